// fb_engine.cu -- the CUDA (sm_100a) engine behind include/figbird_b200.h (kernel revision 3).
//
// One CTA owns one work item = one (gap, candidate length) and runs the whole EM chain of that item on
// chip: initialisation from the partial-read pile-ups, then per round pass 1 (weighted votes), soft
// consensus, pass 2 (hard placement against the consensus), hard consensus / coverage, comp_count and the
// M-step, with the row tables resident in shared memory.  Reference arithmetic:
//   placeReads          Figbird.cpp:3022-4387     computeProbsGap / computeErrorProbsGap  :2090-2137
//   computeSequence     Figbird.cpp:4417-4508     update_partial_prob (initial gap rows)   :1913-2088
//
// What bounds the path and how the kernel is shaped by it.  A pass-1 base term needs the row
// entry {P[r][c], E[r][c]} (16 B) and the shared-memory crossbar moves 128 B/clk/SM, i.e. at most 8 terms per
// clock and SM -- a quarter of what the FP64 pipe could multiply.  So the kernel spends crossbar bytes only on
// terms that need them:
//   * Flank rows never change (one-hot), so the product of a read's bases that land on the flanks depends
//     only on (read, number of bases on that flank): fb_flank_kernel computes these once per gap batch
//     (LF[q][a], RF[q][b]) and every candidate length and every EM round reuses them.
//   * The gap-row terms of all placements of a read form the rectangle rows x read bases; placement x0 owns
//     the diagonal row - j = x0.  A lane walks a diagonal class xi = x0 mod Lg with warp-uniform read base j
//     over a table that is stored cyclically extended, so neighbouring lanes read neighbouring 16-byte
//     entries (conflict-free LDS.128) and a lane that leaves the gap on the right continues with the
//     placement that enters on the left: no lane idles on flank bases.  Per term: one LDS.128 of {P, E-P},
//     one DFMA (P + e*(E-P)), one DMUL.
//   * Only placements inside the admissible insert-size band of a read are stored or visited (the band is
//     an interval in x0, computed once per item).
//   * Pass 2 is exact (products of table entries in read order) but every factor is <= 1, so running
//     products only fall: each read first evaluates the offset that won pass 1, and every other offset is
//     dropped as soon as its running product is below that value.  Results are identical to the unpruned
//     scan (first maximum, ties to the lowest offset).
//
// Numerics.  Pass-2 products, positions, accept tests, votes, consensus and coverage are bit-exact with the
// CPU (separately rounded IEEE operations in the reference's order; the accept test -log10(p) < cutoff is
// p >= accept_min_p with a threshold found by bisection on glibc's log10 at model upload).  Pass-1 products
// are re-associated (flank product x gap product) and use P + e*(E-P): they agree with the reference to
// ~1e-14 relative, far inside the 1e-5 bar on the base weights; every placement is still a fixed function
// of its local inputs and the vote reduction is a fixed-order gather, so results are deterministic and
// candidate lengths that see identical inputs get identical likelihoods, as in the reference.
// There is no CPU path in this file: without a device fb_ctx_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/figbird_b200.h"

namespace {

#ifndef FB_THREADS
#define FB_THREADS 384
#endif
#ifndef FB_CTAS
#define FB_CTAS 2
#endif
#ifndef FB_CHECK
#define FB_CHECK 0      // 1: device-side bounds checks on the staged regions (diagnostic build; compute-sanitizer is not available on the pool)
#endif
#if FB_CHECK
#define FBCHK(cond, what) do { if (!(cond)) { printf("FB_CHECK failed: %s (block %d thread %d)\n", what, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define FBCHK(cond, what) do { } while (0)
#endif
#ifndef FB_PHASES
#define FB_PHASES 0     // 1: thread 0 accumulates the clock cycles between the barriers of an EM round (diagnostic build)
#endif
#ifndef FB_FUSE1
#define FB_FUSE1 0      // 1: the warp that completes a read's last pass-1 unit finishes the read inside the unit loop
#endif
#ifndef FB_FUSE2
#define FB_FUSE2 0      // same for pass 2
#endif
#ifndef FB_PREF
#define FB_PREF 1       // 1: integer mismatch prefilter in front of the pass-2 products
#endif
#ifndef FB_THREADS_BIG
#define FB_THREADS_BIG 512
#endif
constexpr int kThreads = FB_THREADS;
constexpr int kThreadsBig = FB_THREADS_BIG;
constexpr int kMaxReadLen = 255;          // read lengths travel in 8 bits (RMeta::packed); one limit for the model and the batch
constexpr int kBigBucket = 3;             // first shared-memory bucket that leaves room for one CTA per SM only
constexpr int kWarps = kThreads / 32;
constexpr int kMaxSmem = 200 * 1024;      // dynamic shared memory we opt in to (227 KB is the sm_100 limit)
constexpr int kChunkWant = 72 * 1024;     // weights + read-code staging we ask for per CTA when the reads allow it
constexpr int kNumBuckets = 4;
__host__ __device__ constexpr int bucketCap(int b) { return b == 0 ? 36 * 1024 : b == 1 ? 72 * 1024 : b == 2 ? 112 * 1024 : kMaxSmem; }

struct DevGap { long long gap_start; int mode, orig_len, n_reads, read_begin, flank_len, flank_begin, pile_len, pile_begin, flank_pk, pad_; };   // flank_pk: first packed word of the left flank

struct DevItem {
    int kind, gap, Lg, max_rounds, flags, comp_in;
    int n_slots, tables_in_smem, max_len, pad_;
    long long counts_in_off, string_in_off;     // byte offsets into the input arena (-1: none)
    long long out_off;                           // byte offset of the FbItemOut header in the output arena
    long long scratch_off;                       // byte offset of this item's global table scratch (-1: tables in smem)
    long long meta_off;                          // byte offset of this item's per-read metadata in the meta scratch
    long long off_p1, off_p2, off_pos, off_soft, off_hard, off_cov, off_counts;   // relative to out_off
};

struct DevModel {
    const double* e; const double* match;   // [k]
    const double* pdf; int n_insert;
    double etp[25];
    int tmin, tmax, max_read_len;
    int prunable;                           // every pass-2 factor is in [0, 1]: running products are monotone
    int mis_a;                              // every pass-2 mismatch factor is <= 2^-mis_a (0: unknown, no integer prefilter)
    double accept_min_p;
};

struct Params {
    DevModel m;
    const DevGap* gaps;
    const int* read_len; const long long* read_off; const int* read_mate; const int* read_gap;
    const unsigned char* read_flags; const unsigned char* read_jlo; const unsigned char* read_jcut;
    // reads and flanks in HBM: 2 bits per base + an N mask, 16 bases per uint2 {codes, mask}; every read / flank starts a word
    const uint2* codes_pk; const long long* read_pk_off; const uint2* flank_pk;
    const int* pile_l; const int* pile_r;
    double* lfrf;                   // per read: LF[len] then RF[len] at 2*read_off (fb_flank_kernel)
    const DevItem* items;
    const unsigned char* in_arena; unsigned char* out_arena; unsigned char* scratch; unsigned char* meta;
    double e_tab[512];              // e[k] at [k], e[max_read_len-1-i] at [256+i]: read with a warp-uniform index through the constant bank (LDC)
    unsigned long long* counters;   // [0] pass-1 placements, [1] pass-2 placements, [2] base terms, [3] pass-1 lane steps, [4] pass-2 lane steps
};

// Shared-memory plan of one item, computed identically on host (sizing, bucketing) and device (carving).
//   table region (shared memory, or a global scratch slice for very long candidates):
//     UT[5][S] double2 {P, E-P}: plane c = read base code, entry r = gap row r mod Lg for r < S = Lg + maxLenPad
//     (cyclic extension, so that a lane's address is linear in the read base index),
//     C[5][Lg] countsGap, NC[5][Lg] new_counts_gap, G[rows] gapString codes,
//     PREV[Lg] previous hard consensus
//   local region (always shared): ME[k] = e, MER[i] = e[modelLen-1-i] (reverse-strand reads walk it forwards), MT2[k] = {1-e-ins-del, e},
//   ETP[25], then per chunk of reads the records, RC (codes), RC2 (code * S: plane offset of the walk) and W.
struct Plan { int S, rows, oUT, oC, oNC, oG, oPREV, tableBytes, oME, oMER, oMT2, oETP, localFixed, maxLenPad, nMax, perRead; };
__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline Plan makePlan(int Lg, int F, int modelLen, int maxLen, int mode, int bandMax) {
    Plan p;
    p.maxLenPad = (maxLen + 15) & ~15;
    p.S = (Lg + p.maxLenPad + 7) & ~7; p.rows = Lg + 2 * F;      // plane stride: a multiple of 8 entries (RC2 stores code * S / 8 in 16 bits)
    int o = 0;
    p.oUT = o; o += 16 * 5 * p.S;
    p.oC = o; o += 8 * 5 * ((Lg + 3) & ~3);
    p.oNC = o; o += 4 * 5 * Lg;
    p.oG = o; o += al16(p.rows);
    p.oPREV = o; o += al16(Lg);
    p.tableBytes = al16(o);
    if (p.tableBytes < 12 * kThreadsBig) p.tableBytes = al16(12 * kThreadsBig);     // the prologue's prefix-sum scratch lives here
    o = 0;
    p.oME = o; o += al16(8 * modelLen);
    p.oMER = o; o += al16(8 * modelLen);
    p.oMT2 = o; o += 16 * modelLen;
    p.oETP = o; o += 208;
    p.localFixed = al16(o);
    int n = (mode == FB_MODE_UNMAPPED) ? (maxLen + Lg - 1) : (maxLen - 1);
    if (mode == FB_MODE_UNMAPPED && bandMax < n) n = bandMax;     // unmapped reads always pass through the insert-size filter
    p.nMax = n > 1 ? n : 1;
    p.perRead = 8 * p.nMax + 3 * p.maxLenPad + 56;   // weights + codes (bytes and plane offsets) + record + threshold + unit counters
    return p;
}

// Per-read metadata of one item (global scratch, written in the prologue): R+1 entries each.
struct Meta { int* xlo; int* nn; int* woff; int* u1; int* u2; int* x1; double* thr; };
__host__ __device__ inline size_t metaBytes(int R) { return (size_t)al16(4 * (R + 1)) * 6 + (size_t)al16(8 * (R + 1)); }
__device__ __forceinline__ Meta carveMeta(unsigned char* base, int R) {
    Meta m; const size_t s = (size_t)al16(4 * (R + 1));
    m.xlo = (int*)base; m.woff = (int*)(base + s); m.u1 = (int*)(base + 2 * s); m.u2 = (int*)(base + 3 * s); m.x1 = (int*)(base + 4 * s);
    m.nn = (int*)(base + 5 * s); m.thr = (double*)(base + 6 * s);
    return m;
}

__device__ __forceinline__ void window(const DevGap& g, int fl, int len, int Lg, int* lo, int* hi) {
    if (g.mode == FB_MODE_UNMAPPED) { *lo = -(len - 1); *hi = Lg - 1; }
    else if (fl & FB_READ_LEFT) { *lo = -(len - 1); *hi = -1; }
    else { *lo = Lg - len + 1; *hi = Lg - 1; }
}

// Admissible offsets of a read: window intersected with the insert-size filter (Figbird.cpp:3126-3135,3192-3203,
// 3546-3554,3617-3624; finalize :5295).  The implied insert is linear in x0, so the set is an interval.
__device__ __forceinline__ void band(const DevModel& m, const DevGap& g, int fl, int rel, int len, int Lg, bool finalizeRef, int* xlo, int* xhi) {
    int lo, hi; window(g, fl, len, Lg, &lo, &hi);
    const long long off = (long long)Lg - g.orig_len;
    long long a = LLONG_MIN / 4, b = LLONG_MAX / 4;      // filter interval
    if (g.mode == FB_MODE_UNMAPPED) {
        if (fl & FB_READ_LEFT) { a = (long long)m.tmin + rel - len; b = (long long)m.tmax + rel - len; }          // t = x0 - rel + len
        else { const long long A = (long long)rel + off + len; a = A - m.tmax; b = A - m.tmin; }                  // t = rel + off + len - x0
    } else if (fl & FB_READ_LEFT) {
        if (!(fl & FB_READ_NOMATE)) { a = (long long)m.tmin + rel - len; b = (long long)m.tmax + rel - len; }
    } else {
        bool apply = true; long long absref = 0;
        if (fl & FB_READ_NOMATE) { if (!finalizeRef) apply = false; else absref = -1 + off; }
        else absref = (long long)rel + g.gap_start + off;
        if (apply && absref != -1) { const long long A = (absref - g.gap_start) + len; a = A - m.tmax; b = A - m.tmin; }   // t = A - x0
    }
    const long long l2 = a > lo ? a : lo, h2 = b < hi ? b : hi;
    if (l2 > h2) { *xlo = lo; *xhi = lo - 1; } else { *xlo = (int)l2; *xhi = (int)h2; }
}

// Pass-1 weight rows are stored in four planes by (placement index mod 4): the gather's lanes own 4 consecutive gap rows each,
// so at any read base they need indices 4*lane + const -- one plane, consecutive entries, no bank conflicts.
#ifndef FB_FIN
#define FB_FIN 1      // placements per lane and trip in the finish phase (independent log / exp chains)
#endif
#ifndef FB_WTR
#define FB_WTR 0
#endif
#ifndef FB_WTR_BIG
#define FB_WTR_BIG 0   // the same layout for the one-CTA-per-SM launch shape only (shared-memory wavefronts are its busiest unit)
#endif
#ifndef FB_TILEBAR
#define FB_TILEBAR 0   // 1: the parts of one gather tile hand the count rows on through a named barrier of their own (the warps of that tile only)
#endif
#ifndef FB_HINT
#define FB_HINT 0      // 1: a warp looks for the owner of its next work unit next to the owner of its previous one before it bisects
#endif

// base code i of a packed sequence (A0 C1 G2 T3, 4 = N / other)
__device__ __forceinline__ int pkCode(const uint2* p, int i) { const uint2 w = p[i >> 4]; const int sft = i & 15; return ((w.y >> sft) & 1u) ? 4 : (int)((w.x >> (2 * sft)) & 3u); }

// eight consecutive code bytes from an arbitrarily aligned address (three aligned words + byte permutes)
__device__ __forceinline__ void load8(const unsigned char* p, unsigned& lo, unsigned& hi) {
    const unsigned sh = (unsigned)((size_t)p & 3);
    const unsigned* w = (const unsigned*)(p - sh);
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
    const unsigned sel = 0x3210u + 0x1111u * sh;
    lo = __byte_perm(w0, w1, sel); hi = __byte_perm(w1, w2, sel);
}
__device__ __forceinline__ int mismatches8(unsigned alo, unsigned ahi, unsigned blo, unsigned bhi) {
    return __popc(__vcmpne4(alo, blo) & 0x01010101u) + __popc(__vcmpne4(ahi, bhi) & 0x01010101u);
}

// E[j] = sum_{k<4, k!=j} P[k]*ETP[k][j] in k order (Figbird.cpp:2118-2137)
__device__ __forceinline__ void errRow(const double* etp, const double p[4], double e[5]) {
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (j == k) continue; s = __dadd_rn(s, __dmul_rn(p[k], etp[k * 5 + j])); }
        e[j] = s;
    }
}

// soft weight of one placement (Figbird.cpp:3591,3601 unmapped; :3169,3179 partial)
__device__ __forceinline__ double placementWeight(double p, bool unm) {
    if (unm) { const double s = log10(p); return exp(0.5 * s); }
    // partial: w = pow(10, ln p) = exp(ln p * ln 10), the product formed in double-double so that the exponent
    // argument carries no rounding of its own.  These weights run down into the subnormal range and the
    // reference's consensus still sees them (a row whose only vote is 4.9e-324 gets that base), so gradual
    // underflow must round like glibc's pow: evaluate 300 decades higher (exact shift of the exponent argument)
    // and let one IEEE multiply do the final scaling.
    double s = log(p);
    const bool tiny = s < -290.0;
    s = tiny ? s + 300.0 : s;
    const double LN10_HI = 2.302585092994045901e+00, LN10_LO = -2.170756223382249351e-16;
    const double ph = __dmul_rn(s, LN10_HI);
    const double pl = __fma_rn(s, LN10_HI, -ph) + s * LN10_LO;
    double w = exp(ph);
    w = __fma_rn(w, pl, w);
    return __dmul_rn(w, tiny ? 1e-300 : 1.0);      // (x * 1.0 is exact: one code path, no divergence)
}

// ---- flank products, once per gap batch.  One block per read.
//   LF[a] = prod over scored bases j < a            of term(left flank base  lf[F - a + j], read base j)   (a bases on the left flank)
//   RF[b] = prod over scored bases j >= len - b     of term(right flank base rf[j - (len - b)], read base j) (b bases on the right flank)
// term for a one-hot (or all-N) row: P[c] + e[k]*(E[c] - P[c]) with the dense-row values of initialize + computeProbsGap.
__global__ void __launch_bounds__(128) fb_flank_kernel(const Params prm, int nReads) {
    __shared__ double sP[25], sD[25];
    const int q = blockIdx.x;
    if (q >= nReads) return;
    const int tid = threadIdx.x;
    if (tid < 25) {
        const int f = tid / 5, c = tid % 5;
        double p[4], e[5];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = (f == 4) ? 0.25 : (f == k ? 1.0 : 0.0);
        errRow(prm.m.etp, p, e);
        const double P = (c < 4) ? p[c] : 0.0;
        sP[tid] = P; sD[tid] = e[c] - P;
    }
    __syncthreads();
    const int gi = prm.read_gap[q];
    if (gi < 0) return;
    const DevGap g = prm.gaps[gi];
    const int len = prm.read_len[q], fl = prm.read_flags[q];
    const int jlo = prm.read_jlo[q], jhi = len - prm.read_jcut[q];
    const int F = g.flank_len;
    const uint2* rc = prm.codes_pk + prm.read_pk_off[q];
    const uint2* lf = prm.flank_pk + g.flank_pk;
    const uint2* rf = lf + ((F + 15) >> 4);
    const bool rev = fl & FB_READ_REVERSE;
    double* LF = prm.lfrf + 2 * prm.read_off[q];
    double* RF = LF + len;
    for (int a = tid; a < len; a += blockDim.x) {
        double pl = 1.0, pr = 1.0;
        const int je = jhi < a ? jhi : a;
        for (int j = jlo; j < je; j++) {
            const int f = pkCode(lf, F - a + j), c = pkCode(rc, j);
            pl *= __fma_rn(prm.m.e[rev ? (len - 1 - j) : j], sD[f * 5 + c], sP[f * 5 + c]);
        }
        const int js = jlo > len - a ? jlo : len - a;
        for (int j = js; j < jhi; j++) {
            const int f = pkCode(rf, j - (len - a)), c = pkCode(rc, j);
            pr *= __fma_rn(prm.m.e[rev ? (len - 1 - j) : j], sD[f * 5 + c], sP[f * 5 + c]);
        }
        LF[a] = pl; RF[a] = pr;
    }
}

// NT threads per CTA: kThreads (2 CTAs per SM) for items whose shared memory lets two CTAs share an SM, kThreadsBig for the buckets
// that leave room for one CTA only (long candidates: the same number of warps per SM then works on one item).
template <bool TSMEM, int NT>
__global__ void __launch_bounds__(NT, (TSMEM && NT == FB_THREADS) ? FB_CTAS : 1) fb_em_kernel(const Params prm, const int* __restrict__ order, const int smemBytes) {
    constexpr int kThreads = NT, kWarps = NT / 32;
    constexpr bool kWtr = FB_WTR || (FB_WTR_BIG && NT == FB_THREADS_BIG && NT != FB_THREADS);
    auto wtr = [](int i, int np) { return kWtr ? (i & 3) * np + (i >> 2) : i; };
    const DevItem it = prm.items[order[blockIdx.x]];
    const DevGap g = prm.gaps[it.gap];
    const DevModel& m = prm.m;
    const int Lg = it.Lg, F = g.flank_len, R = g.n_reads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool unm = (g.mode == FB_MODE_UNMAPPED);
#if FB_PHASES
    long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long phLast = clock64();
#define PHASE(k) do { if (tid == 0) { const long long t_ = clock64(); ph[k] += t_ - phLast; phLast = t_; } } while (0)
#else
#define PHASE(k) do { } while (0)
#endif
    const Plan pl = makePlan(Lg, F, m.max_read_len, it.max_len, g.mode, m.tmax - m.tmin + 1);
    const int S = pl.S, rows = pl.rows, mlp = pl.maxLenPad;

    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_same[2], s_gchg, s_flags, s_q1, s_next1, s_next2, s_tw[kWarps], s_task[kWarps + 1];
    __shared__ unsigned long long s_lane1, s_lane2, s_terms, s_nsum;

    unsigned char* const tbase = TSMEM ? smem : (prm.scratch + it.scratch_off);
    unsigned char* const lbase = TSMEM ? smem + pl.tableBytes : smem;
    double2* const UT = (double2*)(tbase + pl.oUT);
    double* const C = (double*)(tbase + pl.oC);
    const int LgC = (Lg + 3) & ~3;      // plane stride of C: rows x..x+3 of a gather thread are two aligned 16-byte pairs
    int* const NC = (int*)(tbase + pl.oNC);
    unsigned char* const G = tbase + pl.oG;
    unsigned char* const PREV = tbase + pl.oPREV;
    double* const ME = (double*)(lbase + pl.oME);
    double* const MER = (double*)(lbase + pl.oMER);
    double2* const MT2 = (double2*)(lbase + pl.oMT2);
    double* const ETP = (double*)(lbase + pl.oETP);
    unsigned char* const chunkBase = lbase + pl.localFixed;
    const int chunkBytes = smemBytes - (TSMEM ? pl.tableBytes : 0) - pl.localFixed;
    const Meta mt = carveMeta(prm.meta + it.meta_off, R);

    unsigned char* out = prm.out_arena + it.out_off;
    double* oP1 = (double*)(out + it.off_p1);
    double* oP2 = (double*)(out + it.off_p2);
    int* oPos = (int*)(out + it.off_pos);
    unsigned char* oSoft = out + it.off_soft;
    unsigned char* oHard = out + it.off_hard;
    int* oCov = (int*)(out + it.off_cov);

    const uint2* lf = prm.flank_pk + g.flank_pk;
    const uint2* rf = lf + ((F + 15) >> 4);

    if (tid == 0) { s_same[0] = 1; s_same[1] = 1; s_gchg = 0; s_next1 = 0; s_next2 = 0; s_flags = 0; s_lane1 = 0; s_lane2 = 0; s_terms = 0; s_nsum = 0; }
    int comp = 0;          // comp_count (Figbird.cpp:3919-3927), kept identically by every thread
    bool gChanged = true;  // the gap string differs from the one the cached pass-2 thresholds were computed on
    // ---- model tables, flank part of the gap string
    for (int k = tid; k < m.max_read_len; k += kThreads) { const double e = m.e[k]; ME[k] = e; MER[m.max_read_len - 1 - k] = e; MT2[k] = make_double2(m.match[k], e); }
    if (tid < 25) ETP[tid] = m.etp[tid];
    for (int r = tid; r < rows; r += kThreads) {
        const int x = r - F;
        if (x < 0 || x >= Lg) G[r] = (unsigned char)((r < F) ? pkCode(lf, r) : pkCode(rf, r - F - Lg));     // gap part is written by the consensus / string_in
    }
    __syncthreads();
    // ---- per-read admissible band, prefix sums of band widths (W rows) and of pass-1 / pass-2 work units
    const bool finRef = (it.kind == FB_ITEM_HARD) && (it.flags & FB_FLAG_FINALIZE_REF);
    long long sumN = 0, sumTerms = 0;     // algorithmic placements / base terms of one placeReads call
    {
        const int per = max(1, (R + kThreads - 1) / kThreads);
        const int qb = min(R, tid * per), qe = min(R, qb + per);
        int sw = 0, s1 = 0, s2 = 0; long long st = 0;
        for (int q = qb; q < qe; q++) {
            const int qi = g.read_begin + q;
            const int len = prm.read_len[qi];
            int xlo, xhi; band(m, g, prm.read_flags[qi], prm.read_mate[qi], len, Lg, finRef, &xlo, &xhi);
            const int n = xhi - xlo + 1;
            mt.xlo[q] = xlo; mt.nn[q] = n;
            sw += (n + 3) & ~3; s1 += (Lg > 0) ? ((n < Lg ? n : Lg) + 31) / 32 : 0; s2 += (n + 63) / 64;
            st += (long long)n * (len - prm.read_jcut[qi] - prm.read_jlo[qi]);
        }
        int* const s_scan = (int*)UT;          // [3][kThreads] scratch: the row tables are written after this block
        s_scan[tid] = sw; s_scan[kThreads + tid] = s1; s_scan[2 * kThreads + tid] = s2;
        if (st) atomicAdd(&s_terms, (unsigned long long)st);
        if (qe > qb) { long long ns = 0; for (int q = qb; q < qe; q++) ns += mt.nn[q]; atomicAdd(&s_nsum, (unsigned long long)ns); }
        __syncthreads();
        if (tid < 3) { int acc = 0; for (int i = 0; i < kThreads; i++) { const int v = s_scan[tid * kThreads + i]; s_scan[tid * kThreads + i] = acc; acc += v; } }
        __syncthreads();
        sw = s_scan[tid]; s1 = s_scan[kThreads + tid]; s2 = s_scan[2 * kThreads + tid];
        for (int q = qb; q < qe; q++) {
            const int n = mt.nn[q];
            mt.woff[q] = sw; mt.u1[q] = s1; mt.u2[q] = s2;
            sw += (n + 3) & ~3; s1 += (Lg > 0) ? ((n < Lg ? n : Lg) + 31) / 32 : 0; s2 += (n + 63) / 64;
        }
        // entry R of each prefix array: written by the thread that owns the last read (thread 0 when R == 0)
        const int lastOwner = (R > 0) ? (R - 1) / per : 0;
        if (tid == lastOwner) { mt.woff[R] = sw; mt.u1[R] = s1; mt.u2[R] = s2; mt.nn[R] = 0; }
        __syncthreads();
        sumN = (long long)s_nsum;
        sumTerms = (long long)s_terms;
    }
    const int totalW = mt.woff[R];
    const bool singleChunk = (8LL * totalW + (long long)R * (3 * mlp + 56) + 96) <= (long long)chunkBytes;
    int prevValid = 0;   // previous hard consensus present (uniform)

    auto storeRow = [&](int x, const double p[4], const double e[5]) {     // gap row x and its cyclic copies
        for (int r = x; r < S; r += Lg) {
#pragma unroll
            for (int k = 0; k < 4; k++) UT[k * S + r] = make_double2(p[k], __dadd_rn(e[k], -p[k]));
            UT[4 * S + r] = make_double2(0.0, e[4]);     // read base N: term = e*E[4]
        }
    };

    if (it.kind == FB_ITEM_EM) {
        if (it.flags & FB_FLAG_RESUME) {
            const double* cin = (const double*)(prm.in_arena + it.counts_in_off);
            for (int i = tid; i < 5 * Lg; i += kThreads) { int x = i / 5, k = i % 5; C[k * LgC + x] = cin[i]; }
            if (it.string_in_off >= 0) { const unsigned char* sin = prm.in_arena + it.string_in_off; for (int x = tid; x < Lg; x += kThreads) PREV[x] = sin[x]; prevValid = 1; }
            comp = it.comp_in;
        } else {
            // previous_str carried in from an earlier run() of the same length (Figbird.cpp:3919-3927): seeds comp_count's comparison
            if (it.string_in_off >= 0) { const unsigned char* sin = prm.in_arena + it.string_in_off; for (int x = tid; x < Lg; x += kThreads) PREV[x] = sin[x]; prevValid = 1; }
            // gap rows from the partial pile-ups (update_partial_prob, Figbird.cpp:2039-2081)
            const int* plp = prm.pile_l + 4 * (size_t)g.pile_begin; const int* prp = prm.pile_r + 4 * (size_t)g.pile_begin;
            for (int x = tid; x < Lg; x += kThreads) {
                double cnt[4]; int tot = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int c = 1;
                    if (x < g.pile_len) c += plp[4 * x + k];
                    int u = Lg - 1 - x; if (u < g.pile_len) c += prp[4 * u + k];
                    cnt[k] = (double)c; tot += c;
                }
                double p[4], e[5];
#pragma unroll
                for (int k = 0; k < 4; k++) p[k] = __ddiv_rn(cnt[k], (double)tot);
                errRow(m.etp, p, e);
                storeRow(x, p, e);
            }
        }
    } else {
        const unsigned char* sin = prm.in_arena + it.string_in_off;
        for (int x = tid; x < Lg; x += kThreads) { unsigned char c = it.string_in_off >= 0 ? sin[x] : 4; G[F + x] = c; oSoft[x] = c; oHard[x] = 4; oCov[x] = 0; }
        for (int q = tid; q < R; q += kThreads) { oP1[q] = -1.0; mt.x1[q] = INT_MIN; }
    }

    // ---- chunk region: per-read records RM[nq+1], pass-2 thresholds THR[nq], read codes RC[nq][mlp], weights W
    struct RMeta { int xlo, wrel, n, packed, rel, u1, u2, x1; };     // packed = len | jlo<<8 | jhi<<16 | flags<<24; wrel/u1/u2 relative to the chunk
    RMeta* const RM = (RMeta*)chunkBase;
    int nqCur = 0;                       // reads in the resident chunk (uniform)
    double* THR = nullptr; int* CNT = nullptr; int* X1P = nullptr; unsigned char* RC = nullptr; unsigned short* RC2 = nullptr; double* W = nullptr;
    auto carveChunk = [&](int nq) {
        nqCur = nq;
        THR = (double*)(chunkBase + 32 * (size_t)(nq + 1));      // exact pass-2 product at X1P (pruning threshold)
        CNT = (int*)((unsigned char*)THR + al16(8 * nq));        // [2][nq] finished pass-1 / pass-2 units of a read
        X1P = CNT + 2 * nq;                                      // offset THR was computed at
        RC = (unsigned char*)CNT + al16(12 * nq);
        RC2 = (unsigned short*)(RC + (size_t)nq * mlp);         // [nq][mlp] code * S / 8: plane offset of the walk in units of 8 entries
        W = (double*)(RC + (size_t)nq * 3 * mlp);
        FBCHK((unsigned char*)W <= chunkBase + chunkBytes, "chunk records + codes exceed the chunk region");
    };
    auto stageReads = [&](int q0, int q1) {   // read records and codes of the chunk
        const int nq = q1 - q0;
        carveChunk(nq);
        const int wb = mt.woff[q0], u1b = mt.u1[q0], u2b = mt.u2[q0];
        FBCHK((unsigned char*)(W + (mt.woff[q1] - wb)) <= chunkBase + chunkBytes, "weight rows exceed the chunk region");
        for (int ql = tid; ql <= nq; ql += kThreads) {
            const int q = q0 + ql;
            RMeta r;
            r.wrel = mt.woff[q] - wb; r.u1 = mt.u1[q] - u1b; r.u2 = mt.u2[q] - u2b;
            if (ql < nq) {
                const int qi = g.read_begin + q; const int len = prm.read_len[qi];
                r.xlo = mt.xlo[q]; r.n = mt.nn[q]; r.rel = prm.read_mate[qi]; r.x1 = mt.x1[q];
                r.packed = len | (prm.read_jlo[qi] << 8) | ((len - prm.read_jcut[qi]) << 16) | (prm.read_flags[qi] << 24);
            } else { r.xlo = 0; r.n = 0; r.rel = 0; r.x1 = INT_MIN; r.packed = 0; }
            RM[ql] = r;
            if (ql < nq) { CNT[ql] = 0; CNT[nq + ql] = 0; X1P[ql] = INT_MIN; THR[ql] = 0.0; }
        }
        // one packed word (16 bases: 8-byte load, consecutive threads -> consecutive words) per thread and trip
        const int wpr = mlp >> 4, nw = nq * wpr;
        const unsigned s8 = (unsigned)(S >> 3);
        for (int i = tid; i < nw; i += kThreads) {
            const int ql = i / wpr, wj = i - ql * wpr;
            const int qi = g.read_begin + q0 + ql;
            const int len = prm.read_len[qi];
            uint2 w = make_uint2(0u, 0xffffu);
            if (16 * wj < len) w = prm.codes_pk[prm.read_pk_off[qi] + wj];
            unsigned cb[4] = {0, 0, 0, 0}, co[8];
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const unsigned c = (16 * wj + t >= len || ((w.y >> t) & 1u)) ? 4u : ((w.x >> (2 * t)) & 3u);
                cb[t >> 2] |= c << (8 * (t & 3));
                if (t & 1) co[t >> 1] |= (c * s8) << 16; else co[t >> 1] = c * s8;
            }
            uint4* rcp = (uint4*)(RC + (size_t)ql * mlp + 16 * wj);
            *rcp = make_uint4(cb[0], cb[1], cb[2], cb[3]);
            uint4* r2p = (uint4*)(RC2 + (size_t)ql * mlp + 16 * wj);
            r2p[0] = make_uint4(co[0], co[1], co[2], co[3]); r2p[1] = make_uint4(co[4], co[5], co[6], co[7]);
        }
    };
    // reads [q0, q1) whose records + codes + weight rows fit the chunk region (at least one read)
    auto chunkEnd = [&](int q0) -> int {
        if (singleChunk) return R;
        __syncthreads();
        if (tid == 0) {
            int q = q0; long long bytes = 96;
            while (q < R) { const long long nb = 8LL * (mt.woff[q + 1] - mt.woff[q]) + 3 * mlp + 56; if (q > q0 && bytes + nb > chunkBytes) break; bytes += nb; q++; }
            s_q1 = q;
        }
        __syncthreads();
        return s_q1;
    };
    // read of the resident chunk that owns work unit `target` (largest ql with prefix[ql] <= target); U2 selects the pass-2 prefix
    // (units are handed out in increasing order, so the owner is usually the read of the warp's previous unit or one of the next
    //  few: `hint` = that read; RM[nqCur] holds the total, which no unit reaches)
    auto findRead = [&](bool U2, int target, int hint) -> int {
        int lo = 0, hi = nqCur - 1;
#if FB_HINT
        if ((U2 ? RM[hint].u2 : RM[hint].u1) <= target) {
            lo = hint;
#pragma unroll 1
            for (int s = 0; s < 3; s++) { if ((U2 ? RM[lo + 1].u2 : RM[lo + 1].u1) > target) return lo; lo++; }
        } else hi = hint - 1;
#endif
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; const int v = U2 ? RM[mid].u2 : RM[mid].u1; if (v <= target) lo = mid; else hi = mid - 1; }
        return lo;
    };
    if (it.kind == FB_ITEM_EM) for (int q = tid; q < R; q += kThreads) mt.x1[q] = INT_MIN;
    __syncthreads();
    if (singleChunk) stageReads(0, R);
    __syncthreads();

    // M-step over the gap rows (flank rows never change): computeProbsGap(0) + computeErrorProbsGap
    auto mstepRow = [&](int x) {
        {
            double c[5];
#pragma unroll
            for (int k = 0; k < 5; k++) c[k] = C[k * LgC + x];
            double total = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) total = __dadd_rn(total, c[k]);
            const double nq = __ddiv_rn(c[4], 4.0);
            double p[4], e[5];
#pragma unroll
            for (int k = 0; k < 4; k++) p[k] = (total != 0.0) ? __ddiv_rn(__dadd_rn(c[k], nq), total) : 0.25;
            errRow(m.etp, p, e);
            storeRow(x, p, e);
        }
    };
    if (it.kind == FB_ITEM_EM && (it.flags & FB_FLAG_RESUME)) { for (int x = tid; x < Lg; x += kThreads) mstepRow(x); __syncthreads(); }

    // ---- pass 2: products of exact table entries against the gap string, first maximum, accept, unit votes
    auto pass2 = [&](int slot, bool vote) {
        for (int q0 = 0, q1; q0 < R; q0 = q1) {
            q1 = chunkEnd(q0);
            if (!singleChunk) { stageReads(q0, q1); __syncthreads(); }
            const int nq = q1 - q0;
            // exact product at the offset that won pass 1: every other offset only has to beat it.  EM items additionally
            // stop at the accept threshold: a read whose best product stays below it is rejected whatever the exact value
            // (FbItemOut: p2max / pos2 of rejected reads are unspecified below the threshold); HARD items stay exact.
            const double floorP = (vote && m.prunable && m.accept_min_p < 1e300) ? m.accept_min_p : 0.0;
            if (tid == 0) s_next2 = 0;
            // the threshold of the previous round stays exact when neither the gap string nor the seed offset changed
            const bool reuseThr = singleChunk && !gChanged;
            for (int ql = tid; ql < nq; ql += kThreads) {
                const RMeta r = RM[ql];
                if (reuseThr && X1P[ql] == r.x1) continue;
                double thr = 0.0;
                if (r.x1 != INT_MIN && m.prunable) {
                    const int len = r.packed & 0xff, jlo = (r.packed >> 8) & 0xff, jhi = (r.packed >> 16) & 0xff;
                    const bool rev = (r.packed >> 24) & FB_READ_REVERSE;
                    const unsigned char* rc = RC + ql * mlp;
                    const unsigned char* gs = G + F + r.x1;
                    double p = 1.0;
                    int j = jlo;
                    for (; j + 4 <= jhi; j += 4) {      // factors first (independent loads), then the ordered product
                        double v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int c = rc[j + u], f = gs[j + u];
                            const double2 mk = MT2[rev ? (len - j - u - 1) : (j + u)];
                            v[u] = (f == c) ? mk.x : __dmul_rn(mk.y, ETP[f * 5 + c]);
                        }
                        p = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(p, v[0]), v[1]), v[2]), v[3]);
                    }
                    for (; j < jhi; j++) {
                        const int c = rc[j], f = gs[j];
                        const double2 mk = MT2[rev ? (len - j - 1) : j];
                        p = __dmul_rn(p, (f == c) ? mk.x : __dmul_rn(mk.y, ETP[f * 5 + c]));
                    }
                    thr = p;
                }
                THR[ql] = thr; X1P[ql] = r.x1;
            }
            __syncthreads();
            PHASE(6);
            // first maximum over ascending offsets (strict >), accept test, unit votes (Figbird.cpp:3787-3912): by one warp
            auto finishRead2 = [&](int ql) {
                const RMeta r = RM[ql];
                const int q = q0 + ql, len = r.packed & 0xff;
                const int xlo = r.xlo, n = r.n;
                const double* Wq = W + r.wrel;
                double best = -1.0; int bestI = 0x7fffffff;
                for (int i = lane; i < n; i += 32) { double v = Wq[i]; if (v > best) { best = v; bestI = i; } }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bestI, o);
                    if (ov > best || (ov == best && oi < bestI)) { best = ov; bestI = oi; }
                }
                const int bestx = (best >= 0) ? xlo + bestI : 0;
                if (lane == 0) { oP2[(size_t)slot * R + q] = best; oPos[(size_t)slot * R + q] = bestx; }
                if (vote && unm && best >= m.accept_min_p) {
                    const unsigned char* rc = RC + ql * mlp;
                    for (int j = lane; j < len; j += 32) { int x = bestx + j; if (x >= 0 && x < Lg) atomicAdd(&NC[rc[j] * Lg + x], 1); }
                    if (lane == 0 && g.orig_len <= 30) {
                        int fo = 0; const int p0 = bestx, val = p0 + len - Lg;
                        if (p0 < 0 && val > 0 && -p0 > 3 && val > 3) fo |= 4;
                        if (p0 < 0 && p0 + len > 0 && -p0 > 3) fo |= 1;
                        if (p0 > 0 && p0 < Lg && val > 0 && val > 3) fo |= 2;
                        if (fo) atomicOr(&s_flags, fo);
                    }
                }
            };
            const int units = RM[nq].u2;
            int hint2 = 0;
            for (;;) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&s_next2, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= units) break;
                const int ql = findRead(true, u, hint2);
                hint2 = ql;
                const RMeta r = RM[ql];
                const int ch = u - r.u2;
                const int len = r.packed & 0xff, jlo = (r.packed >> 8) & 0xff, jhi = (r.packed >> 16) & 0xff;
                const int xlo = r.xlo, n = r.n;
                const double thrX = THR[ql];
                const int x1 = (thrX > 0.0) ? r.x1 : INT_MIN;      // known exactly: not walked again
                const double thr = fmax(thrX, floorP);             // running products below this cannot matter
                const int ia = ch * 64 + lane, ib = ia + 32;
                const int xa = xlo + ia, xb = xlo + ib;
                const bool ina = ia < n, inb = ib < n;
                bool acta = ina && xa != x1, actb = inb && xb != x1;
                double pa = 1.0, pb = 1.0;
#if FB_PREF
                // Integer prefilter.  Every factor is <= 1 and a mismatch factor is <= 2^-mis_a, so a placement with M mismatches
                // among its first scored bases has a product <= 2^(-mis_a * M): with thr >= 2^ilogb(thr) it cannot reach thr once
                // mis_a * M > -ilogb(thr).  Such placements are dropped before any FP64 work (they could only lose).
                if (m.prunable && m.mis_a > 0 && thr > 0.0 && jhi - jlo >= 8) {
                    const int Mthr = max(1, (-ilogb(thr)) / m.mis_a + 1);
                    const unsigned char* rq = RC + ql * mlp + jlo;
                    const unsigned char* qa = G + F + (ina ? xa : xlo) + jlo; const unsigned char* qb = G + F + (inb ? xb : (ina ? xa : xlo)) + jlo;
                    unsigned r0, r1, a0, a1, b0, b1;
                    load8(rq, r0, r1); load8(qa, a0, a1); load8(qb, b0, b1);
                    int ma = mismatches8(a0, a1, r0, r1), mb = mismatches8(b0, b1, r0, r1);
                    if (jhi - jlo >= 16) {
                        load8(rq + 8, r0, r1); load8(qa + 8, a0, a1); load8(qb + 8, b0, b1);
                        ma += mismatches8(a0, a1, r0, r1); mb += mismatches8(b0, b1, r0, r1);
                    }
                    if (ma >= Mthr) { acta = false; pa = 0.0; }
                    if (mb >= Mthr) { actb = false; pb = 0.0; }
                }
#endif
                if (__any_sync(0xffffffffu, acta || actb)) {
                    const int ra = F + (ina ? xa : xlo), rb = F + (inb ? xb : (ina ? xa : xlo));
                    const bool rev = (r.packed >> 24) & FB_READ_REVERSE;
                    const unsigned char* rc = RC + ql * mlp;
                    const unsigned char* ga = G + ra; const unsigned char* gb = G + rb;
                    FBCHK(ra + jlo >= 0 && ra + jhi <= rows && rb + jlo >= 0 && rb + jhi <= rows, "pass 2 leaves the gap string");
                    int j = jlo, steps = 0;
                    while (j < jhi) {
                        const int je = min(jhi, j + 4);
#pragma unroll 4
                        for (; j < je; j++) {
                            const int c = rc[j];
                            const double2 mk = MT2[rev ? (len - j - 1) : j];      // x = 1-e-ins-del, y = e
                            const int fa = ga[j], fb = gb[j];
                            const double va = (fa == c) ? mk.x : __dmul_rn(mk.y, ETP[fa * 5 + c]);
                            const double vb = (fb == c) ? mk.x : __dmul_rn(mk.y, ETP[fb * 5 + c]);
                            pa = __dmul_rn(pa, va); pb = __dmul_rn(pb, vb);
                        }
                        steps += 4;
                        if (!__any_sync(0xffffffffu, (acta && pa >= thr) || (actb && pb >= thr))) break;
                    }
                    if (lane == 0) atomicAdd(&s_lane2, (unsigned long long)steps * 64ull);
                }
                if (ina) W[r.wrel + ia] = (xa == x1) ? thrX : pa;      // the seed offset is known exactly; dropped placements keep 0
                if (inb) W[r.wrel + ib] = (xb == x1) ? thrX : pb;
#if FB_FUSE2
                // the warp that completes the last unit of a read finishes the read (its products are all in W by then)
                __threadfence_block();
                int done = 0;
                if (lane == 0) done = atomicAdd(&CNT[nq + ql], 1) + 1;
                done = __shfl_sync(0xffffffffu, done, 0);
                if (done == RM[ql + 1].u2 - r.u2) { __threadfence_block(); if (lane == 0) CNT[nq + ql] = 0; finishRead2(ql); }
#endif
            }
#if FB_FUSE2
            for (int ql = warp; ql < nq; ql += kWarps) if (RM[ql + 1].u2 == RM[ql].u2) finishRead2(ql);      // reads without offsets
#else
            __syncthreads();
            PHASE(7);
            for (int ql = warp; ql < nq; ql += kWarps) finishRead2(ql);
#endif
            __syncthreads();
            PHASE(8);
        }
    };

    int calls = 0;
    if (it.kind == FB_ITEM_HARD) {
        pass2(0, false);
        calls = 1;
    } else {
        const int maxCalls = it.max_rounds + ((it.flags & FB_FLAG_EXTRA_PASS) ? 1 : 0);
        bool emDone = it.max_rounds <= 0;
        const long long offLg = (long long)Lg - g.orig_len;
        // gather geometry: a task = 2 or 4 consecutive gap rows x one part of the reads; the parts add their sums to the
        // count matrix one after the other (fixed order), separated by barriers
        const int rowsPerThread = (Lg > 64) ? 4 : 2;
        const int nG = (Lg + rowsPerThread - 1) / rowsPerThread, tpp = (nG + 31) & ~31;
        // A warp = one tile of 32 * rowsPerThread rows and one part of the reads.  The warps are dealt to the tiles in
        // proportion to the tiles' work (read bases walked), computed once from the admissible bands.
        const int nTiles = tpp >> 5;
        const bool tiled = Lg > 0 && nTiles <= kWarps;
        int split = 1;                 // longest chain of parts (uniform)
        int myTile = 0, myPart = 0, myParts = 1; bool myIdle = false;
        if (tiled) {
            if (warp < nTiles) {
                const int xw0 = warp * 32 * rowsPerThread, xw1 = min(Lg - 1, xw0 + 32 * rowsPerThread - 1);
                int w = 0;
                for (int q = lane; q < R; q += 32) {
                    const int n = mt.nn[q];
                    if (n <= 0) continue;
                    const int xlo = mt.xlo[q], len = prm.read_len[g.read_begin + q];
                    const int ja = max(0, xw0 - xlo - n + 1), jb = min(len - 1, xw1 - xlo);
                    w += max(0, jb - ja + 1);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                if (lane == 0) s_tw[warp] = w + 1;
            }
            __syncthreads();
            if (tid == 0) {
                int parts[kWarps];
                for (int t = 0; t < nTiles; t++) parts[t] = 1;
                const int cap = max(1, R);     // more parts than reads is useless
                for (int used = nTiles; used < kWarps; used++) {
                    int bt = -1;
                    for (int t = 0; t < nTiles; t++) if (parts[t] < cap && (bt < 0 || (long long)s_tw[t] * parts[bt] > (long long)s_tw[bt] * parts[t])) bt = t;
                    if (bt < 0) break;
                    parts[bt]++;
                }
                int w = 0, mp = 1;
                for (int t = 0; t < nTiles; t++) { for (int p = 0; p < parts[t]; p++) s_task[w++] = t | (p << 8) | (parts[t] << 16); mp = max(mp, parts[t]); }
                for (; w < kWarps; w++) s_task[w] = 0xff;
                s_task[kWarps] = mp;
            }
            __syncthreads();
            const int task = s_task[warp];
            myIdle = (task & 0xff) == 0xff; myTile = task & 0xff; myPart = (task >> 8) & 0xff; myParts = max(1, (task >> 16) & 0xff);
            split = s_task[kWarps];
        }
        PHASE(0);
        for (int call = 0; call < maxCalls; call++) {
            const bool extra = emDone;
            const int slot = (it.flags & FB_FLAG_RECORD_ALL) ? call : 0;
            // (no barrier of its own: the walk does not touch C / NC, and a barrier separates it from the gather)
            for (int i = tid; i < 5 * LgC; i += kThreads) C[i] = 0.0;
            for (int i = tid; i < 5 * Lg; i += kThreads) NC[i] = 0;
            if (tid == 0) { s_same[call & 1] = 1; s_gchg = 0; }      // (the other parity may still be read by slow warps of the previous round)
            // ================= pass 1 (Figbird.cpp:3082-3263, 3530-3689) =================
            for (int q0 = 0, q1; q0 < R; q0 = q1) {
                q1 = chunkEnd(q0);
                if (!singleChunk) { stageReads(q0, q1); __syncthreads(); }
                const int nq = q1 - q0;
                // ---- finish a read: insert pdf x left-flank product x gap product x right-flank product for every placement,
                //      per-read maximum (value and offset), soft weight in place.  Run by one warp.
                auto finishRead1 = [&](int ql) {
                    const RMeta r = RM[ql];
                    const int q = q0 + ql, qi = g.read_begin + q;
                    const int len = r.packed & 0xff, jlo = (r.packed >> 8) & 0xff, jhi = (r.packed >> 16) & 0xff, fl = (r.packed >> 24) & 0xff;
                    const int xlo = r.xlo, n = r.n, rel = r.rel;
                    double* const Wq = W + r.wrel;
                    const int np = (n + 3) >> 2;
                    const double* LF = prm.lfrf + 2 * prm.read_off[qi];
                    const double* RF = LF + len;
                    double best = 0.0; int bestI = 0x7fffffff;
                    // FB_FIN placements per lane and trip: the log / exp chains of the weight are long and independent
                    for (int i0 = lane; i0 < n; i0 += 32 * FB_FIN) {
                        double p[FB_FIN], w[FB_FIN];
#pragma unroll
                        for (int u = 0; u < FB_FIN; u++) {
                            const int i = i0 + 32 * u;
                            p[u] = 0.0;
                            if (i < n) {
                                const int x0 = xlo + i;
                                double v = 1.0;
                                if (unm) {
                                    const long long t = (fl & FB_READ_LEFT) ? ((long long)x0 - rel + len) : ((long long)rel + offLg + len - x0);
                                    v = m.pdf[min(max((int)t, 0), m.n_insert - 1)];
                                }
                                if (x0 < 0) v = __dmul_rn(v, LF[-x0]);
                                if (min(jhi, Lg - x0) > max(jlo, -x0)) v = __dmul_rn(v, Wq[wtr(i, np)]);
                                const int b = x0 + len - Lg;
                                if (b > 0) v = __dmul_rn(v, RF[b]);
                                p[u] = (v > 0.0) ? v : 0.0;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < FB_FIN; u++) w[u] = placementWeight(p[u] > 0.0 ? p[u] : 1.0, unm);
#pragma unroll
                        for (int u = 0; u < FB_FIN; u++) {
                            const int i = i0 + 32 * u;
                            if (i < n) {
                                Wq[wtr(i, np)] = (p[u] > 0.0) ? w[u] : 0.0;
                                if (p[u] > best) { best = p[u]; bestI = i; }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        double ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bestI, o);
                        if (ov > best || (ov == best && oi < bestI)) { best = ov; bestI = oi; }
                    }
                    if (lane == 0) {
                        const int x1 = (best > 0.0) ? xlo + bestI : INT_MIN;
                        oP1[(size_t)slot * R + q] = (best > 0.0) ? best : -1.0; RM[ql].x1 = x1; mt.x1[q] = x1;
                    }
                };
                // ---- gap-row products of every admissible placement: cyclic diagonal walk
                const int units = (Lg > 0) ? RM[nq].u1 : 0;
                int hint1 = 0;
                for (;;) {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&s_next1, 1);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= units) break;
                    const int ql = findRead(false, u, hint1);
                    hint1 = ql;
                    const RMeta r = RM[ql];
                    const int uu = u - r.u1;
                    const int len = r.packed & 0xff, jlo = (r.packed >> 8) & 0xff, jhi = (r.packed >> 16) & 0xff;
                    const int xlo = r.xlo, n = r.n;
                    double* const Wq = W + r.wrel;
                    const int np = (n + 3) >> 2;
                    const bool full = n >= Lg;
                    const int nl = full ? Lg : n;
                    const int idx = uu * 32 + lane;
                    const bool active = idx < nl;
                    int js, je;
                    if (full) { js = jlo; je = jhi; }
                    else { const int xf = xlo + uu * 32, xl = min(xf + 31, xlo + n - 1); js = max(jlo, -xl); je = min(jhi, Lg - xf); }
                    if (je > js) {      // (a unit whose lanes own no gap row still counts towards its read below)
                    const int xi = active ? (full ? idx : xlo + idx) : (full ? 0 : xlo);
                    int xm = xi % Lg; if (xm < 0) xm += Lg;
                    const int m0 = (xm + js) / Lg;
                    int x0cur = xm - m0 * Lg;          // placement whose segment contains read base js
                    int jw = (m0 + 1) * Lg - xm;       // read base at which the walk re-enters gap row 0 (> js)
                    const double2* ptr = UT + xm;      // + j: cyclically extended table, plane 0
                    FBCHK(xm >= 0 && xm + je <= S && js >= 0 && je <= mlp, "walk leaves the extended table");
                    const unsigned short* rc = RC2 + ql * mlp;
                    // me[j] = e of read base j: kernel-parameter (constant bank) table, or shared memory when reads are longer than it
                    const bool revRead = (r.packed >> 24) & FB_READ_REVERSE;
                    const int eBase = revRead ? 256 + (m.max_read_len - len) : 0;
                    struct ETab { const Params& p; int base; __device__ __forceinline__ double operator[](int j) const { return p.e_tab[base + j]; } };
                    const ETab me{prm, eBase};
                    double acc = 1.0, save = 1.0;
                    const bool oneWrap = (je - js) <= Lg;     // a lane re-enters row 0 at most once: keep the first product in a register
                    auto mul = [&](const double2 v, double e) { acc = __dmul_rn(acc, __fma_rn(e, v.y, v.x)); };
                    auto wrap1 = [&](int j) { const bool w = (j == jw); save = w ? acc : save; acc = w ? 1.0 : acc; };
                    // general case (a lane re-enters row 0 several times, Lg < read length): branch-free, one predicated store.
                    // slotN = x0cur - xlo for active lanes, out of range for inactive ones
                    int slotN = active ? x0cur - xlo : -1;
                    const int nAct = active ? n : 0;
                    auto wrapN = [&](int j) {
                        const bool w = (j == jw);       // the walk re-enters gap row 0: the running product belongs to placement x0cur
                        const bool st = w && ((unsigned)slotN < (unsigned)nAct);
                        FBCHK(!st || wtr(slotN, np) < ((n + 3) & ~3), "flush slot outside the weight row");
                        if (TSMEM) {
                            const unsigned wa = (unsigned)__cvta_generic_to_shared(Wq) + ((unsigned)wtr(slotN, np) << 3);
                            asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.shared.f64 [%1], %2; }" :: "r"((unsigned)st), "r"(wa), "d"(acc) : "memory");
                        } else if (st) Wq[wtr(slotN, np)] = acc;
                        acc = w ? 1.0 : acc; slotN = w ? slotN - Lg : slotN; x0cur = w ? x0cur - Lg : x0cur; jw = w ? jw + Lg : jw;
                    };
                    // read bases [j, stop): MODE 0 = no lane wraps in this stretch, 1 = single-wrap lanes, 2 = general
                    auto run = [&](auto MODE, int& j, int stop) {
                        auto wr = [&](int jj) { if (MODE.value == 1) wrap1(jj); else if (MODE.value == 2) wrapN(jj); };
                        for (; j < stop && (j & 3); j++) { wr(j); mul(ptr[rc[j] * 8 + j], me[j]); }
                        for (; j + 4 <= stop; j += 4) {
                            const uint2 cw = *(const uint2*)(rc + j);        // four plane offsets (units of 8 entries = 128 bytes)
                            const unsigned char* pb = (const unsigned char*)(ptr + j);       // byte address: mask + shift-add per entry
                            const unsigned o0 = __byte_perm(cw.x, 0, 0x4410), o1 = __byte_perm(cw.x, 0, 0x4432), o2 = __byte_perm(cw.y, 0, 0x4410), o3 = __byte_perm(cw.y, 0, 0x4432);
                            const double2 v0 = *(const double2*)(pb + (o0 << 7)), v1 = *(const double2*)(pb + (o1 << 7) + 16),
                                          v2 = *(const double2*)(pb + (o2 << 7) + 32), v3 = *(const double2*)(pb + (o3 << 7) + 48);
                            const double e0 = me[j], e1 = me[j + 1], e2 = me[j + 2], e3 = me[j + 3];
                            wr(j); mul(v0, e0);
                            wr(j + 1); mul(v1, e1);
                            wr(j + 2); mul(v2, e2);
                            wr(j + 3); mul(v3, e3);
                        }
                        for (; j < stop; j++) { wr(j); mul(ptr[rc[j] * 8 + j], me[j]); }
                    };
                    // lanes hold consecutive diagonals, so their wrap points fill a window of at most 32 consecutive read bases per period
                    int wmin = __reduce_min_sync(0xffffffffu, jw), wmax = __reduce_max_sync(0xffffffffu, jw);
                    int j = js;
                    if (oneWrap) {
                        run(std::integral_constant<int, 0>(), j, min(je, wmin));
                        run(std::integral_constant<int, 1>(), j, min(je, wmax + 1));
                        run(std::integral_constant<int, 0>(), j, je);
                        if (active) {
                            const bool wrapped = jw < je;
                            if (wrapped) {
                                if ((unsigned)(x0cur - xlo) < (unsigned)n) Wq[wtr(x0cur - xlo, np)] = save;
                                if ((unsigned)(x0cur - Lg - xlo) < (unsigned)n) Wq[wtr(x0cur - Lg - xlo, np)] = acc;
                            } else if ((unsigned)(x0cur - xlo) < (unsigned)n) Wq[wtr(x0cur - xlo, np)] = acc;
                        }
                    } else {
                        while (j < je) {
                            run(std::integral_constant<int, 0>(), j, min(je, wmin));
                            run(std::integral_constant<int, 2>(), j, min(je, wmax + 1));
                            wmin += Lg; wmax += Lg;
                        }
                        if (active && (unsigned)(x0cur - xlo) < (unsigned)n) Wq[wtr(x0cur - xlo, np)] = acc;
                    }
                    if (lane == 0) atomicAdd(&s_lane1, (unsigned long long)(je - js) * 32ull);
                    }
#if FB_FUSE1
                    // the warp that completes the last unit of a read finishes the read (all its products are in W by then)
                    __threadfence_block();
                    int done = 0;
                    if (lane == 0) done = atomicAdd(&CNT[ql], 1) + 1;
                    done = __shfl_sync(0xffffffffu, done, 0);
                    if (done == RM[ql + 1].u1 - r.u1) { __threadfence_block(); if (lane == 0) CNT[ql] = 0; finishRead1(ql); }
#endif
                }
#if FB_FUSE1
                for (int ql = warp; ql < nq; ql += kWarps) if (Lg <= 0 || RM[ql + 1].u1 == RM[ql].u1) finishRead1(ql);      // reads without gap-row units
#else
                __syncthreads();
                PHASE(1);
                for (int ql = warp; ql < nq; ql += kWarps) finishRead1(ql);
#endif
                __syncthreads();
                PHASE(2);
                if (tid == 0) s_next1 = 0;      // every warp has left the unit loop; the next one starts after further barriers
                // ---- gather the weights of this chunk into the gap rows in a fixed order (deterministic, no FP atomics).
                // A thread owns 4 consecutive rows x..x+3 and one part of the reads; all lanes of a warp walk the same read
                // base j (uniform code -> uniform branch), row x+b takes the weight of placement x+b-j: one new weight
                // per step slides through four registers.
                auto gather = [&](auto BB) {
                    constexpr int B = BB.value;           // gap rows per thread
                    for (int idx = tid; idx < tpp || tiled; idx += kThreads) {
                        // tiled: one trip, every thread reaches the barriers below; else a loop over the row groups, all reads
                        const int s = tiled ? myPart : 0, stride = tiled ? myParts : 1;
                        const int gi = tiled ? myTile * 32 + lane : idx;
                        const int x = gi * B;
                        const int xw0 = (gi - lane) * B, xw1 = min(Lg - 1, xw0 + 32 * B - 1);      // rows of this warp
                        double a[B][5];
#pragma unroll
                        for (int b = 0; b < B; b++)
#pragma unroll
                            for (int k = 0; k < 5; k++) a[b][k] = 0.0;
                        if (xw0 < Lg && !(tiled && myIdle)) for (int ql = s; ql < nq; ql += stride) {
                            const RMeta r = RM[ql];
                            if (r.n <= 0) continue;
                            const int len = r.packed & 0xff;
                            const int xlo = r.xlo, n = r.n;
                            // placement x0 = row - j in [xlo, xlo + n)  <=>  j in [row - xlo - n + 1, row - xlo]
                            const int ja = max(0, xw0 - xlo - n + 1), jb = min(len - 1, xw1 - xlo);
                            if (ja > jb) continue;
                            const unsigned char* rc = RC + ql * mlp;
                            const double* wr = W + r.wrel;
                            const int np = (n + 3) >> 2;
                            FBCHK(r.wrel >= 0 && (unsigned char*)(wr + ((n + 3) & ~3)) <= chunkBase + chunkBytes, "gather row outside the chunk region");
                            auto ld = [&](int i) -> double { return ((unsigned)i < (unsigned)n) ? wr[wtr(i, np)] : 0.0; };
                            int i0 = x - ja - xlo;                 // weight index of row x at read base ja; row x+b: i0 + b
                            double w[B];
#pragma unroll
                            for (int b = 0; b < B; b++) w[b] = ld(i0 + b);
                            auto add = [&](int c, const double (&v)[B]) {
                                switch (c) {
                                    case 0:
#pragma unroll
                                        for (int b = 0; b < B; b++) a[b][0] = __dadd_rn(a[b][0], v[b]);
                                        break;
                                    case 1:
#pragma unroll
                                        for (int b = 0; b < B; b++) a[b][1] = __dadd_rn(a[b][1], v[b]);
                                        break;
                                    case 2:
#pragma unroll
                                        for (int b = 0; b < B; b++) a[b][2] = __dadd_rn(a[b][2], v[b]);
                                        break;
                                    case 3:
#pragma unroll
                                        for (int b = 0; b < B; b++) a[b][3] = __dadd_rn(a[b][3], v[b]);
                                        break;
                                    default:
#pragma unroll
                                        for (int b = 0; b < B; b++) a[b][4] = __dadd_rn(a[b][4], v[b]);
                                }
                            };
                            int j = ja;
                            for (; j + B - 1 <= jb; j += B) {       // B read bases per trip: the window of weights rotates through the names
                                double nw[B];
#pragma unroll
                                for (int b = 0; b < B; b++) nw[b] = ld(i0 - 1 - b);
#pragma unroll
                                for (int t = 0; t < B; t++) {       // read base j+t: row x+b takes weight index i0 + b - t
                                    double v[B];
#pragma unroll
                                    for (int b = 0; b < B; b++) v[b] = (b - t >= 0) ? w[b - t] : nw[t - b - 1];
                                    add(rc[j + t], v);
                                }
#pragma unroll
                                for (int b = 0; b < B; b++) w[b] = nw[B - 1 - b];
                                i0 -= B;
                            }
                            for (; j <= jb; j++) {
                                add(rc[j], w);
                                i0--;
#pragma unroll
                                for (int b = B - 1; b > 0; b--) w[b] = w[b - 1];
                                w[0] = ld(i0);
                            }
                        }
#if FB_TILEBAR
                        // every tile runs its own chain part 0 -> 1 -> ...: only the warps of that tile meet (named barrier 1 + tile), so a
                        // tile does not wait for the slowest gather warp of another tile; the CTA meets once, after the gather
                        const bool tileBar = tiled && nTiles <= 15;
#else
                        const bool tileBar = false;
#endif
                        for (int sp = 0; sp < (tileBar ? (myIdle ? 0 : myParts) : split); sp++) {
                            if (tileBar && sp > 0) asm volatile("bar.sync %0, %1;" :: "r"(1 + myTile), "r"(32 * myParts) : "memory");
                            if (sp == s && gi < nG && !(tiled && myIdle)) {
                                // rows x .. x+B-1 (x a multiple of B, planes padded to a multiple of 4): aligned pairs, rows beyond Lg are padding
#pragma unroll
                                for (int k = 0; k < 5; k++)
#pragma unroll
                                    for (int b = 0; b < B; b += 2) {
                                        double2* cp = (double2*)(C + k * LgC + x + b);
                                        double2 v = *cp;
                                        v.x = __dadd_rn(v.x, a[b][k]); v.y = __dadd_rn(v.y, a[b + 1][k]);
                                        *cp = v;
                                    }
                            }
                            if (tiled && !tileBar) { __syncthreads(); if (sp == 0) PHASE(3); }
                        }
                        if (tiled) break;
                    }
                };
                if (rowsPerThread == 4) gather(std::integral_constant<int, 4>()); else gather(std::integral_constant<int, 2>());
                __syncthreads();
                PHASE(4);
            }
            // ================= computeSequence(0,0) =================
            for (int x = tid; x < Lg; x += kThreads) {
                double mx = 0; int mi = -1;
#pragma unroll
                for (int k = 0; k < 5; k++) { double v = C[k * LgC + x]; if (v > mx) { mx = v; mi = k; } }
                unsigned char c = (mi >= 0 && mi < 4) ? (unsigned char)mi : 4;
                if (call == 0 || G[F + x] != c) s_gchg = 1;      // benign race: all writers store 1
                G[F + x] = c; oSoft[x] = c;
            }
            __syncthreads();
            PHASE(5);
            gChanged = (call == 0) || (s_gchg != 0);
            // ================= pass 2 =================
            pass2(slot, true);
            // ================= computeSequence(1,1) + comp_count (Figbird.cpp:3916-3927), M-step =================
            // one sweep over the gap rows: hard consensus / coverage from the votes, and (unless this was the extra pass) the
            // M-step of the row; one barrier, after which every thread updates its copy of comp_count from the shared flag
            {
                int same = prevValid;
                for (int x = tid; x < Lg; x += kThreads) {
                    if (unm) {
                        int mx = 0, mi = -1;
#pragma unroll
                        for (int k = 0; k < 5; k++) { int v = TSMEM ? NC[k * Lg + x] : __ldcg(&NC[k * Lg + x]); if (v > mx) { mx = v; mi = k; } }
                        unsigned char h = (mx > 0 && mi >= 0 && mi < 4) ? (unsigned char)mi : 4;
                        oHard[x] = h; oCov[x] = mx;
                        if (!prevValid || PREV[x] != h) same = 0;
                        PREV[x] = h;
                    } else { oHard[x] = 4; oCov[x] = 0; }
                    if (!extra) mstepRow(x);
                }
                if (unm && !same && Lg > 0) s_same[call & 1] = 0;   // benign race: all writers store 0
            }
            calls++;
            __syncthreads();
            PHASE(9);
            if (unm) {
                // an empty previous string equals the new one only when Lg == 0
                const int equal = (Lg == 0) ? 1 : (prevValid && s_same[call & 1]);
                comp = equal ? comp + 1 : 0;
                prevValid = 1;
            }
            if (extra) break;
            if (unm && !(it.flags & FB_FLAG_NO_COMP_STOP) && comp >= 5) emDone = true;
            if (call + 1 >= it.max_rounds) emDone = true;
            if (emDone && !(it.flags & FB_FLAG_EXTRA_PASS)) break;
        }
        if (it.off_counts >= 0) {
            double* oc = (double*)(out + it.off_counts);
            for (int i = tid; i < 5 * Lg; i += kThreads) { int x = i / 5, k = i % 5; oc[i] = C[k * LgC + x]; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        FbItemOut* H = (FbItemOut*)out;
        const long long p1calls = (it.kind == FB_ITEM_HARD) ? 0 : calls;
        H->calls = calls; H->comp_count = comp; H->flags = s_flags; H->placements = sumN * p1calls;
        atomicAdd(&prm.counters[0], (unsigned long long)(sumN * p1calls)); atomicAdd(&prm.counters[1], (unsigned long long)(sumN * calls));
        atomicAdd(&prm.counters[2], (unsigned long long)(sumTerms * (p1calls + calls)));
        atomicAdd(&prm.counters[3], s_lane1); atomicAdd(&prm.counters[4], s_lane2);
#if FB_PHASES
        ph[10] = clock64() - phLast;
        for (int k = 0; k < 12; k++) atomicAdd(&prm.counters[8 + k], (unsigned long long)ph[k]);
        atomicAdd(&prm.counters[20], (unsigned long long)calls);
#endif
    }
}

// ---- diagnostic: FP64 pipe peak without FMA contraction (the ceiling of this path) and with FMA
template <int MODE>
__global__ void __launch_bounds__(256) fb_fp64_peak_kernel(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.9999999, b = 1e-12;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { a0 = __dmul_rn(a0, m); a1 = __dmul_rn(a1, m); a2 = __dmul_rn(a2, m); a3 = __dmul_rn(a3, m); a4 = __dmul_rn(a4, m); a5 = __dmul_rn(a5, m); a6 = __dmul_rn(a6, m); a7 = __dmul_rn(a7, m); }
        else { a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b); a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { c->err = std::string(#x) + ": " + cudaGetErrorString(e_); return FB_ERR_CUDA; } } while (0)

template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) { if (n <= cap) return cudaSuccess; if (p) cudaFree(p); p = nullptr; cap = 0; size_t want = n + n / 2 + 64; cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T)); if (e == cudaSuccess) cap = want; return e; }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct fb_ctx {
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evDone = nullptr;                 // blocking-sync event: the host thread sleeps while a launch runs (several lanes and GPUs share the host cores)
    cudaStream_t bstream[kNumBuckets + 1] = {};   // one stream per shared-memory bucket (+1: global-table items)
    cudaEvent_t bev[kNumBuckets + 1] = {};
    cudaEvent_t bev0[kNumBuckets + 1] = {};      // diagnostics (FIGBIRD_BUCKET_STATS): start of each bucket's launch
    double bucketMs[kNumBuckets + 1] = {}; long long bucketItems[kNumBuckets + 1] = {}, bucketLaunches[kNumBuckets + 1] = {};
    bool bucketStats = false;
    double hostPrepS = 0, hostWaitS = 0, hostPostS = 0; long long hostItems = 0;      // diagnostics: wall seconds of fb_em_run's host phases
    int smemOptin = 0;
    bool haveModel = false, haveBatch = false, latencyCritical = false;
    DevModel dm{};
    DevBuf<double> d_e, d_match, d_pdf, d_lfrf;
    std::vector<DevGap> hGaps; std::vector<int> hGapMaxLen;
    DevBuf<DevGap> d_gaps; DevBuf<int> d_rlen, d_rmate, d_rgap, d_pl, d_pr; DevBuf<long long> d_roff;
    DevBuf<unsigned char> d_rfl, d_jlo, d_jcut;
    DevBuf<uint2> d_codes_pk, d_flank_pk; DevBuf<long long> d_pkoff;
    DevBuf<DevItem> d_items; DevBuf<unsigned char> d_in, d_out, d_scratch, d_meta;
    int nReads = 0; bool flankDirty = true;        // LF/RF products must be (re)computed before the next fb_em_run
    DevBuf<unsigned long long> d_ctr;
    unsigned char* h_out = nullptr; size_t h_out_cap = 0;     // pinned result arena
    unsigned char* h_in = nullptr; size_t h_in_cap = 0;       // pinned staging for items + inputs
    unsigned long long* h_ctr = nullptr;                      // pinned copy of the device counters (a pageable target would make the copy -- and the host thread -- wait spinning)
    FbCounters ctr{};
    double hEtab[512] = {};        // host copy of Params::e_tab
};

// Kernel intervals of all contexts on one physical device, on one time base (an epoch event per device): contexts that
// share a GPU overlap their kernels, so the device time of a run is the union of the intervals, not the sum.
struct DevClock { std::mutex mu; cudaEvent_t epoch = nullptr; double lastEnd = 0, unionMs = 0; };
static DevClock g_clock[64];

extern "C" const char* fb_engine_name(void) { return "cuda-sm100a"; }
extern "C" const char* fb_last_error(const fb_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" fb_status fb_ctx_create(int32_t device, fb_ctx** out) {
    if (!out) return FB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return FB_ERR_NODEVICE;
    fb_ctx* c = new fb_ctx();
    c->device = device;
    *out = c;
    CK(cudaSetDevice(device));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, device));
    if (pr.major < 10) { c->err = "device is not sm_100-class (this library is built for sm_100a only)"; return FB_ERR_NODEVICE; }
    c->smemOptin = (int)pr.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->evDone, cudaEventBlockingSync | cudaEventDisableTiming));
    CK(cudaFuncSetAttribute(fb_em_kernel<true, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CK(cudaFuncSetAttribute(fb_em_kernel<true, kThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CK(cudaFuncSetAttribute(fb_em_kernel<false, kThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    c->bucketStats = getenv("FIGBIRD_BUCKET_STATS") != nullptr;
    for (int b = 0; b <= kNumBuckets; b++) {
        CK(cudaStreamCreateWithFlags(&c->bstream[b], cudaStreamNonBlocking));
        if (c->bucketStats) { CK(cudaEventCreate(&c->bev[b])); CK(cudaEventCreate(&c->bev0[b])); }
        else CK(cudaEventCreateWithFlags(&c->bev[b], cudaEventDisableTiming));
    }
    CK(c->d_ctr.ensure(32)); CK(cudaMemset(c->d_ctr.p, 0, 32 * sizeof(unsigned long long)));
    CK(cudaMallocHost((void**)&c->h_ctr, 32 * sizeof(unsigned long long))); memset(c->h_ctr, 0, 32 * sizeof(unsigned long long));
    if (device < 64) {
        DevClock& k = g_clock[device]; std::lock_guard<std::mutex> l(k.mu);
        if (!k.epoch) { CK(cudaEventCreate(&k.epoch)); CK(cudaEventRecord(k.epoch, c->stream)); CK(cudaEventSynchronize(k.epoch)); }
    }
    return FB_OK;
}

extern "C" void fb_ctx_destroy(fb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->d_e.release(); c->d_match.release(); c->d_pdf.release(); c->d_lfrf.release(); c->d_rgap.release(); c->d_meta.release();
    c->d_gaps.release(); c->d_rlen.release(); c->d_rmate.release(); c->d_pl.release(); c->d_pr.release(); c->d_roff.release();
    c->d_rfl.release(); c->d_jlo.release(); c->d_jcut.release(); c->d_codes_pk.release(); c->d_flank_pk.release(); c->d_pkoff.release();
    c->d_items.release(); c->d_in.release(); c->d_out.release(); c->d_scratch.release(); c->d_ctr.release();
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    for (int b = 0; b <= kNumBuckets; b++) { if (c->bstream[b]) cudaStreamDestroy(c->bstream[b]); if (c->bev[b]) cudaEventDestroy(c->bev[b]); if (c->bev0[b]) cudaEventDestroy(c->bev0[b]); }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evDone) cudaEventDestroy(c->evDone);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" fb_status fb_ctx_set_latency_critical(fb_ctx* c, int32_t on) {
    if (!c) return FB_ERR_ARG;
    CK(cudaSetDevice(c->device));
    int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // (numerically lower = higher priority)
    const int prio = on ? hi : lo;
    if (c->latencyCritical == (on != 0)) return FB_OK;
    CK(cudaStreamSynchronize(c->stream));
    for (int b = 0; b <= kNumBuckets; b++) {
        if (c->bstream[b]) { CK(cudaStreamSynchronize(c->bstream[b])); CK(cudaStreamDestroy(c->bstream[b])); c->bstream[b] = nullptr; }
        CK(cudaStreamCreateWithPriority(&c->bstream[b], cudaStreamNonBlocking, prio));
    }
    CK(cudaStreamDestroy(c->stream)); c->stream = nullptr;
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio));
    c->latencyCritical = (on != 0);
    return FB_OK;
}

template <class T> static fb_status upload(fb_ctx* c, DevBuf<T>& b, const T* src, size_t n) {
    CK(b.ensure(n ? n : 1));
    if (n) { CK(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream)); c->ctr.h2d_bytes += (int64_t)(n * sizeof(T)); }
    return FB_OK;
}

extern "C" fb_status fb_model_upload(fb_ctx* c, const FbModel* m) {
    if (!c || !m || m->max_read_len <= 0 || m->n_insert <= 0 || !m->err_pos || !m->ins_pos || !m->del_pos || !m->insert_pdf) return FB_ERR_ARG;
    // every argument check comes before the first mutation: a rejected upload leaves the context as it was
    if (m->max_read_len > kMaxReadLen) { c->err = "reads longer than 255 bases are not supported by this build"; return FB_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    const int RL = m->max_read_len;
    c->haveModel = false;
    std::vector<double> match(RL);
    bool prunable = true;      // every pass-2 factor in [0, 1] => running products only fall (pass-2 pruning is exact)
    for (int k = 0; k < RL; k++) {
        match[k] = 1 - m->err_pos[k] - m->ins_pos[k] - m->del_pos[k];                 // Figbird.cpp:3400
        if (!(match[k] >= 0.0 && match[k] <= 1.0 && m->err_pos[k] >= 0.0 && m->err_pos[k] <= 1.0)) prunable = false;
    }
    for (int i = 0; i < 25; i++) if (!(m->err_type[i] >= 0.0 && m->err_type[i] <= 1.0)) prunable = false;
    fb_status s;
    if ((s = upload(c, c->d_e, m->err_pos, RL))) return s;
    if ((s = upload(c, c->d_match, match.data(), RL))) return s;
    if ((s = upload(c, c->d_pdf, m->insert_pdf, m->n_insert))) return s;
    CK(cudaStreamSynchronize(c->stream));
    c->dm.e = c->d_e.p; c->dm.match = c->d_match.p; c->dm.pdf = c->d_pdf.p; c->dm.n_insert = m->n_insert;
    memcpy(c->dm.etp, m->err_type, sizeof c->dm.etp);
    c->dm.tmin = m->insert_min; c->dm.tmax = m->insert_max; c->dm.max_read_len = RL; c->dm.prunable = prunable ? 1 : 0;
    {   // largest pass-2 mismatch factor e[k] * ETP[f][c], f != c (the kernel's table lookup, codes 0..4)
        double emax = 0, tmax = 0;
        for (int k = 0; k < RL; k++) emax = std::max(emax, m->err_pos[k]);
        for (int f = 0; f < 5; f++) for (int cc = 0; cc < 5; cc++) if (f != cc) tmax = std::max(tmax, m->err_type[f * 5 + cc]);
        const double mm = emax * tmax * (1.0 + 1e-12);
        c->dm.mis_a = (prunable && mm > 0.0 && mm < 0.5) ? (int)std::floor(-std::log2(mm)) : 0;
        if (c->dm.mis_a < 1) c->dm.mis_a = 0;
    }
    // accept iff -log10(p) < cutoff (Figbird.cpp:3474,3852).  log10 is monotone, so the accepted set is
    // {p >= T}; find T = the smallest double glibc accepts, by bisection on the bit pattern.
    {
        const double cut = (double)m->prob_cutoff;
        auto acc = [&](double p) { return -log10(p) < cut; };
        double lo = pow(10.0, -cut) * 0.5, hi = pow(10.0, -cut) * 2.0;
        if (!(lo > 0)) lo = DBL_TRUE_MIN;
        if (acc(lo)) c->dm.accept_min_p = lo;        // cutoff so large that everything positive passes in range
        else if (!acc(hi)) c->dm.accept_min_p = hi;  // degenerate; keep monotone behaviour
        else {
            unsigned long long a, b; memcpy(&a, &lo, 8); memcpy(&b, &hi, 8);
            while (b - a > 1) { unsigned long long mid = a + (b - a) / 2; double pm; memcpy(&pm, &mid, 8); if (acc(pm)) b = mid; else a = mid; }
            double t; memcpy(&t, &b, 8); c->dm.accept_min_p = t;
        }
        if (m->prob_cutoff <= 0) c->dm.accept_min_p = INFINITY;   // -log10(p) < 0 needs p > 1: never for a probability
    }
    for (int k = 0; k < RL; k++) { c->hEtab[k] = m->err_pos[k]; c->hEtab[256 + (RL - 1 - k)] = m->err_pos[k]; }
    c->haveModel = true; c->flankDirty = true;
    return FB_OK;
}

extern "C" fb_status fb_batch_upload(fb_ctx* c, const FbGapBatch* b) {
    if (!c || !b || b->n_gaps < 0) return FB_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->hGaps.resize(b->n_gaps); c->hGapMaxLen.assign(b->n_gaps, 1);
    std::vector<int> readGap((size_t)std::max(b->n_reads, 1), -1);
    if (b->n_reads < 0 || b->n_codes < 0 || b->n_flank < 0 || b->n_pile_rows < 0 || (b->n_gaps > 0 && !b->gaps)) return FB_ERR_ARG;
    {   // reads must not share code bytes (the per-read flank products are written at 2 x read_code_off)
        bool ascending = true;
        for (int q = 0; q + 1 < b->n_reads && ascending; q++) if (b->read_code_off[q + 1] < b->read_code_off[q] + b->read_len[q]) ascending = false;
        if (!ascending) {
            std::vector<int> ord(b->n_reads); for (int q = 0; q < b->n_reads; q++) ord[q] = q;
            std::sort(ord.begin(), ord.end(), [&](int x, int y) { return b->read_code_off[x] < b->read_code_off[y]; });
            for (int k = 0; k + 1 < b->n_reads; k++) if (b->read_code_off[ord[k + 1]] < b->read_code_off[ord[k]] + b->read_len[ord[k]]) { c->err = "read_code_off ranges overlap"; return FB_ERR_ARG; }
        }
    }
    for (int i = 0; i < b->n_gaps; i++) {
        const FbGap& g = b->gaps[i];
        if (g.mode != FB_MODE_PARTIAL && g.mode != FB_MODE_UNMAPPED) { c->err = "bad gap mode"; return FB_ERR_ARG; }
        if (g.n_reads < 0 || g.flank_len < 0 || g.pile_len < 0 || g.pile_begin < 0 || (long long)g.pile_begin + g.pile_len > b->n_pile_rows) { c->err = "pile-up rows out of range"; return FB_ERR_ARG; }
        DevGap d{}; d.gap_start = g.gap_start; d.mode = g.mode; d.orig_len = g.orig_len; d.n_reads = g.n_reads; d.read_begin = g.read_begin;
        d.flank_len = g.flank_len; d.flank_begin = g.flank_begin; d.pile_len = g.pile_len; d.pile_begin = g.pile_begin;
        c->hGaps[i] = d;
        int ml = 1;
        for (int q = 0; q < g.n_reads; q++) {
            if (g.read_begin + q < 0 || g.read_begin + q >= b->n_reads) { c->err = "read index out of range"; return FB_ERR_ARG; }
            int len = b->read_len[g.read_begin + q];
            if (b->read_code_off[g.read_begin + q] < 0 || b->read_code_off[g.read_begin + q] + len > b->n_codes) { c->err = "read codes out of range"; return FB_ERR_ARG; }
            readGap[g.read_begin + q] = i;
            if (len > g.flank_len + 1) { c->err = "flank_len must be >= read length - 1"; return FB_ERR_ARG; }
            if (len < 0 || len > kMaxReadLen) { c->err = "read length outside [0, 255]"; return FB_ERR_ARG; }
            if ((int)b->read_jlo[g.read_begin + q] + (int)b->read_jcut[g.read_begin + q] > len) { c->err = "read_jlo + read_jcut exceeds the read length"; return FB_ERR_ARG; }
            ml = len > ml ? len : ml;
        }
        c->hGapMaxLen[i] = ml;
    }
    fb_status s;
    if ((s = upload(c, c->d_rlen, b->read_len, (size_t)b->n_reads))) return s;
    if ((s = upload(c, c->d_rmate, b->read_mate, (size_t)b->n_reads))) return s;
    if ((s = upload(c, c->d_rgap, readGap.data(), (size_t)b->n_reads))) return s;
    CK(cudaStreamSynchronize(c->stream));
    CK(c->d_lfrf.ensure(2 * (size_t)std::max<int64_t>(b->n_codes, 1)));
    { std::vector<long long> ro(b->read_code_off, b->read_code_off + b->n_reads); if ((s = upload(c, c->d_roff, ro.data(), ro.size()))) return s; CK(cudaStreamSynchronize(c->stream)); }
    if ((s = upload(c, c->d_rfl, b->read_flags, (size_t)b->n_reads))) return s;
    if ((s = upload(c, c->d_jlo, b->read_jlo, (size_t)b->n_reads))) return s;
    if ((s = upload(c, c->d_jcut, b->read_jcut, (size_t)b->n_reads))) return s;
    {   // reads and flanks go to the device 2-bit packed with an N mask (16 bases per 8-byte word, each sequence word-aligned)
        auto pack = [](const uint8_t* src, int n, std::vector<uint2>& dst) {
            for (int w = 0; w < (n + 15) / 16; w++) {
                uint2 v = make_uint2(0u, 0u);
                for (int t = 0; t < 16 && 16 * w + t < n; t++) {
                    const unsigned cde = src[16 * w + t];
                    if (cde < 4) v.x |= cde << (2 * t); else v.y |= 1u << t;
                }
                dst.push_back(v);
            }
        };
        std::vector<uint2> pk; std::vector<long long> pkoff((size_t)std::max(b->n_reads, 1), 0);
        pk.reserve((size_t)b->n_codes / 16 + (size_t)b->n_reads + 1);
        for (int q = 0; q < b->n_reads; q++) {
            pkoff[q] = (long long)pk.size();
            const long long o = b->read_code_off[q]; const int len = b->read_len[q];
            if (o >= 0 && len > 0 && o + len <= b->n_codes) pack(b->read_codes + o, len, pk);
        }
        if (pk.empty()) pk.push_back(make_uint2(0u, 0xffffu));
        std::vector<uint2> fpk;
        for (int i = 0; i < b->n_gaps; i++) {
            const FbGap& g = b->gaps[i];
            c->hGaps[i].flank_pk = (int)fpk.size();
            if (g.flank_len > 0 && g.flank_begin >= 0 && (long long)g.flank_begin + 2LL * g.flank_len <= b->n_flank) {
                pack(b->flank_codes + g.flank_begin, g.flank_len, fpk);
                pack(b->flank_codes + g.flank_begin + g.flank_len, g.flank_len, fpk);
            } else if (g.flank_len > 0) { c->err = "flank codes out of range"; return FB_ERR_ARG; }
        }
        if (fpk.empty()) fpk.push_back(make_uint2(0u, 0xffffu));
        if ((s = upload(c, c->d_gaps, c->hGaps.data(), c->hGaps.size()))) return s;      // (after flank_pk is known)
        if ((s = upload(c, c->d_codes_pk, pk.data(), pk.size()))) return s;
        if ((s = upload(c, c->d_pkoff, pkoff.data(), pkoff.size()))) return s;
        if ((s = upload(c, c->d_flank_pk, fpk.data(), fpk.size()))) return s;
        CK(cudaStreamSynchronize(c->stream));
    }
    if ((s = upload(c, c->d_pl, b->pile_left, (size_t)b->n_pile_rows * 4))) return s;
    if ((s = upload(c, c->d_pr, b->pile_right, (size_t)b->n_pile_rows * 4))) return s;
    CK(cudaStreamSynchronize(c->stream));
    c->nReads = b->n_reads; c->haveBatch = true; c->flankDirty = true;
    return FB_OK;
}

extern "C" fb_status fb_get_counters(const fb_ctx* c, FbCounters* out) {
    if (!c || !out) return FB_ERR_ARG;
    *out = c->ctr;
    if (c->bucketStats) {
        fprintf(stderr, "bucket stats dev %d (cumulative; buckets overlap on the device): total %.0f ms |", c->device, c->ctr.device_ms);
        for (int b = 0; b <= kNumBuckets; b++) fprintf(stderr, " b%d%s: %.0f ms %lld items %lld launches |", b, b == kNumBuckets ? "(global tables)" : "", c->bucketMs[b], c->bucketItems[b], c->bucketLaunches[b]);
        fprintf(stderr, " fb_em_run host: prep %.3f s, wait %.3f s, post %.3f s, %lld items\n", c->hostPrepS, c->hostWaitS, c->hostPostS, c->hostItems);
    }
    if (c->device < 64) { DevClock& k = g_clock[c->device]; std::lock_guard<std::mutex> l(k.mu); out->device_union_ms = k.unionMs; }
    return FB_OK;
}

extern "C" fb_status fb_em_run(fb_ctx* c, const FbWorkItem* items, int32_t n, const FbItemOut** out) {
    if (!c || !items || !out || n < 0) return FB_ERR_ARG;
    if (!c->haveModel || !c->haveBatch) { c->err = "model/batch not uploaded"; return FB_ERR_STATE; }
    if (n == 0) return FB_OK;
    CK(cudaSetDevice(c->device));
    const auto tRun0 = std::chrono::steady_clock::now();
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    std::vector<DevItem> di(n);
    size_t outTotal = 0, inTotal = 0, scratchTotal = 0, metaTotal = 0;
    const int bandMax = c->dm.tmax - c->dm.tmin + 1;
    std::vector<int> bucketOf(n, 0);
    int bucketSmem[kNumBuckets + 1] = {0, 0, 0, 0, 0};
    int bucketCount[kNumBuckets + 1] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const FbWorkItem& it = items[i];
        if (it.gap < 0 || it.gap >= (int)c->hGaps.size() || it.cand_len < 0 || it.cand_len > 60000) { c->err = "bad work item"; return FB_ERR_ARG; }
        if (it.kind != FB_ITEM_EM && it.kind != FB_ITEM_HARD) { c->err = "bad work item kind"; return FB_ERR_ARG; }
        if (it.max_rounds < 0 || it.max_rounds > 100000) { c->err = "max_rounds outside [0, 100000]"; return FB_ERR_ARG; }
        if (it.kind == FB_ITEM_EM && (it.flags & FB_FLAG_RECORD_ALL) && it.max_rounds + ((it.flags & FB_FLAG_EXTRA_PASS) ? 1 : 0) == 0) { c->err = "RECORD_ALL item without a call to record"; return FB_ERR_ARG; }
        const DevGap& g = c->hGaps[it.gap];
        DevItem d{};
        d.kind = it.kind; d.gap = it.gap; d.Lg = it.cand_len; d.max_rounds = it.max_rounds; d.flags = it.flags; d.comp_in = it.comp_count_in;
        const int Lg = it.cand_len, R = g.n_reads;
        int slots = 1;
        if (it.kind == FB_ITEM_EM && (it.flags & FB_FLAG_RECORD_ALL)) slots = it.max_rounds + ((it.flags & FB_FLAG_EXTRA_PASS) ? 1 : 0);
        d.n_slots = slots;
        d.out_off = (long long)outTotal;
        size_t o = al(sizeof(FbItemOut));
        d.off_p1 = o; o += al(sizeof(double) * slots * R);
        d.off_p2 = o; o += al(sizeof(double) * slots * R);
        d.off_pos = o; o += al(sizeof(int) * slots * R);
        d.off_soft = o; o += al(Lg);
        d.off_hard = o; o += al(Lg);
        d.off_cov = o; o += al(sizeof(int) * Lg);
        if (it.flags & FB_FLAG_WANT_COUNTS) { d.off_counts = o; o += al(sizeof(double) * 5 * Lg); } else d.off_counts = -1;
        outTotal += o;
        d.counts_in_off = -1; d.string_in_off = -1;
        if (it.kind == FB_ITEM_EM && (it.flags & FB_FLAG_RESUME)) {
            if (!it.counts_in) { c->err = "RESUME item without counts_in"; return FB_ERR_ARG; }
            d.counts_in_off = (long long)inTotal; inTotal += al(sizeof(double) * 5 * Lg);
        }
        if (it.string_in && Lg > 0) { d.string_in_off = (long long)inTotal; inTotal += al(Lg); }
        // shared-memory plan: tables + model + at least one read's chunk must fit, else tables go to global scratch
        const int ml = c->hGapMaxLen[it.gap];
        d.max_len = ml;
        d.meta_off = (long long)metaTotal; metaTotal += metaBytes(R);
        const Plan pl = makePlan(Lg, g.flank_len, c->dm.max_read_len, ml, g.mode, bandMax);
        const long long want = std::min<long long>((long long)pl.perRead * std::max(R, 1), kChunkWant);
        const long long chunk = std::max<long long>(want, pl.perRead) + 128;
        if ((long long)pl.tableBytes + pl.localFixed + pl.perRead + 128 <= kMaxSmem) {
            d.tables_in_smem = 1; d.scratch_off = -1;
            const int need = (int)std::min<long long>(kMaxSmem, (long long)pl.tableBytes + pl.localFixed + chunk);
            int b = 0; while (b < kNumBuckets - 1 && need > bucketCap(b)) b++;
            bucketOf[i] = b; bucketSmem[b] = std::max(bucketSmem[b], need);
        } else {
            if ((long long)pl.localFixed + pl.perRead + 128 > kMaxSmem) { c->err = "candidate length too large for one weight row in shared memory"; return FB_ERR_ARG; }
            d.tables_in_smem = 0; d.scratch_off = (long long)scratchTotal; scratchTotal += (size_t)pl.tableBytes;
            bucketOf[i] = kNumBuckets; bucketSmem[kNumBuckets] = std::max(bucketSmem[kNumBuckets], (int)std::min<long long>(kMaxSmem, (long long)pl.localFixed + chunk));
        }
        bucketCount[bucketOf[i]]++;
        di[i] = d;
    }
    // launch order: per bucket, longest chains first (unmapped EM items with many reads), so the tail is short
    std::vector<int> order(n);
    int bucketBegin[kNumBuckets + 2];
    {
        int acc = 0;
        for (int b = 0; b <= kNumBuckets; b++) { bucketBegin[b] = acc; acc += bucketCount[b]; }
        bucketBegin[kNumBuckets + 1] = acc;
        int fill[kNumBuckets + 1];
        for (int b = 0; b <= kNumBuckets; b++) fill[b] = bucketBegin[b];
        for (int i = 0; i < n; i++) order[fill[bucketOf[i]]++] = i;
        // (keys computed once per item, not once per comparison)
        std::vector<float> key(n);
        for (int i = 0; i < n; i++) {
            const DevGap& g = c->hGaps[items[i].gap];
            const float rounds = items[i].kind == FB_ITEM_HARD ? 0.3f : (g.mode == FB_MODE_UNMAPPED ? 12.0f : 3.0f);
            key[i] = rounds * (float)g.n_reads * (float)(items[i].cand_len + 100);
        }
        for (int b = 0; b <= kNumBuckets; b++)
            std::stable_sort(order.begin() + bucketBegin[b], order.begin() + bucketBegin[b + 1], [&](int x, int y) { return key[x] > key[y]; });
    }
    // ---- stage inputs
    const size_t itemsBytes = al(sizeof(DevItem) * (size_t)n);
    const size_t orderOff = itemsBytes + al(inTotal);
    const size_t inBytes = orderOff + al(sizeof(int) * (size_t)n) + 16;
    if (inBytes > c->h_in_cap) { if (c->h_in) cudaFreeHost(c->h_in); c->h_in = nullptr; c->h_in_cap = 0; size_t want = 2 * inBytes; CK(cudaMallocHost((void**)&c->h_in, want)); c->h_in_cap = want; }
    memcpy(c->h_in, di.data(), sizeof(DevItem) * (size_t)n);
    memcpy(c->h_in + orderOff, order.data(), sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const FbWorkItem& it = items[i];
        if (di[i].counts_in_off >= 0) memcpy(c->h_in + itemsBytes + di[i].counts_in_off, it.counts_in, sizeof(double) * 5 * (size_t)it.cand_len);
        if (di[i].string_in_off >= 0) memcpy(c->h_in + itemsBytes + di[i].string_in_off, it.string_in, (size_t)it.cand_len);
    }
    CK(c->d_in.ensure(inBytes));
    CK(c->d_out.ensure(outTotal + 16));
    CK(c->d_scratch.ensure(scratchTotal + 16));
    CK(c->d_meta.ensure(metaTotal + 16));
    if (outTotal + 16 > c->h_out_cap) { if (c->h_out) cudaFreeHost(c->h_out); c->h_out = nullptr; c->h_out_cap = 0; size_t want = 2 * outTotal + 64; CK(cudaMallocHost((void**)&c->h_out, want)); c->h_out_cap = want; }
    CK(cudaMemcpyAsync(c->d_in.p, c->h_in, inBytes, cudaMemcpyHostToDevice, c->stream));
    c->ctr.h2d_bytes += (int64_t)inBytes;

    Params prm{};
    prm.m = c->dm;
    prm.gaps = c->d_gaps.p; prm.read_len = c->d_rlen.p; prm.read_off = c->d_roff.p; prm.read_mate = c->d_rmate.p;
    prm.read_flags = c->d_rfl.p; prm.read_jlo = c->d_jlo.p; prm.read_jcut = c->d_jcut.p;
    prm.codes_pk = c->d_codes_pk.p; prm.read_pk_off = c->d_pkoff.p; prm.flank_pk = c->d_flank_pk.p;
    prm.pile_l = c->d_pl.p; prm.pile_r = c->d_pr.p; prm.read_gap = c->d_rgap.p; prm.lfrf = c->d_lfrf.p; prm.meta = c->d_meta.p;
    prm.items = (const DevItem*)c->d_in.p; prm.in_arena = c->d_in.p + itemsBytes; prm.out_arena = c->d_out.p; prm.scratch = c->d_scratch.p;
    prm.counters = c->d_ctr.p;
    memcpy(prm.e_tab, c->hEtab, sizeof prm.e_tab);

    int launches = 0;
    if (c->flankDirty) {      // flank products of every read of the batch, once (model and batch are both resident now)
        if (c->nReads > 0) { fb_flank_kernel<<<c->nReads, 128, 0, c->stream>>>(prm, c->nReads); CK(cudaGetLastError()); launches++; }
        c->flankDirty = false;
    }
    // one launch per shared-memory bucket, on its own stream, so small items run at high occupancy beside large ones
    CK(cudaEventRecord(c->ev0, c->stream));
    const int* d_order = (const int*)(c->d_in.p + orderOff);
    for (int b = 0; b <= kNumBuckets; b++) {
        if (!bucketCount[b]) continue;
        CK(cudaStreamWaitEvent(c->bstream[b], c->ev0, 0));
        if (c->bucketStats) CK(cudaEventRecord(c->bev0[b], c->bstream[b]));
        if (b < kBigBucket) fb_em_kernel<true, kThreads><<<bucketCount[b], kThreads, bucketSmem[b], c->bstream[b]>>>(prm, d_order + bucketBegin[b], bucketSmem[b]);
        else if (b < kNumBuckets) fb_em_kernel<true, kThreadsBig><<<bucketCount[b], kThreadsBig, bucketSmem[b], c->bstream[b]>>>(prm, d_order + bucketBegin[b], bucketSmem[b]);
        else fb_em_kernel<false, kThreadsBig><<<bucketCount[b], kThreadsBig, bucketSmem[b], c->bstream[b]>>>(prm, d_order + bucketBegin[b], bucketSmem[b]);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->bev[b], c->bstream[b]));
        CK(cudaStreamWaitEvent(c->stream, c->bev[b], 0));
        launches++;
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(c->h_out, c->d_out.p, outTotal, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long* const hc = c->h_ctr;
    CK(cudaMemcpyAsync(hc, c->d_ctr.p, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->evDone, c->stream));
    const auto tRun1 = std::chrono::steady_clock::now();
    CK(cudaEventSynchronize(c->evDone));
    const auto tRun2 = std::chrono::steady_clock::now();
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (c->bucketStats) for (int b = 0; b <= kNumBuckets; b++) if (bucketCount[b]) {
        float bm = 0; if (cudaEventElapsedTime(&bm, c->bev0[b], c->bev[b]) == cudaSuccess) c->bucketMs[b] += bm;
        c->bucketItems[b] += bucketCount[b]; c->bucketLaunches[b]++;
    }
    c->ctr.device_ms += ms; c->ctr.kernel_launches += launches; c->ctr.d2h_bytes += (int64_t)outTotal;
    if (c->device < 64) {
        DevClock& k = g_clock[c->device]; std::lock_guard<std::mutex> l(k.mu);
        float s0 = 0, s1 = 0;
        if (k.epoch && cudaEventElapsedTime(&s0, k.epoch, c->ev0) == cudaSuccess && cudaEventElapsedTime(&s1, k.epoch, c->ev1) == cudaSuccess) {
            const double b = std::max((double)s0, k.lastEnd);
            if (s1 > b) k.unionMs += s1 - b;
            k.lastEnd = std::max(k.lastEnd, (double)s1);
        }
    }
    c->ctr.placements_p1 = (int64_t)hc[0]; c->ctr.placements_p2 = (int64_t)hc[1]; c->ctr.base_terms = (int64_t)hc[2];
    c->ctr.lane_steps_p1 = (int64_t)hc[3]; c->ctr.lane_steps_p2 = (int64_t)hc[4];
#if FB_PHASES
    if (getenv("FIGBIRD_PHASES")) {
        static const char* nm[12] = {"prologue", "walk", "finish1", "gather", "combine", "consensus", "prephase", "pass2", "finish2", "endsweep", "epilogue", "-"};
        double tot = 0; for (int k = 0; k < 11; k++) tot += (double)hc[8 + k];
        fprintf(stderr, "phases (cumulative, %llu rounds):", hc[20]);
        for (int k = 0; k < 11; k++) fprintf(stderr, " %s %.1f%%", nm[k], 100.0 * hc[8 + k] / (tot > 0 ? tot : 1));
        fprintf(stderr, " | cycles/round %.0f\n", tot / (hc[20] ? hc[20] : 1));
    }
#endif
    for (int i = 0; i < n; i++) {
        FbItemOut* H = (FbItemOut*)(c->h_out + di[i].out_off);
        const DevGap& g = c->hGaps[items[i].gap];
        H->n_reads = g.n_reads; H->cand_len = items[i].cand_len; H->n_slots = di[i].n_slots;
        H->off_p1max = di[i].off_p1; H->off_p2max = di[i].off_p2; H->off_pos2 = di[i].off_pos; H->off_soft = di[i].off_soft;
        H->off_hard = di[i].off_hard; H->off_cov = di[i].off_cov; H->off_counts = di[i].off_counts;
        out[i] = H;
    }
    if (c->bucketStats) {
        const auto tRun3 = std::chrono::steady_clock::now();
        c->hostPrepS += std::chrono::duration<double>(tRun1 - tRun0).count(); c->hostWaitS += std::chrono::duration<double>(tRun2 - tRun1).count();
        c->hostPostS += std::chrono::duration<double>(tRun3 - tRun2).count(); c->hostItems += n;
    }
    return FB_OK;
}

// Diagnostic (bench.py roofline denominator): measured FP64 instruction throughput of this GPU.
// out[0] = DMUL-only rate in 1e12 instructions/s (1 flop each: the no-FMA ceiling this path lives under),
// out[1] = DFMA rate in TFLOP/s (2 flop each).
extern "C" fb_status fb_microbench_fp64(fb_ctx* c, double* out2) {
    if (!c || !out2) return FB_ERR_ARG;
    CK(cudaSetDevice(c->device));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, c->device));
    const int blocks = pr.multiProcessorCount * 8, iters = 1 << 14;
    DevBuf<double> buf; CK(buf.ensure((size_t)blocks * 256));
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaEventRecord(c->ev0, c->stream));
            if (mode == 0) fb_fp64_peak_kernel<0><<<blocks, 256, 0, c->stream>>>(buf.p, iters); else fb_fp64_peak_kernel<1><<<blocks, 256, 0, c->stream>>>(buf.p, iters);
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->ev1, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double ops = (double)blocks * 256 * 8 * (double)iters;
        out2[mode] = ops * (mode == 0 ? 1.0 : 2.0) / (best * 1e-3) / 1e12;
    }
    buf.release();
    return FB_OK;
}
