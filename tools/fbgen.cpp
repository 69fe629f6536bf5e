// fbgen -- seeded synthetic data generator for the Figbird gap-fill path (test / bench infrastructure).
//
// Emits, from a known truth genome:
//   <out>/truth.fa         ungapped scaffolds
//   <out>/draft.fa         gapped draft (each gap's true sequence replaced by an N-run whose length is
//                          truth*U(0.8,1.2); negative-overlap gaps duplicate `ov` flank bases around a short N-run)
//   <out>/result1.sam      what `bowtie2 --local -X <x1>` would report (soft clips at gap edges)
//   <out>/result2.sam      what `bowtie2 -X <x2>` (end-to-end) would report (reads touching a gap are unaligned)
//   <out>/truth_gaps.txt   scaffold, draft gap start, N-run length, true length, true sequence
// These are the inputs of the reference's Preprocess (Preprocess.cpp:1861-1870, getSAM :1491-1551), which
// buckets them into the per-gap files the gap filler reads.  bowtie2 itself is not available offline
// (SURVEY.md 8c), so the SAM is written directly from truth, following bowtie2's conventions:
//   * SEQ/QUAL in reference orientation for aligned reads, as sequenced for unaligned ones;
//   * an unaligned read with an aligned mate carries the mate's RNAME/POS, CIGAR "*", YT:Z:UP;
//   * pair flags 99/147/83/163 for concordant pairs (fragment <= -X), 97/145/81/161 otherwise,
//     73/133, 89/165, 69/137, 101/153 for one-end-unaligned, 77/141 for both unaligned;
//   * --local: a read overlapping a gap edge is soft-clipped there if the aligned part scores
//     >= 20+8*ln(L) (match bonus 2, mismatch penalty 6), else unaligned.
// Determinism: splitmix64 + xoshiro256**, no libc rand; same seed => same bytes on any platform.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <mutex>
#include <thread>

struct Rng {
    uint64_t s[4];
    static uint64_t sm(uint64_t& x) { uint64_t z = (x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = sm(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(uni() * (double)n); }
    double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

struct Gap { long tstart; int tlen; int og; int ov; long dstart; };   // truth start/len, N-run len, neg-overlap, draft start
struct Scaf { std::string name, truth, draft; std::vector<Gap> gaps; };

static char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; } return 'N'; }

struct Opt {
    long genome = 30000; int nscaf = 1; int ngaps = 3; int gmin = 10, gmax = 200; int L = 100;
    double mu = 200, sd = 20, cov = 30; uint64_t seed = 1; int x1 = -1, x2 = -1; double e0 = 0.005, e1 = 0.02;
    double negfrac = 0; int negmax = 30; long model_pairs = -1; int near = -1; std::string out = "."; int minsep = -1;
    std::string gaplist;  // optional explicit comma list of true gap lengths
    std::string gappos;   // with --gaplist: explicit truth start positions (>= 0: from the scaffold start; < 0: -d puts the gap end d bases before the scaffold end)
    int lib = 0;          // library tag in read names
    int nreadN = 0;       // 1 in nreadN reads gets an 'N' base (0 = never)
    int threads = -1;     // -1: one RNG stream, one thread (the byte-stable legacy layout the fixtures were made with); >= 0: one RNG stream per
                          // scaffold, scaffolds generated on `threads` threads (0 = all cores); output still deterministic for a seed
};

struct Aln { bool ok = false; long pos = 0; std::string cigar, md; int nm = 0, as = 0; long rstart = 0, rend = 0; };  // pos 1-based draft; rstart/rend: draft extent incl. soft clips

int main(int argc, char** argv) {
    Opt o;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i]; const char* v = argv[i + 1];
        if (k == "--genome") o.genome = atol(v); else if (k == "--scaffolds") o.nscaf = atoi(v); else if (k == "--gaps") o.ngaps = atoi(v);
        else if (k == "--gapmin") o.gmin = atoi(v); else if (k == "--gapmax") o.gmax = atoi(v); else if (k == "--readlen") o.L = atoi(v);
        else if (k == "--insert") o.mu = atof(v); else if (k == "--sd") o.sd = atof(v); else if (k == "--cov") o.cov = atof(v);
        else if (k == "--seed") o.seed = strtoull(v, 0, 10); else if (k == "--x1") o.x1 = atoi(v); else if (k == "--x2") o.x2 = atoi(v);
        else if (k == "--negfrac") o.negfrac = atof(v); else if (k == "--negmax") o.negmax = atoi(v); else if (k == "--model-pairs") o.model_pairs = atol(v);
        else if (k == "--near") o.near = atoi(v); else if (k == "--out") o.out = v; else if (k == "--minsep") o.minsep = atoi(v);
        else if (k == "--gaplist") o.gaplist = v; else if (k == "--gappos") o.gappos = v; else if (k == "--lib") o.lib = atoi(v); else if (k == "--readN") o.nreadN = atoi(v);
        else if (k == "--threads") o.threads = atoi(v); else if (k == "--e0") o.e0 = atof(v); else if (k == "--e1") o.e1 = atof(v);
        else { fprintf(stderr, "fbgen: unknown option %s\n", k.c_str()); return 2; }
    }
    if (o.sd <= 0) o.sd = 0.1 * o.mu;
    if (o.x1 < 0) o.x1 = (int)o.mu;                 // RunFigbird.sh:277  --local -X $maxD1 (= smallest insert size)
    if (o.x2 < 0) o.x2 = (int)(1.15 * o.mu);        // RunFigbird.sh:193-217,333  -X 1.15*insert
    int maxins = (int)(o.mu + 6 * o.sd);
    if (o.minsep < 0) o.minsep = (int)(1.15 * maxins) + o.L + 50;
    Rng rng0(o.seed);
    const int L = o.L;
    const char B[4] = {'A', 'C', 'G', 'T'};

    std::vector<int> explicit_gaps;
    if (!o.gaplist.empty()) { const char* p = o.gaplist.c_str(); while (*p) { explicit_gaps.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p) p++; } o.ngaps = (int)explicit_gaps.size(); }

    std::vector<long> explicit_pos;
    if (!o.gappos.empty()) { const char* p = o.gappos.c_str(); while (*p) { explicit_pos.push_back(atol(p)); while (*p && *p != ',') p++; if (*p) p++; } }

    // ---- scaffolds + gaps
    std::vector<Scaf> sc(o.nscaf);
    long per = o.genome / o.nscaf;
    int gi = 0;
    const bool par = o.threads >= 0;
    auto scafSeed = [&](int s, int what) { return o.seed * 0x9e3779b97f4a7c15ULL + (uint64_t)s * 2 + what + 1; };
    for (int s = 0; s < o.nscaf; s++) {
        Scaf& S = sc[s];
        S.name = "scaffold_" + std::to_string(s + 1);
        S.truth.resize(per);
        Rng srng(scafSeed(s, 0));
        Rng& rng = par ? srng : rng0;
        for (long i = 0; i < per; i++) S.truth[i] = B[rng.next() >> 62];
        int ng = o.ngaps / o.nscaf + (s < o.ngaps % o.nscaf ? 1 : 0);
        if (ng > 0) {
            // evenly sized slots; a gap sits at a random place of its slot, >= minsep from the slot borders
            long slot = per / ng;
            for (int g = 0; g < ng; g++, gi++) {
                Gap G{};
                int tl;
                if (!explicit_gaps.empty()) tl = explicit_gaps[gi];
                else tl = (int)std::lround(std::exp(std::log((double)o.gmin) + rng.uni() * (std::log((double)o.gmax) - std::log((double)o.gmin))));
                bool neg = rng.uni() < o.negfrac;
                long lo = g * slot + o.minsep, hi = (g + 1) * slot - o.minsep - (neg ? 0 : tl);
                if (hi <= lo) { fprintf(stderr, "fbgen: slot too small for gap (slot %ld, minsep %d, len %d)\n", slot, o.minsep, tl); return 2; }
                G.tstart = lo + (long)rng.below(hi - lo);
                if (gi < (int)explicit_pos.size()) { const long v = explicit_pos[gi]; G.tstart = v >= 0 ? v : per + v - (neg ? 0 : tl); }      // scaffold-end cases: the caller keeps them sorted and apart
                if (neg) { G.ov = 5 + (int)rng.below(21); G.tlen = 0; G.og = 1 + (int)rng.below(o.negmax); }
                else { G.ov = 0; G.tlen = tl; G.og = std::max(1, (int)std::lround(tl * (0.8 + 0.4 * rng.uni()))); }
                S.gaps.push_back(G);
            }
        }
        // draft
        long t = 0;
        for (auto& G : S.gaps) {
            S.draft.append(S.truth, t, G.tstart - t);
            G.dstart = (long)S.draft.size();
            S.draft.append(G.og, 'N');
            t = G.tstart + G.tlen - G.ov;           // negative overlap: right flank restarts ov bases early
        }
        S.draft.append(S.truth, t, std::string::npos);
    }
    {
        FILE* f = fopen((o.out + "/truth.fa").c_str(), "w"); FILE* d = fopen((o.out + "/draft.fa").c_str(), "w"); FILE* tg = fopen((o.out + "/truth_gaps.txt").c_str(), "w");
        if (!f || !d || !tg) { fprintf(stderr, "fbgen: cannot write to %s\n", o.out.c_str()); return 1; }
        for (int s = 0; s < o.nscaf; s++) {
            fprintf(f, ">%s\n%s\n", sc[s].name.c_str(), sc[s].truth.c_str());
            fprintf(d, ">%s\n%s\n", sc[s].name.c_str(), sc[s].draft.c_str());
            for (auto& G : sc[s].gaps) fprintf(tg, "%d\t%ld\t%d\t%d\t%d\t%s\n", s, G.dstart, G.og, G.tlen, G.ov, sc[s].truth.substr(G.tstart, G.tlen).c_str());
        }
        fclose(f); fclose(d); fclose(tg);
    }

    // ---- reads
    std::vector<double> err(L); std::string qualf(L, 'I');
    for (int j = 0; j < L; j++) { err[j] = o.e0 + (o.e1 - o.e0) * (L > 1 ? (double)j / (L - 1) : 0); qualf[j] = (char)(33 + (int)std::lround(-10 * std::log10(err[j]))); }
    const double minscore = 20 + 8 * std::log((double)L);

    FILE* s1 = fopen((o.out + "/result1.sam").c_str(), "w"); FILE* s2 = fopen((o.out + "/result2.sam").c_str(), "w");
    static char b1[1 << 22], b2[1 << 22]; setvbuf(s1, b1, _IOFBF, sizeof b1); setvbuf(s2, b2, _IOFBF, sizeof b2);
    for (FILE* f : {s1, s2}) { fprintf(f, "@HD\tVN:1.0\tSO:unsorted\n"); for (auto& S : sc) fprintf(f, "@SQ\tSN:%s\tLN:%zu\n", S.name.c_str(), S.draft.size()); fprintf(f, "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.2.3\n"); }

    struct Buf { std::string a, b; };
    auto appendf = [](std::string& dst, const char* fmt, ...) {
        char tmp[4096]; va_list ap; va_start(ap, fmt); int n = vsnprintf(tmp, sizeof tmp, fmt, ap); va_end(ap);
        if (n >= (int)sizeof tmp) { std::string big((size_t)n + 1, '\0'); va_start(ap, fmt); vsnprintf(&big[0], big.size(), fmt, ap); va_end(ap); dst.append(big.c_str(), (size_t)n); }
        else if (n > 0) dst.append(tmp, (size_t)n);
    };
    auto pairsOf = [&](int s) { return (long)(o.cov * (double)sc[s].truth.size() / (2.0 * L)); };
    // reads of one scaffold -> SAM text of both files.  rng / far_written: the shared stream and counter (legacy) or the scaffold's own
    auto genReads = [&](int s, Rng& rng, long pair_id, long& far_written, long far_cap, Buf& buf) {
        Scaf& S = sc[s];
        long npairs = pairsOf(s);
        // truth -> draft coordinate of a truth position left of / right of each gap
        auto align = [&](long ts, const std::string& rd_ref /*read in reference orientation*/, bool local) -> Aln {
            // ts: truth start of the read (reference orientation), covers [ts, ts+L)
            Aln a; long te = ts + L;
            // find first gap with truth end > ts  (gaps sorted)
            long shift = 0; const Gap* hit = nullptr;
            for (auto& G : S.gaps) {
                long gs = G.tstart, ge = G.tstart + G.tlen;
                if (G.ov > 0) {  // negative-overlap gap: truth is continuous; the draft breaks at gs and repeats ov bases
                    if (te <= gs) break;
                    if (ts >= gs - G.ov) { shift += G.og + G.ov; continue; }   // aligns wholly on the right flank
                    hit = &G; break;
                }
                if (te <= gs) break;
                if (ts >= ge) { shift += (G.og - G.tlen); continue; }
                hit = &G; break;
            }
            int from = 0, to = L;      // aligned read interval [from,to)
            long dpos;                 // draft 0-based position of read base `from`
            if (!hit) { dpos = ts + shift; }
            else {
                if (!local) return a;  // end-to-end: a read touching a gap does not align
                long gs = hit->tstart, ge = hit->tstart + hit->tlen;
                int left = (int)std::max(0L, std::min((long)L, gs - ts));     // bases on the left flank
                int right = hit->ov > 0 ? (int)std::min((long)L, te - (gs - hit->ov)) : (int)std::max(0L, std::min((long)L, te - ge));
                if (left == 0 && right == 0) return a;
                if (left >= right) { from = 0; to = left; dpos = ts + shift; }
                else { from = L - right; to = L; dpos = hit->dstart + hit->og; }
            }
            // mismatches against the draft over the aligned interval
            int mm = 0, run = 0; std::string md;
            for (int j = from; j < to; j++) {
                char ref = S.draft[dpos + (j - from)];
                if (ref == rd_ref[j]) run++; else { md += std::to_string(run); md += ref; run = 0; mm++; }
            }
            md += std::to_string(run);
            int alen = to - from;
            if (local) { double score = 2.0 * (alen - mm) - 6.0 * mm; if (score < minscore) return a; a.as = (int)score; }
            else { if (6 * mm > 0.6 + 0.6 * L) return a; a.as = -6 * mm; }
            a.ok = true; a.pos = dpos + 1; a.nm = mm; a.md = md;
            if (from > 0) a.cigar += std::to_string(from) + "S";
            a.cigar += std::to_string(alen) + "M";
            if (to < L) a.cigar += std::to_string(L - to) + "S";
            a.rstart = dpos - from; a.rend = dpos + (L - from);
            return a;
        };
        for (long p = 0; p < npairs; p++, pair_id++) {
            int isz; do { isz = (int)std::lround(o.mu + o.sd * rng.normal()); } while (isz < L + 1 || isz > (long)S.truth.size() / 2);
            long f = (long)rng.below(S.truth.size() - isz + 1);
            bool m1fwd = rng.next() & 1;
            // left read (forward strand), right read (reverse strand), both in reference orientation
            std::string lr = S.truth.substr(f, L), rr = S.truth.substr(f + isz - L, L);
            // sequencing errors are made in sequencing orientation: position j of the right read is ref index L-1-j
            for (int j = 0; j < L; j++) if (rng.uni() < err[j]) { char c; do c = B[rng.next() >> 62]; while (c == lr[j]); lr[j] = c; }
            for (int j = 0; j < L; j++) if (rng.uni() < err[j]) { int x = L - 1 - j; char c; do c = B[rng.next() >> 62]; while (c == rr[x]); rr[x] = c; }
            if (o.nreadN > 0 && rng.below(o.nreadN) == 0) { lr[rng.below(L)] = 'N'; }
            std::string lq = qualf, rq(qualf.rbegin(), qualf.rend());   // QUAL in reference orientation
            // near-gap filter (bounds SAM size for large configs): keep pairs touching [gap-near, gap+near], plus model sample
            if (o.near >= 0) {
                bool nearg = false;
                for (auto& G : S.gaps) { long a = G.tstart - o.near, b = G.tstart + G.tlen + o.near; if (f < b && f + isz > a) { nearg = true; break; } if (G.tstart - o.near > f + isz) break; }
                if (!nearg) { if (far_cap >= 0 && far_written >= far_cap) continue; far_written++; }
            }
            char qn[64]; snprintf(qn, sizeof qn, "r%d_%ld", o.lib, pair_id);
            for (int mode = 0; mode < 2; mode++) {
                std::string& out = mode == 0 ? buf.a : buf.b; bool local = mode == 0; int X = local ? o.x1 : o.x2;
                Aln al = align(f, lr, local), ar = align(f + isz - L, rr, local);
                // mate1 = left/forward if m1fwd else right/reverse
                const Aln& a1 = m1fwd ? al : ar; const Aln& a2 = m1fwd ? ar : al;
                bool rev1 = !m1fwd, rev2 = m1fwd;
                const std::string& seq1ref = m1fwd ? lr : rr; const std::string& seq2ref = m1fwd ? rr : lr;
                const std::string& q1ref = m1fwd ? lq : rq;   const std::string& q2ref = m1fwd ? rq : lq;
                auto asseq = [&](const std::string& sref, bool rev) { if (!rev) return sref; std::string r(sref.size(), 'N'); for (size_t i = 0; i < sref.size(); i++) r[sref.size() - 1 - i] = comp(sref[i]); return r; };
                auto asq = [&](const std::string& qref, bool rev) { if (!rev) return qref; return std::string(qref.rbegin(), qref.rend()); };
                bool conc = false; long tl = 0;
                if (a1.ok && a2.ok) { long lo = std::min(a1.rstart, a2.rstart), hi = std::max(a1.rend, a2.rend); tl = hi - lo; conc = tl <= X && al.rstart <= ar.rstart; }
                for (int m = 0; m < 2; m++) {
                    const Aln& me = m == 0 ? a1 : a2; const Aln& mate = m == 0 ? a2 : a1; bool rev = m == 0 ? rev1 : rev2, mrev = m == 0 ? rev2 : rev1;
                    int flag = 1 | (m == 0 ? 64 : 128);
                    if (conc) flag |= 2;
                    if (!me.ok) flag |= 4; if (!mate.ok) flag |= 8;
                    if (me.ok && rev) flag |= 16; if (mate.ok && mrev) flag |= 32;
                    const std::string& sref = m == 0 ? seq1ref : seq2ref; const std::string& qref = m == 0 ? q1ref : q2ref;
                    if (me.ok) {
                        long t = 0; if (mate.ok) { t = (me.rstart <= mate.rstart && !(me.rstart == mate.rstart && m == 1)) ? tl : -tl; }
                        appendf(out, "%s\t%d\t%s\t%ld\t%d\t%s\t%s\t%ld\t%ld\t%s\t%s\tAS:i:%d\tXN:i:0\tXM:i:%d\tXO:i:0\tXG:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:%s\n",
                                qn, flag, S.name.c_str(), me.pos, 42, me.cigar.c_str(), mate.ok ? "=" : "=", mate.ok ? mate.pos : me.pos, t,
                                sref.c_str(), qref.c_str(), me.as, me.nm, me.nm, me.md.c_str(), conc ? "CP" : (mate.ok ? "DP" : "UP"));
                    } else {
                        std::string sq = asseq(sref, rev), qq = asq(qref, rev);
                        if (mate.ok) appendf(out, "%s\t%d\t%s\t%ld\t0\t*\t=\t%ld\t0\t%s\t%s\tYT:Z:UP\n", qn, flag, S.name.c_str(), mate.pos, mate.pos, sq.c_str(), qq.c_str());
                        else appendf(out, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\tYT:Z:UP\n", qn, flag, sq.c_str(), qq.c_str());
                    }
                }
            }
        }
    };
    long total_pairs = 0;
    if (!par) {
        long pair_id = 0, far_written = 0;
        for (int s = 0; s < o.nscaf; s++) {
            Buf buf;
            genReads(s, rng0, pair_id, far_written, o.model_pairs, buf);
            pair_id += pairsOf(s);
            fwrite(buf.a.data(), 1, buf.a.size(), s1); fwrite(buf.b.data(), 1, buf.b.size(), s2);
        }
        total_pairs = pair_id;
    } else {
        // one RNG stream and one share of the model pairs per scaffold; worker threads fill buffers, this thread writes them in order
        int nt = o.threads > 0 ? o.threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min(nt, o.nscaf));
        std::vector<long> base(o.nscaf + 1, 0);
        for (int s = 0; s < o.nscaf; s++) base[s + 1] = base[s] + pairsOf(s);
        total_pairs = base[o.nscaf];
        std::vector<Buf> bufs(o.nscaf); std::vector<char> ready(o.nscaf, 0);
        std::mutex mu; std::condition_variable cv; std::atomic<int> next(0); int written = 0;
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([&] {
            for (int s; (s = next++) < o.nscaf;) {
                { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return s < written + 2 * nt; }); }      // bound the text held in memory
                Rng r(scafSeed(s, 1)); long far = 0;
                const long cap = o.model_pairs >= 0 ? o.model_pairs / o.nscaf + (s < o.model_pairs % o.nscaf ? 1 : 0) : -1;
                genReads(s, r, base[s], far, cap, bufs[s]);
                { std::lock_guard<std::mutex> l(mu); ready[s] = 1; }
                cv.notify_all();
            }
        });
        for (int s = 0; s < o.nscaf; s++) {
            { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return ready[s] != 0; }); }
            fwrite(bufs[s].a.data(), 1, bufs[s].a.size(), s1); fwrite(bufs[s].b.data(), 1, bufs[s].b.size(), s2);
            Buf().a.swap(bufs[s].a); Buf().b.swap(bufs[s].b);
            { std::lock_guard<std::mutex> l(mu); written = s + 1; }
            cv.notify_all();
        }
        for (auto& t : th) t.join();
    }
    fclose(s1); fclose(s2);
    fprintf(stderr, "fbgen: %ld pairs, %d scaffolds, %d gaps -> %s\n", total_pairs, o.nscaf, o.ngaps, o.out.c_str());
    return 0;
}
