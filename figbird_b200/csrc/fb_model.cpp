// fb_model.cpp -- the learned insert-size / error-model tables.
//
// Host-side, once per run (the reference repeats this in every worker process).  The tables must be
// numerically identical to the reference's because every threshold downstream is applied to them, so the
// arithmetic below keeps the reference's types and evaluation order:
//   sufficient statistics   Figbird.cpp:846-921 (processMapping), :186-225 (updateInsertCounts),
//                           :291-487 (processErrorTypes), :255-275 (getLength)
//   normalisation           Figbird.cpp:497-844 (computeProbabilites)
//   cut-off / thresholds    Figbird.cpp:952-1376 (computeErrorProb, computeLikelihood), :7155-7200
// including its quirks (the MD string is compared with the CIGAR of an error-free read, so every read takes
// the error path, :297; soft clips count as insertions, :339; computeErrorProb does not split CIGARs at 'S', :988).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <algorithm>
#include <sys/mman.h>
#include <memory>
#include <thread>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "fb_host.h"

namespace fb {
namespace {

// strtok(3) semantics on a private buffer (the reference tokenises every SAM line with strtok).
struct Tokens {
    char* p;
    explicit Tokens(char* s) : p(s) {}
    char* next(const char* delim) {
        if (!p) return nullptr;
        p += strspn(p, delim);
        if (!*p) { p = nullptr; return nullptr; }
        char* tok = p;
        p += strcspn(p, delim);
        if (*p) { *p = '\0'; p++; } else p = nullptr;
        return tok;
    }
};

inline int baseIndex(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }

struct alignas(128) Stats {      // (one per parsing thread, side by side in a vector: no shared cache lines)
    int maxReadLength, MAX_INSERT_SIZE, maxInsertSize;
    std::vector<long> insertCounts;
    long errorTypes[5][5]; long baseCounts[5];
    std::vector<long> errorPos, inPos, inLengths, delPos, delLengths, readLengths;
    long discardedReads = 0, uniqueMappedReads = 0;
};

void bumpInsert(Stats& s, int index) {   // Figbird.cpp:186-225
    if (index <= 0) return;
    if (index < s.maxInsertSize) { s.insertCounts[index]++; return; }
    if (index > s.MAX_INSERT_SIZE) { s.discardedReads++; return; }
    int grown = std::max(s.maxInsertSize * 2, index);
    s.insertCounts.resize(grown, 1);
    s.insertCounts[index]++;
    s.maxInsertSize = grown;
}

// strncpy without the zero fill of the whole buffer (these run four times per SAM line)
static inline void copyBounded(char* dst, const char* src, size_t cap) { size_t n = strnlen(src, cap - 1); memcpy(dst, src, n); dst[n] = 0; }

// One CIGAR walk shared by the two passes.  `delims` is what strtok splits on (the two reference
// functions differ: "IDMS^\t\n " vs "IDM^\t\n ").  The op letter is looked up in the ORIGINAL string at the
// running offset, exactly as the reference does (Figbird.cpp:325-375, 988-1039).
template <class OnOp>
void walkCigar(const char* cigar, const char* delims, OnOp onOp) {
    char buf[1024];
    copyBounded(buf, cigar, sizeof buf);
    Tokens tk(buf);
    int consumed = 0;
    const size_t clen = strlen(cigar);
    for (char* t = tk.next(delims); t; t = tk.next(delims)) {
        int n = atoi(t);
        consumed += (int)strlen(t);
        char op = (size_t)consumed <= clen ? cigar[consumed] : '\0';
        onOp(op, n);
        consumed++;
    }
}

// The MD walk (Figbird.cpp:378-484, 1042-1149): calls onMismatch(from, readIndexBase) for each substitution.
template <class OnMis>
void walkMD(const char* md, const std::vector<int>& inserts, OnMis onMis) {
    char buf[1024];
    copyBounded(buf, md, sizeof buf);
    const unsigned long mdLength = strlen(md) - 5;
    Tokens tk(buf);
    tk.next(":"); tk.next(":");
    int index = 0; unsigned long totalLength = 0;
    for (char* t = tk.next("ACGTN^\t\n "); t; t = tk.next("ACGTN^\t\n ")) {
        totalLength += strlen(t);
        if (totalLength < mdLength) {
            char from = md[5 + totalLength];
            if (from == '^') {
                totalLength++;
                index += atoi(t);
                for (unsigned long i = totalLength; i < mdLength; i++) {
                    from = md[5 + totalLength];
                    if (from == 'A' || from == 'C' || from == 'G' || from == 'T' || from == 'N') totalLength++;
                    else break;
                }
            } else if (from == 'A' || from == 'C' || from == 'G' || from == 'T' || from == 'N') {
                totalLength++;
                index += atoi(t) + 1;
                int curIndex = 0;
                for (int i = 0; i < index && i < (int)inserts.size(); i++) curIndex += inserts[i];
                onMis(from, index, curIndex);
            } else break;
        }
    }
}

struct SamFields { char* qname; int flag; char* rname; int pos; char* cigar; int tlen; char* seq; char md[1000]; int nh; bool haveMd, haveNh; };

// Figbird.cpp:864-901: columns of a myout.sam line (the file has 10 columns, Preprocess.cpp:412-416)
bool splitMyout(char* line, SamFields& f, char* mdKeep /*persisting MD buffer*/) {
    Tokens tk(line);
    f.qname = tk.next("\t"); char* t = tk.next("\t"); if (!f.qname || !t) return false;
    f.flag = atoi(t);
    f.rname = tk.next("\t"); t = tk.next("\t"); if (!f.rname || !t) return false;
    f.pos = atoi(t);
    f.cigar = tk.next("\t"); t = tk.next("\t"); f.seq = tk.next("\t");
    if (!f.cigar || !t || !f.seq) return false;
    f.tlen = atoi(t);
    f.haveMd = f.haveNh = false;
    while ((t = tk.next("\t\n")) != nullptr) {
        if (t[0] == 'M' && t[1] == 'D') { copyBounded(mdKeep, t, 1000); f.haveMd = true; }
        else if (t[0] == 'I' && t[1] == 'H') { f.nh = atoi(t + 5); f.haveNh = true; }
    }
    return true;
}

// ---- fast path for regular lines.  Both passes spend their time tokenising; a line that matches the strict grammar below
// (what Preprocess writes for an ungapped, unclipped alignment: Preprocess.cpp:412-416) is parsed in place, without copies,
// and handed to the same callbacks as the generic strtok-style code, so the arithmetic is shared.  Anything else -- soft
// clips, indels, '^' in MD, missing or reordered tags, over-long lines -- takes the generic path.
//   qname \t flag \t rname \t pos \t <n>M \t [-]tlen \t seq \t qual \t MD:Z:[0-9ACGTN]+ \t IH:i:<n> \n
struct FastLine { const char* qname; int qlen; int flag; long rname; int tlen; const char* seq; int seqLen; const char* md; int mdLen; int nh; };

// what pass 2 needs of a regular line, kept from pass 1 (offsets from the line start)
struct LineRec { uint16_t ok, qlen, flag, seqOff, seqLen, mdOff, mdLen; int tlen; };

inline bool fastDigits(const char*& p, long& v) {      // 1..9 digits followed by a tab (consumed); value as atoi / atol give it
    const char* q = p; long x = 0;
    while (*q >= '0' && *q <= '9' && q - p < 10) x = x * 10 + (*q++ - '0');
    if (q == p || q - p > 9 || *q != '\t') return false;
    v = x; p = q + 1; return true;
}

inline bool fastSplit(const char* ln, const char* end, FastLine& f) {      // end: one past the last byte of the text
    const char* const e = (const char*)memchr(ln, '\n', (size_t)(end - ln));
    if (!e || e - ln > 1022 || e == ln) return false;      // no newline (last line), or fgets(1024) would have cut it: generic path
    auto field = [&](const char* p) -> const char* { return (const char*)memchr(p, '\t', (size_t)(e - p)); };      // the tab that ends the field at p
    const char* p = ln;
    if (*p == '\t') return false;                           // strtok would skip leading tabs
    const char* t = field(p); if (!t) return false;
    f.qname = p; f.qlen = (int)(t - p); p = t + 1;
    long v;
    if (!fastDigits(p, v)) return false; f.flag = (int)v;
    if (!fastDigits(p, v)) return false; f.rname = v;
    if (!fastDigits(p, v)) return false;                         // pos
    if (!(*p >= '0' && *p <= '9')) return false;
    while (*p >= '0' && *p <= '9') p++;
    if (p[0] != 'M' || p[1] != '\t') return false;               // CIGAR = one M operation
    p += 2;
    const bool neg = (*p == '-'); if (neg) p++;
    if (!fastDigits(p, v)) return false; f.tlen = neg ? -(int)v : (int)v;
    t = field(p); if (!t || t == p) return false;
    f.seq = p; f.seqLen = (int)(t - p); p = t + 1;
    t = field(p); if (!t || t == p) return false;               // qual
    p = t + 1;
    if (!(p[0] == 'M' && p[1] == 'D' && p[2] == ':' && p[3] == 'Z' && p[4] == ':')) return false;
    f.md = p; p += 5;
    const char* body = p;
    for (;; p++) { const char c = *p; if ((c >= '0' && c <= '9') || c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N') continue; break; }
    if (p == body || *p != '\t') return false;
    f.mdLen = (int)(p - f.md); if (f.mdLen > 990) return false; p++;
    if (!(p[0] == 'I' && p[1] == 'H' && p[2] == ':' && p[3] == 'i' && p[4] == ':')) return false;
    p += 5;
    const char* d = p; long nh = 0;
    while (*p >= '0' && *p <= '9' && p - d < 10) nh = nh * 10 + (*p++ - '0');
    if (p == d || p - d > 9 || p != e) return false;
    f.nh = (int)nh;
    return true;
}

// walkMD on the body of a regular MD tag (digits and ACGTN only), no insertions in the read: same state machine, in place.
template <class OnMis>
inline void fastWalkMD(const char* body, int bl, OnMis onMis) {
    auto isDelim = [](char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; };
    int index = 0, pos = 0; long totalLength = 0;
    for (;;) {
        while (pos < bl && isDelim(body[pos])) pos++;
        if (pos >= bl) break;
        int val = 0; const int t0 = pos; bool big = false;
        while (pos < bl && !isDelim(body[pos])) { if (pos - t0 < 9) val = val * 10 + (body[pos] - '0'); else big = true; pos++; }
        if (big) val = atoi(std::string(body + t0, body + pos).c_str());
        totalLength += pos - t0;
        if (totalLength < bl) {
            const char from = body[totalLength];
            if (isDelim(from)) { totalLength++; index += val + 1; onMis(from, index, 0); }
            else break;
        }
    }
}

}  // namespace

bool learnModel(const Args& a, const Scaffolds& sc, Model& m, std::string& err, const std::function<bool()>& waitScaffolds) {
    // ---- stat.txt (Figbird.cpp:7084-7093)
    int maxIns = 0;
    {
        FILE* f = fopen((a.tmpDir + "stat.txt").c_str(), "r");
        if (!f) { err = "Can't open " + a.tmpDir + "stat.txt"; return false; }
        if (fscanf(f, "%ld %ld %d %d", &m.totalCount, &m.unCount, &m.maxReadLength, &maxIns) != 4) { fclose(f); err = "bad stat.txt"; return false; }
        fclose(f);
    }
    Stats s;
    s.maxReadLength = m.maxReadLength;
    s.MAX_INSERT_SIZE = maxIns > 20000 ? maxIns : 20000;
    s.maxInsertSize = s.MAX_INSERT_SIZE;
    s.insertCounts.assign(s.maxInsertSize, 1);
    for (auto& r : s.errorTypes) for (auto& v : r) v = 1;
    for (auto& v : s.baseCounts) v = 1;
    const int RL = m.maxReadLength;
    s.errorPos.assign(RL, 1); s.inPos.assign(RL, 1); s.inLengths.assign(RL, 1); s.delPos.assign(RL, 1); s.delLengths.assign(RL, 1); s.readLengths.assign(RL, 0);
    const double inputMean = a.setInputMean == 1 ? (double)a.insertSizeMean : 0.0;   // Figbird.cpp:6973

    const bool timing = getenv("FIGBIRD_MODEL_TIMING") != nullptr;
    const bool useFast = getenv("FIGBIRD_MODEL_GENERIC") == nullptr;      // tests compare the two parsers
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto tprev = tnow();
    auto lap = [&](const char* what) { if (timing) { auto t = tnow(); fprintf(stderr, "learnModel %-10s %.3f s\n", what, std::chrono::duration<double>(t - tprev).count()); tprev = t; } };
    FILE* mf = fopen(a.myout.c_str(), "r");
    if (!mf) { err = "Can't open map file"; return false; }
    // whole file in memory once; both passes walk the same lines (the reference reads it twice per worker).  The read and
    // the search for line starts are cut into blocks over the host threads.
    std::vector<char> textOwned;
    std::vector<char*> lines;
    struct Mapping { void* p = nullptr; size_t len = 0; ~Mapping() { if (p) munmap(p, len); } } mapping;
    char* textBase = nullptr;
    const char* textEnd = nullptr;
    {
        fseek(mf, 0, SEEK_END); const long n = ftell(mf); fseek(mf, 0, SEEK_SET);
        int rt = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("FIGBIRD_HOST_THREADS")) rt = atoi(e);
        rt = std::max(1, std::min(rt, (int)(n / (8 << 20)) + 1));
        const int fd = fileno(mf);
        long total = n;
        struct SetEnd { const char*& e; char*& b; long n; ~SetEnd() { e = b ? b + n : nullptr; } } setEnd{textEnd, textBase, n};
        // map the page cache instead of copying it; the zero tail of the last page is the terminating NUL
        // (a file that ends exactly on a page boundary takes the copying path)
        if (n > 0 && (n % 4096) != 0) {
            void* mp = mmap(nullptr, (size_t)n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (mp != MAP_FAILED) { mapping.p = mp; mapping.len = (size_t)n; textBase = (char*)mp; }
        }
        std::vector<std::thread> th;
        if (!textBase) {
            textOwned.resize((size_t)n + 1);
            std::vector<long> got(rt, 0);
            for (int t = 0; t < rt; t++) th.emplace_back([&, t] {
                const long lo = n * t / rt, hi = n * (t + 1) / rt;
                long off = lo;
                while (off < hi) { ssize_t k = pread(fd, textOwned.data() + off, (size_t)(hi - off), off); if (k <= 0) break; off += k; }
                got[t] = off - lo;
            });
            for (auto& t : th) t.join();
            bool shortRead = false;
            for (int t = 0; t < rt; t++) if (got[t] != n * (t + 1) / rt - n * t / rt) shortRead = true;
            if (shortRead) { fclose(mf); err = "Can't read map file"; return false; }
            textOwned[n] = 0;
            textBase = textOwned.data();
        }
        fclose(mf);
        std::vector<std::vector<char*>> part(rt);
        struct { char* d; char* data() const { return d; } } text{textBase};
        // line starts: position 0 and every position after a '\n' (a trailing '\n' starts no line)
        th.clear();
        for (int t = 0; t < rt; t++) th.emplace_back([&, t] {
            const long lo = total * t / rt, hi = total * (t + 1) / rt;
            std::vector<char*>& v = part[t];
            v.reserve((size_t)(hi - lo) / 150 + 16);
            if (lo == 0 && total > 0) v.push_back(text.data());
            const char* base = text.data();
            for (const char* p = (const char*)memchr(base + lo, '\n', (size_t)(hi - lo)); p; p = (const char*)memchr(p + 1, '\n', (size_t)(base + hi - (p + 1)))) {
                if (p + 1 < base + total) v.push_back(const_cast<char*>(p + 1));
                if (p + 1 >= base + hi) break;
            }
        });
        for (auto& t : th) t.join();
        size_t cnt = 0; for (auto& v : part) cnt += v.size();
        lines.reserve(cnt);
        for (auto& v : part) lines.insert(lines.end(), v.begin(), v.end());
    }
    // NB: lines keep their '\n'; tokenisers treat it as a delimiter where the reference does.
    lap("read+split");
    if (waitScaffolds && !waitScaffolds()) { err = "Can't open contig file"; return false; }

    // ---- pass 1: processMapping.  The statistics are integer counts, so the file is cut into blocks that are
    // counted on separate threads and summed -- identical to the sequential result as long as every line carries its
    // own MD/IH tags and no insert size hits the histogram-growth corner (Figbird.cpp:203); otherwise redo it serially.
    // (uninitialised on purpose: pass 1 visits every line and sets its record's `ok` first; pages are first touched by the
    // thread that owns the block)
    std::unique_ptr<LineRec[]> recs(useFast ? new LineRec[lines.size() + 1] : nullptr);
    auto pass1Line = [&](Stats& st, char* mdKeep, std::vector<char>& scratch, char* ln, bool& irregular, LineRec* rec) {
        if (rec) rec->ok = 0;
        if (ln[0] == '@') return;
        FastLine fl;
        if (useFast && fastSplit(ln, textEnd, fl)) {
            if (rec && fl.flag < 65536) { rec->ok = 1; rec->qlen = (uint16_t)fl.qlen; rec->flag = (uint16_t)fl.flag; rec->seqOff = (uint16_t)(fl.seq - ln); rec->seqLen = (uint16_t)fl.seqLen;
                                          rec->mdOff = (uint16_t)(fl.md - ln); rec->mdLen = (uint16_t)fl.mdLen; rec->tlen = fl.tlen; }
            memcpy(mdKeep, fl.md, (size_t)fl.mdLen); mdKeep[fl.mdLen] = 0;      // (a later irregular line may inherit it, as in the generic path)
            if (fl.nh != 1) return;
            if (fl.rname >= 0 && fl.rname < (long)sc.seq.size() && (double)sc.seq[fl.rname].size() > inputMean) {
                if (fl.tlen >= st.maxInsertSize && fl.tlen <= st.MAX_INSERT_SIZE) irregular = true;
                bumpInsert(st, fl.tlen);
            }
            const int strand = (fl.flag & 16) >> 4;
            const int readLength = fl.seqLen;
            {   // base composition of the read: 16 bases per step (byte compares, per-byte counters summed at the end)
                long nA = 0, nC = 0, nG = 0, nT = 0;
                int i = 0;
#if defined(__SSE2__)
                const __m128i vA = _mm_set1_epi8('A'), vC = _mm_set1_epi8('C'), vG = _mm_set1_epi8('G'), vT = _mm_set1_epi8('T');
                while (i + 16 <= readLength) {
                    __m128i aA = _mm_setzero_si128(), aC = aA, aG = aA, aT = aA;
                    const int stop = std::min(readLength - 15, i + 16 * 255);      // a byte counter holds 255 steps
                    for (; i < stop; i += 16) {
                        const __m128i v = _mm_loadu_si128((const __m128i*)(fl.seq + i));
                        aA = _mm_sub_epi8(aA, _mm_cmpeq_epi8(v, vA)); aC = _mm_sub_epi8(aC, _mm_cmpeq_epi8(v, vC));
                        aG = _mm_sub_epi8(aG, _mm_cmpeq_epi8(v, vG)); aT = _mm_sub_epi8(aT, _mm_cmpeq_epi8(v, vT));
                    }
                    const __m128i z = _mm_setzero_si128();
                    auto hsum = [&](__m128i a) { const __m128i sd = _mm_sad_epu8(a, z); return (long)_mm_cvtsi128_si64(sd) + (long)_mm_cvtsi128_si64(_mm_srli_si128(sd, 8)); };
                    nA += hsum(aA); nC += hsum(aC); nG += hsum(aG); nT += hsum(aT);
                }
#endif
                for (; i < readLength; i++) { const char c = fl.seq[i]; nA += c == 'A'; nC += c == 'C'; nG += c == 'G'; nT += c == 'T'; }
                st.baseCounts[0] += nA; st.baseCounts[1] += nC; st.baseCounts[2] += nG; st.baseCounts[3] += nT; st.baseCounts[4] += readLength - nA - nC - nG - nT;
            }
            if (readLength > RL) { st.uniqueMappedReads++; return; }
            st.readLengths[readLength - 1]++;
            fastWalkMD(fl.md + 5, fl.mdLen - 5, [&](char from, int idx, int cur) {
                int ri = idx - 1 + cur;
                char to = (ri >= 0 && ri < readLength) ? fl.seq[ri] : 'N';
                int at = strand == 0 ? ri : readLength - idx - cur;
                if (at >= 0 && at < RL) st.errorPos[at]++;
                int fi = baseIndex(from), ti = baseIndex(to);
                if (fi != ti) st.errorTypes[fi][ti]++;
            });
            st.uniqueMappedReads++;
            return;
        }
        size_t len = strcspn(ln, "\n"); if (len > 1022) len = 1022;   // fgets(1024)
        scratch.assign(ln, ln + len + 1); scratch[len] = '\n'; scratch.push_back(0);
        SamFields f;
        if (!splitMyout(scratch.data(), f, mdKeep)) { irregular = true; return; }
        if (!f.haveMd || !f.haveNh) irregular = true;
        if (!(f.nh == 1 && mdKeep[5] != '^')) return;
        long contigNo = atol(f.rname);
        if (contigNo >= 0 && contigNo < (long)sc.seq.size() && (double)sc.seq[contigNo].size() > inputMean) {
            if (f.tlen >= st.maxInsertSize && f.tlen <= st.MAX_INSERT_SIZE) irregular = true;
            bumpInsert(st, f.tlen);
        }
        // processErrorTypes
        const int strand = (f.flag & 16) >> 4;
        int readLength = 0;
        for (const char* c = f.seq; *c; c++, readLength++) st.baseCounts[baseIndex(*c)]++;
        if (readLength < 1 || readLength > RL) { st.uniqueMappedReads++; return; }
        st.readLengths[readLength - 1]++;
        static thread_local std::vector<int> inserts;
        inserts.assign(readLength, 0);
        int index = 0, curIndex = 0;
        walkCigar(f.cigar, "IDMS^\t\n ", [&](char op, int n) {
            if (op == 'M') { index += n; curIndex += n; }
            else if (op == 'I' || op == 'S') {
                int at = strand == 0 ? index : readLength - index - 1;
                if (at >= 0 && at < RL) st.inPos[at]++;
                if (n >= 1 && n <= RL) st.inLengths[n - 1]++;
                if (curIndex >= 0 && curIndex < readLength) inserts[curIndex] = n;
                index += n;
            } else if (op == 'D') {
                int at = strand == 0 ? index : readLength - index - 1;
                if (at >= 0 && at < RL) st.delPos[at]++;
                if (n >= 1 && n <= RL) st.delLengths[n - 1]++;
            }
        });
        walkMD(mdKeep, inserts, [&](char from, int idx, int cur) {
            int ri = idx - 1 + cur;
            char to = (ri >= 0 && ri < readLength) ? f.seq[ri] : 'N';
            int at = strand == 0 ? ri : readLength - idx - cur;
            if (at >= 0 && at < RL) st.errorPos[at]++;
            int fi = baseIndex(from), ti = baseIndex(to);
            if (fi != ti) st.errorTypes[fi][ti]++;
        });
        st.uniqueMappedReads++;
    };
    auto zeroStats = [&](Stats& z) {
        z = s;
        std::fill(z.insertCounts.begin(), z.insertCounts.end(), 0);
        for (auto& r : z.errorTypes) for (auto& v : r) v = 0;
        for (auto& v : z.baseCounts) v = 0;
        for (auto* v : {&z.errorPos, &z.inPos, &z.inLengths, &z.delPos, &z.delLengths, &z.readLengths}) std::fill(v->begin(), v->end(), 0);
        z.discardedReads = 0; z.uniqueMappedReads = 0;
    };
    int nThreads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("FIGBIRD_HOST_THREADS")) nThreads = atoi(e);
    size_t blockLines = 20000;      // lines per thread below which threading does not pay (FIGBIRD_MODEL_BLOCK: tests force small blocks)
    if (const char* e = getenv("FIGBIRD_MODEL_BLOCK")) blockLines = (size_t)std::max(1, atoi(e));
    nThreads = std::max(1, std::min(nThreads, (int)(lines.size() / blockLines) + 1));
    bool irregularAny = false;
    if (nThreads > 1) {
        std::vector<Stats> part(nThreads);
        std::vector<char> irr(nThreads, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < nThreads; t++) th.emplace_back([&, t] {
            zeroStats(part[t]);
            std::vector<char> scratch(2048); char mdKeep[1000]; mdKeep[0] = 0; bool ir = false;
            const size_t lo = lines.size() * t / nThreads, hi = lines.size() * (t + 1) / nThreads;
            for (size_t i = lo; i < hi; i++) pass1Line(part[t], mdKeep, scratch, lines[i], ir, useFast ? &recs[i] : nullptr);
            irr[t] = ir;
        });
        for (auto& t : th) t.join();
        for (int t = 0; t < nThreads; t++) irregularAny |= (irr[t] != 0);
        if (!irregularAny) {
            for (int t = 0; t < nThreads; t++) {
                const Stats& q = part[t];
                for (size_t i = 0; i < s.insertCounts.size(); i++) s.insertCounts[i] += q.insertCounts[i];
                for (int i = 0; i < 5; i++) { s.baseCounts[i] += q.baseCounts[i]; for (int j = 0; j < 5; j++) s.errorTypes[i][j] += q.errorTypes[i][j]; }
                for (int i = 0; i < RL; i++) { s.errorPos[i] += q.errorPos[i]; s.inPos[i] += q.inPos[i]; s.inLengths[i] += q.inLengths[i]; s.delPos[i] += q.delPos[i]; s.delLengths[i] += q.delLengths[i]; s.readLengths[i] += q.readLengths[i]; }
                s.discardedReads += q.discardedReads; s.uniqueMappedReads += q.uniqueMappedReads;
            }
        }
    }
    if (nThreads <= 1 || irregularAny) {
        std::vector<char> scratch(2048); char mdKeep[1000]; mdKeep[0] = 0; bool ir = false;
        for (size_t i = 0; i < lines.size(); i++) pass1Line(s, mdKeep, scratch, lines[i], ir, useFast ? &recs[i] : nullptr);
    }

    lap("pass1");
    // ---- computeProbabilites (Figbird.cpp:497-844); only the quantities used downstream are kept
    double baseErrorRates[5];
    for (int i = 0; i < 5; i++) {
        int errorCount = 0;
        for (int j = 0; j < 5; j++) errorCount += s.errorTypes[i][j];
        for (int j = 0; j < 5; j++) m.errorTypeProbs[i][j] = (double)s.errorTypes[i][j] / errorCount;
        baseErrorRates[i] = errorCount / (double)s.baseCounts[i];
    }
    { double sum = 0; for (int i = 0; i < 4; i++) sum += baseErrorRates[i]; for (int i = 0; i < 4; i++) baseErrorRates[i] = 4 * baseErrorRates[i] / sum; baseErrorRates[4] = 1; }
    for (int i = RL - 1; i > 0; i--) s.readLengths[i - 1] = s.readLengths[i] + s.readLengths[i - 1];
    m.errorPosDist.resize(RL); m.inPosDist.resize(RL); m.delPosDist.resize(RL);
    std::vector<double> inLengthDist(RL), delLengthDist(RL);
    for (int i = 0; i < RL; i++) m.errorPosDist[i] = (double)s.errorPos[i] / s.readLengths[i];
    for (int i = 0; i < RL; i++) m.inPosDist[i] = (double)s.inPos[i] / s.readLengths[i];
    { int c = 0; for (int i = 0; i < RL; i++) c += s.inLengths[i]; for (int i = 0; i < RL; i++) inLengthDist[i] = (double)s.inLengths[i] / c; }
    for (int i = 0; i < RL; i++) m.delPosDist[i] = (double)s.delPos[i] / s.readLengths[i];
    { int c = 0; for (int i = 0; i < RL; i++) c += s.delLengths[i]; for (int i = 0; i < RL; i++) delLengthDist[i] = (double)s.delLengths[i] / c; }

    const int MI = s.maxInsertSize;
    m.maxInsertSize = MI;
    std::vector<double> insertLengthDist(MI);
    long insCount = s.discardedReads;
    double sum = 0;
    for (int i = 0; i < MI; i++) { insCount += (s.insertCounts[i] - 1); sum += i * (s.insertCounts[i] - 1); }
    m.insertSizeMean = sum / insCount;
    const double mean = m.insertSizeMean;
    sum = 0;
    for (int i = 0; i < MI; i++) { insertLengthDist[i] = (double)s.insertCounts[i] / insCount; sum += (s.insertCounts[i] - 1) * (mean - i) * (mean - i); }
    std::vector<double> noErrorProbs(RL);
    { double p = 1.0; for (int i = 0; i < RL; i++) { p *= (1 - m.errorPosDist[i] - m.inPosDist[i] - m.delPosDist[i]); noErrorProbs[i] = p; } }

    const int W = 12;   // windowSize, Figbird.cpp:89
    m.insertPdfSmoothed.assign(MI, 0.0);
    {
        std::vector<double>& sm = m.insertPdfSmoothed;
        double windowSum = 0;
        for (int i = 0; i < W; i++) sm[i] = insertLengthDist[i];
        for (int i = 0; i < 2 * W + 1; i++) windowSum += insertLengthDist[i];
        sm[W] = windowSum / (2 * W + 1);
        for (int i = W + 1; i < MI - W; i++) { windowSum -= insertLengthDist[i - W - 1]; windowSum += insertLengthDist[i + W]; sm[i] = windowSum / (2 * W + 1); }
        for (int i = MI - W; i < MI; i++) sm[i] = insertLengthDist[i];
        for (int i = 0; i < MI; i++) sm[i] = sm[i] - 1 / (double)(insCount) + (1 / (double)MI) / (double)(insCount + 1);
    }
    {   // one-sided SDs about the mean (Figbird.cpp:785-802)
        double insertSum = 0, insertCount = 0;
        for (int i = mean + 1; i < MI; i++) { insertSum = insertSum + (s.insertCounts[i] - 1) * (i - mean) * (i - mean); insertCount += (s.insertCounts[i] - 1); }
        m.rightSD = sqrt(insertSum / insertCount);
        insertSum = 0; insertCount = 0;
        for (int i = std::max((int)(mean - 10 * m.rightSD), 0); i < mean; i++) { insertSum = insertSum + (s.insertCounts[i] - 1) * (mean - i) * (mean - i); insertCount += (s.insertCounts[i] - 1); }
        m.leftSD = sqrt(insertSum / insertCount);
    }
    m.uniqueMappedReads = s.uniqueMappedReads;

    // ---- pass 2: computeLikelihood -> gapProbs histogram (Figbird.cpp:1156-1376)
    long totalContigLength = 0;
    for (auto& q : sc.seq) totalContigLength += (long)q.size();
    std::vector<long> effectiveLengths(MI, -1);
    effectiveLengths[0] = totalContigLength;
    auto effLen = [&](int insertSize) -> long {
        if (insertSize < 0) return effectiveLengths[0];
        auto compute = [&]() { long e = 0; for (auto& q : sc.seq) if ((long)q.size() >= insertSize) e += ((long)q.size() - insertSize + 1); return e; };
        if (insertSize >= MI) return compute();
        if (effectiveLengths[insertSize] == -1) effectiveLengths[insertSize] = compute();
        return effectiveLengths[insertSize];
    };
    auto misEp = [&](long double& ep, char from, int idx, int cur, const char* read, int readLength, int strand) {
        int ri = idx - 1 + cur;
        char to = (ri >= 0 && ri < readLength) ? read[ri] : '\0';
        int i = strand == 0 ? ri : readLength - idx - cur;
        if (i >= 0 && i < RL) ep = ep * m.errorPosDist[i] / (1 - m.errorPosDist[i] - m.inPosDist[i] - m.delPosDist[i]);
        int fi = baseIndex(from), ti = baseIndex(to);
        if (fi != ti) ep *= baseErrorRates[fi] * m.errorTypeProbs[fi][ti];
    };
    auto errorProb = [&](const char* cigar, const char* md, const char* read, int strand) -> long double {
        const unsigned long readLength = strlen(read);
        long double ep = (readLength >= 1 && (int)readLength <= RL) ? noErrorProbs[readLength - 1] : 0;
        if (md[5] == '^') return ep;
        static thread_local std::vector<int> inserts;
        inserts.assign(readLength, 0);
        int index = 0, curIndex = 0;
        walkCigar(cigar, "IDM^\t\n ", [&](char op, int n) {
            if (op == 'M') { index += n; curIndex += n; }
            else if (op == 'I') {
                unsigned long i = strand == 0 ? (unsigned long)index : readLength - index - 1;
                if (i < (unsigned long)RL && n >= 1 && n <= RL)
                    ep = ep * m.inPosDist[i] * inLengthDist[n - 1] / (1 - m.errorPosDist[i] - m.inPosDist[i] - m.delPosDist[i]);
                if (curIndex >= 0 && curIndex < (int)readLength) inserts[curIndex] = n;
                index += n;
            } else if (op == 'D') {
                unsigned long i = strand == 0 ? (unsigned long)index : readLength - index - 1;
                if (i < (unsigned long)RL && n >= 1 && n <= RL)
                    ep = ep * m.delPosDist[i] * delLengthDist[n - 1] / (1 - m.errorPosDist[i] - m.inPosDist[i] - m.delPosDist[i]);
            }
        });
        walkMD(md, inserts, [&](char from, int idx, int cur) { misEp(ep, from, idx, cur, read, (int)readLength, strand); });
        return ep;
    };
    auto errorProbFast = [&](const FastLine& fl) -> long double {      // the same for a regular line (one M operation, no '^')
        const int readLength = fl.seqLen, strand = (fl.flag & 16) >> 4;
        long double ep = (readLength >= 1 && readLength <= RL) ? noErrorProbs[readLength - 1] : 0;
        fastWalkMD(fl.md + 5, fl.mdLen - 5, [&](char from, int idx, int cur) { misEp(ep, from, idx, cur, fl.seq, readLength, strand); });
        return ep;
    };
    // one line of a pair: read name, template length and error probability (fast or generic parse)
    struct P2Line { const char* q; int qn; int tlen; long double e; };
    auto parseP2 = [&](size_t li, std::vector<char>& b, char* mdKeep, P2Line& o) -> bool {
        char* const l = lines[li];
        FastLine fl;
        if (useFast && recs[li].ok) {
            const LineRec& r = recs[li];
            fl.qname = l; fl.qlen = r.qlen; fl.flag = r.flag; fl.tlen = r.tlen; fl.seq = l + r.seqOff; fl.seqLen = r.seqLen; fl.md = l + r.mdOff; fl.mdLen = r.mdLen;
            memcpy(mdKeep, fl.md, (size_t)fl.mdLen); mdKeep[fl.mdLen] = 0;
            o.q = fl.qname; o.qn = fl.qlen; o.tlen = fl.tlen; o.e = errorProbFast(fl);
            return true;
        }
        size_t n = strcspn(l, "\n");
        if (n > 1022) n = 1022;
        b.assign(l, l + n + 1); b[n] = '\n'; b.push_back(0);
        SamFields f;
        if (!splitMyout(b.data(), f, mdKeep)) return false;
        o.q = l + (f.qname - b.data()); o.qn = (int)strlen(f.qname); o.tlen = f.tlen;
        o.e = errorProb(f.cigar, mdKeep, f.seq, (f.flag & 16) >> 4);
        return true;
    };

    long gapProbs[1000] = {0};
    lap("tables");
    if (nThreads <= 1 || irregularAny) {
        const char* p1 = "*"; int pn1 = 1; const char* p2 = "*"; int pn2 = 1;
        auto same = [](const char* a, int an, const char* b, int bn) { return an == bn && memcmp(a, b, an) == 0; };
        long double tempProb = 0, gapProb = 0;
        char md1[1000], md2[1000]; md1[0] = md2[0] = 0;
        std::vector<char> b1(2048), b2(2048);
        size_t li = 0;
        while (li < lines.size()) {
            const size_t i1 = li++;
            if (lines[i1][0] == '@') continue;
            if (li >= lines.size()) break;
            const size_t i2 = li++;
            P2Line x1, x2;
            if (!parseP2(i1, b1, md1, x1) || !parseP2(i2, b2, md2, x2)) continue;
            int insertSize = std::max(x1.tlen, x2.tlen);
            long double insertSizeProb = 0;
            if (insertSize >= 0 && insertSize < MI) insertSizeProb = insertLengthDist[insertSize];
            if (insertSizeProb == 0) insertSizeProb = 1 / (double)s.uniqueMappedReads;
            long eff = effLen(insertSize);
            long double prob = (1 / (long double)(eff)) * insertSizeProb * x1.e * x2.e;
            if (same(p1, pn1, x1.q, x1.qn) && same(p2, pn2, x2.q, x2.qn)) {
                if (tempProb < prob) { tempProb = prob; gapProb = x2.e; }
            } else if (!same(p1, pn1, "*", 1) && !same(p2, pn2, "*", 1)) {
                int gapIndex = -std::log10(gapProb);
                gapIndex++;
                if (gapIndex < 1000 && gapIndex >= 0) gapProbs[gapIndex]++; else gapProbs[999]++;
                tempProb = prob; gapProb = x2.e;
            } else { tempProb = prob; gapProb = x2.e; }
            p1 = x1.q; pn1 = x1.qn; p2 = x2.q; pn2 = x2.qn;
        }
    }
    else {
        // same histogram, pairs evaluated on threads (each pair's probabilities depend only on its two lines), then the
        // sequential read-name grouping of Figbird.cpp:1290-1349 over the stored results
        struct PairRes { const char* q1; const char* q2; int n1, n2; long double e2, prob; bool ok; };
        // getEffectiveLength (Figbird.cpp:923-950) for every insert size of the histogram, once: sum over scaffolds at least that long
        std::vector<long> effTab((size_t)MI + 1);
        {
            std::vector<long> lens; for (auto& q : sc.seq) lens.push_back((long)q.size());
            std::sort(lens.begin(), lens.end());
            size_t first = 0; long sumLen = 0; for (long v : lens) sumLen += v;      // scaffolds [first, end) have length >= t
            for (long t = 0; t <= MI; t++) {
                while (first < lens.size() && lens[first] < t) { sumLen -= lens[first]; first++; }
                effTab[(size_t)t] = sumLen - (long)(lens.size() - first) * (t - 1);
            }
        }
        std::vector<std::pair<size_t, size_t>> pairs;
        pairs.reserve(lines.size() / 2 + 1);
        for (size_t li = 0; li < lines.size();) {
            const size_t i1 = li++;
            if (lines[i1][0] == '@') continue;
            if (li >= lines.size()) break;
            pairs.emplace_back(i1, li++);
        }
        const size_t nPairs = pairs.size();
        std::unique_ptr<PairRes[]> res(new PairRes[nPairs + 1]);      // (uninitialised: every element is written by its thread)
        std::vector<std::thread> th;
        for (int t = 0; t < nThreads; t++) th.emplace_back([&, t] {
            char md1[1000], md2[1000]; md1[0] = md2[0] = 0;
            std::vector<char> b1(2048), b2(2048);
            const size_t lo = pairs.size() * t / nThreads, hi = pairs.size() * (t + 1) / nThreads;
            for (size_t i = lo; i < hi; i++) {
                PairRes& r = res[i];
                P2Line x1, x2;
                r.ok = parseP2(pairs[i].first, b1, md1, x1) && parseP2(pairs[i].second, b2, md2, x2);
                if (!r.ok) continue;
                r.q1 = x1.q; r.n1 = x1.qn; r.q2 = x2.q; r.n2 = x2.qn;
                int insertSize = std::max(x1.tlen, x2.tlen);
                long double insertSizeProb = 0;
                if (insertSize >= 0 && insertSize < MI) insertSizeProb = insertLengthDist[insertSize];
                if (insertSizeProb == 0) insertSizeProb = 1 / (double)s.uniqueMappedReads;
                long double e1 = x1.e;
                r.e2 = x2.e;
                long eff;
                if (insertSize < 0) eff = totalContigLength;
                else if (insertSize < (int)effTab.size()) eff = effTab[insertSize];
                else { eff = 0; for (auto& q : sc.seq) if ((long)q.size() >= insertSize) eff += ((long)q.size() - insertSize + 1); }
                r.prob = (1 / (long double)(eff)) * insertSizeProb * e1 * r.e2;
            }
        });
        for (auto& t : th) t.join();
        // The grouping of Figbird.cpp:1290-1349 over the stored results: a group = a maximal run of consecutive pairs with the same
        // two read names; when the next group begins, the e2 of the group's most probable pair (first maximum) is histogrammed,
        // unless one of the group's names is "*" (the initial state of the reference's loop); the last group never is.  Groups are
        // independent, so every thread takes the groups that start in its block (it may read past the block's end).
        auto same = [](const char* a, int an, const char* b, int bn) { return an == bn && memcmp(a, b, an) == 0; };
        std::vector<uint32_t> okIdx; okIdx.reserve(nPairs);
        for (size_t i = 0; i < nPairs; i++) if (res[i].ok) okIdx.push_back((uint32_t)i);
        const size_t M = okIdx.size();
        std::vector<std::vector<long>> hist(nThreads, std::vector<long>(1000 + 16, 0));
        th.clear();
        for (int t = 0; t < nThreads; t++) th.emplace_back([&, t] {
            long* const h = hist[t].data();
            const size_t lo = M * t / nThreads, hi = M * (t + 1) / nThreads;
            for (size_t i = lo; i < hi; i++) {
                const PairRes& r = res[okIdx[i]];
                if (i > 0) { const PairRes& q = res[okIdx[i - 1]]; if (same(q.q1, q.n1, r.q1, r.n1) && same(q.q2, q.n2, r.q2, r.n2)) continue; }
                long double tempProb = r.prob, gapProb = r.e2;
                size_t j = i + 1;
                for (; j < M; j++) {
                    const PairRes& x = res[okIdx[j]];
                    if (!(same(r.q1, r.n1, x.q1, x.n1) && same(r.q2, r.n2, x.q2, x.n2))) break;
                    if (tempProb < x.prob) { tempProb = x.prob; gapProb = x.e2; }
                }
                if (j < M && !same(r.q1, r.n1, "*", 1) && !same(r.q2, r.n2, "*", 1)) {
                    int gapIndex = -std::log10(gapProb);
                    gapIndex++;
                    if (gapIndex < 1000 && gapIndex >= 0) h[gapIndex]++; else h[999]++;
                }
            }
        });
        for (auto& t : th) t.join();
        for (int t = 0; t < nThreads; t++) for (int i = 0; i < 1000; i++) gapProbs[i] += hist[t][i];
    }
    lap("pass2");
    {   // Figbird.cpp:7155-7178
        long gapProbSum = 0; for (int i = 0; i < 1000; i++) gapProbSum += gapProbs[i];
        long gapProbCount = 0; const double value = .8;
        m.gapProbCutOff = 0;
        for (int i = 0; i < 1000; i++) { gapProbCount += gapProbs[i]; if (gapProbCount >= value * gapProbSum) { m.gapProbCutOff = i; break; } }
    }
    // ---- insert thresholds (Figbird.cpp:7188-7200)
    m.insertThresholdMin = std::max((int)(mean - 3 * m.leftSD), 1);
    m.insertThresholdMax = std::min((int)(mean + 3 * m.rightSD), MI);
    if (a.partialFlag) { m.insertThresholdMin -= a.partialReadLen; m.insertThresholdMax += a.partialReadLen; }
    return true;
}

}  // namespace fb
