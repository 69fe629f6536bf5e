// fb_preprocess.cpp -- fb_preprocess_main(): drop-in for the reference `Preprocess` executable, the step that
// turns the bowtie2 SAM into the inputs of FillGaps: gapInfo.txt, stat.txt, stat2.txt, myout.sam and the per-gap read files
// Gaps/gaps_<g>.sam (mode 2) or Gaps/partial_gaps_<g>.sam (mode 1).  Same 13 positional arguments (Preprocess.cpp:1861-1870),
// same files, byte-identical content -- except the two fields of gaps_<g>.sam lines that the reference prints from
// uninitialised heap memory (the MD string of a read without an MD tag and the IH count of reads that never went through
// printVectors, Preprocess.cpp:1493,404-410); FillGaps reads neither (parseUnmapped, Figbird.cpp:5661-5767).  Here they are
// `*` and 0.
//
// What is different by design (the reference is a single sequential pass with per-read linear work over all gaps):
//   * gap lookup: the reference's checkPos2 / checkPos scan every gap for every read (Preprocess.cpp:536-639; with 150-base
//     reads no pair is a "101M" full map, so every pair pays it: O(reads x gaps)).  Here the gaps of a scaffold are sorted
//     interval arrays and a lookup is two binary searches that return the same gap the linear scan returns first.
//   * the SAM is mapped and cut into blocks at read-name changes; blocks are parsed in parallel (pair grouping, myout.sam
//     text, candidate records for the per-gap files); what depends on a gap's history (the 3001-read cap, the duplicate
//     filter, the MIM evidence of stat2.txt) is then replayed per gap, gaps in parallel, candidates in SAM order.
//   * a per-gap file is written once instead of being re-opened for every read (Preprocess.cpp:406,427).
//   * duplicate filter of mode 2 (exact match or containment of the read minus 2 bases at both ends, Preprocess.cpp:362-388):
//     a hash set over the windows of the stored reads; a hit is confirmed with the reference's own comparison.
// Inputs the block scheme cannot reproduce exactly (a read with several alignment lines, a read name that comes back after
// other reads, physical lines of 1023 bytes or more) are detected and the file is then processed as one block, sequentially.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "fb_io.h"
#include "fb_tools.h"

namespace fb {
namespace {

using sv = std::string_view;

struct PGap { int contigNo; long start; int len; };

// ---- one alignment line (getSAM, Preprocess.cpp:1491-1551): strtok on "\t\n " -- blanks split fields and empty fields vanish
struct Sam {
    sv qname, rname, cigar, seq, qual, md;
    int flag = 0, pos = 0, tlen = 0, nm = -1;
    long contigNo = -1;
    long ih = 0;
    bool hasMd = false;
    std::string seqOwn, qualOwn;      // set when the program rewrites the field in place
    sv seqv() const { return seqOwn.empty() ? seq : sv(seqOwn); }
    sv qualv() const { return qualOwn.empty() ? qual : sv(qualOwn); }
};

inline int atoiSv(sv s) {
    size_t i = 0; while (i < s.size() && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    bool neg = false; if (i < s.size() && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    long v = 0; for (; i < s.size() && s[i] >= '0' && s[i] <= '9'; i++) v = v * 10 + (s[i] - '0');
    return (int)(neg ? -v : v);
}
inline bool isDelim(char c) { return c == '\t' || c == '\n' || c == ' '; }
inline sv nextTok(const char*& p, const char* e) {
    while (p < e && isDelim(*p)) p++;
    const char* a = p;
    while (p < e && !isDelim(*p)) p++;
    return sv(a, (size_t)(p - a));
}

struct ContigIndex {
    std::unordered_map<sv, long> byName;
    long lookup(sv name, sv& lastName, long& lastNo) const {
        if (name == lastName) return lastNo;
        auto it = byName.find(name);
        lastName = name; lastNo = it == byName.end() ? -1 : it->second;
        return lastNo;
    }
};

void revcomp(sv s, std::string& out) {      // reverse(), Preprocess.cpp:145-165: upper-case ACGT only, everything else -> N
    out.resize(s.size());
    for (size_t i = 0; i < s.size(); i++) {
        char o; switch (s[i]) { case 'A': o = 'T'; break; case 'C': o = 'G'; break; case 'G': o = 'C'; break; case 'T': o = 'A'; break; default: o = 'N'; }
        out[s.size() - 1 - i] = o;
    }
}

int parseDel(sv cigar) {      // Preprocess.cpp:167-201: the soft clip in front of the first M
    const size_t s = cigar.find('S'), m = cigar.find('M');
    if (s != sv::npos && m != sv::npos && s < m) return atoiSv(cigar.substr(0, s));
    return 0;
}

// parse_Cigar, Preprocess.cpp:203-294: v[0] = leading soft clip, v[1] = first M, v[2] = trailing soft clip (as far as the
// reference finds it: it re-uses a scratch buffer by the original index, reproduced below)
void parseCigar(sv cigar, int readlen, int v[3]) {
    const size_t i1 = cigar.find('S'), i2 = cigar.find('M');
    if (i1 == sv::npos || i2 == sv::npos || !(i1 < i2)) return;
    v[0] = atoiSv(cigar.substr(0, i1));
    v[1] = atoiSv(cigar.substr(i1 + 1, i2 - i1 - 1));
    if (v[0] + v[1] == readlen) return;
    const std::string s(cigar.substr(i2 + 1));
    int lastS = -1;
    for (size_t i = 0; i < s.size(); i++) if (s[i] == 'S') lastS = (int)i;
    if (lastS < 0) return;
    auto hasOp = [](sv t) { return t.find_first_of("DIMX=") != sv::npos; };
    if (!hasOp(s)) { v[2] = readlen - v[0] - v[1]; return; }
    std::string buf = s; buf.push_back('\0');                 // new_s[n + 1]
    std::string two;
    for (int k = lastS - 2; k < lastS; k++) if (k >= 0) two += s[(size_t)k];
    for (size_t k = 0; k < two.size() && k < buf.size(); k++) buf[k] = two[k];      // strcpy(new_s, two chars)
    if (two.size() < buf.size()) buf[two.size()] = '\0';
    std::string num = two;
    if (hasOp(two)) { num.clear(); const char c = (lastS - 1 >= 0 && (size_t)(lastS - 1) < buf.size()) ? buf[(size_t)(lastS - 1)] : '\0'; if (c) num += c; }
    v[2] = atoiSv(num);
}

// checkMIM, Preprocess.cpp:885-925: "aMbIcM" without S, D, =, X -> (1, b + 1)
bool checkMIM(sv cigar, int& gaplen) {
    int i1 = 0, i2 = 0, i3 = 0, mc = 0, ic = 0;
    for (int i = 0; i < (int)cigar.size(); i++) {
        const char c = cigar[(size_t)i];
        if (c == 'S' || c == 'D' || c == '=' || c == 'X') return false;
        if (c == 'M') { if (mc == 0) i1 = i; else if (mc == 1) i3 = i; else return false; mc++; }
        else if (c == 'I') { if (ic == 1) return false; i2 = i; ic++; }
    }
    if (i1 && i2 && i3 && i1 < i2 && i2 < i3) { gaplen = atoiSv(cigar.substr((size_t)i1 + 1, (size_t)(i2 - i1 - 1))) + 1; return true; }
    return false;
}

bool atMostThreeN(sv s) { int n = 0; for (char c : s) if (c == 'N') n++; return n <= 3; }                 // check_Ncount_partial :857
bool badChar(sv s) { for (char c : s) if (!strchr("ACGTNacgtn", c) || c == '\0') return true; return false; }      // checkChar :868
bool mostlyN(sv s) { int n = 0; for (char c : s) if (c == 'N' || c == 'n') n++; return !((double)n / (double)s.size() < 0.8); }

inline void appendInt(std::string& o, long v) { char b[24]; const int n = snprintf(b, sizeof b, "%ld", v); o.append(b, (size_t)n); }

// writeSam / writeSam2, Preprocess.cpp:404-416
void appendSamLine(std::string& o, const Sam& r, bool ihKnown) {
    o.append(r.qname); o += '\t'; appendInt(o, r.flag); o += '\t'; appendInt(o, r.contigNo); o += '\t'; appendInt(o, r.pos); o += '\t';
    o.append(r.cigar); o += '\t'; appendInt(o, r.tlen); o += '\t'; o.append(r.seqv()); o += '\t'; o.append(r.qualv()); o += '\t';
    if (r.hasMd) o.append(r.md); else o += '*';
    o += "\tIH:i:"; appendInt(o, ihKnown ? r.ih : 0); o += '\n';
}

// ---- gap lookup ------------------------------------------------------------------------------------------------
struct GapIndex {
    const std::vector<PGap>* gaps = nullptr;
    struct PerContig { std::vector<int> idx; std::vector<long> start, end; bool sorted = true; };
    std::vector<PerContig> pc;
    int maxDistance = 0, readMean = 0;
    void build(const std::vector<PGap>& g, size_t nContigs) {
        gaps = &g; pc.assign(nContigs, PerContig());
        for (int i = 0; i < (int)g.size(); i++) {
            if (g[i].contigNo < 0 || (size_t)g[i].contigNo >= nContigs) continue;
            PerContig& c = pc[(size_t)g[i].contigNo];
            // the reference's scans use int arithmetic for the gap end (Preprocess.cpp:620)
            const long s = g[i].start, e = (long)(int)(g[i].start + g[i].len);
            if (!c.idx.empty() && (s <= c.start.back() || e <= c.end.back())) c.sorted = false;
            c.idx.push_back(i); c.start.push_back(s); c.end.push_back(e);
        }
    }
    // checkPos2, Preprocess.cpp:616-639: first gap (file order) whose start lies in [pos, pos + readlength - 2] or, for a read
    // with a leading soft clip `del`, whose end lies in [pos - del, pos - 1]
    int partialGap(long contigNo, long pos, int readlength, int del) const {
        if (contigNo < 0 || (size_t)contigNo >= pc.size()) return -1;
        const PerContig& c = pc[(size_t)contigNo];
        if (!c.sorted) {
            for (size_t k = 0; k < c.idx.size(); k++) {
                const long gs = c.start[k]; const int ge = (int)c.end[k];
                if ((pos > gs - readlength + 1 && pos <= gs) || (pos > ge && del && (pos - del) <= ge)) return c.idx[k];
            }
            return -1;
        }
        size_t best = c.idx.size();
        { const size_t a = (size_t)(std::lower_bound(c.start.begin(), c.start.end(), pos) - c.start.begin()); if (a < c.idx.size() && pos > c.start[a] - readlength + 1) best = a; }
        if (del) { const size_t b = (size_t)(std::lower_bound(c.end.begin(), c.end.end(), pos - del) - c.end.begin()); if (b < best && c.end[b] < pos) best = b; }
        return best < c.idx.size() ? c.idx[best] : -1;
    }
    static int closerToMean(int a, int b, double mean) { return std::fabs(mean - a) < std::fabs(mean - b) ? a : b; }      // checkInsert :515
    static bool inRange(int a, int b, int mean) {                                                                          // checkRange :525
        const int lo = mean - 1000, hi = mean + 1000;
        return (a > lo && a < hi) || (b > lo && b < hi) || (a < lo && b > hi) || (b < lo && a > hi);
    }
    // checkPos, Preprocess.cpp:536-614: the gap a one-end-unmapped pair can reach (mate forward: gap start within maxDistance to
    // the right; mate reverse: gap end within maxDistance to the left)
    int unmappedGap(long contigNo, long pos, int strandNo, int readlength) const {
        if (contigNo < 0 || (size_t)contigNo >= pc.size()) return -1;
        const PerContig& c = pc[(size_t)contigNo];
        const std::vector<PGap>& g = *gaps;
        auto matches = [&](size_t k) {
            const long gs = c.start[k], ge = g[(size_t)c.idx[k]].start + g[(size_t)c.idx[k]].len;
            return (strandNo == 0 && pos > gs - maxDistance && pos < gs) || (strandNo == 1 && pos > ge && pos < ge + maxDistance);
        };
        size_t k0 = 0, k1 = c.idx.size();
        if (c.sorted) {      // candidates are a contiguous stretch
            if (strandNo == 0) { k0 = (size_t)(std::upper_bound(c.start.begin(), c.start.end(), pos) - c.start.begin()); k1 = (size_t)(std::lower_bound(c.start.begin(), c.start.end(), pos + maxDistance) - c.start.begin()); }
            else if (strandNo == 1) { k0 = (size_t)(std::upper_bound(c.end.begin(), c.end.end(), pos - maxDistance) - c.end.begin()); k1 = (size_t)(std::lower_bound(c.end.begin(), c.end.end(), pos) - c.end.begin()); }
            else return -1;
        }
        int hits = 0, lastHit = -1, minVal = 1000000, minIdx = -1;
        std::vector<std::pair<int, int>> est;      // (gap, insert estimate) of the matching gaps
        for (size_t k = k0; k < k1; k++) {
            if (!matches(k)) continue;
            const int i = c.idx[k];
            if (maxDistance <= 250) return i;
            const long gs = g[(size_t)i].start; const int gl = g[(size_t)i].len;
            int v0, v1, t = 0;
            if (pos < gs) { v0 = (int)(gs + gl - pos + readlength); v1 = (int)(gs - pos + 1); }
            else { v0 = (int)(pos - gs + 2 * readlength - 1); v1 = (int)(pos - gs - gl + readlength + 1); }
            if (inRange(v0, v1, readMean)) t = closerToMean(v0, v1, (double)readMean);
            if (t != 0) { hits++; lastHit = i; }
            const int d = std::abs(readMean - t);
            if (d < minVal) { minVal = d; minIdx = i; }
            est.emplace_back(i, t);
        }
        if (maxDistance <= 250 || hits == 0) return -1;
        const int minThresh = (int)(readMean - readMean * 0.6);
        const int chosen = hits == 1 ? lastHit : minIdx;
        for (auto& e : est) if (e.first == chosen) return e.second < minThresh ? -1 : chosen;
        return -1;
    }
};

// ---- a candidate (a line, or a pair of lines) for a per-gap file: accepted or not by the gap's history
struct Cand {
    int gap;
    std::string key;      // the string the duplicate filter compares and stores
    std::string text;     // what is appended to the gap's file when the candidate is accepted (may be empty)
    bool mim = false; int mimLen = 0;
};

// polynomial window hashes of `s` for window length w, added to `set`
const uint64_t kHashBase = 0x9E3779B97F4A7C15ULL | 1ULL;
void addWindows(sv s, size_t w, std::unordered_set<uint64_t>& set) {
    if (w == 0 || s.size() < w) return;
    uint64_t pw = 1; for (size_t i = 1; i < w; i++) pw *= kHashBase;
    uint64_t h = 0; for (size_t i = 0; i < w; i++) h = h * kHashBase + (unsigned char)s[i];
    set.insert(h);
    for (size_t i = w; i < s.size(); i++) { h = (h - pw * (unsigned char)s[i - w]) * kHashBase + (unsigned char)s[i]; set.insert(h); }
}
uint64_t windowHash(sv s) { uint64_t h = 0; for (char c : s) h = h * kHashBase + (unsigned char)c; return h; }

// The history of one gap: the cap of 3001 reads (`read_count <= 3000` before every insertion, Preprocess.cpp:1229,1350,1678), the
// duplicate filter (check_duplicate :362-402) and the MIM evidence of stat2.txt (checkMIM :885).
struct GapState {
    int count = 0, perfect = 0, perfectLen = 0;
    std::string text;
    std::vector<std::string> stored;
    std::unordered_set<std::string> seen;                              // mode 1
    std::map<size_t, std::unordered_set<uint64_t>> windows;            // mode 2: window length -> hashes of all such windows of the stored reads
    bool offer(int samflag, const Cand& c) {
        if (count > 3000) return false;
        const sv key(c.key);
        if (samflag != 2) {      // exact match with a stored read
            if (!seen.insert(c.key).second) return false;
            text += c.text; count++;
            if (c.mim) { perfect = 1; perfectLen = c.mimLen; }
            return true;
        }
        // mode 2: exact match, or a stored read contains this one minus two bases at both ends
        bool dup = false;
        if (key.size() < 4) {      // (s2.substr(2, size - 4) with size < 4: unsigned arithmetic, the rest of the string)
            for (const std::string& s : stored) { if (sv(s) == key) { dup = true; break; } if (key.size() >= 2 && s.find(key.substr(2)) != std::string::npos) { dup = true; break; } }
        } else {
            const sv core = key.substr(2, key.size() - 4);
            auto it = windows.find(core.size());
            if (it == windows.end()) { it = windows.emplace(core.size(), std::unordered_set<uint64_t>()).first; for (const std::string& s : stored) addWindows(s, core.size(), it->second); }
            if (core.empty()) dup = !stored.empty();
            else if (it->second.count(windowHash(core))) for (const std::string& s : stored) if (sv(s).find(core) != sv::npos) { dup = true; break; }
        }
        if (dup) return false;
        text += c.text; count++;
        stored.push_back(c.key);
        for (auto& w : windows) addWindows(key, w.first, w.second);
        return true;
    }
};

struct BlockOut {
    std::string myout, red1, red2;
    std::vector<Cand> cands;
    long flushes = 0, unCount = 0, mixedCount = 0; unsigned long maxReadLength = 0;
    std::vector<long> insertHist;      // mode 2 with maxDistance > 250, first pass: tlen histogram of unique, deletion-free pairs
    long discarded = 0;
    std::string firstQ1, firstQ2, lastQ1, lastQ2; bool anyProper = false;
    bool irregular = false;            // something the block scheme must not be trusted with
};

struct Config {
    int samflag = 0, maxDistance = 0;
    bool writeReduced = false;
    bool modelPass = false;            // mode 2, maxDistance > 250, first pass: only myout.sam + the insert histogram
    bool noVectors = false;            // mode 2, maxDistance > 250, second pass: proper pairs are skipped
    const ContigIndex* contigs = nullptr;
    const std::vector<unsigned long>* contigLengths = nullptr;
    const GapIndex* gidx = nullptr;
    std::vector<GapState>* inlineStates = nullptr;      // sequential mode: candidates meet their gap's history at once
};

// The pair-grouping state machine of main(), Preprocess.cpp:2447-2596 (and :2315-2383 for the model pass), over the records
// [begin, end) of the mapped SAM.
class BlockParser {
public:
    BlockParser(const Config& c, BlockOut& o) : cfg(c), out(o), seq_(c.inlineStates != nullptr) {}
    ~BlockParser() { for (Sam* r : all_) delete r; }
    void run(const char* begin, const char* end) {
        p_ = begin; e_ = end;
        const char *rb, *re;
        while (nextRecord(rb, re)) {
            if (*rb == '@') continue;
            Sam* r1 = parse(rb, re);
            bool ended = false;
            while ((r1->flag & 2) == 0) {      // not aligned as a proper pair: the lines of mate 1, then the lines of mate 2
                std::vector<Sam*> m1, m2;
                std::string q(r1->qname); int seg = r1->flag & 192;
                bool eof = false;
                while (r1->qname == sv(q) && (r1->flag & 192) == seg) { m1.push_back(r1); if (nextRecord(rb, re)) r1 = parse(rb, re); else { eof = true; break; } }
                if (eof) { ended = true; break; }      // (the reference never leaves this loop at the end of the file)
                q = std::string(r1->qname); seg = r1->flag & 192;
                while (r1->qname == sv(q) && (r1->flag & 192) == seg) { m2.push_back(r1); if (nextRecord(rb, re)) r1 = parse(rb, re); else { ended = true; break; } }
                if ((m1.size() != 1 || m2.size() != 1) && !seq_) out.irregular = true;
                if (!cfg.modelPass) {
                    if (cfg.writeReduced) rewriteReadset(*m1[0], *m2[0]);
                    mixed(m1, m2);
                }
                release(m1); release(m2);
                if (ended) break;
            }
            if (ended) break;
            if (!nextRecord(rb, re)) break;
            Sam* r2 = parse(rb, re);
            if (r1->qname != sv(preq1_) || r2->qname != sv(preq2_)) {
                if (!out.anyProper) { out.anyProper = true; out.firstQ1 = std::string(r1->qname); out.firstQ2 = std::string(r2->qname); }
                preq1_ = std::string(r1->qname); preq2_ = std::string(r2->qname);
                if (!cfg.noVectors) flushGroup();
                if (cfg.samflag == 1 && !cfg.modelPass) {
                    if (!(r1->cigar == "101M" && r2->cigar == "101M")) {      // everything but the literal "101M" pair is looked up (Preprocess.cpp:2546-2564)
                        collectPartial(*r1, r2->pos);
                        collectPartial(*r2, r1->pos);
                        if (cfg.writeReduced) rewriteReadset(*r1, *r2);
                    }
                }
            } else if (!seq_) out.irregular = true;      // further alignment lines of the same pair
            if (!cfg.noVectors) { g1_.push_back(r1); g2_.push_back(r2); } else { release(r1); release(r2); }
        }
        if (!cfg.noVectors) flushGroup();
        out.lastQ1 = preq1_; out.lastQ2 = preq2_;
    }

private:
    const Config& cfg; BlockOut& out; const bool seq_;
    const char *p_ = nullptr, *e_ = nullptr;
    std::vector<Sam*> free_, all_;      // records are recycled; all_ owns them
    std::vector<Sam*> g1_, g2_;
    std::string preq1_ = "*", preq2_ = "*";
    sv lastName_; long lastNo_ = -1;
    std::string tmp_, tmp2_;

    // fgets(line, 1024, f): a physical line of 1023 bytes or more arrives in pieces
    bool nextRecord(const char*& rb, const char*& re) {
        if (p_ >= e_) return false;
        const char* nl = (const char*)memchr(p_, '\n', (size_t)(e_ - p_));
        const char* le = nl ? nl + 1 : e_;
        if (le - p_ > 1023) { le = p_ + 1023; if (!seq_) out.irregular = true; }
        rb = p_; re = le; p_ = le;
        return true;
    }
    void release(Sam* r) { free_.push_back(r); }
    void release(std::vector<Sam*>& v) { for (Sam* r : v) free_.push_back(r); v.clear(); }
    Sam* parse(const char* b, const char* e) {
        Sam* rec;
        if (!free_.empty()) { rec = free_.back(); free_.pop_back(); *rec = Sam(); } else { rec = new Sam(); all_.push_back(rec); }
        Sam& s = *rec;
        const char* p = b;
        s.qname = nextTok(p, e); s.flag = atoiSv(nextTok(p, e)); s.rname = nextTok(p, e); s.pos = atoiSv(nextTok(p, e));
        nextTok(p, e);      // mapq
        s.cigar = nextTok(p, e);
        nextTok(p, e);      // rnext
        nextTok(p, e);      // pnext
        s.tlen = atoiSv(nextTok(p, e)); s.seq = nextTok(p, e); s.qual = nextTok(p, e);
        for (;;) {
            const sv t = nextTok(p, e);
            if (t.empty()) break;
            if (t.size() >= 2 && t[0] == 'M' && t[1] == 'D') { s.md = t; s.hasMd = true; }
            if (t.size() >= 2 && t[0] == 'N' && t[1] == 'M') s.nm = t.size() > 5 ? atoiSv(t.substr(5)) : 0;
        }
        s.contigNo = cfg.contigs->lookup(s.rname, lastName_, lastNo_);
        return rec;
    }
    void noteLength(sv s) { if (s.size() > out.maxReadLength) out.maxReadLength = s.size(); }
    // a candidate meets its gap's history now (sequential mode) or after all blocks are parsed; the answer is only known now
    bool offer(Cand&& c) {
        if (cfg.inlineStates) return c.gap >= 0 && (size_t)c.gap < cfg.inlineStates->size() && (*cfg.inlineStates)[(size_t)c.gap].offer(cfg.samflag, c);
        out.cands.push_back(std::move(c));
        return false;
    }

    // printVectors, Preprocess.cpp:641-855: the alignments of one properly paired read pair go to myout.sam with IH = their number
    void flushGroup() {
        out.flushes++;      // every call ends in totalCount++ (:843), or in unCount++ and totalCount++ on the early return below
        const size_t ih = g1_.size();
        for (size_t i = 0; i < ih; i++) {
            Sam& a = *g1_[i]; Sam& b = *g2_[i];
            if (a.rname == "*" || b.rname == "*") {
                noteLength(a.seqv()); noteLength(b.seqv());
                if (!mostlyN(a.seqv()) && !mostlyN(b.seqv())) { out.unCount++; break; }
            } else if (a.rname != b.rname) {
            } else {
                noteLength(a.seqv()); a.ih = (long)ih;
                noteLength(b.seqv()); b.ih = (long)ih;
                appendSamLine(out.myout, a, true); appendSamLine(out.myout, b, true);
                if (cfg.modelPass) { modelLine(a); modelLine(b); }
            }
        }
        release(g1_); release(g2_);
    }
    // processMapping + updateInsertCounts on a myout.sam line, Preprocess.cpp:1744-1828 (MAX_FRAGMENT_SIZE = maxInsertSize = 5000)
    void modelLine(const Sam& r) {
        if (r.ih != 1) return;
        if (r.md.size() > 5 && r.md[5] == '^') return;
        if (r.contigNo < 0 || (size_t)r.contigNo >= cfg.contigLengths->size() || !((*cfg.contigLengths)[(size_t)r.contigNo] > 0)) return;
        const int t = r.tlen;
        if (t <= 0) return;
        if (t < 5000) { if (out.insertHist.empty()) out.insertHist.assign(5000, 0); out.insertHist[(size_t)t]++; }
        else if (t > 5000) out.discarded++;
    }

    // reWriteReadset, Preprocess.cpp:1696-1731: FASTQ records in sequencing orientation; QUAL of a reverse-strand read is
    // reversed IN the record, so every later use of the record sees the reversed string
    void rewriteReadset(Sam& a, Sam& b) {
        auto one = [&](Sam& r, std::string& o) {
            o += '@'; o.append(r.qname); o += '\n';
            if ((r.flag & 16) >> 4) {
                revcomp(r.seqv(), tmp_);
                std::string q(r.qualv()); std::reverse(q.begin(), q.end()); r.qualOwn = q;
                o += tmp_; o += "\n+\n"; o.append(r.qualv()); o += '\n';
            } else { o.append(r.seqv()); o += "\n+\n"; o.append(r.qualv()); o += '\n'; }
        };
        one(a, out.red1); one(b, out.red2);
    }

    // collectPartialSAM + writePartialSam, Preprocess.cpp:1667-1694, 425-502 (mode 1)
    void collectPartial(const Sam& r, int pos2) {
        const int strand = (r.flag & 16) >> 4;
        const int del = parseDel(r.cigar);
        const sv seq = r.seqv();
        const int g = cfg.gidx->partialGap(r.contigNo, r.pos, (int)seq.size(), del);
        if (g < 0) return;
        if (!atMostThreeN(seq)) return;
        Cand c; c.gap = g; c.key = std::string(seq);
        const PGap& G = (*cfg.gidx->gaps)[(size_t)g];
        const int gs = (int)G.start, ge = (int)(G.start + G.len);
        const int readlength = (int)seq.size();
        int clipped = 0, match = -1; bool write = false;
        if (r.pos < gs) {
            match = strand == 0 ? 1 : 4;
            int v[3] = {0, 0, 0};
            parseCigar(r.cigar, readlength, v);
            if (v[0]) { if (v[2]) { clipped = readlength - v[2] - 1; write = true; } }      // S..M only: the read is dropped but still counted
            else { clipped = gs - r.pos; write = true; }
        } else if (r.pos > gs) {
            match = strand == 0 ? 2 : 3;
            clipped = ge - 1 - r.pos + del + 2;
            write = true;
        }
        if (write) {
            std::string& t = c.text;
            t.append(seq); t += '\t'; appendInt(t, clipped); t += '\t'; appendInt(t, match); t += '\t'; appendInt(t, r.pos); t += '\t';
            t.append(r.cigar); t += '\t'; appendInt(t, pos2); t += '\t'; t.append(r.qualv()); t += '\n';
        }
        c.mim = checkMIM(r.cigar, c.mimLen);
        offer(std::move(c));
    }

    // printMixedVectors, Preprocess.cpp:999-1489
    void mixed(std::vector<Sam*>& m1, std::vector<Sam*>& m2) {
        for (size_t i = 0; i < m1.size(); i++) {
            for (size_t j = 0; j < m2.size(); j++) {
                Sam& r1 = *m1[i]; Sam& r2 = *m2[j];
                if (i == 0 && j == 0) {
                    noteLength(r1.seqv()); noteLength(r2.seqv());
                    if (!mostlyN(r1.seqv()) && !mostlyN(r2.seqv())) { out.unCount++; out.mixedCount++; }
                    else return;
                }
                const bool u1 = (r1.flag & 4) != 0, u2 = (r2.flag & 4) != 0;
                if (u1 && u2) return;
                if ((!u1 && u2) || (!u1 && !u2 && cfg.maxDistance > 250)) {
                    for (size_t k = 0; k < m1.size(); k++) {
                        Sam& a = *m1[k]; Sam& b = *m2[0];
                        const int strand1 = (a.flag & 16) >> 4;
                        if (cfg.samflag == 2 && !badChar(b.seqv())) {
                            if (u2) unmappedMate(a, b, strand1);
                            else linkedPair(a, b);
                        }
                        if (cfg.samflag == 1) collectPartial(a, -1);
                    }
                    return;
                }
                // (mate 1 unmapped with mate 2 mapped, or both mapped on different scaffolds: nothing is written)
            }
        }
    }
    // mode 2: `a` is aligned, `b` is its unaligned mate.  The pair goes to the gap `a` points at; the filter sees `b` in the
    // orientation FillGaps will score it in (reverse-complemented when `a` is on the forward strand).  Preprocess.cpp:1225-1250
    void unmappedMate(Sam& a, Sam& b, int strand1) {
        const int g = cfg.gidx->unmappedGap(a.contigNo, a.pos, strand1, (int)b.seqv().size());
        if (g < 0) return;
        Cand c; c.gap = g;
        if (strand1 == 1) c.key = std::string(b.seqv()); else revcomp(b.seqv(), c.key);
        appendSamLine(c.text, a, false); appendSamLine(c.text, b, false);
        std::string kept = strand1 == 0 ? c.key : std::string();
        // an accepted pair leaves `b` reverse-complemented in the record (:1241), which a further alignment line of `a` then sees
        if (offer(std::move(c)) && strand1 == 0) b.seqOwn = kept;
    }
    // mode 2, jump library (maxDistance > 250): both mates aligned but not as a proper pair -- each may vouch for the other
    // (Preprocess.cpp:1251-1343).  The reference edits SEQ in place before writing and restores it afterwards.
    void linkedPair(Sam& a, Sam& b) {
        const int strand1 = (a.flag & 16) >> 4, strand2 = (b.flag & 16) >> 4;
        {
            const int g = cfg.gidx->unmappedGap(a.contigNo, a.pos, strand1, (int)b.seqv().size());
            if (g >= 0) {
                Sam bb = b;
                if (strand2 == 1) { revcomp(b.seqv(), tmp_); bb.seqOwn = tmp_; }
                revcomp(bb.seqv(), tmp2_);
                Cand c; c.gap = g; c.key = strand1 == 1 ? std::string(bb.seqv()) : tmp2_;
                appendSamLine(c.text, a, false); appendSamLine(c.text, bb, false);
                offer(std::move(c));
            }
        }
        {
            const int g = cfg.gidx->unmappedGap(b.contigNo, b.pos, strand2, (int)a.seqv().size());
            if (g >= 0) {
                Sam aa = a;
                if (strand1 == 1) { revcomp(a.seqv(), tmp_); aa.seqOwn = tmp_; }
                revcomp(aa.seqv(), tmp2_);
                Cand c; c.gap = g; c.key = strand2 == 1 ? std::string(aa.seqv()) : tmp2_;
                appendSamLine(c.text, b, false); appendSamLine(c.text, aa, false);
                offer(std::move(c));
            }
        }
    }
};

bool writeWhole(const std::string& path, const std::string& data) {
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    size_t off = 0;
    while (off < data.size()) { const ssize_t k = write(fd, data.data() + off, data.size() - off); if (k <= 0) { close(fd); return false; } off += (size_t)k; }
    close(fd);
    return true;
}

// (an exception on a worker thread -- out of memory, say -- is carried to the caller, which turns it into exit status 1)
template <class F>
void parallelFor(size_t n, int threads, F f) {
    threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n));
    if (threads <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    std::mutex emu; std::exception_ptr err;
    for (int t = 0; t < threads; t++) th.emplace_back([&] {
        try { for (size_t i; (i = next++) < n;) f(i); }
        catch (...) { std::lock_guard<std::mutex> l(emu); if (!err) err = std::current_exception(); next = n; }
    });
    for (auto& t : th) t.join();
    if (err) std::rethrow_exception(err);
}

struct Mapped {
    const char* p = nullptr; size_t n = 0; void* base = nullptr; std::string owned;
    bool open(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return false; }
        n = (size_t)st.st_size;
        if (n) {
            base = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (base == MAP_FAILED) { base = nullptr; close(fd); if (!slurp(path, owned)) return false; p = owned.data(); n = owned.size(); return true; }
            p = (const char*)base;
        }
        close(fd);
        return true;
    }
    ~Mapped() { if (base) munmap(base, n); }
};

// Block starts: the start of a line that opens a new read pair -- its name differs from the previous line's, it is a first
// mate (flag & 64) and the previous line is a second mate (flag & 128).  The state machine is between pairs there.
std::vector<size_t> blockCuts(const Mapped& sam, size_t target) {
    std::vector<size_t> cuts{0};
    auto lineStartAfter = [&](size_t o) -> size_t { const char* nl = (const char*)memchr(sam.p + o, '\n', sam.n - o); return nl ? (size_t)(nl - sam.p) + 1 : sam.n; };
    auto nameFlag = [&](size_t o, sv& name, int& flag) { const char* p = sam.p + o; const char* e = sam.p + sam.n; name = nextTok(p, e); flag = atoiSv(nextTok(p, e)); };
    for (size_t want = target; want < sam.n; want += target) {
        size_t prev = lineStartAfter(want);
        if (prev >= sam.n) break;
        size_t o = lineStartAfter(prev);
        bool found = false;
        for (int tries = 0; o < sam.n && tries < 4096; tries++) {
            sv n0, n1; int f0, f1;
            nameFlag(prev, n0, f0); nameFlag(o, n1, f1);
            if (sam.p[prev] != '@' && sam.p[o] != '@' && n0 != n1 && (f0 & 128) && (f1 & 64)) { found = true; break; }
            prev = o; o = lineStartAfter(o);
        }
        if (!found) break;      // (no pair boundary near here: the rest is one block)
        if (o > cuts.back()) cuts.push_back(o);
        if (o > want) want = o - (o % target);
    }
    cuts.push_back(sam.n);
    return cuts;
}

// One pass over the SAM: blocks in parallel, text in block order.  Returns false when the input needs the sequential path.
struct PassResult {
    long flushes = 0, unCount = 0, mixedCount = 0; unsigned long maxReadLength = 0;
    std::vector<long> insertHist; long discarded = 0;
    std::vector<std::vector<Cand>> cands;
};
bool runPass(const Config& cfg, const Mapped& sam, int threads, FILE* myout, FILE* red1, FILE* red2, PassResult& res) {
    const bool sequential = cfg.inlineStates != nullptr;
    size_t target = std::max<size_t>(1u << 20, std::min<size_t>(16u << 20, sam.n / (size_t)(std::max(1, threads) * 8) + 1));
    if (const char* e = getenv("FIGBIRD_PP_BLOCK")) target = (size_t)std::max(1L, atol(e));      // (tests: many small blocks)
    const std::vector<size_t> cuts = sequential ? std::vector<size_t>{0, sam.n} : blockCuts(sam, target);
    const size_t nb = cuts.size() - 1;
    std::vector<BlockOut> outs(nb);
    std::mutex mu; size_t nextWrite = 0; std::vector<char> done(nb, 0);
    parallelFor(nb, threads, [&](size_t b) {
        { BlockParser bp(cfg, outs[b]); bp.run(sam.p + cuts[b], sam.p + cuts[b + 1]); }
        std::lock_guard<std::mutex> l(mu);
        done[b] = 1;
        while (nextWrite < nb && done[nextWrite]) {      // ordered flush: text leaves memory as soon as its turn comes
            BlockOut& o = outs[nextWrite];
            if (myout && !o.myout.empty()) fwrite(o.myout.data(), 1, o.myout.size(), myout);
            if (red1 && !o.red1.empty()) fwrite(o.red1.data(), 1, o.red1.size(), red1);
            if (red2 && !o.red2.empty()) fwrite(o.red2.data(), 1, o.red2.size(), red2);
            std::string().swap(o.myout); std::string().swap(o.red1); std::string().swap(o.red2);
            nextWrite++;
        }
    });
    if (nb > 1) {
        std::string l1 = "*", l2 = "*";
        for (size_t b = 0; b < nb; b++) {
            if (outs[b].irregular) return false;
            // a pair name that comes back after a block boundary would have joined the earlier group in the sequential program
            if (outs[b].anyProper) { if (b > 0 && outs[b].firstQ1 == l1 && outs[b].firstQ2 == l2) return false; l1 = outs[b].lastQ1; l2 = outs[b].lastQ2; }
        }
    }
    for (size_t b = 0; b < nb; b++) {
        res.flushes += outs[b].flushes; res.unCount += outs[b].unCount; res.mixedCount += outs[b].mixedCount;
        res.maxReadLength = std::max(res.maxReadLength, outs[b].maxReadLength);
        res.discarded += outs[b].discarded;
        if (!outs[b].insertHist.empty()) { if (res.insertHist.empty()) res.insertHist.assign(5000, 0); for (size_t i = 0; i < 5000; i++) res.insertHist[i] += outs[b].insertHist[i]; }
        res.cands.push_back(std::move(outs[b].cands));
    }
    // printVectors runs once per new pair name plus once at the end (Preprocess.cpp:2541,2596).  A block flushes its last group
    // at its own end and starts with a flush of nothing: nb blocks make nb - 1 calls more than the one sequential pass.
    if (!cfg.noVectors) res.flushes -= (long)nb - 1;
    return true;
}

}  // namespace

int preprocessMain(int argc, const char* const* argv) {
    if (argc < 14) { fprintf(stderr, "Invalid parameters\n"); return 1; }
    const std::string contigFile = argv[1];
    const int maxDistance = atoi(argv[2]), samflag = atoi(argv[3]);
    const std::string mapFile = argv[4], outName = argv[5], filledName = argv[6];
    const std::string gapsDir = argv[9], tmpDir = argv[10];
    const int def = atoi(argv[11]), genomeReduction = atoi(argv[12]), readReduction = atoi(argv[13]);
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("FIGBIRD_HOST_THREADS")) threads = atoi(e);
    threads = std::max(1, threads);

    // ---- genome_reduction: gap number -> scaffold number of the un-reduced genome (Preprocess.cpp:1888-2007)
    std::map<int, int> contigOfGap;
    if (genomeReduction == 1) {
        RawFasta full;
        if (!loadFastaRaw(filledName, full)) { fprintf(stderr, "Can't open contig file\n"); return 1; }
        int gapcount = 0; bool in = false;
        for (size_t i = 0; i < full.seq.size(); i++) {
            const std::string& c = full.seq[i];
            for (size_t j = 0; j < c.size(); j++) {
                const bool isN = c[j] == 'N' || c[j] == 'n';
                if (isN) in = true;
                if ((!isN && in) || (isN && j == c.size() - 1)) { contigOfGap.emplace(gapcount, (int)i); gapcount++; in = false; }
            }
        }
    }
    RawFasta fa;
    if (!loadFastaRaw(contigFile, fa)) { fprintf(stderr, "Can't open contig file\n"); return 1; }
    const size_t nContigs = fa.seq.size();
    std::vector<unsigned long> contigLengths(nContigs);
    for (size_t i = 0; i < nContigs; i++) contigLengths[i] = fa.seq[i].size();

    // ---- gapInfo.txt (Preprocess.cpp:2091-2154).  The run state is NOT reset at a scaffold end: an N-run that reaches the end
    // of a scaffold is reported when the next scaffold shows a base, under that scaffold's number.
    std::vector<PGap> gaps;
    {
        FILE* gi = fopen((tmpDir + "gapInfo.txt").c_str(), "w");
        if (!gi) { fprintf(stderr, "Can't create gapInfo.txt\n"); return 1; }
        bool in = false; long runStart = 0; int runLen = 0, gapcount = 0;
        for (size_t i = 0; i < nContigs; i++) {
            const std::string& c = fa.seq[i];
            for (size_t j = 0; j < c.size(); j++) {
                if (c[j] == 'N' || c[j] == 'n') { if (!in) { in = true; runLen = 1; runStart = (long)j; } else runLen++; }
                else if (in) {
                    gaps.push_back(PGap{(int)i, runStart, runLen});
                    int toWrite = (int)i;
                    if (genomeReduction == 1) { auto it = contigOfGap.find(gapcount); if (it != contigOfGap.end()) toWrite = it->second; }
                    fprintf(gi, "%d\t%ld\t%d\n", toWrite, runStart, runLen);
                    gapcount++; in = false;
                }
            }
        }
        fclose(gi);
    }
    const size_t nGaps = gaps.size();
    FILE* statFile = fopen((tmpDir + "stat.txt").c_str(), "w");
    FILE* statFile2 = fopen((tmpDir + "stat2.txt").c_str(), "w");
    FILE* myout = fopen(outName.c_str(), "w");
    if (!myout) { fprintf(stderr, "Can't create myout file\n"); if (statFile) fclose(statFile); if (statFile2) fclose(statFile2); return 1; }
    if (!statFile || !statFile2) { if (statFile) fclose(statFile); if (statFile2) fclose(statFile2); fclose(myout); return 1; }
    setvbuf(myout, nullptr, _IOFBF, 1 << 20);

    // ---- reduced read set (Preprocess.cpp:2261-2302)
    const bool writeReduced = def == 1 ? readReduction == 1 : (readReduction == 1 && samflag == 1);
    FILE *red1 = nullptr, *red2 = nullptr;
    if (writeReduced) {
        const std::string s1(argv[7]), s2(argv[8]);
        const size_t f1 = s1.find_last_of("."), f2 = s2.find_last_of(".");
        if (f1 == std::string::npos) { fprintf(stderr, "read file name without extension\n"); return 1; }
        std::string o1 = s1.substr(0, f1), o2 = s2.substr(0, f2);
        const std::string ext = s1.substr(f1, o1.size());
        o1 += "_reduced" + ext; o2 += "_reduced" + ext;
        red1 = fopen(o1.c_str(), "w"); red2 = fopen(o2.c_str(), "w");
        if (!red1 || !red2) { fprintf(stderr, "Can't create reduced read pair during preproscessing...exiting.\n"); return 1; }
        printf("%s\n%s\n", o1.c_str(), o2.c_str());
        fflush(stdout);
    }

    Mapped sam;
    if (!sam.open(mapFile)) { fprintf(stderr, "Can't open alignment file\n"); return 1; }

    ContigIndex cidx;
    for (size_t i = 0; i < nContigs && i < fa.names.size(); i++) cidx.byName.emplace(sv(fa.names[i]), (long)i);      // first of equal names wins
    GapIndex gidx; gidx.build(gaps, nContigs); gidx.maxDistance = maxDistance;
    Config cfg; cfg.samflag = samflag; cfg.maxDistance = maxDistance; cfg.writeReduced = writeReduced; cfg.contigs = &cidx; cfg.contigLengths = &contigLengths; cfg.gidx = &gidx;

    const bool jump = samflag == 2 && maxDistance > 250;
    struct Totals { long totalCount = 0, unCount = 0; unsigned long maxReadLength = 0; std::vector<std::vector<Cand>> cands; };
    auto passes = [&](Totals& out) -> bool {
        out = Totals();
        if (jump) {
            // first pass: myout.sam from the properly paired reads, then the mean insert from that file (Preprocess.cpp:2313-2429)
            Config c1 = cfg; c1.modelPass = true; c1.writeReduced = false;
            PassResult p1;
            if (!runPass(c1, sam, threads, myout, nullptr, nullptr, p1)) return false;
            long insCount = p1.discarded; double sum = 0;
            if (!p1.insertHist.empty()) for (int i = 0; i < 5000; i++) { insCount += p1.insertHist[(size_t)i]; sum += (double)i * (double)p1.insertHist[(size_t)i]; }
            gidx.readMean = insCount != 0 ? (int)(sum / (double)insCount) : 0;
            Config c2 = cfg; c2.noVectors = true;
            PassResult p2;
            if (!runPass(c2, sam, threads, nullptr, red1, red2, p2)) return false;
            out.totalCount = p1.flushes + p2.mixedCount; out.unCount = p1.unCount + p2.unCount;
            out.maxReadLength = std::max(p1.maxReadLength, p2.maxReadLength);
            out.cands = std::move(p2.cands);
            return true;
        }
        PassResult p;
        if (!runPass(cfg, sam, threads, myout, red1, red2, p)) return false;
        out.totalCount = p.flushes + p.mixedCount; out.unCount = p.unCount; out.maxReadLength = p.maxReadLength; out.cands = std::move(p.cands);
        return true;
    };
    Totals total;
    std::vector<GapState> states;
    bool blocksOk = !(getenv("FIGBIRD_PP_SEQUENTIAL") && atoi(getenv("FIGBIRD_PP_SEQUENTIAL")) != 0);      // (tests: force the one-block path)
    if (blocksOk) blocksOk = passes(total);
    if (!blocksOk) {
        // irregular input: start over as one sequential block, every candidate meeting its gap's history at once
        fflush(myout); if (ftruncate(fileno(myout), 0) != 0) { } rewind(myout);
        if (red1) { fflush(red1); if (ftruncate(fileno(red1), 0) != 0) { } rewind(red1); }
        if (red2) { fflush(red2); if (ftruncate(fileno(red2), 0) != 0) { } rewind(red2); }
        states.assign(nGaps, GapState());
        cfg.inlineStates = &states;
        if (!passes(total)) { fprintf(stderr, "figbird_b200: cannot process the alignment file\n"); return 1; }
    }
    fclose(myout);
    if (red1) fclose(red1);
    if (red2) fclose(red2);

    // ---- per-gap history and files
    std::atomic<bool> ioOk(true);
    if (states.empty()) {
        states.assign(nGaps, GapState());
        std::vector<std::vector<const Cand*>> perGap(nGaps);
        for (auto& blk : total.cands) for (const Cand& c : blk) if (c.gap >= 0 && (size_t)c.gap < nGaps) perGap[(size_t)c.gap].push_back(&c);
        parallelFor(nGaps, threads, [&](size_t g) {
            for (const Cand* c : perGap[g]) states[g].offer(samflag, *c);
            GapState& st = states[g];
            std::vector<std::string>().swap(st.stored); st.seen.clear(); st.windows.clear();
        });
    }
    const bool container = getenv("FIGBIRD_CONTAINER") && atoi(getenv("FIGBIRD_CONTAINER")) != 0;
    parallelFor(nGaps, threads, [&](size_t g) {
        const std::string path = gapsDir + (samflag == 2 ? "gaps_" : "partial_gaps_") + std::to_string(g) + ".sam";
        if (!writeWhole(path, states[g].text)) ioOk = false;
        if (!container) std::string().swap(states[g].text);
    });
    if (container) {      // (opt-in) the same bytes once more, all gaps in one file that fb_fillgaps_main maps
        std::vector<std::string> texts(nGaps);
        for (size_t g = 0; g < nGaps; g++) texts[g].swap(states[g].text);
        if (!writeGapContainer(gapsDir + (samflag == 2 ? "gaps.fbc" : "partial_gaps.fbc"), samflag == 2 ? 2u : 1u, texts)) ioOk = false;
    }
    if (!ioOk) { fprintf(stderr, "figbird_b200: cannot write the per-gap files under %s\n", gapsDir.c_str()); return 1; }
    fprintf(statFile, "%ld %ld %ld %ld", total.totalCount, total.unCount, (long)total.maxReadLength, 5000L);
    for (size_t g = 0; g < nGaps; g++) fprintf(statFile2, "%d\t%d\t%d\n", 1, states[g].perfect, states[g].perfectLen);
    fclose(statFile); fclose(statFile2);
    return 0;
}

}  // namespace fb
