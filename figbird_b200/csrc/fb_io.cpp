// fb_io.cpp -- file formats on the gap-fill path: scaffold FASTA, gapInfo/stat2,
// per-gap read files in; gapout.txt / filledContigs.fa / Ncount.txt / draw.txt out.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "fb_host.h"
#include "fb_io.h"

namespace fb {

// The reference reads the FASTA with fgets into a 1024-byte buffer (Figbird.cpp:6986-7046, same code in
// FillGaps.cpp:716-771): a chunk shorter than 1023 chars loses its last char (the newline); a chunk of
// exactly 1023 chars is kept whole -- even when its last char is the newline.  Scaffold lengths and hence
// every coordinate depend on that, so the loader works on the same 1023-char chunks.
bool loadScaffolds(const std::string& path, Scaffolds& out) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat stt;
    if (fstat(fd, &stt) != 0) { close(fd); return false; }
    const size_t total = (size_t)stt.st_size;
    const char* text = nullptr; void* mapped = nullptr; std::string owned;
    if (total > 0) {
        mapped = mmap(nullptr, total, PROT_READ, MAP_PRIVATE, fd, 0);
        if (mapped != MAP_FAILED) text = (const char*)mapped;
        else { mapped = nullptr; owned.resize(total); size_t off = 0; while (off < total) { ssize_t k = pread(fd, &owned[off], total - off, (off_t)off); if (k <= 0) break; off += (size_t)k; } if (off != total) { close(fd); return false; } text = owned.data(); }
    }
    close(fd);
    // Pass 1 (sequential, per physical line, no copying): which bytes go to which scaffold.  A chunk that starts with '>' is a
    // header, one that starts with ';' is skipped; a sequence chunk shorter than 1023 chars loses its last char.
    struct Seg { const char* p; size_t n; };
    std::vector<std::vector<Seg>> segs(1);
    std::vector<char> pushed(1, 0);        // the reference pushes a finished scaffold only when it is non-empty, and the last one always
    const size_t CH = 1023;
    size_t curLen = 0;
    auto addSeg = [&](const char* p, size_t n) { if (n) { auto& v = segs.back(); if (!v.empty() && v.back().p + v.back().n == p) v.back().n += n; else v.push_back(Seg{p, n}); curLen += n; } };
    size_t p = 0;
    while (p < total) {
        const char* e = (const char*)memchr(text + p, '\n', total - p);
        const size_t lineEnd = e ? (size_t)(e - text) + 1 : total;   // physical line incl. '\n'
        const size_t n = lineEnd - p;
        bool special = false;      // any chunk of this line starts with '>' or ';'
        for (size_t c = p; c < lineEnd; c += CH) if (text[c] == '>' || text[c] == ';') { special = true; break; }
        if (!special) addSeg(text + p, (n % CH) == 0 ? n : n - 1);
        else for (size_t c = p; c < lineEnd; c += CH) {
            const size_t m = std::min(CH, lineEnd - c);
            const char* chunk = text + c;
            if (chunk[0] == ';') continue;
            if (chunk[0] == '>') {
                std::string nm(chunk + 1, m >= 2 ? m - 2 : 0);          // drop '>' and the last char
                size_t a = nm.find_first_not_of(" \t\n"), b = a == std::string::npos ? a : nm.find_first_of(" \t\n", a);
                out.names.push_back(a == std::string::npos ? std::string() : nm.substr(a, b == std::string::npos ? b : b - a));
                if (curLen > 0) { pushed.back() = 1; segs.emplace_back(); pushed.push_back(0); curLen = 0; }
            } else addSeg(chunk, m < CH ? m - 1 : m);
        }
        p = lineEnd;
    }
    pushed.back() = 1;
    // Pass 2 (parallel): copy + toupper() of the C locale (the reference never calls setlocale)
    const size_t nS = segs.size();
    out.seq.assign(nS, std::string());
    struct Job { size_t s; size_t dst; const char* p; size_t n; };
    std::vector<Job> jobs;
    out.totalLength = 0;
    const size_t BLK = 4u << 20;
    for (size_t i = 0; i < nS; i++) {
        size_t len = 0; for (auto& g : segs[i]) len += g.n;
        out.seq[i].resize(len);
        out.totalLength += (long)len;
        size_t d = 0;
        for (auto& g : segs[i]) { for (size_t o = 0; o < g.n; o += BLK) jobs.push_back(Job{i, d + o, g.p + o, std::min(BLK, g.n - o)}); d += g.n; }
    }
    auto runJob = [&](const Job& j) { char* d = &out.seq[j.s][j.dst]; const char* sp = j.p; for (size_t i = 0; i < j.n; i++) { const char ch = sp[i]; d[i] = (ch >= 'a' && ch <= 'z') ? (char)(ch - 32) : ch; } };
    int nt = (int)std::thread::hardware_concurrency();
    if (const char* e2 = getenv("FIGBIRD_HOST_THREADS")) nt = atoi(e2);
    nt = std::max(1, std::min(nt, (int)((out.totalLength >> 22) + 1)));
    if (nt <= 1) for (auto& j : jobs) runJob(j);
    else {
        std::atomic<size_t> next(0);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([&] { for (size_t i; (i = next++) < jobs.size();) runJob(jobs[i]); });
        for (auto& t : th) t.join();
    }
    if (mapped) munmap(mapped, total);
    return true;
}

static bool readLines(const std::string& path, std::vector<std::string>& lines) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[1024];   // MAX_REC_LEN: a longer physical line arrives as several fgets chunks, like the reference sees it
    while (fgets(buf, sizeof buf, f)) lines.emplace_back(buf);
    fclose(f);
    return true;
}

bool loadGapRecords(const std::string& tmpDir, std::vector<GapRecord>& gaps, int& totGaps) {
    std::vector<std::string> gi, s2;
    if (!readLines(tmpDir + "gapInfo.txt", gi)) return false;
    readLines(tmpDir + "stat2.txt", s2);
    totGaps = (int)gi.size();
    size_t n = std::min(gi.size(), s2.size());   // Figbird.cpp:7329 stops at the shorter file
    for (size_t i = 0; i < n; i++) {
        GapRecord r; r.gapNo = (int)i;
        long a = 0, b = 0, c = 0;
        sscanf(gi[i].c_str(), "%ld\t%ld\t%ld", &a, &b, &c);
        r.contigNo = (int)a; r.gapStart = b; r.gapLength = (int)c;
        int x = 0, y = 0, z = 0;
        sscanf(s2[i].c_str(), "%d\t%d\t%d", &x, &y, &z);
        r.stat1 = x; r.stat2 = y; r.stat3 = z;
        gaps.push_back(r);
    }
    return true;
}

// ---- per-gap read files.  A file is read with one call and walked in place: records as fgets(buf, 1024) returns them (a physical
// line of 1023 bytes or more arrives in pieces), fields as strtok(line, "\t") yields them (runs of tabs collapse, the newline
// stays in the last field).
namespace {
struct FileText {
    std::string buf;
    bool read(const std::string& path) {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return false; }
        buf.resize((size_t)st.st_size);
        size_t off = 0;
        while (off < buf.size()) { const ssize_t k = ::read(fd, &buf[off], buf.size() - off); if (k <= 0) break; off += (size_t)k; }
        close(fd);
        buf.resize(off);
        return true;
    }
};
struct View { const char* p; size_t n; bool empty() const { return n == 0; } };
struct Records {
    const char* p; const char* e;
    explicit Records(const std::string& t) : p(t.data()), e(t.data() + t.size()) {}
    Records(const char* b, size_t n) : p(b), e(b + n) {}
    bool next(View& r) {
        if (p >= e) return false;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl + 1 : e;
        if (le - p > 1023) le = p + 1023;
        r.p = p; r.n = (size_t)(le - p); p = le;
        // a NUL byte ends the C string the reference works on
        const void* z = memchr(r.p, 0, r.n); if (z) r.n = (size_t)((const char*)z - r.p);
        return true;
    }
};
// up to `want` tab-separated fields of a record; returns how many there are in total (counting stops at 16)
int fields(const View& r, View* out, int want) {
    int n = 0; size_t q = 0;
    while (q < r.n && n < 16) {
        while (q < r.n && r.p[q] == '\t') q++;
        if (q >= r.n) break;
        size_t e = q; while (e < r.n && r.p[e] != '\t') e++;
        if (n < want) out[n] = View{r.p + q, e - q};
        n++; q = e;
    }
    return n;
}
int atoiView(const View& v) {
    size_t i = 0; while (i < v.n && (v.p[i] == ' ' || (v.p[i] >= '\t' && v.p[i] <= '\r'))) i++;
    bool neg = false; if (i < v.n && (v.p[i] == '+' || v.p[i] == '-')) { neg = v.p[i] == '-'; i++; }
    long x = 0; for (; i < v.n && v.p[i] >= '0' && v.p[i] <= '9'; i++) x = x * 10 + (v.p[i] - '0');
    return (int)(neg ? -x : x);
}
}  // namespace

// partial_gaps_<g>.sam: seq, clipped_index, match, pos, cigar, pos2, qual (Preprocess.cpp:454,466,478).
// Every reader in the reference stops after partial_limit+1 = 3001 lines (e.g. Figbird.cpp:1814,2014).
bool loadPartial(const std::string& path, std::vector<PartialRead>& out, bool& exists) {
    FileText ft;
    exists = ft.read(path);
    if (!exists) return false;
    loadPartialText(ft.buf.data(), ft.buf.size(), out);
    return true;
}
void loadPartialText(const char* text, size_t n_, std::vector<PartialRead>& out) {
    Records rec(text, n_);
    View r, t[7];
    while (rec.next(r)) {
        const int n = fields(r, t, 7);
        out.emplace_back();
        PartialRead& pr = out.back();
        if (n >= 1) pr.seq.assign(t[0].p, t[0].n);
        if (n >= 2) pr.clippedIndex = atoiView(t[1]);
        if (n >= 3) pr.match = atoiView(t[2]);
        if (n >= 4) pr.pos = atoiView(t[3]);
        if (n >= 6) pr.refPos = atoiView(t[5]);
        if (n >= 7) { pr.qual.assign(t[6].p, t[6].n); while (!pr.qual.empty() && (pr.qual.back() == '\n' || pr.qual.back() == '\r')) pr.qual.pop_back(); }
        if (n == 1) { while (!pr.seq.empty() && pr.seq.back() == '\n') pr.seq.pop_back(); }
        if (out.size() > 3000) break;
    }
}

static char rc(char ch) {   // reverse(), Figbird.cpp:1427-1449
    switch (ch) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; }
    return 'N';
}

// gaps_<g>.sam (parseUnmapped, Figbird.cpp:5661-5767): line pairs, mapped mate then unmapped mate.
// pairCount = findcount_file(.,0) (Figbird.cpp:6686-6711).
bool loadUnmapped(const std::string& path, int readLen, std::vector<UnmappedRead>& out, int& pairCount) {
    (void)readLen;
    FileText ft;
    pairCount = 0;
    if (!ft.read(path)) return false;
    loadUnmappedText(ft.buf.data(), ft.buf.size(), out, pairCount);
    return true;
}
void loadUnmappedText(const char* text, size_t n_, std::vector<UnmappedRead>& out, int& pairCount) {
    { Records all(text, n_); View r; size_t n = 0; while (all.next(r)) n++; pairCount = (int)(n / 2); }
    Records rec(text, n_);
    View l1, l2, t1[4], t2[8];
    int total = 0;
    while (rec.next(l1)) {
        if (l1.n > 0 && l1.p[0] == '@') continue;
        if (l1.n < 60) continue;
        if (!rec.next(l2)) break;
        if (fields(l1, t1, 4) < 4 || fields(l2, t2, 8) < 8) continue;
        out.emplace_back();
        UnmappedRead& u = out.back();
        const int strand1 = (atoiView(t1[1]) & 16) >> 4;
        u.matePos = atoiView(t1[3]);
        if (strand1 == 0) {
            u.seq.resize(t2[6].n);
            for (size_t k = 0; k < t2[6].n; k++) u.seq[t2[6].n - 1 - k] = rc(t2[6].p[k]);
            u.isReverse = 1;
        } else { u.seq.assign(t2[6].p, t2[6].n); u.isReverse = 0; }
        if (++total == 3000) break;   // unmapped_limit
    }
}

// ---- per-gap container (opt-in, FIGBIRD_CONTAINER=1): what fb_preprocess_main writes beside the per-gap text files
// when asked to -- the same bytes, all gaps in one file: "FBGAPS1\0", kind (1 partial_gaps / 2 gaps), number of gaps, nGaps + 1
// offsets, then the text of gap 0, 1, ...  The text files stay the contract; the container saves a reader 2 x 10^4 opens.
bool GapContainer::open(const std::string& path, uint32_t kind, size_t nGaps) {
    close();
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < 16) { ::close(fd); return false; }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) return false;
    base_ = (const char*)m; size_ = (size_t)st.st_size;
    uint32_t k = 0, n = 0;
    memcpy(&k, base_ + 8, 4); memcpy(&n, base_ + 12, 4);
    const size_t head = 16 + 8 * ((size_t)n + 1);
    bool ok = memcmp(base_, "FBGAPS1", 8) == 0 && k == kind && (size_t)n == nGaps && head <= size_;
    if (ok) {
        off_ = (const uint64_t*)(base_ + 16); n_ = n;
        for (size_t g = 0; g < n_ && ok; g++) ok = off_[g] <= off_[g + 1];
        ok = ok && off_[0] == head && off_[n_] == size_;
    }
    if (!ok) close();
    return ok;
}
void GapContainer::close() { if (base_) munmap((void*)base_, size_); base_ = nullptr; off_ = nullptr; n_ = 0; size_ = 0; }
bool writeGapContainer(const std::string& path, uint32_t kind, const std::vector<std::string>& texts) {
    std::vector<uint64_t> off(texts.size() + 1);
    const uint64_t head = 16 + 8 * (uint64_t)(texts.size() + 1);
    off[0] = head;
    for (size_t g = 0; g < texts.size(); g++) off[g + 1] = off[g] + texts[g].size();
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t n = (uint32_t)texts.size();
    bool ok = fwrite("FBGAPS1", 1, 8, f) == 8 && fwrite(&kind, 4, 1, f) == 1 && fwrite(&n, 4, 1, f) == 1 && fwrite(off.data(), 8, off.size(), f) == off.size();
    for (size_t g = 0; g < texts.size() && ok; g++) ok = texts[g].empty() || fwrite(texts[g].data(), 1, texts[g].size(), f) == texts[g].size();
    return fclose(f) == 0 && ok;
}

// ---- writers -------------------------------------------------------------------------------------------

bool writeGapout(const std::string& path, const std::vector<GapRecord>& gaps, const std::vector<GapResult>& res) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    for (size_t i = 0; i < gaps.size(); i++) {
        // Figbird.cpp:7413 then re-emitted by FillGaps.cpp:204-207 (an empty string leaves a bare tab)
        fprintf(f, "%d\t%d\t%ld\t%d\t%d\t", gaps[i].gapNo, gaps[i].contigNo, gaps[i].gapStart, gaps[i].gapLength, res[i].gapStringLength);
        if (res[i].gapStringLength > 0) fprintf(f, "%s\n", res[i].gapString.c_str()); else fprintf(f, "\n");
    }
    fclose(f);
    return true;
}

// FillGaps.cpp:820-926: splice the gap strings into the scaffolds, trim after negative-overlap gaps,
// and report whether any N is left (with the stale-string quirk of :861-870).
bool writeFilledContigs(const std::string& tmpDir, const Scaffolds& sc, const std::vector<GapRecord>& gaps,
                        const std::vector<GapResult>& res, int totGaps) {
    FILE* out = fopen((tmpDir + "filledContigs.fa").c_str(), "w");
    FILE* nf = fopen((tmpDir + "Ncount.txt").c_str(), "w");
    if (!out || !nf) { if (out) fclose(out); if (nf) fclose(nf); return false; }
    std::vector<char> iobuf(8u << 20); setvbuf(out, iobuf.data(), _IOFBF, iobuf.size());
    std::vector<int> gtf(std::max(totGaps, (int)gaps.size()) + 1, 0);
    for (size_t i = 0; i < res.size(); i++) gtf[i] = res[i].gapToFill;
    int gapCount = -1; size_t nextEntry = 0;
    int nStart = 0;
    long newNcount = 0;
    std::string gapString;          // persists between gaps like the reference's stack buffer
    int gapStringLength = 0;
    std::string buf;
    for (size_t i = 0; i < sc.seq.size(); i++) {
        const std::string& s = sc.seq[i];
        fprintf(out, ">%s\n", i < sc.names.size() ? sc.names[i].c_str() : "");
        buf.clear();
        for (size_t j = 0; j < s.size(); j++) {
            if (nStart == 0 && !(gapCount >= 0 && gapCount < (int)gtf.size() && gtf[gapCount] > 0)) {
                // a run of ordinary bases outside any gap and past the trimmed flank: copied in one piece (the last base of
                // the scaffold and anything from the next N on take the per-character path below)
                // k = the next N / n at or after j, at most the last base (memchr: this loop is 100 Mbp of a C4 draft)
                size_t k = s.size() - 1;
                if (j < k) {
                    const char* b = s.data();
                    if (const void* pN = memchr(b + j, 'N', k - j)) k = (size_t)((const char*)pN - b);
                    if (const void* pn = memchr(b + j, 'n', k - j)) k = (size_t)((const char*)pn - b);
                } else k = j;
                if (k > j) { buf.append(s, j, k - j); j = k - 1; continue; }
            }
            bool isN = (s[j] == 'N' || s[j] == 'n');
            if (isN && nStart == 0) { nStart = 1; gapCount++; }
            if (!isN || j == s.size() - 1) {
                if (nStart == 1) {
                    if (nextEntry < res.size()) {
                        gapStringLength = res[nextEntry].gapStringLength;
                        if (gapStringLength > 0) gapString = res[nextEntry].gapString;
                        nextEntry++;
                    }
                    for (char ch : gapString) if (ch == 'N') newNcount++;
                    fwrite(buf.data(), 1, buf.size(), out);
                    if (gapStringLength > 0) fwrite(gapString.data(), 1, gapString.size(), out);
                    buf.clear();
                    nStart = 0;
                }
                if (gapCount >= 0 && gapCount < (int)gtf.size() && gtf[gapCount] > 0) gtf[gapCount]--;
                else buf.push_back(s[j]);
            }
        }
        fwrite(buf.data(), 1, buf.size(), out); fputc('\n', out);
    }
    fprintf(nf, "%d", newNcount == 0 ? 0 : 1);
    fclose(out); fclose(nf);
    return true;
}

}  // namespace fb
