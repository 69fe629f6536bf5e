// fb_tools.cpp -- the small text tools around the gap-fill path, same argv and byte-identical output
// files as the reference executables RunFigbird.sh builds with bare g++:
//   fb_combinegaps_main  CombineGaps.cpp:169-313   (RunFigbird.sh:777)      gapout_<itr>.txt -> combined_gapstring.txt, Individual_gaps.txt
//   fb_flanktrim_main    FlankTrim.cpp:22-233      (RunFigbird.sh:254,433)  N-out `trim` bases on both sides of small gaps
//   fb_reduce_scf_main   Reduce_SCF.cpp:16-152     (RunFigbird.sh:266,320)  keep only the scaffolds that still hold an N
//   fb_reverse_main      Reverse.cpp:42-120        (RunFigbird.sh:166)      reverse-complement both FASTQ files of a jump library
// Host-only (no device work): they are the callers / consumers on either side of FillGaps.  What is reproduced on purpose is
// the reference's record chunking (fgets into 1024- / 10024-byte buffers) because lengths and counts depend on it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "fb_tools.h"

namespace fb {

bool slurp(const std::string& path, std::string& out) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    out.resize((size_t)st.st_size);
    size_t off = 0;
    while (off < out.size()) { const ssize_t k = read(fd, &out[off], out.size() - off); if (k <= 0) break; off += (size_t)k; }
    close(fd);
    out.resize(off);
    return true;
}

// The records fgets(buf, cap, f) would return: physical lines cut into pieces of at most cap-1 bytes.
template <class F>
static void forEachRecord(const std::string& text, size_t cap, F f) {
    const size_t CH = cap - 1;
    size_t p = 0;
    while (p < text.size()) {
        const char* e = (const char*)memchr(text.data() + p, '\n', text.size() - p);
        const size_t lineEnd = e ? (size_t)(e - text.data()) + 1 : text.size();
        for (size_t c = p; c < lineEnd; c += CH) f(text.data() + c, std::min(CH, lineEnd - c));
        p = lineEnd;
    }
}

// The FASTA reader all reference programs share (Preprocess.cpp:2019-2083, FlankTrim.cpp:66-133, Reduce_SCF.cpp:57-135):
// 1023-byte records; a record that starts with ';' is skipped, one that starts with '>' is a header; a sequence record shorter
// than 1023 bytes loses its last byte (the newline), a full one is kept whole.  A finished scaffold is pushed only when it is
// non-empty (names are pushed always), the last one always.  Case is kept.
bool loadFastaRaw(const std::string& path, RawFasta& out) {
    std::string text;
    if (!slurp(path, text)) return false;
    out.names.clear(); out.headers.clear(); out.seq.clear(); out.hasN.clear();
    std::string cur; bool curN = false;
    forEachRecord(text, 1024, [&](const char* r, size_t n) {
        if (r[0] == ';') return;
        if (r[0] == '>') {
            std::string h(r + 1, n >= 2 ? n - 2 : 0);        // without '>' and the record's last byte
            const size_t nul = h.find('\0'); if (nul != std::string::npos) h.resize(nul);
            const size_t a = h.find_first_not_of(" \t\n"), b = a == std::string::npos ? a : h.find_first_of(" \t\n", a);
            out.names.push_back(a == std::string::npos ? std::string() : h.substr(a, b == std::string::npos ? b : b - a));
            if (!cur.empty()) { out.seq.push_back(std::move(cur)); out.hasN.push_back(curN); out.headers.push_back(out.pendingHeader); cur.clear(); curN = false; }
            out.pendingHeader = h;
            return;
        }
        if (!curN) for (size_t i = 0; i < n; i++) if (r[i] == 'N' || r[i] == 'n') { curN = true; break; }       // (the dropped byte is looked at too)
        cur.append(r, n < 1023 ? n - 1 : n);
    });
    out.seq.push_back(std::move(cur)); out.hasN.push_back(curN); out.headers.push_back(out.pendingHeader);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reduce_SCF <gapped genome> <temp dir/>  ->  <temp dir/>newgenome.fa with the scaffolds that contain an N (Reduce_SCF.cpp:16-152)
// ---------------------------------------------------------------------------------------------------------------------
static int reduceScfMain(int argc, const char* const* argv) {
    if (argc < 3) { fprintf(stderr, "usage: reduce_scf <gapped genome> <temp dir/>\n"); return 1; }
    RawFasta fa;
    if (!loadFastaRaw(argv[1], fa)) { fprintf(stderr, "Can't open gapped genome file during reduction\n"); return 1; }
    FILE* o = fopen((std::string(argv[2]) + "newgenome.fa").c_str(), "w");
    if (!o) return 1;
    // the reference writes the header text as it stood when the scaffold was flushed: the whole line after '>' (Reduce_SCF.cpp:87-89)
    for (size_t i = 0; i < fa.seq.size(); i++)
        if (fa.hasN[i]) { fputc('>', o); fputs(fa.headers[i].c_str(), o); fputc('\n', o); fwrite(fa.seq[i].data(), 1, fa.seq[i].size(), o); fputc('\n', o); }
    fclose(o);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// FlankTrim <gapped genome> <trim> <readlen> <out>  (FlankTrim.cpp:22-233): for every N-run of 2 .. readlen-1 bases that is more
// than 3*trim from the scaffold start and more than 2*trim from its end, and whose `trim` flank bases on both sides hold no 'N',
// those flank bases become 'N'.  The scan then skips trim+1 bases (the reference's `j += trimsize` inside a for loop).
// ---------------------------------------------------------------------------------------------------------------------
static int flankTrimMain(int argc, const char* const* argv) {
    if (argc < 5) { fprintf(stderr, "usage: flanktrim <gapped genome> <trim> <readlen> <trimmed genome>\n"); return 1; }
    const int trim = atoi(argv[2]), readlen = atoi(argv[3]);
    RawFasta fa;
    if (!loadFastaRaw(argv[1], fa)) { printf("Can't open gapped genome file\n"); return 1; }
    FILE* o = fopen(argv[4], "w");
    if (!o) return 1;
    for (size_t i = 0; i < fa.seq.size(); i++) {
        std::string& c = fa.seq[i];
        const long len = (long)c.size();
        fprintf(o, ">%s\n", i < fa.names.size() ? fa.names[i].c_str() : "");
        bool inRun = false; long runStart = 0, runLen = 0;
        for (long j = 0; j < len; j++) {
            const bool isN = c[j] == 'N' || c[j] == 'n';
            if (isN) { if (!inRun) { inRun = true; runLen = 1; runStart = j; } else runLen++; }
            if ((!isN && inRun) || (isN && j == len - 1)) {
                if (trim > 0 && runLen > 1 && (int)runLen < readlen && (int)runStart - trim > 2 * trim && (unsigned long)(len - runStart - runLen) > (unsigned long)(2 * trim)) {
                    bool clean = true;      // strpbrk(flank, "N"): upper case only
                    for (int t = 0; t < trim && clean; t++) if (c[runStart - 1 - t] == 'N' || c[runStart + runLen + t] == 'N') clean = false;
                    if (clean) { for (int t = 0; t < trim; t++) { c[runStart - 1 - t] = 'N'; c[runStart + runLen + t] = 'N'; } j += trim; }
                }
                inRun = false;
            }
        }
        fwrite(c.data(), 1, c.size(), o); fputc('\n', o);
    }
    fclose(o);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reverse <reads_1.fq> <reads_2.fq>  (Reverse.cpp:42-120): writes <stem>_reversed<ext> for both files, every 2nd record of four
// reverse-complemented (a record = an fgets chunk of at most 1023 bytes; its last byte is dropped before reversing), and prints
// the two new paths.  `ext` is cut from the first path and is at most as long as its stem (Reverse.cpp:65).
// ---------------------------------------------------------------------------------------------------------------------
static int reverseMain(int argc, const char* const* argv) {
    if (argc < 3) { fprintf(stderr, "usage: reverse <reads_1.fq> <reads_2.fq>\n"); return 1; }
    std::string t1, t2;
    if (!slurp(argv[1], t1) || !slurp(argv[2], t2)) { fprintf(stderr, "Can't open read pair files during reversing...exiting.\n"); return 1; }
    const std::string s1(argv[1]), s2(argv[2]);
    const size_t f1 = s1.find_last_of("."), f2 = s2.find_last_of(".");
    if (f1 == std::string::npos) { fprintf(stderr, "read file name without extension\n"); return 1; }      // (the reference aborts in substr)
    std::string o1 = s1.substr(0, f1), o2 = s2.substr(0, f2);
    const std::string ext = s1.substr(f1, o1.size());
    o1 += "_reversed" + ext; o2 += "_reversed" + ext;
    FILE* w1 = fopen(o1.c_str(), "w"); FILE* w2 = fopen(o2.c_str(), "w");
    if (!w1 || !w2) { fprintf(stderr, "Can't create new read pair files during reversing...exiting.\n"); if (w1) fclose(w1); if (w2) fclose(w2); return 1; }
    struct Rec { const char* p; size_t n; };
    std::vector<Rec> r1, r2;
    forEachRecord(t1, 1024, [&](const char* p, size_t n) { r1.push_back(Rec{p, n}); });
    forEachRecord(t2, 1024, [&](const char* p, size_t n) { r2.push_back(Rec{p, n}); });
    std::string buf;
    auto emit = [&](FILE* w, const Rec& r, bool rc) {
        // records are C strings for the reference: an embedded NUL would end them; FASTQ text has none
        if (!rc) { fwrite(r.p, 1, r.n, w); return; }
        const size_t L = r.n - 1;
        buf.resize(L);
        for (size_t i = 0; i < L; i++) {
            const char ch = r.p[i]; char o;
            switch (ch) { case 'A': case 'a': o = 'T'; break; case 'C': case 'c': o = 'G'; break; case 'G': case 'g': o = 'C'; break; case 'T': case 't': o = 'A'; break; default: o = 'N'; }
            buf[L - 1 - i] = o;
        }
        fwrite(buf.data(), 1, L, w); fputc('\n', w);
    };
    const size_t n = std::min(r1.size(), r2.size());
    for (size_t i = 0; i < n; i++) { emit(w1, r1[i], i % 4 == 1); emit(w2, r2[i], i % 4 == 1); }
    fclose(w1); fclose(w2);
    printf("%s\n%s\n", o1.c_str(), o2.c_str());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// CombineGaps <iterations> <dir/>  (CombineGaps.cpp:169-313).  gapout_<itr>.txt holds one line per gap that was still open
// when iteration itr ran ("gapNo contigNo gapStart gapLength gapStringLength gapString", FillGaps.cpp:193-219); a gap's string
// may keep one N-run, which the next iteration's string replaces (combine(), CombineGaps.cpp:65-124).  Out: combined_gapstring.txt
// (one line per gap, rewritten after every iteration) and Individual_gaps.txt.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct OpenGap {
    std::string s;
    bool closed = false;
    int leftN = -1, rightN = -1, rSize = 0, original = 0, finalLen = 0;
};
// first / last index of the LAST N-run of a string and the number of N-runs (checkComplete, CombineGaps.cpp:31-63); the positions
// live in globals there and keep their previous values when the string has no N
struct NRuns { int first = 0, last = 0; };
int scanRuns(const std::string& g, NRuns& pos) {
    int runs = 0; bool in = false;
    const int len = (int)g.size();
    for (int i = 0; i < len; i++) {
        if (g[i] == 'N' && !in) { in = true; pos.first = i; }
        else if (g[i] != 'N' && in) { in = false; runs++; pos.last = i - 1; }
        if (i == len - 1 && in) { runs++; pos.last = i; }
    }
    return runs;
}
// whitespace-separated tokens, as the reference's fscanf("%d\t...") / fscanf("%s\n") calls consume them
struct Tokens {
    const std::string& t; size_t p = 0;
    explicit Tokens(const std::string& text) : t(text) {}
    bool next(std::string& out) {
        while (p < t.size() && isspace((unsigned char)t[p])) p++;
        if (p >= t.size()) return false;
        const size_t a = p;
        while (p < t.size() && !isspace((unsigned char)t[p])) p++;
        out.assign(t, a, p - a);
        return true;
    }
};
}  // namespace

static int combineGapsMain(int argc, const char* const* argv) {
    if (argc < 3) { fprintf(stderr, "usage: combinegaps <iterations> <dir/>\n"); return 1; }
    const int iters = atoi(argv[1]);
    const std::string dir(argv[2]);
    std::vector<OpenGap> gaps;
    NRuns pos;
    auto writeCombined = [&]() -> bool {
        FILE* f = fopen((dir + "combined_gapstring.txt").c_str(), "w");
        if (!f) return false;
        for (const OpenGap& g : gaps) { fwrite(g.s.data(), 1, g.s.size(), f); fputc('\n', f); }
        fclose(f);
        return true;
    };
    for (int itr = 1; itr <= iters; itr++) {
        std::string text;
        if (!slurp(dir + "gapout_" + std::to_string(itr) + ".txt", text)) { printf("Can't open gapout txt file\n"); return 1; }
        if (itr == 1) { size_t n = 0; forEachRecord(text, 10024, [&](const char*, size_t) { n++; }); gaps.assign(n, OpenGap()); }
        Tokens tk(text);
        std::string tok;
        int gapLength = 0, strLen = 0;
        for (OpenGap& g : gaps) {
            if (g.closed) continue;      // closed in an earlier iteration: the draft of this iteration no longer had that gap
            int v[5] = {gapLength, gapLength, gapLength, gapLength, strLen};
            for (int k = 0; k < 5; k++) { if (!tk.next(tok)) break; v[k] = atoi(tok.c_str()); }
            gapLength = v[3]; strLen = v[4];
            if (strLen > 0) {
                if (!tk.next(tok)) tok.clear();
                const int runs = scanRuns(tok, pos);
                if (runs > 1) return 0;      // several N-runs in one gap string: the reference gives up silently with status 0 (CombineGaps.cpp:252-256)
                if (itr == 1) g.original = gapLength;
                g.closed = runs == 0;
                if (itr == 1) {
                    g.s = tok; g.finalLen = strLen;
                    if (!g.closed) { scanRuns(g.s, pos); g.leftN = pos.first; g.rightN = pos.last; g.rSize = strLen - g.rightN; }
                } else {
                    // left part up to the old N-run + the new string + what followed the old N-run
                    const int newlen = g.leftN + strLen + g.rSize;
                    std::string ns;
                    ns.reserve(newlen > 0 ? newlen : 0);
                    ns.append(g.s, 0, (size_t)std::max(0, g.leftN));
                    ns.append(tok, 0, (size_t)strLen); if ((int)tok.size() < strLen) ns.append((size_t)(strLen - (int)tok.size()), '\0');
                    if (g.rSize - 1 > 0) ns.append(g.s, (size_t)(g.rightN + 1), (size_t)(g.rSize - 1));
                    const size_t nul = ns.find('\0'); if (nul != std::string::npos) ns.resize(nul);
                    g.s.swap(ns);
                    scanRuns(g.s, pos);
                    g.leftN = pos.first; g.rightN = pos.last; g.finalLen = newlen - 1; g.rSize = g.finalLen - g.rightN;
                }
            } else {
                // a gap closed with length 0 (negative overlap): its record restarts as in iteration 1
                g.original = gapLength; g.closed = true; g.s.clear(); g.finalLen = 0;
            }
        }
        if (!writeCombined()) return 1;
    }
    FILE* o = fopen((dir + "Individual_gaps.txt").c_str(), "w");
    if (!o) return 1;
    std::string combined;
    slurp(dir + "combined_gapstring.txt", combined);
    Tokens tk(combined);
    std::string line;
    fprintf(o, "GapNo\tOriginal_Length\tFilled_Length\n\n");
    for (size_t i = 0; i < gaps.size(); i++) {
        if (gaps[i].finalLen > 0) { std::string t; if (tk.next(t)) line = t; } else line.clear();
        fprintf(o, "%d\t%d\t%d\t%s\n", (int)i, gaps[i].original, gaps[i].finalLen, line.c_str());
    }
    fclose(o);
    return 0;
}

int preprocessMain(int argc, const char* const* argv);      // fb_preprocess.cpp

}  // namespace fb

#define FB_TOOL_ENTRY(name, fn) \
    extern "C" int32_t name(int32_t argc, const char* const* argv) { \
        try { return fb::fn(argc, argv); } catch (const std::exception& e) { fprintf(stderr, "figbird_b200: %s\n", e.what()); return 1; } \
        catch (...) { fprintf(stderr, "figbird_b200: tool failed\n"); return 1; } }
FB_TOOL_ENTRY(fb_combinegaps_main, combineGapsMain)
FB_TOOL_ENTRY(fb_flanktrim_main, flankTrimMain)
FB_TOOL_ENTRY(fb_reduce_scf_main, reduceScfMain)
FB_TOOL_ENTRY(fb_reverse_main, reverseMain)
FB_TOOL_ENTRY(fb_preprocess_main, preprocessMain)
