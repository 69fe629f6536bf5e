// fb_gapfiller.cpp -- host control of one gap: which candidate lengths are evaluated, how the per-read
// results coming back from the device turn into likelihoods, and the final string.
//
// Everything that is *scored* (read x offset x base products, weighted votes, consensus, M-step) happens in
// the engine (include/figbird_b200.h); this file holds only the sequential decision logic of the reference
// worker, re-expressed over device results.  It has to reproduce the reference's discrete decisions exactly
// (the filled sequence is compared byte for byte), so thresholds, tie rules and even benign quirks follow
// GapFiller (Figbird.cpp:1563-6684) -- cited per function.  Window geometry (left/right_maxDistance and their clipping
// at scaffold ends) is not modelled: inside the supported domain (the +-readLength rows around the gap lie inside the
// scaffold, see prepare()) the reference's results do not depend on it.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "fb_gapfiller.h"

namespace fb {
namespace {

inline int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }   // charCodes, Figbird.cpp:7060-7082
inline char letter(int c) { return c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N'; }

// std::string::substr that clamps instead of throwing (the reference would abort on out_of_range)
std::string sub(const std::string& s, long pos, long n) {
    if (pos < 0 || pos > (long)s.size() || n < 0) return std::string();
    return s.substr((size_t)pos, (size_t)n);
}

// getDiff, Figbird.cpp:2429-2443
double diffFrac(const std::string& target, const std::string& ref, int length) {
    double diff = 0;
    for (int i = 0; i < length; i++) {
        char t = i < (int)target.size() ? target[i] : '\0', r = i < (int)ref.size() ? ref[i] : '\0';
        if (toupper((unsigned char)t) != r) diff++;
    }
    return diff / length;
}

// find_partial_match, Figbird.cpp:2445-2490
int partialMatch(const std::string& ref, const std::string& search, int pos, int c, int lenT) {
    const int lenR = (int)ref.size(), lenS = (int)search.size();
    const int lenThresh = c == 0 ? lenT : 4;
    if (lenR >= lenS && lenS >= lenThresh) {
        std::string s3 = pos == 0 ? ref.substr(lenR - lenS, lenS) : ref.substr(0, lenS);
        double frac = diffFrac(s3, search, lenS);
        if (frac < 0.2 && c == 1) return 1;
        if (frac <= 0.08 && c == 0) return 1;
    }
    return 0;
}

size_t cstrlen(const std::vector<char>& b) { return strnlen(b.data(), b.size()); }

}  // namespace

GapFill::GapFill(const Args& a, const Model& m, const Scaffolds& sc, GapInput&& in) : a_(a), m_(m), sc_(sc), in_(std::move(in)) {
    memset(ovl_, 0, sizeof ovl_);
    savedTemp_[0] = savedTemp_[1] = savedFinal_[0] = savedFinal_[1] = -1; sideLimit_ = 30;
}

// ---------------------------------------------------------------------------------------------------------
// prepare(): everything main() + allocate() + parse*() + analyzeGap() decide before the first placement
// (Figbird.cpp:7377-7400, 1638-1778, 6879-6906, 6168-6199), plus the device encoding of the gap.
// ---------------------------------------------------------------------------------------------------------
void GapFill::prepare() {
    const GapRecord& r = in_.rec;
    og_ = r.gapLength;
    unmReadLen_ = a_.readLength;
    partialReadLen_ = a_.partialReadLen;
    midLimitP_ = 2 * a_.partialReadLen;
    midLimitU_ = 400;                      // gapthresh, FillGaps.cpp:22 -> argv[14] of the worker
    negOverlap_ = a_.negOverlap;
    fillflag_ = 1;
    if (a_.unmapped == 1) {
        numReads_ = in_.unmPairCount;
        if (numReads_ > 3000) { numReads_ = 3000; fillflag_ = -1; }   // Figbird.cpp:7380-7385
        if ((int)in_.unm.size() < numReads_) numReads_ = (int)in_.unm.size();
    }
    if (a_.partialFlag == 1) partialReadCount_ = (int)std::min<size_t>(in_.partial.size(), 3001);
    // findFrac (Figbird.cpp:6879-6906); float arithmetic on purpose
    const int factor = 3 * a_.partialReadLen;
    int allocationFactor;
    largeGapFlag_ = 0;
    if (a_.partialFlag) {
        if (og_ <= midLimitP_ / 2) { frac1_ = .00001; frac2_ = (float)factor / og_; allocationFactor = -1; }
        else if (og_ <= midLimitP_) { frac1_ = .00001; frac2_ = 5.0; allocationFactor = 5; }
        else { frac1_ = 1; frac2_ = 1; allocationFactor = 3; }
    } else {
        if (og_ <= midLimitU_ / 3) { frac1_ = .3; frac2_ = (float)factor / og_; allocationFactor = -1; }
        else if (og_ <= midLimitU_) { frac1_ = .5; frac2_ = 2.5; allocationFactor = 3; }
        else { frac1_ = 1; frac2_ = 1; largeGapFlag_ = 1; allocationFactor = 1; }
    }
    allocArg_ = allocationFactor == -1 ? factor * 3 : og_ * allocationFactor;
    sideLimit_ = 30;

    const int maxGap = std::max(allocArg_, 1);
    concensus_.assign(maxGap + 1, 0); concensus_[0] = 'N';
    bestString_.assign(maxGap + 1, 0);
    originalStr_.assign(og_ + 1, 0);
    gapCoverage_.assign(maxGap, 0);
    // countsGap / qual_gap host copies (maxGap rows of 5 doubles each) are allocated where they are first used (large-gap rounds,
    // finalize): most gaps never need them before finalize, and 10^4 gaps x 2 tables x 40 B x maxGap is gigabytes
    counts_.clear(); qualGap_.clear();
    markAccepted_.assign(numReads_, 0); savedReads_.assign(numReads_, 0); mlvNonZero_.assign(numReads_, 0);
    finalReadpos_.assign(numReads_, Pos3{-200, 0, -1}); unmPosOrg_.assign(numReads_, Pos3{-200, 0, 0});
    partialPosOrg_.assign(partialReadCount_, std::array<int, 3>{{0, -200, 0}});

    // parsePartial: per-base error probabilities of partial reads (qualityFilter, Figbird.cpp:1780-1797, 5829-5836)
    // -- evaluated where finalize() adds them up (partialQualityAt): a table per read would be written for every gap and read
    //    for the few reads finalize accepts

    // ---- supported domain + flanks
    const std::string& ctg = sc_.seq[r.contigNo];
    const long clen = (long)ctg.size();
    int maxLen = 1;
    for (auto& u : in_.unm) maxLen = std::max<int>(maxLen, (int)u.seq.size());
    for (auto& p : in_.partial) maxLen = std::max<int>(maxLen, (int)p.seq.size());
    maxLen = std::max(maxLen, std::max(a_.readLength, a_.partialReadLen));
    const int F = maxLen;   // >= maxLen-1 rows on each side
    prep_.mode = a_.unmapped == 1 ? FB_MODE_UNMAPPED : FB_MODE_PARTIAL;
    prep_.origLen = og_; prep_.flankLen = F; prep_.gapStart = r.gapStart;
    // largest candidate ever evaluated: gapMax, the checkGapReads probes (<=3*og or 70) and og itself
    int gapMaxEver = std::max(og_, (int)((float)og_ * frac2_));
    if (a_.unmapped == 1 && og_ <= midLimitU_) gapMaxEver = std::max(gapMaxEver, og_ < 30 ? 70 : 3 * og_);
    // Supported domain = where the reference's own result is defined.  Its window is clipped at the scaffold ends
    // (initialize_start_end, Figbird.cpp:2268-2296): left_maxDistance = min(maxDistance, gapStart), right_maxDistance =
    // min(maxDistance, scaffold end - gapStart - Lg).  Pass 1 skips window rows < 0 (:3157,3578) but pass 2 indexes gapString
    // unguarded (:3386,3767,5036), and no loop bounds the row index above, so a read base outside the window reads memory the
    // worker never initialised.  With gapStart >= read length - 1 and (scaffold end - gapStart - Lg) >= read length - 1 for every
    // evaluated Lg no read base leaves the window and the clip changes nothing that is scored: the gap is filled exactly as the
    // reference fills it (goldens g8 / g9).  Outside that (a gap within a read length of a scaffold end) the reference's output
    // depends on heap contents; such gaps are emitted unfilled.  side_limit = min(30, left_maxDistance) stays 30 here.
    prep_.supported = r.contigNo >= 0 && r.contigNo < (int)sc_.seq.size() && r.gapStart >= F && a_.maxDistance >= F - 1 &&
                      r.gapStart + og_ + F <= clen && r.gapStart + gapMaxEver + F <= clen;
    if (!prep_.supported) { prep_.attempt = false; return; }
    prep_.flank.resize(2 * F);
    for (int i = 0; i < F; i++) prep_.flank[i] = code(ctg[r.gapStart - F + i]);
    for (int i = 0; i < F; i++) prep_.flank[F + i] = code(ctg[r.gapStart + og_ + i]);
    // findGapLeftRight (Figbird.cpp:2151-2174)
    gapLeft_ = ctg.substr(r.gapStart - sideLimit_, sideLimit_);
    gapRight_ = ctg.substr(r.gapStart + og_, sideLimit_);

    prep_.attempt = analyze();

    // ---- reads
    if (prep_.mode == FB_MODE_UNMAPPED) {
        for (int q = 0; q < numReads_; q++) {
            const UnmappedRead& u = in_.unm[q];
            prep_.readLen.push_back((int)u.seq.size());
            prep_.readMate.push_back((int32_t)(u.matePos - r.gapStart));
            prep_.readFlags.push_back((u.matePos < r.gapStart ? FB_READ_LEFT : 0) | (u.isReverse ? FB_READ_REVERSE : 0));
            prep_.readJlo.push_back(0); prep_.readJcut.push_back(0);
            std::vector<uint8_t> c(u.seq.size()); for (size_t k = 0; k < c.size(); k++) c[k] = code(u.seq[k]);
            prep_.readCodes.push_back(std::move(c));
        }
    } else {
        const int n = std::min(partialReadCount_, 3000);   // placeReads stops after partial_limit reads (Figbird.cpp:3122,3363)
        for (int q = 0; q < n; q++) {
            const PartialRead& p = in_.partial[q];
            prep_.readLen.push_back((int)p.seq.size());
            int fl = (p.pos < r.gapStart) ? FB_READ_LEFT : 0;
            if (p.refPos == -1) fl |= FB_READ_NOMATE;
            prep_.readFlags.push_back(fl);
            prep_.readMate.push_back(p.refPos == -1 ? 0 : (int32_t)(p.refPos - r.gapStart));
            bool m14 = (p.match == 1 || p.match == 4);
            prep_.readJlo.push_back(m14 ? kClipThresh : 0); prep_.readJcut.push_back(m14 ? 0 : kClipThresh);
            std::vector<uint8_t> c(p.seq.size()); for (size_t k = 0; k < c.size(); k++) c[k] = code(p.seq[k]);
            prep_.readCodes.push_back(std::move(c));
        }
    }
    // ---- partial-read pile-ups for the initial gap-row probabilities (update_partial_prob, Figbird.cpp:1941-2015)
    {
        int T = 1;
        const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
        for (int p = 0; p < np; p++) T = std::max<int>(T, (int)in_.partial[p].seq.size());
        prep_.pileLen = T;
        prep_.pileL.assign(4 * T, 0); prep_.pileR.assign(4 * T, 0);
        for (int p = 0; p < np; p++) {
            const PartialRead& pr = in_.partial[p];
            const int len = (int)pr.seq.size();
            int ci = pr.clippedIndex;
            if (p < (int)repeatflag_.size() && repeatflag_[p][0] != -1) {
                if (repeatflag_[p][0] == 1) ci = repeatflag_[p][2] + repeatflag_[p][1] - 1;
                if (repeatflag_[p][0] == 2) ci = repeatflag_[p][1];
            }
            if (pr.match == 1 || pr.match == 4) {
                for (int i = ci + 1, t = 0; i < len && t < T; i++, t++) {
                    if (i < 0) continue;
                    int cd = code(pr.seq[i]);
                    if (cd < 4) prep_.pileL[4 * t + cd] += 1; else for (int h = 0; h < 4; h++) prep_.pileL[4 * t + h] += 1;
                }
            } else if (pr.match == 2 || pr.match == 3) {
                for (int i = ci - 1, u = 0; i >= 0 && u < T; i--, u++) {
                    if (i >= len) continue;
                    int cd = code(pr.seq[i]);
                    if (cd < 4) prep_.pileR[4 * u + cd] += 1; else for (int h = 0; h < 4; h++) prep_.pileR[4 * u + h] += 1;
                }
            }
        }
    }
    // ---- cost estimate for sharding: candidates x rounds x reads x offsets
    {
        int gapMin = (int)((float)og_ * frac1_), gapMax = (int)((float)og_ * frac2_);
        double cand = prep_.attempt ? std::max(1, gapMax - gapMin + 1) : 0;
        double rounds = prep_.mode == FB_MODE_PARTIAL ? 3 : 12;
        double nr = (double)prep_.readLen.size();
        double offs = prep_.mode == FB_MODE_PARTIAL ? maxLen : (maxLen + 0.5 * (gapMin + gapMax));
        prep_.cost = cand * rounds * nr * offs * maxLen;
        prep_.sequential = prep_.attempt && prep_.mode == FB_MODE_UNMAPPED && largeGapFlag_ == 1;
    }
}

// 10^(-Q/10), Q = char - 33 (qualityFilter, Figbird.cpp:1780-1797): one pow() per distinct character instead of one per base
double GapFill::phredError(unsigned char ch) {
    static const std::vector<double> tab = [] { std::vector<double> t(256); for (int c = 0; c < 256; c++) { const int Q = (int)(char)c - 33; t[(size_t)c] = pow(10, -Q / 10.0); } return t; }();
    return tab[ch];
}

// findRepeat (Figbird.cpp:1799-1911)
int GapFill::findRepeat() {
    const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
    repeatflag_.assign(np, std::array<int, 3>{{-1, -1, -1}});
    int flag1 = 0, flag2 = 0, flag = 0;
    const int n = 20;
    for (int p = 0; p < np; p++) {
        const std::string& s5 = in_.partial[p].seq;
        std::vector<size_t> positions;
        int lim = (int)gapLeft_.size() - n;
        // every flank piece tried below contains the shortest one (the last suffix of the left flank, the last prefix of the right
        // flank), so a piece can occur twice in the read only if the shortest one does: one search settles almost every read
        auto twice = [&](const std::string& piece) { const size_t a = s5.find(piece); return a != std::string::npos && s5.find(piece, a + 1) != std::string::npos; };
        if (lim > 0 && !twice(gapLeft_.substr(lim - 1))) lim = 0;
        for (int i = 0; i < lim; i++) {
            std::string s3 = gapLeft_.substr(i);
            size_t pos = s5.find(s3, 0);
            while (pos != std::string::npos) { positions.push_back(pos); pos = s5.find(s3, pos + 1); }
            if (positions.size() > 1) {
                repeatflag_[p][0] = 1; repeatflag_[p][1] = (int)s3.length(); repeatflag_[p][2] = (int)positions[0];
                flag1 = 1; flag2 = p; oneSideRepeat_ = 1;
                break;
            }
            positions.clear();
        }
        positions.clear();
        lim = (int)gapRight_.size() - n;
        if (lim > 0 && !twice(gapRight_.substr(0, gapRight_.size() - (lim - 1)))) lim = 0;
        for (int i = 0; i < lim; i++) {
            std::string s4 = gapRight_.substr(0, gapRight_.size() - i);
            size_t pos = s5.find(s4, 0);
            while (pos != std::string::npos) { positions.push_back(pos); pos = s5.find(s4, pos + 1); }
            if (positions.size() > 1) {
                repeatflag_[p][0] = 2; repeatflag_[p][1] = (int)positions[positions.size() - 1];
                if (flag1 == 1 && p == flag2) flag = 1;
                oneSideRepeat_ = 1;
                break;
            }
            positions.clear();
        }
    }
    return flag;
}

// analyzeGap (Figbird.cpp:6168-6199): is the gap attempted at all?
bool GapFill::analyze() {
    repFlag_ = findRepeat();
    if (repFlag_ == 1 && a_.partialFlag) return false;
    if (oneSideRepeat_ == 1 && a_.partialFlag && og_ > 3 * midLimitP_) return false;
    if (fillflag_ == -1) return false;
    if (partialReadCount_ == 0 && numReads_ == 0) return false;
    return true;
}

// find_contig_match (Figbird.cpp:2176-2267): negative-overlap test on the two 30-mers, confirmed by a partial read
int GapFill::findContigMatch() const {
    if (og_ > negOverlap_) return 0;
    const int n = 3, errorThresh = 2;
    const std::string& s1 = gapLeft_; const std::string& s2 = gapRight_;
    for (int i = 0; i < sideLimit_ - n; i++) {
        std::string s3 = sub(s1, i, (long)s1.size()), s4 = sub(s2, 0, (long)s2.size() - i);
        if (s3.find(s4) == std::string::npos) continue;
        std::string rem = sub(s2, (long)s4.size(), (long)s2.size());
        int partCount = 0;
        for (const PartialRead& pr : in_.partial) {
            const std::string& rd = pr.seq;
            const int len = (int)rd.size();
            int maxMatch = -1, maxPos = -1;
            for (int j = 0; j < len - (int)s1.size(); j++) {
                int matchCount = 0, mismatch = 0;
                for (int k = 0; k < (int)s1.size(); k++) {
                    if (rd[j + k] == s1[k]) matchCount++; else mismatch++;
                    if (mismatch > errorThresh) break;
                }
                if (matchCount > maxMatch) { maxMatch = matchCount; maxPos = j; }
            }
            if ((int)s1.size() - maxMatch <= errorThresh) {
                int newpos = maxPos + (int)s1.size(), match = 0;
                for (int j = 0; j < (int)rem.size(); j++) {
                    char ch = (newpos + j) < len ? rd[newpos + j] : '\0';
                    if (rem[j] == ch) match++;
                }
                if ((int)rem.size() - match <= errorThresh) return (int)s4.size();
            }
            partCount++;
            if (partCount > 3000) break;
        }
    }
    return 0;
}

// The string side of update_partial_prob (Figbird.cpp:2036-2084): majority strings inside the pile-up edges.
void GapFill::pileUp(int Lg, bool makeStrings) {
    if (!makeStrings) return;
    // The votes of a partial read depend on the evaluated length only through a clip (left reads: rows counted from the
    // left edge, right reads: from the right edge), so they are piled up once per gap and every length reads the two
    // tables: cnt[i] = 1 + L[i] + R[Lg-1-i].
    if (!hostPileBuilt_) {
        hostPileBuilt_ = true;
        const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
        for (int p = 0; p < np; p++) {
            const PartialRead& pr = in_.partial[p];
            const int len = (int)pr.seq.size();
            int ci = pr.clippedIndex;
            if (repeatflag_[p][0] == 1) ci = repeatflag_[p][2] + repeatflag_[p][1] - 1;
            if (repeatflag_[p][0] == 2) ci = repeatflag_[p][1];
            if (pr.match == 1 || pr.match == 4) {
                const int ext = len - ci - 1;
                hostAnyLeft_ = true; hostMaxExt_ = std::max(hostMaxExt_, ext);
                if (ext > (int)hostPileL_.size()) hostPileL_.resize(ext, std::array<int, 4>{{0, 0, 0, 0}});
                for (int j = 0; j < ext; j++) {
                    const int i = ci + 1 + j;
                    const int cd = (i >= 0 && i < len) ? code(pr.seq[i]) : 4;
                    if (cd < 4) hostPileL_[j][cd] += 1; else for (int h = 0; h < 4; h++) hostPileL_[j][h] += 1;
                }
            } else if (pr.match == 2 || pr.match == 3) {
                hostAnyRight_ = true; hostMaxCi_ = std::max(hostMaxCi_, ci);
                if (ci > (int)hostPileR_.size()) hostPileR_.resize(ci, std::array<int, 4>{{0, 0, 0, 0}});
                for (int u = 0; u < ci; u++) {
                    const int i = ci - 1 - u;
                    const int cd = (i >= 0 && i < len) ? code(pr.seq[i]) : 4;
                    if (cd < 4) hostPileR_[u][cd] += 1; else for (int h = 0; h < 4; h++) hostPileR_[u][h] += 1;
                }
            }
        }
    }
    const int leftMax = hostAnyLeft_ ? std::max(std::min(hostMaxExt_, Lg), 0) - 1 : -kMaxGap;
    const int rightMin = hostAnyRight_ ? Lg - std::min(std::max(hostMaxCi_, 0), Lg) : kMaxGap;
    std::vector<std::array<int, 4>> cnt(Lg, std::array<int, 4>{{1, 1, 1, 1}});
    for (int i = 0; i < Lg; i++) {
        if (i < (int)hostPileL_.size()) for (int k = 0; k < 4; k++) cnt[i][k] += hostPileL_[i][k];
        const int u = Lg - 1 - i;
        if (u < (int)hostPileR_.size()) for (int k = 0; k < 4; k++) cnt[i][k] += hostPileR_[u][k];
    }
    int lc = 0, rc = 0;
    char* pl = pileStr_; char* pr = pileStr_ + 100;
    for (int i = 0; i < Lg; i++) {
        int maxIndex = 0, maxVal = -1;
        for (int k = 0; k < 4; k++) if (cnt[i][k] > maxVal) { maxVal = (int)cnt[i][k]; maxIndex = k; }
        if (i <= leftMax - 5 || i >= rightMin + 5) {
            char ch = maxIndex == 0 ? 'A' : maxIndex == 1 ? 'C' : maxIndex == 2 ? 'G' : 'T';
            if (i <= leftMax - 5) { if (lc < kOvlBytes) pl[lc] = ch; lc++; }
            else { if (100 + rc < kOvlBytes) pr[rc] = ch; rc++; }
        }
    }
    if (lc < kOvlBytes) pl[lc] = '\0';
    if (100 + rc < kOvlBytes) pr[rc] = '\0';
}

// initialize(gapEstimate) runs update_partial_prob for every evaluated length (Figbird.cpp:2111-2115).  The gap rows it
// produces are computed on the device; on the host only its overflow matters: partial_right holds min(longest right-side
// pile-up, Lg) - 5 characters, and from 100 on they land on the saved-read indices and side_limit (see ovl_).  Reads of up
// to ~105 bases can never get there, so the strings are only rebuilt for the (gap, length) pairs that can.
void GapFill::initializeEffects(int Lg) {
    if (maxRightCi_ < 0) {
        maxRightCi_ = 0;
        const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
        for (int p = 0; p < np; p++) {
            const PartialRead& pr = in_.partial[p];
            if (!(pr.match == 2 || pr.match == 3)) continue;
            int ci = pr.clippedIndex;
            if (repeatflag_[p][0] == 1) ci = repeatflag_[p][2] + repeatflag_[p][1] - 1;
            if (repeatflag_[p][0] == 2) ci = repeatflag_[p][1];
            maxRightCi_ = std::max(maxRightCi_, ci);
        }
    }
    if (std::min(maxRightCi_, Lg) - 5 >= 100) pileUp(Lg, true);
}

void GapFill::setConcensus(const std::vector<uint8_t>& codes, int len) { setConcensus(codes.data(), len); }
void GapFill::setConcensus(const uint8_t* codes, int len) {
    if ((int)concensus_.size() < len + 1) concensus_.resize(len + 1, 0);
    for (int i = 0; i < len; i++) concensus_[i] = letter(codes[i]);
    concensus_[len] = '\0';
}

void GapFill::copyStr(std::vector<char>& dst, const std::vector<char>& src) {   // strcpy
    size_t n = cstrlen(src);
    if (dst.size() < n + 1) dst.resize(n + 1, 0);
    memcpy(dst.data(), src.data(), n);
    dst[n] = '\0';
}

ItemSpec GapFill::emSpec(int Lg) const {
    ItemSpec s; s.kind = FB_ITEM_EM; s.candLen = Lg;
    if (prep_.mode == FB_MODE_PARTIAL) { s.maxRounds = 3; s.flags = FB_FLAG_RECORD_ALL | FB_FLAG_NO_COMP_STOP; }
    else { s.maxRounds = kNumItr; s.flags = 0; }
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// Epilogues: the part of placeReads that follows the scored loops.
// ---------------------------------------------------------------------------------------------------------

// findOverlapUnmapped (Figbird.cpp:2945-3019)
double GapFill::findOverlapUnmapped(int Lg) {
    int incorrectPenalty = 0, gapPenalty = 0;
    std::vector<Pos3> vec(finalReadpos_.begin(), finalReadpos_.end());
    std::sort(vec.begin(), vec.end(), [](const Pos3& x, const Pos3& y) { return x.a < y.a; });
    const int n = numReads_;
    for (int i = 0; i < n - 1; i++) {
        if (vec[i].a == -200) continue;
        int diff = vec[i].a + vec[i].b - vec[i + 1].a;
        if (diff >= kMatchDiscont) {}
        else if (diff >= 0) { incorrectPenalty += -250; discont_ = 1; }
        else {
            gapPenalty += -4 * 50;
            if (vec[i].c >= 0) markAccepted_[vec[i].c] = 0;
            if (vec[i + 1].c >= 0) markAccepted_[vec[i + 1].c] = 0;
            if (Lg == og_) {
                if (vec[i].c >= 0) { unmPosOrg_[vec[i].c].a = -200; unmPosOrg_[vec[i].c].b = 0; }
                if (vec[i + 1].c >= 0) { unmPosOrg_[vec[i + 1].c].a = -200; unmPosOrg_[vec[i + 1].c].b = 0; }
            }
            validCount_ -= 2;
        }
    }
    int lr = 0;
    for (int i = 0; i < n; i++) {
        if (vec[i].a == -200) continue;
        if (vec[i].a < 0 && -vec[i].a >= 3 && vec[i].a + vec[i].b > 0) lr++;
        if (vec[i].a < Lg && vec[i].a + vec[i].b - Lg >= 3) lr++;
    }
    return (double)(0 + incorrectPenalty + gapPenalty + lr * 50);
}

// The "update countsGap from rejected reads that match the N-region borders" step (Figbird.cpp:4032-4376).
void GapFill::borderUpdate(int Lg) {
    const int RL = unmReadLen_;
    std::vector<int> indexPair(1000 + 4, -1);
    int startN = 0, pairCount = 0, Ncount = 0, numMatch0 = 0;
    const int matchThreshold = (int)(RL * 0.25), maxSeg = (int)(RL * 0.67), minGapLen = RL / 2 + 1;
    std::vector<std::array<double, 4>> countPos(Lg, std::array<double, 4>{{0, 0, 0, 0}});
    const char* con = concensus_.data();
    const int clen = (int)cstrlen(concensus_);
    for (int i = 0; i < clen; i++) {
        if (con[i] == 'N' && startN == 0) { startN = 1; if (pairCount < 1000) indexPair[pairCount] = i > 0 ? i - 1 : i; pairCount++; Ncount++; }
        else if (con[i] != 'N' && startN == 1) { startN = 0; if (pairCount < 1000) indexPair[pairCount] = i; pairCount++; if (Ncount < minGapLen) pairCount -= 2; Ncount = 0; }
        else if (con[i] == 'N' && startN == 1) Ncount++;
        if (i == clen - 1 && startN == 1) { if (pairCount < 1000) indexPair[pairCount] = i; pairCount++; if (Ncount < minGapLen) pairCount -= 2; }
    }
    if (pairCount > 2) indexPair[1] = indexPair[pairCount - 1];
    int flag1 = 1, flag2 = 1;
    if (pairCount < 2) { flag1 = 0; flag2 = 0; }
    if (!(flag1 == 1 || flag2 == 1)) return;
    const int endIndex1 = indexPair[0], startIndex1 = indexPair[1];
    int indexS = endIndex1 >= maxSeg ? endIndex1 - maxSeg + 1 : 0;
    std::string textLeft, textRight;
    for (int j = indexS; j < clen && con[j] != 'N'; j++) textLeft.push_back(con[j]);
    int stopIndex = (startIndex1 + maxSeg <= clen) ? startIndex1 + maxSeg - 1 : clen - 1;
    for (int j = startIndex1; j <= stopIndex; j++) textRight.push_back(con[j]);
    const int tf1 = flag1, tf2 = flag2;
    const long gapStart = in_.rec.gapStart;
    for (int q = 0; q < numReads_; q++) {
        if (!(markAccepted_[q] == 0 && mlvNonZero_[q])) continue;
        const std::string& rd = in_.unm[q].seq;
        const int R2 = (int)rd.size();
        const int mapped = in_.unm[q].matePos;
        flag1 = tf1; flag2 = tf2;
        for (int g = 0; g < 2; g++) {
            long placed = gapStart + indexPair[g];
            long isz = mapped < gapStart ? placed + R2 - mapped : mapped - placed + R2;
            if (isz < (m_.insertThresholdMin) + 100 || isz > (m_.insertThresholdMax - 100)) { if (g == 0) flag1 = 0; else flag2 = 0; }
        }
        if (flag1 == 1) {
            const int tl = (int)textLeft.size();
            for (int j = 0; j < tl; j++) {
                int match = 0, mf = 1;
                for (int k = 0; k < tl - j; k++) {
                    char rc = k < R2 ? rd[k] : '\0';
                    if (textLeft[j + k] != rc) { mf = 0; break; } else match++;
                }
                if (mf == 1 && match > matchThreshold) {
                    numMatch0++;
                    for (int r = 0; r < R2; r++) {
                        int indRead = code(rd[r]);
                        int indRef = indexPair[0] - match + 1 + r;
                        if (indRef == Lg) break;
                        if (indRef > indexPair[0] && indRef >= 0 && indRef < Lg) {
                            if (indRead < 4) countPos[indRef][indRead] += match; else for (int z = 0; z < 4; z++) countPos[indRef][z] += match;
                        }
                    }
                    break;
                }
            }
        }
        if (flag2 == 1) {
            std::string rr(rd.rbegin(), rd.rend()), tr(textRight.rbegin(), textRight.rend());
            const int tl = (int)tr.size();
            for (int j = 0; j < tl; j++) {
                int match = 0, mf = 1;
                for (int k = 0; k < tl - j; k++) {
                    char rc = k < R2 ? rr[k] : '\0';
                    if (tr[j + k] != rc) { mf = 0; break; } else match++;
                }
                if (mf == 1 && match > matchThreshold) {
                    for (int r = 0; r < R2; r++) {
                        int indRead = code(rd[R2 - r - 1]);
                        int indRef = indexPair[1] + match - 1 - r;
                        if (indRef < 0) break;
                        if (indRef < indexPair[1] && indRef < Lg) {
                            if (indRead < 4) countPos[indRef][indRead] += match; else for (int z = 0; z < 4; z++) countPos[indRef][z] += match;
                        }
                    }
                    break;
                }
            }
        }
    }
    // fall back to the partial pile-up string on the left (the right-hand twin can never fire: its
    // match counter starts at 1, Figbird.cpp:4040,4321)
    const int plen = (int)strnlen(pileStr_, kOvlBytes);      // strlen(partial_left): may run on into what follows it
    if (flag1 == 1 && numMatch0 == 0 && indexPair[0] < plen)
        for (int f = indexPair[0] + 1; f < plen; f++) { int cd = code(pileStr_[f]); if (f >= 0 && f < Lg && cd < 4) countPos[f][cd] += 1; }
    for (int j = 0; j < Lg; j++) {
        int tot = 0;
        for (int k = 0; k < 4; k++) tot += countPos[j][k];
        if (tot > 0) for (int k = 0; k < 4; k++) counts_[j][k] = (countPos[j][k] / tot);
    }
}

// Everything placeReads does in unmapped mode after the scored loops (Figbird.cpp:3848-4377), for one call.
double GapFill::unmappedEpilogue(const ItemResult& r, int slot, int Lg, int finalizeFlag, int updateFlag, int ge) {
    const int R = numReads_;
    double like = 0;
    const double* p1 = r.p1max + (size_t)slot * R;
    const double* p2 = r.p2max + (size_t)slot * R;
    const int32_t* ps = r.pos2 + (size_t)slot * R;
    for (int q = 0; q < R; q++) {
        markAccepted_[q] = 0; finalReadpos_[q] = Pos3{-200, 0, -1};
        if (Lg == og_) unmPosOrg_[q] = Pos3{-200, 0, 0};
        mlvNonZero_[q] = (p1[q] > 0) ? (p1[q] != 1.0) : 0;      // log10(p) != 0: only p == 1 has a zero logarithm (|log10(1 +- ulp)| > 4e-17)
    }
    for (int q = 0; q < R; q++) {
        const double maxProb = p2[q] >= 0 ? p2[q] : -DBL_MAX;
        const double tlv = -log10(maxProb);
        if (tlv < m_.gapProbCutOff) {
            mlvNonZero_[q] = (-tlv != 0);
            like += -tlv;
            validCount_++;
            const int len = prep_.readLen[q];
            markAccepted_[q] = 1;
            finalReadpos_[q] = Pos3{ps[q], len, q};
            if (Lg == og_) unmPosOrg_[q] = Pos3{ps[q], len, q};
        } else like += -50;
    }
    umaxFlags_ |= r.flags;
    // computeSequence(1,1): hard consensus + coverage come back from the device
    setConcensus(r.hard, Lg);
    gapCoverage_.assign(std::max<size_t>(gapCoverage_.size(), (size_t)Lg), 0);
    for (int i = 0; i < Lg; i++) gapCoverage_[i] = r.cov[i];
    compCount_ = r.compCount;
    if (!finalizeFlag) return like;

    // low-coverage regions (Figbird.cpp:3935-3977)
    std::vector<int> region(Lg + 4, 0);
    int regionStart = 0, regionCount = 0;
    for (int i = 0; i < Lg; i++) {
        if (gapCoverage_[i] < kCov2 && regionStart == 0) { region[regionCount] = i; regionStart = 1; }
        else if (gapCoverage_[i] >= kCov2 && regionStart == 1) {
            if (i - 1 - region[regionCount] >= 10) { regionStart = 0; region[regionCount + 1] = i - 1; regionCount += 2; }
        }
        if (i == Lg - 1 && regionStart == 1) {
            if (i - region[regionCount] >= 10) { regionStart = 0; region[regionCount + 1] = i; regionCount += 2; }
        }
    }
    int rl = kMaxGap, rr = -kMaxGap;
    if (regionCount) { rl = region[0]; rr = region[regionCount - 1]; regionPerct_ = (rr - rl * 1.0) / Lg; }
    else regionPerct_ = 0;
    for (int q = 0; q < R; q++) {
        if (markAccepted_[q] != 1) continue;
        if (finalReadpos_[q].a >= rl && finalReadpos_[q].a + finalReadpos_[q].b - 1 < rr) {
            like += -50;
            markAccepted_[q] = 0;
            if (Lg == og_) { unmPosOrg_[q].a = -200; unmPosOrg_[q].b = 0; }
            validCount_--;
            finalReadpos_[q].a = -200; finalReadpos_[q].b = 0;
        }
    }
    like += findOverlapUnmapped(Lg);
    const int condition = compCount_ >= 1 && regionPerct_ != 0 && ge != kNumItr - 1;
    if (condition && updateFlag) borderUpdate(Lg);
    return like;
}

// detect_overlap_gapestimate (Figbird.cpp:2513-2779).  pflag rows: {used, position}.
void GapFill::detectOverlap(const std::vector<std::array<int, 2>>& pflag, const std::vector<std::array<int, 3>>*, int gaplen, int* ret, int lenThresh, bool* touchedSaved) {
    int lMax = -kMaxGap, rMin = kMaxGap;
    std::vector<int>& leftCross = scratchLeft_; std::vector<int>& rightCross = scratchRight_;
    leftCross.clear(); rightCross.clear();
    const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
    const int prc = partialReadCount_;
    std::vector<int>& smFlag = scratchSm_; smFlag.assign(prc, 0);
    if (touchedSaved) *touchedSaved = false;
    const double mismatchThreshold = .1;
    for (int p = 0; p < np && p < prc; p++) {
        if (pflag[p][0] == 0) continue;
        const int match = in_.partial[p].match, len = (int)in_.partial[p].seq.size();
        const int pos = pflag[p][1];
        int j, flag = 0, start = -1;
        for (j = 0; j < len; j++) {
            if (pos + j >= 0 && pos + j < gaplen) { if (flag == 0) { flag = 1; start = pos + j; } }
            if (pos + j == gaplen) break;
        }
        if (match == 1 || match == 4 || pos < 0) { if (pos + j - 1 > lMax) lMax = pos + j - 1; }
        else { if (start < rMin) rMin = start; }
    }
    if (lMax == -kMaxGap) lMax = -1;
    if (rMin == kMaxGap) rMin = gaplen;
    int ovflag = 0;
    for (int k = 0; k < prc; k++) {
        if (pflag[k][0] != 1) continue;
        const std::string& s1 = in_.partial[k].seq;
        const int len = (int)s1.size(), placed = pflag[k][1], match = in_.partial[k].match;
        if (placed < 0 && placed + len > gaplen) {
            std::string l = (-placed < sideLimit_) ? sub(s1, 0, -placed) : sub(s1, -placed - sideLimit_, sideLimit_);
            std::string rr = sub(s1, -placed + gaplen, sideLimit_);
            if (partialMatch(gapLeft_, l, 0, 0, lenThresh) && partialMatch(gapRight_, rr, 1, 0, lenThresh)) ovflag = 1;
        }
        if (placed < 0 && placed + len - 1 >= rMin && placed + len <= gaplen) leftCross.push_back(k);
        if (placed > 0 && placed <= lMax) rightCross.push_back(k);
        if (placed < 0 && placed + len > gaplen && (match == 2 || match == 3)) { rightCross.push_back(k); smFlag[k] = 1; }
    }
    if (ovflag || (in_.rec.stat2 == 1 && og_ <= 20 && gaplen == in_.rec.stat3)) { ret[0] = 300; ret[1] = 0; return; }
    auto readFrag = [&](const std::string& s1, int placed) -> std::string {   // get_read_frag, Figbird.cpp:2492-2511
        int neg = -placed;
        if (placed < 0) return neg < sideLimit_ ? sub(s1, 0, neg) : sub(s1, neg - sideLimit_, sideLimit_);
        return sub(s1, gaplen - placed, sideLimit_);
    };
    if (rMin <= lMax) {
        if (touchedSaved) *touchedSaved = true;      // from here on partial_saved_read_temp is always written (set by a pair, or reset below)
        int maxOverlap = 0, falseFlag = 0;
        for (size_t i = 0; i < leftCross.size(); i++) for (size_t j = 0; j < rightCross.size(); j++) {
            const int li = leftCross[i], ri = rightCross[j];
            if (li == ri) continue;
            const std::string& r1 = in_.partial[li].seq; const std::string& r2 = in_.partial[ri].seq;
            const int p1 = pflag[li][1], p2 = pflag[ri][1], len = (int)r1.size();
            int diffGap = p1 + len - gaplen, overlap;
            if (diffGap > 0) overlap = (p1 + len - 1) - p2 + 1 - diffGap;
            else { overlap = (p1 + len - 1) - p2 + 1; diffGap = 0; }
            if (overlap < overlapThreshold_) continue;
            std::string cl, cr;
            if (smFlag[ri] != 1) {
                if (partialMatch(gapLeft_, readFrag(r1, p1), 0, 1, -1)) cl = sub(r1, (long)r1.size() - overlap - diffGap, overlap);
                if (partialMatch(gapRight_, readFrag(r2, p2), 1, 1, -1)) cr = sub(r2, 0, overlap);
            } else {
                const int x = pflag[ri][1];
                if (partialMatch(gapLeft_, readFrag(r1, p1), 0, 1, -1)) cl = sub(r1, (long)r1.size() - overlap - x, overlap - x);
                std::string s3 = sub(r2, -x + gaplen, sideLimit_);
                if (partialMatch(gapRight_, s3, 1, 1, -1)) cr = sub(r2, -x, overlap + x);
            }
            const int len1 = (int)cl.size(), len2 = (int)cr.size();
            if (len1 > 0 && len2 > 0 && len1 == len2) {
                double mf = diffFrac(cl, cr, len1);
                if (mf <= mismatchThreshold) { if (len1 > maxOverlap) { maxOverlap = len1; savedTemp_[0] = li; savedTemp_[1] = ri; } }
                else falseFlag = -1;
            }
        }
        if ((falseFlag == 0 && maxOverlap >= overlapThreshold_) || (falseFlag == -1 && maxOverlap >= 2 * overlapThreshold_)) { ret[0] = maxOverlap; ret[1] = 0; }
        else if (falseFlag == -1 || maxOverlap < overlapThreshold_) { ret[0] = 0; ret[1] = -1; savedTemp_[0] = savedTemp_[1] = -1; }
        return;
    }
    ret[0] = 0; ret[1] = 0;
}

// partial-mode placeReads after the scored loops (Figbird.cpp:3258-3523), for one call.
// savedOnly: apply nothing but the call's effect on partial_saved_read_temp (see evalCandidate).
double GapFill::partialEpilogue(const ItemResult& r, int slot, int Lg, bool savedOnly, bool* touchedSaved) {
    const int R = (int)prep_.readLen.size();
    const double* p1 = r.p1max + (size_t)slot * R;
    const double* p2 = r.p2max + (size_t)slot * R;
    const int32_t* ps = r.pos2 + (size_t)slot * R;
    double like = 0;
    if (!savedOnly) for (int q = 0; q < R; q++) if (p1[q] > 0) like += log(p1[q]);
    std::vector<std::array<int, 2>>& pflag = scratchPflag_;
    pflag.assign(partialReadCount_, std::array<int, 2>{{1, 0}});
    if (Lg == og_ && !savedOnly) for (auto& o : partialPosOrg_) o = std::array<int, 3>{{0, -200, 0}};
    for (int q = 0; q < R; q++) {
        const double maxProb = p2[q] >= 0 ? p2[q] : -DBL_MAX;
        const double tlv = -log10(maxProb);
        if (tlv < m_.gapProbCutOff) {
            if (!savedOnly) validCount_++;
            pflag[q][1] = ps[q];
            if (Lg == og_ && !savedOnly) partialPosOrg_[q] = std::array<int, 3>{{1, ps[q], prep_.readLen[q]}};
        } else pflag[q][0] = 0;
    }
    int ret[2] = {0, 0};
    detectOverlap(pflag, nullptr, Lg, ret, 8, touchedSaved);
    if (ret[0] == 300) like += ret[0];
    else if (ret[0] >= 1 && ret[0] < 200) like += 30 * ret[0];
    else if (ret[1] == -1) like += -100;
    return like;
}

// Apply the epilogue(s) of one evaluated length; returns the likelihood of its last placeReads call and
// leaves concensus_ = computeSequence(0,0) of that call, validCount_ etc. as the reference has them.
double GapFill::evalCandidate(const ItemResult& r, int Lg, int finalizeFlag) {
    double like = 0;
    refPlacements_ += r.placements;
    initializeEffects(Lg);
    if (prep_.mode == FB_MODE_PARTIAL) {
        // The reference runs this epilogue after each of the three placeReads calls; what survives is the last call's likelihood,
        // valid_count and original-length positions, and partial_saved_read_temp as the last call that *wrote* it left it
        // (detect_overlap_gapestimate writes it exactly when it reaches its overlap block, Figbird.cpp:2671-2775).  So the last call
        // is evaluated in full and earlier ones only while the saved pair is still undetermined.
        bool touched = false;
        validCount_ = 0;
        like = partialEpilogue(r, r.calls - 1, Lg, false, &touched);
        for (int s = r.calls - 2; s >= 0 && !touched; s--) partialEpilogue(r, s, Lg, true, &touched);
    } else {
        validCount_ = 0;
        like = unmappedEpilogue(r, 0, Lg, finalizeFlag, 0, r.calls - 1);
    }
    return like;
}

// Host-driven rounds for gaps > 400 bp in unmapped mode: the border update may rewrite countsGap between
// rounds, and the stop rule looks at region_perct (Figbird.cpp:6323-6344 with large_gap_flag).
double GapFill::largeGapRounds(int Lg, int finalizeFlag, int updateFlag, bool) {
    double like = 0;
    compCount_ = 0;
    pileUp(Lg, true);      // initialize -> update_partial_prob also derives partial_left/right (Figbird.cpp:2053-2084)
    std::vector<uint8_t> prevHard; bool havePrev = false;
    const int presetUnfilled = 2 * unmReadLen_;
    for (int i = 0; i < kNumItr; i++) {
        ItemSpec s; s.kind = FB_ITEM_EM; s.candLen = Lg; s.maxRounds = 1; s.flags = FB_FLAG_WANT_COUNTS;
        if (i > 0) {
            s.flags |= FB_FLAG_RESUME; s.compIn = compCount_;
            s.countsIn.resize((size_t)Lg * 5);
            for (int x = 0; x < Lg; x++) for (int k = 0; k < 5; k++) s.countsIn[(size_t)x * 5 + k] = counts_[x][k];
            if (havePrev) s.stringIn = prevHard;
        }
        std::vector<ItemResult> res;
        dev_->submit(bidx_, {s}, res);
        const ItemResult& r = res[0];
        refPlacements_ += r.placements;
        if ((int)counts_.size() < Lg) counts_.resize(Lg, std::array<double, 5>{{0, 0, 0, 0, 0}});
        for (int x = 0; x < Lg; x++) for (int k = 0; k < 5; k++) counts_[x][k] = r.counts[(size_t)x * 5 + k];
        // comp_count bookkeeping mirrors the device: prev string changes only when the consensus changed
        if (!(havePrev && (int)prevHard.size() == Lg && std::equal(prevHard.begin(), prevHard.end(), r.hard))) { prevHard.assign(r.hard, r.hard + Lg); havePrev = true; }
        validCount_ = 0;
        like = unmappedEpilogue(r, 0, Lg, finalizeFlag, updateFlag, i);
        // computeSequence(0,0) after the loop sees countsGap as the border update left it
        lastSoft_.assign(Lg, 4);
        for (int x = 0; x < Lg; x++) { double mx = 0; int mi = -1; for (int k = 0; k <= 4; k++) if (counts_[x][k] > mx) { mx = counts_[x][k]; mi = k; } lastSoft_[x] = (mi >= 0 && mi < 4) ? mi : 4; }
        if (compCount_ >= 5) break;
        if (largeGapFlag_ == 1 && updateFlag && regionPerct_ * Lg < presetUnfilled) break;
    }
    return like;
}

// GapFiller::run (Figbird.cpp:5913-5965)
double GapFill::runLength(int Lg, int finalizeFlag, int c) {
    double like;
    regionPerct_ = 0;
    ItemSpec s = emSpec(Lg);
    std::vector<ItemResult> res;
    dev_->submit(bidx_, {s}, res);
    like = evalCandidate(res[0], Lg, finalizeFlag);
    lastSoft_.assign(res[0].soft, res[0].soft + Lg);
    regionPerctMax_ = regionPerct_;
    return c == 0 ? (double)validCount_ : like;
}

// checkGapReads (Figbird.cpp:6121-6153): probe a few lengths; >= 3 placed reads anywhere => worth scanning
int GapFill::checkGapReads() {
    std::vector<int> probes;
    const int thresh = 3;
    if (og_ < 30) { int step = og_ < 15 ? 10 : 20; for (int i = 0; i < 80; i += step) probes.push_back(i); }
    else for (int k = 0; k < 4; k++) probes.push_back(k == 0 ? og_ / 2 : og_ * k);
    std::vector<ItemSpec> specs; for (int L : probes) specs.push_back(emSpec(L));
    std::vector<ItemResult> res;
    dev_->submit(bidx_, specs, res);        // speculative: all probes at once, replayed in order below
    for (size_t i = 0; i < probes.size(); i++) {
        regionPerct_ = 0;
        evalCandidate(res[i], probes[i], 1);
        lastSoft_.assign(res[i].soft, res[i].soft + probes[i]);
        // previous_str is a member that run() never resets (Figbird.cpp:3919-3927, 6254): the last probe's final hard consensus
        // is what the first placeReads call of the next evaluated length is compared with
        prevStrLen_ = probes[i]; prevStr_.assign(res[i].hard, res[i].hard + probes[i]);
        regionPerctMax_ = regionPerct_;
        if (og_ < 30) { if (validCount_ > thresh) return -1; }
        else { if (validCount_ >= thresh) return -1; }
    }
    return 1;
}

// computeSequence(1,0) on the host copy of countsGap (Figbird.cpp:4417-4508), over this->gapLength rows
void GapFill::computeSequenceHost(int check) {
    const int Lg = gapLength_;
    if ((int)concensus_.size() < Lg + 1) concensus_.resize(Lg + 1, 0);
    if ((int)gapCoverage_.size() < Lg) gapCoverage_.resize(Lg, 0);
    for (int i = 0; i < Lg; i++) {
        double mx = 0; int mi = -1;
        const std::array<double, 5> zero{{0, 0, 0, 0, 0}};
        const std::array<double, 5>& row = i < (int)counts_.size() ? counts_[i] : zero;
        for (int j = 0; j <= 4; j++) if (row[j] > mx) { mx = row[j]; mi = j; }
        int coverageFlag = 1;
        if (check == 1) { gapCoverage_[i] = (int)mx; if (gapCoverage_[i] <= kCov1) coverageFlag = 0; }
        concensus_[i] = coverageFlag ? letter(mi < 0 ? 4 : mi) : 'N';
    }
    concensus_[Lg] = '\0';
}

int GapFill::findRegion(std::vector<int>& region) const {   // Figbird.cpp:4594-4621
    int Nstart = 0, rc = 0;
    const int len = gapLength_;
    region.assign(2 * (len + 2), 0);
    for (int i = 0; i < len; i++) {
        if (concensus_[i] == 'N' && Nstart == 0) { region[2 * rc] = i; Nstart = 1; }
        else if (concensus_[i] != 'N' && Nstart == 1) { Nstart = 0; region[2 * rc + 1] = i - 1; rc++; }
        if (i == len - 1 && Nstart == 1) { region[2 * rc + 1] = i; rc++; }
    }
    return rc;
}

// recheck_sequence + findDiscontinous (Figbird.cpp:4623-4743)
int GapFill::recheckSequence(const std::vector<Pos3>& pos) {
    std::vector<int> region;
    int regionCount = findRegion(region);
    const int len = gapLength_;
    std::vector<int> nec;
    {
        std::vector<Pos3> vec(pos.begin(), pos.begin() + numReads_);
        std::sort(vec.begin(), vec.end(), [](const Pos3& x, const Pos3& y) { return x.a < y.a; });
        for (int i = 0; i < numReads_ - 1; i++) {
            if (vec[i].a == -200) continue;
            int diff = vec[i].a + vec[i].b - vec[i + 1].a;
            if (diff >= 0 && diff <= kMatchDiscont / 2) nec.push_back(vec[i].a + vec[i].b);
        }
    }
    const int flag = (int)nec.size();
    if (flag > 0) {
        for (int k : nec) if (k >= 0 && k < (int)concensus_.size()) concensus_[k] = 'N';
        regionCount = findRegion(region);
    }
    const double reduction = og_ < 400 ? 1 : og_ < 1200 ? 1.5 : 2;
    const int readchar = 30;
    if (regionCount <= 1) {
        if (regionCount == 1) {
            if (regionPerctMax_ < .75 || flag > 0) {
                int i, j;
                for (i = region[0] - 1; i >= region[0] - reduction * readchar && i >= 0; i--) concensus_[i] = 'N';
                for (j = region[1] + 1; j <= region[1] + reduction * readchar && j < len; j++) concensus_[j] = 'N';
                if (i < 0 && j == len) return 1;
            }
        }
    } else {
        const int start = region[0], end = region[2 * regionCount - 1];
        for (int j = start; j < end; j++) concensus_[j] = 'N';
        int i, j;
        for (i = start - 1; i > start - 1 - reduction * readchar && i >= 0; i--) concensus_[i] = 'N';
        for (j = end + 1; j < end + 1 + reduction * readchar && j < len; j++) concensus_[j] = 'N';
        if (i < 0 && j == len) { gapLength_ = og_; return 1; }
    }
    return 0;
}

// check_update (Figbird.cpp:4535-4581)
int GapFill::checkUpdate(const std::array<double, 5>& arr, int j) const {
    double maxVal = -DBL_MAX, second = -DBL_MAX;
    int maxp = -kMaxGap, secp = -kMaxGap;
    for (int k = 0; k < 4; k++) {
        if (arr[k] > maxVal) { second = maxVal; secp = maxp; maxVal = arr[k]; maxp = k; }
        else if (arr[k] >= second) { second = arr[k]; secp = k; }
    }
    const int diff = (int)(maxVal - second);
    auto q = [&](int p) { return (p >= 0 && p < 5 && j < (int)qualGap_.size()) ? qualGap_[j][p] : 0.0; };
    if (diff >= kPartialThreshold) {
        if (maxVal > 3 && second > 3) return q(maxp) <= q(secp) ? maxp : secp;
        return 50;
    }
    if (maxVal >= 1 && second >= 1) return q(maxp) <= q(secp) ? maxp : secp;
    return -1;
}

// draw_read (Figbird.cpp:2385-2427)
void GapFill::drawHeader(int length) {
    draw_.append(unmReadLen_, ' ');
    draw_ += "====================+Gap = " + std::to_string(in_.rec.gapNo) + " starting,length = " + std::to_string(length) + "===============================\n";
    draw_.append(unmReadLen_, ' ');
    draw_.append(std::max(length, 0), 'N');
    draw_ += "\n";
}
void GapFill::drawRead(int length, const std::string& s, int readno, int isz, char type) {
    if (unmReadLen_ + length > 0) draw_.append(unmReadLen_ + length, ' ');
    draw_ += s + "[" + std::to_string(readno) + " " + std::to_string(length) + " isz = " + std::to_string(isz) + " " + std::string(1, type) + "]\n";
}

// ---------------------------------------------------------------------------------------------------------
// finalize (Figbird.cpp:4929-5659)
// ---------------------------------------------------------------------------------------------------------
void GapFill::finalize(int gl) {
    const long gapStart = in_.rec.gapStart;
    const int gapoffset = gl - og_;
    const int R = (int)prep_.readLen.size();
    std::vector<std::array<int, 3>> partialReadFlag(partialReadCount_, std::array<int, 3>{{0, -200, partialReadLen_}});
    std::vector<Pos3> unmPos(numReads_, Pos3{-200, 0, -1});
    int leftRightCheck[2] = {0, 0};
    int leftcount = 0, rightcount = 0, leftStartZero = 0, rightFinGlen = 0;
    int totalCount = 0, discardedCount = 0;
    int unmMaxLeft = 0, unmMaxRight = 0;

    // hard placement of every read on bestString: one HARD item on the device
    ItemSpec hs; hs.kind = FB_ITEM_HARD; hs.candLen = gl; hs.flags = FB_FLAG_FINALIZE_REF;
    hs.stringIn.resize(gl);
    for (int i = 0; i < gl; i++) hs.stringIn[i] = code(i < (int)bestString_.size() ? bestString_[i] : '\0');
    std::vector<ItemResult> res;
    dev_->submit(bidx_, {hs}, res);
    const ItemResult& hr = res[0];

    const int endLim = allocArg_;
    if ((int)counts_.size() < std::max(endLim, gl)) counts_.resize(std::max(endLim, gl), std::array<double, 5>{{0, 0, 0, 0, 0}});
    for (int i = 0; i < endLim; i++) counts_[i] = std::array<double, 5>{{0, 0, 0, 0, 0}};
    gapLength_ = gl;
    const int W = (int)std::min<long>(a_.maxDistance, gapStart);      // left_maxDistance (Figbird.cpp:2272-2277): an unplaced read keeps maxPos = 0

    if (a_.unmapped) {
        drawHeader(gapLength_);
        for (int q = 0; q < numReads_; q++) {
            totalCount++;
            double maxProb = 0; int posRel = -W;
            if (hr.p2max[q] > 0) { maxProb = hr.p2max[q]; posRel = hr.pos2[q]; }
            const UnmappedRead& u = in_.unm[q];
            const int len = prep_.readLen[q];
            long pos1 = u.matePos;
            int mle;
            if (pos1 < gapStart) mle = (int)(posRel + gapStart - pos1 + len);
            else { pos1 += gapoffset; mle = (int)(pos1 + len - (posRel + gapStart)); }
            if (-log10(maxProb) < m_.gapProbCutOff && savedReads_[q] == 1) {
                drawRead(posRel, u.seq, q, mle, pos1 < gapStart ? 'I' : 'E');
                unmPos[q] = Pos3{posRel, len, q};
                if (posRel == 0) leftStartZero = 1;
                if (posRel + len == gapLength_) rightFinGlen = 1;
                if (posRel < 0 && posRel + len > 0) { leftRightCheck[0] = 1; if (-posRel > unmMaxLeft) unmMaxLeft = -posRel; }
                const int val = posRel + len - gapLength_;
                if (posRel < gapLength_ && val > 0) { leftRightCheck[1] = 1; if (val > unmMaxRight) unmMaxRight = val; }
                for (int j = 0; j < len; j++) { int x = posRel + j; if (x >= 0 && x < gl) counts_[x][code(u.seq[j])] += 1; }
            } else discardedCount++;
        }
    }
    if (a_.partialFlag) {
        drawHeader(gapLength_);
        qualGap_.assign(std::max(allocArg_, 1), std::array<double, 5>{{0, 0, 0, 0, 0}});
        for (int q = 0; q < R; q++) {
            const PartialRead& p = in_.partial[q];
            const int len = prep_.readLen[q];
            totalCount++;
            double maxProb = 0; int posRel = -W;
            if (hr.p2max[q] > 0) { maxProb = hr.p2max[q]; posRel = hr.pos2[q]; }
            long refPos = p.refPos;
            if (!(p.pos < gapStart)) refPos += gapoffset;
            if (-log10(maxProb) < m_.gapProbCutOff || savedFinal_[0] == q || savedFinal_[1] == q) {
                if (posRel < 0) leftcount++; else rightcount++;
                if (posRel < 0 && posRel + len >= gapLength_) { if (-posRel >= 3 && (posRel + len - gapLength_) >= 3) leftcount = rightcount = 10; }
                partialReadFlag[q] = std::array<int, 3>{{1, posRel, len}};
                const long placed = posRel + gapStart;
                int newIsz = -1;
                if (posRel < 0) { if (refPos != -1) newIsz = (int)(placed - refPos + len); }
                else { if (refPos != -1) newIsz = (int)(refPos + len - placed); }
                drawRead(posRel, p.seq, q, newIsz, 'P');
                for (int j = 0; j < len; j++) {
                    int x = posRel + j;
                    if (x >= 0 && x < gl) {
                        int cd = code(p.seq[j]);
                        counts_[x][cd] += 1;
                        if (x < (int)qualGap_.size() && a_.partialFlag) {
                            // quality string cut at partial_read_len (token[partial_read_len] = '\0'); bases beyond it count 0
                            const size_t ql = std::min<size_t>(p.qual.size(), (size_t)std::max(0, partialReadLen_));
                            if ((size_t)j < std::max(ql, p.seq.size())) qualGap_[x][cd] += (size_t)j < ql ? phredError((unsigned char)p.qual[(size_t)j]) : 0.0;
                        }
                    }
                }
            } else discardedCount++;
        }
        if ((int)in_.partial.size() > R) totalCount++;   // the 3001st line is read (and counted) before the loop breaks
    }

    int usedRead = totalCount - discardedCount, recomputeFlag = 0;
    int Nflag[2] = {-1, -1}, lflag[2] = {-1, -1};
    auto clearCounts = [&](int a) { for (int j = 0; j < a && j < (int)counts_.size(); j++) for (int k = 0; k < 4; k++) counts_[j][k] = 0; };

    if (a_.unmapped == 1) {
        const int sideThresh = 4;
        if ((unmMaxLeft < 2 * sideThresh && unmMaxLeft > 0) || (unmMaxRight < 2 * sideThresh && unmMaxRight > 0)) {
            if (regionPerctMax_ > .75) usedRead = 0;
        }
        if ((unmMaxLeft < sideThresh && unmMaxLeft > 0) || (unmMaxRight < sideThresh && unmMaxRight > 0)) {
            std::vector<int> region;
            computeSequenceHost(1);
            int rc = findRegion(region);
            if (rc >= 1) {
                if (unmMaxLeft < sideThresh && unmMaxLeft > 0) lflag[0] = 1;
                if (unmMaxRight < sideThresh && unmMaxRight > 0) lflag[1] = 1;
            } else if (rc == 0) { usedRead = 0; unmMaxLeft = unmMaxRight = -1; }
        }
        if (leftRightCheck[0] == 0 && leftRightCheck[1] == 0 && usedRead != 0) { usedRead = 0; unmMaxLeft = unmMaxRight = -1; }
        if ((leftRightCheck[0] == 0 && leftStartZero != 0) || (leftRightCheck[1] == 0 && rightFinGlen != 0)) {
            std::vector<int> region;
            computeSequenceHost(1);
            int rc = findRegion(region);
            if (rc >= 1) { if (leftRightCheck[0] == 0) Nflag[0] = 1; if (leftRightCheck[1] == 0) Nflag[1] = 1; }
        }
        if (usedRead == 0 || !(leftRightCheck[0] == 1 && leftRightCheck[1] == 1)) {
            gapLength_ = og_;
            int offset = gapLength_ > gl ? 0 : (gl - gapLength_);
            clearCounts(gapLength_ + offset);
            auto recompute2 = [&]() {   // Figbird.cpp:4908-4927
                for (int q = 0; q < numReads_; q++) {
                    if (unmPosOrg_[q].b > 0) {
                        int pos = unmPosOrg_[q].a; const std::string& s = in_.unm[q].seq;
                        for (int j = 0; j < (int)s.size(); j++) if (pos + j >= 0 && pos + j < gapLength_) counts_[pos + j][code(s[j])] += 1;
                    }
                }
            };
            if (!leftRightCheck[0] && leftRightCheck[1] && unmMaxRight >= sideThresh) { recompute2(); recomputeFlag = 1; }
            else if (leftRightCheck[0] && !leftRightCheck[1] && unmMaxLeft >= sideThresh) { recompute2(); recomputeFlag = 1; }
        }
    }

    if (a_.partialFlag == 1) {
        int ret[2] = {0, 0};
        int uFlag = 1;
        std::vector<std::array<int, 2>> pf(partialReadCount_);
        for (int i = 0; i < partialReadCount_; i++) pf[i] = std::array<int, 2>{{partialReadFlag[i][0], partialReadFlag[i][1]}};
        detectOverlap(pf, nullptr, gapLength_, ret, 8);
        int gapCase;
        if ((og_ - gl) > 0 && ret[0] > 0) gapCase = 1;
        else if ((og_ - gl) > 0 && ret[0] == 0) gapCase = 2;
        else if ((og_ - gl) < 0 && ret[0] > 0) gapCase = 3;
        else if ((og_ - gl) < 0 && ret[0] == 0) gapCase = 4;
        else gapCase = 5;
        if (usedRead < kPartialThreshold || gapCase == 2 || gapCase == 4) {
            gapLength_ = og_;
            int offset = gapLength_ > gl ? 0 : (gl - gapLength_);
            clearCounts(gapLength_ + offset);
            if (usedRead < kPartialThreshold || gapCase == 4) uFlag = 0;
            else {
                // recompute1 (Figbird.cpp:4875-4906): unit votes at the placements found for the original length
                const int np = (int)std::min<size_t>(in_.partial.size(), 3001);
                for (int q = 0; q < np && q < (int)partialPosOrg_.size(); q++) {
                    if (partialPosOrg_[q][0] == 1) {
                        int pos = partialPosOrg_[q][1]; const std::string& s = in_.partial[q].seq;
                        for (int j = 0; j < (int)s.size(); j++) if (pos + j >= 0 && pos + j < gapLength_) counts_[pos + j][code(s[j])] += 1;
                    }
                }
                for (int i = 0; i < partialReadCount_; i++) pf[i] = std::array<int, 2>{{partialPosOrg_[i][0], partialPosOrg_[i][1]}};
                detectOverlap(pf, nullptr, gapLength_, ret, 8);
                if (ret[1] == -1) { clearCounts(gapLength_ + offset); uFlag = 0; }
            }
        }
        if (uFlag == 1 && ret[0] == 0 && ret[1] == 0) {
            for (int j = 0; j < gapLength_; j++) {
                int nonZero = 0;
                for (int k = 0; k < 4; k++) if (counts_[j][k] > 0) nonZero++;
                if (nonZero == 0) continue;
                int up = checkUpdate(counts_[j], j);
                if (up != -1) { if (up != 50) counts_[j][up] += 10; }
                else for (int k = 0; k < 4; k++) counts_[j][k] = 0;
            }
        }
    }

    computeSequenceHost(1);

    if (a_.unmapped && (leftRightCheck[0] || leftRightCheck[1] || usedRead != 0)) {
        if (Nflag[0] == 1) concensus_[0] = 'N';
        if (Nflag[1] == 1 && gapLength_ > 0) concensus_[gapLength_ - 1] = 'N';
        if (lflag[0] == 1) concensus_[0] = 'N';
        if (lflag[1] == 1 && gapLength_ > 0) concensus_[gapLength_ - 1] = 'N';
        int clearVal = recomputeFlag == 0 ? recheckSequence(unmPos) : recheckSequence(unmPosOrg_);
        if (clearVal == 1) {
            gapLength_ = og_;
            int offset = gapLength_ > gl ? 0 : (gl - gapLength_);
            clearCounts(gapLength_ + offset);
            computeSequenceHost(1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// fillGap (Figbird.cpp:6201-6570)
// ---------------------------------------------------------------------------------------------------------
GapResult GapFill::run(DeviceQueue& dev, int batchGapIndex) {
    dev_ = &dev; bidx_ = batchGapIndex;
    GapResult out;
    const bool attempt = prep_.attempt;
    const int unmapped = a_.unmapped, partialFlag = a_.partialFlag;
    gapLength_ = og_;

    auto emitUnfilled = [&]() {
        out.gapStringLength = og_; out.gapString.assign(og_, 'N'); out.gapToFill = 0; out.drawText = draw_;
        return out;
    };
    if (!prep_.supported) {
        fprintf(stderr, "figbird_b200: gap %d lies within a read length of a scaffold end (the reference reads outside its window there); left unfilled\n", in_.rec.gapNo);
        return emitUnfilled();
    }
    if (!attempt) return emitUnfilled();    // num_itr=0: one initialize + computeSequence(0,0) => og x 'N' (Figbird.cpp:6223-6233)

    int finalizeFlag = 1;
    if (unmapped) { if (largeGapFlag_ == 0) finalizeFlag = 0; }
    const int gapMin = (int)((float)og_ * frac1_), gapMax = (int)((float)og_ * frac2_);
    int maxGapEstimate = gapMin;
    double maxLikelihood = -DBL_MAX, secondMaxLikelihood = -DBL_MAX;
    double likelihood = 0, prevlikelihood = 0;
    int fillOrNot = 0;
    int sameCount = 0, stuckCount = 0;
    const int sameThresh = unmapped ? 50 : 4;
    int range = gapMax - gapMin + 1;
    std::vector<int> usedReadArr(std::max(range, 1), 0);
    int lessReadFlag = 0;

    if (unmapped && og_ <= midLimitU_) lessReadFlag = checkGapReads();
    if (lessReadFlag == 1) range = 0;

    int prevBest = -1, currBest = 0, prevU = -1, currU = 0, secSame = 0, secSame2 = 0;
    int j = 0;
    int gapEstimate = gapMin;

    // first candidate only: negative-overlap shortcut (initialize -> find_contig_match, Figbird.cpp:2375,6301-6306)
    if (range > 0) {
        fillOrNot = sideLimit_ > 0 ? findContigMatch() : 0;
        if (oneSideRepeat_ == 1) fillOrNot = 0;
    }
    if (fillOrNot != 0) {
        gapLength_ = 0;
        out.gapStringLength = 0; out.gapString.clear(); out.gapToFill = fillOrNot; out.drawText = draw_;
        out.refPlacements = refPlacements_;      // the checkGapReads probes ran before initialize() found the overlap
        return out;
    }

    // speculative evaluation: candidates are scored in chunks on the device, then the reference's
    // sequential scan is replayed over the results; work past an early exit is discarded.
    // candidates per device request: more means fewer engine calls (ticks) but more work past an early exit
    static const int chunkP = [] { const char* e = getenv("FIGBIRD_CHUNK_PARTIAL"); return e ? std::max(1, atoi(e)) : 16; }();
    static const int chunkU = [] { const char* e = getenv("FIGBIRD_CHUNK_UNMAPPED"); return e ? std::max(1, atoi(e)) : 24; }();
    const int chunkBase = largeGapFlag_ ? 1 : (partialFlag ? chunkP : chunkU);
    // tail of a run: once the lane's batch is down to a few gaps a tick no longer fills the device, so the candidates still to scan
    // can be requested further ahead (FIGBIRD_TAIL_ITEMS = items per tick to aim for).  Off by default: measured on C4 / 8 GPUs it
    // saves two or three ticks but every tick still lasts as long as its slowest EM chain (~0.1 s), and the extra candidates
    // cost 2 % more kernel time (profiles/README.md).
    static const int tailItems = [] { const char* e = getenv("FIGBIRD_TAIL_ITEMS"); return e ? std::max(0, atoi(e)) : 0; }();
    std::vector<ItemResult> chunkRes; int chunkFirst = 0;
    bool broke = false;
    for (; j < range; j++) {
        umaxFlags_ = 0;
        discont_ = 0; compCount_ = 0; overlapThreshold_ = 5;
        regionPerct_ = 0;
        if (unmapped && largeGapFlag_) {
            likelihood = largeGapRounds(gapEstimate, finalizeFlag, largeGapFlag_, false);
        } else {
            if (j >= chunkFirst + (int)chunkRes.size()) {
                chunkFirst = j;
                int chunk = chunkBase;
                { const int active = std::max(1, dev_->activeGaps()); if ((long)active * chunkBase < tailItems) chunk = std::min(192, std::max(chunkBase, tailItems / active)); }
                std::vector<ItemSpec> specs;
                for (int c = 0; c < chunk && j + c < range; c++) {
                    ItemSpec s = emSpec(gapEstimate + c);
                    if (unmapped && !finalizeFlag) s.flags |= FB_FLAG_EXTRA_PASS;
                    // the first candidate inherits previous_str from the last checkGapReads probe; it can only match at equal length
                    // (gapMin == og/2, the first probe, for N-runs of 134..400 bases)
                    if (unmapped && j == 0 && c == 0 && prevStrLen_ == gapEstimate && gapEstimate > 0) s.stringIn = prevStr_;
                    specs.push_back(s);
                }
                dev_->submit(bidx_, specs, chunkRes);
            }
            const ItemResult& r = chunkRes[j - chunkFirst];
            if (unmapped && !finalizeFlag) {
                // EM rounds ran with finalize_flag=0; the extra pass is the call with finalize_flag=1 (Figbird.cpp:6348-6352)
                validCount_ = 0; refPlacements_ += r.placements;
                initializeEffects(gapEstimate);
                likelihood = unmappedEpilogue(r, 0, gapEstimate, 1, 0, r.calls - 1);
            } else likelihood = evalCandidate(r, gapEstimate, finalizeFlag);
            lastSoft_.assign(r.soft, r.soft + gapEstimate);
        }
        setConcensus(lastSoft_, gapEstimate);          // computeSequence(0,0)
        gapLength_ = gapEstimate;

        if (likelihood > maxLikelihood) {
            secondMaxLikelihood = maxLikelihood;
            maxLikelihood = likelihood; maxGapEstimate = gapEstimate;
            copyStr(bestString_, concensus_);
            for (int k = 0; k < numReads_; k++) savedReads_[k] = markAccepted_[k];
            regionPerctMax_ = regionPerct_;
            savedFinal_[0] = savedTemp_[0]; savedFinal_[1] = savedTemp_[1];
            currBest = j; prevU = validCount_;
        } else if (likelihood > secondMaxLikelihood) secondMaxLikelihood = likelihood;
        if (gapEstimate == og_) copyStr(originalStr_, concensus_);

        usedReadArr[j] = validCount_;
        const double diff1 = std::abs(prevlikelihood - likelihood);
        if (diff1 <= 0.9) sameCount++; else sameCount = 0;
        prevlikelihood = likelihood;
        auto refreshOriginal = [&]() {
            runLength(og_, 1, 0);
            setConcensus(lastSoft_, og_); gapLength_ = og_;
            copyStr(originalStr_, concensus_);
        };
        if (sameCount == sameThresh) { if (gapLength_ < og_) refreshOriginal(); broke = true; break; }
        if (unmapped) {
            currU = validCount_;
            if (currBest == prevBest && std::abs(currU - prevU) <= 2) secSame++;
            else { prevBest = currBest; secSame = 0; }
            if (secSame >= 2 * sameThresh) { if (gapLength_ < og_) refreshOriginal(); broke = true; break; }
            if (og_ <= 30) {
                if (!(umaxFlags_ & 7)) secSame2++; else secSame2 = 0;
                if (secSame2 >= 1.5 * sameThresh) { if (gapLength_ < og_) refreshOriginal(); broke = true; break; }
            }
            if (discont_ == 1 && validCount_ < 5) stuckCount++; else stuckCount = 0;
            if (stuckCount > 3 * sameThresh) { if (gapLength_ < og_) refreshOriginal(); broke = true; break; }
        }
        gapEstimate++;
    }
    (void)broke;

    if (unmapped) {
        if (lessReadFlag == 1) {
            runLength(og_, 1, 0);
            setConcensus(lastSoft_, og_); gapLength_ = og_;
            copyStr(originalStr_, concensus_);
            finalize(og_);
        } else {
            bool change = false;
            for (int i = 1; i < j; i++) if (usedReadArr[0] != usedReadArr[i]) { change = true; break; }
            if (j == 1) change = false;
            if (change) finalize(maxGapEstimate);
            else { copyStr(bestString_, originalStr_); finalize(og_); }
        }
    } else {
        if (maxGapEstimate == 0) {
            if (usedReadArr[0] != 0) finalize(maxGapEstimate);
            else {
                if (maxGapEstimate < og_) {
                    runLength(og_, 1, 0);
                    setConcensus(lastSoft_, og_); gapLength_ = og_;
                    copyStr(originalStr_, concensus_);
                }
                finalize(og_);
            }
        } else finalize(maxGapEstimate);
    }

    out.gapStringLength = gapLength_;
    out.gapString.assign(concensus_.data(), strnlen(concensus_.data(), (size_t)gapLength_));   // printed with %s
    out.gapToFill = 0;
    out.drawText.swap(draw_);
    out.refPlacements = refPlacements_;
    // the per-gap working set is not needed once the result is out
    std::vector<std::array<double, 5>>().swap(counts_); std::vector<std::array<double, 5>>().swap(qualGap_);
    return out;
}

}  // namespace fb
