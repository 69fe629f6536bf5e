// fb_gapfiller.h -- per-gap host control on top of the device engine.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "fb_host.h"

namespace fb {

// One device request (mirrors FbWorkItem with owned buffers).
struct ItemSpec {
    int kind = FB_ITEM_EM;
    int candLen = 0;
    int maxRounds = 0;
    int flags = 0;
    int compIn = 0;
    std::vector<double> countsIn;
    std::vector<uint8_t> stringIn;
};

// Blocking batch interface: a gap submits a group of items and gets their results; the implementation
// (fb_engine.cpp) merges the groups of all gaps in flight into one fb_em_run per GPU.
class DeviceQueue {
public:
    virtual ~DeviceQueue() {}
    virtual void submit(int batchGapIndex, const std::vector<ItemSpec>& items, std::vector<ItemResult>& results) = 0;
    // gaps that took part in the queue's most recent engine call (a large number before the first one): lets a gap speculate
    // further ahead when the batch has become too small to fill the device (the tail of a run)
    virtual int activeGaps() const { return 1 << 30; }
};

// Inputs of one gap as read from disk.
struct GapInput {
    GapRecord rec;
    std::vector<PartialRead> partial;   // first 3001 lines of partial_gaps_<g>.sam
    bool partialExists = false;
    std::vector<UnmappedRead> unm;      // first 3000 pairs of gaps_<g>.sam (unmapped mode)
    int unmPairCount = 0;
};

// Device-side view of a gap (what goes into FbGapBatch), produced by GapFill::prepare().
struct PreparedGap {
    bool supported = true;              // false: window would leave the scaffold (reference behaviour undefined)
    bool attempt = false;               // analyzeGap verdict (Figbird.cpp:6168-6199)
    int mode = 0, origLen = 0, flankLen = 0;
    long gapStart = 0;
    std::vector<uint8_t> flank;         // 2*flankLen codes
    std::vector<int32_t> readLen, readMate;
    std::vector<uint8_t> readFlags, readJlo, readJcut;
    std::vector<std::vector<uint8_t>> readCodes;
    std::vector<int32_t> pileL, pileR;  // [t][4]
    int pileLen = 0;
    double cost = 0;                    // sharding cost estimate (placements)
    bool sequential = false;            // the gap is a long chain of single-round device requests (unmapped mode, N-run > 400)
};

class GapFill {
public:
    GapFill(const Args& a, const Model& m, const Scaffolds& sc, GapInput&& in);
    void prepare();                               // host-only analysis + encoding (no device)
    const PreparedGap& prepared() const { return prep_; }
    GapResult run(DeviceQueue& dev, int batchGapIndex);
    const GapRecord& record() const { return in_.rec; }

private:
    struct Pos3 { int a, b, c; };
    const Args& a_; const Model& m_; const Scaffolds& sc_;
    GapInput in_;
    PreparedGap prep_;
    DeviceQueue* dev_ = nullptr; int bidx_ = -1;

    // ---- reference GapFiller members (Figbird.cpp:1563-1635), window geometry dropped
    int og_ = 0, unmReadLen_ = 0, partialReadLen_ = 0, midLimitP_ = 0, midLimitU_ = 0, negOverlap_ = 0;
    int largeGapFlag_ = 0, allocArg_ = 0, fillflag_ = 1;
    float frac1_ = 1, frac2_ = 1;
    std::string gapLeft_, gapRight_;
    int partialReadCount_ = 0, numReads_ = 0;
    std::vector<std::array<int, 3>> repeatflag_;
    int repFlag_ = 0, oneSideRepeat_ = 0;
    int validCount_ = 0, discont_ = 0, compCount_ = 0;
    double regionPerct_ = 0, regionPerctMax_ = 0;
    std::vector<int> markAccepted_, savedReads_;
    std::vector<char> mlvNonZero_;
    std::vector<Pos3> finalReadpos_, unmPosOrg_;
    std::vector<std::array<int, 3>> partialPosOrg_;
    int umaxFlags_ = 0;
    int gapLength_ = 0;                     // this->gapLength
    std::vector<char> concensus_, bestString_, originalStr_;   // C buffers with strcpy semantics
    std::vector<int> gapCoverage_;
    std::vector<std::array<double, 5>> counts_;   // countsGap gap rows (host copy, finalize / border update)
    std::vector<std::array<double, 5>> qualGap_;
    // The reference keeps `char partial_left[100], partial_right[100]; int partial_saved_read_temp[2], partial_saved_read_final[2];
    // int side_limit;` side by side (Figbird.cpp:1625-1627) and update_partial_prob writes the pile-up strings without a bound
    // (:2063-2084): with reads longer than ~105 bases partial_left runs into partial_right and partial_right into the saved-read
    // indices and side_limit, which finalize() then reads (:5345).  The layout of the x86-64 build is reproduced byte for byte
    // so that those reads see what the reference's see; writes past side_limit (pointers in the reference) are dropped.
    alignas(8) unsigned char ovl_[224];
    char* const pileStr_ = (char*)ovl_;                       // [0,100) partial_left, [100,200) partial_right
    int32_t* const savedTemp_ = (int32_t*)(ovl_ + 200);      // partial_saved_read_temp[2]
    int32_t* const savedFinal_ = (int32_t*)(ovl_ + 208);     // partial_saved_read_final[2]
    int32_t& sideLimit_ = *(int32_t*)(ovl_ + 216);           // side_limit
    static constexpr int kOvlBytes = 220;
    int overlapThreshold_ = 5;
    std::string draw_;
    std::vector<uint8_t> lastSoft_;         // computeSequence(0,0) of the most recent placeReads call
    int64_t refPlacements_ = 0;
    int prevStrLen_ = -1; std::vector<uint8_t> prevStr_;    // previous_str as the last checkGapReads probe left it

    // constants of setParameters (Figbird.cpp:6157-6165)
    static constexpr int kCov1 = 0, kCov2 = 1, kMatchDiscont = 4, kPartialThreshold = 2, kClipThresh = 2;
    static constexpr int kMaxGap = 100000, kNumItr = 200;

    // ---- helpers
    bool analyze();
    int findRepeat();
    static double phredError(unsigned char ch);
    int findContigMatch() const;
    void pileUp(int Lg, bool makeStrings);
    void initializeEffects(int Lg);     // what initialize() leaves behind on the host besides the gap rows (see ovl_)
    int maxRightCi_ = -1;               // longest right-side pile-up of the gap (-1: not computed yet)
    // update_partial_prob's vote tables of the gap (host copy, built on first use): rows from the left / right edge
    bool hostPileBuilt_ = false, hostAnyLeft_ = false, hostAnyRight_ = false;
    int hostMaxExt_ = -(1 << 30), hostMaxCi_ = -(1 << 30);
    std::vector<std::array<int, 4>> hostPileL_, hostPileR_;
    void setConcensus(const std::vector<uint8_t>& codes, int len);
    void setConcensus(const uint8_t* codes, int len);
    void copyStr(std::vector<char>& dst, const std::vector<char>& src);
    ItemSpec emSpec(int Lg) const;
    double evalCandidate(const ItemResult& r, int Lg, int finalizeFlag);      // epilogue of the last call(s)
    double unmappedEpilogue(const ItemResult& r, int slot, int Lg, int finalizeFlag, int updateFlag, int ge);
    double partialEpilogue(const ItemResult& r, int slot, int Lg, bool savedOnly = false, bool* touchedSaved = nullptr);
    std::vector<std::array<int, 2>> scratchPflag_; std::vector<int> scratchLeft_, scratchRight_, scratchSm_;
    void borderUpdate(int Lg);
    double findOverlapUnmapped(int Lg);
    void detectOverlap(const std::vector<std::array<int, 2>>& pflag3, const std::vector<std::array<int, 3>>* pflagFinal, int gaplen, int* ret, int lenThresh, bool* touchedSaved = nullptr);
    double runLength(int Lg, int finalizeFlag, int c);        // GapFiller::run
    double largeGapRounds(int Lg, int finalizeFlag, int updateFlag, bool extraPass);
    int checkGapReads();
    void finalize(int gapLength);
    void computeSequenceHost(int check);
    int findRegion(std::vector<int>& region) const;
    int recheckSequence(const std::vector<Pos3>& pos);
    int checkUpdate(const std::array<double, 5>& arr, int j) const;
    void drawHeader(int length);
    void drawRead(int length, const std::string& s, int readno, int isz, char type);
};

}  // namespace fb
