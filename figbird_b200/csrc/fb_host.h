// fb_host.h -- host-side data structures of the B200 gap filler (the program behind fb_fillgaps_main).
// The host side re-implements the *control* of the reference worker (Figbird.cpp main + GapFiller) around
// the device engine of include/figbird_b200.h; the scored loops themselves run only on the GPU.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../include/figbird_b200.h"

namespace fb {

// ---- command line of FillGaps (FillGaps.cpp:419-433)
struct Args {
    std::string draft;         // 1 contigFileName
    int maxDistance = 0;       // 2 (atoi)
    int readLength = 0;        // 3 max read length of the library in use
    int scriptItr = 0;         // 4
    int partialFlag = 0;       // 5
    int unmapped = 0;          // 6
    int numThreads = 1;        // 7 (advisory here)
    std::string myout;         // 8 myout.sam
    std::string tmpDir;        // 9 (trailing '/')
    std::string gapsDir;       // 10 (trailing '/')
    int negOverlap = 0;        // 11
    int partialReadLen = 0;    // 12
    int trim = 0;              // 13 (unused downstream)
    int setInputMean = 0;      // 14
    int insertSizeMean = 0;    // 15
};

// ---- scaffolds (FASTA loader reproduces the fgets(1024) chunking of Figbird.cpp:6986-7046)
struct Scaffolds {
    std::vector<std::string> names;
    std::vector<std::string> seq;   // upper-cased (Figbird.cpp:7052-7058)
    long totalLength = 0;
};
bool loadScaffolds(const std::string& path, Scaffolds& out);

// ---- learned model (Figbird.cpp:7087-7200)
struct Model {
    int maxReadLength = 0;
    long totalCount = 0, unCount = 0;
    int maxInsertSize = 0;          // may grow past MAX_INSERT_SIZE exactly like updateInsertCounts
    std::vector<double> errorPosDist, inPosDist, delPosDist;
    double errorTypeProbs[5][5];
    std::vector<double> insertPdfSmoothed;
    double insertSizeMean = 0, leftSD = 0, rightSD = 0;
    int insertThresholdMin = 0, insertThresholdMax = 0;
    int gapProbCutOff = 0;
    long uniqueMappedReads = 0;
};
// returns false (message in err) when an input cannot be opened -- the caller exits 1 like the reference
// waitScaffolds (optional): called once before `sc` is first read; false = the scaffolds could not be loaded
bool learnModel(const Args& a, const Scaffolds& sc, Model& m, std::string& err, const std::function<bool()>& waitScaffolds = nullptr);

// ---- one gap line of gapInfo.txt + stat2.txt (Figbird.cpp:7359-7374)
struct GapRecord {
    int gapNo = 0, contigNo = 0;
    long gapStart = 0;
    int gapLength = 0;             // originalGap
    int stat1 = 0, stat2 = 0, stat3 = 0;   // gaptofill, perfectread_gap, perfectread_gaplen
};

// ---- a partial read line (partial_gaps_<g>.sam, Figbird.cpp:3089-3106)
struct PartialRead {
    std::string seq;
    int clippedIndex = 0, match = 0, pos = 0, refPos = 0;
    std::string qual;
};
// ---- an unmapped pair (gaps_<g>.sam, parseUnmapped Figbird.cpp:5661-5767)
struct UnmappedRead {
    std::string seq;      // reference orientation as the reference scores it
    int matePos = 0;      // pos_reads[q]
    int isReverse = 0;
};

// ---- result of filling one gap (what Figbird.cpp:7411-7413, 7463-7466 write)
struct GapResult {
    int gapStringLength = 0;
    std::string gapString;
    int gapToFill = 0;          // gaptofill[g]
    std::string drawText;       // this gap's part of draw.txt
    int64_t refPlacements = 0;  // reference-equivalent pass-1 placements actually consumed by the scan
};

// ---- device queue: batches EM/HARD items of many gaps into one fb_em_run (one per GPU)
struct ItemResult {      // views into the engine's pinned result arena: valid until the owning gap submits its next request
    int calls = 0, compCount = 0, flags = 0, nReads = 0, candLen = 0, nSlots = 0;
    int64_t placements = 0;
    const double* p1max = nullptr; const double* p2max = nullptr;   // [slot][read]
    const int32_t* pos2 = nullptr;
    const uint8_t* soft = nullptr; const uint8_t* hard = nullptr;    // [candLen]
    const int32_t* cov = nullptr;
    const double* counts = nullptr;     // [row][5] or null
};

class Engine;   // fb_engine.cpp

}  // namespace fb
