#pragma once
#include "fb_host.h"

namespace fb {
bool loadGapRecords(const std::string& tmpDir, std::vector<GapRecord>& gaps, int& totGaps);
bool loadPartial(const std::string& path, std::vector<PartialRead>& out, bool& exists);
bool loadUnmapped(const std::string& path, int readLen, std::vector<UnmappedRead>& out, int& pairCount);
bool writeGapout(const std::string& path, const std::vector<GapRecord>& gaps, const std::vector<GapResult>& res);
bool writeFilledContigs(const std::string& tmpDir, const Scaffolds& sc, const std::vector<GapRecord>& gaps,
                        const std::vector<GapResult>& res, int totGaps);
}  // namespace fb
