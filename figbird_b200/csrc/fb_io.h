#pragma once
#include "fb_host.h"

namespace fb {
bool loadGapRecords(const std::string& tmpDir, std::vector<GapRecord>& gaps, int& totGaps);
bool loadPartial(const std::string& path, std::vector<PartialRead>& out, bool& exists);
bool loadUnmapped(const std::string& path, int readLen, std::vector<UnmappedRead>& out, int& pairCount);
// the same readers over the text of a per-gap file that is already in memory
void loadPartialText(const char* text, size_t n, std::vector<PartialRead>& out);
void loadUnmappedText(const char* text, size_t n, std::vector<UnmappedRead>& out, int& pairCount);
// opt-in per-gap container (FIGBIRD_CONTAINER=1): all per-gap files of one kind in one mapped file
class GapContainer {
public:
    ~GapContainer() { close(); }
    bool open(const std::string& path, uint32_t kind, size_t nGaps);      // false: absent or not a container of this kind / gap count
    void close();
    bool valid() const { return base_ != nullptr; }
    const char* text(size_t g, size_t& n) const { n = (size_t)(off_[g + 1] - off_[g]); return base_ + off_[g]; }
private:
    const char* base_ = nullptr; const uint64_t* off_ = nullptr; size_t n_ = 0, size_ = 0;
};
bool writeGapContainer(const std::string& path, uint32_t kind, const std::vector<std::string>& texts);
bool writeGapout(const std::string& path, const std::vector<GapRecord>& gaps, const std::vector<GapResult>& res);
bool writeFilledContigs(const std::string& tmpDir, const Scaffolds& sc, const std::vector<GapRecord>& gaps,
                        const std::vector<GapResult>& res, int totGaps);
}  // namespace fb
