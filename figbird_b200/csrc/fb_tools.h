// fb_tools.h -- shared helpers of the host-only pipeline tools (fb_tools.cpp, fb_preprocess.cpp)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/figbird_b200.h"

namespace fb {

bool slurp(const std::string& path, std::string& out);

// scaffolds as the reference's shared FASTA reader sees them (1023-byte fgets records, case kept)
struct RawFasta {
    std::vector<std::string> names;      // first token of every header record (pushed for every header)
    std::vector<std::string> headers;    // per pushed scaffold: text of the header record in force when it was flushed
    std::vector<std::string> seq;
    std::vector<char> hasN;              // any 'N' / 'n' in the records of the scaffold
    std::string pendingHeader;
};
bool loadFastaRaw(const std::string& path, RawFasta& out);

}  // namespace fb
