// fillgaps -- executable wrapper: the command RunFigbird.sh:352,480 runs instead of `g++ FillGaps.cpp && ./a.out`.
#include "../../include/figbird_b200.h"
int main(int argc, char** argv) { return fb_fillgaps_main(argc, (const char* const*)argv); }
