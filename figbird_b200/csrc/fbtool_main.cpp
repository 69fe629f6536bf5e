// fbtool -- the host-only pipeline tools as one executable: `fbtool <preprocess|combinegaps|flanktrim|reduce_scf|reverse> <args of the reference program>`
#include <cstdio>
#include <cstring>
#include "../../include/figbird_b200.h"
int main(int argc, char** argv) {
    if (argc >= 2) {
        const char* t = argv[1];
        const int n = argc - 1; const char* const* a = (const char* const*)(argv + 1);
        if (!strcmp(t, "preprocess")) return fb_preprocess_main(n, a);
        if (!strcmp(t, "combinegaps")) return fb_combinegaps_main(n, a);
        if (!strcmp(t, "flanktrim")) return fb_flanktrim_main(n, a);
        if (!strcmp(t, "reduce_scf")) return fb_reduce_scf_main(n, a);
        if (!strcmp(t, "reverse")) return fb_reverse_main(n, a);
    }
    fprintf(stderr, "usage: fbtool <preprocess|combinegaps|flanktrim|reduce_scf|reverse> <arguments of the reference program>\n");
    return 1;
}
