# Mechanical instrumentation of the reference worker (test / bench infrastructure): placement counters.
# Streams /root/reference/Figbird.cpp to stdout with a counter incremented at the four statements that end the
# scoring of one (read, admissible offset) pair in pass 1 of GapFiller::placeReads --
#   Figbird.cpp:3169, 3236  `tempProb = log(tempProb);`    (partial mode, left / right reads)
#   Figbird.cpp:3591, 3658  `tempProb = log10(tempProb);`  (unmapped mode, mate left / right)
# -- i.e. BASELINE.md 3.3 / SURVEY.md 8d's unit of work, and one line written at exit (before main's final
# `return 0;`, Figbird.cpp:7507) to $FB_COUNT_DIR/count_<pid>.txt.  Nothing else changes: every file the worker
# writes stays byte-identical (asserted by tests/make_callcounts.py).
BEGIN { print "#include <unistd.h>"; print "static long fb_p1_placements = 0;"; }
{
    if ($0 ~ /^\treturn 0;[ \t\r]*$/) {
        print "    { const char* fbd = getenv(\"FB_COUNT_DIR\"); if (fbd) { char fbp[1200]; snprintf(fbp, sizeof fbp, \"%s/count_%d.txt\", fbd, (int)getpid());";
        print "        FILE* fbf = fopen(fbp, \"w\"); if (fbf) { fprintf(fbf, \"%ld\\n\", fb_p1_placements); fclose(fbf); } } }";
        dumps++;
    }
    print;
    if ($0 ~ /^[ \t]*tempProb = log\(tempProb\);[ \t\r]*$/ || $0 ~ /^[ \t]*tempProb = log10\(tempProb\);[ \t\r]*$/) { print "                        fb_p1_placements++;"; sites++; }
}
END { if (dumps != 1) { print "#error count_patch.awk: expected one final return 0; found " dumps; }
      if (sites != 4) { print "#error count_patch.awk: expected 4 pass-1 sites, found " sites; } }
