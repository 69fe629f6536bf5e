# Mechanical instrumentation of the reference worker (test infrastructure).
# Streams /root/reference/Figbird.cpp to stdout with ONE insertion: just before the
# `return maxLikelihood;` that ends GapFiller::placeReads (Figbird.cpp:4386) it appends a record to
# the file named by $FB_DUMP:  call sequence number, gap, candidate length, round, likelihood,
# valid_count, then the countsGap gap rows [W, W+Lg) x 5 with %.17g, and the soft/hard consensus.
# Nothing else changes, so every file the worker writes stays byte-identical (checked by tests).
{
    if ($0 ~ /return maxLikelihood;/ && !done) {
        print "    { static FILE* fbd=NULL; static long fbseq=0; const char* fbn=getenv(\"FB_DUMP\");";
        print "      if(fbn){ if(!fbd)fbd=fopen(fbn,\"a\"); fprintf(fbd,\"CALL %ld gap %d Lg %d round %d fin %d like %.17g valid %d\\n\",fbseq++,g,gapLength,ge,finalize_flag,maxLikelihood,valid_count);";
        print "        for(long fi=left_maxDistance;fi<left_maxDistance+gapLength;fi++){fprintf(fbd,\"%.17g %.17g %.17g %.17g %.17g\\n\",countsGap[fi][0],countsGap[fi][1],countsGap[fi][2],countsGap[fi][3],countsGap[fi][4]);}";
        print "        fflush(fbd);} }";
        done = 1;
    }
    if ($0 ~ /Done with insertthteshold/ && !mdone) {
        print "    { const char* fbn=getenv(\"FB_DUMP_MODEL\"); if(fbn){ FILE* fm=fopen(fbn,\"w\");";
        print "      fprintf(fm,\"mean %.17g leftSD %.17g rightSD %.17g tmin %d tmax %d cutoff %d maxins %d maxread %d\\n\",insertSizeMean,leftSD,rightSD,insertThresholdMin,insertThresholdMax,gapProbCutOff,maxInsertSize,maxReadLength);";
        print "      for(int fi=0;fi<maxReadLength;fi++)fprintf(fm,\"pos %d %.17g %.17g %.17g\\n\",fi,errorPosDist[fi],inPosDist[fi],delPosDist[fi]);";
        print "      for(int fi=0;fi<5;fi++)fprintf(fm,\"etp %d %.17g %.17g %.17g %.17g %.17g\\n\",fi,errorTypeProbs[fi][0],errorTypeProbs[fi][1],errorTypeProbs[fi][2],errorTypeProbs[fi][3],errorTypeProbs[fi][4]);";
        print "      for(int fi=0;fi<maxInsertSize;fi++)if(fi<1200||fi%97==0)fprintf(fm,\"pdf %d %.17g\\n\",fi,insertLengthDistSmoothed[fi]);";
        print "      for(int fi=0;fi<60;fi++)fprintf(fm,\"gapprob %d %ld\\n\",fi,gapProbs[fi]);";
        print "      fclose(fm);} }";
        mdone = 1;
    }
    print;
}
