/*
 * figbird_b200.h -- C ABI of the B200-native Figbird gap-fill engine.
 *
 * The reference (SumitTarafder/Figbird) has no library boundary at all: RunFigbird.sh:352,480 compile and run
 * FillGaps.cpp, which spawns Figbird.cpp worker processes; everything is files + argv.  This header is the
 * "thin C-ABI layer" BASELINE.json:north_star asks for.  It has two levels:
 *
 *   Level 1 (drop-in):  fb_fillgaps_main()   == the whole `FillGaps` executable (FillGaps.cpp:371-947 plus the
 *                       workers it spawns, Figbird.cpp:6909-7508) behind one call with the same 15 positional
 *                       arguments, reading and writing the same files.  A cgo/JNI/ctypes host binds this one.
 *
 *   Level 2 (engine):   fb_ctx_* / fb_model_upload / fb_batch_upload / fb_em_run -- the placement loop
 *                       GapFiller::placeReads + computeProbsGap + computeErrorProbsGap + computeSequence
 *                       (Figbird.cpp:3022-4387, 2090-2137, 4417-4508) for a batch of (gap, candidate length)
 *                       work items.  The host program (level 1) is written on top of exactly these calls.
 *
 * Conventions: plain C types, caller-owned inputs, no exceptions, no exit(); every call returns fb_status
 * (0 = ok, <0 = error; fb_last_error() gives the text).  One fb_ctx per GPU; calls on one ctx are not
 * thread-safe; different ctxs are independent.  There is NO CPU fallback: fb_ctx_create fails when no
 * sm_100-class device is present.
 *
 * Base codes follow charCodes[] (Figbird.cpp:7060-7082): A=0 C=1 G=2 T=3, everything else (N, lower case) = 4.
 */
#ifndef FIGBIRD_B200_H
#define FIGBIRD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t fb_status;
#define FB_OK 0
#define FB_ERR_ARG (-1)
#define FB_ERR_NODEVICE (-2)
#define FB_ERR_CUDA (-3)
#define FB_ERR_NOMEM (-4)
#define FB_ERR_STATE (-5)
#define FB_ERR_IO (-6)

typedef struct fb_ctx fb_ctx;

/* ------------------------------------------------------------------------------------------------------
 * Model tables: what Figbird.cpp main():7118-7200 learns from myout.sam (processMapping :846,
 * computeProbabilites :497, computeLikelihood :1156).  Uploaded once per context.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t max_read_len;        /* maxReadLength (stat.txt, Figbird.cpp:7088); length of the three arrays below */
    const double* err_pos;       /* errorPosDist[k]            Figbird.cpp:533-538 */
    const double* ins_pos;       /* inPosDist[k]               Figbird.cpp:540-545 */
    const double* del_pos;       /* delPosDist[k]              Figbird.cpp:561-566 */
    double err_type[25];         /* errorTypeProbs[from][to]   Figbird.cpp:502-512 (row-major 5x5) */
    int32_t n_insert;            /* maxInsertSize; length of insert_pdf */
    const double* insert_pdf;    /* insertLengthDistSmoothed[t] Figbird.cpp:646-677 */
    int32_t insert_min;          /* insertThresholdMin  Figbird.cpp:7193,7198 */
    int32_t insert_max;          /* insertThresholdMax  Figbird.cpp:7194,7199 */
    int32_t prob_cutoff;         /* gapProbCutOff       Figbird.cpp:7170-7178 */
} FbModel;

/* ------------------------------------------------------------------------------------------------------
 * Gap batch: per-gap inputs of GapFiller (flanks from the scaffold, candidate reads from
 * Gaps/gaps_<g>.sam (parseUnmapped, Figbird.cpp:5661) or Gaps/partial_gaps_<g>.sam (placeReads :3089-3106)),
 * already decoded to base codes.  All arrays are concatenations; per-gap / per-read offsets index into them.
 * Coordinates are relative to the gap start: x = scaffold position - gapStart.
 * ---------------------------------------------------------------------------------------------------- */
#define FB_MODE_PARTIAL 0  /* partial_flag=1, unmapped=0 */
#define FB_MODE_UNMAPPED 1 /* partial_flag=0, unmapped=1 */

typedef struct {
    int64_t gap_start;        /* gapStart (0-based scaffold position of the first N; gapInfo.txt column 2) */
    int32_t mode;             /* FB_MODE_* */
    int32_t orig_len;         /* originalGap: length of the N-run (gapInfo.txt column 3) */
    int32_t n_reads;          /* reads scored by placeReads (<= 3000) */
    int32_t read_begin;       /* index of this gap's first read in the read arrays */
    int32_t flank_len;        /* F: number of scaffold bases given on each side (>= max read length - 1) */
    int32_t flank_begin;      /* offset into flank_codes: F left-flank codes (x=-F..-1) then F right-flank codes
                                 (scaffold positions gapStart+orig_len .. +F-1) */
    int32_t pile_len;         /* T: rows of the partial-read pile-ups below */
    int32_t pile_begin;       /* offset (in rows of 4 int32) into pile_left / pile_right */
} FbGap;

typedef struct {
    int32_t n_gaps;
    const FbGap* gaps;
    int32_t n_reads;
    const int32_t* read_len;      /* strlen of the read */
    const int64_t* read_code_off; /* offset of the read's codes in read_codes */
    const int32_t* read_mate;     /* mate position minus gapStart: pos_reads[q] (unmapped, Figbird.cpp:5708) or
                                     ref_pos = column 6 of the partial file (Figbird.cpp:3106); see read_flags */
    const uint8_t* read_flags;    /* FB_READ_* bits */
    const uint8_t* read_jlo;      /* first scored read index (read_start: clip_thresh for match 1/4, Figbird.cpp:3117) */
    const uint8_t* read_jcut;     /* bases not scored at the end (read_end, Figbird.cpp:3118) */
    int64_t n_codes;
    const uint8_t* read_codes;    /* base codes 0..4, reference orientation as the reference scores them */
    int64_t n_flank;
    const uint8_t* flank_codes;
    int64_t n_pile_rows;
    const int32_t* pile_left;     /* [row t][4]: partial-read votes for gap row t counted from the left edge
                                     (update_partial_prob, Figbird.cpp:1981-1996; an N base votes for all four) */
    const int32_t* pile_right;    /* [row u][4]: votes for gap row Lg-1-u counted from the right edge (:1997-2011) */
} FbGapBatch;

#define FB_READ_LEFT 1    /* pos1 < gapStart branch (Figbird.cpp:3124,3546): insert = x0 - mate + len */
#define FB_READ_REVERSE 2 /* isReverse[q] (Figbird.cpp:5727-5736): error-model index runs len-1-j */
#define FB_READ_NOMATE 4  /* partial read whose ref_pos is -1: no insert-size filter (Figbird.cpp:3132) */

/* ------------------------------------------------------------------------------------------------------
 * Work items.  One item = one candidate gap length of one gap, i.e. one `initialize(gapEstimate)` followed
 * by the EM rounds of GapFiller::run / the fillGap candidate loop (Figbird.cpp:5913-5965, 6298-6352), or one
 * hard placement sweep on a given string (finalize, Figbird.cpp:4929-5399).
 * ---------------------------------------------------------------------------------------------------- */
#define FB_ITEM_EM 0
#define FB_ITEM_HARD 1

#define FB_FLAG_EXTRA_PASS 1     /* after the EM loop run one more placeReads without M-step (Figbird.cpp:6348-6352) */
#define FB_FLAG_RECORD_ALL 2     /* keep per-read results of every placeReads call, not only the last one */
#define FB_FLAG_WANT_COUNTS 4    /* return countsGap gap rows of the last call */
#define FB_FLAG_RESUME 8         /* do not initialise from the pile-ups: start with the M-step on counts_in */
#define FB_FLAG_NO_COMP_STOP 16  /* ignore the comp_count>=5 stop rule (partial mode runs exactly max_rounds) */
#define FB_FLAG_FINALIZE_REF 32  /* HARD items: finalize()'s `ref_pos += gapoffset` for every right-side partial
                                    read, also those without a mate (Figbird.cpp:5295) */

typedef struct {
    int32_t kind;               /* FB_ITEM_EM / FB_ITEM_HARD */
    int32_t gap;                /* index into the uploaded batch */
    int32_t cand_len;           /* Lg: candidate gap length (gapEstimate) */
    int32_t max_rounds;         /* EM: number of placeReads+M-step rounds at most (num_itr=200, or 3 for partial) */
    int32_t flags;              /* FB_FLAG_* */
    int32_t comp_count_in;      /* RESUME: comp_count carried in */
    const double* counts_in;    /* RESUME: countsGap gap rows [Lg][5] to start the M-step from */
    const uint8_t* string_in;   /* HARD: the gap string (codes, [Lg]); EM: previous hard consensus ([Lg]) the first call is compared
                                   with (previous_str, Figbird.cpp:3919-3927), or NULL for none */
} FbWorkItem;

/* Result header; arrays follow in the same engine-owned pinned arena at the given byte offsets from the
 * header.  Valid until the next fb_em_run on the same context. */
typedef struct {
    int32_t calls;              /* placeReads calls executed (EM rounds + extra pass); HARD: 1 */
    int32_t comp_count;         /* comp_count after the last call (Figbird.cpp:3919-3927) */
    int32_t flags;              /* bit0 umaxleftf, bit1 umaxrightf, bit2 ucoverf (Figbird.cpp:3894-3911), OR over calls */
    int32_t n_reads;
    int32_t cand_len;
    int32_t n_slots;            /* recorded calls: `calls` with RECORD_ALL, else 1 (the last) */
    int64_t placements;         /* (read, admissible offset) pairs scored in pass 1, summed over calls */
    int64_t off_p1max;          /* double[n_slots][n_reads] largest pass-1 product per read; -1 = no admissible offset */
    int64_t off_p2max;          /* double[n_slots][n_reads] largest pass-2 product per read; -1 = none.  EM items: exact whenever the
                                   read is accepted (-log10(p) < prob_cutoff, Figbird.cpp:3474,3852); for a rejected read only
                                   "below the accept threshold" is guaranteed (the scan stops early).  HARD items: always exact */
    int64_t off_pos2;           /* int32 [n_slots][n_reads] x0 of the first offset reaching p2max (same validity as p2max) */
    int64_t off_soft;           /* uint8 [cand_len] computeSequence(0,0) codes after the last call (4 = N) */
    int64_t off_hard;           /* uint8 [cand_len] computeSequence(1,1) codes (unmapped mode) */
    int64_t off_cov;            /* int32 [cand_len] gap_coverage (unmapped mode) */
    int64_t off_counts;         /* double[cand_len][5] countsGap gap rows after the last call (WANT_COUNTS), else -1 */
} FbItemOut;

typedef struct {
    int64_t placements_p1;      /* pass-1 (read, offset) pairs scored since context creation */
    int64_t placements_p2;      /* pass-2 pairs */
    int64_t base_terms;         /* scored read bases (pass 1 + pass 2) */
    int64_t kernel_launches;
    double device_ms;           /* CUDA-event time of the kernels launched by fb_em_run */
    int64_t h2d_bytes, d2h_bytes;
    int64_t lane_steps_p1;      /* pass-1 gap-row terms actually executed (warp steps x 32 lanes): flank terms come from the
                                   per-gap cache and inadmissible offsets are never visited, so this is below base_terms */
    int64_t lane_steps_p2;      /* pass-2 terms actually executed (pruned against the pass-1 winner) */
    double device_union_ms;     /* union of the kernel intervals of ALL contexts of this process on this physical device
                                   (contexts that share a GPU overlap their kernels; sum of device_ms would double count) */
} FbCounters;

fb_status fb_ctx_create(int32_t device, fb_ctx** out);
void fb_ctx_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx);
const char* fb_engine_name(void);  /* "cuda-sm100a" for the product library */

/* Scheduling hint: a context whose requests are small and sequentially dependent (one EM round per call, the host decides
 * between rounds: the large-gap loop of Figbird.cpp:6323-6344 with the border update :4029-4376) asks for its kernels to be
 * scheduled ahead of the bulk launches of other contexts on the same GPU (CUDA stream priority).  Results are unaffected. */
fb_status fb_ctx_set_latency_critical(fb_ctx* ctx, int32_t on);

fb_status fb_model_upload(fb_ctx* ctx, const FbModel* model);
fb_status fb_batch_upload(fb_ctx* ctx, const FbGapBatch* batch);

/* Runs n items; out[i] points at the i-th result header. */
fb_status fb_em_run(fb_ctx* ctx, const FbWorkItem* items, int32_t n_items, const FbItemOut** out);

fb_status fb_get_counters(const fb_ctx* ctx, FbCounters* out);

/* Diagnostic for the roofline report: measured FP64 throughput of the device.  out2[0] = DMUL-only rate
 * (1e12 instr/s == TFLOP/s at 1 flop per instruction, the ceiling when FMA contraction is forbidden),
 * out2[1] = DFMA rate in TFLOP/s (2 flop per instruction). */
fb_status fb_microbench_fp64(fb_ctx* ctx, double* out2);

/* Level 1: the FillGaps executable as a function.  argv[1..15] as FillGaps.cpp:419-433.  Returns the
 * process exit status the reference would give (0 ok, 1 on unreadable inputs).  GPUs: devices listed in
 * FIGBIRD_GPUS (default: device 0); gaps are sharded cost-balanced across them. */
int32_t fb_fillgaps_main(int32_t argc, const char* const* argv);

/* ------------------------------------------------------------------------------------------------------
 * Level 1, the callers and consumers on either side of FillGaps: each is the reference
 * executable of the same name as a function -- argv[0] is ignored, argv[1..] are that program's positional
 * arguments, the files read and written are the same, the return value is its exit status.  Host-only.
 * ---------------------------------------------------------------------------------------------------- */
/* Preprocess.cpp:1832-2676 (RunFigbird.sh:285,338,451,472).  argv[1..13]: contig file, maxDistance, mode
 * (1 partial / 2 unmapped), bowtie2 SAM, myout.sam to write, gapped genome, reads_1, reads_2, Gaps dir/,
 * Temp dir/, default_setting, genome_reduction, read_reduction.  Writes gapInfo.txt, stat.txt, stat2.txt,
 * myout.sam and Gaps/gaps_<g>.sam or Gaps/partial_gaps_<g>.sam (and the reduced read files when asked). */
int32_t fb_preprocess_main(int32_t argc, const char* const* argv);
/* CombineGaps.cpp:169-313 (RunFigbird.sh:777).  argv[1..2]: number of iterations, directory/ holding
 * gapout_<itr>.txt.  Writes combined_gapstring.txt and Individual_gaps.txt there. */
int32_t fb_combinegaps_main(int32_t argc, const char* const* argv);
/* FlankTrim.cpp:22-233 (RunFigbird.sh:254,433).  argv[1..4]: gapped genome, trim, read length, output FASTA. */
int32_t fb_flanktrim_main(int32_t argc, const char* const* argv);
/* Reduce_SCF.cpp:16-152 (RunFigbird.sh:266,320).  argv[1..2]: gapped genome, Temp dir/ (-> newgenome.fa). */
int32_t fb_reduce_scf_main(int32_t argc, const char* const* argv);
/* Reverse.cpp:42-120 (RunFigbird.sh:166).  argv[1..2]: the two FASTQ files of a jump library; writes
 * <stem>_reversed<ext> beside them and prints the two new paths. */
int32_t fb_reverse_main(int32_t argc, const char* const* argv);

#ifdef __cplusplus
}
#endif
#endif
