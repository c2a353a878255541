/*
 * mipgen_b200.h -- C-ABI of the B200-native MIPgen scoring hot path.
 *
 * One shared library (libmipgen_b200.so), plain pointers and sizes, no C++ or
 * torch types.  Every compute entry point runs hand-written sm_100a CUDA
 * kernels; there is NO CPU fallback -- a missing/failed device is an error
 * (negative status + mg_last_error()).
 *
 * What each entry point replaces in the reference (/root/reference):
 *
 *   mg_long_range_content   Featurev5::get_long_range_content      Featurev5.cpp:18-56
 *   mg_load_svr_model       svm_load_model                         svm.cpp:2759-2973  (mipgen.cpp:409)
 *   mg_svr_predict          svm_predict / svm_predict_values /     svm.cpp:2580-2593, 2504-2522,
 *                           Kernel::k_function (RBF)               328-368            (mipgen.cpp:2016)
 *   mg_score_candidates     SVMipv4::get_score, ::get_parameters   SVMipv4.cpp:114-248, 60-113
 *                           (+ predict_value)                      mipgen.cpp:1948-2019
 *   mg_score_regions        the candidate loop nest of             mipgen.cpp:421-501
 *   mg_panel_*              mipgen::tile_regions + design_mip      mipgen.cpp:599-613
 *                           (Plus/Minus geometry and setters)      PlusSVMipv4.cpp:7-28, MinusSVMipv4.cpp:6-51
 *   mg_tile_replay          the score-dependent skips of the loop  mipgen.cpp:426-437, 494-497
 *   mg_describe_candidates, mg_format_mip_record   print_details   mipgen.cpp:765-794
 *
 * Threading: a context is used by one host thread at a time; calls are
 * synchronous unless stated.  The library never calls rand()/srand() and never
 * changes the locale.
 */
#ifndef MIPGEN_B200_H
#define MIPGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_NFEAT 192 /* SVMipv4.cpp:13 TOTAL_FEATURES */
#define MG_NLRC 44   /* Featurev5.h:4 MER_NUM */

/* status codes */
#define MG_OK 0
#define MG_ERR_CUDA (-1)    /* CUDA runtime / launch failure, or no usable device */
#define MG_ERR_INVALID (-2) /* bad argument / inconsistent input */
#define MG_ERR_MODEL (-3)   /* model file unreadable or not an RBF (epsilon|nu)-SVR */
#define MG_ERR_NOMODEL (-4) /* SVR requested before a model was loaded */
#define MG_ERR_NOCONFIG (-5)/* grid call before mg_set_config */
#define MG_ERR_NOMEM (-6)
#define MG_ERR_UNSUPPORTED (-7) /* the device formatter leaves this input to the host one (score magnitude, record length) */

/* what to compute (bit mask) */
#define MG_WANT_LOGISTIC 1 /* SVMipv4::get_score                        */
#define MG_WANT_SVR 2      /* get_parameters + svm_predict               */
#define MG_WANT_FEATURES 4 /* the 192 doubles of get_parameters themselves */

typedef struct mg_ctx mg_ctx;
typedef struct mg_panel mg_panel;

/* The knobs of mipgen.cpp that shape the candidate grid (parse_arg_values, 190-276). */
typedef struct {
    int max_capture;       /* -max_capture_size                                   */
    int min_capture;       /* -min_capture_size                                   */
    int capture_increment; /* -capture_increment (0 is treated as 1, mipgen.cpp:274) */
    int max_mip_overlap;   /* -max_mip_overlap, used by the static skip at :429   */
    int n_pairs;           /* arm pairs in ENUMERATION order: arm sum descending,  */
    const int *ext_len;    /*   extension length in list order within a sum       */
    const int *lig_len;    /*   (mipgen.cpp:431, 438; lists built at 222-261)     */
    int n_oligo_sizes;     /* set<int> oligo_sizes (mipgen.cpp:77); may be 0      */
    const int *oligo_sizes;
} mg_config;

/* One Featurev5 (Featurev5.h:10-23).  Coordinates are 1-based inclusive. */
typedef struct {
    const char *seq;   /* chromosomal_sequence (upper-cased by the caller, may hold N/IUPAC/-) */
    int seq_len;       /* must equal seq_stop - seq_start + 1                                 */
    int seq_start;     /* chromosomal_sequence_start_position                                 */
    int seq_stop;      /* chromosomal_sequence_stop_position                                  */
    int start_flanked; /* start_position_flanked                                              */
    int stop_flanked;  /* stop_position_flanked                                               */
    const double *lrc; /* long_range_content[44], or NULL (logistic only / zeros)            */
    const int *copies; /* [n_oligo_sizes][seq_len]: copy number of the oligo of size
                          oligo_sizes[k] starting at seq index i (copy_chr_start_stop,
                          mipgen.cpp:83, 612-613; 0 = absent key); NULL => every copy is 1   */
    int scan_begin;    /* optional explicit range of scan starts [scan_begin, scan_end];      */
    int scan_end;      /*   both 0 => the reference's rule (mipgen.cpp:421-425)               */
    /* Selection-only inputs (they change no score, only condense/collapse): NULL = absent.              */
    const char *masked_seq;     /* masked_chromosomal_sequence (TRF output, mipgen.cpp:1045-1084): seq_len
                                   characters, 'N' = masked (mipgen.cpp:606-610); NULL => seq itself,
                                   which is what the reference uses with -trf off (mipgen.cpp:1058-1062) */
    const uint8_t *snp;         /* [seq_len] non-zero where chr_snp_positions holds seq_start + i
                                   (mipgen.cpp:634-636, 698-700): feeds snp_count                        */
    const uint8_t *unmappable;  /* [n_capture_sizes][seq_len], capture index as in the grid: non-zero where
                                   unmappable_positions[capture][chr] holds seq_start + i as a MIP start
                                   (mipgen.cpp:615-625); NULL also stands for -check_copy_number off     */
} mg_region;

/* One SVMipv4 object as mipgen.cpp leaves it after design_mip: strand-oriented strings
 * (already reverse-complemented on '-'), constructor lengths and copy numbers. */
typedef struct {
    const char *ext; int ext_n; /* ext_probe_sequence   */
    const char *lig; int lig_n; /* lig_probe_sequence   */
    const char *tgt; int tgt_n; /* scan_target_sequence */
    int ext_len, lig_len, scan_size;
    int ext_copy, lig_copy;
} mg_candidate;

/* Per-kernel device timings (CUDA events on the launching stream) accumulated since
 * the last mg_reset_timings().  ms_* are sums over launches. */
typedef struct {
    double ms_feat;  long launches_feat;   /* K-feat (+ fused logistic epilogue)      */
    double ms_svr;   long launches_svr;    /* K-svr  (FP64 DMMA contraction + exp)     */
    double ms_other; long launches_other;  /* encode / lrc / misc                      */
    long candidates_feat, candidates_svr;  /* units the timed launches processed       */
    /* work actually issued by the K-svr launches (either form), for the roofline:          */
    double svr_dmma;  /* DMMA.8x8x4 warp instructions (256 FP64 FMA each)                    */
    double svr_exp;   /* kernel values exp(-gamma d) evaluated                               */
    double svr_gather;/* factor triples multiplied and accumulated (factored form only)      */
    double svr_tc_mma;/* tcgen05.mma instructions (M128 N64 K16, FP16 in / FP32 out) of the tensor-core form */
} mg_timings;

/* ---- context ---------------------------------------------------------------- */
const char *mg_version(void);
/* Create a context on CUDA device `device`.  Fails (MG_ERR_CUDA) without a GPU. */
int mg_create(int device, mg_ctx **out);
void mg_destroy(mg_ctx *ctx);
const char *mg_last_error(const mg_ctx *ctx); /* ctx may be NULL: last create error */
int mg_set_config(mg_ctx *ctx, const mg_config *cfg);
int mg_sync(mg_ctx *ctx);
/* Device stopwatch on the context's stream: start records a CUDA event, stop records a
 * second one, waits for it and returns the elapsed device time in milliseconds. */
int mg_timer_start(mg_ctx *ctx);
int mg_timer_stop(mg_ctx *ctx, double *ms);
int mg_reset_timings(mg_ctx *ctx);
int mg_get_timings(mg_ctx *ctx, mg_timings *out); /* synchronises the stream */

/* ---- SVR model ---------------------------------------------------------------- */
/* Parse a libsvm text model (svm.cpp:2759-2973) and upload it as a dense SV matrix. */
int mg_load_svr_model(mg_ctx *ctx, const char *path);
/* Same, from memory: sv is [n_sv][n_feat] row-major (n_feat <= 192). */
int mg_set_svr_model(mg_ctx *ctx, const double *sv, const double *alpha, int n_sv, int n_feat,
                     double gamma, double rho);
int mg_model_info(const mg_ctx *ctx, int *n_sv, double *gamma, double *rho);
/* How region grids are SVR-scored: 0 = automatic (the factored kernel when the arm-pair table
 * fits its shared-memory tables, else the dense contraction), 1 = dense DMMA contraction,
 * 2 = factored (error if it does not fit).  Those compute the same FP64 decision value.
 * 3 = the tcgen05 tensor-core form: split-FP16 contraction with FP32 TMEM accumulators, FP64 exponent /
 * exp / row sum (k_svr_tc.cu); ~1e-9 relative to libsvm instead of ~1e-13, see DESIGN.md.  Error if the
 * model's length / junction columns are not small integers (mg_svr_tensor_core_available). */
int mg_set_svr_mode(mg_ctx *ctx, int mode);
int mg_svr_tensor_core_available(const mg_ctx *ctx);
/* > 0 (the factored kernel's window size) if the current config fits the factored kernel */
int mg_svr_factored_available(const mg_ctx *ctx);
/* svm_predict for n dense rows x[i*ld .. i*ld+191] (features 1..192). out[n]. */
int mg_svr_predict(mg_ctx *ctx, const double *x, long n, long ld, double *out);
/* Same arithmetic order as libsvm (direct (x-s)^2 form, sequential sums), one thread
 * per row: the slow on-device cross-check of the contraction kernel. */
int mg_svr_predict_direct(mg_ctx *ctx, const double *x, long n, long ld, double *out);

/* ---- per-region constants -------------------------------------------------------- */
/* long_range_content of one region: ext_seq is the region +- (max_capture+1000) window,
 * denom = seq_stop - seq_start + 2001 (Featurev5.cpp:49,53). */
int mg_long_range_content(mg_ctx *ctx, const char *ext_seq, int n, int denom, double out[MG_NLRC]);

/* ---- explicit candidates (the SVMipv4 interface) --------------------------------- */
/* lrc: [n][44] (one row per candidate) or NULL.  Any output may be NULL when not wanted.
 * logistic[n], svr[n], features[n][192]. */
int mg_score_candidates(mg_ctx *ctx, const mg_candidate *cands, long n, const double *lrc, int want,
                        double *logistic, double *svr, double *features);

/* ---- region grids (the tile_regions loop nest) ------------------------------------- */
/* Grid of one region, canonical (= reference enumeration) order:
 *   index = (((scan_idx*n_cap + cap_idx)*n_pairs + pair_idx)*2 + strand)      strand 0 '+', 1 '-'
 *   scan_start = first_scan_start + scan_idx, capture = max_capture - cap_idx*increment  */
int64_t mg_grid_size(const mg_ctx *ctx, const mg_region *r);
int mg_first_scan_start(const mg_ctx *ctx, const mg_region *r);
/* the same two, from a bare config (pure host arithmetic, no device needed) */
int64_t mg_config_grid_size(const mg_config *cfg, const mg_region *r);
int mg_config_first_scan_start(const mg_config *cfg, const mg_region *r);

/* Host-buffer call: upload regions, score every grid point on the device, copy back.
 * Region i's grid starts at offsets[i] (offsets may be NULL; out_offsets, if given,
 * receives n+1 prefix sums).  valid[i]=0 marks points removed by the static skips
 * (mipgen.cpp:429, 443-444); their scores are NaN. */
int mg_score_regions(mg_ctx *ctx, const mg_region *regions, int n, int want, int64_t *out_offsets,
                     uint8_t *valid, double *logistic, double *svr, double *features);

/* Device-resident variant: inputs are uploaded once, results stay in HBM. */
int mg_panel_create(mg_ctx *ctx, const mg_region *regions, int n, mg_panel **out);
void mg_panel_destroy(mg_panel *p);
int64_t mg_panel_candidates(const mg_panel *p);
int64_t mg_panel_valid_candidates(const mg_panel *p); /* statically valid grid points */
/* Bytes of arm / insert row tables K-feat hands to the factored SVR per scoring pass (the distinct rows of every work item:
 * 192 B per arm row, 704 B per insert row, 12 B of norm + junction code each); 0 when the factored form does not apply.
 * This, not 1,536 B per candidate, is what the SVR path moves through HBM between its two kernels. */
int64_t mg_panel_row_table_bytes(const mg_panel *p);
/* Launch the kernels for the whole panel on the context's stream (asynchronous). */
int mg_panel_score(mg_ctx *ctx, mg_panel *p, int want);
/* Copy results back (synchronous).  features may be NULL. */
int mg_panel_fetch(mg_ctx *ctx, mg_panel *p, uint8_t *valid, double *logistic, double *svr, double *features);
/* Raw device pointers of the result arrays (NULL if never computed). */
int mg_panel_device_ptrs(const mg_panel *p, const uint8_t **valid, const double **logistic, const double **svr);

/* ---- selection front-end on the device ----------------------------- */
/* condense_mips (mipgen.cpp:1670-1746, on top of the tile loop's score-dependent enumeration,
 * mipgen.cpp:426-497) and collapse_mips (mipgen.cpp:1617-1649) over a scored panel: the best candidate
 * per (scan start, strand) and per (position, strand), as global grid indices of the panel (-1: none).
 * Arm copies come from the copy tables; arm_fraction_masked, snp_count and mapping_failed (mipgen.cpp:606-625,
 * 634-760) from the regions' masked_seq / snp / unmappable inputs. */
typedef struct {
    int method;                 /* 0 logistic, 1 svr, 2 mixed (tile phase = logistic scores)              */
    int heuristic;              /* -logistic_heuristic != "off"                                          */
    double lower_score_limit;   /* -{svr,logistic}_priority_score                                        */
    double upper_score_limit;   /* -{svr,logistic}_optimal_score                                         */
    int max_arm_copy;           /* -max_arm_copy_product (75)                                            */
    int target_arm_copy;        /* -target_arm_copy (20)                                                 */
    double masked_arm_threshold;/* -masked_arm_threshold (0.5): mipgen.cpp:1629, 1701                        */
} mg_select_params;
/* number of scan starts / coverable positions of a region (positions: first scan start ..
 * stop_flanked + max_capture - min arm sum - 1) */
int mg_region_scan_count(const mg_ctx *ctx, const mg_region *r);
int mg_region_position_count(const mg_ctx *ctx, const mg_region *r);
/* scan_best[(scan_offset(region) + scan_idx)*2 + strand], pos_best[(pos_offset(region) + pos_idx)*2 + strand];
 * offsets are prefix sums of the two counts over the panel's regions.  The panel must have been scored
 * with the method's scores (MG_WANT_SVR for method 1, MG_WANT_LOGISTIC otherwise). */
int mg_panel_select(mg_ctx *ctx, mg_panel *p, const mg_select_params *sp, int64_t *scan_best, int64_t *pos_best);

/* ---- host helper: replay of the score-dependent control flow ---------------------- */
/* Walks one scored region grid exactly like the tile loop (mipgen.cpp:426-497) and
 * writes the grid indices the reference would have enumerated.  method: 0 logistic,
 * 1 svr, 2 mixed; heuristic: -logistic_heuristic != "off".  Returns the count, or -1 on
 * a bad config.  Pure host logic: needs no context and no device. */
int64_t mg_tile_replay(const mg_config *cfg, const mg_region *r, const uint8_t *valid, const double *score,
                       int method, int heuristic, double upper_score_limit, int64_t *out_idx, int64_t cap);

/* ---- design-file records for the batched caller ------------------------------
 * A caller that owns the loop (INTEGRATION.md route B) turns grid indices -- the enumeration from
 * mg_tile_replay, the winners from mg_panel_select -- back into what the reference keeps per SVMipv4 object and
 * prints with print_details (mipgen.cpp:765-794).  Host arithmetic only. */
typedef struct mg_mip_info {
    int strand;                     /* 0 '+', 1 '-' */
    int ext_len, lig_len;           /* extension_arm_length, ligation_arm_length */
    int scan_start, scan_stop;      /* scan_start_position, scan_stop_position (chromosomal, 1-based) */
    int ext_start, ext_stop;        /* ext_probe_start/stop  (PlusSVMipv4.cpp:7-14, MinusSVMipv4.cpp:30-37) */
    int lig_start, lig_stop;        /* lig_probe_start/stop */
    int ext_copy, lig_copy;         /* ext/lig_probe_copy (mipgen.cpp:612-613); 1 without a copy table */
} mg_mip_info;

/* Geometry and arm copy numbers of n candidates of one region, given their grid indices.  Returns MG_OK or
 * MG_ERR_INVALID (bad config / index outside the region's grid). */
int mg_describe_candidates(const mg_config *cfg, const mg_region *r, const int64_t *idx, int n, mg_mip_info *out);

/* One record exactly as print_details writes it to all_mips.txt / collapsed_mips.txt (mipgen.cpp:765-794):
 * key, score (ostream default = %g), chr, arm coordinates / copies / sequences (reverse-complemented on '-',
 * MinusSVMipv4.cpp:38-51), scan target, mip_seq = lig + universal_middle + ext (mipgen.cpp:605), feature
 * start - 1 and stop, strand, failure flags "000" (no SNP / TRF / mapping inputs), label_%04d.
 * universal_middle is mipgen.cpp:199-200's string (lig tag Ns + constant + ext tag Ns).
 * Writes at most cap bytes incl. the terminating NUL; returns the record's length (without NUL), or -1. */
int64_t mg_format_mip_record(const mg_region *r, const mg_mip_info *m, double score, const char *chr, const char *label,
                             int feature_start, int feature_stop, const char *universal_middle, int mip_index,
                             char *buf, int64_t cap);

/* The same for n candidates of one region at once (grid indices idx, scores taken from the region's score grid),
 * numbered first_index, first_index + 1, ...: what output_collapsed_mips (mipgen.cpp:1651-1668) or the all_mips
 * writer (mipgen.cpp:474, 488) appends for one feature.  Returns the bytes written (no NUL is appended), or -1 if
 * cap is too small or an argument is invalid; 512 + 3 * max_capture bytes per record are always enough. */
int64_t mg_format_mip_records(const mg_config *cfg, const mg_region *r, const int64_t *idx, int n, const double *score,
                              const char *chr, const char *label, int feature_start, int feature_stop,
                              const char *universal_middle, int first_index, char *buf, int64_t cap);

/* ---- the same records written on the device -------------------------------------------------------------
 * n records of a scored panel, given by panel-global grid indices, numbered first_index, first_index + 1, ...; byte-identical to
 * what mg_format_mip_records / print_details (mipgen.cpp:765-794) produce, incl. the score as `ostream << double` prints it (%g,
 * correctly rounded from the exact binary value).  which: 0 = logistic scores, 1 = SVR scores.  meta[i] describes region i of the
 * panel.  Returns the bytes written into buf (host memory, capacity cap), or a negative status: regions that carry TRF / SNP /
 * mappability inputs are refused (their flags need design_mip). */
typedef struct {
    const char *chr;     /* Featurev5::chr   */
    const char *label;   /* Featurev5::label */
    int feature_start;   /* Featurev5::start_position (printed minus 1) */
    int feature_stop;    /* Featurev5::stop_position  */
} mg_record_meta;
int64_t mg_panel_format_records(mg_ctx *ctx, mg_panel *p, const mg_record_meta *meta, const int64_t *idx, int64_t n, int which,
                                const char *universal_middle, int first_index, char *buf, int64_t cap);
/* all_mips.txt of a scored panel, written on the device: the records of every candidate the tile loop enumerates, in its order
 * (mipgen.cpp:426-497 with the score-dependent shortcuts of lines 430, 434, 494 replayed per scan start by K-replay; sp->method,
 * sp->heuristic and sp->upper_score_limit drive them, mixed mode prints logistic scores as mipgen.cpp:467-468 does), numbered
 * first_index, first_index + 1, ...  The text is handed to `sink` in consecutive pieces (return non-zero from it to abort);
 * records_per_region[n_regions] (may be NULL) receives each region's record count.  Returns the bytes produced or a negative
 * status (MG_ERR_UNSUPPORTED: see above; regions with TRF / SNP / mappability inputs are refused like mg_panel_format_records). */
typedef int (*mg_text_sink)(void *user, const char *text, int64_t len);
int64_t mg_panel_format_enumerated(mg_ctx *ctx, mg_panel *p, const mg_record_meta *meta, const mg_select_params *sp,
                                   const char *universal_middle, int first_index, int64_t *records_per_region, mg_text_sink sink, void *user);
/* printf("%g") of n doubles on the device, 32 bytes per value in out32 (not NUL-terminated), lengths in len (-1: magnitude outside
 * [1e-12, 1e15), which the record formatter leaves to the host).  Exposed for tests. */
int mg_format_g(mg_ctx *ctx, const double *values, int64_t n, char *out32, int *len);

/* ---- the FASTQ files check_copy_numbers writes for BWA (mipgen.cpp:798-840), formatted on the device --------------------------------
 * all_sequences.fq: one read per (region, capture size descending, MIP start ascending) "@<capture>_<chr>_<start>" (mipgen.cpp:808-823);
 * oligo_copy_count.fq: one read per (region, oligo size in the given order, start ascending) "@chr<chr>:<a>-<b>" (mipgen.cpp:824-836).
 * chr[i] names region i's chromosome.  With buf == NULL the calls return the number of bytes needed (host arithmetic); otherwise the
 * bytes written, or a negative status. */
int64_t mg_format_capture_fastq(mg_ctx *ctx, const mg_region *regions, const char *const *chr, int n, int max_capture, int min_capture,
                                int capture_increment, char *buf, int64_t cap);
int64_t mg_format_oligo_fastq(mg_ctx *ctx, const mg_region *regions, const char *const *chr, int n, const int *oligo_sizes, int n_oligo_sizes,
                              char *buf, int64_t cap);

/* ---- one call per batch of Featurev5 objects: what tile_regions does per feature up to collapse_mips ----------
 * (mipgen.cpp:412-505: the candidate loop nest, condense_mips, collapse_mips), for a caller that keeps pick_mips
 * (mipgen.cpp:1506-1614) on the host.  Regions are walked in sub-batches of at most max_batch_candidates grid
 * points (0 = a default of 2^26), so device memory stays bounded however long the region list is; only the
 * winners (and, if asked for, the full grids) come back.  Indices are LOCAL to each region's grid
 * (mg_describe_candidates / mg_format_mip_records take them as they are), -1 = none. */
typedef struct {
    int64_t *scan_best;          /* [2 * scan_off[n]]  scan_strand_best_mip: entry (scan_off[i] + scan_idx)*2 + strand   */
    int64_t *pos_best;           /* [2 * pos_off[n]]   pos_strand_best_mip:  entry (pos_off[i] + pos_idx)*2 + strand     */
    double *scan_best_logistic;  /* [2 * scan_off[n]]  logistic score of each scan_best winner (NaN where -1); or NULL   */
    double *scan_best_svr;       /* [2 * scan_off[n]]  SVR score of each scan_best winner; or NULL                       */
    uint8_t *valid;              /* [grid_off[n]] full grids in region order; each may be NULL                           */
    double *logistic;
    double *svr;
} mg_tile_result;
/* Prefix sums a caller needs to size the arrays of mg_tile_result -- pure host arithmetic: grid points, scan starts and coverable
 * positions of regions[0..n); each array has n + 1 entries and may be NULL. */
int mg_tile_sizes(const mg_config *cfg, const mg_region *regions, int n, int64_t *grid_off, int64_t *scan_off, int64_t *pos_off);
/* want: MG_WANT_LOGISTIC and/or MG_WANT_SVR (what to score); sp selects on sp->method's scores (mixed = logistic,
 * mipgen.cpp:467-468) and may be NULL (score only: scan_best / pos_best are not written). */
int mg_tile_regions(mg_ctx *ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                    int64_t max_batch_candidates, mg_tile_result *out);

/* mg_tile_regions that also writes all_mips.txt on the device: meta[i] describes region i; the text of the sub-batches reaches
 * the sink in region order while the winners are being computed (one context: record numbers run through the whole list). */
int mg_tile_regions_records(mg_ctx *ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                            int64_t max_batch_candidates, mg_tile_result *out, const mg_record_meta *meta, const char *universal_middle,
                            int first_index, int64_t *records_per_region, mg_text_sink sink, void *user);

/* ---- the same over several GPUs of one box ---------------------------------------------------------
 * Regions are independent (mipgen.cpp:412-525), so they are partitioned over the contexts by longest-processing-time
 * on their grid sizes; every context runs on its own host thread, stream and device and writes its regions' slices
 * of the caller's arrays, which therefore come back in the caller's region order.  No collective, no NCCL.  All
 * contexts must carry the same config and model.  Returns the first error of any context. */
int mg_tile_regions_multi(mg_ctx *const *ctxs, int n_ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                          int64_t max_batch_candidates, mg_tile_result *out);
/* mg_score_regions over several GPUs: out_offsets[n+1] (may be NULL), valid / logistic / svr as in mg_score_regions. */
int mg_score_regions_multi(mg_ctx *const *ctxs, int n_ctx, const mg_region *regions, int n, int want, int64_t *out_offsets,
                           uint8_t *valid, double *logistic, double *svr);
/* the partition both calls use: owner[i] = index of the context that scores region i */
int mg_partition_regions(const mg_config *cfg, const mg_region *regions, int n, int n_parts, int *owner);

/* scores of n grid points of a scored panel (panel-global indices; -1 yields NaN); logistic / svr may be NULL */
int mg_panel_gather(mg_ctx *ctx, mg_panel *p, const int64_t *idx, int64_t n, double *logistic, double *svr);

/* ---- SURVEY.md 8(f4), OPT-IN: exact-match arm copy counting on the device -------------------------------------------------------
 * Replaces `bwa aln` + `bwa samse` on <project>.oligo_copy_count.fq and find_copy's parse of the X0 tag (mipgen.cpp:558-596, reads
 * written at 824-836) for what X0 is on reads cut out of the indexed genome: the number of EXACT occurrences of the oligo on either
 * strand.  Nothing in the library calls it by itself: it changes the contract of an external tool (BWA also reports hits with
 * mismatches when a read has no exact hit), so a caller opts in by filling mg_region.copies from it (route C).
 *
 * mg_genome_create   index of the given contigs (plain sequences, any case; every character other than ACGT separates): the sorted
 *                    2-bit packed 32-mers of all positions, resident in HBM (8 bytes per base).
 * mg_count_arm_copies  for every region and oligo size (1..32 bases) a table in the layout of mg_region.copies, [n_oligo_sizes][seq_len]:
 *                    entry (k, i) = occurrences of seq[i .. i + size_k - 1] plus occurrences of its reverse complement, as find_copy
 *                    stores X0; 100 (find_copy's value for a read without X0) when the oligo holds a character other than ACGT or
 *                    does not occur in the index; 0 = absent key for the starts the reference never writes (i >= seq_len - size_k,
 *                    mipgen.cpp:829).  copies_off[n + 1] (may be NULL) receives each region's offset in `copies`; with
 *                    copies == NULL only the offsets are computed. */
typedef struct mg_genome mg_genome;
int mg_genome_create(mg_ctx *ctx, const char *const *seqs, const int64_t *lens, int n_contigs, mg_genome **out);
void mg_genome_destroy(mg_genome *g);
int mg_genome_info(const mg_genome *g, int64_t *positions, int64_t *indexed, int64_t *short_suffixes);
int mg_count_arm_copies(mg_genome *g, const mg_region *regions, int n, const int *oligo_sizes, int n_oligo_sizes, int32_t *copies,
                        int64_t *copies_off);

#ifdef __cplusplus
}
#endif
#endif /* MIPGEN_B200_H */
