/*
 * mipgen_oracle.h -- CPU restatement of the MIPgen scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (mipgen_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg use it, as the checker.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_*.py)
 * against the compiled, unmodified reference objects (oracle/_ref, built by
 * oracle/Makefile from /root/reference) and against golden vectors generated
 * from them (tests/golden/, generator tools/make_golden.py).  The reference
 * itself ships no tests or golden vectors.
 *
 * Each function cites the reference file:line it restates.
 */
#ifndef MIPGEN_ORACLE_H
#define MIPGEN_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NFEAT 192
#define ORC_NLRC 44

/* ---- sequence helpers ---------------------------------------------------- */
/* MinusSVMipv4.cpp:6-29 : reverse, complement ACGT, pass other chars through */
void orc_reverse_comp(const char *in, int n, char *out);
/* SVMipv4.cpp:31-57 : overlapping occurrence count (find(sub, offset+1) loop) */
int orc_count_mer(const char *s, int n, const char *sub, int k);

/* ---- per-region long-range content -------------------------------------- */
/* Featurev5.cpp:18-56 with the k-mer list of mipgen.cpp:32.
 * denom = chromosomal_sequence_stop_position - chromosomal_sequence_start_position + 2001 */
void orc_long_range_content(const char *ext_seq, int n, int denom, double out[ORC_NLRC]);

/* ---- one candidate, given its (strand-oriented) strings ------------------ */
typedef struct {
    const char *ext;  int ext_n;    /* ext_probe_sequence  (already rev-comped on '-') */
    const char *lig;  int lig_n;    /* lig_probe_sequence */
    const char *tgt;  int tgt_n;    /* scan_target_sequence */
    int ext_len, lig_len, scan_size;/* ctor values (SVMipv4.cpp:16-30) */
    int ext_copy, lig_copy;         /* mipgen.cpp:612-613 */
} orc_mip;

/* SVMipv4.cpp:63 / 116 : 'N' inside an arm or '-' in mip_seq (== '-' in an arm) */
int orc_mip_invalid(const orc_mip *m);
/* SVMipv4.cpp:60-113 */
void orc_get_parameters(const orc_mip *m, const double lrc[ORC_NLRC], double out[ORC_NFEAT]);
/* SVMipv4.cpp:114-248 (+ junction table 249-267) */
double orc_get_score(const orc_mip *m);

/* ---- candidate geometry + design --------------------------------------- */
/* PlusSVMipv4.cpp:7-14, MinusSVMipv4.cpp:30-37 */
typedef struct {
    int scan_start, scan_stop, ext_len, lig_len, strand; /* strand 0 '+', 1 '-' */
    int ext_start, ext_stop, lig_start, lig_stop;
} orc_geom;
void orc_geometry(int scan_start, int scan_stop, int ext_len, int lig_len, int strand, orc_geom *g);

/* mipgen.cpp:461-462 + 602-603 (+ Minus setters): cut the three strings out of
 * the region's chromosomal_sequence.  Buffers must hold >= 512 chars.
 * Returns 0, or -1 if a substr would start beyond the sequence (reference throws). */
int orc_design(const char *seq, int seq_len, int seq_start, const orc_geom *g,
               char *ext, char *lig, char *tgt, orc_mip *m);

/* ---- libsvm subset ------------------------------------------------------- */
typedef struct orc_model orc_model;
/* svm.cpp:2759-2973 (text model; sparse idx:val).  NULL on failure. */
orc_model *orc_svm_load_model(const char *path);
void orc_svm_free(orc_model *m);
int orc_svm_nsv(const orc_model *m);
double orc_svm_gamma(const orc_model *m);
double orc_svm_rho(const orc_model *m);
/* svm.cpp:2580-2593 -> 2504-2522 -> 328-368, x dense with 1-based indices 1..n */
double orc_svm_predict(const orc_model *m, const double *x, int n);
/* mipgen.cpp:1948-2019: the "%.17g" text round trip in front of svm_predict */
double orc_predict_value(const orc_model *m, const double *x, int n);

/* ---- region grid + tile replay ------------------------------------------ */
typedef struct {
    const char *seq; int seq_len;
    int seq_start, seq_stop;          /* chromosomal_sequence_{start,stop}_position */
    int start_flanked, stop_flanked;  /* Featurev5 start/stop_position_flanked */
    const double *lrc;                /* [44] or NULL */
    const int *copies;                /* [n_oligo_sizes][seq_len] copy of oligo starting at
                                         seq index i with size oligo_sizes[k]; NULL => 1 */
} orc_region;

typedef struct {
    int max_capture, min_capture, capture_increment, max_mip_overlap;
    int n_pairs;                /* arm pairs in enumeration order:            */
    const int *ext_len;         /*   arm sum descending, ext ascending within */
    const int *lig_len;         /*   (mipgen.cpp:431,438; 249-259)            */
    int n_oligo_sizes; const int *oligo_sizes;
} orc_cfg;

/* number of capture sizes and scan starts (mipgen.cpp:421-427) */
int orc_n_captures(const orc_cfg *c);
int orc_first_scan_start(const orc_region *r, const orc_cfg *c);
int orc_n_scan(const orc_region *r, const orc_cfg *c);
long orc_grid_size(const orc_region *r, const orc_cfg *c);

/* Score the full static grid of a region in canonical order
 * index = (((scan_idx*n_cap + cap_idx)*n_pairs + pair_idx)*2 + strand)
 * valid[i]=0 for points removed by the static skips mipgen.cpp:429,443,444.
 * Any of logistic / svr / feats may be NULL.  feats is [grid][192]. */
void orc_grid_region(const orc_region *r, const orc_cfg *c, const orc_model *model,
                     unsigned char *valid, double *logistic, double *svr, double *feats);

/* Replay the score-dependent control flow of the tile loop (mipgen.cpp:426-497)
 * over a scored grid: writes the grid indices the reference would actually
 * enumerate, in order, to out_idx (capacity cap); returns the count.
 * method: 0 logistic, 1 svr, 2 mixed.  heuristic: -logistic_heuristic != "off". */
long orc_tile_replay(const orc_region *r, const orc_cfg *c, const unsigned char *valid,
                     const double *score, int method, int heuristic, double upper_score_limit,
                     long *out_idx, long cap);

/* ---- selection front-end: best MIP per scan start / per position ---------- */
/* The knobs of condense_mips / collapse_mips (mipgen.cpp:197-198, 264-265, 177) and the selection-only inputs
 * design_mip reads besides the sequence (mipgen.cpp:606-625, 634-760):
 *   masked_seq   masked_chromosomal_sequence (seq_len chars; NULL => the sequence itself, as with -trf off,
 *                mipgen.cpp:1058-1062)                                   -> arm_fraction_masked
 *   snp          [seq_len] non-zero where chr_snp_positions has an entry  -> snp_count
 *   unmappable   [n_captures][seq_len] non-zero where unmappable_positions[capture][chr] holds that MIP start
 *                (NULL also stands for -check_copy_number off)            -> mapping_failed */
typedef struct {
    double lower_score_limit, upper_score_limit;
    int max_arm_copy, target_arm_copy;
    double masked_arm_threshold;
    const char *masked_seq;
    const unsigned char *snp;
    const unsigned char *unmappable;
} orc_sel;

/* Positions the region's candidates can cover: [orc_first_scan_start, stop_flanked + max_capture - min_sum - 1] */
int orc_n_positions(const orc_region *r, const orc_cfg *c);

/* condense_mips (mipgen.cpp:1670-1746) over the enumerated candidates (grid indices in enumeration
 * order, e.g. from orc_tile_replay): scan_best[scan_idx*2 + strand] = grid index of scan_strand_best_mip
 * or -1.  Lists are walked in push_front order (reverse enumeration, mipgen.cpp:475,489). */
void orc_condense(const orc_region *r, const orc_cfg *c, const orc_sel *s, const double *score,
                  const long *enum_idx, long n_enum, long *scan_best);

/* collapse_mips (mipgen.cpp:1617-1649): pos_best[pos_idx*2 + strand] = grid index of pos_strand_best_mip
 * or -1, pos_idx = position - orc_first_scan_start. */
void orc_collapse(const orc_region *r, const orc_cfg *c, const orc_sel *s, const double *score,
                  const long *scan_best, long *pos_best);

#ifdef __cplusplus
}
#endif
#endif
