// mg_common.cuh -- shared device/host definitions for the MIPgen B200 hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mipgen_b200.h"

// ---------------------------------------------------------------------------
// base codes (one byte per base in HBM; the reference distinguishes more than
// four symbols: 'N' invalidates an arm, '-' invalidates via mip_seq, every other
// character simply never matches a k-mer -- SVMipv4.cpp:63,116; MinusSVMipv4.cpp:24-26)
// ---------------------------------------------------------------------------
enum : uint8_t { B_A = 0, B_C = 1, B_G = 2, B_T = 3, B_N = 4, B_DASH = 5, B_OTHER = 6, B_NONE = 7 };

#define MG_MAX_PAIRS 1024
#define MG_MAX_OLIGO 64

// grid configuration as the kernels see it (lives in global memory, read uniformly)
struct DevConfig {
    int max_capture, min_capture, inc, max_mip_overlap;
    int n_cap, n_pairs, max_sum, min_sum;
    int n_oligo, max_arm, min_arm, pad0;
    int ext_len[MG_MAX_PAIRS];
    int lig_len[MG_MAX_PAIRS];
    int oligo_sizes[MG_MAX_OLIGO];
};

// one region as the kernels see it
struct DevRegion {
    int64_t seq_off;    // offset of the region's codes in the panel's code array
    int64_t grid_off;   // first global candidate index of the region's grid
    int64_t copy_off;   // offset into the panel's copy array, or -1
    int seq_len, seq_start, seq_stop;
    int start_flanked, stop_flanked;
    int first_scan, n_scan;
    int has_snp;        // the panel's snp prefix array has entries for this region
    int64_t aux_off;    // first entry of this region in the panel's masked / snp prefix arrays (seq_len + 1 entries each)
    int64_t unmap_off;  // offset of this region's [n_cap][seq_len] unmappable-start bytes, or -1
};

// one K-feat work item: a window of consecutive scan starts of one region, restricted to a range of
// capture sizes (so that a capture sweep does not blow a work item up to 1e5+ candidates)
struct DevTask {
    int64_t g0;      // global candidate index of the window's first grid point (capture index 0)
    int region;      // index into DevRegion[]
    int si0, nsi;    // first scan index of the window, number of scan starts
    int ci0, nci;    // capture-size indices [ci0, ci0 + nci)
    int ft0;         // index of the first factored-SVR work item nested in the window (row-table mode of K-feat)
};

// one explicit candidate (strand-oriented codes in a packed buffer)
struct DevCand {
    int64_t ext_off, lig_off, tgt_off;
    int ext_n, lig_n, tgt_n;
    int ext_len, lig_len, scan_size;
    int ext_copy, lig_copy;
    int pad;
};

// ---------------------------------------------------------------------------
// feature descriptors: what each of the 192 outputs of get_parameters is
// (SVMipv4.cpp:60-113)
// ---------------------------------------------------------------------------
enum : uint32_t {
    FK_RATIO = 0,  // count slot / (len - k + 1)
    FK_LEN = 1,    // raw arm / insert length
    FK_LRC = 2,    // long_range_content[j]
    FK_JUNC = 3,   // one-hot of the ligation junction
    FK_COPY = 4    // log10 copy (ext: j=0, lig: j=1)
};
// packed descriptor: kind[0:3] part[3:5] km1[5:7] slot_fwd[7:15] slot_rc[15:23] j[23:31]
//   part: 0 ext, 1 insert, 2 lig.  km1 = k-1 (the divisor is len - km1).
__host__ __device__ inline uint32_t fd_pack(uint32_t kind, uint32_t part, uint32_t km1, uint32_t sf, uint32_t sr, uint32_t j)
{
    return kind | (part << 3) | (km1 << 5) | (sf << 7) | (sr << 15) | (j << 23);
}

// per-warp shared-memory count slots (genomic orientation)
//   insert: tri[64] di[16] mono[4] gc[1]  -> 85 slots
//   arm:    di[16] mono[4] gc[1]          -> 21 slots
#define SLOT_INS_TRI 0
#define SLOT_INS_DI 64
#define SLOT_INS_MONO 80
#define SLOT_INS_GC 84
#define SLOT_EXT_DI 85
#define SLOT_EXT_MONO 101
#define SLOT_EXT_GC 105
#define SLOT_LIG_DI 106
#define SLOT_LIG_MONO 122
#define SLOT_LIG_GC 126
#define SLOT_COUNT 128

// ---------------------------------------------------------------------------
// SVR contraction tile shape (k_svr.cu)
// ---------------------------------------------------------------------------
#define SVR_BM 64        // candidates per CTA tile
#define SVR_BN 64        // support vectors per chunk
#define SVR_BK 32        // k-slab staged per pipeline stage
#define SVR_LDX 200      // padded row stride of the X tile in doubles (== 8 mod 16: conflict-free LDS.128)
#define SVR_LDB 40       // padded row stride of an SV slab in doubles
#define SVR_STAGES 4
#define SVR_CONSUMER_WARPS 8
#define SVR_THREADS ((SVR_CONSUMER_WARPS + 1) * 32)

// ---------------------------------------------------------------------------
// factored SVR (k_svr_fact.cu): block sizes, padded strides, per-chunk blob layout
// ---------------------------------------------------------------------------
#define FACT_C 16          // support vectors per chunk
#ifndef FACT_MATH_WARPS
#define FACT_MATH_WARPS 11
#define FACT_GATHER_WARPS 5
#define FACT_CPT 3                                      // candidates per gather thread
#endif
#define FACT_THREADS ((FACT_MATH_WARPS + FACT_GATHER_WARPS) * 32)       // 16 warps: 128 registers per thread
#define FACT_K_ARM 24      // 21 arm ratios + arm length + log copy, padded to a multiple of 4 (either role)
#define FACT_K_INS 88      // 86 insert features
#define FACT_LD_ARM 24     // strides == 8 mod 16 doubles: conflict-free LDS.128 fragment loads, no padding needed
#define FACT_LD_INS 88
#define FACT_OFF_EXT 0
#define FACT_OFF_LIG (FACT_C * FACT_LD_ARM)
#define FACT_OFF_INS (FACT_OFF_LIG + FACT_C * FACT_LD_ARM)
#define FACT_OFF_SS (FACT_OFF_INS + FACT_C * FACT_LD_INS)   // ss_ext[C], ss_lig[C] (incl. junction columns), ss_ins[C]
#define FACT_OFF_JT (FACT_OFF_SS + 3 * FACT_C)              // JT[17][C]: 1 - 2 s_i[junction]  (row 16: no junction -> 0)
#define FACT_BLOB (FACT_OFF_JT + 17 * FACT_C)               // doubles per chunk
#define FACT_SMEM_LIMIT (227 * 1024)
#define FACT_MAX_LEN 64
#define FACT_MAX_SPAN 128

// ---------------------------------------------------------------------------
// tensor-core SVR (k_svr_tc.cu): tile shape and operand images
// ---------------------------------------------------------------------------
#define TC_M 128           // candidates per CTA tile (= TMEM lanes)
#define TC_N 64            // support vectors per MMA tile
#define TC_KF 128          // fractional columns (127 used), FP16 hi / lo
#define TC_KF_USED 127
#define TC_KI 64           // integer columns, padded to one 128-byte operand row (19 used, 32 multiplied)
#define TC_KI_USED 19
#define TC_STAGES 3
#define TC_THREADS 640          // 4 service warps (producer, MMA issuer, TMEM allocator, spare) + 16 epilogue warps
#define TC_A_F_BYTES (2 * TC_M * 128)      // two K blocks of 64 FP16 columns x 128 rows
#define TC_A_I_BYTES (TC_M * 128)
#define TC_B_F_BYTES (2 * TC_N * 128)
#define TC_B_I_BYTES (TC_N * 128)
#define TC_IMG_BYTES (2 * TC_B_F_BYTES + TC_B_I_BYTES + 2 * TC_N * 8)   // hi | lo | integer | s_191[64] | s_192[64]
#define TC_STAGE_BYTES 43008               // image + the region's 64 weights, rounded up to a multiple of 1024

struct DevFact {
    int n_pairs, n_cap, n_ext, n_lig, n_sums, min_sum, max_sum, W;
    int cap_FA, cap_FQ, cap_FI, cap_R;   // shared-memory capacities in doubles / rows
    int ext_idx[FACT_MAX_LEN], lig_idx[FACT_MAX_LEN], sum_idx[FACT_MAX_SPAN];
    int ext_of[FACT_MAX_LEN], lig_of[FACT_MAX_LEN], sum_of[FACT_MAX_SPAN];   // the inverse maps: index -> arm length / arm sum
    int pair_e[MG_MAX_PAIRS], pair_l[MG_MAX_PAIRS];
    int blob_doubles;   // stride of one work item's row tables in the K-feat -> K-svr hand-off buffer (FACT_ROWS_DOUBLES)
};
// row tables of one factored-SVR work item as K-feat writes them and K-svr bulk-copies them into shared memory:
// FA[cap_FA] | FQ[cap_FQ] | FI[cap_FI] | xx[cap_R] (doubles), then jc[cap_R] (ints); stride rounded up to 128 bytes
#define FACT_ROWS_DOUBLES(f) ((((f).cap_FA + (f).cap_FQ + (f).cap_FI + (f).cap_R) + ((f).cap_R + 1) / 2 + 15) & ~15)

// one factored-SVR work item: W scan starts of one region, one capture size, one strand
struct DevFTask {
    int64_t g0;   // grid offset of the region
    int region, si0, nsi, ci, strand, pad;
};

// ---------------------------------------------------------------------------
// host context
// ---------------------------------------------------------------------------
struct HostConfig {
    int max_capture = 0, min_capture = 0, inc = 1, max_mip_overlap = 0;
    std::vector<int> ext_len, lig_len, oligo_sizes;
    int n_cap = 0, max_sum = 0, min_sum = 0, max_arm = 0, min_arm = 0;
};

struct EventPair { cudaEvent_t a, b; int which; long units; bool ok; };
struct CachedBlock { void *ptr; size_t bytes; };

struct mg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // config
    bool has_cfg = false;
    long cfg_serial = 0;  // bumped by every mg_set_config: panels remember the one they were built under
    HostConfig cfg;
    DevConfig *d_cfg = nullptr;
    // tables
    uint32_t *d_fdesc = nullptr;    // [192] packed feature descriptors (explicit front-end: count slots)
    uint32_t *d_fdesc_win = nullptr;  // [192] the same with prefix-table rows (window front-end)
    double *d_logcopy = nullptr;    // [102] log10(copy) for copy 0..100 (glibc), [101] = 2.0
    double *d_exp2tab = nullptr;    // [64] 2^(j/64), K-svr's exp table
    double *d_exp2tab256 = nullptr; // [256] 2^(j/256), the tensor-core kernel's
    unsigned long long *d_work = nullptr;  // [0..2] DMMAs, exp elements, gathered triples EXECUTED by the factored K-svr since the
                                           // last mg_reset_timings (tasks that exit after the claim phase add nothing); [3] scratch
    // model
    bool has_model = false;
    int n_sv = 0, n_sv_pad = 0;
    double gamma = 0, rho = 0;
    double *d_sv = nullptr;     // [n_sv_pad][192] row-major (cross-check kernel)
    double *d_sv_tiled = nullptr; // [n_sv_pad/64][6][64][SVR_LDB]: the padded smem slab image K-svr bulk-copies
    double *d_ss = nullptr;     // [n_sv_pad] ||s||^2 (incl. features beyond 192)
    double *d_alpha = nullptr;  // [n_sv_pad], 0 in the padding
    double *d_tail = nullptr;   // [n_sv_pad] sum of squares of SV features with index > 192
    // factored SVR
    int svr_mode = 0;           // 0 auto (factored when the config fits), 1 dense, 2 factored
    bool fact_ok = false;       // config fits the factored kernel's shared-memory tables
    int fact_W = 0;
    size_t fact_smem = 0;
    DevFact *d_fact = nullptr;
    DevFact h_fact{};               // host copy (work accounting)
    double *d_fact_blob = nullptr;  // [n_sv_pad/FACT_C][FACT_BLOB] per-chunk SV blocks + block norms
    double zero_score = 0;          // SVR value of the all-zero vector (invalid candidates)
    // tensor-core SVR (svr_mode 3)
    bool tc_ok = false;             // the model's integer columns are small integers: representable exactly in FP16
    uint8_t *d_tc_img = nullptr;    // [n_sv_pad / TC_N][TC_IMG_BYTES] pre-swizzled operand images
    double *d_tc_centre = nullptr;  // [TC_KF + TC_KI] column centres
    double *d_tc_expc = nullptr;    // [n_sv_pad] exp(-gamma ||s'||^2)
    // workspace
    double *d_x = nullptr;      // feature rows of the chunk in flight
    size_t x_rows_cap = 0;
    double *d_rows = nullptr;   // row tables of the factored-SVR work items in flight (K-feat writes, K-svr reads)
    size_t rows_cap = 0;        // in bytes
    // freed device blocks kept for reuse (the per-call buffers of mg_score_regions / mg_score_candidates)
    std::vector<CachedBlock> pool;
    std::vector<CachedBlock> live;
    size_t pool_bytes = 0;
    // timing
    std::vector<EventPair> ev_pending;
    std::vector<cudaEvent_t> ev_free;
    mg_timings tm{};
    cudaEvent_t sw_a = nullptr, sw_b = nullptr;
    int sm_count = 148;
};

struct mg_panel {
    mg_ctx *ctx = nullptr;
    long cfg_serial = 0;
    int n_regions = 0;
    int64_t n_cand = 0;
    int64_t n_valid_static = 0;
    std::vector<int64_t> offsets;  // n+1
    std::vector<DevRegion> h_regions;
    DevRegion *d_regions = nullptr;
    std::vector<DevTask> h_windows;  // scan-start windows (all capture sizes), ascending in g0: the unit of chunking
    std::vector<DevTask> h_tasks;    // K-feat work items: windows split along the capture dimension
    std::vector<int> task_start;     // [n_windows+1] first K-feat work item of each window
    DevTask *d_tasks = nullptr;
    std::vector<DevFTask> h_ftasks;     // factored-SVR tasks, grouped by K-feat window
    std::vector<int> ftask_start;       // [n_windows+1] first factored task of each window
    int64_t row_table_bytes = 0;        // distinct rows of all factored tasks, in bytes (see mg_panel_row_table_bytes)
    DevFTask *d_ftasks = nullptr;
    double *d_w = nullptr;              // [n_regions][n_sv_pad] alpha * exp(-g d_lrc)
    int w_n_sv_pad = 0;
    double *d_w_tc = nullptr;           // the same times exp(-g ||s'||^2): per-SV weights of the tensor-core kernel
    int w_tc_n_sv_pad = 0;
    int span_cap = 0, pf_stride = 0;
    uint8_t *d_codes = nullptr;
    char *d_ascii = nullptr;       // the sequences as given (the record formatter prints them)
    int64_t n_codes = 0;
    double *d_lrc = nullptr;       // [n_regions][44]
    int *d_copies = nullptr;
    int *d_maskpf = nullptr;       // per region: prefix counts of 'N' in the masked sequence (seq_len + 1 entries)
    int *d_snppf = nullptr;        // per region: prefix counts of SNP positions, or null
    uint8_t *d_unmap = nullptr;    // [n_cap][seq_len] per region that declares unmappable MIP starts, or null
    bool has_sel_inputs = false;   // some region carries masked_seq / snp / unmappable
    uint8_t *d_valid = nullptr;
    uint8_t *d_state = nullptr;   // per grid point: 0 skipped, 1 invalid (N / '-' in an arm), 2 scored -- what the factored SVR reads
    double *d_logistic = nullptr;
    double *d_svr = nullptr;
    double *d_feat = nullptr;      // only when features are fetched for the whole panel
    bool has_logistic = false, has_svr = false, has_feat = false, has_valid = false;
};

#define CUDA_TRY(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return MG_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

// host arithmetic shared by mg_api.cu and mg_tile.cu
int mg_host_config_from(const mg_config *c, HostConfig &h, std::string &err);
int mg_host_first_scan(const HostConfig &c, const mg_region *r);      // mipgen.cpp:421-425
int mg_host_n_scan(const HostConfig &c, const mg_region *r);
int mg_host_n_positions(const HostConfig &c, const mg_region *r);     // first scan start .. stop_flanked + max_capture - min_sum - 1

// caching device allocator (mg_api.cu): stream-ordered reuse on the context's single stream
cudaError_t mg_dev_alloc(mg_ctx *ctx, void **out, size_t bytes);
void mg_dev_free(mg_ctx *ctx, void *ptr);
void mg_dev_trim(mg_ctx *ctx);

// kernels / launchers (defined in k_feat.cu, k_svr.cu)
enum { TM_FEAT = 0, TM_SVR = 1, TM_OTHER = 2 };
int mg_time_begin(mg_ctx *ctx, int which, long units);
int mg_time_end(mg_ctx *ctx);

int launch_encode(mg_ctx *ctx, const char *d_ascii, uint8_t *d_codes, int64_t n);
int launch_lrc(mg_ctx *ctx, const uint8_t *d_codes, int n, int denom, double *d_out44);
// grid front-end: any of valid/logistic/x may be null; x rows are written at row (g - g_base).
// tasks [task0, task1) hold n_cand consecutive candidates starting at global index g_base
// Row-table mode (d_rows != null): instead of feature rows, the distinct arm / insert rows of the factored-SVR work items
// nested in the tasks are written to d_rows (work item ftask_base first) and every grid point's state to d_state.
int launch_feat_grid(mg_ctx *ctx, const mg_panel *p, int task0, int task1, int64_t g_base, int64_t n_cand, uint8_t *d_valid,
                     double *d_logistic, double *d_x, uint8_t *d_state = nullptr, double *d_rows = nullptr, int ftask_base = 0);
int launch_feat_setup(mg_ctx *ctx);
// explicit front-end
int launch_feat_explicit(mg_ctx *ctx, const DevCand *d_cands, const uint8_t *d_codes, const double *d_lrc,
                         int64_t n, double *d_logistic, double *d_x);
// SVR: rows [0, n) of d_x (ld 192, padded to a multiple of SVR_BM rows) -> d_out[n]
// d_valid (nullable): rows with valid[g]==0 are written as NaN.
int launch_svr(mg_ctx *ctx, const double *d_x, int64_t n, const uint8_t *d_valid, double *d_out);
int launch_svr_setup(mg_ctx *ctx);
int launch_fact_setup(mg_ctx *ctx);
int launch_select(mg_ctx *ctx, const mg_panel *p, const int64_t *d_scan_off, const int64_t *d_pos_off, int64_t total_scan,
                  int64_t total_pos, const double *d_score, const mg_select_params *sp, int64_t *d_scan_best, int64_t *d_pos_best);
// K-replay: d_out_idx == null counts the enumerated grid points per scan start into d_count, otherwise fills their indices
int launch_replay(mg_ctx *ctx, const mg_panel *p, const int64_t *d_scan_off, int64_t total_scan, const double *d_score, const mg_select_params *sp,
                  int *d_count, const int64_t *d_out_off, int64_t *d_out_idx);
int launch_gather(mg_ctx *ctx, const int64_t *d_idx, int64_t n, const double *d_a, double *d_out_a, const double *d_b, double *d_out_b);
int launch_count_valid(mg_ctx *ctx, const uint8_t *d_valid, int64_t n, unsigned long long *d_count);
int launch_lrc_weights(mg_ctx *ctx, const mg_panel *p, double *d_w);
// work items [ftask0, ftask1): their row tables are d_rows[0 ..) in order, the grid points' states d_state[] (panel-wide)
int launch_svr_fact(mg_ctx *ctx, const mg_panel *p, int ftask0, int ftask1, const double *d_rows, int64_t n_cand,
                    const uint8_t *d_state, const double *d_w, double *d_out);
int mg_upload_lrc_tables(mg_ctx *ctx, const uint8_t *k, const uint8_t *code);
// tensor-core SVR (k_svr_tc.cu)
int launch_tc_setup(mg_ctx *ctx);
bool mg_tc_prepare_model(const std::vector<double> &sv, int n_sv_pad, int n_sv, double gamma, std::vector<uint8_t> &img,
                         std::vector<double> &centre, std::vector<double> &exp_c);
int launch_lrc_weights_tc(mg_ctx *ctx, const mg_panel *p, double *d_w);
int launch_svr_tc(mg_ctx *ctx, const mg_panel *p, const double *d_x, int64_t g0, int64_t g1, const uint8_t *d_valid, const double *d_w,
                  double *d_out);
int launch_svr_direct(mg_ctx *ctx, const double *d_x, int64_t n, int64_t ld, double *d_out);
