// k_select.cu -- the selection front-end that consumes every candidate right after scoring
// (SURVEY.md section 8f, rank 2), as device reductions over the score grid:
//
//   K-condense  one thread per (region, scan start): replays the score-dependent control flow of the
//               tile loop (mipgen.cpp:426-497: optimal-score shortcuts, logistic heuristic with its int
//               truncation) to find which grid points the reference would have enumerated, then walks them
//               in push_front order (reverse enumeration, mipgen.cpp:475,489) through condense_mips'
//               rules (mipgen.cpp:1670-1746)  ->  scan_strand_best_mip as grid indices.
//   K-collapse  one thread per (region, position, strand): collapse_mips (mipgen.cpp:1617-1649), the best
//               scan-start winner covering the position  ->  pos_strand_best_mip as grid indices.
//
// Not modelled on the device (host-only inputs outside the scoring path): TRF masking (arm_fraction_masked
// = 0), SNP counts (= 0) and mapping failures ('0') -- the state of every run without -trf / -snp_file and
// with uniquely mapping capture sites.  Arm copy numbers come from the same tables K-feat uses.
// Both kernels only compare and copy scores: results are exactly the oracle's on the same score grid.
#include "mg_common.cuh"

namespace {

constexpr int kMaskWords = 128;  // up to 4096 (capture, pair) combinations per scan start

struct SelParams {
    int method, heuristic;
    double lower, upper;
    int max_arm_copy, target_arm_copy;
};

__device__ __forceinline__ int copy_of(const DevConfig *__restrict__ cfg, const DevRegion &r, const int *__restrict__ copies, int start, int len)
{
    if (r.copy_off < 0) return 1;
    for (int k = 0; k < cfg->n_oligo; k++)
        if (cfg->oligo_sizes[k] == len) {
            const int i = start - r.seq_start;
            if (i < 0 || i >= r.seq_len) return 0;
            return copies[r.copy_off + (int64_t)k * r.seq_len + i];
        }
    return 0;
}

// arm copy numbers of grid point (s, capture, pair p, strand)
__device__ __forceinline__ void arm_copies(const DevConfig *__restrict__ cfg, const DevRegion &r, const int *__restrict__ copies, int s,
                                           int cap, int p, int strand, int &ec, int &lc)
{
    const int e = cfg->ext_len[p], l = cfg->lig_len[p], t = s + cap - (e + l) - 1;
    ec = copy_of(cfg, r, copies, strand ? t + 1 : s - e, e);
    lc = copy_of(cfg, r, copies, strand ? s - l : t + 1, l);
}

__global__ void __launch_bounds__(128)
k_condense(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const int64_t *__restrict__ scan_off, int n_regions,
           int64_t total_scan, const int *__restrict__ copies, const uint8_t *__restrict__ valid, const double *__restrict__ score,
           SelParams sp, int64_t *__restrict__ scan_best)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_scan) return;
    int lo = 0, hi = n_regions - 1;  // region of this scan start
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (scan_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const DevRegion r = regions[lo];
    const int si = (int)(t - scan_off[lo]), s = r.first_scan + si;
    const int n_cap = cfg->n_cap, n_pairs = cfg->n_pairs, inc = cfg->inc;
    const int64_t base = r.grid_off + (int64_t)si * n_cap * n_pairs * 2;

    // ---- pass 1: which (capture, pair) combinations does the tile loop reach? ----
    uint32_t mask[kMaskWords];
    const int n_comb = n_cap * n_pairs, n_words = (n_comb + 31) >> 5;
    for (int w = 0; w < n_words; w++) mask[w] = 0;
    double previous_best = 0.0;  // :426
    for (int ci = 0; ci < n_cap; ci++) {
        const int cap = cfg->max_capture - ci * inc;
        if (cap > r.stop_flanked - r.start_flanked + cfg->max_mip_overlap && cap - inc >= cfg->min_capture) continue;  // :429
        if (previous_best > sp.upper) continue;                                                                       // :430
        int p = 0;
        while (p < n_pairs) {
            const int sum = cfg->ext_len[p] + cfg->lig_len[p];
            int q = p;
            while (q < n_pairs && cfg->ext_len[q] + cfg->lig_len[q] == sum) q++;
            if (!(previous_best > sp.upper && sum != cfg->min_sum)) {  // :434
                int prev_minus = 0, prev_plus = 0;                     // ints in the reference (:435-436)
                for (int k = p; k < q; k++) {
                    const int64_t idx = base + (int64_t)(ci * n_pairs + k) * 2;
                    if (!valid[idx]) continue;                         // :443-444
                    const double plus = score[idx], minus = score[idx + 1];
                    mask[(ci * n_pairs + k) >> 5] |= 1u << ((ci * n_pairs + k) & 31);
                    const bool stop = sp.method == 0 && sp.heuristic && plus < prev_plus && minus < prev_minus;  // :494
                    previous_best = minus > plus ? minus : plus;       // :495
                    prev_minus = __double2int_rz(minus);               // :496
                    prev_plus = __double2int_rz(plus);                 // :497
                    if (stop) break;
                }
            }
            p = q;
        }
    }

    // ---- pass 2: condense_mips per strand, lists walked newest first ----
    int chosen_copy = 0;  // declared once per position in the reference, shared by both strands (:1677)
    for (int strand = 0; strand < 2; strand++) {
        int64_t best = -1;
        double best_score = 0.0;
        bool skip_ahead = false;
        for (int k = n_comb - 1; k >= 0 && !skip_ahead; k--) {
            if (!((mask[k >> 5] >> (k & 31)) & 1)) continue;
            const int ci = k / n_pairs, p = k - ci * n_pairs;
            int ec = 1, lc = 1;
            if (r.copy_off >= 0) arm_copies(cfg, r, copies, s, cfg->max_capture - ci * inc, p, strand, ec, lc);
            if (ec * lc > sp.max_arm_copy) continue;                                     // :1689
            const int current = ec > lc ? ec : lc;                                       // :1692
            const int64_t idx = base + (int64_t)k * 2 + strand;
            const double sc = score[idx];
            if (best < 0) { best = idx; best_score = sc; chosen_copy = current; }         // :1695-1700
            else if (current > sp.target_arm_copy && current < chosen_copy) { best = idx; best_score = sc; chosen_copy = current; }  // :1709
            else if (current <= sp.target_arm_copy) {                                    // :1715
                if (sc < sp.lower && sc > best_score) { best = idx; best_score = sc; chosen_copy = current; }
                else if (sc > sp.lower && sc > best_score) {                             // :1723-1737 (equal snp counts)
                    best = idx; best_score = sc;
                    if (sc > sp.upper) skip_ahead = true;
                }
            }
        }
        scan_best[t * 2 + strand] = best;
    }
}

__global__ void __launch_bounds__(128)
k_collapse(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const int64_t *__restrict__ scan_off,
           const int64_t *__restrict__ pos_off, int n_regions, int64_t total_pos, const int *__restrict__ copies,
           const double *__restrict__ score, const int64_t *__restrict__ scan_best, SelParams sp, int64_t *__restrict__ pos_best)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_pos * 2) return;
    const int64_t pt = t >> 1;
    const int strand = (int)(t & 1);
    int lo = 0, hi = n_regions - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pos_off[mid] <= pt) lo = mid; else hi = mid - 1;
    }
    const DevRegion r = regions[lo];
    const int pos = r.first_scan + (int)(pt - pos_off[lo]);
    const int n_cap = cfg->n_cap, n_pairs = cfg->n_pairs, inc = cfg->inc;
    const int max_scan = cfg->max_capture - cfg->min_sum;  // longest scan window
    int s0 = pos - max_scan + 1;
    if (s0 < r.first_scan) s0 = r.first_scan;
    const int s1 = min(pos, r.first_scan + r.n_scan - 1);
    int64_t best = -1;
    double best_score = 0.0;
    for (int s = s0; s <= s1; s++) {  // scan starts ascending, as the std::map is walked (:1620)
        const int64_t cur = scan_best[(scan_off[lo] + (s - r.first_scan)) * 2 + strand];
        if (cur < 0) continue;
        const int64_t local = cur - r.grid_off;
        const int k = (int)((local >> 1) % ((int64_t)n_cap * n_pairs));
        const int ci = k / n_pairs, p = k - ci * n_pairs;
        const int cap = cfg->max_capture - ci * inc, e = cfg->ext_len[p], l = cfg->lig_len[p];
        if (s + cap - (e + l) - 1 < pos) continue;  // does not cover this position
        if (r.copy_off >= 0) {
            int ec, lc;
            arm_copies(cfg, r, copies, s, cap, p, strand, ec, lc);
            if (ec * lc > sp.max_arm_copy || ec > sp.target_arm_copy || lc > sp.target_arm_copy) continue;  // :1628
        }
        const double sc = score[cur];
        if (best < 0 || sc > best_score) { best = cur; best_score = sc; }  // :1634-1645 (equal snp counts)
    }
    pos_best[t] = best;
}

}  // namespace

int launch_select(mg_ctx *ctx, const mg_panel *p, const int64_t *d_scan_off, const int64_t *d_pos_off, int64_t total_scan,
                  int64_t total_pos, const double *d_score, int method, int heuristic, double lower, double upper, int max_arm_copy,
                  int target_arm_copy, int64_t *d_scan_best, int64_t *d_pos_best)
{
    if ((int64_t)ctx->cfg.n_cap * (int64_t)ctx->cfg.ext_len.size() > kMaskWords * 32) {
        ctx->err = "mg_panel_select: more than 4096 (capture, arm pair) combinations per scan start";
        return MG_ERR_INVALID;
    }
    SelParams sp;
    sp.method = method; sp.heuristic = heuristic; sp.lower = lower; sp.upper = upper;
    sp.max_arm_copy = max_arm_copy; sp.target_arm_copy = target_arm_copy;
    if (total_scan > 0) {
        mg_time_begin(ctx, TM_OTHER, total_scan);
        k_condense<<<(unsigned)((total_scan + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, p->n_regions,
                                                                                 total_scan, p->d_copies, p->d_valid, d_score, sp,
                                                                                 d_scan_best);
        mg_time_end(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (total_pos > 0) {
        mg_time_begin(ctx, TM_OTHER, total_pos);
        k_collapse<<<(unsigned)((total_pos * 2 + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, d_pos_off,
                                                                                    p->n_regions, total_pos, p->d_copies, d_score,
                                                                                    d_scan_best, sp, d_pos_best);
        mg_time_end(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return MG_OK;
}
