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
// The selection-only fields of an SVMipv4 object that design_mip fills (mipgen.cpp:606-625, 634-760) are derived
// per grid point from per-region tables: arm_fraction_masked from prefix counts of 'N' in the masked sequence,
// snp_count from prefix counts of SNP positions, mapping_failed from the per-capture-size unmappable MIP starts,
// arm copy numbers from the same tables K-feat uses.
// Both kernels only compare and copy scores: results are exactly the oracle's on the same score grid.
#include <algorithm>

#include "mg_common.cuh"

namespace {

constexpr int kMaskWords = 128;  // up to 4096 (capture, pair) combinations per scan start

struct SelParams {
    int method, heuristic;
    double lower, upper;
    int max_arm_copy, target_arm_copy;
    double masked_thr;
};

struct SelAux {  // per-panel tables behind arm_fraction_masked / snp_count / mapping_failed
    const int *maskpf;
    const int *snppf;
    const uint8_t *unmap;
};

// what design_mip leaves in the object besides the score (mipgen.cpp:606-625, 634-760), for grid point (s, capture, pair, strand)
struct SelMip {
    int ec, lc, snp, mapping_failed;
    double masked;
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

__device__ __forceinline__ SelMip sel_mip(const DevConfig *__restrict__ cfg, const DevRegion &r, const int *__restrict__ copies,
                                          const SelAux &aux, int s, int ci, int cap, int p, int strand)
{
    SelMip m;
    const int e = cfg->ext_len[p], l = cfg->lig_len[p], t = s + cap - (e + l) - 1;
    const int ext_start = strand ? t + 1 : s - e, lig_start = strand ? s - l : t + 1;  // Plus/MinusSVMipv4 ctors
    m.ec = m.lc = 1;
    if (r.copy_off >= 0) { m.ec = copy_of(cfg, r, copies, ext_start, e); m.lc = copy_of(cfg, r, copies, lig_start, l); }
    // arm windows clamped like std::string::substr (statically valid candidates lie inside the sequence anyway)
    const int ea = max(0, min(r.seq_len, ext_start - r.seq_start)), eb = max(ea, min(r.seq_len, ext_start - r.seq_start + e));
    const int la = max(0, min(r.seq_len, lig_start - r.seq_start)), lb = max(la, min(r.seq_len, lig_start - r.seq_start + l));
    const int *mp = aux.maskpf + r.aux_off;
    const int n_masked = (mp[eb] - mp[ea]) + (mp[lb] - mp[la]);
    m.masked = __ddiv_rn((double)n_masked, (double)(l + e));  // mipgen.cpp:610
    m.snp = 0;
    if (r.has_snp) {
        const int *sp = aux.snppf + r.aux_off;
        m.snp = (sp[eb] - sp[ea]) + (sp[lb] - sp[la]);           // mipgen.cpp:634-636, 698-700
    }
    m.mapping_failed = 0;
    if (r.unmap_off >= 0) {
        const int start = (strand ? lig_start : ext_start) - r.seq_start;  // get_mip_start()
        if (start >= 0 && start < r.seq_len) m.mapping_failed = aux.unmap[r.unmap_off + (int64_t)ci * r.seq_len + start] != 0;  // :615-625
    }
    return m;
}

__global__ void __launch_bounds__(128)
k_condense(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const int64_t *__restrict__ scan_off, int n_regions,
           int64_t total_scan, const int *__restrict__ copies, SelAux aux, const uint8_t *__restrict__ valid,
           const double *__restrict__ score, SelParams sp, int64_t *__restrict__ scan_best)
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
    int chosen_copy = 0;         // declared once per position in the reference, shared by both strands (:1677-1680)
    double chosen_masked = 0.0;
    for (int strand = 0; strand < 2; strand++) {
        int64_t best = -1;
        double best_score = 0.0;
        int best_snp = 0;
        bool skip_ahead = false;
        for (int k = n_comb - 1; k >= 0 && !skip_ahead; k--) {
            if (!((mask[k >> 5] >> (k & 31)) & 1)) continue;
            const int ci = k / n_pairs, p = k - ci * n_pairs;
            const SelMip m = sel_mip(cfg, r, copies, aux, s, ci, cfg->max_capture - ci * inc, p, strand);
            if (m.ec * m.lc > sp.max_arm_copy) continue;                                 // :1689
            if (m.mapping_failed) continue;                                              // :1690
            const int current = m.ec > m.lc ? m.ec : m.lc;                               // :1692
            const int64_t idx = base + (int64_t)k * 2 + strand;
            const double sc = score[idx];
            bool take = false, keep_chosen = false;
            if (best < 0) take = true;                                                   // :1695-1700
            else if (m.masked > sp.masked_thr && m.masked < chosen_masked) take = true;  // :1701-1706
            else if (current > sp.target_arm_copy && current < chosen_copy) take = true; // :1709-1714
            else if (current <= sp.target_arm_copy) {                                    // :1715
                if (sc < sp.lower && sc > best_score) take = true;                       // :1717-1722
                else if (sc > sp.lower) {
                    if (m.snp < best_snp) take = true;                                   // :1725-1730
                    else if (m.snp == best_snp && sc > best_score) {                     // :1731-1737
                        take = keep_chosen = true;
                        if (sc > sp.upper) skip_ahead = true;
                    }
                }
            }
            if (take) {
                best = idx; best_score = sc; best_snp = m.snp;
                if (!keep_chosen) { chosen_copy = current; chosen_masked = m.masked; }
            }
        }
        scan_best[t * 2 + strand] = best;
    }
}

__global__ void __launch_bounds__(128)
k_collapse(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const int64_t *__restrict__ scan_off,
           const int64_t *__restrict__ pos_off, int n_regions, int64_t total_pos, const int *__restrict__ copies, SelAux aux,
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
    int best_snp = 0;
    for (int s = s0; s <= s1; s++) {  // scan starts ascending, as the std::map is walked (:1620)
        const int64_t cur = scan_best[(scan_off[lo] + (s - r.first_scan)) * 2 + strand];
        if (cur < 0) continue;
        const int64_t local = cur - r.grid_off;
        const int k = (int)((local >> 1) % ((int64_t)n_cap * n_pairs));
        const int ci = k / n_pairs, p = k - ci * n_pairs;
        const int cap = cfg->max_capture - ci * inc, e = cfg->ext_len[p], l = cfg->lig_len[p];
        if (s + cap - (e + l) - 1 < pos) continue;  // does not cover this position
        const SelMip m = sel_mip(cfg, r, copies, aux, s, ci, cap, p, strand);
        if (m.ec * m.lc > sp.max_arm_copy || m.ec > sp.target_arm_copy || m.lc > sp.target_arm_copy) continue;  // :1628
        if (m.masked > sp.masked_thr) continue;                                                                // :1629
        const double sc = score[cur];
        if (best < 0 || m.snp < best_snp || (sc > best_score && m.snp == best_snp)) { best = cur; best_score = sc; best_snp = m.snp; }  // :1634-1645
    }
    pos_best[t] = best;
}

// out_x[i] = x[idx[i]] (NaN for idx < 0), for up to two score arrays at once
__global__ void __launch_bounds__(256) k_gather_scores(const int64_t *__restrict__ idx, int64_t n, const double *__restrict__ a,
                                                       double *__restrict__ out_a, const double *__restrict__ b, double *__restrict__ out_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t g = idx[i];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (out_a) out_a[i] = g >= 0 ? a[g] : nan;
    if (out_b) out_b[i] = g >= 0 ? b[g] : nan;
}

__global__ void __launch_bounds__(256) k_count_valid(const uint8_t *__restrict__ valid, int64_t n, unsigned long long *__restrict__ count)
{
    unsigned int c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += valid[i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}


// K-replay: which grid points does the tile loop enumerate (and print to all_mips.txt, mipgen.cpp:474, 488), in its order?
// The enumeration order is the grid order, so per scan start the answer is a subset of its (capture, pair) combinations, both
// strands of each: the same replay of the score-dependent shortcuts as K-condense's pass 1.  kFill = false counts them per scan
// start; kFill = true writes their panel-global grid indices at the scan start's offset (exclusive prefix sum of the counts).
template <bool kFill>
__global__ void __launch_bounds__(128)
k_replay(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const int64_t *__restrict__ scan_off, int n_regions,
         int64_t total_scan, const uint8_t *__restrict__ valid, const double *__restrict__ score, SelParams sp, int *__restrict__ count,
         const int64_t *__restrict__ out_off, int64_t *__restrict__ out_idx)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_scan) return;
    int lo = 0, hi = n_regions - 1;  // region of this scan start
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (scan_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const DevRegion r = regions[lo];
    const int si = (int)(t - scan_off[lo]);
    const int n_cap = cfg->n_cap, n_pairs = cfg->n_pairs, inc = cfg->inc;
    const int64_t base = r.grid_off + (int64_t)si * n_cap * n_pairs * 2;
    int n = 0;
    int64_t at = kFill ? out_off[t] : 0;
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
                    if (kFill) { out_idx[at] = idx; out_idx[at + 1] = idx + 1; at += 2; }
                    n += 2;
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
    if (!kFill) count[t] = n;
}

}  // namespace

int launch_select(mg_ctx *ctx, const mg_panel *p, const int64_t *d_scan_off, const int64_t *d_pos_off, int64_t total_scan,
                  int64_t total_pos, const double *d_score, const mg_select_params *msp, int64_t *d_scan_best, int64_t *d_pos_best)
{
    if ((int64_t)ctx->cfg.n_cap * (int64_t)ctx->cfg.ext_len.size() > kMaskWords * 32) {
        ctx->err = "mg_panel_select: more than 4096 (capture, arm pair) combinations per scan start";
        return MG_ERR_INVALID;
    }
    SelParams sp;
    sp.method = msp->method; sp.heuristic = msp->heuristic; sp.lower = msp->lower_score_limit; sp.upper = msp->upper_score_limit;
    sp.max_arm_copy = msp->max_arm_copy; sp.target_arm_copy = msp->target_arm_copy; sp.masked_thr = msp->masked_arm_threshold;
    SelAux aux;
    aux.maskpf = p->d_maskpf; aux.snppf = p->d_snppf; aux.unmap = p->d_unmap;
    if (total_scan > 0) {
        mg_time_begin(ctx, TM_OTHER, total_scan);
        k_condense<<<(unsigned)((total_scan + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, p->n_regions,
                                                                                 total_scan, p->d_copies, aux, p->d_valid, d_score, sp,
                                                                                 d_scan_best);
        mg_time_end(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (total_pos > 0) {
        mg_time_begin(ctx, TM_OTHER, total_pos);
        k_collapse<<<(unsigned)((total_pos * 2 + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, d_pos_off,
                                                                                    p->n_regions, total_pos, p->d_copies, aux, d_score,
                                                                                    d_scan_best, sp, d_pos_best);
        mg_time_end(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return MG_OK;
}

int launch_gather(mg_ctx *ctx, const int64_t *d_idx, int64_t n, const double *d_a, double *d_out_a, const double *d_b, double *d_out_b)
{
    if (n <= 0) return MG_OK;
    mg_time_begin(ctx, TM_OTHER, n);
    k_gather_scores<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_idx, n, d_a, d_out_a, d_b, d_out_b);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

int launch_count_valid(mg_ctx *ctx, const uint8_t *d_valid, int64_t n, unsigned long long *d_count)
{
    if (n <= 0) return MG_OK;
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_count_valid<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_valid, n, d_count);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

int launch_replay(mg_ctx *ctx, const mg_panel *p, const int64_t *d_scan_off, int64_t total_scan, const double *d_score, const mg_select_params *msp,
                  int *d_count, const int64_t *d_out_off, int64_t *d_out_idx)
{
    if (total_scan <= 0) return MG_OK;
    SelParams sp;
    sp.method = msp->method; sp.heuristic = msp->heuristic; sp.lower = msp->lower_score_limit; sp.upper = msp->upper_score_limit;
    sp.max_arm_copy = msp->max_arm_copy; sp.target_arm_copy = msp->target_arm_copy; sp.masked_thr = msp->masked_arm_threshold;
    mg_time_begin(ctx, TM_OTHER, total_scan);
    if (d_out_idx)
        k_replay<true><<<(unsigned)((total_scan + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, p->n_regions, total_scan,
                                                                                       p->d_valid, d_score, sp, nullptr, d_out_off, d_out_idx);
    else
        k_replay<false><<<(unsigned)((total_scan + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, d_scan_off, p->n_regions, total_scan,
                                                                                        p->d_valid, d_score, sp, d_count, nullptr, nullptr);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}
