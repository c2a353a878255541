// mg_tile.cu -- one call per batch of Featurev5 objects, and the same over several GPUs of one box.
//
// mg_tile_regions does for a list of regions what mipgen::tile_regions does per feature up to collapse_mips
// (mipgen.cpp:412-505): the candidate loop nest (K-feat / K-svr), condense_mips and collapse_mips (K-condense /
// K-collapse), in sub-batches of bounded size, returning the winners (scan_strand_best_mip / pos_strand_best_mip
// as region-local grid indices, plus their scores) and optionally the full grids.  pick_mips stays with the caller.
//
// The *_multi calls shard the region list over several contexts (one per GPU): regions are independent for scoring
// (mipgen.cpp:412-525, state cleared at 522-524), so the partition is a plain longest-processing-time assignment on
// grid sizes, every context works on its own host thread / device / stream and writes its regions' slices of the
// caller's arrays; nothing is exchanged between devices (no collective).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <thread>

#include "mg_common.cuh"

extern "C" int mg_tile_sizes(const mg_config *cfg, const mg_region *regions, int n, int64_t *grid_off, int64_t *scan_off, int64_t *pos_off)
{
    HostConfig h;
    std::string err;
    if (n < 0 || (n > 0 && !regions) || mg_host_config_from(cfg, h, err) != MG_OK) return MG_ERR_INVALID;
    int64_t g = 0, s = 0, p = 0;
    for (int i = 0; i <= n; i++) {
        if (grid_off) grid_off[i] = g;
        if (scan_off) scan_off[i] = s;
        if (pos_off) pos_off[i] = p;
        if (i == n) break;
        const int ns = mg_host_n_scan(h, &regions[i]);
        g += (int64_t)ns * h.n_cap * (int64_t)h.ext_len.size() * 2;
        s += ns;
        p += mg_host_n_positions(h, &regions[i]);
    }
    return MG_OK;
}

extern "C" int mg_partition_regions(const mg_config *cfg, const mg_region *regions, int n, int n_parts, int *owner)
{
    if (n_parts <= 0 || n < 0 || (n > 0 && (!regions || !owner))) return MG_ERR_INVALID;
    std::vector<int64_t> goff((size_t)n + 1);
    if (mg_tile_sizes(cfg, regions, n, goff.data(), nullptr, nullptr) != MG_OK) return MG_ERR_INVALID;
    // longest processing time first: regions by descending grid size (ties: lower index), each to the lightest part
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return goff[a + 1] - goff[a] > goff[b + 1] - goff[b]; });
    std::vector<int64_t> load((size_t)n_parts, 0);
    for (int i : order) {
        int best = 0;
        for (int k = 1; k < n_parts; k++)
            if (load[k] < load[best]) best = k;
        owner[i] = best;
        load[best] += goff[i + 1] - goff[i];
    }
    return MG_OK;
}

namespace {

struct Offsets { std::vector<int64_t> g, s, p; };

// optional: all_mips.txt records of every region, written on the device and streamed to the caller in region order
struct TileRecords {
    const mg_record_meta *meta;
    const char *universal_middle;
    int first_index;
    int64_t *records_per_region;
    mg_text_sink sink;
    void *user;
};

int check_args(mg_ctx *ctx, int want, const mg_select_params *sp, const mg_tile_result *out)
{
    if (!ctx->has_cfg) { ctx->err = "mg_set_config has not been called"; return MG_ERR_NOCONFIG; }
    if ((want & MG_WANT_SVR) && !ctx->has_model) { ctx->err = "no SVR model loaded"; return MG_ERR_NOMODEL; }
    if (sp) {
        const int need = sp->method == 1 ? MG_WANT_SVR : MG_WANT_LOGISTIC;
        if (!(want & need)) { ctx->err = "mg_tile_regions: `want` lacks the score the selection method works on"; return MG_ERR_INVALID; }
        if (!out->scan_best || !out->pos_best) { ctx->err = "mg_tile_regions: scan_best / pos_best are required with selection parameters"; return MG_ERR_INVALID; }
    }
    if ((out->scan_best_logistic || out->logistic) && !(want & MG_WANT_LOGISTIC)) { ctx->err = "mg_tile_regions: logistic output without MG_WANT_LOGISTIC"; return MG_ERR_INVALID; }
    if ((out->scan_best_svr || out->svr) && !(want & MG_WANT_SVR)) { ctx->err = "mg_tile_regions: svr output without MG_WANT_SVR"; return MG_ERR_INVALID; }
    return MG_OK;
}

// regions[mine[*]] through one context, results at the caller's offsets
int tile_core(mg_ctx *ctx, const mg_region *regions, const std::vector<int> &mine, int want, const mg_select_params *sp, int64_t batch_cap,
              const mg_tile_result *out, const Offsets &off, const TileRecords *rec = nullptr)
{
    int next_index = rec ? rec->first_index : 0;
    std::vector<mg_record_meta> meta;
    std::vector<int64_t> per_region;
    if (batch_cap <= 0) batch_cap = (int64_t)1 << 26;
    want &= MG_WANT_LOGISTIC | MG_WANT_SVR;
    std::vector<mg_region> batch;
    std::vector<int64_t> sb, pb;
    std::vector<double> wl, ws;
    size_t k = 0;
    while (k < mine.size()) {
        batch.clear();
        const size_t k0 = k;
        int64_t cand = 0;
        while (k < mine.size()) {
            const int64_t g = off.g[mine[k] + 1] - off.g[mine[k]];
            if (!batch.empty() && cand + g > batch_cap) break;
            batch.push_back(regions[mine[k]]);
            cand += g;
            k++;
        }
        mg_panel *p = nullptr;
        int rc = mg_panel_create(ctx, batch.data(), (int)batch.size(), &p);
        if (rc != MG_OK) return rc;
        rc = mg_panel_score(ctx, p, want);
        if (rc == MG_OK && sp) {
            int64_t ns = 0, np = 0;
            for (size_t j = 0; j < batch.size(); j++) { ns += off.s[mine[k0 + j] + 1] - off.s[mine[k0 + j]]; np += off.p[mine[k0 + j] + 1] - off.p[mine[k0 + j]]; }
            sb.resize((size_t)std::max<int64_t>(2 * ns, 1));
            pb.resize((size_t)std::max<int64_t>(2 * np, 1));
            rc = mg_panel_select(ctx, p, sp, sb.data(), pb.data());
            const bool gl = out->scan_best_logistic != nullptr, gs = out->scan_best_svr != nullptr;
            if (rc == MG_OK && (gl || gs)) {
                if (gl) wl.resize(sb.size());
                if (gs) ws.resize(sb.size());
                rc = mg_panel_gather(ctx, p, sb.data(), 2 * ns, gl ? wl.data() : nullptr, gs ? ws.data() : nullptr);
            }
            if (rc == MG_OK) {
                // panel-global indices -> region-local ones, at the caller's offsets
                int64_t s_at = 0, p_at = 0;
                for (size_t j = 0; j < batch.size(); j++) {
                    const int r = mine[k0 + j];
                    const int64_t base = p->offsets[j], n_s = 2 * (off.s[r + 1] - off.s[r]), n_p = 2 * (off.p[r + 1] - off.p[r]);
                    int64_t *dst_s = out->scan_best + 2 * off.s[r], *dst_p = out->pos_best + 2 * off.p[r];
                    for (int64_t i = 0; i < n_s; i++) dst_s[i] = sb[(size_t)(s_at + i)] >= 0 ? sb[(size_t)(s_at + i)] - base : -1;
                    for (int64_t i = 0; i < n_p; i++) dst_p[i] = pb[(size_t)(p_at + i)] >= 0 ? pb[(size_t)(p_at + i)] - base : -1;
                    if (gl) memcpy(out->scan_best_logistic + 2 * off.s[r], &wl[(size_t)s_at], (size_t)n_s * 8);
                    if (gs) memcpy(out->scan_best_svr + 2 * off.s[r], &ws[(size_t)s_at], (size_t)n_s * 8);
                    s_at += n_s;
                    p_at += n_p;
                }
            }
        }
        if (rc == MG_OK && rec) {
            // `mine` is the identity here (one context), so the sub-batches -- and the text -- come in region order
            meta.assign(rec->meta + mine[k0], rec->meta + mine[k0] + batch.size());
            per_region.assign(batch.size(), 0);
            const int64_t got = mg_panel_format_enumerated(ctx, p, meta.data(), sp, rec->universal_middle, next_index, per_region.data(), rec->sink, rec->user);
            if (got < 0) rc = (int)got;
            for (size_t j = 0; j < batch.size() && rc == MG_OK; j++) {
                if (rec->records_per_region) rec->records_per_region[mine[k0 + j]] = per_region[j];
                next_index += (int)per_region[j];
            }
        }
        if (rc == MG_OK && (out->valid || out->logistic || out->svr)) {
            cudaError_t e = cudaSetDevice(ctx->device);
            for (size_t j = 0; j < batch.size() && e == cudaSuccess; j++) {
                const int r = mine[k0 + j];
                const int64_t a = p->offsets[j], m = p->offsets[j + 1] - a;
                if (m == 0) continue;
                if (out->valid) e = cudaMemcpyAsync(out->valid + off.g[r], p->d_valid + a, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream);
                if (out->logistic && e == cudaSuccess)
                    e = cudaMemcpyAsync(out->logistic + off.g[r], p->d_logistic + a, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream);
                if (out->svr && e == cudaSuccess) e = cudaMemcpyAsync(out->svr + off.g[r], p->d_svr + a, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { ctx->err = std::string("mg_tile_regions: result copy: ") + cudaGetErrorString(e); rc = MG_ERR_CUDA; }
        }
        mg_panel_destroy(p);
        if (rc != MG_OK) return rc;
    }
    return MG_OK;
}

int config_of(const mg_ctx *ctx, mg_config *c)
{
    const HostConfig &h = ctx->cfg;
    c->max_capture = h.max_capture; c->min_capture = h.min_capture; c->capture_increment = h.inc; c->max_mip_overlap = h.max_mip_overlap;
    c->n_pairs = (int)h.ext_len.size(); c->ext_len = h.ext_len.data(); c->lig_len = h.lig_len.data();
    c->n_oligo_sizes = (int)h.oligo_sizes.size(); c->oligo_sizes = h.oligo_sizes.data();
    return MG_OK;
}

int tile_multi(mg_ctx *const *ctxs, int n_ctx, const mg_region *regions, int n, int want, const mg_select_params *sp, int64_t batch_cap,
               const mg_tile_result *out, int64_t *out_offsets, const TileRecords *rec = nullptr)
{
    if (!ctxs || n_ctx <= 0 || n < 0 || (n > 0 && !regions) || !out) return MG_ERR_INVALID;
    for (int d = 0; d < n_ctx; d++) {
        if (!ctxs[d]) return MG_ERR_INVALID;
        const int rc = check_args(ctxs[d], want, sp, out);
        if (rc != MG_OK) return rc;
        if (d > 0) {  // same grid everywhere, or the caller's offsets would not hold
            const HostConfig &a = ctxs[0]->cfg, &b = ctxs[d]->cfg;
            if (a.max_capture != b.max_capture || a.min_capture != b.min_capture || a.inc != b.inc || a.max_mip_overlap != b.max_mip_overlap ||
                a.ext_len != b.ext_len || a.lig_len != b.lig_len || a.oligo_sizes != b.oligo_sizes) {
                ctxs[d]->err = "mg_*_multi: the contexts carry different configs";
                return MG_ERR_INVALID;
            }
        }
    }
    mg_config cfg;
    config_of(ctxs[0], &cfg);
    Offsets off;
    off.g.resize((size_t)n + 1); off.s.resize((size_t)n + 1); off.p.resize((size_t)n + 1);
    if (mg_tile_sizes(&cfg, regions, n, off.g.data(), off.s.data(), off.p.data()) != MG_OK) return MG_ERR_INVALID;
    if (out_offsets) memcpy(out_offsets, off.g.data(), ((size_t)n + 1) * sizeof(int64_t));
    std::vector<int> owner((size_t)std::max(n, 1), 0);
    if (n_ctx > 1 && mg_partition_regions(&cfg, regions, n, n_ctx, owner.data()) != MG_OK) return MG_ERR_INVALID;
    std::vector<std::vector<int>> mine((size_t)n_ctx);
    for (int i = 0; i < n; i++) mine[(size_t)owner[i]].push_back(i);
    if (n_ctx == 1) return tile_core(ctxs[0], regions, mine[0], want, sp, batch_cap, out, off, rec);
    std::vector<int> rcs((size_t)n_ctx, MG_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < n_ctx; d++)
        th.emplace_back([&, d]() { rcs[(size_t)d] = tile_core(ctxs[d], regions, mine[(size_t)d], want, sp, batch_cap, out, off); });
    for (auto &t : th) t.join();
    for (int d = 0; d < n_ctx; d++)
        if (rcs[(size_t)d] != MG_OK) return rcs[(size_t)d];
    return MG_OK;
}

}  // namespace

extern "C" int mg_tile_regions(mg_ctx *ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                               int64_t max_batch_candidates, mg_tile_result *out)
{
    mg_ctx *one[1] = {ctx};
    return tile_multi(one, 1, regions, n, want, sp, max_batch_candidates, out, nullptr);
}

extern "C" int mg_tile_regions_records(mg_ctx *ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                                       int64_t max_batch_candidates, mg_tile_result *out, const mg_record_meta *meta, const char *universal_middle,
                                       int first_index, int64_t *records_per_region, mg_text_sink sink, void *user)
{
    if (!ctx || !sp || !meta || !universal_middle || !sink) return MG_ERR_INVALID;
    TileRecords rec = {meta, universal_middle, first_index, records_per_region, sink, user};
    mg_ctx *one[1] = {ctx};
    return tile_multi(one, 1, regions, n, want, sp, max_batch_candidates, out, nullptr, &rec);
}

extern "C" int mg_tile_regions_multi(mg_ctx *const *ctxs, int n_ctx, const mg_region *regions, int n, int want, const mg_select_params *sp,
                                     int64_t max_batch_candidates, mg_tile_result *out)
{
    return tile_multi(ctxs, n_ctx, regions, n, want, sp, max_batch_candidates, out, nullptr);
}

extern "C" int mg_score_regions_multi(mg_ctx *const *ctxs, int n_ctx, const mg_region *regions, int n, int want, int64_t *out_offsets,
                                      uint8_t *valid, double *logistic, double *svr)
{
    mg_tile_result out;
    memset(&out, 0, sizeof out);
    out.valid = valid;
    out.logistic = (want & MG_WANT_LOGISTIC) ? logistic : nullptr;
    out.svr = (want & MG_WANT_SVR) ? svr : nullptr;
    int w = 0;
    if (out.logistic) w |= MG_WANT_LOGISTIC;
    if (out.svr) w |= MG_WANT_SVR;
    return tile_multi(ctxs, n_ctx, regions, n, w, nullptr, 0, &out, out_offsets);
}
