// mg_api.cu -- host side of the C-ABI declared in include/mipgen_b200.h.
//
// Owns device memory, the stream, the model upload and the launch sequence.  There is no
// CPU implementation of any scoring step in this file (or anywhere in the library): if the
// device is missing or a launch fails, the call returns MG_ERR_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <sstream>

#include "mg_common.cuh"

static std::string g_create_err;

// ---------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------
static cudaEvent_t ev_get(mg_ctx *ctx)
{
    if (!ctx->ev_free.empty()) {
        cudaEvent_t e = ctx->ev_free.back();
        ctx->ev_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;  // the launch still runs; only its timing is lost
    return e;
}

int mg_time_begin(mg_ctx *ctx, int which, long units)
{
    EventPair ep;
    ep.a = ev_get(ctx);
    ep.b = ev_get(ctx);
    ep.which = which;
    ep.units = units;
    ep.ok = ep.a && ep.b && cudaEventRecord(ep.a, ctx->stream) == cudaSuccess;
    ctx->ev_pending.push_back(ep);
    return MG_OK;
}

int mg_time_end(mg_ctx *ctx)
{
    EventPair &ep = ctx->ev_pending.back();
    if (ep.ok) ep.ok = cudaEventRecord(ep.b, ctx->stream) == cudaSuccess;
    return MG_OK;
}

static void drain_timings(mg_ctx *ctx)
{
    for (auto &ep : ctx->ev_pending) {
        float ms = 0.f;
        if (ep.ok && cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
            if (ep.which == TM_FEAT) { ctx->tm.ms_feat += ms; ctx->tm.launches_feat++; ctx->tm.candidates_feat += ep.units; }
            else if (ep.which == TM_SVR) { ctx->tm.ms_svr += ms; ctx->tm.launches_svr++; ctx->tm.candidates_svr += ep.units; }
            else { ctx->tm.ms_other += ms; ctx->tm.launches_other++; }
        }
        if (ep.a) ctx->ev_free.push_back(ep.a);
        if (ep.b) ctx->ev_free.push_back(ep.b);
    }
    ctx->ev_pending.clear();
}

// ---------------------------------------------------------------------------
// caching device allocator.  All work of a context is ordered on one stream, so a block freed
// by the host after the calls that used it were enqueued can be handed to later calls at once.
// ---------------------------------------------------------------------------
static const size_t kPoolCap = (size_t)12 << 30;

cudaError_t mg_dev_alloc(mg_ctx *ctx, void **out, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); i++)
        if (ctx->pool[i].bytes >= bytes && ctx->pool[i].bytes <= bytes + bytes / 2 + 4096 &&
            (best < 0 || ctx->pool[i].bytes < ctx->pool[best].bytes))
            best = (int)i;
    if (best >= 0) {
        CachedBlock b = ctx->pool[best];
        ctx->pool.erase(ctx->pool.begin() + best);
        ctx->pool_bytes -= b.bytes;
        ctx->live.push_back(b);
        *out = b.ptr;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {  // give cached memory back and retry once
        cudaGetLastError();
        mg_dev_trim(ctx);
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) ctx->live.push_back({*out, bytes});
    return e;
}

void mg_dev_free(mg_ctx *ctx, void *ptr)
{
    if (!ptr) return;
    for (size_t i = 0; i < ctx->live.size(); i++)
        if (ctx->live[i].ptr == ptr) {
            CachedBlock b = ctx->live[i];
            ctx->live.erase(ctx->live.begin() + i);
            if (ctx->pool_bytes + b.bytes <= kPoolCap) {
                ctx->pool.push_back(b);
                ctx->pool_bytes += b.bytes;
            } else {
                cudaStreamSynchronize(ctx->stream);
                cudaFree(ptr);
            }
            return;
        }
    cudaFree(ptr);  // not ours
}

void mg_dev_trim(mg_ctx *ctx)
{
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->pool) cudaFree(b.ptr);
    ctx->pool.clear();
    ctx->pool_bytes = 0;
}

// ---------------------------------------------------------------------------
// static tables
// ---------------------------------------------------------------------------
// mipgen.cpp:32 (data): the 44 strand-symmetric k-mers of long_range_content
static const char *const kFeatureMers[MG_NLRC] = {
    "A", "AA", "AAA", "AAC", "AAG", "AAT", "AC", "ACA", "ACC", "ACG", "AG", "AGA", "AGC", "AGG", "AGT",
    "AT", "ATA", "ATC", "ATG", "CAG", "CG", "CGG", "G", "GAC", "GAG", "GC", "GCG", "GG", "GGC", "GGG",
    "GTG", "TA", "TAA", "TAC", "TAG", "TC", "TCC", "TCG", "TG", "TGA", "TGC", "TGG", "TTC", "TTG"};

static int base_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

static int revcomp_code(int code, int k)
{
    int rc = 0;
    for (int i = 0; i < k; i++, code >>= 2) rc = (rc << 2) | (3 - (code & 3));
    return rc;
}

// Feature layout of get_parameters (SVMipv4.cpp:60-113): k-mers in trie pre-order
// ("A","AA","AAA",...), a G+C fraction right before "T", then the raw length.
static void emit_block(std::vector<uint32_t> &fd, int part, int depth, int slot_tri, int slot_di, int slot_mono, int slot_gc)
{
    for (int a = 0; a < 4; a++) {
        if (a == 3) fd.push_back(fd_pack(FK_RATIO, part, 0, slot_gc, slot_gc, 0));
        fd.push_back(fd_pack(FK_RATIO, part, 0, slot_mono + a, slot_mono + revcomp_code(a, 1), 0));
        for (int b = 0; b < 4; b++) {
            int di = a * 4 + b;
            fd.push_back(fd_pack(FK_RATIO, part, 1, slot_di + di, slot_di + revcomp_code(di, 2), 0));
            if (depth < 3) continue;
            for (int c = 0; c < 4; c++) {
                int tri = di * 4 + c;
                fd.push_back(fd_pack(FK_RATIO, part, 2, slot_tri + tri, slot_tri + revcomp_code(tri, 3), 0));
            }
        }
    }
    fd.push_back(fd_pack(FK_LEN, part, 0, 0, 0, 0));
}

static std::vector<uint32_t> build_feature_descriptors(bool window)
{
    // explicit front-end: per-warp count slots; window front-end: shared prefix-table rows
    // (tri 0..63, di 64..79, mono 80..83, G+C 84 -- the same tables serve arms and insert)
    std::vector<uint32_t> fd;
    if (window) emit_block(fd, 0, 2, 0, 64, 80, 84);
    else emit_block(fd, 0, 2, 0, SLOT_EXT_DI, SLOT_EXT_MONO, SLOT_EXT_GC);                      // 1..22
    for (int j = 0; j < MG_NLRC; j++) fd.push_back(fd_pack(FK_LRC, 0, 0, 0, 0, j));              // 23..66
    if (window) emit_block(fd, 1, 3, 0, 64, 80, 84);
    else emit_block(fd, 1, 3, SLOT_INS_TRI, SLOT_INS_DI, SLOT_INS_MONO, SLOT_INS_GC);           // 67..152
    if (window) emit_block(fd, 2, 2, 0, 64, 80, 84);
    else emit_block(fd, 2, 2, 0, SLOT_LIG_DI, SLOT_LIG_MONO, SLOT_LIG_GC);                      // 153..174
    for (int j = 0; j < 16; j++) fd.push_back(fd_pack(FK_JUNC, 2, 0, 0, 0, j));                  // 175..190
    fd.push_back(fd_pack(FK_COPY, 0, 0, 0, 0, 0));                                               // 191
    fd.push_back(fd_pack(FK_COPY, 2, 0, 0, 0, 1));                                               // 192
    return fd;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" const char *mg_version(void) { return "mipgen_b200 0.1 (sm_100a)"; }

extern "C" const char *mg_last_error(const mg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int mg_create(int device, mg_ctx **out)
{
    if (!out) return MG_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        g_create_err = std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (this library has no CPU fallback)";
        return MG_ERR_CUDA;
    }
    if (device < 0 || device >= count) { g_create_err = "device ordinal out of range"; return MG_ERR_INVALID; }
    mg_ctx *ctx = new mg_ctx();
    ctx->device = device;
    auto fail = [&](const char *what, cudaError_t err) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return MG_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        g_create_err = "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
        delete ctx;
        return MG_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);

    std::vector<uint32_t> fd = build_feature_descriptors(false), fdw = build_feature_descriptors(true);
    if (fd.size() != MG_NFEAT || fdw.size() != MG_NFEAT) { g_create_err = "internal: feature table size"; delete ctx; return MG_ERR_INVALID; }
    // K-feat's window kernel hard-codes which kind/part each feature index has; verify it against the table
    for (int f = 0; f < MG_NFEAT; f++) {
        uint32_t kind, part = 0;
        if (f < 21) { kind = FK_RATIO; part = 0; } else if (f == 21) { kind = FK_LEN; part = 0; }
        else if (f < 66) kind = FK_LRC;
        else if (f < 151) { kind = FK_RATIO; part = 1; } else if (f == 151) { kind = FK_LEN; part = 1; }
        else if (f < 173) { kind = FK_RATIO; part = 2; } else if (f == 173) { kind = FK_LEN; part = 2; }
        else if (f < 190) kind = FK_JUNC;
        else kind = FK_COPY;
        bool ok = (fdw[f] & 7) == kind && (kind == FK_LRC || kind == FK_JUNC || kind == FK_COPY || ((fdw[f] >> 3) & 3) == part);
        if (kind == FK_LRC) ok = ok && ((fdw[f] >> 23) & 255) == (uint32_t)(f - 22);
        if (kind == FK_JUNC) ok = ok && ((fdw[f] >> 23) & 255) == (uint32_t)(f - 174);
        if (kind == FK_COPY) ok = ok && ((fdw[f] >> 23) & 255) == (uint32_t)(f - 190);
        if (!ok) { g_create_err = "internal: feature layout does not match K-feat's specialisation"; delete ctx; return MG_ERR_INVALID; }
    }
    if ((e = cudaMalloc(&ctx->d_fdesc, MG_NFEAT * sizeof(uint32_t))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_fdesc_win, MG_NFEAT * sizeof(uint32_t))) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMemcpy(ctx->d_fdesc, fd.data(), MG_NFEAT * sizeof(uint32_t), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_fdesc_win, fdw.data(), MG_NFEAT * sizeof(uint32_t), cudaMemcpyHostToDevice);
    // log10 of copy numbers 0..100 from the host libm, so features 191/192 are bit-identical to glibc
    double logtab[102];
    for (int i = 0; i <= 100; i++) logtab[i] = log10((double)i);
    logtab[101] = 2.0;
    if ((e = cudaMalloc(&ctx->d_logcopy, sizeof logtab)) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMemcpy(ctx->d_logcopy, logtab, sizeof logtab, cudaMemcpyHostToDevice);
    double e2tab[64];
    for (int j = 0; j < 64; j++) e2tab[j] = exp2((double)j / 64.0);
    if ((e = cudaMalloc(&ctx->d_exp2tab, sizeof e2tab)) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMemcpy(ctx->d_exp2tab, e2tab, sizeof e2tab, cudaMemcpyHostToDevice);
    double e2tab256[256];
    for (int j = 0; j < 256; j++) e2tab256[j] = exp2((double)j / 256.0);
    if ((e = cudaMalloc(&ctx->d_exp2tab256, sizeof e2tab256)) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMemcpy(ctx->d_exp2tab256, e2tab256, sizeof e2tab256, cudaMemcpyHostToDevice);
    if ((e = cudaMalloc(&ctx->d_cfg, sizeof(DevConfig))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_work, 4 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc", e);
    cudaMemset(ctx->d_work, 0, 4 * sizeof(unsigned long long));
    if ((e = cudaMalloc(&ctx->d_fact, sizeof(DevFact))) != cudaSuccess) return fail("cudaMalloc", e);
    uint8_t lk[MG_NLRC], lc[MG_NLRC];
    for (int i = 0; i < MG_NLRC; i++) {
        int k = (int)strlen(kFeatureMers[i]), code = 0;
        for (int j = 0; j < k; j++) code = code * 4 + base_code(kFeatureMers[i][j]);
        lk[i] = (uint8_t)k;
        lc[i] = (uint8_t)code;
    }
    if (mg_upload_lrc_tables(ctx, lk, lc) != MG_OK || launch_svr_setup(ctx) != MG_OK || launch_feat_setup(ctx) != MG_OK ||
        launch_fact_setup(ctx) != MG_OK || launch_tc_setup(ctx) != MG_OK) {
        g_create_err = ctx->err;
        delete ctx;
        return MG_ERR_CUDA;
    }
    *out = ctx;
    return MG_OK;
}

static void free_model(mg_ctx *ctx)
{
    cudaFree(ctx->d_fact_blob);
    ctx->d_fact_blob = nullptr;
    cudaFree(ctx->d_tc_img); cudaFree(ctx->d_tc_centre); cudaFree(ctx->d_tc_expc);
    ctx->d_tc_img = nullptr; ctx->d_tc_centre = ctx->d_tc_expc = nullptr;
    ctx->tc_ok = false;
    cudaFree(ctx->d_sv); cudaFree(ctx->d_sv_tiled); cudaFree(ctx->d_ss); cudaFree(ctx->d_alpha); cudaFree(ctx->d_tail);
    ctx->d_sv = ctx->d_sv_tiled = ctx->d_ss = ctx->d_alpha = ctx->d_tail = nullptr;
    ctx->has_model = false;
}

extern "C" void mg_destroy(mg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drain_timings(ctx);
    mg_dev_trim(ctx);
    for (auto e : ctx->ev_free) cudaEventDestroy(e);
    if (ctx->sw_a) { cudaEventDestroy(ctx->sw_a); cudaEventDestroy(ctx->sw_b); }
    free_model(ctx);
    cudaFree(ctx->d_x); cudaFree(ctx->d_rows); cudaFree(ctx->d_fdesc); cudaFree(ctx->d_fdesc_win); cudaFree(ctx->d_logcopy); cudaFree(ctx->d_exp2tab); cudaFree(ctx->d_exp2tab256); cudaFree(ctx->d_cfg); cudaFree(ctx->d_fact); cudaFree(ctx->d_work);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int mg_sync(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MG_OK;
}

extern "C" int mg_timer_start(mg_ctx *ctx)
{
    if (!ctx) return MG_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->sw_a) { CUDA_TRY(ctx, cudaEventCreate(&ctx->sw_a)); CUDA_TRY(ctx, cudaEventCreate(&ctx->sw_b)); }
    CUDA_TRY(ctx, cudaEventRecord(ctx->sw_a, ctx->stream));
    return MG_OK;
}

extern "C" int mg_timer_stop(mg_ctx *ctx, double *ms)
{
    if (!ctx || !ctx->sw_a) return MG_ERR_INVALID;
    CUDA_TRY(ctx, cudaEventRecord(ctx->sw_b, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->sw_b));
    float f = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->sw_a, ctx->sw_b));
    if (ms) *ms = f;
    return MG_OK;
}

extern "C" int mg_reset_timings(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    drain_timings(ctx);
    memset(&ctx->tm, 0, sizeof ctx->tm);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_work, 0, 3 * sizeof(unsigned long long), ctx->stream));
    return MG_OK;
}

extern "C" int mg_get_timings(mg_ctx *ctx, mg_timings *out)
{
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    drain_timings(ctx);
    if (out) {
        *out = ctx->tm;
        // the factored K-svr counts the work of the tasks that really ran on the device
        unsigned long long w[3] = {0, 0, 0};
        CUDA_TRY(ctx, cudaMemcpy(w, ctx->d_work, sizeof w, cudaMemcpyDeviceToHost));
        out->svr_dmma += (double)w[0]; out->svr_exp += (double)w[1]; out->svr_gather += (double)w[2];
    }
    return MG_OK;
}

int mg_host_config_from(const mg_config *c, HostConfig &h, std::string &err)
{
    if (!c) return MG_ERR_INVALID;
    if (c->n_pairs <= 0 || c->n_pairs > MG_MAX_PAIRS || c->n_oligo_sizes > MG_MAX_OLIGO || c->n_oligo_sizes < 0 ||
        c->max_capture < c->min_capture || c->min_capture <= 0 || !c->ext_len || !c->lig_len) {
        err = "mg_config: bad capture range or arm pair count";
        return MG_ERR_INVALID;
    }
    h.max_capture = c->max_capture;
    h.min_capture = c->min_capture;
    h.inc = c->capture_increment == 0 ? 1 : c->capture_increment;  // mipgen.cpp:274
    if (h.inc < 0) { err = "mg_config: negative capture_increment"; return MG_ERR_INVALID; }
    h.max_mip_overlap = c->max_mip_overlap;
    h.ext_len.assign(c->ext_len, c->ext_len + c->n_pairs);
    h.lig_len.assign(c->lig_len, c->lig_len + c->n_pairs);
    h.oligo_sizes.clear();
    if (c->n_oligo_sizes > 0) h.oligo_sizes.assign(c->oligo_sizes, c->oligo_sizes + c->n_oligo_sizes);
    h.n_cap = (h.max_capture - h.min_capture) / h.inc + 1;
    h.max_sum = 0;
    h.min_sum = 1 << 30;
    for (int i = 0; i < c->n_pairs; i++) {
        int s = h.ext_len[i] + h.lig_len[i];
        if (h.ext_len[i] <= 0 || h.lig_len[i] <= 0 || s >= h.min_capture) {
            err = "mg_config: arm lengths must be positive and sum below min_capture";
            return MG_ERR_INVALID;
        }
        h.max_sum = std::max(h.max_sum, s);
        h.min_sum = std::min(h.min_sum, s);
        // the tile loop walks arm sums in descending order, each sum's list once (mipgen.cpp:431-438); the replay of its
        // score-dependent skips (mg_tile_replay, K-condense) relies on pairs grouped by sum, sums strictly descending
        if (i > 0 && s > h.ext_len[i - 1] + h.lig_len[i - 1]) {
            err = "mg_config: arm pairs must be grouped by arm sum with sums in descending order (mipgen.cpp:431-438)";
            return MG_ERR_INVALID;
        }
    }
    return MG_OK;
}

extern "C" int mg_set_config(mg_ctx *ctx, const mg_config *c)
{
    if (!ctx || !c) return MG_ERR_INVALID;
    HostConfig h;
    int hrc = mg_host_config_from(c, h, ctx->err);
    if (hrc != MG_OK) return hrc;
    DevConfig *d = new DevConfig();
    memset(d, 0, sizeof *d);
    d->max_capture = h.max_capture; d->min_capture = h.min_capture; d->inc = h.inc; d->max_mip_overlap = h.max_mip_overlap;
    d->n_cap = h.n_cap; d->n_pairs = c->n_pairs; d->max_sum = h.max_sum; d->min_sum = h.min_sum;
    d->n_oligo = (int)h.oligo_sizes.size();
    d->max_arm = 0;
    d->min_arm = 1 << 30;
    for (int i = 0; i < c->n_pairs; i++) {
        d->ext_len[i] = h.ext_len[i]; d->lig_len[i] = h.lig_len[i];
        d->max_arm = std::max(d->max_arm, std::max(h.ext_len[i], h.lig_len[i]));
        d->min_arm = std::min(d->min_arm, std::min(h.ext_len[i], h.lig_len[i]));
    }
    h.max_arm = d->max_arm;
    h.min_arm = d->min_arm;
    for (size_t i = 0; i < h.oligo_sizes.size(); i++) d->oligo_sizes[i] = h.oligo_sizes[i];
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaMemcpy(ctx->d_cfg, d, sizeof *d, cudaMemcpyHostToDevice);
    delete d;
    if (e != cudaSuccess) { ctx->err = std::string("config upload: ") + cudaGetErrorString(e); return MG_ERR_CUDA; }
    ctx->cfg = h;
    ctx->has_cfg = true;
    ctx->cfg_serial++;

    // factored SVR: distinct arm lengths / arm sums, and the largest window whose tables fit in shared memory
    ctx->fact_ok = false;
    {
        DevFact *f = new DevFact();
        memset(f, 0, sizeof *f);
        std::vector<int> exts, ligs, sums;
        bool ok = h.max_sum - h.min_sum < FACT_MAX_SPAN;
        for (int i = 0; i < c->n_pairs && ok; i++) {
            if (h.ext_len[i] >= FACT_MAX_LEN || h.lig_len[i] >= FACT_MAX_LEN) ok = false;
            exts.push_back(h.ext_len[i]); ligs.push_back(h.lig_len[i]); sums.push_back(h.ext_len[i] + h.lig_len[i]);
        }
        auto uniq = [](std::vector<int> &v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
        uniq(exts); uniq(ligs); uniq(sums);
        if (ok) {
            for (int i = 0; i < FACT_MAX_LEN; i++) f->ext_idx[i] = f->lig_idx[i] = -1;
            for (int i = 0; i < FACT_MAX_SPAN; i++) f->sum_idx[i] = -1;
            for (size_t i = 0; i < exts.size(); i++) f->ext_idx[exts[i]] = (int)i;
            for (size_t i = 0; i < ligs.size(); i++) f->lig_idx[ligs[i]] = (int)i;
            for (size_t i = 0; i < sums.size(); i++) f->sum_idx[sums[i] - h.min_sum] = (int)i;
            for (size_t i = 0; i < exts.size(); i++) f->ext_of[i] = exts[i];
            for (size_t i = 0; i < ligs.size(); i++) f->lig_of[i] = ligs[i];
            for (size_t i = 0; i < sums.size(); i++) f->sum_of[i] = sums[i];
            for (int i = 0; i < c->n_pairs; i++) { f->pair_e[i] = h.ext_len[i]; f->pair_l[i] = h.lig_len[i]; }
            f->n_pairs = c->n_pairs; f->n_cap = h.n_cap; f->n_ext = (int)exts.size(); f->n_lig = (int)ligs.size();
            f->n_sums = (int)sums.size(); f->min_sum = h.min_sum; f->max_sum = h.max_sum;
            const int dsum = h.max_sum - h.min_sum;
            auto pad8 = [](int v) { return (v + 15) & ~15; };  // the factored kernel's work units are 16 rows
            for (int W = 8; W >= 1 && !ctx->fact_ok; W /= 2) {
                if (W * c->n_pairs > FACT_CPT * FACT_GATHER_WARPS * 32) continue;  // FACT_CPT candidates per gather thread
                const int RA0 = pad8(W * f->n_ext), RA1 = pad8(W * f->n_lig);
                const int RQ0 = pad8((W + dsum) * f->n_lig), RQ1 = pad8((W + dsum) * f->n_ext), RI = pad8(W * f->n_sums);
                f->cap_FA = std::max(RA0, RA1) * FACT_LD_ARM;
                f->cap_FQ = std::max(RQ0, RQ1) * FACT_LD_ARM;
                f->cap_FI = RI * FACT_LD_INS;
                f->cap_R = std::max(RA0 + RQ0, RA1 + RQ1) + RI;
                if (f->cap_R / 16 > FACT_MATH_WARPS * 24) continue;  // per-warp work-unit lists
                size_t doubles = (size_t)f->cap_FA + f->cap_FQ + f->cap_FI + f->cap_R + 2 * (size_t)f->cap_R * (FACT_C + 1) + 2 * FACT_BLOB +
                                 2 * FACT_C + 64;
                size_t bytes = doubles * 8 + 64 + (size_t)f->cap_R * 8 + FACT_MATH_WARPS * 25 + 64;  // + mbarriers, rep[] and jc[] ints, unit lists
                if (bytes <= FACT_SMEM_LIMIT) { f->W = W; ctx->fact_ok = true; ctx->fact_W = W; ctx->fact_smem = bytes; f->blob_doubles = FACT_ROWS_DOUBLES(*f); }
            }
        }
        if (ctx->fact_ok) { e = cudaMemcpy(ctx->d_fact, f, sizeof *f, cudaMemcpyHostToDevice); ctx->h_fact = *f; }
        delete f;
        if (e != cudaSuccess) { ctx->err = std::string("factored config upload: ") + cudaGetErrorString(e); return MG_ERR_CUDA; }
    }
    return MG_OK;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
static int ensure_x(mg_ctx *ctx, int64_t rows);

static int upload_model(mg_ctx *ctx, const std::vector<double> &dense, const std::vector<double> &tail, const std::vector<double> &alpha,
                        int n_sv, double gamma, double rho)
{
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_model(ctx);
    int pad = std::max(SVR_BN, (n_sv + SVR_BN - 1) / SVR_BN * SVR_BN);
    std::vector<double> sv((size_t)pad * MG_NFEAT, 0.0), ss(pad, 0.0), al(pad, 0.0), tl(pad, 0.0);
    for (int i = 0; i < n_sv; i++) {
        double s = 0.0;
        for (int k = 0; k < MG_NFEAT; k++) {
            double v = dense[(size_t)i * MG_NFEAT + k];
            sv[(size_t)i * MG_NFEAT + k] = v;
            s += v * v;
        }
        ss[i] = s + tail[i];
        tl[i] = tail[i];
        al[i] = alpha[i];
    }
    // slab image: [chunk][slab][row][SVR_LDB], the last SVR_LDB-SVR_BK doubles of a row are padding
    const int n_slabs = MG_NFEAT / SVR_BK;
    std::vector<double> tiled((size_t)(pad / SVR_BN) * n_slabs * SVR_BN * SVR_LDB, 0.0);
    for (int i = 0; i < pad; i++)
        for (int k = 0; k < MG_NFEAT; k++)
            tiled[(((size_t)(i / SVR_BN) * n_slabs + k / SVR_BK) * SVR_BN + i % SVR_BN) * SVR_LDB + k % SVR_BK] = sv[(size_t)i * MG_NFEAT + k];
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_sv_tiled, tiled.size() * 8));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_sv_tiled, tiled.data(), tiled.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_sv, sv.size() * 8));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_ss, ss.size() * 8));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_alpha, al.size() * 8));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_tail, tl.size() * 8));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_tail, tl.data(), tl.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_sv, sv.data(), sv.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_ss, ss.data(), ss.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_alpha, al.data(), al.size() * 8, cudaMemcpyHostToDevice));
    // factored kernel: per chunk of FACT_C support vectors, the ext / lig / insert blocks in the padded
    // shared-memory layout, followed by the blocks' squared norms (one bulk copy per chunk)
    {
        const int n_chunks = pad / FACT_C;
        std::vector<double> blob((size_t)n_chunks * FACT_BLOB, 0.0);
        for (int i = 0; i < pad; i++) {
            double *b = &blob[(size_t)(i / FACT_C) * FACT_BLOB];
            const int r = i % FACT_C;
            const double *srow = &sv[(size_t)i * MG_NFEAT];
            // the blocks carry the factor 2 gamma and the norm tables the factor -gamma, so the kernel's
            // contraction  -g ||x||^2 - g ||s||^2 + (2 g s) . x  ends on the exponent itself
            const double g2 = 2.0 * gamma;
            double se = 0, sl = 0, si = 0;
            for (int k = 0; k < 22; k++) { b[FACT_OFF_EXT + r * FACT_LD_ARM + k] = g2 * srow[k]; se += srow[k] * srow[k]; }
            b[FACT_OFF_EXT + r * FACT_LD_ARM + 22] = g2 * srow[190]; se += srow[190] * srow[190];
            for (int k = 0; k < 22; k++) { b[FACT_OFF_LIG + r * FACT_LD_ARM + k] = g2 * srow[152 + k]; sl += srow[152 + k] * srow[152 + k]; }
            b[FACT_OFF_LIG + r * FACT_LD_ARM + 22] = g2 * srow[191]; sl += srow[191] * srow[191];
            // junction one-hot (features 175..190): ||onehot(jc) - s||^2 = sum_j s_j^2 + (1 - 2 s_jc); the table row of
            // junction code jc holds the whole ligation-role term, row 16 the one for "no junction bit set"
            for (int k = 0; k < 16; k++) sl += srow[174 + k] * srow[174 + k];
            for (int k = 0; k < 16; k++) b[FACT_OFF_JT + k * FACT_C + r] = -gamma * (sl + (1.0 - 2.0 * srow[174 + k]));
            b[FACT_OFF_JT + 16 * FACT_C + r] = -gamma * sl;
            for (int k = 0; k < 86; k++) { b[FACT_OFF_INS + r * FACT_LD_INS + k] = g2 * srow[66 + k]; si += srow[66 + k] * srow[66 + k]; }
            // spare columns: the rows carry -gamma ||row||^2 there (k_svr_fact phase 1)
            b[FACT_OFF_EXT + r * FACT_LD_ARM + FACT_K_ARM - 1] = 1.0;
            b[FACT_OFF_LIG + r * FACT_LD_ARM + FACT_K_ARM - 1] = 1.0;
            b[FACT_OFF_INS + r * FACT_LD_INS + FACT_K_INS - 2] = 1.0;
            b[FACT_OFF_SS + r] = -gamma * se; b[FACT_OFF_SS + FACT_C + r] = -gamma * sl; b[FACT_OFF_SS + 2 * FACT_C + r] = -gamma * si;
        }
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_fact_blob, blob.size() * 8));
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_fact_blob, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice));
    }
    // tensor-core kernel: centred FP16 hi/lo operand images (k_svr_tc.cu)
    {
        std::vector<uint8_t> img;
        std::vector<double> centre, exp_c;
        ctx->tc_ok = mg_tc_prepare_model(sv, pad, n_sv, gamma, img, centre, exp_c);
        if (ctx->tc_ok) {
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_tc_img, img.size()));
            CUDA_TRY(ctx, cudaMemcpy(ctx->d_tc_img, img.data(), img.size(), cudaMemcpyHostToDevice));
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_tc_centre, centre.size() * 8));
            CUDA_TRY(ctx, cudaMemcpy(ctx->d_tc_centre, centre.data(), centre.size() * 8, cudaMemcpyHostToDevice));
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_tc_expc, exp_c.size() * 8));
            CUDA_TRY(ctx, cudaMemcpy(ctx->d_tc_expc, exp_c.data(), exp_c.size() * 8, cudaMemcpyHostToDevice));
        }
    }
    ctx->n_sv = n_sv; ctx->n_sv_pad = pad; ctx->gamma = gamma; ctx->rho = rho;
    ctx->has_model = true;
    // SVR value of the all-zero vector (what an invalid candidate scores, SVMipv4.cpp:63-68), from the dense kernel
    {
        int rc = ensure_x(ctx, SVR_BM);
        if (rc != MG_OK) return rc;
        double *d_z = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&d_z, SVR_BM * 8));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_x, 0, (size_t)SVR_BM * MG_NFEAT * 8, ctx->stream));
        rc = launch_svr(ctx, ctx->d_x, SVR_BM, nullptr, d_z);
        if (rc == MG_OK) {
            cudaError_t e2 = cudaMemcpyAsync(&ctx->zero_score, d_z, 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
            if (e2 != cudaSuccess) { ctx->err = cudaGetErrorString(e2); rc = MG_ERR_CUDA; }
        }
        cudaFree(d_z);
        if (rc != MG_OK) return rc;
    }
    return MG_OK;
}

extern "C" int mg_set_svr_model(mg_ctx *ctx, const double *sv, const double *alpha, int n_sv, int n_feat, double gamma, double rho)
{
    if (!ctx || !sv || !alpha || n_sv < 0 || n_feat <= 0 || n_feat > MG_NFEAT) return MG_ERR_INVALID;
    std::vector<double> dense((size_t)std::max(n_sv, 1) * MG_NFEAT, 0.0), tail(std::max(n_sv, 1), 0.0), al(alpha, alpha + n_sv);
    for (int i = 0; i < n_sv; i++)
        for (int k = 0; k < n_feat; k++) dense[(size_t)i * MG_NFEAT + k] = sv[(size_t)i * n_feat + k];
    return upload_model(ctx, dense, tail, al, n_sv, gamma, rho);
}

// libsvm text model (format written by svm_save_model, svm.cpp:2644-2736; read by
// svm_load_model, svm.cpp:2759-2973): header "key value" lines up to "SV", then one
// line per support vector: "<coef> idx:val idx:val ...", absent idx meaning 0.
extern "C" int mg_load_svr_model(mg_ctx *ctx, const char *path)
{
    if (!ctx || !path) return MG_ERR_INVALID;
    std::ifstream in(path, std::ios::binary);
    if (!in) { ctx->err = std::string("cannot open model file ") + path; return MG_ERR_MODEL; }
    std::string line, svm_type, kernel_type;
    double gamma = 0, rho = 0;
    int nr_class = 0, total_sv = -1;
    bool in_sv = false;
    while (std::getline(in, line)) {
        std::istringstream is(line);
        std::string key;
        if (!(is >> key)) continue;
        if (key == "SV") { in_sv = true; break; }
        if (key == "svm_type") is >> svm_type;
        else if (key == "kernel_type") is >> kernel_type;
        else if (key == "gamma") { std::string v; is >> v; gamma = strtod(v.c_str(), nullptr); }
        else if (key == "rho") { std::string v; is >> v; rho = strtod(v.c_str(), nullptr); }
        else if (key == "nr_class") is >> nr_class;
        else if (key == "total_sv") is >> total_sv;
        else if (key == "degree" || key == "coef0" || key == "label" || key == "nr_sv" || key == "probA" || key == "probB") {}
        else { ctx->err = "unknown text in model file: [" + key + "]"; return MG_ERR_MODEL; }
    }
    if (!in_sv || total_sv < 0) { ctx->err = "model file has no SV section / total_sv"; return MG_ERR_MODEL; }
    if ((svm_type != "epsilon_svr" && svm_type != "nu_svr") || kernel_type != "rbf" || nr_class != 2) {
        ctx->err = "model is not an RBF epsilon_svr/nu_svr (this path implements only what mipgen uses)";
        return MG_ERR_MODEL;
    }
    std::vector<double> dense((size_t)std::max(total_sv, 1) * MG_NFEAT, 0.0), tail(std::max(total_sv, 1), 0.0), alpha(std::max(total_sv, 1), 0.0);
    for (int i = 0; i < total_sv; i++) {
        if (!std::getline(in, line)) { ctx->err = "model file truncated in the SV section"; return MG_ERR_MODEL; }
        const char *p = line.c_str();
        char *end;
        alpha[i] = strtod(p, &end);
        p = end;
        for (;;) {
            while (*p == ' ' || *p == '\t') p++;
            if (*p == 0 || *p == '\n' || *p == '\r') break;
            long idx = strtol(p, &end, 10);
            if (end == p || *end != ':') break;
            p = end + 1;
            double val = strtod(p, &end);
            if (end == p) break;
            p = end;
            if (idx >= 1 && idx <= MG_NFEAT) dense[(size_t)i * MG_NFEAT + (idx - 1)] = val;
            else tail[i] += val * val;  // feature the 192-vector never carries: contributes s^2
        }
    }
    alpha.resize(total_sv);
    return upload_model(ctx, dense, tail, alpha, total_sv, gamma, rho);
}

extern "C" int mg_set_svr_mode(mg_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 3) return MG_ERR_INVALID;
    ctx->svr_mode = mode;
    return MG_OK;
}

extern "C" int mg_svr_tensor_core_available(const mg_ctx *ctx) { return ctx && ctx->has_model && ctx->tc_ok; }

extern "C" int mg_svr_factored_available(const mg_ctx *ctx) { return ctx && ctx->fact_ok ? ctx->fact_W : 0; }

extern "C" int mg_model_info(const mg_ctx *ctx, int *n_sv, double *gamma, double *rho)
{
    if (!ctx || !ctx->has_model) return MG_ERR_NOMODEL;
    if (n_sv) *n_sv = ctx->n_sv;
    if (gamma) *gamma = ctx->gamma;
    if (rho) *rho = ctx->rho;
    return MG_OK;
}

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
static const int64_t kMaxChunkRows = 1 << 22;  // 4 M candidates = 6.4 GB of feature rows in flight (one launch tail per chunk)

static int ensure_x(mg_ctx *ctx, int64_t rows)
{
    rows = (rows + SVR_BM - 1) / SVR_BM * SVR_BM;
    if ((size_t)rows <= ctx->x_rows_cap) return MG_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_x);
    ctx->d_x = nullptr;
    ctx->x_rows_cap = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_x, (size_t)rows * MG_NFEAT * 8));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_x, 0, (size_t)rows * MG_NFEAT * 8, ctx->stream));
    ctx->x_rows_cap = (size_t)rows;
    return MG_OK;
}

// row tables of the factored-SVR work items in flight: at most kMaxRowsBytes of them at a time
static const size_t kMaxRowsBytes = (size_t)2 << 30;

static int ensure_rows(mg_ctx *ctx, size_t items)
{
    const size_t bytes = items * (size_t)ctx->h_fact.blob_doubles * 8;   // the stride changes with the configuration
    if (bytes <= ctx->rows_cap) return MG_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_rows);
    ctx->d_rows = nullptr;
    ctx->rows_cap = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_rows, bytes));
    ctx->rows_cap = bytes;
    return MG_OK;
}

// ---------------------------------------------------------------------------
// svm_predict on dense rows
// ---------------------------------------------------------------------------
static int predict_rows(mg_ctx *ctx, const double *x, long n, long ld, double *out, bool direct)
{
    if (!ctx || !x || !out || n < 0 || ld < MG_NFEAT) return MG_ERR_INVALID;
    if (!ctx->has_model) { ctx->err = "no SVR model loaded"; return MG_ERR_NOMODEL; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *d_out = nullptr;
    for (long i0 = 0; i0 < n; i0 += kMaxChunkRows) {
        long m = std::min<long>(kMaxChunkRows, n - i0);
        int rc = ensure_x(ctx, m);
        if (rc != MG_OK) return rc;
        if (!d_out) CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&d_out, (size_t)std::min<long>(n, kMaxChunkRows) * 8));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->d_x, MG_NFEAT * 8, x + i0 * ld, (size_t)ld * 8, MG_NFEAT * 8, (size_t)m,
                                         cudaMemcpyHostToDevice, ctx->stream));
        rc = direct ? launch_svr_direct(ctx, ctx->d_x, m, MG_NFEAT, d_out) : launch_svr(ctx, ctx->d_x, m, nullptr, d_out);
        if (rc != MG_OK) { mg_dev_free(ctx, d_out); return rc; }
        cudaError_t e = cudaMemcpyAsync(out + i0, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("mg_svr_predict: ") + cudaGetErrorString(e); mg_dev_free(ctx, d_out); return MG_ERR_CUDA; }
    }
    mg_dev_free(ctx, d_out);
    return MG_OK;
}

extern "C" int mg_svr_predict(mg_ctx *ctx, const double *x, long n, long ld, double *out) { return predict_rows(ctx, x, n, ld, out, false); }
extern "C" int mg_svr_predict_direct(mg_ctx *ctx, const double *x, long n, long ld, double *out) { return predict_rows(ctx, x, n, ld, out, true); }

// ---------------------------------------------------------------------------
// long-range content
// ---------------------------------------------------------------------------
extern "C" int mg_long_range_content(mg_ctx *ctx, const char *ext_seq, int n, int denom, double out[MG_NLRC])
{
    if (!ctx || !ext_seq || n < 0 || !out) return MG_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    char *d_a = nullptr;
    uint8_t *d_c = nullptr;
    double *d_o = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_a, std::max(n, 1)));
    CUDA_TRY(ctx, cudaMalloc(&d_c, std::max(n, 1)));
    CUDA_TRY(ctx, cudaMalloc(&d_o, MG_NLRC * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_a, ext_seq, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_encode(ctx, d_a, d_c, n);
    if (rc == MG_OK) rc = launch_lrc(ctx, d_c, n, denom, d_o);
    if (rc == MG_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_o, MG_NLRC * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = MG_ERR_CUDA; }
    }
    cudaFree(d_a); cudaFree(d_c); cudaFree(d_o);
    return rc;
}

// ---------------------------------------------------------------------------
// explicit candidates
// ---------------------------------------------------------------------------
extern "C" int mg_score_candidates(mg_ctx *ctx, const mg_candidate *cands, long n, const double *lrc, int want, double *logistic,
                                   double *svr, double *features)
{
    if (!ctx || (!cands && n > 0) || n < 0) return MG_ERR_INVALID;
    const bool w_log = (want & MG_WANT_LOGISTIC) && logistic, w_svr = (want & MG_WANT_SVR) && svr, w_feat = (want & MG_WANT_FEATURES) && features;
    if (w_svr && !ctx->has_model) { ctx->err = "no SVR model loaded"; return MG_ERR_NOMODEL; }
    if (n == 0) return MG_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = MG_OK;
    for (long i0 = 0; i0 < n && rc == MG_OK; i0 += kMaxChunkRows) {
        const long m = std::min<long>(kMaxChunkRows, n - i0);
        std::vector<DevCand> dc(m);
        size_t total = 0;
        for (long i = 0; i < m; i++) {
            const mg_candidate &c = cands[i0 + i];
            if (c.ext_n < 0 || c.lig_n < 0 || c.tgt_n < 0 || (!c.ext && c.ext_n) || (!c.lig && c.lig_n) || (!c.tgt && c.tgt_n)) {
                ctx->err = "mg_score_candidates: bad string in candidate";
                return MG_ERR_INVALID;
            }
            total += (size_t)c.ext_n + c.lig_n + c.tgt_n;
        }
        std::vector<char> ascii(std::max<size_t>(total, 1));
        size_t off = 0;
        for (long i = 0; i < m; i++) {
            const mg_candidate &c = cands[i0 + i];
            DevCand &d = dc[i];
            d.ext_off = (int64_t)off; memcpy(&ascii[off], c.ext, c.ext_n); off += c.ext_n;
            d.lig_off = (int64_t)off; memcpy(&ascii[off], c.lig, c.lig_n); off += c.lig_n;
            d.tgt_off = (int64_t)off; memcpy(&ascii[off], c.tgt, c.tgt_n); off += c.tgt_n;
            d.ext_n = c.ext_n; d.lig_n = c.lig_n; d.tgt_n = c.tgt_n;
            d.ext_len = c.ext_len; d.lig_len = c.lig_len; d.scan_size = c.scan_size;
            d.ext_copy = c.ext_copy; d.lig_copy = c.lig_copy; d.pad = 0;
        }
        char *d_a = nullptr; uint8_t *d_c = nullptr; DevCand *d_dc = nullptr; double *d_lrc = nullptr, *d_log = nullptr, *d_svr = nullptr;
        auto cleanup = [&]() { cudaFree(d_a); cudaFree(d_c); cudaFree(d_dc); cudaFree(d_lrc); cudaFree(d_log); cudaFree(d_svr); };
#define TRY_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return MG_ERR_CUDA; } } while (0)
        TRY_FREE(cudaMalloc(&d_a, ascii.size()));
        TRY_FREE(cudaMalloc(&d_c, ascii.size()));
        TRY_FREE(cudaMalloc(&d_dc, (size_t)m * sizeof(DevCand)));
        TRY_FREE(cudaMemcpyAsync(d_a, ascii.data(), ascii.size(), cudaMemcpyHostToDevice, ctx->stream));
        TRY_FREE(cudaMemcpyAsync(d_dc, dc.data(), (size_t)m * sizeof(DevCand), cudaMemcpyHostToDevice, ctx->stream));
        if (lrc) {
            TRY_FREE(cudaMalloc(&d_lrc, (size_t)m * MG_NLRC * 8));
            TRY_FREE(cudaMemcpyAsync(d_lrc, lrc + i0 * MG_NLRC, (size_t)m * MG_NLRC * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (w_log) TRY_FREE(cudaMalloc(&d_log, (size_t)m * 8));
        if (w_svr) TRY_FREE(cudaMalloc(&d_svr, (size_t)m * 8));
        const bool need_x = w_svr || w_feat;
        if (need_x && (rc = ensure_x(ctx, m)) != MG_OK) { cleanup(); return rc; }
        rc = launch_encode(ctx, d_a, d_c, (int64_t)ascii.size());
        if (rc == MG_OK) rc = launch_feat_explicit(ctx, d_dc, d_c, d_lrc, m, d_log, need_x ? ctx->d_x : nullptr);
        if (rc == MG_OK && w_svr) rc = launch_svr(ctx, ctx->d_x, m, nullptr, d_svr);
        if (rc == MG_OK) {
            if (w_log) TRY_FREE(cudaMemcpyAsync(logistic + i0, d_log, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (w_svr) TRY_FREE(cudaMemcpyAsync(svr + i0, d_svr, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (w_feat) TRY_FREE(cudaMemcpyAsync(features + i0 * MG_NFEAT, ctx->d_x, (size_t)m * MG_NFEAT * 8, cudaMemcpyDeviceToHost, ctx->stream));
            TRY_FREE(cudaStreamSynchronize(ctx->stream));
        }
#undef TRY_FREE
        cleanup();
    }
    return rc;
}

// ---------------------------------------------------------------------------
// region grids
// ---------------------------------------------------------------------------
int mg_host_first_scan(const HostConfig &c, const mg_region *r)
{
    if (r->scan_begin > 0) return r->scan_begin;
    // mipgen.cpp:421-425 (the loop pre-increments)
    int cur = r->start_flanked - c.max_capture + c.max_sum;
    if (cur < 0) cur = 0;
    return cur + 1;
}

int mg_host_n_scan(const HostConfig &c, const mg_region *r)
{
    int last = (r->scan_begin > 0 && r->scan_end > 0) ? r->scan_end : r->stop_flanked;
    int n = last - mg_host_first_scan(c, r) + 1;
    return n < 0 ? 0 : n;
}

extern "C" int mg_first_scan_start(const mg_ctx *ctx, const mg_region *r) { return (ctx && ctx->has_cfg && r) ? mg_host_first_scan(ctx->cfg, r) : 0; }

extern "C" int64_t mg_grid_size(const mg_ctx *ctx, const mg_region *r)
{
    if (!ctx || !ctx->has_cfg || !r) return 0;
    return (int64_t)mg_host_n_scan(ctx->cfg, r) * ctx->cfg.n_cap * (int64_t)ctx->cfg.ext_len.size() * 2;
}

extern "C" int64_t mg_config_grid_size(const mg_config *cfg, const mg_region *r)
{
    HostConfig h;
    std::string err;
    if (!r || mg_host_config_from(cfg, h, err) != MG_OK) return -1;
    return (int64_t)mg_host_n_scan(h, r) * h.n_cap * (int64_t)h.ext_len.size() * 2;
}

extern "C" int mg_config_first_scan_start(const mg_config *cfg, const mg_region *r)
{
    HostConfig h;
    std::string err;
    if (!r || mg_host_config_from(cfg, h, err) != MG_OK) return -1;
    return mg_host_first_scan(h, r);
}

extern "C" void mg_panel_destroy(mg_panel *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    mg_ctx *c = p->ctx;
    mg_dev_free(c, p->d_ftasks); mg_dev_free(c, p->d_w); mg_dev_free(c, p->d_w_tc);
    mg_dev_free(c, p->d_regions); mg_dev_free(c, p->d_tasks); mg_dev_free(c, p->d_codes); mg_dev_free(c, p->d_lrc); mg_dev_free(c, p->d_copies);
    mg_dev_free(c, p->d_maskpf); mg_dev_free(c, p->d_snppf); mg_dev_free(c, p->d_unmap); mg_dev_free(c, p->d_ascii);
    mg_dev_free(c, p->d_valid); mg_dev_free(c, p->d_state); mg_dev_free(c, p->d_logistic); mg_dev_free(c, p->d_svr); mg_dev_free(c, p->d_feat);
    delete p;
}

extern "C" int mg_panel_create(mg_ctx *ctx, const mg_region *regions, int n, mg_panel **out)
{
    if (!ctx || !out || n < 0 || (!regions && n > 0)) return MG_ERR_INVALID;
    *out = nullptr;
    if (!ctx->has_cfg) { ctx->err = "mg_set_config has not been called"; return MG_ERR_NOCONFIG; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    mg_panel *p = new mg_panel();
    p->ctx = ctx;
    p->cfg_serial = ctx->cfg_serial;
    p->n_regions = n;
    p->offsets.assign(n + 1, 0);
    p->h_regions.resize(std::max(n, 1));
    int64_t codes = 0, copies = 0, unmap = 0;
    bool any_lrc = false, any_snp = false;
    const int n_oligo = (int)ctx->cfg.oligo_sizes.size();
    for (int i = 0; i < n; i++) {
        const mg_region &r = regions[i];
        if (!r.seq || r.seq_len <= 0 || r.seq_len != r.seq_stop - r.seq_start + 1 || r.stop_flanked < r.start_flanked) {
            ctx->err = "mg_panel_create: region " + std::to_string(i) + " has inconsistent sequence coordinates";
            delete p;
            return MG_ERR_INVALID;
        }
        if (r.copies && n_oligo == 0) {
            ctx->err = "mg_panel_create: copy table given but the config has no oligo_sizes";
            delete p;
            return MG_ERR_INVALID;
        }
        DevRegion &d = p->h_regions[i];
        d.seq_off = codes; d.grid_off = p->offsets[i];
        d.copy_off = r.copies ? copies : -1;
        d.seq_len = r.seq_len; d.seq_start = r.seq_start; d.seq_stop = r.seq_stop;
        d.start_flanked = r.start_flanked; d.stop_flanked = r.stop_flanked;
        d.first_scan = mg_host_first_scan(ctx->cfg, &r); d.n_scan = mg_host_n_scan(ctx->cfg, &r);
        d.has_snp = r.snp != nullptr;
        d.aux_off = codes + i;  // seq_len + 1 prefix entries per region
        d.unmap_off = r.unmappable ? unmap : -1;
        codes += r.seq_len;
        if (r.copies) copies += (int64_t)n_oligo * r.seq_len;
        if (r.unmappable) unmap += (int64_t)ctx->cfg.n_cap * r.seq_len;
        any_lrc |= r.lrc != nullptr;
        any_snp |= r.snp != nullptr;
        p->has_sel_inputs |= r.masked_seq || r.snp || r.unmappable;
        p->offsets[i + 1] = p->offsets[i] + mg_grid_size(ctx, &r);
    }
    p->n_cand = p->offsets[n];
    p->n_codes = codes;
    // K-feat work items: windows of W consecutive scan starts; W grows with the panel so that
    // there are several windows per SM but the per-window prefix tables stay well amortised
    {
        int64_t total_scan = 0;
        for (int i = 0; i < n; i++) total_scan += p->h_regions[i].n_scan;
        int W = (int)std::min<int64_t>(64, std::max<int64_t>(8, total_scan / ((int64_t)ctx->sm_count * 6)));
        W = W / 8 * 8;  // factored-SVR windows (8, 4, 2 or 1 scan starts) must nest inside K-feat windows
        const int64_t per_scan = (int64_t)ctx->cfg.n_cap * (int64_t)ctx->cfg.ext_len.size() * 2;
        while (W > 1 && W * per_scan > (1 << 24)) W /= 2;  // keep a window's candidate count in int range
        const int n_cap = ctx->cfg.n_cap, n_pairs2 = (int)ctx->cfg.ext_len.size() * 2;
        // captures per K-feat work item: keep a work item around 8-16 k candidates
        const int nci_max = std::max(1, std::min(n_cap, 16384 / std::max(1, W * n_pairs2)));
        for (int i = 0; i < n; i++)
            for (int si = 0; si < p->h_regions[i].n_scan; si += W) {
                DevTask t;
                t.region = i; t.si0 = si; t.nsi = std::min(W, p->h_regions[i].n_scan - si); t.ft0 = 0;
                t.ci0 = 0; t.nci = n_cap;
                t.g0 = p->h_regions[i].grid_off + (int64_t)si * per_scan;
                p->task_start.push_back((int)p->h_tasks.size());
                p->h_windows.push_back(t);
                for (int c0 = 0; c0 < n_cap; c0 += nci_max) {
                    DevTask u = t;
                    u.ci0 = c0; u.nci = std::min(nci_max, n_cap - c0);
                    p->h_tasks.push_back(u);
                }
            }
        p->task_start.push_back((int)p->h_tasks.size());
        // factored-SVR tasks, grouped by window so a chunk of windows is a contiguous task range
        p->ftask_start.assign(p->h_windows.size() + 1, 0);
        if (ctx->fact_ok && W % ctx->fact_W == 0) {
            for (size_t k = 0; k < p->h_windows.size(); k++) {
                const DevTask &t = p->h_windows[k];
                p->ftask_start[k] = (int)p->h_ftasks.size();
                for (int si = 0; si < t.nsi; si += ctx->fact_W)
                    for (int ci = 0; ci < ctx->cfg.n_cap; ci++)
                        for (int strand = 0; strand < 2; strand++) {
                            DevFTask f;
                            f.g0 = p->h_regions[t.region].grid_off;
                            f.region = t.region; f.si0 = t.si0 + si; f.nsi = std::min(ctx->fact_W, t.nsi - si);
                            f.ci = ci; f.strand = strand; f.pad = 0;
                            p->h_ftasks.push_back(f);
                            const DevFact &hf = ctx->h_fact;
                            const int64_t arm_rows = (int64_t)f.nsi * (strand ? hf.n_lig : hf.n_ext) +
                                                     (int64_t)(f.nsi + hf.max_sum - hf.min_sum) * (strand ? hf.n_ext : hf.n_lig);
                            const int64_t ins_rows = (int64_t)f.nsi * hf.n_sums;
                            p->row_table_bytes += arm_rows * FACT_LD_ARM * 8 + ins_rows * FACT_LD_INS * 8 + (arm_rows + ins_rows) * 12;
                        }
            }
            p->ftask_start[p->h_windows.size()] = (int)p->h_ftasks.size();
            for (size_t k = 0; k < p->h_windows.size(); k++)   // K-feat's row-table mode addresses the work items nested in a window
                for (int u = p->task_start[k]; u < p->task_start[k + 1]; u++) p->h_tasks[u].ft0 = p->ftask_start[k];
        }
        p->span_cap = W + ctx->cfg.max_arm + ctx->cfg.max_capture - ctx->cfg.min_arm + 4;
        int words = (p->span_cap + 2) / 2;
        if (words % 2 == 0) words++;  // odd word stride: table rows spread over the banks
        p->pf_stride = 2 * words;
    }
#define P_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); mg_panel_destroy(p); return MG_ERR_CUDA; } } while (0)
    if (n > 0) {
        std::vector<char> ascii((size_t)codes);
        std::vector<double> lrc(any_lrc ? (size_t)n * MG_NLRC : 0, 0.0);
        std::vector<int> cp((size_t)copies);
        // selection-only inputs: prefix counts of masked ('N') bases and of SNP positions per region, unmappable MIP starts
        std::vector<int> maskpf((size_t)(codes + n)), snppf(any_snp ? (size_t)(codes + n) : 0, 0);
        std::vector<uint8_t> um((size_t)unmap);
        for (int i = 0; i < n; i++) {
            const mg_region &r = regions[i];
            memcpy(&ascii[p->h_regions[i].seq_off], r.seq, r.seq_len);
            if (r.lrc) memcpy(&lrc[(size_t)i * MG_NLRC], r.lrc, MG_NLRC * 8);
            if (r.copies) memcpy(&cp[p->h_regions[i].copy_off], r.copies, (size_t)n_oligo * r.seq_len * sizeof(int));
            const char *ms = r.masked_seq ? r.masked_seq : r.seq;  // -trf off: the masked copy IS the sequence (mipgen.cpp:1058-1062)
            int *mp = &maskpf[(size_t)p->h_regions[i].aux_off];
            mp[0] = 0;
            for (int k = 0; k < r.seq_len; k++) mp[k + 1] = mp[k] + (ms[k] == 'N');
            if (r.snp) {
                int *sp = &snppf[(size_t)p->h_regions[i].aux_off];
                for (int k = 0; k < r.seq_len; k++) sp[k + 1] = sp[k] + (r.snp[k] != 0);
            }
            if (r.unmappable) memcpy(&um[(size_t)p->h_regions[i].unmap_off], r.unmappable, (size_t)ctx->cfg.n_cap * r.seq_len);
        }
        char *&d_ascii = p->d_ascii;  // kept: K-fmt-write prints the sequences
        P_TRY(mg_dev_alloc(ctx, (void **)&d_ascii, (size_t)codes));
        P_TRY(mg_dev_alloc(ctx, (void **)&p->d_codes, (size_t)codes));
        P_TRY(mg_dev_alloc(ctx, (void **)&p->d_regions, (size_t)n * sizeof(DevRegion)));
        P_TRY(cudaMemcpyAsync(d_ascii, ascii.data(), (size_t)codes, cudaMemcpyHostToDevice, ctx->stream));
        P_TRY(cudaMemcpyAsync(p->d_regions, p->h_regions.data(), (size_t)n * sizeof(DevRegion), cudaMemcpyHostToDevice, ctx->stream));
        if (!p->h_ftasks.empty()) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_ftasks, p->h_ftasks.size() * sizeof(DevFTask)));
            P_TRY(cudaMemcpyAsync(p->d_ftasks, p->h_ftasks.data(), p->h_ftasks.size() * sizeof(DevFTask), cudaMemcpyHostToDevice, ctx->stream));
        }
        if (!p->h_tasks.empty()) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_tasks, p->h_tasks.size() * sizeof(DevTask)));
            P_TRY(cudaMemcpyAsync(p->d_tasks, p->h_tasks.data(), p->h_tasks.size() * sizeof(DevTask), cudaMemcpyHostToDevice, ctx->stream));
        }
        if (any_lrc) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_lrc, lrc.size() * 8));
            P_TRY(cudaMemcpyAsync(p->d_lrc, lrc.data(), lrc.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (copies > 0) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_copies, (size_t)copies * sizeof(int)));
            P_TRY(cudaMemcpyAsync(p->d_copies, cp.data(), (size_t)copies * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        }
        P_TRY(mg_dev_alloc(ctx, (void **)&p->d_maskpf, maskpf.size() * sizeof(int)));
        P_TRY(cudaMemcpyAsync(p->d_maskpf, maskpf.data(), maskpf.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (any_snp) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_snppf, snppf.size() * sizeof(int)));
            P_TRY(cudaMemcpyAsync(p->d_snppf, snppf.data(), snppf.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        }
        if (unmap > 0) {
            P_TRY(mg_dev_alloc(ctx, (void **)&p->d_unmap, (size_t)unmap));
            P_TRY(cudaMemcpyAsync(p->d_unmap, um.data(), (size_t)unmap, cudaMemcpyHostToDevice, ctx->stream));
        }
        int rc = launch_encode(ctx, d_ascii, p->d_codes, codes);
        cudaStreamSynchronize(ctx->stream);  // host staging vectors go out of scope
        if (rc != MG_OK) { mg_panel_destroy(p); return rc; }
    }
    if (p->n_cand > 0) P_TRY(mg_dev_alloc(ctx, (void **)&p->d_valid, (size_t)p->n_cand));
#undef P_TRY
    *out = p;
    return MG_OK;
}

extern "C" int64_t mg_panel_candidates(const mg_panel *p) { return p ? p->n_cand : 0; }
extern "C" int64_t mg_panel_row_table_bytes(const mg_panel *p) { return p ? p->row_table_bytes : 0; }

// a panel caches grid offsets, windows and task lists derived from the config it was created under
static const char *const kStalePanel = "the panel was created under an earlier mg_set_config: create it again";

extern "C" int mg_panel_score(mg_ctx *ctx, mg_panel *p, int want)
{
    if (!ctx || !p || p->ctx != ctx) return MG_ERR_INVALID;
    if (p->cfg_serial != ctx->cfg_serial) { ctx->err = kStalePanel; return MG_ERR_INVALID; }
    if ((want & MG_WANT_SVR) && !ctx->has_model) { ctx->err = "no SVR model loaded"; return MG_ERR_NOMODEL; }
    if (p->n_cand == 0) return MG_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool w_log = want & MG_WANT_LOGISTIC, w_svr = want & MG_WANT_SVR, w_feat = want & MG_WANT_FEATURES;
    if (w_log && !p->d_logistic) CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_logistic, (size_t)p->n_cand * 8));
    if (w_svr && !p->d_svr) CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_svr, (size_t)p->n_cand * 8));
    if (w_feat && !p->d_feat) {
        int64_t rows = (p->n_cand + SVR_BM - 1) / SVR_BM * SVR_BM;
        CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_feat, (size_t)rows * MG_NFEAT * 8));
        CUDA_TRY(ctx, cudaMemsetAsync(p->d_feat, 0, (size_t)rows * MG_NFEAT * 8, ctx->stream));
    }
    int rc = MG_OK;
    const int n_win = (int)p->h_windows.size(), n_tasks = (int)p->h_tasks.size();
    // tensor-core form on request (mode 3)
    const bool tc = w_svr && ctx->svr_mode == 3;
    if (tc) {
        if (!ctx->tc_ok || ctx->cfg.max_capture > 1024) {
            ctx->err = "tensor-core SVR requested but the model's length / junction columns (or the capture sizes) are not small integers";
            return MG_ERR_INVALID;
        }
        if (!p->d_w_tc || p->w_tc_n_sv_pad != ctx->n_sv_pad) {
            mg_dev_free(ctx, p->d_w_tc);
            p->d_w_tc = nullptr;
            CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_w_tc, (size_t)p->n_regions * ctx->n_sv_pad * 8));
            p->w_tc_n_sv_pad = ctx->n_sv_pad;
        }
        if ((rc = launch_lrc_weights_tc(ctx, p, p->d_w_tc)) != MG_OK) return rc;
    }
    // factored SVR when the configuration fits its tables (mode 0/2); dense otherwise (mode 1, or as fallback)
    const bool fact = w_svr && !tc && ctx->svr_mode != 1 && ctx->fact_ok && !p->h_ftasks.empty();
    if (w_svr && ctx->svr_mode == 2 && !fact) {
        ctx->err = "factored SVR requested but the configuration does not fit its shared-memory tables";
        return MG_ERR_INVALID;
    }
    if (fact) {
        if (!p->d_w || p->w_n_sv_pad != ctx->n_sv_pad) {
            mg_dev_free(ctx, p->d_w);
            p->d_w = nullptr;
            CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_w, (size_t)p->n_regions * ctx->n_sv_pad * 8));
            p->w_n_sv_pad = ctx->n_sv_pad;
        }
        if ((rc = launch_lrc_weights(ctx, p, p->d_w)) != MG_OK) return rc;
    }
    // windows [w0, w1) <-> candidates [g0, g1)
    auto svr_range = [&](int w0, int w1, const double *xbuf, int64_t g0, int64_t g1) {
        if (tc) return launch_svr_tc(ctx, p, xbuf, g0, g1, p->d_valid, p->d_w_tc, p->d_svr);
        return launch_svr(ctx, xbuf, g1 - g0, p->d_valid + g0, p->d_svr + g0);
    };
    auto end_of = [&](int w) { return w < n_win ? p->h_windows[w].g0 : p->n_cand; };
    if (!w_svr && !w_feat) {
        rc = launch_feat_grid(ctx, p, 0, n_tasks, 0, p->n_cand, p->d_valid, w_log ? p->d_logistic : nullptr, nullptr);
    } else if (fact) {
        // Factored SVR: K-feat hands over the distinct arm / insert rows of every work item (row-table mode) and the grid
        // points' states; no 192-vector is materialised unless the caller asked for the features themselves.
        if (w_feat) rc = launch_feat_grid(ctx, p, 0, n_tasks, 0, p->n_cand, p->d_valid, w_log ? p->d_logistic : nullptr, p->d_feat);
        if (!p->d_state) CUDA_TRY(ctx, mg_dev_alloc(ctx, (void **)&p->d_state, (size_t)p->n_cand));
        const size_t max_items = std::max<size_t>(1, kMaxRowsBytes / ((size_t)ctx->h_fact.blob_doubles * 8));
        int w0 = 0;
        while (w0 < n_win && rc == MG_OK) {
            int w1 = w0 + 1;
            while (w1 < n_win && (size_t)(p->ftask_start[w1 + 1] - p->ftask_start[w0]) <= max_items) w1++;
            const int64_t g0 = p->h_windows[w0].g0, g1 = end_of(w1);
            const int f0 = p->ftask_start[w0], f1 = p->ftask_start[w1];
            if ((rc = ensure_rows(ctx, (size_t)(f1 - f0))) != MG_OK) return rc;
            rc = launch_feat_grid(ctx, p, p->task_start[w0], p->task_start[w1], g0, g1 - g0, p->d_valid,
                                  (w_log && !w_feat) ? p->d_logistic : nullptr, nullptr, p->d_state, ctx->d_rows, f0);
            if (rc == MG_OK) rc = launch_svr_fact(ctx, p, f0, f1, ctx->d_rows, g1 - g0, p->d_state, p->d_w, p->d_svr);
            w0 = w1;
        }
    } else if (w_feat) {
        rc = launch_feat_grid(ctx, p, 0, n_tasks, 0, p->n_cand, p->d_valid, w_log ? p->d_logistic : nullptr, p->d_feat);
        if (rc == MG_OK && w_svr) rc = svr_range(0, n_win, p->d_feat, 0, p->n_cand);
    } else {
        // feature rows live only in a workspace: walk the panel in chunks of whole windows
        const int64_t n_chunks = (p->n_cand + kMaxChunkRows - 1) / kMaxChunkRows;
        const int64_t target = std::min<int64_t>(kMaxChunkRows, p->n_cand / n_chunks + (1 << 16));  // even chunks, no tiny tail
        int w0 = 0;
        while (w0 < n_win && rc == MG_OK) {
            const int64_t g0 = p->h_windows[w0].g0;
            int w1 = w0 + 1;
            while (w1 < n_win && end_of(w1 + 1) - g0 <= target) w1++;
            const int64_t g1 = end_of(w1);
            if ((rc = ensure_x(ctx, g1 - g0)) != MG_OK) return rc;
            rc = launch_feat_grid(ctx, p, p->task_start[w0], p->task_start[w1], g0, g1 - g0, p->d_valid, w_log ? p->d_logistic : nullptr, ctx->d_x);
            if (rc == MG_OK) rc = svr_range(w0, w1, ctx->d_x, g0, g1);
            w0 = w1;
        }
    }
    if (rc == MG_OK) {
        p->has_valid = true;
        p->has_logistic |= w_log;
        p->has_svr |= w_svr;
        p->has_feat |= w_feat;
    }
    return rc;
}

extern "C" int mg_panel_fetch(mg_ctx *ctx, mg_panel *p, uint8_t *valid, double *logistic, double *svr, double *features)
{
    if (!ctx || !p || p->ctx != ctx) return MG_ERR_INVALID;
    if (p->cfg_serial != ctx->cfg_serial) { ctx->err = kStalePanel; return MG_ERR_INVALID; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (p->n_cand > 0) {
        if ((valid && !p->has_valid) || (logistic && !p->has_logistic) || (svr && !p->has_svr) || (features && !p->has_feat)) {
            ctx->err = "mg_panel_fetch: requested output was never computed";
            return MG_ERR_INVALID;
        }
        if (valid) CUDA_TRY(ctx, cudaMemcpyAsync(valid, p->d_valid, (size_t)p->n_cand, cudaMemcpyDeviceToHost, ctx->stream));
        if (logistic) CUDA_TRY(ctx, cudaMemcpyAsync(logistic, p->d_logistic, (size_t)p->n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (svr) CUDA_TRY(ctx, cudaMemcpyAsync(svr, p->d_svr, (size_t)p->n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (features) CUDA_TRY(ctx, cudaMemcpyAsync(features, p->d_feat, (size_t)p->n_cand * MG_NFEAT * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MG_OK;
}

extern "C" int mg_panel_device_ptrs(const mg_panel *p, const uint8_t **valid, const double **logistic, const double **svr)
{
    if (!p) return MG_ERR_INVALID;
    if (valid) *valid = p->has_valid ? p->d_valid : nullptr;
    if (logistic) *logistic = p->has_logistic ? p->d_logistic : nullptr;
    if (svr) *svr = p->has_svr ? p->d_svr : nullptr;
    return MG_OK;
}

extern "C" int64_t mg_panel_valid_candidates(const mg_panel *p)
{
    if (!p || !p->has_valid || p->n_cand == 0) return 0;
    mg_ctx *ctx = p->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    unsigned long long *d_count = ctx->d_work + 3, h = 0;  // the context's scratch counter word
    if (cudaMemsetAsync(d_count, 0, sizeof h, ctx->stream) != cudaSuccess) return -1;
    if (launch_count_valid(ctx, p->d_valid, p->n_cand, d_count) != MG_OK) return -1;
    if (cudaMemcpyAsync(&h, d_count, sizeof h, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return (int64_t)h;
}

extern "C" int mg_score_regions(mg_ctx *ctx, const mg_region *regions, int n, int want, int64_t *out_offsets, uint8_t *valid,
                                double *logistic, double *svr, double *features)
{
    mg_panel *p = nullptr;
    int rc = mg_panel_create(ctx, regions, n, &p);
    if (rc != MG_OK) return rc;
    if (out_offsets) memcpy(out_offsets, p->offsets.data(), (size_t)(n + 1) * sizeof(int64_t));
    int w = 0;
    if ((want & MG_WANT_LOGISTIC) && logistic) w |= MG_WANT_LOGISTIC;
    if ((want & MG_WANT_SVR) && svr) w |= MG_WANT_SVR;
    if ((want & MG_WANT_FEATURES) && features) w |= MG_WANT_FEATURES;
    rc = mg_panel_score(ctx, p, w);
    if (rc == MG_OK)
        rc = mg_panel_fetch(ctx, p, valid, (w & MG_WANT_LOGISTIC) ? logistic : nullptr, (w & MG_WANT_SVR) ? svr : nullptr,
                            (w & MG_WANT_FEATURES) ? features : nullptr);
    mg_panel_destroy(p);
    return rc;
}

// ---------------------------------------------------------------------------
// selection front-end
// ---------------------------------------------------------------------------
extern "C" int mg_region_scan_count(const mg_ctx *ctx, const mg_region *r) { return (ctx && ctx->has_cfg && r) ? mg_host_n_scan(ctx->cfg, r) : 0; }

int mg_host_n_positions(const HostConfig &c, const mg_region *r)
{
    int n = r->stop_flanked + c.max_capture - c.min_sum - 1 - mg_host_first_scan(c, r) + 1;
    return n < 0 ? 0 : n;
}

extern "C" int mg_region_position_count(const mg_ctx *ctx, const mg_region *r) { return (ctx && ctx->has_cfg && r) ? mg_host_n_positions(ctx->cfg, r) : 0; }

extern "C" int mg_panel_select(mg_ctx *ctx, mg_panel *p, const mg_select_params *sp, int64_t *scan_best, int64_t *pos_best)
{
    if (!ctx || !p || p->ctx != ctx || !sp || !scan_best || !pos_best) return MG_ERR_INVALID;
    if (p->cfg_serial != ctx->cfg_serial) { ctx->err = kStalePanel; return MG_ERR_INVALID; }
    const double *d_score = sp->method == 1 ? (p->has_svr ? p->d_svr : nullptr) : (p->has_logistic ? p->d_logistic : nullptr);
    if (!d_score) { ctx->err = "mg_panel_select: the panel has not been scored with the scores this method selects on"; return MG_ERR_INVALID; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int n = p->n_regions;
    std::vector<int64_t> so(n + 1, 0), po(n + 1, 0);
    for (int i = 0; i < n; i++) {
        so[i + 1] = so[i] + p->h_regions[i].n_scan;
        int np = p->h_regions[i].stop_flanked + ctx->cfg.max_capture - ctx->cfg.min_sum - 1 - p->h_regions[i].first_scan + 1;
        po[i + 1] = po[i] + (np < 0 ? 0 : np);
    }
    int64_t *d_so = nullptr, *d_po = nullptr, *d_sb = nullptr, *d_pb = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_so); mg_dev_free(ctx, d_po); mg_dev_free(ctx, d_sb); mg_dev_free(ctx, d_pb); };
#define S_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return MG_ERR_CUDA; } } while (0)
    S_TRY(mg_dev_alloc(ctx, (void **)&d_so, (size_t)(n + 1) * 8));
    S_TRY(mg_dev_alloc(ctx, (void **)&d_po, (size_t)(n + 1) * 8));
    S_TRY(mg_dev_alloc(ctx, (void **)&d_sb, (size_t)std::max<int64_t>(so[n], 1) * 16));
    S_TRY(mg_dev_alloc(ctx, (void **)&d_pb, (size_t)std::max<int64_t>(po[n], 1) * 16));
    S_TRY(cudaMemcpyAsync(d_so, so.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    S_TRY(cudaMemcpyAsync(d_po, po.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_select(ctx, p, d_so, d_po, so[n], po[n], d_score, sp, d_sb, d_pb);
    if (rc == MG_OK) {
        S_TRY(cudaMemcpyAsync(scan_best, d_sb, (size_t)so[n] * 16, cudaMemcpyDeviceToHost, ctx->stream));
        S_TRY(cudaMemcpyAsync(pos_best, d_pb, (size_t)po[n] * 16, cudaMemcpyDeviceToHost, ctx->stream));
        S_TRY(cudaStreamSynchronize(ctx->stream));  // so/po staging vectors go out of scope
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
#undef S_TRY
    cleanup();
    return rc;
}

extern "C" int mg_panel_gather(mg_ctx *ctx, mg_panel *p, const int64_t *idx, int64_t n, double *logistic, double *svr)
{
    if (!ctx || !p || p->ctx != ctx || n < 0 || (n > 0 && !idx)) return MG_ERR_INVALID;
    if ((logistic && !p->has_logistic) || (svr && !p->has_svr)) { ctx->err = "mg_panel_gather: requested score was never computed"; return MG_ERR_INVALID; }
    if (n == 0 || (!logistic && !svr)) return MG_OK;
    for (int64_t i = 0; i < n; i++)
        if (idx[i] >= p->n_cand) { ctx->err = "mg_panel_gather: grid index outside the panel"; return MG_ERR_INVALID; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t *d_idx = nullptr;
    double *d_a = nullptr, *d_b = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_idx); mg_dev_free(ctx, d_a); mg_dev_free(ctx, d_b); };
#define G_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cudaStreamSynchronize(ctx->stream); cleanup(); return MG_ERR_CUDA; } } while (0)
    G_TRY(mg_dev_alloc(ctx, (void **)&d_idx, (size_t)n * 8));
    if (logistic) G_TRY(mg_dev_alloc(ctx, (void **)&d_a, (size_t)n * 8));
    if (svr) G_TRY(mg_dev_alloc(ctx, (void **)&d_b, (size_t)n * 8));
    G_TRY(cudaMemcpyAsync(d_idx, idx, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_gather(ctx, d_idx, n, p->d_logistic, d_a, p->d_svr, d_b);
    if (rc == MG_OK) {
        if (logistic) G_TRY(cudaMemcpyAsync(logistic, d_a, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (svr) G_TRY(cudaMemcpyAsync(svr, d_b, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    G_TRY(cudaStreamSynchronize(ctx->stream));
#undef G_TRY
    cleanup();
    return rc;
}

// ---------------------------------------------------------------------------
// host helper: the score-dependent control flow of the tile loop
// ---------------------------------------------------------------------------
extern "C" int64_t mg_tile_replay(const mg_config *cfg, const mg_region *r, const uint8_t *valid, const double *score, int method,
                                  int heuristic, double upper, int64_t *out_idx, int64_t cap)
{
    HostConfig c;
    std::string err;
    if (!r || !valid || !score || mg_host_config_from(cfg, c, err) != MG_OK) return -1;
    const int n_pairs = (int)c.ext_len.size(), ns = mg_host_n_scan(c, r);
    // arm-sum groups: maximal runs of pairs with equal ext+lig (mipgen.cpp:431, 438)
    std::vector<std::pair<int, int>> groups;
    for (int p = 0; p < n_pairs;) {
        int q = p, sum = c.ext_len[p] + c.lig_len[p];
        while (q < n_pairs && c.ext_len[q] + c.lig_len[q] == sum) q++;
        groups.emplace_back(p, q);
        p = q;
    }
    int64_t n = 0;
    for (int si = 0; si < ns; si++) {
        double best = 0;  // previous_best_score, reset per scan start (:426)
        for (int ci = 0; ci < c.n_cap; ci++) {
            const int capture = c.max_capture - ci * c.inc;
            if (capture > r->stop_flanked - r->start_flanked + c.max_mip_overlap && capture - c.inc >= c.min_capture) continue;  // :429
            if (best > upper) continue;                                                                                        // :430
            for (auto &g : groups) {
                const int sum = c.ext_len[g.first] + c.lig_len[g.first];
                if (best > upper && sum != c.min_sum) continue;  // :434
                int prev_minus = 0, prev_plus = 0;               // ints in the reference (:435-436)
                for (int k = g.first; k < g.second; k++) {
                    const int64_t idx = (((int64_t)si * c.n_cap + ci) * n_pairs + k) * 2;
                    if (!valid[idx]) continue;  // :443-444
                    const double plus = score[idx], minus = score[idx + 1];
                    if (out_idx && n + 2 <= cap) { out_idx[n] = idx; out_idx[n + 1] = idx + 1; }
                    n += 2;
                    best = minus > plus ? minus : plus;  // :495
                    const bool stop = method == 0 && heuristic && plus < prev_plus && minus < prev_minus;  // :494
                    prev_minus = (int)minus;  // :496
                    prev_plus = (int)plus;    // :497
                    if (stop) break;          // skip_ahead: the rest of this arm-sum list is skipped (:440)
                }
            }
        }
    }
    return n;
}

// ---------------------------------------------------------------------------
// host helpers: grid index -> the fields of an SVMipv4 object -> a design-file record
// ---------------------------------------------------------------------------
extern "C" int mg_describe_candidates(const mg_config *cfg, const mg_region *r, const int64_t *idx, int n, mg_mip_info *out)
{
    HostConfig c;
    std::string err;
    if (!r || (n > 0 && (!idx || !out)) || mg_host_config_from(cfg, c, err) != MG_OK) return MG_ERR_INVALID;
    const int n_pairs = (int)c.ext_len.size(), ns = mg_host_n_scan(c, r), s0 = mg_host_first_scan(c, r);
    const int64_t grid = (int64_t)ns * c.n_cap * n_pairs * 2;
    auto copy_of = [&](int start, int len) {  // mipgen.cpp:612-613; an absent key reads as 0
        if (!r->copies) return 1;
        for (size_t k = 0; k < c.oligo_sizes.size(); k++)
            if (c.oligo_sizes[k] == len) {
                const int i = start - r->seq_start;
                return (i < 0 || i >= r->seq_len) ? 0 : r->copies[(int64_t)k * r->seq_len + i];
            }
        return 0;
    };
    for (int k = 0; k < n; k++) {
        int64_t q = idx[k];
        if (q < 0 || q >= grid) return MG_ERR_INVALID;
        mg_mip_info &m = out[k];
        m.strand = (int)(q & 1); q >>= 1;
        const int p = (int)(q % n_pairs); q /= n_pairs;
        const int ci = (int)(q % c.n_cap), si = (int)(q / c.n_cap);
        m.ext_len = c.ext_len[p]; m.lig_len = c.lig_len[p];
        m.scan_start = s0 + si;
        m.scan_stop = m.scan_start + (c.max_capture - ci * c.inc) - m.ext_len - m.lig_len - 1;
        if (m.strand == 0) {  // PlusSVMipv4.cpp:7-14
            m.ext_start = m.scan_start - m.ext_len; m.ext_stop = m.scan_start - 1;
            m.lig_start = m.scan_stop + 1; m.lig_stop = m.scan_stop + m.lig_len;
        } else {              // MinusSVMipv4.cpp:30-37
            m.lig_start = m.scan_start - m.lig_len; m.lig_stop = m.scan_start - 1;
            m.ext_start = m.scan_stop + 1; m.ext_stop = m.scan_stop + m.ext_len;
        }
        m.ext_copy = copy_of(m.ext_start, m.ext_len);
        m.lig_copy = copy_of(m.lig_start, m.lig_len);
    }
    return MG_OK;
}

extern "C" int64_t mg_format_mip_record(const mg_region *r, const mg_mip_info *m, double score, const char *chr, const char *label,
                                        int feature_start, int feature_stop, const char *universal_middle, int mip_index,
                                        char *buf, int64_t cap)
{
    if (!r || !r->seq || !m || !chr || !label || !universal_middle || !buf || cap <= 0) return -1;
    // the failure flags are printed as "000": a region that declares TRF / SNP / mappability inputs may need other flags
    // (and _SNP_ name suffixes, mipgen.cpp:790-792) that only design_mip can derive -- refuse rather than print a wrong record
    if (r->masked_seq || r->snp || r->unmappable) return -1;
    // the three strings as the object holds them: genomic on '+', reverse-complemented on '-'
    auto cut = [&](int start, int stop, std::string &out) {
        const int a = start - r->seq_start, n = stop - start + 1;
        if (a < 0 || n < 0 || a + n > r->seq_len) return false;
        out.assign(r->seq + a, (size_t)n);
        if (m->strand) {
            std::reverse(out.begin(), out.end());
            for (char &ch : out) ch = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : ch;  // MinusSVMipv4.cpp:6-29
        }
        return true;
    };
    std::string ext, lig, tgt;
    if (!cut(m->ext_start, m->ext_stop, ext) || !cut(m->lig_start, m->lig_stop, lig) || !cut(m->scan_start, m->scan_stop, tgt)) return -1;
    const char strand = m->strand ? '-' : '+';
    char head[160], mid[192], tail[160];  // numbers only: the chromosome name, of any length, goes in as a string
    std::string rec = chr;
    snprintf(head, sizeof head, ":%d-%d/%d,%d/%c\t%g\t", m->strand ? m->lig_start : m->ext_start, m->strand ? m->ext_stop : m->lig_stop,
             m->ext_len, m->lig_len, strand, score);
    rec += head; rec += chr;
    snprintf(head, sizeof head, "\t%d\t%d\t%d\t", m->ext_start, m->ext_stop, m->ext_copy);
    rec += head;
    snprintf(mid, sizeof mid, "\t%d\t%d\t%d\t", m->lig_start, m->lig_stop, m->lig_copy);
    rec += ext; rec += mid; rec += lig;
    snprintf(mid, sizeof mid, "\t%d\t%d\t", m->scan_start, m->scan_stop);
    rec += mid; rec += tgt; rec += '\t';
    rec += lig; rec += universal_middle; rec += ext;
    snprintf(tail, sizeof tail, "\t%d\t%d\t%c\t000\t", feature_start - 1, feature_stop, strand);
    rec += tail; rec += label;
    snprintf(tail, sizeof tail, "_%04d\n", mip_index);
    rec += tail;
    if ((int64_t)rec.size() + 1 > cap) return -1;
    memcpy(buf, rec.c_str(), rec.size() + 1);
    return (int64_t)rec.size();
}

extern "C" int64_t mg_format_mip_records(const mg_config *cfg, const mg_region *r, const int64_t *idx, int n, const double *score,
                                         const char *chr, const char *label, int feature_start, int feature_stop,
                                         const char *universal_middle, int first_index, char *buf, int64_t cap)
{
    if (n < 0 || (n > 0 && (!idx || !score || !buf))) return -1;
    std::vector<mg_mip_info> info((size_t)n);
    if (mg_describe_candidates(cfg, r, idx, n, info.data()) != MG_OK) return -1;
    int64_t at = 0;
    for (int k = 0; k < n; k++) {
        const int64_t len = mg_format_mip_record(r, &info[k], score[idx[k]], chr, label, feature_start, feature_stop, universal_middle,
                                                 first_index + k, buf + at, cap - at);
        if (len < 0) return -1;
        at += len;  // the next record overwrites this one's NUL
    }
    return at;
}
