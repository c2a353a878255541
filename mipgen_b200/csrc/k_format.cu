// k_format.cu -- design-file records written on the device.
//
// print_details (mipgen.cpp:765-794) turns one SVMipv4 object into one tab-separated line of all_mips.txt /
// collapsed_mips.txt.  With the candidates, their scores, the region sequences and the copy tables already resident in HBM
// the lines are pure formatting:
//
//   K-fmt-len    one thread per record: geometry from the grid index, the score printed the way `ostream << double` does
//                (== printf "%g": 6 significant digits, correctly rounded from the exact binary value -- 128-bit integer
//                arithmetic, no floating-point rounding anywhere), the record's length;
//   (host)       exclusive prefix sum of the lengths -> byte offsets;
//   K-fmt-write  one warp per record: the record is assembled in shared memory (lanes copy the arm / insert sequences,
//                reverse-complementing on the minus strand as MinusSVMipv4.cpp:6-29 does) and written out with coalesced
//                stores.
// Only records whose failure flags are "000" (no TRF / SNP / mappability inputs on the region) are produced here: the other
// flags and the _SNP_ name suffixes need design_mip's allele logic (mipgen.cpp:634-760) and stay with the caller.
#include <string.h>

#include <algorithm>

#include "mg_common.cuh"

namespace {

struct FmtRegion {       // per-region strings of the records
    int chr_off, chr_len, label_off, label_len;
    int feature_start, feature_stop;
};

struct FmtRec {          // K-fmt-len -> K-fmt-write
    int region, strand, e, l, scan_start, scan_stop, ext_start, lig_start, ext_copy, lig_copy;
    int glen;            // length of the printed score
    char g[28];          // the printed score
};

__device__ __forceinline__ int n_digits(int v)  // characters of "%d"
{
    int n = v < 0 ? 1 : 0;
    unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    do { n++; u /= 10; } while (u);
    return n;
}

__device__ __forceinline__ int put_int(char *out, int v)
{
    char tmp[12];
    int n = 0, k = 0;
    unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) out[n++] = '-';
    while (k) out[n++] = tmp[--k];
    return n;
}

// printf("%g", v) -- what `ss << mip->score` prints (mipgen.cpp:773): 6 significant digits rounded half-to-even from the EXACT
// value m * 2^e2, trailing zeros removed, fixed notation for decimal exponents -4..5, scientific otherwise.  Returns the
// length, or -1 for magnitudes outside [1e-12, 1e15] (never a MIP score; the caller formats those on the host).
__device__ int format_g(double v, char *out)
{
    int n = 0;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (bits >> 63) out[n++] = '-';
    const int be = (int)((bits >> 52) & 0x7ff);
    unsigned long long m = bits & 0xfffffffffffffULL;
    if (be == 0x7ff) {
        const char *s = m ? "nan" : "inf";
        for (int i = 0; i < 3; i++) out[n++] = s[i];
        return n;
    }
    if (be == 0 && m == 0) { out[n++] = '0'; return n; }
    int e2;
    if (be == 0) e2 = -1074; else { m |= 1ULL << 52; e2 = be - 1075; }
    const double a = fabs(v);
    if (!(a >= 1e-12 && a < 1e15)) return -1;
    const unsigned long long p10[18] = {1ULL, 10ULL, 100ULL, 1000ULL, 10000ULL, 100000ULL, 1000000ULL, 10000000ULL, 100000000ULL, 1000000000ULL,
                                        10000000000ULL, 100000000000ULL, 1000000000000ULL, 10000000000000ULL, 100000000000000ULL,
                                        1000000000000000ULL, 10000000000000000ULL, 100000000000000000ULL};
    int X = (int)floor(log10(a));  // decimal exponent, corrected below if the estimate or the rounding moves it
    unsigned long long q = 0;
    for (int tries = 0; tries < 4; tries++) {
        const int k = 5 - X;  // q = round(a * 10^k)
        const int sh = -e2;   // a < 1e15 < 2^50  =>  e2 < 0
        if (k >= 0) {
            const unsigned __int128 N = (unsigned __int128)m * p10[k];  // < 2^53 * 10^17 < 2^110
            q = (unsigned long long)(N >> sh);
            const unsigned __int128 rem = N & ((((unsigned __int128)1) << sh) - 1), half = ((unsigned __int128)1) << (sh - 1);
            if (rem > half || (rem == half && (q & 1))) q++;
        } else {
            const unsigned __int128 den = (unsigned __int128)p10[-k] << sh;  // < 10^10 * 2^52
            q = (unsigned long long)((unsigned __int128)m / den);
            const unsigned __int128 rem2 = ((unsigned __int128)m % den) * 2;
            if (rem2 > den || (rem2 == den && (q & 1))) q++;
        }
        if (q >= 1000000ULL) { X++; continue; }
        if (q < 100000ULL) { X--; continue; }
        break;
    }
    int nd = 6;
    while (nd > 1 && q % 10 == 0) { q /= 10; nd--; }
    char dg[8];
    for (int i = nd - 1; i >= 0; i--) { dg[i] = (char)('0' + q % 10); q /= 10; }
    if (X < -4 || X >= 6) {  // d.ddddde+XX
        out[n++] = dg[0];
        if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; i++) out[n++] = dg[i]; }
        out[n++] = 'e';
        out[n++] = X < 0 ? '-' : '+';
        const int ax = X < 0 ? -X : X;
        if (ax < 10) out[n++] = '0';
        n += put_int(out + n, ax);
    } else if (X >= 0) {
        for (int i = 0; i <= X; i++) out[n++] = i < nd ? dg[i] : '0';
        if (nd > X + 1) { out[n++] = '.'; for (int i = X + 1; i < nd; i++) out[n++] = dg[i]; }
    } else {
        out[n++] = '0'; out[n++] = '.';
        for (int i = 0; i < -X - 1; i++) out[n++] = '0';
        for (int i = 0; i < nd; i++) out[n++] = dg[i];
    }
    return n;
}

__device__ __forceinline__ int copy_at(const DevConfig *__restrict__ cfg, const DevRegion &r, const int *__restrict__ copies, int start, int len)
{
    if (r.copy_off < 0) return 1;
    for (int k = 0; k < cfg->n_oligo; k++)
        if (cfg->oligo_sizes[k] == len) {
            const int i = start - r.seq_start;
            return (i < 0 || i >= r.seq_len) ? 0 : copies[r.copy_off + (int64_t)k * r.seq_len + i];
        }
    return 0;
}

__global__ void __launch_bounds__(128)
k_fmt_len(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, int n_regions, const FmtRegion *__restrict__ fr,
          const int *__restrict__ copies, const int64_t *__restrict__ idx, int64_t n, const double *__restrict__ score, int mid_len, int first_index,
          FmtRec *__restrict__ rec, int *__restrict__ len, int *__restrict__ status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t g = idx[i];
    if (g < 0) { atomicExch(status, 1); len[i] = 0; return; }
    int lo = 0, hi = n_regions - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (regions[mid].grid_off <= g) lo = mid; else hi = mid - 1;
    }
    const DevRegion r = regions[lo];
    const int n_cap = cfg->n_cap, n_pairs = cfg->n_pairs;
    int64_t q = g - r.grid_off;
    FmtRec o;
    o.region = lo;
    o.strand = (int)(q & 1); q >>= 1;
    const int p = (int)(q % n_pairs); q /= n_pairs;
    const int ci = (int)(q % n_cap), si = (int)(q / n_cap);
    if (si >= r.n_scan) { atomicExch(status, 1); len[i] = 0; return; }   // index outside the region's grid
    o.e = cfg->ext_len[p]; o.l = cfg->lig_len[p];
    o.scan_start = r.first_scan + si;
    o.scan_stop = o.scan_start + (cfg->max_capture - ci * cfg->inc) - o.e - o.l - 1;
    o.ext_start = o.strand ? o.scan_stop + 1 : o.scan_start - o.e;   // Plus/MinusSVMipv4 constructors
    o.lig_start = o.strand ? o.scan_start - o.l : o.scan_stop + 1;
    const int first = min(o.ext_start, o.lig_start) - r.seq_start, last = max(o.ext_start + o.e, o.lig_start + o.l) - r.seq_start;
    if (first < 0 || last > r.seq_len) { atomicExch(status, 1); len[i] = 0; return; }   // a window outside the sequence
    o.ext_copy = copy_at(cfg, r, copies, o.ext_start, o.e);
    o.lig_copy = copy_at(cfg, r, copies, o.lig_start, o.l);
    o.glen = format_g(score[g], o.g);
    if (o.glen < 0) { atomicExch(status, 2); o.glen = 0; }   // a score outside the device formatter's range
    const FmtRegion f = fr[lo];
    const int key_lo = o.strand ? o.lig_start : o.ext_start, key_hi = o.strand ? o.ext_start + o.e - 1 : o.lig_start + o.l - 1;
    const int scan_size = o.scan_stop - o.scan_start + 1;
    int L = f.chr_len + 1 + n_digits(key_lo) + 1 + n_digits(key_hi) + 1 + n_digits(o.e) + 1 + n_digits(o.l) + 1 + 1 + 1;   // key
    L += o.glen + 1 + f.chr_len + 1;
    L += n_digits(o.ext_start) + 1 + n_digits(o.ext_start + o.e - 1) + 1 + n_digits(o.ext_copy) + 1 + o.e + 1;
    L += n_digits(o.lig_start) + 1 + n_digits(o.lig_start + o.l - 1) + 1 + n_digits(o.lig_copy) + 1 + o.l + 1;
    L += n_digits(o.scan_start) + 1 + n_digits(o.scan_stop) + 1 + scan_size + 1;
    L += o.l + mid_len + o.e + 1;
    L += n_digits(f.feature_start - 1) + 1 + n_digits(f.feature_stop) + 1 + 1 + 1 + 3 + 1;
    const int index = first_index + (int)i;
    L += f.label_len + 1 + max(4, n_digits(index)) + 1;
    rec[i] = o;
    len[i] = L;
}

constexpr int kFmtWarps = 8, kFmtStage = 1536;

__global__ void __launch_bounds__(kFmtWarps * 32)
k_fmt_write(const DevRegion *__restrict__ regions, const FmtRegion *__restrict__ fr, const char *__restrict__ strings, const char *__restrict__ ascii,
            const FmtRec *__restrict__ rec, const int64_t *__restrict__ off, int64_t n, const char *__restrict__ middle, int mid_len, int first_index,
            char *__restrict__ out)
{
    __shared__ char stage_all[kFmtWarps][kFmtStage];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *st = stage_all[warp];
    for (int64_t i = (int64_t)blockIdx.x * kFmtWarps + warp; i < n; i += (int64_t)gridDim.x * kFmtWarps) {
        const FmtRec o = rec[i];
        const int64_t o0 = off[i], total = off[i + 1] - o0;
        if (total <= 0 || total > kFmtStage) continue;   // (records longer than the staging buffer are reported by the host side)
        const DevRegion r = regions[o.region];
        const FmtRegion f = fr[o.region];
        const char *seq = ascii + r.seq_off;
        const char *chr = strings + f.chr_off, *label = strings + f.label_off;
        const char sc = o.strand ? '-' : '+';
        const int scan_size = o.scan_stop - o.scan_start + 1;
        // positions of the sequence fields are known from the lengths of everything before them: lane 0 writes the numbers,
        // all lanes copy the sequences
        int at = 0, p_ext = 0, p_lig = 0, p_tgt = 0, p_mip = 0, p_tail = 0;
        if (lane == 0) {
            for (int k = 0; k < f.chr_len; k++) st[at++] = chr[k];
            st[at++] = ':';
            at += put_int(st + at, o.strand ? o.lig_start : o.ext_start); st[at++] = '-';
            at += put_int(st + at, o.strand ? o.ext_start + o.e - 1 : o.lig_start + o.l - 1); st[at++] = '/';
            at += put_int(st + at, o.e); st[at++] = ',';
            at += put_int(st + at, o.l); st[at++] = '/'; st[at++] = sc; st[at++] = '\t';
            for (int k = 0; k < o.glen; k++) st[at++] = o.g[k];
            st[at++] = '\t';
            for (int k = 0; k < f.chr_len; k++) st[at++] = chr[k];
            st[at++] = '\t';
            at += put_int(st + at, o.ext_start); st[at++] = '\t';
            at += put_int(st + at, o.ext_start + o.e - 1); st[at++] = '\t';
            at += put_int(st + at, o.ext_copy); st[at++] = '\t';
            p_ext = at; at += o.e; st[at++] = '\t';
            at += put_int(st + at, o.lig_start); st[at++] = '\t';
            at += put_int(st + at, o.lig_start + o.l - 1); st[at++] = '\t';
            at += put_int(st + at, o.lig_copy); st[at++] = '\t';
            p_lig = at; at += o.l; st[at++] = '\t';
            at += put_int(st + at, o.scan_start); st[at++] = '\t';
            at += put_int(st + at, o.scan_stop); st[at++] = '\t';
            p_tgt = at; at += scan_size; st[at++] = '\t';
            p_mip = at; at += o.l + mid_len + o.e; st[at++] = '\t';
            p_tail = at;
            at += put_int(st + at, f.feature_start - 1); st[at++] = '\t';
            at += put_int(st + at, f.feature_stop); st[at++] = '\t';
            st[at++] = sc; st[at++] = '\t'; st[at++] = '0'; st[at++] = '0'; st[at++] = '0'; st[at++] = '\t';
            for (int k = 0; k < f.label_len; k++) st[at++] = label[k];
            st[at++] = '_';
            const int index = first_index + (int)i;
            for (int z = n_digits(index); z < 4; z++) st[at++] = '0';   // setw(4) << setfill('0')
            at += put_int(st + at, index);
            st[at++] = '\n';
        }
        p_ext = __shfl_sync(0xffffffffu, p_ext, 0); p_lig = __shfl_sync(0xffffffffu, p_lig, 0);
        p_tgt = __shfl_sync(0xffffffffu, p_tgt, 0); p_mip = __shfl_sync(0xffffffffu, p_mip, 0);
        (void)p_tail;
        // sequences in probe orientation: genomic on '+', reverse complement on '-' (characters outside ACGT pass unchanged)
        auto put_seq = [&](int dst, int start, int len) {
            const char *w = seq + (start - r.seq_start);
            for (int k = lane; k < len; k += 32) {
                char c = o.strand ? w[len - 1 - k] : w[k];
                if (o.strand) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
                st[dst + k] = c;
            }
        };
        put_seq(p_ext, o.ext_start, o.e);
        put_seq(p_lig, o.lig_start, o.l);
        put_seq(p_tgt, o.scan_start, scan_size);
        put_seq(p_mip, o.lig_start, o.l);                       // mip_seq = lig + universal middle + ext (mipgen.cpp:605)
        for (int k = lane; k < mid_len; k += 32) st[p_mip + o.l + k] = middle[k];
        put_seq(p_mip + o.l + mid_len, o.ext_start, o.e);
        __syncwarp();
        for (int k = lane; k < (int)total; k += 32) out[o0 + k] = st[k];
        __syncwarp();
    }
}

// %g of an array of doubles (test hook for format_g): out[i*32 ..], len[i]
__global__ void k_fmt_g_test(const double *__restrict__ v, int64_t n, char *__restrict__ out, int *__restrict__ len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char buf[32];
    const int l = format_g(v[i], buf);
    len[i] = l;
    for (int k = 0; k < (l < 0 ? 0 : l); k++) out[i * 32 + k] = buf[k];
}

}  // namespace

extern "C" int mg_format_g(mg_ctx *ctx, const double *values, int64_t n, char *out32, int *len)
{
    if (!ctx || n < 0 || (n > 0 && (!values || !out32 || !len))) return MG_ERR_INVALID;
    if (n == 0) return MG_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *d_v = nullptr;
    char *d_o = nullptr;
    int *d_l = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_v); mg_dev_free(ctx, d_o); mg_dev_free(ctx, d_l); };
#define F_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cudaStreamSynchronize(ctx->stream); cleanup(); return MG_ERR_CUDA; } } while (0)
    F_TRY(mg_dev_alloc(ctx, (void **)&d_v, (size_t)n * 8));
    F_TRY(mg_dev_alloc(ctx, (void **)&d_o, (size_t)n * 32));
    F_TRY(mg_dev_alloc(ctx, (void **)&d_l, (size_t)n * 4));
    F_TRY(cudaMemcpyAsync(d_v, values, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    F_TRY(cudaMemsetAsync(d_o, 0, (size_t)n * 32, ctx->stream));
    k_fmt_g_test<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_v, n, d_o, d_l);
    F_TRY(cudaGetLastError());
    F_TRY(cudaMemcpyAsync(out32, d_o, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    F_TRY(cudaMemcpyAsync(len, d_l, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    F_TRY(cudaStreamSynchronize(ctx->stream));
    cleanup();
    return MG_OK;
}

// n records given by panel-global grid indices (on the host, or already on the device) -> text.  Either into buf (capacity cap) or,
// with buf == NULL, into *text (resized).  Returns the bytes written or a negative status.
static int64_t format_records_core(mg_ctx *ctx, mg_panel *p, const mg_record_meta *meta, const int64_t *idx, bool idx_on_device, int64_t n, int which,
                                   const char *universal_middle, int first_index, char *buf, int64_t cap, std::vector<char> *text)
{
    const double *d_score = which == 1 ? (p->has_svr ? p->d_svr : nullptr) : (p->has_logistic ? p->d_logistic : nullptr);
    if (!d_score) { ctx->err = "mg_panel_format_records: the panel has not been scored with the requested scores"; return MG_ERR_INVALID; }
    if (p->has_sel_inputs) {
        ctx->err = "mg_panel_format_records: regions with TRF / SNP / mappability inputs need design_mip's flags (mipgen.cpp:615-760); format them through the caller's print_details";
        return MG_ERR_INVALID;
    }
    if (!p->d_ascii) { ctx->err = "mg_panel_format_records: the panel keeps no ASCII sequences"; return MG_ERR_INVALID; }
    if (n == 0) { if (text) text->clear(); return 0; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MG_ERR_CUDA;
    // per-region strings
    std::vector<FmtRegion> fr((size_t)p->n_regions);
    std::string pool;
    for (int i = 0; i < p->n_regions; i++) {
        if (!meta[i].chr || !meta[i].label) { ctx->err = "mg_panel_format_records: missing chromosome / label"; return MG_ERR_INVALID; }
        fr[i].chr_off = (int)pool.size(); fr[i].chr_len = (int)strlen(meta[i].chr); pool += meta[i].chr;
        fr[i].label_off = (int)pool.size(); fr[i].label_len = (int)strlen(meta[i].label); pool += meta[i].label;
        fr[i].feature_start = meta[i].feature_start; fr[i].feature_stop = meta[i].feature_stop;
    }
    const int mid_len = (int)strlen(universal_middle);
    pool += universal_middle;
    const int mid_off = (int)pool.size() - mid_len;
    FmtRegion *d_fr = nullptr; char *d_str = nullptr, *d_out = nullptr; int64_t *d_idx_own = nullptr, *d_off = nullptr; FmtRec *d_rec = nullptr; int *d_len = nullptr, *d_status = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_fr); mg_dev_free(ctx, d_str); mg_dev_free(ctx, d_out); mg_dev_free(ctx, d_idx_own); mg_dev_free(ctx, d_off); mg_dev_free(ctx, d_rec); mg_dev_free(ctx, d_len); mg_dev_free(ctx, d_status); };
#define R_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cudaStreamSynchronize(ctx->stream); cleanup(); return MG_ERR_CUDA; } } while (0)
    R_TRY(mg_dev_alloc(ctx, (void **)&d_fr, fr.size() * sizeof(FmtRegion)));
    R_TRY(mg_dev_alloc(ctx, (void **)&d_str, pool.size() + 1));
    R_TRY(mg_dev_alloc(ctx, (void **)&d_off, (size_t)(n + 1) * 8));
    R_TRY(mg_dev_alloc(ctx, (void **)&d_rec, (size_t)n * sizeof(FmtRec)));
    R_TRY(mg_dev_alloc(ctx, (void **)&d_len, (size_t)n * 4));
    R_TRY(mg_dev_alloc(ctx, (void **)&d_status, 4));
    R_TRY(cudaMemcpyAsync(d_fr, fr.data(), fr.size() * sizeof(FmtRegion), cudaMemcpyHostToDevice, ctx->stream));
    R_TRY(cudaMemcpyAsync(d_str, pool.data(), pool.size() + 1, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t *d_idx = idx;
    if (!idx_on_device) {
        R_TRY(mg_dev_alloc(ctx, (void **)&d_idx_own, (size_t)n * 8));
        R_TRY(cudaMemcpyAsync(d_idx_own, idx, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_idx = d_idx_own;
    }
    R_TRY(cudaMemsetAsync(d_status, 0, 4, ctx->stream));
    mg_time_begin(ctx, TM_OTHER, n);
    k_fmt_len<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_cfg, p->d_regions, p->n_regions, d_fr, p->d_copies, d_idx, n, d_score, mid_len,
                                                                   first_index, d_rec, d_len, d_status);
    mg_time_end(ctx);
    R_TRY(cudaGetLastError());
    std::vector<int> len((size_t)n);
    int status = 0;
    R_TRY(cudaMemcpyAsync(len.data(), d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    R_TRY(cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    R_TRY(cudaStreamSynchronize(ctx->stream));
    if (status) {
        ctx->err = status == 1 ? "mg_panel_format_records: a grid index lies outside its region's grid or sequence"
                               : "mg_panel_format_records: a score outside [1e-12, 1e15) in magnitude (format it on the host)";
        cleanup();
        return status == 1 ? MG_ERR_INVALID : MG_ERR_UNSUPPORTED;
    }
    std::vector<int64_t> off((size_t)n + 1);
    off[0] = 0;
    int longest = 0;
    for (int64_t i = 0; i < n; i++) { off[(size_t)i + 1] = off[(size_t)i] + len[(size_t)i]; longest = std::max(longest, len[(size_t)i]); }
    const int64_t total = off[(size_t)n];
    if (longest > kFmtStage) { ctx->err = "mg_panel_format_records: a record exceeds the staging buffer (capture size too large)"; cleanup(); return MG_ERR_UNSUPPORTED; }
    if (buf && total > cap) { ctx->err = "mg_panel_format_records: output buffer too small"; cleanup(); return MG_ERR_INVALID; }
    if (!buf) { text->resize((size_t)total); buf = text->data(); }
    R_TRY(mg_dev_alloc(ctx, (void **)&d_out, (size_t)std::max<int64_t>(total, 1)));
    R_TRY(cudaMemcpyAsync(d_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    mg_time_begin(ctx, TM_OTHER, n);
    const int64_t blocks = std::min<int64_t>((n + kFmtWarps - 1) / kFmtWarps, (int64_t)ctx->sm_count * 8);
    k_fmt_write<<<(unsigned)blocks, kFmtWarps * 32, 0, ctx->stream>>>(p->d_regions, d_fr, d_str, p->d_ascii, d_rec, d_off, n, d_str + mid_off, mid_len,
                                                                     first_index, d_out);
    mg_time_end(ctx);
    R_TRY(cudaGetLastError());
    R_TRY(cudaMemcpyAsync(buf, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    R_TRY(cudaStreamSynchronize(ctx->stream));
    cleanup();
    return total;
}

extern "C" int64_t mg_panel_format_records(mg_ctx *ctx, mg_panel *p, const mg_record_meta *meta, const int64_t *idx, int64_t n, int which,
                                           const char *universal_middle, int first_index, char *buf, int64_t cap)
{
    if (!ctx || !p || p->ctx != ctx || !meta || n < 0 || (n > 0 && (!idx || !buf)) || !universal_middle) return MG_ERR_INVALID;
    if (p->cfg_serial != ctx->cfg_serial) { ctx->err = "the panel was created under an earlier mg_set_config: create it again"; return MG_ERR_INVALID; }
    return format_records_core(ctx, p, meta, idx, false, n, which, universal_middle, first_index, buf, cap, nullptr);
}

// all_mips.txt of a scored panel (mipgen.cpp:474, 488): K-replay finds the grid points the tile loop enumerates, in its order
// (count per scan start -> prefix sum -> fill), the record kernels print them; the text goes to the sink in pieces of at most
// kEnumPiece records so that neither side ever holds more than ~1 GB of it.
extern "C" int64_t mg_panel_format_enumerated(mg_ctx *ctx, mg_panel *p, const mg_record_meta *meta, const mg_select_params *sp,
                                              const char *universal_middle, int first_index, int64_t *records_per_region, mg_text_sink sink,
                                              void *user)
{
    if (!ctx || !p || p->ctx != ctx || !meta || !sp || !universal_middle || !sink) return MG_ERR_INVALID;
    if (p->cfg_serial != ctx->cfg_serial) { ctx->err = "the panel was created under an earlier mg_set_config: create it again"; return MG_ERR_INVALID; }
    const int which = sp->method == 1 ? 1 : 0;   // mixed mode enumerates and prints with the logistic score (mipgen.cpp:467-468)
    const double *d_score = which == 1 ? (p->has_svr ? p->d_svr : nullptr) : (p->has_logistic ? p->d_logistic : nullptr);
    if (!d_score) { ctx->err = "mg_panel_format_enumerated: the panel has not been scored with the scores this method prints"; return MG_ERR_INVALID; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MG_ERR_CUDA;
    const int n = p->n_regions;
    std::vector<int64_t> so((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) so[(size_t)i + 1] = so[(size_t)i] + p->h_regions[i].n_scan;
    const int64_t total_scan = so[(size_t)n];
    for (int i = 0; i < n && records_per_region; i++) records_per_region[i] = 0;
    if (total_scan == 0) return 0;
    int64_t *d_so = nullptr, *d_off = nullptr, *d_idx = nullptr;
    int *d_count = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_so); mg_dev_free(ctx, d_off); mg_dev_free(ctx, d_idx); mg_dev_free(ctx, d_count); };
#define E_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cudaStreamSynchronize(ctx->stream); cleanup(); return MG_ERR_CUDA; } } while (0)
    E_TRY(mg_dev_alloc(ctx, (void **)&d_so, (size_t)(n + 1) * 8));
    E_TRY(mg_dev_alloc(ctx, (void **)&d_off, (size_t)total_scan * 8));
    E_TRY(mg_dev_alloc(ctx, (void **)&d_count, (size_t)total_scan * 4));
    E_TRY(cudaMemcpyAsync(d_so, so.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_replay(ctx, p, d_so, total_scan, d_score, sp, d_count, nullptr, nullptr);
    if (rc != MG_OK) { cleanup(); return rc; }
    std::vector<int> count((size_t)total_scan);
    E_TRY(cudaMemcpyAsync(count.data(), d_count, (size_t)total_scan * 4, cudaMemcpyDeviceToHost, ctx->stream));
    E_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> off((size_t)total_scan);
    int64_t n_rec = 0;
    for (int i = 0; i < n; i++)
        for (int64_t t = so[(size_t)i]; t < so[(size_t)i + 1]; t++) {
            off[(size_t)t] = n_rec;
            n_rec += count[(size_t)t];
            if (records_per_region) records_per_region[i] += count[(size_t)t];
        }
    if (n_rec == 0) { cleanup(); return 0; }
    if ((int64_t)first_index + n_rec > 0x7fffffff) { ctx->err = "mg_panel_format_enumerated: record numbers exceed the int range"; cleanup(); return MG_ERR_INVALID; }
    E_TRY(mg_dev_alloc(ctx, (void **)&d_idx, (size_t)n_rec * 8));
    E_TRY(cudaMemcpyAsync(d_off, off.data(), (size_t)total_scan * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_replay(ctx, p, d_so, total_scan, d_score, sp, nullptr, d_off, d_idx);
    if (rc != MG_OK) { cudaStreamSynchronize(ctx->stream); cleanup(); return rc; }
    E_TRY(cudaStreamSynchronize(ctx->stream));  // off[] is read by the copy above
#undef E_TRY
    const int64_t kEnumPiece = (int64_t)2 << 20;
    std::vector<char> text;
    int64_t bytes = 0;
    for (int64_t r0 = 0; r0 < n_rec; r0 += kEnumPiece) {
        const int64_t m = std::min(kEnumPiece, n_rec - r0);
        const int64_t got = format_records_core(ctx, p, meta, d_idx + r0, true, m, which, universal_middle, first_index + (int)r0, nullptr, 0, &text);
        if (got < 0) { cleanup(); return got; }
        if (sink(user, text.data(), got) != 0) { ctx->err = "mg_panel_format_enumerated: the sink reported an error"; cleanup(); return MG_ERR_INVALID; }
        bytes += got;
    }
    cleanup();
    return bytes;
}

// =====================================================================================================================
// The two FASTQ files check_copy_numbers writes for BWA (mipgen.cpp:798-840), formatted on the device:
//   <project>.all_sequences.fq     one read per (feature, capture size, MIP start): "@<capture>_<chr>_<start>\n<seq>\n+\n###..\n"
//   <project>.oligo_copy_count.fq  one read per (feature, oligo size, start):        "@chr<chr>:<a>-<b>\n<seq>\n+\n###..\n"
// Records of one (feature, size) group are consecutive and their lengths differ only through the digit counts of the
// coordinates, so every record's byte offset has a closed form: one warp per record, no prefix-sum pass.
// =====================================================================================================================
namespace {

struct FqGroup {      // one (feature, size) run of consecutive starts
    int64_t out_off;  // byte offset of the group's first record
    int64_t seq_off;  // offset of the feature's sequence in the ASCII array
    int seq_start;    // chromosome coordinate of sequence index 0
    int first, count; // first start coordinate, number of records
    int size;         // capture size / oligo size
    int chr_off, chr_len;
    int64_t rec0;     // index of the group's first record among all records of the launch
};

// sum over j < k of the decimal digit count of (x0 + j), x0 >= 1
__host__ __device__ inline int64_t digit_sum(int64_t x0, int64_t k)
{
    int64_t s = k, p = 10;
    for (int d = 1; d < 11; d++, p *= 10) {
        const int64_t above = x0 + k - p;  // how many of x0 .. x0+k-1 are >= 10^d
        if (above <= 0) break;
        s += above < k ? above : k;
    }
    return s;
}

__host__ __device__ inline int digits_of(int64_t x)
{
    int n = 1;
    while (x >= 10) { x /= 10; n++; }
    return n;
}

template <bool kOligo>
__global__ void __launch_bounds__(256)
k_fastq(const FqGroup *__restrict__ groups, int n_groups, int64_t n_records, const char *__restrict__ ascii, const char *__restrict__ strings,
        char *__restrict__ out)
{
    __shared__ char head_all[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *head = head_all[warp];
    for (int64_t rec = (int64_t)blockIdx.x * 8 + warp; rec < n_records; rec += (int64_t)gridDim.x * 8) {
        int lo = 0, hi = n_groups - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (groups[mid].rec0 <= rec) lo = mid; else hi = mid - 1;
        }
        const FqGroup g = groups[lo];
        const int64_t k = rec - g.rec0;
        const int start = g.first + (int)k;
        // fixed part of a record: '@' + separators + newline, sequence + newline, "+\n", quality + newline
        int64_t off;
        int hl = 0;
        if (kOligo) {
            const int fixed = 1 + 3 + g.chr_len + 1 + 1 + 1 + 2 * (g.size + 1) + 2;   // "@chr" chr ':' a '-' b '\n' ...
            off = g.out_off + k * fixed + digit_sum(g.first, k) + digit_sum((int64_t)g.first + g.size - 1, k);
            if (lane == 0) {
                head[hl++] = '@'; head[hl++] = 'c'; head[hl++] = 'h'; head[hl++] = 'r';
                for (int i = 0; i < g.chr_len; i++) head[hl++] = strings[g.chr_off + i];
                head[hl++] = ':'; hl += put_int(head + hl, start); head[hl++] = '-'; hl += put_int(head + hl, start + g.size - 1); head[hl++] = '\n';
            }
        } else {
            const int fixed = 1 + digits_of(g.size) + 1 + g.chr_len + 1 + 1 + 2 * (g.size + 1) + 2;   // '@' size '_' chr '_' start '\n' ...
            off = g.out_off + k * fixed + digit_sum(g.first, k);
            if (lane == 0) {
                head[hl++] = '@'; hl += put_int(head + hl, g.size); head[hl++] = '_';
                for (int i = 0; i < g.chr_len; i++) head[hl++] = strings[g.chr_off + i];
                head[hl++] = '_'; hl += put_int(head + hl, start); head[hl++] = '\n';
            }
        }
        hl = __shfl_sync(0xffffffffu, hl, 0);
        __syncwarp();
        char *o = out + off;
        for (int i = lane; i < hl; i += 32) o[i] = head[i];
        const char *seq = ascii + g.seq_off + (start - g.seq_start);
        for (int i = lane; i < g.size; i += 32) { o[hl + i] = seq[i]; o[hl + g.size + 3 + i] = '#'; }
        if (lane == 0) { o[hl + g.size] = '\n'; o[hl + g.size + 1] = '+'; o[hl + g.size + 2] = '\n'; o[hl + 2 * g.size + 3] = '\n'; }
        __syncwarp();
    }
}

// host side shared by both files
int64_t fastq_run(mg_ctx *ctx, bool oligo, const mg_region *regions, const char *const *chr, int n, const int *sizes, int n_sizes, char *buf, int64_t cap)
{
    if (!ctx || n < 0 || n_sizes < 0 || (n > 0 && (!regions || !chr)) || (n_sizes > 0 && !sizes)) return MG_ERR_INVALID;
    std::vector<FqGroup> groups;
    std::string pool;
    int64_t bytes = 0, records = 0, seq_total = 0;
    std::vector<int64_t> seq_off((size_t)n);
    for (int i = 0; i < n; i++) {
        const mg_region &r = regions[i];
        if (!r.seq || r.seq_len <= 0 || !chr[i]) { ctx->err = "mg_format_*_fastq: region without sequence / chromosome name"; return MG_ERR_INVALID; }
        seq_off[(size_t)i] = seq_total;
        seq_total += r.seq_len;
        const int chr_off = (int)pool.size(), chr_len = (int)strlen(chr[i]);
        pool += chr[i];
        for (int s = 0; s < n_sizes; s++) {
            const int size = sizes[s];
            FqGroup g;
            g.size = size; g.chr_off = chr_off; g.chr_len = chr_len; g.seq_off = seq_off[(size_t)i]; g.seq_start = r.seq_start;
            if (oligo) {
                // relative starts 0 .. len - size - 1 (mipgen.cpp:826-828)
                g.first = r.seq_start; g.count = r.seq_len - size;
            } else {
                // current_mip_start from start_flanked - capture while < stop_flanked, kept if > 0 and the read fits (mipgen.cpp:811-823)
                const int a = std::max(1, r.start_flanked - size), b = std::min(r.stop_flanked - 1, r.seq_stop - size + 1);
                g.first = a; g.count = b - a + 1;
                if (g.count > 0 && a < r.seq_start) { ctx->err = "mg_format_capture_fastq: a read starts before the region's sequence"; return MG_ERR_INVALID; }
            }
            if (g.count <= 0) continue;
            g.out_off = bytes; g.rec0 = records;
            const int64_t fixed = oligo ? 1 + 3 + chr_len + 1 + 1 + 1 + 2 * (size + 1) + 2 : 1 + digits_of(size) + 1 + chr_len + 1 + 1 + 2 * (size + 1) + 2;
            bytes += (int64_t)g.count * fixed + digit_sum(g.first, g.count) + (oligo ? digit_sum((int64_t)g.first + size - 1, g.count) : 0);
            records += g.count;
            groups.push_back(g);
        }
    }
    if (!buf) return bytes;   // sizing call
    if (bytes > cap) { ctx->err = "mg_format_*_fastq: output buffer too small"; return MG_ERR_INVALID; }
    if (bytes == 0) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MG_ERR_CUDA;
    std::vector<char> ascii((size_t)seq_total);
    for (int i = 0; i < n; i++) memcpy(&ascii[(size_t)seq_off[(size_t)i]], regions[i].seq, (size_t)regions[i].seq_len);
    FqGroup *d_g = nullptr; char *d_a = nullptr, *d_s = nullptr, *d_o = nullptr;
    auto cleanup = [&]() { mg_dev_free(ctx, d_g); mg_dev_free(ctx, d_a); mg_dev_free(ctx, d_s); mg_dev_free(ctx, d_o); };
#define Q_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(_e); cudaStreamSynchronize(ctx->stream); cleanup(); return MG_ERR_CUDA; } } while (0)
    Q_TRY(mg_dev_alloc(ctx, (void **)&d_g, groups.size() * sizeof(FqGroup)));
    Q_TRY(mg_dev_alloc(ctx, (void **)&d_a, ascii.size()));
    Q_TRY(mg_dev_alloc(ctx, (void **)&d_s, pool.size() + 1));
    Q_TRY(mg_dev_alloc(ctx, (void **)&d_o, (size_t)bytes));
    Q_TRY(cudaMemcpyAsync(d_g, groups.data(), groups.size() * sizeof(FqGroup), cudaMemcpyHostToDevice, ctx->stream));
    Q_TRY(cudaMemcpyAsync(d_a, ascii.data(), ascii.size(), cudaMemcpyHostToDevice, ctx->stream));
    Q_TRY(cudaMemcpyAsync(d_s, pool.data(), pool.size() + 1, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t blocks = std::min<int64_t>((records + 7) / 8, (int64_t)ctx->sm_count * 16);
    mg_time_begin(ctx, TM_OTHER, records);
    if (oligo) k_fastq<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_g, (int)groups.size(), records, d_a, d_s, d_o);
    else k_fastq<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_g, (int)groups.size(), records, d_a, d_s, d_o);
    mg_time_end(ctx);
    Q_TRY(cudaGetLastError());
    Q_TRY(cudaMemcpyAsync(buf, d_o, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    Q_TRY(cudaStreamSynchronize(ctx->stream));
#undef Q_TRY
    cleanup();
    return bytes;
}

}  // namespace

extern "C" int64_t mg_format_capture_fastq(mg_ctx *ctx, const mg_region *regions, const char *const *chr, int n, int max_capture, int min_capture,
                                           int capture_increment, char *buf, int64_t cap)
{
    if (max_capture < min_capture || min_capture <= 0 || capture_increment < 0) return MG_ERR_INVALID;
    std::vector<int> sizes;
    const int inc = capture_increment == 0 ? 1 : capture_increment;  // (a zero increment would never terminate in the reference's loop)
    for (int c = max_capture; c >= min_capture; c -= inc) sizes.push_back(c);
    return fastq_run(ctx, false, regions, chr, n, sizes.data(), (int)sizes.size(), buf, cap);
}

extern "C" int64_t mg_format_oligo_fastq(mg_ctx *ctx, const mg_region *regions, const char *const *chr, int n, const int *oligo_sizes, int n_oligo_sizes,
                                         char *buf, int64_t cap)
{
    return fastq_run(ctx, true, regions, chr, n, oligo_sizes, n_oligo_sizes, buf, cap);
}
