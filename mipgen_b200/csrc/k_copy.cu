// k_copy.cu -- SURVEY.md 8(f4), opt-in: exact-match arm copy counting on the device.
//
// The reference gets the copy number of every arm-sized oligo of a region from BWA: check_copy_numbers writes one read per
// (oligo size, start) into <project>.oligo_copy_count.fq (mipgen.cpp:824-836), find_copy runs `bwa aln` + `bwa samse` on it and
// keeps each read's X0 tag = the number of best hits (mipgen.cpp:558-596; no X0 tag -> 100).  A read cut out of the indexed
// genome has edit distance 0 to its own locus, so its best hits are exactly its exact occurrences on either strand: X0 is an
// exact-match count.  That count is what this file computes, without BWA:
//
//   index   every position of the genome with 32 valid (ACGT, case-insensitive) bases ahead contributes its 2-bit packed
//           32-mer to one sorted array of 64-bit keys; the occurrences of any oligo of L <= 32 bases are then one contiguous
//           range of that array (a prefix range).  Positions closer than 32 bases to a non-ACGT character or a contig end
//           ("short suffixes") go to a small side list that is filtered and sorted per oligo size.
//   query   one thread per (region, oligo size, start): two binary searches per strand in the main array and in the side
//           array of its size; copy = occurrences of the oligo + occurrences of its reverse complement.
//
// Differences from BWA that cannot be closed here (BWA is absent from this image, so this row is "parity unpinned" against it;
// tests pin it to a brute-force restatement, oracle/copy_count.py): BWA also reports hits with mismatches when there is NO exact
// hit, and gives up on reads it cannot place.  Both cases only arise for oligos that do not occur in the genome or that hold
// non-ACGT characters; they get copy 100 here, the value find_copy assigns to a read without an X0 tag.
#include <cub/device/device_radix_sort.cuh>

#include <map>

#include "mg_common.cuh"

struct mg_genome {
    mg_ctx *ctx = nullptr;
    uint64_t *d_main = nullptr;   // sorted 32-mers of the positions with 32 valid bases ahead
    int64_t n_main = 0;
    uint64_t *d_side_key = nullptr;  // short suffixes: bases beyond the valid run are A (00)
    uint8_t *d_side_len = nullptr;   // their valid run lengths (1..31)
    int64_t n_side = 0;
    int64_t n_positions = 0;
    struct Side { uint64_t *d = nullptr; int64_t n = 0; };
    std::map<int, Side> side_of;     // per oligo size: sorted prefixes of the short suffixes that are long enough
};

namespace {

__device__ __forceinline__ int base_code(char c)
{
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

// one thread per position of a contig: its valid run (<= 32) and the packed bases of the run, first base in the top two bits
__global__ void __launch_bounds__(256) k_copy_pack(const char *__restrict__ seq, int64_t n, uint64_t *__restrict__ main_keys,
                                                   unsigned long long *__restrict__ n_main, uint64_t *__restrict__ side_keys,
                                                   uint8_t *__restrict__ side_len, unsigned long long *__restrict__ n_side, int pass)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key = 0;
    int v = 0;
    if (p < n) {
        const int64_t end = p + 32 < n ? p + 32 : n;
        for (int64_t j = p; j < end; j++) {
            const int c = base_code(seq[j]);
            if (c < 0) break;
            key |= (uint64_t)c << (62 - 2 * v);
            v++;
        }
    }
    // warp-aggregated slots (the order of the output does not matter: it is sorted afterwards)
    const unsigned full = __ballot_sync(0xffffffffu, v == 32), part = __ballot_sync(0xffffffffu, v > 0 && v < 32);
    const int lane = threadIdx.x & 31;
    unsigned long long base_f = 0, base_p = 0;
    if (lane == 0) {
        if (full) base_f = atomicAdd(n_main, (unsigned long long)__popc(full));
        if (part) base_p = atomicAdd(n_side, (unsigned long long)__popc(part));
    }
    base_f = __shfl_sync(0xffffffffu, base_f, 0);
    base_p = __shfl_sync(0xffffffffu, base_p, 0);
    if (pass == 0) return;   // counting pass
    const unsigned below = (1u << lane) - 1;
    if (v == 32) main_keys[base_f + __popc(full & below)] = key;
    else if (v > 0) {
        const unsigned long long at = base_p + __popc(part & below);
        side_keys[at] = key;
        side_len[at] = (uint8_t)v;
    }
}

// the short suffixes that are at least L bases long, cut to L bases
__global__ void __launch_bounds__(256) k_copy_side(const uint64_t *__restrict__ key, const uint8_t *__restrict__ len, int64_t n, int L,
                                                   uint64_t *__restrict__ out, unsigned long long *__restrict__ n_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool keep = i < n && len[i] >= L;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(n_out, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) out[base + __popc(m & ((1u << lane) - 1))] = L == 32 ? key[i] : key[i] & ~(~0ull >> (2 * L));
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *__restrict__ a, int64_t n, uint64_t x)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// number of keys in [x, y]
__device__ __forceinline__ int64_t range_count(const uint64_t *__restrict__ a, int64_t n, uint64_t x, uint64_t y)
{
    if (n == 0) return 0;
    const int64_t lo = lower_bound_u64(a, n, x);
    const int64_t hi = y == ~0ull ? n : lower_bound_u64(a, n, y + 1);
    return hi - lo;
}

struct CopyRegion { int64_t seq_off, out_off; int seq_len; };

// one thread per (region, oligo size, start).  out[region.out_off + k * seq_len + i]
__global__ void __launch_bounds__(256) k_copy_query(const char *__restrict__ seqs, const CopyRegion *__restrict__ regions, const int64_t *__restrict__ work_off,
                                                    int n_regions, const int *__restrict__ sizes, int n_sizes, const uint64_t *__restrict__ main_keys,
                                                    int64_t n_main, const uint64_t *const *__restrict__ side_keys, const int64_t *__restrict__ side_n,
                                                    int32_t *__restrict__ out, int64_t total)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total) return;
    int lo = 0, hi = n_regions - 1;   // region of this work item
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (work_off[mid] <= w) lo = mid; else hi = mid - 1;
    }
    const CopyRegion r = regions[lo];
    const int64_t rel = w - work_off[lo];
    const int k = (int)(rel / r.seq_len), i = (int)(rel - (int64_t)k * r.seq_len);
    const int L = sizes[k];
    int32_t result = 0;   // a start the reference never queries (mipgen.cpp:829: start < length - size) is an absent key
    if (L >= 1 && L <= 32 && i < r.seq_len - L) {
        uint64_t f = 0, rc = 0;
        bool ok = true;
        for (int j = 0; j < L; j++) {
            const int c = base_code(seqs[r.seq_off + i + j]);
            if (c < 0) { ok = false; break; }
            f |= (uint64_t)c << (62 - 2 * j);
            rc |= (uint64_t)(3 - c) << (62 - 2 * (L - 1 - j));
        }
        if (!ok) result = 100;
        else {
            const uint64_t tail = L == 32 ? 0ull : ~0ull >> (2 * L);
            int64_t cnt = range_count(main_keys, n_main, f, f | tail) + range_count(main_keys, n_main, rc, rc | tail);
            cnt += range_count(side_keys[k], side_n[k], f, f) + range_count(side_keys[k], side_n[k], rc, rc);
            result = cnt == 0 ? 100 : (int32_t)(cnt > 1000000 ? 1000000 : cnt);
        }
    }
    out[r.out_off + (int64_t)k * r.seq_len + i] = result;
}

int sort_keys(mg_ctx *ctx, uint64_t **d_keys, int64_t n)
{
    if (n <= 1) return MG_OK;
    if (n > 0x7fffffffLL * 2) { ctx->err = "genome index: more positions than one radix sort call takes"; return MG_ERR_UNSUPPORTED; }
    uint64_t *d_alt = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&d_alt, (size_t)n * 8));
    cub::DoubleBuffer<uint64_t> buf(*d_keys, d_alt);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, buf, n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    uint64_t *sorted = buf.Current(), *other = sorted == *d_keys ? d_alt : *d_keys;
    cudaFree(other);
    cudaFree(d_tmp);
    *d_keys = sorted;
    if (e != cudaSuccess) { ctx->err = std::string("genome index sort: ") + cudaGetErrorString(e); return MG_ERR_CUDA; }
    return MG_OK;
}

int side_for(mg_genome *g, int L, mg_genome::Side *out)
{
    auto it = g->side_of.find(L);
    if (it != g->side_of.end()) { *out = it->second; return MG_OK; }
    mg_ctx *ctx = g->ctx;
    mg_genome::Side s;
    if (g->n_side > 0) {
        unsigned long long *d_n = ctx->d_work + 3, h_n = 0;
        CUDA_TRY(ctx, cudaMalloc(&s.d, (size_t)g->n_side * 8));
        CUDA_TRY(ctx, cudaMemsetAsync(d_n, 0, 8, ctx->stream));
        k_copy_side<<<(unsigned)((g->n_side + 255) / 256), 256, 0, ctx->stream>>>(g->d_side_key, g->d_side_len, g->n_side, L, s.d, d_n);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(&h_n, d_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        s.n = (int64_t)h_n;
        const int rc = sort_keys(ctx, &s.d, s.n);
        if (rc != MG_OK) return rc;
    }
    g->side_of[L] = s;
    *out = s;
    return MG_OK;
}

}  // namespace

extern "C" int mg_genome_create(mg_ctx *ctx, const char *const *seqs, const int64_t *lens, int n_contigs, mg_genome **out)
{
    if (!ctx || !out || n_contigs < 0 || (n_contigs > 0 && (!seqs || !lens))) return MG_ERR_INVALID;
    *out = nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t longest = 0, total = 0;
    for (int c = 0; c < n_contigs; c++) {
        if (lens[c] < 0 || (lens[c] > 0 && !seqs[c])) { ctx->err = "mg_genome_create: bad contig"; return MG_ERR_INVALID; }
        longest = lens[c] > longest ? lens[c] : longest;
        total += lens[c];
    }
    mg_genome *g = new mg_genome();
    g->ctx = ctx;
    g->n_positions = total;
    char *d_seq = nullptr;
    unsigned long long *d_cnt = nullptr, h_cnt[2] = {0, 0};
    auto fail = [&](int rc) { cudaFree(d_seq); cudaFree(d_cnt); mg_genome_destroy(g); return rc; };
    auto cuda_ok = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess) return true;
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    };
    if (!cuda_ok(cudaMalloc(&d_seq, (size_t)(longest ? longest : 1)), "genome staging buffer")) return fail(MG_ERR_CUDA);
    if (!cuda_ok(cudaMalloc(&d_cnt, 16), "genome counters")) return fail(MG_ERR_CUDA);
    // pass 0 counts the two kinds of positions, pass 1 writes their keys
    for (int pass = 0; pass < 2; pass++) {
        if (!cuda_ok(cudaMemsetAsync(d_cnt, 0, 16, ctx->stream), "genome counters")) return fail(MG_ERR_CUDA);
        for (int c = 0; c < n_contigs; c++) {
            if (lens[c] == 0) continue;
            if (!cuda_ok(cudaMemcpyAsync(d_seq, seqs[c], (size_t)lens[c], cudaMemcpyHostToDevice, ctx->stream), "contig upload")) return fail(MG_ERR_CUDA);
            k_copy_pack<<<(unsigned)((lens[c] + 255) / 256), 256, 0, ctx->stream>>>(d_seq, lens[c], g->d_main, d_cnt, g->d_side_key, g->d_side_len, d_cnt + 1, pass);
            if (!cuda_ok(cudaGetLastError(), "k_copy_pack")) return fail(MG_ERR_CUDA);
            if (!cuda_ok(cudaStreamSynchronize(ctx->stream), "k_copy_pack")) return fail(MG_ERR_CUDA);   // seqs[c] may be pageable: one contig in flight
        }
        if (pass == 0) {
            if (!cuda_ok(cudaMemcpy(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost), "genome counters")) return fail(MG_ERR_CUDA);
            g->n_main = (int64_t)h_cnt[0];
            g->n_side = (int64_t)h_cnt[1];
            if (!cuda_ok(cudaMalloc(&g->d_main, (size_t)(g->n_main ? g->n_main : 1) * 8), "genome index")) return fail(MG_ERR_NOMEM);
            if (!cuda_ok(cudaMalloc(&g->d_side_key, (size_t)(g->n_side ? g->n_side : 1) * 8), "genome index")) return fail(MG_ERR_NOMEM);
            if (!cuda_ok(cudaMalloc(&g->d_side_len, (size_t)(g->n_side ? g->n_side : 1)), "genome index")) return fail(MG_ERR_NOMEM);
        }
    }
    cudaFree(d_seq); d_seq = nullptr;
    cudaFree(d_cnt); d_cnt = nullptr;
    const int rc = sort_keys(ctx, &g->d_main, g->n_main);
    if (rc != MG_OK) return fail(rc);
    *out = g;
    return MG_OK;
}

extern "C" void mg_genome_destroy(mg_genome *g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaFree(g->d_main);
    cudaFree(g->d_side_key);
    cudaFree(g->d_side_len);
    for (auto &kv : g->side_of) cudaFree(kv.second.d);
    delete g;
}

extern "C" int mg_genome_info(const mg_genome *g, int64_t *positions, int64_t *indexed, int64_t *short_suffixes)
{
    if (!g) return MG_ERR_INVALID;
    if (positions) *positions = g->n_positions;
    if (indexed) *indexed = g->n_main;
    if (short_suffixes) *short_suffixes = g->n_side;
    return MG_OK;
}

extern "C" int mg_count_arm_copies(mg_genome *g, const mg_region *regions, int n, const int *oligo_sizes, int n_oligo_sizes, int32_t *copies,
                                   int64_t *copies_off)
{
    if (!g || n < 0 || n_oligo_sizes < 0 || (n > 0 && !regions) || (n_oligo_sizes > 0 && !oligo_sizes)) return MG_ERR_INVALID;
    mg_ctx *ctx = g->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    for (int k = 0; k < n_oligo_sizes; k++)
        if (oligo_sizes[k] < 1 || oligo_sizes[k] > 32) { ctx->err = "mg_count_arm_copies: oligo sizes must lie in 1..32"; return MG_ERR_UNSUPPORTED; }
    std::vector<CopyRegion> h_regions((size_t)n);
    std::vector<int64_t> work_off((size_t)n + 1, 0);
    int64_t n_seq = 0, n_out = 0;
    for (int i = 0; i < n; i++) {
        if (regions[i].seq_len < 0 || (regions[i].seq_len > 0 && !regions[i].seq)) { ctx->err = "mg_count_arm_copies: bad region"; return MG_ERR_INVALID; }
        h_regions[i].seq_off = n_seq;
        h_regions[i].out_off = n_out;
        h_regions[i].seq_len = regions[i].seq_len;
        if (copies_off) copies_off[i] = n_out;
        n_seq += regions[i].seq_len;
        n_out += (int64_t)n_oligo_sizes * regions[i].seq_len;
        work_off[i + 1] = n_out;
    }
    if (copies_off) copies_off[n] = n_out;
    if (n_out == 0 || !copies) return MG_OK;   // sizing call, or nothing to count
    std::vector<char> h_seq((size_t)n_seq);
    for (int i = 0; i < n; i++) memcpy(h_seq.data() + h_regions[i].seq_off, regions[i].seq, (size_t)regions[i].seq_len);
    std::vector<const uint64_t *> h_side((size_t)n_oligo_sizes);
    std::vector<int64_t> h_side_n((size_t)n_oligo_sizes);
    for (int k = 0; k < n_oligo_sizes; k++) {
        mg_genome::Side s;
        const int rc = side_for(g, oligo_sizes[k], &s);
        if (rc != MG_OK) return rc;
        h_side[k] = s.d;
        h_side_n[k] = s.n;
    }
    char *d_seq = nullptr;
    CopyRegion *d_regions = nullptr;
    int64_t *d_work = nullptr, *d_side_n = nullptr;
    int *d_sizes = nullptr;
    const uint64_t **d_side = nullptr;
    int32_t *d_out = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [&](void **dst, const void *src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = mg_dev_alloc(ctx, dst, bytes ? bytes : 1);
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    up((void **)&d_seq, h_seq.data(), (size_t)n_seq);
    up((void **)&d_regions, h_regions.data(), h_regions.size() * sizeof(CopyRegion));
    up((void **)&d_work, work_off.data(), work_off.size() * 8);
    up((void **)&d_sizes, oligo_sizes, (size_t)n_oligo_sizes * sizeof(int));
    up((void **)&d_side, h_side.data(), h_side.size() * sizeof(void *));
    up((void **)&d_side_n, h_side_n.data(), h_side_n.size() * 8);
    if (e == cudaSuccess) e = mg_dev_alloc(ctx, (void **)&d_out, (size_t)n_out * sizeof(int32_t));
    if (e == cudaSuccess) {
        mg_time_begin(ctx, TM_OTHER, n_out);
        k_copy_query<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(d_seq, d_regions, d_work, n, d_sizes, n_oligo_sizes, g->d_main, g->n_main,
                                                                              d_side, d_side_n, d_out, n_out);
        mg_time_end(ctx);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(copies, d_out, (size_t)n_out * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    mg_dev_free(ctx, d_seq); mg_dev_free(ctx, d_regions); mg_dev_free(ctx, d_work); mg_dev_free(ctx, d_sizes);
    mg_dev_free(ctx, d_side); mg_dev_free(ctx, d_side_n); mg_dev_free(ctx, d_out);
    if (e != cudaSuccess) { ctx->err = std::string("mg_count_arm_copies: ") + cudaGetErrorString(e); return MG_ERR_CUDA; }
    return MG_OK;
}
