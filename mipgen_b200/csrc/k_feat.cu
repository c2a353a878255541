// k_feat.cu -- K-feat (+ fused logistic epilogue), K-encode, K-lrc.
//
// K-feat restates, for the GPU, what the reference does per candidate in
//   tile_regions / design_mip          mipgen.cpp:446-462, 602-613
//   Plus/MinusSVMipv4 geometry         PlusSVMipv4.cpp:7-28, MinusSVMipv4.cpp:6-51
//   SVMipv4::get_parameters            SVMipv4.cpp:60-113   (192 FP64 features)
//   SVMipv4::get_score                 SVMipv4.cpp:114-248  (logistic score)
//
// Two front-ends share the feature layout and the logistic epilogue:
//
// (1) k_feat_window -- the region grid (the hot one).  A CTA takes a WINDOW of consecutive
//     scan starts of one region.  It stages the window's span of the region (<= ~400 bases)
//     in shared memory and builds 88 prefix-count tables over it (64 tri-, 16 di-, 4
//     mono-nucleotides, G+C, N-or-'-', non-ACGT, GC/AT class transitions; uint16).  Every
//     k-mer count of every arm / insert of every candidate of the window is then a difference
//     of two table entries: no per-candidate scanning at all.  A warp writes one candidate's
//     192-double row with six coalesced 256-byte stores (each lane: 6 x {2 LDS.U16, subtract,
//     divide}); afterwards each lane evaluates the logistic model for its own candidate.
//     The minus strand is never materialised: counts are taken on the genomic window and
//     looked up through the reverse-complement k-mer's table.
//     Division: count/(len-k+1) with small integer operands is computed as
//     q0 = c*rn, r = fma(-q0, n, c), q = fma(r, rn, q0) with rn = RN(1/n) from a shared-memory
//     table -- bit-identical to IEEE division for every 0 <= c, 1 <= n <= 4096
//     (exhaustively checked on the CPU: tests/test_division_identity.py).
//
// (2) k_feat_explicit -- explicit, already strand-oriented strings (the SVMipv4 object
//     interface; cfg1 and the drop-in's exception path).  One warp per candidate counts
//     k-mers with one shared-memory atomic per base.
//
// All feature values are a single correctly rounded division of two small integers, and the
// logistic exponent is summed in the reference's order with explicit round-to-nearest
// mul/add (no FMA contraction), so both are bit-identical to the CPU code; only pow()
// differs (CUDA vs glibc, <= 2 ulp).
#include "mg_common.cuh"
#include "logistic_terms.inc"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kHistInts = 128;  // 84 insert bins + 20 ext + 20 lig (+4 pad)
constexpr int kWarpSmemInts = kHistInts + SLOT_COUNT;

struct View {
    const uint8_t *ext, *lig, *tgt;  // windows in genomic orientation
    int ext_n, lig_n, tgt_n;         // characters actually present (substr clamps)
    int ext_len, lig_len, scan_size; // constructor values: the divisors
    int rc;                          // 1: the stored strings are reverse complements of the windows
    int ext_copy, lig_copy;
    const double *lrc;               // [44] or null
    int ok;                          // 0: statically skipped grid point
};

struct Stash {  // what the logistic model needs, kept by lane c for candidate c
    int eg, ec, ea, lg, lc, la, tg, tc, ta;
    int runs, ext_len, lig_len, scan_size, jcode, ext_copy, lig_copy, state;  // state: 0 skipped, 1 invalid, 2 ok
};

__device__ __forceinline__ int base_class(uint32_t c)
{
    // G/C -> 0, A/T -> 1, anything else -> 2   (SVMipv4.cpp:123-139)
    return (c == B_C || c == B_G) ? 0 : ((c == B_A || c == B_T) ? 1 : 2);
}

__device__ __forceinline__ double log_copy(int copy, const double *__restrict__ tab)
{
    // SVMipv4.cpp:109-110 / 173-174.  tab = glibc log10 of 0..100 computed on the host.
    if (copy > 100) return 2.0;
    if (copy < 0) return __longlong_as_double(0xfff8000000000000LL);
    return tab[copy];
}

// Count one arm window into hist[20] (bin = X*5 + (Y|none)).  Returns flag bits: 1 = has 'N', 2 = has '-'.
__device__ __forceinline__ uint32_t count_arm(const uint8_t *__restrict__ w, int n, int *hist, int lane)
{
    uint32_t flags = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
        int p = p0 + lane;
        uint32_t c0 = B_NONE, c1 = B_NONE;
        if (p < n) {
            c0 = w[p];
            if (p + 1 < n) c1 = w[p + 1];
        }
        if (c0 < 4) atomicAdd(&hist[c0 * 5 + (c1 < 4 ? c1 : 4)], 1);
        if (__any_sync(0xffffffffu, c0 == B_N)) flags |= 1;
        if (__any_sync(0xffffffffu, c0 == B_DASH)) flags |= 2;
    }
    return flags;
}

// Process one candidate with the whole warp.  sm: this warp's kWarpSmemInts ints.
__device__ __forceinline__ void warp_candidate(const View &v, int *sm, int lane, const uint32_t (&fd)[6],
                                               const double *__restrict__ logtab, double *__restrict__ xrow,
                                               Stash &st, bool mine)
{
    int *hist_ins = sm, *hist_ext = sm + 84, *hist_lig = sm + 104, *slot = sm + kHistInts;

    if (!v.ok) {
        if (mine) st.state = 0;
        if (xrow) {
#pragma unroll
            for (int m = 0; m < 6; m++) xrow[lane + 32 * m] = 0.0;
        }
        return;
    }

#pragma unroll
    for (int i = 0; i < kHistInts / 32; i++) sm[lane + 32 * i] = 0;
    __syncwarp();

    // ---- insert: extended-trimer histogram, class transitions, "other" detection ----
    int trans = 0;
    bool any_other = false;
    for (int p0 = 0; p0 < v.tgt_n; p0 += 32) {
        int p = p0 + lane;
        uint32_t c0 = B_NONE, c1 = B_NONE, c2 = B_NONE, cm = B_NONE;
        if (p < v.tgt_n) {
            c0 = v.tgt[p];
            if (p + 1 < v.tgt_n) c1 = v.tgt[p + 1];
            if (p + 2 < v.tgt_n) c2 = v.tgt[p + 2];
            if (p > 0) cm = v.tgt[p - 1];
        }
        if (c0 < 4) {
            int bin = (c1 < 4) ? (int)(c0 * 21 + c1 * 5 + (c2 < 4 ? c2 : 4)) : (int)(c0 * 21 + 20);
            atomicAdd(&hist_ins[bin], 1);
        }
        bool in = p < v.tgt_n;
        trans += __popc(__ballot_sync(0xffffffffu, in && p > 0 && base_class(c0) != base_class(cm)));
        any_other |= __any_sync(0xffffffffu, in && c0 >= 4) != 0;
    }
    uint32_t fe = count_arm(v.ext, v.ext_n, hist_ext, lane);
    uint32_t fl = count_arm(v.lig, v.lig_n, hist_lig, lane);
    __syncwarp();

    // ---- derive count slots (genomic orientation) ----
#pragma unroll
    for (int t = lane; t < 64; t += 32) slot[SLOT_INS_TRI + t] = hist_ins[(t >> 4) * 21 + ((t >> 2) & 3) * 5 + (t & 3)];
    if (lane < 16) {
        int X = lane >> 2, Y = lane & 3, s = 0;
#pragma unroll
        for (int z = 0; z < 5; z++) s += hist_ins[X * 21 + Y * 5 + z];
        slot[SLOT_INS_DI + lane] = s;
        slot[SLOT_EXT_DI + lane] = hist_ext[X * 5 + Y];
        slot[SLOT_LIG_DI + lane] = hist_lig[X * 5 + Y];
    } else if (lane < 20) {
        int X = lane - 16, s = hist_ins[X * 21 + 20], se = 0, sl = 0;
#pragma unroll
        for (int y = 0; y < 20; y++) s += hist_ins[X * 21 + y];
#pragma unroll
        for (int y = 0; y < 5; y++) { se += hist_ext[X * 5 + y]; sl += hist_lig[X * 5 + y]; }
        slot[SLOT_INS_MONO + X] = s;
        slot[SLOT_EXT_MONO + X] = se;
        slot[SLOT_LIG_MONO + X] = sl;
    }
    __syncwarp();
    if (lane == 0) {
        slot[SLOT_INS_GC] = slot[SLOT_INS_MONO + B_C] + slot[SLOT_INS_MONO + B_G];
        slot[SLOT_EXT_GC] = slot[SLOT_EXT_MONO + B_C] + slot[SLOT_EXT_MONO + B_G];
        slot[SLOT_LIG_GC] = slot[SLOT_LIG_MONO + B_C] + slot[SLOT_LIG_MONO + B_G];
    }
    __syncwarp();

    // ---- ligation junction: first two characters of the (oriented) ligation arm ----
    int jcode = -1;
    if (v.lig_n >= 2) {
        uint32_t j0 = v.rc ? v.lig[v.lig_n - 1] : v.lig[0];
        uint32_t j1 = v.rc ? v.lig[v.lig_n - 2] : v.lig[1];
        if (j0 < 4 && j1 < 4) jcode = v.rc ? (int)((3 - j0) * 4 + (3 - j1)) : (int)(j0 * 4 + j1);
    }
    // 'N' in an arm, or '-' in mip_seq (== '-' in an arm)   SVMipv4.cpp:63, 116
    bool invalid = ((fe | fl) & 3) != 0;

    // ---- run count (SVMipv4.cpp:118-141) ----
    int runs = trans + 1;
    if (any_other && !invalid) {
        // characters outside ACGT make the reference's state machine order dependent:
        // replay it literally, in stored-string order (lane 0; rare path)
        int r = 0;
        if (lane == 0 && v.tgt_n > 0) {
            int n = v.tgt_n < v.scan_size ? v.tgt_n : v.scan_size;
            int last = base_class(v.rc ? v.tgt[v.tgt_n - 1] : v.tgt[0]);
            for (int i = 1; i < n; i++) {
                int cur = base_class(v.rc ? v.tgt[v.tgt_n - 1 - i] : v.tgt[i]);
                if (cur == 0) { if (last != 0) { r++; last = 0; } }
                else { if (last != 1) { r++; last = cur; } }
            }
        }
        runs = __shfl_sync(0xffffffffu, r, 0) + 1;
    }

    if (mine) {
        const int G = v.rc ? B_C : B_G, Cc = v.rc ? B_G : B_C, A = v.rc ? B_T : B_A;
        st.eg = slot[SLOT_EXT_MONO + G]; st.ec = slot[SLOT_EXT_MONO + Cc]; st.ea = slot[SLOT_EXT_MONO + A];
        st.lg = slot[SLOT_LIG_MONO + G]; st.lc = slot[SLOT_LIG_MONO + Cc]; st.la = slot[SLOT_LIG_MONO + A];
        st.tg = slot[SLOT_INS_MONO + G]; st.tc = slot[SLOT_INS_MONO + Cc]; st.ta = slot[SLOT_INS_MONO + A];
        st.runs = runs; st.ext_len = v.ext_len; st.lig_len = v.lig_len; st.scan_size = v.scan_size;
        st.jcode = jcode; st.ext_copy = v.ext_copy; st.lig_copy = v.lig_copy;
        st.state = invalid ? 1 : 2;
    }

    // ---- the 192-double feature row ----
    if (xrow) {
#pragma unroll
        for (int m = 0; m < 6; m++) {
            uint32_t d = fd[m];
            uint32_t kind = d & 7, part = (d >> 3) & 3, km1 = (d >> 5) & 3, j = (d >> 23) & 255;
            int len = part == 0 ? v.ext_len : (part == 1 ? v.scan_size : v.lig_len);
            double val;
            if (invalid) val = 0.0;  // SVMipv4.cpp:63-68: 192 zeros
            else if (kind == FK_RATIO) {
                int s = v.rc ? (int)((d >> 15) & 255) : (int)((d >> 7) & 255);
                val = __ddiv_rn((double)slot[s], (double)(len - (int)km1));
            } else if (kind == FK_LEN) val = (double)len;
            else if (kind == FK_LRC) val = v.lrc ? v.lrc[j] : 0.0;
            else if (kind == FK_JUNC) val = (jcode == (int)j) ? 1.0 : 0.0;
            else val = log_copy(j == 0 ? v.ext_copy : v.lig_copy, logtab);
            xrow[lane + 32 * m] = val;
        }
    }
    __syncwarp();
}

// The logistic model, one thread per candidate (SVMipv4.cpp:142-247).
__device__ __noinline__ double logistic_score(const Stash &s, const double *__restrict__ logtab)
{
    if (s.state == 0) return __longlong_as_double(0x7ff8000000000000LL);
    if (s.state == 1) return -1000.0;
    double v[MG_LOGIT_NVARS];
    const double ext_length = (double)s.ext_len, lig_length = (double)s.lig_len, scan = (double)s.scan_size;
    v[MG_V_BASES_PER_SWITCH] = __ddiv_rn(scan, (double)s.runs);
    v[MG_V_EXT_LENGTH] = ext_length;
    v[MG_V_LIG_LENGTH] = lig_length;
    v[MG_V_TARGET_LENGTH] = s.scan_size > 250 ? 250.0 : scan;
    v[MG_V_EXT_GC_CONTENT] = __ddiv_rn((double)(s.ec + s.eg), ext_length);
    v[MG_V_LIG_GC_CONTENT] = __ddiv_rn((double)(s.lc + s.lg), lig_length);
    v[MG_V_TARGET_GC_CONTENT] = __ddiv_rn((double)(s.tc + s.tg), scan);
    v[MG_V_EXT_G_CONTENT] = __ddiv_rn((double)s.eg, ext_length);
    v[MG_V_LIG_G_CONTENT] = __ddiv_rn((double)s.lg, lig_length);
    v[MG_V_TARGET_G_CONTENT] = __ddiv_rn((double)s.tg, scan);
    v[MG_V_EXT_A_CONTENT] = __ddiv_rn((double)s.ea, ext_length);
    v[MG_V_LIG_A_CONTENT] = __ddiv_rn((double)s.la, lig_length);
    v[MG_V_TARGET_A_CONTENT] = __ddiv_rn((double)s.ta, scan);
    // junction_scores (SVMipv4.cpp:249-267, data); unknown key -> 0.0 (:171)
    const double JT[16] = {0.0, 0.35, 0.046, 0.079, 0.34, 0.22, 0.55, -0.071,
                           0.35, 0.92, 0.24, 0.48, -0.46, -0.35, -0.25, -0.98};
    double js = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) js = (s.jcode == i) ? JT[i] : js;
    v[MG_V_JUNCTION_SCORE] = js;
    v[MG_V_LOG_EXT_COPY] = log_copy(s.ext_copy, logtab);
    v[MG_V_LOG_LIG_COPY] = log_copy(s.lig_copy, logtab);

    double ex = __dsub_rn(MG_LOGIT_C0, (double)MG_LOGIT_C1);
#define LIN(c, a, b) __dmul_rn((c), v[a])
#define PROD(c, a, b) __dmul_rn(__dmul_rn((c), v[a]), v[b])
#define SQ(c, a, b) __dmul_rn((c), __dmul_rn(v[a], v[a]))
#define X(c, kind, a, b) ex = __dadd_rn(ex, kind(c, a, b));
    MG_LOGIT_TERMS(X)
#undef X
#undef LIN
#undef PROD
#undef SQ
    double p = pow(2.71828, ex);  // the literal 2.71828, not e (SVMipv4.cpp:247)
    return __ddiv_rn(p, __dadd_rn(1.0, p));
}

// ------------------------------- grid front-end -------------------------------
// prefix-table rows
constexpr int PF_DI = 64, PF_MONO = 80, PF_GC = 84, PF_NDASH = 85, PF_OTHER = 86, PF_TRANS = 87, PF_ROWS = 88;
constexpr int kRecipN = 512;
// per-warp candidate records (field-major, 32 candidates): 11 ints + 2 doubles
enum { GW_FLAGS = 0, GW_JCODE, GW_EXT_A, GW_EXT_N, GW_EXT_LEN, GW_LIG_A, GW_LIG_N, GW_LIG_LEN, GW_TGT_A, GW_TGT_N, GW_SCAN, GW_GOFF,
       GW_LCE = 12, GW_LCL = 14, GW_FIELDS = 16 };

struct WinSmem {
    const uint16_t *P;   // [PF_ROWS][stride]
    int stride;
    const uint8_t *codes;  // span codes (+3 sentinels)
    const double *rn;    // [kRecipN] RN(1/n)
    const double *lrc;   // [44]
    int span_len;
};

// occurrences of the k-mer of table `row` that lie inside [a, a+n)  (span-relative)
__device__ __forceinline__ int pf_count(const WinSmem &w, int row, int a, int n, int k)
{
    const int end = a + n - k + 1;
    if (end <= a) return 0;
    const uint16_t *r = w.P + row * w.stride;
    return (int)(uint16_t)(r[end] - r[a]);
}

struct Geo {
    int ok;                        // passes the static skips and lies inside the sequence
    int rc;                        // strand
    int ext_a, lig_a, tgt_a;       // span-relative window starts
    int ext_n, lig_n, tgt_n;       // characters present
    int ext_len, lig_len, scan_size;
    int ext_start, lig_start;      // chromosome coordinates (copy look-up)
};

struct WinCfg {  // per-kernel constants, read once
    int max_capture, min_capture, inc, max_mip_overlap, n_cap, n_pairs;
    const int *pair_e, *pair_l;  // shared-memory copies of the arm-pair table
};

// geometry of grid point (scan index si, capture index ci, pair p, strand) -- no divisions
__device__ __forceinline__ Geo candidate_geometry(const WinCfg &c, const DevRegion &r, int span0, int si, int ci, int p, int strand)
{
    Geo g;
    g.rc = strand;
    const int s = r.first_scan + si, cap = c.max_capture - ci * c.inc;
    const int e = c.pair_e[p], l = c.pair_l[p];
    // static skips: mipgen.cpp:429, 443, 444
    bool ok = !(cap > r.stop_flanked - r.start_flanked + c.max_mip_overlap && cap - c.inc >= c.min_capture);
    ok = ok && !(s - e <= 0 || s - l <= 0);
    ok = ok && !(s + cap - e - 1 > r.seq_stop || s + cap - l - 1 > r.seq_stop);
    const int t = s + cap - (e + l) - 1;  // scan_stop (:449)
    g.ext_len = e; g.lig_len = l; g.scan_size = t - s + 1;
    g.ext_start = strand ? t + 1 : s - e;   // Plus/MinusSVMipv4 ctors
    g.lig_start = strand ? s - l : t + 1;
    const int eo = g.ext_start - r.seq_start, lo = g.lig_start - r.seq_start, to = s - r.seq_start;
    // std::string::substr(off,len) throws for off > size; it clamps the length otherwise
    ok = ok && eo >= 0 && lo >= 0 && to >= 0 && eo <= r.seq_len && lo <= r.seq_len && to <= r.seq_len && g.scan_size >= 0;
    g.ext_n = min(e, r.seq_len - eo); g.lig_n = min(l, r.seq_len - lo); g.tgt_n = min(g.scan_size, r.seq_len - to);
    g.ext_a = eo - span0; g.lig_a = lo - span0; g.tgt_a = to - span0;
    ok = ok && g.ext_a >= 0 && g.lig_a >= 0 && g.tgt_a >= 0;  // always true for statically valid candidates
    g.ok = ok;
    return g;
}

__device__ __forceinline__ int copy_lookup(const DevConfig *__restrict__ cfg, const DevRegion &r,
                                           const int *__restrict__ copies, int start, int len)
{
    if (r.copy_off < 0) return 1;
    for (int k = 0; k < cfg->n_oligo; k++)
        if (cfg->oligo_sizes[k] == len) {
            int i = start - r.seq_start;
            if (i < 0 || i >= r.seq_len) return 0;
            return copies[r.copy_off + (int64_t)k * r.seq_len + i];
        }
    return 0;  // absent key: map::operator[] yields 0 (mipgen.cpp:612-613)
}

__device__ __forceinline__ int junction_code(const WinSmem &w, const Geo &g)
{
    if (g.lig_n < 2) return -1;
    const uint32_t j0 = g.rc ? w.codes[g.lig_a + g.lig_n - 1] : w.codes[g.lig_a];
    const uint32_t j1 = g.rc ? w.codes[g.lig_a + g.lig_n - 2] : w.codes[g.lig_a + 1];
    if (j0 >= 4 || j1 >= 4) return -1;
    return g.rc ? (int)((3 - j0) * 4 + (3 - j1)) : (int)(j0 * 4 + j1);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
k_feat_window(const DevConfig *__restrict__ cfg, const DevRegion *__restrict__ regions, const DevTask *__restrict__ tasks,
              int task0, int task1, const uint8_t *__restrict__ codes, const double *__restrict__ lrc_all,
              const int *__restrict__ copies, const uint32_t *__restrict__ fdesc, const double *__restrict__ logtab,
              int64_t g_base, uint8_t *__restrict__ valid, double *__restrict__ logistic, double *__restrict__ x, int stride,
              uint8_t *__restrict__ state, double *__restrict__ rows, const DevFact *__restrict__ fc, int ftask_base, double gamma)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double *rn = reinterpret_cast<double *>(smem_raw);            // [kRecipN]
    double *lrc_s = rn + kRecipN;                                  // [44]
    int *geo_s = reinterpret_cast<int *>(lrc_s + MG_NLRC);         // [warps][GW_FIELDS][32]
    int *pair_e = geo_s + kWarpsPerBlock * GW_FIELDS * 32;         // [n_pairs]
    int *pair_l = pair_e + cfg->n_pairs;                           // [n_pairs]
    uint16_t *P = reinterpret_cast<uint16_t *>(pair_l + cfg->n_pairs + (cfg->n_pairs & 1) * 2);   // [PF_ROWS][stride]
    uint8_t *codes_s = reinterpret_cast<uint8_t *>(P + PF_ROWS * stride);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t fd[6];
#pragma unroll
    for (int m = 0; m < 6; m++) fd[m] = fdesc[lane + 32 * m];
    for (int i = threadIdx.x; i < kRecipN; i += blockDim.x) rn[i] = i ? __ddiv_rn(1.0, (double)i) : 0.0;
    for (int i = threadIdx.x; i < cfg->n_pairs; i += blockDim.x) { pair_e[i] = cfg->ext_len[i]; pair_l[i] = cfg->lig_len[i]; }
    WinCfg wc;
    wc.max_capture = cfg->max_capture; wc.min_capture = cfg->min_capture; wc.inc = cfg->inc; wc.max_mip_overlap = cfg->max_mip_overlap;
    wc.n_cap = cfg->n_cap; wc.n_pairs = cfg->n_pairs; wc.pair_e = pair_e; wc.pair_l = pair_l;
    // per-lane feature constants: table-row offsets for both strands
    int rowoff_f[6], rowoff_r[6];
#pragma unroll
    for (int m = 0; m < 6; m++) { rowoff_f[m] = (int)((fd[m] >> 7) & 255) * stride; rowoff_r[m] = (int)((fd[m] >> 15) & 255) * stride; }

    const int max_arm = cfg->max_arm, min_arm = cfg->min_arm;
    // row-table mode: prefix-table offsets and k-1 of the arm / insert ratio features, by (role, strand):
    //   [0,84) arm table offsets [role][strand][21]; [84,126) arm k-1 [role][21]; [126,296) insert offsets [strand][85]; [296,381) insert k-1
    __shared__ int rt_tab[384];
    if (rows) {
        for (int i = threadIdx.x; i < 381; i += blockDim.x) {
            int f, strand = 0, want_km1 = 0;
            if (i < 84) { const int role = i / 42, k = i % 21; strand = (i / 21) & 1; f = role ? 152 + k : k; }
            else if (i < 126) { const int role = (i - 84) / 21, k = (i - 84) % 21; f = role ? 152 + k : k; want_km1 = 1; }
            else if (i < 296) { strand = (i - 126) / 85; f = 66 + (i - 126) % 85; }
            else { f = 66 + (i - 296); want_km1 = 1; }
            const uint32_t d = fdesc[f];
            rt_tab[i] = want_km1 ? (int)((d >> 5) & 3) : (int)((strand ? (d >> 15) : (d >> 7)) & 255) * stride;
        }
    }

    for (int ti = task0 + blockIdx.x; ti < task1; ti += gridDim.x) {
        const DevTask tk = tasks[ti];
        const DevRegion r = regions[tk.region];
        // span of the region this window can touch: [first scan start - longest arm, last scan start + max capture - shortest arm)
        const int s_first = r.first_scan + tk.si0, s_last = s_first + tk.nsi - 1;
        int span0 = s_first - max_arm - r.seq_start;
        if (span0 < 0) span0 = 0;
        int span1 = s_last + cfg->max_capture - min_arm - r.seq_start;
        if (span1 > r.seq_len) span1 = r.seq_len;
        const int span_len = span1 > span0 ? span1 - span0 : 0;
        __syncthreads();  // previous task's tables are no longer read
        for (int i = threadIdx.x; i < span_len + 3; i += blockDim.x) codes_s[i] = i < span_len ? codes[r.seq_off + span0 + i] : (uint8_t)B_NONE;
        if (threadIdx.x < MG_NLRC) lrc_s[threadIdx.x] = lrc_all ? lrc_all[(int64_t)tk.region * MG_NLRC + threadIdx.x] : 0.0;
        __syncthreads();
        // prefix tables: warp w owns rows w, w+8, ...; 32 positions at a time, the running
        // count of a row is carry + popc(ballot(hit) & lanes-below-or-equal)
        {
            uint32_t carry[(PF_ROWS + kWarpsPerBlock - 1) / kWarpsPerBlock];
#pragma unroll
            for (int i = 0; i < (PF_ROWS + kWarpsPerBlock - 1) / kWarpsPerBlock; i++) carry[i] = 0;
            if (lane == 0)
                for (int row = warp; row < PF_ROWS; row += kWarpsPerBlock) P[row * stride] = 0;
            const uint32_t le_mask = 0xffffffffu >> (31 - lane);
            for (int p0 = 0; p0 < span_len; p0 += 32) {
                const int p = p0 + lane;
                const bool in = p < span_len;
                const uint32_t c0 = in ? codes_s[p] : (uint32_t)B_NONE, c1 = in ? codes_s[p + 1] : (uint32_t)B_NONE,
                               c2 = in ? codes_s[p + 2] : (uint32_t)B_NONE, cm = p > 0 && in ? codes_s[p - 1] : (uint32_t)B_NONE;
                const int tri = (c0 < 4 && c1 < 4 && c2 < 4) ? (int)(c0 * 16 + c1 * 4 + c2) : -1;
                const int di = (c0 < 4 && c1 < 4) ? (int)(c0 * 4 + c1) : -1;
#pragma unroll
                for (int i = 0; i < (PF_ROWS + kWarpsPerBlock - 1) / kWarpsPerBlock; i++) {
                    const int row = warp + i * kWarpsPerBlock;
                    if (row < PF_ROWS) {
                        bool hit;
                        if (row < PF_DI) hit = tri == row;
                        else if (row < PF_MONO) hit = di == row - PF_DI;
                        else if (row < PF_GC) hit = (int)c0 == row - PF_MONO;
                        else if (row == PF_GC) hit = c0 == B_C || c0 == B_G;
                        else if (row == PF_NDASH) hit = c0 == B_N || c0 == B_DASH;
                        else if (row == PF_OTHER) hit = c0 >= 4 && in;
                        else hit = p > 0 && in && base_class(c0) != base_class(cm);
                        const uint32_t m = __ballot_sync(0xffffffffu, hit);
                        if (in) P[row * stride + p + 1] = (uint16_t)(carry[i] + __popc(m & le_mask));
                        carry[i] += __popc(m);
                    }
                }
            }
        }
        __syncthreads();

        // long-range content values this lane writes in groups 0, 1, 2 (constant for the whole window)
        const double lrc_r0 = lane >= 22 ? lrc_s[lane - 22] : 0.0, lrc_r1 = lrc_s[10 + lane], lrc_r2 = lane < 2 ? lrc_s[42 + lane] : 0.0;
        WinSmem w;
        w.P = P; w.stride = stride; w.codes = codes_s; w.rn = rn; w.lrc = lrc_s; w.span_len = span_len;
        const int per_scan = tk.nci * wc.n_pairs * 2;  // this work item's grid points per scan start
        const int n_c = tk.nsi * per_scan;
        int *gw = geo_s + warp * (GW_FIELDS * 32);  // this warp's candidate records, field-major
        for (int blk = warp; blk * 32 < n_c; blk += kWarpsPerBlock) {
            // ---- phase 1: lane c works out candidate c (geometry, validity, junction, copies, logistic) ----
            const int j = blk * 32 + lane;
            Geo g;
            g.ok = 0;
            int flags = 0, jcode = -1, goff = 0;
            double lce = 0.0, lcl = 0.0;  // log10(1)
            if (j < n_c) {
                const int si = j / per_scan;
                const int rem = j - si * per_scan;
                const int cr = (rem >> 1) / wc.n_pairs, pp = (rem >> 1) - cr * wc.n_pairs, ci = tk.ci0 + cr;
                // offset of this grid point from the window's first one, in reference enumeration order
                goff = ((si * wc.n_cap + ci) * wc.n_pairs + pp) * 2 + (rem & 1);
                g = candidate_geometry(wc, r, span0, tk.si0 + si, ci, pp, rem & 1);
                int ext_copy = 1, lig_copy = 1;
                bool invalid = false;
                if (g.ok) {
                    // 'N' in an arm, or '-' in mip_seq (== '-' in an arm)   SVMipv4.cpp:63-68, 116
                    invalid = pf_count(w, PF_NDASH, g.ext_a, g.ext_n, 1) + pf_count(w, PF_NDASH, g.lig_a, g.lig_n, 1) > 0;
                    jcode = junction_code(w, g);
                    if (r.copy_off >= 0) {
                        ext_copy = copy_lookup(cfg, r, copies, g.ext_start, g.ext_len);
                        lig_copy = copy_lookup(cfg, r, copies, g.lig_start, g.lig_len);
                        lce = log_copy(ext_copy, logtab);
                        lcl = log_copy(lig_copy, logtab);
                    }
                    // every divisor len-2 .. len inside the reciprocal table?
                    const bool fast = g.ext_len > 2 && g.lig_len > 2 && g.scan_size > 3 && g.ext_len < kRecipN && g.lig_len < kRecipN &&
                                      g.scan_size < kRecipN;
                    flags = 1 | (invalid ? 2 : 0) | (fast ? 4 : 0) | (g.rc ? 8 : 0);
                }
                if (valid) valid[tk.g0 + goff] = (uint8_t)g.ok;
                if (state) state[tk.g0 + goff] = (uint8_t)(!g.ok ? 0 : (invalid ? 1 : 2));
                if (logistic) {
                    Stash st;
                    st.state = !g.ok ? 0 : (invalid ? 1 : 2);
                    if (st.state == 2) {
                        const int G = PF_MONO + (g.rc ? B_C : B_G), Cc = PF_MONO + (g.rc ? B_G : B_C), A = PF_MONO + (g.rc ? B_T : B_A);
                        st.eg = pf_count(w, G, g.ext_a, g.ext_n, 1); st.ec = pf_count(w, Cc, g.ext_a, g.ext_n, 1); st.ea = pf_count(w, A, g.ext_a, g.ext_n, 1);
                        st.lg = pf_count(w, G, g.lig_a, g.lig_n, 1); st.lc = pf_count(w, Cc, g.lig_a, g.lig_n, 1); st.la = pf_count(w, A, g.lig_a, g.lig_n, 1);
                        st.tg = pf_count(w, G, g.tgt_a, g.tgt_n, 1); st.tc = pf_count(w, Cc, g.tgt_a, g.tgt_n, 1); st.ta = pf_count(w, A, g.tgt_a, g.tgt_n, 1);
                        // run count (SVMipv4.cpp:118-141)
                        const int nt = min(g.tgt_n, g.scan_size);
                        int runs;
                        if (pf_count(w, PF_OTHER, g.tgt_a, nt, 1) == 0) {
                            const uint16_t *tr = P + PF_TRANS * stride;
                            runs = 1 + (nt >= 2 ? (int)(uint16_t)(tr[g.tgt_a + nt] - tr[g.tgt_a + 1]) : 0);
                        } else {
                            // characters outside ACGT make the reference's state machine order dependent:
                            // replay it literally, in stored-string order (rare path)
                            int rr = 0;
                            if (nt > 0) {
                                int last = base_class(g.rc ? codes_s[g.tgt_a + g.tgt_n - 1] : codes_s[g.tgt_a]);
                                for (int i = 1; i < nt; i++) {
                                    const int cur = base_class(g.rc ? codes_s[g.tgt_a + g.tgt_n - 1 - i] : codes_s[g.tgt_a + i]);
                                    if (cur == 0) { if (last != 0) { rr++; last = 0; } }
                                    else { if (last != 1) { rr++; last = cur; } }
                                }
                            }
                            runs = rr + 1;
                        }
                        st.runs = runs; st.ext_len = g.ext_len; st.lig_len = g.lig_len; st.scan_size = g.scan_size;
                        st.jcode = jcode; st.ext_copy = ext_copy; st.lig_copy = lig_copy;
                    }
                    logistic[tk.g0 + goff] = logistic_score(st, logtab);
                }
            }
            if (!x) continue;

            // ---- phase 2: the warp writes the 32 feature rows, reading each record by broadcast ----
            __syncwarp();
            gw[GW_FLAGS * 32 + lane] = flags; gw[GW_JCODE * 32 + lane] = jcode; gw[GW_GOFF * 32 + lane] = goff;
            gw[GW_EXT_A * 32 + lane] = g.ext_a; gw[GW_EXT_N * 32 + lane] = g.ext_n; gw[GW_EXT_LEN * 32 + lane] = g.ext_len;
            gw[GW_LIG_A * 32 + lane] = g.lig_a; gw[GW_LIG_N * 32 + lane] = g.lig_n; gw[GW_LIG_LEN * 32 + lane] = g.lig_len;
            gw[GW_TGT_A * 32 + lane] = g.tgt_a; gw[GW_TGT_N * 32 + lane] = g.tgt_n; gw[GW_SCAN * 32 + lane] = g.scan_size;
            reinterpret_cast<double *>(gw + GW_LCE * 32)[lane] = lce;
            reinterpret_cast<double *>(gw + GW_LCL * 32)[lane] = lcl;
            __syncwarp();
            const int jend = min(32, n_c - blk * 32);
            double *xbase = x + (tk.g0 - g_base) * MG_NFEAT + lane;
            for (int c = 0; c < jend; c++) {
                const int fl = gw[GW_FLAGS * 32 + c];
                double *xrow = xbase + (int64_t)gw[GW_GOFF * 32 + c] * MG_NFEAT;
                if (!(fl & 1) || (fl & 2)) {
                    // statically skipped grid point (never read) or invalid candidate: 192 zeros
#pragma unroll
                    for (int m = 0; m < 6; m++) xrow[32 * m] = 0.0;
                    continue;
                }
                const bool rc = fl & 8;
                const int jc = gw[GW_JCODE * 32 + c];
                const int ext_a = gw[GW_EXT_A * 32 + c], ext_n = gw[GW_EXT_N * 32 + c], ext_len = gw[GW_EXT_LEN * 32 + c];
                const int lig_a = gw[GW_LIG_A * 32 + c], lig_n = gw[GW_LIG_N * 32 + c], lig_len = gw[GW_LIG_LEN * 32 + c];
                const int tgt_a = gw[GW_TGT_A * 32 + c], tgt_n = gw[GW_TGT_N * 32 + c], scan_size = gw[GW_SCAN * 32 + c];
                const double ce = reinterpret_cast<const double *>(gw + GW_LCE * 32)[c], cl = reinterpret_cast<const double *>(gw + GW_LCL * 32)[c];
                if (fl & 4) {
                    // The six 32-feature groups have a fixed structure (SVMipv4.cpp:72-112, checked against the
                    // descriptor table at context creation):
                    //   m=0: ext ratios 0..20 | ext_len 21 | lrc 22..31      m=1: lrc 32..63
                    //   m=2: lrc 64,65 | insert ratios 66..95                  m=3: insert ratios 96..127
                    //   m=4: insert ratios 128..150 | scan_size 151 | lig ratios 152..159
                    //   m=5: lig ratios 160..172 | lig_len 173 | junction one-hot 174..189 | log copies 190,191
                    auto ratio = [&](int m, int a, int n, int len) {
                        const int km1 = (fd[m] >> 5) & 3;
                        const int ro = rc ? rowoff_r[m] : rowoff_f[m];
                        const int end = a + n - km1;
                        const int cnt = (int)(uint16_t)(P[ro + max(end, a)] - P[ro + a]);
                        const int den = len - km1;
                        const double y = rn[den], ad = (double)cnt, bd = (double)den;
                        const double q0 = __dmul_rn(ad, y);
                        return __fma_rn(__fma_rn(-q0, bd, ad), y, q0);
                    };
                    double v0 = ratio(0, ext_a, ext_n, ext_len);
                    v0 = lane < 21 ? v0 : (lane == 21 ? (double)ext_len : lrc_r0);
                    double v2 = ratio(2, tgt_a, tgt_n, scan_size);
                    v2 = lane < 2 ? lrc_r2 : v2;
                    const double v3 = ratio(3, tgt_a, tgt_n, scan_size);
                    const bool lig4 = lane >= 24;
                    double v4 = ratio(4, lig4 ? lig_a : tgt_a, lig4 ? lig_n : tgt_n, lig4 ? lig_len : scan_size);
                    v4 = lane == 23 ? (double)scan_size : v4;
                    double v5 = ratio(5, lig_a, lig_n, lig_len);
                    v5 = lane < 13 ? v5 : (lane == 13 ? (double)lig_len : (lane < 30 ? (jc == lane - 14 ? 1.0 : 0.0) : (lane == 30 ? ce : cl)));
                    xrow[0] = v0; xrow[32] = lrc_r1; xrow[64] = v2; xrow[96] = v3; xrow[128] = v4; xrow[160] = v5;
                } else {
                    // unusual lengths (divisor <= 0 or beyond the reciprocal table): plain IEEE division
#pragma unroll
                    for (int m = 0; m < 6; m++) {
                        const uint32_t d = fd[m];
                        const uint32_t kind = d & 7, part = (d >> 3) & 3, km1 = (d >> 5) & 3, jj = (d >> 23) & 255;
                        const int len = part == 0 ? ext_len : (part == 1 ? scan_size : lig_len);
                        double val;
                        if (kind == FK_RATIO) {
                            const int a = part == 0 ? ext_a : (part == 1 ? tgt_a : lig_a);
                            const int n = part == 0 ? ext_n : (part == 1 ? tgt_n : lig_n);
                            const int row = rc ? (int)((d >> 15) & 255) : (int)((d >> 7) & 255);
                            val = __ddiv_rn((double)pf_count(w, row, a, n, (int)km1 + 1), (double)(len - (int)km1));
                        } else if (kind == FK_LEN) val = (double)len;
                        else if (kind == FK_LRC) val = lrc_s[jj];
                        else if (kind == FK_JUNC) val = (jc == (int)jj) ? 1.0 : 0.0;
                        else val = jj == 0 ? ce : cl;
                        xrow[32 * m] = val;
                    }
                }
            }
        }

        // ---- row-table mode: the distinct arm / insert rows of the factored-SVR work items nested in this task ----
        // A work item (fc->W scan starts, one capture size, one strand; k_svr_fact.cu) needs, instead of its candidates' 192-vectors,
        //   rows [0, RA):        the arm that ends at the scan start      (+: extension [s-e, s-1], -: ligation [s-l, s-1])
        //   rows [RA, RA+RQ):    the arm that starts at q = s + cap - sum (+: ligation, -: extension), q-indexed
        //   rows [RA+RQ, R):     the insert [s, s + cap - sum - 1]
        // each as its block of the feature vector (21 ratios, length, log copy | 85 ratios, scan size) with -gamma ||row||^2
        // in the block's spare column -- exactly the tables k_svr_fact keeps in shared memory, so it fetches them with one bulk copy.
        // One lane per row (32 rows per warp pass): the row's geometry is decoded once, its ratios come out of the prefix tables
        // in a loop over the features, and the norm accumulates in the lane.
        if (rows) {
            const int fW = fc->W, n_ext = fc->n_ext, n_lig = fc->n_lig, n_sums = fc->n_sums, dsum = fc->max_sum - fc->min_sum;
            const int n_pair = ((tk.nsi + fW - 1) / fW) * tk.nci;   // (sub-window, capture size) pairs of this task
            const double kZeroNorm = -gamma * 0.0, kNegInf = __longlong_as_double(0xfff0000000000000LL);
            auto r16 = [](int v) { return (v + 15) & ~15; };
            // row counts of a full sub-window, per strand (a shorter last sub-window uses a prefix of each range)
            const int capA0 = r16(fW * n_ext) + r16((fW + dsum) * n_lig), capA1 = r16(fW * n_lig) + r16((fW + dsum) * n_ext);
            const int capI = r16(fW * n_sums);
            struct Item { int ok, strand, ci, cap, nsi_f, s0; double *base; };
            auto item_of = [&](int pair, int strand) {
                Item it;
                const int sub = pair / tk.nci, cr = pair - sub * tk.nci;
                it.strand = strand; it.ci = tk.ci0 + cr; it.cap = wc.max_capture - it.ci * wc.inc;
                // a capture size ruled out for the whole region (mipgen.cpp:429) scores nothing: k_svr_fact never fetches its tables
                it.ok = !(it.cap > r.stop_flanked - r.start_flanked + wc.max_mip_overlap && it.cap - wc.inc >= wc.min_capture);
                it.nsi_f = min(fW, tk.nsi - sub * fW);
                it.s0 = r.first_scan + tk.si0 + sub * fW;  // the work item's first scan start (chromosome coordinate)
                it.base = rows + (int64_t)(tk.ft0 + (sub * wc.n_cap + it.ci) * 2 + strand - ftask_base) * fc->blob_doubles;
                return it;
            };
            // the characters present: std::string::substr clamps the length (as candidate_geometry does); rows of scored
            // candidates always lie inside the window's span, anything else is a dead row (zeros)
            auto window_of = [&](bool live, int start, int len, int &a, int &n) {
                const int off = start - r.seq_start;
                live = live && len >= 0 && off >= 0 && off <= r.seq_len;
                n = live ? min(len, r.seq_len - off) : 0;
                a = off - span0;
                return live && a >= 0 && a + n <= span_len;
            };
            auto ratio_of = [&](int tab_off, int km1, int a, int n, int len) {
                const int end = a + n - km1;
                const int cnt = (int)(uint16_t)(P[tab_off + max(end, a)] - P[tab_off + a]);
                const int den = len - km1;
                const double ad = (double)cnt, bd = (double)den;
                if (den <= 0 || den >= kRecipN) return __ddiv_rn(ad, bd);
                const double y = rn[den], q0 = __dmul_rn(ad, y);
                return __fma_rn(__fma_rn(-q0, bd, ad), y, q0);
            };
            // ---- arm rows ----
            const int n_arm = n_pair * (capA0 + capA1);
            for (int i = warp * 32 + lane; i < n_arm; i += kWarpsPerBlock * 32) {
                const int pair = i / (capA0 + capA1), rem = i - pair * (capA0 + capA1);
                const int strand = rem >= capA0, row = rem - strand * capA0;
                const Item it = item_of(pair, strand);
                const int nA = strand ? n_lig : n_ext, nQ = strand ? n_ext : n_lig;
                const int RA = r16(it.nsi_f * nA), RQ = r16((it.nsi_f + dsum) * nQ);
                if (!it.ok || row >= RA + RQ) continue;
                int role, len, start;   // role 0 extension arm, 1 ligation arm
                bool live;
                double *dst;
                if (row < RA) {
                    const int s_rel = row / nA, ia = row - s_rel * nA;
                    live = s_rel < it.nsi_f; role = strand ? 1 : 0;
                    len = strand ? fc->lig_of[ia] : fc->ext_of[ia];
                    start = it.s0 + s_rel - len;
                    dst = it.base + row * FACT_LD_ARM;
                } else {
                    const int m = row - RA, j = m / nQ, iq = m - j * nQ;
                    live = j < it.nsi_f + dsum; role = strand ? 0 : 1;
                    len = strand ? fc->ext_of[iq] : fc->lig_of[iq];
                    start = it.s0 + j + it.cap - fc->max_sum;
                    dst = it.base + fc->cap_FA + m * FACT_LD_ARM;
                }
                int a, n;
                live = window_of(live, start, len, a, n);
                double ssum = 0.0;
                int code = 16;
                if (live) {
                    const int *tab = rt_tab + (role * 2 + strand) * 21, *km = rt_tab + 84 + role * 21;
#pragma unroll 3
                    for (int k = 0; k < 21; k++) {
                        const double v = ratio_of(tab[k], km[k], a, n, len);
                        ssum = fma(v, v, ssum);
                        dst[k] = v;
                    }
                    const double vl = (double)len, vc = r.copy_off >= 0 ? log_copy(copy_lookup(cfg, r, copies, start, len), logtab) : 0.0;
                    ssum = fma(vl, vl, ssum);
                    ssum = fma(vc, vc, ssum);
                    dst[21] = vl; dst[22] = vc;
                    if (role == 1 && n >= 2) {
                        const uint32_t j0 = strand ? codes_s[a + n - 1] : codes_s[a], j1 = strand ? codes_s[a + n - 2] : codes_s[a + 1];
                        if (j0 < 4 && j1 < 4) code = strand ? (int)((3 - j0) * 4 + (3 - j1)) : (int)(j0 * 4 + j1);
                    }
                }
                // a non-finite feature (log10(0) = -inf copy, a zero divisor) makes every kernel value of the row 0, as in libsvm:
                // the row is parked at exponent -inf with finite (zero) features so the contraction stays NaN free
                const bool finite = fabs(ssum) <= 1.7976931348623157e308;
                if (!live || !finite)
                    for (int k = 0; k < FACT_K_ARM - 1; k++) dst[k] = 0.0;
                const double nrm = !live ? kZeroNorm : (finite ? -gamma * ssum : kNegInf);
                dst[FACT_K_ARM - 1] = nrm;
                double *xxg = it.base + fc->cap_FA + fc->cap_FQ + fc->cap_FI;
                xxg[row] = nrm;
                reinterpret_cast<int *>(xxg + fc->cap_R)[row] = role == 1 ? code : 16;
            }
            // ---- insert rows ----
            const int n_ins = n_pair * 2 * capI;
            for (int i = warp * 32 + lane; i < n_ins; i += kWarpsPerBlock * 32) {
                const int pair = i / (2 * capI), rem = i - pair * (2 * capI);
                const int strand = rem >= capI, m = rem - strand * capI;
                const Item it = item_of(pair, strand);
                if (!it.ok || m >= r16(it.nsi_f * n_sums)) continue;
                const int nA = strand ? n_lig : n_ext, nQ = strand ? n_ext : n_lig;
                const int row = r16(it.nsi_f * nA) + r16((it.nsi_f + dsum) * nQ) + m;
                const int s_rel = m / n_sums, is = m - s_rel * n_sums;
                const int len = it.cap - fc->sum_of[is];  // scan size
                double *dst = it.base + fc->cap_FA + fc->cap_FQ + m * FACT_LD_INS;
                int a, n;
                const bool live = window_of(s_rel < it.nsi_f, it.s0 + s_rel, len, a, n);
                double ssum = 0.0;
                if (live) {
                    const int *tab = rt_tab + 126 + strand * 85, *km = rt_tab + 296;
#pragma unroll 5
                    for (int k = 0; k < 85; k++) {
                        const double v = ratio_of(tab[k], km[k], a, n, len);
                        ssum = fma(v, v, ssum);
                        dst[k] = v;
                    }
                    const double vl = (double)len;
                    ssum = fma(vl, vl, ssum);
                    dst[85] = vl;
                }
                const bool finite = fabs(ssum) <= 1.7976931348623157e308;
                if (!live || !finite)
                    for (int k = 0; k < FACT_K_INS - 2; k++) dst[k] = 0.0;
                const double nrm = !live ? kZeroNorm : (finite ? -gamma * ssum : kNegInf);
                dst[FACT_K_INS - 2] = nrm;
                dst[FACT_K_INS - 1] = 0.0;
                double *xxg = it.base + fc->cap_FA + fc->cap_FQ + fc->cap_FI;
                xxg[row] = nrm;
                reinterpret_cast<int *>(xxg + fc->cap_R)[row] = 16;
            }
        }
    }
}

// ------------------------------ explicit front-end ------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_feat_explicit(const DevCand *__restrict__ cands, const uint8_t *__restrict__ codes, const double *__restrict__ lrc,
                const uint32_t *__restrict__ fdesc, const double *__restrict__ logtab, int64_t n,
                double *__restrict__ logistic, double *__restrict__ x)
{
    __shared__ int smem[kWarpsPerBlock * kWarpSmemInts];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *sm = smem + warp * kWarpSmemInts;
    uint32_t fd[6];
#pragma unroll
    for (int m = 0; m < 6; m++) fd[m] = fdesc[lane + 32 * m];
    const int64_t n_blocks = (n + 31) >> 5;
    for (int64_t blk = (int64_t)blockIdx.x * kWarpsPerBlock + warp; blk < n_blocks; blk += (int64_t)gridDim.x * kWarpsPerBlock) {
        Stash st;
        st.state = 0;
        for (int c = 0; c < 32; c++) {
            const int64_t i = (blk << 5) + c;
            if (i >= n) break;
            const DevCand dc = cands[i];
            View v;
            v.ext = codes + dc.ext_off; v.lig = codes + dc.lig_off; v.tgt = codes + dc.tgt_off;
            v.ext_n = dc.ext_n; v.lig_n = dc.lig_n; v.tgt_n = dc.tgt_n;
            v.ext_len = dc.ext_len; v.lig_len = dc.lig_len; v.scan_size = dc.scan_size;
            v.rc = 0; v.ext_copy = dc.ext_copy; v.lig_copy = dc.lig_copy;
            v.lrc = lrc ? lrc + i * MG_NLRC : nullptr;
            v.ok = 1;
            warp_candidate(v, sm, lane, fd, logtab, x ? x + i * MG_NFEAT : nullptr, st, lane == c);
        }
        const int64_t i = (blk << 5) + lane;
        if (i < n && logistic) logistic[i] = logistic_score(st, logtab);
    }
}

// ---------------------------------- K-encode ----------------------------------
__global__ void k_encode(const char *__restrict__ in, uint8_t *__restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        char ch = in[i];
        uint8_t c = B_OTHER;
        switch (ch) {
        case 'A': c = B_A; break;
        case 'C': c = B_C; break;
        case 'G': c = B_G; break;
        case 'T': c = B_T; break;
        case 'N': c = B_N; break;
        case '-': c = B_DASH; break;
        default: break;
        }
        out[i] = c;
    }
}

// ----------------------------------- K-lrc ------------------------------------
// Featurev5::get_long_range_content (Featurev5.cpp:18-56): for each of the 44 k-mers of
// mipgen.cpp:32, (count(mer) + count(revcomp) unless palindromic) / denom.
__constant__ uint8_t c_lrc_k[MG_NLRC];
__constant__ uint8_t c_lrc_code[MG_NLRC];

__global__ void __launch_bounds__(256) k_lrc(const uint8_t *__restrict__ codes, int n, int denom, double *__restrict__ out)
{
    __shared__ int h[84];  // tri[64] di[16] mono[4]
    if (threadIdx.x < 84) h[threadIdx.x] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        uint32_t c0 = codes[p];
        if (c0 >= 4) continue;
        atomicAdd(&h[80 + c0], 1);
        uint32_t c1 = p + 1 < n ? codes[p + 1] : B_NONE;
        if (c1 >= 4) continue;
        atomicAdd(&h[64 + c0 * 4 + c1], 1);
        uint32_t c2 = p + 2 < n ? codes[p + 2] : B_NONE;
        if (c2 >= 4) continue;
        atomicAdd(&h[c0 * 16 + c1 * 4 + c2], 1);
    }
    __syncthreads();
    if (threadIdx.x < MG_NLRC) {
        int k = c_lrc_k[threadIdx.x], code = c_lrc_code[threadIdx.x];
        int rc = 0;
        for (int i = 0, q = code; i < k; i++, q >>= 2) rc = (rc << 2) | (3 - (q & 3));
        int base = k == 3 ? 0 : (k == 2 ? 64 : 80);
        double f = (double)h[base + code];
        double v = (rc != code) ? __dadd_rn(f, (double)h[base + rc]) : f;
        out[threadIdx.x] = __ddiv_rn(v, (double)denom);
    }
}

}  // namespace

// ------------------------------------ launchers ------------------------------------
static int feat_grid_dim(mg_ctx *ctx, int64_t n)
{
    int64_t blocks = (n + 32 * kWarpsPerBlock - 1) / (32 * kWarpsPerBlock);
    int64_t cap = (int64_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int launch_encode(mg_ctx *ctx, const char *d_ascii, uint8_t *d_codes, int64_t n)
{
    if (n <= 0) return MG_OK;
    int blocks = (int)((n + 255) / 256 < ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
    mg_time_begin(ctx, TM_OTHER, n);
    k_encode<<<blocks, 256, 0, ctx->stream>>>(d_ascii, d_codes, n);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

int mg_upload_lrc_tables(mg_ctx *ctx, const uint8_t *k, const uint8_t *code)
{
    CUDA_TRY(ctx, cudaMemcpyToSymbol(c_lrc_k, k, MG_NLRC));
    CUDA_TRY(ctx, cudaMemcpyToSymbol(c_lrc_code, code, MG_NLRC));
    return MG_OK;
}

int launch_lrc(mg_ctx *ctx, const uint8_t *d_codes, int n, int denom, double *d_out44)
{
    mg_time_begin(ctx, TM_OTHER, n);
    k_lrc<<<1, 256, 0, ctx->stream>>>(d_codes, n, denom, d_out44);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

size_t feat_window_smem(int stride, int span_cap, int n_pairs)
{
    return (size_t)kRecipN * 8 + MG_NLRC * 8 + (size_t)kWarpsPerBlock * GW_FIELDS * 32 * 4 + (size_t)(2 * n_pairs + 2) * 4 +
           (size_t)PF_ROWS * stride * 2 + (size_t)span_cap + 8;
}

int launch_feat_setup(mg_ctx *ctx)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_feat_window, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return MG_OK;
}

int launch_feat_grid(mg_ctx *ctx, const mg_panel *p, int task0, int task1, int64_t g_base, int64_t n_cand, uint8_t *d_valid,
                     double *d_logistic, double *d_x, uint8_t *d_state, double *d_rows, int ftask_base)
{
    if (task1 <= task0) return MG_OK;
    const size_t smem = feat_window_smem(p->pf_stride, p->span_cap, (int)ctx->cfg.ext_len.size());
    if (smem > 200 * 1024) { ctx->err = "capture size too large for the K-feat window tables"; return MG_ERR_INVALID; }
    int blocks = task1 - task0;
    const int cap = ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    mg_time_begin(ctx, TM_FEAT, n_cand);
    k_feat_window<<<blocks, kWarpsPerBlock * 32, smem, ctx->stream>>>(ctx->d_cfg, p->d_regions, p->d_tasks, task0, task1, p->d_codes,
                                                                      p->d_lrc, p->d_copies, ctx->d_fdesc_win, ctx->d_logcopy, g_base,
                                                                      d_valid, d_logistic, d_x, p->pf_stride, d_state, d_rows, ctx->d_fact,
                                                                      ftask_base, ctx->gamma);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}

int launch_feat_explicit(mg_ctx *ctx, const DevCand *d_cands, const uint8_t *d_codes, const double *d_lrc, int64_t n,
                         double *d_logistic, double *d_x)
{
    if (n <= 0) return MG_OK;
    mg_time_begin(ctx, TM_FEAT, n);
    k_feat_explicit<<<feat_grid_dim(ctx, n), kWarpsPerBlock * 32, 0, ctx->stream>>>(d_cands, d_codes, d_lrc, ctx->d_fdesc,
                                                                                   ctx->d_logcopy, n, d_logistic, d_x);
    mg_time_end(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return MG_OK;
}
