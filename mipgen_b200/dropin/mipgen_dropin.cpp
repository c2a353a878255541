// mipgen_dropin.cpp -- host side of the drop-in boundary: the reference's scoring classes
// (Featurev5, SVMipv4, PlusSVMipv4, MinusSVMipv4) and the libsvm entry points mipgen.cpp
// calls, implemented on top of the C-ABI of libmipgen_b200.so.  With these headers on the
// include path the UNCHANGED /root/reference/mipgen.cpp compiles and links, and every score
// it consumes is produced by the CUDA kernels.  There is no scoring arithmetic in this file:
// a failure to reach the GPU is fatal (message on stderr, exit code 70), never a fallback.
//
// How a per-object API is served by a batched device path
// --------------------------------------------------------
// mipgen.cpp asks for one candidate at a time (tile_regions 466-486) and passes no region
// to the scoring classes.  The shim therefore
//   1. keeps a registry of live Featurev5 objects (their ctor/dtor below) and finds the
//      region a candidate belongs to through chr + current_scan_start_position
//      (mipgen.cpp:425 keeps it equal to the candidate's scan start while tiling);
//   2. learns the candidate pattern -- the set of (capture, ext, lig) combinations the
//      caller enumerates per scan start -- from the calls themselves, so it needs none of
//      mipgen's private configuration;
//   3. on a miss, scores in ONE device call every remaining scan start of the region for
//      the whole pattern (a superset of what the score-dependent pruning will ask for,
//      SURVEY.md F4), and answers the following ~1e5 calls from that grid;
//   4. takes the arm copy numbers of the region from the file mipgen's find_copy wrote just before
//      (<project>.oligo_copy_count.sam, mipgen.cpp:558-596; project name from the process's own command
//      line), since the caller's copy map is private (mipgen.cpp:83) and reaches the scoring classes only
//      one object at a time (mipgen.cpp:612-613);
//   5. verifies each answer's geometry, sequences and copy numbers against the object before using it
//      and re-scores the single candidate explicitly (mg_score_candidates) whenever anything differs.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <set>
#include <unordered_map>

#include "Featurev5.h"
#include "SVMipv4.h"
#include "PlusSVMipv4.h"
#include "MinusSVMipv4.h"
#include "svm.h"
#include "../../include/mipgen_b200.h"

namespace mgshim {

[[noreturn]] static void fatal(const char *what, mg_ctx *ctx)
{
    fprintf(stderr, "[mipgen_b200] fatal: %s: %s\n", what, mg_last_error(ctx));
    fprintf(stderr, "[mipgen_b200] scoring runs on the GPU only; there is no CPU fallback\n");
    exit(70);
}

static mg_ctx *g_ctx = nullptr;
static bool g_have_model = false;
static bool g_has_pending_svr = false;  // batched driver: get_parameters -> predict_value hand-over
static double g_pending_svr = 0;

static mg_ctx *ctx()
{
    if (!g_ctx) {
        // MIPGEN_B200_DEVICE=<ordinal>, or the first entry of MIPGEN_B200_DEVICES (batched driver: "0,1,2" / "0-7")
        const char *dev = getenv("MIPGEN_B200_DEVICE");
        if (!dev || !*dev) dev = getenv("MIPGEN_B200_DEVICES");
        if (mg_create(dev ? atoi(dev) : 0, &g_ctx) != MG_OK) fatal("mg_create", nullptr);
    }
    return g_ctx;
}

static std::set<Featurev5 *> &registry()
{
    static std::set<Featurev5 *> r;
    return r;
}

struct Combo { int capture, ext, lig; };

// Arm copy numbers.  mipgen keeps them in a private map (copy_chr_start_stop, mipgen.cpp:83) and hands them to the scoring
// classes one object at a time (mipgen.cpp:612-613).  A region batch needs them up front, so the shim reads the file the map
// was filled from: find_copy (mipgen.cpp:558-596) has already written <project>.oligo_copy_count.sam when tile_regions starts;
// the project name is taken from this process's own command line.  Every grid look-up is still checked against the copies
// the object carries (same_sequences), so a missing or different file can only cost speed, never correctness.
struct CopyMap {
    bool tried = false, loaded = false;
    std::unordered_map<std::string, std::unordered_map<uint64_t, int>> by_chr;
    static uint64_t key(int start, int stop) { return ((uint64_t)(uint32_t)start << 32) | (uint32_t)stop; }

    void load()
    {
        tried = true;
        std::ifstream cl("/proc/self/cmdline", std::ios::binary);
        std::string all((std::istreambuf_iterator<char>(cl)), std::istreambuf_iterator<char>()), project;
        for (size_t at = 0; at < all.size();) {
            const std::string arg(all.c_str() + at);
            at += arg.size() + 1;
            if (arg == "-project_name" && at < all.size()) project = std::string(all.c_str() + at);
        }
        if (project.empty()) return;
        std::ifstream sam((project + ".oligo_copy_count.sam").c_str());
        if (!sam) return;
        std::string line;
        while (std::getline(sam, line)) {
            if (line.size() <= 1 || line[0] == '@') continue;
            // the reference's own parse (mipgen.cpp:572-590): "chr<chr>:<start>-<stop>\t...", copy = X0:i:<n>, 100 without the tag
            const size_t c0 = line.find("chr", 0);
            if (c0 == std::string::npos) continue;
            const size_t c1 = line.find(':', c0 + 3), d = line.find('-', c1 + 1), t = line.find('\t', d + 1);
            if (c1 == std::string::npos || d == std::string::npos || t == std::string::npos) continue;
            const int start = atoi(line.substr(c1 + 1, d - c1 - 1).c_str()), stop = atoi(line.substr(d + 1, t - d - 1).c_str());
            const size_t x0 = line.find("X0:i:", 0);
            by_chr[line.substr(c0 + 3, c1 - c0 - 3)][key(start, stop)] = x0 != std::string::npos ? atoi(line.c_str() + x0 + 5) : 100;
        }
        loaded = true;
    }

    // copy_chr_start_stop[chr][start][stop]; an absent key reads as 0 there (operator[])
    int get(const std::string &chr, int start, int stop) const
    {
        auto c = by_chr.find(chr);
        if (c == by_chr.end()) return 0;
        auto it = c->second.find(key(start, stop));
        return it == c->second.end() ? 0 : it->second;
    }
};

static CopyMap &copy_map()
{
    static CopyMap m;
    if (!m.tried) m.load();
    return m;
}

// One device batch: scan starts [s0, s1] x the first n_pairs/n_caps of the pattern.
struct Batch {
    const Featurev5 *feature = nullptr;
    std::string chr;
    int s0 = 0, s1 = -1;
    int max_cap = 0, min_cap = 0, inc = 1, n_cap = 0;
    std::vector<int> ext, lig;  // pair table used for this batch
    bool has_logistic = false, has_svr = false;
    std::vector<uint8_t> valid;
    std::vector<double> logistic, svr, feats;
    std::vector<int> oligo_sizes, copies;  // [n_oligo][seq_len] arm copy table the batch was scored with; empty => every copy 1
    int seq_start = 0;
    std::string seq;  // copy of the region sequence the batch was computed from
    double lrc[MG_NLRC];

    long index(int s, int capture, int e, int l, int strand) const
    {
        if (s < s0 || s > s1 || capture > max_cap || capture < min_cap || (max_cap - capture) % inc) return -1;
        int p = -1;
        for (size_t i = 0; i < ext.size(); i++)
            if (ext[i] == e && lig[i] == l) { p = (int)i; break; }
        if (p < 0) return -1;
        int ci = (max_cap - capture) / inc;
        return ((((long)(s - s0)) * n_cap + ci) * (long)ext.size() + p) * 2 + strand;
    }

    // the copy number the batch assumed for the arm [start, start + len)
    int copy_of(int start, int len) const
    {
        if (copies.empty()) return 1;
        for (size_t k = 0; k < oligo_sizes.size(); k++)
            if (oligo_sizes[k] == len) {
                const int i = start - seq_start;
                return (i < 0 || i >= (int)seq.size()) ? 0 : copies[k * seq.size() + (size_t)i];
            }
        return 0;
    }
};

struct Engine {
    std::vector<Combo> pattern;  // every (capture, ext, lig) the caller has asked for so far
    Batch batch;                 // the current region grid
    // the candidate get_parameters() last answered: svm_predict() is always called right
    // after it on the same object (mipgen.cpp:471-472, 485-486, 1525-1526, ...)
    bool last_valid = false;
    double last_feats[MG_NFEAT];
    double last_svr = 0;
    bool last_has_svr = false;
    long n_batches = 0, n_explicit = 0, n_hits = 0;
};

static Engine &engine()
{
    static Engine e;
    return e;
}

static void report_at_exit()
{
    if (getenv("MIPGEN_B200_VERBOSE")) {
        Engine &e = engine();
        fprintf(stderr, "[mipgen_b200] device batches %ld, explicit candidates %ld, grid look-ups %ld\n", e.n_batches, e.n_explicit,
                e.n_hits);
    }
}

static bool pattern_add(const Combo &c)
{
    Engine &e = engine();
    for (auto &p : e.pattern)
        if (p.capture == c.capture && p.ext == c.ext && p.lig == c.lig) return false;
    e.pattern.push_back(c);
    return true;
}

static int gcd(int a, int b) { return b == 0 ? a : gcd(b, a % b); }

static const Featurev5 *find_feature(const SVMipv4 *m)
{
    const Featurev5 *fallback = nullptr;
    int lo = std::min(m->ext_probe_start, m->lig_probe_start), hi = std::max(m->ext_probe_stop, m->lig_probe_stop);
    for (Featurev5 *f : registry()) {
        if (f->chr != m->chr || f->chromosomal_sequence.empty()) continue;
        if (lo < f->chromosomal_sequence_start_position || hi > f->chromosomal_sequence_stop_position) continue;
        if ((int)f->chromosomal_sequence.size() != f->chromosomal_sequence_stop_position - f->chromosomal_sequence_start_position + 1) continue;
        if (f->current_scan_start_position == m->scan_start_position) return f;  // the region being tiled right now
        if (!fallback) fallback = f;
    }
    return fallback;
}

// Does the grid entry describe exactly this object?  (geometry is implied by the index;
// here: the sequences the object carries are the ones the batch was computed from.)
static bool same_sequences(const Batch &b, const SVMipv4 *m)
{
    // the copies the object carries are the ones the grid point was scored with
    if (m->ext_probe_copy != b.copy_of(m->ext_probe_start, m->extension_arm_length) ||
        m->lig_probe_copy != b.copy_of(m->lig_probe_start, m->ligation_arm_length))
        return false;
    if ((int)m->ext_probe_sequence.size() != m->extension_arm_length || (int)m->lig_probe_sequence.size() != m->ligation_arm_length ||
        (int)m->scan_target_sequence.size() != m->scan_size)
        return false;
    const bool minus = m->strand == "-";
    auto eq = [&](const std::string &s, int start) {
        int off = start - b.seq_start, n = (int)s.size();
        if (off < 0 || off + n > (int)b.seq.size()) return false;
        if (!minus) return memcmp(s.data(), b.seq.data() + off, n) == 0;
        for (int i = 0; i < n; i++) {
            char g = b.seq[off + n - 1 - i], c = g;
            if (g == 'A') c = 'T'; else if (g == 'C') c = 'G'; else if (g == 'G') c = 'C'; else if (g == 'T') c = 'A';
            if (s[i] != c) return false;
        }
        return true;
    };
    return eq(m->ext_probe_sequence, m->ext_probe_start) && eq(m->lig_probe_sequence, m->lig_probe_start) &&
           eq(m->scan_target_sequence, m->scan_start_position);
}

// Score scan starts [s0, s1] of feature f for the whole known pattern in one device call.
static void run_batch(const Featurev5 *f, int s0, int s1, bool want_svr, const double *lrc)
{
    Engine &e = engine();
    Batch &b = e.batch;
    b = Batch();
    b.feature = f;
    b.chr = f->chr;
    b.s0 = s0;
    b.s1 = s1;
    b.max_cap = 0;
    b.min_cap = 1 << 30;
    for (auto &c : e.pattern) { b.max_cap = std::max(b.max_cap, c.capture); b.min_cap = std::min(b.min_cap, c.capture); }
    int g = 0;
    for (auto &c : e.pattern) g = gcd(g, b.max_cap - c.capture);
    b.inc = g == 0 ? 1 : g;
    b.n_cap = (b.max_cap - b.min_cap) / b.inc + 1;
    // pair table: arm sum descending, first-seen order within a sum (the caller's own order)
    std::vector<std::pair<int, int>> pairs;
    for (auto &c : e.pattern) {
        std::pair<int, int> p(c.ext, c.lig);
        if (std::find(pairs.begin(), pairs.end(), p) == pairs.end()) pairs.push_back(p);
    }
    std::stable_sort(pairs.begin(), pairs.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &c) { return a.first + a.second > c.first + c.second; });
    for (auto &p : pairs) { b.ext.push_back(p.first); b.lig.push_back(p.second); }

    mg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.max_capture = b.max_cap;
    cfg.min_capture = b.min_cap;
    cfg.capture_increment = b.inc;
    cfg.max_mip_overlap = 1 << 29;  // the static skip of mipgen.cpp:429 is the caller's business here
    cfg.n_pairs = (int)b.ext.size();
    cfg.ext_len = b.ext.data();
    cfg.lig_len = b.lig.data();
    std::vector<int> sizes;  // oligo sizes of this pattern (the copy table below is built for them)
    if (copy_map().loaded) {
        for (auto &p : pairs) { sizes.push_back(p.first); sizes.push_back(p.second); }
        std::sort(sizes.begin(), sizes.end());
        sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
        cfg.n_oligo_sizes = (int)sizes.size();
        cfg.oligo_sizes = sizes.data();
    }
    if (mg_set_config(ctx(), &cfg) != MG_OK) fatal("mg_set_config", ctx());

    b.seq = f->chromosomal_sequence;
    b.seq_start = f->chromosomal_sequence_start_position;
    memcpy(b.lrc, lrc ? lrc : f->long_range_content, sizeof b.lrc);
    // arm copy table of this region from the file find_copy wrote, for every arm length of the pattern
    const CopyMap &cm = copy_map();
    if (cm.loaded) {
        for (auto &p : pairs) { b.oligo_sizes.push_back(p.first); b.oligo_sizes.push_back(p.second); }
        std::sort(b.oligo_sizes.begin(), b.oligo_sizes.end());
        b.oligo_sizes.erase(std::unique(b.oligo_sizes.begin(), b.oligo_sizes.end()), b.oligo_sizes.end());
        b.copies.assign(b.oligo_sizes.size() * b.seq.size(), 0);
        for (size_t k = 0; k < b.oligo_sizes.size(); k++)
            for (size_t i = 0; i < b.seq.size(); i++)
                b.copies[k * b.seq.size() + i] = cm.get(f->chr, b.seq_start + (int)i, b.seq_start + (int)i + b.oligo_sizes[k] - 1);
    }
    mg_region r;
    memset(&r, 0, sizeof r);
    r.seq = b.seq.data();
    r.seq_len = (int)b.seq.size();
    r.seq_start = b.seq_start;
    r.seq_stop = f->chromosomal_sequence_stop_position;
    r.start_flanked = f->start_position_flanked;
    r.stop_flanked = f->stop_position_flanked;
    r.scan_begin = s0;  // explicit scan range: the rest of the region from the caller's position
    r.scan_end = s1;
    r.lrc = b.lrc;
    r.copies = b.copies.empty() ? nullptr : b.copies.data();
    int64_t n = mg_grid_size(ctx(), &r);
    b.valid.resize((size_t)n);
    b.logistic.resize((size_t)n);
    int want = MG_WANT_LOGISTIC;
    if (want_svr) {
        want |= MG_WANT_SVR | MG_WANT_FEATURES;
        b.svr.resize((size_t)n);
        b.feats.resize((size_t)n * MG_NFEAT);
    }
    if (mg_score_regions(ctx(), &r, 1, want, nullptr, b.valid.data(), b.logistic.data(), want_svr ? b.svr.data() : nullptr,
                         want_svr ? b.feats.data() : nullptr) != MG_OK)
        fatal("mg_score_regions", ctx());
    b.has_logistic = true;
    b.has_svr = want_svr;
    e.n_batches++;
}

// Explicit single-candidate path: the object's own strings and copies.
static void score_explicit(const SVMipv4 *m, const double *lrc, double *logistic, double *svr, double *feats)
{
    mg_candidate c;
    c.ext = m->ext_probe_sequence.data(); c.ext_n = (int)m->ext_probe_sequence.size();
    c.lig = m->lig_probe_sequence.data(); c.lig_n = (int)m->lig_probe_sequence.size();
    c.tgt = m->scan_target_sequence.data(); c.tgt_n = (int)m->scan_target_sequence.size();
    c.ext_len = m->extension_arm_length; c.lig_len = m->ligation_arm_length; c.scan_size = m->scan_size;
    c.ext_copy = m->ext_probe_copy; c.lig_copy = m->lig_probe_copy;
    int want = (logistic ? MG_WANT_LOGISTIC : 0) | (svr ? MG_WANT_SVR : 0) | (feats ? MG_WANT_FEATURES : 0);
    if (mg_score_candidates(ctx(), &c, 1, lrc, want, logistic, svr, feats) != MG_OK) fatal("mg_score_candidates", ctx());
    engine().n_explicit++;
}

// memory bound for one batch's host copy of the feature rows
static const long kMaxFeatureBytes = 768L << 20;

// Returns the grid index of m in the current batch, launching a device batch if needed;
// -1 when the candidate cannot be served from a grid (explicit path).
static long locate(const SVMipv4 *m, bool need_svr, const double *lrc)
{
    Engine &e = engine();
    static bool hooked = false;
    if (!hooked) { atexit(report_at_exit); hooked = true; }
    // without the copy file the grids assume copy 1 for every arm: anything else goes the explicit way
    if (!copy_map().loaded && (m->ext_probe_copy != 1 || m->lig_probe_copy != 1)) return -1;
    if (m->strand != "+" && m->strand != "-") return -1;
    const int strand = m->strand == "-";
    const int capture = m->scan_size + m->extension_arm_length + m->ligation_arm_length;
    // the grid derives the arms from (scan start, capture, ext, lig, strand): the object must agree
    const int t = m->scan_stop_position, s = m->scan_start_position;
    const int ext_start = strand ? t + 1 : s - m->extension_arm_length, lig_start = strand ? s - m->ligation_arm_length : t + 1;
    if (m->scan_size != t - s + 1 || m->ext_probe_start != ext_start || m->lig_probe_start != lig_start) return -1;

    Batch &b = e.batch;
    if (b.feature && b.chr == m->chr && (!need_svr || b.has_svr)) {
        long idx = b.index(s, capture, m->extension_arm_length, m->ligation_arm_length, strand);
        if (idx >= 0 && b.valid[(size_t)idx] && (!lrc || !need_svr || memcmp(lrc, b.lrc, sizeof b.lrc) == 0) && same_sequences(b, m)) {
            e.n_hits++;
            return idx;
        }
    }
    const Featurev5 *f = find_feature(m);
    if (!f) return -1;
    Combo c = {capture, m->extension_arm_length, m->ligation_arm_length};
    const bool grew = pattern_add(c);
    // While the pattern is still growing (first scan start of the run) score one row; once a
    // known combination misses, the caller has moved on: score the rest of the region.
    int s1 = s;
    if (!grew && f->current_scan_start_position == s) {
        s1 = f->stop_position_flanked;
        if (need_svr) {
            long per_row = (long)e.pattern.size() * 2 * MG_NFEAT * 8;
            long rows = std::max(1L, kMaxFeatureBytes / std::max(1L, per_row));
            s1 = (int)std::min<long>(s1, s + rows - 1);
        }
        if (s1 < s) s1 = s;
    }
    if (s < 1) return -1;
    run_batch(f, s, s1, need_svr, lrc);
    long idx = e.batch.index(s, capture, m->extension_arm_length, m->ligation_arm_length, strand);
    if (idx >= 0 && e.batch.valid[(size_t)idx] && same_sequences(e.batch, m)) return idx;
    return -1;
}

}  // namespace mgshim

// entry points of the batched driver (mipgen_b200/batched/mipgen_batched.h)
mg_ctx *mipgen_b200_shim_context() { return mgshim::ctx(); }

bool mipgen_b200_take_pending_svr(double *score)
{
    if (!mgshim::g_has_pending_svr) return false;
    mgshim::g_has_pending_svr = false;
    *score = mgshim::g_pending_svr;
    return true;
}

[[noreturn]] void mipgen_b200_fatal(const char *what, mg_ctx *c) { mgshim::fatal(what, c); }

// ------------------------------------------------------------------------------------
// Featurev5
// ------------------------------------------------------------------------------------
Featurev5::Featurev5(string chromosome, int start, int stop, int f, string l)
    : chr(chromosome), label(l), start_position(start), stop_position(stop), flank_size(f), start_position_flanked(start - f),
      stop_position_flanked(stop + f), mip_count(0), current_scan_start_position(0), chromosomal_sequence_start_position(0),
      chromosomal_sequence_stop_position(0)
{
    memset(long_range_content, 0, sizeof long_range_content);
    mgshim::registry().insert(this);
}

Featurev5::Featurev5()
    : start_position(0), stop_position(0), flank_size(0), start_position_flanked(0), stop_position_flanked(0), mip_count(0),
      current_scan_start_position(0), chromosomal_sequence_start_position(0), chromosomal_sequence_stop_position(0)
{
    memset(long_range_content, 0, sizeof long_range_content);
    mgshim::registry().insert(this);
}

Featurev5::Featurev5(const Featurev5 &o)
    : chr(o.chr), label(o.label), start_position(o.start_position), stop_position(o.stop_position), flank_size(o.flank_size),
      start_position_flanked(o.start_position_flanked), stop_position_flanked(o.stop_position_flanked), mip_count(o.mip_count),
      current_scan_start_position(o.current_scan_start_position), chromosomal_sequence(o.chromosomal_sequence),
      masked_chromosomal_sequence(o.masked_chromosomal_sequence), chromosomal_sequence_start_position(o.chromosomal_sequence_start_position),
      chromosomal_sequence_stop_position(o.chromosomal_sequence_stop_position)
{
    memcpy(long_range_content, o.long_range_content, sizeof long_range_content);
    mgshim::registry().insert(this);
}

Featurev5 &Featurev5::operator=(const Featurev5 &o)
{
    if (this != &o) {
        chr = o.chr; label = o.label; start_position = o.start_position; stop_position = o.stop_position; flank_size = o.flank_size;
        start_position_flanked = o.start_position_flanked; stop_position_flanked = o.stop_position_flanked; mip_count = o.mip_count;
        current_scan_start_position = o.current_scan_start_position; chromosomal_sequence = o.chromosomal_sequence;
        masked_chromosomal_sequence = o.masked_chromosomal_sequence;
        chromosomal_sequence_start_position = o.chromosomal_sequence_start_position;
        chromosomal_sequence_stop_position = o.chromosomal_sequence_stop_position;
        memcpy(long_range_content, o.long_range_content, sizeof long_range_content);
    }
    return *this;
}

Featurev5::~Featurev5()
{
    mgshim::registry().erase(this);
    if (mgshim::engine().batch.feature == this) mgshim::engine().batch.feature = nullptr;
}

void Featurev5::get_long_range_content(string extended_sequence, string feature_mers[])
{
    // Featurev5.cpp:18-56 on the device.  The 44 k-mers are fixed (mipgen.cpp:32); the caller's
    // array is accepted for signature compatibility only.
    (void)feature_mers;
    int denom = this->chromosomal_sequence_stop_position - this->chromosomal_sequence_start_position + 2001;
    if (mg_long_range_content(mgshim::ctx(), extended_sequence.data(), (int)extended_sequence.size(), denom, long_range_content) != MG_OK)
        mgshim::fatal("mg_long_range_content", mgshim::ctx());
}

bool Featurev5::operator<(Featurev5 &b)
{
    if (this->chr == b.chr) return this->start_position < b.start_position;
    return this->chr < b.chr;
}

// ------------------------------------------------------------------------------------
// SVMipv4 / PlusSVMipv4 / MinusSVMipv4
// ------------------------------------------------------------------------------------
SVMipv4::SVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length)
    : chr(chromosome), scan_start_position(scan_start), scan_stop_position(scan_stop), scan_size(scan_stop - scan_start + 1),
      extension_arm_length(ext_length), ligation_arm_length(lig_length), ext_probe_start(0), ext_probe_stop(0), lig_probe_start(0),
      lig_probe_stop(0), ext_probe_copy(0), lig_probe_copy(0), arm_fraction_masked(0), translocation_failed('0'), snp_failed('0'),
      mapping_failed('0'), masking_failed('0'), snp_count(0), has_snp_mip(false), score(0), b200_has_svr(false), b200_svr(0)
{
}

void SVMipv4::set_junction_scores()
{
    // The table itself lives on the device (K-feat's logistic epilogue).  The static map is
    // declared for source compatibility; mipgen.cpp defines it and never reads it.
}

double SVMipv4::get_score()
{
    long idx = mgshim::locate(this, false, nullptr);
    if (idx >= 0) return mgshim::engine().batch.logistic[(size_t)idx];
    double v = 0;
    mgshim::score_explicit(this, nullptr, &v, nullptr, nullptr);
    return v;
}

void SVMipv4::get_parameters(vector<double> &parameters, double long_range_content[])
{
    if (b200_has_svr) {
        // batched driver: the object was materialised from a device grid and carries its SVR score; the patched
        // predict_value picks it up instead of round-tripping 192 doubles through text (mipgen.cpp:1948-2019)
        mgshim::g_pending_svr = b200_svr;
        mgshim::g_has_pending_svr = true;
        return;
    }
    mgshim::Engine &e = mgshim::engine();
    parameters.resize(MG_NFEAT);
    long idx = mgshim::g_have_model ? mgshim::locate(this, true, long_range_content) : -1;
    if (idx >= 0) {
        memcpy(parameters.data(), &e.batch.feats[(size_t)idx * MG_NFEAT], MG_NFEAT * sizeof(double));
        e.last_svr = e.batch.svr[(size_t)idx];
        e.last_has_svr = true;
    } else {
        double svr = 0;
        mgshim::score_explicit(this, long_range_content, nullptr, mgshim::g_have_model ? &svr : nullptr, parameters.data());
        e.last_svr = svr;
        e.last_has_svr = mgshim::g_have_model;
    }
    memcpy(e.last_feats, parameters.data(), sizeof e.last_feats);
    e.last_valid = true;
}

PlusSVMipv4::PlusSVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length)
    : SVMipv4(chromosome, scan_start, scan_stop, ext_length, lig_length)
{
    strand = "+";
    ext_probe_stop = scan_start - 1;
    ext_probe_start = ext_probe_stop - ext_length + 1;
    lig_probe_start = scan_stop + 1;
    lig_probe_stop = scan_stop + lig_length;
}
void PlusSVMipv4::set_ext_probe_seq(string seq) { ext_probe_sequence = seq; }
void PlusSVMipv4::set_lig_probe_seq(string seq)
{
    lig_probe_sequence = seq;
    ligation_junction = seq.substr(0, 2);
}
void PlusSVMipv4::set_scan_target_seq(string seq) { scan_target_sequence = seq; }
int PlusSVMipv4::get_mip_start() { return ext_probe_start; }

// probe-orientation copy of a genomic window: reversed, A<->T and C<->G, anything else kept
static string to_minus_strand(const string &s)
{
    string out(s.rbegin(), s.rend());
    for (char &c : out) {
        if (c == 'A') c = 'T';
        else if (c == 'T') c = 'A';
        else if (c == 'C') c = 'G';
        else if (c == 'G') c = 'C';
    }
    return out;
}

MinusSVMipv4::MinusSVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length)
    : SVMipv4(chromosome, scan_start, scan_stop, ext_length, lig_length)
{
    strand = "-";
    ext_probe_start = scan_stop + 1;
    ext_probe_stop = scan_stop + ext_length;
    lig_probe_stop = scan_start - 1;
    lig_probe_start = lig_probe_stop - lig_length + 1;
}
void MinusSVMipv4::set_ext_probe_seq(std::string seq) { ext_probe_sequence = to_minus_strand(seq); }
void MinusSVMipv4::set_lig_probe_seq(std::string seq)
{
    lig_probe_sequence = to_minus_strand(seq);
    ligation_junction = lig_probe_sequence.substr(0, 2);
}
void MinusSVMipv4::set_scan_target_seq(string seq) { scan_target_sequence = to_minus_strand(seq); }
int MinusSVMipv4::get_mip_start() { return lig_probe_start; }

// ------------------------------------------------------------------------------------
// libsvm entry points
// ------------------------------------------------------------------------------------
struct svm_model { int n_sv; };
static svm_model g_model;

extern "C" struct svm_model *svm_load_model(const char *model_file_name)
{
    // svm.cpp:2761-2762: a missing file is a quiet NULL (mipgen.cpp:409 calls this in every
    // mode, and logistic runs ship no model) -- do not touch the device for that.
    FILE *fp = fopen(model_file_name, "rb");
    if (!fp) return nullptr;
    fclose(fp);
    if (mg_load_svr_model(mgshim::ctx(), model_file_name) != MG_OK) {
        fprintf(stderr, "[mipgen_b200] %s\n", mg_last_error(mgshim::ctx()));
        return nullptr;
    }
    mgshim::g_have_model = true;
    mg_model_info(mgshim::ctx(), &g_model.n_sv, nullptr, nullptr);
    return &g_model;
}

extern "C" int svm_get_nr_sv(const struct svm_model *model) { return model ? model->n_sv : 0; }

extern "C" void svm_free_and_destroy_model(struct svm_model **model_ptr_ptr)
{
    if (model_ptr_ptr) *model_ptr_ptr = nullptr;
}

extern "C" double svm_predict(const struct svm_model *model, const struct svm_node *x)
{
    if (!model || !mgshim::g_have_model) {
        fprintf(stderr, "[mipgen_b200] fatal: svm_predict without a loaded model (mipgen_svr.model must sit next to the executable)\n");
        exit(70);
    }
    mgshim::Engine &e = mgshim::engine();
    // densify (svm_node vectors are sparse: absent index == 0)
    double dense[MG_NFEAT];
    memset(dense, 0, sizeof dense);
    bool representable = true;
    for (const svm_node *p = x; p->index != -1; ++p) {
        if (p->index >= 1 && p->index <= MG_NFEAT) dense[p->index - 1] = p->value;
        else representable = false;
    }
    if (e.last_valid && e.last_has_svr && representable && memcmp(dense, e.last_feats, sizeof dense) == 0) return e.last_svr;
    if (!representable) {
        fprintf(stderr, "[mipgen_b200] fatal: svm_predict called with a feature index outside 1..192\n");
        exit(70);
    }
    double out = 0;
    if (mg_svr_predict(mgshim::ctx(), dense, 1, MG_NFEAT, &out) != MG_OK) mgshim::fatal("mg_svr_predict", mgshim::ctx());
    return out;
}
