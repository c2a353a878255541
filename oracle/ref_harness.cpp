/*
 * ref_harness.cpp -- thin C-ABI driver around the UNMODIFIED reference objects
 * (SVMipv4.o PlusSVMipv4.o MinusSVMipv4.o Featurev5.o svm.o, compiled by
 * oracle/Makefile straight from /root/reference into oracle/_ref/).
 *
 * TEST INFRASTRUCTURE ONLY.  It exists to (1) pin the C restatement in
 * mipgen_oracle.c, (2) generate tests/golden/, (3) serve as the CPU baseline
 * ("kind": "reference") in bench.py.  It mirrors what mipgen.cpp does around
 * the scoring classes (tile_regions 446-487, design_mip 602-613) but contains
 * no scoring arithmetic of its own: every number comes out of the reference's
 * get_score / get_parameters / get_long_range_content / svm_predict.
 */
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include "Featurev5.h"
#include "SVMipv4.h"
#include "PlusSVMipv4.h"
#include "MinusSVMipv4.h"
#include "svm.h"

map<string, double> SVMipv4::junction_scores; /* defined in mipgen.cpp:33 in the reference build */

/* mipgen.cpp:32 (data) */
static string feature_mers[] = {"A","AA","AAA","AAC","AAG","AAT","AC","ACA","ACC","ACG","AG","AGA","AGC","AGG","AGT","AT","ATA","ATC","ATG","CAG","CG","CGG","G","GAC","GAG","GC","GCG","GG","GGC","GGG","GTG","TA","TAA","TAC","TAG","TC","TCC","TCG","TG","TGA","TGC","TGG","TTC","TTG"};

static bool g_init = false;
static void init_once()
{
    if (!g_init) { SVMipv4::set_junction_scores(); g_init = true; }
}

extern "C" {

struct ref_region {
    const char *seq; int seq_len;
    int seq_start, seq_stop;
    int start_flanked, stop_flanked;
    const double *lrc;
    const int *copies;
};

struct ref_cfg {
    int max_capture, min_capture, capture_increment, max_mip_overlap;
    int n_pairs;
    const int *ext_len;
    const int *lig_len;
    int n_oligo_sizes; const int *oligo_sizes;
};

struct ref_mip {
    const char *ext; int ext_n;
    const char *lig; int lig_n;
    const char *tgt; int tgt_n;
    int ext_len, lig_len, scan_size;
    int ext_copy, lig_copy;
};

void ref_long_range_content(const char *ext_seq, int n, int seq_start, int seq_stop, double *out)
{
    Featurev5 f("1", seq_start, seq_stop, 0, "x");
    f.chromosomal_sequence_start_position = seq_start;
    f.chromosomal_sequence_stop_position = seq_stop;
    f.get_long_range_content(string(ext_seq, n), feature_mers);
    memcpy(out, f.long_range_content, sizeof(double) * MER_NUM);
}

/* explicit strings: build a PlusSVMipv4 and poke the public fields directly */
static void fill_from_strings(PlusSVMipv4 &mip, const ref_mip *m, int ext_tag, int lig_tag)
{
    mip.ext_probe_sequence = string(m->ext, m->ext_n);
    mip.lig_probe_sequence = string(m->lig, m->lig_n);
    mip.ligation_junction = mip.lig_probe_sequence.substr(0, 2);
    mip.scan_target_sequence = string(m->tgt, m->tgt_n);
    mip.mip_seq = mip.lig_probe_sequence + string(lig_tag, 'N') + "CTTCAGCTTCCCGATATCCGACGGTAGTGT" + string(ext_tag, 'N') + mip.ext_probe_sequence;
    mip.ext_probe_copy = m->ext_copy;
    mip.lig_probe_copy = m->lig_copy;
}

double ref_get_score(const ref_mip *m)
{
    init_once();
    PlusSVMipv4 mip("1", 1000, 1000 + m->scan_size - 1, m->ext_len, m->lig_len);
    fill_from_strings(mip, m, 5, 0);
    return mip.get_score();
}

void ref_get_parameters(const ref_mip *m, const double *lrc, double *out)
{
    init_once();
    PlusSVMipv4 mip("1", 1000, 1000 + m->scan_size - 1, m->ext_len, m->lig_len);
    fill_from_strings(mip, m, 5, 0);
    vector<double> p;
    double l[MER_NUM];
    memcpy(l, lrc, sizeof l);
    mip.get_parameters(p, l);
    memcpy(out, p.data(), sizeof(double) * p.size());
}

void *ref_svm_load_model(const char *path) { return svm_load_model(path); }
void ref_svm_free(void *m) { svm_model *mm = (svm_model *)m; svm_free_and_destroy_model(&mm); }
int ref_svm_nsv(void *m) { return ((svm_model *)m)->l; }

double ref_svm_predict(void *model, const double *x, int n)
{
    /* dense nodes exactly as predict_value builds them (mipgen.cpp:2001-2014) */
    vector<svm_node> nodes(n + 1);
    for (int i = 0; i < n; i++) { nodes[i].index = i + 1; nodes[i].value = x[i]; }
    nodes[n].index = -1;
    return svm_predict((svm_model *)model, nodes.data());
}

static int n_captures(const ref_cfg *c)
{
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int n = 0;
    for (int cap = c->max_capture; cap >= c->min_capture; cap -= inc) n++;
    return n;
}

/* Walk the full static grid like tile_regions does (mipgen.cpp:421-491) but without
 * the score-dependent skips; build each candidate the way tile_regions/design_mip do
 * and score it with the reference methods.  Same output layout as orc_grid_region. */
void ref_grid_region(const ref_region *r, const ref_cfg *c, void *model,
                     unsigned char *valid, double *logistic, double *svr, double *feats)
{
    init_once();
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int ncap = n_captures(c);
    int maxsum = 0;
    for (int i = 0; i < c->n_pairs; i++) maxsum = std::max(maxsum, c->ext_len[i] + c->lig_len[i]);
    string chromosomal_sequence(r->seq, r->seq_len);
    double lrc[MER_NUM];
    for (int i = 0; i < MER_NUM; i++) lrc[i] = r->lrc ? r->lrc[i] : 0.0;
    vector<double> params;
    string middle = string(0, 'N') + "CTTCAGCTTCCCGATATCCGACGGTAGTGT" + string(5, 'N');

    int cur = r->start_flanked - c->max_capture + maxsum;
    if (cur < 0) cur = 0;
    long idx = 0;
    while (cur < r->stop_flanked) {
        cur++;
        for (int ci = 0; ci < ncap; ci++) {
            int cap = c->max_capture - ci * inc;
            bool cap_skip = cap > r->stop_flanked - r->start_flanked + c->max_mip_overlap && cap - inc >= c->min_capture;
            for (int p = 0; p < c->n_pairs; p++) {
                int e = c->ext_len[p], l = c->lig_len[p];
                bool skip = cap_skip;
                if (cur - e <= 0 || cur - l <= 0) skip = true;
                if (cur + cap - e - 1 > r->seq_stop || cur + cap - l - 1 > r->seq_stop) skip = true;
                for (int strand = 0; strand < 2; strand++, idx++) {
                    if (skip) {
                        if (valid) valid[idx] = 0;
                        if (logistic) logistic[idx] = NAN;
                        if (svr) svr[idx] = NAN;
                        if (feats) for (int k = 0; k < 192; k++) feats[idx * 192 + k] = NAN;
                        continue;
                    }
                    SVMipv4 *mip;
                    int scan_stop = cur + cap - (e + l) - 1;
                    if (strand == 0) {
                        PlusSVMipv4 *pm = new PlusSVMipv4("1", cur, scan_stop, e, l);
                        pm->set_scan_target_seq(chromosomal_sequence.substr(pm->scan_start_position - r->seq_start, pm->scan_size));
                        mip = pm;
                    } else {
                        MinusSVMipv4 *mm = new MinusSVMipv4("1", cur, scan_stop, e, l);
                        mm->set_scan_target_seq(chromosomal_sequence.substr(mm->scan_start_position - r->seq_start, mm->scan_size));
                        mip = mm;
                    }
                    mip->set_ext_probe_seq(chromosomal_sequence.substr(mip->ext_probe_start - r->seq_start, mip->extension_arm_length));
                    mip->set_lig_probe_seq(chromosomal_sequence.substr(mip->lig_probe_start - r->seq_start, mip->ligation_arm_length));
                    mip->mip_seq = mip->lig_probe_sequence + middle + mip->ext_probe_sequence;
                    mip->ext_probe_copy = 1;
                    mip->lig_probe_copy = 1;
                    if (r->copies) {
                        mip->ext_probe_copy = 0; mip->lig_probe_copy = 0;
                        for (int k = 0; k < c->n_oligo_sizes; k++) {
                            if (c->oligo_sizes[k] == e) mip->ext_probe_copy = r->copies[(long)k * r->seq_len + (mip->ext_probe_start - r->seq_start)];
                            if (c->oligo_sizes[k] == l) mip->lig_probe_copy = r->copies[(long)k * r->seq_len + (mip->lig_probe_start - r->seq_start)];
                        }
                    }
                    if (valid) valid[idx] = 1;
                    if (logistic) logistic[idx] = mip->get_score();
                    if (svr || feats) {
                        mip->get_parameters(params, lrc);
                        if (feats) memcpy(feats + idx * 192, params.data(), sizeof(double) * 192);
                        if (svr) svr[idx] = model ? ref_svm_predict(model, params.data(), 192) : NAN;
                    }
                    delete mip;
                }
            }
        }
    }
}

} /* extern "C" */
