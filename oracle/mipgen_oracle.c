/*
 * mipgen_oracle.c -- CPU restatement of the MIPgen candidate enumeration +
 * scoring hot path in plain C.  TEST INFRASTRUCTURE ONLY (see mipgen_oracle.h).
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref) and the
 * golden vectors in tests/golden/ -- see tests/test_oracle_vs_ref.py.
 *
 * Compile with -O2 -ffp-contract=off (no FMA contraction: the reference is
 * built for baseline x86-64, where GCC cannot contract either).
 */
#include "mipgen_oracle.h"

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "logistic_terms.inc"

/* ------------------------------------------------------------------------- */
/* sequence helpers                                                           */
/* ------------------------------------------------------------------------- */

void orc_reverse_comp(const char *in, int n, char *out)
{
    /* MinusSVMipv4.cpp:6-29 */
    for (int i = 0; i < n; i++) {
        char c = in[n - 1 - i];
        switch (c) {
        case 'G': c = 'C'; break;
        case 'C': c = 'G'; break;
        case 'A': c = 'T'; break;
        case 'T': c = 'A'; break;
        default: break; /* every other character is passed through unchanged */
        }
        out[i] = c;
    }
    out[n] = 0;
}

int orc_count_mer(const char *s, int n, const char *sub, int k)
{
    /* SVMipv4.cpp:31-57: find(sub), then find(sub, offset+1) => overlapping matches */
    int count = 0;
    for (int i = 0; i + k <= n; i++)
        if (memcmp(s + i, sub, (size_t)k) == 0) count++;
    return count;
}

/* k-mer vocabularies.  arm_mers (SVMipv4.cpp:69) and insert_mers (:70) are the
 * pre-order walks of the ACGT trie to depth 2 and 3; junctions (:71) are the 16
 * dimers in lexicographic order. */
static const char BASES[4] = {'A', 'C', 'G', 'T'};

static int gen_mers(int depth, char out[][4])
{
    int n = 0;
    for (int a = 0; a < 4; a++) {
        out[n][0] = BASES[a]; out[n][1] = 0; n++;
        if (depth < 2) continue;
        for (int b = 0; b < 4; b++) {
            out[n][0] = BASES[a]; out[n][1] = BASES[b]; out[n][2] = 0; n++;
            if (depth < 3) continue;
            for (int c = 0; c < 4; c++) {
                out[n][0] = BASES[a]; out[n][1] = BASES[b]; out[n][2] = BASES[c]; out[n][3] = 0; n++;
            }
        }
    }
    return n;
}

/* mipgen.cpp:32 (data: which 44 strand-symmetric k-mers make up long_range_content) */
static const char *const FEATURE_MERS[ORC_NLRC] = {
    "A", "AA", "AAA", "AAC", "AAG", "AAT", "AC", "ACA", "ACC", "ACG", "AG",
    "AGA", "AGC", "AGG", "AGT", "AT", "ATA", "ATC", "ATG", "CAG", "CG", "CGG",
    "G", "GAC", "GAG", "GC", "GCG", "GG", "GGC", "GGG", "GTG", "TA", "TAA",
    "TAC", "TAG", "TC", "TCC", "TCG", "TG", "TGA", "TGC", "TGG", "TTC", "TTG"};

void orc_long_range_content(const char *ext_seq, int n, int denom, double out[ORC_NLRC])
{
    /* Featurev5.cpp:18-56 */
    for (int i = 0; i < ORC_NLRC; i++) {
        const char *mer = FEATURE_MERS[i];
        int k = (int)strlen(mer);
        double forward_count = orc_count_mer(ext_seq, n, mer, k);
        char rc[8];
        orc_reverse_comp(mer, k, rc);
        if (strcmp(rc, mer) != 0) {
            double reverse_count = orc_count_mer(ext_seq, n, rc, k);
            out[i] = (forward_count + reverse_count) / denom;
        } else {
            out[i] = forward_count / denom;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* one candidate                                                              */
/* ------------------------------------------------------------------------- */

static int find_char_before(const char *s, int n, char c, int limit)
{
    /* string::find(c) < limit */
    for (int i = 0; i < n && i < limit; i++)
        if (s[i] == c) return 1;
    return 0;
}

int orc_mip_invalid(const orc_mip *m)
{
    /* SVMipv4.cpp:63 and :116.  mip_seq = lig + N..backbone..N + ext (mipgen.cpp:199-200,605):
     * the backbone has no '-', so '-' in mip_seq <=> '-' in one of the arms. */
    if (find_char_before(m->ext, m->ext_n, 'N', m->ext_len)) return 1;
    if (find_char_before(m->lig, m->lig_n, 'N', m->lig_len)) return 1;
    if (memchr(m->ext, '-', (size_t)m->ext_n)) return 1;
    if (memchr(m->lig, '-', (size_t)m->lig_n)) return 1;
    return 0;
}

static double log_copy(int copy)
{
    /* SVMipv4.cpp:109-110, 173-174 */
    return copy > 100 ? 2 : log10((double)copy);
}

static int push_block(const char *s, int n, int len, int depth, double *out)
{
    /* one "mer block" of get_parameters (SVMipv4.cpp:72-79, 85-92, 94-101):
     * counts / (len - k + 1) in trie pre-order, with (G+C)/len inserted before "T" */
    char mers[84][4];
    int nm = gen_mers(depth, mers);
    int o = 0;
    for (int i = 0; i < nm; i++) {
        int k = (int)strlen(mers[i]);
        if (k == 1 && mers[i][0] == 'T') {
            double gc = (double)orc_count_mer(s, n, "G", 1) + (double)orc_count_mer(s, n, "C", 1);
            out[o++] = gc / (double)(len - 1 + 1);
        }
        out[o++] = (double)orc_count_mer(s, n, mers[i], k) / (len - k + 1.);
    }
    return o;
}

void orc_get_parameters(const orc_mip *m, const double lrc[ORC_NLRC], double out[ORC_NFEAT])
{
    /* SVMipv4.cpp:60-113 */
    if (orc_mip_invalid(m)) {
        for (int i = 0; i < ORC_NFEAT; i++) out[i] = 0;
        return;
    }
    int o = 0;
    o += push_block(m->ext, m->ext_n, m->ext_len, 2, out + o);         /* 1..21   */
    out[o++] = m->ext_len;                                               /* 22      */
    for (int i = 0; i < ORC_NLRC; i++) out[o++] = lrc[i];                /* 23..66  */
    o += push_block(m->tgt, m->tgt_n, m->scan_size, 3, out + o);         /* 67..151 */
    out[o++] = m->scan_size;                                             /* 152     */
    o += push_block(m->lig, m->lig_n, m->lig_len, 2, out + o);           /* 153..173*/
    out[o++] = m->lig_len;                                               /* 174     */
    for (int a = 0; a < 4; a++)                                          /* 175..190*/
        for (int b = 0; b < 4; b++)
            out[o++] = (m->lig_n >= 2 && m->lig[0] == BASES[a] && m->lig[1] == BASES[b]) ? 1 : 0;
    out[o++] = log_copy(m->ext_copy);                                    /* 191     */
    out[o++] = log_copy(m->lig_copy);                                    /* 192     */
}

static double junction_score(const char *lig, int n)
{
    /* SVMipv4.cpp:249-267 (data).  Unknown key => map::operator[] inserts 0.0 (:171). */
    static const double T[16] = {
        /* AA */ 0.0,   /* AC */ 0.35,  /* AG */ 0.046, /* AT */ 0.079,
        /* CA */ 0.34,  /* CC */ 0.22,  /* CG */ 0.55,  /* CT */ -0.071,
        /* GA */ 0.35,  /* GC */ 0.92,  /* GG */ 0.24,  /* GT */ 0.48,
        /* TA */ -0.46, /* TC */ -0.35, /* TG */ -0.25, /* TT */ -0.98};
    if (n < 2) return 0.0;
    int a = -1, b = -1;
    for (int i = 0; i < 4; i++) {
        if (lig[0] == BASES[i]) a = i;
        if (lig[1] == BASES[i]) b = i;
    }
    if (a < 0 || b < 0) return 0.0;
    return T[a * 4 + b];
}

static double count_char(const char *s, int n, char c)
{
    int k = 0;
    for (int i = 0; i < n; i++) k += (s[i] == c);
    return (double)k;
}

double orc_get_score(const orc_mip *m)
{
    /* SVMipv4.cpp:114-248 */
    if (orc_mip_invalid(m)) return -1000.0;

    /* :118-141 run counting.  last_base only changes on a "switch". */
    char last = m->tgt_n > 0 ? m->tgt[0] : 0;
    double run_count = 0;
    for (int i = 1; i < m->scan_size && i < m->tgt_n; i++) {
        char cur = m->tgt[i];
        if (cur == 'G' || cur == 'C') {
            if (!(last == 'G' || last == 'C')) { run_count++; last = cur; }
        } else {
            if (!(last == 'A' || last == 'T')) { run_count++; last = cur; }
        }
    }
    run_count++;

    double v[MG_LOGIT_NVARS];
    v[MG_V_BASES_PER_SWITCH] = m->scan_size / run_count;
    double ext_g = count_char(m->ext, m->ext_n, 'G');
    double lig_g = count_char(m->lig, m->lig_n, 'G');
    double tgt_g = count_char(m->tgt, m->tgt_n, 'G');
    double ext_gc = count_char(m->ext, m->ext_n, 'C') + ext_g;
    double lig_gc = count_char(m->lig, m->lig_n, 'C') + lig_g;
    double tgt_gc = count_char(m->tgt, m->tgt_n, 'C') + tgt_g;
    double ext_a = count_char(m->ext, m->ext_n, 'A');
    double lig_a = count_char(m->lig, m->lig_n, 'A');
    double tgt_a = count_char(m->tgt, m->tgt_n, 'A');
    double ext_length = m->ext_len, lig_length = m->lig_len;
    v[MG_V_EXT_LENGTH] = ext_length;
    v[MG_V_LIG_LENGTH] = lig_length;
    v[MG_V_TARGET_LENGTH] = m->scan_size > 250 ? 250 : m->scan_size;
    v[MG_V_EXT_GC_CONTENT] = ext_gc / ext_length;
    v[MG_V_LIG_GC_CONTENT] = lig_gc / lig_length;
    v[MG_V_TARGET_GC_CONTENT] = tgt_gc / m->scan_size;
    v[MG_V_EXT_G_CONTENT] = ext_g / ext_length;
    v[MG_V_LIG_G_CONTENT] = lig_g / lig_length;
    v[MG_V_TARGET_G_CONTENT] = tgt_g / m->scan_size;
    v[MG_V_EXT_A_CONTENT] = ext_a / ext_length;
    v[MG_V_LIG_A_CONTENT] = lig_a / lig_length;
    v[MG_V_TARGET_A_CONTENT] = tgt_a / m->scan_size;
    v[MG_V_JUNCTION_SCORE] = junction_score(m->lig, m->lig_n);
    v[MG_V_LOG_EXT_COPY] = log_copy(m->ext_copy);
    v[MG_V_LOG_LIG_COPY] = log_copy(m->lig_copy);

    /* :176-246, summed left to right; products are (coef*a)*b; pow(x,2) == x*x */
    volatile double exponent = MG_LOGIT_C0 - MG_LOGIT_C1;
#define LIN(c, a, b) ((c) * v[a])
#define PROD(c, a, b) (((c) * v[a]) * v[b])
#define SQ(c, a, b) ((c) * (v[a] * v[a]))
#define X(c, kind, a, b) exponent = exponent + kind(c, a, b);
    MG_LOGIT_TERMS(X)
#undef X
#undef LIN
#undef PROD
#undef SQ
    double e = exponent;
    return pow(2.71828, e) / (1 + pow(2.71828, e)); /* :247 -- base is the literal 2.71828 */
}

/* ------------------------------------------------------------------------- */
/* geometry + design                                                          */
/* ------------------------------------------------------------------------- */

void orc_geometry(int scan_start, int scan_stop, int ext_len, int lig_len, int strand, orc_geom *g)
{
    g->scan_start = scan_start; g->scan_stop = scan_stop;
    g->ext_len = ext_len; g->lig_len = lig_len; g->strand = strand;
    if (strand == 0) { /* PlusSVMipv4.cpp:9-12 */
        g->ext_start = scan_start - ext_len; g->ext_stop = scan_start - 1;
        g->lig_start = scan_stop + 1;        g->lig_stop = scan_stop + lig_len;
    } else {           /* MinusSVMipv4.cpp:32-35 */
        g->ext_start = scan_stop + 1;        g->ext_stop = scan_stop + ext_len;
        g->lig_start = scan_start - lig_len; g->lig_stop = scan_start - 1;
    }
}

static int cut(const char *seq, int seq_len, int off, int len, int strand, char *out)
{
    /* std::string::substr(off, len): throws if off > size, clamps len otherwise */
    if (off < 0 || off > seq_len) return -1;
    int n = len;
    if (off + n > seq_len) n = seq_len - off;
    if (n < 0) n = 0;
    if (strand == 0) { memcpy(out, seq + off, (size_t)n); out[n] = 0; }
    else orc_reverse_comp(seq + off, n, out);
    return n;
}

int orc_design(const char *seq, int seq_len, int seq_start, const orc_geom *g,
               char *ext, char *lig, char *tgt, orc_mip *m)
{
    int scan_size = g->scan_stop - g->scan_start + 1; /* SVMipv4.cpp:27 */
    /* mipgen.cpp:461-462 */
    int nt = cut(seq, seq_len, g->scan_start - seq_start, scan_size, g->strand, tgt);
    /* mipgen.cpp:602-603 */
    int ne = cut(seq, seq_len, g->ext_start - seq_start, g->ext_len, g->strand, ext);
    int nl = cut(seq, seq_len, g->lig_start - seq_start, g->lig_len, g->strand, lig);
    if (nt < 0 || ne < 0 || nl < 0) return -1;
    m->ext = ext; m->ext_n = ne; m->lig = lig; m->lig_n = nl; m->tgt = tgt; m->tgt_n = nt;
    m->ext_len = g->ext_len; m->lig_len = g->lig_len; m->scan_size = scan_size;
    m->ext_copy = 1; m->lig_copy = 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* libsvm subset                                                              */
/* ------------------------------------------------------------------------- */

typedef struct { int index; double value; } orc_node;

struct orc_model {
    int svm_type, kernel_type, degree, nr_class, l;
    double gamma, coef0, rho;
    double *sv_coef;   /* [l] */
    orc_node **sv;     /* [l] -> into space, terminated by index -1 */
    orc_node *space;
};

static const char *const SVM_TYPES[] = {"c_svc", "nu_svc", "one_class", "epsilon_svr", "nu_svr", NULL};
static const char *const KERNEL_TYPES[] = {"linear", "polynomial", "rbf", "sigmoid", "precomputed", NULL};

static int lookup(const char *const *tab, const char *s)
{
    for (int i = 0; tab[i]; i++) if (strcmp(tab[i], s) == 0) return i;
    return -1;
}

static char *read_line(FILE *fp, char **buf, size_t *cap)
{
    size_t len = 0;
    if (!*buf) { *cap = 1024; *buf = (char *)malloc(*cap); }
    for (;;) {
        if (!fgets(*buf + len, (int)(*cap - len), fp)) return len ? *buf : NULL;
        len += strlen(*buf + len);
        if (len && (*buf)[len - 1] == '\n') return *buf;
        *cap *= 2; *buf = (char *)realloc(*buf, *cap);
    }
}

orc_model *orc_svm_load_model(const char *path)
{
    /* svm.cpp:2759-2973 */
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    char cmd[81];
    for (;;) {
        if (fscanf(fp, "%80s", cmd) != 1) goto fail;
        if (!strcmp(cmd, "svm_type")) {
            if (fscanf(fp, "%80s", cmd) != 1 || (m->svm_type = lookup(SVM_TYPES, cmd)) < 0) goto fail;
        } else if (!strcmp(cmd, "kernel_type")) {
            if (fscanf(fp, "%80s", cmd) != 1 || (m->kernel_type = lookup(KERNEL_TYPES, cmd)) < 0) goto fail;
        } else if (!strcmp(cmd, "degree")) { if (fscanf(fp, "%d", &m->degree) != 1) goto fail; }
        else if (!strcmp(cmd, "gamma")) { if (fscanf(fp, "%lf", &m->gamma) != 1) goto fail; }
        else if (!strcmp(cmd, "coef0")) { if (fscanf(fp, "%lf", &m->coef0) != 1) goto fail; }
        else if (!strcmp(cmd, "nr_class")) { if (fscanf(fp, "%d", &m->nr_class) != 1) goto fail; }
        else if (!strcmp(cmd, "total_sv")) { if (fscanf(fp, "%d", &m->l) != 1) goto fail; }
        else if (!strcmp(cmd, "rho")) {
            int n = m->nr_class * (m->nr_class - 1) / 2;
            for (int i = 0; i < n; i++) { double r; if (fscanf(fp, "%lf", &r) != 1) goto fail; if (i == 0) m->rho = r; }
        } else if (!strcmp(cmd, "label") || !strcmp(cmd, "nr_sv")) {
            for (int i = 0; i < m->nr_class; i++) { int d; if (fscanf(fp, "%d", &d) != 1) goto fail; }
        } else if (!strcmp(cmd, "probA") || !strcmp(cmd, "probB")) {
            int n = m->nr_class * (m->nr_class - 1) / 2;
            for (int i = 0; i < n; i++) { double d; if (fscanf(fp, "%lf", &d) != 1) goto fail; }
        } else if (!strcmp(cmd, "SV")) {
            int c;
            while ((c = getc(fp)) != EOF && c != '\n') {}
            break;
        } else goto fail; /* "unknown text in model file" */
    }
    if (m->nr_class != 2) goto fail; /* regression / one-class models only on this path */
    {
        long pos = ftell(fp);
        char *buf = NULL; size_t cap = 0;
        long elements = 0;
        while (read_line(fp, &buf, &cap))
            for (char *p = buf; *p; p++) if (*p == ':') elements++;
        elements += m->l;
        fseek(fp, pos, SEEK_SET);
        m->sv_coef = (double *)calloc((size_t)(m->l > 0 ? m->l : 1), sizeof(double));
        m->sv = (orc_node **)calloc((size_t)(m->l > 0 ? m->l : 1), sizeof(orc_node *));
        m->space = (orc_node *)calloc((size_t)(elements > 0 ? elements : 1), sizeof(orc_node));
        long j = 0;
        for (int i = 0; i < m->l; i++) {
            if (!read_line(fp, &buf, &cap)) { free(buf); goto fail; }
            m->sv[i] = &m->space[j];
            char *save = NULL, *endp;
            char *p = strtok_r(buf, " \t", &save);
            m->sv_coef[i] = p ? strtod(p, &endp) : 0.0;
            for (;;) {
                char *idx = strtok_r(NULL, ":", &save);
                char *val = strtok_r(NULL, " \t", &save);
                if (!val) break;
                m->space[j].index = (int)strtol(idx, &endp, 10);
                m->space[j].value = strtod(val, &endp);
                j++;
            }
            m->space[j++].index = -1;
        }
        free(buf);
    }
    fclose(fp);
    return m;
fail:
    fclose(fp);
    orc_svm_free(m);
    return NULL;
}

void orc_svm_free(orc_model *m)
{
    if (!m) return;
    free(m->sv_coef); free(m->sv); free(m->space); free(m);
}

int orc_svm_nsv(const orc_model *m) { return m->l; }
double orc_svm_gamma(const orc_model *m) { return m->gamma; }
double orc_svm_rho(const orc_model *m) { return m->rho; }

static double k_rbf(const double *x, int n, const orc_node *y, double gamma)
{
    /* svm.cpp:328-368, x dense: node j has index j+1 */
    double sum = 0;
    int xi = 0;
    while (xi < n && y->index != -1) {
        int xindex = xi + 1;
        if (xindex == y->index) {
            double d = x[xi] - y->value;
            sum += d * d; ++xi; ++y;
        } else if (xindex > y->index) {
            sum += y->value * y->value; ++y;
        } else {
            sum += x[xi] * x[xi]; ++xi;
        }
    }
    while (xi < n) { sum += x[xi] * x[xi]; ++xi; }
    while (y->index != -1) { sum += y->value * y->value; ++y; }
    return exp(-gamma * sum);
}

double orc_svm_predict(const orc_model *m, const double *x, int n)
{
    /* svm.cpp:2580-2593 -> 2504-2522 (EPSILON_SVR / NU_SVR branch) */
    double sum = 0;
    for (int i = 0; i < m->l; i++)
        sum += m->sv_coef[i] * k_rbf(x, n, m->sv[i], m->gamma);
    sum -= m->rho;
    return sum;
}

double orc_predict_value(const orc_model *m, const double *x, int n)
{
    /* mipgen.cpp:1948-2019: boost::lexical_cast<string>(double) prints 17 significant
     * digits (boost/detail/lcast_precision.hpp:81-83); strtod parses them back. */
    double *y = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    char buf[64];
    for (int i = 0; i < n; i++) {
        snprintf(buf, sizeof buf, "%.17g", x[i]);
        y[i] = strtod(buf, NULL);
    }
    double r = orc_svm_predict(m, y, n);
    free(y);
    return r;
}

/* ------------------------------------------------------------------------- */
/* region grid + tile replay                                                  */
/* ------------------------------------------------------------------------- */

int orc_n_captures(const orc_cfg *c)
{
    /* mipgen.cpp:427 (capture_increment==0 is forced to 1 at :274) */
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int n = 0;
    for (int cap = c->max_capture; cap >= c->min_capture; cap -= inc) n++;
    return n;
}

static int max_sum(const orc_cfg *c)
{
    int m = 0;
    for (int i = 0; i < c->n_pairs; i++) {
        int s = c->ext_len[i] + c->lig_len[i];
        if (s > m) m = s;
    }
    return m;
}

static int min_sum(const orc_cfg *c)
{
    int m = 1 << 30;
    for (int i = 0; i < c->n_pairs; i++) {
        int s = c->ext_len[i] + c->lig_len[i];
        if (s < m) m = s;
    }
    return m;
}

int orc_first_scan_start(const orc_region *r, const orc_cfg *c)
{
    /* mipgen.cpp:421-425: the loop pre-increments, so the first scanned start is +1 */
    int cur = r->start_flanked - c->max_capture + max_sum(c);
    if (cur < 0) cur = 0;
    return cur + 1;
}

int orc_n_scan(const orc_region *r, const orc_cfg *c)
{
    int n = r->stop_flanked - orc_first_scan_start(r, c) + 1;
    return n < 0 ? 0 : n;
}

long orc_grid_size(const orc_region *r, const orc_cfg *c)
{
    return (long)orc_n_scan(r, c) * orc_n_captures(c) * c->n_pairs * 2;
}

static int cap_skipped(const orc_region *r, const orc_cfg *c, int cap)
{
    /* mipgen.cpp:429 */
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    return cap > r->stop_flanked - r->start_flanked + c->max_mip_overlap && cap - inc >= c->min_capture;
}

static int pair_skipped(const orc_region *r, int s, int cap, int e, int l)
{
    /* mipgen.cpp:443-444 */
    if (s - e <= 0 || s - l <= 0) return 1;
    if (s + cap - e - 1 > r->seq_stop || s + cap - l - 1 > r->seq_stop) return 1;
    return 0;
}

static int copy_lookup(const orc_region *r, const orc_cfg *c, int start, int stop)
{
    /* mipgen.cpp:612-613: copy_chr_start_stop[chr][start][stop]; absent key => 0 */
    if (!r->copies) return 1;
    int size = stop - start + 1;
    for (int k = 0; k < c->n_oligo_sizes; k++)
        if (c->oligo_sizes[k] == size) {
            int i = start - r->seq_start;
            if (i < 0 || i >= r->seq_len) return 0;
            return r->copies[(long)k * r->seq_len + i];
        }
    return 0;
}

void orc_grid_region(const orc_region *r, const orc_cfg *c, const orc_model *model,
                     unsigned char *valid, double *logistic, double *svr, double *feats)
{
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int ncap = orc_n_captures(c), nscan = orc_n_scan(r, c), s0 = orc_first_scan_start(r, c);
    static const double zero_lrc[ORC_NLRC] = {0};
    const double *lrc = r->lrc ? r->lrc : zero_lrc;
    char ext[512], lig[512], *tgt = (char *)malloc((size_t)c->max_capture + 16);
    double x[ORC_NFEAT];
    long idx = 0;
    for (int si = 0; si < nscan; si++) {
        int s = s0 + si;
        for (int ci = 0; ci < ncap; ci++) {
            int cap = c->max_capture - ci * inc;
            for (int p = 0; p < c->n_pairs; p++) {
                int e = c->ext_len[p], l = c->lig_len[p];
                int skip = cap_skipped(r, c, cap) || pair_skipped(r, s, cap, e, l);
                for (int strand = 0; strand < 2; strand++, idx++) {
                    orc_geom g; orc_mip m;
                    int ok = !skip;
                    if (ok) {
                        orc_geometry(s, s + cap - (e + l) - 1, e, l, strand, &g); /* :449 */
                        if (orc_design(r->seq, r->seq_len, r->seq_start, &g, ext, lig, tgt, &m) != 0) ok = 0;
                    }
                    if (valid) valid[idx] = (unsigned char)ok;
                    if (!ok) {
                        if (logistic) logistic[idx] = NAN;
                        if (svr) svr[idx] = NAN;
                        if (feats) for (int k = 0; k < ORC_NFEAT; k++) feats[idx * ORC_NFEAT + k] = NAN;
                        continue;
                    }
                    m.ext_copy = copy_lookup(r, c, g.ext_start, g.ext_stop);
                    m.lig_copy = copy_lookup(r, c, g.lig_start, g.lig_stop);
                    if (logistic) logistic[idx] = orc_get_score(&m);
                    if (svr || feats) {
                        orc_get_parameters(&m, lrc, x);
                        if (feats) memcpy(feats + idx * ORC_NFEAT, x, sizeof x);
                        if (svr) svr[idx] = model ? orc_svm_predict(model, x, ORC_NFEAT) : NAN;
                    }
                }
            }
        }
    }
    free(tgt);
}

long orc_tile_replay(const orc_region *r, const orc_cfg *c, const unsigned char *valid,
                     const double *score, int method, int heuristic, double upper_score_limit,
                     long *out_idx, long cap_out)
{
    /* mipgen.cpp:423-501 */
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int ncap = orc_n_captures(c), nscan = orc_n_scan(r, c);
    int smallest = min_sum(c);
    long n = 0;
    for (int si = 0; si < nscan; si++) {
        double previous_best_score = 0;                                        /* :426 */
        for (int ci = 0; ci < ncap; ci++) {
            int cap = c->max_capture - ci * inc;
            if (cap_skipped(r, c, cap)) continue;                               /* :429 */
            if (previous_best_score > upper_score_limit) continue;              /* :430 */
            int p = 0;
            while (p < c->n_pairs) {
                int sum = c->ext_len[p] + c->lig_len[p];
                int q = p;
                while (q < c->n_pairs && c->ext_len[q] + c->lig_len[q] == sum) q++;
                /* pairs [p,q) share one arm_length_sum (:431-438) */
                if (!(previous_best_score > upper_score_limit && sum != smallest)) { /* :434 */
                    int previous_minus_score = 0, previous_plus_score = 0;      /* :435-436 (int!) */
                    int skip_ahead = 0;
                    for (int k = p; k < q; k++) {
                        if (skip_ahead) continue;
                        long idx = ((((long)si * ncap + ci) * c->n_pairs + k) * 2);
                        if (!valid[idx]) continue;                               /* :443-444 */
                        double plus = score[idx], minus = score[idx + 1];
                        if (n + 2 <= cap_out) { out_idx[n] = idx; out_idx[n + 1] = idx + 1; }
                        n += 2;
                        if (method == 0 && heuristic && plus < previous_plus_score && minus < previous_minus_score)
                            skip_ahead = 1;                                      /* :494 */
                        previous_best_score = (minus > plus) ? minus : plus;    /* :495 */
                        previous_minus_score = (int)minus;                       /* :496 */
                        previous_plus_score = (int)plus;                         /* :497 */
                    }
                }
                p = q;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* selection front-end                                                        */
/* ------------------------------------------------------------------------- */

typedef struct { int si, ci, p, strand, s, cap, e, l, scan_stop, ext_copy, lig_copy; } orc_dec;

static void decode_idx(const orc_region *r, const orc_cfg *c, long idx, orc_dec *d)
{
    int inc = c->capture_increment == 0 ? 1 : c->capture_increment;
    int ncap = orc_n_captures(c);
    d->strand = (int)(idx & 1);
    long q = idx >> 1;
    d->p = (int)(q % c->n_pairs); q /= c->n_pairs;
    d->ci = (int)(q % ncap);
    d->si = (int)(q / ncap);
    d->s = orc_first_scan_start(r, c) + d->si;
    d->cap = c->max_capture - d->ci * inc;
    d->e = c->ext_len[d->p]; d->l = c->lig_len[d->p];
    d->scan_stop = d->s + d->cap - (d->e + d->l) - 1;
    orc_geom g;
    orc_geometry(d->s, d->scan_stop, d->e, d->l, d->strand, &g);
    d->ext_copy = copy_lookup(r, c, g.ext_start, g.ext_stop);
    d->lig_copy = copy_lookup(r, c, g.lig_start, g.lig_stop);
}

int orc_n_positions(const orc_region *r, const orc_cfg *c)
{
    int last = r->stop_flanked + c->max_capture - min_sum(c) - 1;
    int n = last - orc_first_scan_start(r, c) + 1;
    return n < 0 ? 0 : n;
}

/* what design_mip leaves in the object for the selection code (mipgen.cpp:606-625, 634-760) */
typedef struct { double masked; int snp, mapping_failed; } orc_selmip;

static int count_in(const char *s, int seq_len, int off, int len, char what)
{
    /* std::string::substr(off, len) clamps the length; std::count over the piece */
    int n = 0;
    if (off < 0 || off > seq_len) return 0;
    for (int i = off; i < off + len && i < seq_len; i++) n += s[i] == what;
    return n;
}

static void sel_fields(const orc_region *r, const orc_cfg *c, const orc_sel *sel, const orc_dec *d, orc_selmip *m)
{
    orc_geom g;
    orc_geometry(d->s, d->scan_stop, d->e, d->l, d->strand, &g);
    const char *masked = sel->masked_seq ? sel->masked_seq : r->seq;
    double ext_n = count_in(masked, r->seq_len, g.ext_start - r->seq_start, d->e, 'N');   /* :606-609 */
    double lig_n = count_in(masked, r->seq_len, g.lig_start - r->seq_start, d->l, 'N');
    m->masked = (ext_n + lig_n) / (d->l + d->e);                                          /* :610 */
    m->mapping_failed = 0;
    m->snp = 0;
    if (sel->unmappable) {                                                                /* :615-625 */
        int mip_start = (d->strand ? g.lig_start : g.ext_start) - r->seq_start;           /* get_mip_start() */
        if (mip_start >= 0 && mip_start < r->seq_len && sel->unmappable[(long)d->ci * r->seq_len + mip_start]) {
            m->mapping_failed = 1;
            return;  /* design_mip returns before the SNP scan */
        }
    }
    if (sel->snp) {                                                                       /* :634-636, 698-700 */
        for (int i = g.ext_start; i <= g.ext_stop; i++) { int o = i - r->seq_start; if (o >= 0 && o < r->seq_len && sel->snp[o]) m->snp++; }
        for (int i = g.lig_start; i <= g.lig_stop; i++) { int o = i - r->seq_start; if (o >= 0 && o < r->seq_len && sel->snp[o]) m->snp++; }
    }
}

void orc_condense(const orc_region *r, const orc_cfg *c, const orc_sel *sel, const double *score,
                  const long *enum_idx, long n_enum, long *scan_best)
{
    /* mipgen.cpp:1670-1746 */
    int nscan = orc_n_scan(r, c);
    for (long i = 0; i < 2L * nscan; i++) scan_best[i] = -1;
    /* the candidates of one scan start are contiguous in enum_idx; walk each block backwards per strand */
    long b = 0;
    while (b < n_enum) {
        orc_dec d0; decode_idx(r, c, enum_idx[b], &d0);
        long e = b;
        while (e < n_enum) { orc_dec d; decode_idx(r, c, enum_idx[e], &d); if (d.si != d0.si) break; e++; }
        int chosen_copy_count = 0;                  /* :1677-1680: declared per position, outside the strand loop */
        double chosen_masked_arm_proportion = 0.0;
        for (int strand = 0; strand < 2; strand++) {
            int skip_ahead = 0;
            long best = -1;
            int best_snp = 0;
            for (long k = e - 1; k >= b; k--) {
                if (skip_ahead) continue;
                orc_dec d; decode_idx(r, c, enum_idx[k], &d);
                if (d.strand != strand) continue;
                if (d.ext_copy * d.lig_copy > sel->max_arm_copy) continue;                 /* :1689 */
                orc_selmip m; sel_fields(r, c, sel, &d, &m);
                if (m.mapping_failed) continue;                                            /* :1690 */
                int current = d.ext_copy > d.lig_copy ? d.ext_copy : d.lig_copy;          /* :1692 */
                double current_masked = m.masked;                                         /* :1693 */
                double sc = score[enum_idx[k]];
                if (best < 0) {                                                            /* :1695-1700 */
                    best = enum_idx[k]; best_snp = m.snp; chosen_masked_arm_proportion = current_masked; chosen_copy_count = current;
                } else if (current_masked > sel->masked_arm_threshold && current_masked < chosen_masked_arm_proportion) {  /* :1701-1706 */
                    best = enum_idx[k]; best_snp = m.snp; chosen_masked_arm_proportion = current_masked; chosen_copy_count = current;
                } else if (current > sel->target_arm_copy && current < chosen_copy_count) {  /* :1709-1714 */
                    best = enum_idx[k]; best_snp = m.snp; chosen_masked_arm_proportion = current_masked; chosen_copy_count = current;
                } else if (current <= sel->target_arm_copy) {                              /* :1715 */
                    if (sc < sel->lower_score_limit && sc > score[best]) {                 /* :1717-1722 */
                        best = enum_idx[k]; best_snp = m.snp; chosen_masked_arm_proportion = current_masked; chosen_copy_count = current;
                    } else if (sc > sel->lower_score_limit) {
                        if (m.snp < best_snp) {                                            /* :1725-1730 */
                            best = enum_idx[k]; best_snp = m.snp; chosen_masked_arm_proportion = current_masked; chosen_copy_count = current;
                        } else if (m.snp == best_snp) {                                    /* :1731-1737 */
                            if (sc > score[best]) {
                                best = enum_idx[k];
                                if (sc > sel->upper_score_limit) skip_ahead = 1;
                            }
                        }
                    }
                }
            }
            scan_best[2L * d0.si + strand] = best;
        }
        b = e;
    }
}

void orc_collapse(const orc_region *r, const orc_cfg *c, const orc_sel *sel, const double *score,
                  const long *scan_best, long *pos_best)
{
    /* mipgen.cpp:1617-1649 */
    int nscan = orc_n_scan(r, c), npos = orc_n_positions(r, c), s0 = orc_first_scan_start(r, c);
    for (long i = 0; i < 2L * npos; i++) pos_best[i] = -1;
    int *pos_snp = (int *)calloc((size_t)(2L * npos + 1), sizeof(int));
    for (int si = 0; si < nscan; si++)
        for (int strand = 0; strand < 2; strand++) {
            long cur = scan_best[2L * si + strand];
            if (cur < 0) continue;
            orc_dec d; decode_idx(r, c, cur, &d);
            if (d.ext_copy * d.lig_copy > sel->max_arm_copy || d.ext_copy > sel->target_arm_copy || d.lig_copy > sel->target_arm_copy) continue;  /* :1628 */
            orc_selmip m; sel_fields(r, c, sel, &d, &m);
            if (m.masked > sel->masked_arm_threshold) continue;                            /* :1629 */
            for (int pos = d.s; pos <= d.scan_stop; pos++) {
                long *slot = &pos_best[2L * (pos - s0) + strand];
                int *ssnp = &pos_snp[2L * (pos - s0) + strand];
                if (*slot < 0) { *slot = cur; *ssnp = m.snp; }                             /* :1634-1637 */
                else if (m.snp < *ssnp) { *slot = cur; *ssnp = m.snp; }                    /* :1638-1641 */
                else if (score[cur] > score[*slot] && m.snp == *ssnp) { *slot = cur; *ssnp = m.snp; }  /* :1642-1645 */
            }
        }
    free(pos_snp);
}
