/* c_abi_example.c -- the C-ABI of libmipgen_b200.so from plain C (C99), the way a binding of the reference would use it
 * (INTEGRATION.md route B): contexts on the devices named on the command line, the default arm table, a libsvm model, a few
 * regions read from a text file, mg_score_regions_multi (regions sharded over the contexts, results in the caller's order), the
 * device-side selection through mg_tile_regions_multi and the opt-in copy counting.  Prints sums the tests compare with the
 * same calls made through the Python view of the C-ABI.
 *
 *   gcc -std=c99 -O1 -I include tests/c/c_abi_example.c -L mipgen_b200 -lmipgen_b200 -Wl,-rpath,$PWD/mipgen_b200 -o c_abi_example
 *   ./c_abi_example <model> <regions.txt> <device> [<device> ...]
 * regions.txt: one region per line: seq_start seq_stop start_flanked stop_flanked SEQUENCE
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mipgen_b200.h"

#define MAX_REGIONS 64
#define MAX_CTX 16

static void die(const char *what, mg_ctx *ctx)
{
    fprintf(stderr, "c_abi_example: %s: %s\n", what, mg_last_error(ctx));
    exit(1);
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s model regions.txt device [device ...]\n", argv[0]); return 2; }
    /* the default arm table of mipgen.cpp:243-259: arm sums 45..40, ligation arm 18..(sum - 16), in enumeration order */
    int ext[256], lig[256], n_pairs = 0, oligo[32], n_oligo = 0;
    for (int sum = 45; sum >= 40; sum--)
        for (int l = sum - 16; l >= 18; l--) {   /* extension length ascending */
            if (sum - l <= 30 && l <= 30) { ext[n_pairs] = sum - l; lig[n_pairs] = l; n_pairs++; }
        }
    for (int s = 16; s <= 29; s++) oligo[n_oligo++] = s;
    mg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.max_capture = 162; cfg.min_capture = 162; cfg.capture_increment = 5; cfg.max_mip_overlap = 30;
    cfg.n_pairs = n_pairs; cfg.ext_len = ext; cfg.lig_len = lig; cfg.n_oligo_sizes = n_oligo; cfg.oligo_sizes = oligo;

    mg_ctx *ctxs[MAX_CTX];
    int n_ctx = 0;
    for (int a = 3; a < argc && n_ctx < MAX_CTX; a++) {
        if (mg_create(atoi(argv[a]), &ctxs[n_ctx]) != MG_OK) die("mg_create", NULL);
        if (mg_set_config(ctxs[n_ctx], &cfg) != MG_OK) die("mg_set_config", ctxs[n_ctx]);
        if (mg_load_svr_model(ctxs[n_ctx], argv[1]) != MG_OK) die("mg_load_svr_model", ctxs[n_ctx]);
        n_ctx++;
    }

    static mg_region regions[MAX_REGIONS];
    static char *seqs[MAX_REGIONS];
    int n = 0;
    FILE *f = fopen(argv[2], "r");
    if (!f) { perror(argv[2]); return 2; }
    static char line[1 << 16];
    while (n < MAX_REGIONS && fgets(line, sizeof line, f)) {
        int a, b, c, d, at = 0;
        if (sscanf(line, "%d %d %d %d %n", &a, &b, &c, &d, &at) < 4) continue;
        char *s = line + at;
        s[strcspn(s, "\r\n")] = 0;
        seqs[n] = strdup(s);
        memset(&regions[n], 0, sizeof regions[n]);
        regions[n].seq = seqs[n]; regions[n].seq_len = (int)strlen(seqs[n]);
        regions[n].seq_start = a; regions[n].seq_stop = b; regions[n].start_flanked = c; regions[n].stop_flanked = d;
        n++;
    }
    fclose(f);

    /* sizes: pure host arithmetic */
    int64_t grid_off[MAX_REGIONS + 1], scan_off[MAX_REGIONS + 1], pos_off[MAX_REGIONS + 1];
    if (mg_tile_sizes(&cfg, regions, n, grid_off, scan_off, pos_off) != MG_OK) { fprintf(stderr, "mg_tile_sizes failed\n"); return 1; }
    const int64_t total = grid_off[n];
    uint8_t *valid = malloc((size_t)total);
    double *lo = malloc((size_t)total * sizeof(double)), *sv = malloc((size_t)total * sizeof(double));
    int64_t offsets[MAX_REGIONS + 1];
    if (mg_score_regions_multi(ctxs, n_ctx, regions, n, MG_WANT_LOGISTIC | MG_WANT_SVR, offsets, valid, lo, sv) != MG_OK)
        die("mg_score_regions_multi", ctxs[0]);
    long n_valid = 0;
    double s_lo = 0, s_sv = 0;
    for (int64_t i = 0; i < total; i++) {
        n_valid += valid[i];
        if (valid[i]) { s_lo += lo[i]; s_sv += sv[i]; }
    }
    printf("regions %d contexts %d candidates %lld valid %ld sum_logistic %.17g sum_svr %.17g\n", n, n_ctx, (long long)total, n_valid, s_lo, s_sv);

    /* condense + collapse on the devices: best MIP per scan start and strand */
    mg_select_params sp;
    memset(&sp, 0, sizeof sp);
    sp.method = 1; sp.heuristic = 1; sp.lower_score_limit = 1.5; sp.upper_score_limit = 2.2; sp.max_arm_copy = 75; sp.target_arm_copy = 20; sp.masked_arm_threshold = 0.5;
    mg_tile_result res;
    memset(&res, 0, sizeof res);
    res.scan_best = malloc((size_t)scan_off[n] * 2 * sizeof(int64_t));
    res.pos_best = malloc((size_t)pos_off[n] * 2 * sizeof(int64_t));
    res.scan_best_svr = malloc((size_t)scan_off[n] * 2 * sizeof(double));
    if (mg_tile_regions_multi(ctxs, n_ctx, regions, n, MG_WANT_SVR, &sp, 0, &res) != MG_OK) die("mg_tile_regions_multi", ctxs[0]);
    long winners = 0;
    double s_win = 0;
    for (int64_t i = 0; i < scan_off[n] * 2; i++)
        if (res.scan_best[i] >= 0) { winners++; s_win += res.scan_best_svr[i]; }
    printf("scan_start_winners %ld sum_winner_svr %.17g\n", winners, s_win);

    /* opt-in: copy numbers of every arm-sized oligo, counted against the regions' own sequences as the "genome" */
    const char *contigs[MAX_REGIONS];
    int64_t lens[MAX_REGIONS];
    for (int i = 0; i < n; i++) { contigs[i] = seqs[i]; lens[i] = regions[i].seq_len; }
    mg_genome *g = NULL;
    if (mg_genome_create(ctxs[0], contigs, lens, n, &g) != MG_OK) die("mg_genome_create", ctxs[0]);
    int64_t copies_off[MAX_REGIONS + 1];
    if (mg_count_arm_copies(g, regions, n, oligo, n_oligo, NULL, copies_off) != MG_OK) die("mg_count_arm_copies (sizing)", ctxs[0]);
    int32_t *copies = malloc((size_t)(copies_off[n] + 1) * sizeof(int32_t));
    if (mg_count_arm_copies(g, regions, n, oligo, n_oligo, copies, copies_off) != MG_OK) die("mg_count_arm_copies", ctxs[0]);
    long ones = 0, queried = 0;
    for (int64_t i = 0; i < copies_off[n]; i++) { queried += copies[i] != 0; ones += copies[i] == 1; }
    printf("oligos %ld single_copy %ld\n", queried, ones);
    mg_genome_destroy(g);

    for (int i = 0; i < n_ctx; i++) mg_destroy(ctxs[i]);
    return 0;
}
