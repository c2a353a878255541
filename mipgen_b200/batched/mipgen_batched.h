// mipgen_batched.h -- file-scope declarations for the batched MIPgen driver (INTEGRATION.md route C).
//
// The batched driver is the reference's own mipgen.cpp with five anchored edits applied at BUILD time by
// make_source.py (no reference code lives in this repository):
//   1. this header is included before `class mipgen`, and batched_members.inc inside it;
//   2. in tile_regions (mipgen.cpp:403-556) the per-feature candidate loop nest + condense_mips + collapse_mips
//      (mipgen.cpp:421-505) is replaced by one call, b200_tile_feature(feature), which takes the winners of a whole
//      batch of features from mg_tile_regions_multi and materialises SVMipv4 objects only for them;
//   3. predict_value (mipgen.cpp:1948-2019) returns the device's SVR score of the object get_parameters was just
//      called on, instead of printing and re-parsing its 192 features;
//   4. check_copy_numbers' FASTQ loops (mipgen.cpp:804-838) become one device call, b200_write_fastqs;
//   5. find_copy (mipgen.cpp:558-596) starts with `if (b200_find_copy()) return;`: a no-op unless MIPGEN_B200_EXACT_COPIES switches the
//      opt-in exact-match arm copy counting on.
// Everything else -- flag parsing, BED / FASTA / BWA / TRF / tabix handling, design_mip, pick_mips and its helpers,
// print_details and the output files -- is the reference's code, compiled unchanged.
#ifndef MIPGEN_B200_BATCHED_H
#define MIPGEN_B200_BATCHED_H
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mipgen_b200.h"

// implemented in dropin/mipgen_dropin.cpp
mg_ctx *mipgen_b200_shim_context();                 // the context svm_load_model / get_score use (device MIPGEN_B200_DEVICE)
bool mipgen_b200_take_pending_svr(double *score);   // SVR score parked by get_parameters on an object that carries one
[[noreturn]] void mipgen_b200_fatal(const char *what, mg_ctx *ctx);

// per-batch storage the member functions in batched_members.inc keep between calls
struct mipgen_b200_batch {
    bool ready = false;
    std::vector<mg_ctx *> ctxs;              // [0] is the shim's context
    std::vector<int> ext_len, lig_len, oligo;
    mg_config cfg;
    size_t first = 0, last = 0;              // features [first, last) of the run are covered by the arrays below
    std::vector<mg_region> regions;
    std::vector<std::vector<int>> copies;    // backing stores of the regions' optional inputs
    std::vector<std::string> masked;
    std::vector<std::vector<uint8_t>> snp, unmappable;
    std::vector<int64_t> grid_off, scan_off, pos_off, scan_best, pos_best;
    std::vector<double> sb_logistic, sb_svr, logistic, svr;
    std::vector<uint8_t> valid;
    bool device_records = false;             // this batch's all_mips.txt records were written on the device:
    std::vector<char> text;                  //   the text of all its features, in order (flushed with the batch's first feature)
    std::vector<int64_t> records_per_region; //   and how many records each feature contributed
    long n_batches = 0, n_objects = 0;
    double t_first_tile = 0;                 // ... and when the tile phase began
    double t_constructed = 0;                // CLOCK_MONOTONIC when `class mipgen` was constructed (start of main)
    double device_seconds = 0, setup_seconds = 0, prep_seconds = 0, objects_seconds = 0, records_seconds = 0;  // MIPGEN_B200_VERBOSE report
};
#endif
