// Featurev5.h -- drop-in replacement for the reference header of the same name.
//
// Same public surface as /root/reference/Featurev5.h:7-28 (mipgen.cpp reads and writes
// these members directly: 414-425, 1023-1030, 1118-1119, 1214-1225), but
//   * get_long_range_content() runs on the GPU (mg_long_range_content, K-lrc), and
//   * every live object registers itself with the scoring shim, which is how a
//     candidate's get_score()/get_parameters() finds the region it was cut from
//     (the reference passes no region pointer to the scoring classes).
#ifndef MIPGEN_B200_DROPIN_FEATUREV5_H
#define MIPGEN_B200_DROPIN_FEATUREV5_H
#include <string>
#define MER_NUM 44

using namespace std;  // the reference header injects this; mipgen.cpp relies on it

class Featurev5
{
  public:
    // --- identity and coordinates (1-based, inclusive) ---
    string chr;
    string label;
    int start_position;
    int stop_position;
    int flank_size;
    int start_position_flanked;
    int stop_position_flanked;
    // --- tiling state owned by mipgen.cpp ---
    int mip_count;
    int current_scan_start_position;
    // --- sequence window: [start_flanked - max_capture, stop_flanked + max_capture + 15] ---
    string chromosomal_sequence;
    string masked_chromosomal_sequence;
    int chromosomal_sequence_start_position;
    int chromosomal_sequence_stop_position;
    // --- features 23..66 of every candidate of this region ---
    double long_range_content[MER_NUM];

    Featurev5(string chromosome, int start, int stop, int f, string l);
    Featurev5();
    Featurev5(const Featurev5 &other);
    Featurev5 &operator=(const Featurev5 &other);
    ~Featurev5();

    void get_long_range_content(string extended_sequence, string feature_mers[]);
    bool operator<(Featurev5 &b);
};
#endif
