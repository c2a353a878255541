// Featurev5.h -- drop-in replacement for the reference header of the same name.
//
// One target region.  The unchanged mipgen.cpp reads and writes the public members below directly
// (mipgen.cpp:414-425, 1023-1030, 1118-1119, 1214-1225), so their names and types are the contract;
// what differs from the reference class is behaviour:
//   * get_long_range_content() runs on the GPU (mg_long_range_content, K-lrc);
//   * every live object registers itself with the scoring shim (constructors / destructor below),
//     which is how a candidate's get_score() / get_parameters() finds the region it was cut from --
//     the reference passes no region pointer to the scoring classes.
#ifndef MIPGEN_B200_DROPIN_FEATUREV5_H
#define MIPGEN_B200_DROPIN_FEATUREV5_H
#include <string>
#define MER_NUM 44

using namespace std;  // the reference header injects this and mipgen.cpp relies on it

class Featurev5
{
  public:
    string chr, label;
    // 1-based inclusive coordinates of the target and of the target +- flank
    int start_position, stop_position, flank_size, start_position_flanked, stop_position_flanked;
    // tiling state owned by mipgen.cpp
    int mip_count, current_scan_start_position;
    // sequence window [start_flanked - max_capture, stop_flanked + max_capture + 15] and its TRF-masked copy
    string chromosomal_sequence, masked_chromosomal_sequence;
    int chromosomal_sequence_start_position, chromosomal_sequence_stop_position;
    // features 23..66 of every candidate of this region
    double long_range_content[MER_NUM];

    Featurev5();
    Featurev5(string chromosome, int start, int stop, int f, string l);
    Featurev5(const Featurev5 &other);             // registration needs explicit copy semantics
    Featurev5 &operator=(const Featurev5 &other);
    ~Featurev5();

    bool operator<(Featurev5 &b);
    void get_long_range_content(string extended_sequence, string feature_mers[]);
};
#endif
