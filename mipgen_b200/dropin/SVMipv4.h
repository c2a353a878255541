// SVMipv4.h -- drop-in replacement for the reference header of the same name.
//
// The candidate-MIP classes.  Public members and methods are the ones the unchanged mipgen.cpp touches
// directly (design_mip 599-762, print_details 765-794, condense/collapse/pick 1506-1939); the two scoring
// methods do no arithmetic on the host:
//   get_score()       -> logistic score computed by K-feat's fused epilogue on the GPU
//   get_parameters()  -> the 192 feature doubles computed by K-feat on the GPU
// both served from a per-region device batch the shim launches lazily (mipgen_dropin.cpp).
// PlusSVMipv4 / MinusSVMipv4 are declared here as well; the headers of those names only forward to
// this file (mipgen.cpp:22-25 includes all three, this one first).
#ifndef MIPGEN_B200_DROPIN_SVMIPV4_H
#define MIPGEN_B200_DROPIN_SVMIPV4_H
#include <map>
#include <string>
#include <vector>
using namespace std;

class SVMipv4
{
  public:
    static map<string, double> junction_scores;  // DEFINED by the caller (mipgen.cpp:33); unused on the host

    // geometry, set by the strand-specific constructors (1-based inclusive chromosome coordinates)
    string chr, strand;
    int scan_start_position, scan_stop_position, scan_size;
    int extension_arm_length, ligation_arm_length;
    int ext_probe_start, ext_probe_stop, lig_probe_start, lig_probe_stop;

    // sequences in probe orientation (reverse-complemented on '-')
    string ext_probe_sequence, lig_probe_sequence, scan_target_sequence, mip_seq, ligation_junction;
    string ext_masked_sequence, lig_masked_sequence;

    // filled in by design_mip / the selection code
    int ext_probe_copy, lig_probe_copy, snp_count;
    double arm_fraction_masked, score;
    char translocation_failed, snp_failed, mapping_failed, masking_failed;
    bool has_snp_mip;
    vector<int> snp_positions;
    string snp_ext_sequence, snp_lig_sequence, snp_mip_sequence;

    // batched driver only (mipgen_b200/batched): the SVR score the device computed for this candidate, handed to the
    // caller's predict_value by get_parameters (mixed mode re-scores picked MIPs, mipgen.cpp:1523-1550, 1873-1877)
    bool b200_has_svr;
    double b200_svr;

    SVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length);
    virtual ~SVMipv4() {}
    virtual int get_mip_start() = 0;
    virtual void set_ext_probe_seq(string seq) = 0;
    virtual void set_lig_probe_seq(string seq) = 0;

    static void set_junction_scores();
    double get_score();
    void get_parameters(vector<double> &parameters, double long_range_content[]);
};

// '+' strand: extension arm upstream of the scan window, ligation arm downstream
class PlusSVMipv4 : public SVMipv4
{
  public:
    PlusSVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length);
    int get_mip_start();
    void set_ext_probe_seq(string seq);
    void set_lig_probe_seq(string seq);
    void set_scan_target_seq(string seq);  // non-virtual in the reference too (mipgen.cpp:461)
};

// '-' strand: arms swapped, every stored sequence is the reverse complement of the genomic window
class MinusSVMipv4 : public SVMipv4
{
  public:
    MinusSVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length);
    int get_mip_start();
    void set_ext_probe_seq(string seq);
    void set_lig_probe_seq(string seq);
    void set_scan_target_seq(string seq);  // (mipgen.cpp:462)
};
#endif
