// SVMipv4.h -- drop-in replacement for the reference header of the same name.
//
// One candidate MIP.  Public members and methods are those of
// /root/reference/SVMipv4.h:6-62 because the unchanged mipgen.cpp touches them directly
// (design_mip 599-762, print_details 765-794, condense/collapse/pick 1506-1939).
// The two scoring methods do no arithmetic on the host:
//   get_score()       -> logistic score computed by K-feat's fused epilogue on the GPU
//   get_parameters()  -> the 192 feature doubles computed by K-feat on the GPU
// served from a per-region batch the shim launches lazily (mipgen_dropin.cpp).
// Like the reference header this one has no include guard and is included once, before
// PlusSVMipv4.h / MinusSVMipv4.h (mipgen.cpp:22-25).
#include <string>
#include <vector>
#include <map>
using namespace std;

class SVMipv4
{
  public:
    // dinucleotide -> junction score table; DEFINED by the caller (mipgen.cpp:33)
    static map<string, double> junction_scores;

    // geometry (set by the Plus/Minus constructors)
    string chr;
    string strand;
    int scan_start_position;
    int scan_stop_position;
    int scan_size;
    int extension_arm_length;
    int ligation_arm_length;
    int ext_probe_start;
    int ext_probe_stop;
    int lig_probe_start;
    int lig_probe_stop;

    // sequences, stored in probe orientation (reverse-complemented on '-')
    string ext_probe_sequence;
    string ext_masked_sequence;
    string lig_probe_sequence;
    string lig_masked_sequence;
    string scan_target_sequence;
    string mip_seq;
    string ligation_junction;

    // design_mip results
    int ext_probe_copy;
    int lig_probe_copy;
    double arm_fraction_masked;
    char translocation_failed;
    char snp_failed;
    char mapping_failed;
    char masking_failed;
    int snp_count;
    vector<int> snp_positions;
    bool has_snp_mip;
    string snp_ext_sequence;
    string snp_lig_sequence;
    string snp_mip_sequence;

    double score;

    SVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length);
    virtual ~SVMipv4() {}
    virtual void set_ext_probe_seq(string seq) = 0;
    virtual void set_lig_probe_seq(string seq) = 0;
    virtual int get_mip_start() = 0;

    void get_parameters(vector<double> &parameters, double long_range_content[]);
    double get_score();
    static void set_junction_scores();
};
