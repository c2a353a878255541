// MinusSVMipv4.h -- drop-in replacement (reference: MinusSVMipv4.h:5-13).  Minus-strand
// candidate: arms swapped, every stored sequence is the reverse complement of the genomic
// window.  Relies on SVMipv4.h having been included first (mipgen.cpp:23-25).
#include <string>
using namespace std;

class MinusSVMipv4 : public SVMipv4
{
  public:
    MinusSVMipv4(string chromosome, int scan_start, int scan_stop, int ext_length, int lig_length);
    void set_ext_probe_seq(std::string);
    void set_lig_probe_seq(std::string);
    int get_mip_start();
    void set_scan_target_seq(string seq);
};
