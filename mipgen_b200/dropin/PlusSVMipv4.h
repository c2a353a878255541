// PlusSVMipv4.h -- forwarding stub.  The drop-in declares PlusSVMipv4 next to its base class in
// SVMipv4.h (which mipgen.cpp:23 has already included); this file only has to exist under the name
// mipgen.cpp:24 includes.
#include "SVMipv4.h"
