// MinusSVMipv4.h -- forwarding stub, see PlusSVMipv4.h (included by mipgen.cpp:25).
#include "SVMipv4.h"
