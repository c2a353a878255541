#!/bin/bash
# Stand-in for tabix so that `-snp_file <plain VCF>` exercises the SNP code paths (mipgen.cpp:875-978, 634-760)
# on a box without tabix: no arguments -> exit 1 (mipgen.cpp:919-924 requires it); `tabix <file> <regions...>`
# prints the whole file (the regions always cover every target).
[ $# -eq 0 ] && exit 1
cat "$1"
