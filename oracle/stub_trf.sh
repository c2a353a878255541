#!/bin/bash
# Stand-in for Tandem Repeats Finder so that `-trf <this script>` exercises the masked-arm code paths
# (mipgen.cpp:1045-1084, 606-610, 626-633, 1629, 1701) on a box without TRF.  Called as
#   trf <project>.feature_sequences.fa 2 7 7 80 10 14 100 -m -h
# it writes <basename>.2.7.7.80.10.14.100.mask into the working directory: the same records with every base
# whose chromosome coordinate x satisfies x % 211 < 17 replaced by 'N' (mirrored by tests/stub_rules.py).
# Record headers are ">chr:start-stop" (mipgen.cpp:1222).  Without arguments it exits 255, which is what mipgen
# probes for at start-up (mipgen.cpp:152-160).
[ $# -eq 0 ] && exit 255
in="$1"
out="$(basename "$in").2.7.7.80.10.14.100.mask"
awk '
  /^>/ { print; split(substr($0, 2), a, ":"); split(a[2], b, "-"); pos = b[1] + 0; next }
  { s = ""
    for (i = 1; i <= length($0); i++) { c = substr($0, i, 1); if (pos % 211 < 17) c = "N"; s = s c; pos++ }
    print s }' "$in" > "$out"
exit 0
