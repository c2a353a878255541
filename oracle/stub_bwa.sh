#!/bin/bash
# Stand-in for `bwa` so the MIPgen CLI (reference, drop-in and batched alike) runs on a box
# without BWA.  mipgen requires `bwa` with no arguments to exit 1 (mipgen.cpp:146-151);
# `aln` output is ignored; `samse <index> <sai> <fq>` must print one SAM record per read.
#
# Default: every read carries X0:i:1 / X1:i:0, so every capture site maps uniquely and every
# arm has copy number 1 (mipgen.cpp:581-587, 857).
# MIPGEN_STUB_RULES=1 switches on deterministic, coordinate-keyed exceptions (mirrored by
# tests/stub_rules.py) so that copy tables and unmappable MIP starts are exercised:
#   arm reads      "chr<c>:<start>-<stop>"   k = (31*start + stop - start) % 97
#                                            X0 = 101 (k==0), 30 (k==1), 3 (k<5), 2 (k<9), else 1
#   capture reads  "<size>_<c>_<start>"      X0 = 2 when (start + size) % 53 == 0 (ambiguous site)
# MIPGEN_STUB_RULES=2: the arm rule only (no ambiguous sites: design_mip leaves masking_failed uninitialised after a
# mapping failure, mipgen.cpp:622-624, so all_mips.txt of such a run is not reproducible even by the reference itself)
# MIPGEN_STUB_RULES=3: arm reads take X0 from the two-column file MIPGEN_STUB_X0_FILE (read name, X0; absent names get 1) -- used to
# replay exact-match copy counts (tests/test_copy_count.py) through the reference's own find_copy
[ $# -eq 0 ] && exit 1
case "$1" in
  aln) exit 0 ;;
  samse)
    awk -v rules="${MIPGEN_STUB_RULES:-0}" -v x0file="${MIPGEN_STUB_X0_FILE:-}" '
      BEGIN { if (rules == 3 && x0file != "") while ((getline line < x0file) > 0) { split(line, f, "\t"); tab[f[1]] = f[2] } }
      NR%4==1 { n = substr($0, 2) }
      NR%4==2 {
        x0 = 1
        if (rules == 3 && (n in tab)) x0 = tab[n] + 0
        if (rules == 1 || rules == 2) {
          if (n ~ /^chr/) {
            split(n, a, ":"); split(a[2], b, "-"); start = b[1] + 0; stop = b[2] + 0
            k = (31 * start + stop - start) % 97
            x0 = (k == 0) ? 101 : (k == 1) ? 30 : (k < 5) ? 3 : (k < 9) ? 2 : 1
          } else if (rules == 1) {
            m = split(n, a, "_"); size = a[1] + 0; start = a[m] + 0
            if ((start + size) % 53 == 0) x0 = 2
          }
        }
        printf "%s\t0\tchr1\t1\t37\t%dM\t*\t0\t0\t%s\t*\tXT:A:U\tNM:i:0\tX0:i:%d\tX1:i:0\n", n, length($0), $0, x0
      }' "$4"
    exit 0 ;;
esac
exit 1
