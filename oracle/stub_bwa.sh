#!/bin/bash
# Stand-in for `bwa` so the MIPgen CLI (reference and drop-in alike) runs on a box
# without BWA.  mipgen requires `bwa` with no arguments to exit 1 (mipgen.cpp:146-151);
# `aln` output is ignored; `samse <index> <sai> <fq>` must print one SAM record per
# read carrying X0:i:1 / X1:i:0 so every capture site maps uniquely and every arm has
# copy number 1 (mipgen.cpp:581-587, 857).
[ $# -eq 0 ] && exit 1
case "$1" in
  aln) exit 0 ;;
  samse)
    awk 'NR%4==1{n=substr($0,2)} NR%4==2{printf "%s\t0\tchr1\t1\t37\t%dM\t*\t0\t0\t%s\t*\tXT:A:U\tNM:i:0\tX0:i:1\tX1:i:0\n", n, length($0), $0}' "$4"
    exit 0 ;;
esac
exit 1
