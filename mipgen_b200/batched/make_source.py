#!/usr/bin/env python3
"""Build-time source transformation for the batched MIPgen driver (INTEGRATION.md route C).

    python make_source.py /root/reference/mipgen.cpp _build/mipgen_batched.cpp

Reads the reference's mipgen.cpp WHERE IT LIES and writes a patched copy into the (git-ignored) build directory; nothing
of the reference is stored in this repository.  Five anchored edits, each checked to match exactly once:

  1. `#include "mipgen_batched.h"` before `class mipgen{`, `#include "batched_members.inc"` right after its `public:`;
  2. tile_regions (mipgen.cpp:403-556): the statements from the initialisation of current_scan_start_position (:421)
     through `collapse_mips();` (:505) become `b200_tile_feature(feature);`;
  3. predict_value (mipgen.cpp:1948): first statement returns the device score parked by get_parameters, if any;
  4. check_copy_numbers (mipgen.cpp:796-873): the loops that print all_sequences.fq / oligo_copy_count.fq (:804-838) become
     `b200_write_fastqs(BWAFQ, ARMSFQ);` (the bwa calls and the SAM parsing that follow are untouched);
  5. find_copy (mipgen.cpp:558-596): first statement `if (b200_find_copy()) return;` -- a no-op unless the user opts in to exact-match
     arm copy counting on the device with MIPGEN_B200_EXACT_COPIES.
"""
import re
import sys


def one(pattern, text, what, flags=0):
    m = list(re.finditer(pattern, text, flags))
    if len(m) != 1:
        sys.exit("make_source.py: anchor for %s matched %d times (reference changed?)" % (what, len(m)))
    return m[0]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    t = open(src, encoding="latin-1", newline="").read()
    # 1. includes
    m = one(r"class\s+mipgen\s*\{\s*\n\s*public:\s*\n", t, "class mipgen")
    t = t[:m.start()] + '#include "mipgen_batched.h"\n' + t[m.start():m.end()] + '#include "batched_members.inc"\n' + t[m.end():]
    # 2. the candidate loop nest + condense + collapse of one feature
    a = one(r"^[ \t]*feature->current_scan_start_position\s*=\s*feature->start_position_flanked\s*-\s*max_capture_size[^\n]*\n", t,
            "start of the tile loop", re.M)
    b = one(r"^[ \t]*collapse_mips\(\);[^\n]*\n", t, "collapse_mips() call", re.M)
    if not (a.start() < b.start()):
        sys.exit("make_source.py: tile loop anchors out of order")
    t = t[:a.start()] + "\t\tb200_tile_feature(feature); // mipgen_b200: loop nest + condense_mips + collapse_mips on the GPU\n" + t[b.end():]
    # 3. predict_value
    m = one(r"double\s+predict_value\s*\(\s*vector<double>\s*&\s*parameters\s*,\s*svm_model\s*\*\s*model\s*\)\s*\{", t, "predict_value")
    t = t[:m.end()] + "\n\t{ double b200_score; if (mipgen_b200_take_pending_svr(&b200_score)) return b200_score; } // mipgen_b200\n" + t[m.end():]
    # 4. check_copy_numbers (mipgen.cpp:796-873): the loops that print the two FASTQ files for BWA (:804-838)
    a = one(r"^[ \t]*Featurev5\s*\*\s*feature;\s*for\s*\(list<Featurev5>::iterator it = features_to_scan\.begin\(\); it != features_to_scan\.end\(\); it\+\+\)\s*\{\s*"
            r"feature = &\*it;\s*string chr = feature->chr;\s*for \(int capture_size = max_capture_size", t, "FASTQ loops of check_copy_numbers", re.M)
    b = one(r"^[ \t]*BWAFQ\.close\(\);", t, "BWAFQ.close()", re.M)
    if not (a.start() < b.start()):
        sys.exit("make_source.py: FASTQ anchors out of order")
    t = t[:a.start()] + "\tb200_write_fastqs(BWAFQ, ARMSFQ); // mipgen_b200: both FASTQ files formatted on the GPU\n" + t[b.start():]
    # 5. find_copy: opt-in replacement of the bwa run on oligo_copy_count.fq
    m = one(r"void\s+find_copy\s*\(\s*\)\s*\{", t, "find_copy")
    t = t[:m.end()] + "\n\tif (b200_find_copy()) return; // mipgen_b200: exact-match copy counting on the GPU when MIPGEN_B200_EXACT_COPIES is set\n" + t[m.end():]
    open(dst, "w", encoding="latin-1", newline="").write(t)


if __name__ == "__main__":
    main()
