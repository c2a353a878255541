"""TEST INFRASTRUCTURE -- CPU restatement of SURVEY.md 8(f4): the copy number find_copy stores for an arm-sized oligo.

The reference (mipgen.cpp:558-596) runs `bwa aln` + `bwa samse` on one read per (oligo size, start) of every region
(reads written at mipgen.cpp:824-836: starts 0 .. length - size - 1) and keeps each read's X0 tag, BWA's number of best hits;
a read without X0 gets 100.  For a read cut out of the indexed genome the best hits are its exact occurrences on either strand,
so this restatement counts exact occurrences by brute force (a dictionary of every k-mer of the genome).

PARITY UNPINNED against BWA itself: BWA (pinned by the reference's README to 0.6+/0.7) is not in this image and the reference
holds no golden vector for this path.  What is pinned: (1) the parse rule of find_copy (X0 -> copy, no tag -> 100) through the
reference CLI run against the rule-driven stub bwa (tests/test_selection_pinning.py); (2) this restatement against hand-counted
known answers (tests/test_copy_count.py).  Only tests/ may import this file.
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Sequence

import numpy as np

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _clean(seq: bytes) -> bytes:
    return seq.upper()


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def kmer_table(contigs: Sequence[bytes], size: int) -> Counter:
    """Occurrences of every pure-ACGT k-mer of the contigs (forward strand; case-insensitive)."""
    tab: Counter = Counter()
    ok = set(b"ACGT")
    for c in contigs:
        c = _clean(c)
        bad = np.frombuffer(c, np.uint8)
        bad = ~np.isin(bad, list(ok))
        badpf = np.concatenate([[0], np.cumsum(bad)])
        for i in range(0, len(c) - size + 1):
            if badpf[i + size] - badpf[i] == 0:
                tab[c[i:i + size]] += 1
    return tab


def count_arm_copies(contigs: Sequence[bytes], seq: bytes, sizes: Sequence[int], tables: Dict[int, Counter] = None) -> np.ndarray:
    """[len(sizes)][len(seq)] int32 in the layout of mg_region.copies for one region's chromosomal_sequence."""
    seq = _clean(seq)
    out = np.zeros((len(sizes), len(seq)), np.int32)
    ok = set(b"ACGT")
    for k, size in enumerate(sizes):
        tab = tables[size] if tables is not None and size in tables else kmer_table(contigs, size)
        for i in range(0, len(seq) - size):   # mipgen.cpp:829: relative_start_position < length - oligo_size
            q = seq[i:i + size]
            if any(ch not in ok for ch in q):
                out[k, i] = 100                # BWA's answer is not an exact-match count: treated like a read without X0
                continue
            n = tab.get(q, 0) + tab.get(revcomp(q), 0)
            out[k, i] = 100 if n == 0 else min(n, 1000000)
    return out
