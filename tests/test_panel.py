"""Synthetic-input generators."""
import numpy as np

from mipgen_b200 import panel


def test_lcg_genome_matches_scalar_recurrence():
    n, seed = 50000, 12345
    s, out = seed, bytearray()
    for _ in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out.append(b"ACGT"[(s >> 33) & 3])
    assert panel.lcg_genome(n, seed) == bytes(out)


def test_default_arm_pairs_are_the_references_57():
    e, l = panel.default_arm_pairs()
    assert len(e) == 57 and min(e) == 16 and max(e) == 27 and min(l) == 18 and max(l) == 29
    sums = [a + b for a, b in zip(e, l)]
    assert sums == sorted(sums, reverse=True)
    for s in set(sums):
        es = [a for a, t in zip(e, sums) if t == s]
        assert es == sorted(es)


def test_cut_region_follows_genome_dir_slicing():
    cfg = panel.Config()
    g = panel.lcg_genome(10000, 1)
    r = panel.cut_region(g, 4000, 4100, cfg)
    assert r.seq_start == 4000 - 162 and r.seq_stop == 4100 + 162 + 15 and len(r.seq) == r.seq_stop - r.seq_start + 1
    assert len(r.flank_seq) == len(r.seq) + 2000
    r2 = panel.cut_region(g, 50, 90, cfg)
    assert r2.seq_start == 1 and r2.flank_seq is None
