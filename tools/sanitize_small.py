#!/usr/bin/env python3
"""A small tour of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Two configs through one context (narrow -> wide arm table), SVR in all three forms, logistic, features, select with selection inputs,
device-written records and all_mips.txt, FASTQ writers, mg_tile_regions, and the opt-in exact-match copy counting (genome index with
N runs, lower case and a contig shorter than 32 bases)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from oracle_api import Oracle  # noqa: E402
from helpers import small_config, synthetic_regions, random_model, tmpdir  # noqa: E402


def main():
    oracle = Oracle()
    ctx = mg.Context(0)
    d = tmpdir()
    for cfg in (small_config((45,), 162, 152, 5), small_config((40, 43, 45), 162, 157, 5)):
        _g, regions = synthetic_regions(oracle, cfg, 2, 40, 70, 5)
        ctx.set_config(cfg)
        ctx.load_svr_model(random_model(oracle, cfg, 80, 3, os.path.join(d, "m%d.model" % len(cfg.ext_len))))
        rng = np.random.default_rng(1)
        regions[0].copies = rng.choice([0, 1, 1, 2, 101], size=(len(cfg.oligo_sizes), len(regions[0].seq))).astype(np.int32)
        for mode in (0, 1, 3):
            if mode == 3 and not ctx.svr_tensor_core_available():
                continue
            ctx.set_svr_mode(mode)
            ctx.score_regions(regions, mg.MG_WANT_SVR)
        ctx.set_svr_mode(0)
        ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
        pnl = ctx.panel(regions)
        pnl.score(mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
        pnl.select(regions, 1, 1.5, 2.2)
        text, per = pnl.format_enumerated(regions, 0, 0.9)
        assert per.sum() > 0 and len(text) > 0
        pnl.close()
        t = mg.tile_regions(ctx, regions, mg.MG_WANT_SVR, select=dict(method=1, lower=1.5, upper=2.2))
        assert (t.scan_best >= -1).all()
        ctx.fastq(regions, "1", oligo=False)
        ctx.fastq(regions, "1", oligo=True)
    g = bytearray(panel.lcg_genome(20000, 3))
    g[5000:5040] = b"N" * 40
    g[9000:9100] = bytes(g[9000:9100]).lower()
    gen = ctx.genome([bytes(g).decode(), "ACGTTGCAAGGCTTAACCGGTTAA", ""])
    regions[0].copies = None
    tabs = gen.count_arm_copies(regions + [panel.Region(10, 20, 1, 30, bytes(g[4990:5020]))], [16, 24, 32, 1])
    assert all(t.min() >= 0 for t in tabs)
    gen.close()
    ctx.close()
    print("sanitize tour done")


if __name__ == "__main__":
    main()
