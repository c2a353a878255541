#!/usr/bin/env python3
"""Scale check (not a pytest): a few thousand regions through the region-grid path -- many workspace
chunks, > 2^31 feature elements -- timed, and spot-checked against the explicit-candidate path.
    python tools/scale_check.py [n_regions]
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from test_gpu_fullsize import decode  # noqa: E402


def revcomp(s: bytes) -> bytes:
    return s[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


def main():
    n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    bench.build_model(ctx, cfg, tempfile.mkdtemp())
    glen = panel.genome_length_for(n_regions, 200, cfg)
    genome = panel.lcg_genome(glen, 99)
    regions = panel.make_regions(genome, n_regions, 100, 200, cfg, 100)
    t0 = time.time()
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    t_lrc = time.time() - t0
    pnl = ctx.panel(regions)
    n = pnl.n_candidates
    for want, name in ((mg.MG_WANT_SVR, "svr"), (mg.MG_WANT_LOGISTIC, "logistic")):
        pnl.score(want)
        ctx.sync()
        ctx.timer_start()
        pnl.score(want)
        ms = ctx.timer_stop()
        print("%s: %d regions, %d candidates, %.1f ms -> %.3e cand/s" % (name, n_regions, n, ms, n / ms * 1e3), flush=True)
    valid, lo, sv, _ = pnl.fetch(valid=True, logistic=True, svr=True)
    offs = pnl.offsets
    rng = np.random.default_rng(5)
    picks = np.sort(np.concatenate([rng.choice(n, 300, replace=False), np.arange(n - 50, n), np.arange(50)]))
    picks = picks[valid[picks].astype(bool)]
    cands, lrc = [], []
    for g in picks:
        ri = int(np.searchsorted(offs, g, side="right") - 1)
        r = regions[ri]
        s, t, e, l, strand = decode(cfg, r, int(g - offs[ri]))
        o = r.seq_start
        if strand == 0:
            c = dict(ext=r.seq[s - e - o:s - o], lig=r.seq[t + 1 - o:t + 1 + l - o], tgt=r.seq[s - o:t + 1 - o])
        else:
            c = dict(ext=revcomp(r.seq[t + 1 - o:t + 1 + e - o]), lig=revcomp(r.seq[s - l - o:s - o]), tgt=revcomp(r.seq[s - o:t + 1 - o]))
        cands.append(c)
        lrc.append(r.lrc)
    lo2, sv2, _ = ctx.score_candidates(cands, np.array(lrc), mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    err = float(np.max(np.abs(sv[picks] - sv2) / np.abs(sv2)))
    print("lrc time %.1f s; valid %.4f; logistic bit-equal %s; svr max rel diff vs explicit path %.2e" %
          (t_lrc, valid.mean(), np.array_equal(lo[picks], lo2), err))
    assert np.array_equal(lo[picks], lo2) and err < 1e-11 and valid.mean() > 0.99
    print("SCALE CHECK OK")


if __name__ == "__main__":
    main()
