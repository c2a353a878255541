#!/usr/bin/env python3
"""Small, fixed workload for ncu: one panel (default 8 regions, ~3e5 candidates), 2048-SV model,
a couple of SVR passes and one logistic pass.  Usage (under gpurun):
    ncu --set full --clock-control none --import-source on -k regex:k_svr -s 1 -c 1 -o gpurun_out/prof python tools/profile_step.py
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402


def main():
    n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # mg_set_svr_mode: 0 factored, 1 dense, 3 tensor cores
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    bench.build_model(ctx, cfg, tempfile.mkdtemp())
    _g, regions = bench.make_panel(cfg, n_regions, bench.GENOME_SEED)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    ctx.set_svr_mode(mode)
    for _ in range(passes):
        pnl.score(mg.MG_WANT_SVR)
    pnl.score(mg.MG_WANT_LOGISTIC)
    ctx.sync()
    t = ctx.timings()
    print("candidates", pnl.n_candidates, "ms_feat", t.ms_feat, "ms_svr", t.ms_svr)


if __name__ == "__main__":
    main()
