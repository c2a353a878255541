#!/usr/bin/env python3
"""K-svr forms side by side on the bench panel (60 regions, 2.53 M candidates, 2048 SV): factored FP64 (default), dense FP64
DMMA, tensor-core split-FP16 (tcgen05).  Prints ms per pass, candidates/s, max / median relative deviation from the factored
FP64 scores and how many condense / collapse winners change.   python tools/bench_modes.py [steps]"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    bench.build_model(ctx, cfg, tempfile.mkdtemp())
    _g, regions = bench.make_panel(cfg, bench.N_REGIONS, bench.GENOME_SEED)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    out = {"candidates": pnl.n_candidates}
    ref = ref_sel = None
    for name, mode in (("factored_fp64", 0), ("tensor_core_fp16x2", 3), ("dense_fp64", 1)):
        ctx.set_svr_mode(mode)
        for _ in range(2):
            pnl.score(mg.MG_WANT_SVR)
        ctx.sync()
        ctx.reset_timings()
        ctx.timer_start()
        for _ in range(steps):
            pnl.score(mg.MG_WANT_SVR)
        ms = ctx.timer_stop() / steps
        t = ctx.timings()
        valid, _l, sv, _ = pnl.fetch(valid=True, svr=True)
        sel = pnl.select(regions, 1, 1.5, 2.2)
        ok = valid.astype(bool)
        row = {"ms_per_pass": ms, "candidates_per_s": pnl.n_candidates / ms * 1e3, "k_feat_ms": t.ms_feat / steps, "k_svr_ms": t.ms_svr / steps}
        if ref is None:
            ref, ref_sel = sv, sel
        else:
            err = np.abs(sv[ok] - ref[ok]) / np.abs(ref[ok])
            row.update(max_rel_vs_factored=float(err.max()), median_rel_vs_factored=float(np.median(err)),
                       scan_winners_changed=int((sel[1] != ref_sel[1]).sum()), pos_winners_changed=int((sel[3] != ref_sel[3]).sum()),
                       threshold_flips=int(((sv[ok] > 2.2) != (ref[ok] > 2.2)).sum() + ((sv[ok] > 1.5) != (ref[ok] > 1.5)).sum()))
        out[name] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
