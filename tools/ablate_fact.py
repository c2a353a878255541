#!/usr/bin/env python3
"""Where does the factored K-svr spend its time?  Runs the bench panel through library variants built with
-DMG_FACT_ABLATE=<mask> (parts of the kernel removed; results are wrong, only the timing is meaningful):

    for v in 0 1 2 4 8 15; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared \
        -Xcompiler -fPIC -DMG_FACT_ABLATE=$v -o tools/_bin/libmg_ablate_$v.so mipgen_b200/csrc/*.cu; done
    gpurun -- python tools/ablate_fact.py

mask bits: 1 no exp, 2 no insert-block DMMA, 4 no gather arithmetic (and loads), 8 no arm DMMA.
The additive picture these timings give, and the restructurings they led to, are in DESIGN.md section 6.
"""
import glob
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib):
    import mipgen_b200._capi as capi
    capi.LIB_PATH = lib
    import bench
    import mipgen_b200 as mg
    from mipgen_b200 import panel
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    bench.build_model(ctx, cfg, tempfile.mkdtemp())
    _g, regions = bench.make_panel(cfg, 60, bench.GENOME_SEED)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_SVR)
    ctx.sync()
    ctx.reset_timings()
    for _ in range(3):
        pnl.score(mg.MG_WANT_SVR)
    ctx.sync()
    t = ctx.timings()
    print("%-28s k_svr %.3f ms/pass  k_feat %.3f ms/pass" % (os.path.basename(lib), t.ms_svr / 3, t.ms_feat / 3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_bin", "libmg_ablate_*.so")),
                      key=lambda p: int(p.rsplit("_", 1)[1].split(".")[0]))
        for lib in libs:
            subprocess.run([sys.executable, os.path.abspath(__file__), lib], check=False)
