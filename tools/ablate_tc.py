#!/usr/bin/env python3
"""Where does the tensor-core K-svr (k_svr_tc.cu) spend its time?  Runs the bench panel through library variants built with
-DMG_TC_ABLATE=<mask> (parts of the kernel removed; results are wrong, only the timing is meaningful):

    for v in 0 1 2 4 8 16; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared \
        -Xcompiler -fPIC -DMG_TC_ABLATE=$v -o tools/_bin/libmg_tcab_$v.so mipgen_b200/csrc/*.cu; done
    gpurun -- python tools/ablate_tc.py

mask bits: 1 no exp arithmetic, 2 no float->double conversions, 4 no MMAs issued, 8 no epilogue arithmetic, 16 no operand build."""
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
    ctx.set_svr_mode(3)
    pnl.score(mg.MG_WANT_SVR)
    ctx.sync()
    ctx.reset_timings()
    for _ in range(3):
        pnl.score(mg.MG_WANT_SVR)
    ctx.sync()
    t = ctx.timings()
    print("%-28s k_svr_tc %.3f ms/pass" % (os.path.basename(lib), t.ms_svr / 3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_bin", "libmg_tcab_*.so")),
                      key=lambda p: int(p.rsplit("_", 1)[1].split(".")[0]))
        for lib in libs:
            subprocess.run([sys.executable, os.path.abspath(__file__), lib], check=False)
