#!/usr/bin/env python3
"""Pipeline timeline of one CTA of the tensor-core K-svr (library built with -DMG_TC_TRACE): per 64-SV tile the cycle at which
the operand copy was issued, the operands had landed, the MMAs were issued, the accumulators were complete, epilogue warp 0 had
read them and had finished the tile.
    nvcc ... -DMG_TC_TRACE -o tools/_bin/libmg_tctrace.so mipgen_b200/csrc/*.cu ; gpurun -- python tools/trace_tc.py"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mipgen_b200._capi as capi  # noqa: E402
capi.LIB_PATH = os.path.join(ROOT, "tools", "_bin", "libmg_tctrace.so")
import bench  # noqa: E402
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402

cfg = panel.Config()
ctx = mg.Context(0)
ctx.set_config(cfg)
bench.build_model(ctx, cfg, tempfile.mkdtemp())
_g, regions = bench.make_panel(cfg, 60, bench.GENOME_SEED)
for r in regions:
    r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
pnl = ctx.panel(regions)
ctx.set_svr_mode(3)
for _ in range(3):
    pnl.score(mg.MG_WANT_SVR)
ctx.sync()
t = np.zeros(64 * 8, np.int64)
lib = C.CDLL(capi.LIB_PATH)
assert lib.mg_tc_trace_fetch(t.ctypes.data_as(C.c_void_p)) == 0
t = t.reshape(64, 8)[:32]
t0 = t[0, 0]
print("tile  copy_issued  operands_landed  mma_issued  accum_complete  epi_read  epi_done   | mma->complete  epi_compute  tile_period")
for j in range(32):
    r = t[j] - t0
    print("%4d %11d %15d %11d %15d %9d %9d   | %12d %12d %12d" % (j, r[0], r[1], r[2], r[3], r[4], r[5], r[3] - r[2], r[5] - r[3], (t[j, 5] - t[j - 1, 5]) if j else 0))
