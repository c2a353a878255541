#!/usr/bin/env python3
"""BASELINE configs[4] at FULL scale on the GPUs of one box: 2e5 regions of U[100,200] bp (~30 Mb of targets, ~6e9 grid points),
capture 162, 57 arm pairs, SVR with the bench's 2048-SV model; score + condense + collapse through mg_tile_regions[_multi] in
sub-batches of <= 2^26 grid points (bounded device memory; host buffers in, winners out).

    gpurun [--gpus N] -- python tools/run_cfg5_full.py [n_regions] > gpurun_out/cfg5_full.json

Prints one JSON line: wall seconds of the pass, grid points per second, device memory high-water mark, and a consistency check of a
sample of regions against their own single-region calls (the streaming must not change any winner)."""
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402


def device_mem_used():
    rt = ctypes.CDLL("libcudart.so")
    free, total = ctypes.c_size_t(), ctypes.c_size_t()
    rt.cudaMemGetInfo(ctypes.byref(free), ctypes.byref(total))
    return total.value - free.value


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    n_gpus = int(os.environ.get("CFG5_GPUS", "1"))
    cfg = panel.Config()
    ctxs = [mg.Context(d) for d in range(n_gpus)]
    work = tempfile.mkdtemp()
    ctxs[0].set_config(cfg)
    model = bench.build_model(ctxs[0], cfg, work)
    for c in ctxs[1:]:
        c.set_config(cfg)
        c.load_svr_model(model)
    t0 = time.perf_counter()
    g = panel.lcg_genome(panel.genome_length_for(n, 200, cfg, gap=300), bench.GENOME_SEED + 5)
    regions = panel.make_regions(g, n, 100, 200, cfg, bench.GENOME_SEED + 6, gap=300)
    lrc0 = ctxs[0].long_range_content(regions[0].flank_seq, regions[0].seq_start, regions[0].seq_stop)
    for r in regions:
        r.lrc = lrc0   # one long-range vector for all (as bench.py's cfg5_sample): K-lrc per region is not what this run measures
    t_gen = time.perf_counter() - t0
    sel = dict(method=1, lower=1.5, upper=2.2)
    who = ctxs if n_gpus > 1 else ctxs[0]
    mg.tile_regions(who, regions[:400 * n_gpus], mg.MG_WANT_SVR, select=sel)   # warm-up: workspaces grow to a full sub-batch
    ctxs[0].reset_timings()
    t0 = time.perf_counter()
    t = mg.tile_regions(who, regions, mg.MG_WANT_SVR, select=sel)
    dt = time.perf_counter() - t0
    mem = device_mem_used()
    tm = ctxs[0].timings()
    n_grid = int(t.grid_off[-1])
    # consistency: a sample of regions on their own
    rng = np.random.default_rng(1)
    pick = sorted(rng.choice(n, min(40, n), replace=False).tolist())
    same = True
    for i in pick:
        one = mg.tile_regions(ctxs[0], [regions[i]], mg.MG_WANT_SVR, select=sel)
        a, b = int(t.scan_off[i]), int(t.scan_off[i + 1])
        same &= bool(np.array_equal(one.scan_best, t.scan_best[a:b]))
        a, b = int(t.pos_off[i]), int(t.pos_off[i + 1])
        same &= bool(np.array_equal(one.pos_best, t.pos_best[a:b]))
        same &= bool(np.array_equal(one.scan_best_svr, t.scan_best_svr[int(t.scan_off[i]):int(t.scan_off[i + 1])], equal_nan=True))
    print(json.dumps({"what": "BASELINE configs[4] at full scale: %d regions of U[100,200] bp (%.1f Mb of targets), capture 162, 57 arm pairs, SVR (2048 SV), "
                              "score + condense + collapse through mg_tile_regions%s in sub-batches of <= 2^26 grid points"
                              % (n, sum(r.stop_flanked - r.start_flanked + 1 for r in regions) / 1e6, "_multi" if n_gpus > 1 else ""),
                      "n_gpus": n_gpus, "regions": n, "grid_points": n_grid, "seconds": t.call_seconds, "value": n_grid / t.call_seconds,
                      "unit": "candidates/s (end to end through the C call: host buffers in, winners out)",
                      "seconds_incl_python_marshalling": dt,
                      "kernel_seconds_gpu0": (tm.ms_feat + tm.ms_svr + tm.ms_other) / 1e3, "scan_start_winners": int((t.scan_best >= 0).sum()),
                      "position_winners": int((t.pos_best >= 0).sum()), "device_memory_in_use_bytes_gpu0": mem,
                      "host_generation_seconds": t_gen, "sampled_regions_equal_their_single_region_calls": same, "sampled_regions": len(pick),
                      "library_build": bench.library_build_id()}))
    assert same


if __name__ == "__main__":
    main()
