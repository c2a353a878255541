"""The C-ABI from plain C: tests/c/c_abi_example.c compiles as C99 against include/mipgen_b200.h and links against the library
(CPU); on a GPU it runs over two contexts and must reproduce the numbers the Python view of the same calls gives."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "c_abi_example.c")


def _build(out_dir: str) -> str:
    from mipgen_b200 import build
    lib = build.build_library()
    exe = os.path.join(out_dir, "c_abi_example")
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    cmd = [cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", os.path.dirname(lib), "-lmipgen_b200", "-Wl,-rpath," + os.path.dirname(lib), "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_compiles_as_c99_and_links(tmp_path):
    exe = _build(str(tmp_path))
    assert os.path.getsize(exe) > 0
    # without a device the program must fail loudly (mg_create), not fall back to anything
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        r = subprocess.run([exe, "none.model", "none.txt", "0"], capture_output=True, text=True)
        assert r.returncode != 0 and "mg_create" in r.stderr


@pytest.mark.gpu
def test_c_example_equals_the_python_view(tmp_path):
    import mipgen_b200 as mg
    from mipgen_b200 import panel
    from helpers import random_model
    from oracle_api import Oracle
    oracle = Oracle()
    cfg = panel.Config()
    genome = panel.lcg_genome(panel.genome_length_for(5, 120, cfg), 808)
    regions = panel.make_regions(genome, 5, 60, 120, cfg, 809)
    model = random_model(oracle, cfg, 200, 7, str(tmp_path / "m.model"))
    with open(tmp_path / "regions.txt", "w") as f:
        for r in regions:
            f.write("%d %d %d %d %s\n" % (r.seq_start, r.seq_stop, r.start_flanked, r.stop_flanked, r.seq.decode()))
    exe = _build(str(tmp_path))
    out = subprocess.run([exe, model, str(tmp_path / "regions.txt"), "0", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = {}
    for line in out.stdout.strip().splitlines():
        t = line.split()
        got.update({t[i]: float(t[i + 1]) for i in range(0, len(t), 2)})
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    _o, valid, lo, sv, _f = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    ok = valid.astype(bool)
    assert got["regions"] == 5 and got["contexts"] == 2 and got["candidates"] == valid.size and got["valid"] == ok.sum()
    assert abs(got["sum_logistic"] - lo[ok].sum()) <= 1e-9 * abs(lo[ok].sum())
    assert abs(got["sum_svr"] - sv[ok].sum()) <= 1e-9 * abs(sv[ok].sum())
    t = mg.tile_regions(ctx, regions, mg.MG_WANT_SVR, select=dict(method=1, lower=1.5, upper=2.2))
    has = t.scan_best >= 0
    assert got["scan_start_winners"] == has.sum()
    assert abs(got["sum_winner_svr"] - t.scan_best_svr[has].sum()) <= 1e-9 * abs(t.scan_best_svr[has].sum())
    g = ctx.genome([r.seq.decode() for r in regions])
    tabs = g.count_arm_copies(regions, cfg.oligo_sizes)
    assert got["oligos"] == sum(int((x != 0).sum()) for x in tabs) and got["single_copy"] == sum(int((x == 1).sum()) for x in tabs)
    g.close()
    ctx.close()
