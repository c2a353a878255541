"""CPU-side checks of the product: the C-ABI library loads, exports every symbol the header
declares, refuses to compute without a GPU, and its host logic (grid sizing, tile replay,
sharding) matches the oracle.  No compute call is made here."""
import os
import re

import numpy as np
import pytest

import mipgen_b200 as mg
from mipgen_b200 import _capi, panel, shard
from mipgen_b200.panel import Config
from oracle_api import Oracle
from helpers import small_config, synthetic_regions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "mipgen_b200.h")).read()
    declared = set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", hdr))
    lib = mg.load_library()
    bound = {s[0] for s in _capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mg.MgError) as e:
        mg.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """Nothing under mipgen_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for base in ("mipgen_b200", "include"):
        for dp, _dn, fn in os.walk(os.path.join(ROOT, base)):
            for f in fn:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_api|mipgen_oracle|liboracle|libmipgen_ref|orc_[a-z_]+\(", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_grid_size_and_first_scan_match_oracle():
    o = Oracle()
    for cfg in (Config(), small_config((40, 43), 250, 120, 5), small_config((45,), 162, 150, 0)):
        genome, regs = synthetic_regions(o, cfg, 3, 10, 300, 17, with_lrc=False)
        regs.append(panel.cut_region(genome, 30, 80, cfg))  # scan start clamp at the chromosome start
        for r in regs:
            assert mg.config_grid_size(cfg, r) == cfg.grid_size(r)
            v, _l, _s, _f = o.grid_region(r, cfg, None)
            assert v.size == cfg.grid_size(r)


def test_tile_replay_matches_oracle_on_synthetic_scores():
    """Host replay (mg_tile_replay) vs the oracle's restatement of mipgen.cpp:426-497 over random
    score grids that trip every rule: optimal-score skips, the logistic heuristic on -1000
    sentinels, int truncation of previous scores."""
    o = Oracle()
    rng = np.random.default_rng(8)
    cfg = small_config((40, 42, 45), 162, 147, 5)
    _g, regs = synthetic_regions(o, cfg, 2, 100, 160, 23, with_lrc=False)
    fired = 0
    for r in regs:
        v, _l, _s, _f = o.grid_region(r, cfg, None)
        for trial in range(6):
            score = rng.uniform(0.2, 1.05, v.size) if trial % 2 == 0 else rng.uniform(0.5, 3.0, v.size)
            score[rng.random(v.size) < 0.05] = -1000.0
            score[~v.astype(bool)] = np.nan
            for method in (0, 1, 2):
                for heur in (True, False):
                    upper = 0.98 if method != 1 else 2.2
                    a = mg.tile_replay(cfg, r, v, score, method, heur, upper)
                    b = o.tile_replay(r, cfg, v, score, method, heur, upper)
                    assert np.array_equal(a, b)
                    fired += b.size < v.sum()
    assert fired > 10


def test_tile_replay_and_grid_geometry_on_random_configs():
    """Random capture ranges / increments / arm-sum sets and regions hard against the chromosome start:
    grid size, enumeration replay and candidate geometry of the host helpers agree with the oracle."""
    o = Oracle()
    rng = np.random.default_rng(2024)
    all_sums = list(range(36, 50))
    for trial in range(12):
        sums = tuple(sorted(rng.choice(all_sums, int(rng.integers(1, 4)), replace=False).tolist()))
        max_cap = int(rng.integers(120, 260))
        min_cap = int(max(max(sums) + 1, max_cap - rng.integers(0, 40)))
        inc = int(rng.choice([0, 1, 3, 5, 7]))
        cfg = small_config(sums, max_cap, min_cap, inc)
        genome, regs = synthetic_regions(o, cfg, 2, 5, 120, 300 + trial, with_lrc=False)
        regs.append(panel.cut_region(genome, int(rng.integers(20, 60)), int(rng.integers(70, 140)), cfg))
        for r in regs:
            v, _l, _s, _f = o.grid_region(r, cfg, None)
            assert mg.config_grid_size(cfg, r) == v.size == cfg.grid_size(r)
            if v.size == 0:
                continue
            score = rng.uniform(0.3, 1.02, v.size)
            score[rng.random(v.size) < 0.03] = -1000.0
            score[~v.astype(bool)] = np.nan
            for method, upper in ((0, 0.98), (1, 0.9), (2, 0.98)):
                a = mg.tile_replay(cfg, r, v, score, method, True, upper)
                b = o.tile_replay(r, cfg, v, score, method, True, upper)
                assert np.array_equal(a, b), (trial, sums, max_cap, min_cap, inc, method)
            # geometry of a few enumerated candidates: every arm and the target lie inside the region's sequence
            idx = a[:: max(1, a.size // 50)]
            for m in mg.describe_candidates(cfg, r, idx):
                lo = min(m["ext_start"], m["lig_start"])
                hi = max(m["ext_stop"], m["lig_stop"])
                assert r.seq_start <= lo and hi <= r.seq_stop
                assert m["scan_stop"] - m["scan_start"] + 1 + m["ext_len"] + m["lig_len"] in cfg.captures
                assert m["ext_stop"] - m["ext_start"] + 1 == m["ext_len"] and m["lig_stop"] - m["lig_start"] + 1 == m["lig_len"]


def test_lpt_sharding_is_a_partition_and_balanced():
    rng = np.random.default_rng(1)
    costs = rng.integers(100, 100000, 61).tolist()
    for n in (1, 2, 4, 8):
        owned = shard.lpt_assign(costs, n)
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(61))
        loads = [sum(costs[i] for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(costs)


def test_tile_sizes_and_partition_are_host_arithmetic():
    """mg_tile_sizes / mg_partition_regions need no device: prefix sums match the Python config arithmetic, the partition is a
    deterministic longest-processing-time assignment that covers every region once and balances the grid sizes."""
    cfg = panel.Config(170, 150, 5)
    genome = panel.lcg_genome(panel.genome_length_for(23, 400, cfg), 12)
    regions = panel.make_regions(genome, 23, 60, 400, cfg, 13)
    g, s, p = mg.tile_sizes(cfg, regions)
    assert [int(x) for x in np.diff(g)] == [cfg.grid_size(r) for r in regions]
    assert [int(x) for x in np.diff(s)] == [cfg.n_scan(r) for r in regions]
    min_sum = min(e + l for e, l in zip(cfg.ext_len, cfg.lig_len))
    assert [int(x) for x in np.diff(p)] == [r.stop_flanked + cfg.max_capture - min_sum - 1 - cfg.first_scan_start(r) + 1 for r in regions]
    for parts in (1, 2, 3, 8):
        owner = mg.partition_regions(cfg, regions, parts)
        assert owner.min() >= 0 and owner.max() < parts and owner.size == len(regions)
        load = np.bincount(owner, weights=np.diff(g), minlength=parts)
        assert load.max() - load.min() <= np.diff(g).max()          # LPT bound
        assert np.array_equal(owner, mg.partition_regions(cfg, regions, parts))
        assert np.array_equal(owner, np.array([next(k for k, o in enumerate(shard.lpt_assign(list(np.diff(g)), parts)) if i in o)
                                               for i in range(len(regions))]))


def test_config_rejects_misordered_arm_pairs():
    """The replay of the tile loop's skips relies on pairs grouped by arm sum, sums descending (mipgen.cpp:431-438)."""
    ok = panel.Config()
    assert mg.config_grid_size(ok, _one_region(ok)) > 0
    bad = panel.Config(ext_len=[16, 20, 17], lig_len=[24, 25, 23])  # sums 40, 45, 40
    assert mg.config_grid_size(bad, _one_region(bad)) == -1


def _one_region(cfg):
    genome = panel.lcg_genome(8000, 3)
    return panel.cut_region(genome, 3001, 3100, cfg)


def test_batched_driver_source_edits_apply_exactly_once(tmp_path):
    """mipgen_b200/batched/make_source.py: each of the five anchored edits matches the reference's mipgen.cpp exactly once (the script
    exits otherwise) and leaves its marker in the patched copy.  Needs the reference sources: skipped on the GPU box."""
    import subprocess
    import sys
    ref = "/root/reference/mipgen.cpp"
    if not os.path.exists(ref):
        pytest.skip("reference sources not present")
    out = tmp_path / "mipgen_batched.cpp"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "mipgen_b200", "batched", "make_source.py"), ref, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = out.read_text(encoding="latin-1")
    for marker in ('#include "mipgen_batched.h"', '#include "batched_members.inc"', "b200_tile_feature(feature);", "mipgen_b200_take_pending_svr(&b200_score)",
                   "b200_write_fastqs(BWAFQ, ARMSFQ);", "if (b200_find_copy()) return;"):
        assert text.count(marker) == 1, marker
    # nothing else changed: the patched copy minus the inserted lines is a subsequence of the reference
    assert len(text) < len(open(ref, encoding="latin-1").read()) + 1200
