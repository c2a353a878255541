"""Full-size checks on the bench workload (BASELINE.json configs[1]/[2]: 60-region panel,
capture 162, 57 arm pairs, ~2.5e6 candidates, 2048-SV model) through size-independent
properties, plus the oracle on a random sample:

  * the grid kernel and the explicit-candidate kernel (different front-ends, same arithmetic)
    agree bit for bit on sampled candidates cut out of the regions on the host;
  * two runs give identical bits (no atomics / order dependence in any score);
  * the DMMA contraction agrees with the libsvm-order cross-check kernel on sampled rows;
  * a 256-candidate sample agrees with the CPU oracle at the usual tolerances;
  * the replayed enumeration (mg_tile_replay) never yields an invalid grid point and, with
    thresholds out of reach, yields exactly the statically valid ones.
"""
import os

import numpy as np
import pytest

import mipgen_b200 as mg
from mipgen_b200 import panel
from helpers import rel_err, tmpdir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import bench
    from oracle_api import Oracle
    ctx = mg.Context(0)
    cfg = panel.Config()
    ctx.set_config(cfg)
    model = bench.build_model(ctx, cfg, tmpdir())
    genome, regions = bench.make_panel(cfg, bench.N_REGIONS, bench.GENOME_SEED)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    offs, valid, lo, sv, _ = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    yield dict(ctx=ctx, cfg=cfg, model=model, regions=regions, offs=offs, valid=valid, lo=lo, sv=sv, oracle=Oracle())
    ctx.close()


def decode(cfg, r, local):
    strand = local & 1
    q = local >> 1
    p = q % cfg.n_pairs
    q //= cfg.n_pairs
    ci = q % len(cfg.captures)
    si = q // len(cfg.captures)
    s = cfg.first_scan_start(r) + si
    cap = cfg.captures[ci]
    e, l = cfg.ext_len[p], cfg.lig_len[p]
    t = s + cap - e - l - 1
    return s, t, e, l, strand


def cut_candidate(oracle, r, s, t, e, l, strand):
    o = r.seq_start
    if strand == 0:
        ext, lig, tgt = r.seq[s - e - o:s - o], r.seq[t + 1 - o:t + 1 + l - o], r.seq[s - o:t + 1 - o]
    else:
        ext = oracle.reverse_comp(r.seq[t + 1 - o:t + 1 + e - o])
        lig = oracle.reverse_comp(r.seq[s - l - o:s - o])
        tgt = oracle.reverse_comp(r.seq[s - o:t + 1 - o])
    return dict(ext=ext, lig=lig, tgt=tgt)


def test_panel_size_and_validity(setup):
    n = int(setup["offs"][-1])
    assert 2.0e6 < n < 3.0e6
    assert setup["valid"].mean() > 0.99
    ok = setup["valid"].astype(bool)
    assert np.isfinite(setup["sv"][ok]).all() and np.isnan(setup["sv"][~ok]).all()
    assert ((setup["lo"][ok] > 0) & (setup["lo"][ok] < 1)).all()


def test_grid_equals_explicit_and_oracle_on_sample(setup):
    ctx, cfg, regions, offs, oracle = setup["ctx"], setup["cfg"], setup["regions"], setup["offs"], setup["oracle"]
    rng = np.random.default_rng(123)
    picks = np.sort(rng.choice(np.nonzero(setup["valid"])[0], 256, replace=False))
    cands, lrc = [], []
    for g in picks:
        ri = int(np.searchsorted(offs, g, side="right") - 1)
        r = regions[ri]
        cands.append(cut_candidate(oracle, r, *decode(cfg, r, int(g - offs[ri]))))
        lrc.append(r.lrc)
    lrc = np.array(lrc)
    lo, sv, ft = ctx.score_candidates(cands, lrc, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
    assert np.array_equal(lo, setup["lo"][picks]), "grid and explicit front-ends must agree bit for bit"
    # the panel was scored by the factored kernel, the explicit candidates by the dense contraction
    assert rel_err(setup["sv"][picks], sv) <= 1e-11
    # oracle on the same sample
    h = oracle.svm_load_model(setup["model"])
    want_ft = np.array([oracle.get_parameters(c["ext"], c["lig"], c["tgt"], lrc[i]) for i, c in enumerate(cands)])
    want_lo = np.array([oracle.get_score(c["ext"], c["lig"], c["tgt"]) for c in cands])
    want_sv = oracle.svm_predict_rows(h, want_ft)
    oracle.svm_free(h)
    assert np.array_equal(ft, want_ft)
    assert rel_err(lo, want_lo) <= 1e-12
    assert rel_err(sv, want_sv) <= 1e-9
    # libsvm-order cross-check kernel on the device
    direct = ctx.svr_predict(ft, direct=True)
    assert rel_err(direct, want_sv) <= 1e-13
    assert rel_err(sv, direct) <= 1e-9
    assert rel_err(setup["sv"][picks], want_sv) <= 1e-9


def test_runs_are_bitwise_reproducible(setup):
    ctx, regions = setup["ctx"], setup["regions"][:12]
    a = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    b = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    n = int(a[0][-1])
    assert np.array_equal(a[2], b[2], equal_nan=True) and np.array_equal(a[3], b[3], equal_nan=True)
    assert np.array_equal(a[3], setup["sv"][:n], equal_nan=True), "a sub-panel scores exactly like the same regions inside the full panel"


def test_replay_enumerates_only_valid_points(setup):
    cfg, regions, offs, valid, sv, lo = (setup[k] for k in ("cfg", "regions", "offs", "valid", "sv", "lo"))
    total = 0
    for i in (0, 7, 33):
        a, b = offs[i], offs[i + 1]
        full = mg.tile_replay(cfg, regions[i], valid[a:b], sv[a:b], 1, True, 1e9)
        assert full.size == valid[a:b].sum() and valid[a:b][full].all()
        pruned = mg.tile_replay(cfg, regions[i], valid[a:b], sv[a:b], 1, True, float(np.nanmedian(sv[a:b])))
        assert pruned.size < full.size and np.all(np.diff(pruned) > 0) and np.isin(pruned, full).all()
        total += pruned.size
    assert total > 0


def test_multi_capture_panel_cfg4_style():
    """BASELINE.json configs[3] shape: capture sweep 120..250 step 5 (27 sizes, 3078 grid points per scan
    start), short regions so the static capture skip (mipgen.cpp:429) removes many sizes.  Checked through
    properties + an oracle sample; factored and dense SVR must agree everywhere."""
    from oracle_api import Oracle
    from helpers import random_model
    oracle = Oracle()
    cfg = panel.Config(250, 120, 5)
    genome = panel.lcg_genome(40000, 4004)
    regions = panel.make_regions(genome, 3, 60, 170, cfg, 4005)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    ctx.load_svr_model(random_model(oracle, panel.Config(), 200, 44, os.path.join(tmpdir(), "m.model")))
    assert ctx.svr_factored_available() > 0
    offs, valid, lo, sv, _ = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    ctx.set_svr_mode(1)
    _o, v2, _l, dense, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    ctx.set_svr_mode(0)
    assert int(offs[-1]) > 1_500_000 and np.array_equal(valid, v2)
    assert 0.2 < valid.mean() < 0.95, "the static capture skip should remove a good share of this grid"
    assert rel_err(sv, dense) <= 1e-11
    # the static skip rule, restated: capture > region length + max_mip_overlap (and not the smallest size)
    for i, r in enumerate(regions):
        g = valid[offs[i]:offs[i + 1]].reshape(-1, len(cfg.captures), cfg.n_pairs, 2)
        for ci, cap in enumerate(cfg.captures):
            skipped = cap > r.stop_flanked - r.start_flanked + cfg.max_mip_overlap and cap - cfg.capture_increment >= cfg.min_capture
            if skipped:
                assert not g[:, ci].any()
            else:
                assert g[5:-5, ci].all()
    # oracle on a random sample of valid grid points
    rng = np.random.default_rng(9)
    picks = np.sort(rng.choice(np.nonzero(valid)[0], 200, replace=False))
    cands, lrc = [], []
    for gidx in picks:
        ri = int(np.searchsorted(offs, gidx, side="right") - 1)
        r = regions[ri]
        cands.append(cut_candidate(oracle, r, *decode(cfg, r, int(gidx - offs[ri]))))
        lrc.append(r.lrc)
    want_lo = np.array([oracle.get_score(c["ext"], c["lig"], c["tgt"]) for c in cands])
    assert rel_err(lo[picks], want_lo) <= 1e-12
    lo2, sv2, _ft = ctx.score_candidates(cands, np.array(lrc), mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    assert np.array_equal(lo2, lo[picks]) and rel_err(sv[picks], sv2) <= 1e-11
    ctx.close()


def test_score_only_harness_cfg1():
    """BASELINE.json configs[0]: 1e5 explicit candidate MIPs (162 bp capture, 16-24 bp arms, consecutive scan
    starts, both strands) through Featurev5 + logistic + SVR.  Full size on the GPU; the oracle checks a sample."""
    from oracle_api import Oracle
    from helpers import random_model
    oracle = Oracle()
    rng = np.random.default_rng(1)
    genome = panel.lcg_genome(60000, 31337)
    n_pairs = 50000
    cands = []
    for i in range(n_pairs):
        e, l = int(rng.integers(16, 25)), int(rng.integers(16, 25))
        s = 1000 + i  # consecutive scan starts (0-based offset of the insert)
        size = 162 - e - l
        ext, tgt, lig = genome[s - e:s], genome[s:s + size], genome[s + size:s + size + l]
        cands.append(dict(ext=ext, lig=lig, tgt=tgt, ext_copy=1, lig_copy=1))
        # the minus-strand MIP over the same insert: arms swap sides and everything is reverse-complemented
        cands.append(dict(ext=oracle.reverse_comp(genome[s + size:s + size + e]), lig=oracle.reverse_comp(genome[s - l:s]),
                          tgt=oracle.reverse_comp(tgt), ext_copy=int(rng.choice([1, 2, 3, 10, 100, 101])), lig_copy=1))
    lrc = np.tile(oracle.long_range_content(genome[:3000], 1001, 2000), (len(cands), 1))
    ctx = mg.Context(0)
    model = random_model(oracle, panel.Config(), 256, 3, os.path.join(tmpdir(), "m.model"))
    ctx.load_svr_model(model)
    lo, sv, ft = ctx.score_candidates(cands, lrc, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
    assert lo.shape == (100000,) and np.isfinite(sv).all() and ((lo > 0) & (lo < 1)).all()
    picks = rng.choice(len(cands), 300, replace=False)
    h = oracle.svm_load_model(model)
    for k in picks:
        c = cands[k]
        wf = oracle.get_parameters(c["ext"], c["lig"], c["tgt"], lrc[k], ext_copy=c["ext_copy"], lig_copy=c["lig_copy"])
        assert np.array_equal(wf, ft[k])
        assert rel_err([lo[k]], [oracle.get_score(c["ext"], c["lig"], c["tgt"], ext_copy=c["ext_copy"], lig_copy=c["lig_copy"])]) <= 1e-12
        assert rel_err([sv[k]], [oracle.svm_predict(h, wf)]) <= 1e-9
    oracle.svm_free(h)
    ctx.close()
