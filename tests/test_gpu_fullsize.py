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


# ----------------------------------------------------------------------------------------------------------
# whole bench panel against the oracle: every candidate, not a sample
# ----------------------------------------------------------------------------------------------------------
def _oracle_region_job(args):
    """One region through the oracle in a forked worker: returns (valid, logistic, sha256 of the feature rows)."""
    import hashlib
    from oracle_api import Oracle
    r, cfg = args
    o = Oracle()
    v, lo, _s, ft = o.grid_region(r, cfg, None, want_logistic=True, want_feats=True)
    ft[~v.astype(bool)] = 0.0
    return v, lo, hashlib.sha256(np.ascontiguousarray(ft).tobytes()).hexdigest()


def test_whole_bench_panel_features_and_logistic_equal_the_oracle(setup):
    """All 2.53 M candidates of the bench panel: validity and the 192 features bit-exact (compared per region through a digest of
    the raw feature bytes), logistic <= 1e-12 relative -- the oracle runs on the host cores, one region per task."""
    import hashlib
    import multiprocessing as mp
    ctx, cfg, regions, offs = setup["ctx"], setup["cfg"], setup["regions"], setup["offs"]
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        want = pool.map(_oracle_region_job, [(r, cfg) for r in regions], chunksize=1)
    worst = 0.0
    for i0 in range(0, len(regions), 6):   # feature rows of six regions at a time (~0.4 GB on the host)
        part = regions[i0:i0 + 6]
        o2, v2, _l, _s, ft = ctx.score_regions(part, mg.MG_WANT_FEATURES)
        for k, r in enumerate(part):
            i = i0 + k
            a, b = int(offs[i]), int(offs[i + 1])
            wv, wl, wdigest = want[i]
            assert np.array_equal(setup["valid"][a:b], wv) and np.array_equal(v2[o2[k]:o2[k + 1]], wv), "validity differs in region %d" % i
            rows = ft[o2[k]:o2[k + 1]].copy()
            rows[~wv.astype(bool)] = 0.0
            assert hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest() == wdigest, "feature rows differ in region %d" % i
            ok = wv.astype(bool)
            worst = max(worst, rel_err(setup["lo"][a:b][ok], wl[ok]))
    print("whole panel: %d candidates, features bit-exact, logistic max rel err %.2e" % (int(offs[-1]), worst))
    assert worst <= 1e-12


def _oracle_svr_job(args):
    from oracle_api import Oracle
    model, rows = args
    o = Oracle()
    h = o.svm_load_model(model)
    out = o.svm_predict_rows(h, rows)
    o.svm_free(h)
    return out


def test_stratified_svr_sample_of_the_bench_panel_equals_the_oracle(setup):
    """12,000 candidates of the bench panel, 200 per region (so every region, both strands and all arm pairs are hit), through
    libsvm's arithmetic (2048 SV): the factored FP64 kernel <= 1e-9, the tensor-core form <= 1e-7 (north star: 1e-6)."""
    import multiprocessing as mp
    ctx, regions, offs, valid = setup["ctx"], setup["regions"], setup["offs"], setup["valid"]
    rng = np.random.default_rng(2024)
    picks = np.concatenate([np.sort(rng.choice(np.nonzero(valid[offs[i]:offs[i + 1]])[0], 200, replace=False)) + offs[i] for i in range(len(regions))])
    rows = []
    for i0 in range(0, len(regions), 6):
        part = regions[i0:i0 + 6]
        o2, _v, _l, _s, ft = ctx.score_regions(part, mg.MG_WANT_FEATURES)
        for k in range(len(part)):
            sel = picks[(picks >= offs[i0 + k]) & (picks < offs[i0 + k + 1])] - offs[i0 + k]
            rows.append(ft[o2[k] + sel])
    rows = np.concatenate(rows)
    n_proc = min(16, os.cpu_count() or 1)
    with mp.get_context("fork").Pool(n_proc) as pool:
        want = np.concatenate(pool.map(_oracle_svr_job, [(setup["model"], c) for c in np.array_split(rows, n_proc)]))
    e64 = rel_err(setup["sv"][picks], want)
    ctx.set_svr_mode(3)
    _o, _v, _l, tc, _f = ctx.score_regions(regions, mg.MG_WANT_SVR)
    ctx.set_svr_mode(0)
    etc = rel_err(tc[picks], want)
    print("SVR vs libsvm arithmetic on %d stratified candidates: FP64 factored kernel %.2e, tensor-core kernel %.2e" % (picks.size, e64, etc))
    assert e64 <= 1e-9 and etc <= 1e-7


def test_gpu_equals_the_compiled_reference_directly(setup):
    """The same comparison against the unmodified reference objects themselves (oracle/_ref/libmipgen_ref.so: SVMipv4::get_parameters,
    get_score, svm_predict), not their restatement: one region of the bench panel, every candidate."""
    from oracle_api import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref = Ref()
    ctx, cfg, regions, offs = setup["ctx"], setup["cfg"], setup["regions"], setup["offs"]
    i = int(np.argmin(np.diff(offs)))   # the smallest region keeps the reference's 2048-SV predictions to a few seconds on one core
    r = regions[i]
    h = ref.svm_load_model(setup["model"])
    wv, wl, ws, wf = ref.grid_region(r, cfg, h, want_logistic=True, want_svr=False, want_feats=True)
    n = 40 * cfg.n_pairs * 2   # the first 40 scan starts = 4,560 candidates through the reference's svm_predict
    want_svr = ref.svm_predict_rows(h, wf[:n])
    ref.svm_free(h)
    a = int(offs[i])
    _o, v, lo, sv, ft = ctx.score_regions([r], mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR | mg.MG_WANT_FEATURES)
    ok = wv.astype(bool)
    assert np.array_equal(v, wv) and np.array_equal(ft[ok], wf[ok])
    assert rel_err(lo[ok], wl[ok]) <= 1e-12
    assert rel_err(sv[:n][ok[:n]], want_svr[ok[:n]]) <= 1e-9
    assert np.array_equal(lo, setup["lo"][a:a + lo.size], equal_nan=True)
    lrc_ref = ref.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    assert np.array_equal(lrc_ref, r.lrc), "K-lrc equals Featurev5::get_long_range_content bit for bit"


def test_cfg5_shape_streams_through_bounded_memory():
    """BASELINE configs[4] shape at 1/100 scale: 2,000 regions of U[100,200] bp, capture 162, SVR, through mg_tile_regions with a
    small sub-batch bound (so the region list is walked in ~60 panels): winners equal the ones of plain per-region panels on a
    sample, and an oracle sample of the winners' scores holds."""
    from oracle_api import Oracle
    from helpers import calibrated_model, small_config
    oracle = Oracle()
    cfg = panel.Config()
    n = 2000
    genome = panel.lcg_genome(panel.genome_length_for(n, 200, cfg, gap=300), 515)
    regions = panel.make_regions(genome, n, 100, 200, cfg, 516, gap=300)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    for r in regions[:5]:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    for k, r in enumerate(regions[5:]):
        r.lrc = regions[k % 5].lrc
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, small_config((40, 45)), 96, 5, os.path.join(tmpdir(), "m.model"), feats[np.isfinite(feats[:, 0])][::97])
    ctx.load_svr_model(model)
    sel = dict(method=1, lower=1.5, upper=2.2)
    t = mg.tile_regions(ctx, regions, mg.MG_WANT_SVR, select=sel, max_batch_candidates=1 << 20)
    assert int(t.grid_off[-1]) > 5.5e7 and (t.scan_best >= 0).mean() > 0.99
    h = oracle.svm_load_model(model)
    for i in (0, 777, 1999):
        r = regions[i]
        pnl = ctx.panel([r])
        pnl.score(mg.MG_WANT_SVR)
        valid, _lo, sv, _ = pnl.fetch(valid=True, svr=True)
        _so, sb, _po, pb = pnl.select([r], 1, 1.5, 2.2)
        pnl.close()
        assert np.array_equal(t.scan_best[t.scan_off[i]:t.scan_off[i + 1]], sb) and np.array_equal(t.pos_best[t.pos_off[i]:t.pos_off[i + 1]], pb)
        has = sb >= 0
        assert np.array_equal(t.scan_best_svr[t.scan_off[i]:t.scan_off[i + 1]][has], sv[sb[has]])
        # the oracle on this region: same winners, scores within tolerance
        wv, _wl, ws, _wf = oracle.grid_region(r, cfg, h, want_logistic=False, want_svr=True)
        assert np.array_equal(valid, wv) and rel_err(sv, ws) <= 1e-9
        enum_idx = oracle.tile_replay(r, cfg, wv, sv, 1, True, 2.2)
        wsb, wpb = oracle.select(r, cfg, sv, enum_idx, 1.5, 2.2)
        assert np.array_equal(sb, wsb) and np.array_equal(pb, wpb)
    oracle.svm_free(h)
    ctx.close()


def test_full_scale_cfg5_runner_at_small_scale():
    """tools/run_cfg5_full.py (the script behind profiles/r02_cfg5_full_*.json) on 300 regions: it runs, its sampled regions equal their
    single-region calls, and it reports the C call's own seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "run_cfg5_full.py"), "300"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["regions"] == 300 and d["grid_points"] > 8e6 and d["sampled_regions_equal_their_single_region_calls"] is True
    assert 0 < d["seconds"] <= d["seconds_incl_python_marshalling"] and d["scan_start_winners"] > 0
