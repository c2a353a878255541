"""Device selection front-end (K-condense, K-collapse) against the oracle's restatement of the tile
replay + condense_mips + collapse_mips, which is itself pinned to the reference CLI's
all_mips.txt / collapsed_mips.txt (tests/test_selection_pinning.py).  The kernels only compare and
copy scores, so on the same score grid the results must be IDENTICAL."""
import os

import numpy as np
import pytest

import mipgen_b200 as mg
from mipgen_b200 import panel
from helpers import small_config, synthetic_regions, calibrated_model, tmpdir
from test_gpu_parity import last_pair_threshold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


def check_against_oracle(oracle, cfg, regions, offs, valid, score, so, sb, po, pb, method, lower, upper, heuristic, mac=75, tac=20,
                         thr=0.5):
    fired = 0
    for i, r in enumerate(regions):
        a, b = offs[i], offs[i + 1]
        enum_idx = oracle.tile_replay(r, cfg, valid[a:b], score[a:b], method, heuristic, upper)
        want_sb, want_pb = oracle.select(r, cfg, score[a:b], enum_idx, lower, upper, mac, tac, thr)
        want_sb = np.where(want_sb >= 0, want_sb + a, -1)
        want_pb = np.where(want_pb >= 0, want_pb + a, -1)
        assert np.array_equal(sb[so[i]:so[i + 1]], want_sb), "scan_strand_best_mip differs (region %d)" % i
        assert np.array_equal(pb[po[i]:po[i + 1]], want_pb), "pos_strand_best_mip differs (region %d)" % i
        fired += enum_idx.size < valid[a:b].sum()
    return fired


def test_select_matches_oracle(oracle):
    cfg = small_config((40, 43, 45), 162, 152, 5)
    rng = np.random.default_rng(4)
    genome, regions = synthetic_regions(oracle, cfg, 3, 110, 170, 61)
    regions.append(panel.cut_region(genome, 150, 260, cfg, 0, "edge"))  # clamped scan range
    regions[-1].lrc = rng.uniform(0, 0.3, 44)
    # copy tables on one region: exercise the copy-number rules of condense (:1689, :1709-1715) and collapse (:1628)
    regions[1].copies = rng.choice([1, 1, 1, 1, 2, 5, 9, 30, 100], size=(len(cfg.oligo_sizes), len(regions[1].seq))).astype(np.int32)
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, cfg, 64, 5, os.path.join(d, "m.model"), feats[np.isfinite(feats[:, 0])][::53])
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    valid, lo, sv, _ = pnl.fetch(valid=True, logistic=True, svr=True)
    offs = pnl.offsets
    a, b = offs[0], offs[1]
    fired = 0
    cases = [(1, sv, 1.5, 2.2, True), (1, sv, 0.5, last_pair_threshold(cfg, valid[a:b], sv[a:b]), True),
             (0, lo, 0.9, 0.98, True), (0, lo, 0.6, last_pair_threshold(cfg, valid[a:b], lo[a:b]), True),
             (0, lo, 0.6, last_pair_threshold(cfg, valid[a:b], lo[a:b]), False), (2, lo, 0.8, 0.9, True)]
    for method, score, lower, upper, heur in cases:
        so, sb, po, pb = pnl.select(regions, method, lower, upper, heur)
        fired += check_against_oracle(oracle, cfg, regions, offs, valid, score, so, sb, po, pb, method, lower, upper, heur)
        assert (sb >= 0).any() and (pb >= 0).any()
    # tighter copy limits
    so, sb, po, pb = pnl.select(regions, 1, 1.5, 2.2, True, max_arm_copy=20, target_arm_copy=4)
    check_against_oracle(oracle, cfg, regions, offs, valid, sv, so, sb, po, pb, 1, 1.5, 2.2, True, 20, 4)
    assert fired >= 4, "pruning must fire in some cases"
    pnl.close()
    ctx.close()


def test_select_with_masked_snp_unmappable_inputs(oracle):
    """arm_fraction_masked / snp_count / mapping_failed (mipgen.cpp:606-625, 634-760) from the regions' masked_seq / snp /
    unmappable inputs: device condense/collapse == the oracle's restatement, which tests/test_selection_pinning.py pins to
    the reference CLI run against the rule-driven stub bwa / trf / tabix."""
    import stub_rules
    cfg = small_config((40, 45), 162, 157, 5)
    genome, regions = synthetic_regions(oracle, cfg, 4, 110, 170, 71)
    snps = stub_rules.snp_positions(genome, regions)
    stub_rules.decorate(cfg, genome, regions[:3], snps=snps)   # region 3 stays plain: mixed panels must work
    regions[2].copies = None                                    # masked / snp / unmappable without a copy table
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[3], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, cfg, 64, 5, os.path.join(d, "m.model"), feats[np.isfinite(feats[:, 0])][::53])
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    ctx.load_svr_model(model)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    valid, lo, sv, _ = pnl.fetch(valid=True, logistic=True, svr=True)
    offs = pnl.offsets
    for method, score, lower, upper, thr in [(1, sv, 1.5, 2.2, 0.5), (0, lo, 0.8, 0.9, 0.5), (1, sv, 0.6, 1.9, 0.3), (0, lo, 0.9, 0.98, 0.05)]:
        so, sb, po, pb = pnl.select(regions, method, lower, upper, True, masked_arm_threshold=thr)
        check_against_oracle(oracle, cfg, regions, offs, valid, score, so, sb, po, pb, method, lower, upper, True, thr=thr)
        # the inputs must matter: a panel of the same sequences without them selects differently somewhere
    plain = [panel.Region(r.start_flanked, r.stop_flanked, r.seq_start, r.seq_stop, r.seq, r.lrc, r.flank_seq, r.label, r.copies) for r in regions]
    pnl2 = ctx.panel(plain)
    pnl2.score(mg.MG_WANT_SVR)
    _so, sb2, _po, pb2 = pnl2.select(plain, 1, 1.5, 2.2)
    so, sb, po, pb = pnl.select(regions, 1, 1.5, 2.2)
    assert not np.array_equal(sb, sb2) and not np.array_equal(pb, pb2)
    # records with failure flags need design_mip: the "000" formatter refuses such regions instead of printing wrong flags
    with pytest.raises(mg.MgError):
        mg.design_records(cfg, regions[0], np.array([0]), sv[offs[0]:offs[1]], "1", "x", 1, 2, 1)
    pnl.close()
    pnl2.close()
    ctx.close()


def test_tile_regions_single_and_multi_context(oracle):
    """mg_tile_regions (sub-batched, region-local winners + their scores) and mg_tile_regions_multi / mg_score_regions_multi
    (LPT partition, one host thread per context) reproduce the plain panel calls exactly, whatever the batching."""
    cfg = small_config((40, 43, 45), 162, 157, 5)
    genome, regions = synthetic_regions(oracle, cfg, 7, 90, 200, 91)
    d = tmpdir()
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, cfg, 48, 5, os.path.join(d, "m.model"), feats[np.isfinite(feats[:, 0])][::53])
    ctxs = [mg.Context(0), mg.Context(0)]   # two contexts (own streams / host threads); on a multi-GPU box they sit on two devices
    for c in ctxs:
        c.set_config(cfg)
        c.load_svr_model(model)
    pnl = ctxs[0].panel(regions)
    pnl.score(mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    valid, lo, sv, _ = pnl.fetch(valid=True, logistic=True, svr=True)
    offs = pnl.offsets
    so, sb, po, pb = pnl.select(regions, 2, 0.8, 0.9)
    want_sb = np.where(sb >= 0, sb - np.repeat(offs[:-1], np.diff(so))[:, None], -1)
    want_pb = np.where(pb >= 0, pb - np.repeat(offs[:-1], np.diff(po))[:, None], -1)
    sel = dict(method=2, lower=0.8, upper=0.9)
    want = mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR
    smallest = int(np.diff(offs).min())
    for who, cap in ((ctxs[0], 0), (ctxs[0], smallest), (ctxs, 0), (ctxs, 3 * smallest)):
        t = mg.tile_regions(who, regions, want, select=sel, full_grids=True, max_batch_candidates=cap)
        assert np.array_equal(t.grid_off, offs) and np.array_equal(t.scan_off, so) and np.array_equal(t.pos_off, po)
        assert np.array_equal(t.scan_best, want_sb) and np.array_equal(t.pos_best, want_pb)
        assert np.array_equal(t.valid, valid) and np.array_equal(t.logistic, lo, equal_nan=True) and np.array_equal(t.svr, sv, equal_nan=True)
        has = sb >= 0
        assert np.array_equal(t.scan_best_logistic[has], lo[sb[has]]) and np.array_equal(t.scan_best_svr[has], sv[sb[has]])
        assert np.isnan(t.scan_best_svr[~has]).all()
    o2, v2, l2, s2 = ctxs[0].score_regions_multi(ctxs[1:], regions, want)
    assert np.array_equal(o2, offs) and np.array_equal(v2, valid) and np.array_equal(l2, lo, equal_nan=True) and np.array_equal(s2, sv, equal_nan=True)
    # a panel created under an earlier config is refused, not mis-indexed
    ctxs[0].set_config(small_config((40, 45)))
    with pytest.raises(mg.MgError):
        pnl.score(mg.MG_WANT_LOGISTIC)
    pnl.close()
    for c in ctxs:
        c.close()


def test_select_full_size_properties(oracle):
    """Bench panel (60 regions, 2.5e6 candidates): every winner is a valid grid point of the right region, strand
    and scan start; every position winner covers its position; winners are enumerated candidates."""
    import bench
    cfg = panel.Config()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    bench.build_model(ctx, cfg, tmpdir())
    _g, regions = bench.make_panel(cfg, bench.N_REGIONS, bench.GENOME_SEED)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_SVR)
    valid, _lo, sv, _ = pnl.fetch(valid=True, svr=True)
    offs = pnl.offsets
    so, sb, po, pb = pnl.select(regions, 1, 1.5, 2.2)
    per_scan = cfg.n_pairs * 2
    for i in (0, 17, 59):
        r = regions[i]
        s_idx = np.arange(so[i + 1] - so[i])
        for strand in (0, 1):
            w = sb[so[i]:so[i + 1], strand]
            ok = w >= 0
            assert ok.all() and valid[w].all()
            assert np.array_equal((w - offs[i]) // per_scan, s_idx) and ((w & 1) == strand).all()
            pw = pb[po[i]:po[i + 1], strand]
            has = pw >= 0
            scan_start = cfg.first_scan_start(r) + (pw[has] - offs[i]) // per_scan
            pidx = ((pw[has] - offs[i]) >> 1) % cfg.n_pairs
            size = 162 - (np.array(cfg.ext_len)[pidx] + np.array(cfg.lig_len)[pidx])
            pos = cfg.first_scan_start(r) + np.nonzero(has)[0]
            assert ((scan_start <= pos) & (pos <= scan_start + size - 1)).all()
            assert np.isin(pw[has], w).all(), "a position winner is one of the scan-start winners"
        # exact check against the oracle on this region
        enum_idx = oracle.tile_replay(r, cfg, valid[offs[i]:offs[i + 1]], sv[offs[i]:offs[i + 1]], 1, True, 2.2)
        wsb, wpb = oracle.select(r, cfg, sv[offs[i]:offs[i + 1]], enum_idx, 1.5, 2.2)
        assert np.array_equal(sb[so[i]:so[i + 1]], np.where(wsb >= 0, wsb + offs[i], -1))
        assert np.array_equal(pb[po[i]:po[i + 1]], np.where(wpb >= 0, wpb + offs[i], -1))
    ctx.reset_timings()
    pnl.select(regions, 1, 1.5, 2.2)
    t = ctx.timings()
    print("select: %.3f ms for %d candidates" % (t.ms_other, pnl.n_candidates))
    pnl.close()
    ctx.close()
