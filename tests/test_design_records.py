"""The batched caller's output side: grid indices -> the records print_details writes
(mipgen.cpp:765-794), through mg_describe_candidates + mg_format_mip_record.

Golden: SHA-256 and line count of the reference CLI's own all_mips.txt / collapsed_mips.txt for the two cases of
test_selection_pinning.py (tests/golden/records_*.json, regenerate with `python tests/test_design_records.py`
where /root/reference is present).  The CPU test takes the scores from the oracle, the GPU test from the device;
both must reproduce the files byte for byte.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipgen_b200 as mg  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from cli_util import REF_CLI, run_cli  # noqa: E402
from helpers import tmpdir  # noqa: E402
import test_selection_pinning as sp  # noqa: E402

# records with failure flags other than "000" need design_mip's SNP / masking logic: the stub-rule cases are out of the
# formatter's scope (it refuses such regions)
PLAIN_CASES = [c for c in sp.CASES if c not in sp.STUB_CASES]

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEADER = (">mip_key\t%s_score\tchr\text_probe_start\text_probe_stop\text_probe_copy\text_probe_sequence\tlig_probe_start\t"
          "lig_probe_stop\tlig_probe_copy\tlig_probe_sequence\tmip_scan_start_position\tmip_scan_stop_position\t"
          "scan_target_sequence\tmip_sequence\tfeature_start_position\tfeature_stop_position\tprobe_strand\tfailure_flags\tmip_name\n")


def file_digest(text: str):
    return {"lines": text.count("\n"), "sha256": hashlib.sha256(text.encode()).hexdigest()}


def write_files(cfg, regions, method, lower, upper, score_of):
    """all_mips.txt and collapsed_mips.txt as the batched caller writes them.  score_of(i, region) ->
    (valid, score, scan_best, pos_best) of region i; enumeration is mg_tile_replay on those scores."""
    name = "svr" if method == 1 else "logistic"
    all_txt, col_txt = HEADER % name, HEADER % name
    n_all = n_col = 0
    for i, r in enumerate(regions):
        valid, score, pos_best = score_of(i, r)
        enum_idx = mg.tile_replay(cfg, r, valid, score, method, True, upper)
        # zero flank: the feature is the flanked region itself (mipgen.cpp:1153-1160 with -feature_flank 0)
        all_txt += mg.design_records(cfg, r, enum_idx, score, "1", r.label, r.start_flanked, r.stop_flanked, n_all + 1)
        n_all += enum_idx.size
        winners = np.array([pos_best[p, s] for p in range(pos_best.shape[0]) for s in (0, 1) if pos_best[p, s] >= 0], np.int64)
        col_txt += mg.design_records(cfg, r, winners, score, "1", r.label, r.start_flanked, r.stop_flanked, n_col + 1)
        n_col += winners.size
    return all_txt, col_txt


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


@pytest.mark.parametrize("case", sorted(PLAIN_CASES))
def test_records_from_oracle_scores_equal_reference_files(oracle, case):
    g = json.load(open(os.path.join(GOLDEN_DIR, "records_%s.json" % case)))
    _flags, method, lower, upper, cfg = sp.CASES[case]
    d = tmpdir()
    model = sp.model_for(oracle, d)
    _genome, regions = sp.inputs(oracle, cfg)
    h = oracle.svm_load_model(model) if method == 1 else None

    def score_of(i, r):
        valid, lo, sv, _ = oracle.grid_region(r, cfg, h, want_logistic=method != 1, want_svr=method == 1)
        score = sv if method == 1 else lo
        enum_idx = oracle.tile_replay(r, cfg, valid, score, method, True, upper)
        _sb, pb = oracle.select(r, cfg, score, enum_idx, lower, upper)
        return valid, score, pb

    all_txt, col_txt = write_files(cfg, regions, method, lower, upper, score_of)
    if h:
        oracle.svm_free(h)
    assert file_digest(all_txt) == g["all_mips"], "all_mips.txt differs from the reference CLI's file"
    assert file_digest(col_txt) == g["collapsed_mips"], "collapsed_mips.txt differs from the reference CLI's file"


def test_describe_candidates_geometry_copies_and_errors(oracle):
    """mg_describe_candidates: Plus/Minus geometry (PlusSVMipv4.cpp:7-14, MinusSVMipv4.cpp:30-37), arm copy numbers
    from the per-oligo-size table (mipgen.cpp:612-613, absent key = 0), and its error path."""
    cfg = sp.small_config((40, 43, 45), 162, 152, 5)
    _genome, regions = sp.inputs(oracle, cfg)
    r = regions[0]
    rng = np.random.default_rng(5)
    r.copies = rng.integers(0, 200, size=(len(cfg.oligo_sizes), len(r.seq))).astype(np.int32)
    n = cfg.grid_size(r)
    idx = rng.choice(n, 500, replace=False).astype(np.int64)
    s0 = cfg.first_scan_start(r)
    for i, m in zip(idx, mg.describe_candidates(cfg, r, idx)):
        strand, q = int(i) & 1, int(i) >> 1
        p, q = q % cfg.n_pairs, q // cfg.n_pairs
        ci, si = q % len(cfg.captures), q // len(cfg.captures)
        e, l = cfg.ext_len[p], cfg.lig_len[p]
        s = s0 + si
        t = s + cfg.captures[ci] - e - l - 1
        want = dict(strand=strand, ext_len=e, lig_len=l, scan_start=s, scan_stop=t)
        if strand == 0:
            want.update(ext_start=s - e, ext_stop=s - 1, lig_start=t + 1, lig_stop=t + l)
        else:
            want.update(lig_start=s - l, lig_stop=s - 1, ext_start=t + 1, ext_stop=t + e)

        def copy(start, length):
            k = cfg.oligo_sizes.index(length)
            a = start - r.seq_start
            return int(r.copies[k, a]) if 0 <= a < len(r.seq) else 0

        want.update(ext_copy=copy(want["ext_start"], e), lig_copy=copy(want["lig_start"], l))
        assert m == want
    r.copies = None
    assert all(m["ext_copy"] == 1 and m["lig_copy"] == 1 for m in mg.describe_candidates(cfg, r, idx[:5]))
    with pytest.raises(mg.MgError):
        mg.describe_candidates(cfg, r, np.array([n], np.int64))
    with pytest.raises(mg.MgError):
        mg.describe_candidates(cfg, r, np.array([-1], np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(PLAIN_CASES))
def test_records_from_device_scores_equal_reference_files(oracle, case):
    g = json.load(open(os.path.join(GOLDEN_DIR, "records_%s.json" % case)))
    _flags, method, lower, upper, cfg = sp.CASES[case]
    d = tmpdir()
    model = sp.model_for(oracle, d)
    _genome, regions = sp.inputs(oracle, cfg)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    if method == 1:
        ctx.load_svr_model(model)
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_SVR if method == 1 else mg.MG_WANT_LOGISTIC)
    valid, lo, sv, _f = pnl.fetch(valid=True, logistic=method != 1, svr=method == 1)
    _so, _sb, po, pb = pnl.select(regions, method, lower, upper)  # condense + collapse on the device
    offs = pnl.offsets

    def score_of(i, r):
        a, b = offs[i], offs[i + 1]
        score = (sv if method == 1 else lo)[a:b]
        local = pb[po[i]:po[i + 1]]
        return valid[a:b], score, np.where(local >= 0, local - a, -1)  # panel-global -> region-local grid indices

    all_txt, col_txt = write_files(cfg, regions, method, lower, upper, score_of)
    assert file_digest(all_txt) == g["all_mips"], "all_mips.txt differs from the reference CLI's file"
    assert file_digest(col_txt) == g["collapsed_mips"], "collapsed_mips.txt differs from the reference CLI's file"


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/mipgen not built (needs /root/reference)")
@pytest.mark.parametrize("case", sorted(PLAIN_CASES))
def test_golden_is_what_the_reference_cli_writes(oracle, case):
    assert reference_files(oracle, case) == json.load(open(os.path.join(GOLDEN_DIR, "records_%s.json" % case)))


def reference_files(oracle, case):
    flags, _m, _lo, _up, cfg = sp.CASES[case]
    d = tmpdir()
    model = sp.model_for(oracle, d)
    genome = panel.lcg_genome(30000, sp.GENOME_SEED)
    gdir = os.path.join(d, "genome_" + case)
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    regions = panel.make_regions(genome, sp.N_REGIONS, 120, 170, cfg, sp.REGION_SEED)
    bed = os.path.join(d, case + ".bed")
    panel.write_bed(bed, regions)
    run, _log = run_cli(REF_CLI, d, "ref_" + case, bed, gdir, flags, model)
    return {"all_mips": file_digest(open(os.path.join(run, "p.all_mips.txt")).read()),
            "collapsed_mips": file_digest(open(os.path.join(run, "p.collapsed_mips.txt")).read())}


if __name__ == "__main__":
    from oracle_api import Oracle
    o = Oracle()
    for case in sorted(PLAIN_CASES):
        g = reference_files(o, case)
        json.dump(g, open(os.path.join(GOLDEN_DIR, "records_%s.json" % case), "w"), indent=1)
        print(case, g)
