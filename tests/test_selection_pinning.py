"""Pin the oracle's restatement of the tile-loop replay + condense_mips + collapse_mips
(mipgen.cpp:421-501, 1670-1746, 1617-1649) against the reference CLI itself:

  * all_mips.txt      = the candidates the reference enumerated, in order, with their scores
  * collapsed_mips.txt = pos_strand_best_mip, positions ascending, '+' before '-'

Runs the unmodified reference binary (oracle/_ref/mipgen) where it has been built, and always
checks the committed golden key lists (tests/golden/selection_*.json) generated from it by this
file's `make_golden()` (python tests/test_selection_pinning.py)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mipgen_b200 import panel
from helpers import small_config, calibrated_model, tmpdir, GOLDEN_DIR
from cli_util import run_cli, read_rows, REF_CLI
import stub_rules

CASES = {
    # name: (cli flags, method, lower, upper, config)
    "logistic_3cap": (["-min_capture_size", "152", "-max_capture_size", "162", "-logistic_optimal_score", "0.9",
                       "-logistic_priority_score", "0.8", "-arm_length_sums", "40,43,45"], 0, 0.8, 0.9,
                      small_config((40, 43, 45), 162, 152, 5)),
    "svr_2cap": (["-min_capture_size", "157", "-max_capture_size", "162", "-score_method", "svr", "-svr_optimal_score", "0.728",
                  "-svr_priority_score", "0.6", "-arm_length_sums", "40,45"], 1, 0.6, 0.728, small_config((40, 45), 162, 157, 5)),
}
# cases that run against the rule-driven stubs (oracle/stub_bwa.sh with MIPGEN_STUB_RULES=1, stub_trf.sh, stub_tabix.sh):
# arm copy numbers other than 1, ambiguously mapping MIP starts, TRF-masked arms and SNPs in arms -- the selection-only
# inputs of condense_mips / collapse_mips (mipgen.cpp:1629, 1634-1645, 1689-1737)
STUB_CASES = {
    "logistic_stubs_2cap": (["-min_capture_size", "157", "-max_capture_size", "162", "-logistic_optimal_score", "0.9",
                             "-logistic_priority_score", "0.8", "-arm_length_sums", "40,45", "-trf", "trf"], 0, 0.8, 0.9,
                            small_config((40, 45), 162, 157, 5)),
    "svr_stubs_1cap": (["-min_capture_size", "162", "-max_capture_size", "162", "-score_method", "svr", "-svr_optimal_score", "0.728",
                        "-svr_priority_score", "0.6", "-arm_length_sums", "41,44", "-trf", "trf", "-masked_arm_threshold", "0.3"],
                       1, 0.6, 0.728, small_config((41, 44), 162, 162, 5)),
}
CASES.update(STUB_CASES)
GENOME_SEED, REGION_SEED, N_REGIONS = 8101, 8102, 2


def masked_threshold(case):
    flags = CASES[case][0]
    return float(flags[flags.index("-masked_arm_threshold") + 1]) if "-masked_arm_threshold" in flags else 0.5


def inputs(oracle, cfg):
    genome = panel.lcg_genome(30000, GENOME_SEED)
    regions = panel.make_regions(genome, N_REGIONS, 120, 170, cfg, REGION_SEED)
    for r in regions:
        r.lrc = oracle.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    return genome, regions


def model_for(oracle, d):
    cfg = small_config((40, 45))
    genome, regions = inputs(oracle, cfg)
    _v, _l, _s, feats = oracle.grid_region(regions[0], cfg, None, want_logistic=False, want_feats=True)
    sample = feats[np.isfinite(feats[:, 0])][::41]
    return calibrated_model(oracle, cfg, 48, 21, os.path.join(d, "mipgen_svr.model"), sample)


def key_of(cfg, r, idx):
    strand = idx & 1
    q = idx >> 1
    p = q % cfg.n_pairs
    q //= cfg.n_pairs
    ci = q % len(cfg.captures)
    si = q // len(cfg.captures)
    s = cfg.first_scan_start(r) + si
    e, l = cfg.ext_len[p], cfg.lig_len[p]
    t = s + cfg.captures[ci] - e - l - 1
    lo, hi = (s - e, t + l) if strand == 0 else (s - l, t + e)
    return "1:%d-%d/%d,%d/%s" % (lo, hi, e, l, "+-"[strand])


def digest(rows):
    """Compact fingerprint of a row list: count, SHA-256 over 'key<TAB>score' lines, first and last rows."""
    text = "".join("%s\t%s\n" % (k, v) for k, v in rows)
    return {"n": len(rows), "sha256": hashlib.sha256(text.encode()).hexdigest(), "head": [list(x) for x in rows[:5]],
            "tail": [list(x) for x in rows[-5:]]}


def oracle_rows(oracle, cfg, regions, model, method, lower, upper, masked_thr=0.5):
    """What the reference writes to all_mips.txt and collapsed_mips.txt, according to the oracle."""
    h = oracle.svm_load_model(model) if method == 1 else None
    all_rows, col_rows = [], []
    for r in regions:
        valid, lo, sv, _ = oracle.grid_region(r, cfg, h, want_logistic=method != 1, want_svr=method == 1)
        score = sv if method == 1 else lo
        enum_idx = oracle.tile_replay(r, cfg, valid, score, method, True, upper)
        all_rows += [(key_of(cfg, r, int(i)), "%g" % score[i]) for i in enum_idx]
        _sb, pb = oracle.select(r, cfg, score, enum_idx, lower, upper, masked_arm_threshold=masked_thr)
        for pos in range(pb.shape[0]):
            for strand in (0, 1):
                if pb[pos, strand] >= 0:
                    col_rows.append((key_of(cfg, r, int(pb[pos, strand])), "%g" % score[pb[pos, strand]]))
    if h:
        oracle.svm_free(h)
    return all_rows, col_rows


def reference_rows(case, d, model):
    flags, _m, _lo, _up, cfg = CASES[case]
    genome = panel.lcg_genome(30000, GENOME_SEED)
    gdir = os.path.join(d, "genome_" + case)
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    regions = panel.make_regions(genome, N_REGIONS, 120, 170, cfg, REGION_SEED)
    bed = os.path.join(d, case + ".bed")
    panel.write_bed(bed, regions)
    env = None
    if case in STUB_CASES:
        vcf = os.path.join(d, case + ".vcf")
        stub_rules.write_vcf(vcf, stub_rules.snp_positions(genome, regions))
        flags = flags + ["-snp_file", vcf]
        env = {"MIPGEN_STUB_RULES": "1"}
    run, _log = run_cli(REF_CLI, d, "ref_" + case, bed, gdir, flags, model, env_extra=env)
    return read_rows(os.path.join(run, "p.all_mips.txt")), read_rows(os.path.join(run, "p.collapsed_mips.txt"))


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_selection_matches_golden_from_reference_cli(oracle, case):
    g = json.load(open(os.path.join(GOLDEN_DIR, "selection_%s.json" % case)))
    _flags, method, lower, upper, cfg = CASES[case]
    d = tmpdir()
    model = model_for(oracle, d)
    genome, regions = inputs(oracle, cfg)
    if case in STUB_CASES:
        stub_rules.decorate(cfg, genome, regions, snps=stub_rules.snp_positions(genome, regions))
    all_rows, col_rows = oracle_rows(oracle, cfg, regions, model, method, lower, upper, masked_threshold(case))
    assert digest(all_rows) == g["all"], "enumeration order / replay / scores differ from the reference's all_mips.txt"
    assert digest(col_rows) == g["collapsed"], "condense+collapse differ from the reference's collapsed_mips.txt"
    if case in STUB_CASES:
        # the selection-only inputs must matter: without them condense/collapse pick different MIPs
        for r in regions:
            r.masked_seq = r.snp = r.unmappable = None
        _a, plain = oracle_rows(oracle, cfg, regions, model, method, lower, upper, masked_threshold(case))
        assert digest(plain) != g["collapsed"]
    else:
        n_valid = sum(int(oracle.grid_region(r, cfg, None)[0].sum()) for r in regions)
        assert g["all"]["n"] < n_valid * 0.97, "the case must exercise the score-dependent pruning"


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/mipgen not built (needs /root/reference)")
@pytest.mark.parametrize("case", sorted(CASES))
def test_golden_is_what_the_reference_cli_writes(oracle, case):
    d = tmpdir()
    model = model_for(oracle, d)
    all_rows, col_rows = reference_rows(case, d, model)
    g = json.load(open(os.path.join(GOLDEN_DIR, "selection_%s.json" % case)))
    assert digest(all_rows) == g["all"] and digest(col_rows) == g["collapsed"]


def make_golden():
    from oracle_api import Oracle
    o = Oracle()
    for case in CASES:
        d = tmpdir()
        model = model_for(o, d)
        a, c = reference_rows(case, d, model)
        json.dump({"all": digest(a), "collapsed": digest(c)}, open(os.path.join(GOLDEN_DIR, "selection_%s.json" % case), "w"), indent=1)
        print(case, len(a), "enumerated,", len(c), "collapsed rows")


if __name__ == "__main__":
    make_golden()
