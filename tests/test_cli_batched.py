"""The batched MIPgen driver (mipgen_b200/batched: the reference's mipgen.cpp with the per-feature candidate loop nest +
condense_mips + collapse_mips replaced by mg_tile_regions_multi at build time, pick_mips unchanged) against the unmodified
reference CLI (oracle/_ref/mipgen): every design file must be byte-identical -- all_mips, collapsed_mips, picked_mips (the
final probe set), snp_mips and the gap BED files -- in logistic, svr and mixed mode, incl. BASELINE cfg4's shape (mixed
scoring, -tag_sizes 4,4, capture sweep 120..250 step 5) and runs with arm copy numbers, ambiguous sites, TRF masks and SNPs
(rule-driven stub bwa / trf / tabix, oracle/stub_*.sh).

Both binaries are prebuilt where /root/reference exists (build()); on the GPU box they are only executed."""
import filecmp
import glob
import os
import time

import numpy as np
import pytest

from mipgen_b200 import panel
from helpers import small_config, calibrated_model, tmpdir
from cli_util import run_cli, REF_CLI
import stub_rules

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BATCHED_CLI = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen_batched")
needs_binaries = pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(BATCHED_CLI)),
                                    reason="reference / batched CLI not prebuilt (needs /root/reference at build time)")
DESIGN_FILES = ["all_mips.txt", "collapsed_mips.txt", "picked_mips.txt", "snp_mips.txt",
                # check_copy_numbers' inputs for BWA (mipgen.cpp:798-840): written on the device in the batched driver
                "all_sequences.fq", "oligo_copy_count.fq"]


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


@pytest.fixture(scope="module")
def ws(oracle):
    d = tmpdir()
    genome = panel.lcg_genome(120000, 9101)
    gdir = os.path.join(d, "genome")
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    cfg = panel.Config(250, 120, 5)
    beds = {}
    for name, n, lo, hi, seed in (("small3", 3, 50, 110, 1), ("cfg4_10", 10, 60, 140, 2), ("cfg4_big2", 2, 280, 320, 3), ("stubs2", 2, 120, 170, 4)):
        regs = panel.make_regions(genome, n, lo, hi, cfg, 9200 + seed)
        beds[name] = (os.path.join(d, name + ".bed"), regs)
        panel.write_bed(beds[name][0], regs)
    r0 = beds["small3"][1][0]
    c162 = small_config((40, 45))
    r0 = panel.cut_region(genome, r0.start_flanked, r0.stop_flanked, c162)
    r0.lrc = oracle.long_range_content(r0.flank_seq, r0.seq_start, r0.seq_stop)
    _v, _l, _s, feats = oracle.grid_region(r0, c162, None, want_logistic=False, want_feats=True)
    model = calibrated_model(oracle, c162, 64, 31, os.path.join(d, "mipgen_svr.model"), feats[np.isfinite(feats[:, 0])][::53])
    vcf = os.path.join(d, "snps.vcf")
    stub_rules.write_vcf(vcf, stub_rules.snp_positions(genome, [panel.cut_region(genome, r.start_flanked, r.stop_flanked, c162) for r in beds["stubs2"][1]]))
    return dict(dir=d, gdir=gdir, beds=beds, model=model, vcf=vcf)


# name: (bed, flags, environment for both binaries, extra environment for the batched binary only)
CASES = {
    "logistic": ("small3", ["-min_capture_size", "162", "-max_capture_size", "162"], {}, {}),
    "logistic_two_contexts_small_batches": ("small3", ["-min_capture_size", "152", "-max_capture_size", "162", "-logistic_heuristic", "off",
                                                       "-logistic_optimal_score", "0.9", "-logistic_priority_score", "0.8"], {},
                                            {"MIPGEN_B200_DEVICES": "0,0", "MIPGEN_B200_BATCH": "20000"}),
    "svr": ("small3", ["-min_capture_size", "152", "-max_capture_size", "162", "-score_method", "svr", "-arm_length_sums", "40,45"], {}, {}),
    "svr_low_threshold_silent": ("small3", ["-min_capture_size", "152", "-max_capture_size", "162", "-score_method", "svr", "-arm_length_sums", "40,45",
                                            "-svr_optimal_score", "1.9", "-svr_priority_score", "1.2", "-silent_mode", "on"], {}, {}),
    # BASELINE configs[3] at reduced scale: mixed scoring, smMIP tags, capture sweep 120..250 step 5
    "cfg4_mixed_tags_sweep": ("cfg4_10", ["-min_capture_size", "120", "-max_capture_size", "250", "-capture_increment", "5", "-score_method", "mixed",
                                          "-tag_sizes", "4,4"], {}, {}),
    "cfg4_mixed_tags_sweep_long_regions_silent": ("cfg4_big2", ["-min_capture_size", "120", "-max_capture_size", "250", "-capture_increment", "5",
                                                                "-score_method", "mixed", "-tag_sizes", "4,4", "-silent_mode", "on",
                                                                "-double_tile_strand_unaware", "on"], {}, {"MIPGEN_B200_BATCH": "1500000"}),
    # arm copies != 1, TRF-masked arms, SNPs in arms (all_mips.txt goes through design_mip objects)
    "stubs_copies_trf_snps": ("stubs2", ["-min_capture_size", "157", "-max_capture_size", "162", "-arm_length_sums", "40,45", "-trf", "trf",
                                         "-logistic_optimal_score", "0.9", "-logistic_priority_score", "0.8", "-snp_file", "@VCF@"],
                              {"MIPGEN_STUB_RULES": "2"}, {}),
    # ... plus ambiguously mapping MIP starts, svr, a lower masked-arm threshold
    "stubs_all_svr_silent": ("stubs2", ["-min_capture_size", "162", "-max_capture_size", "162", "-score_method", "svr", "-arm_length_sums", "41,44",
                                        "-trf", "trf", "-masked_arm_threshold", "0.3", "-snp_file", "@VCF@", "-silent_mode", "on",
                                        "-svr_optimal_score", "0.728", "-svr_priority_score", "0.6"], {"MIPGEN_STUB_RULES": "1"}, {}),
    "stubs_all_mixed_silent": ("stubs2", ["-min_capture_size", "152", "-max_capture_size", "162", "-score_method", "mixed", "-arm_length_sums", "40,43,45",
                                          "-trf", "trf", "-snp_file", "@VCF@", "-silent_mode", "on", "-seal_both_strands", "on"],
                               {"MIPGEN_STUB_RULES": "1"}, {}),
}


@needs_binaries
@pytest.mark.parametrize("case", sorted(CASES))
def test_batched_cli_writes_identical_design_files(ws, case):
    bed_name, flags, env, env_b = CASES[case]
    flags = [ws["vcf"] if f == "@VCF@" else f for f in flags]
    bed = ws["beds"][bed_name][0]
    t0 = time.perf_counter()
    ref_dir, _ = run_cli(REF_CLI, ws["dir"], "ref_" + case, bed, ws["gdir"], flags, ws["model"], env_extra=env)
    t1 = time.perf_counter()
    new_dir, log = run_cli(BATCHED_CLI, ws["dir"], "b200_" + case, bed, ws["gdir"], flags, ws["model"], env_extra=dict(env, **env_b))
    t2 = time.perf_counter()
    names = ["p." + f for f in DESIGN_FILES] + sorted(os.path.basename(p) for p in glob.glob(os.path.join(ref_dir, "p.*.bed")))
    for f in names:
        a, b = os.path.join(ref_dir, f), os.path.join(new_dir, f)
        assert os.path.exists(b), "%s missing in %s" % (f, case)
        assert filecmp.cmp(a, b, shallow=False), "%s differs in %s" % (f, case)
    assert os.path.getsize(os.path.join(ref_dir, "p.picked_mips.txt")) > 400
    if "-silent_mode" not in flags:
        assert os.path.getsize(os.path.join(ref_dir, "p.all_mips.txt")) > 10000
    line = [l for l in log.splitlines() if "batched driver" in l]
    assert line, "the batched driver must report its device batches:\n" + log[-800:]
    if "MIPGEN_B200_DEVICES" in env_b:
        assert "on 2 GPU(s)" in line[-1] and int(line[-1].split("batched driver:")[1].split()[0]) >= 2, line[-1]
    print("%s: reference CLI %.2f s, batched CLI %.2f s; %s" % (case, t1 - t0, t2 - t1, line[-1].split("] ")[-1]))
