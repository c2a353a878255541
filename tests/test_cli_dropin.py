"""End-to-end drop-in test: the UNCHANGED reference mipgen.cpp, compiled once against the
reference's own classes (oracle/_ref/mipgen) and once against the drop-in headers +
libmipgen_b200.so (mipgen_b200/dropin/_build/mipgen), must write byte-identical
all_mips / collapsed_mips / picked_mips / snp_mips files in logistic, svr and mixed mode.

Both binaries are prebuilt where /root/reference exists (build()); on the GPU box they are
only executed.  `bwa` is the stub of oracle/stub_bwa.sh (every arm copy = 1)."""
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

from mipgen_b200 import panel
from helpers import small_config, calibrated_model, tmpdir

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "mipgen")
STUB_DIR = os.path.join(ROOT, "oracle", "_ref")
DROPIN_CLI = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen")
OUTPUTS = ["all_mips.txt", "collapsed_mips.txt", "picked_mips.txt", "snp_mips.txt"]

needs_binaries = pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(DROPIN_CLI)),
                                    reason="reference / drop-in CLI not prebuilt (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def oracle():
    from oracle_api import Oracle
    return Oracle()


@pytest.fixture(scope="module")
def workspace(oracle):
    d = tmpdir()
    cfg = panel.Config(162, 152, 5)
    genome = panel.lcg_genome(40000, 9001)
    gdir = os.path.join(d, "genome")
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    regions = panel.make_regions(genome, 3, 50, 110, cfg, 9002)
    bed = os.path.join(d, "targets.bed")
    panel.write_bed(bed, regions)
    # a model calibrated on this genome so that the 1.5 / 2.2 thresholds are exercised
    r0 = regions[0]
    r0.lrc = oracle.long_range_content(r0.flank_seq, r0.seq_start, r0.seq_stop)
    _v, _l, _s, feats = oracle.grid_region(r0, cfg, None, want_logistic=False, want_feats=True)
    sample = feats[np.isfinite(feats[:, 0])][::53]
    model = calibrated_model(oracle, small_config((40, 45)), 64, 31, os.path.join(d, "mipgen_svr.model"), sample)
    return dict(dir=d, gdir=gdir, bed=bed, model=model)


def first_difference(a, b):
    """For assertion messages: the first line where two output files differ."""
    with open(a, errors="replace") as fa, open(b, errors="replace") as fb:
        for k, (x, y) in enumerate(zip(fa, fb)):
            if x != y:
                return "line %d:\n  ref: %s\n  new: %s" % (k + 1, x.rstrip()[:400], y.rstrip()[:400])
    return "one file is a prefix of the other"


def same_file(a, b, what):
    assert filecmp.cmp(a, b, shallow=False), "%s differs; %s" % (what, first_difference(a, b))


def run_cli(binary, ws, name, extra, env_extra=None):
    run = os.path.join(ws["dir"], name)
    os.makedirs(run)
    exe = os.path.join(run, "mipgen")
    os.symlink(binary, exe)                      # argv[0]'s directory is where the model is looked up
    shutil.copy(ws["model"], os.path.join(run, "mipgen_svr.model"))
    env = dict(os.environ, PATH=STUB_DIR + os.pathsep + os.environ.get("PATH", ""), MIPGEN_B200_VERBOSE="1")
    env.update(env_extra or {})
    cmd = [exe, "-regions_to_scan", ws["bed"], "-project_name", "p", "-bwa_genome_index", os.path.join(ws["gdir"], "chr1.fa"),
           "-genome_dir", ws["gdir"]] + extra
    r = subprocess.run(cmd, cwd=run, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return run, r.stderr


CASES = {
    "logistic": ["-min_capture_size", "162", "-max_capture_size", "162"],
    "logistic_multicap_noheur": ["-min_capture_size", "152", "-max_capture_size", "162", "-logistic_heuristic", "off",
                                 "-logistic_optimal_score", "0.9", "-logistic_priority_score", "0.8"],
    "svr": ["-min_capture_size", "152", "-max_capture_size", "162", "-score_method", "svr", "-arm_length_sums", "40,45"],
    "svr_low_threshold": ["-min_capture_size", "152", "-max_capture_size", "162", "-score_method", "svr", "-arm_length_sums", "40,45",
                          "-svr_optimal_score", "1.9", "-svr_priority_score", "1.2"],
    "mixed": ["-min_capture_size", "157", "-max_capture_size", "162", "-score_method", "mixed", "-arm_length_sums", "41,45",
              "-tag_sizes", "4,4"],
}


@needs_binaries
@pytest.mark.parametrize("mode", ["logistic", "svr"])
def test_dropin_cli_with_arm_copy_numbers_other_than_one(workspace, mode):
    """A bwa that reports extra copies for some arms (oracle/stub_bwa.sh, MIPGEN_STUB_RULES=2): the shim takes the region's copy
    table from the <project>.oligo_copy_count.sam find_copy wrote (mipgen.cpp:558-596), so candidates with copies != 1 are served
    from the device grid too instead of one launch per candidate; files stay byte-identical."""
    flags = CASES["logistic"] if mode == "logistic" else CASES["svr"]
    env = {"MIPGEN_STUB_RULES": "2"}
    ref_dir, _ = run_cli(REF_CLI, workspace, "ref_copies_" + mode, flags, env)
    new_dir, log = run_cli(DROPIN_CLI, workspace, "b200_copies_" + mode, flags, env)
    for f in OUTPUTS:
        same_file(os.path.join(ref_dir, "p." + f), os.path.join(new_dir, "p." + f), f)
    copies = [l.split("\t")[5] for l in open(os.path.join(ref_dir, "p.all_mips.txt")) if not l.startswith(">")]
    assert sum(c != "1" for c in copies) > 100, "the stub must have produced arm copies other than 1"
    line = [l for l in log.splitlines() if "device batches" in l][-1]
    explicit = int(line.split("explicit candidates")[1].split(",")[0])
    lookups = int(line.rsplit(" ", 1)[1])
    assert explicit == 0 and lookups > 1000, line


@needs_binaries
@pytest.mark.parametrize("case", sorted(CASES))
def test_dropin_cli_writes_identical_files(workspace, case):
    ref_dir, _ = run_cli(REF_CLI, workspace, "ref_" + case, CASES[case])
    new_dir, log = run_cli(DROPIN_CLI, workspace, "b200_" + case, CASES[case])
    for f in OUTPUTS:
        a, b = os.path.join(ref_dir, "p." + f), os.path.join(new_dir, "p." + f)
        assert os.path.getsize(a) > 200 or f == "snp_mips.txt", f
        same_file(a, b, "%s in %s" % (f, case))
    # the work really went through device batches, not per-candidate calls
    line = [l for l in log.splitlines() if "device batches" in l][-1]
    batches = int(line.split("device batches")[1].split(",")[0])
    lookups = int(line.rsplit(" ", 1)[1])
    assert lookups > 1000 and batches < lookups / 50, line
