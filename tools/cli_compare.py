#!/usr/bin/env python3
"""Wall-clock comparison of the unmodified reference CLI (oracle/_ref/mipgen) and the drop-in CLI
(mipgen_b200/dropin/_build/mipgen: the same mipgen.cpp against the GPU library) on one synthetic panel.
Outputs must be byte-identical; prints the timings as JSON.   python tools/cli_compare.py [n_regions] [n_sv]
"""
import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mipgen_b200 import panel  # noqa: E402
from oracle_api import Oracle  # noqa: E402
from helpers import calibrated_model, small_config  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "mipgen")
NEW = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen")
STUB = os.path.join(ROOT, "oracle", "_ref")


def run(binary, d, name, bed, gdir, model, extra):
    rd = os.path.join(d, name)
    os.makedirs(rd)
    os.symlink(binary, os.path.join(rd, "mipgen"))
    shutil.copy(model, os.path.join(rd, "mipgen_svr.model"))
    env = dict(os.environ, PATH=STUB + os.pathsep + os.environ["PATH"], MIPGEN_B200_VERBOSE="1")
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(rd, "mipgen"), "-regions_to_scan", bed, "-project_name", "p", "-bwa_genome_index",
                        os.path.join(gdir, "chr1.fa"), "-genome_dir", gdir, "-min_capture_size", "162", "-max_capture_size", "162",
                        "-silent_mode", "on"] + extra, cwd=rd, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-1500:]
    # time of the tile phase: from the first "[mipgen] feature #" line on is not separable here; report total
    return rd, dt, [l for l in r.stderr.splitlines() if "device batches" in l]


def main():
    n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n_sv = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    oracle = Oracle()
    d = tempfile.mkdtemp(prefix="mipgen_cli_")
    cfg = panel.Config()
    genome = panel.lcg_genome(panel.genome_length_for(n_regions, 300, cfg), 555)
    gdir = os.path.join(d, "genome")
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    regions = panel.make_regions(genome, n_regions, 80, 300, cfg, 556)
    bed = os.path.join(d, "t.bed")
    panel.write_bed(bed, regions)
    r0 = regions[0]
    r0.lrc = oracle.long_range_content(r0.flank_seq, r0.seq_start, r0.seq_stop)
    _v, _l, _s, feats = oracle.grid_region(r0, cfg, None, want_logistic=False, want_feats=True)
    sample = feats[np.isfinite(feats[:, 0])][::211]
    model = calibrated_model(oracle, small_config((40, 45)), n_sv, 3, os.path.join(d, "mipgen_svr.model"), sample)
    n_cand = sum(cfg.grid_size(r) for r in regions)
    out = {"regions": n_regions, "candidates": n_cand, "n_sv": n_sv}
    for mode, nreg in (("logistic", n_regions), ("svr", max(2, n_regions // 10))):
        bed_m = bed
        if nreg != n_regions:
            bed_m = os.path.join(d, "t_%s.bed" % mode)
            panel.write_bed(bed_m, regions[:nreg])
        extra = ["-score_method", mode]
        a, ta, _ = run(REF, d, "ref_" + mode, bed_m, gdir, model, extra)
        b, tb, log = run(NEW, d, "b200_" + mode, bed_m, gdir, model, extra)
        same = all(filecmp.cmp(os.path.join(a, "p." + f), os.path.join(b, "p." + f), shallow=False)
                   for f in ("picked_mips.txt", "collapsed_mips.txt", "snp_mips.txt"))
        out[mode] = {"regions": nreg, "candidates": sum(cfg.grid_size(r) for r in regions[:nreg]), "reference_s": round(ta, 2),
                     "dropin_s": round(tb, 2), "identical_outputs": same, "shim": log[-1] if log else ""}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
