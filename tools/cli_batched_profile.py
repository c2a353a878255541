#!/usr/bin/env python3
"""Where the batched driver (route C) spends its wall clock: the cli_compare panel through mipgen_batched only, with the driver's own
MIPGEN_B200_VERBOSE report.   python tools/cli_batched_profile.py [n_regions] [n_sv]"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cli_compare as cc  # noqa: E402
from mipgen_b200 import panel  # noqa: E402
from oracle_api import Oracle  # noqa: E402
from helpers import calibrated_model, small_config  # noqa: E402


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
    out = {}
    for rep in range(2):
        for mode in ("logistic", "svr"):
            for silent in (True, False):
                _rd, dt, log = cc.run(cc.BATCHED, d, "b_%s_%d_%d" % (mode, silent, rep), bed, gdir, model, ["-score_method", mode], silent=silent)
                out["%s_%s_run%d" % (mode, "silent" if silent else "full", rep)] = {"wall_s": round(dt, 3), "driver": log[-1] if log else ""}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
