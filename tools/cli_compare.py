#!/usr/bin/env python3
"""Wall-clock comparison of the unmodified reference CLI (oracle/_ref/mipgen), the drop-in CLI
(mipgen_b200/dropin/_build/mipgen: the same mipgen.cpp against the GPU library) and the batched caller
(INTEGRATION.md route B: mg_panel_score + mg_panel_select + mg_format_mip_records, which writes collapsed_mips.txt
itself) on one synthetic panel.  Outputs must be byte-identical; prints the timings as JSON.
    python tools/cli_compare.py [n_regions] [n_sv] [svr_regions] [quick]
`quick` times only the silent reference run and the batched driver (twice: the first process on an idle GPU also pays the driver's
device initialisation) -- the byte-identity of every output file, silent or not, is what tests/test_cli_batched.py checks.
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
BATCHED = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen_batched")   # route C: loop nest + condense + collapse on the GPU
STUB = os.path.join(ROOT, "oracle", "_ref")


def run(binary, d, name, bed, gdir, model, extra, silent=True):
    rd = os.path.join(d, name)
    os.makedirs(rd)
    os.symlink(binary, os.path.join(rd, "mipgen"))
    shutil.copy(model, os.path.join(rd, "mipgen_svr.model"))
    env = dict(os.environ, PATH=STUB + os.pathsep + os.environ["PATH"], MIPGEN_B200_VERBOSE="1")
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(rd, "mipgen"), "-regions_to_scan", bed, "-project_name", "p", "-bwa_genome_index",
                        os.path.join(gdir, "chr1.fa"), "-genome_dir", gdir, "-min_capture_size", "162", "-max_capture_size", "162"] +
                       (["-silent_mode", "on"] if silent else []) + extra, cwd=rd, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-1500:]
    # time of the tile phase: from the first "[mipgen] feature #" line on is not separable here; report total
    return rd, dt, [l for l in r.stderr.splitlines() if "device batches" in l]


HEADER = (">mip_key\t%s_score\tchr\text_probe_start\text_probe_stop\text_probe_copy\text_probe_sequence\tlig_probe_start\t"
          "lig_probe_stop\tlig_probe_copy\tlig_probe_sequence\tmip_scan_start_position\tmip_scan_stop_position\t"
          "scan_target_sequence\tmip_sequence\tfeature_start_position\tfeature_stop_position\tprobe_strand\tfailure_flags\tmip_name\n")


def batched(cfg, regions, mode, model, out_all, out_collapsed):
    """all_mips.txt and collapsed_mips.txt through the batched caller; returns the stage timings in seconds."""
    import mipgen_b200 as mg
    t = {}
    t0 = time.perf_counter()
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    method = 1 if mode == "svr" else 0
    if method:
        ctx.load_svr_model(model)
    t["context_and_model_s"] = time.perf_counter() - t0
    t1 = time.perf_counter()
    for r in regions:
        r.lrc = ctx.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    pnl = ctx.panel(regions)
    pnl.score(mg.MG_WANT_SVR if method else mg.MG_WANT_LOGISTIC)
    valid, lo, sv, _f = pnl.fetch(valid=True, logistic=not method, svr=bool(method))
    lower, upper = (1.5, 2.2) if method else (0.9, 0.98)  # mipgen.cpp:210-216
    _so, _sb, po, pb = pnl.select(regions, method, lower, upper)
    ctx.sync()
    t["upload_score_select_fetch_s"] = time.perf_counter() - t1
    t2 = time.perf_counter()
    score = sv if method else lo
    offs = pnl.offsets
    n = n_all = 0
    f_col, f_all = open(out_collapsed, "wb"), open(out_all, "wb")
    f_col.write((HEADER % mode).encode())
    f_all.write((HEADER % mode).encode())
    for i, r in enumerate(regions):
        a, b = offs[i], offs[i + 1]
        local = pb[po[i]:po[i + 1]].reshape(-1)
        winners = local[local >= 0] - a
        f_col.write(mg.design_records(cfg, r, winners, score[a:b], "1", r.label, r.start_flanked, r.stop_flanked, n + 1, raw=True))
        n += winners.size
        # all_mips.txt: every candidate the tile loop enumerates, in its order (mipgen.cpp:426-497)
        enum_idx = mg.tile_replay(cfg, r, valid[a:b], score[a:b], method, True, upper)
        f_all.write(mg.design_records(cfg, r, enum_idx, score[a:b], "1", r.label, r.start_flanked, r.stop_flanked, n_all + 1, raw=True))
        n_all += enum_idx.size
    f_col.close()
    f_all.close()
    t["format_and_write_s"] = time.perf_counter() - t2
    t["total_s"] = time.perf_counter() - t0
    t["collapsed_records"] = n
    t["all_mips_records"] = n_all
    return {k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()}


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
    svr_regions = int(sys.argv[3]) if len(sys.argv) > 3 else max(2, n_regions // 10)
    quick = len(sys.argv) > 4 and sys.argv[4] == "quick"
    for mode, nreg in (("logistic", n_regions), ("svr", svr_regions)):
        bed_m = bed
        if nreg != n_regions:
            bed_m = os.path.join(d, "t_%s.bed" % mode)
            panel.write_bed(bed_m, regions[:nreg])
        extra = ["-score_method", mode]
        a, ta, _ = run(REF, d, "ref_" + mode, bed_m, gdir, model, extra)
        if quick:
            runs = []
            for rep in range(2):
                e, te, elog = run(BATCHED, d, "batched_%s_%d" % (mode, rep), bed_m, gdir, model, extra)
                same_c = all(filecmp.cmp(os.path.join(a, "p." + f), os.path.join(e, "p." + f), shallow=False)
                             for f in ("picked_mips.txt", "collapsed_mips.txt", "snp_mips.txt"))
                runs.append({"wall_s": round(te, 2), "speedup": round(ta / te, 1), "identical_picked_collapsed_snp": same_c,
                             "driver_report": elog[-1] if elog else ""})
            out[mode + "_batched_cli"] = {"regions": nreg, "candidates": sum(cfg.grid_size(r) for r in regions[:nreg]),
                                          "reference_silent_s": round(ta, 2), "batched_silent_runs": runs}
            continue
        b, tb, log = run(NEW, d, "b200_" + mode, bed_m, gdir, model, extra)
        same = all(filecmp.cmp(os.path.join(a, "p." + f), os.path.join(b, "p." + f), shallow=False)
                   for f in ("picked_mips.txt", "collapsed_mips.txt", "snp_mips.txt"))
        # default (non-silent) run of the reference: it also writes all_mips.txt and the collapsed records
        c, tc, _ = run(REF, d, "ref_full_" + mode, bed_m, gdir, model, extra, silent=False)
        f_all, f_col = os.path.join(d, "batched_%s_all.txt" % mode), os.path.join(d, "batched_%s_collapsed.txt" % mode)
        bt = batched(cfg, regions[:nreg], mode, model, f_all, f_col)
        same_b = (filecmp.cmp(os.path.join(c, "p.collapsed_mips.txt"), f_col, shallow=False) and
                  filecmp.cmp(os.path.join(c, "p.all_mips.txt"), f_all, shallow=False))
        # route C: the batched driver (reference mipgen.cpp patched at build time), same flags, silent and not
        e, te, elog = run(BATCHED, d, "batched_" + mode, bed_m, gdir, model, extra)
        same_c = all(filecmp.cmp(os.path.join(a, "p." + f), os.path.join(e, "p." + f), shallow=False)
                     for f in ("picked_mips.txt", "collapsed_mips.txt", "snp_mips.txt"))
        e2, te2, _ = run(BATCHED, d, "batched_full_" + mode, bed_m, gdir, model, extra, silent=False)
        same_c2 = all(filecmp.cmp(os.path.join(c, "p." + f), os.path.join(e2, "p." + f), shallow=False)
                      for f in ("picked_mips.txt", "collapsed_mips.txt", "snp_mips.txt", "all_mips.txt"))
        out[mode + "_batched_cli"] = {"regions": nreg, "reference_silent_s": round(ta, 2), "batched_silent_s": round(te, 2),
                                      "speedup_silent": round(ta / te, 1), "identical_picked_collapsed_snp": same_c,
                                      "reference_not_silent_s": round(tc, 2), "batched_not_silent_s": round(te2, 2),
                                      "speedup_not_silent": round(tc / te2, 1), "identical_all_files_not_silent": same_c2,
                                      "driver_report": elog[-1] if elog else ""}
        out[mode] = {"regions": nreg, "candidates": sum(cfg.grid_size(r) for r in regions[:nreg]), "reference_s": round(ta, 2),
                     "dropin_s": round(tb, 2), "identical_outputs": same, "shim": log[-1] if log else "",
                     "reference_not_silent_s": round(tc, 2), "batched_caller": bt,
                     "batched_all_and_collapsed_identical_to_reference": same_b,
                     "all_mips_bytes": os.path.getsize(f_all)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
