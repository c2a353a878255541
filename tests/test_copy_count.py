"""SURVEY.md 8(f4), opt-in exact-match arm copy counting: the brute-force restatement against hand-counted answers (CPU), and
mg_genome_create / mg_count_arm_copies against the restatement (GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import copy_count as cc  # noqa: E402

from mipgen_b200 import panel  # noqa: E402
from mipgen_b200.panel import Region  # noqa: E402


def _region(seq: bytes, start: int = 1000) -> Region:
    return Region(seq=seq, seq_start=start, seq_stop=start + len(seq) - 1, start_flanked=start + 10, stop_flanked=start + len(seq) - 10)


def test_restatement_known_answers():
    genome = [b"ACGTACGTTTTT", b"GGGGACGA"]
    # "ACGT": forward at 0 and 4 of contig 0; it is its own reverse complement, so each locus counts on both strands -> 4
    t = cc.count_arm_copies(genome, b"ACGTAC", [4])
    assert t[0, 0] == 4
    # "CGTA": forward once (contig 0, offset 1); reverse complement "TACG" once (offset 3) -> 2
    assert t[0, 1] == 2
    # the start the reference never writes (i >= len - size) stays 0 = absent key
    assert t[0, 2] == 0 and t[0, 5] == 0
    # an oligo that is not in the genome, and one with a non-ACGT character: 100, find_copy's value for a read without X0
    assert cc.count_arm_copies(genome, b"CACACA", [4])[0, 0] == 100
    assert cc.count_arm_copies(genome, b"ACNTAC", [4])[0, 0] == 100
    # lower case is matched like upper case (BWA indexes are case-blind); N separates
    assert cc.count_arm_copies([b"acgtNacgt"], b"ACGTA", [4])[0, 0] == 4
    # k-mers across a contig boundary do not exist
    assert cc.count_arm_copies([b"AAAC", b"GTTT"], b"ACGTT", [4])[0, 0] == 100
    # "AAAC" + its reverse complement "GTTT"
    assert cc.count_arm_copies([b"AAAC", b"GTTT"], b"AAACG", [4])[0, 0] == 2


def _synthetic_genome(rng: np.random.Generator):
    """Three contigs with planted repeats, an inverted repeat, N runs, lower case and a contig shorter than 32 bases."""
    a = bytearray(panel.lcg_genome(60000, 5))
    rep = bytes(a[1000:1060])
    for at in (7000, 23000, 41000):
        a[at:at + 60] = rep                       # direct repeats of a 60-mer
    a[30000:30060] = cc.revcomp(rep)              # the same on the other strand
    a[12000:12040] = b"N" * 40                    # an N run
    a[12500] = ord("N")
    a[50000:50200] = bytes(a[50000:50200]).lower()
    a[55000:55030] = b"AC" * 15                   # low complexity
    b = bytearray(panel.lcg_genome(5000, 9))
    b[100:160] = rep
    b[4990:5000] = bytes(a[1000:1010])            # a repeat that runs into the contig end (short suffixes)
    c = bytearray(b"ACGTTGCAAGGCTTAACCGGTTAA")   # 24 bases: only short suffixes
    del rng
    return [bytes(a), bytes(b), bytes(c)], rep


@pytest.mark.gpu
def test_device_counts_equal_the_restatement():
    import mipgen_b200 as mg
    rng = np.random.default_rng(3)
    genome, rep = _synthetic_genome(rng)
    ctx = mg.Context(0)
    g = ctx.genome([s.decode() for s in genome])
    pos, indexed, short = g.info()
    assert pos == sum(len(s) for s in genome) and indexed + short <= pos and short > 0
    sizes = [16, 17, 20, 24, 29, 30, 32, 1, 5]
    a = genome[0]
    regions = [
        _region(a[900:1200]),                     # holds the planted repeat
        _region(a[11950:12600]),                  # N run inside
        _region(a[29950:30120]),                  # the inverted copy
        _region(a[49900:50300].upper()),          # region upper-cased by the caller, genome lower case
        _region(a[54950:55100]),                  # low complexity
        _region(genome[1][4900:5000]),            # contig end
        _region(genome[2]),                       # the short contig
        _region(b"ACGTNRYACGTACGTAC-GTACGTACGTTTACGATCGATCGATCGATTTACGACGATC"),   # not from the genome, odd characters
        _region(b""),                             # empty region
        _region(b"ACGT"),                         # shorter than most oligo sizes
    ]
    got = g.count_arm_copies(regions, sizes)
    tables = {s: cc.kmer_table(genome, s) for s in sizes}
    for r, t in zip(regions, got):
        want = cc.count_arm_copies(genome, r.seq, sizes, tables)
        assert t.shape == want.shape
        assert np.array_equal(t, want)
    # the planted 60-mer: 5 direct copies (4 in contig 0 incl. the original, 1 in contig 1) + 1 inverted = 6 for every 30-mer inside it
    k30 = sizes.index(30)
    assert (got[0][k30, 100:131] == 6).all()
    # tables are usable as mg_region.copies: same layout
    assert got[0].dtype == np.int32 and got[0].shape == (len(sizes), len(regions[0].seq))
    # sizes beyond 32 bases are refused, not truncated
    with pytest.raises(mg.MgError):
        g.count_arm_copies(regions[:1], [33])
    g.close()
    ctx.close()


@pytest.mark.gpu
def test_copy_tables_feed_the_scorer(tmp_path):
    """End of the opt-in route: copies counted on the device go into mg_region.copies and change the scores exactly as the
    reference's copy terms do (oracle.grid_region with the same tables)."""
    import mipgen_b200 as mg
    from helpers import small_config, random_model, rel_err
    from oracle_api import Oracle
    oracle = Oracle()
    cfg = small_config((40, 41))
    genome = bytearray(panel.lcg_genome(panel.genome_length_for(2, 80, cfg), 77))
    regions = panel.make_regions(bytes(genome), 2, 60, 80, cfg, 78)
    r0 = regions[0]
    # duplicate a stretch of region 0 elsewhere in the genome so that some arms have copy 2
    src = r0.seq_start - 1 + 200
    genome[100:160] = genome[src:src + 60]
    regions = panel.make_regions(bytes(genome), 2, 60, 80, cfg, 78)
    for r in regions:
        r.lrc = oracle.long_range_content(r.flank_seq, r.seq_start, r.seq_stop)
    ctx = mg.Context(0)
    ctx.set_config(cfg)
    g = ctx.genome([bytes(genome).decode()])
    tabs = g.count_arm_copies(regions, cfg.oligo_sizes)
    assert any((t == 2).any() for t in tabs)
    for r, t in zip(regions, tabs):
        r.copies = t
    model = random_model(oracle, cfg, 48, 2, str(tmp_path / "m.model"))
    ctx.load_svr_model(model)
    _o, valid, lo, sv, _f = ctx.score_regions(regions, mg.MG_WANT_LOGISTIC | mg.MG_WANT_SVR)
    h = oracle.svm_load_model(model)
    want = [oracle.grid_region(r, cfg, h, want_logistic=True, want_svr=True) for r in regions]
    oracle.svm_free(h)
    assert np.array_equal(valid, np.concatenate([w[0] for w in want]))
    assert rel_err(lo, np.concatenate([w[1] for w in want])) <= 1e-12
    assert rel_err(sv, np.concatenate([w[2] for w in want])) <= 1e-9
    g.close()
    ctx.close()


@pytest.mark.gpu
def test_batched_cli_with_exact_copies_equals_reference_fed_the_same_x0(tmp_path):
    """Route C with MIPGEN_B200_EXACT_COPIES=1 (find_copy answered by mg_count_arm_copies, no bwa run on the oligo reads) against the
    unmodified reference CLI whose stub bwa replays, read by read, the exact-match counts of the brute-force restatement as X0
    tags: every design file byte-identical.  The genome holds a direct and an inverted copy of a stretch of one region, so arms
    with copy 2 and 3 exist and move the logistic scores (SVMipv4.cpp:173-174)."""
    import filecmp
    from cli_util import run_cli, REF_CLI
    batched = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen_batched")
    if not (os.path.exists(REF_CLI) and os.path.exists(batched)):
        pytest.skip("reference / batched CLI not prebuilt (needs /root/reference at build time)")
    d = str(tmp_path)
    cfg = panel.Config(162, 157, 5)
    genome = bytearray(panel.lcg_genome(60000, 4242))
    regs = panel.make_regions(bytes(genome), 2, 120, 170, cfg, 4243)
    src = regs[0].start_flanked - 1
    genome[400:480] = genome[src:src + 80]
    genome[700:760] = cc.revcomp(bytes(genome[src + 20:src + 80]))
    genome = bytes(genome)
    gdir = os.path.join(d, "genome")
    os.makedirs(gdir)
    panel.write_fasta(os.path.join(gdir, "chr1.fa"), "chr1", genome)
    bed = os.path.join(d, "r.bed")
    panel.write_bed(bed, regs)
    flags = ["-min_capture_size", "157", "-max_capture_size", "162", "-arm_length_sums", "40,45", "-logistic_optimal_score", "0.9",
             "-logistic_priority_score", "0.8"]
    # 1. the reads the CLI asks BWA about
    probe, _ = run_cli(REF_CLI, d, "probe", bed, gdir, flags)
    lines = open(os.path.join(probe, "p.oligo_copy_count.fq")).read().split("\n")
    names, seqs = lines[0::4], lines[1::4]
    tables = {}
    x0 = os.path.join(d, "x0.tsv")
    n_multi = 0
    with open(x0, "w") as f:
        for name, s in zip(names, seqs):
            if not name:
                continue
            q = s.encode()
            if len(q) not in tables:
                tables[len(q)] = cc.kmer_table([genome], len(q))
            tab = tables[len(q)]
            n = tab.get(q, 0) + tab.get(cc.revcomp(q), 0)
            n_multi += n > 1
            f.write("%s\t%d\n" % (name[1:], n))
    assert n_multi > 50
    # 2. the reference with those X0 tags, the batched driver counting on the device (its stub bwa would say 1 everywhere)
    ref_dir, _ = run_cli(REF_CLI, d, "ref", bed, gdir, flags, env_extra={"MIPGEN_STUB_RULES": "3", "MIPGEN_STUB_X0_FILE": x0})
    new_dir, log = run_cli(batched, d, "b200", bed, gdir, flags, env_extra={"MIPGEN_B200_EXACT_COPIES": "1"})
    for name in ("p.all_mips.txt", "p.collapsed_mips.txt", "p.picked_mips.txt"):
        assert filecmp.cmp(os.path.join(ref_dir, name), os.path.join(new_dir, name), shallow=False), name + " differs"
    # the copies did matter: a run with copy 1 everywhere scores differently
    assert not filecmp.cmp(os.path.join(probe, "p.all_mips.txt"), os.path.join(ref_dir, "p.all_mips.txt"), shallow=False)


def test_stub_bwa_replays_x0_tags_from_a_file(tmp_path):
    """oracle/stub_bwa.sh, MIPGEN_STUB_RULES=3: the X0 tag of an arm read comes from MIPGEN_STUB_X0_FILE, anything else stays 1."""
    import subprocess
    fq = tmp_path / "r.fq"
    fq.write_text("@chr1:100-115\nACGTACGTACGTACGT\n+\n################\n@chr1:101-116\nCGTACGTACGTACGTA\n+\n################\n")
    x0 = tmp_path / "x0.tsv"
    x0.write_text("chr1:100-115\t7\n")
    env = dict(os.environ, MIPGEN_STUB_RULES="3", MIPGEN_STUB_X0_FILE=str(x0))
    out = subprocess.run(["bash", os.path.join(ROOT, "oracle", "stub_bwa.sh"), "samse", "idx", "sai", str(fq)], env=env, capture_output=True, text=True)
    assert out.returncode == 0
    lines = out.stdout.strip().split("\n")
    assert len(lines) == 2 and lines[0].startswith("chr1:100-115\t") and "X0:i:7\t" in lines[0] and "X0:i:1\t" in lines[1]
