"""Helpers to run a MIPgen CLI binary (reference or drop-in) on synthetic inputs with the stub bwa."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "mipgen")
STUB_DIR = os.path.join(ROOT, "oracle", "_ref")
DROPIN_CLI = os.path.join(ROOT, "mipgen_b200", "dropin", "_build", "mipgen")


def run_cli(binary, workdir, name, bed, gdir, extra, model=None, timeout=900, env_extra=None):
    run = os.path.join(workdir, name)
    os.makedirs(run)
    exe = os.path.join(run, "mipgen")
    os.symlink(binary, exe)  # argv[0]'s directory is where mipgen_svr.model is looked up (mipgen.cpp:137-138, 409)
    if model:
        shutil.copy(model, os.path.join(run, "mipgen_svr.model"))
    env = dict(os.environ, PATH=STUB_DIR + os.pathsep + os.environ.get("PATH", ""), MIPGEN_B200_VERBOSE="1")
    env.update(env_extra or {})
    cmd = [exe, "-regions_to_scan", bed, "-project_name", "p", "-bwa_genome_index", os.path.join(gdir, "chr1.fa"),
           "-genome_dir", gdir] + extra
    r = subprocess.run(cmd, cwd=run, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return run, r.stderr


def read_rows(path):
    """(mip_key, score string) per data row of an all/collapsed/picked_mips file."""
    rows = []
    for line in open(path, errors="replace"):  # masking_failed is an uninitialised byte after a mapping failure (mipgen.cpp:622-624)
        if line.startswith(">") or not line.strip():
            continue
        f = line.rstrip("\n").split("\t")
        rows.append((f[0], f[1]))
    return rows
