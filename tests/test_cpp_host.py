"""The C++ `pansim` command line (pansim_b200/host): builds, mirrors the reference's
validation behaviour (print + exit 0, main.rs:194-247) without a GPU, and on a GPU writes
the reference's output files."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pansim_b200", "pansim")


@pytest.fixture(scope="module")
def binary():
    subprocess.run(["make", "-C", os.path.join(ROOT, "pansim_b200", "csrc")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", os.path.join(ROOT, "pansim_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def run(binary, *args):
    return subprocess.run([binary, *args], capture_output=True, text=True)


def test_validation_messages_and_exit_code_zero(binary):
    r = run(binary, "--core_genes", "7000")
    assert r.returncode == 0 and r.stdout == "core_genes must be less than or equal to pan_size\n"
    r = run(binary, "--HR_rate", "-1")
    assert r.returncode == 0 and r.stdout.splitlines() == ["HR_rate and HGT_rate must be above 0.0", "HR_rate: -1",
                                                           "HGT_rate: 0.05"]
    r = run(binary, "--n_gen", "0")
    assert r.returncode == 0 and r.stdout.splitlines()[0].startswith("pop_size, core_size, pan_genes, n_gen")
    r = run(binary, "--avg_gene_freq", "0")
    assert r.returncode == 0 and r.stdout.splitlines()[1] == "avg_gene_freq: 0"
    r = run(binary, "--help")
    assert r.returncode == 0 and "--HGT_rate" in r.stdout and "--no_control_genome_size" in r.stdout
    assert "--all_pairs" in r.stdout                          # extension flag


def test_no_gpu_fails_loudly(binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(binary, "--pop_size", "10", "--core_size", "100", "--n_gen", "1", "--max_distances", "5")
    assert r.returncode == 101 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_end_to_end(binary, tmp_path):
    pref = str(tmp_path / "cpp")
    r = run(binary, "--pop_size", "2e1", "--core_size", "700", "--pan_genes", "80", "--core_genes", "30", "--n_gen", "4",
            "--max_distances", "40", "--outpref", pref, "--print_dist", "--print_matrices", "--print_selection",
            "--prop_positive", "0.2", "--competition_strength", "0.5", "--verbose", "--seed", "3")
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    assert out[0].startswith("avg_gene_freq adjusted to ") and out[1] == "Finished gen: 1"
    rows = [l.split("\t") for l in open(pref + ".tsv").read().splitlines()]
    assert len(rows) == 40 and all(len(x) == 2 and 0.0 <= float(x[0]) <= 1.0 and 0.0 <= float(x[1]) <= 1.0 for x in rows)
    fr = open(pref + "_freqs.txt").read().splitlines()
    assert len(fr) == 80 and fr[-30:] == ["1"] * 30
    assert len(open(pref + "_per_gen.tsv").read().splitlines()) == 4
    assert len(open(pref + "_selection.tsv").read().splitlines()) == 50
    core = open(pref + "_core_genome.csv").read().splitlines()
    assert len(core) == 20 and all(len(l.split(",")) == 700 and set(l.split(",")) <= set("ACGT") for l in core)
    pan = [l.split(",") for l in open(pref + "_pangenome.csv").read().splitlines()]
    assert len(pan) == 20 and all(len(x) == 80 and x[:30] == ["1"] * 30 for x in pan)


@pytest.mark.gpu
def test_cpp_host_all_pairs_and_batched_run(binary, tmp_path):
    """--all_pairs (extension): <outpref>.tsv holds every pair i < j in (i, j) order. Without
    --print_dist / --verbose the generations run as one device-resident batch and must leave the
    same matrices as the generation-by-generation loop."""
    import numpy as np
    common = ["--pop_size", "21", "--core_size", "900", "--pan_genes", "70", "--core_genes", "10", "--n_gen", "4",
              "--max_distances", "30", "--print_matrices", "--seed", "5"]
    pa, pb = str(tmp_path / "a"), str(tmp_path / "b")
    assert run(binary, *common, "--outpref", pa, "--all_pairs").returncode == 0
    assert run(binary, *common, "--outpref", pb, "--print_dist").returncode == 0
    assert open(pa + "_core_genome.csv").read() == open(pb + "_core_genome.csv").read()
    assert open(pa + "_pangenome.csv").read() == open(pb + "_pangenome.csv").read()
    letters = np.array([l.split(",") for l in open(pa + "_core_genome.csv").read().splitlines()])
    pan = np.array([l.split(",") for l in open(pa + "_pangenome.csv").read().splitlines()]).astype(int)[:, 10:]
    rows = [l.split("\t") for l in open(pa + ".tsv").read().splitlines()]
    assert len(rows) == 21 * 20 // 2 and len(open(pb + ".tsv").read().splitlines()) == 30
    k = 0
    for i in range(21):
        for j in range(i + 1, 21):
            assert float(rows[k][0]) == (letters[i] != letters[j]).sum() / 900
            inter, uni = (pan[i] & pan[j]).sum(), (pan[i] | pan[j]).sum()
            assert float(rows[k][1]) == 1.0 - ((inter + 10.0) / (uni + 10.0))
            k += 1
