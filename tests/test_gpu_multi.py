"""Multi-GPU: sharded results equal single-GPU results bit for bit.

* pansim_group (one process, ncclCommInitAll inside the library): a group of ONE device runs on any
  GPU box; the two-device tests need >= 2 GPUs.
* one process per GPU (torchrun) with pansim_comm_init_rank: tests/multi_gpu_worker.py.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import pansim_b200 as pb
from helpers import sample_pairs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _setup(seed=1):
    p = pb.Params(pop_size=96, core_size=8192 * 7 + 123, pan_genes=700, core_genes=200, n_gen=4, max_distances=500,
                  seed=5, prop_positive=0.1, competition_strength=0.3, HR_rate=0.5)
    d = pb.derive(p)
    rng = np.random.default_rng(seed)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < 0.25).astype(np.uint8)
    sel = rng.normal(0, 0.05, d.pan_size)
    r1, r2 = sample_pairs(rng, p.pop_size, p.max_distances)
    return p, core_row, acc_row, sel, r1, r2


def _reference_run(p, core_row, acc_row, sel, r1, r2):
    with pb.Pansim.from_params(p) as whole:
        whole.set_initial(core_row, acc_row)
        whole.set_selection(sel)
        stats = whole.run_generations_stats(0, p.n_gen, r1, r2)
        out = dict(core=whole.download_core(), acc=whole.download_acc(), counts=whole.pair_counts(r1, r2), stats=stats,
                   genes=whole.gene_counts(), csv=whole.export_core_csv(3, 9),
                   all_pairs=[(cd, it, un) for _, _, cd, it, un in whole.iter_all_pairs(chunk_pairs=700, with_indices=False)])
    return out


def _check_group(n_dev):
    p, core_row, acc_row, sel, r1, r2 = _setup()
    want = _reference_run(p, core_row, acc_row, sel, r1, r2)
    with pb.PansimGroup(p, n_dev) as g:
        assert g.size == n_dev
        g.set_initial(core_row, acc_row)
        g.set_selection(sel)
        stats = g.run_generations_stats(0, p.n_gen, r1, r2)
        assert stats.tolist() == want["stats"].tolist()
        assert (g.download_core() == want["core"]).all() and (g.download_acc() == want["acc"]).all()
        for a, b in zip(g.pair_counts(r1, r2), want["counts"]):
            assert (a == b).all()
        assert (g.gene_counts() == want["genes"]).all()
        assert g.export_core_csv(3, 9) == want["csv"]
        blocks = g.all_pairs(chunk_pairs=700)
        assert [b[0] for b in blocks][0] == 0 and blocks[-1][1] == p.pop_size - 1
        assert len(blocks) == len(want["all_pairs"]) > 3
        for (i0, i1, cd, it, un), (wcd, wit, wun) in zip(blocks, want["all_pairs"]):
            assert (cd == wcd).all() and (it == wit).all() and (un == wun).all()
        # a plain batch after the statistics batch keeps working (stream / buffer bookkeeping)
        g.run_generations(p.n_gen, 2)
    with pb.Pansim.from_params(p) as whole:
        whole.set_initial(core_row, acc_row)
        whole.set_selection(sel)
        whole.run_generations(0, p.n_gen + 2)
        with pb.PansimGroup(p, n_dev) as g2:
            g2.set_initial(core_row, acc_row)
            g2.set_selection(sel)
            g2.run_generations(0, p.n_gen + 2)
            assert (g2.download_core() == whole.download_core()).all()


def test_group_of_one_device_equals_single_context():
    _check_group(1)


def test_group_of_two_devices_equals_single_context():
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    _check_group(2)


def test_group_of_all_devices_equals_single_context():
    n = _n_gpus()
    if n < 3:
        pytest.skip("needs >= 3 GPUs")
    _check_group(min(n, 7))            # 7 regions + a ragged one: at most 8 non-empty shards


def test_column_sharded_run_equals_single_gpu():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
