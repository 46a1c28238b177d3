"""PSEV event-dump format: writer/reader round trip (CPU) and replay from a dump file (GPU)."""
import numpy as np
import pytest

import pansim_b200 as pb
from pansim_b200 import event_dump
from oracle import binding as ob


def _make_dump(path, N, L, G, gens, seed):
    rng = np.random.default_rng(seed)
    core0 = (1 << rng.integers(0, 4, (N, L))).astype(np.uint8)
    acc0 = (rng.random((N, G)) < 0.3).astype(np.uint8)
    core, pan = ob.Population(core0, True, 3), ob.Population(acc0, False, 3)
    main = ob.make_rng(seed)
    with open(path, "wb") as f:
        for g in range(gens):
            parents = rng.integers(0, N, N).astype(np.uint32)
            core.next_generation(parents)
            pan.next_generation(parents)
            ev = ob.EventLog()
            core.mutate_alleles([L * 0.05], [(0, L)], seed, g, ev)
            pan.mutate_alleles([G * 0.4], [(0, G)], seed, g, ev)
            core.recombine([L * 0.02], [(0, L)], main, seed, g, ev)
            pan.recombine([G * 0.1], [(0, G)], main, seed, g, ev)
            event_dump.write_generation(f, g, parents, ev.arrays())
    return core0, acc0, core.m, pan.m


def test_round_trip(tmp_path):
    path = str(tmp_path / "run.psev")
    _make_dump(path, 8, 500, 40, 3, 1)
    recs = list(event_dump.read_generations(path))
    assert [r[0] for r in recs] == [0, 1, 2]
    for _g, parents, ev in recs:
        assert parents.dtype == np.uint32 and len(parents) == 8
        assert len(ev["core_mut_row"]) == len(ev["core_mut_site"]) == len(ev["core_mut_allele"]) > 0
        assert len(ev["hr_recipient"]) == len(ev["hr_value"]) > 0
    with open(path, "ab") as f:
        f.write(b"junk")
    with pytest.raises(ValueError):
        list(event_dump.read_generations(path))


@pytest.mark.gpu
def test_replay_from_dump_file(tmp_path):
    path = str(tmp_path / "run.psev")
    N, L, G = 20, 9000, 120
    core0, acc0, core_final, acc_final = _make_dump(path, N, L, G, 3, 2)
    with pb.Pansim.from_params(pb.Params(pop_size=N, core_size=L, pan_genes=G + 3, core_genes=3)) as sim:
        sim.upload(core0, acc0)
        assert event_dump.replay(sim, path) == 3
        assert (sim.download_core() == core_final).all()
        assert (sim.download_acc() == acc_final).all()
