"""Binary event-dump format ("PSEV", version 1) for replay mode.

One record per generation, little-endian, everything the apply step consumes in the
order main.rs:445-464 consumes it (flat lists in APPLY ORDER, see include/pansim_b200.h
`pansim_events`). The same format is written by `integration/event_dump.rs` (the module
a maintainer adds to the Rust reference to dump a real run) and by the oracle-side tests.

    magic  b"PSEV"   u32 version (=1)   u32 generation   u32 pop_size
    u64 n_core_mut   u64 n_acc_flip     u64 n_hr         u64 n_hgt
    u32 parents[pop_size]
    u32 core_mut_row[n_core_mut]   u32 core_mut_site[n_core_mut]   u8 core_mut_allele[n_core_mut]
    u32 acc_flip_row[n_acc_flip]   u32 acc_flip_gene[n_acc_flip]
    u32 hr_recipient[n_hr]         u32 hr_locus[n_hr]              u8 hr_value[n_hr]
    u32 hgt_recipient[n_hgt]       u32 hgt_gene[n_hgt]
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"PSEV"
VERSION = 1
_FIELDS = [("core_mut_row", "u4", "n_core_mut"), ("core_mut_site", "u4", "n_core_mut"),
           ("core_mut_allele", "u1", "n_core_mut"), ("acc_flip_row", "u4", "n_acc_flip"),
           ("acc_flip_gene", "u4", "n_acc_flip"), ("hr_recipient", "u4", "n_hr"), ("hr_locus", "u4", "n_hr"),
           ("hr_value", "u1", "n_hr"), ("hgt_recipient", "u4", "n_hgt"), ("hgt_gene", "u4", "n_hgt")]


def write_generation(f, gen: int, parents, ev: dict) -> None:
    parents = np.ascontiguousarray(parents, "<u4")
    n = dict(n_core_mut=len(ev.get("core_mut_row", ())), n_acc_flip=len(ev.get("acc_flip_row", ())),
             n_hr=len(ev.get("hr_recipient", ())), n_hgt=len(ev.get("hgt_recipient", ())))
    f.write(MAGIC + struct.pack("<III", VERSION, gen, len(parents)))
    f.write(struct.pack("<QQQQ", n["n_core_mut"], n["n_acc_flip"], n["n_hr"], n["n_hgt"]))
    f.write(parents.tobytes())
    for name, dt, cnt in _FIELDS:
        a = np.ascontiguousarray(ev.get(name, np.zeros(0, dt)), "<" + dt)
        if len(a) != n[cnt]:
            raise ValueError(f"{name}: {len(a)} entries, expected {n[cnt]}")
        f.write(a.tobytes())


def read_generations(path: str):
    """Yield (generation, parents, events dict) for every record of the file."""
    with open(path, "rb") as f:
        while True:
            head = f.read(16)
            if not head:
                return
            if len(head) < 16 or head[:4] != MAGIC:
                raise ValueError("not a PSEV event dump")
            version, gen, n_rows = struct.unpack("<III", head[4:])
            if version != VERSION:
                raise ValueError(f"unsupported PSEV version {version}")
            counts = dict(zip(("n_core_mut", "n_acc_flip", "n_hr", "n_hgt"), struct.unpack("<QQQQ", f.read(32))))
            parents = np.frombuffer(f.read(4 * n_rows), "<u4").copy()
            ev = {}
            for name, dt, cnt in _FIELDS:
                size = np.dtype(dt).itemsize * counts[cnt]
                buf = f.read(size)
                if len(buf) != size:
                    raise ValueError("truncated PSEV record")
                ev[name] = np.frombuffer(buf, "<" + dt).copy()
            yield gen, parents, ev


def replay(sim, path: str) -> int:
    """Apply every generation of a dump to `sim` (a Pansim) with pansim_step_replay."""
    n = 0
    for _gen, parents, ev in read_generations(path):
        sim.step_replay(parents,
                        core_mut=(ev["core_mut_row"], ev["core_mut_site"], ev["core_mut_allele"]),
                        acc_flip=(ev["acc_flip_row"], ev["acc_flip_gene"]),
                        hr=(ev["hr_recipient"], ev["hr_locus"], ev["hr_value"]),
                        hgt=(ev["hgt_recipient"], ev["hgt_gene"]))
        n += 1
    return n
