"""Several GPUs in one process: `PansimGroup` over the pansim_group_* C ABI.

The reference is one process whose Population methods are called from the main thread
(main.rs:429-528); this mirror keeps that shape and lets the library shard the core
alignment by columns over the devices of the box (NCCL inside the library, see
include/pansim_b200.h "multi-GPU"). Method names follow `Pansim`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .params import Params
from .population import OK, PansimError, _ptr, make_config


class PansimGroup:
    def __init__(self, p: Params, n_devices: int, devices=None):
        self._lib = _ffi.lib()
        self.cfg = make_config(p)
        devs = None
        if devices is not None:
            devs = (C.c_int * n_devices)(*devices)
        h = C.c_void_p()
        rc = self._lib.pansim_group_create(C.byref(self.cfg), n_devices, devs, C.byref(h))
        if rc != OK:
            raise PansimError(rc, self._lib.pansim_group_last_error(None).decode())
        self._h = h
        self.N, self.G, self.L = self.cfg.pop_size, self.cfg.pan_size, self.cfg.core_size
        self.size = self._lib.pansim_group_size(h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pansim_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise PansimError(rc, self._lib.pansim_group_last_error(self._h).decode())

    def set_initial(self, core_row_onehot, acc_row):
        cr = np.ascontiguousarray(core_row_onehot, np.uint8)
        ar = np.ascontiguousarray(acc_row, np.uint8)
        assert cr.shape == (self.L,) and ar.shape == (self.G,)
        self._check(self._lib.pansim_group_set_initial(self._h, _ptr(cr), _ptr(ar)))

    def set_selection(self, s):
        s = np.ascontiguousarray(s, np.float64)
        self._check(self._lib.pansim_group_set_selection(self._h, _ptr(s)))

    def run_generations(self, gen0: int, n: int):
        self._check(self._lib.pansim_group_run_generations(self._h, gen0, n))

    def pair_counts(self, range1, range2):
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        P = len(r1)
        cd, it, un = (np.empty(P, np.uint32) for _ in range(3))
        self._check(self._lib.pansim_group_pair_counts(self._h, _ptr(r1), _ptr(r2), P, _ptr(cd), _ptr(it), _ptr(un)))
        return cd, it, un

    def run_generations_stats(self, gen0: int, n: int, range1, range2) -> np.ndarray:
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        out = np.empty((n, 4), np.float64)
        self._check(self._lib.pansim_group_run_generations_stats(self._h, gen0, n, _ptr(r1), _ptr(r2), len(r1), _ptr(out)))
        return out

    def all_pairs(self, chunk_pairs: int = 4_000_000):
        """Exact all-pairs mode: (core_diff, inter, union) of every pair i < j in (i, j) order."""
        parts = []

        def cb(_user, i0, i1, n, cd, it, un):
            parts.append((i0, i1, np.ctypeslib.as_array(cd, (n,)).copy(), np.ctypeslib.as_array(it, (n,)).copy(),
                          np.ctypeslib.as_array(un, (n,)).copy()))
            return 0

        self._check(self._lib.pansim_group_all_pairs(self._h, chunk_pairs, _ffi.PAIRS_CB(cb), None))
        return parts

    def walk_all_pairs(self, chunk_pairs: int = 4_000_000, sink=None) -> dict:
        """The same walk without keeping the vectors: `sink(i0, i1, core_diff, inter, union)` per block (views,
        valid during the call) or nothing. Returns wall ms, reduce-scatter device ms on shard 0, pairs."""
        def cb(_user, i0, i1, n, cd, it, un):
            if sink is not None:
                sink(i0, i1, np.ctypeslib.as_array(cd, (n,)), np.ctypeslib.as_array(it, (n,)), np.ctypeslib.as_array(un, (n,)))
            return 0

        self._check(self._lib.pansim_group_all_pairs(self._h, chunk_pairs, _ffi.PAIRS_CB(cb), None))
        wall, nccl, pairs = C.c_float(), C.c_float(), C.c_uint64()
        self._check(self._lib.pansim_group_all_pairs_timing(self._h, C.byref(wall), C.byref(nccl), C.byref(pairs)))
        return dict(wall_ms=wall.value, nccl_ms_shard0=nccl.value, pairs=pairs.value)

    def gene_counts(self) -> np.ndarray:
        out = np.empty(self.G, np.uint32)
        self._check(self._lib.pansim_group_gene_counts(self._h, _ptr(out)))
        return out

    def download_acc(self) -> np.ndarray:
        out = np.empty((self.N, self.G), np.uint8)
        self._check(self._lib.pansim_group_download_acc(self._h, _ptr(out)))
        return out

    def download_core(self) -> np.ndarray:
        out = np.empty((self.N, self.L), np.uint8)
        self._check(self._lib.pansim_group_download_core(self._h, _ptr(out)))
        return out

    def export_core_csv(self, row_begin: int, row_end: int) -> bytes:
        out = np.empty((row_end - row_begin) * 2 * self.L, np.uint8)
        self._check(self._lib.pansim_group_export_core_csv(self._h, row_begin, row_end, _ptr(out)))
        return out.tobytes()
