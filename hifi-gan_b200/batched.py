"""Job tables for hg_prep_batched (include/hifigan_b200.h, "Batched weight preparation"): the per-layer weight
folds / packs / weight-gradient finishes of a whole network described once, resident on the device, and run as one
launch per phase instead of one launch per layer and call."""
from __future__ import annotations

import ctypes
from ctypes import c_int32, c_void_p
from typing import Dict, List, Sequence, Tuple

import torch

from . import _lib

GEN_CONV, GEN_CONVTR, GEN_POST, DISC_ROW, TRANSPOSE_TILE, DISC_DGRAD_TILE, FINISH_ROW, LOSS_SUM = range(8)


class PrepJob(ctypes.Structure):
    _fields_ = [("src0", c_void_p), ("src1", c_void_p), ("src2", c_void_p), ("dst0", c_void_p), ("dst1", c_void_p),
                ("kind", c_int32), ("i", c_int32 * 16), ("tab", c_int32 * 64), ("pad_", c_int32)]


def _ptr(x) -> int:
    if x is None:
        return 0
    return x if isinstance(x, int) else x.data_ptr()


class JobTable:
    """Collect jobs per phase with add(), upload once with finalize(), run a phase with launch(phase).
    The table holds raw device pointers: every tensor passed to add() must outlive the table, unmoved."""

    def __init__(self, device):
        self.device = device
        self._jobs: List[PrepJob] = []
        self._blocks: Dict[str, List[Tuple[int, int]]] = {}
        self._smem: Dict[str, int] = {}
        self.ranges: Dict[str, Tuple[int, int, int]] = {}
        self._dev = None

    def add(self, phase: str, kind: int, nblocks: int, smem: int = 0, src0=None, src1=None, src2=None, dst0=None,
            dst1=None, ints: Sequence[int] = (), tab: Sequence[int] = ()) -> None:
        if self._dev is not None:
            raise RuntimeError("JobTable: add() after finalize()")
        if len(ints) > 16 or len(tab) > 64 or nblocks <= 0:
            raise ValueError("JobTable.add: too many integer fields / empty job")
        j = PrepJob()
        j.src0, j.src1, j.src2, j.dst0, j.dst1 = (_ptr(src0) or None, _ptr(src1) or None, _ptr(src2) or None,
                                                   _ptr(dst0) or None, _ptr(dst1) or None)
        j.kind = kind
        for n, v in enumerate(ints):
            j.i[n] = int(v)
        for n, v in enumerate(tab):
            j.tab[n] = int(v)
        idx = len(self._jobs)
        self._jobs.append(j)
        self._blocks.setdefault(phase, []).extend((idx, b) for b in range(nblocks))
        self._smem[phase] = max(self._smem.get(phase, 0), int(smem))

    def add_loss_sum(self, phase: str, mode: int, a, b, n: int, c: float, out_ptr: int, chunk: int = 16384) -> None:
        """sum |a - b| (mode 0, fp32; mode 2, bf16: n counts 16-byte groups) or sum (c - a)^2 (mode 1) -> *out_ptr +="""
        import struct
        cbits = struct.unpack("<i", struct.pack("<f", float(c)))[0]
        lo = n & 0xFFFFFFFF
        lo = lo - (1 << 32) if lo >= (1 << 31) else lo          # the low word travels as a signed int32
        self.add(phase, LOSS_SUM, (n + chunk - 1) // chunk, 0, a, b, dst0=out_ptr, ints=(mode, lo, n >> 32, chunk, cbits))

    def finalize(self) -> "JobTable":
        if ctypes.sizeof(PrepJob) != _lib.lib().hg_prep_job_size():
            raise RuntimeError("hg_prep_job layout mismatch between batched.py and the library")
        raw = b"".join(bytes(j) for j in self._jobs)
        jobs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        pairs, first = [], 0
        for phase, blocks in self._blocks.items():
            self.ranges[phase] = (first, len(blocks), self._smem[phase])
            pairs.extend(blocks)
            first += len(blocks)
        bj = torch.tensor(pairs, dtype=torch.int32).reshape(-1, 2).to(self.device)
        self._dev = (jobs, bj)
        return self

    def has(self, phase: str) -> bool:
        return phase in self.ranges

    def launch(self, phase: str) -> None:
        first, n, smem = self.ranges[phase]
        jobs, bj = self._dev
        _lib.check(_lib.lib().hg_prep_batched(jobs.data_ptr(), bj.data_ptr(), first, n, smem,
                                              torch.cuda.current_stream().cuda_stream), "hg_prep_batched")
