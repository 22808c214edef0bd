"""Batches of independent LPs across 1/2/4/8 GPUs (BASELINE.json config 4).

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).  LP ``i`` of the
batch goes to rank ``i mod world`` -- the partition SURVEY.md 8e prescribes -- and is solved there by
``vbk_solve_batch`` (csrc/vbk_batch.cu): one factor object + one CUDA stream per LP, ``nstreams`` LPs in
flight per GPU.  The data path has **no collective**; only the per-LP results (status, iteration
count, objectives: five numbers per LP) are gathered at the end so that every rank can return the whole
batch's verdicts.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class LpDesc(C.Structure):
    """include/vbkkt.h: vbk_lp_desc"""
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nz", C.c_int), ("iA", _ip), ("kA", _ip),
                ("A", _dp), ("b", _dp), ("c", _dp), ("f", C.c_double), ("x", _dp), ("y", _dp),
                ("status", C.c_int), ("iterations", C.c_int), ("primal_obj", C.c_double),
                ("dual_obj", C.c_double), ("seconds", C.c_double)]


def declare(lib):
    lib.vbk_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(LpDesc), C.c_int]
    lib.vbk_solve_batch.restype = C.c_int
    lib.vbk_batch_launches.argtypes = []
    lib.vbk_batch_launches.restype = C.c_longlong
    return lib


def shard(nlp: int, rank: int, world: int):
    """Indices of the LPs rank ``rank`` owns: i mod world == rank."""
    return list(range(rank, nlp, world))


def solve_local(lib, lps, method="hsd", device=0, mode=0, nstreams=4):
    """Solve ``lps`` (objects with m, n, nz, kA, iA, A, b, c, f) on one GPU.  ``mode`` 0 = strict (the reference's
    arithmetic, bit for bit -- the mode that carries the parity claim, and the default), 1 = fast (opt in: dense-window
    factorisation, tolerance parity only).  Returns a list of dicts
    (status, iterations, primal_obj, dual_obj, seconds, x, y) in input order."""
    declare(lib)
    keep, descs = [], (LpDesc * max(len(lps), 1))()
    for d, lp in zip(descs, lps):
        arrs = dict(iA=np.ascontiguousarray(lp.iA, dtype=np.int32), kA=np.ascontiguousarray(lp.kA, dtype=np.int32),
                    A=np.ascontiguousarray(lp.A, dtype=np.float64), b=np.ascontiguousarray(lp.b, dtype=np.float64),
                    c=np.ascontiguousarray(lp.c, dtype=np.float64),
                    x=np.zeros(lp.n, dtype=np.float64), y=np.zeros(lp.m, dtype=np.float64))
        keep.append(arrs)
        d.m, d.n, d.nz, d.f = lp.m, lp.n, lp.nz, float(lp.f)
        d.iA, d.kA = arrs["iA"].ctypes.data_as(_ip), arrs["kA"].ctypes.data_as(_ip)
        for k in ("A", "b", "c", "x", "y"):
            setattr(d, k, arrs[k].ctypes.data_as(_dp))
    if lps:
        lib.vbk_solve_batch({"hsd": 0, "intpt": 1, "hsdls": 2}[method], device, mode, len(lps), descs, nstreams)
    return [dict(status=int(d.status), iterations=int(d.iterations), primal_obj=float(d.primal_obj),
                 dual_obj=float(d.dual_obj), seconds=float(d.seconds), x=a["x"], y=a["y"])
            for d, a in zip(descs, keep)]


def solve_batch(lib, make_lp, nlp, method="hsd", device=0, mode=0, nstreams=4, group=None, result_device="cpu"):
    """Distributed batch solve.  ``make_lp(i)`` builds LP ``i`` (only called for the LPs this rank owns,
    so generators run sharded too).  Returns (summary[nlp, 5] = status, iterations, primal_obj, dual_obj,
    seconds for every LP of the batch, local_results)."""
    import torch
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    mine = shard(nlp, rank, world)
    local = solve_local(lib, [make_lp(i) for i in mine], method, device, mode, nstreams)
    per_rank = (nlp + world - 1) // world
    buf = torch.full((per_rank, 5), float("nan"), dtype=torch.float64, device=result_device)
    for k, r in enumerate(local):
        buf[k] = torch.tensor([r["status"], r["iterations"], r["primal_obj"], r["dual_obj"], r["seconds"]],
                              dtype=torch.float64)
    if on and world > 1:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)          # results only: 40 bytes per LP
    else:
        parts = [buf]
    summary = np.full((nlp, 5), np.nan)
    for r, part in enumerate(parts):
        idx = shard(nlp, r, world)
        summary[idx] = part[: len(idx)].cpu().numpy()
    return summary, local
