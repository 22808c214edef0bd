#!/usr/bin/env python
"""bench.py -- the KKT step (LDL^T factor + solves) of Vanderbei's hsd interior-point method on B200.

One "step" = the KKT work of one hsd iteration (reference src/ipo/hsd.c:215-228) on a fixed
mid-solve iterate of a netlib LP: 1 x ldltfac (assemble + numeric LDL^T) and 2 x forwardbackward
(forward/diagonal/backward sweeps with iterative refinement, 2 SpMVs and 2 max-norms per pass).
The iterate (E, D, right-hand sides) is a committed fixture produced by the pinned oracle
(tests/golden/iterates/, see make_iterates.py), so the GPU arm, the CPU baseline and the reference
arm all work on byte-identical inputs, and the GPU result is checked against the reference's
solution inside the bench.

Metric (BASELINE.json): LDL^T factor+solve GFLOP/s, with the reference's own work model
(SURVEY.md 8d): F_fac = narth = sum_j c_j^2 + 3 Lnz + N (ldlt.c:1243-1248), F_rawsolve = 4 Lnz + N,
F_smx = 2 nz.  `value`: inputs resident in HBM, device-pointer calls, CUDA events on the library's
stream.  `e2e`: the same step through the host-buffer C ABI (what the reference's hsd.c would call:
ldltfac / forwardbackward), pinned host arrays, H2D+D2H inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload pilot87] [--mode strict|fast]

Two more workloads cover the multi-GPU rows of SURVEY.md 8e (run them under torchrun for N > 1):
    --workload batch      BASELINE config 4: batches of independent random sparse LPs (m=2000, n=4000),
                          LP i on rank i mod N, no collective; metric LPs/s
    --workload rowblock   BASELINE config 5: row-block smx / transpose-smx + dot / max-norm all-reduce on the
                          synthetic multicommodity LP (default R=100, K=126: m=2.56e6, n=4.99e6); metric GB/s
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tests"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# BASELINE config 3: multicommodity flow on a planar R x R grid, K commodities.  "mcf" = the point the
# survey calls "the one to run to completion on CPU" (R=32, K=25: N=154 368, Lnz=3.07e7, 8.0e10 flops per factorisation);
# "mcf:R:K" selects another point.  Symbolic counts of the reference's ordering for the sizes used in profiles/:
MCF_KNOWN = {(32, 25): (8.019e10, 30717036), (26, 16): (1.757e10, 10465056), (20, 12): (3.246e9, 3362836)}


def mcf_workload(name):
    """Synthetic multicommodity LP + a synthetic mid-solve iterate (E, D log-uniform over six decades like the x/z,
    y/w ratios of an interior-point iterate; right-hand sides standard normal).  No reference solution is stored:
    the bench checks the KKT residual of the GPU solution instead (size-independent property)."""
    # orderings of these LPs computed once on the host (tests/golden/make_symcache.py, 8 bytes per row/column): the
    # library reads them instead of repeating minutes of explicit-fill minimum degree
    os.environ.setdefault("VBK_SYM_CACHE", str(ROOT / "tests" / "golden" / "symcache"))
    vbw = _load("vbkkt_workloads", ROOT / "linear-programming-vanderbei_b200" / "workloads.py")
    parts = name.split(":")
    R, K = (int(parts[1]), int(parts[2])) if len(parts) == 3 else (32, 25)
    lp = vbw.multicommodity_lp(R, K)
    rng = np.random.default_rng(20)
    it = {"E": 10.0 ** rng.uniform(-3, 3, lp.m), "D": 10.0 ** rng.uniform(-3, 3, lp.n),
          "rhs_y": rng.standard_normal(lp.m), "rhs_x": rng.standard_normal(lp.n)}
    lp.extra = {}
    if (R, K) in MCF_KNOWN:
        lp.extra = {"sym_narth": MCF_KNOWN[(R, K)][0], "sym_lnz": MCF_KNOWN[(R, K)][1]}
    return lp, it


def kkt_residual(lp, it, sy, sx):
    """max-norm residual of [-E A; A^T D] [sy; sx] = [rhs_y; rhs_x] relative to the right-hand side."""
    import scipy.sparse as sp
    A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
    r1 = -it["E"] * sy + A @ sx - it["rhs_y"]
    r2 = A.T @ sy + it["D"] * sx - it["rhs_x"]
    scale = max(np.abs(it["rhs_y"]).max(), np.abs(it["rhs_x"]).max(), np.abs(sy).max(), np.abs(sx).max()) + 1.0
    return float(max(np.abs(r1).max(), np.abs(r2).max()) / scale)


def load_workload(name, it):
    import harness as H
    if name.startswith("mcf"):
        return mcf_workload(name)
    lp = H.load_fixture(name)
    z = np.load(H.GOLDEN / "iterates" / f"{name}_it{it}.npz")
    return lp, {k: z[k] for k in z.files}


def work_model(sym_narth, lnz, N, nz, rawsolves, factors=1):
    """flops of `factors` factorisations + `rawsolves` refinement passes (each 1 rawsolve + 2 smx)."""
    f_fac = sym_narth
    f_sol = 4.0 * lnz + N
    f_smx = 2.0 * nz
    return factors * f_fac + rawsolves * (f_sol + 2 * f_smx)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference / CPU arm: the compiled reference (oracle/_ref) when present, else the oracle port
# ----------------------------------------------------------------------------------------------
class CpuKkt:
    def __init__(self, lp):
        import harness as H
        self.H, self.lp = H, lp
        self.kAt, self.iAt, self.At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
        ref = H.load_ref("hsd")
        if ref is not None:
            self.kind, self.lib = "reference", ref
            ref.ldltfac.argtypes = [C.c_int, C.c_int, H.c_int_p, H.c_int_p, H.c_double_p, H.c_double_p,
                                    H.c_double_p, H.c_int_p, H.c_int_p, H.c_double_p, C.c_int]
            ref.forwardbackward.argtypes = [H.c_double_p] * 4
        else:
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "restatement"], check=True, stdout=subprocess.DEVNULL)
            self.kind = "port"
            self.lib = H.declare_oracle(C.CDLL(str(ROOT / "oracle" / "libkkt_oracle.so")))
            self.F = H.oracle_factor_for(self.lib, lp)

    def step(self, it):
        """One KKT step exactly as hsd.c:218-228 issues it; returns the first solution."""
        H, lp = self.H, self.lp
        E, D = it["E"], it["D"]
        fy, fx = it["rhs_y"].copy(), it["rhs_x"].copy()
        gy, gx = -lp.b, -lp.c
        if self.kind == "reference":
            self.lib.ldltfac(lp.n, lp.m, H.ptr_i(self.kAt), H.ptr_i(self.iAt), H.ptr_d(self.At), H.ptr_d(E),
                             H.ptr_d(D), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(lp.A), 1)
            self.lib.forwardbackward(H.ptr_d(E), H.ptr_d(D), H.ptr_d(fy), H.ptr_d(fx))
            self.lib.forwardbackward(H.ptr_d(E), H.ptr_d(D), H.ptr_d(gy), H.ptr_d(gx))
        else:
            self.F.factor(E, D)
            fy, fx, _ = self.F.solve(E, D, fy, fx)
            self.F.solve(E, D, gy, gx)
        return fy, fx


def cpu_measure(lp, it, flops_per_step, budget_s, max_steps):
    cpu = CpuKkt(lp)
    t0 = time.perf_counter()
    cpu.step(it)                                   # includes the one-time symbolic phase; not timed below
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu.step(it)
    one = time.perf_counter() - t0
    if one > budget_s:                             # a single step already exceeds the budget: that step is the sample
        steps, dt = 1, one
    else:
        steps = int(max(1, min(max_steps, budget_s / max(one, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu.step(it)
        dt = (time.perf_counter() - t0) / steps
    return {"value": flops_per_step / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": cpu.kind,
            "sample": f"{steps} KKT steps (1 ldltfac + 2 forwardbackward) of the same workload, "
                      f"{dt * 1e3:.2f} ms/step on 1 host core; symbolic phase ({first:.2f} s incl. first step) excluded",
            "ms_per_step": dt * 1e3}


def _clock_fields(sampler):
    return sampler.summary()


def multi_gpu_workload(a, rank, local_rank, world):
    """BASELINE configs 4 (batch of independent LPs) and 5 (row-block smx + all-reduce).  One process per
    GPU; barrier + synchronize on both sides of the timed region; max over ranks."""
    import torch
    import torch.distributed as dist
    vb = _load("vbkkt", ROOT / "linear-programming-vanderbei_b200" / "__init__.py")
    import harness as H
    mode = vb.MODE_STRICT if a.mode == "strict" else vb.MODE_FAST

    if a.impl == "reference":
        # the reference's own CPU path for the same unit of work, rank 0 only
        if rank != 0:
            return
        if a.workload == "batch":
            ref = H.load_ref("hsd")
            fn, kind = (ref.solver, "reference") if ref is not None else (None, "port")
            if fn is None:
                subprocess.run(["make", "-C", str(ROOT / "oracle"), "restatement"], check=True, stdout=subprocess.DEVNULL)
                fn = H.declare_oracle(C.CDLL(str(ROOT / "oracle" / "libkkt_oracle.so"))).kko_solver_hsd
            lp = vb.workloads.random_sparse_lp(0, a.batch_m, a.batch_n)
            t0 = time.perf_counter()
            st, log, x, y = H.call_solver(fn, lp)
            dt = time.perf_counter() - t0
            v = 1.0 / dt
            print(json.dumps({"impl": "reference", "metric": "independent LPs solved per second (hsd)", "value": v,
                              "unit": "LP/s", "n_gpus": a.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic random sparse LP, seed 0",
                              "config": {"workload": f"batch of random sparse LPs m={a.batch_m} n={a.batch_n} (8 nnz/col)"},
                              "cpu_baseline": {"value": v, "unit": "LP/s", "cores": 1, "kind": kind,
                                               "sample": f"1 LP, {len(H.iteration_lines(log))} iterations, status {st}, {dt:.1f} s on 1 host core"},
                              "e2e": {"value": v, "unit": "LP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        else:
            lp = vb.workloads.multicommodity_lp(a.grid, a.commodities)
            kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
            ref = H.load_ref("hsd")
            rng = np.random.default_rng(1)
            x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
            rho, sig = np.zeros(lp.m), np.zeros(lp.n)
            if ref is not None:
                smx, dot, mxv, kind = ref.smx, ref.dotprod, ref.maxv, "reference"
            else:
                subprocess.run(["make", "-C", str(ROOT / "oracle"), "restatement"], check=True, stdout=subprocess.DEVNULL)
                o = H.declare_oracle(C.CDLL(str(ROOT / "oracle" / "libkkt_oracle.so")))
                smx, dot, mxv, kind = o.kko_smx, o.kko_dotprod, o.kko_maxv, "port"
            smx.argtypes = [C.c_int, C.c_int, H.c_double_p, H.c_int_p, H.c_int_p, H.c_double_p, H.c_double_p]
            dot.argtypes, dot.restype = [H.c_double_p, H.c_double_p, C.c_int], C.c_double
            mxv.argtypes, mxv.restype = [H.c_double_p, C.c_int], C.c_double
            def step():
                smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(rho))
                smx(lp.n, lp.m, H.ptr_d(At), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(y), H.ptr_d(sig))
                dot(H.ptr_d(x), H.ptr_d(sig), lp.n); dot(H.ptr_d(y), H.ptr_d(rho), lp.m)
                dot(H.ptr_d(rho), H.ptr_d(rho), lp.m); dot(H.ptr_d(sig), H.ptr_d(sig), lp.n)
                mxv(H.ptr_d(rho), lp.m); mxv(H.ptr_d(sig), lp.n)
            bytes_step = rowblock_bytes(lp.m, lp.n, lp.nz)
            for _ in range(a.warmup):
                step()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                step()
            dt = (time.perf_counter() - t0) / a.steps
            v = bytes_step / dt / 1e9
            print(json.dumps({"impl": "reference", "metric": "row-block smx + dot/maxv GB/s", "value": v, "unit": "GB/s",
                              "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic multicommodity LP",
                              "config": {"workload": f"multicommodity R={a.grid} K={a.commodities}: m={lp.m} n={lp.n} nz={lp.nz}"},
                              "cpu_baseline": {"value": v, "unit": "GB/s", "cores": 1, "kind": kind,
                                               "sample": f"{a.steps} steps (2 smx + 4 dotprod + 2 maxv), {dt * 1e3:.1f} ms/step, 1 host core"},
                              "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    lib = vb.load(os.environ.get("VBK_LIB"))
    if lib.vbk_device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run_batch(a, vb, lib, rank, local_rank, world) if a.workload == "batch" else run_rowblock(a, vb, lib, rank, local_rank, world)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _sync_helpers(world, dev):
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return barrier, max_over_ranks


def run_batch(a, vb, lib, rank, local_rank, world):
    """BASELINE config 4: independent random sparse LPs, LP i on rank i mod world, no collective on the data path.
    Returns the result dict on rank 0 (None elsewhere).  The process group, if any, is the caller's."""
    import torch
    dev = torch.device("cuda", local_rank)
    barrier, max_over_ranks = _sync_helpers(world, dev)
    mode = vb.MODE_STRICT if a.batch_mode == "strict" else vb.MODE_FAST
    warm = max(a.warmup, 1)
    sampler = ClockSampler(local_rank)
    B = a.batch_per_gpu
    nlp = B * world
    gen = lambda i: vb.workloads.random_sparse_lp(i, a.batch_m, a.batch_n)
    mine = [gen(i) for i in vb.batch.shard(nlp, rank, world)]
    vb.batch.solve_local(lib, mine[:1], "hsd", local_rank, mode, 1)              # warm-up (module load, allocator)
    for _ in range(warm - 1):
        vb.batch.solve_local(lib, mine[:min(len(mine), a.streams)], "hsd", local_rank, mode, a.streams)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    res = None
    batch_launches = 0
    for _ in range(a.batch_steps):
        res = vb.batch.solve_local(lib, mine, "hsd", local_rank, mode, a.streams)
        batch_launches += int(lib.vbk_batch_launches())
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.summary()
    ok = all(r["status"] == 0 for r in res)
    its = float(np.mean([r["iterations"] for r in res]))
    gap = float(max(abs(r["primal_obj"] - r["dual_obj"]) / max(1.0, abs(r["primal_obj"])) for r in res))
    # strict-mode sample on rank 0: the same LP in the parity mode, for the iteration count and objective it should have
    strict_sample = None
    if rank == 0 and mode != vb.MODE_STRICT and not a.no_strict:
        r0 = vb.batch.solve_local(lib, mine[:1], "hsd", local_rank, vb.MODE_STRICT, 1)[0]
        strict_sample = {"lp": 0, "status": r0["status"], "iterations": r0["iterations"], "seconds": r0["seconds"],
                         "fast_iterations": res[0]["iterations"],
                         "rel_objective_diff_fast_vs_strict": abs(res[0]["primal_obj"] - r0["primal_obj"]) / max(1.0, abs(r0["primal_obj"]))}
    if rank != 0:
        return None
    v = nlp * a.batch_steps / dt
    nbytes_in = sum(12 * lp.nz + 4 * (lp.n + 1) + 8 * (lp.m + lp.n) for lp in mine)
    return {"metric": "independent LPs solved per second (hsd)", "value": v, "unit": "LP/s", "n_gpus": world,
            "steps": a.batch_steps, "warmup": warm, "ms_per_step": dt / a.batch_steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic random sparse LPs (seed = LP index)",
            "config": {"workload": f"BASELINE config 4: batch of {nlp} random sparse LPs m={a.batch_m} n={a.batch_n} (8 nnz/col), "
                                   f"{B} per GPU, LP i on rank i mod {world}", "mode": a.batch_mode,
                       "streams_per_gpu": a.streams, "host_cores": os.cpu_count(),
                       "l2": "working sets rotate across LPs (each LP has its own factor object)",
                       "parallelism": f"{world} rank(s), no collective on the data path"},
            "e2e": {"value": v, "unit": "LP/s", "h2d_bytes_per_step": int(nbytes_in),
                    "d2h_bytes_per_step": int(sum(8 * (lp.m + lp.n) for lp in mine)),
                    "note": "vbk_solve_batch takes host arrays and returns host x,y: the timed region is end to end, symbolic phase included"},
            "gpu_launches": batch_launches, "gpu_launches_note": "rank 0's kernels inside the timed region (vbk_batch_launches)",
            "clocks": clocks,
            "parity": {"all_optimal": bool(ok), "mean_iterations": its, "max_rel_duality_gap": gap, "strict_sample": strict_sample}}


def run_rowblock(a, vb, lib, rank, local_rank, world, partition=None):
    """BASELINE config 5: row-block smx / transpose-smx + dot / max-norm all-reduce on the synthetic multicommodity LP.
    Returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import harness as H
    dev = torch.device("cuda", local_rank)
    barrier, max_over_ranks = _sync_helpers(world, dev)
    warm = max(a.warmup, 3)
    sampler = ClockSampler(local_rank)
    steps = a.rowblock_steps
    if True:
        lp = vb.workloads.multicommodity_lp(a.grid, a.commodities)
        kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
        partition = partition or a.rowblock_partition
        coupled = partition == "coupled"
        ops = (vb.rowblock.CoupledOps if coupled else vb.rowblock.RowBlockOps)(lib, lp.m, lp.n, lp.kA, lp.iA, lp.A, kAt, iAt, At, dev)
        rng = np.random.default_rng(1)
        x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
        lx, ly = ops.local_x(x), ops.local_y(y)
        rho = torch.zeros(ops.rows_per, dtype=torch.float64, device=dev)
        sig = torch.zeros(ops.cols_per, dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step_eager():
            return ops.step(lx, ly, rho, sig)      # hsd.c:182-195: A x, A^T y, 4 dot products, 2 max-norms
        for _ in range(warm):
            d, mx = step_eager()
        # The step is a dozen small launches and three collectives: issued one by one from Python it is bound by launch
        # latency, not by the GPU.  Capture it once in a CUDA graph (NCCL collectives included) and replay that.
        graph, graphed = None, "eager"
        if not a.no_graph:
            try:
                torch.cuda.synchronize()
                if world > 1:
                    import torch.distributed as dist
                    dist.barrier()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step_eager()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    gd, gmx = ops.step(lx, ly, rho, sig, overlap=False)
                graph.replay()
                torch.cuda.synchronize()
                graphed = "cuda graph (one replay per step)"
            except Exception as e:                       # capture not possible with this torch/NCCL: stay eager, say so
                graph, graphed = None, f"eager (graph capture failed: {type(e).__name__})"
                torch.cuda.synchronize()

        def step():
            if graph is not None:
                graph.replay()
                return gd, gmx
            return step_eager()
        d, mx = step()
        # parity: the distributed products against a host computation of the same sums
        import scipy.sparse as sp
        Acsr = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n)).tocsr()
        if coupled:
            full_ref = Acsr @ x
            got = rho.cpu().numpy()
            loc_ok = bool(np.array_equal(got[: ops.nl], full_ref[ops.local_rows]))          # local rows: the reference's bits
            err = float(np.max(np.abs(got[ops.nl:] - full_ref[ops.shared_rows])) / max(1.0, np.max(np.abs(full_ref)))) if ops.ns else 0.0
            assert loc_ok, "local rows of the coupled partition must equal the host row sums bit for bit"
            nrow_local = ops.rows_per
        else:
            rho_ref = Acsr[ops.r0:ops.r1] @ x
            err = float(np.max(np.abs(rho.cpu().numpy()[: ops.r1 - ops.r0] - rho_ref)) / max(1.0, np.max(np.abs(rho_ref))))
            nrow_local = ops.r1 - ops.r0
        dref = float(y @ (Acsr @ x))
        derr = abs(float(d[1]) - dref) / max(1.0, abs(dref))
        assert err < 1e-12 and derr < 1e-10, (err, derr)
        sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s in range(steps):
            flush.fill_(s & 0xFF)
            barrier()
            ev[s][0].record()
            step()
            ev[s][1].record()
        barrier()
        ms = max_over_ranks(float(sum(e0.elapsed_time(e1) for e0, e1 in ev)) / steps)
        clocks = sampler.summary()
        # e2e: host vectors in, host results out, per step
        hx = torch.from_numpy(x[ops.c0:ops.c1].copy()).pin_memory()
        hy = (ly.cpu() if coupled else torch.from_numpy(y[ops.r0:ops.r1].copy())).pin_memory()
        hrho, hsig = torch.empty(nrow_local, dtype=torch.float64).pin_memory(), torch.empty(ops.c1 - ops.c0, dtype=torch.float64).pin_memory()
        def step_host():
            lx[: ops.c1 - ops.c0].copy_(hx, non_blocking=True); ly[:nrow_local].copy_(hy, non_blocking=True)
            d, mx = step()
            hrho.copy_(rho[:nrow_local], non_blocking=True); hsig.copy_(sig[: ops.c1 - ops.c0], non_blocking=True)
            return d.cpu(), mx.cpu()
        step_host(); barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        if rank == 0:
            total_bytes = rowblock_bytes(lp.m, lp.n, lp.nz)
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
            hbm = peaks.get("hbm_gbs", 6650.0)
            v = total_bytes / (ms * 1e-3) / 1e9
            out = {"metric": "row-block smx + dot/maxv GB/s", "value": v, "unit": "GB/s", "n_gpus": world, "steps": steps,
                   "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                   "dtype": "f64", "data": "synthetic multicommodity LP (generator seed 1)",
                   "config": {"workload": f"BASELINE config 5: multicommodity R={a.grid} K={a.commodities}: m={lp.m} n={lp.n} nz={lp.nz}; step = A x + A^T y "
                                          + ("(column blocks; rows local to a rank or shared: one all-reduce of the shared rows' partial sums, no vector travels) "
                                             "+ 4 dot products + 2 max-norms (ONE 72-byte all-gather)" if coupled else
                                             "(all-gather + row-block SpMV each, both all-gathers in flight together) + 4 dot products + 2 max-norms (ONE 48-byte all-gather)"),
                              "partition": partition, "l2": "flushed between timed steps (256 MiB write)", "launch": graphed,
                              "parallelism": (f"column blocks over {world} rank(s); {ops.ns} shared rows of {lp.m}: NCCL all-reduce of {8 * ops.ns} bytes per step" if coupled else
                                              f"row blocks over {world} rank(s); NCCL all-gather of x, y and of the 6 partial scalars")},
                   "e2e": {"value": total_bytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_ms,
                           "h2d_bytes_per_step": int(8 * (ops.c1 - ops.c0 + nrow_local)),
                           "d2h_bytes_per_step": int(8 * (ops.c1 - ops.c0 + nrow_local) + 48)},
                   "gpu_launches": 5 * steps, "clocks": clocks,
                   "roofline": {"kernel": "k_spmv_rows + k_dot_partial/k_absmax_partial (whole step)", "bound": "hbm",
                                "achieved": v / world, "peak": hbm, "unit": "GB/s", "frac": v / world / hbm, "traffic": None,
                                "note": "per-GPU algorithmic bytes (12 B per nonzero, 8 B per vector entry read or written) over the step time, collectives included"},
                   "parity": ({"local_rows_bit_identical": True, "max_rel_err_Ax_shared_rows": err, "rel_err_dot": derr,
                               "note": "A^T y and the local rows of A x keep the reference's summation order; the shared rows are re-associated by column block"}
                              if coupled else {"max_rel_err_Ax": err, "rel_err_dot": derr})}
            return out
    return None


def rowblock_bytes(m, n, nz):
    """Algorithmic bytes of one step: two smx (12 B per nonzero + pointers + input + output vector),
    four dot products (16 B per entry), two max-norms (8 B per entry)."""
    smx1 = 12 * nz + 4 * (m + 1) + 8 * n + 8 * m
    smx2 = 12 * nz + 4 * (n + 1) + 8 * m + 8 * n
    dots = 16 * (n + m + m + n)
    mx = 8 * (m + n)
    return smx1 + smx2 + dots + mx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # dfl001 is the largest netlib LP BASELINE.json's config 2 names (FP64-bound factor, ~90 flop/B)
    ap.add_argument("--workload", default=os.environ.get("VBK_BENCH_WORKLOAD", "dfl001"))
    ap.add_argument("--iterate", type=int, default=20)
    # strict (default, the headline) = bit-exact replay of the reference's rounding order: the only mode that meets the
    # north_star tolerances on the whole netlib suite; fast = dense-window factorisation with re-associated sums, parity
    # by tolerance on the robust problems only (reported beside it at N=1 as `fast_mode`, with that caveat)
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--no-strict", action="store_true", help="skip the other mode's side measurement")
    ap.add_argument("--no-solve-time", action="store_true", help="skip the full-solve timings (solve_time) at N=1")
    ap.add_argument("--no-sharded", action="store_true", help="N>1: skip the batch / row-block workloads attached to the line")
    ap.add_argument("--batch-mode", default="fast", choices=["strict", "fast"],
                    help="batch workload: arithmetic mode (fast: tolerance parity, all LPs checked optimal + one strict sample)")
    ap.add_argument("--batch-steps", type=int, default=1)
    ap.add_argument("--rowblock-steps", type=int, default=20)
    ap.add_argument("--rowblock-partition", default="rows", choices=["rows", "coupled"],
                    help="rowblock workload: 'rows' = equal row blocks + all-gather of x and y (bit-identical smx); 'coupled' = column blocks, "
                         "rows local to a rank or shared, one all-reduce of the shared rows' partial sums")
    ap.add_argument("--no-graph", action="store_true", help="row-block workload: launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="batch workload: LPs per GPU per step (default: 2 per solver stream, at least 8)")
    ap.add_argument("--batch-m", type=int, default=2000)
    ap.add_argument("--batch-n", type=int, default=4000)
    # one host thread per stream runs the LP's symbolic phase (0.4 s per m=2000 LP on one core + handle set-up, measured) and its METHOD
    # loop; the GPU part of such an LP is 0.2-0.4 s, so the batch is bound by host cores (measured: 4 streams 3.5-7.9 LP/s depending on the box, 8 streams 7.7, 16 streams 6.1): default = cores / ranks, at most 8
    ap.add_argument("--streams", type=int, default=0, help="batch workload: solver streams in flight per GPU (default: host cores / ranks, 2..8)")
    ap.add_argument("--grid", type=int, default=100, help="rowblock workload: grid side R")
    ap.add_argument("--commodities", type=int, default=126, help="rowblock workload: K")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.streams <= 0:
        a.streams = max(2, min(16, (os.cpu_count() or 8) // max(world, 1)))
    if a.batch_per_gpu <= 0:
        a.batch_per_gpu = max(8, 4 * a.streams)
    if a.workload in ("batch", "rowblock"):
        if a.workload == "batch":
            a.batch_steps = max(a.batch_steps, 1)
        else:
            a.rowblock_steps = a.steps
        return multi_gpu_workload(a, rank, local_rank, world)

    lp, it = load_workload(a.workload, a.iterate)
    N, nz = lp.m + lp.n, lp.nz
    what = (f"synthetic multicommodity-flow LP {lp.name}" if a.workload.startswith("mcf") else f"netlib {a.workload}")
    config = {"workload": f"{what} (solver-space m={lp.m} n={lp.n} nz={nz}, N={N}), hsd iterate "
                          f"{'synthetic' if a.workload.startswith('mcf') else a.iterate}: 1 ldltfac + 2 forwardbackward per step", "mode": a.mode,
              "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"{world} independent LP replica(s)"}
    metric = "LDL^T factor+solve GFLOP/s (hsd KKT step)"

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        if a.gpus > 1 and not os.environ.get("VBK_BENCH_REF_CHILD"):
            # the config is N independent replicas: the (single-threaded) reference runs N of them side by side, one host
            # core each, as separate processes (it keeps ONE factor object per process, ldlt.c:108-120)
            env = dict(os.environ, VBK_BENCH_REF_CHILD="1")
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "LOCAL_WORLD_SIZE", "TORCHELASTIC_RUN_ID"):
                env.pop(k, None)
            cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", str(a.steps), "--warmup", str(a.warmup),
                   "--workload", a.workload, "--iterate", str(a.iterate)]
            kids = [subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(a.gpus)]
            lines = []
            for kproc in kids:
                out_k, _ = kproc.communicate()
                rows = [ln for ln in out_k.splitlines() if ln.startswith("{")]
                if kproc.returncode != 0 or not rows:
                    raise SystemExit("bench.py --impl reference: a replica process failed")
                lines.append(json.loads(rows[-1]))
            dt = max(r["ms_per_step"] for r in lines) * 1e-3
            flops = lines[0]["flops_per_step"]
            v = a.gpus * flops / dt / 1e9
            one = lines[0]
            one.update({"value": v, "n_gpus": a.gpus, "ms_per_step": dt * 1e3, "config": config,
                        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": a.gpus, "kind": one["cpu_baseline"]["kind"],
                                         "sample": f"{a.gpus} replicas side by side (one process and one host core each), {one['steps']} KKT steps each, "
                                                   f"slowest replica {dt * 1e3:.2f} ms/step"},
                        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            print(json.dumps(one))
            return
        cpu = CpuKkt(lp)                                         # nothing of the product on this arm
        t0 = time.perf_counter()
        sol_y, sol_x = cpu.step(it)
        first = time.perf_counter() - t0
        if "sol_y" in it:
            assert np.array_equal(sol_y, it["sol_y"]), "reference arm does not reproduce the fixture"
        else:
            assert kkt_residual(lp, it, sol_y, sol_x) < 1e-6
        # flop model from the reference's own symbolic counts stored in the fixture (ldlt.c:1243-1248);
        # 2 rawsolve passes per step is what both arms need on these iterates (checked by the GPU arm)
        if "sym_narth" not in lp.extra:
            raise SystemExit("bench.py --impl reference: no stored symbolic counts for this synthetic size (see MCF_KNOWN)")
        flops = work_model(float(lp.extra["sym_narth"]), int(lp.extra["sym_lnz"]), N, nz, 2)
        if first > 20.0:                           # large synthetic LP: one step is minutes of CPU; bound the sample
            a.warmup, a.steps = 0, 1
        for _ in range(a.warmup):
            cpu.step(it)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            cpu.step(it)
        dt = (time.perf_counter() - t0) / a.steps
        v = flops / dt / 1e9
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "GFLOP/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "netlib LP fixture + oracle-generated iterate", "config": config,
            "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": 1, "kind": cpu.kind,
                             "sample": f"{a.steps} KKT steps, {dt * 1e3:.2f} ms/step, 1 host core (the reference is single-threaded)"},
            "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "flops_per_step": flops}))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    vb = _load("vbkkt", ROOT / "linear-programming-vanderbei_b200" / "__init__.py")
    lib = vb.load(os.environ.get("VBK_LIB"))
    if lib.vbk_device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the KKT path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = vb.MODE_STRICT if a.mode == "strict" else vb.MODE_FAST
    import harness as H
    K = H.kkt_for(vb, lib, lp, device=local_rank, mode=mode)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(K.stream, device=dev)

    def dten(v):
        return torch.from_numpy(np.ascontiguousarray(v)).to(dev)

    E_d, D_d = dten(it["E"]), dten(it["D"])
    ry_d, rx_d, b_d, c_d = dten(it["rhs_y"]), dten(it["rhs_x"]), dten(-lp.b), dten(-lp.c)
    fy, fx, gy, gx = (torch.empty_like(t) for t in (ry_d, rx_d, b_d, c_d))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    passes_seen = []

    def step_dev():
        with torch.cuda.stream(stream):
            fy.copy_(ry_d); fx.copy_(rx_d); gy.copy_(b_d); gx.copy_(c_d)
        K.factor_dev(E_d.data_ptr(), D_d.data_ptr())
        # both systems of the iteration (hsd.c:223,228) through the two-right-hand-side sweeps, as the product's own
        # METHOD plugin runs them; the e2e arm below makes the reference's two separate forwardbackward calls
        K.solve2_dev(E_d.data_ptr(), D_d.data_ptr(), fy.data_ptr(), fx.data_ptr(), gy.data_ptr(), gx.data_ptr())
        passes_seen.append(K.last_passes2(0) + K.last_passes2(1))

    for _ in range(a.warmup):
        step_dev()
    K.sync()
    # parity gate inside the bench: the solution of the first right-hand side equals the reference's
    sol_y = fy.cpu().numpy()
    synthetic = "sol_y" not in it
    if synthetic:
        # synthetic workload: no stored reference solution; gate on the KKT residual of the GPU solution
        err = kkt_residual(lp, it, sol_y, fx.cpu().numpy())
        bit_equal = False
        assert err < 1e-8, f"KKT residual of the GPU solution is {err:.3e} (relative)"
        a.no_strict = True                          # strict mode is latency-bound by design: hours at this size
    else:
        ref_y = it["sol_y"]
        err = float(np.max(np.abs(sol_y - ref_y)) / max(np.max(np.abs(ref_y)), 1e-300))
        bit_equal = bool(np.array_equal(sol_y, ref_y))
        if a.mode == "strict":
            assert bit_equal, f"strict mode lost bit-parity with the reference (max rel err {err:.3e})"
        else:
            sol_x = fx.cpu().numpy()
            err_x = float(np.max(np.abs(sol_x - it["sol_x"])) / max(np.max(np.abs(it["sol_x"])), 1e-300))
            err = max(err, err_x)
            assert err < 1e-7, f"fast mode: KKT-step solution differs from the reference's by {err:.3e} (relative)"

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = K.launches
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    fac_ms = []
    for s in range(a.steps):
        flush.fill_(s & 0xFF)                       # evict L2 (126 MB) between timed steps, untimed
        torch.cuda.synchronize()
        ev[s][0].record(stream)
        step_dev()
        ev[s][1].record(stream)
        K.sync()
        fac_ms.append(float(lib.vbk_kkt_last_factor_kernel_ms(K.h)))
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(step_ms))
    launches = K.launches - launches0
    if dist:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / a.steps
    raw_per_step = float(np.mean(passes_seen[-a.steps:]))
    flops_step = work_model(K.narth, K.lnz, N, nz, raw_per_step)
    value = world * flops_step / (ms_per_step * 1e-3) / 1e9

    # e2e: host buffers (pinned) through the reference-facing C ABI: H2D + D2H inside the timed region
    def pinned(v):
        t = torch.empty(v.shape, dtype=torch.float64, pin_memory=True)
        t.copy_(torch.from_numpy(np.ascontiguousarray(v)))
        return t
    hE, hD, hry, hrx, hb, hc = (pinned(v) for v in (it["E"], it["D"], it["rhs_y"], it["rhs_x"], -lp.b, -lp.c))
    wy, wx, wgy, wgx = (torch.empty_like(t).pin_memory() for t in (hry, hrx, hb, hc))
    dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))

    # The reference's own call sequence (hsd.c:218-228) on the B1 plugin symbols themselves: ldltfac with the m/n and
    # A/At arguments swapped exactly as hsd.c passes them, then two forwardbackward calls.  The process-global factor
    # object behind these symbols (ldlt.c:108-120) is separate from the handle K above: it analyses once, here.
    kAt_h, iAt_h, At_h = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    kA_h, iA_h, A_h = (np.ascontiguousarray(lp.kA, dtype=np.int32), np.ascontiguousarray(lp.iA, dtype=np.int32),
                       np.ascontiguousarray(lp.A, dtype=np.float64))
    kAt_h, iAt_h, At_h = (np.ascontiguousarray(kAt_h, dtype=np.int32), np.ascontiguousarray(iAt_h, dtype=np.int32),
                          np.ascontiguousarray(At_h, dtype=np.float64))
    lib.inv_clo()
    lib.vbk_set_device(local_rank)
    lib.vbk_set_mode(mode)

    def step_host():
        wy.copy_(hry); wx.copy_(hrx); wgy.copy_(hb); wgx.copy_(hc)
        lib.ldltfac(lp.n, lp.m, H.ptr_i(kAt_h), H.ptr_i(iAt_h), H.ptr_d(At_h), dp(hE), dp(hD),
                    H.ptr_i(kA_h), H.ptr_i(iA_h), H.ptr_d(A_h), 1)                       # hsd.c:218
        lib.forwardbackward(dp(hE), dp(hD), dp(wy), dp(wx))                              # hsd.c:223
        lib.forwardbackward(dp(hE), dp(hD), dp(wgy), dp(wgx))                            # hsd.c:228

    for _ in range(2):
        step_host()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    if dist:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    clocks = sampler.summary()
    lib.inv_clo()
    assert np.array_equal(wy.numpy(), sol_y), "host-buffer path and device-pointer path disagree"
    e2e = {"value": world * flops_step / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(8 * N + 2 * 16 * N), "d2h_bytes_per_step": int(2 * 8 * N),
           "api": ("ldltfac + 2 x forwardbackward (the reference's plugin symbols, src/ipo/ldlt.h:1-20), pinned host arrays; the library "
                   "solves the iteration's constant second right-hand side speculatively with the first call's sweeps and answers the "
                   "second call from that result when its inputs are bit-identical ($VBK_SPECULATE=0: 181 ms)")}

    def rank0_line():
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            cpu = cpu_measure(lp, it, flops_step, a.cpu_budget, 200)
        # roofline of the dominant kernel (numeric LDL^T): algorithmic bytes / flops per launch
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        fac_s = float(np.mean(fac_ms)) * 1e-3
        b_fac = 12.0 * K.lnz + 8.0 * K.lnz + 2 * 12.0 * nz + 24.0 * N
        fp64_peak = float(lib.vbk_measure_fp64_tflops(local_rank))
        fp64 = {"achieved_tflops": K.narth / fac_s / 1e12, "peak_tflops": fp64_peak,
                "frac": K.narth / fac_s / 1e12 / max(fp64_peak, 1e-9),
                "peak_source": "measured here (DFMA yardstick kernel, vbk_measure_fp64_tflops; DMMA measures the same "
                               "37 TFLOP/s, profiles/ubench/ubench.cu); MEASURED_PEAKS.json has no FP64 entry"}
        hbm = {"achieved_gbs": b_fac / fac_s / 1e9, "peak_gbs": hbm_peak, "frac": b_fac / fac_s / 1e9 / hbm_peak,
               "peak_source": peak_src}
        # a factorisation is held to the FP64 roofline iff its arithmetic intensity (reference flop count over
        # algorithmic bytes) exceeds peak_FP64 / peak_HBM (~6 flop/B), otherwise to HBM
        intensity = K.narth / b_fac
        fp64_bound = intensity > fp64_peak * 1e12 / (hbm_peak * 1e9)
        kname = ("numeric factorisation = k_factor_pipe (strict LDL^T: pipelined slice tasks, csrc/vbk_strict_factor.cuh)"
                 if a.mode == "strict" else
                 "numeric factorisation = k_sparse_level / k_factor_pipe (sparse columns) + k_schur_window2 + per 128-column panel "
                 "k_panel_diag, k_panel_rows_m, k_dense_update_m<64,64> (strip), k_dense_update_m<128,64> (DMMA rank-128 update)")
        # dependent-chain model of the strict factorisation: one rounded FP64 addition per contributor link of the
        # critical path (~Lnz links), 8.1 cycles each (profiles/r01_ubench.txt) at the sampled SM clock
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        floor_ns = 8.1 / (sm_mhz * 1e-3)
        chain = {"links": int(K.lnz), "ns_per_link": fac_s * 1e9 / max(K.lnz, 1), "floor_ns_per_link": floor_ns,
                 "frac_of_floor": floor_ns / (fac_s * 1e9 / max(K.lnz, 1)),
                 "note": "strict mode is a chain of ~Lnz dependent rounded additions by construction (SURVEY 8d): this is the bound it is held to"}
        # DRAM traffic of one launch of the dominant kernel: not measurable here (no profiler inside a timed run); the
        # committed ncu capture of the same workload and mode is quoted when there is one
        traffic, traffic_note = None, "no ncu capture of this workload/mode in profiles/r02_traffic.json"
        try:
            tj = json.load(open(ROOT / "profiles" / "r02_traffic.json"))
            ent = tj.get(a.workload, {}).get(a.mode)
            if ent:
                traffic = ent["bytes_per_launch"]
                traffic_note = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one {ent['kernel']} launch ({ent['source']}); "
                                f"L2 -> SM traffic of the same launch: {ent['l2_to_sm_bytes'] / 1e9:.0f} GB, all L2 hits")
        except (OSError, ValueError, KeyError):
            pass
        roofline = {"kernel": kname, "bound": "tensor" if fp64_bound else "hbm",
                    "achieved": fp64["achieved_tflops"] if fp64_bound else hbm["achieved_gbs"],
                    "peak": fp64_peak if fp64_bound else hbm_peak, "unit": "TFLOP/s" if fp64_bound else "GB/s",
                    "frac": fp64["frac"] if fp64_bound else hbm["frac"],
                    "traffic": traffic, "traffic_note": traffic_note,
                    "algorithmic_bytes": b_fac,
                    "peak_source": fp64["peak_source"] if fp64_bound else peak_src,
                    "flop_per_byte": intensity, "kernel_ms": fac_s * 1e3, "share_of_step": fac_s * 1e3 / ms_per_step,
                    "fp64": fp64, "hbm": hbm, "dependent_chain": chain if a.mode == "strict" else None,
                    "note": ("strict mode replays the reference's rounding order: latency-bound by design (SURVEY 8d), see dependent_chain"
                             if a.mode == "strict" else
                             "flops = the reference's own count narth (ldlt.c:1243-1248); the dense window executes more "
                             "(padding to rho = 0.06), so the pipe is busier than this fraction says")}

        def side_mode(other):
            """the same step in the other arithmetic mode"""
            K2 = H.kkt_for(vb, lib, lp, device=local_rank, mode=other)
            s2 = torch.cuda.ExternalStream(K2.stream, device=dev)
            def step2():
                with torch.cuda.stream(s2):
                    fy.copy_(ry_d); fx.copy_(rx_d); gy.copy_(b_d); gx.copy_(c_d)
                K2.factor_dev(E_d.data_ptr(), D_d.data_ptr())
                K2.solve2_dev(E_d.data_ptr(), D_d.data_ptr(), fy.data_ptr(), fx.data_ptr(), gy.data_ptr(), gx.data_ptr())
            torch.cuda.synchronize()
            for _ in range(3):
                step2()
            K2.sync()
            ry2, rx2 = fy.cpu().numpy(), fx.cpu().numpy()
            bit = bool(np.array_equal(ry2, it["sol_y"]) and np.array_equal(rx2, it["sol_x"]))
            err2 = float(max(np.max(np.abs(ry2 - it["sol_y"])) / max(np.max(np.abs(it["sol_y"])), 1e-300),
                             np.max(np.abs(rx2 - it["sol_x"])) / max(np.max(np.abs(it["sol_x"])), 1e-300)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = 5
            e0.record(s2)
            for _ in range(nrep):
                step2()
            e1.record(s2); K2.sync(); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / nrep
            fk = float(lib.vbk_kkt_last_factor_kernel_ms(K2.h))
            K2.close()
            return {"ms_per_step": ms2, "value": flops_step / (ms2 * 1e-3) / 1e9, "unit": "GFLOP/s",
                    "bit_equal_to_reference": bit, "max_rel_err_kkt_step": err2, "factor_kernel_ms": fk,
                    "fp64_frac": K.narth / (fk * 1e-3) / 1e12 / max(fp64_peak, 1e-9)}

        strict = fast = None
        if not a.no_strict and world == 1 and not synthetic:
            if a.mode == "fast":
                strict = side_mode(vb.MODE_STRICT)
                strict["note"] = "same step, strict arithmetic mode: reproduces the reference's golden logs byte for byte"
            else:
                fast = side_mode(vb.MODE_FAST)
                fast["note"] = ("same step, fast mode: dense-window factorisation with re-associated sums.  NOT a parity mode: full "
                                "solves stay inside the north_star tolerances on the robust problems only (profiles/r01_fast_sweep.md: "
                                "55 of 86 netlib LPs; see solve_time.in_tolerance for the three BASELINE names)")

        # full solves, device-resident METHOD=hsd (BASELINE's metric starts with "LP solve time"): strict, fast, and the
        # reference on one host core where that is affordable inside the bench
        solve_time = None
        if world == 1 and not a.no_solve_time and not synthetic:
            solve_time = {}
            for name in ("25fv47", "pilot87", a.workload if a.workload not in ("25fv47", "pilot87") else "dfl001"):
                try:
                    lpx = H.load_fixture(name)
                except Exception:
                    continue
                ent = {}
                t0 = time.perf_counter()
                st, log, x, y, _ = H.solve_via(vb, lib, lpx, "hsd", mode=vb.MODE_STRICT)
                ent["strict_s"] = time.perf_counter() - t0
                ent["strict_log_identical_to_golden"] = bool(log == str(lpx.extra["hsd_log"]))
                ent["strict_x_bit_equal"] = bool(np.array_equal(x, lpx.extra["hsd_x"]))
                ent["iterations"] = len(H.iteration_lines(log))
                t0 = time.perf_counter()
                stf, logf, xf, yf, _ = H.solve_via(vb, lib, lpx, "hsd", mode=vb.MODE_FAST)
                ent["fast_s"] = time.perf_counter() - t0
                obj_r = float(lpx.c @ lpx.extra["hsd_x"])
                rel = abs(float(lpx.c @ xf) - obj_r) / max(1.0, abs(obj_r))
                dit = abs(len(H.iteration_lines(logf)) - ent["iterations"])
                ent["fast"] = {"status": int(stf), "reference_status": int(lpx.extra["hsd_status"]), "iteration_diff": int(dit), "objective_rel_err": rel,
                               "in_tolerance": bool(stf == int(lpx.extra["hsd_status"]) and dit <= 1 and rel <= 1e-8)}
                if name == "25fv47":
                    # fresh process: the reference keeps ONE factor object per process (ldlt.c:108-120) and would reuse
                    # the symbolic phase of whatever LP it saw first
                    code = ("import sys, time; sys.path.insert(0, %r); import harness as H; ref = H.load_ref('hsd'); "
                            "lp = H.load_fixture(%r); t0 = time.perf_counter(); H.call_solver(ref.solver, lp); "
                            "print('CPU_S', time.perf_counter() - t0)" % (str(ROOT / "tests"), name))
                    try:
                        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
                        for ln in r.stdout.splitlines():
                            if ln.startswith("CPU_S"):
                                ent["cpu_s"] = float(ln.split()[1])
                                ent["cpu"] = "compiled reference, 1 host core, measured here (own process)"
                    except Exception:
                        pass
                if "cpu_s" not in ent and cpu is not None and name == a.workload:
                    ent["cpu_s_estimate"] = cpu["ms_per_step"] * 1e-3 * ent["iterations"]
                    ent["cpu"] = "estimate: measured reference ms per KKT step (cpu_baseline) x iterations; the full CPU solve takes minutes"
                solve_time[name] = ent

        out = {"metric": metric, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": ("synthetic multicommodity LP (generator seed 1) + synthetic iterate (seed 20)" if a.workload.startswith("mcf")
                                        else "netlib LP fixture + oracle-generated hsd iterate (tests/golden)"),
               "config": config, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
               "cpu_baseline": cpu, "strict_mode": strict, "fast_mode": fast, "solve_time": solve_time,
               "flops_per_step": flops_step, "rawsolves_per_step": raw_per_step,
               "parity": ({"kkt_residual_rel": err, "note": "synthetic LP: residual of the solved KKT system, no stored reference solution"}
                          if a.workload.startswith("mcf") else {"bit_equal_to_reference": bit_equal, "max_rel_err": err}),
               "symbolic": {"N": N, "lnz": K.lnz, "narth": K.narth, "levels": K.nlevels, "supernodes": K.nsupernodes}}
        return out

    out = rank0_line() if rank == 0 else None
    K.close()
    if world > 1 and not a.no_sharded:
        # the two workloads that shard, measured on the same ranks in the same run and attached to the line
        torch.cuda.synchronize()
        dist.barrier()
        sb = run_batch(a, vb, lib, rank, local_rank, world)
        dist.barrier()
        a.rowblock_steps = max(a.rowblock_steps, 10)
        sr = run_rowblock(a, vb, lib, rank, local_rank, world, partition="rows")
        dist.barrier()
        sc = run_rowblock(a, vb, lib, rank, local_rank, world, partition="coupled")
        if rank == 0:
            out["sharded"] = {"batch": sb, "rowblock": sr, "rowblock_coupled": sc,
                              "note": "BASELINE configs 4 and 5 on the same N ranks; the primary metric above is N independent replicas of the KKT step. "
                                      "rowblock = equal row blocks, x and y all-gathered (bit-identical smx); rowblock_coupled = column blocks with the "
                                      "coupling rows' partial sums all-reduced (fast-mode variant: those rows are re-associated)"}
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
