"""Multi-GPU host logic on CPU: world_size-2 `gloo` process groups, with the kernels of the product
sources running on the host thread emulator (tests/emu) -- the partitioning, the collectives and the
result gathering are exactly the code the NCCL runs use (batch.py, rowblock.py); only the device differs.
The real-GPU versions are in tests/test_gpu.py / bench.py --gpus N."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

import harness as H

ROOT = Path(__file__).resolve().parents[1]
EMU = ROOT / "tests" / "emu" / "libvbkkt_emu.so"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "tests"))
    import conftest
    return conftest._load_pkg(), dist


# ------------------------------------------------------------------------------------------------
def _rowblock_worker(rank, world, port, out):
    vb, dist = _init(rank, world, port)
    try:
        lib = vb.load(EMU)
        lp = H.load_fixture("afiro")
        kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
        ops = vb.rowblock.RowBlockOps(lib, lp.m, lp.n, lp.kA, lp.iA, lp.A, kAt, iAt, At, "cpu")
        rng = np.random.default_rng(7)
        x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
        w = rng.standard_normal(lp.m)
        rho = ops.A_x(ops.local_x(x)).numpy()[: ops.r1 - ops.r0].copy()
        sig = ops.At_y(ops.local_y(y)).numpy()[: ops.c1 - ops.c0].copy()
        d = ops.dots([(ops.local_x(x), ops.local_x(x)), (ops.local_y(y), ops.local_y(w))]).numpy().copy()
        mx = ops.absmax([ops.local_x(x), ops.local_y(y)]).numpy().copy()
        # the fused step (both all-gathers in flight, one small all-gather for the six scalars)
        import torch
        rho2 = torch.zeros(ops.rows_per, dtype=torch.float64)
        sig2 = torch.zeros(ops.cols_per, dtype=torch.float64)
        sums, maxes = ops.step(ops.local_x(x), ops.local_y(y), rho2, sig2)
        np.savez(out + f".{rank}.npz", rho=rho, sig=sig, d=d, mx=mx, r=np.array([ops.r0, ops.r1, ops.c0, ops.c1]),
                 rho2=rho2.numpy()[: ops.r1 - ops.r0].copy(), sig2=sig2.numpy()[: ops.c1 - ops.c0].copy(),
                 sums=sums.numpy().copy(), maxes=maxes.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rowblock_smx_dot_maxv_gloo(tmp_path, emu_lib, oracle_lib, world):
    """A x and A^T y assembled from the ranks' blocks are BIT-identical to the oracle's smx (the row sums
    keep the reference's order, linalg.c:62-70); dot products agree to rounding, max-norms exactly."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "rb")
    mp.spawn(_rowblock_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    lp = H.load_fixture("afiro")
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
    w = rng.standard_normal(lp.m)
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    rho_ref, sig_ref = np.zeros(lp.m), np.zeros(lp.n)
    oracle_lib.kko_smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(rho_ref))
    oracle_lib.kko_smx(lp.n, lp.m, H.ptr_d(At), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(y), H.ptr_d(sig_ref))
    rho, sig = np.zeros(lp.m), np.zeros(lp.n)
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        r0, r1, c0, c1 = z["r"]
        rho[r0:r1], sig[c0:c1] = z["rho"], z["sig"]
        assert abs(z["d"][0] - x @ x) <= 1e-13 * (x @ x)
        assert abs(z["d"][1] - y @ w) <= 1e-13 * max(1.0, np.abs(y * w).sum())
        assert z["mx"][0] == np.abs(x).max() and z["mx"][1] == np.abs(y).max()
        # fused step: same products bit for bit, the six scalars identical on every rank
        assert np.array_equal(z["rho2"], z["rho"]) and np.array_equal(z["sig2"], z["sig"])
        z0 = np.load(out + ".0.npz")
        assert np.array_equal(z["sums"], z0["sums"]) and np.array_equal(z["maxes"], z0["maxes"])
        ref = np.array([x @ sig_ref, y @ rho_ref, rho_ref @ rho_ref, sig_ref @ sig_ref])
        assert np.allclose(z["sums"], ref, rtol=1e-12, atol=1e-12)
        assert z["maxes"][0] == np.abs(rho_ref).max() and z["maxes"][1] == np.abs(sig_ref).max()
    assert np.array_equal(rho, rho_ref) and np.array_equal(sig, sig_ref)


# ------------------------------------------------------------------------------------------------
def _coupled_worker(rank, world, port, out):
    vb, dist = _init(rank, world, port)
    try:
        import torch
        lib = vb.load(EMU)
        lp = vb.workloads.multicommodity_lp(4, 5)          # 5 commodities on a 4 x 4 grid: conservation rows per commodity, 48 coupling rows
        kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
        ops = vb.rowblock.CoupledOps(lib, lp.m, lp.n, lp.kA, lp.iA, lp.A, kAt, iAt, At, "cpu")
        rng = np.random.default_rng(11)
        x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
        rho = torch.zeros(ops.rows_per, dtype=torch.float64)
        sig = torch.zeros(ops.cols_per, dtype=torch.float64)
        sums, maxes = ops.step(ops.local_x(x), ops.local_y(y), rho, sig)
        np.savez(out + f".{rank}.npz", rho=rho.numpy().copy(), sig=sig.numpy()[: ops.c1 - ops.c0].copy(), c=np.array([ops.c0, ops.c1]),
                 local_rows=ops.local_rows, shared_rows=ops.shared_rows, sums=sums.numpy().copy(), maxes=maxes.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_coupled_partition_gloo(tmp_path, vbkkt, emu_lib, oracle_lib, world):
    """The structure-following partition (CoupledOps): local rows of A x and all of A^T y are BIT-identical to the
    oracle's smx, the shared (coupling) rows agree to rounding (their sums are re-associated by column block), every
    row is local to exactly one rank or shared, and all ranks hold the same six scalars."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "cp")
    mp.spawn(_coupled_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    lp = vbkkt.workloads.multicommodity_lp(4, 5)
    rng = np.random.default_rng(11)
    x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    rho_ref, sig_ref = np.zeros(lp.m), np.zeros(lp.n)
    oracle_lib.kko_smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(rho_ref))
    oracle_lib.kko_smx(lp.n, lp.m, H.ptr_d(At), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(y), H.ptr_d(sig_ref))
    zs = [np.load(out + f".{r}.npz") for r in range(world)]
    covered = np.zeros(lp.m, dtype=int)
    sig = np.zeros(lp.n)
    shared = zs[0]["shared_rows"]
    assert 0 < len(shared) < lp.m // 2
    covered[shared] += 1
    for z in zs:
        lr = z["local_rows"]
        covered[lr] += 1
        assert np.array_equal(z["shared_rows"], shared)
        assert np.array_equal(z["rho"][: len(lr)], rho_ref[lr])                                    # local rows: the reference's bits
        assert np.allclose(z["rho"][len(lr):], rho_ref[shared], rtol=1e-13, atol=1e-13)           # coupling rows: re-associated
        assert np.array_equal(z["rho"][len(lr):], zs[0]["rho"][len(zs[0]["local_rows"]):])        # ... but the same on every rank
        c0, c1 = z["c"]
        sig[c0:c1] = z["sig"]
        assert np.array_equal(z["sums"], zs[0]["sums"]) and np.array_equal(z["maxes"], zs[0]["maxes"])
    assert np.all(covered == 1)
    assert np.array_equal(sig, sig_ref)                                                            # A^T y: bit-identical
    ref = np.array([x @ sig_ref, y @ rho_ref, rho_ref @ rho_ref, sig_ref @ sig_ref])
    assert np.allclose(zs[0]["sums"], ref, rtol=1e-12, atol=1e-12)
    assert np.isclose(zs[0]["maxes"][0], np.abs(rho_ref).max(), rtol=1e-13) and zs[0]["maxes"][1] == np.abs(sig_ref).max()


# ------------------------------------------------------------------------------------------------
def _tiny_lp(vb, i):
    return vb.workloads.random_sparse_lp(seed=i, m=12, n=20, nnz_per_col=3)


def _batch_worker(rank, world, port, out):
    vb, dist = _init(rank, world, port)
    try:
        lib = vb.load(EMU)
        summary, local = vb.batch.solve_batch(lib, lambda i: _tiny_lp(vb, i), 5, method="hsd", device=0,
                                              mode=vb.MODE_STRICT, nstreams=2)
        np.savez(out + f".{rank}.npz", summary=summary, nlocal=len(local))
    finally:
        dist.destroy_process_group()


def test_batch_of_lps_sharded_over_two_ranks_gloo(tmp_path, vbkkt, emu_lib, oracle_lib):
    """5 random LPs dealt i mod 2 over two ranks, two solver streams per rank: every rank ends with the
    whole batch's results, equal to the oracle's METHOD run LP by LP (strict mode: same iteration count,
    same objective bits)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "batch")
    mp.spawn(_batch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z0, z1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert int(z0["nlocal"]) == 3 and int(z1["nlocal"]) == 2
    assert np.array_equal(z0["summary"][:, :4], z1["summary"][:, :4])
    for i in range(5):
        lp = _tiny_lp(vbkkt, i)
        st, log, x, y = H.call_solver(oracle_lib.kko_solver_hsd, lp)
        lines = H.iteration_lines(log)
        assert z0["summary"][i, 0] == st
        # an optimal run prints one line per iteration and breaks in the iteration after the last line
        assert z0["summary"][i, 1] == (len(lines) if st != 5 else 200)
        assert np.isclose(z0["summary"][i, 2], float(lp.c @ x) + lp.f, rtol=1e-13, atol=0)
        assert np.isclose(z0["summary"][i, 3], float(lp.b @ y) + lp.f, rtol=1e-13, atol=0)


def test_batch_single_process_matches_individual_solves(vbkkt, emu_lib):
    """vbk_solve_batch with concurrent solver streams returns what solve_lp returns one LP at a time."""
    lps = [_tiny_lp(vbkkt, i) for i in range(4)]
    res = vbkkt.batch.solve_local(emu_lib, lps, method="hsd", mode=vbkkt.MODE_STRICT, nstreams=3)
    for lp, r in zip(lps, res):
        with H.capture_stdout():
            st, x, y, _ = vbkkt.solve_lp("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, lib=emu_lib)
        assert r["status"] == st and np.array_equal(r["x"], x) and np.array_equal(r["y"], y)
        assert r["iterations"] > 0 and np.isclose(r["primal_obj"], float(lp.c @ x) + lp.f, rtol=1e-13, atol=0)
