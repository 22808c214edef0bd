"""GPU parity tests (run with `-m gpu` on a B200): everything goes through the C ABI of the
nvcc-built libvbkkt.so and is compared with the oracle (oracle/libkkt_oracle.so, checker only) and
with the committed golden fixtures of the reference's own logs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
import parity as P

pytestmark = pytest.mark.gpu

# BASELINE config 2: strict mode must reproduce the golden log of EVERY netlib fixture byte for byte (robust and fragile
# problems of SURVEY.md H2 alike: 86 LPs, dfl001 and pds-06 included; ~105 s of GPU time in all, profiles/r02_strict_sweep.jsonl)
FULL_HSD = H.fixture_names()
# fast mode is a tolerance mode: these LPs stayed inside the north_star tolerances (status, iterations +-1, objective 1e-8)
# in both committed sweeps (profiles/r01_fast_sweep.jsonl, r02_fast_sweep.jsonl); the others are NOT claimed
FAST_IN_TOLERANCE = ['25fv47', 'adlittle', 'afiro', 'bandm', 'beaconfd', 'blend', 'boeing1', 'boeing2', 'bore3d', 'cre-a', 'cre-c',
                     'czprob', 'd6cube', 'degen2', 'finnis', 'fit1d', 'fit1p', 'fit2p', 'ganges', 'grow22', 'grow7', 'israel', 'kb2',
                     'recipe', 'sc105', 'sc205', 'sc50a', 'sc50b', 'scfxm2', 'scfxm3', 'scorpion', 'scrs8', 'scsd1', 'scsd8', 'sctap1',
                     'sctap2', 'sctap3', 'seba', 'ship04l', 'ship04s', 'ship08l', 'ship08s', 'ship12l', 'ship12s', 'standata',
                     'standgub', 'standmps', 'stocfor1', 'stocfor2', 'wood1p', 'woodw']
FULL_INTPT = ["afiro", "adlittle", "blend", "sc50a", "share2b", "israel", "25fv47"]


def test_device_is_blackwell(gpu_lib):
    assert gpu_lib.vbk_device_count() >= 1
    assert b"sm_100a" in gpu_lib.vbk_version()


def test_linalg_bit_exact(vbkkt, gpu_lib, oracle_lib):
    P.check_linalg(vbkkt, gpu_lib, oracle_lib, sizes=(0, 1, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 100003))


@pytest.mark.parametrize("name", ["afiro", "25fv47", "fit2p", "ken-07"])
def test_transpose_and_smx(vbkkt, gpu_lib, oracle_lib, name):
    P.check_transpose_and_smx(vbkkt, gpu_lib, oracle_lib, H.load_fixture(name))


def test_ragged_transpose(vbkkt, gpu_lib):
    P.check_ragged_transpose(vbkkt, gpu_lib)


@pytest.mark.parametrize("name,method,it", [
    ("afiro", "hsd", 3), ("afiro", "hsd", 26), ("afiro", "intpt", 14),
    ("25fv47", "hsd", 0), ("25fv47", "hsd", 40), ("25fv47", "hsd", 85),
    ("sctap3", "hsd", 30), ("fit2p", "hsd", 20), ("israel", "hsd", 12), ("grow7", "intpt", 8),
    ("pilot87", "hsd", 60),
])
def test_kkt_step_bit_exact(vbkkt, gpu_lib, oracle_lib, name, method, it):
    """ldltfac + forwardbackward + rawsolve on captured iterates: L, diag, mark, ndep, the number
    of refinement passes and the solution are bit-equal to the oracle's."""
    P.check_kkt_step(vbkkt, gpu_lib, oracle_lib, H.load_fixture(name), method, it)


@pytest.mark.parametrize("name", FULL_HSD)
def test_hsd_log_and_solution_match_golden(vbkkt, gpu_lib, name):
    """BASELINE.json config 2, the whole suite: device-resident METHOD=hsd in strict mode prints the reference's golden
    log byte for byte and returns bit-equal x and y (fixtures without an hsd run -- free variables, status 3 -- are skipped)."""
    lp = H.load_fixture(name)
    if "hsd_log" not in lp.extra:
        pytest.skip("the reference does not iterate on this LP (free variables)")
    P.check_full_solve(vbkkt, gpu_lib, lp, "hsd")


@pytest.mark.parametrize("name", FULL_INTPT)
def test_intpt_log_and_solution_match_reference(vbkkt, gpu_lib, name):
    """BASELINE.json config 1 (afiro: 25 lines, status 0) and friends."""
    lp = H.load_fixture(name)
    P.check_full_solve(vbkkt, gpu_lib, lp, "intpt")
    if name == "afiro":
        assert len(H.iteration_lines(str(lp.extra["intpt_log"]))) == 25


@pytest.mark.parametrize("name", ["afiro", "adlittle", "blend", "sc50a", "sc50b", "sc105", "share2b", "kb2", "israel", "stocfor1",
                                  "scsd1", "e226", "bandm", "sctap1", "25fv47"])
def test_hsdls_log_and_solution_match_reference(vbkkt, gpu_lib, name):
    """METHOD = hsdls (src/ipo/hsdls.c:37-336) device-resident; oracle = the compiled reference's committed
    output (no golden logs exist for it): byte-identical log, bit-equal x and y."""
    P.check_hsdls(vbkkt, gpu_lib, H.load_fixture(name))


def test_north_star_tolerances_hold(vbkkt, gpu_lib):
    for name in ("afiro", "25fv47"):
        lp = H.load_fixture(name)
        st, log, x, y, _ = H.solve_via(vbkkt, gpu_lib, lp, "hsd")
        P.north_star_tolerances(lp, "hsd", x, y, st, log)


@pytest.mark.parametrize("name,it", [("afiro", 5), ("25fv47", 20), ("israel", 12), ("pilot87", 20), ("fit2p", 10)])
def test_fast_mode_kkt_step(vbkkt, gpu_lib, oracle_lib, name, it):
    """Fast mode (dense-window factorisation, re-associated sums) agrees with the oracle to rounding
    on iterates without dependent pivots."""
    P.check_kkt_step_fast(vbkkt, gpu_lib, oracle_lib, H.load_fixture(name), "hsd", it)


@pytest.mark.parametrize("name", FAST_IN_TOLERANCE)
def test_fast_mode_full_solve_north_star_tolerances(vbkkt, gpu_lib, name):
    """The 51 LPs fast mode is claimed for, solved end to end in fast mode: status, iterations +-1, objective 1e-8."""
    P.check_full_solve_fast(vbkkt, gpu_lib, H.load_fixture(name))


def test_strict_mode_big_iterates_bit_equal(vbkkt, gpu_lib):
    """The largest netlib LPs (dfl001: Lnz 7.2 M, window 2766; pds-06) at full size: the strict KKT
    step returns the reference's solution bit for bit (fixture from the pinned oracle)."""
    for name in ("dfl001", "pds-06"):
        lp = H.load_fixture(name)
        z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
        K = H.kkt_for(vbkkt, gpu_lib, lp)
        K.factor(z["E"], z["D"])
        sy, sx, _ = K.solve(z["E"], z["D"], z["rhs_y"], z["rhs_x"])
        assert np.array_equal(sy, z["sol_y"]) and np.array_equal(sx, z["sol_x"])
        K.close()


def test_fast_mode_big_iterates_match_reference_solution(vbkkt, gpu_lib):
    """dfl001 / pds-06 at hsd iterate 20 (committed fixtures from the pinned oracle)."""
    for name in ("dfl001", "pds-06"):
        lp = H.load_fixture(name)
        z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
        K = H.kkt_for(vbkkt, gpu_lib, lp, mode=vbkkt.MODE_FAST)
        K.factor(z["E"], z["D"])
        sy, sx, _ = K.solve(z["E"], z["D"], z["rhs_y"], z["rhs_x"])
        assert P._rel(sy, z["sol_y"]) < 1e-6 and P._rel(sx, z["sol_x"]) < 1e-6
        K.close()


@pytest.mark.parametrize("name,it", [("afiro", 26), ("25fv47", 20), ("pds-02", 20), ("dfl001", 20)])
def test_fast_mode_sparse_columns_bit_exact(vbkkt, gpu_lib, oracle_lib, name, it):
    """Fast mode keeps the reference's arithmetic for the sparse columns j < T (level-scheduled kernels of
    vbk_sparse_level.cuh: contributors in the reference's list order, separately rounded products and sums, exact pivot
    rule): their part of L, diag and mark equals the strict factor -- itself bit-equal to the reference -- bit for
    bit, dependent pivots included (afiro iterate 26 has two)."""
    lp = H.load_fixture(name)
    if name in ("afiro", "25fv47"):
        E, D, *_ = H.capture_step(oracle_lib, lp, "hsd", it)
    else:
        z = np.load(H.GOLDEN / "iterates" / f"{name}_it{it}.npz") if name == "dfl001" else None
        E, D = (z["E"], z["D"]) if z is not None else (np.random.default_rng(3).uniform(0.1, 10.0, lp.m),
                                                      np.random.default_rng(4).uniform(0.1, 10.0, lp.n))
    Ks = H.kkt_for(vbkkt, gpu_lib, lp)
    Kf = H.kkt_for(vbkkt, gpu_lib, lp, mode=vbkkt.MODE_FAST)
    try:
        Ks.factor(E, D)
        Kf.factor(E, D)
        Ls, ds, ms = Ks.get_factor()
        Lf, df, mf = Kf.get_factor()
        T = Kf.dim - Kf.window
        nsp = int(Kf.kAAt[T])
        assert T > 0 and nsp > 0
        assert np.array_equal(Lf[:nsp], Ls[:nsp])
        assert np.array_equal(df[:T], ds[:T]) and np.array_equal(mf[:T], ms[:T])
    finally:
        Ks.close()
        Kf.close()


@pytest.mark.parametrize("R,K", [(8, 5), (12, 8)])
def test_fast_mode_multicommodity_kkt_step(vbkkt, gpu_lib, R, K):
    """BASELINE config 3 at sizes the strict mode finishes in seconds: the fast-mode KKT step on the synthetic
    multicommodity LP leaves a residual at rounding level and agrees with the strict (= reference) solution."""
    import scipy.sparse as sp
    lp = vbkkt.workloads.multicommodity_lp(R, K)
    rng = np.random.default_rng(20)
    E, D = 10.0 ** rng.uniform(-3, 3, lp.m), 10.0 ** rng.uniform(-3, 3, lp.n)
    ry, rx = rng.standard_normal(lp.m), rng.standard_normal(lp.n)
    Ks = H.kkt_for(vbkkt, gpu_lib, lp)
    Kf = H.kkt_for(vbkkt, gpu_lib, lp, mode=vbkkt.MODE_FAST)
    try:
        Ks.factor(E, D); Kf.factor(E, D)
        sy, sx, _ = Ks.solve(E, D, ry, rx)
        fy, fx, _ = Kf.solve(E, D, ry, rx)
        A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
        r1 = -E * fy + A @ fx - ry
        r2 = A.T @ fy + D * fx - rx
        scale = max(np.abs(ry).max(), np.abs(rx).max(), np.abs(fy).max(), np.abs(fx).max()) + 1
        assert max(np.abs(r1).max(), np.abs(r2).max()) <= 1e-8 * scale
        assert P._rel(fy, sy) < 1e-6 and P._rel(fx, sx) < 1e-6
        z = np.random.default_rng(R).standard_normal(lp.m + lp.n)
        assert P._rel(Kf.rawsolve(z), Ks.rawsolve(z)) < 1e-5
    finally:
        Ks.close()
        Kf.close()


def _small_lps(vbkkt, count):
    return [vbkkt.workloads.random_sparse_lp(seed=i, m=60, n=120, nnz_per_col=4) for i in range(count)]


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_batch_of_independent_lps(vbkkt, gpu_lib, oracle_lib, mode):
    """BASELINE config 4 in small: a batch solved by vbk_solve_batch with several solver streams in
    flight.  Strict mode: every LP bit-equal to the oracle's METHOD run; fast mode: north_star tolerances."""
    lps = _small_lps(vbkkt, 8)
    md = vbkkt.MODE_STRICT if mode == "strict" else vbkkt.MODE_FAST
    res = vbkkt.batch.solve_local(gpu_lib, lps, method="hsd", device=0, mode=md, nstreams=4)
    for lp, r in zip(lps, res):
        st, log, x, y = H.call_solver(oracle_lib.kko_solver_hsd, lp)
        assert r["status"] == st == 0
        if mode == "strict":
            assert np.array_equal(r["x"], x) and np.array_equal(r["y"], y)
            assert r["iterations"] == len(H.iteration_lines(log))
        else:
            obj = float(lp.c @ x)
            assert abs(float(lp.c @ r["x"]) - obj) <= 1e-8 * max(1.0, abs(obj))
            assert abs(r["iterations"] - len(H.iteration_lines(log))) <= 2


def test_batch_driver_single_rank_summary(vbkkt, gpu_lib):
    """solve_batch without a process group = one rank owning every LP; the summary has one row per LP."""
    summary, local = vbkkt.batch.solve_batch(gpu_lib, lambda i: _small_lps(vbkkt, i + 1)[i], 5, nstreams=2,
                                             mode=vbkkt.MODE_FAST)
    assert summary.shape == (5, 5) and not np.isnan(summary).any() and len(local) == 5
    assert (summary[:, 0] == 0).all() and (summary[:, 1] > 5).all()
    assert np.allclose(summary[:, 2], summary[:, 3], rtol=1e-6)          # primal = dual objective at the optimum


@pytest.mark.parametrize("name", ["afiro", "25fv47", "ken-07"])
def test_rowblock_ops_on_one_gpu(vbkkt, gpu_lib, oracle_lib, name):
    """BASELINE config 5, the per-rank kernels on a real GPU (world 1; the 2-rank partition is covered by
    tests/test_multi.py under gloo and by bench.py --workload rowblock --gpus N under NCCL): A x and A^T y
    bit-identical to the oracle's smx, dot products to rounding, max-norms exact."""
    import torch
    lp = H.load_fixture(name)
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    ops = vbkkt.rowblock.RowBlockOps(gpu_lib, lp.m, lp.n, lp.kA, lp.iA, lp.A, kAt, iAt, At, "cuda:0")
    rng = np.random.default_rng(11)
    x, y, w = rng.standard_normal(lp.n), rng.standard_normal(lp.m), rng.standard_normal(lp.m)
    rho_ref, sig_ref = np.zeros(lp.m), np.zeros(lp.n)
    oracle_lib.kko_smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(rho_ref))
    oracle_lib.kko_smx(lp.n, lp.m, H.ptr_d(At), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(y), H.ptr_d(sig_ref))
    lx, ly, lw = ops.local_x(x), ops.local_y(y), ops.local_y(w)
    assert np.array_equal(ops.A_x(lx).cpu().numpy()[: lp.m], rho_ref)
    assert np.array_equal(ops.At_y(ly).cpu().numpy()[: lp.n], sig_ref)
    d = ops.dots([(lx, lx), (ly, lw)]).cpu().numpy()
    assert abs(d[0] - x @ x) <= 1e-13 * (x @ x) and abs(d[1] - y @ w) <= 1e-13 * np.abs(y * w).sum()
    mx = ops.absmax([lx, ly]).cpu().numpy()
    assert mx[0] == np.abs(x).max() and mx[1] == np.abs(y).max()
    torch.cuda.synchronize()


def test_coupled_partition_on_one_gpu(vbkkt, gpu_lib, oracle_lib):
    """rowblock.CoupledOps on the device (world size 1: every row is local): A x and A^T y equal the oracle's smx bit for
    bit, the six scalars of the step agree with the host's."""
    import torch
    lp = vbkkt.workloads.multicommodity_lp(6, 4)
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    ops = vbkkt.rowblock.CoupledOps(gpu_lib, lp.m, lp.n, lp.kA, lp.iA, lp.A, kAt, iAt, At, "cuda:0")
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(lp.n), rng.standard_normal(lp.m)
    rho = torch.zeros(ops.rows_per, dtype=torch.float64, device="cuda:0")
    sig = torch.zeros(ops.cols_per, dtype=torch.float64, device="cuda:0")
    sums, maxes = ops.step(ops.local_x(x), ops.local_y(y), rho, sig)
    torch.cuda.synchronize()
    rho_ref, sig_ref = np.zeros(lp.m), np.zeros(lp.n)
    oracle_lib.kko_smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(rho_ref))
    oracle_lib.kko_smx(lp.n, lp.m, H.ptr_d(At), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(y), H.ptr_d(sig_ref))
    got = rho.cpu().numpy()
    assert ops.ns == 0 and np.array_equal(got[: ops.nl], rho_ref[ops.local_rows])
    assert np.array_equal(sig.cpu().numpy()[: lp.n], sig_ref)
    ref = np.array([x @ sig_ref, y @ rho_ref, rho_ref @ rho_ref, sig_ref @ sig_ref])
    assert np.allclose(sums.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)
    assert maxes.cpu().numpy()[0] == np.abs(rho_ref).max() and maxes.cpu().numpy()[1] == np.abs(sig_ref).max()


def test_factor_residual_property(vbkkt, gpu_lib):
    """Size-independent property at a size the oracle is not needed for: K z = rhs after
    forwardbackward, measured with scipy (|r| small relative to |rhs|)."""
    import scipy.sparse as sp
    lp = H.load_fixture("pds-02")
    rng = np.random.default_rng(7)
    E = rng.uniform(0.1, 10.0, lp.m)
    D = rng.uniform(0.1, 10.0, lp.n)
    ry, rx = rng.standard_normal(lp.m), rng.standard_normal(lp.n)
    K = H.kkt_for(vbkkt, gpu_lib, lp)
    K.factor(E, D)
    sy, sx, ok = K.solve(E, D, ry, rx)
    A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
    # [-E A; A^T D] [sy; sx] = [ry; rx]   
    r1 = -E * sy + A @ sx - ry
    r2 = A.T @ sy + D * sx - rx
    scale = max(np.abs(ry).max(), np.abs(rx).max()) + 1
    assert max(np.abs(r1).max(), np.abs(r2).max()) <= 1e-8 * scale
    K.close()


@pytest.mark.parametrize("name", ["afiro", "25fv47"])
def test_b1_speculative_second_rhs(vbkkt, gpu_lib, oracle_lib, name):
    """The B1 symbols answer the second forwardbackward call of a factorisation from a speculative solve (csrc/vbk_capi.cu:
    b1_solve): every call must return the oracle's bits, also when the speculation misses."""
    P.check_b1_speculation(gpu_lib, oracle_lib, H.load_fixture(name))


def test_b1_seam_reference_method_on_gpu_plugins(vbkkt, gpu_lib):
    """The drop-in itself: the reference's UNMODIFIED METHOD object (hsd.c compiled in oracle/_ref)
    linked against libvbkkt.so's ldltfac/forwardbackward/smx/atnum/dotprod/maxv reproduces the
    golden log.  One LP per process (the reference's plugin state is process-global)."""
    lib = H.REF_DIR / "libhsd_b1.so"
    if not lib.exists():
        pytest.skip("oracle/_ref seam objects not built")
    code = (
        "import sys; sys.path.insert(0, r'%s'); import ctypes as C, numpy as np, harness as H\n"
        "lp = H.load_fixture(sys.argv[1]); lib = C.CDLL(r'%s')\n"
        "st, log, x, y = H.call_solver(lib.solver, lp)\n"
        "ok = st == int(lp.extra['hsd_status']) and log == str(lp.extra['hsd_log']) and "
        "np.array_equal(x, lp.extra['hsd_x']) and np.array_equal(y, lp.extra['hsd_y'])\n"
        "sys.exit(0 if ok else 3)\n" % (H.ROOT / "tests", lib))
    for name in ("afiro", "israel", "25fv47"):
        r = subprocess.run([os.sys.executable, "-c", code, name], capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stdout[-400:], r.stderr[-400:])


def test_b2_seam_executable_matches_reference_executable(tmp_path):
    """Whole-program drop-in: reference driver + MPS reader + solvelp linked with the product's
    `solver` (ipo_hsd_b2) prints exactly what the all-reference executable prints, on an MPS file
    written by this test (the netlib MPS files do not travel to the GPU box)."""
    b2, ref = H.REF_DIR / "ipo_hsd_b2", H.REF_DIR / "ipo_hsd_ref"
    if not (b2.exists() and ref.exists()):
        pytest.skip("oracle/_ref executables not built")
    rng = np.random.default_rng(3)
    m, n = 12, 20
    lines = ["NAME          TESTLP", "ROWS", " N  COST"]
    lines += [f" L  R{i:03d}" for i in range(m)]
    lines.append("COLUMNS")
    rowsum = np.zeros(m)
    for j in range(n):
        lines.append(f"    X{j:03d}      COST      {float(rng.integers(1, 9)):12.1f}")
        for i in sorted(rng.choice(m, size=3, replace=False)):
            v = float(rng.integers(1, 6))
            rowsum[i] += v
            lines.append(f"    X{j:03d}      R{i:03d}      {v:12.1f}")
    lines.append("RHS")
    for i in range(m):
        lines.append(f"    RHS       R{i:03d}      {rowsum[i] + 3.0:12.1f}")
    lines.append("ENDATA")
    mps = tmp_path / "testlp.mps"
    mps.write_text("\n".join(lines) + "\n")
    outs = []
    for exe in (ref, b2):
        r = subprocess.run([str(exe), str(mps)], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-400:]
        outs.append([l for l in r.stdout.splitlines() if "Version" not in l])
    assert outs[0] == outs[1]
    assert any("optimal solution" in l for l in outs[1])
