"""CPU tests of the KERNEL LOGIC: the product's .cu sources compiled as C++ on the host thread
emulator (tests/emu/cuda_emu.h; every CUDA thread an OS thread) and checked bit-for-bit against the
oracle on tiny inputs.  This is test infrastructure, not a fallback -- libvbkkt.so never contains
it.  The real-GPU versions of the same checks are in tests/test_gpu.py."""
import pytest

import harness as H
import parity as P


def test_emu_build_is_labelled(emu_lib):
    assert b"TEST build" in emu_lib.vbk_version()


def test_emu_linalg(vbkkt, emu_lib, oracle_lib):
    P.check_linalg(vbkkt, emu_lib, oracle_lib, sizes=(0, 1, 31, 33, 67, 2047, 2049, 4500))


def test_emu_transpose_and_smx(vbkkt, emu_lib, oracle_lib):
    P.check_transpose_and_smx(vbkkt, emu_lib, oracle_lib, H.load_fixture("afiro"))
    P.check_ragged_transpose(vbkkt, emu_lib)


@pytest.mark.parametrize("method,it", [("hsd", 3), ("hsd", 26), ("intpt", 14)])
def test_emu_kkt_step_bit_exact(vbkkt, emu_lib, oracle_lib, method, it):
    """hsd iteration 26 has exact-zero pivots (ndep>0, ldlt.c:600-614); intpt iteration 14 needs a
    second refinement pass (ldlt.c:411)."""
    info = P.check_kkt_step(vbkkt, emu_lib, oracle_lib, H.load_fixture("afiro"), method, it)
    if (method, it) == ("hsd", 26):
        assert info["ndep"] > 0
    if (method, it) == ("intpt", 14):
        assert info["passes"] == 2


@pytest.mark.parametrize("cap,blk", [(4, 4), (2, 3)])
def test_emu_sliced_columns_bit_exact(vbkkt, emu_lib, oracle_lib, monkeypatch, cap, blk):
    """Force the long-column path (several CTAs share one column, pivot hand-off between slices,
    dependent-pivot rule across slices) on a tiny LP by lowering the slicing thresholds."""
    monkeypatch.setenv("VBK_WHOLE_CAP", str(cap))
    monkeypatch.setenv("VBK_ROWBLK", str(blk))
    info = P.check_kkt_step(vbkkt, emu_lib, oracle_lib, H.load_fixture("afiro"), "hsd", 26)
    assert info["ndep"] > 0


@pytest.mark.parametrize("name,it,env", [("afiro", 5, {}), ("sc50b", 12, {}),
                                         ("sc105", 15, {"VBK_WINDOW_RHO": "0.1"}),
                                         ("israel", 12, {"VBK_WINDOW_RHO": "0.1"}),
                                         ("israel", 12, {"VBK_WINDOW_RHO": "0.1", "VBK_SPARSE": "strict"})])
def test_emu_fast_mode_kkt_step(vbkkt, emu_lib, oracle_lib, monkeypatch, name, it, env):
    """Fast mode: sparse part + Schur assembly + blocked dense LDL^T + dense-window sweeps, several
    panels and a padded (rho < 1) window forced on tiny LPs.  israel at rho = 0.1 has a 131-wide window:
    three of the emulated build's 64-column panels (diagonal block with two sub-blocks, rows below,
    trailing update) and two of the 128-row panels of the window sweeps (k_window_tri3 with its four column
    slices and the inverted diagonal blocks of k_window_tinv); VBK_SPARSE=strict runs the sparse columns
    through the strict slice-task kernel instead of the level kernels."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    P.check_kkt_step_fast(vbkkt, emu_lib, oracle_lib, H.load_fixture(name), "hsd", it)


@pytest.mark.parametrize("name,it,env", [("afiro", 26, {"VBK_SPARSE": "level"}),
                                         ("afiro", 26, {"VBK_SPARSE": "level", "VBK_SPARSE_HEAVY": "0"}),
                                         ("sc50b", 12, {"VBK_SPARSE": "level", "VBK_SPARSE_HEAVY": "8"}),
                                         ("israel", 12, {"VBK_SPARSE": "level", "VBK_WINDOW_RHO": "0.5"}),
                                         ("afiro", 26, {})])
def test_emu_fast_mode_sparse_columns_bit_exact(vbkkt, emu_lib, oracle_lib, monkeypatch, name, it, env):
    """The level-scheduled sparse-column kernels of fast mode (warp per light column, CTA per heavy column;
    VBK_SPARSE_HEAVY moves the boundary) reproduce the reference's sparse columns bit for bit, dependent pivots
    included (afiro iterate 26); without VBK_SPARSE the handle measures both paths and keeps one."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    info = P.check_fast_sparse_columns(vbkkt, emu_lib, oracle_lib, H.load_fixture(name), "hsd", it)
    assert info["T"] > 0 and info["nsp"] > 0


def test_emu_fast_mode_full_solve(vbkkt, emu_lib):
    assert P.check_full_solve_fast(vbkkt, emu_lib, H.load_fixture("afiro")) == 0


def test_emu_kkt_step_second_problem(vbkkt, emu_lib, oracle_lib):
    P.check_kkt_step(vbkkt, emu_lib, oracle_lib, H.load_fixture("sc50b"), "hsd", 10)


def test_emu_full_solve_intpt_afiro(vbkkt, emu_lib):
    """BASELINE.json config 1 (intpt on afiro) end to end through the emulated kernels:
    25 log lines, status 0, byte-identical log, bit-equal x and y."""
    lp = H.load_fixture("afiro")
    assert P.check_full_solve(vbkkt, emu_lib, lp, "intpt") == 0
    assert len(H.iteration_lines(str(lp.extra["intpt_log"]))) == 25


@pytest.mark.parametrize("name", ["afiro"])
def test_emu_full_solve_hsdls(vbkkt, emu_lib, name):
    """the long-step METHOD (hsdls.c) end to end through the emulated kernels -- per-component line
    search (hsdls.c:296-336) with the reference's MIN-fold semantics, constant delta, its own status rules -- prints
    the compiled reference's log byte for byte and returns bit-equal x and y."""
    assert P.check_hsdls(vbkkt, emu_lib, H.load_fixture(name)) == 0


def test_emu_b1_speculative_second_rhs(vbkkt, emu_lib, oracle_lib):
    """Seam B1 (ldltfac + two forwardbackward calls per factorisation, hsd.c:218-228): from the second factorisation on the
    library solves the remembered second right-hand side together with the first call and answers the second call from
    that result -- which must be exactly what the oracle computes for each call, also when the second right-hand side
    turns out to differ from the remembered one."""
    P.check_b1_speculation(emu_lib, oracle_lib, H.load_fixture("afiro"))
