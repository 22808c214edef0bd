"""CPU tests that PIN THE ORACLE: the plain-C restatement (oracle/kkt_oracle.c) must reproduce
(a) the reference's own golden logs (evaluate/v1-cf4d5ba/netlib/ipo/<name>.mps.sol, committed as
fixtures by tests/golden/make_golden.py together with the reference's x,y and symbolic arrays) and
(b) the compiled, unmodified reference in oracle/_ref where that is present."""
import hashlib

import numpy as np
import pytest

import harness as H

# fixtures the single-thread oracle finishes in well under a second each
SMALL = ["afiro", "adlittle", "blend", "kb2", "sc50a", "sc50b", "sc105", "sc205", "share2b", "stocfor1",
         "israel", "recipe", "scagr7", "boeing2", "lotfi", "e226", "brandy", "share1b", "beaconfd",
         "agg", "sctap1", "scorpion", "bandm", "scfxm1", "forplan", "25fv47"]


def test_fixtures_are_the_reference_goldens():
    """Every committed fixture was checked line-for-line against the reference's golden log."""
    names = H.fixture_names()
    assert len(names) >= 80
    for name in names:
        lp = H.load_fixture(name)
        assert bool(lp.extra["golden_match"]), name


@pytest.mark.parametrize("name", SMALL)
def test_restatement_reproduces_golden_hsd_log(oracle_lib, name):
    lp = H.load_fixture(name)
    status, log, x, y = H.call_solver(oracle_lib.kko_solver_hsd, lp)
    assert status == int(lp.extra["hsd_status"])
    assert log == str(lp.extra["hsd_log"])          # byte-for-byte, incl. every printed digit
    assert np.array_equal(x, lp.extra["hsd_x"])     # bit-equal solution
    assert np.array_equal(y, lp.extra["hsd_y"])


@pytest.mark.parametrize("name", ["afiro", "adlittle", "blend", "sc50a", "share2b", "israel", "25fv47"])
def test_restatement_reproduces_reference_intpt(oracle_lib, name):
    """No golden logs exist for intpt: the fixture holds the compiled reference's run."""
    lp = H.load_fixture(name)
    status, log, x, y = H.call_solver(oracle_lib.kko_solver_intpt, lp)
    assert status == int(lp.extra["intpt_status"])
    assert log == str(lp.extra["intpt_log"])
    assert np.array_equal(x, lp.extra["intpt_x"])
    assert np.array_equal(y, lp.extra["intpt_y"])


@pytest.mark.parametrize("name", ["afiro", "25fv47", "pilot87"])
def test_known_answer_symbolic_facts(name):
    """SURVEY 8(c) known answers: Lnz, arith_ops, dense window."""
    facts = {"afiro": (67, 230, 1745.0, 8), "25fv47": (2908, 101278, 11742966.0, 239),
             "pilot87": (8944, 582242, 222143962.0, 695)}
    lp = H.load_fixture(name)
    N, lnz, narth, window = facts[name]
    assert lp.m + lp.n == N
    assert int(lp.extra["sym_lnz"]) == lnz
    assert float(lp.extra["sym_narth"]) == narth
    assert N - int(lp.extra["sym_denwin"]) == window


@pytest.mark.parametrize("name", ["afiro", "blend", "israel", "25fv47"])
def test_restatement_symbolic_matches_reference(oracle_lib, name):
    lp = H.load_fixture(name)
    F = H.oracle_factor_for(oracle_lib, lp)
    F.factor(np.ones(lp.m), np.ones(lp.n))
    assert np.array_equal(F.perm, lp.extra["sym_perm"])
    assert np.array_equal(F.kAAt, lp.extra["sym_kAAt"])
    assert hashlib.sha256(F.iAAt.astype(np.int32).tobytes()).hexdigest() == str(lp.extra["sym_iAAt_sha256"])
    assert int(oracle_lib.kko_denwin(F.h)) == int(lp.extra["sym_denwin"])
    assert int(oracle_lib.kko_pdf(F.h)) == int(lp.extra["sym_pdf"])
    F.close()


def test_restatement_against_live_reference(oracle_lib):
    """Where oracle/_ref exists (build container and, prebuilt, the GPU box): run the unmodified
    reference in-process and compare logs, solution and the numeric factor of the last iteration."""
    ref = H.load_ref("hsd")
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for name in ("afiro", "sc105", "israel"):
        lp = H.load_fixture(name)
        s1, log1, x1, y1 = H.call_solver(ref.solver, lp)
        s2, log2, x2, y2 = H.call_solver(oracle_lib.kko_solver_hsd, lp)
        assert (s1, log1) == (s2, log2)
        assert np.array_equal(x1, x2) and np.array_equal(y1, y2)
        ref.ref_ldlt_reset()


def test_linalg_restatement_edge_cases(oracle_lib):
    z = np.zeros(0)
    assert oracle_lib.kko_dotprod(H.ptr_d(z), H.ptr_d(z), 0) == 0.0      # empty input
    assert oracle_lib.kko_maxv(H.ptr_d(z), 0) == 0.0
    x = np.array([1.0, -3.0, 2.0])
    assert oracle_lib.kko_maxv(H.ptr_d(x), 3) == 3.0
    # ragged CSC with an empty column and an empty row
    kA = np.array([0, 2, 2, 3], dtype=np.int32); iA = np.array([0, 3, 1], dtype=np.int32)
    A = np.array([1.0, 2.0, 3.0])
    kAt = np.zeros(5, dtype=np.int32); iAt = np.zeros(3, dtype=np.int32); At = np.zeros(3)
    oracle_lib.kko_atnum(4, 3, H.ptr_i(kA), H.ptr_i(iA), H.ptr_d(A), H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(At))
    k2, i2, a2 = H.transpose_csc(4, 3, kA, iA, A)
    assert np.array_equal(kAt, k2) and np.array_equal(iAt, i2) and np.array_equal(At, a2)
