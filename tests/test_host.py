"""CPU tests of the product's host side: C-ABI surface, symbolic analysis (bit-exact against the
reference's perm/kAAt/iAAt for every fixture), derived structures, loud failure without a GPU."""
import ctypes as C
import hashlib
import re
import subprocess
import sys

import numpy as np
import pytest

import harness as H


def test_library_exports_every_declared_symbol(product_lib):
    header = (H.ROOT / "include" / "vbkkt.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", header))
    names -= {"defined", "C"}
    decl = {n for n in names if re.search(r"\b(void|int|double|long|char|vbk_kkt)\b[\s\*]+" + n + r"\s*\(", header)}
    assert {"ldltfac", "forwardbackward", "smx", "atnum", "dotprod", "maxv", "inv_clo", "inv_num", "solve",
            "vbk_solver_hsd", "vbk_solver_hsdls", "vbk_solver_intpt", "vbk_kkt_create", "vbk_solve_lp"} <= decl
    for n in sorted(decl):
        assert hasattr(product_lib, n), f"libvbkkt.so does not export {n}"


def test_solver_shims_export_solver(product_lib, vbkkt):
    for meth in ("hsd", "hsdls", "intpt"):
        lib = C.CDLL(str(vbkkt.PKG_DIR / f"libvbkkt_{meth}.so"))
        assert hasattr(lib, "solver")


def test_product_does_not_link_oracle(vbkkt):
    out = subprocess.run(["ldd", str(vbkkt.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out and "emu" not in out
    syms = subprocess.run(["nm", "-D", str(vbkkt.LIB_PATH)], capture_output=True, text=True).stdout
    assert "kko_" not in syms


@pytest.mark.parametrize("name", H.fixture_names())
def test_symbolic_is_bit_exact(vbkkt, product_lib, name):
    """perm, iperm, kAAt, iAAt, denwin, pdf of the product's host analysis == the reference's
    (reference src/ipo/ldlt.c:638-1262), for every committed netlib fixture."""
    lp = H.load_fixture(name)
    k = H.kkt_for(vbkkt, product_lib, lp, device=-1)
    assert k.dim == lp.m + lp.n
    perm = k.perm
    assert np.array_equal(perm, lp.extra["sym_perm"])
    assert np.array_equal(k.iperm[perm], np.arange(k.dim))
    assert np.array_equal(k.kAAt, lp.extra["sym_kAAt"])
    iL = k.iAAt.astype(np.int32)
    assert hashlib.sha256(iL.tobytes()).hexdigest() == str(lp.extra["sym_iAAt_sha256"])
    if "sym_iAAt" in lp.extra:
        assert np.array_equal(iL, lp.extra["sym_iAAt"])
    assert k.denwin == int(lp.extra["sym_denwin"])
    assert k.pdf == int(lp.extra["sym_pdf"])
    assert k.lnz == int(lp.extra["sym_lnz"])
    assert k.narth == float(lp.extra["sym_narth"])
    k.close()


def test_pattern_rebuilt_from_cached_ordering_on_every_fixture(vbkkt, product_lib, tmp_path, monkeypatch):
    """Symbolic::pattern_from_ordering (the path a cached ordering takes, vbk_symbolic.cpp): for every committed netlib
    fixture the second analysis -- ordering read from $VBK_SYM_CACHE, fill pattern rebuilt by the elimination-tree pass --
    gives the reference's kAAt and iAAt (sha256 stored with the fixture)."""
    monkeypatch.setenv("VBK_SYM_CACHE", str(tmp_path))
    for name in H.fixture_names():
        lp = H.load_fixture(name)
        for _ in range(2):
            k = H.kkt_for(vbkkt, product_lib, lp, device=-1)
            kA, iL = k.kAAt, k.iAAt.astype(np.int32)
            k.close()
        assert np.array_equal(kA, lp.extra["sym_kAAt"]), name
        assert hashlib.sha256(iL.tobytes()).hexdigest() == str(lp.extra["sym_iAAt_sha256"]), name
    assert len(list(tmp_path.glob("vbksym_*.bin"))) == len(H.fixture_names())


def test_numeric_call_without_gpu_fails_loudly(vbkkt, product_lib):
    """No CPU fallback: on a box without a CUDA device a numeric entry point terminates the process
    with a message (run in a child so the test process survives)."""
    if product_lib.vbk_device_count() > 0:
        pytest.skip("a CUDA device is present")
    code = ("import importlib.util,sys,numpy as np;"
            f"spec=importlib.util.spec_from_file_location('vbkkt', r'{vbkkt.PKG_DIR}/__init__.py');"
            "m=importlib.util.module_from_spec(spec);spec.loader.exec_module(m);"
            "print(m.dotprod(np.ones(4),np.ones(4)))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU path" in r.stderr


def test_multicommodity_bench_workload(vbkkt, product_lib):
    """BASELINE config 3 as bench.py builds it: the generator is deterministic, the synthetic iterate is
    reproducible, and the symbolic counts stored for the reference arm (bench.MCF_KNOWN) are what the
    bit-exact symbolic phase finds (host only, no GPU)."""
    import importlib.util
    root = vbkkt.PKG_DIR.parent
    spec = importlib.util.spec_from_file_location("bench_for_test", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    lp, it = bench.mcf_workload("mcf:4:3")
    lp2, it2 = bench.mcf_workload("mcf:4:3")
    V, E = 16, 48
    assert (lp.m, lp.n, lp.nz) == (2 * 3 * V + E, 3 * E, 5 * 3 * E)
    assert np.array_equal(lp.A, lp2.A) and np.array_equal(lp.iA, lp2.iA) and np.array_equal(it["E"], it2["E"])
    assert it["E"].shape == (lp.m,) and it["D"].shape == (lp.n,) and (it["E"] > 0).all() and (it["D"] > 0).all()
    lp, _ = bench.mcf_workload("mcf:20:12")
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    k = vbkkt.KKT(device=-1, lib=product_lib)
    k.analyze(lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A)
    narth, lnz = bench.MCF_KNOWN[(20, 12)]
    assert k.lnz == lnz and abs(k.narth - narth) <= 1e-3 * narth
    k.close()


def test_symbolic_disk_cache_round_trip(vbkkt, emu_lib, tmp_path, monkeypatch):
    """$VBK_SYM_CACHE: the second analysis of a pattern reads the ORDERING back from disk, rebuilds the fill pattern from
    it (elimination-tree pass) and yields the same arrays as the full analysis -- which in turn equal the
    compiled reference's (fixture); a different pattern gets its own file.  Problems with and without a dense window."""
    monkeypatch.setenv("VBK_SYM_CACHE", str(tmp_path))
    names = ("afiro", "sc50b", "israel", "25fv47", "ken-07", "pilotnov")
    for name in names:
        lp = H.load_fixture(name)
        kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
        outs = []
        for _ in range(2):
            K = vbkkt.KKT(device=-1, lib=emu_lib)
            K.analyze(lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A)
            outs.append((K.perm, K.iperm, K.kAAt, K.iAAt, K.denwin, K.narth))
            K.close()
        for a, b in zip(outs[0], outs[1]):
            assert np.array_equal(a, b), name
        assert np.array_equal(outs[1][0], lp.extra["sym_perm"]) and np.array_equal(outs[1][2], lp.extra["sym_kAAt"])
        assert np.array_equal(outs[1][3], lp.extra["sym_iAAt"])
    files = list(tmp_path.glob("vbksym_*.bin"))
    assert len(files) == len(names)
    sizes = sorted(f.stat().st_size for f in files)
    want = sorted(72 + 8 * (H.load_fixture(nm).m + H.load_fixture(nm).n) for nm in names)   # ordering only: 8 bytes per row/column of K
    assert sizes == want
