"""Build recipes for the native libraries (run by __graft_entry__.build()).

  libvbkkt.so         the product: nvcc, sm_100a only, -fmad=false (strict arithmetic), in-tree
  libvbkkt_hsd.so     one-symbol shims exporting the reference's METHOD entry point `solver`
  libvbkkt_intpt.so
  tests/emu/libvbkkt_emu.so   TEST build of the same sources on the host thread emulator
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SOURCES = ["vbk_symbolic.cpp", "vbk_kkt.cu", "vbk_kkt_fast.cu", "vbk_linalg.cu", "vbk_solver.cu", "vbk_batch.cu", "vbk_rowblock.cu", "vbk_capi.cu"]
HEADERS = sorted(p.name for p in CSRC.glob("*.h")) + sorted(p.name for p in CSRC.glob("*.cuh"))
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",                       # strict mode: no FMA contraction on the device
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default",
    "-Xlinker", "-Bsymbolic",            # our own smx/dotprod/... always bind inside the library
    "-shared", "-cudart", "static",
]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(p).exists():
        raise RuntimeError("nvcc not found")
    return p


def run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def build_product(force=False, verbose_ptxas=False) -> Path:
    out = PKG / "libvbkkt.so"
    deps = [CSRC / s for s in SOURCES + HEADERS] + [ROOT / "include" / "vbkkt.h"]
    if force or _newer(out, deps):
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", CSRC, "-o", out, *[CSRC / s for s in SOURCES]]
        if verbose_ptxas:
            cmd[1:1] = ["-Xptxas", "-v"]
        run(cmd)
    for meth in ("hsd", "hsdls", "intpt"):
        shim = PKG / f"libvbkkt_{meth}.so"
        src = CSRC / f"shim_solver_{meth}.c"
        if force or _newer(shim, [src, out]):
            run(["gcc", "-O2", "-fPIC", "-shared", "-o", shim, src, f"-L{PKG}", "-lvbkkt",
                 "-Wl,-rpath,$ORIGIN"])
    return out


def build_emu(force=False) -> Path:
    emu_dir = ROOT / "tests" / "emu"
    out = emu_dir / "libvbkkt_emu.so"
    deps = [CSRC / s for s in SOURCES + HEADERS] + [emu_dir / "cuda_emu.h", ROOT / "include" / "vbkkt.h"]
    if force or _newer(out, deps):
        cmd = ["g++", "-std=c++20", "-O1", "-g", "-DVBK_EMU", "-ffp-contract=off", "-fPIC", "-shared",
               "-Wl,-Bsymbolic", "-I", emu_dir, "-I", CSRC, "-o", out]
        for s in SOURCES:
            cmd += ["-x", "c++", CSRC / s]
        cmd += ["-lpthread"]
        run(cmd)
    return out


def build_oracle() -> None:
    run(["make", "-C", ROOT / "oracle", "restatement"])
    if Path("/root/reference/src").exists():
        run(["make", "-C", ROOT / "oracle", "ref"])


if __name__ == "__main__":
    what = sys.argv[1:] or ["product", "emu", "oracle"]
    if "product" in what:
        build_product(force="--force" in what, verbose_ptxas="-v" in what)
    if "emu" in what:
        build_emu(force="--force" in what)
    if "oracle" in what:
        build_oracle()
