#!/usr/bin/env python
"""Generate tests/golden/iterates/<name>_it<k>.npz: the inputs (E, D, rhs_y, rhs_x) and the
reference outputs (sol_y, sol_x, refinement passes, ndep) of the KKT step of hsd iteration k.

The producer is the oracle restatement (oracle/kkt_oracle.c), which tests/test_oracle.py pins
bit-for-bit to the compiled reference and to the reference's golden logs.  bench.py uses these as
its (identical) inputs for the GPU arm, the CPU baseline and the reference arm."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import harness as H  # noqa: E402


def main():
    lib = H.declare_oracle(C.CDLL(str(H.ORACLE_DIR / "libkkt_oracle.so")))
    lib.kko_set_itnlim.argtypes = [C.c_int]
    out = H.GOLDEN / "iterates"
    out.mkdir(parents=True, exist_ok=True)
    for spec in sys.argv[1:]:
        name, it = spec.split(":")
        it = int(it)
        lp = H.load_fixture(name)
        lib.kko_set_itnlim(it + 1)
        E, D, ry, rx, sy, sx = H.capture_step(lib, lp, "hsd", it)
        lib.kko_set_itnlim(0)
        np.savez_compressed(out / f"{name}_it{it}.npz", E=E, D=D, rhs_y=ry, rhs_x=rx, sol_y=sy, sol_x=sx,
                            iteration=np.int32(it))
        print(name, it, "saved", flush=True)


if __name__ == "__main__":
    main()
