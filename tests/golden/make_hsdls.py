#!/usr/bin/env python
"""Golden fixtures for METHOD = hsdls (reference src/ipo/hsdls.c; the reference ships no logs for it): the UNMODIFIED
reference compiled in place (oracle/_ref/libref_hsdls.so, `make -C oracle ref`) is run on the solver-space LP arrays of
the committed netlib fixtures; its stdout, status and final x, y are stored under tests/golden/hsdls/.
Each LP runs in its own process: the reference keeps one factor object per process (ldlt.c:108-120).
Run in the BUILD container only.     python tests/golden/make_hsdls.py [names...]"""
import subprocess
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import harness as H  # noqa: E402

NAMES = ["afiro", "adlittle", "blend", "sc50a", "sc50b", "sc105", "share2b", "kb2", "israel", "stocfor1", "scsd1", "e226",
         "bandm", "sctap1", "25fv47"]
OUT = H.GOLDEN / "hsdls"

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        name = sys.argv[2]
        lp = H.load_fixture(name)
        ref = H.load_ref("hsdls")
        st, log, x, y = H.call_solver(ref.solver, lp)
        np.savez_compressed(OUT / f"{name}.npz", status=st, log=log, x=x, y=y)
        print(name, "status", st, "lines", len(H.iteration_lines(log)))
    else:
        OUT.mkdir(parents=True, exist_ok=True)
        for name in (sys.argv[1:] or NAMES):
            subprocess.run([sys.executable, __file__, "--one", name], check=True)
