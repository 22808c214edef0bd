#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/netlib/.

Run in the BUILD container only (needs /root/reference and oracle/_ref built by
`make -C oracle ref`).  For every netlib MPS file present in the reference tree:

  1. oracle/_ref/mps_dump (reference MPS reader + solvelp transforms + a dumping METHOD plugin)
     gives the solver-space LP arrays -- the exact inputs of the hot path;
  2. the UNMODIFIED reference METHOD=hsd (oracle/_ref/libref_hsd.so) is run in-process on those
     arrays; its stdout must equal the reference's own golden log
     evaluate/v1-cf4d5ba/netlib/ipo/<name>.mps.sol line for line (checked here, recorded as
     `golden_match`);
  3. the symbolic arrays of the reference's factor object (perm, kAAt, sha256 of iAAt, denwin,
     pdf, Lnz) and the final x, y are stored beside the log;
  4. METHOD=intpt (no golden logs exist; oracle = compiled reference) is run for small problems.

Output: one compressed .npz per problem + index.json.  Nothing here is reference source code;
the numerical inputs are the public netlib LP data after the reference's own transforms.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import harness as H  # noqa: E402

REF = Path("/root/reference")
MPS_DIR = REF / "problems" / "netlib"
SOL_DIR = REF / "evaluate" / "v1-cf4d5ba" / "netlib" / "ipo"
OUT = H.GOLDEN / "netlib"


def golden_tail(name):
    """Golden log from the second 'm = ' line on (what solver() itself prints) + status line."""
    txt = (SOL_DIR / f"{name}.mps.sol").read_text(errors="replace").splitlines()
    idx = [i for i, l in enumerate(txt) if l.startswith("m = ")]
    return txt[idx[1]:] if len(idx) >= 2 else None


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def one(name, intpt_limit_nz=40000, max_nz=None):
    t0 = time.time()
    with tempfile.TemporaryDirectory() as td:
        dump = Path(td) / "lp.bin"
        env = dict(os.environ, VBK_DUMP=str(dump))
        r = subprocess.run([str(H.REF_DIR / "mps_dump"), str(MPS_DIR / f"{name}.mps")],
                           cwd=td, env=env, capture_output=True, text=True)
        if not dump.exists():
            # solvelp returned before calling solver (status 3: free variables, solve.c:79-87)
            return dict(name=name, skipped="no solver() call: " + r.stdout.strip().splitlines()[-1])
        lp = H.read_dump(dump, name)
    if max_nz is not None and lp.nz > max_nz:
        return dict(name=name, skipped=f"nz={lp.nz} above --max-nz")
    ref = H.load_ref("hsd")
    status, log, x, y = H.call_solver(ref.solver, lp)
    N = lp.m + lp.n
    for fn in ("ref_ldlt_perm", "ref_ldlt_iperm", "ref_ldlt_kAAt", "ref_ldlt_iAAt"):
        getattr(ref, fn).restype = H.c_int_p
    perm = np.ctypeslib.as_array(ref.ref_ldlt_perm(), (N,)).copy()
    kAAt = np.ctypeslib.as_array(ref.ref_ldlt_kAAt(), (N + 1,)).copy()
    lnz = int(kAAt[N])
    iAAt = np.ctypeslib.as_array(ref.ref_ldlt_iAAt(), (max(lnz, 1),))[:lnz].copy()
    sym = dict(denwin=int(ref.ref_ldlt_denwin()), pdf=int(ref.ref_ldlt_pdf()),
               dense=int(ref.ref_ldlt_dense()), lnz=lnz)
    narth = float(np.sum(np.diff(kAAt).astype(np.float64) ** 2) + 3.0 * lnz + N)
    gold = golden_tail(name)
    mine = log.splitlines()
    statmsg = {0: "optimal solution", 2: "primal infeasible", 4: "dual infeasible",
               5: "iteration limit", 1: "primal unbounded", 3: "dual unbounded",
               6: "infinite lower bounds - not implemented", 7: "suboptimal solution"}
    match = gold is not None and [l.rstrip() for l in gold] == [l.rstrip() for l in mine] + [statmsg[status]]
    rec = dict(dims=np.array([lp.m, lp.n, lp.nz], dtype=np.int32), f=np.float64(lp.f),
               kA=lp.kA, iA=lp.iA, A=lp.A, b=lp.b, c=lp.c,
               hsd_log=np.array(log), hsd_status=np.int32(status), hsd_x=x, hsd_y=y,
               sym_perm=perm.astype(np.int32), sym_kAAt=kAAt.astype(np.int32),
               sym_iAAt_sha256=np.array(sha(iAAt.astype(np.int32))),
               sym_denwin=np.int32(sym["denwin"]), sym_pdf=np.int32(sym["pdf"]),
               sym_lnz=np.int64(lnz), sym_narth=np.float64(narth),
               golden_match=np.bool_(match))
    if lnz <= 300000:
        rec["sym_iAAt"] = iAAt.astype(np.int32)
    info = dict(name=name, m=lp.m, n=lp.n, nz=lp.nz, N=N, lnz=lnz, narth=narth,
                denwin=sym["denwin"], pdf=sym["pdf"], hsd_status=status,
                hsd_lines=len(H.iteration_lines(log)), golden_match=bool(match))
    if lp.nz <= intpt_limit_nz:
        refi = H.load_ref("intpt")
        st2, log2, x2, y2 = H.call_solver(refi.solver, lp)
        rec.update(intpt_log=np.array(log2), intpt_status=np.int32(st2), intpt_x=x2, intpt_y=y2)
        info.update(intpt_status=st2, intpt_lines=len(H.iteration_lines(log2)))
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **rec)
    info["seconds"] = round(time.time() - t0, 2)
    info["bytes"] = (OUT / f"{name}.npz").stat().st_size
    return info


def _worker(args):
    name, kw = args
    try:
        return one(name, **kw)
    except Exception as e:  # keep the sweep going
        return dict(name=name, error=repr(e))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--jobs", type=int, default=7)
    ap.add_argument("--max-nz", type=int, default=None)
    a = ap.parse_args()
    names = a.names or sorted(p.stem for p in MPS_DIR.glob("*.mps"))
    # one fresh process per problem: the reference keeps process-global factor state
    with mp.get_context("spawn").Pool(a.jobs, maxtasksperchild=1) as pool:
        res = []
        for info in pool.imap_unordered(_worker, [(n, dict(max_nz=a.max_nz)) for n in names]):
            print(json.dumps(info), flush=True)
            res.append(info)
    idx_path = OUT / "index.json"
    old = {}
    if idx_path.exists():
        old = {r["name"]: r for r in json.loads(idx_path.read_text())}
    for r in res:
        old[r["name"]] = r
    idx_path.write_text(json.dumps(sorted(old.values(), key=lambda r: r["name"]), indent=1))


if __name__ == "__main__":
    main()
