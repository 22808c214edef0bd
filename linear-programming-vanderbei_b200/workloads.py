"""Synthetic LP generators for the configurations BASELINE.json names beyond netlib.

Everything is produced directly in *solver space* -- the form ``max c^T x, A x <= b, x >= 0`` that the
reference's ``solvelp`` (src/common/solve.c:101-205) hands to the METHOD plugin ``solver`` -- as CSC
arrays ``(kA, iA, A)`` with sorted row indices, so that a generated LP can be passed to
``vbkkt.solve_lp`` / ``vbk_solve_batch`` and to the reference's ``solver`` alike.

* ``random_sparse_lp``   config 4: one member of the batch of independent random sparse LPs.
* ``multicommodity_lp``  config 3: multicommodity flow on a planar R x R grid; equality rows are
  emitted as the +/- pair of inequalities ``solvelp`` would make of them (solve.c:127-147).

The reference has no generator ("Generator (new; none in reference)"); determinism comes
from ``numpy.random.default_rng(seed)``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SolverLP:
    """An LP in solver space: maximise c^T x + f subject to A x <= b, x >= 0 (A is m x n, CSC)."""
    name: str
    m: int
    n: int
    nz: int
    kA: np.ndarray
    iA: np.ndarray
    A: np.ndarray
    b: np.ndarray
    c: np.ndarray
    f: float = 0.0


def _csc_from_triplets(m, n, rows, cols, vals):
    """CSC with ascending row indices inside every column (duplicates are not expected)."""
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    kA = np.zeros(n + 1, dtype=np.int32)
    np.add.at(kA, cols + 1, 1)
    kA = np.cumsum(kA, dtype=np.int64).astype(np.int32)
    return kA, rows.astype(np.int32), vals.astype(np.float64)


def random_sparse_lp(seed: int, m: int = 2000, n: int = 4000, nnz_per_col: int = 8) -> SolverLP:
    """m "<=" rows, n columns, ``nnz_per_col`` nonzeros per column at uniform
    random distinct rows, values U[-1,1]; b = A 1 + U[0.1,1] (x = 1 is strictly feasible) and
    c = A^T 1 - U[0.1,1] (y = 1 is strictly dual feasible), so a finite optimum exists."""
    rng = np.random.default_rng(seed)
    k = min(nnz_per_col, m)
    # distinct rows per column: argpartition of random keys (vectorised "sample without replacement")
    keys = rng.random((n, m))
    rows = np.argpartition(keys, k - 1, axis=1)[:, :k].astype(np.int64).ravel()
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    vals = rng.uniform(-1.0, 1.0, size=n * k)
    vals[vals == 0.0] = 0.5
    kA, iA, A = _csc_from_triplets(m, n, rows, cols, vals)
    Ax1 = np.zeros(m)
    np.add.at(Ax1, iA, A)
    Aty1 = np.add.reduceat(A, kA[:-1]) if n else np.zeros(0)
    b = Ax1 + rng.uniform(0.1, 1.0, size=m)
    c = Aty1 - rng.uniform(0.1, 1.0, size=n)
    return SolverLP(f"rand{m}x{n}s{seed}", m, n, int(kA[n]), kA, iA, A, b, c, 0.0)


def multicommodity_lp(R: int, K: int, seed: int = 1) -> SolverLP:
    """K commodities on a planar R x R grid with 4-neighbour bidirectional arcs
    (V = R^2 nodes, E = 4 R (R-1) arcs).  Variables x[k,e] >= 0 (n = K E).  Rows, already in the form
    ``solvelp`` produces: for every commodity and node the conservation equality as a +/- pair of "<="
    rows (2 K V rows), then E joint-capacity rows sum_k x[k,e] <= cap_e.  Costs U[1,10) (negated: the
    solver maximises), one random s-t demand U[1,5) per commodity, cap_e = U[0.5,1] * sum_k d_k."""
    rng = np.random.default_rng(seed)
    V = R * R
    node = lambda r, c: r * R + c
    tail, head = [], []
    for r in range(R):
        for c in range(R):
            if c + 1 < R:
                tail += [node(r, c), node(r, c + 1)]; head += [node(r, c + 1), node(r, c)]
            if r + 1 < R:
                tail += [node(r, c), node(r + 1, c)]; head += [node(r + 1, c), node(r, c)]
    tail, head = np.asarray(tail, dtype=np.int64), np.asarray(head, dtype=np.int64)
    E = len(tail)
    assert E == 4 * R * (R - 1)
    n = K * E
    m = 2 * K * V + E
    cost = rng.uniform(1.0, 10.0, size=n)
    dem = rng.uniform(1.0, 5.0, size=K)
    src = rng.integers(0, V, size=K)
    dst = (src + 1 + rng.integers(0, V - 1, size=K)) % V
    cap = rng.uniform(0.5, 1.0, size=E) * dem.sum()

    # column (k,e): +1 at the head's conservation row, -1 at the tail's (inflow - outflow = supply),
    # mirrored with opposite signs in the twin row block, +1 in capacity row e
    kk = np.repeat(np.arange(K, dtype=np.int64), E)
    ee = np.tile(np.arange(E, dtype=np.int64), K)
    col = kk * E + ee
    r_head = kk * V + head[ee]
    r_tail = kk * V + tail[ee]
    rows = np.concatenate([r_head, r_tail, K * V + r_head, K * V + r_tail, 2 * K * V + ee])
    cols = np.concatenate([col, col, col, col, col])
    ones = np.ones(n)
    vals = np.concatenate([ones, -ones, -ones, ones, ones])
    kA, iA, A = _csc_from_triplets(m, n, rows, cols, vals)
    rhs = np.zeros(K * V)
    rhs[np.arange(K) * V + dst] += dem        # net inflow at the sink
    rhs[np.arange(K) * V + src] -= dem        # net outflow at the source
    b = np.concatenate([rhs, -rhs, cap])
    c = -cost                                 # minimise cost  ==  maximise -cost (solve.c:202-205)
    return SolverLP(f"mcf_R{R}_K{K}_s{seed}", m, n, int(kA[n]), kA, iA, A, b, c, 0.0)
