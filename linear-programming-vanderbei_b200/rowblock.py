"""Row-block partitioned ``smx`` / ``dotprod`` / ``maxv`` over 2/4/8 GPUs (BASELINE.json config 5,
SURVEY.md 8e).

One process per GPU.  Vectors on the constraint side (length m: rho, w, y, dy, ...) are split into equal
contiguous row blocks, vectors on the variable side (length n: sigma, z, x, dx, ...) into equal contiguous
column blocks (the last block is zero-padded so that every rank holds ``ceil(len / world)`` entries and
the all-gather is a single contiguous collective).

* ``A_x``   (reference ``smx(m, n, A, kA, iA, x, rho)``, src/ipo/hsd.c:182): all-gather x (NCCL over NVLink),
  then the rank's row block of A -- a slice of the transpose arrays ``atnum`` built (hsd.c:111) -- times
  the full x with ``vbk_spmv_rows_dev``.  Every output entry is one row sum in ascending column order:
  bit-identical to the reference's scatter loop (linalg.c:62-70) for any number of ranks.
* ``At_y``  (``smx(n, m, At, kAt, iAt, y, sigma)``, hsd.c:191): same with the column block of A.
* ``dots`` / ``absmax`` (``dotprod`` linalg.c:17-25, ``maxv`` linalg.c:108-116): up to 8 local partial
  reductions in one launch, then ONE all-reduce of those few doubles (SUM, resp. MAX).

The device kernels are csrc/vbk_rowblock.cu; torch is only the carrier of device memory, the stream and
the collectives.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def declare(lib):
    lib.vbk_spmv_rows_dev.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vbk_spmv_rows_dev.restype = None
    lib.vbk_dots_partial_dev.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_longlong), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vbk_dots_partial_dev.restype = None
    lib.vbk_absmax_partial_dev.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.c_void_p, C.c_void_p]
    lib.vbk_absmax_partial_dev.restype = None
    lib.vbk_reduce_scratch_doubles.restype = C.c_int
    return lib


def block_bounds(length: int, rank: int, world: int):
    """(block size shared by all ranks, first index, one-past-last index) of rank's contiguous block."""
    per = (length + world - 1) // world
    lo = min(length, rank * per)
    return per, lo, min(length, lo + per)


class RowBlockOps:
    """The rank-local half of the partitioned linear algebra of one LP (solver-space A: m x n, CSC
    ``kA, iA, A`` with its transpose ``kAt, iAt, At`` as ``atnum`` produces it)."""

    def __init__(self, lib, m, n, kA, iA, A, kAt, iAt, At, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.lib, self.group = torch, dist, declare(lib), group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.device = torch.device(device)
        self.m, self.n = m, n
        self.rows_per, self.r0, self.r1 = block_bounds(m, self.rank, self.world)
        self.cols_per, self.c0, self.c1 = block_bounds(n, self.rank, self.world)
        kA, iA, A = np.asarray(kA), np.asarray(iA), np.asarray(A)
        kAt, iAt, At = np.asarray(kAt), np.asarray(iAt), np.asarray(At)

        def slab(ptr, idx, val, lo, hi):
            a, b = int(ptr[lo]), int(ptr[hi])
            t = lambda v, dt: torch.from_numpy(np.ascontiguousarray(v, dtype=dt)).to(self.device)
            return (t(ptr[lo:hi + 1] - ptr[lo], np.int32), t(idx[a:b] if b > a else np.zeros(1), np.int32),
                    t(val[a:b] if b > a else np.zeros(1), np.float64), b - a)
        # rows r0..r1 of A  = columns r0..r1 of the CSC transpose;  columns c0..c1 of A = rows of A^T
        self.rowblk = slab(kAt, iAt, At, self.r0, self.r1)
        self.colblk = slab(kA, iA, A, self.c0, self.c1)
        self.full_x = torch.zeros(self.cols_per * self.world, dtype=torch.float64, device=self.device)
        self.full_y = torch.zeros(self.rows_per * self.world, dtype=torch.float64, device=self.device)
        self.scratch = torch.zeros(int(self.lib.vbk_reduce_scratch_doubles()), dtype=torch.float64, device=self.device)
        self.red = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.red_max = torch.zeros(8, dtype=torch.float64, device=self.device)   # max-norms: own buffer (dots and absmax may be in flight together)
        self.red6 = torch.zeros(6, dtype=torch.float64, device=self.device)            # this rank's 4 sums + 2 maxima (step)
        self.red6_all = torch.zeros(6 * self.world, dtype=torch.float64, device=self.device)
        self.launches = 0

    # -- partition helpers -------------------------------------------------------------------------
    def local_x(self, full):
        """This rank's (zero-padded) block of a length-n host vector."""
        v = np.zeros(self.cols_per)
        v[: self.c1 - self.c0] = np.asarray(full)[self.c0:self.c1]
        return self.torch.from_numpy(v).to(self.device)

    def local_y(self, full):
        v = np.zeros(self.rows_per)
        v[: self.r1 - self.r0] = np.asarray(full)[self.r0:self.r1]
        return self.torch.from_numpy(v).to(self.device)

    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def _gather(self, full, local):
        if self.world > 1:
            self.dist.all_gather_into_tensor(full, local, group=self.group)
        else:
            full.copy_(local)
        return full

    # -- smx -----------------------------------------------------------------------------------------
    def A_x(self, x_local, out=None):
        """rho[rows of this rank] = (A x)[rows]; x_local is this rank's block of x."""
        out = out if out is not None else self.torch.zeros(self.rows_per, dtype=self.torch.float64, device=self.device)
        full = self._gather(self.full_x, x_local)
        ptr, idx, val, _ = self.rowblk
        self.lib.vbk_spmv_rows_dev(self.r1 - self.r0, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                   full.data_ptr(), out.data_ptr(), self._stream())
        self.launches += 1
        return out

    def At_y(self, y_local, out=None):
        """sigma[cols of this rank] = (A^T y)[cols]; y_local is this rank's block of y."""
        out = out if out is not None else self.torch.zeros(self.cols_per, dtype=self.torch.float64, device=self.device)
        full = self._gather(self.full_y, y_local)
        ptr, idx, val, _ = self.colblk
        self.lib.vbk_spmv_rows_dev(self.c1 - self.c0, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                   full.data_ptr(), out.data_ptr(), self._stream())
        self.launches += 1
        return out

    # -- dotprod / maxv --------------------------------------------------------------------------
    def _ptrs(self, vecs):
        arr = (C.c_void_p * len(vecs))(*[v.data_ptr() for v in vecs])
        lens = (C.c_longlong * len(vecs))(*[int(v.numel()) for v in vecs])
        return arr, lens

    def dots(self, pairs):
        """Global dot products of up to 8 pairs of partitioned vectors; returns a device tensor [len(pairs)]
        (every rank gets the same values)."""
        k = len(pairs)
        xs, lens = self._ptrs([p[0] for p in pairs])
        ys, _ = self._ptrs([p[1] for p in pairs])
        out = self.red[:k]
        self.lib.vbk_dots_partial_dev(k, xs, ys, lens, out.data_ptr(), self.scratch.data_ptr(), self._stream())
        self.launches += 2
        if self.world > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM, group=self.group)
        return out.clone()       # the reduction buffer is reused by the next call: hand out a copy

    def absmax(self, vecs):
        """Global max-norms (maxv) of up to 8 partitioned vectors; device tensor [len(vecs)]."""
        k = len(vecs)
        xs, lens = self._ptrs(vecs)
        out = self.red_max[:k]
        self.lib.vbk_absmax_partial_dev(k, xs, lens, out.data_ptr(), self._stream())
        self.launches += 1
        if self.world > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.MAX, group=self.group)
        return out.clone()       # the reduction buffer is reused by the next call: hand out a copy

    # -- one interior-point residual step, collectives overlapped -------------------------------------
    def step(self, x_local, y_local, rho, sig, overlap=True):
        """rho = (A x)[rows], sig = (A^T y)[cols] and the six scalars hsd.c:182-195 needs from them --
        x.sig, y.rho, rho.rho, sig.sig (sums) and max|rho|, max|sig| -- with as little exposed communication as the
        partition allows: both all-gathers are issued at once (NCCL's own stream) and each SpMV waits only for its own
        operand, and the six partial scalars of all ranks travel in ONE small all-gather; every rank then adds (resp.
        maximises) them in rank order, so all ranks hold bit-identical results.  Returns (sums[4], maxes[2])."""
        hy = None
        if self.world > 1 and overlap:
            hx = self.dist.all_gather_into_tensor(self.full_x, x_local, group=self.group, async_op=True)
            hy = self.dist.all_gather_into_tensor(self.full_y, y_local, group=self.group, async_op=True)
            hx.wait()
        elif self.world > 1:                 # stream-ordered (CUDA graph capture): the collectives sit on the current stream
            self.dist.all_gather_into_tensor(self.full_x, x_local, group=self.group)
            self.dist.all_gather_into_tensor(self.full_y, y_local, group=self.group)
        else:
            self.full_x.copy_(x_local); self.full_y.copy_(y_local)
        ptr, idx, val, _ = self.rowblk
        self.lib.vbk_spmv_rows_dev(self.r1 - self.r0, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                   self.full_x.data_ptr(), rho.data_ptr(), self._stream())
        if hy is not None:
            hy.wait()
        ptr, idx, val, _ = self.colblk
        self.lib.vbk_spmv_rows_dev(self.c1 - self.c0, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(),
                                   self.full_y.data_ptr(), sig.data_ptr(), self._stream())
        pairs = [(x_local, sig), (y_local, rho), (rho, rho), (sig, sig)]
        xs, lens = self._ptrs([p[0] for p in pairs])
        ys, _ = self._ptrs([p[1] for p in pairs])
        self.lib.vbk_dots_partial_dev(4, xs, ys, lens, self.red6.data_ptr(), self.scratch.data_ptr(), self._stream())
        vs, vlens = self._ptrs([rho, sig])
        self.lib.vbk_absmax_partial_dev(2, vs, vlens, self.red6[4:].data_ptr(), self._stream())
        self.launches += 5
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.red6_all, self.red6, group=self.group)
            allv = self.red6_all.view(self.world, 6)
            return allv[:, :4].sum(dim=0), allv[:, 4:].amax(dim=0)
        return self.red6[:4].clone(), self.red6[4:].clone()

    # -- algorithmic bytes of one A_x + At_y pair on this rank (SURVEY.md 8d work model) ----------------
    def spmv_bytes(self):
        nz_r, nz_c = self.rowblk[3], self.colblk[3]
        rows, cols = self.r1 - self.r0, self.c1 - self.c0
        return (12 * nz_r + 4 * (rows + 1) + 8 * self.n + 8 * rows) + (12 * nz_c + 4 * (cols + 1) + 8 * self.m + 8 * cols)


class CoupledOps:
    """The same step on a partition that follows the LP's block structure instead of cutting rows blindly (fast-mode
    variant of BASELINE config 5).  Columns are split into contiguous blocks as above.  A row whose entries all lie in
    ONE rank's columns is *local* to that rank; the rows that couple several column blocks (the joint-capacity rows of
    a multicommodity LP, a few per cent of m) are *shared*.

    * ``A x``: every rank multiplies its local rows and ITS columns' part of the shared rows with its own block of x --
      no x travels; the shared rows' partial sums are added with ONE all-reduce (ns doubles: 0.3 MB instead of the
      60 MB of x and y the row-block partition moves per step).  Local rows are bit-identical to the reference's
      ``smx``; shared rows are the same sums re-associated by column block (that is why this is not the strict variant).
    * ``A^T y``: a column only meets rows that are local to its rank or shared, and the shared part of y is replicated:
      no communication, entries summed in ascending row order -- bit-identical to the reference.
    * constraint-side vectors live as ``[local rows of the rank ; shared rows]``; the shared tail holds the same
      values on every rank and is counted once (rank 0) in the dot products.
    """

    def __init__(self, lib, m, n, kA, iA, A, kAt, iAt, At, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.lib, self.group = torch, dist, declare(lib), group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.device = torch.device(device)
        self.m, self.n = m, n
        self.cols_per, self.c0, self.c1 = block_bounds(n, self.rank, self.world)
        kA, iA, A = np.asarray(kA), np.asarray(iA), np.asarray(A, dtype=np.float64)
        kAt, iAt, At = np.asarray(kAt), np.asarray(iAt), np.asarray(At, dtype=np.float64)
        cnt = kAt[1:m + 1] - kAt[:m]
        nonempty = cnt > 0
        first = np.zeros(m, dtype=np.int64)
        last = np.zeros(m, dtype=np.int64)
        first[nonempty] = iAt[kAt[:m][nonempty]]                 # column indices ascend inside a row (atnum's order)
        last[nonempty] = iAt[kAt[1:m + 1][nonempty] - 1]
        own_lo, own_hi = first // self.cols_per, last // self.cols_per
        shared = own_lo != own_hi
        self.shared_rows = np.nonzero(shared)[0]
        self.local_rows = np.nonzero(~shared & (own_lo == self.rank))[0]
        self.nl, self.ns = len(self.local_rows), len(self.shared_rows)
        self.rows_per = self.nl + self.ns                        # length of this rank's constraint-side vectors
        pos = np.full(m, -1, dtype=np.int64)                     # row -> position in [local ; shared]
        pos[self.local_rows] = np.arange(self.nl)
        pos[self.shared_rows] = self.nl + np.arange(self.ns)
        self.pos = pos

        def t(v, dt):
            return torch.from_numpy(np.ascontiguousarray(v, dtype=dt)).to(self.device)

        # rows [local ; shared] x this rank's columns, entries in ascending column order
        rows = np.concatenate([self.local_rows, self.shared_rows])
        starts, ends = kAt[rows].astype(np.int64), kAt[rows + 1].astype(np.int64)
        lens = ends - starts
        tot = int(lens.sum())
        src = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(tot)
        cols, vals = iAt[src], At[src]
        row_of = np.repeat(np.arange(len(rows)), lens)
        keep = (cols >= self.c0) & (cols < self.c1)
        cols, vals, row_of = cols[keep] - self.c0, vals[keep], row_of[keep]
        ptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.add.at(ptr, row_of + 1, 1)
        ptr = np.cumsum(ptr)
        self.rowmat = (t(ptr, np.int32), t(cols if len(cols) else np.zeros(1), np.int32), t(vals if len(vals) else np.zeros(1), np.float64), len(cols))
        # this rank's columns x [local ; shared] rows, entries in ascending (global) row order
        a, b = int(kA[self.c0]), int(kA[self.c1])
        cidx = pos[iA[a:b]]
        if len(cidx) and cidx.min() < 0:
            raise ValueError("a column meets a row that is neither local to its rank nor shared")
        self.colmat = (t(kA[self.c0:self.c1 + 1] - kA[self.c0], np.int32), t(cidx if b > a else np.zeros(1), np.int32),
                       t(A[a:b] if b > a else np.zeros(1), np.float64), b - a)
        self.scratch = torch.zeros(int(self.lib.vbk_reduce_scratch_doubles()), dtype=torch.float64, device=self.device)
        self.red9 = torch.zeros(9, dtype=torch.float64, device=self.device)       # 6 partial sums + 3 partial maxima
        self.red9_all = torch.zeros(9 * self.world, dtype=torch.float64, device=self.device)
        self.launches = 0

    def local_x(self, full):
        v = np.zeros(self.cols_per)
        v[: self.c1 - self.c0] = np.asarray(full)[self.c0:self.c1]
        return self.torch.from_numpy(v).to(self.device)

    def local_y(self, full):
        full = np.asarray(full)
        return self.torch.from_numpy(np.concatenate([full[self.local_rows], full[self.shared_rows]])).to(self.device)

    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def _ptrs(self, vecs):
        arr = (C.c_void_p * len(vecs))(*[v.data_ptr() for v in vecs])
        lens = (C.c_longlong * len(vecs))(*[int(v.numel()) for v in vecs])
        return arr, lens

    def step(self, x_local, y_local, rho, sig, overlap=True):
        """rho = A x on [local ; shared] rows, sig = (A^T y)[cols], and the six scalars of hsd.c:182-195.  One all-reduce
        of the shared rows' partial sums and one 72-byte all-gather of the partial scalars; ``overlap`` is accepted for
        interface compatibility (the collectives are stream-ordered either way, which is what graph capture needs)."""
        nl, ns = self.nl, self.ns
        ptr, idx, val, _ = self.rowmat
        self.lib.vbk_spmv_rows_dev(nl + ns, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), x_local.data_ptr(), rho.data_ptr(), self._stream())
        if self.world > 1 and ns > 0:
            self.dist.all_reduce(rho[nl:], op=self.dist.ReduceOp.SUM, group=self.group)
        ptr, idx, val, _ = self.colmat
        self.lib.vbk_spmv_rows_dev(self.c1 - self.c0, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), y_local.data_ptr(), sig.data_ptr(), self._stream())
        pairs = [(x_local, sig), (y_local[:nl], rho[:nl]), (rho[:nl], rho[:nl]), (sig, sig), (y_local[nl:], rho[nl:]), (rho[nl:], rho[nl:])]
        xs, lens = self._ptrs([p[0] for p in pairs])
        ys, _ = self._ptrs([p[1] for p in pairs])
        self.lib.vbk_dots_partial_dev(6, xs, ys, lens, self.red9.data_ptr(), self.scratch.data_ptr(), self._stream())
        vs, vlens = self._ptrs([rho[:nl], sig, rho[nl:]])
        self.lib.vbk_absmax_partial_dev(3, vs, vlens, self.red9[6:].data_ptr(), self._stream())
        self.launches += 5
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.red9_all, self.red9, group=self.group)
            allv = self.red9_all.view(self.world, 9)
        else:
            allv = self.red9.view(1, 9)
        loc = allv[:, :4].sum(dim=0)                   # rank order: every rank holds the same bits
        sums = self.torch.stack([loc[0], loc[1] + allv[0, 4], loc[2] + allv[0, 5], loc[3]])    # the shared tail counts once
        mx = self.torch.stack([self.torch.maximum(allv[:, 6].amax(), allv[0, 8]), allv[:, 7].amax()])
        return sums, mx

    def exchanged_bytes(self):
        """bytes a rank sends or receives per step besides the scalars"""
        return 8 * self.ns if self.world > 1 else 0

