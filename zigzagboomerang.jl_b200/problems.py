"""Synthetic inputs for the configurations named in BASELINE.json / SURVEY.md 8(d).

All matrices are returned as :class:`CSC` in the memory layout Julia's ``SparseMatrixCSC{Float64,Int64}``
uses (1-based ``colptr`` / ``rowval``, row indices ascending inside a column), because that is what the
C-ABI takes (``include/zzb200.h``) and what the reference iterates over (``src/common.jl:16-24``).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class CSC:
    """Julia-layout compressed sparse column matrix (1-based indices, int64)."""

    n: int
    colptr: np.ndarray  # int64[n+1], 1-based
    rowval: np.ndarray  # int64[nnz], 1-based, ascending per column
    nzval: np.ndarray  # float64[nnz]

    @staticmethod
    def from_scipy(A) -> "CSC":
        import scipy.sparse as sp

        A = sp.csc_matrix(A)
        A.sort_indices()
        return CSC(
            A.shape[0],
            np.ascontiguousarray(A.indptr.astype(np.int64) + 1),
            np.ascontiguousarray(A.indices.astype(np.int64) + 1),
            np.ascontiguousarray(A.data.astype(np.float64)),
        )

    @staticmethod
    def from_dense(M) -> "CSC":
        """Dense -> CSC keeping every non-zero entry (column-major scan like Julia's ``sparse``)."""
        M = np.asarray(M, dtype=np.float64)
        n = M.shape[0]
        colptr = [1]
        rows, vals = [], []
        for j in range(n):
            nz = np.nonzero(M[:, j])[0]
            rows.extend((nz + 1).tolist())
            vals.extend(M[nz, j].tolist())
            colptr.append(len(rows) + 1)
        return CSC(n, np.array(colptr, np.int64), np.array(rows, np.int64), np.array(vals, np.float64))

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.n, self.n))

    def scaled(self, s: float) -> "CSC":
        return CSC(self.n, self.colptr, self.rowval, self.nzval * s)

    def colnorms(self) -> np.ndarray:
        """``[norm(A[:, i], 2) for i in 1:n]`` (``scripts/gaussianrandomfield.jl:33``); every column non-empty."""
        sq = self.nzval * self.nzval
        return np.sqrt(np.add.reduceat(sq, self.colptr[:-1] - 1))

    @property
    def nnz(self) -> int:
        return int(self.nzval.shape[0])


def grid_precision(m: int, n: int | None = None, shift: float = 0.01) -> CSC:
    """``shift*I + gridlaplacian(m, n)`` of ``scripts/gridlaplace.jl:4-21`` /
    ``scripts/gaussianrandomfield.jl:14-15``: 5-point graph Laplacian of an m x n lattice, nodes numbered
    column-major (``LinearIndices((1:m, 1:n))``), built directly in CSC form (no scipy needed at d = 10^6).
    """
    n = m if n is None else n
    d = m * n
    idx = np.arange(d, dtype=np.int64)
    i = idx % m  # row in the lattice (fast index)
    j = idx // m
    has_up = i > 0  # neighbour idx-1
    has_dn = i < m - 1  # neighbour idx+1
    has_lf = j > 0  # neighbour idx-m
    has_rt = j < n - 1  # neighbour idx+m
    deg = has_up.astype(np.int64) + has_dn + has_lf + has_rt
    cnt = deg + 1
    colptr = np.empty(d + 1, np.int64)
    colptr[0] = 1
    np.cumsum(cnt, out=colptr[1:])
    colptr[1:] += 1
    nnz = int(colptr[-1] - 1)
    rowval = np.empty(nnz, np.int64)
    nzval = np.empty(nnz, np.float64)
    pos = colptr[:-1] - 1  # write cursor per column (0-based)
    # ascending row order inside a column: idx-m, idx-1, idx, idx+1, idx+m
    for mask, off, val in ((has_lf, -m, -1.0), (has_up, -1, -1.0), (None, 0, None), (has_dn, 1, -1.0), (has_rt, m, -1.0)):
        if mask is None:
            rowval[pos] = idx + 1
            nzval[pos] = shift + deg.astype(np.float64)
            pos = pos + 1
        else:
            p = pos[mask]
            rowval[p] = idx[mask] + off + 1
            nzval[p] = val
            pos = pos + mask.astype(np.int64)
    return CSC(d, colptr, rowval, nzval)


def gmrf_config(n: int, seed: int = 1, tight: bool = False):
    """Configs 2 / 5 of SURVEY.md 8(d): local ZigZag on ``0.01 I + gridlaplacian(n, n)``.

    Returns ``(Gamma, x0, theta0, c)``; ``c = ||Gamma[:, i]||_2`` as in ``scripts/gaussianrandomfield.jl:33``
    or the tight ``sqrt(eps)`` variant of ``scripts/example.jl:39``.
    """
    G = grid_precision(n, n)
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal(n * n)
    theta0 = rng.choice(np.array([-1.0, 1.0]), size=n * n)
    c = np.full(n * n, np.sqrt(np.finfo(np.float64).eps)) if tight else G.colnorms()
    return G, x0, theta0, c


def random_spd(d: int, seed: int = 2, density: float = 0.1) -> CSC:
    """``S = 1.3 I + 0.5 sprandn(d, d, density); Gamma = S S'`` of ``test/maintest.jl:4-8`` (own RNG)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((d, d)) < density
    S = 1.3 * np.eye(d) + 0.5 * np.where(mask, rng.standard_normal((d, d)), 0.0)
    return CSC.from_dense(S @ S.T)


def random_sparse_spd(d: int, deg: int = 3, seed: int = 0) -> CSC:
    """Random symmetric diagonally dominant sparse precision (general-graph parity cases)."""
    rng = np.random.default_rng(seed)
    M = np.zeros((d, d))
    for i in range(d):
        for j in rng.choice(d, size=deg, replace=False):
            if i != j:
                w = rng.uniform(-1.0, 1.0)
                M[i, j] += w
                M[j, i] += w
    M[np.diag_indices(d)] = np.abs(M).sum(axis=1) + rng.uniform(0.1, 1.0, size=d)
    return CSC.from_dense(M)


@dataclasses.dataclass
class RectCSC:
    """Julia-layout CSC of a rectangular matrix (design matrices of regression targets)."""

    nrows: int
    ncols: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    @staticmethod
    def from_dense(M) -> "RectCSC":
        """``sparse(M)``: column-major scan keeping every non-zero entry."""
        M = np.asarray(M, dtype=np.float64)
        rows, cols = np.nonzero(M.T)[::-1]  # column-major order: sort by column, then row
        colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(cols, minlength=M.shape[1]))]).astype(np.int64)
        return RectCSC(M.shape[0], M.shape[1], colptr, np.ascontiguousarray(rows.astype(np.int64) + 1),
                       np.ascontiguousarray(M[rows, cols]))

    def transpose(self) -> "RectCSC":
        """``SparseMatrixCSC(A')`` (scripts/logistic.jl:31)."""
        cols = np.repeat(np.arange(self.ncols), np.diff(self.colptr))
        order = np.argsort(self.rowval, kind="stable")  # by row, columns ascending inside a row
        colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(self.rowval - 1, minlength=self.nrows))]).astype(np.int64)
        return RectCSC(self.ncols, self.nrows, colptr, np.ascontiguousarray(cols[order].astype(np.int64) + 1),
                       np.ascontiguousarray(self.nzval[order]))

    def to_dense(self) -> np.ndarray:
        M = np.zeros((self.nrows, self.ncols))
        cols = np.repeat(np.arange(self.ncols), np.diff(self.colptr))
        M[self.rowval - 1, cols] = self.nzval
        return M


def sparse_design(d=(5, 5, 5), r: int = 5, m: int = 100, rng=None) -> np.ndarray:
    """``sparse_design(d, r, m)`` of ``scripts/sparsedesign.jl:1-25`` (dense result; own RNG): ``len(d)`` categorical factors
    (one random level per row, switched on with probability ``(d_k - 1) / d_k``), their pairwise interactions
    (``0.3`` when both levels are on) and ``r`` continuous regressors ``0.1 randn``; ``n = m p`` rows."""
    rng = np.random.default_rng(0) if rng is None else rng
    d = tuple(int(v) for v in d)
    K = len(d)
    p = sum(d) + (sum(d) ** 2 - sum(v * v for v in d)) // 2 + r
    n = m * p
    D = np.concatenate([[0], np.cumsum(d)])
    A = np.zeros((n, p))
    rows = np.arange(n)
    j = int(D[-1])
    for k in range(K):
        lev = rng.integers(0, d[k], size=n)
        A[rows, D[k] + lev] = (rng.random(n) < (d[k] - 1) / d[k]).astype(np.float64)
        for k2 in range(k):
            for c2 in range(d[k2]):  # CartesianIndices((d[k], d[k2])): first index fastest
                for c1 in range(d[k]):
                    A[:, j] = 0.3 * ((A[:, D[k] + c1] == 1) & (A[:, D[k2] + c2] == 1))
                    j += 1
    for _ in range(r):
        A[:, j] = 0.1 * rng.standard_normal(n)
        j += 1
    assert j == p
    return A


def logistic_config(levels=(20, 20), r: int = 2, m: int = 20, seed: int = 2, gamma0: float = 0.01, droptol: float = 1e-2,
                    newton_steps: int = 30):
    """Config 3 of SURVEY.md 8(d): the sparse logistic regression of ``scripts/logistic.jl:21-158`` (README: n = 8840,
    p = 442 for ``m = 20``).  Returns a dict with the design ``A`` / ``At`` (:class:`RectCSC`), responses ``y`` / ``ny``, the
    mode ``mu`` (30 Newton steps, :120-125), the Hessian at the mode ``Gamma`` and its sparsified copy ``Gamma_drop``
    (``droptol!(copy(Gamma), 1e-2)``, :137; both :class:`CSC`), ``sigma = sqrt(diag(inv(Gamma)))`` (:150), ``x0 = mu``,
    ``theta0 = +-sigma`` (:158), ``c = 0.01`` (:148) and ``gamma0``.  The Hessian is the analytic
    ``gamma0 I + A' diag(s (1 - s)) A`` (the script differentiates the gradient with ReverseDiff, :134)."""
    rng = np.random.default_rng(seed)
    Ad = sparse_design(levels, r, m, rng)
    n, p = Ad.shape
    xtrue = 5 * rng.standard_normal(p)
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    y = (rng.random(n) < sig(Ad @ xtrue)).astype(np.float64)
    ny = 1.0 - y
    x = 0.1 * rng.random(p)
    for _ in range(newton_steps):
        s = sig(Ad @ x)
        grad = gamma0 * x + Ad.T @ (s - y)  # = gamma0 x - A'(y sigmoidn(Ax)) - A'(ny nsigmoid(Ax)), :102
        H = gamma0 * np.eye(p) + (Ad.T * (s * (1 - s))) @ Ad
        x = x - np.linalg.solve(H, grad)
    s = sig(Ad @ x)
    H = gamma0 * np.eye(p) + (Ad.T * (s * (1 - s))) @ Ad
    H = 0.5 * (H + H.T)
    pattern = (Ad != 0).T.astype(np.float64) @ (Ad != 0).astype(np.float64) > 0  # structure of A'A (:128-129)
    Gamma = CSC.from_dense(np.where(pattern, H, 0.0))
    # the diagonal is always kept: the script asserts nnz(diag(Gamma_drop)) == p (:138), which holds at full size anyway
    Hd = np.where((pattern & (np.abs(H) > droptol)) | np.eye(p, dtype=bool), H, 0.0)
    sigma = np.sqrt(np.diag(np.linalg.inv(H)))
    theta0 = rng.choice(np.array([-1.0, 1.0]), size=p) * sigma
    A = RectCSC.from_dense(Ad)
    return dict(A=A, At=A.transpose(), y=y, ny=ny, mu=x.copy(), Gamma=Gamma, Gamma_drop=CSC.from_dense(Hd), sigma=sigma,
                x0=x.copy(), theta0=theta0, c=np.full(p, 0.01), gamma0=gamma0, n=n, p=p)


def _block_diag_csc(blocks_rows: int, blocks_cols: int, colptr, rowval, nzval, R: int):
    nnz = len(nzval)
    cp = np.concatenate([[1]] + [colptr[1:] + r * nnz for r in range(R)]).astype(np.int64)
    rv = np.concatenate([rowval + r * blocks_rows for r in range(R)]).astype(np.int64)
    return cp, rv, np.tile(nzval, R)


def replicate_logistic(cfg: dict, R: int) -> dict:
    """R independent replicas of a logistic configuration as ONE block-diagonal problem (design ``I_R (x) A``, sampler matrix
    ``I_R (x) Gamma_drop``): the components never interact, every coordinate has its own counter stream, so one device run
    simulates R independent chains side by side (config 3 is "plumbing + replicas": one chain of p = 442 coordinates with a
    complete dependency graph cannot fill a GPU)."""
    A, G = cfg["A"], cfg["Gamma_drop"]
    cp, rv, nz = _block_diag_csc(A.nrows, A.ncols, A.colptr, A.rowval, A.nzval, R)
    AR = RectCSC(A.nrows * R, A.ncols * R, cp, rv, nz)
    gcp, grv, gnz = _block_diag_csc(G.n, G.n, G.colptr, G.rowval, G.nzval, R)
    out = dict(cfg)
    out.update(A=AR, At=AR.transpose(), y=np.tile(cfg["y"], R), ny=np.tile(cfg["ny"], R), mu=np.tile(cfg["mu"], R),
               Gamma_drop=CSC(G.n * R, gcp, grv, gnz), sigma=np.tile(cfg["sigma"], R), x0=np.tile(cfg["x0"], R),
               theta0=np.tile(cfg["theta0"], R), c=np.tile(cfg["c"], R), n=cfg["n"] * R, p=cfg["p"] * R, replicas=R)
    out.pop("Gamma", None)
    out.pop("logistic", None)
    return out
