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
