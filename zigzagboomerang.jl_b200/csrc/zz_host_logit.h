// zz_host_logit.h -- host-side preparation of the subsampled logistic target (zz_logit.h) from the Julia-layout arrays
// the C-ABI receives: the design matrix A (n x d, CSC), its transpose At (d x n, CSC; column r = row r of A), the
// responses y / ny and the control-variate point mu (scripts/logistic.jl:21-31,78-107).
//   * 0-based int32 copies of both matrices for the kernels;
//   * u0[row] = idot(At, row, mu) in storage order (the reference recomputes it at every draw, :89; same value);
//   * the dependency pattern: coordinate j reads every coordinate that shares a design row with it (pattern of A'A),
//     handed to zz_build_graph as the "target" pattern so that those coordinates mark j when their flip lists change.
#ifndef ZZ_HOST_LOGIT_H
#define ZZ_HOST_LOGIT_H

#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "zz_logit.h"

struct ZzHostLogit {
    int32_t d = 0, n = 0, k = 0;
    double gamma0 = 0.0;
    std::vector<int32_t> acp, arow, rp, rcol;
    std::vector<double> aval, rval, y, ny, u0;
    // dependency pattern as a Julia-layout CSC (values are zeros: the kernels never read them)
    std::vector<int64_t> dep_cp, dep_rv;
    std::vector<double> dep_nz;
};

static inline std::string zz_build_logit(ZzHostLogit& L, int64_t d, int64_t n, const int64_t* acp, const int64_t* arv,
                                         const double* anz, const int64_t* tcp, const int64_t* trv, const double* tnz,
                                         const double* y, const double* ny, const double* mu, double gamma0, int64_t k)
{
    if (d <= 0 || n <= 0 || d > 0x7ffffff0LL || n > 0x7ffffff0LL) return "logistic target: dimensions out of range";
    if (k < 1 || k > 4096) return "logistic target: the number of subsamples k must lie in 1..4096";
    if (acp[0] != 1 || tcp[0] != 1) return "logistic target: colptr[1] must be 1 (Julia layout)";
    const int64_t nnz = acp[d] - 1;
    if (tcp[n] - 1 != nnz) return "logistic target: A and At hold a different number of entries";
    if (nnz > 0x7ffffff0LL) return "logistic target: too many non-zeros";
    L.d = (int32_t)d; L.n = (int32_t)n; L.k = (int32_t)k; L.gamma0 = gamma0;
    L.acp.resize(d + 1); L.arow.resize(nnz); L.aval.assign(anz, anz + nnz);
    for (int64_t j = 0; j <= d; ++j) L.acp[j] = (int32_t)(acp[j] - 1);
    for (int64_t j = 0; j < d; ++j) {
        if (acp[j + 1] <= acp[j]) return "logistic target: every column of A needs at least one entry (rand over an empty range)";
        for (int64_t p = acp[j] - 1; p < acp[j + 1] - 1; ++p) {
            if (arv[p] < 1 || arv[p] > n) return "logistic target: row index of A out of range";
            if (p > acp[j] - 1 && arv[p] <= arv[p - 1]) return "logistic target: rows of A not strictly ascending in a column";
            L.arow[p] = (int32_t)(arv[p] - 1);
        }
    }
    L.rp.resize(n + 1); L.rcol.resize(nnz); L.rval.assign(tnz, tnz + nnz);
    for (int64_t r = 0; r <= n; ++r) L.rp[r] = (int32_t)(tcp[r] - 1);
    L.y.assign(y, y + n); L.ny.assign(ny, ny + n); L.u0.assign(n, 0.0);
    for (int64_t r = 0; r < n; ++r) {
        if (tcp[r + 1] < tcp[r]) return "logistic target: colptr of At not monotone";
        double s = 0.0;   // idot(At, r, mu), src/common.jl:16-24
        for (int64_t p = tcp[r] - 1; p < tcp[r + 1] - 1; ++p) {
            if (trv[p] < 1 || trv[p] > d) return "logistic target: row index of At out of range";
            if (p > tcp[r] - 1 && trv[p] <= trv[p - 1]) return "logistic target: rows of At not strictly ascending in a column";
            L.rcol[p] = (int32_t)(trv[p] - 1);
            s += tnz[p] * mu[trv[p] - 1];
        }
        L.u0[r] = s;
    }
    // At must be the transpose of A (pattern and values)
    {
        std::vector<int64_t> cur(tcp, tcp + n);
        for (int64_t j = 0; j < d; ++j)
            for (int64_t p = acp[j] - 1; p < acp[j + 1] - 1; ++p) {
                const int64_t r = arv[p] - 1;
                const int64_t q = cur[r]++ - 1;
                if (q >= tcp[r + 1] - 1 || trv[q] - 1 != j || zz_d2u(tnz[q]) != zz_d2u(anz[p])) return "logistic target: At is not the transpose of A";
            }
    }
    // dependency pattern: dep[j] = union of the rows' coordinate sets over column j of A, ascending, j included
    L.dep_cp.assign(d + 1, 1); L.dep_rv.clear();
    std::vector<int32_t> mark(d, -1), tmp;
    for (int64_t j = 0; j < d; ++j) {
        tmp.clear();
        for (int32_t p = L.acp[j]; p < L.acp[j + 1]; ++p) {
            const int32_t r = L.arow[p];
            for (int32_t q = L.rp[r]; q < L.rp[r + 1]; ++q) {
                const int32_t m = L.rcol[q];
                if (mark[m] != (int32_t)j) { mark[m] = (int32_t)j; tmp.push_back(m); }
            }
        }
        std::sort(tmp.begin(), tmp.end());
        for (int32_t m : tmp) L.dep_rv.push_back((int64_t)m + 1);
        L.dep_cp[j + 1] = (int64_t)L.dep_rv.size() + 1;
    }
    L.dep_nz.assign(L.dep_rv.size(), 0.0);
    return "";
}

#endif  // ZZ_HOST_LOGIT_H
