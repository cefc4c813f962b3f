// zz_host_graph.h -- host-side construction of the neighbour structure the kernels walk, from the two
// Julia-layout CSC matrices the C-ABI receives (target precision Gamma_t and the sampler's Z.Gamma).
//
// For coordinate j the kernel needs, in ONE ascending list (storage order of column j, so that the
// left-to-right accumulation of src/common.jl:16-24 is reproduced bit for bit):
//   * rows of column j of Gamma_t           -> grad phi_j          (flag ZZ_NB_TGT)
//   * rows of column j of Z.Gamma           -> bound a_j, b_j      (flag ZZ_NB_BND)
//   * every k with j in rows(column k of Z.Gamma), i.e. j in G1[k] (src/sfact.jl:170): an accepted flip
//     of k reschedules j (sfact.jl:131-135)                        (flag ZZ_NB_TRIG)
// and the reverse map "who reads j" (dependents) used to propagate dirtiness in the relaxation.
#ifndef ZZ_HOST_GRAPH_H
#define ZZ_HOST_GRAPH_H

#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "zz_core.h"

struct ZzHostGraph {
    int32_t d = 0;
    std::vector<int32_t> nptr, nidx, dptr, didx;
    std::vector<double> nwt, nwb, gmu, h, mu;
    std::vector<uint8_t> nfl;
    int32_t same = 0;
    bool has_h = false;
    bool bnd_eq_tgt = false;   // sampler matrix identical to the target matrix and Z.mu == 0 (needed by LocalBound)
    int32_t maxdeg = 0;
    // 5-point lattice detection (m x n, column-major numbering): lets the kernels use index arithmetic
    int32_t grid_m = 0, grid_n = 0;
    double grid_diag[5] = { 0, 0, 0, 0, 0 };
};

// Is the (single) matrix a 5-point lattice operator  shift*I + gridlaplacian(m, n)  (scripts/gridlaplace.jl:4-21)
// with column-major node numbering?  Then the kernels can replace every index/weight load by arithmetic.
static inline void zz_detect_grid(ZzHostGraph& G, int64_t d, const int64_t* cp, const int64_t* rv, const double* nz)
{
    G.grid_m = G.grid_n = 0;
    if (!G.same || d < 4) return;
    // column 1 of an m x n lattice (m, n >= 2) holds rows {1, 2, 1+m}
    if (cp[1] - cp[0] != 3) return;
    const int64_t m = rv[2] - 1;
    if (m < 2 || d % m != 0 || d / m < 2) return;
    const int64_t n = d / m;
    double diag[5] = { 0, 0, 0, 0, 0 }; bool have[5] = { false, false, false, false, false };
    for (int64_t j = 0; j < d; ++j) {
        const int64_t col = j / m, row = j % m;
        int64_t exp_idx[5]; int cnt = 0, self = 0;
        if (col > 0) exp_idx[cnt++] = j - m;
        if (row > 0) exp_idx[cnt++] = j - 1;
        self = cnt; exp_idx[cnt++] = j;
        if (row < m - 1) exp_idx[cnt++] = j + 1;
        if (col < n - 1) exp_idx[cnt++] = j + m;
        if (cp[j + 1] - cp[j] != cnt) return;
        const int64_t p0 = cp[j] - 1;
        for (int q = 0; q < cnt; ++q) {
            if (rv[p0 + q] - 1 != exp_idx[q]) return;
            if (q == self) {
                const int deg = cnt - 1;
                if (!have[deg]) { have[deg] = true; diag[deg] = nz[p0 + q]; }
                else if (zz_d2u(diag[deg]) != zz_d2u(nz[p0 + q])) return;
            } else if (nz[p0 + q] != -1.0) return;
        }
    }
    G.grid_m = (int32_t)m; G.grid_n = (int32_t)n;
    for (int q = 0; q < 5; ++q) G.grid_diag[q] = diag[q];
}

// Returns "" on success, else an error message (ZZB_E_GRAPH / ZZB_E_ARG material).
static inline std::string zz_build_graph(ZzHostGraph& G, int64_t d, const int64_t* tcp, const int64_t* trv,
                                         const double* tnz, const double* hvec, const int64_t* bcp,
                                         const int64_t* brv, const double* bnz, const double* mu)
{
    if (d <= 0 || d > 0x7ffffff0LL) return "dimension out of range";
    G.d = (int32_t)d;
    auto check = [&](const int64_t* cp, const int64_t* rv, const char* name) -> std::string {
        if (cp[0] != 1) return std::string(name) + ": colptr[1] must be 1 (Julia layout)";
        for (int64_t j = 0; j < d; ++j) {
            if (cp[j + 1] < cp[j]) return std::string(name) + ": colptr not monotone";
            for (int64_t p = cp[j] - 1; p < cp[j + 1] - 1; ++p) {
                if (rv[p] < 1 || rv[p] > d) return std::string(name) + ": row index out of range";
                if (p > cp[j] - 1 && rv[p] <= rv[p - 1]) return std::string(name) + ": rows not strictly ascending in a column";
            }
        }
        return "";
    };
    std::string e = check(tcp, trv, "target");
    if (!e.empty()) return e;
    e = check(bcp, brv, "bound");
    if (!e.empty()) return e;
    // The reference reschedules coordinate i after its own accepted flip only because i is in G1[i] = rows of column i of
    // Z.Gamma (src/sfact.jl:131-135,170); without a stored diagonal entry it would spin on the same queue time for ever.  The
    // device always reschedules the flipping coordinate, so such a matrix would silently sample a different process: refuse it.
    for (int64_t j = 0; j < d; ++j) {
        bool diag = false;
        for (int64_t p = bcp[j] - 1; p < bcp[j + 1] - 1 && !diag; ++p) diag = (brv[p] == j + 1);
        if (!diag) return "bound: column " + std::to_string(j + 1) + " of Z.Gamma has no stored diagonal entry (i must be in G1[i], src/sfact.jl:131-135)";
    }

    // same matrix? (pattern and values identical, no linear term, mu == 0)
    bool same = (tcp[d] == bcp[d]);
    if (same) same = std::equal(tcp, tcp + d + 1, bcp);
    if (same) same = std::equal(trv, trv + (tcp[d] - 1), brv);
    if (same) {
        for (int64_t p = 0; p < tcp[d] - 1 && same; ++p) same = (zz_d2u(tnz[p]) == zz_d2u(bnz[p]));
    }
    bool has_h = false;
    if (hvec) for (int64_t j = 0; j < d && !has_h; ++j) has_h = (hvec[j] != 0.0);
    bool mu0 = true;
    for (int64_t j = 0; j < d && mu0; ++j) mu0 = (zz_d2u(mu[j]) == 0);
    G.same = (same && !has_h && mu0) ? 1 : 0;
    G.bnd_eq_tgt = same && mu0;
    G.has_h = has_h;
    zz_detect_grid(G, d, bcp, brv, bnz);
    if (has_h) G.h.assign(hvec, hvec + d);
    G.mu.assign(mu, mu + d);   // Z.mu itself (the Boomerang flow rotates around it)

    // transpose pattern of the bound matrix: trig[j] = { k : j in rows(col k) }
    std::vector<int64_t> tptr(d + 1, 0);
    for (int64_t p = 0; p < bcp[d] - 1; ++p) tptr[brv[p]]++;  // brv is 1-based -> lands in tptr[row]
    for (int64_t j = 0; j < d; ++j) tptr[j + 1] += tptr[j];
    std::vector<int32_t> tidx(bcp[d] - 1);
    {
        std::vector<int64_t> cur(tptr.begin(), tptr.end() - 1);
        for (int64_t k = 0; k < d; ++k)
            for (int64_t p = bcp[k] - 1; p < bcp[k + 1] - 1; ++p) tidx[cur[brv[p] - 1]++] = (int32_t)k;  // ascending k
    }

    G.nptr.assign(d + 1, 0);
    G.gmu.assign(d, 0.0);
    G.nidx.clear(); G.nwt.clear(); G.nwb.clear(); G.nfl.clear();
    G.maxdeg = 0;
    for (int64_t j = 0; j < d; ++j) {
        int64_t pt = tcp[j] - 1, pte = tcp[j + 1] - 1, pb = bcp[j] - 1, pbe = bcp[j + 1] - 1;
        int64_t pq = tptr[j], pqe = tptr[j + 1];
        double s = 0.0;  // idot(Z.Gamma, j, mu), storage order
        for (int64_t p = pb; p < pbe; ++p) s += bnz[p] * mu[brv[p] - 1];
        G.gmu[j] = s;
        while (pt < pte || pb < pbe || pq < pqe) {
            int64_t kt = pt < pte ? trv[pt] - 1 : INT64_MAX, kb = pb < pbe ? brv[pb] - 1 : INT64_MAX,
                    kq = pq < pqe ? tidx[pq] : INT64_MAX;
            int64_t k = std::min(kt, std::min(kb, kq));
            uint8_t fl = 0; double wt = 0.0, wb = 0.0;
            if (kt == k) { fl |= ZZ_NB_TGT; wt = tnz[pt]; ++pt; }
            if (kb == k) { fl |= ZZ_NB_BND; wb = bnz[pb]; ++pb; }
            if (kq == k) { if (k != j) fl |= ZZ_NB_TRIG; ++pq; }
            G.nidx.push_back((int32_t)k); G.nwt.push_back(wt); G.nwb.push_back(wb); G.nfl.push_back(fl);
        }
        G.nptr[j + 1] = (int32_t)G.nidx.size();
        G.maxdeg = std::max(G.maxdeg, G.nptr[j + 1] - G.nptr[j]);
        if (G.nidx.size() > 0x7ffffff0ULL) return "too many non-zeros";
    }
    // dependents: k reads j  <=>  j in nidx-list of k, k != j
    std::vector<int32_t> cnt(d + 1, 0);
    for (int64_t k = 0; k < d; ++k)
        for (int32_t e2 = G.nptr[k]; e2 < G.nptr[k + 1]; ++e2)
            if (G.nidx[e2] != k) cnt[G.nidx[e2] + 1]++;
    G.dptr.assign(d + 1, 0);
    for (int64_t j = 0; j < d; ++j) G.dptr[j + 1] = G.dptr[j] + cnt[j + 1];
    G.didx.assign(G.dptr[d], 0);
    {
        std::vector<int32_t> cur(G.dptr.begin(), G.dptr.end() - 1);
        for (int64_t k = 0; k < d; ++k)
            for (int32_t e2 = G.nptr[k]; e2 < G.nptr[k + 1]; ++e2)
                if (G.nidx[e2] != k) G.didx[cur[G.nidx[e2]]++] = (int32_t)k;
    }
    return "";
}

#endif  // ZZ_HOST_GRAPH_H
