// zz_host_seq.h -- host-side preparation of the sequential-chain schedule (zz_seq.cuh): 0-based int32 copies of the sampler
// matrix (and of the target precision when it is a different object), and the connected components of the dependency graph
//   i ~ j  when  j is a row of column i of Z.Gamma (bound / reschedule, src/sfact.jl:131-135, fact_samplers.jl:50-54),
//                of the target's column i (partial derivative, src/common.jl:16-24),
//                or shares a design row with i (logistic target, scripts/logistic.jl:78-95).
// Components evolve independently of each other, so each gets a warp of its own.  They must be contiguous index ranges
// (block-diagonal problems: replicas side by side); anything else leaves `ok` false and the windowed kernels are used.
#ifndef ZZ_HOST_SEQ_H
#define ZZ_HOST_SEQ_H

#include <stdint.h>
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

struct ZzHostSeq {
    bool ok = false;
    std::string why = "not prepared";
    std::vector<int32_t> bcp, brow, tcp, trow, comp;
    std::vector<double> bval, tval;
    int32_t ncmax = 0;        // coordinates of the largest component, rounded up to an even number
    int32_t colmax = 32;      // longest column of either matrix (at least 32), rounded up to an even number
    bool have_tgt = false;
};

// dep_cp / dep_rv: an extra dependency pattern (Julia-layout CSC, may be null); tcp == nullptr: no separate target matrix.
static inline void zz_build_seq(ZzHostSeq& S, int64_t d, const int64_t* bcp, const int64_t* brv, const double* bnz,
                                const int64_t* tcp, const int64_t* trv, const double* tnz,
                                const int64_t* dep_cp, const int64_t* dep_rv, int64_t max_nc)
{
    S.ok = false;
    std::vector<int32_t> parent((size_t)d);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto unite = [&](int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
    auto scan = [&](const int64_t* cp, const int64_t* rv) {
        for (int64_t j = 0; j < d; ++j)
            for (int64_t p = cp[j] - 1; p < cp[j + 1] - 1; ++p) unite((int32_t)j, (int32_t)(rv[p] - 1));
    };
    scan(bcp, brv);
    if (tcp) scan(tcp, trv);
    if (dep_cp) scan(dep_cp, dep_rv);
    // roots are the smallest index of their component: contiguous iff the root never decreases along the index
    S.comp.clear();
    int32_t last = -1;
    std::vector<char> seen((size_t)d, 0);
    for (int64_t j = 0; j < d; ++j) {
        const int32_t r = find((int32_t)j);
        if (r != last) {
            if (seen[r]) { S.why = "the connected components of the dependency graph are not contiguous index ranges"; return; }
            seen[r] = 1; last = r;
            S.comp.push_back((int32_t)j);
        }
    }
    S.comp.push_back((int32_t)d);
    int64_t nmax = 0;
    for (size_t q = 0; q + 1 < S.comp.size(); ++q) nmax = std::max<int64_t>(nmax, S.comp[q + 1] - S.comp[q]);
    if (nmax > max_nc) { S.why = "a connected component has " + std::to_string(nmax) + " coordinates (limit of the sequential schedule: " + std::to_string(max_nc) + ")"; return; }
    S.ncmax = (int32_t)((nmax + 1) & ~(int64_t)1);
    auto copy = [&](const int64_t* cp, const int64_t* rv, const double* nz, std::vector<int32_t>& ocp, std::vector<int32_t>& orow, std::vector<double>& oval) {
        const int64_t nnz = cp[d] - 1;
        ocp.resize((size_t)d + 1); orow.resize((size_t)nnz); oval.assign(nz, nz + nnz);
        for (int64_t j = 0; j <= d; ++j) ocp[j] = (int32_t)(cp[j] - 1);
        for (int64_t p = 0; p < nnz; ++p) orow[p] = (int32_t)(rv[p] - 1);
    };
    if (bcp[d] - 1 > 0x7ffffff0LL || (tcp && tcp[d] - 1 > 0x7ffffff0LL)) { S.why = "too many non-zeros"; return; }
    copy(bcp, brv, bnz, S.bcp, S.brow, S.bval);
    S.have_tgt = tcp != nullptr;
    int64_t cm = 32;
    for (int64_t j = 0; j < d; ++j) { cm = std::max(cm, bcp[j + 1] - bcp[j]); if (tcp) cm = std::max(cm, tcp[j + 1] - tcp[j]); }
    S.colmax = (int32_t)((cm + 1) & ~(int64_t)1);
    if (tcp) copy(tcp, trv, tnz, S.tcp, S.trow, S.tval);
    S.ok = true; S.why.clear();
}

#endif  // ZZ_HOST_SEQ_H
