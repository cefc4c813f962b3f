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
    // Coordinates are renumbered component by component ("new" ids; ascending original id inside a component), so that every
    // chain is a contiguous range.  The matrices are stored column by column in the new numbering, rows relabelled, the ORDER of
    // the entries inside a column unchanged (sums run in the reference's storage order).
    std::vector<int32_t> bcp, brow, tcp, trow, comp;
    std::vector<int32_t> orig;    // new id -> original id (draw streams, trace ids and every per-coordinate array keep the original ids)
    std::vector<int32_t> newid;   // original id -> new id
    std::vector<double> bval, tval;
    int32_t ncmax = 0;        // coordinates of the largest component, rounded up to an even number
    int32_t colmax = 32;      // longest column of either matrix (at least 32), rounded up to an even number
    bool have_tgt = false;
    bool prefer = false;      // small or densely coupled components: the sequential schedule is the faster one (see zzb200.cpp)
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
    // components in the order of their smallest member (= their root); members in ascending order
    std::vector<int32_t> size((size_t)d, 0), start((size_t)d, 0);
    for (int64_t j = 0; j < d; ++j) size[find((int32_t)j)]++;
    S.comp.clear();
    int64_t nmax = 0; int32_t at = 0;
    for (int64_t j = 0; j < d; ++j)
        if (size[j]) { start[j] = at; S.comp.push_back(at); at += size[j]; nmax = std::max<int64_t>(nmax, size[j]); }
    S.comp.push_back((int32_t)d);
    if (nmax > max_nc) { S.why = "a connected component of the dependency graph has " + std::to_string(nmax) + " coordinates (limit of the sequential schedule: " + std::to_string(max_nc) + ")"; return; }
    if (bcp[d] - 1 > 0x7ffffff0LL || (tcp && tcp[d] - 1 > 0x7ffffff0LL)) { S.why = "too many non-zeros"; return; }
    S.orig.assign((size_t)d, 0); S.newid.assign((size_t)d, 0);
    {
        std::vector<int32_t> fill(start);
        for (int64_t j = 0; j < d; ++j) { const int32_t r = find((int32_t)j); const int32_t nj = fill[r]++; S.orig[nj] = (int32_t)j; S.newid[j] = nj; }
    }
    S.ncmax = (int32_t)((nmax + 1) & ~(int64_t)1);
    auto copy = [&](const int64_t* cp, const int64_t* rv, const double* nz, std::vector<int32_t>& ocp, std::vector<int32_t>& orow, std::vector<double>& oval) {
        const int64_t nnz = cp[d] - 1;
        ocp.assign((size_t)d + 1, 0); orow.resize((size_t)nnz); oval.resize((size_t)nnz);
        int64_t w = 0;
        for (int64_t jn = 0; jn < d; ++jn) {
            const int64_t j = S.orig[jn];
            ocp[jn] = (int32_t)w;
            for (int64_t p = cp[j] - 1; p < cp[j + 1] - 1; ++p, ++w) { orow[w] = S.newid[rv[p] - 1]; oval[w] = nz[p]; }
        }
        ocp[d] = (int32_t)w;
    };
    copy(bcp, brv, bnz, S.bcp, S.brow, S.bval);
    S.have_tgt = tcp != nullptr;
    int64_t cm = 32;
    for (int64_t j = 0; j < d; ++j) { cm = std::max(cm, bcp[j + 1] - bcp[j]); if (tcp) cm = std::max(cm, tcp[j + 1] - tcp[j]); }
    S.colmax = (int32_t)((cm + 1) & ~(int64_t)1);
    if (80 * (int64_t)S.ncmax + 8 + 8 * 8 * (72 + 2 * ((int64_t)S.colmax + 8)) > 220 * 1024) {   // (state + the scratch of eight warps)
        S.why = "the largest component (" + std::to_string(nmax) + " coordinates, longest column " + std::to_string(cm) + ") does not fit into the shared memory of an SM";
        return;
    }
    if (tcp) copy(tcp, trv, tnz, S.tcp, S.trow, S.tval);
    // measured on a B200 (tools/seq_vs_window.py): a chain costs ~1.7 us per proposal whatever its size; the windowed relaxation
    // 5.4 / 3.8 / 1.6 / 0.5 us at d = 2 / 16 / 64 / 256 on a lattice, and 28 us on a dense d = 32 (complete dependency graph)
    S.prefer = nmax <= 64 || (nmax >= 8 && (bcp[d] - 1) * 4 >= (int64_t)d * nmax);
    S.ok = true; S.why.clear();
}

#endif  // ZZ_HOST_SEQ_H
