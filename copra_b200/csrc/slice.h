// slice.h -- sub-batch views of a copra_b200_problem / copra_b200_results: instances are independent, so a contiguous
// range [b0, b0 + count) of a batch is the same problem with every (pointer, stride) pair and every result pointer
// advanced.  Used by the chunked single-device run (capi.cu) and by the multi-device sharder (capi_multi.cu).
#pragma once
#include "../../include/copra_b200.h"
#include <vector>

namespace cb {

struct ProblemSlice {
    copra_b200_problem p;
    std::vector<copra_b200_cost> costs;
    std::vector<copra_b200_constraint> cstrs;
};

inline copra_b200_array slice_array(copra_b200_array a, long long b0)
{
    if (a.ptr && a.stride) a.ptr += b0 * a.stride;
    return a;
}

inline void slice_problem(const copra_b200_problem& src, long long b0, int count, ProblemSlice& out)
{
    out.p = src;
    out.p.batch = count;
    out.p.A = slice_array(src.A, b0); out.p.B = slice_array(src.B, b0); out.p.d = slice_array(src.d, b0); out.p.x0 = slice_array(src.x0, b0);
    out.p.R = slice_array(src.R, b0); out.p.r = slice_array(src.r, b0);
    out.p.x0lb = slice_array(src.x0lb, b0); out.p.x0ub = slice_array(src.x0ub, b0);
    out.costs.assign(src.costs, src.costs + src.ncost);
    out.cstrs.assign(src.cstrs, src.cstrs + src.ncstr);
    for (auto& c : out.costs) { c.M = slice_array(c.M, b0); c.N = slice_array(c.N, b0); c.p = slice_array(c.p, b0); c.w = slice_array(c.w, b0); }
    for (auto& c : out.cstrs) {
        c.E = slice_array(c.E, b0); c.G = slice_array(c.G, b0); c.f = slice_array(c.f, b0);
        c.lower = slice_array(c.lower, b0); c.upper = slice_array(c.upper, b0);
    }
    out.p.costs = out.costs.data();
    out.p.cstrs = out.cstrs.data();
}

inline copra_b200_results slice_results(const copra_b200_results& r, long long b0, const copra_b200_sizes& sz)
{
    copra_b200_results o = r;
    if (r.control) o.control = r.control + b0 * sz.nU;
    if (r.trajectory) o.trajectory = r.trajectory + b0 * sz.X;
    if (r.x) o.x = r.x + b0 * sz.nvar;
    if (r.status) o.status = r.status + b0;
    if (r.iters) o.iters = r.iters + 2 * b0;
    if (r.nact) o.nact = r.nact + b0;
    if (r.iact) o.iact = r.iact + b0 * sz.nvar;
    return o;
}

} // namespace cb
