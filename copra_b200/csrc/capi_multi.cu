// capi_multi.cu -- the data-parallel sharder of the batched engine: a batch of independent controllers is
// split by instance index into contiguous ranges [g*ceil(B/G), (g+1)*ceil(B/G)), one range per device.  Every device has
// its own engine handle (stream, workspace, pinned staging) driven by its own host thread; parameters go up per shard
// and the results of every shard are written by DMA straight into the caller's (ideally page-locked) result buffers --
// that final gather is the only cross-device step, there is no collective on the path.  Built on the public
// single-device C ABI only, so the results of a shard are bit-identical to a single-device run of the same instances.
#include "../../include/copra_b200.h"
#include "slice.h"

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

using namespace cb;

struct copra_b200_multi {
    std::vector<copra_b200_handle*> h;
    std::vector<int> dev;
    std::vector<int> lo, hi;   // shard ranges of the last run
    std::vector<int> rc;
    std::string err;
    double wall_ms = 0.0;
};

namespace {

int mfail(copra_b200_multi* m, int code, const std::string& msg)
{
    if (m) m->err = msg;
    return code;
}

void shard_ranges(copra_b200_multi* m, int batch)
{
    const int G = int(m->h.size());
    const int per = (batch + G - 1) / G;
    m->lo.assign(G, 0);
    m->hi.assign(G, 0);
    for (int g = 0; g < G; ++g) {
        m->lo[g] = std::min(batch, g * per);
        m->hi[g] = std::min(batch, (g + 1) * per);
    }
}

template <class F> int run_shards(copra_b200_multi* m, F body)
{
    const int G = int(m->h.size());
    m->rc.assign(G, 0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g)
        if (m->hi[g] > m->lo[g]) th.emplace_back([&, g] { m->rc[g] = body(g); });
    if (m->hi[0] > m->lo[0]) m->rc[0] = body(0); // the calling thread drives device 0's shard
    for (auto& t : th) t.join();
    m->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int g = 0; g < G; ++g)
        if (m->rc[g]) return mfail(m, m->rc[g], "device " + std::to_string(m->dev[g]) + ": " + copra_b200_last_error(m->h[g]));
    return 0;
}

} // namespace

extern "C" {

int copra_b200_multi_create(const int* devices, int ndev, copra_b200_multi** out)
{
    if (!out) return COPRA_B200_E_ARG;
    *out = nullptr;
    const int avail = copra_b200_device_count();
    if (avail <= 0) return COPRA_B200_E_NOGPU;
    std::vector<int> devs;
    if (!devices || ndev <= 0) for (int d = 0; d < avail; ++d) devs.push_back(d);
    else devs.assign(devices, devices + ndev);
    copra_b200_multi* m = new copra_b200_multi();
    for (int d : devs) {
        copra_b200_options opt{};
        opt.device = d;
        copra_b200_handle* h = nullptr;
        const int rc = copra_b200_create(&opt, &h);
        if (rc) {
            for (auto* hh : m->h) copra_b200_destroy(hh);
            delete m;
            return rc;
        }
        m->h.push_back(h);
        m->dev.push_back(d);
    }
    *out = m;
    return COPRA_B200_OK;
}

void copra_b200_multi_destroy(copra_b200_multi* m)
{
    if (!m) return;
    for (auto* h : m->h) copra_b200_destroy(h);
    delete m;
}

const char* copra_b200_multi_last_error(const copra_b200_multi* m) { return m ? m->err.c_str() : "null handle"; }
int copra_b200_multi_size(const copra_b200_multi* m) { return m ? int(m->h.size()) : 0; }

int copra_b200_multi_shard(const copra_b200_multi* m, int g, int* device, int* lo, int* hi)
{
    if (!m || g < 0 || g >= int(m->h.size())) return COPRA_B200_E_ARG;
    if (device) *device = m->dev[g];
    if (lo) *lo = g < int(m->lo.size()) ? m->lo[g] : 0;
    if (hi) *hi = g < int(m->hi.size()) ? m->hi[g] : 0;
    return 0;
}

int copra_b200_multi_timing(const copra_b200_multi* m, int g, copra_b200_timing* t, double* wall_ms)
{
    if (!m || g < 0 || g >= int(m->h.size())) return COPRA_B200_E_ARG;
    if (wall_ms) *wall_ms = m->wall_ms;
    return t ? copra_b200_last_timing(m->h[g], t) : 0;
}

long long copra_b200_multi_launch_count(const copra_b200_multi* m)
{
    long long n = 0;
    if (m) for (auto* h : m->h) n += copra_b200_launch_count(h);
    return n;
}

int copra_b200_multi_lmpc_run(copra_b200_multi* m, const copra_b200_problem* p, const copra_b200_results* r)
{
    if (!m) return COPRA_B200_E_ARG;
    if (!p) return mfail(m, COPRA_B200_E_ARG, "null problem");
    if (p->memory != COPRA_B200_HOST || (r && r->memory != COPRA_B200_HOST))
        return mfail(m, COPRA_B200_E_ARG, "the multi-device entry takes HOST arrays (each shard is uploaded to its own device)");
    copra_b200_sizes sz;
    int rc = copra_b200_lmpc_sizes(m->h[0], p, &sz);
    if (rc) return mfail(m, rc, copra_b200_last_error(m->h[0]));
    shard_ranges(m, p->batch);
    return run_shards(m, [&](int g) {
        ProblemSlice q;
        slice_problem(*p, m->lo[g], m->hi[g] - m->lo[g], q);
        copra_b200_results rr{};
        if (r) rr = slice_results(*r, m->lo[g], sz);
        return copra_b200_lmpc_run(m->h[g], &q.p, r ? &rr : nullptr);
    });
}

int copra_b200_multi_set_warm_start(copra_b200_multi* m, int on)
{
    if (!m) return COPRA_B200_E_ARG;
    for (copra_b200_handle* h : m->h) copra_b200_set_warm_start(h, on);
    return 0;
}

int copra_b200_multi_lmpc_resolve(copra_b200_multi* m, copra_b200_array x0, const copra_b200_results* r)
{
    if (!m) return COPRA_B200_E_ARG;
    if (!x0.ptr) return mfail(m, COPRA_B200_E_ARG, "x0 is required");
    if (m->lo.size() != m->h.size()) return mfail(m, COPRA_B200_E_STATE, "copra_b200_multi_lmpc_resolve needs a previous run");
    if (r && r->memory != COPRA_B200_HOST) return mfail(m, COPRA_B200_E_ARG, "the multi-device entry takes HOST arrays");
    return run_shards(m, [&](int g) {
        copra_b200_results rr{};
        if (r) {
            // per-instance result sizes are those of the resident build: recover them from the shard's own handle
            copra_b200_sizes sz{};
            const int rc = copra_b200_lmpc_built_sizes(m->h[g], &sz);
            if (rc) return rc;
            rr = slice_results(*r, m->lo[g], sz);
        }
        return copra_b200_lmpc_resolve(m->h[g], slice_array(x0, m->lo[g]), COPRA_B200_HOST, r ? &rr : nullptr);
    });
}

} // extern "C"
