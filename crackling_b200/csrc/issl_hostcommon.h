// issl_hostcommon.h -- what the host programs (isslScoreOfftargets, isslScoreServer) share above the
// C ABI: choosing GPUs, one issl_device per GPU for a given index, guides partitioned into contiguous
// ranges with one host thread per GPU (index replicated, no cross-GPU reduction: guides are independent,
// ref isslScoreOfftargets.cpp:316-317).
#ifndef ISSL_HOSTCOMMON_H
#define ISSL_HOSTCOMMON_H

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "issl_cuda.h"

namespace issl_host {

inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline std::vector<int> devices_from_env()
{
    std::vector<int> devs;
    if (const char *e = getenv("ISSL_DEVICES")) {
        for (const char *p = e; *p;) {
            char *end;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devs.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    return devs;
}

// ISSL_DEVICES wins; else min(ISSL_GPUS, present); else at most one GPU per 65536 guides.
inline std::vector<int> pick_devices(size_t nGuides)
{
    std::vector<int> devs = devices_from_env();
    if (!devs.empty()) return devs;
    int want = issl_device_count();
    if (const char *e = getenv("ISSL_GPUS")) {
        const int v = atoi(e);
        if (v > 0 && v < want) want = v;
    } else {
        const size_t byWork = (nGuides + 65535) / 65536;
        if ((size_t)want > byWork) want = (int)(byWork ? byWork : 1);
    }
    if (want < 1) want = 1;   // device 0: creation will fail loudly when there is no GPU
    for (int d = 0; d < want; d++) devs.push_back(d);
    return devs;
}

inline int layout_from_env()
{
    const char *e = getenv("ISSL_LAYOUT");
    if (!e) return ISSL_LAYOUT_AUTO;
    if (!strcmp(e, "res32")) return ISSL_LAYOUT_RES32;
    if (!strcmp(e, "sig64")) return ISSL_LAYOUT_SIG64;
    if (!strcmp(e, "gather")) return ISSL_LAYOUT_GATHER;
    if (!strcmp(e, "triple")) return ISSL_LAYOUT_TRIPLE;
    return ISSL_LAYOUT_AUTO;
}

inline bool method_has_mit(int m) { return m == ISSL_METHOD_MIT || m == ISSL_METHOD_AND || m == ISSL_METHOD_OR || m == ISSL_METHOD_AVG; }
inline bool method_has_cfd(int m) { return m == ISSL_METHOD_CFD || m == ISSL_METHOD_AND || m == ISSL_METHOD_OR || m == ISSL_METHOD_AVG; }

// One index resident on a set of GPUs.
class DeviceSet {
public:
    ~DeviceSet() { clear(); }
    void clear()
    {
        for (issl_device *d : handles_) issl_device_destroy(d);
        handles_.clear(); devs_.clear();
    }
    const std::vector<int> &devices() const { return devs_; }
    bool has(int dev) const
    {
        for (int d : devs_) if (d == dev) return true;
        return false;
    }

    // Makes the index resident on every device of `devs` that does not hold it yet (in parallel).
    // Returns ISSL_OK or the first error (message in *err).
    int ensure(const issl_index *index, const std::vector<int> &devs, int layout, std::string *err)
    {
        std::vector<int> missing;
        for (int d : devs) if (!has(d)) missing.push_back(d);
        if (missing.empty()) return ISSL_OK;
        std::vector<issl_device *> made(missing.size(), nullptr);
        std::vector<int> rcs(missing.size(), ISSL_OK);
        std::vector<std::string> errors(missing.size());
        auto worker = [&](size_t k) {
            rcs[k] = issl_device_create(index, missing[k], layout, &made[k]);
            if (rcs[k] != ISSL_OK) errors[k] = issl_last_error();
        };
        run_parallel(missing.size(), worker);
        int rc = ISSL_OK;
        for (size_t k = 0; k < missing.size(); k++) {
            if (rcs[k] == ISSL_OK) { devs_.push_back(missing[k]); handles_.push_back(made[k]); }
            else if (rc == ISSL_OK) { rc = rcs[k]; if (err) *err = errors[k]; }
        }
        return rc;
    }

    // Scores guides[0..n) on the devices of `use` (all of which must be resident): contiguous partitions,
    // disjoint output ranges.  stats (optional) receives the sums over devices.
    int score(const std::vector<int> &use, const uint64_t *guides, size_t n, int maxDist, double threshold, int method,
              double *mit, double *cfd, issl_stats *stats, std::string *err, bool verbose = false)
    {
        const size_t nd = use.size();
        std::vector<int> rcs(nd, ISSL_OK);
        std::vector<std::string> errors(nd);
        std::vector<issl_stats> st(nd);
        auto worker = [&](size_t k) {
            issl_device *dev = handle_of(use[k]);
            const size_t b = n * k / nd, e = n * (k + 1) / nd;
            if (!dev) { rcs[k] = ISSL_ERR_ARG; errors[k] = "index is not resident on the requested device"; return; }
            rcs[k] = issl_score(dev, guides + b, e - b, maxDist, threshold, method, mit ? mit + b : nullptr, cfd ? cfd + b : nullptr);
            if (rcs[k] != ISSL_OK) { errors[k] = issl_last_error(); return; }
            issl_last_stats(dev, &st[k]);
            if (verbose)
                fprintf(stderr, "[issl] gpu %d: guides %zu candidates %llu hits %llu early-exits %llu scan %.3f ms device-total %.3f ms\n",
                        use[k], e - b, (unsigned long long)st[k].candidates, (unsigned long long)st[k].hits,
                        (unsigned long long)st[k].early_exits, st[k].scan_ms, st[k].total_ms);
        };
        run_parallel(nd, worker);
        if (stats) {
            memset(stats, 0, sizeof *stats);
            for (size_t k = 0; k < nd; k++) {
                stats->guides += st[k].guides; stats->candidates += st[k].candidates; stats->hits += st[k].hits;
                stats->early_exits += st[k].early_exits; stats->launches += st[k].launches; stats->scan_launches += st[k].scan_launches;
                stats->streamed += st[k].streamed;
                if (st[k].scan_ms > stats->scan_ms) stats->scan_ms = st[k].scan_ms;
                if (st[k].total_ms > stats->total_ms) stats->total_ms = st[k].total_ms;
            }
        }
        for (size_t k = 0; k < nd; k++)
            if (rcs[k] != ISSL_OK) { if (err) *err = errors[k]; return rcs[k]; }
        return ISSL_OK;
    }

private:
    template <class F> static void run_parallel(size_t n, F &f)
    {
        if (n == 1) { f(0); return; }
        std::vector<std::thread> pool;
        for (size_t k = 0; k < n; k++) pool.emplace_back([&f, k] { f(k); });
        for (auto &t : pool) t.join();
    }
    issl_device *handle_of(int dev) const
    {
        for (size_t k = 0; k < devs_.size(); k++) if (devs_[k] == dev) return handles_[k];
        return nullptr;
    }
    std::vector<int> devs_;
    std::vector<issl_device *> handles_;
};

}  // namespace issl_host
#endif
