// issl_hostcommon.h -- what the host programs (isslScoreOfftargets, isslScoreServer) share above the
// C ABI: choosing GPUs, one issl_device per GPU for a given index, guides partitioned into contiguous
// ranges with one host thread per GPU (index replicated, no cross-GPU reduction: guides are independent,
// ref isslScoreOfftargets.cpp:316-317).
#ifndef ISSL_HOSTCOMMON_H
#define ISSL_HOSTCOMMON_H

#include <algorithm>
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

// ISSL_DEVICES wins; else min(ISSL_GPUS, present); else at most one GPU per 65536 guides.  Ordinals outside the
// machine and repeated ordinals are dropped (two host threads must never drive one handle).
inline std::vector<int> pick_devices(size_t nGuides)
{
    const int present = issl_device_count();
    std::vector<int> devs;
    for (int d : devices_from_env()) {
        bool seen = d < 0 || d >= (present > 0 ? present : 1);
        for (int e : devs) seen |= e == d;
        if (!seen) devs.push_back(d);
    }
    if (!devs.empty()) return devs;
    int want = present;
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

    // Makes the index resident on every device of `devs` that does not hold it yet: the file goes to ONE GPU
    // (issl_device_create: upload, validation, layout), every further GPU receives a peer copy of the finished layout
    // over NVLink (issl_device_clone), fanned out as a tree -- every replica made in one round is a source in the next.
    // ISSL_FANOUT=0: every GPU loads the file itself, in parallel (the round-1 behaviour, kept for comparison).
    // Returns ISSL_OK or the first error (message in *err).  seconds[0] / seconds[1] (optional): first load / fan-out.
    int ensure(const issl_index *index, const std::vector<int> &devs, int layout, std::string *err, double *seconds = nullptr)
    {
        std::vector<int> missing;
        for (int d : devs) {
            bool dup = has(d);
            for (int m : missing) dup |= m == d;
            if (!dup) missing.push_back(d);
        }
        if (seconds) seconds[0] = seconds[1] = 0;
        if (missing.empty()) return ISSL_OK;
        const bool fanout = !(getenv("ISSL_FANOUT") && atoi(getenv("ISSL_FANOUT")) == 0);
        const double t0 = now_s();
        if (handles_.empty() || !fanout) {
            // from the file: one device when replicas follow, all of them otherwise
            const size_t n = fanout ? 1 : missing.size();
            std::vector<issl_device *> made(n, nullptr);
            std::vector<int> rcs(n, ISSL_OK);
            std::vector<std::string> errors(n);
            auto worker = [&](size_t k) {
                rcs[k] = issl_device_create(index, missing[k], layout, &made[k]);
                if (rcs[k] != ISSL_OK) errors[k] = issl_last_error();
            };
            run_parallel(n, worker);
            int rc = ISSL_OK;
            for (size_t k = 0; k < n; k++) {
                if (rcs[k] == ISSL_OK) { devs_.push_back(missing[k]); handles_.push_back(made[k]); }
                else if (rc == ISSL_OK) { rc = rcs[k]; if (err) *err = errors[k]; }
            }
            if (rc != ISSL_OK) return rc;
            missing.erase(missing.begin(), missing.begin() + (long)n);
        }
        const double t1 = now_s();
        if (seconds) seconds[0] = t1 - t0;
        while (!missing.empty()) {
            const size_t n = std::min(missing.size(), handles_.size());
            std::vector<issl_device *> made(n, nullptr);
            std::vector<int> rcs(n, ISSL_OK);
            std::vector<std::string> errors(n);
            auto worker = [&](size_t k) {
                rcs[k] = issl_device_clone(handles_[k], missing[k], &made[k]);
                if (rcs[k] != ISSL_OK) {   // no peer path, or no room for a copy made that way: load the file instead
                    rcs[k] = issl_device_create(index, missing[k], layout, &made[k]);
                    if (rcs[k] != ISSL_OK) errors[k] = issl_last_error();
                }
            };
            run_parallel(n, worker);
            int rc = ISSL_OK;
            for (size_t k = 0; k < n; k++) {
                if (rcs[k] == ISSL_OK) { devs_.push_back(missing[k]); handles_.push_back(made[k]); }
                else if (rc == ISSL_OK) { rc = rcs[k]; if (err) *err = errors[k]; }
            }
            if (rc != ISSL_OK) return rc;
            missing.erase(missing.begin(), missing.begin() + (long)n);
        }
        if (seconds) seconds[1] = now_s() - t1;
        return ISSL_OK;
    }

    // Scores guides[0..n) on the devices of `use` (all of which must be resident) through issl_score_multi: chunks of
    // guides handed out dynamically, one host thread per device, disjoint output ranges.  stats (optional) receives the
    // sums over devices.
    int score(const std::vector<int> &use, const uint64_t *guides, size_t n, int maxDist, double threshold, int method,
              double *mit, double *cfd, issl_stats *stats, std::string *err, bool verbose = false)
    {
        std::vector<issl_device *> hs;
        for (int d : use) {
            issl_device *h = handle_of(d);
            if (!h) { if (err) *err = "index is not resident on the requested device"; return ISSL_ERR_ARG; }
            bool dup = false;
            for (issl_device *o : hs) dup |= o == h;
            if (!dup) hs.push_back(h);
        }
        issl_stats st;
        std::vector<uint64_t> per(hs.size(), 0);
        size_t chunk = 0;
        if (const char *e = getenv("ISSL_CHUNK")) { const long long v = atoll(e); if (v > 0) chunk = (size_t)v; }
        const int rc = issl_score_multi(hs.data(), hs.size(), guides, n, maxDist, threshold, method, mit, cfd, chunk, &st, per.data());
        if (rc != ISSL_OK) { if (err) *err = issl_last_error(); return rc; }
        if (verbose) {
            fprintf(stderr, "[issl] %zu gpu(s), chunks of %zu guides: candidates %llu hits %llu early-exits %llu scan %.3f ms device-total %.3f ms (busiest gpu); guides per gpu:",
                    hs.size(), chunk ? chunk : issl_multi_chunk(n, hs.size()), (unsigned long long)st.candidates, (unsigned long long)st.hits,
                    (unsigned long long)st.early_exits, st.scan_ms, st.total_ms);
            for (uint64_t g : per) fprintf(stderr, " %llu", (unsigned long long)g);
            fprintf(stderr, "\n");
        }
        if (stats) *stats = st;
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
