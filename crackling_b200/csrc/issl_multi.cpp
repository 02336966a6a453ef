// issl_multi.cpp -- one process, several GPUs: the guides of one call are cut into chunks that the devices
// take from a shared counter, one host thread per device.  The index is replicated (issl_device_clone), every
// chunk writes its own range of the output arrays, and there is no cross-GPU reduction: guides are independent
// (ref /root/reference/src/ISSL/isslScoreOfftargets.cpp:308-317, `#pragma omp for` over guides with the default
// schedule).  Chunks are handed out dynamically because the early exit (ref :466-502) makes a guide's cost
// uneven: a device that drew guides of a repeat family finishes its chunk early and takes the next one.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "issl_internal.h"

extern "C" size_t issl_multi_chunk(size_t n, size_t n_devs)
{
    if (n_devs <= 1) return n;
    // about eight chunks per device, never below 65 536 guides (launch-bound below that) nor above one internal batch
    size_t chunk = n / (8 * n_devs);
    chunk = std::min<size_t>(std::max<size_t>(chunk, 65536), 1u << 20);
    return (chunk + 4095) / 4096 * 4096;
}

extern "C" int issl_score_multi(issl_device *const *devs, size_t n_devs, const uint64_t *guides, size_t n, int maxDist,
                                double threshold, int method, double *mit_out, double *cfd_out, size_t chunk, issl_stats *stats_out,
                                uint64_t *guides_per_device)
{
    if (!devs || n_devs == 0) return issl_set_error(ISSL_ERR_ARG, "issl_score_multi: no devices");
    for (size_t k = 0; k < n_devs; k++) {
        if (!devs[k]) return issl_set_error(ISSL_ERR_ARG, "issl_score_multi: null device handle");
        for (size_t j = 0; j < k; j++)
            if (devs[j] == devs[k]) return issl_set_error(ISSL_ERR_ARG, "issl_score_multi: the same device handle was passed twice");
    }
    if (n && !guides) return issl_set_error(ISSL_ERR_ARG, "issl_score_multi: null guide array");
    if (chunk == 0) chunk = issl_multi_chunk(n, n_devs);
    if (stats_out) memset(stats_out, 0, sizeof *stats_out);
    if (guides_per_device) std::fill(guides_per_device, guides_per_device + n_devs, 0ull);
    if (n == 0) return ISSL_OK;

    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    std::vector<int> rcs(n_devs, ISSL_OK);
    std::vector<std::string> errors(n_devs);
    std::vector<issl_stats> st(n_devs);
    for (auto &s : st) memset(&s, 0, sizeof s);
    auto worker = [&](size_t k) {
        for (;;) {
            if (failed.load(std::memory_order_relaxed)) return;
            const size_t b = next.fetch_add(chunk);
            if (b >= n) return;
            const size_t e = std::min(n, b + chunk);
            const int rc = issl_score(devs[k], guides + b, e - b, maxDist, threshold, method, mit_out ? mit_out + b : nullptr,
                                      cfd_out ? cfd_out + b : nullptr);
            if (rc != ISSL_OK) { rcs[k] = rc; errors[k] = issl_last_error(); failed.store(1); return; }
            issl_stats s;
            issl_last_stats(devs[k], &s);
            st[k].guides += s.guides; st[k].candidates += s.candidates; st[k].hits += s.hits; st[k].scan_launches += s.scan_launches;
            st[k].launches += s.launches; st[k].scan_ms += s.scan_ms; st[k].total_ms += s.total_ms; st[k].early_exits += s.early_exits;
            st[k].streamed += s.streamed; st[k].bucket_visits += s.bucket_visits; st[k].heavy_hits += s.heavy_hits; st[k].sorted_hits += s.sorted_hits; st[k].heavy_ms += s.heavy_ms;
        }
    };
    if (n_devs == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (size_t k = 0; k < n_devs; k++) pool.emplace_back(worker, k);
        for (auto &t : pool) t.join();
    }
    for (size_t k = 0; k < n_devs; k++)
        if (rcs[k] != ISSL_OK) return issl_set_error(rcs[k], "%s", errors[k].c_str());
    for (size_t k = 0; k < n_devs; k++) {
        if (guides_per_device) guides_per_device[k] = st[k].guides;
        if (!stats_out) continue;
        stats_out->guides += st[k].guides; stats_out->candidates += st[k].candidates; stats_out->hits += st[k].hits;
        stats_out->scan_launches += st[k].scan_launches; stats_out->launches += st[k].launches; stats_out->early_exits += st[k].early_exits;
        stats_out->streamed += st[k].streamed; stats_out->bucket_visits += st[k].bucket_visits;
        stats_out->heavy_hits += st[k].heavy_hits; stats_out->sorted_hits += st[k].sorted_hits;
        // devices run side by side: the call took as long as the busiest one
        stats_out->scan_ms = std::max(stats_out->scan_ms, st[k].scan_ms);
        stats_out->total_ms = std::max(stats_out->total_ms, st[k].total_ms);
        stats_out->heavy_ms = std::max(stats_out->heavy_ms, st[k].heavy_ms);
    }
    return ISSL_OK;
}
