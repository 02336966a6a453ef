// isslScoreServer -- resident ISSL scorer: keeps .issl indexes in HBM between invocations of
// bin/isslScoreOfftargets, which Crackling's pipeline starts once per page of guides
// (/root/reference/src/crackling/Crackling.py:737-778; the reference re-reads the index every time,
// isslScoreOfftargets.cpp:152-243).  Host program over the C ABI of libissl_cuda; framing in issl_wire.h.
//
//   isslScoreServer <socket path>
//
// Normally started on demand by `ISSL_SERVER=<socket> ISSL_SERVER_AUTOSTART=1 isslScoreOfftargets ...`.
// An index is identified by (real path, size, mtime, inode); a rewritten file is reloaded.
//
// Environment:
//   ISSL_SERVER_IDLE_S=<s>     exit after this many seconds without a request (default 900; 0 = never)
//   ISSL_SERVER_MAX_INDEXES=<n>  resident indexes kept, least recently used dropped first (default 2)
//   ISSL_DEVICES / ISSL_GPUS / ISSL_LAYOUT   as for isslScoreOfftargets, used when a request leaves them open
#include <algorithm>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <poll.h>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "issl_cuda.h"
#include "issl_hostcommon.h"
#include "issl_wire.h"

using namespace issl_wire;
using issl_host::now_s;

namespace {

struct Resident {
    std::string path;
    off_t size = 0;
    ino_t inode = 0;
    timespec mtime{};
    int layout = 0;
    issl_index *index = nullptr;
    issl_host::DeviceSet devices;
    ~Resident() { devices.clear(); if (index) issl_index_close(index); }
};

volatile sig_atomic_t g_stop = 0;
void on_signal(int) { g_stop = 1; }

bool same_file(const Resident &r, const std::string &path, const struct stat &st, int layout)
{
    return r.path == path && r.size == st.st_size && r.inode == st.st_ino && r.layout == layout &&
           r.mtime.tv_sec == st.st_mtim.tv_sec && r.mtime.tv_nsec == st.st_mtim.tv_nsec;
}

void reply(int fd, Response &rsp, const std::string &msg, const double *mit, const double *cfd)
{
    memcpy(rsp.magic, kRspMagic, 8);
    rsp.msgLen = (uint32_t)msg.size();
    if (!write_full(fd, &rsp, sizeof rsp)) return;
    if (!msg.empty() && !write_full(fd, msg.data(), msg.size())) return;
    if (rsp.n) {
        if (!write_full(fd, mit, rsp.n * 8)) return;
        write_full(fd, cfd, rsp.n * 8);
    }
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "Usage: %s [socket path]\n", argv[0]);
        return 1;
    }
    const char *sockPath = argv[1];
    const double idleS = getenv("ISSL_SERVER_IDLE_S") ? atof(getenv("ISSL_SERVER_IDLE_S")) : 900.0;
    const size_t maxIndexes = getenv("ISSL_SERVER_MAX_INDEXES") ? (size_t)std::max(1, atoi(getenv("ISSL_SERVER_MAX_INDEXES"))) : 2;
    if (issl_device_count() < 1) {   // no CPU fallback: a server without a GPU refuses to start
        fprintf(stderr, "isslScoreServer: no usable sm_100 CUDA device\n");
        return 1;
    }

    sockaddr_un addr;
    if (!fill_addr(sockPath, &addr)) { fprintf(stderr, "isslScoreServer: socket path too long\n"); return 1; }
    // a stale socket file from a dead server is replaced; a live server keeps its socket
    if (const int probe = connect_to(sockPath); probe >= 0) {
        close(probe);
        fprintf(stderr, "isslScoreServer: a server is already listening on %s\n", sockPath);
        return 1;
    }
    unlink(sockPath);
    const int lfd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    const mode_t old = umask(0077);
    if (lfd < 0 || bind(lfd, reinterpret_cast<sockaddr *>(&addr), sizeof addr) != 0 || listen(lfd, 16) != 0) {
        perror("isslScoreServer: bind/listen");
        return 1;
    }
    umask(old);
    struct sigaction sa{};
    sa.sa_handler = on_signal;
    sigaction(SIGTERM, &sa, nullptr);
    sigaction(SIGINT, &sa, nullptr);
    signal(SIGPIPE, SIG_IGN);
    fprintf(stderr, "[issl-server] listening on %s (idle timeout %.0f s, %d GPU(s))\n", sockPath, idleS, issl_device_count());

    std::list<std::unique_ptr<Resident>> cache;   // front = most recently used
    double lastActive = now_s();
    bool stop = false;
    while (!stop && !g_stop) {
        pollfd pfd{lfd, POLLIN, 0};
        const int pr = poll(&pfd, 1, 1000);
        if (pr <= 0) {
            if (idleS > 0 && now_s() - lastActive > idleS) { fprintf(stderr, "[issl-server] idle, exiting\n"); break; }
            continue;
        }
        const int fd = accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) continue;
        lastActive = now_s();
        {   // a stalled client must not block the (single-threaded) loop for ever
            timeval tv{};
            tv.tv_sec = 30;
            setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
            setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
        }
        Request req;
        Response rsp{};
        std::string msg;
        if (!read_full(fd, &req, sizeof req) || memcmp(req.magic, kReqMagic, 8) != 0) { close(fd); continue; }
        if (req.op == kPing) { rsp.status = ISSL_OK; rsp.nDevices = (uint32_t)issl_device_count(); reply(fd, rsp, "", nullptr, nullptr); close(fd); continue; }
        if (req.op == kShutdown) { rsp.status = ISSL_OK; reply(fd, rsp, "", nullptr, nullptr); close(fd); stop = true; continue; }
        if (req.op == kDrop) { cache.clear(); rsp.status = ISSL_OK; reply(fd, rsp, "", nullptr, nullptr); close(fd); continue; }
        if (req.op != kScore || req.pathLen == 0 || req.pathLen > 65536 || req.nDevices > (uint32_t)kMaxDevices ||
            req.nGuides > (1ull << 28)) {   // 2^28 guides = 2 GB of packed guides: far beyond a page of the pipeline (5 M)
            rsp.status = ISSL_ERR_ARG; reply(fd, rsp, "malformed request", nullptr, nullptr); close(fd); continue;
        }
        std::string path(req.pathLen, '\0');
        std::vector<uint64_t> guides;
        try { guides.resize(req.nGuides); }
        catch (const std::exception &) { rsp.status = ISSL_ERR_NOMEM; reply(fd, rsp, "out of host memory for the guides", nullptr, nullptr); close(fd); continue; }
        if (!read_full(fd, path.data(), req.pathLen) || (req.nGuides && !read_full(fd, guides.data(), req.nGuides * 8))) { close(fd); continue; }

        struct stat st;
        if (stat(path.c_str(), &st) != 0) {
            rsp.status = ISSL_ERR_IO; reply(fd, rsp, "cannot open " + path, nullptr, nullptr); close(fd); continue;
        }
        // find or load
        Resident *res = nullptr;
        for (auto it = cache.begin(); it != cache.end(); ++it)
            if (same_file(**it, path, st, req.layout)) { cache.splice(cache.begin(), cache, it); res = cache.front().get(); break; }
        rsp.cached = res != nullptr;
        const double t0 = now_s();
        if (!res) {
            // a changed file under the same path replaces the old copy
            cache.remove_if([&](const std::unique_ptr<Resident> &r) { return r->path == path && r->layout == req.layout; });
            while (cache.size() >= maxIndexes) cache.pop_back();
            auto fresh = std::make_unique<Resident>();
            fresh->path = path; fresh->size = st.st_size; fresh->inode = st.st_ino; fresh->mtime = st.st_mtim; fresh->layout = req.layout;
            if (issl_index_open(path.c_str(), &fresh->index) != ISSL_OK) {
                rsp.status = ISSL_ERR_FORMAT; reply(fd, rsp, issl_last_error(), nullptr, nullptr); close(fd); continue;
            }
            cache.push_front(std::move(fresh));
            res = cache.front().get();
        }
        std::vector<int> use;
        for (uint32_t k = 0; k < req.nDevices; k++) {   // ordinals that exist, each once
            bool drop = req.devices[k] < 0 || req.devices[k] >= issl_device_count();
            for (int u : use) drop |= u == req.devices[k];
            if (!drop) use.push_back(req.devices[k]);
        }
        if (use.empty()) use = issl_host::pick_devices(req.nGuides);
        std::string err;
        int rc = res->devices.ensure(res->index, use, req.layout, &err);
        // no room next to the other resident indexes: drop the least recently used ones and try again
        while (rc == ISSL_ERR_NOMEM && cache.size() > 1) {
            fprintf(stderr, "[issl-server] out of device memory, evicting %s\n", cache.back()->path.c_str());
            cache.pop_back();
            rc = res->devices.ensure(res->index, use, req.layout, &err);
        }
        const double t1 = now_s();
        std::vector<double> mit, cfd;
        try { mit.assign(req.nGuides, 0.0); cfd.assign(req.nGuides, 0.0); }
        catch (const std::exception &) { rc = ISSL_ERR_NOMEM; err = "out of host memory for the score arrays"; }
        issl_stats stats{};
        if (rc == ISSL_OK)
            rc = res->devices.score(use, guides.data(), guides.size(), req.maxDist, req.threshold, req.method, mit.data(), cfd.data(), &stats, &err);
        const double t2 = now_s();
        if (rc != ISSL_OK) {
            // a failed load must not stay cached half-resident
            if (res->devices.devices().empty()) cache.pop_front();
            rsp.status = rc; reply(fd, rsp, err, nullptr, nullptr); close(fd); continue;
        }
        rsp.status = ISSL_OK;
        rsp.n = req.nGuides;
        rsp.loadSeconds = t1 - t0; rsp.scoreSeconds = t2 - t1;
        rsp.nDevices = (uint32_t)use.size();
        rsp.candidates = stats.candidates; rsp.hits = stats.hits; rsp.earlyExits = stats.early_exits;
        reply(fd, rsp, "", mit.data(), cfd.data());
        close(fd);
        lastActive = now_s();
        fprintf(stderr, "[issl-server] %s: %llu guides on %zu GPU(s), %s, load %.3f s, score %.3f s\n", path.c_str(),
                (unsigned long long)req.nGuides, use.size(), rsp.cached ? "resident" : "loaded", rsp.loadSeconds, rsp.scoreSeconds);
    }
    cache.clear();
    close(lfd);
    unlink(sockPath);
    return 0;
}
