// isslScoreOfftargets -- drop-in replacement for Crackling's ISSL off-target scorer
// (/root/reference/src/ISSL/isslScoreOfftargets.cpp), host program over the C ABI of libissl_cuda.
//
//   isslScoreOfftargets <index.issl> <guides.txt> <max distance> <score-threshold> <score-method>
//
// Same five positional arguments, same .issl file (as written by the reference isslCreateIndex),
// same guide file, same stdout lines (`SEQ\tMIT\tCFD\n`, %f or the literal -1), diagnostics on
// stderr only, exit status 0 / 1 -- the contract Crackling's pipeline relies on at
// src/crackling/Crackling.py:767-786.  The scoring itself runs on B200 GPUs; there is no CPU path.
//
// Environment (the Python caller cannot pass extra argv):
//   ISSL_GPUS=<n>        number of GPUs to use (default: as many as there are, at most one per 65536 guides)
//   ISSL_DEVICES=a,b,..  explicit CUDA device ordinals (overrides ISSL_GPUS)
//   ISSL_LAYOUT=res32|sig64|gather   HBM layout of the slice lists (default: automatic)
//   ISSL_TIMING=1        phase timings on stderr
//   ISSL_SERVER=<socket> score through a resident isslScoreServer on that unix socket (the index stays in HBM
//                        between invocations -- the pipeline starts this program once per page of guides,
//                        Crackling.py:737-778); ISSL_SERVER_AUTOSTART=1 starts the server when nobody listens
//                        (ISSL_SERVER_LOG=<file> receives its stderr).  Output is the same either way.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include "issl_cuda.h"
#include "issl_hostcommon.h"
#include "issl_wire.h"

namespace {

using issl_host::now_s;

// Guide and score arrays.  From 65 536 guides on they live in pinned, portable host memory (issl_host_alloc) so that
// every GPU copies straight out of / into them; small runs keep plain memory (pinning costs more than it saves).
template <class T> struct HostArray {
    T *p = nullptr;
    size_t n = 0;
    bool pinned = false;
    explicit HostArray(size_t count, bool pin) : n(count)
    {
        void *q = nullptr;
        if (pin && count && issl_host_alloc(count * sizeof(T), &q) == ISSL_OK) { p = static_cast<T *>(q); pinned = true; }
        else p = static_cast<T *>(calloc(count ? count : 1, sizeof(T)));
        if (pinned) memset(p, 0, count * sizeof(T));
    }
    ~HostArray() { if (pinned) issl_host_free(p); else free(p); }
    HostArray(const HostArray &) = delete;
    HostArray &operator=(const HostArray &) = delete;
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
};

// Starts bin/isslScoreServer (next to this executable) detached from the caller's terminal and pipes.
bool spawn_server(const char *sockPath)
{
    char self[PATH_MAX];
    const ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
    if (n <= 0) return false;
    self[n] = 0;
    std::string exe(self);
    exe = exe.substr(0, exe.rfind('/') + 1) + "isslScoreServer";
    if (access(exe.c_str(), X_OK) != 0) return false;
    const pid_t pid = fork();
    if (pid < 0) return false;
    if (pid == 0) {
        setsid();
        if (fork() != 0) _exit(0);   // grandchild is re-parented to init: nobody has to wait for it
        const int nul = open("/dev/null", O_RDWR);
        const char *logPath = getenv("ISSL_SERVER_LOG");
        const int log = logPath ? open(logPath, O_WRONLY | O_CREAT | O_APPEND, 0600) : -1;
        dup2(nul, 0); dup2(nul, 1); dup2(log >= 0 ? log : nul, 2);   // never hold the caller's stdout redirect open
        for (int fd = 3; fd < 1024; fd++) close(fd);
        execl(exe.c_str(), exe.c_str(), sockPath, (char *)nullptr);
        _exit(127);
    }
    int status;
    waitpid(pid, &status, 0);
    return true;
}

// Scores through the resident server.  Returns 0 on success, 1 when the server answered with an error
// (message already on stderr), -1 when no server could be reached (the caller then scores in-process).
int score_remote(const char *sockPath, const char *indexPath, const HostArray<uint64_t> &guides, int maxDist, double threshold,
                 int method, HostArray<double> &mit, HostArray<double> &cfd, bool timing)
{
    using namespace issl_wire;
    int fd = connect_to(sockPath);
    if (fd < 0 && getenv("ISSL_SERVER_AUTOSTART") && atoi(getenv("ISSL_SERVER_AUTOSTART")) != 0 && spawn_server(sockPath)) {
        for (int tries = 0; tries < 600 && fd < 0; tries++) {   // the server's first cudaInit can take a while
            usleep(50 * 1000);
            fd = connect_to(sockPath);
        }
    }
    if (fd < 0) return -1;
    char real[PATH_MAX];
    if (!realpath(indexPath, real)) { close(fd); return -1; }
    Request req{};
    memcpy(req.magic, kReqMagic, 8);
    req.op = kScore; req.maxDist = maxDist; req.threshold = threshold; req.method = method;
    req.layout = issl_host::layout_from_env();
    req.nGuides = guides.size();
    req.pathLen = (uint32_t)strlen(real);
    std::vector<int> devs = issl_host::devices_from_env();
    if (devs.empty() && getenv("ISSL_GPUS")) devs = issl_host::pick_devices(guides.size());
    req.nDevices = (uint32_t)std::min<size_t>(devs.size(), kMaxDevices);
    for (uint32_t k = 0; k < req.nDevices; k++) req.devices[k] = devs[k];
    Response rsp{};
    std::string msg;
    bool ok = write_full(fd, &req, sizeof req) && write_full(fd, real, req.pathLen) &&
              (guides.empty() || write_full(fd, guides.data(), guides.size() * 8)) && read_full(fd, &rsp, sizeof rsp) &&
              memcmp(rsp.magic, kRspMagic, 8) == 0;
    if (ok && rsp.msgLen) { msg.resize(rsp.msgLen); ok = read_full(fd, msg.data(), rsp.msgLen); }
    if (ok && rsp.status == ISSL_OK) {
        ok = rsp.n == guides.size() && (rsp.n == 0 || (read_full(fd, mit.data(), rsp.n * 8) && read_full(fd, cfd.data(), rsp.n * 8)));
    }
    close(fd);
    if (!ok) { fprintf(stderr, "isslScoreOfftargets: the score server on %s hung up\n", sockPath); return 1; }
    if (rsp.status != ISSL_OK) { fprintf(stderr, "%s\n", msg.c_str()); return 1; }
    if (timing)
        fprintf(stderr, "[issl] server: %u GPU(s), index %s (%.3f s), scoring %.3f s, candidates %llu hits %llu early-exits %llu\n",
                rsp.nDevices, rsp.cached ? "resident" : "loaded", rsp.loadSeconds, rsp.scoreSeconds,
                (unsigned long long)rsp.candidates, (unsigned long long)rsp.hits, (unsigned long long)rsp.earlyExits);
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    // The reference guards with argc < 4 but reads argv[4] and argv[5] unconditionally
    // (isslScoreOfftargets.cpp:93-96, :112, :121); all five arguments are required here.
    if (argc < 6) {
        fprintf(stderr, "Usage: %s [issltable] [query file] [max distance] [score-threshold] [score-method]\n", argv[0]);
        return 1;
    }
    const bool timing = getenv("ISSL_TIMING") && atoi(getenv("ISSL_TIMING")) != 0;
    const double t0 = now_s();

    const int maxDist = atoi(argv[3]);          // ref :109
    const double threshold = atof(argv[4]);     // ref :112
    const int method = issl_method_from_string(argv[5]);   // ref :121-143
    const bool calcMit = method == ISSL_METHOD_MIT || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    const bool calcCfd = method == ISSL_METHOD_CFD || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;

    issl_index *index = nullptr;
    if (issl_index_open(argv[1], &index) != ISSL_OK) {
        fprintf(stderr, "%s\n", issl_last_error());
        return 1;
    }
    issl_info info;
    issl_index_info(index, &info);

    // guide file: ref :275-294
    const size_t seqLineLength = info.seqLength + 1;
    struct stat st;
    if (stat(argv[2], &st) != 0) {
        fprintf(stderr, "Failed to read in query file.\n");
        return 1;
    }
    const size_t fileSize = (size_t)st.st_size;
    if (fileSize % seqLineLength != 0) {
        fprintf(stderr, "Error: query file is not a multiple of the expected line length (%zu)\n", seqLineLength);
        fprintf(stderr, "The sequence length may be incorrect; alternatively, the line endings\n");
        fprintf(stderr, "may be something other than LF, or there may be junk at the end of the file.\n");
        return 1;
    }
    const size_t queryCount = fileSize / seqLineLength;
    std::vector<char> queryDataSet(fileSize);
    FILE *fp = fopen(argv[2], "rb");
    if (!fp || fileSize == 0 || fread(queryDataSet.data(), fileSize, 1, fp) < 1) {
        fprintf(stderr, "Failed to read in query file.\n");
        return 1;
    }
    fclose(fp);

    // pinned memory needs a CUDA context; a score server client never creates one
    const bool pin = queryCount >= 65536 && !(getenv("ISSL_SERVER") && *getenv("ISSL_SERVER")) && issl_device_count() > 0;
    HostArray<uint64_t> querySignatures(queryCount, pin);
    if (issl_pack_guides(queryDataSet.data(), fileSize, info.seqLength, querySignatures.data()) != ISSL_OK) {
        fprintf(stderr, "%s\n", issl_last_error());
        return 1;
    }
    HostArray<double> mit(queryCount, pin), cfd(queryCount, pin);
    const double t1 = now_s();

    // index replicated per GPU, guides partitioned into contiguous ranges, no cross-GPU reduction
    double tLoad = 0, tFan = 0, tScore = 0;
    size_t nGpus = 0;
    if (calcMit || calcCfd) {
        int remote = -1;
        const char *server = getenv("ISSL_SERVER");
        if (server && *server) {
            remote = score_remote(server, argv[1], querySignatures, maxDist, threshold, method, mit, cfd, timing);
            if (remote == 1) return 1;
            if (remote < 0) fprintf(stderr, "isslScoreOfftargets: no score server on %s, scoring in-process\n", server);
        }
        if (remote != 0) {
            const std::vector<int> devs = issl_host::pick_devices(queryCount);
            issl_host::DeviceSet set;
            std::string err;
            double sec[2] = {0, 0};
            if (set.ensure(index, devs, issl_host::layout_from_env(), &err, sec) != ISSL_OK) {
                fprintf(stderr, "%s\n", err.c_str());
                return 1;
            }
            const double a1 = now_s();
            nGpus = devs.size();
            if (set.score(devs, querySignatures.data(), queryCount, maxDist, threshold, method, mit.data(), cfd.data(), nullptr, &err, timing) != ISSL_OK) {
                fprintf(stderr, "%s\n", err.c_str());
                return 1;
            }
            tLoad = sec[0]; tFan = sec[1]; tScore = now_s() - a1;
        }
    }
    const double t2 = now_s();

    // ref :514-527: "%s\t" then "%f\t" / "-1\t" then "%f\n" / "-1\n", in input order.  issl_format_lines formats with all
    // cores (the digits of printf's %f); a million lines at a time bounds the buffer.
    {
        const size_t L = info.seqLength, step = 1u << 20;
        std::vector<char> text;
        for (size_t b = 0; b < queryCount; b += step) {
            const size_t n = std::min(step, queryCount - b);
            text.resize(n * (L + 2 + 2 * 24));
            size_t bytes = issl_format_lines(querySignatures.data() + b, mit.data() + b, cfd.data() + b, n, L, method, text.data(), text.size());
            if (bytes > text.size()) {   // scores with very many digits
                text.resize(bytes);
                bytes = issl_format_lines(querySignatures.data() + b, mit.data() + b, cfd.data() + b, n, L, method, text.data(), text.size());
            }
            if (fwrite(text.data(), 1, bytes, stdout) != bytes) {
                fprintf(stderr, "Failed to write results.\n");
                return 1;
            }
        }
        fflush(stdout);
    }
    const double t3 = now_s();
    issl_index_close(index);
    if (timing)
        fprintf(stderr, "[issl] parse+guides %.3f s, index to HBM %.3f s, replicas on %zu more gpu(s) %.3f s, scoring %.3f s, print %.3f s, "
                        "index close %.3f s, total %.3f s\n",
                t1 - t0, tLoad, nGpus ? nGpus - 1 : 0, tFan, tScore, t3 - t2, now_s() - t3, now_s() - t0);
    return 0;
}
