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
int score_remote(const char *sockPath, const char *indexPath, const std::vector<uint64_t> &guides, int maxDist, double threshold,
                 int method, std::vector<double> &mit, std::vector<double> &cfd, bool timing)
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

    std::vector<uint64_t> querySignatures(queryCount);
    if (issl_pack_guides(queryDataSet.data(), fileSize, info.seqLength, querySignatures.data()) != ISSL_OK) {
        fprintf(stderr, "%s\n", issl_last_error());
        return 1;
    }
    std::vector<double> mit(queryCount, 0.0), cfd(queryCount, 0.0);
    const double t1 = now_s();

    // index replicated per GPU, guides partitioned into contiguous ranges, no cross-GPU reduction
    double tLoad = 0, tScore = 0;
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
            const double a0 = now_s();
            if (set.ensure(index, devs, issl_host::layout_from_env(), &err) != ISSL_OK) {
                fprintf(stderr, "%s\n", err.c_str());
                return 1;
            }
            const double a1 = now_s();
            if (set.score(devs, querySignatures.data(), queryCount, maxDist, threshold, method, mit.data(), cfd.data(), nullptr, &err, timing) != ISSL_OK) {
                fprintf(stderr, "%s\n", err.c_str());
                return 1;
            }
            tLoad = a1 - a0; tScore = now_s() - a1;
        }
    }
    const double t2 = now_s();

    // ref :514-527: "%s\t" then "%f\t" / "-1\t" then "%f\n" / "-1\n", in input order.  Lines are formatted in
    // parallel chunks (glibc's %f is the slow part at millions of guides) and written sequentially.
    {
        const size_t L = info.seqLength;
        const size_t chunkLines = 1 << 11;
        const size_t nChunks = (queryCount + chunkLines - 1) / chunkLines;
        std::vector<std::string> chunks(nChunks);
#pragma omp parallel for schedule(dynamic, 1)
        for (long c = 0; c < (long)nChunks; c++) {
            std::string &o = chunks[c];
            const size_t b = (size_t)c * chunkLines, e = std::min(queryCount, b + chunkLines);
            o.reserve((e - b) * (L + 48));
            char num[512];
            std::vector<char> seq(L);
            for (size_t i = b; i < e; i++) {
                issl_unpack_guide(querySignatures[i], L, seq.data());
                o.append(seq.data(), L);
                o.push_back('\t');
                if (calcMit) o.append(num, (size_t)snprintf(num, sizeof num, "%f\t", mit[i])); else o.append("-1\t");
                if (calcCfd) o.append(num, (size_t)snprintf(num, sizeof num, "%f\n", cfd[i])); else o.append("-1\n");
            }
        }
        for (const std::string &o : chunks)
            if (!o.empty() && fwrite(o.data(), 1, o.size(), stdout) != o.size()) {
                fprintf(stderr, "Failed to write results.\n");
                return 1;
            }
        fflush(stdout);
    }
    issl_index_close(index);
    if (timing)
        fprintf(stderr, "[issl] parse+guides %.3f s, index to HBM %.3f s, scoring %.3f s, print %.3f s, total %.3f s\n",
                t1 - t0, tLoad, tScore, now_s() - t2, now_s() - t0);
    return 0;
}
