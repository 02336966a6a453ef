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
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "issl_cuda.h"

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::vector<int> pick_devices(size_t nGuides)
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
        if (!devs.empty()) return devs;
    }
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

int layout_from_env()
{
    const char *e = getenv("ISSL_LAYOUT");
    if (!e) return ISSL_LAYOUT_AUTO;
    if (!strcmp(e, "res32")) return ISSL_LAYOUT_RES32;
    if (!strcmp(e, "sig64")) return ISSL_LAYOUT_SIG64;
    if (!strcmp(e, "gather")) return ISSL_LAYOUT_GATHER;
    return ISSL_LAYOUT_AUTO;
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
        const std::vector<int> devs = pick_devices(queryCount);
        const size_t nd = devs.size();
        std::vector<std::string> errors(nd);
        std::vector<double> loadS(nd, 0), scoreS(nd, 0);
        const int layout = layout_from_env();
        auto worker = [&](size_t k) {
            const size_t b = queryCount * k / nd, e = queryCount * (k + 1) / nd;
            const double a0 = now_s();
            issl_device *dev = nullptr;
            if (issl_device_create(index, devs[k], layout, &dev) != ISSL_OK) { errors[k] = issl_last_error(); return; }
            const double a1 = now_s();
            if (issl_score(dev, querySignatures.data() + b, e - b, maxDist, threshold, method, mit.data() + b, cfd.data() + b) != ISSL_OK)
                errors[k] = issl_last_error();
            const double a2 = now_s();
            loadS[k] = a1 - a0; scoreS[k] = a2 - a1;
            if (timing) {
                issl_stats s;
                issl_last_stats(dev, &s);
                fprintf(stderr, "[issl] gpu %d: guides %zu candidates %llu hits %llu early-exits %llu scan %.3f ms device-total %.3f ms\n",
                        devs[k], e - b, (unsigned long long)s.candidates, (unsigned long long)s.hits,
                        (unsigned long long)s.early_exits, s.scan_ms, s.total_ms);
            }
            issl_device_destroy(dev);
        };
        if (nd == 1) worker(0);
        else {
            std::vector<std::thread> pool;
            for (size_t k = 0; k < nd; k++) pool.emplace_back(worker, k);
            for (auto &t : pool) t.join();
        }
        for (size_t k = 0; k < nd; k++) {
            if (!errors[k].empty()) {
                fprintf(stderr, "%s\n", errors[k].c_str());
                return 1;
            }
            tLoad = loadS[k] > tLoad ? loadS[k] : tLoad;
            tScore = scoreS[k] > tScore ? scoreS[k] : tScore;
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
