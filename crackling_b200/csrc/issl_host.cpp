// issl_host.cpp -- host-side half of libissl_cuda: .issl parsing/validation, guide packing,
// method names, the builder-side MIT score arithmetic, error strings.  No CUDA in this file.
//
// "ref:" = /root/reference/src/ISSL/.

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <numeric>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "issl_internal.h"
#include "issl_triple_tables.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_error[512] = "";

int issl_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *issl_last_error(void) { return g_error; }
extern "C" int issl_abi_version(void) { return ISSL_CUDA_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// .issl image
// ---------------------------------------------------------------------------------------------
// Section order and sizes: ref isslCreateIndex.cpp:256-289; error texts: ref
// isslScoreOfftargets.cpp:164-167, :201-204, :223-226, :237-240.
static int parse_image(issl_index *ix)
{
    const size_t words = ix->bytes / 8;
    if (ix->bytes < 6 * sizeof(uint64_t))
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: header invalid");
    const uint64_t *h = reinterpret_cast<const uint64_t *>(ix->base);
    issl_info &f = ix->info;
    f.offtargetsCount = h[0]; f.seqLength = h[1]; f.seqCount = h[2];
    f.sliceWidth = h[3];      f.sliceCount = h[4]; f.scoresCount = h[5];

    if (f.seqLength == 0 || f.seqLength > 32)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: header invalid (sequence length %llu)",
                              (unsigned long long)f.seqLength);
    if (f.sliceWidth == 0 || f.sliceWidth > 24 || f.sliceCount == 0 || f.sliceCount * f.sliceWidth > 64)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: header invalid (slice width %llu x count %llu)",
                              (unsigned long long)f.sliceWidth, (unsigned long long)f.sliceCount);
    if (f.offtargetsCount >= (1ull << 32))   // ids are 32-bit in the file, isslCreateIndex.cpp:225-230
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: more than 2^32 off-target sites");
    ix->sliceLimit = 1ull << f.sliceWidth;

    uint64_t off = 6;
    if (f.scoresCount > words || off + 2 * f.scoresCount > words)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: header invalid (score table exceeds file)");
    ix->scorePairs = h + off;
    ix->scoresInFile = f.scoresCount;
    off += 2 * f.scoresCount;

    if (f.offtargetsCount == 0 || off + f.offtargetsCount > words)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: loading off-target sequences failed");
    ix->offtargets = h + off;
    off += f.offtargetsCount;

    const uint64_t nLists = f.sliceCount * ix->sliceLimit;
    if (off + nLists > words)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: reading slice list sizes failed");
    ix->sizes = h + off;
    off += nLists;

    // isslCreateIndex.cpp:225-233 puts every site into exactly one list of every slice.
    for (uint64_t i = 0; i < f.sliceCount; i++) {
        uint64_t sum = 0;
        for (uint64_t v = 0; v < ix->sliceLimit; v++) {
            const uint64_t s = ix->sizes[i * ix->sliceLimit + v];
            if (s > f.offtargetsCount)
                return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: reading slice list sizes failed (list longer than the index)");
            sum += s;
        }
        if (sum != f.offtargetsCount)
            return issl_set_error(ISSL_ERR_UNSUPPORTED,
                                  "Error reading index: slice %llu lists hold %llu entries, expected one per off-target (%llu)",
                                  (unsigned long long)i, (unsigned long long)sum, (unsigned long long)f.offtargetsCount);
    }
    if (off + f.sliceCount * f.offtargetsCount > words)
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: reading slice contents failed");
    ix->entries = h + off;
    return ISSL_OK;
}

extern "C" int issl_index_open(const char *path, issl_index **out)
{
    if (!path || !out) return issl_set_error(ISSL_ERR_ARG, "issl_index_open: null argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return issl_set_error(ISSL_ERR_IO, "Error reading index: cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        close(fd);
        return issl_set_error(ISSL_ERR_IO, "Error reading index: cannot stat %s", path);
    }
    if (st.st_size < 48) {
        close(fd);
        return issl_set_error(ISSL_ERR_FORMAT, "Error reading index: header invalid");
    }
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return issl_set_error(ISSL_ERR_IO, "Error reading index: mmap of %s failed: %s", path, strerror(errno));
    madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
    issl_index *ix = new issl_index();
    ix->base = static_cast<const uint8_t *>(p);
    ix->bytes = (size_t)st.st_size;
    ix->mapped = true;
    const int rc = parse_image(ix);
    if (rc != ISSL_OK) { issl_index_close(ix); return rc; }
    *out = ix;
    return ISSL_OK;
}

extern "C" int issl_index_from_memory(const void *image, size_t bytes, issl_index **out)
{
    if (!image || !out) return issl_set_error(ISSL_ERR_ARG, "issl_index_from_memory: null argument");
    *out = nullptr;
    if (reinterpret_cast<uintptr_t>(image) % 8 != 0)
        return issl_set_error(ISSL_ERR_ARG, "issl_index_from_memory: image must be 8-byte aligned");
    issl_index *ix = new issl_index();
    ix->base = static_cast<const uint8_t *>(image);
    ix->bytes = bytes;
    ix->mapped = false;
    const int rc = parse_image(ix);
    if (rc != ISSL_OK) { delete ix; return rc; }
    *out = ix;
    return ISSL_OK;
}

extern "C" int issl_index_info(const issl_index *index, issl_info *out)
{
    if (!index || !out) return issl_set_error(ISSL_ERR_ARG, "issl_index_info: null argument");
    *out = index->info;
    return ISSL_OK;
}

extern "C" void issl_index_close(issl_index *index)
{
    if (!index) return;
    if (index->mapped && index->base) munmap(const_cast<uint8_t *>(index->base), index->bytes);
    delete index;
}

// ---------------------------------------------------------------------------------------------
// guides
// ---------------------------------------------------------------------------------------------
extern "C" int issl_pack_guides(const char *text, size_t bytes, size_t seqLength, uint64_t *out)
{
    if (seqLength == 0 || seqLength > 32) return issl_set_error(ISSL_ERR_ARG, "issl_pack_guides: bad sequence length");
    const size_t line = seqLength + 1;
    if (bytes % line != 0)   // ref :277-282
        return issl_set_error(ISSL_ERR_ARG, "Error: query file is not a multiple of the expected line length (%zu)", line);
    if (bytes && (!text || !out)) return issl_set_error(ISSL_ERR_ARG, "issl_pack_guides: null argument");
    uint8_t code[256] = {0};          // ref :42, :99-102: everything but C, G, T packs as 0
    code[(unsigned char)'C'] = 1; code[(unsigned char)'G'] = 2; code[(unsigned char)'T'] = 3;
    const size_t n = bytes / line;
#pragma omp parallel for schedule(static) if (n > 65536)
    for (size_t i = 0; i < n; i++) {
        const unsigned char *p = reinterpret_cast<const unsigned char *>(text) + i * line;
        uint64_t s = 0;
        for (size_t j = 0; j < seqLength; j++) s |= (uint64_t)code[p[j]] << (2 * j);
        out[i] = s;
    }
    return ISSL_OK;
}

extern "C" void issl_unpack_guide(uint64_t signature, size_t seqLength, char *out)
{
    for (size_t j = 0; j < seqLength; j++) out[j] = "ACGT"[(signature >> (2 * j)) & 3];
}

// "%f" of a double: std::to_chars in fixed notation with six decimals is correctly rounded, as glibc's printf is, so the
// digits are the same (tests/test_lib_cpu.py compares them); it is several times faster.  Non-finite values keep printf.
static inline char *put_f(char *p, double v)
{
    if (!std::isfinite(v)) return p + sprintf(p, "%f", v);
    return std::to_chars(p, p + 400, v, std::chars_format::fixed, 6).ptr;
}

// ref isslScoreOfftargets.cpp:514-527: "%s\t" then "%f\t" or "-1\t" then "%f\n" or "-1\n", in input order
extern "C" size_t issl_format_lines(const uint64_t *guides, const double *mit, const double *cfd, size_t n, size_t seqLength,
                                    int method, char *out, size_t cap)
{
    const bool calcMit = method == ISSL_METHOD_MIT || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    const bool calcCfd = method == ISSL_METHOD_CFD || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    if (n == 0) return 0;
    if (!guides || (calcMit && !mit) || (calcCfd && !cfd)) { issl_set_error(ISSL_ERR_ARG, "issl_format_lines: null argument"); return 0; }
    // pass 1: every chunk of lines into its own buffer, in parallel; pass 2: lengths -> offsets -> out
    const size_t chunkLines = 4096, nChunks = (n + chunkLines - 1) / chunkLines;
    std::vector<std::string> chunks(nChunks);
#pragma omp parallel for schedule(dynamic, 4)
    for (long c = 0; c < (long)nChunks; c++) {
        const size_t b = (size_t)c * chunkLines, e = std::min(n, b + chunkLines);
        std::string &o = chunks[(size_t)c];
        o.reserve((e - b) * (seqLength + 28));
        char line[32 + 2 + 2 * 400];
        for (size_t i = b; i < e; i++) {
            char *p = line;
            issl_unpack_guide(guides[i], seqLength, p);
            p += seqLength;
            *p++ = '\t';
            if (calcMit) p = put_f(p, mit[i]); else { *p++ = '-'; *p++ = '1'; }
            *p++ = '\t';
            if (calcCfd) p = put_f(p, cfd[i]); else { *p++ = '-'; *p++ = '1'; }
            *p++ = '\n';
            o.append(line, (size_t)(p - line));
        }
    }
    std::vector<size_t> off(nChunks + 1, 0);
    for (size_t c = 0; c < nChunks; c++) off[c + 1] = off[c] + chunks[c].size();
    if (out && off[nChunks] <= cap) {
#pragma omp parallel for schedule(static)
        for (long c = 0; c < (long)nChunks; c++) memcpy(out + off[(size_t)c], chunks[(size_t)c].data(), chunks[(size_t)c].size());
    }
    return off[nChunks];
}

extern "C" int issl_method_from_string(const char *name)
{
    if (!name) return ISSL_METHOD_UNKNOWN;
    static const struct { const char *s; int m; } table[] = {
        {"mit", ISSL_METHOD_MIT}, {"cfd", ISSL_METHOD_CFD}, {"and", ISSL_METHOD_AND},
        {"or", ISSL_METHOD_OR},   {"avg", ISSL_METHOD_AVG}};
    for (const auto &e : table)
        if (strcmp(name, e.s) == 0) return e.m;
    return ISSL_METHOD_UNKNOWN;
}

// ---------------------------------------------------------------------------------------------
// score tables
// ---------------------------------------------------------------------------------------------
void issl_sorted_score_table(const uint64_t *pairs, uint64_t n, std::vector<uint64_t> &masks,
                             std::vector<double> &scores)
{
    std::vector<uint64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint64_t a, uint64_t b) { return pairs[2 * a] < pairs[2 * b]; });
    masks.clear(); scores.clear();
    masks.reserve(n); scores.reserve(n);
    for (uint64_t k : order) {
        if (!masks.empty() && masks.back() == pairs[2 * k]) continue;   // first insertion wins
        double s;
        memcpy(&s, &pairs[2 * k + 1], sizeof s);
        masks.push_back(pairs[2 * k]);
        scores.push_back(s);
    }
}

// Local MIT ("Hsu") score of one mismatch pattern.  ref isslCreateIndex.cpp:93-130:
//   T1 = prod(1 - M[pos]),  T2 = 1 / ((19 - meanGap)/19 * 4 + 1),  T3 = 1/m^2,  score = T1*T2*T3*100
// with meanGap = 19 for a single mismatch.  The factors are combined in the same order as the
// reference so the doubles come out bit-identical to the table stored in real .issl files.
extern "C" double issl_local_mit_score(uint64_t mask, size_t seqLength)
{
    static const double hsu[20] = {0.0,   0.0,   0.014, 0.0,   0.0,   0.395, 0.317, 0.0,   0.389, 0.079,
                                   0.445, 0.508, 0.613, 0.851, 0.732, 0.828, 0.615, 0.804, 0.685, 0.583};
    int first = -1, last = -1, m = 0;
    double t1 = 1.0;
    const size_t limit = seqLength < 32 ? seqLength : 32;
    for (size_t pos = 0; pos < limit; pos++) {
        if (((mask >> (2 * pos)) & 3) == 0) continue;
        if (pos >= 20) return 0.0;            // the reference only ever generates masks over 20 positions
        t1 = t1 * (1.0 - hsu[pos]);
        if (first < 0) first = (int)pos;
        last = (int)pos;
        m++;
    }
    if (m == 0) return 0.0;
    double gap = 19.0;
    if (m > 1) gap = (double)(last - first) / (double)(m - 1);   // the gaps telescope; small ints are exact
    const double t2 = 1.0 / ((19.0 - gap) / 19.0 * 4.0 + 1);
    const double t3 = 1.0 / (double)(m * m);
    return t1 * t2 * t3 * 100;
}

// ref isslCreateIndex.cpp:239-252: every mask with 1..sliceCount-1 of 20 positions set, stored in
// ascending mask order (std::map).  Enumerated here per popcount with Gosper's hack on the
// 20-bit position set, spread to bit 2*pos, then sorted.
extern "C" size_t issl_mit_table(size_t seqLength, size_t sliceWidth, uint64_t *masks, double *scores, size_t cap,
                                 uint64_t *scoresCount)
{
    if (scoresCount) *scoresCount = 0;
    if (sliceWidth == 0 || seqLength == 0 || seqLength > 32) return 0;
    const int maxMismatches = (int)(seqLength * 2 / sliceWidth) - 1;
    if (maxMismatches >= 20) { issl_set_error(ISSL_ERR_UNSUPPORTED, "issl_mit_table: slice width too small"); return 0; }
    std::vector<uint64_t> all;
    for (int m = 1; m <= maxMismatches; m++) {
        uint32_t set = (1u << m) - 1;
        while (set < (1u << 20)) {
            uint64_t spread = 0;
            for (int p = 0; p < 20; p++)
                if (set & (1u << p)) spread |= 1ull << (2 * p);
            all.push_back(spread);
            const uint32_t c = set & (0u - set), r = set + c;
            set = (((r ^ set) >> 2) / c) | r;
        }
    }
    std::sort(all.begin(), all.end());
    if (scoresCount) *scoresCount = all.size();
    size_t n = 0;
    for (uint64_t mk : all) {
        if (n >= cap) break;
        if (masks) masks[n] = mk;
        if (scores) scores[n] = issl_local_mit_score(mk, seqLength);
        n++;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// ISSL_LAYOUT_TRIPLE: which sub-buckets a guide has to read
// ---------------------------------------------------------------------------------------------
// The reference scores a site iff it lies in at least one of the guide's five looked-up lists -- i.e. the set E of
// slices on which site and guide agree exactly is non-empty -- and dist <= maxDist, and it meets the site first
// in slice min(E) (ref isslScoreOfftargets.cpp:330-390).  Every slice outside E carries at least one mismatch.
// Each such site is made the responsibility of exactly one slice triple T = resp(E) (issl_triple_tables.h), which
// contains E when |E| <= 3 and is the three lowest slices of E otherwise; within that triple's copy of the index it can
// only sit in a bucket whose key differs from the guide's on the slices of T \ E, by 1..(maxDist - #other non-exact
// slices) mismatches each.  Enumerating those XOR patterns gives the table below; the scan kernel re-derives E for
// every survivor and keeps it only when resp(E) is the triple it was found in, so every hit is produced exactly once.
namespace {
const uint8_t kLayout[10][5] = ISSL_TRIPLE_LAYOUT_INIT;

int byte_pos(int t, int slice)
{
    for (int k = 0; k < 3; k++)
        if (kLayout[t][k] == slice) return k;
    return -1;
}
int ham4(uint32_t x) { return ((x & 3u) != 0) + ((x & 12u) != 0) + ((x & 48u) != 0) + ((x & 192u) != 0); }
}  // namespace

extern "C" void issl_triple_layout(uint8_t slices_out[50], uint8_t resp_out[32])
{
    if (slices_out) memcpy(slices_out, kLayout, 50);
    if (resp_out)
        for (uint32_t E = 0; E < 32; E++) resp_out[E] = (uint8_t)issl_triple_resp(E);
}

static size_t triple_visits(int maxDist, int noExactByte, uint32_t *out, size_t cap, uint32_t waveStart[6]);

extern "C" size_t issl_triple_visits(int maxDist, uint32_t *out, size_t cap, uint32_t waveStart[6])
{
    return triple_visits(maxDist, 0, out, cap, waveStart);
}

extern "C" size_t issl_triple_visits_w4(int maxDist, uint32_t *out, size_t cap, uint32_t waveStart[6])
{
    return triple_visits(maxDist, 1, out, cap, waveStart);
}

static size_t triple_visits(int maxDist, int noExactByte, uint32_t *out, size_t cap, uint32_t waveStart[6])
{
    std::vector<std::pair<uint32_t, uint32_t>> v;   // (wave, entry)
    if (maxDist > 7) { issl_set_error(ISSL_ERR_ARG, "issl_triple_visits: maxDist %d > 7", maxDist); maxDist = -1; }
    const int D = maxDist;
    auto add = [&](int wave, int t, uint32_t pattern, int budget) {
        v.push_back({(uint32_t)wave, pattern | ((uint32_t)t << 24) | ((uint32_t)budget << 28)});
    };
    if (D >= 0) {
        for (int t = 0; t < 10; t++)                                                     // E contains the whole triple
            add(std::min({kLayout[t][0], kLayout[t][1], kLayout[t][2]}), t, 0, D);
        for (int i = 0; i < 5; i++)                                                      // |E| == 2
            for (int j = i + 1; j < 5; j++) {
                const int t = (int)issl_triple_resp((1u << i) | (1u << j));
                const int k = kLayout[t][0], pos = byte_pos(t, k);                       // the third slice: the low key byte
                for (uint32_t x = 1; x < 256; x++)
                    if (ham4(x) <= D - 2) add(i, t, x << (8 * pos), D - ham4(x));
            }
        for (int e = 0; e < 5; e++) {                                                    // |E| == 1
            const int t = (int)issl_triple_resp(1u << e);
            const int j = kLayout[t][0], k = kLayout[t][1], pj = byte_pos(t, j), pk = byte_pos(t, k);
            for (uint32_t xj = 1; xj < 256; xj++)
                for (uint32_t xk = 1; xk < 256; xk++)
                    if (ham4(xj) + ham4(xk) <= D - 2) add(e, t, (xj << (8 * pj)) | (xk << (8 * pk)), D - ham4(xj) - ham4(xk));
        }
        if (noExactByte && D >= 5) {
            // sliceWidth 4: a site may agree with the guide on a 2-base slice without agreeing on any whole byte -- every
            // byte then carries a mismatch (maxDist >= 5).  Those sites belong to triple 0: all three key bytes differ, and
            // so do both residual slices (at least two more mismatches).  Last wave: where the reference meets such a hit
            // is worked out from the site itself (order_slice).
            for (uint32_t x0 = 1; x0 < 256; x0++)
                for (uint32_t x1 = 1; x1 < 256; x1++) {
                    if (ham4(x0) + ham4(x1) > D - 3) continue;
                    for (uint32_t x2 = 1; x2 < 256; x2++)
                        if (ham4(x0) + ham4(x1) + ham4(x2) <= D - 2)
                            add(4, 0, x0 | (x1 << 8) | (x2 << 16), D - ham4(x0) - ham4(x1) - ham4(x2));
                }
        }
    }
    std::sort(v.begin(), v.end());
    if (waveStart) {
        for (int s = 0; s <= 5; s++) {
            size_t lo = 0;
            while (lo < v.size() && (int)v[lo].first < s) lo++;
            waveStart[s] = (uint32_t)lo;
        }
    }
    if (out)
        for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i].second;
    return v.size();
}
