// issl_device.cu -- device-side half of libissl_cuda: index residency (re-layout from an .issl
// image or on-device construction), and the scoring pipeline
//     scan items -> K1 scan -> radix sort of survivor keys -> K2a contributions -> K2b ordered
//     accumulation (with the reference's early exit) -> finalise.
// "ref:" = /root/reference/src/ISSL/.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "issl_internal.h"
#include "issl_device_common.cuh"
#include "issl_kernels.cuh"
#include "issl_triple.cuh"

using namespace issl;

struct issl_device {
    int dev = -1;
    issl_info info{};
    IndexView iv{};
    int layout = 0;
    uint64_t nLists = 0;
    uint64_t hbmBytes = 0;
    int pbits = 0;

    // index storage
    DBuf sig, occ, ids, res32, sig64, listStart, listLen, filePrefix, mitMasks, mitScores;
    uint32_t mitCount = 0;
    std::vector<uint64_t> hMitMasks;   // as loaded/generated (for write_issl)
    std::vector<double> hMitScores;
    std::vector<uint64_t> hListLen, hListStart, hFilePrefix;

    // scoring scratch
    cudaStream_t stream = nullptr;
    DBuf guides, totMit, totCfd, done, pairKeys, pairVals, pairKeysSorted, pairValsSorted, pairCounts, pairOffsets, items, keysA, keysB, sortTemp, scanTemp,
        contribMit, contribCfd, counters, outMit, outCfd, hitId, hitDist, hitOcc, scoredEnd, segBegin;
    unsigned long long *hCounters = nullptr;   // pinned: [0] total candidates, [1] hit count, [2] items, [3] done
    uint64_t hitCap = 0, hitCapAuto = 0, firstHitCap = 0;
    std::vector<cudaEvent_t> evPool;
    issl_stats stats{};
    uint32_t maxBatch = 1u << 20;
    uint32_t maxGroup = kBigGroup;   // ISSL_MAX_GROUP: 32 bit-sliced blocks + register groups (default), 8/4/2 register groups only, 1 no list reuse

    // ISSL_LAYOUT_TRIPLE
    DBuf tripleRes, tripleIds, tripleOffs, tripleBlk, visits, segOff, segCnt, segKeys, segSites, totMit2, totCfd2, done2;
    uint64_t segCap = 0;
    DBuf redo;                       // guides the warp-per-guide kernel left to the CTA-per-guide kernel
    bool listsResident = false;      // the slice lists (ids / res32 / sig64) exist: TRIPLE builds them on demand (ensure_lists)
    DBuf tripleBits;                 // TripleView::nonEmpty (small indexes)
    DBuf ovfBits;                    // non-flush scan: one bit per (CTA, visit) whose bucket overflows its block beyond the shared-memory list
    int tripleSmall = 1;             // ISSL_TRIPLE_SMALL=0: never use the warp-per-guide kernel
    DBuf heavyDesc, heavyFlat;       // k_heavy_finish: one descriptor per heavy guide; the second half of its sort's ping-pong
    DBuf heavyKeys;                  // sort keys of the guides with more hits than a CTA's record list holds (heavy_finish)
    uint64_t heavyCap = 0;
    int tripleHeavy = 1;             // ISSL_TRIPLE_HEAVY=0: such guides go through the general pipeline (device-wide sort) instead
    DBuf mitDense;                   // the score table spread over all 2^20 position sets (seqLength <= 20)
    TripleView tv{};
    int tripleMaxDist = 6;           // ISSL_TRIPLE_MAXDIST: larger maxDist takes the RES32 list scan
    int waves = 1;                   // ISSL_WAVES: how slices are cut into launches when there is an early exit (score_batch)
    int tripleFlush = -1;            // ISSL_TRIPLE_FLUSH: 1 / 0 force the scan variant that flushes full record lists; -1 automatic
    double lastHitsPerGuide = 0;     // of the previous scoring call on this handle
    double lastHeavyFraction = 0;    // share of its hits that belonged to guides with more hits than a CTA's record list holds
    double lastExitFraction = -1.0;  // early exits / guides of the previous call that had an early exit to take; -1: none yet
    int tripleFuse = 2;              // ISSL_TRIPLE_FUSE: 2 = guides are finished inside the scan kernel, 1 = by k_score_segments from
                                     // per-guide segments, 0 = everything through the general sort/score/accumulate kernels
    bool layoutAuto = false;         // TRIPLE was chosen by ISSL_LAYOUT_AUTO: fall back to RES32 if it does not fit
    int visitsDist = -100;           // maxDist the resident visit table was built for
    uint32_t waveStart[6] = {0, 0, 0, 0, 0, 0};
};

// ---------------------------------------------------------------------------------------------
// devices
// ---------------------------------------------------------------------------------------------
extern "C" int issl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int usable = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) usable++;
    }
    return usable;
}

static int select_device(int cuda_device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return issl_set_error(ISSL_ERR_NO_DEVICE, "no CUDA device available (%s); libissl_cuda has no CPU fallback",
                              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (cuda_device < 0 || cuda_device >= n)
        return issl_set_error(ISSL_ERR_NO_DEVICE, "CUDA device %d does not exist (%d present)", cuda_device, n);
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, cuda_device));
    if (p.major != 10)
        return issl_set_error(ISSL_ERR_NO_DEVICE, "CUDA device %d (%s, sm_%d%d) is not an sm_100 part; libissl_cuda is built for sm_100a only",
                              cuda_device, p.name, p.major, p.minor);
    CK(cudaSetDevice(cuda_device));
    return ISSL_OK;
}

static int upload_constants()
{
    CK(cudaMemcpyToSymbol(c_cfdPos, ISSL_CFD_POS, sizeof(double) * 320));
    CK(cudaMemcpyToSymbol(c_cfdPam, ISSL_CFD_PAM, sizeof(double) * 16));
    CK(cudaMemcpyToSymbol(g_cfdPos, ISSL_CFD_POS, sizeof(double) * 320));
    return ISSL_OK;
}

static int choose_layout(const issl_info &f, int requested, int *out)
{
    const uint32_t w = (uint32_t)f.sliceWidth, kb = std::min<uint32_t>(w, 8);
    const bool res32ok = (w % 2 == 0) && (2 * f.seqLength >= kb) && (2 * f.seqLength - kb <= 32);
    const bool tripleok = f.seqLength == 20 && ((w == 8 && f.sliceCount == 5) || (w == 4 && f.sliceCount == 10) || (w == 10 && f.sliceCount == 4));
    if (requested == ISSL_LAYOUT_AUTO) requested = tripleok ? ISSL_LAYOUT_TRIPLE : res32ok ? ISSL_LAYOUT_RES32 : ISSL_LAYOUT_SIG64;
    if (requested == ISSL_LAYOUT_RES32 && !res32ok)
        return issl_set_error(ISSL_ERR_UNSUPPORTED, "layout RES32 needs an even slice width and 2*seqLength - min(width,8) <= 32");
    if (requested == ISSL_LAYOUT_TRIPLE && !tripleok)
        return issl_set_error(ISSL_ERR_UNSUPPORTED, "layout TRIPLE needs seqLength 20 and sliceWidth 8, 4 or 10");
    if (requested != ISSL_LAYOUT_RES32 && requested != ISSL_LAYOUT_SIG64 && requested != ISSL_LAYOUT_GATHER &&
        requested != ISSL_LAYOUT_TRIPLE)
        return issl_set_error(ISSL_ERR_ARG, "unknown layout %d", requested);
    *out = requested;
    return ISSL_OK;
}

// common tail of both constructors: header-derived fields, list geometry, storage
static int alloc_lists(issl_device *d);

// TRIPLE scores from its sub-bucket copies; its slice lists (ids + residuals, 23 GB at human scale) are only needed for
// maxDist above ISSL_TRIPLE_MAXDIST, for issl_device_write_issl and when the copies do not fit: they are built from sig[] the
// first time something asks for them (ensure_lists; ISSL_LISTS_EAGER=1: at load, as in round 1)
static bool lists_are_lazy(int layout)
{
    if (layout != ISSL_LAYOUT_TRIPLE) return false;
    const char *e = getenv("ISSL_LISTS_EAGER");
    return !(e && atoi(e) != 0);
}

static int init_geometry(issl_device *d, const issl_info &f, int layout, const uint64_t *listLen /* host, nLists */)
{
    d->info = f;
    d->layout = layout;
    const uint64_t sliceLimit = 1ull << f.sliceWidth;
    d->nLists = f.sliceCount * sliceLimit;
    d->hListLen.assign(listLen, listLen + d->nLists);
    d->hListStart.resize(d->nLists);
    d->hFilePrefix.resize(d->nLists + 1);
    uint64_t p = 0, q = 0;
    for (uint64_t i = 0; i < d->nLists; i++) {
        d->hListStart[i] = p;
        d->hFilePrefix[i] = q;
        q += listLen[i];
        p += (listLen[i] + kListAlign - 1) / kListAlign * kListAlign;
    }
    d->hFilePrefix[d->nLists] = q;
    const uint64_t P = p + kListAlign;   // slack so that no vector load can leave the allocation
    d->pbits = 1;
    while ((1ull << d->pbits) < P) d->pbits++;
    if (d->pbits > 40) return issl_set_error(ISSL_ERR_UNSUPPORTED, "index too large: %llu list positions", (unsigned long long)P);

    const uint64_t N = f.offtargetsCount;
    CKR(d->sig.exact(N * 8));
    CKR(d->occ.exact(N * 4));
    // TRIPLE's slice lists (maxDist beyond what the sub-bucket scan serves, .issl export): RES32 for sliceWidth 8 and 10,
    // ids only (GATHER) for sliceWidth 4
    const bool res32 = layout == ISSL_LAYOUT_RES32 || (layout == ISSL_LAYOUT_TRIPLE && (f.sliceWidth == 8 || f.sliceWidth == 10));
    CKR(d->listStart.exact(d->nLists * 8));
    CKR(d->listLen.exact(d->nLists * 8));
    CKR(d->filePrefix.exact((d->nLists + 1) * 8));
    CK(cudaMemcpyAsync(d->listStart.p, d->hListStart.data(), d->nLists * 8, cudaMemcpyHostToDevice, d->stream));
    CK(cudaMemcpyAsync(d->listLen.p, d->hListLen.data(), d->nLists * 8, cudaMemcpyHostToDevice, d->stream));
    CK(cudaMemcpyAsync(d->filePrefix.p, d->hFilePrefix.data(), (d->nLists + 1) * 8, cudaMemcpyHostToDevice, d->stream));

    IndexView &v = d->iv;
    v.sig = d->sig.as<uint64_t>();
    v.occ = d->occ.as<uint32_t>();
    v.ids = nullptr; v.res32 = nullptr; v.sig64 = nullptr;   // alloc_lists
    v.listStart = d->listStart.as<uint64_t>();
    v.listLen = d->listLen.as<uint64_t>();
    v.N = N; v.P = P;
    v.seqLength = (uint32_t)f.seqLength; v.sliceWidth = (uint32_t)f.sliceWidth;
    v.sliceCount = (uint32_t)f.sliceCount; v.sliceLimit = (uint32_t)sliceLimit;
    v.sliceMask = (uint32_t)(sliceLimit - 1);
    v.knownBits = std::min<uint32_t>((uint32_t)f.sliceWidth, 8);
    v.layout = res32 ? (int)ISSL_LAYOUT_RES32 : layout == ISSL_LAYOUT_TRIPLE ? (int)ISSL_LAYOUT_GATHER : layout;

    d->hbmBytes = N * 12 + d->nLists * 24;
    d->listsResident = false;
    if (!lists_are_lazy(layout)) CKR(alloc_lists(d));
    return ISSL_OK;
}

// storage of the slice lists in position space; whoever calls this fills them (k_relayout / fill_lists)
static int alloc_lists(issl_device *d)
{
    IndexView &v = d->iv;
    const uint64_t P = v.P;
    CKR(d->ids.exact(P * 4));
    CK(cudaMemsetAsync(d->ids.p, 0xFF, P * 4, d->stream));
    if (v.layout == ISSL_LAYOUT_RES32) { CKR(d->res32.exact(P * 4)); CK(cudaMemsetAsync(d->res32.p, 0, P * 4, d->stream)); }
    if (v.layout == ISSL_LAYOUT_SIG64) { CKR(d->sig64.exact(P * 8)); CK(cudaMemsetAsync(d->sig64.p, 0, P * 8, d->stream)); }
    v.ids = d->ids.as<uint32_t>();
    v.res32 = d->res32.as<uint32_t>();
    v.sig64 = d->sig64.as<uint64_t>();
    d->hbmBytes += P * 4 + (v.layout == ISSL_LAYOUT_RES32 ? P * 4 : 0) + (v.layout == ISSL_LAYOUT_SIG64 ? P * 8 : 0);
    d->listsResident = true;
    return ISSL_OK;
}

// the slice lists from sig[]: per slice a stable 8-bit radix sort of (value, id) and placement -- "all sites whose slice i has
// value v, in id order" (ref isslCreateIndex.cpp:216-234, including the 8-bit truncation), which is exactly what a file that
// passed k_relayout's validation holds
static int fill_lists(issl_device *d)
{
    cudaStream_t st = d->stream;
    const uint64_t N = d->info.offtargetsCount, S = d->info.sliceCount, sliceLimit = 1ull << d->info.sliceWidth;
    const uint32_t w = (uint32_t)d->info.sliceWidth, smask = (uint32_t)(sliceLimit - 1);
    DBuf values, valuesSorted, idsTmp, idsSorted, hist, valueStart, tmp;
    CKR(values.ensure(N)); CKR(valuesSorted.ensure(N)); CKR(idsTmp.ensure(N * 4)); CKR(idsSorted.ensure(N * 4));
    CKR(hist.ensure(256 * 8)); CKR(valueStart.ensure(256 * 8));
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, values.as<uint8_t>(), valuesSorted.as<uint8_t>(), idsTmp.as<uint32_t>(),
                                       idsSorted.as<uint32_t>(), N, 0, 8, st));
    CKR(tmp.ensure(tb));
    for (uint64_t s = 0; s < S; s++) {
        CK(cudaMemsetAsync(hist.p, 0, 256 * 8, st));
        k_slice_values<<<148 * 8, 256, 0, st>>>(d->sig.as<uint64_t>(), N, w, smask, (uint32_t)s, values.as<uint8_t>(),
                                               idsTmp.as<uint32_t>(), hist.as<unsigned long long>());
        CK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, values.as<uint8_t>(), valuesSorted.as<uint8_t>(), idsTmp.as<uint32_t>(),
                                           idsSorted.as<uint32_t>(), N, 0, 8, st));
        uint64_t vs[256], acc = 0;
        for (uint64_t v = 0; v < 256; v++) { vs[v] = acc; acc += v < sliceLimit ? d->hListLen[s * sliceLimit + v] : 0; }
        CK(cudaMemcpyAsync(valueStart.p, vs, sizeof vs, cudaMemcpyHostToDevice, st));
        k_place_slice<<<blocks_for(N, 256), 256, 0, st>>>(d->iv, valuesSorted.as<uint8_t>(), idsSorted.as<uint32_t>(), N,
                                                         (uint32_t)s, valueStart.as<uint64_t>());
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));   // vs[] is a stack buffer
    }
    return ISSL_OK;
}

static int ensure_lists(issl_device *d)
{
    if (d->listsResident) return ISSL_OK;
    CKR(alloc_lists(d));
    const int rc = fill_lists(d);
    if (rc != ISSL_OK) {
        for (DBuf *b : {&d->ids, &d->res32, &d->sig64}) b->release();
        d->iv.ids = nullptr; d->iv.res32 = nullptr; d->iv.sig64 = nullptr; d->listsResident = false;
    }
    return rc;
}

static int upload_mit_table(issl_device *d)
{
    d->mitCount = (uint32_t)d->hMitMasks.size();
    CKR(d->mitMasks.exact(std::max<size_t>(1, d->mitCount) * 8));
    CKR(d->mitScores.exact(std::max<size_t>(1, d->mitCount) * 8));
    if (d->mitCount) {
        CK(cudaMemcpyAsync(d->mitMasks.p, d->hMitMasks.data(), d->mitCount * 8ull, cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->mitScores.p, d->hMitScores.data(), d->mitCount * 8ull, cudaMemcpyHostToDevice, d->stream));
    }
    d->hbmBytes += d->mitCount * 16ull;
    if (d->info.seqLength <= 20) {
        // dense copy: index = set of mismatching positions, 0.0 where the file has no entry (ref :394)
        std::vector<double> dense(1u << 20, 0.0);
        for (size_t k = 0; k < d->hMitMasks.size(); k++) {
            const uint64_t mk = d->hMitMasks[k];
            if ((mk & 0xAAAAAAAAAAAAAAAAull) || (mk >> 40)) continue;   // never looked up: masks only carry even bits below 2*seqLength
            uint32_t idx = 0;
            for (int p = 0; p < 20; p++) idx |= (uint32_t)((mk >> (2 * p)) & 1ull) << p;
            dense[idx] = d->hMitScores[k];
        }
        CKR(d->mitDense.exact(dense.size() * 8));
        CK(cudaMemcpyAsync(d->mitDense.p, dense.data(), dense.size() * 8, cudaMemcpyHostToDevice, d->stream));
        CK(cudaStreamSynchronize(d->stream));   // dense is a local
        d->hbmBytes += dense.size() * 8;
    }
    return ISSL_OK;
}

static ScoreTables score_tables(const issl_device *d)
{
    ScoreTables tb;
    tb.mitMasks = d->mitMasks.as<uint64_t>(); tb.mitScores = d->mitScores.as<double>(); tb.mitCount = d->mitCount;
    tb.mitDense = d->mitDense.as<double>();
    return tb;
}

// ---------------------------------------------------------------------------------------------
// ISSL_LAYOUT_TRIPLE: ten bucketed copies of the sites (issl_triple.cuh), built from sig[] once the
// index itself is resident and validated
// ---------------------------------------------------------------------------------------------
// Split in two so that issl_device_create can run it on a second stream while the slice lists are still crossing PCIe:
// the copies need only sig[] and occ[], which are complete once slice 0 has been re-laid out.
struct TripleBuild {
    DBuf keysIn, keysOut, idsIn, tmp;
    uint32_t pitch = 0, perm10 = 0;
    uint64_t stride = 0, needBlk = 0;
    bool occFlag = false, bitmap = false;
};

// allocations and launches on `st`; nothing is waited for.  ISSL_ERR_NOMEM when the copies do not fit.
static int build_triple_enqueue(issl_device *d, cudaStream_t st, TripleBuild &b)
{
    if (getenv("ISSL_TEST_TRIPLE_NOMEM"))   // test hook: behave as if the copies did not fit
        return issl_set_error(ISSL_ERR_NOMEM, "cudaMalloc: out of memory (simulated by ISSL_TEST_TRIPLE_NOMEM)");
    const uint64_t N = d->info.offtargetsCount;
    const uint64_t stride = (N + 64 + 7) / 8 * 8;
    // blocked copy of the residuals (issl_triple.cuh): block size from the mean bucket occupancy, so that all but
    // a few buckets fit; ISSL_TRIPLE_BLOCKS = 0 (off), 32 / 64 / 128 (forced), unset = automatic
    uint32_t pitch = 0;
    {
        const double lambda = (double)N / kTripleBuckets, want = lambda + 4.5 * std::sqrt(lambda);
        // from one site per five buckets on (3.4 M sites) a visit is one aligned read of the bucket's block; 64-byte blocks up to
        // ~14 sites per bucket.  Measured per 100 000 guides, 5 / 10 / 20 M sites: 2.20 ms each, against 3.42 / 5.70 / 6.04 ms
        // through the offsets (profiles/r02_ab_midsize_blocks.jsonl); at 2 M sites the bitmap below is as fast (2.11 vs 2.19)
        if (lambda >= 0.2) pitch = want <= 31 ? 32 : want <= 62 ? 64 : want <= 124 ? 128 : 0;
        if (const char *e = getenv("ISSL_TRIPLE_BLOCKS")) {
            const long v = atol(e);
            if (v == 0 || v == 32 || v == 64 || v == 128) pitch = (uint32_t)v;
        }
    }
    size_t freeB = 0, totalB = 0;
    CK(cudaMemGetInfo(&freeB, &totalB));
    const uint64_t needBase = kTripleCount * (stride * 6 + (kTripleBuckets + 1ull) * 4) + N * 16 + (64ull << 20);
    uint64_t needBlk = (uint64_t)kTripleCount * kTripleBuckets * pitch * 2;
    if (needBase + needBlk > freeB) { pitch = 0; needBlk = 0; }
    // bit 31 of the stored ids doubles as "occurs more than once" when ids leave it free (ISSL_TRIPLE_OCCFLAG=0: tests
    // of the path indexes with 2^31 sites or more take)
    bool occFlag = N < (1ull << 31);
    if (const char *e = getenv("ISSL_TRIPLE_OCCFLAG")) occFlag = occFlag && atoi(e) != 0;
    CKR(d->tripleRes.exact(kTripleCount * stride * 2));
    CKR(d->tripleIds.exact(kTripleCount * stride * 4));
    CKR(d->tripleOffs.exact(kTripleCount * (kTripleBuckets + 1ull) * 4));
    CK(cudaMemsetAsync(d->tripleRes.p, 0, kTripleCount * stride * 2, st));
    if (pitch) {
        CKR(d->tripleBlk.exact(needBlk));   // every sub-block is written by k_triple_blocks
    }
    // no blocked copy and mostly empty buckets (a bacterial genome): a bitmap of the buckets that hold anything
    b.bitmap = pitch == 0 && (double)N / kTripleBuckets < 0.5 && !getenv("ISSL_TRIPLE_NO_BITMAP");
    if (b.bitmap) CKR(d->tripleBits.exact((size_t)kTripleCount * (kTripleBuckets / 32) * 4));
    const uint32_t perm10 = d->info.sliceWidth == 10 ? 1u : 0u;   // sliceWidth 10: copies of the permuted signatures (issl_triple.cuh)
    DBuf &keysIn = b.keysIn, &keysOut = b.keysOut, &idsIn = b.idsIn, &tmp = b.tmp;
    CKR(keysIn.ensure(N * 4)); CKR(keysOut.ensure(N * 4)); CKR(idsIn.ensure(N * 4));
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keysIn.as<uint32_t>(), keysOut.as<uint32_t>(), idsIn.as<uint32_t>(),
                                       d->tripleIds.as<uint32_t>(), N, 0, 25, st));
    CKR(tmp.ensure(tb));
    CKR(d->counters.ensure(16 * 8));
    for (uint32_t t = 0; t < kTripleCount; t++) {
        uint32_t *ids = d->tripleIds.as<uint32_t>() + t * stride;
        k_triple_keys<<<blocks_for(N, 256), 256, 0, st>>>(d->sig.as<uint64_t>(), d->occ.as<uint32_t>(), N, t, perm10, keysIn.as<uint32_t>(), idsIn.as<uint32_t>());
        // stable: inside a bucket the sites that occur more than once come first, ids ascending within either class
        CK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keysIn.as<uint32_t>(), keysOut.as<uint32_t>(), idsIn.as<uint32_t>(), ids, N, 0, 25, st));
        k_triple_residuals<<<blocks_for(N, 256), 256, 0, st>>>(d->sig.as<uint64_t>(), ids, N, t, perm10, d->tripleRes.as<uint16_t>() + t * stride);
        if (occFlag) k_triple_flag_ids<<<blocks_for(N, 256), 256, 0, st>>>(d->occ.as<uint32_t>(), N, ids);
        k_triple_offsets<<<blocks_for(kTripleBuckets + 1ull, 256), 256, 0, st>>>(keysOut.as<uint32_t>(), N,
                                                                                d->tripleOffs.as<uint32_t>() + t * (kTripleBuckets + 1ull));
        if (b.bitmap)
            k_triple_nonempty<<<kTripleBuckets / 256, 256, 0, st>>>(d->tripleOffs.as<uint32_t>() + t * (kTripleBuckets + 1ull),
                                                                   d->tripleBits.as<uint32_t>() + (size_t)t * (kTripleBuckets / 32));
        if (pitch)
            k_triple_blocks<<<blocks_for((uint64_t)kTripleBuckets * (pitch / 32), 256), 256, 0, st>>>(
                d->tripleRes.as<uint16_t>() + t * stride, d->tripleOffs.as<uint32_t>() + t * (kTripleBuckets + 1ull), keysOut.as<uint32_t>(), pitch / 32,
                d->tripleBlk.as<uint4>() + (uint64_t)t * kTripleBuckets * (pitch / 8));
        CK(cudaGetLastError());
    }
    // are the sites in text order, i.e. are ids text ranks?  (read back by build_triple_finish)
    CK(cudaMemsetAsync(d->counters.as<unsigned long long>() + 15, 0, 8, st));
    k_check_site_order<<<blocks_for(N, 256), 256, 0, st>>>(d->sig.as<uint64_t>(), N, (uint32_t)d->info.seqLength, d->counters.as<unsigned long long>() + 15);
    CK(cudaMemcpyAsync(d->hCounters + 15, d->counters.as<unsigned long long>() + 15, 8, cudaMemcpyDeviceToHost, st));
    b.pitch = pitch; b.perm10 = perm10; b.stride = stride; b.needBlk = needBlk; b.occFlag = occFlag;
    return ISSL_OK;
}

static int build_triple_finish(issl_device *d, cudaStream_t st, TripleBuild &b)
{
    CK(cudaStreamSynchronize(st));
    bool siteOrdered = d->hCounters[15] == 0;   // (ISSL_SITE_ORDER=0: behave as if the sites were not in text order)
    if (const char *e = getenv("ISSL_SITE_ORDER")) siteOrdered = siteOrdered && atoi(e) != 0;
    for (DBuf *x : {&b.keysIn, &b.keysOut, &b.idsIn, &b.tmp}) x->release();
    d->tv.siteOrdered = siteOrdered ? 1u : 0u;
    d->tv.res = d->tripleRes.as<uint16_t>();
    d->tv.ids = d->tripleIds.as<uint32_t>();
    d->tv.offs = d->tripleOffs.as<uint32_t>();
    d->tv.stride = b.stride;
    d->tv.occFlag = b.occFlag ? 1u : 0u;
    d->tv.nibbleOrder = d->info.sliceWidth == 4 ? 1u : 0u;
    d->tv.perm10 = b.perm10;
    d->tv.blk = b.pitch ? d->tripleBlk.as<uint4>() : nullptr;
    d->tv.pitch = b.pitch;
    d->tv.nonEmpty = b.bitmap ? d->tripleBits.as<uint32_t>() : nullptr;
    d->hbmBytes += kTripleCount * (b.stride * 6 + (kTripleBuckets + 1ull) * 4) + b.needBlk + (b.bitmap ? (uint64_t)kTripleCount * (kTripleBuckets / 8) : 0);
    return ISSL_OK;
}

static int build_triple(issl_device *d)
{
    if (d->layout != ISSL_LAYOUT_TRIPLE) return ISSL_OK;
    TripleBuild b;
    CKR(build_triple_enqueue(d, d->stream, b));
    return build_triple_finish(d, d->stream, b);
}

// ISSL_LAYOUT_AUTO promised TRIPLE only "when it fits": when the ten copies cannot be allocated (a second human-scale
// index next to a resident one, a smaller GPU), the index stays usable through its slice lists.
static int triple_fall_back(issl_device *d, int rc)
{
    if (rc != ISSL_ERR_NOMEM || !d->layoutAuto || d->layout != ISSL_LAYOUT_TRIPLE) return rc;
    cudaGetLastError();
    for (DBuf *b : {&d->tripleRes, &d->tripleIds, &d->tripleOffs, &d->tripleBlk, &d->tripleBits}) b->release();
    d->tv = TripleView{};
    d->layout = d->iv.layout;   // RES32 (sliceWidth 8) or the ids-only lists (sliceWidth 4)
    CKR(ensure_lists(d));
    if (getenv("ISSL_DEBUG")) fprintf(stderr, "[issl] the sub-bucket copies do not fit this GPU's free memory: scanning the slice lists instead\n");
    return ISSL_OK;
}

static int build_triple_or_fall_back(issl_device *d)
{
    return triple_fall_back(d, build_triple(d));
}

static int new_device(int cuda_device, issl_device **out)
{
    CKR(select_device(cuda_device));
    issl_device *d = new issl_device();
    d->dev = cuda_device;
    if (const char *e = getenv("ISSL_BATCH")) {
        const long v = atol(e);
        if (v > 0 && v <= (1 << 24)) d->maxBatch = (uint32_t)v;
    }
    if (const char *e = getenv("ISSL_MAX_GROUP")) {
        const long v = atol(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 32) d->maxGroup = (uint32_t)v;
    }
    if (const char *e = getenv("ISSL_WAVES")) { const int v = atoi(e); if (v >= 0 && v <= 2) d->waves = v; }
    if (const char *e = getenv("ISSL_HIT_CAP")) { const long v = atol(e); if (v > 0) d->firstHitCap = (uint64_t)v; }
    if (const char *e = getenv("ISSL_TRIPLE_FLUSH")) d->tripleFlush = atoi(e) != 0;
    if (const char *e = getenv("ISSL_TRIPLE_HEAVY")) d->tripleHeavy = atoi(e) != 0;
    if (const char *e = getenv("ISSL_TRIPLE_SMALL")) d->tripleSmall = atoi(e) != 0;
    if (const char *e = getenv("ISSL_TRIPLE_FUSE")) { const int v = atoi(e); if (v >= 0 && v <= 2) d->tripleFuse = v; }
    if (const char *e = getenv("ISSL_TRIPLE_MAXDIST")) {
        const long v = atol(e);
        if (v >= -1 && v <= 7) d->tripleMaxDist = (int)v;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&d->hCounters, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        delete d;
        return issl_set_error(ISSL_ERR_CUDA, "device %d setup: %s", cuda_device, cudaGetErrorString(e));
    }
    int rc = upload_constants();
    if (rc != ISSL_OK) { issl_device_destroy(d); return rc; }
    *out = d;
    return ISSL_OK;
}

extern "C" void issl_device_destroy(issl_device *d)
{
    if (!d) return;
    cudaSetDevice(d->dev);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (DBuf *b : {&d->sig, &d->occ, &d->ids, &d->res32, &d->sig64, &d->listStart, &d->listLen, &d->filePrefix,
                    &d->mitMasks, &d->mitScores, &d->guides, &d->totMit, &d->totCfd, &d->done, &d->pairKeys, &d->pairVals, &d->pairKeysSorted, &d->pairValsSorted, &d->pairCounts,
                    &d->pairOffsets, &d->items, &d->keysA, &d->keysB, &d->sortTemp, &d->scanTemp, &d->contribMit,
                    &d->contribCfd, &d->counters, &d->outMit, &d->outCfd, &d->hitId, &d->hitDist, &d->hitOcc,
                    &d->scoredEnd, &d->segBegin, &d->heavyKeys, &d->heavyDesc, &d->heavyFlat, &d->redo, &d->ovfBits, &d->tripleBits, &d->tripleRes, &d->tripleIds, &d->tripleOffs, &d->visits, &d->segOff, &d->segCnt, &d->tripleBlk, &d->segKeys, &d->segSites, &d->totMit2, &d->totCfd2, &d->done2,
                    &d->mitDense})
        b->release();
    for (cudaEvent_t ev : d->evPool) cudaEventDestroy(ev);
    if (d->hCounters) cudaFreeHost(d->hCounters);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

// ---------------------------------------------------------------------------------------------
// residency from an .issl image
// ---------------------------------------------------------------------------------------------
extern "C" int issl_device_create(const issl_index *ix, int cuda_device, int layout, issl_device **out)
{
    if (!ix || !out) return issl_set_error(ISSL_ERR_ARG, "issl_device_create: null argument");
    *out = nullptr;
    int lay = 0;
    CKR(choose_layout(ix->info, layout, &lay));
    issl_device *d = nullptr;
    CKR(new_device(cuda_device, &d));
    d->layoutAuto = (layout == ISSL_LAYOUT_AUTO);
    auto fail = [&](int rc) { issl_device_destroy(d); return rc; };

    int rc = init_geometry(d, ix->info, lay, ix->sizes);
    if (rc != ISSL_OK) return fail(rc);
    issl_sorted_score_table(ix->scorePairs, ix->scoresInFile, d->hMitMasks, d->hMitScores);
    if ((rc = upload_mit_table(d)) != ISSL_OK) return fail(rc);

    const uint64_t N = ix->info.offtargetsCount, S = ix->info.sliceCount;
    // signatures first (the re-layout kernel gathers them)
    constexpr size_t kStage = 256ull << 20;
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t freeEv[2] = {nullptr, nullptr};
    DBuf dstage[2];
    unsigned long long *dErr = nullptr;
    TripleBuild tbuild;
    cudaStream_t st2 = nullptr;
    cudaEvent_t evSlice0 = nullptr;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; b++) {
            if (stage[b]) cudaFreeHost(stage[b]);
            if (freeEv[b]) cudaEventDestroy(freeEv[b]);
            dstage[b].release();
        }
        if (dErr) cudaFree(dErr);
        if (st2) { cudaStreamSynchronize(st2); cudaStreamDestroy(st2); st2 = nullptr; }
        if (evSlice0) { cudaEventDestroy(evSlice0); evSlice0 = nullptr; }
    };
#define CKF(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            issl_set_error(ISSL_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            cleanup();                                                                                    \
            return fail(ISSL_ERR_CUDA);                                                                   \
        }                                                                                                 \
    } while (0)
    for (int b = 0; b < 2; b++) {
        CKF(cudaMallocHost(&stage[b], kStage + 8));
        CKF(cudaEventCreateWithFlags(&freeEv[b], cudaEventDisableTiming));
        if (dstage[b].ensure(kStage + 8) != ISSL_OK) { cleanup(); return fail(ISSL_ERR_NOMEM); }
    }
    CKF(cudaMalloc(&dErr, 8));
    CKF(cudaMemsetAsync(dErr, 0, 8, d->stream));
    // TRIPLE: the ten sub-bucket copies are built on a second stream while slices 1.. are still crossing PCIe (they need
    // sig[] and occ[] only, complete after slice 0); ISSL_TRIPLE_LATE=1 builds them after the upload, as round 1 did
    int earlyRc = -1;   // -1: not started
    const bool early = d->layout == ISSL_LAYOUT_TRIPLE && !getenv("ISSL_TRIPLE_LATE");

    int buf = 0;
    // 1) signatures: host image -> pinned -> device
    for (uint64_t o = 0; o < N * 8; o += kStage) {
        const size_t n = (size_t)std::min<uint64_t>(kStage, N * 8 - o);
        CKF(cudaEventSynchronize(freeEv[buf]));
        parallel_copy(stage[buf], reinterpret_cast<const uint8_t *>(ix->offtargets) + o, n);
        CKF(cudaMemcpyAsync(d->sig.as<uint8_t>() + o, stage[buf], n, cudaMemcpyHostToDevice, d->stream));
        CKF(cudaEventRecord(freeEv[buf], d->stream));
        buf ^= 1;
    }
    // 2) lists, slice by slice, in chunks carrying one entry of overlap for the ascending-id check
    const uint64_t chunkEntries = kStage / 8;
    for (uint64_t s = 0; s < S; s++) {
        for (uint64_t q0 = s * N; q0 < (s + 1) * N; q0 += chunkEntries) {
            const uint64_t n = std::min<uint64_t>(chunkEntries, (s + 1) * N - q0);
            const int hasPrev = q0 > s * N;
            CKF(cudaEventSynchronize(freeEv[buf]));
            parallel_copy(stage[buf], ix->entries + q0 - hasPrev, (n + hasPrev) * 8);
            CKF(cudaMemcpyAsync(dstage[buf].p, stage[buf], (n + hasPrev) * 8, cudaMemcpyHostToDevice, d->stream));
            RelayoutArgs a;
            a.iv = d->iv; a.entries = dstage[buf].as<uint64_t>(); a.filePrefix = d->filePrefix.as<uint64_t>();
            a.q0 = q0; a.n = n; a.slice = (uint32_t)s; a.hasPrev = hasPrev; a.errors = dErr;
            k_relayout<<<blocks_for(n, 256), 256, 0, d->stream>>>(a);
            CKF(cudaGetLastError());
            CKF(cudaEventRecord(freeEv[buf], d->stream));
            buf ^= 1;
        }
        if (s == 0 && early) {
            CKF(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
            CKF(cudaEventCreateWithFlags(&evSlice0, cudaEventDisableTiming));
            CKF(cudaEventRecord(evSlice0, d->stream));
            CKF(cudaStreamWaitEvent(st2, evSlice0, 0));
            earlyRc = build_triple_enqueue(d, st2, tbuild);
        }
    }
    CKF(cudaMemcpyAsync(d->hCounters, dErr, 8, cudaMemcpyDeviceToHost, d->stream));
    CKF(cudaStreamSynchronize(d->stream));
    const unsigned long long bad = d->hCounters[0];
    if (earlyRc == ISSL_OK) earlyRc = build_triple_finish(d, st2, tbuild);
#undef CKF
    cleanup();
    for (DBuf *x : {&tbuild.keysIn, &tbuild.keysOut, &tbuild.idsIn, &tbuild.tmp}) x->release();
    if (bad)
        return fail(issl_set_error(ISSL_ERR_UNSUPPORTED,
                                   "Error reading index: %llu list entries violate the isslCreateIndex invariants "
                                   "(id range, list membership, ascending ids or occurrence counts)", bad));
    rc = earlyRc >= 0 ? triple_fall_back(d, earlyRc) : build_triple_or_fall_back(d);
    if (rc != ISSL_OK) return fail(rc);
    *out = d;
    return ISSL_OK;
}

// ---------------------------------------------------------------------------------------------
// on-device construction from sorted keys (synthetic indexes)
// ---------------------------------------------------------------------------------------------
static int collapse_and_build(issl_device *d, int layout, const uint64_t *dKeys, DBuf &flags, bool haveFlags,
                              bool valuesAreSortKeys, uint64_t nRaw, uint32_t L, uint32_t w);

static int build_from_sites(issl_device *d, int layout, uint64_t *dKeys, uint64_t nRaw, uint32_t L, uint32_t w,
                            DBuf &keysAlt)
{
    cudaStream_t st = d->stream;
    // sort raw keys (lexicographic site order)
    {
        cub::DoubleBuffer<uint64_t> db(dKeys, keysAlt.as<uint64_t>());
        size_t tb = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, nRaw, 0, (int)(2 * L), st));
        DBuf tmp; CKR(tmp.ensure(tb));
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, nRaw, 0, (int)(2 * L), st));
        CK(cudaStreamSynchronize(st));
        tmp.release();
        if (db.Current() != dKeys) CK(cudaMemcpyAsync(dKeys, db.Current(), nRaw * 8, cudaMemcpyDeviceToDevice, st));
    }
    DBuf flags;
    const int rc = collapse_and_build(d, layout, dKeys, flags, false, true, nRaw, L, w);
    flags.release();
    return rc;
}

// Everything isslCreateIndex does after reading its input (ref isslCreateIndex.cpp:184-252), on the device:
// run-length collapse of adjacent equal records into (signature, occurrences), per-slice lists, score table.
// dKeys[nRaw]: one value per input record in input order -- sort keys (synthetic path) or signatures (text path).
// flags[nRaw] (optional): 1 where a record differs from its predecessor (the text path compares the text itself,
// as the reference's memcmp at :192 does); computed from value equality otherwise.
static int collapse_and_build(issl_device *d, int layout, const uint64_t *dKeys, DBuf &flags, bool haveFlags,
                              bool valuesAreSortKeys, uint64_t nRaw, uint32_t L, uint32_t w)
{
    cudaStream_t st = d->stream;
    if (nRaw >= (1ull << 32)) return issl_set_error(ISSL_ERR_UNSUPPORTED, "more than 2^32 input records");
    // run-length collapse (ref isslCreateIndex.cpp:184-207)
    DBuf ranks, runStart;
    CKR(flags.ensure(nRaw * 4));
    CKR(ranks.ensure(nRaw * 8));
    if (!haveFlags) k_run_flags<<<blocks_for(nRaw, 256), 256, 0, st>>>(dKeys, nRaw, flags.as<uint32_t>());
    {
        size_t tb = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tb, flags.as<uint32_t>(), ranks.as<uint64_t>(), nRaw, st));
        DBuf tmp; CKR(tmp.ensure(tb));
        CK(cub::DeviceScan::InclusiveSum(tmp.p, tb, flags.as<uint32_t>(), ranks.as<uint64_t>(), nRaw, st));
        CK(cudaMemcpyAsync(d->hCounters, ranks.as<uint64_t>() + (nRaw - 1), 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tmp.release();
    }
    const uint64_t N = d->hCounters[0];
    if (N >= (1ull << 32)) return issl_set_error(ISSL_ERR_UNSUPPORTED, "more than 2^32 distinct sites");

    issl_info f{};
    f.offtargetsCount = N; f.seqLength = L; f.seqCount = nRaw; f.sliceWidth = w; f.sliceCount = (2 * L) / w;
    // score table exactly as the reference builder would store it (ref isslCreateIndex.cpp:239-252)
    {
        uint64_t sc = 0;
        issl_mit_table(L, w, nullptr, nullptr, 0, &sc);
        d->hMitMasks.resize(sc); d->hMitScores.resize(sc);
        issl_mit_table(L, w, d->hMitMasks.data(), d->hMitScores.data(), sc, &sc);
        f.scoresCount = sc;
    }

    // signatures + occurrences into temporaries (init_geometry needs the list lengths first)
    DBuf sigTmp, occTmp;
    CKR(sigTmp.ensure(N * 8)); CKR(occTmp.ensure(N * 4)); CKR(runStart.ensure(N * 8));
    k_run_scatter<<<blocks_for(nRaw, 256), 256, 0, st>>>(dKeys, flags.as<uint32_t>(), ranks.as<uint64_t>(), nRaw, L,
                                                        valuesAreSortKeys ? 1 : 0, sigTmp.as<uint64_t>(), runStart.as<uint64_t>());
    k_run_lengths<<<blocks_for(N, 256), 256, 0, st>>>(runStart.as<uint64_t>(), N, nRaw, occTmp.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ranks.release(); runStart.release();

    // per-slice values, histograms
    const uint64_t S = f.sliceCount, sliceLimit = 1ull << w;
    const uint32_t smask = (uint32_t)(sliceLimit - 1);
    DBuf values, valuesSorted, idsTmp, idsSorted, hist, valueStart;
    CKR(values.ensure(N)); CKR(valuesSorted.ensure(N)); CKR(idsTmp.ensure(N * 4)); CKR(idsSorted.ensure(N * 4));
    CKR(hist.ensure(S * 256 * 8)); CKR(valueStart.ensure(256 * 8));
    CK(cudaMemsetAsync(hist.p, 0, S * 256 * 8, st));
    // pass A: histograms of all slices (list lengths), so that geometry can be fixed
    for (uint64_t s = 0; s < S; s++)
        k_slice_values<<<148 * 8, 256, 0, st>>>(sigTmp.as<uint64_t>(), N, w, smask, (uint32_t)s, values.as<uint8_t>(),
                                               idsTmp.as<uint32_t>(), hist.as<unsigned long long>() + s * 256);
    std::vector<unsigned long long> hHist(S * 256);
    CK(cudaMemcpyAsync(hHist.data(), hist.p, S * 256 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<uint64_t> listLen(S * sliceLimit, 0);
    for (uint64_t s = 0; s < S; s++)
        for (uint64_t v = 0; v < 256 && v < sliceLimit; v++) listLen[s * sliceLimit + v] = hHist[s * 256 + v];

    CKR(init_geometry(d, f, layout, listLen.data()));
    CKR(upload_mit_table(d));
    CK(cudaMemcpyAsync(d->sig.p, sigTmp.p, N * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(d->occ.p, occTmp.p, N * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    sigTmp.release(); occTmp.release();

    for (DBuf *b : {&values, &valuesSorted, &idsTmp, &idsSorted, &hist, &valueStart}) b->release();
    // pass B: per slice, stable 8-bit radix sort of (value, id) and placement (ref :216-234) -- unless the lists can wait
    if (d->listsResident) CKR(fill_lists(d));
    return build_triple_or_fall_back(d);
}

extern "C" int issl_device_create_synthetic_ex(int cuda_device, int layout, const issl_synth_spec *spec, issl_device **out)
{
    if (!out || !spec) return issl_set_error(ISSL_ERR_ARG, "issl_device_create_synthetic_ex: null argument");
    *out = nullptr;
    const uint32_t seqLength = spec->seqLength, sliceWidth = spec->sliceWidth;
    if (seqLength == 0 || seqLength > 32 || sliceWidth < 2 || sliceWidth > 24 || (2 * seqLength) / sliceWidth == 0)
        return issl_set_error(ISSL_ERR_ARG, "issl_device_create_synthetic: bad sequence length / slice width");
    if (spec->family_size_max < spec->family_size_min || !(spec->low_complexity_fraction >= 0.0 && spec->low_complexity_fraction <= 1.0))
        return issl_set_error(ISSL_ERR_ARG, "issl_device_create_synthetic: bad family sizes / low-complexity fraction");
    // family sizes: fixed, or log-uniform in [min, max] from the same counter-based generator as the sites
    std::vector<uint64_t> start(spec->families + 1ull, 0);
    for (uint32_t f = 0; f < spec->families; f++) {
        uint64_t size = spec->family_size_min;
        if (spec->family_size_max > spec->family_size_min && spec->family_size_min > 0) {
            const double u = (double)(rng3(spec->seed, 7, f) >> 11) * (1.0 / 9007199254740992.0);
            size = (uint64_t)std::floor(std::exp(std::log((double)spec->family_size_min) +
                                                 u * (std::log((double)spec->family_size_max) - std::log((double)spec->family_size_min))));
            size = std::min<uint64_t>(std::max<uint64_t>(size, spec->family_size_min), spec->family_size_max);
        }
        start[f + 1] = start[f] + size;
    }
    const uint64_t nFamily = start[spec->families];
    const uint64_t nLow = (uint64_t)std::floor(spec->low_complexity_fraction * (double)(spec->uniform_sites + nFamily));
    const uint64_t nRaw = spec->uniform_sites + nFamily + nLow;
    if (nRaw == 0) return issl_set_error(ISSL_ERR_ARG, "issl_device_create_synthetic: no sites");
    issl_info f{};
    f.seqLength = seqLength; f.sliceWidth = sliceWidth; f.sliceCount = (2 * seqLength) / sliceWidth;
    int lay = 0;
    CKR(choose_layout(f, layout, &lay));
    issl_device *d = nullptr;
    CKR(new_device(cuda_device, &d));
    d->layoutAuto = (layout == ISSL_LAYOUT_AUTO);
    DBuf keys, keysAlt, dStart;
    int rc = keys.ensure(nRaw * 8);
    if (rc == ISSL_OK) rc = keysAlt.ensure(nRaw * 8);
    if (rc == ISSL_OK) rc = dStart.ensure(start.size() * 8);
    if (rc == ISSL_OK && cudaMemcpyAsync(dStart.p, start.data(), start.size() * 8, cudaMemcpyHostToDevice, d->stream) != cudaSuccess)
        rc = issl_set_error(ISSL_ERR_CUDA, "synthetic build: upload of the family table failed");
    if (rc == ISSL_OK) {
        SynthArgs a;
        a.seed = spec->seed; a.nUniform = spec->uniform_sites; a.nFamilySites = nFamily; a.nLow = nLow;
        a.families = spec->families; a.familyStart = dStart.as<uint64_t>(); a.maxSubRate = spec->max_sub_rate;
        a.L = seqLength; a.keys = keys.as<uint64_t>();
        k_synth_sites<<<blocks_for(nRaw, 256), 256, 0, d->stream>>>(a);
        rc = build_from_sites(d, lay, keys.as<uint64_t>(), nRaw, seqLength, sliceWidth, keysAlt);
    }
    if (rc == ISSL_OK) {
        cudaError_t e = cudaStreamSynchronize(d->stream);
        if (e != cudaSuccess) rc = issl_set_error(ISSL_ERR_CUDA, "synthetic build: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(d->stream);   // `start` is a local the upload may still be reading
    }
    keys.release(); keysAlt.release(); dStart.release();
    if (rc != ISSL_OK) { issl_device_destroy(d); return rc; }
    *out = d;
    return ISSL_OK;
}

extern "C" int issl_device_create_synthetic(int cuda_device, int layout, uint64_t seed, uint64_t uniform_sites,
                                            uint32_t families, uint32_t family_size, double max_sub_rate,
                                            uint32_t seqLength, uint32_t sliceWidth, issl_device **out)
{
    issl_synth_spec spec{};
    spec.seed = seed; spec.uniform_sites = uniform_sites; spec.families = families;
    spec.family_size_min = spec.family_size_max = family_size; spec.max_sub_rate = max_sub_rate;
    spec.low_complexity_fraction = 0.0; spec.seqLength = seqLength; spec.sliceWidth = sliceWidth;
    return issl_device_create_synthetic_ex(cuda_device, layout, &spec, out);
}

// ref isslScoreOfftargets.cpp:204-216 (allSlicelistSizes): the length of every slice list, slice-major
extern "C" size_t issl_device_list_lengths(const issl_device *d, uint64_t *out, size_t cap)
{
    if (!d) return 0;
    const size_t n = (size_t)d->nLists;
    if (out) for (size_t i = 0; i < std::min(n, cap); i++) out[i] = d->hListLen[i];
    return n;
}

// internal (issl_sites.cu): an index from `nRaw` site sort keys already on this device, in any order.  Both key
// buffers are scratch afterwards; the caller still owns them.
int issl_internal_device_from_keys(int cuda_device, int layout, uint64_t *dKeys, uint64_t *dKeysAlt, uint64_t nRaw,
                                   uint32_t seqLength, uint32_t sliceWidth, issl_device **out)
{
    *out = nullptr;
    if (seqLength == 0 || seqLength > 32 || sliceWidth < 2 || sliceWidth > 24 || (2 * seqLength) / sliceWidth == 0)
        return issl_set_error(ISSL_ERR_ARG, "bad sequence length / slice width");
    if (nRaw == 0) return issl_set_error(ISSL_ERR_ARG, "no sites");
    issl_info f{};
    f.seqLength = seqLength; f.sliceWidth = sliceWidth; f.sliceCount = (2 * seqLength) / sliceWidth;
    int lay = 0;
    CKR(choose_layout(f, layout, &lay));
    issl_device *d = nullptr;
    CKR(new_device(cuda_device, &d));
    d->layoutAuto = (layout == ISSL_LAYOUT_AUTO);
    DBuf alt;
    alt.view(dKeysAlt, nRaw * 8);   // the caller keeps ownership
    int rc = build_from_sites(d, lay, dKeys, nRaw, seqLength, sliceWidth, alt);
    if (rc == ISSL_OK) {
        cudaError_t e = cudaStreamSynchronize(d->stream);
        if (e != cudaSuccess) rc = issl_set_error(ISSL_ERR_CUDA, "index build: %s", cudaGetErrorString(e));
    }
    if (rc != ISSL_OK) { issl_device_destroy(d); return rc; }
    *out = d;
    return ISSL_OK;
}

// ---------------------------------------------------------------------------------------------
// construction from the text file isslCreateIndex reads (ref isslCreateIndex.cpp:138-207)
// ---------------------------------------------------------------------------------------------
extern "C" int issl_device_create_from_text(const char *text, size_t bytes, uint32_t seqLength, uint32_t sliceWidth,
                                            int cuda_device, int layout, issl_device **out)
{
    if (!out || (bytes && !text)) return issl_set_error(ISSL_ERR_ARG, "issl_device_create_from_text: null argument");
    *out = nullptr;
    if (seqLength == 0 || seqLength > 32)   // ref :142-145
        return issl_set_error(ISSL_ERR_ARG, "Sequence length is greater than 32, which is the maximum supported currently");
    if (sliceWidth < 2 || sliceWidth > 24 || (2 * seqLength) / sliceWidth == 0 || (2 * seqLength) / sliceWidth - 1 >= 20)
        return issl_set_error(ISSL_ERR_ARG, "issl_device_create_from_text: unsupported slice width %u", sliceWidth);
    const size_t line = seqLength + 1;
    if (bytes % line != 0)                  // ref :147-153
        return issl_set_error(ISSL_ERR_ARG, "Error: file does is not a multiple of the expected line length (%zu)", line);
    const uint64_t nRaw = bytes / line;
    if (nRaw == 0) return issl_set_error(ISSL_ERR_IO, "Failed to read in file.");   // ref :176-179
    issl_info f{};
    f.seqLength = seqLength; f.sliceWidth = sliceWidth; f.sliceCount = (2 * seqLength) / sliceWidth;
    int lay = 0;
    CKR(choose_layout(f, layout, &lay));
    issl_device *d = nullptr;
    CKR(new_device(cuda_device, &d));
    d->layoutAuto = (layout == ISSL_LAYOUT_AUTO);

    DBuf sigRaw, flags, dtext[2];
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t freeEv[2] = {nullptr, nullptr};
    int rc = ISSL_OK;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; b++) {
            if (stage[b]) cudaFreeHost(stage[b]);
            if (freeEv[b]) cudaEventDestroy(freeEv[b]);
            dtext[b].release();
        }
        sigRaw.release(); flags.release();
    };
    // chunks of whole lines, each preceded by one line of overlap (for the "differs from predecessor" test)
    const uint64_t linesPerChunk = std::max<uint64_t>(1, (192ull << 20) / line);
    const size_t chunkBytes = (size_t)((linesPerChunk + 1) * line);
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == ISSL_OK) rc = issl_set_error(ISSL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    if ((rc = sigRaw.ensure(nRaw * 8)) == ISSL_OK) rc = flags.ensure(nRaw * 4);
    for (int b = 0; b < 2 && rc == ISSL_OK; b++) {
        cu(cudaMallocHost(&stage[b], chunkBytes), "cudaMallocHost");
        cu(cudaEventCreateWithFlags(&freeEv[b], cudaEventDisableTiming), "cudaEventCreate");
        if (rc == ISSL_OK) rc = dtext[b].ensure(chunkBytes);
    }
    int buf = 0;
    for (uint64_t l0 = 0; l0 < nRaw && rc == ISSL_OK; l0 += linesPerChunk) {
        const uint64_t n = std::min<uint64_t>(linesPerChunk, nRaw - l0);
        const int hasPrev = l0 > 0;
        if (!cu(cudaEventSynchronize(freeEv[buf]), "cudaEventSynchronize")) break;
        parallel_copy(stage[buf], text + (l0 - hasPrev) * line, (size_t)((n + hasPrev) * line));
        if (!cu(cudaMemcpyAsync(dtext[buf].p, stage[buf], (size_t)((n + hasPrev) * line), cudaMemcpyHostToDevice, d->stream), "H2D of site text")) break;
        k_pack_lines<<<blocks_for(n, 256), 256, 0, d->stream>>>(dtext[buf].as<char>(), n, hasPrev, seqLength,
                                                               sigRaw.as<uint64_t>() + l0, flags.as<uint32_t>() + l0);
        cu(cudaGetLastError(), "k_pack_lines");
        cu(cudaEventRecord(freeEv[buf], d->stream), "cudaEventRecord");
        buf ^= 1;
    }
    if (rc == ISSL_OK) cu(cudaStreamSynchronize(d->stream), "site text upload");
    for (int b = 0; b < 2; b++) dtext[b].release();
    if (rc == ISSL_OK) rc = collapse_and_build(d, lay, sigRaw.as<uint64_t>(), flags, true, false, nRaw, seqLength, sliceWidth);
    if (rc == ISSL_OK) cu(cudaStreamSynchronize(d->stream), "index build");
    cleanup();
    if (rc != ISSL_OK) { issl_device_destroy(d); return rc; }
    *out = d;
    return ISSL_OK;
}

// ---------------------------------------------------------------------------------------------
// a second copy of a resident index on another GPU, straight over NVLink: the index is replicated per GPU
// (guides are independent, ref isslScoreOfftargets.cpp:316-317), and a peer copy of the finished layout
// (~0.1 s for 87 GB) replaces another upload of the file from the host plus another build of the ten copies
// ---------------------------------------------------------------------------------------------
namespace {
// the buffers that make up a resident index, in a fixed order
std::vector<DBuf *> index_buffers(issl_device *d)
{
    return {&d->sig, &d->occ, &d->ids, &d->res32, &d->sig64, &d->listStart, &d->listLen, &d->filePrefix, &d->mitMasks,
            &d->mitScores, &d->mitDense, &d->tripleRes, &d->tripleIds, &d->tripleOffs, &d->tripleBlk, &d->tripleBits};
}
template <class T> const T *rebase(const T *p, const DBuf &from, const DBuf &to)
{
    return p ? reinterpret_cast<const T *>(static_cast<const uint8_t *>(to.p) + (reinterpret_cast<const uint8_t *>(p) - static_cast<const uint8_t *>(from.p)))
             : nullptr;
}
}  // namespace

extern "C" int issl_device_clone(const issl_device *src_, int cuda_device, issl_device **out)
{
    if (!src_ || !out) return issl_set_error(ISSL_ERR_ARG, "issl_device_clone: null argument");
    *out = nullptr;
    issl_device *src = const_cast<issl_device *>(src_);   // buffers are only read
    if (cuda_device == src->dev) return issl_set_error(ISSL_ERR_ARG, "issl_device_clone: device %d already holds this index", cuda_device);
    issl_device *d = nullptr;
    CKR(new_device(cuda_device, &d));
    auto fail = [&](int rc) { issl_device_destroy(d); return rc; };
    // direct peer copies need peer access from the copying side; without it the driver stages through the host
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, cuda_device, src->dev) == cudaSuccess && can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(src->dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        cudaGetLastError();
    }
    d->info = src->info; d->layout = src->layout; d->nLists = src->nLists; d->hbmBytes = src->hbmBytes; d->pbits = src->pbits;
    d->mitCount = src->mitCount; d->hMitMasks = src->hMitMasks; d->hMitScores = src->hMitScores;
    d->hListLen = src->hListLen; d->hListStart = src->hListStart; d->hFilePrefix = src->hFilePrefix;
    d->layoutAuto = src->layoutAuto; d->tripleMaxDist = src->tripleMaxDist; d->listsResident = src->listsResident;
    const std::vector<DBuf *> from = index_buffers(src), to = index_buffers(d);
    for (size_t k = 0; k < from.size(); k++) {
        if (!from[k]->p) continue;
        int rc = to[k]->exact(from[k]->used);
        if (rc != ISSL_OK) return fail(rc);
        const cudaError_t e = cudaMemcpyPeerAsync(to[k]->p, cuda_device, from[k]->p, src->dev, from[k]->used, d->stream);
        if (e != cudaSuccess) return fail(issl_set_error(ISSL_ERR_CUDA, "peer copy %d -> %d: %s", src->dev, cuda_device, cudaGetErrorString(e)));
    }
    d->iv = src->iv;
    d->iv.sig = rebase(src->iv.sig, src->sig, d->sig); d->iv.occ = rebase(src->iv.occ, src->occ, d->occ);
    d->iv.ids = rebase(src->iv.ids, src->ids, d->ids); d->iv.res32 = rebase(src->iv.res32, src->res32, d->res32);
    d->iv.sig64 = rebase(src->iv.sig64, src->sig64, d->sig64);
    d->iv.listStart = rebase(src->iv.listStart, src->listStart, d->listStart);
    d->iv.listLen = rebase(src->iv.listLen, src->listLen, d->listLen);
    d->tv = src->tv;
    d->tv.res = rebase(src->tv.res, src->tripleRes, d->tripleRes); d->tv.ids = rebase(src->tv.ids, src->tripleIds, d->tripleIds);
    d->tv.offs = rebase(src->tv.offs, src->tripleOffs, d->tripleOffs); d->tv.blk = rebase(src->tv.blk, src->tripleBlk, d->tripleBlk);
    d->tv.nonEmpty = rebase(src->tv.nonEmpty, src->tripleBits, d->tripleBits);
    const cudaError_t e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) return fail(issl_set_error(ISSL_ERR_CUDA, "peer copy %d -> %d: %s", src->dev, cuda_device, cudaGetErrorString(e)));
    *out = d;
    return ISSL_OK;
}

extern "C" int issl_device_get_info(const issl_device *d, issl_device_info *out)
{
    if (!d || !out) return issl_set_error(ISSL_ERR_ARG, "issl_device_get_info: null argument");
    out->cuda_device = d->dev;
    out->layout = d->layout;
    out->bytes_per_candidate = d->layout == ISSL_LAYOUT_TRIPLE ? 2 : d->layout == ISSL_LAYOUT_RES32 ? 4 : (d->layout == ISSL_LAYOUT_SIG64 ? 8 : 12);
    out->hbm_bytes = d->hbmBytes;
    out->list_entries = d->info.sliceCount * d->info.offtargetsCount;
    out->info = d->info;
    out->triple_block_bytes = d->layout == ISSL_LAYOUT_TRIPLE ? d->tv.pitch * 2 : 0;
    // what the fused tail gathers per hit: an offset pair, an id and the 16-byte record of the non-fused modes -- or
    // nothing, when the site's own text rank orders the hits
    out->triple_hit_bytes = d->layout == ISSL_LAYOUT_TRIPLE ? ((d->tv.siteOrdered && d->tv.pitch) ? 0 : 28) : 0;
    return ISSL_OK;
}

extern "C" int issl_device_read_sites(issl_device *d, const uint64_t *site_ids, uint64_t n, uint64_t *out)
{
    if (!d || (n && (!site_ids || !out))) return issl_set_error(ISSL_ERR_ARG, "issl_device_read_sites: null argument");
    if (n == 0) return ISSL_OK;
    CK(cudaSetDevice(d->dev));
    DBuf in, o;
    CKR(in.ensure(n * 8)); CKR(o.ensure(n * 8));
    CK(cudaMemcpyAsync(in.p, site_ids, n * 8, cudaMemcpyHostToDevice, d->stream));
    k_gather_sites<<<blocks_for(n, 256), 256, 0, d->stream>>>(d->iv.sig, d->iv.N, in.as<uint64_t>(), n, o.as<uint64_t>());
    CK(cudaMemcpyAsync(out, o.p, n * 8, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    in.release(); o.release();
    return ISSL_OK;
}

extern "C" int issl_device_write_issl(issl_device *d, const char *path)
{
    if (!d || !path) return issl_set_error(ISSL_ERR_ARG, "issl_device_write_issl: null argument");
    CK(cudaSetDevice(d->dev));
    CKR(ensure_lists(d));
    FILE *fp = fopen(path, "wb");
    if (!fp) return issl_set_error(ISSL_ERR_IO, "cannot create %s", path);
    auto put = [&](const void *p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, fp) == bytes; };
    bool ok = true;
    const uint64_t header[6] = {d->info.offtargetsCount, d->info.seqLength, d->info.seqCount,
                                d->info.sliceWidth, d->info.sliceCount, d->info.scoresCount};
    ok &= put(header, sizeof header);
    for (size_t k = 0; k < d->hMitMasks.size() && ok; k++) { ok &= put(&d->hMitMasks[k], 8); ok &= put(&d->hMitScores[k], 8); }
    constexpr size_t kStage = 64ull << 20;
    uint8_t *stage = nullptr;
    DBuf dbuf;
    int rc = ISSL_OK;
    if (cudaMallocHost(&stage, kStage) != cudaSuccess) { fclose(fp); return issl_set_error(ISSL_ERR_NOMEM, "pinned staging buffer"); }
    const uint64_t N = d->info.offtargetsCount;
    for (uint64_t o = 0; o < N * 8 && ok; o += kStage) {
        const size_t n = (size_t)std::min<uint64_t>(kStage, N * 8 - o);
        if (cudaMemcpyAsync(stage, d->sig.as<uint8_t>() + o, n, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
            cudaStreamSynchronize(d->stream) != cudaSuccess) { rc = issl_set_error(ISSL_ERR_CUDA, "D2H of signatures failed"); ok = false; break; }
        ok &= put(stage, n);
    }
    ok = ok && put(d->hListLen.data(), d->nLists * 8);
    if (ok && (rc = dbuf.ensure(kStage)) == ISSL_OK) {
        const uint64_t total = d->hFilePrefix[d->nLists], per = kStage / 8;
        for (uint64_t q0 = 0; q0 < total && ok; q0 += per) {
            const uint64_t n = std::min<uint64_t>(per, total - q0);
            k_export_entries<<<blocks_for(n, 256), 256, 0, d->stream>>>(d->iv, d->filePrefix.as<uint64_t>(), d->nLists, q0, n,
                                                                       dbuf.as<uint64_t>());
            if (cudaMemcpyAsync(stage, dbuf.p, n * 8, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
                cudaStreamSynchronize(d->stream) != cudaSuccess) { rc = issl_set_error(ISSL_ERR_CUDA, "D2H of list entries failed"); ok = false; break; }
            ok &= put(stage, n * 8);
        }
    }
    cudaFreeHost(stage);
    dbuf.release();
    if (fclose(fp) != 0) ok = false;
    if (rc != ISSL_OK) return rc;
    if (!ok) return issl_set_error(ISSL_ERR_IO, "short write to %s", path);
    return ISSL_OK;
}

// ---------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------
namespace {

struct HitSink {     // host-side collection for issl_score_hits
    std::vector<uint64_t> guide;
    std::vector<uint32_t> id, occ;
    std::vector<int32_t> dist;
};

struct EventTimer {
    issl_device *d;
    size_t used = 0;
    std::vector<std::pair<size_t, size_t>> scanPairs, heavyPairs;
    int get(cudaEvent_t *ev)
    {
        if (used == d->evPool.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            d->evPool.push_back(e);
        }
        *ev = d->evPool[used++];
        return ISSL_OK;
    }
};

}  // namespace

// fusedTail: guides are finished where their hits are (the scan kernel's fused tail / k_heavy_finish): only what spills from a
// CTA's record list reaches this buffer, so it starts small -- 600 MB of cudaMalloc were most of a first call's time on a
// fresh handle, and in the executable every call is the first; an overflow re-launches the scan with a larger one, as ever
static int ensure_hit_buffers(issl_device *d, uint32_t n, bool fusedTail = false)
{
    uint64_t wantCap = std::min<uint64_t>(std::max<uint64_t>(1ull << 22, fusedTail ? 0ull : 384ull * n), 1ull << 28);
    if (d->firstHitCap) wantCap = d->firstHitCap;   // ISSL_HIT_CAP: start small, so that tests reach the re-launch after an overflow
    if (d->hitCap < wantCap && d->hitCap == d->hitCapAuto) {   // first sizing for this batch size (uniform genomes: ~275 survivors per guide)
        d->hitCap = d->hitCapAuto = wantCap;
        CKR(d->keysA.ensure(d->hitCap * 8)); CKR(d->keysB.ensure(d->hitCap * 8));
    }
    return ISSL_OK;
}

static int ensure_visits(issl_device *d, int maxDist, cudaStream_t st)
{
    if (d->visitsDist != maxDist) {   // the visit table depends on maxDist only
        const bool w4 = d->info.sliceWidth == 4;   // ten 2-base slices: also the buckets without an exact byte (maxDist >= 5)
        std::vector<uint32_t> raw(w4 ? issl_triple_visits_w4(maxDist, nullptr, 0, nullptr) : issl_triple_visits(maxDist, nullptr, 0, nullptr));
        if (w4) issl_triple_visits_w4(maxDist, raw.data(), raw.size(), d->waveStart);
        else issl_triple_visits(maxDist, raw.data(), raw.size(), d->waveStart);
        static const uint8_t slices[10][5] = ISSL_TRIPLE_LAYOUT_INIT;
        std::vector<TripleVisit> v(raw.size());
        for (size_t i = 0; i < raw.size(); i++) {
            const uint32_t t = (raw[i] >> 24) & 15u;
            uint32_t exact = 0;   // slices of the triple on which the bucket agrees with the guide
            for (int k = 0; k < 3; k++)
                if (((raw[i] >> (8 * k)) & 0xFFu) == 0) exact |= 1u << slices[t][k];
            // which of its entries this visit reports: resp(E) == t, for the four ways the residual's two slices can match exactly
            uint32_t keep = 0;
            for (uint32_t c = 0; c < 4; c++) {
                const uint32_t E = exact | ((c & 1u) << slices[t][3]) | ((c >> 1) << slices[t][4]);
                if (issl_triple_resp(E) == t) keep |= 1u << c;
            }
            if (exact == 0) keep = 1u;   // sliceWidth 4, no exact byte: triple 0 reports the entries both of whose residual slices differ
            v[i].x = raw[i];
            v[i].y = exact | ((uint32_t)slices[t][3] << 8) | ((uint32_t)slices[t][4] << 12) | (keep << 16);
        }
        CKR(d->visits.ensure(std::max<size_t>(1, v.size()) * sizeof(TripleVisit)));
        CK(cudaMemcpyAsync(d->visits.p, v.data(), v.size() * sizeof(TripleVisit), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // v is a local
        d->visitsDist = maxDist;
    }
    return ISSL_OK;
}

// ISSL_LAYOUT_TRIPLE: survivors of slices [s0, s0 + ns) for the guides that are still active
template <bool FUSED, bool FLUSH>
static void launch_triple_scan_t(const issl_device *d, const TripleArgs &a, dim3 grid, cudaStream_t st)
{
    const bool gates = d->tv.perm10 != 0;   // sliceWidth 10 (issl_triple.cuh)
    if (d->tv.pitch == 32 && gates) k_scan_triple_blocked<1, FUSED, FLUSH, true><<<grid, kTripleThreads, 0, st>>>(a);
    else if (d->tv.pitch == 32) k_scan_triple_blocked<1, FUSED, FLUSH><<<grid, kTripleThreads, 0, st>>>(a);
    else if (d->tv.pitch == 64 && gates) k_scan_triple_blocked<2, FUSED, FLUSH, true><<<grid, kTripleThreads, 0, st>>>(a);
    else if (d->tv.pitch == 64) k_scan_triple_blocked<2, FUSED, FLUSH><<<grid, kTripleThreads, 0, st>>>(a);
    else if (d->tv.pitch == 128 && gates) k_scan_triple_blocked<4, FUSED, FLUSH, true><<<grid, kTripleThreads, 0, st>>>(a);
    else if (d->tv.pitch == 128) k_scan_triple_blocked<4, FUSED, FLUSH><<<grid, kTripleThreads, 0, st>>>(a);
    else k_scan_triple<FUSED><<<grid, kTripleThreads, 0, st>>>(a);
}

// flush: the variant that empties a full record list into the general pipeline's buffer in the middle of the scan
// (one barrier per round of visits) -- chosen when guides are expected to have more hits than a CTA can hold
static void launch_triple_scan(const issl_device *d, const TripleArgs &a, dim3 grid, bool fused, bool flush, cudaStream_t st)
{
    if (fused) { if (flush) launch_triple_scan_t<true, true>(d, a, grid, st); else launch_triple_scan_t<true, false>(d, a, grid, st); }
    else { if (flush) launch_triple_scan_t<false, true>(d, a, grid, st); else launch_triple_scan_t<false, false>(d, a, grid, st); }
}

struct WaveScoring {   // what the fused tail / k_score_segments need to finish the guides
    int fuse;              // 0, 1, 2 as ISSL_TRIPLE_FUSE
    bool calcMit, calcCfd, checkExit;
    int method;
    double maximumSum;
};

static int triple_wave(issl_device *d, cudaStream_t st, const uint64_t *dGuides, uint32_t n, uint32_t s0, uint32_t ns,
                       const uint8_t *doneMask, int maxDist, const WaveScoring &ws, EventTimer &timer, uint64_t *nHitsOut)
{
    unsigned long long *dc = d->counters.as<unsigned long long>();
    CKR(ensure_visits(d, maxDist, st));
    // the visit table is ordered by byte slice; sliceWidth 4 runs it in one piece (score_batch)
    // ... nor does sliceWidth 10, where the slice a hit is met in depends on the guide's gates
    const bool nibble = d->info.sliceWidth != 8;
    const uint32_t v0 = nibble ? d->waveStart[0] : d->waveStart[s0], v1 = nibble ? d->waveStart[5] : d->waveStart[std::min(s0 + ns, 5u)], nv = v1 - v0;
    CK(cudaMemsetAsync(dc, 0, 16 * 8, st));
    k_wave_candidates<<<blocks_for((uint64_t)n * ns, 256), 256, 0, st>>>(d->iv, dGuides, doneMask, n, s0, ns, dc + 0);
    d->stats.launches += 1;
    *nHitsOut = 0;
    if (nv) {
        // enough CTAs to fill the machine even for a handful of guides
        const uint32_t kOctets = d->tv.pitch ? kTripleThreads / (d->tv.pitch / 32) : kTripleThreads / 8;   // visits in flight per CTA
        uint32_t chunks = std::max<uint32_t>(1, (148u * 16u + n - 1) / n);
        chunks = std::min<uint32_t>(chunks, (nv + kOctets - 1) / kOctets);
        chunks = std::min<uint32_t>(chunks, 65535u);
        // finishing a guide where its hits are needs all of them in one CTA
        const bool inScan = ws.fuse == 2 && chunks == 1, fuse = ws.fuse == 1 && chunks == 1;
        // many hits per guide expected (large maxDist; the previous call saw repeat families): flush variant
        // (on average, or -- a genome with a few repeat families -- for a noticeable share of the hits)
        const bool flush = d->tripleFlush == 1 || (d->tripleFlush < 0 && (maxDist >= 5 || d->lastHitsPerGuide > 0.6 * kTripleHitCap ||
                                                                          d->lastHeavyFraction > 0.02));
        if (fuse) { CKR(d->segOff.ensure(n * 8ull)); CKR(d->segCnt.ensure(n * 4ull)); }
        if (inScan) { CKR(d->totMit2.ensure(n * 8ull)); CKR(d->totCfd2.ensure(n * 8ull)); CKR(d->done2.ensure(n)); }
        ScoreParams sp;
        sp.sig = d->iv.sig; sp.occ = d->iv.occ; sp.nSites = d->info.offtargetsCount; sp.occFlag = d->tv.occFlag; sp.tb = score_tables(d);
        // order keys: site text ranks (40 bits) when the index is in text order, site ids otherwise
        uint32_t idShift = 0;
        while (idShift < 28 && (d->info.offtargetsCount >> idShift) > kKeyBuckets) idShift++;
        sp.keyShift = idShift; sp.byFirstMismatch = d->tv.siteOrdered ? 1u : 0u;
        sp.calcMit = ws.calcMit; sp.calcCfd = ws.calcCfd; sp.method = ws.method; sp.checkExit = ws.checkExit;
        sp.maximumSum = ws.maximumSum;
        sp.totMit = d->totMit.as<double>(); sp.totCfd = d->totCfd.as<double>(); sp.done = d->done.as<uint8_t>();
        // heavy guides are finished inside the scan kernel when it runs fused, in its flush variant, on an index in text order
        const bool heavy = inScan && flush && d->tripleHeavy && d->tv.siteOrdered && d->tv.pitch;
        if (heavy) {
            // first guess: what the previous call saw (chunks + the sort's second copy: up to ~3.5 keys per hit), else 2^24 keys
            const double guess = 4.0 * std::max(d->lastHitsPerGuide, 64.0) * n;
            uint64_t want = (uint64_t)std::min(std::max(guess, (double)(1ull << 24)), (double)(1ull << 31));
            if (d->firstHitCap) want = d->firstHitCap;   // ISSL_HIT_CAP: start small, so that tests reach the re-launch after an overflow
            d->heavyCap = std::max<uint64_t>(d->heavyCap, want);
        }
        // a visit table of a few dozen buckets (maxDist <= 3): a warp per guide instead of a CTA per guide
        const bool small = inScan && !flush && d->tripleSmall && nv <= kSmallMaxVisits && d->tv.pitch && d->tv.siteOrdered && !d->tv.perm10;
        if (small) CKR(d->redo.ensure(n * 4ull));
        for (;;) {
            CKR(ensure_hit_buffers(d, n, inScan));
            if (heavy) { CKR(d->heavyKeys.ensure(d->heavyCap * 8)); CKR(d->heavyDesc.ensure((size_t)n * sizeof(HeavyDesc))); }
            if (fuse && d->segCap < d->hitCap) {
                d->segCap = d->hitCap;
                CKR(d->segKeys.ensure(d->segCap * 8)); CKR(d->segSites.ensure(d->segCap * 8));
            }
            CK(cudaMemsetAsync(dc + 1, 0, 8, st));
            CK(cudaMemsetAsync(dc + 4, 0, 72, st));
            // overflow bitmap of the non-flush scan (issl_triple.cuh): a row per CTA
            const uint32_t visitsPerCta = (nv + chunks - 1) / chunks, ovfWords = (visitsPerCta + 31) / 32;
            const bool ovfBitmap = d->tv.pitch && !flush && visitsPerCta <= 65536;
            if (ovfBitmap) {
                CKR(d->ovfBits.ensure((size_t)n * chunks * ovfWords * 4));
                CK(cudaMemsetAsync(d->ovfBits.p, 0, (size_t)n * chunks * ovfWords * 4, st));
            }
            if (fuse) CK(cudaMemsetAsync(d->segCnt.p, 0, n * 4ull, st));
            TripleArgs a;
            a.tv = d->tv; a.guides = dGuides; a.done = doneMask; a.visits = d->visits.as<TripleVisit>() + v0;
            a.nVisits = nv; a.visitsPerCta = (nv + chunks - 1) / chunks;
            a.hitKeys = d->keysA.as<uint64_t>(); a.hitCount = dc + 1; a.hitCap = d->hitCap; a.streamed = dc + 4;
            a.maxDist = maxDist;
            a.segKeys = d->segKeys.as<uint64_t>(); a.segSites = d->segSites.as<uint64_t>(); a.segCount = dc + 6; a.segCap = d->segCap;
            a.segOff = fuse ? d->segOff.as<uint64_t>() : nullptr; a.segCnt = fuse ? d->segCnt.as<uint32_t>() : nullptr;
            a.fuse = inScan ? 1 : 0; a.sp = sp;
            a.totMitOut = d->totMit2.as<double>(); a.totCfdOut = d->totCfd2.as<double>(); a.doneOut = d->done2.as<uint8_t>();
            a.fusedHits = dc + 7; a.maxRecords = dc + 2;
            a.heavyKeys = heavy ? d->heavyKeys.as<uint64_t>() : nullptr; a.heavyCount = dc + 8; a.heavyCap = heavy ? d->heavyCap : 0;
            a.heavyHits = dc + 9; a.heavyDesc = heavy ? d->heavyDesc.as<HeavyDesc>() : nullptr; a.heavyGuides = dc + 11;
            a.ovfBits = ovfBitmap ? d->ovfBits.as<uint32_t>() : nullptr; a.ovfWords = ovfWords;
            a.nGuides = n; a.redo = small ? d->redo.as<uint32_t>() : nullptr; a.redoCount = dc + 10; a.guideList = nullptr;
            cudaEvent_t e0, e1;
            CKR(timer.get(&e0)); CKR(timer.get(&e1));
            timer.scanPairs.push_back({timer.used - 2, timer.used - 1});
            CK(cudaEventRecord(e0, st));
            if (small) {
                const dim3 sgrid((n + kTripleThreads / 32 - 1) / (kTripleThreads / 32));
                if (d->tv.pitch == 32) k_scan_triple_small<1><<<sgrid, kTripleThreads, 0, st>>>(a);
                else if (d->tv.pitch == 64) k_scan_triple_small<2><<<sgrid, kTripleThreads, 0, st>>>(a);
                else k_scan_triple_small<4><<<sgrid, kTripleThreads, 0, st>>>(a);
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(d->hCounters + 10, dc + 10, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                d->stats.scan_launches += 1;
                d->stats.launches += 1;
                if (d->hCounters[10]) {   // the guides it left: repeat families, buckets that overflow their block
                    a.guideList = d->redo.as<uint32_t>();
                    launch_triple_scan(d, a, dim3((unsigned)d->hCounters[10], chunks), inScan, flush, st);
                }
            } else {
                launch_triple_scan(d, a, dim3(n, chunks), inScan, flush, st);
            }
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1, st));
            CK(cudaMemcpyAsync(d->hCounters, dc, 16 * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (!small || d->hCounters[10]) { d->stats.scan_launches += 1; d->stats.launches += 1; }
            if (d->hCounters[1] <= d->hitCap && d->hCounters[6] <= d->segCap && (!heavy || d->hCounters[8] <= d->heavyCap)) break;
            if (heavy && d->hCounters[8] > d->heavyCap) d->heavyCap = d->hCounters[8] + d->hCounters[8] / 4;
            if (d->hCounters[1] > d->hitCap) {
                d->hitCap = d->hCounters[1] + d->hCounters[1] / 4;
                CKR(d->keysA.ensure(d->hitCap * 8)); CKR(d->keysB.ensure(d->hitCap * 8));
            }
            if (d->hCounters[6] > d->segCap) {
                d->segCap = d->hCounters[6] + d->hCounters[6] / 4;
                CKR(d->segKeys.ensure(d->segCap * 8)); CKR(d->segSites.ensure(d->segCap * 8));
            }
        }
        *nHitsOut = d->hCounters[1];
        if (heavy && d->hCounters[11]) {
            // the guides whose hits did not fit their CTA's record list: sorted and finished by a CTA each (k_heavy_finish)
            CKR(d->heavyFlat.ensure(d->hCounters[9] * 8));
            HeavyArgs ha;
            ha.desc = d->heavyDesc.as<HeavyDesc>(); ha.keys = d->heavyKeys.as<uint64_t>(); ha.flat = d->heavyFlat.as<uint64_t>();
            ha.flatCount = dc + 12; ha.guides = dGuides; ha.sp = sp; ha.fineGroups = d->tv.nibbleOrder ? 0u : 1u;
            ha.totMitOut = d->totMit2.as<double>(); ha.totCfdOut = d->totCfd2.as<double>(); ha.doneOut = d->done2.as<uint8_t>();
            cudaEvent_t h0, h1;
            CKR(timer.get(&h0)); CKR(timer.get(&h1));
            timer.heavyPairs.push_back({timer.used - 2, timer.used - 1});
            CK(cudaEventRecord(h0, st));
            k_heavy_finish<<<(unsigned)d->hCounters[11], kTripleThreads, 0, st>>>(ha);
            CK(cudaGetLastError());
            CK(cudaEventRecord(h1, st));
            d->stats.launches += 1;
        }
        if (getenv("ISSL_DEBUG")) fprintf(stderr, "[issl] wave %u+%u: max records per guide %llu, general-pipeline hits %llu, fused hits %llu, heavy keys %llu\n",
                                          s0, ns, d->hCounters[2], d->hCounters[1], d->hCounters[7], d->hCounters[8]);
        d->stats.streamed += d->hCounters[4];
        d->stats.bucket_visits += d->hCounters[5];
        if (fuse && d->hCounters[6]) {
            // guides whose hits fit a segment are finished there; nHitsOut covers only the others
            SegmentArgs sa;
            sa.segKeys = d->segKeys.as<uint64_t>(); sa.segSites = d->segSites.as<uint64_t>();
            sa.segOff = d->segOff.as<uint64_t>(); sa.segCnt = d->segCnt.as<uint32_t>();
            sa.guides = dGuides; sa.sp = sp;
            sa.sp.keyShift = idShift; sa.sp.byFirstMismatch = 0;   // segments are ordered by id
            k_score_segments<<<n, kTripleThreads, 0, st>>>(sa);
            CK(cudaGetLastError());
            d->stats.launches += 1;
            d->stats.hits += d->hCounters[6];
        }
        if (inScan) {   // the scan kernel wrote every guide's state of after this wave
            d->totMit.swap(d->totMit2); d->totCfd.swap(d->totCfd2); d->done.swap(d->done2);
            d->stats.hits += d->hCounters[7];
            d->stats.heavy_hits += d->hCounters[9];
        }
    } else {
        CK(cudaMemcpyAsync(d->hCounters, dc, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    d->stats.candidates += d->hCounters[0];
    return ISSL_OK;
}

// List scan (RES32 / SIG64 / GATHER): survivors of slices [s0, s0 + ns) for the guides that are still active.
// ref isslScoreOfftargets.cpp:330-390 for every (guide, slice) pair of the wave.
static int list_wave(issl_device *d, cudaStream_t st, const uint64_t *dGuides, uint32_t n, uint32_t s0, uint32_t ns,
                     const uint8_t *doneMask, int maxDist, EventTimer &timer, uint64_t *nHitsOut)
{
    unsigned long long *dc = d->counters.as<unsigned long long>();
    const uint64_t pairs = (uint64_t)n * ns;
    *nHitsOut = 0;
    // group the (guide, slice) pairs by the list they select: radix sort of (list id, guide index)
    const uint32_t nLists = (uint32_t)d->nLists;
    CKR(d->pairKeys.ensure(pairs * 4)); CKR(d->pairVals.ensure(pairs * 4));
    CKR(d->pairKeysSorted.ensure(pairs * 4)); CKR(d->pairValsSorted.ensure(pairs * 4));
    CK(cudaMemsetAsync(dc, 0, 16 * 8, st));
    k_pair_keys<<<blocks_for(pairs, 256), 256, 0, st>>>(d->iv, dGuides, doneMask, n, s0, ns, nLists,
                                                       d->pairKeys.as<uint32_t>(), d->pairVals.as<uint32_t>(), dc + 0);
    int keyBits = 1;
    while ((1ull << keyBits) <= nLists) keyBits++;
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d->pairKeys.as<uint32_t>(), d->pairKeysSorted.as<uint32_t>(),
                                       d->pairVals.as<uint32_t>(), d->pairValsSorted.as<uint32_t>(), pairs, 0, keyBits, st));
    CKR(d->sortTemp.ensure(tb));
    CK(cub::DeviceRadixSort::SortPairs(d->sortTemp.p, tb, d->pairKeys.as<uint32_t>(), d->pairKeysSorted.as<uint32_t>(),
                                       d->pairVals.as<uint32_t>(), d->pairValsSorted.as<uint32_t>(), pairs, 0, keyBits, st));
    CK(cudaMemcpyAsync(d->hCounters, dc, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    d->stats.launches += 1 + (uint64_t)((keyBits + 7) / 8) + 2;
    const uint64_t candidates = d->hCounters[0];
    d->stats.candidates += candidates;
    if (candidates == 0) return ISSL_OK;

    // chunk: aim at a few hundred thousand items at most, never below two quanta
    uint64_t chunk64 = (candidates / (1ull << 19) + kChunkQuantum - 1) / kChunkQuantum * kChunkQuantum;
    chunk64 = std::max<uint64_t>(chunk64, 2 * kChunkQuantum);
    chunk64 = std::min<uint64_t>(chunk64, 1ull << 30);
    const uint32_t chunk = (uint32_t)chunk64;

    uint32_t maxGroup = d->maxGroup;
    if (maxGroup > kMaxGroup && !(d->iv.layout == ISSL_LAYOUT_RES32 && maxDist <= 7)) maxGroup = kMaxGroup;
    CKR(d->pairCounts.ensure((pairs + 1) * 4)); CKR(d->pairOffsets.ensure((pairs + 1) * 4));
    k_group_count<<<blocks_for(pairs, 256), 256, 0, st>>>(d->iv, d->pairKeysSorted.as<uint32_t>(), (uint32_t)pairs, nLists, chunk,
                                                         maxGroup, d->pairCounts.as<uint32_t>(), dc + 4);
    CK(cudaMemsetAsync(d->pairCounts.as<uint32_t>() + pairs, 0, 4, st));
    tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d->pairCounts.as<uint32_t>(), d->pairOffsets.as<uint32_t>(), pairs + 1, st));
    CKR(d->scanTemp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(d->scanTemp.p, tb, d->pairCounts.as<uint32_t>(), d->pairOffsets.as<uint32_t>(), pairs + 1, st));
    CK(cudaMemcpyAsync(d->hCounters + 2, d->pairOffsets.as<uint32_t>() + pairs, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(d->hCounters + 4, dc + 4, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t nItems = *reinterpret_cast<uint32_t *>(d->hCounters + 2);
    d->stats.streamed += d->hCounters[4];
    CKR(d->items.ensure((size_t)nItems * sizeof(ScanItem)));
    k_group_fill<<<blocks_for(pairs, 256), 256, 0, st>>>(d->iv, d->pairKeysSorted.as<uint32_t>(), (uint32_t)pairs, nLists, chunk,
                                                        maxGroup, d->pairOffsets.as<uint32_t>(), d->items.as<ScanItem>());
    d->stats.launches += 4;

    // K1 (re-run with a larger survivor buffer if it overflowed)
    for (;;) {
        CKR(ensure_hit_buffers(d, n));
        CK(cudaMemsetAsync(dc + 1, 0, 8, st));
        ScanArgs a;
        a.iv = d->iv; a.items = d->items.as<ScanItem>(); a.guides = dGuides;
        a.sortedGuide = d->pairValsSorted.as<uint32_t>(); a.hitKeys = d->keysA.as<uint64_t>();
        a.hitCount = dc + 1; a.hitCap = d->hitCap; a.maxDist = maxDist; a.pbits = d->pbits;
        cudaEvent_t e0, e1;
        CKR(timer.get(&e0)); CKR(timer.get(&e1));
        timer.scanPairs.push_back({timer.used - 2, timer.used - 1});
        CK(cudaEventRecord(e0, st));
        if (d->iv.layout == ISSL_LAYOUT_RES32) k_scan<kRes32><<<nItems, kScanThreads, 0, st>>>(a);
        else if (d->iv.layout == ISSL_LAYOUT_SIG64) k_scan<kSig64><<<nItems, kScanThreads, 0, st>>>(a);
        else k_scan<kGather><<<nItems, kScanThreads, 0, st>>>(a);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, st));
        CK(cudaMemcpyAsync(d->hCounters + 1, dc + 1, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        d->stats.scan_launches += 1;
        d->stats.launches += 1;
        *nHitsOut = d->hCounters[1];
        if (*nHitsOut <= d->hitCap) break;
        d->hitCap = *nHitsOut + *nHitsOut / 4;
        CKR(d->keysA.ensure(d->hitCap * 8)); CKR(d->keysB.ensure(d->hitCap * 8));
    }
    return ISSL_OK;
}

// one batch of <= maxBatch guides, already resident at dGuides
static int score_batch(issl_device *d, cudaStream_t st, const uint64_t *dGuides, uint32_t n, uint64_t guideBase, int maxDist,
                       double threshold, int method, double *dMit, double *dCfd, HitSink *sink, EventTimer &timer)
{
    const bool calcMit = method == ISSL_METHOD_MIT || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    const bool calcCfd = method == ISSL_METHOD_CFD || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    if (!calcMit && !calcCfd) return ISSL_OK;   // unknown method: nothing is scored, both columns print as -1
    if (const char *e = getenv("ISSL_TEST_NOMEM_ABOVE"))   // test hook: batches above this size "run out of memory"
        if (n > (uint32_t)atol(e)) return issl_set_error(ISSL_ERR_NOMEM, "cudaMalloc: out of memory (simulated by ISSL_TEST_NOMEM_ABOVE)");

    const double maximumSum = (10000.0 - threshold * 100) / threshold;   // ref :326
    const bool checkExit = !(std::isnan(maximumSum) || (std::isinf(maximumSum) && maximumSum > 0));
    const uint32_t S = d->iv.sliceCount;

    CKR(d->totMit.ensure(n * 8ull)); CKR(d->totCfd.ensure(n * 8ull)); CKR(d->done.ensure(n));
    CKR(d->counters.ensure(16 * 8));
    CK(cudaMemsetAsync(d->totMit.p, 0, n * 8ull, st));
    CK(cudaMemsetAsync(d->totCfd.p, 0, n * 8ull, st));
    CK(cudaMemsetAsync(d->done.p, 0, n, st));
    if (sink) { CKR(d->scoredEnd.ensure(n * 8ull)); CKR(d->segBegin.ensure(n * 8ull)); }
    unsigned long long *dc = d->counters.as<unsigned long long>();
    std::vector<uint64_t> hEnd, hBegin, hKeys;
    std::vector<uint32_t> hId, hOcc;
    std::vector<int32_t> hDist;

    // without early exit all slices go in one wave; with it, one wave per slice so that guides
    // which exited stop generating work (ref :501-502)
    // sliceWidth 4 under TRIPLE: the visit table's waves are by byte, not by the reference's 2-base slices (and from maxDist 5
    // on a hit need not match on any whole byte: issl_triple_visits_w4): one wave, the early exit takes effect in the ordered
    // accumulation
    // sliceWidth 10 under TRIPLE: the slice a hit is met in depends on the guide's gates (fifth base of a slice = A), not on
    // the visit: one wave as well
    const bool nibble = d->layout == ISSL_LAYOUT_TRIPLE && d->info.sliceWidth == 4;
    const bool gates = d->layout == ISSL_LAYOUT_TRIPLE && d->info.sliceWidth == 10;
    const bool useTriple = d->layout == ISSL_LAYOUT_TRIPLE && maxDist >= 0 && maxDist <= d->tripleMaxDist;
    if (!useTriple) CKR(ensure_lists(d));   // (TRIPLE builds its slice lists the first time a call needs them)
    // Waves: one slice at a time pays when many guides leave through the early exit (repeat families: their later slices
    // are never scanned); when few do, every further launch only costs.  So after each single-slice wave the guides that
    // left are counted, and once a wave sends fewer than a tenth of the batch through the exit all remaining slices go in
    // one launch -- the ordered accumulation reproduces the exit points either way.  (ISSL_WAVES: 0 = one launch,
    // 1 = adaptive, 2 = always one slice per wave.)
    //   The previous call on the handle is a good predictor: if fewer than a third of its guides left early, this one
    // starts as one launch right away.
    const bool oneWave = !checkExit || (useTriple && (nibble || gates)) || d->waves == 0 ||
                         (d->waves == 1 && d->lastExitFraction >= 0.0 && d->lastExitFraction < 0.33);
    uint64_t doneBefore = 0;
    bool merged = false;
    for (uint32_t s0 = 0, ns = 0; s0 < S; s0 += ns) {
        ns = (oneWave || merged) ? S - s0 : 1;
        if (!oneWave && !merged && d->waves == 1 && s0 > 0) {
            CK(cudaMemsetAsync(dc + 3, 0, 8, st));
            k_count_done<<<blocks_for(n, 256), 256, 0, st>>>(d->done.as<uint8_t>(), n, dc + 3);
            CK(cudaMemcpyAsync(d->hCounters + 3, dc + 3, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            d->stats.launches += 1;
            if (d->hCounters[3] - doneBefore < n / 10) { merged = true; ns = S - s0; }
            doneBefore = d->hCounters[3];
        }
        const uint8_t *doneMask = checkExit ? d->done.as<uint8_t>() : nullptr;

        const int posBits = useTriple ? kTripleKeyBits : d->pbits;
        uint64_t nHits = 0;
        if (useTriple) {
            WaveScoring ws;
            ws.fuse = sink == nullptr ? d->tripleFuse : 0; ws.calcMit = calcMit; ws.calcCfd = calcCfd; ws.checkExit = checkExit;
            ws.method = method; ws.maximumSum = maximumSum;
            CKR(triple_wave(d, st, dGuides, n, s0, ns, doneMask, maxDist, ws, timer, &nHits));
        } else {
            CKR(list_wave(d, st, dGuides, n, s0, ns, doneMask, maxDist, timer, &nHits));
        }
        if (nHits == 0) continue;

        if (useTriple && nibble) {   // sliceWidth 4: the keys carry the lowest exact byte; the reference orders by 2-base slice
            k_fix_order_slices<<<blocks_for(nHits, 256), 256, 0, st>>>(d->keysA.as<uint64_t>(), nHits, dGuides, d->sig.as<uint64_t>());
            d->stats.launches += 1;
        }
        // canonical order: sort keys (guide, position)
        int gbits = 1;
        while ((1ull << gbits) < n) gbits++;
        cub::DoubleBuffer<uint64_t> db(d->keysA.as<uint64_t>(), d->keysB.as<uint64_t>());
        size_t tb = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, nHits, 0, posBits + gbits, st));
        CKR(d->sortTemp.ensure(tb));
        CK(cub::DeviceRadixSort::SortKeys(d->sortTemp.p, tb, db, nHits, 0, posBits + gbits, st));
        const uint64_t *sorted = db.Current();
        d->stats.launches += (uint64_t)((posBits + gbits + 7) / 8) * 2 + 1;   // histogram + onesweep passes (approximate)

        CKR(d->contribMit.ensure(nHits * 8)); CKR(d->contribCfd.ensure(nHits * 8));
        if (sink) { CKR(d->hitId.ensure(nHits * 4)); CKR(d->hitDist.ensure(nHits * 4)); CKR(d->hitOcc.ensure(nHits * 4)); }
        ContribArgs c;
        c.iv = d->iv; c.keys = sorted; c.nHits = nHits; c.guides = dGuides;
        c.tb = score_tables(d);
        c.pbits = posBits; c.idInKey = useTriple ? 1 : 0; c.calcMit = calcMit; c.calcCfd = calcCfd;
        c.contribMit = d->contribMit.as<double>(); c.contribCfd = d->contribCfd.as<double>();
        c.hitId = sink ? d->hitId.as<uint32_t>() : nullptr;
        c.hitDist = sink ? d->hitDist.as<int32_t>() : nullptr;
        c.hitOcc = sink ? d->hitOcc.as<uint32_t>() : nullptr;
        k_contrib<<<blocks_for(nHits, 256), 256, 0, st>>>(c);

        AccumArgs ac;
        ac.keys = sorted; ac.nHits = nHits; ac.contribMit = c.contribMit; ac.contribCfd = c.contribCfd;
        ac.nGuides = n; ac.pbits = posBits; ac.method = method; ac.checkExit = checkExit; ac.maximumSum = maximumSum;
        ac.totMit = d->totMit.as<double>(); ac.totCfd = d->totCfd.as<double>(); ac.done = d->done.as<uint8_t>();
        ac.scoredEnd = sink ? d->scoredEnd.as<uint64_t>() : nullptr;
        ac.segBegin = sink ? d->segBegin.as<uint64_t>() : nullptr;
        k_accumulate<<<blocks_for(n, 128), 128, 0, st>>>(ac);
        CK(cudaGetLastError());
        d->stats.launches += 2;

        if (sink) {
            hEnd.resize(n); hBegin.resize(n); hKeys.resize(nHits); hId.resize(nHits); hOcc.resize(nHits); hDist.resize(nHits);
            CK(cudaMemcpyAsync(hEnd.data(), d->scoredEnd.p, n * 8ull, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(hBegin.data(), d->segBegin.p, n * 8ull, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(hId.data(), d->hitId.p, nHits * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(hOcc.data(), d->hitOcc.p, nHits * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(hDist.data(), d->hitDist.p, nHits * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (uint32_t g = 0; g < n; g++)
                for (uint64_t j = hBegin[g]; j < hEnd[g]; j++) {
                    sink->guide.push_back(guideBase + g);
                    sink->id.push_back(hId[j]); sink->dist.push_back(hDist[j]); sink->occ.push_back(hOcc[j]);
                }
            d->stats.hits += 0;
        }
        d->stats.hits += nHits;
        d->stats.sorted_hits += nHits;
    }

    if (checkExit) {
        CK(cudaMemsetAsync(dc + 3, 0, 8, st));
        k_count_done<<<blocks_for(n, 256), 256, 0, st>>>(d->done.as<uint8_t>(), n, dc + 3);
        CK(cudaMemcpyAsync(d->hCounters + 3, dc + 3, 8, cudaMemcpyDeviceToHost, st));
        d->stats.launches += 1;
    }
    k_finalize<<<blocks_for(n, 256), 256, 0, st>>>(d->totMit.as<double>(), d->totCfd.as<double>(), n, calcMit ? dMit : nullptr,
                                                  calcCfd ? dCfd : nullptr);
    CK(cudaGetLastError());
    d->stats.launches += 1;
    CK(cudaStreamSynchronize(st));
    if (checkExit) d->stats.early_exits += d->hCounters[3];
    return ISSL_OK;
}

static int score_common(issl_device *d, const uint64_t *guides, bool guidesOnDevice, size_t n, int maxDist, double threshold,
                        int method, double *mitOut, double *cfdOut, bool outOnDevice, cudaStream_t st, HitSink *sink)
{
    if (!d) return issl_set_error(ISSL_ERR_ARG, "issl_score: null device handle");
    if (n && !guides) return issl_set_error(ISSL_ERR_ARG, "issl_score: null guide array");
    CK(cudaSetDevice(d->dev));
    if (!st) st = d->stream;
    d->stats = issl_stats{};
    d->stats.guides = n;
    EventTimer timer{d};
    cudaEvent_t t0, t1;
    CKR(timer.get(&t0)); CKR(timer.get(&t1));
    CK(cudaEventRecord(t0, st));

    const bool calcMit = method == ISSL_METHOD_MIT || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    const bool calcCfd = method == ISSL_METHOD_CFD || method == ISSL_METHOD_AND || method == ISSL_METHOD_OR || method == ISSL_METHOD_AVG;
    if ((calcMit && !mitOut) || (calcCfd && !cfdOut))
        return issl_set_error(ISSL_ERR_ARG, "issl_score: output array missing for a column the method computes");

    // A batch's scratch grows with its hits (general pipeline: two key buffers and two contribution arrays, 32 B per
    // hit): at maxDist 5-6, or on repeat-rich genomes, 2^20 guides can ask for more than is left beside the index.  The
    // batch is bounded by what the previous call saw, and a batch that still runs out of memory is halved and repeated.
    uint32_t batch = d->maxBatch;
    // (cudaMemGetInfo takes milliseconds on a GPU this size: only asked when the call is large enough for the bound to matter)
    if (d->lastHitsPerGuide > 512.0 && (double)n * d->lastHitsPerGuide * 40.0 > 8e9) {
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            const double held = (double)d->keysA.cap + (double)d->keysB.cap + (double)d->contribMit.cap + (double)d->contribCfd.cap + (double)d->heavyKeys.cap;
            const double fit = 0.6 * ((double)freeB + held) / (40.0 * d->lastHitsPerGuide);
            if (fit < (double)batch) batch = (uint32_t)std::max(4096.0, fit);
        }
    }
    for (size_t b0 = 0; b0 < n;) {
        const uint32_t nb = (uint32_t)std::min<size_t>(batch, n - b0);
        const uint64_t *dG = guides + b0;
        if (!guidesOnDevice) {
            CKR(d->guides.ensure(nb * 8ull));
            CK(cudaMemcpyAsync(d->guides.p, guides + b0, nb * 8ull, cudaMemcpyHostToDevice, st));
            dG = d->guides.as<uint64_t>();
        }
        double *dM = mitOut ? mitOut + b0 : nullptr, *dC = cfdOut ? cfdOut + b0 : nullptr;
        if (!outOnDevice) {
            CKR(d->outMit.ensure(nb * 8ull)); CKR(d->outCfd.ensure(nb * 8ull));
            dM = d->outMit.as<double>(); dC = d->outCfd.as<double>();
        }
        const issl_stats before = d->stats;
        const size_t hitsBefore = sink ? sink->guide.size() : 0;
        const int rc = score_batch(d, st, dG, nb, b0, maxDist, threshold, method, dM, dC, sink, timer);
        if (rc == ISSL_ERR_NOMEM && nb > 1024) {
            cudaGetLastError();
            cudaStreamSynchronize(st);
            for (DBuf *b : {&d->keysA, &d->keysB, &d->contribMit, &d->contribCfd, &d->sortTemp, &d->segKeys, &d->segSites, &d->hitId, &d->hitDist, &d->hitOcc,
                            &d->heavyKeys, &d->heavyFlat})
                b->release();
            d->hitCap = d->hitCapAuto = 0; d->segCap = 0; d->heavyCap = 0;
            d->stats = before;
            if (sink) { sink->guide.resize(hitsBefore); sink->id.resize(hitsBefore); sink->dist.resize(hitsBefore); sink->occ.resize(hitsBefore); }
            batch = std::max<uint32_t>(1024, nb / 2);
            if (getenv("ISSL_DEBUG")) fprintf(stderr, "[issl] batch of %u guides ran out of device memory: retrying with %u\n", nb, batch);
            continue;
        }
        CKR(rc);
        if (!outOnDevice) {
            if (calcMit) CK(cudaMemcpyAsync(mitOut + b0, dM, nb * 8ull, cudaMemcpyDeviceToHost, st));
            if (calcCfd) CK(cudaMemcpyAsync(cfdOut + b0, dC, nb * 8ull, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        b0 += nb;
    }
    CK(cudaEventRecord(t1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    d->stats.total_ms = ms;
    d->lastHitsPerGuide = n ? (double)d->stats.hits / (double)n : 0.0;
    d->lastHeavyFraction = d->stats.hits ? (double)(d->stats.heavy_hits + d->stats.sorted_hits) / (double)d->stats.hits : 0.0;
    {
        const double maximumSum = (10000.0 - threshold * 100) / threshold;
        if (n && !(std::isnan(maximumSum) || (std::isinf(maximumSum) && maximumSum > 0)))
            d->lastExitFraction = (double)d->stats.early_exits / (double)n;
    }
    for (auto &pr : timer.scanPairs) {
        CK(cudaEventElapsedTime(&ms, d->evPool[pr.first], d->evPool[pr.second]));
        d->stats.scan_ms += ms;
    }
    for (auto &pr : timer.heavyPairs) {
        CK(cudaEventElapsedTime(&ms, d->evPool[pr.first], d->evPool[pr.second]));
        d->stats.heavy_ms += ms;
    }
    return ISSL_OK;
}

extern "C" int issl_score(issl_device *d, const uint64_t *guides, size_t n, int maxDist, double threshold, int method,
                          double *mit_out, double *cfd_out)
{
    return score_common(d, guides, false, n, maxDist, threshold, method, mit_out, cfd_out, false, nullptr, nullptr);
}

extern "C" int issl_score_device(issl_device *d, const uint64_t *d_guides, size_t n, int maxDist, double threshold,
                                 int method, double *d_mit_out, double *d_cfd_out, void *stream)
{
    return score_common(d, d_guides, true, n, maxDist, threshold, method, d_mit_out, d_cfd_out, true,
                        static_cast<cudaStream_t>(stream), nullptr);
}

extern "C" int issl_score_hits(issl_device *d, const uint64_t *guides, size_t n, int maxDist, double threshold, int method,
                               double *mit_out, double *cfd_out, uint64_t *hit_guide, uint32_t *hit_id, int32_t *hit_dist,
                               uint32_t *hit_occ, size_t cap, size_t *count)
{
    HitSink sink;
    CKR(score_common(d, guides, false, n, maxDist, threshold, method, mit_out, cfd_out, false, nullptr, &sink));
    // waves append per slice; the reference's order is per guide, then slice, then list position
    std::vector<size_t> order(sink.guide.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return sink.guide[a] < sink.guide[b]; });
    if (count) *count = order.size();
    for (size_t i = 0; i < order.size() && i < cap; i++) {
        const size_t k = order[i];
        if (hit_guide) hit_guide[i] = sink.guide[k];
        if (hit_id) hit_id[i] = sink.id[k];
        if (hit_dist) hit_dist[i] = sink.dist[k];
        if (hit_occ) hit_occ[i] = sink.occ[k];
    }
    return ISSL_OK;
}

extern "C" int issl_guide_filters(issl_device *d, const char *text, size_t bytes, uint8_t *flags_out, double *at_out,
                                  uint64_t *packed_out)
{
    if (!d || (bytes && !text)) return issl_set_error(ISSL_ERR_ARG, "issl_guide_filters: null argument");
    if (bytes % 24 != 0) return issl_set_error(ISSL_ERR_ARG, "issl_guide_filters: input is not a multiple of 24 bytes (23 characters + LF)");
    const uint64_t n = bytes / 24;
    if (n == 0) return ISSL_OK;
    CK(cudaSetDevice(d->dev));
    DBuf dText, dFlags, dAt, dPacked;
    int rc = dText.ensure(bytes);
    if (rc == ISSL_OK && flags_out) rc = dFlags.ensure(n);
    if (rc == ISSL_OK && at_out) rc = dAt.ensure(n * 8);
    if (rc == ISSL_OK && packed_out) rc = dPacked.ensure(n * 8);
    cudaError_t e = cudaSuccess;
    if (rc == ISSL_OK) {
        e = cudaMemcpyAsync(dText.p, text, bytes, cudaMemcpyHostToDevice, d->stream);
        k_guide_filters<<<blocks_for(n, 256), 256, 0, d->stream>>>(dText.as<char>(), n, flags_out ? dFlags.as<uint8_t>() : nullptr,
                                                                  at_out ? dAt.as<double>() : nullptr,
                                                                  packed_out ? dPacked.as<uint64_t>() : nullptr);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess && flags_out) e = cudaMemcpyAsync(flags_out, dFlags.p, n, cudaMemcpyDeviceToHost, d->stream);
        if (e == cudaSuccess && at_out) e = cudaMemcpyAsync(at_out, dAt.p, n * 8, cudaMemcpyDeviceToHost, d->stream);
        if (e == cudaSuccess && packed_out) e = cudaMemcpyAsync(packed_out, dPacked.p, n * 8, cudaMemcpyDeviceToHost, d->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    }
    for (DBuf *b : {&dText, &dFlags, &dAt, &dPacked}) b->release();
    if (rc != ISSL_OK) return rc;
    if (e != cudaSuccess) return issl_set_error(ISSL_ERR_CUDA, "issl_guide_filters: %s", cudaGetErrorString(e));
    return ISSL_OK;
}

extern "C" int issl_guide_duplicates(issl_device *d, const char *text, size_t bytes, uint8_t *flags_out, uint64_t *n_later,
                                     uint64_t *n_sequences)
{
    if (!d || !flags_out || (bytes && !text)) return issl_set_error(ISSL_ERR_ARG, "issl_guide_duplicates: null argument");
    if (bytes % 24 != 0) return issl_set_error(ISSL_ERR_ARG, "issl_guide_duplicates: input is not a multiple of 24 bytes (23 characters + LF)");
    const uint64_t n = bytes / 24;
    if (n_later) *n_later = 0;
    if (n_sequences) *n_sequences = 0;
    if (n == 0) return ISSL_OK;
    if (n >= (1ull << 32)) return issl_set_error(ISSL_ERR_ARG, "issl_guide_duplicates: more than 2^32 - 1 targets in one call");
    CK(cudaSetDevice(d->dev));
    cudaStream_t st = d->stream;
    DBuf dText, keys, keysOut, idx, idxOut, tmp, dFlags;
    CKR(dText.ensure(bytes)); CKR(keys.ensure(n * 8)); CKR(keysOut.ensure(n * 8)); CKR(idx.ensure(n * 4)); CKR(idxOut.ensure(n * 4));
    CKR(dFlags.ensure(n)); CKR(d->counters.ensure(16 * 8));
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint64_t>(), keysOut.as<uint64_t>(), idx.as<uint32_t>(), idxOut.as<uint32_t>(), n, 0, 46, st));
    CKR(tmp.ensure(tb));
    unsigned long long *dc = d->counters.as<unsigned long long>();
    CK(cudaMemcpyAsync(dText.p, text, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dc, 0, 16, st));
    k_pack_targets<<<blocks_for(n, 256), 256, 0, st>>>(dText.as<char>(), n, keys.as<uint64_t>(), idx.as<uint32_t>());
    // least-significant-digit radix sort: stable, so inside a run of equal keys positions ascend
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<uint64_t>(), keysOut.as<uint64_t>(), idx.as<uint32_t>(), idxOut.as<uint32_t>(), n, 0, 46, st));
    k_mark_duplicates<<<blocks_for(n, 256), 256, 0, st>>>(keysOut.as<uint64_t>(), idxOut.as<uint32_t>(), n, dFlags.as<uint8_t>(), dc);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flags_out, dFlags.p, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(d->hCounters, dc, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_later) *n_later = d->hCounters[0];
    if (n_sequences) *n_sequences = d->hCounters[1];
    return ISSL_OK;
}

// pinned host memory every device can copy from / into directly (cudaHostAllocPortable): what a host program should
// hold its guide and score arrays in when it drives several GPUs (pageable memory is staged by the driver)
extern "C" int issl_host_alloc(size_t bytes, void **out)
{
    if (!out) return issl_set_error(ISSL_ERR_ARG, "issl_host_alloc: null argument");
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return issl_set_error(e == cudaErrorMemoryAllocation ? ISSL_ERR_NOMEM : ISSL_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return ISSL_OK;
}

extern "C" void issl_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int issl_last_stats(const issl_device *d, issl_stats *out)
{
    if (!d || !out) return issl_set_error(ISSL_ERR_ARG, "issl_last_stats: null argument");
    *out = d->stats;
    return ISSL_OK;
}
