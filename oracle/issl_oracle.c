/*
 * issl_oracle.c -- CPU restatement of Crackling's ISSL off-target scorer and index
 * builder.  TEST INFRASTRUCTURE ONLY: this file is the checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libissl_cuda, the isslScoreOfftargets host program)
 * never links, imports or executes anything under oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so the pin is the reference itself: oracle/Makefile compiles
 * the unmodified reference sources into oracle/_ref/, tests/golden/make_golden.py ran
 * those binaries in the build container and committed their inputs/outputs under
 * tests/golden/, and tests/test_oracle_golden.py checks this restatement against them
 * byte for byte (stdout, .issl images) and tuple for tuple (hit sets).
 *
 * Every function cites the reference lines it follows; paths are relative to
 * /root/reference/src/ISSL/.  The algorithm is deliberately the reference's own
 * (8-byte list entries, signature gather, per-thread toggle bitset cleared per guide,
 * strictly ordered accumulation with early exit) and NOT the layout or the
 * de-duplication rule the CUDA path uses, so agreement between the two is evidence.
 *
 * Plain C11, gcc -O3 -fopenmp -mpopcnt (no -mfma / -ffast-math: each `tot += s*occ` is
 * a rounded multiply followed by a rounded add, as in the reference build, Makefile:5).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cfd_table.inc"

/* method codes = the reference's enum ScoreMethod, isslScoreOfftargets.cpp:44 */
enum { M_UNKNOWN = 0, M_MIT = 1, M_CFD = 2, M_AND = 3, M_OR = 4, M_AVG = 5 };

/* ------------------------------------------------------------------------------------------
 * 2-bit packing.  isslScoreOfftargets.cpp:63-71 and :99-102 (A=0 C=1 G=2 T=3, anything else 0,
 * base j at bits 2j..2j+1); isslCreateIndex.cpp:39-47 is the same function.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t nucleotide_code(unsigned char c)
{
    switch (c) { case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 0; }
}

uint64_t oracle_sequence_to_signature(const char *seq, size_t seqLength)
{
    uint64_t signature = 0;
    for (size_t j = 0; j < seqLength; j++)
        signature |= nucleotide_code((unsigned char)seq[j]) << (j * 2);
    return signature;
}

/* isslScoreOfftargets.cpp:82-89 */
void oracle_signature_to_sequence(uint64_t signature, size_t seqLength, char *out)
{
    static const char letters[4] = { 'A', 'C', 'G', 'T' };
    for (size_t j = 0; j < seqLength; j++)
        out[j] = letters[(signature >> (j * 2)) & 0x3];
}

int oracle_method_from_string(const char *s)
{
    /* isslScoreOfftargets.cpp:121-143 */
    if (!strcmp(s, "and")) return M_AND;
    if (!strcmp(s, "or"))  return M_OR;
    if (!strcmp(s, "avg")) return M_AVG;
    if (!strcmp(s, "mit")) return M_MIT;
    if (!strcmp(s, "cfd")) return M_CFD;
    return M_UNKNOWN;
}

/* ------------------------------------------------------------------------------------------
 * .issl image view.  Layout written by isslCreateIndex.cpp:256-289 and read back by
 * isslScoreOfftargets.cpp:162-240: 6 x u64 header, scoresCount x (u64 mask, f64 score),
 * offtargetsCount x u64 signature, sliceCount * 2^sliceWidth x u64 list length, then the
 * concatenated lists of u64 (occurrences << 32 | signatureId).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t offtargetsCount, seqLength, seqCount, sliceWidth, sliceCount, scoresCount;
    uint64_t sliceLimit;
    const uint64_t *scorePairs;   /* interleaved mask, score-bits */
    const uint64_t *offtargets;
    const uint64_t *sizes;
    const uint64_t *entries;
    uint64_t entriesAvailable;
} issl_view;

static int issl_view_init(issl_view *v, const uint8_t *img, size_t len)
{
    if (len < 48) return 1;                       /* :164-167 header invalid */
    const uint64_t *h = (const uint64_t *)img;
    v->offtargetsCount = h[0]; v->seqLength = h[1]; v->seqCount = h[2];
    v->sliceWidth = h[3]; v->sliceCount = h[4]; v->scoresCount = h[5];
    if (v->sliceWidth >= 32) return 1;
    v->sliceLimit = 1ull << v->sliceWidth;        /* :179 */
    uint64_t off = 6;
    uint64_t words = len / 8;
    if (off + 2 * v->scoresCount > words) return 1;
    v->scorePairs = h + off; off += 2 * v->scoresCount;
    if (v->offtargetsCount == 0 || off + v->offtargetsCount > words) return 2;   /* :201-204 */
    v->offtargets = h + off; off += v->offtargetsCount;
    if (off + v->sliceCount * v->sliceLimit > words) return 3;                    /* :223-226 */
    v->sizes = h + off; off += v->sliceCount * v->sliceLimit;
    v->entries = h + off;
    v->entriesAvailable = words - off;
    if (v->entriesAvailable == 0) return 4;                                       /* :237-240 */
    return 0;
}

int oracle_issl_header(const uint8_t *img, size_t len, uint64_t out[6])
{
    issl_view v;
    int rc = issl_view_init(&v, img, len);
    if (len >= 48) memcpy(out, img, 48);
    return rc;
}

/* sorted (mask, score) table standing in for phmap::flat_hash_map<uint64_t,double>
 * (isslScoreOfftargets.cpp:188-197): insert() keeps the FIRST value of a duplicated key,
 * operator[] on a missing key yields 0.0 (:394). */
typedef struct { uint64_t mask; double score; uint64_t order; } mit_entry;

static int mit_cmp(const void *a, const void *b)
{
    const mit_entry *x = a, *y = b;
    if (x->mask != y->mask) return x->mask < y->mask ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order);
}

static size_t mit_table_build(const issl_view *v, mit_entry **out)
{
    size_t n = v->scoresCount, m = 0;
    mit_entry *t = malloc((n ? n : 1) * sizeof *t);
    for (size_t i = 0; i < n; i++) {
        t[i].mask = v->scorePairs[2 * i];
        memcpy(&t[i].score, &v->scorePairs[2 * i + 1], 8);
        t[i].order = i;
    }
    qsort(t, n, sizeof *t, mit_cmp);
    for (size_t i = 0; i < n; i++)
        if (m == 0 || t[m - 1].mask != t[i].mask) t[m++] = t[i];
    *out = t;
    return m;
}

static inline double mit_lookup(const mit_entry *t, size_t n, uint64_t mask)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (t[mid].mask < mask) lo = mid + 1; else hi = mid;
    }
    return (lo < n && t[lo].mask == mask) ? t[lo].score : 0.0;
}

/* ------------------------------------------------------------------------------------------
 * The scoring loop.  isslScoreOfftargets.cpp:308-511, statement for statement.
 *
 * hits (optional): every (guide index, signatureId, dist, occurrences) tuple that reaches the
 * scoring block at :390, in the order each guide meets them.  hit_count receives the TOTAL
 * number (may exceed hit_cap; only the first hit_cap tuples of the global, per-guide-ordered
 * sequence are stored when threads == 1; with threads > 1 storage order is unspecified).
 * candidates_out (optional, per guide): list entries visited by the loop at :344, the unit of
 * work of SURVEY.md section 8(d).
 * ------------------------------------------------------------------------------------------ */
int oracle_score_issl(const uint8_t *img, size_t len,
                      const uint64_t *guides, size_t nGuides,
                      int maxDist, double threshold, int method, int threads,
                      double *mitOut, double *cfdOut,
                      uint64_t *hitGuide, uint32_t *hitId, int32_t *hitDist, uint32_t *hitOcc,
                      size_t hitCap, size_t *hitCount,
                      uint64_t *candidatesOut)
{
    issl_view v;
    int rc = issl_view_init(&v, img, len);
    if (rc) return rc;

    const int calcMit = (method == M_AND || method == M_OR || method == M_AVG || method == M_MIT);
    const int calcCfd = (method == M_AND || method == M_OR || method == M_AVG || method == M_CFD);

    mit_entry *mit = NULL;
    const size_t mitN = mit_table_build(&v, &mit);

    /* :214 */
    const uint64_t numOfftargetToggles = (v.offtargetsCount / 64) + 1;

    /* :261-270  list start offsets by prefix-walking the size table */
    const uint64_t nLists = v.sliceCount * v.sliceLimit;
    uint64_t *listStart = malloc((nLists + 1) * sizeof *listStart);
    listStart[0] = 0;
    for (uint64_t i = 0; i < nLists; i++) listStart[i + 1] = listStart[i] + v.sizes[i];

    size_t hitTotal = 0;
    int bad = 0;

#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif

#pragma omp parallel
    {
        /* :311-313 per-thread toggle words, addressed from the tail */
        uint64_t *offtargetToggles = calloc(numOfftargetToggles, sizeof(uint64_t));
        uint64_t *offtargetTogglesTail = offtargetToggles + numOfftargetToggles - 1;

#pragma omp for schedule(static)
        for (size_t searchIdx = 0; searchIdx < nGuides; searchIdx++) {
            const uint64_t searchSignature = guides[searchIdx];
            double totScoreMit = 0.0, totScoreCfd = 0.0;
            const double maximum_sum = (10000.0 - threshold * 100) / threshold;   /* :326 */
            int checkNextSlice = 1;
            uint64_t visited = 0;

            for (uint64_t i = 0; i < v.sliceCount && checkNextSlice; i++) {        /* :330, :501 */
                uint64_t sliceMask = v.sliceLimit - 1;
                const int sliceShift = (int)(v.sliceWidth * i);
                sliceMask = sliceShift < 64 ? sliceMask << sliceShift : 0;
                const uint64_t searchSlice = sliceShift < 64 ? (searchSignature & sliceMask) >> sliceShift : 0;
                const uint64_t idx = i * v.sliceLimit + searchSlice;
                const uint64_t signaturesInSlice = v.sizes[idx];
                const uint64_t base = listStart[idx];

                for (uint64_t j = 0; j < signaturesInSlice; j++) {                 /* :344 */
                    if (base + j >= v.entriesAvailable) { bad = 1; break; }
                    visited++;
                    const uint64_t e = v.entries[base + j];
                    const uint64_t signatureId = e & 0xFFFFFFFFull;                /* :347 */
                    const uint32_t occurrences = (uint32_t)(e >> 32);              /* :348 */
                    if (signatureId >= v.offtargetsCount) { bad = 1; continue; }

                    /* :376-380 */
                    const uint64_t xoredSignatures = searchSignature ^ v.offtargets[signatureId];
                    const uint64_t evenBits = xoredSignatures & 0xAAAAAAAAAAAAAAAAull;
                    const uint64_t oddBits = xoredSignatures & 0x5555555555555555ull;
                    const uint64_t mismatches = (evenBits >> 1) | oddBits;
                    const int dist = __builtin_popcountll(mismatches);

                    if (dist >= 0 && dist <= maxDist) {                            /* :382 */
                        uint64_t *ptrOfftargetFlag = offtargetTogglesTail - (signatureId / 64);
                        const uint64_t seen = (*ptrOfftargetFlag >> (signatureId % 64)) & 1ULL;
                        if (!seen) {                                               /* :390 */
                            if (hitCount) {
                                size_t slot;
#pragma omp atomic capture
                                slot = hitTotal++;
                                if (slot < hitCap && hitGuide) {
                                    hitGuide[slot] = searchIdx; hitId[slot] = (uint32_t)signatureId;
                                    hitDist[slot] = dist; hitOcc[slot] = occurrences;
                                }
                            }
                            if (calcMit && dist > 0)                               /* :392-396 */
                                totScoreMit += mit_lookup(mit, mitN, mismatches) * (double)occurrences;

                            if (calcCfd) {                                         /* :399-461 */
                                double cfdScore = 0;
                                if (dist == 0) {
                                    cfdScore = 1;
                                } else if (dist > 0 && dist <= maxDist) {
                                    cfdScore = ORACLE_CFD_PAM[0xA];                /* :411, 0b1010 = GG */
                                    for (size_t pos = 0; pos < 20; pos++) {        /* :413 */
                                        const uint64_t g = (searchSignature >> (pos * 2)) & 3;
                                        const uint64_t o = (v.offtargets[signatureId] >> (pos * 2)) & 3;
                                        const size_t mask = (pos << 4) | (g << 2) | (o ^ 3);   /* :453 */
                                        if (g != o) cfdScore *= ORACLE_CFD_POS[mask];           /* :455-457 */
                                    }
                                }
                                totScoreCfd += cfdScore * (double)occurrences;     /* :460 */
                            }

                            *ptrOfftargetFlag |= (1ULL << (signatureId % 64));     /* :463 */

                            /* :466-496 */
                            int stop = 0;
                            if (method == M_AND) stop = (totScoreMit > maximum_sum && totScoreCfd > maximum_sum);
                            if (method == M_OR)  stop = (totScoreMit > maximum_sum || totScoreCfd > maximum_sum);
                            if (method == M_AVG) stop = (((totScoreMit + totScoreCfd) / 2.0) > maximum_sum);
                            if (method == M_MIT) stop = (totScoreMit > maximum_sum);
                            if (method == M_CFD) stop = (totScoreCfd > maximum_sum);
                            if (stop) { checkNextSlice = 0; break; }
                        }
                    }
                }
            }

            if (mitOut) mitOut[searchIdx] = 10000.0 / (100.0 + totScoreMit);       /* :505 */
            if (cfdOut) cfdOut[searchIdx] = 10000.0 / (100.0 + totScoreCfd);       /* :506 */
            if (candidatesOut) candidatesOut[searchIdx] = visited;
            memset(offtargetToggles, 0, sizeof(uint64_t) * numOfftargetToggles);   /* :508 */
        }
        free(offtargetToggles);
    }

    if (hitCount) *hitCount = hitTotal;
    free(listStart);
    free(mit);
    return bad ? 9 : 0;
}

/* ------------------------------------------------------------------------------------------
 * main() on memory buffers: guide file bytes in, stdout bytes out.
 * isslScoreOfftargets.cpp:275-305 (guide file checks + packing) and :514-527 (printing).
 * Returns 0, or the reference's exit status 1 with *outLen = 0.  outCap must be at least
 * nGuides * (seqLength + 2 * 330) bytes; *outLen receives the bytes written.
 * ------------------------------------------------------------------------------------------ */
int oracle_cli(const uint8_t *img, size_t len,
               const char *guideFile, size_t guideFileSize,
               int maxDist, double threshold, const char *methodStr, int threads,
               char *out, size_t outCap, size_t *outLen)
{
    issl_view v;
    *outLen = 0;
    if (issl_view_init(&v, img, len)) return 1;
    const size_t seqLineLength = v.seqLength + 1;
    if (guideFileSize % seqLineLength != 0) return 1;     /* :277-282 */
    if (guideFileSize == 0) return 1;                     /* :290-293: fread(buf, 0, 1, fp) < 1 */
    const size_t queryCount = guideFileSize / seqLineLength;
    const int method = oracle_method_from_string(methodStr);
    const int calcMit = (method == M_AND || method == M_OR || method == M_AVG || method == M_MIT);
    const int calcCfd = (method == M_AND || method == M_OR || method == M_AVG || method == M_CFD);

    uint64_t *sig = malloc(queryCount * sizeof *sig);
    double *mit = malloc(queryCount * sizeof *mit), *cfd = malloc(queryCount * sizeof *cfd);
    for (size_t i = 0; i < queryCount; i++)
        sig[i] = oracle_sequence_to_signature(guideFile + i * seqLineLength, v.seqLength);

    int rc = oracle_score_issl(img, len, sig, queryCount, maxDist, threshold, method, threads,
                               mit, cfd, NULL, NULL, NULL, NULL, 0, NULL, NULL);
    size_t w = 0;
    if (rc == 0) {
        char seq[72];
        for (size_t i = 0; i < queryCount; i++) {
            if (w + v.seqLength + 700 > outCap) { rc = 8; break; }
            oracle_signature_to_sequence(sig[i], v.seqLength, seq);
            seq[v.seqLength] = 0;
            w += (size_t)sprintf(out + w, "%s\t", seq);                            /* :516 */
            if (calcMit) w += (size_t)sprintf(out + w, "%f\t", mit[i]); else w += (size_t)sprintf(out + w, "-1\t");
            if (calcCfd) w += (size_t)sprintf(out + w, "%f\n", cfd[i]); else w += (size_t)sprintf(out + w, "-1\n");
        }
    }
    *outLen = rc ? 0 : w;
    free(sig); free(mit); free(cfd);
    return rc ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * Index builder.  isslCreateIndex.cpp:132-296.
 * ------------------------------------------------------------------------------------------ */

/* isslCreateIndex.cpp:59-91: all masks with `mismatches` set positions out of `seqLength`,
 * a set position p contributing bit 2p; same recursion, hence the same emission order. */
typedef struct { uint64_t *v; size_t n, cap; } u64vec;
static void u64vec_push(u64vec *a, uint64_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = realloc(a->v, a->cap * sizeof(uint64_t)); }
    a->v[a->n++] = x;
}

static void compute_masks_two_bit(int seqLength, int mismatches, uint64_t prefix, u64vec *out)
{
    if (mismatches < seqLength) {
        if (mismatches > 0) {
            compute_masks_two_bit(seqLength - 1, mismatches - 1, prefix + (1ULL << ((seqLength - 1) * 2)), out);
            compute_masks_two_bit(seqLength - 1, mismatches, prefix, out);
        } else {
            u64vec_push(out, prefix);
        }
    } else {
        uint64_t t = 0;
        for (int i = 0; i < seqLength; i++) t |= (1ULL << (i * 2));
        u64vec_push(out, prefix + t);
    }
}

/* isslCreateIndex.cpp:93-118 */
static double single_score(const int *mismatch_array, int length)
{
    int i;
    double T1 = 1.0, T2, T3, d = 0.0, score;
    static const double M[] = { 0.0, 0.0, 0.014, 0.0, 0.0, 0.395, 0.317, 0.0, 0.389, 0.079,
                                0.445, 0.508, 0.613, 0.851, 0.732, 0.828, 0.615, 0.804, 0.685, 0.583 };
    for (i = 0; i < length; ++i) T1 = T1 * (1.0 - M[mismatch_array[i]]);
    if (length == 1) d = 19.0;
    else {
        for (i = 0; i < length - 1; ++i) d += mismatch_array[i + 1] - mismatch_array[i];
        d = d / (length - 1);
    }
    T2 = 1.0 / ((19.0 - d) / 19.0 * 4.0 + 1);
    T3 = 1.0 / (length * length);
    score = T1 * T2 * T3 * 100;
    return score;
}

/* isslCreateIndex.cpp:120-130 (loop bound is the index's seqLength, a global there) */
double oracle_sscore(uint64_t xoredSignatures, size_t seqLength)
{
    int mismatch_array[32], m = 0;
    for (size_t j = 0; j < seqLength && j < 32; j++)
        if ((xoredSignatures >> (j * 2)) & 0x3) mismatch_array[m++] = (int)j;
    if (m == 0) return 0.0;
    return single_score(mismatch_array, m);
}

static int pair_cmp(const void *a, const void *b)
{
    const uint64_t *x = a, *y = b;     /* {mask, scorebits, order} */
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[2] < y[2] ? -1 : (x[2] > y[2]);
}

/* Builds the .issl image for a sorted, LF-terminated, fixed-width off-target text file.
 * Returns a malloc'ed image in *imgOut (caller frees with oracle_free) or non-zero on the
 * reference's error exits (:142-145, :147-153). */
int oracle_create_index(const char *text, size_t fileSize, size_t seqLength, size_t sliceWidth,
                        uint8_t **imgOut, size_t *lenOut)
{
    *imgOut = NULL; *lenOut = 0;
    if (seqLength > 32 || seqLength == 0 || sliceWidth == 0 || sliceWidth > 24) return 1;
    const size_t seqLineLength = seqLength + 1;
    if (fileSize % seqLineLength != 0) return 1;
    const size_t seqCount = fileSize / seqLineLength;

    /* :184-207 run-length collapse of identical adjacent lines */
    uint64_t *sigs = malloc((seqCount ? seqCount : 1) * sizeof *sigs);
    uint32_t *occs = malloc((seqCount ? seqCount : 1) * sizeof *occs);
    size_t distinct = 0, progress = 0;
    while (progress < seqCount) {
        const char *ptr = text + progress * seqLineLength;
        uint32_t occurrences = 1;
        /* the reference's memcmp at :192 runs one record past the buffer on the last run;
         * here the comparison simply stops at the end of the file. */
        while (progress + occurrences < seqCount &&
               memcmp(ptr, ptr + seqLineLength * occurrences, seqLength) == 0)
            occurrences++;
        sigs[distinct] = oracle_sequence_to_signature(ptr, seqLength);
        occs[distinct] = occurrences;
        distinct++;
        progress += occurrences;
    }

    const size_t sliceLimit = (size_t)1 << sliceWidth;            /* :212 */
    const size_t sliceCount = (seqLength * 2) / sliceWidth;       /* :213 */
    const size_t offtargetsCount = distinct;

    /* :216-234 -- note `uint8_t sliceVal` at :228: the slice value is truncated to 8 bits */
    uint64_t *sizes = calloc(sliceCount * sliceLimit + 1, sizeof *sizes);
    for (size_t i = 0; i < sliceCount; i++) {
        const int sliceShift = (int)(sliceWidth * i);
        const uint64_t sliceMask = (uint64_t)(sliceLimit - 1) << sliceShift;
        for (size_t id = 0; id < offtargetsCount; id++) {
            const uint8_t sliceVal = (uint8_t)((sigs[id] & sliceMask) >> sliceShift);
            sizes[i * sliceLimit + sliceVal]++;
        }
    }

    /* :239-252 score table; std::map => ascending mask order, first insertion wins;
     * scoresCount counts every generated mask. */
    const int maxDist = (int)(seqLength * 2 / sliceWidth) - 1;
    u64vec masks = { 0, 0, 0 };
    for (int i = 1; i <= maxDist; i++) compute_masks_two_bit(20, i, 0, &masks);
    const size_t scoresCount = masks.n;
    uint64_t *pairs = malloc((scoresCount ? scoresCount : 1) * 3 * sizeof *pairs);
    for (size_t k = 0; k < scoresCount; k++) {
        const double s = oracle_sscore(masks.v[k], seqLength);
        pairs[3 * k] = masks.v[k]; memcpy(&pairs[3 * k + 1], &s, 8); pairs[3 * k + 2] = k;
    }
    qsort(pairs, scoresCount, 3 * sizeof *pairs, pair_cmp);
    size_t uniquePairs = 0;
    for (size_t k = 0; k < scoresCount; k++)
        if (uniquePairs == 0 || pairs[3 * (uniquePairs - 1)] != pairs[3 * k]) {
            memmove(&pairs[3 * uniquePairs], &pairs[3 * k], 3 * sizeof *pairs);
            uniquePairs++;
        }

    /* :256-289 */
    const size_t words = 6 + 2 * uniquePairs + offtargetsCount + sliceCount * sliceLimit + offtargetsCount * sliceCount;
    uint64_t *img = malloc(words * 8);
    size_t w = 0;
    img[w++] = offtargetsCount; img[w++] = seqLength; img[w++] = seqCount;
    img[w++] = sliceWidth; img[w++] = sliceCount; img[w++] = scoresCount;
    for (size_t k = 0; k < uniquePairs; k++) { img[w++] = pairs[3 * k]; img[w++] = pairs[3 * k + 1]; }
    memcpy(img + w, sigs, offtargetsCount * 8); w += offtargetsCount;
    memcpy(img + w, sizes, sliceCount * sliceLimit * 8); w += sliceCount * sliceLimit;
    /* lists: for each slice, for each value, the ids in ascending order */
    uint64_t *cursor = malloc((sliceLimit + 1) * sizeof *cursor);
    for (size_t i = 0; i < sliceCount; i++) {
        const int sliceShift = (int)(sliceWidth * i);
        const uint64_t sliceMask = (uint64_t)(sliceLimit - 1) << sliceShift;
        cursor[0] = w;
        for (size_t j = 0; j < sliceLimit; j++) cursor[j + 1] = cursor[j] + sizes[i * sliceLimit + j];
        for (size_t id = 0; id < offtargetsCount; id++) {
            const uint8_t sliceVal = (uint8_t)((sigs[id] & sliceMask) >> sliceShift);
            img[cursor[sliceVal]++] = (((uint64_t)occs[id]) << 32) | (uint64_t)(uint32_t)id;   /* :230 */
        }
        w += offtargetsCount;
    }
    free(cursor); free(pairs); free(masks.v); free(sizes); free(sigs); free(occs);
    *imgOut = (uint8_t *)img; *lenOut = words * 8;
    return 0;
}

void oracle_free(void *p) { free(p); }

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
