/*
 * issl_cuda.h -- C ABI of libissl_cuda, the B200 (sm_100a) implementation of Crackling's
 * ISSL off-target scorer.
 *
 * The reference has no plugin/operator API for this path: the seam is the process boundary of
 * the `isslScoreOfftargets` executable (argv in, stdout out), whose main() is
 * /root/reference/src/ISSL/isslScoreOfftargets.cpp:91-530.  This header splits that main()
 * into the phases a host program (ours: crackling_b200/csrc/isslScoreOfftargets.cpp; a
 * maintainer's: see INTEGRATION.md) calls in order.  Each entry point cites the reference
 * region it replaces; "ref:" paths are relative to /root/reference/src/ISSL/.
 *
 * Conventions: plain C types only; every function returns ISSL_OK (0) or an issl_status
 * error code and records a message retrievable with issl_last_error() (thread-local);
 * no exceptions cross the boundary; there is NO CPU fallback -- without an sm_100 device
 * issl_device_create* fails with ISSL_ERR_NO_DEVICE.
 *
 * Threading: an issl_index is immutable after open and may be shared.  An issl_device may be
 * used by one host thread at a time; different issl_device handles (one per GPU) may be driven
 * concurrently from different host threads (guides partitioned, index replicated, no collective:
 * ref guides are independent, isslScoreOfftargets.cpp:316-317).
 */
#ifndef ISSL_CUDA_H
#define ISSL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISSL_CUDA_ABI_VERSION 4

typedef enum issl_status {
    ISSL_OK = 0,
    ISSL_ERR_IO = 1,          /* file cannot be opened / read                              */
    ISSL_ERR_FORMAT = 2,      /* not a valid .issl image (ref error exits :164-167, :201-204, :223-226, :237-240) */
    ISSL_ERR_CUDA = 3,        /* a CUDA call failed                                        */
    ISSL_ERR_NO_DEVICE = 4,   /* no usable sm_100 device -- the product never falls back to the CPU */
    ISSL_ERR_ARG = 5,         /* bad argument                                              */
    ISSL_ERR_UNSUPPORTED = 6, /* a valid file this implementation refuses (e.g. lists that violate
                                 the invariants of isslCreateIndex.cpp:216-234)            */
    ISSL_ERR_NOMEM = 7
} issl_status;

/* ref enum ScoreMethod, isslScoreOfftargets.cpp:44 and :121-143 */
typedef enum issl_method {
    ISSL_METHOD_UNKNOWN = 0,  /* any other string: both columns print as -1, nothing is scored */
    ISSL_METHOD_MIT = 1,
    ISSL_METHOD_CFD = 2,
    ISSL_METHOD_AND = 3,
    ISSL_METHOD_OR = 4,
    ISSL_METHOD_AVG = 5
} issl_method;

/* The six header words of an .issl file, ref isslScoreOfftargets.cpp:162-174
 * (written by isslCreateIndex.cpp:257-267). */
typedef struct issl_info {
    uint64_t offtargetsCount; /* distinct off-target sites                                 */
    uint64_t seqLength;       /* bases per site (20)                                       */
    uint64_t seqCount;        /* sites before duplicate collapsing                         */
    uint64_t sliceWidth;      /* bits per slice                                            */
    uint64_t sliceCount;      /* slices per site = 2*seqLength / sliceWidth                */
    uint64_t scoresCount;     /* header count of precomputed local MIT scores              */
} issl_info;

/* How the index lies in HBM (DESIGN.md "Data layout"). */
typedef enum issl_layout {
    ISSL_LAYOUT_AUTO = 0,     /* TRIPLE when it applies and fits, else RES32 when the file allows it, else SIG64 */
    ISSL_LAYOUT_RES32 = 1,    /* slice lists hold 32-bit residual signatures inline: 4 B / candidate */
    ISSL_LAYOUT_SIG64 = 2,    /* slice lists hold the 64-bit signature inline: 8 B / candidate      */
    ISSL_LAYOUT_GATHER = 3,   /* slice lists hold 32-bit ids, signatures gathered: 4 + 8 B / candidate
                                 (the layout BASELINE.json's north_star describes; kept for comparison) */
    ISSL_LAYOUT_TRIPLE = 4    /* seqLength 20, sliceWidth 8 (or 4, maxDist <= 4: ids-only lists instead of RES32).  RES32
                                 plus, for each of the 10 slice triples,
                                 the sites bucketed by their three slice values (2^24 buckets, ascending id inside)
                                 with the remaining 16 signature bits inline: every slice list sub-divided by two
                                 more slices.  A guide then reads only the sub-buckets that can hold a site within
                                 maxDist (issl_triple_visits) -- 2 B per entry actually read, ~250x fewer entries
                                 than the whole lists at maxDist 4 -- with identical results.
                                 Used for maxDist <= 6; larger distances take the RES32 path. */
} issl_layout;

typedef struct issl_device_info {
    int cuda_device;
    int layout;                    /* issl_layout actually in use                          */
    uint32_t bytes_per_candidate;  /* algorithmic bytes streamed per list entry read (TRIPLE: 2, of the sub-buckets) */
    uint64_t hbm_bytes;            /* bytes of HBM held by the index                       */
    uint64_t list_entries;         /* sliceCount * offtargetsCount                          */
    issl_info info;
    uint32_t triple_block_bytes;   /* TRIPLE: bytes of one bucket's block in the blocked copy (one read per
                                      bucket visit); 0 = no blocked copy, buckets are read through their offsets */
    uint32_t triple_hit_bytes;     /* TRIPLE: bytes gathered per hit by the fused tail (offset pair, id, record): 28, or 0
                                      when the index is in text order and a hit's own signature orders it */
} issl_device_info;

/* Counters of the last issl_score* call on a device handle. */
typedef struct issl_stats {
    uint64_t guides;
    uint64_t candidates;       /* list entries visited = the unit of work (ref loop at :344)       */
    uint64_t hits;             /* entries with dist <= maxDist that survived de-duplication        */
    uint64_t scan_launches;    /* launches of the candidate-scan kernel                            */
    uint64_t launches;         /* all kernel launches of ours                                      */
    double scan_ms;            /* device time of the scan kernel(s), CUDA events on the call's stream */
    double total_ms;           /* device time of the whole call (setup + scan + sort + score)      */
    uint64_t early_exits;      /* guides that stopped before the last slice (threshold > 0)        */
    uint64_t streamed;         /* list entries actually read from HBM (each chunk once per guide GROUP;
                                  TRIPLE: entries of the sub-buckets visited)                        */
    uint64_t bucket_visits;    /* TRIPLE: (guide, sub-bucket) visits = bucket-offset pairs read; 0 otherwise */
    uint64_t heavy_hits;       /* TRIPLE: hits of guides with more hits than a CTA's record list holds, sorted and
                                  accumulated inside the scan kernel (ABI 4)                          */
    uint64_t sorted_hits;      /* hits that went through the general pipeline's device-wide sort (ABI 4) */
    double heavy_ms;           /* device time of k_heavy_finish, the per-guide sort + accumulation of those guides (ABI 4) */
} issl_stats;

typedef struct issl_index issl_index;     /* a parsed .issl image in host memory          */
typedef struct issl_device issl_device;   /* an index resident in one GPU's HBM           */

/* ---- host side: the .issl file ------------------------------------------------------------ */

/* Maps and validates an .issl file.  Replaces the sequential fread of
 * isslScoreOfftargets.cpp:152-243 (header, score table, signatures, list sizes, lists). */
int issl_index_open(const char *path, issl_index **out);

/* Same, over an image already in memory (not copied; must outlive the handle). */
int issl_index_from_memory(const void *image, size_t bytes, issl_index **out);

int issl_index_info(const issl_index *index, issl_info *out);
void issl_index_close(issl_index *index);

/* 2-bit packs a guide file (fixed-width lines of seqLength bases + LF) into out[bytes/(seqLength+1)].
 * Replaces sequenceToSignature, isslScoreOfftargets.cpp:63-71 with the table of :99-102
 * (A=0 C=1 G=2 T=3, anything else 0), applied as in :297-305.  Fails with ISSL_ERR_ARG when
 * bytes is not a multiple of seqLength+1 (ref :277-282). */
int issl_pack_guides(const char *text, size_t bytes, size_t seqLength, uint64_t *out);

/* Inverse, ref signatureToSequence :82-89; out receives seqLength chars (no terminator). */
void issl_unpack_guide(uint64_t signature, size_t seqLength, char *out);

/* ref :121-143 */
int issl_method_from_string(const char *name);

/* The output lines of isslScoreOfftargets.cpp:514-527 for guides[0..n): the sequence re-decoded from the packed guide
 * (:82-89, :515), a tab, "%f" of the MIT score or the literal -1 when the method does not compute it (:516-520), a tab,
 * the same for CFD (:521-525), a newline.  Formatted by all host cores; the digits are those of printf("%f").
 * Returns the number of bytes the lines take; they are written to out only when that fits cap (call with
 * out = NULL to size the buffer: at most seqLength + 2 + 2 * 330 bytes per line). */
size_t issl_format_lines(const uint64_t *guides, const double *mit, const double *cfd, size_t n, size_t seqLength,
                         int method, char *out, size_t cap);

/* ---- device side -------------------------------------------------------------------------- */

/* Number of usable sm_100 devices (0 when there is none; never an error). */
int issl_device_count(void);

/* One-time re-layout of the index into the HBM of `cuda_device` (call once per GPU).
 * Replaces the in-memory structures built at isslScoreOfftargets.cpp:200-270 (offtargets[],
 * allSlicelistSizes, allSignatures, sliceLists pointer table).  Validates the builder's
 * invariants (isslCreateIndex.cpp:216-234) on the device while copying. */
int issl_device_create(const issl_index *index, int cuda_device, int layout, issl_device **out);

/* Builds a synthetic index directly in HBM (benchmarks: a human-scale .issl is ~28 GB and cannot
 * be shipped to the GPU box).  Sites are drawn from a counter-based RNG: `uniform_sites` i.i.d.
 * uniform 20-mers whose first base is A/C/G (the extractor's regex, extractOfftargets.py:23),
 * plus `families` near-repeat families of `family_size` copies with per-base substitution
 * rate <= max_sub_rate; then sorted, run-length collapsed into occurrence counts and turned into
 * slice lists exactly as isslCreateIndex.cpp:184-252 does (including its 8-bit slice truncation). */
int issl_device_create_synthetic(int cuda_device, int layout, uint64_t seed, uint64_t uniform_sites,
                                 uint32_t families, uint32_t family_size, double max_sub_rate,
                                 uint32_t seqLength, uint32_t sliceWidth, issl_device **out);

/* The same with the repeat structure BASELINE.json configs[3] / SURVEY.md 8d (C4) asks for: family sizes drawn
 * log-uniformly from [family_size_min, family_size_max] (equal: fixed), and low_complexity_fraction x (uniform + family
 * sites) further sites that overlap a poly-A / poly-T / dinucleotide tract (10-20 bases of the repeat on one side of the
 * window, 2 % of them substituted): they skew the slice-list lengths and give sites that occur thousands of times. */
typedef struct issl_synth_spec {
    uint64_t seed;
    uint64_t uniform_sites;
    uint32_t families;
    uint32_t family_size_min, family_size_max;
    double max_sub_rate;
    double low_complexity_fraction;
    uint32_t seqLength, sliceWidth;
} issl_synth_spec;
int issl_device_create_synthetic_ex(int cuda_device, int layout, const issl_synth_spec *spec, issl_device **out);

/* Lengths of the index's slice lists, slice-major (ref allSlicelistSizes, isslScoreOfftargets.cpp:204-216): returns the
 * number of lists (sliceCount << min(sliceWidth, ...)) and fills out[0..min(cap, lists)).  For list-length histograms. */
size_t issl_device_list_lengths(const issl_device *dev, uint64_t *out, size_t cap);

/* Builds the index on the device from the text file isslCreateIndex reads: fixed-width lines of seqLength
 * bases + LF, sorted, duplicates adjacent.  Replaces isslCreateIndex.cpp:138-252 (record packing :39-47,
 * run-length collapse of identical adjacent lines into occurrence counts :184-207, slice lists with the 8-bit
 * slice-value truncation :216-234, local MIT score table :239-252).  Together with issl_device_write_issl
 * (:256-289) this is a drop-in for the isslCreateIndex executable (bin/isslCreateIndex); the handle can also
 * be scored directly, skipping the .issl file. */
int issl_device_create_from_text(const char *text, size_t bytes, uint32_t seqLength, uint32_t sliceWidth,
                                 int cuda_device, int layout, issl_device **out);

/* ---- off-target site extraction (the step before the index) ------------------------------- */

/* The off-target sites of a genome, held as sort keys in one GPU's HBM.  Replaces
 * /root/reference/src/crackling/utils/extractOfftargets.py: FASTA reading (:26-62, :73-90), the two look-ahead
 * regexes (:23-24: forward [ACG][ACGT]{19}[ACGT][AG]G, reverse C[CT][ACGT][ACGT]{19}[TGC]), the slicing
 * (:97-106: match[0:20], reverse-complemented on the reverse strand) and the global sort (:112-191). */
typedef struct issl_sites issl_sites;

int issl_sites_create(int cuda_device, issl_sites **out);
void issl_sites_destroy(issl_sites *sites);

/* Extracts the sites of one FASTA / multi-FASTA / plain-sequence buffer and appends them.  single_input != 0
 * selects the reading rules of the tool's one-input path (explodeMultiFastaFile, :26-62: lines stripped on both
 * sides, every record kept); 0 those of its several-inputs path (processingNode, :73-90: lines stripped on the
 * right only, records keyed by header text so that a repeated header discards the earlier record, :83).
 * Lines end at LF, CR or CRLF (Python's universal newlines); characters are upper-cased (:59, :89). */
int issl_sites_add_fasta(issl_sites *sites, const char *text, size_t bytes, int single_input);

/* Sites found so far (before duplicate collapsing) and sequence characters scanned. */
int issl_sites_count(const issl_sites *sites, uint64_t *n_sites, uint64_t *n_characters);

/* Writes the sorted text file the tool writes (20 bases + LF per site, Python string order; ref :112-191). */
int issl_sites_write_text(issl_sites *sites, const char *path);

/* Copies sorted sites [first, first + n) to the host as sort keys: base 0 in bits 38..39, ..., base 19 in bits
 * 0..1 (A=0 C=1 G=2 T=3), so ascending keys are ascending text lines. */
int issl_sites_read_keys(issl_sites *sites, uint64_t first, uint64_t n, uint64_t *out);

/* Builds the index from the extracted sites on the same GPU, without the text file in between: what
 * extractOfftargets followed by isslCreateIndex (seqLength 20) produces. */
int issl_device_create_from_sites(issl_sites *sites, uint32_t sliceWidth, int layout, issl_device **out);

/* A second copy of a resident index on another GPU, copied device to device over NVLink (no second pass over the
 * file, no second build of the layout).  The reference shares one in-memory index between its OpenMP threads
 * (isslScoreOfftargets.cpp:308-317); across GPUs the index is replicated instead, and this is how a replica is
 * made.  The source handle is only read and may be scoring meanwhile. */
int issl_device_clone(const issl_device *src, int cuda_device, issl_device **out);

int issl_device_get_info(const issl_device *dev, issl_device_info *out);
void issl_device_destroy(issl_device *dev);

/* Serialises the device-resident index back into the reference's .issl byte format
 * (isslCreateIndex.cpp:256-289) -- used to hand a synthetic index to the reference binary. */
int issl_device_write_issl(issl_device *dev, const char *path);

/* Copies the signatures of the sites site_ids[0..n) (taken modulo offtargetsCount) to host --
 * used to draw benchmark guides from a synthetic index. */
int issl_device_read_sites(issl_device *dev, const uint64_t *site_ids, uint64_t n, uint64_t *out);

/* Scores n packed guides.  Replaces the parallel region isslScoreOfftargets.cpp:308-511:
 * mit_out[i] = 10000/(100 + sum of local MIT scores * occurrences), cfd_out[i] likewise for CFD
 * (:505-506), with the reference's order-dependent early exit (:326, :466-496) reproduced exactly.
 * Host buffers; H2D of guides and D2H of scores happen inside the call; blocking.
 * A column the method does not compute is left untouched (the reference prints -1 for it);
 * its pointer may be NULL. */
int issl_score(issl_device *dev, const uint64_t *guides, size_t n, int maxDist, double threshold,
               int method, double *mit_out, double *cfd_out);

/* Same with DEVICE pointers and a caller stream (cudaStream_t passed as void*, NULL = the
 * handle's own stream).  Work is enqueued on that stream; the call returns after the stream
 * has drained (the pipeline reads survivor counts back between phases). */
int issl_score_device(issl_device *dev, const uint64_t *d_guides, size_t n, int maxDist,
                      double threshold, int method, double *d_mit_out, double *d_cfd_out,
                      void *stream);

/* Debug / parity export: scores like issl_score and also returns every scored hit as
 * (guide index, site id, distance, occurrences), sorted by (guide, slice, list position) -- the
 * order in which the reference meets them (hits after a guide's early exit are not reported).
 * *count receives the total; at most cap tuples are stored. */
int issl_score_hits(issl_device *dev, const uint64_t *guides, size_t n, int maxDist, double threshold,
                    int method, double *mit_out, double *cfd_out,
                    uint64_t *hit_guide, uint32_t *hit_id, int32_t *hit_dist, uint32_t *hit_occ,
                    size_t cap, size_t *count);

int issl_last_stats(const issl_device *dev, issl_stats *out);

/* One call, several GPUs (one process): replaces the `#pragma omp for` over guides of isslScoreOfftargets.cpp:308-317
 * across devices.  devs[0..n_devs) hold replicas of the same index (issl_device_create / issl_device_clone), one per
 * GPU.  The guides are cut into chunks of `chunk` guides (0 = issl_multi_chunk(n, n_devs)) that the devices take from a
 * shared counter, one host thread per device -- dynamic, because the early exit makes a guide's cost uneven; every chunk
 * writes its own range of mit_out / cfd_out, so results are in input order and there is no cross-GPU reduction.
 * Host buffers as for issl_score (pinned memory from issl_host_alloc avoids the driver's staging copies).
 * stats_out (optional): counters summed over devices, times of the busiest device; guides_per_device (optional,
 * n_devs entries): how many guides each device ended up scoring. */
int issl_score_multi(issl_device *const *devs, size_t n_devs, const uint64_t *guides, size_t n, int maxDist,
                     double threshold, int method, double *mit_out, double *cfd_out, size_t chunk,
                     issl_stats *stats_out, uint64_t *guides_per_device);
size_t issl_multi_chunk(size_t n, size_t n_devs);

/* Pinned, portable host memory (cudaHostAlloc) for guide and score arrays shared by several devices. */
int issl_host_alloc(size_t bytes, void **out);
void issl_host_free(void *p);

/* ---- guide-side pre-filters of the pipeline (the step before the scorer) ------------------- */

/* Crackling's sequence-only consensus filters, /root/reference/src/crackling/Crackling.py:312-384, and the packing
 * of the 20-mer the scorer is given (:747-752 writes target23[0:20] to the guide file), in one pass on the device.
 * text: n lines of 23 characters + LF (target23 as produced at :151-165: 20-mer + PAM on the forward strand, or the
 * reverse complement of a CC... match).  Per target:
 *   flags_out (optional): ISSL_FILTER_* bits of the filters the target FAILS;
 *   at_out (optional):    AT_percentage(target23[0:20]) = 100.0 * count(A, T) / 20.0 (Helpers.py:21-27);
 *   packed_out (optional): the 20-mer 2-bit packed as issl_pack_guides does, ready for issl_score.
 * bytes must be a multiple of 24 (ISSL_ERR_ARG otherwise). */
#define ISSL_FILTER_G20 1u        /* CHOPCHOP: target23[19] != 'G'                                   (:318-326)  */
#define ISSL_FILTER_LEADING_T 2u  /* mm10db: (ends "GG" and starts 'T') or (starts "CC" and ends 'A') (:336-346) */
#define ISSL_FILTER_AT 4u         /* mm10db: AT % of the 20-mer < 20 or > 65                          (:356-368) */
#define ISSL_FILTER_TTTT 8u       /* mm10db: "TTTT" in target23                                       (:378-384) */
int issl_guide_filters(issl_device *dev, const char *text, size_t bytes, uint8_t *flags_out, double *at_out,
                       uint64_t *packed_out);

/* Duplicate candidate guides, /root/reference/src/crackling/Crackling.py:211-240 and :291-296, as one device pass over the
 * same text (n lines of 23 characters + LF, in the order the pipeline discovers them): the 23-mers are packed to 46-bit
 * keys, sorted stably with their positions, and runs of equal keys are read off.  flags_out[i] (required) gets
 *   ISSL_FILTER_DUPLICATE  when an equal target23 occurs EARLIER in the text: the pipeline keeps only the first
 *                          occurrence in candidateGuides and counts this one in numDuplicateGuides (:221-224);
 *   ISSL_FILTER_NOT_UNIQUE when the target23 occurs more than once at all (it is in duplicateGuides: isUnique = rejected,
 *                          position fields ambiguous, :291-296) -- set on the first occurrence too.
 * n_later (optional) = numDuplicateGuides, n_sequences (optional) = len(duplicateGuides) for this text. */
#define ISSL_FILTER_DUPLICATE 16u
#define ISSL_FILTER_NOT_UNIQUE 32u
int issl_guide_duplicates(issl_device *dev, const char *text, size_t bytes, uint8_t *flags_out, uint64_t *n_later,
                          uint64_t *n_sequences);

/* ---- builder-side arithmetic (used by the synthetic builder; exposed for parity tests) ---- */

/* Local MIT score of a mismatch mask (bit 2*pos set per mismatching position).
 * ref isslCreateIndex.cpp:93-130 (single_score / sscore). */
double issl_local_mit_score(uint64_t mask, size_t seqLength);

/* The score table isslCreateIndex.cpp:239-252 writes: masks ascending; returns the number of
 * entries written (<= cap) and the header's scoresCount through *scoresCount. */
size_t issl_mit_table(size_t seqLength, size_t sliceWidth, uint64_t *masks, double *scores, size_t cap,
                      uint64_t *scoresCount);

/* ISSL_LAYOUT_TRIPLE: the sub-buckets one guide has to read for a given maxDist, as XOR patterns relative to
 * the guide's own bucket.  Entry = pattern24 | triple << 24 | budget << 28: `triple` indexes the ten slice triples
 * (issl_triple_layout); pattern24 is XORed onto the guide's bucket key (the triple's three slices in key bytes
 * 0..2); an entry of that bucket can only be a hit if its 16 residual bits
 * differ from the guide's in at most `budget` bases.  Entries are ordered by the lowest slice on which their hits
 * match the guide exactly -- the slice through which the reference meets them first
 * (isslScoreOfftargets.cpp:330-390) -- and waveStart[s] .. waveStart[s+1] delimits slice s.
 * Returns the number of entries (written up to cap); out may be NULL to size the table.  0 <= maxDist <= 7. */
size_t issl_triple_visits(int maxDist, uint32_t *out, size_t cap, uint32_t waveStart[6]);

/* ... for sliceWidth 4 (ten 2-base slices): from maxDist 5 on a site can agree with the guide on a 2-base slice without
 * agreeing on any whole byte; the buckets of triple 0 whose three key bytes all differ from the guide's are added
 * (1 728 for maxDist 5, 25 056 for maxDist 6), budget = what is left for the residual, both slices of which differ too. */
size_t issl_triple_visits_w4(int maxDist, uint32_t *out, size_t cap, uint32_t waveStart[6]);

/* The fixed tables behind it: slices_out[t*5 + 0..2] = the slices in key bytes 0..2 of triple t, [t*5 + 3..4] the two
 * residual slices; resp_out[E] = the triple responsible for a site whose exactly matching slices are the set E
 * (bit s = slice s; resp_out[0] = 15).  Either pointer may be NULL. */
void issl_triple_layout(uint8_t slices_out[50], uint8_t resp_out[32]);

const char *issl_last_error(void);
int issl_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ISSL_CUDA_H */
