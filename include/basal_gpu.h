/*
 * basal_gpu.h — C-ABI of the B200-native BASAL read-mapping hot path.
 *
 * The reference (JiejunShi/BASAL v1.8.1) has no plugin/FFI API: the seams are the
 * C++ calls `main.cpp` makes on `RefSeq` (index) and `SingleAlign`/`PairAlign`
 * (batches).  Each entry point below names the reference interface it replaces
 * (file:line under /root/reference).  Plain pointers and sizes only; no C++ or
 * torch types; no allocation crosses the boundary; functions never throw and
 * never exit — they return 0 or a negative BSL_E* code and leave a message in
 * bsl_last_error().
 *
 * There is NO CPU fallback behind this interface: every call fails with
 * BSL_ENODEV when no sm_100 class device can be opened.
 */
#ifndef BASAL_GPU_H_
#define BASAL_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSL_ABI_VERSION 1

/* error codes */
#define BSL_OK        0
#define BSL_EINVAL   -1   /* bad argument / bad -M rule / parameter out of range */
#define BSL_ENODEV   -2   /* no usable CUDA device                                */
#define BSL_ENOMEM   -3   /* device or pinned allocation failed                   */
#define BSL_ECUDA    -4   /* CUDA runtime error (message in bsl_last_error)       */
#define BSL_ESTATE   -5   /* call order violated (e.g. align before index build)  */
#define BSL_ELIMIT   -6   /* input exceeds a reference limit (read > 480 bp, ...) */

/* read status in bsl_hit.status */
#define BSL_ST_UNMAPPED 0  /* passed the filter, no hit within the budget (flag 0x4)   */
#define BSL_ST_UNIQUE   1  /* exactly one hit at the best level                        */
#define BSL_ST_MULTI    2  /* >1 equal-best hits; record holds the -S/-r 1 random pick */
#define BSL_ST_FILTERED 3  /* rejected by FilterReads (flag 0x204)                     */
#define BSL_ST_PAIRED   4  /* PE only: record is one end of the reported pair          */

typedef struct bsl_ctx bsl_ctx;

/* Mirrors the fields of the reference's global `Param` that reach the hot path
 * (param.h:58-90, defaults param.cpp:7-68, CLI parsing main.cpp:272-364).      */
typedef struct bsl_params {
    char     from_base;          /* -M: convert-from base, one of ACGT (param.cpp:163-171)            */
    char     to_bases[7];        /* -M: convert-to bases, NUL terminated subset of "ACGT-"            */
    uint32_t seed_size;          /* -s  10..16                  (param.cpp:108-115)                    */
    uint32_t index_interval;     /* -I  1..16                   (main.cpp:319-322)                     */
    uint32_t max_snp_num;        /* -v  as stored by the reference: <100 absolute, 100+pct otherwise   */
    uint32_t gap;                /* -g  0..3                    (main.cpp:300-305)                     */
    uint32_t max_num_hits;       /* -w  1..1000                 (main.cpp:339-341)                     */
    uint32_t min_insert;         /* -m                          (param.cpp:22)                         */
    uint32_t max_insert;         /* -x                          (param.cpp:23)                         */
    uint32_t chains;             /* -n  0,1,2                   (align.cpp:83-84)                      */
    uint32_t report_repeat_hits; /* -r  0,1,2                   (align.cpp:599-610)                    */
    uint32_t randseed;           /* -S  must be non-zero: -S 0 is irreproducible in the reference      */
    uint32_t max_ns;             /* -f                          (align.cpp:560)                        */
    uint32_t min_read_size;      /* Param::min_read_size        (param.cpp:34,112)                     */
    float    max_kmer_ratio;     /* -k                          (refbase.cpp:363)                      */
    uint32_t reserved[4];
} bsl_params;

/* One batch of reads as the loader hands them to SingleAlign::ImportBatchReads
 * (align.cpp:35, reads.h:16-23): post-trim bases, running read number, mate set. */
typedef struct bsl_batch {
    uint32_t        n;           /* number of reads                                                    */
    uint32_t        readset;     /* ReadInf::readset: 0 single-end, 1 mate #1, 2 mate #2               */
    const uint8_t  *bases;       /* concatenated ASCII bases, offsets[n] bytes                         */
    const uint64_t *offsets;     /* n+1 byte offsets into bases                                        */
    const uint32_t *index;       /* ReadInf::index per read (0-based input order) or NULL              */
    uint32_t        first_index; /* used when index==NULL: read i has index first_index+i              */
    uint32_t        n_context;   /* the first n_context reads are CONTEXT: reads that came earlier in the input
                                    and only re-establish the state the reference's aligner object carries from
                                    read to read (xseed_start_offset / xseed_array, align.cpp:476-480, 79-150);
                                    they are packed and scheduled, never mapped; their records are undefined.
                                    Every call starts with fresh aligner objects (zeroed state), reads are taken
                                    in batch order like one SingleAlign / PairAlign object would take them.   */
    const uint16_t *raw_len;     /* length before adapter trimming (align.cpp:420) or NULL = length    */
} bsl_batch;

/* Fixed-size result record (one per read): what StringAlign/s_OutHit (align.cpp:583-669)
 * and StringAlignPair/StringAlignUnpair (pairs.cpp:204-305) need to print the SAM line. */
typedef struct bsl_hit {
    uint32_t loc;        /* gHit::loc, 0-based leftmost forward-strand coordinate on the sequence      */
    uint32_t chr;        /* gHit::chr = 2*sequence_index + reference_strand                            */
    uint32_t n_hits;     /* equal-best hits counted (sum of both read chains, capped by -w semantics)  */
    uint32_t n_chain0;   /* how many of them are on read chain 0 (the rest on chain 1)                 */
    int32_t  gap_size;   /* gHit::gap_size: >0 deletion from read (D), <0 insertion (I), 0 none        */
    uint16_t gap_pos;    /* gHit::gap_pos                                                              */
    uint8_t  nm;         /* mismatch level of the reported hit (NM:i)                                  */
    uint8_t  status;     /* BSL_ST_*                                                                   */
    uint8_t  read_chain; /* 0 read as is, 1 reverse complement                                         */
    uint8_t  max_snp;    /* per-read mismatch budget after FilterReads (align.cpp:548-563)             */
    uint16_t read_len;   /* mapped read length                                                         */
    uint32_t all_first;  /* -r 2: first record of this read in the all-hits array                      */
} bsl_hit;

/* Per-pair result of PairAlign::RunAlign (pairs.cpp:132-177). */
typedef struct bsl_pair {
    uint32_t n_pairs;    /* pairs at the best level; 0 = no pair (out_a/out_b then hold unpaired picks) */
    uint32_t insert;     /* PairHit::insert                                                             */
    uint8_t  chain;      /* PairHit::chain                                                              */
    uint8_t  na, nb;     /* PairHit::na / nb                                                            */
    uint8_t  reserved;
    uint32_t all_first;  /* -r 2: first pair of this read pair in the all-pairs arrays                  */
} bsl_pair;

/* Index statistics (RefSeq after FinishIndex, refbase.h:96-120). */
typedef struct bsl_index_info {
    uint32_t n_seq;          /* RefSeq::total_num                                                  */
    uint32_t n_kmers;        /* RefSeq::total_kmers = 3^seed_size                                  */
    uint64_t sum_length;     /* RefSeq::sum_length                                                 */
    uint64_t n_words;        /* words per strand plane including both 400-word margins             */
    uint64_t n_entries;      /* total seed-table entries (both strands)                            */
    uint32_t max_kmer_num;   /* Param::max_kmer_num (refbase.cpp:363)                              */
    uint32_t reserved;
} bsl_index_info;

/* Work counters of the last align call (the reference's total_seeds / total_candidates,
 * align.cpp:283-285) plus device timings used by bench.py for the roofline.             */
typedef struct bsl_stats {
    uint64_t reads;
    uint64_t seed_lookups;   /* executed (read, chain, phase, mode) bucket look-ups                 */
    uint64_t candidates;     /* bucket entries visited = verification units                         */
    uint64_t hits_added;
    uint64_t heavy_reads;    /* reads re-run on the large-capacity path                             */
    double   ms_pack;        /* CUDA-event ms of pack+seed-select kernels                           */
    double   ms_search;      /* ms_lookup + ms_verify + ms_reduce                                   */
    double   ms_pair;        /* CUDA-event ms of the mate-pairing kernels                           */
    double   ms_total;       /* first H2D to last D2H                                               */
    uint64_t kernel_launches;
    uint64_t verify_bytes;   /* candidates x (4 + 8 W) algorithmic bytes               */
    double   ms_device;      /* first kernel to last kernel (inputs resident, records still on device) */
    uint64_t search_launches;/* verify_candidates launches inside ms_verify                          */
    double   ms_lookup;      /* CUDA-event ms of seed_lookup (bucket look-ups of every round)         */
    double   ms_verify;      /* CUDA-event ms of verify_candidates (the roofline kernel)              */
    double   ms_reduce;      /* CUDA-event ms of reduce_round (AddHit replay + gap search)            */
} bsl_stats;

/* -- context ---------------------------------------------------------------------------- */

/* Replaces: global `Param param` + `param.SetAlign(-M)` (main.cpp:26,629; param.cpp:163-263).
 * Validates the rule exactly like SetAlign and stores the 2-bit code tables.              */
int  bsl_ctx_create(bsl_ctx **out, int device, const bsl_params *params);
void bsl_ctx_destroy(bsl_ctx *ctx);
const char *bsl_last_error(const bsl_ctx *ctx);   /* ctx may be NULL: last creation error */
int  bsl_abi_version(void);
void bsl_params_default(bsl_params *p);           /* Param::Param defaults (param.cpp:7-68) */

/* -- index seam ---------------------------------------------------------------------------
 * Replaces: RefSeq::Run_ConvertBinseq (refbase.cpp:186-252) + Do_Formatdb =
 * InitialIndex/t_CalKmerFreq/AllocIndex/t_FillIndex/FinishIndex (main.cpp:136-151,
 * refbase.cpp:261-439).  Host passes the raw sequence bytes of every FASTA record;
 * the GPU packs both strand planes, hashes every I-th window, radix-sorts the
 * (kmer, location) pairs and derives max_kmer_num.                                        */
int  bsl_index_build(bsl_ctx *ctx, const uint8_t *seq_concat, const uint64_t *seq_off,
                     const uint32_t *seq_len, uint32_t n_seq);
int  bsl_index_info_get(const bsl_ctx *ctx, bsl_index_info *info);
/* Test hook: copy device index arrays back (any pointer may be NULL).
 * bucket_start[n_kmers+1], n_fwd[n_kmers], loc[n_entries], planes[n_words] each.          */
int  bsl_index_download(const bsl_ctx *ctx, uint32_t *bucket_start, uint32_t *n_fwd,
                        uint32_t *loc, uint64_t *fwd_plane, uint64_t *rc_plane);
/* Replaces RefSeq::ReleaseIndex (refbase.cpp:369-385). */
void bsl_index_free(bsl_ctx *ctx);

/* -- batch seam ---------------------------------------------------------------------------
 * Replaces: SingleAlign::ImportBatchReads + Do_Batch (align.cpp:35,565) up to, but not
 * including, SAM text.  out[n]; all_hits (capacity all_cap records, n_all returned) is only
 * written when report_repeat_hits==2 and may be NULL otherwise.                           */
int  bsl_align_se(bsl_ctx *ctx, const bsl_batch *reads, bsl_hit *out,
                  bsl_hit *all_hits, uint64_t all_cap, uint64_t *n_all);
/* Replaces: PairAlign::ImportBatchReads + Do_Batch (pairs.cpp:22,179).  a->n must equal
 * b->n.  all_a/all_b (capacity all_cap each) receive every pair of the best level when
 * report_repeat_hits==2.                                                                  */
int  bsl_align_pe(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b,
                  bsl_hit *out_a, bsl_hit *out_b, bsl_pair *out_pair,
                  bsl_hit *all_a, bsl_hit *all_b, uint64_t all_cap, uint64_t *n_all);
int  bsl_stats_get(const bsl_ctx *ctx, bsl_stats *st);
/* Measurement hook: run the kernels again on the batch the previous bsl_align_se/pe call of this
 * thread left resident in device memory (same descriptors; no H2D, no D2H of records).  Used by
 * bench.py for the inputs-resident throughput; results stay on the device.                        */
int  bsl_align_rerun(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b);

/* Pinned host memory for callers that want zero-copy staging (optional). */
void *bsl_host_alloc(size_t bytes);
void  bsl_host_free(void *p);

/* Host helpers shared with the CLI (pure functions, no device):
 * FilterReads budget (align.cpp:550-561) and myrand (utilities.cpp:38-48).                */
uint32_t bsl_read_budget(const bsl_params *p, uint32_t raw_len, uint32_t len);
uint32_t bsl_myrand(uint32_t read_index, uint32_t randseed);

#ifdef __cplusplus
}
#endif
#endif /* BASAL_GPU_H_ */
