// common.cuh — shared types and bit primitives of the B200 BASAL hot path.
//
// Bit conventions follow the reference so that global coordinates and packed words
// are interchangeable with its `refcat/crefcat` and `xseq` arrays:
//   * 2 bits per base, base k of a 64-bit word at bits 63-2k..62-2k (refbase.cpp:76-79)
//   * convert-from base = code 1; a single non-'-' convert-to base = code 3 (param.cpp:216-233)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/basal_gpu.h"

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;

#define BSL_MAXSNPS 15u        // param.h:18
#define BSL_REF_MARGIN 400u    // refbase.h:16 (words)
#define BSL_MAX_READLEN 480u   // (FIXELEMENT-1)*SEGLEN, param.cpp:58
#define BSL_MAX_WORDS 15u      // words per read plane
#define BSL_SM_COUNT 148       // B200

#define HD __host__ __device__ __forceinline__

// ---- bit primitives (param.h:104-142) -------------------------------------------------
// XT: fold code 3 onto 1 in every 2-bit digit, then read the 16 digits as a base-3 number.
HD u32 bsl_xt(u32 t) {
    t -= (t << 1) & t & 0xAAAAAAAAu;                 // 11 -> 01
    t -= (t >> 2) & 0x33333333u;                     // pairs of digits: 4a+b -> 3a+b
    u32 s = (t & 0xF0F0F0F0u) >> 1; t -= s - (s >> 3);            // nibbles: 16a+b -> 9a+b
    s = (t & 0xFF00FF00u) >> 2; t = (t & 0x00FF00FFu) + s + (s >> 2) + (s >> 6);   // bytes: 256a+b -> 81a+b
    return (t & 0xFFFFu) + (t >> 16) * 6561u;
}
// XC64: per digit 01 -> 01, everything else -> 11 (mask that lets read 01/11 match reference 01)
HD u64 bsl_xc(u64 t) { return ((~t) << 1) | t | 0x5555555555555555ULL; }
// M2_judge: per digit 11 -> 11, everything else -> 00
HD u64 bsl_m2(u64 t) { return t & (((t & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((t & 0x5555555555555555ULL) << 1)); }
// collapse every non-zero digit to 01 (so popcount == XM64 and clz/ctz give positions)
HD u64 bsl_pairs(u64 d) { return (d | (d >> 1)) & 0x5555555555555555ULL; }

// mismatch digits of one read word against the aligned reference word, WITHOUT the N mask.
// single conversion: align.h:126-128 ; multi-way / '-' : align.h:210-236
template <bool SINGLE>
HD u64 bsl_diff(u64 q, u64 cm, u64 r) {
    if (SINGLE) return (q & bsl_xc(r)) ^ r;
    u64 m2 = bsl_xc(r) | cm, m3 = bsl_m2(m2);
    return ((~m3 & m2) | (m3 & q)) ^ r;
}

// myrand for -S != 0 (utilities.cpp:38-48); note the 32-bit wrap of randseed*1000000
HD u32 bsl_rand(u32 index, u32 seed) {
    u32 add = seed * 1000000u;
    u64 v = ((u64)(int64_t)(int32_t)index + (u64)add) * 3935559000370003845ULL + 2691343689449507681ULL;
    v ^= v >> 21; v ^= v << 37; v ^= v >> 4;
    v *= 4768777513237032717ULL;
    v ^= v << 20; v ^= v >> 41; v ^= v << 5;
    return (u32)v;
}

// ---- conversion rule tables (Param::SetAlign, param.cpp:163-263) -----------------------
struct RuleTables {
    u8 code[256];     // alphabet
    u8 rcode[256];    // rev_alphabet
    u8 reg[256];      // reg_alphabet (3 for ACGTacgt)
    u8 conv[256];     // alphabet_Mread
    u8 rconv[256];    // rev_alphabet_Mread
    char letter[8];   // useful_nt
    int single;       // one non-'-' convert-to base -> CountMismatch, else CountMismatch_new
};

// ---- device-resident index (replicated per GPU) ----------------------------------------
struct DevIndex {
    const u64 *plane[2];     // forward / reverse-complement 2-bit planes with 400-word margins
    const u32 *bucket;       // [2K+1]: bucket[2k]=first entry of k-mer k, bucket[2k+1]=end of its forward-strand entries
    const u8 *cnt8;          // [K] saturating bucket sizes for seed selection (0xFF -> use bucket[]); 43 MB at -s 16: L2 resident
    const u32 *loc;          // seed table entries (global coordinates), forward entries first per bucket; behind them, at rec_base, one
                             // 8-word record per k-mer: {entries, forward-strand entries, the entries if <= BSL_REC_INLINE else the first one's index}
    u32 rec_base;
    const u32 *anchor;       // [nseq+1] ref_anchor (refbase.cpp:222-226)
    const u32 *seqlen;       // [nseq]
    const u32 *rcoff;        // [nseq] RefTitle::rc_offset
    u32 nseq, K, maxk;
    u64 n_words, n_entries;
    // screening plane (single-conversion rules): ONE bit per forward-strand base = the low bit of its 2-bit code, which
    // the conversion cannot change (from = 01, to = 11). 32 bases per u32, first base in the top bit; n_words u32 words.
    const u32 *bit1;
    const u32 *nflag;        // 1 bit per 256-base sector of bit1: holds a non-ACGT base, padding or margin
    const uint2 *ctab;       // per 2^20 global coordinates: {2*anchor + rc_offset - 1, anchor + rc_offset} of the sequence that owns
                             // the whole block, {0, 0} when a boundary falls inside (then: binary search over anchor[])
    u32 flip;                // low bit of code(complement(x)) = low bit of code(x) ^ flip
    u32 has_bit1;
};
#define BSL_CTAB_SHIFT 16
#define BSL_REC_INLINE 6u

// ---- per-read ("slot") device state -------------------------------------------------------
// hit record: 16 bytes
//   x: loc            forward coordinate on the sequence
//   y: chr2 | level<<20 | chain<<24 | gapped<<25
//   z: (u32)(int)gap_size
//   w: gap_pos
struct __align__(16) DevHit { u32 loc, tag, gap, gp; };
#define HIT_CHR2(t)  ((t) & 0xFFFFFu)
#define HIT_LEVEL(t) (((t) >> 20) & 15u)
#define HIT_CHAIN(t) (((t) >> 24) & 1u)
#define HIT_GAPPED(t) (((t) >> 25) & 1u)

struct __align__(16) SlotMeta {
    u32 rnd;        // myrand(index)
    u16 len;        // mapped read length
    u8  B;          // mismatch budget
    u8  nseg;       // seedseg_num
    u8  flags;      // bit0 chain0 enabled, bit1 chain1 enabled, bit2 filtered, bit3 overflow (needs heavy pass), bit4 done
    u8  thr;        // current snp_thres
    u16 nhit;       // hits stored
    u32 item;       // this slot's hit storage: index into the hit pool in units of cap, or SLOT_BIG | index of a large block
};
#define SLOT_BIG 0x80000000u
#define SF_CHAIN0 1u
#define SF_CHAIN1 2u
#define SF_FILTERED 4u
#define SF_OVERFLOW 8u
#define SF_DONE 16u
#define SF_ABORT0 32u   // level-0 list reached -w
#define SF_CONTEXT 64u  // context read (bsl_batch::n_context): scheduled, never mapped
#define SF_STALE 128u   // its schedule reaches beyond the end of the read: those seed hashes come from KArgs::stale

struct SlotCounts { u16 c[2][16]; };   // hits per read chain and mismatch level

// ---- work counters -------------------------------------------------------------------------
// one non-empty seed bucket visited in a search round ("item"): owns the flat candidate range [base, base + m), where
// base + m is the base of the next item (items and candidates are allocated by one packed counter, so item order is
// candidate order). The cyclic bucket walk (align.cpp:293-296) starts at entry rot = myrand % m and forward-strand
// entries come first in a bucket, so the walk positions t = flat index - base that lie on the reverse strand form ONE
// interval [x1, x2) (inv = 0, rot < n_fwd) or everything but one interval (inv = 1): strand(t) = (x1 <= t < x2) ^ inv.
//   y: slot (22) | chain << 22 | thr << 23 | inv << 27 | phase << 28
//   z: x1 (23) | L << 23
//   w: x2 (23) | h << 23            (h = read offset of the seed: candidate start g = loc - h, align.cpp:297)
struct __align__(16) ItemHdr { u32 base, y, z, w; };
#define IH_SLOT(y)  ((y) & 0x3FFFFFu)
#define IH_CHAIN(y) (((y) >> 22) & 1u)
#define IH_THR(y)   (((y) >> 23) & 15u)
#define IH_INV(y)   (((y) >> 27) & 1u)
#define IH_PHASE(y) ((y) >> 28)
#define IH_X1(z)    ((z) & 0x7FFFFFu)
#define IH_L(z)     ((z) >> 23)
#define IH_X2(w)    ((w) & 0x7FFFFFu)
#define IH_H(w)     ((w) >> 23)
#define BSL_MAX_SLOTS (1u << 22)       // slots of one sub-range (ItemHdr::y)
#define BSL_MAX_BUCKET (1u << 23)      // largest bucket a walk may visit (ItemHdr x1 / x2)
__host__ __device__ __forceinline__ u32 ih_strand(u32 t, u32 y, u32 z, u32 w) { return ((t >= IH_X1(z) && t < IH_X2(w)) ? 1u : 0u) ^ IH_INV(y); }

// per search round (SE rounds use index r, PE rounds 20+r)
struct RoundCtr {
    unsigned long long alloc;   // items<<40 | candidates allocated by seed_lookup
    unsigned long long limit_inv; // 0, or ~(allocation state at which the flat candidate space ran out)
    u32 active;                 // length of the list this round consumes
    u32 flagged;                // reads left to reduce_round, listed from the back of KArgs::flag_list
    u32 work;                   // work-stealing cursor of reduce_round (grab number)
    u32 wide;                   // pairs of this round left to pair_round_wide (a mate's hit list is long or lives in a large block), front of KArgs::wide_list
    u32 long_n;                 // reads with many marked candidates / long hit lists, listed from the front of KArgs::flag_list: reduce_round starts them first, one per grab
    u32 long_work;              // pairs with a list beyond PW_CAP_SMALL, back of KArgs::wide_list
};
struct DevCounters {
    unsigned long long seed_lookups, candidates, hits_added, heavy, all_n;
    RoundCtr rc[40];
    u32 overflow_n;
    u32 defer_n;                // reads whose seed schedule is left to prepare_deferred
    u32 big_n;                  // large hit-list blocks handed out
};

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(ctx, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return (e_ == cudaErrorMemoryAllocation) ? BSL_ENOMEM : BSL_ECUDA; } } while (0)
