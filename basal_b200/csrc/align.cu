// align.cu — batched read mapping on the GPU (replaces SingleAlign/PairAlign::Do_Batch).
//
// Kernels (all sm_100a integer / LSU work, no tensor cores):
//   prepare_reads      a warp takes 32 reads: FilterReads, 2-bit planes and 1-bit streams of both chains, seed hashes,
//                      bucket-size gathers and the seed schedule (ConvertBinaySeq + ReorderSeed); prepare_deferred finishes
//                      the reads that inherit their start offset from earlier reads (SURVEY trap 3)
//   build_lists        first active lists (SE reads / full pairs / lone mates)
//   per search round r (= SnpAlign mode r of every still-active read, align.cpp:274-316):
//     seed_lookup      thread per (read, chain): the I look-ups of mode r (one 32-byte record per k-mer); every non-empty
//                      bucket becomes an "item" that owns a contiguous range of a flat candidate space; the bucket walks
//                      are copied out to flat_loc in visiting order
//     screen_bits      THE roofline kernel (single-conversion and '-'-only rules, with or without -g): warp per 32
//                      candidates, one 32-byte gather per candidate from the one-bit forward plane (L2 resident),
//                      XOR/popcount lower bound of CountMismatch; survivors are counted exactly on the 2-bit planes
//                      (drain_exact)
//     screen_candidates  the same for multi-way rules without -g, on the 2-bit planes
//     verify_candidates  multi-way rules with -g: whole window of every candidate (CountMismatch + GapAlign's first test)
//     reduce_fast      thread per read with <= 4 marked candidates: AddHit replay from the mark records; lists the reads
//                      it leaves to reduce_round
//     reduce_round     warp per listed read: replays the marked candidates, 32 per trip, in discovery order with the
//                      reference's AddHit semantics (dedup, -w feedback on the threshold, abort) and runs the single-gap
//                      search (GapAlign)
//     pair_round       PE: SortHits4PE + GetPairs replay for level r, thread per pair; pairs with long hit lists go to
//                      pair_round_wide<256> / <2048> (warp per pair)
//   finalize_reads     lowest non-empty level, -S tie-break, result records
//
// Discovery order inside a read is preserved exactly: the flat candidate index of a read's candidates grows in
// (chain, phase, rotated bucket index) order and the reducers walk the marked candidates in that order.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "stdsort.cuh"

namespace {

struct DevTables {
    RuleTables rule;
    u32 budget0[BSL_MAX_READLEN + 1];
    u16 prof[16][16];
    u32 tab_code[2];      // 4 x 2-bit codes indexed by (ascii>>1)&3 for chain 0 / chain 1 (alphabet / rev_alphabet)
    u32 tab_conv[2];      // same for alphabet_Mread / rev_alphabet_Mread
};

struct KArgs {
    DevIndex di;
    const DevTables *tab;
    u32 s, I, gap, w, min_insert, max_insert, chains, report, randseed, max_ns, min_read_size, single;
    // batch
    u32 n_slots, n_a;              // n_a = reads in batch a (PE: slots [n_a, 2 n_a) are the mates)
    u32 pe;                        // paired run
    u32 readset_a, readset_b;
    const u8 *bases; const u64 *off; const u32 *index; const u16 *rawlen; u32 first_index_a, first_index_b; u64 bases_b_shift;
    u32 has_index, has_rawlen;
    u32 n_ctx;                     // the first n_ctx reads of a batch (and of its mate batch) are context: scheduled, never mapped
    u32 carry;                     // the batch holds reads with an empty start-offset range: their schedules are made by prepare_deferred
    u8 *st0arr;                    // [slot*2+chain] start offset chosen by ReorderSeed for reads that have a range (carry mode)
    u32 *defer; u32 *stale;        // slots left to prepare_deferred; [slot*2+chain][16] seed hashes inherited beyond the end of the read
    u64 off_base_a, off_base_b;    // value of offsets[first] of the sub-range (offsets are passed as given)
    u64 all_off;                   // all-hits records already produced by earlier sub-ranges of this call
    u32 Wb;                        // words per plane in this batch
    u32 *bits1;                    // [slot*2+chain][2*Wb]: low code bit per base, then ACGT mask per base (32 bases per word, first base in the top bit)
    u64 *planes; u8 *sched; SlotMeta *meta; SlotCounts *cnt; uint2 *stat; u8 *minlvl;   // stat: executed seed look-ups / candidates per slot
    DevHit *hits; u32 cap;         // hit pool and per-slot capacity
    DevHit *bighits; u32 big_cap, big_blocks;   // large blocks for the few reads whose list outgrows `cap` (repeat families): handed out by add_hit
    u32 *wide_list;                // pairs routed to pair_round_wide
    DevCounters *ctr;
    // per-round scratch
    ItemHdr *hdr; u32 cap_items; u32 cap_cands;
    uint2 *slot_item;              // [slot*2+chain] = {first item, number of items} of this round
    u32 *slot_flag;                // number of candidates verify marked for the slot in this round
    u32 *flag_list;                // slots with marked candidates
    u32 *chunk_first;              // item that owns flat candidate 32 k, for every group k of 32 candidates
    u32 *bitmap;                   // 1 bit per flat candidate
    u32 *flat_loc;                 // seed-table entry of every flat candidate (the bucket walks, in visiting order)
    uint4 *marks;                  // [slot][MK_CAP] the first marked candidates of a slot in this round: {flat index, g, snp | strand<<8 | chain<<9, -}
    bsl_hit *out; bsl_pair *pair_out; bsl_hit *all_a; bsl_hit *all_b; u64 all_cap;
};

__device__ __forceinline__ u64 plane_extract(const u64 *pl, u32 p) {     // 32 bases starting at base p, MSB first
    u32 w = p >> 5, o = (p & 31u) * 2;
    u64 x = pl[w] << o;
    if (o) x |= pl[w + 1] >> (64 - o);
    return x;
}

// Read planes live in global memory as "streams": 64-bit words with their halves swapped, so that the same bytes read
// as a u32 array are in logical order (u32 j = bases 16j..16j+15, first base in the top bits), which is what
// verify_candidates stages. Plane order per (slot, chain): bases, N-mask reduced to 01 per ACGT base, convert-to mask.
__device__ __forceinline__ u64 swap32(u64 x) { return (x << 32) | (x >> 32); }
// hit list of a slot and its capacity
__device__ __forceinline__ DevHit *slot_hits(const KArgs &A, u32 item, u32 &cap) {
    if (item & SLOT_BIG) { cap = A.big_cap; return A.bighits + (u64)(item & ~SLOT_BIG) * A.big_cap; }
    cap = A.cap; return A.hits + (u64)item * A.cap;
}
__device__ __forceinline__ u64 stream_extract(const u64 *pl, u32 p) {    // plane_extract over a stream in global memory
    u32 w = p >> 5, o = (p & 31u) * 2;
    u64 x = swap32(pl[w]) << o;
    if (o) x |= swap32(pl[w + 1]) >> (64 - o);
    return x;
}

// ------------------------------------------------------------------------------------------------
// L2 cache-hint policies (createpolicy): `keep` for the tables every SM gathers from again and again (the one-byte
// bucket-size table, the one-bit screening plane), `stream` for data that passes through once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 l2_policy_keep() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_stream() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u32 ldg_u8_hint(const u8 *p, u64 pol) { u32 v; asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }
__device__ __forceinline__ uint2 ldg_v2_hint(const void *p, u64 pol) { uint2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol)); return v; }
__device__ __forceinline__ u32 ldg_u32_stream(const u32 *p, u64 pol) { u32 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }
__device__ __forceinline__ uint4 ldg_v4_stream(const void *p, u64 pol) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol)); return v; }
__device__ __forceinline__ u64 ldg_u64_stream(const u64 *p, u64 pol) { u64 v; asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v; }

// ------------------------------------------------------------------------------------------------
// prepare_reads : FilterReads (align.cpp:548-563), read planes (ConvertBinaySeq / ConvertBinarySeq, align.cpp:79-226)
// and the seed schedule (ReorderSeed / AdjustSeedStartArray / CountSeeds, align.cpp:468-546). A warp takes 32 reads:
//   A.  (lane per 32-base plane word of any of the 32 reads) five aligned 8-byte loads cover the 32 ASCII bases of the
//       word; four bases at a time are turned into 2-bit codes / ACGT flags / convert-to codes in registers (one PRMT
//       looks up the expected letter, one the code) and packed with a multiply; the words go to global memory (2-bit
//       streams, 1-bit streams) and to the warp's shared slice;
//       the 1-bit streams of the read taken backwards are bit-reversed funnel shifts of the forward ones, fetched by shuffle;
//   B.  (lane per read) per seed segment j: seed hashes and bucket sizes of the read offsets j*s .. j*s + I + ii - 1 (every
//       offset prof[j][i] + v - i the schedule can touch), eight gathers from the one-byte size table in flight, then
//       CountSeeds(j, v) for every start v <= ii from those sizes;
//   C.  (lane per read) ReorderSeed / AdjustSeedStartArray / the (count, segment) sort, literally, as look-ups in the
//       CountSeeds table of step B.
// Shared memory per lane: WQ + Wb + 1 + wd + nseg (ii + 1) words, odd stride so that per-lane rows sit in different banks.
// ------------------------------------------------------------------------------------------------
#define PR_WARPS 4
#define BASES_PAD 64        // bytes in front of and behind the batch's bases in device memory: load32 may touch up to 39 bytes either side

struct Conv4 { u32 cq, cc, r1; };      // per byte: 2-bit code, convert-to code, 0x01 for an ACGT letter
// four ASCII bases (byte k = base k); valid01 has 0x01 in the bytes that belong to the read
__device__ __forceinline__ Conv4 conv4(u32 w, u32 valid01, u32 tqb, u32 tcb) {
    const u32 v = (w >> 1) & 0x03030303u;                                              // (ascii >> 1) & 3: A 0, C 1, T 2, G 3 in either case
    const u32 t = (v | (v >> 4)) & 0x00330033u, sel = (t | (t >> 8)) & 0xFFFFu;         // one PRMT selector nibble per byte
    const u32 y = (w & 0xDFDFDFDFu) ^ __byte_perm(0x47544341u, 0u, sel);               // zero byte <=> the letter is the A / C / T / G its bits select
    const u32 z = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
    Conv4 o; o.r1 = (z >> 7) & valid01;
    const u32 m3 = o.r1 * 3u;
    o.cq = __byte_perm(tqb, 0u, sel) & m3; o.cc = __byte_perm(tcb, 0u, sel) & m3;
    return o;
}
__device__ __forceinline__ u32 pack4x2(u32 x) { return (x * 0x40100401u) >> 24; }      // bytes b0..b3 (2 bits each) -> b0 b1 b2 b3, first base in the top bits
__device__ __forceinline__ u32 pack4x1(u32 x) { return (x * 0x08040201u) >> 24; }      // bytes b0..b3 (1 bit each) -> 4 bits, first base in the top bit
// 2-bit table indexed by (ascii >> 1) & 3 -> the same table with one byte per entry
__device__ __forceinline__ u32 tab_bytes(u32 t2) { return (t2 & 3u) | ((t2 & 12u) << 6) | ((t2 & 48u) << 12) | ((t2 & 192u) << 18); }

// 32 bytes starting at p (any alignment) as eight little-endian words; touches [p & ~7, (p & ~7) + 40)
__device__ __forceinline__ void load32(const u8 *p, u32 (&w)[8]) {
    const u64 a = (u64)p; const uint2 *q = (const uint2 *)(a & ~7ull);
    const uint2 x0 = __ldcs(q), x1 = __ldcs(q + 1), x2 = __ldcs(q + 2), x3 = __ldcs(q + 3), x4 = __ldcs(q + 4);
    const u32 x[10] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y, x4.x, x4.y};
    const bool odd = (a >> 2) & 1u; const u32 sh = ((u32)a & 3u) * 8u;
    u32 y[9];
#pragma unroll
    for (int j = 0; j < 9; j++) y[j] = odd ? x[j + 1] : x[j];
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = __funnelshift_r(y[j], y[j + 1], sh);
}

// CountSeeds(j, v) (align.cpp:526-540) for every start v <= vmax, from the bucket sizes (bit 31 = the seed holds a non-ACGT
// base) of the read offsets j*s + d held in cwj[d]
__device__ __forceinline__ void count_seeds_row(const u32 *cwj, const u16 *profj, u32 j, u32 s, u32 I, u32 vmax, u32 *csj) {
    if (I <= 4) {                                                      // the usual interval: the four offsets stay in registers
        u32 d[4];
#pragma unroll
        for (u32 i = 0; i < 4; i++) d[i] = i < I ? profj[i] - i - j * s : 0u;
        for (u32 v = 0; v <= vmax; v++) {
            u32 total = 0, k = 0;
#pragma unroll
            for (u32 i = 0; i < 4; i++) if (i < I) {
                const u32 e = cwj[d[i] + v];
                if (e >> 31) k = 12;
                total += (e & 0x7fffffffu) << k;
            }
            csj[v] = total == 0 ? 9999999u : total;
        }
        return;
    }
    for (u32 v = 0; v <= vmax; v++) {
        u32 total = 0, k = 0;
        for (u32 i = 0; i < I; i++) {
            const u32 e = cwj[profj[i] + v - i - j * s];
            if (e >> 31) k = 12;
            total += (e & 0x7fffffffu) << k;
        }
        csj[v] = total == 0 ? 9999999u : total;
    }
}
// ReorderSeed / AdjustSeedStartArray / the (count, segment) sort (align.cpp:468-524), literally, as look-ups in the CountSeeds
// table cs[j * ncol + v]. A read with an empty start-offset range (ii == 0) starts from `carried` (align.cpp:476-480 never
// assigns xseed_start_offset then). keys: nseg scratch words or null. Returns the 16 schedule bytes (byte t = segment of
// rank t | its start << 4).
__device__ __forceinline__ uint4 schedule_from_table(const u32 *cs, u32 *keys, u32 ncol, u32 nseg, u32 ii, u32 carried, u32 &st0_out) {
    u32 st0 = carried, best = 0xffffffffu;
    for (u32 v = 0; v < ii; v++) { u32 tt = 0; for (u32 j = 0; j < nseg; j++) tt += cs[j * ncol + v]; if (tt < best) { best = tt; st0 = v; } }
    st0_out = st0;
    u64 stp = 0;                                                       // start[j] in 4 bits each
    for (u32 j = 0; j < nseg; j++) stp |= (u64)st0 << (4 * j);
    for (u32 t = 0; t < nseg; t++) {                                   // AdjustSeedStartArray
        const u32 ptr = (t & 1) ? nseg - 1 - t / 2 : t / 2;
        const u32 lo_ = ptr == 0 ? 0 : (u32)(stp >> (4 * (ptr - 1))) & 15u, hi_ = ptr == nseg - 1 ? ii : (u32)(stp >> (4 * (ptr + 1))) & 15u;
        u32 pick = lo_, b = 0xffffffffu;
        for (u32 x = lo_; x <= hi_; x++) { const u32 tt = cs[ptr * ncol + x]; if (tt < b) { b = tt; pick = x; } }
        stp = (stp & ~(15ull << (4 * ptr))) | ((u64)pick << (4 * ptr));
    }
    if (keys) for (u32 j = 0; j < nseg; j++) keys[j] = cs[j * ncol + ((u32)(stp >> (4 * j)) & 15u)];
    u32 sb[4] = {0, 0, 0, 0};
    for (u32 j = 0; j < nseg; j++) {
        const u32 stj = (u32)(stp >> (4 * j)) & 15u;
        const int kj = (int)cs[j * ncol + stj]; u32 rank = 0;          // pair<int,int> order: (count, segment)
        if (keys) for (u32 y = 0; y < nseg; y++) { const int ky = (int)keys[y]; rank += (ky < kj || (ky == kj && y < j)) ? 1u : 0u; }
        else for (u32 y = 0; y < nseg; y++) { const int ky = (int)cs[y * ncol + ((u32)(stp >> (4 * y)) & 15u)]; rank += (ky < kj || (ky == kj && y < j)) ? 1u : 0u; }
        const u32 v = (j | (stj << 4)) << ((rank & 3u) * 8);
        if ((rank >> 2) == 0) sb[0] |= v; else if ((rank >> 2) == 1) sb[1] |= v; else if ((rank >> 2) == 2) sb[2] |= v; else sb[3] |= v;
    }
    return make_uint4(sb[0], sb[1], sb[2], sb[3]);
}

__global__ void __launch_bounds__(PR_WARPS * 32, 8) prepare_reads(const __grid_constant__ KArgs A, u32 WQ, u32 WDM, u32 CSZ) {
    extern __shared__ u32 psm[];
    __shared__ u16 s_prof[16][16];
    const DevTables *T = A.tab;
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    for (u32 x = threadIdx.x; x < 256; x += blockDim.x) s_prof[x >> 4][x & 15u] = T->prof[x >> 4][x & 15u];
    __syncthreads();
    const u32 Wb = A.Wb, W2 = 2 * Wb, I = A.I, s = A.s;
    const u32 RS = (WQ + Wb + 1 + WDM + CSZ) | 1u;                                  // words per lane row: 2-bit codes, ACGT bits, one segment of bucket sizes, CountSeeds table
    u32 *wsm = psm + (size_t)wid * 32 * RS;
    u32 *row = wsm + lane * RS;
    u32 *sq = row, *smk = row + WQ, *cw = smk + Wb + 1, *cs = cw + WDM;
    const u32 tabq[2] = {T->tab_code[0], T->tab_code[1]}, tabc[2] = {T->tab_conv[0], T->tab_conv[1]};
    const u32 shs = 32 - 2 * s, nfull = (1u << s) - 1u;
    const u32 flipm = A.di.flip ? 0xffffffffu : 0u;
    const u64 keep = l2_policy_keep();
    const u32 rpt = 32u / Wb;                                                       // reads per trip of step A: the words of a read stay in one trip
    const u32 a_rl = lane / Wb, a_wv = lane - a_rl * Wb;
    for (u32 slot0 = (blockIdx.x * PR_WARPS + wid) * 32u; slot0 < A.n_slots; slot0 += gridDim.x * PR_WARPS * 32u) {
        const u32 slot = slot0 + lane;
        const bool valid = slot < A.n_slots;
        u64 b0 = 0; u32 Lraw = 0, readset = 0, index = 0;
        if (valid) {
            const bool mate_b = A.pe && slot >= A.n_a;
            const u32 r = mate_b ? slot - A.n_a : slot;
            const u64 *off = mate_b ? A.off + (A.n_a + 1) : A.off;
            const u64 ob = mate_b ? A.off_base_b : A.off_base_a;
            b0 = off[r] - ob + (mate_b ? A.bases_b_shift : 0);
            Lraw = (u32)(off[r + 1] - off[r]);
            readset = mate_b ? A.readset_b : A.readset_a;
            index = A.has_index ? A.index[slot] : (mate_b ? A.first_index_b : A.first_index_a) + r;
        }
        const u32 L = Lraw > BSL_MAX_READLEN ? BSL_MAX_READLEN : Lraw;
        u32 flags = 0;
        if (valid) {
            if ((A.chains == 1) || ((A.chains <= 1) == (readset < 2))) flags |= SF_CHAIN0;                       // align.cpp:83-84
            if ((A.chains == 1) || ((A.chains <= 1) == (readset == 2))) flags |= SF_CHAIN1;
        }
        // per-level counters, stats: 32 slots x 64 bytes are contiguous
        {
            u32 *cz = (u32 *)(A.cnt + slot0); const u32 nz = min(32u, A.n_slots - slot0) * 16u;
            for (u32 x = lane; x < nz; x += 32) cz[x] = 0;
            if (valid) { A.stat[slot] = make_uint2(0u, 0u); A.minlvl[slot] = 255; }
        }
        u32 B = 0, nseg = 0; bool filtered = false, ns_known = false, deferred = false;
        const u32 ii = (L + 1 >= I) ? (L + 1 - I) % s : 0;
        for (u32 c = 0; c < 2; c++) {
            const bool en = valid && (flags & (c ? SF_CHAIN1 : SF_CHAIN0)) && !filtered;
            if (!__any_sync(0xffffffffu, en)) continue;
            // ---- A: one plane word (32 bases) of one read per lane; rpt reads per trip
            const u32 tqb = tab_bytes(tabq[c]), tcb = tab_bytes(tabc[c]);
            for (u32 r0 = 0; r0 < 32; r0 += rpt) {
                const u32 rl = min(r0 + a_rl, 31u), wv = a_wv;
                const bool mine = a_rl < rpt && r0 + a_rl < 32u;
                const u32 Lr = __shfl_sync(0xffffffffu, L, rl);
                const u32 b0l = __shfl_sync(0xffffffffu, (u32)b0, rl), b0h = __shfl_sync(0xffffffffu, (u32)(b0 >> 32), rl);
                const bool enr = (__shfl_sync(0xffffffffu, (u32)en, rl) != 0) && mine;
                const u8 *src = A.bases + (((u64)b0h << 32) | b0l);
                const u32 p0 = wv * 32;
                const u32 nbase = (enr && Lr > p0) ? min(32u, Lr - p0) : 0u;
                u32 qh = 0, ql = 0, nh = 0, nl = 0, ch = 0, cl = 0, lo = 0, mk = 0;
                if (nbase) {
                    // chain 0: bases p0 .. p0+31 in order; chain 1: the reverse complement, i.e. bytes Lr-1-p0 down to Lr-32-p0
                    u32 w[8];
                    if (c == 0) load32(src + p0, w);
                    else {
                        u32 t[8]; load32(src + (long long)Lr - 32 - (long long)p0, t);       // may start before the read: those bytes are masked out below
#pragma unroll
                        for (int j = 0; j < 8; j++) w[j] = __byte_perm(t[7 - j], 0u, 0x0123u);
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const u32 nb = nbase > 4u * j ? nbase - 4u * j : 0u;
                        const u32 v01 = nb >= 4u ? 0x01010101u : (0x01010101u & ((1u << (8u * nb)) - 1u));
                        const Conv4 k4 = conv4(w[j], v01, tqb, tcb);
                        const u32 q8 = pack4x2(k4.cq), n8 = pack4x2(k4.r1), c8 = pack4x2(k4.cc);
                        if (j < 4) { qh = (qh << 8) | q8; nh = (nh << 8) | n8; ch = (ch << 8) | c8; }
                        else { ql = (ql << 8) | q8; nl = (nl << 8) | n8; cl = (cl << 8) | c8; }
                        lo = (lo << 4) | pack4x1(k4.cq & 0x01010101u); mk = (mk << 4) | pack4x1(k4.r1);
                    }
                }
                // the two 1-bit streams of the read taken backwards (what screen_bits compares with a forward-plane sector when the
                // candidate lies on the reverse strand): reversed word wv = forward bases s0+31 down to s0, from the lanes that hold
                // forward words s0 >> 5 and (s0 >> 5) + 1 of the same read; the low bits carry the complement flip already
                u32 rlo, rmk;
                {
                    const int s0 = (int)Lr - 32 - 32 * (int)wv, w0 = s0 >> 5; const u32 sf = (u32)s0 & 31u;
                    const u32 lb_ = lane - wv;                                           // lane of word 0 of this read
                    const u32 la = __shfl_sync(0xffffffffu, lo, (lb_ + (u32)max(w0, 0)) & 31u), lb = __shfl_sync(0xffffffffu, lo, (lb_ + (u32)min(w0 + 1, (int)Wb - 1)) & 31u);
                    const u32 ma = __shfl_sync(0xffffffffu, mk, (lb_ + (u32)max(w0, 0)) & 31u), mb = __shfl_sync(0xffffffffu, mk, (lb_ + (u32)min(w0 + 1, (int)Wb - 1)) & 31u);
                    const bool oka = w0 >= 0, okb = w0 + 1 >= 0 && w0 + 1 < (int)Wb;
                    rlo = s0 > -32 ? __funnelshift_r(oka ? __brev(la) : 0u, okb ? __brev(lb) : 0u, sf) : 0u;
                    rmk = s0 > -32 ? __funnelshift_r(oka ? __brev(ma) : 0u, okb ? __brev(mb) : 0u, sf) : 0u;
                }
                if (!enr) continue;
                const u32 slr = slot0 + rl;
                u32 *dst = (u32 *)(A.planes + ((u64)slr * 2 + c) * 3 * Wb);           // streams: logical 32-bit words, see KArgs::planes
                __stcs((uint2 *)(dst + 2 * wv), make_uint2(qh, ql));
                __stcs((uint2 *)(dst + W2 + 2 * wv), make_uint2(nh, nl));
                __stcs((uint2 *)(dst + 2 * W2 + 2 * wv), make_uint2(ch, cl));
                u32 *b1 = A.bits1 + ((u64)slr * 2 + c) * 2 * W2;
                __stcs((uint2 *)(b1 + 2 * wv), make_uint2(lo, mk));
                __stcs((uint2 *)(b1 + W2 + 2 * wv), make_uint2(rlo ^ flipm, rmk));
                u32 *r_ = wsm + rl * RS;
                r_[2 * wv] = qh; r_[2 * wv + 1] = ql; r_[WQ + wv] = mk;
            }
            sq[W2] = 0; smk[Wb] = 0;
            __syncwarp();
            if (en && !ns_known) {
                ns_known = true;
                u32 acgt = 0; for (u32 j = 0; j < Wb; j++) acgt += __popc(smk[j]);
                const u32 ns = L - acgt;
                filtered = (L == 0) || (L < A.min_read_size) || (ns > A.max_ns) || (Lraw > BSL_MAX_READLEN);     // align.cpp:559-560
                if (!filtered) {
                    u32 raw = A.has_rawlen ? A.rawlen[slot] : L; if (raw == 0 || raw > BSL_MAX_READLEN) raw = L ? L : 1;
                    B = (T->budget0[raw] + 1) * (L - 1) / raw;                                                    // align.cpp:561
                    nseg = (L + 1 >= I + s) ? min((L + 1 - I) / s, B + 1) : 0;                                    // align.cpp:450
                }
            }
            if (en && !filtered && nseg > 0 && A.carry && ii == 0) deferred = true;          // inherits its start offset: prepare_deferred
            else if (en && !filtered && nseg > 0) {
                const u32 vmax = ii, wd = I + vmax, ncol = vmax + 1;
                // ---- B: per segment, bucket sizes (bit 31 = the seed holds a non-ACGT base) of offsets j*s + d, d < wd, eight gathers
                //      from the one-byte size table in flight, then CountSeeds(j, v) (align.cpp:526-540) for every start v <= vmax
                auto seed_at = [&](u32 p, u32 &flag) -> u32 {
                    const u32 w = p >> 4, o = (p & 15u) * 2, wm = p >> 5, om = p & 31u;
                    const u32 xq = __funnelshift_l(sq[w + 1], sq[w], o) >> shs, xm = __funnelshift_l(smk[wm + 1], smk[wm], om) >> (32 - s);
                    flag = xm != nfull ? 0x80000000u : 0u;
                    return bsl_xt(xq);
                };
                if (wd <= 8) {
                    // the usual case (I + ii <= 8): one batch of eight gathers per segment; the batch of segment j+1 is issued before
                    // the sizes of segment j are consumed, so its latency hides behind CountSeeds; a saturated size (0xFF: reads from repeat
                    // families, but with 224 look-ups per warp and segment nearly every warp meets one) asks the k-mer's record for the
                    // exact one, again one segment ahead of its use
                    // the s + 7 bases the eight seeds of a segment cover are read once; the hash of offset u + 1 follows from the one
                    // of offset u (drop the leading base-3 digit, append one)
                    u32 pw3 = 1; for (u32 x = 1; x < s; x++) pw3 *= 3u;
                    auto hashes = [&](u32 j, u32 (&hk)[8], u32 &flb) {
                        const u32 p0 = j * s, w = p0 >> 4, o = (p0 & 15u) * 2, wm = p0 >> 5, om = p0 & 31u;
                        const u32 q0 = sq[w], q1 = sq[w + 1], q2 = w + 2 <= W2 ? sq[w + 2] : 0u;
                        u32 f0 = __funnelshift_l(q1, q0, o), f1 = __funnelshift_l(q2, q1, o);        // bases p0 .. p0+15, p0+16 .. p0+31
                        f0 -= (f0 << 1) & f0 & 0xAAAAAAAAu; f1 -= (f1 << 1) & f1 & 0xAAAAAAAAu;      // digits: 11 -> 01 (the first step of XT)
                        const u32 fin = s >= 16 ? f1 : __funnelshift_l(f1, f0, 2 * s);             // digits p0+s .. : what the window takes in
                        const u32 zm = ~__funnelshift_l(smk[wm + 1], smk[wm], om);                 // a set bit = a non-ACGT base, from p0
                        u32 h = bsl_xt(f0 >> shs);
                        flb = 0;
#pragma unroll
                        for (u32 u = 0; u < 8; u++) {
                            if (u) h = (h - ((f0 >> (32 - 2 * u)) & 3u) * pw3) * 3u + ((fin >> (32 - 2 * u)) & 3u);
                            hk[u] = h;
                            flb |= (((zm << u) >> (32 - s)) != 0u ? 1u : 0u) << u;
                        }
                    };
                    auto issue = [&](u32 j, u32 (&c8)[8], u32 &flb) {
                        u32 hk[8]; hashes(j, hk, flb);
#pragma unroll
                        for (u32 u = 0; u < 8; u++) { c8[u] = 0; if (u < wd) c8[u] = ldg_u8_hint(A.di.cnt8 + hk[u], keep); }
                    };
                    const u32 *rec = A.di.loc + A.di.rec_base;
                    u32 r8[8], fr = 0;                           // one-byte sizes of the segment after the one being packed (in flight), its non-ACGT flags
                    u32 ex[8];                                   // exact sizes of the saturated offsets of the segment to consume next (in flight)
                    u32 cap0 = 0, cap1 = 0, fa = 0, sa = 0;      // the segment to consume: one-byte sizes (packed), flags, which offsets are saturated
#pragma unroll
                    for (u32 u = 0; u < 8; u++) ex[u] = 0;
                    issue(0, r8, fr);
                    // -I 4 with -s a multiple of 4 (the defaults): CountSeeds(j, v) reads offsets v, v+3, v+2, v+1 of the segment, in that order
                    const bool quad = I == 4u && (s & 3u) == 0u;
                    for (u32 j = 0; j <= nseg; j++) {
                        // (a) segment j-1: its sizes are complete
                        u32 fv[8];
#pragma unroll
                        for (u32 u = 0; u < 8; u++) {
                            u32 v = ((u < 4 ? cap0 : cap1) >> (8u * (u & 3u))) & 0xFFu;
                            if ((sa >> u) & 1u) v = ex[u] & 0x7fffffffu;
                            fv[u] = u < wd ? v | (((fa >> u) & 1u) << 31) : 0u;
                        }
                        if (j > 0 && !quad) {
#pragma unroll
                            for (u32 u = 0; u < 8; u++) if (u < wd) cw[u] = fv[u];
                        }
                        // (b) segment j: the one-byte sizes have arrived; exact sizes of the saturated ones requested
                        u32 sb = 0, n0 = 0, n1 = 0; const u32 fn = fr;
                        if (j < nseg) {
#pragma unroll
                            for (u32 u = 0; u < 8; u++) if (u < wd) {
                                if (u < 4) n0 |= r8[u] << (8u * u); else n1 |= r8[u] << (8u * (u - 4u));
                                if (r8[u] == 0xFFu) sb |= 1u << u;
                            }
                            if (sb) {
                                u32 hk[8], fl_; hashes(j, hk, fl_);
#pragma unroll
                                for (u32 u = 0; u < 8; u++) if ((sb >> u) & 1u) ex[u] = __ldg(rec + 8 * (size_t)hk[u]);
                            }
                        }
                        // (c) one-byte sizes of segment j+1
                        if (j + 1 < nseg) issue(j + 1, r8, fr);
                        // (d) CountSeeds of segment j-1
                        if (j > 0) {
                            if (quad) {
                                u32 *csj = cs + (j - 1) * ncol;
#pragma unroll
                                for (u32 v = 0; v < 5; v++) if (v <= vmax) {                    // wd = 4 + vmax <= 8
                                    const u32 e0 = fv[v], e1 = fv[v + 3], e2 = fv[v + 2], e3 = fv[v + 1];
                                    u32 k = (e0 >> 31) ? 12u : 0u, total = (e0 & 0x7fffffffu) << k;
                                    if (e1 >> 31) k = 12u; total += (e1 & 0x7fffffffu) << k;
                                    if (e2 >> 31) k = 12u; total += (e2 & 0x7fffffffu) << k;
                                    if (e3 >> 31) k = 12u; total += (e3 & 0x7fffffffu) << k;
                                    csj[v] = total == 0u ? 9999999u : total;
                                }
                            } else count_seeds_row(cw, s_prof[j - 1], j - 1, s, I, vmax, cs + (j - 1) * ncol);
                        }
                        cap0 = n0; cap1 = n1; fa = fn; sa = sb;
                    }
                } else for (u32 j = 0; j < nseg; j++) {
                    for (u32 d0 = 0; d0 < wd; d0 += 8) {
                        u32 c8[8], flb = 0;
#pragma unroll
                        for (u32 u = 0; u < 8; u++) {
                            u32 fl; const u32 k = seed_at(j * s + min(d0 + u, wd - 1), fl);
                            c8[u] = ldg_u8_hint(A.di.cnt8 + k, keep); flb |= (fl >> 31) << u;
                        }
#pragma unroll
                        for (u32 u = 0; u < 8; u++) if (d0 + u < wd) cw[d0 + u] = c8[u] | (((flb >> u) & 1u) << 31);
                    }
                    for (u32 d = 0; d < wd; d++) if ((cw[d] & 0xFFu) == 0xFFu) {      // saturated: the exact size from the bucket table (rare)
                        u32 fl; const u32 k = seed_at(j * s + d, fl);
                        cw[d] = (__ldg(A.di.loc + A.di.rec_base + 8 * (size_t)k) & 0x7fffffffu) | fl;
                    }
                    count_seeds_row(cw, s_prof[j], j, s, I, vmax, cs + j * ncol);
                }
                // ---- C: the schedule
                u32 st0;
                // the sort keys go where the read's 2-bit words were (step A of the next chain writes them again)
                const uint4 sbytes = schedule_from_table(cs, nseg <= WQ ? sq : (nseg <= WDM ? cw : nullptr), ncol, nseg, ii, 0u, st0);
                if (A.carry) A.st0arr[(u64)slot * 2 + c] = (u8)st0;
                *(uint4 *)(A.sched + ((u64)slot * 2 + c) * 16) = sbytes;
            }
            __syncwarp();
        }
        if (valid) {
            if (filtered) flags |= SF_FILTERED;
            if ((A.pe && slot >= A.n_a ? slot - A.n_a : slot) < A.n_ctx) flags |= SF_CONTEXT;
            SlotMeta m; m.rnd = bsl_rand(index, A.randseed); m.len = (u16)L; m.B = (u8)B; m.nseg = (u8)nseg; m.flags = (u8)flags; m.thr = (u8)B; m.nhit = 0; m.item = slot;
            A.meta[slot] = m;
            if (deferred) A.defer[atomicAdd(&A.ctr->defer_n, 1u)] = slot;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// prepare_deferred : the seed schedule of reads whose start-offset range is empty ((len - I + 1) % s == 0). The reference
// never assigns xseed_start_offset for them (align.cpp:476-480), so they keep the value of the last read of the same
// aligner object and chain that had a range, AdjustSeedStartArray then moves their segments up to that offset, and seed
// positions beyond len - s still hold the hashes / non-ACGT flags of the last read long enough to have written them
// (xseed_array / xseedreg_array are never cleared, align.cpp:79-150). One aligner object sees the reads of one mate in
// batch order (-p 1 semantics; fresh, zeroed objects per call). Runs after prepare_reads: lane per deferred read.
// ------------------------------------------------------------------------------------------------
#define PD_WARPS 2
__global__ void __launch_bounds__(PD_WARPS * 32) prepare_deferred(const __grid_constant__ KArgs A, u32 WDM, u32 CSZ) {
    extern __shared__ u32 dsm[];
    __shared__ u16 s_prof[16][16];
    const DevTables *T = A.tab;
    for (u32 x = threadIdx.x; x < 256; x += blockDim.x) s_prof[x >> 4][x & 15u] = T->prof[x >> 4][x & 15u];
    __syncthreads();
    const u32 I = A.I, s = A.s, Wb = A.Wb, shs = 64 - 2 * s;
    const u32 RS = (WDM + CSZ) | 1u;
    u32 *cw = dsm + (size_t)threadIdx.x * RS, *cs = cw + WDM;
    const u32 n_def = A.ctr->defer_n;
    // seed hash (bit 31: the seed holds a non-ACGT base) at offset p of the read in `slot`; offsets beyond the end of the read
    // belong to the nearest earlier unfiltered read of the same mate that is long enough, or to nobody (zeroed memory)
    auto seed_of = [&](u32 slot, u32 c, u32 L, u32 base, u32 p) -> u32 {
        u32 src = slot;
        if (p + s > L) {
            src = 0xffffffffu;
            for (u32 t = slot; t-- > base;) { const SlotMeta d = A.meta[t]; if (!(d.flags & SF_FILTERED) && (u32)d.len >= p + s) { src = t; break; } }
            if (src == 0xffffffffu) return 0u;
        }
        const u64 *pl = A.planes + ((u64)src * 2 + c) * 3 * Wb;
        const bool acgt = (stream_extract(pl + Wb, p) >> shs) == (0x5555555555555555ULL >> shs);       // second stream: 01 per ACGT base
        return bsl_xt((u32)(stream_extract(pl, p) >> shs)) | (acgt ? 0u : 0x80000000u);
    };
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n_def; k += gridDim.x * blockDim.x) {
        const u32 slot = A.defer[k];
        const SlotMeta m = A.meta[slot];
        const u32 L = m.len, nseg = m.nseg, base = (A.pe && slot >= A.n_a) ? A.n_a : 0u;
        for (u32 c = 0; c < 2; c++) {
            if (!(m.flags & (c ? SF_CHAIN1 : SF_CHAIN0))) continue;
            u32 carried = 0;                                           // xseed_start_offset as the previous reads left it
            for (u32 t = slot; t-- > base;) {
                const SlotMeta d = A.meta[t];
                if (d.flags & SF_FILTERED) continue;
                const u32 Ld = d.len;
                if (Ld + 1 >= I && (Ld + 1 - I) % s != 0) { carried = A.st0arr[(u64)t * 2 + c]; break; }
            }
            const u32 vmax = carried, wd = I + vmax, ncol = vmax + 1;
            for (u32 j = 0; j < nseg; j++) {
                for (u32 d = 0; d < wd; d++) {
                    const u32 e = seed_of(slot, c, L, base, j * s + d), kmer = e & 0x7fffffffu;
                    u32 cnt = A.di.cnt8[kmer]; if (cnt == 0xFFu) cnt = A.di.loc[A.di.rec_base + 8 * (size_t)kmer];
                    cw[d] = (cnt & 0x7fffffffu) | (e & 0x80000000u);
                }
                count_seeds_row(cw, s_prof[j], j, s, I, vmax, cs + j * ncol);
            }
            u32 st0;
            const uint4 sbytes = schedule_from_table(cs, nullptr, ncol, nseg, 0u, carried, st0);
            *(uint4 *)(A.sched + ((u64)slot * 2 + c) * 16) = sbytes;
            for (u32 x = 0; x < 16; x++) A.stale[((u64)slot * 2 + c) * 16 + x] = x < vmax ? (seed_of(slot, c, L, base, L - s + 1 + x) & 0x7fffffffu) : 0u;
        }
        A.meta[slot].flags = (u8)(m.flags | SF_STALE);
    }
}

// Builds the first active lists. SE: every passing read. PE: full pairs -> pair list, lone passing mates -> SE list.
__global__ void build_lists(const __grid_constant__ KArgs A, u32 *se_list, u32 *pe_list) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!A.pe) {
        if (i >= A.n_slots) return;
        SlotMeta m = A.meta[i];
        if (!(m.flags & (SF_FILTERED | SF_CONTEXT)) && m.nseg > 0) { u32 pos = atomicAdd(&A.ctr->rc[0].active, 1u); se_list[pos] = i; }
        return;
    }
    if (i >= A.n_a) return;
    SlotMeta ma = A.meta[i], mb = A.meta[i + A.n_a];
    // a pair with one filtered mate stays in the pair list: seed_lookup<PE> gives its lone mate SingleAlign's stop rule and
    // pair_round leaves its hit lists alone, so the paired rounds also do what the reference's `_sa.RunAlign` does (pairs.cpp:197-201)
    if ((ma.flags | mb.flags) & SF_CONTEXT) return;
    const bool oka = !(ma.flags & SF_FILTERED), okb = !(mb.flags & SF_FILTERED);
    if (oka || okb) { u32 pos = atomicAdd(&A.ctr->rc[20].active, 1u); pe_list[pos] = i; }
}

// ------------------------------------------------------------------------------------------------
// seed_lookup : the bucket look-ups of mode `round` (align.cpp:279-292) for every active (read, chain)
// ------------------------------------------------------------------------------------------------
#define LK_THREADS 256
#define LK_BIG 64u            // a bucket with more entries is copied by the whole warp on its own
#define CHUNK 256            // candidates per verify chunk
#define ALLOC_SHIFT 40       // RoundCtr::alloc = items << 40 | candidates
#define ALLOC_MASK ((1ULL << ALLOC_SHIFT) - 1)
#define MAX_ITEMS_PER_ROUND ((1u << 24) - (1u << 16))

template <bool PE>
__global__ void __launch_bounds__(LK_THREADS, 4) seed_lookup(const __grid_constant__ KArgs A, u32 round, const u32 *list_in, u32 *list_out, u32 ci) {
    __shared__ u32 s_wi[LK_THREADS / 32], s_wc[LK_THREADS / 32];
    __shared__ u32 s_ibase, s_cbase, s_ok;
    const DevTables *T = A.tab;
    RoundCtr *rc = A.ctr->rc + ci;
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr u32 SH = PE ? 2 : 1;
    const u32 n_in = rc->active;
    const u32 n_thr = n_in << SH;
    const u32 shs = 64 - 2 * A.s;
    for (u32 k0 = blockIdx.x * LK_THREADS; k0 < n_thr; k0 += gridDim.x * LK_THREADS) {
        const u32 k = k0 + threadIdx.x;
        const bool act = k < n_thr;
        const u32 entry = k >> SH, c = k & 1u, mate = PE ? (k >> 1) & 1u : 0u;
        u32 slot = 0, mlvl = 255; SlotMeta m; m.flags = 0; m.nseg = 0; m.rnd = 0; m.len = 0; m.thr = 0;
        bool search = false, keep = false;
        if (act) {
            slot = list_in[entry] + mate * A.n_a;
            m = A.meta[slot];
            bool cont = !(m.flags & (SF_OVERFLOW | SF_FILTERED));
            if (!PE) cont = cont && round < m.nseg && A.minlvl[slot] >= round;          // stop rule of RunAlign (align.cpp:459-463)
            mlvl = A.minlvl[slot];
            search = cont && round < m.nseg && (m.flags & (c ? SF_CHAIN1 : SF_CHAIN0));
            keep = !PE && c == 0 && cont;
        }
        if (PE) {                                                                       // the mate's threads are two lanes away
            const u32 mate_flags = __shfl_xor_sync(0xffffffffu, (u32)m.flags, 2);
            if (act && (mate_flags & SF_FILTERED) && mlvl < round) search = false;          // lone mate: RunAlign's stop rule (align.cpp:459-463)
        }
        if (!PE) {                                                                      // compact the list of reads searched in this round
            const u32 bal = __ballot_sync(0xffffffffu, keep);
            if (bal) {
                u32 base = 0; const u32 leader = __ffs(bal) - 1;
                if (lane == leader) base = atomicAdd(&rc[1].active, (u32)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (keep) list_out[base + __popc(bal & ((1u << lane) - 1u))] = slot;
            }
        }
        // ---- pass 1: bucket sizes
        u32 tot = 0, nne = 0, j = 0, stj = 0; const u64 *pq = nullptr;
        // seed hash at read offset h (xseeds, align.cpp:487-490); offsets beyond the end of the read exist only for reads whose
        // schedule was made by prepare_deferred, and hold the hashes an earlier read left there
        auto kmer_at = [&](u32 h) -> u32 {
            if ((m.flags & SF_STALE) && h + A.s > (u32)m.len) return A.stale[((u64)slot * 2 + c) * 16 + (h + A.s - (u32)m.len - 1u)];
            return bsl_xt((u32)(stream_extract(pq, h) >> shs));
        };
        // a look-up = the first two words of the k-mer's record (entries, forward-strand entries); where the entries are follows in
        // pass 2. The look-ups of the first four phases stay in registers.
        const u32 *rec = A.di.loc + A.di.rec_base;
        u32 kpm[4], knf[4], kk[4];
#pragma unroll
        for (u32 i = 0; i < 4; i++) { kpm[i] = 0; knf[i] = 0; kk[i] = 0; }
        if (search) {
            const u8 sc = A.sched[((u64)slot * 2 + c) * 16 + round]; j = sc & 15u; stj = sc >> 4;
            pq = A.planes + ((u64)slot * 2 + c) * 3 * A.Wb;
#pragma unroll
            for (u32 i = 0; i < 4; i++) {
                if (i < A.I) {
                    const u32 h = T->prof[j][i] + stj - i;
                    const u32 kmer = kmer_at(h);
                    const uint2 r01 = __ldg((const uint2 *)(rec + 8 * (size_t)kmer)); kpm[i] = r01.x; knf[i] = r01.y; kk[i] = kmer;
                }
            }
#pragma unroll
            for (u32 i = 0; i < 4; i++) { const u32 pm = kpm[i]; if (i < A.I && pm != 0 && pm <= A.di.maxk) { tot += pm; nne++; } }
            for (u32 i = 4; i < A.I; i++) {
                const u32 h = T->prof[j][i] + stj - i;
                const u32 kmer = kmer_at(h);
                const u32 pm = __ldg(rec + 8 * (size_t)kmer);
                if (pm != 0 && pm <= A.di.maxk) { tot += pm; nne++; }
            }
        }
        // ---- block-wide exclusive scan of (items, candidates)
        u32 xi = nne, xc = tot;
        for (u32 o = 1; o < 32; o <<= 1) { u32 a = __shfl_up_sync(0xffffffffu, xi, o), b = __shfl_up_sync(0xffffffffu, xc, o); if (lane >= o) { xi += a; xc += b; } }
        if (lane == 31) { s_wi[wid] = xi; s_wc[wid] = xc; }
        __syncthreads();
        u32 wi0 = 0, wc0 = 0, bi = 0, bc = 0;
        for (u32 w = 0; w < LK_THREADS / 32; w++) { if (w < wid) { wi0 += s_wi[w]; wc0 += s_wc[w]; } bi += s_wi[w]; bc += s_wc[w]; }
        if (threadIdx.x == 0) {
            u32 ok = 1; unsigned long long old = 0;
            if (bi) {
                // one atomicAdd per block (a CAS loop serialises: one winner per round trip). A block whose range does not
                // fit records its start in rc->limit_inv (as the complement, so that 0 = none): every earlier range fits, every later one fails too, so
                // [0, limit) is exactly the part of the flat space that exists. Sticky check keeps the packed counter from wrapping.
                if (*(volatile unsigned long long *)&rc->limit_inv != 0ULL) ok = 0;
                else {
                    old = atomicAdd(&rc->alloc, ((unsigned long long)bi << ALLOC_SHIFT) | bc);
                    if ((old & ALLOC_MASK) + bc > A.cap_cands || (old >> ALLOC_SHIFT) + bi > A.cap_items) { ok = 0; atomicMax(&rc->limit_inv, ~old); }
                }
            }
            s_ok = ok; s_ibase = (u32)(old >> ALLOC_SHIFT); s_cbase = (u32)(old & ALLOC_MASK);
        }
        __syncthreads();
        const bool ok = s_ok != 0;
        u32 it = s_ibase + wi0 + xi - nne, cb = s_cbase + wc0 + xc - tot;
        // ---- per-slot bookkeeping (the two chain threads of a slot are neighbouring lanes)
        const u32 tot_o = __shfl_xor_sync(0xffffffffu, tot, 1), se_o = __shfl_xor_sync(0xffffffffu, (u32)search, 1);
        if (act && c == 0) A.slot_flag[slot] = 0;                                       // reduce_round visits every listed slot and skips the unflagged ones
        if (c == 0 && (search || se_o)) {
            if (ok) {
                uint2 ss = A.stat[slot]; ss.x += A.I * ((u32)search + se_o); ss.y += tot + tot_o; A.stat[slot] = ss;
            } else { A.meta[slot].flags = m.flags | SF_OVERFLOW; atomicAdd(&A.ctr->overflow_n, 1u); }
        }
        // ---- pass 2: item headers, and the bucket walk itself: the entries of every non-empty bucket are copied, in
        //      visiting order (cyclic from rot, align.cpp:293-296), to flat_loc[base..base+m), so that verification
        //      reads its loc entries as a plain stream. A warp expands the buckets of its lanes together.
        if (search && ok) A.slot_item[(u64)slot * 2 + c] = make_uint2(it, nne);
        for (u32 i = 0; i < A.I; i++) {
            u32 x_cb = 0, x_pm = 0, x_e0 = 0, x_rot = 0;
            if (search && ok && nne) {
                const u32 h = T->prof[j][i] + stj - i;
                u32 pm, nf, kmer;
                if (i < 4) { pm = i == 0 ? kpm[0] : i == 1 ? kpm[1] : i == 2 ? kpm[2] : kpm[3]; nf = i == 0 ? knf[0] : i == 1 ? knf[1] : i == 2 ? knf[2] : knf[3]; kmer = i == 0 ? kk[0] : i == 1 ? kk[1] : i == 2 ? kk[2] : kk[3]; }
                else { kmer = kmer_at(h); const uint2 r01 = __ldg((const uint2 *)(rec + 8 * (size_t)kmer)); pm = r01.x; nf = r01.y; }
                if (pm != 0 && pm <= A.di.maxk) {
                    // where the entries are, as an index into loc[]: inside the record, or what the record's third word says
                    const u32 e0 = pm <= BSL_REC_INLINE ? A.di.rec_base + 8u * kmer + 2u : __ldg(rec + 8 * (size_t)kmer + 2);
                    // walk positions on the reverse strand: [x1, x2) (inv = 0) or all but [x1, x2) (inv = 1), see ItemHdr
                    const u32 rot = m.rnd % pm;
                    const u32 inv = rot < nf ? 0u : 1u, x1 = inv ? pm - rot : nf - rot, x2 = inv ? pm - rot + nf : pm - rot;
                    *(uint4 *)(A.hdr + it) = make_uint4(cb, slot | (c << 22) | ((u32)m.thr << 23) | (inv << 27) | (i << 28), x1 | ((u32)m.len << 23), x2 | (h << 23));
                    for (u32 cc = (cb + 31u) / 32u; (u64)cc * 32u < (u64)cb + pm; cc++) A.chunk_first[cc] = it;
                    x_cb = cb; x_pm = pm; x_e0 = e0; x_rot = rot;
                    cb += pm; it++;
                }
            }
            // ---- a large bucket (reads from repeat families) is copied by the whole warp, four loads in flight per lane
            u32 bigm = __ballot_sync(0xffffffffu, x_pm > LK_BIG);
            while (bigm) {
                const u32 l = __ffs(bigm) - 1; bigm &= bigm - 1;
                const u32 ycb = __shfl_sync(0xffffffffu, x_cb, l), ypm = __shfl_sync(0xffffffffu, x_pm, l);
                const u32 ye0 = __shfl_sync(0xffffffffu, x_e0, l), yrot = __shfl_sync(0xffffffffu, x_rot, l);
                for (u32 tt = lane; tt < ypm; tt += 128) {
                    u32 vv[4];
#pragma unroll
                    for (u32 q = 0; q < 4; q++) { const u32 t2 = tt + 32 * q; u32 e = yrot + t2; if (e >= ypm) e -= ypm; vv[q] = t2 < ypm ? __ldg(A.di.loc + ye0 + e) : 0u; }
#pragma unroll
                    for (u32 q = 0; q < 4; q++) { const u32 t2 = tt + 32 * q; if (t2 < ypm) A.flat_loc[ycb + t2] = vv[q]; }
                }
            }
            // ---- the others (two or three entries on average): the entries of all lanes' buckets in one flat sequence, a lane per
            //      entry — the owner of entry t is the first lane whose inclusive entry count exceeds t
            const u32 fp = x_pm > LK_BIG ? 0u : x_pm;
            u32 inc = fp;
            for (u32 o = 1; o < 32; o <<= 1) { const u32 a = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += a; }
            const u32 T_ = __shfl_sync(0xffffffffu, inc, 31);
            for (u32 t0 = 0; t0 < T_; t0 += 64) {
                u32 vv[2], dd[2];
#pragma unroll
                for (u32 q = 0; q < 2; q++) {
                    const u32 t = t0 + 32 * q + lane;
                    u32 ow = 0;
#pragma unroll
                    for (u32 o = 16; o; o >>= 1) { const u32 a = __shfl_sync(0xffffffffu, inc, ow + o - 1u); if (a <= t) ow += o; }
                    const u32 ycb = __shfl_sync(0xffffffffu, x_cb, ow), ypm = __shfl_sync(0xffffffffu, fp, ow), yinc = __shfl_sync(0xffffffffu, inc, ow);
                    const u32 ye0 = __shfl_sync(0xffffffffu, x_e0, ow), yrot = __shfl_sync(0xffffffffu, x_rot, ow);
                    dd[q] = 0xffffffffu; vv[q] = 0;
                    if (t < T_) { const u32 idx = t - (yinc - ypm); u32 e = yrot + idx; if (e >= ypm) e -= ypm; vv[q] = __ldg(A.di.loc + ye0 + e); dd[q] = ycb + idx; }
                }
#pragma unroll
                for (u32 q = 0; q < 2; q++) if (dd[q] != 0xffffffffu) A.flat_loc[dd[q]] = vv[q];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// verify_candidates : CountMismatch / CountMismatch_new over the flat candidate space
// ------------------------------------------------------------------------------------------------
#define VF_THREADS 256
#define VF_ITMAX 128u        // items whose read streams are staged at a time (a chunk with more items is done in groups)
#define MK_CAP 4u            // marked candidates per slot and round that reduce_fast can take
#define VF_EAGER 96u         // item headers every CTA prefetches for its next chunk (a chunk with more items loads the rest on demand)

// One THREAD per candidate. The thread gathers the 32-byte sectors its reference window touches with 256-bit loads
// (2-3 sectors for 150 bp), rotates them in registers so that the window starts at register 0, and walks the read's
// words: reference half-words come out of the registers with one funnel shift each, read half-words come out of the
// staged streams in shared memory as aligned 64-bit loads. NS = sectors a window can touch (3 up to 256 bp, 5 up to 480).
//
// Staged read streams: every item owns NPL streams of 2*Wb 32-bit words in LOGICAL order (word j = bases 16j..16j+15,
// first base in the top bits), zero beyond the read:
//   stream 0: read bases (2-bit codes)
//   stream 1: N-mask reduced to one bit per base (01 = ACGT) — ANDed with the mismatch digits before the popcount
//   stream 2: convert-to mask (multi-way / '-' rules only)
//   last    : prefix mask 01 for read positions < h+s (only with -g: GapAlign's first test, align.cpp:353-360)

// mismatch digits (either bit of a digit set) of 16 bases: reference half-word r against read half-word q / convert mask cm
template <bool SINGLE>
__device__ __forceinline__ u32 vf_diff(u32 q, u32 cm, u32 r) {
    const u32 xc = ~(r << 1) | r | 0x55555555u;                          // XC64: digit 01 -> 01, else 11
    u32 d;
    if (SINGLE) d = (q & xc) ^ r;                                        // align.h:126-128
    else { const u32 m2 = xc | cm; const u32 m3 = m2 & (((m2 & 0xAAAAAAAAu) >> 1) | ((m2 & 0x55555555u) << 1)); d = ((~m3 & m2) | (m3 & q)) ^ r; }   // align.h:210-236
    return d | (d >> 1);                                                 // caller ANDs with a 01-per-base mask
}

__device__ __forceinline__ void ldg256(const u64 *p, u32 (&r)[8]) {     // one 32-byte sector -> 8 logical 32-bit words (first base in r[0]'s top bits)
    u64 a, b, c, d;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    r[0] = (u32)(a >> 32); r[1] = (u32)a; r[2] = (u32)(b >> 32); r[3] = (u32)b; r[4] = (u32)(c >> 32); r[5] = (u32)c; r[6] = (u32)(d >> 32); r[7] = (u32)d;
}

// the 32-byte sectors of the reference window of a candidate at global coordinate g (read length L) -> R, 8 logical words
// per sector; kk = logical word of the first sector the window starts in
template <int NS>
__device__ __forceinline__ void vf_gather(const u64 *plane, u32 g, u32 L, u32 (&R)[8 * NS], u32 &kk) {
    const u32 word0 = g >> 5, sb = word0 & ~3u;                          // first 64-bit word of the window, its 32-byte sector
    const u32 nwc = ((g & 31u) + L + 31u) >> 5;                          // 64-bit words the window touches
    const u32 nsec = ((word0 & 3u) + nwc + 3u) >> 2;
    kk = 2 * (word0 & 3u) + ((g & 31u) >> 4);
    const u64 *P = plane + sb;
#pragma unroll
    for (int sct = 0; sct < NS; sct++) if ((u32)sct < nsec) { u32 r8[8]; ldg256(P + 4 * sct, r8);
#pragma unroll
        for (int j = 0; j < 8; j++) R[8 * sct + j] = r8[j]; }
}

// CountMismatch / CountMismatch_new of a whole read against the gathered window (R is rotated in place so that the window
// starts at R[0]); S = the read's staged streams, sh = bit offset of the window start inside its half-word.
// pre (only with GAP) = mismatches under the prefix mask stream
template <bool SINGLE, bool GAP, int NS>
__device__ __forceinline__ void vf_count(u32 (&R)[8 * NS], u32 kk, u32 sh, const u32 *S, u32 W, u32 W2, u32 &snp, u32 &pre) {
    constexpr int NR = 8 * NS;
    constexpr u32 NP = SINGLE ? 2 : 3, PL_NM = 1, PL_CM = 2, PL_PM = NP;
    if (kk & 4u) {
#pragma unroll
        for (int j = 0; j + 4 < NR; j++) R[j] = R[j + 4];
    }
    if (kk & 2u) {
#pragma unroll
        for (int j = 0; j + 2 < NR; j++) R[j] = R[j + 2];
    }
    if (kk & 1u) {
#pragma unroll
        for (int j = 0; j + 1 < NR; j++) R[j] = R[j + 1];
    }
#pragma unroll
    for (int i = 0; i < (NR - 8) / 2; i++) {                                    // read word i = logical words 2i, 2i+1
        if ((u32)i < W) {
            const uint2 qq = *(const uint2 *)(S + 2 * i), nn = *(const uint2 *)(S + PL_NM * W2 + 2 * i);
            uint2 cc = make_uint2(0u, 0u); if (!SINGLE) cc = *(const uint2 *)(S + PL_CM * W2 + 2 * i);
            const u32 r0 = __funnelshift_l(R[2 * i + 1], R[2 * i], sh), r1 = __funnelshift_l(R[2 * i + 2], R[2 * i + 1], sh);
            const u32 d0 = vf_diff<SINGLE>(qq.x, cc.x, r0), d1 = vf_diff<SINGLE>(qq.y, cc.y, r1);
            snp += __popc((d0 & nn.x) | ((d1 & nn.y) << 1));
            if (GAP) { const uint2 pp = *(const uint2 *)(S + PL_PM * W2 + 2 * i); pre += __popc((d0 & pp.x) | ((d1 & pp.y) << 1)); }
        }
    }
}

// Used with -g only (the screens below handle -g 0): GapAlign's first test needs the whole window of every candidate.
template <bool SINGLE, int NS>
__global__ void __launch_bounds__(VF_THREADS, NS == 3 ? 4 : 2) verify_candidates(const __grid_constant__ KArgs A, u32 ci, u32 W) {
    constexpr bool GAP = true;
    extern __shared__ u32 vsm[];                          // staged streams: item x stream x 2*Wb
    __shared__ uint4 s_ha[CHUNK];                         // item headers (ItemHdr)
    __shared__ u32 s_mask[CHUNK / 32], s_nmk;
    __shared__ uint4 s_mk[CHUNK];                         // marked candidates of the chunk: {flat index, g, snp | strand << 8 | chain << 9, slot}
    constexpr u32 NP = SINGLE ? 2 : 3;                    // streams copied from global memory: bases, N-mask, (convert-to mask)
    constexpr u32 NPL = NP + (GAP ? 1 : 0);               // + prefix mask
    constexpr u32 PL_CM = 2, PL_PM = NP;
    constexpr int NR = 8 * NS;                            // logical 32-bit words of the gathered sectors
    RoundCtr *rc = A.ctr->rc + ci;
    const unsigned long long al = min(rc->alloc, ~rc->limit_inv);
    const u32 n_cands = (u32)(al & ALLOC_MASK), n_items = (u32)(al >> ALLOC_SHIFT);
    const u32 n_chunks = (n_cands + CHUNK - 1) / CHUNK;
    const u32 t = threadIdx.x, lane = t & 31u, wid = t >> 5;
    const u32 W2 = 2 * A.Wb, D = NP * W2, IST = NPL * W2;
    if (blockIdx.x >= n_chunks) return;
    // ---- software pipeline over this CTA's chunks: the headers and loc entries of the next chunk are loaded while this one is verified
    u32 chunk = blockIdx.x, first = A.chunk_first[chunk * (CHUNK / 32)];
    uint4 ha = make_uint4(0, 0, 0, 0); bool have = false;
    if (t < VF_EAGER && first + t < n_items) { ha = __ldg((const uint4 *)(A.hdr + first + t)); have = true; }
    u32 nchunk = chunk + gridDim.x, nfirst = nchunk < n_chunks ? A.chunk_first[nchunk * (CHUNK / 32)] : 0u;
    u32 cloc = chunk * CHUNK + t < n_cands ? __ldg(A.flat_loc + chunk * CHUNK + t) : 0u;      // seed-table entry of my candidate (flat_loc is a plain stream)
    auto stream_of = [&](u32 y) -> u32 { return (IH_SLOT(y) * 2 + IH_CHAIN(y)) * 3 * A.Wb; };   // first 64-bit word of an item's read streams
    for (; chunk < n_chunks;) {
        const u32 cbeg = chunk * CHUNK, cend = min(cbeg + CHUNK, n_cands);
        bool mine = have && (t == 0 || ha.x < cend);
        if (t < CHUNK / 32) s_mask[t] = 0;
        if (t == 0) s_nmk = 0;
        u32 n_it = (u32)__syncthreads_count(mine);
        if (n_it == VF_EAGER) {                                                  // (rare) more items than were prefetched
            if (t >= VF_EAGER && first + t < n_items) { ha = __ldg((const uint4 *)(A.hdr + first + t)); mine = ha.x < cend; }
            n_it = (u32)__syncthreads_count(mine);
        }
        if (mine) {
            s_ha[t] = ha;
            const u32 pos = ha.x > cbeg ? ha.x - cbeg : 0u;                  // first candidate of the item inside the chunk
            atomicOr(&s_mask[pos >> 5], 1u << (pos & 31u));
        }
        // prefetch for the next chunk (consumed at the top of the next iteration)
        uint4 pa = make_uint4(0, 0, 0, 0); bool phave = false;
        if (nchunk < n_chunks && t < VF_EAGER && nfirst + t < n_items) { pa = __ldg((const uint4 *)(A.hdr + nfirst + t)); phave = true; }
        const u32 nnchunk = nchunk + gridDim.x; const u32 nnfirst = nnchunk < n_chunks ? __ldg(A.chunk_first + (size_t)nnchunk * (CHUNK / 32)) : 0u;
        const u32 nloc = (nchunk < n_chunks && nchunk * CHUNK + t < n_cands) ? __ldg(A.flat_loc + nchunk * CHUNK + t) : 0u;
        __syncthreads();
        // ---- my candidate: item = (number of item starts at chunk positions <= t) - 1
        const u32 idx = cbeg + t;
        u32 it = 0;
        {
            u32 acc = 0;
#pragma unroll
            for (u32 ww = 0; ww < CHUNK / 32; ww++) { const u32 mk = s_mask[ww]; if (ww < wid) acc += __popc(mk); else if (ww == wid) acc += __popc(mk & (0xffffffffu >> (31u - lane))); }
            it = acc - 1u;
        }
        const bool act = idx < cend;
        bool marked = false;
        u32 g = 0, hy = 0, sig = 0, kk = 0, sh = 0;
        u32 R[NR];
#pragma unroll
        for (int j = 0; j < NR; j++) R[j] = 0;
        if (act) {
            const uint4 xa = s_ha[it];
            sig = ih_strand(idx - xa.x, xa.y, xa.z, xa.w);                       // forward-strand entries come first (align.cpp:296)
            hy = xa.y;
            g = cloc - IH_H(xa.w);                                               // _hit.loc (align.cpp:297)
            sh = (g & 15u) * 2;
            vf_gather<NS>(A.di.plane[sig], g, IH_L(xa.z), R, kk);
        }
        for (u32 grp = 0; grp < n_it; grp += VF_ITMAX) {
            const u32 n_g = min(VF_ITMAX, n_it - grp);
            if (grp) __syncthreads();                                        // the previous group is done with the staging buffer
            // ---- stage the streams of this group's items: a warp copies the D contiguous words of an item (coalesced),
            //      four items per warp in flight
            for (u32 l = lane; l < D; l += 32) {
                for (u32 m0 = wid; m0 < n_g; m0 += 4 * (VF_THREADS / 32)) {
                    u32 v[4];
#pragma unroll
                    for (u32 b = 0; b < 4; b++) { const u32 i2 = m0 + b * (VF_THREADS / 32); if (i2 < n_g) v[b] = __ldg((const u32 *)(A.planes + stream_of(s_ha[grp + i2].y)) + l); }
#pragma unroll
                    for (u32 b = 0; b < 4; b++) { const u32 i2 = m0 + b * (VF_THREADS / 32); if (i2 < n_g) vsm[(size_t)i2 * IST + l] = v[b]; }
                }
            }
            {
                for (u32 j = lane; j < W2; j += 32)
                    for (u32 i2 = wid; i2 < n_g; i2 += VF_THREADS / 32) {
                        const u32 hs = IH_H(s_ha[grp + i2].w) + A.s;
                        vsm[(size_t)i2 * IST + PL_PM * W2 + j] = hs >= 16 * j + 16 ? 0x55555555u : (hs <= 16 * j ? 0u : 0x55555555u & (0xffffffffu << (32 - 2 * (hs - 16 * j))));
                    }
            }
            // ---- L2 prefetch for the NEXT chunk (its headers, loaded at the top of this iteration, have arrived by now):
            //      the read streams of every item, so that the staging copies of the next iteration hit L2 instead of DRAM
            if (grp == 0 && phave) {
                const u32 ncend = min(nchunk * CHUNK + CHUNK, n_cands);
                if (t == 0 || pa.x < ncend) {
                    const char *pp = (const char *)(A.planes + stream_of(pa.y));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(pp));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(pp + 4 * D - 4));
                }
            }
            __syncthreads();                                                                // staged streams visible
            const bool now = act && it >= grp && it < grp + n_g;
            u32 snp = 0, pre = 0;
            if (now) vf_count<SINGLE, GAP, NS>(R, kk, sh, vsm + (size_t)(it - grp) * IST, W, W2, snp, pre);
            // could this candidate add a hit?  ungapped: snp <= thr.  gapped: GapAlign's first test (align.cpp:353-360)
            const u32 thr = IH_THR(hy);
            const bool mark = now && (snp <= thr || (thr >= 2 && pre < thr - 1));
            if (mark) { marked = true; s_mk[atomicAdd(&s_nmk, 1u)] = make_uint4(idx, g, snp | (sig << 8) | (IH_CHAIN(hy) << 9), IH_SLOT(hy)); }
        }
        // ---- 32 verdicts are one word of the bitmap; marked candidates are published for reduce_fast / reduce_round
        {
            const u32 bal = __ballot_sync(0xffffffffu, marked);
            if (lane == 0) A.bitmap[(cbeg >> 5) + wid] = bal;
        }
        __syncthreads();
        if (t < s_nmk) atomicAdd(&A.slot_flag[s_mk[t].w], 1u);          // reduce_round replays the flagged reads from the bitmap
        __syncthreads();
        chunk = nchunk; first = nfirst; ha = pa; have = phave; nchunk = nnchunk; nfirst = nnfirst; cloc = nloc;
    }
}

// ------------------------------------------------------------------------------------------------
// screen_candidates : candidate verification without -g, any conversion rule (2-bit planes).
//
// A WARP owns 32 consecutive flat candidates and never waits for another warp (no block barriers): lanes load the
// headers of the items overlapping the group, every lane resolves its candidate's item with one popcount, issues ONE
// 256-bit gather (the 32-byte sector of its reference window that covers most of the read), the warp stages the items'
// read streams into its private slice of shared memory behind those gathers, and each lane counts the mismatches of
// the read half-words the sector covers. That count is a lower bound of CountMismatch (align.h:118-131 / 199-239: a
// sum over half-words, the same alignment), so a candidate above its threshold is rejected for good. The rest (a few
// per cent) wait in a per-warp queue; whenever 32 are waiting they are counted exactly (drain_exact), one per lane.
// Single-conversion rules use screen_bits (below) instead: a pipelined version on a one-bit plane that stays in L2.
// ------------------------------------------------------------------------------------------------
#define SC_WARPS 8
#define SC_QCAP 64u          // survivors a warp can hold (it empties the queue whenever 32 are waiting)

// Exact count of one queued survivor per lane, sv = {flat index, g, ItemHdr::y with the STRAND in the inv bit, L}: the
// whole CountMismatch / CountMismatch_new (align.h:118-131 / 199-239) over the 2-bit window, reference words and read
// words straight from global memory (three independent loads per word, nothing staged). Passing candidates get their
// bitmap bit and a mark record for reduce_fast.
template <bool SINGLE, bool GAP = false>
__device__ __forceinline__ void drain_exact(const KArgs &A, const uint4 sv, bool have, u64 strm) {
    if (!have) return;
    const u32 y = sv.z, sig = IH_INV(y), slot = IH_SLOT(y), chain = IH_CHAIN(y), L = sv.w & 511u, g = sv.y;
    const u32 hs = GAP ? min((sv.w >> 9) + A.s, L) : 0u;                           // GAP: sv.w = L | h << 9; prefix [0, h + s) of GapAlign's first test
    const u32 Wb = A.Wb, W = (L + 31u) >> 5;
    const u64 *S = A.planes + ((u64)slot * 2 + chain) * 3 * Wb;                    // streams: bases, 01 per ACGT base, convert-to mask
    const u64 *P = A.di.plane[sig] + (g >> 5);
    const u32 off = (g & 31u) * 2;
    u32 snp = 0, pre = 0;
    const u32 thr = IH_THR(y);
    if (GAP && W <= 6u) {
        // -g, reads up to 192 bases: the window (one word before it, two behind) and the read words stay in registers, so that
        // gap_possible — a necessary condition of GapAlign finding anything, see reduce_round — is decided here and random
        // candidates that merely pass GapAlign's first test never reach reduce_round
        u64 win[9], q[6], cm[6];
#pragma unroll
        for (int k = 0; k < 9; k++) win[k] = (u32)k < W + 3u ? ldg_u64_stream(P - 1 + k, strm) : 0ULL;
        const u32 lastb = L & 31u; const u64 endmask = lastb ? (~0ULL << (64 - 2 * lastb)) : ~0ULL;
        // read word i against the window at shift sh (first base of the alignment = window base 32 + (g & 31) + sh)
        auto ref_at = [&](int sh, int i) -> u64 {
            const u32 pos = 32u + (g & 31u) + (u32)sh, w0 = pos >> 5, o = (pos & 31u) * 2;      // w0 in {0, 1, 2}
            const u64 a = w0 == 0 ? win[i] : (w0 == 1 ? win[i + 1] : win[i + 2]), b = w0 == 0 ? win[i + 1] : (w0 == 1 ? win[i + 2] : (i + 3 < 9 ? win[i + 3] : 0ULL));
            return o ? (a << o) | (b >> (64 - o)) : a;
        };
        u64 d0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            q[i] = cm[i] = d0[i] = 0;
            if ((u32)i < W) {
                q[i] = swap32(ldg_u64_stream(S + i, strm)); const u64 n = swap32(ldg_u64_stream(S + Wb + i, strm));
                if (!SINGLE) cm[i] = swap32(ldg_u64_stream(S + 2 * Wb + i, strm));
                u64 d = bsl_pairs(bsl_diff<SINGLE>(q[i], cm[i], ref_at(0, i)));
                snp += __popcll(d & n);
                if ((u32)i == W - 1) d &= endmask;
                d0[i] = d;
                const u32 np = hs > 32u * i ? min(hs - 32u * i, 32u) : 0u;
                if (np) pre += __popcll(d & (~0ULL << (64 - 2 * np)));
            }
        }
        bool mark = snp <= thr;
        if (!mark && thr >= 2 && pre < thr - 1) {                                  // GapAlign's first test passes (align.cpp:353-360): can it find a gap at all?
            const u32 G = A.gap, M = L >> 1, start = G + M - 1;
            u32 a = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) if (32u * i < M && (u32)i < W) { const u32 n = min(M - 32u * i, 32u); a += __popcll(d0[i] & (~0ULL << (64 - 2 * n))); }
            mark = a <= thr - 2;
            for (u32 tt = 1; tt <= 2 * G && !mark; tt++) {
                const u32 t = (tt + 1) >> 1; const int sh = (tt & 1) ? -(int)t : (int)t;
                if (thr < 1 + t) break;
                u32 b = 0;
#pragma unroll
                for (int i = 0; i < 6; i++) if ((u32)i >= (start >> 5) && (u32)i < W) {
                    u64 d = bsl_pairs(bsl_diff<SINGLE>(q[i], cm[i], ref_at(sh, i))); if ((u32)i == W - 1) d &= endmask;
                    const u32 n0 = start > 32u * i ? start - 32u * i : 0u;
                    b += __popcll(n0 ? d & (~0ULL >> (2 * n0)) : d);
                }
                mark = b <= thr - 2;
            }
        }
        if (mark) { atomicOr(&A.bitmap[sv.x >> 5], 1u << (sv.x & 31u)); atomicAdd(&A.slot_flag[slot], 1u); }
        return;
    }
    u64 prev = ldg_u64_stream(P, strm);
#pragma unroll 3
    for (u32 i = 0; i < W; i++) {
        const u64 next = ldg_u64_stream(P + i + 1, strm);
        const u64 r = off ? (prev << off) | (next >> (64 - off)) : prev;
        const u64 q = swap32(ldg_u64_stream(S + i, strm)), n = swap32(ldg_u64_stream(S + Wb + i, strm));
        u64 cm = 0; if (!SINGLE) cm = swap32(ldg_u64_stream(S + 2 * Wb + i, strm));
        const u64 d = bsl_pairs(bsl_diff<SINGLE>(q, cm, r));
        snp += __popcll(d & n);
        if (GAP) {                                                                 // mismatches among the first h + s bases, without the N mask (MismatchPattern0, align.h:133-168)
            const u32 np = hs > 32 * i ? min(hs - 32 * i, 32u) : 0u;
            if (np) pre += __popcll(d & (~0ULL << (64 - 2 * np)));
        }
        prev = next;
    }
    if (snp <= thr || (GAP && thr >= 2 && pre < thr - 1)) {                        // ungapped hit, or GapAlign's first test passes (align.cpp:353-360)
        atomicOr(&A.bitmap[sv.x >> 5], 1u << (sv.x & 31u));
        const u32 pos = atomicAdd(&A.slot_flag[slot], 1u);
        if (!GAP && pos < MK_CAP) A.marks[(size_t)slot * MK_CAP + pos] = make_uint4(sv.x, g, snp | (sig << 8) | (chain << 9), 0u);
    }
}
// queue entry of a survivor
__device__ __forceinline__ uint4 survivor(u32 flat, u32 g, u32 hy, u32 sig, u32 L) { return make_uint4(flat, g, (hy & ~(1u << 27)) | (sig << 27), L); }

template <bool SINGLE>
__global__ void __launch_bounds__(SC_WARPS * 32, 5) screen_candidates(const __grid_constant__ KArgs A, u32 ci, u32 stage_items, u32 rcp_dw) {
    extern __shared__ u32 ssm[];
    __shared__ uint4 s_q[SC_WARPS][SC_QCAP];          // survivors waiting for the exact count
    constexpr u32 NP = SINGLE ? 2 : 3, PL_NM = 1, PL_CM = 2, FULL = 0xffffffffu;
    const RoundCtr *rc = A.ctr->rc + ci;
    const unsigned long long al = min(rc->alloc, ~rc->limit_inv);
    const u32 n_cands = (u32)(al & ALLOC_MASK), n_items = (u32)(al >> ALLOC_SHIFT);
    const u32 n_groups = (n_cands + 31u) >> 5;
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 W2 = 2 * A.Wb, D = NP * W2, DW = NP * A.Wb;
    u32 *S0 = ssm + (size_t)wid * stage_items * D;
    u64 *S64 = (u64 *)S0;
    uint4 *Q = s_q[wid];
    u32 qn = 0;
    const u64 strm = l2_policy_stream();
    for (u32 grp = blockIdx.x * SC_WARPS + wid; grp < n_groups; grp += gridDim.x * SC_WARPS) {
        const u32 gbeg = grp << 5, gend = min(gbeg + 32u, n_cands);
        const u32 first = __ldg(A.chunk_first + grp);
        const u32 last = grp + 1 < n_groups ? __ldg(A.chunk_first + grp + 1) : n_items - 1u;
        const bool act = gbeg + lane < gend;
        const u32 cloc = act ? __ldg(A.flat_loc + gbeg + lane) : 0u;
        uint4 ha = make_uint4(FULL, 0, 0, 0);
        const bool ld = first + lane <= last;
        if (ld) ha = __ldg((const uint4 *)(A.hdr + first + lane));
        if (lane == 0) A.bitmap[grp] = 0u;                                       // bits are set by drain_exact
        const bool mine = ld && (lane == 0 || ha.x < gend);
        const u32 pos = (mine && ha.x > gbeg) ? ha.x - gbeg : 0u;
        const u32 mask = __reduce_or_sync(FULL, mine ? 1u << pos : 0u);
        const u32 n_it = __popc(mask);
        const u32 soff = (IH_SLOT(ha.y) * 2 + IH_CHAIN(ha.y)) * 3 * A.Wb;         // first 64-bit word of the item's read streams
        const u32 it = __popc(mask & (FULL >> (31u - lane))) - 1u;               // my candidate's item = lane `it`
        const u32 ibase = __shfl_sync(FULL, ha.x, it), hy = __shfl_sync(FULL, ha.y, it), hz = __shfl_sync(FULL, ha.z, it), hw = __shfl_sync(FULL, ha.w, it);
        const u32 sig = ih_strand(gbeg + lane - ibase, hy, hz, hw);              // forward-strand entries come first (align.cpp:296)
        const u32 g = cloc - IH_H(hw), sh = (g & 15u) * 2, L = IH_L(hz);         // _hit.loc (align.cpp:297)
        // ---- the sector: read half-word i lines up with reference half-words gh+i, gh+i+1; the sector that starts o
        //      half-words before gh covers i in [0, 6-o], the next one i in [8-o, 14-o]; take the better covered one
        const u32 gh = g >> 4, o = gh & 7u, nh = (L + 15u) >> 4;
        const u32 c0 = min(7u - o, nh), hi1 = min(14u - o, nh - 1u);
        const u32 c1 = hi1 + o >= 8u ? hi1 + o - 7u : 0u;
        const u32 k = c1 > c0 ? 1u : 0u;
        const u32 dlt = k ? 8u - o : 0u - o;                                     // read half-word of sector half-word x = x + dlt
        u32 R[8];
#pragma unroll
        for (int j = 0; j < 8; j++) R[j] = 0;
        if (act) ldg256(A.di.plane[sig] + ((gh - o) >> 1) + 4u * k, R);
        bool pass = false;
        for (u32 s0 = 0; s0 < n_it; s0 += stage_items) {
            const u32 nb = min(stage_items, n_it - s0), nw = nb * DW;
            // ---- stage the read streams of items s0 .. s0+nb-1 (DW contiguous 64-bit words each), two trips in flight
            for (u32 x0 = 0; x0 < nw; x0 += 64) {
                const u32 xa = x0 + lane, xb = xa + 32u;
                const u32 ia = min(__umulhi(xa, rcp_dw), nb - 1u), ib = min(__umulhi(xb, rcp_dw), nb - 1u);
                const u32 sa = __shfl_sync(FULL, soff, s0 + ia), sb = __shfl_sync(FULL, soff, s0 + ib);
                u64 va = 0, vb = 0;
                if (xa < nw) va = __ldg(A.planes + sa + (xa - ia * DW));
                if (xb < nw) vb = __ldg(A.planes + sb + (xb - ib * DW));
                if (xa < nw) S64[xa] = va;
                if (xb < nw) S64[xb] = vb;
            }
            __syncwarp();
            if (act && it >= s0 && it < s0 + nb) {
                const u32 *S = S0 + (size_t)(it - s0) * D;
                u32 snp = 0;
#pragma unroll
                for (int x = 0; x < 7; x++) {
                    const u32 i = (u32)x + dlt;
                    if (i < nh) {
                        const u32 r = __funnelshift_l(R[x + 1], R[x], sh);
                        u32 cc = 0; if (!SINGLE) cc = S[PL_CM * W2 + i];
                        snp += __popc(vf_diff<SINGLE>(S[i], cc, r) & S[PL_NM * W2 + i]);
                    }
                }
                pass = snp <= IH_THR(hy);
            }
            __syncwarp();
        }
        // ---- survivors join the warp's queue; 32 waiting survivors are counted exactly, one per lane
        const u32 bal = __ballot_sync(FULL, pass);
        if (bal) {
            if (pass) Q[qn + __popc(bal & ((1u << lane) - 1u))] = survivor(gbeg + lane, g, hy, sig, L);
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 32u) { qn -= 32u; drain_exact<SINGLE>(A, Q[qn + lane], true, strm); __syncwarp(); }
        }
    }
    if (qn) drain_exact<SINGLE>(A, Q[min(lane, qn - 1u)], lane < qn, strm);
}

// ------------------------------------------------------------------------------------------------
// screen_bits : the screen for single-conversion rules, on the ONE-BIT forward plane (DevIndex::bit1) — the roofline kernel.
//
// A conversion never changes the low bit of a base's code (from = 01, to = 11), so "low bits differ" implies a mismatch
// under CountMismatch (align.h:126-128) and the number of such positions over any part of the read is a lower bound of
// its result. A 32-byte sector of bit1 holds 256 bases: one gather covers at least half of the window, and the whole
// plane of a 500 Mb reference is 62.5 MB. Candidates on the reverse strand are screened on the same plane: rc position
// g+k of sequence i is the complement of forward position 2*anchor_i + rc_offset_i - 1 - (g+k), complementing flips the
// low bit or not (DevIndex::flip, folded into the read's reversed stream by prepare_reads), so the REVERSED read is
// compared with the forward sector. Sectors that hold a non-ACGT base, padding or margin (DevIndex::nflag) and windows
// that cross a sequence boundary are not screened on that strand: they go straight to the exact count.
//
// A warp owns groups of 32 consecutive flat candidates (lane per candidate) and never waits for another warp. Its
// groups move through a software pipeline, one stage per visit; a visit runs C, S, B2, B1, P, F in that order, each for
// a group one further ahead, so every global load is consumed at the code position that issued it one visit earlier:
//   F  chunk_first of the group four visits ahead            (which items overlap the group)
//   P  item headers + loc entries of the group three ahead   (16-byte ItemHdr per lane, 4-byte flat_loc per lane)
//   B1 resolve each lane's item (one popcount over the "an item starts here" mask, four shuffles), strand, window start;
//      reverse-strand lanes issue their strand-table look-up (ctab)
//   B2 mirror reverse-strand windows, pick the sector, issue the 256-bit gather (and the nflag word)
//   S  cp.async the items' 1-bit streams (16 Wb bytes per item: {low bits, ACGT mask} words forward, then reversed)
//      into the warp's shared slice, consecutive lanes copying consecutive 16-byte chunks — issued right after the
//      previous group's compare has left the slice
//   C  XOR / mask / popcount of <= 7 read words against funnel-shifted sector words; survivors join the warp's queue,
//      32 waiting survivors are counted exactly (drain_exact), one per lane
// ------------------------------------------------------------------------------------------------
#define SB_WARPS 8
__device__ __forceinline__ void cp_async16(u32 smem_addr, const void *gptr, u64 pol) { asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(smem_addr), "l"(gptr), "l"(pol)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldg256_keep(const u32 *p, u32 (&r)[8], u64 pol) {      // one 32-byte sector of the one-bit plane, L2 evict-last
    u64 a, b, c, d;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    r[0] = (u32)a; r[1] = (u32)(a >> 32); r[2] = (u32)b; r[3] = (u32)(b >> 32); r[4] = (u32)c; r[5] = (u32)(c >> 32); r[6] = (u32)d; r[7] = (u32)(d >> 32);
}

// SINGLE: one substitution base (CountMismatch) or an empty substitution set like "T:-" (CountMismatch_new compares exactly).
// GAP (-g > 0): a candidate also survives when the low-bit mismatches among its first h + s bases stay below thr - 1, a lower bound
// of what GapAlign's first test counts (align.cpp:353-360); the exact count then decides both tests.
template <bool SINGLE, bool GAP>
__global__ void __launch_bounds__(SB_WARPS * 32, GAP ? 2 : 4) screen_bits(const __grid_constant__ KArgs A, u32 ci, u32 item_bytes, u32 rcp_wb) {
    extern __shared__ __align__(16) unsigned char sbm[];                       // per warp: 32 staged items x item_bytes
    __shared__ uint4 s_q[SB_WARPS][SC_QCAP];
    constexpr u32 FULL = 0xffffffffu;
    const RoundCtr *rc = A.ctr->rc + ci;
    const unsigned long long al = min(rc->alloc, ~rc->limit_inv);
    const u32 n_cands = (u32)(al & ALLOC_MASK), n_items = (u32)(al >> ALLOC_SHIFT);
    const u32 n_groups = (n_cands + 31u) >> 5;
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 stride = gridDim.x * SB_WARPS, g0 = blockIdx.x * SB_WARPS + wid;
    if (g0 >= n_groups) return;
    const u32 Wb = A.Wb, RB = 16u * Wb;                                        // bytes of one (slot, chain) record of bits1
    unsigned char *S0 = sbm + (size_t)wid * 32u * item_bytes;
    const u32 S0a = (u32)__cvta_generic_to_shared(S0);
    const unsigned char *bits = (const unsigned char *)A.bits1;
    const u64 keep = l2_policy_keep(), strm = l2_policy_stream();
    uint4 *Q = s_q[wid];
    u32 qn = 0;

    // ---- pipeline registers (every stage consumes what the same stage position issued one visit earlier)
    u32 fP = 0, lP = 0;                                                        // F -> P   chunk_first pair
    uint4 hB = make_uint4(FULL, 0, 0, 0); u32 clB = 0;                         // P -> B1  item header of this lane, loc entry of this lane's candidate
    u32 g2 = 0, y2 = 0, m2 = 0, rec2 = FULL, h2 = 0; uint2 ce2 = make_uint2(0u, 0u);   // B1 -> B2 / S   (m: L | staged offset << 9 | act << 23; y: ItemHdr::y with the strand in the inv bit)
    u32 R[8], nfl = 0, cx = 0, gC = 0, yC = 0, mC = 0;                         // B2 -> C  (cx: dlt+8 | sft << 5 | nW << 10 | screened << 14 | (sec & 31) << 15 | h << 20)
#pragma unroll
    for (int j = 0; j < 8; j++) R[j] = 0;

    // everything but the screening plane passes through once: evict-first, so that it does not push the plane out of L2
    auto stage_F = [&](u32 G) { fP = ldg_u32_stream(A.chunk_first + G, strm); lP = G + 1 < n_groups ? ldg_u32_stream(A.chunk_first + G + 1, strm) : n_items - 1u; };
    auto stage_P = [&](u32 G) {
        hB = make_uint4(FULL, 0, 0, 0); clB = 0;
        if (fP + lane <= lP) hB = ldg_v4_stream(A.hdr + fP + lane, strm);
        const u32 c = (G << 5) + lane;
        if (c < n_cands) clB = ldg_u32_stream(A.flat_loc + c, strm);
    };
    // B1: item of every candidate, strand, window start; reverse-strand lanes issue their strand-table look-up
    auto stage_B1 = [&](u32 G) {
        const u32 gbeg = G << 5, gend = min(gbeg + 32u, n_cands);
        const bool act = gbeg + lane < gend;
        const bool mine = hB.x != FULL && (lane == 0 || hB.x < gend);
        const u32 pos = (mine && hB.x > gbeg) ? hB.x - gbeg : 0u;
        const u32 mask = __reduce_or_sync(FULL, mine ? 1u << pos : 0u);
        const u32 it = __popc(mask & (FULL >> (31u - lane))) - 1u;               // my candidate's item = lane `it`
        const u32 ibase = __shfl_sync(FULL, hB.x, it), hy = __shfl_sync(FULL, hB.y, it), hz = __shfl_sync(FULL, hB.z, it), hw = __shfl_sync(FULL, hB.w, it);
        const u32 sig = ih_strand(gbeg + lane - ibase, hy, hz, hw);              // forward-strand entries come first (align.cpp:296)
        g2 = clB - IH_H(hw); h2 = IH_H(hw);                                      // _hit.loc (align.cpp:297)
        y2 = (hy & ~(1u << 27)) | (sig << 27);
        m2 = IH_L(hz) | ((it * item_bytes + (sig ? 8u * Wb : 0u)) << 9) | ((act ? 1u : 0u) << 23);
        ce2 = make_uint2(0u, 0u);
        if (act && sig) ce2 = __ldg(A.di.ctab + (g2 >> BSL_CTAB_SHIFT));
        rec2 = mine ? IH_SLOT(hB.y) * 2 + IH_CHAIN(hB.y) : FULL;
        if (lane == 0) A.bitmap[G] = 0u;                                         // bits are set by drain_exact
    };
    // S: the items' 1-bit streams into the warp's slice, 16-byte chunks, consecutive lanes take consecutive chunks
    auto stage_S = [&]() {
        const u32 n_it = __popc(__ballot_sync(FULL, rec2 != FULL)), nch = n_it * Wb;
        for (u32 c0 = 0; c0 < nch; c0 += 32) {
            const u32 c = c0 + lane, item = min(__umulhi(c, rcp_wb), n_it - 1u), part = c - item * Wb;
            const u32 rec = __shfl_sync(FULL, rec2, item);
            if (c < nch) cp_async16(S0a + item * item_bytes + 16u * part, bits + (size_t)rec * RB + 16u * part, strm);
        }
        cp_async_commit();
    };
    // B2: mirror reverse-strand windows, choose the sector, issue the gather
    auto stage_B2 = [&]() {
        const u32 L = m2 & 511u, sig = IH_INV(y2), g = g2; const bool act = (m2 >> 23) & 1u;
        u32 p0 = g; bool screen = act; uint2 ce = ce2;
        if (act && sig) {
            if (ce.y == 0u) {                                                    // a sequence boundary inside the 65 536-coordinate block
                u32 lo = 0, hi = A.di.nseq;
                while (lo + 1 < hi) { const u32 mid = (lo + hi) >> 1; if (g >= __ldg(A.di.anchor + mid)) lo = mid; else hi = mid; }
                const u32 a0 = __ldg(A.di.anchor + lo), P = __ldg(A.di.rcoff + lo);
                ce = make_uint2(2u * a0 + P - 1u, a0 + P);
                if (g < a0) screen = false;                                      // in the leading margin
            }
            if ((u64)g + L > (u64)ce.y) screen = false;                          // crosses into the next sequence
            p0 = ce.x - g - L + 1u;                                              // forward coordinate of the window's lowest base
        }
        // the sector: read word i (32 bases; of the reversed read on the rc strand) lines up with plane words pw+i, pw+i+1;
        // the sector that starts o words before pw covers i in [0, 6-o], the next one [8-o, 14-o]
        const u32 pw = p0 >> 5, o = pw & 7u, nW = (L + 31u) >> 5, sft = p0 & 31u;
        const u32 c0 = min(7u - o, nW), hi1 = min(14u - o, nW - 1u);
        const u32 c1 = hi1 + o >= 8u ? hi1 + o - 7u : 0u;
        const u32 k = c1 > c0 ? 1u : 0u;
        const u32 dlt8 = k ? 16u - o : 8u - o;                                   // (read word of sector word x) - x + 8
        const u32 sec = (pw >> 3) + k;
        nfl = 0;
        if (screen) {
            ldg256_keep(A.di.bit1 + (size_t)sec * 8, R, keep);
            if (sig) nfl = __ldg(A.di.nflag + (sec >> 5));
        }
        cx = dlt8 | (sft << 5) | (nW << 10) | ((screen ? 1u : 0u) << 14) | ((sec & 31u) << 15) | (h2 << 20);
        gC = g2; yC = y2; mC = m2;
    };
    // C: XOR / mask / popcount against the staged streams; survivors join the queue
    auto stage_C = [&](u32 G) {
        cp_async_wait_all();
        __syncwarp();
        const bool act = (mC >> 23) & 1u; const u32 sig = IH_INV(yC);
        bool pass = false;
        if (act) {
            if (!((cx >> 14) & 1u) || (sig && ((nfl >> ((cx >> 15) & 31u)) & 1u))) pass = true;     // not screened on this strand: the exact count decides
            else {
                const uint2 *Sr = (const uint2 *)(S0 + ((mC >> 9) & 0x3FFFu));       // {low bits, ACGT mask} per 32 bases, forward or reversed
                const u32 sft = (cx >> 5) & 31u, nW = (cx >> 10) & 15u, d8 = cx & 31u;
                u32 low = 0, lowp = 0;
                // GAP: the first h + s bases of the read; in the reversed stream (reverse strand) they are the LAST h + s positions
                const u32 L = mC & 511u, hs = GAP ? min((cx >> 20) + A.s, L) : 0u;
#pragma unroll
                for (int x = 0; x < 7; x++) {
                    const u32 i = (u32)x + d8 - 8u;
                    if (i < nW) {
                        const uint2 lm = Sr[i]; const u32 dm = (lm.x ^ __funnelshift_l(R[x + 1], R[x], sft)) & lm.y;
                        low += __popc(dm);
                        if (GAP) {
                            u32 pm;
                            if (!sig) { const u32 np = hs > 32u * i ? min(hs - 32u * i, 32u) : 0u; pm = np ? (0xffffffffu << (32u - np)) : 0u; }
                            else { const u32 z0 = L - hs, sk = z0 > 32u * i ? min(z0 - 32u * i, 32u) : 0u; pm = sk >= 32u ? 0u : (0xffffffffu >> sk); }
                            lowp += __popc(dm & pm);
                        }
                    }
                }
                const u32 thr = IH_THR(yC);
                pass = low <= thr || (GAP && thr >= 2 && lowp < thr - 1);
            }
        }
        const u32 bal = __ballot_sync(FULL, pass);
        if (bal) {
            if (pass) {
                Q[qn + __popc(bal & ((1u << lane) - 1u))] = make_uint4((G << 5) + lane, gC, yC, (mC & 511u) | (GAP ? (cx >> 20) << 9 : 0u));
            }
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 32u) { qn -= 32u; drain_exact<SINGLE, GAP>(A, Q[qn + lane], true, strm); }
        }
        __syncwarp();                                                            // every lane is done with the slice
    };

    // ---- fill the pipeline: group k of this warp is G_k = g0 + k * stride
    const u32 G1 = g0 + stride, G2 = G1 + stride, G3 = G2 + stride;
    stage_F(g0);
    stage_P(g0); if (G1 < n_groups) stage_F(G1);
    stage_B1(g0); if (G1 < n_groups) { stage_P(G1); if (G2 < n_groups) stage_F(G2); }
    stage_S(); stage_B2(); if (G1 < n_groups) { stage_B1(G1); if (G2 < n_groups) { stage_P(G2); if (G3 < n_groups) stage_F(G3); } }
    // ---- steady state: C(G), then S / B2 of the next group, B1 of the one after, P, F
    for (u32 G = g0; G < n_groups; G += stride) {
        stage_C(G);
        const u32 Ga = G + stride, Gb = Ga + stride, Gc = Gb + stride, Gd = Gc + stride;
        if (Ga >= n_groups) break;
        stage_S(); stage_B2();
        if (Gb < n_groups) { stage_B1(Gb); if (Gc < n_groups) { stage_P(Gc); if (Gd < n_groups) stage_F(Gd); } }
    }
    if (qn) drain_exact<SINGLE, GAP>(A, Q[min(lane, qn - 1u)], lane < qn, strm);
}

// ------------------------------------------------------------------------------------------------
// reduce_fast : AddHit replay for the common case, one THREAD per read. A read whose marked candidates of this round
// (at most MK_CAP, recorded by verify_candidates with their mismatch counts) cannot trigger the -w feedback and fit its
// hit list needs no window gather and no warp: sort the marks into discovery order (= flat index), int2hit + AddHit
// each (align.cpp:319-346, align.h:329-347). Everything else (long lists, -w in reach, -g) is left to reduce_round.
// ------------------------------------------------------------------------------------------------
// Reads it leaves to reduce_round are collected in KArgs::flag_list: those with many marked candidates or a long hit list (reads from
// repeat families: their replay is long and serial) from the front, the others from the back, so that reduce_round can start the
// long ones first and hand every warp one read at a time. fast = 0 (-g): nothing is replayed here, the reads are only listed.
#define RR_LONG_MARKS 48u
#define RR_LONG_HITS 64u
__global__ void __launch_bounds__(256) reduce_fast(const __grid_constant__ KArgs A, const u32 *list, u32 list_ci, u32 ci, u32 as_pe, u32 fast) {
    const u32 n_entries = A.ctr->rc[list_ci].active << (as_pe ? 1 : 0);
    RoundCtr *rc = A.ctr->rc + ci;
    const u32 lane = threadIdx.x & 31u;
    u32 added = 0;
    for (u32 kb = blockIdx.x * blockDim.x + threadIdx.x - lane; kb < n_entries; kb += gridDim.x * blockDim.x) {
      const u32 k = kb + lane;
      u32 cls = 0, slot = 0;                                                // 0: nothing left to do, 1: reduce_round, 2: reduce_round, long replay
      if (k < n_entries) do {
        slot = as_pe ? list[k >> 1] + (k & 1u) * A.n_a : list[k];
        const u32 c = A.slot_flag[slot];
        if (c == 0) break;
        SlotMeta m = A.meta[slot];
        cls = (c >= RR_LONG_MARKS || m.nhit >= RR_LONG_HITS) ? 2u : 1u;
        if (!fast || c > MK_CAP) break;
        u32 hcap; DevHit *hits = slot_hits(A, m.item, hcap);
        if (m.nhit + c >= A.w || m.nhit + c > hcap) break;                // a level could reach -w, or the list could outgrow its storage
        cls = 0;
        uint4 e[MK_CAP];
#pragma unroll
        for (u32 i = 0; i < MK_CAP; i++) e[i] = i < c ? A.marks[(size_t)slot * MK_CAP + i] : make_uint4(0xffffffffu, 0, 0, 0);
#pragma unroll
        for (u32 i = 1; i < MK_CAP; i++)                                    // discovery order = flat candidate index
#pragma unroll
            for (u32 j = i; j > 0; j--) if (e[j].x < e[j - 1].x) { const uint4 tmp = e[j]; e[j] = e[j - 1]; e[j - 1] = tmp; }
        u32 nhit = m.nhit, minl = A.minlvl[slot];
        const u32 L = m.len;
#pragma unroll
        for (u32 i = 0; i < MK_CAP; i++) {
            if (i >= c) break;
            const u32 g = e[i].y, snp = e[i].z & 0xffu, sig = (e[i].z >> 8) & 1u, chain = (e[i].z >> 9) & 1u;
            if (snp > m.thr) continue;
            u32 lo = 0, hi = A.di.nseq;
            while (lo + 1 < hi) { const u32 mid = (lo + hi) >> 1; if (g >= A.di.anchor[mid]) lo = mid; else hi = mid; }
            u32 x = g - A.di.anchor[lo], gp = 0;
            if (sig) { x = A.di.rcoff[lo] - L - x; gp = L & 511u; }
            if ((int)x < 0 || x + L > A.di.seqlen[lo]) continue;
            bool dup = false;
            for (u32 h = 0; h < nhit; h++) { const DevHit hh = hits[h]; if (hh.loc == x && HIT_GAPPED(hh.tag) == 0 && (HIT_CHR2(hh.tag) >> 1) == lo) { dup = true; break; } }
            if (dup) continue;
            DevHit hh; hh.loc = x; hh.tag = (lo * 2 + sig) | (snp << 20) | (chain << 24); hh.gap = 0; hh.gp = gp; hits[nhit++] = hh;
            ((u16 *)&A.cnt[slot])[chain * 16 + snp]++;
            minl = min(minl, snp);
        }
        added += nhit - m.nhit;
        if (nhit != m.nhit) { A.meta[slot].nhit = (u16)nhit; A.minlvl[slot] = (u8)minl; }
        A.slot_flag[slot] = 0;                                              // done
      } while (0);
      // ---- list what is left (one atomic per warp and list)
      const u32 bl = __ballot_sync(0xffffffffu, cls == 2u), bn = __ballot_sync(0xffffffffu, cls == 1u);
      if (bl | bn) {
        u32 basel = 0, basen = 0;
        if (lane == 0) { if (bl) basel = atomicAdd(&rc->long_n, (u32)__popc(bl)); if (bn) basen = atomicAdd(&rc->flagged, (u32)__popc(bn)); }
        basel = __shfl_sync(0xffffffffu, basel, 0); basen = __shfl_sync(0xffffffffu, basen, 0);
        const u32 below = (1u << lane) - 1u;
        if (cls == 2u) A.flag_list[basel + __popc(bl & below)] = slot;
        if (cls == 1u) A.flag_list[A.n_slots - 1u - (basen + __popc(bn & below))] = slot;
      }
    }
    for (u32 o = 16; o; o >>= 1) added += __shfl_xor_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31u) == 0 && added) atomicAdd(&A.ctr->hits_added, (unsigned long long)added);
}

// ------------------------------------------------------------------------------------------------
// reduce_round
// ------------------------------------------------------------------------------------------------
#define ROUND_WARPS 8
#define GAP_NONE 0xffffffffu

struct WarpCtx {
    // warp-uniform running state of one read (one SingleAlign object)
    u32 L, W, thr, nhit, chain, cap, item; u32 mycnt;     // mycnt: lane c*16+l holds hits[c][l]
    DevHit *hits; bool overflow;
    u64 *keys; u32 kcap;                            // shared-memory copy of the dedup keys of hits[0 .. min(nhit, kcap))
};
#define RR_KEYS 512u         // dedup keys a warp keeps in shared memory (longer lists scan the rest in global memory)
// what AddHit's std::set compares (align.h:334-339): forward coordinate, sequence, gapped or not
__device__ __forceinline__ u64 hit_key(u32 loc, u32 seq, u32 gapped) { return (u64)loc | ((u64)(seq | (gapped << 20)) << 32); }

// ref word aligned to read word i for an alignment starting at window-relative base `rel`
__device__ __forceinline__ u64 ref_word(const u64 *win, u32 NW, u32 rel, u32 i) {
    u32 pos = rel + 32 * i, wi = pos >> 5, o = (pos & 31u) * 2;
    u64 x = win[wi] << o;
    if (o && wi + 1 < NW) x |= win[wi + 1] >> (64 - o);
    return x;
}

// GapAlign (align.cpp:348-410) over MismatchPattern0/1 (align.h:133-196, 241-327): the result is GAP_NONE or
// level | (shift+4)<<8 | gap_pos<<16.
// ---- GapAlign for ONE candidate by the whole warp (a thread-serial search costs the same issue slots with one lane active).
// The two half-warps take two shifts at a time: lane k of a half holds the k-th mismatch position of its list
// (MismatchPattern0 from the left at shift 0, MismatchPattern1 from the right at the half's shift), found by a prefix sum over the
// per-word mismatch masks and __fns; the (i, j) search of align.cpp:370-405 becomes lane i counting the PR entries below what it
// needs. The reference's answer, including which of several admissible gaps wins (lowest tt, then i, then j).
__device__ __forceinline__ u32 compress_even(u64 x) {                      // bit k of the result = bit 2k of x
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL; x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL; x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffULL; x = (x | (x >> 16)) & 0x00000000ffffffffULL;
    return (u32)x;
}
// mismatches of read word w against the window at offset relx, one bit per base, first base in the top bit; no N mask (align.h:133-196)
template <bool SINGLE>
__device__ __forceinline__ u32 mm_word(const u64 *win, u32 NW, u32 relx, const u64 *q, const u64 *cm, u32 w, u32 W, u64 endmask) {
    if (w >= W) return 0u;
    u64 d = bsl_pairs(bsl_diff<SINGLE>(q[w], cm[w], ref_word(win, NW, relx, w))); if (w == W - 1) d &= endmask;
    return compress_even(d);
}
// half-warp: lane hl holds `bits` of slot hl (slots and the bits inside them in list order, lowest bit first); returns true and the
// (slot, bit) of the k-th set bit, k = hl
__device__ __forceinline__ bool kth_bit(u32 bits, u32 hl, u32 nslots, u32 &slot, u32 &bit) {
    const u32 cnt = __popc(bits); u32 incl = cnt;
    for (u32 o = 1; o < 16; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o, 16); if (hl >= o) incl += v; }
    bool found = false; u32 excl = 0; slot = 0;
    for (u32 w = 0; w < nslots; w++) {
        const u32 iw = __shfl_sync(0xffffffffu, incl, w, 16), cw_ = __shfl_sync(0xffffffffu, cnt, w, 16);
        if (!found && iw > hl) { found = true; slot = w; excl = iw - cw_; }
    }
    const u32 sb = __shfl_sync(0xffffffffu, bits, slot, 16);
    bit = found ? __fns(sb, 0u, (int)(hl - excl + 1u)) : 0u;
    return found;
}
template <bool SINGLE>
__device__ u32 gap_search_warp(const u64 *win, u32 NW, u32 rel, const u64 *q, const u64 *cm, u32 L, u32 W, u64 endmask,
                               u32 thr, u32 h, u32 s, u32 G, u32 lane) {
    if (thr < 2) return GAP_NONE;
    const u32 want = thr - 1, hl = lane & 15u, half = lane >> 4;
    // P0: the first `want` mismatch positions from the left at shift 0 (lane k: position k, or L)
    u32 slot, bit;
    const bool f0 = kth_bit(__brev(mm_word<SINGLE>(win, NW, rel, q, cm, hl, W, endmask)), hl, W, slot, bit);
    const u32 P0k = (f0 && hl < want) ? 32u * slot + bit : L;
    const u32 ret0 = __shfl_sync(0xffffffffu, P0k, want - 1u, 16);           // the last collected position, or L when fewer were found
    if (ret0 < h + s) return GAP_NONE;
    // a gap sits at one of the first thr - t mismatch positions P0[i] with 6 <= P0[i] < L - t - 1 (align.cpp:372-376): none there, nothing to find
    const u32 real = __ballot_sync(0xffffffffu, hl + 2u <= thr && P0k >= 6u && P0k + 2u < L) & 0xffffu;
    if (!real) return GAP_NONE;
    const u32 gpmax = __shfl_sync(0xffffffffu, P0k, 31u - (u32)__clz(real), 16);    // the largest of them (P0 ascends)
    for (u32 tt0 = 1; tt0 <= 2 * G; tt0 += 2) {
        const u32 tt = tt0 + half, t = (tt + 1) >> 1; const int sh = (tt & 1) ? -(int)t : (int)t, sh1 = sh < 0 ? sh : 0;
        const bool active = tt <= 2 * G && thr >= 1 + t;                     // `if (thr < 1 + t) break;`
        // PR: the first `want` mismatch positions from the right at this shift, counted from the read's last base
        const u32 wr = W - 1u - hl;                                          // words right to left (hl < W)
        const u32 c = (active && hl < W) ? mm_word<SINGLE>(win, NW, (u32)((int)rel + sh), q, cm, wr, W, endmask) : 0u;
        // the right part must reach back to the gap: PR[j] >= L + min(sh, 0) - P0[i] for some j < thr - t - i, i.e. at most thr - t - 1 mismatches
        // among the last L + min(sh, 0) - P0[i] bases; cheapest to satisfy with the largest P0[i]. Count them before building the list.
        {
            const int xs = (int)L + sh1 - (int)gpmax;                        // bases counted from the read's end
            u32 cntx = 0;
            if (xs > 0 && hl < W) {
                const int lo_base = (int)L - xs, wb = 32 * (int)wr;          // bases >= lo_base of word wr
                const int skip = lo_base - wb;                               // leading bases of the word to ignore
                cntx = skip <= 0 ? __popc(c) : (skip >= 32 ? 0u : __popc(c & (0xffffffffu >> skip)));
            }
            for (u32 o = 8; o; o >>= 1) cntx += __shfl_xor_sync(0xffffffffu, cntx, o, 16);
            const bool hopeless = !active || cntx + t + 1u > thr;            // needs j <= thr - t - 1 - i mismatches skipped, so count <= thr - t - 1
            const u32 hb = __ballot_sync(0xffffffffu, hopeless);
            if ((hb & 0xffffu) && (hb >> 16)) continue;                      // both shifts of this trip
        }
        const bool f1 = kth_bit(c, hl, W, slot, bit);                        // lowest bit of a word = its last base
        const u32 PRk = (f1 && hl < want) ? L - 1u - (32u * (W - 1u - slot) + 31u - bit) : L;
        const u32 rl = L - t - 1u, i = hl, gp = P0k;
        const bool vg = active && i < thr - t && gp >= 6u && gp < rl;
        const int nd = (int)L + sh1 - (int)gp; const u32 need = nd > 6 ? (u32)nd : 6u;
        u32 jj = 0;
        for (u32 j = 0; j < want; j++) jj += __shfl_sync(0xffffffffu, PRk, j, 16) < need ? 1u : 0u;      // PR ascends: the first admissible j
        const u32 pj = __shfl_sync(0xffffffffu, PRk, min(jj, 15u), 16);
        const bool ok = vg && jj < thr - t - i && jj < want && pj < rl;
        u32 g2 = gp; { const int clip = (int)gp + 6 - (int)L - sh1; if (clip > 0) g2 -= (u32)clip; }
        const u32 res = (i + jj + t) | ((u32)(sh + 4) << 8) | (g2 << 16);
        const u32 bal = __ballot_sync(0xffffffffu, ok);
        if (bal) { const u32 src = (bal & 0xffffu) ? (u32)__ffs(bal & 0xffffu) - 1u : 16u + (u32)__ffs(bal >> 16) - 1u; return __shfl_sync(0xffffffffu, res, src); }
    }
    return GAP_NONE;
}

// A necessary condition of GapAlign finding anything (align.cpp:348-410), cheap enough to run for every candidate.
// A gapped hit at shift d is a left part [0, gp) with i mismatches at shift 0 and a right part of m2 bases with j
// mismatches at shift d, i, j <= thr - 2 and gp + m2 >= L - |d|. Cut the read at M = L / 2: either gp >= M, then the first
// M bases hold at most thr - 2 mismatches at shift 0, or m2 >= L - G - M + 1 =: y, then the last y bases hold at most
// thr - 2 mismatches at that shift. Mismatches are counted like MismatchPattern0/1 do (no N mask, align.h:133-196).
template <bool SINGLE>
__device__ bool gap_possible(const u64 *win, u32 NW, u32 rel, const u64 *q, const u64 *cm, u32 L, u32 W, u64 endmask, u32 thr, u32 G) {
    if (thr < 2) return false;
    const u32 M = L >> 1, start = G + M - 1;                                   // the last y = L - start bases
    u32 a = 0;
    for (u32 i = 0; i < W && 32 * i < M; i++) {
        const u32 n = min(M - 32 * i, 32u);
        u64 d = bsl_pairs(bsl_diff<SINGLE>(q[i], cm[i], ref_word(win, NW, rel, i))); if (i == W - 1) d &= endmask;
        a += __popcll(d & (~0ULL << (64 - 2 * n)));
    }
    if (a <= thr - 2) return true;
    for (u32 tt = 1; tt <= 2 * G; tt++) {
        const u32 t = (tt + 1) >> 1; const int sh = (tt & 1) ? -(int)t : (int)t;
        if (thr < 1 + t) break;
        const u32 rel1 = (u32)((int)rel + sh);
        u32 b = 0;
        for (u32 i = start >> 5; i < W; i++) {
            const u32 n0 = start > 32 * i ? start - 32 * i : 0u;                // bases of this word before the part
            u64 d = bsl_pairs(bsl_diff<SINGLE>(q[i], cm[i], ref_word(win, NW, rel1, i))); if (i == W - 1) d &= endmask;
            b += __popcll(n0 ? d & (~0ULL >> (2 * n0)) : d);
        }
        if (b <= thr - 2) return true;
    }
    return false;
}

// AddHit + int2hit (align.h:329-347, align.cpp:319-346); all arguments warp-uniform.
// returns 0 = continue, 1 = abort this SnpAlign call (level-0 list full, or storage overflow)
__device__ __forceinline__ u32 seq_of(const KArgs &A, u32 g) {               // binary search of ref_anchor (align.cpp:320-327)
    u32 lo = 0, hi = A.di.nseq;
    while (lo + 1 < hi) { u32 mid = (lo + hi) >> 1; if (g >= A.di.anchor[mid]) lo = mid; else hi = mid; }
    return lo;
}
__device__ int add_hit(const KArgs &A, WarpCtx &S, u32 lane, u32 level, u32 g, u32 lo, u32 sig, int sh, u32 gp) {
    u32 x = g - A.di.anchor[lo];
    if (sig) { x = A.di.rcoff[lo] - S.L - x; gp = (u32)((int)S.L + (sh < 0 ? sh : 0) - (int)gp); x -= (u32)sh; }
    gp &= 511u;
    if ((int)x < 0) return 0;
    if (x + S.L > A.di.seqlen[lo]) return 0;
    const u32 gapped = sh != 0;
    const u64 key = hit_key(x, lo, gapped);
    bool dup = false;
    const u32 nk = min(S.nhit, S.kcap);
    for (u32 i = lane; i < nk; i += 32) dup |= S.keys[i] == key;
    for (u32 i = nk + lane; i < S.nhit; i += 32) { DevHit hh = S.hits[i]; if (hh.loc == x && HIT_GAPPED(hh.tag) == gapped && (HIT_CHR2(hh.tag) >> 1) == lo) dup = true; }
    if (__any_sync(0xffffffffu, dup)) return 0;
    if (S.nhit >= S.cap) {
        // the list outgrew its storage: move it into a large block (no re-run); only when none is left does the read take the large-capacity pass
        u32 blk = 0xffffffffu;
        if (!(S.item & SLOT_BIG) && A.bighits && A.big_cap > S.cap) { if (lane == 0) blk = atomicAdd(&A.ctr->big_n, 1u); blk = __shfl_sync(0xffffffffu, blk, 0); }
        if (blk >= A.big_blocks) { S.overflow = true; return 1; }
        DevHit *nh = A.bighits + (u64)blk * A.big_cap;
        for (u32 i = lane; i < S.nhit; i += 32) nh[i] = S.hits[i];
        __syncwarp();
        S.hits = nh; S.cap = A.big_cap; S.item = SLOT_BIG | blk;
    }
    if (lane == 0) {
        DevHit hh; hh.loc = x; hh.tag = (lo * 2 + sig) | (level << 20) | (S.chain << 24) | (gapped << 25); hh.gap = (u32)sh; hh.gp = gp; S.hits[S.nhit] = hh;
        if (S.nhit < S.kcap) S.keys[S.nhit] = key;
    }
    S.nhit++;
    __syncwarp();
    if (lane == S.chain * 16 + level) S.mycnt++;
    u32 tot = __shfl_sync(0xffffffffu, S.mycnt, level) + __shfl_sync(0xffffffffu, S.mycnt, 16 + level);
    if (tot >= A.w) { if (level == 0) return 1; S.thr = level - 1; }
    return 0;
}

// The reads reduce_fast listed in KArgs::flag_list: first the long replays (front of the list), one per grab, so that they start
// early and spread over all warps; then the others (back of the list), RR_GRAB per grab.
#define RR_GRAB 2u
template <bool SINGLE>
__global__ void __launch_bounds__(ROUND_WARPS * 32, 3) reduce_round(const __grid_constant__ KArgs A, u32 round, u32 ci, u32 NW, u32 NWS) {
    extern __shared__ u64 smem[];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u64 *win_all = smem + (size_t)wid * (32 * NWS + 48 + RR_KEYS);
    u64 *pq = win_all + 32 * NWS, *pn = pq + 16, *pc = pn + 16, *keys = pc + 16;
    RoundCtr *rc = A.ctr->rc + ci;
    const unsigned long long al_rc = min(rc->alloc, ~rc->limit_inv);
    const u32 n_cands_rc = (u32)(al_rc & ALLOC_MASK), n_items_rc = (u32)(al_rc >> ALLOC_SHIFT);
    const u32 n_long = rc->long_n, n_norm = rc->flagged;
    const u32 G = A.gap;
    unsigned long long st_hits = 0;
    for (;;) {
        u32 v = 0;
        if (lane == 0) v = atomicAdd(&rc->work, 1u);                       // grab number: n_long grabs of one long replay, then the others
        v = __shfl_sync(0xffffffffu, v, 0);
        u32 my_slot = 0; bool flagged = false;
        if (v < n_long) { if (lane == 0) { my_slot = A.flag_list[v]; flagged = true; } }
        else {
            const u32 k0 = (v - n_long) * RR_GRAB;
            if (k0 >= n_norm) break;
            if (lane < RR_GRAB && k0 + lane < n_norm) { my_slot = A.flag_list[A.n_slots - 1u - (k0 + lane)]; flagged = true; }
        }
        u32 todo = __ballot_sync(0xffffffffu, flagged);
      while (todo) {
        const u32 src_lane = __ffs(todo) - 1; todo &= todo - 1;
        const u32 slot = __shfl_sync(0xffffffffu, my_slot, src_lane);
        SlotMeta m = A.meta[slot];
        WarpCtx S; S.L = m.len; S.W = (S.L + 31) >> 5; S.thr = m.thr; S.nhit = m.nhit; S.overflow = false; S.item = m.item;
        S.hits = slot_hits(A, m.item, S.cap);
        S.mycnt = ((const u16 *)&A.cnt[slot])[lane];
        S.keys = keys; S.kcap = RR_KEYS;
        __syncwarp();
        for (u32 i = lane; i < min(S.nhit, RR_KEYS); i += 32) { const DevHit hh = S.hits[i]; keys[i] = hit_key(hh.loc, HIT_CHR2(hh.tag) >> 1, HIT_GAPPED(hh.tag)); }
        __syncwarp();
        const u32 L = S.L, W = S.W;
        const u32 lastb = L & 31u; const u64 endmask = lastb ? (~0ULL << (64 - 2 * lastb)) : ~0ULL;
        const u32 hits_before = S.nhit;
        bool stop_all = false; u32 un_lookups = 0, un_cand = 0;     // work the reference never did because SnpAlign returned early
        for (u32 c = 0; c < 2; c++) {
            if (!(m.flags & (c ? SF_CHAIN1 : SF_CHAIN0))) continue;
            const uint2 si = A.slot_item[(u64)slot * 2 + c];
            const u32 ni = si.y;
            // ---- the items (non-empty buckets) of this chain, one per lane; an item ends where the next one begins
            u32 pm = 0, py = 0, pz = 0, pw_ = 0, ph_h = 0, pphase = 0, pbase = 0;
            if (lane < ni) {
                const uint4 a = *(const uint4 *)(A.hdr + si.x + lane);
                pbase = a.x; py = a.y; pz = a.z; pw_ = a.w; ph_h = IH_H(a.w); pphase = IH_PHASE(a.y);
            } else if (lane == ni) pbase = si.x + ni < n_items_rc ? A.hdr[si.x + ni].base : n_cands_rc;
            { const u32 nxt = __shfl_down_sync(0xffffffffu, pbase, 1); if (lane < ni) pm = nxt - pbase; }
            u32 incl = pm;
            for (u32 o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const u32 poff = incl - pm; const u32 total = __shfl_sync(0xffffffffu, incl, 31);
            if (stop_all) { un_lookups += A.I; un_cand += total; continue; }            // an abort in chain 0 ends the read's SnpAlign calls of this mode
            if (total == 0) continue;
            const u32 cbase = __shfl_sync(0xffffffffu, pbase, 0);
            S.chain = c;
            __syncwarp();
            if (lane < 16) {
                const u64 *src = A.planes + ((u64)slot * 2 + c) * 3 * A.Wb;
                bool in = lane < A.Wb;
                pq[lane] = in ? swap32(src[lane]) : 0; pn[lane] = in ? swap32(src[A.Wb + lane]) : 0; pc[lane] = in ? swap32(src[2 * A.Wb + lane]) : 0;   // pn: 01 per ACGT base
            }
            __syncwarp();
            // ---- the marked candidates of the chain in discovery (= flat) order, 32 per trip: a lane takes one bitmap word of a
            //      1024-candidate stretch, a prefix sum ranks the set bits, lane k of a trip finds the k-th of them
            const u32 fend = cbase + total;
            for (u32 wc = cbase >> 5; wc * 32u < fend && !stop_all; wc += 32) {
              u32 mbits = 0;
              {
                const u32 f0 = (wc + lane) * 32u;
                if (f0 < fend) {
                    mbits = A.bitmap[wc + lane];
                    if (f0 < cbase) mbits &= 0xffffffffu << (cbase - f0);
                    if (f0 + 32u > fend) mbits &= 0xffffffffu >> (f0 + 32u - fend);
                }
              }
              const u32 mcnt = __popc(mbits); u32 mincl = mcnt;
              for (u32 o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, mincl, o); if (lane >= o) mincl += v; }
              const u32 nmk = __shfl_sync(0xffffffffu, mincl, 31);
              for (u32 k0 = 0; k0 < nmk && !stop_all; k0 += 32) {
                const u32 rank = k0 + lane; const bool valid = rank < nmk;
                u32 wsel = 0;                                                   // words whose inclusive count is <= rank come before the one that holds it
                for (u32 o = 16; o; o >>= 1) { const u32 v = __shfl_sync(0xffffffffu, mincl, wsel + o - 1u); if (v <= rank) wsel += o; }
                const u32 wexcl = __shfl_sync(0xffffffffu, mincl - mcnt, wsel), wbits = __shfl_sync(0xffffffffu, mbits, wsel);
                const u32 flat = (wc + wsel) * 32u + (valid ? __fns(wbits, 0u, (int)(rank - wexcl + 1u)) : 0u);
                const u32 idx = valid ? flat - cbase : 0u;
                u32 ph = 0;
                for (u32 i = 1; i < ni; i++) { u32 o = __shfl_sync(0xffffffffu, poff, i); if (idx >= o) ph = i; }
                const u32 t = idx - __shfl_sync(0xffffffffu, poff, ph);
                const u32 cy = __shfl_sync(0xffffffffu, py, ph), cz = __shfl_sync(0xffffffffu, pz, ph), cw = __shfl_sync(0xffffffffu, pw_, ph);
                const u32 ch = __shfl_sync(0xffffffffu, ph_h, ph);
                u32 g = 0, sig = 0, rel = 0;
                u64 *win = win_all + lane * NWS;
                if (valid) {
                    sig = ih_strand(t, cy, cz, cw);                             // forward-strand entries come first (align.cpp:296)
                    g = A.flat_loc[flat] - ch;                                  // _hit.loc (align.cpp:297)
                    const u32 gb = g - G; const u32 word0 = gb >> 5; rel = (gb & 31u) + G;      // window starts at word0; alignment starts `rel` bases into it
                    const u64 *P = A.di.plane[sig] + word0;
                    for (u32 w = 0; w < NW; w++) win[w] = __ldg(P + w);
                }
                u32 snp = 0xffffu; bool gscr = false;
                if (valid) {
                    snp = 0;
                    for (u32 i = 0; i < W; i++) {                              // CountMismatch / CountMismatch_new
                        const u64 d = bsl_diff<SINGLE>(pq[i], pc[i], ref_word(win, NW, rel, i));
                        snp += __popcll(bsl_pairs(d) & pn[i]);
                    }
                    if (G) gscr = gap_possible<SINGLE>(win, NW, rel, pq, pc, L, W, endmask, S.thr, G);        // necessary for GapAlign to find anything
                }
                // ---- in-order reduction (AddHit semantics need the discovery order)
                const u32 thr0 = S.thr;
                const bool cand = valid && (snp <= thr0 || gscr);
                const u32 myseq = cand ? seq_of(A, g) : 0u;                       // every lane finds the sequence of its own candidate
                u32 pending = __ballot_sync(0xffffffffu, cand);
                while (pending) {
                    const u32 l = __ffs(pending) - 1; pending &= pending - 1;
                    const u32 cg = __shfl_sync(0xffffffffu, g, l), csig = __shfl_sync(0xffffffffu, sig, l), csnp = __shfl_sync(0xffffffffu, snp, l);
                    const u32 cseq = __shfl_sync(0xffffffffu, myseq, l);
                    bool ab = false;
                    if (csnp <= S.thr) ab = add_hit(A, S, lane, csnp, cg, cseq, csig, 0, 0);
                    if (G && !ab && __shfl_sync(0xffffffffu, (u32)gscr, l)) {       // GapAlign with the threshold as AddHit has left it (align.cpp:311-312)
                        const u32 cres = gap_search_warp<SINGLE>(win_all + l * NWS, NW, __shfl_sync(0xffffffffu, rel, l), pq, pc, L, W, endmask, S.thr,
                                                                 __shfl_sync(0xffffffffu, ch, l), A.s, G, lane);
                        if (cres != GAP_NONE) ab = add_hit(A, S, lane, cres & 255u, cg, cseq, csig, (int)((cres >> 8) & 255u) - 4, cres >> 16);
                    }
                    if (ab) {                    // SnpAlign returns here: later phases / the other chain are never looked up
                        const u32 aitem = __shfl_sync(0xffffffffu, ph, l);
                        const u32 aph = __shfl_sync(0xffffffffu, pphase, aitem);
                        un_lookups += A.I - 1 - aph;
                        un_cand += total - (__shfl_sync(0xffffffffu, poff, aitem) + __shfl_sync(0xffffffffu, pm, aitem));
                        stop_all = true; break;
                    }
                }
              }
            }
        }
        st_hits += S.nhit - hits_before;
        // ---- write back
        ((u16 *)&A.cnt[slot])[lane] = (u16)S.mycnt;
        const u32 nz = __ballot_sync(0xffffffffu, S.mycnt > 0);
        if (lane == 0) {
            u32 fl = m.flags;
            if (S.overflow) { fl |= SF_OVERFLOW; atomicAdd(&A.ctr->overflow_n, 1u); }
            SlotMeta *mp = A.meta + slot; mp->thr = (u8)S.thr; mp->nhit = (u16)S.nhit; mp->flags = (u8)fl; if (S.item != m.item) mp->item = S.item;
            if (un_lookups | un_cand) { uint2 ss = A.stat[slot]; ss.x -= un_lookups; ss.y -= un_cand; A.stat[slot] = ss; }
            const u32 lv = (nz | (nz >> 16)) & 0xffffu;
            A.minlvl[slot] = lv ? (u8)(__ffs(lv) - 1) : (u8)255;
        }
        __syncwarp();
      }
    }
    if (lane == 0 && st_hits) atomicAdd(&A.ctr->hits_added, st_hits);
}

// ------------------------------------------------------------------------------------------------
// pair_round : SortHits4PE + GetPairs (align.cpp:412-416, pairs.cpp:29-177), one thread per pair
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hit_less(const DevHit &a, const DevHit &b) {             // HitComp, utilities.cpp:51
    u32 ca = HIT_CHR2(a.tag), cb = HIT_CHR2(b.tag); return ca < cb || (ca == cb && a.loc < b.loc);
}
__device__ __forceinline__ bool tag_is(u32 tag, u32 chain, u32 level) { return HIT_CHAIN(tag) == chain && HIT_LEVEL(tag) == level; }

// stable insertion sort of the sub-sequence of `h[0..n)` with (chain, level), in place
__device__ void sort_level(DevHit *h, u32 n, u32 chain, u32 level) {
    for (u32 i = 0; i < n; i++) {
        if (!tag_is(h[i].tag, chain, level)) continue;
        u32 cur = i;
        for (;;) {
            int p = (int)cur - 1; while (p >= 0 && !tag_is(h[p].tag, chain, level)) p--;
            if (p < 0) break;
            DevHit a = h[p], b = h[cur];
            if (!hit_less(b, a)) break;
            h[p] = b; h[cur] = a; cur = (u32)p;
        }
    }
}

struct PairPick { u32 found; u32 chain, na, nb, insert; DevHit a, b; };

// One GetPairs(na, nb) call. mode 0: count into cnt_level (with the -w return). mode 1: same walk, but
// additionally captures the `target`-th pair (0-based position in the level list) or emits all (-r 2).
__device__ u32 get_pairs(const KArgs &A, const DevHit *ha, u32 nA, u32 Ba, u32 La, const DevHit *hb, u32 nB, u32 Bb, u32 Lb,
                         u32 na, u32 nb, u32 &cnt_level, int mode, u32 target, PairPick *pick, u64 all_base) {
    if (na > Ba || nb > Bb) return 0;
    u32 npair = 0;
    for (u32 chain = 0; chain < 2; chain++) {
        u32 chra = ~0u, bs = 0, be = 0;                     // raw positions in hb delimiting the candidates of the current chr
        for (u32 i = 0; i < nA; i++) {
            if (!tag_is(ha[i].tag, chain, na)) continue;
            const u32 ca = HIT_CHR2(ha[i].tag);
            if (chra != ca) {
                chra = ca;
                for (bs = be; bs < nB; bs++) if (tag_is(hb[bs].tag, 1 - chain, nb) && HIT_CHR2(hb[bs].tag) >= chra) break;
                for (be = bs; be < nB; be++) if (tag_is(hb[be].tag, 1 - chain, nb) && HIT_CHR2(hb[be].tag) > chra) break;
            }
            for (u32 j = bs; j < be; j++) {
                if (!tag_is(hb[j].tag, 1 - chain, nb)) continue;
                const bool a_first = chain == 0 ? !(chra & 1u) : (chra & 1u);
                u32 s0, e0;
                if (a_first) { s0 = ha[i].loc; e0 = hb[j].loc + Lb; } else { s0 = hb[j].loc; e0 = ha[i].loc + La; }
                const u32 ins = e0 - s0;
                if (ins >= A.min_insert && ins <= A.max_insert) {
                    if (mode == 1 && pick) {
                        if (A.report == 2 && A.all_a && all_base + cnt_level < A.all_cap) {
                            bsl_hit ra, rb; memset(&ra, 0, sizeof ra); memset(&rb, 0, sizeof rb);
                            ra.loc = ha[i].loc; ra.chr = HIT_CHR2(ha[i].tag); ra.gap_size = (int)ha[i].gap; ra.gap_pos = (u16)ha[i].gp; ra.nm = (u8)na; ra.read_chain = (u8)chain; ra.read_len = (u16)La; ra.status = BSL_ST_PAIRED; ra.all_first = ins;
                            rb.loc = hb[j].loc; rb.chr = HIT_CHR2(hb[j].tag); rb.gap_size = (int)hb[j].gap; rb.gap_pos = (u16)hb[j].gp; rb.nm = (u8)nb; rb.read_chain = (u8)(1 - chain); rb.read_len = (u16)Lb; rb.status = BSL_ST_PAIRED; rb.all_first = ins;
                            A.all_a[all_base + cnt_level] = ra; A.all_b[all_base + cnt_level] = rb;
                        }
                        if (cnt_level == target) { pick->found = 1; pick->chain = chain; pick->na = na; pick->nb = nb; pick->insert = ins; pick->a = ha[i]; pick->b = hb[j]; }
                    }
                    cnt_level++; npair++;
                    if (cnt_level >= A.w) return npair;
                }
            }
        }
    }
    return npair;
}

__device__ void fill_record(bsl_hit &o, const DevHit &h, u32 chain, u32 level) {
    o.loc = h.loc; o.chr = HIT_CHR2(h.tag); o.gap_size = (int)h.gap; o.gap_pos = (u16)h.gp; o.nm = (u8)level; o.read_chain = (u8)chain;
}

#define PR_LOCAL 4u
#define PR_WIDE_HITS 12u     // a pair with a longer hit list goes to pair_round_wide
#define PW_CAP 2048u        // longest list the shared-memory path handles; longer ones take the serial path (lane 0)
#define PW_CAP_SMALL 256u   // most long lists are shorter than this: their pairs run in a second instance with eight times the resident warps
// Rounds [round, r_hi]: after the last search round no list changes any more, so the remaining rounds of a pair (levels up to its
// budget) are replayed by the same thread in one launch.
__global__ void pair_round(const __grid_constant__ KArgs A, u32 round0, u32 r_hi, const u32 *list_in, u32 *list_out, u32 ci) {
    RoundCtr *rc = A.ctr->rc + ci;
    const u32 n_items = rc->active;
    // one round of one pair; true = no pair yet and levels left: the pair stays listed
    auto one = [&](const u32 p, const u32 round) -> bool {
        const u32 sa = p, sb = p + A.n_a;
        SlotMeta ma = A.meta[sa], mb = A.meta[sb];
        if ((ma.flags | mb.flags) & SF_OVERFLOW) {                 // re-run the whole pair on the large-capacity path
            if (!(ma.flags & SF_OVERFLOW)) { A.meta[sa].flags = ma.flags | SF_OVERFLOW; }
            if (!(mb.flags & SF_OVERFLOW)) { A.meta[sb].flags = mb.flags | SF_OVERFLOW; }
            return false;
        }
        if ((ma.flags | mb.flags) & SF_FILTERED) {                 // a lone mate (the other one was filtered) runs SingleAlign::RunAlign in the reference
            return round < max((u32)ma.B, (u32)mb.B);                // (pairs.cpp:197-201): no SortHits4PE, no pairs
        }
        if (((ma.item | mb.item) & SLOT_BIG) || ma.nhit > PR_WIDE_HITS || mb.nhit > PR_WIDE_HITS) {      // long lists: warp per pair
            if (max((u32)ma.nhit, (u32)mb.nhit) <= PW_CAP_SMALL) A.wide_list[atomicAdd(&rc->wide, 1u)] = p;
            else A.wide_list[A.n_a - 1u - atomicAdd(&rc->long_work, 1u)] = p;                            // RoundCtr::long_work: pairs with a list beyond PW_CAP_SMALL
            return false;
        }
        u32 capa, capb; DevHit *ha = slot_hits(A, ma.item, capa), *hb = slot_hits(A, mb.item, capb);
        const u32 nA = ma.nhit, nB = mb.nhit, Ba = ma.B, Bb = mb.B, La = ma.len, Lb = mb.len;
        const u32 i = round;
        u32 total = 0, best = 0xffffffffu, best_cnt = 0;
        // short lists (the usual case: one or two hits per mate) are replayed on a thread-local copy, so that the scans of
        // sort_level / get_pairs are L1 hits instead of dependent global loads; sorted lists are written back
        DevHit la[PR_LOCAL], lb[PR_LOCAL];
        DevHit *ga = ha, *gb = hb;
        const bool local = nA <= PR_LOCAL && nB <= PR_LOCAL;
        if (local) {
#pragma unroll
            for (u32 x = 0; x < PR_LOCAL; x++) { if (x < nA) la[x] = ga[x]; if (x < nB) lb[x] = gb[x]; }
            ha = la; hb = lb;
        }
        if (i <= Ba) { sort_level(ha, nA, 0, i); sort_level(ha, nA, 1, i); }      // SortHits4PE happens even when the mate has no hit yet
        if (i <= Bb) { sort_level(hb, nB, 0, i); sort_level(hb, nB, 1, i); }
        if (local) {
            if (nA > 1 && i <= Ba) for (u32 x = 0; x < nA; x++) ga[x] = la[x];
            if (nB > 1 && i <= Bb) for (u32 x = 0; x < nB; x++) gb[x] = lb[x];
        }
        if (nA && nB) {
            // pass 1: count per level sum in the reference's call order
            u32 c = 0; total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 0, 0, nullptr, 0);
            if (c) { best = 2 * i; best_cnt = c; }
            for (u32 j = 0; j < i; j++) {
                u32 cj = 0;
                total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, cj, 0, 0, nullptr, 0);
                total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, cj, 0, 0, nullptr, 0);
                if (cj && i + j < best) { best = i + j; best_cnt = cj; }
            }
        }
        if (total == 0) return round < max(Ba, Bb);
        // a pair exists: the search ends here (pairs.cpp:173)
        bsl_pair pr; memset(&pr, 0, sizeof pr); pr.n_pairs = best_cnt;
        if (best_cnt > 1 && A.report == 0) { A.pair_out[p] = pr; return false; }   // suppressed; mates get reported unpaired by finalize
        const u32 target = best_cnt == 1 ? 0 : ma.rnd % best_cnt;
        PairPick pk; pk.found = 0;
        u64 all_base = 0;
        if (A.report == 2 && A.all_a) { all_base = atomicAdd(&A.ctr->all_n, (unsigned long long)best_cnt); pr.all_first = (u32)(all_base + A.all_off); }
        {
            u32 c = 0;
            if (best == 2 * i) get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 1, target, &pk, all_base);
            else { u32 j = best - i;
                get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, c, 1, target, &pk, all_base);
                get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, c, 1, target, &pk, all_base); }
        }
        pr.insert = pk.insert; pr.chain = (u8)pk.chain; pr.na = (u8)pk.na; pr.nb = (u8)pk.nb;
        A.pair_out[p] = pr;
        bsl_hit oa, ob; memset(&oa, 0, sizeof oa); memset(&ob, 0, sizeof ob);
        fill_record(oa, pk.a, pk.chain, pk.na); fill_record(ob, pk.b, 1 - pk.chain, pk.nb);
        oa.status = ob.status = BSL_ST_PAIRED; oa.n_hits = ob.n_hits = best_cnt; oa.read_len = (u16)La; ob.read_len = (u16)Lb; oa.max_snp = (u8)Ba; ob.max_snp = (u8)Bb;
        A.out[sa] = oa; A.out[sb] = ob;
        A.meta[sa].flags = ma.flags | SF_DONE; A.meta[sb].flags = mb.flags | SF_DONE;
        return false;
    };
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n_items; k += gridDim.x * blockDim.x) {
        const u32 p = list_in[k];
        for (u32 r = round0; one(p, r); r++) if (r >= r_hi) { list_out[atomicAdd(&rc[1].active, 1u)] = p; break; }
    }
}

// ------------------------------------------------------------------------------------------------
// pair_round_wide : the same replay with one WARP per pair, for pairs whose hit lists are long (reads from
// repeat families on the large-capacity pass). The per-(chain, level) sub-lists are compacted into shared
// memory, sorted by rank (stable, like sort_level), and the A x window(B) enumeration of GetPairs is spread
// over the lanes; positions in the reference's enumeration order come from prefix sums, so the -w cut, the
// -S pick and the -r 2 listing are the same pairs in the same order.
// ------------------------------------------------------------------------------------------------
template <u32 CAP>
struct PairWideSmemT {
    union { DevHit stage[CAP]; struct { u32 a_loc[CAP], a_chr[CAP], b_loc[CAP], b_chr[CAP]; } k; };
    u16 ia[CAP], ib[CAP], wbs[CAP], wbe[CAP];
    u8 cla[CAP], clb[CAP];                  // chain << 4 | level of every hit of mate a / b (SortHits4PE only permutes hits of equal chain and level)
    u32 hist[2][32];                        // how many hits of mate a / b carry each (chain, level)
    PairPick pick;
};
typedef PairWideSmemT<PW_CAP> PairWideSmem;
#define HIT_CL(tag) (((tag) >> 20) & 31u)

// indices (in list order) of the hits of `h[0..n)` tagged (chain, level) -> idx[]; returns how many (lists longer than PW_CAP are cut: caller checks n first)
__device__ u32 pw_compact(const u8 *cl, u32 n, u32 chain, u32 level, u16 *idx, u32 lane) {
    u32 c = 0; const u32 want = (chain << 4) | level;
    for (u32 t = 0; t < n; t += 32) {
        const u32 i = t + lane; const bool ok = i < n && cl[i] == want;
        const u32 bal = __ballot_sync(0xffffffffu, ok);
        if (ok) idx[c + __popc(bal & ((1u << lane) - 1u))] = (u16)i;
        c += __popc(bal);
    }
    __syncwarp();
    return c;
}

// SortHits4PE's std::sort over one (chain, level) list. Up to 16 elements libstdc++ runs a plain insertion sort (stable), and a
// list without equal (chr, loc) keys has only one sorted order: both are the stable rank sort below. A longer list that holds
// equal keys (a gapped and an ungapped hit at the same place) comes out of introsort in an order of its own, which stdsort()
// reproduces step by step (SURVEY trap 10).
template <class SM>
__device__ void pw_sort_level(DevHit *h, const u8 *cl, const u32 *hist, u32 n, u32 chain, u32 level, SM &sm, u32 lane) {
    if (hist[(chain << 4) | level] <= 1) return;
    const u32 c = pw_compact(cl, n, chain, level, sm.ia, lane);
    for (u32 x = lane; x < c; x += 32) sm.stage[x] = h[sm.ia[x]];
    __syncwarp();
    bool twins = false;
    for (u32 x0 = 0; x0 < c; x0 += 32) {
        const u32 x = x0 + lane; u32 rank = 0; DevHit me; me.loc = 0; me.tag = 0; me.gap = 0; me.gp = 0; bool eq = false;
        if (x < c) {
            me = sm.stage[x]; const u64 kx = ((u64)HIT_CHR2(me.tag) << 32) | me.loc;
            for (u32 f = 0; f < c; f++) { const u64 kf = ((u64)HIT_CHR2(sm.stage[f].tag) << 32) | sm.stage[f].loc; rank += (kf < kx || (kf == kx && f < x)) ? 1u : 0u; eq |= kf == kx && f != x; }
        }
        twins |= __any_sync(0xffffffffu, eq);
        if (x < c && (c <= 16 || !twins)) h[sm.ia[rank]] = me;              // provisional when a later tile finds twins: rewritten below
    }
    __syncwarp();
    if (c > 16 && twins) {
        u16 *p = sm.ib;
        if (lane == 0) {
            for (u32 x = 0; x < c; x++) p[x] = (u16)x;
            stdsort(p, (int)c, [&](u16 a, u16 b) { const DevHit &A_ = sm.stage[a], &B_ = sm.stage[b]; const u32 ca = HIT_CHR2(A_.tag), cb = HIT_CHR2(B_.tag); return ca < cb || (ca == cb && A_.loc < B_.loc); });
        }
        __syncwarp();
        for (u32 x = lane; x < c; x += 32) h[sm.ia[x]] = sm.stage[p[x]];
        __syncwarp();
    }
}

__device__ __forceinline__ bool pw_valid(const KArgs &A, u32 chain, u32 chra, u32 aloc, u32 bloc, u32 La, u32 Lb, u32 &ins) {
    const bool a_first = chain == 0 ? !(chra & 1u) : (chra & 1u);
    u32 s0, e0;
    if (a_first) { s0 = aloc; e0 = bloc + Lb; } else { s0 = bloc; e0 = aloc + La; }
    ins = e0 - s0;
    return ins >= A.min_insert && ins <= A.max_insert;
}

// warp version of get_pairs; every argument and the return value are warp-uniform
template <class SM>
__device__ u32 pw_get_pairs(const KArgs &A, SM &sm, u32 lane, const DevHit *ha, u32 nA, u32 Ba, u32 La, const DevHit *hb, u32 nB, u32 Bb, u32 Lb,
                            u32 na, u32 nb, u32 &cnt_level, int mode, u32 target, u64 all_base) {
    if (na > Ba || nb > Bb) return 0;
    u32 npair = 0;
    for (u32 chain = 0; chain < 2; chain++) {
        if (sm.hist[0][(chain << 4) | na] == 0 || sm.hist[1][((1 - chain) << 4) | nb] == 0) continue;
        const u32 cA = pw_compact(sm.cla, nA, chain, na, sm.ia, lane);
        const u32 cB = pw_compact(sm.clb, nB, 1 - chain, nb, sm.ib, lane);
        for (u32 x = lane; x < cA; x += 32) { const DevHit t = ha[sm.ia[x]]; sm.k.a_loc[x] = t.loc; sm.k.a_chr[x] = HIT_CHR2(t.tag); }
        for (u32 x = lane; x < cB; x += 32) { const DevHit t = hb[sm.ib[x]]; sm.k.b_loc[x] = t.loc; sm.k.b_chr[x] = HIT_CHR2(t.tag); }
        __syncwarp();
        // windows of B per A element: the two-pointer walk of GetPairs (state carried from element to element)
        {
            u32 chra = ~0u, bs = 0, be = 0;
            for (u32 x = 0; x < cA; x++) {
                const u32 ca = sm.k.a_chr[x];
                if (ca != chra) {
                    chra = ca;
                    u32 t = be; bs = cB;
                    for (; t < cB; t += 32) { const u32 bal = __ballot_sync(0xffffffffu, t + lane < cB && sm.k.b_chr[t + lane] >= chra); if (bal) { bs = t + __ffs(bal) - 1; break; } }
                    t = bs; be = cB;
                    for (; t < cB; t += 32) { const u32 bal = __ballot_sync(0xffffffffu, t + lane < cB && sm.k.b_chr[t + lane] > chra); if (bal) { be = t + __ffs(bal) - 1; break; } }
                }
                if (lane == 0) { sm.wbs[x] = (u16)bs; sm.wbe[x] = (u16)be; }
            }
        }
        __syncwarp();
        for (u32 x0 = 0; x0 < cA; x0 += 32) {
            const u32 x = x0 + lane; u32 cnt = 0, aloc = 0, chra = 0, bs = 0, be = 0;
            if (x < cA) {
                aloc = sm.k.a_loc[x]; chra = sm.k.a_chr[x]; bs = sm.wbs[x]; be = sm.wbe[x];
                for (u32 j = bs; j < be; j++) { u32 ins; cnt += pw_valid(A, chain, chra, aloc, sm.k.b_loc[j], La, Lb, ins) ? 1u : 0u; }
            }
            u32 incl = cnt;
            for (u32 o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const u32 tile = __shfl_sync(0xffffffffu, incl, 31);
            if (tile == 0) continue;
            const u32 room = cnt_level >= A.w ? 1u : A.w - cnt_level;          // pairs this call may still take before GetPairs returns
            const u32 take = min(tile, room);
            if (mode == 1 && cnt) {
                u32 q = incl - cnt;                                          // order of my first pair within the tile
                for (u32 j = bs; j < be && q < take; j++) {
                    u32 ins;
                    if (!pw_valid(A, chain, chra, aloc, sm.k.b_loc[j], La, Lb, ins)) continue;
                    const u32 pos = cnt_level + q;
                    const DevHit ta = ha[sm.ia[x]], tb = hb[sm.ib[j]];
                    if (A.report == 2 && A.all_a && all_base + pos < A.all_cap) {
                        bsl_hit ra, rb; memset(&ra, 0, sizeof ra); memset(&rb, 0, sizeof rb);
                        ra.loc = ta.loc; ra.chr = HIT_CHR2(ta.tag); ra.gap_size = (int)ta.gap; ra.gap_pos = (u16)ta.gp; ra.nm = (u8)na; ra.read_chain = (u8)chain; ra.read_len = (u16)La; ra.status = BSL_ST_PAIRED; ra.all_first = ins;
                        rb.loc = tb.loc; rb.chr = HIT_CHR2(tb.tag); rb.gap_size = (int)tb.gap; rb.gap_pos = (u16)tb.gp; rb.nm = (u8)nb; rb.read_chain = (u8)(1 - chain); rb.read_len = (u16)Lb; rb.status = BSL_ST_PAIRED; rb.all_first = ins;
                        A.all_a[all_base + pos] = ra; A.all_b[all_base + pos] = rb;
                    }
                    if (pos == target) { PairPick &pk = sm.pick; pk.found = 1; pk.chain = chain; pk.na = na; pk.nb = nb; pk.insert = ins; pk.a = ta; pk.b = tb; }
                    q++;
                }
            }
            cnt_level += take; npair += take;
            if (take == room) { __syncwarp(); return npair; }               // cnt_level reached -w: GetPairs returns
        }
        __syncwarp();
    }
    return npair;
}

// from_wide = 0: every pair of list_in (large-capacity pass); 1: the pairs pair_round of this round left at the front of
// KArgs::wide_list (no list longer than PW_CAP_SMALL); 2: those it left at the back
template <u32 CAP>
__global__ void __launch_bounds__(32) pair_round_wide(const __grid_constant__ KArgs A, u32 round0, u32 r_hi, const u32 *list_in, u32 *list_out, u32 ci, u32 from_wide) {
    extern __shared__ __align__(16) unsigned char pw_raw[];
    typedef PairWideSmemT<CAP> SM;
    SM &sm = *reinterpret_cast<SM *>(pw_raw);
    RoundCtr *rc = A.ctr->rc + ci;
    const u32 n_items = from_wide == 1 ? rc->wide : (from_wide == 2 ? rc->long_work : rc->active), lane = threadIdx.x;
    // one round of one pair (warp-uniform); true = no pair yet and levels left
    auto one = [&](const u32 p, const u32 round) -> bool {
        const u32 sa = p, sb = p + A.n_a;
        const SlotMeta ma = A.meta[sa], mb = A.meta[sb];
        __syncwarp();
        if ((ma.flags | mb.flags) & SF_OVERFLOW) {
            if (lane == 0) { if (!(ma.flags & SF_OVERFLOW)) A.meta[sa].flags = ma.flags | SF_OVERFLOW; if (!(mb.flags & SF_OVERFLOW)) A.meta[sb].flags = mb.flags | SF_OVERFLOW; }
            return false;
        }
        if ((ma.flags | mb.flags) & SF_FILTERED) return round < max((u32)ma.B, (u32)mb.B);      // lone mate: see pair_round
        u32 capa, capb; DevHit *ha = slot_hits(A, ma.item, capa), *hb = slot_hits(A, mb.item, capb);
        const u32 nA = ma.nhit, nB = mb.nhit, Ba = ma.B, Bb = mb.B, La = ma.len, Lb = mb.len;
        const u32 i = round;
        u32 total = 0, best = 0xffffffffu, best_cnt = 0;
        const bool wide = nA <= CAP && nB <= CAP;
        if (wide) {
            // one pass over both lists: (chain, level) of every hit into shared memory, and how many of each there are — most
            // (chain, level) combinations of a pair are empty and cost nothing below
            __syncwarp();
            sm.hist[0][lane] = 0; sm.hist[1][lane] = 0;
            __syncwarp();
            for (u32 x = lane; x < nA; x += 32) { const u32 cl = HIT_CL(ha[x].tag); sm.cla[x] = (u8)cl; atomicAdd(&sm.hist[0][cl], 1u); }
            for (u32 x = lane; x < nB; x += 32) { const u32 cl = HIT_CL(hb[x].tag); sm.clb[x] = (u8)cl; atomicAdd(&sm.hist[1][cl], 1u); }
            __syncwarp();
            if (i <= Ba) { pw_sort_level(ha, sm.cla, sm.hist[0], nA, 0, i, sm, lane); pw_sort_level(ha, sm.cla, sm.hist[0], nA, 1, i, sm, lane); }
            if (i <= Bb) { pw_sort_level(hb, sm.clb, sm.hist[1], nB, 0, i, sm, lane); pw_sort_level(hb, sm.clb, sm.hist[1], nB, 1, i, sm, lane); }
            if (nA && nB) {
                u32 c = 0; total += pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 0, 0, 0);
                if (c) { best = 2 * i; best_cnt = c; }
                for (u32 j = 0; j < i; j++) {
                    u32 cj = 0;
                    total += pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, cj, 0, 0, 0);
                    total += pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, cj, 0, 0, 0);
                    if (cj && i + j < best) { best = i + j; best_cnt = cj; }
                }
            }
        } else {
            if (lane == 0) {
                if (i <= Ba) { sort_level(ha, nA, 0, i); sort_level(ha, nA, 1, i); }
                if (i <= Bb) { sort_level(hb, nB, 0, i); sort_level(hb, nB, 1, i); }
                if (nA && nB) {
                    u32 c = 0; total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 0, 0, nullptr, 0);
                    if (c) { best = 2 * i; best_cnt = c; }
                    for (u32 j = 0; j < i; j++) {
                        u32 cj = 0;
                        total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, cj, 0, 0, nullptr, 0);
                        total += get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, cj, 0, 0, nullptr, 0);
                        if (cj && i + j < best) { best = i + j; best_cnt = cj; }
                    }
                }
            }
            total = __shfl_sync(0xffffffffu, total, 0); best = __shfl_sync(0xffffffffu, best, 0); best_cnt = __shfl_sync(0xffffffffu, best_cnt, 0);
        }
        if (total == 0) return round < max(Ba, Bb);
        bsl_pair pr; memset(&pr, 0, sizeof pr); pr.n_pairs = best_cnt;
        if (best_cnt > 1 && A.report == 0) { if (lane == 0) A.pair_out[p] = pr; return false; }
        const u32 target = best_cnt == 1 ? 0 : ma.rnd % best_cnt;
        u64 all_base = 0;
        if (A.report == 2 && A.all_a) {
            if (lane == 0) all_base = atomicAdd(&A.ctr->all_n, (unsigned long long)best_cnt);
            all_base = __shfl_sync(0xffffffffu, all_base, 0); pr.all_first = (u32)(all_base + A.all_off);
        }
        if (lane == 0) sm.pick.found = 0;
        __syncwarp();
        if (wide) {
            u32 c = 0;
            if (best == 2 * i) pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 1, target, all_base);
            else { const u32 j = best - i;
                pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, c, 1, target, all_base);
                pw_get_pairs(A, sm, lane, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, c, 1, target, all_base); }
        } else if (lane == 0) {
            u32 c = 0; PairPick pk; pk.found = 0;
            if (best == 2 * i) get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, i, c, 1, target, &pk, all_base);
            else { const u32 j = best - i;
                get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, i, j, c, 1, target, &pk, all_base);
                get_pairs(A, ha, nA, Ba, La, hb, nB, Bb, Lb, j, i, c, 1, target, &pk, all_base); }
            sm.pick = pk;
        }
        __syncwarp();
        if (lane == 0) {
            const PairPick pk = sm.pick;
            pr.insert = pk.insert; pr.chain = (u8)pk.chain; pr.na = (u8)pk.na; pr.nb = (u8)pk.nb;
            A.pair_out[p] = pr;
            bsl_hit oa, ob; memset(&oa, 0, sizeof oa); memset(&ob, 0, sizeof ob);
            fill_record(oa, pk.a, pk.chain, pk.na); fill_record(ob, pk.b, 1 - pk.chain, pk.nb);
            oa.status = ob.status = BSL_ST_PAIRED; oa.n_hits = ob.n_hits = best_cnt; oa.read_len = (u16)La; ob.read_len = (u16)Lb; oa.max_snp = (u8)Ba; ob.max_snp = (u8)Bb;
            A.out[sa] = oa; A.out[sb] = ob;
            A.meta[sa].flags = ma.flags | SF_DONE; A.meta[sb].flags = mb.flags | SF_DONE;
        }
        __syncwarp();
        return false;
    };
    for (u32 k = blockIdx.x; k < n_items; k += gridDim.x) {
        const u32 p = from_wide == 2 ? list_in[A.n_a - 1u - k] : list_in[k];
        for (u32 r = round0; one(p, r); r++) if (r >= r_hi) { if (lane == 0) list_out[atomicAdd(&rc[1].active, 1u)] = p; break; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// finalize_reads : StringAlign / StringAlignUnpair selection (align.cpp:583-612, pairs.cpp:232-259)
// ------------------------------------------------------------------------------------------------
__global__ void finalize_reads(const __grid_constant__ KArgs A, const u32 *only_list, u32 only_n) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = only_list ? only_n : A.n_slots;
    u32 slot = 0; bool live = t < n;
    if (live) slot = only_list ? only_list[t] : t;
    SlotMeta m; m.flags = SF_OVERFLOW;
    if (live) m = A.meta[slot];
    if (m.flags & SF_OVERFLOW) live = false;                     // will be written by the heavy pass
    // work counters: one atomic per warp
    uint2 ss = make_uint2(0u, 0u); if (live) ss = A.stat[slot];
    unsigned long long lx = ss.x, ly = ss.y;
    for (u32 o = 16; o; o >>= 1) { lx += __shfl_xor_sync(0xffffffffu, lx, o); ly += __shfl_xor_sync(0xffffffffu, ly, o); }
    if ((threadIdx.x & 31u) == 0) { if (lx) atomicAdd(&A.ctr->seed_lookups, lx); if (ly) atomicAdd(&A.ctr->candidates, ly); }
    if (!live) return;
    if (m.flags & SF_DONE) return;                               // already written by pair_round
    bsl_hit o; memset(&o, 0, sizeof o); o.read_len = m.len; o.max_snp = m.B;
    if (m.flags & (SF_FILTERED | SF_CONTEXT)) { o.status = BSL_ST_FILTERED; A.out[slot] = o; return; }
    o.status = BSL_ST_UNMAPPED;
    if (m.nhit == 0) { A.out[slot] = o; return; }
    const SlotCounts cn = A.cnt[slot];
    u32 hcap; const DevHit *h = slot_hits(A, m.item, hcap);
    for (u32 l = 0; l <= m.B; l++) {
        u32 n0 = cn.c[0][l], nn = n0 + cn.c[1][l];
        if (!nn) continue;
        u32 pick = nn == 1 ? 0 : m.rnd % nn; u32 chain = pick < n0 ? 0 : 1; u32 want = pick - (chain ? n0 : 0);
        u32 seen = 0;
        for (u32 i = 0; i < m.nhit; i++) if (tag_is(h[i].tag, chain, l)) { if (seen == want) { fill_record(o, h[i], chain, l); break; } seen++; }
        o.n_hits = nn; o.n_chain0 = n0; o.status = nn == 1 ? BSL_ST_UNIQUE : BSL_ST_MULTI;
        if (A.report == 2 && A.all_a && nn > 1) {
            // -r 2: every hit of the level, chain 0 list then chain 1 list (align.cpp:604-607; for a mate that ends unpaired
            // pairs.cpp:270-274, 293-297). Mate #2 of a pair lists into the second array; both share one index space.
            bsl_hit *dst = (A.pe && slot >= A.n_a) ? A.all_b : A.all_a;
            u64 base = atomicAdd(&A.ctr->all_n, (unsigned long long)nn); o.all_first = (u32)(base + A.all_off); u64 w = base;
            for (u32 c = 0; c < 2; c++) for (u32 i = 0; i < m.nhit; i++) if (tag_is(h[i].tag, c, l)) {
                if (w < A.all_cap) { bsl_hit r = o; fill_record(r, h[i], c, l); dst[w] = r; } w++; }
        }
        break;
    }
    A.out[slot] = o;
}

// heavy pass: reset the state of the overflowed items and point them at the large pool
__global__ void collect_overflow(const __grid_constant__ KArgs A, u32 *heavy_list, u32 *n_heavy) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 n = A.pe ? A.n_a : A.n_slots;
    if (i >= n) return;
    bool ov = A.meta[i].flags & SF_OVERFLOW;
    if (A.pe) ov = ov || (A.meta[i + A.n_a].flags & SF_OVERFLOW);
    if (ov) { u32 pos = atomicAdd(n_heavy, 1u); heavy_list[pos] = i; }
}

__global__ void reset_heavy(const __grid_constant__ KArgs A, const u32 *heavy_list, u32 first, u32 count, u32 *se_list, u32 *pe_list) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    u32 i = heavy_list[first + t];
    for (u32 mate = 0; mate < (A.pe ? 2u : 1u); mate++) {
        u32 slot = i + mate * A.n_a;
        SlotMeta m = A.meta[slot];
        m.flags &= ~(SF_OVERFLOW | SF_DONE); m.thr = m.B; m.nhit = 0; m.item = A.pe ? 2 * t + mate : t;
        A.meta[slot] = m;
        for (u32 k = 0; k < 32; k++) ((u16 *)&A.cnt[slot])[k] = 0;
        A.stat[slot] = make_uint2(0u, 0u);
        A.minlvl[slot] = 255;
    }
    if (!A.pe) { u32 pos = atomicAdd(&A.ctr->rc[0].active, 1u); se_list[pos] = i; return; }
    SlotMeta ma = A.meta[i], mb = A.meta[i + A.n_a];
    const bool oka = !(ma.flags & SF_FILTERED), okb = !(mb.flags & SF_FILTERED);
    if (oka || okb) { u32 pos = atomicAdd(&A.ctr->rc[20].active, 1u); pe_list[pos] = i; }
}

__global__ void expand_pairs_to_slots(const u32 *heavy_list, u32 first, u32 count, u32 n_a, u32 pe, u32 *slots) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    u32 i = heavy_list[first + t];
    if (pe) { slots[2 * t] = i; slots[2 * t + 1] = i + n_a; } else slots[t] = i;
}

template <typename T> int grow(bsl_ctx *ctx, T **p, size_t *cap, size_t need, bool pinned = false) {
    if (need <= *cap && *p) return 0;
    size_t ncap = std::max(need, *cap + *cap / 2);
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; }
    cudaError_t e = pinned ? cudaMallocHost((void **)p, ncap * sizeof(T)) : cudaMalloc((void **)p, ncap * sizeof(T));
    if (e != cudaSuccess) { set_error(ctx, "allocation of %zu bytes failed: %s", ncap * sizeof(T), cudaGetErrorString(e)); *cap = 0; return BSL_ENOMEM; }
    *cap = ncap; return 0;
}

} // namespace

void bsl_lane_free(Lane &ln) {
    cudaFree(ln.d_bases); cudaFree(ln.d_off); cudaFree(ln.d_index); cudaFree(ln.d_rawlen); cudaFree(ln.d_meta); cudaFree(ln.d_cnt); cudaFree(ln.d_sched); cudaFree(ln.d_stat);
    cudaFree(ln.d_minlvl); cudaFree(ln.d_slot_item); cudaFree(ln.d_slot_flag); cudaFree(ln.d_flag_list); cudaFree(ln.d_marks);
    cudaFree(ln.d_bits1); cudaFree(ln.d_hdr); cudaFree(ln.d_chunk_first); cudaFree(ln.d_bitmap); cudaFree(ln.d_flat_loc);
    cudaFree(ln.d_planes); cudaFree(ln.d_hits); cudaFree(ln.d_heavy_hits); cudaFree(ln.d_list[0]); cudaFree(ln.d_list[1]); cudaFree(ln.d_heavy_list);
    cudaFree(ln.d_pe_list[0]); cudaFree(ln.d_pe_list[1]); cudaFree(ln.d_out); cudaFree(ln.d_pair); cudaFree(ln.d_all[0]); cudaFree(ln.d_all[1]); cudaFree(ln.d_ctr);
    cudaFree(ln.d_st0); cudaFree(ln.d_defer); cudaFree(ln.d_stale); cudaFree(ln.d_bighits); cudaFree(ln.d_wide_list);
    if (ln.h_ctr) cudaFreeHost(ln.h_ctr);
    for (auto &e : ln.ev) if (e) cudaEventDestroy(e);
    for (auto &e : ln.evk) if (e) cudaEventDestroy(e);
    if (ln.stream) cudaStreamDestroy(ln.stream);
    ln.stream = nullptr;
}

int bsl_upload_params(bsl_ctx *ctx) {
    // device copy of the rule / budget / profile tables (one per context)
    DevTables h; memset(&h, 0, sizeof h);
    h.rule = ctx->rule;
    const bsl_params &P = ctx->P;
    for (u32 raw = 0; raw <= BSL_MAX_READLEN; raw++) {                       // FilterReads, align.cpp:550-556
        u32 B = P.max_snp_num < 100 ? P.max_snp_num : (u32)((P.max_snp_num - 100) / 100.0 * raw + 0.5);
        if (P.gap > 0) B = B + 1 + P.gap;
        if (B > BSL_MAXSNPS) B = BSL_MAXSNPS;
        h.budget0[raw] = B;
    }
    for (u32 i = 0; i < P.index_interval && i < 16; i++) for (u32 j = 0; j <= BSL_MAXSNPS; j++)
        h.prof[j][i] = (u16)(((j * P.seed_size + i + P.index_interval - 1) / P.index_interval) * P.index_interval);   // param.cpp:70-74
    static const char kNt[4] = {'A', 'C', 'G', 'T'};
    for (int k = 0; k < 4; k++) {                                            // 2-bit tables indexed by (ascii >> 1) & 3
        const u32 ix = ((u32)kNt[k] >> 1) & 3u;
        h.tab_code[0] |= (u32)ctx->rule.code[(u8)kNt[k]] << (2 * ix); h.tab_code[1] |= (u32)ctx->rule.rcode[(u8)kNt[k]] << (2 * ix);
        h.tab_conv[0] |= (u32)ctx->rule.conv[(u8)kNt[k]] << (2 * ix); h.tab_conv[1] |= (u32)ctx->rule.rconv[(u8)kNt[k]] << (2 * ix);
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->d_tables) CUDA_TRY(cudaMalloc(&ctx->d_tables, sizeof(DevTables)));
    CUDA_TRY(cudaMemcpy(ctx->d_tables, &h, sizeof h, cudaMemcpyHostToDevice));
    return 0;
}

static int ensure_lane(bsl_ctx *ctx, Lane &ln) {
    if (ln.stream) return 0;
    CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
    for (auto &e : ln.ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : ln.evk) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaMalloc(&ln.d_ctr, sizeof(DevCounters)));
    CUDA_TRY(cudaMallocHost(&ln.h_ctr, sizeof(DevCounters)));
    return 0;
}

// longest read of a sub-range; dims (optional) = shared-memory row sizes of prepare_reads over the read lengths present:
// dims[0] = most read offsets per seed segment the schedule can touch, I + (L - I + 1) % s; dims[1] = largest CountSeeds table,
// segments x ((L - I + 1) % s + 1) with segments <= (L - I + 1) / s (align.cpp:450, 476-480)
static u32 max_len_of(const bsl_batch *b, u32 first, u32 n, const bsl_params *P = nullptr, u32 *dims = nullptr, bool *empty_range = nullptr) {
    u32 mx = 0, last = 0xffffffffu;
    for (u32 i = first; i < first + n; i++) {
        const u64 l = b->offsets[i + 1] - b->offsets[i];
        if (l > mx) mx = (u32)std::min<u64>(l, 0xffffffffu);
        if (P && (u32)l != last) {
            last = (u32)l; const u32 L = (u32)std::min<u64>(l, BSL_MAX_READLEN), I = P->index_interval, s = P->seed_size;
            if (L + 1 >= I + s) {
                const u32 ii = (L + 1 - I) % s;
                if (dims) { dims[0] = std::max(dims[0], I + ii); dims[1] = std::max(dims[1], std::min<u32>((L + 1 - I) / s, 16) * (ii + 1)); }
                if (ii == 0 && empty_range) *empty_range = true;          // SURVEY trap 3: this read inherits its start offset
            }
        }
    }
    return mx;
}

// ---- carried aligner state (SURVEY trap 3) across the sub-ranges of one call: which reads of [0, upto) re-create, when
// they precede read `upto`, the state a single aligner object would have there — per mate the last unfiltered read with
// a start-offset range, and every unfiltered read longer than all unfiltered reads after it (its seed hashes are still
// visible beyond the end of a shorter read). Ascending indices.
static void carry_context(const bsl_params &P, const bsl_batch *a, const bsl_batch *b, u32 upto, std::vector<u32> &idx) {
    idx.clear();
    const u32 I = P.index_interval, s = P.seed_size; const int nm = b ? 2 : 1;
    u32 maxlen[2] = {0, 0}; bool have_def[2] = {false, b == nullptr};
    auto defining = [&](u32 L) { return L + 1 >= I && (L + 1 - I) % s != 0; };
    auto filtered = [&](const bsl_batch *q, u32 i, u32 L) {
        if (L == 0 || L < P.min_read_size || L > BSL_MAX_READLEN) return true;
        const u8 *p = q->bases + q->offsets[i]; u32 ns = 0;
        for (u32 k = 0; k < L; k++) { const u8 c = p[k] & 0xDFu; ns += !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }
        return ns > P.max_ns;
    };
    for (u32 i = upto; i-- > 0;) {
        bool want = false, settled = true;
        for (int m = 0; m < nm; m++) {
            const bsl_batch *q = m ? b : a; const u32 L = (u32)std::min<u64>(q->offsets[i + 1] - q->offsets[i], 0xffffu);
            const bool cand = L > maxlen[m] || (!have_def[m] && defining(L));
            if (cand && !filtered(q, i, L)) { if (L > maxlen[m]) maxlen[m] = L; if (defining(L)) have_def[m] = true; want = true; }
            if (!have_def[m] || maxlen[m] < BSL_MAX_READLEN) settled = false;
        }
        if (want) idx.push_back(i);
        if (settled) break;
    }
    std::reverse(idx.begin(), idx.end());
}

// Opt-in shared-memory sizes are per device: done once per context (under its first lane's lock or any later one: idempotent).
static int configure_kernels(bsl_ctx *ctx) {
    std::lock_guard<std::mutex> g(ctx->stats_mu);
    if (ctx->kernels_configured) return 0;
    CUDA_TRY(cudaFuncSetAttribute(prepare_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(prepare_deferred, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(screen_bits<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(screen_bits<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(screen_bits<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(screen_bits<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(pair_round_wide<PW_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairWideSmem)));
    CUDA_TRY(cudaFuncSetAttribute(reduce_round<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(reduce_round<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(verify_candidates<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(verify_candidates<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(verify_candidates<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(verify_candidates<false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    ctx->kernels_configured = true;
    return 0;
}

// One sub-range [first, first+n_a) of the caller's batch on one lane. Sub-ranges keep the number of items a search
// round can produce below MAX_ITEMS_PER_ROUND and bound the device memory of a call.
struct LenScan { u32 Lmax = 0; u32 dims[2] = {0, 1}; bool empty_range = false; };       // one pass over the read lengths of a (sub-)range
static LenScan scan_lengths(const bsl_batch *a, const bsl_batch *b, u32 first, u32 n, const bsl_params &P) {
    LenScan r; r.dims[0] = P.index_interval;
    r.Lmax = max_len_of(a, first, n, &P, r.dims, &r.empty_range);
    if (b) r.Lmax = std::max(r.Lmax, max_len_of(b, first, n, &P, r.dims, &r.empty_range));
    return r;
}

static int align_range(bsl_ctx *ctx, Lane &ln, const bsl_batch *a, const bsl_batch *b, u32 first, u32 n_a, bsl_hit *out_a, bsl_hit *out_b, bsl_pair *out_pair,
                       bsl_hit *all_a, bsl_hit *all_b, u64 all_cap, u64 all_off, u64 *all_made, int resident, bsl_stats *acc, bool carry, const LenScan *scan = nullptr) {
    int rc = 0;
    const bool pe = b != nullptr; const u32 n_slots = pe ? 2 * n_a : n_a;
    cudaStream_t st = ln.stream;
    const bsl_params &P = ctx->P;

    const u64 off0_a = a->offsets[first], off0_b = pe ? b->offsets[first] : 0;
    const u64 bases_a = a->offsets[first + n_a] - off0_a, bases_b = pe ? b->offsets[first + n_a] - off0_b : 0;
    const LenScan ls = scan ? *scan : scan_lengths(a, b, first, n_a, P);
    const u32 pdims[2] = {ls.dims[0], ls.dims[1]};
    u32 Lmax = ls.Lmax;
    if (carry) {            // st0 per (slot, chain), the deferred list, and 16 inherited seed hashes per (slot, chain)
        size_t c1 = ln.cap_st0; if ((rc = grow(ctx, &ln.d_st0, &c1, (size_t)n_slots * 2 + 16))) return rc; ln.cap_st0 = c1;
        c1 = ln.cap_defer; if ((rc = grow(ctx, &ln.d_defer, &c1, (size_t)n_slots + 16))) return rc; ln.cap_defer = c1;
        c1 = ln.cap_stale; if ((rc = grow(ctx, &ln.d_stale, &c1, (size_t)n_slots * 32 + 16))) return rc; ln.cap_stale = c1;
    }
    if (Lmax > BSL_MAX_READLEN) Lmax = BSL_MAX_READLEN;       // longer reads are flagged filtered by prepare_reads; the CLI truncates like the reference
    const u32 Wb = std::max(1u, (Lmax + 31) / 32);
    const u32 cap = 32;
    const char *env_cap = getenv("BSL_HIT_CAP"); const u32 cap_main = env_cap ? std::max(2, atoi(env_cap)) : cap;

    // ---- buffers
    size_t c0;
    c0 = ln.cap_bases; if ((rc = grow(ctx, &ln.d_bases, &c0, (size_t)(bases_a + bases_b + 2 * BASES_PAD)))) return rc; ln.cap_bases = c0;
    {
        size_t need = n_slots + 2;
        if (need > ln.cap_slots) {
            size_t ncap = std::max(need, ln.cap_slots + ln.cap_slots / 2);
            // free-and-null first: a failed allocation below leaves the lane with null pointers and zero capacity, so the next call starts clean
            auto drop = [](auto *&q) { cudaFree(q); q = nullptr; };
            drop(ln.d_off); drop(ln.d_index); drop(ln.d_rawlen); drop(ln.d_meta); drop(ln.d_cnt); drop(ln.d_sched); drop(ln.d_stat);
            drop(ln.d_minlvl); drop(ln.d_slot_item); drop(ln.d_slot_flag); drop(ln.d_flag_list); drop(ln.d_marks);
            drop(ln.d_list[0]); drop(ln.d_list[1]); drop(ln.d_pe_list[0]); drop(ln.d_pe_list[1]); drop(ln.d_heavy_list); drop(ln.d_out); drop(ln.d_pair);
            ln.cap_slots = 0;
            CUDA_TRY(cudaMalloc(&ln.d_off, (ncap + 4) * 8)); CUDA_TRY(cudaMalloc(&ln.d_index, ncap * 4)); CUDA_TRY(cudaMalloc(&ln.d_rawlen, ncap * 2));
            CUDA_TRY(cudaMalloc(&ln.d_meta, ncap * sizeof(SlotMeta))); CUDA_TRY(cudaMalloc(&ln.d_cnt, ncap * sizeof(SlotCounts))); CUDA_TRY(cudaMalloc(&ln.d_sched, ncap * 32)); CUDA_TRY(cudaMalloc(&ln.d_stat, ncap * sizeof(uint2)));
            CUDA_TRY(cudaMalloc(&ln.d_minlvl, ncap)); CUDA_TRY(cudaMalloc(&ln.d_slot_item, ncap * 2 * sizeof(uint2))); CUDA_TRY(cudaMalloc(&ln.d_slot_flag, ncap * 4)); CUDA_TRY(cudaMalloc(&ln.d_flag_list, ncap * 4)); CUDA_TRY(cudaMalloc(&ln.d_marks, ncap * MK_CAP * sizeof(uint4)));
            for (int k = 0; k < 2; k++) { CUDA_TRY(cudaMalloc(&ln.d_list[k], ncap * 4)); CUDA_TRY(cudaMalloc(&ln.d_pe_list[k], ncap * 4)); }
            CUDA_TRY(cudaMalloc(&ln.d_heavy_list, ncap * 4)); CUDA_TRY(cudaMalloc(&ln.d_out, ncap * sizeof(bsl_hit))); CUDA_TRY(cudaMalloc(&ln.d_pair, ncap * sizeof(bsl_pair)));
            ln.cap_slots = ncap;
        }
    }
    c0 = ln.cap_words; if ((rc = grow(ctx, &ln.d_planes, &c0, (size_t)n_slots * 6 * Wb + 16))) return rc; ln.cap_words = c0;
    c0 = ln.cap_bits1; if ((rc = grow(ctx, &ln.d_bits1, &c0, (size_t)n_slots * 8 * Wb + 16))) return rc; ln.cap_bits1 = c0;
    c0 = ln.cap_hits; if ((rc = grow(ctx, &ln.d_hits, &c0, (size_t)n_slots * cap_main))) return rc; ln.cap_hits = c0;
    // large blocks for lists that outgrow cap_main (reads from repeat families: every level may fill up to -w): room for about
    // 1.5 % of the slots, at most 1 GB; a read that finds none left is re-run on the large-capacity pass
    const u32 big_cap = std::min<u32>(16 * P.max_num_hits + 32, 65535u);
    const u32 big_blocks = (u32)std::min<u64>(std::max<u64>(n_slots / 64, 256), std::max<u64>((1ull << 30) / ((u64)big_cap * sizeof(DevHit)), 16));
    c0 = ln.cap_bighits; if ((rc = grow(ctx, &ln.d_bighits, &c0, (size_t)big_blocks * big_cap))) return rc; ln.cap_bighits = c0;
    c0 = ln.cap_wide; if ((rc = grow(ctx, &ln.d_wide_list, &c0, (size_t)n_a + 16))) return rc; ln.cap_wide = c0;
    // per-round scratch: flat candidate space (1 bit per candidate) and item headers
    const u32 nch = P.chains == 1 ? 2 : 1;
    const u64 worst_slot = 2ull * P.index_interval * std::max<u32>(ctx->di.maxk, 1);      // candidates one read can have in one round
    const char *env_cc = getenv("BSL_CAND_CAP");
    // flat candidate space of one round: sized for ~1.5x the expected bucket walks of every slot (mean bucket occupancy
    // n_entries / K per look-up; 96 per slot at 500 Mb, ~430 at 3.1 Gb); reads that do not fit are re-run on the large-capacity pass
    const u64 per_lookup = ctx->di.K ? ctx->di.n_entries / ctx->di.K + 1 : 1;
    u64 want_cands = env_cc ? (u64)atoll(env_cc) : std::max<u64>(std::max<u64>(96ull * n_slots, 3ull * n_slots * P.index_interval * per_lookup / 2), 1ull << 22);
    want_cands = std::max<u64>(want_cands, 2 * worst_slot + CHUNK);
    want_cands = std::min<u64>(want_cands, 0xfff00000ull);
    want_cands = (want_cands + CHUNK - 1) / CHUNK * CHUNK;
    if (2 * worst_slot + CHUNK > want_cands) { set_error(ctx, "over-represented k-mer cut-off %u is too large for the 32-bit candidate space", ctx->di.maxk); return BSL_ELIMIT; }
    const u64 want_items = std::min<u64>((u64)n_slots * nch * P.index_interval, want_cands) + 16;
    c0 = ln.cap_bitmap; if ((rc = grow(ctx, &ln.d_bitmap, &c0, (size_t)(want_cands / 32 + 8)))) return rc;
    if (c0 != ln.cap_bitmap || !ln.d_chunk_first || !ln.d_flat_loc) {
        cudaFree(ln.d_chunk_first); ln.d_chunk_first = nullptr; cudaFree(ln.d_flat_loc); ln.d_flat_loc = nullptr; ln.cap_bitmap = 0;
        CUDA_TRY(cudaMalloc(&ln.d_chunk_first, (c0 + 8) * 4)); CUDA_TRY(cudaMalloc(&ln.d_flat_loc, (c0 * 32 + 2 * CHUNK) * 4)); }
    ln.cap_bitmap = c0;
    c0 = ln.cap_items; if ((rc = grow(ctx, &ln.d_hdr, &c0, (size_t)want_items))) return rc; ln.cap_items = c0;
    const bool want_all = P.report_repeat_hits == 2 && all_a && (!pe || all_b) && all_cap > 0;
    const u64 sub_cap = all_cap > all_off ? all_cap - all_off : 0;      // room left in the caller's all-hits arrays
    if (want_all) {
        const u64 dcap = std::max<u64>(sub_cap, 1);
        if (dcap > ln.cap_all) { cudaFree(ln.d_all[0]); cudaFree(ln.d_all[1]); ln.d_all[0] = ln.d_all[1] = nullptr; ln.cap_all = 0;
            CUDA_TRY(cudaMalloc(&ln.d_all[0], dcap * sizeof(bsl_hit))); CUDA_TRY(cudaMalloc(&ln.d_all[1], dcap * sizeof(bsl_hit))); ln.cap_all = dcap; }
    }

    // ---- kernel arguments
    KArgs A; memset(&A, 0, sizeof A);
    A.di = ctx->di; A.tab = (const DevTables *)ctx->d_tables;
    A.s = P.seed_size; A.I = P.index_interval; A.gap = P.gap; A.w = P.max_num_hits; A.min_insert = P.min_insert; A.max_insert = P.max_insert;
    A.chains = P.chains; A.report = P.report_repeat_hits; A.randseed = P.randseed; A.max_ns = P.max_ns; A.min_read_size = P.min_read_size; A.single = ctx->rule.single;
    A.n_slots = n_slots; A.n_a = n_a; A.pe = pe; A.readset_a = a->readset; A.readset_b = pe ? b->readset : 0;
    A.bases = ln.d_bases + BASES_PAD; A.off = ln.d_off; A.index = ln.d_index; A.rawlen = ln.d_rawlen; A.first_index_a = a->first_index + first; A.first_index_b = pe ? b->first_index + first : 0;
    A.off_base_a = off0_a; A.off_base_b = off0_b; A.all_off = all_off;
    A.bases_b_shift = bases_a; A.has_index = (a->index != nullptr) && (!pe || b->index != nullptr); A.has_rawlen = (a->raw_len != nullptr) && (!pe || b->raw_len != nullptr);
    A.n_ctx = first == 0 ? std::min(a->n_context, n_a) : 0; A.carry = carry ? 1u : 0u; A.st0arr = ln.d_st0; A.defer = ln.d_defer; A.stale = ln.d_stale;
    A.Wb = Wb; A.planes = ln.d_planes; A.bits1 = ln.d_bits1; A.sched = ln.d_sched; A.meta = ln.d_meta; A.cnt = ln.d_cnt; A.stat = ln.d_stat; A.minlvl = ln.d_minlvl;
    A.hits = ln.d_hits; A.cap = cap_main; A.ctr = ln.d_ctr;
    A.bighits = ln.d_bighits; A.big_cap = big_cap; A.big_blocks = big_blocks; A.wide_list = ln.d_wide_list;
    A.hdr = ln.d_hdr; A.cap_items = (u32)std::min<u64>(want_items - 16, 0xffffffffu); A.cap_cands = (u32)want_cands;
    A.slot_item = ln.d_slot_item; A.marks = ln.d_marks; A.slot_flag = ln.d_slot_flag; A.flag_list = ln.d_flag_list; A.chunk_first = ln.d_chunk_first; A.bitmap = ln.d_bitmap; A.flat_loc = ln.d_flat_loc;
    A.out = ln.d_out; A.pair_out = ln.d_pair; A.all_a = want_all ? ln.d_all[0] : nullptr; A.all_b = want_all ? ln.d_all[1] : nullptr; A.all_cap = want_all ? sub_cap : 0;

    // ---- H2D
    CUDA_TRY(cudaEventRecord(ln.ev[0], st));
    CUDA_TRY(cudaMemsetAsync(ln.d_ctr, 0, sizeof(DevCounters), st));
    if (!resident) {
    CUDA_TRY(cudaMemcpyAsync(ln.d_bases + BASES_PAD, a->bases + off0_a, bases_a, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ln.d_off, a->offsets + first, (size_t)(n_a + 1) * 8, cudaMemcpyHostToDevice, st));
    if (pe) {
        CUDA_TRY(cudaMemcpyAsync(ln.d_bases + BASES_PAD + bases_a, b->bases + off0_b, bases_b, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ln.d_off + (n_a + 1), b->offsets + first, (size_t)(n_a + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    }
    if (pe) CUDA_TRY(cudaMemsetAsync(ln.d_pair, 0, (size_t)n_a * sizeof(bsl_pair), st));
    if (A.has_index && !resident) { CUDA_TRY(cudaMemcpyAsync(ln.d_index, a->index + first, (size_t)n_a * 4, cudaMemcpyHostToDevice, st));
        if (pe) CUDA_TRY(cudaMemcpyAsync(ln.d_index + n_a, b->index + first, (size_t)n_a * 4, cudaMemcpyHostToDevice, st)); }
    if (A.has_rawlen && !resident) { CUDA_TRY(cudaMemcpyAsync(ln.d_rawlen, a->raw_len + first, (size_t)n_a * 2, cudaMemcpyHostToDevice, st));
        if (pe) CUDA_TRY(cudaMemcpyAsync(ln.d_rawlen + n_a, b->raw_len + first, (size_t)n_a * 2, cudaMemcpyHostToDevice, st)); }

    u64 launches = 0;
    const int sms = ctx->sm_count;
    CUDA_TRY(cudaEventRecord(ln.ev[1], st));
    {
        // prepare_reads keeps WQ + Wb + 1 + wd + nseg (ii + 1) words per read in shared memory (<= 200 KB per CTA for any -s / -I / length)
        const u32 WQ = 2 * Wb + 1;
        // -I 4, -s a multiple of 4, at most 8 offsets per segment: the sizes of a segment never leave the registers (no cw[] in the row;
        // 53 words = 8 CTAs per SM at 2x150 bp)
        const u32 wdm = (P.index_interval == 4 && (P.seed_size & 3u) == 0 && pdims[0] <= 8) ? 0u : pdims[0];
        const size_t smem_p = (size_t)PR_WARPS * 32 * ((WQ + Wb + 1 + wdm + pdims[1]) | 1u) * 4;
        if (smem_p > 200 * 1024) { set_error(ctx, "seed schedule does not fit the shared memory of prepare_reads"); return BSL_ELIMIT; }
        if (ctx->prep_smem != smem_p) {
            // The kernel wants L1 more than resident warps (its loads of the read bytes, of the size table and its stores all pass
            // through it): ask for about 132 KB of shared memory — 8 CTAs per SM at 100 bp, 4 at 2x150 bp — and leave the rest of
            // the 256 KB to L1. Measured at 2x150 bp: 1.88 ms with room for 8 or 7 CTAs (28 KB of L1), 1.26 ms for 6, 1.15 for 5,
            // 1.13 for 4, 1.31 for 3 (profiles/README.md).
            const size_t per_cta = smem_p + 1536, total = 228 * 1024;
            const size_t ctas = std::min<size_t>(8, std::max<size_t>(3, (132 * 1024) / per_cta));
            CUDA_TRY(cudaFuncSetAttribute(prepare_reads, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (ctas * per_cta * 100 + total - 1) / total)));
            ctx->prep_smem = smem_p;
        }
        const u32 groups = (n_slots + 31) / 32;
        if (carry) CUDA_TRY(cudaMemsetAsync(ln.d_st0, 0, (size_t)n_slots * 2, st));
        prepare_reads<<<std::min<u32>((groups + PR_WARPS - 1) / PR_WARPS, (u32)sms * 16), PR_WARPS * 32, smem_p, st>>>(A, WQ, wdm, pdims[1]);
        launches++;
        if (carry) {        // reads with an empty start-offset range: schedule from the state the earlier reads of the batch leave behind
            const u32 wdd = P.index_interval + P.seed_size, csz = 16 * P.seed_size;
            prepare_deferred<<<sms * 4, PD_WARPS * 32, (size_t)PD_WARPS * 32 * ((wdd + csz) | 1u) * 4, st>>>(A, wdd, csz);
            launches++;
        }
    }
    build_lists<<<(std::max(n_slots, n_a) + 255) / 256, 256, 0, st>>>(A, ln.d_list[0], ln.d_pe_list[0]); launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ln.ev[2], st));

    const u32 G = P.gap;
    const u32 NW = (31 + 2 * G + Lmax + 31) / 32; const u32 NWS = NW | 1u;
    const size_t smem_r = (size_t)ROUND_WARPS * (32 * NWS + 48 + RR_KEYS) * 8;
    const u32 NP = ctx->rule.single ? 2 : 3;
    const u32 NPL = NP + 1;                                               // verify_candidates (-g) stages a prefix-mask stream as well
    const u32 Wr = (Lmax + 31) / 32;                                     // 64-bit words of the longest read
    const bool ns3 = Wr + 1 + 3 <= 12;                                   // a window (Wr + 1 words at any of 4 word offsets) fits 3 sectors
    const size_t smem_v = (size_t)VF_ITMAX * NPL * 2 * Wb * 4;
    const u32 rounds_se = std::min<u32>((Lmax + 1 >= P.index_interval + P.seed_size) ? (Lmax + 1 - P.index_interval) / P.seed_size : 0, 16);
    const int vkind = ctx->rule.single ? 0 : 1;
    if (G && !ctx->occ_verify[vkind]) {
        int occ = 0;                                     // sized for the common 3-sector variant
        cudaError_t oe = ctx->rule.single ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, verify_candidates<true, 3>, VF_THREADS, smem_v)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, verify_candidates<false, 3>, VF_THREADS, smem_v);
        ctx->occ_verify[vkind] = (oe == cudaSuccess && occ > 0) ? occ : 3;
    }
    const int grid_l = sms * 8, grid_v = sms * std::max(ctx->occ_verify[vkind], 1), grid_r = sms * 4;      // persistent grids: a multiple of the SM count

    int nev = 0; std::vector<char> ev_kind;           // per-launch CUDA-event pairs: 'l' seed_lookup, 'v' verify, 'r' reduce, 'p' pair_round
    const int max_ev = (int)(sizeof ln.evk / sizeof ln.evk[0]);
    auto ev_begin = [&](char kind) { if (nev + 2 <= max_ev) { cudaEventRecord(ln.evk[nev], st); ev_kind.push_back(kind); } };
    auto ev_end = [&]() { if (nev + 2 <= max_ev) { cudaEventRecord(ln.evk[nev + 1], st); nev += 2; } };
    const u32 DWv = NP * Wb;                                             // 64-bit words of one item's read streams
    const u32 stage_items = std::max<u32>(1, std::min<u32>(32, 640 / (2 * DWv)));
    const size_t smem_s = (size_t)SC_WARPS * stage_items * 2 * DWv * 4;
    const u32 rcp_dw = (u32)((0x100000000ull + DWv - 1) / DWv);
    if (!ctx->occ_screen) {
        int occ = 0;
        cudaError_t oe = ctx->rule.single ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_candidates<true>, SC_WARPS * 32, smem_s)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_candidates<false>, SC_WARPS * 32, smem_s);
        ctx->occ_screen = (oe == cudaSuccess && occ > 0) ? occ : 4;
    }
    const int grid_s = sms * ctx->occ_screen;
    // 1-bit screen (single-conversion rules)
    const bool use_bits = ctx->di.has_bit1 != 0;
    const u32 item_bytes = 16 * (Wb | 1u);                                // staged item: {low bits, ACGT mask} x Wb forward, x Wb reversed; stride = 4 x odd words spreads the items over the banks
    const u32 rcp_wb = (u32)((0x100000000ull + Wb - 1) / Wb);
    const size_t smem_b = (size_t)SB_WARPS * 32 * item_bytes;
    if (!ctx->occ_bits || ctx->occ_bits_wb != Wb) {
        int occ = 0; cudaError_t oe;
        if (G) oe = ctx->rule.single ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_bits<true, true>, SB_WARPS * 32, smem_b) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_bits<false, true>, SB_WARPS * 32, smem_b);
        else oe = ctx->rule.single ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_bits<true, false>, SB_WARPS * 32, smem_b) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, screen_bits<false, false>, SB_WARPS * 32, smem_b);
        ctx->occ_bits = (oe == cudaSuccess && occ > 0) ? occ : 1; ctx->occ_bits_wb = Wb;
        // shared memory for exactly the CTAs that fit, the rest of the SM's 256 KB stays L1 (with 28 KB of L1 the kernel takes 2.4 times as long)
        const size_t need = (size_t)ctx->occ_bits * (smem_b + sizeof(uint4) * SB_WARPS * SC_QCAP + 1024), total = 228 * 1024;
        const int pct = (int)std::min<size_t>(100, (need * 100 + total - 1) / total);
        if (G) { if (ctx->rule.single) cudaFuncSetAttribute(screen_bits<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); else cudaFuncSetAttribute(screen_bits<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); }
        else { if (ctx->rule.single) cudaFuncSetAttribute(screen_bits<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); else cudaFuncSetAttribute(screen_bits<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct); }
    }
    const int grid_b = sms * ctx->occ_bits;
    auto launch_verify = [&](KArgs &K, u32 ci) {
        if (use_bits) {                                      // single-conversion and '-'-only rules, with or without -g
            if (ctx->rule.single) { if (G) screen_bits<true, true><<<grid_b, SB_WARPS * 32, smem_b, st>>>(K, ci, item_bytes, rcp_wb); else screen_bits<true, false><<<grid_b, SB_WARPS * 32, smem_b, st>>>(K, ci, item_bytes, rcp_wb); }
            else { if (G) screen_bits<false, true><<<grid_b, SB_WARPS * 32, smem_b, st>>>(K, ci, item_bytes, rcp_wb); else screen_bits<false, false><<<grid_b, SB_WARPS * 32, smem_b, st>>>(K, ci, item_bytes, rcp_wb); }
            return;
        }
        if (!G) {
            if (ctx->rule.single) screen_candidates<true><<<grid_s, SC_WARPS * 32, smem_s, st>>>(K, ci, stage_items, rcp_dw);
            else screen_candidates<false><<<grid_s, SC_WARPS * 32, smem_s, st>>>(K, ci, stage_items, rcp_dw);
            return;
        }
        if (ctx->rule.single) { if (ns3) verify_candidates<true, 3><<<grid_v, VF_THREADS, smem_v, st>>>(K, ci, Wr); else verify_candidates<true, 5><<<grid_v, VF_THREADS, smem_v, st>>>(K, ci, Wr); }
        else { if (ns3) verify_candidates<false, 3><<<grid_v, VF_THREADS, smem_v, st>>>(K, ci, Wr); else verify_candidates<false, 5><<<grid_v, VF_THREADS, smem_v, st>>>(K, ci, Wr); }
    };
    auto search = [&](KArgs &K, bool as_pe, u32 r, const u32 *lin, u32 *lout, u32 ci) {
        ev_begin('l');
        if (as_pe) seed_lookup<true><<<grid_l, LK_THREADS, 0, st>>>(K, r, lin, lout, ci);
        else seed_lookup<false><<<grid_l, LK_THREADS, 0, st>>>(K, r, lin, lout, ci);
        ev_end();
        ev_begin('v');
        launch_verify(K, ci);
        ev_end();
        ev_begin('r');
        {
            const u32 *rl = as_pe ? lin : lout; const u32 rci = as_pe ? ci : ci + 1;
            reduce_fast<<<sms * 8, 256, 0, st>>>(K, rl, rci, ci, as_pe ? 1u : 0u, G ? 0u : 1u);      // -g: lists the reads, replays none
            if (ctx->rule.single) reduce_round<true><<<grid_r, ROUND_WARPS * 32, smem_r, st>>>(K, r, ci, NW, NWS);
            else reduce_round<false><<<grid_r, ROUND_WARPS * 32, smem_r, st>>>(K, r, ci, NW, NWS);
        }
        ev_end();
        launches += 4;
    };
    auto run_passes = [&](KArgs &K) -> int {
        // SE rounds (stop rule evaluated by the next round's seed_lookup)
        if (!K.pe) for (u32 r = 0; r < rounds_se; r++) search(K, false, r, ln.d_list[r & 1], ln.d_list[(r + 1) & 1], r);
        else {
            // paired rounds; pairs with a filtered mate ride along (their lone mate follows SingleAlign's stop rule, see seed_lookup)
            for (u32 r = 0; r <= BSL_MAXSNPS; r++) {
                if (r < rounds_se) search(K, true, r, ln.d_pe_list[r & 1], nullptr, 20 + r);
                // after the last search round nothing changes between the pairing rounds: one launch replays all that are left
                const u32 r_hi = r + 1 >= rounds_se ? BSL_MAXSNPS : r;
                ev_begin('p');
                if (K.hits == ln.d_heavy_hits) { pair_round_wide<PW_CAP><<<sms * 4, 32, sizeof(PairWideSmem), st>>>(K, r, r_hi, ln.d_pe_list[r & 1], ln.d_pe_list[(r + 1) & 1], 20 + r, 0u); launches++; }
                else {
                    pair_round<<<sms * 8, 128, 0, st>>>(K, r, r_hi, ln.d_pe_list[r & 1], ln.d_pe_list[(r + 1) & 1], 20 + r);
                    // the pairs with long lists: first the few with a list beyond PW_CAP_SMALL (they take longest), then the rest
                    pair_round_wide<PW_CAP><<<sms * 4, 32, sizeof(PairWideSmem), st>>>(K, r, r_hi, K.wide_list, ln.d_pe_list[(r + 1) & 1], 20 + r, 2u);
                    pair_round_wide<PW_CAP_SMALL><<<sms * 24, 32, sizeof(PairWideSmemT<PW_CAP_SMALL>), st>>>(K, r, r_hi, K.wide_list, ln.d_pe_list[(r + 1) & 1], 20 + r, 1u);
                    launches += 3;
                }
                ev_end();
                if (r_hi != r) break;
            }
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error(ctx, "kernel launch failed: %s", cudaGetErrorString(e)); return BSL_ECUDA; }
        return 0;
    };
    if (smem_v > 128 * 1024) { set_error(ctx, "read planes do not fit the verify staging buffer"); return BSL_ELIMIT; }
    if ((rc = run_passes(A))) return rc;
    CUDA_TRY(cudaEventRecord(ln.ev[3], st));
    finalize_reads<<<(n_slots + 255) / 256, 256, 0, st>>>(A, nullptr, 0); launches++;
    CUDA_TRY(cudaMemcpyAsync(ln.h_ctr, ln.d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(ln.ev[4], st));
    // ---- D2H: the records leave right behind the kernels; only a call that needs the large-capacity pass (rare: a read found no
    //      large block left, or the candidate space of a round ran out) synchronises twice and copies them again
    auto copy_out = [&]() -> int {
        if (resident) return 0;
        CUDA_TRY(cudaMemcpyAsync(out_a + first, ln.d_out, (size_t)n_a * sizeof(bsl_hit), cudaMemcpyDeviceToHost, st));
        if (pe) { CUDA_TRY(cudaMemcpyAsync(out_b + first, ln.d_out + n_a, (size_t)n_a * sizeof(bsl_hit), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(out_pair + first, ln.d_pair, (size_t)n_a * sizeof(bsl_pair), cudaMemcpyDeviceToHost, st)); }
        return 0;
    };
    if ((rc = copy_out())) return rc;
    CUDA_TRY(cudaEventRecord(ln.ev[5], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    DevCounters c1 = *ln.h_ctr;
    u64 heavy_total = 0;
    if (c1.overflow_n > 0) {
        // ---- heavy pass: re-run the overflowed reads / pairs from scratch with room for every list (16 levels x -w)
        u32 *d_nheavy = &ln.d_ctr->overflow_n;
        CUDA_TRY(cudaMemsetAsync(d_nheavy, 0, 4, st));
        const u32 n_items = pe ? n_a : n_slots;
        collect_overflow<<<(n_items + 255) / 256, 256, 0, st>>>(A, ln.d_heavy_list, d_nheavy); launches++;
        u32 n_heavy = 0; CUDA_TRY(cudaMemcpyAsync(&n_heavy, d_nheavy, 4, cudaMemcpyDeviceToHost, st)); CUDA_TRY(cudaStreamSynchronize(st));
        heavy_total = n_heavy;
        const u32 hcap = std::min<u32>(16 * P.max_num_hits + 32, 65535u);
        const u32 per_item = pe ? 2 : 1;
        u32 chunk = (u32)std::max<u64>(1, (1ull << 30) / ((u64)hcap * 16 * per_item));
        chunk = (u32)std::min<u64>(chunk, std::max<u64>(1, (want_cands - CHUNK) / (worst_slot * per_item)));     // the candidate space must hold every read of the chunk
        chunk = std::min(chunk, n_heavy);
        c0 = ln.cap_heavy_hits; if ((rc = grow(ctx, &ln.d_heavy_hits, &c0, (size_t)chunk * per_item * hcap))) return rc; ln.cap_heavy_hits = c0;
        KArgs H = A; H.hits = ln.d_heavy_hits; H.cap = hcap;
        u32 *d_slots = ln.d_list[0];
        for (u32 first = 0; first < n_heavy; first += chunk) {
            const u32 count = std::min(chunk, n_heavy - first);
            CUDA_TRY(cudaMemsetAsync(ln.d_ctr->rc, 0, sizeof(ln.d_ctr->rc), st));
            reset_heavy<<<(count + 127) / 128, 128, 0, st>>>(H, ln.d_heavy_list, first, count, ln.d_list[0], ln.d_pe_list[0]); launches++;
            if ((rc = run_passes(H))) return rc;
            d_slots = ln.d_list[0];
            expand_pairs_to_slots<<<(count + 127) / 128, 128, 0, st>>>(ln.d_heavy_list, first, count, n_a, pe, d_slots); launches++;
            finalize_reads<<<(count * per_item + 255) / 256, 256, 0, st>>>(H, d_slots, count * per_item); launches++;
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpyAsync(ln.h_ctr, ln.d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(ln.ev[4], st));
        if ((rc = copy_out())) return rc;
        CUDA_TRY(cudaEventRecord(ln.ev[5], st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    DevCounters c2 = *ln.h_ctr;
    if (want_all && !resident) {
        u64 na_ = std::min<u64>(c2.all_n, sub_cap);
        if (na_) { CUDA_TRY(cudaMemcpy(all_a + all_off, ln.d_all[0], na_ * sizeof(bsl_hit), cudaMemcpyDeviceToHost));
            if (pe) CUDA_TRY(cudaMemcpy(all_b + all_off, ln.d_all[1], na_ * sizeof(bsl_hit), cudaMemcpyDeviceToHost)); }
    }
    if (all_made) *all_made = c2.all_n;
    float ms_pack = 0, ms_total = 0, ms_dev = 0; double ms_l = 0, ms_v = 0, ms_r = 0, ms_p = 0; u64 n_v = 0;
    cudaEventElapsedTime(&ms_pack, ln.ev[1], ln.ev[2]); cudaEventElapsedTime(&ms_total, ln.ev[0], ln.ev[5]); cudaEventElapsedTime(&ms_dev, ln.ev[1], ln.ev[4]);
    for (int k = 0; k + 1 < nev; k += 2) {
        float t = 0; cudaEventElapsedTime(&t, ln.evk[k], ln.evk[k + 1]);
        switch (ev_kind[k / 2]) { case 'l': ms_l += t; break; case 'v': ms_v += t; n_v++; break; case 'r': ms_r += t; break; default: ms_p += t; }
    }
    {
        bsl_stats &S = *acc;
        S.reads += n_slots; S.seed_lookups += c2.seed_lookups; S.candidates += c2.candidates; S.hits_added += c2.hits_added; S.heavy_reads += heavy_total;
        S.ms_pack += ms_pack; S.ms_search += ms_l + ms_v + ms_r; S.ms_pair += ms_p; S.ms_total += ms_total; S.ms_device += ms_dev; S.kernel_launches += launches; S.search_launches += n_v;
        S.ms_lookup += ms_l; S.ms_verify += ms_v; S.ms_reduce += ms_r;
        const u32 Wd = (Lmax + 31) / 32 + 1 + (G ? 1 : 0);      // 4 + 8 W bytes per candidate, W+1 words with -g
        S.verify_bytes += c2.candidates * (4 + 8ull * Wd);
    }
    return 0;
}

int bsl_align_impl(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b, bsl_hit *out_a, bsl_hit *out_b, bsl_pair *out_pair,
                   bsl_hit *all_a, bsl_hit *all_b, u64 all_cap, u64 *n_all, int resident) {
    if (!ctx->has_index) { set_error(ctx, "bsl_align: no index (call bsl_index_build first)"); return BSL_ESTATE; }
    if (!a || (!resident && (!out_a || (b && (!out_b || !out_pair))))) { set_error(ctx, "bsl_align: null argument"); return BSL_EINVAL; }
    if (b && a->n != b->n) { set_error(ctx, "bsl_align_pe: batches differ in size (%u vs %u)", a->n, b->n); return BSL_EINVAL; }
    if (n_all) *n_all = 0;
    const u32 n = a->n; const bool pe = b != nullptr;
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // pick a lane
    Lane *lnp = nullptr; std::unique_lock<std::mutex> lk;
    for (int t = 0; t < BSL_NLANES && !lnp; t++) { std::unique_lock<std::mutex> l(ctx->lanes[t].mu, std::try_to_lock); if (l.owns_lock()) { lk = std::move(l); lnp = &ctx->lanes[t]; } }
    if (!lnp) { lk = std::unique_lock<std::mutex>(ctx->lanes[0].mu); lnp = &ctx->lanes[0]; }
    Lane &ln = *lnp;
    int rc = ensure_lane(ctx, ln); if (rc) return rc;
    if ((rc = configure_kernels(ctx))) return rc;
    if (ctx->di.maxk >= BSL_MAX_BUCKET) { set_error(ctx, "over-represented k-mer cut-off %u exceeds the largest bucket a walk can visit (%u)", ctx->di.maxk, BSL_MAX_BUCKET - 1); return BSL_ELIMIT; }
    // sub-ranges: a search round may create at most MAX_ITEMS_PER_ROUND items (slots x enabled chains x -I look-ups)
    const bsl_params &P = ctx->P;
    const u32 per_read = (P.chains == 1 ? 2u : 1u) * P.index_interval * (pe ? 2u : 1u);
    u32 sub = std::min<u32>(std::max<u32>(1024u, MAX_ITEMS_PER_ROUND / per_read), BSL_MAX_SLOTS / (pe ? 2u : 1u));
    const char *env_sub = getenv("BSL_SUB_BATCH"); if (env_sub && atoi(env_sub) > 0) sub = (u32)atoi(env_sub);
    if (resident && n > sub) { set_error(ctx, "bsl_align_rerun: the batch was split into sub-ranges and is not resident"); return BSL_ESTATE; }
    bsl_stats acc; memset(&acc, 0, sizeof acc);
    u64 all_off = 0;
    const LenScan whole = scan_lengths(a, b, 0, n, P);                // one pass over the lengths; reused when the call is one sub-range
    const bool carry = whole.empty_range;                           // any read with an empty start-offset range in this call?
    for (u32 first = 0; first < n; first += sub) {
        const u32 cnt = std::min(sub, n - first);
        u64 made = 0;
        std::vector<u32> cidx;
        if (carry && first > 0) carry_context(P, a, b, first, cidx);
        if (cidx.empty()) rc = align_range(ctx, ln, a, b, first, cnt, out_a, out_b, out_pair, all_a, all_b, all_cap, all_off, &made, resident, &acc, carry, n <= sub ? &whole : nullptr);
        else {
            // a later sub-range in carry mode: the reads that define its inherited state go in front of it as context
            if (resident) { set_error(ctx, "bsl_align_rerun: the batch was split into sub-ranges and is not resident"); return BSL_ESTATE; }
            const u32 nc = (u32)cidx.size(), nt = nc + cnt;
            struct Tmp { std::vector<u8> bases; std::vector<u64> off; std::vector<u32> index; std::vector<u16> raw; bsl_batch q; std::vector<bsl_hit> out; } t[2];
            std::vector<bsl_pair> tpair(pe ? nt : 0);
            for (int m = 0; m < (pe ? 2 : 1); m++) {
                const bsl_batch *q = m ? b : a; Tmp &T = t[m];
                T.off.resize((size_t)nt + 1); T.index.resize(nt); if (q->raw_len) T.raw.resize(nt);
                u64 tot = 0; for (u32 k = 0; k < nt; k++) { const u32 i = k < nc ? cidx[k] : first + (k - nc); tot += q->offsets[i + 1] - q->offsets[i]; }
                T.bases.resize(tot + 1); u64 pos = 0;
                for (u32 k = 0; k < nt; k++) {
                    const u32 i = k < nc ? cidx[k] : first + (k - nc); const u64 l = q->offsets[i + 1] - q->offsets[i];
                    T.off[k] = pos; memcpy(T.bases.data() + pos, q->bases + q->offsets[i], l); pos += l;
                    T.index[k] = q->index ? q->index[i] : q->first_index + i; if (q->raw_len) T.raw[k] = q->raw_len[i];
                }
                T.off[nt] = pos; T.out.resize(nt);
                T.q = *q; T.q.n = nt; T.q.bases = T.bases.data(); T.q.offsets = T.off.data(); T.q.index = T.index.data(); T.q.raw_len = q->raw_len ? T.raw.data() : nullptr; T.q.n_context = nc;
            }
            rc = align_range(ctx, ln, &t[0].q, pe ? &t[1].q : nullptr, 0, nt, t[0].out.data(), pe ? t[1].out.data() : nullptr, pe ? tpair.data() : nullptr,
                             all_a, all_b, all_cap, all_off, &made, 0, &acc, true);
            if (rc == 0) {
                memcpy(out_a + first, t[0].out.data() + nc, (size_t)cnt * sizeof(bsl_hit));
                if (pe) { memcpy(out_b + first, t[1].out.data() + nc, (size_t)cnt * sizeof(bsl_hit)); memcpy(out_pair + first, tpair.data() + nc, (size_t)cnt * sizeof(bsl_pair)); }
            }
        }
        if (rc) return rc;
        all_off += made;
    }
    if (n_all) *n_all = all_off;
    { std::lock_guard<std::mutex> g(ctx->stats_mu); ctx->stats = acc; }
    return 0;
}
