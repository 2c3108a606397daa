// index.cu — GPU index builder (replaces RefSeq::Run_ConvertBinseq + Do_Formatdb).
//
//   reference ASCII --pack_planes--> 2-bit forward / reverse-complement planes
//                   --nx_transitions--> N/X run boundaries (UnmaskRegion blocks finished on the host)
//   blocks --gen_seeds--> (kmer*2+strand, global coordinate) for every I-th window
//          --cub radix sort--> seed table `loc[]`   (stable: forward entries first, ascending)
//          --bucket_bounds--> bucket[2K+1], cnt8[K]; counts --radix sort--> max_kmer_num
//
// Reference behaviour restated here: refbase.cpp:63-128 (BinSeq/cBinSeq/UnmaskRegion),
// :186-252 (anchors, margins), :254-255 (s_MakeSeed_1), :303-367 (count/alloc/cut-off),
// :419-439 (fill order).  All kernels are HBM-bound integer work; no tensor cores.
#include <algorithm>
#include <cstring>
#include <cub/cub.cuh>

#include "ctx.hpp"

namespace {

__constant__ u8 c_code[256];
__constant__ u8 c_rcode[256];
__constant__ u8 c_reg[256];

// ---- pack both strand planes ---------------------------------------------------------------
// One thread per 64-bit word of one sequence. Forward word w holds bases 32w..32w+31 (code 0
// beyond the sequence: the reference pads with 'N'); reverse word w holds the complement codes
// of padded positions P-1-32w .. P-32-32w (refbase.cpp:85-101).
__global__ void pack_planes(const u8 *__restrict__ ascii, const u64 *__restrict__ aoff, const u32 *__restrict__ alen,
                            const u64 *__restrict__ wstart, u32 nseq, u64 total_words, u64 *__restrict__ fwd, u64 *__restrict__ rc,
                            u32 *__restrict__ bit1, u32 *__restrict__ reg1) {
    u64 gw = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= total_words) return;
    u32 lo = 0, hi = nseq;                       // sequence containing word gw
    while (lo + 1 < hi) { u32 mid = (lo + hi) >> 1; if (gw >= wstart[mid]) lo = mid; else hi = mid; }
    u32 w = (u32)(gw - wstart[lo]); u32 n = alen[lo]; u32 nw = (u32)(wstart[lo + 1] - wstart[lo]); u32 P = nw * 32;
    const u8 *s = ascii + aoff[lo];
    u64 f = 0, r = 0;
    u32 b0 = w * 32, lowbits = 0, regular = 0;
#pragma unroll 8
    for (u32 k = 0; k < 32; k++) {
        u32 p = b0 + k; u32 cf = p < n ? c_code[s[p]] : 0u;
        f = (f << 2) | cf;
        lowbits = (lowbits << 1) | (cf & 1u);
        regular = (regular << 1) | ((p < n && c_reg[s[p]]) ? 1u : 0u);
        u32 q = P - 1 - p; u32 cr = q < n ? c_rcode[s[q]] : 0u;
        r = (r << 2) | cr;
    }
    fwd[BSL_REF_MARGIN + gw] = f; rc[BSL_REF_MARGIN + gw] = r;
    bit1[BSL_REF_MARGIN + gw] = lowbits; reg1[BSL_REF_MARGIN + gw] = regular;
}

// one bit per 256-base sector of the screening plane: set unless all 256 positions are ACGT bases of a sequence
__global__ void sector_flags(const u32 *__restrict__ reg1, u64 n_sectors, u32 *__restrict__ nflag) {
    u64 sct = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (sct < n_sectors) { const uint4 *p = (const uint4 *)(reg1 + sct * 8); uint4 a = p[0], b = p[1]; bad = (a.x & a.y & a.z & a.w & b.x & b.y & b.z & b.w) != 0xffffffffu; }
    u32 bal = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31u) == 0 && sct < n_sectors + 32) nflag[sct >> 5] = bal;
}

// ---- N/X run boundaries ----------------------------------------------------------------------
// Emits (pos<<1 | is_end) for every k where isNX(k) != isNX(k-1), with isNX(-1) = true.
__device__ __forceinline__ bool is_nx(u8 c) { return c == 'N' || c == 'X' || c == 'n' || c == 'x'; }

__global__ void nx_transitions(const u8 *__restrict__ ascii, u64 total, const u64 *__restrict__ aoff, u32 nseq,
                               u64 *__restrict__ out, u32 cap, u32 *__restrict__ n_out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    u32 lo = 0, hi = nseq;
    while (lo + 1 < hi) { u32 mid = (lo + hi) >> 1; if (i >= aoff[mid]) lo = mid; else hi = mid; }
    bool first = (i == aoff[lo]);
    bool cur = is_nx(ascii[i]); bool prev = first ? true : is_nx(ascii[i - 1]);
    if (cur != prev) { u32 k = atomicAdd(n_out, 1u); if (k < cap) out[k] = (i << 1) | (cur ? 1u : 0u); }
}

// ---- seed generation ----------------------------------------------------------------------------
struct SeedBlock { u32 g0; u32 strand; u64 first; };   // first global coordinate, strand, first entry index

__global__ void gen_seeds(const SeedBlock *__restrict__ blk, u32 nblk, u64 n_entries, u32 I, u32 s,
                          const u64 *__restrict__ fwd, const u64 *__restrict__ rc, u32 *__restrict__ keys, u32 *__restrict__ vals) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    u32 lo = 0, hi = nblk;
    while (lo + 1 < hi) { u32 mid = (lo + hi) >> 1; if (e >= blk[mid].first) lo = mid; else hi = mid; }
    SeedBlock b = blk[lo];
    u32 g = b.g0 + (u32)(e - b.first) * I;
    const u64 *pl = b.strand ? rc : fwd;
    u32 wi = g >> 5, o = (g & 31u) * 2;
    u64 w0 = pl[wi], w1 = pl[wi + 1];
    u64 x = o ? ((w0 << o) | (w1 >> (64 - o))) : w0;                       // s_MakeSeed_1 (refbase.cpp:254-255)
    u32 kmer = bsl_xt((u32)(x >> (64 - 2 * s)));
    keys[e] = kmer * 2 + b.strand; vals[e] = g;
}

// ---- bucket boundaries --------------------------------------------------------------------------
// keys sorted ascending; bucket[j] = lower_bound(keys, j) for j in 0..2K
__global__ void bucket_bounds(const u32 *__restrict__ keys, u64 n, u32 twoK, u32 *__restrict__ bucket) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    u32 prev = (i == 0) ? 0u : keys[i - 1] + 1;        // first j not yet covered
    u32 cur = (i == n) ? twoK + 1 : keys[i] + 1;       // one past the last j whose lower bound is i
    for (u32 j = prev; j < cur; j++) bucket[j] = (u32)i;
}

__global__ void bucket_counts(const u32 *__restrict__ bucket, u32 K, u32 *__restrict__ cnt, u8 *__restrict__ cnt8) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    u32 m = bucket[2 * k + 2] - bucket[2 * k];
    cnt[k] = m; cnt8[k] = m >= 0xFFu ? (u8)0xFFu : (u8)m;
}

// One 32-byte record per k-mer for seed_lookup: {entries, forward-strand entries, then the entries themselves when there are at
// most BSL_REC_INLINE of them, else the index of the first one in loc[]}. A look-up then costs one sector instead of one in
// bucket[] plus one in loc[]; the records sit behind loc[] in the same allocation, so that "where the entries are" is one index.
__global__ void bucket_records(const u32 *__restrict__ bucket, const u32 *__restrict__ loc, u32 K, u32 *__restrict__ rec) {
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const u32 e0 = bucket[2 * k], pm = bucket[2 * k + 2] - e0, nf = bucket[2 * k + 1] - e0;
    u32 w[6] = {e0, 0, 0, 0, 0, 0};
    if (pm <= BSL_REC_INLINE) for (u32 i = 0; i < BSL_REC_INLINE; i++) w[i] = i < pm ? loc[e0 + i] : 0u;
    uint4 *dst = (uint4 *)(rec + 8 * (size_t)k);
    dst[0] = make_uint4(pm, nf, w[0], w[1]); dst[1] = make_uint4(w[2], w[3], w[4], w[5]);
}

__global__ void split_bucket(const u32 *__restrict__ bucket, u32 K, u32 *__restrict__ start, u32 *__restrict__ nfwd) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K) return;
    if (start) start[k] = bucket[2 * k];
    if (nfwd && k < K) nfwd[k] = bucket[2 * k + 1] - bucket[2 * k];
}

template <typename T> int dmalloc(bsl_ctx *ctx, T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) { set_error(ctx, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return BSL_ENOMEM; }
    return 0;
}

} // namespace

void bsl_index_free_impl(bsl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree((void *)ctx->di.plane[0]); cudaFree((void *)ctx->di.bucket); cudaFree((void *)ctx->di.cnt8);
    cudaFree((void *)ctx->di.loc); cudaFree((void *)ctx->di.anchor); cudaFree((void *)ctx->di.seqlen); cudaFree((void *)ctx->di.rcoff);
    cudaFree((void *)ctx->di.bit1); cudaFree((void *)ctx->di.nflag); cudaFree((void *)ctx->di.ctab);
    memset(&ctx->di, 0, sizeof ctx->di); ctx->has_index = false;
}

int bsl_index_build_impl(bsl_ctx *ctx, const u8 *cat, const u64 *off, const u32 *len, u32 n) {
    if (!cat || !off || !len || n == 0) { set_error(ctx, "bsl_index_build: empty reference"); return BSL_EINVAL; }
    if (n > (1u << 17)) { set_error(ctx, "bsl_index_build: more than 131072 sequences (gHit::chr is 18 bits, param.h:37)"); return BSL_ELIMIT; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    bsl_index_free_impl(ctx);
    const u32 s = ctx->P.seed_size, I = ctx->P.index_interval;
    // ---- layout (refbase.cpp:222-244)
    std::vector<u64> wstart(n + 1), aoff(n + 1);
    ctx->anchor.assign(n + 1, 0); ctx->seqlen.assign(len, len + n); ctx->rcoff.assign(n, 0);
    u64 words = 0, bases = 0;
    for (u32 c = 0; c < n; c++) {
        if (len[c] >= 0x7fffff00u) { set_error(ctx, "sequence %u too long (must be < 2^31)", c); return BSL_ELIMIT; }
        u32 nw = (len[c] + 31) / 32 + 2;
        wstart[c] = words; aoff[c] = bases; ctx->rcoff[c] = nw * 32;
        ctx->anchor[c] = (u32)((words + BSL_REF_MARGIN) * 32);
        words += nw; bases += len[c];
    }
    wstart[n] = words; aoff[n] = bases;
    u64 n_words = words + 2 * BSL_REF_MARGIN;
    if (n_words * 32 >= (1ull << 32)) { set_error(ctx, "reference too large: concatenated coordinates must stay below 2^32 (refbase.cpp:222)"); return BSL_ELIMIT; }
    ctx->anchor[n] = (u32)((words + BSL_REF_MARGIN) * 32);

    cudaMemcpyToSymbol(c_code, ctx->rule.code, 256); cudaMemcpyToSymbol(c_rcode, ctx->rule.rcode, 256); cudaMemcpyToSymbol(c_reg, ctx->rule.reg, 256);

    u8 *d_ascii = nullptr; u64 *d_aoff = nullptr, *d_wstart = nullptr; u32 *d_alen = nullptr;
    u64 *d_fwd = nullptr, *d_rc = nullptr;
    int rc_ = 0;
    if ((rc_ = dmalloc(ctx, &d_ascii, bases + 64)) || (rc_ = dmalloc(ctx, &d_aoff, n + 1)) || (rc_ = dmalloc(ctx, &d_wstart, n + 1)) ||
        (rc_ = dmalloc(ctx, &d_alen, n)) || (rc_ = dmalloc(ctx, &d_fwd, 2 * ((n_words + 3) & ~3ull) + 8))) return rc_;
    d_rc = d_fwd + ((n_words + 3) & ~3ull);   // 32-byte aligned: the kernels gather whole sectors of either plane                 // both strand planes in one allocation: one L2 access-policy window covers them
    // screening plane + a temporary "is an ACGT base of a sequence" plane (margins and padding stay 0)
    const u64 n_words8 = (n_words + 7) / 8 * 8, n_sectors = n_words8 / 8;
    u32 *d_bit1 = nullptr, *d_reg1 = nullptr, *d_nflag = nullptr; uint2 *d_ctab = nullptr;
    if ((rc_ = dmalloc(ctx, &d_bit1, n_words8 + 64)) || (rc_ = dmalloc(ctx, &d_reg1, n_words8 + 64)) || (rc_ = dmalloc(ctx, &d_nflag, n_sectors / 32 + 8))) return rc_;
    CUDA_TRY(cudaMemset(d_bit1, 0, (n_words8 + 64) * 4)); CUDA_TRY(cudaMemset(d_reg1, 0, (n_words8 + 64) * 4));
    for (u32 c = 0; c < n; c++) CUDA_TRY(cudaMemcpy(d_ascii + aoff[c], cat + off[c], len[c], cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_aoff, aoff.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_wstart, wstart.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_alen, len, n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(d_fwd, 0, n_words * 8)); CUDA_TRY(cudaMemset(d_rc, 0, n_words * 8));
    {
        u64 blocks = (words + 255) / 256;
        pack_planes<<<(unsigned)blocks, 256>>>(d_ascii, d_aoff, d_alen, d_wstart, n, words, d_fwd, d_rc, d_bit1, d_reg1);
        CUDA_TRY(cudaGetLastError());
        sector_flags<<<(unsigned)((n_sectors + 32 + 255) / 256), 256>>>(d_reg1, n_sectors, d_nflag);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d_reg1); d_reg1 = nullptr;
    }
    // ---- UnmaskRegion blocks (refbase.cpp:103-128): GPU finds N/X run boundaries, host finishes
    const u32 tcap = 1u << 24;
    u64 *d_tr = nullptr; u32 *d_ntr = nullptr;
    if ((rc_ = dmalloc(ctx, &d_tr, tcap)) || (rc_ = dmalloc(ctx, &d_ntr, 1))) return rc_;
    CUDA_TRY(cudaMemset(d_ntr, 0, 4));
    if (bases) { nx_transitions<<<(unsigned)((bases + 255) / 256), 256>>>(d_ascii, bases, d_aoff, n, d_tr, tcap, d_ntr); CUDA_TRY(cudaGetLastError()); }
    u32 ntr = 0; CUDA_TRY(cudaMemcpy(&ntr, d_ntr, 4, cudaMemcpyDeviceToHost));
    if (ntr > tcap) { set_error(ctx, "reference has more than %u N/X run boundaries", tcap); return BSL_ELIMIT; }
    std::vector<u64> tr(ntr); if (ntr) CUDA_TRY(cudaMemcpy(tr.data(), d_tr, (size_t)ntr * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_tr); cudaFree(d_ntr); cudaFree(d_ascii); d_ascii = nullptr;
    std::sort(tr.begin(), tr.end());
    struct Blk { u32 chr, b, e; };
    std::vector<Blk> fblk;                                 // forward blocks in (chr, begin) order
    {
        size_t t = 0;
        for (u32 c = 0; c < n; c++) {
            const u8 *q = cat + off[c];
            while (t < tr.size() && (tr[t] >> 1) < aoff[c + 1]) {
                if (tr[t] & 1) { t++; continue; }           // an end without a start cannot happen; skip defensively
                u32 rs = (u32)((tr[t] >> 1) - aoff[c]); t++;
                u32 re = len[c];
                if (t < tr.size() && (tr[t] >> 1) < aoff[c + 1] && (tr[t] & 1)) { re = (u32)((tr[t] >> 1) - aoff[c]); t++; }
                u32 b = rs; while (b < re && !ctx->rule.reg[q[b]]) b++;     // block starts at the first ACGT of the run
                if (b >= re || re - b < 16) continue;
                fblk.push_back({c, b, re});
            }
        }
    }
    std::vector<SeedBlock> sb; u64 ne = 0;
    auto add_block = [&](u32 c, u32 strand, u32 b, u32 e) {
        u32 p0 = (b / I) * I, last = ((e - s) / I) * I;
        if (last < p0) return;
        sb.push_back({ctx->anchor[c] + p0, strand, ne}); ne += (u64)(last - p0) / I + 1;
    };
    for (const Blk &k : fblk) add_block(k.chr, 0, k.b, k.e);
    u64 ne_fwd = ne;
    {   // reverse-strand blocks [P-e, P-b), sorted by (chr, begin): per sequence in reverse order of the forward ones
        size_t i = 0;
        while (i < fblk.size()) {
            size_t j = i; while (j < fblk.size() && fblk[j].chr == fblk[i].chr) j++;
            for (size_t k = j; k-- > i;) { u32 P = ctx->rcoff[fblk[k].chr]; add_block(fblk[k].chr, 1, P - fblk[k].e, P - fblk[k].b); }
            i = j;
        }
    }
    (void)ne_fwd;
    u32 K = 1; for (u32 i = 0; i < s; i++) K *= 3;
    if (ne + 8ull * K + 64 >= (1ull << 32)) { set_error(ctx, "seed table would exceed 2^32 entries; use a larger -I"); return BSL_ELIMIT; }

    // ---- seeds -> sort -> table
    const size_t rec_base = (ne + 32 + 7) / 8 * 8;                       // the bucket records follow loc[] (32-byte aligned)
    SeedBlock *d_sb = nullptr; u32 *d_keys = nullptr, *d_vals = nullptr, *d_keys2 = nullptr, *d_loc = nullptr, *d_bucket = nullptr; u8 *d_cnt8 = nullptr;
    if ((rc_ = dmalloc(ctx, &d_sb, sb.size())) || (rc_ = dmalloc(ctx, &d_keys, ne)) || (rc_ = dmalloc(ctx, &d_vals, ne)) ||
        (rc_ = dmalloc(ctx, &d_keys2, ne)) || (rc_ = dmalloc(ctx, &d_loc, rec_base + 8 * (size_t)K)) || (rc_ = dmalloc(ctx, &d_bucket, 2 * (size_t)K + 2)) ||
        (rc_ = dmalloc(ctx, &d_cnt8, (size_t)K + 16))) return rc_;
    if (!sb.empty()) CUDA_TRY(cudaMemcpy(d_sb, sb.data(), sb.size() * sizeof(SeedBlock), cudaMemcpyHostToDevice));
    if (ne) {
        gen_seeds<<<(unsigned)((ne + 255) / 256), 256>>>(d_sb, (u32)sb.size(), ne, I, s, d_fwd, d_rc, d_keys, d_vals);
        CUDA_TRY(cudaGetLastError());
        int end_bit = 1; while ((1ull << end_bit) < 2ull * K) end_bit++;
        size_t tmp_bytes = 0; void *d_tmp = nullptr;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_loc, (long long)ne, 0, end_bit);
        CUDA_TRY(cudaMalloc(&d_tmp, tmp_bytes + 16));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_loc, (long long)ne, 0, end_bit));
        CUDA_TRY(cudaGetLastError());
        cudaFree(d_tmp);
    }
    CUDA_TRY(cudaMemset(d_loc + ne, 0, 32 * 4));
    bucket_bounds<<<(unsigned)((ne + 1 + 255) / 256), 256>>>(d_keys2, ne, 2 * K, d_bucket);
    CUDA_TRY(cudaGetLastError());
    cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_sb);
    // ---- counts, cnt8, over-represented k-mer cut-off (refbase.cpp:362-363)
    u32 *d_cnt = nullptr, *d_cnt_sorted = d_keys2;          // reuse: keys2 has >= K entries only if ne >= K; allocate otherwise
    if ((rc_ = dmalloc(ctx, &d_cnt, K))) return rc_;
    bucket_counts<<<(K + 255) / 256, 256>>>(d_bucket, K, d_cnt, d_cnt8);
    CUDA_TRY(cudaGetLastError());
    bucket_records<<<(K + 255) / 256, 256>>>(d_bucket, d_loc, K, d_loc + rec_base);
    CUDA_TRY(cudaGetLastError());
    u32 *d_sorted = nullptr; if ((rc_ = dmalloc(ctx, &d_sorted, K))) return rc_;
    (void)d_cnt_sorted;
    {
        size_t tmp_bytes = 0; void *d_tmp = nullptr;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_cnt, d_sorted, (int)(K - 1));
        CUDA_TRY(cudaMalloc(&d_tmp, tmp_bytes + 16));
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_cnt, d_sorted, (int)(K - 1)));
        cudaFree(d_tmp);
    }
    u32 rank = (u32)((float)K * (1 - ctx->P.max_kmer_ratio)) - 1;      // float32 on purpose
    u32 maxk = 0;
    if (rank < K - 1) CUDA_TRY(cudaMemcpy(&maxk, d_sorted + rank, 4, cudaMemcpyDeviceToHost));
    else if (rank == K - 1) CUDA_TRY(cudaMemcpy(&maxk, d_cnt + (K - 1), 4, cudaMemcpyDeviceToHost));   // the unsorted last element
    else { set_error(ctx, "-k ratio %g puts the cut-off rank outside the table", (double)ctx->P.max_kmer_ratio); return BSL_EINVAL; }
    cudaFree(d_sorted); cudaFree(d_cnt); cudaFree(d_keys2); cudaFree(d_aoff); cudaFree(d_wstart); cudaFree(d_alen);

    u32 *d_anchor = nullptr, *d_len = nullptr, *d_rcoff = nullptr;
    if ((rc_ = dmalloc(ctx, &d_anchor, n + 1)) || (rc_ = dmalloc(ctx, &d_len, n)) || (rc_ = dmalloc(ctx, &d_rcoff, n))) return rc_;
    CUDA_TRY(cudaMemcpy(d_anchor, ctx->anchor.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_len, len, n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_rcoff, ctx->rcoff.data(), n * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());

    // ---- strand table for the screen: which sequence owns a block of 2^BSL_CTAB_SHIFT global coordinates
    {
        const u32 nblk = 1u << (32 - BSL_CTAB_SHIFT);
        std::vector<uint2> ctab(nblk, make_uint2(0u, 0u));
        for (u32 c = 0; c < n; c++) {
            const u64 a0 = ctx->anchor[c], a1 = a0 + ctx->rcoff[c];          // [a0, a1) = coordinates of sequence c (both strands)
            const u64 b0 = (a0 + (1ull << BSL_CTAB_SHIFT) - 1) >> BSL_CTAB_SHIFT, b1 = a1 >> BSL_CTAB_SHIFT;   // blocks wholly inside
            for (u64 b = b0; b < b1 && b < nblk; b++) ctab[b] = make_uint2((u32)(2 * a0 + ctx->rcoff[c] - 1), (u32)a1);
        }
        if ((rc_ = dmalloc(ctx, &d_ctab, nblk))) return rc_;
        CUDA_TRY(cudaMemcpy(d_ctab, ctab.data(), nblk * sizeof(uint2), cudaMemcpyHostToDevice));
    }
    DevIndex &di = ctx->di;
    di.bit1 = d_bit1; di.nflag = d_nflag; di.ctab = d_ctab;
    {   // the screen needs: one convert-to base (from = 01, to = 11 share the low bit) and complement = XOR with a constant on the low bit
        const u8 *cd = ctx->rule.code, *rd = ctx->rule.rcode;
        const u32 f = (cd['A'] ^ rd['A']) & 1u;
        const bool uniform = ((cd['C'] ^ rd['C']) & 1u) == f && ((cd['G'] ^ rd['G']) & 1u) == f && ((cd['T'] ^ rd['T']) & 1u) == f;
        // also rules whose substitution set is empty ("T:-"): CountMismatch_new (align.h:199-239) then compares every read base exactly,
        // so differing low bits are a mismatch there as well
        const bool dash_only = strcmp(ctx->P.to_bases, "-") == 0;
        di.flip = f; di.has_bit1 = ((ctx->rule.single || dash_only) && uniform) ? 1u : 0u;
    }
    di.plane[0] = d_fwd; di.plane[1] = d_rc; di.bucket = d_bucket; di.cnt8 = d_cnt8; di.loc = d_loc; di.rec_base = (u32)rec_base;
    di.anchor = d_anchor; di.seqlen = d_len; di.rcoff = d_rcoff; di.nseq = n; di.K = K; di.maxk = maxk; di.n_words = n_words; di.n_entries = ne;
    memset(&ctx->info, 0, sizeof ctx->info);
    ctx->info.n_seq = n; ctx->info.n_kmers = K; ctx->info.sum_length = bases; ctx->info.n_words = n_words; ctx->info.n_entries = ne; ctx->info.max_kmer_num = maxk;
    ctx->has_index = true;
    return bsl_upload_params(ctx);
}

int bsl_index_download_impl(const bsl_ctx *cctx, u32 *bucket_start, u32 *n_fwd, u32 *loc, u64 *fwd, u64 *rc) {
    bsl_ctx *ctx = const_cast<bsl_ctx *>(cctx);
    if (!ctx->has_index) { set_error(ctx, "no index"); return BSL_ESTATE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const DevIndex &di = ctx->di; u32 K = di.K;
    if (bucket_start || n_fwd) {
        u32 *d_s = nullptr, *d_f = nullptr;
        CUDA_TRY(cudaMalloc(&d_s, ((size_t)K + 1) * 4)); CUDA_TRY(cudaMalloc(&d_f, (size_t)K * 4));
        split_bucket<<<(K + 1 + 255) / 256, 256>>>(di.bucket, K, d_s, d_f);
        CUDA_TRY(cudaGetLastError());
        if (bucket_start) CUDA_TRY(cudaMemcpy(bucket_start, d_s, ((size_t)K + 1) * 4, cudaMemcpyDeviceToHost));
        if (n_fwd) CUDA_TRY(cudaMemcpy(n_fwd, d_f, (size_t)K * 4, cudaMemcpyDeviceToHost));
        cudaFree(d_s); cudaFree(d_f);
    }
    if (loc) CUDA_TRY(cudaMemcpy(loc, di.loc, di.n_entries * 4, cudaMemcpyDeviceToHost));
    if (fwd) CUDA_TRY(cudaMemcpy(fwd, di.plane[0], di.n_words * 8, cudaMemcpyDeviceToHost));
    if (rc) CUDA_TRY(cudaMemcpy(rc, di.plane[1], di.n_words * 8, cudaMemcpyDeviceToHost));
    return 0;
}
