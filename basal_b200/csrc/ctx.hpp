// ctx.hpp — host-side context object behind the opaque `bsl_ctx` of include/basal_gpu.h.
#pragma once
#include <mutex>
#include <string>
#include <vector>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

// One "lane" = a CUDA stream plus every device/pinned buffer one in-flight batch needs,
// so that several host threads can overlap H2D / kernels / D2H on one GPU.
#define BSL_NLANES 6
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t evk[640] = {};   // per-launch timing of search_round / pair_round
    std::mutex mu;
    // capacities
    size_t cap_slots = 0, cap_bases = 0, cap_words = 0, cap_hits = 0, cap_heavy_hits = 0, cap_pairs = 0, cap_bitmap = 0, cap_items = 0;
    // device buffers
    u8 *d_bases = nullptr; u64 *d_off = nullptr; u32 *d_index = nullptr; u16 *d_rawlen = nullptr;
    SlotMeta *d_meta = nullptr; SlotCounts *d_cnt = nullptr; uint2 *d_stat = nullptr; u8 *d_sched = nullptr; u64 *d_planes = nullptr; u32 *d_bits1 = nullptr; size_t cap_bits1 = 0;
    DevHit *d_hits = nullptr; DevHit *d_heavy_hits = nullptr;
    u32 *d_list[2] = {nullptr, nullptr}; u32 *d_heavy_list = nullptr; u32 *d_pe_list[2] = {nullptr, nullptr};
    bsl_hit *d_out = nullptr; bsl_pair *d_pair = nullptr; bsl_hit *d_all[2] = {nullptr, nullptr}; size_t cap_all = 0;
    u8 *d_minlvl = nullptr; uint2 *d_slot_item = nullptr; u32 *d_slot_flag = nullptr; u32 *d_flag_list = nullptr; uint4 *d_marks = nullptr;
    ItemHdr *d_hdr = nullptr; u32 *d_chunk_first = nullptr; u32 *d_bitmap = nullptr; u32 *d_flat_loc = nullptr;      // per-round flat candidate space
    DevCounters *d_ctr = nullptr;
    DevHit *d_bighits = nullptr; size_t cap_bighits = 0; u32 *d_wide_list = nullptr; size_t cap_wide = 0;   // large hit-list blocks, pairs for pair_round_wide
    u8 *d_st0 = nullptr; u32 *d_defer = nullptr; u32 *d_stale = nullptr; size_t cap_st0 = 0, cap_defer = 0, cap_stale = 0;   // carried seed-start state (mixed read lengths only)
    // pinned staging
    DevCounters *h_ctr = nullptr;
    u8 *h_bases = nullptr; u64 *h_off = nullptr; size_t hcap_bases = 0, hcap_off = 0;
    bsl_hit *h_out = nullptr; size_t hcap_out = 0; bsl_pair *h_pair = nullptr; size_t hcap_pair = 0;
};

struct bsl_ctx {
    int device = -1;
    bsl_params P;
    RuleTables rule;
    void *d_tables = nullptr;  // device copy of rule / budget / profile tables (DevTables in align.cu)
    // index
    bool has_index = false;
    DevIndex di;               // device pointers
    std::vector<u32> anchor, seqlen, rcoff;
    bsl_index_info info;
    // lanes
    Lane lanes[BSL_NLANES];
    bsl_stats stats;
    std::mutex stats_mu;
    char err[512];
    int sm_count = BSL_SM_COUNT;
    int occ_verify[4] = {0, 0, 0, 0};   // resident CTAs per SM of the verify_candidates variants
    int occ_bits = 0;                   // resident CTAs per SM of screen_bits (for occ_bits_wb words per read plane)
    size_t prep_smem = 0;                    // dynamic shared memory of the last prepare_reads launch (its carveout hint follows it)
    u32 occ_bits_wb = 0;
    bool kernels_configured = false;    // opt-in shared-memory sizes set on this context's device
    int occ_screen = 0;                 // resident CTAs per SM of screen_candidates
};

static inline void set_error(bsl_ctx *ctx, const char *fmt, ...) {
    static thread_local char tmp[512];
    va_list ap; va_start(ap, fmt); vsnprintf(tmp, sizeof tmp, fmt, ap); va_end(ap);
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", tmp);
    extern char g_bsl_last_error[512];
    snprintf(g_bsl_last_error, 512, "%s", tmp);
}

// rule.cpp
int  bsl_make_rule(const char from, const char *to, RuleTables *rt, char *err, size_t errlen);
// index.cu
int  bsl_index_build_impl(bsl_ctx *ctx, const u8 *cat, const u64 *off, const u32 *len, u32 n);
void bsl_index_free_impl(bsl_ctx *ctx);
int  bsl_index_download_impl(const bsl_ctx *ctx, u32 *bucket_start, u32 *n_fwd, u32 *loc, u64 *fwd, u64 *rc);
// align.cu
int  bsl_align_impl(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b, bsl_hit *out_a, bsl_hit *out_b, bsl_pair *out_pair,
                    bsl_hit *all_a, bsl_hit *all_b, u64 all_cap, u64 *n_all, int resident);
void bsl_lane_free(Lane &ln);
int  bsl_upload_params(bsl_ctx *ctx);
