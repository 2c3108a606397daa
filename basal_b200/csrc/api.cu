// api.cu — the extern "C" boundary declared in include/basal_gpu.h.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.hpp"

char g_bsl_last_error[512] = "";

extern "C" {

int bsl_abi_version(void) { return BSL_ABI_VERSION; }

// Param::Param defaults (param.cpp:7-68)
void bsl_params_default(bsl_params *p) {
    memset(p, 0, sizeof *p);
    p->from_base = 'C'; strcpy(p->to_bases, "T");
    p->seed_size = 16; p->index_interval = 4; p->max_snp_num = 110; p->gap = 0; p->max_num_hits = 100;
    p->min_insert = 28; p->max_insert = 1000; p->chains = 0; p->report_repeat_hits = 1; p->randseed = 0;
    p->max_ns = 5; p->min_read_size = 16; p->max_kmer_ratio = 5e-7f;
}

const char *bsl_last_error(const bsl_ctx *ctx) { return ctx ? ctx->err : g_bsl_last_error; }

int bsl_ctx_create(bsl_ctx **out, int device, const bsl_params *p) {
    if (!out || !p) { set_error(nullptr, "bsl_ctx_create: null argument"); return BSL_EINVAL; }
    *out = nullptr;
    // the reference's own range checks
    if (p->seed_size > 16 || p->seed_size < 10) { set_error(nullptr, "seed size must be between 10 and 16"); return BSL_EINVAL; }       // param.cpp:109
    if (p->index_interval < 1 || p->index_interval > 16) { set_error(nullptr, "index interval exceeds max value:16"); return BSL_EINVAL; }   // main.cpp:321
    if (p->gap > 3) { set_error(nullptr, "gap length exceeds max value:3"); return BSL_EINVAL; }                                        // main.cpp:301
    if (p->max_num_hits > 1000 || p->max_num_hits < 1) { set_error(nullptr, "number of multi-hits exceeds max value:1000"); return BSL_EINVAL; }   // main.cpp:340
    if (p->report_repeat_hits > 2) { set_error(nullptr, "invalid -r value: %u, must be 0, 1, or 2.", p->report_repeat_hits); return BSL_EINVAL; }
    if (p->randseed == 0) { set_error(nullptr, "-S 0 draws from rand_r in the reference and is not reproducible; pass a non-zero seed"); return BSL_EINVAL; }
    if (p->max_snp_num >= 200) { set_error(nullptr, "max_snp_num encoding out of range"); return BSL_EINVAL; }
    RuleTables rt; char msg[256];
    int rc = bsl_make_rule(p->from_base, p->to_bases, &rt, msg, sizeof msg);
    if (rc) { set_error(nullptr, "%s", msg); return rc; }
    int ndev = 0; cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error(nullptr, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e)); return BSL_ENODEV; }
    if (device < 0 || device >= ndev) { set_error(nullptr, "device %d out of range (0..%d)", device, ndev - 1); return BSL_ENODEV; }
    cudaDeviceProp prop; e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { set_error(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return BSL_ENODEV; }
    if (prop.major != 10) { set_error(nullptr, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor); return BSL_ENODEV; }
    bsl_ctx *c = new (std::nothrow) bsl_ctx();
    if (!c) return BSL_ENOMEM;
    c->device = device; c->P = *p; c->rule = rt; c->err[0] = 0; c->sm_count = prop.multiProcessorCount;
    memset(&c->di, 0, sizeof c->di); memset(&c->info, 0, sizeof c->info); memset(&c->stats, 0, sizeof c->stats);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { set_error(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e)); delete c; return BSL_ENODEV; }
    rc = bsl_upload_params(c);
    if (rc) { snprintf(g_bsl_last_error, sizeof g_bsl_last_error, "%s", c->err); delete c; return rc; }
    *out = c; return BSL_OK;
}

void bsl_ctx_destroy(bsl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    bsl_index_free_impl(ctx);
    for (auto &ln : ctx->lanes) bsl_lane_free(ln);
    cudaFree(ctx->d_tables);
    delete ctx;
}

int bsl_index_build(bsl_ctx *ctx, const uint8_t *seq_concat, const uint64_t *seq_off, const uint32_t *seq_len, uint32_t n_seq) {
    if (!ctx) return BSL_EINVAL;
    return bsl_index_build_impl(ctx, seq_concat, seq_off, seq_len, n_seq);
}

int bsl_index_info_get(const bsl_ctx *ctx, bsl_index_info *info) {
    if (!ctx || !info) return BSL_EINVAL;
    if (!ctx->has_index) return BSL_ESTATE;
    *info = ctx->info; return 0;
}

int bsl_index_download(const bsl_ctx *ctx, uint32_t *bucket_start, uint32_t *n_fwd, uint32_t *loc, uint64_t *fwd_plane, uint64_t *rc_plane) {
    if (!ctx) return BSL_EINVAL;
    return bsl_index_download_impl(ctx, bucket_start, n_fwd, loc, fwd_plane, rc_plane);
}

void bsl_index_free(bsl_ctx *ctx) { bsl_index_free_impl(ctx); }

int bsl_align_se(bsl_ctx *ctx, const bsl_batch *reads, bsl_hit *out, bsl_hit *all_hits, uint64_t all_cap, uint64_t *n_all) {
    if (!ctx) return BSL_EINVAL;
    return bsl_align_impl(ctx, reads, nullptr, out, nullptr, nullptr, all_hits, nullptr, all_cap, n_all, 0);
}

int bsl_align_pe(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b, bsl_hit *out_a, bsl_hit *out_b, bsl_pair *out_pair,
                 bsl_hit *all_a, bsl_hit *all_b, uint64_t all_cap, uint64_t *n_all) {
    if (!ctx) return BSL_EINVAL;
    if (!b) { set_error(ctx, "bsl_align_pe: second batch is null"); return BSL_EINVAL; }
    return bsl_align_impl(ctx, a, b, out_a, out_b, out_pair, all_a, all_b, all_cap, n_all, 0);
}

int bsl_align_rerun(bsl_ctx *ctx, const bsl_batch *a, const bsl_batch *b) {
    if (!ctx) return BSL_EINVAL;
    return bsl_align_impl(ctx, a, b, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 1);
}

int bsl_stats_get(const bsl_ctx *ctx, bsl_stats *st) {
    if (!ctx || !st) return BSL_EINVAL;
    bsl_ctx *c = const_cast<bsl_ctx *>(ctx);
    std::lock_guard<std::mutex> g(c->stats_mu);
    *st = ctx->stats; return 0;
}

void *bsl_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr; return p; }
void bsl_host_free(void *p) { if (p) cudaFreeHost(p); }

// FilterReads budget (align.cpp:550-561)
uint32_t bsl_read_budget(const bsl_params *p, uint32_t raw_len, uint32_t len) {
    if (!raw_len || !len) return 0;
    uint32_t B = p->max_snp_num < 100 ? p->max_snp_num : (uint32_t)((p->max_snp_num - 100) / 100.0 * raw_len + 0.5);
    if (p->gap > 0) B = B + 1 + p->gap;
    if (B > BSL_MAXSNPS) B = BSL_MAXSNPS;
    return (B + 1) * (len - 1) / raw_len;
}

uint32_t bsl_myrand(uint32_t read_index, uint32_t randseed) { return bsl_rand(read_index, randseed); }

} // extern "C"
