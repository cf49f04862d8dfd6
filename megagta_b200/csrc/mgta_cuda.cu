// mgta_cuda.cu -- C ABI (include/mgta_cuda.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side schedule of one context (= one GPU = one contiguous lv1-bucket shard).  It replaces
// CX1::run() (reference cx1.h:443-623) for the read2sdbg plug-in: instead of ~8 lv1 passes sized
// from host RAM that each re-scan all reads into int32 offset deltas and then sort bucket by bucket
// on CPU threads, a stage is
//     histogram  ->  [per HBM-sized bucket-range batch]  extract+scatter -> MSD levels (only if a
//     bucket exceeds the on-chip tile) -> on-chip sort + count/emit
// with everything between the histogram read-back and the batch's result resident in HBM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mgta_cuda.h"
#include "kernels.cuh"

using namespace mgta;

namespace {

std::string g_create_error;

enum { CTR_TICKET = 0, CTR_NLIST0 = 1, CTR_NLIST1 = 2, CTR_NGIANTS = 3, CTR_ERR = 4, CTR_COUNT = 8 };
enum { PH_HIST = 0, PH_EXTRACT = 1, PH_PARTITION = 2, PH_SORT = 3, PH_COUNT = 4 };

struct Timed {
    int phase;
    cudaEvent_t a, b;
};

}  // namespace

struct mgta_ctx {
    mgta_opts opt;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int sm_count = 148;
    // reads
    uint32_t *d_seq = nullptr;
    uint64_t *d_start = nullptr;
    uint64_t n_words = 0, n_reads = 0, n_short = 0, total_bases = 0;
    int max_len = 0;
    uint32_t *d_solid = nullptr;
    uint64_t solid_words = 0;
    // small device state
    unsigned long long *d_hist = nullptr, *d_cursor = nullptr, *d_meta = nullptr, *d_totals = nullptr, *d_ec = nullptr;
    unsigned *d_ctr = nullptr;
    unsigned long long *h_pin = nullptr;   // pinned: hist / cursor staging [2 * 65536] + misc [64]
    unsigned char *h_out = nullptr;        // pinned output staging
    size_t h_out_bytes = 0;
    // arena
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;
    // results
    std::vector<int64_t> hist;             // last histogram (whole bucket space)
    int shard_lo = 0, shard_hi = NUM_BUCKETS;
    mgta_stage_stats stats[2];
    std::vector<Timed> timed;
    uint64_t n_dollar = 0;                 // stage-2 items with a == $ (from the last stage-2 histogram)
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                             \
            return MGTA_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define FAIL(code, ...)                              \
    do {                                             \
        char buf_[512];                              \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);   \
        ctx->err = buf_;                             \
        return (code);                               \
    } while (0)

namespace {

#define W_SWITCH(W, STMT)                                   \
    switch (W) {                                            \
        case 1: { constexpr int WW = 1; STMT; } break;      \
        case 2: { constexpr int WW = 2; STMT; } break;      \
        case 3: { constexpr int WW = 3; STMT; } break;      \
        case 4: { constexpr int WW = 4; STMT; } break;      \
        case 5: { constexpr int WW = 5; STMT; } break;      \
        case 6: { constexpr int WW = 6; STMT; } break;      \
        case 7: { constexpr int WW = 7; STMT; } break;      \
        case 8: { constexpr int WW = 8; STMT; } break;      \
        case 9: { constexpr int WW = 9; STMT; } break;      \
        default: break;                                     \
    }

template <int STAGE, int MODE>
void launch_walk(int W, const WalkParams &P, unsigned grid, cudaStream_t st) {
    W_SWITCH(W, (k_walk<WW, STAGE, MODE><<<grid, WALK_THREADS, 0, st>>>(P)));
}

struct Plan {
    int stage, W, IW, k;
    unsigned CAPI, T, C;
    size_t chunk_smem;
    std::vector<std::pair<int, int>> levels;   // (msb start bit, nbits)
    int n_pass;
    short pass_lsb[MAX_PASSES];
    unsigned char pass_nb[MAX_PASSES];
    // arena carve (byte offsets)
    uint64_t cap;                              // items per batch
    size_t off_a, off_b, off_flags, off_win, off_state, off_list0, off_list1, off_giants, off_out, total;
    unsigned list_cap, giants_cap;
    uint64_t out_cap;
};

size_t chunk_smem_bytes(int IW, unsigned capi) {
    return (size_t)IW * capi * 4 + CHUNK_WARPS * 256 * 2 + 256 * 4 + 2 * (capi / 32 + 2) * 4 + (CHUNK_WARPS + 2) * 4 +
           2 * (size_t)capi * 2;
}

void make_plan(Plan &pl, int stage, int k, int cap_override) {
    pl.stage = stage;
    pl.k = k;
    pl.W = stage == 1 ? key_words_s1(k) : key_words_s2(k);
    pl.IW = pl.W + (stage == 1 ? 2 : 0);
    // on-chip tile: two CTAs per SM (<= ~112 KB dynamic shared memory each)
    unsigned capi = 8192;
    while (capi > 512 && chunk_smem_bytes(pl.IW, capi) > 112 * 1024) capi -= 512;
    if (cap_override > 0) capi = std::max(64, (cap_override / 64) * 64);
    pl.CAPI = capi;
    pl.T = capi / 2;
    pl.C = capi / 2;
    pl.chunk_smem = chunk_smem_bytes(pl.IW, capi);
    // MSD digit levels over the (k-1)-mer bits below the 16-bit lv1 prefix (never split a group)
    const int GB = 2 * (k - 1);
    pl.levels.clear();
    for (int s = 16; s < GB;) {
        int nb = std::min(8, GB - s);
        pl.levels.push_back({s, nb});
        s += nb;
    }
    // LSD digit passes of the on-chip sort: flag bits, skip the zero padding, then the sequence bits
    const int TB = 32 * pl.W, FB = stage == 1 ? 6 : 4, SB = stage == 1 ? 2 * (k - 1) : 2 * k;
    int n = 0, lsb = 0;
    if (TB - SB > FB) {
        pl.pass_lsb[n] = 0; pl.pass_nb[n] = (unsigned char)FB; ++n;
        lsb = TB - SB;
    }
    for (; lsb < TB && n < MAX_PASSES; lsb += 8) {
        pl.pass_lsb[n] = (short)lsb; pl.pass_nb[n] = (unsigned char)std::min(8, TB - lsb); ++n;
    }
    pl.n_pass = n;
}

// carve the arena for batches of up to `cap` items; returns total bytes
size_t carve(Plan &pl, uint64_t cap, uint64_t n_dollar) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    cap = (cap + 31) & ~(uint64_t)31;
    pl.cap = cap;
    const uint64_t n_win = cap / pl.C + 2;
    pl.list_cap = (unsigned)std::min<uint64_t>(cap / pl.T + NUM_BUCKETS + 16, 0x7FFFFFFFu);
    pl.giants_cap = (unsigned)std::min<uint64_t>(cap / pl.T + 16, 0x7FFFFFFFu);
    size_t o = 0;
    pl.off_a = o; o += al((size_t)pl.IW * cap * 4);
    pl.off_b = o; o += al((size_t)pl.IW * cap * 4);
    pl.off_flags = o; o += al((cap / 32 + 64) * 4);
    pl.off_win = o; o += al(n_win * 4);
    pl.off_state = o; o += al(n_win * 8);
    pl.off_list0 = o; o += al((size_t)pl.list_cap * sizeof(Seg));
    pl.off_list1 = o; o += al((size_t)pl.list_cap * sizeof(Seg));
    pl.off_giants = o; o += al((size_t)pl.giants_cap * sizeof(Giant));
    pl.off_out = o;
    // stage-2 record stream, guaranteed bound: every record is a run of >= 1 items (2 B), a u16 multiplicity
    // needs a run of > 254 items, a tip label needs an item with a == $ (counted by the histogram pass)
    const int wpt = (2 * pl.k + 31) / 32;
    pl.out_cap = pl.stage == 2 ? 2 * cap + 2 * (cap / 255 + 1) + 4ull * wpt * std::min<uint64_t>(cap, n_dollar) + 4096 : 0;
    o += al(pl.out_cap);
    pl.total = o;
    return o;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int mgta_abi_version(void) { return 1; }

extern "C" int mgta_words_per_key(int stage, int kmer_k) { return stage == 1 ? key_words_s1(kmer_k) : key_words_s2(kmer_k); }

extern "C" const char *mgta_last_error(const mgta_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int mgta_ctx_create(const mgta_opts *opts, mgta_ctx **out) {
    if (!opts || !out) { g_create_error = "null argument"; return MGTA_ERR_ARG; }
    *out = nullptr;
    if (opts->kmer_k < 9 || opts->kmer_k > MGTA_MAX_K) { g_create_error = "kmer_k must be in [9, 127]"; return MGTA_ERR_ARG; }
    if (opts->min_count < 1) { g_create_error = "min_count must be >= 1"; return MGTA_ERR_ARG; }
    if (opts->world < 1 || opts->rank < 0 || opts->rank >= opts->world) { g_create_error = "bad rank/world"; return MGTA_ERR_ARG; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return MGTA_ERR_CUDA;
    }
    if (opts->device < 0 || opts->device >= n_dev) { g_create_error = "bad device ordinal"; return MGTA_ERR_ARG; }
    mgta_ctx *ctx = new mgta_ctx();
    ctx->opt = *opts;
    memset(ctx->stats, 0, sizeof(ctx->stats));
    auto fail = [&](const char *what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete ctx;
        return (int)MGTA_ERR_CUDA;
    };
    if ((e = cudaSetDevice(opts->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, opts->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    ctx->sm_count = prop.multiProcessorCount;
    if (opts->stream) ctx->stream = (cudaStream_t)opts->stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
        ctx->own_stream = true;
    }
    if ((e = cudaMalloc(&ctx->d_hist, NUM_BUCKETS * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_cursor, NUM_BUCKETS * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_meta, NUM_BUCKETS * 3 * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_totals, 16 * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_ec, NUM_BUCKETS * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_ctr, CTR_COUNT * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&ctx->h_pin, (2 * NUM_BUCKETS + 64) * 8, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    *out = ctx;
    return MGTA_OK;
}

extern "C" void mgta_ctx_destroy(mgta_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->opt.device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_seq); cudaFree(ctx->d_start); cudaFree(ctx->d_solid);
    cudaFree(ctx->d_hist); cudaFree(ctx->d_cursor); cudaFree(ctx->d_meta); cudaFree(ctx->d_totals); cudaFree(ctx->d_ec);
    cudaFree(ctx->d_ctr); cudaFree(ctx->arena);
    cudaFreeHost(ctx->h_pin); cudaFreeHost(ctx->h_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

namespace {
int alloc_reads(mgta_ctx *ctx, uint64_t n_words, uint64_t n_reads, uint64_t n_short, uint64_t total, int max_len) {
    if (n_reads == 0 || n_short > n_reads) FAIL(MGTA_ERR_ARG, "reads: bad counts");
    if (total == 0) FAIL(MGTA_ERR_ARG, "reads: no bases");
    if (n_words * 16 < total) FAIL(MGTA_ERR_ARG, "reads: packed_seq shorter than start_idx says");
    if (total >= (1ull << 40) - 1) FAIL(MGTA_ERR_ARG, "reads: more than 2^40 bases");
    CK(cudaSetDevice(ctx->opt.device));
    const uint64_t padded = ((n_words + 3) & ~3ull) + SEQ_PAD_WORDS;
    const uint64_t solid_words = (total + 31) / 32 + 4;
    if (n_words != ctx->n_words || n_reads != ctx->n_reads || total != ctx->total_bases || !ctx->d_seq) {     // reuse buffers of equal shape
        cudaFree(ctx->d_seq); cudaFree(ctx->d_start); cudaFree(ctx->d_solid);
        ctx->d_seq = nullptr; ctx->d_start = nullptr; ctx->d_solid = nullptr;
        CK(cudaMalloc(&ctx->d_seq, padded * 4));
        CK(cudaMalloc(&ctx->d_start, (n_reads + 1) * 8));
        CK(cudaMalloc(&ctx->d_solid, solid_words * 4));
    }
    ctx->solid_words = solid_words;
    CK(cudaMemsetAsync(ctx->d_seq + n_words, 0, (padded - n_words) * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_solid, 0, ctx->solid_words * 4, ctx->stream));
    ctx->n_words = n_words; ctx->n_reads = n_reads; ctx->n_short = n_short; ctx->total_bases = total;
    ctx->max_len = max_len;
    return MGTA_OK;
}
}  // namespace

extern "C" int mgta_set_reads(mgta_ctx *ctx, const uint32_t *packed_seq, uint64_t n_words, const uint64_t *start_idx,
                              uint64_t n_reads, uint64_t n_short_reads, int32_t max_read_len) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!packed_seq || !start_idx || n_reads == 0) FAIL(MGTA_ERR_ARG, "set_reads: bad arguments");
    int rc = alloc_reads(ctx, n_words, n_reads, n_short_reads, start_idx[n_reads], max_read_len);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_seq, packed_seq, n_words * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_start, start_idx, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MGTA_OK;
}

extern "C" int mgta_alloc_reads(mgta_ctx *ctx, uint64_t n_words, uint64_t n_reads, uint64_t n_short_reads,
                                uint64_t total_bases, int32_t max_read_len) {
    if (!ctx) return MGTA_ERR_ARG;
    int rc = alloc_reads(ctx, n_words, n_reads, n_short_reads, total_bases, max_read_len);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return MGTA_OK;
}

extern "C" int mgta_reads_device_buffers(mgta_ctx *ctx, void **seq_dev, uint64_t *seq_bytes, void **start_dev,
                                         uint64_t *start_bytes) {
    if (!ctx || !seq_dev || !seq_bytes || !start_dev || !start_bytes) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads or mgta_alloc_reads first");
    *seq_dev = ctx->d_seq; *seq_bytes = ctx->n_words * 4;
    *start_dev = ctx->d_start; *start_bytes = (ctx->n_reads + 1) * 8;
    return MGTA_OK;
}

namespace {

WalkParams walk_params(mgta_ctx *ctx) {
    WalkParams P;
    memset(&P, 0, sizeof(P));
    P.seq = ctx->d_seq; P.start = ctx->d_start; P.n_reads = ctx->n_reads; P.n_short = ctx->n_short;
    P.total_bases = ctx->total_bases; P.k = ctx->opt.kmer_k; P.all_solid = ctx->opt.min_count == 1;
    P.solid = ctx->d_solid; P.hist = ctx->d_hist; P.cursor = ctx->d_cursor; P.n_dollar = ctx->d_totals + 12;
    P.b_lo = 0; P.b_hi = NUM_BUCKETS;
    return P;
}

int begin_timed(mgta_ctx *ctx, int phase) {
    Timed t;
    t.phase = phase;
    CK(cudaEventCreate(&t.a));
    CK(cudaEventCreate(&t.b));
    CK(cudaEventRecord(t.a, ctx->stream));
    ctx->timed.push_back(t);
    return MGTA_OK;
}
int end_timed(mgta_ctx *ctx) {
    CK(cudaEventRecord(ctx->timed.back().b, ctx->stream));
    return MGTA_OK;
}

int histogram(mgta_ctx *ctx, int stage, mgta_stage_stats *st) {
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    CK(cudaSetDevice(ctx->opt.device));
    const int W = stage == 1 ? key_words_s1(ctx->opt.kmer_k) : key_words_s2(ctx->opt.kmer_k);
    WalkParams P = walk_params(ctx);
    const unsigned grid = (unsigned)((ctx->total_bases + WALK_TILE - 1) / WALK_TILE);
    CK(cudaMemsetAsync(ctx->d_hist, 0, NUM_BUCKETS * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals + 12, 0, 8, ctx->stream));
    int rc = begin_timed(ctx, PH_HIST);
    if (rc) return rc;
    if (stage == 1) launch_walk<1, MODE_HIST>(W, P, grid, ctx->stream);
    else launch_walk<2, MODE_HIST>(W, P, grid, ctx->stream);
    CK(cudaGetLastError());
    if ((rc = end_timed(ctx))) return rc;
    if (st) st->n_launches++;
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_hist, NUM_BUCKETS * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_pin + NUM_BUCKETS, ctx->d_totals + 12, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (stage == 2) ctx->n_dollar = ctx->h_pin[NUM_BUCKETS];
    ctx->hist.assign(NUM_BUCKETS, 0);
    uint64_t total = 0;
    for (int b = 0; b < NUM_BUCKETS; ++b) { ctx->hist[b] = (int64_t)ctx->h_pin[b]; total += ctx->h_pin[b]; }
    // shard = contiguous bucket range balanced by item count
    auto boundary = [&](int r) {
        if (r <= 0) return 0;
        if (r >= ctx->opt.world) return (int)NUM_BUCKETS;
        const long double target = (long double)total * r / ctx->opt.world;
        uint64_t acc = 0;
        for (int b = 0; b < NUM_BUCKETS; ++b) {
            if ((long double)acc >= target) return b;
            acc += (uint64_t)ctx->hist[b];
        }
        return (int)NUM_BUCKETS;
    };
    ctx->shard_lo = boundary(ctx->opt.rank);
    ctx->shard_hi = boundary(ctx->opt.rank + 1);
    return MGTA_OK;
}

int finish_timing(mgta_ctx *ctx, mgta_stage_stats *st) {
    for (auto &t : ctx->timed) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t.a, t.b));
        switch (t.phase) {
            case PH_HIST: st->ms_hist += ms; break;
            case PH_EXTRACT: st->ms_extract += ms; break;
            case PH_PARTITION: st->ms_partition += ms; break;
            case PH_SORT: st->ms_sort_emit += ms; break;
            default: break;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    ctx->timed.clear();
    return MGTA_OK;
}

// One stage over this context's bucket shard.
int run_stage(mgta_ctx *ctx, int stage, int64_t *edge_counting, mgta_bucket_sink sink, void *user, int64_t *totals) {
    mgta_stage_stats *st = &ctx->stats[stage - 1];
    memset(st, 0, sizeof(*st));
    cudaEvent_t ev0, ev1;
    CK(cudaSetDevice(ctx->opt.device));
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, ctx->stream));
    int rc = histogram(ctx, stage, st);
    if (rc) return rc;

    Plan pl;
    make_plan(pl, stage, ctx->opt.kmer_k, ctx->opt.sort_items_cap);
    st->key_words = pl.W; st->item_words = pl.IW; st->sort_cap = (int)pl.CAPI;

    uint64_t shard_items = 0, max_bucket = 0;
    for (int b = ctx->shard_lo; b < ctx->shard_hi; ++b) {
        shard_items += (uint64_t)ctx->hist[b];
        max_bucket = std::max<uint64_t>(max_bucket, (uint64_t)ctx->hist[b]);
    }
    st->n_items = shard_items;
    if (max_bucket >= 0xFFFFFFFFull) FAIL(MGTA_ERR_MEM, "a single lv1 bucket holds %llu items (limit 2^32-1)", (unsigned long long)max_bucket);

    CK(cudaMemsetAsync(ctx->d_meta, 0, NUM_BUCKETS * 3 * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals, 0, 10 * 8, ctx->stream));
    if (stage == 1) CK(cudaMemsetAsync(ctx->d_ec, 0, NUM_BUCKETS * 8, ctx->stream));

    // ---- HBM budget -> items per batch
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = ctx->opt.hbm_budget_bytes > 0 ? (size_t)ctx->opt.hbm_budget_bytes : (size_t)(0.9 * (double)(free_b + ctx->arena_bytes));
    {
        uint64_t cap = std::max<uint64_t>(shard_items, 1024);
        while (carve(pl, cap, ctx->n_dollar) > budget) {
            if (cap <= std::max<uint64_t>(max_bucket, 1024)) break;
            cap = std::max<uint64_t>(std::max<uint64_t>(max_bucket, 1024), (uint64_t)((double)cap * 0.9));
        }
        if (carve(pl, cap, ctx->n_dollar) > budget)
            FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the largest lv1 bucket (%llu items, need %zu B)", budget,
                 (unsigned long long)max_bucket, pl.total);
        if (pl.total > ctx->arena_bytes) {
            CK(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->arena);
            ctx->arena = nullptr; ctx->arena_bytes = 0;
            CK(cudaMalloc(&ctx->arena, pl.total));
            ctx->arena_bytes = pl.total;
        }
        uint32_t *bufA = reinterpret_cast<uint32_t *>(ctx->arena + pl.off_a);
        uint32_t *bufB = reinterpret_cast<uint32_t *>(ctx->arena + pl.off_b);
        uint32_t *flags = reinterpret_cast<uint32_t *>(ctx->arena + pl.off_flags);
        uint32_t *win = reinterpret_cast<uint32_t *>(ctx->arena + pl.off_win);
        unsigned long long *state = reinterpret_cast<unsigned long long *>(ctx->arena + pl.off_state);
        Seg *lists[2] = {reinterpret_cast<Seg *>(ctx->arena + pl.off_list0), reinterpret_cast<Seg *>(ctx->arena + pl.off_list1)};
        Giant *giants = reinterpret_cast<Giant *>(ctx->arena + pl.off_giants);
        unsigned char *outbuf = ctx->arena + pl.off_out;

        if (stage == 1) CK(cudaFuncSetAttribute(k_chunk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.chunk_smem));
        else CK(cudaFuncSetAttribute(k_chunk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.chunk_smem));
        int occ = 1;
        if (stage == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chunk<1>, CHUNK_THREADS, pl.chunk_smem));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_chunk<2>, CHUNK_THREADS, pl.chunk_smem));
        occ = std::max(1, occ);

        std::vector<int64_t> meta_host;
        std::vector<Seg> seg_host;
        int b0 = ctx->shard_lo;
        while (b0 < ctx->shard_hi) {
            // ---- batch [b0, b1): greedy prefix of buckets that fits `cap`
            int b1 = b0;
            uint64_t n_items = 0, batch_max = 0;
            while (b1 < ctx->shard_hi && n_items + (uint64_t)ctx->hist[b1] <= pl.cap) {
                n_items += (uint64_t)ctx->hist[b1];
                batch_max = std::max<uint64_t>(batch_max, (uint64_t)ctx->hist[b1]);
                ++b1;
            }
            if (b1 == b0) FAIL(MGTA_ERR_MEM, "bucket %d (%lld items) exceeds the batch capacity %llu", b0, (long long)ctx->hist[b0], (unsigned long long)pl.cap);
            st->n_batches++;
            if (n_items == 0) {
                if (stage == 2 && sink) {
                    meta_host.assign((size_t)(b1 - b0) * 3, 0);
                    if (sink(user, b0, b1, nullptr, 0, meta_host.data()) != 0) FAIL(MGTA_ERR_ARG, "sink aborted");
                }
                b0 = b1;
                continue;
            }
            const unsigned n_windows = (unsigned)((n_items + pl.C - 1) / pl.C);
            // batch-relative bucket offsets -> device cursors
            unsigned long long *h_cur = ctx->h_pin + NUM_BUCKETS;
            {
                uint64_t acc = 0;
                for (int b = 0; b < NUM_BUCKETS; ++b) {
                    h_cur[b] = acc;
                    if (b >= b0 && b < b1) acc += (uint64_t)ctx->hist[b];
                }
            }
            CK(cudaMemcpyAsync(ctx->d_cursor, h_cur, NUM_BUCKETS * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemsetAsync(flags, 0, (n_items / 32 + 64) * 4, ctx->stream));
            CK(cudaMemsetAsync(win, 0, ((size_t)n_windows + 2) * 4, ctx->stream));
            CK(cudaMemsetAsync(state, 0, ((size_t)n_windows + 2) * 8, ctx->stream));
            CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
            k_flags_level0<<<(b1 - b0 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_cursor, ctx->d_hist, b0, b1, flags);
            CK(cudaGetLastError());
            st->n_launches++;
            // ---- extraction + scatter
            WalkParams WP = walk_params(ctx);
            WP.dst = bufA; WP.cap = pl.cap; WP.b_lo = b0; WP.b_hi = b1;
            if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
            const unsigned wgrid = (unsigned)((ctx->total_bases + WALK_TILE - 1) / WALK_TILE);
            if (stage == 1) launch_walk<1, MODE_SCATTER>(pl.W, WP, wgrid, ctx->stream);
            else launch_walk<2, MODE_SCATTER>(pl.W, WP, wgrid, ctx->stream);
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            st->n_launches++;
            // ---- MSD levels when a bucket does not fit the on-chip tile
            const uint32_t *sorted_src = bufA;
            int depth_min = 16;
            if (batch_max > pl.T) {
                seg_host.clear();
                uint64_t acc = 0;
                for (int b = b0; b < b1; ++b) {
                    const uint64_t c = (uint64_t)ctx->hist[b];
                    if (c && (!pl.levels.empty() || c > pl.T)) seg_host.push_back(Seg{acc, (unsigned)c, 0});
                    acc += c;
                }
                // the H2D below reads seg_host asynchronously from pageable memory: CUDA stages it before returning
                CK(cudaMemcpyAsync(lists[0], seg_host.data(), seg_host.size() * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
                const unsigned n0 = (unsigned)seg_host.size();
                CK(cudaMemcpyAsync(ctx->d_ctr + CTR_NLIST0, &n0, 4, cudaMemcpyHostToDevice, ctx->stream));
                if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
                if (pl.levels.empty()) {
                    k_register_giants<<<(n0 + 255) / 256, 256, 0, ctx->stream>>>(lists[0], n0, pl.C, giants, ctx->d_ctr + CTR_NGIANTS,
                                                                              pl.giants_cap, win, ctx->d_ctr + CTR_ERR);
                    CK(cudaGetLastError());
                    st->n_launches++;
                } else {
                    for (size_t l = 0; l < pl.levels.size(); ++l) {
                        MsdParams MP;
                        memset(&MP, 0, sizeof(MP));
                        MP.src = l == 0 ? bufA : bufB;
                        MP.dst = l == 0 ? bufB : bufA;
                        MP.cap = pl.cap; MP.IW = pl.IW;
                        MP.list = lists[l & 1]; MP.n_list = ctx->d_ctr + CTR_NLIST0 + (l & 1);
                        MP.word = pl.levels[l].first >> 5;
                        MP.shift = 32 - (pl.levels[l].first & 31) - pl.levels[l].second;
                        MP.bins = 1 << pl.levels[l].second;
                        MP.flags = flags;
                        MP.next_list = lists[(l + 1) & 1]; MP.next_count = ctx->d_ctr + CTR_NLIST0 + ((l + 1) & 1);
                        MP.next_cap = pl.list_cap;
                        MP.last_level = l + 1 == pl.levels.size();
                        MP.copy_back = l > 0;
                        MP.T = pl.T; MP.C = pl.C;
                        MP.giants = giants; MP.n_giants = ctx->d_ctr + CTR_NGIANTS; MP.giants_cap = pl.giants_cap;
                        MP.win_giant = win; MP.err = ctx->d_ctr + CTR_ERR;
                        CK(cudaMemsetAsync(ctx->d_ctr + CTR_NLIST0 + ((l + 1) & 1), 0, 4, ctx->stream));
                        const unsigned grid = l == 0 ? std::min<unsigned>(std::max(1u, n0), (unsigned)ctx->sm_count * 8) : (unsigned)ctx->sm_count * 4;
                        k_msd<<<grid, MSD_THREADS, 0, ctx->stream>>>(MP);
                        CK(cudaGetLastError());
                        st->n_launches++;
                        st->msd_levels = std::max<int>(st->msd_levels, (int)l + 1);
                    }
                    sorted_src = bufB;
                    depth_min = 16 + pl.levels[0].second;
                }
                if ((rc = end_timed(ctx))) return rc;
            }
            // ---- on-chip sort + count / emit
            ChunkParams CP;
            memset(&CP, 0, sizeof(CP));
            CP.src = sorted_src; CP.cap = pl.cap; CP.n_items = n_items; CP.W = pl.W; CP.IW = pl.IW; CP.k = pl.k;
            CP.CAPI = pl.CAPI; CP.C = pl.C; CP.n_windows = n_windows; CP.ticket = ctx->d_ctr + CTR_TICKET;
            CP.flags = flags; CP.win_giant = win; CP.giants = giants; CP.depth_min = depth_min;
            CP.n_pass = pl.n_pass;
            memcpy(CP.pass_lsb, pl.pass_lsb, sizeof(CP.pass_lsb));
            memcpy(CP.pass_nb, pl.pass_nb, sizeof(CP.pass_nb));
            CP.g_full = (pl.k - 1) / 16;
            CP.g_rem_shift = (pl.k - 1) % 16 ? (16 - (pl.k - 1) % 16) * 2 : 32;
            CP.solid = ctx->d_solid; CP.edge_counting = ctx->d_ec; CP.m = (unsigned)ctx->opt.min_count;
            CP.aw = (pl.k - 1) >> 4; CP.ash = (15 - ((pl.k - 1) & 15)) * 2; CP.wpt = (2 * pl.k + 31) / 32;
            CP.out = outbuf; CP.out_cap = pl.out_cap; CP.state = state; CP.meta = ctx->d_meta; CP.totals = ctx->d_totals;
            CP.err = ctx->d_ctr + CTR_ERR;
            const unsigned cgrid = std::min<unsigned>(n_windows, (unsigned)(ctx->sm_count * occ));
            if ((rc = begin_timed(ctx, PH_SORT))) return rc;
            if (stage == 1) k_chunk<1><<<cgrid, CHUNK_THREADS, pl.chunk_smem, ctx->stream>>>(CP);
            else k_chunk<2><<<cgrid, CHUNK_THREADS, pl.chunk_smem, ctx->stream>>>(CP);
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            st->n_launches++;
            // ---- batch epilogue: error flags, giants, output
            unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
            unsigned long long *h_state = ctx->h_pin + 2 * NUM_BUCKETS + 8;
            CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (stage == 2) CK(cudaMemcpyAsync(h_state, state + (n_windows - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            st->n_giants += h_ctr[CTR_NGIANTS];
            const unsigned dev_err = h_ctr[CTR_ERR];
            if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (stage %d, buckets [%d,%d))", dev_err, stage, b0, b1);
            if (stage == 2) {
                const unsigned long long bytes = *h_state & ((1ull << 62) - 1);
                st->out_bytes += bytes;
                if (sink) {
                    if (bytes > ctx->h_out_bytes) {
                        cudaFreeHost(ctx->h_out);
                        ctx->h_out = nullptr; ctx->h_out_bytes = 0;
                        CK(cudaHostAlloc(&ctx->h_out, bytes + bytes / 4 + 4096, cudaHostAllocDefault));
                        ctx->h_out_bytes = bytes + bytes / 4 + 4096;
                    }
                    meta_host.resize((size_t)(b1 - b0) * 3);
                    if (bytes) CK(cudaMemcpyAsync(ctx->h_out, outbuf, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaMemcpyAsync(meta_host.data(), ctx->d_meta + (size_t)b0 * 3, (size_t)(b1 - b0) * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream));
                    if (sink(user, b0, b1, ctx->h_out, bytes, meta_host.data()) != 0) FAIL(MGTA_ERR_ARG, "sink aborted");
                }
            }
            b0 = b1;
        }
    }
    // ---- stage epilogue
    if (stage == 1 && edge_counting) {
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_ec, NUM_BUCKETS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < NUM_BUCKETS; ++i) edge_counting[i] = (int64_t)ctx->h_pin[i];
    }
    if (stage == 2) {
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals, 10 * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        uint64_t edges = 0;
        for (int i = 0; i < 9; ++i) edges += ctx->h_pin[i];
        st->n_edges = edges;
        if (totals) for (int i = 0; i < 10; ++i) totals[i] = (int64_t)ctx->h_pin[i];
    }
    CK(cudaEventRecord(ev1, ctx->stream));
    CK(cudaEventSynchronize(ev1));
    CK(cudaEventElapsedTime(&st->ms_total, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return finish_timing(ctx, st);
}

}  // namespace

extern "C" int mgta_stage1_histogram(mgta_ctx *ctx, int64_t *hist) {
    if (!ctx || !hist) return MGTA_ERR_ARG;
    int rc = histogram(ctx, 1, nullptr);
    if (rc) return rc;
    memcpy(hist, ctx->hist.data(), NUM_BUCKETS * 8);
    mgta_stage_stats tmp;
    memset(&tmp, 0, sizeof(tmp));
    return finish_timing(ctx, &tmp);
}

extern "C" int mgta_stage2_histogram(mgta_ctx *ctx, int64_t *hist) {
    if (!ctx || !hist) return MGTA_ERR_ARG;
    int rc = histogram(ctx, 2, nullptr);
    if (rc) return rc;
    memcpy(hist, ctx->hist.data(), NUM_BUCKETS * 8);
    mgta_stage_stats tmp;
    memset(&tmp, 0, sizeof(tmp));
    return finish_timing(ctx, &tmp);
}

extern "C" int mgta_stage1(mgta_ctx *ctx, int64_t *edge_counting) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    if (ctx->opt.min_count == 1) {
        if (edge_counting) memset(edge_counting, 0, NUM_BUCKETS * 8);
        memset(&ctx->stats[0], 0, sizeof(ctx->stats[0]));
        return MGTA_OK;
    }
    if (ctx->opt.need_mercy) FAIL(MGTA_ERR_ARG, "need_mercy is not implemented on the device yet");
    CK(cudaSetDevice(ctx->opt.device));
    CK(cudaMemsetAsync(ctx->d_solid, 0, ctx->solid_words * 4, ctx->stream));
    return run_stage(ctx, 1, edge_counting, nullptr, nullptr, nullptr);
}

extern "C" int mgta_stage2(mgta_ctx *ctx, mgta_bucket_sink sink, void *user, int64_t *totals) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    return run_stage(ctx, 2, nullptr, sink, user, totals);
}

extern "C" int mgta_solid_device_buffer(mgta_ctx *ctx, void **dev_ptr, uint64_t *n_bytes) {
    if (!ctx || !dev_ptr || !n_bytes) return MGTA_ERR_ARG;
    if (!ctx->d_solid) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    *dev_ptr = ctx->d_solid;
    *n_bytes = ctx->solid_words * 4;
    return MGTA_OK;
}

extern "C" int mgta_get_is_solid(mgta_ctx *ctx, uint8_t *host, uint64_t n_bytes) {
    if (!ctx || !host) return MGTA_ERR_ARG;
    if (!ctx->d_solid) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    const int nk1 = ctx->max_len - ctx->opt.kmer_k;
    const uint64_t bits = nk1 > 0 ? (uint64_t)nk1 * ctx->n_short : 0;
    if (n_bytes < (bits + 7) / 8) FAIL(MGTA_ERR_ARG, "get_is_solid: buffer too small");
    CK(cudaSetDevice(ctx->opt.device));
    const uint64_t words = (bits + 31) / 32 + 1;
    uint32_t *tmp = nullptr;
    CK(cudaMalloc(&tmp, words * 4));
    CK(cudaMemsetAsync(tmp, 0, words * 4, ctx->stream));
    if (bits) {
        k_solid_export<<<(unsigned)((ctx->n_short + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_solid, ctx->d_start, ctx->n_short,
                                                                                      ctx->opt.kmer_k, nk1, tmp);
        CK(cudaGetLastError());
    }
    std::vector<uint32_t> h(words);
    CK(cudaMemcpyAsync(h.data(), tmp, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    memset(host, 0, n_bytes);
    memcpy(host, h.data(), std::min<uint64_t>(n_bytes, (bits + 7) / 8));
    return MGTA_OK;
}

extern "C" int mgta_set_is_solid(mgta_ctx *ctx, const uint8_t *host, uint64_t n_bytes) {
    if (!ctx || !host) return MGTA_ERR_ARG;
    if (!ctx->d_solid) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    const int nk1 = ctx->max_len - ctx->opt.kmer_k;
    const uint64_t bits = nk1 > 0 ? (uint64_t)nk1 * ctx->n_short : 0;
    if (n_bytes < (bits + 7) / 8) FAIL(MGTA_ERR_ARG, "set_is_solid: buffer too small");
    CK(cudaSetDevice(ctx->opt.device));
    const uint64_t words = (bits + 31) / 32 + 1;
    std::vector<uint32_t> h(words, 0);
    memcpy(h.data(), host, (bits + 7) / 8);
    uint32_t *tmp = nullptr;
    CK(cudaMalloc(&tmp, words * 4));
    CK(cudaMemcpyAsync(tmp, h.data(), words * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_solid, 0, ctx->solid_words * 4, ctx->stream));
    if (bits) {
        k_solid_import<<<(unsigned)((ctx->n_short + 255) / 256), 256, 0, ctx->stream>>>(tmp, ctx->d_start, ctx->n_short,
                                                                                      ctx->opt.kmer_k, nk1, ctx->d_solid);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return MGTA_OK;
}

extern "C" int mgta_get_mercy_candidates(mgta_ctx *ctx, uint64_t *host, uint64_t cap, uint64_t *n) {
    if (!ctx || !n) return MGTA_ERR_ARG;
    (void)host; (void)cap;
    *n = 0;
    FAIL(MGTA_ERR_ARG, "need_mercy is not implemented on the device yet");
}

extern "C" int mgta_shard_range(mgta_ctx *ctx, int32_t *bucket_begin, int32_t *bucket_end) {
    if (!ctx || !bucket_begin || !bucket_end) return MGTA_ERR_ARG;
    *bucket_begin = ctx->shard_lo;
    *bucket_end = ctx->shard_hi;
    return MGTA_OK;
}

extern "C" int mgta_get_stats(mgta_ctx *ctx, int stage, mgta_stage_stats *out) {
    if (!ctx || !out || stage < 1 || stage > 2) return MGTA_ERR_ARG;
    *out = ctx->stats[stage - 1];
    return MGTA_OK;
}
