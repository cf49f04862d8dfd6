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
#include "v2_kernels.cuh"
#include "emit_kernels.cuh"
#include "mercy_kernels.cuh"
#include "node_kernels.cuh"

using namespace mgta;

namespace {

std::string g_create_error;

enum { CTR_TICKET = 0, CTR_NLIST0 = 1, CTR_NLIST1 = 2, CTR_NGIANTS = 3, CTR_ERR = 4, CTR_TICKET2 = 5, CTR_TICKET3 = 6, CTR_NOVF = 7,
       CTR_NOVF2 = 8, CTR_COUNT = 12 };
enum { PH_HIST = 0, PH_EXTRACT = 1, PH_PARTITION = 2, PH_SORT = 3, PH_NODES = 4, PH_COUNT = 5 };

struct Timed {
    int phase;
    cudaEvent_t a, b;
};

}  // namespace

namespace {
struct CountPlan {
    int k, WE, PW, IW, bits;
    bool plus, has_assist, stage1_mode, mark_mode;
    unsigned tab_cap, big_cap, tab_limit, lb1, lb2, B1, T, r_lo, r_hi;
    uint64_t n_pos;
};

// scan-sharded stage 1 between mgta_stage1_scan and mgta_stage1_count
struct ExchangeState {
    bool valid = false;
    CountPlan cp;
    uint64_t slab_items = 0;
    size_t send_off = 0, recv_off = 0, xbytes = 0;      // arena offsets of the send / receive slabs, bytes of each
    std::vector<uint64_t> send_counts;
    unsigned round = 0, n_rounds = 1;
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
    bool sink_device = false;              // stage 2 hands the sink DEVICE pointers (mgta_stage2_into_sdbg): no D2H of the records
    uint32_t *d_lut = nullptr;             // read lookup table (k_build_read_lut), built on first use after the reads change
    uint64_t n_lut = 0;
    bool lut_valid = false;
    // small device state
    unsigned long long *d_hist = nullptr, *d_cursor = nullptr, *d_meta = nullptr, *d_totals = nullptr, *d_ec = nullptr, *d_ec_bak = nullptr;
    unsigned *d_ctr = nullptr;
    unsigned long long *h_pin = nullptr;   // pinned: hist / cursor staging [2 * 65536] + misc [64]
    unsigned char *h_out = nullptr;        // pinned output staging
    size_t h_out_bytes = 0;
    // arena
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t budget_cached = 0;              // see hbm_budget()
    // results
    std::vector<int64_t> hist;             // last histogram (whole bucket space)
    int shard_lo = 0, shard_hi = NUM_BUCKETS;
    mgta_stage_stats stats[2];
    std::vector<Timed> timed;
    uint64_t n_dollar = 0;                 // stage-2 items with a == $ (from the last stage-2 histogram)
    // edge-centric state: {(canonical (k+1)-mer, multiplicity)} rows of edge_row_words u32, and the histogram of the
    // stage-2 items they generate by PB-bit key prefix
    uint32_t *d_edges = nullptr;
    uint64_t n_edges = 0, edges_cap = 0;
    // world > 1: the stage-2 items this shard received from all shards (sharded.inc): `world` slabs of slab_items items at
    // the start of the arena, counts per source; valid until the edges change
    struct ItemSlabs {
        bool valid = false;
        uint64_t slab_items = 0, n_dollar = 0;
        size_t bytes = 0;
        std::vector<uint64_t> counts;
    } s2x;
    unsigned long long *d_xs = nullptr;    // scan-sharded exchange: input regions of the level-1 split
    int edge_row_words = 0;
    bool edges_valid = false;
    bool edges_complete = false;           // d_edges holds the edges of the WHOLE hash space (one shard, or the general counting mode)
    bool solid_valid = false;              // d_solid holds the is_solid vector of the last stage 1 (derived on demand)
    bool stage1_done = false;
    uint32_t *d_hist_s2 = nullptr;
    int PB = 19;                           // prefix bits of a stage-2 tile: min(19 + log2(world), 24, 2(k-1)), >= 16
    uint32_t *h_hist2 = nullptr;           // pinned copy of d_hist_s2
    uint64_t n_positions = 0;              // edge offsets over all reads
    bool n_positions_valid = false;
    ExchangeState xch;
    uint64_t slab_suggest = 0;             // cached mgta_stage1_slab_items() (0 = not computed for the current reads)
    // asynchronous upload (mgta_set_reads_async): start_idx first, then packed_seq in chunks on a copy stream; stage 1
    // extracts chunk c as soon as chunk c + 1 has landed
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    std::vector<cudaEvent_t> ev_out;       // stage 2: one per part of a batch's output (D2H pipelined behind the sort)
    std::vector<uint64_t> chunk_end;       // end base of each chunk (multiples of 16384 except the last = total_bases)
    bool copy_pending = false;
    // mercy (need_mercy): candidates of the last stage 1 and the number of is_solid bits the per-read scan added
    unsigned long long *d_cand = nullptr;
    uint64_t cand_cap = 0, n_cand = 0, num_mercy = 0;
    bool mercy_valid = false;
    // node pass (stage 2): the $-items of the tip k-mers, rows of key_words_s2 + 1 words; see node_kernels.cuh
    bool node_pass = false;
    uint32_t *d_tips = nullptr;
    uint64_t n_tips = 0, tips_cap = 0;
    bool tips_valid = false;
    uint32_t *d_hist_bak = nullptr;        // d_hist_s2 before the node pass added the tip items (restored when the pass restarts)
    // sharded build (mgta_sharded_*): the protocol state between two collectives
    struct Sharded {
        int stage = 0, state = 0, attempt = 0;
        bool timing = false;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        mgta_bucket_sink sink = nullptr;
        void *user = nullptr;
        unsigned long long *d_small = nullptr, *h_small = nullptr;   // world * (world + 2) u64: counts tables
        uint64_t slab = 0, n_ops_all = 0;
        unsigned round = 0, n_rounds = 1;                      // stage-1 exchange rounds (HBM-limited inputs)
        std::vector<uint64_t> recv_counts, expect;
        std::vector<unsigned> bnd;                              // first lv1 bucket of every shard (world + 1 entries)
        std::vector<int64_t> ec, totals;
    } sh;
};

// Host waits on the context's stream.  A blocking cudaStreamSynchronize puts the thread to sleep; on a loaded host the
// wake-up can cost milliseconds while the GPU sits idle, and a step has ~8 such points.  Polling keeps the thread hot.
static inline cudaError_t mgta_stream_wait(cudaStream_t s) {
    static const int spin = [] { const char *e = getenv("MGTA_SPIN_SYNC"); return e ? atoi(e) : 1; }();
    if (!spin) return cudaStreamSynchronize(s);
    cudaError_t e;
    while ((e = cudaStreamQuery(s)) == cudaErrorNotReady) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    return e;
}

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

#define W_SWITCH(W, ...)                                           \
    switch (W) {                                                   \
        case 1: { constexpr int WW = 1; __VA_ARGS__; } break;      \
        case 2: { constexpr int WW = 2; __VA_ARGS__; } break;      \
        case 3: { constexpr int WW = 3; __VA_ARGS__; } break;      \
        case 4: { constexpr int WW = 4; __VA_ARGS__; } break;      \
        case 5: { constexpr int WW = 5; __VA_ARGS__; } break;      \
        case 6: { constexpr int WW = 6; __VA_ARGS__; } break;      \
        case 7: { constexpr int WW = 7; __VA_ARGS__; } break;      \
        case 8: { constexpr int WW = 8; __VA_ARGS__; } break;      \
        case 9: { constexpr int WW = 9; __VA_ARGS__; } break;      \
        default: break;                                            \
    }

#define W_SWITCH3(W, ...)                                          \
    switch (W) {                                                   \
        case 1: { constexpr int WW = 1; __VA_ARGS__; } break;      \
        case 2: { constexpr int WW = 2; __VA_ARGS__; } break;      \
        case 3: { constexpr int WW = 3; __VA_ARGS__; } break;      \
        default: break;                                            \
    }

template <int STAGE, int MODE>
void launch_walk(int W, const WalkParams &P, unsigned grid, cudaStream_t st) {
    W_SWITCH(W, (k_walk<WW, STAGE, MODE><<<grid, WALK_THREADS, 0, st>>>(P)));
}

struct Plan {
    int stage, W, IW, k;
    unsigned CAPI, T, C;
    size_t chunk_smem;
    // arena carve (byte offsets)
    uint64_t cap;                              // items per batch
    size_t off_a, off_b, off_flags, off_win, off_state, off_list0, off_list1, off_giants, off_out, total;
    unsigned list_cap, giants_cap;
    uint64_t out_cap;
};

void make_plan(Plan &pl, int stage, int k, int cap_override) {
    pl.stage = stage;
    pl.k = k;
    pl.W = stage == 1 ? key_words_s1(k) : key_words_s2(k);
    pl.IW = pl.W + (stage == 1 ? 2 : 1);       // stage 2: key + multiplicity of the generating edge
    // on-chip window: two CTAs per SM (<= ~110 KB dynamic shared memory each), at most 4096 items (8 per thread)
    unsigned capi = 4096;
    while (capi > 512 && sort_emit_smem_bytes(pl.IW, capi) > 110 * 1024) capi -= 512;
    if (cap_override > 0) capi = std::min(4096, std::max(128, (cap_override / 64) * 64));   // groups hold <= 48 items: T >= 64
    pl.CAPI = capi;
    pl.T = capi / 2;                                               // largest leaf the MSD levels leave alone
    pl.C = capi / 2;                                               // window stride: C + largest leaf (>= one 48-item group) <= CAPI
    if (cap_override > 0) pl.T = std::max(16, std::min<int>(cap_override, (int)capi) / 2);
    pl.chunk_smem = sort_emit_smem_bytes(pl.IW, capi);
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
    // stage-2 record stream, guaranteed bound: every record is a run of >= 1 items (2 B); the items are WEIGHTED (one
    // item per distinct edge carrying its multiplicity), so any single item can need the u16 multiplicity (2 B more);
    // a tip label needs an item with a == $ (n_dollar bounds them)
    const int wpt = (2 * pl.k + 31) / 32;
    pl.out_cap = pl.stage == 2 ? 4 * cap + 4ull * wpt * std::min<uint64_t>(cap, n_dollar) + 4096 : 0;
    o += al(pl.out_cap);
    pl.total = o;
    return o;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int mgta_abi_version(void) { return 3; }

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
    if ((e = cudaMalloc(&ctx->d_totals, 32 * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_ec, NUM_BUCKETS * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_ec_bak, NUM_BUCKETS * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_ctr, CTR_COUNT * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    // stage-2 prefix tiles: 2^19 for one GPU; the item count grows with the shards (weak scaling), so each doubling of the
    // world adds a prefix bit and the tiles keep their size (measured at 8 GPUs with 2^20 tiles: every tile overflows
    // the on-chip window, MSD levels run for all of them and the sort slows from 67 to 96 ms).
    {
        int wb = 0;
        while ((1 << wb) < opts->world) ++wb;
        // 2^19 tiles per shard: 512 level-1 bins (k_item_part writes runs twice as long as with 1024: 22.1 -> 14.7 ms at
        // 20M x 150) and room below the 1024-bin batch limit for unevenly wide shard ranges; the few tiles that then exceed
        // the on-chip window take the MSD levels (+3.5 ms)
        ctx->PB = std::min(std::min(19 + wb, 24), 2 * (opts->kmer_k - 1));
        if (const char *e2 = getenv("MGTA_S2_PB")) ctx->PB = std::max(16, std::min(std::min(atoi(e2), 24), 2 * (opts->kmer_k - 1)));   // test hook
    }
    if ((e = cudaMalloc(&ctx->d_hist_s2, ((size_t)1 << ctx->PB) * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_hist_bak, ((size_t)1 << ctx->PB) * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    // stage 2 takes its $-items from the node pass (a third of the items to partition and sort).  MGTA_NODE_PASS=0: every
    // edge generates all six items and the group logic drops the covered $-items (one shard only; A/B switch)
    ctx->node_pass = true;
    if (const char *e3 = getenv("MGTA_NODE_PASS")) ctx->node_pass = atoi(e3) != 0 || opts->world > 1;
    if ((e = cudaHostAlloc(&ctx->h_hist2, ((size_t)1 << ctx->PB) * 4, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    if ((e = cudaMalloc(&ctx->d_xs, (size_t)(MAX_OWNERS + 1) * 24)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&ctx->h_pin, (2 * NUM_BUCKETS + 64) * 8, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    memset(ctx->h_pin, 0, (2 * NUM_BUCKETS + 64) * 8);
    if (opts->world > 1) {
        if (opts->world > MAX_OWNERS) { g_create_error = "at most 16 shards"; delete ctx; return MGTA_ERR_ARG; }
        const size_t small = (size_t)opts->world * (opts->world + 2) * 8;
        if ((e = cudaMalloc(&ctx->sh.d_small, small)) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaHostAlloc(&ctx->sh.h_small, small, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    }
    *out = ctx;
    return MGTA_OK;
}

extern "C" void mgta_ctx_destroy(mgta_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->opt.device);
    mgta_stream_wait(ctx->stream);
    cudaFree(ctx->d_seq); cudaFree(ctx->d_start); cudaFree(ctx->d_solid); cudaFree(ctx->d_lut);
    cudaFree(ctx->d_hist); cudaFree(ctx->d_cursor); cudaFree(ctx->d_meta); cudaFree(ctx->d_totals); cudaFree(ctx->d_ec); cudaFree(ctx->d_ec_bak);
    cudaFree(ctx->d_ctr); cudaFree(ctx->arena); cudaFree(ctx->d_edges); cudaFree(ctx->d_hist_s2);
    cudaFree(ctx->d_xs); cudaFree(ctx->d_cand); cudaFree(ctx->d_tips); cudaFree(ctx->d_hist_bak);
    cudaFreeHost(ctx->h_pin); cudaFreeHost(ctx->h_out); cudaFreeHost(ctx->h_hist2);
    cudaFree(ctx->sh.d_small); cudaFreeHost(ctx->sh.h_small);
    if (ctx->sh.ev0) cudaEventDestroy(ctx->sh.ev0);
    if (ctx->sh.ev1) cudaEventDestroy(ctx->sh.ev1);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (auto e : ctx->ev_chunk) cudaEventDestroy(e);
    for (auto e : ctx->ev_out) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

namespace {
// everything but the pipelined stage-1 extraction sees the reads only once the whole upload has landed
int wait_reads(mgta_ctx *ctx) {
    if (ctx->copy_pending) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk.back(), 0));
        ctx->copy_pending = false;
    }
    return MGTA_OK;
}

int alloc_reads(mgta_ctx *ctx, uint64_t n_words, uint64_t n_reads, uint64_t n_short, uint64_t total, int max_len) {
    if (n_reads == 0 || n_short > n_reads) FAIL(MGTA_ERR_ARG, "reads: bad counts");
    if (total == 0) FAIL(MGTA_ERR_ARG, "reads: no bases");
    if (n_words * 16 < total) FAIL(MGTA_ERR_ARG, "reads: packed_seq shorter than start_idx says");
    if (total >= (1ull << 40) - 1) FAIL(MGTA_ERR_ARG, "reads: more than 2^40 bases");
    if (n_reads >= 0xFFFFFFFFull) FAIL(MGTA_ERR_ARG, "reads: more than 2^32 - 2 reads");
    CK(cudaSetDevice(ctx->opt.device));
    if (ctx->copy_pending) {                                       // a previous asynchronous upload still owns the buffers
        CK(cudaStreamSynchronize(ctx->copy_stream));
        ctx->copy_pending = false;
    }
    const uint64_t padded = ((n_words + 3) & ~3ull) + SEQ_PAD_WORDS;
    const uint64_t solid_words = (total + 31) / 32 + 4;
    if (n_words != ctx->n_words || n_reads != ctx->n_reads || total != ctx->total_bases || !ctx->d_seq) {     // reuse buffers of equal shape
        cudaFree(ctx->d_seq); cudaFree(ctx->d_start); cudaFree(ctx->d_solid); cudaFree(ctx->d_lut);
        ctx->d_seq = nullptr; ctx->d_start = nullptr; ctx->d_solid = nullptr; ctx->d_lut = nullptr;
        CK(cudaMalloc(&ctx->d_seq, padded * 4));
        CK(cudaMalloc(&ctx->d_start, (n_reads + 1) * 8));
        CK(cudaMalloc(&ctx->d_solid, solid_words * 4));
        CK(cudaMalloc(&ctx->d_lut, ((total >> READ_LUT_SHIFT) + 2) * 4));
        ctx->budget_cached = 0;
    }
    ctx->solid_words = solid_words;
    CK(cudaMemsetAsync(ctx->d_seq + n_words, 0, (padded - n_words) * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_solid, 0, ctx->solid_words * 4, ctx->stream));
    ctx->n_words = n_words; ctx->n_reads = n_reads; ctx->n_short = n_short; ctx->total_bases = total;
    ctx->max_len = max_len;
    ctx->edges_valid = false;
    ctx->s2x.valid = false;
    ctx->solid_valid = false;
    ctx->stage1_done = false;
    ctx->n_positions_valid = false;
    ctx->n_lut = (total >> READ_LUT_SHIFT) + 2;
    ctx->lut_valid = false;
    ctx->xch.valid = false;
    ctx->slab_suggest = 0;
    return MGTA_OK;
}

// the read lookup table of the current start_idx (start_idx has landed on the device or its copy is ordered before
// ev_start); every kernel that walks base positions takes it
int ensure_read_lut(mgta_ctx *ctx) {
    if (ctx->lut_valid) return MGTA_OK;
    if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_start, 0));
    k_build_read_lut<<<(unsigned)((ctx->n_lut + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_start, ctx->n_reads, ctx->n_lut, ctx->d_lut);
    CK(cudaGetLastError());
    ctx->lut_valid = true;
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
    CK(mgta_stream_wait(ctx->stream));
    return MGTA_OK;
}

extern "C" int mgta_set_reads_async(mgta_ctx *ctx, const uint32_t *packed_seq, uint64_t n_words, const uint64_t *start_idx,
                                    uint64_t n_reads, uint64_t n_short_reads, int32_t max_read_len) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!packed_seq || !start_idx || n_reads == 0) FAIL(MGTA_ERR_ARG, "set_reads: bad arguments");
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, packed_seq) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    constexpr int NCH = 8;
    constexpr uint64_t ALIGN_WORDS = 1024;                         // 16384 bases: a multiple of every extraction tile
    if (!pinned || n_words < NCH * ALIGN_WORDS * 64)               // pageable memory copies synchronously anyway; tiny inputs
        return mgta_set_reads(ctx, packed_seq, n_words, start_idx, n_reads, n_short_reads, max_read_len);
    int rc = alloc_reads(ctx, n_words, n_reads, n_short_reads, start_idx[n_reads], max_read_len);
    if (rc) return rc;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        ctx->ev_chunk.resize(NCH);
        for (auto &e : ctx->ev_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));               // the copies follow the buffer set-up on the main stream
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_start, 0));
    CK(cudaMemcpyAsync(ctx->d_start, start_idx, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_start, ctx->copy_stream));
    ctx->chunk_end.assign(NCH, 0);
    const uint64_t per = ((n_words + NCH - 1) / NCH + ALIGN_WORDS - 1) / ALIGN_WORDS * ALIGN_WORDS;
    for (int c = 0; c < NCH; ++c) {
        const uint64_t w0 = std::min<uint64_t>(n_words, (uint64_t)c * per), w1 = c + 1 == NCH ? n_words : std::min<uint64_t>(n_words, w0 + per);
        if (w1 > w0) CK(cudaMemcpyAsync(ctx->d_seq + w0, packed_seq + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
        ctx->chunk_end[c] = c + 1 == NCH ? ctx->total_bases : std::min<uint64_t>(ctx->total_bases, w1 * 16);
    }
    ctx->copy_pending = true;
    return MGTA_OK;
}

extern "C" int mgta_alloc_reads(mgta_ctx *ctx, uint64_t n_words, uint64_t n_reads, uint64_t n_short_reads,
                                uint64_t total_bases, int32_t max_read_len) {
    if (!ctx) return MGTA_ERR_ARG;
    int rc = alloc_reads(ctx, n_words, n_reads, n_short_reads, total_bases, max_read_len);
    if (rc) return rc;
    CK(mgta_stream_wait(ctx->stream));
    return MGTA_OK;
}

extern "C" int mgta_reads_device_buffers(mgta_ctx *ctx, void **seq_dev, uint64_t *seq_bytes, void **start_dev,
                                         uint64_t *start_bytes) {
    if (!ctx || !seq_dev || !seq_bytes || !start_dev || !start_bytes) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads or mgta_alloc_reads first");
    if (ctx->copy_pending) { CK(cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }   // the caller's streams are not ours
    *seq_dev = ctx->d_seq; *seq_bytes = ctx->n_words * 4;
    *start_dev = ctx->d_start; *start_bytes = (ctx->n_reads + 1) * 8;
    return MGTA_OK;
}

namespace {

WalkParams walk_params(mgta_ctx *ctx) {
    WalkParams P;
    memset(&P, 0, sizeof(P));
    P.seq = ctx->d_seq; P.start = ctx->d_start; P.lut = ctx->d_lut; P.n_lut = ctx->n_lut; P.n_reads = ctx->n_reads; P.n_short = ctx->n_short;
    P.total_bases = ctx->total_bases; P.k = ctx->opt.kmer_k; P.all_solid = ctx->opt.min_count == 1;
    P.solid = ctx->d_solid; P.hist = ctx->d_hist; P.n_dollar = ctx->d_totals + 12;
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

void set_shard_range(mgta_ctx *ctx, uint64_t total);

int histogram(mgta_ctx *ctx, int stage, mgta_stage_stats *st) {
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    CK(cudaSetDevice(ctx->opt.device));
    { int rcw = wait_reads(ctx); if (rcw) return rcw; }
    { int rcl = ensure_read_lut(ctx); if (rcl) return rcl; }
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
    CK(mgta_stream_wait(ctx->stream));
    if (stage == 2) ctx->n_dollar = ctx->h_pin[NUM_BUCKETS];
    ctx->hist.assign(NUM_BUCKETS, 0);
    uint64_t total = 0;
    for (int b = 0; b < NUM_BUCKETS; ++b) { ctx->hist[b] = (int64_t)ctx->h_pin[b]; total += ctx->h_pin[b]; }
    set_shard_range(ctx, total);
    return MGTA_OK;
}

// shard = contiguous bucket range balanced by item count, from ctx->hist: first bucket of shard r
int shard_boundary(const mgta_ctx *ctx, uint64_t total, int r) {
    if (r <= 0) return 0;
    if (r >= ctx->opt.world) return (int)NUM_BUCKETS;
    const long double target = (long double)total * r / ctx->opt.world;
    uint64_t acc = 0;
    for (int b = 0; b < NUM_BUCKETS; ++b) {
        if ((long double)acc >= target) return b;
        acc += (uint64_t)ctx->hist[b];
    }
    return (int)NUM_BUCKETS;
}

void set_shard_range(mgta_ctx *ctx, uint64_t total) {
    ctx->shard_lo = shard_boundary(ctx, total, ctx->opt.rank);
    ctx->shard_hi = shard_boundary(ctx, total, ctx->opt.rank + 1);
}

// destroys every pending phase timer (error paths, and the start of the next stage after one)
void drop_timed(mgta_ctx *ctx) {
    for (auto &t : ctx->timed) {
        if (t.a) cudaEventDestroy(t.a);
        if (t.b) cudaEventDestroy(t.b);
    }
    ctx->timed.clear();
}

int finish_timing(mgta_ctx *ctx, mgta_stage_stats *st) {
    for (auto &t : ctx->timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) {  // e.g. an event that was never recorded: nothing to report, nothing to keep
            cudaGetLastError();
            drop_timed(ctx);
            FAIL(MGTA_ERR_CUDA, "phase timer could not be read");
        }
        switch (t.phase) {
            case PH_HIST: st->ms_hist += ms; break;
            case PH_EXTRACT: st->ms_extract += ms; break;
            case PH_PARTITION: st->ms_partition += ms; break;
            case PH_SORT: st->ms_sort_emit += ms; break;
            case PH_NODES: st->ms_nodes += ms; break;
            default: break;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    ctx->timed.clear();
    return MGTA_OK;
}

// ---- small helpers -------------------------------------------------------------------------------
struct Carver {
    size_t o = 0;
    size_t take(size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; }
};

// HBM budget of the context: --gpu_mem, or 90 % of what is free plus what the arena already holds.  cudaMemGetInfo goes
// through the kernel driver and was measured to stall for tens of milliseconds on a busy host -- with the GPU idle, at the
// start of every stage -- so the value is cached and refreshed only when the arena has to grow (the stream is drained
// there anyway).
size_t hbm_budget(mgta_ctx *ctx, bool refresh = false) {
    if (ctx->opt.hbm_budget_bytes > 0) return (size_t)ctx->opt.hbm_budget_bytes;
    if (!ctx->budget_cached || refresh) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        ctx->budget_cached = (size_t)(0.9 * (double)(free_b + ctx->arena_bytes));
    }
    return ctx->budget_cached;
}

int ensure_arena(mgta_ctx *ctx, size_t bytes) {
    if (bytes > ctx->arena_bytes) {
        CK(mgta_stream_wait(ctx->stream));
        cudaFree(ctx->arena);
        ctx->arena = nullptr; ctx->arena_bytes = 0;
        if (cudaMalloc(&ctx->arena, bytes) != cudaSuccess) {
            cudaGetLastError();
            const size_t now = hbm_budget(ctx, true);      // the cached budget was stale (other buffers grew since)
            FAIL(MGTA_ERR_MEM, "cannot allocate a %zu B arena (budget now %zu B): call again, the plan will use the smaller budget", bytes, now);
        }
        ctx->arena_bytes = bytes;
        hbm_budget(ctx, true);
    }
    return MGTA_OK;
}

int count_positions(mgta_ctx *ctx) {
    { int rcl = ensure_read_lut(ctx); if (rcl) return rcl; }
    if (ctx->n_positions_valid) return MGTA_OK;
    if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_start, 0));     // start_idx has landed; packed_seq may not have
    CK(cudaMemsetAsync(ctx->d_totals + 13, 0, 8, ctx->stream));
    k_count_positions<<<(unsigned)((ctx->n_reads + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_start, 0, ctx->n_reads, ctx->opt.kmer_k,
                                                                                     ctx->d_totals + 13);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals + 13, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    ctx->n_positions = ctx->h_pin[0];
    ctx->n_positions_valid = true;
    return MGTA_OK;
}

#define WE_SWITCH(WE, ...)                                         \
    switch (WE) {                                                  \
        case 1: { constexpr int EE = 1; __VA_ARGS__; } break;      \
        case 2: { constexpr int EE = 2; __VA_ARGS__; } break;      \
        case 3: { constexpr int EE = 3; __VA_ARGS__; } break;      \
        case 4: { constexpr int EE = 4; __VA_ARGS__; } break;      \
        case 5: { constexpr int EE = 5; __VA_ARGS__; } break;      \
        case 6: { constexpr int EE = 6; __VA_ARGS__; } break;      \
        case 7: { constexpr int EE = 7; __VA_ARGS__; } break;      \
        case 8: { constexpr int EE = 8; __VA_ARGS__; } break;      \
        default: break;                                            \
    }

template <int PW, int TP>
int launch_edge_part_t(int WE, const EdgePartParams &P, unsigned grid, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    WE_SWITCH(WE, {
        const size_t smem = bin_smem_bytes(EE + PW, TP);
        e = cudaFuncSetAttribute(k_edge_part<EE, PW, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_edge_part<EE, PW, TP><<<grid, PART_THREADS, smem, st>>>(P);
    });
    return e == cudaSuccess ? 0 : -1;
}

// positions per k_edge_part CTA: as many as shared memory allows with >= 2 CTAs per SM (1024 level-1 bins per CTA: a small
// tile leaves runs of 1-2 items per bin, i.e. one cursor atomic and one partial sector per item)
int edge_tile_positions(int IW) {
    int tp = IW <= 4 ? 4096 : (IW <= 9 ? 2048 : 1024);
    if (const char *e = getenv("MGTA_EDGE_TP")) { const int v = atoi(e); if (v == 1024 || v == 2048 || v == 4096) tp = v; }   // A/B switch
    return tp;
}

int launch_edge_part(int WE, int PW, const EdgePartParams &P, uint64_t total_bases, cudaStream_t st) {
    const int TP = edge_tile_positions(WE + PW);
    const unsigned grid = (unsigned)((total_bases + TP - 1) / TP);
#define EP_CASE(PWV)                                                        \
    if (PW == PWV) {                                                        \
        if (TP == 4096) return launch_edge_part_t<PWV, 4096>(WE, P, grid, st); \
        if (TP == 2048) return launch_edge_part_t<PWV, 2048>(WE, P, grid, st); \
        return launch_edge_part_t<PWV, 1024>(WE, P, grid, st);               \
    }
    EP_CASE(0) EP_CASE(1) EP_CASE(2)
#undef EP_CASE
    return -1;
}

int launch_count(int WE, bool plus, const CountParams &P, unsigned grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    WE_SWITCH(WE, {
        if (plus) {
            e = cudaFuncSetAttribute(k_count<EE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_count<EE, true><<<grid, COUNT_THREADS, smem, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(k_count<EE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_count<EE, false><<<grid, COUNT_THREADS, smem, st>>>(P);
        }
    });
    return e == cudaSuccess ? 0 : -1;
}

template <int PER>
int launch_item_part_t(int WE, bool plus, const ItemPartParams &P, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    WE_SWITCH(WE, {
        const int iw = EE + (plus ? 1 : 0) + 1;
        const unsigned per_cta = item_slots(iw, PER) / PER;
        const unsigned grid = (unsigned)((P.n_edges + per_cta - 1) / per_cta);
        const size_t smem = bin_smem_bytes(iw, item_slots(iw, PER));
        if (plus) {
            e = cudaFuncSetAttribute(k_item_part<EE, true, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_item_part<EE, true, PER><<<grid, PART_THREADS, smem, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(k_item_part<EE, false, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_item_part<EE, false, PER><<<grid, PART_THREADS, smem, st>>>(P);
        }
    });
    return e == cudaSuccess ? 0 : -1;
}

// real_only: the two real items per edge (the node pass supplies the $-items); else all six
int launch_item_part(int WE, bool plus, bool real_only, const ItemPartParams &P, cudaStream_t st) {
    return real_only ? launch_item_part_t<2>(WE, plus, P, st) : launch_item_part_t<6>(WE, plus, P, st);
}

#define KW_SWITCH(KW, ...)                                         \
    switch (KW) {                                                  \
        case 1: { constexpr int KK = 1; __VA_ARGS__; } break;      \
        case 2: { constexpr int KK = 2; __VA_ARGS__; } break;      \
        case 3: { constexpr int KK = 3; __VA_ARGS__; } break;      \
        case 4: { constexpr int KK = 4; __VA_ARGS__; } break;      \
        case 5: { constexpr int KK = 5; __VA_ARGS__; } break;      \
        case 6: { constexpr int KK = 6; __VA_ARGS__; } break;      \
        case 7: { constexpr int KK = 7; __VA_ARGS__; } break;      \
        case 8: { constexpr int KK = 8; __VA_ARGS__; } break;      \
        default: break;                                            \
    }

int launch_node_part(int KW, bool eplus, const NodePartParams &P, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    KW_SWITCH(KW, {
        const unsigned grid = (unsigned)((P.n_edges + node_edges(KK) - 1) / node_edges(KK));
        const size_t smem = bin_smem_bytes(KK + 1, node_edges(KK) * 2);
        if (eplus) {
            e = cudaFuncSetAttribute(k_node_part<KK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_node_part<KK, true><<<grid, PART_THREADS, smem, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(k_node_part<KK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_node_part<KK, false><<<grid, PART_THREADS, smem, st>>>(P);
        }
    });
    return e == cudaSuccess ? 0 : -1;
}

int launch_node_count(int KW, bool plus, const NodeCountParams &P, unsigned grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    KW_SWITCH(KW, {
        if (plus) {
            e = cudaFuncSetAttribute(k_node_count<KK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_node_count<KK, true><<<grid, COUNT_THREADS, smem, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(k_node_count<KK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) k_node_count<KK, false><<<grid, COUNT_THREADS, smem, st>>>(P);
        }
    });
    return e == cudaSuccess ? 0 : -1;
}

unsigned split_chunk_items(int IW) { return IW <= 4 ? 4096u : (IW <= 9 ? 2048u : 1024u); }

int launch_scans(mgta_ctx *ctx, const ScanParams &SP) {
    const unsigned B1 = SP.NT >> SP.lb2;
    k_scan_local<<<B1, 256, 0, ctx->stream>>>(SP);
    k_scan_top<<<1, 1024, 0, ctx->stream>>>(SP);
    k_scan_apply<<<(SP.NT + 255) / 256, 256, 0, ctx->stream>>>(SP);
    CK(cudaGetLastError());
    return MGTA_OK;
}

int launch_split(mgta_ctx *ctx, const SplitParams &SP) {
    const size_t smem = bin_smem_bytes(SP.IW, (int)SP.T + 4);      // + 4: the bulk copy of a chunk starts on a 16-byte boundary
    CK(cudaFuncSetAttribute(k_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_split, PART_THREADS, smem));
    k_split<<<(unsigned)(ctx->sm_count * std::max(1, occ)), PART_THREADS, smem, ctx->stream>>>(SP);
    CK(cudaGetLastError());
    return MGTA_OK;
}

// ---- the counting pipeline ------------------------------------------------------------------------
// reads -> canonical (k+1)-mer items -> two hash partition levels -> per-tile hash tables.
//   stage-1 mode : marks is_solid, fills edge_counting, lists edges with count >= m (or assist occurrences)
//   general mode : counts the occurrences the is_solid vector calls solid (no marking), lists every edge
// Both leave {(canonical edge, multiplicity)} in ctx->d_edges and the stage-2 key-prefix histogram in d_hist_s2.
//
// The level-1 hash bins of a batch are filled either by k_edge_part straight from the reads (one GPU, or every shard
// scanning all reads for its own hash range), or -- scan-sharded exchange, world > 1 -- by k_split (mode 2) from the items
// the shards sent each other: shard r extracts the items of ITS reads only, binned by owner shard, one all-to-all moves
// them (mgta_stage1_scan / _exchange_buffers / _count), and everything after the level-1 bins is the same code.
enum CountMode { CM_STAGE1 = 0, CM_GENERAL = 1, CM_MARK = 2 };

struct CountLay {
    size_t R, A, B, hist2, loc, off2, cur2, tot, base, in_start, chunk_pref, cur1, ovf, total;
    uint64_t slab_cap, capA, capB;
    unsigned bins, NT;
};

int make_count_plan(mgta_ctx *ctx, CountMode mode, CountPlan &cp) {
    cp.stage1_mode = mode != CM_GENERAL;                           // counts every occurrence of this shard's hash range
    cp.mark_mode = mode == CM_MARK;                                // only (re)derives the is_solid vector
    cp.k = ctx->opt.kmer_k;
    int rc = count_positions(ctx);
    if (rc) return rc;
    cp.n_pos = ctx->n_positions;
    cp.WE = edge_words(cp.k);
    cp.plus = key_words_s2(cp.k) > cp.WE;
    cp.has_assist = cp.stage1_mode && ctx->n_short < ctx->n_reads;
    // the base position rides along only when it is needed: to mark is_solid, or to tell assist occurrences apart
    cp.PW = (cp.mark_mode || cp.has_assist) ? (ctx->total_bases >= 0xFFFFFFFEull ? 2 : 1) : 0;
    cp.IW = cp.WE + cp.PW;
    // per-tile table: sized so that two CTAs share an SM; overflow tiles get the largest table that fits one SM
    const size_t slot_bytes = 4 * (size_t)(2 + (cp.has_assist ? 1 : 0) + cp.WE) + 2;
    cp.tab_cap = 4096; cp.big_cap = 16384;
    while (cp.tab_cap > 1024 && cp.tab_cap * slot_bytes > 106 * 1024) cp.tab_cap >>= 1;       // 2 x (106 KB + static) fit one SM
    while (cp.big_cap > cp.tab_cap && cp.big_cap * slot_bytes > 200 * 1024) cp.big_cap >>= 1;
    cp.tab_limit = cp.tab_cap - 640;                               // COUNT_THREADS inserts may be in flight past the check
    if (ctx->opt.sort_items_cap > 0) cp.tab_limit = std::min<unsigned>(cp.tab_limit, std::max(8, ctx->opt.sort_items_cap));   // test hook: force the overflow pass
    // tiles: level-2 fan-out <= 1024, level-1 bins as many as it takes (a batch handles <= 1024).
    // mean tile = up to 3/4 of the table (a Poisson tail of 6 sigma still fits below tab_limit; denser tiles only pay the
    // overflow pass).  Keeping the mean high matters: one more bit doubles the level-1 bins, and past MAX_BINS the reads
    // are scanned once per batch of bins.
    int bits = 2;
    unsigned fill_pct = 75;
    if (const char *e = getenv("MGTA_TILE_FILL_PCT")) fill_pct = (unsigned)std::max(10, std::min(400, atoi(e)));   // A/B switch
    while (bits < 28 && (cp.n_pos >> bits) > (uint64_t)cp.tab_cap * fill_pct / 100) ++bits;
    cp.bits = bits;
    cp.lb2 = (unsigned)std::min(10, bits / 2);
    cp.lb1 = (unsigned)bits - cp.lb2;
    cp.B1 = 1u << cp.lb1;
    cp.T = split_chunk_items(cp.IW);
    // this shard's level-1 hash bins (hash ranges balance the shards without a histogram pass)
    // (the general mode has no exchange step after it, so there every shard counts the whole hash space)
    cp.r_lo = cp.stage1_mode ? (unsigned)((uint64_t)cp.B1 * ctx->opt.rank / ctx->opt.world) : 0u;
    cp.r_hi = cp.stage1_mode ? (unsigned)((uint64_t)cp.B1 * (ctx->opt.rank + 1) / ctx->opt.world) : cp.B1;
    return MGTA_OK;
}

// arena carve for batches of `nb` level-1 bins.  recv_bytes != 0: room for the items received from the other shards
// (first, at a fixed offset: it must survive a slab-overflow restart); with a single batch it doubles as buffer B.
void count_layout(const CountPlan &cp, CountLay &L, unsigned nb, double slack, size_t send_bytes, size_t recv_bytes, uint64_t n_recv_cap) {
    L.bins = (cp.r_hi - cp.r_lo + nb - 1) / nb;
    L.NT = L.bins << cp.lb2;
    L.slab_cap = ((uint64_t)((double)cp.n_pos / cp.B1 * slack) + 1024 + 31) & ~(uint64_t)31;
    L.capA = L.slab_cap * L.bins;
    uint64_t most = (cp.n_pos + 31) & ~(uint64_t)31;
    if (recv_bytes) most = std::min<uint64_t>(most, (n_recv_cap + 31) & ~(uint64_t)31);
    L.capB = std::min<uint64_t>(L.capA, most);
    Carver c;
    const size_t bytesA = std::max((size_t)cp.IW * L.capA * 4, send_bytes), bytesB = (size_t)cp.IW * L.capB * 4;
    if (recv_bytes && nb == 1) { L.B = c.take(std::max(bytesB, recv_bytes)); L.R = L.B; }
    else if (recv_bytes) { L.R = c.take(recv_bytes); L.B = c.take(bytesB); }
    else { L.R = 0; L.B = c.take(bytesB); }
    L.A = c.take(bytesA);
    L.hist2 = c.take((size_t)L.NT * 4); L.loc = c.take((size_t)L.NT * 4);
    L.off2 = c.take(((size_t)L.NT + 1) * 8); L.cur2 = c.take((size_t)L.NT * 8);
    const size_t regions = std::max<size_t>(cp.B1, MAX_OWNERS) + 1;
    L.tot = c.take(regions * 8); L.base = c.take(regions * 8); L.in_start = c.take(regions * 8);
    L.chunk_pref = c.take(regions * 4); L.cur1 = c.take(regions * 8);
    L.ovf = c.take((size_t)L.NT * 4);
    L.total = c.o;
}

// grows the arena keeping bytes [0, keep) (the received items of a scan-sharded exchange)
int ensure_arena_keep(mgta_ctx *ctx, size_t bytes, size_t keep) {
    if (bytes <= ctx->arena_bytes) return MGTA_OK;
    if (!keep || !ctx->arena) return ensure_arena(ctx, bytes);
    CK(mgta_stream_wait(ctx->stream));
    unsigned char *na = nullptr;
    CK(cudaMalloc(&na, bytes));
    CK(cudaMemcpyAsync(na, ctx->arena, std::min(keep, ctx->arena_bytes), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    cudaFree(ctx->arena);
    ctx->arena = na; ctx->arena_bytes = bytes;
    return MGTA_OK;
}

// One batch of level-1 bins [b_lo, b_hi) whose slabs (buffer A, cursors cur1) and tile histogram hist2 are filled:
// exact tile offsets, level-2 split, per-tile hash counting, epilogue.  overflow = a level-1 slab was too small (nothing
// was counted); need_slack then holds the slack that fits what was seen.
int count_batch_tail(mgta_ctx *ctx, const CountPlan &cp, const CountLay &L, unsigned b_lo, unsigned b_hi, unsigned n_batches,
                     unsigned batch, double slack, mgta_stage_stats *st, bool &overflow, double &need_slack) {
    int rc;
    overflow = false;
    const int WE = cp.WE, PW = cp.PW, IW = cp.IW, k = cp.k;
    uint32_t *bufA = reinterpret_cast<uint32_t *>(ctx->arena + L.A), *bufB = reinterpret_cast<uint32_t *>(ctx->arena + L.B);
    uint32_t *hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
    unsigned long long *cur1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
    unsigned long long *off2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.off2);
    const size_t smem_count = count_smem_bytes(WE, cp.tab_cap, cp.has_assist), smem_big = count_smem_bytes(WE, cp.big_cap, cp.has_assist);
    int occ = 1;
    // ---- exact tile offsets, K3a: level-2 split
    ScanParams SP;
    memset(&SP, 0, sizeof(SP));
    const unsigned NTb = (b_hi - b_lo) << cp.lb2;
    SP.hist = hist2; SP.NT = NTb; SP.lb2 = cp.lb2; SP.t_lo = 0; SP.t_hi = NTb;
    SP.loc = reinterpret_cast<uint32_t *>(ctx->arena + L.loc);
    SP.tot = reinterpret_cast<unsigned long long *>(ctx->arena + L.tot);
    SP.base = reinterpret_cast<unsigned long long *>(ctx->arena + L.base);
    SP.off2 = off2; SP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur2);
    SP.cursor1 = nullptr; SP.chunk_pref = reinterpret_cast<unsigned *>(ctx->arena + L.chunk_pref);
    SP.T = cp.T; SP.slab_cap = L.slab_cap; SP.b1_lo = 0;
    SP.in_start = reinterpret_cast<unsigned long long *>(ctx->arena + L.in_start);
    if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
    if ((rc = launch_scans(ctx, SP))) return rc;
    SplitParams XP;
    memset(&XP, 0, sizeof(XP));
    XP.src = bufA; XP.dst = bufB; XP.cap_src = L.capA; XP.cap_dst = L.capB; XP.IW = IW; XP.WE = WE; XP.mode = 0;
    XP.sh2 = 32 - cp.bits; XP.lb2 = cp.lb2; XP.in_start = SP.in_start; XP.in_count = SP.tot; XP.chunk_pref = SP.chunk_pref;
    XP.B1 = b_hi - b_lo; XP.cursor2 = SP.cursor2; XP.ticket = ctx->d_ctr + CTR_TICKET; XP.T = cp.T; XP.err = ctx->d_ctr + CTR_ERR;
    if ((rc = launch_split(ctx, XP))) return rc;
    if ((rc = end_timed(ctx))) return rc;
    st->n_launches += 4;
    // ---- K4: per-tile hash counting (edge rows are staged in buffer A, dead after the split)
    CountParams CP;
    memset(&CP, 0, sizeof(CP));
    CP.src = bufB; CP.cap = L.capB; CP.PW = PW; CP.k = k; CP.off2 = off2; CP.t_lo = SP.t_lo; CP.t_hi = SP.t_hi;
    CP.ticket = ctx->d_ctr + CTR_TICKET2; CP.tab_cap = cp.tab_cap; CP.tab_limit = cp.tab_limit; CP.m = (unsigned)ctx->opt.min_count;
    CP.mark = cp.mark_mode ? 1 : 0; CP.threshold = cp.stage1_mode ? 1 : 0; CP.has_assist = cp.has_assist ? 1 : 0;
    CP.emit = cp.mark_mode ? 0 : 1;
    CP.solid = ctx->d_solid; CP.edge_counting = (cp.stage1_mode && !cp.mark_mode) ? ctx->d_ec : nullptr;
    CP.edges_out = bufA; CP.n_edges = ctx->d_totals + 14; CP.edges_cap = (uint64_t)IW * L.capA / (WE + 1);
    CP.hist_s2 = ctx->d_hist_s2; CP.s2_shift = 32 - ctx->PB; CP.real_only = ctx->node_pass ? 1 : 0;
    CP.ovf_list = reinterpret_cast<unsigned *>(ctx->arena + L.ovf); CP.n_ovf = ctx->d_ctr + CTR_NOVF; CP.ovf_cap = L.NT;
    CP.err = ctx->d_ctr + CTR_ERR;
    if ((rc = begin_timed(ctx, PH_SORT))) return rc;
    {
        cudaError_t e = cudaSuccess;
        WE_SWITCH(WE, {
            if (cp.plus) {
                e = cudaFuncSetAttribute(k_count<EE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_count<EE, true>, COUNT_THREADS, smem_count);
            } else {
                e = cudaFuncSetAttribute(k_count<EE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_count<EE, false>, COUNT_THREADS, smem_count);
            }
        });
        if (e != cudaSuccess) occ = 1;
        cudaGetLastError();
    }
    if (launch_count(WE, cp.plus, CP, (unsigned)(ctx->sm_count * std::max(1, std::min(occ, 4))), smem_count, ctx->stream))
        FAIL(MGTA_ERR_CUDA, "k_count launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CountParams CB = CP;                                                             // overflow tiles: one CTA per SM, large table
    CB.tile_list = CP.ovf_list; CB.n_tile_list = CP.n_ovf; CB.ticket = ctx->d_ctr + CTR_TICKET3;
    CB.tab_cap = cp.big_cap; CB.tab_limit = cp.big_cap - 1024; CB.ovf_cap = 0; CB.n_ovf = ctx->d_ctr + CTR_NOVF2;
    if (launch_count(WE, cp.plus, CB, (unsigned)ctx->sm_count, smem_big, ctx->stream))
        FAIL(MGTA_ERR_CUDA, "k_count (overflow pass) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaGetLastError());
    if ((rc = end_timed(ctx))) return rc;
    st->n_launches += 2;
    // ---- batch epilogue
    unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
    unsigned long long *h_ne = ctx->h_pin + 2 * NUM_BUCKETS + 8;
    CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_ne, ctx->d_totals + 14, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_ne + 1, off2 + NTb, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    const unsigned dev_err = h_ctr[CTR_ERR];
    if (dev_err & ERR_SLAB_OVERFLOW) {                    // this batch was not counted (k_split / k_count bail out)
        std::vector<unsigned long long> hc(b_hi - b_lo);   // the cursors kept counting past the slab ends: exact bin sizes
        CK(cudaMemcpy(hc.data(), cur1, hc.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (size_t i = 0; i < hc.size(); ++i) mx = std::max(mx, hc[i] - (unsigned long long)i * L.slab_cap);
        need_slack = std::max(slack * 1.5, (double)mx / ((double)cp.n_pos / cp.B1) * 1.05);
        overflow = true;
        return MGTA_OK;
    }
    if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (count pipeline, bins [%u,%u))", dev_err, b_lo, b_hi);
    st->n_items += h_ne[1];
    st->n_batches++;
    st->msd_levels = std::max<int>(st->msd_levels, (int)h_ctr[CTR_NOVF]);          // overflow tiles (informational)
    const uint64_t ne = h_ne[0];
    if (ne) {
        const size_t row = (size_t)(WE + 1) * 4;
        if (ctx->n_edges + ne > ctx->edges_cap) {
            // room for the batches still to come, extrapolated from the share counted so far (hash batches are balanced);
            // if that much is not available, exactly what is needed now
            const uint64_t need = ctx->n_edges + ne;
            uint64_t ncap = n_batches > 1 ? std::max<uint64_t>(need, (uint64_t)((double)need * n_batches / (batch + 1) * 1.08)) : need;
            uint32_t *nbuf = nullptr;
            if (ncap > need && cudaMalloc(&nbuf, ncap * row) != cudaSuccess) {
                cudaGetLastError();
                nbuf = nullptr; ncap = need;
            }
            if (!nbuf) CK(cudaMalloc(&nbuf, ncap * row));
            if (ctx->n_edges) CK(cudaMemcpyAsync(nbuf, ctx->d_edges, ctx->n_edges * row, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(mgta_stream_wait(ctx->stream));
            cudaFree(ctx->d_edges);
            ctx->d_edges = nbuf; ctx->edges_cap = ncap;
        }
        CK(cudaMemcpyAsync(reinterpret_cast<unsigned char *>(ctx->d_edges) + ctx->n_edges * row, bufA, ne * row, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->n_edges += ne;
    }
    return MGTA_OK;
}

int count_reset_outputs(mgta_ctx *ctx, const CountPlan &cp) {
    if (cp.mark_mode) {
        CK(cudaMemsetAsync(ctx->d_solid, 0, ctx->solid_words * 4, ctx->stream));
    } else {
        ctx->n_edges = 0;
        ctx->s2x.valid = false;
        ctx->edges_complete = false;
        ctx->tips_valid = false;
        ctx->n_tips = 0;
        CK(cudaMemsetAsync(ctx->d_hist_s2, 0, ((size_t)1 << ctx->PB) * 4, ctx->stream));
        if (cp.stage1_mode) CK(cudaMemsetAsync(ctx->d_ec, 0, NUM_BUCKETS * 8, ctx->stream));
    }
    return MGTA_OK;
}

int count_batch_begin(mgta_ctx *ctx, const CountLay &L, unsigned n_bins) {
    CK(cudaMemsetAsync(ctx->arena + L.hist2, 0, (size_t)L.NT * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals + 14, 0, 8, ctx->stream));                      // edge rows written by this batch
    k_init_slab_cursors<<<(n_bins + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1), n_bins, L.slab_cap);
    CK(cudaGetLastError());
    return MGTA_OK;
}

int run_count(mgta_ctx *ctx, CountMode mode, mgta_stage_stats *st) {
    CountPlan cp;
    int rc = make_count_plan(ctx, mode, cp);
    if (rc) return rc;
    const int k = cp.k, WE = cp.WE, PW = cp.PW, IW = cp.IW;
    st->key_words = WE; st->item_words = IW;
    if (!cp.mark_mode) {
        ctx->edges_valid = false;
        ctx->edge_row_words = WE + 1;
    }
    if (cp.n_pos == 0 || cp.r_lo >= cp.r_hi) {
        if (!cp.mark_mode) {
            if ((rc = count_reset_outputs(ctx, cp))) return rc;
            ctx->edges_valid = true;
            ctx->edges_complete = mode == CM_GENERAL || ctx->opt.world == 1;
        }
        return MGTA_OK;
    }
    st->sort_cap = (int)cp.tab_cap;
    const size_t budget = hbm_budget(ctx);
    double slack = 1.15;
    for (int attempt = 0;; ++attempt) {          // a level-1 slab overflow restarts the pipeline with slabs sized from what was seen
        bool retry = false;
        memset(st, 0, sizeof(*st));
        st->key_words = WE; st->item_words = IW; st->sort_cap = (int)cp.tab_cap; st->n_giants = (uint64_t)attempt;
        if ((rc = count_reset_outputs(ctx, cp))) return rc;
        CountLay L;
        unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
        count_layout(cp, L, n_batches, slack, 0, 0, 0);
        while (L.total > budget && L.bins > 1) { n_batches *= 2; count_layout(cp, L, n_batches, slack, 0, 0, 0); }
        if (L.total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold one level-1 hash bin (%zu B)", budget, L.total);
        if ((rc = ensure_arena(ctx, L.total))) return rc;
        for (unsigned batch = 0; batch < n_batches; ++batch) {
            const unsigned b_lo = cp.r_lo + batch * L.bins, b_hi = std::min(cp.r_hi, b_lo + L.bins);
            if (b_lo >= b_hi) break;
            if ((rc = count_batch_begin(ctx, L, b_hi - b_lo))) return rc;
            // ---- K1+K2: extraction + level-1 hash partition
            EdgePartParams EP;
            memset(&EP, 0, sizeof(EP));
            EP.seq = ctx->d_seq; EP.start = ctx->d_start; EP.lut = ctx->d_lut; EP.n_lut = ctx->n_lut; EP.n_reads = ctx->n_reads; EP.n_short = ctx->n_short;
            EP.total_bases = ctx->total_bases; EP.k = k; EP.g_begin = 0; EP.g_end = ctx->total_bases;
            EP.filter = cp.stage1_mode ? 0 : 1; EP.all_solid = ctx->opt.min_count == 1; EP.solid = ctx->d_solid;
            EP.sh1 = 32 - (int)cp.lb1; EP.sh2 = 32 - cp.bits; EP.lb2 = cp.lb2; EP.b_lo = b_lo; EP.b_hi = b_hi;
            EP.cursor1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1); EP.slab_cap = L.slab_cap;
            EP.hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
            EP.dst = reinterpret_cast<uint32_t *>(ctx->arena + L.A); EP.cap = L.capA;
            EP.err = ctx->d_ctr + CTR_ERR;
            if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
            if (ctx->copy_pending && cp.stage1_mode && !cp.mark_mode) {
                // the upload is still in flight: extract chunk c once chunk c + 1 has landed (a tile stages a few words
                // past its end), so the copy hides behind the extraction
                const int nch = (int)ctx->ev_chunk.size();
                uint64_t g0 = 0;
                for (int c = 0; c < nch; ++c) {
                    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[std::min(c + 1, nch - 1)], 0));
                    EP.g_begin = g0; EP.g_end = ctx->chunk_end[c];
                    if (EP.g_end > EP.g_begin) {
                        if (launch_edge_part(WE, PW, EP, EP.g_end - EP.g_begin, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_edge_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                        st->n_launches++;
                    }
                    g0 = EP.g_end;
                }
                ctx->copy_pending = false;
            } else {
                if ((rc = wait_reads(ctx))) return rc;
                if (launch_edge_part(WE, PW, EP, ctx->total_bases, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_edge_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                st->n_launches++;
            }
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            bool overflow = false;
            double need = slack;
            if ((rc = count_batch_tail(ctx, cp, L, b_lo, b_hi, n_batches, batch, slack, st, overflow, need))) return rc;
            if (overflow) {
                if (attempt >= 6) FAIL(MGTA_ERR_MEM, "level-1 hash bins overflow their slabs even with %.1fx slack", slack);
                slack = need;
                retry = true;
                break;
            }
        }
        if (!retry) break;
    }
    if (cp.mark_mode) ctx->solid_valid = true;
    else { ctx->edges_valid = true; ctx->edges_complete = mode == CM_GENERAL || ctx->opt.world == 1; }
    return MGTA_OK;
}

// ---- scan-sharded stage 1 (world > 1): scan, [caller: all-to-all], count ---------------------------------------------
// round / n_rounds: the hash range of every shard is cut into n_rounds slices of level-1 bins and one exchange handles
// one slice (scan of the same reads again, 1 / n_rounds of the items): for inputs whose items do not fit the HBM at once.
// dry_run: only compute the arena bytes this exchange would need (*needed), nothing is launched.
int exchange_scan(mgta_ctx *ctx, uint64_t r_begin, uint64_t r_end, uint64_t slab_in, uint64_t *needed, mgta_stage_stats *st,
                  unsigned round = 0, unsigned n_rounds = 1, bool dry_run = false) {
    ExchangeState &X = ctx->xch;
    X.valid = false;
    if (r_begin > r_end || r_end > ctx->n_reads) FAIL(MGTA_ERR_ARG, "stage1_scan: bad read range");
    { int rcw = wait_reads(ctx); if (rcw) return rcw; }
    const int world = ctx->opt.world;
    if (world > MAX_OWNERS) FAIL(MGTA_ERR_ARG, "stage1_scan: at most %d shards", (int)MAX_OWNERS);
    CountPlan cp;
    int rc = make_count_plan(ctx, CM_STAGE1, cp);
    if (rc) return rc;
    st->key_words = cp.WE; st->item_words = cp.IW; st->sort_cap = (int)cp.tab_cap;
    {                                                              // this round's slice of my level-1 bins
        const unsigned lo = cp.r_lo, cnt = cp.r_hi - cp.r_lo;
        cp.r_lo = lo + (unsigned)((uint64_t)cnt * round / n_rounds);
        cp.r_hi = lo + (unsigned)((uint64_t)cnt * (round + 1) / n_rounds);
    }
    if (dry_run) {
        const uint64_t slab_items = (slab_in + 31) & ~(uint64_t)31;
        const size_t xbytes = (size_t)world * cp.IW * slab_items * 4;
        unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
        if (cp.r_hi <= cp.r_lo) n_batches = 1;
        CountLay L;
        count_layout(cp, L, n_batches, 1.15, xbytes, xbytes, (uint64_t)world * slab_items);
        *needed = L.total;
        return MGTA_OK;
    }
    ctx->edges_valid = false;
    ctx->edge_row_words = cp.WE + 1;
    // edge offsets of the local reads, base range of the scan
    uint64_t n_local = 0, g_begin = 0, g_end = 0;
    if (r_begin < r_end) {
        CK(cudaMemsetAsync(ctx->d_totals + 13, 0, 8, ctx->stream));
        k_count_positions<<<(unsigned)((r_end - r_begin + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_start, r_begin, r_end, cp.k, ctx->d_totals + 13);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals + 13, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pin + 1, ctx->d_start + r_begin, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pin + 2, ctx->d_start + r_end, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
        n_local = ctx->h_pin[0]; g_begin = ctx->h_pin[1]; g_end = ctx->h_pin[2];
    }
    std::vector<unsigned> owner_lo(world + 1);
    for (int d = 0; d <= world; ++d) owner_lo[d] = (unsigned)((uint64_t)cp.B1 * d / world);
    const size_t budget = hbm_budget(ctx);
    // send slabs: hash ranges are balanced, so a slab is the mean share plus a small margin.  All shards must use the
    // same slab size (equal-split all-to-all), so the caller agrees on it: slab_in == 0 only reports the size this shard
    // would like; the cursors keep counting past a full slab, so a skewed input (one k-mer dominating) reports the exact
    // size that fits and costs one rescan.
    if (slab_in == 0) {
        *needed = ((uint64_t)((double)n_local / world * 1.02) + 4096 + 31) & ~(uint64_t)31;
        return MGTA_OK;
    }
    const uint64_t slab_items = (slab_in + 31) & ~(uint64_t)31;
    {
        const size_t xbytes = (size_t)world * cp.IW * slab_items * 4;
        // the count phase's arena layout, sized for the worst case this shard can receive
        unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
        if (cp.r_hi <= cp.r_lo) n_batches = 1;
        CountLay L;
        count_layout(cp, L, n_batches, 1.15, xbytes, xbytes, (uint64_t)world * slab_items);
        if (L.total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the scan-sharded exchange (%zu B)", budget, L.total);
        if ((rc = ensure_arena(ctx, L.total))) return rc;
        uint32_t *send = reinterpret_cast<uint32_t *>(ctx->arena + L.A);
        unsigned long long *cur = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
        const unsigned long long stride = (unsigned long long)cp.IW * slab_items;
        CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
        k_init_slab_cursors<<<1, 256, 0, ctx->stream>>>(cur, (unsigned)world, stride);
        CK(cudaGetLastError());
        if (g_begin < g_end) {
            EdgePartParams EP;
            memset(&EP, 0, sizeof(EP));
            EP.seq = ctx->d_seq; EP.start = ctx->d_start; EP.lut = ctx->d_lut; EP.n_lut = ctx->n_lut; EP.n_reads = ctx->n_reads; EP.n_short = ctx->n_short;
            EP.total_bases = ctx->total_bases; EP.k = cp.k; EP.filter = 0; EP.all_solid = 0; EP.solid = ctx->d_solid;
            EP.sh1 = 32 - (int)cp.lb1; EP.sh2 = 32 - cp.bits; EP.lb2 = cp.lb2; EP.b_lo = 0; EP.b_hi = cp.B1;
            EP.cursor1 = cur; EP.slab_cap = slab_items; EP.slab_stride = stride; EP.hist2 = nullptr;
            EP.dst = send; EP.cap = slab_items; EP.err = ctx->d_ctr + CTR_ERR;
            EP.n_owner = world;
            for (int d = 0; d <= world; ++d) EP.owner_lo[d] = owner_lo[d];
            EP.n_rounds = n_rounds; EP.round = round;
            EP.g_begin = g_begin & ~(uint64_t)1023; EP.g_end = g_end; EP.r_begin = r_begin;
            if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
            if (launch_edge_part(cp.WE, cp.PW, EP, g_end - EP.g_begin, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_edge_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            st->n_launches++;
        }
        unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
        CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pin, cur, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
        X.send_counts.assign(world, 0);
        uint64_t mx = 0;
        for (int d = 0; d < world; ++d) { X.send_counts[d] = ctx->h_pin[d] - (unsigned long long)d * stride; mx = std::max(mx, X.send_counts[d]); }
        const unsigned dev_err = h_ctr[CTR_ERR];
        *needed = std::max<uint64_t>(slab_items, (mx + 31) & ~(uint64_t)31);
        if (dev_err & ERR_SLAB_OVERFLOW) { st->n_giants++; return MGTA_OK; }         // *needed > slab: the caller rescans
        if (dev_err & ~(unsigned)ERR_SLAB_OVERFLOW) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (stage1_scan)", dev_err);
        X.cp = cp; X.slab_items = slab_items; X.send_off = L.A; X.recv_off = L.R; X.xbytes = xbytes;
        X.round = round; X.n_rounds = n_rounds;
        X.valid = true;
        return MGTA_OK;
    }
}

int exchange_count(mgta_ctx *ctx, const uint64_t *recv_counts, mgta_stage_stats *st) {
    ExchangeState &X = ctx->xch;
    if (!X.valid) FAIL(MGTA_ERR_STATE, "stage1_count: call mgta_stage1_scan (and exchange the items) first");
    X.valid = false;
    const CountPlan &cp = X.cp;
    const int world = ctx->opt.world;
    int rc;
    uint64_t n_recv = 0;
    for (int s = 0; s < world; ++s) {
        if (recv_counts[s] > X.slab_items) FAIL(MGTA_ERR_ARG, "stage1_count: shard %d sent %llu items, a slab holds %llu", s,
                                                (unsigned long long)recv_counts[s], (unsigned long long)X.slab_items);
        n_recv += recv_counts[s];
    }
    if (X.round == 0 && (rc = count_reset_outputs(ctx, cp))) return rc;         // later rounds add to what the earlier ones left
    if (n_recv == 0 || cp.r_lo >= cp.r_hi) { ctx->edges_valid = true; return MGTA_OK; }
    // input regions of the level-1 split = the slabs received from the shards
    std::vector<unsigned long long> in_start(world + 1, 0), in_count(world + 1, 0);
    std::vector<unsigned> chunk_pref(world + 1, 0);
    for (int s = 0; s < world; ++s) {
        in_start[s] = (unsigned long long)s * cp.IW * X.slab_items;
        in_count[s] = recv_counts[s];
        chunk_pref[s + 1] = chunk_pref[s] + (unsigned)((recv_counts[s] + cp.T - 1) / cp.T);
    }
    unsigned long long *d_xs = ctx->d_xs;                          // [in_start | in_count | chunk_pref]: 3 small arrays
    CK(cudaMemcpyAsync(d_xs, in_start.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_xs + (world + 1), in_count.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_xs + 2 * (world + 1), chunk_pref.data(), (size_t)(world + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    const size_t budget = hbm_budget(ctx);
    // state a restart of this round must go back to (a level-1 slab overflow is detected after earlier batches counted)
    const uint64_t edges0 = ctx->n_edges;
    CK(cudaMemcpyAsync(ctx->d_hist_bak, ctx->d_hist_s2, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ec_bak, ctx->d_ec, NUM_BUCKETS * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    double slack = 1.15;
    for (int attempt = 0;; ++attempt) {
        bool retry = false;
        const uint64_t giants0 = st->n_giants;
        if (X.round == 0) { st->n_items = 0; st->n_batches = 0; }
        if (attempt) {
            ctx->n_edges = edges0;
            CK(cudaMemcpyAsync(ctx->d_hist_s2, ctx->d_hist_bak, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_ec, ctx->d_ec_bak, NUM_BUCKETS * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        CountLay L;
        unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
        count_layout(cp, L, n_batches, slack, X.xbytes, X.xbytes, (uint64_t)world * X.slab_items);
        if (L.R != X.recv_off) { FAIL(MGTA_ERR_INTERNAL, "stage1_count: receive buffer moved"); }
        if (L.total > budget) { FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the level-1 slabs (%zu B)", budget, L.total); }
        if ((rc = ensure_arena_keep(ctx, L.total, X.recv_off + X.xbytes))) return rc;
        for (unsigned batch = 0; batch < n_batches; ++batch) {
            const unsigned b_lo = cp.r_lo + batch * L.bins, b_hi = std::min(cp.r_hi, b_lo + L.bins);
            if (b_lo >= b_hi) break;
            if ((rc = count_batch_begin(ctx, L, b_hi - b_lo))) return rc;
            SplitParams XP;
            memset(&XP, 0, sizeof(XP));
            XP.src = reinterpret_cast<uint32_t *>(ctx->arena + L.R); XP.dst = reinterpret_cast<uint32_t *>(ctx->arena + L.A);
            XP.cap_src = X.slab_items; XP.cap_dst = L.capA; XP.IW = cp.IW; XP.WE = cp.WE; XP.mode = 2;
            XP.sh1 = 32 - (int)cp.lb1; XP.sh2 = 32 - cp.bits; XP.lb2 = cp.lb2;
            XP.in_start = d_xs; XP.in_count = d_xs + (world + 1); XP.chunk_pref = reinterpret_cast<unsigned *>(d_xs + 2 * (world + 1));
            XP.B1 = (unsigned)world; XP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
            XP.ticket = ctx->d_ctr + CTR_NLIST0; XP.T = cp.T; XP.err = ctx->d_ctr + CTR_ERR;
            XP.b_lo = b_lo; XP.b_hi = b_hi; XP.slab_cap = L.slab_cap; XP.hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
            if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
            if ((rc = launch_split(ctx, XP))) return rc;
            if ((rc = end_timed(ctx))) return rc;
            st->n_launches++;
            bool overflow = false;
            double need = slack;
            if ((rc = count_batch_tail(ctx, cp, L, b_lo, b_hi, n_batches, batch, slack, st, overflow, need))) return rc;
            if (overflow) {
                if (attempt >= 6) { FAIL(MGTA_ERR_MEM, "level-1 hash bins overflow their slabs even with %.1fx slack", slack); }
                slack = need;
                st->n_giants = giants0 + 1;
                retry = true;
                break;
            }
        }
        if (!retry) break;
    }
    ctx->edges_valid = true;
    return MGTA_OK;
}

// ---- mercy edges (need_mercy) ------------------------------------------------------------------------
// stage-1 items of the reference (s1_position) -> hash partition by (k-1)-mer -> per-tile tables -> candidates
// (s1.cpp:671-830); then candidates -> position bit vectors -> per-read scan adding is_solid bits (s2.cpp:106-250).
template <int TP>
int launch_ctx_part_t(int W, const CtxPartParams &P, unsigned grid, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    W_SWITCH(W, {
        const size_t smem = bin_smem_bytes(WW + 2, 2 * TP);
        e = cudaFuncSetAttribute(k_ctx_part<WW, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_ctx_part<WW, TP><<<grid, PART_THREADS, smem, st>>>(P);
    });
    return e == cudaSuccess ? 0 : -1;
}

int mercy_apply(mgta_ctx *ctx, mgta_stage_stats *st);

// sharded: only the (k-1)-mers whose hash lies in this shard's 1 / world of the 32-bit hash space are grouped (every shard
// scans all reads; the ranges are cut in hash space, not in bins, so a shard that refines its partition still owns the
// same keys); the candidates stay in ctx->d_cand and the caller gathers them before mercy_apply().
int run_mercy(mgta_ctx *ctx, mgta_stage_stats *st, bool sharded = false) {
    const int k = ctx->opt.kmer_k;
    const int W = key_words_s1(k), IW = W + 2;
    int rc = count_positions(ctx);
    if (rc) return rc;
    if ((rc = wait_reads(ctx))) return rc;
    ctx->mercy_valid = false;
    ctx->n_cand = 0; ctx->num_mercy = 0;
    const uint64_t n_items_max = ctx->n_positions + 4 * ctx->n_reads;      // L - k + 4 per read with L >= k + 1
    if (ctx->n_positions == 0) { ctx->mercy_valid = true; return MGTA_OK; }
    // table: as many slots as one CTA can hold; tiles of mean <= cap / 2 ITEMS can never hold more than tab_limit distinct keys
    unsigned tab_cap = 4096;
    // the table capacity (and with it the tile plan) is that of the byte tables; with min_count <= 3 the counts fit 2-bit
    // fields, the same table takes a third of the shared memory and three CTAs share an SM
    const bool compact = ctx->opt.min_count <= 3 && !getenv("MGTA_MERCY_BYTE_TABLES");
    while (tab_cap > 256 && mercy_smem_bytes(W, tab_cap) > 200 * 1024) tab_cap >>= 1;
    const unsigned tab_limit = tab_cap - 600;
    // tiles of mean <= 0.6 * cap ITEMS: even if every item were a distinct S a 6-sigma tile stays below tab_limit; a tile
    // that does not (ERR_TABLE_FULL) restarts the pass with one more partition bit
    int bits = 2, lb2i = 0, lb1i = 0;
    unsigned lb2 = 0, lb1 = 0, B1 = 0;
    auto set_bits = [&](int extra) {
        bits = 2;
        while (bits < 30 && (n_items_max >> bits) > (uint64_t)tab_cap * 6 / 10) ++bits;
        bits = std::min(30, bits + extra);
        lb2i = std::min(10, bits / 2); lb1i = bits - lb2i;
        lb2 = (unsigned)lb2i; lb1 = (unsigned)lb1i; B1 = 1u << lb1;
    };
    int bits_extra = 0;
    set_bits(0);
    const unsigned T = split_chunk_items(IW);
    const int TP = IW <= 5 ? 2048 : 1024;
    const size_t budget = hbm_budget(ctx);
    const size_t vec_bytes = (((ctx->total_bases + 31) / 32 + 4) * 4 + 255) & ~(size_t)255;     // one position bit vector
    const unsigned long long ha_lo = sharded ? (1ull << 32) * (unsigned)ctx->opt.rank / (unsigned)ctx->opt.world : 0ull;
    const unsigned long long ha_hi = sharded ? (1ull << 32) * ((unsigned)ctx->opt.rank + 1) / (unsigned)ctx->opt.world : (1ull << 32);
    double slack = 1.25;
    uint64_t cand_cap = std::max<uint64_t>(ctx->cand_cap, std::max<uint64_t>(1 << 20, n_items_max / 8));
    for (int attempt = 0;; ++attempt) {
        if (attempt >= 8) FAIL(MGTA_ERR_MEM, "mercy: the partition does not settle");
        if (cand_cap > ctx->cand_cap) {
            CK(mgta_stream_wait(ctx->stream));
            cudaFree(ctx->d_cand);
            ctx->d_cand = nullptr; ctx->cand_cap = 0;
            CK(cudaMalloc(&ctx->d_cand, cand_cap * 8));
            ctx->cand_cap = cand_cap;
        }
        CK(cudaMemsetAsync(ctx->d_totals + 11, 0, 8, ctx->stream));                            // candidate counter
        bool retry = false;
        // level-1 bins that hold hashes of [ha_lo, ha_hi) (the first and the last may be shared with a neighbour shard)
        const unsigned R_lo = (unsigned)(ha_lo >> (32 - lb1)), R_hi = ha_hi > ha_lo ? (unsigned)((ha_hi - 1) >> (32 - lb1)) + 1 : R_lo;
        unsigned n_batches = std::max(1u, (R_hi - R_lo + MAX_BINS - 1) / MAX_BINS);
        struct { size_t A, B, hist2, loc, off2, cur2, tot, base, in_start, chunk_pref, cur1, total; uint64_t slab_cap, capA, capB; unsigned bins, NT; } L;
        auto layout = [&](unsigned nb) {
            L.bins = std::max(1u, (R_hi - R_lo + nb - 1) / nb);
            L.NT = L.bins << lb2;
            L.slab_cap = ((uint64_t)((double)n_items_max / B1 * slack) + 2048 + 31) & ~(uint64_t)31;
            L.capA = L.slab_cap * L.bins;
            L.capB = std::min<uint64_t>(L.capA, (n_items_max + 31) & ~(uint64_t)31);
            Carver c;
            L.A = c.take((size_t)IW * L.capA * 4); L.B = c.take((size_t)IW * L.capB * 4);
            L.hist2 = c.take((size_t)L.NT * 4); L.loc = c.take((size_t)L.NT * 4);
            L.off2 = c.take(((size_t)L.NT + 1) * 8); L.cur2 = c.take((size_t)L.NT * 8);
            L.tot = c.take((MAX_BINS + 1) * 8); L.base = c.take((MAX_BINS + 1) * 8); L.in_start = c.take((MAX_BINS + 1) * 8);
            L.chunk_pref = c.take((MAX_BINS + 1) * 4); L.cur1 = c.take((MAX_BINS + 1) * 8);
            L.total = c.o;
        };
        layout(n_batches);
        while (L.total + 3 * vec_bytes > budget && L.bins > 1) { n_batches *= 2; layout(n_batches); }
        if (L.total + 3 * vec_bytes > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold one level-1 bin of mercy items (%zu B)", budget, L.total);
        if ((rc = ensure_arena(ctx, L.total + 3 * vec_bytes))) return rc;
        const size_t smem_m = mercy_smem_bytes(W, tab_cap, compact);
        for (unsigned batch = 0; batch < n_batches && !retry; ++batch) {
            const unsigned b_lo = R_lo + batch * L.bins, b_hi = std::min(R_hi, b_lo + L.bins);
            if (b_lo >= b_hi) break;
            uint32_t *bufA = reinterpret_cast<uint32_t *>(ctx->arena + L.A), *bufB = reinterpret_cast<uint32_t *>(ctx->arena + L.B);
            uint32_t *hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
            unsigned long long *cur1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
            unsigned long long *off2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.off2);
            CK(cudaMemsetAsync(hist2, 0, (size_t)L.NT * 4, ctx->stream));
            CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
            k_init_slab_cursors<<<(b_hi - b_lo + 255) / 256, 256, 0, ctx->stream>>>(cur1, b_hi - b_lo, L.slab_cap);
            CK(cudaGetLastError());
            CtxPartParams XP0;
            memset(&XP0, 0, sizeof(XP0));
            XP0.seq = ctx->d_seq; XP0.start = ctx->d_start; XP0.lut = ctx->d_lut; XP0.n_lut = ctx->n_lut; XP0.n_reads = ctx->n_reads; XP0.n_short = ctx->n_short;
            XP0.total_bases = ctx->total_bases; XP0.k = k; XP0.sh1 = 32 - (int)lb1; XP0.sh2 = 32 - bits; XP0.lb2 = lb2;
            XP0.b_lo = b_lo; XP0.b_hi = b_hi; XP0.cursor1 = cur1; XP0.slab_cap = L.slab_cap; XP0.hist2 = hist2; XP0.dst = bufA;
            XP0.cap = L.capA; XP0.err = ctx->d_ctr + CTR_ERR;
            XP0.ha_lo = (uint32_t)ha_lo; XP0.ha_last = (uint32_t)(ha_hi - 1);
            if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
            const unsigned grid = (unsigned)((ctx->total_bases + TP - 1) / TP);
            if (TP == 2048 ? launch_ctx_part_t<2048>(W, XP0, grid, ctx->stream) : launch_ctx_part_t<1024>(W, XP0, grid, ctx->stream))
                FAIL(MGTA_ERR_CUDA, "k_ctx_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            ScanParams SP;
            memset(&SP, 0, sizeof(SP));
            const unsigned NTb = (b_hi - b_lo) << lb2;
            SP.hist = hist2; SP.NT = NTb; SP.lb2 = lb2; SP.t_lo = 0; SP.t_hi = NTb;
            SP.loc = reinterpret_cast<uint32_t *>(ctx->arena + L.loc);
            SP.tot = reinterpret_cast<unsigned long long *>(ctx->arena + L.tot);
            SP.base = reinterpret_cast<unsigned long long *>(ctx->arena + L.base);
            SP.off2 = off2; SP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur2);
            SP.cursor1 = nullptr; SP.chunk_pref = reinterpret_cast<unsigned *>(ctx->arena + L.chunk_pref);
            SP.T = T; SP.slab_cap = L.slab_cap; SP.b1_lo = 0;
            SP.in_start = reinterpret_cast<unsigned long long *>(ctx->arena + L.in_start);
            if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
            if ((rc = launch_scans(ctx, SP))) return rc;
            SplitParams XP;
            memset(&XP, 0, sizeof(XP));
            XP.src = bufA; XP.dst = bufB; XP.cap_src = L.capA; XP.cap_dst = L.capB; XP.IW = IW; XP.WE = W; XP.mode = 0;
            XP.drop_last = ~S1_FLAG_MASK;
            XP.sh2 = 32 - bits; XP.lb2 = lb2; XP.in_start = SP.in_start; XP.in_count = SP.tot; XP.chunk_pref = SP.chunk_pref;
            XP.B1 = b_hi - b_lo; XP.cursor2 = SP.cursor2; XP.ticket = ctx->d_ctr + CTR_TICKET; XP.T = T; XP.err = ctx->d_ctr + CTR_ERR;
            if ((rc = launch_split(ctx, XP))) return rc;
            if ((rc = end_timed(ctx))) return rc;
            MercyParams MP;
            memset(&MP, 0, sizeof(MP));
            MP.src = bufB; MP.cap = L.capB; MP.off2 = off2; MP.t_lo = 0; MP.t_hi = NTb; MP.ticket = ctx->d_ctr + CTR_TICKET2;
            MP.tab_cap = tab_cap; MP.tab_limit = tab_limit; MP.m = (unsigned)std::min(ctx->opt.min_count, 255);
            MP.cand_out = ctx->d_cand; MP.n_cand = ctx->d_totals + 11; MP.cand_cap = ctx->cand_cap; MP.err = ctx->d_ctr + CTR_ERR;
            if ((rc = begin_timed(ctx, PH_SORT))) return rc;
            {
                cudaError_t e = cudaSuccess;
                W_SWITCH(W, {
                    int occ_m = 1;
                    if (compact) {
                        e = cudaFuncSetAttribute(k_mercy<WW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
                        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, k_mercy<WW, true>, COUNT_THREADS, smem_m);
                        if (e == cudaSuccess) k_mercy<WW, true><<<(unsigned)(ctx->sm_count * std::max(1, occ_m)), COUNT_THREADS, smem_m, ctx->stream>>>(MP);
                    } else {
                        e = cudaFuncSetAttribute(k_mercy<WW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
                        if (e == cudaSuccess) k_mercy<WW, false><<<(unsigned)ctx->sm_count, COUNT_THREADS, smem_m, ctx->stream>>>(MP);
                    }
                });
                if (e != cudaSuccess) FAIL(MGTA_ERR_CUDA, "k_mercy launch failed: %s", cudaGetErrorString(e));
            }
            CK(cudaGetLastError());
            if ((rc = end_timed(ctx))) return rc;
            st->n_launches += 6;
            unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
            CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(mgta_stream_wait(ctx->stream));
            const unsigned dev_err = h_ctr[CTR_ERR];
            if (dev_err & ERR_SLAB_OVERFLOW) {
                std::vector<unsigned long long> hc(b_hi - b_lo);
                CK(cudaMemcpy(hc.data(), cur1, hc.size() * 8, cudaMemcpyDeviceToHost));
                unsigned long long mx = 0;
                for (size_t i = 0; i < hc.size(); ++i) mx = std::max(mx, hc[i] - (unsigned long long)i * L.slab_cap);
                slack = std::max(slack * 1.5, (double)mx / ((double)n_items_max / B1) * 1.05);
                retry = true;
                break;
            }
            if (dev_err == ERR_TABLE_FULL && bits_extra < 4) {    // a tile with more distinct (k-1)-mers than the table takes: finer tiles
                set_bits(++bits_extra);
                retry = true;
                break;
            }
            if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (mercy pipeline, bins [%u,%u))", dev_err, b_lo, b_hi);
        }
        if (retry) continue;
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals + 11, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
        const uint64_t n_cand = ctx->h_pin[0];
        if (n_cand > ctx->cand_cap) { cand_cap = n_cand + n_cand / 16 + 1024; continue; }     // counted past the end: rerun with room
        ctx->n_cand = n_cand;
        break;
    }
    if (sharded) return MGTA_OK;                                   // the caller gathers the candidates of all shards first
    return mercy_apply(ctx, st);
}

// ---- candidates (ctx->d_cand, ctx->n_cand) -> position bit vectors -> per-read scan extending is_solid (s2.cpp:106-250).
// The candidate buffers of the partition are dead by now: the three vectors take the start of the arena.
int mercy_apply(mgta_ctx *ctx, mgta_stage_stats *st) {
    int rc;
    const size_t vec_bytes = (((ctx->total_bases + 31) / 32 + 4) * 4 + 255) & ~(size_t)255;     // one position bit vector
    if ((rc = ensure_arena(ctx, 3 * vec_bytes))) return rc;
    uint32_t *v_in = reinterpret_cast<uint32_t *>(ctx->arena), *v_out = reinterpret_cast<uint32_t *>(ctx->arena + vec_bytes),
             *v_any = reinterpret_cast<uint32_t *>(ctx->arena + 2 * vec_bytes);
    CK(cudaMemsetAsync(v_in, 0, 3 * vec_bytes, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals + 11, 0, 8, ctx->stream));
    if (ctx->n_cand) {
        k_mercy_bits<<<(unsigned)((ctx->n_cand + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_cand, ctx->n_cand, v_in, v_out, v_any);
        k_mercy_reads<<<(unsigned)((ctx->n_short + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_start, ctx->n_short, ctx->opt.kmer_k, v_in, v_out, v_any,
                                                                                    ctx->d_solid, ctx->d_totals + 11);
        CK(cudaGetLastError());
        st->n_launches += 2;
    }
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals + 11, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    ctx->num_mercy = ctx->h_pin[0];
    ctx->mercy_valid = true;
    return MGTA_OK;
}

// is_solid is derived lazily: the hot path (stage 1 -> edge list -> stage 2) never reads it
int ensure_solid(mgta_ctx *ctx) {
    if (ctx->solid_valid || !ctx->stage1_done || ctx->opt.min_count == 1) return MGTA_OK;
    mgta_stage_stats tmp;
    int rc = run_count(ctx, CM_MARK, &tmp);
    if (rc) return rc;
    mgta_stage_stats dummy;
    memset(&dummy, 0, sizeof(dummy));
    return finish_timing(ctx, &dummy);
}

// ---- the node pass of stage 2 (node_kernels.cuh) ----------------------------------------------------------
// {(canonical edge, multiplicity)} -> 2 ops per edge keyed by canonical k-mer -> two hash partition levels -> per-tile
// tables of (out, in) weights -> the $-items of the tip k-mers (ctx->d_tips) + their share of the stage-2 histogram.
// One shard: run_nodes().  Several: node_exchange_scan() bins the ops of this shard's edges by the shard that owns the
// k-mer's hash range, one all-to-all moves them, node_exchange_count() finishes like the one-shard pass.
struct NodeShape {
    int k, KW, WE, W2, IW;
    bool eplus, plus2;
    size_t smem_count, smem_big, row;
};

NodeShape node_shape(int k) {
    NodeShape n;
    n.k = k; n.KW = kmer_words(k); n.WE = edge_words(k); n.W2 = key_words_s2(k); n.IW = n.KW + 1;
    n.eplus = n.WE > n.KW; n.plus2 = n.W2 > n.KW;
    n.smem_count = n.smem_big = 0;
    n.row = (size_t)(n.W2 + 1) * 4;
    return n;
}

// n_ops: ops of ALL shards (2 per distinct solid edge); every shard derives the same plan from it
void make_node_plan(mgta_ctx *ctx, uint64_t n_ops, CountPlan &cp, NodeShape &ns) {
    ns = node_shape(ctx->opt.kmer_k);
    memset(&cp, 0, sizeof(cp));
    cp.k = ns.k; cp.WE = ns.KW; cp.PW = 1; cp.IW = ns.IW; cp.has_assist = true; cp.stage1_mode = true; cp.n_pos = n_ops;
    const size_t slot_bytes = 4 * (size_t)(3 + ns.KW) + 2;
    cp.tab_cap = 4096; cp.big_cap = 16384;
    while (cp.tab_cap > 1024 && cp.tab_cap * slot_bytes > 106 * 1024) cp.tab_cap >>= 1;
    while (cp.big_cap > cp.tab_cap && cp.big_cap * slot_bytes > 200 * 1024) cp.big_cap >>= 1;
    cp.tab_limit = cp.tab_cap - 640;
    if (ctx->opt.sort_items_cap > 0) cp.tab_limit = std::min<unsigned>(cp.tab_limit, std::max(8, ctx->opt.sort_items_cap));   // test hook
    int bits = 2;
    while (bits < 28 && (n_ops >> bits) > (uint64_t)cp.tab_cap * 3 / 4) ++bits;
    cp.bits = bits;
    cp.lb2 = (unsigned)std::min(10, bits / 2);
    cp.lb1 = (unsigned)bits - cp.lb2;
    cp.B1 = 1u << cp.lb1;
    cp.T = split_chunk_items(ns.IW);
    cp.r_lo = (unsigned)((uint64_t)cp.B1 * ctx->opt.rank / ctx->opt.world);
    cp.r_hi = (unsigned)((uint64_t)cp.B1 * (ctx->opt.rank + 1) / ctx->opt.world);
    ns.smem_count = count_smem_bytes(ns.KW, cp.tab_cap, 1);
    ns.smem_big = count_smem_bytes(ns.KW, cp.big_cap, 1);
}

enum { NODE_OK = 0, NODE_RETRY_SLACK = 1, NODE_RETRY_TIPS = 2 };

// One batch of level-1 bins [b_lo, b_hi) whose slabs (buffer A, cursors cur1) and tile histogram hist2 are filled: exact
// tile offsets, level-2 split, per-tile (out, in) tables, tips.  verdict: NODE_OK, or what the caller must enlarge.
int node_batch_tail(mgta_ctx *ctx, const CountPlan &cp, const NodeShape &ns, const CountLay &L, unsigned b_lo, unsigned b_hi,
                    unsigned n_batches, unsigned batch, mgta_stage_stats *st, int &verdict, double &slack, uint64_t &tips_cap) {
    int rc;
    verdict = NODE_OK;
    uint32_t *bufA = reinterpret_cast<uint32_t *>(ctx->arena + L.A), *bufB = reinterpret_cast<uint32_t *>(ctx->arena + L.B);
    uint32_t *hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
    unsigned long long *cur1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
    unsigned long long *off2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.off2);
    ScanParams SP;
    memset(&SP, 0, sizeof(SP));
    const unsigned NTb = (b_hi - b_lo) << cp.lb2;
    SP.hist = hist2; SP.NT = NTb; SP.lb2 = cp.lb2; SP.t_lo = 0; SP.t_hi = NTb;
    SP.loc = reinterpret_cast<uint32_t *>(ctx->arena + L.loc);
    SP.tot = reinterpret_cast<unsigned long long *>(ctx->arena + L.tot);
    SP.base = reinterpret_cast<unsigned long long *>(ctx->arena + L.base);
    SP.off2 = off2; SP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur2);
    SP.cursor1 = nullptr; SP.chunk_pref = reinterpret_cast<unsigned *>(ctx->arena + L.chunk_pref);
    SP.T = cp.T; SP.slab_cap = L.slab_cap; SP.b1_lo = 0;
    SP.in_start = reinterpret_cast<unsigned long long *>(ctx->arena + L.in_start);
    if ((rc = launch_scans(ctx, SP))) return rc;
    SplitParams XP;
    memset(&XP, 0, sizeof(XP));
    XP.src = bufA; XP.dst = bufB; XP.cap_src = L.capA; XP.cap_dst = L.capB; XP.IW = ns.IW; XP.WE = ns.KW; XP.mode = 0;
    XP.sh2 = 32 - cp.bits; XP.lb2 = cp.lb2; XP.in_start = SP.in_start; XP.in_count = SP.tot; XP.chunk_pref = SP.chunk_pref;
    XP.B1 = b_hi - b_lo; XP.cursor2 = SP.cursor2; XP.ticket = ctx->d_ctr + CTR_TICKET; XP.T = cp.T; XP.err = ctx->d_ctr + CTR_ERR;
    if ((rc = launch_split(ctx, XP))) return rc;
    NodeCountParams CP;
    memset(&CP, 0, sizeof(CP));
    CP.src = bufB; CP.cap = L.capB; CP.k = ns.k; CP.off2 = off2; CP.t_lo = 0; CP.t_hi = NTb; CP.ticket = ctx->d_ctr + CTR_TICKET2;
    CP.tab_cap = cp.tab_cap; CP.tab_limit = cp.tab_limit; CP.tips_out = ctx->d_tips; CP.n_tips = ctx->d_totals + 11;
    CP.tips_cap = ctx->tips_cap; CP.hist_s2 = ctx->d_hist_s2; CP.s2_shift = 32 - ctx->PB;
    CP.ovf_list = reinterpret_cast<unsigned *>(ctx->arena + L.ovf); CP.n_ovf = ctx->d_ctr + CTR_NOVF; CP.ovf_cap = L.NT;
    CP.err = ctx->d_ctr + CTR_ERR;
    if (launch_node_count(ns.KW, ns.plus2, CP, (unsigned)(ctx->sm_count * 2), ns.smem_count, ctx->stream))
        FAIL(MGTA_ERR_CUDA, "k_node_count launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    NodeCountParams CB = CP;                                                     // overflow tiles: one CTA per SM, large table
    CB.tile_list = CP.ovf_list; CB.n_tile_list = CP.n_ovf; CB.ticket = ctx->d_ctr + CTR_TICKET3;
    CB.tab_cap = cp.big_cap; CB.tab_limit = cp.big_cap - 1024; CB.ovf_cap = 0; CB.n_ovf = ctx->d_ctr + CTR_NOVF2;
    if (launch_node_count(ns.KW, ns.plus2, CB, (unsigned)ctx->sm_count, ns.smem_big, ctx->stream))
        FAIL(MGTA_ERR_CUDA, "k_node_count (overflow pass) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaGetLastError());
    st->n_launches += 6;
    unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
    unsigned long long *h_nt = ctx->h_pin + 2 * NUM_BUCKETS + 8;
    CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_nt, ctx->d_totals + 11, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    const unsigned dev_err = h_ctr[CTR_ERR];
    if (dev_err & ERR_SLAB_OVERFLOW) {
        std::vector<unsigned long long> hc(b_hi - b_lo);
        CK(cudaMemcpy(hc.data(), cur1, hc.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (size_t i = 0; i < hc.size(); ++i) mx = std::max(mx, hc[i] - (unsigned long long)i * L.slab_cap);
        slack = std::max(slack * 1.5, (double)mx / ((double)cp.n_pos / cp.B1) * 1.05);
        verdict = NODE_RETRY_SLACK;
        return MGTA_OK;
    }
    if (dev_err & ERR_EDGE_LIST_FULL) {                             // the tip list was too small: the counter kept counting
        tips_cap = std::max<uint64_t>(2 * ctx->tips_cap, (uint64_t)((double)h_nt[0] * n_batches / (batch + 1) * 1.1) + 1024);
        verdict = NODE_RETRY_TIPS;
        return MGTA_OK;
    }
    if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (node pass, bins [%u,%u))", dev_err, b_lo, b_hi);
    ctx->n_tips = h_nt[0];
    return MGTA_OK;
}

int node_ensure_tips(mgta_ctx *ctx, uint64_t tips_cap, size_t row) {
    if (tips_cap > ctx->tips_cap || !ctx->d_tips) {
        CK(mgta_stream_wait(ctx->stream));
        cudaFree(ctx->d_tips);
        ctx->d_tips = nullptr; ctx->tips_cap = 0;
        CK(cudaMalloc(&ctx->d_tips, tips_cap * row));
        ctx->tips_cap = tips_cap;
    }
    return MGTA_OK;
}

int run_nodes(mgta_ctx *ctx, mgta_stage_stats *st) {
    int rc;
    ctx->tips_valid = false;
    ctx->n_tips = 0;
    const uint64_t n_edges = ctx->n_edges, n_ops = 2 * n_edges;
    st->n_node_ops = n_ops;
    st->n_tip_items = 0;
    if (n_edges == 0) { ctx->tips_valid = true; return MGTA_OK; }
    CountPlan cp;
    NodeShape ns;
    make_node_plan(ctx, n_ops, cp, ns);
    cp.r_lo = 0; cp.r_hi = cp.B1;                                   // the edge list is complete: this shard sees every k-mer
    const size_t budget = hbm_budget(ctx);
    CK(cudaMemcpyAsync(ctx->d_hist_bak, ctx->d_hist_s2, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    uint64_t tips_cap = std::max<uint64_t>(ctx->tips_cap, std::max<uint64_t>(1u << 20, n_edges / 4));
    double slack = 1.15;
    for (int attempt = 0;; ++attempt) {
        if (attempt >= 10) FAIL(MGTA_ERR_MEM, "node pass: the partition does not settle");
        if (attempt) CK(cudaMemcpyAsync(ctx->d_hist_s2, ctx->d_hist_bak, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = node_ensure_tips(ctx, tips_cap, ns.row))) return rc;
        CK(cudaMemsetAsync(ctx->d_totals + 11, 0, 8, ctx->stream));                  // tip rows (free after the mercy pass of stage 1)
        int verdict = NODE_OK;
        CountLay L;
        unsigned n_batches = (cp.B1 + MAX_BINS - 1) / MAX_BINS;
        count_layout(cp, L, n_batches, slack, 0, 0, 0);
        while (L.total > budget && L.bins > 1) { n_batches *= 2; count_layout(cp, L, n_batches, slack, 0, 0, 0); }
        if (L.total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold one level-1 bin of the node pass (%zu B)", budget, L.total);
        if ((rc = ensure_arena(ctx, L.total))) return rc;
        for (unsigned batch = 0; batch < n_batches && verdict == NODE_OK; ++batch) {
            const unsigned b_lo = batch * L.bins, b_hi = std::min(cp.B1, b_lo + L.bins);
            if (b_lo >= b_hi) break;
            if ((rc = count_batch_begin(ctx, L, b_hi - b_lo))) return rc;
            if ((rc = begin_timed(ctx, PH_NODES))) return rc;
            NodePartParams NP;
            memset(&NP, 0, sizeof(NP));
            NP.edges = ctx->d_edges; NP.n_edges = n_edges; NP.k = ns.k; NP.sh1 = 32 - (int)cp.lb1; NP.sh2 = 32 - cp.bits; NP.lb2 = cp.lb2;
            NP.b_lo = b_lo; NP.b_hi = b_hi; NP.cursor1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1); NP.slab_cap = L.slab_cap;
            NP.hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2); NP.dst = reinterpret_cast<uint32_t *>(ctx->arena + L.A); NP.cap = L.capA;
            NP.err = ctx->d_ctr + CTR_ERR;
            if (launch_node_part(ns.KW, ns.eplus, NP, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_node_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            CK(cudaGetLastError());
            st->n_launches++;
            if ((rc = node_batch_tail(ctx, cp, ns, L, b_lo, b_hi, n_batches, batch, st, verdict, slack, tips_cap))) return rc;
            if ((rc = end_timed(ctx))) return rc;
        }
        if (verdict == NODE_OK) break;
    }
    st->n_tip_items = ctx->n_tips;
    ctx->tips_valid = true;
    return MGTA_OK;
}

// ---- sharded node pass: scan (ops of my edges binned by owner), [caller: all-to-all], count ---------------------------
int node_exchange_scan(mgta_ctx *ctx, uint64_t n_ops_all, uint64_t slab_in, uint64_t *needed, mgta_stage_stats *st) {
    ExchangeState &X = ctx->xch;
    X.valid = false;
    const int world = ctx->opt.world;
    CountPlan cp;
    NodeShape ns;
    make_node_plan(ctx, n_ops_all, cp, ns);
    int rc;
    const size_t budget = hbm_budget(ctx);
    const uint64_t slab_items = (slab_in + 31) & ~(uint64_t)31;
    const size_t xbytes = (size_t)world * ns.IW * slab_items * 4;
    unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
    if (cp.r_hi <= cp.r_lo) n_batches = 1;
    CountLay L;
    count_layout(cp, L, n_batches, 1.15, xbytes, xbytes, (uint64_t)world * slab_items);
    if (L.total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the node-pass exchange (%zu B)", budget, L.total);
    if ((rc = ensure_arena(ctx, L.total))) return rc;
    uint32_t *send = reinterpret_cast<uint32_t *>(ctx->arena + L.A);
    unsigned long long *cur = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
    const unsigned long long stride = (unsigned long long)ns.IW * slab_items;
    CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
    k_init_slab_cursors<<<1, 256, 0, ctx->stream>>>(cur, (unsigned)world, stride);
    CK(cudaGetLastError());
    if (ctx->n_edges) {
        NodePartParams NP;
        memset(&NP, 0, sizeof(NP));
        NP.edges = ctx->d_edges; NP.n_edges = ctx->n_edges; NP.k = ns.k; NP.sh1 = 32 - (int)cp.lb1; NP.sh2 = 32 - cp.bits; NP.lb2 = cp.lb2;
        NP.b_lo = 0; NP.b_hi = cp.B1; NP.cursor1 = cur; NP.slab_cap = slab_items; NP.slab_stride = stride; NP.hist2 = nullptr;
        NP.dst = send; NP.cap = slab_items; NP.err = ctx->d_ctr + CTR_ERR; NP.n_owner = world;
        for (int d = 0; d <= world; ++d) NP.owner_lo[d] = (unsigned)((uint64_t)cp.B1 * d / world);
        if ((rc = begin_timed(ctx, PH_NODES))) return rc;
        if (launch_node_part(ns.KW, ns.eplus, NP, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_node_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        CK(cudaGetLastError());
        if ((rc = end_timed(ctx))) return rc;
        st->n_launches++;
    }
    unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
    CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_pin, cur, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    X.send_counts.assign(world, 0);
    uint64_t mx = 0;
    for (int d = 0; d < world; ++d) { X.send_counts[d] = ctx->h_pin[d] - (unsigned long long)d * stride; mx = std::max(mx, X.send_counts[d]); }
    const unsigned dev_err = h_ctr[CTR_ERR];
    *needed = std::max<uint64_t>(slab_items, (mx + 31) & ~(uint64_t)31);
    if (dev_err & ERR_SLAB_OVERFLOW) return MGTA_OK;                 // *needed > slab: the caller rescans
    if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (node scan)", dev_err);
    X.cp = cp; X.slab_items = slab_items; X.send_off = L.A; X.recv_off = L.R; X.xbytes = xbytes;
    X.valid = true;
    return MGTA_OK;
}

int node_exchange_count(mgta_ctx *ctx, const uint64_t *recv_counts, mgta_stage_stats *st) {
    ExchangeState &X = ctx->xch;
    if (!X.valid) FAIL(MGTA_ERR_STATE, "node count: no exchange pending");
    X.valid = false;
    const CountPlan &cp = X.cp;
    NodeShape ns = node_shape(ctx->opt.kmer_k);
    ns.smem_count = count_smem_bytes(ns.KW, cp.tab_cap, 1);
    ns.smem_big = count_smem_bytes(ns.KW, cp.big_cap, 1);
    const int world = ctx->opt.world;
    int rc;
    ctx->tips_valid = false;
    ctx->n_tips = 0;
    uint64_t n_recv = 0;
    for (int s = 0; s < world; ++s) {
        if (recv_counts[s] > X.slab_items) FAIL(MGTA_ERR_ARG, "node count: shard %d sent %llu ops, a slab holds %llu", s,
                                                (unsigned long long)recv_counts[s], (unsigned long long)X.slab_items);
        n_recv += recv_counts[s];
    }
    st->n_node_ops = n_recv;
    if (n_recv == 0 || cp.r_lo >= cp.r_hi) { ctx->tips_valid = true; return MGTA_OK; }
    std::vector<unsigned long long> in_start(world + 1, 0), in_count(world + 1, 0);
    std::vector<unsigned> chunk_pref(world + 1, 0);
    for (int s = 0; s < world; ++s) {
        in_start[s] = (unsigned long long)s * cp.IW * X.slab_items;
        in_count[s] = recv_counts[s];
        chunk_pref[s + 1] = chunk_pref[s] + (unsigned)((recv_counts[s] + cp.T - 1) / cp.T);
    }
    unsigned long long *d_xs = ctx->d_xs;
    CK(cudaMemcpyAsync(d_xs, in_start.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_xs + (world + 1), in_count.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_xs + 2 * (world + 1), chunk_pref.data(), (size_t)(world + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));                               // the host vectors die with this frame
    const size_t budget = hbm_budget(ctx);
    CK(cudaMemcpyAsync(ctx->d_hist_bak, ctx->d_hist_s2, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    uint64_t tips_cap = std::max<uint64_t>(ctx->tips_cap, std::max<uint64_t>(1u << 20, n_recv / 8));
    double slack = 1.15;
    for (int attempt = 0;; ++attempt) {
        if (attempt >= 10) FAIL(MGTA_ERR_MEM, "node pass: the partition does not settle");
        if (attempt) CK(cudaMemcpyAsync(ctx->d_hist_s2, ctx->d_hist_bak, ((size_t)1 << ctx->PB) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = node_ensure_tips(ctx, tips_cap, ns.row))) return rc;
        CK(cudaMemsetAsync(ctx->d_totals + 11, 0, 8, ctx->stream));
        int verdict = NODE_OK;
        CountLay L;
        unsigned n_batches = (cp.r_hi - cp.r_lo + MAX_BINS - 1) / MAX_BINS;
        count_layout(cp, L, n_batches, slack, X.xbytes, X.xbytes, (uint64_t)world * X.slab_items);
        if (L.R != X.recv_off) FAIL(MGTA_ERR_INTERNAL, "node count: receive buffer moved");
        if (L.total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the level-1 slabs of the node pass (%zu B)", budget, L.total);
        if ((rc = ensure_arena_keep(ctx, L.total, X.recv_off + X.xbytes))) return rc;
        for (unsigned batch = 0; batch < n_batches && verdict == NODE_OK; ++batch) {
            const unsigned b_lo = cp.r_lo + batch * L.bins, b_hi = std::min(cp.r_hi, b_lo + L.bins);
            if (b_lo >= b_hi) break;
            if ((rc = count_batch_begin(ctx, L, b_hi - b_lo))) return rc;
            SplitParams XP;
            memset(&XP, 0, sizeof(XP));
            XP.src = reinterpret_cast<uint32_t *>(ctx->arena + L.R); XP.dst = reinterpret_cast<uint32_t *>(ctx->arena + L.A);
            XP.cap_src = X.slab_items; XP.cap_dst = L.capA; XP.IW = cp.IW; XP.WE = cp.WE; XP.mode = 2;
            XP.sh1 = 32 - (int)cp.lb1; XP.sh2 = 32 - cp.bits; XP.lb2 = cp.lb2;
            XP.in_start = d_xs; XP.in_count = d_xs + (world + 1); XP.chunk_pref = reinterpret_cast<unsigned *>(d_xs + 2 * (world + 1));
            XP.B1 = (unsigned)world; XP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
            XP.ticket = ctx->d_ctr + CTR_NLIST0; XP.T = cp.T; XP.err = ctx->d_ctr + CTR_ERR;
            XP.b_lo = b_lo; XP.b_hi = b_hi; XP.slab_cap = L.slab_cap; XP.hist2 = reinterpret_cast<uint32_t *>(ctx->arena + L.hist2);
            if ((rc = begin_timed(ctx, PH_NODES))) return rc;
            if ((rc = launch_split(ctx, XP))) return rc;
            st->n_launches++;
            if ((rc = node_batch_tail(ctx, cp, ns, L, b_lo, b_hi, n_batches, batch, st, verdict, slack, tips_cap))) return rc;
            if ((rc = end_timed(ctx))) return rc;
        }
        if (verdict == NODE_OK) break;
    }
    st->n_tip_items = ctx->n_tips;
    ctx->tips_valid = true;
    return MGTA_OK;
}

// ---- the emission pipeline -------------------------------------------------------------------------
// {(canonical edge, multiplicity)} -> stage-2 items (S a | flags, multiplicity) -> two key-prefix partition levels
// (exact offsets) -> per-window on-chip sort + W/last/tip/multiplicity records (k_sort_emit) -> sink.
// from_slabs: the items come from the slabs the shards sent each other (ctx->s2x, at the start of the arena, which the
// buffers below then leave alone) instead of from this shard's own edge and tip lists.
int run_emit(mgta_ctx *ctx, mgta_bucket_sink sink, void *user, int64_t *totals, mgta_stage_stats *st, bool from_slabs = false) {
    const int k = ctx->opt.kmer_k;
    const int WE = edge_words(k);
    const int W = key_words_s2(k), IW = W + 1;
    const bool plus = W > WE;
    const int PB = ctx->PB;
    // two partition levels: level-2 fan-out <= 1024 tiles per level-1 bin; a batch handles <= MAX_BINS level-1 bins,
    // numbered relative to its first one (g1_lo), so any number of global level-1 bins works
    const unsigned lb1 = (unsigned)std::max(PB - 10, PB / 2), lb2 = (unsigned)PB - lb1, NT = 1u << PB;
    const unsigned B1 = std::min(1u << lb1, (unsigned)MAX_BINS), NTB = B1 << lb2;      // per-batch maxima
    const unsigned tiles_per_bucket = 1u << (PB - 16);
    int rc;
    st->key_words = W; st->item_words = IW;

    Plan pl;
    make_plan(pl, 2, k, ctx->opt.sort_items_cap);
    st->sort_cap = (int)pl.CAPI;

    // tile histogram -> host: lv1 bucket sizes, shard range, batches
    const uint32_t *h2 = ctx->h_hist2;
    CK(cudaMemcpyAsync(ctx->h_hist2, ctx->d_hist_s2, (size_t)NT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    ctx->hist.assign(NUM_BUCKETS, 0);
    uint64_t total = 0;
    for (unsigned t = 0; t < NT; ++t) { ctx->hist[t >> (PB - 16)] += h2[t]; total += h2[t]; }
    set_shard_range(ctx, total);
    uint64_t shard_items = 0, max_bucket = 0;
    for (int b = ctx->shard_lo; b < ctx->shard_hi; ++b) {
        shard_items += (uint64_t)ctx->hist[b];
        max_bucket = std::max<uint64_t>(max_bucket, (uint64_t)ctx->hist[b]);
    }
    st->n_items = shard_items;
    if (max_bucket >= 0xFFFFFFFFull) FAIL(MGTA_ERR_MEM, "a single lv1 bucket holds %llu items (limit 2^32-1)", (unsigned long long)max_bucket);

    CK(cudaMemsetAsync(ctx->d_meta, 0, NUM_BUCKETS * 3 * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals, 0, 10 * 8, ctx->stream));

    const size_t budget = hbm_budget(ctx);
    struct Lay { size_t A, B, flags, win, state, list0, list1, giants, loc, off2, cur2, tot, base, in_start, chunk_pref, cur1, out, tmp, total; } L;
    const bool tips_in = ctx->node_pass && ctx->tips_valid && !from_slabs;             // $-items come from the node pass
    const size_t arena_base = from_slabs ? ((ctx->s2x.bytes + 255) & ~(size_t)255) : 0;
    auto layout = [&](uint64_t cap) {
        carve(pl, cap, from_slabs ? ctx->s2x.n_dollar : (tips_in ? ctx->n_tips : cap / 3 + 1));
        Carver c;
        c.o = arena_base;
        L.A = c.take((size_t)IW * pl.cap * 4); L.B = c.take((size_t)IW * pl.cap * 4);
        L.flags = c.take((pl.cap / 32 + 64) * 4);
        const uint64_t n_win = pl.cap / pl.C + 2;
        L.win = c.take(n_win * 4); L.state = c.take((2 * n_win + 4) * 8);
        L.list0 = c.take((size_t)pl.list_cap * sizeof(Seg)); L.list1 = c.take((size_t)pl.list_cap * sizeof(Seg));
        L.giants = c.take((size_t)pl.giants_cap * sizeof(Giant));
        L.loc = c.take((size_t)NTB * 4); L.off2 = c.take(((size_t)NTB + 1) * 8); L.cur2 = c.take((size_t)NTB * 8);
        L.tot = c.take((B1 + 1) * 8); L.base = c.take((B1 + 1) * 8); L.in_start = c.take((B1 + 1) * 8);
        L.chunk_pref = c.take((B1 + 1) * 4); L.cur1 = c.take((B1 + 1) * 8);
        L.out = c.take(pl.out_cap);
        L.tmp = c.take(pl.out_cap);
        L.total = c.o;
    };
    uint64_t cap = std::max<uint64_t>(shard_items, 1024);
    layout(cap);
    while (L.total > budget && cap > std::max<uint64_t>(max_bucket, 1024)) {
        cap = std::max<uint64_t>(std::max<uint64_t>(max_bucket, 1024), (uint64_t)((double)cap * 0.9));
        layout(cap);
    }
    if (L.total > budget)
        FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the largest lv1 bucket (%llu items, need %zu B)", budget, (unsigned long long)max_bucket, L.total);
    if ((rc = ensure_arena_keep(ctx, L.total, arena_base))) return rc;
    uint32_t *bufA = reinterpret_cast<uint32_t *>(ctx->arena + L.A), *bufB = reinterpret_cast<uint32_t *>(ctx->arena + L.B);
    uint32_t *flags = reinterpret_cast<uint32_t *>(ctx->arena + L.flags), *win = reinterpret_cast<uint32_t *>(ctx->arena + L.win);
    unsigned long long *state = reinterpret_cast<unsigned long long *>(ctx->arena + L.state);
    Seg *lists[2] = {reinterpret_cast<Seg *>(ctx->arena + L.list0), reinterpret_cast<Seg *>(ctx->arena + L.list1)};
    Giant *giants = reinterpret_cast<Giant *>(ctx->arena + L.giants);
    unsigned long long *off2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.off2);
    unsigned char *outbuf = ctx->arena + L.out, *tmpbuf = ctx->arena + L.tmp;
    const unsigned T = split_chunk_items(IW);

    int occ = 1;
    const bool capi_const = pl.CAPI == 4096 && W <= 3;             // the usual shape (k <= 45) gets the constant-stride instantiation
    {
        cudaError_t e = cudaSuccess;
        if (capi_const) {
            W_SWITCH3(W, {
                e = cudaFuncSetAttribute(k_sort_emit<WW, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.chunk_smem);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sort_emit<WW, 4096>, CHUNK_THREADS, pl.chunk_smem);
            });
        } else {
            W_SWITCH(W, {
                e = cudaFuncSetAttribute(k_sort_emit<WW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.chunk_smem);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sort_emit<WW, 0>, CHUNK_THREADS, pl.chunk_smem);
            });
        }
        if (e != cudaSuccess) FAIL(MGTA_ERR_CUDA, "k_sort_emit setup failed: %s", cudaGetErrorString(e));
    }
    occ = std::max(1, occ);

    std::vector<int64_t> meta_host;
    std::vector<Seg> seg_host;
    int b0 = ctx->shard_lo;
    while (b0 < ctx->shard_hi) {
        int b1 = b0;
        uint64_t n_items = 0;
        const unsigned g1_lo = ((unsigned)b0 * tiles_per_bucket) >> lb2;                    // first (global) level-1 bin of the batch
        while (b1 < ctx->shard_hi && n_items + (uint64_t)ctx->hist[b1] <= pl.cap &&
               ((((unsigned)b1 + 1) * tiles_per_bucket - 1) >> lb2) - g1_lo < (unsigned)MAX_BINS) { n_items += (uint64_t)ctx->hist[b1]; ++b1; }
        if (b1 == b0) FAIL(MGTA_ERR_MEM, "bucket %d (%lld items) exceeds the batch capacity %llu", b0, (long long)ctx->hist[b0], (unsigned long long)pl.cap);
        st->n_batches++;
        if (n_items == 0) {
            if (sink) {
                meta_host.assign((size_t)(b1 - b0) * 3, 0);
                if (sink(user, b0, b1, nullptr, 0, meta_host.data()) != 0) FAIL(MGTA_ERR_ARG, "sink aborted");
            }
            b0 = b1;
            continue;
        }
        const unsigned t_lo = (unsigned)b0 * tiles_per_bucket, t_hi = (unsigned)b1 * tiles_per_bucket;
        const unsigned nb1 = ((t_hi - 1) >> lb2) - g1_lo + 1;                               // level-1 bins the batch touches (<= MAX_BINS)
        const unsigned rt_lo = t_lo - (g1_lo << lb2), rt_hi = t_hi - (g1_lo << lb2);          // its tiles, batch relative
        const unsigned n_windows = (unsigned)((n_items + pl.C - 1) / pl.C);
        CK(cudaMemsetAsync(flags, 0, (n_items / 32 + 64) * 4, ctx->stream));
        CK(cudaMemsetAsync(win, 0, ((size_t)n_windows + 2) * 4, ctx->stream));
        CK(cudaMemsetAsync(state, 0, (2 * (size_t)n_windows + 4) * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
        // ---- exact offsets of the batch's tiles
        ScanParams SP;
        memset(&SP, 0, sizeof(SP));
        SP.hist = ctx->d_hist_s2 + ((size_t)g1_lo << lb2); SP.NT = nb1 << lb2; SP.lb2 = lb2; SP.t_lo = rt_lo; SP.t_hi = rt_hi;
        SP.loc = reinterpret_cast<uint32_t *>(ctx->arena + L.loc);
        SP.tot = reinterpret_cast<unsigned long long *>(ctx->arena + L.tot);
        SP.base = reinterpret_cast<unsigned long long *>(ctx->arena + L.base);
        SP.off2 = off2; SP.cursor2 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur2);
        SP.cursor1 = reinterpret_cast<unsigned long long *>(ctx->arena + L.cur1);
        SP.chunk_pref = reinterpret_cast<unsigned *>(ctx->arena + L.chunk_pref);
        SP.T = T; SP.slab_cap = 0; SP.b1_lo = 0;
        SP.in_start = reinterpret_cast<unsigned long long *>(ctx->arena + L.in_start);
        if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
        if ((rc = launch_scans(ctx, SP))) return rc;
        if ((rc = end_timed(ctx))) return rc;
        // ---- K2': items into the batch's level-1 prefix bins at exact offsets: from the distinct edges (+ tip list), or
        //      from the slabs the shards sent each other
        if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
        if (from_slabs) {
            const int world = ctx->opt.world;
            const auto &X = ctx->s2x;
            std::vector<unsigned long long> in_start(world + 1, 0), in_count(world + 1, 0);
            std::vector<unsigned> chunk_pref(world + 1, 0);
            for (int sidx = 0; sidx < world; ++sidx) {
                in_start[sidx] = (unsigned long long)sidx * IW * X.slab_items;
                in_count[sidx] = X.counts[sidx];
                chunk_pref[sidx + 1] = chunk_pref[sidx] + (unsigned)((X.counts[sidx] + T - 1) / T);
            }
            unsigned long long *d_xs = ctx->d_xs;
            CK(cudaMemcpyAsync(d_xs, in_start.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(d_xs + (world + 1), in_count.data(), (size_t)(world + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(d_xs + 2 * (world + 1), chunk_pref.data(), (size_t)(world + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(mgta_stream_wait(ctx->stream));                       // the host vectors die with this scope
            SplitParams X3;
            memset(&X3, 0, sizeof(X3));
            X3.src = reinterpret_cast<uint32_t *>(ctx->arena); X3.dst = bufA; X3.cap_src = X.slab_items; X3.cap_dst = pl.cap;
            X3.IW = IW; X3.WE = W; X3.mode = 3; X3.sh1 = 32 - (int)lb1; X3.lb2 = lb2;
            X3.in_start = d_xs; X3.in_count = d_xs + (world + 1); X3.chunk_pref = reinterpret_cast<unsigned *>(d_xs + 2 * (world + 1));
            X3.B1 = (unsigned)world; X3.cursor2 = SP.cursor1; X3.ticket = ctx->d_ctr + CTR_TICKET3; X3.T = T; X3.err = ctx->d_ctr + CTR_ERR;
            X3.b_lo = g1_lo; X3.b_hi = g1_lo + nb1; X3.slab_cap = 0; X3.bkt_lo = (unsigned)b0; X3.bkt_hi = (unsigned)b1;
            if ((rc = launch_split(ctx, X3))) return rc;
        } else {
            ItemPartParams IP;
            memset(&IP, 0, sizeof(IP));
            IP.edges = ctx->d_edges; IP.n_edges = ctx->n_edges; IP.k = k; IP.sh1 = 32 - (int)lb1;
            IP.bkt_lo = (unsigned)b0; IP.bkt_hi = (unsigned)b1; IP.cursor1 = SP.cursor1; IP.NB = nb1; IP.b1_lo = g1_lo; IP.dst = bufA; IP.cap = pl.cap;
            IP.err = ctx->d_ctr + CTR_ERR;
            if (launch_item_part(WE, plus, tips_in, IP, ctx->stream))
                FAIL(MGTA_ERR_CUDA, "k_item_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            CK(cudaGetLastError());
            if (tips_in && ctx->n_tips) {
                RowPartParams RP;
                memset(&RP, 0, sizeof(RP));
                RP.rows = ctx->d_tips; RP.n_rows = ctx->n_tips; RP.IW = IW; RP.sh1 = IP.sh1; RP.bkt_lo = IP.bkt_lo; RP.bkt_hi = IP.bkt_hi;
                RP.cursor1 = IP.cursor1; RP.NB = IP.NB; RP.b1_lo = IP.b1_lo; RP.dst = IP.dst; RP.cap = IP.cap; RP.err = IP.err;
                const size_t smem = bin_smem_bytes(IW, ROW_SLOTS);
                CK(cudaFuncSetAttribute(k_row_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_row_part<<<(unsigned)((ctx->n_tips + ROW_SLOTS - 1) / ROW_SLOTS), PART_THREADS, smem, ctx->stream>>>(RP);
                CK(cudaGetLastError());
                st->n_launches++;
            }
        }
        if ((rc = end_timed(ctx))) return rc;
        // ---- K3a: level-2 prefix split, leaf flags, MSD levels for oversize tiles
        SplitParams XP;
        memset(&XP, 0, sizeof(XP));
        XP.src = bufA; XP.dst = bufB; XP.cap_src = pl.cap; XP.cap_dst = pl.cap; XP.IW = IW; XP.WE = W; XP.mode = 1;
        XP.sh2 = 32 - PB; XP.lb2 = lb2; XP.in_start = SP.in_start; XP.in_count = SP.tot; XP.chunk_pref = SP.chunk_pref;
        XP.B1 = nb1; XP.cursor2 = SP.cursor2; XP.ticket = ctx->d_ctr + CTR_TICKET2; XP.T = T; XP.err = ctx->d_ctr + CTR_ERR;
        if ((rc = begin_timed(ctx, PH_PARTITION))) return rc;
        if ((rc = launch_split(ctx, XP))) return rc;
        k_flags_tiles<<<(rt_hi - rt_lo + 255) / 256, 256, 0, ctx->stream>>>(off2, rt_lo, rt_hi, flags);
        CK(cudaGetLastError());
        st->n_launches += 6;
        seg_host.clear();
        {
            uint64_t acc = 0;
            for (unsigned t = t_lo; t < t_hi; ++t) {
                if (h2[t] > pl.T) seg_host.push_back(Seg{acc, h2[t], 0});
                acc += h2[t];
            }
        }
        if (!seg_host.empty()) {
            CK(cudaMemcpyAsync(lists[0], seg_host.data(), seg_host.size() * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
            const unsigned n0 = (unsigned)seg_host.size();
            CK(cudaMemcpyAsync(ctx->d_ctr + CTR_NLIST0, &n0, 4, cudaMemcpyHostToDevice, ctx->stream));
            // digit levels over the (k-1)-mer bits below the tile prefix (never split a group)
            std::vector<std::pair<int, int>> levels;
            for (int s = PB; s < 2 * (k - 1);) {                  // digits end on byte boundaries: never straddle a key word
                const int nb = std::min(8 - (s & 7), 2 * (k - 1) - s);
                levels.push_back({s, nb});
                s += nb;
            }
            if (levels.empty()) {
                k_register_giants<<<(n0 + 255) / 256, 256, 0, ctx->stream>>>(lists[0], n0, pl.C, giants, ctx->d_ctr + CTR_NGIANTS,
                                                                          pl.giants_cap, win, ctx->d_ctr + CTR_ERR);
                CK(cudaGetLastError());
                st->n_launches++;
            }
            for (size_t l = 0; l < levels.size(); ++l) {
                MsdParams MP;
                memset(&MP, 0, sizeof(MP));
                MP.src = bufB; MP.dst = bufA; MP.cap = pl.cap; MP.IW = IW;
                MP.list = lists[l & 1]; MP.n_list = ctx->d_ctr + CTR_NLIST0 + (l & 1);
                MP.word = levels[l].first >> 5;
                MP.shift = 32 - (levels[l].first & 31) - levels[l].second;
                MP.bins = 1 << levels[l].second;
                MP.flags = flags;
                MP.next_list = lists[(l + 1) & 1]; MP.next_count = ctx->d_ctr + CTR_NLIST0 + ((l + 1) & 1);
                MP.next_cap = pl.list_cap;
                MP.last_level = l + 1 == levels.size();
                MP.copy_back = 1;
                MP.T = pl.T; MP.C = pl.C;
                MP.giants = giants; MP.n_giants = ctx->d_ctr + CTR_NGIANTS; MP.giants_cap = pl.giants_cap;
                MP.win_giant = win; MP.err = ctx->d_ctr + CTR_ERR;
                CK(cudaMemsetAsync(ctx->d_ctr + CTR_NLIST0 + ((l + 1) & 1), 0, 4, ctx->stream));
                k_msd<<<(unsigned)ctx->sm_count * 4, MSD_THREADS, 0, ctx->stream>>>(MP);
                CK(cudaGetLastError());
                st->n_launches++;
                st->msd_levels = std::max<int>(st->msd_levels, (int)l + 1);
            }
        }
        if ((rc = end_timed(ctx))) return rc;
        // ---- K3b+K5: on-chip sort + record emission
        ChunkParams CP;
        memset(&CP, 0, sizeof(CP));
        CP.src = bufB; CP.cap = pl.cap; CP.n_items = n_items; CP.W = pl.W; CP.IW = pl.IW; CP.k = pl.k;
        CP.CAPI = pl.CAPI; CP.C = pl.C; CP.n_windows = n_windows; CP.ticket = ctx->d_ctr + CTR_TICKET;
        CP.flags = flags; CP.win_giant = win; CP.giants = giants; CP.depth_min = std::min(PB, 2 * (k - 1));
        CP.n_lsd = ctx->d_ctr + CTR_NOVF;
        CP.big_bin = 64;
        if (const char *e = getenv("MGTA_SORT_BIG_BIN")) CP.big_bin = (unsigned)std::max(1, atoi(e));
        CP.bin_bits = 0;
        while (CP.bin_bits < 12 && (2u << CP.bin_bits) <= pl.CAPI) ++CP.bin_bits;             // bins <= CAPI counters (the code array)
        if (const char *e = getenv("MGTA_SORT_BIN_BITS")) CP.bin_bits = std::max(0, std::min(CP.bin_bits, atoi(e)));   // A/B switch: 0 = LSD only
        CP.g_full = (pl.k - 1) / 16;
        CP.g_rem_shift = (pl.k - 1) % 16 ? (16 - (pl.k - 1) % 16) * 2 : 32;
        CP.m = (unsigned)ctx->opt.min_count;
        CP.aw = (pl.k - 1) >> 4; CP.ash = (15 - ((pl.k - 1) & 15)) * 2; CP.wpt = (2 * pl.k + 31) / 32;
        CP.out = tmpbuf; CP.out_cap = pl.out_cap; CP.state = state; CP.meta = ctx->d_meta; CP.totals = ctx->d_totals;
        CP.err = ctx->d_ctr + CTR_ERR;
        // The windows are sorted and emitted in a few launches.  With a sink, the bytes of part c go to the host on the copy
        // stream while part c + 1 is sorted (a part's bytes are final and contiguous once its scan + gather ran).
        const unsigned n_parts = sink && !ctx->sink_device ? std::max(1u, std::min(8u, n_windows / 2048u)) : 1u;
        bool piped = sink && n_parts > 1 && ctx->h_out_bytes > 0;
        if (piped && !ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        while (piped && ctx->ev_out.size() < n_parts) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->ev_out.push_back(e);
        }
        unsigned long long *h_end = ctx->h_pin + 2 * NUM_BUCKETS + 16;                // [n_parts] running byte totals (pinned)
        unsigned long long *d_end = ctx->d_totals + 16;                             // [8]
        unsigned long long copied = 0;
        auto retire = [&](unsigned c) -> int {                                      // part c has been enqueued: ship its bytes
            if (!piped) return MGTA_OK;
            cudaError_t qe;
            while ((qe = cudaEventQuery(ctx->ev_out[c])) == cudaErrorNotReady) {}
            if (qe != cudaSuccess) { ctx->err = std::string("stage 2 part event: ") + cudaGetErrorString(qe); return MGTA_ERR_CUDA; }
            const unsigned long long end = h_end[c];
            if (end > ctx->h_out_bytes) { piped = false; return MGTA_OK; }           // staging too small: one copy at the end
            if (end > copied) {
                CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_out[c], 0));
                CK(cudaMemcpyAsync(ctx->h_out + copied, outbuf + copied, end - copied, cudaMemcpyDeviceToHost, ctx->copy_stream));
                copied = end;
            }
            return MGTA_OK;
        };
        CK(cudaMemsetAsync(ctx->d_totals + 15, 0, 8, ctx->stream));
        if ((rc = begin_timed(ctx, PH_SORT))) return rc;
        for (unsigned c = 0; c < n_parts; ++c) {
            CP.win_lo = (unsigned)((uint64_t)n_windows * c / n_parts);
            CP.win_hi = (unsigned)((uint64_t)n_windows * (c + 1) / n_parts);
            const unsigned cnt = CP.win_hi - CP.win_lo;
            if (c) CK(cudaMemsetAsync(ctx->d_ctr + CTR_TICKET, 0, 4, ctx->stream));
            const unsigned cgrid = std::max(1u, std::min<unsigned>(cnt, (unsigned)(ctx->sm_count * occ)));
            if (capi_const) { W_SWITCH3(W, (k_sort_emit<WW, 4096><<<cgrid, CHUNK_THREADS, pl.chunk_smem, ctx->stream>>>(CP))); }
            else { W_SWITCH(W, (k_sort_emit<WW, 0><<<cgrid, CHUNK_THREADS, pl.chunk_smem, ctx->stream>>>(CP))); }
            // windows took their space in completion order: scan the byte counts and copy every window to its place in bucket order
            k_out_scan<<<1, 1024, 0, ctx->stream>>>(state, CP.win_lo, CP.win_hi, ctx->d_totals + 15, d_end + c);
            k_out_gather<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(state, CP.win_lo, CP.win_hi, tmpbuf, outbuf);
            CK(cudaGetLastError());
            st->n_launches += 3;
            if (piped) {
                CK(cudaMemcpyAsync(h_end + c, d_end + c, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaEventRecord(ctx->ev_out[c], ctx->stream));
                if (c >= 1 && (rc = retire(c - 1))) return rc;
            }
        }
        if ((rc = end_timed(ctx))) return rc;
        // ---- batch epilogue: error flags, output
        unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
        unsigned long long *h_state = ctx->h_pin + 2 * NUM_BUCKETS + 8;
        CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_state, ctx->d_totals + 15, 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (sink) {
            meta_host.resize((size_t)(b1 - b0) * 3);
            CK(cudaMemcpyAsync(meta_host.data(), ctx->d_meta + (size_t)b0 * 3, (size_t)(b1 - b0) * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (piped && (rc = retire(n_parts - 1))) return rc;
        CK(mgta_stream_wait(ctx->stream));
        st->n_giants += h_ctr[CTR_NGIANTS] + h_ctr[CTR_NOVF];      // windows sorted by the LSD passes (low-complexity or forced)
        const unsigned dev_err = h_ctr[CTR_ERR];
        if (dev_err) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (stage 2, buckets [%d,%d))", dev_err, b0, b1);
        const unsigned long long bytes = *h_state;
        st->out_bytes += bytes;
        if (sink && ctx->sink_device) {                        // the consumer parses the records where they lie
            if (sink(user, b0, b1, outbuf, bytes, meta_host.data()) != 0) FAIL(MGTA_ERR_ARG, "sink aborted");
        } else if (sink) {
            if (piped && copied == bytes) {
                CK(mgta_stream_wait(ctx->copy_stream));
            } else {
                if (ctx->copy_stream) CK(mgta_stream_wait(ctx->copy_stream));
                if (bytes > ctx->h_out_bytes) {
                    cudaFreeHost(ctx->h_out);
                    ctx->h_out = nullptr; ctx->h_out_bytes = 0;
                    CK(cudaHostAlloc(&ctx->h_out, bytes + bytes / 4 + 4096, cudaHostAllocDefault));
                    ctx->h_out_bytes = bytes + bytes / 4 + 4096;
                    copied = 0;
                }
                if (bytes > copied) CK(cudaMemcpyAsync(ctx->h_out + copied, outbuf + copied, bytes - copied, cudaMemcpyDeviceToHost, ctx->stream));
                CK(mgta_stream_wait(ctx->stream));
            }
            if (sink(user, b0, b1, ctx->h_out, bytes, meta_host.data()) != 0) FAIL(MGTA_ERR_ARG, "sink aborted");
        }
        b0 = b1;
    }
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_totals, 10 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    uint64_t edges = 0;
    for (int i = 0; i < 9; ++i) edges += ctx->h_pin[i];
    st->n_edges = edges;
    if (totals) for (int i = 0; i < 10; ++i) totals[i] = (int64_t)ctx->h_pin[i];
    return MGTA_OK;
}

// ---- sharded stage 2: the items of my edges and tips, binned by the shard that emits their lv1 bucket ------------------
// bnd[d] = first bucket of shard d (d = 0 .. world).  Arena: [recv slabs | send slabs | cursors]; the receive slabs stay
// at the start of the arena for run_emit(from_slabs), whose buffers overlay the (then dead) send slabs.
// expect[d]: items this shard must produce for shard d according to its own histogram (consistency check).
int item_exchange_send(mgta_ctx *ctx, const unsigned *bnd, uint64_t slab_in, const uint64_t *expect, mgta_stage_stats *st,
                       void **send_out, void **recv_out, uint64_t *slab_bytes) {
    const int k = ctx->opt.kmer_k, WE = edge_words(k), W = key_words_s2(k), IW = W + 1, world = ctx->opt.world;
    const bool plus = W > WE;
    int rc;
    const uint64_t slab_items = std::max<uint64_t>(32, (slab_in + 31) & ~(uint64_t)31);
    const size_t xbytes = (size_t)world * IW * slab_items * 4, xal = (xbytes + 255) & ~(size_t)255;
    const size_t total = 2 * xal + 4096;
    const size_t budget = hbm_budget(ctx);
    if (total > budget) FAIL(MGTA_ERR_MEM, "HBM budget %zu B cannot hold the stage-2 item exchange (%zu B)", budget, total);
    if ((rc = ensure_arena(ctx, total))) return rc;
    uint32_t *send = reinterpret_cast<uint32_t *>(ctx->arena + xal);
    unsigned long long *cur = reinterpret_cast<unsigned long long *>(ctx->arena + 2 * xal);
    const unsigned long long stride = (unsigned long long)IW * slab_items;
    CK(cudaMemsetAsync(ctx->d_ctr, 0, CTR_COUNT * 4, ctx->stream));
    k_init_slab_cursors<<<1, 256, 0, ctx->stream>>>(cur, (unsigned)world, stride);
    CK(cudaGetLastError());
    if ((rc = begin_timed(ctx, PH_EXTRACT))) return rc;
    if (ctx->n_edges) {
        ItemPartParams IP;
        memset(&IP, 0, sizeof(IP));
        IP.edges = ctx->d_edges; IP.n_edges = ctx->n_edges; IP.k = k; IP.cursor1 = cur; IP.dst = send; IP.cap = slab_items;
        IP.err = ctx->d_ctr + CTR_ERR; IP.n_owner = world; IP.slab_cap = slab_items; IP.slab_stride = stride;
        for (int d = 0; d <= world; ++d) IP.bnd[d] = bnd[d];
        if (launch_item_part(WE, plus, true, IP, ctx->stream)) FAIL(MGTA_ERR_CUDA, "k_item_part launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        CK(cudaGetLastError());
        st->n_launches++;
    }
    if (ctx->n_tips) {
        RowPartParams RP;
        memset(&RP, 0, sizeof(RP));
        RP.rows = ctx->d_tips; RP.n_rows = ctx->n_tips; RP.IW = IW; RP.cursor1 = cur; RP.dst = send; RP.cap = slab_items;
        RP.err = ctx->d_ctr + CTR_ERR; RP.n_owner = world; RP.slab_cap = slab_items; RP.slab_stride = stride;
        for (int d = 0; d <= world; ++d) RP.bnd[d] = bnd[d];
        const size_t smem = bin_smem_bytes(IW, ROW_SLOTS);
        CK(cudaFuncSetAttribute(k_row_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_row_part<<<(unsigned)((ctx->n_tips + ROW_SLOTS - 1) / ROW_SLOTS), PART_THREADS, smem, ctx->stream>>>(RP);
        CK(cudaGetLastError());
        st->n_launches++;
    }
    if ((rc = end_timed(ctx))) return rc;
    unsigned *h_ctr = reinterpret_cast<unsigned *>(ctx->h_pin + 2 * NUM_BUCKETS);
    CK(cudaMemcpyAsync(h_ctr, ctx->d_ctr, CTR_COUNT * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_pin, cur, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));
    if (h_ctr[CTR_ERR]) FAIL(MGTA_ERR_INTERNAL, "device consistency flags 0x%x (stage-2 item exchange)", h_ctr[CTR_ERR]);
    for (int d = 0; d < world; ++d) {
        const uint64_t got = ctx->h_pin[d] - (unsigned long long)d * stride;
        if (got != expect[d]) FAIL(MGTA_ERR_INTERNAL, "stage-2 item exchange: %llu items for shard %d, the histogram says %llu",
                                   (unsigned long long)got, d, (unsigned long long)expect[d]);
    }
    ctx->s2x.valid = false;
    ctx->s2x.slab_items = slab_items;
    ctx->s2x.bytes = xbytes;
    *send_out = send; *recv_out = ctx->arena; *slab_bytes = (uint64_t)IW * slab_items * 4;
    return MGTA_OK;
}

// the two events that time a whole stage; destroyed on every exit path (an error return between stage_begin and
// stage_end used to leak them)
struct StageTimer {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    StageTimer() = default;
    StageTimer(const StageTimer &) = delete;
    StageTimer &operator=(const StageTimer &) = delete;
    ~StageTimer() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

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
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    int rc = ensure_solid(ctx);
    if (rc) return rc;
    rc = histogram(ctx, 2, nullptr);
    if (rc) return rc;
    memcpy(hist, ctx->hist.data(), NUM_BUCKETS * 8);
    mgta_stage_stats tmp;
    memset(&tmp, 0, sizeof(tmp));
    return finish_timing(ctx, &tmp);
}

namespace {
int stage_begin(mgta_ctx *ctx, mgta_stage_stats *st, StageTimer &tm) {
    memset(st, 0, sizeof(*st));
    CK(cudaSetDevice(ctx->opt.device));
    drop_timed(ctx);                                               // phase timers a failed stage left behind
    CK(cudaEventCreate(&tm.ev0));
    CK(cudaEventCreate(&tm.ev1));
    CK(cudaEventRecord(tm.ev0, ctx->stream));
    return MGTA_OK;
}
int stage_end(mgta_ctx *ctx, mgta_stage_stats *st, StageTimer &tm) {
    CK(cudaEventRecord(tm.ev1, ctx->stream));
    CK(mgta_stream_wait(ctx->stream));                               // ev1 is the last thing on the stream
    CK(cudaEventElapsedTime(&st->ms_total, tm.ev0, tm.ev1));
    return finish_timing(ctx, st);                                 // the StageTimer destroys its events
}
}  // namespace

extern "C" int mgta_stage1(mgta_ctx *ctx, int64_t *edge_counting) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    if (ctx->opt.min_count == 1) {
        if (edge_counting) memset(edge_counting, 0, NUM_BUCKETS * 8);
        memset(&ctx->stats[0], 0, sizeof(ctx->stats[0]));
        return MGTA_OK;
    }
    if (ctx->opt.need_mercy && ctx->opt.world > 1) FAIL(MGTA_ERR_ARG, "need_mercy runs on one shard only (world == 1)");
    if (ctx->opt.need_mercy && ctx->opt.min_count > 255) FAIL(MGTA_ERR_ARG, "need_mercy is offered for min_count <= 255");
    mgta_stage_stats *st = &ctx->stats[0];
    StageTimer tm;
    int rc = stage_begin(ctx, st, tm);
    if (rc) return rc;
    ctx->solid_valid = false;
    ctx->stage1_done = false;
    ctx->mercy_valid = false;
    if ((rc = run_count(ctx, CM_STAGE1, st))) return rc;
    ctx->stage1_done = true;
    st->n_edges = ctx->n_edges;                                    // distinct solid edges listed for stage 2
    if (ctx->opt.need_mercy) {
        // mercy edges extend is_solid read by read (s2.cpp:106-250): derive the vector, add the mercy bits, and let stage 2
        // recount the solid occurrences from it (the edge list above no longer describes them)
        if ((rc = ensure_solid(ctx))) return rc;
        if ((rc = run_mercy(ctx, st))) return rc;
        ctx->edges_valid = false;
        ctx->s2x.valid = false;
    }
    if (edge_counting) {
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_ec, NUM_BUCKETS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
        for (int i = 0; i < NUM_BUCKETS; ++i) edge_counting[i] = (int64_t)ctx->h_pin[i];
    }
    return stage_end(ctx, st, tm);
}

static int stage1_slab_items(mgta_ctx *ctx, uint64_t *slab_items) {
    if (!ctx || !slab_items) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    const int world = ctx->opt.world;
    if (world > MAX_OWNERS) FAIL(MGTA_ERR_ARG, "at most %d shards", (int)MAX_OWNERS);
    if (!ctx->slab_suggest) {
        CK(cudaSetDevice(ctx->opt.device));
        { int rcw = wait_reads(ctx); if (rcw) return rcw; }
        unsigned long long *d_out = ctx->d_xs;                       // (MAX_OWNERS + 1) * 24 bytes: room for `world` counters
        CK(cudaMemsetAsync(d_out, 0, (size_t)world * 8, ctx->stream));
        k_count_positions_parts<<<(unsigned)((ctx->n_reads + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_start, ctx->n_reads, ctx->opt.kmer_k,
                                                                                              (unsigned)world, d_out);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pin, d_out, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
        uint64_t mx = 0;
        for (int d = 0; d < world; ++d) mx = std::max<uint64_t>(mx, ctx->h_pin[d]);
        ctx->slab_suggest = ((uint64_t)((double)mx / world * 1.02) + 4096 + 31) & ~(uint64_t)31;
    }
    *slab_items = ctx->slab_suggest;
    return MGTA_OK;
}

extern "C" int mgta_stage2(mgta_ctx *ctx, mgta_bucket_sink sink, void *user, int64_t *totals) {
    if (!ctx) return MGTA_ERR_ARG;
    if (!ctx->d_seq) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    // several shards: this shard's edge list alone would give a self-consistent but partial graph
    if (ctx->opt.world > 1 && ctx->edges_valid && !ctx->edges_complete)
        FAIL(MGTA_ERR_STATE, "this shard holds the solid edges of its hash range only: stage 2 on %d shards runs through mgta_sharded_begin / _step", ctx->opt.world);
    mgta_stage_stats *st = &ctx->stats[1];
    StageTimer tm;
    int rc = stage_begin(ctx, st, tm);
    if (rc) return rc;
    if (!ctx->edges_valid) {
        // no edge list from stage 1 (min_count == 1, or is_solid was set / merged from outside): count the occurrences the
        // is_solid vector calls solid.  Its kernels are reported in the stage-2 statistics.
        mgta_stage_stats tmp;
        memset(&tmp, 0, sizeof(tmp));
        if ((rc = run_count(ctx, CM_GENERAL, &tmp))) return rc;
        st->n_launches += tmp.n_launches;
    }
    if (ctx->node_pass && !ctx->tips_valid && (rc = run_nodes(ctx, st))) return rc;
    if ((rc = run_emit(ctx, sink, user, totals, st))) return rc;
    if (ctx->node_pass && ctx->tips_valid) { st->n_node_ops = 2 * ctx->n_edges; st->n_tip_items = ctx->n_tips; }
    return stage_end(ctx, st, tm);
}

extern "C" int mgta_stage2_into_sdbg(mgta_ctx *ctx, mgta_sdbg *g, int64_t *totals) {
    if (!ctx || !g) return MGTA_ERR_ARG;
    if (ctx->opt.world != 1) FAIL(MGTA_ERR_ARG, "stage2_into_sdbg: one shard only (the in-memory graph is one address space)");
    ctx->sink_device = true;
    const int rc = mgta_stage2(ctx, mgta_sdbg_sink, g, totals);
    ctx->sink_device = false;
    if (rc == MGTA_ERR_ARG && ctx->err == "sink aborted") ctx->err = std::string("sdbg builder: ") + mgta_sdbg_last_error(g);
    return rc;
}

extern "C" int mgta_solid_device_buffer(mgta_ctx *ctx, void **dev_ptr, uint64_t *n_bytes) {
    if (!ctx || !dev_ptr || !n_bytes) return MGTA_ERR_ARG;
    if (!ctx->d_solid) FAIL(MGTA_ERR_STATE, "no reads: call mgta_set_reads first");
    { int rc = ensure_solid(ctx); if (rc) return rc; }
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
    { int rc = ensure_solid(ctx); if (rc) return rc; }
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
    CK(mgta_stream_wait(ctx->stream));
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
    CK(mgta_stream_wait(ctx->stream));
    cudaFree(tmp);
    ctx->edges_valid = false;
    ctx->s2x.valid = false;
    ctx->solid_valid = true;
    return MGTA_OK;
}

extern "C" int mgta_get_mercy_candidates(mgta_ctx *ctx, uint64_t *host, uint64_t cap, uint64_t *n) {
    if (!ctx || !n) return MGTA_ERR_ARG;
    *n = 0;
    if (!ctx->mercy_valid) FAIL(MGTA_ERR_STATE, "no mercy candidates: run mgta_stage1 with need_mercy first");
    *n = ctx->n_cand;
    const uint64_t take = std::min<uint64_t>(cap, ctx->n_cand);
    if (host && take) {
        CK(cudaSetDevice(ctx->opt.device));
        CK(cudaMemcpyAsync(host, ctx->d_cand, take * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(mgta_stream_wait(ctx->stream));
    }
    return MGTA_OK;
}

extern "C" int mgta_get_num_mercy(mgta_ctx *ctx, uint64_t *num_mercy) {
    if (!ctx || !num_mercy) return MGTA_ERR_ARG;
    if (!ctx->mercy_valid) FAIL(MGTA_ERR_STATE, "no mercy result: run mgta_stage1 with need_mercy first");
    *num_mercy = ctx->num_mercy;
    return MGTA_OK;
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

#include "sharded.inc"
