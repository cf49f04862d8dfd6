// kernels.cuh -- shared pieces of the sm_100a kernels of the CX1 reads -> SdBG path.
//
//   k_walk       per-base-position item generation from shared-memory-staged read words, MODE_HIST: the reference's
//                lv1 bucket histograms (s1.cpp:177-229 / s2.cpp:252-315) behind mgta_stage{1,2}_histogram.
//   k_msd        most-significant-digit partition of oversize stage-2 prefix tiles (up to 8 more key bits per level) so
//                that every leaf fits the on-chip window of k_sort_emit (emit_kernels.cuh).
//   window / flag / scan helpers and the parameter blocks of k_sort_emit; is_solid layout conversion.
// The hot path lives in v2_kernels.cuh (extraction, partition, counting), emit_kernels.cuh (on-chip sort + record
// emission) and mercy_kernels.cuh (need_mercy).
//
// Item arrays are SoA: word w of item i lives at buf[w * cap + i] (coalesced per word).
#pragma once
#include <cuda_runtime.h>
#include "cx1_emit.cuh"

namespace mgta {

constexpr int NUM_BUCKETS = 65536;
constexpr int WALK_TILE = 2048;       // base positions per CTA
constexpr int WALK_THREADS = 256;
constexpr int WALK_BACK_WORDS = 4;    // staged words before the tile (prev/head chars, keeps 16 B alignment)
constexpr int WALK_SMEM_WORDS = WALK_TILE / 16 + WALK_BACK_WORDS + 12;   // + (k+1 <= 128 chars) + 1 lookahead, mult of 4
constexpr int SEQ_PAD_WORDS = 256;    // zero words appended to the device copy of packed_seq
constexpr int MSD_THREADS = 512;
constexpr int CHUNK_THREADS = 512;
constexpr int CHUNK_WARPS = CHUNK_THREADS / 32;

enum { MODE_HIST = 0 };
enum { ERR_NEXT_LIST_FULL = 1, ERR_GIANT_LIST_FULL = 2, ERR_CHUNK_TOO_BIG = 4, ERR_OUT_OVERFLOW = 8, ERR_SORT_ORDER = 256 };

struct WalkParams {
    const uint32_t *seq;
    const uint64_t *start;
    const uint32_t *lut;            // read lookup table (k_build_read_lut)
    uint64_t n_lut;
    uint64_t n_reads, n_short, total_bases;
    int k, all_solid;
    const uint32_t *solid;          // stage 2: one bit per base position (edge offset o of read r <=> start[r]+o)
    unsigned long long *hist;       // MODE_HIST
    unsigned long long *n_dollar;   // MODE_HIST, stage 2: number of items with a == $ (bounds the tip records)
};

// largest r in [lo, hi] with start[r] <= g
__device__ __forceinline__ uint64_t find_read(const uint64_t *__restrict__ start, uint64_t lo, uint64_t hi, uint64_t g) {
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(start + mid) <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Read lookup table (the device counterpart of SequencePackage::pos_to_id_, sequence_package.h:145-188):
// lut[i] = largest r with start[r] <= i * 1024, for i in [0, (total_bases >> 10) + 2).  A CTA tile reads two entries
// instead of running two binary searches over all of start_idx with dependent global loads (measured: 42 % of the stall
// samples of k_edge_part sat on the barrier behind those searches).
constexpr int READ_LUT_SHIFT = 10;

__global__ void __launch_bounds__(256) k_build_read_lut(const uint64_t *__restrict__ start, uint64_t n_reads, uint64_t n_lut, uint32_t *lut) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n_lut) lut[i] = (uint32_t)find_read(start, 0, n_reads - 1, i << READ_LUT_SHIFT);
}

// reads of the tile [g0, gend): r_lo = the read of base g0 (exact when g0 is a multiple of 1024, else a lower bound),
// r_hi >= the read of base gend - 1
__device__ __forceinline__ void tile_read_span(const uint32_t *__restrict__ lut, uint64_t n_lut, uint64_t g0, uint64_t gend,
                                               uint64_t &r_lo, uint64_t &r_hi) {
    r_lo = __ldg(lut + (g0 >> READ_LUT_SHIFT));
    const uint64_t j = (gend + ((1u << READ_LUT_SHIFT) - 1)) >> READ_LUT_SHIFT;
    r_hi = __ldg(lut + (j < n_lut - 1 ? j : n_lut - 1));
}

__device__ __forceinline__ bool bit_at(const uint32_t *__restrict__ bits, uint64_t i) {
    return (__ldg(bits + (i >> 5)) >> (i & 31)) & 1u;
}

template <int W, int STAGE, int MODE>
__global__ void __launch_bounds__(WALK_THREADS) k_walk(const WalkParams P) {
    __shared__ __align__(16) uint32_t sw[WALK_SMEM_WORDS];
    const uint64_t g0 = (uint64_t)blockIdx.x * WALK_TILE;
    const uint64_t gend = min(g0 + (uint64_t)WALK_TILE, P.total_bases);
    const uint64_t w_lo = (g0 >> 4) >= WALK_BACK_WORDS ? (g0 >> 4) - WALK_BACK_WORDS : 0;
    // 128-bit coalesced staging of the tile's read words (the device copy of seq is zero padded)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.seq + w_lo);
        uint4 *dst = reinterpret_cast<uint4 *>(sw);
        for (int i = threadIdx.x; i < WALK_SMEM_WORDS / 4; i += WALK_THREADS) dst[i] = __ldg(src + i);
    }
    uint64_t r_lo, r_hi;
    tile_read_span(P.lut, P.n_lut, g0, gend, r_lo, r_hi);
    __syncthreads();
    const int k = P.k;
    unsigned n_dollar = 0;
#pragma unroll 1
    for (int it = 0; it < WALK_TILE / WALK_THREADS; ++it) {
        const uint64_t g = g0 + (uint64_t)it * WALK_THREADS + threadIdx.x;
        if (g >= gend) continue;
        const uint64_t r = find_read(P.start, r_lo, r_hi, g);
        const uint64_t s = __ldg(P.start + r);
        const int64_t Lw = (int64_t)(__ldg(P.start + r + 1) - s);
        const int64_t pw = (int64_t)(g - s);
        if (Lw < k + 1) continue;
        const int L = (int)Lw, p = (int)pw;
        const uint32_t q = (uint32_t)(g - 16 * w_lo);
        auto sink = [&](const uint32_t(&key)[W], uint64_t v) {
            const int b = (int)(key[0] >> 16);
            atomicAdd(P.hist + b, 1ull);
            if (STAGE == 2 && !(key[W - 1] & 8u)) ++n_dollar;
        };
        if (STAGE == 1) {
            if (p > L - k + 1) continue;
            s1_position<W>(sw, q, g, p, L, k, r < P.n_short, sink);
        } else {
            if (p >= L - k) continue;
            const bool all = P.all_solid || r >= P.n_short;
            if (!(all || bit_at(P.solid, g))) continue;
            const bool sp = p > 0 && (all || bit_at(P.solid, g - 1));
            const bool sn = p < L - k - 1 && (all || bit_at(P.solid, g + 1));
            s2_position<W>(sw, q, p, L, k, sp, sn, [&](const uint32_t(&key)[W]) { sink(key, 0); });
        }
    }
    if (STAGE == 2 && MODE == MODE_HIST) {
        n_dollar = __reduce_add_sync(0xFFFFFFFFu, n_dollar);
        if ((threadIdx.x & 31) == 0 && n_dollar) atomicAdd(P.n_dollar, (unsigned long long)n_dollar);
    }
}

// ------------------------------------------------------------------------------------------------
struct Seg {
    unsigned long long start;
    unsigned int count, pad;
};
struct Giant {
    unsigned long long start, end;
};

__device__ __forceinline__ void register_giant(unsigned long long g, unsigned long long g2, unsigned C, Giant *giants,
                                               unsigned *n_giants, unsigned giants_cap, uint32_t *win_giant, unsigned *err) {
    unsigned idx = atomicAdd(n_giants, 1u);
    if (idx >= giants_cap) { atomicOr(err, (unsigned)ERR_GIANT_LIST_FULL); return; }
    giants[idx].start = g;
    giants[idx].end = g2;
    // windows whose start lies strictly inside the giant resolve their lower bound to g
    for (unsigned long long j = g / C + 1; j * C < g2; ++j) win_giant[j] = idx + 1;
}

__global__ void k_register_giants(const Seg *list, unsigned n, unsigned C, Giant *giants, unsigned *n_giants,
                                  unsigned giants_cap, uint32_t *win_giant, unsigned *err) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) register_giant(list[i].start, list[i].start + list[i].count, C, giants, n_giants, giants_cap, win_giant, err);
}

struct MsdParams {
    const uint32_t *src;
    uint32_t *dst;
    uint64_t cap;
    int IW;
    const Seg *list;
    const unsigned *n_list;
    int word, shift, bins;
    uint32_t *flags;
    Seg *next_list;
    unsigned *next_count;
    unsigned next_cap;
    int last_level, copy_back;
    unsigned T, C;
    Giant *giants;
    unsigned *n_giants;
    unsigned giants_cap;
    uint32_t *win_giant;
    unsigned *err;
};

// One CTA per segment: histogram of the next digit, exclusive scan, scatter src -> dst (same index
// range), leaf flags for the children, oversize children queued for the next level.
__global__ void __launch_bounds__(MSD_THREADS) k_msd(const MsdParams P) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_cur[256];
    __shared__ unsigned s_wsum[8];
    const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const unsigned n_list = *P.n_list;
    for (unsigned si = blockIdx.x; si < n_list; si += gridDim.x) {
        const Seg sg = P.list[si];
        const uint32_t *dsrc = P.src + (uint64_t)P.word * P.cap + sg.start;
        for (int i = threadIdx.x; i < 256; i += MSD_THREADS) s_hist[i] = 0;
        __syncthreads();
        // pass A: digit histogram (warp-aggregated shared atomics: skewed digits are the common case)
        for (unsigned base = 0; base < sg.count; base += MSD_THREADS) {
            const unsigned i = base + threadIdx.x;
            const bool valid = i < sg.count;
            const unsigned d = valid ? ((dsrc[i] >> P.shift) & (unsigned)(P.bins - 1)) : (0x1000u + lane);
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
            if (valid && (peers & lt) == 0) atomicAdd(&s_hist[d], (unsigned)__popc(peers));
        }
        __syncthreads();
        // exclusive scan of 256 bins
        unsigned v = 0, x = 0;
        if (threadIdx.x < 256) {
            v = s_hist[threadIdx.x];
            x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= (unsigned)o) x += y;
            }
            if (lane == 31) s_wsum[threadIdx.x >> 5] = x;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            unsigned add = 0;
            for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) add += s_wsum[w];
            const unsigned excl = x - v + add;
            s_cur[threadIdx.x] = excl;
            if (v > 0 && (int)threadIdx.x < P.bins) {
                const unsigned long long cs = sg.start + excl;
                atomicOr(P.flags + (cs >> 5), 1u << (cs & 31));
                if (v > P.T) {
                    if (!P.last_level) {
                        unsigned slot = atomicAdd(P.next_count, 1u);
                        if (slot < P.next_cap) { P.next_list[slot].start = cs; P.next_list[slot].count = v; P.next_list[slot].pad = 0; }
                        else atomicOr(P.err, (unsigned)ERR_NEXT_LIST_FULL);
                    } else {
                        register_giant(cs, cs + v, P.C, P.giants, P.n_giants, P.giants_cap, P.win_giant, P.err);
                    }
                }
            }
        }
        __syncthreads();
        // pass B: scatter all item words
        for (unsigned base = 0; base < sg.count; base += MSD_THREADS) {
            const unsigned i = base + threadIdx.x;
            const bool valid = i < sg.count;
            const unsigned d = valid ? ((dsrc[i] >> P.shift) & (unsigned)(P.bins - 1)) : (0x1000u + lane);
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
            const int leader = __ffs(peers) - 1;
            unsigned pos = 0;
            if (valid && (int)lane == leader) pos = atomicAdd(&s_cur[d], (unsigned)__popc(peers));
            pos = __shfl_sync(0xFFFFFFFFu, pos, leader) + __popc(peers & lt);
            if (valid) {
                for (int w = 0; w < P.IW; ++w)
                    P.dst[(uint64_t)w * P.cap + sg.start + pos] = P.src[(uint64_t)w * P.cap + sg.start + i];
            }
        }
        __syncthreads();
        if (P.copy_back) {
            uint32_t *back = const_cast<uint32_t *>(P.src);
            for (int w = 0; w < P.IW; ++w)
                for (unsigned i = threadIdx.x; i < sg.count; i += MSD_THREADS)
                    back[(uint64_t)w * P.cap + sg.start + i] = P.dst[(uint64_t)w * P.cap + sg.start + i];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct ChunkParams {
    const uint32_t *src;
    uint64_t cap, n_items;
    int W, IW, k;
    unsigned CAPI, C, n_windows;
    unsigned win_lo, win_hi;                 // windows of this launch (a batch is emitted in a few launches so that the D2H
                                             // copy of one part overlaps the sort of the next)
    unsigned *ticket;
    const uint32_t *flags;
    const uint32_t *win_giant;
    const Giant *giants;
    int depth_min;
    unsigned *n_lsd;                         // windows that took the LSD passes (statistics)
    unsigned big_bin;                        // largest bin the comparison rank accepts
    int bin_bits;                            // bucket + rank sort: key bits of the counting pass (0 = LSD passes only)
    int g_full, g_rem_shift;                 // (k-1)-mer compare: full words, shift of the partial word (32 = none)
    unsigned m;
    int aw, ash, wpt;
    unsigned char *out;
    unsigned long long out_cap;
    unsigned long long *state;               // decoupled look-back: [63:62] 0 empty 1 aggregate 2 prefix | bytes
    unsigned long long *meta;                // [65536][3]
    unsigned long long *totals;              // [10]
    unsigned *err;
};

__device__ __forceinline__ unsigned next_flag(const uint32_t *f, unsigned x, unsigned n) {
    if (x >= n) return n;
    unsigned w = x >> 5;
    const unsigned nw = (n + 31) >> 5;
    uint32_t v = f[w] & (0xFFFFFFFFu << (x & 31));
    while (!v) {
        if (++w >= nw) return n;
        v = f[w];
    }
    const unsigned r = w * 32 + __ffs(v) - 1;
    return r < n ? r : n;
}

// lower bound of window j: the last leaf start <= j*C (warp-cooperative backward scan of the flag bits)
__device__ __forceinline__ unsigned long long window_lo(const ChunkParams &P, unsigned j) {
    if (j >= P.n_windows) return P.n_items;
    const unsigned gi = P.win_giant[j];
    if (gi) return P.giants[gi - 1].start;
    const unsigned long long pos = (unsigned long long)j * P.C;
    const unsigned lane = threadIdx.x & 31;
    long long w = (long long)(pos >> 5);
    uint32_t first_mask = (pos & 31) == 31 ? 0xFFFFFFFFu : ((1u << ((pos & 31) + 1)) - 1);
    while (true) {
        const long long wi = w - lane;
        uint32_t v = wi >= 0 ? __ldg(P.flags + wi) : 0u;
        if (lane == 0) v &= first_mask;
        if (wi == 0) v |= 1u;                       // item 0 always starts a leaf
        const unsigned hit = __ballot_sync(0xFFFFFFFFu, v != 0);
        if (hit) {
            const int src = __ffs(hit) - 1;
            const uint32_t vv = __shfl_sync(0xFFFFFFFFu, v, src);
            return (unsigned long long)(w - src) * 32 + (31 - __clz(vv));
        }
        first_mask = 0xFFFFFFFFu;
        w -= 32;
    }
}

struct ChunkSmem {
    uint32_t *keys;       // [IW][CAPI]
    uint16_t *pa, *pb;    // [CAPI] each
    uint16_t *whist;      // [CHUNK_WARPS][256]
    uint32_t *tot;        // [256]
    uint32_t *gflag, *rflag;   // [CAPI/32 + 1]
    uint32_t *scan;       // [CHUNK_WARPS + 1]
};

__device__ __forceinline__ bool group_diff(const uint32_t *keys, unsigned capi, unsigned a, unsigned b, int full, int rem_shift) {
    for (int w = 0; w < full; ++w)
        if (keys[w * capi + a] != keys[w * capi + b]) return true;
    if (rem_shift < 32 && (keys[full * capi + a] >> rem_shift) != (keys[full * capi + b] >> rem_shift)) return true;
    return false;
}

__device__ __forceinline__ unsigned block_exclusive_scan(ChunkSmem &S, unsigned v, unsigned &total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    __syncthreads();
    if (lane == 31) S.scan[warp] = x;
    __syncthreads();
    unsigned add = 0, tot = 0;
    for (unsigned w = 0; w < CHUNK_WARPS; ++w) {
        const unsigned s = S.scan[w];
        if (w < warp) add += s;
        tot += s;
    }
    total = tot;
    return x - v + add;
}

// is_solid between our layout (bit per base position) and the reference's ((max_len-k)*read + offset)
__global__ void k_solid_export(const uint32_t *__restrict__ solid, const uint64_t *__restrict__ start, uint64_t n_short, int k,
                               int nk1, uint32_t *out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_short) return;
    const uint64_t s = start[r];
    const int L = (int)(start[r + 1] - s);
    for (int o = 0; o < L - k; ++o)
        if ((solid[(s + o) >> 5] >> ((s + o) & 31)) & 1) {
            const uint64_t bit = (uint64_t)nk1 * r + o;
            atomicOr(out + (bit >> 5), 1u << (bit & 31));
        }
}
__global__ void k_solid_import(const uint32_t *__restrict__ in, const uint64_t *__restrict__ start, uint64_t n_short, int k,
                               int nk1, uint32_t *solid) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_short) return;
    const uint64_t s = start[r];
    const int L = (int)(start[r + 1] - s);
    for (int o = 0; o < L - k; ++o) {
        const uint64_t bit = (uint64_t)nk1 * r + o;
        if ((in[bit >> 5] >> (bit & 31)) & 1) atomicOr(solid + ((s + o) >> 5), 1u << ((s + o) & 31));
    }
}

}  // namespace mgta
