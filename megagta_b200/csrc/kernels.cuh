// kernels.cuh -- sm_100a kernels of the CX1 reads -> SdBG path.
//
//   k_walk       K1/K2: per-base-position item generation from shared-memory-staged read words;
//                MODE_HIST = lv1 bucket histogram (reference s1.cpp:177-229 / s2.cpp:252-315),
//                MODE_SCATTER = extraction straight into bucket-contiguous order (fuses the
//                reference's lv1 offset pass s1.cpp:408-513 / s2.cpp:475-584 with its lv2 extract
//                s1.cpp:515-596 / s2.cpp:586-677).
//   k_msd        K3a: most-significant-digit partition of oversize segments (8 more key bits per
//                level) so that every leaf fits the on-chip sort.
//   k_chunk      K3b+K4/K5: per tile of <= CAP items: multi-word LSD radix sort of a u16
//                permutation in shared memory (warp match/ballot ranking, per-warp digit
//                histograms; replaces lv2_cpu_radix_sort_st lv2_cpu_sort.h:113-151), then stage-1
//                counting + is_solid marking (s1.cpp:671-830) or stage-2 W/last/tip/multiplicity
//                record emission (s2.cpp:742-835, sdbg_multi_io.h:83-112) straight from the sorted
//                tile.  Sorted keys never go back to HBM.
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
constexpr int MAX_PASSES = 40;

enum { MODE_HIST = 0, MODE_SCATTER = 1 };
enum { ERR_NEXT_LIST_FULL = 1, ERR_GIANT_LIST_FULL = 2, ERR_CHUNK_TOO_BIG = 4, ERR_OUT_OVERFLOW = 8, ERR_SORT_ORDER = 256 };

struct WalkParams {
    const uint32_t *seq;
    const uint64_t *start;
    uint64_t n_reads, n_short, total_bases;
    int k, all_solid;
    const uint32_t *solid;          // stage 2: one bit per base position (edge offset o of read r <=> start[r]+o)
    unsigned long long *hist;       // MODE_HIST
    unsigned long long *n_dollar;   // MODE_HIST, stage 2: number of items with a == $ (bounds the tip records)
    unsigned long long *cursor;     // MODE_SCATTER: next free slot per bucket (batch relative)
    uint32_t *dst;
    uint64_t cap;
    int b_lo, b_hi;
};

// largest r in [lo, hi] with start[r] <= g
__device__ __forceinline__ uint64_t find_read(const uint64_t *__restrict__ start, uint64_t lo, uint64_t hi, uint64_t g) {
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(start + mid) <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ bool bit_at(const uint32_t *__restrict__ bits, uint64_t i) {
    return (__ldg(bits + (i >> 5)) >> (i & 31)) & 1u;
}

template <int W, int STAGE, int MODE>
__global__ void __launch_bounds__(WALK_THREADS) k_walk(const WalkParams P) {
    __shared__ __align__(16) uint32_t sw[WALK_SMEM_WORDS];
    __shared__ uint64_t s_r[2];
    const uint64_t g0 = (uint64_t)blockIdx.x * WALK_TILE;
    const uint64_t gend = min(g0 + (uint64_t)WALK_TILE, P.total_bases);
    const uint64_t w_lo = (g0 >> 4) >= WALK_BACK_WORDS ? (g0 >> 4) - WALK_BACK_WORDS : 0;
    // 128-bit coalesced staging of the tile's read words (the device copy of seq is zero padded)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.seq + w_lo);
        uint4 *dst = reinterpret_cast<uint4 *>(sw);
        for (int i = threadIdx.x; i < WALK_SMEM_WORDS / 4; i += WALK_THREADS) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) s_r[0] = find_read(P.start, 0, P.n_reads - 1, g0);
    if (threadIdx.x == 32) s_r[1] = find_read(P.start, 0, P.n_reads - 1, gend - 1);
    __syncthreads();
    const uint64_t r_lo = s_r[0], r_hi = s_r[1];
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
            if (MODE == MODE_HIST) {
                atomicAdd(P.hist + b, 1ull);
                if (STAGE == 2 && !(key[W - 1] & 8u)) ++n_dollar;
            } else if (b >= P.b_lo && b < P.b_hi) {
                const uint64_t pos = atomicAdd(P.cursor + b, 1ull);
#pragma unroll
                for (int w = 0; w < W; ++w) P.dst[(uint64_t)w * P.cap + pos] = key[w];
                if (STAGE == 1) {
                    P.dst[(uint64_t)W * P.cap + pos] = (uint32_t)v;
                    P.dst[(uint64_t)(W + 1) * P.cap + pos] = (uint32_t)(v >> 32);
                }
            }
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

// leaf-start flags for the lv1 buckets of a batch (cursor still holds the batch-relative starts)
__global__ void k_flags_level0(const unsigned long long *__restrict__ cursor, const unsigned long long *__restrict__ sizes,
                               int b_lo, int b_hi, uint32_t *flags) {
    int b = b_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (b < b_hi && sizes[b] > 0) {
        unsigned long long s = cursor[b];
        atomicOr(flags + (s >> 5), 1u << (s & 31));
    }
}

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
    unsigned *ticket;
    const uint32_t *flags;
    const uint32_t *win_giant;
    const Giant *giants;
    int depth_min;
    unsigned *n_lsd;                         // windows that took the LSD passes (statistics)
    unsigned big_bin;                        // largest bin the comparison rank accepts
    int bin_bits;                            // bucket + rank sort: key bits of the counting pass (0 = LSD passes only)
    int n_pass;
    short pass_lsb[MAX_PASSES];
    unsigned char pass_nb[MAX_PASSES];
    int g_full, g_rem_shift;                 // (k-1)-mer compare: full words, shift of the partial word (32 = none)
    // stage 1
    uint32_t *solid;
    unsigned long long *edge_counting;
    unsigned m;
    int need_mercy;
    unsigned long long *mercy_out;
    unsigned long long *mercy_count;
    unsigned long long mercy_cap;
    // stage 2
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

__device__ __forceinline__ unsigned key_digit(const uint32_t *keys, unsigned capi, unsigned idx, int wi, int sh, bool two,
                                              uint32_t mask) {
    uint32_t v = keys[wi * capi + idx] >> sh;
    if (two) v |= keys[(wi - 1) * capi + idx] << (32 - sh);
    return v & mask;
}

// Stable LSD radix sort of the permutation pa[0..n) by the low `sort_bits` key bits.  Each warp owns a
// contiguous slice; per-warp digit histograms give stable global ranks, __match_any_sync ranks inside
// a warp.  Returns with the sorted permutation in S.pa.
__device__ void block_sort(ChunkSmem &S, const ChunkParams &P, unsigned n, int sort_bits) {
    const unsigned capi = P.CAPI, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lt = (1u << lane) - 1;
    const unsigned slice = (((n + CHUNK_WARPS - 1) / CHUNK_WARPS) + 31) & ~31u;
    const unsigned beg = min(n, warp * slice), end = min(n, beg + slice);
    for (int pass = 0; pass < P.n_pass; ++pass) {
        const int lsb = P.pass_lsb[pass], nb = P.pass_nb[pass];
        if (lsb >= sort_bits) break;
        const int wi = P.W - 1 - (lsb >> 5), sh = lsb & 31;
        const uint32_t mask = (1u << nb) - 1;
        const bool two = (sh + nb > 32) && wi > 0;
        for (unsigned i = tid; i < CHUNK_WARPS * 256; i += CHUNK_THREADS) S.whist[i] = 0;
        __syncthreads();
        for (unsigned i0 = beg; i0 < end; i0 += 32) {
            const unsigned i = i0 + lane;
            const bool valid = i < end;
            const unsigned d = valid ? key_digit(S.keys, capi, S.pa[i], wi, sh, two, mask) : (0x1000u + lane);
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
            if (valid && (peers & lt) == 0) S.whist[warp * 256 + d] += (uint16_t)__popc(peers);
            __syncwarp();
        }
        __syncthreads();
        unsigned v = 0, x = 0;
        if (tid < 256) {
            unsigned sum = 0;
#pragma unroll
            for (int w = 0; w < CHUNK_WARPS; ++w) {
                const unsigned c = S.whist[w * 256 + tid];
                S.whist[w * 256 + tid] = (uint16_t)sum;
                sum += c;
            }
            v = sum;
            x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= (unsigned)o) x += y;
            }
            if (lane == 31) S.scan[warp] = x;
        }
        __syncthreads();
        if (tid < 256) {
            unsigned add = 0;
            for (unsigned w = 0; w < warp; ++w) add += S.scan[w];
            const unsigned excl = x - v + add;
#pragma unroll
            for (int w = 0; w < CHUNK_WARPS; ++w) S.whist[w * 256 + tid] += (uint16_t)excl;
        }
        __syncthreads();
        for (unsigned i0 = beg; i0 < end; i0 += 32) {
            const unsigned i = i0 + lane;
            const bool valid = i < end;
            const unsigned idx = valid ? S.pa[i] : 0;
            const unsigned d = valid ? key_digit(S.keys, capi, idx, wi, sh, two, mask) : (0x1000u + lane);
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
            const int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if (valid && (int)lane == leader) {
                base = S.whist[warp * 256 + d];
                S.whist[warp * 256 + d] = (uint16_t)(base + __popc(peers));
            }
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (valid) S.pb[base + __popc(peers & lt)] = (uint16_t)idx;
            __syncwarp();
        }
        __syncthreads();
        uint16_t *t = S.pa; S.pa = S.pb; S.pb = t;
    }
}

// group (k-1)-mer bits differ?
__device__ __forceinline__ bool group_diff(const uint32_t *keys, unsigned capi, unsigned a, unsigned b, int full, int rem_shift) {
    for (int w = 0; w < full; ++w)
        if (keys[w * capi + a] != keys[w * capi + b]) return true;
    if (rem_shift < 32 && (keys[full * capi + a] >> rem_shift) != (keys[full * capi + b] >> rem_shift)) return true;
    return false;
}

__device__ __forceinline__ void boundary_flags(ChunkSmem &S, const ChunkParams &P, unsigned n) {
    const unsigned capi = P.CAPI, lane = threadIdx.x & 31;
    for (unsigned b = (threadIdx.x >> 5) * 32; b < ((n + 31) & ~31u) + 32; b += CHUNK_THREADS) {
        const unsigned i = b + lane;
        bool g = false, r = false;
        if (i < n) {
            if (i == 0) { g = r = true; }
            else {
                const unsigned x = S.pa[i], y = S.pa[i - 1];
                g = group_diff(S.keys, capi, x, y, P.g_full, P.g_rem_shift);
                r = g;
                if (!r)
                    for (int w = P.g_full; w < P.W; ++w)
                        if (S.keys[w * capi + x] != S.keys[w * capi + y]) { r = true; break; }
            }
        }
        const unsigned gb = __ballot_sync(0xFFFFFFFFu, g), rb = __ballot_sync(0xFFFFFFFFu, r);
        if (lane == 0 && (b >> 5) <= (capi >> 5)) { S.gflag[b >> 5] = gb; S.rflag[b >> 5] = rb; }
    }
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

// ---- stage-2 run iterators and sinks over the sorted tile
struct TileRuns {
    const uint32_t *keys; const uint16_t *perm; const uint32_t *rflag;
    unsigned capi, i, e, cur; int W, aw, ash;
    __device__ void reset() { cur = i; }
    __device__ bool next(S2Run &r) {
        if (cur >= e) return false;
        const unsigned idx = perm[cur];
        const unsigned j = next_flag(rflag, cur + 1, e);
        const uint32_t lw = keys[(W - 1) * capi + idx];
        unsigned long long sum = 0;                        // items carry the multiplicity of their edge (payload word W)
        for (unsigned t = cur; t < j; ++t) sum += keys[W * capi + perm[t]];
        r.a = ((lw >> 3) & 1) ? (int)((keys[aw * capi + idx] >> ash) & 3) : SENT;
        r.b = (int)(lw & 7);
        r.cnt = sum > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)sum;
        r.item = idx;
        cur = j;
        return true;
    }
};
struct CountRuns {                      // giant group: runs from the 24-entry (a,b) count table
    const unsigned *cnt; int code;
    __device__ void reset() { code = 0; }
    __device__ bool next(S2Run &r) {
        while (code < 24 && cnt[code] == 0) ++code;
        if (code >= 24) return false;
        if (code < 4) { r.a = SENT; r.b = code; } else { r.a = (code - 4) / 5; r.b = (code - 4) % 5; }
        r.cnt = cnt[code];
        r.item = (uint32_t)code;
        ++code;
        return true;
    }
};
struct SizeSink {
    unsigned bytes; int wpt;
    __device__ void record(int, int, int tip, uint32_t mult, uint32_t) { bytes += s2_record_bytes(tip, mult, wpt); }
};
struct WriteSink {
    unsigned char *p;                   // next byte to write (2-byte aligned)
    const uint32_t *label;              // label word w of item x at label[w * lstride + x]
    unsigned lstride; int wpt;
    unsigned n_items, n_tips, n_large;
    unsigned *s_tot;                    // shared [10]
    __device__ void record(int w, int last, int tip, uint32_t mult, uint32_t item) {
        unsigned short *o = reinterpret_cast<unsigned short *>(p);
        *o++ = s2_record_word(w, last, tip, mult);
        if (mult > 254u) { *o++ = (unsigned short)mult; ++n_large; }
        if (tip) {
            for (int i = 0; i < wpt; ++i) {
                const uint32_t x = label[i * lstride + item];
                *o++ = (unsigned short)(x & 0xFFFFu);
                *o++ = (unsigned short)(x >> 16);
            }
            ++n_tips;
        }
        p = reinterpret_cast<unsigned char *>(o);
        ++n_items;
        atomicAdd(&s_tot[w], 1u);
        if (last) atomicAdd(&s_tot[9], 1u);
    }
};

__device__ __forceinline__ void flush_meta(unsigned long long *meta, unsigned bucket, WriteSink &ws) {
    if (ws.n_items) atomicAdd(meta + bucket * 3 + 0, (unsigned long long)ws.n_items);
    if (ws.n_tips) atomicAdd(meta + bucket * 3 + 1, (unsigned long long)ws.n_tips);
    if (ws.n_large) atomicAdd(meta + bucket * 3 + 2, (unsigned long long)ws.n_large);
    ws.n_items = ws.n_tips = ws.n_large = 0;
}

template <int STAGE>
__global__ void __launch_bounds__(CHUNK_THREADS) k_chunk(const ChunkParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem S;
    const unsigned capi = P.CAPI, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        unsigned char *p = smem_raw;
        S.keys = reinterpret_cast<uint32_t *>(p); p += (size_t)P.IW * capi * 4;
        S.whist = reinterpret_cast<uint16_t *>(p); p += CHUNK_WARPS * 256 * 2;
        S.tot = reinterpret_cast<uint32_t *>(p); p += 256 * 4;
        S.gflag = reinterpret_cast<uint32_t *>(p); p += (capi / 32 + 2) * 4;
        S.rflag = reinterpret_cast<uint32_t *>(p); p += (capi / 32 + 2) * 4;
        S.scan = reinterpret_cast<uint32_t *>(p); p += (CHUNK_WARPS + 2) * 4;
        S.pa = reinterpret_cast<uint16_t *>(p); p += capi * 2;
        S.pb = reinterpret_cast<uint16_t *>(p);
    }
    __shared__ unsigned long long s_lo, s_hi, s_base;
    __shared__ unsigned s_j, s_gcnt[64], s_tot10[10], s_ec[256], s_gbytes;
    __shared__ uint32_t s_glabel[4 * 9];
    __shared__ int s_sortbits;
    for (unsigned i = tid; i < 256; i += CHUNK_THREADS) s_ec[i] = 0;
    if (tid < 10) s_tot10[tid] = 0;
    __syncthreads();

    while (true) {
        if (tid == 0) s_j = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const unsigned j = s_j;
        if (j >= P.n_windows) break;
        if (warp == 0) { unsigned long long v = window_lo(P, j); if (lane == 0) s_lo = v; }
        if (warp == 1) { unsigned long long v = window_lo(P, j + 1); if (lane == 0) s_hi = v; }
        if (tid == 64) s_gbytes = 0;
        __syncthreads();
        unsigned long long lo = s_lo;
        const unsigned long long hi = s_hi;
        bool bad = false;
        // ---- giant group leading the chunk: counted in place, never sorted
        bool giant = false;
        if (hi > lo && hi - lo > capi) {
            const unsigned gi = P.win_giant[j];
            if (gi && P.giants[gi - 1].start == lo) giant = true;
            else { bad = true; if (tid == 0) atomicOr(P.err, (unsigned)ERR_CHUNK_TOO_BIG); }
        }
        unsigned long long g_lo = 0, g_hi = 0;
        if (giant) {
            g_lo = lo; g_hi = P.giants[P.win_giant[j] - 1].end;
            if (tid < 64) s_gcnt[tid] = 0;
            __syncthreads();
            const uint32_t *lastw = P.src + (uint64_t)(P.W - 1) * P.cap;
            for (unsigned long long b0 = g_lo; b0 < g_hi; b0 += CHUNK_THREADS) {
                const unsigned long long i = b0 + tid;
                const bool valid = i < g_hi;
                unsigned code = 0x1000u + lane;
                if (valid) {
                    const uint32_t lw = lastw[i];
                    if (STAGE == 1) code = lw & 63;
                    else {
                        const int b = lw & 7;
                        if ((lw >> 3) & 1) code = 4 + 5 * ((P.src[(uint64_t)P.aw * P.cap + i] >> P.ash) & 3) + b; else code = b;
                    }
                }
                if (STAGE == 2) {
                    if (valid) {
                        atomicAdd(&s_gcnt[code], P.src[(uint64_t)P.W * P.cap + i]);
                        if (code < 4)
                            for (int w = 0; w < P.wpt; ++w) s_glabel[w * 4 + code] = P.src[(uint64_t)w * P.cap + i];
                    }
                } else {
                    const unsigned peers = __match_any_sync(0xFFFFFFFFu, code);
                    if (valid && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&s_gcnt[code], (unsigned)__popc(peers));
                }
            }
            __syncthreads();
            if (STAGE == 1) {
                if (tid < 64) {
                    const unsigned c = s_gcnt[tid];
                    if (c && (tid >> 3) != SENT && (tid & 7) != SENT) {
                        if (c < 256) atomicAdd(&s_ec[c], 1u); else atomicAdd(P.edge_counting + (c < 65535u ? c : 65535u), 1ull);
                    }
                }
                for (unsigned long long i = g_lo + tid; i < g_hi; i += CHUNK_THREADS) {
                    const unsigned ht = lastw[i] & 63;
                    if ((ht >> 3) != SENT && (ht & 7) != SENT && s_gcnt[ht] >= P.m) {
                        const uint32_t v0 = P.src[(uint64_t)P.W * P.cap + i], v1 = P.src[(uint64_t)(P.W + 1) * P.cap + i];
                        const unsigned long long kpos = ((unsigned long long)v1 << 24) | (v0 >> 8);
                        if (kpos != S1_NO_EDGE) { const unsigned long long e = kpos - 1; atomicOr(P.solid + (e >> 5), 1u << (e & 31)); }
                    }
                }
            } else if (tid == 0) {
                CountRuns runs{s_gcnt, 0};
                SizeSink sz{0, P.wpt};
                s2_emit_group(runs, sz);
                s_gbytes = sz.bytes;
            }
            __syncthreads();
            lo = g_hi;
        }
        const unsigned n = bad ? 0u : (unsigned)(hi - lo);
        // ---- load the tile (coalesced per word array)
        for (int w = 0; w < P.IW; ++w) {
            const uint32_t *s = P.src + (uint64_t)w * P.cap + lo;
            for (unsigned i = tid; i < n; i += CHUNK_THREADS) S.keys[w * capi + i] = s[i];
        }
        for (unsigned i = tid; i < n; i += CHUNK_THREADS) S.pa[i] = (uint16_t)i;
        __syncthreads();
        if (tid == 0) {
            int cpl = 32 * P.W;
            if (n > 1)
                for (int w = 0; w < P.W; ++w) {
                    const uint32_t x = S.keys[w * capi] ^ S.keys[w * capi + n - 1];
                    if (x) { cpl = 32 * w + __clz(x); break; }
                }
            s_sortbits = 32 * P.W - min(cpl, P.depth_min);
        }
        __syncthreads();
        if (n > 1) block_sort(S, P, n, s_sortbits);
        boundary_flags(S, P, n);
        __syncthreads();

        if (STAGE == 1) {
            for (unsigned i = tid; i < n; i += CHUNK_THREADS) {
                if (!((S.rflag[i >> 5] >> (i & 31)) & 1)) continue;
                const unsigned e = next_flag(S.rflag, i + 1, n), cnt = e - i;
                const unsigned ht = S.keys[(P.W - 1) * capi + S.pa[i]] & 63;
                if ((ht >> 3) == SENT || (ht & 7) == SENT) continue;
                if (cnt < 256) atomicAdd(&s_ec[cnt], 1u); else atomicAdd(P.edge_counting + (cnt < 65535u ? cnt : 65535u), 1ull);
                if (cnt >= P.m)
                    for (unsigned t = i; t < e; ++t) {
                        const unsigned id = S.pa[t];
                        const uint32_t v0 = S.keys[P.W * capi + id], v1 = S.keys[(P.W + 1) * capi + id];
                        const unsigned long long kpos = ((unsigned long long)v1 << 24) | (v0 >> 8);
                        if (kpos != S1_NO_EDGE) { const unsigned long long eb = kpos - 1; atomicOr(P.solid + (eb >> 5), 1u << (eb & 31)); }
                    }
            }
            __syncthreads();
        } else {
            // ---- sizing pass: thread t owns the groups that START in its block of sorted positions
            const unsigned Q = (n + CHUNK_THREADS - 1) / CHUNK_THREADS;
            const unsigned pb = min(n, tid * Q), pe = min(n, pb + Q);
            SizeSink sz{0, P.wpt};
            for (unsigned pos = next_flag(S.gflag, pb, pe); pos < pe; pos = next_flag(S.gflag, pos + 1, pe)) {
                TileRuns runs{S.keys, S.pa, S.rflag, capi, pos, next_flag(S.gflag, pos + 1, n), pos, P.W, P.aw, P.ash};
                s2_emit_group(runs, sz);
            }
            unsigned total = 0;
            const unsigned my_off = block_exclusive_scan(S, sz.bytes, total);
            const unsigned gbytes = s_gbytes;
            // ---- decoupled look-back over windows for the chunk's byte offset in the batch stream
            if (tid == 0) {
                const unsigned long long mine = (unsigned long long)total + gbytes;
                unsigned long long prefix = 0;
                if (j == 0) {
                    atomicExch(P.state + j, (2ull << 62) | mine);
                } else {
                    atomicExch(P.state + j, (1ull << 62) | mine);
                    long long q = (long long)j - 1;
                    while (true) {
                        unsigned long long s;
                        do { s = *reinterpret_cast<volatile unsigned long long *>(P.state + q); } while ((s >> 62) == 0);
                        prefix += s & ((1ull << 62) - 1);
                        if ((s >> 62) == 2) break;
                        --q;
                    }
                    atomicExch(P.state + j, (2ull << 62) | (prefix + mine));
                }
                s_base = prefix;
                if (prefix + mine > P.out_cap) atomicOr(P.err, (unsigned)ERR_OUT_OVERFLOW);
            }
            __syncthreads();
            const unsigned long long base = s_base;
            if (base + total + gbytes <= P.out_cap) {
                if (giant && tid == 0) {
                    CountRuns runs{s_gcnt, 0};
                    WriteSink ws{P.out + base, s_glabel, 4u, P.wpt, 0, 0, 0, s_tot10};
                    s2_emit_group(runs, ws);
                    flush_meta(P.meta, P.src[g_lo] >> 16, ws);
                }
                WriteSink ws{P.out + base + gbytes + my_off, S.keys, capi, P.wpt, 0, 0, 0, s_tot10};
                for (unsigned pos = next_flag(S.gflag, pb, pe); pos < pe; pos = next_flag(S.gflag, pos + 1, pe)) {
                    TileRuns runs{S.keys, S.pa, S.rflag, capi, pos, next_flag(S.gflag, pos + 1, n), pos, P.W, P.aw, P.ash};
                    s2_emit_group(runs, ws);
                    flush_meta(P.meta, S.keys[S.pa[pos]] >> 16, ws);
                }
            }
            __syncthreads();
        }
    }
    // ---- per-CTA flush of the small counters
    __syncthreads();
    if (STAGE == 1) {
        for (unsigned i = tid; i < 256; i += CHUNK_THREADS)
            if (s_ec[i]) atomicAdd(P.edge_counting + i, (unsigned long long)s_ec[i]);
    } else if (tid < 10 && s_tot10[tid]) {
        atomicAdd(P.totals + tid, (unsigned long long)s_tot10[tid]);
    }
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
