// mercy_kernels.cuh -- mercy edges (`--need_mercy`) on the device: candidate emission of stage 1
// (reference s1_lv2_output_ s1.cpp:671-830) and the per-read scan that turns candidates into extra
// is_solid bits (s2_read_mercy_prepare s2.cpp:106-250).
//
// The candidate rules need, for every (k-1)-mer S, the (prev, head) / (tail, next) / (head, tail) count tables
// over ALL items of S on the strand rule of s1.cpp:482-495 (read ends on both strands), i.e. the reference's own
// stage-1 items (cx1_items.cuh s1_position), not the canonical (k+1)-mers of the fast counting path.  Only
// GROUPING by S is needed, never order, so the items take the same route as the fast path: hash partition into
// tiles (k_ctx_part -> k_split mode 0 with the head/tail flag bits masked out of the hash), one shared-memory
// hash table per tile (k_mercy) holding the three 4x4 tables as saturating bytes, group masks, then one
// classification per item (cx1_emit.cuh s1_mercy_item, the code the CPU logic test pins on the oracle).
// The per-read scan needs no sort either: candidate flags are idempotent, so three bit vectors over base
// positions (no-in, no-out, touched) replace the reference's sorted candidate files.
#pragma once
#include "v2_kernels.cuh"
#include "cx1_emit.cuh"

namespace mgta {

constexpr uint32_t S1_FLAG_MASK = ~63u;          // key without head<<3|tail (s1.cpp:575-588)

struct CtxPartParams {
    const uint32_t *seq;
    const uint64_t *start;
    const uint32_t *lut;              // read lookup table (k_build_read_lut)
    uint64_t n_lut;
    uint64_t n_reads, n_short, total_bases;
    int k;
    int sh1, sh2;
    unsigned lb2, b_lo, b_hi;
    unsigned long long *cursor1;
    unsigned long long slab_cap;
    uint32_t *hist2;
    uint32_t *dst;
    uint64_t cap;
    unsigned *err;
    uint32_t ha_lo, ha_last;          // only (k-1)-mers with hash in [ha_lo, ha_last] (a shard's slice of the hash space; all: 0, 2^32 - 1)
};

// One CTA = TP base positions -> <= 2 TP stage-1 items (key W words | value lo, hi), binned by the hash of S.
template <int W, int TP>
__global__ void __launch_bounds__(PART_THREADS) k_ctx_part(const CtxPartParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int IW = W + 2, SLOTS = 2 * TP;
    constexpr int SW_WORDS = TP / 16 + WALK_BACK_WORDS + 12;
    __shared__ __align__(16) uint32_t sw[SW_WORDS];
    BinSmem S;
    bin_smem_carve(S, smem_raw, IW, SLOTS);
    const int tid = threadIdx.x;
    const int NB = (int)(P.b_hi - P.b_lo);
    const uint64_t g0 = (uint64_t)blockIdx.x * TP;
    const uint64_t gend = min(g0 + (uint64_t)TP, P.total_bases);
    const uint64_t w_lo = (g0 >> 4) >= WALK_BACK_WORDS ? (g0 >> 4) - WALK_BACK_WORDS : 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.seq + w_lo);
        uint4 *dstw = reinterpret_cast<uint4 *>(sw);
        for (int i = tid; i < SW_WORDS / 4; i += PART_THREADS) dstw[i] = __ldg(src + i);
    }
    for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
    for (int i = tid; i < SLOTS; i += PART_THREADS) S.bin[i] = 0xFFFFu;
    uint64_t r_lo, r_hi;
    tile_read_span(P.lut, P.n_lut, g0, gend, r_lo, r_hi);
    __syncthreads();
    const unsigned sub_mask = (1u << P.lb2) - 1u;
    for (int i = tid; i < TP; i += PART_THREADS) {
        const uint64_t g = g0 + (uint64_t)i;
        if (g >= gend) break;
        const uint64_t r = find_read(P.start, r_lo, r_hi, g);
        const uint64_t s0 = __ldg(P.start + r);
        const int L = (int)(__ldg(P.start + r + 1) - s0), p = (int)(g - s0);
        if (L < P.k + 1 || p > L - P.k + 1) continue;
        int n_out = 0;
        s1_position<W>(sw, (uint32_t)(g - 16 * w_lo), g, p, L, P.k, r < P.n_short, [&](const uint32_t(&key)[W], uint64_t val) {
            uint32_t ha, hb;
            edge_hash([&](int w) { return w == W - 1 ? key[w] & S1_FLAG_MASK : key[w]; }, W, ha, hb);
            const unsigned b1 = ha >> P.sh1;
            if (b1 >= P.b_lo && b1 < P.b_hi && ha >= P.ha_lo && ha <= P.ha_last) {
                const unsigned bin = b1 - P.b_lo;
                const int slot = 2 * i + n_out;
                atomicAdd(P.hist2 + ((bin << P.lb2) | ((ha >> P.sh2) & sub_mask)), 1u);
#pragma unroll
                for (int w = 0; w < W; ++w) S.stage[w * SLOTS + slot] = key[w];
                S.stage[W * SLOTS + slot] = (uint32_t)val;
                S.stage[(W + 1) * SLOTS + slot] = (uint32_t)(val >> 32);
                S.bin[slot] = (uint16_t)bin;
                S.rank[slot] = (uint16_t)atomicAdd(&S.cnt[bin], 1u);
            }
            ++n_out;
        });
    }
    __syncthreads();
    bin_scatter(S, SLOTS, IW, SLOTS, NB, P.cursor1, P.dst, P.cap, P.slab_cap, P.err);
}

// ------------------------------------------------------------------------------------------------
struct MercyParams {
    const uint32_t *src;
    uint64_t cap;
    const unsigned long long *off2;
    unsigned t_lo, t_hi;
    unsigned *ticket;
    unsigned tab_cap, tab_limit;      // power of two; distinct keys accepted per tile
    unsigned m;
    unsigned long long *cand_out;     // packed ((start_idx + kmer_offset) << 2) | flag  (s1.cpp:764)
    unsigned long long *n_cand;       // keeps counting past cand_cap: the host retries with a buffer that fits
    unsigned long long cand_cap;
    unsigned *err;
};

// compact: the three 4x4 count tables as saturating 2-bit fields (one u32 each) -- enough for min_count <= 3, and a third
// of the shared memory per (k-1)-mer, so three CTAs share an SM instead of one
__host__ __device__ inline size_t mercy_smem_bytes(int W, unsigned cap, bool compact = false) {
    return (size_t)cap * (4 + 4 * (size_t)W + (compact ? 12 : 48) + 2 + 2);
}

// saturating increment of the 2-bit field `x` (0..15) of *w
__device__ __forceinline__ void sat_inc2(uint32_t *w, int x) {
    uint32_t old = *(volatile uint32_t *)w;
    while (((old >> (2 * x)) & 3u) != 3u) {
        const uint32_t prev = atomicCAS(w, old, old + (1u << (2 * x)));
        if (prev == old) return;
        old = prev;
    }
}

// saturating increment of byte `b` of *w (counts only matter up to min_count <= 255)
__device__ __forceinline__ void sat_inc(uint32_t *w, int b) {
    uint32_t old = *(volatile uint32_t *)w;
    while (((old >> (8 * b)) & 0xFFu) != 0xFFu) {
        const uint32_t prev = atomicCAS(w, old, old + (1u << (8 * b)));
        if (prev == old) return;
        old = prev;
    }
}

// COMPACT: tables of 2-bit saturating fields, TW = 3 words per slot (cph | ctn | cht, field index hi * 4 + lo); else bytes, 12 words
template <int W, bool COMPACT>
__global__ void __launch_bounds__(COUNT_THREADS) k_mercy(const MercyParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TW = COMPACT ? 3 : 12;
    const unsigned cap = P.tab_cap, mask = cap - 1, tid = threadIdx.x, lane = tid & 31;
    uint32_t *tag = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *keys = tag + cap;                                   // [W][cap]
    uint32_t *tabs = keys + (size_t)W * cap;                      // [cap][TW]
    uint16_t *gmask = reinterpret_cast<uint16_t *>(tabs + (size_t)TW * cap);   // has_in | has_out << 4 | l_has_out << 8 | r_has_in << 12
    // count of entry x (0..15) of table t (0 cph, 1 ctn, 2 cht) of a slot
    auto count_of = [&](const uint32_t *tb, int t, int x) -> unsigned {
        return COMPACT ? (tb[t] >> (2 * x)) & 3u : reinterpret_cast<const unsigned char *>(tb)[16 * t + x];
    };
    uint16_t *list = gmask + cap;
    __shared__ unsigned s_tile2[2], s_ndist;
    volatile uint32_t *vtag = tag;
    volatile uint32_t *vkeys = keys;
    if (*P.err & ERR_SLAB_OVERFLOW) return;
    for (unsigned i = tid; i < cap; i += COUNT_THREADS) tag[i] = TAG_EMPTY;
    for (unsigned i = tid; i < TW * cap; i += COUNT_THREADS) tabs[i] = 0;
    const unsigned n_tiles = P.t_hi - P.t_lo;

    auto load_key = [&](unsigned long long i, uint32_t (&key)[W], uint32_t &flags) {
#pragma unroll
        for (int w = 0; w < W; ++w) key[w] = P.src[(uint64_t)w * P.cap + i];
        flags = key[W - 1] & 63u;
        key[W - 1] &= S1_FLAG_MASK;
    };
    auto find = [&](const uint32_t (&key)[W], uint32_t hb) -> unsigned {   // first entry in probe order
        const uint32_t fp = (hb >> 4) + 1u;
        unsigned slot = hb & mask;
        while (true) {
            const uint32_t t = vtag[slot];
            if (t == fp) {
                bool eq = true;
#pragma unroll
                for (int w = 0; w < W; ++w) eq = eq && (vkeys[w * cap + slot] == key[w]);
                if (eq) return slot;
            } else if (t == TAG_EMPTY) {
                return 0xFFFFFFFFu;
            }
            slot = (slot + 1) & mask;
        }
    };

    for (unsigned iter = 0;; ++iter) {
        if (tid == 0) { s_tile2[iter & 1] = atomicAdd(P.ticket, 1u); s_ndist = 0; }
        __syncthreads();
        const unsigned s_tile = s_tile2[iter & 1];
        if (s_tile >= n_tiles) break;
        const unsigned t = P.t_lo + s_tile;
        const unsigned long long lo = P.off2[t], hi = P.off2[t + 1];
        if (hi == lo) continue;
        // ---- phase A: claim one slot per distinct S (an insert that skips a locked slot may leave a second entry further
        //      along the probe chain: harmless, every later lookup stops at the first one)
        for (unsigned long long i = lo + tid; i < hi; i += COUNT_THREADS) {
            uint32_t key[W], flags;
            load_key(i, key, flags);
            uint32_t ha, hb;
            edge_hash([&](int w) { return key[w]; }, W, ha, hb);
            const uint32_t fp = (hb >> 4) + 1u;
            unsigned slot = hb & mask;
            while (true) {
                const uint32_t tg = vtag[slot];
                if (tg == fp) {
                    bool eq = true;
#pragma unroll
                    for (int w = 0; w < W; ++w) eq = eq && (vkeys[w * cap + slot] == key[w]);
                    if (eq) break;
                } else if (tg == TAG_EMPTY) {
                    if (*(volatile unsigned *)&s_ndist >= P.tab_limit) break;
                    if (atomicCAS(&tag[slot], TAG_EMPTY, TAG_LOCK) == TAG_EMPTY) {
#pragma unroll
                        for (int w = 0; w < W; ++w) vkeys[w * cap + slot] = key[w];
                        __threadfence_block();
                        vtag[slot] = fp;
                        list[atomicAdd(&s_ndist, 1u)] = (uint16_t)slot;
                        break;
                    }
                    continue;
                }
                slot = (slot + 1) & mask;
            }
        }
        __syncthreads();
        const unsigned nd = s_ndist;
        if (nd >= P.tab_limit) {                                  // uniform; cannot happen with tiles of mean <= cap / 2 items
            if (tid == 0) atomicOr(P.err, (unsigned)ERR_TABLE_FULL);
            for (unsigned i = tid; i < nd; i += COUNT_THREADS) { const unsigned sl = list[i]; tag[sl] = TAG_EMPTY; }
            __syncthreads();
            continue;
        }
        // ---- phase B: the three count tables of every S (s1.cpp:700-704)
        for (unsigned long long i = lo + tid; i < hi; i += COUNT_THREADS) {
            uint32_t key[W], flags;
            load_key(i, key, flags);
            uint32_t ha, hb;
            edge_hash([&](int w) { return key[w]; }, W, ha, hb);
            const unsigned slot = find(key, hb);
            if (slot == 0xFFFFFFFFu) continue;
            const uint32_t v0 = P.src[(uint64_t)W * P.cap + i];
            const int head = flags >> 3, tail = flags & 7, prev = (v0 >> 3) & 7, next = v0 & 7;
            uint32_t *tb = tabs + (size_t)TW * slot;
            if (COMPACT) {
                if (prev < 4 && head < 4) sat_inc2(tb, prev * 4 + head);
                if (tail < 4 && next < 4) sat_inc2(tb + 1, tail * 4 + next);
                if (head < 4 && tail < 4) sat_inc2(tb + 2, head * 4 + tail);
            } else {
                if (prev < 4 && head < 4) { const int x = prev * 4 + head; sat_inc(tb + (x >> 2), x & 3); }
                if (tail < 4 && next < 4) { const int x = tail * 4 + next; sat_inc(tb + 4 + (x >> 2), x & 3); }
                if (head < 4 && tail < 4) { const int x = head * 4 + tail; sat_inc(tb + 8 + (x >> 2), x & 3); }
            }
        }
        __syncthreads();
        // ---- phase C: group masks (s1.cpp:709-737)
        for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
            const unsigned s = list[li];
            const uint32_t *tb = tabs + (size_t)TW * s;
            unsigned has_in = 0, has_out = 0, l_has_out = 0, r_has_in = 0;
            for (int j = 0; j < 4; ++j)
                for (int x = 0; x < 4; ++x) {
                    if (count_of(tb, 0, x * 4 + j) >= P.m) has_in |= 1u << j;               // count_prev_head[x][j]
                    if (count_of(tb, 1, j * 4 + x) >= P.m) has_out |= 1u << j;              // count_tail_next[j][x]
                    if (count_of(tb, 2, j * 4 + x) >= P.m) { l_has_out |= 1u << j; r_has_in |= 1u << x; }
                }
            gmask[s] = (uint16_t)(has_in | (has_out << 4) | (l_has_out << 8) | (r_has_in << 12));
        }
        __syncthreads();
        // ---- phase D: candidates of every item of a short read (s1.cpp:739-826)
        for (unsigned long long i0 = lo; i0 < hi; i0 += COUNT_THREADS) {
            const unsigned long long i = i0 + tid;
            unsigned long long c_val[2];
            int n_c = 0;
            if (i < hi) {
                uint32_t key[W], flags;
                load_key(i, key, flags);
                const uint32_t v0 = P.src[(uint64_t)W * P.cap + i], v1 = P.src[(uint64_t)(W + 1) * P.cap + i];
                const unsigned long long kpos = ((unsigned long long)v1 << 24) | (v0 >> 8);
                if (kpos != S1_NO_EDGE) {
                    uint32_t ha, hb;
                    edge_hash([&](int w) { return key[w]; }, W, ha, hb);
                    const unsigned slot = find(key, hb);
                    if (slot != 0xFFFFFFFFu) {
                        const unsigned gm = gmask[slot];
                        S1GroupMasks g = {(int)(gm & 15u), (int)((gm >> 4) & 15u), (int)((gm >> 8) & 15u), (int)(gm >> 12)};
                        const int head = flags >> 3, tail = flags & 7, strand = (v0 >> 6) & 1;
                        bool solid = false;
                        if (head < 4 && tail < 4) {
                            solid = count_of(tabs + (size_t)TW * slot, 2, head * 4 + tail) >= P.m;
                        }
                        s1_mercy_item(g, solid, head, tail, strand, kpos, [&](uint64_t pos, int flag) {
                            if (n_c < 2) c_val[n_c] = ((unsigned long long)pos << 2) | (unsigned)flag;
                            ++n_c;
                        });
                    }
                }
            }
            // warp-aggregated append (an item pushes at most two candidates)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool have = n_c > q;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, have);
                if (bal) {
                    unsigned long long base = 0;
                    if (lane == (unsigned)(__ffs(bal) - 1)) base = atomicAdd(P.n_cand, (unsigned long long)__popc(bal));
                    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(bal) - 1);
                    const unsigned long long at = base + __popc(bal & ((1u << lane) - 1));
                    if (have && at < P.cand_cap) P.cand_out[at] = c_val[q];
                }
            }
        }
        __syncthreads();
        // ---- clear what this tile used
        for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
            const unsigned s = list[li];
            tag[s] = TAG_EMPTY;
            for (int j = 0; j < TW; ++j) tabs[(size_t)TW * s + j] = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// candidates -> three bit vectors over base positions (s2.cpp:186-199: flag 2 = no-out, 1 = no-in, any = touched)
__global__ void k_mercy_bits(const unsigned long long *__restrict__ cand, unsigned long long n, uint32_t *no_in, uint32_t *no_out,
                             uint32_t *touched) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long c = cand[i], pos = c >> 2;
    const unsigned flag = (unsigned)(c & 3u);
    const uint32_t bit = 1u << (pos & 31);
    if (flag == 2) atomicOr(no_out + (pos >> 5), bit);
    else if (flag == 1) atomicOr(no_in + (pos >> 5), bit);
    atomicOr(touched + (pos >> 5), bit);
}

// one thread per short read: the scan of s2.cpp:201-238 over the read's k-mer offsets; sets is_solid bits
// (bit start_idx[r] + edge offset) and counts them ("Number mercy")
__global__ void k_mercy_reads(const uint64_t *__restrict__ start, uint64_t n_short, int k, const uint32_t *__restrict__ no_in,
                              const uint32_t *__restrict__ no_out, const uint32_t *__restrict__ touched, uint32_t *solid,
                              unsigned long long *num_mercy) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long added = 0;
    if (r < n_short) {
        const uint64_t s0 = start[r];
        const int L = (int)(start[r + 1] - s0);
        const volatile uint32_t *vsolid = solid;                   // written by this kernel: plain loads, not the read-only path
        added = mercy_scan_read(L, k,
                                [&](int v, int i) {
                                    const uint64_t g = s0 + (uint64_t)i;
                                    if (v == 3) return (bool)((vsolid[g >> 5] >> (g & 31)) & 1u);
                                    return bit_at(v == 0 ? no_in : (v == 1 ? no_out : touched), g);
                                },
                                [&](int j) {
                                    const uint64_t e = s0 + (uint64_t)j;
                                    atomicOr(solid + (e >> 5), 1u << (e & 31));
                                });
    }
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xFFFFFFFFu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(num_mercy, added);
}

}  // namespace mgta
