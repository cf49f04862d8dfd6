// node_kernels.cuh -- the node pass of stage 2: which k-mers are TIPS of the solid graph.
//
// output_() (reference s2.cpp:801-818) drops a $-item whenever a real edge covers it; the only $-items that become
// records are those of tip k-mers.  Instead of generating four $-items per distinct edge and letting the on-chip sort
// throw 97 % of them away, this pass accumulates, per canonical k-mer, the summed multiplicities of the solid edges
// leaving (out) and entering (in) it and emits the two $-items of every k-mer with exactly one of them zero.
// No order is needed, only grouping by k-mer, so the ops take the route of the stage-1 items:
//
//   k_node_part   2 ops per distinct solid edge (node_ops_of_edge, cx1_items.cuh): key = canonical k-mer, payload =
//                 min(mult, 65535) << 1 | dir; binned in shared memory by the top bits of the key's hash into the
//                 level-1 slabs (or, scan-sharded, by owner shard), like k_edge_part
//   k_split       (v2_kernels.cuh, mode 0) level-1 bin -> its hash tiles at exact offsets
//   k_node_count  per tile: shared-memory hash table of the distinct k-mers with the two weights; every tip k-mer
//                 appends its two stage-2 items {key words, weight} to the tip list and counts them in the stage-2
//                 key-prefix histogram
//   k_row_part    stage-2 level-1 prefix partition of the tip list (the edges' real items go through k_item_part)
#pragma once
#include "v2_kernels.cuh"

namespace mgta {

struct NodePartParams {
    const uint32_t *edges;            // rows of WE + 1 words
    unsigned long long n_edges;
    int k;
    int sh1, sh2;                     // level-1 bin = ha >> sh1, level-2 bin = (ha >> sh2) & (2^lb2 - 1)
    unsigned lb2;
    unsigned b_lo, b_hi;              // level-1 bins of this batch
    unsigned long long *cursor1;      // [b_hi - b_lo] absolute next index in dst, slab b starts at b * slab_cap
    unsigned long long slab_cap;
    uint32_t *hist2;                  // [(b_hi - b_lo) << lb2]
    uint32_t *dst;
    uint64_t cap;
    unsigned *err;
    // sharded (n_owner > 0): the bin of an op is the SHARD that owns its level-1 hash bin (owner d holds bins
    // [owner_lo[d], owner_lo[d + 1])); slab d starts at item index d * slab_stride, holds slab_cap ops; no tile histogram
    int n_owner;
    unsigned owner_lo[MAX_OWNERS + 1];
    unsigned long long slab_stride;
};

// edges per CTA (2 op slots each): with up to 1024 bins per CTA a small tile leaves runs of two ops per bin (one cursor atomic
// and one partial sector per op); 2048 edges while the staged ops fit two CTAs per SM
__host__ __device__ constexpr int node_edges(int KW) { return KW + 1 <= 4 ? 2048 : 1024; }

// KW = kmer_words(k); EPLUS: the edge has one word more than the k-mer (k % 16 == 0)
template <int KW, bool EPLUS>
__global__ void __launch_bounds__(PART_THREADS) k_node_part(const NodePartParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NODE_EDGES = node_edges(KW);
    constexpr int WE = KW + (EPLUS ? 1 : 0), IW = KW + 1, SLOTS = NODE_EDGES * 2;
    BinSmem S;
    bin_smem_carve(S, smem_raw, IW, SLOTS);
    const int tid = threadIdx.x;
    const int NB = P.n_owner ? P.n_owner : (int)(P.b_hi - P.b_lo);
    for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
    for (int i = tid; i < SLOTS; i += PART_THREADS) S.bin[i] = 0xFFFFu;
    __syncthreads();
    const unsigned long long e0 = (unsigned long long)blockIdx.x * NODE_EDGES;
    const unsigned sub_mask = (1u << P.lb2) - 1u;
    for (int el = tid; el < NODE_EDGES; el += PART_THREADS) {
        const unsigned long long e = e0 + el;
        if (e >= P.n_edges) break;
        const uint32_t *row = P.edges + e * (WE + 1);
        uint32_t key[WE];
#pragma unroll
        for (int w = 0; w < WE; ++w) key[w] = __ldg(row + w);
        const uint32_t mult = __ldg(row + WE);
        const uint32_t wgt = mult < 65535u ? mult : 65535u;
        int j = 0;
        node_ops_of_edge<WE>(key, P.k, [&](const uint32_t(&c)[WE], int dir) {
            uint32_t ha, hb;
            edge_hash([&](int w) { return c[w]; }, KW, ha, hb);
            const unsigned b1 = ha >> P.sh1;
            if (P.n_owner || (b1 >= P.b_lo && b1 < P.b_hi)) {
                const int slot = el * 2 + j;
                unsigned bin;
                if (P.n_owner) {
                    bin = 0;
#pragma unroll
                    for (int i = 1; i < MAX_OWNERS; ++i) bin += (i < P.n_owner && b1 >= P.owner_lo[i]) ? 1u : 0u;
                } else {
                    bin = b1 - P.b_lo;
                    atomicAdd(P.hist2 + ((bin << P.lb2) | ((ha >> P.sh2) & sub_mask)), 1u);
                }
#pragma unroll
                for (int w = 0; w < KW; ++w) S.stage[w * SLOTS + slot] = c[w];
                S.stage[KW * SLOTS + slot] = (wgt << 1) | (uint32_t)dir;
                S.bin[slot] = (uint16_t)bin;
                S.rank[slot] = (uint16_t)atomicAdd(&S.cnt[bin], 1u);
            }
            ++j;
        });
    }
    __syncthreads();
    bin_scatter(S, SLOTS, IW, SLOTS, NB, P.cursor1, P.dst, P.cap, P.slab_cap, P.err, P.n_owner ? P.slab_stride : 0ull);
}

// ------------------------------------------------------------------------------------------------
struct NodeCountParams {
    const uint32_t *src;              // ops by tile: KW key words + 1 payload word, SoA
    uint64_t cap;
    int k;
    const unsigned long long *off2;
    unsigned t_lo, t_hi;
    const unsigned *tile_list;        // null: all tiles [t_lo, t_hi); else the overflow list of the first launch
    const unsigned *n_tile_list;
    unsigned *ticket;
    unsigned tab_cap, tab_limit;
    uint32_t *tips_out;               // rows of W2 + 1 words: stage-2 key, weight
    unsigned long long *n_tips;       // keeps counting past tips_cap (the host then reruns with room)
    unsigned long long tips_cap;
    uint32_t *hist_s2;
    int s2_shift;
    unsigned *ovf_list, *n_ovf, ovf_cap;
    unsigned *err;
};

// W2 = key_words_s2(k) (KW or KW + 1)
template <int KW, bool PLUS>
__global__ void __launch_bounds__(COUNT_THREADS) k_node_count(const NodeCountParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W2 = PLUS ? KW + 1 : KW;
    const unsigned cap = P.tab_cap, mask = cap - 1, tid = threadIdx.x;
    CountSmem S;                      // cnt = out weight, acnt = in weight
    {
        unsigned char *p = smem_raw;
        S.tag = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
        S.cnt = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
        S.keys = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4 * KW;
        S.acnt = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
        S.list = reinterpret_cast<uint16_t *>(p);
    }
    __shared__ unsigned s_tile2[2], s_ndist, s_sawlock, s_nemit, s_wsum[COUNT_THREADS / 32];
    __shared__ unsigned long long s_ebase;
    const unsigned n_tiles = P.tile_list ? *P.n_tile_list : P.t_hi - P.t_lo;
    volatile uint32_t *vtag = S.tag;
    volatile uint32_t *vkeys = S.keys;
    if (*P.err & ERR_SLAB_OVERFLOW) return;
    for (unsigned i = tid; i < cap; i += COUNT_THREADS) { S.tag[i] = TAG_EMPTY; S.cnt[i] = 0; S.acnt[i] = 0; }

    for (unsigned iter = 0;; ++iter) {
        if (tid == 0) {
            s_tile2[iter & 1] = atomicAdd(P.ticket, 1u);
            s_ndist = 0; s_sawlock = 0; s_nemit = 0;
        }
        __syncthreads();
        const unsigned s_tile = s_tile2[iter & 1];
        if (s_tile >= n_tiles) break;
        const unsigned t = P.tile_list ? P.tile_list[s_tile] : P.t_lo + s_tile;
        const unsigned long long lo = P.off2[t], hi = P.off2[t + 1];
        if (hi == lo) continue;
        // ---- phase A: insert / accumulate (the lock-free table of k_count)
        constexpr int U = KW <= 2 ? 4 : (KW <= 4 ? 2 : 1);       // loads of U ops in flight before the first probe (see k_count)
        for (unsigned long long i0 = lo + tid; i0 < hi; i0 += (unsigned long long)U * COUNT_THREADS) {
            uint32_t keys[U][KW], pays[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned long long i = i0 + (unsigned long long)u * COUNT_THREADS;
                const bool in = i < hi;
#pragma unroll
                for (int w = 0; w < KW; ++w) keys[u][w] = in ? __ldcs(P.src + (uint64_t)w * P.cap + i) : 0u;
                pays[u] = in ? __ldcs(P.src + (uint64_t)KW * P.cap + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
            if (i0 + (unsigned long long)u * COUNT_THREADS >= hi) break;
            uint32_t key[KW];
#pragma unroll
            for (int w = 0; w < KW; ++w) key[w] = keys[u][w];
            const uint32_t pay = pays[u];
            uint32_t ha, hb;
            edge_hash([&](int w) { return key[w]; }, KW, ha, hb);
            const uint32_t fp = (hb >> 4) + 1u;
            unsigned slot = hb & mask;
            bool placed = false;
            while (!placed) {
                const uint32_t tg = vtag[slot];
                if (tg == fp) {
                    bool eq = true;
#pragma unroll
                    for (int w = 0; w < KW; ++w) eq = eq && (vkeys[w * cap + slot] == key[w]);
                    if (eq) { placed = true; break; }
                } else if (tg == TAG_EMPTY) {
                    if (*(volatile unsigned *)&s_ndist >= P.tab_limit) break;
                    if (atomicCAS(&S.tag[slot], TAG_EMPTY, TAG_LOCK) == TAG_EMPTY) {
#pragma unroll
                        for (int w = 0; w < KW; ++w) vkeys[w * cap + slot] = key[w];
                        __threadfence_block();
                        vtag[slot] = fp;
                        S.list[atomicAdd(&s_ndist, 1u)] = (uint16_t)slot;
                        placed = true;
                        break;
                    }
                    continue;
                } else if (tg == TAG_LOCK) {
                    s_sawlock = 1;
                }
                slot = (slot + 1) & mask;
            }
            if (placed) atomicAdd((pay & 1u) ? &S.acnt[slot] : &S.cnt[slot], pay >> 1);
            }
        }
        __syncthreads();
        const unsigned nd = s_ndist;
        if (nd >= P.tab_limit) {                                  // uniform: the tile goes to the overflow pass
            if (tid == 0) {
                const unsigned o = atomicAdd(P.n_ovf, 1u);
                if (o < P.ovf_cap) P.ovf_list[o] = t; else atomicOr(P.err, (unsigned)(P.tile_list ? ERR_TABLE_FULL : ERR_OVF_LIST_FULL));
                if (P.tile_list) atomicOr(P.err, (unsigned)ERR_TABLE_FULL);
            }
            for (unsigned i = tid; i < nd; i += COUNT_THREADS) { const unsigned sl = S.list[i]; S.tag[sl] = TAG_EMPTY; S.cnt[sl] = 0; S.acnt[sl] = 0; }
            __syncthreads();
            continue;
        }
        if (s_sawlock) {                                          // uniform: fold duplicates into the first entry in probe order
            for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
                const unsigned s = S.list[li];
                uint32_t key[KW];
#pragma unroll
                for (int w = 0; w < KW; ++w) key[w] = S.keys[w * cap + s];
                uint32_t ha, hb;
                edge_hash([&](int w) { return key[w]; }, KW, ha, hb);
                const unsigned first = table_find<KW>(S, cap, key, hb);
                if (first != s && first != 0xFFFFFFFFu) {
                    atomicAdd(&S.cnt[first], S.cnt[s]);
                    atomicAdd(&S.acnt[first], S.acnt[s]);
                    vtag[s] = TAG_DEAD;
                }
            }
            __syncthreads();
        }
        // ---- tips: k-mers with exactly one of (out, in) zero; two rows each
        unsigned my_rows = 0;
        for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
            const unsigned s = S.list[li];
            if (S.tag[s] == TAG_DEAD) continue;
            const unsigned o = S.cnt[s], in = S.acnt[s];
            if ((o == 0) == (in == 0)) continue;
            uint32_t C[W2], RC[W2];
#pragma unroll
            for (int w = 0; w < W2; ++w) C[w] = w < KW ? S.keys[(w < KW ? w : 0) * cap + s] : 0u;
            revcomp<W2>(C, P.k, RC);
            if (cmp_words<W2>(C, RC) == 0) continue;              // palindromic k-mer: in == out by symmetry, never a tip
            my_rows += 2;
        }
        unsigned row0;
        {
            const unsigned lane = tid & 31, warp = tid >> 5;
            unsigned x = my_rows;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= (unsigned)o) x += y;
            }
            if (lane == 31) s_wsum[warp] = x;
            __syncthreads();
            unsigned add = 0, total = 0;
#pragma unroll
            for (int w = 0; w < COUNT_THREADS / 32; ++w) {
                const unsigned v = s_wsum[w];
                if ((unsigned)w < warp) add += v;
                total += v;
            }
            row0 = x - my_rows + add;
            if (tid == 0) s_nemit = total;
        }
        __syncthreads();
        const unsigned ne = s_nemit;
        if (ne) {                                                 // uniform
            if (tid == 0) {
                const unsigned long long b = atomicAdd(P.n_tips, (unsigned long long)ne);
                s_ebase = b;
                if (b + ne > P.tips_cap) atomicOr(P.err, (unsigned)ERR_EDGE_LIST_FULL);
            }
            __syncthreads();
            const unsigned long long eb = s_ebase;
            if (eb + ne <= P.tips_cap && my_rows) {
                unsigned long long r = eb + row0;
                for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
                    const unsigned s = S.list[li];
                    if (S.tag[s] == TAG_DEAD) continue;
                    const unsigned o = S.cnt[s], in = S.acnt[s];
                    if ((o == 0) == (in == 0)) continue;
                    uint32_t C[W2], RC[W2];
#pragma unroll
                    for (int w = 0; w < W2; ++w) C[w] = w < KW ? S.keys[(w < KW ? w : 0) * cap + s] : 0u;
                    revcomp<W2>(C, P.k, RC);
                    if (cmp_words<W2>(C, RC) == 0) continue;
                    const uint32_t wgt = (o + in) < 65535u ? (o + in) : 65535u;
                    s2_tip_items<W2>(C, RC, in == 0, P.k, [&](const uint32_t(&y)[W2]) {
                        uint32_t *row = P.tips_out + r * (W2 + 1);
                        ++r;
#pragma unroll
                        for (int w = 0; w < W2; ++w) row[w] = y[w];
                        row[W2] = wgt;
                        atomicAdd(P.hist_s2 + (y[0] >> P.s2_shift), 1u);
                    });
                }
            }
            __syncthreads();
        }
        for (unsigned i = tid; i < nd; i += COUNT_THREADS) { const unsigned sl = S.list[i]; S.tag[sl] = TAG_EMPTY; S.cnt[sl] = 0; S.acnt[sl] = 0; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// level-1 prefix partition of ready-made stage-2 items held as AoS rows of IW words (the tip list)
struct RowPartParams {
    const uint32_t *rows;
    unsigned long long n_rows;
    int IW, sh1;
    unsigned bkt_lo, bkt_hi;
    unsigned long long *cursor1;
    unsigned NB, b1_lo;
    uint32_t *dst;
    uint64_t cap;
    unsigned *err;
    // sharded (n_owner > 0): bin = the shard whose lv1-bucket range [bnd[d], bnd[d + 1]) holds the item; slab d starts at
    // item index d * slab_stride and holds slab_cap items
    int n_owner;
    unsigned bnd[MAX_OWNERS + 1];
    unsigned long long slab_cap, slab_stride;
};

constexpr int ROW_SLOTS = 2048;

__global__ void __launch_bounds__(PART_THREADS) k_row_part(const RowPartParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BinSmem S;
    const int IW = P.IW;
    bin_smem_carve(S, smem_raw, IW, ROW_SLOTS);
    const int tid = threadIdx.x;
    const int NB = P.n_owner ? P.n_owner : (int)P.NB;
    for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
    __syncthreads();
    const unsigned long long r0 = (unsigned long long)blockIdx.x * ROW_SLOTS;
    for (int sl = tid; sl < ROW_SLOTS; sl += PART_THREADS) {
        const unsigned long long r = r0 + sl;
        unsigned bin = 0xFFFFu;
        if (r < P.n_rows) {
            const uint32_t *row = P.rows + r * IW;
            const uint32_t y0 = __ldg(row);
            const unsigned bkt = y0 >> 16;
            if (P.n_owner || (bkt >= P.bkt_lo && bkt < P.bkt_hi)) {
                if (P.n_owner) {
                    bin = 0;
#pragma unroll
                    for (int i = 1; i < MAX_OWNERS; ++i) bin += (i < P.n_owner && bkt >= P.bnd[i]) ? 1u : 0u;
                } else {
                    bin = (y0 >> P.sh1) - P.b1_lo;
                }
                for (int w = 0; w < IW; ++w) S.stage[w * ROW_SLOTS + sl] = __ldg(row + w);
                S.rank[sl] = (uint16_t)atomicAdd(&S.cnt[bin], 1u);
            }
        }
        S.bin[sl] = (uint16_t)bin;
    }
    __syncthreads();
    bin_scatter(S, ROW_SLOTS, IW, ROW_SLOTS, NB, P.cursor1, P.dst, P.cap, P.n_owner ? P.slab_cap : 0ull, P.err, P.n_owner ? P.slab_stride : 0ull);
}

}  // namespace mgta
