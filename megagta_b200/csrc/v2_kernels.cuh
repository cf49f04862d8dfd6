// v2_kernels.cuh -- edge-centric sm_100a kernels of the CX1 reads -> SdBG path (DESIGN.md sections 3-5).
//
//   k_edge_part   K1+K2  one pass over the 2-bit packed reads: every edge offset yields the canonical
//                 (k+1)-mer [+ its base position]; items are binned in shared memory by the top bits
//                 of a hash and leave the CTA as contiguous runs (one global cursor atomic per bin per
//                 CTA, no per-item global atomics).  Also accumulates the tile histogram.
//                 Replaces s1_lv0_calc_bucket_size + s1_lv1_fill_offset + s1_extract_subtstr_
//                 (reference s1.cpp:177-229, 408-596).
//   k_split       K3a    second partition level: a level-1 bin is cut into its tiles at exact offsets.
//   k_count       K4     one tile (<= a few thousand items, all keys with equal top hash bits) per CTA
//                 pass: shared-memory hash table of the distinct canonical edges, multiplicity
//                 counting, solid marking (s1.cpp:744-760), edge_counting, and the list of solid
//                 edges with multiplicities that feeds stage 2.  Replaces lv2 sort + s1_lv2_output_.
//   k_item_part   K2'    stage-2 items (s2.cpp:586-677) generated from the distinct solid edges instead
//                 of from every read occurrence, binned by key prefix.
//   scans         exact tile offsets from the tile histograms.
//
// Item arrays are SoA: word w of item i lives at buf[w * cap + i].
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "tma_bulk.cuh"

namespace mgta {

constexpr int PART_THREADS = 512;
constexpr int MAX_BINS = 1024;
constexpr int MAX_OWNERS = 16;        // shards of one scan-sharded exchange (GPUs of one NVSwitch domain)
constexpr int COUNT_THREADS = 512;
constexpr uint32_t TAG_EMPTY = 0u, TAG_LOCK = 0xFFFFFFFFu, TAG_DEAD = 0xFFFFFFFEu;
enum { ERR_SLAB_OVERFLOW = 16, ERR_EDGE_LIST_FULL = 32, ERR_OVF_LIST_FULL = 64, ERR_TABLE_FULL = 128 };

// two independent 32-bit hashes of a key: `ha` drives the two partition levels (top bits), `hb`
// the slot and the fingerprint inside a tile's table.
template <class KeyAt>
MGTA_HD void edge_hash(KeyAt key_at, int WE, uint32_t &ha, uint32_t &hb) {
    uint32_t a = 0x9E3779B9u, b = 0x7F4A7C15u;
    for (int w = 0; w < WE; ++w) {
        const uint32_t x = key_at(w);
        a = (a ^ x) * 0x85EBCA6Bu; a ^= a >> 15;
        b = (b + x) * 0xC2B2AE35u; b ^= b >> 13;
    }
    a *= 0x2C1B3C6Du; a ^= a >> 16; a *= 0x297A2D39u; a ^= a >> 15;
    b *= 0x846CA68Bu; b ^= b >> 16; b *= 0x9E3779B1u; b ^= b >> 14;
    ha = a; hb = b;
}

// ------------------------------------------------------------------------------------------------
// Shared-memory binning of one CTA tile of item slots and coalesced scatter of the runs.
struct BinSmem {
    uint32_t *stage;              // [IW][P]  item words by slot
    uint16_t *bin;                // [P]      bin of slot, 0xFFFF = empty slot
    uint16_t *rank;               // [P]      arrival rank of the slot inside its bin
    uint16_t *perm;               // [P]      slot at sorted position j
    uint32_t *cnt;                // [MAX_BINS]
    uint32_t *lbase;              // [MAX_BINS + 1]
    unsigned long long *gbase;    // [MAX_BINS]
    uint32_t *wsum;               // [PART_THREADS / 32]
};

__host__ __device__ inline size_t bin_smem_bytes(int IW, int P) {
    return (size_t)IW * P * 4 + 3 * (size_t)P * 2 + MAX_BINS * 4 + (MAX_BINS + 8) * 4 + MAX_BINS * 8 + 64;
}

__device__ __forceinline__ void bin_smem_carve(BinSmem &S, unsigned char *p, int IW, int P) {
    S.gbase = reinterpret_cast<unsigned long long *>(p); p += MAX_BINS * 8;
    S.stage = reinterpret_cast<uint32_t *>(p); p += (size_t)IW * P * 4;
    S.cnt = reinterpret_cast<uint32_t *>(p); p += MAX_BINS * 4;
    S.lbase = reinterpret_cast<uint32_t *>(p); p += (MAX_BINS + 8) * 4;
    S.wsum = reinterpret_cast<uint32_t *>(p); p += 64;
    S.bin = reinterpret_cast<uint16_t *>(p); p += (size_t)P * 2;
    S.rank = reinterpret_cast<uint16_t *>(p); p += (size_t)P * 2;
    S.perm = reinterpret_cast<uint16_t *>(p);
}

// Precondition: S.cnt / S.bin / S.rank / S.stage filled for slots [0, n_slots), block synchronised.
// cursor[b]: next free absolute item index of bin b in dst; slab_cap != 0: bin b may only use indices
// [b * slab_stride, b * slab_stride + slab_cap) (optimistic fixed-capacity slabs; overflow is flagged, never written;
// slab_stride defaults to slab_cap).
__device__ __forceinline__ void bin_scatter(BinSmem &S, int n_slots, int IW, int P, int NB, unsigned long long *cursor,
                                            uint32_t *dst, uint64_t cap, unsigned long long slab_cap, unsigned *err,
                                            unsigned long long slab_stride = 0) {
    if (slab_stride == 0) slab_stride = slab_cap;                 // slab b holds indices [b * stride, b * stride + slab_cap)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (NB + PART_THREADS - 1) / PART_THREADS;       // <= 4
    unsigned local[4], c_[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = tid * per + j;
        const unsigned c = (j < per && b < NB) ? S.cnt[b] : 0u;
        c_[j] = c; local[j] = sum; sum += c;
    }
    unsigned x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) S.wsum[warp] = x;
    __syncthreads();
    unsigned add = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PART_THREADS / 32; ++w) {
        const unsigned s = S.wsum[w];
        if (w < warp) add += s;
        total += s;
    }
    const unsigned excl = x - sum + add;
    unsigned long long g_[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g_[j] = c_[j] ? atomicAdd(cursor + (tid * per + j), (unsigned long long)c_[j]) : 0ull;
    // gbase[b] = (global index of the bin's first item) - (its sorted position), so that item j of the sorted order goes
    // to gbase[b] + j (modulo 2^64).  A run that does not fit its slab is not written at all (gbase = 2^63, a value no real
    // run can have): the cursor has counted it, the host restarts the pass with slabs that fit.
    bool over = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = tid * per + j;
        if (j < per && b < NB) {
            const unsigned lb = excl + local[j];
            const bool fits = !slab_cap || g_[j] + c_[j] <= (unsigned long long)b * slab_stride + slab_cap;
            if (!fits && c_[j]) over = true;
            S.lbase[b] = lb;
            S.gbase[b] = fits ? g_[j] - lb : 0x8000000000000000ull;
        }
    }
    if (tid == 0) S.lbase[NB] = total;
    __syncthreads();
    // Both loops are chains of dependent shared-memory reads (slot -> bin -> base -> words); four positions per thread are
    // in flight at a time so that the chains overlap (the kernels ran at 50-60 % of the issue rate on short_scoreboard
    // stalls with one chain per thread).
    constexpr int UF = 4;
    for (int i0 = tid; i0 < n_slots; i0 += UF * PART_THREADS) {
        unsigned b_[UF], at_[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            const int i = i0 + u * PART_THREADS;
            b_[u] = i < n_slots ? (unsigned)S.bin[i] : 0xFFFFu;
        }
#pragma unroll
        for (int u = 0; u < UF; ++u) at_[u] = b_[u] != 0xFFFFu ? S.lbase[b_[u]] + S.rank[i0 + u * PART_THREADS] : 0u;
#pragma unroll
        for (int u = 0; u < UF; ++u)
            if (b_[u] != 0xFFFFu) S.perm[at_[u]] = (uint16_t)(i0 + u * PART_THREADS);
    }
    __syncthreads();
    for (unsigned j0 = tid; j0 < total; j0 += UF * PART_THREADS) {
        unsigned i_[UF];
        unsigned long long gb_[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            const unsigned j = j0 + u * PART_THREADS;
            i_[u] = j < total ? (unsigned)S.perm[j] : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < UF; ++u)                               // gbase may have wrapped below zero (first index < sorted position): gb + j is exact
            gb_[u] = i_[u] != 0xFFFFFFFFu ? S.gbase[S.bin[i_[u]]] : 0x8000000000000000ull;
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            if (gb_[u] == 0x8000000000000000ull) continue;
            uint32_t *d = dst + (gb_[u] + (j0 + u * PART_THREADS));
            for (int w = 0; w < IW; ++w, d += cap) *d = S.stage[w * P + i_[u]];
        }
    }
    if (over) atomicOr(err, (unsigned)ERR_SLAB_OVERFLOW);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
struct EdgePartParams {
    const uint32_t *seq;
    const uint64_t *start;
    const uint32_t *lut;            // read lookup table (k_build_read_lut)
    uint64_t n_lut;
    uint64_t n_reads, n_short, total_bases;
    int k;
    int filter, all_solid;          // filter: keep only solid occurrences (stage-2 counting from an is_solid vector)
    const uint32_t *solid;
    int sh1, sh2;                   // level-1 bin = ha >> sh1 (global id), level-2 bin = (ha >> sh2) & (2^lb2 - 1)
    unsigned lb2;
    unsigned b_lo, b_hi;            // level-1 bins of this batch (at most MAX_BINS); tiles are numbered batch-relative
    unsigned long long *cursor1;    // [b_hi - b_lo] absolute next index in dst, slab b starts at b * slab_cap
    unsigned long long slab_cap;
    uint32_t *hist2;                // [(b_hi - b_lo) << lb2]
    uint32_t *dst;
    uint64_t cap;
    unsigned *err;
    // scan-sharded mode (n_owner > 0): only reads [r_begin, r_end) are scanned (CTA tiles start at base g_begin, a multiple
    // of 1024), and the bin of an item is the SHARD that owns its level-1 hash bin: owner d holds bins
    // [owner_lo[d], owner_lo[d + 1]).  Slab d starts at item index d * slab_stride; no tile histogram is taken.
    int n_owner;
    unsigned owner_lo[MAX_OWNERS + 1];
    uint64_t g_begin, g_end;        // base range of this launch (g_begin a multiple of 1024); all modes
    uint64_t r_begin;
    unsigned long long slab_stride;
    // rounds (n_rounds > 1, HBM too small for all items at once): this scan keeps only the items whose level-1 bin lies in
    // slice `round` of its owner's bins, [lo + cnt * round / n_rounds, lo + cnt * (round + 1) / n_rounds)
    unsigned n_rounds, round;
};

// PW = payload words: 0 none, 1 = base position (u32), 2 = base position (lo, hi)
template <int WE, int PW, int TP>
__global__ void __launch_bounds__(PART_THREADS, TP == 4096 && WE + PW <= 2 ? 3 : 1) k_edge_part(const EdgePartParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int IW = WE + PW;
    constexpr int SW_WORDS = TP / 16 + WALK_BACK_WORDS + 12;
    constexpr int NS = 192;                                       // start_idx entries staged per tile (reads of >= ~22-44 bases)
    __shared__ __align__(16) uint32_t sw[SW_WORDS];
    __shared__ uint64_t s_start[NS];
    BinSmem S;
    bin_smem_carve(S, smem_raw, IW, TP);
    const int tid = threadIdx.x;
    const int NB = P.n_owner ? P.n_owner : (int)(P.b_hi - P.b_lo);
    const uint64_t g0 = P.g_begin + (uint64_t)blockIdx.x * TP;    // [g_begin, g_end): the whole read set, a shard's slice of it,
    const uint64_t gend = min(g0 + (uint64_t)TP, P.g_end);         // or one upload chunk (copy / extract pipeline)
    const uint64_t w_lo = (g0 >> 4) >= WALK_BACK_WORDS ? (g0 >> 4) - WALK_BACK_WORDS : 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.seq + w_lo);
        uint4 *dstw = reinterpret_cast<uint4 *>(sw);
        for (int i = tid; i < SW_WORDS / 4; i += PART_THREADS) dstw[i] = __ldg(src + i);
    }
    for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
    // the reads of this tile: two table entries bound them, their start offsets are staged next to the read words
    // (s_start[i] = start[r_lo + i], i <= r_hi + 1 - r_lo), so locating a position never waits on global memory
    uint64_t r_lo, r_hi;
    tile_read_span(P.lut, P.n_lut, g0, gend, r_lo, r_hi);
    const bool st_smem = r_hi - r_lo + 2 <= (uint64_t)NS;          // uniform; tiles of very short reads fall back to global loads
    if (st_smem)
        for (int i = tid; i < (int)(r_hi - r_lo + 2); i += PART_THREADS) s_start[i] = __ldg(P.start + r_lo + i);
    __syncthreads();
    auto start_at = [&](uint64_t rr) -> uint64_t { return st_smem ? s_start[rr - r_lo] : __ldg(P.start + rr); };
    const int k = P.k;
    // Each thread ROLLS over RUN consecutive edge offsets: the (k+1)-mer E and its reverse complement R are cut out of
    // the staged words once and then advanced one base at a time (the device counterpart of the reference's
    // ShiftAppend / ShiftPreappend, megahit_kmer.h:151-174), and the read the position lies in is tracked incrementally.
    constexpr int RUN = TP / PART_THREADS;
    const uint64_t gt = g0 + (uint64_t)tid * RUN;
    const uint32_t q0 = (uint32_t)(gt - 16 * w_lo);
    uint32_t E[WE], R[WE];
    uint64_t r = 0, s_next = 0;
    if (gt < gend) {
        uint64_t lo = r_lo, hi = r_hi;                            // largest r in [r_lo, r_hi] with start[r] <= gt
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) >> 1;
            if (start_at(mid) <= gt) lo = mid; else hi = mid - 1;
        }
        r = lo;
        s_next = start_at(r + 1);
        load_chars<WE>(sw, q0, k + 1, E);
        revcomp<WE>(E, k + 1, R);
    }
    const int cw = k >> 4, csh = (15 - (k & 15)) * 2;             // word / shift of char index k (the last char of the edge)
    const uint32_t tail_mask = head_mask(k + 1 - 16 * (WE - 1));
    const unsigned sub_mask = (1u << P.lb2) - 1u;
#pragma unroll
    for (int j = 0; j < RUN; ++j) {
        const int slot = j * PART_THREADS + tid;                  // any slot numbering works: this one is bank-conflict free
        const uint64_t g = gt + (uint64_t)j;
        unsigned bin = 0xFFFFu;
        if (g < gend) {
            if (j > 0) {
                const uint32_t c = (uint32_t)char_at(sw, q0 + (uint32_t)(j + k));
#pragma unroll
                for (int w = 0; w < WE; ++w) {
                    E[w] = (w + 1 < WE) ? __funnelshift_l(E[w + 1 < WE ? w + 1 : w], E[w], 2) : (E[w] << 2);
                    if (w == cw) E[w] |= c << csh;
                }
#pragma unroll
                for (int w = WE - 1; w >= 0; --w)
                    R[w] = (w > 0) ? __funnelshift_r(R[w], R[w > 0 ? w - 1 : 0], 2) : ((R[0] >> 2) | ((3u - c) << 30));
                R[WE - 1] &= tail_mask;
            }
            while (g >= s_next) { ++r; s_next = start_at(r + 1); }
            bool ok = g + (uint64_t)k + 1 <= s_next;              // the whole (k+1)-mer lies inside read r
            const bool assist = r >= P.n_short;
            if (ok && P.filter) ok = P.all_solid || assist || bit_at(P.solid, g);
            if (ok && P.n_owner) ok = r >= P.r_begin;             // lead-in of the first tile belongs to the previous read
            if (ok) {
                const bool fw = cmp_words<WE>(E, R) <= 0;
                uint32_t key[WE];
#pragma unroll
                for (int w = 0; w < WE; ++w) key[w] = fw ? E[w] : R[w];
                uint32_t ha, hb;
                edge_hash([&](int w) { return key[w]; }, WE, ha, hb);
                const unsigned b1 = ha >> P.sh1;
                bool mine;
                if (P.n_owner) {
                    unsigned d = 0;
#pragma unroll
                    for (int i = 1; i < MAX_OWNERS; ++i) d += (i < P.n_owner && b1 >= P.owner_lo[i]) ? 1u : 0u;
                    mine = true;
                    if (P.n_rounds > 1) {
                        const unsigned lo = P.owner_lo[d], cnt = P.owner_lo[d + 1] - lo, off = b1 - lo;
                        mine = off >= cnt * P.round / P.n_rounds && off < cnt * (P.round + 1) / P.n_rounds;
                    }
                    if (mine) bin = d;
                } else {
                    mine = b1 >= P.b_lo && b1 < P.b_hi;
                    if (mine) {
                        bin = b1 - P.b_lo;
                        atomicAdd(P.hist2 + ((bin << P.lb2) | ((ha >> P.sh2) & sub_mask)), 1u);
                    }
                }
                if (mine) {
#pragma unroll
                    for (int w = 0; w < WE; ++w) S.stage[w * TP + slot] = key[w];
                    if (PW >= 1) S.stage[WE * TP + slot] = assist ? 0xFFFFFFFFu : (uint32_t)g;
                    if (PW >= 2) S.stage[(WE + 1) * TP + slot] = assist ? 0xFFFFFFFFu : (uint32_t)(g >> 32);
                    S.rank[slot] = (uint16_t)atomicAdd(&S.cnt[bin], 1u);
                }
            }
        }
        S.bin[slot] = (uint16_t)bin;
    }
    __syncthreads();
    bin_scatter(S, TP, IW, TP, NB, P.cursor1, P.dst, P.cap, P.slab_cap, P.err, P.n_owner ? P.slab_stride : 0ull);
}

// ------------------------------------------------------------------------------------------------
// exact offsets from a tile histogram: hist[t] for t in [t_lo, t_hi) (zero outside), NT = B1 << lb2
struct ScanParams {
    const uint32_t *hist;
    unsigned NT, lb2, t_lo, t_hi;
    uint32_t *loc;                   // [NT] exclusive scan inside the level-1 bin
    unsigned long long *tot;         // [B1]
    unsigned long long *base;        // [B1 + 1]
    unsigned long long *off2;        // [NT + 1]
    unsigned long long *cursor2;     // [NT]
    unsigned long long *cursor1;     // [B1]  (exact mode: start of each level-1 bin) may be null
    unsigned *chunk_pref;            // [B1 + 1] chunks of T items per level-1 bin
    unsigned T;
    // slab mode (stage 1): level-1 input regions are slabs; exact mode: they are the bins themselves
    unsigned long long slab_cap;
    unsigned b1_lo;                  // first level-1 bin of the batch (slab mode: slab index = b1 - b1_lo)
    unsigned long long *in_start;    // [B1]
};

__global__ void __launch_bounds__(256) k_scan_local(const ScanParams P) {
    __shared__ unsigned s_w[8];
    const unsigned b1 = blockIdx.x, B2 = 1u << P.lb2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned per = (B2 + 255) / 256;
    unsigned sum = 0;
    for (unsigned j = 0; j < per; ++j) {
        const unsigned b2 = tid * per + j, t = (b1 << P.lb2) + b2;
        if (b2 < B2 && t >= P.t_lo && t < P.t_hi) sum += P.hist[t];
    }
    unsigned x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    unsigned add = 0, total = 0;
    for (unsigned w = 0; w < 8; ++w) { if (w < warp) add += s_w[w]; total += s_w[w]; }
    unsigned run = x - sum + add;
    for (unsigned j = 0; j < per; ++j) {
        const unsigned b2 = tid * per + j, t = (b1 << P.lb2) + b2;
        if (b2 < B2) {
            P.loc[t] = run;
            if (t >= P.t_lo && t < P.t_hi) run += P.hist[t];
        }
    }
    if (tid == 0) P.tot[b1] = total;
}

__global__ void __launch_bounds__(1024) k_scan_top(const ScanParams P) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned s_c[32];
    const unsigned B1 = P.NT >> P.lb2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long v = tid < B1 ? P.tot[tid] : 0ull;
    const unsigned c = (unsigned)((v + P.T - 1) / P.T);
    unsigned long long x = v;
    unsigned xc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        const unsigned yc = __shfl_up_sync(0xFFFFFFFFu, xc, o);
        if (lane >= (unsigned)o) { x += y; xc += yc; }
    }
    if (lane == 31) { s_w[warp] = x; s_c[warp] = xc; }
    __syncthreads();
    unsigned long long add = 0, total = 0;
    unsigned addc = 0, totc = 0;
    for (unsigned w = 0; w < 32; ++w) {
        if (w < warp) { add += s_w[w]; addc += s_c[w]; }
        total += s_w[w]; totc += s_c[w];
    }
    if (tid < B1) {
        const unsigned long long excl = x - v + add;
        P.base[tid] = excl;
        P.chunk_pref[tid] = xc - c + addc;
        if (P.cursor1) P.cursor1[tid] = excl;
        P.in_start[tid] = P.slab_cap ? (unsigned long long)(tid >= P.b1_lo ? tid - P.b1_lo : 0u) * P.slab_cap : excl;
    }
    if (tid == 0) { P.base[B1] = total; P.chunk_pref[B1] = totc; P.off2[P.NT] = total; }
}

__global__ void __launch_bounds__(256) k_scan_apply(const ScanParams P) {
    const unsigned t = blockIdx.x * 256 + threadIdx.x;
    if (t < P.NT) {
        const unsigned long long o = P.base[t >> P.lb2] + P.loc[t];
        P.off2[t] = o;
        P.cursor2[t] = o;
    }
}

// ------------------------------------------------------------------------------------------------
struct SplitParams {
    const uint32_t *src;
    uint32_t *dst;
    uint64_t cap_src, cap_dst;
    int IW, WE;
    int mode;                            // 0: hash of the WE key words, 1: prefix bits of key word 0,
                                         // 2: level-1 hash bin of items received from the other shards (see below)
                                         // 3: level-1 PREFIX bin of stage-2 items received from the other shards
    int sh2;                             // tile = x >> sh2 (x = ha or key word 0)
    unsigned lb2;
    const unsigned long long *in_start;  // [B1]
    const unsigned long long *in_count;  // [B1]
    const unsigned *chunk_pref;          // [B1 + 1]
    unsigned B1;
    unsigned long long *cursor2;         // [B1 << lb2] absolute into dst
    unsigned *ticket;
    unsigned T;
    unsigned *err;
    // mode 2: the input regions are the slabs received from the shards; all of them feed ONE set of level-1 bins
    // [b_lo, b_hi) (bin = ha >> sh1; items of other bins are skipped: a later batch takes them), written to fixed-capacity
    // slabs through cursor2[bin] like k_edge_part does, and the tile histogram hist2 is accumulated on the way.
    int sh1;
    unsigned b_lo, b_hi;
    unsigned long long slab_cap;
    uint32_t *hist2;
    uint32_t drop_last;                  // bits of the last hashed key word that do not take part in the hash (mercy items: head/tail flags)
    // mode 3: like mode 2 with bin = key word 0 >> sh1 (global level-1 prefix bin; [b_lo, b_hi) = the batch's bins), only
    // items of lv1 buckets [bkt_lo, bkt_hi), exact cursors (cursor2[bin], slab_cap = 0), no histogram
    unsigned bkt_lo, bkt_hi;
};

// The chunk of a job arrives through the TMA engine: thread 0 takes the ticket, arms the mbarrier with the byte count and
// issues one bulk copy per word plane (cp.async.bulk, 16-byte granules: the chunk is widened to the enclosing 4-item
// boundaries, slots [0, a) and [a + n, ...) of the stage are never binned); the CTA waits on the barrier phase.  T4 = plane
// stride of the stage = T + 4.
__global__ void __launch_bounds__(PART_THREADS, 3) k_split(const SplitParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_job, s_b1, s_n, s_a;
    __shared__ __align__(8) uint64_t s_bar;
    BinSmem S;
    const int T = (int)P.T + 4, IW = P.IW, tid = threadIdx.x;
    bin_smem_carve(S, smem_raw, IW, T);
    const int NB = P.mode >= 2 ? (int)(P.b_hi - P.b_lo) : 1 << P.lb2;
    const unsigned n_jobs = P.chunk_pref[P.B1];
    if (*P.err & ERR_SLAB_OVERFLOW) return;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    unsigned phase = 0;
    // bulk copies need 16-byte aligned planes (every caller carves them so); anything else takes the plain load loop
    const bool bulk_ok = (P.cap_src & 3ull) == 0 && (reinterpret_cast<unsigned long long>(P.src) & 15ull) == 0;
    while (true) {
        if (tid == 0) {
            const unsigned j = atomicAdd(P.ticket, 1u);
            s_job = j;
            if (j < n_jobs) {                                   // level-1 bin of chunk j: last b1 with chunk_pref[b1] <= j
                unsigned lo = 0, hi = P.B1 - 1;
                while (lo < hi) {
                    const unsigned mid = (lo + hi + 1) >> 1;
                    if (P.chunk_pref[mid] <= j) lo = mid; else hi = mid - 1;
                }
                s_b1 = lo;
                const unsigned long long c0 = (unsigned long long)(j - P.chunk_pref[lo]) * P.T;
                const unsigned n = (unsigned)min((unsigned long long)P.T, P.in_count[lo] - c0);
                const unsigned long long s0 = P.in_start[lo] + c0;
                const unsigned a = bulk_ok ? (unsigned)(s0 & 3ull) : 0u, n_ld = (n + a + 3u) & ~3u;
                s_n = n; s_a = a;
                if (bulk_ok) {
                    fence_proxy_async_smem();                   // the previous job's reads of the stage (ordered by the barrier at its end)
                    mbar_arrive_expect_tx(&s_bar, (unsigned)IW * n_ld * 4u);
                    for (int w = 0; w < IW; ++w)
                        bulk_g2s(S.stage + w * T, P.src + (uint64_t)w * P.cap_src + (s0 - a), n_ld * 4u, &s_bar);
                }
            }
        }
        for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
        __syncthreads();
        const unsigned job = s_job;
        if (job >= n_jobs) break;
        const unsigned b1 = s_b1;
        const int a = (int)s_a, n = (int)s_n + a;               // binned slots: [a, n)
        for (int i = tid; i < a; i += PART_THREADS) S.bin[i] = 0xFFFFu;
        if (bulk_ok) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
        } else {
            const unsigned long long s0 = P.in_start[b1] + (unsigned long long)(job - P.chunk_pref[b1]) * P.T;
            for (int w = 0; w < IW; ++w) {
                const uint32_t *sp = P.src + (uint64_t)w * P.cap_src + s0;
                for (int i = tid; i < n; i += PART_THREADS) S.stage[w * T + i] = sp[i];
            }
            __syncthreads();
        }
        for (int i = a + tid; i < n; i += PART_THREADS) {
            uint32_t x;
            if (P.mode == 0 || P.mode == 2) {
                uint32_t hb;
                edge_hash([&](int w) { const uint32_t v = S.stage[w * T + i]; return w == P.WE - 1 ? v & ~P.drop_last : v; }, P.WE, x, hb);
            } else {
                x = S.stage[i];
            }
            if (P.mode >= 2) {
                const unsigned bb = x >> P.sh1;
                const bool in_bkt = P.mode == 2 || ((x >> 16) >= P.bkt_lo && (x >> 16) < P.bkt_hi);
                if (bb >= P.b_lo && bb < P.b_hi && in_bkt) {
                    const unsigned b = bb - P.b_lo;
                    if (P.mode == 2) atomicAdd(P.hist2 + ((b << P.lb2) | ((x >> P.sh2) & ((1u << P.lb2) - 1u))), 1u);
                    S.bin[i] = (uint16_t)b;
                    S.rank[i] = (uint16_t)atomicAdd(&S.cnt[b], 1u);
                } else {
                    S.bin[i] = 0xFFFFu;
                }
                continue;
            }
            const unsigned b2 = (x >> P.sh2) & (unsigned)(NB - 1);
            S.bin[i] = (uint16_t)b2;
            S.rank[i] = (uint16_t)atomicAdd(&S.cnt[b2], 1u);
        }
        __syncthreads();
        if (P.mode >= 2) bin_scatter(S, n, IW, T, NB, P.cursor2, P.dst, P.cap_dst, P.slab_cap, P.err);
        else bin_scatter(S, n, IW, T, NB, P.cursor2 + ((size_t)b1 << P.lb2), P.dst, P.cap_dst, 0ull, P.err);
    }
}

// ------------------------------------------------------------------------------------------------
struct CountParams {
    const uint32_t *src;
    uint64_t cap;
    int PW, k;
    const unsigned long long *off2;
    unsigned t_lo, t_hi;              // tiles of this batch
    const unsigned *tile_list;        // null: all tiles [t_lo, t_hi); else the overflow list of a previous launch
    const unsigned *n_tile_list;
    unsigned *ticket;
    unsigned tab_cap;                 // power of two
    unsigned tab_limit;               // max distinct keys accepted per tile (< tab_cap)
    unsigned m;
    int mark, threshold, has_assist;  // mark: set is_solid bits; threshold: mult = c >= m ? c : assist count (else c)
    int emit;                         // append solid edges to edges_out and their stage-2 items to hist_s2
    uint32_t *solid;
    unsigned long long *edge_counting;
    uint32_t *edges_out;              // rows of WE + 1 words
    unsigned long long *n_edges;
    unsigned long long edges_cap;
    uint32_t *hist_s2;                // histogram of the stage-2 items of the emitted edges by key prefix
    int s2_shift;
    int real_only;                    // count only the real items b S a: the node pass adds the $-items of the tip k-mers
    unsigned *ovf_list, *n_ovf, ovf_cap;
    unsigned *err;
};

struct CountSmem {
    uint32_t *tag, *cnt, *acnt, *keys;
    uint16_t *list;                   // slots claimed for this tile (the only ones that need clearing)
};

__host__ __device__ inline size_t count_smem_bytes(int WE, unsigned cap, int has_assist) {
    return (size_t)cap * 4 * (2 + (has_assist ? 1 : 0) + WE) + (size_t)cap * 2;
}

template <int WE>
__device__ __forceinline__ unsigned table_find(const CountSmem &S, unsigned cap, const uint32_t (&key)[WE], uint32_t hb) {
    const unsigned mask = cap - 1;
    const uint32_t fp = (hb >> 4) + 1u;
    unsigned slot = hb & mask;
    const volatile uint32_t *vtag = S.tag;
    const volatile uint32_t *vkeys = S.keys;
    while (true) {
        const uint32_t t = vtag[slot];
        if (t == fp) {
            bool eq = true;
#pragma unroll
            for (int w = 0; w < WE; ++w) eq = eq && (vkeys[w * cap + slot] == key[w]);
            if (eq) return slot;
        } else if (t == TAG_EMPTY) {
            return 0xFFFFFFFFu;
        }
        slot = (slot + 1) & mask;
    }
}

template <int WE, bool PLUS>
__global__ void __launch_bounds__(COUNT_THREADS) k_count(const CountParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W2 = PLUS ? WE + 1 : WE;
    const unsigned cap = P.tab_cap, mask = cap - 1, tid = threadIdx.x;
    CountSmem S;
    {
        unsigned char *p = smem_raw;
        S.tag = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
        S.cnt = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
        S.keys = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4 * WE;
        S.acnt = S.cnt;
        if (P.has_assist) { S.acnt = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4; }
        S.list = reinterpret_cast<uint16_t *>(p);
    }
    __shared__ unsigned s_tile2[2], s_ndist, s_sawlock, s_nemit, s_ec[256], s_wsum[COUNT_THREADS / 32];
    __shared__ unsigned long long s_ebase;
    for (unsigned i = tid; i < 256; i += COUNT_THREADS) s_ec[i] = 0;
    const unsigned n_tiles = P.tile_list ? *P.n_tile_list : P.t_hi - P.t_lo;
    volatile uint32_t *vtag = S.tag;
    volatile uint32_t *vkeys = S.keys;
    if (*P.err & ERR_SLAB_OVERFLOW) return;                       // the partition is incomplete: the host retries with larger slabs
    for (unsigned i = tid; i < cap; i += COUNT_THREADS) { S.tag[i] = TAG_EMPTY; S.cnt[i] = 0; }
    if (P.has_assist) for (unsigned i = tid; i < cap; i += COUNT_THREADS) S.acnt[i] = 0;

    for (unsigned iter = 0;; ++iter) {
        if (tid == 0) {
            s_tile2[iter & 1] = atomicAdd(P.ticket, 1u);          // double-buffered: laggards of the previous tile still read theirs
            s_ndist = 0; s_sawlock = 0; s_nemit = 0;
        }
        __syncthreads();                                          // the table is clean here (cleared at the end of every tile)
        const unsigned s_tile = s_tile2[iter & 1];
        if (s_tile >= n_tiles) break;
        const unsigned t = P.tile_list ? P.tile_list[s_tile] : P.t_lo + s_tile;
        const unsigned long long lo = P.off2[t], hi = P.off2[t + 1];
        if (hi == lo) continue;                                   // uniform: no divergent barrier
        // ---- phase A: insert / count.  The loads of U items are issued back to back before the first probe: the probe loop
        //      (volatile shared-memory reads, CAS) is a scheduling barrier for the compiler, and one dependent global load
        //      per item left the warps waiting on HBM latency with nothing else in flight.
        constexpr int U = WE <= 2 ? 4 : (WE <= 4 ? 2 : 1);
        for (unsigned long long i0 = lo + tid; i0 < hi; i0 += (unsigned long long)U * COUNT_THREADS) {
            uint32_t keys[U][WE];
            bool assists[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned long long i = i0 + (unsigned long long)u * COUNT_THREADS;
                assists[u] = false;
                if (i < hi) {
#pragma unroll
                    for (int w = 0; w < WE; ++w) keys[u][w] = __ldcs(P.src + (uint64_t)w * P.cap + i);
                    if (P.has_assist && P.PW) assists[u] = P.src[(uint64_t)WE * P.cap + i] == 0xFFFFFFFFu &&
                                                           (P.PW < 2 || P.src[(uint64_t)(WE + 1) * P.cap + i] == 0xFFFFFFFFu);
                } else {
#pragma unroll
                    for (int w = 0; w < WE; ++w) keys[u][w] = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
            if (i0 + (unsigned long long)u * COUNT_THREADS >= hi) break;
            uint32_t key[WE];
#pragma unroll
            for (int w = 0; w < WE; ++w) key[w] = keys[u][w];
            const bool assist = assists[u];
            uint32_t ha, hb;
            edge_hash([&](int w) { return key[w]; }, WE, ha, hb);
            const uint32_t fp = (hb >> 4) + 1u;
            unsigned slot = hb & mask;
            bool placed = false;
            while (!placed) {
                const uint32_t tg = vtag[slot];
                if (tg == fp) {
                    bool eq = true;
#pragma unroll
                    for (int w = 0; w < WE; ++w) eq = eq && (vkeys[w * cap + slot] == key[w]);
                    if (eq) { placed = true; break; }
                } else if (tg == TAG_EMPTY) {
                    if (*(volatile unsigned *)&s_ndist >= P.tab_limit) break;   // table full: tile goes to the overflow list
                    if (atomicCAS(&S.tag[slot], TAG_EMPTY, TAG_LOCK) == TAG_EMPTY) {
#pragma unroll
                        for (int w = 0; w < WE; ++w) vkeys[w * cap + slot] = key[w];
                        __threadfence_block();
                        vtag[slot] = fp;
                        S.list[atomicAdd(&s_ndist, 1u)] = (uint16_t)slot;
                        placed = true;
                        break;
                    }
                    continue;                                     // lost the race: look at the slot again
                } else if (tg == TAG_LOCK) {
                    s_sawlock = 1;                                // may create a duplicate entry further on: fixed below
                }
                slot = (slot + 1) & mask;
            }
            if (placed) {
                atomicAdd(&S.cnt[slot], 1u);
                if (assist) atomicAdd(&S.acnt[slot], 1u);
            }
            }
        }
        __syncthreads();
        const unsigned nd = s_ndist;                              // slots claimed (<= tab_limit + COUNT_THREADS - 1 < cap)
        if (nd >= P.tab_limit) {                                  // uniform
            if (tid == 0) {
                const unsigned o = atomicAdd(P.n_ovf, 1u);
                if (o < P.ovf_cap) P.ovf_list[o] = t; else atomicOr(P.err, (unsigned)(P.tile_list ? ERR_TABLE_FULL : ERR_OVF_LIST_FULL));
                if (P.tile_list) atomicOr(P.err, (unsigned)ERR_TABLE_FULL);
            }
            for (unsigned i = tid; i < nd; i += COUNT_THREADS) { const unsigned sl = S.list[i]; S.tag[sl] = TAG_EMPTY; S.cnt[sl] = 0; S.acnt[sl] = 0; }
            __syncthreads();
            continue;
        }
        // ---- duplicates left behind by skipped locked slots: fold into the first entry in probe order
        if (s_sawlock) {                                          // uniform
            for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
                const unsigned s = S.list[li];
                uint32_t key[WE];
#pragma unroll
                for (int w = 0; w < WE; ++w) key[w] = S.keys[w * cap + s];
                uint32_t ha, hb;
                edge_hash([&](int w) { return key[w]; }, WE, ha, hb);
                const unsigned first = table_find<WE>(S, cap, key, hb);
                if (first != s && first != 0xFFFFFFFFu) {
                    atomicAdd(&S.cnt[first], S.cnt[s]);
                    if (P.has_assist) atomicAdd(&S.acnt[first], S.acnt[s]);
                    vtag[s] = TAG_DEAD;
                }
            }
            __syncthreads();
        }
        // ---- phase B: solid marking of every occurrence
        if (P.mark) {
            for (unsigned long long i = lo + tid; i < hi; i += COUNT_THREADS) {
                const uint32_t p0 = P.src[(uint64_t)WE * P.cap + i];
                const uint32_t p1 = P.PW >= 2 ? P.src[(uint64_t)(WE + 1) * P.cap + i] : 0u;
                if (p0 == 0xFFFFFFFFu && (P.PW < 2 || p1 == 0xFFFFFFFFu)) continue;       // assist read: never marked (s1.cpp:757)
                uint32_t key[WE];
#pragma unroll
                for (int w = 0; w < WE; ++w) key[w] = P.src[(uint64_t)w * P.cap + i];
                uint32_t ha, hb;
                edge_hash([&](int w) { return key[w]; }, WE, ha, hb);
                const unsigned slot = table_find<WE>(S, cap, key, hb);
                if (slot != 0xFFFFFFFFu && S.cnt[slot] >= P.m) {
                    const unsigned long long g = ((unsigned long long)p1 << 32) | p0;
                    atomicOr(P.solid + (g >> 5), 1u << (g & 31));
                }
            }
        }
        // ---- phase C: per distinct edge: edge_counting (s1.cpp:744-746) and the solid edge list.  Two walks over this
        //      thread's claimed slots: count the rows it will write, block scan + one global reservation, write them.
        unsigned my_rows = 0;
        for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
            const unsigned s = S.list[li];
            if (S.tag[s] == TAG_DEAD) continue;
            const unsigned c = S.cnt[s];
            if (P.edge_counting) {
                if (c < 256) atomicAdd(&s_ec[c], 1u);
                else atomicAdd(P.edge_counting + (c < 65535u ? c : 65535u), 1ull);
            }
            const unsigned mult = P.threshold ? (c >= P.m ? c : (P.has_assist ? S.acnt[s] : 0u)) : c;
            if (mult && P.emit) ++my_rows;
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
                const unsigned long long b = atomicAdd(P.n_edges, (unsigned long long)ne);
                s_ebase = b;
                if (b + ne > P.edges_cap) atomicOr(P.err, (unsigned)ERR_EDGE_LIST_FULL);
            }
            __syncthreads();
            const unsigned long long eb = s_ebase;
            if (eb + ne <= P.edges_cap && my_rows) {
                unsigned long long r = eb + row0;
                for (unsigned li = tid; li < nd; li += COUNT_THREADS) {
                    const unsigned s = S.list[li];
                    if (S.tag[s] == TAG_DEAD) continue;
                    const unsigned c = S.cnt[s];
                    const unsigned mult = P.threshold ? (c >= P.m ? c : (P.has_assist ? S.acnt[s] : 0u)) : c;
                    if (!mult) continue;
                    uint32_t key[WE];
#pragma unroll
                    for (int w = 0; w < WE; ++w) key[w] = S.keys[w * cap + s];
                    uint32_t *row = P.edges_out + r * (WE + 1);
                    ++r;
#pragma unroll
                    for (int w = 0; w < WE; ++w) row[w] = key[w];
                    row[WE] = mult;
                    s2_items_of_edge<W2, WE>(key, P.k, [&](const uint32_t(&y)[W2]) { atomicAdd(P.hist_s2 + (y[0] >> P.s2_shift), 1u); },
                                             !P.real_only);
                }
            }
            __syncthreads();
        }
        for (unsigned i = tid; i < nd; i += COUNT_THREADS) { const unsigned sl = S.list[i]; S.tag[sl] = TAG_EMPTY; S.cnt[sl] = 0; S.acnt[sl] = 0; }
        __syncthreads();
    }
    __syncthreads();
    if (P.edge_counting)
        for (unsigned i = tid; i < 256; i += COUNT_THREADS)
            if (s_ec[i]) atomicAdd(P.edge_counting + i, (unsigned long long)s_ec[i]);
}

// ------------------------------------------------------------------------------------------------
struct ItemPartParams {
    const uint32_t *edges;            // rows of WE + 1 words
    unsigned long long n_edges;
    int k, sh1;                       // level-1 bin = key word 0 >> sh1
    unsigned bkt_lo, bkt_hi;          // lv1 buckets (top 16 key bits) of this batch
    unsigned long long *cursor1;      // [NB] exact absolute starts of the batch's level-1 bins
    unsigned NB, b1_lo;               // level-1 bins of the batch: global bins [b1_lo, b1_lo + NB)
    uint32_t *dst;
    uint64_t cap;
    unsigned *err;
    // sharded (n_owner > 0): bin = the shard whose lv1-bucket range [bnd[d], bnd[d + 1]) holds the item (no bucket filter);
    // slab d starts at item index d * slab_stride and holds slab_cap items
    int n_owner;
    unsigned bnd[MAX_OWNERS + 1];
    unsigned long long slab_cap, slab_stride;
};

// item slots per CTA: 512 edges x 6 items; 2048 edges x 2 real items while the staged items leave two CTAs per SM (IW <= 4),
// else 1536 x 2
__host__ __device__ constexpr int item_slots(int IW, int PER) { return PER == 2 && IW <= 4 ? 4096 : 3072; }

// PER = 6: all stage-2 items of an edge ($-items unconditionally; the group logic drops the covered ones);
// PER = 2: the real items only -- the $-items of the tip k-mers come from the node pass (k_row_part)
template <int WE, bool PLUS, int PER>
__global__ void __launch_bounds__(PART_THREADS) k_item_part(const ItemPartParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W2 = PLUS ? WE + 1 : WE;
    constexpr int IW = W2 + 1, SLOTS = item_slots(IW, PER), EDGES = SLOTS / PER;
    BinSmem S;
    bin_smem_carve(S, smem_raw, IW, SLOTS);
    const int tid = threadIdx.x;
    const int NB = P.n_owner ? P.n_owner : (int)P.NB;
    for (int i = tid; i < NB; i += PART_THREADS) S.cnt[i] = 0;
    for (int i = tid; i < SLOTS; i += PART_THREADS) S.bin[i] = 0xFFFFu;
    __syncthreads();
    const unsigned long long e0 = (unsigned long long)blockIdx.x * EDGES;
    for (int el = tid; el < EDGES; el += PART_THREADS) {
        const unsigned long long e = e0 + el;
        if (e >= P.n_edges) break;
        const uint32_t *row = P.edges + e * (WE + 1);
        uint32_t key[WE];
#pragma unroll
        for (int w = 0; w < WE; ++w) key[w] = __ldg(row + w);
        const uint32_t mult = __ldg(row + WE);
        int j = 0;
        s2_items_of_edge<W2, WE>(key, P.k, [&](const uint32_t(&y)[W2]) {
            const unsigned bkt = y[0] >> 16;
            if (P.n_owner || (bkt >= P.bkt_lo && bkt < P.bkt_hi)) {
                const int slot = el * PER + j;
                unsigned b;
                if (P.n_owner) {
                    b = 0;
#pragma unroll
                    for (int i = 1; i < MAX_OWNERS; ++i) b += (i < P.n_owner && bkt >= P.bnd[i]) ? 1u : 0u;
                } else {
                    b = (y[0] >> P.sh1) - P.b1_lo;
                }
#pragma unroll
                for (int w = 0; w < W2; ++w) S.stage[w * SLOTS + slot] = y[w];
                S.stage[W2 * SLOTS + slot] = mult;
                S.bin[slot] = (uint16_t)b;
                S.rank[slot] = (uint16_t)atomicAdd(&S.cnt[b], 1u);
            }
            ++j;
        }, PER == 6);
    }
    __syncthreads();
    bin_scatter(S, SLOTS, IW, SLOTS, NB, P.cursor1, P.dst, P.cap, P.n_owner ? P.slab_cap : 0ull, P.err, P.n_owner ? P.slab_stride : 0ull);
}

// leaf-start flags of the non-empty tiles of a batch (the on-chip sort's windows never straddle a leaf)
__global__ void k_flags_tiles(const unsigned long long *__restrict__ off2, unsigned t_lo, unsigned t_hi, uint32_t *flags) {
    const unsigned t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < t_hi) {
        const unsigned long long s = off2[t];
        if (off2[t + 1] > s) atomicOr(flags + (s >> 5), 1u << (s & 31));
    }
}

// edges * 1: number of edge offsets (k+1)-mer positions over all reads
__global__ void k_init_slab_cursors(unsigned long long *cursor, unsigned n, unsigned long long slab_cap) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cursor[i] = (unsigned long long)i * slab_cap;
}

__global__ void k_count_positions(const uint64_t *__restrict__ start, uint64_t r_begin, uint64_t n_reads, int k, unsigned long long *out) {
    const uint64_t r = r_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (r < n_reads) {
        const int64_t L = (int64_t)(start[r + 1] - start[r]);
        if (L >= k + 1) v = (unsigned long long)(L - k);
    }
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}


// edge offsets per equal slice of the reads: slice d = reads [n_reads * d / parts, n_reads * (d + 1) / parts)
__global__ void k_count_positions_parts(const uint64_t *__restrict__ start, uint64_t n_reads, int k, unsigned parts, unsigned long long *out) {
    const uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = r0 < n_reads;
    const uint64_t r = valid ? r0 : n_reads - 1;
    unsigned long long v = 0;
    if (valid) {
        const int64_t L = (int64_t)(start[r + 1] - start[r]);
        if (L >= k + 1) v = (unsigned long long)(L - k);
    }
    unsigned d = (unsigned)(((unsigned __int128)r * parts) / n_reads);       // the d with n_reads*d/parts <= r < n_reads*(d+1)/parts
    while (d + 1 < parts && (uint64_t)(((unsigned __int128)n_reads * (d + 1)) / parts) <= r) ++d;
    while (d > 0 && (uint64_t)(((unsigned __int128)n_reads * d) / parts) > r) --d;
    int same = 0;
    __match_all_sync(0xFFFFFFFFu, d, &same);
    if (same) {                                                              // the usual case: one slice per warp
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + d, v);
    } else if (v) {
        atomicAdd(out + d, v);
    }
}

}  // namespace mgta
