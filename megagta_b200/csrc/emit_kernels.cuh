// emit_kernels.cuh -- K3b + K5 of stage 2: per window of <= CAPI weighted items (prefix tiles never straddle a
// window) sort by the whole key in shared memory, then turn every group of equal (k-1)-mer S into SdBG records.
//
// Replaces lv2_cpu_radix_sort_st (reference lv2_cpu_sort.h:113-151), output_() (s2.cpp:742-835) and
// SdbgWriter::write (sdbg_multi_io.h:83-112) for the items k_item_part generated from the distinct solid edges.
//
// Sort = LSD radix over the whole key below the window's shared prefix, 8-bit digits, on a u16 permutation in shared
// memory: one __match_any_sync per item and pass (digit, warp-local rank and item stay in registers between the
// counting and the scatter half), per-warp digit counters in (digit, warp) order so that one 16-byte load per
// thread scans all 4096 of them.  Pass 0 is the 6-bit in-group order (a slot, a != $, b), the rest are S bytes.
// Emission: one thread per group walks a per-position code word (a, b, multiplicity, run/group bits) twice
// (masks, then records -- the two passes of s2.cpp:766-830) and leaves a record descriptor at each run start;
// a block scan over descriptor sizes + the decoupled look-back over windows gives every record its byte offset.
#pragma once
#include "kernels.cuh"

namespace mgta {

__host__ __device__ inline size_t sort_emit_smem_bytes(int IW, unsigned capi) {
    return (size_t)IW * capi * 4 + (size_t)capi * 4 + 2 * (size_t)capi * 2 + CHUNK_WARPS * 256 * 2 + 2 * (capi / 32 + 2) * 4 +
           (CHUNK_WARPS + 2) * 4 + 64;
}

template <int W>
__device__ __forceinline__ int key_cmp(const uint32_t *keys, unsigned capi, unsigned x, unsigned y) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t a = keys[w * capi + x], b = keys[w * capi + y];
        if (a != b) return a < b ? -1 : 1;
    }
    return 0;
}

// Bucket + rank sort of one window: the items are distinct-edge items, so below the `kb` key bits the window shares they
// are close to uniformly spread.  One counting pass over the next D key bits (shared-memory atomics give the arrival
// rank inside a bin), an exclusive scan of the <= 4096 bin counters, and a comparison rank inside each bin (bins hold
// the few items of one (k-1)-mer group plus chance collisions) replace the 1 + ceil((2(k-1) - kb) / 8) LSD passes.
// Returns false (uniformly, nothing written to S.pa) when a bin holds more than P.big_bin items: low-complexity windows
// take the LSD path below.

// The bin bits are taken from (key word 0 - base0) << kb, base0 = the prefix of the window's first tile and kb = the leading
// zeros of the window's span in that word: a window of a few consecutive prefix tiles then spreads over at least half of
// the bins wherever it lies.  (Taking the bits below the prefix the window SHARES wasted bins whenever the window crossed
// an aligned boundary -- 2^j-fold with probability 2^-j, so every level cost the same: simulated mean bin occupancy 9.5
// instead of 2.2, and every window past 64 per bin fell to the LSD passes.)  Subtracting a constant from the first word
// keeps the key order, so everything below works on K' = (word 0 - base0, word 1, ...).
template <int W>
__device__ __forceinline__ bool bucket_rank_sort(ChunkSmem &S, uint32_t *qk, unsigned capi, unsigned n, int kb, uint32_t base0, int D,
                                                 unsigned *s_big, unsigned big_bin, int k) {
    // bin counters: u16 pairs packed in the (otherwise idle) per-warp LSD counter array -- 32-bit shared atomics on the
    // half that belongs to the bin return the arrival rank; <= 4096 items per window, so a half never carries over.
    // qk[i]: the 32 key bits that follow the bin bits.  When what is left of `S a` fits in 28 bits (k = 31: always) the
    // low 4 bits hold the flag nibble ((a != $) << 3 | b) and qk alone orders a bin: the comparison rank reads ONE word
    // per pair.  Longer keys compare the remaining key words only when qk ties.
    uint32_t *hist = reinterpret_cast<uint32_t *>(S.whist);       // [NB / 2] packed (low half = even bin)
    D = min(D, 32 - kb);                                          // the bin bits come from key word 0 only (kb <= 24: D >= 8)
    const unsigned NB = 1u << D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (unsigned i = tid; i < NB / 2; i += CHUNK_THREADS) hist[i] = 0;
    if (tid == 0) *s_big = 0;
    __syncthreads();
    const int q_off = kb + D;                                     // first key bit below the bin bits
    const int left = 2 * (k - 1) + 2 - q_off;                     // bits of `S a` below the bin bits
    const bool exact = left <= 28;
    const bool semi = !exact && left <= 32;                       // qk holds all of `S a`: ties differ in the flag nibble only
    const int qw = q_off >> 5, qs = q_off & 31;
    unsigned short dg[8], rk[8];
    bool big = false;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const unsigned i = tid + r * CHUNK_THREADS;
        if (i < n) {
            const uint32_t w0 = S.keys[i] - base0;
            const unsigned d = (w0 << kb) >> (32 - D);
            const unsigned sh = (d & 1u) * 16u;
            const unsigned a = (atomicAdd(&hist[d >> 1], 1u << sh) >> sh) & 0xFFFFu;
            dg[r] = (unsigned short)d; rk[r] = (unsigned short)a;
            big = big || a >= big_bin;
            const uint32_t k0 = qw == 0 ? w0 : (qw < W ? S.keys[qw * capi + i] : 0u);
            const uint32_t k1 = qw + 1 < W ? S.keys[(qw + 1) * capi + i] : 0u;
            uint32_t q = __funnelshift_l(k1, k0, qs);             // 32 key bits from bit q_off on
            if (exact) q = left > 0 ? (((q >> (32 - left)) << 4) | (S.keys[(W - 1) * capi + i] & 15u)) : (S.keys[(W - 1) * capi + i] & 15u);
            qk[i] = q;
        }
    }
    if (big) *s_big = 1;
    __syncthreads();
    if (*s_big) return false;
    // exclusive scan of the bin counters: `per` consecutive bins per thread (per even, or NB < CHUNK_THREADS * 2)
    {
        const unsigned per = NB >= 2 * CHUNK_THREADS ? NB / CHUNK_THREADS : 2u;   // 2 .. 8 bins = 1 .. 4 packed words
        unsigned c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned b = tid * per + 2 * j;
            const uint32_t v = ((unsigned)(2 * j) < per && b < NB) ? hist[b >> 1] : 0u;
            c[2 * j] = v & 0xFFFFu; c[2 * j + 1] = v >> 16;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const unsigned t = c[j]; c[j] = sum; sum += t; }
        unsigned x = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if (lane >= (unsigned)off) x += y;
        }
        if (lane == 31) S.scan[warp] = x;
        __syncthreads();
        unsigned ws = lane < CHUNK_WARPS ? S.scan[lane] : 0u;                      // inclusive scan of the warp totals
#pragma unroll
        for (int off = 1; off < CHUNK_WARPS; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xFFFFFFFFu, ws, off);
            if (lane >= (unsigned)off) ws += y;
        }
        const unsigned add = warp ? __shfl_sync(0xFFFFFFFFu, ws, warp - 1) : 0u;
        const unsigned excl = x - sum + add;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned b = tid * per + 2 * j;
            if ((unsigned)(2 * j) < per && b < NB) hist[b >> 1] = (excl + c[2 * j]) | ((excl + c[2 * j + 1]) << 16);   // starts <= 4096 fit 16 bits
        }
    }
    __syncthreads();
    const uint16_t *hs = reinterpret_cast<const uint16_t *>(hist);                 // bin start by bin index (little endian halves)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const unsigned i = tid + r * CHUNK_THREADS;
        if (i < n) S.pb[(unsigned)hs[dg[r]] + rk[r]] = (uint16_t)i;
    }
    __syncthreads();
    const int tw = (q_off + 32) >> 5;                             // first key word not fully covered by qk
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const unsigned i = tid + r * CHUNK_THREADS;
        if (i < n) {
            const unsigned d = dg[r];
            const unsigned s = hs[d], e = d + 1 < NB ? (unsigned)hs[d + 1] : n;
            unsigned pos = s;
            if (e - s > 1) {
                const uint32_t mq = qk[i];
                const uint32_t mf = S.keys[(W - 1) * capi + i] & 15u;
                for (unsigned j = s; j < e; ++j) {
                    const unsigned o = S.pb[j];
                    if (o == i) continue;                                          // itself: it would take the whole tie-break below
                    const uint32_t oq = qk[o];
                    bool less = oq < mq;
                    if (oq == mq) {
                        less = o < i;                                              // equal keys: by item index
                        if (semi) {                                                // k = 31 with 2^19 tiles: one more word, not the key
                            const uint32_t of = S.keys[(W - 1) * capi + o] & 15u;
                            less = of < mf || (of == mf && less);
                        } else if (!exact) {
#pragma unroll
                            for (int w = W - 1; w >= 0; --w) {
                                if (w >= tw) {
                                    const uint32_t x = S.keys[w * capi + o], y = S.keys[w * capi + i];
                                    less = x < y || (x == y && less);
                                }
                            }
                        }
                    }
                    pos += less ? 1u : 0u;
                }
            }
            S.pa[pos] = (uint16_t)i;
        }
    }
    __syncthreads();
    return true;
}

// CAPI_T: the window capacity as a compile-time constant (0 = P.CAPI at run time).  Every access to the staged keys is
// keys[w * capi + idx]: with a constant plane stride the multiplications become immediate offsets (they were 17 % of the
// kernel's instructions, attributed to the lines that define `capi` and `S.keys`).
template <int W, int CAPI_T>
__global__ void __launch_bounds__(CHUNK_THREADS, 2) k_sort_emit(const ChunkParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int IW = W + 1;
    const unsigned capi = CAPI_T ? (unsigned)CAPI_T : P.CAPI, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lt = (1u << lane) - 1;
    ChunkSmem S;
    uint32_t *fld;
    {
        unsigned char *p = smem_raw;
        S.keys = reinterpret_cast<uint32_t *>(p); p += (size_t)IW * capi * 4;
        fld = reinterpret_cast<uint32_t *>(p); p += (size_t)capi * 4;
        S.whist = reinterpret_cast<uint16_t *>(p); p += CHUNK_WARPS * 256 * 2;
        S.gflag = reinterpret_cast<uint32_t *>(p); p += (capi / 32 + 2) * 4;
        S.rflag = reinterpret_cast<uint32_t *>(p); p += (capi / 32 + 2) * 4;
        S.scan = reinterpret_cast<uint32_t *>(p); p += (CHUNK_WARPS + 2) * 4;
        p += 64 - ((CHUNK_WARPS + 2) * 4) % 64;
        S.pa = reinterpret_cast<uint16_t *>(p); p += (size_t)capi * 2;
        S.pb = reinterpret_cast<uint16_t *>(p);
        S.tot = nullptr;
    }
    uint32_t *code = fld;                                      // per sorted position: group / run bits, a, b, multiplicity
    __shared__ unsigned long long s_lo, s_hi, s_base;
    __shared__ unsigned s_j2[2], s_tot10[10], s_big;
    if (tid < 10) s_tot10[tid] = 0;

    for (unsigned iter = 0;; ++iter) {
        if (tid == 0) s_j2[iter & 1] = P.win_lo + atomicAdd(P.ticket, 1u);
        __syncthreads();
        const unsigned j = s_j2[iter & 1];
        if (j >= P.win_hi) break;
        if (warp == 0) { unsigned long long v = window_lo(P, j); if (lane == 0) s_lo = v; }
        if (warp == 1) { unsigned long long v = window_lo(P, j + 1); if (lane == 0) s_hi = v; }
        __syncthreads();
        const unsigned long long lo = s_lo, hi = s_hi;
        unsigned n = (unsigned)(hi - lo);
        if (hi - lo > capi) {                                  // cannot happen once the MSD levels ran (groups hold <= 48 items)
            if (tid == 0) atomicOr(P.err, (unsigned)ERR_CHUNK_TOO_BIG);
            n = 0;
        }
        // key bits all items of the window share: the tiles of a window are in prefix order, so first ^ last bounds them
        // (the LSD passes stop there); the bucket + rank sort takes its bins from the window's SPAN instead (see above)
        int kb = P.depth_min, kspan = P.depth_min;
        uint32_t base0 = 0;
        if (n > 1) {
            const uint32_t f0 = P.src[lo], l0 = P.src[lo + n - 1];
            const uint32_t low = kb >= 32 ? 0u : (0xFFFFFFFFu >> kb);             // bits below the tile prefix
            base0 = f0 & ~low;
            kspan = min(kspan, __clz(((l0 | low) - base0) | 1u));
            kb = min(kb, __clz(f0 ^ l0));
        }
        // ---- load the window (coalesced per word array)
        for (unsigned i = tid; i < n; i += CHUNK_THREADS) {
#pragma unroll
            for (int w = 0; w < IW; ++w) S.keys[w * capi + i] = P.src[(uint64_t)w * P.cap + lo + i];
            S.pa[i] = (uint16_t)i;
        }
        __syncthreads();
        // ---- LSD radix sort of the permutation by the whole key: pass 0 = the 6-bit in-group order (a slot, a != $, b),
        //      then the S bits [kb, 2(k-1)) from the least significant byte up.  One __match_any_sync per item and pass;
        //      digit, warp-local rank and item stay in registers between the counting and the scatter half.
        bool sorted = n <= 1;
        if (!sorted && P.bin_bits > 0) sorted = bucket_rank_sort<W>(S, fld, capi, n, kspan, base0, P.bin_bits, &s_big, P.big_bin, P.k);
        if (!sorted) {
            if (tid == 0 && P.n_lsd) atomicAdd(P.n_lsd, 1u);
            const unsigned slice = (((n + CHUNK_WARPS - 1) / CHUNK_WARPS) + 31) & ~31u;
            const unsigned beg = min(n, warp * slice), end = min(n, beg + slice);
            const unsigned rounds = (end - beg + 31) >> 5;    // <= capi / 512 <= 8
            const int SB = 2 * (P.k - 1);
            const int n_pass = 1 + (SB - kb + 7) / 8;
#pragma unroll 1
            for (int pass = 0; pass < n_pass; ++pass) {
                // digit of this pass: `wd` bits whose most significant one is key bit `o` (pass 0: see above)
                const int hi_bit = SB - 8 * (pass - 1);
                const int wd = pass == 0 ? 6 : min(8, hi_bit - kb);
                const int o = hi_bit - wd, wi = o >> 5, sh = o & 31;
                {
                    uint4 *z = reinterpret_cast<uint4 *>(S.whist);
                    z[tid] = make_uint4(0, 0, 0, 0);           // 512 threads x 16 B = 16 warps x 256 u16
                }
                __syncthreads();
                unsigned short idx_r[8], loc_r[8];
                unsigned char dig_r[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if ((unsigned)r < rounds) {
                        const unsigned p = beg + r * 32 + lane;
                        const bool valid = p < end;
                        const unsigned idx = valid ? S.pa[p] : 0u;
                        unsigned d;
                        if (pass == 0) {
                            d = (((S.keys[P.aw * capi + idx] >> P.ash) & 3u) << 4) | (S.keys[(W - 1) * capi + idx] & 15u);
                        } else {
                            const uint32_t k0 = S.keys[wi * capi + idx];
                            const uint32_t k1 = (sh + wd > 32) ? S.keys[(wi + 1) * capi + idx] : 0u;
                            d = __funnelshift_l(k1, k0, sh) >> (32 - wd);
                        }
                        if (!valid) d = 256u + lane;
                        const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
                        const int leader = __ffs(peers) - 1;
                        unsigned old = 0;
                        if (valid && (int)lane == leader) {
                            old = S.whist[d * CHUNK_WARPS + warp];
                            S.whist[d * CHUNK_WARPS + warp] = (uint16_t)(old + __popc(peers));
                        }
                        old = __shfl_sync(0xFFFFFFFFu, old, leader);
                        idx_r[r] = (unsigned short)idx;
                        dig_r[r] = (unsigned char)d;
                        loc_r[r] = (unsigned short)(old + __popc(peers & lt));
                        __syncwarp();
                    }
                }
                __syncthreads();
                // exclusive scan of the 4096 counters in (digit, warp) order: 8 consecutive u16 per thread
                {
                    uint4 *h4 = reinterpret_cast<uint4 *>(S.whist);
                    uint4 v = h4[tid];
                    unsigned c[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16, v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
                    unsigned sum = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const unsigned t = c[i]; c[i] = sum; sum += t; }
                    unsigned x = sum;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, off);
                        if (lane >= (unsigned)off) x += y;
                    }
                    if (lane == 31) S.scan[warp] = x;
                    __syncthreads();
                    unsigned add = 0;
#pragma unroll
                    for (int w = 0; w < CHUNK_WARPS; ++w) add += (unsigned)w < warp ? S.scan[w] : 0u;
                    const unsigned excl = x - sum + add;
#pragma unroll
                    for (int i = 0; i < 8; ++i) c[i] += excl;
                    v.x = c[0] | (c[1] << 16); v.y = c[2] | (c[3] << 16); v.z = c[4] | (c[5] << 16); v.w = c[6] | (c[7] << 16);
                    h4[tid] = v;
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if ((unsigned)r < rounds) {
                        const unsigned p = beg + r * 32 + lane;
                        if (p < end) S.pb[S.whist[(unsigned)dig_r[r] * CHUNK_WARPS + warp] + loc_r[r]] = idx_r[r];
                    }
                }
                __syncthreads();
                uint16_t *t = S.pa; S.pa = S.pb; S.pb = t;
            }
        }
        const unsigned Q = (n + CHUNK_THREADS - 1) / CHUNK_THREADS;
        const unsigned qb = min(n, tid * Q), qe = min(n, qb + Q);
        // ---- per sorted position: group / run bits, a, b, capped multiplicity
        for (unsigned b = warp * 32; b < ((n + 31) & ~31u); b += CHUNK_THREADS) {
            const unsigned i = b + lane;
            bool g = false;
            uint32_t c = 0;
            if (i < n) {
                const unsigned x = S.pa[i];
                bool r;
                if (i == 0) { g = r = true; }
                else {
                    const unsigned y = S.pa[i - 1];
                    g = group_diff(S.keys, capi, x, y, P.g_full, P.g_rem_shift);
                    r = g;
                    if (!r)
                        for (int w = P.g_full; w < W; ++w)
                            if (S.keys[w * capi + x] != S.keys[w * capi + y]) { r = true; break; }
                    // sortedness is checked on every window of every run (one position in 16: warp 0's share -- a mis-sorted
                    // window is wrong all over, and the full check costs 3 ms of 67): an unsorted window would silently
                    // split groups into extra records
                    if (warp == 0 && r && key_cmp<W>(S.keys, capi, x, y) < 0) atomicOr(P.err, (unsigned)ERR_SORT_ORDER);
                }
                const uint32_t lw = S.keys[(W - 1) * capi + x];
                const uint32_t a = ((lw >> 3) & 1) ? ((S.keys[P.aw * capi + x] >> P.ash) & 3u) : (uint32_t)SENT;
                const uint32_t mult = S.keys[W * capi + x];
                c = ((uint32_t)g << 31) | ((uint32_t)r << 30) | (a << 19) | ((lw & 7u) << 16) | (mult < 65535u ? mult : 65535u);
            }
            const unsigned gb = __ballot_sync(0xFFFFFFFFu, g);
            if (lane == 0) S.gflag[b >> 5] = gb;
            if (i < n) code[i] = c;
        }
        __syncthreads();
        // ---- emission: the groups that start in this thread's block of positions
        unsigned long long accA = 0, accB = 0;                 // records per w: 5 + 4 fields of 12 bits
        unsigned accL = 0;                                     // records with last = 1
        {
            unsigned cur_bucket = 0xFFFFFFFFu, m_items = 0, m_tips = 0, m_large = 0;
            for (unsigned gs = next_flag(S.gflag, qb, qe); gs < qe; gs = next_flag(S.gflag, gs + 1, qe)) {
                const unsigned ge = next_flag(S.gflag, gs + 1, n);
                // pass 1 (s2.cpp:766-780): which a / b have a real edge, last qualifying run per a
                unsigned hsa = 0, hsb = 0, last_run = 0xFFFFFFFFu;
                int ridx = -1;
                for (unsigned t = gs; t < ge; ++t) {
                    const uint32_t c = code[t];
                    ridx += (c >> 30) & 1;
                    const unsigned a = (c >> 19) & 7, b = (c >> 16) & 7;
                    if (a != SENT && b != SENT) { hsa |= 1u << a; hsb |= 1u << b; }
                    if (a != SENT && (b != SENT || !((hsa >> a) & 1))) last_run = (last_run & ~(0xFFu << (8 * a))) | ((unsigned)ridx << (8 * a));
                }
                // pass 2 (s2.cpp:801-830): one record per surviving run, descriptor at the run start
                unsigned outb = 0;
                ridx = -1;
                unsigned t = gs;
                unsigned g_items = 0, g_tips = 0, g_large = 0;
                while (t < ge) {
                    const uint32_t c0 = code[t];
                    unsigned sum = c0 & 0xFFFFu, t2 = t + 1;
                    while (t2 < ge && !((code[t2] >> 30) & 1)) { sum += code[t2] & 0xFFFFu; code[t2] = 0; ++t2; }
                    ++ridx;
                    const unsigned a = (c0 >> 19) & 7, b = (c0 >> 16) & 7;
                    const unsigned tip = a == SENT;
                    const bool skip = (tip && ((hsb >> b) & 1)) || (b == SENT && ((hsa >> a) & 1));
                    uint32_t d = 0;
                    if (!skip) {
                        const unsigned w = b == SENT ? 0u : (((outb >> b) & 1) ? b + 5 : b + 1);
                        const unsigned last = tip ? 0u : (unsigned)(((last_run >> (8 * a)) & 0xFFu) == (unsigned)ridx);
                        outb |= 1u << b;
                        const unsigned mult = sum > 65535u ? 65535u : sum;
                        d = s2_record_word((int)w, (int)last, (int)tip, mult) | (mult << 16);
                        ++g_items; g_tips += tip; g_large += mult > 254u;
                        if (w < 5) accA += 1ull << (12 * w); else accB += 1ull << (12 * (w - 5));
                        accL += last;
                    }
                    code[t] = d;
                    t = t2;
                }
                const unsigned bucket = S.keys[S.pa[gs]] >> 16;
                if (bucket != cur_bucket) {
                    if (m_items) atomicAdd(P.meta + cur_bucket * 3 + 0, (unsigned long long)m_items);
                    if (m_tips) atomicAdd(P.meta + cur_bucket * 3 + 1, (unsigned long long)m_tips);
                    if (m_large) atomicAdd(P.meta + cur_bucket * 3 + 2, (unsigned long long)m_large);
                    m_items = m_tips = m_large = 0;
                    cur_bucket = bucket;
                }
                m_items += g_items; m_tips += g_tips; m_large += g_large;
            }
            if (m_items) atomicAdd(P.meta + cur_bucket * 3 + 0, (unsigned long long)m_items);
            if (m_tips) atomicAdd(P.meta + cur_bucket * 3 + 1, (unsigned long long)m_tips);
            if (m_large) atomicAdd(P.meta + cur_bucket * 3 + 2, (unsigned long long)m_large);
        }
        // w / last totals of the window: warp reduce, one shared atomic per counter and warp
#pragma unroll
        for (int w = 0; w < 10; ++w) {
            const unsigned v = w < 5 ? (unsigned)(accA >> (12 * w)) & 4095u : (w < 9 ? (unsigned)(accB >> (12 * (w - 5))) & 4095u : accL);
            const unsigned s = __reduce_add_sync(0xFFFFFFFFu, v);
            if (lane == 0 && s) atomicAdd(&s_tot10[w], s);
        }
        __syncthreads();
        // ---- sizes of this thread's block of positions -> offsets
        unsigned bytes = 0;
        for (unsigned p = qb; p < qe; ++p) {
            const uint32_t d = code[p];
            if (d) bytes += s2_record_bytes((d >> 5) & 1, d >> 16, P.wpt);
        }
        unsigned total = 0;
        const unsigned my_off = block_exclusive_scan(S, bytes, total);
        if (tid == 0) {                                        // any order: k_out_gather restores window order afterwards
            const unsigned long long b0 = atomicAdd(P.state, (unsigned long long)total);
            s_base = b0;
            P.state[1 + 2 * (size_t)j] = b0;
            P.state[2 + 2 * (size_t)j] = total;
            if (b0 + total > P.out_cap) atomicOr(P.err, (unsigned)ERR_OUT_OVERFLOW);
        }
        __syncthreads();
        const unsigned long long base = s_base;
        if (base + total <= P.out_cap) {
            unsigned short *o = reinterpret_cast<unsigned short *>(P.out + base + my_off);
            for (unsigned p = qb; p < qe; ++p) {
                const uint32_t d = code[p];
                if (!d) continue;
                *o++ = (unsigned short)d;
                if ((d >> 16) > 254u) *o++ = (unsigned short)(d >> 16);
                if ((d >> 5) & 1) {
                    const unsigned item = S.pa[p];
                    for (int i = 0; i < P.wpt; ++i) {
                        const uint32_t x = S.keys[i * capi + item];
                        *o++ = (unsigned short)(x & 0xFFFFu);
                        *o++ = (unsigned short)(x >> 16);
                    }
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid < 10 && s_tot10[tid]) atomicAdd(P.totals + tid, (unsigned long long)s_tot10[tid]);
}

// ---- window order: exclusive scan of the window byte counts, then a copy of every window to its final place
// state layout: [0] cursor, then per window (temp offset, bytes); after the scan the pair holds (temp offset, final offset)
__global__ void __launch_bounds__(1024) k_out_scan(unsigned long long *state, unsigned w_lo, unsigned w_hi, unsigned long long *total_io,
                                                   unsigned long long *end_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = *total_io;                         // bytes of the windows before w_lo (earlier launches of the batch)
    __syncthreads();
    for (unsigned base = w_lo; base < w_hi; base += 1024) {
        const unsigned j = base + tid;
        const unsigned long long v = j < w_hi ? state[2 + 2 * (size_t)j] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        unsigned long long add = s_carry;
        for (unsigned w = 0; w < warp; ++w) add += s_w[w];
        if (j < w_hi) state[2 + 2 * (size_t)j] = x - v + add | (v << 40);            // final offset (40 bits) | bytes (24 bits)
        __syncthreads();
        if (tid == 1023) s_carry = add + x;
        __syncthreads();
    }
    if (tid == 0) { *total_io = s_carry; if (end_out) *end_out = s_carry; }
}

__global__ void __launch_bounds__(256) k_out_gather(const unsigned long long *__restrict__ state, unsigned w_lo, unsigned w_hi,
                                                    const unsigned char *__restrict__ tmp, unsigned char *__restrict__ out) {
    for (unsigned j = w_lo + blockIdx.x; j < w_hi; j += gridDim.x) {
        const unsigned long long src = state[1 + 2 * (size_t)j], pk = state[2 + 2 * (size_t)j];
        const unsigned long long dst = pk & ((1ull << 40) - 1);
        const unsigned n2 = (unsigned)(pk >> 40) >> 1;                               // u16 units (all record sizes are even)
        const unsigned short *s = reinterpret_cast<const unsigned short *>(tmp + src);
        unsigned short *d = reinterpret_cast<unsigned short *>(out + dst);
        for (unsigned i = threadIdx.x; i < n2; i += 256) d[i] = s[i];
    }
}

}  // namespace mgta
