// tests/cpu/logic_host.cpp -- TEST INFRASTRUCTURE.  Runs the product's __host__ __device__ item and
// emission logic (megagta_b200/csrc/{kmer_ops,cx1_items,cx1_emit}.cuh) on the CPU with std::sort in
// place of the device sort, so that logic can be checked against the oracle in the GPU-less build
// container.  Not shipped, not a fallback: libmgta_cuda.so never contains this file.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../megagta_b200/csrc/cx1_emit.cuh"

using namespace mgta;

struct Reads {
    const uint32_t *seq; const uint64_t *start; int64_t n_reads, n_short; int max_len, k, m;
};

template <int W> struct Item1 { uint32_t key[W]; uint64_t val; };

template <int W>
static void stage1_t(const Reads &rd, uint8_t *is_solid, int64_t *edge_counting, bool mercy, std::vector<uint64_t> &cands,
                     int64_t *hist) {
    std::vector<Item1<W>> items;
    const int k = rd.k;
    for (int64_t r = 0; r < rd.n_reads; ++r) {
        uint64_t s = rd.start[r]; int L = (int)(rd.start[r + 1] - s);
        if (L < k + 1) continue;
        for (int p = 0; p <= L - k + 1; ++p) {
            uint64_t g = s + p;
            // stage a window exactly like the kernel: words from 4 words (64 chars) before the position
            uint64_t w0 = (g >> 4) >= 4 ? (g >> 4) - 4 : 0;
            s1_position<W>(rd.seq + w0, (uint32_t)(g - 16 * w0), g, p, L, k, r < rd.n_short,
                           [&](const uint32_t(&key)[W], uint64_t v) {
                               Item1<W> it; memcpy(it.key, key, sizeof(it.key)); it.val = v; items.push_back(it);
                               if (hist) hist[key[0] >> 16]++;
                           });
        }
    }
    if (!is_solid) return;
    std::sort(items.begin(), items.end(), [](const Item1<W> &a, const Item1<W> &b) {
        for (int i = 0; i < W; ++i) if (a.key[i] != b.key[i]) return a.key[i] < b.key[i];
        return false;
    });
    const int full = (k - 1) / 16, rem = (k - 1) % 16;
    auto same_group = [&](const Item1<W> &a, const Item1<W> &b) {
        for (int w = 0; w < full; ++w) if (a.key[w] != b.key[w]) return false;
        if (rem && (a.key[full] >> (16 - rem) * 2) != (b.key[full] >> (16 - rem) * 2)) return false;
        return true;
    };
    const int64_t nk1 = rd.max_len - k;
    size_t n = items.size();
    for (size_t i = 0, e; i < n; i = e) {
        for (e = i + 1; e < n && same_group(items[i], items[e]); ++e) {}
        Sat16 cph, ctn, cht; cph.clear(); ctn.clear(); cht.clear();
        if (mercy)
            for (size_t j = i; j < e; ++j) {
                int ht = items[j].key[W - 1] & 63, pn = items[j].val & 63;
                int head = ht >> 3, tail = ht & 7, prev = pn >> 3, next = pn & 7;
                if (prev < 4 && head < 4) cph.add(prev * 4 + head, 1);
                if (tail < 4 && next < 4) ctn.add(tail * 4 + next, 1);
                if (head < 4 && tail < 4) cht.add(head * 4 + tail, 1);
            }
        S1GroupMasks gm = s1_group_masks(cph, ctn, cht, (uint32_t)rd.m);
        for (size_t j = i, j2; j < e; j = j2) {                    // runs of equal full key
            for (j2 = j + 1; j2 < e && items[j2].key[W - 1] == items[j].key[W - 1]; ++j2) {}
            int ht = items[j].key[W - 1] & 63, head = ht >> 3, tail = ht & 7;
            uint32_t cnt = (uint32_t)(j2 - j);
            bool real = head != SENT && tail != SENT;
            if (real) edge_counting[cnt < 65535 ? cnt : 65535]++;
            bool solid = real && cnt >= (uint32_t)rd.m;
            for (size_t t = j; t < j2; ++t) {
                uint64_t kpos = items[t].val >> 8; int strand = (items[t].val >> 6) & 1;
                if (kpos == S1_NO_EDGE) continue;
                if (solid) {
                    // internal layout: bit per absolute base position (kpos - 1); convert to the reference's
                    uint64_t edge = kpos - 1;
                    int64_t r = std::upper_bound(rd.start, rd.start + rd.n_reads + 1, edge) - rd.start - 1;
                    int64_t bit = nk1 * r + (int64_t)(edge - rd.start[r]);
                    is_solid[bit >> 3] |= 1u << (bit & 7);
                }
                if (mercy) s1_mercy_item(gm, solid, head, tail, strand, kpos, [&](uint64_t pos, int flag) { cands.push_back((pos << 2) | flag); });
            }
        }
    }
    std::sort(cands.begin(), cands.end());
}

template <int W> struct Item2 { uint32_t key[W]; };

struct Out2 {
    std::vector<uint8_t> bytes; int64_t *meta; int64_t *totals; int wpt;
};

template <int W>
static void stage2_t(const Reads &rd, const uint8_t *is_solid, Out2 &out, int64_t *hist) {
    std::vector<Item2<W>> items;
    const int k = rd.k; const int64_t nk1 = rd.max_len - k;
    for (int64_t r = 0; r < rd.n_reads; ++r) {
        uint64_t s = rd.start[r]; int L = (int)(rd.start[r + 1] - s);
        if (L < k + 1) continue;
        bool all = rd.m == 1 || r >= rd.n_short;
        auto solid = [&](int o) { int64_t bit = nk1 * r + o; return all || ((is_solid[bit >> 3] >> (bit & 7)) & 1); };
        for (int o = 0; o < L - k; ++o) {
            if (!solid(o)) continue;
            uint64_t g = s + o; uint64_t w0 = (g >> 4) >= 4 ? (g >> 4) - 4 : 0;
            s2_position<W>(rd.seq + w0, (uint32_t)(g - 16 * w0), o, L, k, o > 0 && solid(o - 1), o < L - k - 1 && solid(o + 1),
                           [&](const uint32_t(&key)[W]) { Item2<W> it; memcpy(it.key, key, sizeof(it.key)); items.push_back(it);
                                                         if (hist) hist[key[0] >> 16]++; });
        }
    }
    if (!out.meta) return;
    std::sort(items.begin(), items.end(), [](const Item2<W> &a, const Item2<W> &b) {
        for (int i = 0; i < W; ++i) if (a.key[i] != b.key[i]) return a.key[i] < b.key[i];
        return false;
    });
    const int full = (k - 1) / 16, rem = (k - 1) % 16;
    const int aw = (k - 1) >> 4, ash = (15 - ((k - 1) & 15)) * 2;
    auto same_group = [&](const Item2<W> &a, const Item2<W> &b) {
        for (int w = 0; w < full; ++w) if (a.key[w] != b.key[w]) return false;
        if (rem && (a.key[full] >> (16 - rem) * 2) != (b.key[full] >> (16 - rem) * 2)) return false;
        return true;
    };
    size_t n = items.size();
    struct Runs {
        const std::vector<Item2<W>> *it; size_t i, e, cur; int aw, ash;
        void reset() { cur = i; }
        bool next(S2Run &r) {
            if (cur >= e) return false;
            const Item2<W> &x = (*it)[cur];
            size_t j = cur + 1;
            while (j < e && memcmp((*it)[j].key, x.key, sizeof(x.key)) == 0) ++j;
            uint32_t lw = x.key[W - 1];
            r.a = (lw >> 3 & 1) ? (int)((x.key[aw] >> ash) & 3) : SENT; r.b = lw & 7; r.cnt = (uint32_t)(j - cur); r.item = (uint32_t)cur;
            cur = j; return true;
        }
    };
    struct Sink {
        Out2 *o; const std::vector<Item2<W>> *it;
        void record(int w, int last, int tip, uint32_t mult, uint32_t item) {
            const Item2<W> &x = (*it)[item];
            int bucket = x.key[0] >> 16;
            uint16_t rec = s2_record_word(w, last, tip, mult);
            o->bytes.insert(o->bytes.end(), (uint8_t *)&rec, (uint8_t *)&rec + 2);
            o->meta[bucket * 3]++; o->totals[w]++; o->totals[9] += last;
            if (mult > 254) { uint16_t mm = (uint16_t)mult; o->bytes.insert(o->bytes.end(), (uint8_t *)&mm, (uint8_t *)&mm + 2); o->meta[bucket * 3 + 2]++; }
            if (tip) { o->bytes.insert(o->bytes.end(), (uint8_t *)x.key, (uint8_t *)x.key + 4 * o->wpt); o->meta[bucket * 3 + 1]++; }
        }
    };
    for (size_t i = 0, e; i < n; i = e) {
        for (e = i + 1; e < n && same_group(items[i], items[e]); ++e) {}
        Runs runs{&items, i, e, i, aw, ash};
        Sink sink{&out, &items};
        s2_emit_group(runs, sink);
    }
}

#define DISPATCH(W, CALL) switch (W) { case 1: { constexpr int WW = 1; CALL; } break; case 2: { constexpr int WW = 2; CALL; } break; \
    case 3: { constexpr int WW = 3; CALL; } break; case 4: { constexpr int WW = 4; CALL; } break; case 5: { constexpr int WW = 5; CALL; } break; \
    case 6: { constexpr int WW = 6; CALL; } break; case 7: { constexpr int WW = 7; CALL; } break; case 8: { constexpr int WW = 8; CALL; } break; \
    case 9: { constexpr int WW = 9; CALL; } break; default: return -1; }

extern "C" {
int logic_stage1(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short, int max_len, int k, int m,
                 uint8_t *is_solid, int64_t *edge_counting, int need_mercy, uint64_t **cand_out, int64_t *n_cand, int64_t *hist) {
    Reads rd{seq, start, n_reads, n_short, max_len, k, m};
    std::vector<uint64_t> cands;
    if (edge_counting) memset(edge_counting, 0, 65536 * 8);
    if (hist) memset(hist, 0, 65536 * 8);
    DISPATCH(key_words_s1(k), (stage1_t<WW>(rd, is_solid, edge_counting, need_mercy != 0, cands, hist)));
    if (cand_out) {
        *cand_out = (uint64_t *)malloc(cands.size() * 8 + 8);
        memcpy(*cand_out, cands.data(), cands.size() * 8);
        *n_cand = (int64_t)cands.size();
    }
    return 0;
}
int logic_stage2(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short, int max_len, int k, int m,
                 const uint8_t *is_solid, uint8_t **stream, int64_t *stream_bytes, int64_t *meta, int64_t *totals, int64_t *hist) {
    Reads rd{seq, start, n_reads, n_short, max_len, k, m};
    Out2 out; out.meta = meta; out.totals = totals; out.wpt = (2 * k + 31) / 32;
    if (meta) { memset(meta, 0, 65536 * 3 * 8); memset(totals, 0, 10 * 8); }
    if (hist) memset(hist, 0, 65536 * 8);
    DISPATCH(key_words_s2(k), (stage2_t<WW>(rd, is_solid, out, hist)));
    if (stream) {
        *stream = (uint8_t *)malloc(out.bytes.size() + 8);
        memcpy(*stream, out.bytes.data(), out.bytes.size());
        *stream_bytes = (int64_t)out.bytes.size();
    }
    return 0;
}
void logic_free(void *p) { free(p); }

// The device's per-read mercy scan (cx1_emit.cuh mercy_scan_read, what k_mercy_bits + k_mercy_reads run) on the CPU:
// candidates -> three flag vectors over base positions -> scan of every short read.  is_solid: the reference layout
// ((max_len - k) * read + offset), updated in place.  Returns "Number mercy".
int64_t logic_mercy(const uint64_t *start, int64_t n_short, int max_len, int k, const uint64_t *cands, int64_t n_cand,
                    uint8_t *is_solid) {
    const uint64_t total = start[n_short];
    std::vector<uint8_t> v[3];
    for (auto &x : v) x.assign(total + 1, 0);
    for (int64_t i = 0; i < n_cand; ++i) {
        const uint64_t pos = cands[i] >> 2;
        const int flag = (int)(cands[i] & 3);
        if (pos > total) return -1;
        if (flag == 1) v[0][pos] = 1;
        if (flag == 2) v[1][pos] = 1;
        v[2][pos] = 1;
    }
    const int64_t nk1 = max_len - k;
    int64_t num = 0;
    for (int64_t r = 0; r < n_short; ++r) {
        const uint64_t s0 = start[r];
        const int L = (int)(start[r + 1] - s0);
        num += (int64_t)mercy_scan_read(L, k,
                                        [&](int w, int i) {
                                            if (w == 3) { const int64_t b = nk1 * r + i; return (bool)((is_solid[b >> 3] >> (b & 7)) & 1); }
                                            return (bool)v[w][s0 + (uint64_t)i];
                                        },
                                        [&](int j) { const int64_t b = nk1 * r + j; is_solid[b >> 3] |= (uint8_t)(1u << (b & 7)); });
    }
    return num;
}
}

// ---- edge-centric path (v2): canonical (k+1)-mer multiset -> stage-1 outputs, and
// {(canonical edge, solid occurrences)} -> stage-2 records with multiplicity-weighted runs.
#include <map>
#include <array>
typedef std::array<uint32_t, 9> EKey;

template <int WE>
static void count_edges_t(const Reads &rd, const uint8_t *is_solid_filter, std::map<EKey, std::pair<uint32_t, uint32_t>> &tab,
                          std::vector<std::pair<EKey, uint64_t>> *occ) {
    const int k = rd.k; const int64_t nk1 = rd.max_len - k;
    for (int64_t r = 0; r < rd.n_reads; ++r) {
        uint64_t s = rd.start[r]; int L = (int)(rd.start[r + 1] - s);
        if (L < k + 1) continue;
        const bool assist = r >= rd.n_short;
        for (int o = 0; o < L - k; ++o) {
            if (is_solid_filter) {
                int64_t bit = nk1 * r + o;
                if (!(rd.m == 1 || assist || ((is_solid_filter[bit >> 3] >> (bit & 7)) & 1))) continue;
            }
            uint64_t g = s + o; uint64_t w0 = (g >> 4) >= 4 ? (g >> 4) - 4 : 0;
            uint32_t key[WE];
            canonical_edge<WE>(rd.seq + w0, (uint32_t)(g - 16 * w0), k, key);
            EKey ek; ek.fill(0); for (int w = 0; w < WE; ++w) ek[w] = key[w];
            auto &c = tab[ek]; c.first++; if (assist) c.second++;
            if (occ && !assist) occ->push_back({ek, (uint64_t)(nk1 * r + o)});
        }
    }
}

template <int W> struct Item2W { uint32_t key[W]; uint32_t mult; };

static int64_t g_last_items = 0;      // stage-2 items the last logic_stage2_edges call generated (before the group logic)

#include <set>
template <int W, int WE>
static void emit_edges_t(const Reads &rd, const std::map<EKey, std::pair<uint32_t, uint32_t>> &tab, bool apply_threshold, Out2 &out,
                         bool node_filter = false, bool node_pass = false) {
    const int k = rd.k;
    std::vector<Item2W<W>> items;
    auto mult_of = [&](const std::pair<uint32_t, uint32_t> &c) {
        return apply_threshold ? (c.first >= (uint32_t)rd.m ? c.first : c.second) : c.first;
    };
    // node filter: a $-item is dropped by output_() exactly when a solid edge enters (left $) /
    // leaves (right $) the k-mer it hangs off, so it need not be generated then.  in_nodes / out_nodes: the k-mers some solid
    // oriented edge ends in / starts from.
    typedef std::array<uint32_t, W> NKey;
    std::set<NKey> in_nodes, out_nodes;
    auto node = [&](const uint32_t (&X)[W], int c) { uint32_t Y[W]; sub_chars<W>(X, c, k, Y); NKey n; for (int w = 0; w < W; ++w) n[w] = Y[w]; return n; };
    if (node_filter)
        for (auto &kv : tab) {
            if (!mult_of(kv.second)) continue;
            uint32_t E[W], R[W];
            for (int w = 0; w < W; ++w) E[w] = w < WE ? kv.first[w] : 0u;
            revcomp<W>(E, k + 1, R);
            out_nodes.insert(node(E, 0)); in_nodes.insert(node(E, 1));
            out_nodes.insert(node(R, 0)); in_nodes.insert(node(R, 1));
        }
    // node pass (the product's formulation): out / in weights per canonical k-mer from
    // node_ops_of_edge, the $-items of the tip k-mers from s2_tip_items, real items only from the edges
    if (node_pass) {
        std::map<NKey, std::pair<uint64_t, uint64_t>> nodes;          // canonical k-mer -> (out, in)
        for (auto &kv : tab) {
            const uint32_t mult = mult_of(kv.second);
            if (!mult) continue;
            uint32_t E[WE];
            for (int w = 0; w < WE; ++w) E[w] = kv.first[w];
            node_ops_of_edge<WE>(E, k, [&](const uint32_t(&c)[WE], int dir) {
                NKey n; n.fill(0);
                for (int w = 0; w < kmer_words(k); ++w) n[w] = c[w];
                auto &acc = nodes[n];
                (dir ? acc.second : acc.first) += std::min<uint32_t>(mult, 65535u);
            });
            auto put = [&](const uint32_t(&y)[W]) { Item2W<W> it; memcpy(it.key, y, sizeof(it.key)); it.mult = mult; items.push_back(it); };
            uint32_t key[WE]; for (int w = 0; w < WE; ++w) key[w] = kv.first[w];
            s2_items_of_edge<W, WE>(key, k, put, false);
        }
        for (auto &kv : nodes) {
            const uint64_t o = kv.second.first, i = kv.second.second;
            uint32_t C[W], RC[W];
            for (int w = 0; w < W; ++w) C[w] = kv.first[w];
            revcomp<W>(C, k, RC);
            if (cmp_words<W>(C, RC) == 0) continue;
            if ((o == 0) == (i == 0)) continue;
            const uint32_t wgt = (uint32_t)std::min<uint64_t>(o + i, 65535u);
            s2_tip_items<W>(C, RC, i == 0, k, [&](const uint32_t(&y)[W]) { Item2W<W> it; memcpy(it.key, y, sizeof(it.key)); it.mult = wgt; items.push_back(it); });
        }
    }
    for (auto &kv : tab) {
        if (node_pass) break;
        uint32_t mult = mult_of(kv.second);
        if (!mult) continue;
        auto put = [&](const uint32_t(&y)[W]) { Item2W<W> it; memcpy(it.key, y, sizeof(it.key)); it.mult = mult; items.push_back(it); };
        if (node_filter) {
            uint32_t E[W], R[W];
            for (int w = 0; w < W; ++w) E[w] = w < WE ? kv.first[w] : 0u;
            revcomp<W>(E, k + 1, R);
            const bool left = !in_nodes.count(node(E, 0)), right = !out_nodes.count(node(E, 1));
            s2_edge_items<W>(E, R, cmp_words<W>(E, R) == 0, left, right, k, put);
        } else {
            uint32_t key[WE]; for (int w = 0; w < WE; ++w) key[w] = kv.first[w];
            s2_items_of_edge<W, WE>(key, k, put);
        }
    }
    g_last_items = (int64_t)items.size();
    std::sort(items.begin(), items.end(), [](const Item2W<W> &a, const Item2W<W> &b) {
        for (int i = 0; i < W; ++i) if (a.key[i] != b.key[i]) return a.key[i] < b.key[i];
        return false;
    });
    const int full = (k - 1) / 16, rem = (k - 1) % 16;
    const int aw = (k - 1) >> 4, ash = (15 - ((k - 1) & 15)) * 2;
    auto same_group = [&](const Item2W<W> &a, const Item2W<W> &b) {
        for (int w = 0; w < full; ++w) if (a.key[w] != b.key[w]) return false;
        if (rem && (a.key[full] >> (16 - rem) * 2) != (b.key[full] >> (16 - rem) * 2)) return false;
        return true;
    };
    struct Runs {
        const std::vector<Item2W<W>> *it; size_t i, e, cur; int aw, ash;
        void reset() { cur = i; }
        bool next(S2Run &r) {
            if (cur >= e) return false;
            const Item2W<W> &x = (*it)[cur];
            size_t j = cur; uint64_t sum = 0;
            while (j < e && memcmp((*it)[j].key, x.key, sizeof(x.key)) == 0) { sum += (*it)[j].mult; ++j; }
            uint32_t lw = x.key[W - 1];
            r.a = (lw >> 3 & 1) ? (int)((x.key[aw] >> ash) & 3) : SENT; r.b = lw & 7; r.cnt = (uint32_t)std::min<uint64_t>(sum, 0xFFFFFFFFu); r.item = (uint32_t)cur;
            cur = j; return true;
        }
    };
    struct Sink {
        Out2 *o; const std::vector<Item2W<W>> *it;
        void record(int w, int last, int tip, uint32_t mult, uint32_t item) {
            const Item2W<W> &x = (*it)[item];
            int bucket = x.key[0] >> 16;
            uint16_t rec = s2_record_word(w, last, tip, mult);
            o->bytes.insert(o->bytes.end(), (uint8_t *)&rec, (uint8_t *)&rec + 2);
            o->meta[bucket * 3]++; o->totals[w]++; o->totals[9] += last;
            if (mult > 254) { uint16_t mm = (uint16_t)mult; o->bytes.insert(o->bytes.end(), (uint8_t *)&mm, (uint8_t *)&mm + 2); o->meta[bucket * 3 + 2]++; }
            if (tip) { o->bytes.insert(o->bytes.end(), (uint8_t *)x.key, (uint8_t *)x.key + 4 * o->wpt); o->meta[bucket * 3 + 1]++; }
        }
    };
    size_t n = items.size();
    for (size_t i = 0, e; i < n; i = e) {
        for (e = i + 1; e < n && same_group(items[i], items[e]); ++e) {}
        Runs runs{&items, i, e, i, aw, ash};
        Sink sink{&out, &items};
        s2_emit_group(runs, sink);
    }
}

#define DISPATCH2(W, WE, CALL) DISPATCH(W, { constexpr int W2 = WW; if (WE == W2) { constexpr int WEE = W2; CALL; } else { constexpr int WEE = W2 > 1 ? W2 - 1 : 1; CALL; } })

extern "C" {
// stage 1 from the canonical-edge multiset: edge_counting + is_solid (reference bit layout)
int logic_stage1_edges(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short, int max_len, int k, int m,
                       uint8_t *is_solid, int64_t *edge_counting) {
    Reads rd{seq, start, n_reads, n_short, max_len, k, m};
    std::map<EKey, std::pair<uint32_t, uint32_t>> tab;
    std::vector<std::pair<EKey, uint64_t>> occ;
    memset(edge_counting, 0, 65536 * 8);
    DISPATCH(edge_words(k), (count_edges_t<WW>(rd, nullptr, tab, &occ)));
    for (auto &kv : tab) edge_counting[kv.second.first < 65535 ? kv.second.first : 65535]++;
    for (auto &o : occ) if (tab[o.first].first >= (uint32_t)m) is_solid[o.second >> 3] |= 1u << (o.second & 7);
    return 0;
}
// stage 2 from {(canonical edge, solid occurrences)}.  fused != 0: take the multiplicities straight from the unfiltered
// stage-1 counts (count >= m ? count : assist occurrences) instead of re-counting under the is_solid filter.
int logic_stage2_edges(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short, int max_len, int k, int m,
                       const uint8_t *is_solid, int fused, uint8_t **stream, int64_t *stream_bytes, int64_t *meta, int64_t *totals) {
    Reads rd{seq, start, n_reads, n_short, max_len, k, m};
    Out2 out; out.meta = meta; out.totals = totals; out.wpt = (2 * k + 31) / 32;
    memset(meta, 0, 65536 * 3 * 8); memset(totals, 0, 10 * 8);
    std::map<EKey, std::pair<uint32_t, uint32_t>> tab;
    static uint8_t dummy[8];
    DISPATCH(edge_words(k), (count_edges_t<WW>(rd, (fused & 1) ? nullptr : (is_solid ? is_solid : dummy), tab, nullptr)));
    const int WE = edge_words(k);
    DISPATCH2(key_words_s2(k), WE, (emit_edges_t<W2, WEE>(rd, tab, (fused & 1) != 0, out, (fused & 2) != 0, (fused & 4) != 0)));
    *stream = (uint8_t *)malloc(out.bytes.size() + 8);
    memcpy(*stream, out.bytes.data(), out.bytes.size());
    *stream_bytes = (int64_t)out.bytes.size();
    return 0;
}
int64_t logic_last_items(void) { return g_last_items; }
}
