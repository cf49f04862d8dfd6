// cx1_emit.cuh -- per-group SdBG record emission (stage 2) and mercy classification (stage 1),
// written over RUNS of equal keys so the same code serves the on-chip sorted tile, the counted
// "giant group" path and the CPU logic test.
//
// Stage 2 follows reference output_() s2.cpp:742-835 and SdbgWriter::write sdbg_multi_io.h:83-112.
// A group = all items with equal (k-1)-mer S; inside it keys sort by (a-slot, a!=$, b), i.e. runs
// arrive as ($,A) ($,C) ($,G) ($,T) (A,A)..(A,T) (A,$) (C,A) ... (T,$)  [written (a,b)].
// The reference's per-item `last_a[a]` (s2.cpp:766-780) is constant over a run, so tracking the
// last qualifying RUN per `a` is exact.
#pragma once
#include "cx1_items.cuh"

namespace mgta {

struct S2Run {
    int a, b;         // 0..3 or SENT
    uint32_t cnt;     // occurrences (uncapped)
    uint32_t item;    // handle of one item of the run (for the tip label words)
};

// Runs: void reset(); bool next(S2Run&).   Sink: void record(int w, int last, int tip, uint32_t mult, uint32_t item)
template <class Runs, class Sink>
MGTA_HD void s2_emit_group(Runs &runs, Sink &sink) {
    int hsa = 0, hsb = 0;
    uint32_t last_run = 0xFFFFFFFFu;                       // 4 x 8-bit: index of last qualifying run per a
    S2Run r;
    int idx = 0;
    runs.reset();
    while (runs.next(r)) {
        if (r.a != SENT && r.b != SENT) { hsa |= 1 << r.a; hsb |= 1 << r.b; }
        if (r.a != SENT && (r.b != SENT || !((hsa >> r.a) & 1)))
            last_run = (last_run & ~(0xFFu << (8 * r.a))) | ((uint32_t)idx << (8 * r.a));
        ++idx;
    }
    int outb = 0;
    idx = 0;
    runs.reset();
    while (runs.next(r)) {
        const int tip = r.a == SENT;
        const bool skip = (tip && ((hsb >> r.b) & 1)) || (r.b == SENT && ((hsa >> r.a) & 1));   // s2.cpp:801-818
        if (!skip) {
            const int w = r.b == SENT ? 0 : (((outb >> r.b) & 1) ? r.b + 5 : r.b + 1);          // s2.cpp:820
            const int last = tip ? 0 : (int)(((last_run >> (8 * r.a)) & 0xFFu) == (uint32_t)idx);
            outb |= 1 << r.b;
            sink.record(w, last, tip, r.cnt > 65535u ? 65535u : r.cnt, r.item);
        }
        ++idx;
    }
}

// bytes of one record (sdbg_multi_io.h:93-111)
MGTA_HD uint32_t s2_record_bytes(int tip, uint32_t mult, int words_per_tip) {
    return 2u + (mult > 254u ? 2u : 0u) + (tip ? 4u * (uint32_t)words_per_tip : 0u);
}
MGTA_HD uint16_t s2_record_word(int w, int last, int tip, uint32_t mult) {
    return (uint16_t)(w | (last << 4) | (tip << 5) | ((mult < 255u ? mult : 255u) << 8));
}

// ---- stage 1 mercy classification (s1.cpp:705-829) over saturating (prev,head)/(tail,next)/(head,tail) tables.
// Table = 16 counters of 8 bits (index hi*4+lo, both < 4) saturating at 255; thresholds above 255 are
// clamped by the caller (mercy is only offered for min_count <= 255).
struct Sat16 {
    uint64_t lo, hi;
    MGTA_HD void clear() { lo = hi = 0; }
    MGTA_HD void add(int idx, uint32_t n) {
        uint64_t &w = idx < 8 ? lo : hi;
        const int sh = (idx & 7) * 8;
        uint64_t c = (w >> sh) & 0xFF;
        c = c + n > 255 ? 255 : c + n;
        w = (w & ~(0xFFull << sh)) | (c << sh);
    }
    MGTA_HD uint32_t get(int idx) const { return (uint32_t)(((idx < 8 ? lo : hi) >> ((idx & 7) * 8)) & 0xFF); }
};

struct S1GroupMasks {
    int has_in, has_out, l_has_out, r_has_in;
};

MGTA_HD S1GroupMasks s1_group_masks(const Sat16 &cph, const Sat16 &ctn, const Sat16 &cht, uint32_t m) {
    S1GroupMasks g = {0, 0, 0, 0};
    for (int j = 0; j < 4; ++j)
        for (int x = 0; x < 4; ++x) {
            if (cph.get(x * 4 + j) >= m) g.has_in |= 1 << j;          // count_prev_head[x][j]
            if (ctn.get(j * 4 + x) >= m) g.has_out |= 1 << j;         // count_tail_next[j][x]
            if (cht.get(j * 4 + x) >= m) { g.l_has_out |= 1 << j; g.r_has_in |= 1 << x; }
        }
    return g;
}

// candidates of one item; push(pos, flag) with pos = absolute k-mer position (start_idx + offset);
// kpos = absolute position of the item's (k-1)-mer (payload field), so edge offset = kpos - 1.
template <class Push>
MGTA_HD void s1_mercy_item(const S1GroupMasks &g, bool solid, int head, int tail, int strand, uint64_t kpos,
                           Push &&push) {
    const uint64_t lo = strand == 0 ? kpos - 1 : kpos, ro = strand == 0 ? kpos : kpos - 1;
    if (solid) {
        if (!((g.has_in >> head) & 1)) push(lo, 1 + strand);
        if (!((g.has_out >> tail) & 1)) push(ro, 2 - strand);
    } else {
        if (head != SENT && ((g.l_has_out >> head) & 1)) push(lo, ((g.has_in >> head) & 1) ? 0 : 1 + strand);
        else if (head != SENT && ((g.has_in >> head) & 1)) push(lo, 2 - strand);
        if (tail != SENT && ((g.r_has_in >> tail) & 1)) push(ro, ((g.has_out >> tail) & 1) ? 0 : 2 - strand);
        else if (tail != SENT && ((g.has_out >> tail) & 1)) push(ro, 1 + strand);
    }
}

// ---- the per-read mercy scan (s2_read_mercy_prepare, s2.cpp:170-238) over position flags instead of sorted candidates.
// bit(v, i): flag vector v at k-mer offset i of this read: 0 = a candidate with flag 1 ("no in"), 1 = a candidate with flag 2
// ("no out"), 2 = any candidate ("has solid k-mer"), 3 = is_solid of edge offset i (asked only for i + k < L, and never
// for an offset this scan has set).  set_solid(j): mark edge offset j.  Returns the number of edges marked ("Number mercy").
template <class Bit, class SetSolid>
MGTA_HD unsigned long long mercy_scan_read(int L, int k, Bit &&bit, SetSolid &&set_solid) {
    if (L < k + 1) return 0;
    int first_0_out = 1 << 30, last_0_in = -1;
    bool any = false;
    for (int i = 0; i + k <= L; ++i) {
        if (bit(2, i)) any = true;
        if (bit(1, i) && i < first_0_out) first_0_out = i;
        if (bit(0, i)) last_0_in = i;
    }
    if (!any || last_0_in < first_0_out) return 0;                         // s2.cpp:203-205
    unsigned long long added = 0;
    int last_no_out = -1;
    bool carry = false;                                                    // original is_solid[i - 1]: marks has_solid_kmer[i] too (s2.cpp:216-220)
    for (int i = 0; i + k <= L; ++i) {
        const bool sol = (i + k < L) && bit(3, i);                         // still the original value: this scan only sets offsets < i
        const bool hs = bit(2, i) || sol || carry;
        if (bit(0, i) && last_no_out != -1) {
            for (int j = last_no_out; j < i; ++j) set_solid(j);
            added += (unsigned long long)(i - last_no_out);
        }
        if (hs) last_no_out = -1;
        if (bit(1, i)) last_no_out = i;
        carry = sol;
    }
    return added;
}

}  // namespace mgta
