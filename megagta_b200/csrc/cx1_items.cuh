// cx1_items.cuh -- which sort items one base position of one read contributes, and their keys.
//
// Stage 1 (reference s1_lv0_calc_bucket_size s1.cpp:177-229, s1_lv1_fill_offset :408-513,
// s1_extract_subtstr_ :515-596): the (k-1)-mer S at read offset p with its neighbours
// prev,head | S | tail,next.  Read ends (p == 0, p == L-k+1) go in on BOTH strands, interior
// positions on the smaller of (S, rc(S)), palindromes by `head <= 3 - tail` (s1.cpp:482-495).
//   key  = S (or rc S) | zero pad | head<<3|tail in the low 6 bits of the last word (s1.cpp:575-588)
//   value (ours, the reference's is an lv1 offset): kpos<<8 | strand<<6 | prev<<3 | next, where
//          kpos = absolute base index of S = start_idx[read] + p (the (k+1)-mer head S tail starts at
//          kpos-1), or S1_NO_EDGE for assist reads (never marked solid, no mercy: s1.cpp:757,785).
//
// Stage 2 (reference s2_lv0_calc_bucket_size s2.cpp:252-315, s2_lv1_fill_offset :475-584,
// s2_lv2_extract_substr_ :586-677): a solid edge e = R[o..o+k] with r = rc(e) contributes
//   solid  : (b=e[0], S=e[1..k-1], a=e[k])          and, unless e is a palindrome, the same from r
//   left $ : (b=$,   S=e[0..k-2], a=e[k-1]) ; (b=r[1], S=r[2..k], a=$)      if o==0 or !solid(o-1)
//   right $: (b=e[1], S=e[2..k],  a=$)      ; (b=$,   S=r[0..k-2], a=r[k-1]) if o==L-k-1 or !solid(o+1)
//   key = S a | zero pad | (a!=$)<<3 | b in the low 4 bits of the last word (s2.cpp:639-641,668-670)
//
// __host__ __device__ so the CPU logic test can drive the same code against the oracle.
#pragma once
#include "kmer_ops.cuh"

namespace mgta {

constexpr int SENT = 4;                                 // kSentinelValue, cx1_read2sdbg.h:71
constexpr uint64_t S1_NO_EDGE = (1ull << 40) - 1;       // payload marker: item never marks a solid edge

MGTA_HD int comp_char(int c) { return c == SENT ? SENT : 3 - c; }

MGTA_HD int key_words_s1(int k) { return (2 * (k - 1) + 6 + 31) / 32; }   // s1.cpp:246
MGTA_HD int key_words_s2(int k) { return (2 * k + 4 + 31) / 32; }         // s2.cpp:331

// words: staged read words; q: char offset of the position inside `words`; g: absolute base index
// of the position; p: offset inside its read of length L.  Caller guarantees L >= k+1, p <= L-k+1.
// emit(key[W], value64)
template <int W, class Emit>
MGTA_HD void s1_position(const uint32_t *words, uint32_t q, uint64_t g, int p, int L, int k, bool short_read,
                         Emit &&emit) {
    uint32_t S[W], R[W];
    load_chars<W>(words, q, k - 1, S);
    revcomp<W>(S, k - 1, R);
    const int head = p > 0 ? char_at(words, q - 1) : SENT;
    const int prev = p > 1 ? char_at(words, q - 2) : SENT;
    const int tail = p + k - 1 < L ? char_at(words, q + k - 1) : SENT;
    const int next = p + k < L ? char_at(words, q + k) : SENT;
    const bool ends = (p == 0) || (p == L - k + 1);
    const int c = ends ? 0 : cmp_words<W>(S, R);
    const bool tie_fw = head <= 3 - tail;                               // s1.cpp:486
    const bool fw = ends || c < 0 || (c == 0 && tie_fw);
    const bool rv = ends || c > 0 || (c == 0 && !tie_fw);
    const uint64_t edge = short_read ? g : S1_NO_EDGE;
    if (fw) {
        S[W - 1] |= (uint32_t)((head << 3) | tail);
        emit(S, (edge << 8) | (0u << 6) | (uint64_t)((prev << 3) | next));
    }
    if (rv) {
        R[W - 1] |= (uint32_t)((comp_char(tail) << 3) | comp_char(head));
        emit(R, (edge << 8) | (1u << 6) | (uint64_t)((comp_char(next) << 3) | comp_char(prev)));
    }
}

// The <= 6 stage-2 items of one solid edge E (k+1 chars, zero padded, W = key_words_s2(k) words) with
// R = rc(E), pal = (E == R).  `left` / `right`: also emit the $-items of the edge's first / last k-mer.
// emit(key[W])
template <int W, class Emit>
MGTA_HD void s2_edge_items(const uint32_t (&E)[W], const uint32_t (&R)[W], bool pal, bool left, bool right, int k,
                           Emit &&emit) {
    auto put = [&](const uint32_t(&X)[W], int c, bool has_a) {
        uint32_t Y[W];
        sub_chars<W>(X, c, has_a ? k : k - 1, Y);
        const int b = c ? (int)((X[0] >> (32 - 2 * c)) & 3u) : SENT;
        Y[W - 1] |= (uint32_t)(((has_a ? 1 : 0) << 3) | b);
        emit(Y);
    };
    if (left) {
        put(E, 0, true);
        if (!pal) put(R, 2, false);
    }
    put(E, 1, true);
    if (!pal) put(R, 1, true);
    if (right) {
        put(E, 2, false);
        if (!pal) put(R, 0, true);
    }
}

// Caller guarantees L >= k+1, o < L-k and solid(o).  emit(key[W])
template <int W, class Emit>
MGTA_HD void s2_position(const uint32_t *words, uint32_t q, int o, int L, int k, bool solid_prev, bool solid_next,
                         Emit &&emit) {
    uint32_t E[W], R[W];
    load_chars<W>(words, q, k + 1, E);
    revcomp<W>(E, k + 1, R);
    const bool pal = cmp_words<W>(E, R) == 0;
    s2_edge_items<W>(E, R, pal, (o == 0) || !solid_prev, (o == L - k - 1) || !solid_next, k, emit);
}

// ---- edge-centric path (v2).  An EDGE is a (k+1)-mer; its canonical form is min(e, rc(e)) as zero padded
// words.  Stage 1 without mercy only needs the multiset of canonical edges (s1.cpp:744-760 count
// (head, S, tail) groups, whose strand rule s1.cpp:482-495 is one particular canonical form: any
// other gives the same counts and the same solid occurrences); stage 2's records are a function of
// {(oriented edge, number of solid occurrences)}.
MGTA_HD int edge_words(int k) { return (2 * (k + 1) + 31) / 32; }

template <int WE>
MGTA_HD void canonical_edge(const uint32_t *words, uint32_t q, int k, uint32_t (&key)[WE]) {
    uint32_t E[WE], R[WE];
    load_chars<WE>(words, q, k + 1, E);
    revcomp<WE>(E, k + 1, R);
    const bool fw = cmp_words<WE>(E, R) <= 0;
#pragma unroll
    for (int w = 0; w < WE; ++w) key[w] = fw ? E[w] : R[w];
}

// all stage-2 items of a canonical edge (both orientations, $-items unconditionally: the emission
// rules s2.cpp:801-818 drop a $-item whenever a real edge covers it, which is exactly the case in
// which the reference would not have produced it at every occurrence).  W = key_words_s2(k) >= WE.
// tips = false: only the (up to two) real items b S a -- the node pass below supplies the $-items that survive.
template <int W, int WE, class Emit>
MGTA_HD void s2_items_of_edge(const uint32_t (&key)[WE], int k, Emit &&emit, bool tips = true) {
    uint32_t E[W], R[W];
#pragma unroll
    for (int w = 0; w < W; ++w) E[w] = w < WE ? key[w < WE ? w : 0] : 0u;
    revcomp<W>(E, k + 1, R);
    s2_edge_items<W>(E, R, cmp_words<W>(E, R) == 0, tips, tips, k, emit);
}

// ---- node pass.  output_() (s2.cpp:801-818) drops a $-item whenever a real edge covers it, so the
// only $-items that reach the records are those of TIP k-mers: oriented k-mers X with solid edges leaving but none
// entering (left tip: item ($, X)) or entering but none leaving (right tip: item (X, $)), with multiplicity = the summed
// multiplicities of the edges on the other side.  Both are functions of two weights per CANONICAL k-mer c:
//     out(c) = sum of mult over oriented solid edges that start with c,   in(c) = ... that end in c,
// and by reverse-complement symmetry out(rc c) = in(c), in(rc c) = out(c).  A canonical edge E contributes to the
// canonical forms of its first and last k-mer (its reverse complement contributes the same again, so it is not visited;
// a palindromic edge starts with X and ends in rc X, one and the same contribution).
MGTA_HD int kmer_words(int k) { return (2 * k + 31) / 32; }

// emit(c[WE], dir): canonical k-mer zero padded to WE words (the first kmer_words(k) hold it), dir 0: out(c) += mult,
// dir 1: in(c) += mult.  A palindromic k-mer (k even) gets both kinds on one entry; its in and out are equal by
// symmetry, so it is never a tip (s2_tip_items is not called for it).
template <int WE, class Emit>
MGTA_HD void node_ops_of_edge(const uint32_t (&E)[WE], int k, Emit &&emit) {
    uint32_t R[WE], X[WE], XR[WE];
    revcomp<WE>(E, k + 1, R);
    const bool pal = cmp_words<WE>(E, R) == 0;
    sub_chars<WE>(E, 0, k, X);                          // first k-mer; its reverse complement is the last k-mer of R
    sub_chars<WE>(R, 1, k, XR);
    {
        const bool fw = cmp_words<WE>(X, XR) <= 0;
        if (fw) emit(X, 0); else emit(XR, 1);
    }
    if (!pal) {
        sub_chars<WE>(E, 1, k, X);                      // last k-mer; reverse complement = first k-mer of R
        sub_chars<WE>(R, 0, k, XR);
        const bool fw = cmp_words<WE>(X, XR) <= 0;
        if (fw) emit(X, 1); else emit(XR, 0);
    }
}

// The two stage-2 items of a tip k-mer C (zero padded to W = key_words_s2(k) words, RC its reverse complement, C != RC).
// no_in: nothing enters C (so nothing leaves RC): C is a left tip, RC a right tip; else the other way round.
//   left tip  L: ($, S = L[0..k-2], a = L[k-1])   key = L | (a != $) << 3 | $
//   right tip T: (b = T[0], S = T[1..k-1], a = $)  key = T[1..k-1] | b
template <int W, class Emit>
MGTA_HD void s2_tip_items(const uint32_t (&C)[W], const uint32_t (&RC)[W], bool no_in, int k, Emit &&emit) {
    uint32_t Y[W];
#pragma unroll
    for (int w = 0; w < W; ++w) Y[w] = no_in ? C[w] : RC[w];
    Y[W - 1] |= (uint32_t)((1 << 3) | SENT);
    emit(Y);
    uint32_t T[W];
#pragma unroll
    for (int w = 0; w < W; ++w) T[w] = no_in ? RC[w] : C[w];
    sub_chars<W>(T, 1, k - 1, Y);
    Y[W - 1] |= (T[0] >> 30) & 3u;
    emit(Y);
}

}  // namespace mgta
