/*
 * oracle/cx1_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's CX1 reads -> SdBG construction
 * (`megagta buildgraph`).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the CHECKER.  The product path (libmgta_cuda.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against golden hashes that
 * were produced by the unmodified reference binary (oracle/_ref/megagta_ref, built by
 * oracle/Makefile from /root/reference/src) with tests/golden/make_golden.py; SURVEY.md Appendix C
 * lists the same hashes for the shared cases; tests/test_oracle_fuzz.py compares it with the reference
 * binary run live on 52 seeded random read sets (k = 9 ... 127, min-count 1 ... 3, mercy, assist reads).
 *
 * It deliberately does NOT follow the reference's schedule (lv1 passes, int32 delta offsets,
 * kt_dfor work stealing: /root/reference/src/cx1.h:443-623) -- only its observable results:
 *   - stage-1 items and keys       cx1_read2sdbg_s1.cpp:177-229, 408-513, 515-596
 *   - stage-1 group counting       cx1_read2sdbg_s1.cpp:671-830, 905-930
 *   - mercy edges                  cx1_read2sdbg_s2.cpp:106-250
 *   - stage-2 items and keys       cx1_read2sdbg_s2.cpp:252-315, 475-584, 586-677
 *   - stage-2 group emission       cx1_read2sdbg_s2.cpp:742-835
 *   - record format                sdbg_multi_io.h:83-112
 *   - sort order                   lv2_cpu_sort.h:87-151 (ascending by the entire key)
 *
 * Reads are given as the reference holds them in memory: 2-bit, bit-contiguous, REVERSED (not
 * complemented) reads (sequence_package.h:247-252,341-367), start_idx in bases.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NB 65536
#define SENT 4 /* kSentinelValue, cx1_read2sdbg.h:71 */

typedef struct {
    const uint32_t *seq;
    const uint64_t *start; /* n_reads + 1 */
    int64_t n_reads, n_short;
    int max_len, k, m;
} cx1o_in;

static inline int base_at(const cx1o_in *in, uint64_t pos) { /* sequence_package.h:129-132 */
    return (in->seq[pos >> 4] >> ((15 - (pos & 15)) * 2)) & 3;
}
static inline int cm(int c) { return c == SENT ? SENT : 3 - c; }

/* write chars c[0..n) MSB-first into zero-initialised words (packed_reads.h:44-176 semantics) */
static inline void put_chars(uint32_t *w, const uint8_t *c, int n) {
    for (int i = 0; i < n; ++i) w[i >> 4] |= (uint32_t)c[i] << ((15 - (i & 15)) * 2);
}

int cx1o_cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
static int g_w; /* words compared by qsort */
static int cmp_item(const void *a, const void *b) {
    const uint32_t *x = a, *y = b;
    for (int i = 0; i < g_w; ++i)
        if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}

/* ---- growable u64 / byte vectors ---- */
typedef struct { uint64_t *p; int64_t n, cap; } vec64;
static void v64_push(vec64 *v, uint64_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = realloc(v->p, v->cap * 8); }
    v->p[v->n++] = x;
}
typedef struct { uint8_t *p; int64_t n, cap; } vec8;
static void v8_put(vec8 *v, const void *src, int64_t n) {
    if (v->n + n > v->cap) { while (v->n + n > v->cap) v->cap = v->cap ? v->cap * 2 : 4096; v->p = realloc(v->p, v->cap); }
    memcpy(v->p + v->n, src, n); v->n += n;
}

int cx1o_words_s1(int k) { return (2 * (k - 1) + 6 + 31) / 32; } /* s1.cpp:246 */
int cx1o_words_s2(int k) { return (2 * k + 4 + 31) / 32; }       /* s2.cpp:331 */

/* ===================================================================================== stage 1 */

/* One stage-1 item: (k-1)-mer at position p of read rid on `strand` (s1.cpp:515-596).
 * item layout: W key words, then 2 words value = (rid<<?)... we keep (rid, off, strand, prev, next)
 * explicitly: val0 = rid (low 32), val1 = rid_hi<<24 | p<<8 | strand<<6 | prev<<3 | next. */
static void s1_make(const cx1o_in *in, int64_t rid, int p, int strand, uint32_t *item, int W) {
    int k = in->k, L = (int)(in->start[rid + 1] - in->start[rid]);
    uint64_t s0 = in->start[rid];
    uint8_t S[128];
    for (int i = 0; i < k - 1; ++i) S[i] = base_at(in, s0 + p + i);
    int head = p > 0 ? base_at(in, s0 + p - 1) : SENT;
    int prev = p > 1 ? base_at(in, s0 + p - 2) : SENT;
    int tail = p + k - 1 < L ? base_at(in, s0 + p + k - 1) : SENT;
    int next = p + k < L ? base_at(in, s0 + p + k) : SENT;
    memset(item, 0, 4 * (W + 2));
    if (strand == 0) {
        put_chars(item, S, k - 1);
        item[W - 1] |= (head << 3) | tail;                 /* s1.cpp:575-580 */
    } else {
        uint8_t R[128];
        for (int i = 0; i < k - 1; ++i) R[i] = 3 - S[k - 2 - i];
        put_chars(item, R, k - 1);
        item[W - 1] |= (cm(tail) << 3) | cm(head);         /* s1.cpp:583-588 */
        int t = prev; prev = cm(next); next = cm(t);
    }
    item[W] = (uint32_t)rid;
    item[W + 1] = ((uint32_t)(rid >> 32) << 24) | ((uint32_t)p << 8) | (strand << 6) | (prev << 3) | next;
}

/* strands of the (k-1)-mer at position p: returns bitmask 1=fwd 2=rc (s1.cpp:459-507) */
static int s1_strands(const cx1o_in *in, int64_t rid, int p) {
    int k = in->k, L = (int)(in->start[rid + 1] - in->start[rid]);
    uint64_t s0 = in->start[rid];
    if (p == 0 || p == L - k + 1) return 3;
    for (int i = 0; i < k - 1; ++i) {
        int f = base_at(in, s0 + p + i), r = 3 - base_at(in, s0 + p + k - 2 - i);
        if (f < r) return 1;
        if (f > r) return 2;
    }
    int head = base_at(in, s0 + p - 1), tail = base_at(in, s0 + p + k - 1);
    return head <= 3 - tail ? 1 : 2;                        /* s1.cpp:482-495 */
}

/* 65536-bin histogram of stage-1 items (s1.cpp:177-229) */
void cx1o_s1_hist(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int k, int64_t *hist) {
    cx1o_in in = {seq, start, n_reads, n_reads, 0, k, 2};
    int W = cx1o_words_s1(k);
    uint32_t item[16];
    memset(hist, 0, NB * sizeof(int64_t));
    for (int64_t r = 0; r < n_reads; ++r) {
        int L = (int)(start[r + 1] - start[r]);
        if (L < k + 1) continue;
        for (int p = 0; p <= L - k + 1; ++p) {
            int st = s1_strands(&in, r, p);
            for (int s = 0; s < 2; ++s)
                if (st >> s & 1) { s1_make(&in, r, p, s, item, W); hist[item[0] >> 16]++; }
        }
    }
}

/* Stage 1.  is_solid: zero-initialised, ceil(n_short*(max_len-k)/8) bytes, bit i at byte i/8 bit
 * i%8 (atomic_bit_vector.h:58-60).  edge_counting: int64[65536].  mercy: optional list of packed
 * candidates ((start_idx+off)<<2|flag, s1.cpp:764), returned unsorted-then-sorted ascending.
 * Returns 0, or -1 on allocation failure. */
int cx1o_stage1(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short,
                int max_len, int k, int m, uint8_t *is_solid, int64_t *edge_counting,
                int need_mercy, uint64_t **mercy_out, int64_t *n_mercy_out) {
    cx1o_in in = {seq, start, n_reads, n_short, max_len, k, m};
    int W = cx1o_words_s1(k), IW = W + 2;
    int64_t nk1 = max_len - k;
    int64_t n_items = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int L = (int)(start[r + 1] - start[r]);
        if (L >= k + 1) n_items += L - k + 4;              /* L-k+4 items per read */
    }
    uint32_t *items = malloc((size_t)(n_items ? n_items : 1) * IW * 4);
    if (!items) return -1;
    int64_t n = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int L = (int)(start[r + 1] - start[r]);
        if (L < k + 1) continue;
        for (int p = 0; p <= L - k + 1; ++p) {
            int st = s1_strands(&in, r, p);
            for (int s = 0; s < 2; ++s)
                if (st >> s & 1) s1_make(&in, r, p, s, items + (n++) * IW, W);
        }
    }
    g_w = W;
    qsort(items, n, IW * 4, cmp_item);                     /* lv2_cpu_sort.h:133-151: by entire key */
    memset(edge_counting, 0, NB * sizeof(int64_t));
    vec64 cand = {0, 0, 0};
    int full = (k - 1) / 16, rem = (k - 1) % 16;
    for (int64_t i = 0, e; i < n; i = e) {
        /* group of equal (k-1)-mer: IsDiffKMinusOneMer, s1.cpp:59-80 */
        const uint32_t *a = items + i * IW;
        for (e = i + 1; e < n; ++e) {
            const uint32_t *b = items + e * IW;
            int diff = 0;
            for (int w = 0; w < full; ++w) if (a[w] != b[w]) { diff = 1; break; }
            if (!diff && rem && (a[full] >> (16 - rem) * 2) != (b[full] >> (16 - rem) * 2)) diff = 1;
            if (diff) break;
        }
        int cph[5][5] = {{0}}, ctn[5][5] = {{0}}, cht[64] = {0};
        for (int64_t j = i; j < e; ++j) {
            const uint32_t *it = items + j * IW;
            int ht = it[W - 1] & 63, pn = it[W + 1] & 63;
            cph[pn >> 3][ht >> 3]++; ctn[ht & 7][pn & 7]++; cht[ht]++;
        }
        int has_in = 0, has_out = 0, l_has_out = 0, r_has_in = 0;   /* s1.cpp:705-733 */
        for (int j = 0; j < 4; ++j)
            for (int x = 0; x < 4; ++x) {
                if (cph[x][j] >= m) has_in |= 1 << j;
                if (ctn[j][x] >= m) has_out |= 1 << j;
                if (cht[(j << 3) | x] >= m) { l_has_out |= 1 << j; r_has_in |= 1 << x; }
            }
        int64_t j = i;
        while (j < e) {                                    /* s1.cpp:735-829 */
            const uint32_t *it0 = items + j * IW;
            int ht = it0[W - 1] & 63, head = ht >> 3, tail = ht & 7, c = cht[ht];
            if (head != SENT && tail != SENT) edge_counting[c < 65535 ? c : 65535]++;
            int solid = head != SENT && tail != SENT && c >= m;
            for (int q = 0; q < c; ++q, ++j) {
                const uint32_t *it = items + j * IW;
                int64_t rid = (int64_t)it[W] | ((int64_t)(it[W + 1] >> 24) << 32);
                int p = (it[W + 1] >> 8) & 0xffff, strand = (it[W + 1] >> 6) & 1;
                int off = p - 1, lo = strand == 0 ? off : off + 1, ro = strand == 0 ? off + 1 : off;
                if (rid >= n_short) continue;              /* s1.cpp:757,785 */
                uint64_t s0 = start[rid];
                if (solid) {
                    int64_t bit = nk1 * rid + off;
                    is_solid[bit >> 3] |= 1u << (bit & 7);
                    if (need_mercy) {
                        if (!(has_in >> head & 1)) v64_push(&cand, ((s0 + lo) << 2) | (1 + strand));
                        if (!(has_out >> tail & 1)) v64_push(&cand, ((s0 + ro) << 2) | (2 - strand));
                    }
                } else if (need_mercy) {
                    if (l_has_out >> head & 1) v64_push(&cand, ((s0 + lo) << 2) | ((has_in >> head & 1) ? 0 : 1 + strand));
                    else if (has_in >> head & 1) v64_push(&cand, ((s0 + lo) << 2) | (2 - strand));
                    if (r_has_in >> tail & 1) v64_push(&cand, ((s0 + ro) << 2) | ((has_out >> tail & 1) ? 0 : 2 - strand));
                    else if (has_out >> tail & 1) v64_push(&cand, ((s0 + ro) << 2) | (1 + strand));
                }
            }
        }
    }
    free(items);
    if (mercy_out) {
        if (cand.n > 1) qsort(cand.p, cand.n, 8, cx1o_cmp_u64);   /* ascending, s2.cpp:138 */
        *mercy_out = cand.p; *n_mercy_out = cand.n;
    } else free(cand.p);
    return 0;
}
void cx1o_free(void *p) { free(p); }

/* Mercy edges (s2.cpp:106-250).  cand sorted ascending.  Returns num_mercy. */
int64_t cx1o_mercy(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short,
                   int max_len, int k, uint8_t *is_solid, const uint64_t *cand, int64_t n_cand) {
    (void)seq;
    int64_t nk1 = max_len - k, num_mercy = 0, i = 0, rid = 0;
    uint8_t *no_in = malloc(max_len + 2), *no_out = malloc(max_len + 2), *hs = malloc(max_len + 2);
    while (i < n_cand) {
        uint64_t pos = cand[i] >> 2;
        while (!(start[rid] <= pos && pos < start[rid + 1])) ++rid;   /* package.get_id */
        int first_0_out = max_len + 1, last_0_in = -1;
        memset(no_in, 0, max_len + 2); memset(no_out, 0, max_len + 2); memset(hs, 0, max_len + 2);
        while (i < n_cand && (cand[i] >> 2) < start[rid + 1]) {
            int off = (int)((cand[i] >> 2) - start[rid]), fl = cand[i] & 3;
            if (fl == 2) { no_out[off] = 1; if (off < first_0_out) first_0_out = off; }
            else if (fl == 1) { no_in[off] = 1; if (off > last_0_in) last_0_in = off; }
            hs[off] = 1; ++i;
        }
        if (last_0_in < first_0_out) continue;
        int L = (int)(start[rid + 1] - start[rid]), last_no_out = -1;
        for (int o = 0; o + k < L; ++o) {
            int64_t bit = rid * nk1 + o;
            if (is_solid[bit >> 3] >> (bit & 7) & 1) hs[o] = hs[o + 1] = 1;
        }
        for (int o = 0; o + k <= L; ++o) {
            if (no_in[o] && last_no_out != -1) {
                for (int j = last_no_out; j < o; ++j) { int64_t bit = rid * nk1 + j; is_solid[bit >> 3] |= 1u << (bit & 7); }
                num_mercy += o - last_no_out;
            }
            if (hs[o]) last_no_out = -1;
            if (no_out[o]) last_no_out = o;
        }
    }
    (void)n_reads; (void)n_short;
    free(no_in); free(no_out); free(hs);
    return num_mercy;
}

/* ===================================================================================== stage 2 */

static inline int solid_at(const cx1o_in *in, const uint8_t *is_solid, int64_t rid, int o) {
    if (in->m == 1 || rid >= in->n_short) return 1;        /* s2.cpp:276,529 */
    int64_t bit = (int64_t)(in->max_len - in->k) * rid + o;
    return is_solid[bit >> 3] >> (bit & 7) & 1;
}

/* key of item (b, S, a): S = c[0..k-1) ; a == SENT leaves the a-slot zero (s2.cpp:586-677) */
static void s2_make(const uint8_t *S, int a, int b, int k, uint32_t *item, int W) {
    memset(item, 0, 4 * W);
    put_chars(item, S, k - 1);
    if (a != SENT) item[(k - 1) >> 4] |= (uint32_t)a << ((15 - ((k - 1) & 15)) * 2);
    item[W - 1] |= ((a != SENT) << 3) | b;                 /* s2.cpp:639-641,668-670 */
}

/* enumerate the stage-2 items of one read into out (or only count them into hist) */
static int64_t s2_items_of_read(const cx1o_in *in, const uint8_t *is_solid, int64_t rid,
                                uint32_t *out, int W, int64_t *hist) {
    int k = in->k, L = (int)(in->start[rid + 1] - in->start[rid]);
    if (L < k + 1) return 0;
    uint64_t s0 = in->start[rid];
    int64_t n = 0;
    uint8_t e[130], r[130];
    uint32_t tmp[16];
#define EMIT(Sptr, a, b) do { uint32_t *dst = out ? out + (n * W) : tmp; s2_make((Sptr), (a), (b), k, dst, W); \
                              if (hist) { hist[dst[0] >> 16]++; }                                            \
                              ++n; } while (0)
    for (int o = 0; o < L - k; ++o) {
        if (!solid_at(in, is_solid, rid, o)) continue;
        for (int i = 0; i <= k; ++i) e[i] = base_at(in, s0 + o + i);
        for (int i = 0; i <= k; ++i) r[i] = 3 - e[k - i];
        int pal = memcmp(e, r, k + 1) == 0;
        if (o == 0 || !solid_at(in, is_solid, rid, o - 1)) {       /* left $  (s2.cpp:540-548) */
            EMIT(e, e[k - 1], SENT);
            if (!pal) EMIT(r + 2, SENT, r[1]);
        }
        EMIT(e + 1, e[k], e[0]);                                    /* solid   (s2.cpp:550-557) */
        if (!pal) EMIT(r + 1, r[k], r[0]);
        if (o == L - k - 1 || !solid_at(in, is_solid, rid, o + 1)) { /* right $ (s2.cpp:559-567) */
            EMIT(e + 2, SENT, e[1]);
            if (!pal) EMIT(r, r[k - 1], SENT);
        }
    }
#undef EMIT
    return n;
}

/* 65536-bin histogram of stage-2 items (s2.cpp:252-315) */
void cx1o_s2_hist(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short,
                  int max_len, int k, int m, const uint8_t *is_solid, int64_t *hist) {
    cx1o_in in = {seq, start, n_reads, n_short, max_len, k, m};
    memset(hist, 0, NB * sizeof(int64_t));
    for (int64_t r = 0; r < n_reads; ++r) s2_items_of_read(&in, is_solid, r, NULL, cx1o_words_s2(k), hist);
}

/* Stage 2.  Outputs: *stream (malloc'd, bucket-ordered concatenation of every bucket's records),
 * meta[b*3+{0,1,2}] = num_items, num_tips, num_large_mul (sdbg_multi_io.h:178-185),
 * totals[0..8] = num_w[0..8], totals[9] = num_last1. */
int cx1o_stage2(const uint32_t *seq, const uint64_t *start, int64_t n_reads, int64_t n_short,
                int max_len, int k, int m, const uint8_t *is_solid,
                uint8_t **stream, int64_t *stream_bytes, int64_t *meta, int64_t *totals) {
    cx1o_in in = {seq, start, n_reads, n_short, max_len, k, m};
    int W = cx1o_words_s2(k), wpt = (2 * k + 31) / 32;
    int64_t cap = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int L = (int)(start[r + 1] - start[r]);
        if (L >= k + 1) cap += 2 * (int64_t)(L - k) + 4;
    }
    /* upper bound 2(L-k)+4 holds per maximal solid run; count exactly instead */
    int64_t n = 0;
    for (int64_t r = 0; r < n_reads; ++r) n += s2_items_of_read(&in, is_solid, r, NULL, W, NULL);
    (void)cap;
    uint32_t *items = malloc((size_t)(n ? n : 1) * W * 4);
    if (!items) return -1;
    int64_t q = 0;
    for (int64_t r = 0; r < n_reads; ++r) q += s2_items_of_read(&in, is_solid, r, items + q * W, W, NULL);
    g_w = W;
    qsort(items, n, W * 4, cmp_item);
    memset(meta, 0, NB * 3 * sizeof(int64_t));
    memset(totals, 0, 10 * sizeof(int64_t));
    vec8 out = {0, 0, 0};
    int full = (k - 1) / 16, rem = (k - 1) % 16;
    int aw = (k - 1) >> 4, ash = (15 - ((k - 1) & 15)) * 2;
    for (int64_t i = 0, e; i < n; i = e) {
        const uint32_t *g0 = items + i * W;
        for (e = i + 1; e < n; ++e) {                       /* s2.cpp:52-73 */
            const uint32_t *b = items + e * W;
            int diff = 0;
            for (int w = 0; w < full; ++w) if (g0[w] != b[w]) { diff = 1; break; }
            if (!diff && rem && (g0[full] >> (16 - rem) * 2) != (b[full] >> (16 - rem) * 2)) diff = 1;
            if (diff) break;
        }
        int hsa = 0, hsb = 0, outb = 0;
        int64_t last_a[4] = {-1, -1, -1, -1};
#define A_OF(it) (((it)[W - 1] >> 3 & 1) ? (int)(((it)[aw] >> ash) & 3) : SENT)   /* s2.cpp:83-94 */
#define B_OF(it) ((int)((it)[W - 1] & 7))                                         /* s2.cpp:96-98 */
        for (int64_t j = i; j < e; ++j) {                   /* s2.cpp:766-780 */
            const uint32_t *it = items + j * W;
            int a = A_OF(it), b = B_OF(it);
            if (a != SENT && b != SENT) { hsa |= 1 << a; hsb |= 1 << b; }
            if (a != SENT && (b != SENT || !(hsa >> a & 1))) last_a[a] = j;
        }
        for (int64_t j = i, j2; j < e; j = j2) {            /* s2.cpp:782-834 */
            const uint32_t *it = items + j * W;
            int a = A_OF(it), b = B_OF(it);
            for (j2 = j + 1; j2 < e; ++j2) {
                const uint32_t *nx = items + j2 * W;
                if (A_OF(nx) != a || B_OF(nx) != b) break;
            }
            int64_t cnt = j2 - j; if (cnt > 65535) cnt = 65535;
            int tip = 0;
            if (a == SENT) { if (hsb >> b & 1) continue; tip = 1; }
            if (b == SENT) { if (hsa >> a & 1) continue; }
            int w = b == SENT ? 0 : ((outb >> b & 1) ? b + 5 : b + 1);
            int last = a == SENT ? 0 : (last_a[a] == j2 - 1);
            outb |= 1 << b;
            int bucket = it[0] >> 16;
            uint16_t rec = (uint16_t)(w | (last << 4) | (tip << 5) | ((cnt < 255 ? cnt : 255) << 8));  /* sdbg_multi_io.h:93 */
            v8_put(&out, &rec, 2);
            meta[bucket * 3 + 0]++; totals[w]++; totals[9] += last;
            if (cnt > 254) { uint16_t mm = (uint16_t)cnt; v8_put(&out, &mm, 2); meta[bucket * 3 + 2]++; }
            if (tip) { v8_put(&out, it, 4 * wpt); meta[bucket * 3 + 1]++; }
        }
#undef A_OF
#undef B_OF
    }
    free(items);
    *stream = out.p; *stream_bytes = out.n;
    return 0;
}
