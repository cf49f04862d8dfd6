// seq_tools.cu -- the streaming passes on either side of the graph build.
//
//   mgta_pack_reads   `megagta buildlib`: ASCII bases -> the records of <X>.bin (u32 length, ceil(len / 16) u32 words, first
//                     base in bits 31..30, unused low bits zero).  Replaces SequencePackage::AddSeqToPackedSeq_
//                     (reference sequence_package.h:254-270, character map :67-69) + SequenceManager::WriteBinarySequences
//                     (sequence_manager.cpp:375-410); the FASTA/Q parsing stays on the host (it is gzip / file bound).
//
// One thread per output word: 16 characters of one read, fetched as bytes (reads start at arbitrary byte offsets), mapped
// through a 256-entry table in shared memory; the record index of a word comes from a binary search over the record
// offsets of the batch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/mgta_cuda.h"

namespace {

thread_local std::string g_tools_error;

__global__ void __launch_bounds__(256) k_pack_reads(const unsigned char *__restrict__ bases, const unsigned long long *__restrict__ seq_off,
                                                    const unsigned long long *__restrict__ rec_off, unsigned long long n_reads,
                                                    unsigned long long n_out, uint32_t *__restrict__ out) {
    __shared__ unsigned char map[256];
    for (int i = threadIdx.x; i < 256; i += 256) {
        unsigned char v = 0;                                       // the reference leaves other bytes undefined (uninitialised table)
        if (i == 'C' || i == 'c') v = 1;
        if (i == 'G' || i == 'g' || i == 'N' || i == 'n') v = 2;   // N -> G (sequence_package.h:67-69)
        if (i == 'T' || i == 't') v = 3;
        map[i] = v;
    }
    __syncthreads();
    const unsigned long long j = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= n_out) return;
    unsigned long long lo = 0, hi = n_reads - 1;                   // the record of output word j: last r with rec_off[r] <= j
    while (lo < hi) {
        const unsigned long long mid = (lo + hi + 1) >> 1;
        if (rec_off[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const unsigned long long s = seq_off[lo], len = seq_off[lo + 1] - s, w = j - rec_off[lo];
    if (w == 0) { out[j] = (uint32_t)len; return; }               // the length word of the record
    const unsigned long long c0 = (w - 1) * 16;
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (c0 + i < len) x |= (uint32_t)map[bases[s + c0 + i]] << (30 - 2 * i);
    out[j] = x;
}

}  // namespace

extern "C" const char *mgta_tools_last_error(void) { return g_tools_error.c_str(); }

#define TCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            g_tools_error = buf_;                                                                        \
            cudaFree(d_bases); cudaFree(d_off); cudaFree(d_rec); cudaFree(d_out);                        \
            return MGTA_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

extern "C" int mgta_pack_reads(int device, const char *bases, const uint64_t *seq_off, uint64_t n_reads, uint32_t *out_records,
                               uint64_t out_words) {
    unsigned char *d_bases = nullptr;
    unsigned long long *d_off = nullptr, *d_rec = nullptr;
    uint32_t *d_out = nullptr;
    if (!seq_off || !out_records || (!bases && n_reads && seq_off[n_reads])) { g_tools_error = "pack_reads: null argument"; return MGTA_ERR_ARG; }
    if (n_reads == 0) return MGTA_OK;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        g_tools_error = "pack_reads: no CUDA device (there is no CPU fallback)";
        return MGTA_ERR_CUDA;
    }
    // record offsets in u32 words: 1 length word + ceil(len / 16) data words per read
    std::string rec_bytes((size_t)(n_reads + 1) * 8, '\0');
    unsigned long long *rec = reinterpret_cast<unsigned long long *>(&rec_bytes[0]);
    unsigned long long acc = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        if (seq_off[r + 1] < seq_off[r] || seq_off[r + 1] - seq_off[r] > 0xFFFFFFFFull) { g_tools_error = "pack_reads: bad offsets"; return MGTA_ERR_ARG; }
        rec[r] = acc;
        acc += 1 + (seq_off[r + 1] - seq_off[r] + 15) / 16;
    }
    rec[n_reads] = acc;
    if (acc != out_words) { g_tools_error = "pack_reads: out_words must be the sum of 1 + ceil(len / 16) over the reads"; return MGTA_ERR_ARG; }
    const uint64_t n_bases = seq_off[n_reads] - seq_off[0];
    TCK(cudaSetDevice(device));
    TCK(cudaMalloc(&d_bases, n_bases + 16));
    TCK(cudaMalloc(&d_off, (n_reads + 1) * 8));
    TCK(cudaMalloc(&d_rec, (n_reads + 1) * 8));
    TCK(cudaMalloc(&d_out, acc * 4));
    if (n_bases) TCK(cudaMemcpy(d_bases, bases + seq_off[0], n_bases, cudaMemcpyHostToDevice));
    {
        // offsets relative to the first base of the batch
        std::string rel_bytes((size_t)(n_reads + 1) * 8, '\0');
        unsigned long long *rel = reinterpret_cast<unsigned long long *>(&rel_bytes[0]);
        for (uint64_t r = 0; r <= n_reads; ++r) rel[r] = seq_off[r] - seq_off[0];
        TCK(cudaMemcpy(d_off, rel, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    }
    TCK(cudaMemcpy(d_rec, rec, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    k_pack_reads<<<(unsigned)((acc + 255) / 256), 256>>>(d_bases, d_off, d_rec, n_reads, acc, d_out);
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(out_records, d_out, acc * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_bases); cudaFree(d_off); cudaFree(d_rec); cudaFree(d_out);
    return MGTA_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
//   mgta_find_seeds   `megagta findstart`: every read, both strands, all three frames is translated codon by codon and each
//                     window of aa_k amino acids is looked up in the set of the reference proteins' k-mers.  Replaces the
//                     OpenMP loop of find_start / ProcessSequenceMulti (reference fast_kmer_filter.cpp:117-146,196-218:
//                     per read two std::string copies, three translated strings, one hash lookup per window).
//
// One thread per (read, strand, frame): the protein k-mer (5 bits per residue, two u64) rolls along the frame, one probe
// of an open-addressing table in global memory (L2 resident: the k-mers of a gene's reference alignment) per window.
namespace {

// standard genetic code in the residue codes of ProtKmer::setUp (prot_kmer.h:26-43): ARNDCQEGHI = 0..9, LKMFPSTWYV = 10..19,
// * = 20; codon index = 16 b0 + 4 b1 + b2 with A C G T = 0 1 2 3 (sequence/Codon.C:8-90)
__constant__ unsigned char c_codon[64] = {
    11, 2, 11, 2,   16, 16, 16, 16,  1, 15, 1, 15,   9, 9, 12, 9,      // AAx ACx AGx ATx
    5, 8, 5, 8,     14, 14, 14, 14,  1, 1, 1, 1,     10, 10, 10, 10,   // CAx CCx CGx CTx
    6, 3, 6, 3,     0, 0, 0, 0,      7, 7, 7, 7,     19, 19, 19, 19,   // GAx GCx GGx GTx
    20, 18, 20, 18, 15, 15, 15, 15,  20, 4, 17, 4,   10, 13, 10, 13};  // TAx TCx TGx TTx

struct SeedParams {
    const uint32_t *rec;               // .bin records of the batch, back to back
    const unsigned long long *rec_off; // [n_reads + 1] word offset of every record
    unsigned long long n_reads;
    int aa_k, min_len;
    const unsigned long long *tab;     // [cap][2] keys, empty = (~0, ~0)
    const int *tab_val;                // [cap] index of the model k-mer
    unsigned cap_mask;
    unsigned long long *hit_pos;       // read << 24 | strand << 23 | nucleotide offset
    unsigned *hit_model;
    unsigned long long *n_hits, hits_cap;
};

__device__ __forceinline__ unsigned long long seed_hash(unsigned long long a, unsigned long long b) {
    unsigned long long x = a * 0x9E3779B97F4A7C15ull ^ (b + 0xC2B2AE3D27D4EB4Full);
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}

__global__ void __launch_bounds__(256) k_find_seeds(const SeedParams P) {
    const unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long r = t / 6;
    if (r >= P.n_reads) return;
    const int strand = (int)(t % 6) / 3, frame = (int)(t % 3);
    const uint32_t *rec = P.rec + P.rec_off[r];
    const int len = (int)rec[0];
    if (len < P.min_len) return;
    const uint32_t *w = rec + 1;
    auto base = [&](int j) -> unsigned {                           // base j of the strand this thread reads
        const int q = strand ? len - 1 - j : j;
        const unsigned b = (w[q >> 4] >> (30 - 2 * (q & 15))) & 3u;
        return strand ? 3u - b : b;
    };
    const int n_aa = (len - frame) / 3;
    const int k = P.aa_k;
    // k-mer layout: residues 0..11 in k0 (first residue most significant), 12..23 in k1; only equality matters
    const int n1 = k > 12 ? k - 12 : 0, n0 = k - n1;
    const unsigned long long m0 = n0 * 5 >= 64 ? ~0ull : ((1ull << (n0 * 5)) - 1ull), m1 = n1 ? ((1ull << (n1 * 5)) - 1ull) : 0ull;
    unsigned long long k0 = 0, k1 = 0;
    for (int a = 0; a < n_aa; ++a) {
        const int p = frame + 3 * a;
        const unsigned aa = c_codon[base(p) * 16 + base(p + 1) * 4 + base(p + 2)];
        if (n1) {
            const unsigned long long carry = (k1 >> ((n1 - 1) * 5)) & 31ull;
            k1 = ((k1 << 5) | aa) & m1;
            k0 = ((k0 << 5) | carry) & m0;
        } else {
            k0 = ((k0 << 5) | aa) & m0;
        }
        if (a + 1 < k) continue;
        unsigned slot = (unsigned)seed_hash(k0, k1) & P.cap_mask;
        for (;;) {
            const unsigned long long t0 = P.tab[2ull * slot], t1 = P.tab[2ull * slot + 1];
            if (t0 == k0 && t1 == k1) {
                const unsigned long long at = atomicAdd(P.n_hits, 1ull);
                if (at < P.hits_cap) {
                    P.hit_pos[at] = (r << 24) | ((unsigned long long)strand << 23) | (unsigned long long)(frame + 3 * (a + 1 - k));
                    P.hit_model[at] = (unsigned)P.tab_val[slot];
                }
                break;
            }
            if (t0 == ~0ull && t1 == ~0ull) break;
            slot = (slot + 1) & P.cap_mask;
        }
    }
}

}  // namespace

#undef TCK
#define TCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char buf_[512];                                                                              \
            snprintf(buf_, sizeof(buf_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            g_tools_error = buf_;                                                                        \
            for (void *p_ : to_free) cudaFree(p_);                                                       \
            return MGTA_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

// model_kmers: [n_model][2] u64 in the layout above (mgta_seed_pack builds it from residue codes); records: the .bin records
// of n_reads reads (n_words u32); hits: hit_pos / hit_model of capacity hits_cap; *n_hits keeps counting past the capacity
extern "C" int mgta_find_seeds(int device, const uint64_t *model_kmers, uint64_t n_model, int aa_k, const uint32_t *records,
                               uint64_t n_words, uint64_t n_reads, int min_len, uint64_t *hit_pos, uint32_t *hit_model,
                               uint64_t hits_cap, uint64_t *n_hits) {
    std::vector<void *> to_free;
    if (!n_hits || aa_k < 1 || aa_k > 24 || (n_model && !model_kmers) || (n_reads && !records)) { g_tools_error = "find_seeds: bad argument"; return MGTA_ERR_ARG; }
    *n_hits = 0;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        g_tools_error = "find_seeds: no CUDA device (there is no CPU fallback)";
        return MGTA_ERR_CUDA;
    }
    if (n_reads == 0 || n_model == 0) return MGTA_OK;
    // open-addressing table of the model k-mers (first occurrence wins, like HashSetST::insert -> insert_unique)
    unsigned cap = 1024;
    while ((uint64_t)cap < 2 * n_model) cap <<= 1;
    std::vector<unsigned long long> tab((size_t)cap * 2, ~0ull);
    std::vector<int> val(cap, -1);
    auto hash = [](unsigned long long a, unsigned long long b) {
        unsigned long long x = a * 0x9E3779B97F4A7C15ull ^ (b + 0xC2B2AE3D27D4EB4Full);
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        return x;
    };
    for (uint64_t i = 0; i < n_model; ++i) {
        const unsigned long long a = model_kmers[2 * i], b = model_kmers[2 * i + 1];
        unsigned slot = (unsigned)hash(a, b) & (cap - 1);
        for (;;) {
            if (tab[2ull * slot] == a && tab[2ull * slot + 1] == b) break;
            if (val[slot] < 0) { tab[2ull * slot] = a; tab[2ull * slot + 1] = b; val[slot] = (int)i; break; }
            slot = (slot + 1) & (cap - 1);
        }
    }
    // record offsets of the batch
    std::vector<unsigned long long> off(n_reads + 1);
    uint64_t p = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        if (p >= n_words) { g_tools_error = "find_seeds: the records end before n_reads reads"; return MGTA_ERR_ARG; }
        off[r] = p;
        p += 1 + ((uint64_t)records[p] + 15) / 16;
    }
    off[n_reads] = p;
    if (p > n_words) { g_tools_error = "find_seeds: the last record is truncated"; return MGTA_ERR_ARG; }
    uint32_t *d_rec = nullptr; unsigned long long *d_off = nullptr, *d_tab = nullptr, *d_pos = nullptr, *d_n = nullptr;
    int *d_val = nullptr; unsigned *d_model = nullptr;
    TCK(cudaSetDevice(device));
    TCK(cudaMalloc(&d_rec, (p + 4) * 4)); to_free.push_back(d_rec);
    TCK(cudaMalloc(&d_off, (n_reads + 1) * 8)); to_free.push_back(d_off);
    TCK(cudaMalloc(&d_tab, (size_t)cap * 16)); to_free.push_back(d_tab);
    TCK(cudaMalloc(&d_val, (size_t)cap * 4)); to_free.push_back(d_val);
    TCK(cudaMalloc(&d_pos, (hits_cap + 1) * 8)); to_free.push_back(d_pos);
    TCK(cudaMalloc(&d_model, (hits_cap + 1) * 4)); to_free.push_back(d_model);
    TCK(cudaMalloc(&d_n, 8)); to_free.push_back(d_n);
    TCK(cudaMemcpy(d_rec, records, p * 4, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(d_off, off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(d_tab, tab.data(), (size_t)cap * 16, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(d_val, val.data(), (size_t)cap * 4, cudaMemcpyHostToDevice));
    TCK(cudaMemset(d_n, 0, 8));
    SeedParams P;
    memset(&P, 0, sizeof(P));
    P.rec = d_rec; P.rec_off = d_off; P.n_reads = n_reads; P.aa_k = aa_k; P.min_len = min_len; P.tab = d_tab; P.tab_val = d_val;
    P.cap_mask = cap - 1; P.hit_pos = d_pos; P.hit_model = d_model; P.n_hits = d_n; P.hits_cap = hits_cap;
    const unsigned long long threads = n_reads * 6;
    k_find_seeds<<<(unsigned)((threads + 255) / 256), 256>>>(P);
    TCK(cudaGetLastError());
    unsigned long long n = 0;
    TCK(cudaMemcpy(&n, d_n, 8, cudaMemcpyDeviceToHost));
    *n_hits = n;
    const uint64_t take = n < hits_cap ? n : hits_cap;
    if (take && hit_pos) TCK(cudaMemcpy(hit_pos, d_pos, take * 8, cudaMemcpyDeviceToHost));
    if (take && hit_model) TCK(cudaMemcpy(hit_model, d_model, take * 4, cudaMemcpyDeviceToHost));
    for (void *q : to_free) cudaFree(q);
    return MGTA_OK;
}
