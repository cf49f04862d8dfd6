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
