// findstart_b200.cpp -- host side of `megagta_b200 findstart <ref_seq> <reads.bin> <k_size> [num_threads] [contigs.fa]`, the
// drop-in for `megagta findstart` (reference src/fast_kmer_filter.cpp:49-218).
//
// The reference's protein k-mers of an aligned reference (.faa; model columns only, ProtKmerGenerator with model_only = true,
// prot_kmer_generator.h:33-137) are collected on the host -- a few hundred thousand at most -- and every read of the packed
// read file (plus, optionally, contigs from a FASTA / FASTQ file) is scanned on the GPU (mgta_find_seeds: both strands, three
// frames, one table probe per amino-acid window).  Output: the reference's seed lines (fast_kmer_filter.cpp:186-188), one
// per distinct nucleotide k-mer, here in sorted order (the reference shuffles them, fast_kmer_filter.cpp:183).
#include <ctype.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../../include/mgta_cuda.h"
#include "fastx_reader.h"
#include "prot_kmers.h"

namespace {

using namespace mgta_host;

[[noreturn]] void die_fs(const std::string &msg) {
    fprintf(stderr, "[ERROR] %s\n", msg.c_str());
    exit(1);
}

struct Hit {
    std::string nucl;
    unsigned model;
    bool operator<(const Hit &o) const { return nucl < o.nucl; }
};

// scans the records of one batch and appends the seeds it yields
void scan_batch(int device, const std::vector<ModelKmer> &model, const std::vector<uint64_t> &keys, int kmer_size,
                const std::vector<uint32_t> &rec, uint64_t n_reads, std::vector<Hit> &hits) {
    if (n_reads == 0) return;
    uint64_t cap = 1 << 22, n = 0;
    std::vector<uint64_t> pos;
    std::vector<uint32_t> mod;
    for (;;) {
        pos.resize(cap); mod.resize(cap);
        if (mgta_find_seeds(device, keys.data(), model.size(), kmer_size / 3, rec.data(), rec.size(), n_reads, kmer_size, pos.data(), mod.data(), cap, &n) != 0)
            die_fs(std::string("mgta_find_seeds: ") + mgta_tools_last_error());
        if (n <= cap) break;
        cap = n + n / 8 + 1024;                                    // counted past the end: once more with room
    }
    std::vector<uint64_t> off(n_reads + 1);
    uint64_t p = 0;
    for (uint64_t r = 0; r < n_reads; ++r) { off[r] = p; p += 1 + ((uint64_t)rec[p] + 15) / 16; }
    std::string s;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t r = pos[i] >> 24;
        const int strand = (int)((pos[i] >> 23) & 1), at = (int)(pos[i] & 0x7FFFFF);
        const uint32_t *w = rec.data() + off[r] + 1;
        const int len = (int)rec[off[r]];
        const int take = std::min(kmer_size, len - at);            // std::string::substr clamps at the end of the read
        s.resize(take);
        for (int j = 0; j < take; ++j) {
            const int q = strand ? len - 1 - (at + j) : at + j;
            const unsigned b = (w[q >> 4] >> (30 - 2 * (q & 15))) & 3u;
            s[j] = "ACGT"[strand ? 3 - b : b];
        }
        hits.push_back({s, mod[i]});
    }
}

}  // namespace

int find_start_b200(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "Usage: %s <ref_seq> <read.lib> <k_size> [num_threads=0] [contigs.fa]\n", argv[0]);
        exit(1);
    }
    const std::string ref_file = argv[1], bin_file = argv[2];
    const int kmer_size = atoi(argv[3]);
    if (kmer_size < 3 || kmer_size / 3 > 24) die_fs("K-mer size cannot be larger than 24 amino acids (k_size / 3)");
    for (const std::string &f : {ref_file, bin_file}) {
        FILE *t = fopen(f.c_str(), "rb");
        if (!t) { fprintf(stderr, "File %s doesn't exist\n", f.c_str()); exit(1); }
        fclose(t);
    }
    const int device = 0, aa_k = kmer_size / 3;
    // ---- the k-mers of the reference proteins
    std::vector<ModelKmer> model;
    {
        FastxReader fr(ref_file);
        std::string seq;
        while (fr.next(seq)) model_kmers_of(seq, aa_k, model);
    }
    std::vector<uint64_t> keys(model.size() * 2);
    for (size_t i = 0; i < model.size(); ++i) { keys[2 * i] = model[i].w[0]; keys[2 * i + 1] = model[i].w[1]; }
    {
        std::vector<std::pair<uint64_t, uint64_t>> u(model.size());
        for (size_t i = 0; i < model.size(); ++i) u[i] = {model[i].w[0], model[i].w[1]};
        std::sort(u.begin(), u.end());
        fprintf(stderr, "[B200] reference kmer set size: %zu\n", (size_t)(std::unique(u.begin(), u.end()) - u.begin()));
    }
    std::vector<Hit> hits;
    // ---- the packed reads, in batches of <= 2^22 reads / 2^28 bases (the reference's batch shape)
    {
        gzFile gz = gzopen(bin_file.c_str(), "rb");
        if (!gz) die_fs("cannot open " + bin_file);
        gzbuffer(gz, 1 << 22);
        std::vector<uint32_t> rec;
        uint64_t n_reads = 0, bases = 0;
        for (;;) {
            uint32_t len;
            const int got = gzread(gz, &len, 4);
            if (got == 4) {
                const uint32_t nw = (len + 15) / 16;
                const size_t at = rec.size();
                rec.resize(at + 1 + nw);
                rec[at] = len;
                if (nw && gzread(gz, rec.data() + at + 1, nw * 4) != (int)(nw * 4)) die_fs("truncated record in " + bin_file);
                ++n_reads; bases += len;
            } else if (got != 0) {
                die_fs("truncated record header in " + bin_file);
            }
            if (got == 0 || n_reads >= (1u << 22) || bases >= (1u << 28)) {
                if (n_reads) fprintf(stderr, "[B200] Processing %llu reads\n", (unsigned long long)n_reads);
                scan_batch(device, model, keys, kmer_size, rec, n_reads, hits);
                rec.clear(); n_reads = 0; bases = 0;
                if (got == 0) break;
            }
        }
        gzclose(gz);
    }
    // ---- optional contigs (fast[aq]): packed on the device, then the same scan
    if (argc > 5) {
        FastxReader fr(argv[5]);
        std::string seq, all;
        std::vector<uint64_t> off{0};
        auto flush = [&]() {
            const uint64_t n = off.size() - 1;
            if (!n) return;
            uint64_t words = 0;
            for (uint64_t r = 0; r < n; ++r) words += 1 + (off[r + 1] - off[r] + 15) / 16;
            std::vector<uint32_t> rec(words);
            if (mgta_pack_reads(device, all.data(), off.data(), n, rec.data(), words) != 0) die_fs(std::string("mgta_pack_reads: ") + mgta_tools_last_error());
            fprintf(stderr, "[B200] Processing %llu contigs\n", (unsigned long long)n);
            scan_batch(device, model, keys, kmer_size, rec, n, hits);
            all.clear(); off.assign(1, 0);
        };
        while (fr.next(seq)) {
            if (seq.size() >= (1u << 23)) die_fs("a contig of 8 Mbp or more does not fit the seed encoding");
            all += seq; off.push_back(all.size());
            if (off.size() - 1 >= (1u << 22) || all.size() >= (1u << 28)) flush();
        }
        flush();
    }
    // ---- one line per distinct nucleotide k-mer (fast_kmer_filter.cpp:181-188)
    std::sort(hits.begin(), hits.end());
    hits.erase(std::unique(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.nucl == b.nucl; }), hits.end());
    for (const Hit &h : hits)
        printf("dump_gene_name\tdump_seq_name\tdump\t%s\ttrue\t%d\t%s\t%d\n", h.nucl.c_str(), 1, unpack_key(model[h.model].w, aa_k).c_str(),
               model[h.model].model_pos);
    return 0;
}
