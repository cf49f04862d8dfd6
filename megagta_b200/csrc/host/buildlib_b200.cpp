// buildlib_b200.cpp -- host side of `megagta_b200 buildlib <read_lib_file> <out_prefix>`, the drop-in for `megagta buildlib`
// (reference src/build_read_lib.cpp:30-41 -> ReadAndWriteMultipleLibs, read_lib_functions-inl.h:119-226).
//
// Same input (a library list: per library one free-text line, then `pe f1 f2` | `se f` | `interleaved f`; FASTA / FASTQ,
// optionally gzip'ed, parsed with kseq's record rules, kseq.h:168-207), same output (<P>.bin: per read u32 length + packed
// words; <P>.lib_info: totals + two lines per library, read_lib_functions-inl.h:216-225).  The host parses the text (file
// and gzip bound); the 2-bit packing of every batch runs on the GPU (mgta_pack_reads).  No CPU packing path exists.
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <fstream>
#include <string>
#include <vector>

#include "../../../include/mgta_cuda.h"
#include "fastx_reader.h"

namespace {

[[noreturn]] void die_lib(const std::string &msg) {
    fprintf(stderr, "[ERROR] %s\n", msg.c_str());
    exit(1);
}

struct LibInfo {
    std::string metadata;
    long long from, to;
    int max_read_len;
    bool is_pe;
};

struct Batch {
    std::string bases;
    std::vector<uint64_t> off{0};
    void add(const std::string &s) { bases += s; off.push_back(bases.size()); }
    size_t reads() const { return off.size() - 1; }
    void clear() { bases.clear(); off.assign(1, 0); }
};

void flush(Batch &b, FILE *bin, int device) {
    if (b.reads() == 0) return;
    uint64_t words = 0;
    for (size_t r = 0; r < b.reads(); ++r) words += 1 + (b.off[r + 1] - b.off[r] + 15) / 16;
    std::vector<uint32_t> rec(words);
    if (mgta_pack_reads(device, b.bases.data(), b.off.data(), b.reads(), rec.data(), words) != 0)
        die_lib(std::string("mgta_pack_reads: ") + mgta_tools_last_error());
    if (fwrite(rec.data(), 4, words, bin) != words) die_lib("short write to the .bin file");
    b.clear();
}

}  // namespace

int build_lib_b200(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "Usage %s <read_lib_file> <out_prefix>\n", argv[0]);
        exit(1);
    }
    const std::string lib_file = argv[1], out_prefix = argv[2];
    std::ifstream cfg(lib_file);
    if (!cfg.is_open()) die_lib("File to open read_lib file: " + lib_file);
    FILE *bin = fopen((out_prefix + ".bin").c_str(), "wb");
    if (!bin) die_lib("cannot open " + out_prefix + ".bin: " + strerror(errno));
    const int device = 0;
    const size_t kBatchReads = 1 << 22, kBatchBases = 1 << 28;     // the reference's batch shape; it does not show in the files
    std::vector<LibInfo> libs;
    long long total_reads = 0, total_bases = 0;
    std::string metadata, type, f1, f2, seq1, seq2, rest;
    Batch batch;
    while (std::getline(cfg, metadata)) {
        if (!(cfg >> type)) die_lib("read_lib file: library type expected after \"" + metadata + "\"");
        const bool pe = type == "pe";
        if (pe) { if (!(cfg >> f1 >> f2)) die_lib("read_lib file: pe needs two files"); }
        else if (type == "se" || type == "interleaved") { if (!(cfg >> f1)) die_lib("read_lib file: file name expected"); }
        else {
            fprintf(stderr, "Cannot identify read library type %s\n", type.c_str());
            die_lib("Valid types: pe, se, interleaved");
        }
        const long long start = total_reads;
        int max_len = 0;
        {
            FastxReader r1(f1);
            FastxReader *r2 = pe ? new FastxReader(f2) : nullptr;
            for (;;) {
                if (!r1.next(seq1)) {
                    if (r2 && r2->next(seq2)) die_lib("paired files differ in length: " + f1 + " ended first");
                    break;
                }
                batch.add(seq1);
                max_len = std::max<int>(max_len, (int)seq1.size());
                ++total_reads; total_bases += (long long)seq1.size();
                if (r2) {
                    if (!r2->next(seq2)) die_lib("paired files differ in length: " + f2 + " ended first");
                    batch.add(seq2);
                    max_len = std::max<int>(max_len, (int)seq2.size());
                    ++total_reads; total_bases += (long long)seq2.size();
                }
                if (batch.reads() >= kBatchReads || batch.bases.size() >= kBatchBases) flush(batch, bin, device);
            }
            delete r2;
        }
        flush(batch, bin, device);
        if (type != "se" && (total_reads - start) % 2 != 0) {
            fprintf(stderr, "PE library number of reads is odd: %lld!\n", total_reads - start);
            die_lib("File(s): " + metadata);
        }
        fprintf(stderr, "[B200] Lib %zu (%s): %s, %lld reads, %d max length\n", libs.size(), metadata.c_str(), type.c_str(), total_reads - start, max_len);
        libs.push_back({metadata, start, total_reads - 1, max_len, type != "se"});
        std::getline(cfg, rest);                                   // the rest of the file-name line
    }
    fclose(bin);
    FILE *info = fopen((out_prefix + ".lib_info").c_str(), "w");
    if (!info) die_lib("cannot open " + out_prefix + ".lib_info: " + strerror(errno));
    fprintf(info, "%zu %zu\n", (size_t)total_bases, (size_t)total_reads);
    for (auto &l : libs) {
        fprintf(info, "%s\n", l.metadata.c_str());
        fprintf(info, "%lld %lld %d %s\n", l.from, l.to, l.max_read_len, l.is_pe ? "pe" : "se");
    }
    fclose(info);
    return 0;
}
