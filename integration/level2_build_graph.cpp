// integration/level2_build_graph.cpp -- INTEGRATION.md "Level 2" as a program that compiles: the reference's build_graph()
// (src/build_graph.cpp:33-135) with its two cx1.run() calls replaced by the C ABI of include/mgta_cuda.h.
//
// Everything around the hot path stays the UNMODIFIED reference, compiled from /root/reference/src where it lies
// (oracle/Makefile, target level2): OptionsDescription parses the command line, s1_read_input_prepare
// (cx1_read2sdbg_s1.cpp:96-175) loads the read library and the --assist_seq file into its SequencePackage, and
// SdbgWriter (sdbg_multi_io.h:34-199) writes <P>.sdbg.0 and <P>.sdbg_info.  No reference source is changed or copied:
// the package's two arrays are public members, and the records are replayed through SdbgWriter::write -- or, built with
// -DMGTA_HAVE_APPEND_RAW against a tree that carries integration/sdbg_writer_append_raw.patch, appended with one fwrite per delivery.
// (One GPU; for several GPUs run the loop of megagta_b200/csrc/host/buildgraph_b200.cpp gpu_thread() per device.)
//
//   megagta_level2 buildgraph -k 31 -m 2 --host_mem 8e9 --num_cpu_threads 4 --read_lib_file X --output_prefix P
#include <omp.h>
#include <stdio.h>
#include <string.h>

#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cx1_read2sdbg.h"
#include "definitions.h"
#include "options_description.h"
#include "utils.h"

#include "level2_sink.h"
#include "mgta_cuda.h"

static int build_graph_b200(int argc, char **argv) {
    AutoMaxRssRecorder recorder;                                          // prints the reference's "Real: ..." line (utils.h:124)
    OptionsDescription desc;
    read2sdbg_opt_t opt;
    // the option table of build_graph.cpp:38-48 -- the command-line contract
    desc.AddOption("kmer_k", "k", opt.kmer_k, "kmer size");
    desc.AddOption("min_kmer_frequency", "m", opt.kmer_freq_threshold, "min frequency to output an edge");
    desc.AddOption("host_mem", "", opt.host_mem, "Max memory to be used. 90% of the free memory is recommended.");
    desc.AddOption("gpu_mem", "", opt.gpu_mem, "HBM to be used. 0 for 90% of the free device memory.");
    desc.AddOption("num_cpu_threads", "", opt.num_cpu_threads, "number of CPU threads. At least 2.");
    desc.AddOption("num_output_threads", "", opt.num_output_threads, "number of threads for output. Must be less than num_cpu_threads");
    desc.AddOption("read_lib_file", "", opt.read_lib_file, "input read library prefix (from `buildlib`)");
    desc.AddOption("assist_seq", "", opt.assist_seq_file, "input assisting fast[aq] file (FILE_NAME.info should exist), can be gzip'ed.");
    desc.AddOption("output_prefix", "", opt.output_prefix, "output prefix");
    desc.AddOption("mem_flag", "", opt.mem_flag, "accepted for compatibility");
    desc.AddOption("need_mercy", "", opt.need_mercy, "to add mercy edges.");
    try {
        desc.Parse(argc, argv);
        if (opt.read_lib_file == "") throw std::logic_error("No input file!");
        if (opt.num_cpu_threads == 0) opt.num_cpu_threads = omp_get_max_threads();
        if (opt.num_output_threads == 0) opt.num_output_threads = std::max(1, opt.num_cpu_threads / 3);
        if (opt.host_mem == 0) throw std::logic_error("Please specify the host memory!");
        if (opt.num_cpu_threads == 1) throw std::logic_error("Number of CPU threads should be at least 2!");
        if (opt.num_output_threads >= opt.num_cpu_threads) throw std::logic_error("Number of output threads must be less than number of CPU threads!");
    } catch (std::exception &e) {
        std::cerr << e.what() << std::endl << "Usage: sdbg_builder read2sdbg --read_lib_file fastx_file -o out" << std::endl
                  << "Options:" << std::endl << desc << std::endl;
        exit(1);
    }

    // ---- the reference's own loader: reads reversed into a SequencePackage, assist sequences appended (s1.cpp:96-134)
    cx1_read2sdbg::read2sdbg_global_t *g = new cx1_read2sdbg::read2sdbg_global_t();
    g->kmer_k = opt.kmer_k; g->kmer_freq_threshold = 1;                   // 1: no host-side is_solid vector, it lives in HBM
    g->host_mem = opt.host_mem; g->gpu_mem = opt.gpu_mem; g->num_cpu_threads = opt.num_cpu_threads;
    g->num_output_threads = opt.num_output_threads; g->read_lib_file = opt.read_lib_file;
    g->assist_seq_file = opt.assist_seq_file; g->output_prefix = opt.output_prefix; g->mem_flag = opt.mem_flag;
    g->need_mercy = opt.need_mercy;
    cx1_read2sdbg::s1::s1_read_input_prepare(*g);
    SequencePackage &package = g->package;

    // ---- the B200 path behind the C ABI
    mgta_opts mo;
    memset(&mo, 0, sizeof(mo));
    mo.kmer_k = opt.kmer_k; mo.min_count = opt.kmer_freq_threshold; mo.need_mercy = opt.need_mercy && opt.kmer_freq_threshold > 1;
    mo.device = 0; mo.rank = 0; mo.world = 1; mo.hbm_budget_bytes = (int64_t)opt.gpu_mem;
    mgta_ctx *ctx = NULL;
    if (mgta_ctx_create(&mo, &ctx) != 0) { xerr_and_exit("%s\n", mgta_last_error(NULL)); }     // the macro is several statements
    if (mgta_set_reads(ctx, &package.packed_seq[0], package.packed_seq.size(), &package.start_idx_[0], (uint64_t)g->num_reads,
                       (uint64_t)g->num_short_reads, g->max_read_length) != 0) {
        xerr_and_exit("%s\n", mgta_last_error(ctx));
    }

    if (opt.kmer_freq_threshold > 1) {                                    // replaces cx1.run() #1 + s1_post_proc
        std::vector<int64_t> edge_counting(MGTA_NUM_BUCKETS, 0);
        if (mgta_stage1(ctx, &edge_counting[0]) != 0) { xerr_and_exit("%s\n", mgta_last_error(ctx)); }
        long long num_solid_edges = 0;
        for (int i = opt.kmer_freq_threshold; i <= kMaxMulti_t; ++i) num_solid_edges += edge_counting[i];
        xlog("Total number of solid edges: %llu\n", num_solid_edges);
        FILE *counting_file = OpenFileAndCheck((opt.output_prefix + ".counting").c_str(), "w");       // s1.cpp:925-930
        long long acc = 0;
        for (int i = 1; i <= kMaxMulti_t; ++i) { acc += edge_counting[i]; fprintf(counting_file, "%lld %lld\n", (long long)i, acc); }
        fclose(counting_file);
        if (mo.need_mercy) {
            uint64_t num_mercy = 0;
            mgta_get_num_mercy(ctx, &num_mercy);
            xlog("Number mercy: %llu\n", (unsigned long long)num_mercy);                              // s2.cpp:241
        }
    }

    {                                                                     // replaces cx1.run() #2; the writer is the reference's
        SdbgWriter writer;
        writer.set_num_threads(1);
        writer.set_file_prefix(opt.output_prefix);
        writer.set_kmer_size(opt.kmer_k);
        writer.set_num_buckets(MGTA_NUM_BUCKETS);
        writer.init_files();
        Level2Sink sink = {&writer, (2 * opt.kmer_k + 31) / 32, 0};
        int64_t totals[10];
        if (mgta_stage2(ctx, LEVEL2_SINK, &sink, totals) != 0) { xerr_and_exit("%s\n", mgta_last_error(ctx)); }
        xlog("Number of $ A C G T A- C- G- T-:\n");                       // the reference's closing log (s2.cpp:905-915)
        xlog("");
        for (int i = 0; i < 9; ++i) xlog_ext("%lld ", (long long)totals[i]);
        xlog_ext("\n");
        xlog("Total number of edges: %lld\n", (long long)writer.num_edges());
        xlog("Total number of ONEs: %lld\n", (long long)writer.num_last1());
        xlog("Total number of $v edges: %lld\n", (long long)writer.num_tips());
#ifndef MGTA_HAVE_APPEND_RAW                                              // (append_raw does not count W: the totals are the library's)
        for (int i = 0; i < 9; ++i)
            if (writer.num_w(i) != totals[i]) { xerr_and_exit("internal: W totals of the library and of the writer differ\n"); }
#endif
    }                                                                     // ~SdbgWriter closes the files and writes sdbg_info
    mgta_ctx_destroy(ctx);
    delete g;
    return 0;
}

// replay <stream file> <meta file: int64[65536][3]> <k> <out prefix> <buckets per delivery>: a recorded bucket-ordered record
// stream handed to the sink in deliveries (no GPU involved) -- how the tests check the sink and the reference's writer
static int replay(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: replay <stream> <meta> <k> <out_prefix> <buckets_per_delivery>\n"); return 1; }
    const int k = atoi(argv[3]), per = atoi(argv[5]) > 0 ? atoi(argv[5]) : 65536, wpt = (2 * k + 31) / 32;
    std::vector<unsigned char> stream;
    std::vector<int64_t> meta(65536 * 3);
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END); stream.resize(ftell(f)); fseek(f, 0, SEEK_SET);
    if (!stream.empty() && fread(&stream[0], 1, stream.size(), f) != stream.size()) return 1;
    fclose(f);
    f = fopen(argv[2], "rb");
    if (!f || fread(&meta[0], 8, meta.size(), f) != meta.size()) return 1;
    fclose(f);
    SdbgWriter writer;
    writer.set_num_threads(1); writer.set_file_prefix(argv[4]); writer.set_kmer_size(k); writer.set_num_buckets(65536);
    writer.init_files();
    Level2Sink sink = {&writer, wpt, 0};
    size_t at = 0;
    for (int b0 = 0; b0 < 65536; b0 += per) {
        const int b1 = b0 + per < 65536 ? b0 + per : 65536;
        size_t bytes = 0;
        for (int b = b0; b < b1; ++b) bytes += meta[b * 3] * 2 + meta[b * 3 + 2] * 2 + meta[b * 3 + 1] * 4 * wpt;
        if (at + bytes > stream.size()) bytes = stream.size() - at;      // a table that claims more than the stream holds
        const int rc = LEVEL2_SINK(&sink, b0, b1, stream.empty() ? NULL : &stream[at], bytes, &meta[b0 * 3]);
        if (rc) { fprintf(stderr, "sink returned %d\n", rc); return 3; }
        at += bytes;
    }
    return at == stream.size() ? 0 : 4;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "buildgraph") == 0) return build_graph_b200(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "replay") == 0) return replay(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s buildgraph [options of `megagta buildgraph`]\n", argv[0]);
    return 1;
}
