// TEST INFRASTRUCTURE ONLY (oracle/): a multiplexer over the reference's own sub-programs (declared at
// /root/reference/src/megagta.cpp:8-15): `buildlib` / `buildgraph` are the path under test, `denovo` / `findstart` / `search`
// are the stages that consume its files (downstream acceptance).
#include <stdio.h>
#include <string.h>
// `sdbgdump` looks inside SuccinctDBG / RankAndSelect after the reference's own LoadFromMultiFile (succinct_dbg.cpp:595-723)
// has filled them: the golden arrays for the device-side SdBG load + rank/select build.  Layout-neutral.
#define private public
#include "succinct_dbg.h"
#undef private
#include "cx1_read2sdbg.h"
int build_lib(int argc, char **argv);
int build_graph(int argc, char **argv);
int main_assemble(int argc, char **argv);
int find_start(int argc, char **argv);
int search(int argc, char **argv);
static void put(FILE *f, const char *name, const void *p, size_t bytes) {
    char nm[24];
    memset(nm, 0, sizeof(nm));
    strncpy(nm, name, sizeof(nm) - 1);
    unsigned long long n = bytes;
    fwrite(nm, 1, sizeof(nm), f);
    fwrite(&n, 8, 1, f);
    if (bytes) fwrite(p, 1, bytes, f);
}

// sdbgdump <prefix> <need_multiplicity 0|1> <out>: sections {char name[24]; u64 bytes; data}
static int sdbg_dump(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: sdbgdump <prefix> <need_mult> <out>\n"); return 1; }
    const bool need_mul = atoi(argv[2]) != 0;
    SuccinctDBG g;
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    g.LoadFromMultiFile(argv[1], need_mul);
    gettimeofday(&t1, NULL);
    fprintf(stderr, "load_seconds %.6f\n", (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_usec - t0.tv_usec));
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 1;
    char nm[64];
    const int64_t size = g.size;
    const size_t ww = (size + 15) / 16, wl = (size + 63) / 64;
    int64_t hdr[4] = {size, g.kmer_k, g.num_tip_nodes_, g.uint32_per_tip_nodes_};
    put(f, "hdr", hdr, sizeof(hdr));
    put(f, "f", g.f_, sizeof(g.f_));
    put(f, "rank_f", g.rank_f_, sizeof(g.rank_f_));
    put(f, "w", g.w_, ww * 8);
    put(f, "last", g.last_, wl * 8);
    put(f, "is_tip", g.is_tip_, wl * 8);
    put(f, "invalid", g.invalid_, wl * 8);
    put(f, "tip_seq", g.tip_node_seq_, (size_t)g.num_tip_nodes_ * g.uint32_per_tip_nodes_ * 4);
    if (g.edge_multi_) put(f, "edge_multi", g.edge_multi_, size);
    if (g.edge_large_multi_) put(f, "edge_large_multi", g.edge_large_multi_, (size_t)size * 2);
    if (g.is_multi_1_) put(f, "is_multi_1", g.is_multi_1_, wl * 8);
    if (need_mul && g.edge_multi_) {                               // the khash of large multiplicities as sorted (edge, mult) pairs
        std::vector<std::pair<long long, long long> > v;
        for (khint_t k = kh_begin(g.large_multi_h_); k != kh_end(g.large_multi_h_); ++k)
            if (kh_exist(g.large_multi_h_, k)) v.push_back(std::make_pair((long long)kh_key(g.large_multi_h_, k), (long long)kh_value(g.large_multi_h_, k)));
        std::sort(v.begin(), v.end());
        put(f, "large_multi", v.empty() ? NULL : &v[0], v.size() * 16);
    }
    {   // rank/select of W (rank_and_select.h:33-330)
        RankAndSelect4Bits &r = g.rs_w_;
        const int64_t ni = (size + 255) / 256 + 1, nmaj = (size + 65535) / 65536 + 1;
        put(f, "w_freq", r.char_frequency, sizeof(r.char_frequency));
        for (int c = 0; c < 9; ++c) {
            snprintf(nm, sizeof(nm), "w_major_%d", c); put(f, nm, r.occ_value_explicit_major_[c], nmaj * 8);
            snprintf(nm, sizeof(nm), "w_minor_%d", c); put(f, nm, r.occ_value_explicit_minor_[c], ni * 2);
            const int64_t ns = (r.char_frequency[c] + 255) / 256 + 1;
            snprintf(nm, sizeof(nm), "w_sel_%d", c); put(f, nm, r.rank_to_interval_explicit_[c], ns * 4);
        }
    }
    {   // rank/select of last (rank_and_select.h:400-)
        const int64_t ni = (size + 255) / 256 + 1, nmaj = (size + 65535) / 65536 + 1;
        put(f, "last_ones", &g.rs_last_.total_num_ones, 8);
        put(f, "last_major", g.rs_last_.occ_value_explicit_major_, nmaj * 8);
        put(f, "last_minor", g.rs_last_.occ_value_explicit_minor_, ni * 2);
        put(f, "last_sel", g.rs_last_.rank_to_interval_explicit_, ((g.rs_last_.total_num_ones + 255) / 256 + 1) * 4);
        put(f, "tip_ones", &g.rs_is_tip_.total_num_ones, 8);
        put(f, "tip_major", g.rs_is_tip_.occ_value_explicit_major_, nmaj * 8);
        put(f, "tip_minor", g.rs_is_tip_.occ_value_explicit_minor_, ni * 2);
    }
    fclose(f);
    return 0;
}


// readsdump <read_lib_prefix> <assist_seq or ""> <out>: the read set as the reference holds it in memory after its own
// s1_read_input_prepare (cx1_read2sdbg_s1.cpp:96-175: ReadBinaryLibs with is_reverse = true, then --assist_seq appended) --
// the two arrays the C ABI's mgta_set_reads takes.  Pins the oracle's numpy loader and the driver's C++ loader.
static int reads_dump(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: readsdump <read_lib_prefix> <assist_seq|\"\"> <out>\n"); return 1; }
    cx1_read2sdbg::read2sdbg_global_t *g = new cx1_read2sdbg::read2sdbg_global_t();
    g->kmer_k = 21; g->kmer_freq_threshold = 1; g->host_mem = (int64_t)1 << 40; g->gpu_mem = 0; g->num_cpu_threads = 2;
    g->num_output_threads = 1; g->mem_flag = 1; g->need_mercy = false;
    g->read_lib_file = argv[1]; g->assist_seq_file = argv[2]; g->output_prefix = argv[3];
    cx1_read2sdbg::s1::s1_read_input_prepare(*g);
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 1;
    int64_t hdr[4] = {g->num_reads, g->num_short_reads, g->max_read_length, (int64_t)g->package.base_size()};
    put(f, "hdr", hdr, sizeof(hdr));
    put(f, "packed_seq", &g->package.packed_seq[0], g->package.packed_seq.size() * 4);
    put(f, "start_idx", &g->package.start_idx_[0], g->package.start_idx_.size() * 8);
    fclose(f);
    return 0;
}


int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "readsdump") == 0) return reads_dump(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "sdbgdump") == 0) return sdbg_dump(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "buildlib") == 0) return build_lib(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "buildgraph") == 0) return build_graph(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "denovo") == 0) return main_assemble(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "findstart") == 0) return find_start(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "search") == 0) return search(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s buildlib|buildgraph|denovo|findstart|search [options]\n", argv[0]);
    return 1;
}
