// buildgraph_b200.cpp -- host-side drop-in for `megagta buildgraph` (reference src/build_graph.cpp:33-135).
//
// Same command line (option table of build_graph.cpp:38-48 with the defaults of cx1_read2sdbg.h:48-59), same input
// (<X>.lib_info + <X>.bin written by `megagta buildlib`, read_lib_functions-inl.h:216-261, sequence_manager.cpp:186-213)
// and same output files (<P>.sdbg.<i>, <P>.sdbg_info: sdbg_multi_io.h:83-112,154-198; <P>.counting: s1.cpp:925-930), so
// the unchanged `megagta denovo / findstart / search` stages load the result as-is.  Everything between "reads in host
// memory" and "record bytes in host memory" runs on the GPU through the C ABI of include/mgta_cuda.h; this file only
// parses options, loads and reverses the packed reads, and writes files.  There is no CPU fallback.
//
//   megagta_b200 buildgraph -k 31 -m 2 --host_mem 8e9 --read_lib_file X --output_prefix P [--num_cpu_threads T ...]
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <chrono>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>      // device count, streams, pinned host memory: plumbing only, every kernel sits behind the C ABI
#include <nccl.h>

#include "../../../include/mgta_cuda.h"

namespace {

struct Options {                       // read2sdbg_opt_t, cx1_read2sdbg.h:36-59
    int kmer_k = 21;
    int min_count = 2;
    double host_mem = 0;
    double gpu_mem = 0;
    int num_cpu_threads = 0;
    int num_output_threads = 0;
    std::string read_lib_file;
    std::string assist_seq_file;
    std::string output_prefix = "out";
    int mem_flag = 1;
    bool need_mercy = false;
};

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "[ERROR] %s\n", msg.c_str());
    exit(1);
}

void usage_and_exit(const char *why) {           // build_graph.cpp:77-84
    fprintf(stderr, "%s\n", why);
    fputs("Usage: sdbg_builder read2sdbg --read_lib_file fastx_file -o out\nOptions:\n"
                    "  -k, --kmer_k arg                kmer size\n"
                    "  -m, --min_kmer_frequency arg    min frequency to output an edge\n"
                    "      --host_mem arg              Max memory to be used. 90% of the free memory is recommended.\n"
                    "      --gpu_mem arg               gpu memory to be used. 0 for auto detect.\n"
                    "      --num_cpu_threads arg       number of CPU threads. At least 2.\n"
                    "      --num_output_threads arg    number of threads for output. Must be less than num_cpu_threads\n"
                    "      --read_lib_file arg         input read library prefix (from `buildlib`)\n"
                    "      --assist_seq arg            input assisting fast[aq] file (FILE_NAME.info should exist), can be gzip'ed.\n"
                    "      --output_prefix arg         output prefix\n"
                    "      --mem_flag arg              memory options (accepted for compatibility; HBM is budgeted by --gpu_mem)\n"
                    "      --need_mercy                to add mercy edges.\n", stderr);
    exit(1);
}

Options parse(int argc, char **argv) {
    Options o;
    static const struct option longopts[] = {
        {"kmer_k", required_argument, nullptr, 'k'},          {"min_kmer_frequency", required_argument, nullptr, 'm'},
        {"host_mem", required_argument, nullptr, 1},          {"gpu_mem", required_argument, nullptr, 2},
        {"num_cpu_threads", required_argument, nullptr, 3},   {"num_output_threads", required_argument, nullptr, 4},
        {"read_lib_file", required_argument, nullptr, 5},     {"assist_seq", required_argument, nullptr, 6},
        {"output_prefix", required_argument, nullptr, 7},     {"mem_flag", required_argument, nullptr, 8},
        {"need_mercy", no_argument, nullptr, 9},              {nullptr, 0, nullptr, 0}};
    opterr = 0;
    int ch;
    while ((ch = getopt_long(argc, argv, "k:m:", longopts, nullptr)) != -1) {
        switch (ch) {
            case 'k': o.kmer_k = atoi(optarg); break;
            case 'm': o.min_count = atoi(optarg); break;
            case 1: o.host_mem = atof(optarg); break;
            case 2: o.gpu_mem = atof(optarg); break;
            case 3: o.num_cpu_threads = atoi(optarg); break;
            case 4: o.num_output_threads = atoi(optarg); break;
            case 5: o.read_lib_file = optarg; break;
            case 6: o.assist_seq_file = optarg; break;
            case 7: o.output_prefix = optarg; break;
            case 8: o.mem_flag = atoi(optarg); break;
            case 9: o.need_mercy = true; break;
            default: usage_and_exit("uknown option");           // options_description.cpp:69-70 (sic)
        }
    }
    // the checks of build_graph.cpp:53-75, same messages
    if (o.read_lib_file.empty()) usage_and_exit("No input file!");
    if (o.num_cpu_threads == 0) o.num_cpu_threads = omp_get_max_threads();
    if (o.num_output_threads == 0) o.num_output_threads = std::max(1, o.num_cpu_threads / 3);
    if (o.host_mem == 0) usage_and_exit("Please specify the host memory!");
    if (o.num_cpu_threads == 1) usage_and_exit("Number of CPU threads should be at least 2!");
    if (o.num_output_threads >= o.num_cpu_threads) usage_and_exit("Number of output threads must be less than number of CPU threads!");
    return o;
}

// ---- read library -> reversed, bit-contiguous packed reads (what ReadBinaryLibs(..., is_reverse = true) builds:
// read_lib_functions-inl.h:233-261, SequencePackage::AppendRevSeq sequence_package.h:247-252,341-367)
struct Reads {
    uint32_t *seq = nullptr;           // n_words (+ slack), zero initialised; pinned when `pinned`
    bool pinned = false;
    uint64_t n_words = 0;
    std::vector<uint64_t> start;       // n_reads + 1
    uint64_t n_reads = 0;
    int max_len = 0;
    void alloc(uint64_t words, bool want_pinned) {
        seq = nullptr; pinned = false;
        if (want_pinned && cudaHostAlloc((void **)&seq, (words + 64) * 4, cudaHostAllocDefault) == cudaSuccess) {
            pinned = true;
            memset(seq, 0, (words + 64) * 4);
        } else {
            cudaGetLastError();
            seq = (uint32_t *)calloc(words + 64, 4);
        }
    }
    void release() {
        if (pinned) cudaFreeHost(seq); else free(seq);
        seq = nullptr;
    }
};

// the 16 bases of a packed word in reverse order (2-bit groups swapped end to end)
inline uint32_t reverse_bases(uint32_t x) {
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}

// The records of <prefix>.bin in memory.  A plain file is mapped (the records are read where the page cache holds them);
// a gzip'ed one is inflated in large blocks -- zlib decides which it is, as in the reference (gzread, sequence_manager.cpp:186-196).
struct BinFile {
    const uint32_t *words = nullptr;
    size_t n_words = 0, tail_bytes = 0;        // tail_bytes: bytes after the last whole word (a truncated file)
    void *map = nullptr;
    size_t map_len = 0;
    std::vector<uint32_t> inflated;
    void open(const std::string &path) {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) die("cannot open " + path);
        if (gzdirect(gz)) {                                          // not compressed
            gzclose(gz);
            const int fd = ::open(path.c_str(), O_RDONLY);
            struct stat sb;
            if (fd < 0 || fstat(fd, &sb) != 0) die("cannot open " + path + ": " + strerror(errno));
            map_len = (size_t)sb.st_size;
            if (map_len) {
                map = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
                if (map == MAP_FAILED) die("cannot map " + path + ": " + strerror(errno));
                madvise(map, map_len, MADV_WILLNEED);
            }
            ::close(fd);
            words = (const uint32_t *)map; n_words = map_len / 4; tail_bytes = map_len % 4;
            return;
        }
        gzbuffer(gz, 1 << 22);
        size_t bytes = 0;
        for (;;) {
            const size_t block = (size_t)64 << 20;
            inflated.resize((bytes + block + 3) / 4);
            const int got = gzread(gz, (char *)inflated.data() + bytes, (unsigned)block);
            if (got < 0) die("cannot inflate " + path);
            bytes += (size_t)got;
            if ((size_t)got < block) break;
        }
        gzclose(gz);
        words = inflated.data(); n_words = bytes / 4; tail_bytes = bytes % 4;
    }
    void close() {
        if (map && map_len) munmap(map, map_len);
        map = nullptr;
        std::vector<uint32_t>().swap(inflated);
    }
};

Reads load_read_lib(const std::string &prefix, int threads, bool want_pinned) {
    Reads R;
    long long total_bases = 0, num_reads = 0;
    {
        FILE *f = fopen((prefix + ".lib_info").c_str(), "r");
        if (!f) die("cannot open " + prefix + ".lib_info: " + strerror(errno));
        if (fscanf(f, "%lld %lld", &total_bases, &num_reads) != 2) die("bad first line in " + prefix + ".lib_info");
        fclose(f);
    }
    BinFile bin;
    bin.open(prefix + ".bin");
    // pass 1: where every record lies (u32 len, ceil(len/16) words, forward orientation)
    std::vector<uint64_t> rec_off;
    rec_off.reserve((size_t)num_reads);
    R.start.reserve((size_t)num_reads + 1);
    R.start.push_back(0);
    uint64_t bases = 0;
    for (size_t at = 0; at < bin.n_words;) {
        const uint32_t len = bin.words[at];
        const size_t nw = ((size_t)len + 15) / 16;
        if (at + 1 + nw > bin.n_words) die("truncated record in " + prefix + ".bin");
        rec_off.push_back(at + 1);
        bases += len;
        R.start.push_back(bases);
        R.max_len = std::max<int>(R.max_len, (int)len);
        at += 1 + nw;
    }
    if (bin.tail_bytes) die("truncated record header in " + prefix + ".bin");
    R.n_reads = rec_off.size();
    if ((long long)R.n_reads != num_reads || (long long)bases != total_bases)
        die("read library " + prefix + ": .bin holds " + std::to_string(R.n_reads) + " reads / " + std::to_string(bases) +
            " bases, .lib_info says " + std::to_string(num_reads) + " / " + std::to_string(total_bases));
    R.n_words = bases / 16 + 1;
    R.alloc(R.n_words, want_pinned);
    if (!R.seq) die("out of host memory for the packed reads");
    // pass 2: every read reversed (not complemented) into its bit range, a word at a time: the record's words in reverse
    // order with their bases reversed are the reversed read behind `pad` empty slots; they stream through a 64-bit
    // accumulator that is aligned with the destination.  Only the first and the last destination word can be shared with the
    // neighbouring reads (atomic OR into the zeroed buffer); the words in between belong to this read alone.
#pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
    for (long long r = 0; r < (long long)R.n_reads; ++r) {
        const uint32_t *w = bin.words + rec_off[(size_t)r];
        const uint64_t s = R.start[(size_t)r];
        const int L = (int)(R.start[(size_t)r + 1] - s);
        if (L == 0) continue;
        const int nw = (L + 15) / 16, pad = nw * 16 - L;
        uint64_t word = s >> 4;
        const uint64_t last_word = (s + (uint64_t)L - 1) >> 4;
        uint64_t acc = 0;
        int nbits = (int)(s & 15) * 2;                         // bits of the first word that belong to earlier reads
        for (int m = 0; m < nw; ++m) {
            const uint32_t v = reverse_bases(w[nw - 1 - m]);
            const int nb = m == 0 ? 32 - 2 * pad : 32;         // the first reversed word starts with the padding slots
            acc = (acc << nb) | (nb == 32 ? (uint64_t)v : (uint64_t)(v & ((1u << nb) - 1u)));
            nbits += nb;
            if (nbits >= 32) {
                const uint32_t out = (uint32_t)(acc >> (nbits - 32));
                nbits -= 32;
                acc &= nbits ? ((1ull << nbits) - 1ull) : 0ull;
                if (word == (s >> 4) || word == last_word) __atomic_fetch_or(&R.seq[word], out, __ATOMIC_RELAXED);
                else R.seq[word] = out;
                ++word;
            }
        }
        if (nbits) __atomic_fetch_or(&R.seq[word], (uint32_t)(acc << (32 - nbits)), __ATOMIC_RELAXED);
    }
    bin.close();
    return R;
}

// ---- --assist_seq (reference s1.cpp:104-134): a fast[aq] file (optionally gzip'ed) whose sequences are appended after the
// short reads through SequencePackage::AppendReverseSeq -- reversed, un-trimmed, chars mapped by dna_map
// (ACGTNacgtn -> 0123201232, sequence_package.h:67-69; anything else -> 0).  `<file>.info` holds `num_seq num_bases`.
// Assist reads count in stage 1 but never get is_solid bits; all their edges are solid in stage 2 (s2.cpp:276,529).
void append_assist(Reads &R, const std::string &file) {
    long long n_seq_info = 0, n_bases_info = 0;
    {
        FILE *f = fopen((file + ".info").c_str(), "r");
        if (!f) die("cannot open " + file + ".info: " + strerror(errno));
        if (fscanf(f, "%lld%lld", &n_seq_info, &n_bases_info) != 2) die("bad " + file + ".info (expected: num_seq num_bases)");
        fclose(f);
    }
    gzFile gz = gzopen(file.c_str(), "rb");
    if (!gz) die("cannot open " + file);
    gzbuffer(gz, 1 << 20);
    unsigned char map[256];
    memset(map, 0, sizeof(map));
    { const char *a = "ACGTNacgtn", *b = "0123201232"; for (int i = 0; i < 10; ++i) map[(unsigned char)a[i]] = (unsigned char)(b[i] - '0'); }
    // kseq-style record reader: '>' or '@' header; sequence lines up to the next header ('>' / '@') or '+' (then as many
    // quality characters as bases are skipped)
    std::vector<std::string> seqs;
    std::string line, cur;
    bool in_seq = false;
    auto getline_gz = [&](std::string &out) {
        out.clear();
        char buf[1 << 16];
        bool any = false;
        while (gzgets(gz, buf, sizeof(buf))) {
            any = true;
            const size_t n = strlen(buf);
            out.append(buf, n);
            if (n && buf[n - 1] == '\n') break;
        }
        while (!out.empty() && (out.back() == '\n' || out.back() == '\r')) out.pop_back();
        return any;
    };
    while (getline_gz(line)) {
        if (line.empty()) continue;
        if (line[0] == '>' || line[0] == '@') {
            if (in_seq) seqs.push_back(cur);
            cur.clear();
            in_seq = true;
        } else if (line[0] == '+' && in_seq) {
            size_t q = 0;
            while (q < cur.size() && getline_gz(line)) q += line.size();
            seqs.push_back(cur);
            cur.clear();
            in_seq = false;
        } else if (in_seq) {
            for (char c : line) if (!isspace((unsigned char)c)) cur.push_back(c);
        }
    }
    if (in_seq) seqs.push_back(cur);
    gzclose(gz);
    unsigned long long extra = 0;
    for (auto &q : seqs) extra += q.size();
    if ((long long)seqs.size() != n_seq_info || (long long)extra != n_bases_info)
        fprintf(stderr, "[B200] warning: %s holds %zu sequences / %llu bases, its .info says %lld / %lld\n", file.c_str(), seqs.size(),
                extra, n_seq_info, n_bases_info);
    const uint64_t total = R.start.back() + extra;
    const uint64_t new_words = total / 16 + 1;
    Reads N2;
    N2.alloc(new_words, R.pinned);
    uint32_t *ns = N2.seq;
    if (!ns) die("out of host memory for the assist sequences");
    memcpy(ns, R.seq, R.n_words * 4);
    R.release();
    R.seq = ns; R.pinned = N2.pinned; R.n_words = new_words;
    uint64_t g = R.start.back();
    for (auto &q : seqs) {
        for (size_t i = q.size(); i-- > 0;) {                  // reversed, not complemented
            ns[g >> 4] |= (uint32_t)map[(unsigned char)q[i]] << ((15 - (g & 15)) * 2);
            ++g;
        }
        R.start.push_back(g);
    }
    R.n_reads += seqs.size();
}

// ---- SdbgWriter equivalent (sdbg_multi_io.h:34-199): one record file per GPU (= per lv1-bucket shard; a bucket's
// records are contiguous in exactly one file, :86-91), one sdbg_info over all of them in the reference's exact text format
struct Writer {
    int file_id = 0;
    FILE *f = nullptr;
    long long offset = 0;
    int wpt = 0;
    std::vector<long long> fid, start, n_items, n_tips, n_large;
    Writer() : fid(MGTA_NUM_BUCKETS, -1), start(MGTA_NUM_BUCKETS, 0), n_items(MGTA_NUM_BUCKETS, 0), n_tips(MGTA_NUM_BUCKETS, 0),
               n_large(MGTA_NUM_BUCKETS, 0) {}
};

int sink(void *user, int32_t b0, int32_t b1, const void *bytes, uint64_t n_bytes, const int64_t *meta) {
    Writer *w = (Writer *)user;
    if (n_bytes && fwrite(bytes, 1, n_bytes, w->f) != n_bytes) return -1;
    long long off = w->offset;
    for (int b = b0; b < b1; ++b) {
        const int64_t *m = meta + (size_t)(b - b0) * 3;
        w->n_items[b] = m[0]; w->n_tips[b] = m[1]; w->n_large[b] = m[2];
        if (m[0]) { w->fid[b] = w->file_id; w->start[b] = off; }
        off += m[0] * 2 + m[2] * 2 + m[1] * 4LL * w->wpt;       // u16 record (+ u16 multiplicity) (+ tip label words)
    }
    if (off != w->offset + (long long)n_bytes) return -2;      // the table must add up to the bytes delivered
    w->offset = off;
    return 0;
}

// sdbg_info over all record files (sdbg_multi_io.h:160-187): num_threads = number of files, every row names its file;
// empty buckets carry file_id -1 (the reader skips only those rows, :356-358).  -> total number of edges
long long write_sdbg_info(const std::string &prefix, int kmer_k, const std::vector<const Writer *> &files) {
    const int wpt = (2 * kmer_k + 31) / 32;
    FILE *info = fopen((prefix + ".sdbg_info").c_str(), "w");
    if (!info) die("cannot write " + prefix + ".sdbg_info");
    long long te = 0, tt = 0, tl = 0;
    for (const Writer *W : files)
        for (int b = 0; b < MGTA_NUM_BUCKETS; ++b) { te += W->n_items[b]; tt += W->n_tips[b]; tl += W->n_large[b]; }
    fprintf(info, "k %d\n", kmer_k);
    fprintf(info, "words_per_tip_label %d\n", wpt);
    fprintf(info, "num_buckets %d\n", MGTA_NUM_BUCKETS);
    fprintf(info, "num_threads %d\n", (int)files.size());
    fprintf(info, "total_size %lld\n", te);
    fprintf(info, "num_tips %lld\n", tt);
    fprintf(info, "large_multi %lld\n", tl);
    for (int b = 0; b < MGTA_NUM_BUCKETS; ++b) {
        const Writer *own = nullptr;
        for (const Writer *W : files) if (W->n_items[b]) { if (own) die("internal: bucket " + std::to_string(b) + " emitted by two GPUs"); own = W; }
        if (own) fprintf(info, "%d %d %lld %lld %lld %lld\n", b, (int)own->fid[b], own->start[b], own->n_items[b], own->n_tips[b], own->n_large[b]);
        else fprintf(info, "%d %d %lld %lld %lld %lld\n", b, -1, 0LL, 0LL, 0LL, 0LL);
    }
    fclose(info);
    return te;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] void die_now(const std::string &msg) {            // from a GPU thread: the other threads may sit in a collective
    fprintf(stderr, "[ERROR] %s\n", msg.c_str());
    fflush(stderr);
    _exit(1);
}

// the collective the library asked for, among the `world` GPU threads of this process, on the context's stream
void run_collective(const mgta_collective &c, int world, ncclComm_t comm, cudaStream_t st) {
    ncclResult_t r = ncclSuccess;
    switch (c.op) {
        case MGTA_COLL_ALL_TO_ALL:
            ncclGroupStart();
            for (int p = 0; p < world && r == ncclSuccess; ++p) {
                r = ncclSend((const char *)c.send + (size_t)p * c.bytes, c.bytes, ncclChar, p, comm, st);
                if (r == ncclSuccess) r = ncclRecv((char *)c.recv + (size_t)p * c.bytes, c.bytes, ncclChar, p, comm, st);
            }
            ncclGroupEnd();
            break;
        case MGTA_COLL_ALL_GATHER: r = ncclAllGather(c.send, c.recv, c.bytes, ncclChar, comm, st); break;
        case MGTA_COLL_ALL_REDUCE_SUM_U32: r = ncclAllReduce(c.recv, c.recv, c.bytes / 4, ncclUint32, ncclSum, comm, st); break;
        case MGTA_COLL_ALL_REDUCE_SUM_U64: r = ncclAllReduce(c.recv, c.recv, c.bytes / 8, ncclUint64, ncclSum, comm, st); break;
        default: die_now("unknown collective requested by the library");
    }
    if (r != ncclSuccess) die_now(std::string("NCCL: ") + ncclGetErrorString(r));
}

struct GpuJob {
    int g = 0, G = 1;
    const Options *opt = nullptr;
    const Reads *R = nullptr;
    uint64_t n_short = 0;
    ncclComm_t comm = nullptr;
    Writer W;
    std::vector<int64_t> ec;
    int64_t totals[10] = {0};
    uint64_t num_mercy = 0;
    mgta_stage_stats s1, s2;
    double t_upload = 0, t_s1 = 0, t_s2 = 0;
};

// one GPU = one lv1-bucket shard = one host thread: upload my slice of the reads, all-gather them over NVLink, then walk
// the library's protocol for both stages
void gpu_thread(GpuJob *J) {
    const Options &opt = *J->opt;
    const Reads &R = *J->R;
    const int g = J->g, G = J->G;
    if (cudaSetDevice(g) != cudaSuccess) die_now("cudaSetDevice(" + std::to_string(g) + ") failed");
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) die_now("cudaStreamCreate failed");
    mgta_opts mo;
    memset(&mo, 0, sizeof(mo));
    mo.kmer_k = opt.kmer_k; mo.min_count = opt.min_count; mo.need_mercy = opt.need_mercy && opt.min_count > 1 ? 1 : 0;
    mo.device = g; mo.rank = g; mo.world = G;
    mo.hbm_budget_bytes = (int64_t)opt.gpu_mem;                 // 0 = 90 % of the free HBM, as the reference's "auto detect"
    mo.stream = st;
    mgta_ctx *ctx = nullptr;
    if (mgta_ctx_create(&mo, &ctx) != 0) die_now(std::string("mgta_ctx_create: ") + mgta_last_error(nullptr));
    auto ck = [&](int rc, const char *what) { if (rc != 0) die_now(std::string(what) + " (GPU " + std::to_string(g) + "): " + mgta_last_error(ctx)); };

    double t = now();
    if (G == 1) {
        // pinned host buffers: the upload runs on a copy stream and hides behind the stage-1 extraction
        ck(mgta_set_reads_async(ctx, R.seq, R.n_words, R.start.data(), R.n_reads, J->n_short, R.max_len), "mgta_set_reads_async");
    } else {
        ck(mgta_alloc_reads(ctx, R.n_words, R.n_reads, J->n_short, R.start.back(), R.max_len), "mgta_alloc_reads");
        void *d_seq, *d_start;
        uint64_t seq_bytes, start_bytes;
        ck(mgta_reads_device_buffers(ctx, &d_seq, &seq_bytes, &d_start, &start_bytes), "mgta_reads_device_buffers");
        // every GPU uploads 1/G of both arrays over its own PCIe link; one group of broadcasts (an all-gather with exact
        // part sizes) over NVLink completes the buffers everywhere
        auto part = [&](uint64_t bytes, int r) { return (bytes * (uint64_t)r / G) & ~(uint64_t)15; };
        for (int pass = 0; pass < 2; ++pass) {
            const char *src = pass ? (const char *)R.start.data() : (const char *)R.seq;
            char *dst = pass ? (char *)d_start : (char *)d_seq;
            const uint64_t bytes = pass ? start_bytes : seq_bytes;
            const uint64_t lo = part(bytes, g), hi = g + 1 == G ? bytes : part(bytes, g + 1);
            if (hi > lo && cudaMemcpyAsync(dst + lo, src + lo, hi - lo, cudaMemcpyHostToDevice, st) != cudaSuccess) die_now("H2D of the reads failed");
            ncclGroupStart();
            for (int r = 0; r < G; ++r) {
                const uint64_t a = part(bytes, r), b = r + 1 == G ? bytes : part(bytes, r + 1);
                if (b > a && ncclBroadcast(dst + a, dst + a, b - a, ncclChar, r, J->comm, st) != ncclSuccess) die_now("ncclBroadcast of the reads failed");
            }
            ncclGroupEnd();
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) die_now("read distribution failed");
    }
    J->t_upload = now() - t;

    auto run_stage = [&](int stage) {
        ck(mgta_sharded_begin(ctx, stage, stage == 2 ? sink : nullptr, stage == 2 ? (void *)&J->W : nullptr), "mgta_sharded_begin");
        for (;;) {
            mgta_collective c;
            ck(mgta_sharded_step(ctx, &c), stage == 1 ? "stage 1" : "stage 2");
            if (c.op == MGTA_COLL_NONE) break;
            run_collective(c, G, J->comm, st);
        }
    };
    t = now();
    if (opt.min_count > 1) {
        run_stage(1);
        J->ec.assign(MGTA_NUM_BUCKETS, 0);
        ck(mgta_sharded_result(ctx, J->ec.data(), nullptr), "mgta_sharded_result");
        if (mo.need_mercy) ck(mgta_get_num_mercy(ctx, &J->num_mercy), "mgta_get_num_mercy");
    }
    J->t_s1 = now() - t;
    t = now();
    J->W.file_id = g;
    J->W.wpt = (2 * opt.kmer_k + 31) / 32;
    const std::string fn = opt.output_prefix + ".sdbg." + std::to_string(g);
    J->W.f = fopen(fn.c_str(), "wb");
    if (!J->W.f) die_now("cannot write " + fn);
    run_stage(2);
    ck(mgta_sharded_result(ctx, nullptr, J->totals), "mgta_sharded_result");
    fclose(J->W.f);
    J->t_s2 = now() - t;
    mgta_get_stats(ctx, 1, &J->s1);
    mgta_get_stats(ctx, 2, &J->s2);
    mgta_ctx_destroy(ctx);
    cudaStreamDestroy(st);
}

}  // namespace

int build_graph(int argc, char **argv) {
    const double t0 = now();
    Options opt = parse(argc, argv);

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) die("no CUDA device (there is no CPU fallback)");
    int G = n_dev;                                               // all visible GPUs (CUDA_VISIBLE_DEVICES); MGTA_NUM_GPUS caps it
    if (const char *e = getenv("MGTA_NUM_GPUS")) G = std::max(1, std::min(n_dev, atoi(e)));
    G = std::min(G, 16);
    if (opt.need_mercy && opt.min_count > 1 && G > 1 && !opt.assist_seq_file.empty()) {
        fprintf(stderr, "[B200] --need_mercy together with --assist_seq runs on one GPU: using GPU 0 only\n");
        G = 1;
    }

    Reads R = load_read_lib(opt.read_lib_file, opt.num_cpu_threads, G == 1);
    fprintf(stderr, "[B200] %llu reads, %llu bases, max length %d, loaded in %.2f s\n", (unsigned long long)R.n_reads,
            (unsigned long long)R.start.back(), R.max_len, now() - t0);
    if (R.n_reads == 0) die("empty read library");
    const uint64_t n_short = R.n_reads;                          // max_len stays that of the short reads (s1.cpp:119)
    if (!opt.assist_seq_file.empty()) {
        append_assist(R, opt.assist_seq_file);
        fprintf(stderr, "[B200] %llu assist sequences, %llu bases in all\n", (unsigned long long)(R.n_reads - n_short),
                (unsigned long long)R.start.back());
    }

    std::vector<ncclComm_t> comms(G, nullptr);
    if (G > 1) {
        std::vector<int> devs(G);
        for (int g = 0; g < G; ++g) devs[g] = g;
        ncclResult_t r = ncclCommInitAll(comms.data(), G, devs.data());
        if (r != ncclSuccess) die(std::string("ncclCommInitAll: ") + ncclGetErrorString(r));
    }
    const double t1 = now();
    std::vector<GpuJob> jobs(G);
    std::vector<std::thread> threads;
    for (int g = 0; g < G; ++g) {
        jobs[g].g = g; jobs[g].G = G; jobs[g].opt = &opt; jobs[g].R = &R; jobs[g].n_short = n_short; jobs[g].comm = comms[g];
        threads.emplace_back(gpu_thread, &jobs[g]);
    }
    for (auto &th : threads) th.join();
    const double t2 = now();
    for (auto c : comms) if (c) ncclCommDestroy(c);

    if (opt.min_count > 1) {
        const std::vector<int64_t> &ec = jobs[0].ec;                // all-reduced: whole on every GPU
        FILE *cf = fopen((opt.output_prefix + ".counting").c_str(), "w");     // s1.cpp:925-930
        if (!cf) die("cannot write " + opt.output_prefix + ".counting");
        long long acc = 0;
        for (int i = 1; i <= 65535; ++i) { acc += ec[i]; fprintf(cf, "%lld %lld\n", (long long)i, acc); }
        fclose(cf);
        long long solid = 0;
        for (int i = opt.min_count; i <= 65535; ++i) solid += ec[i];
        fprintf(stderr, "[B200] Total number of solid edges: %lld\n", solid);
        if (opt.need_mercy) fprintf(stderr, "[B200] Number mercy: %llu\n", (unsigned long long)jobs[0].num_mercy);   // s2.cpp:241
    }

    std::vector<const Writer *> files;
    for (auto &J : jobs) files.push_back(&J.W);
    const long long te = write_sdbg_info(opt.output_prefix, opt.kmer_k, files);

    for (auto &J : jobs)
        fprintf(stderr, "[B200] GPU %d: reads on the device in %.2f s; stage 1 %.1f ms device (%llu items) / %.2f s wall; stage 2 %.1f ms device "
                        "(%llu items, %llu edges) / %.2f s wall\n", J.g, J.t_upload, J.s1.ms_total, (unsigned long long)J.s1.n_items, J.t_s1,
                J.s2.ms_total, (unsigned long long)J.s2.n_items, (unsigned long long)J.s2.n_edges, J.t_s2);
    fprintf(stderr, "[B200] %d GPU(s): reads-in-host-memory -> files: %.2f s; %lld edges\n", G, t2 - t1, te);
    fprintf(stderr, "Real: %.4f\n", now() - t0);                 // utils.h:124 prints the same line
    R.release();
    return 0;
}

int build_lib_b200(int argc, char **argv);       // buildlib_b200.cpp
int find_start_b200(int argc, char **argv);      // findstart_b200.cpp

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "buildgraph") == 0) return build_graph(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "buildlib") == 0) return build_lib_b200(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "findstart") == 0) return find_start_b200(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s buildlib <read_lib_file> <out_prefix> | buildgraph [options] | findstart <ref_seq> <reads.bin> <k_size> [threads] [contigs]\n"
                    "       (drop-ins for the `megagta` sub-programs of the same "
                    "names, reference src/megagta.cpp:28-60)\n", argv[0]);
    return 1;
}
