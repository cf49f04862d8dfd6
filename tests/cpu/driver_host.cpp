// tests/cpu/driver_host.cpp -- TEST INFRASTRUCTURE.  Exposes the HOST-side logic of the C++ drivers (megagta_b200/csrc/host/)
// to the CPU tests: the read-library loader (reversed, bit-contiguous packing), --assist_seq, the record-file writer and
// sdbg_info, the kseq-rule FASTA / FASTQ reader and the model k-mer rules of findstart.  The driver sources are included as
// they are (their main() renamed); nothing here touches a kernel, and the product never contains this file.
#define main mgta_driver_main_unused
#include "../../megagta_b200/csrc/host/buildgraph_b200.cpp"
#undef main
int build_lib_b200(int, char **) { return -1; }       // the other two sub-programs are not part of this harness
int find_start_b200(int, char **) { return -1; }
#include "../../megagta_b200/csrc/host/fastx_reader.h"
#include "../../megagta_b200/csrc/host/prot_kmers.h"

extern "C" {

// -> number of reads; *seq (malloc'ed, n_words u32), *start (malloc'ed, n_reads + 1 u64)
int64_t hd_load_reads(const char *prefix, const char *assist, int threads, uint32_t **seq, uint64_t *n_words, uint64_t **start,
                      int *max_len, uint64_t *n_short) {
    Reads R = load_read_lib(prefix, threads, false);
    *n_short = R.n_reads;
    if (assist && *assist) append_assist(R, assist);
    *seq = (uint32_t *)malloc(R.n_words * 4);
    memcpy(*seq, R.seq, R.n_words * 4);
    *start = (uint64_t *)malloc(R.start.size() * 8);
    memcpy(*start, R.start.data(), R.start.size() * 8);
    *n_words = R.n_words;
    *max_len = R.max_len;
    const int64_t n = (int64_t)R.n_reads;
    R.release();
    return n;
}

void hd_free(void *p) { free(p); }

// the writer: `n_files` record files; file f takes the deliveries whose first bucket lies in [cut[f], cut[f + 1]).
// deliveries: d_b0/d_b1 bucket ranges in ascending order, bytes concatenated, meta = rows of the covered buckets
int hd_write_graph(const char *prefix, int k, int n_files, const int32_t *cut, int n_del, const int32_t *d_b0, const int32_t *d_b1,
                   const uint8_t *bytes, const uint64_t *d_bytes, const int64_t *meta) {
    std::vector<Writer> W(n_files);
    for (int f = 0; f < n_files; ++f) {
        W[f].file_id = f;
        W[f].wpt = (2 * k + 31) / 32;
        W[f].f = fopen((std::string(prefix) + ".sdbg." + std::to_string(f)).c_str(), "wb");
        if (!W[f].f) return -10;
    }
    uint64_t at = 0;
    for (int d = 0; d < n_del; ++d) {
        int f = 0;
        while (f + 1 < n_files && d_b0[d] >= cut[f + 1]) ++f;
        const int rc = sink(&W[f], d_b0[d], d_b1[d], bytes + at, d_bytes[d], meta + (size_t)d_b0[d] * 3);
        if (rc) return rc;
        at += d_bytes[d];
    }
    std::vector<const Writer *> files;
    for (auto &w : W) { fclose(w.f); files.push_back(&w); }
    return (int)(write_sdbg_info(prefix, k, files) & 0x7FFFFFFF);
}

// sequences of a FASTA / FASTQ file as the drivers see them: concatenated into *out (malloc'ed), offsets into *off
int64_t hd_fastx(const char *path, char **out, uint64_t **off) {
    FastxReader fr(path);
    std::string seq, all;
    std::vector<uint64_t> o{0};
    while (fr.next(seq)) { all += seq; o.push_back(all.size()); }
    *out = (char *)malloc(all.size() + 1);
    memcpy(*out, all.data(), all.size());
    *off = (uint64_t *)malloc(o.size() * 8);
    memcpy(*off, o.data(), o.size() * 8);
    return (int64_t)o.size() - 1;
}

// model k-mers of an aligned reference (findstart): rows of (w0, w1, model position) in file order, + the decoded text
int64_t hd_model_kmers(const char *faa, int k, uint64_t **rows, char **text) {
    FastxReader fr(faa);
    std::string seq, txt;
    std::vector<mgta_host::ModelKmer> m;
    while (fr.next(seq)) mgta_host::model_kmers_of(seq, k, m);
    *rows = (uint64_t *)malloc(m.size() * 24 + 8);
    for (size_t i = 0; i < m.size(); ++i) {
        (*rows)[3 * i] = m[i].w[0]; (*rows)[3 * i + 1] = m[i].w[1]; (*rows)[3 * i + 2] = (uint64_t)m[i].model_pos;
        txt += mgta_host::unpack_key(m[i].w, k);
    }
    *text = (char *)malloc(txt.size() + 1);
    memcpy(*text, txt.data(), txt.size());
    return (int64_t)m.size();
}
}
