/*
 * mgta_cuda.h -- C ABI of libmgta_cuda.so: the B200 (sm_100a) implementation of MegaGTA's CX1
 * reads -> succinct de Bruijn graph construction (`megagta buildgraph`).
 *
 * This is the drop-in boundary.  The reference has no FFI for this path: it is one process whose
 * `build_graph()` (reference src/build_graph.cpp:33-135, declared src/megagta.cpp:9) installs 12
 * CX1 callbacks twice (stage 1: build_graph.cpp:100-113, stage 2: :120-132) and calls
 * `CX1::run()` (src/cx1.h:443-623).  The only historic GPU seam is the undeclared
 * `lv2_gpu_sort(...)` / `alloc_gpu_buffers` / `free_gpu_buffers` under `#ifdef USE_GPU`
 * (src/cx1_read2sdbg_s1.cpp:308,619,945; src/cx1_read2sdbg_s2.cpp:391,697,926).  A host
 * `build_graph()` replacement (megagta_b200/csrc/host/buildgraph_b200.cpp) keeps the reference's
 * option table, read-library loader and output writer and calls the entry points below; see
 * INTEGRATION.md for the patch a reference maintainer would apply.
 *
 * Conventions: plain C, POD arguments, no exceptions cross the boundary.  Every function returns
 * 0 on success and a negative mgta_status on failure; mgta_last_error() gives the message.
 * One caller thread per context.  The caller owns all host buffers; the library owns all device
 * memory.  There is NO CPU fallback: without a CUDA device mgta_ctx_create() fails.
 */
#ifndef MGTA_CUDA_H_
#define MGTA_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGTA_NUM_BUCKETS 65536 /* reference cx1_read2sdbg.h:66 (kNumBuckets = 4^8) */
#define MGTA_MAX_K 127         /* reference definitions.h:56 (kMaxK) */

typedef enum {
    MGTA_OK = 0,
    MGTA_ERR_ARG = -1,      /* bad argument / option (k range, missing reads, ...) */
    MGTA_ERR_CUDA = -2,     /* CUDA runtime error, incl. "no device" */
    MGTA_ERR_MEM = -3,      /* HBM budget too small for the largest bucket */
    MGTA_ERR_STATE = -4,    /* call order (stage 2 before reads, ...) */
    MGTA_ERR_INTERNAL = -5  /* device-side consistency check failed */
} mgta_status;

typedef struct mgta_ctx mgta_ctx;

/* Options of one context = one GPU = one contiguous lv1-bucket shard.
 * Replaces read2sdbg_opt_t / read2sdbg_global_t fields (reference cx1_read2sdbg.h:36-59,76-89). */
typedef struct {
    int32_t kmer_k;           /* -k: graph order; edges are (k+1)-mers.  9 <= k <= 127 */
    int32_t min_count;        /* -m: solid (k+1)-mer threshold; 1 skips stage 1 */
    int32_t need_mercy;       /* --need_mercy: also emit mercy candidates in stage 1 */
    int32_t device;           /* CUDA device ordinal */
    int32_t rank;             /* shard index in [0, world) */
    int32_t world;            /* number of bucket shards (GPUs); 1 = whole graph */
    int64_t hbm_budget_bytes; /* --gpu_mem: 0 = 90 % of the currently free device memory */
    void *stream;             /* cudaStream_t to launch on; NULL = a stream owned by the context */
    int32_t sort_items_cap;   /* items per on-chip sort tile (test hook); 0 = auto */
    int32_t reserved;
} mgta_opts;

/* One delivery of stage-2 output: the records of buckets [bucket_begin, bucket_end), concatenated
 * in ascending bucket order (the layout SdbgWriter::write produces per bucket,
 * sdbg_multi_io.h:83-112).  `meta` holds 3 int64 per bucket of the range: num_items, num_tips,
 * num_large_mul (the sdbg_info row, sdbg_multi_io.h:178-185).  Pointers are valid only during the
 * call.  Deliveries arrive in ascending bucket order. */
typedef int (*mgta_bucket_sink)(void *user, int32_t bucket_begin, int32_t bucket_end,
                                const void *bytes, uint64_t n_bytes, const int64_t *meta);

/* Per-stage statistics (device-timed with CUDA events on the context's stream). */
typedef struct {
    uint64_t n_items;        /* items sorted (stage-1 (k-1)-mer contexts / stage-2 edge contexts) */
    uint64_t n_batches;      /* bucket-range batches the HBM budget forced */
    uint64_t n_launches;     /* kernels launched */
    uint64_t n_giants;       /* groups too large for the on-chip sort (counted, not sorted) */
    uint64_t out_bytes;      /* stage 2: bytes of SdBG records emitted by this shard */
    uint64_t n_edges;        /* stage 2: records emitted by this shard (total_size share) */
    float ms_total;          /* whole stage, device time */
    float ms_hist;           /* lv1 bucket histogram kernel */
    float ms_extract;        /* item extraction + bucket scatter */
    float ms_partition;      /* MSD digit partition levels */
    float ms_sort_emit;      /* on-chip multi-word LSD radix sort + counting / emission */
    int32_t key_words;       /* u32 words per key */
    int32_t item_words;      /* u32 words per item (key + payload) */
    int32_t sort_cap;        /* items per on-chip sort tile */
    int32_t msd_levels;      /* digit partition levels that ran */
    float ms_nodes;          /* stage 2: node pass (tip k-mers), device time */
    int32_t reserved;
    uint64_t n_node_ops;     /* stage 2: k-mer ops of the node pass (2 per distinct solid edge) */
    uint64_t n_tip_items;    /* stage 2: $-items the node pass produced (2 per tip k-mer) */
} mgta_stage_stats;

int mgta_ctx_create(const mgta_opts *opts, mgta_ctx **out);
void mgta_ctx_destroy(mgta_ctx *ctx);
const char *mgta_last_error(const mgta_ctx *ctx); /* ctx may be NULL: last create error */

/* Reads as the reference holds them after ReadBinaryLibs(..., is_reverse=true)
 * (read_lib_functions-inl.h:233-261; sequence_package.h:34-406): 2-bit bases, MSB first, 16 per
 * u32, bit-contiguous, each read REVERSED; start_idx[i] = first base of read i, start_idx[n_reads]
 * = total bases.  Reads [n_short_reads, n_reads) are assist sequences (always solid, s2.cpp:276).
 * max_read_len = longest short read (s1.cpp:119).  Copies host -> device. */
int mgta_set_reads(mgta_ctx *ctx, const uint32_t *packed_seq, uint64_t n_words, const uint64_t *start_idx,
                   uint64_t n_reads, uint64_t n_short_reads, int32_t max_read_len);

/* The same, asynchronous: returns once the copies are queued (start_idx first, then packed_seq in chunks on a copy
 * stream of the context); mgta_stage1 extracts a chunk as soon as the next one has landed, so the upload hides behind
 * the extraction.  Every other entry point waits for the whole upload.  The host buffers must be pinned
 * (cudaHostAlloc / cudaHostRegister) and stay valid and unchanged until the next mgta_stage1 / mgta_stage2 /
 * histogram call on the context has returned; pageable buffers take the synchronous path of mgta_set_reads. */
int mgta_set_reads_async(mgta_ctx *ctx, const uint32_t *packed_seq, uint64_t n_words, const uint64_t *start_idx,
                         uint64_t n_reads, uint64_t n_short_reads, int32_t max_read_len);

/* Multi-GPU read distribution (north star: "packed reads are broadcast with NCCL over NVLink"):
 * a non-root shard allocates device buffers of the right size, the caller broadcasts the root's
 * buffers into them (ncclBroadcast / torch.distributed.broadcast) and no host copy is needed.
 * seq buffer = n_words u32 (+ zero padding owned by the library), start buffer = (n_reads+1) u64. */
int mgta_alloc_reads(mgta_ctx *ctx, uint64_t n_words, uint64_t n_reads, uint64_t n_short_reads,
                     uint64_t total_bases, int32_t max_read_len);
int mgta_reads_device_buffers(mgta_ctx *ctx, void **seq_dev, uint64_t *seq_bytes, void **start_dev,
                              uint64_t *start_bytes);

/* lv1 bucket histograms (reference s1_lv0_calc_bucket_size s1.cpp:177-229 and
 * s2_lv0_calc_bucket_size s2.cpp:252-315).  hist: int64[65536], whole bucket space. */
int mgta_stage1_histogram(mgta_ctx *ctx, int64_t *hist);
int mgta_stage2_histogram(mgta_ctx *ctx, int64_t *hist);

/* Stage 1 (reference cx1.run() with the s1 callbacks): marks solid (k+1)-mers of this shard's
 * buckets in the device-resident is_solid vector and accumulates edge_counting
 * (int64[65536], may be NULL; s1.cpp:744-746).  No-op success when min_count == 1. */
int mgta_stage1(mgta_ctx *ctx, int64_t *edge_counting);

/* The device bit vector (one bit per base position, bit start_idx[r]+o <=> edge offset o of read r).  With world > 1 and
 * mgta_stage1 (every shard scanning all reads for its hash range) each bit is set by exactly one shard, so an
 * all-reduce SUM over uint32 words merges the shards.  Not on the hot path: the sharded build below exchanges items. */
int mgta_solid_device_buffer(mgta_ctx *ctx, void **dev_ptr, uint64_t *n_bytes);

/* is_solid in the reference's layout (AtomicBitVector, atomic_bit_vector.h:58-60; bit index
 * (max_read_len-k)*read_id + offset, s1.cpp:151,760).  n_bytes >= ceil(n_short*(max_len-k)/8). */
int mgta_get_is_solid(mgta_ctx *ctx, uint8_t *host, uint64_t n_bytes);
int mgta_set_is_solid(mgta_ctx *ctx, const uint8_t *host, uint64_t n_bytes);

/* Mercy edges (opts.need_mercy, min_count > 1; on one shard through mgta_stage1, on several through the sharded build
 * below, which exchanges the candidates): stage 1 then also (a) emits the mercy candidates of
 * s1_lv2_output_ (s1.cpp:762-826: packed ((start_idx + kmer_offset) << 2) | flag; the reference spreads them over
 * <prefix>.mercy_cand.N files, here they stay on the device) and (b) runs the per-read scan of s2_read_mercy_prepare
 * (s2.cpp:106-250) that extends is_solid; mgta_stage2 then builds the graph from the extended vector.
 * mgta_get_mercy_candidates: *n receives the count; copies min(*n, cap) values (any order). */
int mgta_get_mercy_candidates(mgta_ctx *ctx, uint64_t *host, uint64_t cap, uint64_t *n);
/* Number of is_solid bits the per-read mercy scan added (the reference logs it as "Number mercy", s2.cpp:241). */
int mgta_get_num_mercy(mgta_ctx *ctx, uint64_t *num_mercy);

/* ---- Sharded build, world > 1: the protocol lives in the library, the caller only runs collectives ----------------------
 * One context per GPU (rank r of world), one caller thread or process per context.  Every context holds ALL reads
 * (upload 1/world each + all-gather, or broadcast: mgta_alloc_reads / mgta_reads_device_buffers).  Then, per stage:
 *
 *     mgta_sharded_begin(ctx, stage, sink, user);
 *     for (;;) {
 *         mgta_collective c;
 *         mgta_sharded_step(ctx, &c);            // runs this shard's kernels up to the next exchange
 *         if (c.op == MGTA_COLL_NONE) break;     // stage finished
 *         <run collective c among the `world` contexts, ordered on the context's stream>
 *     }
 *     mgta_sharded_result(ctx, edge_counting, totals);
 *
 * All ranks see the same sequence of collectives.  Buffers are device memory owned by the library.  The collective must
 * be enqueued on (or otherwise ordered with) the stream of the context (mgta_opts.stream): the library launches the
 * kernels that produce `send` and consume `recv` on that stream and does not synchronise with any other.
 *   ALL_TO_ALL     send / recv hold `world` slabs of `bytes` bytes: slab d of send goes to rank d, slab s of recv comes
 *                  from rank s (ncclSend / ncclRecv in one group; torch.distributed.all_to_all_single)
 *   ALL_GATHER     every rank contributes `bytes` bytes at send == recv + rank * bytes (in place); recv holds world * bytes
 *   ALL_REDUCE_*   in place (send == recv), `bytes` / 4 resp. / 8 elements, operator SUM
 * Stage 1 replaces cx1.run() with the s1 callbacks (build_graph.cpp:100-113): shard r extracts the canonical (k+1)-mers
 * of ITS 1/world of the reads, binned by the shard that owns their hash range; one all-to-all moves them; every shard
 * counts what it received and keeps the solid edges of its hash range; edge_counting is all-reduced (every rank gets the
 * whole array).  Stage 2 replaces the s2 run (build_graph.cpp:120-132) and replicates nothing: the node pass runs over
 * the k-mers of each shard's hash range (ops all-to-all), the key-prefix histogram is all-reduced, every shard sends the
 * stage-2 items of its edges and tips to the shard that emits their lv1 bucket (all-to-all), and every rank sorts and
 * emits its bucket range to its sink (ascending buckets; ranks in order give the whole graph). */
typedef enum {
    MGTA_COLL_NONE = 0,
    MGTA_COLL_ALL_TO_ALL = 1,
    MGTA_COLL_ALL_GATHER = 2,
    MGTA_COLL_ALL_REDUCE_SUM_U32 = 3,
    MGTA_COLL_ALL_REDUCE_SUM_U64 = 4
} mgta_coll_op;

typedef struct {
    int32_t op;       /* mgta_coll_op */
    int32_t reserved;
    void *send;       /* device */
    void *recv;       /* device */
    uint64_t bytes;   /* see above */
} mgta_collective;

int mgta_sharded_begin(mgta_ctx *ctx, int stage /*1|2*/, mgta_bucket_sink sink, void *user);
int mgta_sharded_step(mgta_ctx *ctx, mgta_collective *next);
/* after the stage finished: stage 1 -> edge_counting int64[65536] (whole graph, may be NULL);
 * stage 2 -> totals int64[10] of THIS shard (may be NULL; sum over the shards) */
int mgta_sharded_result(mgta_ctx *ctx, int64_t *edge_counting, int64_t *totals);

/* Stage 2 (cx1.run() with the s2 callbacks + SdbgWriter): emits this shard's buckets in ascending
 * order.  sink may be NULL (device-resident run: records are produced and counted, not copied).
 * totals: int64[10] = num_w[0..8], num_last1 (sdbg_multi_io.h:114-143); may be NULL. */
int mgta_stage2(mgta_ctx *ctx, mgta_bucket_sink sink, void *user, int64_t *totals);

/* First/last+1 bucket of this shard for the stage whose histogram was computed last. */
int mgta_shard_range(mgta_ctx *ctx, int32_t *bucket_begin, int32_t *bucket_end);

int mgta_get_stats(mgta_ctx *ctx, int stage /*1|2*/, mgta_stage_stats *out);

/* u32 words per sort key: ceil((2(k-1)+6)/32) stage 1 (s1.cpp:246), ceil((2k+4)/32) stage 2 (s2.cpp:331) */
int mgta_words_per_key(int stage, int kmer_k);
int mgta_abi_version(void);

/* ---- SdBG load + rank/select build on the device -----------------------------------------------------
 * Replaces SuccinctDBG::LoadFromMultiFile (reference succinct_dbg.cpp:595-723: a serial u16-at-a-time loop over the whole
 * record stream), SuccinctDBG::init (succinct_dbg.h:62-83) and the table builds RankAndSelect4Bits::Build
 * (rank_and_select.h:81-150) / RankAndSelect1Bit::Build (rank_and_select.h:420-487).  The builder eats the stage-2
 * deliveries (mgta_sdbg_sink has the mgta_bucket_sink signature; `bytes` may also be a DEVICE pointer, so a stream that
 * is still in HBM never visits the host) or the bucket ranges of <prefix>.sdbg.<i> files, and leaves every array of the
 * in-memory graph on the device in the reference's own layout, bit for bit:
 *   w (4 bits per edge, 16 per u64), last / is_tip / invalid (= is_tip | W == 0) / is_multi_1 (bit vectors, 64 per u64),
 *   edge_multi (u8 per edge; 255 = see the large list) + (edge, multiplicity) pairs of the large multiplicities in edge
 *   order (the reference keeps them in a khash; its alternative u16-per-edge array for graphs with > 8 % large
 *   multiplicities is the same information), tip_node_seq, f / rank_f, and the sampled rank (major i64 every 65536 +
 *   minor u16 every 256) and select (interval of every 256th occurrence) tables of W per character, of last and of is_tip.
 * need_multiplicity mirrors LoadFromMultiFile's flag: 1 -> edge_multi + large list, 0 -> is_multi_1. */
typedef struct mgta_sdbg mgta_sdbg;

typedef struct {
    int64_t size;                  /* edges (total_size of sdbg_info) */
    int32_t kmer_k, words_per_tip_label;
    int64_t num_tips, num_large_mul;
    int64_t f[6], rank_f[6];       /* SdbgReader::read_info f_ (sdbg_multi_io.h:253-268); rank_f = ones of last before f[i] */
    int64_t w_freq[9];             /* RankAndSelect4Bits::char_frequency */
    int64_t last_ones, tip_ones;   /* RankAndSelect1Bit::total_num_ones */
    int64_t n_minor, n_major;      /* entries of every minor / major table */
} mgta_sdbg_header_t;

typedef enum {
    MGTA_SDBG_W = 0, MGTA_SDBG_LAST = 1, MGTA_SDBG_IS_TIP = 2, MGTA_SDBG_INVALID = 3, MGTA_SDBG_IS_MULTI_1 = 4,
    MGTA_SDBG_EDGE_MULTI = 5, MGTA_SDBG_LARGE_EDGE = 6 /* u64 */, MGTA_SDBG_LARGE_VALUE = 7 /* u16 */, MGTA_SDBG_TIP_SEQ = 8,
    /* tables (after mgta_sdbg_finish); W_* take the character c in [0, 9) */
    MGTA_SDBG_W_MINOR = 16, MGTA_SDBG_W_MAJOR = 17, MGTA_SDBG_W_SELECT = 18, MGTA_SDBG_LAST_MINOR = 19, MGTA_SDBG_LAST_MAJOR = 20,
    MGTA_SDBG_LAST_SELECT = 21, MGTA_SDBG_TIP_MINOR = 22, MGTA_SDBG_TIP_MAJOR = 23
} mgta_sdbg_array_id;

int mgta_sdbg_create(int device, void *stream /* cudaStream_t or NULL */, int kmer_k, int need_multiplicity, mgta_sdbg **out);
void mgta_sdbg_destroy(mgta_sdbg *g);
const char *mgta_sdbg_last_error(const mgta_sdbg *g);
/* records of buckets [bucket_begin, bucket_end) in bucket order, deliveries in ascending order starting at bucket 0;
 * bytes: host or device; meta: host, 3 int64 per bucket (num_items, num_tips, num_large_mul) */
int mgta_sdbg_append(mgta_sdbg *g, int32_t bucket_begin, int32_t bucket_end, const void *bytes, uint64_t n_bytes, const int64_t *meta);
int mgta_sdbg_sink(void *sdbg, int32_t bucket_begin, int32_t bucket_end, const void *bytes, uint64_t n_bytes, const int64_t *meta);
int mgta_sdbg_finish(mgta_sdbg *g);            /* rank / select tables, f, rank_f */
int mgta_sdbg_header(const mgta_sdbg *g, mgta_sdbg_header_t *h);
int mgta_sdbg_array(mgta_sdbg *g, int which, int c, const void **dev_ptr, uint64_t *n_bytes);
int mgta_sdbg_copy(mgta_sdbg *g, int which, int c, void *host, uint64_t n_bytes);
/* Stage 2 straight into the builder: the records of every batch are parsed where they lie in HBM (no D2H, no sink). */
int mgta_stage2_into_sdbg(mgta_ctx *ctx, mgta_sdbg *g, int64_t *totals);

/* ---- the streaming passes next to the graph build ----------------------------------------------------
 * `megagta buildlib`: ASCII bases -> the records of <X>.bin (per read u32 length, then ceil(len / 16) u32 words, first base
 * in bits 31..30, unused low bits zero; A C G T N a c g t n -> 0 1 2 3 2 0 1 2 3 2).  Replaces
 * SequencePackage::AddSeqToPackedSeq_ (reference sequence_package.h:254-270, map :67-69) + SequenceManager::
 * WriteBinarySequences (sequence_manager.cpp:375-410).  bases: host, the reads back to back; seq_off[r] .. seq_off[r + 1] =
 * the bytes of read r; out_records: host, out_words = sum of 1 + ceil(len / 16).  FASTA/Q parsing stays with the caller. */
int mgta_pack_reads(int device, const char *bases, const uint64_t *seq_off, uint64_t n_reads, uint32_t *out_records, uint64_t out_words);
const char *mgta_tools_last_error(void);
/* `megagta findstart` (reference fast_kmer_filter.cpp:49-218): every read of a batch of .bin records (n_reads records in
 * n_words u32) of at least min_len bases is translated on both strands in all three frames (standard code, sequence/
 * Codon.C:8-90) and every window of aa_k residues is looked up among the model k-mers.  model_kmers: [n_model][2] u64, 5 bits
 * per residue in the codes of ProtKmer::setUp (prot_kmer.h:26-43: ARNDCQEGHI = 0..9, LKMFPSTWYV = 10..19, * = 20), residues
 * 0..11 in word 0 (first residue most significant), 12..23 in word 1; of equal k-mers the first counts (HashSetST::insert).
 * A hit: hit_pos = read << 24 | strand << 23 | nucleotide offset in the strand's own direction, hit_model = index of the
 * model k-mer.  *n_hits keeps counting past hits_cap (call again with room). */
int mgta_find_seeds(int device, const uint64_t *model_kmers, uint64_t n_model, int aa_k, const uint32_t *records, uint64_t n_words,
                    uint64_t n_reads, int min_len, uint64_t *hit_pos, uint32_t *hit_model, uint64_t hits_cap, uint64_t *n_hits);

#ifdef __cplusplus
}
#endif
#endif /* MGTA_CUDA_H_ */
