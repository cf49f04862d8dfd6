// integration/level2_sink.h -- the stage-2 sink of the Level-2 binding: a delivery of the C ABI replayed
// through the reference's own, unmodified SdbgWriter::write (sdbg_multi_io.h:83-112), so that file layout and sdbg_info stay
// the reference's code.  A delivery holds the records of buckets [bucket_begin, bucket_end) back to back; `meta` gives the
// number of records per bucket (include/mgta_cuda.h, mgta_bucket_sink).
#pragma once
#include <stdint.h>
#include <string.h>

#include "definitions.h"
#include "sdbg_multi_io.h"

struct Level2Sink {
    SdbgWriter *writer;
    int words_per_tip_label;
    int file_id;                       // the writer's `tid`: one file per GPU
};

inline int level2_replay_sink(void *user, int32_t bucket_begin, int32_t bucket_end, const void *bytes, uint64_t n_bytes,
                              const int64_t *meta) {
    Level2Sink *s = static_cast<Level2Sink *>(user);
    const uint16_t *p = static_cast<const uint16_t *>(bytes), *end = p + n_bytes / 2;
    uint32_t label[8];                 // ceil(2 * kMaxK / 32) words
    for (int32_t b = bucket_begin; b < bucket_end; ++b) {
        const int64_t n_items = meta[(size_t)(b - bucket_begin) * 3];
        for (int64_t i = 0; i < n_items; ++i) {
            if (p >= end) return -1;
            const uint16_t rec = *p++;                                   // w | last << 4 | tip << 5 | min(mult, 255) << 8
            const int w = rec & 15, last = (rec >> 4) & 1, tip = (rec >> 5) & 1;
            multi_t mult = (multi_t)(rec >> 8);
            if (mult == kMulti2Sp) {                                     // the real multiplicity follows (> kMaxMulti2_t)
                if (p >= end) return -1;
                mult = *p++;
            }
            if (tip) {
                if (p + 2 * s->words_per_tip_label > end) return -1;
                memcpy(label, p, 4 * (size_t)s->words_per_tip_label);
                p += 2 * s->words_per_tip_label;
            }
            s->writer->write(s->file_id, b, w, last, tip, mult, label);
        }
    }
    return p == end ? 0 : -2;                                            // the table must add up to the bytes delivered
}

#ifdef MGTA_HAVE_APPEND_RAW
// The fast binding, for a reference tree that carries integration/sdbg_writer_append_raw.patch: one fwrite per delivery.
inline int level2_raw_sink(void *user, int32_t bucket_begin, int32_t bucket_end, const void *bytes, uint64_t n_bytes,
                           const int64_t *meta) {
    Level2Sink *s = static_cast<Level2Sink *>(user);
    uint64_t expect = 0;
    for (int32_t b = bucket_begin; b < bucket_end; ++b) {
        const int64_t *m = meta + (size_t)(b - bucket_begin) * 3;
        expect += (uint64_t)m[0] * 2 + (uint64_t)m[2] * 2 + (uint64_t)m[1] * 4 * (uint64_t)s->words_per_tip_label;
    }
    if (expect != n_bytes) return -2;                                    // the table must add up to the bytes delivered
    s->writer->append_raw(s->file_id, bucket_begin, bucket_end, bytes, n_bytes, meta);
    return 0;
}
#define LEVEL2_SINK level2_raw_sink
#else
#define LEVEL2_SINK level2_replay_sink
#endif
