// fastx_reader.h -- kseq's view of a FASTA / FASTQ stream (optionally gzip'ed), shared by the buildlib and findstart drivers.
#pragma once
#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <string>
#include <vector>

// kseq's view of a FASTA / FASTQ stream: the record rules of kseq_read (reference kseq.h:168-207), restated over gzread
class FastxReader {
  public:
    explicit FastxReader(const std::string &path) : buf_(1 << 20) {
        gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) { fprintf(stderr, "[ERROR] cannot open %s: %s\n", path.c_str(), strerror(errno)); exit(1); }
        gzbuffer(gz_, 1 << 20);
    }
    ~FastxReader() { if (gz_) gzclose(gz_); }
    FastxReader(const FastxReader &) = delete;
    FastxReader &operator=(const FastxReader &) = delete;

    // next record's sequence -> seq; false at the end of the file (or at a truncated FASTQ record, which the reference
    // treats like the end: kseq_read() < 0)
    bool next(std::string &seq) {
        int c;
        if (last_ == 0) {                                          // jump to the next header character
            while ((c = getc_()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return false;
            last_ = c;
        }
        seq.clear();
        // name up to the first white space; the rest of the header line is the comment
        bool any = false;
        while ((c = getc_()) != -1) { any = true; if (isspace(c)) break; }
        if (!any) return false;                                    // EOF right after a header character
        if (c != '\n' && c != -1) skip_line_();
        while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;                               // empty line
            seq.push_back((char)c);
            rest_of_line_(seq);
        }
        if (c == '>' || c == '@') last_ = c;
        if (c != '+') return finish_(c);                           // FASTA
        while ((c = getc_()) != -1 && c != '\n') {}                // rest of the '+' line
        if (c == -1) return false;                                 // no quality string
        size_t q = 0;
        qual_.clear();
        for (;;) {                                                 // quality lines until they cover the sequence
            if (at_eof_()) break;
            rest_of_line_(qual_);
            q = qual_.size();
            if (q >= seq.size()) break;
        }
        last_ = 0;
        return q == seq.size();
    }

  private:
    bool finish_(int c) {
        if (c == -1) last_ = -2;                                   // the file ended with this record: nothing follows
        return true;
    }
    bool fill_() {
        if (eof_) return false;
        const int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (n <= 0) { eof_ = true; return false; }
        beg_ = 0; end_ = (size_t)n;
        return true;
    }
    bool at_eof_() { return beg_ >= end_ && !fill_(); }
    int getc_() {
        if (last_ == -2) return -1;
        if (beg_ >= end_ && !fill_()) return -1;
        return (unsigned char)buf_[beg_++];
    }
    void skip_line_() { int c; while ((c = getc_()) != -1 && c != '\n') {} }
    // appends up to (not including) the next '\n'; a trailing '\r' of the accumulated string is dropped (kseq.h:132)
    void rest_of_line_(std::string &s) {
        for (;;) {
            if (beg_ >= end_ && !fill_()) break;
            const char *p = buf_.data() + beg_;
            const char *nl = (const char *)memchr(p, '\n', end_ - beg_);
            const size_t n = nl ? (size_t)(nl - p) : end_ - beg_;
            s.append(p, n);
            beg_ += n + (nl ? 1 : 0);
            if (nl) break;
        }
        if (s.size() > 1 && s.back() == '\r') s.pop_back();
    }

    gzFile gz_ = nullptr;
    std::vector<char> buf_;
    size_t beg_ = 0, end_ = 0;
    bool eof_ = false;
    int last_ = 0;                                                 // header character read ahead; -2: the file has ended
    std::string qual_;
};

