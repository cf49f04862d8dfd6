// prot_kmers.h -- host rules of `findstart`: residue codes, the packed protein k-mer key and the k-mers of an aligned
// reference protein over its model columns (reference prot_kmer.h:26-43, prot_kmer_generator.h:58-137).  Shared by the
// findstart driver and the CPU test harness (tests/cpu/seqtools_host.cpp).
#pragma once
#include <ctype.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace mgta_host {

// residue codes of ProtKmer::setUp (prot_kmer.h:26-43): ARNDCQEGHI = 0..9, LKMFPSTWYV = 10..19, * = 20; 31 = not a residue
inline int residue_code(unsigned char c) {
    static int map[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; ++i) map[i] = 31;
        const char *up = "ARNDCQEGHILKMFPSTWYV", *lo = "arndcqeghilkmfpstwyv";
        for (int i = 0; i < 20; ++i) { map[(unsigned char)up[i]] = i; map[(unsigned char)lo[i]] = i; }
        map[(unsigned char)'*'] = 20;
        init = true;
    }
    return c < 127 ? map[c] : 31;
}

struct ModelKmer {
    uint64_t w[2];
    int model_pos;
};

// k residue codes -> the two-word key mgta_find_seeds uses (residues 0..11 in w[0], first residue most significant; 12.. in w[1])
inline void pack_key(const int *codes, int k, uint64_t w[2]) {
    w[0] = w[1] = 0;
    for (int i = 0; i < k; ++i) {
        if (i < 12) w[0] = (w[0] << 5) | (uint64_t)codes[i];
        else w[1] = (w[1] << 5) | (uint64_t)codes[i];
    }
}

inline std::string unpack_key(const uint64_t w[2], int k) {               // ProtKmer::decodePacked: lower-case residues
    static const char *lo = "arndcqeghilkmfpstwyv*";
    std::string s(k, '?');
    const int n1 = k > 12 ? k - 12 : 0, n0 = k - n1;
    for (int i = 0; i < n0; ++i) s[i] = lo[(w[0] >> (5 * (n0 - 1 - i))) & 31];
    for (int i = 0; i < n1; ++i) s[12 + i] = lo[(w[1] >> (5 * (n1 - 1 - i))) & 31];
    return s;
}

// The k-mers of one aligned reference protein over its model columns, with the model position of each
// (prot_kmer_generator.h:58-137 with model_only = true): lower case (insert states), '-', 'X', 'x' break the window ('-' and
// 'X' are model columns, so they advance the position); '.', '*' and anything that is not a residue are skipped without
// breaking it; a window of k residues yields a k-mer at position (columns consumed so far + 1 - k).
inline void model_kmers_of(const std::string &seq, int k, std::vector<ModelKmer> &out) {
    std::vector<int> win;                                          // the residues of the current unbroken run
    int position = 1;
    for (unsigned char base : seq) {
        if (islower(base) || base == '-' || base == 'X' || base == 'x') {
            if (base == '-' || base == 'X') ++position;
            win.clear();
            continue;
        }
        const int code = residue_code(base);
        if (base == '.' || code == 31 || base == '*') continue;
        win.push_back(code);
        ++position;
        if ((int)win.size() >= k) {
            ModelKmer m;
            pack_key(win.data() + win.size() - k, k, m.w);
            m.model_pos = position - k;
            out.push_back(m);
        }
    }
}

}  // namespace mgta_host
