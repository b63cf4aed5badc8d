// packed2bam_main.cpp -- fixture tool: turns packed rows (raw little-endian u32, ms_row_words(L) per read) into an
// aligned BAM with =/X/D/N CIGARs against a reference given as a text file of ACGT.  Used to time the
// from-BAM region (R3) of juliet/fuse on synthetic data; 'N' states become base N under an X op, the
// insertion flag becomes a 1-base insertion "A".  Host tool, no GPU.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include "bgzf_bam.hpp"

int main(int argc, char** argv) {
    if (argc != 6) { fputs("Usage: packed2bam <packed.bin> <L> <R> <ref.txt> <out.bam>\n", stderr); return 1; }
    const int L = atoi(argv[2]);
    const long R = atol(argv[3]);
    const int nblk = (L + 31) / 32;
    std::ifstream rf(argv[4]);
    std::string ref;
    rf >> ref;
    if (static_cast<int>(ref.size()) < L) { fputs("reference shorter than L\n", stderr); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fputs("cannot open packed file\n", stderr); return 1; }
    try {
    msbam::BamWriter w(argv[5], "@HD\tVN:1.5\tSO:unknown\n@SQ\tSN:synthetic_ref\tLN:" + std::to_string(L) + "\n", {{"synthetic_ref", L}});
    std::vector<uint32_t> row(4 * nblk);
    static const char B[] = "ACGT";
    for (long r = 0; r < R; ++r) {
        if (fread(row.data(), 4, row.size(), f) != row.size()) { fputs("short packed file\n", stderr); return 1; }
        msbam::Record rec;
        rec.ref_id = 0; rec.flag = 0; rec.mapq = 60;
        rec.name = "m/" + std::to_string(r) + "/ccs";
        int first = -1, last = -1;
        std::vector<uint8_t> st(L), ins(L);
        for (int c = 0; c < L; ++c) {
            const uint32_t* q = &row[4 * (c >> 5)];
            const int sh = c & 31;
            st[c] = ((q[0] >> sh) & 1) | (((q[1] >> sh) & 1) << 1) | (((q[2] >> sh) & 1) << 2);
            ins[c] = (q[3] >> sh) & 1;
            if (st[c] != 7) { if (first < 0) first = c; last = c; }
        }
        if (first < 0) continue;
        rec.pos = first;
        auto push = [&](uint32_t op, uint32_t n) {
            if (!rec.cigar.empty() && (rec.cigar.back() & 15u) == op) rec.cigar.back() += n << 4;
            else rec.cigar.push_back((n << 4) | op);
        };
        for (int c = first; c <= last; ++c) {
            if (st[c] == 7) push(3, 1);
            else if (st[c] == 4) push(2, 1);
            else if (st[c] == 5) { push(8, 1); rec.seq += 'N'; }
            else { push(B[st[c]] == ref[c] ? 7 : 8, 1); rec.seq += B[st[c]]; }
            if (ins[c]) { push(1, 1); rec.seq += 'A'; }
        }
        w.write(rec);
    }
    w.close();
    fclose(f);
    } catch (const std::exception& e) {     // cannot create / write error on the output
        fprintf(stderr, "packed2bam: %s\n", e.what());
        return 1;
    }
    return 0;
}
