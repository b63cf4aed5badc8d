// mixdata_main.cpp -- in-silico mixing of clonal alignments, the rule of minorseq's mixdata script
// (/root/reference/doc/MIXDATA.md:9-22): "the first file is the major", every other file contributes
// PERCENTAGE % of COVERAGE reads.  Same environment variables (COVERAGE=3000 PERCENTAGE=1 OUTPUT_PREFIX=mix);
// reads are taken from the head of each input (the script shuffles with samtools; here `--seed` picks a
// deterministic pseudo-random subset instead).  Writes <OUTPUT_PREFIX>.bam.  Host tool, no GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "bgzf_bam.hpp"

int main(int argc, char** argv) {
    const char* cov_s = getenv("COVERAGE");
    const char* per_s = getenv("PERCENTAGE");
    const char* pre_s = getenv("OUTPUT_PREFIX");
    const long coverage = cov_s ? atol(cov_s) : 3000;
    const double percentage = per_s ? atof(per_s) : 1.0;
    const std::string prefix = pre_s ? pre_s : "mix";
    unsigned seed = 42;
    std::vector<std::string> in;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--seed" && i + 1 < argc) seed = static_cast<unsigned>(atol(argv[++i]));
        else if (a == "-h" || a == "--help") {
            puts("Usage: COVERAGE=3000 PERCENTAGE=1 OUTPUT_PREFIX=mix mixdata [--seed n] major.bam minor1.bam [minor2.bam ...]");
            return 0;
        } else in.push_back(a);
    }
    if (in.size() < 2) { fputs("mixdata: need a major and at least one minor BAM\n", stderr); return 1; }
    try {
        const long per_minor = static_cast<long>(coverage * percentage / 100.0 + 0.5);
        const long major = coverage - per_minor * static_cast<long>(in.size() - 1);
        if (major <= 0) { fputs("mixdata: minors exceed the coverage\n", stderr); return 1; }
        std::vector<msbam::Record> out;
        std::vector<msbam::RefSeq> refs;
        std::string text;
        std::mt19937 rng(seed);
        for (size_t f = 0; f < in.size(); ++f) {
            msbam::BamReader rd(in[f]);
            if (f == 0) { refs = rd.refs(); text = rd.header_text(); }
            std::vector<msbam::Record> all;
            msbam::Record r;
            while (rd.next(r)) all.push_back(r);
            const size_t want = static_cast<size_t>(f == 0 ? major : per_minor);
            if (all.size() < want) { fprintf(stderr, "mixdata: %s has only %zu reads, %zu needed\n", in[f].c_str(), all.size(), want); return 1; }
            std::shuffle(all.begin(), all.end(), rng);
            out.insert(out.end(), all.begin(), all.begin() + static_cast<long>(want));
        }
        std::stable_sort(out.begin(), out.end(), [](const msbam::Record& a, const msbam::Record& b) { return a.pos < b.pos; });
        msbam::BamWriter w(prefix + ".bam", text, refs);
        for (const msbam::Record& r : out) w.write(r);
        w.close();
        fprintf(stderr, "mixdata: %zu reads -> %s.bam (%ld major, %ld per minor)\n", out.size(), prefix.c_str(), major, per_minor);
    } catch (const std::exception& e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
