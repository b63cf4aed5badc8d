// fuse_main.cpp -- `fuse in.bam out.fasta` on top of the C ABI (/root/reference/doc/FUSE.md:17-32):
// consensus of an alignment, in-frame insertions at a distance from each other included, major deletions
// removed, one FASTA record per input.  The thresholds are the restatement's choices U6-U8 (options).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>
#include "host_common.hpp"

#define CK(h, call)                                                                            \
    do {                                                                                       \
        int rc_ = (call);                                                                      \
        if (rc_ != MS_OK) mshost::die(std::string(#call) + " failed: " + ms_last_error(h));     \
    } while (0)

int main(int argc, char** argv) {
    ms_fuse_params prm;
    ms_fuse_params_default(&prm);
    int device = 0;
    mshost::QvFilter qv;
    qv.threshold = 0;  // fuse takes every base
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> std::string {
            if (i + 1 >= argc) mshost::die(std::string("option ") + name + " needs a value");
            return argv[++i];
        };
        if (a == "-h" || a == "--help") {
            puts("Usage: fuse [--min-coverage n] [--ins-fraction f] [--ins-distance d] [--device n] <in.bam> <out.fasta>");
            return 0;
        } else if (a == "--version") { puts("minorseq_b200 fuse 0.1.0 (B200-native restatement; not PacBio fuse)"); return 0; }
        else if (a == "--min-coverage") prm.min_coverage = atoi(need("--min-coverage").c_str());
        else if (a == "--ins-fraction") prm.ins_fraction = atof(need("--ins-fraction").c_str());
        else if (a == "--ins-distance") prm.ins_distance = atoi(need("--ins-distance").c_str());
        else if (a == "--device") device = atoi(need("--device").c_str());
        else if (!a.empty() && a[0] == '-') mshost::die("unknown option " + a);
        else pos.push_back(a);
    }
    if (pos.size() != 2) { puts("Usage: fuse <in.bam> <out.fasta>"); return 1; }
    try {
        // the CUDA context comes up (~0.5 s) on this thread while a helper thread inflates and indexes the BAM
        ms_handle* h = nullptr;
        mshost::Alignments aln;
        mshost::load_alignments_overlapped(pos[0], qv, false, true, aln, [&] {
            if (ms_create(device, &h) != MS_OK) mshost::die(ms_last_error(nullptr));   // there is no CPU path
        });
        if (aln.nreads == 0) mshost::die("no primary or supplementary alignments in " + pos[0]);
        CK(h, ms_set_layout(h, aln.L, nullptr));
        CK(h, ms_pileup_host(h, aln.rows, aln.nreads, nullptr));
        std::string seq(static_cast<size_t>(aln.L) + aln.ins_pool.size() + 16, '\0');
        int64_t len = 0;
        CK(h, ms_fuse(h, &prm, aln.ins_col.data(), aln.ins_off.data(), aln.ins_len.data(), static_cast<int64_t>(aln.ins_col.size()),
                      aln.ins_pool.data(), static_cast<int64_t>(aln.ins_pool.size()), &seq[0], static_cast<int64_t>(seq.size()), &len));
        seq.resize(static_cast<size_t>(len));
        std::string stem = pos[0];
        const size_t sl = stem.find_last_of('/');
        if (sl != std::string::npos) stem = stem.substr(sl + 1);
        std::ofstream f(pos[1]);
        if (!f) mshost::die("cannot write " + pos[1]);
        f << ">" << stem << "|fuse|" << aln.ref_name << "\n";
        for (size_t i = 0; i < seq.size(); i += 70) f << seq.substr(i, 70) << "\n";
        fprintf(stderr, "fuse: %lld reads, consensus length %lld\n", static_cast<long long>(aln.nreads), static_cast<long long>(len));
        f.close();
        fflush(nullptr);
        if (!getenv("MS_FULL_TEARDOWN")) _exit(0);
        ms_destroy(h);
    } catch (const std::exception& e) {
        mshost::die(e.what());
    }
    return 0;
}
