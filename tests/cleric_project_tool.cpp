// Test-only driver for minorseq_b200/host/cleric.hpp: reads "ops b pos cigar seq" lines, prints the projected
// alignment.  Lets the CPU suite check the product's per-read projection against the oracle without a GPU
// (the path is given; in the product it comes from ms_align_refs).
#include <cstdio>
#include <iostream>
#include <sstream>
#include "../minorseq_b200/host/cleric.hpp"

int main() {
    std::string line;
    static const char* opc = "MIDNSHP=X";
    while (std::getline(std::cin, line)) {
        std::istringstream ls(line);
        std::string ops, b, cig, seq;
        int pos;
        if (!(ls >> ops >> b >> pos >> cig >> seq)) continue;
        if (ops == "-") ops.clear();
        if (b == "-") b.clear();
        msbam::Record r;
        r.pos = pos;
        r.seq = seq == "-" ? "" : seq;
        for (size_t i = 0; i < cig.size();) {
            size_t j = i;
            while (j < cig.size() && isdigit(static_cast<unsigned char>(cig[j]))) ++j;
            const uint32_t len = static_cast<uint32_t>(std::stoul(cig.substr(i, j - i)));
            const char* p = strchr(opc, cig[j]);
            r.cigar.push_back((len << 4) | static_cast<uint32_t>(p - opc));
            i = j + 1;
        }
        const mscleric::Path path(ops);
        switch (mscleric::project_read(path, b, r)) {
        case mscleric::Projected::Ok: {
            printf("ok %d ", r.pos);
            for (uint32_t c : r.cigar) printf("%u%c", c >> 4, opc[c & 15u]);
            printf("\n");
            break;
        }
        case mscleric::Projected::Unmapped: printf("unmapped\n"); break;
        case mscleric::Projected::Unsupported: printf("unsupported\n"); break;
        case mscleric::Projected::Inconsistent: printf("inconsistent\n"); break;
        }
    }
    return 0;
}
