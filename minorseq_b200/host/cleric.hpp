// cleric.hpp -- host side of `cleric`: FASTA input and the transitive re-expression of one read's alignment.
//
// "Current scope of Cleric is converting a given alignment to a different reference.  This is done by aligning the
// original and target reference sequences.  A transitive alignment is used to generate the new alignment."
// (/root/reference/doc/CLERIC.md:19-23).  The reference-to-reference alignment comes from the GPU (ms_align_refs,
// csrc/nw.cu); what is here is per-read CIGAR bookkeeping, the same kind of host work as CIGAR expansion.
// Rules (restatement choice U13, oracle/ms_oracle.h mso_project_read): a base over an original column that has a
// partner on the target becomes '=' / 'X' against the target, over a column without partner an insertion; target
// columns without partner inside the read become deletions; insertions and clips stay; the alignment may not begin or
// end with 'I' (becomes a clip) or 'D' (dropped); 'M', 'N', 'P' are refused (doc/CLERIC.md:14-15).
#pragma once
#include <cctype>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "bgzf_bam.hpp"

namespace mscleric {

struct Fasta { std::string name, seq; };

inline std::vector<Fasta> read_fasta(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<Fasta> out;
    std::string line;
    while (std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            Fasta r;
            size_t e = 1;
            while (e < line.size() && !isspace(static_cast<unsigned char>(line[e]))) ++e;
            r.name = line.substr(1, e - 1);
            out.push_back(r);
        } else {
            if (out.empty()) throw std::runtime_error(path + ": sequence data before the first '>' line");
            for (char c : line) out.back().seq.push_back(static_cast<char>(toupper(static_cast<unsigned char>(c))));
        }
    }
    return out;
}

// The reference-to-reference path with the lookups a read needs: for every original column, the op that consumes it
// and how many target columns lie before that op.
struct Path {
    std::string ops;                 // 'M' both advance, 'D' only the original advances, 'I' only the target advances
    std::vector<int32_t> op_of_a;    // [la]   index into ops
    std::vector<int32_t> b_before;   // [nops] target columns consumed before op x
    int32_t la = 0, lb = 0;
    explicit Path(std::string o) : ops(std::move(o)) {
        b_before.resize(ops.size());
        for (size_t x = 0; x < ops.size(); ++x) {
            b_before[x] = lb;
            if (ops[x] != 'I') { op_of_a.push_back(static_cast<int32_t>(x)); ++la; }
            if (ops[x] != 'D') ++lb;
        }
    }
};

enum class Projected { Ok, Unmapped, Unsupported, Inconsistent };

namespace detail {
constexpr uint32_t opI = 1, opD = 2, opS = 4, opH = 5, opEQ = 7, opX = 8;
inline void push(std::vector<uint32_t>& cig, uint32_t op, uint32_t len) {
    if (len == 0) return;
    if (!cig.empty() && (cig.back() & 15u) == op) cig.back() += len << 4;
    else cig.push_back((len << 4) | op);
}
}  // namespace detail

// Re-expresses rec (aligned to the original at rec.pos with rec.cigar) against the target sequence b.
inline Projected project_read(const Path& path, const std::string& b, msbam::Record& rec) {
    using namespace detail;
    const std::vector<uint32_t>& in = rec.cigar;
    size_t c0 = 0, c1 = in.size();
    uint32_t hard0 = 0, soft0 = 0, hard1 = 0, soft1 = 0;
    while (c0 < c1 && ((in[c0] & 15u) == opS || (in[c0] & 15u) == opH)) { ((in[c0] & 15u) == opS ? soft0 : hard0) += in[c0] >> 4; ++c0; }
    while (c1 > c0 && ((in[c1 - 1] & 15u) == opS || (in[c1 - 1] & 15u) == opH)) { ((in[c1 - 1] & 15u) == opS ? soft1 : hard1) += in[c1 - 1] >> 4; --c1; }
    if (rec.pos < 0 || rec.pos > path.la) return Projected::Inconsistent;
    // body: ops against the target, one event at a time; `first_b` = target column of the first reference-consuming op
    std::vector<uint32_t> body;
    std::vector<int32_t> body_b;     // target column at which each body op starts (-1 for insertions)
    auto emit = [&](uint32_t op, int32_t bj) {
        if (!body.empty() && (body.back() & 15u) == op) { body.back() += 1u << 4; return; }
        body.push_back((1u << 4) | op);
        body_b.push_back(bj);
    };
    size_t q = soft0;                // next query base
    int32_t ai = rec.pos;            // next original column
    bool inside = false;
    for (size_t c = c0; c < c1; ++c) {
        const uint32_t op = in[c] & 15u, len = in[c] >> 4;
        if (op == opI) {
            if (!body.empty() && (body.back() & 15u) == opI) body.back() += len << 4;
            else { body.push_back((len << 4) | opI); body_b.push_back(-1); }
            q += len;
            continue;
        }
        if (op != opEQ && op != opX && op != opD) return Projected::Unsupported;
        for (uint32_t t = 0; t < len; ++t, ++ai) {
            if (ai >= path.la) return Projected::Inconsistent;
            const int32_t x = path.op_of_a[ai];
            // target-only columns between the previous original column and this one lie inside the read
            if (inside) {
                int32_t y = x;
                while (y > 0 && path.ops[y - 1] == 'I') --y;
                for (int32_t z = y; z < x; ++z) emit(opD, path.b_before[z]);
            }
            inside = true;
            if (path.ops[x] == 'M') {
                const int32_t bj = path.b_before[x];
                if (op == opD) emit(opD, bj);
                else {
                    if (q >= rec.seq.size() || bj >= static_cast<int32_t>(b.size())) return Projected::Inconsistent;
                    emit(rec.seq[q] == b[bj] ? opEQ : opX, bj);
                    ++q;
                }
            } else if (op != opD) {   // the original column has no partner: the base is extra relative to the target
                emit(opI, -1);
                ++q;
            }
        }
    }
    if (q + soft1 != rec.seq.size()) return Projected::Inconsistent;
    size_t lo = 0, hi = body.size();
    while (lo < hi && ((body[lo] & 15u) == opD || (body[lo] & 15u) == opI)) { if ((body[lo] & 15u) == opI) soft0 += body[lo] >> 4; ++lo; }
    while (hi > lo && ((body[hi - 1] & 15u) == opD || (body[hi - 1] & 15u) == opI)) { if ((body[hi - 1] & 15u) == opI) soft1 += body[hi - 1] >> 4; --hi; }
    if (lo == hi) return Projected::Unmapped;
    std::vector<uint32_t> out;
    push(out, opH, hard0);
    push(out, opS, soft0);
    for (size_t k = lo; k < hi; ++k) push(out, body[k] & 15u, body[k] >> 4);
    push(out, opS, soft1);
    push(out, opH, hard1);
    rec.pos = body_b[lo];
    rec.cigar.swap(out);
    return Projected::Ok;
}

}  // namespace mscleric
