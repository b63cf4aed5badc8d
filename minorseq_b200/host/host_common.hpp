// host_common.hpp -- what juliet and fuse share in front of the C ABI: the BAM loop.
//
// For every record: admission filter (/root/reference/doc/JULIET.md:58), CIGAR walk into one packed row
// (:49-58, `M` rejected), rich-QV base filter -> 'N' (:256-259,:273-276), optional insertion events (fuse).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "../../include/minorseq_b200.h"
#include "bgzf_bam.hpp"
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define MSHOST_RANGE(name) mshost::NvtxRange nvtx_range_(name)
namespace mshost { struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } }; }
#endif
#endif
#ifndef MSHOST_RANGE
#define MSHOST_RANGE(name) do {} while (0)     // built without the CUDA toolkit's headers: no NVTX ranges on the host stages
#endif

namespace mshost {

struct QvFilter {                      // restatement choice U4: which per-base tracks, which threshold
    std::vector<std::string> tags{"dq", "iq", "sq"};
    int threshold = 20;                // a base whose QV in any present track is below this becomes 'N'; 0 = off

    // mask[i] = 1 for every base of rec (SEQ orientation) whose value in any present per-base track is below the
    // threshold; returns false when no usable track exists (doc/JULIET.md:256-259,273-276: no filter then).  Tracks
    // are stored in native orientation, SEQ in reference orientation.  Works on the raw tag bytes: this runs once per
    // read and track on the decode path, which is what bounds the tools end to end.
    bool apply(const msbam::Record& rec, std::vector<uint8_t>& mask) const {
        if (threshold <= 0) return false;
        const size_t n = rec.seq.size();
        const bool rev = rec.flag & 0x10;
        bool any = false;
        for (const std::string& tg : tags) {
            const uint8_t* t = rec.find_tag(tg.c_str());
            if (!t || n == 0) continue;
            auto track = [&](auto value) {     // value(i) = QV of native base i; branch-free so that the loops vectorise
                if (!any) mask.assign(n, 0);
                any = true;
                uint8_t* m = mask.data();
                const int thr = threshold;
                if (rev) { for (size_t i = 0; i < n; ++i) m[n - 1 - i] |= static_cast<uint8_t>(value(i) < thr); }
                else { for (size_t i = 0; i < n; ++i) m[i] |= static_cast<uint8_t>(value(i) < thr); }
            };
            if (*t == 'Z') {
                const uint8_t* p = t + 1;
                if (strnlen(reinterpret_cast<const char*>(p), n + 1) != n) continue;
                track([&](size_t i) { return static_cast<int>(p[i]) - 33; });
            } else if (*t == 'B') {
                if (msbam::detail::u32(t + 2) != n) continue;
                const uint8_t* p = t + 6;
                switch (t[1]) {
                case 'c': track([&](size_t i) { return static_cast<int>(static_cast<int8_t>(p[i])); }); break;
                case 'C': track([&](size_t i) { return static_cast<int>(p[i]); }); break;
                case 's': track([&](size_t i) { return static_cast<int>(static_cast<int16_t>(msbam::detail::u16(p + 2 * i))); }); break;
                case 'S': track([&](size_t i) { return static_cast<int>(msbam::detail::u16(p + 2 * i)); }); break;
                default: break;
                }
            }
        }
        return any;
    }
};

struct Alignments {
    int32_t L = 0;
    int32_t ref_id = -1;
    std::string ref_name;
    int64_t nreads = 0, nskipped = 0;
    uint32_t* rows = nullptr;          // pinned, nreads * ms_row_words(L)
    size_t cap_rows = 0;
    std::vector<std::string> names;
    // insertion events (only when want_insertions)
    std::vector<int32_t> ins_col, ins_len;
    std::vector<int64_t> ins_off;
    std::string ins_pool;
    ~Alignments() { ms_free_pinned(rows); }
};

inline void die(const std::string& m) {
    fprintf(stderr, "ERROR: %s\n", m.c_str());
    exit(1);
}

// f(rank) for rank 0..n-1: on this thread when n == 1, else one thread per rank (one rank per GPU; the ranks meet inside the
// library's collectives, so they have to run concurrently)
template <class F>
inline void run_ranks(int n, F f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int r = 0; r < n; ++r) th.emplace_back([&f, r] { f(r); });
    for (std::thread& t : th) t.join();
}

// One record -> one packed row (+ optional insertion events appended to the caller's vectors).
inline void expand_record(const msbam::Record& rec, const QvFilter& qv, int32_t L, uint32_t* row, bool want_insertions,
                          std::vector<uint8_t>& mask, std::vector<int32_t>& ic, std::vector<int64_t>& io, std::vector<int32_t>& il,
                          std::string& pool, std::vector<int32_t>& out_col, std::vector<int32_t>& out_len, std::vector<int64_t>& out_off,
                          std::string& out_pool) {
    const uint8_t* maskp = qv.apply(rec, mask) ? mask.data() : nullptr;   // rich-QV filter
    int64_t ni = 0, pu = 0;
    int rc;
    for (;;) {
        ni = 0; pu = 0;
        rc = ms_expand_cigar(rec.cigar.data(), static_cast<int32_t>(rec.cigar.size()), rec.pos, rec.seq.data(), maskp,
                             static_cast<int32_t>(rec.seq.size()), L, row, want_insertions ? ic.data() : nullptr, io.data(), il.data(),
                             static_cast<int64_t>(ic.size()), &ni, want_insertions ? &pool[0] : nullptr, static_cast<int64_t>(pool.size()), &pu);
        if (rc != MS_ERR_CAPACITY) break;
        ic.resize(ic.size() * 2); io.resize(io.size() * 2); il.resize(il.size() * 2); pool.resize(pool.size() * 2);
    }
    // thrown, not die(): this runs on worker threads (an exit() from there would race the other workers)
    if (rc == MS_ERR_FORMAT) throw std::runtime_error("record " + rec.name + ": BAM files have to be PacBio-compliant, cigar M is forbidden");
    if (rc != MS_OK) throw std::runtime_error("record " + rec.name + ": cannot expand CIGAR");
    for (int64_t i = 0; i < ni; ++i) {
        out_col.push_back(ic[i]);
        out_len.push_back(il[i]);
        out_off.push_back(static_cast<int64_t>(out_pool.size()) + io[i]);
    }
    out_pool.append(pool.data(), static_cast<size_t>(pu));
}

// Whole-file loader in two phases so that the first one (no CUDA involved) can overlap the creation of the
// CUDA context: decode() = parallel BGZF inflate + record index + admission filter; expand() = all host
// threads parse and expand the admitted records straight into one pinned row buffer.
struct Decoded {
    msbam::Bytes u;
    msbam::BamIndexed bx;
    std::vector<size_t> keep;
    unsigned nt = 1;
};

inline void decode_alignments(const std::string& path, Decoded& d, Alignments& out) {
    MSHOST_RANGE("BAM inflate + index + admission");
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("MS_HOST_THREADS")) nt = static_cast<unsigned>(atoi(e));
    if (nt < 1) nt = 1;
    d.nt = nt;
    d.u = msbam::inflate_file(path, nt);
    d.bx = msbam::index_stream(d.u);
    // admission (doc/JULIET.md:58) + one reference per run
    for (size_t i = 0; i < d.bx.records.size(); ++i) {
        const uint8_t* p = d.u.data() + d.bx.records[i].first;
        if (d.bx.records[i].second < 32) throw std::runtime_error("BAM record too small");
        const int32_t ref_id = msbam::detail::i32(p);
        const uint16_t flag = msbam::detail::u16(p + 14);
        if (!ms_read_admitted(flag) || ref_id < 0) { ++out.nskipped; continue; }
        if (out.ref_id < 0) {
            out.ref_id = ref_id;
            if (ref_id >= static_cast<int32_t>(d.bx.refs.size())) throw std::runtime_error("record refers to an unknown reference");
            out.L = d.bx.refs[ref_id].length;
            out.ref_name = d.bx.refs[ref_id].name;
            if (out.L < 3) throw std::runtime_error("reference too short");
        }
        if (ref_id != out.ref_id) { ++out.nskipped; continue; }
        d.keep.push_back(i);
    }
    out.nreads = static_cast<int64_t>(d.keep.size());
}

inline void expand_alignments(const Decoded& d, const QvFilter& qv, bool want_names, bool want_insertions, Alignments& out) {
    MSHOST_RANGE("CIGAR expansion + QV filter");
    const msbam::Bytes& u = d.u;
    const msbam::BamIndexed& bx = d.bx;
    const std::vector<size_t>& keep = d.keep;
    unsigned nt = d.nt;
    if (keep.empty()) return;
    const int32_t rw = ms_row_words(out.L);
    out.rows = static_cast<uint32_t*>(ms_alloc_pinned(keep.size() * rw * sizeof(uint32_t)));
    if (!out.rows) throw std::runtime_error("cannot allocate pinned host memory for the packed rows");
    out.cap_rows = keep.size();
    if (want_names) out.names.resize(keep.size());
    struct Part { std::vector<int32_t> col, len; std::vector<int64_t> off; std::string pool; };
    nt = static_cast<unsigned>(std::min<size_t>(nt, keep.size()));
    std::vector<Part> parts(nt);
    std::vector<std::string> errs(nt);      // a corrupt record throws on its worker: carried to the caller after the join
    auto work = [&](unsigned t) { try {
        const size_t i0 = keep.size() * t / nt, i1 = keep.size() * (t + 1) / nt;
        msbam::Record rec;
        std::vector<uint8_t> mask;
        std::vector<int32_t> ic(4096), il(4096);
        std::vector<int64_t> io(4096);
        std::string pool(1 << 16, '\0');
        for (size_t k = i0; k < i1; ++k) {
            const auto& rr = bx.records[keep[k]];
            msbam::BamReader::parse_record(u.data() + rr.first, rr.second, rec);
            expand_record(rec, qv, out.L, out.rows + k * rw, want_insertions, mask, ic, io, il, pool, parts[t].col, parts[t].len, parts[t].off, parts[t].pool);
            if (want_names) out.names[k] = rec.name;
        }
    } catch (const std::exception& e) { errs[t] = e.what(); } };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (const std::string& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
    for (const Part& p : parts) {   // read order is preserved: thread t holds a contiguous range
        const int64_t base = static_cast<int64_t>(out.ins_pool.size());
        out.ins_col.insert(out.ins_col.end(), p.col.begin(), p.col.end());
        out.ins_len.insert(out.ins_len.end(), p.len.begin(), p.len.end());
        for (int64_t o : p.off) out.ins_off.push_back(base + o);
        out.ins_pool += p.pool;
    }
}

// decode on a helper thread while the caller brings up the CUDA context, then expand
template <class CreateFn>
inline void load_alignments_overlapped(const std::string& path, const QvFilter& qv, bool want_names, bool want_insertions, Alignments& out,
                                       CreateFn create_context) {
    Decoded d;
    std::string err;
    if (getenv("MS_SERIAL_LOAD")) {   // A/B switch for the overlap (tools/from_bam_timing.py)
        create_context();
        decode_alignments(path, d, out);
        expand_alignments(d, qv, want_names, want_insertions, out);
        return;
    }
    std::thread dec([&] { try { decode_alignments(path, d, out); } catch (const std::exception& e) { err = e.what(); } });
    create_context();      // dies with the CUDA error when there is no usable GPU: that message wins
    dec.join();
    if (!err.empty()) throw std::runtime_error(err);
    expand_alignments(d, qv, want_names, want_insertions, out);
}

}  // namespace mshost
