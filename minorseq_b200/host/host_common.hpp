// host_common.hpp -- what juliet and fuse share in front of the C ABI: the BAM loop.
//
// For every record: admission filter (/root/reference/doc/JULIET.md:58), CIGAR walk into one packed row
// (:49-58, `M` rejected), rich-QV base filter -> 'N' (:256-259,:273-276), optional insertion events (fuse).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/minorseq_b200.h"
#include "bgzf_bam.hpp"

namespace mshost {

struct QvFilter {                      // restatement choice U4: which per-base tracks, which threshold
    std::vector<std::string> tags{"dq", "iq", "sq"};
    int threshold = 20;                // a base whose QV in any present track is below this becomes 'N'; 0 = off
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

inline void load_alignments(const std::string& path, const QvFilter& qv, bool want_names, bool want_insertions, Alignments& out) {
    msbam::BamReader bam(path);
    msbam::Record rec;
    std::vector<uint8_t> mask;
    std::vector<int32_t> ic(4096);
    std::vector<int64_t> io(4096);
    std::vector<int32_t> il(4096);
    std::string pool(1 << 16, '\0');
    while (bam.next(rec)) {
        if (!ms_read_admitted(rec.flag) || rec.ref_id < 0) { ++out.nskipped; continue; }
        if (out.ref_id < 0) {
            out.ref_id = rec.ref_id;
            if (rec.ref_id >= static_cast<int32_t>(bam.refs().size())) die("record refers to an unknown reference");
            out.L = bam.refs()[rec.ref_id].length;
            out.ref_name = bam.refs()[rec.ref_id].name;
            if (out.L < 3) die("reference too short");
        }
        if (rec.ref_id != out.ref_id) { ++out.nskipped; continue; }   // one reference per run
        const int32_t rw = ms_row_words(out.L);
        if (static_cast<size_t>(out.nreads + 1) > out.cap_rows) {
            const size_t ncap = out.cap_rows ? out.cap_rows * 2 : 4096;
            uint32_t* nr = static_cast<uint32_t*>(ms_alloc_pinned(ncap * rw * sizeof(uint32_t)));
            if (!nr) die("out of pinned host memory");
            if (out.rows) memcpy(nr, out.rows, static_cast<size_t>(out.nreads) * rw * sizeof(uint32_t));
            ms_free_pinned(out.rows);
            out.rows = nr; out.cap_rows = ncap;
        }
        // rich-QV filter: tracks are stored in native orientation, SEQ in reference orientation
        const uint8_t* maskp = nullptr;
        if (qv.threshold > 0) {
            bool any = false;
            mask.assign(rec.seq.size(), 0);
            for (const std::string& t : qv.tags) {
                std::vector<int> track = rec.tag_per_base(t.c_str());
                if (track.size() != rec.seq.size()) continue;
                any = true;
                const bool rev = rec.flag & 0x10;
                for (size_t i = 0; i < track.size(); ++i)
                    if (track[i] < qv.threshold) mask[rev ? track.size() - 1 - i : i] = 1;
            }
            if (any) maskp = mask.data();
        }
        int64_t ni = 0, pu = 0;
        int rc;
        for (;;) {
            ni = 0; pu = 0;
            rc = ms_expand_cigar(rec.cigar.data(), static_cast<int32_t>(rec.cigar.size()), rec.pos, rec.seq.data(), maskp,
                                 static_cast<int32_t>(rec.seq.size()), out.L, out.rows + static_cast<size_t>(out.nreads) * rw,
                                 want_insertions ? ic.data() : nullptr, io.data(), il.data(), static_cast<int64_t>(ic.size()), &ni,
                                 want_insertions ? &pool[0] : nullptr, static_cast<int64_t>(pool.size()), &pu);
            if (rc != MS_ERR_CAPACITY) break;
            ic.resize(ic.size() * 2); io.resize(io.size() * 2); il.resize(il.size() * 2); pool.resize(pool.size() * 2);
        }
        if (rc == MS_ERR_FORMAT) die("record " + rec.name + ": BAM files have to be PacBio-compliant, cigar M is forbidden");
        if (rc != MS_OK) die("record " + rec.name + ": cannot expand CIGAR");
        for (int64_t i = 0; i < ni; ++i) {
            out.ins_col.push_back(ic[i]);
            out.ins_len.push_back(il[i]);
            out.ins_off.push_back(static_cast<int64_t>(out.ins_pool.size()) + io[i]);
        }
        out.ins_pool.append(pool.data(), static_cast<size_t>(pu));
        if (want_names) out.names.push_back(rec.name);
        ++out.nreads;
    }
}

}  // namespace mshost
