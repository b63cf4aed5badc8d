// bgzf_bam.hpp -- minimal BGZF + BAM reader/writer on zlib (host side of the hot path).
//
// juliet and fuse "operate on (aligned) CCS records in the BAM format" (/root/reference/doc/JULIET.md:49-58,
// /root/reference/doc/FUSE.md:13-15).  north_star names pbbam/htslib for this; neither exists in the image
// (SURVEY App. E), so this is a from-scratch reader of the public SAM/BAM specification: BGZF blocks
// (RFC 1952 members with a 'BC' extra field), the BAM header, alignment records, and typed aux tags.
// Real PacBio BAM compatibility is unverified here (no fixture is available offline); the tests round-trip
// files written by BamWriter below.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <cstdlib>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

#include "fast_crc32.hpp"
#include "fast_inflate.hpp"

namespace msbam {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct RefSeq { std::string name; int32_t length = 0; };

struct Record {
    int32_t ref_id = -1, pos = -1;
    int32_t next_ref_id = -1, next_pos = -1, tlen = 0;   // mate fields, carried through unchanged
    uint8_t mapq = 0;
    uint16_t flag = 0;
    std::string name;
    std::vector<uint32_t> cigar;   // BAM encoding: len << 4 | op  (MIDNSHP=X)
    std::string seq;               // one base per byte
    std::vector<uint8_t> qual;     // phred, 0xff when absent
    std::vector<uint8_t> aux;      // raw aux block

    // find a tag; returns pointer to the type byte or nullptr
    const uint8_t* find_tag(const char tag[2]) const;
    // Z/H string tag
    bool tag_string(const char tag[2], std::string& out) const;
    // per-base QV track: 'Z' (phred+33 text) or 'B' array of c/C/s/S; empty when absent
    std::vector<int> tag_per_base(const char tag[2]) const;
    // numeric scalar (c C s S i I f)
    bool tag_number(const char tag[2], double& out) const;
};

namespace detail {
inline uint16_t u16(const uint8_t* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
inline uint32_t u32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline int32_t i32(const uint8_t* p) { return static_cast<int32_t>(u32(p)); }
inline void put16(std::vector<uint8_t>& v, uint16_t x) { v.push_back(x & 0xff); v.push_back(x >> 8); }
inline void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((x >> (8 * i)) & 0xff); }

// size in bytes of the value of an aux field starting at its type byte; 0 on malformed
inline size_t aux_value_size(const uint8_t* t, const uint8_t* end) {
    if (t >= end) return 0;
    switch (*t) {
    case 'A': case 'c': case 'C': return 1 + 1;
    case 's': case 'S': return 1 + 2;
    case 'i': case 'I': case 'f': return 1 + 4;
    case 'Z': case 'H': {
        const uint8_t* p = t + 1;
        while (p < end && *p) ++p;
        return p < end ? static_cast<size_t>(p - t) + 1 : 0;
    }
    case 'B': {
        if (t + 6 > end) return 0;
        const uint32_t n = u32(t + 2);
        size_t w = 0;
        switch (t[1]) { case 'c': case 'C': w = 1; break; case 's': case 'S': w = 2; break; case 'i': case 'I': case 'f': w = 4; break; default: return 0; }
        return 1 + 1 + 4 + static_cast<size_t>(n) * w;
    }
    default: return 0;
    }
}
}  // namespace detail

inline const uint8_t* Record::find_tag(const char tag[2]) const {
    const uint8_t* p = aux.data();
    const uint8_t* end = p + aux.size();
    while (p + 3 <= end) {
        const size_t vs = detail::aux_value_size(p + 2, end);
        if (vs == 0 || vs > static_cast<size_t>(end - (p + 2))) return nullptr;     // malformed, or a value that runs past the record
        if (p[0] == static_cast<uint8_t>(tag[0]) && p[1] == static_cast<uint8_t>(tag[1])) return p + 2;
        p += 2 + vs;
    }
    return nullptr;
}

inline bool Record::tag_string(const char tag[2], std::string& out) const {
    const uint8_t* t = find_tag(tag);
    if (!t || (*t != 'Z' && *t != 'H')) return false;
    out.assign(reinterpret_cast<const char*>(t + 1));
    return true;
}

inline bool Record::tag_number(const char tag[2], double& out) const {
    const uint8_t* t = find_tag(tag);
    if (!t) return false;
    switch (*t) {
    case 'c': out = static_cast<int8_t>(t[1]); return true;
    case 'C': out = t[1]; return true;
    case 's': out = static_cast<int16_t>(detail::u16(t + 1)); return true;
    case 'S': out = detail::u16(t + 1); return true;
    case 'i': out = detail::i32(t + 1); return true;
    case 'I': out = detail::u32(t + 1); return true;
    case 'f': { float f; memcpy(&f, t + 1, 4); out = f; return true; }
    default: return false;
    }
}

inline std::vector<int> Record::tag_per_base(const char tag[2]) const {
    std::vector<int> v;
    const uint8_t* t = find_tag(tag);
    if (!t) return v;
    if (*t == 'Z') {
        for (const uint8_t* p = t + 1; *p; ++p) v.push_back(static_cast<int>(*p) - 33);
    } else if (*t == 'B') {
        const uint32_t n = detail::u32(t + 2);
        const uint8_t* p = t + 6;
        for (uint32_t i = 0; i < n; ++i) {
            switch (t[1]) {
            case 'c': v.push_back(static_cast<int8_t>(p[i])); break;
            case 'C': v.push_back(p[i]); break;
            case 's': v.push_back(static_cast<int16_t>(detail::u16(p + 2 * i))); break;
            case 'S': v.push_back(detail::u16(p + 2 * i)); break;
            default: return std::vector<int>();
            }
        }
    }
    return v;
}

// ------------------------------------------------------------------ BGZF reader
class BgzfReader {
public:
    explicit BgzfReader(const std::string& path) : f_(fopen(path.c_str(), "rb")) {
        if (!f_) throw Error("cannot open " + path);
    }
    ~BgzfReader() { if (f_) fclose(f_); }
    BgzfReader(const BgzfReader&) = delete;
    // read exactly n bytes of the uncompressed stream; returns false on clean EOF at a boundary
    bool read(void* dst, size_t n) {
        uint8_t* out = static_cast<uint8_t*>(dst);
        size_t got = 0;
        while (got < n) {
            if (off_ == buf_.size() && !next_block()) {
                if (got == 0) return false;
                throw Error("truncated BAM stream");
            }
            const size_t k = std::min(n - got, buf_.size() - off_);
            memcpy(out + got, buf_.data() + off_, k);
            off_ += k; got += k;
        }
        return true;
    }
private:
    bool next_block() {
        for (;;) {
            uint8_t hd[18];
            const size_t r = fread(hd, 1, 18, f_);
            if (r == 0) return false;
            if (r != 18 || hd[0] != 31 || hd[1] != 139 || hd[2] != 8 || !(hd[3] & 4)) throw Error("not a BGZF block");
            const uint16_t xlen = detail::u16(hd + 10);
            std::vector<uint8_t> extra(xlen);
            memcpy(extra.data(), hd + 12, std::min<size_t>(6, xlen));
            if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f_) != static_cast<size_t>(xlen - 6)) throw Error("truncated BGZF header");
            int bsize = -1;
            for (size_t i = 0; i + 4 <= extra.size();) {
                const uint16_t slen = detail::u16(extra.data() + i + 2);
                if (extra[i] == 66 && extra[i + 1] == 67 && slen == 2) bsize = detail::u16(extra.data() + i + 4);
                i += 4 + slen;
            }
            if (bsize < 0) throw Error("BGZF block without BC field");
            const long cdata = static_cast<long>(bsize) + 1 - 12 - xlen - 8;
            if (cdata < 0) throw Error("bad BGZF block size");
            std::vector<uint8_t> comp(static_cast<size_t>(cdata) + 8);
            if (fread(comp.data(), 1, comp.size(), f_) != comp.size()) throw Error("truncated BGZF block");
            const uint32_t crc = detail::u32(comp.data() + cdata), isize = detail::u32(comp.data() + cdata + 4);
            buf_.resize(isize);
            off_ = 0;
            if (isize == 0) continue;  // empty block (EOF marker)
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) throw Error("inflateInit2 failed");
            zs.next_in = comp.data(); zs.avail_in = static_cast<uInt>(cdata);
            zs.next_out = buf_.data(); zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END || zs.avail_out != 0) throw Error("BGZF inflate failed");
            if (crc32(crc32(0L, Z_NULL, 0), buf_.data(), isize) != crc) throw Error("BGZF CRC mismatch");
            return true;
        }
    }
    FILE* f_;
    std::vector<uint8_t> buf_;
    size_t off_ = 0;
};

// ------------------------------------------------------------------ BAM reader
class BamReader {
public:
    explicit BamReader(const std::string& path) : z_(path) {
        uint8_t magic[4];
        if (!z_.read(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) throw Error(path + ": not a BAM file");
        uint8_t b4[4];
        z_.read(b4, 4);
        text_.resize(detail::u32(b4));
        if (!text_.empty()) z_.read(&text_[0], text_.size());
        z_.read(b4, 4);
        const uint32_t nref = detail::u32(b4);
        for (uint32_t i = 0; i < nref; ++i) {
            z_.read(b4, 4);
            std::string nm(detail::u32(b4), '\0');
            z_.read(&nm[0], nm.size());
            if (!nm.empty() && nm.back() == '\0') nm.pop_back();
            z_.read(b4, 4);
            refs_.push_back({nm, detail::i32(b4)});
        }
    }
    const std::string& header_text() const { return text_; }
    const std::vector<RefSeq>& refs() const { return refs_; }
    bool next(Record& r) {
        uint8_t b4[4];
        if (!z_.read(b4, 4)) return false;
        const uint32_t bs = detail::u32(b4);
        if (bs < 32) throw Error("BAM record too small");
        raw_.resize(bs);
        z_.read(raw_.data(), bs);
        parse_record(raw_.data(), bs, r);
        return true;
    }
    // decode one alignment record from its raw bytes (after the 4-byte block_size)
    static void parse_record(const uint8_t* p, uint32_t bs, Record& r) {
        if (bs < 32) throw Error("BAM record too small");
        r.ref_id = detail::i32(p); r.pos = detail::i32(p + 4);
        const uint8_t l_name = p[8];
        r.mapq = p[9];
        const uint16_t ncig = detail::u16(p + 12);
        r.flag = detail::u16(p + 14);
        const uint32_t lseq = detail::u32(p + 16);
        r.next_ref_id = detail::i32(p + 20); r.next_pos = detail::i32(p + 24); r.tlen = detail::i32(p + 28);
        size_t o = 32;
        if (o + l_name + 4ull * ncig + (lseq + 1) / 2 + lseq > bs) throw Error("corrupt BAM record");
        r.name.assign(reinterpret_cast<const char*>(p + o), l_name ? l_name - 1 : 0);
        o += l_name;
        r.cigar.resize(ncig);
        if (ncig) memcpy(r.cigar.data(), p + o, 4ull * ncig);   // little-endian host
        o += 4ull * ncig;
        static const char dec[] = "=ACMGRSVTWYHKDBN";
        r.seq.resize(lseq);
        char* sq = lseq ? &r.seq[0] : nullptr;
        for (uint32_t i = 0; i + 1 < lseq; i += 2) { const uint8_t b = p[o + i / 2]; sq[i] = dec[b >> 4]; sq[i + 1] = dec[b & 15]; }
        if (lseq & 1) sq[lseq - 1] = dec[p[o + lseq / 2] >> 4];
        o += (lseq + 1) / 2;
        r.qual.assign(p + o, p + o + lseq);
        o += lseq;
        r.aux.assign(p + o, p + bs);
    }
private:
    BgzfReader z_;
    std::string text_;
    std::vector<RefSeq> refs_;
    std::vector<uint8_t> raw_;
};

// Byte buffer that is NOT zero-filled on allocation: the inflated stream of a large BAM is hundreds of megabytes, and a
// std::vector would first write all of it serially; here the inflating threads are the first to touch their pages.
struct Bytes {
    std::unique_ptr<uint8_t[]> p;
    size_t n = 0;
    Bytes() = default;
    explicit Bytes(size_t size) : p(new uint8_t[size ? size : 1]), n(size) {}
    uint8_t* data() { return p.get(); }
    const uint8_t* data() const { return p.get(); }
    size_t size() const { return n; }
};

// ------------------------------------------------------------------ whole-file parallel inflate
// BGZF blocks are independent gzip members: read the file, find the block boundaries from the BSIZE fields,
// inflate all blocks concurrently into one buffer.  Returns the uncompressed BAM stream.
inline Bytes inflate_file(const std::string& path, unsigned nthreads) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open " + path);
    fseek(f, 0, SEEK_END);
    const long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    Bytes file(static_cast<size_t>(fsz));
    if (fsz && fread(file.data(), 1, file.size(), f) != file.size()) { fclose(f); throw Error("short read on " + path); }
    fclose(f);
    struct Blk { size_t cpos, clen, upos; uint32_t isize, crc; };
    std::vector<Blk> blks;
    size_t o = 0, utotal = 0;
    while (o + 18 <= file.size()) {
        const uint8_t* hd = file.data() + o;
        if (hd[0] != 31 || hd[1] != 139 || hd[2] != 8 || !(hd[3] & 4)) throw Error("not a BGZF block");
        const uint16_t xlen = detail::u16(hd + 10);
        if (o + 12 + static_cast<size_t>(xlen) > file.size()) throw Error("truncated BGZF header");
        int bsize = -1;
        for (size_t i = 0; i + 4 <= xlen;) {       // the extra subfields lie inside the file (checked above)
            const uint8_t* x = hd + 12 + i;
            const uint16_t slen = detail::u16(x + 2);
            if (x[0] == 66 && x[1] == 67 && slen == 2 && i + 6 <= xlen) bsize = detail::u16(x + 4);
            i += 4 + static_cast<size_t>(slen);
        }
        // a block is header (12 + xlen) + deflate data + CRC32 + ISIZE (8): anything shorter would wrap clen below
        if (bsize < 0 || o + static_cast<size_t>(bsize) + 1 > file.size() || static_cast<size_t>(bsize) + 1 < 12 + static_cast<size_t>(xlen) + 8)
            throw Error("bad BGZF block");
        const size_t total = static_cast<size_t>(bsize) + 1, cpos = o + 12 + xlen, clen = total - 12 - xlen - 8;
        const uint32_t crc = detail::u32(file.data() + o + total - 8), isize = detail::u32(file.data() + o + total - 4);
        blks.push_back({cpos, clen, utotal, isize, crc});
        utotal += isize;
        o += total;
    }
    Bytes out(utotal);
    const bool use_fast = getenv("MS_ZLIB_INFLATE") == nullptr;
    std::vector<std::string> errs(nthreads ? nthreads : 1);
    auto work = [&](unsigned t, unsigned nt) {
        for (size_t i = t; i < blks.size(); i += nt) {
            const Blk& b = blks[i];
            if (!b.isize) continue;
            // the project's own decoder first (about twice zlib's speed on BAM blocks); the block's CRC-32 decides whether
            // its output stands, zlib decodes the block again otherwise
            if (use_fast && msinflate::fast_inflate(file.data() + b.cpos, b.clen, out.data() + b.upos, b.isize) &&
                mscrc::crc32(out.data() + b.upos, b.isize) == b.crc)
                continue;
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { errs[t] = "inflateInit2 failed"; return; }
            zs.next_in = file.data() + b.cpos; zs.avail_in = static_cast<uInt>(b.clen);
            zs.next_out = out.data() + b.upos; zs.avail_out = b.isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END || zs.avail_out != 0) { errs[t] = "BGZF inflate failed"; return; }
            if (crc32(crc32(0L, Z_NULL, 0), out.data() + b.upos, b.isize) != b.crc) { errs[t] = "BGZF CRC mismatch"; return; }
        }
    };
    if (nthreads <= 1) work(0, 1);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(work, t, nthreads);
        for (auto& x : th) x.join();
    }
    for (const std::string& e : errs) if (!e.empty()) throw Error(e);
    return out;
}

// BAM header + record offsets inside an uncompressed stream
struct BamIndexed {
    std::string text;
    std::vector<RefSeq> refs;
    std::vector<std::pair<size_t, uint32_t>> records;   // (offset of the record body, block_size)
};
inline BamIndexed index_stream(const Bytes& u) {
    BamIndexed x;
    size_t o = 0;
    auto need = [&](size_t n) { if (o + n > u.size()) throw Error("truncated BAM stream"); };
    need(12);
    if (memcmp(u.data(), "BAM\1", 4) != 0) throw Error("not a BAM file");
    const uint32_t lt = detail::u32(u.data() + 4);
    o = 8; need(lt + 4);
    x.text.assign(reinterpret_cast<const char*>(u.data() + o), lt);
    o += lt;
    const uint32_t nref = detail::u32(u.data() + o);
    o += 4;
    for (uint32_t i = 0; i < nref; ++i) {
        need(4);
        const uint32_t ln = detail::u32(u.data() + o);
        o += 4; need(ln + 4);
        std::string nm(reinterpret_cast<const char*>(u.data() + o), ln);
        if (!nm.empty() && nm.back() == '\0') nm.pop_back();
        o += ln;
        x.refs.push_back({nm, detail::i32(u.data() + o)});
        o += 4;
    }
    while (o + 4 <= u.size()) {
        const uint32_t bs = detail::u32(u.data() + o);
        o += 4; need(bs);
        x.records.emplace_back(o, bs);
        o += bs;
    }
    return x;
}

// ------------------------------------------------------------------ BAM writer (fixtures, mixdata)
// removes every aux field whose two-letter tag is in `tags` (tags of a record that describe its OLD alignment: NM, MD);
// a malformed aux block is left as it is
inline void strip_tags(std::vector<uint8_t>& aux, std::initializer_list<const char*> tags) {
    std::vector<uint8_t> out;
    out.reserve(aux.size());
    const uint8_t* p = aux.data();
    const uint8_t* end = p + aux.size();
    while (p + 3 <= end) {
        const size_t vs = detail::aux_value_size(p + 2, end);
        if (vs == 0 || p + 2 + vs > end) return;
        bool drop = false;
        for (const char* t : tags) drop = drop || (p[0] == static_cast<uint8_t>(t[0]) && p[1] == static_cast<uint8_t>(t[1]));
        if (!drop) out.insert(out.end(), p, p + 2 + vs);
        p += 2 + vs;
    }
    if (p != end) return;
    aux.swap(out);
}

class BamWriter {
public:
    BamWriter(const std::string& path, const std::string& header_text, const std::vector<RefSeq>& refs) : f_(fopen(path.c_str(), "wb")) {
        if (!f_) throw Error("cannot create " + path);
        std::vector<uint8_t> h;
        encode_header(header_text, refs, h);
        append(h.data(), h.size());
    }
    void write(const Record& r) {
        std::vector<uint8_t> b;
        encode(r, b);
        append(b.data(), b.size());
    }
    // UCSC binning scheme of the SAM specification (section 5.3): the bin of the zero-based half-open interval [beg, end)
    static uint16_t reg2bin(int64_t beg, int64_t end) {
        --end;
        if (beg >> 14 == end >> 14) return static_cast<uint16_t>(((1 << 15) - 1) / 7 + (beg >> 14));
        if (beg >> 17 == end >> 17) return static_cast<uint16_t>(((1 << 12) - 1) / 7 + (beg >> 17));
        if (beg >> 20 == end >> 20) return static_cast<uint16_t>(((1 << 9) - 1) / 7 + (beg >> 20));
        if (beg >> 23 == end >> 23) return static_cast<uint16_t>(((1 << 6) - 1) / 7 + (beg >> 23));
        if (beg >> 26 == end >> 26) return static_cast<uint16_t>(((1 << 3) - 1) / 7 + (beg >> 26));
        return 0;
    }
    // one alignment record in BAM's binary layout, appended to b
    static void encode(const Record& r, std::vector<uint8_t>& b) {
        if (r.name.size() > 254) throw Error("read name longer than 254 characters: " + r.name.substr(0, 40) + "...");
        if (r.cigar.size() > 0xffff) throw Error("more than 65535 CIGAR operations in " + r.name + " (the CG-tag form is not written)");
        const size_t at = b.size();
        const uint32_t lseq = static_cast<uint32_t>(r.seq.size());
        int64_t reflen = 0;                     // reference bases the alignment covers: M D N = X
        for (uint32_t c : r.cigar) {
            const uint32_t op = c & 15u;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += c >> 4;
        }
        detail::put32(b, 0);  // block size, patched below
        detail::put32(b, static_cast<uint32_t>(r.ref_id));
        detail::put32(b, static_cast<uint32_t>(r.pos));
        b.push_back(static_cast<uint8_t>(r.name.size() + 1));
        b.push_back(r.mapq);
        // unplaced records: reg2bin(-1, 0) = 4680 (SAM specification, section 4.2.1); no reference bases: one position
        detail::put16(b, r.pos < 0 ? 4680 : reg2bin(r.pos, static_cast<int64_t>(r.pos) + std::max<int64_t>(1, reflen)));
        detail::put16(b, static_cast<uint16_t>(r.cigar.size()));
        detail::put16(b, r.flag);
        detail::put32(b, lseq);
        detail::put32(b, static_cast<uint32_t>(r.next_ref_id));
        detail::put32(b, static_cast<uint32_t>(r.next_pos));
        detail::put32(b, static_cast<uint32_t>(r.tlen));
        b.insert(b.end(), r.name.begin(), r.name.end());
        b.push_back(0);
        for (uint32_t c : r.cigar) detail::put32(b, c);
        static const char dec[] = "=ACMGRSVTWYHKDBN";
        for (uint32_t i = 0; i < lseq; i += 2) {
            auto code = [&](char ch) { const char* q = strchr(dec, ch >= 'a' ? ch - 32 : ch); return static_cast<uint8_t>(q && *q ? q - dec : 15); };
            const uint8_t hi = code(r.seq[i]), lo = i + 1 < lseq ? code(r.seq[i + 1]) : 0;
            b.push_back(static_cast<uint8_t>((hi << 4) | lo));
        }
        if (r.qual.size() == lseq) b.insert(b.end(), r.qual.begin(), r.qual.end());
        else b.insert(b.end(), lseq, 0xff);
        b.insert(b.end(), r.aux.begin(), r.aux.end());
        const uint32_t bs = static_cast<uint32_t>(b.size() - at - 4);
        for (int i = 0; i < 4; ++i) b[at + i] = (bs >> (8 * i)) & 0xff;
    }
    // the BAM header block (magic, text, reference list), uncompressed
    static void encode_header(const std::string& header_text, const std::vector<RefSeq>& refs, std::vector<uint8_t>& h) {
        h.insert(h.end(), {'B', 'A', 'M', 1});
        detail::put32(h, static_cast<uint32_t>(header_text.size()));
        h.insert(h.end(), header_text.begin(), header_text.end());
        detail::put32(h, static_cast<uint32_t>(refs.size()));
        for (const RefSeq& r : refs) {
            detail::put32(h, static_cast<uint32_t>(r.name.size() + 1));
            h.insert(h.end(), r.name.begin(), r.name.end());
            h.push_back(0);
            detail::put32(h, static_cast<uint32_t>(r.length));
        }
    }
    // one BGZF member holding n <= 0xff00 bytes, appended to out
    static void bgzf_member(const uint8_t* p, size_t n, std::vector<uint8_t>& out, int level = 6) {
        std::vector<uint8_t> z(compressBound(static_cast<uLong>(n)) + 64);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Error("deflateInit2 failed");
        zs.next_in = const_cast<uint8_t*>(p); zs.avail_in = static_cast<uInt>(n);
        zs.next_out = z.data(); zs.avail_out = static_cast<uInt>(z.size());
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) throw Error("deflate failed");
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        const size_t total = 18 + clen + 8;
        if (total - 1 > 0xffff) throw Error("BGZF block too large");
        const uint8_t hd[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, static_cast<uint8_t>((total - 1) & 0xff),
                                static_cast<uint8_t>((total - 1) >> 8)};
        out.insert(out.end(), hd, hd + 18);
        out.insert(out.end(), z.data(), z.data() + clen);
        detail::put32(out, static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), p, static_cast<uInt>(n))));
        detail::put32(out, static_cast<uint32_t>(n));
    }
    ~BamWriter() { try { close(); } catch (...) {} }   // call close() yourself to see a write error
    void close() {
        if (!f_) return;
        FILE* f = f_;
        f_ = nullptr;
        bool ok = true;
        if (!pend_.empty()) {
            std::vector<uint8_t> out;
            bgzf_member(pend_.data(), pend_.size(), out);
            ok = fwrite(out.data(), 1, out.size(), f) == out.size();
            pend_.clear();
        }
        static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        ok = ok && fwrite(eof, 1, 28, f) == 28;
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw Error("write error on the BAM output (disk full?)");
    }
    static void aux_string(std::vector<uint8_t>& aux, const char tag[2], const std::string& s) {
        aux.push_back(tag[0]); aux.push_back(tag[1]); aux.push_back('Z');
        aux.insert(aux.end(), s.begin(), s.end());
        aux.push_back(0);
    }
    static void aux_float(std::vector<uint8_t>& aux, const char tag[2], float f) {
        aux.push_back(tag[0]); aux.push_back(tag[1]); aux.push_back('f');
        uint8_t b[4]; memcpy(b, &f, 4);
        aux.insert(aux.end(), b, b + 4);
    }
private:
    void append(const uint8_t* p, size_t n) {
        while (n) {
            const size_t k = std::min(n, static_cast<size_t>(0xff00) - pend_.size());
            pend_.insert(pend_.end(), p, p + k);
            p += k; n -= k;
            if (pend_.size() >= 0xff00) flush_block();
        }
    }
    void flush_block() {
        if (pend_.empty()) return;
        std::vector<uint8_t> out;
        bgzf_member(pend_.data(), pend_.size(), out);
        if (fwrite(out.data(), 1, out.size(), f_) != out.size()) throw Error("write error on the BAM output (disk full?)");
        pend_.clear();
    }
    FILE* f_;
    std::vector<uint8_t> pend_;
};

// Whole-file writer for tools that hold every record in memory (cleric): records are encoded and the stream is
// compressed in 0xff00-byte BGZF members on all host threads, then written in order.
inline void write_bam_parallel(const std::string& path, const std::string& header_text, const std::vector<RefSeq>& refs,
                               const std::vector<const Record*>& recs, unsigned nthreads) {
    nthreads = std::max(1u, nthreads);
    std::vector<std::vector<uint8_t>> part(nthreads);
    std::vector<std::string> errs(nthreads);        // an exception must not leave a worker thread: carried to the caller
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                try {
                    for (size_t k = recs.size() * t / nthreads; k < recs.size() * (t + 1) / nthreads; ++k) BamWriter::encode(*recs[k], part[t]);
                } catch (const std::exception& e) { errs[t] = e.what(); }
            });
        for (auto& x : th) x.join();
    }
    for (const std::string& e : errs)
        if (!e.empty()) throw Error(e);
    std::vector<uint8_t> u;
    BamWriter::encode_header(header_text, refs, u);
    size_t total = u.size();
    for (auto& p : part) total += p.size();
    u.reserve(total);
    for (auto& p : part) { u.insert(u.end(), p.begin(), p.end()); std::vector<uint8_t>().swap(p); }
    const size_t nblocks = (u.size() + 0xff00 - 1) / 0xff00;
    std::vector<std::vector<uint8_t>> z(nthreads);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                try {
                    for (size_t b = nblocks * t / nthreads; b < nblocks * (t + 1) / nthreads; ++b)
                        BamWriter::bgzf_member(u.data() + b * 0xff00, std::min<size_t>(0xff00, u.size() - b * 0xff00), z[t]);
                } catch (const std::exception& e) { errs[t] = e.what(); }
            });
        for (auto& x : th) x.join();
    }
    for (const std::string& e : errs)
        if (!e.empty()) throw Error(e);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Error("cannot create " + path);
    bool ok = true;
    for (auto& p : z) ok = ok && fwrite(p.data(), 1, p.size(), f) == p.size();
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    ok = ok && fwrite(eof, 1, 28, f) == 28;
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Error("write error on " + path + " (disk full?)");
}

}  // namespace msbam
