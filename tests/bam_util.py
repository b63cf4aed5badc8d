"""Tiny BAM writer for test fixtures (python zlib): turns synthetic per-column states into aligned records
with =/X/D/I CIGARs.  An independent implementation of the format the C++ reader (host/bgzf_bam.hpp) parses."""
import struct
import zlib

import numpy as np

_OPS = {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4, "H": 5, "P": 6, "=": 7, "X": 8}
_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    total = 18 + len(comp) + 8
    return (b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def record(name, flag, pos, cigar, seq, aux=b"", ref_id=0, mapq=60):
    """cigar: list of (len, op-char)."""
    l = len(seq)
    packed = bytearray((l + 1) // 2)
    for i, ch in enumerate(seq):
        packed[i // 2] |= _CODE.get(ch.upper(), 15) << (0 if i & 1 else 4)
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name) + 1, mapq, 4680, len(cigar), flag, l, -1, -1, 0)
    body += name.encode() + b"\x00" + b"".join(struct.pack("<I", (n << 4) | _OPS[o]) for n, o in cigar) + bytes(packed) + b"\xff" * l + aux
    return struct.pack("<I", len(body)) + body


def aux_z(tag, s):
    return tag.encode() + b"Z" + s.encode() + b"\x00"


def reg2bin(beg, end):
    """SAM specification section 5.3"""
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def write_bam(path, refname, reflen, records, sort_order="unknown"):
    text = f"@HD\tVN:1.5\tSO:{sort_order}\n@SQ\tSN:{refname}\tLN:{reflen}\n".encode()
    stream = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1)
    stream += struct.pack("<I", len(refname) + 1) + refname.encode() + b"\x00" + struct.pack("<I", reflen)
    stream += b"".join(records)
    with open(path, "wb") as f:
        for i in range(0, len(stream), 0xFF00):
            f.write(_bgzf_block(stream[i: i + 0xFF00]))
        f.write(bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))


def states_to_record(name, st, ref, insertions=None, flag=0, n_via_qv=False):
    """st: uint8[L] (bits 0-2 state, bit 3 insertion flag); ref: uint8[L] reference bases 0..3.
    insertions: {col: string}.  Returns (record bytes, qv aux or b'')."""
    L = len(st)
    s = st & 7
    cov = np.nonzero(s != 7)[0]
    if len(cov) == 0:
        return None
    b, e = int(cov[0]), int(cov[-1]) + 1
    cigar, seq, qv = [], [], []

    def push(op, n=1):
        if cigar and cigar[-1][1] == op:
            cigar[-1] = (cigar[-1][0] + n, op)
        else:
            cigar.append((n, op))
    for j in range(b, e):
        x = int(s[j])
        if x == 7:
            push("N")
        elif x == 4:
            push("D")
        elif x == 5:
            if n_via_qv:                       # a real base with a failing QV: the reader must turn it into N
                push("=")
                seq.append("ACGT"[int(ref[j])])
                qv.append("!")
            else:
                push("X")
                seq.append("N")
                qv.append("I")
        else:
            push("=" if x == int(ref[j]) else "X")
            seq.append("ACGT"[x])
            qv.append("I")
        if st[j] & 8:
            ins = (insertions or {}).get(j, "A")
            push("I", len(ins))
            seq.extend(ins)
            qv.extend("I" * len(ins))
    seq = "".join(seq)
    aux = aux_z("sq", "".join(qv)) if n_via_qv else b""
    if flag & 0x10 and aux:
        aux = aux_z("sq", "".join(qv)[::-1])   # per-base tags are stored in native orientation
    return record(name, flag, b, cigar, seq, aux)


def read_bam(path):
    """Independent minimal BAM reader (BGZF = multi-member gzip): returns (header text, [(name, length)], records) with
    records as dicts(name, flag, ref_id, pos, mapq, cigar [(len, op)], seq, bin, aux bytes)."""
    import gzip
    raw = gzip.decompress(open(path, "rb").read())
    assert raw[:4] == b"BAM\x01"
    lt = struct.unpack_from("<I", raw, 4)[0]
    text = raw[8:8 + lt].decode()
    o = 8 + lt
    nref = struct.unpack_from("<I", raw, o)[0]; o += 4
    refs = []
    for _ in range(nref):
        ln = struct.unpack_from("<I", raw, o)[0]; o += 4
        nm = raw[o:o + ln - 1].decode(); o += ln
        refs.append((nm, struct.unpack_from("<I", raw, o)[0])); o += 4
    ops = "MIDNSHP=X"
    dec = "=ACMGRSVTWYHKDBN"
    recs = []
    while o < len(raw):
        bs = struct.unpack_from("<I", raw, o)[0]; o += 4
        ref_id, pos, lname, mapq, _bin, ncig, flag, lseq = struct.unpack_from("<iiBBHHHI", raw, o)
        p = o + 32
        name = raw[p:p + lname - 1].decode(); p += lname
        cig = [((v >> 4), ops[v & 15]) for v in struct.unpack_from("<%dI" % ncig, raw, p)]; p += 4 * ncig
        sq = "".join(dec[(raw[p + i // 2] >> (0 if i & 1 else 4)) & 15] for i in range(lseq))
        p += (lseq + 1) // 2 + lseq
        recs.append(dict(name=name, flag=flag, ref_id=ref_id, pos=pos, mapq=mapq, cigar=cig, seq=sq, bin=_bin, aux=bytes(raw[p:o + bs])))
        o += bs
    return text, refs, recs
