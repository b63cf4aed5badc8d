"""Seeded synthetic HIV-like amplicon alignments.

Strain mixing follows mixdata's rule (/root/reference/doc/MIXDATA.md:12-13): the first
strain is the major, the others are minors at given fractions.  Per-base noise is a
pure function of (seed, read, column) built on the splitmix64 finaliser, so this numpy
generator and the CUDA one (csrc/synth.cu, ms_synth_dev) emit identical reads.
"""
from dataclasses import dataclass, field

import numpy as np

K1 = np.uint64(0x9E3779B97F4A7C15)
K2 = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix64(x):
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
    return x


@dataclass
class SynthConfig:
    L: int = 3000
    seed: int = 20240001
    minor_fracs: tuple = (0.10, 0.05, 0.01)   # one entry per minor strain
    variants_per_minor: tuple = (3, 5)        # inclusive range of private codon substitutions
    sub: float = 5e-4
    dele: float = 3e-3
    homopolymer_mult: float = 10.0
    ins: float = 1e-3
    n_rate: float = 2e-2
    trunc: float = 0.02
    dense_sites: int = 0                      # >0: phasing stress layout (C5): this many shared variant sites
    dense_strains: int = 64
    frame: int = 0


@dataclass
class SynthTables:
    cfg: SynthConfig
    strain_base: np.ndarray        # [S, L] uint8
    thr_del: np.ndarray            # [L] uint32
    strain_cum: np.ndarray         # [S] uint32
    thr_N: int = 0
    thr_sub: int = 0
    thr_ins20: int = 0
    thr_trunc16: int = 0
    truth: list = field(default_factory=list)   # (strain, start col, codon) planted variants

    @property
    def nstrains(self):
        return self.strain_base.shape[0]

    @property
    def refseq(self):
        return "".join("ACGT"[b] for b in self.strain_base[0])


def make_tables(cfg: SynthConfig) -> SynthTables:
    rng = np.random.default_rng(cfg.seed)
    L = cfg.L
    ref = rng.integers(0, 4, size=L, dtype=np.uint8)
    ncodon = (L - cfg.frame) // 3
    truth = []
    if cfg.dense_sites > 0:
        S = cfg.dense_strains
        sites = np.sort(rng.choice(ncodon, size=min(cfg.dense_sites, ncodon), replace=False))
        alt = np.empty((len(sites), 3), dtype=np.uint8)
        for i, g in enumerate(sites):
            c = ref[cfg.frame + 3 * g: cfg.frame + 3 * g + 3].copy()
            k = rng.integers(0, 3)
            c[k] = (c[k] + 1 + rng.integers(0, 3)) & 3
            alt[i] = c
        strains = np.repeat(ref[None, :], S, axis=0)
        carry = rng.random((S, len(sites))) < 0.5
        carry[0, :] = False
        for s in range(S):
            for i in np.nonzero(carry[s])[0]:
                col = cfg.frame + 3 * int(sites[i])
                strains[s, col: col + 3] = alt[i]
                truth.append((s, col, int(16 * alt[i][0] + 4 * alt[i][1] + alt[i][2])))
        fr = np.full(S, 1.0 / S)
    else:
        S = 1 + len(cfg.minor_fracs)
        strains = np.repeat(ref[None, :], S, axis=0)
        used = set()
        for s in range(1, S):
            nv = int(rng.integers(cfg.variants_per_minor[0], cfg.variants_per_minor[1] + 1))
            for _ in range(nv):
                g = int(rng.integers(0, ncodon))
                while g in used:
                    g = int(rng.integers(0, ncodon))
                used.add(g)
                col = cfg.frame + 3 * g
                c = strains[s, col: col + 3].copy()
                k = rng.integers(0, 3)
                c[k] = (c[k] + 1 + rng.integers(0, 3)) & 3
                strains[s, col: col + 3] = c
                truth.append((s, col, int(16 * c[0] + 4 * c[1] + c[2])))
        fr = np.array([1.0 - sum(cfg.minor_fracs)] + list(cfg.minor_fracs))
    cum = np.minimum(np.floor(np.cumsum(fr) * 2.0 ** 32), 2.0 ** 32 - 1).astype(np.uint64)
    cum[-1] = 2 ** 32 - 1
    # homopolymer-biased deletions (screenshot juliet_hiv-context.png: 339 '-' inside AAA)
    hp = np.zeros(L, dtype=bool)
    i = 0
    while i < L:
        j = i
        while j + 1 < L and ref[j + 1] == ref[i]:
            j += 1
        if j - i + 1 >= 3:
            hp[i: j + 1] = True
        i = j + 1
    pdel = np.where(hp, cfg.dele * cfg.homopolymer_mult, cfg.dele)
    thr_del = np.floor(pdel * 2.0 ** 32).astype(np.uint32)
    return SynthTables(cfg=cfg, strain_base=np.ascontiguousarray(strains), thr_del=thr_del,
                       strain_cum=cum.astype(np.uint32),
                       thr_N=int(cfg.n_rate * 2.0 ** 32), thr_sub=int(cfg.sub * 2.0 ** 32),
                       thr_ins20=int(cfg.ins * 2.0 ** 20), thr_trunc16=int(cfg.trunc * 2.0 ** 16), truth=truth)


def read_strains(t: SynthTables, read0: int, R: int):
    """strain index, begin, end per read (the read-level half of the generator)."""
    L = t.cfg.L
    r = np.arange(read0, read0 + R, dtype=np.uint64)
    with np.errstate(over="ignore"):
        y = mix64(np.uint64(t.cfg.seed) + r * K1)
    us = (y & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    strain = np.minimum(np.searchsorted(t.strain_cum.astype(np.uint64), us, side="right"), t.nstrains - 1).astype(np.int64)
    trunc = ((y >> np.uint64(32)) & np.uint64(0xFFFF)) < np.uint64(t.thr_trunc16)
    z = (y >> np.uint64(48)).astype(np.uint64)
    amount = (((z >> np.uint64(1)) * np.uint64(L // 2)) >> np.uint64(15)).astype(np.int64)
    begin = np.where(trunc & ((z & np.uint64(1)) == 1), amount, 0)
    end = np.where(trunc & ((z & np.uint64(1)) == 0), L - amount, L)
    return strain, begin, end


def synth_states(t: SynthTables, read0: int, R: int, chunk: int = 1024) -> np.ndarray:
    """[R, L] uint8, bits0-2 state, bit3 insertion-follows (numpy twin of ms_synth_dev)."""
    L = t.cfg.L
    out = np.empty((R, L), dtype=np.uint8)
    cols = (np.arange(L, dtype=np.uint64) + np.uint64(1)) * K2
    seed = np.uint64(t.cfg.seed)
    tN = np.uint64(t.thr_N)
    tD = tN + t.thr_del.astype(np.uint64)
    tS = tD + np.uint64(t.thr_sub)
    colidx = np.arange(L)
    for c0 in range(0, R, chunk):
        n = min(chunk, R - c0)
        strain, begin, end = read_strains(t, read0 + c0, n)
        r = np.arange(read0 + c0, read0 + c0 + n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            x = mix64((seed + r * K1)[:, None] + cols[None, :])
        u = x & np.uint64(0xFFFFFFFF)
        base = t.strain_base[strain]                      # [n, L]
        subst = ((base.astype(np.uint64) + np.uint64(1) + (x >> np.uint64(32)) % np.uint64(3)) & np.uint64(3)).astype(np.uint8)
        st = np.where(u < tN, np.uint8(5), np.where(u < tD[None, :], np.uint8(4), np.where(u < tS[None, :], subst, base)))
        ins = (((x >> np.uint64(44)) & np.uint64(0xFFFFF)) < np.uint64(t.thr_ins20)).astype(np.uint8)
        covered = (colidx[None, :] >= begin[:, None]) & (colidx[None, :] < end[:, None])
        out[c0: c0 + n] = np.where(covered, st | (ins << 3), np.uint8(7))
    return out


def pack_states(states: np.ndarray) -> np.ndarray:
    """numpy packer to the planar 4-bit format: [R, L] uint8 -> [R, 4*ceil(L/32)] uint32."""
    R, L = states.shape
    nblk = (L + 31) // 32
    padded = np.full((R, nblk * 32), 7, dtype=np.uint8)
    padded[:, :L] = states
    v = padded.reshape(R, nblk, 32)
    w = (np.uint32(1) << np.arange(32, dtype=np.uint32))[None, None, :]
    out = np.empty((R, nblk, 4), dtype=np.uint32)
    for p in range(4):
        out[:, :, p] = (((v >> p) & 1).astype(np.uint32) * w).sum(axis=2, dtype=np.uint64).astype(np.uint32)
    return out.reshape(R, nblk * 4)


def unpack_states(packed: np.ndarray, L: int) -> np.ndarray:
    R = packed.shape[0]
    nblk = (L + 31) // 32
    w = packed.reshape(R, nblk, 4)
    sh = np.arange(32, dtype=np.uint32)[None, None, :]
    st = np.zeros((R, nblk, 32), dtype=np.uint8)
    for p in range(4):
        st |= (((w[:, :, p][:, :, None] >> sh) & 1).astype(np.uint8) << p)
    return st.reshape(R, nblk * 32)[:, :L].copy()


def tile_rows(packed: np.ndarray) -> np.ndarray:
    """[R, row_words] plain rows -> the device tile layout (csrc/rows.cuh, numpy twin of ms_tile_rows): a flat uint32
    array of ceil(R/8) tiles; block b of read 8t+i sits at 16-byte slot (t*nblk + b)*8 + (i ^ (b & 7)); the slots of
    reads >= R in the last tile hold "not spanned"."""
    R, rw = packed.shape
    nblk = rw // 4
    T = (R + 7) // 8
    full = np.empty((T * 8, nblk, 4), dtype=np.uint32)
    full[:R] = packed.reshape(R, nblk, 4)
    full[R:] = np.array([0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0], dtype=np.uint32)
    t = full.reshape(T, 8, nblk, 4).transpose(0, 2, 1, 3)             # [T, nblk, 8 reads, 4]
    src = np.arange(8)[None, :] ^ (np.arange(nblk)[:, None] & 7)        # position p of block b holds read p ^ (b & 7)
    out = np.take_along_axis(t, src[None, :, :, None], axis=2)
    return np.ascontiguousarray(out).reshape(-1)


def untile_rows(tiled: np.ndarray, R: int, L: int) -> np.ndarray:
    """inverse of tile_rows: flat tile array -> [R, row_words] plain rows"""
    nblk = (L + 31) // 32
    T = (R + 7) // 8
    t = np.asarray(tiled, dtype=np.uint32).reshape(-1)[: T * nblk * 32].reshape(T, nblk, 8, 4)
    src = np.arange(8)[None, :] ^ (np.arange(nblk)[:, None] & 7)        # read i of block b sits at position i ^ (b & 7)
    rows = np.take_along_axis(t, src[None, :, :, None], axis=2).transpose(0, 2, 1, 3)
    return np.ascontiguousarray(rows).reshape(T * 8, nblk * 4)[:R]


def start_mask_words(L: int, genes, region=None) -> np.ndarray:
    """bit j set where a codon of some gene (1-based [begin,end)) starts."""
    nblk = (L + 31) // 32
    m = np.zeros(nblk * 32, dtype=bool)
    lo, hi = 0, L
    if region:
        lo, hi = max(0, region[0] - 1), min(L, region[1] - 1)
    for (b, e) in genes:
        gb, ge = b - 1, min(e - 1, L)
        s = gb
        while s + 3 <= ge:
            if s >= lo and s + 3 <= hi and s >= 0:
                m[s] = True
            s += 3
    bits = m.reshape(nblk, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :]
    return bits.sum(axis=1).astype(np.uint32)
