// pileup.cu -- K1 kernels.  See pileup.cuh for the design and reference citations.
#include "pileup.cuh"

namespace ms {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// bulk global->shared copy completing on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// carry-save adder: (h,l) = a + b + c per bit position.  Two LOP3.
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

// ---------------------------------------------------------------- per-thread state
template <int NM>
struct Vert {
    uint32_t c[NM][kPlanes];  // vertical counters, plane k has weight 2^k
    uint32_t p3[NM], p4[NM], p5[NM];  // pending carry-save inputs of weight 8, 16, 32
};

template <int NM>
__device__ __forceinline__ void ripple(Vert<NM>& v, int i, int level, uint32_t x) {
#pragma unroll
    for (int k = 0; k < kPlanes; ++k) {
        if (k >= level) {
            uint32_t t = v.c[i][k] & x;
            v.c[i][k] ^= x;
            x = t;
        }
    }
}

// Per-thread constants for codon work
struct CodonCtx {
    uint32_t r0, r1, r0n, r1n;  // pivot planes of this block and the next
    uint32_t start;             // codon start columns in this block
    uint32_t* codon;            // global [L][64]
    int32_t colbase;            // 32*blk
};

// Build the NM one-bit masks of one read for this thread's 32 columns.
// m[0..2] raw state planes, m[3..5] pair ANDs, m[6] insertion flag,
// m[7] = "codon starting here is not the clean pivot codon".
template <bool CODON>
__device__ __forceinline__ void read_masks(uint32_t addr, const CodonCtx& cx, uint32_t (&m)[kMasks]) {
    const uint4 q = lds128(addr);
    m[0] = q.x;
    m[1] = q.y;
    m[2] = q.z;
    m[3] = q.x & q.y;
    m[4] = q.x & q.z;
    m[5] = q.y & q.z;
    m[6] = q.w;
    if (CODON) {
        const uint4 n = lds128(addr + 16);  // look-ahead block (garbage past the row end is masked by `start`)
        const uint32_t X = ((q.x ^ cx.r0) | q.z) | (q.y ^ cx.r1);     // column is not the clean pivot base
        const uint32_t Xn = ((n.x ^ cx.r0n) | n.z) | (n.y ^ cx.r1n);
        const uint32_t nm = X | __funnelshift_r(X, Xn, 1) | __funnelshift_r(X, Xn, 2);
        m[7] = nm;
        const uint32_t dirty = q.z | __funnelshift_r(q.z, n.z, 1) | __funnelshift_r(q.z, n.z, 2);
        uint32_t e = ~dirty & nm & cx.start;  // clean codon that is not the pivot codon: rare
        while (e) {
            const int j = __ffs(e) - 1;
            e &= e - 1;
            const uint32_t b0 = __funnelshift_r(q.x, n.x, j) & 7u;  // bit0 of the 3 states
            const uint32_t b1 = __funnelshift_r(q.y, n.y, j) & 7u;  // bit1 of the 3 states
            const uint32_t cod = ((b0 & 1u) << 4) | ((b1 & 1u) << 5) | ((b0 & 2u) << 1) | ((b1 & 2u) << 2) |
                                 ((b0 & 4u) >> 2) | ((b1 & 4u) >> 1);
            atomicAdd(cx.codon + (static_cast<size_t>(cx.colbase + j) * 64 + cod), 1u);
        }
    } else {
        m[7] = 0;
    }
}

template <bool CODON>
__device__ __forceinline__ void block8(uint32_t addr, uint32_t row_bytes, const CodonCtx& cx,
                                       Vert<kMasks>& v, uint32_t bi) {
    constexpr int NM = CODON ? kMasks : kMasks - 1;
    uint32_t m0[kMasks], m1[kMasks], twosA[kMasks], twosB[kMasks], foursA[kMasks], foursB[kMasks];
    // reads 0..3
    read_masks<CODON>(addr, cx, m0);
    read_masks<CODON>(addr + row_bytes, cx, m1);
#pragma unroll
    for (int i = 0; i < NM; ++i) csa(twosA[i], v.c[i][0], v.c[i][0], m0[i], m1[i]);
    read_masks<CODON>(addr + 2 * row_bytes, cx, m0);
    read_masks<CODON>(addr + 3 * row_bytes, cx, m1);
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        csa(twosB[i], v.c[i][0], v.c[i][0], m0[i], m1[i]);
        csa(foursA[i], v.c[i][1], v.c[i][1], twosA[i], twosB[i]);
    }
    // reads 4..7
    read_masks<CODON>(addr + 4 * row_bytes, cx, m0);
    read_masks<CODON>(addr + 5 * row_bytes, cx, m1);
#pragma unroll
    for (int i = 0; i < NM; ++i) csa(twosA[i], v.c[i][0], v.c[i][0], m0[i], m1[i]);
    read_masks<CODON>(addr + 6 * row_bytes, cx, m0);
    read_masks<CODON>(addr + 7 * row_bytes, cx, m1);
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        csa(twosB[i], v.c[i][0], v.c[i][0], m0[i], m1[i]);
        csa(foursB[i], v.c[i][1], v.c[i][1], twosA[i], twosB[i]);
        uint32_t e8;
        csa(e8, v.c[i][2], v.c[i][2], foursA[i], foursB[i]);
        twosA[i] = e8;  // weight-8 carry out of this block
    }
    // second level: combine the weight-8 words of successive blocks lazily (branches are CTA-uniform)
    if (bi & 1u) {
        if (bi & 2u) {
            if (bi & 4u) {
#pragma unroll
                for (int i = 0; i < NM; ++i) {
                    uint32_t x16, x32, x64;
                    csa(x16, v.c[i][3], v.c[i][3], v.p3[i], twosA[i]);
                    csa(x32, v.c[i][4], v.c[i][4], v.p4[i], x16);
                    csa(x64, v.c[i][5], v.c[i][5], v.p5[i], x32);
                    ripple(v, i, 6, x64);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NM; ++i) {
                    uint32_t x16;
                    csa(x16, v.c[i][3], v.c[i][3], v.p3[i], twosA[i]);
                    csa(v.p5[i], v.c[i][4], v.c[i][4], v.p4[i], x16);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NM; ++i) csa(v.p4[i], v.c[i][3], v.c[i][3], v.p3[i], twosA[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NM; ++i) v.p3[i] = twosA[i];
    }
}

// Fold pendings, extract per-column counts, write (or add to) this thread's private slice.
template <bool CODON>
__device__ __forceinline__ void flush(Vert<kMasks>& v, uint32_t bi, uint32_t n, uint32_t start, uint32_t* pc,
                                   uint32_t* pp, bool first) {
    constexpr int NM = CODON ? kMasks : kMasks - 1;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        if (bi & 1u) ripple(v, i, 3, v.p3[i]);
        if (bi & 2u) ripple(v, i, 4, v.p4[i]);
        if (bi & 4u) ripple(v, i, 5, v.p5[i]);
    }
    for (int j = 0; j < 32; ++j) {
        uint32_t s[kMasks];
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < kPlanes; ++k) acc |= ((v.c[i][k] >> j) & 1u) << k;
            s[i] = acc;
        }
        // states: A=000 C=001 G=010 T=011 -=100 N=101 U=111 ; s0=sum p0, s1=sum p1, s2=sum p2,
        // s3=sum p0&p1, s4=sum p0&p2, s5=sum p1&p2
        const uint32_t nU = s[5];
        const uint32_t nN = s[4] - nU;
        const uint32_t nT = s[3] - nU;
        const uint32_t nD = s[2] - nN - nU;
        const uint32_t nG = s[1] - nT - nU;
        const uint32_t nC = s[0] - nT - nN - nU;
        const uint32_t cov = n - nU;
        const uint32_t nA = cov - (nC + nG + nT + nD + nN);
        uint4 lo = make_uint4(nA, nC, nG, nT);
        uint4 hi = make_uint4(nD, nN, s[6], cov);
        uint4* dst = reinterpret_cast<uint4*>(pc + j * 8);
        if (!first) {
            const uint4 a = dst[0], b = dst[1];
            lo.x += a.x; lo.y += a.y; lo.z += a.z; lo.w += a.w;
            hi.x += b.x; hi.y += b.y; hi.z += b.z; hi.w += b.w;
        }
        dst[0] = lo;
        dst[1] = hi;
        if (CODON) {
            uint32_t piv = ((start >> j) & 1u) ? n - s[7] : 0u;
            if (!first) piv += pp[j];
            pp[j] = piv;
        }
    }
#pragma unroll
    for (int i = 0; i < kMasks; ++i) {
#pragma unroll
        for (int k = 0; k < kPlanes; ++k) v.c[i][k] = 0;
        v.p3[i] = v.p4[i] = v.p5[i] = 0;
    }
}

template <bool CODON>
__device__ __forceinline__ void pileup_body(const PileupArgs& a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = a.warps_per_group, G = a.groups, U = a.blocks8, S = a.stages;
    const int ncw = W * G;
    const uint32_t row_bytes = static_cast<uint32_t>(a.nblk) * 16u;
    const uint32_t bar0 = smem_u32(smem);          // full[s] at bar0+8s, empty[s] at bar0+8(S+s)
    const uint32_t data0 = bar0 + 128;             // stage s at data0 + s*stage_bytes
    const int64_t T = static_cast<int64_t>(G) * U * 8;
    const int64_t ntiles = (a.R + T - 1) / T;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (S + s), ncw);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == ncw) {
        // ---------------- producer: one lane streams tiles into the ring
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(bar0 + 8 * (S + stage), phase ^ 1u);
                const int64_t r0 = t * T;
                const int64_t valid = (a.R - r0 < T) ? (a.R - r0) : T;
                const uint32_t full = bar0 + 8 * stage;
                mbar_expect_tx(full, static_cast<uint32_t>(valid) * row_bytes);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(a.packed) + static_cast<size_t>(r0) * row_bytes;
                const uint32_t dst = data0 + stage * static_cast<uint32_t>(a.stage_bytes);
                for (int64_t off = 0; off < valid; off += 8) {
                    const uint32_t nr = static_cast<uint32_t>((valid - off < 8) ? (valid - off) : 8);
                    bulk_g2s(dst + static_cast<uint32_t>(off) * row_bytes, src + static_cast<size_t>(off) * row_bytes,
                             nr * row_bytes, full);
                }
                if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------- consumers: group g of W warps walks its reads of every tile
    const int group = warp / W;
    int blk = (warp - group * W) * 32 + lane;
    const bool active = blk < a.nblk;
    if (!active) blk = 0;

    CodonCtx cx;
    cx.codon = a.codon;
    cx.colbase = blk * 32;
    cx.r0 = cx.r1 = cx.r0n = cx.r1n = 0;
    cx.start = 0;
    if (CODON) {
        const uint2 p = a.pivot[blk], pn = a.pivot[blk + 1];
        cx.r0 = p.x; cx.r1 = p.y; cx.r0n = pn.x; cx.r1n = pn.y;
        cx.start = active ? a.start_mask[blk] : 0u;
    }

    Vert<kMasks> v;
#pragma unroll
    for (int i = 0; i < kMasks; ++i) {
#pragma unroll
        for (int k = 0; k < kPlanes; ++k) v.c[i][k] = 0;
        v.p3[i] = v.p4[i] = v.p5[i] = 0;
    }
    uint32_t bi = 0, n = 0;
    bool first = true;
    const size_t slice = static_cast<size_t>(blockIdx.x) * G + group;
    uint32_t* pc = a.part_col + (slice * a.nblk + blk) * 256;
    uint32_t* pp = a.part_piv + (slice * a.nblk + blk) * 32;

    uint32_t stage = 0, phase = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar0 + 8 * stage, phase);
        const int64_t r0 = t * T;
        const int valid = static_cast<int>((a.R - r0 < T) ? (a.R - r0) : T);
        for (int ub = 0; ub < U; ++ub) {
            const int first_read = (group * U + ub) * 8;
            int nv = valid - first_read;
            nv = nv < 0 ? 0 : (nv > 8 ? 8 : nv);
            const uint32_t addr = data0 + stage * static_cast<uint32_t>(a.stage_bytes) +
                                  static_cast<uint32_t>(first_read) * row_bytes + static_cast<uint32_t>(blk) * 16u;
            if (nv == 8) {
                block8<CODON>(addr, row_bytes, cx, v, bi);
                ++bi;
                n += 8;
            } else {
                for (int i = 0; i < nv; ++i) {
                    uint32_t m[kMasks];
                    read_masks<CODON>(addr + i * row_bytes, cx, m);
#pragma unroll
                    for (int q = 0; q < kMasks; ++q) ripple(v, q, 0, m[q]);
                    ++n;
                }
            }
            if (n > static_cast<uint32_t>(kMaxReadsPerFlush - 8)) {
                if (active) flush<CODON>(v, bi, n, cx.start, pc, pp, first);
                first = false;
                bi = 0;
                n = 0;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (S + stage));
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1u; }
    }
    if (active) flush<CODON>(v, bi, n, cx.start, pc, pp, first);
}

__global__ void __launch_bounds__(kPileupMaxThreads, 1) pileup_csa_kernel(PileupArgs a) {
    if (a.count_codons) pileup_body<true>(a);
    else pileup_body<false>(a);
}

// ---------------------------------------------------------------- pivot sampling
// One CTA per 32-column block; thread t looks at sample read t*R/blockDim.  Per column the
// majority of A/C/G/T among the sample becomes the pivot base.  The pivot only decides which
// codon is counted by the bit-sliced fast path; results are exact for any pivot.
__global__ void pivot_sample_kernel(const uint32_t* packed, int64_t R, int32_t nblk, int32_t L, uint2* pivot,
                                    uint8_t* pivot_state) {
    __shared__ uint32_t cnt[32][4];
    const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    if (tid < 128) cnt[tid >> 2][tid & 3] = 0;
    __syncthreads();
    const int64_t ns = R < blockDim.x ? R : blockDim.x;
    uint4 q = make_uint4(0, 0, 0xffffffffu, 0);  // state 4+: does not vote
    if (tid < ns) {
        const int64_t r = static_cast<int64_t>(tid) * R / ns;
        q = *reinterpret_cast<const uint4*>(packed + (static_cast<size_t>(r) * nblk + blk) * 4);
    }
    for (int j = 0; j < 32; ++j) {
        const uint32_t st = ((q.x >> j) & 1u) | (((q.y >> j) & 1u) << 1) | (((q.z >> j) & 1u) << 2);
#pragma unroll
        for (uint32_t s = 0; s < 4; ++s) {
            const uint32_t b = __ballot_sync(0xffffffffu, st == s);
            if (lane == 0 && b) atomicAdd(&cnt[j][s], __popc(b));
        }
    }
    __syncthreads();
    if (tid < 32) {
        uint32_t best = 0;
#pragma unroll
        for (uint32_t s = 1; s < 4; ++s)
            if (cnt[tid][s] > cnt[tid][best]) best = s;
        const uint32_t r0 = __ballot_sync(0xffffffffu, best & 1u);
        const uint32_t r1 = __ballot_sync(0xffffffffu, best & 2u);
        if (tid == 0) pivot[blk] = make_uint2(r0, r1);
        if (blk * 32 + tid < L) pivot_state[blk * 32 + tid] = static_cast<uint8_t>(best);
        if (blk == 0 && tid == 0) pivot[nblk] = make_uint2(0, 0);  // look-ahead of the last block
    }
}

// ---------------------------------------------------------------- finalize
// counts += sum over slices; pivot-codon bin of every start column += its bit-sliced count.
__global__ void pileup_finalize_kernel(const uint32_t* part_col, const uint32_t* part_piv, int32_t slices,
                                       int32_t nblk, int32_t L, const uint8_t* pivot_state,
                                       const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                       int32_t count_codons) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t ncol = static_cast<int64_t>(L) * 8;
    const size_t cstride = static_cast<size_t>(nblk) * 256, pstride = static_cast<size_t>(nblk) * 32;
    if (i < ncol) {
        uint32_t s = 0;
        for (int k = 0; k < slices; ++k) s += part_col[k * cstride + i];
        col[i] += s;
    } else if (count_codons && i < ncol + L) {
        const int64_t j = i - ncol;
        if (j + 2 < L && ((start_mask[j >> 5] >> (j & 31)) & 1u)) {
            uint32_t s = 0;
            for (int k = 0; k < slices; ++k) s += part_piv[k * pstride + j];
            const uint32_t cod = 16u * pivot_state[j] + 4u * pivot_state[j + 1] + pivot_state[j + 2];
            codon[j * 64 + cod] += s;
        }
    }
}

// ---------------------------------------------------------------- A/B baseline
// The textbook histogram north_star describes: one thread per (read, 32-column block), shared-
// memory bins flushed once per CTA, codons straight to global atomics.  Kept only so ncu can
// show why K1 is not built this way; selected with ms_set_pileup_variant(h, 1).
__global__ void pileup_atomic_kernel(const uint32_t* packed, int64_t R, int32_t L, int32_t nblk,
                                     const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                     int32_t count_codons) {
    extern __shared__ uint32_t bins[];  // [32 columns][8] per warp-block... one block column per CTA.y
    const int blk = blockIdx.y;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const uint32_t start = count_codons ? start_mask[blk] : 0u;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < R;
         r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint4 q = *reinterpret_cast<const uint4*>(packed + (static_cast<size_t>(r) * nblk + blk) * 4);
        uint4 n = make_uint4(0, 0, 0xffffffffu, 0);
        if (blk + 1 < nblk) n = *reinterpret_cast<const uint4*>(packed + (static_cast<size_t>(r) * nblk + blk + 1) * 4);
        for (int j = 0; j < 32; ++j) {
            const uint32_t st = ((q.x >> j) & 1u) | (((q.y >> j) & 1u) << 1) | (((q.z >> j) & 1u) << 2);
            if (st <= 5u) {
                atomicAdd(&bins[j * 8 + st], 1u);
                if ((q.w >> j) & 1u) atomicAdd(&bins[j * 8 + 6], 1u);
            }
            if ((start >> j) & 1u) {
                const uint32_t z = __funnelshift_r(q.z, n.z, j) & 7u;
                if (z == 0u) {
                    const uint32_t b0 = __funnelshift_r(q.x, n.x, j) & 7u, b1 = __funnelshift_r(q.y, n.y, j) & 7u;
                    const uint32_t cod = ((b0 & 1u) << 4) | ((b1 & 1u) << 5) | ((b0 & 2u) << 1) | ((b1 & 2u) << 2) |
                                         ((b0 & 4u) >> 2) | ((b1 & 4u) >> 1);
                    atomicAdd(codon + (static_cast<size_t>(blk * 32 + j) * 64 + cod), 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int c = blk * 32 + (i >> 3);
        if (c < L && (i & 7) != 7 && bins[i]) atomicAdd(col + static_cast<size_t>(c) * 8 + (i & 7), bins[i]);
    }
}

// coverage column for the atomic variant (the CSA path writes it in its flush)
__global__ void coverage_kernel(uint32_t* col, int32_t L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < L) {
        uint32_t* h = col + static_cast<size_t>(j) * 8;
        h[7] = h[0] + h[1] + h[2] + h[3] + h[4] + h[5];
    }
}

}  // namespace ms
