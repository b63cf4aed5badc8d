// cooc_tc.cu -- co-occurrence C = B^T B on the 5th-generation tensor cores (tcgen05, kind::i8).
//
// SURVEY.md row a13 / BASELINE.json north_star item (3): "using int8 tensor-core MMA only if the co-occurrence
// matrix is large and dense enough to be a true contraction".  The phasing stress configuration (500k reads x
// 2048 sites, ~50 % dense) is: 2048 x 2048 x 500k = 2.1e12 multiply-adds, POPC-pipe bound at ~9 ms on the
// popcount-AND kernel of phase.cu.  Here the same contraction runs as u8 x u8 -> s32 UMMA with the accumulator
// tile in tensor memory; integer arithmetic, so the result is bit-identical to the popcount kernel.
//
// Shape of one CTA: a 256 (variants v) x 256 (variants w) tile of C as two 128 x 256 accumulators that fill tensor
// memory (2 x 256 columns) and share the expanded B rows, and a contiguous range of reads (split-K over reads so that
// tiles x splits fills the 148 SMs; partial tiles are combined with red.global.add.s32).  Only tiles that touch the
// upper triangle are launched and only v <= w is written (mirrored), like the popcount kernel.
//
//   16 expander warps : one operand row each (256 rows of A, 256 of B).  Per stage a thread loads 128 reads of its
//                       variant as 16 bytes of the transposed bit matrix and expands them to 128 bytes of 0/1 in
//                       shared memory, in the K-major SWIZZLE_128B canonical layout the UMMA descriptor names
//                       (16-byte chunk c of row r lives at chunk c ^ (r & 7) of the row's 128-byte line).
//    1 MMA warp       : one elected lane issues 8 x tcgen05.mma (M128 N256 K32; 4 K-steps x 2 accumulators) per stage
//                       and commits the stage's "empty" mbarrier; after the last stage it commits the accumulator
//                       barrier.
//   epilogue          : the expander warps read the accumulators with tcgen05.ld (32 lanes x 32 columns per
//                       instruction, warp w owns TMEM lanes 32*(w%4)..) and add them into C.
//
// What bounds it is the expansion (3 integer instructions per 4 bytes: (nibble * 0x00204081) & 0x01010101) and
// shared-memory bandwidth (64 KB written, 96 KB read per stage); one accumulator per CTA ran the tensor pipe at 47 %.
#include <algorithm>
#include <vector>
#include "handle.h"

namespace ms {

constexpr int kTcAcc = 2;                                      // accumulator tiles per CTA (they share the expanded B rows)
constexpr int kTcMma = 128;                                    // M of one MMA
constexpr int kTcM = kTcAcc * kTcMma, kTcN = 256, kTcKStage = 128;   // reads per stage = one 128-byte swizzle line of u8
constexpr int kTcStages = 3;
constexpr int kTcExpanders = kTcM + kTcN;                      // 512 threads, one operand row each
constexpr int kTcThreads = kTcExpanders + 32;                  // + the MMA warp
constexpr uint32_t kTcStageBytes = kTcExpanders * 128;         // 64 KB: A rows then B rows
constexpr uint32_t kTcTmemCols = kTcAcc * kTcN;                // 512: all of tensor memory
constexpr uint32_t kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /* alignment slack */ + 256 /* barriers, tmem address */;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// K-major SWIZZLE_128B operand: rows 128 bytes apart, 8-row groups 1024 bytes apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);          // start address, 16-byte units
    d |= static_cast<uint64_t>(1) << 16;                         // leading byte offset (unused with swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
    d |= static_cast<uint64_t>(1) << 46;                         // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                         // SWIZZLE_128B
    return d;
}

// u8 x u8 -> s32, both operands K-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t tc_instr_desc() {
    return (2u << 4)              // c_format: S32
           | (0u << 7)            // a_format: unsigned 8 bit
           | (0u << 10)           // b_format: unsigned 8 bit
           | (0u << 15) | (0u << 16)   // K-major A and B
           | (static_cast<uint32_t>(kTcN >> 3) << 17) | (static_cast<uint32_t>(kTcMma >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 4 bits -> 4 bytes of 0/1 (bit i -> byte i): the partial products land on distinct bit positions, no carries
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

struct TcTile { int32_t m0, n0; };

// bt: transposed bit matrix, row v = reads as bits, `rstride` words apart (multiple of 4, zero padded).
// Tile list: tiles[blockIdx.x / splits]; this CTA's stages: split blockIdx.x % splits of `nstages` in total.
__global__ void __launch_bounds__(kTcThreads, 1) cooccurrence_tc_kernel(const uint32_t* __restrict__ bt, int32_t V, int64_t rstride,
                                                                         int64_t nstages, int32_t splits, const TcTile* __restrict__ tiles,
                                                                         int32_t* __restrict__ C) {
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc_smem_u32(tc_smem_raw);
    const uint32_t data0 = (raw + 1023u) & ~1023u;                     // SWIZZLE_128B wants 1024-byte aligned tiles
    const uint32_t bars = data0 + kTcStages * kTcStageBytes;           // full[s] at +8s, empty[s] at +64+8s, acc at +128, tmem ptr at +136
    const uint32_t full0 = bars, empty0 = bars + 64, accbar = bars + 128, tmem_slot = bars + 136;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TcTile tile = tiles[blockIdx.x / splits];
    const int32_t split = static_cast<int32_t>(blockIdx.x % splits);
    const int64_t s_begin = nstages * split / splits, s_end = nstages * (split + 1) / splits;
    const int64_t my_stages = s_end - s_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc_mbar_init(full0 + 8 * s, kTcExpanders / 32);     // one arrive per expander warp
            tc_mbar_init(empty0 + 8 * s, 1);                    // tcgen05.commit
        }
        tc_mbar_init(accbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcExpanders / 32) {                            // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < kTcExpanders / 32) {
        // ---------------- expanders: operand row `row` (A: 0..127 -> variant m0+row, B: 128..383 -> variant n0+row-128)
        const int row = threadIdx.x;
        const int32_t v = row < kTcM ? tile.m0 + row : tile.n0 + (row - kTcM);
        const bool live = v < V;
        const uint4* src = reinterpret_cast<const uint4*>(bt + static_cast<size_t>(live ? v : 0) * rstride) + s_begin;
        const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
        const uint32_t sw = static_cast<uint32_t>(row & 7);
        uint4 next = make_uint4(0, 0, 0, 0);
        if (live && my_stages > 0) next = src[0];
        for (int64_t i = 0; i < my_stages; ++i) {
            const int s = static_cast<int>(i % kTcStages);
            const uint32_t ph = static_cast<uint32_t>((i / kTcStages) & 1);
            const uint4 cur = next;
            if (live && i + 1 < my_stages) next = src[i + 1];            // next stage's bits are in flight while this one expands
            if (i >= kTcStages) tc_mbar_wait(empty0 + 8 * s, ph ^ 1u);   // the MMAs that read this slot have completed
            const uint32_t dst = data0 + static_cast<uint32_t>(s) * kTcStageBytes + row_off;
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // word q = reads 32q..32q+31 -> bytes 32q.. : chunks 2q and 2q+1
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const uint32_t x = w[q] >> (16 * hlf);
                    const uint32_t b0 = spread4(x & 15u), b1 = spread4((x >> 4) & 15u), b2 = spread4((x >> 8) & 15u), b3 = spread4((x >> 12) & 15u);
                    const uint32_t chunk = static_cast<uint32_t>(2 * q + hlf) ^ sw;
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + chunk * 16u), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(full0 + 8 * s);
        }
    } else {
        // ---------------- MMA issuer
        const uint32_t idesc = tc_instr_desc();
        for (int64_t i = 0; i < my_stages; ++i) {
            const int s = static_cast<int>(i % kTcStages);
            const uint32_t ph = static_cast<uint32_t>((i / kTcStages) & 1);
            tc_mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a0 = data0 + static_cast<uint32_t>(s) * kTcStageBytes, b0 = a0 + kTcM * 128;
#pragma unroll
                for (int k = 0; k < kTcKStage / 32; ++k)
#pragma unroll
                    for (int acc = 0; acc < kTcAcc; ++acc)   // accumulator `acc`: A rows 128*acc.., TMEM columns 256*acc..
                        tc_mma(tmem_base + static_cast<uint32_t>(acc * kTcN), tc_smem_desc(a0 + acc * kTcMma * 128 + k * 32),
                               tc_smem_desc(b0 + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);
                tc_commit(empty0 + 8 * s);                       // arrives when these MMAs have read the slot
                if (i + 1 == my_stages) tc_commit(accbar);       // ... and when the accumulator is final
            }
            __syncwarp();
        }
    }

    // ---------------- epilogue: accumulator -> C (v <= w only, mirrored), added because of the read split
    if (warp < kTcExpanders / 32 && my_stages > 0) {
        tc_mbar_wait(accbar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int part = warp >> 2;             // column range of this warp
        constexpr int kParts = kTcExpanders / 128;
        for (int it = 0; it < kTcAcc * (kTcN / 32) / kParts; ++it) {
            const int idx = it * kParts + part;                 // (accumulator, 32-column group)
            const int acc = idx / (kTcN / 32), c0 = (idx % (kTcN / 32)) * 32;
            const int32_t v = tile.m0 + acc * kTcMma + q * 32 + lane;
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kTcN + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (v < V) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int32_t w = tile.n0 + c0 + j;
                    if (w < V && v <= w && r[j] != 0u) {
                        atomicAdd(C + static_cast<size_t>(v) * V + w, static_cast<int32_t>(r[j]));
                        if (v != w) atomicAdd(C + static_cast<size_t>(w) * V + v, static_cast<int32_t>(r[j]));
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTcExpanders / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    }
}

}  // namespace ms

// opt in to the large dynamic shared memory on the current device (called once per handle by ms_create)
void ms_cooc_tc_set_smem_attr() {
    cudaFuncSetAttribute(ms::cooccurrence_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ms::kTcSmemBytes));
}

// C (V*V int32, zeroed by the caller) += B^T B over the transposed bit matrix; enqueued on the handle's stream
int ms_cooccurrence_tc_launch(ms_handle* h, const uint32_t* bt, int32_t V, int64_t rstride, int64_t R, int32_t* C) {
    const int64_t nstages = (R + ms::kTcKStage - 1) / ms::kTcKStage;
    if (V <= 0 || nstages <= 0) return MS_OK;
    // tiles that touch the upper triangle: rows [m0, m0+256) x columns [n0, n0+256) with n0 + 255 >= m0
    std::vector<ms::TcTile> tiles;
    for (int32_t n0 = 0; n0 < V; n0 += ms::kTcN)
        for (int32_t m0 = 0; m0 < V && m0 <= n0 + ms::kTcN - 1; m0 += ms::kTcM) tiles.push_back({m0, n0});
    const int64_t ntiles = static_cast<int64_t>(tiles.size());
    int64_t splits = std::max<int64_t>(1, h->num_sms / ntiles);
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, nstages / 8));      // at least 8 stages per CTA
    if (h->tc_tiles_V != V) {   // the tile list depends on V only: uploaded once per variant count, not per call
        MS_CUDA(h, h->b_tc_tiles.ensure(tiles.size() * sizeof(ms::TcTile)));
        MS_CUDA(h, cudaMemcpyAsync(h->b_tc_tiles.p, tiles.data(), tiles.size() * sizeof(ms::TcTile), cudaMemcpyHostToDevice, h->stream));
        MS_CUDA(h, cudaStreamSynchronize(h->stream));    // `tiles` is pageable host memory about to go out of scope
        h->tc_tiles_V = V;
    }
    ms::cooccurrence_tc_kernel<<<static_cast<unsigned>(ntiles * splits), ms::kTcThreads, ms::kTcSmemBytes, h->stream>>>(
        bt, V, rstride, nstages, static_cast<int32_t>(splits), h->b_tc_tiles.as<ms::TcTile>(), C);
    h->launches++;
    MS_CUDA(h, cudaGetLastError());
    return MS_OK;
}
