// tiles.cu -- plain rows <-> the device tile layout (rows.cuh).
//
// Host code produces plain rows, one per read (ms_expand_cigar replaces juliet/fuse's CIGAR walk,
// /root/reference/doc/JULIET.md:49-58); the kernels read tiles of 8 reads.  The conversion is a pure permutation of
// 16-byte blocks: on the GPU behind the upload of ms_pileup_host, or on the host for callers that keep device buffers.
#include <algorithm>
#include <cstring>
#include "handle.h"
#include "rows.cuh"

namespace ms {

// one thread per 16-byte slot of the output; a warp covers 4 consecutive blocks of the 8 reads of a tile, i.e. reads
// 64 contiguous bytes from each of 8 rows and writes 512 contiguous bytes
__global__ void __launch_bounds__(256) tile_rows_kernel(const uint4* __restrict__ rows, int64_t R, int32_t nblk, uint4* __restrict__ tiled) {
    const int64_t nslots = ((R + 7) >> 3) * static_cast<int64_t>(nblk) * 8;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nslots; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t tb = i >> 3;                      // tile * nblk + block
        const int64_t t = tb / nblk;
        const int32_t b = static_cast<int32_t>(tb - t * nblk);
        const int64_t r = t * 8 + ((i & 7) ^ (b & 7));
        tiled[i] = r < R ? rows[static_cast<size_t>(r) * nblk + b] : make_uint4(~0u, ~0u, ~0u, 0u);   // not spanned
    }
}

}  // namespace ms

int ms_tile_rows_launch(ms_handle* h, const uint32_t* d_rows, int64_t R, uint32_t* d_tiled) {
    if (R <= 0) return MS_OK;
    const int64_t nslots = ms::tiles_of(R) * static_cast<int64_t>(h->nblk) * 8;
    const int grid = static_cast<int>(std::min<int64_t>((nslots + 255) / 256, static_cast<int64_t>(h->num_sms) * 16));
    ms::tile_rows_kernel<<<grid, 256, 0, h->stream>>>(reinterpret_cast<const uint4*>(d_rows), R, h->nblk, reinterpret_cast<uint4*>(d_tiled));
    h->launches++;
    MS_CUDA(h, cudaGetLastError());
    return MS_OK;
}

extern "C" {

int64_t ms_tiled_words(int32_t L, int64_t R) {
    if (L <= 0 || R < 0) return 0;
    return ms::tiles_of(R) * 8 * static_cast<int64_t>(4 * ((L + 31) / 32));
}

int ms_tile_rows(const uint32_t* rows, int64_t R, int32_t L, uint32_t* tiled) {
    if (!rows || !tiled || R < 0 || L <= 0) return MS_ERR_ARG;
    const int32_t nblk = (L + 31) / 32;
    static const uint32_t kNotSpanned[4] = {~0u, ~0u, ~0u, 0u};
    for (int64_t r = 0; r < ms::tiles_of(R) * 8; ++r)
        for (int32_t b = 0; b < nblk; ++b)
            memcpy(tiled + 4 * ms::tile_slot(r, b, nblk), r < R ? rows + (static_cast<size_t>(r) * nblk + b) * 4 : kNotSpanned, 16);
    return MS_OK;
}

int ms_untile_rows(const uint32_t* tiled, int64_t R, int32_t L, uint32_t* rows) {
    if (!rows || !tiled || R < 0 || L <= 0) return MS_ERR_ARG;
    const int32_t nblk = (L + 31) / 32;
    for (int64_t r = 0; r < R; ++r)
        for (int32_t b = 0; b < nblk; ++b) memcpy(rows + (static_cast<size_t>(r) * nblk + b) * 4, tiled + 4 * ms::tile_slot(r, b, nblk), 16);
    return MS_OK;
}

int ms_tile_rows_dev(ms_handle* h, const uint32_t* d_rows, int64_t R, uint32_t* d_tiled) {
    if (!h || h->nblk <= 0 || R < 0 || (R > 0 && (!d_rows || !d_tiled))) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    return ms_tile_rows_launch(h, d_rows, R, d_tiled);
}

}  // extern "C"
