// nw.cu -- cleric's alignment step: global Needleman-Wunsch of the original and the target reference on the GPU.
//
// "The alignment step runs a Needleman-Wunsch; with NxM runtime" (/root/reference/doc/CLERIC.md:41-44); the two
// references are amplicon-sized (a few kb to ~10 kb), so N x M is up to ~1e8 cells.  SURVEY.md section 8f row 4.
// Scoring and tie-breaking are the restatement's choice U13 (oracle/ms_oracle.h): match +2, mismatch -3, linear gap -4,
// cell = max(diagonal, up, left) with ties resolved in that order -- the direction bits must be identical to the
// oracle's for the path (and with it every projected CIGAR) to be identical.
//
// Wavefront in two levels: the DP matrix is cut into 128 x 128 tiles; all tiles of one tile anti-diagonal are
// independent and run as one launch (one CTA per tile), exchanging their bottom rows / right columns through two small
// boundary arrays in global memory; inside a tile, thread t owns row t and the 255 cell anti-diagonals are stepped
// with __syncthreads, the upper neighbour's value coming through a double-buffered shared array.  Two direction bits
// per cell go to global memory (25 MB at 10 kb x 10 kb); the host reads them back and walks the path once.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "handle.h"

namespace ms {

constexpr int kNwTile = 128;
constexpr int32_t kNwMatch = 2, kNwMismatch = -3, kNwGap = -4;

// hrow[ti][j] = H[min(ti*128, la)][j], hcol[tj][i] = H[i][min(tj*128, lb)]; row strides lb+1 / la+1
__global__ void __launch_bounds__(kNwTile) nw_tile_kernel(const char* __restrict__ a, int32_t la, const char* __restrict__ b, int32_t lb,
                                                          int32_t diag, int32_t ti_first, int32_t* __restrict__ hrow, int32_t* __restrict__ hcol,
                                                          uint32_t* __restrict__ dir, int64_t dw) {
    __shared__ int32_t top[kNwTile + 1];
    __shared__ int32_t sh[2][kNwTile];
    __shared__ char bt[kNwTile];
    const int t = threadIdx.x;
    const int32_t ti = ti_first + static_cast<int32_t>(blockIdx.x), tj = diag - ti;
    const int32_t i0 = ti * kNwTile, j0 = tj * kNwTile;       // the tile covers DP rows i0+1.., columns j0+1..
    const int32_t rows = la - i0 < kNwTile ? la - i0 : kNwTile, cols = lb - j0 < kNwTile ? lb - j0 : kNwTile;
    const size_t rs = static_cast<size_t>(lb) + 1, cs = static_cast<size_t>(la) + 1;
    for (int c = t; c <= cols; c += kNwTile) top[c] = hrow[static_cast<size_t>(ti) * rs + j0 + c];
    if (t < cols) bt[t] = b[j0 + t];
    const int32_t i = i0 + t + 1;
    const bool live = t < rows;
    const char ai = live ? a[i - 1] : 'N';
    int32_t left = live ? hcol[static_cast<size_t>(tj) * cs + i] : 0;
    int32_t dg = 0;
    if (live) dg = hcol[static_cast<size_t>(tj) * cs + i - 1];   // H[i-1][j0]; for t == 0 this equals top[0]
    uint32_t bits[kNwTile / 16];
#pragma unroll
    for (int w = 0; w < kNwTile / 16; ++w) bits[w] = 0u;
    __syncthreads();
    for (int s = 0; s < rows + cols - 1; ++s) {
        const int c = s - t;                                  // column within the tile
        if (live && c >= 0 && c < cols) {
            const int32_t up = t == 0 ? top[c + 1] : sh[(s + 1) & 1][t - 1];
            const int32_t d = dg + (ai == bt[c] ? kNwMatch : kNwMismatch);
            const int32_t u = up + kNwGap, l = left + kNwGap;
            int32_t best = d;
            uint32_t k = 0;
            if (u > best) { best = u; k = 1; }
            if (l > best) { best = l; k = 2; }
            sh[s & 1][t] = best;
            bits[c >> 4] |= k << (2 * (c & 15));
            dg = up;
            left = best;
            if (t == rows - 1) hrow[static_cast<size_t>(ti + 1) * rs + j0 + c + 1] = best;
            if (c == cols - 1) hcol[static_cast<size_t>(tj + 1) * cs + i] = best;
        }
        __syncthreads();
    }
    if (live) {
        uint32_t* dst = dir + static_cast<size_t>(i - 1) * dw + (j0 >> 4);
#pragma unroll
        for (int w = 0; w < kNwTile / 16; ++w)
            if (w * 16 < cols) dst[w] = bits[w];
    }
}

__global__ void nw_init_kernel(int32_t* __restrict__ hrow, int32_t* __restrict__ hcol, int32_t la, int32_t lb, int32_t nti, int32_t ntj) {
    const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t rs = static_cast<size_t>(lb) + 1, cs = static_cast<size_t>(la) + 1;
    if (x <= lb) hrow[x] = static_cast<int32_t>(x) * kNwGap;                         // H[0][j]
    if (x <= la) hcol[x] = static_cast<int32_t>(x) * kNwGap;                         // H[i][0]
    if (x <= nti) hrow[static_cast<size_t>(x) * rs] = static_cast<int32_t>(x * kNwTile < la ? x * kNwTile : la) * kNwGap;   // column 0 of every boundary row
    if (x <= ntj) hcol[static_cast<size_t>(x) * cs] = static_cast<int32_t>(x * kNwTile < lb ? x * kNwTile : lb) * kNwGap;   // row 0 of every boundary column
}

}  // namespace ms

extern "C" int ms_align_refs(ms_handle* h, const char* a, int32_t la, const char* b, int32_t lb, char* ops, int64_t cap, int64_t* nops,
                             int64_t* score) {
    MsRange nvtx_range("K5 align references");
    if (!h || la < 0 || lb < 0 || (la > 0 && !a) || (lb > 0 && !b) || !nops || (cap > 0 && !ops)) return MS_ERR_ARG;
    *nops = static_cast<int64_t>(la) + lb;     // upper bound, refined below
    if (cap < static_cast<int64_t>(la) + lb) return MS_ERR_CAPACITY;
    if (la == 0 || lb == 0) {
        for (int32_t x = 0; x < lb; ++x) ops[x] = 'I';
        for (int32_t x = 0; x < la; ++x) ops[x] = 'D';
        *nops = la + lb;
        if (score) *score = static_cast<int64_t>(la + lb) * ms::kNwGap;
        return MS_OK;
    }
    MS_CUDA(h, cudaSetDevice(h->device));
    const int32_t nti = (la + ms::kNwTile - 1) / ms::kNwTile, ntj = (lb + ms::kNwTile - 1) / ms::kNwTile;
    const int64_t dw = (static_cast<int64_t>(lb) + 15) / 16 + 8;          // words per direction row (tile writes are 8 words wide)
    const size_t rs = static_cast<size_t>(lb) + 1, cs = static_cast<size_t>(la) + 1;
    MS_CUDA(h, h->b_nw_seq.ensure(static_cast<size_t>(la) + lb + 16));
    MS_CUDA(h, h->b_nw_hrow.ensure((static_cast<size_t>(nti) + 1) * rs * 4));
    MS_CUDA(h, h->b_nw_hcol.ensure((static_cast<size_t>(ntj) + 1) * cs * 4));
    MS_CUDA(h, h->b_nw_dir.ensure(static_cast<size_t>(la) * dw * 4));
    char* d_a = h->b_nw_seq.as<char>();
    char* d_b = d_a + la;
    MS_CUDA(h, cudaMemcpyAsync(d_a, a, la, cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(d_b, b, lb, cudaMemcpyHostToDevice, h->stream));
    const int64_t ninit = std::max<int64_t>(std::max(la, lb), std::max(nti, ntj)) + 1;
    ms::nw_init_kernel<<<static_cast<unsigned>((ninit + 255) / 256), 256, 0, h->stream>>>(h->b_nw_hrow.as<int32_t>(), h->b_nw_hcol.as<int32_t>(), la, lb,
                                                                                       nti, ntj);
    h->launches++;
    for (int32_t d = 0; d < nti + ntj - 1; ++d) {
        const int32_t first = std::max(0, d - ntj + 1), last = std::min(nti - 1, d);
        ms::nw_tile_kernel<<<last - first + 1, ms::kNwTile, 0, h->stream>>>(d_a, la, d_b, lb, d, first, h->b_nw_hrow.as<int32_t>(),
                                                                           h->b_nw_hcol.as<int32_t>(), h->b_nw_dir.as<uint32_t>(), dw);
        h->launches++;
    }
    MS_CUDA(h, cudaGetLastError());
    std::vector<uint32_t> dir(static_cast<size_t>(la) * dw);
    int32_t sc = 0;
    MS_CUDA(h, cudaMemcpyAsync(dir.data(), h->b_nw_dir.p, dir.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(&sc, h->b_nw_hrow.as<int32_t>() + static_cast<size_t>(nti) * rs + lb, 4, cudaMemcpyDeviceToHost, h->stream));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (score) *score = sc;
    // walk back from (la, lb); row 0 / column 0 are implicit
    int64_t n = 0;
    int32_t i = la, j = lb;
    while (i > 0 || j > 0) {
        uint32_t k;
        if (i == 0) k = 2;
        else if (j == 0) k = 1;
        else k = (dir[static_cast<size_t>(i - 1) * dw + ((j - 1) >> 4)] >> (2 * ((j - 1) & 15))) & 3u;
        if (k == 0) { ops[n++] = 'M'; --i; --j; }
        else if (k == 1) { ops[n++] = 'D'; --i; }
        else { ops[n++] = 'I'; --j; }
    }
    std::reverse(ops, ops + n);
    *nops = n;
    return MS_OK;
}
