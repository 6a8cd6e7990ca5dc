// lz_csr.cu -- sparse operator objects: CSR (new, in the reference's container style), the
// reference's ELLPACK (objects/ell_matrix.hpp:10-21), the row-block schedule used by the SpMV /
// SpMM kernels, and the on-device generators of BASELINE.json's synthetic operators.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <limits.h>
#include <vector>
#include <algorithm>

#include <stdlib.h>

#include "lz_common.cuh"

// ---------------------------------------------------------------------------------------------
// schedule: chunk c covers rows [chunk_row[c], chunk_row[c+1]) where chunk_row[c] is the first
// row whose rowptr is >= c * LZ_SPMV_TILE.  Chunks hold < TILE + max_row_nnz non-zeros.
// ---------------------------------------------------------------------------------------------
__global__ void k_chunk_rows(int64_t n_rows, int64_t nnz, const int32_t *__restrict__ rowptr, int n_chunks, int tile,
                             int32_t *__restrict__ chunk_row, int32_t *__restrict__ chunk_ptr)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) { chunk_row[c] = (int32_t)n_rows; chunk_ptr[c] = (int32_t)nnz; return; }
    int64_t target = (int64_t)c * tile;
    int64_t lo = 0, hi = n_rows;   // first r in [0, n_rows] with rowptr[r] >= target
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    chunk_row[c] = (int32_t)lo;
    chunk_ptr[c] = rowptr[lo];
}

// chunk_ulen[c] = L when every row of chunk c has exactly L entries and L is odd (stencil interiors: 5, 7, 27 ...), else 0.
// With LZ_TRANSPOSE=1 the SpMV gather warps walk such a chunk TRANSPOSED (entry s of 32 consecutive rows per warp
// load): the gathered x entries of a warp are contiguous instead of L scattered groups (L1 wavefronts per load: ~L -> 2).
// Measured SLOWER than the storage-order walk (256^3 fused step 0.549 vs 0.438 ms, 4096^2 0.462 vs 0.338 ms: the index
// arithmetic, the strided shared-memory accesses and half the loads in flight cost more than the wavefronts saved), so
// it is an opt-in experiment, not the shipped path (profiles/r02_spmv.md).
__global__ void k_chunk_ulen(int n_chunks, const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr, int32_t *__restrict__ ulen)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int r0 = chunk_row[c], r1 = chunk_row[c + 1];
    int L = 0;
    if (r1 > r0) {
        L = rowptr[r0 + 1] - rowptr[r0];
        for (int r = r0 + 1; r < r1 && L; ++r)
            if (rowptr[r + 1] - rowptr[r] != L) L = 0;
        if (!(L & 1) || L > 63) L = 0;
    }
    ulen[c] = L;
}

__global__ void k_max_row(int64_t n_rows, const int32_t *__restrict__ rowptr, int *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (i < n_rows) len = rowptr[i + 1] - rowptr[i];
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_down_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}

// pieces per row, then virtual row pointers (row r's k-th piece starts at rowptr[r] + k*LZ_SPLIT_L)
__global__ void k_split_count(int64_t n_rows, const int32_t *__restrict__ rowptr, int32_t *__restrict__ pieces, int split_l)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { pieces[r] = 0; return; }
    const int len = rowptr[r + 1] - rowptr[r];
    pieces[r] = len <= split_l ? 1 : (len + split_l - 1) / split_l;
}
__global__ void k_split_fill(int64_t n_rows, int64_t nnz, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ vstart,
                             int32_t *__restrict__ vrowptr, int split_l)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { vrowptr[vstart[n_rows]] = (int32_t)nnz; return; }
    const int v0 = vstart[r], v1 = vstart[r + 1], s = rowptr[r];
    for (int v = v0; v < v1; ++v) vrowptr[v] = s + (v - v0) * split_l;
}

// ---------------------------------------------------------------------------------------------
// Row-length binning for power-law operators.  After the split every virtual row has <= LZ_SPLIT_L entries, but
// their lengths still range over 1..LZ_SPLIT_L and the SpMV / SpMM kernels give one thread (a 4..8-lane group) a
// row at a time: a warp is as slow as its longest row (R-MAT scale 24: mean 32, so ~35 % of the lanes work).
// The virtual rows are therefore re-ordered by DESCENDING LENGTH INSIDE WINDOWS of LZ_BIN_WINDOW rows -- a stable
// radix sort on (window, 65535 - length) -- and colidx / vals are copied into that order (owned by the library):
// rows that share a warp trip now have near-equal lengths, chunks stay contiguous slices for the bulk copies, and
// the partial sums of a window stay within a few KB of each other (locality of the combine).  x / X keep their
// order: only the ORDER OF THE ROWS changes, and the combine kernels find piece v of a row at ybar[vpos[v]].
// ---------------------------------------------------------------------------------------------
#define LZ_BIN_WINDOW 8192

__global__ void k_bin_keys(int64_t nv, const int32_t *__restrict__ vrowptr, uint32_t *__restrict__ keys, int32_t *__restrict__ ids)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int len = vrowptr[v + 1] - vrowptr[v];
    keys[v] = ((uint32_t)(v / LZ_BIN_WINDOW) << 9) | (uint32_t)(511 - min(len, 511));    // LZ_SPLIT_L <= 256 < 512
    ids[v] = (int32_t)v;
}
__global__ void k_bin_lens(int64_t nv, const int32_t *__restrict__ vrowptr, const int32_t *__restrict__ perm, int32_t *__restrict__ lens,
                           int32_t *__restrict__ vpos)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nv) return;
    if (i == nv) { lens[i] = 0; return; }
    const int v = perm[i];                    // new position i holds old virtual row v
    lens[i] = vrowptr[v + 1] - vrowptr[v];
    vpos[v] = (int32_t)i;
}
// one warp per new row: copy its entries from the old position
__global__ void __launch_bounds__(256)
k_bin_copy(int64_t nv, const int32_t *__restrict__ old_ptr, const int32_t *__restrict__ new_ptr, const int32_t *__restrict__ perm,
           const int32_t *__restrict__ colidx, const double *__restrict__ vals, int32_t *__restrict__ ncol, double *__restrict__ nval)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (i >= nv) return;
    const int src = old_ptr[perm[i]], dst = new_ptr[i], len = new_ptr[i + 1] - dst;
    for (int k = lane; k < len; k += 32) { ncol[dst + k] = colidx[src + k]; nval[dst + k] = vals[src + k]; }
}

__global__ void k_long_rows(int64_t n_rows, const int32_t *__restrict__ vstart, int32_t *__restrict__ list, int cap, int *__restrict__ counter)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    if (vstart[r + 1] - vstart[r] > LZ_LONG_PIECES) {
        const int i = atomicAdd(counter, 1);
        if (i < cap) list[i] = (int32_t)r;
    }
}

__global__ void k_split_dst(int64_t n_rows, const int32_t *__restrict__ vstart, const int32_t *__restrict__ vpos, int32_t *__restrict__ dst)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int v0 = vstart[r], v1 = vstart[r + 1];
    for (int v = v0; v < v1; ++v) {
        const int i = vpos ? vpos[v] : v;
        dst[i] = (v1 - v0 == 1) ? (int32_t)r : ~i;
    }
}
static int build_split_dst(lz_ctx *ctx, const lz_matrix *A, LzSplit *S)
{
    LZ_CUDA(cudaMalloc(&S->dst, sizeof(int32_t) * ((size_t)S->n_virtual + 8)));
    k_split_dst<<<(unsigned)((A->n_rows + 255) / 256), 256, 0, ctx->stream>>>(A->n_rows, S->vstart, S->vpos, S->dst);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// One row split of A with virtual rows of <= split_l entries into S (all arrays owned by S; bin_* stay NULL when the
// length binning is switched off)
static int build_split(lz_ctx *ctx, const lz_matrix *A, int split_l, LzSplit *S)
{
    const int64_t n = A->n_rows;
    memset(S, 0, sizeof(*S));
    int32_t *pieces;
    LZ_CUDA(cudaMalloc(&pieces, sizeof(int32_t) * (n + 1)));
    LZ_CUDA(cudaMalloc(&S->vstart, sizeof(int32_t) * (n + 1)));
    k_split_count<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, A->rowptr, pieces, split_l);
    LZ_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pieces, S->vstart, (int)(n + 1), ctx->stream);
    void *tmp;
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pieces, S->vstart, (int)(n + 1), ctx->stream);
    ctx->launches++;
    int32_t nv = 0;
    LZ_CUDA(cudaMemcpyAsync(&nv, S->vstart + n, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp));
    LZ_CUDA(cudaFree(pieces));
    S->n_virtual = nv;
    {   // the hub rows (more than LZ_LONG_PIECES pieces), in no particular order
        const int cap = nv / LZ_LONG_PIECES + 1;
        int *counter = ctx->flags + 20;
        LZ_CUDA(cudaMalloc(&S->long_rows, sizeof(int32_t) * (size_t)cap));
        LZ_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
        k_long_rows<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, S->vstart, S->long_rows, cap, counter);
        LZ_LAUNCH_CHECK(ctx);
        LZ_CUDA(cudaMemcpyAsync(&S->n_long, counter, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (S->n_long > cap) S->n_long = cap;
    }
    LZ_CUDA(cudaMalloc(&S->vrowptr, sizeof(int32_t) * ((size_t)nv + 8)));
    k_split_fill<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, A->csr_nnz, A->rowptr, S->vstart, S->vrowptr, split_l);
    LZ_LAUNCH_CHECK(ctx);
    if (!ctx->knobs.rmat_reorder || (int64_t)nv / LZ_BIN_WINDOW >= (1 << 22)) return build_split_dst(ctx, A, S);
    // ---- length binning of the virtual rows (see above) ----
    const unsigned gv = (unsigned)(((int64_t)nv + 1 + 255) / 256);
    uint32_t *keys, *keys2;
    int32_t *ids, *perm, *lens, *nptr, *ncol;
    double *nval;
    LZ_CUDA(cudaMalloc(&keys, sizeof(uint32_t) * (size_t)nv)); LZ_CUDA(cudaMalloc(&keys2, sizeof(uint32_t) * (size_t)nv));
    LZ_CUDA(cudaMalloc(&ids, sizeof(int32_t) * (size_t)nv)); LZ_CUDA(cudaMalloc(&perm, sizeof(int32_t) * (size_t)nv));
    LZ_CUDA(cudaMalloc(&lens, sizeof(int32_t) * ((size_t)nv + 1))); LZ_CUDA(cudaMalloc(&nptr, sizeof(int32_t) * ((size_t)nv + 8)));
    LZ_CUDA(cudaMalloc(&S->vpos, sizeof(int32_t) * (size_t)nv));
    k_bin_keys<<<gv, 256, 0, ctx->stream>>>(nv, S->vrowptr, keys, ids);
    LZ_LAUNCH_CHECK(ctx);
    tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, perm, nv, 0, 32, ctx->stream);
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, perm, nv, 0, 32, ctx->stream);   // stable: pieces of one row stay in order
    ctx->launches++;
    k_bin_lens<<<gv, 256, 0, ctx->stream>>>(nv, S->vrowptr, perm, lens, S->vpos);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp));
    tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens, nptr, nv + 1, ctx->stream);
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, lens, nptr, nv + 1, ctx->stream);
    ctx->launches++;
    LZ_CUDA(cudaMalloc(&ncol, sizeof(int32_t) * ((size_t)A->csr_nnz + 8)));
    LZ_CUDA(cudaMalloc(&nval, sizeof(double) * ((size_t)A->csr_nnz + 8)));
    k_bin_copy<<<(unsigned)(((int64_t)nv * 32 + 255) / 256), 256, 0, ctx->stream>>>(nv, S->vrowptr, nptr, perm, A->colidx, A->vals, ncol, nval);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp)); LZ_CUDA(cudaFree(keys)); LZ_CUDA(cudaFree(keys2)); LZ_CUDA(cudaFree(ids)); LZ_CUDA(cudaFree(perm)); LZ_CUDA(cudaFree(lens));
    LZ_CUDA(cudaFree(S->vrowptr));
    S->vrowptr = nptr;              // virtual row pointers in binned order, over the binned copies below
    S->bin_colidx = ncol; S->bin_vals = nval;
    return build_split_dst(ctx, A, S);
}

// the SpMM kernel amortises its per-chunk cost over wider rows: its own, coarser schedule, over the rows of ITS split
static int build_mm_schedule(lz_ctx *ctx, lz_matrix *A)
{
    const int64_t nnz = A->csr_nnz;
    const int32_t *rp = A->mm.vrowptr ? A->mm.vrowptr : A->rowptr;
    const int64_t rows = A->mm.vrowptr ? A->mm.n_virtual : A->n_rows;
    A->mm_k_colidx = A->mm.bin_colidx ? A->mm.bin_colidx : A->colidx;
    A->mm_k_vals = A->mm.bin_vals ? A->mm.bin_vals : A->vals;
    int64_t mch = (nnz + LZ_SPMM_TILE - 1) / LZ_SPMM_TILE;
    if (mch < 1) mch = 1;
    A->mm_n_chunks = (int)mch;
    LZ_CUDA(cudaMalloc(&A->mm_chunk_row, sizeof(int32_t) * (mch + 1)));
    LZ_CUDA(cudaMalloc(&A->mm_chunk_ptr, sizeof(int32_t) * (mch + 1)));
    LZ_CUDA(cudaMalloc(&A->mm_chunk_ulen, sizeof(int32_t) * (mch + 1)));
    k_chunk_rows<<<(unsigned)((mch + 1 + 255) / 256), 256, 0, ctx->stream>>>(rows, nnz, rp, (int)mch, LZ_SPMM_TILE, A->mm_chunk_row, A->mm_chunk_ptr);
    LZ_LAUNCH_CHECK(ctx);
    k_chunk_ulen<<<(unsigned)((mch + 255) / 256), 256, 0, ctx->stream>>>((int)mch, A->mm_chunk_row, rp, A->mm_chunk_ulen);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// X-window schedule (lz_spmm_xs.cuh).  One CTA per chunk: sort the chunk's column indices, cut the sorted list into
// segments wherever two neighbours are more than LZ_XS_MERGE apart, lay the segments out back to back (the chunk's
// window of X rows) and give every entry the window row of its column.  A stencil chunk of 73 rows of a 7-point
// operator comes out as 5 segments / ~370 rows; an operator whose chunks need more rows than shared memory holds
// (checked by the caller against xs_max_wrows) keeps the gathering kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_xs_build(const int32_t *__restrict__ chunk_ptr, const int32_t *__restrict__ colidx, int2 *__restrict__ meta, int2 *__restrict__ seg,
           uint16_t *__restrict__ lidx, int *__restrict__ stats /* [0] max window rows, [1] failures */,
           const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowmap, uint16_t *__restrict__ dli)
{
    constexpr int IPT = LZ_XS_ECAP / 256;
    typedef cub::BlockRadixSort<int, 256, IPT> Sort;
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ int sk[LZ_XS_ECAP + 1];
    __shared__ int sfirst[LZ_XS_SEGCAP], slast[LZ_XS_SEGCAP], soff[LZ_XS_SEGCAP + 1], s_nseg;
    const int c = blockIdx.x, tid = threadIdx.x;
    const int p0 = chunk_ptr[c], p1 = chunk_ptr[c + 1], cnt = p1 - p0;
    if (cnt > LZ_XS_ECAP) {                       // (uniform) too many entries for the sort: the operator does not qualify
        if (tid == 0) { atomicAdd(stats + 1, 1); meta[c] = make_int2(0, 0); }
        return;
    }
    int keys[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) { const int e = tid * IPT + i; keys[i] = e < cnt ? colidx[p0 + e] : INT_MAX; }
    Sort(tmp.sort).Sort(keys);
#pragma unroll
    for (int i = 0; i < IPT; ++i) sk[tid * IPT + i] = keys[i];
    if (tid == 0) sk[LZ_XS_ECAP] = INT_MAX;
    __syncthreads();
    // segment heads among the valid sorted keys, segment id = (inclusive count of heads) - 1
    int heads[IPT], ids[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int e = tid * IPT + i;
        heads[i] = (e < cnt && (e == 0 || sk[e] - sk[e - 1] > LZ_XS_MERGE)) ? 1 : 0;
    }
    int total_heads;
    Scan(tmp.scan).InclusiveSum(heads, ids, total_heads);
    if (tid == 0) s_nseg = total_heads;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int e = tid * IPT + i, id = ids[i] - 1;
        if (e < cnt && id < LZ_XS_SEGCAP) {
            if (heads[i]) sfirst[id] = sk[e];
            if (e == cnt - 1 || sk[e + 1] - sk[e] > LZ_XS_MERGE) slast[id] = sk[e];
        }
    }
    __syncthreads();
    const int nseg = s_nseg;
    if (nseg > LZ_XS_SEGCAP) {
        if (tid == 0) { atomicAdd(stats + 1, 1); meta[c] = make_int2(0, 0); }
        return;
    }
    if (tid == 0) {
        int off = 0;
        for (int s2 = 0; s2 < nseg; ++s2) { soff[s2] = off; off += slast[s2] - sfirst[s2] + 1; }
        soff[nseg] = off;
        meta[c] = make_int2(nseg, off);
        atomicMax(stats, off);
        if (off > 65535) atomicAdd(stats + 1, 1);
    }
    __syncthreads();
    if (tid < LZ_XS_SEGCAP) seg[(size_t)c * LZ_XS_SEGCAP + tid] = tid < nseg ? make_int2(sfirst[tid], slast[tid] - sfirst[tid] + 1) : make_int2(0, 0);
    if (soff[nseg] > 65535) return;
    for (int e = tid; e < cnt; e += 256) {
        const int col = colidx[p0 + e];
        int s2 = 0;
        while (s2 + 1 < nseg && col >= sfirst[s2 + 1]) ++s2;
        lidx[p0 + e] = (uint16_t)(soff[s2] + col - sfirst[s2]);
    }
    // window row of every walked row's OWN row of X (the Gram epilogue's A fragments), when the chunk references it
    const int r0 = chunk_row[c], r1 = chunk_row[c + 1];
    for (int i = r0 + tid; i < r1; i += 256) {
        const int col = rowmap[i];
        int s2 = 0;
        while (s2 + 1 < nseg && col >= sfirst[s2 + 1]) ++s2;
        dli[i] = (nseg > 0 && col >= sfirst[s2] && col <= slast[s2]) ? (uint16_t)(soff[s2] + col - sfirst[s2]) : (uint16_t)0xFFFF;
    }
}

__global__ void k_xs_max_rows(int n_chunks, const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ chunk_ptr, int *__restrict__ stats)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    atomicMax(stats + 2, chunk_row[c + 1] - chunk_row[c]);
    atomicMax(stats + 3, chunk_ptr[c + 1] - chunk_ptr[c]);
}

__global__ void k_xs_perm_lens(int64_t n, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ rowmap, int32_t *__restrict__ lens)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { lens[i] = 0; return; }
    const int r = rowmap[i];
    lens[i] = rowptr[r + 1] - rowptr[r];
}
// the last row of every chunk is padded with zero-valued entries so that the chunk's entry count is a multiple of 8:
// every chunk then starts on a 16-byte boundary of the value / window-index streams and the bulk copies carry it whole
__global__ void k_xs_pad_lens(int n_chunks, const int32_t *__restrict__ chunk_row, int32_t *__restrict__ lens)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int r0 = chunk_row[c], r1 = chunk_row[c + 1];
    if (r1 <= r0) return;
    int total = 0;
    for (int r = r0; r < r1; ++r) total += lens[r];
    lens[r1 - 1] += (8 - (total & 7)) & 7;
}
// one warp per row of the walked operator: its entries from the original row, then the padding (value 0, the row's own
// index as column: always a valid row of X, and 0 * x adds exactly nothing)
__global__ void __launch_bounds__(256)
k_xs_copy(int64_t n, const int32_t *__restrict__ old_ptr, const int32_t *__restrict__ new_ptr, const int32_t *__restrict__ rowmap,
          const int32_t *__restrict__ colidx, const double *__restrict__ vals, int32_t *__restrict__ ncol, double *__restrict__ nval)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (i >= n) return;
    const int r = rowmap[i];
    const int src = old_ptr[r], len = old_ptr[r + 1] - src, dst = new_ptr[i], nlen = new_ptr[i + 1] - dst;
    for (int k = lane; k < nlen; k += 32) {
        ncol[dst + k] = k < len ? colidx[src + k] : r;
        nval[dst + k] = k < len ? vals[src + k] : 0.0;
    }
}
__global__ void k_xs_desc(int n_chunks, const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ rowmap,
                          int32_t *__restrict__ chunk_ptr, int4 *__restrict__ desc, int32_t *__restrict__ oseg)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    chunk_ptr[c] = rowptr[chunk_row[c]];
    if (c == n_chunks) return;
    const int r0 = chunk_row[c], r1 = chunk_row[c + 1];
    desc[c] = make_int4(rowptr[r0], rowptr[r1], r0, r1);
    for (int g = 0; g < LZ_XS_OGROUPS; ++g) oseg[(size_t)c * LZ_XS_OGROUPS + g] = r0 + 8 * g < r1 ? rowmap[r0 + 8 * g] : -1;
}

// ---- host logic of the box schedule (pure host code behind two C-ABI entry points, so the CPU test suite covers it) ----

// Structured-grid detection on a sample of rows (HOST arrays: rowptr[0..rows] of the sample, its column indices, the
// operator row of sample row 0): the offsets (column - row) present in at least half of the sampled rows.  A 7-point
// operator on an nx x ny x nz grid gives {0, +-1, +-nx, +-nx*ny}: strides 1 | nx | nx*ny.  Returns the number of nested
// strides found (1: only the unit stride, i.e. a banded operator; 0: not even that).
static int xs_strides_from_sample(int64_t rows, int64_t row0, const int32_t *rp, const int32_t *ci, int64_t stride[3])
{
    if (rows < 64) return 0;
    const int64_t cnt = (int64_t)rp[rows] - rp[0];
    if (cnt <= 0 || cnt > rows * 64) return 0;
    std::vector<int64_t> offs;
    offs.reserve(cnt);
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t k = rp[r] - rp[0]; k < rp[r + 1] - rp[0]; ++k) offs.push_back((int64_t)ci[k] - (row0 + r));
    std::sort(offs.begin(), offs.end());
    std::vector<int64_t> common;
    for (size_t a = 0; a < offs.size();) {
        size_t b = a;
        while (b < offs.size() && offs[b] == offs[a]) ++b;
        if ((int64_t)(b - a) * 2 >= rows && offs[a] > 0) common.push_back(offs[a]);
        a = b;
    }
    if (common.empty() || common[0] != 1) return 0;
    int ns = 1;
    stride[0] = 1;
    for (size_t a = 1; a < common.size() && ns < 3; ++a)
        if (common[a] % stride[ns - 1] == 0 && common[a] / stride[ns - 1] >= 8) stride[ns++] = common[a];
    return ns;
}

// Rows of an nx x ny x nz grid (strides 1 | stride[1] | stride[2], n rows in all) in BOX order: boxes of lx x ty x tz grid
// points, x fastest, boxes clipped at the grid faces and at row n.  rowmap[i] = operator row of walked row i;
// chunk_row = first walked row of every box (+ n at the end).
static void xs_box_order(int64_t n, int ns, const int64_t stride[3], int lx, int ty, int tz, std::vector<int32_t> &rowmap, std::vector<int32_t> &chunk_row)
{
    const int64_t nx = stride[1], ny = ns == 3 ? stride[2] / stride[1] : (n + nx - 1) / nx, nz = ns == 3 ? (n + stride[2] - 1) / stride[2] : 1;
    if (ns < 3) tz = 1;
    rowmap.clear(); chunk_row.clear();
    rowmap.reserve(n);
    for (int64_t kz = 0; kz < nz; kz += tz)
        for (int64_t jy = 0; jy < ny; jy += ty)
            for (int64_t ix = 0; ix < nx; ix += lx) {
                const size_t before = rowmap.size();
                for (int64_t k = kz; k < std::min<int64_t>(kz + tz, nz); ++k)
                    for (int64_t j = jy; j < std::min<int64_t>(jy + ty, ny); ++j) {
                        const int64_t line = (ns == 3 ? k * stride[2] : 0) + j * nx;
                        const int64_t base = line + ix, end = std::min<int64_t>(std::min<int64_t>(base + lx, line + nx), n);
                        for (int64_t r = base; r < end; ++r) rowmap.push_back((int32_t)r);
                    }
                if (rowmap.size() > before) chunk_row.push_back((int32_t)before);
            }
    chunk_row.push_back((int32_t)rowmap.size());
}

extern "C" {

int lz_grid_strides_host(int64_t sample_rows, int64_t first_row, const int32_t *rowptr_host, const int32_t *colidx_host, int64_t strides[3])
{
    LZ_CHECK(rowptr_host && colidx_host && strides && sample_rows >= 0, LZ_ERR_INVALID, "lz_grid_strides_host: bad arguments");
    strides[0] = strides[1] = strides[2] = 0;
    return xs_strides_from_sample(sample_rows, first_row, rowptr_host, colidx_host, strides);
}

int lz_box_order_host(int64_t n_rows, int n_strides, const int64_t strides[3], int lx, int ty, int tz, int32_t *rowmap_host, int64_t chunk_cap,
                      int32_t *chunk_row_host, int64_t *n_chunks)
{
    LZ_CHECK(rowmap_host && chunk_row_host && n_chunks && strides && n_rows > 0 && n_rows < ((int64_t)1 << 31), LZ_ERR_INVALID, "lz_box_order_host: bad arguments");
    LZ_CHECK((n_strides == 2 || n_strides == 3) && strides[0] == 1 && strides[1] >= 2 && (n_strides == 2 || (strides[2] > 0 && strides[2] % strides[1] == 0)) &&
                 lx >= 1 && ty >= 1 && tz >= 1,
             LZ_ERR_INVALID, "lz_box_order_host: strides must be nested (1 | nx | nx*ny) and the box positive");
    std::vector<int32_t> rowmap, crow;
    xs_box_order(n_rows, n_strides, strides, lx, ty, tz, rowmap, crow);
    LZ_CHECK((int64_t)rowmap.size() == n_rows, LZ_ERR_INVALID, "lz_box_order_host: the boxes do not cover the rows");
    LZ_CHECK((int64_t)crow.size() <= chunk_cap, LZ_ERR_INVALID, "lz_box_order_host: %lld chunk boundaries do not fit %lld", (long long)crow.size(), (long long)chunk_cap);
    memcpy(rowmap_host, rowmap.data(), sizeof(int32_t) * rowmap.size());
    memcpy(chunk_row_host, crow.data(), sizeof(int32_t) * crow.size());
    *n_chunks = (int64_t)crow.size() - 1;
    return LZ_OK;
}

}  // extern "C"

// the same detection on a device operator: a sample of rows from the middle goes to the host
static int xs_detect_strides(lz_ctx *ctx, const lz_matrix *A, int64_t stride[3])
{
    (void)ctx;
    const int64_t n = A->n_rows;
    const int64_t S = std::min<int64_t>(4096, n / 2);
    if (S < 64) return 0;
    const int64_t r0 = n / 2;
    std::vector<int32_t> rp(S + 1);
    if (cudaMemcpy(rp.data(), A->rowptr + r0, sizeof(int32_t) * (S + 1), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    const int64_t cnt = (int64_t)rp[S] - rp[0];
    if (cnt <= 0 || cnt > S * 64) return 0;
    std::vector<int32_t> ci(cnt);
    if (cudaMemcpy(ci.data(), A->colidx + rp[0], sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return xs_strides_from_sample(S, r0, rp.data(), ci.data(), stride);
}

// Builds the X-window schedule of an unsplit, unsharded, square CSR operator once (first panel product).  Sets
// xs_state = 1 when every chunk got a window, -1 otherwise; the launch code checks xs_max_wrows against the shared
// memory of its panel width.
int lz_matrix_prepare_xs(lz_ctx *ctx, const lz_matrix *Ac)
{
    if (Ac->xs_state != 0) return LZ_OK;
    lz_matrix *A = const_cast<lz_matrix *>(Ac);
    A->xs_state = -1;
    if (ctx->knobs.no_xs || !A->rowptr || A->vrowptr || A->csr_nnz <= 0 || A->max_row_nnz > 64 || A->max_row_nnz < 1 || A->n_rows != A->n_cols ||
        A->halo_lo || A->halo_hi)
        return LZ_OK;
    const int64_t nnz = A->csr_nnz, n = A->n_rows;
    int *stats = ctx->flags + 16;                           // 4 ints of the context's flag bank
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(int), ctx->stream));

    // ---- chunks: boxes of grid points when the operator has nested strides, runs of rows otherwise ----
    int64_t stride[3] = {1, 0, 0};
    const int ns = ctx->knobs.xs_no_tiles ? 1 : xs_detect_strides(ctx, A, stride);
    // Measured (profiles/r02_spmm.md): with chunks that are runs of rows the staged kernel is no faster than the gathering
    // one (the same 5x window traffic, now through the copy engine) and slower on the Maxwell operator; only box-shaped
    // chunks pay.  Operators without nested strides keep the gathering kernel unless LZ_XS_FORCE is set.
    if (ns < 2 && !ctx->knobs.xs_force) return LZ_OK;
    const int avg = (int)std::max<int64_t>(1, nnz / std::max<int64_t>(n, 1));
    int rows_target = (int)std::min<int64_t>(128, std::min<int64_t>((LZ_XS_ECAP - 8) / A->max_row_nnz, ctx->knobs.xs_tile * 2 / avg));
    rows_target &= ~7;
    if (rows_target < 8) return LZ_OK;
    std::vector<int32_t> rowmap, crow;
    rowmap.reserve(n);
    if (ns >= 2) {
        // box shape: lx rows along the unit stride, ty runs along the second stride, tz along the third
        int lx = ns == 3 ? 32 : 16, ty = ns == 3 ? 2 : 8, tz = ns == 3 ? 2 : 1;      // (profiles/r02_spmm.md: box sweep)
        while (lx * ty * tz > rows_target && ty > 1) ty /= 2;
        while (lx * ty * tz > rows_target && tz > 1) tz /= 2;
        while (lx * ty * tz > rows_target && lx > 8) lx /= 2;
        if (const char *e = getenv("LZ_XS_BOX")) {
            int a, b, c2;
            if (sscanf(e, "%d,%d,%d", &a, &b, &c2) == 3 && a >= 8 && b >= 1 && c2 >= 1 && a * b * c2 <= rows_target) { lx = a; ty = b; tz = c2; }
        }
        if (ns < 3) tz = 1;
        A->xs_tile_dims[0] = lx; A->xs_tile_dims[1] = ty; A->xs_tile_dims[2] = tz;
        xs_box_order(n, ns, stride, lx, ty, tz, rowmap, crow);
        crow.pop_back();                                                 // (the end marker is appended below for both kinds of chunks)
        if ((int64_t)rowmap.size() != n) return LZ_OK;                 // (cannot happen for nested strides; keep the gathering kernel)
    } else {
        for (int64_t r = 0; r < n; ++r) { if (r % rows_target == 0) crow.push_back((int32_t)r); rowmap.push_back((int32_t)r); }
    }
    crow.push_back((int32_t)n);
    const int64_t nch = (int64_t)crow.size() - 1;
    LZ_CUDA(cudaMalloc(&A->xs_rowmap, sizeof(int32_t) * ((size_t)n + 8)));
    LZ_CUDA(cudaMemset(A->xs_rowmap, 0, sizeof(int32_t) * ((size_t)n + 8)));
    LZ_CUDA(cudaMemcpy(A->xs_rowmap, rowmap.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    LZ_CUDA(cudaMalloc(&A->xs_chunk_row, sizeof(int32_t) * (nch + 1)));
    LZ_CUDA(cudaMemcpy(A->xs_chunk_row, crow.data(), sizeof(int32_t) * (nch + 1), cudaMemcpyHostToDevice));
    // the operator's rows in chunk order, every chunk padded to a multiple of 8 entries: lengths, scan, copy
    int32_t *lens;
    LZ_CUDA(cudaMalloc(&lens, sizeof(int32_t) * (n + 1)));
    LZ_CUDA(cudaMalloc(&A->xs_rowptr, sizeof(int32_t) * ((size_t)n + 8)));
    k_xs_perm_lens<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, A->rowptr, A->xs_rowmap, lens);
    LZ_LAUNCH_CHECK(ctx);
    k_xs_pad_lens<<<(unsigned)((nch + 255) / 256), 256, 0, ctx->stream>>>((int)nch, A->xs_chunk_row, lens);
    LZ_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens, A->xs_rowptr, (int)(n + 1), ctx->stream);
    void *tmp;
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, lens, A->xs_rowptr, (int)(n + 1), ctx->stream);
    ctx->launches++;
    int32_t nnz2 = 0;
    LZ_CUDA(cudaMemcpyAsync(&nnz2, A->xs_rowptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp)); LZ_CUDA(cudaFree(lens));
    int32_t *cols2;
    LZ_CUDA(cudaMalloc(&cols2, sizeof(int32_t) * ((size_t)nnz2 + 8)));
    LZ_CUDA(cudaMalloc(&A->xs_vals, sizeof(double) * ((size_t)nnz2 + 8)));
    k_xs_copy<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(n, A->rowptr, A->xs_rowptr, A->xs_rowmap, A->colidx, A->vals, cols2, A->xs_vals);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaMalloc(&A->xs_chunk_ptr, sizeof(int32_t) * (nch + 1)));
    LZ_CUDA(cudaMalloc(&A->xs_desc, sizeof(int4) * nch));
    LZ_CUDA(cudaMalloc(&A->xs_oseg, sizeof(int32_t) * nch * LZ_XS_OGROUPS));
    k_xs_desc<<<(unsigned)((nch + 1 + 255) / 256), 256, 0, ctx->stream>>>((int)nch, A->xs_chunk_row, A->xs_rowptr, A->xs_rowmap, A->xs_chunk_ptr, A->xs_desc,
                                                                          A->xs_oseg);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaMalloc(&A->xs_meta, sizeof(int2) * nch));
    LZ_CUDA(cudaMalloc(&A->xs_seg, sizeof(int2) * nch * LZ_XS_SEGCAP));
    LZ_CUDA(cudaMalloc(&A->xs_lidx, sizeof(uint16_t) * ((size_t)nnz2 + 16)));
    LZ_CUDA(cudaMemsetAsync(A->xs_lidx, 0, sizeof(uint16_t) * ((size_t)nnz2 + 16), ctx->stream));
    k_xs_max_rows<<<(unsigned)((nch + 255) / 256), 256, 0, ctx->stream>>>((int)nch, A->xs_chunk_row, A->xs_chunk_ptr, stats);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaMalloc(&A->xs_dli, sizeof(uint16_t) * ((size_t)n + 16)));
    LZ_CUDA(cudaMemsetAsync(A->xs_dli, 0xFF, sizeof(uint16_t) * ((size_t)n + 16), ctx->stream));
    k_xs_build<<<(unsigned)nch, 256, 0, ctx->stream>>>(A->xs_chunk_ptr, cols2, A->xs_meta, A->xs_seg, A->xs_lidx, stats, A->xs_chunk_row, A->xs_rowmap,
                                                       A->xs_dli);
    LZ_LAUNCH_CHECK(ctx);
    int h[4];
    LZ_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(cols2);                                        // the product streams window indices, not columns
    A->xs_n_chunks = (int)nch; A->xs_max_wrows = h[0]; A->xs_max_rows = h[2]; A->xs_max_entries = h[3];
    if (h[1] == 0 && h[0] > 0) { A->xs_state = 1; return LZ_OK; }
    cudaFree(A->xs_chunk_row); cudaFree(A->xs_chunk_ptr); cudaFree(A->xs_meta); cudaFree(A->xs_seg); cudaFree(A->xs_lidx);
    cudaFree(A->xs_rowptr); cudaFree(A->xs_rowmap); cudaFree(A->xs_vals); cudaFree(A->xs_desc); cudaFree(A->xs_oseg); cudaFree(A->xs_dli);
    A->xs_chunk_row = A->xs_chunk_ptr = nullptr; A->xs_meta = A->xs_seg = nullptr; A->xs_lidx = nullptr;
    A->xs_rowptr = A->xs_rowmap = nullptr; A->xs_vals = nullptr; A->xs_desc = nullptr; A->xs_oseg = nullptr; A->xs_dli = nullptr;
    return LZ_OK;
}

// First panel product on a row-split operator whose SpMM split differs from the SpMV's: build it now (one radix sort and
// one copy of the entries; the operator handle is logically const, this is its lazily built cache)
int lz_matrix_prepare_mm(lz_ctx *ctx, const lz_matrix *Ac)
{
    if (!Ac->mm_pending) return LZ_OK;
    lz_matrix *A = const_cast<lz_matrix *>(Ac);
    LZ_TRY(build_split(ctx, A, ctx->knobs.split_l_mm, &A->mm));
    LZ_TRY(build_mm_schedule(ctx, A));
    A->mm_pending = 0;
    return LZ_OK;
}

static int build_schedule(lz_ctx *ctx, lz_matrix *A)
{
    if (A->format == LZ_FMT_CSR) A->csr_nnz = A->nnz;
    const int64_t nnz = A->csr_nnz;
    A->tile = LZ_SPMV_TILE;
    A->cap = 1024;                       // shared-memory slots per ring stage of the fine-schedule kernel
    int *d_max = ctx->flags + 8;
    LZ_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    k_max_row<<<(unsigned)((A->n_rows + 255) / 256), 256, 0, ctx->stream>>>(A->n_rows, A->rowptr, d_max);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaMemcpyAsync(&A->max_row_nnz, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (A->max_row_nnz > A->cap - A->tile && !ctx->knobs.no_split) {                   // a long row would push chunks off the streaming path
        LzSplit S;
        LZ_TRY(build_split(ctx, A, ctx->knobs.split_l, &S));
        A->n_virtual = S.n_virtual; A->vstart = S.vstart; A->vrowptr = S.vrowptr; A->vpos = S.vpos;
        A->bin_colidx = S.bin_colidx; A->bin_vals = S.bin_vals;
        A->sv_long_rows = S.long_rows; A->sv_dst = S.dst;                                                  // (kept only to be freed: the SpMV combine does not use it)
        LZ_CUDA(cudaMalloc(&A->ybar, sizeof(double) * ((size_t)S.n_virtual + 8)));
        if (ctx->knobs.split_l_mm == ctx->knobs.split_l) { A->mm = S; A->mm_shared = 1; }
        else A->mm_pending = 1;                                                         // built by the first panel product
    }
    const int32_t *rp = A->vrowptr ? A->vrowptr : A->rowptr;
    const int64_t rows = A->vrowptr ? A->n_virtual : A->n_rows;
    int64_t nch = (nnz + A->tile - 1) / A->tile;
    if (nch < 1) nch = 1;
    LZ_CHECK(nch * 2 <= LZ_PARTIALS_CAP, LZ_ERR_UNSUPPORTED, "matrix too large for the reduction scratch (%lld chunks)", (long long)nch);
    A->n_chunks = (int)nch;
    LZ_CUDA(cudaMalloc(&A->chunk_row, sizeof(int32_t) * (nch + 1)));
    LZ_CUDA(cudaMalloc(&A->chunk_ptr, sizeof(int32_t) * (nch + 1)));
    LZ_CUDA(cudaMalloc(&A->chunk_ulen, sizeof(int32_t) * (nch + 1)));
    A->k_colidx = A->bin_colidx ? A->bin_colidx : A->colidx;       // what the SpMV kernels stream
    A->k_vals = A->bin_vals ? A->bin_vals : A->vals;
    A->tma_ok = ((uintptr_t)A->k_vals % 16 == 0) && ((uintptr_t)A->k_colidx % 16 == 0) && ((uintptr_t)rp % 16 == 0);   // bulk copies: 16-byte sources
    k_chunk_rows<<<(unsigned)((nch + 1 + 255) / 256), 256, 0, ctx->stream>>>(rows, nnz, rp, (int)nch, A->tile, A->chunk_row, A->chunk_ptr);
    LZ_LAUNCH_CHECK(ctx);
    k_chunk_ulen<<<(unsigned)((nch + 255) / 256), 256, 0, ctx->stream>>>((int)nch, A->chunk_row, rp, A->chunk_ulen);
    LZ_LAUNCH_CHECK(ctx);
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (A->mm_pending) return LZ_OK;
    return build_mm_schedule(ctx, A);
}

static lz_matrix *new_matrix(lz_ctx *ctx, int fmt, int64_t n_rows, int64_t n_cols, int64_t nnz)
{
    lz_matrix *A = new lz_matrix();
    memset(A, 0, sizeof(*A));
    A->ctx = ctx;
    A->device = ctx->device;
    A->next = ctx->matrices;                 // register with the context (lz_ctx_destroy orphans what is still alive)
    if (ctx->matrices) ctx->matrices->prev = A;
    ctx->matrices = A;
    A->format = fmt;
    A->n_rows = n_rows;
    A->n_cols = n_cols;
    A->nnz = nnz;
    A->global_rows = n_rows;
    return A;
}

static int check_dims(int64_t n_rows, int64_t n_cols, int64_t nnz)
{
    LZ_CHECK(n_rows > 0 && n_cols > 0 && nnz >= 0, LZ_ERR_INVALID, "bad matrix dimensions %lld x %lld, nnz %lld",
             (long long)n_rows, (long long)n_cols, (long long)nnz);
    LZ_CHECK(n_rows < 2147483647LL && n_cols < 2147483647LL && nnz < 2147483647LL, LZ_ERR_UNSUPPORTED,
             "int32 index space exceeded (rows %lld, nnz %lld)", (long long)n_rows, (long long)nnz);
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// ELL helpers
// ---------------------------------------------------------------------------------------------
// count the structurally non-zero entries of every ELL row (layout 0: data[r + k*n], 1: data[w*r + k])
__global__ void k_ell_count(int64_t n, int width, int layout, const double *__restrict__ data, int32_t *__restrict__ cnt)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int c = 0;
    for (int k = 0; k < width; ++k) {
        double v = layout == 0 ? data[r + (int64_t)k * n] : data[(int64_t)width * r + k];
        c += (v != 0.0);
    }
    cnt[r] = c;
}

__global__ void k_ell_fill(int64_t n, int width, int layout, const double *__restrict__ data,
                           const uint32_t *__restrict__ idx, const int32_t *__restrict__ rowptr,
                           int32_t *__restrict__ colidx, double *__restrict__ vals)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int p = rowptr[r];
    for (int k = 0; k < width; ++k) {
        int64_t src = layout == 0 ? r + (int64_t)k * n : (int64_t)width * r + k;
        double v = data[src];
        if (v != 0.0) { colidx[p] = (int32_t)idx[src]; vals[p] = v; ++p; }
    }
}

// column-major width-4 ELL -> row-interleaved (the job of lm::change_major, ell_kernels.hpp:99-121)
__global__ void k_ell4_interleave(int64_t n, const double *__restrict__ data, const uint32_t *__restrict__ idx,
                                  double *__restrict__ odata, uint32_t *__restrict__ oidx)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        odata[4 * r + k] = data[r + (int64_t)k * n];
        oidx[4 * r + k] = idx[r + (int64_t)k * n];
    }
}

// ---------------------------------------------------------------------------------------------
// generators
// ---------------------------------------------------------------------------------------------
// rows [row0, row0 + n_local) of the nx*ny 5-point Laplacian, column ids shifted by col_shift
__device__ __host__ inline int64_t lz_lap2d_before(int64_t i, int64_t nx, int64_t ny)
{
    // entries before row i = 5 i - (#rows<i on each of the four boundaries)
    int64_t y = i / nx, x = i - y * nx;
    if (i >= nx * ny) { y = ny; x = 0; }
    const int64_t b0 = i < nx ? i : nx, b1 = i - (ny - 1) * nx > 0 ? i - (ny - 1) * nx : 0;
    return 5 * i - b0 - b1 - (y + (x > 0)) - y;
}

__global__ void k_lap2d(int64_t nx, int64_t ny, int64_t row0, int64_t n_local, int64_t col_shift,
                        int32_t *__restrict__ rowptr, int32_t *__restrict__ colidx, double *__restrict__ vals)
{
    int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li > n_local) return;
    const int64_t i = row0 + li;
    int64_t p = lz_lap2d_before(i, nx, ny) - lz_lap2d_before(row0, nx, ny);
    rowptr[li] = (int32_t)p;
    if (li == n_local) return;
    const int64_t y = i / nx, x = i - y * nx, c = i - col_shift;
    if (y > 0)      { colidx[p] = (int32_t)(c - nx); vals[p++] = -1.0; }
    if (x > 0)      { colidx[p] = (int32_t)(c - 1);  vals[p++] = -1.0; }
    colidx[p] = (int32_t)c; vals[p++] = 4.0;
    if (x < nx - 1) { colidx[p] = (int32_t)(c + 1);  vals[p++] = -1.0; }
    if (y < ny - 1) { colidx[p] = (int32_t)(c + nx); vals[p++] = -1.0; }
}

// rows [row0, row0 + n_local) of the nx*ny*nz 7-point Laplacian.  Column ids are written as
// (global - col_shift) so a shard sees [lower halo | local | upper halo] as one index space.
__global__ void k_lap3d(int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_local, int64_t col_shift,
                        int32_t *__restrict__ rowptr, int32_t *__restrict__ colidx, double *__restrict__ vals)
{
    int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li > n_local) return;
    const int64_t sxy = nx * ny;
    auto before = [&](int64_t i) -> int64_t {
        int64_t z = i / sxy, rem = i - z * sxy, y = rem / nx, x = rem - y * nx;
        if (i >= sxy * nz) { z = nz; y = 0; x = 0; }
        int64_t bz0 = min(i, sxy), bz1 = max((int64_t)0, i - (nz - 1) * sxy);
        int64_t by0 = z * nx + (y > 0 ? nx : x);
        int64_t by1 = z * nx + (y == ny - 1 ? x : 0);
        int64_t bx0 = z * ny + y + (x > 0), bx1 = z * ny + y;
        return 7 * i - bz0 - bz1 - by0 - by1 - bx0 - bx1;
    };
    const int64_t i = row0 + li;
    const int64_t base = before(row0);
    int64_t p = before(i) - base;
    rowptr[li] = (int32_t)p;
    if (li == n_local) return;
    int64_t z = i / sxy, rem = i - z * sxy, y = rem / nx, x = rem - y * nx;
    const int64_t c = i - col_shift;
    if (z > 0)      { colidx[p] = (int32_t)(c - sxy); vals[p++] = -1.0; }
    if (y > 0)      { colidx[p] = (int32_t)(c - nx);  vals[p++] = -1.0; }
    if (x > 0)      { colidx[p] = (int32_t)(c - 1);   vals[p++] = -1.0; }
    colidx[p] = (int32_t)c; vals[p++] = 6.0;
    if (x < nx - 1) { colidx[p] = (int32_t)(c + 1);   vals[p++] = -1.0; }
    if (y < ny - 1) { colidx[p] = (int32_t)(c + nx);  vals[p++] = -1.0; }
    if (z < nz - 1) { colidx[p] = (int32_t)(c + sxy); vals[p++] = -1.0; }
}

__global__ void k_start_vector(int64_t n, uint64_t seed, int64_t offset, double *__restrict__ v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = 2.0 * lz_u01(lz_splitmix64(seed ^ (uint64_t)(i + offset))) - 1.0;
}

__global__ void k_start_block(int64_t n, int b, int64_t ld, uint64_t seed, double *__restrict__ V)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < b; ++c) V[i + (int64_t)c * ld] = 2.0 * lz_u01(lz_splitmix64(seed ^ (uint64_t)(i * b + c))) - 1.0;
}

// R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05): one counter-based draw per (edge, level)   [SURVEY 8d, config 4]
__global__ void k_rmat_edges(int scale, int64_t n_edges, uint64_t seed, int32_t *__restrict__ src, int32_t *__restrict__ dst)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    uint32_t r = 0, c = 0;
    for (int l = 0; l < scale; ++l) {
        const double u = lz_u01(lz_splitmix64(seed ^ ((uint64_t)e * 64ULL + (uint64_t)l)));
        const int q = (u < 0.57) ? 0 : (u < 0.76) ? 1 : (u < 0.95) ? 2 : 3;
        r = (r << 1) | (uint32_t)(q >> 1);
        c = (c << 1) | (uint32_t)(q & 1);
    }
    src[e] = (int32_t)r;
    dst[e] = (int32_t)c;
}

__global__ void k_fill(int64_t n, double value, double *__restrict__ x)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = value;
}

// ---------------------------------------------------------------------------------------------
static int finish_csr(lz_ctx *ctx, lz_matrix *A, lz_matrix **out)
{
    int s = build_schedule(ctx, A);
    if (s != LZ_OK) { lz_matrix_destroy(A); return s; }
    *out = A;
    return LZ_OK;
}

int lz_gen_lap3d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_local,
                      int64_t col_shift, int64_t n_cols, lz_matrix **out)
{
    const int64_t sxy = nx * ny;
    // nnz of the slab, from the same closed form the kernel uses (host mirror)
    auto before = [&](int64_t i) -> int64_t {
        int64_t z = i / sxy, rem = i - z * sxy, y = rem / nx, x = rem - y * nx;
        if (i >= sxy * nz) { z = nz; y = 0; x = 0; }
        int64_t bz0 = i < sxy ? i : sxy, bz1 = i - (nz - 1) * sxy > 0 ? i - (nz - 1) * sxy : 0;
        int64_t by0 = z * nx + (y > 0 ? nx : x);
        int64_t by1 = z * nx + (y == ny - 1 ? x : 0);
        int64_t bx0 = z * ny + y + (x > 0), bx1 = z * ny + y;
        return 7 * i - bz0 - bz1 - by0 - by1 - bx0 - bx1;
    };
    const int64_t nnz = before(row0 + n_local) - before(row0);
    LZ_TRY(check_dims(n_local, n_cols, nnz));
    lz_matrix *A = new_matrix(ctx, LZ_FMT_CSR, n_local, n_cols, nnz);
    A->owns = 1;
    int32_t *rp, *ci;
    double *va;
    LZ_CUDA(cudaMalloc(&rp, sizeof(int32_t) * (n_local + 1)));
    LZ_CUDA(cudaMalloc(&ci, sizeof(int32_t) * (nnz + 8)));
    LZ_CUDA(cudaMalloc(&va, sizeof(double) * (nnz + 8)));
    A->rowptr = rp; A->colidx = ci; A->vals = va;
    k_lap3d<<<(unsigned)((n_local + 1 + 255) / 256), 256, 0, ctx->stream>>>(nx, ny, nz, row0, n_local, col_shift, rp, ci, va);
    LZ_LAUNCH_CHECK(ctx);
    return finish_csr(ctx, A, out);
}

int lz_gen_lap2d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t row0, int64_t n_local, int64_t col_shift,
                      int64_t n_cols, lz_matrix **out)
{
    const int64_t nnz = lz_lap2d_before(row0 + n_local, nx, ny) - lz_lap2d_before(row0, nx, ny);
    LZ_TRY(check_dims(n_local, n_cols, nnz));
    lz_matrix *A = new_matrix(ctx, LZ_FMT_CSR, n_local, n_cols, nnz);
    A->owns = 1;
    int32_t *rp, *ci;
    double *va;
    LZ_CUDA(cudaMalloc(&rp, sizeof(int32_t) * (n_local + 1)));
    LZ_CUDA(cudaMalloc(&ci, sizeof(int32_t) * (nnz + 8)));
    LZ_CUDA(cudaMalloc(&va, sizeof(double) * (nnz + 8)));
    A->rowptr = rp; A->colidx = ci; A->vals = va;
    k_lap2d<<<(unsigned)((n_local + 1 + 255) / 256), 256, 0, ctx->stream>>>(nx, ny, row0, n_local, col_shift, rp, ci, va);
    LZ_LAUNCH_CHECK(ctx);
    return finish_csr(ctx, A, out);
}

// CSR shadow of a row-interleaved width-4 ELL operator (explicit zeros dropped, ELL column order kept) with the chunk
// schedules: the block path runs the staged SpMM on it instead of a width-4 ELL kernel; the vector path keeps k_ell4_spmv
int lz_ell4_build_shadow(lz_ctx *ctx, lz_matrix *A)
{
    const int64_t n_rows = A->n_rows;
    const unsigned grid = (unsigned)((n_rows + 255) / 256);
    int32_t *cnt, *rp;
    LZ_CUDA(cudaMalloc(&cnt, sizeof(int32_t) * (n_rows + 1)));
    LZ_CUDA(cudaMalloc(&rp, sizeof(int32_t) * (n_rows + 1)));
    LZ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (n_rows + 1), ctx->stream));
    k_ell_count<<<grid, 256, 0, ctx->stream>>>(n_rows, 4, 1, A->ell_data, cnt);
    LZ_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, rp, (int)(n_rows + 1), ctx->stream);
    void *tmp;
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, rp, (int)(n_rows + 1), ctx->stream);
    ctx->launches++;
    int32_t nnz32 = 0;
    LZ_CUDA(cudaMemcpyAsync(&nnz32, rp + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp));
    LZ_CUDA(cudaFree(cnt));
    int32_t *ci; double *va;
    LZ_CUDA(cudaMalloc(&ci, sizeof(int32_t) * ((size_t)nnz32 + 8)));
    LZ_CUDA(cudaMalloc(&va, sizeof(double) * ((size_t)nnz32 + 8)));
    A->rowptr = rp; A->colidx = ci; A->vals = va;
    A->owns_csr = 1; A->csr_nnz = nnz32;
    k_ell_fill<<<grid, 256, 0, ctx->stream>>>(n_rows, 4, 1, A->ell_data, A->ell_idx, rp, ci, va);
    LZ_LAUNCH_CHECK(ctx);
    return build_schedule(ctx, A);
}

lz_matrix *lz_new_matrix(lz_ctx *ctx, int fmt, int64_t n_rows, int64_t n_cols, int64_t nnz) { return new_matrix(ctx, fmt, n_rows, n_cols, nnz); }

extern "C" {

int lz_matrix_ell_view(const lz_matrix *A, const double **data, const uint32_t **idx)
{
    LZ_CHECK(A, LZ_ERR_INVALID, "lz_matrix_ell_view: A is NULL");
    if (data) *data = A->format == LZ_FMT_ELL4 ? A->ell_data : nullptr;
    if (idx) *idx = A->format == LZ_FMT_ELL4 ? A->ell_idx : nullptr;
    return LZ_OK;
}

int lz_csr_create(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *rowptr,
                  const int32_t *colidx, const double *vals, lz_matrix **out)
{
    LZ_CHECK(ctx && out && rowptr && (nnz == 0 || (colidx && vals)), LZ_ERR_INVALID, "lz_csr_create: NULL argument");
    LZ_TRY(check_dims(n_rows, n_cols, nnz));
    LZ_CHECK(((uintptr_t)vals % 16 == 0) && ((uintptr_t)colidx % 8 == 0) && ((uintptr_t)rowptr % 4 == 0), LZ_ERR_INVALID,
             "lz_csr_create: vals must be 16-byte and colidx 8-byte aligned (vectorised loads)");
    LZ_CUDA(cudaSetDevice(ctx->device));
    lz_matrix *A = new_matrix(ctx, LZ_FMT_CSR, n_rows, n_cols, nnz);
    A->rowptr = rowptr; A->colidx = colidx; A->vals = vals;
    return finish_csr(ctx, A, out);
}

int lz_csr_create_host(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *rowptr_host,
                       const int32_t *colidx_host, const double *vals_host, lz_matrix **out)
{
    LZ_CHECK(ctx && out && rowptr_host && (nnz == 0 || (colidx_host && vals_host)), LZ_ERR_INVALID, "lz_csr_create_host: NULL argument");
    LZ_TRY(check_dims(n_rows, n_cols, nnz));
    LZ_CUDA(cudaSetDevice(ctx->device));
    lz_matrix *A = new_matrix(ctx, LZ_FMT_CSR, n_rows, n_cols, nnz);
    A->owns = 1;
    int32_t *rp, *ci;
    double *va;
    LZ_CUDA(cudaMalloc(&rp, sizeof(int32_t) * (n_rows + 1)));
    LZ_CUDA(cudaMalloc(&ci, sizeof(int32_t) * (nnz + 8)));
    LZ_CUDA(cudaMalloc(&va, sizeof(double) * (nnz + 8)));
    A->rowptr = rp; A->colidx = ci; A->vals = va;
    LZ_CUDA(cudaMemcpyAsync(rp, rowptr_host, sizeof(int32_t) * (n_rows + 1), cudaMemcpyHostToDevice, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(ci, colidx_host, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(va, vals_host, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    return finish_csr(ctx, A, out);
}

int lz_ell_create(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int width, int layout, const double *data,
                  const uint32_t *idx, lz_matrix **out)
{
    LZ_CHECK(ctx && out && data && idx, LZ_ERR_INVALID, "lz_ell_create: NULL argument");
    LZ_CHECK(width >= 1 && width <= 1024 && (layout == 0 || layout == 1), LZ_ERR_INVALID, "lz_ell_create: width %d layout %d", width, layout);
    LZ_TRY(check_dims(n_rows, n_cols, n_rows * width));
    LZ_CUDA(cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((n_rows + 255) / 256);
    if (width == 4) {
        lz_matrix *A = new_matrix(ctx, LZ_FMT_ELL4, n_rows, n_cols, n_rows * 4);
        if (layout == 1 && ((uintptr_t)data % 32 == 0) && ((uintptr_t)idx % 16 == 0)) {
            A->ell_data = data; A->ell_idx = idx;
        } else {
            A->owns = 1;
            double *od; uint32_t *oi;
            LZ_CUDA(cudaMalloc(&od, sizeof(double) * n_rows * 4));
            LZ_CUDA(cudaMalloc(&oi, sizeof(uint32_t) * n_rows * 4));
            A->ell_data = od; A->ell_idx = oi;
            if (layout == 0) {
                k_ell4_interleave<<<grid, 256, 0, ctx->stream>>>(n_rows, data, idx, od, oi);
                LZ_LAUNCH_CHECK(ctx);
            } else {
                LZ_CUDA(cudaMemcpyAsync(od, data, sizeof(double) * n_rows * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                LZ_CUDA(cudaMemcpyAsync(oi, idx, sizeof(uint32_t) * n_rows * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        A->max_row_nnz = 4;
        {
            int st = lz_ell4_build_shadow(ctx, A);
            if (st != LZ_OK) { lz_matrix_destroy(A); return st; }
        }
        *out = A;
        return LZ_OK;
    }
    // generic width: compact to CSR on the device (explicit zeros dropped, ELL column order kept)
    lz_matrix *A = new_matrix(ctx, LZ_FMT_CSR, n_rows, n_cols, 0);
    A->owns = 1;
    int32_t *cnt, *rp;
    LZ_CUDA(cudaMalloc(&cnt, sizeof(int32_t) * (n_rows + 1)));
    LZ_CUDA(cudaMalloc(&rp, sizeof(int32_t) * (n_rows + 1)));
    LZ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (n_rows + 1), ctx->stream));
    k_ell_count<<<grid, 256, 0, ctx->stream>>>(n_rows, width, layout, data, cnt);
    LZ_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, rp, (int)(n_rows + 1), ctx->stream);
    void *tmp;
    LZ_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, rp, (int)(n_rows + 1), ctx->stream);
    ctx->launches++;
    int32_t nnz32 = 0;
    LZ_CUDA(cudaMemcpyAsync(&nnz32, rp + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    LZ_CUDA(cudaFree(tmp));
    LZ_CUDA(cudaFree(cnt));
    A->nnz = nnz32;
    A->rowptr = rp;
    int32_t *ci; double *va;
    LZ_CUDA(cudaMalloc(&ci, sizeof(int32_t) * ((size_t)nnz32 + 8)));
    LZ_CUDA(cudaMalloc(&va, sizeof(double) * ((size_t)nnz32 + 8)));
    A->colidx = ci; A->vals = va;
    k_ell_fill<<<grid, 256, 0, ctx->stream>>>(n_rows, width, layout, data, idx, rp, ci, va);
    LZ_LAUNCH_CHECK(ctx);
    return finish_csr(ctx, A, out);
}

int lz_matrix_destroy(lz_matrix *A)
{
    if (!A) return LZ_OK;
    cudaSetDevice(A->device);
    if (A->ctx) {                            // still attached: wait for its stream and leave the context's list
        cudaStreamSynchronize(A->ctx->stream);
        if (A->prev) A->prev->next = A->next; else A->ctx->matrices = A->next;
        if (A->next) A->next->prev = A->prev;
    }                                        // orphaned (context destroyed first): cudaFree synchronises by itself
    if (A->owns || A->owns_csr) {
        cudaFree((void *)A->rowptr);
        cudaFree((void *)A->colidx);
        cudaFree((void *)A->vals);
    }
    if (A->owns) {
        cudaFree((void *)A->ell_data);
        cudaFree((void *)A->ell_idx);
    }
    cudaFree(A->chunk_row);
    cudaFree(A->chunk_ptr);
    cudaFree(A->mm_chunk_row);
    cudaFree(A->mm_chunk_ptr);
    cudaFree(A->chunk_ulen);
    cudaFree(A->mm_chunk_ulen);
    cudaFree(A->vrowptr);
    cudaFree(A->vstart);
    cudaFree(A->ybar);
    cudaFree(A->vpos);
    cudaFree(A->bin_colidx);
    cudaFree(A->bin_vals);
    cudaFree(A->xs_chunk_row); cudaFree(A->xs_chunk_ptr); cudaFree(A->xs_meta); cudaFree(A->xs_seg); cudaFree(A->xs_lidx);
    cudaFree(A->xs_rowptr); cudaFree(A->xs_rowmap); cudaFree(A->xs_vals); cudaFree(A->xs_desc); cudaFree(A->xs_oseg); cudaFree(A->xs_dli);
    if (!A->mm_shared) {
        cudaFree(A->mm.vstart); cudaFree(A->mm.vrowptr); cudaFree(A->mm.vpos);
        cudaFree(A->mm.bin_colidx); cudaFree(A->mm.bin_vals);
    }
    cudaFree(A->mm.long_rows);          // (the SpMV's split keeps its lists in the same struct when the two are shared)
    cudaFree(A->mm.dst);
    if (!A->mm_shared) { cudaFree(A->sv_long_rows); cudaFree(A->sv_dst); }
    delete A;
    return LZ_OK;
}

int lz_matrix_info(const lz_matrix *A, int64_t *n_rows, int64_t *n_cols, int64_t *nnz)
{
    LZ_CHECK(A, LZ_ERR_INVALID, "lz_matrix_info: A is NULL");
    if (n_rows) *n_rows = A->n_rows;
    if (n_cols) *n_cols = A->n_cols;
    if (nnz) *nnz = A->nnz;
    return LZ_OK;
}

int lz_matrix_csr_view(const lz_matrix *A, const int32_t **rowptr, const int32_t **colidx, const double **vals)
{
    LZ_CHECK(A, LZ_ERR_INVALID, "lz_matrix_csr_view: A is NULL");
    const bool csr = A->format == LZ_FMT_CSR;       // (an ELL4 operator's internal CSR shadow is not part of the interface)
    if (rowptr) *rowptr = csr ? A->rowptr : nullptr;
    if (colidx) *colidx = csr ? A->colidx : nullptr;
    if (vals) *vals = csr ? A->vals : nullptr;
    return LZ_OK;
}

int lz_gen_laplacian2d(lz_ctx *ctx, int64_t nx, int64_t ny, lz_matrix **out)
{
    LZ_CHECK(ctx && out && nx >= 2 && ny >= 2, LZ_ERR_INVALID, "lz_gen_laplacian2d: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    return lz_gen_lap2d_rows(ctx, nx, ny, 0, nx * ny, 0, nx * ny, out);
}

int lz_gen_laplacian3d(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, lz_matrix **out)
{
    LZ_CHECK(ctx && out && nx >= 2 && ny >= 2 && nz >= 2, LZ_ERR_INVALID, "lz_gen_laplacian3d: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    return lz_gen_lap3d_rows(ctx, nx, ny, nz, 0, nx * ny * nz, 0, nx * ny * nz, out);
}

int lz_gen_start_vector(lz_ctx *ctx, int64_t n, uint64_t seed, double *v)
{
    LZ_CHECK(ctx && v && n > 0, LZ_ERR_INVALID, "lz_gen_start_vector: bad arguments");
    k_start_vector<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, seed, 0, v);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_gen_start_block(lz_ctx *ctx, int64_t n, int b, int64_t ld, uint64_t seed, double *V)
{
    LZ_CHECK(ctx && V && n > 0 && b > 0 && ld >= n, LZ_ERR_INVALID, "lz_gen_start_block: bad arguments");
    k_start_block<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, b, ld, seed, V);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_gen_rmat_edges(lz_ctx *ctx, int scale, int64_t n_edges, uint64_t seed, int32_t *src, int32_t *dst)
{
    LZ_CHECK(ctx && src && dst && scale >= 1 && scale <= 30 && n_edges > 0, LZ_ERR_INVALID, "lz_gen_rmat_edges: bad arguments");
    k_rmat_edges<<<(unsigned)((n_edges + 255) / 256), 256, 0, ctx->stream>>>(scale, n_edges, seed, src, dst);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_fill(lz_ctx *ctx, int64_t n, double value, double *x)
{
    LZ_CHECK(ctx && (x || n == 0) && n >= 0, LZ_ERR_INVALID, "lz_fill: bad arguments");
    if (n == 0) return LZ_OK;
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, value, x);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

}  // extern "C"

// used by lz_multi.cu for the sharded start vector
int lz_gen_start_vector_offset(lz_ctx *ctx, int64_t n, uint64_t seed, int64_t offset, double *v)
{
    k_start_vector<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, seed, offset, v);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}
