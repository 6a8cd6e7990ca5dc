// lz_common.cuh -- shared internals of liblanczos_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lanczos_b200.h"

#define LZ_SM_COUNT_DEFAULT 148
#define LZ_WARP 32

// ---- error plumbing ------------------------------------------------------------------
void lz_set_error(const char *fmt, ...);

#define LZ_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            lz_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (e__ == cudaErrorMemoryAllocation) ? LZ_ERR_ALLOC : LZ_ERR_CUDA;            \
        }                                                                                      \
    } while (0)

#define LZ_CHECK(cond, code, ...)                                                              \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            lz_set_error(__VA_ARGS__);                                                         \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

#define LZ_TRY(call)                                                                           \
    do {                                                                                       \
        int s__ = (call);                                                                      \
        if (s__ != LZ_OK) return s__;                                                          \
    } while (0)

// launch check: reference never checks launches; we do, cheaply (no sync).
#define LZ_LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t e__ = cudaPeekAtLastError();                                               \
        if (e__ != cudaSuccess) {                                                              \
            lz_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return LZ_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// ---- context -------------------------------------------------------------------------
struct lz_comm;   // lz_multi.cu

struct lz_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    int64_t launches;
    // reduction scratch: per-CTA partials + ticket counters for the last-block-done pattern
    double *partials;       // LZ_PARTIALS_CAP doubles
    unsigned int *tickets;  // LZ_TICKETS ints, zero-initialised, self-resetting
    double *scalars;        // device scalar bank (LZ_SCALARS doubles)
    int *flags;             // device int flags (LZ_FLAGS)
    // generic workspace (grown on demand, outside timed regions after the first call)
    void *work;
    size_t work_bytes;
    void *scratch;          // small second scratch (Gram partials) that never aliases `work`
    size_t scratch_bytes;
    // Krylov basis slab kept by the full-reorth drivers
    double *basis;
    size_t basis_bytes;
    int64_t basis_ts, basis_cs, basis_rows;   // element (i,k) at basis[(i>>5)*ts + k*cs + (i&31)]
    int basis_cols;
    lz_comm *comm;
    // optional per-kernel-class timing (bench.py's roofline leg): CUDA events on ctx->stream
    int prof_on;
    int prof_used;
    cudaEvent_t *prof_ev;   // 2 * LZ_PROF_CAP events
    int *prof_cls;          // class id per pair
    double *prof_bytes;     // algorithmic bytes per pair
    int spmv_variant;       // dev-time A/B knob (env LZ_SPMV_VARIANT: 3 coarse schedule, 20 fine schedule, 9 no staged SpMM)
};

#define LZ_PROF_CAP 16384
// kernel classes for the profiler
enum { LZ_K_SPMV = 0, LZ_K_PASSB = 1, LZ_K_PROJECT = 2, LZ_K_UPDATE = 3, LZ_K_SPMM = 4, LZ_K_GRAM = 5, LZ_K_PANEL = 6, LZ_K_SMALL = 7, LZ_K_COMM = 8, LZ_K_UPDPROJ = 9, LZ_K_CLASSES = 10 };
static_assert(LZ_K_CLASSES == LZ_PROFILE_CLASSES, "profiler class count is part of the C-ABI");
void lz_prof_begin(lz_ctx *ctx, int cls, double bytes);
void lz_prof_end(lz_ctx *ctx);
#define LZ_PARTIALS_CAP (1 << 22)
#define LZ_TICKETS 64
#define LZ_SCALARS 8192
#define LZ_FLAGS 64

int lz_ctx_workspace(lz_ctx *ctx, size_t bytes, void **out);   // grow-only scratch
int lz_ctx_basis(lz_ctx *ctx, int64_t rows, int cols, double **out);            // vector path: row-tiled slab
int lz_ctx_basis_blocks(lz_ctx *ctx, int64_t pan, int blocks, double **out);  // block path: back-to-back panels
int lz_ctx_scratch(lz_ctx *ctx, size_t bytes, void **out);
// communicator hooks (lz_multi.cu); all enqueue on ctx->stream
int lz_comm_world(const lz_ctx *ctx);
int lz_comm_rank(const lz_ctx *ctx);
int lz_comm_allreduce_sum(lz_ctx *ctx, double *buf, size_t count);
// u points at the local rows; hlo entries are received just below it, hhi entries just above u[n)
int lz_comm_halo_exchange(lz_ctx *ctx, double *u, int64_t n, int64_t hlo, int64_t hhi);

// ---- sparse operator -----------------------------------------------------------------
enum { LZ_FMT_CSR = 0, LZ_FMT_ELL4 = 1 };

// row-aligned nnz chunks: chunk c covers rows [chunk_row[c], chunk_row[c+1])
#define LZ_SPMV_THREADS 256
#define LZ_SPMV_TILE 768         // target nnz per SpMV chunk (profiles/r01_spmv_variants.md)
#define LZ_SPMM_TILE 1536        // target nnz per SpMM chunk (k_spmm_ws stages 2048 entries per slot)
#define LZ_SPLIT_L 256           // rows longer than this are split into virtual rows
#define LZ_SPMV_CAP 4096         // shared-memory product slots per CTA (32 KB)

struct lz_matrix {
    lz_ctx *ctx;
    int format;
    int64_t n_rows, n_cols, nnz;
    const int32_t *rowptr;   // CSR
    const int32_t *colidx;
    const double *vals;
    const double *ell_data;  // ELL4 row-interleaved
    const uint32_t *ell_idx;
    int owns;                // arrays owned by the library (freed on destroy)
    int n_chunks;
    int32_t *chunk_row;      // n_chunks + 1
    int32_t *chunk_ptr;      // rowptr[chunk_row[c]], n_chunks + 1
    int tile, cap;           // nnz per chunk (target) and shared-memory product slots per CTA
    int mm_n_chunks;         // second schedule with LZ_SPMM_TILE-sized chunks for the SpMM kernel
    int32_t *mm_chunk_row, *mm_chunk_ptr;
    int tma_ok;              // vals / colidx 16-byte aligned: bulk-copy staged kernel usable
    int max_row_nnz;
    // row-split view for operators with long rows (NULL otherwise): virtual row pointers over the same
    // colidx/vals, first virtual row of every real row, and the per-virtual-row partial sums
    int64_t n_virtual;
    int32_t *vrowptr;        // n_virtual + 1
    int32_t *vstart;         // n_rows + 1
    double *ybar;            // n_virtual
    // sharded operators: local rows only, columns in [0, n_local + halo_lo + halo_hi)
    int64_t halo_lo, halo_hi;        // halo entries below / above the local range
    int64_t global_rows, row_begin;  // position in the global operator
};

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double lz_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0.  `red` is >= 32 doubles of shared memory.
template <int THREADS>
__device__ __forceinline__ double lz_block_sum(double v, double *red)
{
    v = lz_warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < THREADS / 32) ? red[lane] : 0.0;
        v = lz_warp_sum(v);
    }
    return v;
}

// Deterministic grid reduction, last-block-done: thread 0 of every CTA stores its NS partials, the
// last CTA to arrive sums all partials in a fixed order (independent of arrival order) and gets
// `true` back in all its threads with the totals valid in thread 0.  The ticket resets itself, so
// kernels that run one after another on a stream can share it.
template <int THREADS, int NS>
__device__ __forceinline__ bool lz_grid_sum(const double *mine, double *partials, unsigned int *ticket,
                                            double *red, double *total)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) partials[(size_t)blockIdx.x * NS + s] = mine[s];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double acc = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) acc += __ldcg(&partials[(size_t)i * NS + s]);
        acc = lz_block_sum<THREADS>(acc, red);
        if (threadIdx.x == 0) total[s] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0;
    return true;
}

__device__ __forceinline__ uint64_t lz_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double lz_u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

// 256-bit global load of four doubles (sm_100: LDG.E.ENL2.256); pointer must be 32-byte aligned
__device__ __forceinline__ void lz_ld256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// streaming variant for data that is read exactly once (the Krylov basis): evict-first
__device__ __forceinline__ void lz_ld256_stream(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void lz_st256(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// streaming store (written once, not re-read before it would be evicted anyway)
__device__ __forceinline__ void lz_st256_stream(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

#endif  // __CUDACC__
