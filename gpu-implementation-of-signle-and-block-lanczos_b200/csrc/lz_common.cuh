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

struct lz_matrix;
struct LzVecRun;
#define LZ_ATTR_CAP 64
struct LzKnobs {
    int spmv_hint, spmm_hint /* -1 auto; bits: 1 matrix streams evict-first, 2 W stores evict-first, 4 Q0 evict-first, 8 X gathers evict-last */, spmm_run, no_split, cgs_shape_order, cgs_upd_mult, cgs_one_cta, no_cgs_fuse, cgs_no_slices;
    int comm_mode;          // LZ_COMM: 0 auto (peer memory when IPC works, else NCCL), 1 NCCL only, 2 peer required
    int no_overlap;         // LZ_NO_OVERLAP: halo exchange on the compute stream, one SpMV launch
    int no_fold;            // LZ_NO_FOLD: keep pass B + separate alpha in the full-reorth vector path
    int spmm_kernel;        // LZ_SPMM_KERNEL: 0 default choice, 1 k_spmm_ws (round-robin chunks), 2 k_spmm_win (staged X window)
    int block_cgs_fuse;     // LZ_BLOCK_CGS_FUSE: 1 (default) fused update+project in the block CGS2, 0 four streams
    int split_l;            // LZ_SPLIT_L: longest virtual row of a row-split (power-law) operator, SpMV
    int no_xs;              // LZ_NO_XS: never use the operand-staging SpMM
    int xs_stages;          // LZ_XS_STAGES: ring depth of the operand-staging SpMM (0 = as many as fit, <= 8)
    int no_direct_rows;     // LZ_NO_DIRECT_ROWS: row-split SpMM writes every row as a partial row (A/B)
    int xs_force;           // LZ_XS_FORCE: use the operand-staging SpMM wherever it can run (default: b = 16 on structured-grid operators)
    int xs_no_tiles;        // LZ_XS_NO_TILES: chunks of consecutive rows even on structured-grid operators
    int xs_tile;            // LZ_XS_TILE: target entries per chunk of its schedule (128..512, default 512)
    int split_l_mm;         // LZ_SPLIT_L_MM: the same for the SpMM's own split (default 32)
    int no_transpose;       // 1 unless LZ_TRANSPOSE is set: the SpMV gather warps walk a chunk in storage order (the transposed walk of uniform chunks measured slower)
    int cgs_rpt;            // LZ_CGS_RPT: rows per thread of the streaming CGS kernels (0 auto, 4, 8)
    int cgs_fuse_min_k;     // LZ_CGS_FUSE_MIN_K: smallest number of basis columns for which CGS2 uses the fused update+project kernel
    int spmm_shape;         // LZ_SPMM_SHAPE: b = 16 SpMM kernel shape (0 shipped; 1: 8 gathers in flight, 7-8 warps; 2: 6 in flight, 9-10 warps)
    int spmm_slice;         // LZ_SPMM_SLICE: columns per pass of the SpMM on row-split (power-law) operators (default 8; >= b: one pass)
    int panel_pad;          // LZ_PANEL_PAD: extra doubles between the block driver's panels (they are 2^k bytes apart on 2^k grids)
    int no_spmm_gram;       // LZ_NO_SPMM_GRAM: the fused b = 16 SpMM leaves Q_j^T W to a separate Gram pass
    int no_spmm_fuse;       // LZ_NO_SPMM_FUSE: b = 16 runs the plain staged SpMM + two-Gram formulation instead of the fused subtraction
    int rmat_reorder;       // LZ_REORDER: locality reordering at lz_csr_create time for power-law operators
};

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
    LzVecRun *vrun;         // state of the current single-vector run (lz_vector_lanczos_begin / _advance / checkpoints)
    int last_coupling_slot; // index into `scalars` of beta_m left by the last driver run (lz_last_coupling)
    // optional per-kernel-class timing (bench.py's roofline leg): CUDA events on ctx->stream
    int prof_on;
    int prof_used;
    cudaEvent_t *prof_ev;   // 2 * LZ_PROF_CAP events
    int *prof_cls;          // class id per pair
    double *prof_bytes;     // algorithmic bytes per pair
    int spmv_variant;       // dev-time A/B knob (env LZ_SPMV_VARIANT: 3 coarse schedule, 20 fine schedule, 9 no staged SpMM)
    // per-class sums drained from the event ring whenever it fills (lz_ctx.cu), so long timed regions are not truncated
    int64_t prof_acc_launches[16];
    double prof_acc_ms[16], prof_acc_bytes[16];
    // environment knobs, read ONCE per context in lz_ctx_create -- no process-wide statics
    LzKnobs knobs;
    // kernels whose dynamic shared-memory opt-in has been applied on THIS context's device
    // (cudaFuncSetAttribute is per device: a process-wide flag would leave a second GPU without it)
    const void *attr_funcs[LZ_ATTR_CAP];
    int n_attr;
    // operators created on this context (intrusive list): lz_ctx_destroy orphans them, so destroying a
    // matrix after its context is safe (lz_matrix_destroy never dereferences a dead context)
    lz_matrix *matrices;
    // side stream + events for the halo exchange overlapped with the interior SpMV (lz_multi.cu)
    cudaStream_t side_stream;
    cudaEvent_t ev_ready, ev_halo;
};

int lz_func_smem_optin(lz_ctx *ctx, const void *func, int bytes, bool carveout_max = false);

// ---- persistent state of a single-vector run (lz_vector.cu): begin once, advance in pieces, checkpoint, restart
struct LzCgs {
    double *V; int64_t ts, cs; double *cpart; double *c; unsigned grid;      // basis element (i,k): V[(i>>5)*ts + k*cs + (i&31)]
    int rpt;                     // rows per thread of the streaming CGS kernels (8, or 4 for small shards)
};
struct LzVecRun {
    const lz_matrix *A;
    int64_t n, hlo, hhi, n_below, lc, stride;
    int m;                       // capacity: steps / basis columns this run was set up for
    int reorth;
    bool sharded, fold, overlap;
    double *u_prev, *u_cur, *w;  // rotating work vectors (unnormalised q_{j-1}, q_j, scratch)
    LzCgs g;
    double *alpha, *beta, *invb; // device scalar series in the context's bank: alpha[m], beta[m+1], invb[m+1]
    double *q;                   // optional receiver-row series (device, m)
    double *om[3];               // selective reorthogonalisation: omega rows (old, current, new)
    int j;                       // next step
    int first_next;              // the next step has no q_{j-1} term (start of a run / first step after a thick restart)
};

#define LZ_PROF_CAP 32768
// kernel classes for the profiler
enum { LZ_K_SPMV = 0, LZ_K_PASSB = 1, LZ_K_PROJECT = 2, LZ_K_UPDATE = 3, LZ_K_SPMM = 4, LZ_K_GRAM = 5, LZ_K_PANEL = 6, LZ_K_SMALL = 7, LZ_K_COMM = 8, LZ_K_UPDPROJ = 9, LZ_K_CLASSES = 10 };
static_assert(LZ_K_CLASSES == LZ_PROFILE_CLASSES, "profiler class count is part of the C-ABI");
void lz_prof_begin(lz_ctx *ctx, int cls, double bytes);
void lz_prof_end(lz_ctx *ctx);
#define LZ_PARTIALS_CAP (1 << 22)
#define LZ_TICKETS 64
#define LZ_SCALARS 16384
#define LZ_FLAGS 64

// ---- peer-memory communication (lz_multi.cu) ---------------------------------------------------
// Every rank owns one cudaMalloc'ed ARENA that all other ranks of the box map through CUDA IPC:
//   [flags | all-reduce slots (2 parities x world senders x cap doubles) | user region (work vectors / panels)]
// Small reductions and the halo exchange are then plain NVLink stores into the peer's arena followed by a
// release-store of a sequence number; the receiver spins on its own flag with acquire loads and adds the
// deposits in rank order (bit-identical result on every rank).  The descriptor lives in device memory.
#define LZ_MAX_RANKS 16
struct LzPeerDesc {
    int world, rank;
    unsigned long long cap;                               // doubles per all-reduce slot
    double *slot_at[2][LZ_MAX_RANKS];                     // where THIS rank deposits at peer q (q's slot[parity][rank])
    unsigned long long *flag_at[2][LZ_MAX_RANKS];         // peer q's flag[parity][rank]
    double *my_slot[2];                                   // my slot[parity][0]; sender r at + r * cap
    unsigned long long *my_flag[2];                       // my flag[parity][0 .. world)
    unsigned long long *halo_flag_at[2];                  // [0] lower neighbour's "from above" flag, [1] upper neighbour's "from below" flag
    unsigned long long *my_halo_flag;                     // [0] set by my lower neighbour, [1] by my upper neighbour
    int *err;                                             // set to 1 when a wait gave up (peer lost): results are garbage, no hang
};
#define LZ_PEER_SPIN_LIMIT (1u << 27)

// single-vector run in pieces (lz_vector.cu), used by the checkpoint / thick-restart code in lz_eigs.cu
int lz_vec_setup(lz_ctx *ctx, const lz_matrix *A, int m, int64_t lc, int reorth, double *q);
int lz_vec_start(lz_ctx *ctx, const double *b);
int lz_vec_steps(lz_ctx *ctx, int j_end);
int lz_vec_report(lz_ctx *ctx, double *alpha_host, double *beta_host, int *steps_done);

int lz_ctx_workspace(lz_ctx *ctx, size_t bytes, void **out);   // grow-only scratch
int lz_ctx_basis(lz_ctx *ctx, int64_t rows, int cols, double **out);            // vector path: row-tiled slab
int lz_ctx_basis_blocks(lz_ctx *ctx, int64_t pan, int blocks, double **out);  // block path: back-to-back panels
int lz_ctx_scratch(lz_ctx *ctx, size_t bytes, void **out);
// communicator hooks (lz_multi.cu); all enqueue on ctx->stream
int lz_comm_world(const lz_ctx *ctx);
int lz_comm_rank(const lz_ctx *ctx);
// epilogue applied by the LAST step of an all-reduce (inside the peer kernel, or a one-thread kernel after NCCL)
struct LzArEpi {
    double *copy_dst; int copy_idx;     // copy_dst[0] = buf[copy_idx]           (alpha_j = c1[j]); NULL: skip
    double *beta, *invb; int *flags;    // beta[jn] = sqrt(buf[0]), invb[jn] = 1/beta[jn], breakdown flag; beta NULL: skip
    int jn;
    int *dgks_flag; const double *dgks_before;   // *dgks_flag = buf[0] < 0.5 * *dgks_before (DGKS test on the reduced norm); NULL: skip
};
int lz_comm_allreduce_sum(lz_ctx *ctx, double *buf, size_t count, const LzArEpi *epi = nullptr);
// u points at the local rows; hlo entries are received just below it, hhi entries just above u[n).
// n_below = number of entries the LOWER neighbour owns (its upper halo starts n_below past its own u); peer mode only.
// side: run on the context's side stream between ev_ready (recorded on the compute stream by the caller) and ev_halo
int lz_comm_halo_exchange(lz_ctx *ctx, double *u, int64_t n, int64_t hlo, int64_t hhi, int64_t n_below = -1, bool side = false);
// peer-memory mode: 1 when the arena is mapped on every rank
int lz_comm_peer(const lz_ctx *ctx);
const LzPeerDesc *lz_comm_desc(const lz_ctx *ctx);
unsigned long long lz_comm_next_seq(lz_ctx *ctx);     // sequence number of the next in-kernel all-reduce
// collective: a user region of at least user_bytes in the symmetric arena and all-reduce slots of ar_cap doubles
// (falls back to the context's private workspace + NCCL when peer mapping is unavailable)
int lz_comm_arena(lz_ctx *ctx, size_t user_bytes, size_t ar_cap, void **user);
// host-side all-gather of four int64 per rank (layout negotiation at the start of a sharded solve)
int lz_comm_gather4(lz_ctx *ctx, const int64_t mine[4], int64_t *all /* world * 4 */);

// ---- sparse operator -----------------------------------------------------------------
enum { LZ_FMT_CSR = 0, LZ_FMT_ELL4 = 1 };

// row-aligned nnz chunks: chunk c covers rows [chunk_row[c], chunk_row[c+1])
#define LZ_SPMV_THREADS 256
#define LZ_SPMV_TILE 768         // target nnz per SpMV chunk (profiles/r01_spmv_variants.md)
#define LZ_XS_TILE 512          // target nnz per chunk of the operand-staging SpMM
#define LZ_XS_SEGCAP 32         // column segments per chunk (one bulk copy per producer lane)
#define LZ_XS_ECAP 1024         // entries per chunk the build kernel sorts (tile + longest row must fit)
#define LZ_XS_OGROUPS 16        // 8-row output groups per chunk (chunks hold <= 128 rows)
#define LZ_XS_MERGE 8           // columns at most this far apart share a segment
#define LZ_SPMM_TILE 1536        // target nnz per SpMM chunk (k_spmm_ws stages 2048 entries per slot)
#define LZ_SPLIT_L 256           // rows longer than this are split into virtual rows
#define LZ_SPMV_CAP 4096         // shared-memory product slots per CTA (32 KB)

// one row split of a power-law operator: virtual rows of <= L entries, length-binned (see lz_csr.cu)
struct LzSplit {
    int n_virtual;
    int32_t *vstart, *vrowptr, *vpos, *bin_colidx;
    double *bin_vals;
    // rows cut into more than LZ_LONG_PIECES pieces (the hubs): their ordered combine gets a CTA each, a thread walking
    // 12 000 pieces one after the other was the longest thing in the whole product
    int32_t *long_rows;
    int n_long;
    // where the product's row at (binned) position i goes: >= 0 the operator row it IS (a row that was not cut: written
    // straight into W), < 0 position ~i of the partial-row buffer (combined afterwards)
    int32_t *dst;
};
#define LZ_LONG_PIECES 32

struct lz_matrix {
    lz_ctx *ctx;             // NULL once the owning context has been destroyed (orphaned: only destroy is legal)
    lz_matrix *next, *prev;  // the context's list of live operators
    int device;              // copied at creation so destroy never needs the context
    int format;
    int64_t n_rows, n_cols, nnz;
    const int32_t *rowptr;   // CSR
    const int32_t *colidx;
    const double *vals;
    const double *ell_data;  // ELL4 row-interleaved
    const uint32_t *ell_idx;
    int owns;                // arrays owned by the library (freed on destroy)
    int owns_csr;            // ELL4 operators: the CSR shadow (rowptr/colidx/vals + schedules) built for the block path
    int64_t csr_nnz;         // entries of the CSR arrays (== nnz for CSR operators; true non-zeros of an ELL4 shadow)
    int n_chunks;
    int32_t *chunk_row;      // n_chunks + 1
    int32_t *chunk_ptr;      // rowptr[chunk_row[c]], n_chunks + 1
    int tile, cap;           // nnz per chunk (target) and shared-memory product slots per CTA
    int mm_n_chunks;         // second schedule with LZ_SPMM_TILE-sized chunks for the SpMM kernel
    int32_t *mm_chunk_row, *mm_chunk_ptr;
    int32_t *chunk_ulen, *mm_chunk_ulen;   // per chunk: the common (odd) row length, or 0 (lz_csr.cu: k_chunk_ulen)
    int tma_ok;              // vals / colidx 16-byte aligned: bulk-copy staged kernel usable
    int max_row_nnz;
    // row-split view for operators with long rows (NULL otherwise): virtual row pointers over the same
    // colidx/vals, first virtual row of every real row, and the per-virtual-row partial sums
    int64_t n_virtual;
    int32_t *vrowptr;        // n_virtual + 1
    int32_t *vstart;         // n_rows + 1
    double *ybar;            // n_virtual
    // length-binned order of the virtual rows (power-law operators, lz_csr.cu): vrowptr then indexes the binned copies
    // bin_colidx / bin_vals, and piece v (in vstart numbering) of a row is found at position vpos[v]
    int32_t *vpos, *bin_colidx;
    double *bin_vals;
    const int32_t *k_colidx; // the arrays the SpMV kernels stream (binned copies when present)
    const double *k_vals;
    // The SpMM's own row split (lz_csr.cu, lz_matrix_prepare_mm): the panel product wants SHORT virtual rows (a lane group
    // gathers one 128..256-byte panel row per entry and a 256-entry row leaves the other groups of its chunk idle), the
    // SpMV wants long ones (fewer partial sums).  Built on the first panel product unless both lengths agree (then shared).
    LzSplit mm;
    int32_t *sv_long_rows, *sv_dst;
    int mm_pending, mm_shared;
    const int32_t *mm_k_colidx;
    const double *mm_k_vals;
    // X-window schedule of the operand-staging SpMM (lz_spmm_xs.cuh; built by lz_matrix_prepare_xs on the first panel
    // product): chunks of ~LZ_XS_TILE entries; per chunk the distinct columns it references, merged into <= LZ_XS_SEGCAP
    // contiguous segments (first column, rows) -- the rows of X the chunk gathers, bulk-copied into shared memory -- and
    // per entry the 16-bit index of its column's row inside that window (streamed instead of the 32-bit column index)
    int xs_state;                      // 0 not tried yet, 1 usable, -1 this operator does not fit (irregular / long rows)
    int xs_n_chunks, xs_max_wrows, xs_max_rows, xs_max_entries;
    int32_t *xs_chunk_row, *xs_chunk_ptr;
    int2 *xs_meta;                     // per chunk: (segments, window rows)
    int2 *xs_seg;                      // [chunk][LZ_XS_SEGCAP]: (first column, rows)
    uint16_t *xs_lidx;                 // per entry
    // On structured-grid operators the chunks are small BOXES of grid points (lx x ty x tz rows: ty*tz short runs of
    // consecutive rows) instead of runs of consecutive rows, because a box shares far more of its neighbours (7-point
    // stencil: 2.7 window rows per row instead of 5.1).  The product walks a row-permuted copy of the operator --
    // xs_rowptr / xs_vals in chunk order, every chunk padded to a multiple of 8 entries; xs_rowmap[i] = original row of
    // walked row i; column numbering unchanged -- and writes row xs_rowmap[i] of W.
    int32_t *xs_rowptr, *xs_rowmap;
    double *xs_vals;
    int4 *xs_desc;                     // per chunk: (first entry, end entry, first row, end row) of the walked operator
    uint16_t *xs_dli;                  // per walked row: window row of X[its own row] (0xFFFF: not in the chunk's window)
    int32_t *xs_oseg;                  // [chunk][LZ_XS_OGROUPS]: original row of every 8th walked row (-1 past the chunk): L2 prefetch of Q0
    int xs_tile_dims[3];               // lx, ty, tz (0: not tiled)
    // sharded operators: local rows only, columns in [0, n_local + halo_lo + halo_hi)
    int64_t halo_lo, halo_hi;        // halo entries below / above the local range
    // chunks [bnd_lo, bnd_hi) of the fine schedule (mm_*: of the coarse one) hold only rows that reference no halo
    // column: they can run while the halo exchange is still in flight (lz_multi.cu fills these for shards)
    int has_split, bnd_lo, bnd_hi, mm_bnd_lo, mm_bnd_hi;
    int64_t global_rows, row_begin;  // position in the global operator
};

int lz_ell4_build_shadow(lz_ctx *ctx, lz_matrix *A);                      // lz_csr.cu
int lz_matrix_prepare_xs(lz_ctx *ctx, const lz_matrix *A);                // lz_csr.cu: build the X-window schedule (sets xs_state)
int lz_matrix_prepare_mm(lz_ctx *ctx, const lz_matrix *A);                // lz_csr.cu: build the SpMM's row split if it is pending
lz_matrix *lz_new_matrix(lz_ctx *ctx, int fmt, int64_t n_rows, int64_t n_cols, int64_t nnz);

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

// ---- system-scope flag traffic of the peer-memory collectives
__device__ __forceinline__ void lz_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lz_ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double lz_ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lz_st_relaxed_sys(double *p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// wait until *flag >= seq (bounded: a lost peer must not hang the GPU)
__device__ __forceinline__ void lz_peer_wait(const unsigned long long *flag, unsigned long long seq, int *err)
{
    unsigned int spins = 0;
    while (lz_ld_acquire_sys(flag) < seq) {
        if (++spins > LZ_PEER_SPIN_LIMIT) { *err = 1; break; }
    }
}
// One thread all-reduces NS scalars over the ranks of the box: deposit at every peer, release the flag, wait
// for every peer's deposit, add in rank order.  Called by thread 0 of the last CTA of a reduction kernel.
template <int NS>
__device__ __forceinline__ void lz_peer_sum_thread(const LzPeerDesc *pd, unsigned long long seq, double *v)
{
    const int p = (int)(seq & 1ull), R = pd->world, me = pd->rank;
    for (int q = 0; q < R; ++q) {
        if (q == me) continue;
#pragma unroll
        for (int s = 0; s < NS; ++s) lz_st_relaxed_sys(pd->slot_at[p][q] + s, v[s]);
    }
    __threadfence_system();
    for (int q = 0; q < R; ++q)
        if (q != me) lz_st_release_sys(pd->flag_at[p][q], seq);
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0;
    for (int q = 0; q < R; ++q) {
        if (q == me) {
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] += v[s];
        } else {
            lz_peer_wait(pd->my_flag[p] + q, seq, pd->err);
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] += lz_ld_relaxed_sys(pd->my_slot[p] + (size_t)q * pd->cap + s);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) v[s] = acc[s];
}

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
