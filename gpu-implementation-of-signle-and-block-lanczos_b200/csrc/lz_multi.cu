// lz_multi.cu -- multi-GPU plumbing: one process per GPU, rows of A and of every Krylov vector
// sharded in contiguous blocks.  The reference has no distributed code at all; this
// file is an addition in its style.
//
// Per Lanczos step the path has two real exchanges: (1) the boundary planes of q_j go to the two
// neighbouring ranks before the SpMV (halo exchange: grouped ncclSend/ncclRecv over NVLink, 2 MB per
// neighbour at 512^3), (2) the alpha / beta^2 / CGS-coefficient partial sums are all-reduced.
// NCCL is resolved with dlopen at lz_comm_init time so the single-GPU library has no link-time
// dependency on it and shares whichever libnccl the host process (torch) has already loaded.
#include <dlfcn.h>

#include "lz_common.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_INT8 = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *handle;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char *(*GetErrorString)(int);
};

static NcclApi g_nccl = {};

static int load_nccl()
{
    if (g_nccl.handle) return LZ_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        lz_set_error("lz_comm: cannot load libnccl.so.2 (%s)", dlerror());
        return LZ_ERR_COMM;
    }
#define LOAD(field, sym)                                                         \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                                   \
    if (!g_nccl.field) { lz_set_error("lz_comm: symbol %s missing in libnccl", sym); return LZ_ERR_COMM; }
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(AllGather, "ncclAllGather");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.handle = h;
    return LZ_OK;
}

#define LZ_NCCL(call)                                                                           \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != NCCL_SUCCESS) {                                                              \
            lz_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return LZ_ERR_COMM;                                                                 \
        }                                                                                       \
    } while (0)

#define LZ_ARENA_FLAG_BYTES 4096          // flag[2][LZ_MAX_RANKS] + halo flags, padded

struct lz_comm {
    ncclComm_t comm;
    int world, rank;
    // bootstrap staging (device): 64-byte IPC handles / four int64 per rank
    char *stage;                          // (1 + world) * 64 bytes
    // peer-memory arena
    int peer_ok;                          // arena mapped on every rank
    int peer_tried;                       // IPC mapping failed once: stay on NCCL
    char *arena;                          // my arena
    char *peer[LZ_MAX_RANKS];             // rank q's arena mapped into this process (peer[rank] = arena)
    size_t ar_cap;                        // doubles per all-reduce slot
    size_t user_off, user_bytes;          // user region
    LzPeerDesc *desc_dev;                 // device copy of the descriptor
    int *err_dev;
    unsigned long long ar_seq, halo_seq;  // collectives issued so far (identical on every rank)
    unsigned int *halo_ticket;
};

int lz_comm_world(const lz_ctx *ctx) { return ctx->comm ? ctx->comm->world : 1; }
int lz_comm_rank(const lz_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }
int lz_comm_peer(const lz_ctx *ctx) { return ctx->comm ? ctx->comm->peer_ok : 0; }
const LzPeerDesc *lz_comm_desc(const lz_ctx *ctx) { return (ctx->comm && ctx->comm->peer_ok) ? ctx->comm->desc_dev : nullptr; }
unsigned long long lz_comm_next_seq(lz_ctx *ctx) { return ++ctx->comm->ar_seq; }

// ---------------------------------------------------------------------------------------------
// epilogue of a reduction: alpha_j = c1[j] and / or beta_{j+1} = sqrt(||w||^2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lz_ar_epilogue(const LzArEpi &e, const double *buf)
{
    if (e.copy_dst) *e.copy_dst = buf[e.copy_idx];
    if (e.beta) {
        const double t = buf[0], b = sqrt(t);
        e.beta[e.jn] = b;
        e.invb[e.jn] = 1.0 / b;
        if (!isfinite(t) || t == 0.0) atomicMin(e.flags, e.jn);
    }
    if (e.dgks_flag) *e.dgks_flag = (buf[0] < 0.5 * (*e.dgks_before)) ? 1 : 0;
}

__global__ void k_ar_epilogue(const LzArEpi e, const double *buf) { lz_ar_epilogue(e, buf); }

// One-shot all-reduce over peer memory, one CTA: deposit buf at every peer (NVLink stores), publish the
// sequence number, wait for every peer's deposit, add in rank order (same bits on every rank).
__global__ void __launch_bounds__(512)
k_peer_allreduce(const LzPeerDesc *__restrict__ pd, unsigned long long seq, double *__restrict__ buf, int count, const LzArEpi epi, int has_epi)
{
    const int p = (int)(seq & 1ull), R = pd->world, me = pd->rank, tid = threadIdx.x;
    for (int q = 0; q < R; ++q) {
        if (q == me) continue;
        double *dst = pd->slot_at[p][q];
        for (int i = tid; i < count; i += 512) lz_st_relaxed_sys(dst + i, buf[i]);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < R && tid != me) {
        lz_st_release_sys(pd->flag_at[p][tid], seq);
        lz_peer_wait(pd->my_flag[p] + tid, seq, pd->err);
    }
    __syncthreads();
    const double *slots = pd->my_slot[p];
    const size_t cap = pd->cap;
    for (int i = tid; i < count; i += 512) {
        double acc = 0.0;
        for (int q = 0; q < R; ++q) acc += (q == me) ? buf[i] : lz_ld_relaxed_sys(slots + (size_t)q * cap + i);
        buf[i] = acc;
    }
    if (has_epi) {
        __syncthreads();
        if (tid == 0) lz_ar_epilogue(epi, buf);
    }
}

int lz_comm_allreduce_sum(lz_ctx *ctx, double *buf, size_t count, const LzArEpi *epi)
{
    lz_comm *c = ctx->comm;
    lz_prof_begin(ctx, LZ_K_COMM, 8.0 * (double)count);
    LzArEpi e;
    memset(&e, 0, sizeof(e));
    if (epi) e = *epi;
    if (c->peer_ok && count <= c->ar_cap) {
        k_peer_allreduce<<<1, 512, 0, ctx->stream>>>(c->desc_dev, ++c->ar_seq, buf, (int)count, e, epi ? 1 : 0);
        LZ_LAUNCH_CHECK(ctx);
    } else {
        LZ_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, c->comm, ctx->stream));
        if (epi) {
            k_ar_epilogue<<<1, 1, 0, ctx->stream>>>(e, buf);
            LZ_LAUNCH_CHECK(ctx);
        }
    }
    lz_prof_end(ctx);
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// halo exchange
// ---------------------------------------------------------------------------------------------
// peer mode: every CTA copies a slice of this rank's boundary rows straight into the neighbours' halo regions;
// the last CTA to finish publishes the sequence number at both neighbours and waits for theirs, so the halo of
// THIS rank is complete when the kernel retires.
__global__ void __launch_bounds__(256)
k_peer_halo(const LzPeerDesc *__restrict__ pd, unsigned long long seq, const double *__restrict__ src_lo, double *__restrict__ dst_lo,
            int64_t cnt_lo, const double *__restrict__ src_hi, double *__restrict__ dst_hi, int64_t cnt_hi, unsigned int *ticket)
{
    const int64_t stride = (int64_t)gridDim.x * 256, t0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool v2 = (((uintptr_t)src_lo | (uintptr_t)dst_lo | (uintptr_t)src_hi | (uintptr_t)dst_hi) % 16 == 0) && ((cnt_lo | cnt_hi) % 2 == 0);
    if (v2) {
        for (int64_t i = t0; i < cnt_lo / 2; i += stride) reinterpret_cast<double2 *>(dst_lo)[i] = reinterpret_cast<const double2 *>(src_lo)[i];
        for (int64_t i = t0; i < cnt_hi / 2; i += stride) reinterpret_cast<double2 *>(dst_hi)[i] = reinterpret_cast<const double2 *>(src_hi)[i];
    } else {
        for (int64_t i = t0; i < cnt_lo; i += stride) dst_lo[i] = src_lo[i];
        for (int64_t i = t0; i < cnt_hi; i += stride) dst_hi[i] = src_hi[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        *ticket = 0;
        if (cnt_lo > 0) lz_st_release_sys(pd->halo_flag_at[0], seq);
        if (cnt_hi > 0) lz_st_release_sys(pd->halo_flag_at[1], seq);
        if (cnt_lo > 0) lz_peer_wait(pd->my_halo_flag + 0, seq, pd->err);
        if (cnt_hi > 0) lz_peer_wait(pd->my_halo_flag + 1, seq, pd->err);
    }
}

int lz_comm_halo_exchange(lz_ctx *ctx, double *u, int64_t n, int64_t hlo, int64_t hhi, int64_t n_below, bool side)
{
    lz_comm *c = ctx->comm;
    cudaStream_t st = ctx->stream;
    if (side) {
        if (!ctx->side_stream) {
            LZ_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
            LZ_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
            LZ_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        }
        st = ctx->side_stream;
        LZ_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
        LZ_CUDA(cudaStreamWaitEvent(st, ctx->ev_ready, 0));
    } else {
        lz_prof_begin(ctx, LZ_K_COMM, 8.0 * (double)(hlo + hhi) * 2.0);
    }
    const bool lower = c->rank > 0 && hlo > 0, upper = c->rank < c->world - 1 && hhi > 0;
    const char *ub = (const char *)u;
    const bool in_arena = c->peer_ok && ub >= c->arena + c->user_off && ub + sizeof(double) * (size_t)(n + hhi) <= c->arena + c->user_off + c->user_bytes;
    if (in_arena && n_below >= 0) {
        // the arena layout is identical on every rank, so the neighbour's copy of this buffer sits at the same offset
        const size_t off = (size_t)(ub - c->arena);
        double *dst_lo = lower ? reinterpret_cast<double *>(c->peer[c->rank - 1] + off) + n_below : nullptr;   // its upper halo
        double *dst_hi = upper ? reinterpret_cast<double *>(c->peer[c->rank + 1] + off) - hhi : nullptr;       // its lower halo
        const int64_t bytes = 8 * ((lower ? hlo : 0) + (upper ? hhi : 0));
        int grid = (int)((bytes + 65535) / 65536);
        if (grid < 1) grid = 1;
        if (grid > ctx->sm_count) grid = ctx->sm_count;
        k_peer_halo<<<grid, 256, 0, st>>>(c->desc_dev, ++c->halo_seq, u, dst_lo, lower ? hlo : 0, u + n - hhi, dst_hi, upper ? hhi : 0, c->halo_ticket);
        LZ_LAUNCH_CHECK(ctx);
    } else {
        LZ_NCCL(g_nccl.GroupStart());
        if (lower) {
            // my first hlo rows are the lower neighbour's upper halo; its last rows are my lower halo
            LZ_NCCL(g_nccl.Send(u, (size_t)hlo, NCCL_FLOAT64, c->rank - 1, c->comm, st));
            LZ_NCCL(g_nccl.Recv(u - hlo, (size_t)hlo, NCCL_FLOAT64, c->rank - 1, c->comm, st));
        }
        if (upper) {
            LZ_NCCL(g_nccl.Send(u + n - hhi, (size_t)hhi, NCCL_FLOAT64, c->rank + 1, c->comm, st));
            LZ_NCCL(g_nccl.Recv(u + n, (size_t)hhi, NCCL_FLOAT64, c->rank + 1, c->comm, st));
        }
        LZ_NCCL(g_nccl.GroupEnd());
    }
    if (side) LZ_CUDA(cudaEventRecord(ctx->ev_halo, st));
    else lz_prof_end(ctx);
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// bootstrap helpers: host-visible all-gathers over NCCL (once per solve, never inside the iteration)
// ---------------------------------------------------------------------------------------------
static int gather_bytes64(lz_ctx *ctx, const void *mine64, void *all /* world * 64 */)
{
    lz_comm *c = ctx->comm;
    LZ_CUDA(cudaMemcpyAsync(c->stage, mine64, 64, cudaMemcpyHostToDevice, ctx->stream));
    LZ_NCCL(g_nccl.AllGather(c->stage, c->stage + 64, 64, NCCL_INT8, c->comm, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(all, c->stage + 64, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_comm_gather4(lz_ctx *ctx, const int64_t mine[4], int64_t *all)
{
    int64_t rec[8] = {mine[0], mine[1], mine[2], mine[3], 0, 0, 0, 0};
    int64_t tmp[8 * LZ_MAX_RANKS];
    LZ_TRY(gather_bytes64(ctx, rec, tmp));
    for (int r = 0; r < ctx->comm->world; ++r)
        for (int k = 0; k < 4; ++k) all[4 * r + k] = tmp[8 * r + k];
    return LZ_OK;
}

static void arena_release(lz_comm *c)
{
    for (int q = 0; q < c->world; ++q)
        if (q != c->rank && c->peer[q]) { cudaIpcCloseMemHandle(c->peer[q]); c->peer[q] = nullptr; }
    if (c->arena) cudaFree(c->arena);
    c->arena = nullptr; c->peer[c->rank] = nullptr;
    c->peer_ok = 0; c->user_bytes = 0; c->ar_cap = 0;
}

// Collective.  Makes sure every rank's arena offers `user_bytes` of user space and slots of `ar_cap` doubles,
// (re)allocating and re-mapping on ALL ranks when any of them needs more.  Returns the user region, or -- when
// peer mapping is unavailable (LZ_COMM=1, IPC refused) -- the context's private workspace: callers then get
// NCCL collectives from lz_comm_allreduce_sum / lz_comm_halo_exchange without any further change.
int lz_comm_arena(lz_ctx *ctx, size_t user_bytes, size_t ar_cap, void **user)
{
    lz_comm *c = ctx->comm;
    LZ_CHECK(c, LZ_ERR_COMM, "lz_comm_arena: no communicator");
    if (ar_cap < 4096) ar_cap = 4096;
    user_bytes = (user_bytes + 4095) & ~(size_t)4095;
    const bool want_peer = ctx->knobs.comm_mode != 1 && !c->peer_tried && c->world > 1 && c->world <= LZ_MAX_RANKS;
    if (want_peer) {
        // does any rank need a bigger arena?  (sizes are made identical everywhere: the layout must match)
        const int64_t mine[4] = {(int64_t)user_bytes, (int64_t)ar_cap, (int64_t)c->user_bytes, (int64_t)c->ar_cap};
        int64_t all[4 * LZ_MAX_RANKS];
        LZ_TRY(lz_comm_gather4(ctx, mine, all));
        int64_t need_user = 0, need_cap = 0;
        bool grow = !c->peer_ok;
        for (int r = 0; r < c->world; ++r) {
            need_user = all[4 * r] > need_user ? all[4 * r] : need_user;
            need_cap = all[4 * r + 1] > need_cap ? all[4 * r + 1] : need_cap;
        }
        for (int r = 0; r < c->world; ++r)
            if (all[4 * r + 2] < need_user || all[4 * r + 3] < need_cap) grow = true;
        if (grow) {
            arena_release(c);
            const size_t slots = sizeof(double) * 2 * (size_t)c->world * (size_t)need_cap;
            const size_t uoff = (LZ_ARENA_FLAG_BYTES + slots + 4095) & ~(size_t)4095;
            const size_t total = uoff + (size_t)need_user;
            int ok = 1;
            cudaIpcMemHandle_t h;
            memset(&h, 0, sizeof(h));
            if (cudaMalloc(&c->arena, total) != cudaSuccess) { cudaGetLastError(); c->arena = nullptr; ok = 0; }
            if (ok && cudaMemsetAsync(c->arena, 0, uoff, ctx->stream) != cudaSuccess) ok = 0;
            if (ok && cudaIpcGetMemHandle(&h, c->arena) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
            cudaIpcMemHandle_t hs[LZ_MAX_RANKS];
            LZ_TRY(gather_bytes64(ctx, &h, hs));
            c->peer[c->rank] = c->arena;
            for (int q = 0; ok && q < c->world; ++q) {
                if (q == c->rank) continue;
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, hs[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                c->peer[q] = (char *)ptr;
            }
            // every rank must have succeeded, otherwise all of them stay on NCCL
            const int64_t okrec[4] = {ok, 0, 0, 0};
            LZ_TRY(lz_comm_gather4(ctx, okrec, all));
            for (int r = 0; r < c->world; ++r) ok = ok && all[4 * r] != 0;
            if (!ok) {
                arena_release(c);
                c->peer_tried = 1;
                LZ_CHECK(ctx->knobs.comm_mode != 2, LZ_ERR_COMM, "lz_comm_arena: LZ_COMM=2 but peer memory (CUDA IPC) is unavailable");
            } else {
                c->ar_cap = (size_t)need_cap; c->user_off = uoff; c->user_bytes = (size_t)need_user;
                c->ar_seq = 0; c->halo_seq = 0;
                LzPeerDesc d;
                memset(&d, 0, sizeof(d));
                d.world = c->world; d.rank = c->rank; d.cap = (unsigned long long)need_cap; d.err = c->err_dev;
                auto flags_of = [&](int q) { return reinterpret_cast<unsigned long long *>(c->peer[q]); };
                auto slots_of = [&](int q) { return reinterpret_cast<double *>(c->peer[q] + LZ_ARENA_FLAG_BYTES); };
                for (int p = 0; p < 2; ++p) {
                    d.my_flag[p] = flags_of(c->rank) + p * LZ_MAX_RANKS;
                    d.my_slot[p] = slots_of(c->rank) + (size_t)p * c->world * need_cap;
                    for (int q = 0; q < c->world; ++q) {
                        d.flag_at[p][q] = flags_of(q) + p * LZ_MAX_RANKS + c->rank;
                        d.slot_at[p][q] = slots_of(q) + ((size_t)p * c->world + c->rank) * need_cap;
                    }
                }
                d.my_halo_flag = flags_of(c->rank) + 2 * LZ_MAX_RANKS;
                if (c->rank > 0) d.halo_flag_at[0] = flags_of(c->rank - 1) + 2 * LZ_MAX_RANKS + 1;
                if (c->rank < c->world - 1) d.halo_flag_at[1] = flags_of(c->rank + 1) + 2 * LZ_MAX_RANKS + 0;
                LZ_CUDA(cudaMemcpyAsync(c->desc_dev, &d, sizeof(d), cudaMemcpyHostToDevice, ctx->stream));
                LZ_CUDA(cudaStreamSynchronize(ctx->stream));
                c->peer_ok = 1;
            }
        }
    }
    if (c->peer_ok) {
        *user = c->arena + c->user_off;
        return LZ_OK;
    }
    return lz_ctx_workspace(ctx, user_bytes, user);
}

int lz_gen_lap3d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_local,
                      int64_t col_shift, int64_t n_cols, lz_matrix **out);
int lz_gen_lap2d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t row0, int64_t n_local, int64_t col_shift,
                      int64_t n_cols, lz_matrix **out);

// rows [0, lo_end) reference the lower halo, rows [hi_begin, n) the upper one: find the chunks of both
// schedules that lie strictly between (host binary search over a copy of the chunk-row tables)
static int set_split(lz_ctx *ctx, lz_matrix *A, int64_t lo_end, int64_t hi_begin)
{
    A->has_split = 0;
    if (A->format != LZ_FMT_CSR || !A->chunk_row || !A->mm_chunk_row || A->vrowptr) return LZ_OK;
    auto find = [&](const int32_t *dev, int n_chunks, int *c_lo, int *c_hi) -> int {
        int32_t *h = new int32_t[n_chunks + 1];
        cudaError_t e = cudaMemcpy(h, dev, sizeof(int32_t) * (n_chunks + 1), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { delete[] h; lz_set_error("set_split: %s", cudaGetErrorString(e)); return LZ_ERR_CUDA; }
        int lo = 0;
        while (lo < n_chunks && h[lo] < lo_end) ++lo;                  // first chunk starting at or after lo_end
        int hi = n_chunks;
        while (hi > lo && h[hi] > hi_begin) --hi;                      // chunks [lo, hi) end at or before hi_begin
        delete[] h;
        *c_lo = lo; *c_hi = hi;
        return LZ_OK;
    };
    LZ_TRY(find(A->chunk_row, A->n_chunks, &A->bnd_lo, &A->bnd_hi));
    LZ_TRY(find(A->mm_chunk_row, A->mm_n_chunks, &A->mm_bnd_lo, &A->mm_bnd_hi));
    A->has_split = 1;
    return LZ_OK;
}

extern "C" {

int lz_comm_unique_id(void *id128_host)
{
    LZ_CHECK(id128_host, LZ_ERR_INVALID, "lz_comm_unique_id: NULL argument");
    LZ_TRY(load_nccl());
    ncclUniqueId id;
    LZ_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128_host, &id, sizeof(id));
    return LZ_OK;
}

int lz_comm_init(lz_ctx *ctx, int world_size, int rank, const void *id128_host)
{
    LZ_CHECK(ctx && id128_host && world_size >= 1 && rank >= 0 && rank < world_size, LZ_ERR_INVALID, "lz_comm_init: bad arguments");
    LZ_CHECK(!ctx->comm, LZ_ERR_INVALID, "lz_comm_init: context already has a communicator");
    LZ_TRY(load_nccl());
    LZ_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    lz_comm *c = new lz_comm();
    memset(c, 0, sizeof(*c));
    c->world = world_size;
    c->rank = rank;
    int r = g_nccl.CommInitRank(&c->comm, world_size, id, rank);
    if (r != NCCL_SUCCESS) {
        lz_set_error("lz_comm_init: ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
        delete c;
        return LZ_ERR_COMM;
    }
    ctx->comm = c;
    LZ_CUDA(cudaMalloc(&c->stage, 64 * (size_t)(1 + LZ_MAX_RANKS)));
    LZ_CUDA(cudaMalloc(&c->desc_dev, sizeof(LzPeerDesc)));
    LZ_CUDA(cudaMalloc(&c->err_dev, sizeof(int)));
    LZ_CUDA(cudaMalloc(&c->halo_ticket, sizeof(unsigned int)));
    LZ_CUDA(cudaMemset(c->err_dev, 0, sizeof(int)));
    LZ_CUDA(cudaMemset(c->halo_ticket, 0, sizeof(unsigned int)));
    return LZ_OK;
}

// 0 = fine; 1 = a peer-memory wait gave up (a rank was lost): every result since is invalid
int lz_comm_status(lz_ctx *ctx, int *peer_mode, int *timed_out)
{
    LZ_CHECK(ctx && ctx->comm, LZ_ERR_COMM, "lz_comm_status: no communicator");
    int e = 0;
    LZ_CUDA(cudaMemcpyAsync(&e, ctx->comm->err_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (peer_mode) *peer_mode = ctx->comm->peer_ok;
    if (timed_out) *timed_out = e;
    return LZ_OK;
}

int lz_comm_destroy(lz_ctx *ctx)
{
    if (!ctx || !ctx->comm) return LZ_OK;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    arena_release(ctx->comm);
    cudaFree(ctx->comm->stage); cudaFree(ctx->comm->desc_dev); cudaFree(ctx->comm->err_dev); cudaFree(ctx->comm->halo_ticket);
    if (g_nccl.handle) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
    return LZ_OK;
}

// rows are dealt in whole granules (grid planes / lines for the stencil operators): rank r owns
// granules [G*r/world, G*(r+1)/world)
int lz_partition_rows(int64_t n_rows, int64_t granule, int world_size, int rank, int64_t *begin, int64_t *end)
{
    LZ_CHECK(begin && end && n_rows > 0 && granule > 0 && world_size >= 1 && rank >= 0 && rank < world_size, LZ_ERR_INVALID,
             "lz_partition_rows: bad arguments");
    LZ_CHECK(n_rows % granule == 0, LZ_ERR_INVALID, "lz_partition_rows: %lld rows are not a multiple of the granule %lld",
             (long long)n_rows, (long long)granule);
    const int64_t G = n_rows / granule;
    LZ_CHECK(G >= world_size, LZ_ERR_INVALID, "lz_partition_rows: fewer granules (%lld) than ranks (%d)", (long long)G, world_size);
    *begin = (G * rank / world_size) * granule;
    *end = (G * (rank + 1) / world_size) * granule;
    return LZ_OK;
}

int lz_gen_laplacian3d_shard(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int world_size, int rank, lz_matrix **out)
{
    LZ_CHECK(ctx && out && nx >= 2 && ny >= 2 && nz >= 2, LZ_ERR_INVALID, "lz_gen_laplacian3d_shard: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const int64_t sxy = nx * ny;
    int64_t r0, r1;
    LZ_TRY(lz_partition_rows(sxy * nz, sxy, world_size, rank, &r0, &r1));
    const int64_t hlo = rank > 0 ? sxy : 0, hhi = rank < world_size - 1 ? sxy : 0;
    LZ_TRY(lz_gen_lap3d_rows(ctx, nx, ny, nz, r0, r1 - r0, r0 - hlo, hlo + (r1 - r0) + hhi, out));
    (*out)->halo_lo = hlo; (*out)->halo_hi = hhi;
    (*out)->global_rows = sxy * nz; (*out)->row_begin = r0;
    return set_split(ctx, *out, hlo, (r1 - r0) - hhi);     // first / last plane of the slab touch the halos
}

// A row slab of ANY square operator from host CSR arrays: rows [row_begin, row_begin + n_local) of a global operator
// with global_rows rows, column ids already shifted into the local index space [lower halo | local | upper halo]
// (halo_lo / halo_hi entries owned by the neighbouring ranks).  Rows [0, bnd_lo_rows) are the ones that reference the
// lower halo, rows [bnd_hi_rows, n_local) the upper one (pass 0 / n_local when unknown: no overlap, still correct).
int lz_csr_create_shard_host(lz_ctx *ctx, int64_t n_local, int64_t nnz, const int32_t *rowptr_host, const int32_t *colidx_host,
                             const double *vals_host, int64_t halo_lo, int64_t halo_hi, int64_t global_rows, int64_t row_begin,
                             int64_t bnd_lo_rows, int64_t bnd_hi_rows, lz_matrix **out)
{
    LZ_CHECK(ctx && out && halo_lo >= 0 && halo_hi >= 0 && row_begin >= 0 && row_begin + n_local <= global_rows, LZ_ERR_INVALID,
             "lz_csr_create_shard_host: bad arguments");
    LZ_CHECK(bnd_lo_rows >= 0 && bnd_hi_rows <= n_local, LZ_ERR_INVALID,
             "lz_csr_create_shard_host: bad boundary row ranges");
    LZ_TRY(lz_csr_create_host(ctx, n_local, halo_lo + n_local + halo_hi, nnz, rowptr_host, colidx_host, vals_host, out));
    (*out)->halo_lo = halo_lo; (*out)->halo_hi = halo_hi;
    (*out)->global_rows = global_rows; (*out)->row_begin = row_begin;
    if (bnd_lo_rows <= bnd_hi_rows) return set_split(ctx, *out, bnd_lo_rows, bnd_hi_rows);
    return LZ_OK;
}

int lz_gen_laplacian2d_shard(lz_ctx *ctx, int64_t nx, int64_t ny, int world_size, int rank, lz_matrix **out)
{
    // the 5-point operator on nx x ny is the 7-point generator's z-slab structure with planes = lines;
    // it is generated by its own kernel path: rows [r0, r1) of lz_gen_laplacian2d with shifted columns
    LZ_CHECK(ctx && out && nx >= 2 && ny >= 2, LZ_ERR_INVALID, "lz_gen_laplacian2d_shard: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    int64_t r0, r1;
    LZ_TRY(lz_partition_rows(nx * ny, nx, world_size, rank, &r0, &r1));
    const int64_t hlo = rank > 0 ? nx : 0, hhi = rank < world_size - 1 ? nx : 0;
    LZ_TRY(lz_gen_lap2d_rows(ctx, nx, ny, r0, r1 - r0, r0 - hlo, hlo + (r1 - r0) + hhi, out));
    (*out)->halo_lo = hlo; (*out)->halo_hi = hhi;
    (*out)->global_rows = nx * ny; (*out)->row_begin = r0;
    return set_split(ctx, *out, hlo, (r1 - r0) - hhi);     // first / last grid line of the slab touch the halos
}

}  // extern "C"
