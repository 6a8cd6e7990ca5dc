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
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *handle;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
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

struct lz_comm {
    ncclComm_t comm;
    int world, rank;
};

int lz_comm_world(const lz_ctx *ctx) { return ctx->comm ? ctx->comm->world : 1; }
int lz_comm_rank(const lz_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }

int lz_comm_allreduce_sum(lz_ctx *ctx, double *buf, size_t count)
{
    lz_prof_begin(ctx, LZ_K_COMM, 8.0 * (double)count);
    LZ_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, ctx->comm->comm, ctx->stream));
    lz_prof_end(ctx);
    return LZ_OK;
}

int lz_comm_halo_exchange(lz_ctx *ctx, double *u, int64_t n, int64_t hlo, int64_t hhi)
{
    const lz_comm *c = ctx->comm;
    lz_prof_begin(ctx, LZ_K_COMM, 8.0 * (double)(hlo + hhi) * 2.0);
    LZ_NCCL(g_nccl.GroupStart());
    if (c->rank > 0 && hlo > 0) {
        // my first hlo rows are the lower neighbour's upper halo; its last rows are my lower halo
        LZ_NCCL(g_nccl.Send(u, (size_t)hlo, NCCL_FLOAT64, c->rank - 1, c->comm, ctx->stream));
        LZ_NCCL(g_nccl.Recv(u - hlo, (size_t)hlo, NCCL_FLOAT64, c->rank - 1, c->comm, ctx->stream));
    }
    if (c->rank < c->world - 1 && hhi > 0) {
        LZ_NCCL(g_nccl.Send(u + n - hhi, (size_t)hhi, NCCL_FLOAT64, c->rank + 1, c->comm, ctx->stream));
        LZ_NCCL(g_nccl.Recv(u + n, (size_t)hhi, NCCL_FLOAT64, c->rank + 1, c->comm, ctx->stream));
    }
    LZ_NCCL(g_nccl.GroupEnd());
    lz_prof_end(ctx);
    return LZ_OK;
}

int lz_gen_lap3d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_local,
                      int64_t col_shift, int64_t n_cols, lz_matrix **out);
int lz_gen_lap2d_rows(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t row0, int64_t n_local, int64_t col_shift,
                      int64_t n_cols, lz_matrix **out);

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
    c->world = world_size;
    c->rank = rank;
    int r = g_nccl.CommInitRank(&c->comm, world_size, id, rank);
    if (r != NCCL_SUCCESS) {
        lz_set_error("lz_comm_init: ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
        delete c;
        return LZ_ERR_COMM;
    }
    ctx->comm = c;
    return LZ_OK;
}

int lz_comm_destroy(lz_ctx *ctx)
{
    if (!ctx || !ctx->comm) return LZ_OK;
    cudaStreamSynchronize(ctx->stream);
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
    return LZ_OK;
}

}  // extern "C"
