// lz_ctx.cu -- context, error state, raw memory entry points of the C-ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "lz_common.cuh"

static thread_local char g_err[512] = "";

void lz_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int lz_version(void) { return 100; }
const char *lz_last_error(void) { return g_err; }

int lz_ctx_create(int device, void *stream, lz_ctx **out)
{
    LZ_CHECK(out != nullptr, LZ_ERR_INVALID, "lz_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        lz_set_error("lz_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return LZ_ERR_CUDA;
    }
    LZ_CHECK(device >= 0 && device < count, LZ_ERR_INVALID, "lz_ctx_create: device %d out of range", device);
    LZ_CUDA(cudaSetDevice(device));
    lz_ctx *c = new lz_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->stream = (cudaStream_t)stream;
    cudaDeviceProp prop;
    LZ_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        lz_set_error("lz_ctx_create: device %d is sm_%d%d; this build targets sm_100a only", device,
                     prop.major, prop.minor);
        delete c;
        return LZ_ERR_UNSUPPORTED;
    }
    LZ_CUDA(cudaMalloc(&c->partials, sizeof(double) * LZ_PARTIALS_CAP));
    LZ_CUDA(cudaMalloc(&c->tickets, sizeof(unsigned int) * LZ_TICKETS));
    LZ_CUDA(cudaMalloc(&c->scalars, sizeof(double) * LZ_SCALARS));
    LZ_CUDA(cudaMalloc(&c->flags, sizeof(int) * LZ_FLAGS));
    LZ_CUDA(cudaMemset(c->tickets, 0, sizeof(unsigned int) * LZ_TICKETS));
    LZ_CUDA(cudaMemset(c->scalars, 0, sizeof(double) * LZ_SCALARS));
    LZ_CUDA(cudaMemset(c->flags, 0, sizeof(int) * LZ_FLAGS));
    if (const char *e = getenv("LZ_SPMV_VARIANT")) c->spmv_variant = atoi(e);
    auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    auto env_set = [](const char *name) { return getenv(name) ? 1 : 0; };
    LzKnobs &k = c->knobs;
    k.spmv_hint = env_set("LZ_SPMV_HINT");
    k.spmm_hint = env_int("LZ_SPMM_HINT", -1);
    k.spmm_run = env_int("LZ_SPMM_RUN", 1);
    k.no_split = env_set("LZ_NO_SPLIT");
    k.cgs_shape_order = env_int("LZ_CGS_SHAPE_ORDER", 1);
    if (k.cgs_shape_order < 0 || k.cgs_shape_order > 2) k.cgs_shape_order = 1;
    k.cgs_upd_mult = env_int("LZ_CGS_UPD_MULT", 3);
    if (k.cgs_upd_mult < 1) k.cgs_upd_mult = 1;
    k.cgs_one_cta = env_set("LZ_CGS_ONE_CTA");
    k.no_cgs_fuse = env_set("LZ_NO_CGS_FUSE");
    k.cgs_no_slices = env_set("LZ_CGS_NO_SLICES");
    k.comm_mode = env_int("LZ_COMM", 0);
    k.no_overlap = env_set("LZ_NO_OVERLAP");
    k.no_fold = env_set("LZ_NO_FOLD");
    k.spmm_kernel = env_int("LZ_SPMM_KERNEL", 0);
    k.block_cgs_fuse = env_int("LZ_BLOCK_CGS_FUSE", 1);
    k.rmat_reorder = env_int("LZ_REORDER", 1);
    k.panel_pad = env_int("LZ_PANEL_PAD", 0);
    k.spmm_slice = env_int("LZ_SPMM_SLICE", 0);
    k.spmm_shape = env_int("LZ_SPMM_SHAPE", 0);
    k.cgs_fuse_min_k = env_int("LZ_CGS_FUSE_MIN_K", 64);
    k.cgs_rpt = env_int("LZ_CGS_RPT", 0);
    k.split_l = env_int("LZ_SPLIT_L", LZ_SPLIT_L);
    if (k.split_l < 8) k.split_l = 8;
    if (k.split_l > 256) k.split_l = 256;
    k.no_xs = env_set("LZ_NO_XS");
    k.xs_stages = env_int("LZ_XS_STAGES", 0);
    k.no_direct_rows = env_set("LZ_NO_DIRECT_ROWS");
    k.xs_force = env_set("LZ_XS_FORCE");
    k.xs_no_tiles = env_set("LZ_XS_NO_TILES");
    k.xs_tile = env_int("LZ_XS_TILE", LZ_XS_TILE);
    if (k.xs_tile < 128) k.xs_tile = 128;
    if (k.xs_tile > LZ_XS_TILE) k.xs_tile = LZ_XS_TILE;
    k.split_l_mm = env_int("LZ_SPLIT_L_MM", 32);
    if (k.split_l_mm < 8) k.split_l_mm = 8;
    if (k.split_l_mm > 256) k.split_l_mm = 256;
    k.no_transpose = env_set("LZ_TRANSPOSE") ? 0 : 1;     // measured slower (profiles/r02_spmv.md): opt-in only
    k.no_spmm_fuse = env_set("LZ_NO_SPMM_FUSE");
    k.no_spmm_gram = env_set("LZ_NO_SPMM_GRAM");
    *out = c;
    return LZ_OK;
}

}  // extern "C"

// one-time (per context, i.e. per device) opt-in to more than 48 KB of dynamic shared memory
int lz_func_smem_optin(lz_ctx *ctx, const void *func, int bytes, bool carveout_max)
{
    for (int i = 0; i < ctx->n_attr; ++i)
        if (ctx->attr_funcs[i] == func) return LZ_OK;
    LZ_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (carveout_max) LZ_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (ctx->n_attr < LZ_ATTR_CAP) ctx->attr_funcs[ctx->n_attr++] = func;    // (a full table only costs repeated calls)
    return LZ_OK;
}

extern "C" {

int lz_ctx_destroy(lz_ctx *ctx)
{
    if (!ctx) return LZ_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) lz_comm_destroy(ctx);
    delete ctx->vrun;
    ctx->vrun = nullptr;
    // orphan the operators still alive: they keep their device arrays and can be destroyed later
    for (lz_matrix *A = ctx->matrices; A;) {
        lz_matrix *nx = A->next;
        A->ctx = nullptr; A->next = A->prev = nullptr;
        A = nx;
    }
    ctx->matrices = nullptr;
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    cudaFree(ctx->partials);
    cudaFree(ctx->tickets);
    cudaFree(ctx->scalars);
    cudaFree(ctx->flags);
    cudaFree(ctx->work);
    cudaFree(ctx->scratch);
    cudaFree(ctx->basis);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * LZ_PROF_CAP; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        delete[] ctx->prof_ev; delete[] ctx->prof_cls; delete[] ctx->prof_bytes;
    }
    delete ctx;
    return LZ_OK;
}

int lz_ctx_sync(lz_ctx *ctx)
{
    LZ_CHECK(ctx, LZ_ERR_INVALID, "lz_ctx_sync: ctx is NULL");
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_ctx_device(const lz_ctx *ctx) { return ctx ? ctx->device : -1; }
int64_t lz_ctx_launch_count(const lz_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lz_malloc(lz_ctx *ctx, size_t bytes, void **dptr)
{
    LZ_CHECK(ctx && dptr, LZ_ERR_INVALID, "lz_malloc: NULL argument");
    LZ_CUDA(cudaSetDevice(ctx->device));
    *dptr = nullptr;
    if (bytes == 0) return LZ_OK;
    LZ_CUDA(cudaMalloc(dptr, bytes));
    return LZ_OK;
}

int lz_free(lz_ctx *ctx, void *dptr)
{
    LZ_CHECK(ctx, LZ_ERR_INVALID, "lz_free: ctx is NULL");
    LZ_CUDA(cudaFree(dptr));
    return LZ_OK;
}

int lz_memcpy(lz_ctx *ctx, void *dst, const void *src, size_t bytes, int kind)
{
    LZ_CHECK(ctx, LZ_ERR_INVALID, "lz_memcpy: ctx is NULL");
    cudaMemcpyKind k = kind == LZ_H2D ? cudaMemcpyHostToDevice
                       : kind == LZ_D2H ? cudaMemcpyDeviceToHost
                                        : cudaMemcpyDeviceToDevice;
    LZ_CHECK(kind >= LZ_H2D && kind <= LZ_D2D, LZ_ERR_INVALID, "lz_memcpy: bad kind %d", kind);
    if (bytes == 0) return LZ_OK;
    LZ_CUDA(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_memset(lz_ctx *ctx, void *dptr, int value, size_t bytes)
{
    LZ_CHECK(ctx, LZ_ERR_INVALID, "lz_memset: ctx is NULL");
    LZ_CUDA(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return LZ_OK;
}

}  // extern "C"

// fold the recorded pairs into the per-class sums (waits for the last recorded event)
static void prof_drain(lz_ctx *ctx)
{
    if (ctx->prof_used == 0) return;
    cudaEventSynchronize(ctx->prof_ev[2 * (ctx->prof_used - 1) + 1]);
    for (int i = 0; i < ctx->prof_used; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) != cudaSuccess) continue;
        const int c = ctx->prof_cls[i];
        ctx->prof_acc_launches[c]++; ctx->prof_acc_ms[c] += t; ctx->prof_acc_bytes[c] += ctx->prof_bytes[i];
    }
    ctx->prof_used = 0;
}

void lz_prof_begin(lz_ctx *ctx, int cls, double bytes)
{
    if (!ctx->prof_on) return;
    if (ctx->prof_used >= LZ_PROF_CAP) prof_drain(ctx);       // ring full: one host wait per LZ_PROF_CAP launches
    ctx->prof_cls[ctx->prof_used] = cls;
    ctx->prof_bytes[ctx->prof_used] = bytes;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used], ctx->stream);
}

void lz_prof_end(lz_ctx *ctx)
{
    if (!ctx->prof_on || ctx->prof_used >= LZ_PROF_CAP) return;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used + 1], ctx->stream);
    ctx->prof_used++;
}

extern "C" {

// enable (1) / disable (0) per-kernel-class event timing; enabling resets the log
int lz_ctx_profile(lz_ctx *ctx, int enable)
{
    LZ_CHECK(ctx, LZ_ERR_INVALID, "lz_ctx_profile: ctx is NULL");
    if (enable && !ctx->prof_ev) {
        ctx->prof_ev = new cudaEvent_t[2 * LZ_PROF_CAP];
        ctx->prof_cls = new int[LZ_PROF_CAP];
        ctx->prof_bytes = new double[LZ_PROF_CAP];
        for (int i = 0; i < 2 * LZ_PROF_CAP; ++i) LZ_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    }
    ctx->prof_on = enable;
    if (enable) {
        ctx->prof_used = 0;
        for (int c = 0; c < LZ_K_CLASSES; ++c) { ctx->prof_acc_launches[c] = 0; ctx->prof_acc_ms[c] = 0.0; ctx->prof_acc_bytes[c] = 0.0; }
    }
    return LZ_OK;
}

// sums per class since the last enable: launches, milliseconds, algorithmic bytes (arrays of LZ_K_CLASSES = LZ_PROFILE_CLASSES)
int lz_ctx_profile_read(lz_ctx *ctx, int64_t *launches, double *ms, double *bytes)
{
    LZ_CHECK(ctx && launches && ms && bytes, LZ_ERR_INVALID, "lz_ctx_profile_read: NULL argument");
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    prof_drain(ctx);
    for (int c = 0; c < LZ_K_CLASSES; ++c) {
        launches[c] = ctx->prof_acc_launches[c]; ms[c] = ctx->prof_acc_ms[c]; bytes[c] = ctx->prof_acc_bytes[c];
        ctx->prof_acc_launches[c] = 0; ctx->prof_acc_ms[c] = 0.0; ctx->prof_acc_bytes[c] = 0.0;
    }
    return LZ_OK;
}

}  // extern "C"

int lz_ctx_workspace(lz_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->work_bytes) {
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->work) LZ_CUDA(cudaFree(ctx->work));
        ctx->work = nullptr;
        ctx->work_bytes = 0;
        size_t want = bytes + (bytes >> 3);
        LZ_CUDA(cudaMalloc(&ctx->work, want));
        ctx->work_bytes = want;
    }
    *out = ctx->work;
    return LZ_OK;
}

int lz_ctx_scratch(lz_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->scratch_bytes) {
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) LZ_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        LZ_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return LZ_OK;
}

static int basis_alloc(lz_ctx *ctx, size_t bytes)
{
    bytes += 4096;
    if (bytes > ctx->basis_bytes) {
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->basis) LZ_CUDA(cudaFree(ctx->basis));
        ctx->basis = nullptr;
        ctx->basis_bytes = 0;
        LZ_CUDA(cudaMalloc(&ctx->basis, bytes));
        ctx->basis_bytes = bytes;
    }
    return LZ_OK;
}

// Krylov basis of the single-vector path, row-tiled: 32-row tiles, the `cols` columns of a tile back to
// back, so element (i, k) sits at basis[(i >> 5) * ts + k * cs + (i & 31)] with cs = 32, ts = 32 * cols.
// A tile's first K columns are one contiguous block (one bulk copy in the fused CGS kernel) and column
// accesses of the streaming kernels stay 256-byte segments.
int lz_ctx_basis(lz_ctx *ctx, int64_t rows, int cols, double **out)
{
    const int64_t tiles = (rows + 31) / 32;
    LZ_TRY(basis_alloc(ctx, sizeof(double) * (size_t)tiles * 32 * (size_t)cols));
    ctx->basis_cs = 32;
    ctx->basis_ts = 32 * (int64_t)cols;
    ctx->basis_rows = rows;
    ctx->basis_cols = cols;
    *out = ctx->basis;
    return LZ_OK;
}

int lz_ctx_basis_blocks(lz_ctx *ctx, int64_t pan, int blocks, double **out)
{
    LZ_TRY(basis_alloc(ctx, sizeof(double) * (size_t)pan * (size_t)blocks));
    ctx->basis_cs = 0; ctx->basis_ts = 0; ctx->basis_rows = 0; ctx->basis_cols = 0;
    *out = ctx->basis;
    return LZ_OK;
}
