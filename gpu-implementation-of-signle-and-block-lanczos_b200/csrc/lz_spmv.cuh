// lz_spmv.cuh -- fp64 SpMV kernels for sm_100a with the Lanczos "pass A" fused into the epilogue.
//
// Replaces ell::SpMV (kernels/spmv_spmm.hpp:105-135), lm::spmv_basic (kernels/ell_kernels.hpp:13-34)
// and, in fused mode, the spmv + Vector::add + Vector::dot sequence of
// methods/vector_lanczos.hpp:51-57 (one HBM pass instead of three).
//
// Default CSR kernel (k_csr_spmv_ws, 16-byte aligned vals/colidx): persistent, warp-specialised CTAs walk a
// schedule of row-aligned chunks -- a producer warp streams each chunk's vals/colidx slice into a
// shared-memory ring with bulk async copies (cp.async.bulk + mbarrier), gather warps turn it into products
// in place (x through LDG), row warps add the products of each row left to right and run the epilogue.
// Fallback CSR kernel (k_csr_spmv, unaligned arrays): one CTA per row-aligned chunk.
//   phase 1: the CTA walks its contiguous slice of vals/colidx with 128-bit/64-bit coalesced loads,
//            gathers x and parks val*x products in shared memory (conflict-free 16-byte stores);
//   phase 2: one thread per row adds its products left to right (the reference Host order,
//            objects/ell_matrix.hpp:246-251) and runs the epilogue.
//   chunks that contain a long row (more products than shared memory holds) fall back to
//   warp-per-row / CTA-per-row walks of global memory.
// ELL4 kernel: one thread per row, one 256-bit load of the four values and one 128-bit load of the
// four column indices (the reference's row-interleaved width-4 layout).
#pragma once
#include "lz_common.cuh"

// SCALED: gather scaling of LANCZOS, plain store.  EULER: y = x + dt * (A x), the time step of the fdtd
// validator (methods/fdtd.hpp:17-25: spmv followed by Vector::add) in one pass.
enum { LZ_EPI_PLAIN = 0, LZ_EPI_LANCZOS = 1, LZ_EPI_SCALED = 2, LZ_EPI_EULER = 3 };

// Arguments of the fused Lanczos epilogue (pass A of step j):
//   q_j[i]  = x_own[i] * invb[j]                 (lazy normalisation: the same product the reference
//                                                 materialises with mult_scalar, vector_lanczos.hpp:48)
//   w[i]    = sum_k A[i,k] q_j[col_k]  -  beta[j] * (u_prev[i] * invb[j-1])       (:51,:54)
//   alpha_j = sum_i w[i] q_j[i]                                                   (:57)
struct LzPassA {
    const double *x_own;    // x restricted to the locally owned rows (x + halo_lo for shards)
    const double *u_prev;   // unnormalised q_{j-1}
    const double *invb;     // invb[j] = 1 / beta_j
    const double *beta;     // beta[j]
    double *alpha_out;      // &alpha[j] (written by the last CTA) or NULL when the caller all-reduces
    double *alpha_partial;  // where the last CTA leaves the local sum (always)
    double *vcol;           // basis column j to fill with q_j (element i at vcol[(i >> 5) * vts + (i & 31)]), or NULL
    int64_t vts;
    double *qout;           // &q[j]: receives q_j[lc] (copy_vector_element, copy_functions.hpp:116-133)
    int64_t lc;
    int j;
    int first;              // step 0: no q_{-1} term
    double *partials;
    unsigned int *ticket;
    double dt;              // EULER mode: step length
    // sharded runs (peer-memory mode): the last CTA all-reduces the alpha partial over the ranks before publishing it
    const LzPeerDesc *pd;
    unsigned long long seq;
    int alpha_accum;        // add to *alpha_partial instead of overwriting it (second launch of an interior/boundary pair)
    int alpha_hold;         // first launch of such a pair: only leave the partial, no all-reduce, no alpha_out
};

// what the last CTA of a fused SpMV does with the grid total of w.q
__device__ __forceinline__ void lz_alpha_publish(const LzPassA &a, double total)
{
    if (a.alpha_accum) total += *a.alpha_partial;
    if (a.alpha_hold) { *a.alpha_partial = total; return; }
    if (a.pd) lz_peer_sum_thread<1>(a.pd, a.seq, &total);
    *a.alpha_partial = total;
    if (a.alpha_out) *a.alpha_out = total;
}

// contiguous sub-ranges of the chunk schedule a launch works on: virtual chunk v < n0 is chunk c0 + v, the rest
// c1 + (v - n0).  Whole operator: {0, n_chunks, 0, n_chunks}.  Sharded operators launch the interior chunks
// (no halo columns) while the halo exchange is in flight and the two boundary ranges afterwards.
struct LzChunkRange { int c0, n0, c1, total; };

template <int MODE>
struct LzRowEpi {
    // the fields of LzPassA the epilogue needs, held by value: a reference to the kernel parameter
    // would force the whole struct through local memory
    double sx, sprev, beta;
    const double *x_own, *u_prev;
    double *vcol, *qout;
    int64_t lc, vts;
    bool first;
    __device__ __forceinline__ LzRowEpi(const LzPassA &args)
    {
        sx = 1.0; sprev = 0.0; beta = 0.0;
        x_own = args.x_own; u_prev = args.u_prev; vcol = args.vcol; vts = args.vts; qout = args.qout; lc = args.lc; first = args.first != 0;
        if (MODE == LZ_EPI_SCALED) sx = args.invb[args.j];
        if (MODE == LZ_EPI_EULER) beta = args.dt;
        if (MODE == LZ_EPI_LANCZOS) {
            sx = args.invb[args.j];
            if (!first) { sprev = args.invb[args.j - 1]; beta = args.beta[args.j]; }
        }
    }
    // scale applied to every gathered x value
    __device__ __forceinline__ double xs(double xv) const { return (MODE == LZ_EPI_LANCZOS || MODE == LZ_EPI_SCALED) ? __dmul_rn(xv, sx) : xv; }
    // epilogue operands of row i (fetched early by the pipelined kernels)
    __device__ __forceinline__ void load(int64_t i, double &xo, double &up) const
    {
        xo = 0.0; up = 0.0;
        if (MODE == LZ_EPI_LANCZOS) {
            xo = x_own[i];
            if (!first) up = u_prev[i];
        }
        if (MODE == LZ_EPI_EULER) xo = x_own[i];
    }
    // finish row i whose raw sum is t; returns the row's contribution to alpha
    __device__ __forceinline__ double finish(int64_t i, double t, double *__restrict__ y, double xo, double up) const
    {
        if (MODE == LZ_EPI_EULER) { y[i] = fma(beta, t, xo); return 0.0; }
        if (MODE != LZ_EPI_LANCZOS) { y[i] = t; return 0.0; }
        const double qi = __dmul_rn(xo, sx);
        double w = t;
        if (!first) w = __dadd_rn(t, __dmul_rn(-beta, __dmul_rn(up, sprev)));
        y[i] = w;
        if (vcol) vcol[(i >> 5) * vts + (i & 31)] = qi;
        if (i == lc && qout) *qout = qi;
        return __dmul_rn(w, qi);
    }
    __device__ __forceinline__ double finish(int64_t i, double t, double *__restrict__ y) const
    {
        double xo, up;
        load(i, xo, up);
        return finish(i, t, y, xo, up);
    }
};

template <int MODE>
__device__ __forceinline__ void lz_spmv_finalize(double acc, const LzPassA &a, double *red)
{
    if (MODE != LZ_EPI_LANCZOS) return;
    acc = lz_block_sum<LZ_SPMV_THREADS>(acc, red);
    double total;
    if (lz_grid_sum<LZ_SPMV_THREADS, 1>(&acc, a.partials, a.ticket, red, &total)) {
        if (threadIdx.x == 0) lz_alpha_publish(a, total);
    }
}

template <int MODE>
__global__ void __launch_bounds__(LZ_SPMV_THREADS)
k_csr_spmv(const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr,
           const int32_t *__restrict__ colidx, const double *__restrict__ vals,
           const double *__restrict__ x, double *__restrict__ y, const LzPassA args)
{
    __shared__ __align__(16) double prod[LZ_SPMV_CAP + 2];
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int r0 = chunk_row[blockIdx.x], r1 = chunk_row[blockIdx.x + 1];
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;
    if (r1 > r0) {
        const int p0 = rowptr[r0], p1 = rowptr[r1];
        const int a0 = p0 & ~1;                       // products live at prod[p - a0]
        if (p1 - a0 <= LZ_SPMV_CAP) {
            // ---- phase 1: coalesced product generation --------------------------------------
            const int pe = p1 & ~1;                   // pairs cover [a0, pe)
            for (int p = a0 + 2 * tid; p < pe; p += 2 * LZ_SPMV_THREADS) {
                const double2 v = *reinterpret_cast<const double2 *>(vals + p);
                const int2 c = *reinterpret_cast<const int2 *>(colidx + p);
                double2 pr;
                // the first pair may start one entry before p0 (belongs to the previous chunk):
                // its column is valid memory, its product is never read
                pr.x = __dmul_rn(v.x, epi.xs(__ldg(x + c.x)));
                pr.y = __dmul_rn(v.y, epi.xs(__ldg(x + c.y)));
                *reinterpret_cast<double2 *>(prod + (p - a0)) = pr;
            }
            if (tid == 0 && pe < p1) prod[pe - a0] = __dmul_rn(vals[pe], epi.xs(__ldg(x + colidx[pe])));
            __syncthreads();
            // ---- phase 2: one thread per row, left-to-right sum -------------------------------
            for (int r = r0 + tid; r < r1; r += LZ_SPMV_THREADS) {
                const int s = rowptr[r] - a0, e = rowptr[r + 1] - a0;
                double t = 0.0;
                for (int k = s; k < e; ++k) t = __dadd_rn(t, prod[k]);
                acc += epi.finish(r, t, y);
            }
        } else {
            // ---- long-row fallback: warp per row, then CTA per very long row -------------------
            const int lane = tid & 31, warp = tid >> 5;
            for (int r = r0 + warp; r < r1; r += LZ_SPMV_THREADS / 32) {
                const int s = rowptr[r], e = rowptr[r + 1];
                if (e - s > 4096) continue;
                double t = 0.0;
                for (int k = s + lane; k < e; k += 32) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                t = lz_warp_sum(t);
                if (lane == 0) acc += epi.finish(r, t, y);
            }
            for (int r = r0; r < r1; ++r) {
                const int s = rowptr[r], e = rowptr[r + 1];
                if (e - s <= 4096) continue;
                double t = 0.0;
                for (int k = s + tid; k < e; k += LZ_SPMV_THREADS) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                t = lz_block_sum<LZ_SPMV_THREADS>(t, red);
                if (tid == 0) acc += epi.finish(r, t, y);
            }
        }
    }
    lz_spmv_finalize<MODE>(acc, args, red);
}

// ---------------------------------------------------------------------------------------------
// TMA-pipelined persistent CSR kernel (the default path).
//
// grid = a multiple of the SM count; CTA c walks chunks c, c + grid, ...  The contiguous vals /
// colidx slice of a chunk is brought into a shared-memory ring by two bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) issued by one elected thread STAGES-1
// chunks ahead, so HBM streaming never waits on the compute threads.  Consumers turn the staged
// values into products in place (x gathered with LDG, mostly L1/L2 hits), then one thread per row
// adds its products left to right and runs the epilogue.  Per-row operands of the epilogue
// (rowptr pair, u_prev, x_own) are fetched into registers before the wait on the ring.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lz_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lz_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lz_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lz_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lz_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LZ_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LZ_DONE_%=;\n\t"
        "bra LZ_WAIT_%=;\n\t"
        "LZ_DONE_%=:\n\t}"
        ::"r"(lz_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void lz_bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lz_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(lz_smem_u32(bar)) : "memory");
}

// same copy with an L2 evict-first policy: the matrix streams are read once per product and must not
// push the gathered operand (x / the X panel) out of L2
__device__ __forceinline__ uint64_t lz_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void lz_bulk_g2s_hint(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(lz_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(lz_smem_u32(bar)), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised streaming kernel: no CTA-wide barrier inside the chunk loop.
//   warp 0          : producer -- waits for a free ring slot, issues the two bulk copies
//   warps 1..GW     : gather   -- wait for the slice, turn it into products in place (x via LDG)
//   warps GW+1..    : rows     -- wait for the products, add them per row, run the epilogue
// Slots cycle through three mbarriers (full -> prod -> free), so the three roles work on
// different chunks at the same time and every role's global-load latency overlaps the others'.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lz_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lz_smem_u32(bar)) : "memory");
}

template <int MODE, int GW, int RW, int STAGES, int CAP, int MINB = 1>
__global__ void __launch_bounds__((1 + GW + RW) * 32, MINB)
k_csr_spmv_ws(const LzChunkRange cr, const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ chunk_ptr,
              const int32_t *__restrict__ chunk_ulen /* NULL: storage order */,
              const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
              const LzPassA args, const int blocked, const int hint)
{
    const int n_chunks = cr.total;
    auto cmap = [&](int v) { return v < cr.n0 ? cr.c0 + v : cr.c1 + (v - cr.n0); };
    constexpr int THREADS = (1 + GW + RW) * 32, GT = GW * 32, RT = RW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *vals_s = reinterpret_cast<double *>(smem_raw);
    int *cols_s = reinterpret_cast<int *>(smem_raw + sizeof(double) * (size_t)CAP * STAGES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + 12 * (size_t)CAP * STAGES);
    uint64_t *prod = full + STAGES, *freeb = prod + STAGES;
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { lz_mbar_init(&full[s], 1); lz_mbar_init(&prod[s], GW); lz_mbar_init(&freeb[s], RW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // chunk -> CTA map.  blocked: every CTA owns a contiguous range of chunks, so the x entries gathered for
    // the +-1 / +-nx neighbours of a stencil row are this SM's own recent rows (L1 hits) and all CTAs move
    // through their ranges in step (the +-nx*ny neighbours are a neighbouring CTA's rows: L2 hits).
    // strided: chunk c, c + grid, ... (long rows are spread over the CTAs).
    const int per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
    const int first = blocked ? blockIdx.x * per_cta : blockIdx.x, step = blocked ? 1 : gridDim.x;
    const int last = blocked ? min(n_chunks, first + per_cta) : n_chunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int p0 = 0, p1 = 0;
            if (first < last) { p0 = chunk_ptr[cmap(first)]; p1 = chunk_ptr[cmap(first) + 1]; }
            int it = 0;
            const uint64_t pol = lz_policy_evict_first();
            for (int c = first; c < last; c += step, ++it) {
                const int slot = it % STAGES;
                const int cp0 = p0, cp1 = p1;
                if (c + step < last) { p0 = chunk_ptr[cmap(c + step)]; p1 = chunk_ptr[cmap(c + step) + 1]; }
                lz_mbar_wait(&freeb[slot], ((it / STAGES) & 1) ^ 1);
                const int a0 = cp0 & ~3;
                const int cnt4 = (cp1 - a0) & ~3;
                if (cp1 - a0 > CAP || cnt4 == 0) { lz_mbar_arrive(&full[slot]); continue; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                lz_mbar_expect_tx(&full[slot], (uint32_t)cnt4 * 12u);
                if (hint) {     // the matrix is read once per product: keep L2 for x / w (re-read by pass B and the next pass A)
                    lz_bulk_g2s_hint(vals_s + (size_t)slot * CAP, vals + a0, (uint32_t)cnt4 * 8u, &full[slot], pol);
                    lz_bulk_g2s_hint(cols_s + (size_t)slot * CAP, colidx + a0, (uint32_t)cnt4 * 4u, &full[slot], pol);
                } else {
                    lz_bulk_g2s(vals_s + (size_t)slot * CAP, vals + a0, (uint32_t)cnt4 * 8u, &full[slot]);
                    lz_bulk_g2s(cols_s + (size_t)slot * CAP, colidx + a0, (uint32_t)cnt4 * 4u, &full[slot]);
                }
            }
        }
    } else if (warp <= GW) {
        // ------------------------------------------------------------------ gather warps
        const int gtid = tid - 32;
        int p0 = 0, p1 = 0, nr0 = 0, nr1 = 0, nul = 0;
        if (first < last) {
            const int c0 = cmap(first);
            p0 = chunk_ptr[c0]; p1 = chunk_ptr[c0 + 1];
            if (chunk_ulen) { nul = chunk_ulen[c0]; nr0 = chunk_row[c0]; nr1 = chunk_row[c0 + 1]; }
        }
        int it = 0;
        for (int c = first; c < last; c += step, ++it) {
            const int slot = it % STAGES;
            const int cp0 = p0, cp1 = p1, L = nul, R = nr1 - nr0;
            if (c + step < last) {
                const int cn = cmap(c + step);
                p0 = chunk_ptr[cn]; p1 = chunk_ptr[cn + 1];
                if (chunk_ulen) { nul = chunk_ulen[cn]; nr0 = chunk_row[cn]; nr1 = chunk_row[cn + 1]; }
            }
            const int a0 = cp0 & ~3, cnt = cp1 - a0, cnt4 = cnt & ~3;
            double *vs = vals_s + (size_t)slot * CAP;
            const int *cs = cols_s + (size_t)slot * CAP;
            lz_mbar_wait(&full[slot], (it / STAGES) & 1);
            if (cnt <= CAP) {
                if (gtid < cnt - cnt4) {       // tail the 16-byte-granular bulk copies cannot carry
                    const int k = cnt4 + gtid;
                    vs[k] = __dmul_rn(vals[a0 + k], epi.xs(__ldg(x + colidx[a0 + k])));
                }
                if (L > 0) {
                    // uniform chunk: R rows of L entries.  Walk it transposed -- t = s * R + r -> entry (cp0 - a0) + r * L + s --
                    // so a warp load gathers entry s of 32 consecutive rows (contiguous x for a stencil)
                    const int head = cp0 - a0, total = R * L;
                    const unsigned inv = 0xFFFFFFFFu / (unsigned)R + 1u;          // floor(t / R) = umulhi(t, inv) for t * R < 2^32
                    auto entry = [&](int t) -> int {                              // -1: past the chunk / in the tail
                        const int s = (int)__umulhi((unsigned)t, inv), r = t - s * R;
                        const int k = head + r * L + s;
                        return (t < total && k < cnt4) ? k : -1;
                    };
                    for (int base = gtid; base < total; base += 4 * GT) {       // (4 in flight: the kernel is register-tight)
                        double xv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = entry(base + u * GT);
                            xv[u] = k >= 0 ? __ldg(x + cs[k]) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = entry(base + u * GT);
                            if (k >= 0) vs[k] = __dmul_rn(vs[k], epi.xs(xv[u]));
                        }
                    }
                } else
                for (int base = gtid; base < cnt4; base += 8 * GT) {
                    double xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = base + u * GT;
                        xv[u] = 0.0;
                        if (k < cnt4) xv[u] = __ldg(x + cs[k]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = base + u * GT;
                        if (k < cnt4) vs[k] = __dmul_rn(vs[k], epi.xs(xv[u]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) lz_mbar_arrive(&prod[slot]);
        }
    } else {
        // ------------------------------------------------------------------ row warps
        const int rtid = tid - 32 * (1 + GW), rwarp = rtid >> 5;
        int nr0 = 0, nr1 = 0, np0 = 0, np1 = 0;
        if (first < last) { nr0 = chunk_row[cmap(first)]; nr1 = chunk_row[cmap(first) + 1]; np0 = chunk_ptr[cmap(first)]; np1 = chunk_ptr[cmap(first) + 1]; }
        int it = 0;
        for (int c = first; c < last; c += step, ++it) {
            const int slot = it % STAGES;
            const int r0 = nr0, r1 = nr1, cp0 = np0, cp1 = np1;
            if (c + step < last) {
                nr0 = chunk_row[cmap(c + step)]; nr1 = chunk_row[cmap(c + step) + 1];
                np0 = chunk_ptr[cmap(c + step)]; np1 = chunk_ptr[cmap(c + step) + 1];
            }
            const int a0 = cp0 & ~3, cnt = cp1 - a0;
            const double *vs = vals_s + (size_t)slot * CAP;
            // operands of this thread's first row before the wait: their latency is hidden by it
            const int rf = r0 + rtid;
            int s0 = 0, e0 = 0;
            double xo = 0.0, up = 0.0;
            if (rf < r1) { s0 = rowptr[rf] - a0; e0 = rowptr[rf + 1] - a0; epi.load(rf, xo, up); }
            lz_mbar_wait(&prod[slot], (it / STAGES) & 1);
            if (cnt <= CAP) {
                if (rf < r1) {
                    const int len = e0 - s0;
                    double pr[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) pr[k] = (k < len) ? vs[s0 + k] : 0.0;
                    double t = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k < len) t = __dadd_rn(t, pr[k]);
                    for (int k = s0 + 8; k < e0; ++k) t = __dadd_rn(t, vs[k]);
                    acc += epi.finish(rf, t, y, xo, up);
                }
                for (int r = rf + RT; r < r1; r += RT) {
                    const int s = rowptr[r] - a0, e = rowptr[r + 1] - a0;
                    double t = 0.0;
                    for (int k = s; k < e; ++k) t = __dadd_rn(t, vs[k]);
                    acc += epi.finish(r, t, y);
                }
            } else {
                // long-row chunk: the row warps walk global memory (warp per row; all of them per huge row)
                for (int r = r0 + rwarp; r < r1; r += RW) {
                    const int s = rowptr[r], e = rowptr[r + 1];
                    if (e - s > 4096) continue;
                    double t = 0.0;
                    for (int k = s + lane; k < e; k += 32) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                    t = lz_warp_sum(t);
                    if (lane == 0) acc += epi.finish(r, t, y);
                }
                for (int r = r0; r < r1; ++r) {
                    const int s = rowptr[r], e = rowptr[r + 1];
                    if (e - s <= 4096) continue;
                    double t = 0.0;
                    for (int k = s + rtid; k < e; k += RT) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                    t = lz_warp_sum(t);
                    // combine the RW warp partials through the (idle) product slots of this stage
                    double *scratch = vals_s + (size_t)slot * CAP;
                    asm volatile("bar.sync 1, %0;" ::"n"(RT));
                    if (lane == 0) scratch[rwarp] = t;
                    asm volatile("bar.sync 1, %0;" ::"n"(RT));
                    if (rtid == 0) {
                        double tt = 0.0;
                        for (int wv = 0; wv < RW; ++wv) tt += scratch[wv];
                        acc += epi.finish(r, tt, y);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) lz_mbar_arrive(&freeb[slot]);
        }
    }
    if (MODE == LZ_EPI_LANCZOS) {
        acc = lz_block_sum<THREADS>(acc, red);
        double total;
        if (lz_grid_sum<THREADS, 1>(&acc, args.partials, args.ticket, red, &total)) {
            if (tid == 0) lz_alpha_publish(args, total);
        }
    }
}

template <int MODE, int GW, int RW, int STAGES, int CAP, int MINB = 1>
static inline int lz_launch_ws_variant(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y, const LzPassA &args, int ctas_per_sm,
                                       int blocked = 0, bool coarse = false, int part = 0)
{
    // coarse: the LZ_SPMM_TILE schedule (plain SpMV prefers the larger chunks, profiles/r01_spmv_variants.md)
    const int n_chunks = coarse ? A->mm_n_chunks : A->n_chunks;
    const int32_t *chunk_row = coarse ? A->mm_chunk_row : A->chunk_row, *chunk_ptr = coarse ? A->mm_chunk_ptr : A->chunk_ptr;
    const size_t smem = (size_t)STAGES * CAP * 12 + 24 * STAGES;
    LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_csr_spmv_ws<MODE, GW, RW, STAGES, CAP, MINB>, (int)smem));
    // part 0: all chunks; 1: interior chunks of a sharded operator; 2: its two boundary ranges
    const int lo = coarse ? A->mm_bnd_lo : A->bnd_lo, hi = coarse ? A->mm_bnd_hi : A->bnd_hi;
    LzChunkRange cr = {0, n_chunks, 0, n_chunks};
    if (part == 1) cr = {lo, hi - lo, 0, hi - lo};
    if (part == 2) cr = {0, lo, hi, lo + (n_chunks - hi)};
    if (cr.total <= 0) return LZ_OK;
    int grid = ctx->sm_count * ctas_per_sm;
    if (grid > cr.total) grid = cr.total;
    const int32_t *ulen = ctx->knobs.no_transpose ? nullptr : (coarse ? A->mm_chunk_ulen : A->chunk_ulen);
    k_csr_spmv_ws<MODE, GW, RW, STAGES, CAP, MINB><<<grid, (1 + GW + RW) * 32, smem, ctx->stream>>>(
        cr, chunk_row, chunk_ptr, ulen, A->vrowptr ? A->vrowptr : A->rowptr, A->k_colidx, A->k_vals, x, y, args, blocked,
        ctx->knobs.spmv_hint);      // evict-first on the matrix streams measured 4-5 % slower: off
    return LZ_OK;
}

// width-4 row-interleaved ELL (data[4r+k], idx[4r+k]); zero padding entries carry idx 0
template <int MODE>
__global__ void __launch_bounds__(LZ_SPMV_THREADS)
k_ell4_spmv(int64_t n_rows, const double *__restrict__ data, const uint32_t *__restrict__ idx,
            const double *__restrict__ x, double *__restrict__ y, const LzPassA args)
{
    __shared__ double red[32];
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;
    const int64_t r = (int64_t)blockIdx.x * LZ_SPMV_THREADS + threadIdx.x;
    if (r < n_rows) {
        double v0, v1, v2, v3;
        lz_ld256(data + 4 * r, v0, v1, v2, v3);
        const uint4 c = *reinterpret_cast<const uint4 *>(idx + 4 * r);
        double t = __dmul_rn(v0, epi.xs(__ldg(x + c.x)));
        t = __dadd_rn(t, __dmul_rn(v1, epi.xs(__ldg(x + c.y))));
        t = __dadd_rn(t, __dmul_rn(v2, epi.xs(__ldg(x + c.z))));
        t = __dadd_rn(t, __dmul_rn(v3, epi.xs(__ldg(x + c.w))));
        acc = epi.finish(r, t, y);
    }
    lz_spmv_finalize<MODE>(acc, args, red);
}

// host-side launcher shared by lz_spmv() and the drivers
template <int MODE>
static inline int lz_launch_spmv(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y, const LzPassA &args, int part = 0)
{
    if (part != 0 && !(A->format == LZ_FMT_CSR && A->tma_ok && A->has_split)) {
        lz_set_error("interior/boundary launches need a sharded CSR operator with a chunk split");
        return LZ_ERR_INVALID;
    }
    if (A->format == LZ_FMT_ELL4) {
        const unsigned grid = (unsigned)((A->n_rows + LZ_SPMV_THREADS - 1) / LZ_SPMV_THREADS);
        k_ell4_spmv<MODE><<<grid, LZ_SPMV_THREADS, 0, ctx->stream>>>(A->n_rows, A->ell_data, A->ell_idx, x, y, args);
    } else if (A->tma_ok) {
        // fused modes: 1 producer + 3 gather + 5 row warps, 3-slot ring of 1024 entries, 5 CTAs per SM on the fine
        // (768-entry) schedule.  A stand-alone SpMV (x not L2-resident from a preceding pass B) runs faster with
        // 1 + 6 + 9 warps on the coarse (1536-entry) schedule.  LZ_SPMV_VARIANT=3 / 20 force coarse / fine for
        // every mode (profiles/r01_spmv_variants.md).
        const int v = ctx->spmv_variant;
        const bool coarse = A->mm_chunk_row && (!A->vrowptr || A->mm_shared) && (v == 3 || (MODE == LZ_EPI_PLAIN && v != 20));   // the coarse schedule of a row-split operator may index the SpMM's own split
        if (coarse) LZ_TRY((lz_launch_ws_variant<MODE, 6, 9, 3, 2048, 3>(ctx, A, x, y, args, 3, 0, true, part)));
        else LZ_TRY((lz_launch_ws_variant<MODE, 3, 5, 3, 1024, 5>(ctx, A, x, y, args, 5, 0, false, part)));
    } else {
        k_csr_spmv<MODE><<<A->n_chunks, LZ_SPMV_THREADS, 0, ctx->stream>>>(A->chunk_row, A->vrowptr ? A->vrowptr : A->rowptr, A->k_colidx, A->k_vals, x, y, args);
    }
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}


// ---------------------------------------------------------------------------------------------
// Row-split operators (power-law graphs): rows longer than LZ_SPLIT_L are cut into virtual rows of at
// most LZ_SPLIT_L entries (same colidx/vals arrays, finer row pointers), so every chunk of the
// schedule takes the streaming path and a 400k-entry hub row is spread over hundreds of CTAs.
// The kernels above then produce one partial sum per virtual row; this kernel adds the pieces of each
// real row in order (deterministic) and runs the epilogue.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
k_split_combine(int64_t n_rows, const int32_t *__restrict__ vstart, const int32_t *__restrict__ vpos, const double *__restrict__ ybar,
                double *__restrict__ y, const LzPassA args)
{
    __shared__ double red[32];
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n_rows; r += stride) {
        const int v0 = vstart[r], v1 = vstart[r + 1];       // piece v of the row sits at ybar[vpos[v]] in the binned order
        double t = ybar[vpos ? vpos[v0] : v0];
        for (int v = v0 + 1; v < v1; ++v) t += ybar[vpos ? vpos[v] : v];
        acc += epi.finish(r, t, y);
    }
    if (MODE == LZ_EPI_LANCZOS) {
        acc = lz_block_sum<256>(acc, red);
        double total;
        if (lz_grid_sum<256, 1>(&acc, args.partials, args.ticket, red, &total)) {
            if (threadIdx.x == 0) lz_alpha_publish(args, total);
        }
    }
}

template <int MODE>
static inline int lz_spmv_any(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y, const LzPassA &args, int part = 0)
{
    if (!A->vrowptr) return lz_launch_spmv<MODE>(ctx, A, x, y, args, part);
    LZ_CHECK(part == 0, LZ_ERR_INVALID, "row-split operators cannot be launched in parts");
    if (MODE == LZ_EPI_LANCZOS) LZ_TRY(lz_launch_spmv<LZ_EPI_SCALED>(ctx, A, x, A->ybar, args));
    else LZ_TRY(lz_launch_spmv<LZ_EPI_PLAIN>(ctx, A, x, A->ybar, args));
    int64_t want = (A->n_rows + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    k_split_combine<MODE><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(A->n_rows, A->vstart, A->vpos, A->ybar, y, args);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}
