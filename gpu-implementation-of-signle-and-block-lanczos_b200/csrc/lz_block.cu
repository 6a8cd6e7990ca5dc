// lz_block.cu -- block path: SpMM kernels and the block Lanczos driver.
//
// Replaces ell::SpMM (kernels/spmv_spmm.hpp:137-199), lm::spmm_basic (kernels/ell_kernels.hpp:37-61)
// and block_lanczos / block_lanczos_blas (methods/block_lanczos.hpp:13-167).
//
// Internal layout of every n x bw block is ROW-MAJOR (bw contiguous doubles per row): one gathered
// row of the dense block is a single 64..256-byte segment, where the reference's column-major
// layout costs bw separate sectors per non-zero.  The C-ABI keeps the reference's
// column-major layout; the driver converts once on the way in.
#include <math.h>

#include <algorithm>
#include <vector>

#include "lz_dense_host.cuh"
#include "lz_spmv.cuh"

int lz_sym_eig_full_c(int N, double *A, std::vector<double> &d, std::vector<double> &Zt);   // lz_ritz.cu

#define SPMM_THREADS 256
#define SPMM_U 2            // independent rows per lane group in flight
#define SPMM_SLAB 512     // rows per CTA slab of the row-major SpMM

// ---------------------------------------------------------------------------------------------
// Row-major SpMM, CSR:  W[i,:] = sum_k A[i,k] X[col_k,:]   (- Q0[i,:] B  when FUSE_SUB)
// A group of LW = BW/2 lanes owns a row; each lane carries two adjacent columns (128-bit loads of
// the gathered row, 128-bit stores).  Rows are dealt to groups in contiguous slabs so neighbouring
// rows -- which share most of their gathered rows for banded operators -- sit in the same warp/CTA
// and hit L1.  The B (bw x bw, column-major) operand of the fused subtraction lives in shared memory.
// ---------------------------------------------------------------------------------------------
template <int BW, bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm(int64_t n_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
          const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ W,
          const double *__restrict__ Q0, const double *__restrict__ Bm)
{
    constexpr int LW = BW / 2;                       // lanes per row
    constexpr int RPW = 32 / LW;                     // rows per warp at a time
    constexpr int WARPS = SPMM_THREADS / 32;
    // B of the fused subtraction, packed so that one conflict-free 128-bit shared load per lane
    // brings the two rows (2p, 2p+1) of one of the lane's two columns: blo[p][l] = B[2p..2p+1, 2l],
    // bhi[p][l] = B[2p..2p+1, 2l+1]
    __shared__ double2 blo[FUSE_SUB ? LW * LW : 1], bhi[FUSE_SUB ? LW * LW : 1];
    if (FUSE_SUB) {
        for (int e = threadIdx.x; e < LW * LW; e += SPMM_THREADS) {
            const int p = e / LW, l = e % LW;
            blo[e] = make_double2(Bm[(2 * p) + (2 * l) * BW], Bm[(2 * p + 1) + (2 * l) * BW]);
            bhi[e] = make_double2(Bm[(2 * p) + (2 * l + 1) * BW], Bm[(2 * p + 1) + (2 * l + 1) * BW]);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, sub = lane / LW, l = lane % LW, warp = threadIdx.x >> 5;
    // A CTA owns contiguous slabs of SPMM_SLAB rows (slab index strided by the grid) and sweeps each
    // slab front to back with all its warps, so the rows gathered for neighbouring matrix rows --
    // the +-1 and +-nx neighbours of a stencil -- are still in this SM's L1 when they are needed
    // again, and concurrently running CTAs stay within a narrow window of X (L2 reuse of +-nx*ny).
    // Every warp keeps U independent rows per lane group in flight and fetches the row extents of
    // its next trip while it works on the current one: the rowptr -> (col,val) -> X[col] chain is
    // three dependent global loads, and this kernel lives or dies by how many of them overlap.
    constexpr int U = SPMM_U;
    constexpr int TRIPS = SPMM_SLAB / (WARPS * RPW * U);
    const int64_t n_slabs = (n_rows + SPMM_SLAB - 1) / SPMM_SLAB;
    auto row_of = [&](int64_t slab, int t, int u) -> int64_t {
        return slab * SPMM_SLAB + (int64_t)((t * U + u) * WARPS + warp) * RPW + sub;
    };
    int ns[U], ne[U];
    {
        const int64_t slab = blockIdx.x;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = row_of(slab, 0, u);
            ns[u] = (slab < n_slabs && r < n_rows) ? rowptr[r] : 0;
            ne[u] = (slab < n_slabs && r < n_rows) ? rowptr[r + 1] : 0;
        }
    }
    for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        for (int t = 0; t < TRIPS; ++t) {
            int64_t r[U]; int s[U], e[U]; bool valid[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { r[u] = row_of(slab, t, u); valid[u] = r[u] < n_rows; s[u] = ns[u]; e[u] = ne[u]; }
            {   // extents of the next trip (possibly in the next slab of this CTA)
                const int64_t nslab = (t + 1 < TRIPS) ? slab : slab + gridDim.x;
                const int nt = (t + 1 < TRIPS) ? t + 1 : 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t rn = row_of(nslab, nt, u);
                    const bool ok = nslab < n_slabs && rn < n_rows;
                    ns[u] = ok ? rowptr[rn] : 0;
                    ne[u] = ok ? rowptr[rn + 1] : 0;
                }
            }
            double2 q[U];
            if (FUSE_SUB) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    q[u] = valid[u] ? __ldcs(reinterpret_cast<const double2 *>(Q0 + r[u] * BW) + l) : make_double2(0.0, 0.0);
            }
            // uniform trip count across the warp: the group shuffles below need every lane
            int maxlen = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) maxlen = max(maxlen, e[u] - s[u]);
#pragma unroll
            for (int o = LW; o < 32; o <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
            double a0[U], a1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) a0[u] = a1[u] = 0.0;
            for (int k0 = 0; k0 < maxlen; k0 += LW) {
                // the group's lanes fetch LW consecutive (col, val) pairs of each row in one coalesced
                // load, then hand them round with shuffles; streaming loads: read exactly once
                int my_c[U]; double my_v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = s[u] + k0 + l;
                    my_c[u] = (k < e[u]) ? __ldcs(colidx + k) : -1;
                    my_v[u] = (k < e[u]) ? __ldcs(vals + k) : 0.0;
                }
                constexpr int G = LW < 4 ? LW : 4;   // U*G gathers in flight per lane
#pragma unroll
                for (int u0 = 0; u0 < LW; u0 += G) {
                    int c[U][G]; double v[U][G]; double2 x[U][G];
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            c[u][g] = __shfl_sync(0xffffffffu, my_c[u], sub * LW + u0 + g);
                            v[u][g] = __shfl_sync(0xffffffffu, my_v[u], sub * LW + u0 + g);
                        }
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int g = 0; g < G; ++g)
                            x[u][g] = c[u][g] >= 0 ? __ldg(reinterpret_cast<const double2 *>(X + (int64_t)c[u][g] * BW) + l)
                                                   : make_double2(0.0, 0.0);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int g = 0; g < G; ++g) { a0[u] = fma(v[u][g], x[u][g].x, a0[u]); a1[u] = fma(v[u][g], x[u][g].y, a1[u]); }
                }
            }
            if (FUSE_SUB) {
                // W[i, 2l..2l+1] -= sum_p Q0[i,p] B[p, 2l..2l+1]; Q0 row handed round the group's lanes
#pragma unroll
                for (int p = 0; p < LW; ++p) {
                    const double2 b0 = blo[p * LW + l], b1 = bhi[p * LW + l];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const double q0 = __shfl_sync(0xffffffffu, q[u].x, sub * LW + p);
                        const double q1 = __shfl_sync(0xffffffffu, q[u].y, sub * LW + p);
                        a0[u] = fma(-q0, b0.x, a0[u]); a0[u] = fma(-q1, b0.y, a0[u]);
                        a1[u] = fma(-q0, b1.x, a1[u]); a1[u] = fma(-q1, b1.y, a1[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (valid[u]) __stcs(reinterpret_cast<double2 *>(W + r[u] * BW) + l, make_double2(a0[u], a1[u]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged row-major SpMM (default for BW in {8,16,32} on operators with a chunk schedule).
// The rowptr -> (col,val) -> X[col] chain of the kernel above is three dependent global loads per
// row; here warp 0 streams each chunk's vals / colidx / rowptr slices into a shared-memory ring
// with bulk async copies (the same producer as k_csr_spmv_ws), so a compute warp's only global
// loads are the gathered rows of X (plus the independent Q0 row and the W store): one round trip
// per row.  A group of LW lanes owns a row, two rows per group in flight.
// ---------------------------------------------------------------------------------------------
#define SPMM_WS_RCAP 1024      // rowptr entries staged per chunk
#define LZ_SPMM_GRAM_CW 11     // compute warps of the variant that also accumulates the Gram block (16 more registers per thread)

__device__ __forceinline__ void lz_ld256_ro(const double *p, double &a, double &b, double &c, double &d)
{
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// the same loads / stores with an L2 eviction policy (createpolicy descriptor): the gathered panel is the only
// operand of the SpMM that is re-read, so it can be pinned (evict-last) while the streams pass through (evict-first)
__device__ __forceinline__ void lz_ld256_ro_pol(const double *p, double &a, double &b, double &c, double &d, uint64_t pol)
{
    asm("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void lz_ld256_stream_pol(const double *p, double &a, double &b, double &c, double &d, uint64_t pol)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void lz_st256_pol(double *p, double a, double b, double c, double d, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t lz_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// A group of LW = BW/4 lanes owns a row and every lane carries FOUR adjacent columns (one 256-bit
// load per gathered row): the kernel is bound by instruction issue, and four columns per lane halve
// the instructions per non-zero against the two-column layout of k_spmm_rm.
//
// FSUB (BW = 16 only): W = A X - Q0 B in the same pass, the reference's order of the block recurrence
// (W = A Q_j; W -= Q_{j-1} beta_j BEFORE alpha_j is formed, methods/block_lanczos.hpp:149-155).  With 4 lanes per
// row a warp trip is an 8-row x 16-column tile whose accumulators already ARE the C fragments of two
// mma.m8n8k4 n-tiles if the tile columns are labelled  n-tile t, slot s  <->  column 4(s>>1) + 2t + (s&1);
// the Q0 row segment a lane loads (columns 4l..4l+3, one 256-bit load) is its A fragment for the K labelling
// k-tile t, slot l  <->  column 4l + t.  The subtraction is then 8 DMMAs per trip with no data movement;
// only the B fragments of -B (fragment-ordered in shared memory) follow the two labellings.
//
// GRAM (with FSUB): the CTA also accumulates  G_partial = Xown^T W  over its rows -- the Q_j^T (A Q_j - Q_{j-1} beta_j) of
// the block recurrence (:155), which otherwise costs a separate pass over two panels.  The 8 x 16 tile of a trip goes
// through a per-warp shared-memory tile (the MMA's K index runs over ROWS, which live in different lanes of the
// accumulator layout), the A fragments are the trip's own rows of X read straight from global (L1 hits: a stencil row
// has just gathered its diagonal neighbour), 8 more DMMAs per trip.
#define SPMM_GST 20            // row stride (doubles) of the per-warp Gram staging tile: conflict-free fragment reads
// DST (row-split operators): row r of the walked operator goes to row dst[r] of Wd when dst[r] >= 0 (a row that was not
// cut: no partial row, no combine), to row r of W (the partial-row buffer) otherwise.
template <int BW, int CW, int STAGES, int CAP, int MINB, bool FSUB, bool GRAM = false, int GIN = 4, bool DST = false>
__global__ void __launch_bounds__((1 + CW) * 32, MINB)
k_spmm_ws(int n_chunks, int64_t n_rows, const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ chunk_ptr,
          const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
          const double *__restrict__ X, double *__restrict__ W, const double *__restrict__ Q0, const double *__restrict__ Bm,
          const LzChunkRange cr, const int run, const int hint, const int64_t ldx, const int64_t ldw,
          const double *__restrict__ Xown = nullptr, double *__restrict__ gpart = nullptr,
          const int32_t *__restrict__ dst = nullptr, double *__restrict__ Wd = nullptr)
{
    static_assert(!DST || !FSUB, "direct rows are for the plain product");
    static_assert(!GRAM || FSUB, "the fused Gram rides on the 8-row trips of the fused subtraction");
    // ldx / ldw: doubles between consecutive rows of X / W (= BW for a whole panel; a BW-column slice of a wider
    // row-major panel otherwise: power-law operators run wide panels slice by slice so that the rows gathered again
    // and again -- the hubs' -- fit in L2)
    static_assert(!FSUB || BW == 16, "the fused subtraction is written for 16-column panels");
    constexpr int LW = BW / 4, RPW = 32 / LW, NG = CW * RPW, G = GIN;     // G gathered rows in flight per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *vals_s = reinterpret_cast<double *>(smem_raw);
    int *cols_s = reinterpret_cast<int *>(smem_raw + 8 * (size_t)CAP * STAGES);
    int *rptr_s = reinterpret_cast<int *>(smem_raw + 12 * (size_t)CAP * STAGES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + 12 * (size_t)CAP * STAGES + 4 * (size_t)SPMM_WS_RCAP * STAGES);
    uint64_t *freeb = full + STAGES;
    __shared__ double sbs[FSUB ? 8 * 32 : 1];           // -B in fragment order: sbs[(kt*2 + nt)*32 + lane]
    __shared__ __align__(16) double gst[GRAM ? CW * 8 * SPMM_GST : 1];    // per-warp 8 x 16 tile of W for the Gram fragments
    double gacc[GRAM ? 2 : 1][GRAM ? 2 : 1][2];
    if (GRAM) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) gacc[a][b][0] = gacc[a][b][1] = 0.0;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { lz_mbar_init(&full[s], 1); lz_mbar_init(&freeb[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (FSUB) {
        for (int e = tid; e < 8 * 32; e += (1 + CW) * 32) {
            const int ln = e & 31, nt = (e >> 5) & 1, kt = e >> 6;
            const int kk = ln & 3, mm = ln >> 2;
            const int kphys = 4 * kk + kt, nphys = 4 * (mm >> 1) + 2 * nt + (mm & 1);
            sbs[e] = -Bm[kphys + nphys * BW];
        }
    }
    __syncthreads();
    // chunk -> CTA map: runs of `run` consecutive chunks dealt round-robin to the CTAs.  Inside a run the rows
    // gathered for the near neighbours of a stencil row are this CTA's own recent rows (L1 hits); all CTAs
    // together sweep a window of grid*run chunks, so the far (+-nx*ny) neighbours were touched one window
    // earlier and are still in L2 (with one contiguous range per CTA they were DRAM misses: ncu, L2 hit 15 %).
    // Virtual chunk v of the launch is chunk cmap(v) of the schedule (interior / boundary launches of a shard).
    const int n_virtual = cr.total;
    auto vchunk = [&](int it) { return ((it / run) * (int)gridDim.x + (int)blockIdx.x) * run + (it % run); };
    auto cmap = [&](int v) { return v < cr.n0 ? cr.c0 + v : cr.c1 + (v - cr.n0); };
    (void)n_chunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int p0 = 0, p1 = 0, r0 = 0, r1 = 0;
            int v = vchunk(0);
            if (v < n_virtual) { const int c = cmap(v); p0 = chunk_ptr[c]; p1 = chunk_ptr[c + 1]; r0 = chunk_row[c]; r1 = chunk_row[c + 1]; }
            const uint64_t pol = lz_policy_evict_first();
            for (int it = 0; v < n_virtual; ++it) {
                const int slot = it % STAGES;
                const int cp0 = p0, cp1 = p1, cr0 = r0, cr1 = r1;
                v = vchunk(it + 1);
                if (v < n_virtual) { const int c = cmap(v); p0 = chunk_ptr[c]; p1 = chunk_ptr[c + 1]; r0 = chunk_row[c]; r1 = chunk_row[c + 1]; }
                lz_mbar_wait(&freeb[slot], ((it / STAGES) & 1) ^ 1);
                const int a0 = cp0 & ~3, cnt4 = (cp1 - a0) & ~3;
                const int ra = cr0 & ~3;
                const int rcnt = ((cr1 + 1 - ra) + 3) & ~3;               // rowptr[ra .. cr1] rounded up to 16 bytes
                const bool rows_ok = rcnt <= SPMM_WS_RCAP && (int64_t)ra + rcnt <= n_rows + 1;
                if (cp1 - a0 > CAP || cnt4 == 0) { lz_mbar_arrive(&full[slot]); continue; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                lz_mbar_expect_tx(&full[slot], (uint32_t)cnt4 * 12u + (rows_ok ? (uint32_t)rcnt * 4u : 0u));
                if (hint & 1) {
                    lz_bulk_g2s_hint(vals_s + (size_t)slot * CAP, vals + a0, (uint32_t)cnt4 * 8u, &full[slot], pol);
                    lz_bulk_g2s_hint(cols_s + (size_t)slot * CAP, colidx + a0, (uint32_t)cnt4 * 4u, &full[slot], pol);
                    if (rows_ok) lz_bulk_g2s_hint(rptr_s + (size_t)slot * SPMM_WS_RCAP, rowptr + ra, (uint32_t)rcnt * 4u, &full[slot], pol);
                } else {
                    lz_bulk_g2s(vals_s + (size_t)slot * CAP, vals + a0, (uint32_t)cnt4 * 8u, &full[slot]);
                    lz_bulk_g2s(cols_s + (size_t)slot * CAP, colidx + a0, (uint32_t)cnt4 * 4u, &full[slot]);
                    if (rows_ok) lz_bulk_g2s(rptr_s + (size_t)slot * SPMM_WS_RCAP, rowptr + ra, (uint32_t)rcnt * 4u, &full[slot]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int sub = lane / LW, l = lane % LW;
        int nr0 = 0, nr1 = 0, np0 = 0, np1 = 0, trip_base = 0;
        const bool rotate = !(hint & 16);         // hint bit 16 (LZ_SPMM_HINT=16): A/B switch, restart the deal per chunk
        int v = vchunk(0);
        if (v < n_virtual) { const int c = cmap(v); nr0 = chunk_row[c]; nr1 = chunk_row[c + 1]; np0 = chunk_ptr[c]; np1 = chunk_ptr[c + 1]; }
        for (int it = 0; v < n_virtual; ++it) {
            const int slot = it % STAGES;
            const int r0 = nr0, r1 = nr1, cp0 = np0, cp1 = np1;
            v = vchunk(it + 1);
            if (v < n_virtual) {
                const int c = cmap(v);
                nr0 = chunk_row[c]; nr1 = chunk_row[c + 1];
                np0 = chunk_ptr[c]; np1 = chunk_ptr[c + 1];
            }
            const int a0 = cp0 & ~3, cnt = cp1 - a0, cnt4 = cnt & ~3;
            const int ra = r0 & ~3;
            const int rcnt = ((r1 + 1 - ra) + 3) & ~3;
            const bool staged = cnt <= CAP && cnt4 > 0;
            const bool rows_ok = staged && rcnt <= SPMM_WS_RCAP && (int64_t)ra + rcnt <= n_rows + 1;
            const double *vs = vals_s + (size_t)slot * CAP;
            const int *cs = cols_s + (size_t)slot * CAP;
            const int *rs = rptr_s + (size_t)slot * SPMM_WS_RCAP;
            lz_mbar_wait(&full[slot], (it / STAGES) & 1);
            // warp-uniform trips over the chunk's rows: RPW rows per trip, row of this lane group = rb + sub.  The trips of
            // ALL chunks of this CTA are dealt round-robin to the warps (trip_base continues across chunks): a chunk has
            // 27.5 trips on 11-12 warps, and restarting the deal at warp 0 for every chunk gave the same warps the third
            // trip each time (the ring lets a warp run one chunk ahead, so the rotation evens the load out)
            const int trips = (int)((r1 - r0 + RPW - 1) / RPW);
            const int t0 = (((warp - 1) - trip_base) % CW + CW) % CW;
            trip_base = rotate ? (trip_base + trips) % CW : 0;
            for (int64_t rb = (int64_t)r0 + (int64_t)t0 * RPW; rb < r1; rb += NG) {
                const int64_t r = rb + sub;
                const bool valid = r < r1;
                int s = 0, e = 0;
                if (valid) {
                    if (rows_ok) { s = rs[r - ra]; e = rs[r - ra + 1]; }
                    else { s = rowptr[r]; e = rowptr[r + 1]; }
                }
                int drow = -1;
                if (DST && valid) drow = dst[r];
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                const double *Xl = X + 4 * l;
                const int64_t ldx_ = FSUB ? (int64_t)BW : ldx, ldw_ = FSUB ? (int64_t)BW : ldw;
                for (int k0 = s; k0 < e; k0 += G) {
                    // (col,val) straight from the staged slice: a broadcast shared load per entry
                    int cc[G]; double vv[G]; double x0[G], x1[G], x2[G], x3[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int k = k0 + g, ks = k - a0;
                        cc[g] = -1; vv[g] = 0.0;
                        if (k < e) {
                            if (staged && ks < cnt4) { cc[g] = cs[ks]; vv[g] = vs[ks]; }
                            else { cc[g] = __ldcs(colidx + k); vv[g] = __ldcs(vals + k); }
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        x0[g] = x1[g] = x2[g] = x3[g] = 0.0;
                        if (cc[g] >= 0) lz_ld256_ro(Xl + (int64_t)cc[g] * ldx_, x0[g], x1[g], x2[g], x3[g]);
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        acc0 = fma(vv[g], x0[g], acc0); acc1 = fma(vv[g], x1[g], acc1);
                        acc2 = fma(vv[g], x2[g], acc2); acc3 = fma(vv[g], x3[g], acc3);
                    }
                }
                if (FSUB) {
                    // this lane's Q0 row segment = its A fragments.  Streams (Q0, W) carry an L2 evict-first policy: without
                    // it they push the gathered panel's reuse window out of L2 (2.4 -> 2.0 ms, profiles/r02_spmm.md)
                    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                    if (valid) lz_ld256_stream_pol(Q0 + r * BW + 4 * l, q0, q1, q2, q3, lz_policy_evict_first());
                    __syncwarp();     // the gather loop above has per-row trip counts: the MMAs need the whole warp
                    // (acc0,acc1) / (acc2,acc3) are the C fragments of n-tiles 0 / 1, q0..q3 the A fragments of k-tiles 0..3
                    lz_dmma(acc0, acc1, q0, sbs[(0 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q0, sbs[(0 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q1, sbs[(1 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q1, sbs[(1 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q2, sbs[(2 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q2, sbs[(2 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q3, sbs[(3 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q3, sbs[(3 * 2 + 1) * 32 + lane]);
                }
                if (valid) {
                    if (FSUB) lz_st256_pol(W + r * ldw_ + 4 * l, acc0, acc1, acc2, acc3, lz_policy_evict_first());
                    else if (DST && drow >= 0) lz_st256(Wd + (int64_t)drow * ldw_ + 4 * l, acc0, acc1, acc2, acc3);
                    else lz_st256(W + r * ldw_ + 4 * l, acc0, acc1, acc2, acc3);
                }
                if (GRAM) {
                    // G += Xown[rb .. rb+8, :]^T W[rb .. rb+8, :]   (rows past the chunk contribute zeros: their acc is 0)
                    double *gt = gst + (warp - 1) * 8 * SPMM_GST;
                    *reinterpret_cast<double2 *>(gt + sub * SPMM_GST + 4 * l) = make_double2(acc0, acc1);
                    *reinterpret_cast<double2 *>(gt + sub * SPMM_GST + 4 * l + 2) = make_double2(acc2, acc3);
                    const int kk = lane & 3, mm = lane >> 2;
                    double xa[2][2];
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int64_t rr = rb + 4 * ks + kk;
#pragma unroll
                        for (int a = 0; a < 2; ++a) xa[ks][a] = rr < r1 ? __ldg(Xown + rr * BW + 8 * a + mm) : 0.0;
                    }
                    __syncwarp();
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const double wb0 = gt[(4 * ks + kk) * SPMM_GST + mm], wb1 = gt[(4 * ks + kk) * SPMM_GST + 8 + mm];
#pragma unroll
                        for (int a = 0; a < 2; ++a) {
                            lz_dmma(gacc[a][0][0], gacc[a][0][1], xa[ks][a], wb0);
                            lz_dmma(gacc[a][1][0], gacc[a][1][1], xa[ks][a], wb1);
                        }
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            if (lane == 0) lz_mbar_arrive(&freeb[slot]);
        }
    }
    if (GRAM) {
        // CTA partial = sum of the compute warps' accumulators in warp order (fixed), through shared memory
        __shared__ double gsm[16 * 16];
        const int kk = lane & 3, mm = lane >> 2;
        for (int w = 1; w <= CW; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                        if (w == 1) { gsm[p + q * 16] = gacc[a][b][0]; gsm[p + (q + 1) * 16] = gacc[a][b][1]; }
                        else { gsm[p + q * 16] += gacc[a][b][0]; gsm[p + (q + 1) * 16] += gacc[a][b][1]; }
                    }
            }
            __syncthreads();
        }
        for (int e = tid; e < 256; e += (1 + CW) * 32) gpart[(size_t)blockIdx.x * 256 + e] = gsm[e];
    }
}

template <int BW, int CW, int STAGES, int MINB, bool FSUB, bool GRAM = false, int GIN = 4, bool DST = false>
static int launch_spmm_ws_shape(lz_ctx *ctx, const lz_matrix *A, const int32_t *rowptr, int64_t n_rows, const double *X, double *W,
                                const double *Q0, const double *Bm, int run, int part, int64_t ldx = BW, int64_t ldw = BW,
                                const double *Xown = nullptr, double *gpart = nullptr, int *grid_out = nullptr,
                                const int32_t *dst = nullptr, double *Wd = nullptr)
{
    constexpr int CAP = 2048;
    const size_t smem = (size_t)STAGES * (CAP * 12 + SPMM_WS_RCAP * 4) + 16 * STAGES;
    LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_spmm_ws<BW, CW, STAGES, CAP, MINB, FSUB, GRAM, GIN, DST>, (int)smem));
    const int nch = A->mm_n_chunks;
    LzChunkRange cr = {0, nch, 0, nch};
    if (part == 1) cr = {A->mm_bnd_lo, A->mm_bnd_hi - A->mm_bnd_lo, 0, A->mm_bnd_hi - A->mm_bnd_lo};
    if (part == 2) cr = {0, A->mm_bnd_lo, A->mm_bnd_hi, A->mm_bnd_lo + (nch - A->mm_bnd_hi)};
    if (cr.total <= 0) return LZ_OK;
    int grid = ctx->sm_count * MINB;
    if (grid > cr.total) grid = cr.total;
    const int per_cta = (cr.total + grid - 1) / grid;
    if (grid_out) *grid_out = grid;
    k_spmm_ws<BW, CW, STAGES, CAP, MINB, FSUB, GRAM, GIN, DST><<<grid, (1 + CW) * 32, smem, ctx->stream>>>(
        nch, n_rows, A->mm_chunk_row, A->mm_chunk_ptr, rowptr, A->mm_k_colidx, A->mm_k_vals, X, W, Q0, Bm, cr, run > 0 ? run : per_cta,
        ctx->knobs.spmm_hint >= 0 ? ctx->knobs.spmm_hint : 0, ldx, ldw, Xown, gpart, dst, Wd);   // hint bit 1: evict-first on the matrix streams
    return LZ_OK;
}

template <int BW>
static int launch_spmm_ws(lz_ctx *ctx, const lz_matrix *A, const int32_t *rowptr, int64_t n_rows, const double *X, double *W,
                          const double *Q0, const double *Bm, int part, const double *Xown = nullptr, double *gpart = nullptr,
                          int *grid_out = nullptr)
{
    // 12 compute warps, 2-slot ring, 2 CTAs per SM (other shapes: profiles/r01_spmv_variants.md).
    // dev-time knob LZ_SPMM_RUN = chunks per run of the chunk map (default 1; 0: one contiguous range per CTA)
    const int run = ctx->knobs.spmm_run;
    if constexpr (BW == 16) {
        // LZ_SPMM_SHAPE=1 (A/B): 8 gathered rows in flight per lane (a 7-point row in ONE round trip instead of two) at
        // 7-8 compute warps per CTA, against 4 in flight at 11-12 warps
        if (ctx->knobs.spmm_shape == 1) {
            if (Q0 && gpart) return launch_spmm_ws_shape<16, 7, 2, 2, true, true, 8>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part, 16, 16, Xown, gpart, grid_out);
            if (Q0) return launch_spmm_ws_shape<16, 8, 2, 2, true, false, 8>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part);
            return launch_spmm_ws_shape<16, 8, 2, 2, false, false, 8>(ctx, A, rowptr, n_rows, X, W, nullptr, nullptr, run, part);
        }
        if (ctx->knobs.spmm_shape == 2) {
            if (Q0 && gpart) return launch_spmm_ws_shape<16, 9, 2, 2, true, true, 6>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part, 16, 16, Xown, gpart, grid_out);
            if (Q0) return launch_spmm_ws_shape<16, 10, 2, 2, true, false, 6>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part);
            return launch_spmm_ws_shape<16, 10, 2, 2, false, false, 6>(ctx, A, rowptr, n_rows, X, W, nullptr, nullptr, run, part);
        }
        if (Q0 && gpart) return launch_spmm_ws_shape<16, LZ_SPMM_GRAM_CW, 2, 2, true, true>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part, 16, 16, Xown, gpart, grid_out);
        if (Q0) return launch_spmm_ws_shape<16, 12, 2, 2, true>(ctx, A, rowptr, n_rows, X, W, Q0, Bm, run, part);
    }
    return launch_spmm_ws_shape<BW, 12, 2, 2, false>(ctx, A, rowptr, n_rows, X, W, nullptr, nullptr, run, part);
}

// any bw <= 32 (odd widths, bw = 1): one lane per column
template <bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm_any(int bw, int64_t n_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ W,
              const double *__restrict__ Q0, const double *__restrict__ Bm)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (SPMM_THREADS / 32);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        if (lane >= bw) continue;
        double a = 0.0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) a = fma(vals[k], __ldg(X + (int64_t)colidx[k] * bw + lane), a);
        if (FUSE_SUB)
            for (int p = 0; p < bw; ++p) a = fma(-Q0[r * bw + p], Bm[p + lane * bw], a);
        W[r * bw + lane] = a;
    }
}

// width-4 row-interleaved ELL (the reference's device format) as the operator of the row-major SpMM
template <int BW, bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm_ell4(int64_t n_rows, const double *__restrict__ data, const uint32_t *__restrict__ idx,
               const double *__restrict__ X, double *__restrict__ W, const double *__restrict__ Q0,
               const double *__restrict__ Bm)
{
    const int64_t r = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) / BW;
    const int c = threadIdx.x % BW;                  // BW divides 256 for BW in {1,2,4,8,16,32}
    if (r >= n_rows) return;
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) a = fma(data[4 * r + k], __ldg(X + (int64_t)idx[4 * r + k] * BW + c), a);
    if (FUSE_SUB)
        for (int p = 0; p < BW; ++p) a = fma(-Q0[r * BW + p], Bm[p + c * BW], a);
    W[r * BW + c] = a;
}

// column-major SpMM of the C-ABI (reference layout, spmm(): kernels/spmv_spmm.hpp:262-333): one thread per row keeps
// its row of A in registers and walks the b columns FOUR at a time, so 4 x (row length) independent gathers are in
// flight per thread; for banded operators the gathers of a warp's 32 consecutive rows coalesce into 256-byte
// segments of each column, and Y is written in full lines.  Products and sums are separate roundings in the
// row's storage order (the reference Host loop, objects/ell_matrix.hpp:246-251): bit-identical to that loop.
template <int CG>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_cm(int64_t n_rows, int b, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
          const double *__restrict__ vals, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy)
{
    const int64_t r = (int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x;
    if (r >= n_rows) return;
    const int s = rowptr[r], e = rowptr[r + 1];
    if (e - s <= 8) {
        int c[8]; double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = (s + k < e) ? colidx[s + k] : 0; v[k] = (s + k < e) ? vals[s + k] : 0.0; }
        int col = 0;
        for (; col + CG <= b; col += CG) {
            double t[CG], x[CG][8];
#pragma unroll
            for (int u = 0; u < CG; ++u) {
                const double *xc = X + (int64_t)(col + u) * ldx;
#pragma unroll
                for (int k = 0; k < 8; ++k) x[u][k] = (s + k < e) ? __ldg(xc + c[k]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CG; ++u) {
                t[u] = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) if (s + k < e) t[u] = __dadd_rn(t[u], __dmul_rn(v[k], x[u][k]));
                Y[r + (int64_t)(col + u) * ldy] = t[u];
            }
        }
        for (; col < b; ++col) {
            const double *xc = X + (int64_t)col * ldx;
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (s + k < e) t = __dadd_rn(t, __dmul_rn(v[k], __ldg(xc + c[k])));
            Y[r + (int64_t)col * ldy] = t;
        }
    } else {
        for (int col = 0; col < b; ++col) {
            const double *xc = X + (int64_t)col * ldx;
            double t = 0.0;
            for (int k = s; k < e; ++k) t = __dadd_rn(t, __dmul_rn(vals[k], __ldg(xc + colidx[k])));
            Y[r + (int64_t)col * ldy] = t;
        }
    }
}

__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_cm_ell4(int64_t n_rows, int b, const double *__restrict__ data, const uint32_t *__restrict__ idx,
               const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy)
{
    const int64_t r = (int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x;
    if (r >= n_rows) return;
    double v0, v1, v2, v3;
    lz_ld256(data + 4 * r, v0, v1, v2, v3);
    const uint4 c = *reinterpret_cast<const uint4 *>(idx + 4 * r);
    for (int col = 0; col < b; ++col) {
        const double *xc = X + (int64_t)col * ldx;
        double t = __dmul_rn(v0, __ldg(xc + c.x));
        t = __dadd_rn(t, __dmul_rn(v1, __ldg(xc + c.y)));
        t = __dadd_rn(t, __dmul_rn(v2, __ldg(xc + c.z)));
        t = __dadd_rn(t, __dmul_rn(v3, __ldg(xc + c.w)));
        Y[r + (int64_t)col * ldy] = t;
    }
}

// column-major (ld) <-> row-major (bw) through a shared-memory tile
__global__ void __launch_bounds__(256) k_cm_to_rm(int64_t n, int bw, const double *__restrict__ src, int64_t ld, double *__restrict__ dst)
{
    __shared__ double t[32][33];
    const int64_t base = (int64_t)blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * bw; e += 256) {
        const int r = e % 32, c = e / 32;
        t[r][c] = (base + r < n) ? src[base + r + (int64_t)c * ld] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * bw; e += 256) {
        const int r = e / bw, c = e % bw;
        if (base + r < n) dst[(base + r) * bw + c] = t[r][c];
    }
}

#include "lz_spmm_xs.cuh"

// can W = A X - Q0 B run as ONE pass on this operator?  (staged kernel with the DMMA subtraction: 16 columns)
static bool spmm_can_fuse(const lz_ctx *ctx, const lz_matrix *A, int bw)
{
    return bw == 16 && A->rowptr && A->tma_ok && A->mm_chunk_row && !A->vrowptr && ctx->spmv_variant != 9 && !ctx->knobs.no_spmm_fuse;
}

// part: 0 all rows, 1 interior chunks of a shard, 2 its boundary chunks (staged kernel only)
static int spmm_rm_rows(lz_ctx *ctx, const lz_matrix *A, const int32_t *rowptr, int64_t n, int bw, const double *X, double *W,
                       const double *Q0, const double *Bm, int part = 0, const double *Xown = nullptr, double *gpart = nullptr,
                       int *grid_out = nullptr)
{
    const bool fuse = Q0 != nullptr;
    lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 4.0 * (double)n + 16.0 * (double)n * bw + (fuse ? 8.0 * (double)n * bw : 0.0));
    // ELL4 operators carry a CSR shadow with a chunk schedule (lz_ell_create), so the block path runs the same
    // staged kernel on them; the width-4 ELL kernel is only the last resort
    if (A->format == LZ_FMT_ELL4 && !A->rowptr) {
        LZ_CHECK(part == 0, LZ_ERR_INVALID, "spmm: partial launches need a chunk schedule");
        const unsigned grid = (unsigned)((n * bw + SPMM_THREADS - 1) / SPMM_THREADS);
#define ELL_CASE(B)                                                                                               \
    if (fuse) k_spmm_rm_ell4<B, true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->ell_data, A->ell_idx, X, W, Q0, Bm); \
    else k_spmm_rm_ell4<B, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->ell_data, A->ell_idx, X, W, Q0, Bm)
        if (bw == 1) { ELL_CASE(1); } else if (bw == 2) { ELL_CASE(2); } else if (bw == 4) { ELL_CASE(4); }
        else if (bw == 8) { ELL_CASE(8); } else if (bw == 16) { ELL_CASE(16); } else if (bw == 32) { ELL_CASE(32); }
        else { lz_set_error("block width %d is not supported on an ELL4 operator (use 1,2,4,8,16,32)", bw); return LZ_ERR_UNSUPPORTED; }
#undef ELL_CASE
    } else {
        int64_t want = (n + SPMM_SLAB - 1) / SPMM_SLAB;
        int64_t cap = (int64_t)ctx->sm_count * 8;
        const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
        // operators whose chunks reference a few contiguous column ranges (stencils, banded): the rows of X a chunk needs
        // are bulk-copied into shared memory and gathered from there (lz_spmm_xs.cuh)
        int xs_st = 0, xs_sb = 0, xs_xw = 0;
        if (rowptr == A->rowptr && A->rowptr && (bw == 4 || bw == 8 || bw == 16 || bw == 32) && (!fuse || (bw == 16 && !ctx->knobs.no_spmm_fuse)) &&
            (!fuse || (uintptr_t)Q0 % 32 == 0) && spmm_xs_plan(ctx, A, bw, X, W, part, &xs_st, &xs_sb, &xs_xw)) {
            if (bw == 4) LZ_TRY((launch_spmm_xs<4, false, false>(ctx, A, X, W, nullptr, nullptr, xs_st, xs_sb, xs_xw)));
            else if (bw == 8) LZ_TRY((launch_spmm_xs<8, false, false>(ctx, A, X, W, nullptr, nullptr, xs_st, xs_sb, xs_xw)));
            else if (bw == 32) LZ_TRY((launch_spmm_xs<32, false, false>(ctx, A, X, W, nullptr, nullptr, xs_st, xs_sb, xs_xw)));
            else if (fuse && gpart) LZ_TRY((launch_spmm_xs<16, true, true>(ctx, A, X, W, Q0, Bm, xs_st, xs_sb, xs_xw, Xown, gpart, grid_out)));
            else if (fuse) LZ_TRY((launch_spmm_xs<16, true, false>(ctx, A, X, W, Q0, Bm, xs_st, xs_sb, xs_xw)));
            else LZ_TRY((launch_spmm_xs<16, false, false>(ctx, A, X, W, nullptr, nullptr, xs_st, xs_sb, xs_xw)));
            LZ_LAUNCH_CHECK(ctx);
            lz_prof_end(ctx);
            return LZ_OK;
        }
        const bool ws = A->tma_ok && A->mm_chunk_row && ctx->spmv_variant != 9 && (!fuse || spmm_can_fuse(ctx, A, bw)) &&
                        ((uintptr_t)X % 32 == 0) && ((uintptr_t)W % 32 == 0) && (!fuse || (uintptr_t)Q0 % 32 == 0);
        if (ws && (bw == 8 || bw == 16 || bw == 32)) {
            if (bw == 8) LZ_TRY(launch_spmm_ws<8>(ctx, A, rowptr, n, X, W, Q0, Bm, part));
            else if (bw == 16) LZ_TRY(launch_spmm_ws<16>(ctx, A, rowptr, n, X, W, Q0, Bm, part, Xown, gpart, grid_out));
            else LZ_TRY(launch_spmm_ws<32>(ctx, A, rowptr, n, X, W, Q0, Bm, part));
            LZ_LAUNCH_CHECK(ctx);
            lz_prof_end(ctx);
            return LZ_OK;
        }
        LZ_CHECK(part == 0, LZ_ERR_INVALID, "spmm: partial launches need the staged kernel");
#define CSR_CASE(B)                                                                                                     \
    if (fuse) k_spmm_rm<B, true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, rowptr, A->mm_k_colidx, A->mm_k_vals, X, W, Q0, Bm); \
    else k_spmm_rm<B, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, rowptr, A->mm_k_colidx, A->mm_k_vals, X, W, Q0, Bm)
        if (bw == 4) { CSR_CASE(4); } else if (bw == 8) { CSR_CASE(8); } else if (bw == 16) { CSR_CASE(16); }
        else if (bw == 32) { CSR_CASE(32); }
        else {
            if (fuse) k_spmm_rm_any<true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(bw, n, rowptr, A->mm_k_colidx, A->mm_k_vals, X, W, Q0, Bm);
            else k_spmm_rm_any<false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(bw, n, rowptr, A->mm_k_colidx, A->mm_k_vals, X, W, Q0, Bm);
        }
#undef CSR_CASE
    }
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

// W[r,:] = sum of the partial rows of r's virtual pieces (row-split operators)
__global__ void __launch_bounds__(256)
k_split_combine_rows(int64_t n_rows, int bw, const int32_t *__restrict__ vstart, const int32_t *__restrict__ vpos,
                     const double *__restrict__ Wbar, double *__restrict__ W)
{
    const int64_t total = n_rows * bw, stride = (int64_t)gridDim.x * 256;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
        const int64_t r = e / bw;
        const int c = (int)(e - r * bw);
        const int v0 = vstart[r], v1 = vstart[r + 1];
        double t = Wbar[(int64_t)(vpos ? vpos[v0] : v0) * bw + c];
        for (int v = v0 + 1; v < v1; ++v) t += Wbar[(int64_t)(vpos ? vpos[v] : v) * bw + c];
        W[e] = t;
    }
}

// the same for the panel widths the block drivers use: BW threads per row (no division per element), four rows per thread
// in flight -- the chain vstart -> vpos -> Wbar is three dependent loads, and one row at a time left the kernel waiting
// on them (R-MAT scale 24, b = 32: the combine cost as much as half of the product it follows)
template <int BW>
__global__ void __launch_bounds__(256)
k_split_combine_rows_t(int64_t n_rows, const int32_t *__restrict__ vstart, const int32_t *__restrict__ vpos,
                       const double *__restrict__ Wbar, double *__restrict__ W, const int skip_single)
{
    constexpr int RPB = 256 / BW, U = 4;
    const int c = threadIdx.x % BW, sub = threadIdx.x / BW;
    for (int64_t rb = (int64_t)blockIdx.x * (RPB * U); rb < n_rows; rb += (int64_t)gridDim.x * (RPB * U)) {
        int v0[U], v1[U];
        int64_t pos[U];
        double t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = rb + u * RPB + sub;
            v0[u] = v1[u] = 0;
            if (r < n_rows) { v0[u] = vstart[r]; v1[u] = vstart[r + 1]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v1[u] - v0[u] > LZ_LONG_PIECES) v1[u] = v0[u];             // hub rows: k_split_combine_long
            if (skip_single && v1[u] - v0[u] == 1) v1[u] = v0[u];          // uncut rows: the product wrote them itself
            pos[u] = v1[u] > v0[u] ? (vpos ? (int64_t)vpos[v0[u]] : (int64_t)v0[u]) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = pos[u] >= 0 ? __ldcs(Wbar + pos[u] * BW + c) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            for (int v = v0[u] + 1; v < v1[u]; v += 4) {                    // pieces in row order, four loads in flight
                double x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = v + j < v1[u] ? __ldcs(Wbar + (int64_t)(vpos ? vpos[v + j] : v + j) * BW + c) : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) if (v + j < v1[u]) t[u] += x[j];
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = rb + u * RPB + sub;
            if (r < n_rows && pos[u] >= 0) W[r * BW + c] = t[u];
        }
    }
}

// hub rows: one CTA per row, warp w adds the pieces v0 + w, v0 + w + 8, ... (four loads in flight), the eight partial rows
// are added in warp order -- a fixed order, so the result does not depend on the launch
__global__ void __launch_bounds__(256)
k_split_combine_long(int bw, const int32_t *__restrict__ long_rows, const int32_t *__restrict__ vstart, const int32_t *__restrict__ vpos,
                     const double *__restrict__ Wbar, double *__restrict__ W)
{
    __shared__ double part[8][32];
    const int64_t r = long_rows[blockIdx.x];
    const int v0 = vstart[r], v1 = vstart[r + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double t = 0.0;
    if (lane < bw) {
        for (int v = v0 + w; v < v1; v += 32) {
            double x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = v + 8 * j < v1 ? __ldcs(Wbar + (int64_t)(vpos ? vpos[v + 8 * j] : v + 8 * j) * bw + lane) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) t += x[j];
        }
    }
    part[w][lane] = t;
    __syncthreads();
    if (w == 0 && lane < bw) {
        double s = part[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += part[k][lane];
        W[r * bw + lane] = s;
    }
}

static int spmm_rm(lz_ctx *ctx, const lz_matrix *A, int bw, const double *X, double *W, const double *Q0, const double *Bm, int part = 0)
{
    if (!A->vrowptr) return spmm_rm_rows(ctx, A, A->rowptr, A->n_rows, bw, X, W, Q0, Bm, part);
    LZ_CHECK(part == 0, LZ_ERR_INVALID, "row-split operators cannot be launched in parts");
    LZ_TRY(lz_matrix_prepare_mm(ctx, A));
    // row-split operator: partial rows per virtual row, then an ordered combine
    LZ_CHECK(Q0 == nullptr, LZ_ERR_UNSUPPORTED, "fused subtraction is not available on a row-split operator");
    void *wbar;
    int direct = 0;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)A->mm.n_virtual * bw + 64, &wbar));
    // Power-law operator, wide panel: the gathered rows are random.  Running the panel in 8-column slices (64-byte row
    // pieces, four times as many hub rows resident in L2, the matrix streamed once per slice) was measured: every slice
    // costs as much as the whole 32-column pass (R-MAT scale 24: 4 x 54 ms instead of 55 ms, profiles/r02_rmat.md) -- the
    // product is bound by the RATE of random row gathers (~10 G/s), not by their bytes.  Off unless LZ_SPMM_SLICE=8.
    const int slice = ctx->knobs.spmm_slice > 0 ? ctx->knobs.spmm_slice : bw;
    if (bw > slice && bw % slice == 0 && slice == 8 && A->tma_ok && A->mm_chunk_row && ctx->spmv_variant != 9 &&
        ((uintptr_t)X % 32 == 0) && ((uintptr_t)wbar % 32 == 0)) {
        lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 4.0 * (double)A->mm.n_virtual + 16.0 * (double)A->mm.n_virtual * bw);
        for (int c0 = 0; c0 < bw; c0 += slice) {
            LZ_TRY((launch_spmm_ws_shape<8, 12, 2, 2, false>(ctx, A, A->mm.vrowptr, A->mm.n_virtual, X + c0, (double *)wbar + c0, nullptr, nullptr,
                                                             ctx->knobs.spmm_run, 0, bw, bw)));
            LZ_LAUNCH_CHECK(ctx);
        }
        lz_prof_end(ctx);
    } else if ((bw == 8 || bw == 16 || bw == 32) && A->mm.dst && A->tma_ok && A->mm_chunk_row && ctx->spmv_variant != 9 && !ctx->knobs.no_direct_rows &&
               ((uintptr_t)X % 32 == 0) && ((uintptr_t)wbar % 32 == 0) && ((uintptr_t)W % 32 == 0)) {
        // rows that were not cut (three quarters of an R-MAT operator's) go straight to W: no partial row, no combine
        lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 8.0 * (double)A->mm.n_virtual + 16.0 * (double)A->mm.n_virtual * bw);
        const int run = ctx->knobs.spmm_run;
        if (bw == 8) LZ_TRY((launch_spmm_ws_shape<8, 12, 2, 2, false, false, 4, true>(ctx, A, A->mm.vrowptr, A->mm.n_virtual, X, (double *)wbar, nullptr, nullptr, run, 0, 8, 8, nullptr, nullptr, nullptr, A->mm.dst, W)));
        else if (bw == 16) LZ_TRY((launch_spmm_ws_shape<16, 12, 2, 2, false, false, 4, true>(ctx, A, A->mm.vrowptr, A->mm.n_virtual, X, (double *)wbar, nullptr, nullptr, run, 0, 16, 16, nullptr, nullptr, nullptr, A->mm.dst, W)));
        else LZ_TRY((launch_spmm_ws_shape<32, 12, 2, 2, false, false, 4, true>(ctx, A, A->mm.vrowptr, A->mm.n_virtual, X, (double *)wbar, nullptr, nullptr, run, 0, 32, 32, nullptr, nullptr, nullptr, A->mm.dst, W)));
        LZ_LAUNCH_CHECK(ctx);
        lz_prof_end(ctx);
        direct = 1;
    } else
    LZ_TRY(spmm_rm_rows(ctx, A, A->mm.vrowptr, A->mm.n_virtual, bw, X, (double *)wbar, nullptr, nullptr));
    // ordered combine of the partial rows (counted in the SpMM class: it is part of the product on such operators)
    lz_prof_begin(ctx, LZ_K_SPMM, 8.0 * ((double)A->mm.n_virtual + (double)A->n_rows) * bw);
    const unsigned cg = (unsigned)ctx->sm_count * 8;
    if (bw == 32) k_split_combine_rows_t<32><<<cg, 256, 0, ctx->stream>>>(A->n_rows, A->mm.vstart, A->mm.vpos, (const double *)wbar, W, direct);
    else if (bw == 16) k_split_combine_rows_t<16><<<cg, 256, 0, ctx->stream>>>(A->n_rows, A->mm.vstart, A->mm.vpos, (const double *)wbar, W, direct);
    else if (bw == 8) k_split_combine_rows_t<8><<<cg, 256, 0, ctx->stream>>>(A->n_rows, A->mm.vstart, A->mm.vpos, (const double *)wbar, W, direct);
    else if (bw == 4) k_split_combine_rows_t<4><<<cg, 256, 0, ctx->stream>>>(A->n_rows, A->mm.vstart, A->mm.vpos, (const double *)wbar, W, direct);
    else k_split_combine_rows<<<cg, 256, 0, ctx->stream>>>(A->n_rows, bw, A->mm.vstart, A->mm.vpos, (const double *)wbar, W);
    LZ_LAUNCH_CHECK(ctx);
    if (A->mm.n_long > 0 && (bw == 32 || bw == 16 || bw == 8 || bw == 4)) {
        k_split_combine_long<<<A->mm.n_long, 256, 0, ctx->stream>>>(bw, A->mm.long_rows, A->mm.vstart, A->mm.vpos, (const double *)wbar, W);
        LZ_LAUNCH_CHECK(ctx);
    }
    lz_prof_end(ctx);
    return LZ_OK;
}

// W = A X - Q0 Bm  and  G = sym(Xown^T W)  in ONE pass over the operator (b = 16 on a schedule-carrying CSR operator):
// the SpMM kernel accumulates the per-CTA Gram partials in its epilogue, k_gram_reduce sums and symmetrises them
static int spmm_fused_gram(lz_ctx *ctx, const lz_matrix *A, const double *X, double *W, const double *Q0, const double *Bm,
                           const double *Xown, double *G)
{
    void *gp;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)ctx->sm_count * 2 * 256, &gp));
    int grid = 0;
    LZ_TRY(spmm_rm_rows(ctx, A, A->rowptr, A->n_rows, 16, X, W, Q0, Bm, 0, Xown, (double *)gp, &grid));
    LZ_CHECK(grid > 0, LZ_ERR_INVALID, "spmm_fused_gram: the fused kernel did not run");
    k_gram_reduce<<<(256 + 7) / 8, 256, 0, ctx->stream>>>(16, grid, (const double *)gp, 256, G, 1);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// basis block j (row-major) -> stored basis; block CGS sweep over the stored blocks
__global__ void k_flag_init(int *flags) { flags[0] = 0x7fffffff; flags[2] = 0x7fffffff; }

extern "C" {

__global__ void __launch_bounds__(256) k_euler_panel(int64_t count, double dt, double *__restrict__ U, const double *__restrict__ W)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < count) U[i] = fma(dt, W[i], U[i]);
}

// Block fdtd validator (methods/fdtd.hpp:33-56): U <- U + dt * A U, nsteps times; result = row lc of U.
// Panels are kept row-major like everywhere inside the block path.
int lz_fdtd_block(lz_ctx *ctx, const lz_matrix *A, const double *U0, int64_t ldu, int bw, int64_t nsteps, double t_end,
                  int64_t lc, double *result_host)
{
    LZ_CHECK(ctx && A && U0 && result_host && nsteps >= 1 && bw >= 1 && bw <= 32, LZ_ERR_INVALID, "lz_fdtd_block: bad arguments");
    const int64_t n = A->n_rows;
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_fdtd_block: the operator belongs to another (or a destroyed) context");
    LZ_CHECK(A->n_cols == n && A->halo_lo == 0 && A->halo_hi == 0, LZ_ERR_INVALID, "lz_fdtd_block: operator must be square and unsharded");
    LZ_CHECK(ldu >= n && lc >= 0 && lc < n, LZ_ERR_INVALID, "lz_fdtd_block: bad leading dimension or lc");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const size_t panel = (size_t)((n * bw + 15) / 16 * 16);
    void *work;
    LZ_TRY(lz_ctx_workspace(ctx, sizeof(double) * panel * 2, &work));
    double *U = (double *)work, *W = U + panel;
    k_cm_to_rm<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, bw, U0, ldu, U);
    LZ_LAUNCH_CHECK(ctx);
    const double dt = t_end / (double)nsteps;
    const int64_t count = n * bw;
    for (int64_t i = 0; i < nsteps; ++i) {
        LZ_TRY(spmm_rm(ctx, A, bw, U, W, nullptr, nullptr));
        k_euler_panel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(count, dt, U, W);
        LZ_LAUNCH_CHECK(ctx);
    }
    LZ_CUDA(cudaMemcpyAsync(result_host, U + lc * bw, sizeof(double) * bw, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_spmm(lz_ctx *ctx, const lz_matrix *A, int b, const double *X, int64_t ldx, double *Y, int64_t ldy)
{
    LZ_CHECK(ctx && A && X && Y && b >= 1, LZ_ERR_INVALID, "lz_spmm: bad arguments");
    LZ_CHECK(ldx >= A->n_cols && ldy >= A->n_rows, LZ_ERR_INVALID, "lz_spmm: leading dimensions too small");
    LZ_CHECK(X != Y, LZ_ERR_INVALID, "lz_spmm: X and Y must not alias");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_spmm: the operator belongs to another (or a destroyed) context");
    const unsigned grid = (unsigned)((A->n_rows + SPMM_THREADS - 1) / SPMM_THREADS);
    lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 4.0 * (double)A->n_rows + 16.0 * (double)A->n_rows * b);
    if (A->format == LZ_FMT_ELL4) k_spmm_cm_ell4<<<grid, SPMM_THREADS, 0, ctx->stream>>>(A->n_rows, b, A->ell_data, A->ell_idx, X, ldx, Y, ldy);
    else k_spmm_cm<2><<<grid, SPMM_THREADS, 0, ctx->stream>>>(A->n_rows, b, A->rowptr, A->colidx, A->vals, X, ldx, Y, ldy);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

// copy of a b x b block (device -> device), used where the reference copies small matrices
__global__ void k_copy_small(int count, const double *__restrict__ src, double *__restrict__ dst)
{
    for (int e = threadIdx.x; e < count; e += blockDim.x) dst[e] = src[e];
}

enum { SB_GLAST = 4096, SB_BLAST = 4096 + 1024, SB_G1 = 4096 + 2048, SB_G2 = 4096 + 3072 };   // slots in ctx->scalars (b*b <= 1024 each)

int lz_block_lanczos(lz_ctx *ctx, const lz_matrix *A, const double *B, int64_t ldb, int bw, int m, int64_t lc,
                     int reorth, double *alpha, double *beta, double *q)
{
    LZ_CHECK(ctx && A && B && alpha && beta && m >= 1, LZ_ERR_INVALID, "lz_block_lanczos: bad arguments");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_block_lanczos: the operator belongs to another (or a destroyed) context");
    LZ_CHECK(bw >= 1 && bw <= 32, LZ_ERR_INVALID, "lz_block_lanczos: block width %d outside 1..32", bw);
    const int64_t n = A->n_rows;
    // with a communicator attached (lz_comm_init) A is this rank's row slab, B holds the local rows, the
    // panels carry [lower halo | local | upper halo] rows and every Gram matrix is all-reduced
    const bool sharded = ctx->comm != nullptr && lz_comm_world(ctx) > 1;
    const int64_t hlo = A->halo_lo, hhi = A->halo_hi;
    LZ_CHECK(A->n_cols == n + hlo + hhi && ldb >= n, LZ_ERR_INVALID, "lz_block_lanczos: operator must be square and ldb >= n");
    LZ_CHECK(sharded || (hlo == 0 && hhi == 0), LZ_ERR_INVALID, "lz_block_lanczos: a sharded operator needs lz_comm_init");
    LZ_CHECK(lc >= -1 && lc < n, LZ_ERR_INVALID, "lz_block_lanczos: lc out of range");
    LZ_CHECK(reorth == LZ_REORTH_NONE || reorth == LZ_REORTH_FULL || reorth == LZ_REORTH_FULL_DGKS, LZ_ERR_INVALID, "lz_block_lanczos: reorth mode %d", reorth);
    const bool dgks = reorth == LZ_REORTH_FULL_DGKS;
    LZ_CHECK(!dgks || bw == 8 || bw == 16 || bw == 32, LZ_ERR_UNSUPPORTED, "lz_block_lanczos: the DGKS mode needs a block width of 8, 16 or 32");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const size_t pan = (size_t)n * bw, bb = (size_t)bw * bw;
    // panel layout [lower halo | local | upper halo]; sharded: identical offsets on every rank (peer halo pushes)
    int64_t off_rows = hlo, span_rows = hlo + n + hhi, n_below = -1;
    if (sharded) {
        const int64_t mine[4] = {hlo, n, hhi, 0};
        int64_t all[4 * LZ_MAX_RANKS];
        LZ_TRY(lz_comm_gather4(ctx, mine, all));
        const int world = lz_comm_world(ctx), rank = lz_comm_rank(ctx);
        off_rows = 0; span_rows = 0;
        for (int r = 0; r < world; ++r) off_rows = all[4 * r] > off_rows ? all[4 * r] : off_rows;
        for (int r = 0; r < world; ++r) { const int64_t sp = off_rows + all[4 * r + 1] + all[4 * r + 2]; span_rows = sp > span_rows ? sp : span_rows; }
        n_below = rank > 0 ? all[4 * (rank - 1) + 1] * bw : 0;
    }
    const size_t pstride = ((size_t)span_rows * bw + (size_t)(ctx->knobs.panel_pad > 0 ? ctx->knobs.panel_pad : 0) + 15) & ~(size_t)15;   // one panel incl. halo rows, 128-byte multiple
    // Unsharded full reorthogonalisation keeps Q_j only inside the stored basis (block j at V + j*pan): the
    // normalisation writes it there and the SpMM gathers from there -- no per-step device-to-device copy.
    const bool inplace = reorth && !sharded;
    const size_t n_panels = inplace ? 1 : 3;
    const size_t work_doubles = n_panels * pstride + (reorth ? 2 * bb * m : 0) + 64;
    void *work;
    if (sharded) LZ_TRY(lz_comm_arena(ctx, sizeof(double) * work_doubles, (size_t)(reorth ? m : 1) * bb + 8, &work));
    else LZ_TRY(lz_ctx_workspace(ctx, sizeof(double) * work_doubles, &work));
    double *W = (double *)work + (size_t)off_rows * bw;                     // W never needs halo rows; same offset keeps alignment
    double *Qa = inplace ? nullptr : W + pstride, *Qb = inplace ? nullptr : Qa + pstride;
    double *C = (double *)work + n_panels * pstride;
    double *V = nullptr;
    if (reorth) LZ_TRY(lz_ctx_basis_blocks(ctx, (int64_t)pan, m, &V));     // block j at V + j*pan, row-major
    double *binv = beta + bb * m;                                         // beta[m]: scratch inverse (block_lanczos.hpp:111)
    int *flag = ctx->flags + 2;
    k_flag_init<<<1, 1, 0, ctx->stream>>>(ctx->flags);
    LZ_LAUNCH_CHECK(ctx);
    auto reduce_small = [&](double *M) -> int { return sharded ? lz_comm_allreduce_sum(ctx, M, bb) : LZ_OK; };
    auto halo = [&](double *Q) -> int {
        return sharded ? lz_comm_halo_exchange(ctx, Q, (int64_t)pan, hlo * bw, hhi * bw, n_below, false) : LZ_OK;
    };
    int *sweep_flag = ctx->flags + 6;
    // Full reorthogonalisation of W against the stored blocks and the next W^T W into `gn`.
    //   FULL: two block CGS sweeps, always (CGS2), then the Gram.
    //   DGKS: one sweep, the Gram, and a second sweep + Gram only if some column of W lost more than half of its squared
    //         norm in the first ("twice is enough", decided on the device from the two Gram diagonals; `gb` = W^T W before)
    auto reorthogonalise = [&](int nblocks, double *gb, double *gn) -> int {
        LZ_TRY(lz_block_cgs(ctx, n, bw, nblocks, V, (int64_t)pan, W, C, sharded));
        if (!dgks) {
            LZ_TRY(lz_block_cgs(ctx, n, bw, nblocks, V, (int64_t)pan, W, C, sharded));
            return lz_gram(ctx, n, bw, true, W, 0, W, 0, gn, 0);
        }
        double *ga = sharded ? ctx->scalars + SB_G2 : gn;                 // Gram after the first sweep
        LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, ga, 0));
        if (sharded) LZ_TRY(lz_comm_allreduce_sum(ctx, ga, bb));
        LZ_TRY(lz_block_dgks_test(ctx, bw, gb, ga, sweep_flag));
        LZ_TRY(lz_block_cgs(ctx, n, bw, nblocks, V, (int64_t)pan, W, C, sharded, sweep_flag));
        if (sharded) return lz_gram(ctx, n, bw, true, W, 0, W, 0, gn, 0);   // (its all-reduce follows unconditionally)
        return lz_gram_if(ctx, n, bw, W, gn, sweep_flag);
    };
    auto qslot = [&](int j) -> double * { return inplace ? V + pan * j : ((j & 1) ? Qb : Qa); };
    const bool qrow = q != nullptr && lc >= 0;
    const bool fuse = spmm_can_fuse(ctx, A, bw);                           // W = A Q_j - Q_{j-1} beta_j in one pass
    double *sc = ctx->scalars;
    // where the next step's W^T W lands: straight in its beta slot (the square root is taken in place); the one
    // after the last step goes to the context (lz_block_last_coupling) so that beta[m] keeps the last inverse
    auto gslot = [&](int jn) -> double * { return jn < m ? beta + bb * jn : sc + SB_GLAST; };

    // W <- B in row-major; beta[0] = (B^T B)^{1/2}; Q0 = B beta[0]^{-1}                     (:106-114)
    k_cm_to_rm<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, bw, B, ldb, W);
    LZ_LAUNCH_CHECK(ctx);
    LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, beta, 0));
    LZ_TRY(reduce_small(beta));
    LZ_TRY(lz_sqrtm_launch(ctx, bw, beta, binv, flag, 0));
    double *Q0 = qslot(0);
    LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, binv, 0.0, 1.0, Q0, 0, nullptr));
    if (qrow) LZ_TRY(lz_copy_row_launch(ctx, lc, bw, true, Q0, 0, q, 0));                     // :117
    if (V && !inplace) LZ_CUDA(cudaMemcpyAsync(V, Q0, sizeof(double) * pan, cudaMemcpyDeviceToDevice, ctx->stream));
    LZ_TRY(halo(Q0));
    LZ_TRY(spmm_rm(ctx, A, bw, Q0 - hlo * bw, W, nullptr, nullptr));                          // :121
    LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, Q0, 0, alpha, 1));                                 // :124
    LZ_TRY(reduce_small(alpha));
    double *gbefore = sc + SB_G1;                          // DGKS: W^T W before the sweeps (fused into the panel pass)
    LZ_TRY(lz_panel(ctx, n, bw, true, Q0, 0, alpha, 1.0, -1.0, W, 0, reorth ? (dgks ? gbefore : nullptr) : gslot(1)));  // :128 (+ :137 fused)
    if (reorth) {
        if (dgks) LZ_TRY(reduce_small(gbefore));
        LZ_TRY(reorthogonalise(1, gbefore, gslot(1)));
    }
    LZ_TRY(reduce_small(gslot(1)));
    for (int j = 1; j < m; ++j) {                                                             // :132-166
        double *bj = beta + bb * j, *aj = alpha + bb * j;                                     // bj holds W^T W (:137)
        LZ_TRY(lz_sqrtm_launch(ctx, bw, bj, binv, flag, j));                                  // :142
        double *Q1 = qslot(j);
        LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, binv, 0.0, 1.0, Q1, 0, nullptr));             // :145
        if (qrow) LZ_TRY(lz_copy_row_launch(ctx, lc, bw, true, Q1, 0, q, (int64_t)j * bw));   // :165
        if (V && !inplace) LZ_CUDA(cudaMemcpyAsync(V + pan * j, Q1, sizeof(double) * pan, cudaMemcpyDeviceToDevice, ctx->stream));
        LZ_TRY(halo(Q1));
        double *gn = gslot(j + 1);
        if (fuse) {
            // the reference's order: W = A Q_j - Q_{j-1} beta_j (:149,:152), alpha_j = sym(W^T Q_j) (:155),
            // W -= Q_j alpha_j (:159) with the next W^T W accumulated in the same pass (:137)
            if (!ctx->knobs.no_spmm_gram) {
                LZ_TRY(spmm_fused_gram(ctx, A, Q1 - hlo * bw, W, Q0, bj, Q1, aj));        // ... and Q_j^T W in the same pass
            } else {
                LZ_TRY(spmm_rm(ctx, A, bw, Q1 - hlo * bw, W, Q0, bj));
                LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, Q1, 0, aj, 1));
            }
            LZ_TRY(reduce_small(aj));
            LZ_TRY(lz_panel(ctx, n, bw, true, Q1, 0, aj, 1.0, -1.0, W, 0, reorth ? (dgks ? gbefore : nullptr) : gn));
        } else {
            // same quantities when the SpMM cannot subtract: G1 = Q_j^T (A Q_j) and G2 = Q_j^T Q_{j-1} from one pass,
            // alpha_j = sym(G1 - G2 beta_j), then one pass subtracts both Q_{j-1} beta_j and Q_j alpha_j (+ next W^T W)
            LZ_TRY(spmm_rm(ctx, A, bw, Q1 - hlo * bw, W, nullptr, nullptr));
            LZ_TRY(lz_gram2(ctx, n, bw, Q1, W, Q0, sc + SB_G1, sc + SB_G2));
            if (sharded) LZ_TRY(lz_comm_allreduce_sum(ctx, sc + SB_G1, 2048));            // G1 and G2 sit back to back
            LZ_TRY(lz_alpha_from_grams(ctx, bw, sc + SB_G1, sc + SB_G2, bj, aj));
            LZ_TRY(lz_panel2(ctx, n, bw, Q0, bj, Q1, aj, W, reorth ? (dgks ? gbefore : nullptr) : gn));
        }
        Q0 = Q1;                                                                              // :162 (no copy)
        if (V) {
            if (dgks) LZ_TRY(reduce_small(gbefore));
            LZ_TRY(reorthogonalise(j + 1, gbefore, gn));
        }
        LZ_TRY(reduce_small(gn));
    }
    // beta_m = (W_m^T W_m)^{1/2}: the coupling to the next (unbuilt) block, for residual estimates and restarts
    k_copy_small<<<1, 256, 0, ctx->stream>>>((int)bb, sc + SB_GLAST, sc + SB_BLAST);
    LZ_LAUNCH_CHECK(ctx);
    LZ_TRY(lz_sqrtm_launch(ctx, bw, sc + SB_BLAST, sc + SB_G1, ctx->flags + 3, 0));
    ctx->last_coupling_slot = SB_BLAST;
    return LZ_OK;
}

// Pre-sizes everything lz_block_lanczos would allocate on its first call for this operator and these sizes (three
// panels or the basis slab, coefficient blocks, reduction scratch) and loads the kernels' modules, so that a timed
// first call measures the iteration and not cudaMalloc / lazy module loading ("scratch is (re)sized outside timed
// regions by the *_workspace calls", include/lanczos_b200.h).  Unsharded contexts only; sharded runs size their
// arena collectively inside the driver.
int lz_block_lanczos_workspace(lz_ctx *ctx, const lz_matrix *A, int bw, int m, int reorth)
{
    LZ_CHECK(ctx && A && bw >= 1 && bw <= 32 && m >= 1, LZ_ERR_INVALID, "lz_block_lanczos_workspace: bad arguments");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_block_lanczos_workspace: the operator belongs to another (or a destroyed) context");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = A->n_rows;
    const size_t bb = (size_t)bw * bw, pan = (size_t)n * bw;
    const size_t pstride = ((size_t)(A->halo_lo + n + A->halo_hi) * bw + (size_t)(ctx->knobs.panel_pad > 0 ? ctx->knobs.panel_pad : 0) + 15) & ~(size_t)15;
    const size_t n_panels = reorth ? 1 : 3;
    void *p;
    LZ_TRY(lz_ctx_workspace(ctx, sizeof(double) * (n_panels * pstride + (reorth ? 2 * bb * m : 0) + 64), &p));
    double *V;
    if (reorth) LZ_TRY(lz_ctx_basis_blocks(ctx, (int64_t)pan, m, &V));
    // reduction scratch: Gram partials (two products) and, with reorthogonalisation, the projection partials of m blocks
    size_t scratch = (size_t)ctx->sm_count * 4 * 2 * bb;
    if (reorth) scratch = std::max(scratch, (size_t)((m + 1) / 2 + 1) * ctx->sm_count * 2 * 8 * bb + (size_t)m * bb);
    if (A->vrowptr) {
        LZ_TRY(lz_matrix_prepare_mm(ctx, A));
        scratch = std::max(scratch, (size_t)A->mm.n_virtual * bw + 64);
    }
    if (bw == 16 || ctx->knobs.xs_force) LZ_TRY(lz_matrix_prepare_xs(ctx, A));      // the staged SpMM's window schedule (built once per operator)
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * scratch, &p));
    return LZ_OK;
}

int lz_matrix_spmm_schedule(lz_ctx *ctx, const lz_matrix *A, int bw, int *kind, int box[3], int *window_rows)
{
    LZ_CHECK(ctx && A && kind && bw >= 1 && bw <= 32, LZ_ERR_INVALID, "lz_matrix_spmm_schedule: bad arguments");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_matrix_spmm_schedule: the operator belongs to another (or a destroyed) context");
    LZ_CUDA(cudaSetDevice(ctx->device));
    int st = 0, sb = 0, xw = 0;
    // (aligned dummy panel pointers: the plan only looks at their alignment)
    const bool xs = A->rowptr && (bw == 4 || bw == 8 || bw == 16 || bw == 32) &&
                    spmm_xs_plan(ctx, A, bw, (const double *)(uintptr_t)256, (const double *)(uintptr_t)256, 0, &st, &sb, &xw);
    *kind = xs ? (A->xs_tile_dims[0] > 0 ? 2 : 1) : 0;
    if (box) { box[0] = xs ? A->xs_tile_dims[0] : 0; box[1] = xs ? A->xs_tile_dims[1] : 0; box[2] = xs ? A->xs_tile_dims[2] : 0; }
    if (window_rows) *window_rows = xs ? A->xs_max_wrows : 0;
    return LZ_OK;
}

// Number of valid blocks of the last lz_block_lanczos run on this context: m when every W^T W was positive
// definite to working precision, otherwise the index j of the first beta_j that was singular or not finite
// (alpha[0..j), beta[0..j] are valid, LZ_ERR_BREAKDOWN is returned).  Synchronises.
int lz_block_status(lz_ctx *ctx, int m, int *blocks_done)
{
    LZ_CHECK(ctx && m >= 1, LZ_ERR_INVALID, "lz_block_status: bad arguments");
    int flag = 0;
    LZ_CUDA(cudaMemcpyAsync(&flag, ctx->flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    const int done = flag < m ? flag : m;
    if (blocks_done) *blocks_done = done;
    if (done < m) {
        lz_set_error("lz_block_lanczos: breakdown, beta[%d] is singular or not finite", done);
        return LZ_ERR_BREAKDOWN;
    }
    return LZ_OK;
}

// beta_m of the last run (bw = 1: the vector driver's beta_m; otherwise the bw x bw block, column-major) to HOST
int lz_last_coupling(lz_ctx *ctx, int bw, double *beta_last_host)
{
    LZ_CHECK(ctx && beta_last_host && bw >= 1 && bw <= 32, LZ_ERR_INVALID, "lz_last_coupling: bad arguments");
    const double *src = ctx->scalars + ctx->last_coupling_slot;
    LZ_CUDA(cudaMemcpyAsync(beta_last_host, src, sizeof(double) * bw * bw, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

// row-major (bw) -> column-major (ld), `cols` leading columns only
__global__ void __launch_bounds__(256) k_rm_to_cm(int64_t n, int bw, int cols, const double *__restrict__ src, double *__restrict__ dst, int64_t ld)
{
    __shared__ double t[32][33];
    const int64_t base = (int64_t)blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * bw; e += 256) {
        const int r = e / bw, c = e % bw;
        t[r][c] = (base + r < n) ? src[(base + r) * bw + c] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * cols; e += 256) {
        const int r = e % 32, c = e / 32;
        if (base + r < n) dst[base + r + (int64_t)c * ld] = t[r][c];
    }
}

// ---------------------------------------------------------------------------------------------
// Block thick-restart Lanczos (BASELINE config 3: block size 16, k = 64): k extremal eigenpairs with
// a basis of p blocks.  Degenerate / clustered eigenvalues -- the cubic Laplacian has up to six-fold ones -- need a
// block method: a single Lanczos vector sees one direction per eigenspace.  Every step is the block recurrence with
// full block CGS2 against ALL stored blocks; after p blocks the projected matrix H (block tridiagonal plus, after a
// restart, the coupling rows to the kept Ritz vectors) is diagonalised on the host, the basis is compressed to the
// wanted Ritz vectors plus a share of their neighbours (a whole number of blocks) with DMMA block combinations, and
// the normalised residual block continues the recurrence.  The reorthogonalisation removes the coupling to the kept
// vectors, so the restarted step needs no special kernel; H's entries are known in closed form.
// Unsharded operators, bw in {8, 16, 32}.
// ---------------------------------------------------------------------------------------------
int lz_block_eigs_thick_restart(lz_ctx *ctx, const lz_matrix *A, const double *B, int64_t ldb, int bw, int k, int which,
                                int p_blocks, double tol, int max_restarts, double *theta_host, double *resid_host,
                                double *X, int64_t ldx, int *info4)
{
    LZ_CHECK(ctx && A && B && theta_host && k >= 1 && which >= 0 && which <= 2 && tol > 0.0 && max_restarts >= 0, LZ_ERR_INVALID,
             "lz_block_eigs_thick_restart: bad arguments");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_block_eigs_thick_restart: the operator belongs to another (or a destroyed) context");
    LZ_CHECK(bw == 8 || bw == 16 || bw == 32, LZ_ERR_UNSUPPORTED, "lz_block_eigs_thick_restart: block width %d (use 8, 16 or 32)", bw);
    const int64_t n = A->n_rows;
    LZ_CHECK(A->n_cols == n && A->halo_lo == 0 && A->halo_hi == 0 && !(ctx->comm && lz_comm_world(ctx) > 1), LZ_ERR_UNSUPPORTED,
             "lz_block_eigs_thick_restart: unsharded square operators only");
    const int p = p_blocks, N = p * bw;
    const int kb = (k + bw - 1) / bw;                                   // blocks needed for the wanted pairs
    LZ_CHECK(p >= kb + 3 && N <= 2048 && ldb >= n && (!X || ldx >= n), LZ_ERR_INVALID,
             "lz_block_eigs_thick_restart: need p_blocks >= ceil(k / bw) + 3, p_blocks * bw <= 2048, ldb / ldx >= n");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const size_t pan = (size_t)n * bw, bb = (size_t)bw * bw;
    // basis: p blocks + p scratch blocks for the compressed basis (copied back to the front) ; one work panel W
    double *V;
    LZ_TRY(lz_ctx_basis_blocks(ctx, (int64_t)pan, 2 * p, &V));
    void *work;
    LZ_TRY(lz_ctx_workspace(ctx, sizeof(double) * (pan + 4 * bb * (size_t)(p + 2) + 64), &work));
    double *W = (double *)work, *C = W + pan;                          // C: CGS coefficients / combination blocks (2 bb p)
    double *dsm = C + 2 * bb * (size_t)(p + 1);                         // device b x b blocks: alpha, G / beta, beta^{-1}
    double *d_alpha = dsm, *d_beta = dsm + bb, *d_binv = dsm + 2 * bb, *d_gb = dsm + 3 * bb;
    int *flag = ctx->flags + 2;
    k_flag_init<<<1, 1, 0, ctx->stream>>>(ctx->flags);
    LZ_LAUNCH_CHECK(ctx);
    auto block_at = [&](int j) { return V + pan * (size_t)j; };
    std::vector<double> H((size_t)N * N, 0.0), Hw, d, Zt, hb(bb), beta_p(bb);
    auto put_block = [&](int bi, int bj, const double *blk /* col-major b x b */, bool transpose) {
        for (int c = 0; c < bw; ++c)
            for (int r = 0; r < bw; ++r)
                H[(size_t)(bi * bw + r) + (size_t)(bj * bw + c) * N] = transpose ? blk[c + r * bw] : blk[r + c * bw];
    };
    // Q_0 = B (B^T B)^{-1/2}
    k_cm_to_rm<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, bw, B, ldb, W);
    LZ_LAUNCH_CHECK(ctx);
    LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, d_beta, 0));
    LZ_TRY(lz_sqrtm_launch(ctx, bw, d_beta, d_binv, flag, 0));
    LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, d_binv, 0.0, 1.0, block_at(0), 0, nullptr));
    int j0 = 0, restarts = 0, matvecs = 0, nconv = 0;
    std::vector<int> order(N), wanted, keep;
    for (;;) {
        for (int j = j0; j < p; ++j) {
            double *Qj = block_at(j);
            LZ_TRY(spmm_rm(ctx, A, bw, Qj, W, nullptr, nullptr));
            LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, Qj, 0, d_alpha, 1));
            LZ_CUDA(cudaMemcpyAsync(hb.data(), d_alpha, sizeof(double) * bb, cudaMemcpyDeviceToHost, ctx->stream));
            LZ_TRY(lz_panel(ctx, n, bw, true, Qj, 0, d_alpha, 1.0, -1.0, W, 0, d_gb));
            // block Gram-Schmidt against every stored block: removes Q_{j-1} beta_j, the coupling to the kept Ritz vectors
            // and whatever rounding left elsewhere; the second sweep runs only when the first removed more than half of
            // some column (it always does right after a restart, almost never otherwise)
            LZ_TRY(lz_block_cgs(ctx, n, bw, j + 1, V, (int64_t)pan, W, C, false));
            LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, d_beta, 0));
            LZ_TRY(lz_block_dgks_test(ctx, bw, d_gb, d_beta, ctx->flags + 6));
            LZ_TRY(lz_block_cgs(ctx, n, bw, j + 1, V, (int64_t)pan, W, C, false, ctx->flags + 6));
            LZ_TRY(lz_gram_if(ctx, n, bw, W, d_beta, ctx->flags + 6));
            LZ_TRY(lz_sqrtm_launch(ctx, bw, d_beta, d_binv, flag, j + 1));
            LZ_CUDA(cudaStreamSynchronize(ctx->stream));
            put_block(j, j, hb.data(), false);
            LZ_CUDA(cudaMemcpy(hb.data(), d_beta, sizeof(double) * bb, cudaMemcpyDeviceToHost));
            if (j + 1 < p) {
                put_block(j + 1, j, hb.data(), false);
                put_block(j, j + 1, hb.data(), true);
                LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, d_binv, 0.0, 1.0, block_at(j + 1), 0, nullptr));
            } else {
                beta_p = hb;                                            // coupling to the unbuilt block p
            }
        }
        matvecs += (p - j0) * bw;
        int fl = 0;
        LZ_CUDA(cudaMemcpy(&fl, flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (fl <= p) {
            lz_set_error("lz_block_eigs_thick_restart: breakdown, block %d is rank deficient (invariant subspace or dependent start block)", fl);
            return LZ_ERR_BREAKDOWN;
        }
        Hw = H;
        LZ_TRY(lz_sym_eig_full_c(N, Hw.data(), d, Zt));                 // Zt[r * N + i] = component r of eigenvector i
        for (int i = 0; i < N; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
        double scale = 0.0;
        for (int i = 0; i < N; ++i) scale = std::max(scale, fabs(d[i]));
        const int lo = which == 0 ? k : which == 1 ? 0 : k / 2, hi = k - lo;
        wanted.clear();
        for (int t = 0; t < lo; ++t) wanted.push_back(order[t]);
        for (int t = 0; t < hi; ++t) wanted.push_back(order[N - hi + t]);
        auto resid_of = [&](int idx) {          // || beta_p Y[last block rows, idx] ||
            double r2 = 0.0;
            for (int r = 0; r < bw; ++r) {
                double sacc = 0.0;
                for (int c = 0; c < bw; ++c) sacc += beta_p[r + c * bw] * Zt[(size_t)(N - bw + c) * N + idx];
                r2 += sacc * sacc;
            }
            return sqrt(r2);
        };
        nconv = 0;
        for (int idx : wanted)
            if (resid_of(idx) <= tol * scale) ++nconv;
        if (nconv == k || restarts == max_restarts) break;
        // keep the wanted pairs plus a share of their neighbours, rounded up to whole blocks
        int kk = k + std::max(bw, (N - k) * 2 / 5);
        kk = ((kk + bw - 1) / bw) * bw;
        kk = std::min(kk, (p - 2) * bw);
        const int extra = kk - k;
        const int elo = which == 0 ? extra : which == 1 ? 0 : extra / 2, ehi = extra - elo;
        keep.clear();
        for (int t = 0; t < lo + elo; ++t) keep.push_back(order[t]);
        for (int t = 0; t < hi + ehi; ++t) keep.push_back(order[N - (hi + ehi) + t]);
        const int kkb = kk / bw;
        // compressed basis block c = sum_j V_j Y[j-th block rows, kept columns of block c]  -> scratch blocks p .. p+kkb-1
        std::vector<double> negY((size_t)p * bb);
        for (int c = 0; c < kkb; ++c) {
            for (int j = 0; j < p; ++j)
                for (int s2 = 0; s2 < bw; ++s2)
                    for (int r = 0; r < bw; ++r)
                        negY[(size_t)j * bb + r + (size_t)s2 * bw] = -Zt[(size_t)(j * bw + r) * N + keep[c * bw + s2]];
            LZ_CUDA(cudaMemcpyAsync(C, negY.data(), sizeof(double) * negY.size(), cudaMemcpyHostToDevice, ctx->stream));
            LZ_TRY(lz_block_combine(ctx, n, bw, p, V, (int64_t)pan, C, block_at(p + c)));
            LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        LZ_CUDA(cudaMemcpyAsync(block_at(0), block_at(p), sizeof(double) * pan * kkb, cudaMemcpyDeviceToDevice, ctx->stream));
        // the residual block W beta_p^{-1} becomes block kkb  (d_binv still holds beta_p^{-1})
        LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, d_binv, 0.0, 1.0, block_at(kkb), 0, nullptr));
        std::fill(H.begin(), H.end(), 0.0);
        for (int c = 0; c < kk; ++c) {
            H[c + (size_t)c * N] = d[keep[c]];
            for (int r = 0; r < bw; ++r) {              // S = beta_p Y[last block rows, keep]:  H[kk + r, c] = S[r, c]
                double sacc = 0.0;
                for (int cc = 0; cc < bw; ++cc) sacc += beta_p[r + cc * bw] * Zt[(size_t)(N - bw + cc) * N + keep[c]];
                H[(size_t)(kk + r) + (size_t)c * N] = sacc;
                H[c + (size_t)(kk + r) * N] = sacc;
            }
        }
        j0 = kkb;
        ++restarts;
    }
    std::sort(wanted.begin(), wanted.end(), [&](int x, int y) { return d[x] < d[y]; });
    for (int t = 0; t < k; ++t) {
        theta_host[t] = d[wanted[t]];
        if (resid_host) {
            double r2 = 0.0;
            for (int r = 0; r < bw; ++r) {
                double sacc = 0.0;
                for (int c = 0; c < bw; ++c) sacc += beta_p[r + c * bw] * Zt[(size_t)(N - bw + c) * N + wanted[t]];
                r2 += sacc * sacc;
            }
            resid_host[t] = sqrt(r2);
        }
    }
    if (X) {
        std::vector<double> negY((size_t)p * bb);
        for (int c = 0; c < kb; ++c) {
            for (int j = 0; j < p; ++j)
                for (int s2 = 0; s2 < bw; ++s2)
                    for (int r = 0; r < bw; ++r) {
                        const int col = c * bw + s2;
                        negY[(size_t)j * bb + r + (size_t)s2 * bw] = col < k ? -Zt[(size_t)(j * bw + r) * N + wanted[col]] : 0.0;
                    }
            LZ_CUDA(cudaMemcpyAsync(C, negY.data(), sizeof(double) * negY.size(), cudaMemcpyHostToDevice, ctx->stream));
            LZ_TRY(lz_block_combine(ctx, n, bw, p, V, (int64_t)pan, C, block_at(p)));
            const int cols = std::min(bw, k - c * bw);
            k_rm_to_cm<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, bw, cols, block_at(p), X + (size_t)c * bw * ldx, ldx);
            LZ_LAUNCH_CHECK(ctx);
            LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (info4) { info4[0] = nconv; info4[1] = restarts; info4[2] = matvecs; info4[3] = N; }
    return LZ_OK;
}

}  // extern "C"
