// lz_spmm_xs.cuh -- operand-staging SpMM for structured-grid operators (the reference's spmm(),
// kernels/spmv_spmm.hpp:137-199, on its stencil matrices; BASELINE config 3).
//
// What bounds the gathering kernel (k_spmm_ws) is not the latency of its gathers but the bytes they pull from L2 to the
// SMs: a chunk of consecutive rows of a 7-point operator touches every row of X 5.1 times (ncu: 15.4 GB L2 -> SM for
// 7.7 GB of algorithmic traffic; profiles/r02_spmm.md).  Two changes together, neither alone:
//   * chunks are BOXES of grid points (lz_matrix_prepare_xs, lz_csr.cu: nested strides found from the operator itself; the
//     product walks a chunk-ordered, 8-entry-padded copy of the operator and writes row rowmap[i]), so a row of X is needed
//     by 3.1 chunks instead of 5.1;
//   * the rows a chunk needs -- its X WINDOW, <= 32 contiguous row ranges worked out once per operator by k_xs_build --
//     are bulk-copied (cp.async.bulk, one copy per segment, issued by the lanes of the producer warp) into the ring slot
//     together with the chunk's values, row pointers, row map, 16-bit window indices and its descriptor, two or three
//     chunks ahead; the compute warps gather from shared memory.  No load of the chunk loop waits for a load issued in
//     the same iteration (descriptors two chunks ahead, the Q0 rows prefetched into L2 by the producer), and the matrix
//     stream shrinks from 12 to 10 bytes per entry (the 32-bit column index is not read at all).
// One 16-warp CTA per SM.  The kernel ends up bound by the shared-memory data pipe (81 %).
//
// Bank conflicts: a lane owns four adjacent columns (32 bytes) of a row and reads them as two 128-bit loads; a quarter
// warp (one wavefront) spans two row groups, whose rows sit at arbitrary multiples of 128 bytes.  Lane groups with odd
// (lane >> 2) read their upper 16 bytes first: the eight lanes of a wavefront then cover all 32 banks.  The accumulators
// of those lanes are kept in the swapped order and put back once per row.
//
// FSUB / GRAM: the fused DMMA subtraction W = A X - Q0 B and the Gram epilogue of k_spmm_ws (same fragment labelling,
// same reduction order inside a CTA); the Q0 rows of a trip are fetched BEFORE the entry loop (L2 hits thanks to the
// prefetch), the trip's own rows of X (Gram A fragments) come from the staged window when the chunk references them.
#pragma once

#define LZ_XS_ES 1040           // entries staged per ring slot (chunk entries + alignment slack), multiple of 8
#define LZ_XS_RCAP 160          // row pointers staged per ring slot
#define LZ_XS_CW 15             // compute warps per CTA (one CTA per SM)

__device__ __forceinline__ void lz_mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = lz_smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();          // a lost copy must not hang the GPU
    }
}

template <int BW, int CW, bool FSUB, bool GRAM>
__global__ void __launch_bounds__((1 + CW) * 32, 1)
k_spmm_xs(const int n_chunks, const int64_t n_rows, const int4 *__restrict__ desc, const int2 *__restrict__ meta, const int2 *__restrict__ seg,
          const int32_t *__restrict__ rowptr, const uint16_t *__restrict__ lidx, const double *__restrict__ vals, const int32_t *__restrict__ rowmap,
          const double *__restrict__ X, double *__restrict__ W, const double *__restrict__ Q0, const double *__restrict__ Bm, const int stages,
          const int stage_bytes, const int xw_bytes, const int hint, const double *__restrict__ Xown, double *__restrict__ gpart,
          const int32_t *__restrict__ oseg, const uint16_t *__restrict__ dli)
{
    static_assert(!GRAM || FSUB, "the fused Gram rides on the 8-row trips of the fused subtraction");
    static_assert(!FSUB || BW == 16, "the fused subtraction is written for 16-column panels");
    constexpr int LW = BW / 4, RPW = 32 / LW, NG = CW * RPW, G = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // ring slot s: [ X window (xw_bytes) | values (ES*8) | window indices (ES*2) | row pointers (RCAP*4) | row map (RCAP*4) | own-row window index (RCAP*2) | chunk descriptor ]
    // rowmap: row i of the walked (chunk-ordered, padded) operator is row rowmap[i] of W / Q0 / Xown
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stages * stage_bytes);
    uint64_t *freeb = full + 8;
    __shared__ double sbs[FSUB ? 8 * 32 : 1];                             // -B in fragment order (see k_spmm_ws)
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
        for (int s = 0; s < stages; ++s) { lz_mbar_init(&full[s], 1); lz_mbar_init(&freeb[s], CW); }
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
    // chunk -> CTA map: round robin, all CTAs sweep a window of gridDim.x chunks together (the far neighbours of a stencil
    // row were touched one window earlier and are still in L2)
    auto vchunk = [&](int it) { return it * (int)gridDim.x + (int)blockIdx.x; };

    if (warp == 0) {
        // ------------------------------------------------------------------ producer warp
        // descriptors, segment tables: fetched TWO chunks ahead (nothing in an iteration waits for a load it issued itself)
        int4 d0 = make_int4(0, 0, 0, 0), d1 = d0;
        int2 m0 = make_int2(0, 0), m1 = m0, s0 = m0, s1 = m0;
        int o0 = -1, o1 = -1;               // FSUB: first original row of this lane's 8-row output group (lanes < LZ_XS_OGROUPS)
        auto fetch = [&](int c, int4 &d, int2 &m, int2 &sg, int &og) {
            if (c < n_chunks) {
                d = desc[c]; m = meta[c]; sg = seg[(size_t)c * LZ_XS_SEGCAP + lane];
                if (FSUB && lane < LZ_XS_OGROUPS) og = oseg[(size_t)c * LZ_XS_OGROUPS + lane];
            }
        };
        fetch(vchunk(0), d0, m0, s0, o0);
        fetch(vchunk(1), d1, m1, s1, o1);
        const uint64_t pol = lz_policy_evict_first();
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; vchunk(it) < n_chunks; ++it) {
            const int c = vchunk(it);
            const int4 cd = d0;
            const int2 cmt = m0, csg = s0;
            const int cog = o0;
            d0 = d1; m0 = m1; s0 = s1; o0 = o1; o1 = -1;
            fetch(vchunk(it + 2), d1, m1, s1, o1);
            unsigned char *st = smem_raw + (size_t)slot * stage_bytes;
            lz_mbar_wait_bounded(&freeb[slot], phase ^ 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const int a0 = cd.x, cnt = cd.y - cd.x;                 // multiples of 8 (the chunks are padded)
            const int ra = cd.z & ~7;
            const int rcnt = ((cd.w + 1 - ra) + 7) & ~7;
            const bool rows_ok = rcnt <= LZ_XS_RCAP && (int64_t)ra + rcnt <= n_rows + 8;     // (the row arrays carry 8 spare entries)
            // window offset of this lane's segment: exclusive prefix of the segment sizes
            const int rows = lane < cmt.x ? csg.y : 0;
            int incl = rows;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int off = incl - rows;
            if (lane == 0)
                lz_mbar_expect_tx(&full[slot], (uint32_t)cnt * 10u + (rows_ok ? (uint32_t)rcnt * (GRAM ? 10u : 8u) : 0u) + 16u + (uint32_t)cmt.y * (uint32_t)(BW * 8));
            __syncwarp();
            unsigned char *sm = st + xw_bytes;
            if (lane == 0 && cnt > 0) {
                if (hint & 1) {
                    lz_bulk_g2s_hint(sm, vals + a0, (uint32_t)cnt * 8u, &full[slot], pol);
                    lz_bulk_g2s_hint(sm + LZ_XS_ES * 8, lidx + a0, (uint32_t)cnt * 2u, &full[slot], pol);
                } else {
                    lz_bulk_g2s(sm, vals + a0, (uint32_t)cnt * 8u, &full[slot]);
                    lz_bulk_g2s(sm + LZ_XS_ES * 8, lidx + a0, (uint32_t)cnt * 2u, &full[slot]);
                }
            }
            if (lane == 1 && rows_ok) lz_bulk_g2s(sm + LZ_XS_ES * 10, rowptr + ra, (uint32_t)rcnt * 4u, &full[slot]);
            if (lane == 2 && rows_ok) lz_bulk_g2s(sm + LZ_XS_ES * 10 + LZ_XS_RCAP * 4, rowmap + ra, (uint32_t)rcnt * 4u, &full[slot]);
            if (lane == 3) lz_bulk_g2s(sm + LZ_XS_ES * 10 + LZ_XS_RCAP * 10, desc + c, 16u, &full[slot]);
            if (GRAM && lane == 4 && rows_ok) lz_bulk_g2s(sm + LZ_XS_ES * 10 + LZ_XS_RCAP * 8, dli + ra, (uint32_t)rcnt * 2u, &full[slot]);
            if (rows > 0)
                lz_bulk_g2s(st + (size_t)off * (BW * 8), X + (int64_t)csg.x * BW, (uint32_t)rows * (uint32_t)(BW * 8), &full[slot]);
            // the Q0 rows this chunk subtracts are streamed from DRAM by the compute warps a ring depth later: pull them into
            // L2 now, so that the trips wait for an L2 hit instead of a DRAM access
            if (FSUB && !(hint & 64) && lane < LZ_XS_OGROUPS && cog >= 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Q0 + (int64_t)cog * BW), "r"(8 * BW * 8) : "memory");
            if (++slot == stages) { slot = 0; phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int sub = lane / LW, l = lane % LW;
        const int p = (lane >> 2) & 1;                      // bank swizzle: odd groups read their upper 16 bytes first
        const int offA = 4 * l + 2 * p, offB = 4 * l + 2 * (1 - p);
        int trip_base = 0;
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; vchunk(it) < n_chunks; ++it) {
            const unsigned char *st = smem_raw + (size_t)slot * stage_bytes;
            const double *xw = reinterpret_cast<const double *>(st);
            const double *vs = reinterpret_cast<const double *>(st + xw_bytes);
            const uint16_t *ls = reinterpret_cast<const uint16_t *>(st + xw_bytes + LZ_XS_ES * 8);
            const int *rs = reinterpret_cast<const int *>(st + xw_bytes + LZ_XS_ES * 10);
            const int *rm = reinterpret_cast<const int *>(st + xw_bytes + LZ_XS_ES * 10 + LZ_XS_RCAP * 4);
            lz_mbar_wait_bounded(&full[slot], phase);
            const uint16_t *ds = reinterpret_cast<const uint16_t *>(st + xw_bytes + LZ_XS_ES * 10 + LZ_XS_RCAP * 8);
            const int4 cd = *reinterpret_cast<const int4 *>(st + xw_bytes + LZ_XS_ES * 10 + LZ_XS_RCAP * 10);   // the chunk's descriptor rode along
            const int a0 = cd.x, r0 = cd.z, r1 = cd.w;
            const int ra = r0 & ~7;
            const int rcnt = ((r1 + 1 - ra) + 7) & ~7;
            const bool rows_ok = rcnt <= LZ_XS_RCAP && (int64_t)ra + rcnt <= n_rows + 8;     // (the row arrays carry 8 spare entries)
            const int trips = (int)((r1 - r0 + RPW - 1) / RPW);
            const int t0 = (((warp - 1) - trip_base) % CW + CW) % CW;     // trips of all chunks dealt round-robin to the warps
            trip_base = (trip_base + trips) % CW;
            const int64_t rb0 = (int64_t)r0 + (int64_t)t0 * RPW;
            for (int64_t rb = (hint & 32) ? (int64_t)r1 : rb0; rb < r1; rb += NG) {     // hint bit 32 (dev): copies only, no compute
                const int64_t r = rb + sub;
                const bool valid = r < r1;
                // output row (tiled schedules walk a row-permuted operator; the map is in the slot, or read from global when
                // the chunk's rows did not fit the staging area)
                int64_t ro = 0;
                if (valid) ro = rows_ok ? rm[r - ra] : rowmap[r];
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                if (FSUB && valid) lz_ld256_stream_pol(Q0 + ro * BW + 4 * l, q0, q1, q2, q3, lz_policy_evict_first());
                double xa[2][2];
                if (GRAM) {
                    const int kk = lane & 3, mm = lane >> 2;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int64_t rr = rb + 4 * ks + kk;
                        // the row's own row of X: from the window when the chunk references it (an operator with a diagonal), else global
                        const unsigned dw = (rr < r1 && rows_ok) ? ds[rr - ra] : 0xFFFFu;
                        if (dw != 0xFFFFu) {
#pragma unroll
                            for (int a = 0; a < 2; ++a) xa[ks][a] = xw[dw * (unsigned)BW + 8 * a + mm];
                        } else {
                            int64_t rro = 0;
                            if (rr < r1) rro = rows_ok ? rm[rr - ra] : rowmap[rr];
#pragma unroll
                            for (int a = 0; a < 2; ++a) xa[ks][a] = rr < r1 ? __ldg(Xown + rro * BW + 8 * a + mm) : 0.0;
                        }
                    }
                }
                int s = 0, e = 0;                         // this row's entries, as slot indices
                if (valid) {
                    if (rows_ok) { s = rs[r - ra] - a0; e = rs[r - ra + 1] - a0; }
                    else { s = rowptr[r] - a0; e = rowptr[r + 1] - a0; }
                }
                double aA0 = 0.0, aA1 = 0.0, aB0 = 0.0, aB1 = 0.0;
                // every entry of the chunk is in the slot.  Slots past the row's end carry a zero value and a zero row (they add
                // exactly nothing; the order of the real products is the row's own)
                for (int k0 = s; k0 < e; k0 += G) {
                    double vv[G]; double2 xA[G], xB[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const bool on = k0 + g < e;
                        const int kk = min(k0 + g, e - 1);
                        const unsigned li = ls[kk];
                        vv[g] = on ? vs[kk] : 0.0;
                        const double *row = xw + li * (unsigned)BW;
                        // (predicated: the shared-memory data pipe is what bounds this kernel, a dead slot must not cost wavefronts)
                        xA[g] = on ? *reinterpret_cast<const double2 *>(row + offA) : make_double2(0.0, 0.0);
                        xB[g] = on ? *reinterpret_cast<const double2 *>(row + offB) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        aA0 = fma(vv[g], xA[g].x, aA0); aA1 = fma(vv[g], xA[g].y, aA1);
                        aB0 = fma(vv[g], xB[g].x, aB0); aB1 = fma(vv[g], xB[g].y, aB1);
                    }
                }
                double acc0 = p ? aB0 : aA0, acc1 = p ? aB1 : aA1, acc2 = p ? aA0 : aB0, acc3 = p ? aA1 : aB1;
                if (FSUB) {
                    __syncwarp();     // the entry loop has per-row trip counts: the MMAs need the whole warp
                    lz_dmma(acc0, acc1, q0, sbs[(0 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q0, sbs[(0 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q1, sbs[(1 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q1, sbs[(1 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q2, sbs[(2 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q2, sbs[(2 * 2 + 1) * 32 + lane]);
                    lz_dmma(acc0, acc1, q3, sbs[(3 * 2 + 0) * 32 + lane]); lz_dmma(acc2, acc3, q3, sbs[(3 * 2 + 1) * 32 + lane]);
                }
                if (valid) {
                    if (FSUB) lz_st256_pol(W + ro * BW + 4 * l, acc0, acc1, acc2, acc3, lz_policy_evict_first());
                    else lz_st256(W + ro * BW + 4 * l, acc0, acc1, acc2, acc3);
                }
                if (GRAM) {
                    // G += Xown[rb .. rb+8, :]^T W[rb .. rb+8, :]   (rows past the chunk contribute zeros: their acc is 0)
                    double *gt = gst + (warp - 1) * 8 * SPMM_GST;
                    *reinterpret_cast<double2 *>(gt + sub * SPMM_GST + 4 * l) = make_double2(acc0, acc1);
                    *reinterpret_cast<double2 *>(gt + sub * SPMM_GST + 4 * l + 2) = make_double2(acc2, acc3);
                    const int kk = lane & 3, mm = lane >> 2;
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
            if (++slot == stages) { slot = 0; phase ^= 1; }
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
                        const int pp = a * 8 + mm, q = b * 8 + 2 * kk;
                        if (w == 1) { gsm[pp + q * 16] = gacc[a][b][0]; gsm[pp + (q + 1) * 16] = gacc[a][b][1]; }
                        else { gsm[pp + q * 16] += gacc[a][b][0]; gsm[pp + (q + 1) * 16] += gacc[a][b][1]; }
                    }
            }
            __syncthreads();
        }
        for (int e = tid; e < 256; e += (1 + CW) * 32) gpart[(size_t)blockIdx.x * 256 + e] = gsm[e];
    }
}

// true when the operand-staging kernel can run this product: whole operator (no interior / boundary parts), contiguous
// panels, a window schedule that fits at least two ring slots for this panel width
static bool spmm_xs_plan(lz_ctx *ctx, const lz_matrix *A, int bw, const double *X, const double *W, int part, int *stages, int *stage_bytes,
                         int *xw_bytes)
{
    if (part != 0 || ctx->knobs.no_xs || ctx->spmv_variant == 9 || !A->tma_ok) return false;
    if (bw != 16 && !ctx->knobs.xs_force) return false;       // narrower panels: the per-chunk cost outweighs the gathers saved (measured)
    if (lz_matrix_prepare_xs(ctx, A) != LZ_OK || A->xs_state != 1) return false;
    if (((uintptr_t)X % 16) || ((uintptr_t)W % 32)) return false;
    if (A->xs_max_entries + 8 > LZ_XS_ES || A->xs_max_rows + 8 > LZ_XS_RCAP) return false;
    const int xw = ((A->xs_max_wrows * bw * 8) + 127) & ~127;
    const int sb = (xw + LZ_XS_ES * 10 + LZ_XS_RCAP * 10 + 16 + 127) & ~127;
    const int budget = 200 * 1024;                    // dynamic shared memory left beside the static Gram / fragment tiles
    int st = budget / sb;
    if (st > 8) st = 8;
    if (ctx->knobs.xs_stages >= 2 && ctx->knobs.xs_stages < st) st = ctx->knobs.xs_stages;
    if (st < 2) return false;
    *stages = st; *stage_bytes = sb; *xw_bytes = xw;
    return true;
}

template <int BW, bool FSUB, bool GRAM>
static int launch_spmm_xs(lz_ctx *ctx, const lz_matrix *A, const double *X, double *W, const double *Q0, const double *Bm, int stages,
                          int stage_bytes, int xw_bytes, const double *Xown = nullptr, double *gpart = nullptr, int *grid_out = nullptr)
{
    const size_t smem = (size_t)stages * stage_bytes + 128;
    LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_spmm_xs<BW, LZ_XS_CW, FSUB, GRAM>, 200 * 1024 + 128));   // (the size varies with the operator)
    int grid = ctx->sm_count;
    if (grid > A->xs_n_chunks) grid = A->xs_n_chunks;
    if (grid_out) *grid_out = grid;
    k_spmm_xs<BW, LZ_XS_CW, FSUB, GRAM><<<grid, (1 + LZ_XS_CW) * 32, smem, ctx->stream>>>(
        A->xs_n_chunks, A->n_rows, A->xs_desc, A->xs_meta, A->xs_seg, A->xs_rowptr, A->xs_lidx, A->xs_vals, A->xs_rowmap, X, W, Q0, Bm,
        stages, stage_bytes, xw_bytes, ctx->knobs.spmm_hint >= 0 ? ctx->knobs.spmm_hint : 0, Xown, gpart, A->xs_oseg, A->xs_dli);
    return LZ_OK;
}
