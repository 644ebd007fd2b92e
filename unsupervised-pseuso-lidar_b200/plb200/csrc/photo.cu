// Fused photometric reprojection loss (live mode: warp + L1 mean), forward and
// gradients in one pass.  See DESIGN.md "photo_l1_kernel".
//
// Work decomposition.  What one warp evaluates per target pixel is a COMBO: two (source, scale) samples riding the
// two halves of Blackwell's packed fp32 pipe - two sources at one scale, or ONE source at TWO scales (same pixel
// ray, two depths) - or a single sample on the scalar pipe when a job has an odd number of them.  The unit of work
// is one ROW SEGMENT of one combo: 32 consecutive pixels of one row of one target image of one job (direction),
// processed by one warp: every global access of the warp - target pixels, disparity, the 2x2x3 bilinear taps of
// each sample - is a (nearly) contiguous 128-byte piece of a planar NCHW row.  Units are ordered (job, image,
// 32-px column strip, combo, row), so a warp that walks its units moves DOWN a strip and re-uses the source rows it
// has just pulled into L1.
//
// The grid is persistent: SMs x (resident blocks per SM) blocks, and the weighted unit list is cut into equal
// contiguous ranges, one per warp, so the tail is a few row segments long instead of one tile.  A block's range
// touches at most two (job, image) pairs; K^-1 and P = K.[R|t] for both are built once in the block prologue and
// kept in shared memory.
//
// Low-resolution scales (F.interpolate of the DEPTH, losses.py:214-215): the warp walks down rows, so the two
// low-resolution depth rows it interpolates between live in registers and are replaced as the window advances
// (streaming upsample: two FMAs per pixel instead of four loads and four reciprocals); the TRANSPOSED upsample of
// the gradient runs the same way in the other direction - every lane accumulates its column's contribution to the
// two low-resolution rows in registers and leaves one value per low-resolution row; photo_lowres_merge_kernel then
// gathers the columns.  No full-resolution scratch plane is written or read for these scales.
//
// Reductions (loss, 3x4 projection-matrix gradient per source) are kept in registers across a warp's run of rows,
// reduced warp -> block record in a fixed order; a small finalize kernel (one block per image) sums the records of
// each (job, image) pair, runs the pose chain and adds the loss: no atomics on the results, bitwise repeatable, and
// no fence / ticket traffic in the main kernel.
#include "photo_common.cuh"
#include <mutex>
#include <unordered_map>

namespace plb {

#ifdef PLB_DEBUG_TIMERS
__device__ unsigned long long g_dbg_t[32];
__device__ unsigned long long g_dbg_blk[2048];
__device__ unsigned int g_dbg_sm[2048];
__device__ __forceinline__ void dbg_stamp(int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_dbg_t[slot] = t;
}
#define DBG_STAMP(cond, slot) do { if (cond) dbg_stamp(slot); } while (0)
#else
#define DBG_STAMP(cond, slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Per-pixel stages.  The two halves of a combo ride the packed fp32 pipe (FFMA2 / FADD2 / FMUL2: half 0 in the
// low, half 1 in the high half of a 64-bit register pair), which halves the issue slots of the arithmetic; the
// coordinates of both halves are computed first, ONE warp vote decides between the unpredicated and the
// predicated tap loads, all 12 x NS loads are issued back to back, and only then is anything blended.
//
// Projection: cam = D * A + p3 with A = Q . (x, y, 1), Q = P[:, :3] . K^-1 (composed in fp64 in the block
// prologue, rounded once): the ray is never formed, and the x part of A (the lane's column is fixed for a whole
// run) is hoisted out of the row loop.  The perspective divide is one MUFU.RCP + one Newton step + a residual
// correction (as accurate as an IEEE divide: the sample position is the difference of two ~W-sized numbers, every
// ulp of px is 6e-5 px of bilinear weight); the normalise / un-normalise chain of the reference
// (transform.py:143-148 + grid_sample) is the identity and is dropped; the bilinear blend is written as nested
// lerps whose intermediates ARE the coordinate derivatives (d proj / d iy = bot - top).
//
// Depth gradient: d loss / d D = g_cam . A.  Because g_cam . (cx, cy, ze) = 0 identically (the perspective divide
// is scale invariant) and (cx, cy, ze) = D * A + p3', this equals -(g_cam . p3') / D - the analytically cancelled
// form, free of the ~W-sized cancellation the chain-rule form carries, and A need not stay live across the loads.
// ---------------------------------------------------------------------------------------------
// where the constants of a combo live: T2 table `tab` for a packed combo, Q[k0] for a single sample
struct ComboRef { int tab, k0; };

// hoisted x part of A: Ax[r] = qx * x + qz
__device__ __forceinline__ void load_ax(const PairConst& pc, ComboRef c, float xf, float (&Ax)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) { const float4 q = pc.Q[c.k0][r]; Ax[r] = fmaf(q.x, xf, q.z); }
}
__device__ __forceinline__ void load_ax(const PairConst& pc, ComboRef c, float xf, float2 (&Ax)[3]) {
    const float2 xv = make_float2(xf, xf);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float4 t = pc.T2[c.tab][r][0];
        Ax[r] = __ffma2_rn(make_float2(t.x, t.y), xv, make_float2(t.z, t.w));
    }
}
// per row: y coefficient and p3 of cam row r
__device__ __forceinline__ void load_qy(const PairConst& pc, ComboRef c, int r, float& qy, float& p3) {
    const float4 q = pc.Q[c.k0][r];
    qy = q.y; p3 = q.w;
}
__device__ __forceinline__ void load_qy(const PairConst& pc, ComboRef c, int r, float2& qy, float2& p3) {
    const float4 t = pc.T2[c.tab][r][1];
    qy = make_float2(t.x, t.y); p3 = make_float2(t.z, t.w);
}
// p3 re-read from shared memory (asm volatile: a fresh load, not a value held in registers across the taps)
__device__ __forceinline__ void load_p3(const PairConst& pc, ComboRef c, int r, float& w) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(&pc.Q[c.k0][r].w);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(a));
}
__device__ __forceinline__ void load_p3(const PairConst& pc, ComboRef c, int r, float2& w) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(&pc.T2[c.tab][r][1].z);
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w.x), "=f"(w.y) : "r"(a));
}

// coordinates of the samples of one combo at one pixel (stage A output)
template <int NS>
struct GState {
    typename Vec<NS>::T inv, px, py, fx, fy;
    int x0[NS], y0[NS];
    bool all_in;
};

// stage A: project the pixel into the source of every half
template <int NS>
__device__ __forceinline__ void group_project(const PairConst& pc, ComboRef c, int H, int W,
                                              const typename Vec<NS>::T (&Ax)[3], float yf, typename Vec<NS>::T Dv,
                                              GState<NS>& g) {
    typedef typename Vec<NS>::T V;
    V yv, eps, neg1, two;
    v_bc(yv, yf); v_bc(eps, 1e-5f); v_bc(neg1, -1.0f); v_bc(two, 2.0f);
    V cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V qy, p3;
        load_qy(pc, c, r, qy, p3);
        cam[r] = v_fma(Dv, v_fma(qy, yv, Ax[r]), p3);
    }
    const V ze = v_add(cam[2], eps);
    const V nze = v_mul(ze, neg1);
    V inv;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v_get(ze, k)));
        v_set(inv, k, r);
    }
    inv = v_mul(inv, v_fma(nze, inv, two));                  // Newton step
    V px = v_mul(cam[0], inv), py = v_mul(cam[1], inv);
    px = v_fma(v_fma(px, nze, cam[0]), inv, px);             // residual correction: IEEE-accurate quotient
    py = v_fma(v_fma(py, nze, cam[1]), inv, py);
    bool all_in = true;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        // clamp keeps float->int defined; NaN maps to -2 (out of the image)
        const float ixc = fminf(fmaxf(v_get(px, k), -2.0f), (float)(W + 1));
        const float iyc = fminf(fmaxf(v_get(py, k), -2.0f), (float)(H + 1));
        const float xfl = floorf(ixc), yfl = floorf(iyc);
        g.x0[k] = (int)xfl; g.y0[k] = (int)yfl;
        v_set(g.fx, k, ixc - xfl); v_set(g.fy, k, iyc - yfl);
        all_in = all_in && ((unsigned)g.x0[k] < (unsigned)(W - 1)) && ((unsigned)g.y0[k] < (unsigned)(H - 1));
    }
    g.inv = inv; g.px = px; g.py = py; g.all_in = all_in;
}

// stage B: issue the 12 x NS tap loads (one warp vote picks unpredicated or predicated loads).  With compile-time
// image dimensions every tap of a sample is ONE 64-bit address plus an immediate offset.
template <int NS, bool ALL_VALID>
__device__ __forceinline__ void group_load(const float* const (&cbp)[NS], int plane, int H, int W, bool valid, bool pf,
                                           const GState<NS>& g, typename Vec<NS>::T (&v)[3][4], unsigned (&msk)[NS]) {
    const bool all_in = g.all_in;
    if (__all_sync(0xffffffffu, all_in || !valid)) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            // (when every lane is a pixel the vote already says that every footprint is interior)
            const int o00 = (ALL_VALID || (all_in && valid)) ? g.y0[k] * W + g.x0[k] : 0;
            const float* __restrict__ q = cbp[k] + o00;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v_set(v[c][0], k, __ldg(q + c * plane)); v_set(v[c][1], k, __ldg(q + (c * plane + 1)));
                v_set(v[c][2], k, __ldg(q + (c * plane + W))); v_set(v[c][3], k, __ldg(q + (c * plane + W + 1)));
            }
            if (PH_PF_SRC > 0 && pf && (NS == 1 || k == 0 || cbp[NS - 1] != cbp[0])) {
                // the warp walks DOWN a strip: the source row the next unit(s) will newly touch, one line per channel
                // (a scale pair samples ONE source twice, a few pixels apart: one prefetch serves both halves)
                const int opf = W + PH_PF_SRC * W;
                prefetch_l1(q + opf); prefetch_l1(q + (opf + plane)); prefetch_l1(q + (opf + 2 * plane));
            }
            msk[k] = valid ? 15u : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const float* __restrict__ cb = cbp[k];
            const bool vx0 = (unsigned)g.x0[k] < (unsigned)W, vx1 = (unsigned)(g.x0[k] + 1) < (unsigned)W;
            const bool vy0 = (unsigned)g.y0[k] < (unsigned)H, vy1 = (unsigned)(g.y0[k] + 1) < (unsigned)H;
            const bool mnw = valid && vx0 && vy0, mne = valid && vx1 && vy0;
            const bool msw = valid && vx0 && vy1, mse = valid && vx1 && vy1;
            const int o00 = g.y0[k] * W + g.x0[k];
            const int o01 = o00 + W, o10 = o00 + plane, o11 = o10 + W, o20 = o10 + plane, o21 = o20 + W;
            v_set(v[0][0], k, ldg_pred(cb + o00, mnw)); v_set(v[0][1], k, ldg_pred(cb + o00 + 1, mne));
            v_set(v[0][2], k, ldg_pred(cb + o01, msw)); v_set(v[0][3], k, ldg_pred(cb + o01 + 1, mse));
            v_set(v[1][0], k, ldg_pred(cb + o10, mnw)); v_set(v[1][1], k, ldg_pred(cb + o10 + 1, mne));
            v_set(v[1][2], k, ldg_pred(cb + o11, msw)); v_set(v[1][3], k, ldg_pred(cb + o11 + 1, mse));
            v_set(v[2][0], k, ldg_pred(cb + o20, mnw)); v_set(v[2][1], k, ldg_pred(cb + o20 + 1, mne));
            v_set(v[2][2], k, ldg_pred(cb + o21, msw)); v_set(v[2][3], k, ldg_pred(cb + o21 + 1, mse));
            msk[k] = (mnw ? 1u : 0u) | (mne ? 2u : 0u) | (msw ? 4u : 0u) | (mse ? 8u : 0u);
        }
    }
}

// stage C: blend, L1, gradient terms.  `gpv`: per half, g_cam . p3' (the depth-gradient numerator).
template <bool GRAD, bool IMG_GRAD, int NS>
__device__ __forceinline__ void group_blend(const PairConst& pc, ComboRef cr, float* const (&gsp)[NS], int plane, int W,
                                            float yf, typename Vec<NS>::T Dv, const float (&t)[3], float w_e, bool valid,
                                            const GState<NS>& g, const typename Vec<NS>::T (&v)[3][4],
                                            const unsigned (&msk)[NS], typename Vec<NS>::T (&acc)[9], float& l1acc,
                                            typename Vec<NS>::T& gpv, float (&gt)[3]) {
    typedef typename Vec<NS>::T V;
    V yv, eps, neg1;
    v_bc(yv, yf); v_bc(eps, 1e-5f); v_bc(neg1, -1.0f);
    V Gx, Gy;
    v_bc(Gx, 0.0f); v_bc(Gy, 0.0f);
    float l1 = 0.0f;
    float e[NS][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const V dA = v_sub(v[c][1], v[c][0]), dB = v_sub(v[c][3], v[c][2]);
        const V top = v_fma(g.fx, dA, v[c][0]), bot = v_fma(g.fx, dB, v[c][2]);
        const V dV = v_sub(bot, top);
        const V proj = v_fma(g.fy, dV, top);
        V sg;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const float d = v_get(proj, k) - t[c];
            l1 += fabsf(d);
            if (GRAD) {
                // sign(d) with sign(0) = 0 (nn.L1Loss): one compare + one bit merge
                const float ne = (d != 0.0f) ? 1.0f : 0.0f;
                const float s = __int_as_float(__float_as_int(ne) | (__float_as_int(d) & 0x80000000));
                v_set(sg, k, s);
                e[k][c] = s;
            }
        }
        if (GRAD) {
            Gx = v_fma(sg, v_fma(g.fy, v_sub(dB, dA), dA), Gx);
            Gy = v_fma(sg, dV, Gy);
        }
    }
    l1acc += valid ? l1 : 0.0f;
    if (GRAD) {
        V gi;
        V s = v_fma(Gx, g.px, v_mul(Gy, g.py));
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const bool use = msk[k] != 0u;
            // the selects also keep the inf/NaN of a degenerate z out of the sums
            v_set(gi, k, use ? w_e * v_get(g.inv, k) : 0.0f);
            v_set(s, k, use ? v_get(s, k) : 0.0f);
        }
        const V gcx = v_mul(Gx, gi), gcy = v_mul(Gy, gi);
        const V gcz = v_mul(v_mul(s, gi), neg1);
        {
            V q3[3];   // p3 again (shared memory): cheaper than six registers held across the loads
#pragma unroll
            for (int r = 0; r < 3; ++r) load_p3(pc, cr, r, q3[r]);
            gpv = v_fma(gcx, q3[0], v_fma(gcy, q3[1], v_mul(gcz, v_add(q3[2], eps))));
        }
        // d loss / d P[r][:] = sum g_cam[r] * (D * ray, 1), ray = K^-1 (x, y, 1): accumulated in pixel
        // coordinates - sum h_r, sum h_r * y (and x * sum h_r at the flush, x being fixed per lane) - and
        // mapped through K^-1 by the finalize kernel
        const V hx = v_mul(gcx, Dv), hy = v_mul(gcy, Dv), hz = v_mul(gcz, Dv);
        acc[0] = v_add(acc[0], hx); acc[1] = v_add(acc[1], hy); acc[2] = v_add(acc[2], hz);
        acc[3] = v_fma(hx, yv, acc[3]); acc[4] = v_fma(hy, yv, acc[4]); acc[5] = v_fma(hz, yv, acc[5]);
        acc[6] = v_add(acc[6], gcx); acc[7] = v_add(acc[7], gcy); acc[8] = v_add(acc[8], gcz);
        if (IMG_GRAD) {
            // deterministic mode: contributions in units of the call's largest weight (pc.det_rho = this job's share),
            // rounded once to 2^-30 and added as 64-bit integers (photo_common.cuh)
            const bool det = pc.det_rho > 0.0f;
            const float m = valid ? (det ? pc.det_rho * PH_DET_ONE : w_e) : 0.0f;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) gt[c] -= m * e[k][c];
                float* gbase = gsp[k];
                if (gbase != nullptr) {
                    const float fxk = v_get(g.fx, k), fyk = v_get(g.fy, k);
                    const float wnw = (1.0f - fxk) * (1.0f - fyk), wne = fxk * (1.0f - fyk);
                    const float wsw = (1.0f - fxk) * fyk, wse = fxk * fyk;
                    const int og = g.y0[k] * W + g.x0[k];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float ec = m * e[k][c];
                        if (det) {
                            unsigned long long* q = reinterpret_cast<unsigned long long*>(gbase) + (og + c * plane);
                            if (msk[k] & 1u) atomicAdd(q, (unsigned long long)__float2ll_rn(wnw * ec));
                            if (msk[k] & 2u) atomicAdd(q + 1, (unsigned long long)__float2ll_rn(wne * ec));
                            if (msk[k] & 4u) atomicAdd(q + W, (unsigned long long)__float2ll_rn(wsw * ec));
                            if (msk[k] & 8u) atomicAdd(q + W + 1, (unsigned long long)__float2ll_rn(wse * ec));
                        } else {
                            float* q = gbase + (og + c * plane);
                            if (msk[k] & 1u) atomicAdd(q, wnw * ec);
                            if (msk[k] & 2u) atomicAdd(q + 1, wne * ec);
                            if (msk[k] & 4u) atomicAdd(q + W, wsw * ec);
                            if (msk[k] & 8u) atomicAdd(q + W + 1, wse * ec);
                        }
                    }
                }
            }
        }
    }
}

// Per-lane accumulators of one source -> this warp's shared record (ACCUMULATED in place).
// Record of source i: [0..2] sum h_r, [3..5] sum h_r * x, [6..8] sum h_r * y, [9..11] sum g_cam[r];
// value PLB_MAX_SRC*12: sum |diff| (added once per run: `l1` is zero in every other call).
__device__ __forceinline__ void flush_source(const float (&a9)[9], int i, float l1, float xlane, float* rec, int lane) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        v[q] = a9[q]; v[3 + q] = a9[q] * xlane;
        v[6 + q] = a9[3 + q]; v[9 + q] = a9[6 + q];
    }
    v[12] = l1;
    v[13] = v[14] = v[15] = 0.0f;
    int which;
    const float r = warp_reduce16(v, lane, which);
    if ((lane & 1) == 0) {
        if (which < 12) rec[i * 12 + which] += r;
        else if (which == 12) rec[PLB_MAX_SRC * 12] += r;
    }
}

// ---------------------------------------------------------------------------------------------
// Depth streams.  A combo reads one depth per pixel (two for a scale pair).
//   FULL scale: the disparity of the next row is loaded one iteration ahead (software pipeline).
//   Low-resolution scale: F.interpolate(depth, [H, W], bilinear, align_corners=False) (losses.py:214-215; the
//   reference interpolates DEPTH, not disparity).  A run is walked in CHUNKS of at most PH_CHUNK rows that never
//   cross a multiple of PH_CHUNK.  Before the rows of a chunk a short pre-pass walks down the chunk once with a
//   two-row window of x-interpolated low-res depths (the window advances by exactly one row at a time because the
//   factor is >= 1) and leaves the upsampled depth of every pixel of the lane's column in a per-lane slot of shared
//   memory - D = ly0 * r0 + ly1 * r1 with the same association as the 4-tap form ly0 * (lx0 v00 + lx1 v01) +
//   ly1 * (lx0 v10 + lx1 v11).  The row loop then costs one shared load per pixel and holds NO low-resolution
//   state in registers; it puts the gradient with respect to the upsampled depth back into the same slot, and a
//   post-pass runs the TRANSPOSED upsample over the chunk: a_lo / a_hi collect ly0 * g and ly1 * g of the rows and
//   are emitted - one value per low-res row and lane - when the window advances.
// ---------------------------------------------------------------------------------------------
constexpr int PH_CHUNK = PH_ROW_ALIGN;

template <bool HEAD>
__device__ __forceinline__ float to_depth(const plb_photo_args& a, float raw) {
    if (HEAD) raw = head_disp(raw, a.head_alpha, a.head_beta);
    return a.input_is_depth == PLB_INPUT_DEPTH ? raw : rcp_nr(fmaf(a.disp_a, raw, a.disp_b));
}

// x-interpolated depth of low-res row j of scale s at column xc
template <bool HEAD>
__device__ __forceinline__ float low_row(const plb_photo_args& a, const PairConst& pc, int s, int j, int xc) {
    int x0, x1; float lx0, lx1;
    const int dw = pc.dw[s];
    up_coord(xc, pc.sx[s], dw, x0, x1, lx0, lx1);
    const float* row = pc.disp[s] + j * dw;
    const float v0 = to_depth<HEAD>(a, __ldg(row + x0)), v1 = to_depth<HEAD>(a, __ldg(row + x1));
    return lx0 * v0 + lx1 * v1;
}

// pre-pass, any ratio (up_coord per row): upsampled depth of rows [yc, yc + nr) of the lane's column -> slot[r * 32]
template <bool HEAD>
__device__ __forceinline__ void low_prepass(const plb_photo_args& a, const PairConst& pc, int s, int xc, int yc, int nr,
                                            float* slot) {
    const int dh = pc.dh[s];
    const float sy = pc.sy[s];
    int y0, y1; float l0, l1;
    up_coord(yc, sy, dh, y0, y1, l0, l1);
    int jc = y0;
    float r0 = low_row<HEAD>(a, pc, s, y0, xc), r1 = low_row<HEAD>(a, pc, s, y1, xc);
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
        up_coord(yc + r, sy, dh, y0, y1, l0, l1);
        if (y0 != jc) {                                // the window advances by one low-res row (warp-uniform)
            jc = y0;
            r0 = r1;
            r1 = low_row<HEAD>(a, pc, s, y1, xc);
        }
        slot[r * 32] = l0 * r0 + l1 * r1;
    }
}

// One partial row of the transposed upsample leaves the warp.  The chunk [ya, yb) holds either all of the
// full-resolution rows that feed low-res row j or a part of them; chunks end on multiples of PH_CHUNK >= 2 * factor
// rows, so at most ONE chunk boundary crosses the (at most 2.5 * factor rows long) footprint of a low-res row and the
// row has at most two contributors: slot 0 takes the part that starts the footprint, slot 1 the part that ends it,
// and a chunk that holds the whole footprint zeroes slot 1.  Every (row, column) of both slots is written exactly
// once per launch, in no particular order, and summed in a fixed order by the merge kernel: the result does not
// depend on how the unit list was cut.
__device__ __forceinline__ void low_emit(float* g, int dh, int f, int H, int W, int j, float val, int x, bool valid, int ya, int yb) {
    if (j >= dh || !valid) return;
    const int hf = f >> 1;
    const int first = (j <= 1) ? 0 : (j - 1) * f + hf;                 // first row with y0 == j - 1 (or the image top)
    const int last = (j == dh - 1) ? H - 1 : (j + 1) * f + hf - 1;     // last row with y0 == j
    const bool before = first >= ya, after = last < yb;
    float* row0 = g + ((size_t)j * W + x);
    float* row1 = row0 + (size_t)dh * W;
    if (before) {
        *row0 = val;
        if (after) *row1 = 0.0f;
    } else {
        *row1 = val;
    }
}

// ---- chunk passes of the staged (LOW) kernels.  `slot` = this lane's column of the stream's [PH_CHUNK][32] staging
//      area; row r of the chunk lives at slot[r * 32].  Before the rows of a chunk a pre-pass leaves the DEPTH of every
//      pixel there; the row loop replaces it by w = d loss / d depth; a post-pass turns w into the gradient the caller
//      asked for. ------------------------------------------------------------------------------------------------------

// FULL scale: the chunk's disparities in flight together instead of one software-pipelined load per row
template <bool HEAD>
__device__ __forceinline__ void pre_full(const plb_photo_args& a, const float* disp, int W, int nr, float* slot) {
#pragma unroll 4
    for (int r = 0; r < nr; ++r) slot[r * 32] = to_depth<HEAD>(a, __ldg(disp + r * W));
}

// FULL scale: w -> d loss / d disparity (the map re-read from L1: D is not kept)
template <bool HEAD>
__device__ __forceinline__ void post_full(const plb_photo_args& a, const float* disp, float* g, bool shared, int W, int nr,
                                          const float* slot) {
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
        float gv = slot[r * 32];
        if (a.input_is_depth != PLB_INPUT_DEPTH) {
            const float D = to_depth<HEAD>(a, __ldg(disp + r * W));
            gv *= -a.disp_a * D * D;                       // d D / d disp = -disp_a * D^2
            if (HEAD) gv *= head_chain_from_depth(D, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
        }
        if (shared) atomicAdd(g + r * W, gv); else g[r * W] = gv;
    }
}

// LOWSCRATCH scale: w into the full-resolution scratch plane (photo_upsample_T_kernel gathers it)
__device__ __forceinline__ void post_scratch(float* g, bool shared, int W, int nr, const float* slot) {
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
        if (shared) atomicAdd(g + r * W, slot[r * 32]); else g[r * W] = slot[r * 32];
    }
}

// Low-resolution scale with an integer power-of-two factor f: row y has y0 = (y - f/2) >> lg and
// l1 = ((y - f/2) mod f + 0.5) / f - up_coord's values exactly (dyadic rationals) - so walking down the rows is a
// phase counter: k counts the rows of a band (the f rows that share y0), and the two-row window of x-interpolated
// depths advances when k wraps.  At the image top the first f/2 rows clamp to (row 0, l1 = 0): t < 0.  The raw taps
// of the row the NEXT advance needs are loaded one band ahead.
template <bool HEAD>
__device__ __forceinline__ void pre_low_p2(const plb_photo_args& a, const PairConst& pc, int s, int f, int xc, int yc, int nr,
                                           float* slot) {
    const int dh = pc.dh[s], dw = pc.dw[s], lg = 31 - __clz(f);
    const float inv_f = pc.sy[s];                          // 1 / f exactly
    int x0, x1; float lx0, lx1;
    up_coord(xc, pc.sx[s], dw, x0, x1, lx0, lx1);
    const float* base = pc.disp[s];
    int t = yc - (f >> 1);
    int j = t >> lg, k = t & (f - 1);                      // (arithmetic shift: -1 above the first band)
    auto raw = [&](int jj, float& u, float& v) {
        jj = min(max(jj, 0), dh - 1);
        u = __ldg(base + (jj * dw + x0)); v = __ldg(base + (jj * dw + x1));
    };
    float u0, v0, u1, v1, un, vn;
    raw(j, u0, v0); raw(j + 1, u1, v1); raw(j + 2, un, vn);
    float r0 = lx0 * to_depth<HEAD>(a, u0) + lx1 * to_depth<HEAD>(a, v0);
    float r1 = lx0 * to_depth<HEAD>(a, u1) + lx1 * to_depth<HEAD>(a, v1);
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
        const float l1 = t < 0 ? 0.0f : ((float)k + 0.5f) * inv_f;
        const float l0 = 1.0f - l1;
        slot[r * 32] = l0 * r0 + l1 * r1;
        ++t; ++k;
        if (k == f) {                                      // the window advances by one low-res row (warp-uniform)
            k = 0; ++j;
            r0 = r1;
            r1 = lx0 * to_depth<HEAD>(a, un) + lx1 * to_depth<HEAD>(a, vn);
            raw(j + 2, un, vn);
        }
    }
}

// ... and its transpose: w of the chunk's rows -> partial low-res rows (see low_emit)
__device__ __forceinline__ void post_low_p2(const PairConst& pc, int s, int f, float* g, int H, int W, int x, bool valid,
                                            int yc, int nr, const float* slot) {
    const int dh = pc.dh[s], lg = 31 - __clz(f);
    const float inv_f = pc.sy[s];
    int t = yc - (f >> 1);
    int j = t >> lg, k = t & (f - 1);
    float a_lo = 0.0f, a_hi = 0.0f;
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
        const float l1 = t < 0 ? 0.0f : ((float)k + 0.5f) * inv_f;
        const float gz = slot[r * 32];
        a_lo = fmaf(1.0f - l1, gz, a_lo);
        // bottom border (y1 == y0 == dh - 1): both weights act on row j
        if (j >= dh - 1) a_lo = fmaf(l1, gz, a_lo); else a_hi = fmaf(l1, gz, a_hi);
        ++t; ++k;
        if (k == f) {
            k = 0;
            if (j >= 0) { low_emit(g, dh, f, H, W, j, a_lo, x, valid, yc, yc + nr); a_lo = a_hi; }
            else a_lo += a_hi;                             // the clamped rows of the image top belong to low-res row 0
            a_hi = 0.0f; ++j;
        }
    }
    if (j >= 0) {
        low_emit(g, dh, f, H, W, j, a_lo, x, valid, yc, yc + nr);
        low_emit(g, dh, f, H, W, j + 1, a_hi, x, valid, yc, yc + nr);
    } else {
        low_emit(g, dh, f, H, W, 0, a_lo + a_hi, x, valid, yc, yc + nr);
    }
}

// One run: consecutive rows [y, y + rows) of one 32-px strip of one combo of one (job, image) pair.
//   NS = 2, ND = 1: two sources (k0, k0 + 1) at scale s0;  NS = 2, ND = 2: source k0 at scales s0, s1;
//   NS = 1, ND = 1: source k0 at scale s0 on the scalar pipe.
// LOW = false (every scale of the launch is full resolution): the disparity of the next row is loaded one iteration
// ahead and the gradient is written from the row loop.  LOW = true: the staged form described above - the row loop
// is the same for every kind of scale.  `stage`: this warp's [2][PH_CHUNK][32] floats of shared memory (LOW only).
template <bool GRAD, bool IMG_GRAD, int NS, int ND, bool HEAD, bool LOW, int CH, int CW>
__device__ __forceinline__ void run_combo(const plb_photo_args& a, const PairConst& pc, const PhotoCombo cb, int strip,
                                          int y, int rows, int lane, float* rec, float* stage) {
    typedef typename Vec<NS>::T V;
    const int H = CH > 0 ? CH : a.H, W = CW > 0 ? CW : a.W, plane = H * W;
    constexpr bool ALL_VALID = CW > 0 && (CW % 32) == 0;
    constexpr int SST = PH_CHUNK * 32;                     // floats per stream of the staging area
    const float w_e = pc.w_e;
    const int x = strip * 32 + lane;
    const bool valid = ALL_VALID ? true : (x < W);
    const int xc = min(x, W - 1);
    const float xf = (float)x;
    const float* __restrict__ tgt_b = pc.tgt;
    const int yb = y + rows;

    ComboRef cr;
    cr.k0 = cb.k0;
    cr.tab = (ND == 2) ? (PLB_MAX_SRC / 2 + cb.k0) : (cb.k0 >> 1);
    const float* cbp[NS];
    float* gsp[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ks = (ND == 2) ? cb.k0 : cb.k0 + k;
        cbp[k] = pc.src[ks];
        gsp[k] = IMG_GRAD ? pc.g_src[ks] : nullptr;
    }
    V Ax[3];
    load_ax(pc, cr, xf, Ax);

    // per depth stream: scale index and PH_SM_* (| 4: a second combo adds into the same map)
    int sc[ND], md[ND];
    float u0[ND];                                      // !LOW: raw value of the NEXT row
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const int s = d == 0 ? cb.s0 : cb.s1;
        sc[d] = s;
        md[d] = pc.smode[s];
        u0[d] = 0.0f;
        if (!LOW) u0[d] = __ldg(pc.disp[s] + (y * W + xc));
    }

    V acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v_bc(acc[k], 0.0f);
    float l1acc = 0.0f;

    // software pipeline: the target pixel of the NEXT row is loaded while this row is processed, so a warp pays one
    // memory round trip per row (the taps), not two
    int o = y * W + xc;
    float tn[3];
    tn[0] = __ldg(tgt_b + o); tn[1] = __ldg(tgt_b + (o + plane)); tn[2] = __ldg(tgt_b + (o + 2 * plane));
#pragma unroll 1
    while (y < yb) {
        // ---- one chunk: rows [yc, yc + nr), never across a multiple of PH_CHUNK ---------------------------------
        const int yc = y;
        const int nr = LOW ? (min(yb, (yc / PH_CHUNK + 1) * PH_CHUNK) - yc) : (yb - yc);
        if (LOW) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const int s = sc[d], f = pc.fac[s];
                float* slot = stage + d * SST + lane;
                if ((md[d] & 3) == PH_SM_FULL) pre_full<HEAD>(a, pc.disp[s] + o, W, nr, slot);
                else if (f > 0) pre_low_p2<HEAD>(a, pc, s, f, xc, yc, nr, slot);
                else low_prepass<HEAD>(a, pc, s, xc, yc, nr, slot);
            }
        }
        const int ye = yc + nr;
        const int oc = o;                                  // offset of the chunk's first pixel
        float* sp = stage + lane;
#pragma unroll 1
        for (; y < ye; ++y, o += W, sp += 32) {
            float t[3], gt[3] = {0.0f, 0.0f, 0.0f};
            t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
            const bool pf = y + 1 < yb;
            float D[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) D[d] = LOW ? sp[d * SST] : to_depth<HEAD>(a, u0[d]);
            if (pf) {
                const int on = o + W;
                tn[0] = __ldg(tgt_b + on); tn[1] = __ldg(tgt_b + (on + plane)); tn[2] = __ldg(tgt_b + (on + 2 * plane));
                if (!LOW) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) u0[d] = __ldg(pc.disp[sc[d]] + on);
                }
            }
            V Dv;
            if (ND == 2) { v_set(Dv, 0, D[0]); v_set(Dv, 1, D[ND - 1]); } else v_bc(Dv, D[0]);
            const float yf = (float)y;
            GState<NS> g;
            V v[3][4];
            unsigned msk[NS];
            V gpv;
            v_bc(gpv, 0.0f);
            group_project<NS>(pc, cr, H, W, Ax, yf, Dv, g);
            group_load<NS, ALL_VALID>(cbp, plane, H, W, valid, pf, g, v, msk);
            group_blend<GRAD, IMG_GRAD, NS>(pc, cr, gsp, plane, W, yf, Dv, t, w_e, valid, g, v, msk, acc, l1acc, gpv, gt);
            if (GRAD) {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const float gp = (ND == 2) ? v_get(gpv, d) : v_hsum(gpv);
                    const float Dd = D[d];
                    if (LOW) {
                        // d loss / d D = -gp / D, for every kind of scale; the post-pass takes it from here
                        sp[d * SST] = valid ? -gp * rcp_nr(Dd) : 0.0f;
                    } else {
                        float* gdst = pc.g_disp[sc[d]];
                        if (gdst == nullptr) continue;
                        // d loss / d D = -gp / D;  d D / d disp = -disp_a * D^2
                        float gv = a.input_is_depth == PLB_INPUT_DEPTH ? -gp * rcp_nr(Dd) : a.disp_a * Dd * gp;
                        if (HEAD) gv *= head_chain_from_depth(Dd, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
                        if (valid) { if (md[d] & 4) atomicAdd(gdst + o, gv); else gdst[o] = gv; }
                    }
                }
            }
            if (GRAD && IMG_GRAD && valid && pc.g_tgt != nullptr) {
                if (pc.det_rho > 0.0f) {
                    unsigned long long* gq = reinterpret_cast<unsigned long long*>(pc.g_tgt) + o;
                    atomicAdd(gq, (unsigned long long)__float2ll_rn(gt[0]));
                    atomicAdd(gq + plane, (unsigned long long)__float2ll_rn(gt[1]));
                    atomicAdd(gq + 2 * plane, (unsigned long long)__float2ll_rn(gt[2]));
                } else {
                    float* gq = pc.g_tgt + o;
                    atomicAdd(gq, gt[0]); atomicAdd(gq + plane, gt[1]); atomicAdd(gq + 2 * plane, gt[2]);
                }
            }
        }
        if (LOW && GRAD) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const int s = sc[d], f = pc.fac[s], mode = md[d] & 3;
                float* gdst = pc.g_disp[s];
                if (gdst == nullptr || mode == PH_SM_LOWNOGRAD) continue;
                const float* slot = stage + d * SST + lane;
                const bool shared = (md[d] & 4) != 0;
                if (mode == PH_SM_FULL) {
                    if (valid) post_full<HEAD>(a, pc.disp[s] + oc, gdst + oc, shared, W, nr, slot);
                } else if (mode == PH_SM_LOWSCRATCH) {
                    if (valid) post_scratch(gdst + oc, shared, W, nr, slot);
                } else {
                    // this contributor's [2][dh][W] partial rows
                    if ((d == 0 ? cb.c0 : cb.c1) != 0) gdst += (size_t)a.B * 2 * pc.dh[s] * W;
                    post_low_p2(pc, s, f, gdst, H, W, x, valid, yc, nr, slot);     // LOWFAST: f is 2, 4 or 8
                }
            }
        }
    }
    // end of the run: the pose / loss sums
    if (NS == 2 && ND == 1) {
        float a9[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a9[k] = v_get(acc[k], 0);
        flush_source(a9, cb.k0, l1acc, xf, rec, lane);
#pragma unroll
        for (int k = 0; k < 9; ++k) a9[k] = v_get(acc[k], 1);
        flush_source(a9, cb.k0 + 1, 0.0f, xf, rec, lane);
    } else {
        float a9[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a9[k] = v_hsum(acc[k]);      // both halves of a scale pair belong to the one source
        flush_source(a9, cb.k0, l1acc, xf, rec, lane);
    }
    __syncwarp();
}

// inverse of pos_of: the warp whose weight range holds `pos`
__host__ __device__ inline int photo_warp_of(const PhotoLaunch& p, long long pos) {
    const long long big = (long long)p.share_rem * (p.share + 1);
    if (pos < big) return (int)(pos / (p.share + 1));
    return p.share > 0 ? p.share_rem + (int)((pos - big) / p.share) : p.n_warps - 1;
}

// The pose chain  S -> dP (through K^-1) -> K^T.dP -> (rigid inverse) -> Rodrigues / Euler vjp  is linear
// in the 12 record sums S of a (pair, source): row k of its Jacobian is the chain applied to e_k.
// S: [0..2] sum h_r, [3..5] sum h_r x, [6..8] sum h_r y, [9..11] sum g_cam[r].
__device__ inline void photo_pose_jacobian(const plb_photo_args& a, const plb_photo_job& job, int b, int i, int k,
                                           float* J /* [12][6] of this (pair, source) */) {
    const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
    float kinv[9], S[12], dP[12], dM[12], g6[6];
    kinv_f32(Kb, a.k_is_f64, kinv);
#pragma unroll
    for (int m = 0; m < 12; ++m) S[m] = (m == k) ? 1.0f : 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int m = 0; m < 3; ++m)
            dP[r * 4 + m] = kinv[m * 3 + 0] * S[3 + r] + kinv[m * 3 + 1] * S[6 + r] + kinv[m * 3 + 2] * S[r];
        dP[r * 4 + 3] = S[9 + r];
    }
    kT_times_dP(Kb, a.k_is_f64, dP, dM);
    pose_to_M_vjp(a.poses + ((size_t)b * a.n_pose + job.pose_index[i]) * 6, a.rotation_mode, job.pose_inv[i], dM, g6);
#pragma unroll
    for (int m = 0; m < 6; ++m) J[k * 6 + m] = g6[m];
}

// first unit whose weight interval starts at or after `pos`, moved up to the next cut position of its column.
// Units are ordered (job, image, strip, combo, row); a row segment weighs combo_w[job][combo].
__host__ __device__ inline int photo_unit_of(const PhotoLaunch& p, int H, int pos) {
    int j = 0;
    while (j + 1 < p.a.n_jobs && pos >= (int)p.weight_start[j + 1]) ++j;
    const int rel = pos - (int)p.weight_start[j];
    const int nc = p.n_combos[j];
    const int scw = p.combo_cw[j][nc] * H;                  // weight of one (image, strip) super column
    const int sci = rel / scw, r2 = rel - sci * scw;
    int c = 0;
    while (c + 1 < nc && r2 >= p.combo_cw[j][c + 1] * H) ++c;
    const int w = p.combo_w[j][c];
    const int row = (r2 - p.combo_cw[j][c] * H + w - 1) / w;   // 0 .. H (H = first row of the next column)
    int u = p.unit_start[j] + (sci * nc + c) * H + row;
    const int al = p.combo_align[j][c];
    if (al > 1 && row < H) {
        const int r = row & (al - 1);                       // al is a power of two
        if (r) { const int up = al - r, left = H - row; u += up < left ? up : left; }
    }
    return u;
}

// ---------------------------------------------------------------------------------------------
// Per-pair constants, once per (job, image) pair instead of once per block of the main kernel: K^-1 (fp64
// adjugate, `transform.py:92`), the pose matrix (Rodrigues with the +1e-7 / Euler, optional rigid inverse,
// `pose_geometry.py:110-199`), P = K.[R|t], Q = P[:, :3].K^-1 composed in fp64 from the two fp32 matrices the
// reference multiplies a pixel by and rounded once, the packed tables of the combos and every base pointer already
// offset to the image.  One block of two warps per pair (two short dependent chains side by side: this code runs
// cold); the main kernel is launched as a programmatic dependent and copies the entries it needs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
photo_pairs_kernel(const __grid_constant__ PhotoLaunch p, int grad, int img_grad) {
    const plb_photo_args& a = p.a;
    asm volatile("griddepcontrol.launch_dependents;");
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ PairConst pc;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, plane = H * W;
    const int jb = pair / a.B, b = pair - jb * a.B;
    const plb_photo_job& job = a.jobs[jb];
    char* ws = (char*)a.workspace;
    float* gup = (float*)(ws + p.L.gup);
    float* ylow = (float*)(ws + p.L.ylow);
    const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
    const size_t img = (size_t)b * 3 * plane;
    for (int k = tid; k < (int)(sizeof(PairConst) / 4); k += 64) reinterpret_cast<float*>(&pc)[k] = 0.0f;
    __syncthreads();
    if (warp == 0) {
        if (lane == 1) {
            pc.tgt = job.tgt + img;
            pc.g_tgt = (grad && img_grad && job.g_tgt) ? job.g_tgt + img : nullptr;
            if (pc.g_tgt != nullptr && p.det)   // (a float* that holds the address of the image's int64 accumulators)
                pc.g_tgt = reinterpret_cast<float*>(reinterpret_cast<long long*>(ws + p.L.detacc) +
                                                    ((size_t)p.det_tgt[jb] * a.B * 3 * plane + img));
            pc.det_rho = (grad && img_grad && p.det) ? p.det_rho[jb] : 0.0f;
            pc.n_src = job.n_src; pc.n_scales = job.n_scales; pc.lowres = p.lowres[jb];
            pc.w_e = p.w_e[jb] * (a.upstream ? __ldg(a.upstream) : 1.0f);
        }
        if (lane >= 4 && lane < 4 + job.n_scales) {
            const int sc = lane - 4;
            const int dh = job.dh[sc], dw = job.dw[sc];
            const int sm = p.smode[jb][sc];
            pc.dh[sc] = dh; pc.dw[sc] = dw; pc.smode[sc] = sm;
            {
                const int f = H / dh;
                pc.fac[sc] = (f * dh == H && f * dw == W && f >= 2 && (f & (f - 1)) == 0) ? f : 0;
            }
            pc.sx[sc] = (float)dw / (float)W; pc.sy[sc] = (float)dh / (float)H;
            pc.disp[sc] = job.disp[sc] + (size_t)b * dh * dw;
            float* g = nullptr;
            if (grad && job.g_disp[sc] != nullptr) {
                if ((sm & 3) == PH_SM_FULL) g = job.g_disp[sc] + (size_t)b * plane;
                else if ((sm & 3) == PH_SM_LOWSCRATCH) g = gup + ((size_t)(jb * PLB_MAX_SCALES + sc) * a.B + b) * plane;
                else if ((sm & 3) == PH_SM_LOWFAST) g = ylow + p.L.ylow_off[jb][sc] + (size_t)b * 2 * dh * W;
            }
            pc.g_disp[sc] = g;
        }
        if (lane == 8) {
            float ki[9];
            kinv_f32(Kb, a.k_is_f64, ki);
#pragma unroll
            for (int k = 0; k < 9; ++k) pc.kinv[k] = ki[k];
        }
    } else {
        if (lane < job.n_src) {
            float M[12], P[12];
            pose_to_M(a.poses + ((size_t)b * a.n_pose + job.pose_index[lane]) * 6, a.rotation_mode, job.pose_inv[lane], M);
            k_times_M(Kb, a.k_is_f64, M, P);
#pragma unroll
            for (int r = 0; r < 3; ++r) pc.P[lane][r] = make_float4(P[r * 4], P[r * 4 + 1], P[r * 4 + 2], P[r * 4 + 3]);
            pc.src[lane] = job.src[lane] + img;
            pc.g_src[lane] = (grad && img_grad && job.g_src[lane]) ? job.g_src[lane] + img : nullptr;
            if (pc.g_src[lane] != nullptr && p.det)
                pc.g_src[lane] = reinterpret_cast<float*>(reinterpret_cast<long long*>(ws + p.L.detacc) +
                                                          ((size_t)p.det_src[jb][lane] * a.B * 3 * plane + img));
        }
    }
    __syncthreads();
    if (warp == 1) {
        const int n_src = job.n_src;
        // Q = P[:, :3] . fl32(K^-1): the exact product of the two fp32 matrices the reference multiplies a
        // pixel by (transform.py:92,137), rounded once; one lane per (source, row)
        if (lane < 3 * n_src) {
            const int i = lane / 3, r = lane - 3 * i;
            const float4 P = pc.P[i][r];
            float q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                q[c] = (float)((double)P.x * (double)pc.kinv[0 + c] + (double)P.y * (double)pc.kinv[3 + c] +
                               (double)P.z * (double)pc.kinv[6 + c]);
            pc.Q[i][r] = make_float4(q[0], q[1], q[2], P.w);
        }
        __syncwarp();
        // packed tables for the two halves of a combo: sources (2g, 2g+1), and every source with itself
        if (lane < PH_NT2 * 3) {
            const int tb = lane / 3, r = lane - tb * 3;
            const int i0 = tb < PLB_MAX_SRC / 2 ? 2 * tb : tb - PLB_MAX_SRC / 2;
            const int i1 = tb < PLB_MAX_SRC / 2 ? 2 * tb + 1 : i0;
            if (i1 < n_src) {
                const float4 q0 = pc.Q[i0][r], q1 = pc.Q[i1][r];
                pc.T2[tb][r][0] = make_float4(q0.x, q1.x, q0.z, q1.z);
                pc.T2[tb][r][1] = make_float4(q0.y, q1.y, q0.w, q1.w);
            }
        }
    }
    __syncthreads();
    constexpr int N4 = (int)(sizeof(PairConst) / sizeof(float4));
    float4* out = reinterpret_cast<float4*>(ws + p.L.pairs) + (size_t)pair * N4;
    for (int k = tid; k < N4; k += 64) out[k] = reinterpret_cast<const float4*>(&pc)[k];
}

template <bool GRAD, bool IMG_GRAD, bool HEAD, bool LOW, int CH, int CW>
#ifdef PH_MAXNREG
__global__ void __maxnreg__(PH_MAXNREG)
#else
__global__ void __launch_bounds__(PH_THREADS, PH_MIN_BLOCKS)
#endif
photo_l1_kernel(const __grid_constant__ PhotoLaunch p) {
    const plb_photo_args& a = p.a;
    // the finalize grid (programmatic dependent launch) may be scheduled from now on; it waits for
    // this grid to complete before it reads the records
    asm volatile("griddepcontrol.launch_dependents;");
    if (skip_launch(a.skip_if_unit)) return;
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 0 : 4);

    char* ws = (char*)a.workspace;
    float* records = (float*)(ws + p.L.records);
    float* gup = (float*)(ws + p.L.gup);
    float* ylow = (float*)(ws + p.L.ylow);

    const int H = CH > 0 ? CH : a.H, W = CW > 0 ? CW : a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane = H * W;

    __shared__ PairConst s_pc[2];
    __shared__ float s_rec[2][PH_WARPS][PH_NREC + 3];
    __shared__ int s_pair[2];
    __shared__ float s_stage[LOW ? PH_WARPS : 1][LOW ? 2 * PH_CHUNK * 32 : 1];   // per-lane depth / gradient slots of a chunk

    // ---- this warp's unit range: equal shares of the weighted unit list (32-bit maths) ---------
    auto pos_of = [&](int w) -> int { return w * p.share + min(w, p.share_rem); };
    auto pair_of = [&](int u) -> int {
        const int jb = (a.n_jobs > 1 && u >= p.unit_start[1]) ? 1 : 0;      // PLB_MAX_JOBS == 2
        return jb * a.B + (u - p.unit_start[jb]) / p.units_per_pair[jb];
    };
    // virtual block index: which share of the unit list this block owns.  The hardware deals blocks to the SMs round
    // robin, so block i and block i + SMs share an SM: with perm_sms set, the blocks resident on ONE SM own ADJACENT
    // shares - the same job and combo kind (one code path in the SM's instruction cache instead of all of them) and
    // neighbouring rows of one image (shared L1 lines).  Any mapping is correct; this one is only faster.
    const int vblk = p.perm_sms > 0 ? (int)(blockIdx.x % p.perm_sms) * p.perm_bps + (int)(blockIdx.x / p.perm_sms) : (int)blockIdx.x;
    const int gw = vblk * PH_WARPS + warp;
    const int blk_u0 = photo_unit_of(p, H, pos_of(vblk * PH_WARPS));
    const int blk_u1 = photo_unit_of(p, H, pos_of(vblk * PH_WARPS + PH_WARPS));
    const int u0 = photo_unit_of(p, H, pos_of(gw));
    const int u1 = photo_unit_of(p, H, pos_of(gw + 1));
    const bool empty_block = blk_u1 <= blk_u0;  // more blocks than work: still publishes (empty) records
    const int pairA = empty_block ? 0 : pair_of(blk_u0);
    const int pairB = empty_block ? 0 : pair_of(blk_u1 - 1);

    // ---- block prologue: context of the (at most two) pairs this block touches, copied from the table that
    //      photo_pairs_kernel (the primary of this programmatic dependent launch) filled - everything above ran
    //      while that grid was still working -------------------------------------------------------------------
    if (tid < 2) s_pair[tid] = empty_block ? -1 : ((tid == 0) ? pairA : (pairB != pairA ? pairB : -1));
    for (int k = tid; k < 2 * PH_WARPS * (PH_NREC + 3); k += PH_THREADS) (&s_rec[0][0][0])[k] = 0.0f;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!empty_block) {
        constexpr int N4 = (int)(sizeof(PairConst) / sizeof(float4));
        const float4* table = reinterpret_cast<const float4*>(ws + p.L.pairs);
        const int n_set = pairB != pairA ? 2 : 1;
        for (int k = tid; k < n_set * N4; k += PH_THREADS) {
            const int set = k / N4, q = k - set * N4;
            reinterpret_cast<float4*>(&s_pc[set])[q] = __ldcg(table + (size_t)(set == 0 ? pairA : pairB) * N4 + q);
        }
    }
    __syncthreads();
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 1 : 5);

    // ---- runs: consecutive rows of one 32-px strip of one combo of one (job, image) pair --------
    int u = u0;
#ifdef PLB_DEBUG_SKIP_UNITS
    u = u1;   // measurement only: fixed cost of prologue + epilogue
#endif
#pragma unroll 1
    while (u < u1) {
        const int jb = (a.n_jobs > 1 && u >= p.unit_start[1]) ? 1 : 0;
        const int loc = u - p.unit_start[jb];
        const int b = loc / p.units_per_pair[jb];
        const int rem = loc - b * p.units_per_pair[jb];
        const int col = rem / H;                              // (strip, combo) column of the image
        const int y = rem - col * H;
        const int strip = col / p.n_combos[jb];
        const PhotoCombo cb = p.combo[jb][col - strip * p.n_combos[jb]];
        const int rows = min(u1 - u, H - y);
        const int set = (jb * a.B + b == pairA) ? 0 : 1;
        const PairConst& pc = s_pc[set];
        float* rec = s_rec[set][warp];
        float* stage = s_stage[LOW ? warp : 0];
        if (cb.kind == PH_KIND_SRCPAIR) run_combo<GRAD, IMG_GRAD, 2, 1, HEAD, LOW, CH, CW>(a, pc, cb, strip, y, rows, lane, rec, stage);
        else if (cb.kind == PH_KIND_SCALEPAIR) run_combo<GRAD, IMG_GRAD, 2, 2, HEAD, LOW, CH, CW>(a, pc, cb, strip, y, rows, lane, rec, stage);
        else run_combo<GRAD, IMG_GRAD, 1, 1, HEAD, LOW, CH, CW>(a, pc, cb, strip, y, rows, lane, rec, stage);
        u += rows;
    }

    // ---- block records: fixed-order sum over the warps, one record per touched pair; the
    //      finalize kernel combines them (no fences, no tickets here) --------------------------
    __syncthreads();
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 2 : 6);
#ifdef PLB_DEBUG_TIMERS
    if (threadIdx.x == 0 && blockIdx.x < 2048) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g_dbg_blk[blockIdx.x] = t; unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); g_dbg_sm[blockIdx.x] = sm; }
#endif
    float* my_rec = records + (size_t)vblk * 2 * PH_REC_STRIDE;
    for (int k = tid; k < 2 * PH_REC_STRIDE; k += PH_THREADS) {
        const int set = k / PH_REC_STRIDE, c = k - set * PH_REC_STRIDE;
        float v;
        if (c == PH_REC_ID) {
            v = __int_as_float(s_pair[set]);
        } else if (c < PH_NREC) {
            v = 0.0f;
#pragma unroll
            for (int w = 0; w < PH_WARPS; ++w) v += s_rec[set][w][c];
        } else {
            v = 0.0f;
        }
        my_rec[k] = v;
    }
    if (tid < 2) {   // compact copy of (pair id, sum |diff|) for the loss reduction
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < PH_WARPS; ++w) v += s_rec[tid][w][PLB_MAX_SRC * 12];
        reinterpret_cast<float2*>(ws + p.L.lossrec)[vblk * 2 + tid] = make_float2(__int_as_float(s_pair[tid]), v);
    }
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 3 : 7);
}

// ---------------------------------------------------------------------------------------------
// Low-resolution scales, second half of the transposed upsample: the main kernel left, per (job, scale,
// contributor, image), two slots of [dh][W] rows already reduced over the full-resolution ROWS; one thread per
// low-resolution pixel gathers the full-resolution COLUMNS of its footprint (the align_corners=False weights of
// up_coord, the same the forward used), sums slots and contributors in a fixed order and applies the
// depth -> disparity (-> head) chain.  Deterministic; reads (2 / factor) of a plane per scale.
// ---------------------------------------------------------------------------------------------
struct LowMergeItem { int jb, s, first_block, n_contrib, f, col_chunks; };
struct LowMergeLaunch {
    int n_items;
    int total_blocks;
    LowMergeItem items[PLB_MAX_JOBS * PLB_MAX_SCALES];
};
constexpr int LM_THREADS = 256, LM_COLS = 64, LM_ROWS = LM_THREADS / LM_COLS;   // a block: 4 low-res rows x 64 columns

// Footprint of low-res column i for factor F: the F columns with x0 == i - 1 (weight l1) and the F columns with
// x0 == i (weight l0 = 1 - l1), l1 = ((x - F/2) mod F + 0.5) / F - a triangle; at the left border the first F/2
// columns clamp to x0 = 0 with weight 1, at the right border the last F/2 columns put both weights on column dw - 1
// (up_coord's rules; the values are dyadic rationals, identical to its fp32 arithmetic).  Interior columns: constant
// weights, immediate offsets, all loads in flight before the first use.
template <int F, int NC>
__device__ __forceinline__ float lowres_gather(const float* row0, size_t slot, size_t cstride, int i, int dw, int W) {
    constexpr int HF = F / 2;
    const int xs = i * F - HF;
    float acc = 0.0f;
    if (i > 0 && i < dw - 1) {
        const float* q = row0 + xs;
        float v[2 * F];
#pragma unroll
        for (int t = 0; t < 2 * F; ++t) {
            v[t] = __ldg(q + t) + __ldg(q + slot + t);
            if (NC > 1) v[t] += __ldg(q + cstride + t) + __ldg(q + cstride + slot + t);
        }
#pragma unroll
        for (int t = 0; t < 2 * F; ++t) {
            const float w = t < F ? ((float)t + 0.5f) * (1.0f / (float)F) : 1.0f - ((float)(t - F) + 0.5f) * (1.0f / (float)F);
            acc = fmaf(w, v[t], acc);
        }
    } else {
#pragma unroll 1
        for (int t = 0; t < 2 * F; ++t) {
            const int x = xs + t;
            if (x < 0 || x >= W) continue;
            float w = t < F ? ((float)t + 0.5f) * (1.0f / (float)F) : 1.0f - ((float)(t - F) + 0.5f) * (1.0f / (float)F);
            if (x < HF) w = (i == 0) ? 1.0f : 0.0f;
            if (i == dw - 1 && x >= W - HF) w = 1.0f;
            float v = __ldg(row0 + x) + __ldg(row0 + slot + x);
            if (NC > 1) v += __ldg(row0 + cstride + x) + __ldg(row0 + cstride + slot + x);
            acc = fmaf(w, v, acc);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(LM_THREADS)
photo_lowres_merge_kernel(const __grid_constant__ PhotoLaunch p, const __grid_constant__ LowMergeLaunch u) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    int it = 0;
#pragma unroll
    for (int k = 1; k < PLB_MAX_JOBS * PLB_MAX_SCALES; ++k)
        if (k < u.n_items && (int)blockIdx.x >= u.items[k].first_block) it = k;
    const LowMergeItem item = u.items[it];
    const plb_photo_job& job = a.jobs[item.jb];
    const int s = item.s, dh = job.dh[s], dw = job.dw[s], W = a.W;
    // block -> (group of LM_ROWS rows of the [B * dh] low-res rows, chunk of LM_COLS columns): two 32-bit divides per thread
    const unsigned local = blockIdx.x - (unsigned)item.first_block;
    const unsigned rg = local / (unsigned)item.col_chunks, cc = local - rg * (unsigned)item.col_chunks;
    const unsigned bj = rg * LM_ROWS + (threadIdx.x / LM_COLS);            // b * dh + j
    const int i = (int)(cc * LM_COLS + (threadIdx.x % LM_COLS));
    if (bj >= (unsigned)a.B * (unsigned)dh || i >= dw) return;
    const int b = (int)(bj / (unsigned)dh), j = (int)(bj - (unsigned)b * (unsigned)dh);
    const float* base = (const float*)((const char*)a.workspace + p.L.ylow) + p.L.ylow_off[item.jb][s];
    const size_t cstride = (size_t)a.B * 2 * dh * W;     // one contributor
    const float* row0 = base + ((size_t)b * 2 * dh + j) * W;
    const size_t slot = (size_t)dh * W;
    float acc;
    if (item.n_contrib > 1) {
        if (item.f == 2) acc = lowres_gather<2, 2>(row0, slot, cstride, i, dw, W);
        else if (item.f == 4) acc = lowres_gather<4, 2>(row0, slot, cstride, i, dw, W);
        else acc = lowres_gather<8, 2>(row0, slot, cstride, i, dw, W);
    } else {
        if (item.f == 2) acc = lowres_gather<2, 1>(row0, slot, cstride, i, dw, W);
        else if (item.f == 4) acc = lowres_gather<4, 1>(row0, slot, cstride, i, dw, W);
        else acc = lowres_gather<8, 1>(row0, slot, cstride, i, dw, W);
    }
    float chain = 1.0f;
    const size_t o = (size_t)bj * dw + i;
    if (a.input_is_depth != PLB_INPUT_DEPTH) {
        float d = __ldg(job.disp[s] + o), hc = 1.0f;
        if (a.input_is_depth == PLB_INPUT_LOGIT) {
            const float sg = 1.0f / (1.0f + expf(-d));
            d = fmaf(a.head_alpha, sg, a.head_beta);
            hc = a.head_alpha * sg * (1.0f - sg);
        }
        const float D = 1.0f / (a.disp_a * d + a.disp_b);
        chain = -a.disp_a * D * D * hc;
    }
    job.g_disp[s][o] = acc * chain;
}

// The same merge with the partial rows staged in shared memory (W a multiple of 4, 16-byte aligned planes, a row group
// that fits 16 KB).  The kernel above lets every low-resolution pixel gather its 2 F taps straight from the planes:
// lanes F floats apart, so one load instruction touches 8 - 32 sectors and the launch is bound by L1 sector traffic
// (69 us for 110 MB at batch 64).  Here a block owns whole rows: the slots (and contributors) of each are summed while
// they stream in through coalesced 128-bit loads - in the order the gather used, (a0 + a1) + (b0 + b1) - into one
// W-float row of shared memory (one pad word per 32, so the stride-F reads of the second phase hit distinct banks), and
// the pixels then take their taps from there with the same weights in the same order: bitwise the same gradients.
constexpr int LS_MAX_FLOATS = 4096;            // staged floats per block (16 KB + padding)
__device__ __forceinline__ int ls_pad(int x) { return x + (x >> 5); }

template <int F>
__device__ __forceinline__ float lowres_gather_smem(const float* v, int i, int dw, int W) {
    constexpr int HF = F / 2;
    const int xs = i * F - HF;
    float acc = 0.0f;
    if (i > 0 && i < dw - 1) {
#pragma unroll
        for (int t = 0; t < 2 * F; ++t) {
            const float w = t < F ? ((float)t + 0.5f) * (1.0f / (float)F) : 1.0f - ((float)(t - F) + 0.5f) * (1.0f / (float)F);
            acc = fmaf(w, v[ls_pad(xs + t)], acc);
        }
    } else {
#pragma unroll 1
        for (int t = 0; t < 2 * F; ++t) {
            const int x = xs + t;
            if (x < 0 || x >= W) continue;
            float w = t < F ? ((float)t + 0.5f) * (1.0f / (float)F) : 1.0f - ((float)(t - F) + 0.5f) * (1.0f / (float)F);
            if (x < HF) w = (i == 0) ? 1.0f : 0.0f;
            if (i == dw - 1 && x >= W - HF) w = 1.0f;
            acc = fmaf(w, v[ls_pad(x)], acc);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(LM_THREADS)
photo_lowres_merge_smem_kernel(const __grid_constant__ PhotoLaunch p, const __grid_constant__ LowMergeLaunch u) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ float sv[LS_MAX_FLOATS + LS_MAX_FLOATS / 32 + 8];
    int it = 0;
#pragma unroll
    for (int k = 1; k < PLB_MAX_JOBS * PLB_MAX_SCALES; ++k)
        if (k < u.n_items && (int)blockIdx.x >= u.items[k].first_block) it = k;
    const LowMergeItem item = u.items[it];
    const plb_photo_job& job = a.jobs[item.jb];
    const int s = item.s, dh = job.dh[s], dw = job.dw[s], W = a.W;
    const int rows_total = a.B * dh, lr = item.col_chunks;                 // (col_chunks holds the rows per block here)
    const int r0 = ((int)blockIdx.x - item.first_block) * lr;
    const int nr = min(lr, rows_total - r0);
    const float* base = (const float*)((const char*)a.workspace + p.L.ylow) + p.L.ylow_off[item.jb][s];
    const size_t cstride = (size_t)a.B * 2 * dh * W, slot = (size_t)dh * W;
    const int w4 = W >> 2, wp = ls_pad(W) + 1;                             // padded row pitch in shared memory
    // phase 1: v[r][x] = sum over slots (and contributors) of the partial planes
    for (int k = threadIdx.x; k < nr * w4; k += LM_THREADS) {
        const int r = k / w4, x = (k - r * w4) << 2;
        const int bj = r0 + r, b = bj / dh, j = bj - b * dh;
        const float* q = base + ((size_t)b * 2 * dh + j) * W + x;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(q)), a1 = __ldg(reinterpret_cast<const float4*>(q + slot));
        float4 v = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        if (item.n_contrib > 1) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(q + cstride)), b1 = __ldg(reinterpret_cast<const float4*>(q + cstride + slot));
            v.x += b0.x + b1.x; v.y += b0.y + b1.y; v.z += b0.z + b1.z; v.w += b0.w + b1.w;
        }
        float* d = sv + r * wp + ls_pad(x);                                // x is a multiple of 4: the four words share a pad group
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    // phase 2: the low-resolution pixels of the rows
    for (int k = threadIdx.x; k < nr * dw; k += LM_THREADS) {
        const int r = k / dw, i = k - r * dw;
        const float* v = sv + r * wp;
        const float acc = item.f == 2 ? lowres_gather_smem<2>(v, i, dw, W) : item.f == 4 ? lowres_gather_smem<4>(v, i, dw, W)
                                                                                         : lowres_gather_smem<8>(v, i, dw, W);
        float chain = 1.0f;
        const size_t o = (size_t)(r0 + r) * dw + i;
        if (a.input_is_depth != PLB_INPUT_DEPTH) {
            float d = __ldg(job.disp[s] + o), hc = 1.0f;
            if (a.input_is_depth == PLB_INPUT_LOGIT) {
                const float sg = 1.0f / (1.0f + expf(-d));
                d = fmaf(a.head_alpha, sg, a.head_beta);
                hc = a.head_alpha * sg * (1.0f - sg);
            }
            const float D = 1.0f / (a.disp_a * d + a.disp_b);
            chain = -a.disp_a * D * D * hc;
        }
        job.g_disp[s][o] = acc * chain;
    }
}

// ---------------------------------------------------------------------------------------------
// Finalize: one block per image.  For every job of the image: fixed-order sum of the block records
// of that (job, image) pair -> d loss / d P per source (through K^-1) -> K^T . dP -> (inverse) ->
// Rodrigues / Euler vjp -> the 6-vector; pose gradients of the image are written directly, the
// loss is the fp64 sum of the per-image partials, added in image order by the last block (one
// ticket per block).  Loss, pose and disparity gradients are bitwise repeatable.
// ---------------------------------------------------------------------------------------------
constexpr int PF_THREADS = 512;
constexpr int PF_GROUPS = PF_THREADS / 64;         // groups of 64 value lanes summing the records
constexpr int PF_COMBOS = PLB_MAX_JOBS * PLB_MAX_SRC;
constexpr int PF_PREP_T0 = PF_THREADS - 96;        // last three warps: 12 lanes per (job, source)
static_assert(PF_COMBOS * 12 <= 96, "three warps must cover every (job, source)");

__global__ void __launch_bounds__(PF_THREADS)
photo_finalize_kernel(const __grid_constant__ PhotoLaunch p, int want_grad) {
    const plb_photo_args& a = p.a;
    const int tid = threadIdx.x, b = blockIdx.x;
    __shared__ float s_part[PF_GROUPS][PLB_MAX_JOBS][64];
    __shared__ float s_red[PLB_MAX_JOBS][64];
    __shared__ double s_lpart[PF_THREADS];
    __shared__ float s_J[PF_COMBOS][72];
    __shared__ float s_g6[PF_COMBOS][6];
    __shared__ int s_col[PF_COMBOS];                   // pose column of each (job, source), -1 = unused
    const bool grads = want_grad && a.g_poses != nullptr;
    // guarded relaunch (backward with unit upstream): leave before the Jacobians, not after them.  The upstream scalars
    // were written before the MAIN kernel was launched (an ordinary launch), so they are readable ahead of the wait.
    if (skip_launch(a.skip_if_unit)) return;
    if (tid < PF_COMBOS) {
        const int jb = tid / PLB_MAX_SRC, i = tid - jb * PLB_MAX_SRC;
        s_col[tid] = (jb < a.n_jobs && i < a.jobs[jb].n_src) ? a.jobs[jb].pose_index[i] : -1;
    }
    DBG_STAMP(b == 0 && (tid == 0 || tid == PF_PREP_T0), tid == 0 ? 8 : 9);
    // ---- before the wait (inputs only): pose-chain Jacobians, 12 lanes per (job, source), on the last
    //      three warps.  Launched as a programmatic dependent of the main kernel, the block may become
    //      resident while the main kernel is still draining; then this overlaps its tail -----------------
    if (grads && tid >= PF_PREP_T0) {
        const int q = tid - PF_PREP_T0;
        const int combo = q / 12, k = q - combo * 12;
        const int jb = combo / PLB_MAX_SRC, i = combo - jb * PLB_MAX_SRC;
        if (combo < PF_COMBOS && jb < a.n_jobs && i < a.jobs[jb].n_src)
            photo_pose_jacobian(a, a.jobs[jb], b, i, k, s_J[combo]);
    }
    // ---- still before the wait (launch constants only): where this thread will read.  After the wait every
    //      load is base + compile-time offset under one count compare - the post-wait section is issue-bound
    //      (16 warps on one SM), so its address arithmetic is what the caller waits for -----------------------
    const float* records = (const float*)((const char*)a.workspace + p.L.records);
    const int c = tid & 63, grp = tid >> 6;
    constexpr int PF_FIRST = 12;                      // records per (thread, job) loaded in the unrolled batch
    const float* rp[PLB_MAX_JOBS];
    int rcnt[PLB_MAX_JOBS], rcnt_all[PLB_MAX_JOBS], rpair[PLB_MAX_JOBS];
#pragma unroll
    for (int jb = 0; jb < PLB_MAX_JOBS; ++jb) {
        rp[jb] = records; rcnt[jb] = 0; rcnt_all[jb] = 0; rpair[jb] = -1;
        if (jb < a.n_jobs) {
            const int pr = jb * a.B + b;
            // blocks whose range can overlap this pair (widened by one block on each side; records carry the pair id)
            const long long pw = (long long)p.strips * p.combo_cw[jb][p.n_combos[jb]] * a.H;   // weight of one image of the job
            const long long w0 = p.weight_start[jb] + (long long)b * pw;
            const long long w1 = w0 + pw;
            int k_lo = photo_warp_of(p, w0) / p.warps_per_block - 1;
            int k_hi = photo_warp_of(p, w1) / p.warps_per_block + 1;
            k_lo = max(k_lo, 0); k_hi = min(k_hi, p.grid - 1);
            const int n_rec = (k_hi - k_lo + 1) * 2;
            rp[jb] = records + ((size_t)k_lo * 2 + grp) * PH_REC_STRIDE;
            rcnt_all[jb] = n_rec > grp ? (n_rec - grp + PF_GROUPS - 1) / PF_GROUPS : 0;               // r_i = grp, grp + 8, ...
            rcnt[jb] = c < PH_NREC ? rcnt_all[jb] : 0;
            rpair[jb] = pr;
        }
    }
    constexpr int LQ = 6;                             // covers grids up to 6 * 512 / 2 = 1536 blocks
    const float2* lp = reinterpret_cast<const float2*>((const char*)a.workspace + p.L.lossrec) + tid;
    const int lcnt = (b == 0 && p.grid * 2 > tid) ? (p.grid * 2 - tid + PF_THREADS - 1) / PF_THREADS : 0;
    DBG_STAMP(b == 0 && tid == PF_PREP_T0, 10);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (skip_launch(a.skip_if_unit)) return;
    DBG_STAMP(b == 0 && (tid == 0 || tid == PF_PREP_T0), tid == 0 ? 11 : 12);
    {
        // ---- loss (block 0): every record's sum |diff|, weighted by its job; the loads are issued first
        //      so that they share one L2 round trip with the record loads below --------------------------
        float lval[LQ];
        int lid[LQ];
#pragma unroll
        for (int m = 0; m < LQ; ++m) {
            lid[m] = -1; lval[m] = 0.0f;
            if (m < lcnt) {
                const float2 q = __ldcg(lp + m * PF_THREADS);
                lid[m] = __float_as_int(q.x);
                lval[m] = q.y;
            }
        }
        // ---- fixed-order sum of the block records of each (job, image b) pair -------------------------
        float fv[PLB_MAX_JOBS][PF_FIRST];
        int ids[PLB_MAX_JOBS];
        const int lane = tid & 31;
#pragma unroll
        for (int jb = 0; jb < PLB_MAX_JOBS; ++jb) {
            // pair ids of this group's records: lane m loads the id of record m, shared by shuffle below
            const float* r = rp[jb] + min(lane, PF_FIRST - 1) * (PF_GROUPS * PH_REC_STRIDE) + PH_REC_ID;
            ids[jb] = (lane < rcnt_all[jb]) ? __float_as_int(__ldcg(r)) : -2;
#pragma unroll
            for (int m = 0; m < PF_FIRST; ++m) {
                fv[jb][m] = 0.0f;
                if (m < rcnt[jb]) fv[jb][m] = __ldcg(rp[jb] + m * (PF_GROUPS * PH_REC_STRIDE) + c);   // one aligned line per warp
            }
        }
        DBG_STAMP(b == 0 && tid == 0, 16);
#pragma unroll
        for (int jb = 0; jb < PLB_MAX_JOBS; ++jb) {
            float v = 0.0f;
#pragma unroll
            for (int m = 0; m < PF_FIRST; ++m) {
                const int id = __shfl_sync(0xffffffffu, ids[jb], m);
                v += (id == rpair[jb]) ? fv[jb][m] : 0.0f;
            }
            if (jb < a.n_jobs) {
                for (int m = PF_FIRST; m < rcnt[jb]; ++m) {       // grids with more than 8 * PF_FIRST / 2 blocks per pair
                    const float* r = rp[jb] + (size_t)m * (PF_GROUPS * PH_REC_STRIDE);
                    const int id = __float_as_int(__ldcg(r + PH_REC_ID));
                    const float val = __ldcg(r + c);
                    v += (id == rpair[jb]) ? val : 0.0f;
                }
                s_part[grp][jb][c] = v;
            }
        }
        if (b == 0) {
            double part = 0.0;
#pragma unroll
            for (int m = 0; m < LQ; ++m) {
                const float w = (lid[m] >= a.B) ? p.w_e[1] : p.w_e[0];     // PLB_MAX_JOBS == 2
                part += (lid[m] >= 0) ? (double)lval[m] * (double)w : 0.0;
            }
            // grids beyond LQ * PF_THREADS / 2 blocks (never launched today: <= 148 x 8)
            for (int q = tid + LQ * PF_THREADS; q < p.grid * 2; q += PF_THREADS) {
                const float2 r = __ldcg(lp + (q - tid));
                const int id = __float_as_int(r.x);
                const float w = (id >= a.B) ? p.w_e[1] : p.w_e[0];
                part += (id >= 0) ? (double)r.y * (double)w : 0.0;
            }
            s_lpart[tid] = part;
        }
    }
    DBG_STAMP(b == 0 && tid == 0, 13);
    __syncthreads();
    DBG_STAMP(b == 0 && tid == 0, 14);
    if (b == 0 && tid >= PF_THREADS - 32) {
        const int lane = tid - (PF_THREADS - 32);
        double part = 0.0;
#pragma unroll
        for (int m = 0; m < PF_THREADS / 32; ++m) part += s_lpart[lane + 32 * m];
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) part += __shfl_xor_sync(0xffffffffu, part, k);
        if (lane == 0 && a.loss != nullptr) *a.loss = (float)part;
    }
    if (!grads) return;
    if (tid < 64 * a.n_jobs) {
        const int jb = tid >> 6, c = tid & 63;
        float v = s_part[0][jb][c];
#pragma unroll
        for (int g = 1; g < PF_GROUPS; ++g) v += s_part[g][jb][c];
        s_red[jb][c] = v;
    }
    __syncthreads();
    // 6-vector of every (job, source): J^T . S, one thread per element
    if (tid < PF_COMBOS * 6) {
        const int combo = tid / 6, cc = tid - combo * 6;
        const int jb = combo / PLB_MAX_SRC, i = combo - jb * PLB_MAX_SRC;
        float g = 0.0f;
        if (s_col[combo] >= 0) {
            const float* S = &s_red[jb][i * 12];
#pragma unroll
            for (int m = 0; m < 12; ++m) g = fmaf(s_J[combo][m * 6 + cc], S[m], g);
        }
        s_g6[combo][cc] = g;
    }
    __syncthreads();
    // g_poses[b, col, cc] = sum over the (job, source) pairs that use pose column col (job order, then source order)
    for (int k = tid; k < a.n_pose * 6; k += PF_THREADS) {
        const int col = k / 6, cc = k - col * 6;
        float v = 0.0f;
#pragma unroll
        for (int combo = 0; combo < PF_COMBOS; ++combo)
            if (s_col[combo] == col) v += s_g6[combo][cc];
        a.g_poses[((size_t)b * a.n_pose + col) * 6 + cc] = v;
    }
    DBG_STAMP(b == 0 && tid == 0, 15);
}

#ifdef PLB_DEBUG_TIMERS
extern "C" int plb_debug_timers(unsigned long long* out32) {
    return (int)cudaMemcpyFromSymbol(out32, g_dbg_t, sizeof(unsigned long long) * 32);
}
extern "C" int plb_debug_block_ends(unsigned long long* out2048) {
    return (int)cudaMemcpyFromSymbol(out2048, g_dbg_blk, sizeof(unsigned long long) * 2048);
}
extern "C" int plb_debug_block_sms(unsigned int* out2048) {
    return (int)cudaMemcpyFromSymbol(out2048, g_dbg_sm, sizeof(unsigned int) * 2048);
}
#endif

// Transposed bilinear upsample (gather form, deterministic) + disp->depth chain:
// g_disp[s][b,j,i] = dD/dd * sum over the full-resolution pixels whose align_corners=False
// footprint touches low-res pixel (j,i).  Separable and streaming: one block owns `rows`
// consecutive low-res rows of one image (about UT_SPAN full-res rows) and a chunk of UT_OWN
// full-res columns (+ halo).
// Stage 1: every thread walks DOWN four adjacent full-res columns of the scratch plane ONCE (one
// 128-bit load per row, UT_UNROLL rows in flight) and adds each value into the two low-res rows it
// feeds - two sliding accumulators per column, emitted to shared memory when the low-res row index
// advances (it is monotone); the per-row weights come from a shared table built with the exact
// up_coord rule.  Stage 2: one thread per low-res column gathers the ~2f column sums of its footprint
// for all rows of the block (the column weights are read once, the row accumulators stay in registers).
// The launch is a compact list of (job, scale, image, row group, chunk) work items.
constexpr int UT_THREADS = 192;
constexpr int UT_CW = UT_THREADS * 4;          // full-res columns staged per block, four per thread
constexpr int UT_HALO = 32;                    // >= 1.5 * factor + 2 for factor <= 16
constexpr int UT_OWN = UT_CW - 2 * UT_HALO;    // full-res columns owned per block
#ifndef UT_MAXROWS_DEF
#define UT_MAXROWS_DEF 8
#endif
constexpr int UT_MAXROWS = UT_MAXROWS_DEF;     // low-res rows per block (stage-2 accumulators)
#ifndef UT_SPAN
#define UT_SPAN 48                             // full-res rows walked per block (the per-thread serial chain), halo included
#endif
#ifndef UT_UNROLL
#define UT_UNROLL 8
#endif
constexpr int UT_WIN = (UT_MAXROWS + 1) * 16 + 8;   // full-res rows feeding one block at factor <= 16

struct UpTItem { int jb, s, first_block, groups, chunks, rows; };
struct UpTLaunch {
    int n_items;
    int total_blocks;
    UpTItem items[PLB_MAX_JOBS * PLB_MAX_SCALES];
};

constexpr size_t UT_SMEM = sizeof(float) * ((size_t)UT_MAXROWS * UT_CW + 3 * UT_CW) + sizeof(float4) * UT_WIN;

#ifndef UT_MINBLOCKS
#define UT_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(UT_THREADS, UT_MINBLOCKS)
photo_upsample_T_kernel(const __grid_constant__ PhotoLaunch p, const __grid_constant__ UpTLaunch u) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    int it = 0;
    while (it + 1 < u.n_items && (int)blockIdx.x >= u.items[it + 1].first_block) ++it;
    const UpTItem item = u.items[it];
    const int local = blockIdx.x - item.first_block;
    const int chunk = local % item.chunks;
    const int grp_all = local / item.chunks;            // (image, row group)
    const int b = grp_all / item.groups, grp = grp_all - b * item.groups;
    const int jb = item.jb, s = item.s;
    const plb_photo_job& job = a.jobs[jb];
    const int dh = job.dh[s], dw = job.dw[s], H = a.H, W = a.W;
    const int j0 = grp * item.rows, j1 = min(j0 + item.rows, dh);   // low-res rows [j0, j1)
    const int xc0 = chunk * UT_OWN;
    const int tid = threadIdx.x;
    const float sx = (float)dw / (float)W, sy = (float)dh / (float)H;
    const float fy = (float)H / (float)dh, fx = (float)W / (float)dw;

    extern __shared__ __align__(16) unsigned char ut_smem[];
    float4* s_tab = reinterpret_cast<float4*>(ut_smem);        // per full-res row: weight to row y0, to row y0 + 1, y0
    float* s_col = reinterpret_cast<float*>(s_tab + UT_WIN);   // [UT_MAXROWS][UT_CW] column sums per low-res row
    float* s_l0 = s_col + UT_MAXROWS * UT_CW;                  // per staged column: weight to x0, to x0 + 1, x0
    float* s_l1 = s_l0 + UT_CW;
    int* s_x0 = reinterpret_cast<int*>(s_l1 + UT_CW);

    const float* gup = (const float*)((const char*)a.workspace + p.L.gup);
    const float* g = gup + ((size_t)(jb * PLB_MAX_SCALES + s) * a.B + b) * (size_t)H * W;
    // conservative full-res row window of the block; exact up_coord weights (zero outside the footprint)
    const int ylo = max((int)floorf(((float)j0 - 0.5f) * fy - 0.5f) - 1, 0);
    const int yhi = min((int)ceilf(((float)(j1 - 1) + 1.5f) * fy - 0.5f) + 1, H - 1);
    const int nwin = min(yhi - ylo + 1, UT_WIN);
    for (int t = tid; t < nwin; t += UT_THREADS) {
        int y0, y1; float ly0, ly1;
        up_coord(ylo + t, sy, dh, y0, y1, ly0, ly1);
        if (y1 == y0) { ly0 += ly1; ly1 = 0.0f; }      // clamped at the bottom border: both taps are y0
        s_tab[t] = make_float4(ly0, ly1, __int_as_float(y0), 0.0f);
    }
    {
        float4* z = reinterpret_cast<float4*>(s_col);
        for (int q = tid; q < (j1 - j0) * (UT_CW / 4); q += UT_THREADS) z[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const int xq = xc0 - UT_HALO + 4 * tid;            // first of this thread's four columns
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int x = xq + c;
        float l0 = 0.0f, l1 = 0.0f;
        int x0 = -1000000;
        if (x >= 0 && x < W) {
            int x1;
            up_coord(x, sx, dw, x0, x1, l0, l1);
            if (x1 == x0) { l0 += l1; l1 = 0.0f; }   // clamped at the right border: both taps are x0
        }
        s_x0[4 * tid + c] = x0; s_l0[4 * tid + c] = l0; s_l1[4 * tid + c] = l1;
    }
    __syncthreads();
    if (xq + 3 >= 0 && xq < W) {
        const bool vec = ((W & 3) == 0) && xq >= 0 && xq + 3 < W;   // rows are 16-byte aligned when W % 4 == 0
        const float* row = g + ((size_t)ylo * W + xq);
        int jcur = __float_as_int(s_tab[0].z);    // a_lo belongs to low-res row jcur, a_hi to jcur + 1
        float4 a_lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), a_hi = a_lo;
        float4* out = reinterpret_cast<float4*>(s_col) + tid;
        auto emit = [&](int j, const float4& v) { if (j >= j0 && j < j1) out[(j - j0) * (UT_CW / 4)] = v; };
        auto step = [&](int t, const float4& v) {
            const float4 e = s_tab[t];
            const int y0 = __float_as_int(e.z);
            if (y0 > jcur) {                           // y0 advances by <= 1 per row (block-uniform branch)
                emit(jcur, a_lo);
                a_lo = a_hi; a_hi = make_float4(0.0f, 0.0f, 0.0f, 0.0f); ++jcur;
            }
            a_lo.x = fmaf(e.x, v.x, a_lo.x); a_lo.y = fmaf(e.x, v.y, a_lo.y);
            a_lo.z = fmaf(e.x, v.z, a_lo.z); a_lo.w = fmaf(e.x, v.w, a_lo.w);
            a_hi.x = fmaf(e.y, v.x, a_hi.x); a_hi.y = fmaf(e.y, v.y, a_hi.y);
            a_hi.z = fmaf(e.y, v.z, a_hi.z); a_hi.w = fmaf(e.y, v.w, a_hi.w);
        };
        if (vec) {
            int t = 0;
#pragma unroll 1
            for (; t + UT_UNROLL <= nwin; t += UT_UNROLL) {
                float4 v[UT_UNROLL];
#pragma unroll
                for (int q = 0; q < UT_UNROLL; ++q) { v[q] = __ldcs(reinterpret_cast<const float4*>(row)); row += W; }
#pragma unroll
                for (int q = 0; q < UT_UNROLL; ++q) step(t + q, v[q]);
            }
            for (; t < nwin; ++t) { step(t, __ldcs(reinterpret_cast<const float4*>(row))); row += W; }
        } else {
            const bool in0 = xq >= 0 && xq < W, in1 = xq + 1 >= 0 && xq + 1 < W;
            const bool in2 = xq + 2 >= 0 && xq + 2 < W, in3 = xq + 3 >= 0 && xq + 3 < W;
#pragma unroll 2
            for (int t = 0; t < nwin; ++t, row += W) {
                float4 v;
                v.x = in0 ? __ldcs(row) : 0.0f; v.y = in1 ? __ldcs(row + 1) : 0.0f;
                v.z = in2 ? __ldcs(row + 2) : 0.0f; v.w = in3 ? __ldcs(row + 3) : 0.0f;
                step(t, v);
            }
        }
        emit(jcur, a_lo);
        emit(jcur + 1, a_hi);
    }
    __syncthreads();
    // low-res columns whose centre of mass lies in the owned chunk: i in [i_lo, i_hi)
    const int i_lo = (int)ceilf((float)xc0 * sx - 1e-4f);
    const int i_hi = min((int)ceilf((float)min(xc0 + UT_OWN, W) * sx - 1e-4f), dw);
    const int nr = j1 - j0;
    for (int i = i_lo + tid; i < i_hi; i += UT_THREADS) {
        const int xlo = max((int)floorf(((float)i - 0.5f) * fx - 0.5f) - 1, 0);
        const int xhi = min((int)ceilf(((float)i + 1.5f) * fx - 0.5f) + 1, W - 1);
        float acc[UT_MAXROWS];
#pragma unroll
        for (int r = 0; r < UT_MAXROWS; ++r) acc[r] = 0.0f;
        const int k_lo = max(xlo - (xc0 - UT_HALO), 0), k_hi = min(xhi - (xc0 - UT_HALO), UT_CW - 1);
        for (int k = k_lo; k <= k_hi; ++k) {
            const int x0 = s_x0[k];
            const float w = (x0 == i ? s_l0[k] : 0.0f) + (x0 + 1 == i ? s_l1[k] : 0.0f);
            if (w != 0.0f) {
#pragma unroll
                for (int r = 0; r < UT_MAXROWS; ++r)
                    if (r < nr) acc[r] = fmaf(w, s_col[r * UT_CW + k], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < UT_MAXROWS; ++r) {
            if (r < nr) {
                float chain = 1.0f;
                const size_t o = (size_t)b * dh * dw + (size_t)(j0 + r) * dw + i;
                if (a.input_is_depth != PLB_INPUT_DEPTH) {
                    float d = __ldg(job.disp[s] + o), hc = 1.0f;
                    if (a.input_is_depth == PLB_INPUT_LOGIT) {
                        const float sg = 1.0f / (1.0f + expf(-d));
                        d = fmaf(a.head_alpha, sg, a.head_beta);
                        hc = a.head_alpha * sg * (1.0f - sg);
                    }
                    const float D = 1.0f / (a.disp_a * d + a.disp_b);
                    chain = -a.disp_a * D * D * hc;
                }
                job.g_disp[s][o] = acc[r] * chain;
            }
        }
    }
}

int validate_photo(const plb_photo_args* a) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 2 || a->W < 2 || a->n_jobs < 1 || a->n_jobs > PLB_MAX_JOBS || a->n_pose < 1)
        return PLB_EINVAL;
    if ((long long)a->H * a->W * 3 >= (1LL << 31)) return PLB_EINVAL;
    // the weighted unit list is indexed with 32-bit integers: <= PH_MAX_COMBOS combos per strip, <= 64 weight units each
    if ((long long)a->H * ((a->W + 31) / 32) * a->B * a->n_jobs * PH_MAX_COMBOS * 64 >= (1LL << 31)) return PLB_EINVAL;
    if (a->rotation_mode != PLB_ROT_AXISANGLE && a->rotation_mode != PLB_ROT_EULER) return PLB_EINVAL;
    if (a->poses == nullptr || a->K == nullptr || a->loss == nullptr) return PLB_ENULL;
    for (int j = 0; j < a->n_jobs; ++j) {
        const plb_photo_job& job = a->jobs[j];
        if (job.n_src < 1 || job.n_src > PLB_MAX_SRC || job.n_scales < 1 || job.n_scales > PLB_MAX_SCALES)
            return PLB_EINVAL;
        if (job.tgt == nullptr) return PLB_ENULL;
        for (int i = 0; i < job.n_src; ++i) {
            if (job.src[i] == nullptr) return PLB_ENULL;
            if (job.pose_index[i] < 0 || job.pose_index[i] >= a->n_pose) return PLB_EINVAL;
        }
        for (int s = 0; s < job.n_scales; ++s) {
            if (job.disp[s] == nullptr) return PLB_ENULL;
            if (job.dh[s] < 1 || job.dw[s] < 1 || job.dh[s] > a->H || job.dw[s] > a->W) return PLB_EINVAL;
            if (job.dw[s] * 16 < a->W || job.dh[s] * 16 < a->H) return PLB_EINVAL;  // upsampling factor <= 16
        }
    }
    if (a->workspace == nullptr) return PLB_EWORKSPACE;
    if (a->workspace_bytes < photo_layout(*a).total) return PLB_EWORKSPACE;
    return PLB_OK;
}

int photo_upsample_T_launch(const PhotoLaunch& p, cudaStream_t st) {
    const plb_photo_args* a = &p.a;
    static bool attr_set[PLB_MAX_DEVICES] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {                                    // per device: a process may drive several GPUs
        const cudaError_t e = cudaFuncSetAttribute(photo_upsample_T_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UT_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = true;
    }
    {
        UpTLaunch u;
        u.n_items = 0;
        u.total_blocks = 0;
        // the blocks of a coarse scale walk more full-resolution rows: they go first (no long tail)
        for (int pass = 0; pass < PLB_MAX_JOBS * PLB_MAX_SCALES; ++pass) {
            int bj = -1, bs = -1;
            for (int j = 0; j < a->n_jobs; ++j)
                for (int s = 0; s < a->jobs[j].n_scales; ++s) {
                    const plb_photo_job& job = a->jobs[j];
                    if (photo_scale_mode(*a, j, s) != PH_SM_LOWSCRATCH) continue;
                    bool taken = false;
                    for (int k = 0; k < u.n_items; ++k) taken = taken || (u.items[k].jb == j && u.items[k].s == s);
                    if (taken) continue;
                    if (bj < 0 || job.dh[s] < a->jobs[bj].dh[bs]) { bj = j; bs = s; }
                }
            if (bj < 0) break;
            const plb_photo_job& job = a->jobs[bj];
            UpTItem& it = u.items[u.n_items++];
            it.jb = bj; it.s = bs; it.first_block = u.total_blocks;
            int rows = (int)((float)UT_SPAN * (float)job.dh[bs] / (float)a->H) - 1;
            it.rows = rows < 1 ? 1 : (rows > UT_MAXROWS ? UT_MAXROWS : rows);
            it.groups = (job.dh[bs] + it.rows - 1) / it.rows;
            it.chunks = (a->W + UT_OWN - 1) / UT_OWN;
            u.total_blocks += it.groups * it.chunks * a->B;
        }
        photo_upsample_T_kernel<<<u.total_blocks, UT_THREADS, UT_SMEM, st>>>(p, u);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    return PLB_OK;
}

// ---- per-device launch constants (a process may drive several GPUs) ---------------------------------------
static int sm_count() {
    static int cached[PLB_MAX_DEVICES] = {};
    const int dev = current_device();
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = 148;
        }
        cached[dev] = n;
    }
    return cached[dev];
}

template <bool GRAD, bool IMG, bool HEAD, bool LOW, int CH, int CW>
static int launch_variant(const PhotoLaunch& p, int query_only, cudaStream_t st) {
    if (query_only) {
        static int cached[PLB_MAX_DEVICES] = {};
        const int dev = current_device();
        if (cached[dev] == 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photo_l1_kernel<GRAD, IMG, HEAD, LOW, CH, CW>, PH_THREADS, 0) != cudaSuccess || n < 1) {
                (void)cudaGetLastError();
                n = 2;
            }
            cached[dev] = n;
        }
        return cached[dev];
    }
    // programmatic dependent of photo_pairs_kernel: this grid's own set-up overlaps the tail of that one
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(PH_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = PH_USE_PDL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, photo_l1_kernel<GRAD, IMG, HEAD, LOW, CH, CW>, p);
}

// the kernel variant of a launch: gradients / image gradients / folded disparity head / any low-resolution scale,
// and - for the hot configuration (gradients, no image gradients, no head) - the image size as a compile-time
// constant for the KITTI shapes the reference trains at (every tap address is then base + immediate)
static int dispatch_variant(const PhotoLaunch& p, bool img_grad, bool head, bool low, int query_only, cudaStream_t st) {
    const plb_photo_args& a = p.a;
#define PLB_GO(G, I, HD, CH, CW) (low ? launch_variant<G, I, HD, true, CH, CW>(p, query_only, st) : launch_variant<G, I, HD, false, CH, CW>(p, query_only, st))
    if (!a.want_grad) return head ? PLB_GO(false, false, true, 0, 0) : PLB_GO(false, false, false, 0, 0);
    if (img_grad) return PLB_GO(true, true, false, 0, 0);
    if (head) return PLB_GO(true, false, true, 0, 0);
#ifndef PH_NO_FIXED_DIMS
    if (a.H == 192 && a.W == 640) return PLB_GO(true, false, false, 192, 640);
    if (a.H == 320 && a.W == 1024) return PLB_GO(true, false, false, 320, 1024);
#endif
    return PLB_GO(true, false, false, 0, 0);
#undef PLB_GO
}

// combos of one job: source pairs at every scale, the odd source at pairs of scales, at most one single sample
static int build_combos(const plb_photo_job& job, PhotoCombo* out, int* contrib /* [MAX_SCALES] */) {
    int n = 0;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) contrib[s] = 0;
    for (int g = 0; g + 1 < job.n_src; g += 2)
        for (int s = 0; s < job.n_scales; ++s) {
            PhotoCombo c = {};
            c.kind = PH_KIND_SRCPAIR; c.k0 = (unsigned char)g; c.s0 = c.s1 = (unsigned char)s;
            c.c0 = c.c1 = (unsigned char)contrib[s]++;
            out[n++] = c;
        }
    if (job.n_src & 1) {
        const int k = job.n_src - 1;
        int s = 0;
        for (; s + 1 < job.n_scales; s += 2) {
            PhotoCombo c = {};
            c.kind = PH_KIND_SCALEPAIR; c.k0 = (unsigned char)k; c.s0 = (unsigned char)s; c.s1 = (unsigned char)(s + 1);
            c.c0 = (unsigned char)contrib[s]++; c.c1 = (unsigned char)contrib[s + 1]++;
            out[n++] = c;
        }
        if (s < job.n_scales) {
            PhotoCombo c = {};
            c.kind = PH_KIND_SINGLE; c.k0 = (unsigned char)k; c.s0 = c.s1 = (unsigned char)s;
            c.c0 = c.c1 = (unsigned char)contrib[s]++;
            out[n++] = c;
        }
    }
    return n;
}

// zero-fill of a gradient map two combos add into - skipped, like every launch of a guarded backward relaunch,
// when all upstream gradients are 1 (the values written by the forward pass then stand)
__global__ void guarded_zero_kernel(float* g, size_t n, const float* skip0, const float* skip1) {
    const float* const flags[2] = {skip0, skip1};
    if (skip_launch(flags)) return;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n && (reinterpret_cast<size_t>(g) & 15) == 0) {
        *reinterpret_cast<float4*>(g + i) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
        for (size_t k = i; k < n && k < i + 4; ++k) g[k] = 0.0f;
    }
}

// Deterministic image gradients, second half: fixed-point accumulators -> the caller's float buffers (ACCUMULATED,
// one writer per element), accumulators re-zeroed for the next call.
struct DetConvert { float* out[PH_DET_MAX]; int n; float scale; };
__global__ void __launch_bounds__(256)
photo_det_convert_kernel(long long* __restrict__ acc, const __grid_constant__ DetConvert d, size_t per, const float* upstream) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= per) return;
    const float scale = d.scale * (upstream ? __ldg(upstream) : 1.0f);
#pragma unroll
    for (int u = 0; u < PH_DET_MAX; ++u) {
        if (u >= d.n) break;
        long long* q = acc + (size_t)u * per + i;
        const long long v = *q;
        if (v != 0) {
            d.out[u][i] += (float)((double)v * (double)scale);
            *q = 0;
        }
    }
}

int photo_det_convert_launch(const plb_photo_args& a, const PhotoLayout& L, float unit, cudaStream_t st) {
    const PhotoDetSlots dslots = photo_det_slots(a);
    if (dslots.n == 0) return PLB_OK;
    DetConvert d;
    d.n = dslots.n;
    for (int k = 0; k < PH_DET_MAX; ++k) d.out[k] = dslots.out[k];
    d.scale = unit;
    const size_t per = (size_t)a.B * 3 * a.H * a.W;
    photo_det_convert_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<long long*>((char*)a.workspace + L.detacc), d, per, a.upstream);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

static int photo_lowres_merge_launch(const PhotoLaunch& p, cudaStream_t st) {
    const plb_photo_args& a = p.a;
    LowMergeLaunch u;
    u.n_items = 0;
    u.total_blocks = 0;
    // rows staged in shared memory when they are whole 128-bit packets and a row fits the stage
    const float* ylow = (const float*)((const char*)a.workspace + p.L.ylow);
    bool staged = (a.W & 3) == 0 && a.W <= LS_MAX_FLOATS && ((uintptr_t)ylow & 15) == 0;
    for (int j = 0; j < a.n_jobs && staged; ++j)
        for (int s = 0; s < a.jobs[j].n_scales; ++s)
            if ((p.smode[j][s] & 3) == PH_SM_LOWFAST && (p.L.ylow_off[j][s] & 3)) staged = false;
#ifdef PH_MERGE_GATHER
    staged = false;
#endif
    const int lr = staged ? (LS_MAX_FLOATS / a.W > 8 ? 8 : LS_MAX_FLOATS / a.W) : 0;
    for (int j = 0; j < a.n_jobs; ++j)
        for (int s = 0; s < a.jobs[j].n_scales; ++s) {
            if ((p.smode[j][s] & 3) != PH_SM_LOWFAST) continue;
            LowMergeItem& it = u.items[u.n_items++];
            it.jb = j; it.s = s; it.first_block = u.total_blocks;
            it.n_contrib = (p.smode[j][s] & 4) ? 2 : 1;
            it.f = a.W / a.jobs[j].dw[s];
            const long long rows = (long long)a.B * a.jobs[j].dh[s];                      // < 2^31: validate_photo
            if (staged) {
                it.col_chunks = lr;                                                        // rows per block
                u.total_blocks += (int)((rows + lr - 1) / lr);
            } else {
                it.col_chunks = (a.jobs[j].dw[s] + LM_COLS - 1) / LM_COLS;
                u.total_blocks += (int)((rows + LM_ROWS - 1) / LM_ROWS) * it.col_chunks;
            }
        }
    if (u.n_items == 0) return PLB_OK;
    if (staged) photo_lowres_merge_smem_kernel<<<u.total_blocks, LM_THREADS, 0, st>>>(p, u);
    else photo_lowres_merge_kernel<<<u.total_blocks, LM_THREADS, 0, st>>>(p, u);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int photo_l1_launch(const plb_photo_args* a, cudaStream_t st) {
    if (a != nullptr && a->n_jobs >= 1 && a->n_jobs <= PLB_MAX_JOBS && a->jobs[0].mode == PLB_PHOTO_MIN_REPROJ)
        return photo_min_launch(a, st);
    int rc = validate_photo(a);
    if (rc != PLB_OK) return rc;
    PhotoLaunch p;
    p.a = *a;
    p.L = photo_layout(*a);
    rc = photo_workspace_prepare(*a, p.L, st);
    if (rc != PLB_OK) return rc;
    bool img_grad = false, low = false, lowfast = false;
    for (int j = 0; j < a->n_jobs; ++j) {
        if (a->jobs[j].g_tgt) img_grad = true;
        for (int i = 0; i < a->jobs[j].n_src; ++i)
            if (a->jobs[j].g_src[i]) img_grad = true;
    }
    // the disparity head folded in (PLB_INPUT_LOGIT) is its own set of kernel variants: the default variants keep
    // their register budget
    const bool head = a->input_is_depth == PLB_INPUT_LOGIT;
    if (head && img_grad) return PLB_EINVAL;       // image gradients: disparity / depth inputs only
    p.strips = (a->W + 31) / 32;
    p.n_pairs = a->n_jobs * a->B;
    int max_src = 0;
    for (int j = 0; j < a->n_jobs; ++j) max_src = a->jobs[j].n_src > max_src ? a->jobs[j].n_src : max_src;
    long long wsum = 0, pair_w_min = 1LL << 62;
    int usum = 0;
    for (int j = 0; j < PLB_MAX_JOBS; ++j) {
        p.weight_start[j] = wsum;
        p.unit_start[j] = usum;
        p.units_per_pair[j] = 1;
        for (int c = 0; c < PH_MAX_COMBOS; ++c) { p.combo_w[j][c] = 1; p.combo_cw[j][c] = c; p.combo_align[j][c] = 1; }
        p.combo_cw[j][PH_MAX_COMBOS] = PH_MAX_COMBOS;
        p.n_combos[j] = 0;
        p.w_e[j] = 0.0f;
        p.lowres[j] = 0;
        for (int s = 0; s < PLB_MAX_SCALES; ++s) p.smode[j][s] = PH_SM_FULL;
        if (j < a->n_jobs) {
            const plb_photo_job& job = a->jobs[j];
            int contrib[PLB_MAX_SCALES];
            p.n_combos[j] = build_combos(job, p.combo[j], contrib);
            // cost model of a row segment: a packed combo costs less than two single samples on the scalar pipe; every
            // low-resolution depth stream adds its pre- / post-pass
            p.combo_cw[j][0] = 0;
            for (int c = 0; c < p.n_combos[j]; ++c) {
                const PhotoCombo& cb = p.combo[j][c];
                int w = cb.kind == PH_KIND_SINGLE ? PH_W_ODD : PH_W_PAIR;
                if (photo_scale_mode(*a, j, cb.s0) != PH_SM_FULL) w += PH_W_LOW;
                if (cb.kind == PH_KIND_SCALEPAIR) {
                    // (measured, profiles/README.md: with at most two sources per job a scale pair is cheaper relative
                    // to a source pair than in the three- / four-source kernel variants)
                    w += max_src <= 2 ? PH_W_SPAIR2 : PH_W_SPAIR;
                    if (photo_scale_mode(*a, j, cb.s1) != PH_SM_FULL) w += PH_W_LOW;
                }
                p.combo_w[j][c] = w;
                p.combo_cw[j][c + 1] = p.combo_cw[j][c] + w;
                // a chunk boundary or a cut may cross the footprint of a low-res row (<= 2.5 x factor rows) at most once
                int al = 1;
                if (photo_scale_mode(*a, j, cb.s0) == PH_SM_LOWFAST) al = 2 * (a->H / job.dh[cb.s0]);
                if (cb.kind == PH_KIND_SCALEPAIR && photo_scale_mode(*a, j, cb.s1) == PH_SM_LOWFAST && 2 * (a->H / job.dh[cb.s1]) > al)
                    al = 2 * (a->H / job.dh[cb.s1]);
                p.combo_align[j][c] = al;
            }
            p.units_per_pair[j] = p.strips * p.n_combos[j] * a->H;
            const long long pw = (long long)p.combo_cw[j][p.n_combos[j]] * p.strips * a->H;
            if (pw < pair_w_min) pair_w_min = pw;
            wsum += pw * a->B;
            usum += p.units_per_pair[j] * a->B;
            p.w_e[j] = job.term_weight / (3.0f * (float)a->B * (float)a->H * (float)a->W);
            for (int s = 0; s < job.n_scales; ++s) {
                const int sm = photo_scale_mode(*a, j, s);
                if (sm != PH_SM_FULL) { p.lowres[j] |= 1 << s; low = true; }
                if (sm == PH_SM_LOWFAST) lowfast = true;
                const bool shared = a->want_grad && job.g_disp[s] != nullptr && contrib[s] > 1;
                p.smode[j][s] = (unsigned char)(sm | (shared ? 4 : 0));
                if (shared && sm != PH_SM_LOWFAST) {
                    // two combos add into this map (two addends: the sum does not depend on their order): zero it first
                    float* g; size_t n;
                    if (sm == PH_SM_FULL) { g = job.g_disp[s]; n = (size_t)a->B * a->H * a->W; }
                    else { g = (float*)((char*)a->workspace + p.L.gup) + (size_t)(j * PLB_MAX_SCALES + s) * a->B * a->H * a->W; n = (size_t)a->B * a->H * a->W; }
                    guarded_zero_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(g, n, p.a.skip_if_unit[0], p.a.skip_if_unit[1]);
                    ++g_launches;
                    PLB_CHECK_LAUNCH();
                }
            }
        }
    }
    p.weight_start[PLB_MAX_JOBS] = wsum;
    p.unit_start[PLB_MAX_JOBS] = usum;
    // deterministic image gradients: accumulator slots and the weight of every job relative to the largest one
    const PhotoDetSlots dslots = photo_det_slots(*a);
    p.det = (img_grad && dslots.n > 0) ? 1 : 0;
    float w_max = 0.0f;
    for (int j = 0; j < a->n_jobs; ++j) w_max = p.w_e[j] > w_max ? p.w_e[j] : w_max;
    for (int j = 0; j < PLB_MAX_JOBS; ++j) {
        p.det_tgt[j] = dslots.tgt[j];
        for (int i = 0; i < PLB_MAX_SRC; ++i) p.det_src[j][i] = dslots.src[j][i];
        p.det_rho[j] = (j < a->n_jobs && w_max > 0.0f) ? p.w_e[j] / w_max : 0.0f;
    }

    const int bps = dispatch_variant(p, img_grad, head, low, 1, st);
    // more blocks than are resident: the hardware hands a waiting block to whichever SM retires one first, which evens
    // out the finishing times of the SMs (blocks of equal work do not take equal time)
    int sms = sm_count();
    if (a->sm_limit > 0 && a->sm_limit < sms) sms = a->sm_limit;     // leave SMs to kernels running beside this one
    long long grid = (long long)sms * bps * PH_GRID_MULT;
    // a block's weight range must not exceed the lightest pair, so that it touches at most two pairs
    const long long need = (wsum + pair_w_min - 1) / pair_w_min + 1;
    if (grid > usum) grid = usum;                  // tiny problems: no more blocks than units ...
    if (grid < need) grid = need;                  // ... but never so few that a block spans three pairs
    if (grid > photo_max_grid(*a)) grid = photo_max_grid(*a);
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.perm_sms = p.perm_bps = 0;
#ifdef PH_SM_GROUP   // measured (profiles/README.md): c2 192 vs 185 us, c3 461 vs 451, c5 947 vs 951 - off
    if (grid == (long long)sms * bps && bps > 1) { p.perm_sms = sms; p.perm_bps = bps; }
#endif
    p.warps_per_block = PH_WARPS;
    p.n_warps = p.grid * p.warps_per_block;
    p.share = (int)(wsum / p.n_warps);
    p.share_rem = (int)(wsum % p.n_warps);

    photo_pairs_kernel<<<p.n_pairs, 64, 0, st>>>(p, (int)a->want_grad, (int)img_grad);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    {
        const int e = dispatch_variant(p, img_grad, head, low, 0, st);
        if (e != 0) return e;
    }
    ++g_launches;
    PLB_CHECK_LAUNCH();
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(a->B);
        cfg.blockDim = dim3(PF_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = PH_USE_PDL;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, photo_finalize_kernel, p, (int)a->want_grad);
        if (e != cudaSuccess) return (int)e;
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    if (p.det) {
        const int rc2 = photo_det_convert_launch(*a, p.L, w_max / PH_DET_ONE, st);
        if (rc2 != PLB_OK) return rc2;
    }
    if (a->want_grad) {
        if (lowfast) {
            const int rc2 = photo_lowres_merge_launch(p, st);
            if (rc2 != PLB_OK) return rc2;
        }
        if (photo_has_lowres_grad(*a)) {
            const int rc2 = photo_upsample_T_launch(p, st);
            if (rc2 != PLB_OK) return rc2;
        }
    }
    return PLB_OK;
}

size_t photo_workspace_bytes(const plb_photo_args* a) { return photo_layout(*a).total; }

// The workspace cleans up after itself for ONE layout: tickets go back to zero, the fixed-point accumulators of the
// deterministic mode are re-zeroed by their conversion kernel, everything else is written before it is read.  A call
// with a DIFFERENT layout (another batch, image size, mode, set of gradients) finds the regions it expects to be zero
// somewhere else - possibly under the records or statistics an earlier call left - so the library remembers the
// layout every workspace pointer was last used with and clears the buffer when it changes (the first use of a pointer
// trusts the caller's zero fill).
int photo_workspace_prepare(const plb_photo_args& a, const PhotoLayout& L, cudaStream_t st) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix(L.tickets); mix(L.records); mix(L.lossrec); mix(L.ws_pose); mix(L.ws_loss); mix(L.gup); mix(L.pairs); mix(L.ylow);
    for (int j = 0; j < PLB_MAX_JOBS; ++j)
        for (int s = 0; s < PLB_MAX_SCALES; ++s) mix(L.ylow_off[j][s]);
    mix(L.pm_thr); mix(L.pm_stat); mix(L.detacc); mix((uint64_t)L.det_n); mix(L.total);
    static std::mutex mu;
    static std::unordered_map<const void*, uint64_t> seen;
    std::lock_guard<std::mutex> lock(mu);
    auto it = seen.find(a.workspace);
    if (it == seen.end()) { seen.emplace(a.workspace, h); return PLB_OK; }
    if (it->second == h) return PLB_OK;
    it->second = h;
    return (int)cudaMemsetAsync(a.workspace, 0, L.total, st);
}

}  // namespace plb
