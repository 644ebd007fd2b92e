// Fused photometric reprojection loss (live mode: warp + L1 mean), forward and
// gradients in one pass.  See DESIGN.md "photo_l1_kernel".
//
// Work decomposition.  The unit of work is one ROW SEGMENT: 32 consecutive
// pixels of one row of one target image of one job (direction), processed by
// one warp: every global access of the warp - target pixels, disparity, the
// 2x2x3 bilinear taps of each source - is a (nearly) contiguous 128-byte piece
// of a planar NCHW row.  Units are ordered (job, image, 32-px column strip, row),
// so a warp that walks its units moves DOWN a strip and re-uses the source rows
// it has just pulled into L1.
//
// The grid is persistent: 148 x (resident blocks per SM) blocks, and the weighted
// unit list (a unit costs n_scales x n_src warps) is cut into equal contiguous
// ranges, one per warp, so the tail is one row segment long instead of one tile.
// A block's range touches at most two (job, image) pairs; K^-1 and P = K.[R|t]
// for both are built once in the block prologue and kept in shared memory.
//
// Reductions (loss, 3x4 projection-matrix gradient per source) are kept in
// registers across a warp's run of rows, reduced warp -> block record in a fixed
// order; a small finalize kernel (one block per image) sums the records of each
// (job, image) pair, runs the pose chain and adds the loss: no atomics on the
// results, bitwise repeatable, and no fence / ticket traffic in the main kernel.
#include "photo_common.cuh"

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
// Per-pixel stages.  A "group" is one or two sources handled together.  For a pair of sources all
// floating-point work runs on Blackwell's packed fp32 pipe (FFMA2 / FADD2 / FMUL2: source 0 in the
// low half, source 1 in the high half of a 64-bit register pair), which halves the issue slots of
// the arithmetic; the coordinates of every source of the group are computed first, ONE warp vote
// decides between the unpredicated and the predicated tap loads, all 12 x NS loads are issued back
// to back, and only then is anything blended.
//
// Projection: cam = D * A + p3 with A = Q . (x, y, 1), Q = P[:, :3] . K^-1 (composed in fp64 in
// the block prologue, rounded once): the ray is never formed.  The perspective divide is one
// MUFU.RCP + one Newton step + a residual correction (as accurate as an IEEE divide: the sample
// position is the difference of two ~W-sized numbers, every ulp of px is 6e-5 px of bilinear
// weight); the normalise / un-normalise chain of the reference (transform.py:143-148 +
// grid_sample) is the identity and is dropped; the bilinear blend is written as nested lerps
// whose intermediates ARE the coordinate derivatives (d proj / d iy = bot - top).
//
// Depth gradient: d loss / d D = g_cam . A.  Because g_cam . (cx, cy, ze) = 0 identically (the
// perspective divide is scale invariant) and (cx, cy, ze) = D * A + p3', this equals
// -(g_cam . p3') / D - the analytically cancelled form, free of the ~W-sized cancellation the
// chain-rule form carries, and A need not stay live across the loads.
// ---------------------------------------------------------------------------------------------
template <int NS> struct Vec;
template <> struct Vec<1> { typedef float T; };
template <> struct Vec<2> { typedef float2 T; };

__device__ __forceinline__ float v_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 v_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float v_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float2 v_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float v_add(float a, float b) { return a + b; }
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float v_sub(float a, float b) { return a - b; }
__device__ __forceinline__ float2 v_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ void v_bc(float& v, float s) { v = s; }
__device__ __forceinline__ void v_bc(float2& v, float s) { v = make_float2(s, s); }
__device__ __forceinline__ float v_get(float v, int) { return v; }
__device__ __forceinline__ float v_get(float2 v, int k) { return k == 0 ? v.x : v.y; }
__device__ __forceinline__ void v_set(float& v, int, float s) { v = s; }
__device__ __forceinline__ void v_set(float2& v, int k, float s) { if (k == 0) v.x = s; else v.y = s; }
__device__ __forceinline__ float v_hsum(float v) { return v; }
__device__ __forceinline__ float v_hsum(float2 v) { return v.x + v.y; }

// row r of [Q | p3] for the group starting at source i0: x / y / constant coefficient and p3
__device__ __forceinline__ void load_q(const PairConst& pc, int i0, int r, float& qx, float& qy, float& qz, float& qw) {
    const float4 q = pc.Q[i0][r];
    qx = q.x; qy = q.y; qz = q.z; qw = q.w;
}
__device__ __forceinline__ void load_q(const PairConst& pc, int i0, int r, float2& qx, float2& qy, float2& qz, float2& qw) {
    const float4 a = pc.Q2[i0 >> 1][r][0], b = pc.Q2[i0 >> 1][r][1];
    qx = make_float2(a.x, a.y); qy = make_float2(a.z, a.w); qz = make_float2(b.x, b.y); qw = make_float2(b.z, b.w);
}

// p3 re-read from shared memory (asm volatile: a fresh load, not a value held in registers across the taps)
__device__ __forceinline__ void load_p3(const PairConst& pc, int i0, int r, float& w) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(&pc.Q[i0][r].w);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(a));
}
__device__ __forceinline__ void load_p3(const PairConst& pc, int i0, int r, float2& w) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(&pc.Q2[i0 >> 1][r][1].z);
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w.x), "=f"(w.y) : "r"(a));
}

// coordinates of one group of sources at one depth (stage A output)
template <int NS>
struct GState {
    typename Vec<NS>::T inv, px, py, fx, fy;
    int x0[NS], y0[NS];
    bool all_in;
};

// stage A: project the pixel into every source of the group
template <int NS>
__device__ __forceinline__ void group_project(const PairConst& pc, int i0, int H, int W, float xf, float yf, float D,
                                              GState<NS>& g) {
    typedef typename Vec<NS>::T V;
    V xv, yv, Dv, eps, neg1, two;
    v_bc(xv, xf); v_bc(yv, yf); v_bc(Dv, D); v_bc(eps, 1e-5f); v_bc(neg1, -1.0f); v_bc(two, 2.0f);
    V cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V qx, qy, qz;
        V p3;
        load_q(pc, i0, r, qx, qy, qz, p3);
        cam[r] = v_fma(Dv, v_fma(qx, xv, v_fma(qy, yv, qz)), p3);
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

// stage B: issue the 12 x NS tap loads (one warp vote picks unpredicated or predicated loads)
template <int NS>
__device__ __forceinline__ void group_load(const PairConst& pc, int i0, int plane, int H, int W, bool valid, bool pf,
                                           const GState<NS>& g, typename Vec<NS>::T (&v)[3][4], unsigned (&msk)[NS]) {
    typedef typename Vec<NS>::T V;
    const bool all_in = g.all_in;
    if (__all_sync(0xffffffffu, all_in || !valid)) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const float* __restrict__ cb = pc.src[i0 + k];
            const int o00 = (all_in && valid) ? g.y0[k] * W + g.x0[k] : 0;
            const int o01 = o00 + W, o10 = o00 + plane, o11 = o10 + W, o20 = o10 + plane, o21 = o20 + W;
            v_set(v[0][0], k, __ldg(cb + o00)); v_set(v[0][1], k, __ldg(cb + o00 + 1));
            v_set(v[0][2], k, __ldg(cb + o01)); v_set(v[0][3], k, __ldg(cb + o01 + 1));
            v_set(v[1][0], k, __ldg(cb + o10)); v_set(v[1][1], k, __ldg(cb + o10 + 1));
            v_set(v[1][2], k, __ldg(cb + o11)); v_set(v[1][3], k, __ldg(cb + o11 + 1));
            v_set(v[2][0], k, __ldg(cb + o20)); v_set(v[2][1], k, __ldg(cb + o20 + 1));
            v_set(v[2][2], k, __ldg(cb + o21)); v_set(v[2][3], k, __ldg(cb + o21 + 1));
            if (PH_PF_SRC > 0 && pf) {
                // the warp walks DOWN a strip: the source row the next unit(s) will newly touch, one line per channel
                const int opf = o01 + PH_PF_SRC * W;
                prefetch_l1(cb + opf); prefetch_l1(cb + (opf + plane)); prefetch_l1(cb + (opf + 2 * plane));
            }
            msk[k] = valid ? 15u : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const float* __restrict__ cb = pc.src[i0 + k];
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

// stage C: blend, L1, gradient terms
template <bool GRAD, bool IMG_GRAD, int NS>
__device__ __forceinline__ void group_blend(const PairConst& pc, int i0, int plane, int W, float yf, float D,
                                            const float (&t)[3], float w_e, bool valid, const GState<NS>& g,
                                            const typename Vec<NS>::T (&v)[3][4], const unsigned (&msk)[NS],
                                            typename Vec<NS>::T (&acc)[9], float& l1acc, float& gp, float (&gt)[3]) {
    typedef typename Vec<NS>::T V;
    V yv, Dv, eps, neg1;
    v_bc(yv, yf); v_bc(Dv, D); v_bc(eps, 1e-5f); v_bc(neg1, -1.0f);
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
            for (int r = 0; r < 3; ++r) load_p3(pc, i0, r, q3[r]);
            gp += v_hsum(v_fma(gcx, q3[0], v_fma(gcy, q3[1], v_mul(gcz, v_add(q3[2], eps)))));
        }
        // d loss / d P[r][:] = sum g_cam[r] * (D * ray, 1), ray = K^-1 (x, y, 1): accumulated in pixel
        // coordinates - sum h_r, sum h_r * y (and x * sum h_r at the flush, x being fixed per lane) - and
        // mapped through K^-1 by the finalize kernel
        const V hx = v_mul(gcx, Dv), hy = v_mul(gcy, Dv), hz = v_mul(gcz, Dv);
        acc[0] = v_add(acc[0], hx); acc[1] = v_add(acc[1], hy); acc[2] = v_add(acc[2], hz);
        acc[3] = v_fma(hx, yv, acc[3]); acc[4] = v_fma(hy, yv, acc[4]); acc[5] = v_fma(hz, yv, acc[5]);
        acc[6] = v_add(acc[6], gcx); acc[7] = v_add(acc[7], gcy); acc[8] = v_add(acc[8], gcz);
        if (IMG_GRAD) {
            const float m = valid ? w_e : 0.0f;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) gt[c] -= m * e[k][c];
                float* gbase = pc.g_src[i0 + k];
                if (gbase != nullptr) {
                    const float fxk = v_get(g.fx, k), fyk = v_get(g.fy, k);
                    const float wnw = (1.0f - fxk) * (1.0f - fyk), wne = fxk * (1.0f - fyk);
                    const float wsw = (1.0f - fxk) * fyk, wse = fxk * fyk;
                    const int og = g.y0[k] * W + g.x0[k];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float ec = m * e[k][c];
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


template <bool GRAD, bool IMG_GRAD, int NS>
__device__ __forceinline__ void group_pixel(const PairConst& pc, int i0, int plane, int H, int W, float xf, float yf,
                                            float D, const float (&t)[3], float w_e, bool valid, bool pf,
                                            typename Vec<NS>::T (&acc)[9], float& l1acc, float& gp, float (&gt)[3]) {
    GState<NS> g;
    typename Vec<NS>::T v[3][4];
    unsigned msk[NS];
    group_project<NS>(pc, i0, H, W, xf, yf, D, g);
    group_load<NS>(pc, i0, plane, H, W, valid, pf, g, v, msk);
    group_blend<GRAD, IMG_GRAD, NS>(pc, i0, plane, W, yf, D, t, w_e, valid, g, v, msk, acc, l1acc, gp, gt);
}

// Per-lane accumulators of one group -> this warp's shared record (ACCUMULATED in place).
// Record of source i: [0..2] sum h_r, [3..5] sum h_r * x, [6..8] sum h_r * y, [9..11] sum g_cam[r];
// value PLB_MAX_SRC*12: sum |diff|.
template <typename V, int NS>
__device__ __forceinline__ void flush_group(const V (&acc)[9], int i0, float l1, float xlane, float* rec, int lane) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int i = i0 + k;
        float v[16];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            v[q] = v_get(acc[q], k); v[3 + q] = v_get(acc[q], k) * xlane;
            v[6 + q] = v_get(acc[3 + q], k); v[9 + q] = v_get(acc[6 + q], k);
        }
        v[12] = (i == 0) ? l1 : 0.0f;
        v[13] = v[14] = v[15] = 0.0f;
        int which;
        const float r = warp_reduce16(v, lane, which);
        if ((lane & 1) == 0) {
            if (which < 12) rec[i * 12 + which] += r;
            else if (which == 12 && i == 0) rec[PLB_MAX_SRC * 12] += r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Single-source directions (3-4-source kernel variants only: the <= 2-source variants have no registers to spare,
// measured 196 vs 183 us on c2): TWO ROWS of the strip ride the packed pipe (low half = row y, high half = row
// y + 1 of the same source) instead of two sources of one row - the scalar path costs 250 warp instructions
// per row segment against 165 for half of a packed pair.  Same stages as above; the projection constants are
// broadcast, the target pixel / depth / row coordinate differ per half, and the depth-gradient numerator is
// kept per half (each row writes its own gradient).  `vk[k]`: half k is a real pixel (x < W and, for the odd
// last row of a run, k == 0).
// ---------------------------------------------------------------------------------------------
template <bool GRAD>
__device__ __forceinline__ void rowpair_pixel(const PairConst& pc, int i0, int plane, int H, int W, float xf, float2 yv,
                                              float2 Dv, const float (&t)[2][3], float w_e, const bool (&vk)[2], bool pf,
                                              float2 (&acc)[9], float& l1acc, float2& gpv) {
    typedef float2 V;
    V xv, eps, neg1, two;
    v_bc(xv, xf); v_bc(eps, 1e-5f); v_bc(neg1, -1.0f); v_bc(two, 2.0f);
    // ---- stage A: project both rows ----------------------------------------------------------------
    V cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float4 q = pc.Q[i0][r];
        V qx, qy, qz, p3;
        v_bc(qx, q.x); v_bc(qy, q.y); v_bc(qz, q.z); v_bc(p3, q.w);
        cam[r] = v_fma(Dv, v_fma(qx, xv, v_fma(qy, yv, qz)), p3);
    }
    const V ze = v_add(cam[2], eps);
    const V nze = v_mul(ze, neg1);
    V inv;
    {
        float r0, r1;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(ze.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(ze.y));
        inv = make_float2(r0, r1);
    }
    inv = v_mul(inv, v_fma(nze, inv, two));                  // Newton step
    V px = v_mul(cam[0], inv), py = v_mul(cam[1], inv);
    px = v_fma(v_fma(px, nze, cam[0]), inv, px);             // residual correction: IEEE-accurate quotient
    py = v_fma(v_fma(py, nze, cam[1]), inv, py);
    int x0[2], y0[2];
    V fx, fy;
    bool all_in = true;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float ixc = fminf(fmaxf(v_get(px, k), -2.0f), (float)(W + 1));
        const float iyc = fminf(fmaxf(v_get(py, k), -2.0f), (float)(H + 1));
        const float xfl = floorf(ixc), yfl = floorf(iyc);
        x0[k] = (int)xfl; y0[k] = (int)yfl;
        v_set(fx, k, ixc - xfl); v_set(fy, k, iyc - yfl);
        all_in = all_in && (!vk[k] || (((unsigned)x0[k] < (unsigned)(W - 1)) && ((unsigned)y0[k] < (unsigned)(H - 1))));
    }
    // ---- stage B: 24 tap loads ----------------------------------------------------------------------
    V v[3][4];
    unsigned msk[2];
    const float* __restrict__ cb = pc.src[i0];
    if (__all_sync(0xffffffffu, all_in)) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int o00 = vk[k] ? y0[k] * W + x0[k] : 0;
            const int o01 = o00 + W, o10 = o00 + plane, o11 = o10 + W, o20 = o10 + plane, o21 = o20 + W;
            v_set(v[0][0], k, __ldg(cb + o00)); v_set(v[0][1], k, __ldg(cb + o00 + 1));
            v_set(v[0][2], k, __ldg(cb + o01)); v_set(v[0][3], k, __ldg(cb + o01 + 1));
            v_set(v[1][0], k, __ldg(cb + o10)); v_set(v[1][1], k, __ldg(cb + o10 + 1));
            v_set(v[1][2], k, __ldg(cb + o11)); v_set(v[1][3], k, __ldg(cb + o11 + 1));
            v_set(v[2][0], k, __ldg(cb + o20)); v_set(v[2][1], k, __ldg(cb + o20 + 1));
            v_set(v[2][2], k, __ldg(cb + o21)); v_set(v[2][3], k, __ldg(cb + o21 + 1));
            if (PH_PF_SRC > 0 && pf && k == 1) {
                // the next pair of rows touches two new source rows below the lower footprint
                const int opf = o01 + PH_PF_SRC * W;
                prefetch_l1(cb + opf); prefetch_l1(cb + (opf + plane)); prefetch_l1(cb + (opf + 2 * plane));
                prefetch_l1(cb + (opf + W)); prefetch_l1(cb + (opf + W + plane)); prefetch_l1(cb + (opf + W + 2 * plane));
            }
            msk[k] = vk[k] ? 15u : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const bool vx0 = (unsigned)x0[k] < (unsigned)W, vx1 = (unsigned)(x0[k] + 1) < (unsigned)W;
            const bool vy0 = (unsigned)y0[k] < (unsigned)H, vy1 = (unsigned)(y0[k] + 1) < (unsigned)H;
            const bool mnw = vk[k] && vx0 && vy0, mne = vk[k] && vx1 && vy0;
            const bool msw = vk[k] && vx0 && vy1, mse = vk[k] && vx1 && vy1;
            const int o00 = y0[k] * W + x0[k];
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
    // ---- stage C: blend, L1, gradient terms ----------------------------------------------------------
    V Gx, Gy;
    v_bc(Gx, 0.0f); v_bc(Gy, 0.0f);
    float l1[2] = {0.0f, 0.0f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const V dA = v_sub(v[c][1], v[c][0]), dB = v_sub(v[c][3], v[c][2]);
        const V top = v_fma(fx, dA, v[c][0]), bot = v_fma(fx, dB, v[c][2]);
        const V dV = v_sub(bot, top);
        const V proj = v_fma(fy, dV, top);
        V sg;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float d = v_get(proj, k) - t[k][c];
            l1[k] += fabsf(d);
            if (GRAD) {
                const float ne = (d != 0.0f) ? 1.0f : 0.0f;
                v_set(sg, k, __int_as_float(__float_as_int(ne) | (__float_as_int(d) & 0x80000000)));
            }
        }
        if (GRAD) {
            Gx = v_fma(sg, v_fma(fy, v_sub(dB, dA), dA), Gx);
            Gy = v_fma(sg, dV, Gy);
        }
    }
    l1acc += (vk[0] ? l1[0] : 0.0f) + (vk[1] ? l1[1] : 0.0f);
    if (GRAD) {
        V gi;
        V s = v_fma(Gx, px, v_mul(Gy, py));
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const bool use = msk[k] != 0u;
            v_set(gi, k, use ? w_e * v_get(inv, k) : 0.0f);
            v_set(s, k, use ? v_get(s, k) : 0.0f);
        }
        const V gcx = v_mul(Gx, gi), gcy = v_mul(Gy, gi);
        const V gcz = v_mul(v_mul(s, gi), neg1);
        {
            V q3[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { float w3; load_p3(pc, i0, r, w3); v_bc(q3[r], w3); }
            gpv = v_add(gpv, v_fma(gcx, q3[0], v_fma(gcy, q3[1], v_mul(gcz, v_add(q3[2], eps)))));
        }
        const V hx = v_mul(gcx, Dv), hy = v_mul(gcy, Dv), hz = v_mul(gcz, Dv);
        acc[0] = v_add(acc[0], hx); acc[1] = v_add(acc[1], hy); acc[2] = v_add(acc[2], hz);
        acc[3] = v_fma(hx, yv, acc[3]); acc[4] = v_fma(hy, yv, acc[4]); acc[5] = v_fma(hz, yv, acc[5]);
        acc[6] = v_add(acc[6], gcx); acc[7] = v_add(acc[7], gcy); acc[8] = v_add(acc[8], gcz);
    }
}

// One run of a SINGLE-source pair (no image gradients): rows two at a time on the packed pipe.
template <bool GRAD, bool MULTI, bool HEAD>
__device__ __forceinline__ void run_rows_single(const plb_photo_args& a, const PairConst& pc, int strip, int y, int rows,
                                                int lane, float* rec) {
    const int H = a.H, W = a.W, plane = H * W;
    const float w_e = pc.w_e;
    const int x = strip * 32 + lane;
    const bool valid = x < W;
    const float xf = (float)x;
    const float* __restrict__ tgt_b = pc.tgt;
    const int xc = min(x, W - 1);
    float2 acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = make_float2(0.0f, 0.0f);
    float l1acc = 0.0f;
    const int y_end = y + rows;
    // software pipeline over row pairs: the target pixels of the NEXT pair are loaded while this one is processed
    auto row_of = [&](int yy) -> int { return min(yy, H - 1) * W + xc; };   // the odd last row of a run repeats itself (masked)
    float tn[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int o = row_of(y + k);
        tn[k][0] = __ldg(tgt_b + o); tn[k][1] = __ldg(tgt_b + (o + plane)); tn[k][2] = __ldg(tgt_b + (o + 2 * plane));
    }
#pragma unroll 1
    for (; y < y_end; y += 2) {
        float t[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k) { t[k][0] = tn[k][0]; t[k][1] = tn[k][1]; t[k][2] = tn[k][2]; }
        const bool pf = y + 2 < y_end;
        if (pf) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int o = row_of(y + 2 + k);
                tn[k][0] = __ldg(tgt_b + o); tn[k][1] = __ldg(tgt_b + (o + plane)); tn[k][2] = __ldg(tgt_b + (o + 2 * plane));
            }
        }
        const bool vk[2] = {valid, valid && (y + 1 < y_end)};
        const int o0 = row_of(y), o1 = row_of(y + 1);
        const float2 yv = make_float2((float)y, (float)(y + 1));
        const int n_scales = MULTI ? pc.n_scales : 1, lowres = MULTI ? pc.lowres : 0;
#pragma unroll 1
        for (int s = 0; s < n_scales; ++s) {
            const float* disp_b = pc.disp[s];
            const bool full = !((lowres >> s) & 1);
            float2 Dv;
            if (full) {
                float d0 = __ldg(disp_b + o0), d1 = __ldg(disp_b + o1);
                if (HEAD) { d0 = head_disp(d0, a.head_alpha, a.head_beta); d1 = head_disp(d1, a.head_alpha, a.head_beta); }
                Dv = a.input_is_depth == PLB_INPUT_DEPTH ? make_float2(d0, d1)
                                                         : make_float2(rcp_nr(fmaf(a.disp_a, d0, a.disp_b)), rcp_nr(fmaf(a.disp_a, d1, a.disp_b)));
            } else {
                const int dh = pc.dh[s], dw = pc.dw[s];
                int x0, x1; float lx0, lx1;
                up_coord(xc, pc.sx[s], dw, x0, x1, lx0, lx1);
                float Dk[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    int y0, y1; float ly0, ly1;
                    up_coord(min(y + k, H - 1), pc.sy[s], dh, y0, y1, ly0, ly1);
                    float v00 = __ldg(disp_b + (y0 * dw + x0)), v01 = __ldg(disp_b + (y0 * dw + x1));
                    float v10 = __ldg(disp_b + (y1 * dw + x0)), v11 = __ldg(disp_b + (y1 * dw + x1));
                    if (HEAD) {
                        v00 = head_disp(v00, a.head_alpha, a.head_beta); v01 = head_disp(v01, a.head_alpha, a.head_beta);
                        v10 = head_disp(v10, a.head_alpha, a.head_beta); v11 = head_disp(v11, a.head_alpha, a.head_beta);
                    }
                    if (a.input_is_depth != PLB_INPUT_DEPTH) {
                        v00 = rcp_nr(fmaf(a.disp_a, v00, a.disp_b)); v01 = rcp_nr(fmaf(a.disp_a, v01, a.disp_b));
                        v10 = rcp_nr(fmaf(a.disp_a, v10, a.disp_b)); v11 = rcp_nr(fmaf(a.disp_a, v11, a.disp_b));
                    }
                    Dk[k] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
                }
                Dv = make_float2(Dk[0], Dk[1]);
            }
            float2 gpv = make_float2(0.0f, 0.0f);
            rowpair_pixel<GRAD>(pc, 0, plane, H, W, xf, yv, Dv, t, w_e, vk, pf && s == 0, acc, l1acc, gpv);
            if (GRAD) {
                float* g = pc.g_disp[s];
                if (g != nullptr) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (vk[k]) {
                            const float D = v_get(Dv, k), gp = v_get(gpv, k);
                            float gv = (full && a.input_is_depth != PLB_INPUT_DEPTH) ? a.disp_a * D * gp : -gp * rcp_nr(D);
                            if (full && HEAD) gv *= head_chain_from_depth(D, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
                            g[k == 0 ? o0 : o1] = gv;
                        }
                    }
                }
            }
        }
    }
    // both halves belong to the one source: fold them, then the usual warp reduction into the record
    float accs[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) accs[k] = acc[k].x + acc[k].y;
    flush_group<float, 1>(accs, 0, l1acc, xf, rec, lane);
    __syncwarp();
}

// One run: consecutive rows [y, y + rows) of one 32-px strip of one (job, image) pair, NSRC sources.
template <bool GRAD, bool IMG_GRAD, int NSRC, bool MULTI, bool HEAD>
__device__ __forceinline__ void run_rows(const plb_photo_args& a, const PairConst& pc, int strip, int y, int rows,
                                         int lane, float* rec) {
    constexpr int NP = NSRC / 2, ODD = NSRC & 1;
    const int H = a.H, W = a.W, plane = H * W;
    const float w_e = pc.w_e;
    const int x = strip * 32 + lane;
    const bool valid = x < W;
    const float xf = (float)x;
    const float* __restrict__ tgt_b = pc.tgt;
    int o = y * W + min(x, W - 1);

    float2 accp[NP > 0 ? NP : 1][9];
    float accs[9];
    float l1acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int g = 0; g < (NP > 0 ? NP : 1); ++g) accp[g][k] = make_float2(0.0f, 0.0f);
        accs[k] = 0.0f;
    }

    // software pipeline: the target pixel (and the full-resolution disparity) of the NEXT row are loaded while
    // this row is processed, so a warp pays one memory round trip per row (the taps), not two
    float tn[3], dn = 0.0f;
    tn[0] = __ldg(tgt_b + o); tn[1] = __ldg(tgt_b + (o + plane)); tn[2] = __ldg(tgt_b + (o + 2 * plane));
    if (!MULTI) dn = __ldg(pc.disp[0] + o);
#pragma unroll 1
    for (int r = 0; r < rows; ++r, ++y, o += W) {
        float t[3], gt[3] = {0.0f, 0.0f, 0.0f};
        t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
        const float d_cur = dn;
        const bool pf = r + 1 < rows;
        if (pf) {
            const int on = o + W;
            tn[0] = __ldg(tgt_b + on); tn[1] = __ldg(tgt_b + (on + plane)); tn[2] = __ldg(tgt_b + (on + 2 * plane));
            if (!MULTI) dn = __ldg(pc.disp[0] + on);
        }
        const float yf = (float)y;
        if (!MULTI) {
            const float d0 = HEAD ? head_disp(d_cur, a.head_alpha, a.head_beta) : d_cur;
            const float D = a.input_is_depth == PLB_INPUT_DEPTH ? d0 : rcp_nr(fmaf(a.disp_a, d0, a.disp_b));
            float gp = 0.0f;
#pragma unroll
            for (int g = 0; g < NP; ++g)
                group_pixel<GRAD, IMG_GRAD, 2>(pc, 2 * g, plane, H, W, xf, yf, D, t, w_e, valid, pf, accp[g], l1acc, gp, gt);
            if (ODD) group_pixel<GRAD, IMG_GRAD, 1>(pc, NSRC - 1, plane, H, W, xf, yf, D, t, w_e, valid, pf, accs, l1acc, gp, gt);
            if (GRAD) {
                float* g = pc.g_disp[0];
                // d loss / d D = -gp / D;  d D / d disp = -disp_a * D^2
                if (valid && g != nullptr) {
                    float gv = a.input_is_depth == PLB_INPUT_DEPTH ? -gp * rcp_nr(D) : a.disp_a * D * gp;
                    if (HEAD) gv *= head_chain_from_depth(D, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
                    g[o] = gv;
                }
            }
        } else {
            const int n_scales = pc.n_scales, lowres = pc.lowres;
#pragma unroll 1
            for (int s = 0; s < n_scales; ++s) {
                const float* disp_b = pc.disp[s];
                const bool full = !((lowres >> s) & 1);
                float D, gp = 0.0f;
                if (full) {
                    float d = __ldg(disp_b + o);
                    if (HEAD) d = head_disp(d, a.head_alpha, a.head_beta);
                    D = a.input_is_depth == PLB_INPUT_DEPTH ? d : rcp_nr(fmaf(a.disp_a, d, a.disp_b));
                } else {
                    const int dh = pc.dh[s], dw = pc.dw[s];
                    int x0, x1, y0, y1; float lx0, lx1, ly0, ly1;
                    up_coord(min(x, W - 1), pc.sx[s], dw, x0, x1, lx0, lx1);
                    up_coord(y, pc.sy[s], dh, y0, y1, ly0, ly1);
                    float v00 = __ldg(disp_b + (y0 * dw + x0)), v01 = __ldg(disp_b + (y0 * dw + x1));
                    float v10 = __ldg(disp_b + (y1 * dw + x0)), v11 = __ldg(disp_b + (y1 * dw + x1));
                    if (HEAD) {
                        v00 = head_disp(v00, a.head_alpha, a.head_beta); v01 = head_disp(v01, a.head_alpha, a.head_beta);
                        v10 = head_disp(v10, a.head_alpha, a.head_beta); v11 = head_disp(v11, a.head_alpha, a.head_beta);
                    }
                    if (a.input_is_depth != PLB_INPUT_DEPTH) {
                        v00 = rcp_nr(fmaf(a.disp_a, v00, a.disp_b)); v01 = rcp_nr(fmaf(a.disp_a, v01, a.disp_b));
                        v10 = rcp_nr(fmaf(a.disp_a, v10, a.disp_b)); v11 = rcp_nr(fmaf(a.disp_a, v11, a.disp_b));
                    }
                    D = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
                }
#pragma unroll
                for (int g = 0; g < NP; ++g)
                    group_pixel<GRAD, IMG_GRAD, 2>(pc, 2 * g, plane, H, W, xf, yf, D, t, w_e, valid, pf && s == 0, accp[g], l1acc, gp, gt);
                if (ODD) group_pixel<GRAD, IMG_GRAD, 1>(pc, NSRC - 1, plane, H, W, xf, yf, D, t, w_e, valid, pf && s == 0, accs, l1acc, gp, gt);
                if (GRAD) {
                    float* g = pc.g_disp[s];
                    if (valid && g != nullptr)
                    {
                        float gv = (full && a.input_is_depth != PLB_INPUT_DEPTH) ? a.disp_a * D * gp : -gp * rcp_nr(D);
                        if (full && HEAD) gv *= head_chain_from_depth(D, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
                        g[o] = gv;
                    }
                }
            }
        }
        if (GRAD && IMG_GRAD && valid && pc.g_tgt != nullptr) {
            float* g = pc.g_tgt + o;
            atomicAdd(g, gt[0]); atomicAdd(g + plane, gt[1]); atomicAdd(g + 2 * plane, gt[2]);
        }
    }
    // end of the run: the lane's column (and possibly the pair) changes
#pragma unroll
    for (int g = 0; g < NP; ++g) flush_group<float2, 2>(accp[g], 2 * g, l1acc, xf, rec, lane);
    if (ODD) flush_group<float, 1>(accs, NSRC - 1, l1acc, xf, rec, lane);
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

template <bool GRAD, bool IMG_GRAD, int MAXSRC, bool MULTI, bool HEAD>
__global__ void __launch_bounds__(photo_threads(MAXSRC, MULTI), (MAXSRC <= 2) ? PH_MIN_BLOCKS : 2)
photo_l1_kernel(const __grid_constant__ PhotoLaunch p) {
    constexpr int PH_THREADS = photo_threads(MAXSRC, MULTI), PH_WARPS = PH_THREADS / 32;
    const plb_photo_args& a = p.a;
    // the finalize grid (programmatic dependent launch) may be scheduled from now on; it waits for
    // this grid to complete before it reads the records
    asm volatile("griddepcontrol.launch_dependents;");
    if (skip_launch(a.skip_if_unit)) return;
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 0 : 4);
#if defined(PLB_DEBUG_EXIT_AT) && PLB_DEBUG_EXIT_AT == 1
    return;
#endif

    char* ws = (char*)a.workspace;
    float* records = (float*)(ws + p.L.records);
    float* gup = (float*)(ws + p.L.gup);

    const int H = a.H, W = a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane = H * W;

    __shared__ PairConst s_pc[2];
    __shared__ float s_rec[2][PH_WARPS][PH_NREC + 3];
    __shared__ int s_pair[2];

    // ---- this warp's unit range: equal shares of the weighted unit list (32-bit maths) ---------
    auto pos_of = [&](int w) -> int { return w * p.share + min(w, p.share_rem); };
    auto unit_of = [&](int pos) -> int {  // first unit whose weight interval starts at or after pos
        int j = 0;
        while (j + 1 < a.n_jobs && pos >= (int)p.weight_start[j + 1]) ++j;
        const int rel = pos - (int)p.weight_start[j];
        return p.unit_start[j] + (rel + p.unit_weight[j] - 1) / p.unit_weight[j];
    };
#ifdef PLB_DEBUG_REVERSE_BLOCKS
    const int vblk = gridDim.x - 1 - blockIdx.x;
#else
    const int vblk = blockIdx.x;     // virtual block index: which share of the unit list this block owns
#endif
    const int gw = vblk * PH_WARPS + warp;
    const int blk_u0 = unit_of(pos_of(vblk * PH_WARPS));
    const int blk_u1 = unit_of(pos_of(vblk * PH_WARPS + PH_WARPS));
    const int u0 = unit_of(pos_of(gw));
    const int u1 = unit_of(pos_of(gw + 1));
    const bool empty_block = blk_u1 <= blk_u0;  // more blocks than work: still publishes (empty) records
    const int pairA = empty_block ? 0 : blk_u0 / p.units_per_pair;
    const int pairB = empty_block ? 0 : (blk_u1 - 1) / p.units_per_pair;

#if defined(PLB_DEBUG_EXIT_AT) && PLB_DEBUG_EXIT_AT == 2
    if (u0 >= 0) return;
#endif
    // ---- block prologue: context of the (at most two) pairs this block touches ------------------
    if (tid < 2) s_pair[tid] = empty_block ? -1 : ((tid == 0) ? pairA : (pairB != pairA ? pairB : -1));
    for (int k = tid; k < 2 * PH_WARPS * (PH_NREC + 3); k += PH_THREADS) (&s_rec[0][0][0])[k] = 0.0f;
    {
        const int set = warp >> 1;  // warps 0,1 -> set 0; warps 2,3 -> set 1
        const int pair = set == 0 ? pairA : pairB;
        if (!empty_block && warp < 4 && (set == 0 || pairB != pairA)) {
            const int jb = pair / a.B, b = pair - jb * a.B;
            const plb_photo_job& job = a.jobs[jb];
            PairConst& pc = s_pc[set];
            const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
            const size_t img = (size_t)b * 3 * plane;
            if ((warp & 1) == 0) {
                if (lane == 1) {
                    pc.tgt = job.tgt + img;
                    pc.g_tgt = (GRAD && IMG_GRAD && job.g_tgt) ? job.g_tgt + img : nullptr;
                    pc.n_src = job.n_src; pc.n_scales = job.n_scales; pc.lowres = p.lowres[jb];
                    pc.w_e = p.w_e[jb] * (a.upstream ? __ldg(a.upstream) : 1.0f);
                }
                if (lane >= 4 && lane < 4 + job.n_scales) {
                    const int sc = lane - 4;
                    const int dh = job.dh[sc], dw = job.dw[sc];
                    pc.dh[sc] = dh; pc.dw[sc] = dw;
                    pc.sx[sc] = (float)dw / (float)W; pc.sy[sc] = (float)dh / (float)H;
                    pc.disp[sc] = job.disp[sc] + (size_t)b * dh * dw;
                    float* g = nullptr;
                    if (GRAD && job.g_disp[sc] != nullptr)
                        g = ((p.lowres[jb] >> sc) & 1)
                                ? gup + ((size_t)(jb * PLB_MAX_SCALES + sc) * a.B + b) * plane
                                : job.g_disp[sc] + (size_t)b * plane;
                    pc.g_disp[sc] = g;
                }
                if (lane == 8) {
                    // K^-1 (fp64 adjugate, rounded to fp32 as transform.py:92 does) - in parallel with the pose chain
                    // of the odd warp: executed once per block, this code runs cold, and two short dependent
                    // chains on two warps finish sooner than one long chain
                    float ki[9];
                    kinv_f32(Kb, a.k_is_f64, ki);
#pragma unroll
                    for (int k = 0; k < 9; ++k) pc.kinv[k] = ki[k];
                }
            } else {
                if (lane < job.n_src) {
                    float M[12], P[12];
                    pose_to_M(a.poses + ((size_t)b * a.n_pose + job.pose_index[lane]) * 6, a.rotation_mode,
                              job.pose_inv[lane], M);
                    k_times_M(Kb, a.k_is_f64, M, P);
#pragma unroll
                    for (int r = 0; r < 3; ++r) pc.P[lane][r] = make_float4(P[r * 4], P[r * 4 + 1], P[r * 4 + 2], P[r * 4 + 3]);
                    pc.src[lane] = job.src[lane] + img;
                    pc.g_src[lane] = (GRAD && IMG_GRAD && job.g_src[lane]) ? job.g_src[lane] + img : nullptr;
                }
            }
        }
    }
    __syncthreads();
    {
        const int set = warp >> 1;
        if (!empty_block && warp < 4 && (warp & 1) == 1 && (set == 0 || pairB != pairA)) {
            PairConst& pc = s_pc[set];
            const int n_src = pc.n_src;
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
            // packed copy for source pairs (2g, 2g+1): low half = even source, high half = odd source
            if (lane < (PLB_MAX_SRC / 2) * 3) {
                const int g = lane / 3, r = lane - g * 3;
                if (2 * g + 1 < n_src) {
                    const float4 q0 = pc.Q[2 * g][r], q1 = pc.Q[2 * g + 1][r];
                    pc.Q2[g][r][0] = make_float4(q0.x, q1.x, q0.y, q1.y);
                    pc.Q2[g][r][1] = make_float4(q0.z, q1.z, q0.w, q1.w);
                }
            }
        }
    }
    __syncthreads();
    DBG_STAMP(threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1), blockIdx.x == 0 ? 1 : 5);
#if defined(PLB_DEBUG_EXIT_AT) && PLB_DEBUG_EXIT_AT == 3
    if (u0 >= 0) return;
#endif

    // ---- runs: consecutive rows of one 32-px strip of one (job, image) pair ---------------------
    int u = u0;
#ifdef PLB_DEBUG_SKIP_UNITS
    u = u1;   // measurement only: fixed cost of prologue + epilogue
#endif
#pragma unroll 1
    while (u < u1) {
        const int pair = u / p.units_per_pair;
        const int local = u - pair * p.units_per_pair;
        const int strip = local / H;
        const int y = local - strip * H;
        const int rows = min(u1, u - y + H) - u;
        const int set = (pair == pairA) ? 0 : 1;
        const PairConst& pc = s_pc[set];
        float* rec = s_rec[set][warp];
        const int n_src = pc.n_src;
        if (MAXSRC <= 2) {
            if (n_src == 2) run_rows<GRAD, IMG_GRAD, 2, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else if (PH_ROWPAIR && MAXSRC > 2 && !IMG_GRAD) run_rows_single<GRAD, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else run_rows<GRAD, IMG_GRAD, 1, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
        } else {
            if (n_src == 4) run_rows<GRAD, IMG_GRAD, 4, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else if (n_src == 3) run_rows<GRAD, IMG_GRAD, 3, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else if (n_src == 2) run_rows<GRAD, IMG_GRAD, 2, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else if (PH_ROWPAIR && MAXSRC > 2 && !IMG_GRAD) run_rows_single<GRAD, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
            else run_rows<GRAD, IMG_GRAD, 1, MULTI, HEAD>(a, pc, strip, y, rows, lane, rec);
        }
        u += rows;
    }

    // ---- block records: fixed-order sum over the 8 warps, one record per touched pair; the
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
            const long long w0 = p.weight_start[jb] + (long long)(pr * p.units_per_pair - p.unit_start[jb]) * p.unit_weight[jb];
            const long long w1 = w0 + (long long)p.units_per_pair * p.unit_weight[jb];
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
    if ((long long)a->H * ((a->W + 31) / 32) * a->B * a->n_jobs * PLB_MAX_SCALES * PLB_MAX_SRC * 8 >= (1LL << 31)) return PLB_EINVAL;
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

template <bool GRAD, bool IMG, int MS, bool MULTI, bool HEAD>
static int blocks_per_sm() {
    static int cached = 0;
    if (cached == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photo_l1_kernel<GRAD, IMG, MS, MULTI, HEAD>, photo_threads(MS, MULTI), 0) != cudaSuccess ||
            n < 1) {
            (void)cudaGetLastError();
            n = 2;
        }
        cached = n;
    }
    return cached;
}

static int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = 148;
        }
        cached = n;
    }
    return cached;
}

int photo_upsample_T_launch(const PhotoLaunch& p, cudaStream_t st) {
    const plb_photo_args* a = &p.a;
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(photo_upsample_T_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UT_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
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
                    if (!job.g_disp[s] || (job.dh[s] == a->H && job.dw[s] == a->W)) continue;
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

int photo_l1_launch(const plb_photo_args* a, cudaStream_t st) {
    if (a != nullptr && a->n_jobs >= 1 && a->n_jobs <= PLB_MAX_JOBS && a->jobs[0].mode == PLB_PHOTO_MIN_REPROJ)
        return photo_min_launch(a, st);
    int rc = validate_photo(a);
    if (rc != PLB_OK) return rc;
    PhotoLaunch p;
    p.a = *a;
    p.L = photo_layout(*a);
    bool img_grad = false;
    int maxsrc = 1;
    for (int j = 0; j < a->n_jobs; ++j) {
        if (a->jobs[j].g_tgt) img_grad = true;
        for (int i = 0; i < a->jobs[j].n_src; ++i)
            if (a->jobs[j].g_src[i]) img_grad = true;
        if (a->jobs[j].n_src > maxsrc) maxsrc = a->jobs[j].n_src;
    }
    const bool lowres_grad = photo_has_lowres_grad(*a);
    // the disparity head folded in (PLB_INPUT_LOGIT) is its own set of kernel variants: the default variants keep
    // their register budget (the single-scale one has none to spare)
    const bool head = a->input_is_depth == PLB_INPUT_LOGIT;
    if (head && img_grad) return PLB_EINVAL;       // image gradients: disparity / depth inputs only
    p.strips = (a->W + 31) / 32;
    p.units_per_pair = p.strips * a->H;
    p.n_pairs = a->n_jobs * a->B;
    long long wsum = 0;
    int usum = 0, min_w = 1 << 30;
    for (int j = 0; j < PLB_MAX_JOBS; ++j) {
        p.weight_start[j] = wsum;
        p.unit_start[j] = usum;
        p.unit_weight[j] = 1;
        p.w_e[j] = 0.0f;
        p.lowres[j] = 0;
        if (j < a->n_jobs) {
            // cost model of a unit: a PAIR of sources runs on the packed fp32 pipe and costs less than two single
            // (scalar) sources; the ratio is tuned per kernel variant (profiles/README.md)
            {
                const int wp = maxsrc <= 2 ? PH_W_PAIR : PH_W_PAIR4, wo = maxsrc <= 2 ? PH_W_ODD : PH_W_ODD4;
                const int ws = PH_W_SINGLE4;      // a single-source job beside a 3-4-source one: row pairs on the packed pipe
                const int n = a->jobs[j].n_src;
                p.unit_weight[j] = a->jobs[j].n_scales * ((n == 1 && PH_ROWPAIR && maxsrc > 2 && !img_grad) ? ws : wp * (n / 2) + wo * (n & 1));
            }
            if (p.unit_weight[j] < min_w) min_w = p.unit_weight[j];
            wsum += (long long)p.unit_weight[j] * p.units_per_pair * a->B;
            usum += p.units_per_pair * a->B;
            p.w_e[j] = a->jobs[j].term_weight / (3.0f * (float)a->B * (float)a->H * (float)a->W);
            for (int s = 0; s < a->jobs[j].n_scales; ++s)
                if (a->jobs[j].dh[s] != a->H || a->jobs[j].dw[s] != a->W) p.lowres[j] |= 1 << s;
        }
    }
    p.weight_start[PLB_MAX_JOBS] = wsum;
    p.unit_start[PLB_MAX_JOBS] = usum;

    bool multi = false;
    for (int j = 0; j < a->n_jobs; ++j)
        if (a->jobs[j].n_scales != 1 || p.lowres[j] != 0) multi = true;
    int bps;
#define PLB_BPS(G, I, M) (head ? (multi ? blocks_per_sm<G, false, M, true, true>() : blocks_per_sm<G, false, M, false, true>()) \
                               : (multi ? blocks_per_sm<G, I, M, true, false>() : blocks_per_sm<G, I, M, false, false>()))
    if (!a->want_grad) bps = maxsrc <= 2 ? PLB_BPS(false, false, 2) : PLB_BPS(false, false, 4);
    else if (img_grad) bps = maxsrc <= 2 ? PLB_BPS(true, true, 2) : PLB_BPS(true, true, 4);
    else bps = maxsrc <= 2 ? PLB_BPS(true, false, 2) : PLB_BPS(true, false, 4);
#undef PLB_BPS
    long long grid = (long long)sm_count() * bps;
    // a block's weight range must not exceed the lightest pair, so that it touches at most two pairs
    const long long pair_w_min = (long long)min_w * p.units_per_pair;
    const long long need = (wsum + pair_w_min - 1) / pair_w_min + 1;
    if (grid > usum) grid = usum;                  // tiny problems: no more blocks than units ...
    if (grid < need) grid = need;                  // ... but never so few that a block spans three pairs
    if (grid > photo_max_grid(*a)) grid = photo_max_grid(*a);
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.warps_per_block = photo_threads(maxsrc <= 2 ? 2 : 4, multi) / 32;
    p.n_warps = p.grid * p.warps_per_block;
    p.share = (int)(wsum / p.n_warps);
    p.share_rem = (int)(wsum % p.n_warps);

    dim3 g(p.grid), block(photo_threads(maxsrc <= 2 ? 2 : 4, multi));
#define PLB_LAUNCH(G, I, M)                                           \
    do {                                                              \
        if (head) {                                                   \
            if (multi) photo_l1_kernel<G, false, M, true, true><<<g, block, 0, st>>>(p); \
            else photo_l1_kernel<G, false, M, false, true><<<g, block, 0, st>>>(p);      \
        } else if (multi) photo_l1_kernel<G, I, M, true, false><<<g, block, 0, st>>>(p); \
        else photo_l1_kernel<G, I, M, false, false><<<g, block, 0, st>>>(p);             \
    } while (0)
    if (!a->want_grad) { if (maxsrc <= 2) PLB_LAUNCH(false, false, 2); else PLB_LAUNCH(false, false, 4); }
    else if (img_grad) { if (maxsrc <= 2) PLB_LAUNCH(true, true, 2); else PLB_LAUNCH(true, true, 4); }
    else { if (maxsrc <= 2) PLB_LAUNCH(true, false, 2); else PLB_LAUNCH(true, false, 4); }
#undef PLB_LAUNCH
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
    if (lowres_grad) {
        const int rc2 = photo_upsample_T_launch(p, st);
        if (rc2 != PLB_OK) return rc2;
    }
    return PLB_OK;
}

size_t photo_workspace_bytes(const plb_photo_args* a) { return photo_layout(*a).total; }

}  // namespace plb
