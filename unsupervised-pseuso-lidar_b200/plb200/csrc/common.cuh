// Shared device code: intrinsics inverse, pose matrices (+ vjp), bilinear
// sampling helpers, the align_corners=False upsampling rule, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "plb200.h"

namespace plb {

extern unsigned long long g_launches;  // host-side counter (api.cu)

// per-device host-side caches (function attributes, occupancy, SM count): a process may drive several GPUs
constexpr int PLB_MAX_DEVICES = 64;
static inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); d = 0; }
    return (d >= 0 && d < PLB_MAX_DEVICES) ? d : 0;
}

#define PLB_CHECK_LAUNCH()                            \
    do {                                              \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) return (int)e__;      \
    } while (0)

__device__ __forceinline__ double load_k(const void* K, int is_f64, int i) {
    return is_f64 ? ((const double*)K)[i] : (double)((const float*)K)[i];
}

// K^-1 in K's precision (fp64 maths), cast to fp32: `Kinv = K.inverse().float()`
// (geometry/transform.py:92).  Closed-form adjugate instead of the LU the
// library call runs.
__device__ inline void kinv_f64(const void* K, int is_f64, double* out9) {
    double a = load_k(K, is_f64, 0), b = load_k(K, is_f64, 1), c = load_k(K, is_f64, 2);
    double d = load_k(K, is_f64, 3), e = load_k(K, is_f64, 4), f = load_k(K, is_f64, 5);
    double g = load_k(K, is_f64, 6), h = load_k(K, is_f64, 7), i = load_k(K, is_f64, 8);
    double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * Bc + c * C;
    double r = 1.0 / det;
    out9[0] = A * r;
    out9[1] = -(b * i - c * h) * r;
    out9[2] = (b * f - c * e) * r;
    out9[3] = Bc * r;
    out9[4] = (a * i - c * g) * r;
    out9[5] = -(a * f - c * d) * r;
    out9[6] = C * r;
    out9[7] = -(a * h - b * g) * r;
    out9[8] = (a * e - b * d) * r;
}

__device__ inline void kinv_f32(const void* K, int is_f64, float* out9) {
    double t[9];
    kinv_f64(K, is_f64, t);
    for (int i = 0; i < 9; ++i) out9[i] = (float)t[i];
}

// ---------------------------------------------------------------------------
// pose6 (rot3 | trans3) -> M = [R|t] 3x4 row-major, optional rigid inverse.
// axis-angle: geometry/pose_geometry.py:155-199 (axis = v/(|v|+1e-7)), M = T @ R (:139);
// euler: R = Rx @ Ry @ Rz (:38-68); inverse: [R^T | -R^T t] (:110-115).
// ---------------------------------------------------------------------------
__device__ inline void rot_axisangle(const float* r, float* R) {
    float th = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    float den = th + 1e-7f;
    float x = r[0] / den, y = r[1] / den, z = r[2] / den;
    float sa, ca;
    sincosf(th, &sa, &ca);
    float C = 1.0f - ca;
    float xs = x * sa, ys = y * sa, zs = z * sa;
    float xC = x * C, yC = y * C, zC = z * C;
    float xyC = x * yC, yzC = y * zC, zxC = z * xC;
    R[0] = x * xC + ca; R[1] = xyC - zs;    R[2] = zxC + ys;
    R[3] = xyC + zs;    R[4] = y * yC + ca; R[5] = yzC - xs;
    R[6] = zxC - ys;    R[7] = yzC + xs;    R[8] = z * zC + ca;
}

__device__ inline void rot_axisangle_vjp(const float* r, const float* G, float* gr) {
    float th = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    float den = th + 1e-7f;
    float inv = 1.0f / den;
    float x = r[0] / den, y = r[1] / den, z = r[2] / den;
    float sa, ca;
    sincosf(th, &sa, &ca);
    float C = 1.0f - ca;
    float xC = x * C, yC = y * C, zC = z * C;
    float d_ca = G[0] + G[4] + G[8];
    float d_x = G[0] * xC, d_xC = G[0] * x;
    float d_y = G[4] * yC, d_yC = G[4] * y;
    float d_z = G[8] * zC, d_zC = G[8] * z;
    float d_xyC = G[1] + G[3], d_zs = G[3] - G[1];
    float d_zxC = G[2] + G[6], d_ys = G[2] - G[6];
    float d_yzC = G[5] + G[7], d_xs = G[7] - G[5];
    d_x += d_xyC * yC; d_yC += d_xyC * x;
    d_y += d_yzC * zC; d_zC += d_yzC * y;
    d_z += d_zxC * xC; d_xC += d_zxC * z;
    float d_C = d_xC * x + d_yC * y + d_zC * z;
    d_x += d_xC * C; d_y += d_yC * C; d_z += d_zC * C;
    float d_sa = d_xs * x + d_ys * y + d_zs * z;
    d_x += d_xs * sa; d_y += d_ys * sa; d_z += d_zs * sa;
    d_ca -= d_C;
    float d_th = -sa * d_ca + ca * d_sa;
    float d_inv = d_x * r[0] + d_y * r[1] + d_z * r[2];
    d_th += -inv * inv * d_inv;
    // d|r|/dr = r/|r|, and 0 at r = 0 (the subgradient torch.norm uses)
    float s = th > 0.0f ? d_th / th : 0.0f;
    gr[0] = d_x * inv + s * r[0];
    gr[1] = d_y * inv + s * r[1];
    gr[2] = d_z * inv + s * r[2];
}

__device__ inline void mat3_mul(const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// C = A^T B
__device__ inline void mat3_tmul(const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// C = A B^T
__device__ inline void mat3_mult(const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}

__device__ inline void euler_mats(const float* a, float* X, float* Y, float* Z) {
    float sx, cx, sy, cy, sz, cz;
    sincosf(a[0], &sx, &cx);
    sincosf(a[1], &sy, &cy);
    sincosf(a[2], &sz, &cz);
    Z[0] = cz; Z[1] = -sz; Z[2] = 0; Z[3] = sz; Z[4] = cz; Z[5] = 0; Z[6] = 0; Z[7] = 0; Z[8] = 1;
    Y[0] = cy; Y[1] = 0; Y[2] = sy; Y[3] = 0; Y[4] = 1; Y[5] = 0; Y[6] = -sy; Y[7] = 0; Y[8] = cy;
    X[0] = 1; X[1] = 0; X[2] = 0; X[3] = 0; X[4] = cx; X[5] = -sx; X[6] = 0; X[7] = sx; X[8] = cx;
}

__device__ inline void rot_euler(const float* a, float* R) {
    float X[9], Y[9], Z[9], XY[9];
    euler_mats(a, X, Y, Z);
    mat3_mul(X, Y, XY);
    mat3_mul(XY, Z, R);
}

__device__ inline void rot_euler_vjp(const float* a, const float* G, float* ga) {
    float X[9], Y[9], Z[9], XY[9], dXY[9], dZ[9], dX[9], dY[9];
    euler_mats(a, X, Y, Z);
    mat3_mul(X, Y, XY);
    mat3_mult(G, Z, dXY);   // dXY = G Z^T
    mat3_tmul(XY, G, dZ);   // dZ  = XY^T G
    mat3_mult(dXY, Y, dX);  // dX  = dXY Y^T
    mat3_tmul(X, dXY, dY);  // dY  = X^T dXY
    float sx = X[7], cx = X[4], sy = Y[2], cy = Y[0], sz = Z[3], cz = Z[0];
    ga[0] = -sx * dX[4] - cx * dX[5] + cx * dX[7] - sx * dX[8];
    ga[1] = -sy * dY[0] + cy * dY[2] - cy * dY[6] - sy * dY[8];
    ga[2] = -sz * dZ[0] - cz * dZ[1] + cz * dZ[3] - sz * dZ[4];
}

// M (3x4 row-major) from pose6.
__device__ inline void pose_to_M(const float* p, int rotation_mode, int invert, float* M) {
    float R[9];
    if (rotation_mode == PLB_ROT_EULER) rot_euler(p, R); else rot_axisangle(p, R);
    const float* t = p + 3;
    if (!invert) {
        for (int i = 0; i < 3; ++i) {
            M[i * 4 + 0] = R[i * 3 + 0]; M[i * 4 + 1] = R[i * 3 + 1]; M[i * 4 + 2] = R[i * 3 + 2];
            M[i * 4 + 3] = t[i];
        }
    } else {
        for (int i = 0; i < 3; ++i) {
            M[i * 4 + 0] = R[0 * 3 + i]; M[i * 4 + 1] = R[1 * 3 + i]; M[i * 4 + 2] = R[2 * 3 + i];
            M[i * 4 + 3] = (-R[0 * 3 + i]) * t[0] + (-R[1 * 3 + i]) * t[1] + (-R[2 * 3 + i]) * t[2];
        }
    }
}

// vjp of pose_to_M: dM (3x4) -> g6.
__device__ inline void pose_to_M_vjp(const float* p, int rotation_mode, int invert, const float* dM,
                                     float* g6) {
    float R[9], dR[9], dt[3];
    if (rotation_mode == PLB_ROT_EULER) rot_euler(p, R); else rot_axisangle(p, R);
    const float* t = p + 3;
    if (!invert) {
        for (int i = 0; i < 3; ++i) {
            dR[i * 3 + 0] = dM[i * 4 + 0]; dR[i * 3 + 1] = dM[i * 4 + 1]; dR[i * 3 + 2] = dM[i * 4 + 2];
            dt[i] = dM[i * 4 + 3];
        }
    } else {
        // M_R[i][j] = R[j][i];  M_t[i] = -sum_j R[j][i] t[j]
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int i = 0; i < 3; ++i) {
                dR[j * 3 + i] = dM[i * 4 + j] - dM[i * 4 + 3] * t[j];
                acc += R[j * 3 + i] * dM[i * 4 + 3];
            }
            dt[j] = -acc;
        }
    }
    if (rotation_mode == PLB_ROT_EULER) rot_euler_vjp(p, dR, g6); else rot_axisangle_vjp(p, dR, g6);
    g6[3] = dt[0]; g6[4] = dt[1]; g6[5] = dt[2];
}

// P = K(f32) @ M, 3x4 (geometry/transform.py:137-139; K_hom is fp32, :108-111).
__device__ inline void k_times_M(const void* K, int is_f64, const float* M, float* P) {
    float Kf[9];
    for (int i = 0; i < 9; ++i) Kf[i] = (float)load_k(K, is_f64, i);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
            P[r * 4 + c] = Kf[r * 3] * M[c] + Kf[r * 3 + 1] * M[4 + c] + Kf[r * 3 + 2] * M[8 + c];
}
// dM = K^T dP
__device__ inline void kT_times_dP(const void* K, int is_f64, const float* dP, float* dM) {
    float Kf[9];
    for (int i = 0; i < 9; ++i) Kf[i] = (float)load_k(K, is_f64, i);
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 4; ++c)
            dM[k * 4 + c] = Kf[k] * dP[c] + Kf[3 + k] * dP[4 + c] + Kf[6 + k] * dP[8 + c];
}

// ---------------------------------------------------------------------------
// Projection of one back-projected pixel and its bilinear footprint.
// ---------------------------------------------------------------------------
struct Taps {
    float wx0, wx1, wy0, wy1;  // (x1-ix), (ix-x0), (y1-iy), (iy-y0)
    int x0, y0;                // north-west tap (may be -1)
    bool any;                  // false: every tap is out of the image (or coords not finite)
    bool vx0, vx1, vy0, vy1;
};

// cam = P . (X,1);  pix = cam_xy / (cam_z + 1e-5);  then the reference's
// normalise (geometry/transform.py:143-148) and grid_sample's un-normalise
// (align_corners=True) in the same operation order, so floor() sees the same value.
__device__ __forceinline__ void project_pixel(const float* __restrict__ P, float X, float Y, float Z,
                                              float wm1, float hm1, float& cx, float& cy, float& ze,
                                              float& ix, float& iy) {
    cx = fmaf(P[2], Z, fmaf(P[1], Y, P[0] * X)) + P[3];
    cy = fmaf(P[6], Z, fmaf(P[5], Y, P[4] * X)) + P[7];
    float cz = fmaf(P[10], Z, fmaf(P[9], Y, P[8] * X)) + P[11];
    ze = cz + 1e-5f;
    float px = __fdiv_rn(cx, ze), py = __fdiv_rn(cy, ze);
    float gx = __fmul_rn(__fsub_rn(__fdiv_rn(px, wm1), 0.5f), 2.0f);
    float gy = __fmul_rn(__fsub_rn(__fdiv_rn(py, hm1), 0.5f), 2.0f);
    ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), wm1);
    iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), hm1);
}

__device__ __forceinline__ void make_taps(float ix, float iy, int W, int H, Taps& t) {
    // also false for NaN / inf
    t.any = (ix > -1.0f) && (ix < (float)W) && (iy > -1.0f) && (iy < (float)H);
    float x0f = floorf(ix), y0f = floorf(iy);
    t.wx1 = ix - x0f; t.wx0 = (x0f + 1.0f) - ix;
    t.wy1 = iy - y0f; t.wy0 = (y0f + 1.0f) - iy;
    if (!t.any) { t.wx0 = t.wx1 = t.wy0 = t.wy1 = 0.0f; }  // keeps inf/NaN coordinates out of the sums
    t.x0 = t.any ? (int)x0f : 0;
    t.y0 = t.any ? (int)y0f : 0;
    t.vx0 = t.any && t.x0 >= 0;
    t.vx1 = t.any && t.x0 + 1 <= W - 1;
    t.vy0 = t.any && t.y0 >= 0;
    t.vy1 = t.any && t.y0 + 1 <= H - 1;
}

// 1/z: MUFU.RCP (1 ulp) + one Newton step; z = 0 gives inf/NaN, which callers clamp away.
__device__ __forceinline__ float rcp_nr(float z) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
    return r * fmaf(-z, r, 2.0f);
}

// Disparity head folded into the loss (PLB_INPUT_LOGIT): disp = alpha * sigmoid(x) + beta (models/depth/disp_net.py:121).
__device__ __forceinline__ float head_disp(float x, float alpha, float beta) {
    return fmaf(alpha, rcp_nr(1.0f + expf(-x)), beta);
}
// d disp / d x = alpha * s * (1 - s), from the DEPTH the kernels keep: disp = (1/D - b) / a, s = (disp - beta) / alpha
__device__ __forceinline__ float head_chain_from_depth(float D, float a, float b, float alpha, float beta) {
    const float d = (rcp_nr(D) - b) / a;
    return (d - beta) * (alpha + beta - d) / alpha;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float ldg_pred(const float* __restrict__ p, bool ok) {
    return ok ? __ldg(p) : 0.0f;
}

// F.interpolate(.., mode='bilinear', align_corners=False) source rule
// (losses.py:215): src = max(scale*(dst+0.5)-0.5, 0), scale = in/out.
__device__ __forceinline__ void up_coord(int dst, float scale, int in_size, int& i0, int& i1, float& l0,
                                         float& l1) {
    float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.0f);
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = fminf(fmaxf(src - (float)i0, 0.0f), 1.0f);
    l0 = 1.0f - l1;
}

// Backward relaunch guard: true when every upstream gradient equals 1, i.e. the
// gradients the forward pass already wrote (unit upstream) are exact.
__device__ __forceinline__ bool skip_launch(const float* const (&flags)[2]) {
    if (flags[0] == nullptr && flags[1] == nullptr) return false;
    if (flags[0] != nullptr && __ldg(flags[0]) != 1.0f) return false;
    if (flags[1] != nullptr && __ldg(flags[1]) != 1.0f) return false;
    return true;
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Reduce 16 per-lane values across the warp with 16 shuffles (instead of 80):
// every step each lane hands half of its live values to its partner.  On return
// lane L holds in v[0] the full sum of value index idx16(L) (see below); lanes L
// and L^1... the mapping is: bit4 of L picks the upper half first, etc.
// Returns the value index held by this lane in `which` (0..15); two lanes
// (L and L with bit0 flipped... ) - both lanes of each pair hold the same sum.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane, int& which) {
    // step 1: partner = lane ^ 16 ; keep 8
    {
        bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float send = up ? v[k] : v[k + 8];
            float keep = up ? v[k + 8] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float send = up ? v[k] : v[k + 4];
            float keep = up ? v[k + 4] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float send = up ? v[k] : v[k + 2];
            float keep = up ? v[k + 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        bool up = lane & 2;
        float send = up ? v[0] : v[1];
        float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    which = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
    return v[0];
}

}  // namespace plb
