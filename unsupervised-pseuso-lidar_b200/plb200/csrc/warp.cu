// Stand-alone geometry entry points: inverse_warp forward / vjp
// (geometry/pose_geometry.py:201-228), Transform.reconstruct / project
// (geometry/transform.py:74-150), pose matrices and disp_to_depth.  The fused
// loss (photo.cu) does not call these; they exist so that the reference's
// individual functions keep working as drop-ins.
#include "common.cuh"

namespace plb {

constexpr int WP_TILE_W = 64, WP_TILE_H = 16, WP_THREADS = 256, WP_ROWS = 4;

struct WarpLayout {
    size_t tickets, partials, total;
    int tiles_x, tiles;
};

__host__ __device__ inline WarpLayout warp_layout(const plb_warp_args& a) {
    WarpLayout L;
    L.tiles_x = (a.W + WP_TILE_W - 1) / WP_TILE_W;
    L.tiles = L.tiles_x * ((a.H + WP_TILE_H - 1) / WP_TILE_H);
    L.tickets = 0;
    L.partials = ((size_t)a.B * 4 + 255) / 256 * 256;
    L.total = L.partials + ((size_t)a.B * L.tiles * 12 * 4 + 255) / 256 * 256;
    return L;
}

template <bool BWD>
__global__ void __launch_bounds__(WP_THREADS)
warp_kernel(const __grid_constant__ plb_warp_args a) {
    const WarpLayout L = warp_layout(a);
    const int b = blockIdx.y, tile = blockIdx.x;
    const int H = a.H, W = a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)H * W;

    __shared__ float s_kinv[9];
    __shared__ float s_P[12];
    __shared__ float s_acc[WP_THREADS / 32][12];
    __shared__ float s_red[12];
    __shared__ int s_flag;

    const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
    const float* pose_b = a.pose + (size_t)b * a.pose_stride;
    if (tid == 0) kinv_f32(Kb, a.k_is_f64, s_kinv);
    if (tid == 32) {
        float M[12];
        pose_to_M(pose_b, a.rotation_mode, a.pose_inv, M);
        k_times_M(Kb, a.k_is_f64, M, s_P);
    }
    __syncthreads();

    const int x = (tile % L.tiles_x) * WP_TILE_W + (warp & 1) * 32 + lane;
    const int ybase = (tile / L.tiles_x) * WP_TILE_H + (warp >> 1) * WP_ROWS;
    const float* img = a.img + (size_t)b * 3 * plane;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;

#pragma unroll
    for (int j = 0; j < WP_ROWS; ++j) {
        const int y = ybase + j;
        const bool valid = x < W && y < H;
        const size_t o = (size_t)min(y, H - 1) * W + min(x, W - 1);
        const float D = __ldg(a.depth + (size_t)b * plane + o);
        const float xf = (float)x, yf = (float)y;
        const float rx = fmaf(s_kinv[1], yf, s_kinv[0] * xf) + s_kinv[2];
        const float ry = fmaf(s_kinv[4], yf, s_kinv[3] * xf) + s_kinv[5];
        const float rz = fmaf(s_kinv[7], yf, s_kinv[6] * xf) + s_kinv[8];
        const float X = rx * D, Y = ry * D, Z = rz * D;
        float cx, cy, ze, ix, iy;
        project_pixel(s_P, X, Y, Z, (float)(W - 1), (float)(H - 1), cx, cy, ze, ix, iy);
        Taps tp;
        make_taps(ix, iy, W, H, tp);
        const int xa = max(tp.x0, 0), xb = min(tp.x0 + 1, W - 1);
        const int ya = max(tp.y0, 0), yb = min(tp.y0 + 1, H - 1);
        const bool mnw = valid && tp.vx0 && tp.vy0, mne = valid && tp.vx1 && tp.vy0;
        const bool msw = valid && tp.vx0 && tp.vy1, mse = valid && tp.vx1 && tp.vy1;
        const float wnw = tp.wx0 * tp.wy0, wne = tp.wx1 * tp.wy0, wsw = tp.wx0 * tp.wy1, wse = tp.wx1 * tp.wy1;
        const float* r0 = img + (size_t)ya * W;
        const float* r1 = img + (size_t)yb * W;
        float Gx = 0.0f, Gy = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float vnw = ldg_pred(r0 + c * plane + xa, mnw), vne = ldg_pred(r0 + c * plane + xb, mne);
            const float vsw = ldg_pred(r1 + c * plane + xa, msw), vse = ldg_pred(r1 + c * plane + xb, mse);
            if (!BWD) {
                if (valid) a.out[(size_t)b * 3 * plane + c * plane + o] = vnw * wnw + vne * wne + vsw * wsw + vse * wse;
            } else {
                const float e = valid ? __ldg(a.g_out + (size_t)b * 3 * plane + c * plane + o) : 0.0f;
                Gx += e * ((vne - vnw) * tp.wy0 + (vse - vsw) * tp.wy1);
                Gy += e * ((vsw - vnw) * tp.wx0 + (vse - vne) * tp.wx1);
                if (a.g_img != nullptr) {
                    float* q0 = a.g_img + (size_t)b * 3 * plane + c * plane + (size_t)ya * W;
                    float* q1 = a.g_img + (size_t)b * 3 * plane + c * plane + (size_t)yb * W;
                    if (mnw) atomicAdd(q0 + xa, wnw * e);
                    if (mne) atomicAdd(q0 + xb, wne * e);
                    if (msw) atomicAdd(q1 + xa, wsw * e);
                    if (mse) atomicAdd(q1 + xb, wse * e);
                }
            }
        }
        if (BWD) {
            const float iz = 1.0f / ze;
            const float px = cx * iz, py = cy * iz;
            float gcx = Gx * iz, gcy = Gy * iz, gcz = -(Gx * px + Gy * py) * iz;
            if (!(valid && tp.any)) { gcx = 0.0f; gcy = 0.0f; gcz = 0.0f; }
            if (a.g_depth != nullptr && valid)
                a.g_depth[(size_t)b * plane + o] =
                    gcx * (s_P[0] * rx + s_P[1] * ry + s_P[2] * rz) + gcy * (s_P[4] * rx + s_P[5] * ry + s_P[6] * rz) +
                    gcz * (s_P[8] * rx + s_P[9] * ry + s_P[10] * rz);
            acc[0] += gcx * X; acc[1] += gcx * Y; acc[2] += gcx * Z; acc[3] += gcx;
            acc[4] += gcy * X; acc[5] += gcy * Y; acc[6] += gcy * Z; acc[7] += gcy;
            acc[8] += gcz * X; acc[9] += gcz * Y; acc[10] += gcz * Z; acc[11] += gcz;
        }
    }
    if (!BWD) return;
    if (a.g_pose == nullptr) return;

    int which;
    const float r = warp_reduce16(acc, lane, which);
    if ((lane & 1) == 0 && which < 12) s_acc[warp][which] = r;
    __syncthreads();
    int32_t* tickets = (int32_t*)((char*)a.workspace + L.tickets);
    float* partials = (float*)((char*)a.workspace + L.partials);
    if (tid < 12) {
        float v = 0.0f;
        for (int w = 0; w < WP_THREADS / 32; ++w) v += s_acc[w][tid];
        __stcg(partials + ((size_t)b * L.tiles + tile) * 12 + tid, v);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(&tickets[b], 1) == L.tiles - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (tid < 12) {
        float v = 0.0f;
        for (int t = 0; t < L.tiles; ++t) v += __ldcg(partials + ((size_t)b * L.tiles + t) * 12 + tid);
        s_red[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        float dP[12], dM[12], g6[6];
        for (int k = 0; k < 12; ++k) dP[k] = s_red[k];
        kT_times_dP(Kb, a.k_is_f64, dP, dM);
        pose_to_M_vjp(pose_b, a.rotation_mode, a.pose_inv, dM, g6);
        for (int k = 0; k < 6; ++k) a.g_pose[(size_t)b * 6 + k] = g6[k];
        tickets[b] = 0;
    }
}

static int validate_warp(const plb_warp_args* a, bool bwd) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 2 || a->W < 2 || a->pose_stride < 6) return PLB_EINVAL;
    if (a->rotation_mode != PLB_ROT_AXISANGLE && a->rotation_mode != PLB_ROT_EULER) return PLB_EINVAL;
    if (!a->img || !a->depth || !a->pose || !a->K) return PLB_ENULL;
    if (!bwd && !a->out) return PLB_ENULL;
    if (bwd) {
        if (!a->g_out) return PLB_ENULL;
        if (a->g_pose && (!a->workspace || a->workspace_bytes < warp_layout(*a).total)) return PLB_EWORKSPACE;
    }
    return PLB_OK;
}

int warp_forward_launch(const plb_warp_args* a, cudaStream_t st) {
    int rc = validate_warp(a, false);
    if (rc) return rc;
    const WarpLayout L = warp_layout(*a);
    warp_kernel<false><<<dim3(L.tiles, a->B), WP_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int warp_backward_launch(const plb_warp_args* a, cudaStream_t st) {
    int rc = validate_warp(a, true);
    if (rc) return rc;
    const WarpLayout L = warp_layout(*a);
    warp_kernel<true><<<dim3(L.tiles, a->B), WP_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t warp_workspace_bytes(const plb_warp_args* a) { return warp_layout(*a).total; }

// --------------------------------------------------------------------------
__global__ void reconstruct_kernel(const float* __restrict__ depth, const void* K, int k_is_f64, int B, int H, int W,
                                   float* __restrict__ Xc) {
    __shared__ float s_kinv[9];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) kinv_f32((const char*)K + (size_t)b * 9 * (k_is_f64 ? 8 : 4), k_is_f64, s_kinv);
    __syncthreads();
    const size_t plane = (size_t)H * W;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += (size_t)gridDim.x * blockDim.x) {
        const float xf = (float)(p % W), yf = (float)(p / W);
        const float D = __ldg(depth + b * plane + p);
        Xc[(b * 3 + 0) * plane + p] = (fmaf(s_kinv[1], yf, s_kinv[0] * xf) + s_kinv[2]) * D;
        Xc[(b * 3 + 1) * plane + p] = (fmaf(s_kinv[4], yf, s_kinv[3] * xf) + s_kinv[5]) * D;
        Xc[(b * 3 + 2) * plane + p] = (fmaf(s_kinv[7], yf, s_kinv[6] * xf) + s_kinv[8]) * D;
    }
}

__global__ void project_kernel(const float* __restrict__ X, const void* K, int k_is_f64, const float* __restrict__ Tcw,
                               int B, int H, int W, float* __restrict__ grid) {
    __shared__ float s_P[12];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) k_times_M((const char*)K + (size_t)b * 9 * (k_is_f64 ? 8 : 4), k_is_f64, Tcw + (size_t)b * 16, s_P);
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += (size_t)gridDim.x * blockDim.x) {
        const float x = __ldg(X + (b * 3 + 0) * plane + p), y = __ldg(X + (b * 3 + 1) * plane + p);
        const float z = __ldg(X + (b * 3 + 2) * plane + p);
        const float cx = fmaf(s_P[2], z, fmaf(s_P[1], y, s_P[0] * x)) + s_P[3];
        const float cy = fmaf(s_P[6], z, fmaf(s_P[5], y, s_P[4] * x)) + s_P[7];
        const float ze = fmaf(s_P[10], z, fmaf(s_P[9], y, s_P[8] * x)) + s_P[11] + 1e-5f;
        const float gx = __fmul_rn(__fsub_rn(__fdiv_rn(__fdiv_rn(cx, ze), wm1), 0.5f), 2.0f);
        const float gy = __fmul_rn(__fsub_rn(__fdiv_rn(__fdiv_rn(cy, ze), hm1), 0.5f), 2.0f);
        reinterpret_cast<float2*>(grid)[b * plane + p] = make_float2(gx, gy);
    }
}

__global__ void pose_matrix_kernel(const float* pose, int stride, int B, int mode, int invert, float* M44) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float M[12];
    pose_to_M(pose + (size_t)b * stride, mode, invert, M);
    float* o = M44 + (size_t)b * 16;
    for (int k = 0; k < 12; ++k) o[k] = M[k];
    o[12] = 0.0f; o[13] = 0.0f; o[14] = 0.0f; o[15] = 1.0f;
}

__global__ void pose_matrix_bwd_kernel(const float* pose, int stride, int B, int mode, int invert, const float* gM44,
                                       float* g_pose) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float g6[6];
    pose_to_M_vjp(pose + (size_t)b * stride, mode, invert, gM44 + (size_t)b * 16, g6);
    for (int k = 0; k < 6; ++k) g_pose[(size_t)b * 6 + k] = g6[k];
}

__global__ void disp_to_depth_kernel(const float* __restrict__ disp, int64_t n, float a, float b,
                                     float* __restrict__ depth) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        depth[i] = 1.0f / (a * __ldg(disp + i) + b);
}

__global__ void disp_to_depth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ g, int64_t n, float a,
                                         float b, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float D = 1.0f / (a * __ldg(disp + i) + b);
        out[i] = __ldg(g + i) * (-a * D * D);
    }
}

static inline int blocks_for(int64_t n, int threads, int cap) {
    int64_t nb = (n + threads - 1) / threads;
    return (int)(nb < cap ? (nb < 1 ? 1 : nb) : cap);
}

int reconstruct_launch(const float* depth, const void* K, int k64, int B, int H, int W, float* Xc, cudaStream_t st) {
    if (!depth || !K || !Xc) return PLB_ENULL;
    if (B < 1 || H < 1 || W < 1) return PLB_EINVAL;
    reconstruct_kernel<<<dim3(blocks_for((int64_t)H * W, 256, 148 * 8), B), 256, 0, st>>>(depth, K, k64, B, H, W, Xc);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int project_launch(const float* X, const void* K, int k64, const float* Tcw, int B, int H, int W, float* grid,
                   cudaStream_t st) {
    if (!X || !K || !Tcw || !grid) return PLB_ENULL;
    if (B < 1 || H < 2 || W < 2) return PLB_EINVAL;
    project_kernel<<<dim3(blocks_for((int64_t)H * W, 256, 148 * 8), B), 256, 0, st>>>(X, K, k64, Tcw, B, H, W, grid);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int pose_matrix_launch(const float* pose, int stride, int B, int mode, int invert, float* M44, cudaStream_t st) {
    if (!pose || !M44) return PLB_ENULL;
    if (B < 1 || stride < 6 || (mode != PLB_ROT_AXISANGLE && mode != PLB_ROT_EULER)) return PLB_EINVAL;
    pose_matrix_kernel<<<(B + 63) / 64, 64, 0, st>>>(pose, stride, B, mode, invert, M44);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int pose_matrix_bwd_launch(const float* pose, int stride, int B, int mode, int invert, const float* gM, float* g_pose,
                           cudaStream_t st) {
    if (!pose || !gM || !g_pose) return PLB_ENULL;
    if (B < 1 || stride < 6 || (mode != PLB_ROT_AXISANGLE && mode != PLB_ROT_EULER)) return PLB_EINVAL;
    pose_matrix_bwd_kernel<<<(B + 63) / 64, 64, 0, st>>>(pose, stride, B, mode, invert, gM, g_pose);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int disp_to_depth_launch(const float* disp, int64_t n, float a, float b, float* depth, cudaStream_t st) {
    if (!disp || !depth) return PLB_ENULL;
    if (n < 0) return PLB_EINVAL;
    if (n == 0) return PLB_OK;
    disp_to_depth_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, st>>>(disp, n, a, b, depth);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

int disp_to_depth_bwd_launch(const float* disp, const float* g, int64_t n, float a, float b, float* out,
                             cudaStream_t st) {
    if (!disp || !g || !out) return PLB_ENULL;
    if (n < 0) return PLB_EINVAL;
    if (n == 0) return PLB_OK;
    disp_to_depth_bwd_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, st>>>(disp, g, n, a, b, out);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
