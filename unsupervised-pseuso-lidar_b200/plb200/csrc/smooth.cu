// Second-order depth smoothness (losses.py:242-260), all scales of the pyramid,
// loss and gradient in one launch.  disp -> depth is folded in.
//
// One block = one 64x16 tile of one scale of one image.  The depth tile plus a
// 2-pixel halo is staged in shared memory once; every pixel then evaluates the
// four second differences anchored on it (forward sums) and gathers the signs
// of the <= 14 second differences it takes part in (gradient) - no atomics, no
// scatter, bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int SM_TW = 64, SM_TH = 16, SM_HALO = 2, SM_THREADS = 256, SM_ROWS = 4;
constexpr int SM_SW = SM_TW + 2 * SM_HALO;  // 68
constexpr int SM_SH = SM_TH + 2 * SM_HALO;  // 20

struct SmoothLayout {
    size_t ticket;    // int32[1]
    size_t partials;  // float [blocks][4]
    size_t total;
    int tiles_x[PLB_MAX_SCALES], tiles[PLB_MAX_SCALES], first_block[PLB_MAX_SCALES + 1];
};

__host__ __device__ inline SmoothLayout smooth_layout(const plb_smooth_args& a) {
    SmoothLayout L;
    int nb = 0;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) {
        L.first_block[s] = nb;
        if (s < a.n_scales) {
            L.tiles_x[s] = (a.dw[s] + SM_TW - 1) / SM_TW;
            L.tiles[s] = L.tiles_x[s] * ((a.dh[s] + SM_TH - 1) / SM_TH);
            nb += L.tiles[s] * a.B;
        } else {
            L.tiles_x[s] = L.tiles[s] = 0;
        }
    }
    L.first_block[PLB_MAX_SCALES] = nb;
    L.ticket = 0;
    L.partials = 256;
    L.total = 256 + ((size_t)nb * 4 * sizeof(float) + 255) / 256 * 256;
    return L;
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(SM_THREADS)
smooth_kernel(const __grid_constant__ plb_smooth_args a) {
    if (skip_launch(a.skip_if_unit, a.skip_n)) return;
    const SmoothLayout L = smooth_layout(a);
    int32_t* ticket = (int32_t*)((char*)a.workspace + L.ticket);
    float* partials = (float*)((char*)a.workspace + L.partials);

    int s = 0;
    while (s + 1 < a.n_scales && (int)blockIdx.x >= L.first_block[s + 1]) ++s;
    const int local = blockIdx.x - L.first_block[s];
    const int b = local / L.tiles[s], tile = local % L.tiles[s];
    const int h = a.dh[s], w = a.dw[s];
    const int tx0 = (tile % L.tiles_x[s]) * SM_TW, ty0 = (tile / L.tiles_x[s]) * SM_TH;
    const float* disp = a.disp[s] + (size_t)b * h * w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ float sD[SM_SH][SM_SW];
    __shared__ float s_part[SM_THREADS / 32][4];
    __shared__ double s_fin[SM_THREADS];
    __shared__ int s_flag;

    for (int k = tid; k < SM_SH * SM_SW; k += SM_THREADS) {
        const int ly = k / SM_SW, lx = k - ly * SM_SW;
        const int gy = ty0 + ly - SM_HALO, gx = tx0 + lx - SM_HALO;
        float v = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
            v = __ldg(disp + (size_t)gy * w + gx);
            if (!a.input_is_depth) v = 1.0f / (a.disp_a * v + a.disp_b);
        }
        sD[ly][lx] = v;
    }
    __syncthreads();

    // weights of the four terms: weight_s / element count of each difference map
    float wscale = 1.0f;
    for (int k = 0; k < s; ++k) wscale /= a.scale_decay;
    const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
    const float n1 = (float)a.B * (float)h * (float)(w - 2);
    const float n2 = (float)a.B * (float)(h - 1) * (float)(w - 1);
    const float n3 = (float)a.B * (float)(h - 2) * (float)w;
    const float c1 = wscale * up / n1, c2 = wscale * up / n2, c3 = wscale * up / n3;

    const int lx = (warp & 1) * 32 + lane + SM_HALO;
    const int x = tx0 + (warp & 1) * 32 + lane;
    float sum_dx2 = 0.0f, sum_dxdy = 0.0f, sum_dydx = 0.0f, sum_dy2 = 0.0f;
#define DD(yy, xx) sD[ly + (yy)][lx + (xx)]
#define DX2(yy, xx) ((DD(yy, (xx) + 2) - DD(yy, (xx) + 1)) - (DD(yy, (xx) + 1) - DD(yy, xx)))
#define DY2(yy, xx) ((DD((yy) + 2, xx) - DD((yy) + 1, xx)) - (DD((yy) + 1, xx) - DD(yy, xx)))
#define DXDY(yy, xx) ((DD((yy) + 1, (xx) + 1) - DD((yy) + 1, xx)) - (DD(yy, (xx) + 1) - DD(yy, xx)))
#define DYDX(yy, xx) ((DD((yy) + 1, (xx) + 1) - DD(yy, (xx) + 1)) - (DD((yy) + 1, xx) - DD(yy, xx)))
#pragma unroll
    for (int j = 0; j < SM_ROWS; ++j) {
        const int ly = (warp >> 1) * SM_ROWS + j + SM_HALO;
        const int y = ty0 + (warp >> 1) * SM_ROWS + j;
        if (x >= w || y >= h) continue;
        if (x <= w - 3) sum_dx2 += fabsf(DX2(0, 0));
        if (y <= h - 3) sum_dy2 += fabsf(DY2(0, 0));
        if (x <= w - 2 && y <= h - 2) {
            sum_dxdy += fabsf(DXDY(0, 0));
            sum_dydx += fabsf(DYDX(0, 0));
        }
        if (a.want_grad && a.g_disp[s] != nullptr) {
            float g = 0.0f;
            // d^2/dx^2 terms anchored at x, x-1, x-2
            if (x <= w - 3) g += c1 * sgnf(DX2(0, 0));
            if (x >= 1 && x <= w - 2) g -= 2.0f * c1 * sgnf(DX2(0, -1));
            if (x >= 2) g += c1 * sgnf(DX2(0, -2));
            if (y <= h - 3) g += c3 * sgnf(DY2(0, 0));
            if (y >= 1 && y <= h - 2) g -= 2.0f * c3 * sgnf(DY2(-1, 0));
            if (y >= 2) g += c3 * sgnf(DY2(-2, 0));
            // mixed terms anchored at (x,y), (x-1,y), (x,y-1), (x-1,y-1)
            const bool xr = x <= w - 2, xl = x >= 1, yd = y <= h - 2, yu = y >= 1;
            if (xr && yd) g += c2 * (sgnf(DXDY(0, 0)) + sgnf(DYDX(0, 0)));
            if (xl && yd) g -= c2 * (sgnf(DXDY(0, -1)) + sgnf(DYDX(0, -1)));
            if (xr && yu) g -= c2 * (sgnf(DXDY(-1, 0)) + sgnf(DYDX(-1, 0)));
            if (xl && yu) g += c2 * (sgnf(DXDY(-1, -1)) + sgnf(DYDX(-1, -1)));
            if (!a.input_is_depth) { const float D = DD(0, 0); g *= -a.disp_a * D * D; }
            float* o = a.g_disp[s] + (size_t)b * h * w + (size_t)y * w + x;
            if (a.accumulate) *o += g; else *o = g;
        }
    }
#undef DD
#undef DX2
#undef DY2
#undef DXDY
#undef DYDX
    sum_dx2 = warp_sum(sum_dx2); sum_dxdy = warp_sum(sum_dxdy);
    sum_dydx = warp_sum(sum_dydx); sum_dy2 = warp_sum(sum_dy2);
    if (lane == 0) { s_part[warp][0] = sum_dx2; s_part[warp][1] = sum_dxdy; s_part[warp][2] = sum_dydx; s_part[warp][3] = sum_dy2; }
    __syncthreads();
    if (tid < 4) {
        float v = 0.0f;
        for (int k = 0; k < SM_THREADS / 32; ++k) v += s_part[k][tid];
        // pre-scale by weight/count so the final pass is a plain sum
        const float c = tid == 0 ? wscale / n1 : (tid == 3 ? wscale / n3 : wscale / n2);
        __stcg(partials + (size_t)blockIdx.x * 4 + tid, v * c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double acc = 0.0;
    for (int k = tid; k < (int)gridDim.x * 4; k += SM_THREADS) acc += (double)__ldcg(partials + k);
    s_fin[tid] = acc;
    __syncthreads();
    for (int st = SM_THREADS / 2; st > 0; st >>= 1) {
        if (tid < st) s_fin[tid] += s_fin[tid + st];
        __syncthreads();
    }
    if (tid == 0) {
        if (a.loss != nullptr) *a.loss = (float)s_fin[0];
        *ticket = 0;
    }
}

int smooth_launch(const plb_smooth_args* a, cudaStream_t st) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->n_scales < 1 || a->n_scales > PLB_MAX_SCALES) return PLB_EINVAL;
    for (int s = 0; s < a->n_scales; ++s) {
        if (a->disp[s] == nullptr) return PLB_ENULL;
        if (a->dh[s] < 3 || a->dw[s] < 3) return PLB_EINVAL;
    }
    if (a->loss == nullptr) return PLB_ENULL;
    const SmoothLayout L = smooth_layout(*a);
    if (a->workspace == nullptr || a->workspace_bytes < L.total) return PLB_EWORKSPACE;
    smooth_kernel<<<L.first_block[PLB_MAX_SCALES], SM_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t smooth_workspace_bytes(const plb_smooth_args* a) { return smooth_layout(*a).total; }

}  // namespace plb
