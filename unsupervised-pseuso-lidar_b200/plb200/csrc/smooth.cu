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

// sign codes of the four second differences anchored on one pixel, 2 bits each (0:-1, 1:0, 2:+1)
__device__ __forceinline__ unsigned sgn2(float v) { return v > 0.0f ? 2u : (v < 0.0f ? 0u : 1u); }
__device__ __forceinline__ float unsgn(unsigned code, int shift) { return (float)(int)((code >> shift) & 3u) - 1.0f; }

__global__ void __launch_bounds__(SM_THREADS)
smooth_kernel(const __grid_constant__ plb_smooth_args a, int n_tiles) {
    if (skip_launch(a.skip_if_unit)) return;
    const SmoothLayout L = smooth_layout(a);
    int32_t* ticket = (int32_t*)((char*)a.workspace + L.ticket);
    float* partials = (float*)((char*)a.workspace + L.partials);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ float sD[SM_SH][SM_SW];                       // depth, tile + 2-pixel halo
    // signs of the second differences anchored at (x-2.., y-2..): d2/dx2, d2/dy2, dxdy + dydx
    __shared__ signed char sS1[SM_TH + SM_HALO][SM_TW + SM_HALO + 2], sS3[SM_TH + SM_HALO][SM_TW + SM_HALO + 2],
        sSm[SM_TH + SM_HALO][SM_TW + SM_HALO + 2];
    __shared__ double s_fin[SM_THREADS];
    __shared__ int s_flag;

    const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
    double total = 0.0;   // this thread's share of the (weight/count-scaled) forward sums

    // persistent: block k takes tiles k, k + grid, ... (a tile = 64x16 pixels of one scale of one image)
#pragma unroll 1
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int s = 0;
        while (s + 1 < a.n_scales && t >= L.first_block[s + 1]) ++s;
        const int local = t - L.first_block[s];
        const int b = local / L.tiles[s], tile = local % L.tiles[s];
        const int h = a.dh[s], w = a.dw[s];
        const int tx0 = (tile % L.tiles_x[s]) * SM_TW, ty0 = (tile / L.tiles_x[s]) * SM_TH;
        const float* disp = a.disp[s] + (size_t)b * h * w;   // (image planes stay below 2^31 elements: checked at launch)
        const bool want = a.want_grad && a.g_disp[s] != nullptr;

        // issue every global read of the tile up front: the depth tile and (accumulate mode) the
        // gradient values this thread will add to - one memory round trip per tile instead of two
        const int px = tx0 + (warp & 1) * 32 + lane;
        float old[SM_ROWS];
#pragma unroll
        for (int j = 0; j < SM_ROWS; ++j) {
            const int y = ty0 + (warp >> 1) * SM_ROWS + j;
            old[j] = (want && a.accumulate && px < w && y < h)
                         ? __ldcg(a.g_disp[s] + (b * h * w + y * w + px)) : 0.0f;
        }
        // tile + halo: warp w takes rows w, w+8, w+16; lanes take columns lane, lane+32, lane+64
        for (int ly = warp; ly < SM_SH; ly += SM_THREADS / 32) {
            const int gy = ty0 + ly - SM_HALO;
            const bool rowin = gy >= 0 && gy < h;
            const float* drow = disp + gy * w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int lx = lane + 32 * c;
                if (lx < SM_SW) {
                    const int gx = tx0 + lx - SM_HALO;
                    float v = 0.0f;
                    if (rowin && gx >= 0 && gx < w) {
                        v = __ldg(drow + gx);
                        if (!a.input_is_depth) v = rcp_nr(fmaf(a.disp_a, v, a.disp_b));
                    }
                    sD[ly][lx] = v;
                }
            }
        }
        __syncthreads();

        // weights of the four terms: weight_s / element count of each difference map
        float wscale = 1.0f;
        for (int k = 0; k < s; ++k) wscale /= a.scale_decay;
        const float n1 = (float)a.B * (float)h * (float)(w - 2);
        const float n2 = (float)a.B * (float)(h - 1) * (float)(w - 1);
        const float n3 = (float)a.B * (float)(h - 2) * (float)w;
        const float c1 = wscale / n1, c2 = wscale / n2, c3 = wscale / n3;

        // ---- pass 1: every anchor (tile + the 2 columns / rows before it) evaluates its four
        // second differences once: |.| -> forward sums (anchors inside the tile), signs -> sS planes
        float sum_dx2 = 0.0f, sum_mixed = 0.0f, sum_dy2 = 0.0f;
        for (int ay = warp; ay < SM_TH + SM_HALO; ay += SM_THREADS / 32) {
            const int gy = ty0 + ay - SM_HALO;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int ax = lane + 32 * c;
                if (ax < SM_TW + SM_HALO) {
                    const int gx = tx0 + ax - SM_HALO;
                    int s1 = 0, s3 = 0, sm = 0;
                    if (gx >= 0 && gy >= 0 && gx < w && gy < h) {
                        const float d00 = sD[ay][ax], d01 = sD[ay][ax + 1], d02 = sD[ay][ax + 2];
                        const float d10 = sD[ay + 1][ax], d11 = sD[ay + 1][ax + 1], d20 = sD[ay + 2][ax];
                        const bool own = ax >= SM_HALO && ay >= SM_HALO;
                        const float o1 = own ? 1.0f : 0.0f;
                        float v;
                        v = (gx <= w - 3) ? (d02 - d01) - (d01 - d00) : 0.0f;
                        s1 = (v > 0.0f) - (v < 0.0f); sum_dx2 = fmaf(o1, fabsf(v), sum_dx2);
                        v = (gy <= h - 3) ? (d20 - d10) - (d10 - d00) : 0.0f;
                        s3 = (v > 0.0f) - (v < 0.0f); sum_dy2 = fmaf(o1, fabsf(v), sum_dy2);
                        const bool mixed = gx <= w - 2 && gy <= h - 2;
                        v = mixed ? (d11 - d10) - (d01 - d00) : 0.0f;
                        sm = (v > 0.0f) - (v < 0.0f); sum_mixed = fmaf(o1, fabsf(v), sum_mixed);
                        v = mixed ? (d11 - d01) - (d10 - d00) : 0.0f;
                        sm += (v > 0.0f) - (v < 0.0f); sum_mixed = fmaf(o1, fabsf(v), sum_mixed);
                    }
                    sS1[ay][ax] = (signed char)s1; sS3[ay][ax] = (signed char)s3; sSm[ay][ax] = (signed char)sm;
                }
            }
        }
        total += (double)(sum_dx2 * c1) + (double)(sum_mixed * c2) + (double)(sum_dy2 * c3);
        __syncthreads();

        // ---- pass 2: gradient of every tile pixel = signed stencil over the anchors containing it
        if (want) {
            const int lx = (warp & 1) * 32 + lane + SM_HALO;
            const float g1 = c1 * up, g2 = c2 * up, g3 = c3 * up;
#pragma unroll
            for (int j = 0; j < SM_ROWS; ++j) {
                const int ly = (warp >> 1) * SM_ROWS + j + SM_HALO;
                const int y = ty0 + (warp >> 1) * SM_ROWS + j;
                if (px >= w || y >= h) continue;
                const int t1 = (int)sS1[ly][lx] - 2 * (int)sS1[ly][lx - 1] + (int)sS1[ly][lx - 2];
                const int t3 = (int)sS3[ly][lx] - 2 * (int)sS3[ly - 1][lx] + (int)sS3[ly - 2][lx];
                const int tm = (int)sSm[ly][lx] - (int)sSm[ly][lx - 1] - (int)sSm[ly - 1][lx] + (int)sSm[ly - 1][lx - 1];
                float g = fmaf(g1, (float)t1, fmaf(g3, (float)t3, g2 * (float)tm));
                if (!a.input_is_depth) { const float D = sD[ly][lx]; g *= -a.disp_a * D * D; }
                a.g_disp[s][b * h * w + y * w + px] = old[j] + g;
            }
        }
        __syncthreads();   // sD / sS are rewritten by the next tile
    }

    // ---- block partial (fixed order), then the last block sums all partials in double ----------
    s_fin[tid] = total;
    __syncthreads();
    for (int st = SM_THREADS / 2; st > 0; st >>= 1) {
        if (tid < st) s_fin[tid] += s_fin[tid + st];
        __syncthreads();
    }
    if (tid == 0) {
        __stcg((double*)partials + blockIdx.x, s_fin[0]);
        __threadfence();
        s_flag = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double acc0 = 0.0, acc1 = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += 2 * SM_THREADS) {
        acc0 += __ldcg((const double*)partials + k);
        if (k + SM_THREADS < (int)gridDim.x) acc1 += __ldcg((const double*)partials + k + SM_THREADS);
    }
    s_fin[tid] = acc0 + acc1;
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
        if ((long long)a->B * a->dh[s] * a->dw[s] >= (1LL << 31)) return PLB_EINVAL;
    }
    if (a->loss == nullptr) return PLB_ENULL;
    const SmoothLayout L = smooth_layout(*a);
    if (a->workspace == nullptr || a->workspace_bytes < L.total) return PLB_EWORKSPACE;
    const int n_tiles = L.first_block[PLB_MAX_SCALES];
    const int grid = n_tiles < 148 * 4 ? n_tiles : 148 * 4;   // persistent: at most 4 blocks per SM
    smooth_kernel<<<grid, SM_THREADS, 0, st>>>(*a, n_tiles);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t smooth_workspace_bytes(const plb_smooth_args* a) { return smooth_layout(*a).total; }

}  // namespace plb
