// Edge-aware first-order disparity smoothness.  ABSENT from the reference
// (north_star asks for it): the formula is the monodepth2 lineage's that the reference's model files
// cite (models/depth/layers.py:1-2) - PARITY UNPINNED, the oracle is our own torch restatement
// (oracle/restated.py::edge_aware_smooth_loss):
//
//   per scale s (factor f = H / h_s, weight 1 / f):
//     d' = d / (mean_hw(d) + 1e-7)                       (per image; skipped when normalize = 0)
//     I_s = avg_pool2d(tgt, f, f)
//     loss_s = mean(|dx d'| * exp(-mean_c |dx I_s|)) + mean(|dy d'| * exp(-mean_c |dy I_s|))
//
// Launches (all scales in each; a block owns 1024 pixels, four per thread, lane = pixel): (1) pool the target image to every low scale and reduce the
// per-image disparity sums (block partials); (2) per pixel: the <= 4 differences it takes part in ->
// loss partials, unnormalised gradient g', partials of sum(g' * d); (3) only when normalising:
// g = g' / (m + eps) - sum(g' d) / (hw (m + eps)^2).  Every reduction is block partials summed in
// block order by the consumer: no atomics, bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int EG_THREADS = 256;
constexpr int EG_PX = 16;                      // pixels per thread: the per-block work (which scale / image am I, block sums)
constexpr int EG_TILE = EG_THREADS * EG_PX;    // is amortised over 4096 pixels (at 1024 the sum / fix-up launches were issue-bound on it)
constexpr int EV_OWN = 30 * 4;                 // edge_main_vec_kernel: columns a warp owns (lanes 1..30, four each)
#ifndef EV_WARPS_PER_SM
#define EV_WARPS_PER_SM 12                    // the row chunks shrink until the launch has this many warps per SM
#endif

// Launch-time constants, computed once on the host.
struct EdgeLayout {
    size_t pooled[PLB_MAX_SCALES];   // float [B,3,h,w] (scales with f > 1)
    size_t part_mean;                // double [blocks]
    size_t part_loss;                // double [blocks][2]   (loss partial, sum g' d partial)
    size_t img_inv;                  // float [PLB_MAX_SCALES][B]: 1 / (mean_hw(d) + 1e-7) of every image (normalize)
    size_t img_c;                    // float [PLB_MAX_SCALES][B]: sum(g' d) inv^2 / (h w) of every image (normalize, want_grad)
    // edge_main_vec_kernel: one warp per (scale, image, row chunk, 120-column strip); its partials replace the block partials
    int vec_main;                    // 1: every width is a multiple of 4 and every map 16-byte aligned
    int v_rows;                      // rows per chunk
    int v_first[PLB_MAX_SCALES + 1]; // first unit of every scale
    int v_strips[PLB_MAX_SCALES], v_per_img[PLB_MAX_SCALES];
    size_t total;
    int first_block[PLB_MAX_SCALES + 1];   // blocks of launch 1 / 2 / 3: EG_TILE pixels of one image of one scale
    int blocks_per_image[PLB_MAX_SCALES];
    int f[PLB_MAX_SCALES];                 // pooling factor H / h
    float cx[PLB_MAX_SCALES], cy[PLB_MAX_SCALES];   // weight_s / element count of the x / y difference maps
    int tile_pooled[PLB_MAX_SCALES];       // 1: factor 2, 4, 8, 16 or 32 - pooled by edge_pool_kernel (one pass over the image)
    int level_scale[6];                    // scale pooled at level l (factor 2^l), -1 = none
    int any_tile_pooled;
};

__host__ inline EdgeLayout edge_layout(const plb_edge_args& a) {
    EdgeLayout L;
    size_t off = 0;
    int nb = 0;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) {
        L.first_block[s] = nb;
        L.pooled[s] = off;
        L.blocks_per_image[s] = 0;
        L.f[s] = 1; L.cx[s] = L.cy[s] = 0.0f;
        if (s < a.n_scales && a.dh[s] > 0 && a.dw[s] > 0) {
            const int h = a.dh[s], w = a.dw[s];
            const size_t n = (size_t)h * w;
            if (h != a.H || w != a.W) off += ((size_t)a.B * 3 * n * sizeof(float) + 255) / 256 * 256;
            L.blocks_per_image[s] = (int)((n + EG_TILE - 1) / EG_TILE);
            nb += L.blocks_per_image[s] * a.B;
            L.f[s] = a.H / h > 0 ? a.H / h : 1;
            const float wscale = 1.0f / (float)L.f[s];                 // monodepth2: scale s weighs 1 / 2^s
            L.cx[s] = w > 1 ? wscale / ((float)a.B * (float)h * (float)(w - 1)) : 0.0f;   // mean over [B,1,h,w-1]
            L.cy[s] = h > 1 ? wscale / ((float)a.B * (float)(h - 1) * (float)w) : 0.0f;
        }
    }
    for (int l = 0; l < 6; ++l) L.level_scale[l] = -1;
    L.any_tile_pooled = 0;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) {
        L.tile_pooled[s] = 0;
        if (s >= a.n_scales) continue;
        for (int l = 1; l <= 5; ++l)
            if (L.f[s] == (1 << l) && L.level_scale[l] < 0 && a.dw[s] * L.f[s] <= a.W) {
                L.tile_pooled[s] = 1; L.level_scale[l] = s; L.any_tile_pooled = 1;
            }
    }
    L.first_block[PLB_MAX_SCALES] = nb;
    // vector main kernel: possible when rows are whole 128-bit packets
    L.vec_main = 1;
    for (int s = 0; s < a.n_scales && s < PLB_MAX_SCALES; ++s) {
        if (a.dw[s] % 4 != 0 || a.dw[s] < 4) L.vec_main = 0;
        if (((size_t)a.disp[s] | (size_t)a.g_disp[s] | (size_t)a.g_scratch[s]) & 15) L.vec_main = 0;
    }
    if (((size_t)a.tgt & 15) || (a.W & 3)) L.vec_main = 0;
    int nv = 0;
    L.v_rows = 16;
    for (int rows = 16; rows >= 4; rows >>= 1) {      // chunks shrink with the batch: ~12 warps per SM or more
        nv = 0;
        L.v_rows = rows;
        for (int s = 0; s < PLB_MAX_SCALES; ++s) {
            L.v_first[s] = nv;
            L.v_strips[s] = L.v_per_img[s] = 0;
            if (s < a.n_scales && a.dh[s] > 0 && a.dw[s] > 0) {
                L.v_strips[s] = (a.dw[s] + EV_OWN - 1) / EV_OWN;
                L.v_per_img[s] = L.v_strips[s] * ((a.dh[s] + rows - 1) / rows);
                nv += L.v_per_img[s] * a.B;
            }
        }
        L.v_first[PLB_MAX_SCALES] = nv;
        if (nv >= 148 * EV_WARPS_PER_SM) break;
    }
    const int n_part = L.vec_main && nv > nb ? nv : nb;          // partial records: blocks (scalar) or warps (vector)
    L.part_mean = off; off += ((size_t)nb * sizeof(double) + 255) / 256 * 256;
    L.part_loss = off; off += ((size_t)n_part * 2 * sizeof(double) + 255) / 256 * 256;
    L.img_inv = off; off += ((size_t)PLB_MAX_SCALES * a.B * sizeof(float) + 255) / 256 * 256;
    L.img_c = off; off += ((size_t)PLB_MAX_SCALES * a.B * sizeof(float) + 255) / 256 * 256;
    L.total = off;
    return L;
}

struct EdgeWork { int s, b, h, w, f, o0; };   // this thread's pixels: o0 + j * EG_THREADS, j < EG_PX (lane = pixel: coalesced)

__device__ __forceinline__ EdgeWork edge_work(const plb_edge_args& a, const EdgeLayout& L) {
    EdgeWork k;
    int s = 0;
    while (s + 1 < a.n_scales && (int)blockIdx.x >= L.first_block[s + 1]) ++s;
    const int local = blockIdx.x - L.first_block[s];
    const int b = local / L.blocks_per_image[s];
    k.s = s; k.b = b;
    k.h = a.dh[s]; k.w = a.dw[s]; k.f = L.f[s];
    k.o0 = (local - b * L.blocks_per_image[s]) * EG_TILE + threadIdx.x;
    return k;
}

// fixed-order block sum: warp butterflies, then the warp totals in warp order; every thread gets the result
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < EG_THREADS / 32; ++w) r += sh[w];
    __syncthreads();
    return r;
}

// f x f box sum at `src` (row pitch W): every load is issued before the first add (the loop is fully unrolled for the
// factors a depth pyramid has), vector loads when the row pitch allows
template <int F>
__device__ __forceinline__ float box_sum(const float* __restrict__ src, int W, bool vec) {
    float acc = 0.0f;
    if (vec && F == 2) {
        float2 v[2];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) v[dy] = __ldg(reinterpret_cast<const float2*>(src + (size_t)dy * W));
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) acc += v[dy].x + v[dy].y;
    } else if (vec && (F == 4 || F == 8)) {
        constexpr int Q = F >= 4 ? F / 4 : 1;
        float4 v[F][Q];
#pragma unroll
        for (int dy = 0; dy < F; ++dy)
#pragma unroll
            for (int q = 0; q < Q; ++q) v[dy][q] = __ldg(reinterpret_cast<const float4*>(src + (size_t)dy * W) + q);
#pragma unroll
        for (int dy = 0; dy < F; ++dy)
#pragma unroll
            for (int q = 0; q < Q; ++q) acc += (v[dy][q].x + v[dy][q].y) + (v[dy][q].z + v[dy][q].w);
    } else {
#pragma unroll
        for (int dy = 0; dy < F; ++dy)
#pragma unroll
            for (int dx = 0; dx < F; ++dx) acc += __ldg(src + (size_t)dy * W + dx);
    }
    return acc;
}
__device__ __forceinline__ float box_sum_any(const float* __restrict__ src, int W, int f, bool vec) {
    switch (f) {
        case 2: return box_sum<2>(src, W, vec);
        case 4: return box_sum<4>(src, W, vec);
        case 8: return box_sum<8>(src, W, vec);
        default: {
            float acc = 0.0f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) acc += __ldg(src + (size_t)dy * W + dx);
            return acc;
        }
    }
}

// Pooled target images of every power-of-two scale in ONE pass over the image: a block owns a 32 x 32 full-resolution
// tile, builds the 2 x 2 box-sum pyramid of its three channels in shared memory and writes level l to the scale whose
// factor is 2^l.  (Pooling each scale straight from the image read it once per scale, and the few blocks of the
// coarsest scale - 64 loads per output - were the tail of the whole launch.)
constexpr int EP_T = 32;
__global__ void __launch_bounds__(EG_THREADS)
edge_pool_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ float t0[3][EP_T][EP_T + 1];
    __shared__ float t1[3][EP_T / 2][EP_T / 2 + 1];
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * EP_T, y0 = blockIdx.y * EP_T;
    const size_t plane = (size_t)a.H * a.W;
    for (int q = tid; q < 3 * EP_T * EP_T; q += EG_THREADS) {
        const int c = q / (EP_T * EP_T), r = q - c * (EP_T * EP_T), ly = r / EP_T, lx = r - ly * EP_T;
        const int gx = x0 + lx, gy = y0 + ly;
        t0[c][ly][lx] = (gx < a.W && gy < a.H) ? __ldg(a.tgt + (size_t)(b * 3 + c) * plane + (size_t)gy * a.W + gx) : 0.0f;
    }
    __syncthreads();
    // level 1 from t0, then each level from the sums of the one below (ping-pong between t1 and t0's storage)
    float (*src)[EP_T + 1] = nullptr;
    int side = EP_T;
#pragma unroll 1
    for (int l = 1; l <= 5; ++l) {
        const int half = side / 2;
        const int sidx = L.level_scale[l];
        float vals[3];
        const bool mine = tid < half * half;
        const int oy = mine ? tid / half : 0, ox = mine ? tid - oy * half : 0;
        if (mine) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (l & 1) vals[c] = (t0[c][2 * oy][2 * ox] + t0[c][2 * oy][2 * ox + 1]) + (t0[c][2 * oy + 1][2 * ox] + t0[c][2 * oy + 1][2 * ox + 1]);
                else vals[c] = (t1[c][2 * oy][2 * ox] + t1[c][2 * oy][2 * ox + 1]) + (t1[c][2 * oy + 1][2 * ox] + t1[c][2 * oy + 1][2 * ox + 1]);
            }
        }
        __syncthreads();
        if (mine) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (l & 1) t1[c][oy][ox] = vals[c]; else t0[c][oy][ox] = vals[c];
            }
            if (sidx >= 0) {
                const int f = 1 << l, h = a.dh[sidx], w = a.dw[sidx];
                const int px = x0 / f + ox, py = y0 / f + oy;
                if (px < w && py < h) {
                    float* pooled = (float*)((char*)a.workspace + L.pooled[sidx]);
                    const float inv = 1.0f / (float)(f * f);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pooled[((size_t)(b * 3 + c) * h + py) * w + px] = vals[c] * inv;
                }
            }
        }
        __syncthreads();
        side = half;
    }
    (void)src;
}

// The same pyramid without shared memory or barriers, for images whose width is a multiple of 32 and height of 16 and
// factors up to 16 (the KITTI pyramids): a warp owns 32 x 16 pixels, every lane a 4 x 4 patch (four 128-bit loads per
// channel) - levels 1 and 2 are sums inside the lane, levels 3 and 4 two butterfly shuffles each (lane bits: 0 -> x + 4,
// 1 -> y + 4, 2 -> x + 8, 3 -> y + 8, 4 -> x + 16).  Same pairings as edge_pool_kernel - (a + b) + (c + d) of the
// 2 x 2 children at every level - so the pooled images are bitwise the same.  (The tiled kernel spent 60 instructions
// per input value and ten barriers per 12 KB tile: 74 us for the 94 MB of a 64-image batch.)
__global__ void __launch_bounds__(EG_THREADS)
edge_pool_fast_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.z;
    const int X0 = blockIdx.x * 64 + (warp & 1) * 32, Y0 = blockIdx.y * 64 + (warp >> 1) * 16;
    if (X0 >= a.W || Y0 >= a.H) return;
    const int x = X0 + 4 * ((lane & 1) + ((lane >> 2) & 1) * 2 + ((lane >> 4) & 1) * 4);
    const int y = Y0 + 4 * (((lane >> 1) & 1) + ((lane >> 3) & 1) * 2);
    float* out[5];
    int ow[5], oh[5];
#pragma unroll
    for (int l = 1; l <= 4; ++l) {
        const int sidx = L.level_scale[l];
        out[l] = sidx >= 0 ? (float*)((char*)a.workspace + L.pooled[sidx]) : nullptr;
        ow[l] = sidx >= 0 ? a.dw[sidx] : 0;
        oh[l] = sidx >= 0 ? a.dh[sidx] : 0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* src = a.tgt + ((size_t)(b * 3 + c) * a.H + y) * a.W + x;
        float4 r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)i * a.W));
        const float s00 = (r[0].x + r[0].y) + (r[1].x + r[1].y), s01 = (r[0].z + r[0].w) + (r[1].z + r[1].w);
        const float s10 = (r[2].x + r[2].y) + (r[3].x + r[3].y), s11 = (r[2].z + r[2].w) + (r[3].z + r[3].w);
        const float q = (s00 + s01) + (s10 + s11);
        float e = q + __shfl_xor_sync(0xffffffffu, q, 1);
        e = e + __shfl_xor_sync(0xffffffffu, e, 2);
        float h = e + __shfl_xor_sync(0xffffffffu, e, 4);
        h = h + __shfl_xor_sync(0xffffffffu, h, 8);
        if (out[1] != nullptr) {
            float* d = out[1] + ((size_t)(b * 3 + c) * oh[1] + (y >> 1)) * ow[1] + (x >> 1);
            *reinterpret_cast<float2*>(d) = make_float2(s00 * 0.25f, s01 * 0.25f);
            *reinterpret_cast<float2*>(d + ow[1]) = make_float2(s10 * 0.25f, s11 * 0.25f);
        }
        if (out[2] != nullptr) out[2][((size_t)(b * 3 + c) * oh[2] + (y >> 2)) * ow[2] + (x >> 2)] = q * (1.0f / 16.0f);
        if (out[3] != nullptr && (lane & 3) == 0) out[3][((size_t)(b * 3 + c) * oh[3] + (y >> 3)) * ow[3] + (x >> 3)] = e * (1.0f / 64.0f);
        if (out[4] != nullptr && (lane & 15) == 0) out[4][((size_t)(b * 3 + c) * oh[4] + (y >> 4)) * ow[4] + (x >> 4)] = h * (1.0f / 256.0f);
    }
}

// launch 1: pooled target images of the low scales + per-block disparity sums
__global__ void __launch_bounds__(EG_THREADS)
edge_prep_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ double sh[EG_THREADS / 32];
    const EdgeWork k = edge_work(a, L);
    const int n = k.h * k.w;
    const float* disp = a.disp[k.s] + (size_t)k.b * n;
    float dsum = 0.0f;
    const bool vec = (a.W & 3) == 0 && ((size_t)a.tgt & 15) == 0;
    float* pooled = (float*)((char*)a.workspace + L.pooled[k.s]);
    const float inv = 1.0f / (float)(k.f * k.f);
    if ((k.f == 1 || L.tile_pooled[k.s]) && (n & 3) == 0 && ((size_t)a.disp[k.s] & 15) == 0) {
        // nothing to pool here: the block only sums its 1024 disparities - one 128-bit load per thread
        const int o0 = (k.o0 - (int)threadIdx.x) + 4 * (int)threadIdx.x;
        float4 v[EG_PX / 4];
#pragma unroll
        for (int j = 0; j < EG_PX / 4; ++j) {
            const int o = o0 + j * (4 * EG_THREADS);
            v[j] = o < n ? __ldg(reinterpret_cast<const float4*>(disp + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < EG_PX / 4; ++j) dsum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        const double tot = block_sum((double)dsum, sh);
        if (threadIdx.x == 0) ((double*)((char*)a.workspace + L.part_mean))[blockIdx.x] = tot;
        return;
    }
#pragma unroll 4
    for (int j = 0; j < EG_PX; ++j) {
        const int o = k.o0 + j * EG_THREADS;
        if (o < n) {
            dsum += __ldg(disp + o);
            if (k.f > 1 && !L.tile_pooled[k.s]) {
                const int y = o / k.w, x = o - y * k.w;
                const float* src = a.tgt + ((size_t)(k.b * 3) * a.H + (size_t)y * k.f) * a.W + (size_t)x * k.f;
                const size_t plane = (size_t)a.H * a.W;
                const float p0 = box_sum_any(src, a.W, k.f, vec), p1 = box_sum_any(src + plane, a.W, k.f, vec),
                            p2 = box_sum_any(src + 2 * plane, a.W, k.f, vec);
                float* dst = pooled + (size_t)(k.b * 3) * n + o;
                dst[0] = p0 * inv; dst[n] = p1 * inv; dst[2 * (size_t)n] = p2 * inv;
            }
        }
    }
    const double tot = block_sum((double)dsum, sh);
    if (threadIdx.x == 0) ((double*)((char*)a.workspace + L.part_mean))[blockIdx.x] = tot;
}

// sum of the block partials of image b at scale s, by the whole block (thread t takes partials t, t + 256, ... in
// order, then the fixed-order block sum): every thread gets the result.  (One thread walking the partials alone
// cost 100+ us PER BLOCK - a dependent L2 load each.)
__device__ __forceinline__ double range_partial_sum(const double* parts, int first, int count, int stride, int offset, double* sh) {
    double m = 0.0;
    for (int q = threadIdx.x; q < count; q += EG_THREADS) m += __ldcg(parts + (size_t)(first + q) * stride + offset);
    return block_sum(m, sh);
}
__device__ __forceinline__ double image_partial_sum(const double* parts, const EdgeLayout& L, int s, int b, int stride,
                                                    int offset, double* sh) {
    return range_partial_sum(parts, L.first_block[s] + b * L.blocks_per_image[s], L.blocks_per_image[s], stride, offset, sh);
}

__device__ __forceinline__ float sgn1(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// Per-image scalars, ONE block per (scale, image) instead of every block of the image repeating the walk over its
// partials (at batch 64 that was two block-wide sums + up to 240 dependent L2 loads in each of 10 240 blocks of the
// main and the final launch): (a) after launch 1: inv = 1 / (mean + 1e-7); (b) after launch 2: the constant of the
// normalisation's gradient, and (block 0) the loss.  Same partials, same fixed order: bitwise the same values.
__global__ void __launch_bounds__(EG_THREADS)
edge_mean_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ double sh[EG_THREADS / 32];
    const int s = blockIdx.x / a.B, b = blockIdx.x - s * a.B;
    const int n = a.dh[s] * a.dw[s];
    const double m = image_partial_sum((const double*)((const char*)a.workspace + L.part_mean), L, s, b, 1, 0, sh) / (double)n;
    if (threadIdx.x == 0) ((float*)((char*)a.workspace + L.img_inv))[s * a.B + b] = 1.0f / ((float)m + 1e-7f);
}

__global__ void __launch_bounds__(EG_THREADS)
edge_gsum_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ double sh[EG_THREADS / 32];
    const double* parts = (const double*)((const char*)a.workspace + L.part_loss);
    if (blockIdx.x == 0) {
        double v = 0.0;
        const int n_part = L.vec_main ? L.v_first[PLB_MAX_SCALES] : L.first_block[PLB_MAX_SCALES];
        for (int q = threadIdx.x; q < n_part; q += EG_THREADS) v += __ldcg(parts + (size_t)q * 2);
        const double tot = block_sum(v, sh);
        if (threadIdx.x == 0 && a.loss != nullptr) *a.loss = (float)tot;
    }
    if (!(a.normalize && a.want_grad)) return;
    const int s = blockIdx.x / a.B, b = blockIdx.x - s * a.B;
    const int n = a.dh[s] * a.dw[s];
    const float s_inv = ((const float*)((const char*)a.workspace + L.img_inv))[s * a.B + b];
    const double gd = L.vec_main ? range_partial_sum(parts, L.v_first[s] + b * L.v_per_img[s], L.v_per_img[s], 2, 1, sh)
                                 : image_partial_sum(parts, L, s, b, 2, 1, sh);
    if (threadIdx.x == 0) ((float*)((char*)a.workspace + L.img_c))[s * a.B + b] = (float)(gd * (double)s_inv * (double)s_inv / (double)n);
}

// launch 2: loss partials and the unnormalised gradient
__global__ void __launch_bounds__(EG_THREADS)
edge_main_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    __shared__ double sh[EG_THREADS / 32];
    const EdgeWork k = edge_work(a, L);
    const int n = k.h * k.w, w = k.w, h = k.h;
    const float inv = a.normalize ? __ldg((const float*)((const char*)a.workspace + L.img_inv) + k.s * a.B + k.b) : 1.0f;
    const float* disp = a.disp[k.s] + (size_t)k.b * n;
    const float* img = (k.f > 1) ? (const float*)((const char*)a.workspace + L.pooled[k.s]) + (size_t)k.b * 3 * n
                                 : a.tgt + (size_t)k.b * 3 * n;
    const float cx = L.cx[k.s], cy = L.cy[k.s];
    const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
    const bool store = a.want_grad && a.g_disp[k.s] != nullptr;
    float lsum = 0.0f, gd = 0.0f;
    // edge weight of the difference between pixels p and q of the (pooled) image: exp(-mean_c |I_p - I_q|)
    auto edge_w = [&](int p, int q) -> float {
        const float gi = fabsf(__ldg(img + p) - __ldg(img + q)) + fabsf(__ldg(img + n + p) - __ldg(img + n + q)) +
                         fabsf(__ldg(img + 2 * n + p) - __ldg(img + 2 * n + q));
        return expf(-gi * (1.0f / 3.0f));
    };
#pragma unroll 4
    for (int j = 0; j < EG_PX; ++j) {
        const int o = k.o0 + j * EG_THREADS;
        if (o < n) {
            const int y = o / w, x = o - y * w;
            const float draw = __ldg(disp + o);
            const float d0 = draw * inv;
            float g = 0.0f;
            // the difference anchored on this pixel towards +x / +y (forward sum) and the one anchored on its
            // -x / -y neighbour (gradient only)
            if (x + 1 < w) {
                const float dd = d0 - __ldg(disp + o + 1) * inv;
                const float e = edge_w(o, o + 1) * cx;
                lsum = fmaf(fabsf(dd), e, lsum);
                g = fmaf(sgn1(dd), e, g);
            }
            if (x > 0) {
                const float dd = __ldg(disp + o - 1) * inv - d0;
                g = fmaf(-sgn1(dd), edge_w(o - 1, o) * cx, g);
            }
            if (y + 1 < h) {
                const float dd = d0 - __ldg(disp + o + w) * inv;
                const float e = edge_w(o, o + w) * cy;
                lsum = fmaf(fabsf(dd), e, lsum);
                g = fmaf(sgn1(dd), e, g);
            }
            if (y > 0) {
                const float dd = __ldg(disp + o - w) * inv - d0;
                g = fmaf(-sgn1(dd), edge_w(o - w, o) * cy, g);
            }
            g *= up;
            if (store) {
                const size_t go = (size_t)k.b * n + o;
                if (a.normalize) a.g_scratch[k.s][go] = g;    // g' (w.r.t. d'): fixed up by launch 3
                else a.g_disp[k.s][go] = a.accumulate ? a.g_disp[k.s][go] + g : g;
            }
            gd = fmaf(g, draw, gd);
        }
    }
    const double l = block_sum((double)lsum, sh);
    const double q = block_sum((double)gd, sh);
    if (threadIdx.x == 0) {
        double* parts = (double*)((char*)a.workspace + L.part_loss);
        parts[(size_t)blockIdx.x * 2] = l;
        parts[(size_t)blockIdx.x * 2 + 1] = q;
    }
}

// ---------------------------------------------------------------------------------------------
// Vector form of launch 2 (every width a multiple of 4, 16-byte aligned maps - the KITTI pyramids).  The kernel above
// evaluates each edge weight twice (once from either end: 24 image loads and four expf per pixel, 288 instructions per
// pixel, issue-bound).  Here a warp walks DOWN a strip of 120 columns (lanes 1..30 own four adjacent columns each,
// lanes 0 / 31 are the halo), rows y .. y + 3 in four register sets that rotate without moves: every lane
// evaluates the x- and y-difference anchored on each of its pixels ONCE - |dd| e into the forward sum, sgn(dd) e kept
// for the gradient of the pixel itself and (by one shuffle / one row of registers) of its +x / +y neighbour.  One
// 128-bit load of the disparity and three of the image per four pixels, one 128-bit store; per-warp partials in
// fixed order: bitwise repeatable.
// ---------------------------------------------------------------------------------------------
struct EdgeRow { float d[4], i0[4], i1[4], i2[4]; };

__device__ __forceinline__ float edge_sgn(float v) {
    return __int_as_float((__float_as_int(v) & 0x80000000) | (v != 0.0f ? 0x3f800000 : 0));
}

__global__ void __launch_bounds__(EG_THREADS)
edge_main_vec_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x * (EG_THREADS / 32) + warp;
    if (u >= L.v_first[PLB_MAX_SCALES]) return;
    int s = 0;
    while (s + 1 < a.n_scales && u >= L.v_first[s + 1]) ++s;
    const int local = u - L.v_first[s];
    const int b = local / L.v_per_img[s], rem = local - b * L.v_per_img[s];
    const int strips = L.v_strips[s], chunk = rem / strips, strip = rem - chunk * strips;
    const int h = a.dh[s], w = a.dw[s], n = h * w;
    const int x = strip * EV_OWN - 4 + 4 * lane;                  // first of this lane's four columns
    const bool colin = x >= 0 && x < w;                           // all four (w is a multiple of 4)
    const bool own_lane = lane >= 1 && lane <= 30 && colin;
    const int y0 = chunk * L.v_rows, y1 = min(y0 + L.v_rows, h);
    const int ystart = max(y0 - 1, 0), ylast = min(y1, h - 1);    // rows read: the one above (its y-term), the one below
    const float inv = a.normalize ? __ldg((const float*)((const char*)a.workspace + L.img_inv) + s * a.B + b) : 1.0f;
    const float* disp = a.disp[s] + (size_t)b * n + (colin ? x : 0);
    const float* img = ((L.f[s] > 1) ? (const float*)((const char*)a.workspace + L.pooled[s]) : a.tgt) + (size_t)b * 3 * n + (colin ? x : 0);
    const float cx = L.cx[s], cy = L.cy[s];
    const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
    float* gout = nullptr;
    if (a.want_grad && a.g_disp[s] != nullptr) gout = (a.normalize ? a.g_scratch[s] : a.g_disp[s]) + (size_t)b * n + (colin ? x : 0);
    const bool rmw = !a.normalize && a.accumulate != 0;
    bool xok[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xok[c] = colin && x + c + 1 < w;

    auto load = [&](EdgeRow& r, int y) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f), p0 = d, p1 = d, p2 = d;
        if (colin && y <= ylast) {
            const size_t o = (size_t)y * w;
            d = __ldg(reinterpret_cast<const float4*>(disp + o));
            p0 = __ldg(reinterpret_cast<const float4*>(img + o));
            p1 = __ldg(reinterpret_cast<const float4*>(img + n + o));
            p2 = __ldg(reinterpret_cast<const float4*>(img + 2 * (size_t)n + o));
        }
        r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.d[3] = d.w;
        r.i0[0] = p0.x; r.i0[1] = p0.y; r.i0[2] = p0.z; r.i0[3] = p0.w;
        r.i1[0] = p1.x; r.i1[1] = p1.y; r.i1[2] = p1.z; r.i1[3] = p1.w;
        r.i2[0] = p2.x; r.i2[1] = p2.y; r.i2[2] = p2.z; r.i2[3] = p2.w;
    };
    float ty_up[4] = {0.f, 0.f, 0.f, 0.f};                        // sgn(ddy) * ey of the row above
    float lsum = 0.0f, gd = 0.0f;

    auto step = [&](int y, const EdgeRow& cur, const EdgeRow& nxt) {
        // the +x neighbour of the lane's last pixel: the first pixel of the lane to the right
        const float dR = __shfl_down_sync(0xffffffffu, cur.d[0], 1);
        const float r0 = __shfl_down_sync(0xffffffffu, cur.i0[0], 1), r1 = __shfl_down_sync(0xffffffffu, cur.i1[0], 1),
                    r2 = __shfl_down_sync(0xffffffffu, cur.i2[0], 1);
        const bool yok = y + 1 < h;
        const bool owned = own_lane && y >= y0;
        float tx[4], ty[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float dn = c < 3 ? cur.d[c < 3 ? c + 1 : 3] : dR;
            const float n0 = c < 3 ? cur.i0[c < 3 ? c + 1 : 3] : r0, n1 = c < 3 ? cur.i1[c < 3 ? c + 1 : 3] : r1,
                        n2 = c < 3 ? cur.i2[c < 3 ? c + 1 : 3] : r2;
            const float gx = fabsf(cur.i0[c] - n0) + fabsf(cur.i1[c] - n1) + fabsf(cur.i2[c] - n2);
            const float gy = fabsf(cur.i0[c] - nxt.i0[c]) + fabsf(cur.i1[c] - nxt.i1[c]) + fabsf(cur.i2[c] - nxt.i2[c]);
            const float ex = xok[c] ? expf(-gx * (1.0f / 3.0f)) * cx : 0.0f;
            const float ey = (yok && colin) ? expf(-gy * (1.0f / 3.0f)) * cy : 0.0f;
            const float d0 = cur.d[c] * inv;
            const float ddx = d0 - dn * inv, ddy = d0 - nxt.d[c] * inv;
            tx[c] = edge_sgn(ddx) * ex;
            ty[c] = edge_sgn(ddy) * ey;
            if (owned) { lsum = fmaf(fabsf(ddx), ex, lsum); lsum = fmaf(fabsf(ddy), ey, lsum); }
        }
        const float txL = __shfl_up_sync(0xffffffffu, tx[3], 1);  // the x-term anchored on the pixel to the left
        if (owned && gout != nullptr) {
            float g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float left = c > 0 ? tx[c > 0 ? c - 1 : 0] : txL;
                g[c] = ((tx[c] - left) + (ty[c] - ty_up[c])) * up;
                gd = fmaf(g[c], cur.d[c], gd);
            }
            float4* q = reinterpret_cast<float4*>(gout + (size_t)y * w);
            if (rmw) { const float4 o = *q; g[0] += o.x; g[1] += o.y; g[2] += o.z; g[3] += o.w; }
            *q = make_float4(g[0], g[1], g[2], g[3]);
        } else if (owned) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float left = c > 0 ? tx[c > 0 ? c - 1 : 0] : txL;
                gd = fmaf(((tx[c] - left) + (ty[c] - ty_up[c])) * up, cur.d[c], gd);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) ty_up[c] = ty[c];
    };

    // four register sets: a row is loaded two full steps before its first use (with three, one step - ~1000 cycles -
    // was less than a DRAM round trip under load: the first use of every row held 30 % of the stall samples)
    EdgeRow A, Bq, C, Dq;
    load(A, ystart); load(Bq, ystart + 1); load(C, ystart + 2); load(Dq, ystart + 3);
    int y = ystart;
#pragma unroll 1
    while (true) {
        step(y, A, Bq); if (++y >= y1) break; load(A, y + 3);
        step(y, Bq, C); if (++y >= y1) break; load(Bq, y + 3);
        step(y, C, Dq); if (++y >= y1) break; load(C, y + 3);
        step(y, Dq, A); if (++y >= y1) break; load(Dq, y + 3);
    }
    // the warp's partials, lanes in a fixed (butterfly) order
    double l = (double)lsum, q = (double)gd;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
        double* parts = (double*)((char*)a.workspace + L.part_loss);
        parts[(size_t)u * 2] = l;
        parts[(size_t)u * 2 + 1] = q;
    }
}

// launch 3 (only when normalising with gradients): g = g' inv - sum(g' d) inv^2 / (h w)
__global__ void __launch_bounds__(EG_THREADS)
edge_final_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    if (skip_launch(a.skip_if_unit)) return;
    const EdgeWork k = edge_work(a, L);
    const int n = k.h * k.w;
    if (!(a.normalize && a.want_grad && a.g_disp[k.s] != nullptr)) return;
    const float s_inv = __ldg((const float*)((const char*)a.workspace + L.img_inv) + k.s * a.B + k.b);
    const float s_c = __ldg((const float*)((const char*)a.workspace + L.img_c) + k.s * a.B + k.b);
    if ((n & 3) == 0 && (((size_t)a.g_scratch[k.s] | (size_t)a.g_disp[k.s]) & 15) == 0) {
        // four consecutive pixels per thread: 128-bit loads and stores
        const int o0 = (k.o0 - (int)threadIdx.x) + 4 * (int)threadIdx.x;
#pragma unroll
        for (int j = 0; j < EG_PX / 4; ++j) {
            const int o = o0 + j * (4 * EG_THREADS);
            if (o < n) {
                const size_t go = (size_t)k.b * n + o;
                const float4 v = __ldcs(reinterpret_cast<const float4*>(a.g_scratch[k.s] + go));
                float4 g = make_float4(v.x * s_inv - s_c, v.y * s_inv - s_c, v.z * s_inv - s_c, v.w * s_inv - s_c);
                float4* out = reinterpret_cast<float4*>(a.g_disp[k.s] + go);
                if (a.accumulate) { const float4 w = *out; g.x += w.x; g.y += w.y; g.z += w.z; g.w += w.w; }
                *out = g;
            }
        }
        return;
    }
#pragma unroll 4
    for (int j = 0; j < EG_PX; ++j) {
        const int o = k.o0 + j * EG_THREADS;
        if (o < n) {
            const size_t go = (size_t)k.b * n + o;
            const float g = a.g_scratch[k.s][go] * s_inv - s_c;
            float* out = a.g_disp[k.s] + go;
            *out = a.accumulate ? *out + g : g;
        }
    }
}

static int validate_edge(const plb_edge_args* a) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 1 || a->W < 1 || a->n_scales < 1 || a->n_scales > PLB_MAX_SCALES) return PLB_EINVAL;
    if (a->tgt == nullptr || a->loss == nullptr) return PLB_ENULL;
    for (int s = 0; s < a->n_scales; ++s) {
        if (a->disp[s] == nullptr) return PLB_ENULL;
        if (a->dh[s] < 1 || a->dw[s] < 1) return PLB_EINVAL;
        // avg_pool2d(tgt, f, f) must give exactly the disparity's resolution
        const int f = a->H / a->dh[s];
        if (f < 1 || f * a->dh[s] != a->H || a->W / f != a->dw[s] || f > 64) return PLB_EINVAL;
        if ((long long)a->B * 3 * a->dh[s] * a->dw[s] >= (1LL << 31)) return PLB_EINVAL;
        if (a->want_grad && a->normalize && a->g_disp[s] != nullptr && a->g_scratch[s] == nullptr) return PLB_ENULL;
    }
    if (a->workspace == nullptr || a->workspace_bytes < edge_layout(*a).total) return PLB_EWORKSPACE;
    return PLB_OK;
}

int edge_launch(const plb_edge_args* a, cudaStream_t st) {
    const int rc = validate_edge(a);
    if (rc != PLB_OK) return rc;
    const EdgeLayout L = edge_layout(*a);
    const int nb = L.first_block[PLB_MAX_SCALES];
    if (L.any_tile_pooled) {
        if (a->B > 65535) return PLB_EINVAL;
        const bool fast = (a->W % 32) == 0 && (a->H % 16) == 0 && L.level_scale[5] < 0 && ((size_t)a->tgt & 15) == 0;
        if (fast) {
            edge_pool_fast_kernel<<<dim3((a->W + 63) / 64, (a->H + 63) / 64, a->B), EG_THREADS, 0, st>>>(*a, L);
        } else {
            dim3 pg((a->W + EP_T - 1) / EP_T, (a->H + EP_T - 1) / EP_T, a->B);
            edge_pool_kernel<<<pg, EG_THREADS, 0, st>>>(*a, L);
        }
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    edge_prep_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (a->normalize) {
        edge_mean_kernel<<<a->n_scales * a->B, EG_THREADS, 0, st>>>(*a, L);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    if (L.vec_main) {
        const int warps = L.v_first[PLB_MAX_SCALES];
        edge_main_vec_kernel<<<(warps + EG_THREADS / 32 - 1) / (EG_THREADS / 32), EG_THREADS, 0, st>>>(*a, L);
    } else {
        edge_main_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
    }
    ++g_launches;
    PLB_CHECK_LAUNCH();
    edge_gsum_kernel<<<(a->normalize && a->want_grad) ? a->n_scales * a->B : 1, EG_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (a->normalize && a->want_grad) {
        edge_final_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    return PLB_OK;
}

size_t edge_workspace_bytes(const plb_edge_args* a) { return edge_layout(*a).total; }

}  // namespace plb
