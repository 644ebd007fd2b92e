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
constexpr int EG_PX = 4;                       // consecutive pixels per thread: the per-block work (which scale / image am
constexpr int EG_TILE = EG_THREADS * EG_PX;    // I, image mean, block sums) is amortised over 1024 pixels

// Launch-time constants, computed once on the host.
struct EdgeLayout {
    size_t pooled[PLB_MAX_SCALES];   // float [B,3,h,w] (scales with f > 1)
    size_t part_mean;                // double [blocks]
    size_t part_loss;                // double [blocks][2]   (loss partial, sum g' d partial)
    size_t img_inv;                  // float [PLB_MAX_SCALES][B]: 1 / (mean_hw(d) + 1e-7) of every image (normalize)
    size_t img_c;                    // float [PLB_MAX_SCALES][B]: sum(g' d) inv^2 / (h w) of every image (normalize, want_grad)
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
    L.part_mean = off; off += ((size_t)nb * sizeof(double) + 255) / 256 * 256;
    L.part_loss = off; off += ((size_t)nb * 2 * sizeof(double) + 255) / 256 * 256;
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

// launch 1: pooled target images of the low scales + per-block disparity sums
__global__ void __launch_bounds__(EG_THREADS)
edge_prep_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    __shared__ double sh[EG_THREADS / 32];
    const EdgeWork k = edge_work(a, L);
    const int n = k.h * k.w;
    const float* disp = a.disp[k.s] + (size_t)k.b * n;
    float dsum = 0.0f;
    const bool vec = (a.W & 3) == 0 && ((size_t)a.tgt & 15) == 0;
    float* pooled = (float*)((char*)a.workspace + L.pooled[k.s]);
    const float inv = 1.0f / (float)(k.f * k.f);
#pragma unroll
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
__device__ __forceinline__ double image_partial_sum(const double* parts, const EdgeLayout& L, int s, int b, int stride,
                                                    int offset, double* sh) {
    double m = 0.0;
    const int first = L.first_block[s] + b * L.blocks_per_image[s];
    for (int q = threadIdx.x; q < L.blocks_per_image[s]; q += EG_THREADS) m += __ldcg(parts + (size_t)(first + q) * stride + offset);
    return block_sum(m, sh);
}

__device__ __forceinline__ float sgn1(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// Per-image scalars, ONE block per (scale, image) instead of every block of the image repeating the walk over its
// partials (at batch 64 that was two block-wide sums + up to 240 dependent L2 loads in each of 10 240 blocks of the
// main and the final launch): (a) after launch 1: inv = 1 / (mean + 1e-7); (b) after launch 2: the constant of the
// normalisation's gradient, and (block 0) the loss.  Same partials, same fixed order: bitwise the same values.
__global__ void __launch_bounds__(EG_THREADS)
edge_mean_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    __shared__ double sh[EG_THREADS / 32];
    const int s = blockIdx.x / a.B, b = blockIdx.x - s * a.B;
    const int n = a.dh[s] * a.dw[s];
    const double m = image_partial_sum((const double*)((const char*)a.workspace + L.part_mean), L, s, b, 1, 0, sh) / (double)n;
    if (threadIdx.x == 0) ((float*)((char*)a.workspace + L.img_inv))[s * a.B + b] = 1.0f / ((float)m + 1e-7f);
}

__global__ void __launch_bounds__(EG_THREADS)
edge_gsum_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    __shared__ double sh[EG_THREADS / 32];
    const double* parts = (const double*)((const char*)a.workspace + L.part_loss);
    if (blockIdx.x == 0) {
        double v = 0.0;
        for (int q = threadIdx.x; q < L.first_block[PLB_MAX_SCALES]; q += EG_THREADS) v += __ldcg(parts + (size_t)q * 2);
        const double tot = block_sum(v, sh);
        if (threadIdx.x == 0 && a.loss != nullptr) *a.loss = (float)tot;
    }
    if (!(a.normalize && a.want_grad)) return;
    const int s = blockIdx.x / a.B, b = blockIdx.x - s * a.B;
    const int n = a.dh[s] * a.dw[s];
    const float s_inv = ((const float*)((const char*)a.workspace + L.img_inv))[s * a.B + b];
    const double gd = image_partial_sum(parts, L, s, b, 2, 1, sh);
    if (threadIdx.x == 0) ((float*)((char*)a.workspace + L.img_c))[s * a.B + b] = (float)(gd * (double)s_inv * (double)s_inv / (double)n);
}

// launch 2: loss partials and the unnormalised gradient
__global__ void __launch_bounds__(EG_THREADS)
edge_main_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
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
#pragma unroll
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

// launch 3 (only when normalising with gradients): g = g' inv - sum(g' d) inv^2 / (h w)
__global__ void __launch_bounds__(EG_THREADS)
edge_final_kernel(const __grid_constant__ plb_edge_args a, const __grid_constant__ EdgeLayout L) {
    const EdgeWork k = edge_work(a, L);
    const int n = k.h * k.w;
    if (!(a.normalize && a.want_grad && a.g_disp[k.s] != nullptr)) return;
    const float s_inv = __ldg((const float*)((const char*)a.workspace + L.img_inv) + k.s * a.B + k.b);
    const float s_c = __ldg((const float*)((const char*)a.workspace + L.img_c) + k.s * a.B + k.b);
#pragma unroll
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
        dim3 pg((a->W + EP_T - 1) / EP_T, (a->H + EP_T - 1) / EP_T, a->B);
        if (pg.z > 65535) return PLB_EINVAL;
        edge_pool_kernel<<<pg, EG_THREADS, 0, st>>>(*a, L);
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
    edge_main_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
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
