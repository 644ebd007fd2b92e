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
// Launches (all scales in each): (1) pool the target image to every low scale and reduce the
// per-image disparity sums (block partials); (2) per pixel: the <= 4 differences it takes part in ->
// loss partials, unnormalised gradient g', partials of sum(g' * d); (3) only when normalising:
// g = g' / (m + eps) - sum(g' d) / (hw (m + eps)^2).  Every reduction is block partials summed in
// block order by the consumer: no atomics, bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int EG_THREADS = 256;

struct EdgeLayout {
    size_t pooled[PLB_MAX_SCALES];   // float [B,3,h,w] (scales with f > 1)
    size_t part_mean;                // double [blocks1]
    size_t part_loss;                // double [blocks2][2]   (loss partial, sum g' d partial)
    size_t total;
    int first_block[PLB_MAX_SCALES + 1];   // blocks of launch 1 / 2 / 3: one thread per low-res pixel
    int blocks_per_image[PLB_MAX_SCALES];
};

__host__ __device__ inline EdgeLayout edge_layout(const plb_edge_args& a) {
    EdgeLayout L;
    size_t off = 0;
    int nb = 0;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) {
        L.first_block[s] = nb;
        L.pooled[s] = off;
        L.blocks_per_image[s] = 0;
        if (s < a.n_scales) {
            const size_t n = (size_t)a.dh[s] * a.dw[s];
            if (a.dh[s] != a.H || a.dw[s] != a.W) off += ((size_t)a.B * 3 * n * sizeof(float) + 255) / 256 * 256;
            L.blocks_per_image[s] = (int)((n + EG_THREADS - 1) / EG_THREADS);
            nb += L.blocks_per_image[s] * a.B;
        }
    }
    L.first_block[PLB_MAX_SCALES] = nb;
    L.part_mean = off; off += ((size_t)nb * sizeof(double) + 255) / 256 * 256;
    L.part_loss = off; off += ((size_t)nb * 2 * sizeof(double) + 255) / 256 * 256;
    L.total = off;
    return L;
}

struct EdgeWork { int s, b, blk, h, w, f, o; bool in; };

__device__ __forceinline__ EdgeWork edge_work(const plb_edge_args& a, const EdgeLayout& L) {
    EdgeWork k;
    int s = 0;
    while (s + 1 < a.n_scales && (int)blockIdx.x >= L.first_block[s + 1]) ++s;
    const int local = blockIdx.x - L.first_block[s];
    k.s = s; k.b = local / L.blocks_per_image[s]; k.blk = local - k.b * L.blocks_per_image[s];
    k.h = a.dh[s]; k.w = a.dw[s]; k.f = a.H / k.h;
    k.o = k.blk * EG_THREADS + threadIdx.x;
    k.in = k.o < k.h * k.w;
    return k;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int k = EG_THREADS / 2; k > 0; k >>= 1) {
        if (tid < k) sh[tid] += sh[tid + k];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// launch 1: pooled target images of the low scales + per-block disparity sums
__global__ void __launch_bounds__(EG_THREADS)
edge_prep_kernel(const __grid_constant__ plb_edge_args a, const EdgeLayout L) {
    __shared__ double sh[EG_THREADS];
    const EdgeWork k = edge_work(a, L);
    float d = 0.0f;
    if (k.in) {
        d = __ldg(a.disp[k.s] + (size_t)k.b * k.h * k.w + k.o);
        if (k.f > 1) {
            const int y = k.o / k.w, x = k.o - y * k.w;
            float* pooled = (float*)((char*)a.workspace + L.pooled[k.s]);
            const float inv = 1.0f / (float)(k.f * k.f);
            for (int c = 0; c < 3; ++c) {
                const float* src = a.tgt + ((size_t)(k.b * 3 + c) * a.H + (size_t)y * k.f) * a.W + (size_t)x * k.f;
                float acc = 0.0f;
                for (int dy = 0; dy < k.f; ++dy)
                    for (int dx = 0; dx < k.f; ++dx) acc += __ldg(src + dy * a.W + dx);
                pooled[(size_t)(k.b * 3 + c) * k.h * k.w + k.o] = acc * inv;
            }
        }
    }
    const double tot = block_sum((double)d, sh);
    if (threadIdx.x == 0) ((double*)((char*)a.workspace + L.part_mean))[blockIdx.x] = tot;
}

__device__ __forceinline__ double image_partial_sum(const double* parts, const EdgeLayout& L, int s, int b, int stride,
                                                    int offset) {
    // block order: fixed
    double m = 0.0;
    const int first = L.first_block[s] + b * L.blocks_per_image[s];
    for (int q = 0; q < L.blocks_per_image[s]; ++q) m += parts[(size_t)(first + q) * stride + offset];
    return m;
}

__device__ __forceinline__ float sgn1(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// launch 2: loss partials and the unnormalised gradient
__global__ void __launch_bounds__(EG_THREADS)
edge_main_kernel(const __grid_constant__ plb_edge_args a, const EdgeLayout L) {
    __shared__ double sh[EG_THREADS];
    __shared__ float s_inv;
    const EdgeWork k = edge_work(a, L);
    const size_t n = (size_t)k.h * k.w;
    if (threadIdx.x == 0) {
        float inv = 1.0f;
        if (a.normalize) {
            const double m = image_partial_sum((const double*)((const char*)a.workspace + L.part_mean), L, k.s, k.b, 1, 0) / (double)n;
            inv = 1.0f / ((float)m + 1e-7f);
        }
        s_inv = inv;
    }
    __syncthreads();
    const float inv = s_inv;
    const float* disp = a.disp[k.s] + (size_t)k.b * n;
    const float* img = (k.f > 1) ? (const float*)((const char*)a.workspace + L.pooled[k.s]) + (size_t)k.b * 3 * n
                                 : a.tgt + (size_t)k.b * 3 * n;
    double lsum = 0.0, gd = 0.0;
    if (k.in) {
        const int y = k.o / k.w, x = k.o - y * k.w;
        const float wscale = 1.0f / (float)k.f;                        // monodepth2: scale s weighs 1 / 2^s
        const float cx = k.w > 1 ? wscale / ((float)a.B * (float)k.h * (float)(k.w - 1)) : 0.0f;   // mean over [B,1,h,w-1]
        const float cy = k.h > 1 ? wscale / ((float)a.B * (float)(k.h - 1) * (float)k.w) : 0.0f;
        const float d0 = __ldg(disp + k.o) * inv;
        float i0[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) i0[c] = __ldg(img + c * n + k.o);
        float g = 0.0f;
        // the difference anchored on this pixel towards +x / +y (forward sum) and the one anchored on
        // its -x / -y neighbour (gradient only)
        if (x + 1 < k.w) {
            const float dd = d0 - __ldg(disp + k.o + 1) * inv;
            float gi = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gi += fabsf(i0[c] - __ldg(img + c * n + k.o + 1));
            const float e = expf(-gi * (1.0f / 3.0f));
            lsum += (double)(fabsf(dd) * e * cx);
            g += sgn1(dd) * e * cx;
        }
        if (x > 0) {
            const float dd = __ldg(disp + k.o - 1) * inv - d0;
            float gi = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gi += fabsf(__ldg(img + c * n + k.o - 1) - i0[c]);
            g -= sgn1(dd) * expf(-gi * (1.0f / 3.0f)) * cx;
        }
        if (y + 1 < k.h) {
            const float dd = d0 - __ldg(disp + k.o + k.w) * inv;
            float gi = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gi += fabsf(i0[c] - __ldg(img + c * n + k.o + k.w));
            const float e = expf(-gi * (1.0f / 3.0f));
            lsum += (double)(fabsf(dd) * e * cy);
            g += sgn1(dd) * e * cy;
        }
        if (y > 0) {
            const float dd = __ldg(disp + k.o - k.w) * inv - d0;
            float gi = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gi += fabsf(__ldg(img + c * n + k.o - k.w) - i0[c]);
            g -= sgn1(dd) * expf(-gi * (1.0f / 3.0f)) * cy;
        }
        const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
        g *= up;
        if (a.want_grad && a.g_disp[k.s] != nullptr) {
            float* out = a.g_disp[k.s] + (size_t)k.b * n + k.o;
            if (a.normalize) a.g_scratch[k.s][(size_t)k.b * n + k.o] = g;    // g' (w.r.t. d'): fixed up by launch 3
            else *out = a.accumulate ? *out + g : g;
        }
        gd = (double)g * (double)__ldg(disp + k.o);
    }
    const double l = block_sum(lsum, sh);
    const double q = block_sum(gd, sh);
    if (threadIdx.x == 0) {
        double* parts = (double*)((char*)a.workspace + L.part_loss);
        parts[(size_t)blockIdx.x * 2] = l;
        parts[(size_t)blockIdx.x * 2 + 1] = q;
    }
}

// launch 3: loss scalar (block 0) and, when normalising, g = g' inv - sum(g' d) inv^2 / (h w)
__global__ void __launch_bounds__(EG_THREADS)
edge_final_kernel(const __grid_constant__ plb_edge_args a, const EdgeLayout L) {
    __shared__ double sh[EG_THREADS];
    __shared__ float s_inv, s_c;
    const EdgeWork k = edge_work(a, L);
    const size_t n = (size_t)k.h * k.w;
    const double* parts = (const double*)((const char*)a.workspace + L.part_loss);
    if (blockIdx.x == 0) {
        double v = 0.0;
        for (int q = threadIdx.x; q < L.first_block[PLB_MAX_SCALES]; q += EG_THREADS) v += parts[(size_t)q * 2];
        const double tot = block_sum(v, sh);
        if (threadIdx.x == 0 && a.loss != nullptr) *a.loss = (float)tot;
    }
    if (!(a.normalize && a.want_grad && a.g_disp[k.s] != nullptr)) return;
    if (threadIdx.x == 0) {
        const double m = image_partial_sum((const double*)((const char*)a.workspace + L.part_mean), L, k.s, k.b, 1, 0) / (double)n;
        const float inv = 1.0f / ((float)m + 1e-7f);
        const double gd = image_partial_sum(parts, L, k.s, k.b, 2, 1);
        s_inv = inv;
        s_c = (float)(gd * (double)inv * (double)inv / (double)n);
    }
    __syncthreads();
    if (k.in) {
        const size_t o = (size_t)k.b * n + k.o;
        const float g = a.g_scratch[k.s][o] * s_inv - s_c;
        float* out = a.g_disp[k.s] + o;
        *out = a.accumulate ? *out + g : g;
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
    edge_prep_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    edge_main_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    edge_final_kernel<<<nb, EG_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t edge_workspace_bytes(const plb_edge_args* a) { return edge_layout(*a).total; }

}  // namespace plb
