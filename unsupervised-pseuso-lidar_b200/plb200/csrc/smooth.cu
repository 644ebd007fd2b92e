// Second-order depth smoothness (losses.py:242-260), all scales of the pyramid,
// loss and gradient in one launch.  disp -> depth is folded in.
//
// Streaming stencil, no shared-memory tiles and no block barriers in the loop.  One warp owns a strip of 28
// columns (lanes 2..29; lanes 0, 1, 30, 31 are the halo) and a chunk of SW_ROWS rows, and walks DOWN it with the
// depth rows y, y+1, y+2 in registers (one new row per step, loaded two steps ahead).  Per step every lane
// evaluates the four second differences ANCHORED on its pixel of row y once (neighbours by shuffle): |.| goes to
// the forward sums of the owned anchors, the signs stay in registers.  The gradient of pixel (x, y) is the signed
// stencil over the anchors that contain it - columns x, x-1, x-2 of this row (shuffles) and rows y-1, y-2 (the
// signs kept from the previous two steps) - so it is complete the moment row y has been processed; a chunk starts
// two rows early to have those signs.  No atomics, fixed-order reductions: bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int SW_THREADS = 256, SW_WARPS = SW_THREADS / 32;
constexpr int SW_OWN = 28;                 // owned columns per warp (lanes 2..29)
#ifndef SW_AHEAD
#define SW_AHEAD 2                         // rows in flight beyond the three in use
#endif
#ifndef SW_ROWS
#define SW_ROWS 16                         // owned rows per warp
#endif

struct SmoothLaunch {
    int n_units, grid, rows;
    int first_unit[PLB_MAX_SCALES + 1];
    int strips[PLB_MAX_SCALES], chunks[PLB_MAX_SCALES];
    float c1[PLB_MAX_SCALES], c2[PLB_MAX_SCALES], c3[PLB_MAX_SCALES];   // weight_s / element count of each difference map
};

static SmoothLaunch smooth_plan(const plb_smooth_args& a, int own = SW_OWN, int rows = SW_ROWS) {
    SmoothLaunch L;
    L.rows = rows;
    int n = 0;
    float wscale = 1.0f;
    for (int s = 0; s < PLB_MAX_SCALES; ++s) {
        L.first_unit[s] = n;
        L.strips[s] = L.chunks[s] = 0;
        L.c1[s] = L.c2[s] = L.c3[s] = 0.0f;
        if (s < a.n_scales) {
            const int h = a.dh[s], w = a.dw[s];
            L.strips[s] = (w + own - 1) / own;
            L.chunks[s] = (h + rows - 1) / rows;
            n += L.strips[s] * L.chunks[s] * a.B;
            L.c1[s] = wscale / ((float)a.B * (float)h * (float)(w - 2));
            L.c2[s] = wscale / ((float)a.B * (float)(h - 1) * (float)(w - 1));
            L.c3[s] = wscale / ((float)a.B * (float)(h - 2) * (float)w);
            wscale /= a.scale_decay;
        }
    }
    L.first_unit[PLB_MAX_SCALES] = n;
    L.n_units = n;
    L.grid = (n + SW_WARPS - 1) / SW_WARPS;
    return L;
}

// workspace: int32 ticket | double partials[grid]
static size_t smooth_ws_bytes(const SmoothLaunch& L) { return 256 + ((size_t)L.grid * sizeof(double) + 255) / 256 * 256; }

// sign(v) with sign(0) = 0 (and sign(NaN) = +-1, as any non-zero): one compare, one select, one bit merge
__device__ __forceinline__ float sgnf(float v) {
    return __int_as_float((__float_as_int(v) & 0x80000000) | (v != 0.0f ? 0x3f800000 : 0));
}

// block partial (fixed order), then the last block sums all partials in double
__device__ __forceinline__ void smooth_finish(const plb_smooth_args& a, double total) {
    int32_t* ticket = (int32_t*)a.workspace;
    double* partials = (double*)((char*)a.workspace + 256);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_fin[SW_THREADS];
    __shared__ int s_flag;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) s_fin[warp] = total;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < SW_WARPS; ++k) t += s_fin[k];
        __stcg(partials + blockIdx.x, t);
        __threadfence();
        s_flag = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double acc = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += SW_THREADS) acc += __ldcg(partials + k);
    __syncthreads();
    s_fin[tid] = acc;
    __syncthreads();
    for (int st = SW_THREADS / 2; st > 0; st >>= 1) {
        if (tid < st) s_fin[tid] += s_fin[tid + st];
        __syncthreads();
    }
    if (tid == 0) {
        if (a.loss != nullptr) *a.loss = (float)s_fin[0];
        *ticket = 0;
    }
}

__global__ void __launch_bounds__(SW_THREADS)
smooth_kernel(const __grid_constant__ plb_smooth_args a, const __grid_constant__ SmoothLaunch L) {
    if (skip_launch(a.skip_if_unit)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    double total = 0.0;   // this lane's share of the (weight / count-scaled) forward sums
    const int u = blockIdx.x * SW_WARPS + warp;
    if (u < L.n_units) {
        int s = 0;
        while (s + 1 < a.n_scales && u >= L.first_unit[s + 1]) ++s;
        const int local = u - L.first_unit[s];
        const int strips = L.strips[s], per_img = strips * L.chunks[s];
        const int b = local / per_img, rem = local - b * per_img;
        const int chunk = rem / strips, strip = rem - chunk * strips;      // consecutive warps: adjacent strips of the same rows
        const int h = a.dh[s], w = a.dw[s];
        const int x = strip * SW_OWN - 2 + lane;
        const int y0 = chunk * SW_ROWS, y1 = min(y0 + SW_ROWS, h);
        const int ystart = max(y0 - 2, 0);
        const bool colin = x >= 0 && x < w;
        const int ylast = min(y1 + 1, h - 1);
        const bool own_lane = lane >= 2 && lane < 2 + SW_OWN && colin;
        const float* disp = a.disp[s] + (size_t)b * h * w + (colin ? x : 0);
        float* gout = (a.want_grad && a.g_disp[s] != nullptr) ? a.g_disp[s] + (size_t)b * h * w + (colin ? x : 0) : nullptr;
        const bool is_depth = a.input_is_depth == PLB_INPUT_DEPTH, is_logit = a.input_is_depth == PLB_INPUT_LOGIT;
        const float ha = a.head_alpha, hb = a.head_beta;
        const float da = a.disp_a, db = a.disp_b;
        const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
        const float c1 = L.c1[s], c2 = L.c2[s], c3 = L.c3[s];
        const float g1 = c1 * up, g2 = c2 * up, g3 = c3 * up;
        const bool rmw = own_lane && gout != nullptr && a.accumulate != 0;

        // register queue: rows y, y+1, y+2 in use, SW_AHEAD more in flight; the loop is unrolled by the queue length so
        // that the rotation costs nothing, and the row pointers advance by w per step (no per-step multiplications)
        constexpr int NQ = 3 + SW_AHEAD;
        float q[NQ], oq[NQ];
        const float* prow = disp + (size_t)ystart * w;       // next depth row to load
        int yload = ystart;
        auto loadrow = [&]() -> float {                       // rows the chunk needs: [ystart, y1 + 1] inside the image
            float v = 0.0f;
            if (yload <= ylast) {                             // warp-uniform
                if (colin) v = __ldg(prow);                   // RAW value: converted when it enters the window, not here -
                                                              // a use right behind the load would stall the warp on it
            }
            prow += w; ++yload;
            return v;
        };
        const float* pold = gout != nullptr ? gout + (size_t)ystart * w : nullptr;   // next gradient row to pre-load (accumulate mode)
        int yold = ystart;
        auto loadold = [&]() -> float {
            float v = 0.0f;
            if (yold >= y0 && yold < y1) {                    // warp-uniform
                if (rmw) v = __ldcg(pold);
            }
            pold += w; ++yold;
            return v;
        };
#pragma unroll
        for (int k = 0; k < NQ; ++k) { q[k] = loadrow(); oq[k] = loadold(); }
        float s3_m1 = 0.0f, s3_m2 = 0.0f, sm_m1 = 0.0f;      // signs of the anchors one / two rows up
        float sum1 = 0.0f, summ = 0.0f, sum3 = 0.0f;
        const bool x1ok = colin && x <= w - 3, xmok = colin && x <= w - 2;
        float* pout = gout != nullptr ? gout + (size_t)ystart * w : nullptr;
        auto conv = [&](float v, int y) -> float {            // disparity -> depth; 0 outside the image, as the loads give
            if (is_depth || !colin || y > ylast) return v;
            if (is_logit) v = head_disp(v, ha, hb);
            return rcp_nr(fmaf(da, v, db));
        };
        float r0 = conv(q[0], ystart), r1 = conv(q[1], ystart + 1);      // rows y, y + 1 of the window (converted)
        auto step = [&](int y, float raw2, float old) {
            const float r2 = conv(raw2, y + 2);
            const bool owned = own_lane && y >= y0;
            const float d01 = __shfl_down_sync(0xffffffffu, r0, 1), d02 = __shfl_down_sync(0xffffffffu, r0, 2);
            const float d11 = __shfl_down_sync(0xffffffffu, r1, 1);
            const float e0 = d01 - r0, e1 = r1 - r0;          // first differences anchored on (x, y)
            const float v1 = x1ok ? (d02 - d01) - e0 : 0.0f;
            const float v3 = (colin && y <= h - 3) ? (r2 - r1) - e1 : 0.0f;
            const bool mixed = xmok && y <= h - 2;
            const float vm1 = mixed ? (d11 - r1) - e0 : 0.0f;
            const float vm2 = mixed ? (d11 - d01) - e1 : 0.0f;
            const float s1 = sgnf(v1), s3 = sgnf(v3), sm = sgnf(vm1) + sgnf(vm2);
            if (owned) {
                sum1 += fabsf(v1); sum3 += fabsf(v3); summ += fabsf(vm1) + fabsf(vm2);
            }
            const float s1_l1 = __shfl_up_sync(0xffffffffu, s1, 1), s1_l2 = __shfl_up_sync(0xffffffffu, s1, 2);
            const float sm_l1 = __shfl_up_sync(0xffffffffu, sm, 1), smp_l1 = __shfl_up_sync(0xffffffffu, sm_m1, 1);
            if (owned && pout != nullptr) {
                const float t1 = s1 - 2.0f * s1_l1 + s1_l2;
                const float t3 = s3 - 2.0f * s3_m1 + s3_m2;
                const float tm = (sm - sm_l1) - (sm_m1 - smp_l1);
                float g = fmaf(g1, t1, fmaf(g3, t3, g2 * tm));
                if (!is_depth) g *= -da * r0 * r0;
                if (is_logit) g *= head_chain_from_depth(r0, da, db, ha, hb);
                *pout = old + g;
            }
            if (pout != nullptr) pout += w;
            s3_m2 = s3_m1; s3_m1 = s3; sm_m1 = sm;
            r0 = r1; r1 = r2;
        };
        int y = ystart;
#pragma unroll 1
        while (y < y1) {
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                if (y < y1) {                                 // warp-uniform
                    step(y, q[(k + 2) % NQ], oq[k]);
                    q[k] = loadrow(); oq[k] = loadold();      // slot k now holds row y + NQ
                    ++y;
                }
            }
        }
        total = (double)(sum1 * c1) + (double)(summ * c2) + (double)(sum3 * c3);
    }

    smooth_finish(a, total);
}

// ---------------------------------------------------------------------------------------------
// Vector variant (every scale's width a multiple of 4, 16-byte aligned maps - the KITTI pyramids): a lane owns FOUR
// adjacent columns, so a row costs one 128-bit load, one 128-bit read-modify-write of the gradient and 5 shuffles
// per four pixels (the scalar kernel: 7 shuffles and ~150 instructions per pixel).  Lanes 1..30 own 120 columns,
// lanes 0 / 31 are the halo (the two columns to the left whose signs the first owned pixels need, the two to the
// right whose depths the last ones need).  Same anchors, same signs, same fixed-order sums: bitwise repeatable;
// the row chunk (L.rows) shrinks with the batch so that a small batch still fills the SMs.
// ---------------------------------------------------------------------------------------------
constexpr int SV_OWN = 30 * 4;

__global__ void __launch_bounds__(SW_THREADS)
smooth_vec_kernel(const __grid_constant__ plb_smooth_args a, const __grid_constant__ SmoothLaunch L) {
    if (skip_launch(a.skip_if_unit)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double total = 0.0;
    const int u = blockIdx.x * SW_WARPS + warp;
    if (u < L.n_units) {
        int s = 0;
        while (s + 1 < a.n_scales && u >= L.first_unit[s + 1]) ++s;
        const int local = u - L.first_unit[s];
        const int strips = L.strips[s], per_img = strips * L.chunks[s];
        const int b = local / per_img, rem = local - b * per_img;
        const int chunk = rem / strips, strip = rem - chunk * strips;
        const int h = a.dh[s], w = a.dw[s], w4 = w >> 2;
        const int x = strip * SV_OWN - 4 + 4 * lane;              // first of this lane's four columns
        const int y0 = chunk * L.rows, y1 = min(y0 + L.rows, h);
        const int ystart = max(y0 - 2, 0);
        const bool colin = x >= 0 && x < w;                       // all four columns (w is a multiple of 4)
        const int ylast = min(y1 + 1, h - 1);
        const bool own_lane = lane >= 1 && lane <= 30 && colin;
        const size_t img = (size_t)b * h * w + (colin ? x : 0);
        const float4* prow = reinterpret_cast<const float4*>(a.disp[s] + img) + (size_t)ystart * w4;
        float* gbase = (a.want_grad && a.g_disp[s] != nullptr) ? a.g_disp[s] + img : nullptr;
        const bool is_depth = a.input_is_depth == PLB_INPUT_DEPTH, is_logit = a.input_is_depth == PLB_INPUT_LOGIT;
        const float ha = a.head_alpha, hb = a.head_beta, da = a.disp_a, db = a.disp_b;
        const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
        const float c1 = L.c1[s], c2 = L.c2[s], c3 = L.c3[s];
        const float g1 = c1 * up, g2 = c2 * up, g3 = c3 * up;
        const bool rmw = own_lane && gbase != nullptr && a.accumulate != 0;
        bool x1ok[4], xmok[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { x1ok[c] = colin && x + c <= w - 3; xmok[c] = colin && x + c <= w - 2; }

        // Loads are UNCONDITIONAL, from clamped rows (the value of a row beyond the chunk, or of a lane beyond the image,
        // is masked when it enters the window, steps later): behind `cond ? load : 0` the compiler loads into temporaries
        // and copies them into the queue slot at once - a move that waits for the load, one exposed DRAM latency per step.
        const float4* rowbase = reinterpret_cast<const float4*>(a.disp[s] + img);
        (void)prow;
        int yload = ystart;
        auto loadrow = [&]() -> float4 {                          // RAW rows [ystart, ylast]
            const int yy = min(yload, ylast);
            ++yload;
            return __ldg(rowbase + (size_t)yy * w4);
        };
        // old gradients (accumulate mode): rows [y0, y1); without a gradient map / without accumulation the disparity
        // rows stand in as a valid address and the value is dropped at its use
        const bool use_old = gbase != nullptr && a.accumulate != 0;
        const float4* oldbase = use_old ? reinterpret_cast<const float4*>(gbase) : rowbase;
        int yold = ystart;
        auto loadold = [&]() -> float4 {
            const int yy = min(max(yold, y0), y1 - 1);
            ++yold;
            return __ldcg(oldbase + (size_t)yy * w4);
        };
        (void)rmw;
        // a row in the window: its four converted depths and the first two of the lane to the right
        auto conv = [&](const float4 raw, int y, float (&r)[6]) {
            const bool live = colin && y <= ylast;               // (the load was clamped: mask here, where the data has long arrived)
            r[0] = live ? raw.x : 0.0f; r[1] = live ? raw.y : 0.0f; r[2] = live ? raw.z : 0.0f; r[3] = live ? raw.w : 0.0f;
            if (!(is_depth || !colin || y > ylast)) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v = r[c];
                    if (is_logit) v = head_disp(v, ha, hb);
                    r[c] = rcp_nr(fmaf(da, v, db));
                }
            }
            r[4] = __shfl_down_sync(0xffffffffu, r[0], 1);
            r[5] = __shfl_down_sync(0xffffffffu, r[1], 1);
        };
        float r0[6], r1[6];
        { const float4 t0 = loadrow(), t1 = loadrow(); conv(t0, ystart, r0); conv(t1, ystart + 1, r1); }
        constexpr int NQ = 3;
        float4 q[NQ], oq[NQ];                                     // raw rows y + 2 .. y + 4, old gradients of rows y .. y + 2
#pragma unroll
        for (int k = 0; k < NQ; ++k) { q[k] = loadrow(); oq[k] = loadold(); }
        float s3_m1[4], s3_m2[4], sm_m1[4], smL_m1 = 0.0f;        // signs of the anchors one / two rows up
#pragma unroll
        for (int c = 0; c < 4; ++c) s3_m1[c] = s3_m2[c] = sm_m1[c] = 0.0f;
        float sum1 = 0.0f, summ = 0.0f, sum3 = 0.0f;
        float4* pout = reinterpret_cast<float4*>(gbase) + (size_t)ystart * w4;

        auto step = [&](int y, const float4 raw2, const float4 old) {
            float r2[6];
            conv(raw2, y + 2, r2);
            const bool owned = own_lane && y >= y0;
            const bool y3ok = colin && y <= h - 3, ymok = y <= h - 2;
            float s1[4], s3[4], sm[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float e0 = r0[c + 1] - r0[c], e1 = r1[c] - r0[c];     // first differences anchored on (x + c, y)
                const float v1 = x1ok[c] ? (r0[c + 2] - r0[c + 1]) - e0 : 0.0f;
                const float v3 = y3ok ? (r2[c] - r1[c]) - e1 : 0.0f;
                const bool mixed = xmok[c] && ymok;
                const float vm1 = mixed ? (r1[c + 1] - r1[c]) - e0 : 0.0f;
                const float vm2 = mixed ? (r1[c + 1] - r0[c + 1]) - e1 : 0.0f;
                s1[c] = sgnf(v1); s3[c] = sgnf(v3); sm[c] = sgnf(vm1) + sgnf(vm2);
                if (owned) { sum1 += fabsf(v1); sum3 += fabsf(v3); summ += fabsf(vm1) + fabsf(vm2); }
            }
            const float s1_L3 = __shfl_up_sync(0xffffffffu, s1[3], 1), s1_L2 = __shfl_up_sync(0xffffffffu, s1[2], 1);
            const float sm_L3 = __shfl_up_sync(0xffffffffu, sm[3], 1);
            if (owned && gbase != nullptr) {
                float g[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float a1 = c >= 1 ? s1[c - 1] : s1_L3, a2 = c >= 2 ? s1[c - 2] : (c == 1 ? s1_L3 : s1_L2);
                    const float bl = c >= 1 ? sm[c - 1] : sm_L3, bu = c >= 1 ? sm_m1[c - 1] : smL_m1;
                    const float t1 = s1[c] - 2.0f * a1 + a2;
                    const float t3 = s3[c] - 2.0f * s3_m1[c] + s3_m2[c];
                    const float tm = (sm[c] - bl) - (sm_m1[c] - bu);
                    float gg = fmaf(g1, t1, fmaf(g3, t3, g2 * tm));
                    if (!is_depth) gg *= -da * r0[c] * r0[c];
                    if (is_logit) gg *= head_chain_from_depth(r0[c], da, db, ha, hb);
                    g[c] = gg;
                }
                *pout = use_old ? make_float4(old.x + g[0], old.y + g[1], old.z + g[2], old.w + g[3]) : make_float4(g[0], g[1], g[2], g[3]);
            }
            pout += w4;
            smL_m1 = sm_L3;
#pragma unroll
            for (int c = 0; c < 4; ++c) { s3_m2[c] = s3_m1[c]; s3_m1[c] = s3[c]; sm_m1[c] = sm[c]; }
#pragma unroll
            for (int c = 0; c < 6; ++c) { r0[c] = r1[c]; r1[c] = r2[c]; }
        };
        int y = ystart;
#pragma unroll 1
        while (true) {                                            // three steps per trip: the queue rotates without moves
            step(y, q[0], oq[0]); if (++y >= y1) break; q[0] = loadrow(); oq[0] = loadold();
            step(y, q[1], oq[1]); if (++y >= y1) break; q[1] = loadrow(); oq[1] = loadold();
            step(y, q[2], oq[2]); if (++y >= y1) break; q[2] = loadrow(); oq[2] = loadold();
        }
        total = (double)(sum1 * c1) + (double)(summ * c2) + (double)(sum3 * c3);
    }
    smooth_finish(a, total);
}


// the vector variant needs 16-byte rows: every width a multiple of 4 and aligned maps
static bool smooth_vec_ok(const plb_smooth_args& a) {
    for (int s = 0; s < a.n_scales && s < PLB_MAX_SCALES; ++s) {
        if (a.dw[s] % 4 != 0 || a.dw[s] < 4) return false;
        if (((uintptr_t)a.disp[s] & 15) || ((uintptr_t)a.g_disp[s] & 15)) return false;
    }
    return true;
}
// rows per chunk: 16 when that already gives every SM ~12 warps (16 fit), else 8 / 4 (a chunk re-evaluates 2 rows of anchors)
static int smooth_vec_rows(const plb_smooth_args& a) {
    for (int rows = 16; rows > 4; rows >>= 1)
        if (smooth_plan(a, SV_OWN, rows).n_units >= 148 * 12) return rows;
    return 4;
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
    const bool vec = smooth_vec_ok(*a);
    const SmoothLaunch L = vec ? smooth_plan(*a, SV_OWN, smooth_vec_rows(*a)) : smooth_plan(*a);
    if (a->workspace == nullptr || a->workspace_bytes < smooth_ws_bytes(L)) return PLB_EWORKSPACE;
    if (vec) smooth_vec_kernel<<<L.grid, SW_THREADS, 0, st>>>(*a, L);
    else smooth_kernel<<<L.grid, SW_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t smooth_workspace_bytes(const plb_smooth_args* a) {
    return smooth_vec_ok(*a) ? smooth_ws_bytes(smooth_plan(*a, SV_OWN, smooth_vec_rows(*a))) : smooth_ws_bytes(smooth_plan(*a));
}

}  // namespace plb
