// Stand-alone photometric maps of the reference's dormant path:
//   SSIM.standard_loss            (losses.py:12-54)   out = clamp((1 - SSIM3x3(x, y)) / 2, 0, 1)
//   Losses.compute_photometric_loss (losses.py:66-84) out = 0.85 * that + 0.15 * |y - x|, then
//                                                      clamp(max = mean + 0.5 * std)  (threshold detached)
// and their vjp with respect to both images.  The fused min-reprojection kernel (photo_min.cu) does
// not go through these; they exist so that the reference's individual functions drop in.
//
// Forward: one thread per (b, c, y, x); the 3x3 window is read with ReflectionPad2d(1) indices.
// The clip needs the mean and the unbiased std of the whole map: block partial sums (fp64) are
// combined in block order by the last block to finish, which writes the threshold to a device
// scalar; a second launch clamps in place.  No host synchronisation (the reference's float()
// forces one).
// Backward: tiled; the derivative coefficients of every window centre are evaluated once per tile (+ 1 halo)
// into shared memory, then every input pixel q gathers the <= 9 centres p that contain it (reflection makes
// q appear up to 4 times in a border window).  Gather form: no atomics, bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int PMAP_THREADS = 256;

__device__ __forceinline__ int reflect1(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

struct WinStats { float sx, sy, sxx, syy, sxy; };

__device__ __forceinline__ WinStats window(const float* __restrict__ x, const float* __restrict__ y, int px, int py,
                                           int H, int W) {
    WinStats s = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int ry = reflect1(py + dy, H) * W;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int o = ry + reflect1(px + dx, W);
            const float a = __ldg(x + o), b = __ldg(y + o);
            s.sx += a; s.sy += b;
            s.sxx = fmaf(a, a, s.sxx); s.syy = fmaf(b, b, s.syy); s.sxy = fmaf(a, b, s.sxy);
        }
    }
    return s;
}

// value of the map at one window centre and the coefficients of its derivative:
//   d out / d x_q = mult_q * (ax + bx * x_q + cx * y_q)  (+ lx at the centre)
//   d out / d y_q = mult_q * (ay + by * y_q + cy * x_q)  (+ ly at the centre)
struct MapTerm { float v, ax, bx, cx, lx, ay, by, cy, ly; };

__device__ __forceinline__ MapTerm map_term(const WinStats s, float xc, float yc, float C1, float C2, float w_ssim,
                                            float w_l1, bool grad) {
    MapTerm r;
    const float i9 = 1.0f / 9.0f;
    const float mux = s.sx * i9, muy = s.sy * i9;
    const float mxx = mux * mux, myy = muy * muy, mxy = mux * muy;
    const float sigx = s.sxx * i9 - mxx, sigy = s.syy * i9 - myy, sigxy = s.sxy * i9 - mxy;
    const float N1 = 2.0f * mxy + C1, N2 = 2.0f * sigxy + C2;
    const float D1 = mxx + myy + C1, D2 = sigx + sigy + C2;
    const float ssim = (N1 * N2) / (D1 * D2);
    const float h = (1.0f - ssim) * 0.5f;
    const float diff = yc - xc;          // torch.abs(target - pred)
    r.v = w_ssim * fminf(fmaxf(h, 0.0f), 1.0f) + w_l1 * fabsf(diff);
    if (grad) {
        // d ssim / d x_q = (2/9) { mu_y (N2 - N1) / (D1 D2) - ssim mu_x (1/D1 - 1/D2) } - x_q (2/9) ssim / D2
        //                  + y_q (2/9) N1 / (D1 D2); symmetric in (x, y).  d out = -w_ssim/2 * d ssim inside the clamp.
        const float k = (h >= 0.0f && h <= 1.0f) ? (-0.5f * w_ssim * 2.0f * i9) : 0.0f;
        const float iD1 = 1.0f / D1, iD2 = 1.0f / D2, iD12 = iD1 * iD2;
        r.ax = k * (muy * (N2 - N1) * iD12 - ssim * mux * (iD1 - iD2));
        r.ay = k * (mux * (N2 - N1) * iD12 - ssim * muy * (iD1 - iD2));
        r.bx = r.by = k * (-ssim * iD2);
        r.cx = r.cy = k * (N1 * iD12);
        const float sg = (diff > 0.0f ? 1.0f : 0.0f) - (diff < 0.0f ? 1.0f : 0.0f);
        r.lx = -w_l1 * sg;
        r.ly = w_l1 * sg;
    }
    return r;
}

struct PmapLayout { size_t ticket, partials, total; int blocks; };

static inline PmapLayout pmap_layout(const plb_photomap_args& a) {
    PmapLayout L;
    const long long n = (long long)a.B * a.C * a.H * a.W;
    L.blocks = (int)((n + PMAP_THREADS - 1) / PMAP_THREADS);
    L.ticket = 0;
    L.partials = 256;
    L.total = 256 + ((size_t)L.blocks * 2 * sizeof(double) + 255) / 256 * 256;
    return L;
}

__global__ void __launch_bounds__(PMAP_THREADS)
photomap_fwd_kernel(const __grid_constant__ plb_photomap_args a, const PmapLayout L) {
    const int H = a.H, W = a.W, plane = H * W;
    const long long n = (long long)a.B * a.C * plane;
    const long long idx = (long long)blockIdx.x * PMAP_THREADS + threadIdx.x;
    float v = 0.0f;
    if (idx < n) {
        const long long img = idx / plane;
        const int o = (int)(idx - img * plane);
        const int py = o / W, px = o - py * W;
        const float* x = a.x + img * plane;
        const float* y = a.y + img * plane;
        const float xc = __ldg(x + o), yc = __ldg(y + o);
        if (a.w_ssim != 0.0f) {
            v = map_term(window(x, y, px, py, H, W), xc, yc, a.C1, a.C2, a.w_ssim, a.w_l1, false).v;
        } else {
            v = a.w_l1 * fabsf(yc - xc);
        }
        a.out[idx] = v;
    }
    if (a.clip < 0.0f) return;
    // ---- mean / unbiased std of the whole map: block partials, combined in block order -------------
    __shared__ double s_a[PMAP_THREADS], s_b[PMAP_THREADS];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    s_a[tid] = (idx < n) ? (double)v : 0.0;
    s_b[tid] = (idx < n) ? (double)v * (double)v : 0.0;
    __syncthreads();
    for (int k = PMAP_THREADS / 2; k > 0; k >>= 1) {
        if (tid < k) { s_a[tid] += s_a[tid + k]; s_b[tid] += s_b[tid + k]; }
        __syncthreads();
    }
    int32_t* ticket = (int32_t*)((char*)a.workspace + L.ticket);
    double* partials = (double*)((char*)a.workspace + L.partials);
    if (tid == 0) {
        __stcg(partials + 2 * blockIdx.x, s_a[0]);
        __stcg(partials + 2 * blockIdx.x + 1, s_b[0]);
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += PMAP_THREADS) { sa += __ldcg(partials + 2 * k); sb += __ldcg(partials + 2 * k + 1); }
    s_a[tid] = sa; s_b[tid] = sb;
    __syncthreads();
    for (int k = PMAP_THREADS / 2; k > 0; k >>= 1) {
        if (tid < k) { s_a[tid] += s_a[tid + k]; s_b[tid] += s_b[tid + k]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double nn = (double)n;
        const double mean = s_a[0] / nn;
        const double var = nn > 1.0 ? fmax((s_b[0] - nn * mean * mean) / (nn - 1.0), 0.0) : 0.0;
        *a.threshold = (float)((float)mean + a.clip * (float)sqrt(var));
        *ticket = 0;
    }
}

__global__ void __launch_bounds__(PMAP_THREADS)
photomap_clip_kernel(float* out, long long n, const float* thr) {
    const long long idx = (long long)blockIdx.x * PMAP_THREADS + threadIdx.x;
    if (idx >= n) return;
    const float t = __ldg(thr);
    out[idx] = fminf(out[idx], t);
}

// Backward, tiled: one block = one 32 x 8 tile of one (image, channel) plane.  The derivative coefficients of
// every window centre of tile + 1 halo are evaluated ONCE (scaled by the upstream gradient, clip applied) into
// shared memory; then every tile pixel q gathers its <= 9 centres.  (One thread per pixel recomputing the
// statistics of all nine centres did the 18-load window nine times over: 430 us for a 12 x 3 x 192 x 640 map.)
constexpr int PB_TW = 32, PB_TH = 8;
constexpr int PB_W1 = PB_TW + 2, PB_H1 = PB_TH + 2, PB_N1 = PB_W1 * PB_H1;   // centres: tile + 1

__global__ void __launch_bounds__(PMAP_THREADS)
photomap_bwd_kernel(const __grid_constant__ plb_photomap_args a, int tiles_x) {
    const int H = a.H, W = a.W, plane = H * W;
    const int tile = blockIdx.x, tx0 = (tile % tiles_x) * PB_TW, ty0 = (tile / tiles_x) * PB_TH;
    const long long img = blockIdx.y;
    const float* x = a.x + img * plane;
    const float* y = a.y + img * plane;
    const float* g = a.g_out + img * plane;
    const int tid = threadIdx.x;
    const bool clip = a.clip >= 0.0f;
    const float thr = clip ? __ldg(a.threshold) : 0.0f;
    // per centre p: go * (ax, bx, cx) and go * (ay, by, cy); the centre-only terms go * lx, go * ly
    __shared__ float4 sX[PB_N1], sY[PB_N1];      // .x = a, .y = b, .z = c, .w = l
    for (int k = tid; k < PB_N1; k += PMAP_THREADS) {
        const int ly = k / PB_W1, lx = k - ly * PB_W1;
        const int px = tx0 + lx - 1, py = ty0 + ly - 1;
        float4 cx4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), cy4 = cx4;
        if (px >= 0 && px < W && py >= 0 && py < H) {
            const int po = py * W + px;
            const float go = __ldg(g + po);
            if (go != 0.0f) {
                const float xc = __ldg(x + po), yc = __ldg(y + po);
                MapTerm t;
                if (a.w_ssim != 0.0f) {
                    t = map_term(window(x, y, px, py, H, W), xc, yc, a.C1, a.C2, a.w_ssim, a.w_l1, true);
                } else {
                    const float diff = yc - xc;
                    const float sg = (diff > 0.0f ? 1.0f : 0.0f) - (diff < 0.0f ? 1.0f : 0.0f);
                    t.v = a.w_l1 * fabsf(diff);
                    t.ax = t.bx = t.cx = t.ay = t.by = t.cy = 0.0f;
                    t.lx = -a.w_l1 * sg; t.ly = a.w_l1 * sg;
                }
                if (!clip || t.v <= thr) {            // torch.clamp passes the gradient where v <= max
                    cx4 = make_float4(go * t.ax, go * t.bx, go * t.cx, go * t.lx);
                    cy4 = make_float4(go * t.ay, go * t.by, go * t.cy, go * t.ly);
                }
            }
        }
        sX[k] = cx4; sY[k] = cy4;
    }
    __syncthreads();
    const int qx = tx0 + (tid & 31), qy = ty0 + (tid >> 5);
    if (qx >= W || qy >= H) return;
    const int o = qy * W + qx;
    const float xq = __ldg(x + o), yq = __ldg(y + o);
    const int k0 = ((tid >> 5) + 1) * PB_W1 + (tid & 31) + 1;
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int py = qy + dy;
        // multiplicity of row qy in the reflection-padded window of centre row py
        const float my = 1.0f + ((py == 0 && dy == -1) ? 1.0f : 0.0f) + ((py == H - 1 && dy == 1) ? 1.0f : 0.0f);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int px = qx + dx;
            const float mx = 1.0f + ((px == 0 && dx == -1) ? 1.0f : 0.0f) + ((px == W - 1 && dx == 1) ? 1.0f : 0.0f);
            const float4 cx4 = sX[k0 + dy * PB_W1 + dx], cy4 = sY[k0 + dy * PB_W1 + dx];   // zero outside the image
            const float m = mx * my;
            float dxq = m * fmaf(cx4.y, xq, fmaf(cx4.z, yq, cx4.x));
            float dyq = m * fmaf(cy4.y, yq, fmaf(cy4.z, xq, cy4.x));
            if (dx == 0 && dy == 0) { dxq += cx4.w; dyq += cy4.w; }
            gx += dxq;
            gy += dyq;
        }
    }
    const long long idx = img * plane + o;
    if (a.g_x != nullptr) a.g_x[idx] = gx;
    if (a.g_y != nullptr) a.g_y[idx] = gy;
}

static int validate_pmap(const plb_photomap_args* a, bool bwd) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->C < 1 || a->H < 2 || a->W < 2) return PLB_EINVAL;   // ReflectionPad2d(1) needs >= 2
    if ((long long)a->H * a->W >= (1LL << 31)) return PLB_EINVAL;
    if (a->x == nullptr || a->y == nullptr) return PLB_ENULL;
    if (!bwd && a->out == nullptr) return PLB_ENULL;
    if (bwd && a->g_out == nullptr) return PLB_ENULL;
    if (a->clip >= 0.0f) {
        if (a->threshold == nullptr) return PLB_ENULL;
        if (!bwd && (a->workspace == nullptr || a->workspace_bytes < pmap_layout(*a).total)) return PLB_EWORKSPACE;
    }
    return PLB_OK;
}

int photomap_launch(const plb_photomap_args* a, cudaStream_t st) {
    const int rc = validate_pmap(a, false);
    if (rc != PLB_OK) return rc;
    const PmapLayout L = pmap_layout(*a);
    photomap_fwd_kernel<<<L.blocks, PMAP_THREADS, 0, st>>>(*a, L);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (a->clip >= 0.0f) {
        photomap_clip_kernel<<<L.blocks, PMAP_THREADS, 0, st>>>(a->out, (long long)a->B * a->C * a->H * a->W, a->threshold);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    return PLB_OK;
}

int photomap_bwd_launch(const plb_photomap_args* a, cudaStream_t st) {
    const int rc = validate_pmap(a, true);
    if (rc != PLB_OK) return rc;
    const int tiles_x = (a->W + PB_TW - 1) / PB_TW, tiles_y = (a->H + PB_TH - 1) / PB_TH;
    if ((long long)a->B * a->C > 65535) return PLB_EINVAL;
    dim3 grid(tiles_x * tiles_y, a->B * a->C);
    photomap_bwd_kernel<<<grid, PMAP_THREADS, 0, st>>>(*a, tiles_x);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

size_t photomap_workspace_bytes(const plb_photomap_args* a) { return pmap_layout(*a).total; }

}  // namespace plb
