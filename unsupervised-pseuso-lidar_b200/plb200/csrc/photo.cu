// Fused photometric reprojection loss (live mode: warp + L1 mean), forward and
// gradients in one pass.  See DESIGN.md "photo_l1_kernel".
//
// One thread block = one 64x16 tile of one target image of one job (direction).
// Each warp covers 32 consecutive columns x 4 rows, so every global access of
// the warp - target pixels, disparity, the 4x3 bilinear taps of every source -
// is a (nearly) contiguous 128-byte row segment of a planar NCHW tensor.
// Nothing is staged or saved: K^-1, the pose matrices and P = K.[R|t] are
// rebuilt in the block prologue; loss and pose-gradient partials are reduced
// warp -> block -> (job, image) -> launch in a fixed order (bitwise repeatable)
// by "last block done" epilogues, so the whole op is ONE launch.
#include "common.cuh"

namespace plb {

constexpr int PH_TILE_W = 64;
constexpr int PH_TILE_H = 16;
constexpr int PH_THREADS = 256;
constexpr int PH_ROWS = 4;        // rows per thread
constexpr int PH_NACC = 13;       // [0] = sum |diff|, [1..12] = dP (3x4)
constexpr int PH_SLOTS = PLB_MAX_SCALES * PLB_MAX_SRC;
constexpr int PH_PSTRIDE = PH_SLOTS * PH_NACC;  // floats per block partial

struct PhotoLayout {
    size_t tickets;   // int32 [n_jobs*B + 1]
    size_t partials;  // float [n_jobs][B][tiles][PH_PSTRIDE]
    size_t ws_pose;   // float [n_jobs][B][MAX_SRC][6]
    size_t ws_loss;   // float [n_jobs][B][PH_SLOTS]
    size_t gup;       // float [n_jobs][MAX_SCALES][B*H*W]  (only when a low scale exists)
    size_t total;
    int tiles_x, tiles_y, tiles;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline bool photo_has_lowres(const plb_photo_args& a) {
    for (int j = 0; j < a.n_jobs; ++j)
        for (int s = 0; s < a.jobs[j].n_scales; ++s)
            if (a.jobs[j].dh[s] != a.H || a.jobs[j].dw[s] != a.W) return true;
    return false;
}

__host__ __device__ inline PhotoLayout photo_layout(const plb_photo_args& a) {
    PhotoLayout L;
    L.tiles_x = (a.W + PH_TILE_W - 1) / PH_TILE_W;
    L.tiles_y = (a.H + PH_TILE_H - 1) / PH_TILE_H;
    L.tiles = L.tiles_x * L.tiles_y;
    size_t off = 0;
    L.tickets = off; off = align_up(off + sizeof(int32_t) * ((size_t)a.n_jobs * a.B + 1), 256);
    L.partials = off; off = align_up(off + sizeof(float) * (size_t)a.n_jobs * a.B * L.tiles * PH_PSTRIDE, 256);
    L.ws_pose = off; off = align_up(off + sizeof(float) * (size_t)a.n_jobs * a.B * PLB_MAX_SRC * 6, 256);
    L.ws_loss = off; off = align_up(off + sizeof(float) * (size_t)a.n_jobs * a.B * PH_SLOTS, 256);
    L.gup = off;
    if (a.want_grad && photo_has_lowres(a))
        off = align_up(off + sizeof(float) * (size_t)a.n_jobs * PLB_MAX_SCALES * a.B * a.H * a.W, 256);
    L.total = off;
    return L;
}

template <bool GRAD, bool IMG_GRAD>
__device__ __forceinline__ void photo_pixel(const float* __restrict__ src, float* __restrict__ g_src,
                                            int H, int W, const float* __restrict__ P, float rx, float ry,
                                            float rz, float D, const float (&t)[3], float w_e, bool valid,
                                            float (&acc)[16], float& gD, float (&e_out)[3]) {
    const size_t plane = (size_t)H * W;
    float X = rx * D, Y = ry * D, Z = rz * D;
    float cx, cy, ze, ix, iy;
    project_pixel(P, X, Y, Z, (float)(W - 1), (float)(H - 1), cx, cy, ze, ix, iy);
    Taps tp;
    make_taps(ix, iy, W, H, tp);
    const int xa = max(tp.x0, 0), xb = min(tp.x0 + 1, W - 1);
    const int ya = max(tp.y0, 0), yb = min(tp.y0 + 1, H - 1);
    const bool mnw = valid && tp.vx0 && tp.vy0, mne = valid && tp.vx1 && tp.vy0;
    const bool msw = valid && tp.vx0 && tp.vy1, mse = valid && tp.vx1 && tp.vy1;
    const float* r0 = src + (size_t)ya * W;
    const float* r1 = src + (size_t)yb * W;
    float vnw[3], vne[3], vsw[3], vse[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        vnw[c] = ldg_pred(r0 + c * plane + xa, mnw);
        vne[c] = ldg_pred(r0 + c * plane + xb, mne);
        vsw[c] = ldg_pred(r1 + c * plane + xa, msw);
        vse[c] = ldg_pred(r1 + c * plane + xb, mse);
    }
    const float wnw = tp.wx0 * tp.wy0, wne = tp.wx1 * tp.wy0, wsw = tp.wx0 * tp.wy1, wse = tp.wx1 * tp.wy1;
    float Gx = 0.0f, Gy = 0.0f, l1 = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float proj = vnw[c] * wnw + vne[c] * wne + vsw[c] * wsw + vse[c] * wse;
        float d = proj - t[c];
        l1 += fabsf(d);
        if (GRAD) {
            float e = d > 0.0f ? w_e : (d < 0.0f ? -w_e : 0.0f);
            e_out[c] = e;
            Gx += e * ((vne[c] - vnw[c]) * tp.wy0 + (vse[c] - vsw[c]) * tp.wy1);
            Gy += e * ((vsw[c] - vnw[c]) * tp.wx0 + (vse[c] - vne[c]) * tp.wx1);
        }
    }
    if (valid) acc[0] += l1;
    if (GRAD) {
        const float iz = 1.0f / ze;
        const float px = cx * iz, py = cy * iz;
        float gcx = Gx * iz, gcy = Gy * iz, gcz = -(Gx * px + Gy * py) * iz;
        if (!(valid && tp.any)) { gcx = 0.0f; gcy = 0.0f; gcz = 0.0f; }
        gD += gcx * (P[0] * rx + P[1] * ry + P[2] * rz) + gcy * (P[4] * rx + P[5] * ry + P[6] * rz) +
              gcz * (P[8] * rx + P[9] * ry + P[10] * rz);
        acc[1] += gcx * X; acc[2] += gcx * Y; acc[3] += gcx * Z; acc[4] += gcx;
        acc[5] += gcy * X; acc[6] += gcy * Y; acc[7] += gcy * Z; acc[8] += gcy;
        acc[9] += gcz * X; acc[10] += gcz * Y; acc[11] += gcz * Z; acc[12] += gcz;
        if (IMG_GRAD && g_src != nullptr) {
            float* q0 = g_src + (size_t)ya * W;
            float* q1 = g_src + (size_t)yb * W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (mnw) atomicAdd(q0 + c * plane + xa, wnw * e_out[c]);
                if (mne) atomicAdd(q0 + c * plane + xb, wne * e_out[c]);
                if (msw) atomicAdd(q1 + c * plane + xa, wsw * e_out[c]);
                if (mse) atomicAdd(q1 + c * plane + xb, wse * e_out[c]);
            }
        }
    }
}

template <bool GRAD, bool IMG_GRAD>
__global__ void __launch_bounds__(PH_THREADS)
photo_l1_kernel(const __grid_constant__ plb_photo_args a) {
    if (skip_launch(a.skip_if_unit, a.skip_n)) return;

    const PhotoLayout L = photo_layout(a);
    char* ws = (char*)a.workspace;
    int32_t* tickets = (int32_t*)(ws + L.tickets);
    float* partials = (float*)(ws + L.partials);
    float* ws_pose = (float*)(ws + L.ws_pose);
    float* ws_loss = (float*)(ws + L.ws_loss);
    float* gup = (float*)(ws + L.gup);

    const int jb = blockIdx.z, b = blockIdx.y, tile = blockIdx.x;
    const plb_photo_job& job = a.jobs[jb];
    const int H = a.H, W = a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)H * W;

    __shared__ float s_kinv[9];
    __shared__ float s_P[PLB_MAX_SRC][12];
    __shared__ float s_acc[PH_THREADS / 32][PH_SLOTS][PH_NACC];
    __shared__ float s_red[PH_PSTRIDE];
    __shared__ int s_flag;

    if (tid == 0) kinv_f32((const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4), a.k_is_f64, s_kinv);
    if (tid >= 32 && tid < 32 + job.n_src) {
        const int i = tid - 32;
        float M[12];
        pose_to_M(a.poses + ((size_t)b * a.n_pose + job.pose_index[i]) * 6, a.rotation_mode, job.pose_inv[i], M);
        k_times_M((const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4), a.k_is_f64, M, s_P[i]);
    }
    for (int k = tid; k < (PH_THREADS / 32) * PH_SLOTS * PH_NACC; k += PH_THREADS) (&s_acc[0][0][0])[k] = 0.0f;
    __syncthreads();

    const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
    const float w_e = job.term_weight * up / (3.0f * (float)a.B * (float)H * (float)W);

    const int tx = tile % L.tiles_x, ty = tile / L.tiles_x;
    const int x = tx * PH_TILE_W + (warp & 1) * 32 + lane;
    const int ybase = ty * PH_TILE_H + (warp >> 1) * PH_ROWS;
    const bool xin = x < W;

    float t[PH_ROWS][3], rx[PH_ROWS], ry[PH_ROWS], rz[PH_ROWS];
    bool valid[PH_ROWS];
    const float* tgt_b = job.tgt + (size_t)b * 3 * plane;
#pragma unroll
    for (int j = 0; j < PH_ROWS; ++j) {
        const int y = ybase + j;
        valid[j] = xin && y < H;
        const size_t o = (size_t)min(y, H - 1) * W + min(x, W - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) t[j][c] = __ldg(tgt_b + c * plane + o);
        const float xf = (float)x, yf = (float)y;
        rx[j] = fmaf(s_kinv[1], yf, s_kinv[0] * xf) + s_kinv[2];
        ry[j] = fmaf(s_kinv[4], yf, s_kinv[3] * xf) + s_kinv[5];
        rz[j] = fmaf(s_kinv[7], yf, s_kinv[6] * xf) + s_kinv[8];
    }
    float gt[PH_ROWS][3];
    if (GRAD && IMG_GRAD) {
#pragma unroll
        for (int j = 0; j < PH_ROWS; ++j) gt[j][0] = gt[j][1] = gt[j][2] = 0.0f;
    }

#pragma unroll 1
    for (int s = 0; s < job.n_scales; ++s) {
        const int dh = job.dh[s], dw = job.dw[s];
        const bool full = (dh == H && dw == W);
        const float* disp_b = job.disp[s] + (size_t)b * dh * dw;
        float D[PH_ROWS], gD[PH_ROWS];
        if (full) {
#pragma unroll
            for (int j = 0; j < PH_ROWS; ++j) {
                const size_t o = (size_t)min(ybase + j, H - 1) * W + min(x, W - 1);
                float d = __ldg(disp_b + o);
                D[j] = a.input_is_depth ? d : 1.0f / (a.disp_a * d + a.disp_b);
                gD[j] = 0.0f;
            }
        } else {
            int x0, x1; float lx0, lx1;
            up_coord(min(x, W - 1), (float)dw / (float)W, dw, x0, x1, lx0, lx1);
#pragma unroll
            for (int j = 0; j < PH_ROWS; ++j) {
                int y0, y1; float ly0, ly1;
                up_coord(min(ybase + j, H - 1), (float)dh / (float)H, dh, y0, y1, ly0, ly1);
                float v00 = __ldg(disp_b + (size_t)y0 * dw + x0), v01 = __ldg(disp_b + (size_t)y0 * dw + x1);
                float v10 = __ldg(disp_b + (size_t)y1 * dw + x0), v11 = __ldg(disp_b + (size_t)y1 * dw + x1);
                if (!a.input_is_depth) {
                    v00 = 1.0f / (a.disp_a * v00 + a.disp_b); v01 = 1.0f / (a.disp_a * v01 + a.disp_b);
                    v10 = 1.0f / (a.disp_a * v10 + a.disp_b); v11 = 1.0f / (a.disp_a * v11 + a.disp_b);
                }
                D[j] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
                gD[j] = 0.0f;
            }
        }
#pragma unroll 1
        for (int i = 0; i < job.n_src; ++i) {
            float acc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
            const float* src_b = job.src[i] + (size_t)b * 3 * plane;
            float* g_src_b = (GRAD && IMG_GRAD && job.g_src[i]) ? job.g_src[i] + (size_t)b * 3 * plane : nullptr;
#pragma unroll
            for (int j = 0; j < PH_ROWS; ++j) {
                float e[3] = {0.0f, 0.0f, 0.0f};
                photo_pixel<GRAD, IMG_GRAD>(src_b, g_src_b, H, W, s_P[i], rx[j], ry[j], rz[j], D[j], t[j], w_e,
                                            valid[j], acc, gD[j], e);
                if (GRAD && IMG_GRAD) { gt[j][0] -= e[0]; gt[j][1] -= e[1]; gt[j][2] -= e[2]; }
            }
            if (GRAD) {
                int which;
                float r = warp_reduce16(acc, lane, which);
                if ((lane & 1) == 0 && which < PH_NACC) s_acc[warp][s * PLB_MAX_SRC + i][which] = r;
            } else {
                float r = warp_sum(acc[0]);
                if (lane == 0) s_acc[warp][s * PLB_MAX_SRC + i][0] = r;
            }
        }
        if (GRAD) {
            if (full) {
                if (job.g_disp[s] != nullptr) {
                    float* g = job.g_disp[s] + (size_t)b * plane;
#pragma unroll
                    for (int j = 0; j < PH_ROWS; ++j)
                        if (valid[j]) {
                            const float chain = a.input_is_depth ? 1.0f : -a.disp_a * D[j] * D[j];
                            g[(size_t)(ybase + j) * W + x] = gD[j] * chain;
                        }
                }
            } else if (job.g_disp[s] != nullptr) {
                float* g = gup + ((size_t)(jb * PLB_MAX_SCALES + s) * a.B + b) * plane;
#pragma unroll
                for (int j = 0; j < PH_ROWS; ++j)
                    if (valid[j]) g[(size_t)(ybase + j) * W + x] = gD[j];
            }
        }
    }
    if (GRAD && IMG_GRAD && job.g_tgt != nullptr) {
        float* g = job.g_tgt + (size_t)b * 3 * plane;
#pragma unroll
        for (int j = 0; j < PH_ROWS; ++j)
            if (valid[j]) {
                const size_t o = (size_t)(ybase + j) * W + x;
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(g + c * plane + o, gt[j][c]);
            }
    }

    // ---- block partial: fixed-order sum over the 8 warps ---------------------
    __syncthreads();
    float* my_partial = partials + ((size_t)(jb * a.B + b) * L.tiles + tile) * PH_PSTRIDE;
    for (int k = tid; k < PH_PSTRIDE; k += PH_THREADS) {
        const int slot = k / PH_NACC, c = k - slot * PH_NACC;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < PH_THREADS / 32; ++w) v += s_acc[w][slot][c];
        __stcg(my_partial + k, v);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(&tickets[jb * a.B + b], 1) == L.tiles - 1);
    __syncthreads();
    if (!s_flag) return;

    // ---- last block of this (job, image): reduce its tiles, pose chain --------
    __threadfence();
    const float* base = partials + (size_t)(jb * a.B + b) * L.tiles * PH_PSTRIDE;
    for (int k = tid; k < PH_PSTRIDE; k += PH_THREADS) {
        float v = 0.0f;
        for (int tI = 0; tI < L.tiles; ++tI) v += __ldcg(base + (size_t)tI * PH_PSTRIDE + k);
        s_red[k] = v;
    }
    __syncthreads();
    if (tid < PH_SLOTS) ws_loss[(size_t)(jb * a.B + b) * PH_SLOTS + tid] = s_red[tid * PH_NACC];
    if (GRAD && tid < job.n_src) {
        const int i = tid;
        float dP[12], dM[12], g6[6];
        for (int k = 0; k < 12; ++k) {
            float v = 0.0f;
            for (int s = 0; s < job.n_scales; ++s) v += s_red[(s * PLB_MAX_SRC + i) * PH_NACC + 1 + k];
            dP[k] = v;
        }
        const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
        kT_times_dP(Kb, a.k_is_f64, dP, dM);
        pose_to_M_vjp(a.poses + ((size_t)b * a.n_pose + job.pose_index[i]) * 6, a.rotation_mode, job.pose_inv[i],
                      dM, g6);
        float* o = ws_pose + ((size_t)(jb * a.B + b) * PLB_MAX_SRC + i) * 6;
        for (int k = 0; k < 6; ++k) o[k] = g6[k];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        tickets[jb * a.B + b] = 0;  // self-cleaning for the next launch
        s_flag = (atomicAdd(&tickets[a.n_jobs * a.B], 1) == a.n_jobs * a.B - 1);
    }
    __syncthreads();
    if (!s_flag) return;

    // ---- last block of the launch: loss scalar and pose gradients -------------
    __threadfence();
    const double inv_n = 1.0 / (3.0 * (double)a.B * (double)H * (double)W);
    if (tid < a.n_jobs * PLB_MAX_SCALES) {
        const int j2 = tid / PLB_MAX_SCALES, s = tid % PLB_MAX_SCALES;
        double e = 0.0;
        if (s < a.jobs[j2].n_scales) {
            for (int i = 0; i < a.jobs[j2].n_src; ++i) {
                double v = 0.0;
                for (int bb = 0; bb < a.B; ++bb)
                    v += (double)__ldcg(ws_loss + (size_t)(j2 * a.B + bb) * PH_SLOTS + s * PLB_MAX_SRC + i);
                e += v;
            }
            e *= inv_n;
        }
        s_red[tid] = (float)(e * (double)a.jobs[j2].term_weight);
        if (a.entry_loss != nullptr) a.entry_loss[tid] = (float)(e / (double)max(a.jobs[j2].n_src, 1));
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int k = 0; k < a.n_jobs * PLB_MAX_SCALES; ++k) tot += (double)s_red[k];
        if (a.loss != nullptr) *a.loss = (float)tot;
        tickets[a.n_jobs * a.B] = 0;
    }
    if (GRAD && a.g_poses != nullptr) {
        for (int k = tid; k < a.B * a.n_pose * 6; k += PH_THREADS) {
            const int bb = k / (a.n_pose * 6), col = (k / 6) % a.n_pose, c = k % 6;
            float v = 0.0f;
            for (int j2 = 0; j2 < a.n_jobs; ++j2)
                for (int i = 0; i < a.jobs[j2].n_src; ++i)
                    if (a.jobs[j2].pose_index[i] == col)
                        v += __ldcg(ws_pose + ((size_t)(j2 * a.B + bb) * PLB_MAX_SRC + i) * 6 + c);
            a.g_poses[k] = v;
        }
    }
}

// Transposed bilinear upsample (gather form, deterministic) + disp->depth chain:
// g_disp[s][b,j,i] = dD/dd * sum over the full-resolution pixels whose
// align_corners=False footprint touches low-res pixel (j,i).
__global__ void __launch_bounds__(256)
photo_upsample_T_kernel(const __grid_constant__ plb_photo_args a) {
    if (skip_launch(a.skip_if_unit, a.skip_n)) return;
    const PhotoLayout L = photo_layout(a);
    const float* gup = (const float*)((const char*)a.workspace + L.gup);
    const int jb = blockIdx.z / PLB_MAX_SCALES, s = blockIdx.z % PLB_MAX_SCALES, b = blockIdx.y;
    if (jb >= a.n_jobs) return;
    const plb_photo_job& job = a.jobs[jb];
    if (s >= job.n_scales || job.g_disp[s] == nullptr) return;
    const int dh = job.dh[s], dw = job.dw[s], H = a.H, W = a.W;
    if (dh == H && dw == W) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dh * dw) return;
    const int i = idx % dw, j = idx / dw;
    const float sx = (float)dw / (float)W, sy = (float)dh / (float)H;
    const float fx = (float)W / (float)dw, fy = (float)H / (float)dh;
    const int xlo = max((int)floorf(((float)i - 1.0f + 0.5f) * fx - 0.5f) - 1, 0);
    const int xhi = min((int)ceilf(((float)i + 1.0f + 0.5f) * fx - 0.5f) + 1, W - 1);
    const int ylo = max((int)floorf(((float)j - 1.0f + 0.5f) * fy - 0.5f) - 1, 0);
    const int yhi = min((int)ceilf(((float)j + 1.0f + 0.5f) * fy - 0.5f) + 1, H - 1);
    const float* g = gup + ((size_t)(jb * PLB_MAX_SCALES + s) * a.B + b) * (size_t)H * W;
    float acc = 0.0f;
    for (int y = ylo; y <= yhi; ++y) {
        int y0, y1; float ly0, ly1;
        up_coord(y, sy, dh, y0, y1, ly0, ly1);
        float wy = (y0 == j ? ly0 : 0.0f) + (y1 == j ? ly1 : 0.0f);
        if (wy == 0.0f) continue;
        float row = 0.0f;
        for (int x = xlo; x <= xhi; ++x) {
            int x0, x1; float lx0, lx1;
            up_coord(x, sx, dw, x0, x1, lx0, lx1);
            float wx = (x0 == i ? lx0 : 0.0f) + (x1 == i ? lx1 : 0.0f);
            if (wx != 0.0f) row += wx * __ldcg(g + (size_t)y * W + x);
        }
        acc += wy * row;
    }
    float chain = 1.0f;
    if (!a.input_is_depth) {
        float d = __ldg(job.disp[s] + (size_t)b * dh * dw + idx);
        float D = 1.0f / (a.disp_a * d + a.disp_b);
        chain = -a.disp_a * D * D;
    }
    job.g_disp[s][(size_t)b * dh * dw + idx] = acc * chain;
}

static int validate_photo(const plb_photo_args* a) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 2 || a->W < 2 || a->n_jobs < 1 || a->n_jobs > PLB_MAX_JOBS || a->n_pose < 1)
        return PLB_EINVAL;
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
        }
    }
    if (a->workspace == nullptr) return PLB_EWORKSPACE;
    if (a->workspace_bytes < photo_layout(*a).total) return PLB_EWORKSPACE;
    return PLB_OK;
}

int photo_l1_launch(const plb_photo_args* a, cudaStream_t st) {
    int rc = validate_photo(a);
    if (rc != PLB_OK) return rc;
    const PhotoLayout L = photo_layout(*a);
    bool img_grad = false, lowres_grad = false;
    for (int j = 0; j < a->n_jobs; ++j) {
        if (a->jobs[j].g_tgt) img_grad = true;
        for (int i = 0; i < a->jobs[j].n_src; ++i)
            if (a->jobs[j].g_src[i]) img_grad = true;
        for (int s = 0; s < a->jobs[j].n_scales; ++s)
            if (a->jobs[j].g_disp[s] && (a->jobs[j].dh[s] != a->H || a->jobs[j].dw[s] != a->W)) lowres_grad = true;
    }
    dim3 grid(L.tiles, a->B, a->n_jobs), block(PH_THREADS);
    if (!a->want_grad)
        photo_l1_kernel<false, false><<<grid, block, 0, st>>>(*a);
    else if (img_grad)
        photo_l1_kernel<true, true><<<grid, block, 0, st>>>(*a);
    else
        photo_l1_kernel<true, false><<<grid, block, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (a->want_grad && lowres_grad) {
        int maxpx = 0;
        for (int j = 0; j < a->n_jobs; ++j)
            for (int s = 0; s < a->jobs[j].n_scales; ++s)
                if (a->jobs[j].dh[s] != a->H || a->jobs[j].dw[s] != a->W)
                    maxpx = max(maxpx, a->jobs[j].dh[s] * a->jobs[j].dw[s]);
        dim3 g2((maxpx + 255) / 256, a->B, a->n_jobs * PLB_MAX_SCALES);
        photo_upsample_T_kernel<<<g2, 256, 0, st>>>(*a);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    return PLB_OK;
}

size_t photo_workspace_bytes(const plb_photo_args* a) { return photo_layout(*a).total; }

}  // namespace plb
