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

// One target pixel against one source at one depth: project, sample, L1, and
// (GRAD) the gradient terms.  Written for instruction count: the projection is
// cam = D * (P[:, :3].ray) + P[:, 3] (12 FMA), the perspective divide is one
// MUFU.RCP + one Newton step, the normalise/un-normalise chain of the reference
// (transform.py:143-148 + grid_sample) is the identity and is dropped, and the
// bilinear blend is written as nested lerps whose intermediates ARE the
// coordinate derivatives (d proj/d iy = bot - top).  A warp whose 32 pixels all
// land strictly inside the source takes a branch with unpredicated loads.
template <bool GRAD, bool IMG_GRAD>
__device__ __forceinline__ void photo_pixel(const float* const (&cb)[3], float* const (&gb)[3],
                                            int H, int W, const float* __restrict__ P, float rx, float ry,
                                            float rz, float D, const float (&t)[3], float w_e, bool valid,
                                            float (&acc)[16], float& gD, float (&gt)[3]) {
    const float Ax = fmaf(P[2], rz, fmaf(P[1], ry, P[0] * rx));
    const float Ay = fmaf(P[6], rz, fmaf(P[5], ry, P[4] * rx));
    const float Az = fmaf(P[10], rz, fmaf(P[9], ry, P[8] * rx));
    const float cx = fmaf(D, Ax, P[3]), cy = fmaf(D, Ay, P[7]);
    const float ze = fmaf(D, Az, P[11]) + 1e-5f;
    const float inv = rcp_nr(ze);
    const float px = cx * inv, py = cy * inv;
    // clamp keeps float->int defined; NaN maps to -2 (out of the image)
    const float ixc = fminf(fmaxf(px, -2.0f), (float)(W + 1));
    const float iyc = fminf(fmaxf(py, -2.0f), (float)(H + 1));
    const float xf = floorf(ixc), yf = floorf(iyc);
    const int x0 = (int)xf, y0 = (int)yf;
    const float fx = ixc - xf, fy = iyc - yf;
    const bool inter = ((unsigned)x0 < (unsigned)(W - 1)) && ((unsigned)y0 < (unsigned)(H - 1));
    float v[3][4];
    bool use;
    bool mnw = true, mne = true, msw = true, mse = true;
    if (__all_sync(0xffffffffu, inter || !valid)) {
        // 32-bit element offsets: one IMAD.WIDE per (channel, row), +4 B as an immediate
        const int off0 = (inter && valid) ? y0 * W + x0 : 0;
        const int off1 = off0 + W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* q0 = cb[c] + off0;
            const float* q1 = cb[c] + off1;
            v[c][0] = __ldg(q0);
            v[c][1] = __ldg(q0 + 1);
            v[c][2] = __ldg(q1);
            v[c][3] = __ldg(q1 + 1);
        }
        use = valid;
    } else {
        const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
        mnw = valid && vx0 && vy0; mne = valid && vx1 && vy0;
        msw = valid && vx0 && vy1; mse = valid && vx1 && vy1;
        const int off0 = y0 * W + x0, off1 = off0 + W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* q0 = cb[c] + off0;
            const float* q1 = cb[c] + off1;
            v[c][0] = ldg_pred(q0, mnw);
            v[c][1] = ldg_pred(q0 + 1, mne);
            v[c][2] = ldg_pred(q1, msw);
            v[c][3] = ldg_pred(q1 + 1, mse);
        }
        use = mnw || mne || msw || mse;
    }
    float Gx = 0.0f, Gy = 0.0f, l1 = 0.0f;
    float e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float dA = v[c][1] - v[c][0], dB = v[c][3] - v[c][2];
        const float top = fmaf(fx, dA, v[c][0]), bot = fmaf(fx, dB, v[c][2]);
        const float dV = bot - top;
        const float proj = fmaf(fy, dV, top);
        const float d = proj - t[c];
        l1 += fabsf(d);
        if (GRAD) {
            const float sg = (d > 0.0f ? 1.0f : 0.0f) - (d < 0.0f ? 1.0f : 0.0f);
            e[c] = sg;
            Gx = fmaf(sg, fmaf(fy, dB - dA, dA), Gx);
            Gy = fmaf(sg, dV, Gy);
        }
    }
    acc[0] += valid ? l1 : 0.0f;
    if (GRAD) {
        const float gi = use ? w_e * inv : 0.0f;       // also keeps inf/NaN of a degenerate z out
        const float gcx = Gx * gi, gcy = Gy * gi;
        const float gcz = use ? -(gcx * px + gcy * py) : 0.0f;
        gD += fmaf(gcx, Ax, fmaf(gcy, Ay, gcz * Az));
        const float hx = gcx * D, hy = gcy * D, hz = gcz * D;
        acc[1] = fmaf(hx, rx, acc[1]); acc[2] = fmaf(hx, ry, acc[2]); acc[3] = fmaf(hx, rz, acc[3]); acc[4] += gcx;
        acc[5] = fmaf(hy, rx, acc[5]); acc[6] = fmaf(hy, ry, acc[6]); acc[7] = fmaf(hy, rz, acc[7]); acc[8] += gcy;
        acc[9] = fmaf(hz, rx, acc[9]); acc[10] = fmaf(hz, ry, acc[10]); acc[11] = fmaf(hz, rz, acc[11]); acc[12] += gcz;
        if (IMG_GRAD) {
            const float m = valid ? w_e : 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gt[c] -= m * e[c];
            if (gb[0] != nullptr) {
                const float wnw = (1.0f - fx) * (1.0f - fy), wne = fx * (1.0f - fy);
                const float wsw = (1.0f - fx) * fy, wse = fx * fy;
                const int off0 = y0 * W + x0;   // masks carry the per-tap bounds (all true on the fast path)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ec = m * e[c];
                    float* q = gb[c] + off0;
                    if (valid && mnw) atomicAdd(q, wnw * ec);
                    if (valid && mne) atomicAdd(q + 1, wne * ec);
                    if (valid && msw) atomicAdd(q + W, wsw * ec);
                    if (valid && mse) atomicAdd(q + W + 1, wse * ec);
                }
            }
        }
    }
}

template <bool GRAD, bool IMG_GRAD>
__global__ void __launch_bounds__(PH_THREADS, GRAD ? 3 : 4)
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
#pragma unroll
    for (int j = 0; j < PH_ROWS; ++j) gt[j][0] = gt[j][1] = gt[j][2] = 0.0f;

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
            const float* const cb[3] = {src_b, src_b + plane, src_b + 2 * plane};
            float* const gb[3] = {g_src_b, g_src_b ? g_src_b + plane : nullptr, g_src_b ? g_src_b + 2 * plane : nullptr};
#pragma unroll
            for (int j = 0; j < PH_ROWS; ++j) {
                photo_pixel<GRAD, IMG_GRAD>(cb, gb, H, W, s_P[i], rx[j], ry[j], rz[j], D[j], t[j], w_e,
                                            valid[j], acc, gD[j], gt[j]);
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
