// Fused photometric reprojection loss (live mode: warp + L1 mean), forward and
// gradients in one pass.  See DESIGN.md "photo_l1_kernel".
//
// Work decomposition.  The unit of work is one ROW SEGMENT: 32 consecutive
// pixels of one row of one target image of one job (direction), processed by
// one warp: every global access of the warp - target pixels, disparity, the
// 2x2x3 bilinear taps of each source - is a (nearly) contiguous 128-byte piece
// of a planar NCHW row.  Units are ordered (job, image, 32-px column strip, row),
// so a warp that walks its units moves DOWN a strip and re-uses the source rows
// it has just pulled into L1.
//
// The grid is persistent: 148 x (resident blocks per SM) blocks, and the weighted
// unit list (a unit costs n_scales x n_src warps) is cut into equal contiguous
// ranges, one per warp, so the tail is one row segment long instead of one tile.
// A block's range touches at most two (job, image) pairs; K^-1 and P = K.[R|t]
// for both are built once in the block prologue and kept in shared memory.
//
// Reductions (loss, 3x4 projection-matrix gradient per source) are kept in
// registers across all of a warp's units, reduced warp -> block -> (job, image)
// -> launch in a fixed order by "last block done" epilogues: one launch, no
// atomics on the results, bitwise repeatable.
#include "photo_common.cuh"

namespace plb {

// ---------------------------------------------------------------------------------------------
// One target pixel against one source at one depth: project, sample, L1, and (GRAD) the
// gradient terms.  Written for instruction count: the projection is
// cam = D * (P[:, :3].ray) + P[:, 3], the perspective divide is one MUFU.RCP + one Newton step,
// the normalise/un-normalise chain of the reference (transform.py:143-148 + grid_sample) is the
// identity and is dropped, and the bilinear blend is written as nested lerps whose
// intermediates ARE the coordinate derivatives (d proj / d iy = bot - top).  A warp whose 32
// pixels all land strictly inside the source takes a branch with unpredicated loads.
// ---------------------------------------------------------------------------------------------
template <bool GRAD, bool IMG_GRAD>
__device__ __forceinline__ void photo_pixel(const float* __restrict__ cb, float* gbase, int plane, int pf_rows,
                                            int H, int W, const float4 Pa, const float4 Pb, const float4 Pc,
                                            float rx, float ry, float rz, float D, const float (&t)[3], float w_e,
                                            bool valid, float (&acc)[12], float& l1acc, float& gD, float (&gt)[3]) {
    const float Ax = fmaf(Pa.z, rz, fmaf(Pa.y, ry, Pa.x * rx));
    const float Ay = fmaf(Pb.z, rz, fmaf(Pb.y, ry, Pb.x * rx));
    const float Az = fmaf(Pc.z, rz, fmaf(Pc.y, ry, Pc.x * rx));
    const float cx = fmaf(D, Ax, Pa.w), cy = fmaf(D, Ay, Pb.w);
    const float ze = fmaf(D, Az, Pc.w) + 1e-5f;
    const float inv = rcp_nr(ze);
    // quotient + one residual correction: as accurate as an IEEE divide (the sample position is the
    // difference of two ~W-sized numbers, so every ulp of px is 6e-5 px of bilinear weight)
    float px = cx * inv, py = cy * inv;
    px = fmaf(fmaf(-px, ze, cx), inv, px);
    py = fmaf(fmaf(-py, ze, cy), inv, py);
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
    // 32-bit element offsets from ONE base pointer: an IMAD.WIDE per (channel, row), +4 B as an immediate
    int o00 = y0 * W + x0;
    if (__all_sync(0xffffffffu, inter || !valid)) {
        o00 = (inter && valid) ? o00 : 0;
        const int o01 = o00 + W, o10 = o00 + plane, o11 = o10 + W, o20 = o10 + plane, o21 = o20 + W;
        v[0][0] = __ldg(cb + o00); v[0][1] = __ldg(cb + o00 + 1); v[0][2] = __ldg(cb + o01); v[0][3] = __ldg(cb + o01 + 1);
        v[1][0] = __ldg(cb + o10); v[1][1] = __ldg(cb + o10 + 1); v[1][2] = __ldg(cb + o11); v[1][3] = __ldg(cb + o11 + 1);
        v[2][0] = __ldg(cb + o20); v[2][1] = __ldg(cb + o20 + 1); v[2][2] = __ldg(cb + o21); v[2][3] = __ldg(cb + o21 + 1);
        if (pf_rows > 0) {
            // the warp walks DOWN a strip: the source row needed pf_rows units from now, one line per channel
            const int opf = o01 + pf_rows * W;
            prefetch_l1(cb + opf); prefetch_l1(cb + (opf + plane)); prefetch_l1(cb + (opf + 2 * plane));
        }
        use = valid;
    } else {
        const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
        mnw = valid && vx0 && vy0; mne = valid && vx1 && vy0;
        msw = valid && vx0 && vy1; mse = valid && vx1 && vy1;
        const int o01 = o00 + W, o10 = o00 + plane, o11 = o10 + W, o20 = o10 + plane, o21 = o20 + W;
        v[0][0] = ldg_pred(cb + o00, mnw); v[0][1] = ldg_pred(cb + o00 + 1, mne);
        v[0][2] = ldg_pred(cb + o01, msw); v[0][3] = ldg_pred(cb + o01 + 1, mse);
        v[1][0] = ldg_pred(cb + o10, mnw); v[1][1] = ldg_pred(cb + o10 + 1, mne);
        v[1][2] = ldg_pred(cb + o11, msw); v[1][3] = ldg_pred(cb + o11 + 1, mse);
        v[2][0] = ldg_pred(cb + o20, mnw); v[2][1] = ldg_pred(cb + o20 + 1, mne);
        v[2][2] = ldg_pred(cb + o21, msw); v[2][3] = ldg_pred(cb + o21 + 1, mse);
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
    l1acc += valid ? l1 : 0.0f;
    if (GRAD) {
        const float gi = use ? w_e * inv : 0.0f;  // also keeps the inf/NaN of a degenerate z out of the sums
        const float gcx = Gx * gi, gcy = Gy * gi;
        const float gcz = use ? -(gcx * px + gcy * py) : 0.0f;
        gD += fmaf(gcx, Ax, fmaf(gcy, Ay, gcz * Az));
        const float hx = gcx * D, hy = gcy * D, hz = gcz * D;
        acc[0] = fmaf(hx, rx, acc[0]); acc[1] = fmaf(hx, ry, acc[1]); acc[2] = fmaf(hx, rz, acc[2]); acc[3] += gcx;
        acc[4] = fmaf(hy, rx, acc[4]); acc[5] = fmaf(hy, ry, acc[5]); acc[6] = fmaf(hy, rz, acc[6]); acc[7] += gcy;
        acc[8] = fmaf(hz, rx, acc[8]); acc[9] = fmaf(hz, ry, acc[9]); acc[10] = fmaf(hz, rz, acc[10]); acc[11] += gcz;
        if (IMG_GRAD) {
            const float m = valid ? w_e : 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) gt[c] -= m * e[c];
            if (gbase != nullptr) {
                const float wnw = (1.0f - fx) * (1.0f - fy), wne = fx * (1.0f - fy);
                const float wsw = (1.0f - fx) * fy, wse = fx * fy;
                const int og = y0 * W + x0;  // masks carry the per-tap bounds (all true on the fast path)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ec = m * e[c];
                    float* q = gbase + (og + c * plane);
                    if (valid && mnw) atomicAdd(q, wnw * ec);
                    if (valid && mne) atomicAdd(q + 1, wne * ec);
                    if (valid && msw) atomicAdd(q + W, wsw * ec);
                    if (valid && mse) atomicAdd(q + W + 1, wse * ec);
                }
            }
        }
    }
}

template <int MAXSRC>
__device__ __forceinline__ void flush_acc(float (&acc)[MAXSRC][12], float& l1acc, float* rec, int lane) {
    // rec: this warp's shared record for one set, [PH_NREC] floats, ACCUMULATED in place
#pragma unroll
    for (int i = 0; i < MAXSRC; ++i) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = acc[i][k];
        v[12] = (i == 0) ? l1acc : 0.0f;
        v[13] = v[14] = v[15] = 0.0f;
        int which;
        const float r = warp_reduce16(v, lane, which);
        if ((lane & 1) == 0) {
            if (which < 12) rec[i * 12 + which] += r;
            else if (which == 12 && i == 0) rec[PLB_MAX_SRC * 12] += r;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[i][k] = 0.0f;
    }
    l1acc = 0.0f;
    __syncwarp();
}

template <bool GRAD, bool IMG_GRAD, int MAXSRC>
__global__ void __launch_bounds__(PH_THREADS, (MAXSRC <= 2) ? 3 : 2)
photo_l1_kernel(const __grid_constant__ PhotoLaunch p) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;

    char* ws = (char*)a.workspace;
    int32_t* tickets = (int32_t*)(ws + p.L.tickets);
    float* records = (float*)(ws + p.L.records);
    float* ws_pose = (float*)(ws + p.L.ws_pose);
    float* ws_loss = (float*)(ws + p.L.ws_loss);
    float* gup = (float*)(ws + p.L.gup);

    const int H = a.H, W = a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane = H * W;

    __shared__ PairConst s_pc[2];
    __shared__ float s_rec[2][PH_WARPS][PH_NREC + 3];
    __shared__ float s_red[PH_NREC + 3];
    __shared__ float s_part4[4][64];
    __shared__ int s_pair[2];
    __shared__ int s_flag;

    // ---- this warp's unit range: equal shares of the weighted unit list (32-bit maths) ---------
    const int wt_total = (int)p.weight_start[a.n_jobs];
    auto pos_of = [&](int w) -> int { return w * p.share + min(w, p.share_rem); };
    auto unit_of = [&](int pos) -> int {  // first unit whose weight interval starts at or after pos
        int j = 0;
        while (j + 1 < a.n_jobs && pos >= (int)p.weight_start[j + 1]) ++j;
        const int rel = pos - (int)p.weight_start[j];
        return p.unit_start[j] + (rel + p.unit_weight[j] - 1) / p.unit_weight[j];
    };
    const int gw = blockIdx.x * PH_WARPS + warp;
    const int blk_u0 = unit_of(pos_of(blockIdx.x * PH_WARPS));
    const int blk_u1 = unit_of(pos_of(blockIdx.x * PH_WARPS + PH_WARPS));
    const int u0 = unit_of(pos_of(gw));
    const int u1 = unit_of(pos_of(gw + 1));
    const bool empty_block = blk_u1 <= blk_u0;  // more blocks than work: still publishes (empty) records
    const int pairA = empty_block ? 0 : blk_u0 / p.units_per_pair;
    const int pairB = empty_block ? 0 : (blk_u1 - 1) / p.units_per_pair;

    // ---- block prologue: context of the (at most two) pairs this block touches ------------------
    if (tid < 2) s_pair[tid] = empty_block ? -1 : ((tid == 0) ? pairA : (pairB != pairA ? pairB : -1));
    for (int k = tid; k < 2 * PH_WARPS * (PH_NREC + 3); k += PH_THREADS) (&s_rec[0][0][0])[k] = 0.0f;
    {
        const int set = warp >> 1;  // warps 0,1 -> set 0; warps 2,3 -> set 1
        const int pair = set == 0 ? pairA : pairB;
        if (!empty_block && warp < 4 && (set == 0 || pairB != pairA)) {
            const int jb = pair / a.B, b = pair - jb * a.B;
            const plb_photo_job& job = a.jobs[jb];
            PairConst& pc = s_pc[set];
            const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
            const size_t img = (size_t)b * 3 * plane;
            if ((warp & 1) == 0) {
                if (lane == 0) kinv_f32(Kb, a.k_is_f64, pc.kinv);
                if (lane == 1) {
                    pc.tgt = job.tgt + img;
                    pc.g_tgt = (GRAD && IMG_GRAD && job.g_tgt) ? job.g_tgt + img : nullptr;
                    pc.n_src = job.n_src; pc.n_scales = job.n_scales; pc.lowres = p.lowres[jb];
                    pc.w_e = p.w_e[jb] * (a.upstream ? __ldg(a.upstream) : 1.0f);
                }
                if (lane >= 4 && lane < 4 + job.n_scales) {
                    const int sc = lane - 4;
                    const int dh = job.dh[sc], dw = job.dw[sc];
                    pc.dh[sc] = dh; pc.dw[sc] = dw;
                    pc.sx[sc] = (float)dw / (float)W; pc.sy[sc] = (float)dh / (float)H;
                    pc.disp[sc] = job.disp[sc] + (size_t)b * dh * dw;
                    float* g = nullptr;
                    if (GRAD && job.g_disp[sc] != nullptr)
                        g = ((p.lowres[jb] >> sc) & 1)
                                ? gup + ((size_t)(jb * PLB_MAX_SCALES + sc) * a.B + b) * plane
                                : job.g_disp[sc] + (size_t)b * plane;
                    pc.g_disp[sc] = g;
                }
            } else if (lane < job.n_src) {
                float M[12], P[12];
                pose_to_M(a.poses + ((size_t)b * a.n_pose + job.pose_index[lane]) * 6, a.rotation_mode,
                          job.pose_inv[lane], M);
                k_times_M(Kb, a.k_is_f64, M, P);
                pc.P[lane][0] = make_float4(P[0], P[1], P[2], P[3]);
                pc.P[lane][1] = make_float4(P[4], P[5], P[6], P[7]);
                pc.P[lane][2] = make_float4(P[8], P[9], P[10], P[11]);
                pc.src[lane] = job.src[lane] + img;
                pc.g_src[lane] = (GRAD && IMG_GRAD && job.g_src[lane]) ? job.g_src[lane] + img : nullptr;
            }
        }
    }
    __syncthreads();

    float acc[MAXSRC][12];
    float l1acc = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXSRC; ++i)
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[i][k] = 0.0f;

    // decode the first unit, then step incrementally (no division in the loop)
    int pair = u0 / p.units_per_pair;
    int local = u0 - pair * p.units_per_pair;
    int strip = local / H;
    int y = local - strip * H;
    int set = (pair == pairA) ? 0 : 1;

#pragma unroll 1
    for (int u = u0; u < u1; ++u) {
        const PairConst& pc = s_pc[set];
        const float w_e = pc.w_e;
        const int x = strip * 32 + lane;
        const bool valid = x < W;
        const int o = y * W + min(x, W - 1);

        float t[3], gt[3] = {0.0f, 0.0f, 0.0f};
        const float* tgt_b = pc.tgt;
        t[0] = __ldg(tgt_b + o); t[1] = __ldg(tgt_b + (o + plane)); t[2] = __ldg(tgt_b + (o + 2 * plane));
        const int pf = (y + PH_PREFETCH_ROWS < H) ? PH_PREFETCH_ROWS : 0;   // stay inside this image
        if (pf > 0) {
            const int opf = o + pf * W;
            prefetch_l1(tgt_b + opf); prefetch_l1(tgt_b + (opf + plane)); prefetch_l1(tgt_b + (opf + 2 * plane));
            if (!(pc.lowres & 1)) prefetch_l1(pc.disp[0] + opf);
        }
        const float xf = (float)x, yf = (float)y;
        const float rx = fmaf(pc.kinv[1], yf, pc.kinv[0] * xf) + pc.kinv[2];
        const float ry = fmaf(pc.kinv[4], yf, pc.kinv[3] * xf) + pc.kinv[5];
        const float rz = fmaf(pc.kinv[7], yf, pc.kinv[6] * xf) + pc.kinv[8];
        const int n_scales = pc.n_scales, n_src = pc.n_src, lowres = pc.lowres;

#pragma unroll 1
        for (int s = 0; s < n_scales; ++s) {
            const float* disp_b = pc.disp[s];
            const bool full = !((lowres >> s) & 1);
            float D, gD = 0.0f;
            if (full) {
                const float d = __ldg(disp_b + o);
                D = a.input_is_depth ? d : rcp_nr(fmaf(a.disp_a, d, a.disp_b));
            } else {
                const int dh = pc.dh[s], dw = pc.dw[s];
                int x0, x1, y0, y1; float lx0, lx1, ly0, ly1;
                up_coord(min(x, W - 1), pc.sx[s], dw, x0, x1, lx0, lx1);
                up_coord(y, pc.sy[s], dh, y0, y1, ly0, ly1);
                float v00 = __ldg(disp_b + (y0 * dw + x0)), v01 = __ldg(disp_b + (y0 * dw + x1));
                float v10 = __ldg(disp_b + (y1 * dw + x0)), v11 = __ldg(disp_b + (y1 * dw + x1));
                if (!a.input_is_depth) {
                    v00 = rcp_nr(fmaf(a.disp_a, v00, a.disp_b)); v01 = rcp_nr(fmaf(a.disp_a, v01, a.disp_b));
                    v10 = rcp_nr(fmaf(a.disp_a, v10, a.disp_b)); v11 = rcp_nr(fmaf(a.disp_a, v11, a.disp_b));
                }
                D = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
            }
#pragma unroll
            for (int i = 0; i < MAXSRC; ++i) {
                if (i < n_src) {
                    photo_pixel<GRAD, IMG_GRAD>(pc.src[i], (GRAD && IMG_GRAD) ? pc.g_src[i] : nullptr, plane, pf, H, W,
                                                pc.P[i][0], pc.P[i][1], pc.P[i][2], rx, ry, rz, D, t, w_e, valid,
                                                acc[i], l1acc, gD, gt);
                }
            }
            if (GRAD) {
                float* g = pc.g_disp[s];
                if (valid && g != nullptr) {
                    const float chain = (full && !a.input_is_depth) ? -a.disp_a * D * D : 1.0f;
                    g[o] = gD * chain;
                }
            }
        }
        if (GRAD && IMG_GRAD && valid && pc.g_tgt != nullptr) {
            float* g = pc.g_tgt + o;
            atomicAdd(g, gt[0]); atomicAdd(g + plane, gt[1]); atomicAdd(g + 2 * plane, gt[2]);
        }
        // next unit: down the strip, then the next strip, then the next (job, image)
        if (++y == H) {
            y = 0;
            if (++strip == p.strips) {
                strip = 0;
                flush_acc<MAXSRC>(acc, l1acc, s_rec[set][warp], lane);
                ++pair;
                set = 1;
            }
        }
    }
    if (u1 > u0) flush_acc<MAXSRC>(acc, l1acc, s_rec[set][warp], lane);

    // ---- block records: fixed-order sum over the 8 warps, one record per touched pair --------
    __syncthreads();
    float* my_rec = records + (size_t)blockIdx.x * 2 * PH_REC_STRIDE;
    for (int k = tid; k < 2 * PH_REC_STRIDE; k += PH_THREADS) {
        const int set = k / PH_REC_STRIDE, c = k - set * PH_REC_STRIDE;
        float v;
        if (c == 0) {
            v = __int_as_float(s_pair[set]);
        } else if (c <= PH_NREC) {
            v = 0.0f;
#pragma unroll
            for (int w = 0; w < PH_WARPS; ++w) v += s_rec[set][w][c - 1];
        } else {
            v = 0.0f;
        }
        __stcg(my_rec + k, v);
    }
    __threadfence();
    __syncthreads();

    // ---- per-pair tickets count finished UNITS; whoever completes a pair reduces it ----------
#pragma unroll 1
    for (int set = 0; set < 2; ++set) {
        const int pr = s_pair[set];
        if (pr < 0) continue;
        const int lo = max(blk_u0, pr * p.units_per_pair), hi = min(blk_u1, (pr + 1) * p.units_per_pair);
        if (hi <= lo) continue;
        if (tid == 0) s_flag = (atomicAdd(&tickets[pr], hi - lo) + (hi - lo) == p.units_per_pair);
        __syncthreads();
        const bool last = s_flag != 0;
        __syncthreads();
        if (!last) continue;
        __threadfence();
        const int jb = pr / a.B, b = pr - jb * a.B;
        const plb_photo_job& job = a.jobs[jb];
        // blocks whose range can overlap this pair (widened by one block on each side; records carry the pair id)
        const int w0 = (int)p.weight_start[jb] + (pr * p.units_per_pair - p.unit_start[jb]) * p.unit_weight[jb];
        const int w1 = w0 + p.units_per_pair * p.unit_weight[jb];
        auto warp_of = [&](int pos) -> int {  // inverse of pos_of (the warp whose weight range holds pos)
            const int big = p.share_rem * (p.share + 1);
            if (pos < big) return pos / (p.share + 1);
            return p.share > 0 ? p.share_rem + (pos - big) / p.share : p.n_warps - 1;
        };
        int k_lo = warp_of(w0) / PH_WARPS - 1;
        int k_hi = warp_of(w1) / PH_WARPS + 1;
        k_lo = max(k_lo, 0); k_hi = min(k_hi, p.grid - 1);
        {
            // 256 threads = 4 groups x 64 value lanes; group g takes records g, g+4, ... (fixed order),
            // then the four partial sums are added in group order: deterministic and latency-parallel.
            const int c = tid & 63, grp = tid >> 6;
            const int n_rec = (k_hi - k_lo + 1) * 2;
            float v = 0.0f;
            if (c < PH_NREC) {
#pragma unroll 4
                for (int r_i = grp; r_i < n_rec; r_i += 4) {
                    const float* r = records + ((size_t)k_lo * 2 + r_i) * PH_REC_STRIDE;
                    const int id = __float_as_int(__ldcg(r));
                    const float val = __ldcg(r + 1 + c);
                    v += (id == pr) ? val : 0.0f;
                }
            }
            s_part4[grp][c] = v;
            __syncthreads();
            if (tid < PH_NREC) s_red[tid] = ((s_part4[0][tid] + s_part4[1][tid]) + s_part4[2][tid]) + s_part4[3][tid];
        }
        __syncthreads();
        if (tid == 0) {
            ws_loss[pr] = s_red[PLB_MAX_SRC * 12];
            tickets[pr] = 0;  // self-cleaning for the next launch
        }
        if (GRAD && tid < job.n_src) {
            float dP[12], dM[12], g6[6];
            for (int k = 0; k < 12; ++k) dP[k] = s_red[tid * 12 + k];
            const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
            kT_times_dP(Kb, a.k_is_f64, dP, dM);
            pose_to_M_vjp(a.poses + ((size_t)b * a.n_pose + job.pose_index[tid]) * 6, a.rotation_mode,
                          job.pose_inv[tid], dM, g6);
            float* o = ws_pose + ((size_t)pr * PLB_MAX_SRC + tid) * 6;
            for (int k = 0; k < 6; ++k) o[k] = g6[k];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_flag = (atomicAdd(&tickets[p.n_pairs], 1) == p.n_pairs - 1);
        __syncthreads();
        const bool all_done = s_flag != 0;
        __syncthreads();
        if (!all_done) continue;

        // ---- last pair of the launch: loss scalar and pose gradients --------------------------
        __threadfence();
        if (tid == 0) {
            double tot = 0.0;
            for (int q = 0; q < p.n_pairs; ++q)
                tot += (double)__ldcg(ws_loss + q) * (double)p.w_e[q / a.B];
            if (a.loss != nullptr) *a.loss = (float)tot;
            tickets[p.n_pairs] = 0;
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
}

// Transposed bilinear upsample (gather form, deterministic) + disp->depth chain:
// g_disp[s][b,j,i] = dD/dd * sum over the full-resolution pixels whose align_corners=False
// footprint touches low-res pixel (j,i).  Separable and streaming: one block owns UT_ROWS
// consecutive low-res rows of one image and a chunk of UT_CHUNK full-res columns (+ halo).
// Stage 1: every thread walks DOWN one full-res column of the scratch plane once (coalesced,
// independent loads) and adds each value into the (at most two) low-res rows it feeds; stage 2:
// every low-res pixel gathers its ~2f column sums from shared memory with the exact up_coord
// weights.  The launch is a compact list of (job, scale, image, row group, chunk) work items.
constexpr int UT_THREADS = 256;
constexpr int UT_CHUNK = 192;   // full-res columns owned per block (a multiple of every factor <= 64)
constexpr int UT_HALO = 32;     // >= 1.5 * factor + 2 for factor <= 16
constexpr int UT_ROWS = 8;      // low-res rows per block

struct UpTItem { int jb, s, first_block, groups, chunks; };
struct UpTLaunch {
    int n_items;
    int total_blocks;
    UpTItem items[PLB_MAX_JOBS * PLB_MAX_SCALES];
};

__global__ void __launch_bounds__(UT_THREADS)
photo_upsample_T_kernel(const __grid_constant__ PhotoLaunch p, const __grid_constant__ UpTLaunch u) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    int it = 0;
    while (it + 1 < u.n_items && (int)blockIdx.x >= u.items[it + 1].first_block) ++it;
    const UpTItem item = u.items[it];
    const int local = blockIdx.x - item.first_block;
    const int chunk = local % item.chunks;
    const int grp_all = local / item.chunks;            // (image, row group)
    const int b = grp_all / item.groups, grp = grp_all - b * item.groups;
    const int jb = item.jb, s = item.s;
    const plb_photo_job& job = a.jobs[jb];
    const int dh = job.dh[s], dw = job.dw[s], H = a.H, W = a.W;
    const int j0 = grp * UT_ROWS, j1 = min(j0 + UT_ROWS, dh);   // low-res rows [j0, j1)
    const int xc0 = chunk * UT_CHUNK;
    const int tid = threadIdx.x;
    const float sx = (float)dw / (float)W, sy = (float)dh / (float)H;
    const float fy = (float)H / (float)dh, fx = (float)W / (float)dw;

    constexpr int CW = UT_CHUNK + 2 * UT_HALO;   // 256 columns staged per block
    __shared__ float s_col[UT_ROWS][CW];
    __shared__ float s_l0[CW], s_l1[CW];
    __shared__ int s_x0[CW];
    constexpr int UT_TAPS = 40;                  // >= 2*16 + 5 full-res rows feed one low-res row
    __shared__ float s_wr[UT_ROWS][UT_TAPS];     // per local low-res row: weights of its full-res rows
    __shared__ int s_ybeg[UT_ROWS], s_ycnt[UT_ROWS];

    const float* gup = (const float*)((const char*)a.workspace + p.L.gup);
    const float* g = gup + ((size_t)(jb * PLB_MAX_SCALES + s) * a.B + b) * (size_t)H * W;
    // per low-res row j: conservative full-res window, exact up_coord weights (zero outside the footprint)
    for (int q = tid; q < UT_ROWS * UT_TAPS; q += UT_THREADS) {
        const int r = q / UT_TAPS, k = q - r * UT_TAPS;
        const int j = j0 + r;
        const int ylo = max((int)floorf(((float)j - 0.5f) * fy - 0.5f) - 1, 0);
        const int yhi = min((int)ceilf(((float)j + 1.5f) * fy - 0.5f) + 1, H - 1);
        float wgt = 0.0f;
        if (j < j1 && ylo + k <= yhi) {
            int y0, y1; float ly0, ly1;
            up_coord(ylo + k, sy, dh, y0, y1, ly0, ly1);
            wgt = (y0 == j ? ly0 : 0.0f) + (y1 == j ? ly1 : 0.0f);
        }
        s_wr[r][k] = wgt;
        if (k == 0) { s_ybeg[r] = ylo; s_ycnt[r] = (j < j1) ? min(yhi - ylo + 1, UT_TAPS) : 0; }
    }
    __syncthreads();
    {
        const int k = tid;             // CW == UT_THREADS: one staged column per thread
        const int x = xc0 - UT_HALO + k;
        float l0 = 0.0f, l1 = 0.0f;
        int x0 = -1000000;
        const bool xin = x >= 0 && x < W;
        if (xin) {
            int x1;
            up_coord(x, sx, dw, x0, x1, l0, l1);
            if (x1 == x0) { l0 += l1; l1 = 0.0f; }   // clamped at the right border: both taps are x0
        }
        s_x0[k] = x0; s_l0[k] = l0; s_l1[k] = l1;
#pragma unroll 1
        for (int r = 0; r < UT_ROWS; ++r) {
            float acc = 0.0f;
            if (xin) {
                const float* gx = g + (s_ybeg[r] * W + x);
                const int cnt = s_ycnt[r];
#pragma unroll 4
                for (int t = 0; t < cnt; ++t) acc = fmaf(s_wr[r][t], __ldg(gx + t * W), acc);
            }
            s_col[r][k] = acc;
        }
    }
    __syncthreads();
    // low-res columns whose centre of mass lies in the owned chunk: i in [i_lo, i_hi)
    const int i_lo = (int)ceilf((float)xc0 * sx - 1e-4f);
    const int i_hi = min((int)ceilf((float)min(xc0 + UT_CHUNK, W) * sx - 1e-4f), dw);
    const int ni = i_hi - i_lo;
    for (int q = tid; q < ni * (j1 - j0); q += UT_THREADS) {
        const int r = q / ni, i = i_lo + (q - r * ni);
        const int xlo = max((int)floorf(((float)i - 0.5f) * fx - 0.5f) - 1, 0);
        const int xhi = min((int)ceilf(((float)i + 1.5f) * fx - 0.5f) + 1, W - 1);
        float acc = 0.0f;
        for (int x = xlo; x <= xhi; ++x) {
            const int k = x - (xc0 - UT_HALO);
            if (k < 0 || k >= CW) continue;   // cannot happen while factor <= 16
            const int x0 = s_x0[k];
            const float w = (x0 == i ? s_l0[k] : 0.0f) + (x0 + 1 == i ? s_l1[k] : 0.0f);
            acc = fmaf(w, s_col[r][k], acc);
        }
        float chain = 1.0f;
        const size_t o = (size_t)b * dh * dw + (size_t)(j0 + r) * dw + i;
        if (!a.input_is_depth) {
            const float D = 1.0f / (a.disp_a * __ldg(job.disp[s] + o) + a.disp_b);
            chain = -a.disp_a * D * D;
        }
        job.g_disp[s][o] = acc * chain;
    }
}

int validate_photo(const plb_photo_args* a) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 2 || a->W < 2 || a->n_jobs < 1 || a->n_jobs > PLB_MAX_JOBS || a->n_pose < 1)
        return PLB_EINVAL;
    if ((long long)a->H * a->W * 3 >= (1LL << 31)) return PLB_EINVAL;
    if ((long long)a->H * ((a->W + 31) / 32) * a->B * a->n_jobs * PLB_MAX_SCALES * PLB_MAX_SRC >= (1LL << 31)) return PLB_EINVAL;
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
            if (job.dw[s] * 16 < a->W || job.dh[s] * 16 < a->H) return PLB_EINVAL;  // upsampling factor <= 16
        }
    }
    if (a->workspace == nullptr) return PLB_EWORKSPACE;
    if (a->workspace_bytes < photo_layout(*a).total) return PLB_EWORKSPACE;
    return PLB_OK;
}

template <bool GRAD, bool IMG, int MS>
static int blocks_per_sm() {
    static int cached = 0;
    if (cached == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photo_l1_kernel<GRAD, IMG, MS>, PH_THREADS, 0) != cudaSuccess ||
            n < 1) {
            (void)cudaGetLastError();
            n = 2;
        }
        cached = n;
    }
    return cached;
}

static int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = 148;
        }
        cached = n;
    }
    return cached;
}

int photo_upsample_T_launch(const PhotoLaunch& p, cudaStream_t st) {
    const plb_photo_args* a = &p.a;
    {
        UpTLaunch u;
        u.n_items = 0;
        u.total_blocks = 0;
        for (int j = 0; j < a->n_jobs; ++j)
            for (int s = 0; s < a->jobs[j].n_scales; ++s) {
                const plb_photo_job& job = a->jobs[j];
                if (!job.g_disp[s] || (job.dh[s] == a->H && job.dw[s] == a->W)) continue;
                UpTItem& it = u.items[u.n_items++];
                it.jb = j; it.s = s; it.first_block = u.total_blocks;
                it.groups = (job.dh[s] + UT_ROWS - 1) / UT_ROWS;
                it.chunks = (a->W + UT_CHUNK - 1) / UT_CHUNK;
                u.total_blocks += it.groups * it.chunks * a->B;
            }
        photo_upsample_T_kernel<<<u.total_blocks, UT_THREADS, 0, st>>>(p, u);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    return PLB_OK;
}

int photo_l1_launch(const plb_photo_args* a, cudaStream_t st) {
    if (a != nullptr && a->n_jobs >= 1 && a->n_jobs <= PLB_MAX_JOBS && a->jobs[0].mode == PLB_PHOTO_MIN_REPROJ)
        return photo_min_launch(a, st);
    int rc = validate_photo(a);
    if (rc != PLB_OK) return rc;
    PhotoLaunch p;
    p.a = *a;
    p.L = photo_layout(*a);
    bool img_grad = false;
    int maxsrc = 1;
    for (int j = 0; j < a->n_jobs; ++j) {
        if (a->jobs[j].g_tgt) img_grad = true;
        for (int i = 0; i < a->jobs[j].n_src; ++i)
            if (a->jobs[j].g_src[i]) img_grad = true;
        if (a->jobs[j].n_src > maxsrc) maxsrc = a->jobs[j].n_src;
    }
    const bool lowres_grad = photo_has_lowres_grad(*a);
    p.strips = (a->W + 31) / 32;
    p.units_per_pair = p.strips * a->H;
    p.n_pairs = a->n_jobs * a->B;
    long long wsum = 0;
    int usum = 0, min_w = 1 << 30;
    for (int j = 0; j < PLB_MAX_JOBS; ++j) {
        p.weight_start[j] = wsum;
        p.unit_start[j] = usum;
        p.unit_weight[j] = 1;
        p.w_e[j] = 0.0f;
        p.lowres[j] = 0;
        if (j < a->n_jobs) {
            p.unit_weight[j] = a->jobs[j].n_scales * a->jobs[j].n_src;
            if (p.unit_weight[j] < min_w) min_w = p.unit_weight[j];
            wsum += (long long)p.unit_weight[j] * p.units_per_pair * a->B;
            usum += p.units_per_pair * a->B;
            p.w_e[j] = a->jobs[j].term_weight / (3.0f * (float)a->B * (float)a->H * (float)a->W);
            for (int s = 0; s < a->jobs[j].n_scales; ++s)
                if (a->jobs[j].dh[s] != a->H || a->jobs[j].dw[s] != a->W) p.lowres[j] |= 1 << s;
        }
    }
    p.weight_start[PLB_MAX_JOBS] = wsum;
    p.unit_start[PLB_MAX_JOBS] = usum;

    int bps;
    if (!a->want_grad) bps = maxsrc <= 2 ? blocks_per_sm<false, false, 2>() : blocks_per_sm<false, false, 4>();
    else if (img_grad) bps = maxsrc <= 2 ? blocks_per_sm<true, true, 2>() : blocks_per_sm<true, true, 4>();
    else bps = maxsrc <= 2 ? blocks_per_sm<true, false, 2>() : blocks_per_sm<true, false, 4>();
    long long grid = (long long)sm_count() * bps;
    // a block's weight range must not exceed the lightest pair, so that it touches at most two pairs
    const long long pair_w_min = (long long)min_w * p.units_per_pair;
    const long long need = (wsum + pair_w_min - 1) / pair_w_min + 1;
    if (grid > usum) grid = usum;                  // tiny problems: no more blocks than units ...
    if (grid < need) grid = need;                  // ... but never so few that a block spans three pairs
    if (grid > photo_max_grid(*a)) grid = photo_max_grid(*a);
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.n_warps = p.grid * PH_WARPS;
    p.share = (int)(wsum / p.n_warps);
    p.share_rem = (int)(wsum % p.n_warps);

    dim3 g(p.grid), block(PH_THREADS);
#define PLB_LAUNCH(G, I, M) photo_l1_kernel<G, I, M><<<g, block, 0, st>>>(p)
    if (!a->want_grad) { if (maxsrc <= 2) PLB_LAUNCH(false, false, 2); else PLB_LAUNCH(false, false, 4); }
    else if (img_grad) { if (maxsrc <= 2) PLB_LAUNCH(true, true, 2); else PLB_LAUNCH(true, true, 4); }
    else { if (maxsrc <= 2) PLB_LAUNCH(true, false, 2); else PLB_LAUNCH(true, false, 4); }
#undef PLB_LAUNCH
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (lowres_grad) {
        const int rc2 = photo_upsample_T_launch(p, st);
        if (rc2 != PLB_OK) return rc2;
    }
    return PLB_OK;
}

size_t photo_workspace_bytes(const plb_photo_args* a) { return photo_layout(*a).total; }

}  // namespace plb
