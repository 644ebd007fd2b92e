// Fused SSIM + L1 photometric loss with per-pixel minimum reprojection and automasking
// (the reference's dormant path: losses.py:12-84,94-96,154-162 and
// notes/toy_problem/losses.py:107-129), forward and gradients in one launch.
//
//   rp_i  = 0.85 * clamp((1 - SSIM3x3(warp_i, tgt)) / 2, 0, 1) + 0.15 * |tgt - warp_i|     per channel
//   m     = min_i rp_i ;  mu = [m < min_i photo(src_i, tgt)]  (automask) ;  loss = mean_px max_c (mu * m)
//   scales are averaged (losses.py:181); low scales warp with the bilinearly upsampled depth.
//
// One block = one 32x8 tile of one target image, all scales and sources.  SSIM needs the warped
// image on a 3x3 neighbourhood and its gradient needs the SSIM terms of the 3x3 neighbours, so
// per scale the block (1) warps the tile + 2-pixel halo into shared memory (reflected
// coordinates at the image border, exactly ReflectionPad2d(1)), (2) evaluates SSIM / the
// photometric mix / min / mask / channel max on tile + 1 halo, keeping for every pixel the
// selected (channel, source) and the three coefficients of d rp / d warped(q) = a + b*x_q + c*y_q,
// (3) gathers d loss / d warped for the tile pixels from their 3x3 neighbours (with the
// multiplicities reflection padding induces) and pushes it through the bilinear sampler and the
// projection as the L1 kernel does.  Reductions are fixed-order (bitwise repeatable).
//
// The clip of compute_photometric_loss (losses.py:79-82: every map clamped at mean + 0.5 std of that whole map,
// threshold detached) is a grid-wide dependency, so PLB_PHOTO_CLIP runs the kernel twice: the STATS variant stops
// after stage (2) and leaves per-block (sum, sum of squares) of every map - each warped source at each scale, each
// automask reference - pmin_thresholds_kernel reduces them in block order (fp64) into one threshold per map, and
// the main variant clamps each term (a clamped term carries no gradient) before the min / mask / max selections.
//
// Image gradients (IMG variant): d loss / d warped is scattered through the bilinear weights into the source
// gradients (float atomics, or the 64-bit fixed-point accumulators of the deterministic mode); the target enters
// SSIM as y and the L1 term, d rp / d y_q = da + cb * y_q + cc * x_q - cl at the centre, gathered with the same 3x3
// pass and written directly (every target pixel has one owner).  The automask references only feed a comparison.
#include "photo_common.cuh"

namespace plb {

int validate_photo(const plb_photo_args* a);  // photo.cu

constexpr int PM_THREADS = 256;
constexpr int PM_W2 = PM_TW + 4, PM_H2 = PM_TH + 4;   // tile + 2 halo: 36 x 12
constexpr int PM_W1 = PM_TW + 2, PM_H1 = PM_TH + 2;   // tile + 1 halo: 34 x 10
constexpr int PM_N2 = PM_W2 * PM_H2;                  // 432
constexpr int PM_N1 = PM_W1 * PM_H1;                  // 340

struct PminLaunch {
    plb_photo_args a;
    PhotoLayout L;
    int tiles_x, tiles;
    float w_e;   // term_weight / (B*H*W): weight of one pixel's max_c in the loss
    float C1, C2;
    int clip;    // PLB_PHOTO_CLIP: thresholds at L.pm_thr (written by pmin_thresholds_kernel)
    int det;     // deterministic source-image gradients: int64 accumulators at L.detacc
    int det_src[PLB_MAX_SRC];
};
constexpr float PM_DET_ONE = 16777216.0f;   // 2^24: the source-gradient accumulators count 2^-24 of (w_e x upstream)
enum { PM_VAR_MAIN = 0, PM_VAR_STATS = 1, PM_VAR_IMG = 2 };

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // ReflectionPad2d(1) index rule extended so that any tile position maps inside the image
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

struct MinProj { float Ax, Ay, Az, inv, px, py, fx, fy; int o00; unsigned mask; };

__device__ __forceinline__ void pm_project(const float4 Pa, const float4 Pb, const float4 Pc, float rx, float ry,
                                           float rz, float D, int H, int W, MinProj& q) {
    q.Ax = fmaf(Pa.z, rz, fmaf(Pa.y, ry, Pa.x * rx));
    q.Ay = fmaf(Pb.z, rz, fmaf(Pb.y, ry, Pb.x * rx));
    q.Az = fmaf(Pc.z, rz, fmaf(Pc.y, ry, Pc.x * rx));
    const float cx = fmaf(D, q.Ax, Pa.w), cy = fmaf(D, q.Ay, Pb.w);
    const float ze = fmaf(D, q.Az, Pc.w) + 1e-5f;
    q.inv = rcp_nr(ze);
    float px = cx * q.inv, py = cy * q.inv;
    px = fmaf(fmaf(-px, ze, cx), q.inv, px);
    py = fmaf(fmaf(-py, ze, cy), q.inv, py);
    q.px = px; q.py = py;
    const float ixc = fminf(fmaxf(px, -2.0f), (float)(W + 1));
    const float iyc = fminf(fmaxf(py, -2.0f), (float)(H + 1));
    const float xf = floorf(ixc), yf = floorf(iyc);
    const int x0 = (int)xf, y0 = (int)yf;
    q.fx = ixc - xf; q.fy = iyc - yf;
    q.o00 = y0 * W + x0;
    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
    q.mask = (vx0 && vy0 ? 1u : 0u) | (vx1 && vy0 ? 2u : 0u) | (vx0 && vy1 ? 4u : 0u) | (vx1 && vy1 ? 8u : 0u);
}

__device__ __forceinline__ void pm_taps(const float* __restrict__ cb, int plane, int W, const MinProj& q,
                                        float (&v)[3][4]) {
    // one vote of the lanes that are here: every footprint inside the source (the common case) -> plain loads, six row
    // pointers and immediate offsets instead of twelve predicated loads with their own address arithmetic
    if (__all_sync(__activemask(), q.mask == 15u)) {
        const float* __restrict__ p0 = cb + q.o00;
        const float* __restrict__ p1 = p0 + W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[c][0] = __ldg(p0 + c * plane); v[c][1] = __ldg(p0 + c * plane + 1);
            v[c][2] = __ldg(p1 + c * plane); v[c][3] = __ldg(p1 + c * plane + 1);
        }
        return;
    }
    const bool mnw = q.mask & 1u, mne = q.mask & 2u, msw = q.mask & 4u, mse = q.mask & 8u;
    const int o00 = q.o00, o01 = o00 + W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[c][0] = ldg_pred(cb + (o00 + c * plane), mnw); v[c][1] = ldg_pred(cb + (o00 + c * plane) + 1, mne);
        v[c][2] = ldg_pred(cb + (o01 + c * plane), msw); v[c][3] = ldg_pred(cb + (o01 + c * plane) + 1, mse);
    }
}

// depth at full-resolution pixel (x, y) of scale s (direct, or align_corners=False upsample of depth)
template <bool HEAD>
__device__ __forceinline__ float pm_depth(const plb_photo_args& a, const PairConst& pc, int s, bool full, int x, int y,
                                          int W) {
    if (full) {
        float d = __ldg(pc.disp[s] + (y * W + x));
        if (HEAD) d = head_disp(d, a.head_alpha, a.head_beta);
        return a.input_is_depth == PLB_INPUT_DEPTH ? d : rcp_nr(fmaf(a.disp_a, d, a.disp_b));
    }
    const float* disp_b = pc.disp[s];
    const int dh = pc.dh[s], dw = pc.dw[s];
    int x0, x1, y0, y1; float lx0, lx1, ly0, ly1;
    up_coord(x, pc.sx[s], dw, x0, x1, lx0, lx1);
    up_coord(y, pc.sy[s], dh, y0, y1, ly0, ly1);
    float v00 = __ldg(disp_b + (y0 * dw + x0)), v01 = __ldg(disp_b + (y0 * dw + x1));
    float v10 = __ldg(disp_b + (y1 * dw + x0)), v11 = __ldg(disp_b + (y1 * dw + x1));
    if (HEAD) {
        v00 = head_disp(v00, a.head_alpha, a.head_beta); v01 = head_disp(v01, a.head_alpha, a.head_beta);
        v10 = head_disp(v10, a.head_alpha, a.head_beta); v11 = head_disp(v11, a.head_alpha, a.head_beta);
    }
    if (a.input_is_depth != PLB_INPUT_DEPTH) {
        v00 = rcp_nr(fmaf(a.disp_a, v00, a.disp_b)); v01 = rcp_nr(fmaf(a.disp_a, v01, a.disp_b));
        v10 = rcp_nr(fmaf(a.disp_a, v10, a.disp_b)); v11 = rcp_nr(fmaf(a.disp_a, v11, a.disp_b));
    }
    return ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
}

// Shared memory of one block (dynamic: ~70 KB with two sources, three blocks per SM).  Sources live in PAIRS - the two
// halves of a float2 - so that stage 2 runs both on the packed fp32 pipe with one 64-bit load per value; the target is
// stored twice (t, t) for the same reason.
constexpr int PM_NT = PM_TW * PM_TH;        // tile pixels = threads
template <int NSMAX, int VAR>
struct __align__(16) PminSmem {
    static constexpr bool kImg = VAR == PM_VAR_IMG, kStats = VAR == PM_VAR_STATS;
    static constexpr int kState = kStats ? 1 : PM_NT, kImgState = kImg ? PM_NT : 1, kStat = kStats ? PM_MAXMAPS : 1;
    float4 coef[3][PM_N1];                  // per channel, the min-source term: d rp / d x_q = ca + cb x_q + cc y_q (+ cl at the centre)
    PairConst pc;
    float2 T2[3][PM_N2];                    // target (t, t), tile + 2 (reflected at the border)
    float2 X2[NSMAX / 2][3][PM_N2];         // warped (or, for the automask pass, raw) sources (2g, 2g + 1)
    float aut[3][PM_N1];                    // min_i photo(src_i, tgt) per channel
    float val[3][PM_N1];                    // min_i rp_i per channel
    signed char src[3][PM_N1];              // argmin_i
    signed char sel[PM_N1];                 // channel + 4 * source of the term max_c selects, -1 = none
    float rec[PM_THREADS / 32][PH_NREC + 3];
    float red[PH_NREC + 3];
    float part4[4][64];
    float thr[PM_MAXMAPS];                  // clip thresholds (3e38: no clip)
    // sampler state of the tile pixels, left by stage 1 for stage 3 (no second projection / tap gather):
    // d warped_c / d px, d warped_c / d py, (1/z, px, py) - zero where no tap is inside the source - and the depth
    float sdx[NSMAX][3][kState], sdy[NSMAX][3][kState], sgeo[NSMAX][3][kState], sD[kState];
    float sfx[NSMAX][kImgState], sfy[NSMAX][kImgState];      // IMG: bilinear fractions, first tap and tap mask
    int so00[NSMAX][kImgState];
    unsigned smask[NSMAX][kImgState];
    float da[3][kImg ? PM_N1 : 1];          // IMG: constant term of d rp / d y_q (the target's side of SSIM)
    double stat[PM_THREADS / 32][kStat][2]; // STATS: per-warp (sum, sum of squares) of every map
    int flag;
};

// photometric value of one (pixel, channel) from its 3x3 window sums, and the coefficients of its
// derivative w.r.t. the predicted image at window position q:
//   d rp / d x_q = ca + cb * x_q + cc * y_q   (+ cl at the centre)
// evaluated for the two sources of a pair at once: every arithmetic step is one packed instruction (the target's
// sums ride along in both halves), only compares / clamps / the reciprocal seed are per half.
struct Photo { float rp, ca, cb, cc, cl, da; };
struct Photo2 { float2 rp, ca, cb, cc, cl, da; };
__device__ __forceinline__ float2 f2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 v_neg(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ Photo photo_half(const Photo2& t, int q) {
    Photo r;
    r.rp = v_get(t.rp, q); r.ca = v_get(t.ca, q); r.cb = v_get(t.cb, q);
    r.cc = v_get(t.cc, q); r.cl = v_get(t.cl, q); r.da = v_get(t.da, q);
    return r;
}
template <bool WANT_DA>
__device__ __forceinline__ Photo2 photo_from_sums2(float2 sx, float2 sxx, float2 sxy, float2 sy, float2 syy, float2 xc,
                                                   float2 yc, float C1, float C2, bool no_ssim) {
    Photo2 r;
    const float2 diff = v_add(xc, v_neg(yc));
    float2 sg;
    sg.x = (diff.x > 0.0f ? 1.0f : 0.0f) - (diff.x < 0.0f ? 1.0f : 0.0f);
    sg.y = (diff.y > 0.0f ? 1.0f : 0.0f) - (diff.y < 0.0f ? 1.0f : 0.0f);
    const float2 ad = make_float2(fabsf(diff.x), fabsf(diff.y));
    if (no_ssim) {
        r.rp = ad; r.ca = r.cb = r.cc = r.da = f2(0.0f); r.cl = sg;
        return r;
    }
    // (negations are operand modifiers of the packed instructions: free)
    const float2 i9 = f2(1.0f / 9.0f);
    const float2 mux = v_mul(sx, i9), muy = v_mul(sy, i9);
    const float2 mxx = v_mul(mux, mux), myy = v_mul(muy, muy), mxy = v_mul(mux, muy);
    const float2 sigx = v_fma(sxx, i9, v_neg(mxx)), sigy = v_fma(syy, i9, v_neg(myy)), sigxy = v_fma(sxy, i9, v_neg(mxy));
    const float2 c1 = f2(C1), c2 = f2(C2);
    const float2 N1 = v_fma(f2(2.0f), mxy, c1), N2 = v_fma(f2(2.0f), sigxy, c2);
    const float2 D1 = v_add(v_add(mxx, myy), c1), D2 = v_add(v_add(sigx, sigy), c2);
    const float2 DD = v_mul(D1, D2);
    float2 iD;                                           // one reciprocal for both denominators
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iD.x) : "f"(DD.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iD.y) : "f"(DD.y));
    iD = v_mul(iD, v_fma(v_neg(DD), iD, f2(2.0f)));      // Newton step
    const float2 iD1 = v_mul(D2, iD), iD2 = v_mul(D1, iD);
    const float2 ssim = v_mul(v_mul(N1, N2), iD);
    const float2 h = v_fma(ssim, f2(-0.5f), f2(0.5f));   // (1 - ssim) / 2
    const float2 hc = make_float2(fminf(fmaxf(h.x, 0.0f), 1.0f), fminf(fmaxf(h.y, 0.0f), 1.0f));
    r.rp = v_fma(f2(0.85f), hc, v_mul(f2(0.15f), ad));
    // d ssim / d x_q = (2/9) { [mu_y (N2 - N1)]/(D1 D2) - ssim mu_x (1/D1 - 1/D2) }  +  x_q (-(2/9) ssim / D2)
    //                  + y_q ((2/9) N1 / (D1 D2));   d rp / d x_q = -0.425 * (that) inside the clamp
    const float kk = -0.425f * 2.0f * (1.0f / 9.0f);
    const float2 k = make_float2((h.x > 0.0f && h.x < 1.0f) ? kk : 0.0f, (h.y > 0.0f && h.y < 1.0f) ? kk : 0.0f);
    const float2 dN = v_mul(v_add(N2, v_neg(N1)), iD), ndI = v_add(iD2, v_neg(iD1));
    r.ca = v_mul(k, v_fma(v_mul(ssim, mux), ndI, v_mul(muy, dN)));
    // (SSIM is symmetric in x and y: d ssim / d y_q has the same x_q / y_q coefficients swapped and this constant)
    r.da = WANT_DA ? v_mul(k, v_fma(v_mul(ssim, muy), ndI, v_mul(mux, dN))) : f2(0.0f);
    r.cb = v_mul(v_neg(k), v_mul(ssim, iD2));
    r.cc = v_mul(k, v_mul(N1, iD));
    r.cl = v_mul(f2(0.15f), sg);
    return r;
}

// Stage 2 work item = (channel, column of tile + 1, group of 5 rows): 3 x 34 x 2 = 204 threads.  The
// thread walks DOWN its column keeping the last three horizontal 3-sums of y, y^2 and, for the source pair, x,
// x^2, x y in registers (separable box filter: 3 shared loads per row for BOTH sources instead of 18 each), and
// emits one photometric term per row: min over the sources of the pair.
//   MODE 0: first pair of the scale: write val / src / coef;  MODE 1: second pair: update when smaller;
//   MODE 2 / 3: the same for the automask reference (raw sources), value only.
constexpr int PM_S2_ITEMS = 3 * PM_W1 * 2;
template <int MODE, int VAR, class SM>
__device__ __forceinline__ void pm_stage2(SM& S, int g, int n_here, int map0, int tid, int tx0, int ty0, int H, int W, float C1,
                                          float C2, bool no_ssim) {
    const int i0 = 2 * g;
    float st[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) st[q][0] = st[q][1] = 0.0f;
    if (tid < PM_S2_ITEMS) {
        const int ch = tid / (2 * PM_W1), r = tid - ch * (2 * PM_W1);
        const int grp = r / PM_W1, col = r - grp * PM_W1;
        const float2* __restrict__ T = S.T2[ch];
        const float2* __restrict__ X = S.X2[g][ch];
        int k = (5 * grp) * PM_W2 + col;        // tile + 2 index of the window's top-left corner
        float2 hy[3], hyy[3], hx[3], hxx[3], hxy[3];
        float2 yc = f2(0.0f), xc = f2(0.0f);
        float thr[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) thr[q] = S.thr[map0 + q];
        const int gx = tx0 + col - 1;
        const bool colin = gx >= 0 && gx < W;
#pragma unroll
        for (int rr = 0; rr < 7; ++rr, k += PM_W2) {
            const int slot = rr % 3;
            const float2 y0 = T[k], y1 = T[k + 1], y2 = T[k + 2];
            const float2 x0 = X[k], x1 = X[k + 1], x2 = X[k + 2];
            hy[slot] = v_add(v_add(y0, y1), y2);
            hyy[slot] = v_fma(y2, y2, v_fma(y1, y1, v_mul(y0, y0)));
            hx[slot] = v_add(v_add(x0, x1), x2);
            hxx[slot] = v_fma(x2, x2, v_fma(x1, x1, v_mul(x0, x0)));
            hxy[slot] = v_fma(x2, y2, v_fma(x1, y1, v_mul(x0, y0)));
            if (rr >= 2) {
                const int o = 5 * grp + rr - 2;            // row of tile + 1
                const int p1 = o * PM_W1 + col;
                const int gy = ty0 + o - 1;
                const bool in = colin && gy >= 0 && gy < H;
                const Photo2 t2 = photo_from_sums2<VAR == PM_VAR_IMG>(
                    v_add(v_add(hx[0], hx[1]), hx[2]), v_add(v_add(hxx[0], hxx[1]), hxx[2]), v_add(v_add(hxy[0], hxy[1]), hxy[2]),
                    v_add(v_add(hy[0], hy[1]), hy[2]), v_add(v_add(hyy[0], hyy[1]), hyy[2]), xc, yc, C1, C2, no_ssim);
                Photo m;
                int mi = i0;
                if (VAR == PM_VAR_STATS) {
                    // the map's own statistics: every pixel of the image once (the tile's interior)
                    if (in && col >= 1 && col <= PM_TW && o >= 1 && o <= PM_TH) {
                        st[0][0] += t2.rp.x; st[0][1] = fmaf(t2.rp.x, t2.rp.x, st[0][1]);
                        st[1][0] += t2.rp.y; st[1][1] = fmaf(t2.rp.y, t2.rp.y, st[1][1]);
                    }
                    m.rp = 0.0f; m.ca = m.cb = m.cc = m.cl = m.da = 0.0f;
                } else {
                    // torch.clamp(max=thr): a clamped value is the threshold and passes no gradient (losses.py:82); then
                    // the smaller of the two halves (first one on ties, a NaN never wins) - one select per field
                    const bool c0 = t2.rp.x > thr[0], c1 = t2.rp.y > thr[1];
                    const float r0 = c0 ? thr[0] : t2.rp.x, r1 = c1 ? thr[1] : t2.rp.y;
                    const bool take0 = r0 < 3.0e38f;
                    const float base = take0 ? r0 : 3.0e38f;
                    const bool take1 = n_here >= 2 && r1 < base;
                    const bool dead = take1 ? c1 : (c0 || !take0);      // the chosen term carries no gradient
                    m.rp = take1 ? r1 : base;
                    mi = take1 ? i0 + 1 : i0;
                    m.ca = dead ? 0.0f : (take1 ? t2.ca.y : t2.ca.x);
                    m.cb = dead ? 0.0f : (take1 ? t2.cb.y : t2.cb.x);
                    m.cc = dead ? 0.0f : (take1 ? t2.cc.y : t2.cc.x);
                    m.cl = dead ? 0.0f : (take1 ? t2.cl.y : t2.cl.x);
                    m.da = dead ? 0.0f : (take1 ? t2.da.y : t2.da.x);
                }
                if (VAR == PM_VAR_STATS) {
                } else if (MODE >= 2) {
                    S.aut[ch][p1] = (MODE == 2) ? m.rp : fminf(m.rp, S.aut[ch][p1]);
                } else {
                    const bool take = (MODE == 0) || m.rp < S.val[ch][p1];
                    if (take) {
                        S.val[ch][p1] = in ? m.rp : 3.0e38f;
                        S.src[ch][p1] = (signed char)mi;
                        S.coef[ch][p1] = make_float4(m.ca, m.cb, m.cc, m.cl);
                        if (VAR == PM_VAR_IMG) S.da[ch][p1] = m.da;
                    }
                }
            }
            yc = y1;
            xc = x1;
        }
    }
    if (VAR == PM_VAR_STATS) {
        // warp butterfly (fixed order), then this warp's own slot: no atomics, repeatable
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q == 1 && n_here < 2) continue;
            double a = (double)st[q][0], b = (double)st[q][1];
#pragma unroll
            for (int k = 16; k > 0; k >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, k);
                b += __shfl_xor_sync(0xffffffffu, b, k);
            }
            if (lane == 0) { S.stat[warp][map0 + q][0] += a; S.stat[warp][map0 + q][1] += b; }
        }
    }
}

// MODE0 = 0: the warped sources of a scale (maps map0 + i); MODE0 = 2: the raw sources (automask reference)
template <int MODE0, int VAR, int NSMAX, class SM>
__device__ __forceinline__ void pm_stage2_all(SM& S, int n_src, int map0, int tid, int tx0, int ty0, int H, int W,
                                              float C1, float C2, bool no_ssim) {
    // sources two at a time (they share the target's window sums); block-uniform control flow
    pm_stage2<MODE0, VAR>(S, 0, min(n_src, 2), map0, tid, tx0, ty0, H, W, C1, C2, no_ssim);
    if (NSMAX > 2 && n_src > 2) {
        __syncthreads();
        pm_stage2<MODE0 + 1, VAR>(S, NSMAX > 2 ? 1 : 0, n_src - 2, map0 + 2, tid, tx0, ty0, H, W, C1, C2, no_ssim);
    }
}

template <bool GRAD, int NSMAX, bool HEAD, int VAR>
__global__ void __launch_bounds__(PM_THREADS, 3)
photo_min_kernel(const __grid_constant__ PminLaunch p) {
    constexpr bool STATS = VAR == PM_VAR_STATS, IMG = VAR == PM_VAR_IMG;
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    char* ws = (char*)a.workspace;
    int32_t* tickets = (int32_t*)(ws + p.L.tickets);
    float* records = (float*)(ws + p.L.records);
    float* ws_pose = (float*)(ws + p.L.ws_pose);
    float* ws_loss = (float*)(ws + p.L.ws_loss);
    float* gup = (float*)(ws + p.L.gup);

    const int H = a.H, W = a.W, plane = H * W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int tx0 = (tile % p.tiles_x) * PM_TW, ty0 = (tile / p.tiles_x) * PM_TH;
    const plb_photo_job& job = a.jobs[0];
    const bool no_ssim = job.flags & PLB_PHOTO_NO_SSIM, automask = !(job.flags & PLB_PHOTO_NO_AUTOMASK);

    extern __shared__ __align__(16) unsigned char pm_smem[];
    typedef PminSmem<NSMAX, VAR> Smem;
    Smem& S = *reinterpret_cast<Smem*>(pm_smem);
    PairConst& pc = S.pc;

    // ---- prologue: K^-1, P per source, base pointers of image b ------------------------------------
    {
        const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
        const size_t img = (size_t)b * 3 * plane;
        if (warp == 0) {
            if (lane == 0) kinv_f32(Kb, a.k_is_f64, pc.kinv);
            if (lane == 1) {
                pc.tgt = job.tgt + img;
                pc.g_tgt = (GRAD && IMG && job.g_tgt) ? job.g_tgt + img : nullptr;
                pc.n_src = job.n_src; pc.n_scales = job.n_scales;
                pc.w_e = p.w_e * (a.upstream ? __ldg(a.upstream) : 1.0f);
            }
            if (lane >= 4 && lane < 4 + job.n_scales) {
                const int sc = lane - 4;
                const int dh = job.dh[sc], dw = job.dw[sc];
                pc.dh[sc] = dh; pc.dw[sc] = dw;
                pc.sx[sc] = (float)dw / (float)W; pc.sy[sc] = (float)dh / (float)H;
                pc.disp[sc] = job.disp[sc] + (size_t)b * dh * dw;
                float* g = nullptr;
                if (GRAD && job.g_disp[sc] != nullptr)
                    g = (dh != H || dw != W) ? gup + ((size_t)sc * a.B + b) * plane : job.g_disp[sc] + (size_t)b * plane;
                pc.g_disp[sc] = g;
            }
        } else if (warp == 1 && lane < job.n_src) {
            float M[12], P[12];
            pose_to_M(a.poses + ((size_t)b * a.n_pose + job.pose_index[lane]) * 6, a.rotation_mode, job.pose_inv[lane], M);
            k_times_M(Kb, a.k_is_f64, M, P);
            pc.P[lane][0] = make_float4(P[0], P[1], P[2], P[3]);
            pc.P[lane][1] = make_float4(P[4], P[5], P[6], P[7]);
            pc.P[lane][2] = make_float4(P[8], P[9], P[10], P[11]);
            pc.src[lane] = job.src[lane] + img;
            float* g = nullptr;
            if (GRAD && IMG && job.g_src[lane] != nullptr) {
                g = job.g_src[lane] + img;
                if (p.det)   // (a float* that holds the address of the image's int64 accumulators)
                    g = reinterpret_cast<float*>(reinterpret_cast<long long*>(ws + p.L.detacc) +
                                                 ((size_t)p.det_src[lane] * a.B * 3 * plane + img));
            }
            pc.g_src[lane] = g;
        } else if (warp == 2 && lane < PM_MAXMAPS) {
            S.thr[lane] = (!STATS && p.clip) ? __ldcg((const float*)(ws + p.L.pm_thr) + lane) : 3.0e38f;
        }
    }
    for (int k = tid; k < (PM_THREADS / 32) * (PH_NREC + 3); k += PM_THREADS) (&S.rec[0][0])[k] = 0.0f;
    if (STATS)
        for (int k = tid; k < (PM_THREADS / 32) * Smem::kStat * 2; k += PM_THREADS) (&S.stat[0][0][0])[k] = 0.0;
    __syncthreads();
    const int n_src = pc.n_src, n_scales = pc.n_scales;
    const float w_e = pc.w_e / (float)n_scales;        // scales are averaged
    // does the tile (+1 halo) touch the image border?  (reflection multiplicities in stage 3)
    const bool border = tx0 == 0 || ty0 == 0 || tx0 + PM_TW + 1 >= W || ty0 + PM_TH + 1 >= H;

    // ---- target tile (+2, reflected) and the automask reference min_i photo(src_i, tgt) -------------
    {
        // the image pointers live in shared memory (pc): read them ONCE into registers and issue every load of a pixel
        // before the first store - with the pointers re-read behind every shared-memory store (possible aliasing) the
        // loads of a pixel ran one after the other, each at the full latency of a global load
        const float* tg = pc.tgt;
        const float* sp[NSMAX];
#pragma unroll
        for (int i = 0; i < NSMAX; ++i) sp[i] = pc.src[i < n_src ? i : 0];
        for (int k = tid; k < PM_N2; k += PM_THREADS) {
            const int ly = k / PM_W2, lx = k - ly * PM_W2;
            const int gx = reflect_idx(tx0 + lx - 2, W), gy = reflect_idx(ty0 + ly - 2, H);
            const int o = gy * W + gx;
            float t[3], x[NSMAX][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) t[c] = __ldg(tg + (o + c * plane));
            if (automask) {
#pragma unroll
                for (int i = 0; i < NSMAX; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) x[i][c] = i < n_src ? __ldg(sp[i] + (o + c * plane)) : 0.0f;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) S.T2[c][k] = f2(t[c]);
            if (automask) {
#pragma unroll
                for (int g = 0; g < NSMAX / 2; ++g) {
                    if (2 * g >= n_src) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c) S.X2[g][c][k] = make_float2(x[2 * g][c], x[2 * g + 1][c]);
                }
            }
        }
    }
    __syncthreads();
    if (automask) pm_stage2_all<2, VAR, NSMAX>(S, n_src, PLB_MAX_SCALES * PLB_MAX_SRC, tid, tx0, ty0, H, W, p.C1, p.C2, no_ssim);
    __syncthreads();

    // pose sums of this thread's pixel over the scales: sum h_r = sum g_cam[r] * D and sum g_cam[r] per source (the
    // pixel's ray is the same at every scale: it multiplies the sums once, at the end)
    float acc[NSMAX][6];
    float lsum = 0.0f;
    float gta[3] = {0.0f, 0.0f, 0.0f};                 // IMG: d loss / d target of this thread's pixel, all scales
#pragma unroll
    for (int i = 0; i < NSMAX; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[i][k] = 0.0f;

    // this thread's tile pixel
    const int qx = tx0 + lane, qy = ty0 + warp;
    const bool qin = qx < W && qy < H;
    const int qk1 = (warp + 1) * PM_W1 + (lane + 1), qk2 = (warp + 2) * PM_W2 + (lane + 2);

#pragma unroll 1
    for (int s = 0; s < n_scales; ++s) {
        const bool full = pc.dh[s] == H && pc.dw[s] == W;
        // ---- (1) warp tile + 2 halo of every source into shared memory; the tile's own pixels also leave the
        //      derivatives of the sampler (stage 3 needs no second projection / gather) ---------------
        for (int k = tid; k < PM_N2; k += PM_THREADS) {
            const int ly = k / PM_W2, lx = k - ly * PM_W2;
            const int gx = reflect_idx(tx0 + lx - 2, W), gy = reflect_idx(ty0 + ly - 2, H);
            const float D = pm_depth<HEAD>(a, pc, s, full, gx, gy, W);
            const float xf = (float)gx, yf = (float)gy;
            const float rx = fmaf(pc.kinv[1], yf, pc.kinv[0] * xf) + pc.kinv[2];
            const float ry = fmaf(pc.kinv[4], yf, pc.kinv[3] * xf) + pc.kinv[5];
            const float rz = fmaf(pc.kinv[7], yf, pc.kinv[6] * xf) + pc.kinv[8];
            const bool own = GRAD && !STATS && lx >= 2 && lx < 2 + PM_TW && ly >= 2 && ly < 2 + PM_TH;
            const int q = own ? (ly - 2) * PM_TW + (lx - 2) : 0;
            if (own) S.sD[q] = D;
#pragma unroll
            for (int g = 0; g < NSMAX / 2; ++g) {
                if (2 * g >= n_src) continue;
                float w[2][3];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int i = 2 * g + hh;
                    w[hh][0] = w[hh][1] = w[hh][2] = 0.0f;
                    if (i < n_src) {
                        MinProj pq;
                        pm_project(pc.P[i][0], pc.P[i][1], pc.P[i][2], rx, ry, rz, D, H, W, pq);
                        float v[3][4];
                        pm_taps(pc.src[i], plane, W, pq, v);
                        const bool use = pq.mask != 0u;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float dA = v[c][1] - v[c][0], dB = v[c][3] - v[c][2];
                            const float top = fmaf(pq.fx, dA, v[c][0]), bot = fmaf(pq.fx, dB, v[c][2]);
                            const float dV = bot - top;
                            w[hh][c] = fmaf(pq.fy, dV, top);
                            if (own) { S.sdx[i][c][q] = fmaf(pq.fy, dB - dA, dA); S.sdy[i][c][q] = dV; }
                        }
                        if (own) {
                            // (the selects also keep the inf / NaN of a degenerate z out of the sums)
                            S.sgeo[i][0][q] = use ? pq.inv : 0.0f; S.sgeo[i][1][q] = use ? pq.px : 0.0f; S.sgeo[i][2][q] = use ? pq.py : 0.0f;
                            if (IMG) { S.sfx[i][q] = pq.fx; S.sfy[i][q] = pq.fy; S.so00[i][q] = pq.o00; S.smask[i][q] = pq.mask; }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) S.X2[g][c][k] = make_float2(w[0][c], w[1][c]);
            }
        }
        __syncthreads();
        // ---- (2) photometric mix and min over sources per channel (separable, sliding) ------------
        pm_stage2_all<0, VAR, NSMAX>(S, n_src, s * PLB_MAX_SRC, tid, tx0, ty0, H, W, p.C1, p.C2, no_ssim);
        __syncthreads();
        if (STATS) continue;                              // the statistics pass ends here
        // ---- (2b) automask, max over channels on tile + 1 ------------------------------------------
        for (int k1 = tid; k1 < PM_N1; k1 += PM_THREADS) {
            const int ly = k1 / PM_W1, lx = k1 - ly * PM_W1;
            float best = 0.0f;
            int bsel = -1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float rp = S.val[c][k1];
                if (rp < 1.0e38f) {                       // inside the image
                    const bool keep = !automask || rp < S.aut[c][k1];
                    const float val = keep ? rp : 0.0f;
                    // torch.max over channels: first maximal channel wins; a masked channel carries no gradient
                    if (bsel == -1 || val > best) {
                        best = val;
                        bsel = keep ? (c + 4 * (int)S.src[c][k1]) : -2;
                    }
                }
            }
            const bool interior = lx >= 1 && lx <= PM_TW && ly >= 1 && ly <= PM_TH;
            if (interior) lsum += best;
            S.sel[k1] = (signed char)(bsel >= 0 ? bsel : -1);
        }
        __syncthreads();
        // ---- (3) gradient of the tile pixels ------------------------------------------------------
        if (GRAD) {
            if (qin) {
                // d loss / d warped_i(q, c) = the selected terms of the 3x3 neighbours p; each neighbour selects ONE
                // (source, channel), so its term goes straight through that channel's sampler derivative into
                // (G_x, G_y) of that source
                float Gx[NSMAX], Gy[NSMAX];
                float e[IMG ? NSMAX : 1][3];
                float gt[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < NSMAX; ++i) Gx[i] = Gy[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < (IMG ? NSMAX : 1); ++i) e[i][0] = e[i][1] = e[i][2] = 0.0f;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int pk1 = qk1 + dy * PM_W1 + dx;
                        const int sel = S.sel[pk1];
                        const int sc = max(sel, 0), c = sc & 3, i = sc >> 2;
                        const float4 cf = S.coef[c][pk1];
                        const float xq = reinterpret_cast<const float*>(&S.X2[i >> 1][c][qk2])[i & 1], tq = S.T2[c][qk2].x;
                        float g = fmaf(cf.y, xq, fmaf(cf.z, tq, cf.x));
                        float gy_ = 0.0f;
                        if (IMG) gy_ = fmaf(cf.y, tq, fmaf(cf.z, xq, S.da[c][pk1]));
                        if (border) {
                            // multiplicity of q in p's reflection-padded window
                            const int pxg = qx + dx, pyg = qy + dy;
                            const float mx = 1.0f + ((pxg == 0 && dx == -1) ? 1.0f : 0.0f) + ((pxg == W - 1 && dx == 1) ? 1.0f : 0.0f);
                            const float my = 1.0f + ((pyg == 0 && dy == -1) ? 1.0f : 0.0f) + ((pyg == H - 1 && dy == 1) ? 1.0f : 0.0f);
                            g *= mx * my;
                            gy_ *= mx * my;
                        }
                        if (dx == 0 && dy == 0) { g += cf.w; gy_ -= cf.w; }
                        g = sel >= 0 ? g : 0.0f;
                        const float gdx = g * S.sdx[i][c][tid], gdy = g * S.sdy[i][c][tid];
#pragma unroll
                        for (int ii = 0; ii < NSMAX; ++ii) {
                            Gx[ii] += (i == ii) ? gdx : 0.0f;
                            Gy[ii] += (i == ii) ? gdy : 0.0f;
                        }
                        if (IMG) {
#pragma unroll
                            for (int ii = 0; ii < NSMAX; ++ii)
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) e[IMG ? ii : 0][cc] += (sel == cc + 4 * ii) ? g : 0.0f;
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) gt[cc] += (sel >= 0 && c == cc) ? gy_ : 0.0f;
                        }
                    }
                if (IMG) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) gta[cc] = fmaf(w_e, gt[cc], gta[cc]);
                }
                const float D = S.sD[tid];
                const float xf = (float)qx, yf = (float)qy;
                const float rx = fmaf(pc.kinv[1], yf, pc.kinv[0] * xf) + pc.kinv[2];
                const float ry = fmaf(pc.kinv[4], yf, pc.kinv[3] * xf) + pc.kinv[5];
                const float rz = fmaf(pc.kinv[7], yf, pc.kinv[6] * xf) + pc.kinv[8];
                float gD = 0.0f;
#pragma unroll
                for (int i = 0; i < NSMAX; ++i) {
                    if (i < n_src) {
                        const float4 Pa = pc.P[i][0], Pb = pc.P[i][1], Pc = pc.P[i][2];
                        const float Ax = fmaf(Pa.z, rz, fmaf(Pa.y, ry, Pa.x * rx));
                        const float Ay = fmaf(Pb.z, rz, fmaf(Pb.y, ry, Pb.x * rx));
                        const float Az = fmaf(Pc.z, rz, fmaf(Pc.y, ry, Pc.x * rx));
                        const float gi = w_e * S.sgeo[i][0][tid];
                        const float gcx = Gx[i] * gi, gcy = Gy[i] * gi;
                        const float gcz = -(gcx * S.sgeo[i][1][tid] + gcy * S.sgeo[i][2][tid]);
                        gD += fmaf(gcx, Ax, fmaf(gcy, Ay, gcz * Az));
                        acc[i][0] = fmaf(gcx, D, acc[i][0]); acc[i][1] = fmaf(gcy, D, acc[i][1]); acc[i][2] = fmaf(gcz, D, acc[i][2]);
                        acc[i][3] += gcx; acc[i][4] += gcy; acc[i][5] += gcz;
                        if (IMG && pc.g_src[i] != nullptr) {
                            // d loss / d source: e scattered through the bilinear weights (zero-padded taps get nothing)
                            const float fx = S.sfx[i][tid], fy = S.sfy[i][tid];
                            const int o00 = S.so00[i][tid];
                            const unsigned mask = S.smask[i][tid];
                            const float wnw = (1.0f - fx) * (1.0f - fy), wne = fx * (1.0f - fy);
                            const float wsw = (1.0f - fx) * fy, wse = fx * fy;
                            const float m = p.det ? PM_DET_ONE : w_e;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float ec = m * e[IMG ? i : 0][c];
                                if (ec == 0.0f) continue;
                                if (p.det) {
                                    unsigned long long* g = reinterpret_cast<unsigned long long*>(pc.g_src[i]) + (o00 + c * plane);
                                    if (mask & 1u) atomicAdd(g, (unsigned long long)__float2ll_rn(wnw * ec));
                                    if (mask & 2u) atomicAdd(g + 1, (unsigned long long)__float2ll_rn(wne * ec));
                                    if (mask & 4u) atomicAdd(g + W, (unsigned long long)__float2ll_rn(wsw * ec));
                                    if (mask & 8u) atomicAdd(g + W + 1, (unsigned long long)__float2ll_rn(wse * ec));
                                } else {
                                    float* g = pc.g_src[i] + (o00 + c * plane);
                                    if (mask & 1u) atomicAdd(g, wnw * ec);
                                    if (mask & 2u) atomicAdd(g + 1, wne * ec);
                                    if (mask & 4u) atomicAdd(g + W, wsw * ec);
                                    if (mask & 8u) atomicAdd(g + W + 1, wse * ec);
                                }
                            }
                        }
                    }
                }
                float* g = pc.g_disp[s];
                if (g != nullptr) {
                    float chain = (full && a.input_is_depth != PLB_INPUT_DEPTH) ? -a.disp_a * D * D : 1.0f;
                    if (full && HEAD) chain *= head_chain_from_depth(D, a.disp_a, a.disp_b, a.head_alpha, a.head_beta);
                    g[qy * W + qx] = gD * chain;
                }
            }
        }
        __syncthreads();   // X / selection arrays are rewritten by the next scale
    }

    if (STATS) {
        // ---- per-block (sum, sum of squares) of every map: fixed-order sum over the warps --------------
        double* out = (double*)(ws + p.L.pm_stat) + ((size_t)b * p.tiles + tile) * PM_MAXMAPS * 2;
        if (tid < PM_MAXMAPS * 2) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < PM_THREADS / 32; ++w) v += (&S.stat[w][0][0])[tid];
            out[tid] = v;
        }
        return;
    }
    if (GRAD && IMG && qin && pc.g_tgt != nullptr) {
        // every target pixel has exactly one owner: plain accumulation into the caller's (zeroed) buffer
        float* g = pc.g_tgt + (qy * W + qx);
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c * plane] += gta[c];
    }

    // ---- block record: warp butterflies, fixed-order sum over warps --------------------------------
    const float rxq = fmaf(pc.kinv[1], (float)qy, pc.kinv[0] * (float)qx) + pc.kinv[2];
    const float ryq = fmaf(pc.kinv[4], (float)qy, pc.kinv[3] * (float)qx) + pc.kinv[5];
    const float rzq = fmaf(pc.kinv[7], (float)qy, pc.kinv[6] * (float)qx) + pc.kinv[8];
#pragma unroll
    for (int i = 0; i < NSMAX; ++i) {
        float v[16];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            v[4 * r] = acc[i][r] * rxq; v[4 * r + 1] = acc[i][r] * ryq; v[4 * r + 2] = acc[i][r] * rzq; v[4 * r + 3] = acc[i][3 + r];
        }
        v[12] = (i == 0) ? lsum : 0.0f;
        v[13] = v[14] = v[15] = 0.0f;
        int which;
        const float r = warp_reduce16(v, lane, which);
        if ((lane & 1) == 0) {
            if (which < 12) S.rec[warp][i * 12 + which] = r;
            else if (which == 12 && i == 0) S.rec[warp][PLB_MAX_SRC * 12] = r;
        }
    }
    __syncthreads();
    float* my_rec = records + ((size_t)b * p.tiles + tile) * PH_REC_STRIDE;
    if (tid < PH_NREC) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < PM_THREADS / 32; ++w) v += S.rec[w][tid];
        __stcg(my_rec + tid, v);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) S.flag = (atomicAdd(&tickets[b], 1) == p.tiles - 1);
    __syncthreads();
    if (!S.flag) return;

    // ---- last tile of image b: sum its records (4 groups x 64 lanes, fixed order), pose chain -------
    __threadfence();
    {
        const int c = tid & 63, grp = tid >> 6;
        float v = 0.0f;
        if (c < PH_NREC) {
#pragma unroll 4
            for (int t = grp; t < p.tiles; t += 4) v += __ldcg(records + ((size_t)b * p.tiles + t) * PH_REC_STRIDE + c);
        }
        S.part4[grp][c] = v;
        __syncthreads();
        if (tid < PH_NREC) S.red[tid] = ((S.part4[0][tid] + S.part4[1][tid]) + S.part4[2][tid]) + S.part4[3][tid];
    }
    __syncthreads();
    if (tid == 0) { ws_loss[b] = S.red[PLB_MAX_SRC * 12]; tickets[b] = 0; }
    if (GRAD && tid < n_src) {
        float dP[12], dM[12], g6[6];
        for (int k = 0; k < 12; ++k) dP[k] = S.red[tid * 12 + k];
        const void* Kb = (const char*)a.K + (size_t)b * 9 * (a.k_is_f64 ? 8 : 4);
        kT_times_dP(Kb, a.k_is_f64, dP, dM);
        pose_to_M_vjp(a.poses + ((size_t)b * a.n_pose + job.pose_index[tid]) * 6, a.rotation_mode, job.pose_inv[tid], dM, g6);
        float* o = ws_pose + ((size_t)b * PLB_MAX_SRC + tid) * 6;
        for (int k = 0; k < 6; ++k) o[k] = g6[k];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) S.flag = (atomicAdd(&tickets[a.B], 1) == a.B - 1);
    __syncthreads();
    if (!S.flag) return;

    // ---- last image of the launch -----------------------------------------------------------------
    __threadfence();
    if (tid == 0) {
        double tot = 0.0;
        for (int q = 0; q < a.B; ++q) tot += (double)__ldcg(ws_loss + q);
        if (a.loss != nullptr) *a.loss = (float)(tot * (double)p.w_e / (double)n_scales);
        tickets[a.B] = 0;
    }
    if (GRAD && a.g_poses != nullptr) {
        for (int k = tid; k < a.B * a.n_pose * 6; k += PM_THREADS) {
            const int bb = k / (a.n_pose * 6), col = (k / 6) % a.n_pose, c = k % 6;
            float v = 0.0f;
            for (int i = 0; i < job.n_src; ++i)
                if (job.pose_index[i] == col) v += __ldcg(ws_pose + ((size_t)bb * PLB_MAX_SRC + i) * 6 + c);
            a.g_poses[k] = v;
        }
    }
}

// Clip thresholds: one block per map sums the per-block partials in block order (fp64) and writes
// thr = mean + clip * std (unbiased), rounded as the reference's float() of fp32 tensors rounds it (losses.py:80-82).
constexpr int PT_THREADS = 256;
__global__ void __launch_bounds__(PT_THREADS)
pmin_thresholds_kernel(const __grid_constant__ PminLaunch p, int n_blocks, double n_elems, float clip) {
    const plb_photo_args& a = p.a;
    if (skip_launch(a.skip_if_unit)) return;
    const int m = blockIdx.x, tid = threadIdx.x;
    const double* part = (const double*)((const char*)a.workspace + p.L.pm_stat);
    __shared__ double s_a[PT_THREADS], s_b[PT_THREADS];
    double sa = 0.0, sb = 0.0;
    for (int k = tid; k < n_blocks; k += PT_THREADS) {
        sa += __ldcg(part + ((size_t)k * PM_MAXMAPS + m) * 2);
        sb += __ldcg(part + ((size_t)k * PM_MAXMAPS + m) * 2 + 1);
    }
    s_a[tid] = sa; s_b[tid] = sb;
    __syncthreads();
    for (int k = PT_THREADS / 2; k > 0; k >>= 1) {
        if (tid < k) { s_a[tid] += s_a[tid + k]; s_b[tid] += s_b[tid + k]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double mean = s_a[0] / n_elems;
        const double var = n_elems > 1.0 ? fmax((s_b[0] - n_elems * mean * mean) / (n_elems - 1.0), 0.0) : 0.0;
        ((float*)((char*)a.workspace + p.L.pm_thr))[m] = (float)((float)mean + clip * (float)sqrt(var));
    }
}

template <bool GRAD, int NSMAX, bool HEAD, int VAR>
static int pm_launch_variant(const PminLaunch& p, dim3 grid, cudaStream_t st) {
    static bool attr_set[PLB_MAX_DEVICES] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {                                    // per device: a process may drive several GPUs
        const cudaError_t e = cudaFuncSetAttribute(photo_min_kernel<GRAD, NSMAX, HEAD, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)sizeof(PminSmem<NSMAX, VAR>));
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = true;
    }
    photo_min_kernel<GRAD, NSMAX, HEAD, VAR><<<grid, PM_THREADS, sizeof(PminSmem<NSMAX, VAR>), st>>>(p);
    return PLB_OK;
}

int photo_min_launch(const plb_photo_args* a, cudaStream_t st) {
    int rc = validate_photo(a);
    if (rc != PLB_OK) return rc;
    if (a->n_jobs != 1) return PLB_EINVAL;            // the dormant composition has one direction
    const plb_photo_job& job = a->jobs[0];
    bool img = a->want_grad && job.g_tgt != nullptr;
    for (int i = 0; i < job.n_src; ++i) img = img || (a->want_grad && job.g_src[i] != nullptr);
    const bool head = a->input_is_depth == PLB_INPUT_LOGIT;
    if (img && head) return PLB_EINVAL;               // image gradients: disparity / depth inputs only
    PminLaunch p;
    p.a = *a;
    p.L = photo_layout(*a);
    rc = photo_workspace_prepare(*a, p.L, st);
    if (rc != PLB_OK) return rc;
    p.tiles_x = (a->W + PM_TW - 1) / PM_TW;
    p.tiles = photo_min_tiles(*a);
    p.w_e = job.term_weight / ((float)a->B * (float)a->H * (float)a->W);
    p.C1 = 1e-4f; p.C2 = 9e-4f;
    p.clip = (job.flags & PLB_PHOTO_CLIP) ? 1 : 0;
    const PhotoDetSlots dslots = photo_det_slots(*a);
    p.det = (img && dslots.n > 0) ? 1 : 0;
    for (int i = 0; i < PLB_MAX_SRC; ++i) p.det_src[i] = dslots.src[0][i];
    dim3 grid(p.tiles, a->B);
#define PM_GO(G, N, V) (head ? pm_launch_variant<G, N, true, V>(p, grid, st) : pm_launch_variant<G, N, false, V>(p, grid, st))
    if (p.clip) {
        // phase A: the statistics of every map, then one threshold per map
        rc = job.n_src <= 2 ? PM_GO(false, 2, PM_VAR_STATS) : PM_GO(false, PLB_MAX_SRC, PM_VAR_STATS);
        if (rc != PLB_OK) return rc;
        ++g_launches;
        PLB_CHECK_LAUNCH();
        pmin_thresholds_kernel<<<PM_MAXMAPS, PT_THREADS, 0, st>>>(p, p.tiles * a->B, (double)a->B * 3.0 * a->H * a->W, job.clip_loss);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    if (img) rc = job.n_src <= 2 ? pm_launch_variant<true, 2, false, PM_VAR_IMG>(p, grid, st)
                                 : pm_launch_variant<true, PLB_MAX_SRC, false, PM_VAR_IMG>(p, grid, st);
    else if (a->want_grad) rc = job.n_src <= 2 ? PM_GO(true, 2, PM_VAR_MAIN) : PM_GO(true, PLB_MAX_SRC, PM_VAR_MAIN);
    else rc = job.n_src <= 2 ? PM_GO(false, 2, PM_VAR_MAIN) : PM_GO(false, PLB_MAX_SRC, PM_VAR_MAIN);
#undef PM_GO
    if (rc != PLB_OK) return rc;
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (p.det) {
        // one accumulator count = 2^-24 of (w_e / n_scales) x upstream (the upstream is applied by the conversion)
        rc = photo_det_convert_launch(*a, p.L, p.w_e / (float)job.n_scales / PM_DET_ONE, st);
        if (rc != PLB_OK) return rc;
    }
    if (photo_has_lowres_grad(*a)) {
        PhotoLaunch pl;
        pl.a = *a;
        pl.L = p.L;
        rc = photo_upsample_T_launch(pl, st);
        if (rc != PLB_OK) return rc;
    }
    return PLB_OK;
}

}  // namespace plb
