// Declarations shared by the photometric kernels (photo.cu: live L1 mode; photo_min.cu: SSIM /
// min-reprojection / automask mode): workspace layout, launch constants, per-pair shared context.
#pragma once
#include "common.cuh"

namespace plb {

// Packed fp32 helpers: the two halves of a float2 ride Blackwell's packed pipe (FFMA2 / FADD2 / FMUL2), each half
// rounded exactly like the scalar operation.
template <int NS> struct Vec;
template <> struct Vec<1> { typedef float T; };
template <> struct Vec<2> { typedef float2 T; };

__device__ __forceinline__ float v_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 v_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float v_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float2 v_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float v_add(float a, float b) { return a + b; }
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float v_sub(float a, float b) { return a - b; }
__device__ __forceinline__ float2 v_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ void v_bc(float& v, float s) { v = s; }
__device__ __forceinline__ void v_bc(float2& v, float s) { v = make_float2(s, s); }
__device__ __forceinline__ float v_get(float v, int) { return v; }
__device__ __forceinline__ float v_get(float2 v, int k) { return k == 0 ? v.x : v.y; }
__device__ __forceinline__ void v_set(float& v, int, float s) { v = s; }
__device__ __forceinline__ void v_set(float2& v, int k, float s) { if (k == 0) v.x = s; else v.y = s; }
__device__ __forceinline__ float v_hsum(float v) { return v; }
__device__ __forceinline__ float v_hsum(float2 v) { return v.x + v.y; }


// Threads per block of the fused L1 kernel (tuned on B200, profiles/README.md): 4 x 160 threads per SM = 5 warps per
// scheduler, the most that 96 registers per thread allow (a scheduler's quarter of the register file holds
// 16384 / (32 x 96) = 5.3 warps; 3 x 192 leaves two schedulers a warp short).
#ifndef PH_THREADS
#define PH_THREADS 160
#endif
#ifndef PH_MIN_BLOCKS
#define PH_MIN_BLOCKS 4           // resident blocks per SM the kernel is compiled for
#endif
constexpr int PH_WARPS = PH_THREADS / 32;
constexpr int PH_NREC = PLB_MAX_SRC * 12 + 1;  // per (job,image) record: dP[src][12], sum|diff|
constexpr int PH_REC_STRIDE = 64;  // floats per (block, set) record = two 128-byte lines: [0..PH_NREC) values, [PH_REC_ID] pair id
constexpr int PH_REC_ID = 63;
#ifndef PH_USE_PDL
#define PH_USE_PDL 1               // finalize kernel launched as a programmatic dependent of the main kernel
#endif
#ifndef PH_PF_SRC
#define PH_PF_SRC 1               // L1 prefetch of the source row this many rows below the current footprint (0 = off)
#endif
// cost model of one row segment of a combo (the unit list is cut into equal WEIGHT shares, one per warp)
#ifndef PH_W_PAIR
#define PH_W_PAIR 27              // a packed combo (two (source, scale) samples on the packed fp32 pipe) ...
#endif
#ifndef PH_W_ODD
#define PH_W_ODD 19               // ... a single sample on the scalar pipe ...
#endif
#ifndef PH_W_LOW
#define PH_W_LOW 3                // ... and what every low-resolution depth stream of the combo adds (pre- and post-pass of a chunk)
#endif
#ifndef PH_W_SPAIR
#define PH_W_SPAIR 14              // ... and a scale pair (two gradient maps instead of one)
#endif
#ifndef PH_W_SPAIR2
#define PH_W_SPAIR2 8              // ... the same when no job has more than two sources (c2 187 -> 180 us, c5 950 -> 891 us)
#endif
#ifndef PH_GRID_MULT
#define PH_GRID_MULT 1            // blocks launched per resident block slot
#endif
#ifndef PH_ROW_ALIGN
#define PH_ROW_ALIGN 16           // rows of a chunk (shared-memory staging of a low-resolution stream): >= 2 x the largest LOWFAST factor
#endif

// A COMBO is what one warp evaluates per target pixel: two (source, scale) samples on the two halves of the packed
// fp32 pipe - two sources at one scale, or one source at two scales (same pixel ray, two depths) - or, when a job
// has an odd number of samples, one sample on the scalar pipe.
constexpr int PH_MAX_COMBOS = PLB_MAX_SRC * PLB_MAX_SCALES / 2 + 1;
enum { PH_KIND_SRCPAIR = 0, PH_KIND_SCALEPAIR = 1, PH_KIND_SINGLE = 2 };
struct PhotoCombo {
    unsigned char kind, k0, s0, s1;   // sources k0 (and k0 + 1 for a source pair), scales s0 (and s1 for a scale pair)
    unsigned char c0, c1, pad[2];     // which contributor (0 / 1) of scale s0 / s1 this combo is: with more than two sources
                                      // two combos of a job feed the gradient of one scale
};

// How a scale of a job's pyramid is read and how its gradient leaves the main kernel.
enum {
    PH_SM_FULL = 0,        // full resolution: one disparity per pixel, gradient written in place
    PH_SM_LOWFAST = 1,     // low resolution, integer factor 2/4/8: depth upsampled on the fly (streaming rows), the gradient
                           // reduced over the rows in registers and left as [2][dh][W] partial rows for photo_lowres_merge_kernel
    PH_SM_LOWSCRATCH = 2,  // any other ratio: per-pixel gradient into a full-resolution scratch plane + photo_upsample_T_kernel
    PH_SM_LOWNOGRAD = 3    // low resolution, no gradient wanted
};

// maps of the SSIM-mode composition that carry their own clip threshold: (scale, source) and the automask references
constexpr int PM_MAXMAPS = PLB_MAX_SCALES * PLB_MAX_SRC + PLB_MAX_SRC;
constexpr size_t PH_PAIRCONST_BYTES = 2048;   // >= sizeof(PairConst) (static_assert below)
struct PhotoLayout {
    size_t tickets;   // int32 [n_pairs + 1]
    size_t records;   // float [grid][2][PH_REC_STRIDE]
    size_t lossrec;   // float2 [grid][2]: (pair id bits, sum |diff|) of every record again, compact: block 0 of the
                      // finalize kernel reads ALL of them, and a 224-byte stride costs it 32 L1 wavefronts per load
    size_t ws_pose;   // float [n_pairs][MAX_SRC][6] (photo_min.cu)
    size_t ws_loss;   // double [n_pairs]
    size_t gup;       // float [n_jobs][MAX_SCALES][B*H*W]  (only when a scale is PH_SM_LOWSCRATCH)
    size_t pairs;     // PairConst [n_pairs] (photo_pairs_kernel -> photo_l1_kernel)
    size_t ylow;      // float: per PH_SM_LOWFAST (job, scale) [B][2][dh][W] y-reduced gradient rows (two partial slots)
    size_t ylow_off[PLB_MAX_JOBS][PLB_MAX_SCALES];   // float offset of each (job, scale) from `ylow`
    size_t pm_thr;    // float [PM_MAXMAPS]: clip thresholds of the SSIM-mode maps (photo_min.cu, PLB_PHOTO_CLIP)
    size_t pm_stat;   // double [tiles * B][PM_MAXMAPS][2]: per-block (sum, sum of squares) of every map
    size_t detacc;    // int64 [det_n][B*3*H*W]: fixed-point accumulators of the image gradients (deterministic mode)
    int det_n;        // distinct image-gradient buffers of the call (0: not deterministic / no image gradients)
    size_t total;
};

// Deterministic image gradients.  The bilinear scatter of d loss / d source (and the per-combo sums into
// d loss / d target) are the only results that several warps add into in an order the hardware picks.  In
// deterministic mode every contribution is rounded ONCE to a 2^-30 fixed-point fraction of the largest per-pixel
// weight of the call and added with 64-bit INTEGER atomics - integer addition is associative, so the sums are
// bitwise repeatable whatever the order - and photo_det_convert_kernel turns the accumulators into floats (and
// re-zeroes them).  |contribution| <= 2^30 and at most B*H*W*samples < 2^32 of them meet in one pixel: no overflow.
constexpr float PH_DET_ONE = 1073741824.0f;           // 2^30
constexpr int PH_DET_MAX = PLB_MAX_SRC + 2;           // distinct image-gradient buffers of a call (sources + targets)
struct PhotoDetSlots {
    int n;
    float* out[PH_DET_MAX];                           // the caller's buffers, in first-seen order
    int src[PLB_MAX_JOBS][PLB_MAX_SRC], tgt[PLB_MAX_JOBS];   // slot of every job's g_src[i] / g_tgt, -1 = not wanted
};
static inline PhotoDetSlots photo_det_slots(const plb_photo_args& a) {
    PhotoDetSlots d;
    d.n = 0;
    for (int k = 0; k < PH_DET_MAX; ++k) d.out[k] = nullptr;
    auto slot = [&](float* g) -> int {
        if (g == nullptr) return -1;
        for (int k = 0; k < d.n; ++k) if (d.out[k] == g) return k;
        if (d.n >= PH_DET_MAX) return -1;
        d.out[d.n] = g;
        return d.n++;
    };
    for (int j = 0; j < PLB_MAX_JOBS; ++j) {
        d.tgt[j] = -1;
        for (int i = 0; i < PLB_MAX_SRC; ++i) d.src[j][i] = -1;
        if (j >= a.n_jobs || !a.want_grad || !a.deterministic) continue;
        // (the SSIM-mode kernel owns every target pixel: its target gradient is written directly, no accumulator)
        if (a.jobs[j].mode != PLB_PHOTO_MIN_REPROJ) d.tgt[j] = slot(a.jobs[j].g_tgt);
        for (int i = 0; i < a.jobs[j].n_src && i < PLB_MAX_SRC; ++i) d.src[j][i] = slot(a.jobs[j].g_src[i]);
    }
    return d;
}

// Launch-time constants computed once on the host (kept out of the kernel's instruction stream).
struct PhotoLaunch {
    plb_photo_args a;
    PhotoLayout L;
    int grid;                            // number of blocks
    int perm_sms, perm_bps;              // > 0: block i owns share (i % perm_sms) * perm_bps + i / perm_sms (grid == perm_sms * perm_bps)
    int warps_per_block;                 // PH_WARPS
    int n_warps;                         // grid * warps_per_block
    int strips;                          // ceil(W / 32)
    int n_pairs;                         // n_jobs * B
    int n_combos[PLB_MAX_JOBS];
    PhotoCombo combo[PLB_MAX_JOBS][PH_MAX_COMBOS];
    unsigned char smode[PLB_MAX_JOBS][PLB_MAX_SCALES];   // PH_SM_*, | 4 when two combos of the job feed the scale's gradient
    int units_per_pair[PLB_MAX_JOBS];    // strips * n_combos * H: units (rows of one combo of one strip) of one image of the job
    int combo_w[PLB_MAX_JOBS][PH_MAX_COMBOS];       // weight of one row segment of each combo
    int combo_align[PLB_MAX_JOBS][PH_MAX_COMBOS];   // cuts inside a column of the combo fall on multiples of this many rows:
                                                    // 2 x the largest factor of its PH_SM_LOWFAST streams (1: anywhere)
    int combo_cw[PLB_MAX_JOBS][PH_MAX_COMBOS + 1];  // prefix sums of combo_w: a (strip, all combos) super column weighs combo_cw[n] * H
    long long weight_start[PLB_MAX_JOBS + 1];  // cumulative weight at the start of each job
    int unit_start[PLB_MAX_JOBS + 1];    // cumulative unit index at the start of each job
    float w_e[PLB_MAX_JOBS];             // term_weight / (3*B*H*W)
    int det;                             // deterministic image gradients (fixed-point accumulators, see PhotoLayout::detacc)
    int det_src[PLB_MAX_JOBS][PLB_MAX_SRC], det_tgt[PLB_MAX_JOBS];   // accumulator slot of g_src[i] / g_tgt, -1 = none
    float det_rho[PLB_MAX_JOBS];         // w_e[j] / max_j w_e[j]: a contribution of job j in units of the largest weight
    int lowres[PLB_MAX_JOBS];            // bit s set: scale s is not full resolution
    int share, share_rem;                // warp w starts at weight w*share + min(w, share_rem)
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// PH_SM_* of scale s of job j (host).  The in-register path needs an integer power-of-two factor <= 8 in both
// directions; the SSIM / min-reprojection kernel (photo_min.cu) always uses the scratch planes.
static inline int photo_scale_mode(const plb_photo_args& a, int j, int s) {
    const plb_photo_job& job = a.jobs[j];
    if (job.dh[s] == a.H && job.dw[s] == a.W) return PH_SM_FULL;
    if (!a.want_grad || job.g_disp[s] == nullptr) return PH_SM_LOWNOGRAD;
    if (job.mode == PLB_PHOTO_MIN_REPROJ) return PH_SM_LOWSCRATCH;
    if (job.dh[s] < 1 || job.dw[s] < 1 || a.H % job.dh[s] != 0 || a.W % job.dw[s] != 0) return PH_SM_LOWSCRATCH;
    const int f = a.H / job.dh[s];
    if (f != a.W / job.dw[s] || (f != 2 && f != 4 && f != 8)) return PH_SM_LOWSCRATCH;
#ifdef PH_NO_LOWFAST
    return PH_SM_LOWSCRATCH;
#endif
    return PH_SM_LOWFAST;
}

// a low scale whose gradient goes through the scratch planes + photo_upsample_T_kernel
static inline bool photo_has_lowres_grad(const plb_photo_args& a) {
    for (int j = 0; j < a.n_jobs; ++j)
        for (int s = 0; s < a.jobs[j].n_scales; ++s)
            if (photo_scale_mode(a, j, s) == PH_SM_LOWSCRATCH) return true;
    return false;
}

static inline int photo_max_grid(const plb_photo_args& a) {
    // upper bound used for sizing the workspace (the launch may use fewer blocks)
    const long long strips = (a.W + 31) / 32;
    const long long units = strips * a.H * (long long)a.B * a.n_jobs;
    (void)units;
    long long g = 160LL * 8 * PH_GRID_MULT;          // SMs x at most 8 resident blocks (B200: 148 SMs) x waves
    const long long pairs_bound = 1LL * a.n_jobs * a.B * PLB_MAX_SCALES * PLB_MAX_SRC + 2;
    if (g < pairs_bound) g = pairs_bound;
    return (int)g;
}

static inline PhotoLayout photo_layout(const plb_photo_args& a) {
    PhotoLayout L;
    const size_t n_pairs = (size_t)a.n_jobs * a.B;
    size_t off = 0;
    L.tickets = off; off = align_up(off + sizeof(int32_t) * (n_pairs + 1), 256);
    size_t n_rec = (size_t)photo_max_grid(a) * 2;
    if (a.n_jobs >= 1 && a.jobs[0].mode == PLB_PHOTO_MIN_REPROJ) {
        const size_t tiles = (size_t)((a.W + 31) / 32) * ((a.H + 7) / 8);   // photo_min.cu tile 32x8
        if (tiles * a.B > n_rec) n_rec = tiles * a.B;
    }
    L.records = off; off = align_up(off + sizeof(float) * n_rec * PH_REC_STRIDE, 256);
    L.lossrec = off; off = align_up(off + sizeof(float) * 2 * n_rec, 256);
    L.ws_pose = off; off = align_up(off + sizeof(float) * n_pairs * PLB_MAX_SRC * 6, 256);
    L.ws_loss = off; off = align_up(off + sizeof(double) * n_pairs, 256);
    L.gup = off;
    if (photo_has_lowres_grad(a))
        off = align_up(off + sizeof(float) * (size_t)a.n_jobs * PLB_MAX_SCALES * a.B * a.H * a.W, 256);
    L.pairs = off; off = align_up(off + PH_PAIRCONST_BYTES * n_pairs, 256);
    L.ylow = off;
    size_t yl = 0;
    for (int j = 0; j < PLB_MAX_JOBS; ++j)
        for (int s = 0; s < PLB_MAX_SCALES; ++s) {
            L.ylow_off[j][s] = yl;
            if (j < a.n_jobs && s < a.jobs[j].n_scales && photo_scale_mode(a, j, s) == PH_SM_LOWFAST) {
                const int n_contrib = a.jobs[j].n_src / 2 + (a.jobs[j].n_src & 1);   // combos that feed one scale (<= 2)
                yl += align_up((size_t)n_contrib * a.B * 2 * a.jobs[j].dh[s] * a.W, 64);
            }
        }
    off = align_up(off + sizeof(float) * yl, 256);
    L.pm_thr = off; L.pm_stat = off;
    if (a.n_jobs >= 1 && a.jobs[0].mode == PLB_PHOTO_MIN_REPROJ && (a.jobs[0].flags & PLB_PHOTO_CLIP)) {
        off = align_up(off + sizeof(float) * PM_MAXMAPS, 256);
        L.pm_stat = off;
        off = align_up(off + sizeof(double) * (size_t)((a.W + 31) / 32) * ((a.H + 7) / 8) * a.B * PM_MAXMAPS * 2, 256);
    }
    L.detacc = off;
    L.det_n = photo_det_slots(a).n;
    off = align_up(off + sizeof(long long) * (size_t)L.det_n * a.B * 3 * a.H * a.W, 256);
    L.total = off;
    return L;
}

// shared-memory context of one (job, image) pair, built once per block: K^-1, P per source and
// every base pointer already offset to image b, so the unit loop does no 64-bit address maths.
constexpr int PH_NT2 = PLB_MAX_SRC / 2 + PLB_MAX_SRC;   // packed constant tables: source pairs, then every source twice
struct __align__(16) PairConst {
    float4 P[PLB_MAX_SRC][3];        // K . [R|t] (photo_min.cu)
    float4 Q[PLB_MAX_SRC][3];        // [P[:, :3] . K^-1 | P[:, 3]] (photo.cu: cam = D * Q.(x, y, 1) + p3)
    // the same for the two halves of a packed combo, per cam row r: [0] = (qx_lo, qx_hi, qz_lo, qz_hi),
    // [1] = (qy_lo, qy_hi, p3_lo, p3_hi).  Table g < MAX_SRC/2: sources (2g, 2g+1); table MAX_SRC/2 + k: source k twice
    float4 T2[PH_NT2][3][2];
    float kinv[12];
    const float* tgt;
    float* g_tgt;
    const float* src[PLB_MAX_SRC];
    float* g_src[PLB_MAX_SRC];
    const float* disp[PLB_MAX_SCALES];
    float* g_disp[PLB_MAX_SCALES];   // FULL: the user's buffer; LOWSCRATCH: the gup scratch plane; LOWFAST: the [2][dh][W] partial rows
    int dh[PLB_MAX_SCALES], dw[PLB_MAX_SCALES];
    float sx[PLB_MAX_SCALES], sy[PLB_MAX_SCALES];
    int smode[PLB_MAX_SCALES];
    int fac[PLB_MAX_SCALES];         // integer power-of-two upsampling factor of the scale (both directions), else 0
    int n_src, n_scales, lowres, pad;
    float w_e, det_rho, pad2[2];     // det_rho > 0: g_src / g_tgt point at int64 fixed-point accumulators (deterministic mode)
};


static_assert(sizeof(PairConst) <= PH_PAIRCONST_BYTES && sizeof(PairConst) % 16 == 0, "pair table entry");

// tile of the SSIM-mode kernel (photo_min.cu); needed here to size the record area
constexpr int PM_TW = 32, PM_TH = 8;
static inline int photo_min_tiles(const plb_photo_args& a) {
    return ((a.W + PM_TW - 1) / PM_TW) * ((a.H + PM_TH - 1) / PM_TH);
}

int photo_upsample_T_launch(const PhotoLaunch& p, cudaStream_t st);   // photo.cu
// fixed-point accumulators (PhotoLayout::detacc) -> the caller's buffers; `unit`: value of one accumulator count
int photo_det_convert_launch(const plb_photo_args& a, const PhotoLayout& L, float unit, cudaStream_t st);   // photo.cu
int photo_min_launch(const plb_photo_args* a, cudaStream_t st);       // photo_min.cu
// clears the workspace when the layout it was last used with differs from this call's (photo.cu)
int photo_workspace_prepare(const plb_photo_args& a, const PhotoLayout& L, cudaStream_t st);

}  // namespace plb
