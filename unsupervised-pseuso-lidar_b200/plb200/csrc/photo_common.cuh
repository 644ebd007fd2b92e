// Declarations shared by the photometric kernels (photo.cu: live L1 mode; photo_min.cu: SSIM /
// min-reprojection / automask mode): workspace layout, launch constants, per-pair shared context.
#pragma once
#include "common.cuh"

namespace plb {

// Threads per block of the fused L1 kernel, per variant (tuned on B200, profiles/README.md): the
// single-scale <= 2-source kernel runs best with 96 registers and no spills (3 x 192 threads per SM);
// the multi-scale / many-source variants prefer 256-thread blocks.  >= 128: the prologue uses warps 0-3.
#ifndef PH_THREADS_SINGLE
#define PH_THREADS_SINGLE 192
#endif
#ifndef PH_THREADS_MULTI
#define PH_THREADS_MULTI 256
#endif
__host__ __device__ constexpr int photo_threads(int maxsrc, bool multi) {
    return (maxsrc <= 2 && !multi) ? PH_THREADS_SINGLE : PH_THREADS_MULTI;
}
constexpr int PH_MAX_WARPS = 8;
constexpr int PH_NREC = PLB_MAX_SRC * 12 + 1;  // per (job,image) record: dP[src][12], sum|diff|
constexpr int PH_REC_STRIDE = 64;  // floats per (block, set) record = two 128-byte lines: [0..PH_NREC) values, [PH_REC_ID] pair id
constexpr int PH_REC_ID = 63;
#ifndef PH_USE_PDL
#define PH_USE_PDL 1               // finalize kernel launched as a programmatic dependent of the main kernel
#endif
#ifndef PH_PF_SRC
#define PH_PF_SRC 1               // L1 prefetch of the source row this many rows below the current footprint (0 = off)
#endif
#ifndef PH_ROWPAIR
#define PH_ROWPAIR 1                // single-source directions of the 3-4-source kernels: two rows per iteration on the packed pipe
#endif
#ifndef PH_MIN_BLOCKS
#define PH_MIN_BLOCKS 3           // resident blocks per SM the <= 2-source kernels are compiled for
#endif

#ifndef PH_W_PAIR
#define PH_W_PAIR 8               // relative cost of a source pair (packed pipe) ...
#endif
#ifndef PH_W_ODD
#define PH_W_ODD 5                // ... and of a single source (scalar pipe) in the unit weights, <= 2-source kernels
#endif
#ifndef PH_W_SINGLE4
#define PH_W_SINGLE4 4             // a single-source job in the 3-4-source kernels (two rows per iteration on the packed pipe)
#endif
#ifndef PH_W_PAIR4
#define PH_W_PAIR4 4              // the same for the 3-4-source kernels (both <= 8)
#endif
#ifndef PH_W_ODD4
#define PH_W_ODD4 5                // (measured: the odd source of the 4-source variant costs MORE than a packed pair)
#endif

struct PhotoLayout {
    size_t tickets;   // int32 [n_pairs + 1]
    size_t records;   // float [grid][2][PH_REC_STRIDE]
    size_t lossrec;   // float2 [grid][2]: (pair id bits, sum |diff|) of every record again, compact: block 0 of the
                      // finalize kernel reads ALL of them, and a 224-byte stride costs it 32 L1 wavefronts per load
    size_t ws_pose;   // float [n_pairs][MAX_SRC][6] (photo_min.cu)
    size_t ws_loss;   // double [n_pairs]
    size_t gup;       // float [n_jobs][MAX_SCALES][B*H*W]  (only when a low scale carries a gradient)
    size_t total;
};

// Launch-time constants computed once on the host (kept out of the kernel's instruction stream).
struct PhotoLaunch {
    plb_photo_args a;
    PhotoLayout L;
    int grid;                            // number of blocks
    int warps_per_block;                 // photo_threads() / 32 of the launched variant
    int n_warps;                         // grid * warps_per_block
    int strips;                          // ceil(W / 32)
    int units_per_pair;                  // strips * H
    int n_pairs;                         // n_jobs * B
    int unit_weight[PLB_MAX_JOBS];       // n_scales * (PH_W_PAIR * pairs + PH_W_ODD * odd source)
    long long weight_start[PLB_MAX_JOBS + 1];  // cumulative weight at the start of each job
    int unit_start[PLB_MAX_JOBS + 1];    // cumulative unit index at the start of each job
    float w_e[PLB_MAX_JOBS];             // term_weight / (3*B*H*W)
    int lowres[PLB_MAX_JOBS];            // bit s set: scale s is not full resolution
    int share, share_rem;                // warp w starts at weight w*share + min(w, share_rem)
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline bool photo_has_lowres_grad(const plb_photo_args& a) {
    if (!a.want_grad) return false;
    for (int j = 0; j < a.n_jobs; ++j)
        for (int s = 0; s < a.jobs[j].n_scales; ++s)
            if (a.jobs[j].g_disp[s] && (a.jobs[j].dh[s] != a.H || a.jobs[j].dw[s] != a.W)) return true;
    return false;
}

static inline int photo_max_grid(const plb_photo_args& a) {
    // upper bound used for sizing the workspace (the launch may use fewer blocks)
    const long long strips = (a.W + 31) / 32;
    const long long units = strips * a.H * (long long)a.B * a.n_jobs;
    (void)units;
    long long g = 148LL * 8;                         // 148 SMs x at most 8 resident blocks
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
    L.total = off;
    return L;
}

// shared-memory context of one (job, image) pair, built once per block: K^-1, P per source and
// every base pointer already offset to image b, so the unit loop does no 64-bit address maths.
struct __align__(16) PairConst {
    float4 P[PLB_MAX_SRC][3];        // K . [R|t] (photo_min.cu)
    float4 Q[PLB_MAX_SRC][3];        // [P[:, :3] . K^-1 | P[:, 3]] (photo.cu: cam = D * Q.(x, y, 1) + p3)
    float4 Q2[PLB_MAX_SRC / 2][3][2]; // the same for source pairs, interleaved (even, odd) for packed fp32 maths
    float kinv[12];
    const float* tgt;
    float* g_tgt;
    const float* src[PLB_MAX_SRC];
    float* g_src[PLB_MAX_SRC];
    const float* disp[PLB_MAX_SCALES];
    float* g_disp[PLB_MAX_SCALES];   // full-res scales: the user's buffer; low-res scales: the gup scratch plane
    int dh[PLB_MAX_SCALES], dw[PLB_MAX_SCALES];
    float sx[PLB_MAX_SCALES], sy[PLB_MAX_SCALES];
    int n_src, n_scales, lowres, pad;
    float w_e, pad2[3];
};


// tile of the SSIM-mode kernel (photo_min.cu); needed here to size the record area
constexpr int PM_TW = 32, PM_TH = 8;
static inline int photo_min_tiles(const plb_photo_args& a) {
    return ((a.W + PM_TW - 1) / PM_TW) * ((a.H + PM_TH - 1) / PM_TH);
}

int photo_upsample_T_launch(const PhotoLaunch& p, cudaStream_t st);   // photo.cu
int photo_min_launch(const plb_photo_args* a, cudaStream_t st);       // photo_min.cu

}  // namespace plb
