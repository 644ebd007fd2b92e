/*
 * plb200.h - C ABI of libplb200.so: the B200 (sm_100a) implementation of the
 * view-synthesis hot path of unsupervised-pseuso-LiDAR.
 *
 * The reference is pure Python over stock torch ops and has no FFI of its own
 * (SURVEY.md section 0.1); the "interface each entry point replaces" is
 * therefore the Python function named beside it (paths relative to the
 * reference tree).  Every function
 *   - takes plain device pointers, sizes and a cudaStream_t (passed as void*),
 *   - never allocates, never synchronises, and is CUDA-graph capturable,
 *   - returns 0 on success, a negative PLB_E* code for a bad argument, or a
 *     positive cudaError_t if the launch failed.
 * All image tensors are contiguous NCHW fp32 as `trainer.py:291-299` produces
 * them; intrinsics may be fp64 (`dataloaders.py:98`) or fp32.
 */
#ifndef PLB200_H
#define PLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_MAX_SRC 4
#define PLB_MAX_SCALES 4
#define PLB_MAX_JOBS 2

#define PLB_OK 0
#define PLB_EINVAL (-1)     /* bad shape / count / flag                     */
#define PLB_ENULL (-2)      /* a required pointer is NULL                    */
#define PLB_EWORKSPACE (-3) /* workspace missing or too small                */

/* rotation_mode */
#define PLB_ROT_AXISANGLE 0 /* live: transformation_from_parameters, geometry/pose_geometry.py:124-199 */
#define PLB_ROT_EULER 1     /* dormant: pose_vec2mat / euler2mat, geometry/pose_geometry.py:38-108 */
#define PLB_INPUT_DISP 0
#define PLB_INPUT_DEPTH 1
#define PLB_INPUT_LOGIT 2

/* plb_photo_job.mode */
#define PLB_PHOTO_L1_MEAN 0 /* live: nn.L1Loss per (scale, source), mean over sources: losses.py:223-228 */
#define PLB_PHOTO_MIN_REPROJ 1 /* dormant: 0.85*SSIM+0.15*L1, per-pixel min over sources, automask,
                                  max over channel: losses.py:12-84,94-96,154-162,
                                  notes/toy_problem/losses.py:107-129 */
/* plb_photo_job.flags (MIN_REPROJ mode) */
#define PLB_PHOTO_NO_SSIM 1u
#define PLB_PHOTO_NO_AUTOMASK 2u
#define PLB_PHOTO_CLIP 4u  /* clamp every photometric map (each warped source at each scale, each automask
                              reference) at mean + clip_loss * std of that whole [B,3,H,W] map, the threshold
                              detached, as compute_photometric_loss does (losses.py:79-82); two more launches
                              (map statistics, thresholds) precede the fused kernel */

/*
 * One "direction" of Losses.reprojection_loss (losses.py:190-228): a target
 * frame, n_src source frames warped into it with one depth pyramid.
 */
typedef struct plb_photo_job {
    const float* tgt;                   /* [B,3,H,W] target frame                                  */
    const float* src[PLB_MAX_SRC];      /* [B,3,H,W] source frames                                 */
    int32_t pose_index[PLB_MAX_SRC];    /* column of `poses` [B,n_pose,6] used for source i        */
    int32_t pose_inv[PLB_MAX_SRC];      /* 1: invert_pose() the matrix (losses.py:202)             */
    const float* disp[PLB_MAX_SCALES];  /* [B,1,dh,dw] disparity (or depth) pyramid of the target  */
    int32_t dh[PLB_MAX_SCALES];
    int32_t dw[PLB_MAX_SCALES];
    float* g_disp[PLB_MAX_SCALES];      /* out (written): d loss / d disp[s]; NULL = not wanted    */
    float* g_src[PLB_MAX_SRC];          /* out (ACCUMULATED, caller zeroes): d loss / d src[i]     */
    float* g_tgt;                       /* out (ACCUMULATED, caller zeroes): d loss / d tgt        */
    int32_t n_src;
    int32_t n_scales;
    float term_weight;                  /* weight of each (scale, source) mean in the final loss   */
    int32_t mode;                       /* PLB_PHOTO_*                                             */
    uint32_t flags;
    float clip_loss;                    /* PLB_PHOTO_CLIP: 0.5 in the reference (losses.py:58)     */
} plb_photo_job;

/*
 * Fused photometric reprojection loss, forward (+ gradients in the same pass).
 * Replaces Losses.reprojection_loss (losses.py:183-240) together with everything
 * it calls: disp_to_depth (geometry/pose_geometry.py:70-95), F.interpolate of
 * the low scales (losses.py:214-215), inverse_warp (geometry/pose_geometry.py:
 * 201-228), Transform.reconstruct/project (geometry/transform.py:74-150),
 * transformation_from_parameters / invert_pose (geometry/pose_geometry.py:
 * 110-199), F.grid_sample and nn.L1Loss - and their autograd backward.
 */
typedef struct plb_photo_args {
    int32_t B, H, W;
    int32_t n_jobs;
    int32_t n_pose;            /* poses is [B,n_pose,6] (rot3 | trans3)                           */
    int32_t rotation_mode;     /* PLB_ROT_*                                                        */
    int32_t k_is_f64;          /* intrinsics dtype                                                 */
    int32_t input_is_depth;    /* PLB_INPUT_DISP 0: depth = 1/(disp_a*disp+disp_b); PLB_INPUT_DEPTH 1: `disp` already holds
                                  depth; PLB_INPUT_LOGIT 2: `disp` holds the disparity head's pre-activation x and
                                  disp = head_alpha*sigmoid(x)+head_beta is folded in too (models/depth/disp_net.py:121-139;
                                  gradients are then with respect to x)                             */
    float disp_a, disp_b;      /* 10, 0.01 in the reference                                        */
    float head_alpha, head_beta; /* 10, 0.01 in the reference's DispNet (PLB_INPUT_LOGIT only)      */
    int32_t want_grad;         /* 0: loss only (no_grad / eval)                                    */
    int32_t deterministic;     /* loss, pose and disparity gradients are always bitwise repeatable.  The image gradients
                                  (g_src / g_tgt: the scatter that replaces grid_sampler_2d_backward behind
                                  geometry/pose_geometry.py:227) use float atomics when 0; when 1 every contribution is
                                  rounded once to a 2^-30 fixed-point fraction of the largest per-pixel weight and summed
                                  with 64-bit integer atomics (order independent => bitwise repeatable), then converted;
                                  costs 8 B of workspace per image-gradient element                  */
    int32_t sm_limit;          /* 0: the persistent grid fills every SM; n > 0: it is sized for n SMs, leaving the rest
                                  to kernels that run beside it (the CTAs of an NCCL all-reduce need whole SMs)      */
    int32_t reserved2;
    const float* poses;        /* [B,n_pose,6]                                                     */
    const void* K;             /* [B,3,3] f64 or f32                                               */
    float* g_poses;            /* out (written) [B,n_pose,6]; NULL = not wanted                    */
    float* loss;               /* out (written) [1]                                                */
    const float* upstream;     /* device scalar d L / d loss; NULL = 1                             */
    const float* skip_if_unit[2]; /* device scalars or NULL: when at least one is given and every
                                  given one equals 1, the launch returns at once (the gradients
                                  written by the forward pass with unit upstream are then already
                                  exact) - see DESIGN.md                                          */
    void* workspace;           /* plb_photo_workspace_bytes() bytes, zero-filled ONCE by the caller; it cleans up
                                  after itself, and the library clears it when a later call brings a different
                                  layout (another batch / image size / mode / set of gradients)         */
    size_t workspace_bytes;
    plb_photo_job jobs[PLB_MAX_JOBS];
} plb_photo_args;

size_t plb_photo_workspace_bytes(const plb_photo_args* args);
int plb_photo_loss(const plb_photo_args* args, void* stream);

/*
 * Second-order depth smoothness, forward + gradient in one pass.
 * Replaces Losses.smooth_loss (losses.py:242-260) on the target-frame depth
 * pyramid, with disp_to_depth folded in.
 */
typedef struct plb_smooth_args {
    int32_t B;
    int32_t n_scales;
    const float* disp[PLB_MAX_SCALES];  /* [B,1,dh,dw]                                            */
    int32_t dh[PLB_MAX_SCALES];
    int32_t dw[PLB_MAX_SCALES];
    float* g_disp[PLB_MAX_SCALES];      /* out; NULL = not wanted                                  */
    int32_t accumulate;                 /* 1: g_disp += grad, 0: g_disp = grad                      */
    int32_t input_is_depth;             /* PLB_INPUT_* as in plb_photo_args                         */
    float disp_a, disp_b;
    float head_alpha, head_beta;
    float scale_decay;                  /* 2.3 in the reference (losses.py:259)                     */
    int32_t want_grad;
    float* loss;                        /* out (written) [1]                                        */
    const float* upstream;              /* device scalar or NULL (=1)                               */
    const float* skip_if_unit[2];       /* as in plb_photo_args                                     */
    void* workspace;                    /* plb_smooth_workspace_bytes() bytes, zero-filled once     */
    size_t workspace_bytes;
} plb_smooth_args;

size_t plb_smooth_workspace_bytes(const plb_smooth_args* args);
int plb_smooth_loss(const plb_smooth_args* args, void* stream);

/*
 * Edge-aware first-order disparity smoothness, forward + gradient.
 * NOT IN THE REFERENCE (its live smoothness is plb_smooth_loss); north_star asks for it.  Formula of the
 * monodepth2 lineage the reference's model files cite (models/depth/layers.py:1-2):
 *   sum_s (1/f_s) [ mean(|dx d'| exp(-mean_c |dx I_s|)) + mean(|dy d'| exp(-mean_c |dy I_s|)) ],
 *   d' = d / (mean_hw(d) + 1e-7) when `normalize`, I_s = avg_pool2d(tgt, f_s), f_s = H / dh[s].
 * Parity unpinned: checked against oracle/restated.py::edge_aware_smooth_loss only.
 */
typedef struct plb_edge_args {
    int32_t B, H, W;
    int32_t n_scales;
    const float* tgt;                   /* [B,3,H,W] target image                                  */
    const float* disp[PLB_MAX_SCALES];  /* [B,1,dh,dw] disparity pyramid; H % dh == 0, dw == W / (H/dh) */
    int32_t dh[PLB_MAX_SCALES];
    int32_t dw[PLB_MAX_SCALES];
    float* g_disp[PLB_MAX_SCALES];      /* out; NULL = not wanted                                   */
    float* g_scratch[PLB_MAX_SCALES];   /* [B,1,dh,dw] scratch, required when normalize && g_disp   */
    int32_t accumulate;                 /* 1: g_disp += grad, 0: g_disp = grad                      */
    int32_t normalize;
    int32_t want_grad;
    int32_t reserved;
    float* loss;                        /* out (written) [1]                                        */
    const float* upstream;              /* device scalar or NULL (=1)                               */
    void* workspace;                    /* plb_edge_smooth_workspace_bytes() bytes                  */
    size_t workspace_bytes;
    const float* skip_if_unit[2];       /* as in plb_photo_args: the fused step relaunches behind this guard */
} plb_edge_args;

size_t plb_edge_smooth_workspace_bytes(const plb_edge_args* args);
int plb_edge_smooth_loss(const plb_edge_args* args, void* stream);

/*
 * Stand-alone inverse warp (image out) and its vjp.
 * Replaces inverse_warp (geometry/pose_geometry.py:201-228) incl. F.grid_sample
 * (bilinear, zeros padding, align_corners=True).
 */
typedef struct plb_warp_args {
    int32_t B, H, W;
    int32_t rotation_mode;
    int32_t pose_inv;
    int32_t k_is_f64;
    const float* img;          /* [B,3,H,W] source                                                */
    const float* depth;        /* [B,H,W] depth of the target view                                */
    const float* pose;         /* [B,6] (contiguous rows; row stride pose_stride floats)          */
    int32_t pose_stride;
    int32_t reserved;
    const void* K;             /* [B,3,3]                                                          */
    float* out;                /* fwd: [B,3,H,W] warped image                                      */
    /* backward only */
    const float* g_out;        /* [B,3,H,W] cotangent                                              */
    float* g_img;              /* out (ACCUMULATED, caller zeroes) or NULL                         */
    float* g_depth;            /* out (written) [B,H,W] or NULL                                    */
    float* g_pose;             /* out (written) [B,6] or NULL                                      */
    void* workspace;           /* plb_warp_workspace_bytes() bytes, zero-filled once (bwd only)    */
    size_t workspace_bytes;
} plb_warp_args;

size_t plb_warp_workspace_bytes(const plb_warp_args* args);
int plb_warp_forward(const plb_warp_args* args, void* stream);
int plb_warp_backward(const plb_warp_args* args, void* stream);

/* Transform.reconstruct (geometry/transform.py:74-105): Xc[B,3,H,W] = (K^-1 . pixel) * depth. */
int plb_reconstruct(const float* depth, const void* K, int32_t k_is_f64, int32_t B, int32_t H,
                    int32_t W, float* Xc, void* stream);
/* Transform.project (geometry/transform.py:114-150): X[B,3,H,W], Tcw[B,4,4] -> grid[B,H,W,2] in [-1,1]. */
int plb_project(const float* X, const void* K, int32_t k_is_f64, const float* Tcw, int32_t B,
                int32_t H, int32_t W, float* grid, void* stream);
/* pose [B,6] -> [B,4,4] (axis-angle: transformation_from_parameters, geometry/pose_geometry.py:124-136;
 * euler: pose_vec2mat :97-108 padded with [0 0 0 1]); invert!=0 applies invert_pose (:110-115). */
int plb_pose_matrix(const float* pose, int32_t pose_stride, int32_t B, int32_t rotation_mode,
                    int32_t invert, float* M44, void* stream);
/* vjp of plb_pose_matrix: g_M44 [B,4,4] -> g_pose [B,6]. */
int plb_pose_matrix_backward(const float* pose, int32_t pose_stride, int32_t B, int32_t rotation_mode,
                             int32_t invert, const float* g_M44, float* g_pose, void* stream);
/* disp_to_depth (geometry/pose_geometry.py:70-95), elementwise; g != NULL also writes d depth/d disp. */
int plb_disp_to_depth(const float* disp, int64_t n, float a, float b, float* depth, void* stream);
int plb_disp_to_depth_backward(const float* disp, const float* g_depth, int64_t n, float a, float b,
                               float* g_disp, void* stream);

/*
 * Stand-alone photometric maps (the reference's dormant functions) and their vjp.
 *   w_ssim=1,    w_l1=0,    clip<0 : SSIM.standard_loss (losses.py:12-54)
 *   w_ssim=0.85, w_l1=0.15, clip=0.5: Losses.compute_photometric_loss (losses.py:66-84);
 *   w_ssim=0,    w_l1=1             : its no_ssim=True branch (losses.py:73-74)
 * out = w_ssim * clamp((1 - SSIM3x3(x, y)) / 2, 0, 1) + w_l1 * |y - x|, ReflectionPad2d(1) borders;
 * clip >= 0 additionally clamps at mean + clip * std (unbiased) of the whole map, the threshold being a
 * detached device scalar (`threshold`, written by the forward call, read by the backward call).
 */
typedef struct plb_photomap_args {
    int32_t B, C, H, W;
    const float* x;            /* [B,C,H,W] predicted image                                       */
    const float* y;            /* [B,C,H,W] target image                                          */
    float C1, C2;              /* 1e-4, 9e-4 in the reference                                      */
    float w_ssim, w_l1;
    float clip;                /* < 0: no clip                                                     */
    int32_t reserved;
    float* out;                /* fwd out [B,C,H,W]                                                */
    float* threshold;          /* device scalar (clip >= 0)                                        */
    const float* g_out;        /* bwd in  [B,C,H,W]                                                */
    float* g_x;                /* bwd out (written) or NULL                                        */
    float* g_y;                /* bwd out (written) or NULL                                        */
    void* workspace;           /* plb_photometric_map_workspace_bytes() bytes, zero-filled once (clip >= 0) */
    size_t workspace_bytes;
} plb_photomap_args;

size_t plb_photometric_map_workspace_bytes(const plb_photomap_args* args);
int plb_photometric_map(const plb_photomap_args* args, void* stream);
int plb_photometric_map_backward(const plb_photomap_args* args, void* stream);

/*
 * Depth image -> pseudo-LiDAR point cloud in the velodyne frame.
 * Replaces PseudoLiDAR.project_PL (pseudo-lidar/utils/PseudoLiDAR.py:69-110),
 * fp64 arithmetic in the reference's operation order, order-preserving
 * compaction of the valid points, optional [0::sparsity] decimation.
 */
typedef struct plb_cloud_args {
    int32_t B, H, W;           /* B independent depth images                                       */
    int32_t sparsity;          /* 0 = keep all valid points, n = keep every n-th valid point       */
    const float* depth;        /* [B,H,W] f32                                                      */
    double P[12];              /* P_rect_02 3x4 row-major (PseudoLiDAR.py:60)                      */
    double Tinv[16];           /* camera->velodyne 4x4 row-major exactly as inverse_rigid_trans builds
                                  it (PseudoLiDAR.py:39-46): [R^T | -R^T t] with an ALL-ZERO last row  */
    double* cloud_f64;         /* out [B, H*W, 4] f64 (parity layout) or NULL                      */
    float* cloud_f32;          /* out [B, H*W, 4] f32 x,y,z,i (PointCloud2 layout,
                                  PseudoLidarPipeline.py:51-54) or NULL                            */
    int32_t* index;            /* out [B, H*W] row-major pixel index of each kept point, or NULL   */
    uint8_t* valid;            /* out [B, H*W] mask (cloud_x>=0 & cloud_z<1) before decimation, or NULL */
    int32_t* count;            /* out [B] number of points written per image                       */
    void* workspace;           /* plb_cloud_workspace_bytes() bytes (no initialisation needed)     */
    size_t workspace_bytes;
} plb_cloud_args;

size_t plb_cloud_workspace_bytes(const plb_cloud_args* args);
int plb_cloud_project(const plb_cloud_args* args, void* stream);

/*
 * Velodyne sweep -> sparse depth image: the inverse of plb_cloud_project.
 * Replaces Transform.project_velo_to_img (pseudo-lidar/Transform/Transform.py:69-104): per point
 * dist = sqrt(x^2+y^2+z^2) in fp32, xyz = T.[x y z 1], uv = P.xyz in fp64 (the rounding order of the
 * per-point np.matmul), uv /= uv[2]; kept when 0 <= u < W, 0 <= v < H, dist <= 120, x > 0; the cell
 * (int(v), int(u)) takes xyz[2] of the LAST kept point in sweep order (largest index), 0 elsewhere.
 */
typedef struct plb_velo_args {
    int32_t B, N;              /* B sweeps of up to N points each                                   */
    int32_t H, W;              /* image height (Transform.height) and width (Transform.width)       */
    int32_t point_stride;      /* floats per point: 4 = KITTI .bin x,y,z,reflectance (16-byte aligned), >= 3 */
    int32_t reserved;
    const float* points;       /* [B, N, point_stride] f32                                          */
    const int32_t* counts;     /* [B] points used per sweep (<= N), or NULL = N everywhere          */
    double T[16];              /* velodyne->camera 4x4 row-major, last row 0 0 0 1 (Transform.py:61-62) */
    double P[12];              /* camera->image 3x4 row-major (Transform.py:65)                     */
    double* depth_f64;         /* out [B,H,W] f64 (parity layout) or NULL                           */
    float* depth_f32;          /* out [B,H,W] f32 or NULL (at least one of the two)                 */
    int32_t* winner;           /* out [B,H,W] index of the point that owns the cell, -1 = empty, or NULL */
    void* workspace;           /* plb_velo_workspace_bytes() bytes, zero-filled once (self-cleaning) */
    size_t workspace_bytes;
} plb_velo_args;

size_t plb_velo_workspace_bytes(const plb_velo_args* args);
int plb_velo_project(const plb_velo_args* args, void* stream);

/*
 * Loader-side frame preparation.  Replaces, for a batch of decoded uint8 frames,
 * the transform chain `ToTensor -> ToPILImage -> Resize((H, W)) -> ToTensor -> Normalize` (trainer.py:97-103) that
 * KittiDataset.load_img (dataloaders.py:32-49) applies to `np.asarray(Image.open(path), float32) / 255.0`, and the
 * intrinsics scaling of dataloaders.py:95-98.  Bit-exact: the float32 round trip in front of ToPILImage truncates,
 * Resize is Pillow's antialiased bilinear ImagingResample in 8-bit fixed point (coefficient tables in fp64, horizontal
 * pass then vertical pass, each rounded to uint8), ToTensor / Normalize are IEEE fp32 operations.
 */
typedef struct plb_prep_args {
    int32_t B;                 /* frames                                                            */
    int32_t in_h, in_w;        /* size of the decoded frames                                        */
    int32_t H, W;              /* network resolution (config: image_height, image_width)            */
    int32_t n_K;               /* intrinsics matrices in K_in / K_out: 0 = one per frame (B); a training sample has ONE
                                  matrix for its 1 + n_src frames, so a batch of 3 B frames comes with n_K = B.
                                  n_K > B: PLB_EINVAL                                               */
    const uint8_t* frames;     /* [B, in_h, in_w, 3] uint8 RGB, HWC as PIL decodes them             */
    float mean[3];             /* 0.485, 0.456, 0.406 in the reference                              */
    float stdev[3];            /* 0.229, 0.224, 0.225                                               */
    float* out_planar;         /* out [B,3,H,W] fp32 (the layout the networks and the loss read) or NULL */
    float* out_nhwc4;          /* out [B,H,W,4] fp32 (r,g,b,0) or NULL (at least one of the two)     */
    const double* K_in;        /* [n_K,3,3] f64 intrinsics at the decoded size, or NULL             */
    double* K_out;             /* out [n_K,3,3] f64: row 0 * W / in_w, row 1 * H / in_h (NULL iff K_in is NULL) */
    void* workspace;           /* plb_prep_workspace_bytes() bytes (no initialisation needed)       */
    size_t workspace_bytes;
} plb_prep_args;

size_t plb_prep_workspace_bytes(const plb_prep_args* args);
int plb_prep_frames(const plb_prep_args* args, void* stream);

/* Library identification: "plb200 <version> sm_100a". */
const char* plb_version(void);
/* Number of kernel launches issued by this library since load (all entry points). */
uint64_t plb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PLB200_H */
