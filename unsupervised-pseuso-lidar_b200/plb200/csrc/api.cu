// extern "C" surface of libplb200.so (include/plb200.h).  Thin: argument
// checks live next to each kernel; nothing here allocates or synchronises.
#include "common.cuh"

namespace plb {
unsigned long long g_launches = 0;

int photo_l1_launch(const plb_photo_args*, cudaStream_t);
size_t photo_workspace_bytes(const plb_photo_args*);
int smooth_launch(const plb_smooth_args*, cudaStream_t);
size_t smooth_workspace_bytes(const plb_smooth_args*);
int warp_forward_launch(const plb_warp_args*, cudaStream_t);
int warp_backward_launch(const plb_warp_args*, cudaStream_t);
size_t warp_workspace_bytes(const plb_warp_args*);
int reconstruct_launch(const float*, const void*, int, int, int, int, float*, cudaStream_t);
int project_launch(const float*, const void*, int, const float*, int, int, int, float*, cudaStream_t);
int pose_matrix_launch(const float*, int, int, int, int, float*, cudaStream_t);
int pose_matrix_bwd_launch(const float*, int, int, int, int, const float*, float*, cudaStream_t);
int disp_to_depth_launch(const float*, int64_t, float, float, float*, cudaStream_t);
int disp_to_depth_bwd_launch(const float*, const float*, int64_t, float, float, float*, cudaStream_t);
int edge_launch(const plb_edge_args*, cudaStream_t);
size_t edge_workspace_bytes(const plb_edge_args*);
int photomap_launch(const plb_photomap_args*, cudaStream_t);
int photomap_bwd_launch(const plb_photomap_args*, cudaStream_t);
size_t photomap_workspace_bytes(const plb_photomap_args*);
int cloud_launch(const plb_cloud_args*, cudaStream_t);
size_t cloud_workspace_bytes(const plb_cloud_args*);
int velo_launch(const plb_velo_args*, cudaStream_t);
size_t velo_workspace_bytes(const plb_velo_args*);
int prep_launch(const plb_prep_args*, cudaStream_t);
size_t prep_workspace_bytes(const plb_prep_args*);
}  // namespace plb

extern "C" {

size_t plb_photo_workspace_bytes(const plb_photo_args* a) { return a ? plb::photo_workspace_bytes(a) : 0; }
int plb_photo_loss(const plb_photo_args* a, void* stream) { return plb::photo_l1_launch(a, (cudaStream_t)stream); }

size_t plb_smooth_workspace_bytes(const plb_smooth_args* a) { return a ? plb::smooth_workspace_bytes(a) : 0; }
int plb_smooth_loss(const plb_smooth_args* a, void* stream) { return plb::smooth_launch(a, (cudaStream_t)stream); }

size_t plb_edge_smooth_workspace_bytes(const plb_edge_args* a) { return a ? plb::edge_workspace_bytes(a) : 0; }
int plb_edge_smooth_loss(const plb_edge_args* a, void* stream) { return plb::edge_launch(a, (cudaStream_t)stream); }

size_t plb_warp_workspace_bytes(const plb_warp_args* a) { return a ? plb::warp_workspace_bytes(a) : 0; }
int plb_warp_forward(const plb_warp_args* a, void* stream) { return plb::warp_forward_launch(a, (cudaStream_t)stream); }
int plb_warp_backward(const plb_warp_args* a, void* stream) { return plb::warp_backward_launch(a, (cudaStream_t)stream); }

int plb_reconstruct(const float* depth, const void* K, int32_t k_is_f64, int32_t B, int32_t H, int32_t W, float* Xc,
                    void* stream) {
    return plb::reconstruct_launch(depth, K, k_is_f64, B, H, W, Xc, (cudaStream_t)stream);
}
int plb_project(const float* X, const void* K, int32_t k_is_f64, const float* Tcw, int32_t B, int32_t H, int32_t W,
                float* grid, void* stream) {
    return plb::project_launch(X, K, k_is_f64, Tcw, B, H, W, grid, (cudaStream_t)stream);
}
int plb_pose_matrix(const float* pose, int32_t pose_stride, int32_t B, int32_t rotation_mode, int32_t invert,
                    float* M44, void* stream) {
    return plb::pose_matrix_launch(pose, pose_stride, B, rotation_mode, invert, M44, (cudaStream_t)stream);
}
int plb_pose_matrix_backward(const float* pose, int32_t pose_stride, int32_t B, int32_t rotation_mode, int32_t invert,
                             const float* g_M44, float* g_pose, void* stream) {
    return plb::pose_matrix_bwd_launch(pose, pose_stride, B, rotation_mode, invert, g_M44, g_pose, (cudaStream_t)stream);
}
int plb_disp_to_depth(const float* disp, int64_t n, float a, float b, float* depth, void* stream) {
    return plb::disp_to_depth_launch(disp, n, a, b, depth, (cudaStream_t)stream);
}
int plb_disp_to_depth_backward(const float* disp, const float* g_depth, int64_t n, float a, float b, float* g_disp,
                               void* stream) {
    return plb::disp_to_depth_bwd_launch(disp, g_depth, n, a, b, g_disp, (cudaStream_t)stream);
}

size_t plb_photometric_map_workspace_bytes(const plb_photomap_args* a) { return a ? plb::photomap_workspace_bytes(a) : 0; }
int plb_photometric_map(const plb_photomap_args* a, void* stream) { return plb::photomap_launch(a, (cudaStream_t)stream); }
int plb_photometric_map_backward(const plb_photomap_args* a, void* stream) {
    return plb::photomap_bwd_launch(a, (cudaStream_t)stream);
}

size_t plb_cloud_workspace_bytes(const plb_cloud_args* a) { return a ? plb::cloud_workspace_bytes(a) : 0; }
int plb_cloud_project(const plb_cloud_args* a, void* stream) { return plb::cloud_launch(a, (cudaStream_t)stream); }

size_t plb_velo_workspace_bytes(const plb_velo_args* a) { return a ? plb::velo_workspace_bytes(a) : 0; }
int plb_velo_project(const plb_velo_args* a, void* stream) { return plb::velo_launch(a, (cudaStream_t)stream); }

size_t plb_prep_workspace_bytes(const plb_prep_args* a) { return a ? plb::prep_workspace_bytes(a) : 0; }
int plb_prep_frames(const plb_prep_args* a, void* stream) { return plb::prep_launch(a, (cudaStream_t)stream); }

const char* plb_version(void) { return "plb200 0.2 sm_100a"; }
uint64_t plb_launch_count(void) { return plb::g_launches; }

}  // extern "C"
