// Velodyne sweep -> sparse depth image (pseudo-lidar/Transform/Transform.py:69-104), the inverse of
// the pseudo-LiDAR back-projection.
//
// The reference walks the points in order and lets LATER points overwrite earlier ones in the cell
// (int(u), int(v)); here the winner of a cell is the kept point with the LARGEST index, decided by an
// integer atomicMax - order independent, so the result is bitwise repeatable and equal to the
// sequential loop's.  Two launches:
//   (1) velo_scatter_kernel: one thread per point: dist (fp32, the cloud's dtype), T . [x y z 1] and
//       P . xyz in fp64 with the rounding order of the per-point np.matmul (rounded products, then
//       (p0 + p2) + (p1 + p3)), IEEE divisions, the reference's six tests, atomicMax(index + 1) into a
//       zero-filled int32 cell map;
//   (2) velo_resolve_kernel: one thread per PAIR of cells: depth = xyz[2] of the winner, which launch (1) left in a
//       per-point fp64 scratch row (written by the points that reach an image cell only; re-deriving it from the
//       point cost every warp two divergent fp64 dot products and three conversions for its ~8 % of owned cells and
//       made the stream issue-bound), winner index out, and the cell map is zeroed again (self-cleaning workspace).
#include "common.cuh"

namespace plb {

constexpr int VL_THREADS = 256;
constexpr int VL_UNROLL = 4;                     // cell pairs per thread of the resolve launch

// workspace: the int32 cell map [B,H,W] (zero between calls), then the fp64 depth of every point [B,N] (no
// initialisation needed: a cell only ever names a point that wrote its row in the same call)
__host__ __device__ inline size_t velo_cells_bytes(const plb_velo_args& a) {
    return ((size_t)a.B * a.H * a.W * sizeof(int32_t) + 255) / 256 * 256;
}
__device__ __forceinline__ double* velo_scratch(const plb_velo_args& a) {
    return (double*)((char*)a.workspace + velo_cells_bytes(a));
}

// sum of the four rounded products in the order the 4-wide SIMD product + horizontal add produces
__device__ __forceinline__ double dot4_np(const double* m, double x, double y, double z, double w) {
    const double p0 = __dmul_rn(x, m[0]), p1 = __dmul_rn(y, m[1]), p2 = __dmul_rn(z, m[2]), p3 = __dmul_rn(w, m[3]);
    return __dadd_rn(__dadd_rn(p0, p2), __dadd_rn(p1, p3));
}

__device__ __forceinline__ void velo_load(const float* pts, int stride, size_t i, float& x, float& y, float& z) {
    if (stride == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pts) + i);
        x = q.x; y = q.y; z = q.z;
    } else {
        const float* p = pts + i * (size_t)stride;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
}

__global__ void __launch_bounds__(VL_THREADS)
velo_scatter_kernel(const __grid_constant__ plb_velo_args a) {
    const int b = blockIdx.y;
    const int n = a.counts != nullptr ? min(__ldg(a.counts + b), a.N) : a.N;
    const float* pts = a.points + (size_t)b * a.N * a.point_stride;
    int32_t* cells = (int32_t*)a.workspace + (size_t)b * a.H * a.W;
    double* zs = velo_scratch(a) + (size_t)b * a.N;
    for (int i = blockIdx.x * VL_THREADS + threadIdx.x; i < n; i += gridDim.x * VL_THREADS) {
        float x, y, z;
        velo_load(pts, a.point_stride, (size_t)i, x, y, z);
        // Transform.py:84-87: fp32, x**2 + y**2 + z**2 left to right, no contraction
        const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        if (!(dist <= 120.0f) || !(x > 0.0f)) continue;           // the two tests that need no projection
        const double xd = (double)x, yd = (double)y, zd = (double)z;
        double c[4], uvw[3];
#pragma unroll
        for (int r = 0; r < 4; ++r) c[r] = dot4_np(a.T + 4 * r, xd, yd, zd, 1.0);
#pragma unroll
        for (int r = 0; r < 3; ++r) uvw[r] = dot4_np(a.P + 4 * r, c[0], c[1], c[2], c[3]);
        const double u = __ddiv_rn(uvw[0], uvw[2]), v = __ddiv_rn(uvw[1], uvw[2]);
        if (u >= 0.0 && u < (double)a.W && v >= 0.0 && v < (double)a.H) {    // NaN fails every test
            atomicMax(cells + ((int)v * a.W + (int)u), i + 1);
            zs[i] = c[2];
        }
    }
}

// depth (xyz[2]) of the point that owns a cell: the scratch row launch 1 wrote
__device__ __forceinline__ double velo_cell_depth(const double* zs, int w) { return __ldcg(zs + (w - 1)); }

// Two cells per thread: one 8-byte load of the cell map, one 16-byte store of the fp64 image (most cells are empty;
// the kernel is a 12-byte-per-cell stream with a rare side trip to the owning point).
__global__ void __launch_bounds__(VL_THREADS)
velo_resolve_kernel(const __grid_constant__ plb_velo_args a) {
    const int b = blockIdx.y;
    const int ncell = a.H * a.W;
    const double* zs = velo_scratch(a) + (size_t)b * a.N;
    int32_t* cells = (int32_t*)a.workspace + (size_t)b * ncell;
    const size_t base = (size_t)b * ncell;
    // pairs are aligned when the image's first cell is: (b * ncell) even -> 8-byte cell pairs, 16-byte fp64 pairs
    const bool paired = ((base & 1) == 0);
    const int npair = paired ? ncell / 2 : 0;
    // VL_UNROLL pairs per thread, a warp's pairs of one round adjacent (coalesced 8-byte loads, 16-byte stores); every
    // cell-map load of the thread is in flight before the first is used
    const int stride = gridDim.x * VL_THREADS;
    for (int q0 = blockIdx.x * VL_THREADS + threadIdx.x; q0 < npair; q0 += stride * VL_UNROLL) {
        int2 w[VL_UNROLL];
#pragma unroll
        for (int r = 0; r < VL_UNROLL; ++r) {
            const int q2 = q0 + r * stride;
            w[r] = q2 < npair ? *reinterpret_cast<const int2*>(cells + 2 * q2) : make_int2(0, 0);
        }
#pragma unroll
        for (int r = 0; r < VL_UNROLL; ++r) {
            const int q2 = q0 + r * stride;
            if (q2 >= npair) break;
            const int q = 2 * q2;
            double d0 = 0.0, d1 = 0.0;
            if (w[r].x > 0 || w[r].y > 0) {
                if (w[r].x > 0) d0 = velo_cell_depth(zs, w[r].x);
                if (w[r].y > 0) d1 = velo_cell_depth(zs, w[r].y);
                *reinterpret_cast<int2*>(cells + q) = make_int2(0, 0);
            }
            const size_t o = base + q;
            if (a.depth_f64 != nullptr) __stcs(reinterpret_cast<double2*>(a.depth_f64 + o), make_double2(d0, d1));
            if (a.depth_f32 != nullptr) __stcs(reinterpret_cast<float2*>(a.depth_f32 + o), make_float2((float)d0, (float)d1));
            if (a.winner != nullptr) __stcs(reinterpret_cast<int2*>(a.winner + o), make_int2(w[r].x - 1, w[r].y - 1));
        }
    }
    // the cells the pairs do not cover: the odd last one, or the whole image when it starts on an odd cell
    for (int q = 2 * npair + blockIdx.x * VL_THREADS + threadIdx.x; q < ncell; q += gridDim.x * VL_THREADS) {
        const int w = cells[q];
        double depth = 0.0;
        if (w > 0) {
            depth = velo_cell_depth(zs, w);
            cells[q] = 0;
        }
        const size_t o = base + q;
        if (a.depth_f64 != nullptr) __stcs(a.depth_f64 + o, depth);
        if (a.depth_f32 != nullptr) __stcs(a.depth_f32 + o, (float)depth);
        if (a.winner != nullptr) __stcs(a.winner + o, w - 1);
    }
}

size_t velo_workspace_bytes(const plb_velo_args* a) {
    return velo_cells_bytes(*a) + ((size_t)a->B * a->N * sizeof(double) + 255) / 256 * 256;
}

int velo_launch(const plb_velo_args* a, cudaStream_t st) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 1 || a->W < 1 || a->N < 0 || a->B > 65535) return PLB_EINVAL;
    if (a->point_stride < 3) return PLB_EINVAL;
    if ((int64_t)a->H * a->W > (int64_t)1 << 30 || a->N > (1 << 30)) return PLB_EINVAL;
    if ((a->N > 0 && !a->points) || (!a->depth_f64 && !a->depth_f32)) return PLB_ENULL;
    if (a->point_stride == 4 && ((uintptr_t)a->points & 15)) return PLB_EINVAL;
    if (((uintptr_t)a->depth_f64 & 15) || ((uintptr_t)a->depth_f32 & 7) || ((uintptr_t)a->winner & 7) ||
        ((uintptr_t)a->workspace & 7)) return PLB_EINVAL;      // two cells per thread: paired stores
    if (!a->workspace || a->workspace_bytes < velo_workspace_bytes(a)) return PLB_EWORKSPACE;
    if (a->N > 0) {
        dim3 grid(min((a->N + VL_THREADS - 1) / VL_THREADS, 148 * 8), a->B);
        velo_scatter_kernel<<<grid, VL_THREADS, 0, st>>>(*a);
        ++g_launches;
        PLB_CHECK_LAUNCH();
    }
    dim3 grid2(min((a->H * a->W / 2 + VL_THREADS * VL_UNROLL - 1) / (VL_THREADS * VL_UNROLL) + 1, 148 * 8), a->B);
    velo_resolve_kernel<<<grid2, VL_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
