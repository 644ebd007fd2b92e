// Depth image -> pseudo-LiDAR cloud (pseudo-lidar/utils/PseudoLiDAR.py:69-110).
//
// fp64, the reference's operation order, no FMA contraction in the image->camera
// step (numpy evaluates ((u-c_u)*d)/f_u + b_x one ufunc at a time) and the FMA
// chain acc=p0*r0; acc=fma(p_k,r_k,acc) that the BLAS dgemm behind np.matmul
// runs for the [N,4]x[4,4] product - so x,y,z and the validity mask
// (cloud_x >= 0 & cloud_z < 1) are bit-identical to the reference's.
//
// Order-preserving compaction in three launches with no inter-block waiting:
//   (1) per-tile valid counts.  Validity (cloud_x >= 0 & cloud_z < 1) is decided from a division-free
//       fp64 evaluation whenever the value is further than 1e-9 from its threshold (the two
//       evaluations differ by < 1e-12 for |x|, |y|, d below 1e4); only the rare borderline pixel runs
//       the exact chain - so the mask stays bit-identical to the reference's at a third of the work;
//   (2) exclusive scan of the tile counts of each image (one warp per image);
//   (3) every tile computes its points exactly once, and writes them at their final row-major rank;
//       [0::sparsity] keeps ranks divisible by `sparsity`.  Integer prefix sums: bitwise repeatable.
#include "common.cuh"

namespace plb {

constexpr int CL_THREADS = 256;
constexpr int CL_ITEMS = 4;                       // consecutive pixels per thread (4 x 4 doubles in registers)
constexpr int CL_TILE = CL_THREADS * CL_ITEMS;    // 1024 pixels per block

struct CloudConst {
    double c_u, c_v, f_u, f_v, b_x, b_y;
    double rf_u, rf_v;                 // RN(1 / f_u), RN(1 / f_v): IEEE divisions done once on the host
    double Ti[16];
};

// n / f for a fixed divisor: q0 = n * RN(1/f); exact remainder by FMA; one correction (Markstein).  The
// result is the correctly rounded quotient except possibly when n / f lies within ~2^-105 (relative) of
// a rounding boundary, so it replaces the ~25-instruction IEEE division routine in the per-pixel chain.
__device__ __forceinline__ double div_const(double n, double f, double rf) {
    const double q0 = __dmul_rn(n, rf);
    const double rem = __fma_rn(-q0, f, n);
    return __fma_rn(rem, rf, q0);
}

template <bool EXACT>
__device__ __forceinline__ void cloud_point(const CloudConst& cc, int col, int row, float depth, double (&o)[4]) {
    const double d = (double)depth;
    const double nx = __dmul_rn(__dsub_rn((double)col, cc.c_u), d), ny = __dmul_rn(__dsub_rn((double)row, cc.c_v), d);
    const double x = __dadd_rn(EXACT ? __ddiv_rn(nx, cc.f_u) : div_const(nx, cc.f_u, cc.rf_u), cc.b_x);
    const double y = __dadd_rn(EXACT ? __ddiv_rn(ny, cc.f_v) : div_const(ny, cc.f_v, cc.rf_v), cc.b_y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double acc = __dmul_rn(x, cc.Ti[k * 4 + 0]);
        acc = __fma_rn(y, cc.Ti[k * 4 + 1], acc);
        acc = __fma_rn(d, cc.Ti[k * 4 + 2], acc);
        acc = __fma_rn(1.0, cc.Ti[k * 4 + 3], acc);
        o[k] = acc;
    }
}

// One pixel: the point (reference operation order) and its validity (cloud_x >= 0 & cloud_z < 1).
// The mask must be bit-identical to the reference's and identical in the count and the write launch:
// a value closer than 1e-9 to its threshold (or a huge / non-finite input) is re-evaluated with IEEE
// divisions, so the cheap division can never flip a decision.
__device__ __forceinline__ bool cloud_eval(const CloudConst& cc, int col, int row, float depth, double (&o)[4]) {
    cloud_point<false>(cc, col, row, depth, o);
    const bool safe = fabs(o[0]) > 1e-9 && fabs(o[2] - 1.0) > 1e-9 && fabs(o[0]) < 1e12 && fabs(o[2]) < 1e12;
    if (!safe) cloud_point<true>(cc, col, row, depth, o);
    return o[0] >= 0.0 && o[2] < 1.0;
}

struct CloudHost { double rf_u, rf_v, b_x, b_y; };

__device__ __forceinline__ CloudConst cloud_const(const plb_cloud_args& a, const CloudHost& h) {
    CloudConst cc;
    cc.c_u = a.P[2]; cc.c_v = a.P[6]; cc.f_u = a.P[0]; cc.f_v = a.P[5];
    cc.b_x = h.b_x; cc.b_y = h.b_y; cc.rf_u = h.rf_u; cc.rf_v = h.rf_v;
#pragma unroll
    for (int k = 0; k < 16; ++k) cc.Ti[k] = a.Tinv[k];
    return cc;
}

__device__ __forceinline__ void cloud_load(const float* depth, size_t img_off, int p0, int npx, float (&dv)[CL_ITEMS]) {
    if (p0 + CL_ITEMS <= npx && (((img_off + p0) & 3) == 0)) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(depth + p0));
        dv[0] = q.x; dv[1] = q.y; dv[2] = q.z; dv[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) dv[k] = (p0 + k < npx) ? __ldg(depth + p0 + k) : 0.0f;
    }
}

// launch 1: valid points per tile
__global__ void __launch_bounds__(CL_THREADS)
cloud_count_kernel(const __grid_constant__ plb_cloud_args a, const CloudHost h) {
    const int b = blockIdx.y, blk = blockIdx.x, tiles = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npx = a.H * a.W;
    const CloudConst cc = cloud_const(a, h);
    const float* depth = a.depth + (size_t)b * npx;
    __shared__ int s_warp[CL_THREADS / 32];
    const int p0 = blk * CL_TILE + tid * CL_ITEMS;
    float dv[CL_ITEMS];
    cloud_load(depth, (size_t)b * npx, p0, npx, dv);
    int row = p0 / a.W, col = p0 - row * a.W;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; ++k) {
        double o[4];
        if (p0 + k < npx && cloud_eval(cc, col, row, dv[k], o)) ++mine;
        if (++col == a.W) { col = 0; ++row; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_warp[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
#pragma unroll
        for (int w = 0; w < CL_THREADS / 32; ++w) total += s_warp[w];
        ((int32_t*)a.workspace)[(size_t)b * tiles + blk] = total;
    }
}

// launch 2: counts -> exclusive prefix per image (in place), one warp per image; also the image's point count
__global__ void __launch_bounds__(32)
cloud_scan_kernel(const __grid_constant__ plb_cloud_args a, int tiles) {
    const int b = blockIdx.x, lane = threadIdx.x;
    int32_t* counts = (int32_t*)a.workspace + (size_t)b * tiles;
    int carry = 0;
    for (int base = 0; base < tiles; base += 32) {
        const int k = base + lane;
        const int c = k < tiles ? counts[k] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (k < tiles) counts[k] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && a.count != nullptr) {
        const int sp = a.sparsity > 0 ? a.sparsity : 1;
        a.count[b] = (carry + sp - 1) / sp;
    }
}

// launch 3: points at their final rank
__global__ void __launch_bounds__(CL_THREADS)
cloud_write_kernel(const __grid_constant__ plb_cloud_args a, const CloudHost h) {
    const int b = blockIdx.y, blk = blockIdx.x, tiles = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npx = a.H * a.W;
    const CloudConst cc = cloud_const(a, h);
    const float* depth = a.depth + (size_t)b * npx;
    __shared__ int s_warp[CL_THREADS / 32];
    const int base = __ldg((const int32_t*)a.workspace + (size_t)b * tiles + blk);

    const int p0 = blk * CL_TILE + tid * CL_ITEMS;
    double pts[CL_ITEMS][4];
    unsigned vmask = 0;
    {
        float dv[CL_ITEMS];
        cloud_load(depth, (size_t)b * npx, p0, npx, dv);
        int row = p0 / a.W, col = p0 - row * a.W;
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) {
            if (p0 + k < npx) {
                if (cloud_eval(cc, col, row, dv[k], pts[k])) vmask |= 1u << k;
            }
            if (++col == a.W) { col = 0; ++row; }
        }
    }
    const int mine = __popc(vmask);
    // inclusive scan of `mine` across the warp, then across warps
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0;
#pragma unroll
    for (int w = 0; w < CL_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) warp_off += c;
    }
    const int lrank0 = warp_off + incl - mine;     // rank of this thread's first valid point inside the tile
    const int sp = a.sparsity > 0 ? a.sparsity : 1;
    if (a.valid != nullptr) {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k)
            if (p0 + k < npx) a.valid[(size_t)b * npx + p0 + k] = (vmask >> k) & 1u;
    }
    int rank = base + lrank0;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; ++k) {
        if ((vmask >> k) & 1u) {
            if (rank % sp == 0) {
                const size_t pos = (size_t)b * npx + rank / sp;
                if (a.cloud_f64 != nullptr) {
                    double2* o = reinterpret_cast<double2*>(a.cloud_f64 + pos * 4);
                    __stcs(o, make_double2(pts[k][0], pts[k][1]));
                    __stcs(o + 1, make_double2(pts[k][2], pts[k][3]));
                }
                if (a.cloud_f32 != nullptr)
                    __stcs(reinterpret_cast<float4*>(a.cloud_f32) + pos,
                           make_float4((float)pts[k][0], (float)pts[k][1], (float)pts[k][2], (float)pts[k][3]));
                if (a.index != nullptr) a.index[pos] = p0 + k;
            }
            ++rank;
        }
    }
}

static inline int cloud_blocks(const plb_cloud_args* a) { return (a->H * a->W + CL_TILE - 1) / CL_TILE; }

size_t cloud_workspace_bytes(const plb_cloud_args* a) {
    return ((size_t)a->B * cloud_blocks(a) * sizeof(int32_t) + 255) / 256 * 256;
}

int cloud_launch(const plb_cloud_args* a, cudaStream_t st) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->H < 1 || a->W < 1 || a->sparsity < 0) return PLB_EINVAL;
    if ((int64_t)a->H * a->W > (int64_t)1 << 30) return PLB_EINVAL;
    if (a->B > 65535) return PLB_EINVAL;
    if (!a->depth || !a->count) return PLB_ENULL;
    if (!a->workspace || a->workspace_bytes < cloud_workspace_bytes(a)) return PLB_EWORKSPACE;
    const int tiles = cloud_blocks(a);
    dim3 grid(tiles, a->B);
    CloudHost h;                                    // IEEE double divisions, as numpy does them (PseudoLiDAR.py:84-85)
    h.rf_u = 1.0 / a->P[0]; h.rf_v = 1.0 / a->P[5];
    h.b_x = a->P[3] / (-a->P[0]); h.b_y = a->P[7] / (-a->P[5]);
    cloud_count_kernel<<<grid, CL_THREADS, 0, st>>>(*a, h);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    cloud_scan_kernel<<<a->B, 32, 0, st>>>(*a, tiles);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    cloud_write_kernel<<<grid, CL_THREADS, 0, st>>>(*a, h);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
