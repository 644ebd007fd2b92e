// Depth image -> pseudo-LiDAR cloud (pseudo-lidar/utils/PseudoLiDAR.py:69-110).
//
// fp64, the reference's operation order, no FMA contraction in the image->camera
// step (numpy evaluates ((u-c_u)*d)/f_u + b_x one ufunc at a time) and the FMA
// chain acc=p0*r0; acc=fma(p_k,r_k,acc) that the BLAS dgemm behind np.matmul
// runs for the [N,4]x[4,4] product - so x,y,z and the validity mask
// (cloud_x >= 0 & cloud_z < 1) are bit-identical to the reference's.
//
// Order-preserving compaction in two launches: (1) per-tile valid counts,
// (2) every tile sums the counts of the tiles before it in its image (a few
// hundred ints), scans its own pixels and writes its points at their final
// row-major rank; [0::sparsity] keeps ranks divisible by `sparsity`.
#include "common.cuh"

namespace plb {

constexpr int CL_THREADS = 256;
constexpr int CL_ITEMS = 8;                       // consecutive pixels per thread
constexpr int CL_TILE = CL_THREADS * CL_ITEMS;    // 2048 pixels per block

struct CloudConst {
    double c_u, c_v, f_u, f_v, b_x, b_y;
    double Ti[16];
};

__device__ __forceinline__ void cloud_point(const CloudConst& cc, int col, int row, float depth, double (&o)[4]) {
    const double d = (double)depth;
    const double x = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)col, cc.c_u), d), cc.f_u), cc.b_x);
    const double y = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)row, cc.c_v), d), cc.f_v), cc.b_y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double acc = __dmul_rn(x, cc.Ti[k * 4 + 0]);
        acc = __fma_rn(y, cc.Ti[k * 4 + 1], acc);
        acc = __fma_rn(d, cc.Ti[k * 4 + 2], acc);
        acc = __fma_rn(1.0, cc.Ti[k * 4 + 3], acc);
        o[k] = acc;
    }
}

__device__ __forceinline__ CloudConst cloud_const(const plb_cloud_args& a) {
    CloudConst cc;
    cc.c_u = a.P[2]; cc.c_v = a.P[6]; cc.f_u = a.P[0]; cc.f_v = a.P[5];
    cc.b_x = __ddiv_rn(a.P[3], -cc.f_u);
    cc.b_y = __ddiv_rn(a.P[7], -cc.f_v);
#pragma unroll
    for (int k = 0; k < 16; ++k) cc.Ti[k] = a.Tinv[k];
    return cc;
}

template <bool WRITE>
__global__ void __launch_bounds__(CL_THREADS)
cloud_kernel(const __grid_constant__ plb_cloud_args a) {
    const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npx = a.H * a.W;
    const CloudConst cc = cloud_const(a);
    int32_t* counts = (int32_t*)a.workspace + (size_t)b * nblk;
    const float* depth = a.depth + (size_t)b * npx;

    __shared__ int s_warp[CL_THREADS / 32];
    __shared__ int s_base;

    const int p0 = blk * CL_TILE + tid * CL_ITEMS;
    double pts[CL_ITEMS][4];
    unsigned vmask = 0;
#pragma unroll
    for (int k = 0; k < CL_ITEMS; ++k) {
        const int p = p0 + k;
        if (p < npx) {
            const int row = p / a.W, col = p - row * a.W;
            cloud_point(cc, col, row, __ldg(depth + p), pts[k]);
            if (pts[k][0] >= 0.0 && pts[k][2] < 1.0) vmask |= 1u << k;
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
    int warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CL_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) warp_off += c;
        total += c;
    }
    if (!WRITE) {
        if (tid == 0) counts[blk] = total;
        return;
    }
    // rank of this tile's first valid point = sum of the counts of earlier tiles
    if (warp == 0) {
        int acc = 0;
        for (int k = lane; k < blk; k += 32) acc += counts[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_base = acc;
    }
    __syncthreads();
    int rank = s_base + warp_off + incl - mine;
    const int sp = a.sparsity > 0 ? a.sparsity : 1;
    if (a.valid != nullptr) {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k)
            if (p0 + k < npx) a.valid[(size_t)b * npx + p0 + k] = (vmask >> k) & 1u;
    }
#pragma unroll
    for (int k = 0; k < CL_ITEMS; ++k) {
        if ((vmask >> k) & 1u) {
            if (rank % sp == 0) {
                const size_t pos = (size_t)b * npx + rank / sp;
                if (a.cloud_f64 != nullptr) {
                    double2* o = reinterpret_cast<double2*>(a.cloud_f64 + pos * 4);
                    o[0] = make_double2(pts[k][0], pts[k][1]);
                    o[1] = make_double2(pts[k][2], pts[k][3]);
                }
                if (a.cloud_f32 != nullptr)
                    reinterpret_cast<float4*>(a.cloud_f32)[pos] =
                        make_float4((float)pts[k][0], (float)pts[k][1], (float)pts[k][2], (float)pts[k][3]);
                if (a.index != nullptr) a.index[pos] = p0 + k;
            }
            ++rank;
        }
    }
    if (blk == nblk - 1 && tid == 0 && a.count != nullptr) {
        const int all = s_base + total;
        a.count[b] = (all + sp - 1) / sp;
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
    if (!a->depth || !a->count) return PLB_ENULL;
    if (!a->workspace || a->workspace_bytes < cloud_workspace_bytes(a)) return PLB_EWORKSPACE;
    dim3 grid(cloud_blocks(a), a->B);
    cloud_kernel<false><<<grid, CL_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    cloud_kernel<true><<<grid, CL_THREADS, 0, st>>>(*a);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
