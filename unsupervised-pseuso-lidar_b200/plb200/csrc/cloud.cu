// Depth image -> pseudo-LiDAR cloud (pseudo-lidar/utils/PseudoLiDAR.py:69-110).
//
// fp64, the reference's operation order, no FMA contraction in the image->camera
// step (numpy evaluates ((u-c_u)*d)/f_u + b_x one ufunc at a time) and the FMA
// chain acc=p0*r0; acc=fma(p_k,r_k,acc) that the BLAS dgemm behind np.matmul
// runs for the [N,4]x[4,4] product - so x,y,z and the validity mask
// (cloud_x >= 0 & cloud_z < 1) are bit-identical to the reference's.
//
// Order-preserving compaction in two launches with no inter-block waiting:
//   (1) cloud_count_kernel - validity of every pixel, ONE WARP per 1024-pixel tile in a rolled loop of eight 128-pixel
//       groups (the tile prefetched into L2 by one instruction; no shared memory, no block barrier): cloud_x and
//       cloud_z are affine in the depth, so validity (cloud_x >= 0 & cloud_z < 1) is decided by the sign of an fp32
//       value whenever that lies beyond a rigorous error bound of its threshold, by two fp64 FMAs per coordinate
//       otherwise, and only the rare borderline pixel runs the exact chain - the mask stays bit-identical to the
//       reference's.  Out: the tile's count and its 1024 validity bits (32 words, one per lane);
//   (2) cloud_write_direct_kernel (no decimation: the default) - a warp per tile walks it 32 pixels at a time, one
//       pixel per lane: validity bit by shuffle, output row = tile base + rank in the ballot, the exact fp64 chain in
//       place, one 256-bit store per row; or cloud_write_kernel ([0::sparsity]) - the tile compacts (pixel, depth) of
//       its valid pixels in shared memory and evaluates the kept ranks densely, one per thread.  Integer prefix sums:
//       bitwise repeatable.
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

// x, y, z only (the caller knows the 4th row of T_inv is zero and checks that the point is finite)
__device__ __forceinline__ void cloud_point3(const CloudConst& cc, int col, int row, float depth, double (&o)[4]) {
    const double d = (double)depth;
    const double nx = __dmul_rn(__dsub_rn((double)col, cc.c_u), d), ny = __dmul_rn(__dsub_rn((double)row, cc.c_v), d);
    const double x = __dadd_rn(div_const(nx, cc.f_u, cc.rf_u), cc.b_x);
    const double y = __dadd_rn(div_const(ny, cc.f_v, cc.rf_v), cc.b_y);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double acc = __dmul_rn(x, cc.Ti[k * 4 + 0]);
        acc = __fma_rn(y, cc.Ti[k * 4 + 1], acc);
        acc = __fma_rn(d, cc.Ti[k * 4 + 2], acc);
        acc = __fma_rn(1.0, cc.Ti[k * 4 + 3], acc);
        o[k] = acc;
    }
}

// streaming stores under a predicate (no branch in the instruction stream)
__device__ __forceinline__ void st_cs_f64x4_if(bool p, void* q, double a, double b, double c, double d) {   // 32-byte aligned
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.cs.v4.f64 [%1], {%2, %3, %4, %5}; }"
                 :: "r"((unsigned)p), "l"(q), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void st_cs_f32x4_if(bool p, void* q, float a, float b, float c, float d) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.cs.v4.f32 [%1], {%2, %3, %4, %5}; }"
                 :: "r"((unsigned)p), "l"(q), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// (double)i for 0 <= i < 2^31 without the conversion unit: the bits (0x43300000, i) are 2^52 + i exactly, and the
// subtraction is exact - one DADD on the fp64 pipe instead of an I2F.F64 on the (16x narrower) XU pipe
__device__ __forceinline__ double int_as_double_exact(int i) {
    return __dsub_rn(__hiloint2double(0x43300000, i), 4503599627370496.0);
}
// the same chain with the column / row already in fp64
__device__ __forceinline__ void cloud_point3d(const CloudConst& cc, double dcol, double drow, double d, double (&o)[4]) {
    const double nx = __dmul_rn(__dsub_rn(dcol, cc.c_u), d), ny = __dmul_rn(__dsub_rn(drow, cc.c_v), d);
    const double x = __dadd_rn(div_const(nx, cc.f_u, cc.rf_u), cc.b_x);
    const double y = __dadd_rn(div_const(ny, cc.f_v, cc.rf_v), cc.b_y);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double acc = __dmul_rn(x, cc.Ti[k * 4 + 0]);
        acc = __fma_rn(y, cc.Ti[k * 4 + 1], acc);
        acc = __fma_rn(d, cc.Ti[k * 4 + 2], acc);
        acc = __fma_rn(1.0, cc.Ti[k * 4 + 3], acc);
        o[k] = acc;
    }
}

// Validity of one pixel (cloud_x >= 0 & cloud_z < 1), bit-identical to the reference's and identical in the
// count and the write launch.  cloud_x and cloud_z are affine in the depth, cloud_k = d * A_k(col, row) + C_k, with
// A_k affine in (col, row): the host folds the calibration into eight doubles (CloudHost) and a pixel costs two FMAs
// per tested coordinate.  This re-associated value differs from the reference's chain by < 1e-10 for |d| < 1e5, so it
// decides whenever it is further than 1e-8 from its threshold; the rare borderline (or huge / non-finite) pixel runs
// the exact chain with IEEE divisions.
struct CloudHost {
    double rf_u, rf_v, b_x, b_y;
    double a0x, a0y, a0c, c0;          // cloud_x = d * (col * a0x + row * a0y + a0c) + c0
    double a2x, a2y, a2c, c2;          // cloud_z likewise
    // the same affine forms in fp32 with a rigorous error bound: a pixel whose |value| exceeds m * |d| + t is decided
    // by the fp32 sign (two FFMAs per coordinate, no float -> double conversion), every other pixel takes the fp64 test
    float fa0x, fa0y, fa0c, fc0, fm0, ft0;
    float fa2x, fa2y, fa2c, fc2, fm2, ft2;     // fc2 = c2 - 1: the test is cloud_z - 1 < 0
};

__device__ __forceinline__ bool cloud_valid(const CloudConst& cc, const CloudHost& h, int col, int row, double A0, double A2,
                                            float depth) {
    const double d = (double)depth;
    const double v0 = __fma_rn(d, A0, h.c0), v2 = __fma_rn(d, A2, h.c2);
    const bool safe = fabs(v0) > 1e-8 && fabs(v2 - 1.0) > 1e-8 && fabs(d) < 1e5;      // NaN fails every test
    if (safe) return v0 >= 0.0 && v2 < 1.0;
    double o[4];
    cloud_point<true>(cc, col, row, depth, o);
    return o[0] >= 0.0 && o[2] < 1.0;
}

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
    } else if (p0 + CL_ITEMS <= npx && (((img_off + p0) & 1) == 0)) {
        // an odd image of an H * W = 2 (mod 4) batch: 8-byte aligned
        const float2 q0 = __ldg(reinterpret_cast<const float2*>(depth + p0)), q1 = __ldg(reinterpret_cast<const float2*>(depth + p0 + 2));
        dv[0] = q0.x; dv[1] = q0.y; dv[2] = q1.x; dv[3] = q1.y;
    } else {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) dv[k] = (p0 + k < npx) ? __ldg(depth + p0 + k) : 0.0f;
    }
}

// the same with the image's alignment class decided once per warp (`al` = 4 / 2 / 1 floats: p0 is a multiple of 4, so
// the class depends on the image offset only) and a limit that also ends the tile
__device__ __forceinline__ void cloud_load_al(const float* depth, int al, int p0, int lim, float (&dv)[CL_ITEMS]) {
    if (p0 + CL_ITEMS <= lim) {
        if (al == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(depth + p0));
            dv[0] = q.x; dv[1] = q.y; dv[2] = q.z; dv[3] = q.w;
        } else if (al == 2) {
            const float2 q0 = __ldg(reinterpret_cast<const float2*>(depth + p0)), q1 = __ldg(reinterpret_cast<const float2*>(depth + p0 + 2));
            dv[0] = q0.x; dv[1] = q0.y; dv[2] = q1.x; dv[3] = q1.y;
        } else {
#pragma unroll
            for (int k = 0; k < CL_ITEMS; ++k) dv[k] = __ldg(depth + p0 + k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) dv[k] = (p0 + k < lim) ? __ldg(depth + p0 + k) : 0.0f;
    }
}

// validity bits of this thread's CL_ITEMS consecutive pixels
__device__ __forceinline__ unsigned cloud_mask(const CloudConst& cc, const CloudHost& h, int W, int p0, int npx,
                                               int row, int col, const float (&dv)[CL_ITEMS]) {
    unsigned vmask = 0;
    if (p0 + CL_ITEMS <= npx && col + CL_ITEMS <= W) {
        // all four pixels exist and share a row (all but ~0.4 % of the calls): no per-pixel bound or wrap logic.
        // fp32 first: |fp32 value - exact value| <= 4 u (Amax |d| + |c|) (inputs rounded once, two FMAs for A, one for the
        // value); the host passes m = 16 u Amax, t = 16 u |c| + 1e-6, so a value beyond m |d| + t has the exact sign and
        // lies further than the fp64 test's own margin from its threshold.  NaN / inf / |d| >= 1e5 fail the comparison.
        const float r0 = fmaf((float)row, h.fa0y, h.fa0c), r2 = fmaf((float)row, h.fa2y, h.fa2c);
        bool all_dec = true;
        unsigned fmask = 0;
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) {
            const float d = dv[k], ad = fabsf(d), xc = (float)(col + k);
            const float v0 = fmaf(d, fmaf(xc, h.fa0x, r0), h.fc0), v2 = fmaf(d, fmaf(xc, h.fa2x, r2), h.fc2);
            const bool dec = fabsf(v0) > fmaf(h.fm0, ad, h.ft0) && fabsf(v2) > fmaf(h.fm2, ad, h.ft2) && ad < 1.0e5f;
            all_dec = all_dec && dec;
            if (v0 >= 0.0f && v2 < 0.0f) fmask |= 1u << k;
        }
        if (all_dec) return fmask;
        double A0 = __fma_rn((double)col, h.a0x, __fma_rn((double)row, h.a0y, h.a0c));
        double A2 = __fma_rn((double)col, h.a2x, __fma_rn((double)row, h.a2y, h.a2c));
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) {
            if (cloud_valid(cc, h, col + k, row, A0, A2, dv[k])) vmask |= 1u << k;
            A0 += h.a0x; A2 += h.a2x;
        }
        return vmask;
    }
    double A0 = __fma_rn((double)col, h.a0x, __fma_rn((double)row, h.a0y, h.a0c));
    double A2 = __fma_rn((double)col, h.a2x, __fma_rn((double)row, h.a2y, h.a2c));
#pragma unroll
    for (int k = 0; k < CL_ITEMS; ++k) {
        if (p0 + k < npx && cloud_valid(cc, h, col, row, A0, A2, dv[k])) vmask |= 1u << k;
        if (++col == W) {
            col = 0; ++row;
            A0 = __fma_rn((double)row, h.a0y, h.a0c); A2 = __fma_rn((double)row, h.a2y, h.a2c);
        } else {
            A0 += h.a0x; A2 += h.a2x;
        }
    }
    return vmask;
}

// Workspace: int32 counts [B][tiles], then uint32 validity words [B][tiles][32]: bit 4 * it + k of word `lane` is
// pixel tile * 1024 + it * 128 + lane * 4 + k - i.e. the four pixels thread t = it * 32 + lane of launch 2 owns.
// (the per-image count rows are padded to a multiple of four tiles: 16-byte loads of the counts)
__host__ __device__ inline int cloud_tiles_pad(int tiles) { return (tiles + 3) & ~3; }
__host__ __device__ inline size_t cloud_words_offset(int B, int tiles) {
    return ((size_t)B * cloud_tiles_pad(tiles) * sizeof(int32_t) + 255) / 256 * 256;
}

// launch 1: validity bits and valid points per tile; one warp per tile
constexpr int CC_WARPS = 4;
__global__ void __launch_bounds__(CC_WARPS * 32)
cloud_count_kernel(const __grid_constant__ plb_cloud_args a, const CloudHost h, int tiles) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * CC_WARPS + (threadIdx.x >> 5);
    if (blk >= tiles) return;
    const int npx = a.H * a.W, W = a.W;
    const CloudConst cc = cloud_const(a, h);
    const float* depth = a.depth + (size_t)b * npx;
    const int pbase = blk * CL_TILE + lane * CL_ITEMS;
    // ONE copy of the mask code, iterated (an unrolled tile is ~1000 instructions = 16 KB that every warp walks once:
    // the 6 KB L0 instruction cache never hits).  The warp's 4 KB of depths are sent for in one instruction (one
    // 128-byte line per lane into L2); the loop's own 128-bit loads then pay an L2 hit that the other warps cover.
    constexpr int NIT = CL_TILE / 128;
    if (blk * CL_TILE + 32 * lane < npx) prefetch_l2(depth + blk * CL_TILE + 32 * lane);
    int row = pbase / W, col = pbase - row * W;
    unsigned word = 0;
    float cur[CL_ITEMS];
    const size_t img_off = (size_t)b * npx;
    const int al = (((uintptr_t)a.depth & 15) == 0 && (img_off & 3) == 0) ? 4 : (((uintptr_t)a.depth & 7) == 0 && (img_off & 1) == 0) ? 2 : 1;
    const int lim = min(npx, (blk + 1) * CL_TILE);            // the tile's end: zeros beyond it
    cloud_load_al(depth, al, pbase, lim, cur);
    int p0 = pbase;
#pragma unroll 1
    for (int it = 0; it < NIT; ++it, p0 += 128) {
        float nxt[CL_ITEMS];
        cloud_load_al(depth, al, p0 + 128, lim, nxt);
        if (p0 < npx) word |= cloud_mask(cc, h, W, p0, npx, row, col, cur) << (4 * it);
        col += 128;
        while (col >= W) { col -= W; ++row; }
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k) cur[k] = nxt[k];
    }
    int mine = __popc(word);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) ((int32_t*)a.workspace)[(size_t)b * cloud_tiles_pad(tiles) + blk] = mine;
    uint32_t* words = (uint32_t*)((char*)a.workspace + cloud_words_offset(a.B, tiles));
    words[((size_t)b * tiles + blk) * 32 + lane] = word;
}

// launch 2: every tile sums the counts of the tiles before it (a few hundred L2-resident integers: cheaper than a
// scan launch), compacts the (pixel, depth) pairs of its valid pixels in shared memory in rank order - validity
// from the same cheap test as launch 1 - and only then evaluates the points, one KEPT point per thread: the exact
// fp64 chain runs on the ~60 % of the pixels that survive (and, with [0::sparsity], on the kept ranks only),
// without divergence, and thread j writes output row j.
__global__ void __launch_bounds__(CL_THREADS)
cloud_write_kernel(const __grid_constant__ plb_cloud_args a, const CloudHost h) {
    const int b = blockIdx.y, blk = blockIdx.x, tiles = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npx = a.H * a.W;
    const CloudConst cc = cloud_const(a, h);
    const float* depth = a.depth + (size_t)b * npx;
    __shared__ int s_warp[CL_THREADS / 32], s_pre[CL_THREADS / 32];
    __shared__ int s_pix[CL_TILE];
    __shared__ float s_dep[CL_TILE];

    const int p0 = blk * CL_TILE + tid * CL_ITEMS;
    float dv[CL_ITEMS];
    cloud_load(depth, (size_t)b * npx, p0, npx, dv);
    // exclusive prefix of this tile: counts of tiles [0, blk) of image b, fixed order (integers)
    {
        const int32_t* counts = (const int32_t*)a.workspace + (size_t)b * cloud_tiles_pad(tiles);
        int part = 0;
        for (int k = tid; k < blk; k += CL_THREADS) part += __ldg(counts + k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_pre[warp] = part;
    }
    // this thread's four validity bits, as launch 1 decided them
    const uint32_t* words = (const uint32_t*)((const char*)a.workspace + cloud_words_offset(a.B, tiles));
    const unsigned vmask = (__ldg(words + ((size_t)b * tiles + blk) * 32 + lane) >> (4 * warp)) & 15u;
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
    int warp_off = 0, base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CL_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) warp_off += c;
        total += c;
        base += s_pre[w];
    }
    const int sp = a.sparsity > 0 ? a.sparsity : 1;
    if (blk == tiles - 1 && tid == 0 && a.count != nullptr) a.count[b] = (base + total + sp - 1) / sp;
    if (a.valid != nullptr) {
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k)
            if (p0 + k < npx) a.valid[(size_t)b * npx + p0 + k] = (vmask >> k) & 1u;
    }
    // kept points of this tile: local ranks r with (base + r) % sp == 0, i.e. r = r_first + j * sp
    int r_first = 0, n_keep = total;
    if (sp != 1) {
        r_first = (sp - base % sp) % sp;
        n_keep = total > r_first ? (total - r_first + sp - 1) / sp : 0;
    }
    if (n_keep == 0) return;
    {
        int r = warp_off + incl - mine;     // local rank of this thread's first valid point
#pragma unroll
        for (int k = 0; k < CL_ITEMS; ++k)
            if ((vmask >> k) & 1u) { s_pix[r] = p0 + k; s_dep[r] = dv[k]; ++r; }
    }
    __syncthreads();
    // the homogeneous row of T_inv is all zeros in the reference (zeros_like of a 4x4, PseudoLiDAR.py:43): the 4th
    // output column is then exactly 0 for every finite point
    const bool w_zero = a.Tinv[12] == 0.0 && a.Tinv[13] == 0.0 && a.Tinv[14] == 0.0 && a.Tinv[15] == 0.0;
    const int tile_row0 = (blk * CL_TILE) / a.W;
    const size_t out0 = (size_t)b * npx + (sp == 1 ? (size_t)base : (size_t)(base + r_first) / sp);   // first output row of the tile
    for (int j = tid; j < n_keep; j += CL_THREADS) {
        const int r = r_first + j * sp;
        const int pix = s_pix[r];
        // row of the pixel: a tile spans at most CL_TILE / W + 2 rows - a few compares instead of a division
        int row = tile_row0, col = pix - tile_row0 * a.W;
        while (col >= a.W) { col -= a.W; ++row; }
        const float d = s_dep[r];
        double o[4];
        // the reference's operation order with constant-divisor divisions
        if (w_zero) {
            cloud_point3(cc, col, row, d, o);
            o[3] = 0.0;
        } else {
            cloud_point<false>(cc, col, row, d, o);
        }
        // (an infinite depth can be a kept point under odd calibrations: the constant-divisor division turns it into NaN
        // where the IEEE division gives inf.  One fp32 compare instead of range checks on the fp64 results: for a
        // finite depth the quotients are finite and the correction step is exact)
        if (!(fabsf(d) < 1.0e30f)) cloud_point<true>(cc, col, row, d, o);
        const size_t pos = out0 + j;
        if (a.cloud_f64 != nullptr) {
            double2* q = reinterpret_cast<double2*>(a.cloud_f64 + pos * 4);
            __stcs(q, make_double2(o[0], o[1]));
            __stcs(q + 1, make_double2(o[2], o[3]));
        }
        if (a.cloud_f32 != nullptr)
            __stcs(reinterpret_cast<float4*>(a.cloud_f32) + pos, make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]));
        if (a.index != nullptr) a.index[pos] = pix;
    }
}

// launch 2 without decimation (sparsity 0 / 1, the reference's default): no compaction through shared memory, no
// block at all.  A warp owns 1024 / CD_WPT consecutive pixels and walks them 32 at a time, ONE pixel per lane, in a
// rolled loop (one copy of the fp64 chain in the instruction cache, its constants hoisted): the pixel's validity bit
// comes from the word of the lane that owned it in launch 1 (one shuffle), its output row is the tile's base (the counts
// of the tiles before it: 16-byte loads of the padded count row) + the valid pixels of the tile before the warp's chunk
// (one popc + warp reduction of the words) + its rank in the ballot.  Valid lanes run the exact chain in place and
// write rows that are consecutive across the warp.  No shared memory, no barrier, no per-point row / column
// reconstruction; depths are loaded CD_AHEAD steps ahead.
// v, as a value the compiler cannot trace back to the constant bank (blockIdx.z is 0: the grid is two-dimensional)
__device__ __forceinline__ double pin_reg(double v) { return __longlong_as_double(__double_as_longlong(v) ^ (long long)blockIdx.z); }
constexpr int CD_WPT = 1;                         // warps per 1024-pixel tile
constexpr int CD_WARPS = 4;                       // warps per block
#ifndef CD_MINB
#define CD_MINB 5                                 // resident blocks per SM the register allocation aims for
#endif
constexpr int CD_GROUPS = 8 / CD_WPT;             // 128-pixel groups of launch 1 per warp (4 validity bits of every word each)

// F64 / F32: which cloud layouts are written; AUX: the call also wants the index and / or the validity mask
template <bool F64, bool F32, bool AUX>
__global__ void __launch_bounds__(CD_WARPS * 32, CD_MINB)
cloud_write_direct_kernel(const __grid_constant__ plb_cloud_args a, const CloudHost h, int tiles) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * CD_WARPS + (threadIdx.x >> 5);
    const int blk = gw / CD_WPT, part = gw - blk * CD_WPT;
    if (blk >= tiles) return;
    const int npx = a.H * a.W, W = a.W;
    const int chunk = blk * CL_TILE + part * (128 * CD_GROUPS);
    const float* dp = a.depth + (size_t)b * npx + chunk + lane;          // this lane's pixel of the current step
    int p = chunk + lane;
    float ring[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) ring[s] = p + 32 * s < npx ? __ldg(dp + 32 * s) : 0.0f;
    // the rest of the warp's depths on their way into L2 (one 128-byte line per lane: the loads a group ahead of their
    // use then pay an L2 hit, not a DRAM round trip under load)
    if (chunk + 32 * lane < npx) prefetch_l2(a.depth + (size_t)b * npx + chunk + 32 * lane);
    const uint32_t* words = (const uint32_t*)((const char*)a.workspace + cloud_words_offset(a.B, tiles));
    const unsigned word = __ldg(words + ((size_t)b * tiles + blk) * 32 + lane);
    int base = 0;
    {
        const int4* c4 = reinterpret_cast<const int4*>((const int32_t*)a.workspace + (size_t)b * cloud_tiles_pad(tiles));
        for (int k = lane; 4 * k < blk; k += 32) {
            const int4 c = __ldg(c4 + k);
            const int left = blk - 4 * k;
            base += c.x + (left > 1 ? c.y : 0) + (left > 2 ? c.z : 0) + (left > 3 ? c.w : 0);
        }
        base = __reduce_add_sync(0xffffffffu, base);
    }
    const int below = __reduce_add_sync(0xffffffffu, __popc(word & ((1u << (4 * CD_GROUPS * part)) - 1u)));
    if (blk == tiles - 1 && part == CD_WPT - 1) {
        const int total = __reduce_add_sync(0xffffffffu, __popc(word));
        if (lane == 0 && a.count != nullptr) a.count[b] = base + total;
    }
    if (chunk >= npx) return;
    // the chain's constants live in registers for the whole loop (left to itself the compiler re-loads each one from
    // the constant bank inside the divergent block: ~14 uniform loads per step)
    CloudConst cc = cloud_const(a, h);
    cc.c_u = pin_reg(cc.c_u); cc.c_v = pin_reg(cc.c_v); cc.f_u = pin_reg(cc.f_u); cc.f_v = pin_reg(cc.f_v);
    cc.b_x = pin_reg(cc.b_x); cc.b_y = pin_reg(cc.b_y); cc.rf_u = pin_reg(cc.rf_u); cc.rf_v = pin_reg(cc.rf_v);
#pragma unroll
    for (int k = 0; k < 12; ++k) cc.Ti[k] = pin_reg(cc.Ti[k]);
    const bool w_zero = a.Tinv[12] == 0.0 && a.Tinv[13] == 0.0 && a.Tinv[14] == 0.0 && a.Tinv[15] == 0.0;
    unsigned mine = word >> (4 * CD_GROUPS * part);            // bits 4 g + k: pixel 128 g + 4 lane + k of the chunk
    const unsigned lt = (1u << lane) - 1u;
    const int src = lane >> 2;
    int orow = base + below;                                  // output row (within image b) of the step's first valid pixel
    double* const o64 = F64 ? a.cloud_f64 + (size_t)b * npx * 4 : nullptr;
    float4* const o32 = F32 ? reinterpret_cast<float4*>(a.cloud_f32) + (size_t)b * npx : nullptr;
    int32_t* const oidx = AUX && a.index != nullptr ? a.index + (size_t)b * npx : nullptr;
    uint8_t* const oval = AUX && a.valid != nullptr ? a.valid + (size_t)b * npx : nullptr;
    int row = chunk / W, col = chunk - row * W + lane;
#pragma unroll 1
    for (int g = 0; g < CD_GROUPS; ++g) {
        // the next group's four depths, all issued here: they land while this group's four steps run (the copy into
        // `ring` at the end of the group must not wait for a load issued one step earlier)
        float nxt[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) nxt[s] = (g + 1 < CD_GROUPS && p + 128 + 32 * s < npx) ? __ldg(dp + 128 + 32 * s) : 0.0f;
        // pixel p = chunk + 128 g + 32 s + lane of step s: its bit sits with lane 8 s + lane / 4 of launch 1
        bool odd = false;                                     // a depth the constant-divisor chain cannot take (inf / huge / NaN)
#pragma unroll
        for (int s = 0; s < 4; ++s) odd = odd || !(fabsf(ring[s]) < 1.0e30f);
        if (w_zero && !__any_sync(0xffffffffu, odd)) {
            // the common case, ONE basic block for the four steps: every lane runs the chain (the warp would anyway),
            // only the stores are predicated - the four independent chains interleave
            unsigned own[4], ball[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) own[s] = __shfl_sync(0xffffffffu, mine, 8 * s + src);      // four shuffles in flight
#pragma unroll
            for (int s = 0; s < 4; ++s) ball[s] = __ballot_sync(0xffffffffu, (own[s] >> (lane & 3)) & 1u);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const float d = ring[s];
                if (col >= W) { col -= W; ++row; }            // W >= 32 (host dispatch): one wrap per step at most
                const bool v = (ball[s] >> lane) & 1u;
                if (AUX && oval != nullptr && p < npx) oval[p] = v;
                const int r = orow + __popc(ball[s] & lt);
                double o[4];
                cloud_point3d(cc, int_as_double_exact(col), int_as_double_exact(row), (double)d, o);
                // PREDICATED stores, not a branch: behind `if (v)` the compiler sinks the whole chain into the divergent
                // block and the four steps run one after the other, each at the full latency of its dependent chain
                if (F64) st_cs_f64x4_if(v, o64 + (size_t)r * 4, o[0], o[1], o[2], 0.0);     // one 256-bit store: a whole sector
                if (F32) st_cs_f32x4_if(v, o32 + r, (float)o[0], (float)o[1], (float)o[2], 0.0f);
                if (AUX && oidx != nullptr && v) oidx[r] = p;
                orow = __shfl_sync(0xffffffffu, r + (v ? 1 : 0), 31);      // rows after this step (no second popc: XU pipe)
                col += 32; p += 32;
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {
                const float d = s == 0 ? ring[0] : s == 1 ? ring[1] : s == 2 ? ring[2] : ring[3];
                while (col >= W) { col -= W; ++row; }
                const bool v = (__shfl_sync(0xffffffffu, mine, 8 * s + src) >> (lane & 3)) & 1u;
                const unsigned ball = __ballot_sync(0xffffffffu, v);
                if (AUX && oval != nullptr && p < npx) oval[p] = v;
                if (v) {
                    const int r = orow + __popc(ball & lt);
                    double o[4];
                    cloud_point<false>(cc, col, row, d, o);
                    if (!(fabsf(d) < 1.0e30f)) cloud_point<true>(cc, col, row, d, o);      // see cloud_write_kernel
                    if (F64) {
                        double2* q = reinterpret_cast<double2*>(o64) + (size_t)r * 2;
                        __stcs(q, make_double2(o[0], o[1]));
                        __stcs(q + 1, make_double2(o[2], o[3]));
                    }
                    if (F32) __stcs(o32 + r, make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]));
                    if (AUX && oidx != nullptr) oidx[r] = p;
                }
                orow += __popc(ball);
                col += 32; p += 32;
            }
        }
        dp += 128;
        mine >>= 4;
#pragma unroll
        for (int s = 0; s < 4; ++s) ring[s] = nxt[s];
    }
}

static inline int cloud_blocks(const plb_cloud_args* a) { return (a->H * a->W + CL_TILE - 1) / CL_TILE; }

size_t cloud_workspace_bytes(const plb_cloud_args* a) {
    const int tiles = cloud_blocks(a);
    return cloud_words_offset(a->B, tiles) + ((size_t)a->B * tiles * 32 * sizeof(uint32_t) + 255) / 256 * 256;
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
    {
        // cloud_k = x T[k][0] + y T[k][1] + d T[k][2] + T[k][3],  x = (col - c_u) d / f_u + b_x,  y = (row - c_v) d / f_v + b_y
        const double c_u = a->P[2], c_v = a->P[6], f_u = a->P[0], f_v = a->P[5];
        const double* T0 = a->Tinv; const double* T2 = a->Tinv + 8;
        h.a0x = T0[0] / f_u; h.a0y = T0[1] / f_v; h.a0c = T0[2] - c_u * h.a0x - c_v * h.a0y;
        h.c0 = h.b_x * T0[0] + h.b_y * T0[1] + T0[3];
        h.a2x = T2[0] / f_u; h.a2y = T2[1] / f_v; h.a2c = T2[2] - c_u * h.a2x - c_v * h.a2y;
        h.c2 = h.b_x * T2[0] + h.b_y * T2[1] + T2[3];
        const double u = 1.0 / 16777216.0;          // 2^-24
        const double amax0 = fabs(h.a0x) * a->W + fabs(h.a0y) * a->H + fabs(h.a0c);
        const double amax2 = fabs(h.a2x) * a->W + fabs(h.a2y) * a->H + fabs(h.a2c);
        h.fa0x = (float)h.a0x; h.fa0y = (float)h.a0y; h.fa0c = (float)h.a0c; h.fc0 = (float)h.c0;
        h.fa2x = (float)h.a2x; h.fa2y = (float)h.a2y; h.fa2c = (float)h.a2c; h.fc2 = (float)(h.c2 - 1.0);
        h.fm0 = (float)(16.0 * u * amax0 * 1.0001); h.ft0 = (float)(16.0 * u * fabs(h.c0) + 1e-6);
        h.fm2 = (float)(16.0 * u * amax2 * 1.0001); h.ft2 = (float)(16.0 * u * fabs(h.c2 - 1.0) + 1e-6);
    }
    cloud_count_kernel<<<dim3((tiles + CC_WARPS - 1) / CC_WARPS, a->B), CC_WARPS * 32, 0, st>>>(*a, h, tiles);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    if (a->sparsity <= 1 && a->W >= 32) {
        const dim3 dgrid((tiles * CD_WPT + CD_WARPS - 1) / CD_WARPS, a->B);
        const bool f64 = a->cloud_f64 != nullptr, f32 = a->cloud_f32 != nullptr, aux = a->index != nullptr || a->valid != nullptr;
#define PLB_CLOUD_DIRECT(A, B_, C) cloud_write_direct_kernel<A, B_, C><<<dgrid, CD_WARPS * 32, 0, st>>>(*a, h, tiles)
        if (f64 && !f32 && !aux) PLB_CLOUD_DIRECT(true, false, false);
        else if (!f64 && f32 && !aux) PLB_CLOUD_DIRECT(false, true, false);
        else if (f64 && !f32) PLB_CLOUD_DIRECT(true, false, true);
        else if (!f64 && f32) PLB_CLOUD_DIRECT(false, true, true);
        else if (f64 && f32) PLB_CLOUD_DIRECT(true, true, true);
        else PLB_CLOUD_DIRECT(false, false, true);
#undef PLB_CLOUD_DIRECT
    }
    else cloud_write_kernel<<<grid, CL_THREADS, 0, st>>>(*a, h);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
