// Loader-side frame preparation on the GPU: the reference's transform chain
//   ToTensor -> ToPILImage -> Resize((H, W)) -> ToTensor -> Normalize(mean, std)        (trainer.py:97-103)
// applied by KittiDataset.load_img to `np.asarray(Image.open(path), float32) / 255.0` (dataloaders.py:32-49), and the
// intrinsics scaling of dataloaders.py:95-98 - uint8 HWC frames in, normalised planar fp32 (and optionally NHWC4) out.
//
// Bit-exact by construction:
//   * `img / 255.0` then ToPILImage's `mul(255).byte()` is a float32 round trip that TRUNCATES: v -> (uint8)(fl32(v / 255) * 255)
//     (a handful of values come back one lower) - evaluated with the same IEEE operations;
//   * Resize on a PIL image is Pillow's ImagingResample (bilinear, antialiased = support scaled by the reduction
//     factor): per axis a table of fixed-point coefficients (22 fractional bits) computed in double precision,
//     horizontal pass first, every pass rounded to uint8 through `(1 << 21) + sum >> 22` and clipped.  The tables are
//     built on the device with explicitly rounded fp64 operations in Pillow's operation order (prep_tables_kernel);
//   * ToTensor's `/ 255` and Normalize's `(x - mean) / std` are IEEE fp32 operations.
#include "common.cuh"

namespace plb {

constexpr int PR_KMAX = 32;                    // coefficients per output sample: ceil(scale) * 2 + 1 <= 32 (reduction <= 15x)
constexpr int PR_TAB = PR_KMAX + 2;            // table row: first input sample, count, coefficients
constexpr int PR_PRECISION_BITS = 32 - 8 - 2;  // Pillow's PRECISION_BITS for 8-bit channels

__host__ __device__ inline int prep_ksize(int in_size, int out_size) {
    // support = 1.0 * max(in / out, 1);  ksize = (int)ceil(support) * 2 + 1      (Resample.c: precompute_coeffs)
    double scale = (double)in_size / (double)out_size;
    if (scale < 1.0) scale = 1.0;
    int c = (int)scale;
    if ((double)c < scale) ++c;
    return c * 2 + 1;
}

// one thread per output sample of one axis: bounds + fixed-point coefficients, Pillow's arithmetic in fp64
__global__ void prep_tables_kernel(int in_w, int out_w, int in_h, int out_h, int* tab_x, int* tab_y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= out_w + out_h) return;
    const bool is_x = t < out_w;
    const int xx = is_x ? t : t - out_w;
    const int in_size = is_x ? in_w : in_h, out_size = is_x ? out_w : out_h;
    int* row = (is_x ? tab_x : tab_y) + (size_t)xx * PR_TAB;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                        // bilinear: support 1.0 * filterscale
    const double center = __dmul_rn((double)xx + 0.5, scale);  // in0 = 0
    const double ss = __ddiv_rn(1.0, filterscale);
    int xmin = (int)__dadd_rn(__dadd_rn(center, -support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > PR_KMAX) xmax = PR_KMAX;                        // (validated on the host: cannot happen)
    double k[PR_KMAX];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double arg = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss);
        if (arg < 0.0) arg = -arg;
        const double w = arg < 1.0 ? __dadd_rn(1.0, -arg) : 0.0;
        k[x] = w;
        ww = __dadd_rn(ww, w);
    }
    row[0] = xmin;
    row[1] = xmax;
    for (int x = 0; x < PR_KMAX; ++x) {
        int c = 0;
        if (x < xmax) {
            const double v = ww != 0.0 ? __ddiv_rn(k[x], ww) : k[x];
            const double s = __dmul_rn(v, (double)(1 << PR_PRECISION_BITS));
            c = v < 0.0 ? (int)__dadd_rn(-0.5, s) : (int)__dadd_rn(0.5, s);
        }
        row[2 + x] = c;
    }
}

__device__ __forceinline__ int clip8(int v) {               // Pillow's clip8 lookup: saturate after the shift
    v >>= PR_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// `np.asarray(img, float32) / 255.0` -> ToTensor (no scaling for float input) -> ToPILImage: `mul(255).byte()`
__device__ __forceinline__ int roundtrip_u8(int v) {
    return (int)__fmul_rn(__fdiv_rn((float)v, 255.0f), 255.0f);
}

// One block = one tile of PR_TW x PR_TH output pixels of one frame: the input footprint of the tile is staged in
// shared memory (after the uint8 round trip), the horizontal pass leaves its uint8 rows there, the vertical pass
// produces the output pixels, which are normalised and written planar (coalesced along x) and / or NHWC4.
constexpr int PR_TW = 32, PR_TH = 8, PR_THREADS = 256;

struct PrepLaunch {
    plb_prep_args a;
    const int* tab_x;
    const int* tab_y;
    int tiles_x, tiles_y;
    int max_fw, max_fh;            // largest input footprint of a tile (columns, rows)
};

__global__ void __launch_bounds__(PR_THREADS)
prep_resize_kernel(const __grid_constant__ PrepLaunch p) {
    const plb_prep_args& a = p.a;
    extern __shared__ unsigned char pr_smem[];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int ox0 = tx * PR_TW, oy0 = ty * PR_TH;
    const int nx = min(PR_TW, a.W - ox0), ny = min(PR_TH, a.H - oy0);
    // input footprint of the tile
    const int* rx0 = p.tab_x + (size_t)ox0 * PR_TAB;
    const int* rx1 = p.tab_x + (size_t)(ox0 + nx - 1) * PR_TAB;
    const int* ry0 = p.tab_y + (size_t)oy0 * PR_TAB;
    const int* ry1 = p.tab_y + (size_t)(oy0 + ny - 1) * PR_TAB;
    const int fx0 = rx0[0], fx1 = rx1[0] + rx1[1];           // [fx0, fx1) input columns (tables are monotone)
    const int fy0 = ry0[0], fy1 = ry1[0] + ry1[1];
    const int fw = fx1 - fx0, fh = fy1 - fy0;
    unsigned char* s_in = pr_smem;                                     // [fh][fw][3]
    unsigned char* s_h = pr_smem + (size_t)p.max_fh * p.max_fw * 3;    // [fh][PR_TW][3]
    const unsigned char* frame = a.frames + (size_t)b * a.in_h * a.in_w * 3;
    const int row_bytes = fw * 3;
    for (int k = tid; k < fh * row_bytes; k += PR_THREADS) {
        const int r = k / row_bytes, c = k - r * row_bytes;
        s_in[r * row_bytes + c] = (unsigned char)roundtrip_u8(frame[((size_t)(fy0 + r) * a.in_w + fx0) * 3 + c]);
    }
    __syncthreads();
    // horizontal pass (only when the width changes: Resample.c skips a pass whose size is unchanged)
    const bool need_h = a.in_w != a.W, need_v = a.in_h != a.H;
    for (int k = tid; k < fh * nx * 3; k += PR_THREADS) {
        const int r = k / (nx * 3), q = k - r * (nx * 3);
        const int xo = q / 3, c = q - xo * 3;
        int v;
        if (need_h) {
            const int* row = p.tab_x + (size_t)(ox0 + xo) * PR_TAB;
            const int xmin = row[0] - fx0, n = row[1];
            int ss = 1 << (PR_PRECISION_BITS - 1);
            for (int x = 0; x < n; ++x) ss += (int)s_in[r * row_bytes + (xmin + x) * 3 + c] * row[2 + x];
            v = clip8(ss);
        } else {
            v = s_in[r * row_bytes + (ox0 + xo - fx0) * 3 + c];
        }
        s_h[(r * PR_TW + xo) * 3 + c] = (unsigned char)v;
    }
    __syncthreads();
    // vertical pass + ToTensor (/255) + Normalize
    for (int k = tid; k < ny * nx; k += PR_THREADS) {
        const int yo = k / nx, xo = k - yo * nx;
        const int* row = p.tab_y + (size_t)(oy0 + yo) * PR_TAB;
        const int ymin = row[0] - fy0, n = row[1];
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v;
            if (need_v) {
                int ss = 1 << (PR_PRECISION_BITS - 1);
                for (int y = 0; y < n; ++y) ss += (int)s_h[((ymin + y) * PR_TW + xo) * 3 + c] * row[2 + y];
                v = clip8(ss);
            } else {
                v = s_h[((oy0 + yo - fy0) * PR_TW + xo) * 3 + c];
            }
            const float f = __fdiv_rn((float)v, 255.0f);                       // ToTensor
            o[c] = __fdiv_rn(__fsub_rn(f, a.mean[c]), a.stdev[c]);               // Normalize: sub_(mean).div_(std)
        }
        const int X = ox0 + xo, Y = oy0 + yo;
        if (a.out_planar != nullptr) {
            float* q = a.out_planar + ((size_t)b * 3 * a.H + Y) * a.W + X;
            q[0] = o[0]; q[(size_t)a.H * a.W] = o[1]; q[2 * (size_t)a.H * a.W] = o[2];
        }
        if (a.out_nhwc4 != nullptr)
            reinterpret_cast<float4*>(a.out_nhwc4)[((size_t)b * a.H + Y) * a.W + X] = make_float4(o[0], o[1], o[2], 0.0f);
    }
    // intrinsics (dataloaders.py:95-98): K[0] *= W / og_w, K[1] *= H / og_h, fp64 like numpy
    if (a.K_in != nullptr && a.K_out != nullptr && tile == 0 && tid < 9 && b < (a.n_K > 0 ? a.n_K : a.B)) {
        const double sx = __ddiv_rn((double)a.W, (double)a.in_w), sy = __ddiv_rn((double)a.H, (double)a.in_h);
        const double v = a.K_in[(size_t)b * 9 + tid];
        a.K_out[(size_t)b * 9 + tid] = tid < 3 ? __dmul_rn(v, sx) : (tid < 6 ? __dmul_rn(v, sy) : v);
    }
}

static int validate_prep(const plb_prep_args* a) {
    if (a == nullptr) return PLB_ENULL;
    if (a->B < 1 || a->in_h < 1 || a->in_w < 1 || a->H < 1 || a->W < 1) return PLB_EINVAL;
    if (a->frames == nullptr || (a->out_planar == nullptr && a->out_nhwc4 == nullptr)) return PLB_ENULL;
    if (prep_ksize(a->in_w, a->W) > PR_KMAX || prep_ksize(a->in_h, a->H) > PR_KMAX) return PLB_EINVAL;
    if ((a->K_in == nullptr) != (a->K_out == nullptr)) return PLB_ENULL;
    if (a->n_K < 0 || a->n_K > a->B) return PLB_EINVAL;
    if (a->B > 65535) return PLB_EINVAL;
    if (a->workspace == nullptr) return PLB_EWORKSPACE;
    return PLB_OK;
}

size_t prep_workspace_bytes(const plb_prep_args* a) {
    return (sizeof(int) * (size_t)(a->W + a->H) * PR_TAB + 255) / 256 * 256;
}

int prep_launch(const plb_prep_args* a, cudaStream_t st) {
    int rc = validate_prep(a);
    if (rc != PLB_OK) return rc;
    if (a->workspace_bytes < prep_workspace_bytes(a)) return PLB_EWORKSPACE;
    int* tab_x = (int*)a->workspace;
    int* tab_y = tab_x + (size_t)a->W * PR_TAB;
    const int n = a->W + a->H;
    prep_tables_kernel<<<(n + 127) / 128, 128, 0, st>>>(a->in_w, a->W, a->in_h, a->H, tab_x, tab_y);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    PrepLaunch p;
    p.a = *a;
    p.tab_x = tab_x; p.tab_y = tab_y;
    p.tiles_x = (a->W + PR_TW - 1) / PR_TW;
    p.tiles_y = (a->H + PR_TH - 1) / PR_TH;
    // footprint bound of a tile: (tile - 1) * scale + 2 * support + 2, support = max(scale, 1)
    auto bound = [](int in_size, int out_size, int tile) {
        const double scale = (double)in_size / (double)out_size;
        const double support = scale < 1.0 ? 1.0 : scale;
        int v = (int)((tile - 1) * scale + 2.0 * support + 3.0);
        return v > in_size ? in_size : v;
    };
    p.max_fw = bound(a->in_w, a->W, PR_TW);
    p.max_fh = bound(a->in_h, a->H, PR_TH);
    const size_t smem = (size_t)p.max_fh * p.max_fw * 3 + (size_t)p.max_fh * PR_TW * 3;
    if (smem > 200 * 1024) return PLB_EINVAL;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(prep_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    prep_resize_kernel<<<dim3(p.tiles_x * p.tiles_y, a->B), PR_THREADS, smem, st>>>(p);
    ++g_launches;
    PLB_CHECK_LAUNCH();
    return PLB_OK;
}

}  // namespace plb
