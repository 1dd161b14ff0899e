// clipk - C ABI implementation (see include/clipk.h).  sm_100a only; no CPU path, no other backend.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/clipk.h"
#include "gemm_core.cuh"

namespace clipk {

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail(int(e__), "%s: %s", #expr, cudaGetErrorString(e__));  \
    } while (0)

// ------------------------------------------------------------------------------------------------ device info
struct DevInfo {
    int ok = 0;       // 1 = CC 10.x
    int sms = 0;
    int queried = 0;
};
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static int device_info(DevInfo* out) {
    int dev = 0;
    CK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CLIPK_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[dev].queried) {
        int major = 0, sms = 0;
        CK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        CK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        g_dev[dev].ok = (major == 10);
        g_dev[dev].sms = sms;
        g_dev[dev].queried = 1;
    }
    *out = g_dev[dev];
    if (!out->ok) return fail(CLIPK_EARCH, "device %d is not compute capability 10.x (B200 / sm_100a required)", dev);
    return CLIPK_OK;
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn encode_fn() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return g_encode;
}

// 2D bf16 tensor map, SWIZZLE_128B, zero fill out of bounds.  inner = contiguous extent (elements).
static int make_tmap_bf16(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld_elems,
                          int box_inner, int box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CLIPK_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(CLIPK_EINVAL, "operand pointer not 16-byte aligned");
    if ((ld_elems * 2) % 16 != 0) return fail(CLIPK_EINVAL, "leading dimension %lld not a multiple of 8 elements", ld_elems);
    cuuint64_t dims[2] = {cuuint64_t(inner), cuuint64_t(outer)};
    cuuint64_t strides[1] = {cuuint64_t(ld_elems) * 2};
    cuuint32_t box[2] = {cuuint32_t(box_inner), cuuint32_t(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CLIPK_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return CLIPK_OK;
}
// operand [rows, K] row-major, consumed K-major: box = 64 (K) x box_rows
static int tmap_kmajor(CUtensorMap* m, const void* base, long long rows, long long K, long long ld, int box_rows) {
    return make_tmap_bf16(m, base, K, rows, ld, BK, box_rows);
}
// operand stored [K, mn] row-major (mn contiguous), consumed MN-major: box = 64 (mn) x 64 (K)
static int tmap_mnmajor(CUtensorMap* m, const void* base, long long mn, long long K, long long ld) {
    return make_tmap_bf16(m, base, mn, K, ld, 64, BK);
}

// ------------------------------------------------------------------------------------------------ small kernels
// merge the per-unit partial row statistics (log2-scaled domain) into natural-log (max, sum).
__global__ void merge_row_parts_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                       int nparts, int rows, float* __restrict__ row_max, float* __restrict__ row_sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float m = -CUDART_INF_F;
    for (int p = 0; p < nparts; ++p) m = fmaxf(m, part_max[(size_t)p * rows + i]);
    float l = 0.f;
    for (int p = 0; p < nparts; ++p) {
        const float pm = part_max[(size_t)p * rows + i];
        if (pm > -CUDART_INF_F) l += part_sum[(size_t)p * rows + i] * exp2f(pm - m);
    }
    row_max[i] = m * LN2;
    row_sum[i] = l;
}

__device__ __forceinline__ float merge_col(const float* __restrict__ cmax, const float* __restrict__ csum, int nparts,
                                           long long stride, long long j) {
    float m = -CUDART_INF_F;
    for (int p = 0; p < nparts; ++p) m = fmaxf(m, cmax[(size_t)p * stride + j]);
    float l = 0.f;
    for (int p = 0; p < nparts; ++p) {
        const float pm = cmax[(size_t)p * stride + j];
        if (pm > -CUDART_INF_F) l += csum[(size_t)p * stride + j] * expf(pm - m);
    }
    return m + logf(l);
}

__global__ void finalize_kernel(const float* __restrict__ row_max, const float* __restrict__ row_sum,
                                const float* __restrict__ pos, int rows, const float* __restrict__ cmax,
                                const float* __restrict__ csum, int nparts, long long stride, int cols,
                                long long diag_offset,
                                float* __restrict__ lse_row, float* __restrict__ lse_col, float* __restrict__ loss_sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float a = 0.f, b = 0.f;
    if (i < cols) lse_col[i] = merge_col(cmax, csum, nparts, stride, i);
    if (i < rows) {
        const float lr = row_max[i] + logf(row_sum[i]);
        lse_row[i] = lr;
        const float p = pos[i];
        a = lr - p;
        const long long j = diag_offset + i;
        if (j >= 0 && j < cols) b = merge_col(cmax, csum, nparts, stride, j) - p;
    }
    __shared__ float sa[32], sb[32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = (l < (blockDim.x >> 5)) ? sa[l] : 0.f;
        b = (l < (blockDim.x >> 5)) ? sb[l] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        if (l == 0) {
            atomicAdd(loss_sums + 0, a);
            atomicAdd(loss_sums + 1, b);
        }
    }
}

__global__ void cast_kernel(const float* __restrict__ src, void* __restrict__ dst, long long n, int dtype) {
    const long long i = (long long)(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long base = i * 4;
    if (base >= n) return;
    if (base + 4 <= n && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(src + base);
        if (dtype == CLIPK_BF16) {
            uint2 o = make_uint2(ptx::pack_bf16x2(v.x, v.y), ptx::pack_bf16x2(v.z, v.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + base) = o;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + base) = v;
        }
    } else {
        for (long long k = base; k < n && k < base + 4; ++k) {
            if (dtype == CLIPK_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[k] = __float2bfloat16(src[k]);
            else reinterpret_cast<float*>(dst)[k] = src[k];
        }
    }
}

// |x| maximum of a [rows, d] matrix as the bit pattern of a non-negative float (monotonic under integer max).
template <typename T>
__global__ void amax_kernel(const T* __restrict__ src, long long rows, long long d, long long ld,
                            unsigned int* __restrict__ amax_bits) {
    float m = 0.f;
    const long long n = rows * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d, k = i - r * d;
        const float v = fabsf(float(src[r * ld + k]));
        if (v < CUDART_INF_F) m = fmaxf(m, v);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

// [rows, d] bf16 / fp32  ->  scaled fp16 planes laid side by side (plane width dpad, zero padded):
//   planes = 1:  dst = fp16(x * 2^e)                      exact for bf16 sources (8 significant bits fit in 11)
//   planes = 2:  dst = [hi | lo], hi + lo = x * 2^e to 22 bits
// 2^e maps the largest |x| into [2^13, 2^14), so entries down to 2^-28 of the maximum stay normal fp16 numbers.
// scale_io[0] holds the amax bits on entry; scale_io[1] receives 2^-e (true value = stored * scale_io[1]).
template <typename T>
__global__ void to_f16_kernel(const T* __restrict__ src, __half* __restrict__ dst, long long rows, long long d,
                              long long ld_src, long long dpad, int planes, float* __restrict__ scale_io) {
    const float amax = scale_io[0];
    int ex = 0;
    float scale = 1.f;
    if (amax > 0.f && amax < CUDART_INF_F) {
        frexpf(amax, &ex);                 // amax = m * 2^ex, m in [0.5, 1)
        int e = 14 - ex;
        e = e > 120 ? 120 : (e < -120 ? -120 : e);
        scale = ldexpf(1.f, e);
    }
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) scale_io[1] = 1.f / scale;
    if (idx >= rows * dpad) return;
    const long long r = idx / dpad, k = idx - r * dpad;
    const float x = (k < d) ? float(src[r * ld_src + k]) * scale : 0.f;
    __half* o = dst + r * (planes * dpad) + k;
    const __half hi = __float2half_rn(x);
    o[0] = hi;
    if (planes == 2) o[dpad] = __float2half_rn(x - __half2float(hi));
}

// ------------------------------------------------------------------------------------------------ launch helpers
static inline int cdiv(long long a, long long b) { return int((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

constexpr int MAX_SPLIT = 32;
constexpr long long PANEL_ROWS = 4096;
constexpr long long PANEL_COLS = 4096;

// How many CTAs share the column sweep of one 128-row block: fill the SMs in as few equal waves as possible.
static int choose_split(int m_blocks, int n_tiles, int sms) {
    int best = 1;
    double best_cost = 1e30;
    const int lim = n_tiles < MAX_SPLIT ? n_tiles : MAX_SPLIT;
    for (int s = 1; s <= lim; ++s) {
        const int per = cdiv(n_tiles, s);
        const int waves = cdiv((long long)m_blocks * s, sms);
        const double cost = waves * (per + 0.35);   // 0.35 tile-times of prologue/epilogue per CTA
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    return best;
}

template <int MODE, int A_MN, int B_MN, int F16>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const KArgs& a, dim3 grid, cudaStream_t st) {
    auto kfn = gemm_kernel<MODE, A_MN, B_MN, F16>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); });
    if (attr_err != cudaSuccess) return fail(int(attr_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    kfn<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ta, tb, a);
    CK_CUDA(cudaGetLastError());
    return CLIPK_OK;
}

// plane pairs of the split-precision product, SMALLEST terms first: the tensor core truncates when it adds into the
// fp32 accumulator, so the big hi.hi products go last, when only d/16 more additions can bias the sum.
static const int kPairA[3] = {1, 0, 0};
static const int kPairB[3] = {0, 1, 0};

static inline int planes_of(int dtype) { return dtype == CLIPK_F16X2 ? 2 : 1; }
static inline bool is_f16(int dtype) { return dtype == CLIPK_F16 || dtype == CLIPK_F16X2; }

// one K segment for single-plane operands; three for two-plane ones (planes `a_plane` / `b_plane` inner elements apart)
static void set_segments(KArgs& a, int planes, int k_extent, long long a_plane, long long b_plane) {
    a.kb_per_seg = cdiv(k_extent, BK);
    a.nseg = (planes == 2) ? 3 : 1;
    a.num_kb = a.nseg * a.kb_per_seg;
    for (int i = 0; i < 6; ++i) { a.a_off[i] = 0; a.b_off[i] = 0; }
    if (planes == 2)
        for (int i = 0; i < 3; ++i) {
            a.a_off[i] = int(kPairA[i] * a_plane);
            a.b_off[i] = int(kPairB[i] * b_plane);
        }
}

static int check_common(const void* X, const void* Y, int rows, int cols, int d, long long ldx, long long ldy, int dtype) {
    if (!X || !Y) return fail(CLIPK_EINVAL, "null operand pointer");
    if (rows <= 0 || cols <= 0 || d <= 0) return fail(CLIPK_EINVAL, "rows, cols and d must be positive (got %d, %d, %d)", rows, cols, d);
    if (dtype != CLIPK_BF16 && dtype != CLIPK_F16 && dtype != CLIPK_F16X2)
        return fail(CLIPK_EUNSUPPORTED, "dtype %d: operands must be CLIPK_BF16, CLIPK_F16 or CLIPK_F16X2 (fp32 inputs go through clipk_to_f16 first)", dtype);
    const long long need = (dtype == CLIPK_BF16) ? d : planes_of(dtype) * round_up(d, BK);
    if (ldx < need || ldy < need) return fail(CLIPK_EINVAL, "leading dimension smaller than the row width (%lld)", need);
    if (d % 8 != 0) return fail(CLIPK_EUNSUPPORTED, "d = %d is not a multiple of 8", d);
    return CLIPK_OK;
}

}  // namespace clipk

using namespace clipk;

// ================================================================================================ C ABI
extern "C" {

int clipk_version(void) { return CLIPK_VERSION; }
const char* clipk_last_error(void) { return g_err; }

int clipk_check_device(void) {
    DevInfo di;
    return device_info(&di);
}

int clipk_to_f16(const void* src, int src_dtype, long long rows, long long d, long long ld_src, void* dst, int planes,
                 long long ld_dst, float* scale_io, void* stream) {
    if (!src || !dst || !scale_io || rows <= 0 || d <= 0) return fail(CLIPK_EINVAL, "bad argument");
    if (planes != 1 && planes != 2) return fail(CLIPK_EINVAL, "planes must be 1 or 2");
    if (src_dtype != CLIPK_BF16 && src_dtype != CLIPK_F32) return fail(CLIPK_EUNSUPPORTED, "source dtype %d", src_dtype);
    const long long dpad = round_up(d, BK);
    if (ld_src < d || ld_dst != planes * dpad) return fail(CLIPK_EINVAL, "ld_dst must be planes * round_up(d, 64) = %lld", planes * dpad);
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK_CUDA(cudaMemsetAsync(scale_io, 0, 2 * sizeof(float), st));
    const long long n = rows * dpad;
    const int rblocks = int(n / 1024 < 1 ? 1 : (n / 1024 > 4 * di.sms ? 4 * di.sms : n / 1024));
    unsigned int* bits = reinterpret_cast<unsigned int*>(scale_io);
    __half* out = static_cast<__half*>(dst);
    if (src_dtype == CLIPK_BF16) {
        const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(src);
        amax_kernel<<<rblocks, 256, 0, st>>>(p, rows, d, ld_src, bits);
        to_f16_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, out, rows, d, ld_src, dpad, planes, scale_io);
    } else {
        const float* p = static_cast<const float*>(src);
        amax_kernel<<<rblocks, 256, 0, st>>>(p, rows, d, ld_src, bits);
        to_f16_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, out, rows, d, ld_src, dpad, planes, scale_io);
    }
    CK_CUDA(cudaGetLastError());
    return CLIPK_OK;
}

size_t clipk_fwd_workspace_bytes(int rows, int cols, int d, int dtype) {
    (void)cols; (void)d; (void)dtype;
    if (rows <= 0) return 0;
    return size_t(2) * MAX_SPLIT * size_t(rows) * sizeof(float) + 256;
}

int clipk_fwd_stats(const void* X, const void* Y, int rows, int cols, int d, long long ldx, long long ldy, int dtype,
                    const float* x_inv_scale, const float* y_inv_scale, const float* logit_scale,
                    long long diag_offset, float* row_max, float* row_sum, float* pos_logit, void* workspace,
                    size_t workspace_bytes, void* stream) {
    int rc = check_common(X, Y, rows, cols, d, ldx, ldy, dtype);
    if (rc) return rc;
    if (!logit_scale || !row_max || !row_sum || !workspace) return fail(CLIPK_EINVAL, "null pointer argument");
    if (workspace_bytes < clipk_fwd_workspace_bytes(rows, cols, d, dtype)) return fail(CLIPK_EWORKSPACE, "workspace too small");
    DevInfo di;
    if ((rc = device_info(&di))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int planes = planes_of(dtype);
    const long long dpad = round_up(d, BK);
    const long long kext = (dtype == CLIPK_BF16) ? d : planes * dpad;
    CUtensorMap ta, tb;
    if ((rc = tmap_kmajor(&ta, X, rows, kext, ldx, BM))) return rc;
    if ((rc = tmap_kmajor(&tb, Y, cols, kext, ldy, BN))) return rc;
    KArgs a{};
    a.M = rows; a.N = cols; a.n_tiles = cdiv(cols, BN);
    set_segments(a, planes, d, dpad, dpad);
    const int m_blocks = cdiv(rows, BM);
    const int split = choose_split(m_blocks, a.n_tiles, di.sms);
    a.tiles_per_unit = cdiv(a.n_tiles, split);
    const int units = cdiv(a.n_tiles, a.tiles_per_unit);
    a.scale = logit_scale; a.xs = x_inv_scale; a.ys = y_inv_scale; a.diag_offset = diag_offset;
    a.part_max = static_cast<float*>(workspace);
    a.part_sum = a.part_max + size_t(MAX_SPLIT) * rows;
    a.pos = pos_logit;
    if (!pos_logit) a.diag_offset = -(1LL << 40);   // no row has a positive inside [0, cols)
    if (is_f16(dtype)) rc = launch_gemm<MODE_STATS, 0, 0, 1>(ta, tb, a, dim3(units, m_blocks), st);
    else rc = launch_gemm<MODE_STATS, 0, 0, 0>(ta, tb, a, dim3(units, m_blocks), st);
    if (rc) return rc;
    merge_row_parts_kernel<<<cdiv(rows, 256), 256, 0, st>>>(a.part_max, a.part_sum, units, rows, row_max, row_sum);
    CK_CUDA(cudaGetLastError());
    return CLIPK_OK;
}

int clipk_finalize(const float* row_max, const float* row_sum, const float* pos_logit, int rows,
                   const float* col_max_parts, const float* col_sum_parts, int nparts, long long part_stride,
                   int cols, long long diag_offset, float* lse_row, float* lse_col, float* loss_sums, void* stream) {
    if (!row_max || !row_sum || !pos_logit || !col_max_parts || !col_sum_parts || !lse_row || !lse_col || !loss_sums)
        return fail(CLIPK_EINVAL, "null pointer argument");
    if (rows <= 0 || cols <= 0 || nparts <= 0) return fail(CLIPK_EINVAL, "rows, cols and nparts must be positive");
    if (part_stride < cols) return fail(CLIPK_EINVAL, "part_stride smaller than cols");
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK_CUDA(cudaMemsetAsync(loss_sums, 0, 2 * sizeof(float), st));
    const int n = rows > cols ? rows : cols;
    finalize_kernel<<<cdiv(n, 256), 256, 0, st>>>(row_max, row_sum, pos_logit, rows, col_max_parts, col_sum_parts, nparts,
                                                  part_stride, cols, diag_offset, lse_row, lse_col, loss_sums);
    CK_CUDA(cudaGetLastError());
    return CLIPK_OK;
}

size_t clipk_bwd_workspace_bytes(int rows, int cols, int d, int g_dtype) {
    (void)d;
    if (rows <= 0 || cols <= 0) return 0;
    const long long rp = round_up(rows < PANEL_ROWS ? rows : PANEL_ROWS, BM);
    const long long cp = round_up(cols < PANEL_COLS ? cols : PANEL_COLS, BN);
    return size_t(rp) * size_t(cp) * 2 * planes_of(g_dtype) + 1024;
}

int clipk_bwd(const void* X, const void* Y, int rows, int cols, int d, long long ldx, long long ldy, int dtype,
              const float* x_inv_scale, const float* y_inv_scale, const void* Xg, const void* Yg, long long ldxg,
              long long ldyg, int g_dtype, const float* xg_inv_scale, const float* yg_inv_scale,
              const float* logit_scale, long long diag_offset, const float* lse_row, const float* lse_col,
              float alpha, float beta, const float* gscale, float* dX_acc, float* dY_acc, float* ds_acc,
              float* ds_col, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(X, Y, rows, cols, d, ldx, ldy, dtype);
    if (rc) return rc;
    if (g_dtype != CLIPK_F16 && g_dtype != CLIPK_F16X2) return fail(CLIPK_EUNSUPPORTED, "g_dtype must be CLIPK_F16 or CLIPK_F16X2");
    if ((rc = check_common(Xg, Yg, rows, cols, d, ldxg, ldyg, g_dtype))) return rc;
    if (!logit_scale || !lse_row || !lse_col || !gscale || !ds_acc || !workspace) return fail(CLIPK_EINVAL, "null pointer argument");
    if (workspace_bytes < clipk_bwd_workspace_bytes(rows, cols, d, g_dtype)) return fail(CLIPK_EWORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(CLIPK_EINVAL, "workspace must be 256-byte aligned");
    DevInfo di;
    if ((rc = device_info(&di))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int planes = planes_of(dtype);                // S-GEMM operand planes
    const int gplanes = planes_of(g_dtype);             // planes of G and of the gradient-GEMM features
    const long long dpad = round_up(d, BK);
    const long long kext = (dtype == CLIPK_BF16) ? d : planes * dpad;   // inner extent of X / Y rows
    const long long gext = gplanes * dpad;                              // inner extent of Xg / Yg rows
    const long long rp_max = rows < PANEL_ROWS ? rows : PANEL_ROWS;
    const long long cp_max = cols < PANEL_COLS ? cols : PANEL_COLS;
    const int ncp = int(round_up(cp_max, BN));          // padded panel width = G plane stride
    const int ldg = gplanes * ncp;
    __half* G = static_cast<__half*>(workspace);
    const size_t esz = 2;
    const char* Xb = static_cast<const char*>(X);
    const char* Yb = static_cast<const char*>(Y);
    const char* Xgb = static_cast<const char*>(Xg);
    const char* Ygb = static_cast<const char*>(Yg);

    CK_CUDA(cudaMemsetAsync(ds_acc, 0, 2 * sizeof(float), st));
    if (ds_col) CK_CUDA(cudaMemsetAsync(ds_col, 0, size_t(cols) * sizeof(float), st));

    for (long long r0 = 0; r0 < rows; r0 += rp_max) {
        const int nr = int(rows - r0 < rp_max ? rows - r0 : rp_max);
        for (long long c0 = 0; c0 < cols; c0 += cp_max) {
            const int nc = int(cols - c0 < cp_max ? cols - c0 : cp_max);
            // ---- A: recompute S on the panel, write G (fp16, x 2^14) and the dlogit_scale sums
            {
                CUtensorMap ta, tb;
                if ((rc = tmap_kmajor(&ta, Xb + r0 * ldx * esz, nr, kext, ldx, BM))) return rc;
                if ((rc = tmap_kmajor(&tb, Yb + c0 * ldy * esz, nc, kext, ldy, BN))) return rc;
                KArgs a{};
                a.M = nr; a.N = nc; a.n_tiles = cdiv(nc, BN);
                set_segments(a, planes, d, dpad, dpad);
                const int m_blocks = cdiv(nr, BM);
                const int split = choose_split(m_blocks, a.n_tiles, di.sms);
                a.tiles_per_unit = cdiv(a.n_tiles, split);
                const int units = cdiv(a.n_tiles, a.tiles_per_unit);
                a.scale = logit_scale; a.xs = x_inv_scale; a.ys = y_inv_scale;
                a.diag_offset = diag_offset + r0 - c0;
                a.lse_row = lse_row + r0; a.lse_col = lse_col + c0;
                a.alpha = alpha; a.beta = beta; a.gscale = gscale;
                a.G = G; a.ldg = ldg; a.g_planes = gplanes; a.g_plane_stride = ncp;
                a.ds_acc = ds_acc; a.ds_col = ds_col ? ds_col + c0 : nullptr;
                if (is_f16(dtype)) rc = launch_gemm<MODE_GRAD, 0, 0, 1>(ta, tb, a, dim3(units, m_blocks), st);
                else rc = launch_gemm<MODE_GRAD, 0, 0, 0>(ta, tb, a, dim3(units, m_blocks), st);
                if (rc) return rc;
            }
            // ---- B: dX[r0:r0+nr, :] (+)= G[nr, nc] * Yg[c0:c0+nc, :]      (A K-major, B MN-major, fp16 x fp16)
            if (dX_acc) {
                CUtensorMap ta, tb;
                if ((rc = tmap_kmajor(&ta, G, nr, gplanes == 2 ? ldg : nc, ldg, BM))) return rc;
                if ((rc = tmap_mnmajor(&tb, Ygb + c0 * ldyg * esz, gext, nc, ldyg))) return rc;
                KArgs a{};
                a.M = nr; a.N = d; a.n_tiles = cdiv(d, BN); a.tiles_per_unit = 1;
                set_segments(a, gplanes, nc, ncp, dpad);
                a.out = dX_acc + r0 * d; a.ldo = d; a.accumulate = (c0 > 0);
                a.oscale0 = logit_scale; a.oscale1 = gscale; a.oscale2 = yg_inv_scale; a.oconst = 1.f / 16384.f;
                if ((rc = launch_gemm<MODE_OUT, 0, 1, 1>(ta, tb, a, dim3(a.n_tiles, cdiv(nr, BM)), st))) return rc;
            }
            // ---- C: dY[c0:c0+nc, :] (+)= G^T[nc, nr] * Xg[r0:r0+nr, :]    (A MN-major, B MN-major, fp16 x fp16)
            if (dY_acc) {
                CUtensorMap ta, tb;
                if ((rc = tmap_mnmajor(&ta, G, gplanes == 2 ? ldg : nc, nr, ldg))) return rc;
                if ((rc = tmap_mnmajor(&tb, Xgb + r0 * ldxg * esz, gext, nr, ldxg))) return rc;
                KArgs a{};
                a.M = nc; a.N = d; a.n_tiles = cdiv(d, BN); a.tiles_per_unit = 1;
                set_segments(a, gplanes, nr, ncp, dpad);
                a.out = dY_acc + c0 * d; a.ldo = d; a.accumulate = (r0 > 0);
                a.oscale0 = logit_scale; a.oscale1 = gscale; a.oscale2 = xg_inv_scale; a.oconst = 1.f / 16384.f;
                if ((rc = launch_gemm<MODE_OUT, 1, 1, 1>(ta, tb, a, dim3(a.n_tiles, cdiv(nc, BM)), st))) return rc;
            }
        }
    }
    return CLIPK_OK;
}

int clipk_cast(const float* src, void* dst, long long n, int dtype, void* stream) {
    if (!src || !dst || n < 0) return fail(CLIPK_EINVAL, "bad argument");
    if (dtype != CLIPK_BF16 && dtype != CLIPK_F32) return fail(CLIPK_EUNSUPPORTED, "dtype %d", dtype);
    if (n == 0) return CLIPK_OK;
    cast_kernel<<<cdiv(cdiv(n, 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, dtype);
    CK_CUDA(cudaGetLastError());
    return CLIPK_OK;
}

int clipk_gemm16(const void* A, const void* B, float* D, int M, int N, int K, long long lda, long long ldb,
                 long long ldd, int a_mn, int b_mn, int f16, int accumulate, void* stream) {
    if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) return fail(CLIPK_EINVAL, "bad argument");
    if (N % 4 != 0 || ldd % 4 != 0) return fail(CLIPK_EUNSUPPORTED, "N and ldd must be multiples of 4");
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb;
    if (a_mn) rc = tmap_mnmajor(&ta, A, M, K, lda); else rc = tmap_kmajor(&ta, A, M, K, lda, BM);
    if (rc) return rc;
    if (b_mn) rc = tmap_mnmajor(&tb, B, N, K, ldb); else rc = tmap_kmajor(&tb, B, N, K, ldb, BN);
    if (rc) return rc;
    KArgs a{};
    a.M = M; a.N = N; a.n_tiles = cdiv(N, BN); a.tiles_per_unit = 1;
    set_segments(a, 1, K, 0, 0);
    a.out = D; a.ldo = int(ldd); a.accumulate = accumulate; a.oconst = 1.f;
    dim3 grid(a.n_tiles, cdiv(M, BM));
    if (f16) {
        if (!a_mn && !b_mn) return launch_gemm<MODE_OUT, 0, 0, 1>(ta, tb, a, grid, st);
        if (!a_mn && b_mn) return launch_gemm<MODE_OUT, 0, 1, 1>(ta, tb, a, grid, st);
        if (a_mn && b_mn) return launch_gemm<MODE_OUT, 1, 1, 1>(ta, tb, a, grid, st);
        return fail(CLIPK_EUNSUPPORTED, "fp16 operands: (a_mn=1, b_mn=0) is not built");
    }
    if (!a_mn && !b_mn) return launch_gemm<MODE_OUT, 0, 0, 0>(ta, tb, a, grid, st);
    if (!a_mn && b_mn) return launch_gemm<MODE_OUT, 0, 1, 0>(ta, tb, a, grid, st);
    if (a_mn && b_mn) return launch_gemm<MODE_OUT, 1, 1, 0>(ta, tb, a, grid, st);
    return launch_gemm<MODE_OUT, 1, 0, 0>(ta, tb, a, grid, st);
}

}  // extern "C"
