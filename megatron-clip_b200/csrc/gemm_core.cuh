// Warp-specialised tcgen05 GEMM core for the ClipLoss path (sm_100a only).
//
//   D[M, N] = A[M, K] * B[N, K]^T      bf16 operands, fp32 accumulation in TMEM
//
// One CTA = 6 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  Three pipelines: a 4-stage shared-memory ring
// (TMA -> MMA, mbarrier full/empty), a 2-stage TMEM accumulator ring (MMA -> epilogue) and the unit's tile loop.
// A CTA tile is 128 x 256 (UMMA M=128, N=256, K=16), K is streamed in 64-element (128-byte, SWIZZLE_128B) blocks.
//
// Epilogue modes (all read the accumulator with tcgen05.ld, thread == one row of the tile):
//   MODE_STATS : online row max / sum-of-exp of s*A*B^T over a run of column tiles + the positive (diagonal) logit.
//                This is the logits + cross-entropy forward of loss.py:112-119,135-138 without storing the logits.
//   MODE_GRAD  : recompute the tile, form G = s*g*(alpha*(P_row - Id) + beta*(P_col - Id)) in registers, write it as
//                bf16 into an L2-resident panel, and accumulate the dlogit_scale sums.
//   MODE_OUT   : plain fp32 output (optionally accumulating) - the two gradient GEMMs dX = G*Y and dY = G^T*X.
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the latter lets the gradient
// GEMMs read Y, X and the G panel in place, without transposes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ptx.cuh"

namespace clipk {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int NUM_EPI_WARPS = 4;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;               // 2 accumulator stages x 256 fp32 columns
constexpr int MISC_BYTES = 4096;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + MISC_BYTES;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

enum Mode { MODE_STATS = 0, MODE_GRAD = 1, MODE_OUT = 2 };

struct KArgs {
    int M, N;             // output extent (rows of A, rows of B)
    int num_kb;           // K blocks of 64 elements over all segments (= nseg * kb_per_seg)
    // K segments: block kb belongs to segment kb / kb_per_seg; a_off/b_off shift the INNER (contiguous) coordinate of
    // the operand for that segment.  16-bit inputs use one segment; fp32 values are held as two scaled fp16 planes
    // [hi | lo] laid side by side and contracted as three plane pairs (lo.hi, hi.lo, hi.hi - small terms first).
    int nseg, kb_per_seg;
    int a_off[6], b_off[6];
    int n_tiles;          // ceil(N / BN)
    int tiles_per_unit;   // column tiles swept by one CTA (STATS / GRAD); 1 for OUT
    // STATS / GRAD
    const float* scale;        // device scalar logit_scale
    const float* xs;           // device scalars: true value = stored value * (*xs) for A, (*ys) for B; null = 1
    const float* ys;
    long long diag_offset;     // the positive of row i is column diag_offset + i
    float* part_max;           // [gridDim.x][M]  running max of s*log2e*acc
    float* part_sum;           // [gridDim.x][M]
    float* pos;                // [M] natural-log units, written by the CTA whose run contains the diagonal
    // GRAD
    const float* lse_row;      // [M] natural log
    const float* lse_col;      // [N] natural log
    float alpha, beta;
    const float* gscale;       // device scalar: grad_output * c
    __half* G;                 // panel of G * 2^14 in fp16, row-major, rows padded to BM and ldg to BN
    int ldg;
    int g_planes;              // 1, or 2 (G split into hi/lo fp16 planes, g_plane_stride columns apart)
    int g_plane_stride;
    float* ds_acc;             // [2]: sum (P_row-Id).Sraw, sum (P_col-Id).Sraw   (atomicAdd)
    float* ds_col;             // [N] per-column sum (P_col-Id).Sraw, or nullptr
    // OUT
    float* out;
    int ldo;
    int accumulate;
    const float* oscale0;      // out = acc * (*oscale0) * (*oscale1) * (*oscale2) * oconst   (null pointers count as 1)
    const float* oscale1;
    const float* oscale2;
    float oconst;
};

// lane j ends with sum over the warp's lanes of x[j] (31 shuffles: 16 + 8 + 4 + 2 + 1).
__device__ __forceinline__ float warp_transpose_reduce32(float (&x)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? x[i] : x[i + off];
            const float keep = upper ? x[i + off] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return x[0];
}

template <int MODE, int A_MN, int B_MN, int F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw_u32);
    const uint32_t sA = base;
    const uint32_t sB = sA + STAGES * A_STAGE_BYTES;
    const uint32_t sBar = sB + STAGES * B_STAGE_BYTES;
    const uint32_t bar_full = sBar;
    const uint32_t bar_empty = sBar + STAGES * 8;
    const uint32_t bar_tfull = sBar + 2 * STAGES * 8;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t sTmemPtr = bar_tempty + 16;
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + (2 * STAGES + 4) * 8);
    float* misc = reinterpret_cast<float*>(base_ptr + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_blk = blockIdx.y;
    const int t0 = blockIdx.x * args.tiles_per_unit;
    const int t1 = min(args.n_tiles, t0 + args.tiles_per_unit);
    const int num_kb = args.num_kb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(bar_tfull + 8 * a, 1);
            ptx::mbar_init(bar_tempty + 8 * a, NUM_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
    }
    if (warp == 1) {
        ptx::tmem_alloc(sTmemPtr, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = t0; t < t1; ++t) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int seg = kb / args.kb_per_seg;
                    const int kw = (kb - seg * args.kb_per_seg) * BK;
                    const int ao = args.a_off[seg], bo = args.b_off[seg];
                    ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t full = bar_full + 8 * s;
                    ptx::mbar_arrive_expect_tx(full, A_STAGE_BYTES + B_STAGE_BYTES);
                    const uint32_t a_dst = sA + s * A_STAGE_BYTES;
                    const uint32_t b_dst = sB + s * B_STAGE_BYTES;
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            ptx::tma_load_2d(a_dst + i * 8192, &tmA, ao + m_blk * BM + i * 64, kw, full);
                    } else {
                        ptx::tma_load_2d(a_dst, &tmA, ao + kw, m_blk * BM, full);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            ptx::tma_load_2d(b_dst + i * 8192, &tmB, bo + t * BN + i * 64, kw, full);
                    } else {
                        ptx::tma_load_2d(b_dst, &tmB, bo + kw, t * BN, full);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_16bit(BM, BN, A_MN, B_MN, /*a_is_bf16=*/!F16, /*b_is_bf16=*/!F16);
            // K-major SW128: 8-row groups 1024 B apart (SBO); MN-major SW128: 64-wide MN blocks 8192 B apart (LBO),
            // 8-row K groups 1024 B apart (SBO).
            constexpr uint64_t adesc_hi = A_MN ? ptx::make_smem_desc_sw128(8192, 1024) : ptx::make_smem_desc_sw128(16, 1024);
            constexpr uint64_t bdesc_hi = B_MN ? ptx::make_smem_desc_sw128(8192, 1024) : ptx::make_smem_desc_sw128(16, 1024);
            constexpr uint32_t a_kstep = A_MN ? 2048 : 32;   // bytes per UMMA_K = 16 elements
            constexpr uint32_t b_kstep = B_MN ? 2048 : 32;
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = t0; t < t1; ++t, ++it) {
                const int a = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * a, aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(bar_full + 8 * s, ph);
                    ptx::tc_fence_after();
                    const uint32_t a_src = sA + s * A_STAGE_BYTES;
                    const uint32_t b_src = sB + s * B_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        ptx::mma_f16_ss(d_tmem, ptx::desc_with_addr(adesc_hi, a_src + k * a_kstep),
                                        ptx::desc_with_addr(bdesc_hi, b_src + k * b_kstep), idesc, (kb | k) != 0);
                    }
                    ptx::mma_commit(bar_empty + 8 * s);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::mma_commit(bar_tfull + 8 * a);
            }
        }
    } else {
        // ================================ epilogue warps ================================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row_in_tile = q * 32 + lane;
        const int row = m_blk * BM + row_in_tile;
        const int epi_tid = (warp - 2) * 32 + lane;   // 0..127
        const bool row_ok = row < args.M;
        const long long dcol = args.diag_offset + row;

        // acc holds the dot product of the STORED operands; dequant = xs*ys turns it into the true x.y
        float sc = 0.f, s_nat = 0.f, oscale = 1.f, dequant = 1.f;
        if (MODE != MODE_OUT) {
            if (args.xs) dequant *= __ldg(args.xs);
            if (args.ys) dequant *= __ldg(args.ys);
            s_nat = __ldg(args.scale) * dequant;      // S = s_nat * acc
            sc = s_nat * LOG2E;
        } else {
            oscale = args.oconst;
            if (args.oscale0) oscale *= __ldg(args.oscale0);
            if (args.oscale1) oscale *= __ldg(args.oscale1);
            if (args.oscale2) oscale *= __ldg(args.oscale2);
        }
        // STATS state
        float run_m = -CUDART_INF_F, run_l = 0.f, pos_raw = 0.f;
        bool have_pos = false;
        // GRAD state
        float Lr = 0.f, ga = 0.f, gb = 0.f, u_acc = 0.f, w_acc = 0.f;
        if (MODE == MODE_GRAD) {
            if (row_ok) Lr = __ldg(args.lse_row + row) * LOG2E;
            // G is stored as 2^14 * (alpha*(P_row-Id) + beta*(P_col-Id)) in fp16: |.| <= 2^15 and entries down to
            // ~4e-9 keep the full 11-bit mantissa; logit_scale * gscale * 2^-14 is applied by the gradient GEMMs.
            ga = 16384.f * args.alpha;
            gb = 16384.f * args.beta;
        }

        int it = 0;
        for (int t = t0; t < t1; ++t, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int n0 = t * BN;
            float* lc_s = misc + a * BN;   // staged column LSE (log2 units) for this tile
            if (MODE == MODE_GRAD) {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j) {
                    const int c = n0 + epi_tid + j * 128;
                    lc_s[epi_tid + j * 128] = (c < args.N) ? __ldg(args.lse_col + c) * LOG2E : CUDART_INF_F;
                }
                ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
            }
            ptx::mbar_wait(bar_tfull + 8 * a, aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + a * BN + (uint32_t(q * 32) << 16);
            // does this tile contain any positive (diagonal) entry of this CTA's rows?
            const long long d_lo = args.diag_offset + (long long)m_blk * BM;
            const bool diag_tile = (MODE != MODE_OUT) && (d_lo + BM - 1 >= n0) && (d_lo < (long long)n0 + BN);

#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr + c * 32, r);
                ptx::tmem_ld_wait();
                const int col0 = n0 + c * 32;

                if (MODE == MODE_STATS) {
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) * sc;
                    if (col0 + 32 > args.N) {
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (col0 + k >= args.N) v[k] = -CUDART_INF_F;
                    }
                    float cmax = v[0];
#pragma unroll
                    for (int k = 1; k < 32; ++k) cmax = fmaxf(cmax, v[k]);
                    const float m_new = fmaxf(run_m, cmax);
                    if (m_new > -CUDART_INF_F) {
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < 32; ++k) acc += ptx::ex2(v[k] - m_new);
                        run_l = run_l * ptx::ex2(run_m - m_new) + acc;
                        run_m = m_new;
                    }
                    if (diag_tile) {
                        const bool hit = (dcol >= col0) && (dcol < (long long)col0 + 32) && (dcol < args.N);
                        if (__any_sync(0xffffffffu, hit)) {
                            const int idx = hit ? int(dcol - col0) : -1;
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (k == idx) pos_raw = __uint_as_float(r[k]);
                            have_pos = have_pos || hit;
                        }
                    }
                } else if (MODE == MODE_GRAD) {
                    float qv[32];
                    float gv[32];
                    uint32_t packed[16];
                    const bool tail = (col0 + 32 > args.N);
                    const int didx = (diag_tile && dcol >= col0 && dcol < (long long)col0 + 32) ? int(dcol - col0) : -1;
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        float g2[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int kk = k + h;
                            const float sraw = __uint_as_float(r[kk]);
                            float vv = sraw * sc;
                            if (tail && col0 + kk >= args.N) vv = -CUDART_INF_F;
                            float pr = ptx::ex2(vv - Lr);
                            float pc = ptx::ex2(vv - lc_s[c * 32 + kk]);
                            if (diag_tile && kk == didx) { pr -= 1.f; pc -= 1.f; }
                            g2[h] = ga * pr + gb * pc;
                            if (tail && col0 + kk >= args.N) g2[h] = 0.f;
                            gv[kk] = g2[h];
                            u_acc = fmaf(pr, sraw, u_acc);
                            const float qq = pc * sraw;
                            w_acc += qq;
                            qv[kk] = qq;
                        }
                        packed[k >> 1] = ptx::pack_f16x2(g2[0], g2[1]);
                    }
                    // G panel rows are padded to BM and columns to BN, so no bounds checks on the store.
                    uint4* gp = reinterpret_cast<uint4*>(args.G + (size_t)row * args.ldg + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        gp[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    if (args.g_planes == 2) {
                        // residual plane: g = hi + lo with each piece exactly representable in fp16 (22 bits)
#pragma unroll 1
                        for (int pl = 1; pl < 2; ++pl) {
#pragma unroll
                            for (int k = 0; k < 32; k += 2) {
                                const uint32_t prev = packed[k >> 1];
                                gv[k] -= __half2float(__ushort_as_half((unsigned short)(prev & 0xffffu)));
                                gv[k + 1] -= __half2float(__ushort_as_half((unsigned short)(prev >> 16)));
                                packed[k >> 1] = ptx::pack_f16x2(gv[k], gv[k + 1]);
                            }
                            uint4* gq = gp + (size_t)pl * args.g_plane_stride / 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                gq[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                        }
                    }
                    if (args.ds_col != nullptr) {
                        if (!row_ok) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) qv[k] = 0.f;
                        }
                        const float colsum = warp_transpose_reduce32(qv, lane) * dequant;
                        if (col0 + lane < args.N) atomicAdd(args.ds_col + col0 + lane, colsum);
                    }
                } else {  // MODE_OUT
                    if (row_ok) {
                        float* op = args.out + (size_t)row * args.ldo + col0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (col0 + 4 * j < args.N) {
                                float4 o = make_float4(__uint_as_float(r[4 * j]) * oscale, __uint_as_float(r[4 * j + 1]) * oscale,
                                                       __uint_as_float(r[4 * j + 2]) * oscale, __uint_as_float(r[4 * j + 3]) * oscale);
                                if (args.accumulate) {
                                    const float4 old = *reinterpret_cast<const float4*>(op + 4 * j);
                                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                }
                                *reinterpret_cast<float4*>(op + 4 * j) = o;
                            }
                        }
                    }
                }
            }
            // accumulator stage drained: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * a);
        }

        if (MODE == MODE_STATS) {
            if (row_ok) {
                args.part_max[(size_t)blockIdx.x * args.M + row] = run_m;
                args.part_sum[(size_t)blockIdx.x * args.M + row] = run_l;
                if (have_pos) args.pos[row] = pos_raw * s_nat;
            }
        }
        if (MODE == MODE_GRAD) {
            if (!row_ok) { u_acc = 0.f; w_acc = 0.f; }
            u_acc *= dequant;
            w_acc *= dequant;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                u_acc += __shfl_xor_sync(0xffffffffu, u_acc, off);
                w_acc += __shfl_xor_sync(0xffffffffu, w_acc, off);
            }
            if (lane == 0 && t1 > t0) {
                atomicAdd(args.ds_acc + 0, u_acc);
                atomicAdd(args.ds_acc + 1, w_acc);
            }
        }
    }

    // ================================ teardown ================================
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace clipk
