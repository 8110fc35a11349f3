// FP16-split helpers shared by the round-2 tensor-core kernels (solve_tc.cu, solve_dopri5.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "umma.cuh"

namespace tcde {
namespace tc {

using namespace umma;

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {          // (lo, hi) -> f16x2, round to nearest even
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void unpack_h2(uint32_t h, float& lo, float& hi) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h));
}
__device__ __forceinline__ int exponent_of(float m) { return (int)((__float_as_uint(m) >> 23) & 0xFFu); }   // biased
__device__ __forceinline__ float pow2_biased(int e) { return __uint_as_float((uint32_t)e << 23); }          // 2^(e-127)

// byte offset of (row, 16-byte chunk) in a K-major no-swizzle tile of 32-byte rows: 8-row groups 256 B apart (SBO),
// the two K chunks 128 B apart (LBO), rows 16 B apart
__device__ __forceinline__ uint32_t aug_off(int row, int chunk) { return (uint32_t)((row >> 3) * 256 + chunk * 128 + (row & 7) * 16); }


// ---- TMA (cp.async.bulk.tensor, SASS UTMALDG): the spline rows of one interval for the 128 paths of a tile ------------
// The control is a 2-D tensor [paths][n_rows * row_floats] (row stride = one path's whole coefficient block); a box of
// {row_floats, 128 paths} at coordinates (interval * row_floats, first path) is the 128 rows x 128 bytes (cubic, 8
// channels: a | b | 2c | 3d) that one Runge-Kutta stage of a tile reads.  One thread issues the copy; the bytes land in
// shared memory with the 128-byte swizzle (16-byte chunk c of row r at position c ^ (r & 7)): every thread then reads
// its own row with conflict-free 128-bit loads.  Rows beyond the batch are zero-filled by the TMA unit.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global tensor store (bulk group completion): box at (c0, c1, c2) of a 3-D tensor
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// host: the tensor map of a control tensor (float32).  cuTensorMapEncodeTiled is taken from the driver through the runtime
// (no link-time dependency on libcuda).  Returns cudaSuccess or an error.
inline int make_rows_tensor_map(CUtensorMap* map, const float* control, int64_t n_paths, int64_t n_rows, int row_floats, int box_rows = 128) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)(n_rows * row_floats), (cuuint64_t)n_paths};
    const cuuint64_t strides[1] = {(cuuint64_t)(n_rows * row_floats) * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)row_floats, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUtensorMapSwizzle swz = (row_floats * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(control), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// host: the tensor map of a stage trajectory [n_stages][n_paths][32] (float32), box = one tile of 128 paths of one stage,
// 128-byte swizzle in shared memory (rows beyond n_paths are clipped by the TMA unit).
inline int make_stage_tensor_map(CUtensorMap* map, float* stages, int64_t n_paths, int64_t n_stage_total, int box_rows = 128) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t dims[3] = {32u, (cuuint64_t)n_paths, (cuuint64_t)n_stage_total};
    const cuuint64_t strides[2] = {128u, (cuuint64_t)n_paths * 128u};
    const cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, stages, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// ---- building blocks of a [128 paths x 32 hidden] . [32 x 256] tile product with the 2xFP16 operand split ------------
constexpr int kHid = 32, kCh = 8, kCols = kHid * kCh, kRows = 128;

// Operand B = the weights [256][32] scaled by one power of two and split into hi | lo halves: row n = 128 bytes,
// swizzle 128B; bias block: K-major no-swizzle [256 rows][16 halves] with (bias_hi, bias_lo, 0...).  Called by every
// thread of the CTA (contains __syncthreads).  Returns the scales through the references.
__device__ __forceinline__ void prepare_b_fp16(unsigned char* b_tile, unsigned char* b_aug, const float* weight, const float* bias,
                                               float* red, int tid, int n_threads, float& w_scale, float& inv_w_scale, float& beta) {
    float wmax = 0.f, bmax = 0.f;
    for (int e = tid; e < kCols * kHid; e += n_threads) wmax = fmaxf(wmax, fabsf(weight[e]));
    for (int e = tid; e < kCols; e += n_threads) bmax = fmaxf(bmax, fabsf(bias[e]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = wmax; red[16 + (tid >> 5)] = bmax; }
    __syncthreads();
    wmax = 0.f; bmax = 0.f;
    for (int w = 0; w < n_threads / 32; ++w) { wmax = fmaxf(wmax, red[w]); bmax = fmaxf(bmax, red[16 + w]); }
    const int eb = min(max(exponent_of(bmax), 40), 215);
    const int ew = (wmax > 0.f) ? min(max(exponent_of(wmax), 40), 215) : (bmax > 0.f ? eb : 127);
    w_scale = pow2_biased(127 + 13 - (ew - 127));                                // max |W| * w_scale in [2^13, 2^14)
    inv_w_scale = pow2_biased(127 - 13 + (ew - 127));
    beta = (bmax > 0.f) ? pow2_biased(min(max(127 + eb - ew, 2), 250)) : 0.f;    // 2^(ex(bmax) - ex(wmax))
    __half* bt = reinterpret_cast<__half*>(b_tile);
    for (int e = tid; e < kCols * kHid; e += n_threads) {
        const int n = e >> 5, k = e & 31;
        const float w = weight[e] * w_scale;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        bt[n * 64 + ((((k >> 3)) ^ (n & 7)) << 3) + (k & 7)] = hi;
        bt[n * 64 + ((((k >> 3) + 4) ^ (n & 7)) << 3) + (k & 7)] = lo;
    }
    const float bias_scale = (beta > 0.f) ? w_scale / beta : 0.f;                // bias * w_scale / beta in [2^13, 2^14)
    for (int row = tid; row < kCols; row += n_threads) {
        const float bv = bias[row] * bias_scale;
        const __half hi = __float2half_rn(bv);
        const __half lo = __float2half_rn(bv - __half2float(hi));
        *reinterpret_cast<uint4*>(b_aug + aug_off(row, 0)) =
            make_uint4((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(b_aug + aug_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// Row r of operand A: z (32 floats) times the path's own power of two, split into hi | lo halves (one 128-byte swizzled
// row) plus the row's entry of the bias block.  Returns 1 / (row scale * weight scale), to be folded into dX/dt.
__device__ __forceinline__ float split_store_fp16(const float* z, unsigned char* a_tile, unsigned char* a_aug, int r, float beta,
                                                  float inv_w_scale) {
    float mx[8];                          // a tree, not a 32-long dependent chain
#pragma unroll
    for (int k = 0; k < 8; ++k) mx[k] = fmaxf(fmaxf(fabsf(z[k]), fabsf(z[k + 8])), fmaxf(fabsf(z[k + 16]), fabsf(z[k + 24])));
    const float m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])), beta));
    const int e = min(max(exponent_of(m), 30), 224);
    const float s = pow2_biased(127 + 13 - (e - 127));
    const f2 s2 = pk(s, s);
    uint32_t hi_h[16], lo_h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f2 sc = mul2(pk(z[2 * j], z[2 * j + 1]), s2);
        float s0, s1, h0, h1, l0, l1;
        upk(sc, s0, s1);
        hi_h[j] = pack_h2(s0, s1);
        unpack_h2(hi_h[j], h0, h1);
        upk(sub2(sc, pk(h0, h1)), l0, l1);
        lo_h[j] = pack_h2(l0, l1);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4*>(a_tile + r * 128 + ((c ^ (r & 7)) << 4)) =
            make_uint4(hi_h[4 * c], hi_h[4 * c + 1], hi_h[4 * c + 2], hi_h[4 * c + 3]);
        *reinterpret_cast<uint4*>(a_tile + r * 128 + (((c + 4) ^ (r & 7)) << 4)) =
            make_uint4(lo_h[4 * c], lo_h[4 * c + 1], lo_h[4 * c + 2], lo_h[4 * c + 3]);
    }
    const float sb = s * beta;            // <= 2^13 by construction; exact; flushes to 0 far below the row maximum
    *reinterpret_cast<uint32_t*>(a_aug + aug_off(r, 0)) = pack_h2(sb, sb);
    return pow2_biased(127 - 13 + (e - 127)) * inv_w_scale;
}

// The seven MMAs of one tile-stage (issued by one thread): z_lo.W_hi + z_hi.W_lo + z_hi.W_hi + row_scale * bias
__device__ __forceinline__ void issue_fp16(uint32_t tmem_d, const unsigned char* a_tile, const unsigned char* a_aug,
                                           const unsigned char* b_tile, const unsigned char* b_aug, uint64_t* d_ready) {
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(kCols >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);   // F16 x F16 -> F32
    const uint64_t da = make_desc(a_tile), db = make_desc(b_tile);
    const uint64_t da_aug = (uint64_t)((smem_u32(a_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
    const uint64_t db_aug = (uint64_t)((smem_u32(b_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
    mma_f16(tmem_d, da + 4, db + 0, idesc, 0);
    mma_f16(tmem_d, da + 6, db + 2, idesc, 1);
    mma_f16(tmem_d, da + 0, db + 4, idesc, 1);
    mma_f16(tmem_d, da + 2, db + 6, idesc, 1);
    mma_f16(tmem_d, da + 0, db + 0, idesc, 1);
    mma_f16(tmem_d, da + 2, db + 2, idesc, 1);
    mma_f16(tmem_d, da_aug, db_aug, idesc, 1);
    mma_commit(d_ready);
}

// kv[h] = sum_c D[h * 8 + c] * dx[c] for the thread's own accumulator row (256 TMEM columns from taddr); W = columns per
// TMEM load (16 or 32), the next load in flight while the current one is consumed (see umma.cuh)
template <int W>
__device__ __forceinline__ void contract_row(uint32_t taddr, const f2* dx2, float* kv) {
    uint32_t va[W], vb[W];
    if (W == 32) tmem_ld32_issue(taddr, va); else tmem_ld16_issue(taddr, va);
#pragma unroll
    for (int j = 0; j < kCols / W; ++j) {
        uint32_t* cur = (j & 1) ? vb : va;
        if (W == 32) tmem_ld32_wait(cur); else tmem_ld16_wait(cur);
        if (j + 1 < kCols / W) {
            if (W == 32) tmem_ld32_issue(taddr + (uint32_t)(W * (j + 1)), (j & 1) ? va : vb);
            else tmem_ld16_issue(taddr + (uint32_t)(W * (j + 1)), (j & 1) ? va : vb);
        }
#pragma unroll
        for (int hh = 0; hh < W / 8; ++hh) {
            f2 acc = mul2(pk(__uint_as_float(cur[8 * hh + 0]), __uint_as_float(cur[8 * hh + 1])), dx2[0]);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 2]), __uint_as_float(cur[8 * hh + 3])), dx2[1], acc);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 4]), __uint_as_float(cur[8 * hh + 5])), dx2[2], acc);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 6]), __uint_as_float(cur[8 * hh + 7])), dx2[3], acc);
            float lo, hi;
            upk(acc, lo, hi);
            kv[(W / 8) * j + hh] = lo + hi;
        }
    }
}

// seg_first_out[s] = number of requested outputs that fall before segment s's first step (host schedule: out_step[j] is the
// step in which output j is produced, -1 for outputs at the initial time); one tiny launch per solve that is cut in segments
static __global__ void segment_cursor_kernel(const int32_t* out_step, int n_out, int n_seg, int steps_per_seg, int32_t* cursor) {
    const int s = threadIdx.x;
    if (s >= n_seg) return;
    int j = 0;
    while (j < n_out && out_step[j] < s * steps_per_seg) ++j;
    cursor[s] = j;
}

}  // namespace tc
}  // namespace tcde
