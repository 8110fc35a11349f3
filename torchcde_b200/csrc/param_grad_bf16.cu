// The parameter gradients of a whole fixed-step backward solve on the tensor cores, round 2 (sm_100a).
//
//   dL/dW[hC+c][k] = sum_e sum_p  w_e a_e[p][h] dX_e[p][c] * z_e[p][k]          (e: stages, p: paths)
//   dL/db[hC+c]    = sum_e sum_p  w_e a_e[p][h] dX_e[p][c]
//
// is ONE matrix product over the (stage, path) pairs:  D[m][n] = sum_pairs Zx[pair][m] * U[pair][n]  with  Zx = (z | 1)
// (33 of 128 rows used) and U = w a (x) dX formed on the fly (256 columns).  Round 1 (param_grad_umma.cu, kept as variant 1)
// put U^T on the M side: 24 MMAs of N = 48 per 32 pairs -- at ~95 cycles of issue per tcgen05.mma whatever its size that is
// 2,200 cycles per 32 pairs, 24 ms for the BASELINE backward; and every thread scattered 144 scalar shared-memory stores per
// item (a transpose through shared memory).  Here:
//   * U is the N side: ONE MMA covers all 256 columns -> 6 MMAs (M 128, N 256, K 16) per 32 pairs = 768 cycles of tensor pipe;
//   * both operands are MN-major (the pair index is K): a thread's 8 products a[h] * dX[0..7] ARE one 16-byte chunk of its
//     pair's row -- 128-bit stores straight into the canonical layout (8 k-rows x 128 B atoms, 128B swizzle), no transpose;
//   * BF16 operands with a two-way split (hi + lo: z_lo.U_hi + z_hi.U_lo + z_hi.U_hi, error ~2^-16 per product -- a gradient
//     needs no more, and unlike FP16 no scaling is needed: a per-pair scale cannot be factored out of a sum over pairs);
//   * a, z and the spline rows of an item arrive by TMA (three tensor boxes of 32 paths, 128-byte swizzle: the producers'
//     128-bit reads of their own row are bank-conflict free -- linear rows cost 8 wavefronts per load, ncu: 58 % LSU) into
//     a 3-deep staging ring; the sixteen producer warps share an item (warp = MN block g of U = hidden units 8g..8g+7,
//     a quarter of it: two 16-byte chunks).  An item is a serial chain per warp (barrier wait, loads, products, stores,
//     proxy fence, arrive) -- four producer warps needed ~1,900 cycles per item, eight 1,390, against 768 of tensor pipe:
//     more warps per scheduler hide that chain.
// The tensor core adds into its fp32 accumulator with truncation (measured in round 1: 2.6e-3 relative drift over ~10^5
// accumulations), so accumulation runs in chunks of kChunk items into two alternating TMEM sets; two fold warps add each
// finished chunk into fp32 sums in shared memory with round-to-nearest adds, one chunk behind the tensor pipe.
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace tcde {
namespace pg2 {

using namespace umma;

constexpr int H = 32, C = 8;
constexpr int kPairs = 32;                 // (stage, path) pairs per item == K of one operand buffer
constexpr int kN = 256, kM = 128;
// template parameter Q: parts an MN block of U is cut in = producer warps / 4 (2: warps 0-7, four hidden units per thread;
// 4: warps 0-15, two per thread); then two fold warps, the MMA issuer warp, the TMA warp
constexpr int kChunk = 16;
constexpr int kParams = H * C * H + H * C;

// shared memory map (bytes)
constexpr int oA = 16384;                                // per buffer: A_hi (8 KB) | A_lo (8 KB) | B_hi (16 KB) | B_lo (16 KB)
constexpr int kBufBytes = 49152;
constexpr int kStgBytes = 12288;                         // per staging slot: a (4 KB) | z (4 KB) | rows (4 KB)
// kBuf operand buffers, then kStg staging slots, then the fp32 sums [256 n][33 m], then the mbarriers
constexpr int smem_bytes(int kBuf, int kStg) { return kBuf * kBufBytes + kStg * kStgBytes + kN * 33 * 4 + 256; }

// byte offset of (pair k, 16-byte chunk `chunk` of MN block `blk`) in an MN-major 128B-swizzled tile with 4 k-groups
__device__ __forceinline__ uint32_t mn_off(int blk, int k, int chunk) {
    return (uint32_t)(blk * 4096 + (k >> 3) * 1024 + (k & 7) * 128 + ((chunk ^ (k & 7)) << 4));
}
__device__ __forceinline__ uint64_t desc_mn(const void* tile) {     // LBO = 4096 B between MN blocks, SBO = 1024 B between k-groups
    return (uint64_t)((smem_u32(tile) & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// (x0, x1) -> packed bf16 hi parts and packed bf16 lo parts (x - hi, exact in fp32, then rounded)
__device__ __forceinline__ void split_bf2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf2(x0 - h0, x1 - h1);
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int Q, int kBuf, int kStg>
__global__ void __launch_bounds__(128 * Q + 128, 1)
param_grad_bf16_kernel(const float* __restrict__ control, int control_kind, int64_t n_rows, const float* __restrict__ z_stages,
                       const float* __restrict__ a_stages, const int32_t* __restrict__ stage_index, const float* __restrict__ stage_frac,
                       const float* __restrict__ stage_weight, int n_stage_total, float* __restrict__ scratch, int64_t n_paths,
                       const __grid_constant__ CUtensorMap rows_map, const __grid_constant__ CUtensorMap a_map,
                       const __grid_constant__ CUtensorMap z_map, const int issuer_fence) {
    constexpr int kProducers = 128 * Q, kThreads = kProducers + 128;
    constexpr int kFoldWarp = 4 * Q, kIssuerWarp = 4 * Q + 2, kTmaWarp = 4 * Q + 3;
    constexpr int HU = 8 / Q;                                             // hidden units (16-byte chunks of U) per producer thread
    constexpr int oStg = kBuf * kBufBytes, oAcc = oStg + kStg * kStgBytes, oBars = oAcc + kN * 33 * 4;
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    uint64_t* stg_full = reinterpret_cast<uint64_t*>(smem + oBars);      // [kStg]
    uint64_t* stg_free = stg_full + kStg;                                 // [kStg]
    uint64_t* full = stg_free + kStg;                                     // [kBuf]
    uint64_t* empty = full + kBuf;                                        // [kBuf]
    uint64_t* chunk_done = empty + kBuf;                                  // [2]
    uint64_t* set_free = chunk_done + 2;                                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(set_free + 2);
    float* acc = reinterpret_cast<float*>(smem + oAcc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool cubic = (control_kind == TCDE_CONTROL_CUBIC);
    const int row_floats = cubic ? 4 * C : C;

    // ---- one-time setup: zero everything, then the constant parts of the A tiles (the ones row m = 32) -----------------
    for (int e = tid; e < (oAcc + kN * 33 * 4) / 16; e += kThreads) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int e = tid; e < kBuf * kPairs; e += kThreads) {                 // A_hi[pair][m = 32] = 1.0 (bf16 0x3F80): chunk 4, element 0
        const int b = e / kPairs, k = e % kPairs;
        *reinterpret_cast<uint32_t*>(smem + b * kBufBytes + mn_off(0, k, 4)) = 0x3F80u;
    }
    if (tid == 0) {
        for (int i = 0; i < kStg; ++i) { mbar_init(&stg_full[i], 1); mbar_init(&stg_free[i], kProducers); }
        for (int i = 0; i < kBuf; ++i) { mbar_init(&full[i], kProducers); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&chunk_done[i], 1); mbar_init(&set_free[i], 33); }
        fence_barrier_init();
        tc::tma_prefetch_desc(&rows_map);
        tc::tma_prefetch_desc(&a_map);
        tc::tma_prefetch_desc(&z_map);
    }
    if (warp == kIssuerWarp) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_blocks = (n_paths + kPairs - 1) / kPairs;
    const int64_t n_items = n_blocks * n_stage_total;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t n_mine = first < n_items ? (n_items - first + stride - 1) / stride : 0;
    const int64_t n_chunks = (n_mine + kChunk - 1) / kChunk;
    // item -> (stage e, block of 32 paths) without a 64-bit division per item (ncu: ~45 of the ~330 instructions a producer warp
    // spent per item): the walker advances by `stride` blocks and carries into the stage index
    const int n_blocks_i = (int)n_blocks, stride_i = (int)stride;
    int walk_e = (int)(first / n_blocks), walk_blk = (int)(first - (int64_t)walk_e * n_blocks);
    auto advance = [&]() {
        walk_blk += stride_i;
        while (walk_blk >= n_blocks_i) { walk_blk -= n_blocks_i; ++walk_e; }
    };

    if (warp == kTmaWarp) {
        // ================================ TMA: a, z, spline rows of item j -> staging ================================
        if (lane == 0) {
            int s = 0;
            uint32_t par = 1;                                             // parity of "slot s is free" (free at the start)
            for (int64_t j = 0; j < n_mine; ++j) {
                mbar_wait_relaxed(&stg_free[s], par, 100);
                const int e = walk_e;
                const int64_t path0 = (int64_t)walk_blk * kPairs;
                advance();
                unsigned char* dst = smem + oStg + s * kStgBytes;
                // boxes are always delivered whole (rows beyond n_paths arrive as zeros)
                tc::mbar_expect_tx(&stg_full[s], (uint32_t)(2 * kPairs * H * 4 + kPairs * row_floats * 4));
                tc::tma_load_3d(dst, &a_map, 0, (int)path0, e, &stg_full[s]);
                tc::tma_load_3d(dst + 4096, &z_map, 0, (int)path0, e, &stg_full[s]);
                tc::tma_load_2d(dst + 8192, &rows_map, stage_index[e] * row_floats, (int)path0, &stg_full[s]);
                if (++s == kStg) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == kIssuerWarp) {
        // ================================ MMA issuer ===================================================================
        if (lane == 0) {
            // D = F32, A = B = BF16, both MN-major, N = 256, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
            int b = 0;
            uint32_t par = 0;                                             // parity of "buffer b is full"
            for (int64_t j = 0; j < n_mine; ++j) {
                const int64_t chunk = j / kChunk;                         // kChunk is a power of two: shifts
                const int set = (int)(chunk & 1);
                const bool opens = (j % kChunk) == 0, closes = (j % kChunk) == kChunk - 1 || j == n_mine - 1;
                if (opens && chunk >= 2) mbar_wait(&set_free[set], (uint32_t)(((chunk >> 1) & 1) ^ 1));     // chunk - 2 has been folded
                mbar_wait(&full[b], par);
                if (issuer_fence) fence_proxy_async_smem();               // alternative measured in profiles/r02_pg_modes.txt (off)
                tc_fence_after();
                unsigned char* buf = smem + b * kBufBytes;
                const uint32_t d = tmem_base + (uint32_t)(set * kN);
#pragma unroll
                for (int kb = 0; kb < kPairs / 16; ++kb) {                // 16 pairs = 2 k-groups = 2 KB further into every MN block
                    const uint64_t a_hi = desc_mn(buf + kb * 2048), a_lo = desc_mn(buf + 8192 + kb * 2048);
                    const uint64_t b_hi = desc_mn(buf + oA + kb * 2048), b_lo = desc_mn(buf + oA + 16384 + kb * 2048);
                    mma_bf16(d, a_lo, b_hi, idesc, (opens && kb == 0) ? 0u : 1u);      // small terms first
                    mma_bf16(d, a_hi, b_lo, idesc, 1u);
                    mma_bf16(d, a_hi, b_hi, idesc, 1u);
                }
                mma_commit(&empty[b]);
                if (closes) mma_commit(&chunk_done[set]);
                if (++b == kBuf) { b = 0; par ^= 1; }
            }
        }
    } else if (warp >= kFoldWarp) {
        // ================================ fold: finished chunks TMEM -> fp32 sums in shared memory ====================
        // first fold warp: lane m = row m of D (z index k); second: lane 0 = row 32 (the ones row: dL/db)
        const bool active = (warp == kFoldWarp) || (lane == 0);
        const int m = (warp == kFoldWarp) ? lane : 32;
        for (int64_t chunk = 0; chunk < n_chunks; ++chunk) {
            const int set = (int)(chunk & 1);
            if (active) {
                mbar_wait_relaxed(&chunk_done[set], (uint32_t)((chunk >> 1) & 1), 500);
                tc_fence_after();
            }
            __syncwarp();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(set * kN);
#pragma unroll 1
            for (int q = 0; q < kN / 32; ++q) {
                uint32_t v[32];
                tmem_ld32_issue(taddr + 32 * q, v);                       // warp-wide (.sync.aligned); idle lanes discard
                tmem_ld32_wait(v);
                if (active) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[(32 * q + i) * 33 + m] += __uint_as_float(v[i]);
                }
            }
            if (active) {
                tc_fence_before();
                mbar_arrive(&set_free[set]);
            }
        }
    } else {
        // ================================ producers: warp = (MN block g, part of it), lane = pair =======================
        const int g = warp & 3, part = warp >> 2, p = lane;
        const int x = p & 7;                                              // 128-byte swizzle: chunk c of row p sits at c ^ x
        int s = 0, b = 0;
        uint32_t par_s = 0, par_b = 1;                                    // "slot s is full" / "buffer b is empty" (empty at the start)
        for (int64_t j = 0; j < n_mine; ++j) {
            const int e = walk_e;
            advance();
            const float we = stage_weight[e], fr = stage_frac[e];
            mbar_wait(&stg_full[s], par_s);
            const unsigned char* stg = smem + oStg + s * kStgBytes;
            // a[8g + HU part .. + HU) and z likewise: 4 HU bytes of chunk 2g + (HU part) / 4 of the pair's row
            const uint32_t az_off = (uint32_t)(p * 128 + (((2 * g + (HU * part) / 4) ^ x) << 4) + 4 * ((HU * part) & 3));
            float av[HU], zv[HU];
            if (HU == 4) {
                const float4 a0 = *reinterpret_cast<const float4*>(stg + az_off), z0 = *reinterpret_cast<const float4*>(stg + 4096 + az_off);
                av[0] = a0.x; av[1] = a0.y; av[HU - 2] = a0.z; av[HU - 1] = a0.w;
                zv[0] = z0.x; zv[1] = z0.y; zv[HU - 2] = z0.z; zv[HU - 1] = z0.w;
            } else {
                const float2 a0 = *reinterpret_cast<const float2*>(stg + az_off), z0 = *reinterpret_cast<const float2*>(stg + 4096 + az_off);
                av[0] = a0.x; av[1] = a0.y;
                zv[0] = z0.x; zv[1] = z0.y;
            }
            float dx[8];
            if (cubic) {
                const unsigned char* row = stg + 8192 + p * 128;          // [a | b | 2c | 3d]
                const float4 b0 = *reinterpret_cast<const float4*>(row + ((2 ^ x) << 4)), b1 = *reinterpret_cast<const float4*>(row + ((3 ^ x) << 4));
                const float4 c0 = *reinterpret_cast<const float4*>(row + ((4 ^ x) << 4)), c1 = *reinterpret_cast<const float4*>(row + ((5 ^ x) << 4));
                const float4 d0 = *reinterpret_cast<const float4*>(row + ((6 ^ x) << 4)), d1 = *reinterpret_cast<const float4*>(row + ((7 ^ x) << 4));
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int c = 0; c < C; ++c)               // interpolation_cubic.py:331-336, then the stage's Runge-Kutta weight
                    dx[c] = __fadd_rn(bb[c], __fmul_rn(__fadd_rn(cc[c], __fmul_rn(dd[c], fr)), fr)) * we;
            } else {
                const float4 b0 = *reinterpret_cast<const float4*>(stg + 8192 + p * 32), b1 = *reinterpret_cast<const float4*>(stg + 8192 + p * 32 + 16);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int c = 0; c < C; ++c) dx[c] = bb[c] * we;
            }
            mbar_arrive(&stg_free[s]);                                    // staging read into registers: the TMA warp may refill it
            mbar_wait(&empty[b], par_b);                                  // the MMAs that last read this buffer are done
            unsigned char* buf = smem + b * kBufBytes;
            // U[pair][n = (8g + hh) * 8 + c] = a[hh] * dx[c]: one 16-byte chunk per hidden unit, MN block g, chunk hh
            // (rows beyond n_paths arrived as zeros)
#pragma unroll
            for (int i = 0; i < HU; ++i) {
                const int hh = HU * part + i;
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_bf2(av[i] * dx[2 * q], av[i] * dx[2 * q + 1], hi[q], lo[q]);
                const uint32_t off = mn_off(g, p, hh);
                *reinterpret_cast<uint4*>(buf + oA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(buf + oA + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            {   // Zx[pair][m = 8g + HU part .. + HU) = z: 2 HU bytes of chunk g of MN block 0
                uint32_t hi[HU / 2], lo[HU / 2];
#pragma unroll
                for (int q = 0; q < HU / 2; ++q) split_bf2(zv[2 * q], zv[2 * q + 1], hi[q], lo[q]);
                const uint32_t off = mn_off(0, p, g) + 2 * HU * part;
                if (HU == 4) {
                    *reinterpret_cast<uint2*>(buf + off) = make_uint2(hi[0], hi[HU / 2 - 1]);
                    *reinterpret_cast<uint2*>(buf + 8192 + off) = make_uint2(lo[0], lo[HU / 2 - 1]);
                } else {
                    *reinterpret_cast<uint32_t*>(buf + off) = hi[0];
                    *reinterpret_cast<uint32_t*>(buf + 8192 + off) = lo[0];
                }
            }
            if (!issuer_fence) fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&full[b]);
            if (++s == kStg) { s = 0; par_s ^= 1; }
            if (++b == kBuf) { b = 0; par_b ^= 1; }
        }
    }

    // ---- epilogue: the CTA's sums -> its slot of the partial sums (field_vjp_reduce_kernel adds the slots up) -----------
    tc_fence_before();
    __syncthreads();
    float* mine = scratch + (size_t)blockIdx.x * kParams;
    for (int e = tid; e < kN * H; e += kThreads) {                        // dL/dW[n][k] = D[k][n]
        const int n = e >> 5, k = e & 31;
        mine[e] = acc[n * 33 + k];
    }
    for (int n = tid; n < kN; n += kThreads) mine[H * C * H + n] = acc[n * 33 + 32];
    if (warp == kIssuerWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace pg2

int param_grad_bf16_grid(int64_t n_paths, int64_t n_stage_total) {
    const int64_t items = ((n_paths + pg2::kPairs - 1) / pg2::kPairs) * n_stage_total;
    int64_t g = sm_count();
    if (g > items) g = items;
    if (g < 1) g = 1;
    return (int)g;
}

int param_grad_bf16_f32(const float* control, int control_kind, int64_t n_rows, const float* z_stages, const float* a_stages,
                        const int32_t* stage_index, const float* stage_frac, const float* stage_weight, int n_stage_total,
                        float* scratch, int64_t n_paths, int grid, cudaStream_t stream) {
    // 8 producer warps, 3 operand buffers + 4 staging slots, proxy fence in the producers: the A/B of the alternatives (16 warps,
    // fence in the issuer, 2 + 7 / 3 + 3 rings) is in profiles/r02_pg_modes.txt -- all within 10 %, this one fastest
    constexpr int kQ = 2, kBufs = 3, kSlots = 4;
    constexpr int smem = pg2::smem_bytes(kBufs, kSlots) + 1024;   // + slack for the 1024-byte alignment of the tiles
    alignas(64) CUtensorMap rows_map;
    const int row_floats = (control_kind == TCDE_CONTROL_CUBIC) ? 4 * pg2::C : pg2::C;
    const int rc = tc::make_rows_tensor_map(&rows_map, control, n_paths, n_rows, row_floats, pg2::kPairs);
    TCDE_CHECK_SUPPORTED(rc == 0, "parameter gradients: cuTensorMapEncodeTiled failed (%d)", rc);
    alignas(64) CUtensorMap a_map, z_map;
    const int rca = tc::make_stage_tensor_map(&a_map, const_cast<float*>(a_stages), n_paths, n_stage_total, pg2::kPairs);
    const int rcz = tc::make_stage_tensor_map(&z_map, const_cast<float*>(z_stages), n_paths, n_stage_total, pg2::kPairs);
    TCDE_CHECK_SUPPORTED(rca == 0 && rcz == 0, "parameter gradients: cuTensorMapEncodeTiled failed (%d, %d) for the stage trajectories", rca, rcz);
    auto kern = pg2::param_grad_bf16_kernel<kQ, kBufs, kSlots>;
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128 * kQ + 128, smem, stream>>>(control, control_kind, n_rows, z_stages, a_stages, stage_index, stage_frac, stage_weight,
                                                 n_stage_total, scratch, n_paths, rows_map, a_map, z_map, 0);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

}  // namespace tcde
