// The parameter gradients of a whole fixed-step backward solve on the tensor cores (sm_100a).
//
//   dL/dW[hC+c][k] = sum_e sum_p  w_e a_e[p][h] dX_e[p][c] * z_e[p][k]          (e: stages, p: paths)
//   dL/db[hC+c]    = sum_e sum_p  w_e a_e[p][h] dX_e[p][c]
//
// is ONE matrix product with the (stage, path) pairs as the reduction dimension:  G = U^T [256 x SP] . Zx [SP x 33],
// U[(e,p)][hC+c] = w_e a dX (formed on the fly, never stored), Zx = (z | 1).  1.1 TFLOP at the BASELINE shapes --
// 55 ms on the CUDA cores (param_grad_kernel in solve_vjp.cu), a few ms here.  Same recipe as the forward solve:
// tcgen05.mma kind::tf32 with the 3xTF32 split (lo.hi + hi.lo + hi.hi) for fp32-class accuracy, fp32 accumulators
// in TMEM for the whole launch.
//
// CTA = 4 producer warps + 1 issuer warp, persistent over work items (stage, block of 32 paths = one K = 32 slab):
//   * producers: lane = path of the block, warp g = hidden units 8g .. 8g+7.  A thread fetches its 32 bytes of a,
//     its 32 bytes of z and its path's spline row with cp.async three items ahead (only its own data: no barrier
//     needed to read it back), forms 64 products a[h] * (w dX[c]), splits them into hi / lo and writes them as
//     column p of the K-major 128B-swizzled A tiles (rows = (h, c), two halves of 128 rows), and its 8 z values as
//     column p of the B tile (rows = k; row 32 is all ones -> column 32 of the accumulator is dL/db);
//   * issuer: per item and half 12 MMAs M = 128, N = 48, K = 8, then tcgen05.commit frees the operand buffer.
//   Two operand buffers: the producers fill one while the tensor pipe reads the other.
// The tensor core adds into its fp32 accumulator with truncation, harmless over the 13 MMAs of a solve stage but a
// systematic drift over the ~10^5 accumulations of this launch (measured: 2.6e-3 relative on dL/dW).  The
// accumulation therefore runs in chunks of kChunk items into two alternating TMEM accumulator sets; the producers
// fold each finished chunk into registers with ordinary round-to-nearest adds (one chunk behind the tensor pipe,
// so nobody waits).  At the end every producer thread owns one row of dL/dW (both halves) and stores it to the
// per-CTA partial sums that field_vjp_reduce_kernel adds up.
#include "umma.cuh"

namespace tcde {
namespace pgu {

using namespace umma;

constexpr int H = 32, C = 8;
constexpr int kPaths = 32;                 // paths per item == K of one operand slab
constexpr int kN = 48;                     // z (32) | ones (1) | zero padding to a multiple of 16
constexpr int kProd = 128;
constexpr int kThreads = kProd + 32;
constexpr int kDepth = 3;                  // cp.async items in flight per thread
constexpr int kChunk = 8;                  // items accumulated in TMEM before the sum moves to registers
constexpr int kVec = 10;                   // 16-byte pieces per thread and item: a 2, z 2, spline row 6
constexpr int kParams = H * C * H + H * C;

// shared memory map (bytes); operand tiles on 1024-byte boundaries
constexpr int oAhi = 0;                                  // [2 buffers][2 halves][128 rows][128 B]
constexpr int oAlo = oAhi + 2 * 2 * 128 * 128;
constexpr int oBhi = oAlo + 2 * 2 * 128 * 128;           // [2 buffers][8 KB] (48 rows x 128 B used)
constexpr int oBlo = oBhi + 2 * 8192;
constexpr int oRing = oBlo + 2 * 8192;                   // [kDepth][kVec][kProd] x 16 B
constexpr int oBars = oRing + kDepth * kVec * kProd * 16;
constexpr int kSmem = oBars + 128;                       // 8 mbarriers + the TMEM slot

__global__ void __launch_bounds__(kThreads, 1)
param_grad_umma_kernel(const float* __restrict__ control, int control_kind, int64_t n_rows,
                       const float* __restrict__ z_stages, const float* __restrict__ a_stages,
                       const int32_t* __restrict__ stage_index, const float* __restrict__ stage_frac,
                       const float* __restrict__ stage_weight, int n_stage_total, float* __restrict__ scratch,
                       int64_t n_paths) {
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + oBars);      // [2]
    uint64_t* empty = full + 2;                                       // [2]
    uint64_t* chunk_done = empty + 2;                                 // [2 accumulator sets]
    uint64_t* set_free = chunk_done + 2;                              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(set_free + 2);
    const int tid = threadIdx.x, warp = tid >> 5;

    // constant rows of the B tiles: row 32 = ones (hi), everything else of rows 32 .. 47 zero
    for (int e = tid; e < 2 * 16 * 32; e += kThreads) {
        const int buf = e >> 9, row = 32 + ((e >> 5) & 15), k = e & 31;
        reinterpret_cast<float*>(smem + oBhi + buf * 8192)[swz(row, k)] = (row == 32) ? 1.f : 0.f;
        reinterpret_cast<float*>(smem + oBlo + buf * 8192)[swz(row, k)] = 0.f;
    }
    if (tid == 0) {
        mbar_init(&full[0], kProd);
        mbar_init(&full[1], kProd);
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(&chunk_done[0], 1);
        mbar_init(&chunk_done[1], 1);
        mbar_init(&set_free[0], kProd);
        mbar_init(&set_free[1], kProd);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_blocks = (n_paths + kPaths - 1) / kPaths;
    const int64_t n_items = n_blocks * n_stage_total;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t n_mine = first < n_items ? (n_items - first + stride - 1) / stride : 0;

    if (warp == 4) {
        // ================================ MMA issuer ==============================================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int64_t j = 0; j < n_mine; ++j) {
            const int buf = (int)(j & 1);
            const int64_t chunk = j / kChunk;
            const int set = (int)(chunk & 1);
            const bool opens = (j % kChunk) == 0, closes = (j % kChunk) == kChunk - 1 || j == n_mine - 1;
            if (opens && chunk >= 2) mbar_wait(&set_free[set], (uint32_t)(((chunk >> 1) + 1) & 1));   // chunk - 2 was read
            mbar_wait(&full[buf], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if ((tid & 31) == 0) {
                const uint64_t dbh = make_desc(smem + oBhi + buf * 8192), dbl = make_desc(smem + oBlo + buf * 8192);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint64_t dah = make_desc(smem + oAhi + (buf * 2 + half) * 16384);
                    const uint64_t dal = make_desc(smem + oAlo + (buf * 2 + half) * 16384);
                    const uint32_t d = tmem_base + (uint32_t)(set * 128 + half * 64);
                    // small terms first; a k-block is 8 tf32 = 32 bytes = +2 in the descriptor's address field
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dal + 2 * kb, dbh + 2 * kb, idesc, (!opens || kb > 0) ? 1u : 0u);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, dbl + 2 * kb, idesc, 1);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, dbh + 2 * kb, idesc, 1);
                }
                mma_commit(&empty[buf]);
                if (closes) mma_commit(&chunk_done[set]);
            }
            __syncwarp();
        }
    } else {
        // ================================ producers ==============================================
        const int p = tid & 31, g = tid >> 5;
        const bool cubic = (control_kind == TCDE_CONTROL_CUBIC);
        const int row_stride = cubic ? 4 * C : C;
        float4* ring = reinterpret_cast<float4*>(smem + oRing);
        // schedule entries travel ahead of the copies that need them: queue slot i belongs to item j + i
        int idx_q[2 * kDepth];
        float w_q[kDepth], f_q[kDepth];
        auto stage_of = [&](int64_t j) { return (int)((first + j * stride) / n_blocks); };
#pragma unroll
        for (int i = 0; i < 2 * kDepth; ++i) idx_q[i] = (i < n_mine) ? stage_index[stage_of(i)] : 0;

        auto issue = [&](int64_t j, int slot, int index) {       // this thread's 160 bytes of item j -> ring[slot]
            if (j < n_mine) {
                const int64_t item = first + j * stride;
                const int e = (int)(item / n_blocks);
                int64_t path = (item - (int64_t)e * n_blocks) * kPaths + p;
                if (path >= n_paths) path = n_paths - 1;
                const float* as = a_stages + ((int64_t)e * n_paths + path) * H + 8 * g;
                const float* zs = z_stages + ((int64_t)e * n_paths + path) * H + 8 * g;
                const float* cr = control + (path * n_rows + index) * row_stride + (cubic ? C : 0);
                float4* dst = ring + (size_t)slot * kVec * kProd + tid;
                cp_async16(dst + 0 * kProd, as);
                cp_async16(dst + 1 * kProd, as + 4);
                cp_async16(dst + 2 * kProd, zs);
                cp_async16(dst + 3 * kProd, zs + 4);
                const int parts = cubic ? 6 : 2;
                for (int v = 0; v < parts; ++v) cp_async16(dst + (4 + v) * kProd, cr + 4 * v);
            }
            cp_async_commit();                                     // always: keeps the group count uniform
        };
#pragma unroll
        for (int i = 0; i < kDepth; ++i) {
            w_q[i] = (i < n_mine) ? stage_weight[stage_of(i)] : 0.f;
            f_q[i] = (i < n_mine) ? stage_frac[stage_of(i)] : 0.f;
            issue(i, i, idx_q[i]);
        }

        // this thread's rows (h, c) = tid and 128 + tid of dL/dW: 32 columns k, then dL/db
        float acc[2][H + 1];
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int q = 0; q <= H; ++q) acc[half][q] = 0.f;
        auto fold = [&](int64_t chunk) {                          // finished chunk: TMEM -> registers, set released
            const int set = (int)(chunk & 1);
            mbar_wait(&chunk_done[set], (uint32_t)((chunk >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[kN];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(set * 128 + half * 64);
#pragma unroll
                for (int q = 0; q < kN / 16; ++q) tmem_ld16_issue(taddr + 16 * q, v + 16 * q);
#pragma unroll
                for (int q = 0; q < kN / 16; ++q) tmem_ld16_wait(v + 16 * q);
#pragma unroll
                for (int q = 0; q <= H; ++q) acc[half][q] += __uint_as_float(v[q]);
            }
            tc_fence_before();
            mbar_arrive(&set_free[set]);
        };

        int slot = 0;
        for (int64_t j = 0; j < n_mine; ++j) {
            const int buf = (int)(j & 1);
            cp_async_wait<kDepth - 1>();                            // item j's copies have landed
            const float4* src = ring + (size_t)slot * kVec * kProd + tid;
            const float4 a0 = src[0], a1 = src[kProd], z0 = src[2 * kProd], z1 = src[3 * kProd];
            const float4 b0 = src[4 * kProd], b1 = src[5 * kProd];
            float4 c0 = b0, c1 = b1, d0 = b0, d1 = b1;
            if (cubic) { c0 = src[6 * kProd]; c1 = src[7 * kProd]; d0 = src[8 * kProd]; d1 = src[9 * kProd]; }
            const float we = w_q[0], fr = f_q[0];
            // refill the slot with item j + kDepth and advance the queues
            issue(j + kDepth, slot, idx_q[kDepth]);
#pragma unroll
            for (int i = 0; i + 1 < kDepth; ++i) { w_q[i] = w_q[i + 1]; f_q[i] = f_q[i + 1]; }
            {
                const int64_t jn = j + kDepth;
                w_q[kDepth - 1] = (jn < n_mine) ? stage_weight[stage_of(jn)] : 0.f;
                f_q[kDepth - 1] = (jn < n_mine) ? stage_frac[stage_of(jn)] : 0.f;
            }
#pragma unroll
            for (int i = 0; i + 1 < 2 * kDepth; ++i) idx_q[i] = idx_q[i + 1];
            {
                const int64_t jn = j + 2 * kDepth;
                idx_q[2 * kDepth - 1] = (jn < n_mine) ? stage_index[stage_of(jn)] : 0;
            }
            slot = (slot + 1 == kDepth) ? 0 : slot + 1;

            const int64_t item = first + j * stride;
            const int e = (int)(item / n_blocks);
            const bool live = (item - (int64_t)e * n_blocks) * kPaths + p < n_paths;
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float dx[8];
            {
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int c = 0; c < C; ++c) {      // interpolation_cubic.py:331-336, then the stage's weight
                    const float d = cubic ? __fadd_rn(bb[c], __fmul_rn(__fadd_rn(cc[c], __fmul_rn(dd[c], fr)), fr)) : bb[c];
                    dx[c] = d * we;
                }
            }
            if (!live) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { av[i] = 0.f; zv[i] = 0.f; }
            }
            if (j >= 2) mbar_wait(&empty[buf], (uint32_t)(((j >> 1) + 1) & 1));      // the MMAs of item j - 2 are done
            float* a_hi = reinterpret_cast<float*>(smem + oAhi + (buf * 2 + (g >> 1)) * 16384);
            float* a_lo = reinterpret_cast<float*>(smem + oAlo + (buf * 2 + (g >> 1)) * 16384);
            float* b_hi = reinterpret_cast<float*>(smem + oBhi + buf * 8192);
            float* b_lo = reinterpret_cast<float*>(smem + oBlo + buf * 8192);
#pragma unroll
            for (int hh = 0; hh < 8; ++hh) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float u = av[hh] * dx[c];
                    const float hi = tf32_hi(u);
                    const int row = ((g & 1) * 8 + hh) * C + c;          // row (h, c) within its half of 128
                    const uint32_t off = swz(row, p);
                    a_hi[off] = hi;
                    a_lo[off] = u - hi;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float hi = tf32_hi(zv[i]);
                const uint32_t off = swz(8 * g + i, p);
                b_hi[off] = hi;
                b_lo[off] = zv[i] - hi;
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&full[buf]);
            // a chunk behind the tensor pipe: when the last item of chunk c has been handed over, fold chunk c - 1
            if (((j % kChunk) == kChunk - 1 || j == n_mine - 1) && j / kChunk >= 1) fold(j / kChunk - 1);
        }
        cp_async_wait<0>();
        if (n_mine > 0) fold((n_mine - 1) / kChunk);

        // ---- epilogue: thread = row of dL/dW (both halves) -> per-CTA partial sums ----
        float* mine = scratch + (size_t)blockIdx.x * kParams;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int n = half * 128 + tid;                              // = h * C + c
            float4* dst = reinterpret_cast<float4*>(mine + (size_t)n * H);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4)
                dst[k4] = make_float4(acc[half][4 * k4], acc[half][4 * k4 + 1], acc[half][4 * k4 + 2], acc[half][4 * k4 + 3]);
            mine[H * C * H + n] = acc[half][H];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 256);
}

}  // namespace pgu

int param_grad_umma_grid(int64_t n_paths, int64_t n_stage_total) {
    const int64_t items = ((n_paths + pgu::kPaths - 1) / pgu::kPaths) * n_stage_total;
    int64_t g = sm_count();
    if (g > items) g = items;
    if (g < 1) g = 1;
    return (int)g;
}

int param_grad_umma_f32(const float* control, int control_kind, int64_t n_rows, const float* z_stages,
                        const float* a_stages, const int32_t* stage_index, const float* stage_frac,
                        const float* stage_weight, int n_stage_total, float* scratch, int64_t n_paths, int grid,
                        cudaStream_t stream) {
    constexpr int smem = pgu::kSmem + 1024;                // slack for the 1024-byte alignment of the tiles
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(pgu::param_grad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pgu::param_grad_umma_kernel<<<grid, pgu::kThreads, smem, stream>>>(control, control_kind, n_rows, z_stages, a_stages,
                                                                      stage_index, stage_frac, stage_weight,
                                                                      n_stage_total, scratch, n_paths);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

}  // namespace tcde
