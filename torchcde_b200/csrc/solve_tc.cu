// Hot path (ii), round 2: the fused fixed-step CDE solve on tcgen05 with the Runge-Kutta state in registers.
//
// What changed against solve_umma.cu (round 1, kept as variant 2) and why -- profiles/r01_umma_trace.txt showed one
// tile-stage as a 4,065-cycle serial chain of which only 1,664 are tensor work (13 MMAs x 128 cycles: the issue loop
// finishes after ~1,330 because the MMA queue holds the last three), so two ping-ponging tiles kept the pipe 82 % busy:
//   * the slopes k1 and k2 (+k3) never leave registers (one CTA per SM: the register file is there to be used): the
//     64 LDS + 32-64 STS per thread and stage of the old "park" and their latency are gone;
//   * MODE 1 halves the tensor work itself: operands are split into TWO FP16 pieces instead of two TF32 pieces.  An
//     FP16 MMA carries K = 16 per instruction at the cycle cost of a K = 8 TF32 one, so the three partial products
//     z_lo.W_hi + z_hi.W_lo + z_hi.W_hi are 6 MMAs instead of 12 (+1 for the bias in both modes).  FP16 has the same
//     11-bit significand as TF32 (hi + lo = 22 bits, error ~2^-22 like 3xTF32) but only 5 exponent bits, so every
//     path's row is scaled by its own power of two (exact) so that max(|z_k|, bias floor) lands in [2^13, 2^14); the
//     weights get one global power of two; elements more than 2^-27 below the row maximum lose relative -- not
//     absolute -- precision, which is what a dot product needs.  The scale is undone for free by folding its inverse
//     into the path's dX/dt before the contraction.  hi | lo of a row share ONE 128-byte swizzled A row (K = 64).
//
//   * no dedicated issuer warp: 9 warps are allocated like 12 (168 registers per thread), 8 warps get 255.  Lane 0 of a
//     tile's first warp issues the tile's MMAs as soon as the tile's 128 rows have arrived on a_ready[t].
//
// CTA anatomy: two tiles of 128 paths, thread = path (256 threads); accumulators [128 x 256] fp32 per tile in TMEM (all
// 512 columns); a_ready[t] (128 arrivals) / d_ready[t] (tcgen05.commit) mbarriers.
#include "tc_common.cuh"

namespace tcde {

namespace tc {

using namespace umma;

constexpr int kH = 32;
constexpr int kC = 8;
constexpr int kN = kH * kC;       // 256 accumulator columns per tile
constexpr int kTile = 128;
constexpr int kTiles = 2;
constexpr int kRowThreads = kTile * kTiles;     // 256 row threads: thread = path
constexpr int kIssuerThreads = 128;            // ISSUER variant: one more warpgroup (warp 8 issues, 9-11 only donate registers)

// shared-memory map (bytes).  MODE 0: TF32 hi / lo tiles (K = 32 floats = one 128-byte row each);
// MODE 1: one FP16 tile per operand, row = [hi(32) | lo(32)] halves = 128 bytes.
template <int MODE> struct Smem {
    static constexpr int b_bytes = (MODE == 0 ? 2 : 1) * kN * 128;
    static constexpr int a_tile_bytes = (MODE == 0 ? 2 : 1) * kTile * 128;          // per tile: hi (+ lo tile in MODE 0)
    static constexpr int b = 0;
    static constexpr int a = b + b_bytes;
    static constexpr int raw = a + kTiles * a_tile_bytes;                            // [kTiles][2 buffers][128 rows x 128 B] (TMA, swizzled)
    static constexpr int raw_buf = kTile * 128;
    static constexpr int b_aug = raw + kTiles * 2 * raw_buf;                         // N rows x 32 B (no swizzle, K-major)
    static constexpr int a_aug = b_aug + kN * 32;                                    // [kTiles][128 rows x 32 B] (MODE 0: one 8-row group)
    static constexpr int red = a_aug + kTiles * kTile * 32;                          // 64 floats of reduction scratch
    static constexpr int bars = red + 256;                                           // a_ready[2], d_ready[2], raw_full[2][2], tmem slot
    static constexpr int total = bars + 128;
};

__device__ __forceinline__ void reg_dealloc_40() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::: "memory"); }
__device__ __forceinline__ void reg_alloc_232() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n" ::: "memory"); }

// The MMAs of one tile-stage, issued by ONE thread: D[128 x 256] = split(A) . split(B)^T + bias, then commit -> d_ready
template <int MODE>
__device__ __forceinline__ void issue_tile(uint32_t d, unsigned char* smem, unsigned char* a_tile, unsigned char* a_aug, uint64_t* d_ready) {
    using S = Smem<MODE>;
    // instruction descriptor: D = F32 (bit 4), A/B format (bits 7 / 10: 2 = TF32, 0 = F16), N >> 3 (bit 17), M >> 4 (bit 24)
    constexpr uint32_t fmt = (MODE == 0) ? 2u : 0u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    const uint64_t db = make_desc(smem + S::b);
    // K-major no-swizzle descriptors: LBO = 128 B (next 16 bytes of K), SBO = 256 B (next 8 rows), version 1
    const uint64_t db_aug = (uint64_t)((smem_u32(smem + S::b_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
    if (MODE == 0) {
        const uint64_t db_lo = make_desc(smem + S::b + kN * 128);
        const uint64_t dah = make_desc(a_tile);
        const uint64_t dal = make_desc(a_tile + kTile * 128);
        const uint64_t da_aug = (uint64_t)((smem_u32(smem + S::a_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (0ull << 32) | (1ull << 46);
        // small terms first; each K block is 8 tf32 = 32 bytes = +2 in the descriptor's address field
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dal + 2 * kb, db + 2 * kb, idesc, kb > 0);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, db_lo + 2 * kb, idesc, 1);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, db + 2 * kb, idesc, 1);
        mma_tf32(d, da_aug, db_aug, idesc, 1);                // + bias (1 * bias_hi + 1 * bias_lo)
    } else {
        const uint64_t da = make_desc(a_tile);
        const uint64_t da_aug = (uint64_t)((smem_u32(a_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
        // K block = 16 halves = 32 bytes = +2; A row = [hi: blocks 0,1 | lo: blocks 2,3], B row likewise
        mma_f16(d, da + 4, db + 0, idesc, 0);                 // z_lo . W_hi
        mma_f16(d, da + 6, db + 2, idesc, 1);
        mma_f16(d, da + 0, db + 4, idesc, 1);                 // z_hi . W_lo
        mma_f16(d, da + 2, db + 6, idesc, 1);
        mma_f16(d, da + 0, db + 0, idesc, 1);                 // z_hi . W_hi
        mma_f16(d, da + 2, db + 2, idesc, 1);
        mma_f16(d, da_aug, db_aug, idesc, 1);                 // + row_scale * bias
    }
    mma_commit(d_ready);
}

// ISSUER = false: 256 threads, lane 0 of each tile's first warp issues its tile's MMAs (255 registers per thread, but the
// ~650 cycles of issue sit on that warp's chain).  ISSUER = true: 384 threads, warp 8 does nothing but wait and issue;
// its warpgroup gives its registers to the row warps (setmaxnreg: 40 / 232).
template <int MODE, bool ISSUER, bool TRACE, bool DUMP>
__global__ void __launch_bounds__(kRowThreads + (ISSUER ? kIssuerThreads : 0), 1) cdeint_tc_kernel(const UmmaArgs a, const __grid_constant__ CUtensorMap rows_map) {
    constexpr int kThreads = kRowThreads + (ISSUER ? kIssuerThreads : 0);
    constexpr int kAllocWarp = ISSUER ? 8 : 0;
    using S = Smem<MODE>;
    using E = exact<float>;
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(smem + S::bars);
    uint64_t* d_ready = a_ready + kTiles;
    uint64_t* raw_full = d_ready + kTiles;                    // [tile][buffer]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_full + 2 * kTiles);
    float* red = reinterpret_cast<float*>(smem + S::red);
    const bool cubic = (a.control_kind == TCDE_CONTROL_CUBIC);
    const int row_floats = cubic ? 4 * kC : kC;
    const uint32_t row_bytes_tile = (uint32_t)(kTile * row_floats * 4);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int64_t cta_path0 = (int64_t)blockIdx.x * (kTile * kTiles);
    const int total = a.n_steps * a.n_stages;

    // ---- one-time setup: operand B (the weights, split) and the bias K-block -------------------------------------
    float w_scale = 1.f, inv_w_scale = 1.f, beta = 0.f;       // MODE 1: global power-of-two scales
    if (MODE == 1) {
        float wmax = 0.f, bmax = 0.f;
        for (int e = tid; e < kN * kH; e += kThreads) wmax = fmaxf(wmax, fabsf(a.weight[e]));
        for (int e = tid; e < kN; e += kThreads) bmax = fmaxf(bmax, fabsf(a.bias[e]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
            bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        }
        if ((tid & 31) == 0) { red[warp] = wmax; red[16 + warp] = bmax; }
        __syncthreads();
        wmax = 0.f; bmax = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) { wmax = fmaxf(wmax, red[w]); bmax = fmaxf(bmax, red[16 + w]); }
        // exponents clamped so that every derived power of two is a normal fp32 number; an all-zero weight takes the
        // bias's exponent (the product is then the bias alone)
        const int eb = min(max(exponent_of(bmax), 40), 215);
        const int ew = (wmax > 0.f) ? min(max(exponent_of(wmax), 40), 215) : (bmax > 0.f ? eb : 127);
        w_scale = pow2_biased(127 + 13 - (ew - 127));                                // max |W| * w_scale in [2^13, 2^14)
        inv_w_scale = pow2_biased(127 - 13 + (ew - 127));
        if (bmax > 0.f) beta = pow2_biased(min(max(127 + eb - ew, 2), 250));         // 2^(ex(bmax) - ex(wmax))
    }
    if (MODE == 0) {
        float* b_hi = reinterpret_cast<float*>(smem + S::b);
        float* b_lo = b_hi + kN * 32;
        for (int e = tid; e < kN * kH; e += kThreads) {
            const int n = e >> 5, k = e & 31;
            const float w = a.weight[e];                     // weight[n][k], n = h*C + c
            const float hi = tf32_hi(w);
            b_hi[swz(n, k)] = hi;
            b_lo[swz(n, k)] = w - hi;
        }
        float* a_aug = reinterpret_cast<float*>(smem + S::a_aug);
        float* b_aug = reinterpret_cast<float*>(smem + S::b_aug);
        for (int e = tid; e < 8 * 8; e += kThreads) {
            const int row = e >> 3, k = e & 7;
            a_aug[(k >> 2) * 32 + row * 4 + (k & 3)] = (k < 2) ? 1.f : 0.f;         // (1, 1, 0...): SBO = 0 serves all rows
        }
        for (int e = tid; e < kN * 8; e += kThreads) {
            const int row = e >> 3, k = e & 7;
            const float bv = a.bias[row];
            const float hi = tf32_hi(bv);
            b_aug[(row >> 3) * 64 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3)] = (k == 0) ? hi : (k == 1) ? (bv - hi) : 0.f;
        }
    } else {
        __half* bt = reinterpret_cast<__half*>(smem + S::b);
        for (int e = tid; e < kN * kH; e += kThreads) {
            const int n = e >> 5, k = e & 31;
            const float w = a.weight[e] * w_scale;
            const __half hi = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(hi));
            // row n = 128 bytes: halves 0..31 = hi, 32..63 = lo; 16-byte chunk c (8 halves) sits at c ^ (n & 7)
            bt[n * 64 + ((((k >> 3)) ^ (n & 7)) << 3) + (k & 7)] = hi;
            bt[n * 64 + ((((k >> 3) + 4) ^ (n & 7)) << 3) + (k & 7)] = lo;
        }
        unsigned char* b_aug = smem + S::b_aug;
        unsigned char* a_aug = smem + S::a_aug;
        const float bias_scale = (beta > 0.f) ? w_scale / beta : 0.f;                // bias * w_scale / beta in [2^13, 2^14)
        for (int row = tid; row < kN; row += kThreads) {
            const float bv = a.bias[row] * bias_scale;
            const __half hi = __float2half_rn(bv);
            const __half lo = __float2half_rn(bv - __half2float(hi));
            uint4 c0 = make_uint4((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16), 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(b_aug + aug_off(row, 0)) = c0;
            *reinterpret_cast<uint4*>(b_aug + aug_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int e = tid; e < kTiles * kTile * 2; e += kThreads)                     // per-row scale goes to k = 0, 1 each stage
            *reinterpret_cast<uint4*>(a_aug + (e >> 8) * (kTile * 32) + aug_off((e >> 1) & 127, e & 1)) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int t = 0; t < kTiles; ++t) {
            mbar_init(&a_ready[t], kTile);
            mbar_init(&d_ready[t], 1);
            mbar_init(&raw_full[2 * t], 1);
            mbar_init(&raw_full[2 * t + 1], 1);
        }
        fence_barrier_init();
        tma_prefetch_desc(&rows_map);
    }
    if (warp == kAllocWarp) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    bool tile_live[kTiles];
#pragma unroll
    for (int t = 0; t < kTiles; ++t) tile_live[t] = (cta_path0 + (int64_t)t * kTile) < a.n_paths;

    // the spline rows of stage `st` for tile t -> raw[t][st & 1]; completes raw_full[t][st & 1] (one thread calls this)
    auto fetch_rows = [&](int t, int st, int interval) {
        if (a.debug & 1) return;
        uint64_t* bar = &raw_full[2 * t + (st & 1)];
        mbar_expect_tx(bar, row_bytes_tile);
        tma_load_2d(smem + S::raw + (2 * t + (st & 1)) * S::raw_buf, &rows_map, interval * row_floats,
                    (int)(cta_path0 + (int64_t)t * kTile), bar);
    };

    if (ISSUER && warp >= kRowThreads / 32) {
        // ================================ MMA issuer warpgroup ================================
        reg_dealloc_40();
        // one issuer warp PER TILE (warps 8 and 9): a tile's MMAs start as soon as its own 128 rows have arrived, whatever the
        // other tile's issuer is doing (one warp serving both tiles serialised them: 2 x ~1,250 cycles per stage)
        const int t = warp - kRowThreads / 32;
        if (t < kTiles && tile_live[t]) {
            uint32_t phase = 0;
            int idx_next = (total > 1) ? a.stage_index[1] : 0;          // schedule entries are read a stage ahead of their use
            if ((tid & 31) == 0) fetch_rows(t, 0, a.stage_index[0]);
            for (int st = 0; st < total; ++st) {
                const int idx_fetch = idx_next;
                if (st + 2 < total) idx_next = a.stage_index[st + 2];
                mbar_wait(&a_ready[t], phase);
                phase ^= 1;
                tc_fence_after();
                if ((tid & 31) == 0) {
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 0] = clock64();
                    issue_tile<MODE>(tmem_base + (uint32_t)(t * kN), smem, smem + S::a + t * S::a_tile_bytes,
                                     smem + S::a_aug + t * kTile * 32, &d_ready[t]);
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 1] = clock64();
                    // the rows of the NEXT stage: its buffer was last read two stages ago, before the arrivals just waited for
                    if (st + 1 < total) fetch_rows(t, st + 1, idx_fetch);
                }
                __syncwarp();
            }
        }
    } else {
        if (ISSUER) reg_alloc_232();
        // ================================ row threads =========================================
        const int t = warp >> 2;                          // tile of this thread
        const int r = tid & (kTile - 1);                  // row (path) within the tile
        const int64_t path = cta_path0 + (int64_t)t * kTile + r;
        const bool live = path < a.n_paths;
        const int64_t lpath = live ? path : a.n_paths - 1;
        if (tile_live[t]) {
            unsigned char* a_tile = smem + S::a + t * S::a_tile_bytes;
            unsigned char* a_aug = smem + S::a_aug + t * kTile * 32;
            const unsigned char* raw_tile = smem + S::raw + 2 * t * S::raw_buf;      // two buffers, stage st in buffer st & 1
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * kN);
            const float sign = (a.sign < 0.f) ? -1.f : 1.f;
            float inv_scale = 1.f;                        // MODE 1: 1 / (row scale * weight scale) of the stage in flight

            // ---- ISSUER = false: the tile's MMAs are issued by lane 0 of the tile's first warp once all 128 rows have arrived
            const bool issuer_warp = !ISSUER && (warp & 3) == 0;
            uint32_t phase_a = 0;
            int idx_issue = (total > 1) ? a.stage_index[1] : 0;
            auto issue_stage = [&](int st) {
                const int idx_fetch = idx_issue;
                if (st + 2 < total) idx_issue = a.stage_index[st + 2];
                mbar_wait(&a_ready[t], phase_a);
                phase_a ^= 1;
                tc_fence_after();
                if ((tid & 31) == 0) {
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 0] = clock64();
                    if (st == 0) fetch_rows(t, 0, a.stage_index[0]);
                    issue_tile<MODE>(tmem_base + (uint32_t)(t * kN), smem, a_tile, a_aug, &d_ready[t]);
                    if (st + 1 < total) fetch_rows(t, st + 1, idx_fetch);
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 1] = clock64();
                }
                __syncwarp();
            };

            auto write_a = [&](const float* z, int stage_no) {   // next stage input -> split operand rows
                if (DUMP && live) {                       // ... and, for the adjoint, to the trajectory in HBM
                    float4* dst = reinterpret_cast<float4*>(a.stage_dump + ((int64_t)stage_no * a.n_paths + path) * kH);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_float4(z[4 * c4], z[4 * c4 + 1], z[4 * c4 + 2], z[4 * c4 + 3]);
                }
                if (MODE == 0) {
                    float* a_hi = reinterpret_cast<float*>(a_tile);
                    float* a_lo = a_hi + kTile * 32;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 hi, lo;
                        hi.x = tf32_hi(z[4 * c4 + 0]); hi.y = tf32_hi(z[4 * c4 + 1]);
                        hi.z = tf32_hi(z[4 * c4 + 2]); hi.w = tf32_hi(z[4 * c4 + 3]);
                        upk(sub2(pk(z[4 * c4 + 0], z[4 * c4 + 1]), pk(hi.x, hi.y)), lo.x, lo.y);
                        upk(sub2(pk(z[4 * c4 + 2], z[4 * c4 + 3]), pk(hi.z, hi.w)), lo.z, lo.w);
                        const uint32_t off = (uint32_t)r * 32u + (uint32_t)((c4 ^ (r & 7)) << 2);
                        *reinterpret_cast<float4*>(a_hi + off) = hi;
                        *reinterpret_cast<float4*>(a_lo + off) = lo;
                    }
                } else {
                    // the path's own power of two: max(|z_k|, bias floor) * s in [2^13, 2^14)
                    float mx[8];                          // tree instead of a 32-long dependent chain
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mx[k] = fmaxf(fmaxf(fabsf(z[k]), fabsf(z[k + 8])), fmaxf(fabsf(z[k + 16]), fabsf(z[k + 24])));
                    const float m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                                          fmaxf(fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])), beta));
                    const int e = min(max(exponent_of(m), 30), 224);
                    const float s = pow2_biased(127 + 13 - (e - 127));
                    inv_scale = pow2_biased(127 - 13 + (e - 127)) * inv_w_scale;
                    const f2 s2 = pk(s, s);
                    if (!(a.debug & 8)) {
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
                    const float sb = s * beta;            // <= 2^13 by construction; exact, flushes to 0 far below the row maximum
                    *reinterpret_cast<uint32_t*>(a_aug + aug_off(r, 0)) = pack_h2(sb, sb);
                    }
                }
                if (!(a.debug & 2)) fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&a_ready[t]);
            };
            auto write_out = [&](int j, const float* v) {
                if (!live) return;
                float4* dst = reinterpret_cast<float4*>(a.out + (path * a.n_out + j) * kH);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            };

            float y[kH], k1[kH], s23[kH];
            {
                const float4* zp = reinterpret_cast<const float4*>(a.z0 + lpath * kH);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 v = zp[c4];
                    y[4 * c4] = v.x; y[4 * c4 + 1] = v.y; y[4 * c4 + 2] = v.z; y[4 * c4 + 3] = v.w;
                }
#pragma unroll
                for (int h = 0; h < kH; ++h) { k1[h] = 0.f; s23[h] = 0.f; }
            }
            int jn = 0;
            int next_out = (a.n_out > 0) ? a.out_step[0] : 0x7fffffff;
            while (jn < a.n_out && next_out < 0) {
                write_out(jn, y);
                ++jn;
                next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
            }
            write_a(y, 0);
            if (issuer_warp) issue_stage(0);

            const float third = (float)(1.0 / 3.0);
            int step = 0, sub = 0;
            float dt = a.step_dt[0];
            float dt_next = (a.n_steps > 1) ? a.step_dt[1] : 0.f;
            float frac0 = a.stage_frac[0];                        // this stage's fraction
            float frac1 = (total > 1) ? a.stage_frac[1] : 0.f;    // next stage's, read a full stage ahead
            uint32_t phase = 0;
            for (int st = 0; st < total; ++st) {
                const bool more = st + 1 < total;
                // ---- in the MMA's shadow: dX/dt from the prefetched row (interpolation_cubic.py:331-336), times the
                //      sign of the time direction and (MODE 1) the inverse of the operand scales
                if (!(a.debug & 1)) mbar_wait(&raw_full[2 * t + (st & 1)], (uint32_t)((st >> 1) & 1));
                f2 dx2[kC / 2];
                {
                    const unsigned char* rows = raw_tile + (st & 1) * S::raw_buf;
                    if (cubic) {
                        // row r of the TMA box = [a | b | 2c | 3d], 16-byte chunk c at position c ^ (r & 7)
                        const unsigned char* row = rows + r * 128;
                        const int x = r & 7;
                        const float4 b0 = *reinterpret_cast<const float4*>(row + ((2 ^ x) << 4)), b1 = *reinterpret_cast<const float4*>(row + ((3 ^ x) << 4));
                        const float4 c0 = *reinterpret_cast<const float4*>(row + ((4 ^ x) << 4)), c1 = *reinterpret_cast<const float4*>(row + ((5 ^ x) << 4));
                        const float4 d0 = *reinterpret_cast<const float4*>(row + ((6 ^ x) << 4)), d1 = *reinterpret_cast<const float4*>(row + ((7 ^ x) << 4));
                        const f2 fr = pk(frac0, frac0);
                        dx2[0] = add2(pk(b0.x, b0.y), mul2(add2(pk(c0.x, c0.y), mul2(pk(d0.x, d0.y), fr)), fr));
                        dx2[1] = add2(pk(b0.z, b0.w), mul2(add2(pk(c0.z, c0.w), mul2(pk(d0.z, d0.w), fr)), fr));
                        dx2[2] = add2(pk(b1.x, b1.y), mul2(add2(pk(c1.x, c1.y), mul2(pk(d1.x, d1.y), fr)), fr));
                        dx2[3] = add2(pk(b1.z, b1.w), mul2(add2(pk(c1.z, c1.w), mul2(pk(d1.z, d1.w), fr)), fr));
                    } else {
                        const float4 b0 = *reinterpret_cast<const float4*>(rows + r * 32), b1 = *reinterpret_cast<const float4*>(rows + r * 32 + 16);
                        dx2[0] = pk(b0.x, b0.y); dx2[1] = pk(b0.z, b0.w); dx2[2] = pk(b1.x, b1.y); dx2[3] = pk(b1.z, b1.w);
                    }
                    const float post = sign * inv_scale;          // +-1 in MODE 0; a power of two (exact) in MODE 1
                    const f2 p2 = pk(post, post);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dx2[q] = mul2(dx2[q], p2);
                }
                frac0 = frac1;
                if (st + 2 < total) frac1 = a.stage_frac[st + 2];

                const bool tr = TRACE && a.trace && blockIdx.x == 0 && t == 0 && r == 0 && st < 64;
                if (tr) a.trace[st * 8 + 2] = clock64();
                mbar_wait(&d_ready[t], phase);
                phase ^= 1;
                tc_fence_after();
                if (tr) a.trace[st * 8 + 3] = clock64();

                // ---- kv[h] = sum_c D[h*C + c] * dX[c]  (the bias is already in D); packed FFMA2, next TMEM
                //      load in flight while the current 16 columns are consumed
                float kv[kH];
                if (a.debug & 4) {
#pragma unroll
                    for (int h = 0; h < kH; ++h) kv[h] = y[h] * 1e-3f;
                } else {
                    // 32-column loads, the next one in flight while the current one is consumed (see umma.cuh)
                    uint32_t va[32], vb[32];
                    tmem_ld32_issue(taddr, va);
#pragma unroll
                    for (int j = 0; j < kN / 32; ++j) {
                        uint32_t* cur = (j & 1) ? vb : va;
                        tmem_ld32_wait(cur);
                        if (j + 1 < kN / 32) tmem_ld32_issue(taddr + (uint32_t)(32 * (j + 1)), (j & 1) ? va : vb);
#pragma unroll
                        for (int hh = 0; hh < 4; ++hh) {
                            f2 acc = mul2(pk(__uint_as_float(cur[8 * hh + 0]), __uint_as_float(cur[8 * hh + 1])), dx2[0]);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 2]), __uint_as_float(cur[8 * hh + 3])), dx2[1], acc);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 4]), __uint_as_float(cur[8 * hh + 5])), dx2[2], acc);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 6]), __uint_as_float(cur[8 * hh + 7])), dx2[3], acc);
                            float lo, hi;
                            upk(acc, lo, hi);
                            kv[4 * j + hh] = lo + hi;
                        }
                    }
                }

                if (tr) a.trace[st * 8 + 4] = clock64();
                // ---- Runge-Kutta combination in registers, one rounding per operation in the order of
                //      oracle/odeint_port.py, two hidden units per instruction; zn ends up in kv's registers
                bool step_done = false;
                const f2 dt2 = pk(dt, dt);
                if (a.method == TCDE_RK4_38) {
                    const f2 th2 = pk(third, third);
                    if (sub == 0) {
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            k1[h] = kv[h]; k1[h + 1] = kv[h + 1];
                            upk(add2(pk(y[h], y[h + 1]), mul2(mul2(dt2, pk(kv[h], kv[h + 1])), th2)), kv[h], kv[h + 1]);
                        }
                    } else if (sub == 1) {
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            s23[h] = kv[h]; s23[h + 1] = kv[h + 1];
                            upk(add2(pk(y[h], y[h + 1]), mul2(dt2, sub2(pk(kv[h], kv[h + 1]), mul2(pk(k1[h], k1[h + 1]), th2)))),
                                kv[h], kv[h + 1]);
                        }
                    } else if (sub == 2) {
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            const f2 k2p = pk(s23[h], s23[h + 1]), k3p = pk(kv[h], kv[h + 1]);
                            upk(add2(pk(y[h], y[h + 1]), mul2(dt2, add2(sub2(pk(k1[h], k1[h + 1]), k2p), k3p))), kv[h], kv[h + 1]);
                            upk(add2(k2p, k3p), s23[h], s23[h + 1]);
                        }
                    } else {
                        const f2 three = pk(3.f, 3.f), eighth = pk(0.125f, 0.125f);
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            const f2 sum = add2(add2(pk(k1[h], k1[h + 1]), mul2(three, pk(s23[h], s23[h + 1]))), pk(kv[h], kv[h + 1]));
                            upk(add2(pk(y[h], y[h + 1]), mul2(mul2(sum, dt2), eighth)), kv[h], kv[h + 1]);
                        }
                        step_done = true;
                    }
                } else if (a.method == TCDE_MIDPOINT) {
                    if (sub == 0) {
                        const float half = E::mul(0.5f, dt);
#pragma unroll
                        for (int h = 0; h < kH; ++h) kv[h] = E::add(y[h], E::mul(kv[h], half));
                    } else {
#pragma unroll
                        for (int h = 0; h < kH; ++h) kv[h] = E::add(y[h], E::mul(dt, kv[h]));
                        step_done = true;
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < kH; ++h) kv[h] = E::add(y[h], E::mul(dt, kv[h]));
                    step_done = true;
                }
                if (tr) a.trace[st * 8 + 5] = clock64();
                if (more) {                               // hand the next stage to the tensor core first ...
                    write_a(kv, st + 1);
                    if (issuer_warp) issue_stage(st + 1);
                }
                if (tr) a.trace[st * 8 + 6] = clock64();
                if (step_done) {                          // ... then the bookkeeping that nobody waits for
                    while (next_out == step) {
                        const int mode = a.out_mode[jn];
                        if (mode == 0) write_out(jn, y);
                        else if (mode == 1) write_out(jn, kv);
                        else {
                            const float slope_w = a.out_slope[jn];
                            float v[kH];
#pragma unroll
                            for (int h = 0; h < kH; ++h) v[h] = E::add(y[h], E::mul(slope_w, E::sub(kv[h], y[h])));
                            write_out(jn, v);
                        }
                        ++jn;
                        next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
                    }
#pragma unroll
                    for (int h = 0; h < kH; ++h) y[h] = kv[h];
                    ++step;
                    sub = 0;
                    dt = dt_next;
                    if (step + 1 < a.n_steps) dt_next = a.step_dt[step + 1];
                } else {
                    ++sub;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

template <int MODE, bool ISSUER> static int launch_tc(const UmmaArgs& a, cudaStream_t stream) {
    const int64_t per_cta = tc::kTile * tc::kTiles;
    const int64_t ctas = (a.n_paths + per_cta - 1) / per_cta;
    TCDE_CHECK_SUPPORTED(ctas < (1ll << 31), "too many paths");
    auto kern = a.stage_dump ? tc::cdeint_tc_kernel<MODE, ISSUER, false, true>
                             : a.trace ? tc::cdeint_tc_kernel<MODE, ISSUER, true, false> : tc::cdeint_tc_kernel<MODE, ISSUER, false, false>;
    constexpr int smem = tc::Smem<MODE>::total + 1024;      // slack for the 1024-byte alignment of the tiles
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    alignas(64) CUtensorMap rows_map;
    const int row_floats = (a.control_kind == TCDE_CONTROL_CUBIC) ? 4 * tc::kC : tc::kC;
    const int rc = tc::make_rows_tensor_map(&rows_map, a.control, a.n_paths, a.n_rows, row_floats);
    TCDE_CHECK_SUPPORTED(rc == 0, "tensor-core solve: cuTensorMapEncodeTiled failed (%d) for control [%lld][%lld x %d floats]", rc,
                         (long long)a.n_paths, (long long)a.n_rows, row_floats);
    kern<<<(unsigned)ctas, tc::kRowThreads + (ISSUER ? tc::kIssuerThreads : 0), smem, stream>>>(a, rows_map);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

int solve_tc_f32(const UmmaArgs& a, int H, int C, int mode, cudaStream_t stream) {
    TCDE_CHECK_SUPPORTED(H == tc::kH && C == tc::kC, "tensor-core solve: built for hidden=32, channels=8 (got %d, %d)", H, C);
    TCDE_CHECK_SUPPORTED((reinterpret_cast<uintptr_t>(a.control) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.z0) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
                         "tensor-core solve: control, z0 and out must be 16-byte aligned");
    TCDE_CHECK_SUPPORTED(a.stage_dump == nullptr || (reinterpret_cast<uintptr_t>(a.stage_dump) & 15) == 0,
                         "tensor-core solve: the stage dump must be 16-byte aligned");
    // mode: bit 0 = operand split (0 = 3xTF32, 1 = 2xFP16), bit 1 = dedicated issuer warpgroup
    switch (mode & 3) {
        case 0: return launch_tc<0, false>(a, stream);
        case 1: return launch_tc<1, false>(a, stream);
        case 2: return launch_tc<0, true>(a, stream);
        default: return launch_tc<1, true>(a, stream);
    }
}

}  // namespace tcde
