// Hot path (ii), round 2, second kernel: the fused fixed-step CDE solve on tcgen05 with TWO threads per path.
//
// solve_tc.cu (thread = path) is paced by a serial chain per tile and stage, measured with clock64 stamps
// (profiles/r02_trace_tc.txt): 7 MMAs issued -> accumulator ready 1,090 cycles | TMEM -> contraction with dX/dt 642 |
// Runge-Kutta combination 256 | operand split + store + fence + arrive 568 | wake-ups ~200 = 2,770 cycles, of which the
// tensor pipe works 2 tiles x 896.  TMEM's 512 columns hold two [128 x 256] accumulators, so no third tile can hide the
// 1,470 cycles of per-row CUDA-core work -- but that work splits cleanly down the middle of the hidden state:
//   thread (row r, half f) owns hidden units 16f .. 16f+15 of its path: accumulator columns 128f .. 128f+127 of lane r
//   (contraction), those units' Runge-Kutta state, and their 64 bytes of the path's operand row (2 hi + 2 lo chunks).
// The two halves of a row live in warps w and w+4 of a tile's eight row warps (same TMEM lane quarter); the only thing
// they exchange is the row's maximum for the power-of-two scale: one float through shared memory and one 64-thread named
// barrier per stage.  512 row threads leave 112 registers each, so the slopes k1 and k2 (+k3) are parked in shared memory
// ([pair of hidden units][thread] float2: conflict-free 64-bit accesses, 16 per stage on average) -- there is room:
// 64 KB next to the 144 KB of operands.
//
// CTA anatomy (640 threads): warps 0-7 / 8-15 = tile 0 / 1 (warps 0-3 half 0, 4-7 half 1 of tile 0, ...); warps 16 / 17 =
// MMA + TMA issuer of tile 0 / 1; warps 18, 19 only donate registers (setmaxnreg: 96 -> 24 / 112).  Everything else --
// the 2xFP16 operand split with per-path scaling, the bias block, TMA-fed spline rows, persistent CTAs over (tile pair,
// time segment) units with the state handed over through HBM -- is solve_tc.cu's MODE 1, whose comments apply.
#include "tc_common.cuh"

namespace tcde {

namespace tc2 {

using namespace umma;
using namespace tc;

constexpr int kH = 32;
constexpr int kC = 8;
constexpr int kN = kH * kC;       // 256 accumulator columns per tile
constexpr int kTile = 128;
constexpr int kTiles = 2;
constexpr int kHalf = kH / 2;     // hidden units per thread
constexpr int kRowThreads = 2 * kTile * kTiles;   // 512: thread = (path, half of the hidden state)
constexpr int kPairPaths = kTile * kTiles;        // 256 paths per CTA-unit
constexpr int kThreads = kRowThreads + 128;       // + the issuer warpgroup

struct Units {
    int n_seg;
    int steps_per_seg;
    const int32_t* seg_first_out;
    float* ystate;
    int* progress;
};

// shared-memory map (bytes)
struct Smem {
    static constexpr int b = 0;                                   // weights: 256 rows x [hi(32) | lo(32)] halves
    static constexpr int a = b + kN * 128;                        // [kTiles][128 rows x 128 B]
    static constexpr int raw = a + kTiles * kTile * 128;          // [kTiles][2 buffers][128 rows x 128 B] (TMA, swizzled)
    static constexpr int raw_buf = kTile * 128;
    static constexpr int b_aug = raw + kTiles * 2 * raw_buf;      // 256 rows x 32 B
    static constexpr int a_aug = b_aug + kN * 32;                 // [kTiles][128 rows x 32 B]
    static constexpr int red = a_aug + kTiles * kTile * 32;       // reduction scratch of the setup
    static constexpr int rowmax = red + 256;                      // [2 parities][kTiles][2 halves][128] floats
    static constexpr int bars = rowmax + 2 * kTiles * 2 * kTile * 4;
    static constexpr int slopes = bars + 128;                     // k1: [8][512] float2, then s23: [8][512] float2
    static constexpr int total = slopes + 2 * (kHalf / 2) * kRowThreads * 8;
};
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory");

__device__ __forceinline__ void reg_dealloc_24() { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n" ::: "memory"); }
__device__ __forceinline__ void reg_alloc_112() { asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// kv[h] = sum_c D[h * 8 + c] * dx[c] for this thread's 128 accumulator columns (16 hidden units), 16 columns per TMEM load,
// the next load in flight while the current one is consumed; same summation order as tc::contract_row
__device__ __forceinline__ void contract_half(uint32_t taddr, const f2* dx2, float* kv) {
    uint32_t va[16], vb[16];
    tmem_ld16_issue(taddr, va);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t* cur = (j & 1) ? vb : va;
        tmem_ld16_wait(cur);
        if (j + 1 < 8) tmem_ld16_issue(taddr + (uint32_t)(16 * (j + 1)), (j & 1) ? va : vb);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            f2 acc = mul2(pk(__uint_as_float(cur[8 * hh + 0]), __uint_as_float(cur[8 * hh + 1])), dx2[0]);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 2]), __uint_as_float(cur[8 * hh + 3])), dx2[1], acc);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 4]), __uint_as_float(cur[8 * hh + 5])), dx2[2], acc);
            acc = fma2(pk(__uint_as_float(cur[8 * hh + 6]), __uint_as_float(cur[8 * hh + 7])), dx2[3], acc);
            float lo, hi;
            upk(acc, lo, hi);
            kv[2 * j + hh] = lo + hi;
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1) cdeint_tc2_kernel(const UmmaArgs a, const Units un, const __grid_constant__ CUtensorMap rows_map) {
    using S = Smem;
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
    constexpr int kAllocWarp = kRowThreads / 32;              // warp 16

    // ---- one-time setup: operand B (the weights, split) and the bias K-block -------------------------------------
    float inv_w_scale = 1.f, beta = 0.f, w_scale;
    prepare_b_fp16(smem + S::b, smem + S::b_aug, a.weight, a.bias, red, tid, kThreads, w_scale, inv_w_scale, beta);
    for (int e = tid; e < kTiles * kTile * 2; e += kThreads)  // per-row scale goes to k = 0, 1 each stage
        *reinterpret_cast<uint4*>(smem + S::a_aug + (e >> 8) * (kTile * 32) + aug_off((e >> 1) & 127, e & 1)) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int t = 0; t < kTiles; ++t) {
            mbar_init(&a_ready[t], 2 * kTile);
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

    const bool is_row = warp < kRowThreads / 32;
    const int n_pairs = (int)((a.n_paths + kPairPaths - 1) / kPairPaths);
    const int n_units = n_pairs * un.n_seg;
    const int t = is_row ? (warp >> 3) : (warp - kRowThreads / 32);      // tile served by this warp (issuers: 0, 1; donors: 2, 3)
    const int hf = (warp >> 2) & 1;                           // row threads: which half of the hidden state
    const int r = tid & (kTile - 1);
    uint32_t phase_a = 0, phase_d = 0;
    uint32_t kcount = 0;

#define TCDE_UNIT_PROLOGUE()                                                                                          \
        const int seg = u / n_pairs, pair = u - seg * n_pairs;                                                        \
        const int step_lo = seg * un.steps_per_seg;                                                                   \
        const int step_hi = min(a.n_steps, step_lo + un.steps_per_seg);                                               \
        const int st_lo = step_lo * a.n_stages, st_hi = step_hi * a.n_stages;                                         \
        const int64_t tile_path0 = (int64_t)pair * kPairPaths + (int64_t)t * kTile;                                   \
        const bool tile_live = t < kTiles && tile_path0 < a.n_paths;                                                  \
        if (seg > 0) {                      /* the unit that ends where this one starts must be done */              \
            if (tid == 0)                                                                                             \
                while (ld_acquire(un.progress + pair) < seg) __nanosleep(64);                                         \
            __syncthreads();                                                                                          \
        }
#define TCDE_UNIT_EPILOGUE()                                                                                          \
        if (seg + 1 < un.n_seg) {                                                                                     \
            __syncthreads();                                                                                          \
            if (tid == 0) st_release(un.progress + pair, seg + 1);                                                    \
        }

    if (!is_row) {
        reg_dealloc_24();
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            TCDE_UNIT_PROLOGUE()
            // ================================ MMA + TMA issuer of tile t ==========================
            if (tile_live && (tid & 31) == 0) {
                auto fetch_rows = [&](uint32_t k, int interval) {     // rows of one interval -> raw[t][k & 1]
                    uint64_t* bar = &raw_full[2 * t + (k & 1)];
                    mbar_expect_tx(bar, row_bytes_tile);
                    tma_load_2d(smem + S::raw + (2 * t + (k & 1)) * S::raw_buf, &rows_map, interval * row_floats, (int)tile_path0, bar);
                };
                int idx_next = (st_lo + 1 < st_hi) ? a.stage_index[st_lo + 1] : 0;
                fetch_rows(kcount, a.stage_index[st_lo]);
                for (int st = st_lo; st < st_hi; ++st) {
                    const int idx_fetch = idx_next;
                    if (st + 2 < st_hi) idx_next = a.stage_index[st + 2];
                    mbar_wait(&a_ready[t], phase_a);
                    phase_a ^= 1;
                    tc_fence_after();
                    const bool tr = TRACE && a.trace && u == 0 && t == 0 && st < 64;
                    if (tr) a.trace[st * 8 + 0] = clock64();
                    issue_fp16(tmem_base + (uint32_t)(t * kN), smem + S::a + t * kTile * 128, smem + S::a_aug + t * kTile * 32,
                               smem + S::b, smem + S::b_aug, &d_ready[t]);
                    if (tr) a.trace[st * 8 + 1] = clock64();
                    ++kcount;
                    if (st + 1 < st_hi) fetch_rows(kcount, idx_fetch);
                }
            }
            __syncwarp();
            TCDE_UNIT_EPILOGUE()
        }
    } else {
        reg_alloc_112();
        f2* k1s = reinterpret_cast<f2*>(smem + S::slopes) + tid;                      // [q][512]: pair q of this thread's units
        f2* s23s = k1s + (kHalf / 2) * kRowThreads;
        float* rowmax = reinterpret_cast<float*>(smem + S::rowmax);
        const int barrier_id = 1 + t * 4 + (warp & 3);
        uint32_t mx_parity = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            TCDE_UNIT_PROLOGUE()
            if (tile_live) {
            // ================================ row threads =========================================
            const int64_t path = tile_path0 + r;
            const bool live = path < a.n_paths;
            const int64_t lpath = live ? path : a.n_paths - 1;
            unsigned char* a_row = smem + S::a + t * kTile * 128 + r * 128;
            unsigned char* a_aug = smem + S::a_aug + t * kTile * 32;
            const unsigned char* raw_tile = smem + S::raw + 2 * t * S::raw_buf;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * kN + hf * (kN / 2));
            const float sign = (a.sign < 0.f) ? -1.f : 1.f;
            float inv_scale = 1.f;

            auto write_a = [&](const float* z) {              // next stage input -> this thread's 64 bytes of the operand row
                float mx[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) mx[k] = fmaxf(fmaxf(fabsf(z[k]), fabsf(z[k + 4])), fmaxf(fabsf(z[k + 8]), fabsf(z[k + 12])));
                const float m_own = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                float* slot = rowmax + mx_parity * (kTiles * 2 * kTile) + t * (2 * kTile);
                slot[hf * kTile + r] = m_own;
                pair_barrier(barrier_id);                     // the two warps that share these 32 rows
                const float m = fmaxf(fmaxf(m_own, slot[(hf ^ 1) * kTile + r]), beta);
                mx_parity ^= 1;
                const int e = min(max(exponent_of(m), 30), 224);
                const float s = pow2_biased(127 + 13 - (e - 127));
                const f2 s2 = pk(s, s);
                uint32_t hi_h[8], lo_h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const f2 sc = mul2(pk(z[2 * j], z[2 * j + 1]), s2);
                    float s0, s1, h0, h1, l0, l1;
                    upk(sc, s0, s1);
                    hi_h[j] = pack_h2(s0, s1);
                    unpack_h2(hi_h[j], h0, h1);
                    upk(sub2(sc, pk(h0, h1)), l0, l1);
                    lo_h[j] = pack_h2(l0, l1);
                }
                const int x = r & 7;
#pragma unroll
                for (int c = 0; c < 2; ++c) {                 // chunks 2f, 2f+1 (hi) and 4+2f, 4+2f+1 (lo) of the 128-byte row
                    *reinterpret_cast<uint4*>(a_row + (((2 * hf + c) ^ x) << 4)) = make_uint4(hi_h[4 * c], hi_h[4 * c + 1], hi_h[4 * c + 2], hi_h[4 * c + 3]);
                    *reinterpret_cast<uint4*>(a_row + (((4 + 2 * hf + c) ^ x) << 4)) = make_uint4(lo_h[4 * c], lo_h[4 * c + 1], lo_h[4 * c + 2], lo_h[4 * c + 3]);
                }
                if (hf == 0) {
                    const float sb = s * beta;
                    *reinterpret_cast<uint32_t*>(a_aug + aug_off(r, 0)) = pack_h2(sb, sb);
                }
                inv_scale = pow2_biased(127 - 13 + (e - 127)) * inv_w_scale;
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&a_ready[t]);
            };
            auto write_out = [&](int j, const float* v) {
                if (!live) return;
                float4* dst = reinterpret_cast<float4*>(a.out + (path * a.n_out + j) * kH + hf * kHalf);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            };

            float y[kHalf];
            {
                const float4* zp = reinterpret_cast<const float4*>((seg == 0 ? a.z0 : un.ystate) + lpath * kH + hf * kHalf);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 v = (seg == 0) ? zp[c4] : __ldcg(zp + c4);
                    y[4 * c4] = v.x; y[4 * c4 + 1] = v.y; y[4 * c4 + 2] = v.z; y[4 * c4 + 3] = v.w;
                }
            }
            int jn = (seg == 0) ? 0 : un.seg_first_out[seg];
            int next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
            while (seg == 0 && jn < a.n_out && next_out < 0) {         // outputs at the initial time
                write_out(jn, y);
                ++jn;
                next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
            }
            if (t == 1 && !(a.debug & 64)) __nanosleep(600);  // anti-phase of the two tiles (see solve_tc.cu)
            write_a(y);

            const float third = (float)(1.0 / 3.0);
            int step = step_lo, sub = 0;
            float dt = a.step_dt[step_lo];
            float dt_next = (step_lo + 1 < a.n_steps) ? a.step_dt[step_lo + 1] : 0.f;
            float frac0 = a.stage_frac[st_lo];
            float frac1 = (st_lo + 1 < st_hi) ? a.stage_frac[st_lo + 1] : 0.f;
            for (int st = st_lo; st < st_hi; ++st) {
                const bool more = st + 1 < st_hi;
                // ---- in the MMA's shadow: dX/dt from the rows the TMA unit delivered, times sign / operand scales
                mbar_wait(&raw_full[2 * t + (kcount & 1)], (kcount >> 1) & 1);
                f2 dx2[kC / 2];
                {
                    const unsigned char* rows = raw_tile + (kcount & 1) * S::raw_buf;
                    if (cubic) {
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
                    const float post = sign * inv_scale;
                    const f2 p2 = pk(post, post);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dx2[q] = mul2(dx2[q], p2);
                }
                ++kcount;
                frac0 = frac1;
                if (st + 2 < st_hi) frac1 = a.stage_frac[st + 2];

                const bool tr = TRACE && a.trace && u == 0 && t == 0 && r == 0 && hf == 0 && st < 64;
                if (tr) a.trace[st * 8 + 2] = clock64();
                mbar_wait(&d_ready[t], phase_d);
                phase_d ^= 1;
                tc_fence_after();
                if (tr) a.trace[st * 8 + 3] = clock64();

                float kv[kHalf];
                contract_half(taddr, dx2, kv);
                if (tr) a.trace[st * 8 + 4] = clock64();

                // ---- Runge-Kutta combination, one rounding per operation in the order of oracle/odeint_port.py; the
                //      slopes of earlier stages come from / go to this thread's column of the shared-memory park
                bool step_done = false;
                const f2 dt2 = pk(dt, dt);
                if (a.method == TCDE_RK4_38) {
                    const f2 th2 = pk(third, third);
                    if (sub == 0) {
#pragma unroll
                        for (int q = 0; q < kHalf / 2; ++q) {
                            const f2 kq = pk(kv[2 * q], kv[2 * q + 1]);
                            k1s[q * kRowThreads] = kq;
                            upk(add2(pk(y[2 * q], y[2 * q + 1]), mul2(mul2(dt2, kq), th2)), kv[2 * q], kv[2 * q + 1]);
                        }
                    } else if (sub == 1) {
#pragma unroll
                        for (int q = 0; q < kHalf / 2; ++q) {
                            const f2 kq = pk(kv[2 * q], kv[2 * q + 1]);
                            s23s[q * kRowThreads] = kq;
                            upk(add2(pk(y[2 * q], y[2 * q + 1]), mul2(dt2, sub2(kq, mul2(k1s[q * kRowThreads], th2)))), kv[2 * q], kv[2 * q + 1]);
                        }
                    } else if (sub == 2) {
#pragma unroll
                        for (int q = 0; q < kHalf / 2; ++q) {
                            const f2 k2p = s23s[q * kRowThreads], k3p = pk(kv[2 * q], kv[2 * q + 1]);
                            upk(add2(pk(y[2 * q], y[2 * q + 1]), mul2(dt2, add2(sub2(k1s[q * kRowThreads], k2p), k3p))), kv[2 * q], kv[2 * q + 1]);
                            s23s[q * kRowThreads] = add2(k2p, k3p);
                        }
                    } else {
                        const f2 three = pk(3.f, 3.f), eighth = pk(0.125f, 0.125f);
#pragma unroll
                        for (int q = 0; q < kHalf / 2; ++q) {
                            const f2 sum = add2(add2(k1s[q * kRowThreads], mul2(three, s23s[q * kRowThreads])), pk(kv[2 * q], kv[2 * q + 1]));
                            upk(add2(pk(y[2 * q], y[2 * q + 1]), mul2(mul2(sum, dt2), eighth)), kv[2 * q], kv[2 * q + 1]);
                        }
                        step_done = true;
                    }
                } else if (a.method == TCDE_MIDPOINT) {
                    if (sub == 0) {
                        const float half = E::mul(0.5f, dt);
#pragma unroll
                        for (int h = 0; h < kHalf; ++h) kv[h] = E::add(y[h], E::mul(kv[h], half));
                    } else {
#pragma unroll
                        for (int h = 0; h < kHalf; ++h) kv[h] = E::add(y[h], E::mul(dt, kv[h]));
                        step_done = true;
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < kHalf; ++h) kv[h] = E::add(y[h], E::mul(dt, kv[h]));
                    step_done = true;
                }
                if (tr) a.trace[st * 8 + 5] = clock64();
                if (more) write_a(kv);                        // hand the next stage to the tensor core first ...
                if (tr) a.trace[st * 8 + 6] = clock64();
                if (step_done) {                              // ... then the bookkeeping that nobody waits for
                    while (next_out == step) {
                        const int mode = a.out_mode[jn];
                        if (mode == 0) write_out(jn, y);
                        else if (mode == 1) write_out(jn, kv);
                        else {
                            const float slope_w = a.out_slope[jn];
                            float v[kHalf];
#pragma unroll
                            for (int h = 0; h < kHalf; ++h) v[h] = E::add(y[h], E::mul(slope_w, E::sub(kv[h], y[h])));
                            write_out(jn, v);
                        }
                        ++jn;
                        next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
                    }
#pragma unroll
                    for (int h = 0; h < kHalf; ++h) y[h] = kv[h];
                    ++step;
                    sub = 0;
                    dt = dt_next;
                    if (step + 1 < a.n_steps) dt_next = a.step_dt[step + 1];
                } else {
                    ++sub;
                }
            }
            if (seg + 1 < un.n_seg && live) {                 // hand the state over to whoever runs the next segment
                float4* dst = reinterpret_cast<float4*>(un.ystate + path * kH + hf * kHalf);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(y[4 * c4], y[4 * c4 + 1], y[4 * c4 + 2], y[4 * c4 + 3]);
                __threadfence();
            }
            }
            TCDE_UNIT_EPILOGUE()
        }
    }
#undef TCDE_UNIT_PROLOGUE
#undef TCDE_UNIT_EPILOGUE
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc2

static int choose_segments2(int64_t n_pairs, int n_steps, int sms) {
    int best = 1;
    double best_cost = (double)((n_pairs + sms - 1) / sms);
    for (int s = 2; s <= 8; s *= 2) {
        if (n_steps / s < 8) break;
        const int64_t units = n_pairs * s;
        const double cost = (double)((units + sms - 1) / sms) / s;
        if (cost < best_cost * 0.97) { best_cost = cost; best = s; }
    }
    return best;
}

// The two-threads-per-path kernel: fp32, hidden 32, channels 8, 2xFP16 split, no stage dump (the adjoint's dumping solves
// stay on solve_tc.cu).  Same arguments and results as solve_tc_f32(mode 1).
int solve_tc2_f32(const UmmaArgs& a, int H, int C, cudaStream_t stream) {
    TCDE_CHECK_SUPPORTED(H == tc2::kH && C == tc2::kC, "tensor-core solve: built for hidden=32, channels=8 (got %d, %d)", H, C);
    TCDE_CHECK_SUPPORTED(a.stage_dump == nullptr, "tensor-core solve (two threads per path): no stage dump");
    TCDE_CHECK_SUPPORTED((reinterpret_cast<uintptr_t>(a.control) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.z0) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
                         "tensor-core solve: control, z0 and out must be 16-byte aligned");
    const int64_t n_pairs = (a.n_paths + tc2::kPairPaths - 1) / tc2::kPairPaths;
    TCDE_CHECK_SUPPORTED(n_pairs < (1ll << 28), "too many paths");
    const int sms = sm_count();
    const int n_seg = (a.debug & 16) ? 1 : (a.debug & 32) ? (a.n_steps >= 4 ? 4 : a.n_steps) : choose_segments2(n_pairs, a.n_steps, sms);
    const int steps_per_seg = (a.n_steps + n_seg - 1) / n_seg;
    const int64_t n_units = n_pairs * n_seg;
    const int grid = (int)(n_units < sms ? n_units : sms);
    auto kern = a.trace ? tc2::cdeint_tc2_kernel<true> : tc2::cdeint_tc2_kernel<false>;
    constexpr int smem = tc2::Smem::total + 1024;
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    alignas(64) CUtensorMap rows_map;
    const int row_floats = (a.control_kind == TCDE_CONTROL_CUBIC) ? 4 * tc2::kC : tc2::kC;
    const int rc = tc::make_rows_tensor_map(&rows_map, a.control, a.n_paths, a.n_rows, row_floats);
    TCDE_CHECK_SUPPORTED(rc == 0, "tensor-core solve: cuTensorMapEncodeTiled failed (%d) for control [%lld][%lld x %d floats]", rc,
                         (long long)a.n_paths, (long long)a.n_rows, row_floats);
    tc2::Units un{n_seg, steps_per_seg, nullptr, nullptr, nullptr};
    void* workspace = nullptr;
    if (n_seg > 1) {
        const size_t y_bytes = (size_t)a.n_paths * tc2::kH * sizeof(float);
        const size_t flag_bytes = ((size_t)n_pairs * sizeof(int) + 255) & ~(size_t)255;
        const size_t cur_bytes = ((size_t)n_seg * sizeof(int32_t) + 255) & ~(size_t)255;
        TCDE_CHECK_CUDA(cudaMallocAsync(&workspace, y_bytes + flag_bytes + cur_bytes, stream));
        un.ystate = static_cast<float*>(workspace);
        un.progress = reinterpret_cast<int*>(static_cast<char*>(workspace) + y_bytes);
        int32_t* cursor = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + y_bytes + flag_bytes);
        un.seg_first_out = cursor;
        TCDE_CHECK_CUDA(cudaMemsetAsync(un.progress, 0, flag_bytes, stream));
        tc::segment_cursor_kernel<<<1, 32, 0, stream>>>(a.out_step, a.n_out, n_seg, steps_per_seg, cursor);
    }
    kern<<<grid, tc2::kThreads, smem, stream>>>(a, un, rows_map);
    const cudaError_t launch_err = cudaGetLastError();
    if (workspace) cudaFreeAsync(workspace, stream);
    TCDE_CHECK_CUDA(launch_err);
    return TCDE_OK;
}

}  // namespace tcde
