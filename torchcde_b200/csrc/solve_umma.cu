// Hot path (ii) on the 5th-generation tensor cores: the fused fixed-step CDE solve with the
// vector-field GEMM on tcgen05.mma (kind::tf32, 3xTF32 split) and accumulators in TMEM.
//
// Why tensor cores here.  At (L=256, C=8, H=32) the solve is ~700 flop/byte:
// the CUDA-core kernel (solve_simt.cu) is bound by the FP32 FMA pipe at ~1/100 of the HBM
// roofline.  Per Runge-Kutta stage the work is one [paths x 32] . [32 x 256] product followed by
// a cheap contraction over channels -- GEMM-shaped, so it belongs on the tensor pipe.  fp32
// parity (rtol 1e-4 after 1020 chaotic stages) rules out plain TF32 (10-bit mantissa); the
// operands are therefore split  x = hi + lo  (hi = x rounded to TF32, lo = x - hi, exact) and
//   z.W^T  ~=  z_lo.W_hi + z_hi.W_lo + z_hi.W_hi        (error ~2^-21 relative, fp32-class)
// is accumulated in fp32 in TMEM by twelve 128x256x8 MMAs per stage.
//
// CTA anatomy (288 threads, one CTA per SM, all 512 TMEM columns):
//   * two independent tiles of 128 paths; tile T owns TMEM columns [256T, 256T+256);
//   * warps 0-3 / 4-7: the 128 "row" threads of tile 0 / 1 -- thread r owns path r of its tile:
//     its hidden state y[32] lives in registers for all steps; per stage it forms its own dX/dt (8
//     values from the spline row it prefetched with cp.async one stage ahead), reads its 256
//     accumulators from TMEM (tcgen05.ld, next load in flight), contracts them with dX/dt (packed
//     FFMA2), does the Runge-Kutta combination in the reference's operation order (packed
//     FADD2/FMUL2, parked slopes in shared memory), splits the next stage input into hi / lo and
//     writes both into the K-major 128B-swizzled A tiles in shared memory;
//   * warp 8: allocates TMEM, then only issues: wait "A ready" -> 13 x tcgen05.mma (12 for the
//     3xTF32 split, 1 that adds the bias through a K-augmentation tile) -> tcgen05.commit ->
//     "D ready".  While tile 0's MMAs run, tile 1's rows do their epilogue and vice versa.
//   W^T (hi and lo, 64 KB) stays resident in shared memory for the whole solve.
// Measured chain of one tile-stage (scripts/trace_umma.py, profiles/r01_umma_trace.txt):
// MMAs 1331 cycles -> visible to the rows +525 -> contraction 669 -> Runge-Kutta 890 -> split, store,
// fence, arrive 551 -> issuer wakes +97; the tensor pipe is busy 72 % of the time.
#include "umma.cuh"

namespace tcde {

namespace umma {

constexpr int kH = 32;            // hidden channels == K of the MMA == one 128-byte swizzled row
constexpr int kTile = 128;        // paths per tile == UMMA M
constexpr int kTiles = 2;         // tiles per CTA
constexpr int kThreads = kTile * kTiles + 32;

// shared memory map (bytes)
template <int N> struct Smem {
    static constexpr int b_hi = 0;
    static constexpr int b_lo = b_hi + N * 128;
    static constexpr int a_hi = b_lo + N * 128;                         // [kTiles][128 rows][128 B]
    static constexpr int a_lo = a_hi + kTiles * kTile * 128;
    static constexpr int park = a_lo + kTiles * kTile * 128;            // [kTiles][2][kH][kTile] floats
    static constexpr int raw = park + kTiles * 2 * kH * kTile * 4;      // [kTiles][6][kTile] float4
    // bias as a 13th MMA (K augmentation), K-major without swizzle: B_aug[N][8] = (bias_hi, bias_lo, 0...);
    // A_aug = (1, 1, 0...) for every row, so ONE 8-row core-matrix pair serves all 128 rows (SBO = 0)
    static constexpr int b_aug = raw + kTiles * 6 * kTile * 16;         // N rows x 32 B
    static constexpr int a_aug = b_aug + N * 32;                        // 8 rows x 32 B
    static constexpr int bars = a_aug + 256;                            // a_ready[2], d_ready[2], tmem slot
    static constexpr int total = bars + 64;
};

template <int C, bool TRACE, bool DUMP>
__global__ void __launch_bounds__(kThreads, 1) cdeint_umma_kernel(const UmmaArgs a) {
    constexpr int N = kH * C;
    static_assert(N % 16 == 0 && N <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
    static_assert(C == 8, "row prefetch and contraction below are written for 8 channels");
    using S = Smem<N>;
    using E = exact<float>;
    extern __shared__ unsigned char smem_unaligned[];
    // operand tiles of the 128-byte swizzle must start on 1024-byte boundaries of the shared window
    unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    float* b_hi = reinterpret_cast<float*>(smem + S::b_hi);
    float* b_lo = reinterpret_cast<float*>(smem + S::b_lo);
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(smem + S::bars);
    uint64_t* d_ready = a_ready + kTiles;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + kTiles);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int64_t cta_path0 = (int64_t)blockIdx.x * (kTile * kTiles);
    const int total = a.n_steps * a.n_stages;

    // ---- one-time setup --------------------------------------------------------------------
    for (int e = tid; e < N * kH; e += kThreads) {
        const int n = e >> 5, k = e & 31;
        const float w = a.weight[e];                     // weight[n][k], n = h*C + c
        const float hi = tf32_hi(w);
        b_hi[swz(n, k)] = hi;
        b_lo[swz(n, k)] = w - hi;
    }
    {
        float* a_aug = reinterpret_cast<float*>(smem + S::a_aug);
        float* b_aug = reinterpret_cast<float*>(smem + S::b_aug);
        for (int e = tid; e < 8 * 8; e += kThreads) {
            const int row = e >> 3, k = e & 7;
            a_aug[(k >> 2) * 32 + row * 4 + (k & 3)] = (k < 2) ? 1.f : 0.f;
        }
        for (int e = tid; e < N * 8; e += kThreads) {
            const int row = e >> 3, k = e & 7;
            const float bv = a.bias[row];
            const float hi = tf32_hi(bv);
            b_aug[(row >> 3) * 64 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3)] = (k == 0) ? hi : (k == 1) ? (bv - hi) : 0.f;
        }
    }
    if (tid == 0) {
        for (int t = 0; t < kTiles; ++t) {
            mbar_init(&a_ready[t], kTile);
            mbar_init(&d_ready[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == kTiles * 4) tmem_alloc(tmem_slot, 512);
    fence_proxy_async_smem();                             // operand tiles were written by the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    bool tile_live[kTiles];
#pragma unroll
    for (int t = 0; t < kTiles; ++t) tile_live[t] = (cta_path0 + (int64_t)t * kTile) < a.n_paths;

    if (warp == kTiles * 4) {
        // ================================ MMA issuer ==========================================
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
        const uint64_t dbh = make_desc(b_hi), dbl = make_desc(b_lo);
        // no-swizzle K-major descriptors: LBO = 128 B (next 4 k), SBO = 256 B (next 8 rows; 0 for A_aug:
        // all rows are the same (1, 1, 0, ...)), version 1, layout type 0
        const uint64_t db_aug = (uint64_t)((smem_u32(smem + S::b_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
        const uint64_t da_aug = (uint64_t)((smem_u32(smem + S::a_aug) & 0x3FFFF) >> 4) | (8ull << 16) | (0ull << 32) | (1ull << 46);
        uint32_t phase[kTiles] = {0, 0};
        for (int st = 0; st < total; ++st) {
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
                if (!tile_live[t]) continue;
                mbar_wait(&a_ready[t], phase[t]);
                phase[t] ^= 1;
                tc_fence_after();
                if ((tid & 31) == 0) {
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 0] = clock64();
                    const uint64_t dah = make_desc(smem + S::a_hi + t * kTile * 128);
                    const uint64_t dal = make_desc(smem + S::a_lo + t * kTile * 128);
                    const uint32_t d = tmem_base + (uint32_t)(t * N);
                    // small terms first; each k-block is 8 tf32 = 32 bytes = +2 in the descriptor's address field
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dal + 2 * kb, dbh + 2 * kb, idesc, kb > 0);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, dbl + 2 * kb, idesc, 1);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) mma_tf32(d, dah + 2 * kb, dbh + 2 * kb, idesc, 1);
                    mma_tf32(d, da_aug, db_aug, idesc, 1);                // + bias (1 * bias_hi + 1 * bias_lo)
                    mma_commit(&d_ready[t]);
                    if (TRACE && a.trace && blockIdx.x == 0 && t == 0 && st < 64) a.trace[st * 8 + 1] = clock64();
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ row threads =========================================
        const int t = warp >> 2;                          // tile of this thread
        const int r = tid & (kTile - 1);                  // row (path) within the tile
        const int64_t path = cta_path0 + (int64_t)t * kTile + r;
        const bool live = path < a.n_paths;
        const int64_t lpath = live ? path : a.n_paths - 1;
        if (tile_live[t]) {
            float* a_hi = reinterpret_cast<float*>(smem + S::a_hi + t * kTile * 128);
            float* a_lo = reinterpret_cast<float*>(smem + S::a_lo + t * kTile * 128);
            float* park = reinterpret_cast<float*>(smem + S::park) + (size_t)t * 2 * kH * kTile + r;
            float4* raw = reinterpret_cast<float4*>(smem + S::raw) + (size_t)t * 6 * kTile + r;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * N);
            const bool cubic = (a.control_kind == TCDE_CONTROL_CUBIC);
            const int row_stride = cubic ? 4 * C : C;
            const float* crow = a.control + lpath * a.n_rows * row_stride + (cubic ? C : 0);
            const bool negate = a.sign < 0.f;

            auto fetch_row = [&](int idx) {               // (b | 2c | 3d) of interval idx -> raw[0..5]
                const float* src = crow + (int64_t)idx * row_stride;
                const int parts = cubic ? 6 : 2;
                for (int j = 0; j < parts; ++j) cp_async16(&raw[j * kTile], src + 4 * j);
                cp_async_commit();
            };
            auto write_a = [&](const float* z, int stage_no) {   // next stage input -> swizzled hi / lo rows
                if (DUMP && live) {                       // ... and, for the adjoint, to the trajectory in HBM
                    float4* dst = reinterpret_cast<float4*>(a.stage_dump + ((int64_t)stage_no * a.n_paths + path) * kH);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_float4(z[4 * c4], z[4 * c4 + 1], z[4 * c4 + 2], z[4 * c4 + 3]);
                }
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
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&a_ready[t]);
            };
            auto write_out = [&](int j, const float* v) {
                if (!live) return;
                float4* dst = reinterpret_cast<float4*>(a.out + (path * a.n_out + j) * kH);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            };

            float y[kH];
            {
                const float4* zp = reinterpret_cast<const float4*>(a.z0 + lpath * kH);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 v = zp[c4];
                    y[4 * c4] = v.x; y[4 * c4 + 1] = v.y; y[4 * c4 + 2] = v.z; y[4 * c4 + 3] = v.w;
                }
            }
            int jn = 0;
            int next_out = (a.n_out > 0) ? a.out_step[0] : 0x7fffffff;
            while (jn < a.n_out && next_out < 0) {
                write_out(jn, y);
                ++jn;
                next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
            }
            fetch_row(a.stage_index[0]);
            write_a(y, 0);

            const float third = (float)(1.0 / 3.0);
            int step = 0, sub = 0;
            float dt = a.step_dt[0];
            float dt_next = (a.n_steps > 1) ? a.step_dt[1] : 0.f;
            float frac0 = a.stage_frac[0];                        // this stage's fraction
            int idx1 = (total > 1) ? a.stage_index[1] : 0;        // next stage's schedule entry
            float frac1 = (total > 1) ? a.stage_frac[1] : 0.f;
            uint32_t phase = 0;
            for (int st = 0; st < total; ++st) {
                const bool more = st + 1 < total;
                // ---- in the MMA's shadow: dX/dt from the prefetched row (interpolation_cubic.py:331-336) ----
                cp_async_wait<0>();
                f2 dx2[C / 2];
                {
                    const float4 b0 = raw[0], b1 = raw[kTile];
                    if (cubic) {
                        const float4 c0 = raw[2 * kTile], c1 = raw[3 * kTile], d0 = raw[4 * kTile], d1 = raw[5 * kTile];
                        const f2 fr = pk(frac0, frac0);
                        dx2[0] = add2(pk(b0.x, b0.y), mul2(add2(pk(c0.x, c0.y), mul2(pk(d0.x, d0.y), fr)), fr));
                        dx2[1] = add2(pk(b0.z, b0.w), mul2(add2(pk(c0.z, c0.w), mul2(pk(d0.z, d0.w), fr)), fr));
                        dx2[2] = add2(pk(b1.x, b1.y), mul2(add2(pk(c1.x, c1.y), mul2(pk(d1.x, d1.y), fr)), fr));
                        dx2[3] = add2(pk(b1.z, b1.w), mul2(add2(pk(c1.z, c1.w), mul2(pk(d1.z, d1.w), fr)), fr));
                    } else {
                        dx2[0] = pk(b0.x, b0.y); dx2[1] = pk(b0.z, b0.w); dx2[2] = pk(b1.x, b1.y); dx2[3] = pk(b1.z, b1.w);
                    }
                }
                if (more) fetch_row(idx1);
                frac0 = frac1;
                if (st + 2 < total) {                     // schedule entries are read a full stage ahead
                    idx1 = a.stage_index[st + 2];
                    frac1 = a.stage_frac[st + 2];
                }

                const bool tr = TRACE && a.trace && blockIdx.x == 0 && t == 0 && r == 0 && st < 64;
                if (tr) a.trace[st * 8 + 2] = clock64();
                mbar_wait(&d_ready[t], phase);
                phase ^= 1;
                tc_fence_after();
                if (tr) a.trace[st * 8 + 3] = clock64();

                // ---- kv[h] = sum_c D[h*C + c] * dX[c]  (the bias is already in D); packed FFMA2, next TMEM
                //      load in flight while the current 32 columns are consumed
                float kv[kH];
                {
                    uint32_t va[16], vb[16];
                    tmem_ld16_issue(taddr, va);
#pragma unroll
                    for (int j = 0; j < N / 16; ++j) {
                        uint32_t* cur = (j & 1) ? vb : va;
                        tmem_ld16_wait(cur);
                        if (j + 1 < N / 16) tmem_ld16_issue(taddr + (uint32_t)(16 * (j + 1)), (j & 1) ? va : vb);
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            f2 acc = mul2(pk(__uint_as_float(cur[8 * hh + 0]), __uint_as_float(cur[8 * hh + 1])), dx2[0]);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 2]), __uint_as_float(cur[8 * hh + 3])), dx2[1], acc);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 4]), __uint_as_float(cur[8 * hh + 5])), dx2[2], acc);
                            acc = fma2(pk(__uint_as_float(cur[8 * hh + 6]), __uint_as_float(cur[8 * hh + 7])), dx2[3], acc);
                            float lo, hi;
                            upk(acc, lo, hi);
                            const float sum = lo + hi;
                            kv[2 * j + hh] = negate ? -sum : sum;
                        }
                    }
                }

                if (tr) a.trace[st * 8 + 4] = clock64();
                // ---- Runge-Kutta combination, one rounding per operation (oracle/odeint_port.py), two hidden
                //      units per instruction; parked slopes are all loaded first, then combined, then stored
                bool step_done = false;
                float zn[kH];
                const f2 dt2 = pk(dt, dt);
                if (a.method == TCDE_RK4_38) {
                    const f2 th2 = pk(third, third);
                    if (sub == 0) {
#pragma unroll
                        for (int h = 0; h < kH; h += 2)
                            upk(add2(pk(y[h], y[h + 1]), mul2(mul2(dt2, pk(kv[h], kv[h + 1])), th2)), zn[h], zn[h + 1]);
#pragma unroll
                        for (int h = 0; h < kH; ++h) park[(size_t)h * kTile] = kv[h];
                    } else if (sub == 1) {
                        float k1[kH];
#pragma unroll
                        for (int h = 0; h < kH; ++h) k1[h] = park[(size_t)h * kTile];
#pragma unroll
                        for (int h = 0; h < kH; h += 2)
                            upk(add2(pk(y[h], y[h + 1]), mul2(dt2, sub2(pk(kv[h], kv[h + 1]), mul2(pk(k1[h], k1[h + 1]), th2)))),
                                zn[h], zn[h + 1]);
#pragma unroll
                        for (int h = 0; h < kH; ++h) park[(size_t)(kH + h) * kTile] = kv[h];
                    } else if (sub == 2) {
                        float k1[kH], k2[kH];
#pragma unroll
                        for (int h = 0; h < kH; ++h) {
                            k1[h] = park[(size_t)h * kTile];
                            k2[h] = park[(size_t)(kH + h) * kTile];
                        }
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            const f2 k2p = pk(k2[h], k2[h + 1]), k3p = pk(kv[h], kv[h + 1]);
                            upk(add2(pk(y[h], y[h + 1]), mul2(dt2, add2(sub2(pk(k1[h], k1[h + 1]), k2p), k3p))), zn[h], zn[h + 1]);
                            upk(add2(k2p, k3p), k2[h], k2[h + 1]);
                        }
#pragma unroll
                        for (int h = 0; h < kH; ++h) park[(size_t)(kH + h) * kTile] = k2[h];
                    } else {
                        float k1[kH], s23[kH];
#pragma unroll
                        for (int h = 0; h < kH; ++h) {
                            k1[h] = park[(size_t)h * kTile];
                            s23[h] = park[(size_t)(kH + h) * kTile];
                        }
                        const f2 three = pk(3.f, 3.f), eighth = pk(0.125f, 0.125f);
#pragma unroll
                        for (int h = 0; h < kH; h += 2) {
                            const f2 sum = add2(add2(pk(k1[h], k1[h + 1]), mul2(three, pk(s23[h], s23[h + 1]))), pk(kv[h], kv[h + 1]));
                            upk(add2(pk(y[h], y[h + 1]), mul2(mul2(sum, dt2), eighth)), zn[h], zn[h + 1]);
                        }
                        step_done = true;
                    }
                } else if (a.method == TCDE_MIDPOINT) {
                    if (sub == 0) {
                        const float half = E::mul(0.5f, dt);
#pragma unroll
                        for (int h = 0; h < kH; ++h) zn[h] = E::add(y[h], E::mul(kv[h], half));
                    } else {
#pragma unroll
                        for (int h = 0; h < kH; ++h) zn[h] = E::add(y[h], E::mul(dt, kv[h]));
                        step_done = true;
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < kH; ++h) zn[h] = E::add(y[h], E::mul(dt, kv[h]));
                    step_done = true;
                }
                if (tr) a.trace[st * 8 + 5] = clock64();
                if (more) write_a(zn, st + 1);            // hand the next stage to the tensor core first ...
                if (tr) a.trace[st * 8 + 6] = clock64();
                if (step_done) {                          // ... then the bookkeeping that nobody waits for
                    while (next_out == step) {
                        const int mode = a.out_mode[jn];
                        if (mode == 0) write_out(jn, y);
                        else if (mode == 1) write_out(jn, zn);
                        else {
                            const float slope_w = a.out_slope[jn];
                            float v[kH];
#pragma unroll
                            for (int h = 0; h < kH; ++h) v[h] = E::add(y[h], E::mul(slope_w, E::sub(zn[h], y[h])));
                            write_out(jn, v);
                        }
                        ++jn;
                        next_out = (jn < a.n_out) ? a.out_step[jn] : 0x7fffffff;
                    }
#pragma unroll
                    for (int h = 0; h < kH; ++h) y[h] = zn[h];
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
    if (warp == kTiles * 4) tmem_dealloc(tmem_base, 512);
}


}  // namespace umma

bool solve_umma_supported(int H, int C) { return H == umma::kH && C == 8; }

int solve_umma_f32(const UmmaArgs& a, int H, int C, cudaStream_t stream) {
    TCDE_CHECK_SUPPORTED(solve_umma_supported(H, C), "tensor-core solve: built for hidden=32, channels=8 (got %d, %d)", H, C);
    TCDE_CHECK_SUPPORTED((reinterpret_cast<uintptr_t>(a.control) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.z0) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
                         "tensor-core solve: control, z0 and out must be 16-byte aligned");
    const int64_t per_cta = umma::kTile * umma::kTiles;
    const int64_t ctas = (a.n_paths + per_cta - 1) / per_cta;
    TCDE_CHECK_SUPPORTED(ctas < (1ll << 31), "too many paths");
    auto kern = a.stage_dump ? umma::cdeint_umma_kernel<8, false, true>
                             : a.trace ? umma::cdeint_umma_kernel<8, true, false> : umma::cdeint_umma_kernel<8, false, false>;
    TCDE_CHECK_SUPPORTED(a.stage_dump == nullptr || (reinterpret_cast<uintptr_t>(a.stage_dump) & 15) == 0,
                         "tensor-core solve: the stage dump must be 16-byte aligned");
    constexpr int smem = umma::Smem<256>::total + 1024;     // slack for the 1024-byte alignment of the tiles
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)ctas, umma::kThreads, smem, stream>>>(a);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

}  // namespace tcde
