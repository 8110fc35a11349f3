// Hot path (i): batched spline-coefficient construction (sm_100a).
//
// All kernels here are HBM-bound streaming kernels (~1 flop per byte).  The
// design rules that matter are the memory ones: coalesced loads, results staged in shared
// memory and written back as one contiguous 1-D bulk (TMA) store per tile, persistent CTAs
// sized from the SM count.  Arithmetic uses tcde::exact<> (one rounding per operation, no
// FMA contraction) wherever the header promises bit-identical results.
#include "common.cuh"

namespace tcde {

static constexpr int kThreads = 256;

// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2): two IEEE-rounded operations per instruction
typedef uint64_t f2;
__device__ __forceinline__ f2 pk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// =========================================================================================
// Hermite cubic with backward differences  (interpolation_hermite_cubic_bdiff.py:5-44)
// =========================================================================================
// Work item = (path, tile of TR consecutive intervals).  The CTA stages the TR+2 knot rows it
// needs in shared memory, every thread produces the four coefficients of its (interval,
// channel) elements into a shared output tile laid out exactly like global memory
// ([row][a|b|2c|3d][channel]), and one thread issues a single bulk store of the tile (the
// tile is a contiguous byte range of the output).  Two output tiles alternate so that the
// store of tile n overlaps the computation of tile n+1.
template <typename T, bool UNIT>
__global__ void __launch_bounds__(kThreads)
hermite_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out, int64_t n_paths, int L, int C,
               int TR, int tiles_per_path, int use_bulk, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<T>;
    const int row_elems = 4 * C;
    T* ot0 = reinterpret_cast<T*>(smem_raw);
    T* ot1 = ot0 + (size_t)TR * row_elems;
    T* xs = ot1 + (size_t)TR * row_elems;
    T* ts = xs + (size_t)(TR + 2) * C;

    const int tid = threadIdx.x;
    const int di = kThreads / C, dc = kThreads - di * C;   // (row, channel) advance per thread-stride
    const int i_first = tid / C, c_first = tid - i_first * C;
    const int64_t n_items = n_paths * tiles_per_path;
    bool saw_nan = false;
    int buf = 0;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
        const int64_t p = item / tiles_per_path;
        const int tile = (int)(item - p * tiles_per_path);
        const int r0 = tile * TR;
        const int nr = min(TR, L - 1 - r0);
        if (use_bulk && tid == 0) bulk_wait_read<1>();   // the store that last read ot[buf] has drained

        // knot rows r0-1 .. r0+nr ; row -1 does not exist for the first tile
        const T* xp = x + (p * L + r0 - 1) * C;
        const int ne = (nr + 2) * C;
        for (int e = (r0 == 0 ? C : 0) + tid; e < ne; e += kThreads) xs[e] = xp[e];
        if (!UNIT) {
            for (int e = (r0 == 0 ? 1 : 0) + tid; e < nr + 2; e += kThreads) ts[e] = t[r0 - 1 + e];
        }
        __syncthreads();

        T* ot = buf ? ot1 : ot0;
        int i = i_first, c = c_first;
        for (int e = tid; e < nr * C; e += kThreads) {
            const T xl = xs[(i + 1) * C + c];
            const T xh = xs[(i + 2) * C + c];
            saw_nan |= is_nan(xl) | is_nan(xh);
            const bool first = (r0 + i == 0);
            T b, two_c, three_d;
            if (UNIT) {
                // dt == 1 exactly: every division by dt and the 1/dt^2 factor are exact identities
                const T dn = E::sub(xh, xl);
                const T dp = first ? dn : E::sub(xl, xs[i * C + c]);
                const T bend = E::sub(dn, dp);
                const T inner = E::add(E::sub(E::mul(T(3), bend), dn), dp);
                two_c = E::mul(T(2), inner);
                three_d = E::sub(bend, two_c);
                b = dp;
            } else {
                const T dt = E::sub(ts[i + 2], ts[i + 1]);
                const T dn = E::div(E::sub(xh, xl), dt);                       // bdiff.py:39
                const T dp = first ? dn : E::div(E::sub(xl, xs[i * C + c]), E::sub(ts[i + 1], ts[i]));
                const T inner = E::add(E::sub(E::mul(T(3), E::sub(dn, dp)), dn), dp);
                two_c = E::div(E::mul(T(2), inner), dt);                        // bdiff.py:17
                const T inv_sq = E::div(T(1), E::mul(dt, dt));
                three_d = E::sub(E::mul(inv_sq, E::sub(dn, dp)), E::div(two_c, dt));   // bdiff.py:18
                b = dp;
            }
            // four stores per thread; rotating which coefficient goes first by row spreads a
            // warp's stores over all 32 banks (rows are 4C words apart)
            T* row = ot + (size_t)i * row_elems + c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int w = (k + i) & 3;
                const T v = (w == 0) ? xl : (w == 1) ? b : (w == 2) ? two_c : three_d;
                row[w * C] = v;
            }
            i += di;
            c += dc;
            if (c >= C) { c -= C; ++i; }
        }
        T* gp = out + (p * (int64_t)(L - 1) + r0) * row_elems;
        if (use_bulk) {
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                bulk_store(gp, ot, (uint32_t)((size_t)nr * row_elems * sizeof(T)));
                bulk_commit();
            }
        } else {
            __syncthreads();
            for (int e = tid; e < nr * row_elems; e += kThreads) gp[e] = ot[e];
        }
    }
    if (use_bulk && tid == 0) bulk_wait_read<0>();
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// Vectorised fp32 variant for channels % 4 == 0 (the BASELINE shapes): one thread produces a whole
// (interval, 4-channel) block -- three 128-bit read-only loads of the neighbouring knot rows
// straight from global memory (each row is reused by three intervals and hits L1), 4-wide
// arithmetic, four 128-bit shared stores -- so the instruction count per output byte is ~4x
// lower than the scalar kernel's (which ncu showed to be issue-bound at ~40% of HBM peak), and
// a tile needs one CTA barrier instead of two.
template <bool UNIT>
__global__ void __launch_bounds__(kThreads)
hermite_vec4_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ out,
                    int64_t n_paths, int L, int C, int TR, int tiles_per_path, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<float>;
    const int Q = C >> 2;                                   // 16-byte chunks per coefficient part
    const int row_chunks = 4 * Q;
    float4* ot0 = reinterpret_cast<float4*>(smem_raw);
    float4* ot1 = ot0 + (size_t)TR * row_chunks;
    const int tid = threadIdx.x;
    const int di = kThreads / Q, dq = kThreads - di * Q;
    const int i_first = tid / Q, q_first = tid - i_first * Q;
    const int64_t n_items = n_paths * tiles_per_path;
    bool saw_nan = false;
    int buf = 0;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
        const int64_t p = item / tiles_per_path;
        const int tile = (int)(item - p * tiles_per_path);
        const int r0 = tile * TR;
        const int nr = min(TR, L - 1 - r0);
        const float4* xp = reinterpret_cast<const float4*>(x + p * (int64_t)L * C);
        float4* ot = buf ? ot1 : ot0;
        if (tid == 0) bulk_wait_read<1>();
        __syncthreads();                                    // ot[buf] is free again (and visible to all)
        int i = i_first, q = q_first;
        for (int e = tid; e < nr * Q; e += kThreads) {
            const int r = r0 + i;
            const float4 lo = __ldg(xp + (size_t)r * Q + q);
            const float4 hi = __ldg(xp + (size_t)(r + 1) * Q + q);
            const float4 pp = (r > 0) ? __ldg(xp + (size_t)(r - 1) * Q + q) : lo;
            saw_nan |= is_nan(lo.x) | is_nan(lo.y) | is_nan(lo.z) | is_nan(lo.w) | is_nan(hi.x) | is_nan(hi.y) |
                       is_nan(hi.z) | is_nan(hi.w);
            const float xl[4] = {lo.x, lo.y, lo.z, lo.w}, xh[4] = {hi.x, hi.y, hi.z, hi.w};
            const float xq[4] = {pp.x, pp.y, pp.z, pp.w};
            float b[4], c2[4], d3[4];
            float dt = 1.f, dtp = 1.f, inv_sq = 1.f;
            if (!UNIT) {
                dt = E::sub(t[r + 1], t[r]);
                dtp = (r > 0) ? E::sub(t[r], t[r - 1]) : dt;
                inv_sq = E::div(1.f, E::mul(dt, dt));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (UNIT) {
                    const float dn = E::sub(xh[j], xl[j]);
                    const float dp = (r > 0) ? E::sub(xl[j], xq[j]) : dn;
                    const float bend = E::sub(dn, dp);
                    c2[j] = E::mul(2.f, E::add(E::sub(E::mul(3.f, bend), dn), dp));
                    d3[j] = E::sub(bend, c2[j]);
                    b[j] = dp;
                } else {
                    const float dn = E::div(E::sub(xh[j], xl[j]), dt);
                    const float dp = (r > 0) ? E::div(E::sub(xl[j], xq[j]), dtp) : dn;
                    const float inner = E::add(E::sub(E::mul(3.f, E::sub(dn, dp)), dn), dp);
                    c2[j] = E::div(E::mul(2.f, inner), dt);
                    d3[j] = E::sub(E::mul(inv_sq, E::sub(dn, dp)), E::div(c2[j], dt));
                    b[j] = dp;
                }
            }
            float4* row = ot + (size_t)i * row_chunks + q;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                   // rotate the part order by row: conflict-free 128-bit stores
                const int w = (k + i) & 3;
                float4 v;
                if (w == 0) v = lo;
                else if (w == 1) v = make_float4(b[0], b[1], b[2], b[3]);
                else if (w == 2) v = make_float4(c2[0], c2[1], c2[2], c2[3]);
                else v = make_float4(d3[0], d3[1], d3[2], d3[3]);
                row[w * Q] = v;
            }
            i += di;
            q += dq;
            if (q >= Q) { q -= Q; ++i; }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_store(out + (p * (int64_t)(L - 1) + r0) * 4 * C, ot, (uint32_t)((size_t)nr * row_chunks * 16));
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// =========================================================================================
// Per-series scans: linear gap fill, forward fill, rectilinear preparation
// =========================================================================================
// One thread per scalar series, walking the length dimension.  A warp covers 32/C paths x C
// channels, so every load/store instruction touches whole 32-byte sectors (C >= 8 floats) and
// the four sectors of a 128-byte line are consumed by four consecutive iterations (L1 hits).
// Loads do not depend on the scan state, so they are issued eight steps ahead.
template <typename T, bool UNIT>
__global__ void __launch_bounds__(kThreads)
linear_fill_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out, int64_t n_series, int L,
                   int C) {
    using E = exact<T>;
    const int64_t g = blockIdx.x * (int64_t)kThreads + threadIdx.x;
    if (g >= n_series) return;
    const int64_t p = g / C;
    const int c = (int)(g - p * C);
    const T* xs = x + p * L * C + c;
    T* os = out + p * L * C + c;

    auto time_of = [&](int i) -> T { return UNIT ? T(i) : t[i]; };
    // interpolation_linear.py:60-69: x[j] = lo + ((t_j - t_lo) / (t_hi - t_lo)) * (hi - lo)
    auto bridge = [&](int lo, T vlo, int hi, T vhi) {
        const T tl = time_of(lo);
        const T span = E::sub(time_of(hi), tl);
        const T rise = E::sub(vhi, vlo);
        for (int j = lo + 1; j < hi; ++j) {
            const T ratio = E::div(E::sub(time_of(j), tl), span);
            os[(int64_t)j * C] = E::add(vlo, E::mul(ratio, rise));
        }
    };

    int prev = -1;
    T vprev = T(0);
    for (int i0 = 0; i0 < L; i0 += 8) {
        T ahead[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ahead[k] = (i0 + k < L) ? xs[(int64_t)(i0 + k) * C] : T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = i0 + k;
            if (i < L && !is_nan(ahead[k])) {
                const T v = ahead[k];
                if (prev < 0) {
                    if (i > 0) {            // :31-32 the first entry takes the first observation
                        os[0] = v;
                        bridge(0, v, i, v);
                    }
                } else if (i - prev > 1) {
                    bridge(prev, vprev, i, v);
                }
                os[(int64_t)i * C] = v;
                prev = i;
                vprev = v;
            }
        }
    }
    if (prev < 0) {                         // :19-21 nothing observed: the zero path
        for (int i = 0; i < L; ++i) os[(int64_t)i * C] = T(0);
    } else if (prev < L - 1) {              // :33-34 the last entry takes the last observation
        os[(int64_t)(L - 1) * C] = vprev;
        bridge(prev, vprev, L - 1, vprev);
    }
}

// Warp-per-path variant of the gap fill for length <= 32 * kFillRounds.  ncu on the
// thread-per-series kernel above: 70% of issue slots busy at 16% of HBM peak -- the data-dependent
// bridge loops diverge within a warp.  Here a warp stages its path in shared memory
// ([channel][position], conflict-free) and lane l owns positions l, l+32, ...; "nearest
// observation before / after" comes from warp ballots and bit scans (no loops, no divergence),
// every lane then applies the reference's interpolation formula once per element, and the
// filled tile is copied out with 128-bit coalesced stores.
constexpr int kFillRounds = 8;
template <typename T, bool UNIT>
__global__ void __launch_bounds__(kThreads)
linear_fill_warp_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out, int64_t n_paths, int L,
                        int C, int Lp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<T>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* tile = reinterpret_cast<T*>(smem_raw) + (size_t)warp * C * Lp;
    const int rounds = (L + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / 32);
    auto time_of = [&](int i) -> T { return UNIT ? T(i) : t[i]; };

    for (int64_t p = (int64_t)blockIdx.x * (kThreads / 32) + warp; p < n_paths; p += warps_total) {
        const T* xg = x + p * (int64_t)L * C;
        T* og = out + p * (int64_t)L * C;
        __syncwarp();
        const int di = 32 / C, dc = 32 - di * C;
        const bool vec4 = (sizeof(T) == 4) && ((C & 3) == 0) && (32 % (C >> 2) == 0) &&
                          ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
        if (vec4) {                                         // 128-bit coalesced loads, transposed to [c][i]
            const int Q = C >> 2, qi = 32 / Q;              // lane -> (row offset lane / Q, quad lane % Q)
            const int q = lane % Q;
            const float4* xg4 = reinterpret_cast<const float4*>(xg);
            for (int i = lane / Q; i < L; i += qi) {
                const float4 v4 = xg4[(size_t)i * Q + q];
                T* dst = tile + (4 * q) * Lp + i;
                dst[0] = (T)v4.x; dst[Lp] = (T)v4.y; dst[2 * Lp] = (T)v4.z; dst[3 * Lp] = (T)v4.w;
            }
        } else {
            int i = lane / C, c = lane - (lane / C) * C;
            for (int e = lane; e < L * C; e += 32) {        // coalesced load, transposed to [c][i]
                tile[c * Lp + i] = xg[e];
                i += di;
                c += dc;
                if (c >= C) { c -= C; ++i; }
            }
        }
        __syncwarp();
        for (int c = 0; c < C; ++c) {
            T* row = tile + c * Lp;
            T v[kFillRounds];
            uint32_t m[kFillRounds];
#pragma unroll
            for (int k = 0; k < kFillRounds; ++k) {
                const int i = 32 * k + lane;
                v[k] = (k < rounds && i < L) ? row[i] : T(0);
                m[k] = __ballot_sync(0xffffffffu, k < rounds && i < L && !is_nan(v[k]));
            }
            // first / last observation of the series, and per round the nearest ones outside it
            int first = L, last = -1;
#pragma unroll
            for (int k = 0; k < kFillRounds; ++k) {
                if (m[k]) {
                    if (first == L) first = 32 * k + __ffs(m[k]) - 1;
                    last = 32 * k + 31 - __clz(m[k]);
                }
            }
            if (first == L) {                               // nothing observed: the zero path
#pragma unroll
                for (int k = 0; k < kFillRounds; ++k)
                    if (k < rounds && 32 * k + lane < L) row[32 * k + lane] = T(0);
                continue;
            }
            const T v_first = row[first], v_last = row[last];
            __syncwarp();
            // first observation in any LATER round, per round (one backward pass instead of a search per round)
            int later_first[kFillRounds];
            {
                int carry = L;
#pragma unroll
                for (int k = kFillRounds - 1; k >= 0; --k) {
                    later_first[k] = carry;
                    if (m[k]) carry = 32 * k + __ffs(m[k]) - 1;
                }
            }
            int carry_prev = -1;
            const uint32_t le_mask = 0xffffffffu >> (31 - lane);
#pragma unroll
            for (int k = 0; k < kFillRounds; ++k) {
                const int i = 32 * k + lane;
                const bool hole = (k < rounds) && (i < L) && is_nan(v[k]);
                if (__any_sync(0xffffffffu, hole)) {        // whole rounds without a gap are skipped (warp-uniform)
                    const uint32_t below = m[k] & le_mask;
                    const int prv = below ? 32 * k + 31 - __clz(below) : carry_prev;
                    const uint32_t above = m[k] >> lane;
                    const int nxt = above ? i + __ffs(above) - 1 : later_first[k];
                    if (hole) {
                        int lo_i, hi_i;
                        T lo_v, hi_v;
                        if (prv < 0) {                      // before the first observation
                            lo_i = 0; hi_i = first; lo_v = v_first; hi_v = v_first;
                        } else if (nxt >= L) {              // after the last observation
                            lo_i = last; hi_i = L - 1; lo_v = v_last; hi_v = v_last;
                        } else {
                            lo_i = prv; hi_i = nxt; lo_v = row[prv]; hi_v = row[nxt];
                        }
                        T filled;
                        if (i == lo_i) filled = lo_v;       // an imputed end point itself
                        else if (i == hi_i) filled = hi_v;
                        else {
                            const T tl = time_of(lo_i);
                            const T ratio = E::div(E::sub(time_of(i), tl), E::sub(time_of(hi_i), tl));
                            filled = E::add(lo_v, E::mul(ratio, E::sub(hi_v, lo_v)));
                        }
                        v[k] = filled;
                    }
                }
                if (m[k]) carry_prev = 32 * k + 31 - __clz(m[k]);
            }
            __syncwarp();                                   // every gather from row[] is done
#pragma unroll
            for (int k = 0; k < kFillRounds; ++k)
                if (k < rounds && 32 * k + lane < L) row[32 * k + lane] = v[k];
        }
        __syncwarp();
        if (vec4) {
            const int Q = C >> 2, qi = 32 / Q;
            const int q = lane % Q;
            float4* og4 = reinterpret_cast<float4*>(og);
            for (int i = lane / Q; i < L; i += qi) {
                const T* src = tile + (4 * q) * Lp + i;
                og4[(size_t)i * Q + q] = make_float4((float)src[0], (float)src[Lp], (float)src[2 * Lp], (float)src[3 * Lp]);
            }
        } else {
            int i = lane / C, c = lane - (lane / C) * C;
            for (int e = lane; e < L * C; e += 32) {
                og[e] = tile[c * Lp + i];
                i += di;
                c += dc;
                if (c >= C) { c -= C; ++i; }
            }
        }
    }
}

// Scan variant of the gap fill: lane = (channel, chunk of G consecutive positions) of one path, for
// channels <= 32.  ncu on the ballot kernel above: 115 warp-instructions per 32 elements, issue
// bound at 19% of HBM peak.  Here the warp's tile stays in the global [position][channel] order
// (128-bit loads and stores on both sides, chunks padded so that the 32 lanes hit 32 banks) and
// every lane walks its chunk backward once: each hole is overwritten with a NaN whose payload holds
// the distances to the next observation and to the next hole of the chunk, which threads the holes
// into a list.  A few shuffles hand every chunk the nearest observation of the chunks before and
// after it.  Then the lane hops along its list of holes only (30% of the positions in the
// benchmark) and replaces each by the reference's interpolation formula; the end points of a gap
// are fetched once per gap.
template <typename T> struct nan_code;
template <> struct nan_code<float> {
    static constexpr uint32_t quiet = 0x7FC00000u;
    __device__ static int get(float v) { return (int)(__float_as_uint(v) & 0x3FFFFFu); }
    __device__ static float make(int i) { return __uint_as_float(quiet | (uint32_t)i); }
};
template <> struct nan_code<double> {
    static constexpr unsigned long long quiet = 0x7FF8000000000000ull;
    __device__ static int get(double v) { return (int)((unsigned long long)__double_as_longlong(v) & 0x3FFFFFull); }
    __device__ static double make(int i) { return __longlong_as_double((long long)(quiet | (unsigned long long)i)); }
};
constexpr int kFillNone = 0x3FFFFF;          // no observation
constexpr int kFillDist = 11;                // payload = distance to next observation | distance to next hole << 11

template <typename T, bool UNIT>
__global__ void __launch_bounds__(kThreads)
linear_fill_scan_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out, int64_t n_paths, int L,
                        int C, int lgG, int padw, int tile_words, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<T>;
    T* ts = reinterpret_cast<T*>(smem_raw);                 // knot times (absent for unit knots)
    T* tiles = ts + (UNIT ? 0 : ((L + 3) & ~3));
    if (!UNIT) {
        for (int i = threadIdx.x; i < L; i += blockDim.x) ts[i] = t[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* tile = tiles + (size_t)warp * tile_words;
    const int G = 1 << lgG;
    const int nch = (L + G - 1) >> lgG;                     // chunks in use (<= 32 / C)
    const int c = lane % C, j = lane / C;
    const bool active = j < nch;
    const int g0 = j << lgG, g1 = min(g0 + G, L);
    const unsigned full = 0xffffffffu;
    auto word = [&](int i) { return i * C + (i >> lgG) * padw; };
    auto time_of = [&](int i) -> T { return UNIT ? T(i) : ts[i]; };
    const bool vec4 = (sizeof(T) == 4) && ((C & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int Q = C >> 2;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / 32);
    bool saw_nan = false;

    for (int64_t p = (int64_t)blockIdx.x * (kThreads / 32) + warp; p < n_paths; p += warps_total) {
        const T* xg = x + p * (int64_t)L * C;
        T* og = out + p * (int64_t)L * C;
        __syncwarp();                                       // the previous path has been copied out
        if (vec4) {
            const float4* xg4 = reinterpret_cast<const float4*>(xg);
            const int dq_i = 32 / Q, dq_q = 32 - dq_i * Q;
            int i = lane / Q, q = lane - (lane / Q) * Q;
#pragma unroll 8
            for (int e = lane; e < L * Q; e += 32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(tile) + word(i) + 4 * q) = __ldg(xg4 + e);
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
        } else {
            const int di = 32 / C, dc = 32 - di * C;
            int i = lane / C, cc = lane - (lane / C) * C;
#pragma unroll 4
            for (int e = lane; e < L * C; e += 32) {
                tile[word(i) + cc] = xg[e];
                i += di;
                cc += dc;
                if (cc >= C) { cc -= C; ++i; }
            }
        }
        __syncwarp();
        // backward: thread the holes (payload: distance to the next observation / next hole of the chunk, 0 = none)
        int first_idx = kFillNone, last_idx = -1, head = -1;
        T last_val = T(0);
        if (active) {
            T* ptr = tile + word(g1 - 1) + c;
            for (int i = g1 - 1; i >= g0; --i, ptr -= C) {
                const T w = *ptr;
                if (is_nan(w)) {
                    const int d_obs = first_idx == kFillNone ? 0 : first_idx - i;
                    const int d_hole = head < 0 ? 0 : head - i;
                    *ptr = nan_code<T>::make(d_obs | (d_hole << kFillDist));
                    head = i;
                } else {
                    if (last_idx < 0) { last_idx = i; last_val = w; }
                    first_idx = i;
                }
            }
        }
        saw_nan |= head >= 0;
        // nearest observation in the chunks after (position) and before (position, value) this one
        int after = kFillNone, before = -1;
        T before_val = T(0);
        for (int d = 1; d < nch; ++d) {
            const int fa = __shfl_down_sync(full, first_idx, C * d);
            const int la = __shfl_up_sync(full, last_idx, C * d);
            const T lv = __shfl_up_sync(full, last_val, C * d);
            if (after == kFillNone && j + d < nch) after = fa;
            if (before < 0 && j - d >= 0) { before = la; before_val = lv; }
        }
        const int series_first = __shfl_sync(full, first_idx != kFillNone ? first_idx : after, c);
        const int series_last = __shfl_sync(full, last_idx >= 0 ? last_idx : before, c + C * (nch - 1));
        if (active) {
            if (series_first == kFillNone) {                // nothing observed: the zero path (linear.py:19-21)
                T* ptr = tile + word(g0) + c;
                for (int i = g0; i < g1; ++i, ptr += C) *ptr = T(0);
            } else {
                const T v_first = tile[word(series_first) + c], v_last = tile[word(series_last) + c];
                // the ends of the series count as observations carrying the first / last value (linear.py:31-34)
                int prev_idx = before >= 0 ? before : 0;
                T prev_val = before >= 0 ? before_val : v_first;
                const int far_idx = after != kFillNone ? after : L - 1;
                const T far_val = after != kFillNone ? tile[word(after) + c] : v_last;
                T* base = tile + word(g0) + c;              // a chunk has no padding inside
                T lo_t = T(0), span = T(1), rise = T(0);
                for (int i = head, visited = -2; i >= 0;) {
                    T* ptr = base + (i - g0) * C;
                    const int code = nan_code<T>::get(*ptr);
                    const int d_obs = code & ((1 << kFillDist) - 1), d_hole = code >> kFillDist;
                    if (i != visited + 1) {                 // a new gap: fetch its end points
                        if (i > g0) {
                            prev_idx = i - 1;
                            prev_val = ptr[-C];
                        }
                        const int hi_i = d_obs ? i + d_obs : far_idx;
                        const T hi_v = d_obs ? ptr[d_obs * C] : far_val;
                        lo_t = time_of(prev_idx);
                        span = E::sub(time_of(hi_i), lo_t);
                        rise = E::sub(hi_v, prev_val);
                    }
                    // linear.py:60-69: x[j] = lo + ((t_j - t_lo) / (t_hi - t_lo)) * (hi - lo)
                    *ptr = E::add(prev_val, E::mul(E::div(E::sub(time_of(i), lo_t), span), rise));
                    visited = i;
                    i = d_hole ? i + d_hole : -1;
                }
                // an imputed end point is a copy of the observation, not an interpolation
                if (g0 == 0 && series_first > 0) tile[c] = v_first;
                if (g1 == L && series_last < L - 1) tile[word(L - 1) + c] = v_last;
            }
        }
        __syncwarp();
        if (vec4) {
            float4* og4 = reinterpret_cast<float4*>(og);
            const int dq_i = 32 / Q, dq_q = 32 - dq_i * Q;
            int i = lane / Q, q = lane - (lane / Q) * Q;
#pragma unroll 8
            for (int e = lane; e < L * Q; e += 32) {
                og4[e] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tile) + word(i) + 4 * q);
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
        } else {
            const int di = 32 / C, dc = 32 - di * C;
            int i = lane / C, cc = lane - (lane / C) * C;
#pragma unroll 4
            for (int e = lane; e < L * C; e += 32) {
                og[e] = tile[word(i) + cc];
                i += di;
                cc += dc;
                if (cc >= C) { cc -= C; ++i; }
            }
        }
    }
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// misc.forward_fill (misc.py:103-126) and _prepare_rectilinear_interpolation
// (interpolation_linear.py:87-128).  RECT = false: out has L rows; RECT = true: 2L-1 rows, row
// 2i = held[i], row 2i+1 = held[i] except the time channel which takes held[i+1].
template <typename T, bool RECT>
__global__ void __launch_bounds__(kThreads)
hold_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n_series, int L, int C, int time_index,
            int32_t* __restrict__ flags) {
    const int64_t g = blockIdx.x * (int64_t)kThreads + threadIdx.x;
    if (g >= n_series) return;
    const int64_t p = g / C;
    const int c = (int)(g - p * C);
    const T* xs = x + p * L * C + c;
    const int out_rows = RECT ? 2 * L - 1 : L;
    T* os = out + p * out_rows * C + c;
    const bool is_time = RECT && (c == time_index);
    int32_t seen = 0;
    T held = xs[0];
    if (is_nan(held)) seen |= TCDE_FLAG_NAN_SEEN | TCDE_FLAG_NAN_FIRST_ROW | (is_time ? TCDE_FLAG_NAN_TIME : 0);
    for (int i0 = 0; i0 < L; i0 += 8) {
        T ahead[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ahead[k] = (i0 + k < L) ? xs[(int64_t)(i0 + k) * C] : T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = i0 + k;
            if (i < L) {
                const T v = ahead[k];
                const T before = held;
                if (!is_nan(v)) held = v;
                else seen |= TCDE_FLAG_NAN_SEEN | (is_time ? TCDE_FLAG_NAN_TIME : 0);
                if (RECT) {
                    if (i > 0) os[(int64_t)(2 * i - 1) * C] = is_time ? held : before;
                    os[(int64_t)(2 * i) * C] = held;
                } else {
                    os[(int64_t)i * C] = held;
                }
            }
        }
    }
    if (seen && flags != nullptr) atomicOr(flags, seen);
}

// =========================================================================================
// Natural cubic spline, NaN-free knots  (interpolation_cubic.py:7-53, misc.py:13-67)
// =========================================================================================
// The knots are shared by the whole batch (misc.py:83), so the eliminated diagonal and the
// forward multipliers of the Thomas solve are batch independent: one tiny kernel forms them
// once (the reference recomputes them for every series).  workspace = [rdt | rdt2 | mult | rnd].
template <typename T>
__global__ void natural_prep_kernel(const T* __restrict__ t, T* __restrict__ ws, int L) {
    using E = exact<T>;
    T* rdt = ws;
    T* rdt2 = ws + L;
    T* mult = ws + 2 * L;
    T* rnd = ws + 3 * L;
    __shared__ int s_wf, s_wb;
    for (int i = threadIdx.x; i < L - 1; i += blockDim.x) {
        const T t0 = t ? t[i] : T(i), t1 = t ? t[i + 1] : T(i + 1);
        const T r = E::div(T(1), E::sub(t1, t0));
        rdt[i] = r;
        rdt2[i] = E::mul(r, r);
    }
    if (threadIdx.x == 0) { s_wf = 0; s_wb = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        rdt[L - 1] = T(0);
        rdt2[L - 1] = T(0);
        // diag[i] = 2 * (rdt[i] + rdt[i-1]) with the out-of-range terms absent (cubic.py:31-35)
        T nd = E::mul(T(2), rdt[0]);
        T r_prev = rdt[0];
        // mult[0] is unused by the sweeps (0); for two knots it carries t1 - t0 (cubic.py:18)
        mult[0] = (L > 2) ? T(0) : E::sub(t ? t[1] : T(1), t ? t[0] : T(0));
        rnd[0] = E::div(T(1), nd);
        for (int i = 1; i < L; ++i) {
            const T r_cur = (i < L - 1) ? rdt[i] : T(0);
            const T diag = E::mul(T(2), E::add(r_cur, r_prev));
            const T w = E::div(r_prev, nd);                      // misc.py:59
            nd = E::sub(diag, E::mul(w, r_prev));                // misc.py:60
            mult[i] = w;
            rnd[i] = E::div(T(1), nd);
            r_prev = r_cur;
        }
    }
    __syncthreads();
    // Window sizes for the parallel sweeps of natural_win_kernel: the forward recurrence
    // f[i] = rhs[i] - mult[i] f[i-1] forgets f[i-w] by the factor prod |mult|, the backward one
    // k[i] = (f[i] - rdt[i] k[i+1]) rnd[i] forgets k[i+w] by prod |rdt rnd|; both factors are
    // ~0.27 per knot for a diagonally dominant system.  A window is long enough when that
    // product is below eps/16 (or it reaches the end of the series, where the start is exact).
    if (L > 2) {
        const T tol = (sizeof(T) == 4) ? T(3.7e-9) : T(1.4e-17);
        int wf = 0, wb = 0;
        for (int i = 1 + threadIdx.x; i < L; i += blockDim.x) {
            T prod = T(1);
            int w = 0;
            for (int j = i; j >= 1 && prod > tol; --j) { prod *= fabs(mult[j]); ++w; }
            wf = max(wf, w);
        }
        for (int i = threadIdx.x; i < L - 1; i += blockDim.x) {
            T prod = T(1);
            int w = 0;
            for (int j = i; j < L - 1 && prod > tol; ++j) { prod *= fabs(rdt[j] * rnd[j]); ++w; }
            wb = max(wb, w);
        }
        atomicMax(&s_wf, wf);
        atomicMax(&s_wb, wb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ws[4 * L] = T(s_wf);
        ws[4 * L + 1] = T(s_wb);
    }
}

// Natural cubic spline, parallel version.  One path at a time per CTA; thread = (series, chunk of
// G consecutive knots).  Each thread runs the two Thomas recurrences over its chunk preceded by
// a warm-up window (sizes from natural_prep_kernel), so all 256 threads sweep concurrently
// instead of one thread per series walking all L knots: the sweeps stop being the
// latency-bound phase that ncu showed (issue slots 28% busy, DRAM 11%).  The result equals the
// sequential recurrence up to a truncated influence below eps/16 -- the same few-ulp class as
// replacing the division by a multiplication with 1/nd.
template <typename T>
__global__ void __launch_bounds__(kThreads)
natural_win_kernel(const T* __restrict__ x, const T* __restrict__ ws, T* __restrict__ out, int64_t n_paths, int L,
                   int C, int Lp, int G, int TR, int use_bulk, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<T>;
    const int row_elems = 4 * C;
    T* ot0 = reinterpret_cast<T*>(smem_raw);
    T* ot1 = ot0 + (size_t)TR * row_elems;
    T* rdt = ot1 + (size_t)TR * row_elems;
    T* rdt2 = rdt + L;
    T* mult = rdt2 + L;
    T* rnd = mult + L;
    T* xs = rnd + L + 8;                 // [C][Lp]
    T* fs = xs + (size_t)C * Lp;         // forward-sweep values
    T* ks = fs + (size_t)C * Lp;         // knot slopes

    const int tid = threadIdx.x;
    for (int e = tid; e < 4 * L + 2; e += kThreads) rdt[e] = ws[e];
    __syncthreads();
    const int wf = (int)rdt[4 * L], wb = (int)rdt[4 * L + 1];
    const int n_chunks = (L + G - 1) / G;
    const int n_items = C * n_chunks;
    const int di = kThreads / C, dc = kThreads - di * C;
    const bool vec4 = (sizeof(T) == 4) && ((C & 3) == 0) && L > 2 && use_bulk &&
                      (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    bool saw_nan = false;
    int buf = 0;

    for (int64_t p = blockIdx.x; p < n_paths; p += gridDim.x) {
        __syncthreads();                 // the previous path's coefficient phase is done with xs / ks
        if (vec4) {                      // 128-bit coalesced loads, transposed to [c][i]
            const float4* xg4 = reinterpret_cast<const float4*>(x + p * (int64_t)L * C);
            const int Q = C >> 2;
            int i = tid / Q, q = tid - (tid / Q) * Q;
            const int dq_i = kThreads / Q, dq_q = kThreads - dq_i * Q;
            for (int e = tid; e < L * Q; e += kThreads) {
                const float4 v4 = xg4[e];
                saw_nan |= is_nan(v4.x) | is_nan(v4.y) | is_nan(v4.z) | is_nan(v4.w);
                T* dst = xs + (4 * q) * Lp + i;
                dst[0] = (T)v4.x; dst[Lp] = (T)v4.y; dst[2 * Lp] = (T)v4.z; dst[3 * Lp] = (T)v4.w;
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
        } else {
            const T* xg = x + p * (int64_t)L * C;
            int i = tid / C, c = tid - (tid / C) * C;
            for (int e = tid; e < L * C; e += kThreads) {
                const T v = xg[e];
                saw_nan |= is_nan(v);
                xs[c * Lp + i] = v;
                i += di;
                c += dc;
                if (c >= C) { c -= C; ++i; }
            }
        }
        __syncthreads();
        if (L > 2) {
            const bool pairs = (sizeof(T) == 4) && ((C & 1) == 0);
            if (pairs) {
                // two series per thread, packed arithmetic: half the instructions of the scalar sweeps
                const int C2 = C >> 1;
                const int n_pair_items = C2 * n_chunks;
                for (int it = tid; it < n_pair_items; it += kThreads) {     // forward sweep (misc.py:58-61)
                    const int c = 2 * (it % C2), j = it / C2;
                    const int g0 = j * G, g1 = min(g0 + G, L);
                    const float* x0 = reinterpret_cast<const float*>(xs) + c * Lp;
                    float* f0 = reinterpret_cast<float*>(fs) + c * Lp;
                    int i = max(g0 - wf, 0);
                    f2 f = pk2(0.f, 0.f);
                    const f2 three = pk2(3.f, 3.f);
                    f2 x_lo = pk2(x0[i], x0[Lp + i]);
                    f2 sc_prev = pk2(0.f, 0.f);
                    if (i > 0) {
                        const float r2 = (float)rdt2[i - 1];
                        sc_prev = mul2(mul2(three, sub2(x_lo, pk2(x0[i - 1], x0[Lp + i - 1]))), pk2(r2, r2));
                    }
                    for (; i < g1; ++i) {
                        // scaled difference 3 (x[i+1] - x[i]) / dt_i^2; rdt2[L-1] = 0 closes the system (cubic.py:36-39)
                        const int in = min(i + 1, L - 1);
                        const f2 x_hi = pk2(x0[in], x0[Lp + in]);
                        const float r2 = (float)rdt2[i];
                        const f2 sc = mul2(mul2(three, sub2(x_hi, x_lo)), pk2(r2, r2));
                        x_lo = x_hi;
                        const float m = (float)mult[i];
                        f = sub2(add2(sc, sc_prev), mul2(pk2(m, m), f));
                        sc_prev = sc;
                        if (i >= g0) upk2(f, f0[i], f0[Lp + i]);
                    }
                }
            } else {
                for (int it = tid; it < n_items; it += kThreads) {
                    const int c = it % C, j = it / C;
                    const int g0 = j * G, g1 = min(g0 + G, L);
                    const T* xr = xs + c * Lp;
                    T* fr = fs + c * Lp;
                    int i = max(g0 - wf, 0);
                    T f = T(0);
                    T x_lo = xr[i];
                    T sc_prev = (i > 0) ? E::mul(E::mul(T(3), E::sub(x_lo, xr[i - 1])), rdt2[i - 1]) : T(0);
                    for (; i < g1; ++i) {
                        const T x_hi = xr[min(i + 1, L - 1)];
                        const T sc = E::mul(E::mul(T(3), E::sub(x_hi, x_lo)), rdt2[i]);      // rdt2[L-1] = 0
                        x_lo = x_hi;
                        f = E::sub(E::add(sc, sc_prev), E::mul(mult[i], f));     // cubic.py:36-39, misc.py:61
                        sc_prev = sc;
                        if (i >= g0) fr[i] = f;
                    }
                }
            }
            __syncthreads();
            if (pairs) {
                const int C2 = C >> 1;
                const int n_pair_items = C2 * n_chunks;
                for (int it = tid; it < n_pair_items; it += kThreads) {     // back substitution (misc.py:63-65)
                    const int c = 2 * (it % C2), j = it / C2;
                    const int g0 = j * G, g1 = min(g0 + G, L);
                    const float* f0 = reinterpret_cast<const float*>(fs) + c * Lp;
                    float* k0 = reinterpret_cast<float*>(ks) + c * Lp;
                    f2 k = pk2(0.f, 0.f);
                    for (int i = min(g1 - 1 + wb, L - 1); i >= g0; --i) {
                        const float rd = (float)rdt[i], rn = (float)rnd[i];
                        k = mul2(sub2(pk2(f0[i], f0[Lp + i]), mul2(pk2(rd, rd), k)), pk2(rn, rn));
                        if (i < g1) upk2(k, k0[i], k0[Lp + i]);
                    }
                }
            } else {
                for (int it = tid; it < n_items; it += kThreads) {
                    const int c = it % C, j = it / C;
                    const int g0 = j * G, g1 = min(g0 + G, L);
                    const T* fr = fs + c * Lp;
                    T* kr = ks + c * Lp;
                    T k = T(0);
                    for (int i = min(g1 - 1 + wb, L - 1); i >= g0; --i) {
                        k = E::mul(E::sub(fr[i], E::mul(rdt[i], k)), rnd[i]);    // misc.py:63-65 (rdt[L-1] = 0)
                        if (i < g1) kr[i] = k;
                    }
                }
            }
        }
        __syncthreads();
        if (vec4) {
            // coefficient rows straight from registers to global memory: thread = (interval, 4 channels), four
            // 128-bit stores that each fill whole 32-byte sectors (a warp completes its 128-byte lines across
            // the four instructions).  No staging tile: the shared memory saved doubles the resident CTAs.
            const int Q = C >> 2;
            float4* og4 = reinterpret_cast<float4*>(out + p * (int64_t)(L - 1) * row_elems);
            int i = tid / Q, q = tid - (tid / Q) * Q;
            const int dq_i = kThreads / Q, dq_q = kThreads - dq_i * Q;
            for (int e = tid; e < (L - 1) * Q; e += kThreads) {
                const float rd = (float)rdt[i], rd2 = (float)rdt2[i];
                const f2 rdp = pk2(rd, rd), rd2p = pk2(rd2, rd2);
                float av[4], bv[4], cv[4], dv[4];
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const float* xr0 = reinterpret_cast<const float*>(xs) + (4 * q + j) * Lp + i;
                    const float* kr0 = reinterpret_cast<const float*>(ks) + (4 * q + j) * Lp + i;
                    const f2 xl = pk2(xr0[0], xr0[Lp]), xh = pk2(xr0[1], xr0[Lp + 1]);
                    const f2 kl = pk2(kr0[0], kr0[Lp]), kh = pk2(kr0[1], kr0[Lp + 1]);
                    const f2 six = mul2(pk2(2.f, 2.f), mul2(pk2(3.f, 3.f), sub2(xh, xl)));
                    const f2 sr = mul2(six, rdp);
                    const f2 c2 = mul2(sub2(sub2(sr, mul2(pk2(4.f, 4.f), kl)), mul2(pk2(2.f, 2.f), kh)), rdp);   // :45-47
                    const f2 d3 = mul2(sub2(mul2(pk2(3.f, 3.f), add2(kl, kh)), sr), rd2p);                          // :48-50
                    upk2(xl, av[j], av[j + 1]);
                    upk2(kl, bv[j], bv[j + 1]);
                    upk2(c2, cv[j], cv[j + 1]);
                    upk2(d3, dv[j], dv[j + 1]);
                }
                float4* row = og4 + (size_t)i * (4 * Q) + q;
                row[0] = make_float4(av[0], av[1], av[2], av[3]);
                row[Q] = make_float4(bv[0], bv[1], bv[2], bv[3]);
                row[2 * Q] = make_float4(cv[0], cv[1], cv[2], cv[3]);
                row[3 * Q] = make_float4(dv[0], dv[1], dv[2], dv[3]);
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
            continue;
        }
        for (int r0 = 0; r0 < L - 1; r0 += TR, buf ^= 1) {
            const int nr = min(TR, L - 1 - r0);
            if (use_bulk && tid == 0) bulk_wait_read<1>();
            __syncthreads();
            T* ot = buf ? ot1 : ot0;
            int i = tid / C, c = tid - (tid / C) * C;
            for (int e = tid; e < nr * C; e += kThreads) {
                const int r = r0 + i;
                const T* xr = xs + c * Lp + r;
                const T* kr = ks + c * Lp + r;
                const T xl = xr[0], xh = xr[1];
                T b, two_c, three_d;
                if (L == 2) {
                    b = E::div(E::sub(xh, xl), mult[0]);
                    two_c = T(0);
                    three_d = T(0);
                } else {
                    const T kl = kr[0], kh = kr[1];
                    const T six = E::mul(T(2), E::mul(T(3), E::sub(xh, xl)));
                    const T sr = E::mul(six, rdt[r]);
                    b = kl;
                    two_c = E::mul(E::sub(E::sub(sr, E::mul(T(4), kl)), E::mul(T(2), kh)), rdt[r]);
                    three_d = E::mul(E::add(-sr, E::mul(T(3), E::add(kl, kh))), rdt2[r]);
                }
                T* row = ot + (size_t)i * row_elems + c;
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    const int w = (kq + i) & 3;
                    const T v = (w == 0) ? xl : (w == 1) ? b : (w == 2) ? two_c : three_d;
                    row[w * C] = v;
                }
                i += di;
                c += dc;
                if (c >= C) { c -= C; ++i; }
            }
            T* gp = out + (p * (int64_t)(L - 1) + r0) * row_elems;
            if (use_bulk) {
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store(gp, ot, (uint32_t)((size_t)nr * row_elems * sizeof(T)));
                    bulk_commit();
                }
            } else {
                __syncthreads();
                for (int e = tid; e < nr * row_elems; e += kThreads) gp[e] = ot[e];
            }
        }
    }
    if (use_bulk && tid == 0) bulk_wait_read<0>();
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// Natural cubic spline, one WARP per path (fp32, channels % 4 == 0).  No CTA barrier after the
// prologue: a warp loads its path with 128-bit loads into its own shared-memory tile kept in
// the global [knot][channel] order, sweeps it with lane = (channel pair, chunk of G knots) --
// 64-bit shared loads feed the packed fp32x2 pipe directly -- and writes the coefficient rows
// with four 128-bit stores per (interval, 4 channels).  The back substitution overwrites the
// forward-sweep values in place (warm-up reads of the neighbouring chunk happen before a
// __syncwarp, the chunk's own stores after it), so a path needs two tiles, not three.
// Chunk c of a tile starts at word c * (G * C + padw) with padw chosen so that consecutive
// chunks start C banks apart: the sweeps' strided accesses are conflict free.  The per-knot
// constants are stored already duplicated for the packed pipe, at index i + i / G (one slot of
// padding per chunk, same reason).  ~2.2 K warp-instructions per path against ~10 K for the
// CTA-per-path kernel (whose chunks of 4 knots re-ran a 16-knot warm-up each).
static constexpr int kNatPrefetch = 16;     // 128-bit loads per lane kept in flight for the next path
static constexpr int kNatMaxWarps = 12;

__global__ void __launch_bounds__(32 * kNatMaxWarps)
natural_warp_kernel(const float* __restrict__ x, const float* __restrict__ ws, float* __restrict__ out,
                    int64_t n_paths, int L, int C, int lgG, int padw, int tile_words, int stage_words,
                    int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = 1 << lgG;
    const int nct = (L + G - 1) >> lgG;
    const int Lc = L + nct + 1;
    const int Lc2 = (Lc + 1) & ~1;
    float2* cA = reinterpret_cast<float2*>(smem_raw);          // [i + i/G] = (rdt2, mult)
    float2* cB = cA + Lc2;                                      // [i + i/G] = (rdt, rnd)
    float2* cC = cB + Lc2;                                      // [i] = (rdt, rdt2)
    float* tiles = reinterpret_cast<float*>(cC + ((L + 1) & ~1));
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float rd = ws[i], rd2 = ws[L + i], m = ws[2 * L + i], rn = ws[3 * L + i];
        const int pi = i + (i >> lgG);
        cA[pi] = make_float2(rd2, m);
        cB[pi] = make_float2(rd, rn);
        cC[i] = make_float2(rd, rd2);
    }
    __syncthreads();
    const int wf = (int)ws[4 * L], wb = (int)ws[4 * L + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    float* xs = tiles + (size_t)warp * (2 * tile_words + stage_words);
    float* fk = xs + tile_words;
    float* stage = fk + tile_words;                     // stage_words: coefficient rows staged for a bulk store
    const int C2 = C >> 1, Q = C >> 2;
    const int n_items = C2 * nct;
    const int n_vec = L * Q;
    const int dq_i = 32 / Q, dq_q = 32 - dq_i * Q;
    const int row_step = C + padw;                     // from the last knot of a chunk to the first of the next
    const f2 three = pk2(3.f, 3.f), two = pk2(2.f, 2.f), four = pk2(4.f, 4.f);
    bool saw_nan = false;
    const int RR = stage_words / (4 * C);              // rows per bulk store
    int rl_pre[2], q_pre[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        rl_pre[u] = (lane + 32 * u) / Q;
        q_pre[u] = lane + 32 * u - rl_pre[u] * Q;
    }
    auto word = [&](int i) { return i * C + (i >> lgG) * padw; };

    // The next path's knots are requested before the current path is swept: the DRAM latency of a
    // warp's only global loads is covered by its own arithmetic, not by other warps.
    float4 pre[kNatPrefetch];
    auto request = [&](int64_t path) {
        const float4* xg4 = reinterpret_cast<const float4*>(x + path * (int64_t)L * C);
#pragma unroll
        for (int k = 0; k < kNatPrefetch; ++k)
            if (lane + 32 * k < n_vec) pre[k] = __ldg(xg4 + lane + 32 * k);
    };
    const int64_t stride = (int64_t)gridDim.x * n_warps;
    int64_t p = (int64_t)blockIdx.x * n_warps + warp;
    if (p < n_paths) request(p);

    for (; p < n_paths; p += stride) {
        __syncwarp();                                   // the previous path's coefficient phase has read xs / fk
        {
            int i = lane / Q, q = lane - (lane / Q) * Q;
#pragma unroll
            for (int k = 0; k < kNatPrefetch; ++k) {
                if (lane + 32 * k < n_vec) {
                    const float4 v = pre[k];
                    saw_nan |= is_nan(v.x) | is_nan(v.y) | is_nan(v.z) | is_nan(v.w);
                    *reinterpret_cast<float4*>(xs + word(i) + 4 * q) = v;
                }
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
            const float4* xg4 = reinterpret_cast<const float4*>(x + p * (int64_t)L * C);
#pragma unroll 4
            for (int e = lane + 32 * kNatPrefetch; e < n_vec; e += 32) {       // longer paths: the rest, directly
                const float4 v = __ldg(xg4 + e);
                saw_nan |= is_nan(v.x) | is_nan(v.y) | is_nan(v.z) | is_nan(v.w);
                *reinterpret_cast<float4*>(xs + word(i) + 4 * q) = v;
                i += dq_i;
                q += dq_q;
                if (q >= Q) { q -= Q; ++i; }
            }
        }
        if (p + stride < n_paths) request(p + stride);
        __syncwarp();
        for (int base = 0; base < n_items; base += 32) {             // forward sweep (misc.py:58-61)
            const int it = base + lane;
            if (it >= n_items) continue;
            const int pr = it % C2, j = it / C2;
            const int g0 = j << lgG, g1 = min(g0 + G, L);
            int i = max(g0 - wf, 0);
            const float* px = xs + 2 * pr + word(i);               // x[i]; the forward value of knot i goes to
            const float2* pa = cA + i + (i >> lgG);                // px + tile_words
            f2 f = pk2(0.f, 0.f), sc_prev = pk2(0.f, 0.f);
            f2 x_lo = *reinterpret_cast<const f2*>(px);
            if (i > 0) {
                const float2 a = cA[i - 1 + ((i - 1) >> lgG)];
                const f2 x_before = *reinterpret_cast<const f2*>(xs + 2 * pr + word(i - 1));
                sc_prev = mul2(mul2(three, sub2(x_lo, x_before)), pk2(a.x, a.x));
            }
            // scaled difference 3 (x[i+1] - x[i]) / dt_i^2; rdt2[L-1] = 0 closes the system (cubic.py:36-39)
            auto knot = [&](const f2 x_hi, const float2 a, float* dst, bool keep) {
                const f2 sc = mul2(mul2(three, sub2(x_hi, x_lo)), pk2(a.x, a.x));
                x_lo = x_hi;
                f = sub2(add2(sc, sc_prev), mul2(pk2(a.y, a.y), f));
                sc_prev = sc;
                if (keep) *reinterpret_cast<f2*>(dst) = f;
            };
            while (i < g1) {                                         // chunk by chunk: warm-up chunks, then its own
                const int seg_end = min((i | (G - 1)) + 1, g1);
                const bool keep = i >= g0;
                int r = seg_end - 1 - i;
                for (; r >= 4; r -= 4) {                             // loads of four knots ahead of their stores
                    f2 xh[4];
                    float2 aa[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        xh[u] = *reinterpret_cast<const f2*>(px + (u + 1) * C);
                        aa[u] = pa[u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) knot(xh[u], aa[u], const_cast<float*>(px) + u * C + tile_words, keep);
                    px += 4 * C;
                    pa += 4;
                }
                for (; r > 0; --r) {
                    knot(*reinterpret_cast<const f2*>(px + C), *pa, const_cast<float*>(px) + tile_words, keep);
                    px += C;
                    ++pa;
                }
                const int step = (seg_end == L) ? 0 : row_step;      // the right neighbour of the last knot is itself
                knot(*reinterpret_cast<const f2*>(px + step), *pa, const_cast<float*>(px) + tile_words, keep);
                px += step;
                pa += 2;
                i = seg_end;
            }
        }
        __syncwarp();
        for (int base = 0; base < n_items; base += 32) {             // back substitution (misc.py:63-65), in place
            const int it = base + lane;
            const bool active = it < n_items;
            const int pr = active ? it % C2 : 0, j = active ? it / C2 : 0;
            const int g0 = j << lgG, g1 = min(g0 + G, L);
            int i = min(g1 - 1 + wb, L - 1);
            float* pf = fk + 2 * pr + word(i);
            const float2* pb = cB + i + (i >> lgG);
            f2 k = pk2(0.f, 0.f);
            if (active) {
                while (i >= g1) {                                    // warm-up in the chunks to the right: reads only
                    const int seg_lo = max(i & ~(G - 1), g1);        // g1 is a chunk boundary whenever this loop runs
                    for (; i > seg_lo; --i) {
                        const float2 b = *pb;
                        k = mul2(sub2(*reinterpret_cast<const f2*>(pf), mul2(pk2(b.x, b.x), k)), pk2(b.y, b.y));
                        pf -= C;
                        --pb;
                    }
                    const float2 b = *pb;
                    k = mul2(sub2(*reinterpret_cast<const f2*>(pf), mul2(pk2(b.x, b.x), k)), pk2(b.y, b.y));
                    pf -= row_step;
                    pb -= 2;
                    i = seg_lo - 1;
                }
            }
            __syncwarp();
            if (active) {
                for (; i >= g0 + 3; i -= 4) {                       // loads of four knots ahead of their stores
                    f2 fv[4];
                    float2 bb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        fv[u] = *reinterpret_cast<const f2*>(pf - u * C);
                        bb[u] = *(pb - u);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        k = mul2(sub2(fv[u], mul2(pk2(bb[u].x, bb[u].x), k)), pk2(bb[u].y, bb[u].y));
                        *reinterpret_cast<f2*>(pf - u * C) = k;
                    }
                    pf -= 4 * C;
                    pb -= 4;
                }
                for (; i >= g0; --i) {
                    const float2 b = *pb;
                    k = mul2(sub2(*reinterpret_cast<const f2*>(pf), mul2(pk2(b.x, b.x), k)), pk2(b.y, b.y));
                    *reinterpret_cast<f2*>(pf) = k;
                    pf -= C;
                    --pb;
                }
            }
        }
        __syncwarp();
        {
            // coefficient rows: RR whole rows per round (two items per lane, loads of both ahead of the math), staged
            // in shared memory exactly as they lie in global memory and written by one bulk (TMA) store of the warp's
            // elected lane.  Lane = (row, 4 channels); its four 16-byte blocks go to the staging row in an order
            // rotated by the row number, which spreads the lanes of one store instruction over the banks (rows are
            // 16 C bytes apart).  The staging buffer is reused once the previous bulk store has read it.
            float* og = out + p * (int64_t)(L - 1) * 4 * C;
            for (int r0 = 0; r0 < L - 1; r0 += RR) {
                const int n_it = min(RR, L - 1 - r0) * Q;
                for (int base = 0; base < n_it; base += 64) {
                    int rl[2], qq[2];
                    bool valid[2];
                    float4 xl[2], xh[2], kl[2], kh[2];
                    float2 rr[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int it = base + lane + 32 * u;
                        rl[u] = base == 0 ? rl_pre[u] : it / Q;
                        qq[u] = base == 0 ? q_pre[u] : it - (it / Q) * Q;
                        valid[u] = it < n_it;
                        if (valid[u]) {
                            const int i = r0 + rl[u];
                            const int w0 = word(i) + 4 * qq[u], w1 = word(i + 1) + 4 * qq[u];
                            xl[u] = *reinterpret_cast<const float4*>(xs + w0);
                            xh[u] = *reinterpret_cast<const float4*>(xs + w1);
                            kl[u] = *reinterpret_cast<const float4*>(fk + w0);
                            kh[u] = *reinterpret_cast<const float4*>(fk + w1);
                            rr[u] = cC[i];
                        }
                    }
                    if (base == 0) {
                        if (lane == 0) bulk_wait_read<0>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (!valid[u]) continue;
                        const f2 rdp = pk2(rr[u].x, rr[u].x), rd2p = pk2(rr[u].y, rr[u].y);
                        float4 cv, dv;
                        {
                            const f2 l = pk2(kl[u].x, kl[u].y), h = pk2(kh[u].x, kh[u].y);
                            const f2 sr = mul2(mul2(two, mul2(three, sub2(pk2(xh[u].x, xh[u].y), pk2(xl[u].x, xl[u].y)))), rdp);
                            upk2(mul2(sub2(sub2(sr, mul2(four, l)), mul2(two, h)), rdp), cv.x, cv.y);   // cubic.py:45-47
                            upk2(mul2(sub2(mul2(three, add2(l, h)), sr), rd2p), dv.x, dv.y);             // cubic.py:48-50
                        }
                        {
                            const f2 l = pk2(kl[u].z, kl[u].w), h = pk2(kh[u].z, kh[u].w);
                            const f2 sr = mul2(mul2(two, mul2(three, sub2(pk2(xh[u].z, xh[u].w), pk2(xl[u].z, xl[u].w)))), rdp);
                            upk2(mul2(sub2(sub2(sr, mul2(four, l)), mul2(two, h)), rdp), cv.z, cv.w);
                            upk2(mul2(sub2(mul2(three, add2(l, h)), sr), rd2p), dv.z, dv.w);
                        }
                        // rotate (a, b, 2c, 3d) left by rl & 3: two conditional stages of register selects
                        const bool r1 = rl[u] & 1, r2 = rl[u] & 2;
                        const float4 v0 = r1 ? kl[u] : xl[u], v1 = r1 ? cv : kl[u], v2 = r1 ? dv : cv, v3 = r1 ? xl[u] : dv;
                        const float4 u0 = r2 ? v2 : v0, u1 = r2 ? v3 : v1, u2 = r2 ? v0 : v2, u3 = r2 ? v1 : v3;
                        float4* row = reinterpret_cast<float4*>(stage) + (size_t)rl[u] * (4 * Q) + qq[u];
                        const int b0 = rl[u] & 3;
                        row[b0 * Q] = u0;
                        row[((b0 + 1) & 3) * Q] = u1;
                        row[((b0 + 2) & 3) * Q] = u2;
                        row[((b0 + 3) & 3) * Q] = u3;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    bulk_store(og + (size_t)r0 * 4 * C, stage, (uint32_t)(n_it * 64));     // an item is 4 x 16 bytes
                    bulk_commit();
                }
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// One CTA per group of S paths.  Phase 1: coalesced load, transposed into shared memory as
// [series][knot] so that the thread that owns a series walks contiguous words.  Phase 2: one
// thread per series runs the forward sweep and the back substitution in shared memory.
// Phase 3: all threads turn (x, slope) into the [a|b|2c|3d] rows, staged per tile and written
// with bulk stores exactly like the Hermite kernel.
template <typename T>
__global__ void __launch_bounds__(kThreads)
natural_kernel(const T* __restrict__ x, const T* __restrict__ ws, T* __restrict__ out, int64_t n_paths, int L, int C,
               int S, int Lp, int TR, int use_bulk, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using E = exact<T>;
    const int row_elems = 4 * C;
    const int path_stride = C * Lp + 1;
    T* ot0 = reinterpret_cast<T*>(smem_raw);
    T* ot1 = ot0 + (size_t)TR * row_elems;
    T* rdt = ot1 + (size_t)TR * row_elems;
    T* rdt2 = rdt + L;
    T* mult = rdt2 + L;
    T* rnd = mult + L;
    T* xs = rnd + L;
    T* ks = xs + (size_t)S * path_stride;

    const int tid = threadIdx.x;
    for (int e = tid; e < 4 * L; e += kThreads) rdt[e] = ws[e];
    bool saw_nan = false;
    int buf = 0;
    const int64_t n_groups = (n_paths + S - 1) / S;
    const int di = kThreads / C, dc = kThreads - di * C;

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p0 = grp * S;
        const int np = (int)min((int64_t)S, n_paths - p0);
        __syncthreads();   // previous group's phase 3 is done with xs / ks
        {
            const T* xg = x + p0 * L * C;
            int i = tid / C, c = tid - (tid / C) * C, p = 0;
            while (i >= L) { i -= L; ++p; }
            const int total = np * L * C;
            for (int e = tid; e < total; e += kThreads) {
                const T v = xg[e];
                saw_nan |= is_nan(v);
                xs[p * path_stride + c * Lp + i] = v;
                i += di;
                c += dc;
                if (c >= C) { c -= C; ++i; }
                while (i >= L) { i -= L; ++p; }
            }
        }
        __syncthreads();
        if (tid < np * C && L > 2) {
            const int p = tid / C, c = tid - p * C;
            const T* xr = xs + p * path_stride + c * Lp;
            T* kr = ks + p * path_stride + c * Lp;
            // forward sweep (misc.py:58-61); rhs[i] = scaled[i] + scaled[i-1] (cubic.py:36-39)
            T x_lo = xr[0], x_hi = xr[1];
            T scaled_prev = E::mul(E::mul(T(3), E::sub(x_hi, x_lo)), rdt2[0]);
            T f = scaled_prev;
            kr[0] = f;
            for (int i = 1; i < L; ++i) {
                T rhs;
                if (i < L - 1) {
                    x_lo = x_hi;
                    x_hi = xr[i + 1];
                    const T scaled = E::mul(E::mul(T(3), E::sub(x_hi, x_lo)), rdt2[i]);
                    rhs = E::add(scaled, scaled_prev);
                    scaled_prev = scaled;
                } else {
                    rhs = E::add(T(0), scaled_prev);
                }
                f = E::sub(rhs, E::mul(mult[i], f));
                kr[i] = f;
            }
            // back substitution (misc.py:63-65), dividing by nd as a multiplication by 1/nd
            T k = E::mul(f, rnd[L - 1]);
            kr[L - 1] = k;
            for (int i = L - 2; i >= 0; --i) {
                k = E::mul(E::sub(kr[i], E::mul(rdt[i], k)), rnd[i]);
                kr[i] = k;
            }
        }
        __syncthreads();
        // phase 3
        for (int p = 0; p < np; ++p) {
            for (int r0 = 0; r0 < L - 1; r0 += TR, buf ^= 1) {
                const int nr = min(TR, L - 1 - r0);
                if (use_bulk && tid == 0) bulk_wait_read<1>();
                __syncthreads();
                T* ot = buf ? ot1 : ot0;
                int i = tid / C, c = tid - (tid / C) * C;
                for (int e = tid; e < nr * C; e += kThreads) {
                    const int r = r0 + i;
                    const T* xr = xs + p * path_stride + c * Lp + r;
                    const T* kr = ks + p * path_stride + c * Lp + r;
                    const T xl = xr[0], xh = xr[1];
                    T b, two_c, three_d;
                    if (L == 2) {                       // cubic.py:16-20
                        b = E::div(E::sub(xh, xl), mult[0]);
                        two_c = T(0);
                        three_d = T(0);
                    } else {
                        const T kl = kr[0], kh = kr[1];
                        const T six = E::mul(T(2), E::mul(T(3), E::sub(xh, xl)));
                        const T sr = E::mul(six, rdt[r]);
                        b = kl;
                        two_c = E::mul(E::sub(E::sub(sr, E::mul(T(4), kl)), E::mul(T(2), kh)), rdt[r]);      // :45-47
                        three_d = E::mul(E::add(-sr, E::mul(T(3), E::add(kl, kh))), rdt2[r]);                // :48-50
                    }
                    T* row = ot + (size_t)i * row_elems + c;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int w = (k + i) & 3;
                        const T v = (w == 0) ? xl : (w == 1) ? b : (w == 2) ? two_c : three_d;
                        row[w * C] = v;
                    }
                    i += di;
                    c += dc;
                    if (c >= C) { c -= C; ++i; }
                }
                T* gp = out + ((p0 + p) * (int64_t)(L - 1) + r0) * row_elems;
                if (use_bulk) {
                    fence_proxy_async_smem();
                    __syncthreads();
                    if (tid == 0) {
                        bulk_store(gp, ot, (uint32_t)((size_t)nr * row_elems * sizeof(T)));
                        bulk_commit();
                    }
                } else {
                    __syncthreads();
                    for (int e = tid; e < nr * row_elems; e += kThreads) gp[e] = ot[e];
                }
            }
        }
    }
    if (use_bulk && tid == 0) bulk_wait_read<0>();
    if (saw_nan && flags != nullptr) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// =========================================================================================
// Natural cubic spline with missing values, one thread per series  (cubic.py:56-167)
// =========================================================================================
// Irregular path: each series has its own set of observed knots, so the tridiagonal system
// differs per series.  scratch holds, per series, 3 rows of length L in a [row][knot][series]
// layout (coalesced across the warp): compacted rhs/solution, eliminated diagonal, and the
// knot slopes.  Exact arithmetic (the reference order) throughout.
template <typename T, bool UNIT>
__global__ void __launch_bounds__(128)
natural_missing_kernel(const T* __restrict__ x, const T* __restrict__ t, T* __restrict__ out, T* __restrict__ scratch,
                       int64_t n_series, int L, int C, int version) {
    using E = exact<T>;
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_series) return;
    const int64_t p = g / C;
    const int c = (int)(g - p * C);
    const T* xs = x + p * L * C + c;
    T* os = out + p * (int64_t)(L - 1) * 4 * C + c;
    auto time_of = [&](int i) -> T { return UNIT ? T(i) : t[i]; };
    // scratch rows, element (row, j) of this series at scratch[(row * L + j) * n_series + g]
    auto S0 = [&](int j) -> T& { return scratch[((int64_t)0 * L + j) * n_series + g]; };   // rhs -> fwd -> slope
    auto S1 = [&](int j) -> T& { return scratch[((int64_t)1 * L + j) * n_series + g]; };   // eliminated diagonal
    auto S2 = [&](int j) -> T& { return scratch[((int64_t)2 * L + j) * n_series + g]; };   // compacted value
    auto S3 = [&](int j) -> T& { return scratch[((int64_t)3 * L + j) * n_series + g]; };   // compacted time

    // first / last observation
    int first = -1, last = -1;
    for (int i = 0; i < L; ++i) {
        if (!is_nan(xs[(int64_t)i * C])) {
            if (first < 0) first = i;
            last = i;
        }
    }
    const int rows = L - 1;
    if (first < 0) {                                             // cubic.py:85-92
        for (int i = 0; i < rows; ++i)
            for (int w = 0; w < 4; ++w) os[((int64_t)i * 4 + w) * C] = T(0);
        return;
    }
    const T vfirst = xs[(int64_t)first * C], vlast = xs[(int64_t)last * C];
    // compact the (imputed) observed knots: cubic.py:101-133
    int m = 0;
    for (int i = 0; i < L; ++i) {
        T v = xs[(int64_t)i * C];
        if (is_nan(v)) {
            if (version == 0) {
                if (i == 0) v = vfirst;
                else if (i == L - 1) v = vlast;
            } else {
                if (i < first) v = vfirst;
                else if (i > last) v = vlast;
            }
        }
        if (!is_nan(v)) {
            S2(m) = v;
            S3(m) = time_of(i);
            ++m;
        }
    }
    // m >= 2 always (both ends are observed after imputation)
    if (m > 2) {
        // system of cubic.py:23-39 on the compacted knots, Thomas solve of misc.py:52-65
        T rprev = E::div(T(1), E::sub(S3(1), S3(0)));
        T scaled_prev = E::mul(E::mul(T(3), E::sub(S2(1), S2(0))), E::mul(rprev, rprev));
        T nd = E::mul(T(2), rprev);
        T f = scaled_prev;
        S0(0) = f;
        S1(0) = nd;
        for (int j = 1; j < m; ++j) {
            T diag, rhs, rcur = T(0);
            if (j < m - 1) {
                rcur = E::div(T(1), E::sub(S3(j + 1), S3(j)));
                const T scaled = E::mul(E::mul(T(3), E::sub(S2(j + 1), S2(j))), E::mul(rcur, rcur));
                diag = E::mul(T(2), E::add(rcur, rprev));
                rhs = E::add(scaled, scaled_prev);
                scaled_prev = scaled;
            } else {
                diag = E::mul(T(2), E::add(T(0), rprev));
                rhs = E::add(T(0), scaled_prev);
            }
            const T w = E::div(rprev, nd);
            nd = E::sub(diag, E::mul(w, rprev));
            f = E::sub(rhs, E::mul(w, f));
            S0(j) = f;
            S1(j) = nd;
            rprev = rcur;
        }
        T k = E::div(f, nd);
        S0(m - 1) = k;
        for (int j = m - 2; j >= 0; --j) {
            const T up = E::div(T(1), E::sub(S3(j + 1), S3(j)));
            k = E::div(E::sub(S0(j), E::mul(up, k)), S1(j));
            S0(j) = k;
        }
    }
    // walk the original intervals; piece q spans compacted knots q, q+1   (cubic.py:147-162)
    int q = -1;
    T pa = T(0), pb = T(0), pc = T(0), pd = T(0), anchor = T(0);
    int next_obs = 0;   // index into compacted knots of the next observed time
    for (int i = 0; i < rows; ++i) {
        const T ti = time_of(i);
        if (next_obs < m && ti >= S3(next_obs)) {
            q = next_obs;
            ++next_obs;
            anchor = S3(q);
            const T xl = S2(q), xh = S2(q + 1);
            if (m == 2) {                                        // cubic.py:16-20
                pa = xl;
                pb = E::div(E::sub(xh, xl), E::sub(S3(1), S3(0)));
                pc = T(0);
                pd = T(0);
            } else {
                const T r = E::div(T(1), E::sub(S3(q + 1), S3(q)));
                const T r2 = E::mul(r, r);
                const T kl = S0(q), kh = S0(q + 1);
                const T six = E::mul(T(2), E::mul(T(3), E::sub(xh, xl)));
                const T sr = E::mul(six, r);
                pa = xl;
                pb = kl;
                pc = E::mul(E::sub(E::sub(sr, E::mul(T(4), kl)), E::mul(T(2), kh)), r);
                pd = E::mul(E::add(-sr, E::mul(T(3), E::add(kl, kh))), r2);
            }
        }
        const T off = E::sub(anchor, ti);
        const T inner = E::mul(E::sub(E::mul(T(0.5), pc), E::div(E::mul(pd, off), T(3))), off);
        os[((int64_t)i * 4 + 0) * C] = E::add(pa, E::mul(E::sub(inner, pb), off));
        os[((int64_t)i * 4 + 1) * C] = E::add(pb, E::mul(E::sub(E::mul(pd, off), pc), off));
        os[((int64_t)i * 4 + 2) * C] = E::sub(pc, E::mul(E::mul(T(2), pd), off));
        os[((int64_t)i * 4 + 3) * C] = pd;
    }
}

// torch.isnan(x).any() of interpolation_linear.py:169 as one pass that only sets a flag.
template <typename T>
__global__ void __launch_bounds__(kThreads)
nan_flag_kernel(const T* __restrict__ x, int64_t n, int32_t* __restrict__ flags) {
    bool seen = false;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += stride) seen |= is_nan(x[i]);
    if (__syncthreads_or(seen) && threadIdx.x == 0) atomicOr(flags, TCDE_FLAG_NAN_SEEN);
}

// =========================================================================================
// launchers
// =========================================================================================
static int g_fill_variant = 0;        // 0 = scan / ballot warp-per-path gap fill when they fit, 1 = one thread per
                                      // series, 2 = never the scan kernel
static int g_natural_variant = 0;     // 0 = warp per path / windowed CTA sweeps when they fit, 1 = one thread per
                                      // series, 2 = never the warp-per-path kernel

static int persistent_grid(const void* kernel, int threads, size_t smem, int64_t n_items) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int64_t g = (int64_t)sm_count() * per_sm;
    if (g > n_items) g = n_items;
    if (g < 1) g = 1;
    return (int)g;
}

static constexpr size_t kMaxSmem = 200 * 1024;

static int launch_hermite_vec4(const float* x, const float* t, float* out, int64_t n_paths, int L, int C,
                               int32_t* flags, cudaStream_t stream) {
    const size_t row_bytes = (size_t)16 * C;
    int TR = (int)(16384 / row_bytes);
    if (TR < 1) TR = 1;
    if (TR > L - 1) TR = L - 1;
    const size_t smem = 2 * TR * row_bytes + 16;
    const int tiles = (L - 1 + TR - 1) / TR;
    auto kern = t ? hermite_vec4_kernel<false> : hermite_vec4_kernel<true>;
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = persistent_grid((const void*)kern, kThreads, smem, n_paths * tiles);
    kern<<<grid, kThreads, smem, stream>>>(x, t, out, n_paths, L, C, TR, tiles, flags);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

template <typename T>
static int launch_hermite(const T* x, const T* t, T* out, int64_t n_paths, int L, int C, int32_t* flags,
                          cudaStream_t stream) {
    if (sizeof(T) == 4 && (C & 3) == 0 && C <= 512 && aligned16(x) && aligned16(out))
        return launch_hermite_vec4((const float*)x, (const float*)t, (float*)out, n_paths, L, C, flags, stream);
    const size_t row_bytes = (size_t)4 * C * sizeof(T);
    int TR = (int)(16384 / row_bytes);
    if (TR < 1) TR = 1;
    if (TR > L - 1) TR = L - 1;
    const size_t smem = 2 * TR * row_bytes + (size_t)(TR + 2) * C * sizeof(T) + (size_t)(TR + 2) * sizeof(T) + 16;
    TCDE_CHECK_SUPPORTED(smem <= kMaxSmem, "hermite: channels=%d needs %zu bytes of shared memory (max %zu)", C, smem,
                         kMaxSmem);
    const int tiles = (L - 1 + TR - 1) / TR;
    const int use_bulk = aligned16(out) ? 1 : 0;
    auto kern = t ? hermite_kernel<T, false> : hermite_kernel<T, true>;
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = persistent_grid((const void*)kern, kThreads, smem, n_paths * tiles);
    kern<<<grid, kThreads, smem, stream>>>(x, t, out, n_paths, L, C, TR, tiles, use_bulk, flags);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

template <typename T>
static int launch_natural(const T* x, const T* t, T* out, T* ws, int64_t n_paths, int L, int C, int32_t* flags,
                          cudaStream_t stream) {
    natural_prep_kernel<T><<<1, 128, 0, stream>>>(t, ws, L);
    TCDE_CHECK_CUDA(cudaGetLastError());
    const size_t row_bytes = (size_t)4 * C * sizeof(T);
    int TR = (int)(8192 / row_bytes);
    if (TR < 1) TR = 1;
    if (TR > L - 1) TR = L - 1;
    if constexpr (sizeof(T) == 4) {
        // warp per path: fp32, 128-bit rows, two tiles per warp in shared memory
        if ((C & 3) == 0 && L > 2 && aligned16(out) && aligned16(x) && g_natural_variant == 0) {
            const int C2 = C / 2;
            const int n_chunks = C2 >= 32 ? 1 : 32 / C2;
            int lgG = 0;
            while ((1 << lgG) * n_chunks < L) ++lgG;
            const int G = 1 << lgG;
            const int nct = (L + G - 1) / G;
            const int padw = (int)((((int64_t)C - (int64_t)G * C) % 32 + 32) % 32);
            const int tile_words = L * C + nct * padw;
            const int rows_per_round = C / 4 >= 32 ? 2 : 64 / (C / 4);       // two (row, 4 channels) items per lane
            const int stage_words = rows_per_round * 4 * C;
            const size_t fixed = (size_t)2 * ((L + nct + 2) & ~1) * 8 + (size_t)((L + 1) & ~1) * 8;
            const size_t per_warp = (size_t)(2 * tile_words + stage_words) * 4;
            int wpc = fixed < 220 * 1024 ? (int)((220 * 1024 - fixed) / per_warp) : 0;
            if (wpc > kNatMaxWarps) wpc = kNatMaxWarps;
            if (wpc >= 2) {
                const size_t smem_k = fixed + wpc * per_warp;
                auto kw = natural_warp_kernel;
                TCDE_CHECK_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k));
                int64_t grid = (n_paths + wpc - 1) / wpc;
                if (grid > sm_count()) grid = sm_count();
                kw<<<(int)grid, 32 * wpc, smem_k, stream>>>((const float*)x, (const float*)ws, (float*)out, n_paths, L, C,
                                                           lgG, padw, tile_words, stage_words, flags);
                TCDE_CHECK_CUDA(cudaGetLastError());
                return TCDE_OK;
            }
        }
    }
    {
        // parallel windowed sweeps: one path per CTA iteration, thread = (series, chunk of G knots)
        const int Lpw = ((L + 31) / 32) * 32 + 1;
        int G = (int)(((int64_t)L * C + kThreads - 1) / kThreads);
        if (G < 4) G = 4;
        // coefficient tiles: at least one (interval, 4-channel) item per thread and tile (a tile costs two
        // CTA barriers and a bulk store), at most 16 KB per staging buffer
        int TRw = TR;
        while (TRw < L - 1 && (size_t)TRw * (C / 4 > 0 ? C / 4 : 1) < (size_t)kThreads && (size_t)(2 * TRw) * row_bytes <= 16384) TRw *= 2;
        if (TRw > L - 1) TRw = L - 1;
        const bool direct = (sizeof(T) == 4) && ((C & 3) == 0) && L > 2 && aligned16(out) && aligned16(x);
        const int TR = direct ? 1 : TRw;      // the vectorised path stores rows straight to global memory
        const size_t smem_w = 2 * TR * row_bytes + (size_t)(4 * L + 8) * sizeof(T) + (size_t)3 * C * Lpw * sizeof(T) + 16;
        if (smem_w <= 64 * 1024 && g_natural_variant != 1) {
            auto kw = natural_win_kernel<T>;
            TCDE_CHECK_CUDA(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            const int grid = persistent_grid((const void*)kw, kThreads, smem_w, n_paths);
            kw<<<grid, kThreads, smem_w, stream>>>(x, ws, out, n_paths, L, C, Lpw, G, TR, aligned16(out) ? 1 : 0, flags);
            TCDE_CHECK_CUDA(cudaGetLastError());
            return TCDE_OK;
        }
    }
    int cp = 1;
    while (cp < C && cp < 32) cp <<= 1;
    const int Lp = ((L + 31) / 32) * 32 + 32 / cp;
    const size_t per_path = (size_t)2 * ((size_t)C * Lp + 1) * sizeof(T);
    const size_t fixed = 2 * TR * row_bytes + (size_t)4 * L * sizeof(T) + 16;
    int S = (32 + C - 1) / C;
    while (S > 1 && fixed + S * per_path > 100 * 1024) --S;
    const size_t smem = fixed + S * per_path;
    TCDE_CHECK_SUPPORTED(smem <= kMaxSmem,
                         "natural cubic: length=%d x channels=%d needs %zu bytes of shared memory (max %zu)", L, C,
                         smem, kMaxSmem);
    const int use_bulk = aligned16(out) ? 1 : 0;
    auto kern = natural_kernel<T>;
    TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_groups = (n_paths + S - 1) / S;
    const int grid = persistent_grid((const void*)kern, kThreads, smem, n_groups);
    kern<<<grid, kThreads, smem, stream>>>(x, ws, out, n_paths, L, C, S, Lp, TR, use_bulk, flags);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

static int check_shape(const void* x, const void* out, int64_t n_paths, int64_t length, int64_t channels, int dtype) {
    TCDE_CHECK_ARG(x != nullptr && out != nullptr, "null data pointer");
    TCDE_CHECK_ARG(n_paths >= 0 && channels >= 1, "n_paths=%lld channels=%lld", (long long)n_paths,
                   (long long)channels);
    TCDE_CHECK_ARG(length >= 2, "length=%lld (need at least 2 knots, misc.py:96-98)", (long long)length);
    TCDE_CHECK_ARG(dtype == TCDE_F32 || dtype == TCDE_F64, "dtype=%d", dtype);
    TCDE_CHECK_SUPPORTED(length < (1 << 24) && channels < (1 << 20), "length / channels too large");
    return TCDE_OK;
}

}  // namespace tcde

using namespace tcde;

extern "C" int tcde_hermite_bdiff_coeffs(const void* x, const void* t, void* coeffs, int64_t n_paths, int64_t length,
                                         int64_t channels, int dtype, int32_t* flags, void* stream) {
    int rc = check_shape(x, coeffs, n_paths, length, channels, dtype);
    if (rc != TCDE_OK) return rc;
    if (n_paths == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TCDE_F32)
        return launch_hermite<float>((const float*)x, (const float*)t, (float*)coeffs, n_paths, (int)length,
                                     (int)channels, flags, s);
    return launch_hermite<double>((const double*)x, (const double*)t, (double*)coeffs, n_paths, (int)length,
                                  (int)channels, flags, s);
}

extern "C" int tcde_linear_fill(const void* x, const void* t, void* out, int64_t n_paths, int64_t length,
                                int64_t channels, int dtype, int32_t* flags, void* stream) {
    int rc = check_shape(x, out, n_paths, length, channels, dtype);
    if (rc != TCDE_OK) return rc;
    const int64_t n_series = n_paths * channels;
    if (n_series == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int L = (int)length, C = (int)channels;
    if (C <= 32 && L < kFillNone && g_fill_variant == 0) {
        // scan kernel: lane = (channel, chunk of positions), tile in the global layout
        const size_t elem = (dtype == TCDE_F32) ? 4 : 8;
        const int n_chunks = 32 / C;
        int lgG = 0;
        while ((1 << lgG) * n_chunks < L) ++lgG;
        const int G = 1 << lgG;
        const int nct = (L + G - 1) / G;
        const int bank_words = (int)(128 / elem);
        const int padw = (int)((((int64_t)C - (int64_t)G * C) % bank_words + bank_words) % bank_words);
        const int tile_words = (L * C + nct * padw + 3) & ~3;
        const size_t smem = (t ? (size_t)((L + 3) & ~3) * elem : 0) + (size_t)(kThreads / 32) * tile_words * elem;
        if (smem <= 100 * 1024 && lgG < kFillDist) {          // hole-list distances are 11-bit
            const void* kern = (dtype == TCDE_F32)
                ? (t ? (const void*)linear_fill_scan_kernel<float, false> : (const void*)linear_fill_scan_kernel<float, true>)
                : (t ? (const void*)linear_fill_scan_kernel<double, false> : (const void*)linear_fill_scan_kernel<double, true>);
            TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = persistent_grid(kern, kThreads, smem, (n_paths + kThreads / 32 - 1) / (kThreads / 32));
            const int Li = L, Ci = C;
            void* args[] = {(void*)&x, (void*)&t, (void*)&out, (void*)&n_paths, (void*)&Li, (void*)&Ci, (void*)&lgG,
                            (void*)&padw, (void*)&tile_words, (void*)&flags};
            TCDE_CHECK_CUDA(cudaLaunchKernel(kern, dim3(grid), dim3(kThreads), args, smem, s));
            return TCDE_OK;
        }
    }
    if (flags != nullptr) {               // the other kernels do not report: a separate pass over x
        rc = tcde_nan_flag(x, n_series * L, dtype, flags, stream);
        if (rc != TCDE_OK) return rc;
    }
    {
        // warp-per-path kernel when a path fits a warp's shared-memory tile
        const int Lp = ((L + 31) / 32) * 32 + 1;
        const size_t elem = (dtype == TCDE_F32) ? 4 : 8;
        const size_t smem = (size_t)(kThreads / 32) * C * Lp * elem;
        if (L <= 32 * kFillRounds && smem <= 72 * 1024 && g_fill_variant != 1) {
            const void* kern = (dtype == TCDE_F32)
                ? (t ? (const void*)linear_fill_warp_kernel<float, false> : (const void*)linear_fill_warp_kernel<float, true>)
                : (t ? (const void*)linear_fill_warp_kernel<double, false> : (const void*)linear_fill_warp_kernel<double, true>);
            TCDE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = persistent_grid(kern, kThreads, smem, (n_paths + kThreads / 32 - 1) / (kThreads / 32));
            const int Li = L, Ci = C, Lpi = Lp;
            void* args[] = {(void*)&x, (void*)&t, (void*)&out, (void*)&n_paths, (void*)&Li, (void*)&Ci, (void*)&Lpi};
            TCDE_CHECK_CUDA(cudaLaunchKernel(kern, dim3(grid), dim3(kThreads), args, smem, s));
            return TCDE_OK;
        }
    }
    const int64_t blocks = (n_series + kThreads - 1) / kThreads;
    TCDE_CHECK_SUPPORTED(blocks < (1ll << 31), "too many series");
    if (dtype == TCDE_F32) {
        if (t) linear_fill_kernel<float, false><<<(unsigned)blocks, kThreads, 0, s>>>((const float*)x, (const float*)t, (float*)out, n_series, L, C);
        else linear_fill_kernel<float, true><<<(unsigned)blocks, kThreads, 0, s>>>((const float*)x, nullptr, (float*)out, n_series, L, C);
    } else {
        if (t) linear_fill_kernel<double, false><<<(unsigned)blocks, kThreads, 0, s>>>((const double*)x, (const double*)t, (double*)out, n_series, L, C);
        else linear_fill_kernel<double, true><<<(unsigned)blocks, kThreads, 0, s>>>((const double*)x, nullptr, (double*)out, n_series, L, C);
    }
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

template <bool RECT>
static int launch_hold(const void* x, void* out, int64_t n_paths, int64_t length, int64_t channels, int64_t time_index,
                       int dtype, int32_t* flags, void* stream) {
    int rc = check_shape(x, out, n_paths, length, channels, dtype);
    if (rc != TCDE_OK) return rc;
    if (RECT) TCDE_CHECK_ARG(time_index >= 0 && time_index < channels, "time_index=%lld", (long long)time_index);
    const int64_t n_series = n_paths * channels;
    if (n_series == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t blocks = (n_series + kThreads - 1) / kThreads;
    TCDE_CHECK_SUPPORTED(blocks < (1ll << 31), "too many series");
    if (dtype == TCDE_F32)
        hold_kernel<float, RECT><<<(unsigned)blocks, kThreads, 0, s>>>((const float*)x, (float*)out, n_series, (int)length, (int)channels, (int)time_index, flags);
    else
        hold_kernel<double, RECT><<<(unsigned)blocks, kThreads, 0, s>>>((const double*)x, (double*)out, n_series, (int)length, (int)channels, (int)time_index, flags);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

extern "C" int tcde_forward_fill(const void* x, void* out, int64_t n_paths, int64_t length, int64_t channels,
                                 int dtype, int32_t* flags, void* stream) {
    return launch_hold<false>(x, out, n_paths, length, channels, 0, dtype, flags, stream);
}

extern "C" int tcde_rectilinear_prepare(const void* x, void* out, int64_t n_paths, int64_t length, int64_t channels,
                                        int64_t time_index, int dtype, int32_t* flags, void* stream) {
    return launch_hold<true>(x, out, n_paths, length, channels, time_index, dtype, flags, stream);
}

extern "C" int tcde_natural_cubic_coeffs(const void* x, const void* t, void* coeffs, void* workspace, int64_t n_paths,
                                         int64_t length, int64_t channels, int dtype, int32_t* flags, void* stream) {
    int rc = check_shape(x, coeffs, n_paths, length, channels, dtype);
    if (rc != TCDE_OK) return rc;
    TCDE_CHECK_ARG(workspace != nullptr, "null workspace");
    if (n_paths == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TCDE_F32)
        return launch_natural<float>((const float*)x, (const float*)t, (float*)coeffs, (float*)workspace, n_paths,
                                     (int)length, (int)channels, flags, s);
    return launch_natural<double>((const double*)x, (const double*)t, (double*)coeffs, (double*)workspace, n_paths,
                                  (int)length, (int)channels, flags, s);
}

extern "C" int64_t tcde_natural_cubic_missing_scratch_bytes(int64_t n_paths, int64_t length, int64_t channels,
                                                            int dtype) {
    const int64_t elem = (dtype == TCDE_F64) ? 8 : 4;
    return 4 * length * n_paths * channels * elem;
}

extern "C" int tcde_natural_cubic_coeffs_missing(const void* x, const void* t, void* coeffs, void* scratch,
                                                 int64_t n_paths, int64_t length, int64_t channels, int version,
                                                 int dtype, void* stream) {
    int rc = check_shape(x, coeffs, n_paths, length, channels, dtype);
    if (rc != TCDE_OK) return rc;
    TCDE_CHECK_ARG(scratch != nullptr, "null scratch");
    TCDE_CHECK_ARG(version == 0 || version == 1, "version=%d", version);
    const int64_t n_series = n_paths * channels;
    if (n_series == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int threads = 128;
    const int64_t blocks = (n_series + threads - 1) / threads;
    TCDE_CHECK_SUPPORTED(blocks < (1ll << 31), "too many series");
    const int L = (int)length, C = (int)channels;
    if (dtype == TCDE_F32) {
        if (t) natural_missing_kernel<float, false><<<(unsigned)blocks, threads, 0, s>>>((const float*)x, (const float*)t, (float*)coeffs, (float*)scratch, n_series, L, C, version);
        else natural_missing_kernel<float, true><<<(unsigned)blocks, threads, 0, s>>>((const float*)x, nullptr, (float*)coeffs, (float*)scratch, n_series, L, C, version);
    } else {
        if (t) natural_missing_kernel<double, false><<<(unsigned)blocks, threads, 0, s>>>((const double*)x, (const double*)t, (double*)coeffs, (double*)scratch, n_series, L, C, version);
        else natural_missing_kernel<double, true><<<(unsigned)blocks, threads, 0, s>>>((const double*)x, nullptr, (double*)coeffs, (double*)scratch, n_series, L, C, version);
    }
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

extern "C" int tcde_nan_flag(const void* x, int64_t n, int dtype, int32_t* flags, void* stream) {
    TCDE_CHECK_ARG(x != nullptr && flags != nullptr && n >= 0, "null pointer or negative size");
    TCDE_CHECK_ARG(dtype == TCDE_F32 || dtype == TCDE_F64, "dtype=%d", dtype);
    if (n == 0) return TCDE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t blocks = (n + kThreads * 8 - 1) / (kThreads * 8);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == TCDE_F32) nan_flag_kernel<float><<<(unsigned)blocks, kThreads, 0, s>>>((const float*)x, n, flags);
    else nan_flag_kernel<double><<<(unsigned)blocks, kThreads, 0, s>>>((const double*)x, n, flags);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}

extern "C" int tcde_set_natural_variant(int variant) {
    TCDE_CHECK_ARG(variant >= 0 && variant <= 2,
                   "variant=%d (0 parallel kernels, 1 one thread per series, 2 CTA-per-path natural kernel)", variant);
    g_natural_variant = variant;
    g_fill_variant = variant;
    return TCDE_OK;
}
