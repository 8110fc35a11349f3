// Log-ODE preprocessing: the logsignature of every window of every path in ONE launch.
//
// The reference (torchcde/log_ode.py:56-68) calls ``signatory.Logsignature(depth)`` -- an optional third-party C++ package
// that is not vendored and not installable here -- once per window in a Python loop.  This kernel restates the published
// algorithm that class implements for a piecewise-linear path (Chen's identity on truncated tensor exponentials, the
// logarithm in the truncated tensor algebra, coefficients of the Lyndon words = signatory's default "words" basis):
//   thread = (path, window); the signature S of the window is built increment by increment,
//       S <- S (x) exp(dx),   exp(dx)_j = dx^(x)j / j!,
//   then log(1 + x) = x (x) (1 - x (x) (1/2 - x (x) (1/3 - ...))) (Horner; powers of one element commute), and the entries
//   at the Lyndon words (level, flat index: a table built on the host) are written out.
// The truncated tensor algebra of a thread lives in local memory: levels 1..depth of size C^k, at most kMaxSig floats
// (C = 8 at depth 3 is 584; C = 3 at depth 4 is 120) -- preprocessing, run once per dataset, not a roofline kernel.
#include "common.cuh"

namespace tcde {
namespace logsig {

constexpr int kMaxSig = 640;
constexpr int kMaxDepth = 6;

struct Levels {
    int depth, channels;
    int off[kMaxDepth + 2];      // off[k] = start of level k (k = 1..depth) in the flat buffer; off[depth + 1] = total
    int size[kMaxDepth + 1];     // size[k] = channels^k
};

// out_k += sum_{j = 1..k-1} a_j (x) b_{k-j}   (levels 1..k-1 of both operands)
template <typename T>
__device__ __forceinline__ void add_cross_terms(const Levels& lv, int k, const T* a, const T* b, T* out_k) {
    for (int j = 1; j < k; ++j) {
        const T* aj = a + lv.off[j];
        const T* bl = b + lv.off[k - j];
        const int nb = lv.size[k - j];
        for (int u = 0; u < lv.size[j]; ++u) {
            const T au = aj[u];
            T* dst = out_k + (size_t)u * nb;
            for (int v = 0; v < nb; ++v) dst[v] = fma(au, bl[v], dst[v]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
logsig_windows_kernel(const T* __restrict__ x, int64_t n_paths, int64_t length, const int32_t* __restrict__ window_index,
                      int n_windows, const Levels lv, const int32_t* __restrict__ words, int n_words, T* __restrict__ out) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_paths * n_windows) return;
    const int64_t path = item / n_windows;
    const int w = (int)(item - path * n_windows);
    const int C = lv.channels, depth = lv.depth, total = lv.off[depth + 1];
    T S[kMaxSig], E[kMaxSig], R[kMaxSig];
    for (int i = 0; i < total; ++i) S[i] = T(0);
    const T* px = x + path * length * C;
    const int first = window_index[w], last = window_index[w + 1];
    T dx[16];
    for (int p = first; p < last; ++p) {
        for (int c = 0; c < C; ++c) dx[c] = px[(int64_t)(p + 1) * C + c] - px[(int64_t)p * C + c];
        // E = exp(dx) (levels 1..depth)
        for (int c = 0; c < C; ++c) E[lv.off[1] + c] = dx[c];
        for (int j = 2; j <= depth; ++j) {
            const T inv = T(1) / T(j);
            const T* prev = E + lv.off[j - 1];
            T* cur = E + lv.off[j];
            for (int u = 0; u < lv.size[j - 1]; ++u) {
                const T pu = prev[u] * inv;
                for (int c = 0; c < C; ++c) cur[(size_t)u * C + c] = pu * dx[c];
            }
        }
        // S <- S (x) E, highest level first (a level only needs the lower levels of the OLD S)
        for (int k = depth; k >= 1; --k) {
            T* sk = S + lv.off[k];
            const T* ek = E + lv.off[k];
            for (int i = 0; i < lv.size[k]; ++i) sk[i] += ek[i];
            add_cross_terms(lv, k, S, E, sk);
        }
    }
    // log(1 + S): R <- 1/n - S (x) R from n = depth down to 1, then L = S (x) R (R has a level 0: r0)
    T r0 = T(1) / T(depth);
    for (int i = 0; i < total; ++i) R[i] = T(0);
    for (int n = depth - 1; n >= 1; --n) {
        // E <- -(S (x) R) restricted to levels 1..depth, with R = (r0, R_1..)
        for (int k = 1; k <= depth; ++k) {
            T* ek = E + lv.off[k];
            const T* sk = S + lv.off[k];
            for (int i = 0; i < lv.size[k]; ++i) ek[i] = sk[i] * r0;
        }
        for (int k = depth; k >= 2; --k) add_cross_terms(lv, k, S, R, E + lv.off[k]);
        for (int i = 0; i < total; ++i) R[i] = -E[i];
        r0 = T(1) / T(n);
    }
    for (int k = 1; k <= depth; ++k) {
        T* ek = E + lv.off[k];
        const T* sk = S + lv.off[k];
        for (int i = 0; i < lv.size[k]; ++i) ek[i] = sk[i] * r0;
    }
    for (int k = depth; k >= 2; --k) add_cross_terms(lv, k, S, R, E + lv.off[k]);
    T* po = out + item * n_words;
    for (int q = 0; q < n_words; ++q) po[q] = E[lv.off[words[2 * q]] + words[2 * q + 1]];
}

}  // namespace logsig
}  // namespace tcde

using namespace tcde;

extern "C" int64_t tcde_logsignature_max_terms(void) { return logsig::kMaxSig; }

extern "C" int tcde_logsignature_windows(const void* x, int64_t n_paths, int64_t length, int64_t channels,
                                         const int32_t* window_index, int64_t n_windows, int depth, const int32_t* words,
                                         int64_t n_words, void* out, int dtype, void* stream) {
    TCDE_CHECK_ARG(x && window_index && words && out, "null pointer");
    TCDE_CHECK_ARG(n_paths >= 0 && length >= 2 && channels >= 1 && n_windows >= 1 && n_words >= 1, "bad sizes");
    TCDE_CHECK_ARG(depth >= 1, "depth=%d", depth);
    TCDE_CHECK_ARG(dtype == TCDE_F32 || dtype == TCDE_F64, "dtype=%d", dtype);
    TCDE_CHECK_SUPPORTED(depth <= logsig::kMaxDepth && channels <= 16, "logsignature: depth <= %d and channels <= 16 (got %d, %lld)",
                         logsig::kMaxDepth, depth, (long long)channels);
    logsig::Levels lv;
    lv.depth = depth;
    lv.channels = (int)channels;
    int64_t total = 0, pw = 1;
    lv.off[0] = 0;
    lv.size[0] = 1;
    for (int k = 1; k <= depth; ++k) {
        pw *= channels;
        TCDE_CHECK_SUPPORTED(total + pw <= logsig::kMaxSig,
                             "logsignature: channels^1 + ... + channels^depth = more than %d terms (channels=%lld, depth=%d)",
                             logsig::kMaxSig, (long long)channels, depth);
        lv.off[k] = (int)total;
        lv.size[k] = (int)pw;
        total += pw;
    }
    lv.off[depth + 1] = (int)total;
    if (n_paths == 0) return TCDE_OK;
    const int64_t items = n_paths * n_windows;
    const int threads = 128;
    const int64_t blocks = (items + threads - 1) / threads;
    TCDE_CHECK_SUPPORTED(blocks < (1ll << 31), "too many (path, window) pairs");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TCDE_F32)
        logsig::logsig_windows_kernel<float><<<(unsigned)blocks, threads, 0, s>>>((const float*)x, n_paths, length, window_index,
                                                                                 (int)n_windows, lv, words, (int)n_words, (float*)out);
    else
        logsig::logsig_windows_kernel<double><<<(unsigned)blocks, threads, 0, s>>>((const double*)x, n_paths, length, window_index,
                                                                                  (int)n_windows, lv, words, (int)n_words, (double*)out);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}
