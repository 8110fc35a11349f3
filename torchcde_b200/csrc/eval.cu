// CubicSpline / LinearInterpolation evaluate and derivative at many query times
// (interpolation_cubic.py:324-336, interpolation_linear.py:212-225), one rounding per op so
// that the results are bit-identical to the reference's chain of torch ops.  The interval
// index and fraction of every query come from the host (torchcde_b200/schedule.py runs the
// reference's own bucketize arithmetic, which is what makes the indices bit-exact).
#include "common.cuh"

namespace tcde {

template <typename T>
__global__ void __launch_bounds__(256)
spline_eval_kernel(const T* __restrict__ control, const T* __restrict__ knot_t, const int32_t* __restrict__ index,
                   const T* __restrict__ frac, T* __restrict__ out, int64_t total, int64_t n_rows, int C,
                   int64_t n_times, int cubic, int derivative) {
    using E = exact<T>;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int64_t pj = e / C;
    const int64_t j = pj % n_times;
    const int64_t p = pj / n_times;
    const int idx = index[j];
    const T f = frac[j];
    T v;
    if (cubic) {
        const T* r = control + (p * n_rows + idx) * 4 * C + c;
        const T b = r[C], two_c = r[2 * C], three_d = r[3 * C];
        if (derivative) {
            v = E::add(b, E::mul(E::add(two_c, E::mul(three_d, f)), f));
        } else {
            T inner = E::add(E::mul(T(0.5), two_c), E::div(E::mul(three_d, f), T(3)));
            inner = E::add(b, E::mul(inner, f));
            v = E::add(r[0], E::mul(inner, f));
        }
    } else {
        const T* r = control + (p * n_rows + idx) * C + c;
        const T lo = r[0], hi = r[C];
        const T t0 = knot_t ? knot_t[idx] : T(idx), t1 = knot_t ? knot_t[idx + 1] : T(idx + 1);
        const T width = E::sub(t1, t0);
        if (derivative) v = E::div(E::sub(hi, lo), width);
        else v = E::add(lo, E::div(E::mul(f, E::sub(hi, lo)), width));
    }
    out[e] = v;
}

}  // namespace tcde

using namespace tcde;

extern "C" int tcde_spline_eval(const void* control, const void* knot_t, const int32_t* index, const void* frac,
                                void* out, int64_t n_paths, int64_t n_rows, int64_t channels, int64_t n_times,
                                int control_kind, int derivative, int dtype, void* stream) {
    TCDE_CHECK_ARG(control && index && frac && out, "null data pointer");
    TCDE_CHECK_ARG(n_paths >= 0 && n_rows >= 1 && channels >= 1 && n_times >= 0, "bad sizes");
    TCDE_CHECK_ARG(dtype == TCDE_F32 || dtype == TCDE_F64, "dtype=%d", dtype);
    TCDE_CHECK_ARG(control_kind == TCDE_CONTROL_CUBIC || control_kind == TCDE_CONTROL_LINEAR, "control_kind=%d",
                   control_kind);
    const int64_t total = n_paths * n_times * channels;
    if (total == 0) return TCDE_OK;
    const int64_t blocks = (total + 255) / 256;
    TCDE_CHECK_SUPPORTED(blocks < (1ll << 31), "too many outputs");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int cubic = control_kind == TCDE_CONTROL_CUBIC;
    if (dtype == TCDE_F32)
        spline_eval_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)control, (const float*)knot_t, index,
                                                                    (const float*)frac, (float*)out, total, n_rows,
                                                                    (int)channels, n_times, cubic, derivative);
    else
        spline_eval_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const double*)control, (const double*)knot_t,
                                                                     index, (const double*)frac, (double*)out, total,
                                                                     n_rows, (int)channels, n_times, cubic, derivative);
    TCDE_CHECK_CUDA(cudaGetLastError());
    return TCDE_OK;
}
