// Shared device/host helpers for the torchcde_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/torchcde_b200.h"

namespace tcde {

// ---- error plumbing (nothing throws across the C ABI) ------------------------------------
void set_error(const char* fmt, ...);
int cuda_failed(cudaError_t e, const char* what);

#define TCDE_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            tcde::set_error(__VA_ARGS__);         \
            return TCDE_ERR_ARGUMENT;             \
        }                                         \
    } while (0)

#define TCDE_CHECK_SUPPORTED(cond, ...)           \
    do {                                          \
        if (!(cond)) {                            \
            tcde::set_error(__VA_ARGS__);         \
            return TCDE_ERR_UNSUPPORTED;          \
        }                                         \
    } while (0)

#define TCDE_CHECK_CUDA(call)                                         \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return tcde::cuda_failed(e_, #call);   \
    } while (0)

int sm_count();

// ---- arithmetic that must round exactly once per operation --------------------------------
// The reference is a chain of separate torch elementwise ops, i.e. one IEEE rounding per
// +,-,*,/ and never a fused multiply-add.  Kernels that promise bit-identical results use
// these wrappers so that nvcc cannot contract a*b+c into an FMA.
template <typename T> struct exact;
template <> struct exact<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct exact<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

template <typename T> __device__ __forceinline__ bool is_nan(T v) { return v != v; }

// ---- 1-D bulk (TMA) shared -> global store -------------------------------------------------
// cp.async.bulk moves a contiguous, 16-byte-aligned byte range with one instruction issued by
// one thread; SASS shows it as UBLKCP.  The async proxy must see the generic-proxy writes to
// shared memory first, hence the proxy fence before the CTA barrier that precedes the issue.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem, const void* smem, uint32_t bytes) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---- 16-byte cp.async (LDGSTS) global -> shared --------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- arguments of the tensor-core solve (solve_umma.cu), filled in by the C entry points of solve_simt.cu ----
struct UmmaArgs {
    const float* control;
    const float* weight;
    const float* bias;
    const float* z0;
    float* out;
    const float* step_dt;
    const int32_t* stage_index;
    const float* stage_frac;
    const int32_t* out_step;
    const int32_t* out_mode;
    const float* out_slope;
    int64_t n_paths;
    int64_t n_rows;
    int control_kind, method, n_stages, n_steps, n_out;
    float sign;
    long long* trace;     // optional [64][8] clock64 stamps of CTA 0 / tile 0 (profiling aid), else nullptr
    float* stage_dump;    // optional [n_steps * n_stages][n_paths][32]: the input of every stage (for the adjoint)
    int debug;            // solve_tc.cu: profiling experiments with WRONG results (1 no row fetch, 4 no TMEM reads,
                          // 8 no split) and switches with right results (2 proxy fence in every row thread too, 16 one segment, 32 force four segments, 64 no stagger)
};
bool solve_umma_supported(int H, int C);
int solve_umma_f32(const UmmaArgs& a, int H, int C, cudaStream_t stream);
// round-2 kernel (solve_tc.cu): Runge-Kutta state in registers; mode 0 = 3xTF32 (13 MMAs per stage), mode 1 = 2xFP16 with
// per-path power-of-two scaling (7 MMAs per stage)
int solve_tc_f32(const UmmaArgs& a, int H, int C, int mode, cudaStream_t stream);
int current_solve_variant();          // tcde_set_solve_variant: 0 auto, 1 CUDA-core, 2 tcgen05 round 1, 3 tcgen05 TF32 r2, 4 tcgen05 FP16 r2

// parameter gradients of a whole backward solve on the tensor cores (param_grad_umma.cu)
int param_grad_umma_grid(int64_t n_paths, int64_t n_stage_total);
int param_grad_umma_f32(const float* control, int control_kind, int64_t n_rows, const float* z_stages,
                        const float* a_stages, const int32_t* stage_index, const float* stage_frac,
                        const float* stage_weight, int n_stage_total, float* scratch, int64_t n_paths, int grid,
                        cudaStream_t stream);

// round 2: the same product with U on the N side (6 MMAs of N = 256 per 32 pairs instead of 24 of N = 48), BF16 two-way split,
// MN-major operands, TMA-fed (param_grad_bf16.cu)
int param_grad_bf16_grid(int64_t n_paths, int64_t n_stage_total);
int param_grad_bf16_f32(const float* control, int control_kind, int64_t n_rows, const float* z_stages, const float* a_stages,
                        const int32_t* stage_index, const float* stage_frac, const float* stage_weight, int n_stage_total,
                        float* scratch, int64_t n_paths, int grid, cudaStream_t stream);

}  // namespace tcde
