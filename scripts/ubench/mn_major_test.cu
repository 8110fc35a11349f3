// Stand-alone check of the MN-major, 128B-swizzled BF16 operand layout used by param_grad_bf16.cu:
//   D[m][n] = sum_k A[k][m] * B[k][n],   A: K x 128 (m contiguous), B: K x 256 (n contiguous), K = 32, bf16 -> fp32 in TMEM
// against a host reference.  Canonical layout (cute mma_traits_sm100.hpp, Major::MN, SWIZZLE_128B, in 16-byte units):
//   ((8, n), (8, k)) : ((1, LBO), (8, SBO))  -> atom = 8 k-rows of 128 bytes (64 elements), chunk c of row r at c ^ r.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../torchcde_b200/csrc/umma.cuh"

using namespace tcde::umma;

constexpr int K = 32, M = 128, N = 256;

__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
// MN-major SW128 descriptor: LBO = bytes between 64-element MN blocks, SBO = bytes between groups of 8 k
__device__ __forceinline__ uint64_t desc_mn(const void* tile, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_u32(tile) & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of element (k, mn) in a tile of `blocks` MN blocks of 64 and K = 32 (4 k-groups): block-major, then k-group
__host__ __device__ inline uint32_t off_mn(int k, int mn) {
    const int blk = mn >> 6, within = mn & 63, chunk = within >> 3, e = within & 7, g = k >> 3, row = k & 7;
    return (uint32_t)(blk * (K / 8) * 1024 + g * 1024 + row * 128 + ((chunk ^ row) << 4) + e * 2);
}

__global__ void test_kernel(const __nv_bfloat16* a_in, const __nv_bfloat16* b_in, float* d_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_tile = smem;                    // 2 blocks x 4 k-groups x 1 KB = 8 KB
    unsigned char* b_tile = smem + 8192;             // 4 blocks x 4 KB = 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < K * M; e += blockDim.x) *reinterpret_cast<__nv_bfloat16*>(a_tile + off_mn(e / M, e % M)) = a_in[e];
    for (int e = tid; e < K * N; e += blockDim.x) *reinterpret_cast<__nv_bfloat16*>(b_tile + off_mn(e / N, e % N)) = b_in[e];
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 256);
    tcde::fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        // idesc: D = F32 (bit 4), A = B = BF16 (1 at bits 7, 10), A and B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < K / 16; ++kb) {        // one MMA = 16 k = 2 k-groups = 2 KB further into every MN block
            const uint64_t da = desc_mn(a_tile + kb * 2048, (K / 8) * 1024, 1024);
            const uint64_t db = desc_mn(b_tile + kb * 2048, (K / 8) * 1024, 1024);
            mma_bf16(tmem, da, db, idesc, kb > 0);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (warp < 4) {                                   // thread = row m of D
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t v[16];
        for (int q = 0; q < N / 16; ++q) {
            tmem_ld16_issue(taddr + 16 * q, v);
            tmem_ld16_wait(v);
            for (int i = 0; i < 16; ++i) d_out[(size_t)tid * N + 16 * q + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    __nv_bfloat16 *ha = new __nv_bfloat16[K * M], *hb = new __nv_bfloat16[K * N];
    float* ref = new float[M * N]();
    srand(1);
    for (int i = 0; i < K * M; ++i) ha[i] = __float2bfloat16((rand() % 17 - 8) * 0.125f);
    for (int i = 0; i < K * N; ++i) hb[i] = __float2bfloat16((rand() % 13 - 6) * 0.25f);
    for (int k = 0; k < K; ++k)
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) ref[m * N + n] += __bfloat162float(ha[k * M + m]) * __bfloat162float(hb[k * N + n]);
    __nv_bfloat16 *da, *db;
    float* dd;
    cudaMalloc(&da, K * M * 2); cudaMalloc(&db, K * N * 2); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, ha, K * M * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb, K * N * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, M * N * 4);
    const int smem = 8192 + 16384 + 64 + 1024;
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    test_kernel<<<1, 128, smem>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    float* got = new float[M * N];
    cudaMemcpy(got, dd, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
        const double d = fabs(got[i] - ref[i]);
        if (d > worst) worst = d;
        if (d > 1e-3) ++bad;
    }
    printf("MN-major bf16 UMMA test: max |D - ref| = %.3e, mismatches %d of %d  => %s\n", worst, bad, M * N, bad ? "LAYOUT WRONG" : "ok");
    printf("sample D[0][0..3] = %g %g %g %g   ref %g %g %g %g\n", got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3]);
    printf("sample D[33][64..66] = %g %g %g   ref %g %g %g\n", got[33 * N + 64], got[33 * N + 65], got[33 * N + 66], ref[33 * N + 64], ref[33 * N + 65], ref[33 * N + 66]);
    return bad ? 2 : 0;
}
