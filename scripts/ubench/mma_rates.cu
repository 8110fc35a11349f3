// Micro-benchmarks that size the round-2 solve kernel (B200, sm_100a).  Stand-alone: nvcc -> binary, run under gpurun.
//   1. tcgen05.mma issue-to-completion rate: kind::tf32 (K=8) vs kind::f16 (K=16), M=128, N in {256,128,64}, one or two accumulators
//   2. tcgen05.ld bandwidth: 4 / 8 warps reading 256 columns, alone and under a running MMA stream
//   3. FP32 side: FFMA2, cvt.rna.tf32, cvt.rn.f16x2 per-SMSP rates at 1/2/3 warps per SMSP
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../torchcde_b200/csrc/umma.cuh"

using namespace tcde::umma;

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode bit0: 0 = tf32, 1 = f16; n = N; accs = accumulators alternated; reads = concurrent TMEM readers (0/4/8 warps)
__global__ void __launch_bounds__(32 * 9, 1) mma_kernel(int kind, int n, int accs, int rounds, int per_round, int readers, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* a_tile = reinterpret_cast<float*>(smem);                 // 128 rows x 128 B
    float* b_tile = reinterpret_cast<float*>(smem + 16384);         // 256 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (16384 + 32768) / 4; e += blockDim.x) reinterpret_cast<float*>(smem)[e] = 0.f;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); *stop = 0; }
    if (warp == 8) tmem_alloc(slot, 512);
    tcde::fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0, r0 = 0, r1 = 0;
    long long nread = 0;
    if (warp == 8) {
        const uint32_t fmt = kind ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(a_tile), db = make_desc(b_tile);
        uint32_t phase = 0;
        if ((tid & 31) == 0) {
            t0 = clock64();
            for (int r = 0; r < rounds; ++r) {
                for (int i = 0; i < per_round; ++i) {
                    const uint32_t d = tmem + (uint32_t)((i % accs) * 256);
                    if (kind) mma_f16(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
                    else mma_tf32(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
                }
                mma_commit(bar);
                mbar_wait(bar, phase);
                phase ^= 1;
            }
            t1 = clock64();
            *stop = 1;
            out[blockIdx.x * 8 + 0] = t1 - t0;
        }
        __syncwarp();
    } else if (warp < readers) {
        // TMEM readers: warp w reads lanes 32*(w%4).., all 256 columns of accumulator (w/4)
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
        uint32_t va[16], vb[16];
        uint32_t sink = 0;
        r0 = clock64();
        while (!*stop || nread < 64) {
            tmem_ld16_issue(taddr, va);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t* cur = (j & 1) ? vb : va;
                tmem_ld16_wait(cur);
                if (j + 1 < 16) tmem_ld16_issue(taddr + 16u * (j + 1), (j & 1) ? va : vb);
#pragma unroll
                for (int q = 0; q < 16; ++q) sink ^= cur[q];
            }
            ++nread;
            if (rounds == 0 && nread >= 256) break;
        }
        r1 = clock64();
        if ((tid & 31) == 0) {
            out[blockIdx.x * 8 + 1 + (warp & 3)] = (r1 - r0) / (nread ? nread : 1);
            if (sink == 0x12345) out[0] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// FP32-side issue rates: op 0 = FFMA2 (8 independent chains), 1 = cvt.rna.tf32, 2 = cvt.rn.f16x2.f32 + back, 3 = FFMA, 4 = FMNMX
__global__ void fp32_kernel(int op, int iters, long long* out, float* sink_out) {
    f2 acc[8];
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = pk(1.f + threadIdx.x + i, 2.f + i); x[i] = 1.f + 0.001f * (threadIdx.x + i); }
    const f2 m = pk(1.0000001f, 0.9999999f), c = pk(1e-7f, -1e-7f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (op == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], m, c);
        } else if (op == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = tf32_hi(x[i]) + 1e-3f;
        } else if (op == 2) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                uint32_t h;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[i + 1]), "f"(x[i]));
                float lo, hi;
                asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h));
                x[i] = lo + 1e-3f;
                x[i + 1] = hi + 1e-3f;
            }
        } else if (op == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], 1.0000001f, 1e-7f);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], x[(i + 1) & 7] * 0.5f);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; upk(acc[i], a, b); s += a + b + x[i]; }
    if (s == 12345.678f) sink_out[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    long long* out;
    float* sink;
    cudaMalloc(&out, 148 * 8 * sizeof(long long));
    cudaMalloc(&sink, 16);
    long long host[148 * 8];
    const int smem = 16384 + 32768 + 1024 + 64;
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("# tcgen05.mma M=128: cycles per MMA (issue + completion of a 64-MMA round, 16 rounds), all 148 SMs\n");
    printf("kind n accs readers cyc_per_mma  tmem_read_cyc_per_256col_row(warps 0-3)\n");
    for (int kind = 0; kind < 2; ++kind)
        for (int n : {256, 128, 64, 32})
            for (int accs : {1, 2})
                for (int readers : {0, 4, 8}) {
                    if (n != 256 && (accs == 2 || readers)) continue;
                    cudaMemset(out, 0, 148 * 8 * sizeof(long long));
                    mma_kernel<<<148, 32 * 9, smem>>>(kind, n, accs, 16, 64, readers, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
                    printf("%s %3d %d %d  %.1f   %lld %lld %lld %lld\n", kind ? "f16 " : "tf32", n, accs, readers, host[0] / (16.0 * 64), host[1], host[2],
                           host[3], host[4]);
                }
    printf("# tcgen05.ld alone (no MMA): cycles per 256-column row read per warp\n");
    for (int readers : {4, 8}) {
        cudaMemset(out, 0, 148 * 8 * sizeof(long long));
        mma_kernel<<<148, 32 * 9, smem>>>(0, 256, 1, 0, 0, readers, out);
        cudaDeviceSynchronize();
        cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
        printf("readers %d: %lld %lld %lld %lld\n", readers, host[1], host[2], host[3], host[4]);
    }
    printf("# FP32 side: cycles per warp-instruction-group of 8 (divide by 8 for one instruction), by warps per SMSP\n");
    const char* names[] = {"FFMA2", "cvt.rna.tf32+FADD", "cvt.f16x2+2cvt.f32+2FADD per pair", "FFMA", "FMNMX+FMUL"};
    for (int op = 0; op < 5; ++op)
        for (int warps : {4, 8, 16}) {
            fp32_kernel<<<148, warps * 32>>>(op, 4096, out, sink);
            cudaDeviceSynchronize();
            cudaMemcpy(host, out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            printf("%-36s warps/SM %2d: %.2f cycles per 8-op group per warp => %.2f cyc/instr/SMSP\n", names[op], warps, host[0] / 4096.0,
                   host[0] / 4096.0 / 8.0 / (warps / 4));
        }
    return 0;
}
