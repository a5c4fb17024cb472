// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M = 128, K = 8, both operands from shared memory) as a function
// of N and of the operand layout.  One CTA per SM, one issuing thread, `reps` back-to-back MMAs into one accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma tools/ubench_mma.cu && ./ubench_mma
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>
#include "../zerovox_b200/csrc/tc_ptx.cuh"

using namespace zvx;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo16 & 0x3FFFu) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t sbo_bytes, uint64_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

// mode 2: 32-byte swizzle (rows of 32 B = one K = 8 step, 8-row atoms of 256 B); mode 3: 64-byte swizzle (rows of 64 B, 512 B atoms)
// mode 0: no-swizzle, every MMA reads a different row offset of A (the vocoder pattern); 1: 128B swizzle (GEMM pattern)
__device__ int g_lbo_a = 1024, g_lbo_b = 0;   // no-swizzle mode: K-half distance of A / B in 16-byte rows (0: B uses N)
__global__ void __launch_bounds__(128) bench(int N, int mode, int reps, int same_a, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar = sb + 200 * 1024, slot = bar + 8;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 8);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t sA = sb, sW = sb + 128 * 1024;
        const int Rp = g_lbo_a;
        // eight operand descriptor pairs built BEFORE the timed loop: the loop body is eight back-to-back tcgen05.mma with
        // register operands, so the figure is the tensor pipe's, not the issuing thread's address arithmetic
        uint64_t da[8], db[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (mode == 0) {
                const int row = same_a ? 0 : (r * 37) % 512;
                da[r] = desc_nosw(sA + row * 16, Rp);
                db[r] = desc_nosw(sW, g_lbo_b ? g_lbo_b : N);
            } else if (mode == 2) {
                const int row = same_a ? 0 : (r * 37) % 512;
                da[r] = desc_sw(sA + row * 32, 256, 6);
                db[r] = desc_sw(sW, 256, 6);
            } else if (mode == 3) {
                const int row = same_a ? 0 : (r * 37) % 512;
                da[r] = desc_sw(sA + row * 64, 512, 4) + (uint64_t)(2 * (r & 1));
                db[r] = desc_sw(sW, 512, 4) + (uint64_t)(2 * (r & 1));
            } else {
                da[r] = desc_sw128(sA + (same_a ? 0 : ((r & 3) * 16384))) + (uint64_t)(2 * (r & 3));
                db[r] = desc_sw128(sW) + (uint64_t)(2 * (r & 3));
            }
        }
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) umma_tf32(tmem, da[u], db[u], idesc, (r | u) ? 1u : 0u);
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// --peak: the tensor pipe's TF32 ceiling as a throughput — every SM issues N = 256, 128B-swizzled MMAs back to back from
// resident shared-memory operands (no loads), timed with CUDA events: best of 10 launches (burst) and launches back to back
// for 4 s (sustained, power-limited clocks).  One JSON line.
int peak_mode() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int reps = 100000, N = 256;
    const double flop = 2.0 * 128 * N * 8 * (double)reps * sms;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int it = 0; it < 12; ++it) {
        cudaEventRecord(a);
        bench<<<sms, 128, 202 * 1024 + 1024>>>(N, 1, reps, 0, d);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it >= 2 && ms < best) best = ms;
    }
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    int launches = 0;
    float total = 0.f;
    cudaEventRecord(a);
    while (total < 4000.f) {
        for (int i = 0; i < 20; ++i) bench<<<sms, 128, 202 * 1024 + 1024>>>(N, 1, reps, 0, d);
        launches += 20;
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&total, a, b);
    }
    printf("{\"tcgen05_tf32_issue_tflops\": %.1f, \"tcgen05_tf32_issue_tflops_sustained\": %.1f, \"cycles_per_mma_n256\": %.2f, "
           "\"sms\": %d, \"flop_per_clk_per_sm\": %.0f}\n",
           flop / (best * 1e-3) / 1e12, flop * launches / (total * 1e-3) / 1e12, (double)cyc / reps, sms,
           2.0 * 128 * N * 8 / ((double)cyc / reps));
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--peak") return peak_mode();
    if (argc > 1 && std::string(argv[1]) == "--swz") {
        long long* d;
        cudaMalloc(&d, 8);
        cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const int reps = 4096;
        for (int mode : {2, 3})
            for (int same = 0; same < 2; ++same)
                for (int N : {16, 32, 64, 96, 128, 256}) {
                    long long best = 1LL << 60;
                    for (int it = 0; it < 3; ++it) {
                        bench<<<148, 128, 202 * 1024 + 1024>>>(N, mode, reps, same, d);
                        long long h = 0;
                        if (cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
                        if (h < best) best = h;
                    }
                    printf("layout=%s a=%s N=%3d : %.1f cycles/MMA\n", mode == 2 ? "sw32" : "sw64", same ? "same" : "moving (any row)", N,
                           (double)best / reps);
                }
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "--lbo") {
        // no-swizzle K-major operands: does the distance between the two K-halves (bank alignment) change the MMA cost?
        long long* d;
        cudaMalloc(&d, 8);
        cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        const int reps = 4096;
        for (int lboa : {1024, 1025, 1026, 1028, 1032, 1040, 128, 132, 2324, 1162, 585}) {
            for (int lbob : {0, 40, 104, 132, 129, 352}) {
                for (int N : {64, 128}) {
                    cudaMemcpyToSymbol(g_lbo_a, &lboa, 4);
                    cudaMemcpyToSymbol(g_lbo_b, &lbob, 4);
                    long long best = 1LL << 60;
                    for (int it = 0; it < 3; ++it) {
                        bench<<<148, 128, 202 * 1024 + 1024>>>(N, 0, reps, 0, d);
                        long long h = 0;
                        if (cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
                        if (h < best) best = h;
                    }
                    printf("nosw lbo_a=%4d lbo_b=%4d N=%3d : %.1f cycles/MMA\n", lboa, lbob ? lbob : N, N, (double)best / reps);
                }
            }
        }
        return 0;
    }
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int reps = 4096;
    printf("tcgen05.mma kind::tf32 M=128 K=8, SS operands, %d back-to-back MMAs, cycles per MMA (148 CTAs running)\n", reps);
    for (int mode = 0; mode < 2; ++mode)
        for (int same = 0; same < 2; ++same)
            for (int N : {16, 32, 64, 128, 176, 256}) {
                long long best = 1LL << 60;
                for (int it = 0; it < 3; ++it) {
                    bench<<<148, 128, 202 * 1024 + 1024>>>(N, mode, reps, same, d);
                    long long h = 0;
                    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    if (h < best) best = h;
                }
                printf("layout=%s a=%s N=%3d : %.1f cycles/MMA  (math floor N/2 = %d)\n", mode ? "sw128" : "nosw", same ? "same" : "moving",
                       N, (double)best / reps, N / 2);
            }
    return 0;
}
