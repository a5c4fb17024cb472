// Shared helpers for the zerovox_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <cstdio>
#include <cstdlib>

namespace zvx {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define ZVX_CUDA_CHECK(expr)                                                                 \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _buf[512];                                                                  \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr,                     \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                            \
            throw ::zvx::Error(_buf);                                                        \
        }                                                                                    \
    } while (0)

#define ZVX_REQUIRE(cond, msg)                                                               \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            char _buf[512];                                                                  \
            snprintf(_buf, sizeof(_buf), "%s [%s] (%s:%d)", msg, #cond, __FILE__, __LINE__); \
            throw ::zvx::Error(_buf);                                                        \
        }                                                                                    \
    } while (0)

// Tuning / ablation switches (ZVX_* environment variables) exist only in -DZVX_DEBUG builds (ZVX_BUILD_DEBUG=1 python
// __graft_entry__.py; tools/gpu_ab_*.sh): the release library has the measured defaults baked in and reads no environment.
#ifdef ZVX_DEBUG
static inline int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
static inline long long env_ll(const char* name, long long dflt) { const char* v = getenv(name); return v ? atoll(v) : dflt; }
static inline bool env_set(const char* name) { return getenv(name) != nullptr; }
#define ZVX_DBG_PTR(p) ((p).dbg)
#define ZVX_DBG_SKIP(p) ((p).dbg_skip)
#else
static inline constexpr int env_int(const char*, int dflt) { return dflt; }
static inline constexpr long long env_ll(const char*, long long dflt) { return dflt; }
static inline constexpr bool env_set(const char*) { return false; }
#define ZVX_DBG_PTR(p) (static_cast<long long*>(nullptr))
#define ZVX_DBG_SKIP(p) (0)
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// Launch counter (reported through zvx_launch_count; bench.py's gpu_launches).
extern long long g_launches;
#define ZVX_LAUNCHED() (++::zvx::g_launches)

#define ZVX_POST_LAUNCH()                                                                    \
    do {                                                                                     \
        ZVX_LAUNCHED();                                                                      \
        ZVX_CUDA_CHECK(cudaGetLastError());                                                  \
    } while (0)

// Programmatic dependent launch (sm_90+), opt-in.  The hot kernels are launched through launch_k — with the stream-serialisation
// attribute when the option is on.  Each calls pdl_trigger() (first thing; gemm_tc when a CTA starts its last tile): its successor
// in the stream may then be scheduled as soon as every CTA of this grid has got that far or is done, i.e. on the SMs the grid's last
// wave leaves idle.  And each calls pdl_wait() after its own set-up (barrier init, TMEM allocation, shared-memory clearing,
// descriptor prefetch: nothing that touches global memory), which returns when the predecessor grid has COMPLETED and its writes
// are visible.  So the data dependences are exactly those of plain stream order; what overlaps is the launch latency and the
// prologue of kernel N+1 with the tail of kernel N.  Without the attribute both instructions are no-ops — the default
// (zvx_set_option("pdl", 0)): measured on configs[1], -0.6 % with every kernel triggering first thing, +0.3 % with gemm_tc
// triggering at its last tile (profiles/r02_ab_pdl.txt).
extern int g_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1u : 0u;
    ZVX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// GEMM / implicit-GEMM convolution arguments shared by the fp32 SIMT kernel (gemm_simt.cu) and the
// tcgen05 TF32 kernel (gemm_tc.cu).
//
//   C[z][m, n] = epilogue( sum_tap sum_k A[z][rowmap(m, tap), k] * W[z][tap][n, k] )
//
// rowmap modes
//   ROW_PLAIN : rowmap(m, 0) = m
//   ROW_CONV1D: m = s*Lout + t ; ti = t*stride + tap*dil - pad ; valid iff 0 <= ti < Lin ; row = s*Lin + ti
//   ROW_CONV2D: m = (img*Ho + y)*Wo + x ; tap = dy*ksize + dx ; yi = y*stride + dy - pad, xi likewise;
//               valid iff inside [0,Hi)x[0,Wi) ; row = (img*Hi + yi)*Wi + xi
// invalid rows contribute zero (the convolutions' zero padding).
//
// epilogue: v = acc + bias[n]; if relu_first v = max(v,0); if scale v = v*scale[n] + shift[n];
//           if R v += R[m, n]; if relu_last v = max(v,0); v *= post_scale.
// batching: z in [0, nz); zb = z / nzh, zh = z % nzh; pointer offsets zb*s?_b + zh*s?_h.
// ---------------------------------------------------------------------------------------------
enum RowMode { ROW_PLAIN = 0, ROW_CONV1D = 1, ROW_CONV2D = 2 };

struct GemmArgs {
    const float* A = nullptr; int lda = 0;
    const float* W = nullptr; int ldw = 0;     // !b_kn: W[tap][n][k] (k contiguous, row pitch ldw); b_kn: W[k][n] (n contiguous)
    long long w_tap_stride = 0;
    float* C = nullptr; int ldc = 0;
    const float* bias = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    const float* R = nullptr; int ldr = 0;
    int M = 0, N = 0, K = 0;
    int taps = 1;
    int mode = ROW_PLAIN;
    int Lout = 0, Lin = 0, stride = 1, pad = 0, dil = 1;
    int Ho = 0, Wo = 0, Hi = 0, Wi = 0, ksize = 1;
    int relu_first = 0, relu_last = 0;
    float post_scale = 1.f;
    int b_kn = 0;
    int nz = 1, nzh = 1;
    long long sA_b = 0, sA_h = 0, sW_b = 0, sW_h = 0, sC_b = 0, sC_h = 0;
};

void gemm_simt(const GemmArgs& a, cudaStream_t st);   // dispatches skinny problems (M <= 64, plain) to gemm_skinny
bool gemm_skinny_supported(const GemmArgs& a);
void gemm_skinny(const GemmArgs& a, cudaStream_t st);

}  // namespace zvx
