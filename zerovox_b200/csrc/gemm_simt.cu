// fp32 FMA GEMM / implicit-GEMM convolution (CUDA cores).  This is the *accuracy* path: the encoder and
// the variance predictors must reproduce the reference's fp32 arithmetic closely enough that the integer
// decisions downstream (pitch/energy buckets fs2.py:639/649, rounded durations fs2.py:678-681) do not flip,
// so their contractions never go through TF32.  It is also the bring-up path for every other contraction.
//
// Tiling: 128 x BN x 16 block tile, 256 threads, 8 x (BN/16) register tile per thread, double-buffered
// shared memory with register-staged global loads (one __syncthreads per k-step).
#include "common.cuh"

#include <algorithm>

namespace zvx {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;

struct RowMap {
    // Returns the input row for output row m and tap `tap`, or -1 when the tap falls in the zero padding.
    __device__ static __forceinline__ long long map(const GemmArgs& a, int m, int tap) {
        if (a.mode == ROW_PLAIN) return m;
        if (a.mode == ROW_CONV1D) {
            int s = m / a.Lout, t = m - s * a.Lout;
            int ti = t * a.stride + tap * a.dil - a.pad;
            if (ti < 0 || ti >= a.Lin) return -1;
            return (long long)s * a.Lin + ti;
        }
        int x = m % a.Wo;
        int q = m / a.Wo;
        int y = q % a.Ho;
        int img = q / a.Ho;
        int dy = tap / a.ksize, dx = tap - dy * a.ksize;
        int yi = y * a.stride + dy - a.pad, xi = x * a.stride + dx - a.pad;
        if (yi < 0 || yi >= a.Hi || xi < 0 || xi >= a.Wi) return -1;
        return ((long long)img * a.Hi + yi) * a.Wi + xi;
    }
};

template <int BN>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmArgs a, const int a_vec, const int w_vec) {
    constexpr int TN = BN / 16;          // 8 or 4
    constexpr int WF4 = (BN * BK / 4) / NT;  // float4 loads of the W tile per thread: 2 or 1
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z;
    const int zb = z / a.nzh, zh = z - zb * a.nzh;
    const float* __restrict__ Ab = a.A + zb * a.sA_b + zh * a.sA_h;
    const float* __restrict__ Wb = a.W + zb * a.sW_b + zh * a.sW_h;
    float* __restrict__ Cb = a.C + zb * a.sC_b + zh * a.sC_h;
    const float* __restrict__ Rb = a.R ? a.R + zb * a.sC_b + zh * a.sC_h : nullptr;

    const int KT = (a.K + BK - 1) / BK;
    const int iters = a.taps * KT;

    // global -> register staging
    const int lr = tid >> 2;  // tile row 0..63 (+64)
    const int kq = tid & 3;   // float4 index along K
    const float* arow[2];
    float4 ra[2], rw[2];

    auto set_tap = [&](int tap) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int m = m0 + lr + i * 64;
            long long r = (m < a.M) ? RowMap::map(a, m, tap) : -1;
            arow[i] = (r >= 0) ? Ab + r * a.lda : nullptr;
        }
    };

    auto load_tiles = [&](int it) {
        const int tap = it / KT;
        const int k0 = (it - tap * KT) * BK;
        if (it - tap * KT == 0) set_tap(tap);
        const int k = k0 + kq * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (arow[i]) {
                if (a_vec && k + 4 <= a.K) {
                    v = __ldg(reinterpret_cast<const float4*>(arow[i] + k));
                } else {
                    if (k + 0 < a.K) v.x = __ldg(arow[i] + k + 0);
                    if (k + 1 < a.K) v.y = __ldg(arow[i] + k + 1);
                    if (k + 2 < a.K) v.z = __ldg(arow[i] + k + 2);
                    if (k + 3 < a.K) v.w = __ldg(arow[i] + k + 3);
                }
            }
            ra[i] = v;
        }
        if (!a.b_kn) {
            const float* Wt = Wb + (long long)tap * a.w_tap_stride;
#pragma unroll
            for (int i = 0; i < WF4; ++i) {
                int n = n0 + lr + i * 64;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < a.N) {
                    const float* p = Wt + (long long)n * a.ldw;
                    if (w_vec && k + 4 <= a.K) {
                        v = __ldg(reinterpret_cast<const float4*>(p + k));
                    } else {
                        if (k + 0 < a.K) v.x = __ldg(p + k + 0);
                        if (k + 1 < a.K) v.y = __ldg(p + k + 1);
                        if (k + 2 < a.K) v.z = __ldg(p + k + 2);
                        if (k + 3 < a.K) v.w = __ldg(p + k + 3);
                    }
                }
                rw[i] = v;
            }
        } else {
#pragma unroll
            for (int i = 0; i < WF4; ++i) {
                int f = tid + i * NT;
                int kk = f / (BN / 4), n4 = f - kk * (BN / 4);
                int kg = k0 + kk, n = n0 + n4 * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kg < a.K) {
                    const float* p = Wb + (long long)kg * a.ldw + n;
                    if (w_vec && n + 4 <= a.N) {
                        v = __ldg(reinterpret_cast<const float4*>(p));
                    } else {
                        if (n + 0 < a.N) v.x = __ldg(p + 0);
                        if (n + 1 < a.N) v.y = __ldg(p + 1);
                        if (n + 2 < a.N) v.z = __ldg(p + 2);
                        if (n + 3 < a.N) v.w = __ldg(p + 3);
                    }
                }
                rw[i] = v;
            }
        }
    };

    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = lr + i * 64;
            As[buf][kq * 4 + 0][r] = ra[i].x;
            As[buf][kq * 4 + 1][r] = ra[i].y;
            As[buf][kq * 4 + 2][r] = ra[i].z;
            As[buf][kq * 4 + 3][r] = ra[i].w;
        }
        if (!a.b_kn) {
#pragma unroll
            for (int i = 0; i < WF4; ++i) {
                int r = lr + i * 64;
                Bs[buf][kq * 4 + 0][r] = rw[i].x;
                Bs[buf][kq * 4 + 1][r] = rw[i].y;
                Bs[buf][kq * 4 + 2][r] = rw[i].z;
                Bs[buf][kq * 4 + 3][r] = rw[i].w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < WF4; ++i) {
                int f = tid + i * NT;
                int kk = f / (BN / 4), n4 = f - kk * (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kk][n4 * 4]) = rw[i];
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();

    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) load_tiles(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[8], bv[TN];
            float4 t0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 t1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            av[0] = t0.x; av[1] = t0.y; av[2] = t0.z; av[3] = t0.w;
            av[4] = t1.x; av[5] = t1.y; av[6] = t1.z; av[7] = t1.w;
            float4 u0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            bv[0] = u0.x; bv[1] = u0.y; bv[2] = u0.z; bv[3] = u0.w;
            if (TN == 8) {
                float4 u1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
                bv[TN - 4] = u1.x; bv[TN - 3] = u1.y; bv[TN - 2] = u1.z; bv[TN - 1] = u1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < iters) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
    const bool c_vec = ((a.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0) &&
                       (!Rb || (((a.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(Rb) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= a.M) continue;
#pragma unroll
        for (int jg = 0; jg < TN / 4; ++jg) {
            const int n = n0 + (jg == 0 ? tx * 4 : BN / 2 + tx * 4);
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jg * 4 + j];
                const int nn = n + j;
                if (nn < a.N) {
                    if (a.bias) x += __ldg(a.bias + nn);
                    if (a.relu_first) x = fmaxf(x, 0.f);
                    if (a.scale) x = fmaf(x, __ldg(a.scale + nn), __ldg(a.shift + nn));
                }
                v[j] = x;
            }
            float* cp = Cb + (long long)m * a.ldc + n;
            const float* rp = Rb ? Rb + (long long)m * a.ldr + n : nullptr;
            if (c_vec && n + 4 <= a.N) {
                if (rp) {
                    float4 r = *reinterpret_cast<const float4*>(rp);
                    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
                }
                if (a.relu_last) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= a.post_scale;
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < a.N) {
                        float x = v[j];
                        if (rp) x += rp[j];
                        if (a.relu_last) x = fmaxf(x, 0.f);
                        cp[j] = x * a.post_scale;
                    }
                }
            }
        }
    }
}

// Skinny GEMM (M <= 32 rows per pass, any K): lane = output row, warp = 8 output columns, block = 64 columns.  A [32 x 128]
// and W [64 x 128] slabs go through shared memory once per block (A reads conflict-free across lanes, W reads are warp
// broadcasts), so the weight matrix is streamed exactly once by the whole grid.  Used when N gives enough 64-column blocks
// (SCLN / AdaIN affine stacks); narrow outputs (speaker-net fc, N = 528) use the one-warp-per-column kernel below.
// fp32 FMA throughout, deterministic summation order.
constexpr int SK_KC = 64, SK_NB = 64;
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmArgs a) {
    __shared__ float As[32][SK_KC + 1];
    __shared__ float Ws[SK_NB][SK_KC + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n0 = blockIdx.x * SK_NB, m0 = blockIdx.y * 32;
    const int mrows = min(32, a.M - m0);
    const int kc0 = 0, kc1 = (a.K + SK_KC - 1) / SK_KC;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int kc = kc0; kc < kc1; ++kc) {
        const int k0 = kc * SK_KC;
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * SK_KC; i += 256) {
            const int m = i / SK_KC, k = i - m * SK_KC;
            As[m][k] = (m < mrows && k0 + k < a.K) ? __ldg(a.A + (long long)(m0 + m) * a.lda + k0 + k) : 0.f;
        }
        for (int i = threadIdx.x; i < SK_NB * SK_KC; i += 256) {
            const int n = i / SK_KC, k = i - n * SK_KC;
            Ws[n][k] = (n0 + n < a.N && k0 + k < a.K) ? __ldg(a.W + (long long)(n0 + n) * a.ldw + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < SK_KC; ++k) {
            const float av = As[lane][k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(av, Ws[wid * 8 + j][k], acc[j]);
        }
    }
    if (lane >= mrows) return;
    const int m = m0 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + wid * 8 + j;
        if (n >= a.N) continue;
        float x = acc[j];
        if (a.bias) x += __ldg(a.bias + n);
        if (a.relu_first) x = fmaxf(x, 0.f);
        if (a.scale) x = fmaf(x, __ldg(a.scale + n), __ldg(a.shift + n));
        if (a.R) x += a.R[(long long)m * a.ldr + n];
        if (a.relu_last) x = fmaxf(x, 0.f);
        a.C[(long long)m * a.ldc + n] = x * a.post_scale;
    }
}

// one warp per output column, lanes stride over K (narrow outputs)
constexpr int SKC_KC = 128;
__global__ void __launch_bounds__(256) gemm_skinny_col_kernel(const GemmArgs a) {
    __shared__ float As[32][SKC_KC + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + wid;
    const int m0 = blockIdx.y * 32;
    const int mrows = min(32, a.M - m0);
    float acc[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[m] = 0.f;
    const float* wrow = a.W + (long long)min(n, a.N - 1) * a.ldw;
    for (int k0 = 0; k0 < a.K; k0 += SKC_KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * SKC_KC; i += 256) {
            const int m = i / SKC_KC, k = i - m * SKC_KC;
            As[m][k] = (m < mrows && k0 + k < a.K) ? __ldg(a.A + (long long)(m0 + m) * a.lda + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SKC_KC; kk += 32) {
            const int k = k0 + kk + lane;
            const float w = (k < a.K) ? __ldg(wrow + k) : 0.f;
#pragma unroll
            for (int m = 0; m < 32; ++m) acc[m] = fmaf(As[m][kk + lane], w, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[m] = warp_sum(acc[m]);
    if (n >= a.N) return;
    // lane m finishes row m0 + m
    float x = 0.f;
#pragma unroll
    for (int m = 0; m < 32; ++m)
        if (lane == m) x = acc[m];
    if (lane < mrows) {
        const int m = m0 + lane;
        if (a.bias) x += __ldg(a.bias + n);
        if (a.relu_first) x = fmaxf(x, 0.f);
        if (a.scale) x = fmaf(x, __ldg(a.scale + n), __ldg(a.shift + n));
        if (a.R) x += a.R[(long long)m * a.ldr + n];
        if (a.relu_last) x = fmaxf(x, 0.f);
        a.C[(long long)m * a.ldc + n] = x * a.post_scale;
    }
}

}  // namespace

// Streaming skinny GEMM (M <= 32): the A block [M][K-chunk] sits in shared memory, every warp owns whole output columns and
// streams their weight rows straight from HBM with 128-bit loads (no shared-memory round trip, no barrier per k-slab), each lane
// accumulating all M rows for its k's; one shuffle reduction per column.  The weight matrix is read exactly once by the grid and
// enough loads are in flight to approach HBM speed — the two kernels above run at 0.1-0.4 TB/s because every 64 / 128-wide
// k-slab costs a block barrier (SCLN affine stack: 26.8 MB in 70 us; speaker-net fc: 10.8 MB in 121 us).
// Needs K % 4 == 0 and 16-byte aligned operands.  When K exceeds the shared-memory chunk, a warp keeps its column's accumulators
// across chunks, so every warp may own at most one column (N <= warps of the grid): the speaker-net fc.
constexpr int SKS_KC = 1056;                      // floats per A-chunk row (32 rows -> 132 KB of shared memory)
__global__ void __launch_bounds__(256) gemm_skinny_stream_kernel(const GemmArgs a, int cols_per_warp) {
    extern __shared__ float4 As4[];               // [32][kc4 + 1] float4, row pitch padded against bank conflicts
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps = gridDim.x * 8, w0 = blockIdx.x * 8 + wid;
    const int m0 = blockIdx.y * 32, mrows = min(32, a.M - m0);
    const int nchunks = (a.K + SKS_KC - 1) / SKS_KC;
    float acc[32];
    for (int ci = 0; ci < cols_per_warp; ++ci) {
        const int n = w0 + ci * warps;             // (nchunks > 1 => cols_per_warp == 1: the accumulators survive the chunk loop)
#pragma unroll
        for (int m = 0; m < 32; ++m) acc[m] = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int k0 = ch * SKS_KC, kc = min(SKS_KC, a.K - k0), kc4 = kc >> 2, pitch = (SKS_KC >> 2) + 1;
            if (ci == 0 || nchunks > 1) {          // (single chunk: staged once, reused for every column of the warp)
                __syncthreads();
                for (int i = threadIdx.x; i < 32 * kc4; i += 256) {
                    const int m = i / kc4, k4 = i - m * kc4;
                    As4[m * pitch + k4] = (m < mrows) ? __ldg(reinterpret_cast<const float4*>(a.A + (long long)(m0 + m) * a.lda + k0) + k4)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                __syncthreads();
            }
            if (n < a.N) {
                const float4* __restrict__ wrow = reinterpret_cast<const float4*>(a.W + (long long)n * a.ldw + k0);
                for (int k4 = lane; k4 < kc4; k4 += 32) {
                    const float4 w = __ldg(wrow + k4);
#pragma unroll
                    for (int m = 0; m < 32; ++m) {
                        const float4 x = As4[m * pitch + k4];
                        acc[m] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[m]))));
                    }
                }
            }
        }
        if (n >= a.N) continue;                    // (warp-uniform)
#pragma unroll
        for (int m = 0; m < 32; ++m) acc[m] = warp_sum(acc[m]);
        float x = 0.f;
#pragma unroll
        for (int m = 0; m < 32; ++m)
            if (lane == m) x = acc[m];
        if (lane < mrows) {
            const int m = m0 + lane;
            if (a.bias) x += __ldg(a.bias + n);
            if (a.relu_first) x = fmaxf(x, 0.f);
            if (a.scale) x = fmaf(x, __ldg(a.scale + n), __ldg(a.shift + n));
            if (a.R) x += a.R[(long long)m * a.ldr + n];
            if (a.relu_last) x = fmaxf(x, 0.f);
            a.C[(long long)m * a.ldc + n] = x * a.post_scale;
        }
    }
}

bool gemm_skinny_supported(const GemmArgs& a) {
    return a.mode == ROW_PLAIN && a.taps == 1 && a.nz == 1 && !a.b_kn && a.M >= 1 && a.M <= 64 && a.K >= 256 && a.N >= 8;
}

void gemm_skinny(const GemmArgs& a, cudaStream_t st) {
    ZVX_REQUIRE(gemm_skinny_supported(a), "gemm_skinny: unsupported problem");
    {   // streaming kernel when the layout allows 128-bit loads and the column -> warp assignment fits its accumulator rule
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            ZVX_CUDA_CHECK(cudaGetDevice(&dev));
            ZVX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        const bool vec = (a.K % 4 == 0) && (a.lda % 4 == 0) && (a.ldw % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W)) & 15) == 0;
        const int nchunks = cdiv(a.K, SKS_KC);
        const int gx = (int)std::min<long long>(cdiv(a.N, 8), 1LL * sms);   // one CTA per SM (132 KB of shared memory each)
        const int cols_per_warp = cdiv(a.N, gx * 8);
        if (vec && a.M <= 32 && (nchunks == 1 || cols_per_warp == 1)) {
            const int smem = 32 * ((SKS_KC >> 2) + 1) * (int)sizeof(float4);
            static bool attr = false;
            if (!attr) {
                ZVX_CUDA_CHECK(cudaFuncSetAttribute(gemm_skinny_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                attr = true;
            }
            gemm_skinny_stream_kernel<<<dim3(gx, cdiv(a.M, 32)), 256, smem, st>>>(a, cols_per_warp);
            ZVX_POST_LAUNCH();
            return;
        }
    }
    if (cdiv(a.N, SK_NB) >= 64) {
        gemm_skinny_kernel<<<dim3(cdiv(a.N, SK_NB), cdiv(a.M, 32)), 256, 0, st>>>(a);
    } else {
        gemm_skinny_col_kernel<<<dim3(cdiv(a.N, 8), cdiv(a.M, 32)), 256, 0, st>>>(a);
    }
    ZVX_POST_LAUNCH();
}

void gemm_simt(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.nz <= 0) return;
    ZVX_REQUIRE(a.K > 0 && a.taps >= 1, "gemm: bad K/taps");
    ZVX_REQUIRE(!a.b_kn || a.taps == 1, "gemm: b_kn operands cannot have taps");
    ZVX_REQUIRE(a.R == nullptr || a.ldr > 0, "gemm: residual needs ldr");
    if (gemm_skinny_supported(a)) return gemm_skinny(a, st);
    auto al4 = [](long long v) { return (v & 3) == 0; };
    const int a_vec = al4(a.lda) && al4(a.sA_b) && al4(a.sA_h) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    const int w_vec = al4(a.ldw) && al4(a.sW_b) && al4(a.sW_h) && al4(a.w_tap_stride) &&
                      ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
    // pick the N tile that wastes fewer columns
    const long long waste128 = round_up(a.N, 128) - a.N, waste64 = round_up(a.N, 64) - a.N;
    const bool use64 = waste64 < waste128;
    if (use64) {
        dim3 grid(cdiv(a.N, 64), cdiv(a.M, BM), a.nz);
        gemm_simt_kernel<64><<<grid, NT, 0, st>>>(a, a_vec, w_vec);
    } else {
        dim3 grid(cdiv(a.N, 128), cdiv(a.M, BM), a.nz);
        gemm_simt_kernel<128><<<grid, NT, 0, st>>>(a, a_vec, w_vec);
    }
    ZVX_POST_LAUNCH();
}

}  // namespace zvx
