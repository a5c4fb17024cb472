// HiFi-GAN generator kernels (hifigan.py:114-130), fp32 FMA, channel-first [B, C, T] (time contiguous).
//
// conv1d_cf: dilated 'same' Conv1d with leaky-ReLU fused on the input load and bias / residual / MRF
//            accumulation (xs += resblock(x); x = xs / num_kernels, hifigan.py:119-125) / tanh fused on the
//            store.  A block owns TT*NT consecutive samples of TCO output channels; the input tile (+ halo)
//            and the weight slice are staged in shared memory in chunks of CI_CHUNK input channels; every
//            thread keeps a TT x TCO register tile (time strided by NT -> conflict-free scalar LDS, weights
//            are warp-broadcast LDS.128).
// conv_transpose1d_cf: polyphase form of ConvTranspose1d(k = 2u): out[q*u + r - p] = W[r] x[q] + W[r+u] x[q-1].
#include "kernels.cuh"

namespace zvx {

namespace {

constexpr int NT = 128;
constexpr int TT = 4;
constexpr int CI_CHUNK = 16;
constexpr int TILE = NT * TT;  // 512 samples per block

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

template <int TCO>
__global__ void __launch_bounds__(NT) conv1d_cf_kernel(const Conv1dArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int halo = (a.k - 1) * a.dil;
    const int pad = halo / 2;
    const int xw = TILE + halo;                 // staged samples per input channel
    float* xs = smem;                           // [CI_CHUNK][xw]
    float* ws = smem + CI_CHUNK * xw;           // [CI_CHUNK][k][TCO]
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TILE;
    const int co0 = blockIdx.y * TCO;
    const int b = blockIdx.z;
    const float* xb = a.x + (long long)b * a.Cin * a.T;

    float acc[TT][TCO];
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int c = 0; c < TCO; ++c) acc[i][c] = 0.f;

    for (int ci0 = 0; ci0 < a.Cin; ci0 += CI_CHUNK) {
        const int nci = min(CI_CHUNK, a.Cin - ci0);
        __syncthreads();
        for (int idx = tid; idx < nci * xw; idx += NT) {
            const int ci = idx / xw, p = idx - ci * xw;
            const int t = t0 - pad + p;
            float v = 0.f;
            if (t >= 0 && t < a.T) v = lrelu(__ldg(xb + (long long)(ci0 + ci) * a.T + t), a.in_slope);
            xs[idx] = v;
        }
        for (int idx = tid; idx < nci * a.k * TCO; idx += NT) {
            const int c = idx % TCO;
            const int cj = idx / TCO;  // ci*k + j
            const int co = co0 + c;
            ws[idx] = (co < a.Cout) ? __ldg(a.w + ((long long)ci0 * a.k + cj) * a.Cout + co) : 0.f;
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ++ci) {
            const float* xr = xs + ci * xw + tid;
            const float* wr = ws + ci * a.k * TCO;
            for (int j = 0; j < a.k; ++j) {
                float w[TCO];
#pragma unroll
                for (int c4 = 0; c4 < TCO / 4; ++c4) {
                    const float4 t = *reinterpret_cast<const float4*>(wr + j * TCO + c4 * 4);
                    w[c4 * 4 + 0] = t.x; w[c4 * 4 + 1] = t.y; w[c4 * 4 + 2] = t.z; w[c4 * 4 + 3] = t.w;
                }
                const int off = j * a.dil;
#pragma unroll
                for (int i = 0; i < TT; ++i) {
                    const float xv = xr[off + i * NT];
#pragma unroll
                    for (int c = 0; c < TCO; ++c) acc[i][c] = fmaf(xv, w[c], acc[i][c]);
                }
            }
        }
    }

#pragma unroll
    for (int c = 0; c < TCO; ++c) {
        const int co = co0 + c;
        if (co >= a.Cout) continue;
        const float bias = a.bias ? __ldg(a.bias + co) : 0.f;
        const long long base = ((long long)b * a.Cout + co) * a.T;
#pragma unroll
        for (int i = 0; i < TT; ++i) {
            const int t = t0 + tid + i * NT;
            if (t >= a.T) continue;
            float v = acc[i][c] + bias;
            if (a.res) v += a.res[base + t];
            if (a.tanh_out) v = tanhf(v);
            if (a.out) a.out[base + t] = v;
            if (a.acc) {
                const float s = v * a.acc_scale;
                a.acc[base + t] = a.acc_init ? s : a.acc[base + t] + s;
            }
        }
    }
}

template <int U, int TQ, int TCO>
__global__ void __launch_bounds__(NT) conv_transpose1d_cf_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias, int Cin, int Cout,
                                                                 int T, float in_slope, float* __restrict__ out) {
    // positions q in [q0, q0 + TQ*NT) of the *input* axis; each produces U outputs t = q*U + r - p
    constexpr int K = 2 * U;
    constexpr int P = (K - U) / 2;
    constexpr int QT = TQ * NT;
    __shared__ __align__(16) float xs[CI_CHUNK][QT + 1];     // xs[ci][i] = x[q0 - 1 + i]
    __shared__ __align__(16) float ws[CI_CHUNK][K][TCO];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * QT;
    const int co0 = blockIdx.y * TCO;
    const int b = blockIdx.z;
    const float* xb = x + (long long)b * Cin * T;

    float acc[TQ][U][TCO];
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int r = 0; r < U; ++r)
#pragma unroll
            for (int c = 0; c < TCO; ++c) acc[i][r][c] = 0.f;

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_CHUNK) {
        const int nci = min(CI_CHUNK, Cin - ci0);
        __syncthreads();
        for (int idx = tid; idx < nci * (QT + 1); idx += NT) {
            const int ci = idx / (QT + 1), p = idx - ci * (QT + 1);
            const int s = q0 - 1 + p;
            float v = 0.f;
            if (s >= 0 && s < T) v = lrelu(__ldg(xb + (long long)(ci0 + ci) * T + s), in_slope);
            xs[ci][p] = v;
        }
        for (int idx = tid; idx < nci * K * TCO; idx += NT) {
            const int c = idx % TCO;
            const int cj = idx / TCO;
            const int co = co0 + c;
            (&ws[0][0][0])[idx] = (co < Cout) ? __ldg(w + ((long long)ci0 * K + cj) * Cout + co) : 0.f;
        }
        __syncthreads();
        for (int ci = 0; ci < nci; ++ci) {
            float xq[TQ], xm[TQ];
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                xq[i] = xs[ci][tid + i * NT + 1];
                xm[i] = xs[ci][tid + i * NT];
            }
#pragma unroll
            for (int r = 0; r < U; ++r) {
                float w0[TCO], w1[TCO];
#pragma unroll
                for (int c4 = 0; c4 < TCO / 4; ++c4) {
                    const float4 t0 = *reinterpret_cast<const float4*>(&ws[ci][r][c4 * 4]);
                    const float4 t1 = *reinterpret_cast<const float4*>(&ws[ci][r + U][c4 * 4]);
                    w0[c4 * 4 + 0] = t0.x; w0[c4 * 4 + 1] = t0.y; w0[c4 * 4 + 2] = t0.z; w0[c4 * 4 + 3] = t0.w;
                    w1[c4 * 4 + 0] = t1.x; w1[c4 * 4 + 1] = t1.y; w1[c4 * 4 + 2] = t1.z; w1[c4 * 4 + 3] = t1.w;
                }
#pragma unroll
                for (int i = 0; i < TQ; ++i)
#pragma unroll
                    for (int c = 0; c < TCO; ++c)
                        acc[i][r][c] = fmaf(xm[i], w1[c], fmaf(xq[i], w0[c], acc[i][r][c]));
            }
        }
    }

    const int Tout = T * U;
#pragma unroll
    for (int c = 0; c < TCO; ++c) {
        const int co = co0 + c;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
        float* ob = out + ((long long)b * Cout + co) * Tout;
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const int q = q0 + tid + i * NT;
            if (q > T) continue;
#pragma unroll
            for (int r = 0; r < U; ++r) {
                const int t = q * U + r - P;
                if (t >= 0 && t < Tout) ob[t] = acc[i][r][c] + bv;
            }
        }
    }
}

template <int U, int TQ>
void launch_convtr(const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int T,
                   float in_slope, float* out, cudaStream_t st) {
    constexpr int TCO = 8;
    dim3 grid(cdiv(T + 1, TQ * NT), cdiv(Cout, TCO), B);
    conv_transpose1d_cf_kernel<U, TQ, TCO><<<grid, NT, 0, st>>>(x, w, bias, Cin, Cout, T, in_slope, out);
    ZVX_POST_LAUNCH();
}

// conv_post on channel-last input: wav[t] = tanh(bias + sum_{j, c} w[j][c] * lrelu(x[t + j - (k-1)/2, c])).  One CTA = 1024
// consecutive samples of one utterance: the activated input rows (1024 + k - 1) x C are staged in shared memory once (every value
// was loaded and activated k times, once per output that needs it, by the one-output-per-thread version: 143 us for 215 MB =
// 1.5 TB/s), then every thread produces 4 samples.
constexpr int CP_TILE = 1024;
__global__ void __launch_bounds__(256) conv_post_cl_kernel(const float* __restrict__ x, long long x_bs,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int T, int C, int k, float slope, float* __restrict__ wav) {
    extern __shared__ float sm[];    // [k][C] weights, then [(CP_TILE + k - 1)][C + 4] activated rows (row pitch padded)
    float* wsm = sm;
    float* xs = sm + k * C;
    const int half = (k - 1) / 2, pitch = C + 4, C4 = C >> 2;
    const int t0 = blockIdx.x * CP_TILE, b = blockIdx.y;
    for (int i = threadIdx.x; i < k * C; i += blockDim.x) wsm[i] = w[i];
    const float* xb = x + (long long)b * x_bs;
    const int rows = CP_TILE + k - 1;
    for (int i = threadIdx.x; i < rows * C4; i += blockDim.x) {
        const int r = i / C4, cq = i - r * C4, t = t0 - half + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T) {
            v = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            v = make_float4(lrelu(v.x, slope), lrelu(v.y, slope), lrelu(v.z, slope), lrelu(v.w, slope));
        }
        *reinterpret_cast<float4*>(xs + r * pitch + cq * 4) = v;
    }
    __syncthreads();
    // thread tid produces the samples tid, tid + 256, tid + 512, tid + 768 of the tile: consecutive lanes read consecutive rows
    // (pitch C + 4 floats: conflict-free 128-bit loads) and write consecutive samples
    const float b0 = __ldg(bias);
#pragma unroll
    for (int o = 0; o < CP_TILE / 256; ++o) {
        const int tl = threadIdx.x + o * 256;
        float acc = b0;
        for (int j = 0; j < k; ++j) {
            const float* row = xs + (tl + j) * pitch;
            const float* ww = wsm + j * C;
            for (int cq = 0; cq < C4; ++cq) {
                const float4 v = *reinterpret_cast<const float4*>(row + cq * 4);
                const float4 w4 = *reinterpret_cast<const float4*>(ww + cq * 4);   // one broadcast 128-bit load per four FMAs
                acc = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, acc))));
            }
        }
        if (t0 + tl < T) wav[(long long)b * T + t0 + tl] = tanhf(acc);
    }
}

// The shape every HiFi-GAN config ends in (C = 8 after the last upsampler of V1 / V2, k = 7) with the k * C weights in REGISTERS:
// the generic kernel reads one broadcast weight quad per data quad, so half of its shared-memory traffic was weights and the
// kernel was bound by the shared-memory pipe (28 LDS.128 per sample: 136 us for 242 MB = 1.8 TB/s).  Same FMA order, same result.
template <int C, int K>
__global__ void __launch_bounds__(256) conv_post_cl_reg_kernel(const float* __restrict__ x, long long x_bs,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               int T, float slope, float* __restrict__ wav) {
    extern __shared__ float sm[];    // [(CP_TILE + K - 1)][C + 4] activated rows (row pitch padded)
    float* xs = sm;
    constexpr int half = (K - 1) / 2, pitch = C + 4, C4 = C >> 2, rows = CP_TILE + K - 1;
    const int t0 = blockIdx.x * CP_TILE, b = blockIdx.y;
    pdl_trigger();
    pdl_wait();
    const float* xb = x + (long long)b * x_bs;
#pragma unroll
    for (int i = threadIdx.x; i < rows * C4; i += 256) {   // (every thread's 8-9 loads are in flight together)
        const int r = i / C4, cq = i - r * C4, t = t0 - half + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T) {
            v = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            v = make_float4(lrelu(v.x, slope), lrelu(v.y, slope), lrelu(v.z, slope), lrelu(v.w, slope));
        }
        *reinterpret_cast<float4*>(xs + r * pitch + cq * 4) = v;
    }
    float4 wr[K * C4];   // (fetched while the slower warps still stage their rows)
#pragma unroll
    for (int i = 0; i < K * C4; ++i) wr[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    const float b0 = __ldg(bias);
    __syncthreads();
#pragma unroll
    for (int o = 0; o < CP_TILE / 256; ++o) {
        const int tl = threadIdx.x + o * 256;
        float acc = b0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
#pragma unroll
            for (int cq = 0; cq < C4; ++cq) {
                const float4 v = *reinterpret_cast<const float4*>(xs + (tl + j) * pitch + cq * 4);
                const float4 w4 = wr[j * C4 + cq];
                acc = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, acc))));
            }
        }
        if (t0 + tl < T) wav[(long long)b * T + t0 + tl] = tanhf(acc);
    }
}

}  // namespace

void conv_post_cl(const float* x, long long x_bs, const float* w, const float* bias, int B, int T, int C, int k,
                  float slope, float* wav, cudaStream_t st) {
    if (B == 0 || T == 0) return;
    ZVX_REQUIRE(C % 4 == 0 && (k & 1) == 1 && B <= 65535, "conv_post_cl: bad shape");
    const size_t smem = ((size_t)k * C + (size_t)(CP_TILE + k - 1) * (C + 4)) * sizeof(float);
    ZVX_REQUIRE(smem <= 200 * 1024, "conv_post_cl: channel count too large for the staged tile");
    static bool attr = false;
    if (!attr) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(conv_post_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    dim3 grid(cdiv(T, CP_TILE), B);
    if (C == 8 && k == 7 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
        static bool attr_reg = false;
        if (!attr_reg) {
            ZVX_CUDA_CHECK(cudaFuncSetAttribute(conv_post_cl_reg_kernel<8, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_reg = true;
        }
        launch_k(conv_post_cl_reg_kernel<8, 7>, grid, dim3(256), (size_t)(CP_TILE + 6) * 12 * sizeof(float), st, x, x_bs, w, bias, T, slope, wav);
        ZVX_POST_LAUNCH();
        return;
    }
    conv_post_cl_kernel<<<grid, 256, smem, st>>>(x, x_bs, w, bias, T, C, k, slope, wav);
    ZVX_POST_LAUNCH();
}

void conv1d_cf(const Conv1dArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return;
    ZVX_REQUIRE(a.k % 2 == 1, "conv1d_cf: odd kernel sizes only ('same' padding)");
    ZVX_REQUIRE(a.out || a.acc, "conv1d_cf: no output");
    const int halo = (a.k - 1) * a.dil;
    const bool wide = (a.Cout % 16 == 0);
    const int tco = wide ? 16 : 8;
    const size_t smem = (size_t)(CI_CHUNK * (TILE + halo) + CI_CHUNK * a.k * tco) * sizeof(float);
    ZVX_REQUIRE(smem <= 200 * 1024, "conv1d_cf: tile does not fit shared memory");
    dim3 grid(cdiv(a.T, TILE), cdiv(a.Cout, tco), a.B);
    ZVX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv1d_cf: grid too large");
    if (wide) {
        static bool attr16 = false;
        if (!attr16) {
            ZVX_CUDA_CHECK(cudaFuncSetAttribute(conv1d_cf_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr16 = true;
        }
        conv1d_cf_kernel<16><<<grid, NT, smem, st>>>(a);
    } else {
        static bool attr8 = false;
        if (!attr8) {
            ZVX_CUDA_CHECK(cudaFuncSetAttribute(conv1d_cf_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr8 = true;
        }
        conv1d_cf_kernel<8><<<grid, NT, smem, st>>>(a);
    }
    ZVX_POST_LAUNCH();
}

void conv_transpose1d_cf(const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int T, int k,
                         int u, float in_slope, float* out, cudaStream_t st) {
    if (B == 0 || T == 0) return;
    ZVX_REQUIRE(k == 2 * u, "conv_transpose1d_cf: only kernel = 2*stride (every HiFi-GAN config) is built");
    switch (u) {
        case 8: launch_convtr<8, 2>(x, w, bias, B, Cin, Cout, T, in_slope, out, st); break;
        case 4: launch_convtr<4, 4>(x, w, bias, B, Cin, Cout, T, in_slope, out, st); break;
        case 2: launch_convtr<2, 4>(x, w, bias, B, Cin, Cout, T, in_slope, out, st); break;
        default: ZVX_REQUIRE(false, "conv_transpose1d_cf: upsample rate must be 2, 4 or 8");
    }
}

}  // namespace zvx
