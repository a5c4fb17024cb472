// ResNetSE34V2 speaker-embedding kernels that are not GEMM-shaped (ResNetSE34V2.py:176-212).
// Activations are channel-last [B, H(mel), W(time), C]; the 3x3 / 1x1 convolutions run through the
// implicit-GEMM path (ROW_CONV2D) with ReLU / folded-BatchNorm epilogues.
#include "kernels.cuh"

#include <algorithm>

namespace zvx {

// InstanceNorm1d(n_mels) over time, no affine (ResNetSE34V2.py:123, 182); also performs the
// transpose(1,2) of line 178: in [B,T,M] -> out [B,M,T].  One warp per (utterance, mel channel); two passes (mean,
// then biased variance) like ATen.
__global__ void __launch_bounds__(256) instance_norm_time_kernel(const float* __restrict__ in, int T, int M,
                                                                 float* __restrict__ out) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, nw = (blockDim.x >> 5) * gridDim.y;
    const int wid = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);   // warp index across the blocks of this utterance
    const float* x = in + (long long)b * T * M;
    float* o = out + (long long)b * M * T;
    for (int m = wid; m < M; m += nw) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s += x[(long long)t * M + m];
        const float mu = warp_sum(s) / (float)T;
        float q = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float d = x[(long long)t * M + m] - mu;
            q += d * d;
        }
        const float inv = rsqrtf(warp_sum(q) / (float)T + 1e-5f);
        for (int t = lane; t < T; t += 32) o[(long long)m * T + t] = (x[(long long)t * M + m] - mu) * inv;
    }
}

void instance_norm_time(const float* ref_mel, int B, int T, int n_mels, float* out, cudaStream_t st) {
    if (B == 0) return;
    instance_norm_time_kernel<<<dim3(B, cdiv(n_mels, 8)), 256, 0, st>>>(ref_mel, T, n_mels, out);
    ZVX_POST_LAUNCH();
}

// stem Conv2d(1 -> C, 3x3, pad 1) + bias -> ReLU -> BN (ResNetSE34V2.py:184-186).  Lane = output channel (coalesced
// 128-byte stores per pixel); each thread produces XT consecutive pixels of one row so that its 9 weights stay in
// registers and the 3 x (XT+2) input window is loaded once (warp-broadcast loads).
constexpr int STEM_XT = 8;
__global__ void __launch_bounds__(256) stem_conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, int H, int W, int C, int xtiles,
                                                           long long total, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long p = i / C;
    const int xt = (int)(p % xtiles);
    p /= xtiles;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    const int x0 = xt * STEM_XT;
    const float* ib = in + (long long)b * H * W;
    float wv[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) wv[j] = __ldg(w + j * C + c);
    float win[3][STEM_XT + 2];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yi = y + dy - 1;
#pragma unroll
        for (int dx = 0; dx < STEM_XT + 2; ++dx) {
            const int xi = x0 + dx - 1;
            win[dy][dx] = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? __ldg(ib + (long long)yi * W + xi) : 0.f;
        }
    }
    const float bc = __ldg(bias + c), sc = __ldg(scale + c), sh = __ldg(shift + c);
    float* op = out + (((long long)b * H + y) * W + x0) * C + c;
#pragma unroll
    for (int px = 0; px < STEM_XT; ++px) {
        if (x0 + px >= W) break;
        float acc = 0.f;   // same tap order as before: dy-major, dx-minor
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) acc = fmaf(win[dy][px + dx], wv[dy * 3 + dx], acc);
        op[(long long)px * C] = fmaf(fmaxf(acc + bc, 0.f), sc, sh);
    }
}

void stem_conv3x3(const float* in, const float* w, const float* bias, const float* scale, const float* shift, int B,
                  int H, int W, int C, float* out, cudaStream_t st) {
    const int xtiles = cdiv(W, STEM_XT);
    const long long total = (long long)B * H * xtiles * C;
    if (total == 0) return;
    stem_conv3x3_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, w, bias, scale, shift, H, W, C, xtiles, total, out);
    ZVX_POST_LAUNCH();
}

// SE squeeze (AdaptiveAvgPool2d(1), ResNetSE34V2.py:63-64): x [B, HW, C] -> sums over HW, split into S partial sums so
// that the whole GPU streams the tensor once: grid (S, B); block = 256 threads = (256 / (C/4)) rows x C/4 float4 columns.
// out[b][s][c] = sum over the s-th slice of HW (deterministic order); the excitation adds the S partials and divides.
__device__ __forceinline__ void hw_sum_partial_body(const float* __restrict__ x, int HW, int C4, int rows_per,
                                                    float* __restrict__ out) {
    __shared__ float4 red[256];
    const int s = blockIdx.x, b = blockIdx.y, S = gridDim.x;
    const int c4 = threadIdx.x % C4, r0 = threadIdx.x / C4, nr = 256 / C4;
    const int lo = s * rows_per, hi = min(HW, lo + rows_per);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 < nr) {
        const float4* p = reinterpret_cast<const float4*>(x) + (long long)b * HW * C4 + c4;
        // four independent loads in flight per thread (one per iteration left the kernel latency-bound at 0.54 of the HBM
        // peak, profiles/r02_ncu_hbm_kernels.csv); fixed summation order
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = lo + r0; i < hi; i += 4 * nr) {
            const float4 v0 = __ldg(p + (long long)i * C4);
            const float4 v1 = (i + nr < hi) ? __ldg(p + (long long)(i + nr) * C4) : z4;
            const float4 v2 = (i + 2 * nr < hi) ? __ldg(p + (long long)(i + 2 * nr) * C4) : z4;
            const float4 v3 = (i + 3 * nr < hi) ? __ldg(p + (long long)(i + 3 * nr) * C4) : z4;
            acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
            acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (r0 == 0) {
        for (int r = 1; r < nr; ++r) {
            const float4 v = red[r * C4 + c4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[((long long)b * S + s) * C4 + c4] = acc;
    }
}

int hw_mean_splits(int B, int HW) {
    // 8 CTAs of 256 threads per SM: with 4 (608 CTAs at B = 32) the squeeze ran at 0.54 of the HBM peak, too few loads in flight
    return std::max(1, std::min(cdiv(HW, 32), cdiv(8 * 148, std::max(B, 1))));
}

// SE excitation (ResNetSE34V2.py:55-60, 65) for utterance b from the S partial sums: called by every thread of one block.
__device__ __forceinline__ void se_excite_block(const float* p, int b, int S, float inv_hw, const float* __restrict__ w1,
                                                const float* __restrict__ b1, const float* __restrict__ w2,
                                                const float* __restrict__ b2, int C, int R, float* __restrict__ y, float* sh) {
    float* ps = sh;        // [C] pooled
    float* hs = sh + C;    // [R] hidden
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int s = 0; s < S; ++s) t += __ldcg(p + ((long long)b * S + s) * C + c);   // fixed order; written by other blocks: L2
        ps[c] = t * inv_hw;
    }
    __syncthreads();
    for (int r = wid; r < R; r += nw) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w1 + (long long)r * C + c), ps[c], s);
        s = warp_sum(s);
        if (lane == 0) hs[r] = fmaxf(s + __ldg(b1 + r), 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = __ldg(b2 + c);
        for (int r = 0; r < R; ++r) s = fmaf(__ldg(w2 + (long long)c * R + r), hs[r], s);
        y[(long long)b * C + c] = 1.f / (1.f + expf(-s));
    }
}

// Squeeze and excitation in ONE launch: grid (S, B) streams the tensor once (hw_sum_partial_body); the block of utterance b
// that finishes last (a ticket per utterance, release / acquire fences around it) adds the S partials in index order —
// the gate does not depend on which block that is — and runs the two small linears.  `ticket` [B] must be zero at
// the first launch; the last block leaves it zero again.
__global__ void __launch_bounds__(256) se_squeeze_excite_kernel(const float* __restrict__ x, int HW, int C4, int rows_per,
                                                                float* __restrict__ partial, int* __restrict__ ticket,
                                                                const float* __restrict__ w1, const float* __restrict__ b1,
                                                                const float* __restrict__ w2, const float* __restrict__ b2,
                                                                int R, float* __restrict__ y) {
    extern __shared__ float sh[];
    __shared__ int last;
    pdl_trigger();
    pdl_wait();
    hw_sum_partial_body(x, HW, C4, rows_per, partial);
    __threadfence();   // this block's partial before its ticket
    __syncthreads();
    const int b = blockIdx.y, S = gridDim.x;
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket + b, 1);
        last = (t == S - 1);
        if (last) ticket[b] = 0;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();   // every other block's partial after the ticket
    se_excite_block(partial, b, S, 1.f / (float)HW, w1, b1, w2, b2, 4 * C4, R, y, sh);
}

void se_squeeze_excite(const float* x, int B, int HW, int C, int S, float* partial, int* ticket, const float* w1,
                       const float* b1, const float* w2, const float* b2, int R, float* y, cudaStream_t st) {
    if (B == 0) return;
    ZVX_REQUIRE(C % 4 == 0 && C / 4 <= 256, "se_squeeze_excite: C must be a multiple of 4, <= 1024");
    launch_k(se_squeeze_excite_kernel, dim3(S, B), dim3(256), (C + R) * sizeof(float), st, x, HW, C / 4, cdiv(HW, S), partial, ticket,
             w1, b1, w2, b2, R, y);
    ZVX_POST_LAUNCH();
}

// out = relu(x * y[b, c] + res)   (ResNetSE34V2.py:66-67, 97-98)
__global__ void se_scale_add_relu_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                         const float* __restrict__ res, long long n4, long long HWC4, int C4,
                                         float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int b = (int)(i / HWC4);
    const int c4 = (int)(i % C4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(y) + (long long)b * C4 + c4);
    const float4 r = __ldg(reinterpret_cast<const float4*>(res) + i);
    float4 o;
    o.x = fmaxf(fmaf(a.x, g.x, r.x), 0.f);
    o.y = fmaxf(fmaf(a.y, g.y, r.y), 0.f);
    o.z = fmaxf(fmaf(a.z, g.z, r.z), 0.f);
    o.w = fmaxf(fmaf(a.w, g.w, r.w), 0.f);
    reinterpret_cast<float4*>(out)[i] = o;
}

void se_scale_add_relu(const float* x, const float* y, const float* res, int B, int HW, int C, float* out,
                       cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0, "se_scale_add_relu: C % 4");
    const long long n4 = (long long)B * HW * C / 4;
    if (n4 == 0) return;
    launch_k(se_scale_add_relu_kernel, dim3(cdiv(n4, 256)), dim3(256), 0, st, x, y, res, n4, (long long)HW * C / 4, C / 4, out);
    ZVX_POST_LAUNCH();
}

// [B, Hh, W, C] -> [B, W, C*Hh], feature f = c*Hh + h   (x.reshape(B, -1, W) of an NCHW tensor, line 196)
__global__ void spk_flatten_kernel(const float* __restrict__ x, int Hh, int W, int C, long long total,
                                   float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int D = C * Hh;
    const int f = (int)(i % D);
    long long p = i / D;
    const int w = (int)(p % W);
    const int b = (int)(p / W);
    const int c = f / Hh, h = f - c * Hh;
    out[i] = __ldg(x + (((long long)b * Hh + h) * W + w) * C + c);
}

void spk_flatten(const float* x, int B, int Hh, int W, int C, float* out, cudaStream_t st) {
    const long long total = (long long)B * Hh * W * C;
    if (total == 0) return;
    spk_flatten_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, Hh, W, C, total, out);
    ZVX_POST_LAUNCH();
}

// Softmax over time + attentive statistics pooling (ResNetSE34V2.py:139-147, 198-204).
// One thread per (b, feature); consecutive threads = consecutive features -> coalesced over D.
__global__ void attentive_pool_kernel(const float* __restrict__ feat, const float* __restrict__ logits, int W, int D,
                                      int asp, long long total, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const int b = (int)(i / D);
    const float* lg = logits + (long long)b * W * D + d;
    const float* ft = feat + (long long)b * W * D + d;
    float m = -INFINITY;
    for (int t = 0; t < W; ++t) m = fmaxf(m, lg[(long long)t * D]);
    float den = 0.f;
    for (int t = 0; t < W; ++t) den += expf(lg[(long long)t * D] - m);
    float mu = 0.f, sq = 0.f;
    for (int t = 0; t < W; ++t) {
        const float wgt = expf(lg[(long long)t * D] - m) / den;
        const float x = ft[(long long)t * D];
        mu = fmaf(x, wgt, mu);
        sq = fmaf(x * x, wgt, sq);
    }
    if (asp) {
        out[(long long)b * 2 * D + d] = mu;
        out[(long long)b * 2 * D + D + d] = sqrtf(fmaxf(sq - mu * mu, 1e-5f));
    } else {
        out[(long long)b * D + d] = mu;
    }
}

void attentive_pool(const float* feat, const float* logits, int B, int W, int D, int asp, float* out,
                    cudaStream_t st) {
    const long long total = (long long)B * D;
    if (total == 0) return;
    attentive_pool_kernel<<<cdiv(total, 128), 128, 0, st>>>(feat, logits, W, D, asp, total, out);
    ZVX_POST_LAUNCH();
}

__global__ void __launch_bounds__(256) l2_normalize_kernel(float* __restrict__ x, int C) {
    __shared__ float red[8];
    __shared__ float inv;
    float* r = x + (long long)blockIdx.x * C;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += 256) s = fmaf(r[c], r[c], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        inv = 1.f / fmaxf(sqrtf(t), 1e-12f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) r[c] *= inv;
}

void l2_normalize(float* x, int B, int C, cudaStream_t st) {
    if (B == 0) return;
    l2_normalize_kernel<<<B, 256, 0, st>>>(x, C);
    ZVX_POST_LAUNCH();
}

}  // namespace zvx
