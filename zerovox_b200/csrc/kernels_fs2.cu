// HBM-bound kernels of the FastSpeech2 acoustic model: embedding gather, (SC)LayerNorm, masked softmax,
// variance-adaptor bucketing, duration rounding / scan and the LengthRegulator gather.
// All row kernels use 128-bit accesses where the channel count allows and one warp per row.
#include "kernels.cuh"

namespace zvx {

long long g_launches = 0;
int g_pdl = 0;   // measured: +-0.5 % on configs[1] — opt-in with zvx_set_option("pdl", 1)

// ------------------------------------------------------------------------------------------------
// K1 embedding + position encoding (fs2.py:372-392)
// ------------------------------------------------------------------------------------------------
__global__ void embed_posenc_kernel(const int32_t* __restrict__ phoneme, const int32_t* __restrict__ puncts,
                                    const float* __restrict__ phon_emb, const float* __restrict__ punct_emb,
                                    const float* __restrict__ pos, int rows, int T, int E, int P,
                                    float* __restrict__ out) {
    const int row = blockIdx.x;
    if (row >= rows) return;
    const int t = row % T;
    const int C = E + P;
    const float* pe = phon_emb + (long long)phoneme[row] * E;
    const float* qe = punct_emb + (long long)puncts[row] * P;
    const float* ps = pos + (long long)t * C;
    float* o = out + (long long)row * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = (c < E) ? __ldg(pe + c) : __ldg(qe + (c - E));
        o[c] = v + __ldg(ps + c);
    }
}

void embed_posenc(const int32_t* phoneme, const int32_t* puncts, const float* phon_emb, const float* punct_emb,
                  const float* pos, int B, int T, int E, int P, float* out, cudaStream_t st) {
    if (B * T == 0) return;
    embed_posenc_kernel<<<B * T, 128, 0, st>>>(phoneme, puncts, phon_emb, punct_emb, pos, B * T, T, E, P, out);
    ZVX_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// K6 LayerNorm / SCLN, one warp per row, row held in registers (C <= 1024)
// ------------------------------------------------------------------------------------------------
constexpr int NORM_MAXV = 8;  // float4 per lane -> C <= 32*4*8 = 1024

// NV = float4 per lane the row needs (ceil(C / 128)): the 528-wide rows of the FFT blocks take 5, the 256-wide rows of the variance
// predictors 2 — sized for 1024 columns the kernel held 68 registers and 24 warps per SM (profiles/r02_ncu_hbm_kernels.csv: warps
// active 35 %, 0.62 of the HBM peak); fewer live registers = more rows in flight per SM.
template <int NV>
__global__ void __launch_bounds__(256, NV <= 2 ? 6 : NV <= 5 ? 5 : 3) layer_norm_kernel(const NormArgs a) {
    pdl_trigger();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= a.rows) return;
    const int row = warp;
    const bool masked = a.mask && a.mask[row];
    if (masked) {
        if (a.dot_w) {
            if (lane == 0) a.dot_out[row] = 0.f;
        } else {
            float4* o = reinterpret_cast<float4*>(a.out + (long long)row * a.C);
            for (int c4 = lane; c4 < a.C / 4; c4 += 32) o[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    const int C4 = a.C >> 2;
    const float4* xr = reinterpret_cast<const float4*>(a.x + (long long)row * a.C);
    const float4* rr = a.res ? reinterpret_cast<const float4*>(a.res + (long long)row * a.C) : nullptr;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        if (c4 < C4) {
            v[i] = xr[c4];
            if (rr) {
                const float4 t = rr[c4];
                v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
            }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mu = warp_sum(s) / (float)a.C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        if (c4 < C4) {
            float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
            q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    q = warp_sum(q);
    float inv;
    const float *gp, *bp;
    if (a.scln) {
        const float sigma = sqrtf(q / (float)(a.C - 1));  // torch.std: unbiased
        inv = 1.f / (sigma + a.eps);
        const float* gb = a.gb + (long long)(row / a.rows_per_batch) * a.gb_ld;
        bp = gb;          // bias rows first  (fs2.py:85)
        gp = gb + a.C;    // gain rows second
    } else {
        inv = rsqrtf(q / (float)a.C + a.eps);
        gp = a.gamma;
        bp = a.beta;
    }
    float dot = 0.f;
    float4* o = a.dot_w ? nullptr : reinterpret_cast<float4*>(a.out + (long long)row * a.C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        if (c4 < C4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gp) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(bp) + c4);
            float4 y;
            if (a.scln) {
                // o = g * ((x - mu) / (sigma + eps)) + b
                y.x = fmaf(g.x, (v[i].x - mu) * inv, b.x);
                y.y = fmaf(g.y, (v[i].y - mu) * inv, b.y);
                y.z = fmaf(g.z, (v[i].z - mu) * inv, b.z);
                y.w = fmaf(g.w, (v[i].w - mu) * inv, b.w);
            } else {
                y.x = fmaf((v[i].x - mu) * inv, g.x, b.x);
                y.y = fmaf((v[i].y - mu) * inv, g.y, b.y);
                y.z = fmaf((v[i].z - mu) * inv, g.z, b.z);
                y.w = fmaf((v[i].w - mu) * inv, g.w, b.w);
            }
            if (a.dot_w) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.dot_w) + c4);
                dot += (y.x * w.x + y.y * w.y) + (y.z * w.z + y.w * w.w);
            } else {
                o[c4] = y;
            }
        }
    }
    if (a.dot_w) {
        dot = warp_sum(dot);
        if (lane == 0) a.dot_out[row] = dot + (a.dot_b ? __ldg(a.dot_b) : 0.f);
    }
}

void layer_norm(const NormArgs& a, cudaStream_t st) {
    if (a.rows == 0) return;
    ZVX_REQUIRE(a.C % 4 == 0 && a.C <= 128 * NORM_MAXV, "layer_norm: C must be a multiple of 4 and <= 1024");
    ZVX_REQUIRE(!a.scln || (a.gb && (a.gb_ld % 4) == 0 && a.C > 1), "layer_norm: SCLN needs per-batch affine rows");
    const int warps_per_block = 8;
    const int nv = cdiv(a.C, 128);
    const dim3 grid(cdiv(a.rows, warps_per_block)), block(warps_per_block * 32);
    if (nv <= 2) launch_k(layer_norm_kernel<2>, grid, block, 0, st, a);
    else if (nv <= 5) launch_k(layer_norm_kernel<5>, grid, block, 0, st, a);
    else launch_k(layer_norm_kernel<NORM_MAXV>, grid, block, 0, st, a);
    ZVX_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// K4 masked softmax over keys, one block per (z, q) row  (fs2.py:49-55)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_softmax_kernel(float* __restrict__ S, int nh, int Lq, int L, int ldS,
                                                           const uint8_t* __restrict__ key_mask, int mask_ld,
                                                           float temperature) {
    const long long rowid = blockIdx.x;  // z*Lq + q
    const int z = (int)(rowid / Lq);
    const int b = z / nh;
    float* s = S + rowid * ldS;
    const uint8_t* km = key_mask ? key_mask + (long long)b * mask_ld : nullptr;
    __shared__ float red[4];
    __shared__ float bcast;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    float m = -INFINITY;
    for (int j = tid; j < L; j += 128) {
        float v = s[j] / temperature;
        if (km && km[j]) v = -INFINITY;
        s[j] = v;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    if (lane == 0) red[wid] = m;
    __syncthreads();
    if (tid == 0) bcast = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    m = bcast;
    float sum = 0.f;
    for (int j = tid; j < L; j += 128) {
        float e = expf(s[j] - m);
        s[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[wid] = sum;
    __syncthreads();
    if (tid == 0) bcast = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    const float tot = bcast;
    for (int j = tid; j < ldS; j += 128) s[j] = (j < L) ? s[j] / tot : 0.f;
}

// Same arithmetic, one warp per row with the row held in registers (one read + one write of S): rows of up to
// 32 * 4 * NV keys.  8 rows per block.
template <int NV>
__global__ void __launch_bounds__(256) attn_softmax_warp_kernel(float* __restrict__ S, int nh, int Lq, int L, int ldS,
                                                                const uint8_t* __restrict__ key_mask, int mask_ld,
                                                                float temperature, long long rows) {
    const long long rowid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (rowid >= rows) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(rowid / Lq) / nh;
    float4* s4 = reinterpret_cast<float4*>(S + rowid * ldS);
    const uint8_t* km = key_mask ? key_mask + (long long)b * mask_ld : nullptr;
    const int n4 = ldS >> 2;
    // exp(x/T - m) = exp2((x - max) * log2(e) / T): one multiply per element, ex2.approx (2 ulp) instead of a division
    // and a full-precision expf; the final 1/sum is one reciprocal per row
    const float sc = 1.4426950408889634f / temperature;
    float v[NV][4];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < n4) t = s4[idx];
        const float in[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = idx * 4 + e;
            float x = -INFINITY;
            if (j < L && !(km && km[j])) x = in[e];
            v[i][e] = x;
            m = fmaxf(m, x);
        }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[i][e] = exp2f((v[i][e] - m) * sc);
            sum += v[i][e];
        }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (idx * 4 + e < L) ? v[i][e] * inv : 0.f;
            s4[idx] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

void attn_softmax(float* S, int nz, int nh, int Lq, int L, int ldS, const uint8_t* key_mask, int mask_ld,
                  float temperature, cudaStream_t st) {
    const long long rows = (long long)nz * Lq;
    if (rows == 0) return;
    ZVX_REQUIRE(rows < 2147483647LL, "attn_softmax: too many rows");
    const int n4 = ldS / 4;
    const bool vec = (ldS % 4 == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && n4 <= 32 * 16;
    if (vec) {
        const unsigned grid = (unsigned)((rows + 7) / 8);
#define ZVX_SOFTMAX(NV) attn_softmax_warp_kernel<NV><<<grid, 256, 0, st>>>(S, nh, Lq, L, ldS, key_mask, mask_ld, temperature, rows)
        if (n4 <= 32) ZVX_SOFTMAX(1);
        else if (n4 <= 64) ZVX_SOFTMAX(2);
        else if (n4 <= 128) ZVX_SOFTMAX(4);
        else if (n4 <= 256) ZVX_SOFTMAX(8);
        else ZVX_SOFTMAX(16);
#undef ZVX_SOFTMAX
    } else {
        attn_softmax_kernel<<<(unsigned)rows, 128, 0, st>>>(S, nh, Lq, L, ldS, key_mask, mask_ld, temperature);
    }
    ZVX_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
__global__ void add_batch_vector_kernel(float* __restrict__ x, const float* __restrict__ v, long long n4, int TC4,
                                        int C4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int b = (int)(i / TC4);
    const int c4 = (int)(i % C4);
    float4 a = reinterpret_cast<float4*>(x)[i];
    const float4 s = __ldg(reinterpret_cast<const float4*>(v) + (long long)b * C4 + c4);
    a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
    reinterpret_cast<float4*>(x)[i] = a;
}

void add_batch_vector(float* x, const float* v, int B, int T, int C, cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0, "add_batch_vector: C % 4");
    const long long n4 = (long long)B * T * C / 4;
    if (n4 == 0) return;
    add_batch_vector_kernel<<<cdiv(n4, 256), 256, 0, st>>>(x, v, n4, T * C / 4, C / 4);
    ZVX_POST_LAUNCH();
}

__global__ void bucket_embed_add_kernel(float* __restrict__ x, const float* __restrict__ pred,
                                        const float* __restrict__ table, int rows, int C, int n_bins,
                                        int32_t* __restrict__ bucket_out) {
    const int row = blockIdx.x;
    if (row >= rows) return;
    // torch.round = round-half-to-even = rintf in the default rounding mode; the product is a separate fp32
    // multiply in the reference, so keep it un-fused.
    const float scaled = __fmul_rn(pred[row], (float)(n_bins - 1));
    float r = rintf(scaled);
    r = fminf(fmaxf(r, 0.f), (float)(n_bins - 1));
    const int bkt = (int)r;
    if (bucket_out && threadIdx.x == 0) bucket_out[row] = bkt;
    const float4* e = reinterpret_cast<const float4*>(table + (long long)bkt * C);
    float4* o = reinterpret_cast<float4*>(x + (long long)row * C);
    for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) {
        float4 a = o[c4];
        const float4 b = __ldg(e + c4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        o[c4] = a;
    }
}

void bucket_embed_add(float* x, const float* pred, const float* table, int rows, int C, int n_bins,
                      int32_t* bucket_out, cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0, "bucket_embed_add: C % 4");
    if (rows == 0) return;
    bucket_embed_add_kernel<<<rows, 128, 0, st>>>(x, pred, table, rows, C, n_bins, bucket_out);
    ZVX_POST_LAUNCH();
}

__global__ void duration_round_kernel(const float* __restrict__ log_d, const int32_t* __restrict__ forced,
                                      int32_t* __restrict__ dur, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (forced) {
        dur[i] = forced[i];
    } else {
        float d = rintf(__fsub_rn(expf(log_d[i]), 1.f));
        d = fminf(fmaxf(d, 0.f), 1.0e6f);
        dur[i] = (int32_t)d;
    }
}

void duration_round(const float* log_d, const int32_t* forced, int32_t* dur, int n, cudaStream_t st) {
    if (n == 0) return;
    duration_round_kernel<<<cdiv(n, 256), 256, 0, st>>>(log_d, forced, dur, n);
    ZVX_POST_LAUNCH();
}

// one block per utterance; chunked block scan
__global__ void __launch_bounds__(256) duration_scan_kernel(const int32_t* __restrict__ dur, int T,
                                                            int32_t* __restrict__ cum, int64_t* __restrict__ mel_len) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int wsum[8];
    __shared__ int carry_s;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += 256) {
        const int t = t0 + tid;
        int v = (t < T) ? max(dur[(long long)b * T + t], 0) : 0;  // max(int(d), 0)  (fs2.py:452)
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += n;
        }
        if (lane == 31) wsum[wid] = s;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < wid; ++w) woff += wsum[w];
        const int carry = carry_s;
        if (t < T) cum[(long long)b * T + t] = carry + woff + s;
        __syncthreads();
        if (tid == 255) carry_s = carry + woff + s;
        __syncthreads();
    }
    if (tid == 0 && mel_len) mel_len[b] = (int64_t)carry_s;
}

void duration_scan(const int32_t* dur, int B, int T, int32_t* cum, int64_t* mel_len, cudaStream_t st) {
    if (B == 0) return;
    duration_scan_kernel<<<B, 256, 0, st>>>(dur, T, cum, mel_len);
    ZVX_POST_LAUNCH();
}

// one warp per output frame: binary search of the run that covers it, then a coalesced 128-bit row copy
__global__ void __launch_bounds__(256) length_regulate_gather_kernel(const float* __restrict__ x,
                                                                     const int32_t* __restrict__ cum, int T, int C4,
                                                                     int frame0, int L_max, float* __restrict__ features,
                                                                     int32_t* __restrict__ src_index) {
    const int b = blockIdx.y;
    const int fo = blockIdx.x * 8 + (threadIdx.x >> 5);   // frame within the chunk [frame0, frame0 + L_max)
    const int lane = threadIdx.x & 31;
    if (fo >= L_max) return;
    const int f = frame0 + fo;
    const int32_t* c = cum + (long long)b * T;
    const int total = T > 0 ? c[T - 1] : 0;
    float4* o = reinterpret_cast<float4*>(features) + ((long long)b * L_max + fo) * C4;
    if (f >= total) {
        for (int c4 = lane; c4 < C4; c4 += 32) o[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src_index && lane == 0) src_index[(long long)b * L_max + fo] = -1;
        return;
    }
    int lo = 0, hi = T - 1;  // first i with cum[i] > f
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] > f) hi = mid; else lo = mid + 1;
    }
    const float4* src = reinterpret_cast<const float4*>(x) + ((long long)b * T + lo) * C4;
    for (int c4 = lane; c4 < C4; c4 += 32) o[c4] = __ldg(src + c4);
    if (src_index && lane == 0) src_index[(long long)b * L_max + fo] = lo;
}

void length_regulate_gather(const float* x, const int32_t* cum, int B, int T, int C, int frame0, int L_max,
                            float* features, int32_t* src_index, cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0, "length_regulate: C % 4");
    if (B == 0 || L_max == 0) return;
    dim3 grid(cdiv(L_max, 8), B);
    length_regulate_gather_kernel<<<grid, 256, 0, st>>>(x, cum, T, C / 4, frame0, L_max, features, src_index);
    ZVX_POST_LAUNCH();
}

__global__ void add_posenc_kernel(const float* __restrict__ x, const float* __restrict__ pos, long long n4, int LC4,
                                  float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + (i % LC4));
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    reinterpret_cast<float4*>(out)[i] = a;
}

void add_posenc(const float* x, const float* pos, int B, int L, int C, float* out, cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0, "add_posenc: C % 4");
    const long long n4 = (long long)B * L * C / 4;
    if (n4 == 0) return;
    add_posenc_kernel<<<cdiv(n4, 256), 256, 0, st>>>(x, pos, n4, L * C / 4, out);
    ZVX_POST_LAUNCH();
}

__global__ void mask_from_lengths_kernel(const int64_t* __restrict__ mel_len, int B, int L, uint8_t* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * L) return;
    const int b = (int)(i / L), l = (int)(i % L);
    mask[i] = (l >= mel_len[b]) ? 1 : 0;
}

void mask_from_lengths(const int64_t* mel_len, int B, int L, uint8_t* mask, cudaStream_t st) {
    if ((long long)B * L == 0) return;
    mask_from_lengths_kernel<<<cdiv((long long)B * L, 256), 256, 0, st>>>(mel_len, B, L, mask);
    ZVX_POST_LAUNCH();
}

// 32x32 smem-tiled transpose [B,L,C] -> [B,C,L]
__global__ void transpose_mel_kernel(const float* __restrict__ in, const uint8_t* __restrict__ mask, int zero_masked,
                                     int L, int C, float* __restrict__ out, float* __restrict__ inplace) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int l = l0 + i, c = c0 + tx;
        float v = 0.f;
        if (l < L && c < C) {
            const long long idx = ((long long)b * L + l) * C + c;
            v = in[idx];
            if (zero_masked && mask && mask[(long long)b * L + l]) {
                v = 0.f;
                if (inplace) inplace[idx] = 0.f;
            }
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    if (!out) return;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, l = l0 + tx;
        if (l < L && c < C) out[((long long)b * C + c) * L + l] = tile[tx][i];
    }
}

void transpose_mel(const float* mel_BLC, const uint8_t* mask, int zero_masked, int B, int L, int C, float* mel_BCL,
                   float* mel_BLC_inplace, cudaStream_t st) {
    if (B == 0 || L == 0) return;
    dim3 grid(cdiv(L, 32), cdiv(C, 32), B);
    transpose_mel_kernel<<<grid, dim3(32, 8), 0, st>>>(mel_BLC, mask, zero_masked, L, C, mel_BCL, mel_BLC_inplace);
    ZVX_POST_LAUNCH();
}

}  // namespace zvx

namespace zvx {

// ------------------------------------------------------------------------------------------------
// StyleTTS decoder helpers (styletts.py): InstanceNorm1d over time on channel-last activations
// ------------------------------------------------------------------------------------------------
// mean / 1/sqrt(biased var + eps) per (utterance, channel) over the L rows of x[b] ([L, ld], channels [0, C)).
// One block per (32-channel group, utterance): 32 channels x 8 row lanes; two passes (mean, then centred squares) like
// ATen's instance_norm.
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const float* __restrict__ x, int L, int C, int ld,
                                                             float eps, float* __restrict__ mean, float* __restrict__ rstd) {
    __shared__ float red[8][33];
    __shared__ float mu_s[32];
    const int b = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const bool ok = c < C;
    const float* p = x + (long long)b * L * ld + c;
    float s = 0.f;
    if (ok)
        for (int l = threadIdx.y; l < L; l += 8) s += p[(long long)l * ld];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        mu_s[threadIdx.x] = t / (float)L;
    }
    __syncthreads();
    const float mu = mu_s[threadIdx.x];
    float q = 0.f;
    if (ok)
        for (int l = threadIdx.y; l < L; l += 8) {
            const float d = p[(long long)l * ld] - mu;
            q = fmaf(d, d, q);
        }
    __syncthreads();
    red[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        mean[(long long)b * C + c] = mu;
        rstd[(long long)b * C + c] = rsqrtf(t / (float)L + eps);
    }
}

void instnorm_stats(const float* x, int B, int L, int C, int ld, float eps, float* mean, float* rstd, cudaStream_t st) {
    if (B == 0 || C == 0) return;
    instnorm_stats_kernel<<<dim3(cdiv(C, 32), B), dim3(32, 8), 0, st>>>(x, L, C, ld, eps, mean, rstd);
    ZVX_POST_LAUNCH();
}

// y[b,l,c] = act( (x - mean) * rstd * gamma + beta ),  act = leaky-ReLU(slope) (slope 1 = identity)
//   affine InstanceNorm1d : gamma = g[c], beta = bt[c]                        (g_bs = 0)
//   AdaIN1d               : gamma = 1 + h[b, c], beta = h[b, C + c]           (styletts.py:87-92; g = h, bt = h + C, g_bs = 2C)
__global__ void instnorm_apply_kernel(const float4* __restrict__ x, int ld4, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, const float* __restrict__ g,
                                      const float* __restrict__ bt, int g_bs, float g_add, float slope, int L, int C4,
                                      long long total, float4* __restrict__ y, int ldy4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c4 = (int)(i % C4);
    const long long row = i / C4;
    const int b = (int)(row / L);
    const float4 v = __ldg(x + row * ld4 + c4);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + (long long)b * C4 * 4) + c4);
    const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd + (long long)b * C4 * 4) + c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(g + (long long)b * g_bs) + c4);
    const float4 be = __ldg(reinterpret_cast<const float4*>(bt + (long long)b * g_bs) + c4);
    auto f = [&](float xv, float m, float r, float gg, float bb) {
        const float o = (g_add + gg) * ((xv - m) * r) + bb;
        return o > 0.f ? o : o * slope;
    };
    y[row * ldy4 + c4] = make_float4(f(v.x, mu.x, rs.x, ga.x, be.x), f(v.y, mu.y, rs.y, ga.y, be.y),
                                     f(v.z, mu.z, rs.z, ga.z, be.z), f(v.w, mu.w, rs.w, ga.w, be.w));
}

void instnorm_apply(const float* x, int ld, const float* mean, const float* rstd, const float* g, const float* bt,
                    int g_bs, float g_add, float slope, int B, int L, int C, float* y, int ldy, cudaStream_t st) {
    ZVX_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ldy % 4 == 0 && g_bs % 4 == 0, "instnorm_apply: sizes must be multiples of 4");
    const long long total = (long long)B * L * (C / 4);
    if (total == 0) return;
    instnorm_apply_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), ld / 4, mean, rstd, g, bt,
                                                             g_bs, g_add, slope, L, C / 4, total,
                                                             reinterpret_cast<float4*>(y), ldy / 4);
    ZVX_POST_LAUNCH();
}

}  // namespace zvx
