// HiFi-GAN residual-block convolutions on the tensor cores for the narrow stages (C = 8, 16, 32 channels):
// one CTA computes, for one tile of consecutive samples of one utterance, the fused pair of ResBlock1
//     x_new = x + conv_{k,1}( lrelu( conv_{k,d}( lrelu(x) ) ) )            (hifigan.py:49-56)
// (or the single dilated conv of ResBlock2, hifigan.py:80-84) without the intermediate ever leaving the SM.
//
// Layout.  Activations are channel-last [B][T][C] in HBM.  In shared memory a tile is stored "16-byte-chunk major":
//     A[cq][row][4 floats],   cq = channel / 4, row = sample
// which is the canonical K-major / no-swizzle UMMA layout with SBO = 128 B (8 rows x 16 B core matrices) and
// LBO = rows*16 B.  Rows are 16 bytes apart for every chunk, so the operand of filter tap j is the SAME tile with the
// descriptor start address advanced by j*dilation rows: the k taps of a dilated Conv1d read one staged copy of the
// input (k accumulating tcgen05.mma per 8 input channels, M = 128 samples, N = C_out, fp32 accumulators in TMEM).
// The first conv's accumulators are read back with tcgen05.ld, biased, leaky-ReLU'd, rounded to TF32 and written
// into a second tile of the same layout, which the second conv consumes directly; its epilogue adds bias and the
// residual and writes channel-last rows (optionally the MRF mean accumulation xs/num_kernels, hifigan.py:119-125,
// and a leaky-ReLU'd copy for the next upsampler).
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace zvx {

namespace {

constexpr int NT = 128;

struct VocPlan {
    int m1;            // 128-row MMA tiles per CTA
    int TT;            // valid output samples per CTA
    int h1, h2;        // halo (one side) of conv1 / conv2 in samples
    int R1, R1p;       // rows of the input tile (needed / allocated)
    int R2, R2p;       // rows of the intermediate tile
    int Np;            // MMA N (C_out padded to >= 16)
    int tmem_cols;
    uint32_t offA1, offA2, offW1, offW2, offBar;
    uint32_t idesc;
    int smem_bytes;
};

__device__ __forceinline__ float lrelu(float v, float s) { return v > 0.f ? v : v * s; }

// K-major, no-swizzle matrix descriptor: 8-row x 16-byte core matrices, rows 16 B apart (SBO = 128 B),
// the two 16-byte K chunks of one MMA `lbo16` 16-byte units apart.
__device__ __forceinline__ uint64_t nosw_desc(uint32_t saddr, uint32_t lbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo16 & 0x3FFFu) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <int C>
__device__ __forceinline__ void tmem_ld_c(uint32_t taddr, uint32_t* v) {
    if constexpr (C == 8) tmem_ld8(taddr, v);
    else if constexpr (C == 16) tmem_ld16(taddr, v);
    else { tmem_ld16(taddr, v); tmem_ld16(taddr + 16, v + 16); }
}

// All taps of one conv for every M tile: D[m] (+)= A[rows m*128 + j*dil ...] x W[j]
template <int C>
__device__ __forceinline__ void issue_conv(uint32_t sA, int Rp, uint32_t sW, int k, int dil, const VocPlan& p,
                                           uint32_t tmem_base) {
    constexpr int CQ = C / 4;
    for (int m = 0; m < p.m1; ++m) {
        for (int j = 0; j < k; ++j) {
#pragma unroll
            for (int pp = 0; pp < C / 8; ++pp) {
                const uint32_t a_addr = sA + (uint32_t)(((pp * 2) * Rp + m * 128 + j * dil) * 16);
                const uint32_t b_addr = sW + (uint32_t)(((j * CQ + pp * 2) * p.Np) * 16);
                umma_tf32(tmem_base + (uint32_t)(m * p.Np), nosw_desc(a_addr, (uint32_t)Rp), nosw_desc(b_addr, (uint32_t)p.Np),
                          p.idesc, (j | pp) ? 1u : 0u);
            }
        }
    }
}

template <int C>
__global__ void __launch_bounds__(NT) voc_pair_kernel(const VocPairArgs a, const VocPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sb = smem_u32(smem);
    const uint32_t sA1 = sb + p.offA1, sA2 = sb + p.offA2, sW1 = sb + p.offW1, sW2 = sb + p.offW2;
    const uint32_t bar = sb + p.offBar, slot = bar + 8;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.offBar + 8);
    constexpr int CQ = C / 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * p.TT;
    const bool pair = (a.w2 != nullptr);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, (uint32_t)p.tmem_cols);

    // weights: global image == shared image ([tap][cq][n][4]); asynchronous 16-byte copies
    {
        const int n16 = a.k * CQ * p.Np;
        for (int i = tid; i < n16; i += NT) cp_async16(sW1 + (uint32_t)i * 16u, a.w1 + (long long)i * 4);
        if (pair)
            for (int i = tid; i < n16; i += NT) cp_async16(sW2 + (uint32_t)i * 16u, a.w2 + (long long)i * 4);
    }
    // input tile: lrelu + TF32 rounding on the way in; samples outside [0, T) are the conv's zero padding
    {
        const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
        const int tA = t0 - p.h2 - p.h1;
        const int total = p.R1 * CQ;
        for (int base = 0; base < total; base += NT * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT + tid;
                const int row = idx / CQ, cq = idx - row * CQ;
                const int t = tA + row;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total && t >= 0 && t < a.T)
                    v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT + tid;
                if (idx < total) {
                    const int row = idx / CQ, cq = idx - row * CQ;
                    float4 o;
                    o.x = rn_tf32(lrelu(v[u].x, a.in_slope)); o.y = rn_tf32(lrelu(v[u].y, a.in_slope));
                    o.z = rn_tf32(lrelu(v[u].z, a.in_slope)); o.w = rn_tf32(lrelu(v[u].w, a.in_slope));
                    st_shared_v4(sA1 + (uint32_t)((cq * p.R1p + row) * 16), o);
                }
            }
        }
        if (pair) {  // rows of the intermediate tile past the computed ones feed discarded outputs only: keep them finite
            const int extra = (p.R2 - 128 * p.m1) * CQ;
            for (int idx = tid; idx < extra; idx += NT) {
                const int row = 128 * p.m1 + idx / CQ, cq = idx % CQ;
                st_shared_v4(sA2 + (uint32_t)((cq * p.R2p + row) * 16), make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);

    if (tid == 0) {
        issue_conv<C>(sA1, p.R1p, sW1, a.k, a.d1, p, tmem_base);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();

    if (pair) {
        // epilogue 1: accumulators -> bias, lrelu, TF32 -> intermediate tile (zero outside the utterance: conv2's padding)
        for (int m = 0; m < p.m1; ++m) {
            uint32_t v[C];
            __syncwarp();
            tmem_ld_c<C>(trow + (uint32_t)(m * p.Np), v);
            tmem_wait_ld();
            const int i = m * 128 + warp * 32 + lane;
            const int t = t0 - p.h2 + i;
            const bool inside = (t >= 0) && (t < a.T);
#pragma unroll
            for (int cq = 0; cq < CQ; ++cq) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inside) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b1) + cq);
                    o.x = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 0]) + bb.x, a.mid_slope));
                    o.y = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 1]) + bb.y, a.mid_slope));
                    o.z = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 2]) + bb.z, a.mid_slope));
                    o.w = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 3]) + bb.w, a.mid_slope));
                }
                st_shared_v4(sA2 + (uint32_t)((cq * p.R2p + i) * 16), o);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            issue_conv<C>(sA2, p.R2p, sW2, a.k, 1, p, tmem_base);
            umma_commit(bar);
        }
        mbar_wait(bar, 1);
        tc_fence_after();
    }

    // final epilogue: + bias + residual, channel-last rows out
    {
        const float* __restrict__ bias = pair ? a.b2 : a.b1;
        const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
        for (int m = 0; m < p.m1; ++m) {
            uint32_t v[C];
            __syncwarp();
            tmem_ld_c<C>(trow + (uint32_t)(m * p.Np), v);
            tmem_wait_ld();
            const int o = m * 128 + warp * 32 + lane;
            const int t = t0 + o;
            if (o >= p.TT || t >= a.T) continue;
            const long long roff = (long long)t * C;
#pragma unroll
            for (int cq = 0; cq < CQ; ++cq) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cq);
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.residual) r = __ldg(reinterpret_cast<const float4*>(xb + roff) + cq);
                float4 y;
                y.x = __uint_as_float(v[cq * 4 + 0]) + bb.x + r.x;
                y.y = __uint_as_float(v[cq * 4 + 1]) + bb.y + r.y;
                y.z = __uint_as_float(v[cq * 4 + 2]) + bb.z + r.z;
                y.w = __uint_as_float(v[cq * 4 + 3]) + bb.w + r.w;
                if (a.out) reinterpret_cast<float4*>(a.out + (long long)b * a.out_bs + roff)[cq] = y;
                if (a.acc) {
                    float4* ap = reinterpret_cast<float4*>(a.acc + (long long)b * a.acc_bs + roff) + cq;
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!a.acc_init) s = *ap;
                    y.x = fmaf(y.x, a.acc_scale, s.x); y.y = fmaf(y.y, a.acc_scale, s.y);
                    y.z = fmaf(y.z, a.acc_scale, s.z); y.w = fmaf(y.w, a.acc_scale, s.w);
                    *ap = y;
                }
                if (a.act_out)
                    reinterpret_cast<float4*>(a.act_out + (long long)b * a.act_bs + roff)[cq] =
                        make_float4(lrelu(y.x, a.act_slope), lrelu(y.y, a.act_slope), lrelu(y.z, a.act_slope),
                                    lrelu(y.w, a.act_slope));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

int pad_rows(int rows, int C) {
    // store pattern of the tile loader: a quarter-warp writes the C/4 chunks of 32/C consecutive rows; the chunk
    // planes are rows*16 B apart, so rows = 8/(C/4) (mod 8) spreads the 8 writes over all 32 banks
    const int want = 8 / (C / 4);
    int r = rows;
    while ((r & 7) != want) ++r;
    return r;
}

bool make_plan(const VocPairArgs& a, VocPlan* out) {
    const bool pair = a.w2 != nullptr;
    const int C = a.C, CQ = C / 4;
    VocPlan best{};
    bool found = false;
    for (int pass = 0; pass < 2 && !found; ++pass) {
        const int limit = pass == 0 ? 110 * 1024 : 224 * 1024;
        for (int m1 = 4; m1 >= (pass == 0 ? 4 : 1); --m1) {
            VocPlan p{};
            p.m1 = m1;
            p.Np = std::max(C, 16);
            p.h1 = (a.k - 1) / 2 * a.d1;
            p.h2 = pair ? (a.k - 1) / 2 : 0;
            p.TT = 128 * m1 - 2 * p.h2;
            p.R1 = 128 * m1 + 2 * p.h1;
            p.R1p = pad_rows(p.R1, C);
            p.R2 = pair ? 128 * m1 + 2 * p.h2 : 0;
            p.R2p = p.R2;
            uint32_t off = 0;
            p.offA1 = off; off += (uint32_t)(CQ * p.R1p * 16);
            p.offA2 = off; off += (uint32_t)(CQ * p.R2p * 16);
            p.offW1 = off; off += (uint32_t)(a.k * CQ * p.Np * 16);
            p.offW2 = off; off += pair ? (uint32_t)(a.k * CQ * p.Np * 16) : 0u;
            p.offBar = off; off += 16;
            p.smem_bytes = (int)off + 128;
            int cols = 32;
            while (cols < m1 * p.Np) cols <<= 1;
            p.tmem_cols = cols;
            p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            if (p.smem_bytes <= limit && p.TT > 0) { best = p; found = true; break; }
        }
    }
    if (found) *out = best;
    return found;
}

template <int C>
void launch(const VocPairArgs& a, const VocPlan& p, cudaStream_t st) {
    static int attr_done = 0;
    if (!attr_done) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(voc_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = 1;
    }
    dim3 grid(cdiv(a.T, p.TT), a.B);
    voc_pair_kernel<C><<<grid, NT, p.smem_bytes, st>>>(a, p);
    ZVX_POST_LAUNCH();
}

}  // namespace

bool voc_pair_supported(int C, int k, int dil) {
    if (!(C == 8 || C == 16 || C == 32) || (k & 1) == 0 || k < 1 || dil < 1) return false;
    VocPairArgs a;
    a.C = C; a.k = k; a.d1 = dil;
    a.w2 = reinterpret_cast<const float*>(1);
    VocPlan p;
    return make_plan(a, &p);
}

void voc_pair_tc(const VocPairArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return;
    ZVX_REQUIRE(a.C == 8 || a.C == 16 || a.C == 32, "voc_pair_tc: C must be 8, 16 or 32");
    ZVX_REQUIRE((a.k & 1) == 1 && a.x && a.w1 && a.b1 && (a.out || a.acc), "voc_pair_tc: bad arguments");
    ZVX_REQUIRE(a.B <= 65535, "voc_pair_tc: batch too large");
    VocPlan p;
    ZVX_REQUIRE(make_plan(a, &p), "voc_pair_tc: tile does not fit shared memory");
    switch (a.C) {
        case 8: launch<8>(a, p, st); break;
        case 16: launch<16>(a, p, st); break;
        default: launch<32>(a, p, st); break;
    }
}

// [Cout][Cin][k] (PyTorch Conv1d) -> the kernel's shared-memory image [k][Cin/4][Np][4], TF32-rounded (nearest even)
std::vector<float> voc_pack_weight(const float* w, int cout, int cin, int k) {
    const int Np = std::max(cout, 16), CQ = cin / 4;
    std::vector<float> o((size_t)k * CQ * Np * 4, 0.f);
    for (int n = 0; n < cout; ++n)
        for (int ci = 0; ci < cin; ++ci)
            for (int j = 0; j < k; ++j) {
                float v = w[((size_t)n * cin + ci) * k + j];
                uint32_t u;
                memcpy(&u, &v, 4);
                u = (u + 0x0FFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
                memcpy(&v, &u, 4);
                o[(((size_t)j * CQ + ci / 4) * Np + n) * 4 + (ci & 3)] = v;
            }
    return o;
}

}  // namespace zvx
