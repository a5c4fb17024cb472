// HiFi-GAN MRF residual block, fully fused, on the tensor cores for the narrow stages (C = 8, 16, 32 channels).
//
// One CTA computes, for one tile of consecutive samples of one utterance, a whole ResBlock1
//     for d in dilations:  x = x + conv_{k,1}( lrelu( conv_{k,d}( lrelu(x) ) ) )          (hifigan.py:49-56)
// (or ResBlock2: x = x + conv_{k,d}(lrelu(x)), hifigan.py:80-84) and the MRF bookkeeping of Generator.forward
// (xs += resblock(x); x = xs / num_kernels, hifigan.py:119-125): the activation tile is read from HBM once and the
// result written once; the 2*nd - 1 intermediates never leave the SM.
//
// Shared-memory tiles (channel-last activations [B][T][C] in HBM):
//     X[cq][row][4 floats]   fp32 residual stream, row r <-> sample t0 - H + r (H = total one-sided halo of all convs)
//     A[cq][row][4 floats]   TF32 MMA operand, "16-byte-chunk major" = canonical K-major / no-swizzle UMMA layout
//                            (8-row x 16-byte core matrices, SBO = 128 B, LBO = rows*16 B)
// Rows of A are 16 bytes apart in every chunk plane, so filter tap j of a dilated conv is the SAME tile with the
// descriptor start address advanced by j*dilation rows: k accumulating tcgen05.mma per 8 input channels
// (M = 128 samples, N = C_out, fp32 accumulators in TMEM).  Each conv consumes a one-sided halo h = (k-1)/2*dil, so
// after conv s the operand row i stands for sample t0 - H + off_{s+1} + i with off_{s+1} = off_s + h_s; the epilogue
// writes the next operand IN PLACE over A (every MMA of the step has completed) and keeps X in fixed coordinates.
// Weights are streamed per conv step with cp.async while the previous step's epilogue runs.
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace zvx {

namespace {

constexpr int NT = 256;              // 8 warps: warp w works on TMEM lane quarter w % 4, M tiles (w / 4), (w / 4) + 2, ...
constexpr int MAX_STEPS = VocResArgs::MAX_STEPS;

struct ResPlan {
    int nsteps;
    int H;                     // total one-sided halo
    int TT;                    // valid output samples per CTA
    int R0, Rx, Rp;            // rows loaded; allocated rows of X / of A
    int Np;                    // MMA N (C_out padded to >= 16)
    int h[MAX_STEPS];          // one-sided halo of step s
    int off1[MAX_STEPS];       // off_{s+1}
    int rows_out[MAX_STEPS];   // R_{s+1}: valid output rows of step s
    int mt[MAX_STEPS];         // 128-row MMA tiles of step s
    int tmem_cols;
    uint32_t offX, offA, offW, offBar;
    uint32_t idesc;
    int smem_bytes;
    int ctas_per_sm;
};

__device__ __forceinline__ float lrelu(float v, float s) { return v > 0.f ? v : v * s; }

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

template <int C>
__global__ void __launch_bounds__(NT) voc_resblock_kernel(const VocResArgs a, const ResPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sb = smem_u32(smem);
    const uint32_t sA = sb + p.offA, sW = sb + p.offW;
    const uint32_t bar = sb + p.offBar, slot = bar + 8;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.offBar + 8);
    float4* Xs = reinterpret_cast<float4*>(smem + p.offX);
    constexpr int CQ = C / 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * p.TT;
    const int n16 = a.k * CQ * p.Np;   // 16-byte chunks of one conv's weight image

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, (uint32_t)p.tmem_cols);

    // weights of step 0 (global image == shared image [tap][cq][n][4])
    for (int i = tid; i < n16; i += NT) cp_async16(sW + (uint32_t)i * 16u, a.steps[0].w + (long long)i * 4);

    // input tile: raw -> X, lrelu + TF32 rounding -> A; samples outside [0, T) are the convs' zero padding
    {
        const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
        const int tA = t0 - p.H;
        const int total = p.R0 * CQ;
        for (int base = 0; base < total; base += NT * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT + tid;
                const int row = idx / CQ, cq = idx - row * CQ;
                const int t = tA + row;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total && t >= 0 && t < a.T) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT + tid;
                if (idx < total) {
                    const int row = idx / CQ, cq = idx - row * CQ;
                    Xs[cq * p.Rx + row] = v[u];
                    float4 o;
                    o.x = rn_tf32(lrelu(v[u].x, a.in_slope)); o.y = rn_tf32(lrelu(v[u].y, a.in_slope));
                    o.z = rn_tf32(lrelu(v[u].z, a.in_slope)); o.w = rn_tf32(lrelu(v[u].w, a.in_slope));
                    st_shared_v4(sA + (uint32_t)((cq * p.Rp + row) * 16), o);
                }
            }
        }
        // operand rows past the loaded ones only feed discarded outputs: keep them finite
        const int extra = (p.Rp - p.R0) * CQ;
        for (int idx = tid; idx < extra; idx += NT) {
            const int row = p.R0 + idx / CQ, cq = idx % CQ;
            st_shared_v4(sA + (uint32_t)((cq * p.Rp + row) * 16), make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    const int q = warp & 3, g = warp >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

    for (int s = 0; s < p.nsteps; ++s) {
        const int dil = a.steps[s].dil;
        const int mt = p.mt[s];
        if (warp == 0 && elect_one()) {
            // all taps of the conv for every M tile: D[m] (+)= A[rows m*128 + j*dil ...] x W[j]
            const uint64_t da0 = nosw_desc(sA, (uint32_t)p.Rp), db0 = nosw_desc(sW, (uint32_t)p.Np);
            for (int m = 0; m < mt; ++m) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(m * p.Np);
                for (int j = 0; j < a.k; ++j) {
#pragma unroll
                    for (int pp = 0; pp < C / 8; ++pp) {
                        // descriptor address field is in 16-byte units = rows of a chunk plane
                        const uint64_t da = da0 + (uint64_t)((pp * 2) * p.Rp + m * 128 + j * dil);
                        const uint64_t db = db0 + (uint64_t)((j * CQ + pp * 2) * p.Np);
                        umma_tf32(d_tmem, da, db, p.idesc, (j | pp) ? 1u : 0u);
                    }
                }
            }
            umma_commit(bar);
        }
        mbar_wait(bar, (uint32_t)(s & 1));
        tc_fence_after();

        // the weight buffer is free again: stream the next conv's weights behind this step's epilogue
        if (s + 1 < p.nsteps)
            for (int i = tid; i < n16; i += NT) cp_async16(sW + (uint32_t)i * 16u, a.steps[s + 1].w + (long long)i * 4);

        const float* __restrict__ bias = a.steps[s].b;
        const int kind = a.steps[s].kind;
        const bool last = (s == p.nsteps - 1);
        const int rows_out = p.rows_out[s];
        const int tbase = t0 - p.H + p.off1[s];
        for (int m = g; m < mt; m += 2) {
            uint32_t v[C];
            __syncwarp();
            tmem_ld_c<C>(trow + (uint32_t)(m * p.Np), v);
            tmem_wait_ld();
            const int i = m * 128 + q * 32 + lane;
            const int t = tbase + i;
            const bool inside = (t >= 0) && (t < a.T) && (i < rows_out);
            if (kind == 0) {
                // first conv of a pair: bias, lrelu, TF32 -> next operand (zero outside the utterance: conv2's padding)
#pragma unroll
                for (int cq = 0; cq < CQ; ++cq) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inside) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cq);
                        o.x = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 0]) + bb.x, a.mid_slope));
                        o.y = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 1]) + bb.y, a.mid_slope));
                        o.z = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 2]) + bb.z, a.mid_slope));
                        o.w = rn_tf32(lrelu(__uint_as_float(v[cq * 4 + 3]) + bb.w, a.mid_slope));
                    }
                    st_shared_v4(sA + (uint32_t)((cq * p.Rp + i) * 16), o);
                }
            } else if (!last) {
                // residual step: x += conv + bias (fp32, in X), next operand = lrelu(x) in TF32
                const int xr = p.off1[s] + i;
#pragma unroll
                for (int cq = 0; cq < CQ; ++cq) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inside) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cq);
                        float4 xo = Xs[cq * p.Rx + xr];
                        xo.x += __uint_as_float(v[cq * 4 + 0]) + bb.x;
                        xo.y += __uint_as_float(v[cq * 4 + 1]) + bb.y;
                        xo.z += __uint_as_float(v[cq * 4 + 2]) + bb.z;
                        xo.w += __uint_as_float(v[cq * 4 + 3]) + bb.w;
                        Xs[cq * p.Rx + xr] = xo;
                        o.x = rn_tf32(lrelu(xo.x, a.in_slope)); o.y = rn_tf32(lrelu(xo.y, a.in_slope));
                        o.z = rn_tf32(lrelu(xo.z, a.in_slope)); o.w = rn_tf32(lrelu(xo.w, a.in_slope));
                    }
                    st_shared_v4(sA + (uint32_t)((cq * p.Rp + i) * 16), o);
                }
            } else if (inside) {
                // last step: x += conv + bias, then the MRF bookkeeping and the store (channel-last rows)
                const int xr = p.off1[s] + i;   // == H + i, t == t0 + i
                const long long roff = (long long)t * C;
                const float* __restrict__ ain = a.acc_in ? a.acc_in + (long long)b * a.acc_in_bs + roff : nullptr;
                float* __restrict__ op = a.out + (long long)b * a.out_bs + roff;
#pragma unroll
                for (int cq = 0; cq < CQ; ++cq) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cq);
                    const float4 xo = Xs[cq * p.Rx + xr];
                    float4 y;
                    y.x = xo.x + (__uint_as_float(v[cq * 4 + 0]) + bb.x);
                    y.y = xo.y + (__uint_as_float(v[cq * 4 + 1]) + bb.y);
                    y.z = xo.z + (__uint_as_float(v[cq * 4 + 2]) + bb.z);
                    y.w = xo.w + (__uint_as_float(v[cq * 4 + 3]) + bb.w);
                    float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ain) sacc = *(reinterpret_cast<const float4*>(ain) + cq);
                    y.x = fmaf(y.x, a.out_scale, sacc.x); y.y = fmaf(y.y, a.out_scale, sacc.y);
                    y.z = fmaf(y.z, a.out_scale, sacc.z); y.w = fmaf(y.w, a.out_scale, sacc.w);
                    reinterpret_cast<float4*>(op)[cq] = make_float4(lrelu(y.x, a.out_slope), lrelu(y.y, a.out_slope),
                                                                   lrelu(y.z, a.out_slope), lrelu(y.w, a.out_slope));
                }
            }
        }
        cp_async_wait_all();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int pad_rows(int rows, int C) {
    // store pattern of the tile loader: a quarter-warp writes the C/4 chunks of 32/C consecutive rows; the chunk planes
    // are rows*16 B apart, so rows = 8/(C/4) (mod 8) spreads the 8 writes over all 32 banks
    const int want = 8 / (C / 4);
    int r = rows;
    while ((r & 7) != want) ++r;
    return r;
}

bool make_plan(const VocResArgs& a, ResPlan* out) {
    const int C = a.C, CQ = C / 4, k = a.k, ns = a.nsteps;
    if (ns < 1 || ns > MAX_STEPS) return false;
    ResPlan best{};
    bool found = false;
    // prefer two co-resident CTAs per SM (their phases overlap) when the tile stays efficient, else one big CTA
    for (int pass = 0; pass < 2 && !found; ++pass) {
        const int limit = pass == 0 ? 112 * 1024 : 226 * 1024;
        for (int mt0 = 8; mt0 >= 1 && !found; --mt0) {
            ResPlan p{};
            p.nsteps = ns;
            p.Np = std::max(C, 16);
            if (mt0 * p.Np > 512) continue;
            int H = 0;
            for (int s = 0; s < ns; ++s) { p.h[s] = (k - 1) / 2 * a.steps[s].dil; H += p.h[s]; }
            p.H = H;
            const int R1 = 128 * mt0;
            p.R0 = R1 + 2 * p.h[0];
            p.TT = p.R0 - 2 * H;
            if (p.TT < 32) continue;
            int cols = 32;
            while (cols < mt0 * p.Np) cols <<= 1;
            if (pass == 0 && (p.TT * 4 < R1 * 3 || cols > 256)) continue;   // < 75 % useful rows, or more than half the TMEM
            int rows = p.R0, off = 0, need = p.R0;
            for (int s = 0; s < ns; ++s) {
                rows -= 2 * p.h[s];
                off += p.h[s];
                p.rows_out[s] = rows;
                p.off1[s] = off;
                p.mt[s] = cdiv(rows, 128);
                need = std::max(need, 128 * p.mt[s] + 2 * p.h[s]);
            }
            p.Rx = pad_rows(p.R0, C);
            p.Rp = pad_rows(need, C);
            uint32_t o = 0;
            p.offX = o; o += (uint32_t)(CQ * p.Rx * 16);
            p.offA = o; o += (uint32_t)(CQ * p.Rp * 16);
            p.offW = o; o += (uint32_t)(k * CQ * p.Np * 16);
            p.offBar = o; o += 16;
            p.smem_bytes = (int)o + 128;
            p.tmem_cols = cols;
            p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            p.ctas_per_sm = pass == 0 ? 2 : 1;
            if (p.smem_bytes <= limit) { best = p; found = true; }
        }
    }
    if (found) *out = best;
    return found;
}

template <int C>
void launch(const VocResArgs& a, const ResPlan& p, cudaStream_t st) {
    static int attr_done = 0;
    if (!attr_done) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(voc_resblock_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = 1;
    }
    dim3 grid(cdiv(a.T, p.TT), a.B);
    voc_resblock_kernel<C><<<grid, NT, p.smem_bytes, st>>>(a, p);
    ZVX_POST_LAUNCH();
}

}  // namespace

bool voc_resblock_supported(int C, int k, const int* dils, int nd, bool pair) {
    if (!(C == 8 || C == 16 || C == 32) || (k & 1) == 0 || k < 1 || nd < 1 || nd * (pair ? 2 : 1) > MAX_STEPS) return false;
    VocResArgs a;
    a.C = C; a.k = k; a.nsteps = 0;
    for (int i = 0; i < nd; ++i) {
        if (dils[i] < 1) return false;
        if (pair) { a.steps[a.nsteps].dil = dils[i]; a.steps[a.nsteps++].kind = 0; a.steps[a.nsteps].dil = 1; a.steps[a.nsteps++].kind = 1; }
        else { a.steps[a.nsteps].dil = dils[i]; a.steps[a.nsteps++].kind = 1; }
    }
    ResPlan p;
    return make_plan(a, &p);
}

void voc_resblock_tc(const VocResArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return;
    ZVX_REQUIRE(a.C == 8 || a.C == 16 || a.C == 32, "voc_resblock_tc: C must be 8, 16 or 32");
    ZVX_REQUIRE((a.k & 1) == 1 && a.x && a.out && a.nsteps >= 1 && a.nsteps <= MAX_STEPS, "voc_resblock_tc: bad arguments");
    ZVX_REQUIRE(a.steps[a.nsteps - 1].kind == 1, "voc_resblock_tc: the last step must be a residual step");
    ZVX_REQUIRE(a.B <= 65535, "voc_resblock_tc: batch too large");
    ResPlan p;
    ZVX_REQUIRE(make_plan(a, &p), "voc_resblock_tc: tile does not fit shared memory");
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
