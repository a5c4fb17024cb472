// HiFi-GAN ResBlock1 (hifigan.py:25-56) on the tensor cores for the narrow stages (C = 8 / 16 / 32 channels): one launch =
// one whole residual block  x -> [conv_{k,d} -> conv_{k,1} -> + x] x n_dilations  plus the MRF mean (hifigan.py:119-125),
// activations resident in shared memory, the fp32 residual stream in registers.
//
// Formulation (second generation of voc_poly.cu).  tcgen05.mma costs ~80-100 cycles per instruction in the K-major
// no-swizzle layout whatever N is, so the kernel is bound by the NUMBER of MMAs.  As in voc_poly.cu one accumulator row holds
// P = 128 / C output samples (N = P * C = 128) and the weights are Toeplitz-expanded; new here:
//
//  * Dilated convs cost what undilated ones cost.  A conv of dilation d only couples samples of equal residue mod d, so its
//    operand is laid out "residue-major": sample tau = d * (q*P + s) + rho  (rho = tau mod d, s = sub-index in its residue
//    class mod P, q = block) lives in phase block s, row n = q*d + rho.  Accumulator row n then holds the P samples
//    {d*(q*P + r) + rho, r < P} and the conv is an UNDILATED Toeplitz product over the sub-index:  K runs over the P + k - 1
//    offsets w of the sub-index instead of P + (k-1)*d sample offsets (C = 32, k = 11, d = 5: 56 MMAs instead of 176), the
//    A operand of offset w is phase block (w mod P) shifted by floor(w / P) * d rows — a descriptor start address — and
//    the weight image is the same dilation-free Toeplitz array for every dilation.  The epilogue of a step writes the next
//    operand directly in the layout of the next step's dilation (a scatter of 16-byte chunks by index arithmetic).
//  * Trimmed Toeplitz.  Offset w only reaches the sub-indices r in [w - c, w + c] (c = (k-1)/2): edge offsets issue MMAs
//    of N = (#r) * C columns at accumulator column r_lo * C, so the weight image needs the k tap planes only (plus one
//    zero plane each side when C = 8, to keep N a multiple of 16) — 45 KB instead of 70 KB at C = 32, k = 11.
//  * Two CTAs per SM wherever shared memory allows (everything but C = 32, k = 11): while one CTA's epilogue warps move an
//    accumulator TMEM -> registers -> shared memory, the other CTA's MMAs keep the tensor pipe busy.  One 128-column TMEM
//    accumulator per CTA; a step's MMAs cannot start before its operand is complete, so the issuing thread is simply lane 0
//    of epilogue warp 0 (8 warps per CTA, 128 registers per thread); weights single- or double-buffered as space permits.
//
// ResBlock2 stages and shapes outside this plan keep using voc_poly.cu / voc_res.cu.
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>

namespace zvx {

namespace {

constexpr int EPI_THREADS = 256;         // 8 loader / epilogue warps; lane 0 of warp 0 also issues the MMAs and streams the weights
constexpr int NT = EPI_THREADS;          // (no dedicated issuer warp: 2 x 8 warps per SM leave 128 registers per thread)
constexpr int MAX_STEPS = VocResArgs::MAX_STEPS;
constexpr int SMEM_PER_SM = 227 * 1024;

struct PairPlan {
    int nsteps, P, lgP, S, G, Rtot, R, TT, lo;
    int ld[MAX_STEPS];          // dilation of the layout the step's operand is stored in (= the conv's dilation)
    uint32_t mg[MAX_STEPS];     // ceil(2^20 / ld): tau / ld == (tau * mg) >> 20 for every tau of a tile (checked on the host)
    int n16[MAX_STEPS];         // 16-byte chunks of the step's weight image
    int ZC;                     // rows per 16-byte channel-chunk plane of a weight image (tap planes * C)
    int wbuf_bytes, nwbuf;
    int sched_off[MAX_STEPS + 1];
    int sched_total;
    uint32_t offA, offW, offZero, offBar, offSched;   // offZero: 256 zero rows of 16 B (operands of the accumulator-clearing MMA)
    int smem_bytes, ctas;
};

// leaky ReLU for slopes in (0, 1]: max(v, v*s)
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }
// fp32 -> tf32 round-to-nearest, ties away from zero (= cvt.rna.tf32.f32 for finite values) with two integer instructions
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float4 act4(float4 v, float s) {
    return make_float4(rna_tf32(lrelu(v.x, s)), rna_tf32(lrelu(v.y, s)), rna_tf32(lrelu(v.z, s)), rna_tf32(lrelu(v.w, s)));
}

// Row (16-byte units inside one channel-chunk plane) of sample tau in the residue-major layout of dilation d, or -1 when
// the layout does not hold that sample (tau < 0, or past the last row).
__device__ __forceinline__ int layout_row(int tau, int d, uint32_t mg, int P, int lgP, int S, int G) {
    if (tau < 0) return -1;
    const int v = (int)(((uint32_t)tau * mg) >> 20);      // tau / d
    const int rho = tau - v * d;
    const int n = (v >> lgP) * d + rho;
    return n < 128 ? (v & (P - 1)) * S + G + n : -1;
}

template <int C, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) voc_pair_kernel(const VocResArgs a, const PairPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sb = smem_u32(smem);
    const uint32_t sA = sb + p.offA;
    constexpr int CQ = C / 4, P = 128 / C, NPH = P / 2;
    // barriers: mma_done | w_full[2] | op_ready (operand of the next step complete: one arrival per epilogue warp)
    const uint32_t bars = sb + p.offBar;
    const uint32_t mma_bar = bars, op_bar = bars + 24u, slot = bars + 32u;
    auto w_bar = [&](int i) { return bars + 8u + 8u * (uint32_t)i; };
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.offBar + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int tbase = blockIdx.x * p.TT - p.lo;        // global sample index of tile sample tau = 0
    const int S = p.S, G = p.G, Rtot = p.Rtot, lgP = p.lgP;
#ifdef ZVX_DEBUG   // clock64 checkpoints of one CTA's thread 0 (ZVX_VOC_DBG=1): setup | load | per step: MMAs done, epilogue done
    const bool dbg = a.dbg && tid == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2;
    int dbg_i = 0;
#define ZVX_STAMP() do { if (dbg) a.dbg[dbg_i++] = clock64(); } while (0)
#else
#define ZVX_STAMP() do { } while (0)
#endif
    ZVX_STAMP();

    if (tid == 0) {
        mbar_init(mma_bar, 1);
        mbar_init(w_bar(0), 1);
        mbar_init(w_bar(1), 1);
        mbar_init(op_bar, EPI_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 128);
    // MMA schedule (host-built once per block shape): global {A offset | B offset << 16, D column | accumulate << 8 | N << 16}
    // -> shared {A descriptor low word, B descriptor low word, D column | accumulate << 8, instruction descriptor}
    {
        const uint2* __restrict__ gs = reinterpret_cast<const uint2*>(a.sched);
        uint4* ss = reinterpret_cast<uint4*>(smem + p.offSched);
        for (int i = tid; i < p.sched_total; i += NT) {
            const uint2 e = __ldg(gs + i);
            int st = 0;
            while (st + 1 < p.nsteps && i >= p.sched_off[st + 1]) ++st;
            const uint32_t sWs = sb + p.offW + (uint32_t)((p.nwbuf == 2 ? (st & 1) : 0) * p.wbuf_bytes);
            uint32_t alo = (((sA & 0x3FFFFu) >> 4) + (e.x & 0xFFFFu)) | ((uint32_t)Rtot << 16);
            uint32_t blo = (((sWs & 0x3FFFFu) >> 4) + (e.x >> 16)) | ((uint32_t)p.ZC << 16);
            if (e.y & 0x200u) alo = blo = (((sb + p.offZero) & 0x3FFFFu) >> 4) | (128u << 16);   // clear: 0 x 0, K-halves 128 rows apart
            const uint32_t n = e.y >> 16;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            ss[i] = make_uint4(alo, blo, e.y & 0x1FFu, idesc);
        }
    }
    // the whole operand tile starts as zeros: guard rows and layout rows no step writes only feed halo outputs, but must
    // stay finite
    for (int idx = tid; idx < CQ * Rtot; idx += NT) st_shared_v4(sA + (uint32_t)idx * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int idx = tid; idx < 256; idx += NT) st_shared_v4(sb + p.offZero + (uint32_t)idx * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    ZVX_STAMP();

    // ---- MMA issue + weight streaming: lane 0 of warp 0, between its epilogue duties ---------------------------------
    auto load_w = [&](int s) {
        const int buf = p.nwbuf == 2 ? (s & 1) : 0;
        const uint32_t bytes = (uint32_t)p.n16[s] * 16u;
        mbar_arrive_expect_tx(w_bar(buf), bytes);
        bulk_load_1d(sb + p.offW + (uint32_t)(buf * p.wbuf_bytes), a.steps[s].w_pair, bytes, w_bar(buf));
    };
    auto issue_step = [&](int s) {
        const int buf = p.nwbuf == 2 ? (s & 1) : 0;
        mbar_wait_spin(w_bar(buf), (uint32_t)((p.nwbuf == 2 ? (s >> 1) : s) & 1));
        mbar_wait_spin(op_bar, (uint32_t)(s & 1));   // operand of step s written (and the accumulator of step s-1 drained)
        tc_fence_after();
        // double-buffered weights: the other buffer was last read by the MMAs of step s-1, complete since the epilogue of
        // s-1 has run
        if (p.nwbuf == 2 && s + 1 < p.nsteps) load_w(s + 1);
        const uint4* tab = reinterpret_cast<const uint4*>(smem + p.offSched);
        constexpr uint64_t DESC_HI = ((uint64_t)(128 >> 4) | ((uint64_t)1 << 14)) << 32;   // SBO = 128 B, version 1
        auto issue = [&](const uint4 e) {
            umma_tf32(tmem_base + (e.z & 0xFFu), DESC_HI | e.x, DESC_HI | e.y, e.w, (e.z >> 8) & 1u);
        };
        const int i1 = p.sched_off[s + 1];
        int i = p.sched_off[s];
        for (; i + 4 <= i1; i += 4) {   // four entries in registers before the first MMA: MMAs go back to back
            const uint4 e0 = tab[i], e1 = tab[i + 1], e2 = tab[i + 2], e3 = tab[i + 3];
            issue(e0); issue(e1); issue(e2); issue(e3);
        }
        for (; i < i1; ++i) issue(tab[i]);
        umma_commit(mma_bar);
    };
    if (tid == 0) load_w(0);
    {
        // ================================================================ loader + epilogue warps (0..7)
        // Thread (q, lane, g) reads accumulator row n = 32q + lane, columns [64g, 64g + 64) = the NPH sub-indices
        // g*NPH + ri of that row, all C channels: 16 float4 (ri, cq).  In the d = 1 layout row n holds the samples
        // P*n .. P*n + P - 1: the fp32 residual stream of the thread's NPH consecutive samples stays in registers.
        const int q = warp & 3, g = warp >> 2;
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64);
        const int n = q * 32 + lane;
        const int tau0 = P * n + g * NPH;            // first owned sample (d = 1 ownership)
        float4 xo[16];
        {   // input tile: coalesced global loads staged raw in the d = 1 layout; every thread takes the samples it owns into
            // registers (the residual stream); then the operand lrelu(x) is written in the layout of the first step
            const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int idx = u * EPI_THREADS + tid;
                const int tau = idx / CQ, cq = idx - tau * CQ;
                const int t = tbase + tau;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t >= 0 && t < a.T) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int idx = u * EPI_THREADS + tid;
                const int tau = idx / CQ, cq = idx - tau * CQ;
                st_shared_v4(sA + (uint32_t)((cq * Rtot + (tau & (P - 1)) * S + G + (tau >> lgP)) * 16), v[u]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // the 8 loader warps only
            const float4* As = reinterpret_cast<const float4*>(smem + p.offA);
#pragma unroll
            for (int idx = 0; idx < 16; ++idx) {
                const int ri = idx / CQ, cq = idx % CQ;
                xo[idx] = As[cq * Rtot + (g * NPH + ri) * S + G + n];
            }
            const int d0 = p.ld[0];
            if (d0 != 1) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // scatter below: everyone has read
#pragma unroll
            for (int ri = 0; ri < NPH; ++ri) {
                const int row = d0 == 1 ? (g * NPH + ri) * S + G + n : layout_row(tau0 + ri, d0, p.mg[0], P, lgP, S, G);
#pragma unroll
                for (int cq = 0; cq < CQ; ++cq)
                    if (row >= 0) st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), act4(xo[ri * CQ + cq], a.in_slope));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(op_bar);
        }
        ZVX_STAMP();

        for (int s = 0; s < p.nsteps; ++s) {
            const int kind = a.steps[s].kind;
            const bool last = (s == p.nsteps - 1);
            const float4* __restrict__ bias = reinterpret_cast<const float4*>(a.steps[s].b);
            // sample of (this thread's row, sub-index g*NPH + ri) in the layout of THIS step: tau = ds*(qs*P + r) + rho
            const int ds = p.ld[s];
            int tau_first, tau_step;
            if (ds == 1) { tau_first = tau0; tau_step = 1; }
            else {
                const int qs = (int)(((uint32_t)n * p.mg[s]) >> 20), rho = n - qs * ds;
                tau_first = ds * (qs * P + g * NPH) + rho; tau_step = ds;
            }
            // layout of the operand this epilogue produces
            const int dn = last ? 1 : p.ld[s + 1];
            const uint32_t mgn = last ? 0u : p.mg[s + 1];
            if (warp == 0) {
                if (lane == 0) issue_step(s);
                __syncwarp();
            }
            ZVX_STAMP();
            mbar_wait(mma_bar, (uint32_t)(s & 1));   // suspending wait: leaves the issue slots to the co-resident CTA
            tc_fence_after();
            ZVX_STAMP();
            if (tid == 0 && p.nwbuf == 1 && s + 1 < p.nsteps) load_w(s + 1);   // single weight buffer: this step's MMAs have read it

            auto epilogue = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;   // 0: first conv of a pair, 1: residual step, 2: last step
                uint32_t vbuf[2][16];
                tmem_ld16(tcol, vbuf[0]);
#pragma unroll
                for (int idx = 0; idx < 16; ++idx) {
                    if ((idx & 3) == 0) {
                        __syncwarp();
                        tmem_wait_ld();
                        if (idx + 4 < 16) tmem_ld16(tcol + (uint32_t)(4 * (idx + 4)), vbuf[((idx >> 2) + 1) & 1]);
                        else tc_fence_before();   // orders the TMEM reads before the next step's MMAs overwrite the accumulator
                    }
                    const uint32_t* v = vbuf[(idx >> 2) & 1] + 4 * (idx & 3);
                    const int ri = idx / CQ, cq = idx % CQ;
                    const int tau = tau_first + ri * tau_step;
                    const int t = tbase + tau;
                    const bool inside = (t >= 0) && (t < a.T);
                    const float4 bb = __ldg(bias + cq);
                    float4 c4;
                    c4.x = __uint_as_float(v[0]) + bb.x; c4.y = __uint_as_float(v[1]) + bb.y;
                    c4.z = __uint_as_float(v[2]) + bb.z; c4.w = __uint_as_float(v[3]) + bb.w;
                    if constexpr (MODE == 0) {
                        // first conv of a pair: bias, lrelu, TF32 -> operand of the d = 1 conv (zero outside the utterance:
                        // that conv's padding)
                        const int row = (ds == 1) ? (g * NPH + ri) * S + G + n
                                                  : ((tau >> lgP) < 128 ? (tau & (P - 1)) * S + G + (tau >> lgP) : -1);
                        float4 o = act4(c4, a.mid_slope);
                        if (!inside) o = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row >= 0) st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), o);
                    } else if constexpr (MODE == 1) {
                        // residual step: x += conv + bias (fp32, registers), next operand = lrelu(x) in TF32, stored in the
                        // layout of the next step's dilation
                        float4 x4 = xo[idx];
                        x4.x = inside ? x4.x + c4.x : x4.x; x4.y = inside ? x4.y + c4.y : x4.y;
                        x4.z = inside ? x4.z + c4.z : x4.z; x4.w = inside ? x4.w + c4.w : x4.w;
                        xo[idx] = x4;   // stays 0 outside the utterance
                        const int row = (dn == 1) ? (g * NPH + ri) * S + G + n : layout_row(tau, dn, mgn, P, lgP, S, G);
                        if (row >= 0) st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), act4(x4, a.in_slope));
                    } else {
                        // last step: x += conv + bias, then the MRF bookkeeping and the store (channel-last rows)
                        if (inside && tau >= p.lo && tau < p.lo + p.TT) {
                            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (a.acc_in) sa = *(reinterpret_cast<const float4*>(a.acc_in + (long long)b * a.acc_in_bs + (long long)t * C) + cq);
                            float4 y;
                            y.x = xo[idx].x + c4.x; y.y = xo[idx].y + c4.y; y.z = xo[idx].z + c4.z; y.w = xo[idx].w + c4.w;
                            y.x = fmaf(y.x, a.out_scale, sa.x); y.y = fmaf(y.y, a.out_scale, sa.y);
                            y.z = fmaf(y.z, a.out_scale, sa.z); y.w = fmaf(y.w, a.out_scale, sa.w);
                            float* op = a.out + (long long)b * a.out_bs + (long long)t * C;
                            reinterpret_cast<float4*>(op)[cq] = make_float4(lrelu(y.x, a.out_slope), lrelu(y.y, a.out_slope),
                                                                           lrelu(y.z, a.out_slope), lrelu(y.w, a.out_slope));
                        }
                    }
                }
                if constexpr (MODE != 2) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(op_bar);
                }
            };
            if (kind == 0) epilogue(std::integral_constant<int, 0>{});
            else if (!last) epilogue(std::integral_constant<int, 1>{});
            else epilogue(std::integral_constant<int, 2>{});
            ZVX_STAMP();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

int up_to_mod8(int v, int want) {
    while ((v & 7) != want) ++v;
    return v;
}

int planes_of(int C, int k) { return k + (C == 8 ? 2 : 0); }   // tap planes of the trimmed Toeplitz image (+ a zero plane each side for C = 8)

bool make_plan(const VocResArgs& a, PairPlan* out) {
    const int C = a.C, CQ = C / 4, k = a.k, ns = a.nsteps, P = 128 / C, c = (k - 1) / 2;
    if (!(C == 8 || C == 16 || C == 32) || ns < 2 || ns > MAX_STEPS || (ns & 1) || (k & 1) == 0 || k < 3) return false;
    PairPlan p{};
    p.nsteps = ns; p.P = P; p.R = 128 * P;
    p.lgP = 0;
    while ((1 << p.lgP) < P) ++p.lgP;
    int lo = 0, hi = p.R, dmax = 1;
    for (int s = 0; s < ns; ++s) {
        const int d = a.steps[s].dil;
        // ResBlock1: even steps = first conv of a pair (kind 0, any dilation), odd steps = residual step with dilation 1
        // (its accumulator rows must hold the register-resident residual stream's samples)
        if (a.steps[s].kind != (s & 1) || d < 1 || d > 16 || ((s & 1) && d != 1)) return false;
        p.ld[s] = d;
        p.mg[s] = (uint32_t)(((1u << 20) + d - 1) / d);
        for (int tau = 0; tau < p.R + 64; ++tau)
            if ((int)(((uint32_t)tau * p.mg[s]) >> 20) != tau / d) return false;
        dmax = std::max(dmax, d);
        lo += c * d;
        hi = std::min(hi, (128 / d) * d * P) - c * d;   // the layout of dilation d holds floor(128/d) complete blocks of P*d samples
    }
    p.lo = lo;
    p.TT = hi - lo;
    if (p.TT < p.R / 2) return false;   // halo would dominate
    p.G = ((c + P - 1) / P) * dmax;     // row shift of an offset: floor(w / P) * d, |floor| <= ceil(c / P)
    p.S = up_to_mod8(128 + 2 * p.G, 1);
    p.Rtot = up_to_mod8(P * p.S, 8 / CQ == 8 ? 0 : 8 / CQ);
    if (CQ == 1) return false;
    const int planes = planes_of(C, k);
    p.ZC = planes * C;
    const int wbytes = CQ * p.ZC * 16;
    for (int s = 0; s < ns; ++s) p.n16[s] = CQ * p.ZC;
    p.wbuf_bytes = (int)round_up(wbytes, 128);
    // MMA schedule size: per step, one accumulator-clearing MMA, then the P + k - 1 offsets x C/8 k-chunks
    const int per_step = 1 + (P + k - 1) * (C / 8);
    for (int s = 0; s <= ns; ++s) p.sched_off[s] = s * per_step;
    p.sched_total = (ns * per_step + 1) & ~1;
    auto layout = [&](int nwbuf) {
        uint32_t o = 0;
        p.nwbuf = nwbuf;
        p.offA = o; o += (uint32_t)(CQ * p.Rtot * 16);
        p.offW = o; o += (uint32_t)(nwbuf * p.wbuf_bytes);
        p.offZero = o; o += 256 * 16;
        p.offBar = o; o += 48;
        o = (uint32_t)round_up(o, 16);
        p.offSched = o; o += (uint32_t)p.sched_total * 16;
        p.smem_bytes = (int)o + 128;
    };
    // two CTAs per SM when both fit (double-buffered weights first), else one CTA with double-buffered weights
    const int per_cta2 = (SMEM_PER_SM - 2 * 1024) / 2;
    layout(2);
    if (p.smem_bytes <= per_cta2) p.ctas = 2;
    else {
        layout(1);
        if (p.smem_bytes <= per_cta2) p.ctas = 2;
        else { layout(2); p.ctas = 1; if (p.smem_bytes > SMEM_PER_SM - 1024) { layout(1); } }
    }
    if (p.smem_bytes > SMEM_PER_SM - 1024) return false;
    *out = p;
    return true;
}

// The tcgen05.mma list of one block shape in issue order: x = A offset | B offset << 16 (16-byte units relative to the
// operand bases), y = accumulator column | accumulate << 8 | N << 16.
std::vector<uint2> build_schedule(const VocResArgs& a, const PairPlan& p) {
    const int C = a.C, P = p.P, k = a.k, c = (k - 1) / 2, e = (C == 8) ? 1 : 0;
    auto fdiv = [](int x, int y) { return (x >= 0) ? x / y : -((-x + y - 1) / y); };
    std::vector<uint2> t;
    for (int s = 0; s < p.nsteps; ++s) {
        const int d = p.ld[s];
        // Offset w only reaches the output sub-indices [w - c, w + c]: no single MMA need cover the whole accumulator, so the
        // step opens with one N = 128 MMA of zero operands that clears it (flag 0x200) and every real MMA accumulates.
        t.push_back(make_uint2(0u, 0x200u | (128u << 16)));
        for (int w = -c; w <= P - 1 + c; ++w) {
            int r_lo = std::max(0, w - c), r_hi = std::min(P - 1, w + c);
            if (C == 8) {   // N and the accumulator column must be multiples of 16: even r_lo, odd r_hi — a borrowed neighbour's
                if (r_lo & 1) --r_lo;            // tap is a zero plane of the weight image
                if (!(r_hi & 1)) ++r_hi;
            }
            ZVX_REQUIRE(r_lo >= 0 && r_hi <= P - 1 && r_lo - w >= -c - e && r_hi - w <= c + e, "voc_pair: column range out of plan");
            const int al = fdiv(w, P), rp = w - al * P;
            const uint32_t ncols = (uint32_t)((r_hi - r_lo + 1) * C);
            for (int pp = 0; pp < C / 8; ++pp) {
                const uint32_t ao = (uint32_t)((2 * pp) * p.Rtot + rp * p.S + p.G + al * d);
                const uint32_t bo = (uint32_t)((2 * pp) * p.ZC + (r_lo - w + c + e) * C);
                t.push_back(make_uint2(ao | (bo << 16), (uint32_t)(r_lo * C) | (1u << 8) | (ncols << 16)));
            }
        }
        ZVX_REQUIRE((int)t.size() == p.sched_off[s + 1], "voc_pair: schedule size mismatch");
    }
    for (const uint2& en : t) ZVX_REQUIRE((en.x & 0xFFFFu) < 16384u && (en.x >> 16) < 16384u, "voc_pair: operand offset out of range");
    if (t.size() & 1) t.push_back(make_uint2(0u, 0u));
    return t;
}

// Device copies of the schedules, one per (device, block shape); built on first use.
const uint2* schedule_for(const VocResArgs& a, const PairPlan& p) {
    static std::map<std::string, uint2*> cache;
    int dev = 0;
    ZVX_CUDA_CHECK(cudaGetDevice(&dev));
    std::string key = std::to_string(dev) + ":" + std::to_string(a.C) + ":" + std::to_string(a.k);
    for (int s = 0; s < a.nsteps; ++s) key += ":" + std::to_string(a.steps[s].dil);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const std::vector<uint2> t = build_schedule(a, p);
    uint2* d = nullptr;
    ZVX_CUDA_CHECK(cudaMalloc(&d, t.size() * sizeof(uint2)));
    ZVX_CUDA_CHECK(cudaMemcpy(d, t.data(), t.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    cache[key] = d;
    return d;
}

template <int C, int CTAS>
void launch(const VocResArgs& a, const PairPlan& p, cudaStream_t st) {
    static int attr_done = 0;
    if (!attr_done) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(voc_pair_kernel<C, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PER_SM));
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(voc_pair_kernel<C, CTAS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_done = 1;
    }
    dim3 grid(cdiv(a.T, p.TT), a.B);
    voc_pair_kernel<C, CTAS><<<grid, NT, p.smem_bytes, st>>>(a, p);
    ZVX_POST_LAUNCH();
}

}  // namespace

bool voc_pair_supported(int C, int k, const int* dils, int nd) {
    if (nd < 1 || 2 * nd > MAX_STEPS) return false;
    VocResArgs a;
    a.C = C; a.k = k; a.nsteps = 0;
    for (int i = 0; i < nd; ++i) {
        a.steps[a.nsteps].dil = dils[i]; a.steps[a.nsteps++].kind = 0;
        a.steps[a.nsteps].dil = 1; a.steps[a.nsteps++].kind = 1;
    }
    PairPlan p;
    return make_plan(a, &p);
}

// Steps must carry the images of voc_pair_pack_weight in `w_pair`.  Returns false (nothing launched) outside the plan.
bool voc_pair_tc(const VocResArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return true;
    PairPlan p;
    if (!make_plan(a, &p)) return false;
    for (int s = 0; s < a.nsteps; ++s)
        if (!a.steps[s].w_pair) return false;
    ZVX_REQUIRE(a.x && a.out && a.B <= 65535, "voc_pair_tc: bad arguments");
    VocResArgs b = a;
    b.sched = schedule_for(a, p);
    switch (a.C * 10 + p.ctas) {
        case 81: launch<8, 1>(b, p, st); break;
        case 82: launch<8, 2>(b, p, st); break;
        case 161: launch<16, 1>(b, p, st); break;
        case 162: launch<16, 2>(b, p, st); break;
        case 321: launch<32, 1>(b, p, st); break;
        default: launch<32, 2>(b, p, st); break;
    }
    return true;
}

// Weight image of one conv for voc_pair_kernel, TF32-rounded: the dilation-free trimmed Toeplitz array
// [cq][plane][co][4 ci] with plane z + c + e <-> tap j = c - z (zero planes at both ends when C = 8).
std::vector<float> voc_pair_pack_weight(const float* w, int C, int k) {
    auto rn = [](float v) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u = (u + 0x0FFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
        memcpy(&v, &u, 4);
        return v;
    };
    const int CQ = C / 4, c = (k - 1) / 2, e = (C == 8) ? 1 : 0, planes = planes_of(C, k);
    std::vector<float> o((size_t)CQ * planes * C * 4, 0.f);
    for (int z = -c; z <= c; ++z) {
        const int j = c - z;
        for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < C; ++ci)
                o[(((size_t)(ci / 4) * planes + (z + c + e)) * C + co) * 4 + (ci & 3)] = rn(w[((size_t)co * C + ci) * k + j]);
    }
    return o;
}

}  // namespace zvx
