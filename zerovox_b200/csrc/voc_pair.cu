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
#include <memory>
#include <string>
#include <type_traits>

namespace zvx {

namespace {

constexpr int EPI_THREADS = 256;         // 8 loader / epilogue warps; lane 0 of warp 0 also issues the MMAs and streams the weights
constexpr int NT = EPI_THREADS;          // (no dedicated issuer warp: 2 x 8 warps per SM leave 128 registers per thread)
constexpr int MAX_STEPS = VocResArgs::MAX_STEPS;
constexpr int SMEM_PER_SM = 227 * 1024;
constexpr int MAX_SCHED = 480;           // MMAs of one launch (8 steps x 57 at C = 32, k = 11) + 8 entries of read-ahead padding

struct PairPlan {
    int nsteps, P, lgP, S, G, Rtot, R, TT, lo;
    int ld[MAX_STEPS];          // dilation of the layout the step's operand is stored in (= the conv's dilation)
    uint32_t mg[MAX_STEPS];     // ceil(2^20 / ld): tau / ld == (tau * mg) >> 20 for every tau of a tile (checked on the host)
    int n16[MAX_STEPS];         // 16-byte chunks of the step's weight image
    int ZC;                     // rows per 16-byte channel-chunk plane of a weight image (tap planes * C)
    int wbuf_bytes, nwbuf;
    int sched_off[MAX_STEPS + 1];
    int sched_total;
    uint32_t offA, offW, offOnes, offBar;   // offOnes: 128 rows (1, 1, 0, 0) then 128 zero rows (accumulator-initialising MMA)
    int smem_bytes, ctas;
    int wait_ns;                // suspend-time hint of the epilogue warps' barrier waits
    // The tcgen05.mma list in issue order, per step, ready to use: x = A descriptor low word minus the tile's base (row offset
    // | LBO << 16), y = the same for B (relative to the weight buffer), z = instruction descriptor, w = accumulator column.  It
    // travels as a kernel PARAMETER: the issuing thread reads an entry from the constant bank straight into uniform registers
    // (one LDCU.128) and needs three uniform adds per MMA — read from shared memory every MMA cost four R2UR moves plus the
    // unpacking, and the single issuing thread, not the tensor pipe (64 cycles per N = 128 MMA), set the pace (~110-150 cycles).
    uint4 sched[MAX_SCHED];
};

// leaky ReLU for slopes in (0, 1]: max(v, v*s)
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }
// fp32 -> tf32 round-to-nearest, ties away from zero (= cvt.rna.tf32.f32 for finite values) with two integer instructions
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// Operand values are read by the tensor core as TF32 = the fp32 bits with the low 13 mantissa bits IGNORED: adding half a
// TF32 ulp to the bit pattern makes that truncation a round-to-nearest (ties away) — one integer add, no mask.
__device__ __forceinline__ float half_ulp_up(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float4 act4(float4 v, float s) {
    return make_float4(half_ulp_up(lrelu(v.x, s)), half_ulp_up(lrelu(v.y, s)), half_ulp_up(lrelu(v.z, s)), half_ulp_up(lrelu(v.w, s)));
}
// Lean barrier wait for the epilogue warps: try_wait with a suspend-time hint (the warp sleeps in hardware instead of spinning
// through the issue slots of the co-resident CTA); a protocol bug traps after ~2^22 wake-ups instead of hanging the GPU.
__device__ __forceinline__ void wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    int spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(ns)
            : "memory");
        if (done) return;
        if (++spins > (1 << 22)) __trap();
    }
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
__global__ void __launch_bounds__(NT, CTAS) voc_pair_kernel(const __grid_constant__ VocResArgs a, const __grid_constant__ PairPlan p) {
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
    int dbg_j = 32;
#define ZVX_STAMP() do { if (dbg) a.dbg[dbg_i++] = clock64(); } while (0)
#define ZVX_STAMP2() do { if (dbg && dbg_j < 64) a.dbg[dbg_j++] = clock64(); } while (0)   // issuer: waits over | MMAs issued
#else
#define ZVX_STAMP() do { } while (0)
#define ZVX_STAMP2() do { } while (0)
#endif
    ZVX_STAMP();
    pdl_trigger();

    if (tid == 0) {
        mbar_init(mma_bar, 1);
        mbar_init(w_bar(0), 1);
        mbar_init(w_bar(1), 1);
        mbar_init(op_bar, EPI_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 128);
    // the whole operand tile starts as zeros: guard rows and layout rows no step writes only feed halo outputs, but must
    // stay finite
    for (int idx = tid; idx < CQ * Rtot; idx += NT) st_shared_v4(sA + (uint32_t)idx * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int idx = tid; idx < 256; idx += NT)
        st_shared_v4(sb + p.offOnes + (uint32_t)idx * 16u, idx < 128 ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    pdl_wait();   // (tile cleared, barriers and TMEM set up while the previous launch drains)
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
        ZVX_STAMP2();
        // double-buffered weights: the other buffer was last read by the MMAs of step s-1, complete since the epilogue of
        // s-1 has run
        if (p.nwbuf == 2 && s + 1 < p.nsteps) load_w(s + 1);
        constexpr uint32_t DESC_HI = (uint32_t)(128 >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
        constexpr uint32_t IDESC0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);   // f32 += tf32 x tf32, M = 128
        const uint32_t a_rows = (sA & 0x3FFFFu) >> 4;
        const uint32_t w_rows = ((sb + p.offW + (uint32_t)(buf * p.wbuf_bytes)) & 0x3FFFFu) >> 4;
        {   // accumulator := bias.  A = 128 rows (1, 1, 0, 0) with the zero rows as second K-half; B = the bias block at the end
            // of the step's weight image, rows (hi(b[co]), lo(b[co]), 0, 0), second K-half = the same zero rows
            const uint32_t ones = ((sb + p.offOnes) & 0x3FFFFu) >> 4, bias_rows = w_rows + (uint32_t)(CQ * p.ZC);
            umma_tf32_lo(tmem_base, ones | (128u << 16), bias_rows | ((ones + 128u - bias_rows) << 16), DESC_HI, IDESC0 | (128u << 14), 0u);
        }
        const int i1 = p.sched_off[s + 1];
        int i = p.sched_off[s];
        auto issue = [&](const uint4 e) {   // (row offsets stay below 2^14: no carry into the LBO field)
            umma_tf32_lo(tmem_base + e.w, a_rows + e.x, w_rows + e.y, DESC_HI, e.z, 1u);
        };
        // entries are fetched one group of four ahead of the MMAs that use them (the asm statements are memory barriers to the
        // compiler: without the explicit prefetch every MMA would wait for its own constant-bank load)
        uint4 n0 = p.sched[i], n1 = p.sched[i + 1], n2 = p.sched[i + 2], n3 = p.sched[i + 3];
        for (; i + 4 <= i1; i += 4) {
            const uint4 e0 = n0, e1 = n1, e2 = n2, e3 = n3;
            n0 = p.sched[i + 4]; n1 = p.sched[i + 5]; n2 = p.sched[i + 6]; n3 = p.sched[i + 7];   // (the table is padded by 8)
            issue(e0); issue(e1); issue(e2); issue(e3);
        }
        if (i < i1) issue(n0);
        if (i + 1 < i1) issue(n1);
        if (i + 2 < i1) issue(n2);
        ZVX_STAMP2();
        umma_commit(mma_bar);
    };
    if (warp == 0 && elect_one()) load_w(0);
    {
        // ================================================================ loader + epilogue warps (0..7)
        // Thread (q, lane, g) reads accumulator row n = 32q + lane, columns [64g, 64g + 64) = the NPH sub-indices
        // g*NPH + ri of that row, all C channels: 16 float4 (ri, cq).  In the d = 1 layout row n holds the samples
        // P*n .. P*n + P - 1: the fp32 residual stream of the thread's NPH consecutive samples stays in registers.
        const int q = warp & 3, g = warp >> 2;
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64);
        const int n = q * 32 + lane;
        const int tau0 = P * n + g * NPH;            // first owned sample (d = 1 ownership)
        const bool interior = tbase >= 0 && tbase + 130 * P <= a.T;   // every sample any layout of this tile holds is inside the utterance
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
                if (elect_one()) issue_step(s);
                __syncwarp();
            }
            if (s + 2 >= p.nsteps) {
                // While the MMAs of the last two steps run (the residual stream is final after step nsteps-3's epilogue):
                // x <- x * out_scale + acc_in (the MRF running sum, hifigan.py:119-125), one batch of 8 float4 per MMA phase —
                // the global-load latency hides behind the tensor pipe instead of stalling the final epilogue.
                auto fold = [&](auto h_tag) {
                    constexpr int h = decltype(h_tag)::value;
                    float4 sa[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int idx = h * 8 + j, ri = idx / CQ, cq = idx % CQ;
                        const int tau = tau0 + ri, t = tbase + tau;
                        sa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.acc_in && t >= 0 && t < a.T && tau >= p.lo && tau < p.lo + p.TT)
                            sa[j] = *(reinterpret_cast<const float4*>(a.acc_in + (long long)b * a.acc_in_bs + (long long)t * C) + cq);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4& x4 = xo[h * 8 + j];
                        x4.x = fmaf(x4.x, a.out_scale, sa[j].x); x4.y = fmaf(x4.y, a.out_scale, sa[j].y);
                        x4.z = fmaf(x4.z, a.out_scale, sa[j].z); x4.w = fmaf(x4.w, a.out_scale, sa[j].w);
                    }
                };
                if (last) fold(std::integral_constant<int, 1>{}); else fold(std::integral_constant<int, 0>{});
            }
            ZVX_STAMP();
            wait_sleep(mma_bar, (uint32_t)(s & 1), (uint32_t)p.wait_ns);   // sleeping wait: leaves the issue slots to the co-resident CTA
            tc_fence_after();
            ZVX_STAMP();
            if (warp == 0 && p.nwbuf == 1 && s + 1 < p.nsteps && elect_one()) load_w(s + 1);   // single weight buffer: this step's MMAs have read it

            // The accumulator already holds conv + bias (the step's first MMA initialises it with the bias).  Variants, all
            // straight-line: MODE 0 first conv of a pair / 1 residual step / 2 last step; SCAT: the thread's samples (MODE 0) or
            // the produced operand (MODE 1) are in a dilated layout -> index arithmetic + predicated stores; INTERIOR: the whole
            // tile lies inside the utterance -> no zero-padding selects.
            auto epilogue = [&](auto mode_tag, auto scat_tag, auto int_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
                constexpr bool SCAT = decltype(scat_tag)::value, INTERIOR = decltype(int_tag)::value;
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
                    const bool inside = INTERIOR || ((t >= 0) && (t < a.T));
                    const float4 c4 = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                    const int own_row = (g * NPH + ri) * S + G + n;   // the thread's own position in the d = 1 layout
                    if constexpr (MODE == 0) {
                        // first conv of a pair: lrelu, TF32 -> operand of the d = 1 conv (zero outside the utterance: its padding)
                        float4 o = act4(c4, a.mid_slope);
                        if (!inside) o = make_float4(0.f, 0.f, 0.f, 0.f);
                        if constexpr (SCAT) {
                            const bool ok = (tau >> lgP) < 128;
                            const int row = (tau & (P - 1)) * S + G + (tau >> lgP);
                            if (ok) st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), o);
                        } else {
                            st_shared_v4(sA + (uint32_t)((cq * Rtot + own_row) * 16), o);
                        }
                    } else if constexpr (MODE == 1) {
                        // residual step: x += conv (fp32, registers), next operand = lrelu(x) in TF32, stored in the layout of the
                        // next step's dilation
                        float4 x4 = xo[idx];
                        if (inside) { x4.x += c4.x; x4.y += c4.y; x4.z += c4.z; x4.w += c4.w; }
                        xo[idx] = x4;   // stays 0 outside the utterance
                        const float4 o = act4(x4, a.in_slope);
                        if constexpr (SCAT) {
                            const int row = layout_row(tau, dn, mgn, P, lgP, S, G);
                            if (row >= 0) st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), o);
                        } else {
                            st_shared_v4(sA + (uint32_t)((cq * Rtot + own_row) * 16), o);
                        }
                    } else {
                        // last step: out = lrelu(x * out_scale + acc_in + conv * out_scale) — xo already holds the first two terms
                        if (inside && tau >= p.lo && tau < p.lo + p.TT) {
                            float4 y;
                            y.x = fmaf(c4.x, a.out_scale, xo[idx].x); y.y = fmaf(c4.y, a.out_scale, xo[idx].y);
                            y.z = fmaf(c4.z, a.out_scale, xo[idx].z); y.w = fmaf(c4.w, a.out_scale, xo[idx].w);
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
            using T_ = std::true_type;
            using F_ = std::false_type;
            auto run = [&](auto mode_tag, bool scat) {
                if (scat) { if (interior) epilogue(mode_tag, T_{}, T_{}); else epilogue(mode_tag, T_{}, F_{}); }
                else { if (interior) epilogue(mode_tag, F_{}, T_{}); else epilogue(mode_tag, F_{}, F_{}); }
            };
            if (kind == 0) run(std::integral_constant<int, 0>{}, ds != 1);
            else if (!last) run(std::integral_constant<int, 1>{}, dn != 1);
            else run(std::integral_constant<int, 2>{}, false);
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
    const int wbytes = CQ * p.ZC * 16 + 128 * 16;           // tap planes + the bias block (128 rows)
    for (int s = 0; s < ns; ++s) p.n16[s] = CQ * p.ZC + 128;
    p.wait_ns = env_int("ZVX_PAIR_WAIT_NS", 2000);
    p.wbuf_bytes = (int)round_up(wbytes, 128);
    // MMA schedule size: per step the P + k - 1 offsets x C/8 k-chunks (after the bias-initialising MMA the kernel issues itself)
    const int per_step = (P + k - 1) * (C / 8);
    for (int s = 0; s <= ns; ++s) p.sched_off[s] = s * per_step;
    p.sched_total = ns * per_step;
    if (p.sched_total + 8 > MAX_SCHED) return false;
    auto layout = [&](int nwbuf) {
        uint32_t o = 0;
        p.nwbuf = nwbuf;
        p.offA = o; o += (uint32_t)(CQ * p.Rtot * 16);
        p.offW = o; o += (uint32_t)(nwbuf * p.wbuf_bytes);
        p.offOnes = o; o += 256 * 16;
        p.offBar = o; o += 48;
        p.smem_bytes = (int)round_up(o, 16) + 128;
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
    // The tcgen05.mma list.  Offset w of the sub-index only reaches the output sub-indices [w - c, w + c]: edge offsets issue
    // N = (#r) * C columns at accumulator column r_lo * C; all of them accumulate onto the bias the step's first MMA wrote.
    {
        const int e = (C == 8) ? 1 : 0;
        auto fdiv = [](int x, int y) { return (x >= 0) ? x / y : -((-x + y - 1) / y); };
        int i = 0;
        for (int s = 0; s < ns; ++s) {
            const int d = p.ld[s];
            for (int w = -c; w <= P - 1 + c; ++w) {
                int r_lo = std::max(0, w - c), r_hi = std::min(P - 1, w + c);
                if (C == 8) {   // N and the accumulator column must be multiples of 16: even r_lo, odd r_hi — a borrowed
                    if (r_lo & 1) --r_lo;        // neighbour's tap is a zero plane of the weight image
                    if (!(r_hi & 1)) ++r_hi;
                }
                if (!(r_lo >= 0 && r_hi <= P - 1 && r_lo - w >= -c - e && r_hi - w <= c + e)) return false;
                const int al = fdiv(w, P), rp = w - al * P;
                const uint32_t ncols = (uint32_t)((r_hi - r_lo + 1) * C);
                for (int pp = 0; pp < C / 8; ++pp) {
                    const uint32_t ao = (uint32_t)((2 * pp) * p.Rtot + rp * p.S + p.G + al * d);
                    const uint32_t bo = (uint32_t)((2 * pp) * p.ZC + (r_lo - w + c + e) * C);
                    if (ao + 4096u >= 16384u || bo + 8192u >= 16384u) return false;   // + the largest base (shared memory < 228 KB)... 
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                    p.sched[i++] = make_uint4(ao | ((uint32_t)p.Rtot << 16), bo | ((uint32_t)p.ZC << 16), idesc, (uint32_t)(r_lo * C));
                }
            }
            if (i != p.sched_off[s + 1]) return false;
        }
    }
    *out = p;
    return true;
}

// Plans (tile geometry + MMA list) are pure functions of the block shape: built once per (C, k, dilations).
const PairPlan* plan_for(const VocResArgs& a) {
    static std::map<std::string, std::unique_ptr<PairPlan>> cache;
    std::string key = std::to_string(a.C) + ":" + std::to_string(a.k) + ":" + std::to_string(a.nsteps);
    for (int s = 0; s < a.nsteps && s < MAX_STEPS; ++s) key += ":" + std::to_string(a.steps[s].dil) + "/" + std::to_string(a.steps[s].kind);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::unique_ptr<PairPlan> p(new PairPlan());
        if (!make_plan(a, p.get())) p.reset();
        it = cache.emplace(key, std::move(p)).first;
    }
    return it->second.get();
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
    launch_k(voc_pair_kernel<C, CTAS>, grid, dim3(NT), (size_t)p.smem_bytes, st, a, p);
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
    return plan_for(a) != nullptr;
}

// Steps must carry the images of voc_pair_pack_weight in `w_pair`.  Returns false (nothing launched) outside the plan.
bool voc_pair_tc(const VocResArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return true;
    const PairPlan* pp = plan_for(a);
    if (!pp) return false;
    const PairPlan& p = *pp;
    for (int s = 0; s < a.nsteps; ++s)
        if (!a.steps[s].w_pair) return false;
    ZVX_REQUIRE(a.x && a.out && a.B <= 65535, "voc_pair_tc: bad arguments");
    const VocResArgs& b = a;
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
// [cq][plane][co][4 ci] with plane z + c + e <-> tap j = c - z (zero planes at both ends when C = 8), followed by the
// bias block (128 rows of 4 floats).
std::vector<float> voc_pair_pack_weight(const float* w, const float* bias, int C, int k) {
    auto rn = [](float v) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u = (u + 0x0FFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
        memcpy(&v, &u, 4);
        return v;
    };
    const int CQ = C / 4, c = (k - 1) / 2, e = (C == 8) ? 1 : 0, planes = planes_of(C, k);
    std::vector<float> o((size_t)CQ * planes * C * 4 + 128 * 4, 0.f);
    // bias block: row (r, co) = (hi, lo, 0, 0) with hi = the bias truncated to TF32 (what the tensor core reads), lo = the
    // TF32-rounded remainder: ones(1, 1, 0, 0) x this row = the bias to ~2^-21 relative
    for (int row = 0; row < 128; ++row) {
        const float bv = bias[row % C];
        uint32_t u;
        memcpy(&u, &bv, 4);
        u &= ~0x1FFFu;
        float hi;
        memcpy(&hi, &u, 4);
        o[(size_t)CQ * planes * C * 4 + (size_t)row * 4 + 0] = hi;
        o[(size_t)CQ * planes * C * 4 + (size_t)row * 4 + 1] = rn(bv - hi);
    }
    for (int z = -c; z <= c; ++z) {
        const int j = c - z;
        for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < C; ++ci)
                o[(((size_t)(ci / 4) * planes + (z + c + e)) * C + co) * 4 + (ci & 3)] = rn(w[((size_t)co * C + ci) * k + j]);
    }
    return o;
}

}  // namespace zvx
