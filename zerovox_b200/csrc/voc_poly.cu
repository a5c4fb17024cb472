// HiFi-GAN MRF residual block on the tensor cores, polyphase formulation (narrow stages, C = 8 / 16 / 32 channels).
//
// Why.  tcgen05.mma with both operands in shared memory costs ~100 cycles per instruction in the no-swizzle K-major layout
// whatever N is (up to 128; profiles/r01_ubench_tcgen05_tf32_mma.txt): a conv written as "M = 128 samples, N = C_out,
// one MMA per tap per 8 input channels" (voc_res.cu) is bound by the MMA COUNT.  Here one MMA produces P = 128 / C
// consecutive output samples per row: with t = P*n + r,
//     D[n, (r, co)] = sum_w sum_ci  x[P*n + w, ci] * B_w[(r, co), ci],      B_w[(r, co), ci] = W_j[co, ci],  (w - r) = (j - c)*d
// i.e. M = 128 values of n, N = P*C = 128 and K runs over the P + (k-1)*d input offsets w: (P + (k-1)d) * C/8 MMAs per
// 128*P samples instead of P*k*C/8 — 6.8x fewer for C = 8, k = 11.
//
// Layouts.  Activations live in shared memory "phase-major": sample tau of the tile is row  (tau % P)*S + G + tau / P  of
// every 16-byte channel-chunk plane [cq][row][4 floats], so the 128 rows n of the A operand for input offset w are the
// contiguous rows of phase (w mod P) shifted by floor(w / P) — a descriptor start address, as in voc_res.cu; G guard
// rows per phase block absorb the shifts.  B_w is a window of ONE zero-stuffed, tap-reversed weight array
//     E[z][co] = W_{c - z/d}[co]  if d | z and 0 <= c - z/d < k, else 0,     z = r - w,
// laid out [cq][z][co][4]: B_w starts (-w - z_min)*C entries into it — again only a start address.  Dilations d >= P (or
// arrays too large for shared memory) fall back, inside the same kernel and layout, to the per-phase direct form
// (N = C, one MMA per tap per 8 channels, accumulator columns r*C...).
// The rest is voc_res.cu's scheme: fp32 residual stream X next to the TF32 operand A, operands rewritten in place by the
// epilogue after each conv, weights streamed per conv with cp.async behind the previous epilogue, one launch per
// ResBlock including the MRF mean (hifigan.py:49-56, 80-84, 119-125).
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>

namespace zvx {

namespace {

constexpr int EPI_THREADS = 256;   // 8 loader / epilogue warps
constexpr int NT = EPI_THREADS + 32;   // + 1 warp: MMA issuer and weight streamer
constexpr int MAX_STEPS = VocResArgs::MAX_STEPS;
constexpr int kRounds = 2;   // operand hand-over rounds per conv step (4 measured slower twice: 6.7 vs 6.5 ms; each round costs a proxy fence)

struct PolyPlan {
    int nsteps, P, H, TT, R;
    int G, S, Rtot;
    int mode[MAX_STEPS];       // 0 Toeplitz, 1 direct
    int cd[MAX_STEPS];         // (k-1)/2 * dilation
    int n16[MAX_STEPS];        // 16-byte chunks of the step's weight image
    int ZC[MAX_STEPS];         // Toeplitz: entries per chunk plane (Z * C)
    int Np;                    // direct: rows of a dense weight plane
    int wbuf_bytes;            // one of the two weight buffers
    int sched_off[MAX_STEPS][5];   // MMA schedule entries of step s: round rd = [off[rd], off[rd + 1]), NR <= 4 rounds
    int sched_total;           // entries (padded to an even count)
    uint32_t offSched;
    uint32_t offA, offW, offBar;
    uint32_t idesc_t, idesc_d;
    int smem_bytes;
};

// leaky ReLU for slopes in (0, 1]: max(v, v*s)  (2 full-rate instructions)
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }
// fp32 -> tf32 round-to-nearest, ties away from zero (= cvt.rna.tf32.f32 for finite values) with two integer
// instructions; the cvt runs on the quarter-rate conversion pipe and would dominate the epilogue
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

template <int C>
__global__ void __launch_bounds__(NT) voc_poly_kernel(const VocResArgs a, const PolyPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sb = smem_u32(smem);
    const uint32_t sA = sb + p.offA;
    constexpr int CQ = C / 4, P = 128 / C, NPH = P / 2;
    constexpr int NR = kRounds < NPH ? kRounds : NPH, RPH = NPH / NR;   // operand hand-over in NR rounds of RPH phases per warp group
    // barriers: mma_done[2] (one per accumulator buffer) | w_full[2] (weight buffers) | round[NPH] (operand phases ready)
    const uint32_t bars = sb + p.offBar;
    auto mma_bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    auto w_bar = [&](int i) { return bars + 16u + 8u * (uint32_t)i; };
    auto round_bar = [&](int i) { return bars + 32u + 8u * (uint32_t)i; };
    const uint32_t slot = bars + 32u + 8u * NR;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.offBar + 32 + 8 * NR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * p.TT;
    const int S = p.S, G = p.G, Rtot = p.Rtot;

    const bool dbg = a.dbg && (tid == 0 || tid == EPI_THREADS) && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2;
    int dbg_i = (tid == 0) ? 0 : 32;
    auto stamp = [&]() { if (dbg) a.dbg[dbg_i++] = clock64(); };
    stamp();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(mma_bar(i), 1); mbar_init(w_bar(i), 1); }
        for (int i = 0; i < NR; ++i) mbar_init(round_bar(i), EPI_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == EPI_THREADS / 32) tmem_alloc(slot, 256);
    // MMA schedule (operand offsets of every tcgen05.mma, computed on the host once per block shape): global -> shared
    // Entries are expanded to {low word of the A descriptor, low word of the B descriptor, D column | flags} so that the
    // issuing thread spends no arithmetic per MMA (its scalar work would serialise with the tensor pipe).
    {
        const uint2* __restrict__ gs = reinterpret_cast<const uint2*>(a.sched);
        uint4* ss = reinterpret_cast<uint4*>(smem + p.offSched);
        for (int i = tid; i < p.sched_total; i += NT) {
            const uint2 e = __ldg(gs + i);
            int st = 0;
            while (st + 1 < p.nsteps && i >= p.sched_off[st + 1][0]) ++st;
            const uint32_t sWs = sb + p.offW + (uint32_t)((st & 1) * p.wbuf_bytes);
            const uint32_t lboB = (uint32_t)(p.mode[st] == 0 ? p.ZC[st] : p.Np);
            const uint32_t alo = (((sA & 0x3FFFFu) >> 4) + (e.x & 0xFFFFu)) | ((uint32_t)Rtot << 16);
            const uint32_t blo = (((sWs & 0x3FFFFu) >> 4) + (e.x >> 16)) | (lboB << 16);
            ss[i] = make_uint4(alo, blo, e.y, 0u);
        }
    }
    // guard rows of the operand tile: zero (they only feed halo outputs, but must stay finite)
    for (int idx = tid; idx < P * 2 * G * CQ; idx += NT) {
        const int cq = idx % CQ, gr = (idx / CQ) % (2 * G), r = idx / (CQ * 2 * G);
        const int row = r * S + (gr < G ? gr : 128 + gr);
        st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), make_float4(0.f, 0.f, 0.f, 0.f));
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    stamp();

    if (warp == EPI_THREADS / 32) {
        // ================================================================ MMA issuer + weight streamer (one thread)
        if (elect_one()) {
            auto load_w = [&](int s) {
                const uint32_t bytes = (uint32_t)p.n16[s] * 16u;
                mbar_arrive_expect_tx(w_bar(s & 1), bytes);
                bulk_load_1d(sb + p.offW + (uint32_t)((s & 1) * p.wbuf_bytes), a.steps[s].w_poly, bytes, w_bar(s & 1));
            };
            load_w(0);
            for (int s = 0; s < p.nsteps; ++s) {
                const uint32_t d_tmem = tmem_base + (uint32_t)((s & 1) * 128);
                const uint32_t rpar = (uint32_t)(s & 1);            // round barriers complete once per step (load = step -1)
                // every MMA of the step comes from the precomputed schedule: {A offset | B offset << 16, D column | flags};
                // a round's entries are issued as soon as the previous step's epilogue has handed that part of the operand over
                const uint32_t idesc = p.mode[s] == 0 ? p.idesc_t : p.idesc_d;
                const uint4* tab = reinterpret_cast<const uint4*>(smem + p.offSched);
                constexpr uint64_t DESC_HI = ((uint64_t)(128 >> 4) | ((uint64_t)1 << 14)) << 32;   // SBO = 128 B, version 1
                auto issue = [&](const uint4 e) {
                    umma_tf32(d_tmem + (e.z & 0xFFu), DESC_HI | e.x, DESC_HI | e.y, idesc, (e.z >> 8) & 1u);
                };
                mbar_wait_spin(w_bar(s & 1), (uint32_t)((s >> 1) & 1));
                for (int rd = 0; rd < NR; ++rd) {
                    mbar_wait_spin(round_bar(rd), rpar);
                    tc_fence_after();
                    if (rd == 0 && s + 1 < p.nsteps) load_w(s + 1);   // its buffer was last read by the MMAs of step s-1
                    const int i1 = p.sched_off[s][rd + 1];
                    int i = p.sched_off[s][rd];
                    for (; i + 4 <= i1; i += 4) {   // four entries in registers before the first MMA: loads overlap, MMAs go back to back
                        const uint4 e0 = tab[i], e1 = tab[i + 1], e2 = tab[i + 2], e3 = tab[i + 3];
                        issue(e0); issue(e1); issue(e2); issue(e3);
                    }
                    for (; i < i1; ++i) issue(tab[i]);
                }
                umma_commit(mma_bar(s & 1));
                stamp();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ loader + epilogue warps (0..7)
        // Thread (q, lane) owns accumulator row n and, with its warp group g, the 64 columns [64g, 64g + 64) = the NPH
        // consecutive samples tau0 .. tau0 + NPH - 1 (phases g*NPH + ri), all C channels: 16 float4 groups (ri, cq).  The
        // fp32 residual stream of those samples stays in registers for the whole block.
        const int q = warp & 3, g = warp >> 2;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int n = q * 32 + lane;
        const int tau0 = P * n + g * NPH;
        const int tg0 = t0 - p.H + tau0;
        const int row0 = (g * NPH) * S + G + n;
        float4 xo[16];
        {   // input tile: coalesced global loads (consecutive threads = consecutive 16-byte chunks of consecutive samples,
            // zero outside [0, T)) staged raw in the operand tile; then every thread takes the samples it owns into
            // registers (the fp32 residual stream) and rewrites them in place as the TF32 operand lrelu(x)
            const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
            const int tA = t0 - p.H;
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int idx = u * EPI_THREADS + tid;
                const int tau = idx / CQ, cq = idx - tau * CQ;
                const int t = tA + tau;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t >= 0 && t < a.T) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * C) + cq);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int idx = u * EPI_THREADS + tid;
                const int tau = idx / CQ, cq = idx - tau * CQ;
                st_shared_v4(sA + (uint32_t)((cq * Rtot + (tau % P) * S + G + tau / P) * 16), v[u]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // the 8 loader warps only
            const float4* As = reinterpret_cast<const float4*>(smem + p.offA);
#pragma unroll
            for (int idx = 0; idx < 16; ++idx) xo[idx] = As[(idx % CQ) * Rtot + row0 + (idx / CQ) * S];
#pragma unroll
            for (int idx = 0; idx < 16; ++idx) {
                const int ri = idx / CQ, cq = idx % CQ;
                const float4 x4 = xo[idx];
                float4 o;
                o.x = rna_tf32(lrelu(x4.x, a.in_slope)); o.y = rna_tf32(lrelu(x4.y, a.in_slope));
                o.z = rna_tf32(lrelu(x4.z, a.in_slope)); o.w = rna_tf32(lrelu(x4.w, a.in_slope));
                st_shared_v4(sA + (uint32_t)((cq * Rtot + row0 + ri * S) * 16), o);
                if (cq == CQ - 1 && (ri % RPH) == RPH - 1) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(round_bar(ri / RPH));
                }
            }
        }
        stamp();

        for (int s = 0; s < p.nsteps; ++s) {
            const int kind = a.steps[s].kind;
            const bool last = (s == p.nsteps - 1);
            float4 bb[CQ];
#pragma unroll
            for (int cq = 0; cq < CQ; ++cq) bb[cq] = __ldg(reinterpret_cast<const float4*>(a.steps[s].b) + cq);
            mbar_wait(mma_bar(s & 1), (uint32_t)((s >> 1) & 1));   // suspending wait: leaves the issue slots to the MMA issuer
            tc_fence_after();
            stamp();

            // accumulator read in 4 chunks of 16 columns, the next chunk in flight while the current one is processed.
            // The three step kinds are separate straight-line instantiations (MODE: 0 first conv of a pair, 1 residual
            // step, 2 last step); samples outside the utterance are handled with selects, not branches.
            const uint32_t tcol = trow + (uint32_t)((s & 1) * 128 + g * 64);
            auto epilogue = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
                uint32_t vbuf[2][16];
                float4 sbuf[2][4];   // MODE 2: MRF accumulator values of the chunk, fetched one chunk ahead
                auto fetch_acc = [&](int chunk) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int id = chunk * 4 + j, ri = id / CQ, cq = id % CQ;
                        const int t = tg0 + ri, tau = tau0 + ri;
                        sbuf[chunk & 1][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.acc_in && t >= 0 && t < a.T && tau >= p.H && tau < p.R - p.H)
                            sbuf[chunk & 1][j] = *(reinterpret_cast<const float4*>(a.acc_in + (long long)b * a.acc_in_bs + (long long)t * C) + cq);
                    }
                };
                if constexpr (MODE == 2) fetch_acc(0);
                __syncwarp();
                tmem_ld16(tcol, vbuf[0]);
#pragma unroll
                for (int idx = 0; idx < 16; ++idx) {
                    if ((idx & 3) == 0) {
                        if constexpr (MODE == 2) { if (idx + 4 < 16) fetch_acc((idx >> 2) + 1); }
                        __syncwarp();
                        tmem_wait_ld();
                        if (idx + 4 < 16) tmem_ld16(tcol + (uint32_t)(4 * (idx + 4)), vbuf[((idx >> 2) + 1) & 1]);
                        else tc_fence_before();   // orders the TMEM reads before the MMAs that overwrite this buffer two steps later
                    }
                    const uint32_t* v = vbuf[(idx >> 2) & 1] + 4 * (idx & 3);
                    const int ri = idx / CQ, cq = idx % CQ;
                    const int t = tg0 + ri, tau = tau0 + ri;
                    const bool inside = (t >= 0) && (t < a.T);
                    const int row = row0 + ri * S;
                    float4 c4;
                    c4.x = __uint_as_float(v[0]) + bb[cq].x; c4.y = __uint_as_float(v[1]) + bb[cq].y;
                    c4.z = __uint_as_float(v[2]) + bb[cq].z; c4.w = __uint_as_float(v[3]) + bb[cq].w;
                    if constexpr (MODE == 0) {
                        // first conv of a pair: bias, lrelu, TF32 -> next operand (zero outside the utterance: conv2's padding)
                        float4 o;
                        o.x = inside ? rna_tf32(lrelu(c4.x, a.mid_slope)) : 0.f; o.y = inside ? rna_tf32(lrelu(c4.y, a.mid_slope)) : 0.f;
                        o.z = inside ? rna_tf32(lrelu(c4.z, a.mid_slope)) : 0.f; o.w = inside ? rna_tf32(lrelu(c4.w, a.mid_slope)) : 0.f;
                        st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), o);
                    } else if constexpr (MODE == 1) {
                        // residual step: x += conv + bias (fp32, registers), next operand = lrelu(x) in TF32
                        float4 x4 = xo[idx];
                        x4.x = inside ? x4.x + c4.x : x4.x; x4.y = inside ? x4.y + c4.y : x4.y;
                        x4.z = inside ? x4.z + c4.z : x4.z; x4.w = inside ? x4.w + c4.w : x4.w;
                        xo[idx] = x4;   // stays 0 outside the utterance
                        float4 o;
                        o.x = rna_tf32(lrelu(x4.x, a.in_slope)); o.y = rna_tf32(lrelu(x4.y, a.in_slope));
                        o.z = rna_tf32(lrelu(x4.z, a.in_slope)); o.w = rna_tf32(lrelu(x4.w, a.in_slope));
                        st_shared_v4(sA + (uint32_t)((cq * Rtot + row) * 16), o);
                    } else {
                        // last step: x += conv + bias, then the MRF bookkeeping and the store (channel-last rows)
                        if (inside && tau >= p.H && tau < p.R - p.H) {
                            const float4 sa = sbuf[(idx >> 2) & 1][idx & 3];
                            float4 y;
                            y.x = xo[idx].x + c4.x; y.y = xo[idx].y + c4.y; y.z = xo[idx].z + c4.z; y.w = xo[idx].w + c4.w;
                            y.x = fmaf(y.x, a.out_scale, sa.x); y.y = fmaf(y.y, a.out_scale, sa.y);
                            y.z = fmaf(y.z, a.out_scale, sa.z); y.w = fmaf(y.w, a.out_scale, sa.w);
                            float* op = a.out + (long long)b * a.out_bs + (long long)t * C;
                            reinterpret_cast<float4*>(op)[cq] = make_float4(lrelu(y.x, a.out_slope), lrelu(y.y, a.out_slope),
                                                                           lrelu(y.z, a.out_slope), lrelu(y.w, a.out_slope));
                        }
                    }
                    if (MODE != 2 && cq == CQ - 1 && (ri % RPH) == RPH - 1) {   // RPH more phases of the next operand are complete
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(round_bar(ri / RPH));
                    }
                }
            };
            if (kind == 0) epilogue(std::integral_constant<int, 0>{});
            else if (!last) epilogue(std::integral_constant<int, 1>{});
            else epilogue(std::integral_constant<int, 2>{});
            stamp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EPI_THREADS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

int up_to_mod8(int v, int want) {
    while ((v & 7) != want) ++v;
    return v;
}

constexpr int kToeplitzMaxBytes = 72 * 1024;

bool toeplitz_ok(int C, int k, int d) {
    const int P = 128 / C, cd = (k - 1) / 2 * d;
    const int Z = 2 * (P - 1) + 2 * cd + 1;
    return d < P && (long long)Z * C * 16 * (C / 4) <= kToeplitzMaxBytes;
}

bool make_plan(const VocResArgs& a, PolyPlan* out) {
    const int C = a.C, CQ = C / 4, k = a.k, ns = a.nsteps, P = 128 / C;
    if (ns < 1 || ns > MAX_STEPS || (k & 1) == 0) return false;
    PolyPlan p{};
    p.nsteps = ns; p.P = P; p.R = 128 * P;
    p.Np = std::max(C, 16);
    int H = 0, cdmax = 0, wbytes = 0;
    for (int s = 0; s < ns; ++s) {
        const int d = a.steps[s].dil, cd = (k - 1) / 2 * d;
        p.cd[s] = cd;
        H += cd;
        cdmax = std::max(cdmax, cd);
        if (toeplitz_ok(C, k, d)) {
            const int Z = 2 * (P - 1) + 2 * cd + 1;
            p.mode[s] = 0; p.ZC[s] = Z * C; p.n16[s] = CQ * Z * C;
        } else {
            if (C < 16) return false;   // the direct form needs N = C >= 16
            p.mode[s] = 1; p.ZC[s] = 0; p.n16[s] = k * CQ * p.Np;
        }
        wbytes = std::max(wbytes, p.n16[s] * 16);
    }
    p.H = H;
    p.TT = p.R - 2 * H;
    if (p.TT < p.R / 2) return false;   // halo would dominate: leave it to voc_res.cu
    p.G = (cdmax + P - 1) / P + 1;
    p.S = up_to_mod8(128 + 2 * p.G, 1);
    p.Rtot = up_to_mod8(P * p.S, 8 / CQ == 8 ? 0 : 8 / CQ);
    if (CQ == 1) return false;
    uint32_t o = 0;
    p.offA = o; o += (uint32_t)(CQ * p.Rtot * 16);
    p.wbuf_bytes = (int)round_up(wbytes, 128);
    p.offW = o; o += (uint32_t)(2 * p.wbuf_bytes);
    p.offBar = o; o += 32 + 8 * 4 + 16;
    o = (uint32_t)round_up(o, 16);
    // MMA schedule: round rd of step s hands over the input phases {rd*RPH .. rd*RPH + RPH - 1} + {0, NPH}
    {
        const int NPH = P / 2, NR = std::min(kRounds, NPH), RPH = NPH / NR;
        int n = 0;
        for (int s = 0; s < ns; ++s) {
            const int cd = p.cd[s];
            p.sched_off[s][0] = n;
            if (p.mode[s] == 0) {
                for (int rd = 0; rd < NR; ++rd) {
                    const int wmin = -cd, wmax = P - 1 + cd;
                    for (int hq = 0; hq < 2 * RPH; ++hq) {
                        const int rp = (hq / RPH) * NPH + rd * RPH + (hq % RPH);
                        for (int w = wmin + (((rp - wmin) % P) + P) % P; w <= wmax; w += P) n += C / 8;
                    }
                    p.sched_off[s][rd + 1] = n;
                }
            } else {
                for (int rd = 1; rd < NR; ++rd) p.sched_off[s][rd] = n;   // direct form: every output phase reads several
                n += P * k * (C / 8);                                    // input phases -> everything in the last round
                p.sched_off[s][NR] = n;
            }
        }
        p.sched_total = (n + 1) & ~1;
    }
    p.offSched = o; o += (uint32_t)p.sched_total * 16;
    p.smem_bytes = (int)o + 128;
    p.idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.idesc_d = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (p.smem_bytes > 226 * 1024) return false;
    *out = p;
    return true;
}

// The tcgen05.mma list of one block shape in issue order (see the kernel): x = A offset | B offset << 16 (16-byte units
// relative to the operand bases), y = accumulator column | accumulate flag << 8.
std::vector<uint2> build_schedule(const VocResArgs& a, const PolyPlan& p) {
    const int C = a.C, CQ = C / 4, P = p.P, NPH = P / 2, NR = std::min(kRounds, NPH), RPH = NPH / NR, k = a.k;
    auto fdiv = [](int x, int y) { return (x >= 0) ? x / y : -((-x + y - 1) / y); };
    std::vector<uint2> t;
    for (int s = 0; s < p.nsteps; ++s) {
        const int cd = p.cd[s], dil = a.steps[s].dil;
        if (p.mode[s] == 0) {
            const int wmin = -cd, wmax = P - 1 + cd, zmin = -(P - 1) - cd;
            bool first = true;
            for (int rd = 0; rd < NR; ++rd) {
                // offsets whose input phase this round hands over, in ascending order (consecutive offsets read
                // overlapping weight windows: measurably faster than grouping the MMAs by phase)
                std::vector<int> ws;
                for (int hq = 0; hq < 2 * RPH; ++hq) {
                    const int rp = (hq / RPH) * NPH + rd * RPH + (hq % RPH);
                    for (int w = wmin + (((rp - wmin) % P) + P) % P; w <= wmax; w += P) ws.push_back(w);
                }
                std::sort(ws.begin(), ws.end());
                for (int w : ws) {
                    const int al = fdiv(w, P), rp = w - al * P;
                    for (int pp = 0; pp < C / 8; ++pp) {
                        const uint32_t ao = (uint32_t)((2 * pp) * p.Rtot + rp * p.S + p.G + al);
                        const uint32_t bo = (uint32_t)((2 * pp) * p.ZC[s] + (-w - zmin) * C);
                        t.push_back(make_uint2(ao | (bo << 16), first ? 0u : (1u << 8)));
                        first = false;
                    }
                }
            }
        } else {
            const int c = (k - 1) / 2;
            for (int r = 0; r < P; ++r)
                for (int j = 0; j < k; ++j) {
                    const int w = r + (j - c) * dil;
                    const int al = fdiv(w, P), rp = w - al * P;
                    for (int pp = 0; pp < C / 8; ++pp) {
                        const uint32_t ao = (uint32_t)((2 * pp) * p.Rtot + rp * p.S + p.G + al);
                        const uint32_t bo = (uint32_t)((j * CQ + 2 * pp) * p.Np);
                        t.push_back(make_uint2(ao | (bo << 16), (uint32_t)(r * C) | ((j | pp) ? (1u << 8) : 0u)));
                    }
                }
        }
        ZVX_REQUIRE((int)t.size() == p.sched_off[s][NR], "voc_poly: schedule size mismatch");
    }
    for (const uint2& e : t) ZVX_REQUIRE((e.x & 0xFFFFu) < 16384u && (e.x >> 16) < 16384u, "voc_poly: operand offset out of range");
    if (t.size() & 1) t.push_back(make_uint2(0u, 0u));
    return t;
}

// Device copies of the schedules, one per (device, block shape); built on first use.
const uint2* schedule_for(const VocResArgs& a, const PolyPlan& p) {
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

template <int C>
void launch(const VocResArgs& a, const PolyPlan& p, cudaStream_t st) {
    static int attr_done = 0;
    if (!attr_done) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(voc_poly_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = 1;
    }
    dim3 grid(cdiv(a.T, p.TT), a.B);
    voc_poly_kernel<C><<<grid, NT, p.smem_bytes, st>>>(a, p);
    ZVX_POST_LAUNCH();
}

}  // namespace

bool voc_poly_supported(int C, int k, const int* dils, int nd, bool pair) {
    if (!(C == 8 || C == 16 || C == 32) || nd < 1 || nd * (pair ? 2 : 1) > MAX_STEPS) return false;
    VocResArgs a;
    a.C = C; a.k = k; a.nsteps = 0;
    for (int i = 0; i < nd; ++i) {
        if (dils[i] < 1) return false;
        a.steps[a.nsteps++].dil = dils[i];
        if (pair) a.steps[a.nsteps++].dil = 1;
    }
    PolyPlan p;
    return make_plan(a, &p);
}

bool voc_poly_tc(const VocResArgs& a, cudaStream_t st) {
    if (a.B == 0 || a.T == 0) return true;
    PolyPlan p;
    if (!(a.C == 8 || a.C == 16 || a.C == 32) || !make_plan(a, &p)) return false;
    for (int s = 0; s < a.nsteps; ++s)
        if (!a.steps[s].w_poly) return false;
    ZVX_REQUIRE(a.x && a.out && a.steps[a.nsteps - 1].kind == 1 && a.B <= 65535, "voc_poly_tc: bad arguments");
    VocResArgs b = a;
    b.sched = schedule_for(a, p);
    switch (a.C) {
        case 8: launch<8>(b, p, st); break;
        case 16: launch<16>(b, p, st); break;
        default: launch<32>(b, p, st); break;
    }
    return true;
}

// Weight image of one conv for the polyphase kernel, TF32-rounded: the zero-stuffed tap-reversed Toeplitz array
// [cq][z][co][4] when the (C, k, dilation) combination uses the Toeplitz form, else the dense [tap][cq][Np][4] image.
std::vector<float> voc_poly_pack_weight(const float* w, int C, int k, int dil) {
    auto rn = [](float v) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u = (u + 0x0FFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
        memcpy(&v, &u, 4);
        return v;
    };
    const int CQ = C / 4, P = 128 / C, c = (k - 1) / 2, cd = c * dil;
    if (toeplitz_ok(C, k, dil)) {
        const int Z = 2 * (P - 1) + 2 * cd + 1, zmin = -(P - 1) - cd;
        std::vector<float> o((size_t)CQ * Z * C * 4, 0.f);
        for (int z = zmin; z < zmin + Z; ++z) {
            if (z % dil != 0) continue;
            const int j = c - z / dil;
            if (j < 0 || j >= k) continue;
            for (int co = 0; co < C; ++co)
                for (int ci = 0; ci < C; ++ci)
                    o[(((size_t)(ci / 4) * Z + (z - zmin)) * C + co) * 4 + (ci & 3)] = rn(w[((size_t)co * C + ci) * k + j]);
        }
        return o;
    }
    return voc_pack_weight(w, C, C, k);
}

}  // namespace zvx
