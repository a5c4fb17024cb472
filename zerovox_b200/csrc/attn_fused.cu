// Fused scaled-dot-product attention of the decoder's FFT blocks (fs2.py:101-163: bmm -> / temperature -> masked_fill(-inf)
// -> softmax -> bmm) on tcgen05, without the score matrix ever leaving the SM.
//
//   out[b, q, h*dk + c] = sum_j softmax_j( <Q[b,q,h,:], K[b,j,h,:]> / temperature  |  key j not masked ) * V[b,j,h,c]
//
// One CTA tile = 128 query rows of one (utterance, head); keys are visited in blocks of 128:
//   warp 0      TMA producer : Q and K k-chunks (128 rows x 32 fp32, 128B swizzle) for S = Q K^T, then the block's V^T chunks
//                              (dk channel rows x 32 keys) into a 4-stage ring — in exactly the order the issuer consumes them;
//   warp 1      MMA issuer   : S(j+1) = Q K_{j+1}^T is issued as soon as the softmax warps have read S(j) out of TMEM, and
//                              O += P(j) V_j as soon as P(j) is in shared memory: the tensor pipe runs QK^T of the next block
//                              while the softmax of this one is computed (TMEM: S 128 columns + O round16(dk) columns);
//   warps 2..9  softmax      : a query row per pair of threads (TMEM lane = row; warps w and w + 4 split the row's 128 scores and
//                              exchange the block maximum through shared memory).  Online softmax with a
//                              LAZY reference maximum: the row's reference m only moves when the block maximum exceeds it by
//                              more than 2^8 in the exponent, and only then is the O accumulator rescaled in TMEM
//                              (tcgen05.ld / st) — exact in real arithmetic, the final division by the row sum l uses the same m.
//                              P = exp2((s - m) * log2e / T) is rounded to TF32 (round-to-nearest, what the TMA unit does for the
//                              unfused PV product) and written to shared memory in the K-major 128B-swizzled layout the MMA reads.
// Q|K come row-major [B*L, 2H] (one GEMM), V transposed per utterance Vt[b][c][t] (row pitch Lp) — the layouts the unfused
// path already uses (engine.cu fft_block).
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <cmath>
#include <mutex>
#include <string>

namespace zvx {

namespace {

constexpr int AT_THREADS = 320;          // producer warp, issuer warp, 8 softmax warps
constexpr int KB = 112;                  // keys per block = columns of S: S + P + O = 112 + 112 + 272 TMEM columns
constexpr int QT = 128;                  // query rows per tile (MMA M)
constexpr int CH = 32;                   // fp32 per 128-byte swizzle row
constexpr int HK = KB / 2;               // keys per softmax thread
constexpr int AT_STAGES = 6;             // no P buffer in shared memory: the whole 227 KB is operand ring
constexpr int Q_CHUNK_BYTES = QT * CH * 4, K_CHUNK_BYTES = KB * CH * 4;
constexpr int QK_STAGE_BYTES = Q_CHUNK_BYTES + K_CHUNK_BYTES;
constexpr int S_COL = 0, P_COL = KB, O_COL = 2 * KB;   // TMEM columns
constexpr float LAZY_EXP2 = 8.f;         // the reference maximum moves when a block maximum exceeds it by 2^8

struct AttnParams {
    int B, L, n_head, dk, H;
    int qtiles, num_tiles, nkb;
    int kchunks, last_ksteps;
    int NV, n_lo, n_hi;
    int stage_bytes, stages;
    int narrow_last;                     // pair kernel: the last k-chunk of Q / K is 8 columns wide, 32-byte swizzle (d_k % 32 == 8)
    int q_bytes;                         // pair kernel: resident Q tile
    uint32_t idesc_qk, idesc_lo, idesc_hi;
    float sc;                            // log2(e) / temperature
    float* out;
    const uint8_t* mask;
    int mask_ld;
    int dbg_skip;                        // -DZVX_DEBUG experiments (wrong results): 1 = reload no Q after the first key block
    long long* dbg;                      // -DZVX_DEBUG: wait-cycle counters of CTA 0's roles (tools/attn_bench.py)
    const int4* tiles;                   // optional tile list built on the device (attn_plan): {q tile or pair, head, utterance, key blocks}
    const int* ntiles_dev;               // its length
};

__device__ __forceinline__ uint64_t at_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// K-major operand whose rows hold ONE k-step (8 tf32 = 32 bytes, 32-byte swizzle; 8-row groups 256 bytes apart): the last k-chunk
// of a head dimension with d_k % 32 == 8 takes a quarter of the shared memory of a 128-byte-swizzled chunk
__device__ __forceinline__ uint64_t at_sw32_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 8 tf32) is read from TMEM columns [a_tmem, a_tmem + 8), lane = row
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// 2^x, MUFU.EX2 directly (2 ulp; arguments here are <= 8, results below 2^-126 flush to zero)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// -DZVX_DEBUG: cycles a role of CTA 0 spends in a wait, accumulated per wait site
#ifdef ZVX_DEBUG
#define AT_TIMED(acc, stmt) do { const long long t0_ = clock64(); stmt; (acc) += clock64() - t0_; } while (0)
#else
#define AT_TIMED(acc, stmt) do { stmt; } while (0)
#endif

// -DZVX_DEBUG: time stamps of CTA 0's first tile, dbg[16 + role*64 + j*4 + k] (role 0 producer, 1 issuer, 2 softmax warp 2)
#ifdef ZVX_DEBUG
#define AT_STAMP(cond, role, j, k) do { if (p.dbg && blockIdx.x == 0 && (cond) && (j) < 16) p.dbg[16 + (role) * 64 + (j) * 4 + (k)] = clock64(); } while (0)
#else
#define AT_STAMP(cond, role, j, k) do { } while (0)
#endif

// K = 8 steps of block j's PV product (keys < L, rounded up to 8: the rest of P is zero, of V zero-filled)
__device__ __forceinline__ int valid_ksteps(const AttnParams& p, int j) { return (min(p.L - j * KB, KB) + 7) / 8; }
// chunks of 32 keys (one V^T load each) that contain at least one key < L
__device__ __forceinline__ int valid_chunks(const AttnParams& p, int j) {
    const int left = min(p.L - j * KB, KB);
    return (left + CH - 1) / CH;
}

// Tiles of the persistent loops.  Single-CTA kernel: tile = (q tile, head, utterance), CTA i takes tiles i, i + grid, ...
// Pair kernel: tile = (PAIR of adjacent q tiles, head, utterance), cluster i takes i, i + clusters, ...; CTA `rank` of the pair owns
// q tile 2 * pair + rank (p.qtiles counts pairs there).
template <bool PAIR> __device__ __forceinline__ int first_tile() { return PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
template <bool PAIR> __device__ __forceinline__ int tile_stride() { return PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x; }
struct Tile { int qt, h, b, nkb; };
// number of tiles of this launch: the static enumeration, or the device-built list (masked work left out, attn_plan)
__device__ __forceinline__ int tile_count(const AttnParams& p) { return p.tiles ? *p.ntiles_dev : p.num_tiles; }
template <bool PAIR>
__device__ __forceinline__ Tile get_tile(const AttnParams& p, int t, uint32_t rank) {
    Tile tl;
    int pq;
    if (p.tiles) {
        const int4 e = p.tiles[t];
        pq = e.x; tl.h = e.y; tl.b = e.z; tl.nkb = e.w;
    } else {
        pq = t % p.qtiles;
        const int z = t / p.qtiles;
        tl.h = z % p.n_head;
        tl.b = z / p.n_head;
        tl.nkb = p.nkb;
    }
    tl.qt = PAIR ? 2 * pq + (int)rank : pq;
    return tl;
}
// The softmax role (warps 2..9) of both kernels.  s_full / pv_done: this CTA's barriers (completed by tcgen05.commit); s_empty /
// p_full / o_empty: the issuer's barriers (arrive_issuer); aux: the CTA's max / sum exchange buffer.
template <bool PAIR>
__device__ __forceinline__ void softmax_role(const AttnParams& p, uint8_t* aux, uint32_t tmem_base, uint32_t s_full, uint32_t s_empty,
                                             uint32_t p_full, uint32_t pv_done, uint32_t o_empty, uint32_t rank) {
    // ================================================================ softmax warps: a query row per PAIR of threads
    // Warps w and w + 4 own the same TMEM lane quarter (hardware rule: warp id % 4) and split the row's 128 scores /
    // the O columns in halves; they exchange the block maximum through shared memory (one named barrier per block), so
    // both halves keep the same reference maximum; the partial row sums only meet in the epilogue.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    volatile float* xmax = reinterpret_cast<volatile float*>(aux);   // [2][2][128]
    volatile float* xsum = xmax + 2 * 2 * QT;                                                                       // [2][128]
    const int NVh = p.NV >> 1;                         // O columns of this half: [half * NVh, half * NVh + NVh)
    uint32_t g = 0;
    long long w_sfull = 0, w_pv = 0, w_bar = 0, w_epi = 0;
    const long long t_start = clock64();
    const int ntiles = tile_count(p);
    for (int t = first_tile<PAIR>(); t < ntiles; t += tile_stride<PAIR>()) {
        const Tile tl = get_tile<PAIR>(p, t, rank);
        const int h = tl.h, b = tl.b, nkb = tl.nkb;
        const int q = tl.qt * QT + row;
        const uint8_t* km = p.mask ? p.mask + (long long)b * p.mask_ld : nullptr;
        auto key_ok = [&](int j, int i) {
            const int key = j * KB + half * HK + i * 32 + lane;
            return i * 32 + lane < HK && key < p.L && !(km && km[key]);
        };
        float m_ref = -INFINITY, l = 0.f;
        bool ok0 = key_ok(0, 0), ok1 = key_ok(0, 1);
        for (int j = 0; j < nkb; ++j) {
            const uint32_t gq = g + (uint32_t)j;
            // key validity of this half of the block (bounds + padding mask), one bit per key, the same words in every lane
            const uint32_t vw0 = __ballot_sync(0xFFFFFFFFu, ok0), vw1 = __ballot_sync(0xFFFFFFFFu, ok1);
            const bool all_valid = vw0 == 0xFFFFFFFFu && vw1 == (0xFFFFFFFFu >> (64 - HK));

            AT_TIMED(w_sfull, mbar_wait_hint(s_full, gq & 1u, 32));
            AT_STAMP(g == 0 && threadIdx.x == 64, 2, j, 0);
            tc_fence_after();
            float v[HK];
            static_assert(HK % 16 == 8, "56 columns per thread: three x16 loads and one x8");
#pragma unroll
            for (int i = 0; i < HK / 16; ++i)
                tmem_ld16(trow + (uint32_t)(S_COL + half * HK + i * 16), reinterpret_cast<uint32_t*>(v + i * 16));
            tmem_ld8(trow + (uint32_t)(S_COL + half * HK + HK - 8), reinterpret_cast<uint32_t*>(v + HK - 8));
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_issuer<PAIR>(s_empty);                 // the issuer may overwrite S with the next block's scores
            AT_STAMP(g == 0 && threadIdx.x == 64, 2, j, 1);
            if (j + 1 < nkb) { ok0 = key_ok(j + 1, 0); ok1 = key_ok(j + 1, 1); }   // (the loads fly during the arithmetic)

            float mb0 = -INFINITY, mb1 = -INFINITY;
            if (all_valid) {
#pragma unroll
                for (int i = 0; i < HK; i += 2) { mb0 = fmaxf(mb0, v[i]); mb1 = fmaxf(mb1, v[i + 1]); }
            } else {
#pragma unroll
                for (int i = 0; i < HK; i += 2) {
                    if (!(((i < 32 ? vw0 : vw1) >> (i & 31)) & 1u)) v[i] = -INFINITY;
                    if (!(((i < 32 ? vw0 : vw1) >> ((i + 1) & 31)) & 1u)) v[i + 1] = -INFINITY;
                    mb0 = fmaxf(mb0, v[i]); mb1 = fmaxf(mb1, v[i + 1]);
                }
            }
            float mb = fmaxf(mb0, mb1);
            xmax[((gq & 1u) * 2 + (uint32_t)half) * QT + row] = mb;
            AT_TIMED(w_bar, asm volatile("bar.sync 1, 256;" ::: "memory"));
            mb = fmaxf(mb, xmax[((gq & 1u) * 2 + (uint32_t)(half ^ 1)) * QT + row]);
            // lazy reference maximum (in units of the exponent: s * sc)
            float factor = 1.f;
            bool rescale = false;
            if (mb * p.sc > m_ref * p.sc + LAZY_EXP2) {          // (also the first finite maximum: m_ref = -inf)
                if (m_ref != -INFINITY) {
                    factor = ex2_approx((m_ref - mb) * p.sc);
                    rescale = j > 0;
                }
                m_ref = mb;
            }
            const float off = (m_ref == -INFINITY) ? 0.f : m_ref * p.sc;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < HK; i += 4) {
                const float e0 = ex2_approx(fmaf(v[i], p.sc, -off)), e1 = ex2_approx(fmaf(v[i + 1], p.sc, -off));
                const float e2 = ex2_approx(fmaf(v[i + 2], p.sc, -off)), e3 = ex2_approx(fmaf(v[i + 3], p.sc, -off));
                s0 += e0; s1 += e1; s2 += e2; s3 += e3;
                v[i] = rn_tf32(e0); v[i + 1] = rn_tf32(e1); v[i + 2] = rn_tf32(e2); v[i + 3] = rn_tf32(e3);
            }
            l = fmaf(l, factor, (s0 + s1) + (s2 + s3));

            // P buffer free and O stable: the previous block's PV product has completed
            AT_STAMP(g == 0 && threadIdx.x == 64, 2, j, 2);
            if (gq > 0) AT_TIMED(w_pv, mbar_wait_hint(pv_done, (gq - 1) & 1u, 32));
            tc_fence_after();
            if (__any_sync(0xFFFFFFFFu, rescale)) {
                for (int c = half * NVh; c < (half + 1) * NVh; c += 8) {
                    uint32_t o[8];
                    tmem_ld8(trow + (uint32_t)(O_COL + c), o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                    tmem_st8(trow + (uint32_t)(O_COL + c), o);
                }
                tmem_wait_st();
            }
            // P goes to TMEM (the PV product reads its A operand there: lane = query row, column = key)
#pragma unroll
            for (int i = 0; i < HK / 16; ++i)
                tmem_st16(trow + (uint32_t)(P_COL + half * HK + i * 16), reinterpret_cast<const uint32_t*>(v + i * 16));
            tmem_st8(trow + (uint32_t)(P_COL + half * HK + HK - 8), reinterpret_cast<const uint32_t*>(v + HK - 8));
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_issuer<PAIR>(p_full);
            AT_STAMP(g == 0 && threadIdx.x == 64, 2, j, 3);
        }
        g += (uint32_t)nkb;
        // ---- epilogue: O / l -> out[b, q, h*dk + :], this half's columns; the two partial row sums meet here
        xsum[half * QT + row] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.f / (l + xsum[(half ^ 1) * QT + row]);
        AT_TIMED(w_epi, mbar_wait_hint(pv_done, (g - 1) & 1u, 32));
        tc_fence_after();
        float* orow = p.out + ((long long)b * p.L + q) * p.H + (long long)h * p.dk;
        const int c_end = min((half + 1) * NVh, p.dk);
        for (int c = half * NVh; c < c_end; c += 32) {
            uint32_t o[32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (c + i * 8 < c_end) tmem_ld8(trow + (uint32_t)(O_COL + c + i * 8), o + i * 8);
            tmem_wait_ld();
            if (q < p.L) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (c + i < c_end)
                        *reinterpret_cast<float4*>(orow + c + i) =
                            make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, __uint_as_float(o[i + 2]) * inv,
                                        __uint_as_float(o[i + 3]) * inv);
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_issuer<PAIR>(o_empty);
        asm volatile("bar.sync 1, 256;" ::: "memory");          // xsum is rewritten by the next tile
    }
    if (ZVX_DBG_PTR(p) && blockIdx.x == 0 && threadIdx.x == 64) {
        long long* d = ZVX_DBG_PTR(p);
        d[8] = clock64() - t_start; d[9] = w_sfull; d[10] = w_pv; d[11] = w_bar; d[12] = w_epi;
    }
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapVlo, const __grid_constant__ CUtensorMap mapVhi, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sStage = smem_u32(smem);
    const uint32_t bars = sStage + (uint32_t)(AT_STAGES * p.stage_bytes);
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(AT_STAGES + s); };
    const uint32_t s_full = bars + 8u * (2 * AT_STAGES + 0), s_empty = bars + 8u * (2 * AT_STAGES + 1);
    const uint32_t p_full = bars + 8u * (2 * AT_STAGES + 2), pv_done = bars + 8u * (2 * AT_STAGES + 3);
    const uint32_t o_empty = bars + 8u * (2 * AT_STAGES + 4);
    const uint32_t tmem_slot = bars + 8u * (2 * AT_STAGES + 5);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem + AT_STAGES * p.stage_bytes + 8 * (2 * AT_STAGES + 5));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&mapQ);
        prefetch_tmap(&mapK);
        prefetch_tmap(&mapVlo);
        prefetch_tmap(&mapVhi);
        for (int s = 0; s < AT_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 8);
        mbar_init(p_full, 8);
        mbar_init(pv_done, 1);
        mbar_init(o_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t DESC_HI = (uint32_t)(at_sw128_desc(0) >> 32);

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            long long w_empty = 0;
            const long long t_start = clock64();
            const int ntiles = tile_count(p);
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const Tile tl = get_tile<false>(p, t, 0u);
                const int qt = tl.qt, h = tl.h, b = tl.b, nkb = tl.nkb;
                const bool first = t == (int)blockIdx.x;
                auto load_qk = [&](int j) {
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        AT_TIMED(w_empty, mbar_wait_spin(empty_bar(stage), phase ^ 1u));
                        if (kc == 0) AT_STAMP(first, 0, j, 0);
                        if (kc == p.kchunks - 1) AT_STAMP(first, 0, j, 1);
                        const uint32_t fb = full_bar(stage), dst = sStage + (uint32_t)(stage * p.stage_bytes);
                        const bool skip_q = (ZVX_DBG_SKIP(p) & 1) && j > 0;
                        mbar_arrive_expect_tx(fb, (uint32_t)(skip_q ? K_CHUNK_BYTES : QK_STAGE_BYTES));
                        if (!skip_q) tma_load_4d(&mapQ, fb, dst, kc * CH, qt * QT, h, b);
                        tma_load_4d(&mapK, fb, dst + (uint32_t)Q_CHUNK_BYTES, kc * CH, j * KB, h, b);
                        if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                };
                auto load_v = [&](int j) {
                    const int nc = valid_chunks(p, j);
                    for (int c = 0; c < nc; ++c) {
                        AT_TIMED(w_empty, mbar_wait_spin(empty_bar(stage), phase ^ 1u));
                        if (c == 0) AT_STAMP(first, 0, j, 2);
                        if (c == nc - 1) AT_STAMP(first, 0, j, 3);
                        const uint32_t fb = full_bar(stage), dst = sStage + (uint32_t)(stage * p.stage_bytes);
                        mbar_arrive_expect_tx(fb, (uint32_t)(p.NV * CH * 4));
                        tma_load_4d(&mapVlo, fb, dst, j * KB + c * CH, 0, h, b);
                        if (p.n_hi) tma_load_4d(&mapVhi, fb, dst + (uint32_t)(p.n_lo * CH * 4), j * KB + c * CH, p.n_lo, h, b);
                        if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                };
                load_qk(0);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) load_qk(j + 1);
                    load_v(j);
                }
            }
            if (ZVX_DBG_PTR(p) && blockIdx.x == 0) { ZVX_DBG_PTR(p)[0] = clock64() - t_start; ZVX_DBG_PTR(p)[1] = w_empty; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t g = 0, tcount = 0;          // key blocks / tiles this CTA has started
            long long w_sempty = 0, w_pfull = 0, w_oempty = 0, w_full_qk = 0, w_full_v = 0;
            const long long t_start = clock64();
            const uint32_t dS = tmem_base + S_COL, dO = tmem_base + O_COL, dP = tmem_base + P_COL;
            const uint32_t st_lo0 = (uint32_t)(at_sw128_desc(sStage) & 0xFFFFFFFFull);
            const int ntiles = tile_count(p);
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
                const int nkb = get_tile<false>(p, t, 0u).nkb;
                const bool first = tcount == 0;
                auto issue_qk = [&](uint32_t gq) {
                    if (gq > 0) AT_TIMED(w_sempty, mbar_wait_spin(s_empty, (gq - 1) & 1u));   // S(gq-1) has been read out of TMEM
                    tc_fence_after();
                    AT_STAMP(first, 1, (int)(gq - g), 0);
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        AT_TIMED(w_full_qk, mbar_wait_spin(full_bar(stage), phase));
                        const uint32_t a = st_lo0 + (uint32_t)((stage * p.stage_bytes) >> 4), bb = a + (uint32_t)(Q_CHUNK_BYTES >> 4);
                        const int nk = (kc == p.kchunks - 1) ? p.last_ksteps : CH / 8;
                        for (int k = 0; k < nk; ++k)
                            umma_tf32_lo(dS, a + 2 * k, bb + 2 * k, DESC_HI, p.idesc_qk, (kc | k) ? 1u : 0u);
                        umma_commit(empty_bar(stage));
                        if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(s_full);
                    AT_STAMP(first, 1, (int)(gq - g), 1);
                };
                auto issue_pv = [&](int j, uint32_t gq) {
                    AT_TIMED(w_pfull, mbar_wait_spin(p_full, gq & 1u));   // P(gq) is in shared memory (and O rescaled if needed)
                    if (j == 0 && tcount > 0) AT_TIMED(w_oempty, mbar_wait_spin(o_empty, (tcount - 1) & 1u));   // previous tile's O has been read
                    tc_fence_after();
                    AT_STAMP(first, 1, j, 2);
                    const int nc = valid_chunks(p, j), nks = valid_ksteps(p, j);
                    for (int c = 0; c < nc; ++c) {
                        AT_TIMED(w_full_v, mbar_wait_spin(full_bar(stage), phase));
                        const uint32_t a = dP + (uint32_t)(c * CH);                 // P columns of this chunk's keys
                        const uint32_t bb = st_lo0 + (uint32_t)((stage * p.stage_bytes) >> 4);
                        const uint32_t bh = bb + (uint32_t)((p.n_lo * CH * 4) >> 4);
                        const int nk = min(CH / 8, nks - c * (CH / 8));
                        for (int k = 0; k < nk; ++k) {
                            const uint32_t acc = (j | c | k) ? 1u : 0u;
                            umma_tf32_ts(dO, a + 8 * k, bb + 2 * k, DESC_HI, p.idesc_lo, acc);
                            if (p.n_hi) umma_tf32_ts(dO + (uint32_t)p.n_lo, a + 8 * k, bh + 2 * k, DESC_HI, p.idesc_hi, acc);
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == AT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(pv_done);
                    AT_STAMP(first, 1, j, 3);
                };
                issue_qk(g);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) issue_qk(g + (uint32_t)j + 1u);
                    issue_pv(j, g + (uint32_t)j);
                }
                g += (uint32_t)nkb;
            }
            if (ZVX_DBG_PTR(p) && blockIdx.x == 0) {
                long long* d = ZVX_DBG_PTR(p);
                d[2] = clock64() - t_start; d[3] = w_sempty; d[4] = w_pfull; d[5] = w_oempty; d[6] = w_full_qk; d[7] = w_full_v;
            }
        }
        __syncwarp();
    } else {
        softmax_role<false>(p, smem + AT_STAGES * p.stage_bytes + 128, tmem_base, s_full, s_empty, p_full, pv_done, o_empty, 0u);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// Two CTAs of a cluster (the two SMs of a TPC) take two adjacent 128-row query tiles of one (utterance, head) and run every
// product as ONE tcgen05.mma.cta_group::2 (M = 256): the B operand — the K block of S = Q K^T, the V^T block of O += P V — is
// split between the two CTAs' shared memories (56 keys / half of the value channels each), so each SM fetches HALF of K and V
// per block.  That is what the single-CTA kernel cannot have: its operand ring (K + V + the re-streamed Q, 375 KB per block of
// 3.8 k tensor-pipe cycles) is bound by fetch latency x ring depth (profiles/r02_attn_fused_timeline.txt).  Here Q (128 x dk,
// 144 KB) stays RESIDENT in each CTA for the whole tile and the ring carries 131 KB per block and CTA.
//   warp 0 (both CTAs)  TMA producer of its CTA's halves; all bytes are counted on the LEADER's barriers (.cta_group::2 loads)
//   warp 1 (leader)     issues every MMA of the pair; tcgen05.commit multicasts "stage free" / "S ready" / "PV done" to both CTAs
//   warps 2..9          softmax of the CTA's own 128 rows (softmax_role<true>): arrivals go to the leader's barriers
constexpr int KH_CHUNK_BYTES = (KB / 2) * CH * 4;   // this CTA's 56 keys of a 32-wide k-chunk
constexpr int AP_MAX_STAGES = 8;

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapVlo, const __grid_constant__ CUtensorMap mapVhi,
                 const __grid_constant__ CUtensorMap mapQ8, const __grid_constant__ CUtensorMap mapK8, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int q_bytes = p.q_bytes;
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sRing = sQ + (uint32_t)q_bytes;
    const int bar_off = q_bytes + p.stages * p.stage_bytes;
    const uint32_t bars = sQ + (uint32_t)bar_off;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(AP_MAX_STAGES + s); };
    const uint32_t s_full = bars + 8u * (2 * AP_MAX_STAGES + 0), s_empty = bars + 8u * (2 * AP_MAX_STAGES + 1);
    const uint32_t p_full = bars + 8u * (2 * AP_MAX_STAGES + 2), pv_done = bars + 8u * (2 * AP_MAX_STAGES + 3);
    const uint32_t o_empty = bars + 8u * (2 * AP_MAX_STAGES + 4);
    const uint32_t q_full = bars + 8u * (2 * AP_MAX_STAGES + 5), q_empty = bars + 8u * (2 * AP_MAX_STAGES + 6);
    const uint32_t tmem_slot = bars + 8u * (2 * AP_MAX_STAGES + 7);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (2 * AP_MAX_STAGES + 7));

    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        prefetch_tmap(&mapQ);
        prefetch_tmap(&mapK);
        prefetch_tmap(&mapVlo);
        prefetch_tmap(&mapVhi);
        prefetch_tmap(&mapQ8);
        prefetch_tmap(&mapK8);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);          // the leader's producer arrives (expect_tx of both CTAs' bytes)
            mbar_init(empty_bar(s), 1);         // one multicast commit
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 16);                 // 8 softmax warps of each CTA
        mbar_init(p_full, 16);
        mbar_init(pv_done, 1);
        mbar_init(o_empty, 16);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();   // (the tile list and Q / K / V come from the launches before this one)
    const uint32_t DESC_HI = (uint32_t)(at_sw128_desc(0) >> 32), DESC_HI32 = (uint32_t)(at_sw32_desc(0) >> 32);

    if (warp == 0) {
        // ================================================================ TMA producer (both CTAs, the same sequence)
        if (elect_one()) {
            const uint32_t L_full0 = mapa_u32(full_bar(0), 0), L_qfull = mapa_u32(q_full, 0);
            int stage = 0;
            uint32_t phase = 0, tcount = 0;
            long long w_empty = 0;
            const long long t_start = clock64();
            auto advance = [&]() { if (++stage == p.stages) { stage = 0; phase ^= 1u; } };
            const int ntiles = tile_count(p);
            auto load_q = [&](int t, uint32_t tc) {
                const Tile tq = get_tile<true>(p, t, rank);
                const int qt = tq.qt, h = tq.h, b = tq.b;
                if (tc > 0) mbar_wait_spin(q_empty, (tc - 1) & 1u);       // every QK^T of the previous tile has completed
                if (leader) mbar_arrive_expect_tx(q_full, (uint32_t)(2 * q_bytes));
                const int nfull = p.kchunks - p.narrow_last;
                for (int kc = 0; kc < nfull; ++kc)
                    tma_load_4d_pair(&mapQ, L_qfull, sQ + (uint32_t)(kc * Q_CHUNK_BYTES), kc * CH, qt * QT, h, b);
                if (p.narrow_last) tma_load_4d_pair(&mapQ8, L_qfull, sQ + (uint32_t)(nfull * Q_CHUNK_BYTES), nfull * CH, qt * QT, h, b);
            };
            const int t0 = first_tile<true>(), dt = tile_stride<true>();
            if (t0 < ntiles) load_q(t0, 0);
            for (int t = t0; t < ntiles; t += dt, ++tcount) {
                const Tile tl = get_tile<true>(p, t, rank);
                const int h = tl.h, b = tl.b, nkb = tl.nkb;
                auto load_k = [&](int j) {                                // two k-chunks of this CTA's 56 keys per stage
                    for (int kc = 0; kc < p.kchunks; kc += 2) {
                        AT_TIMED(w_empty, mbar_wait_spin(empty_bar(stage), phase ^ 1u));
                        if (kc == 0) AT_STAMP(tcount == 0, 0, j, 0);
                        if (kc + 2 >= p.kchunks) AT_STAMP(tcount == 0, 0, j, 1);
                        const int nch = min(2, p.kchunks - kc);
                        const uint32_t dst = sRing + (uint32_t)(stage * p.stage_bytes);
                        const bool narrow = p.narrow_last && kc + nch == p.kchunks;   // the stage's last chunk is the 8-column one
                        if (leader)
                            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(2 * (nch * KH_CHUNK_BYTES - (narrow ? KH_CHUNK_BYTES * 3 / 4 : 0))));
                        for (int c = 0; c < nch; ++c)
                            tma_load_4d_pair((narrow && c == nch - 1) ? &mapK8 : &mapK, L_full0 + 8u * (uint32_t)stage,
                                             dst + (uint32_t)(c * KH_CHUNK_BYTES), (kc + c) * CH, j * KB + (int)rank * (KB / 2), h, b);
                        advance();
                    }
                };
                auto load_v = [&](int j) {                                // this CTA's halves of the two V^T column groups, 32 keys
                    const int nc = valid_chunks(p, j);
                    for (int c = 0; c < nc; ++c) {
                        AT_TIMED(w_empty, mbar_wait_spin(empty_bar(stage), phase ^ 1u));
                        if (c == 0) AT_STAMP(tcount == 0, 0, j, 2);
                        if (c == nc - 1) AT_STAMP(tcount == 0, 0, j, 3);
                        const uint32_t dst = sRing + (uint32_t)(stage * p.stage_bytes);
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.NV * CH * 4));
                        tma_load_4d_pair(&mapVlo, L_full0 + 8u * (uint32_t)stage, dst, j * KB + c * CH, (int)rank * (p.n_lo / 2), h, b);
                        if (p.n_hi)
                            tma_load_4d_pair(&mapVhi, L_full0 + 8u * (uint32_t)stage, dst + (uint32_t)((p.n_lo / 2) * CH * 4), j * KB + c * CH,
                                             p.n_lo + (int)rank * (p.n_hi / 2), h, b);
                        advance();
                    }
                };
                load_k(0);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) load_k(j + 1);
                    // the next tile's Q goes out before the last V block: by then every QK^T of this tile has been issued and
                    // the transfer overlaps the last two PV products
                    if (j == nkb - 1 && t + dt < ntiles) load_q(t + dt, tcount + 1);
                    load_v(j);
                }
            }
            if (ZVX_DBG_PTR(p) && blockIdx.x == 0) { ZVX_DBG_PTR(p)[0] = clock64() - t_start; ZVX_DBG_PTR(p)[1] = w_empty; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer (leader CTA only)
        if (leader && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t g = 0, tcount = 0;
            long long w_sempty = 0, w_pfull = 0, w_oempty = 0, w_full_qk = 0, w_full_v = 0, w_qfull = 0;
            const long long t_start = clock64();
            auto advance = [&]() { if (++stage == p.stages) { stage = 0; phase ^= 1u; } };
            const uint32_t dS = tmem_base + S_COL, dO = tmem_base + O_COL, dP = tmem_base + P_COL;
            const uint32_t q_lo0 = (uint32_t)(at_sw128_desc(sQ) & 0xFFFFFFFFull);
            const uint32_t ring_lo0 = (uint32_t)(at_sw128_desc(sRing) & 0xFFFFFFFFull);
            const int ntiles = tile_count(p);
            for (int t = first_tile<true>(); t < ntiles; t += tile_stride<true>(), ++tcount) {
                const int nkb = get_tile<true>(p, t, 0u).nkb;
                auto issue_qk = [&](uint32_t gq, bool last) {
                    if (gq > 0) AT_TIMED(w_sempty, mbar_wait_spin_cluster(s_empty, (gq - 1) & 1u));   // both CTAs have read S(gq - 1) out of TMEM
                    tc_fence_after();
                    AT_STAMP(tcount == 0, 1, (int)(gq - g), 0);
                    for (int kc = 0; kc < p.kchunks; kc += 2) {
                        AT_TIMED(w_full_qk, mbar_wait_spin(full_bar(stage), phase));
                        tc_fence_after();
                        const int nch = min(2, p.kchunks - kc);
                        for (int c = 0; c < nch; ++c) {
                            const uint32_t a = q_lo0 + (uint32_t)(((kc + c) * Q_CHUNK_BYTES) >> 4);
                            const uint32_t bb = ring_lo0 + (uint32_t)((stage * p.stage_bytes + c * KH_CHUNK_BYTES) >> 4);
                            const bool lastc = kc + c == p.kchunks - 1;
                            const int nk = lastc ? p.last_ksteps : CH / 8;
                            const uint32_t dhi = (lastc && p.narrow_last) ? DESC_HI32 : DESC_HI;
                            for (int k = 0; k < nk; ++k)
                                umma_pair_tf32_ss(dS, a + 2 * k, bb + 2 * k, dhi, p.idesc_qk, (kc | c | k) ? 1u : 0u);
                        }
                        umma_commit_pair(empty_bar(stage));
                        advance();
                    }
                    umma_commit_pair(s_full);
                    AT_STAMP(tcount == 0, 1, (int)(gq - g), 1);
                    if (last) umma_commit_pair(q_empty);                             // Q may be replaced by the next tile's
                };
                auto issue_pv = [&](int j, uint32_t gq) {
                    AT_TIMED(w_pfull, mbar_wait_spin_cluster(p_full, gq & 1u));      // P(gq) is in both CTAs' TMEM
                    if (j == 0 && tcount > 0) AT_TIMED(w_oempty, mbar_wait_spin_cluster(o_empty, (tcount - 1) & 1u));   // the previous tile's O has been read
                    tc_fence_after();
                    AT_STAMP(tcount == 0, 1, j, 2);
                    const int nc = valid_chunks(p, j), nks = valid_ksteps(p, j);
                    for (int c = 0; c < nc; ++c) {
                        AT_TIMED(w_full_v, mbar_wait_spin(full_bar(stage), phase));
                        tc_fence_after();
                        const uint32_t a = dP + (uint32_t)(c * CH);
                        const uint32_t bb = ring_lo0 + (uint32_t)((stage * p.stage_bytes) >> 4);
                        const uint32_t bh = bb + (uint32_t)(((p.n_lo / 2) * CH * 4) >> 4);
                        const int nk = min(CH / 8, nks - c * (CH / 8));
                        for (int k = 0; k < nk; ++k) {
                            const uint32_t acc = (j | c | k) ? 1u : 0u;
                            umma_pair_tf32_ts(dO, a + 8 * k, bb + 2 * k, DESC_HI, p.idesc_lo, acc);
                            if (p.n_hi) umma_pair_tf32_ts(dO + (uint32_t)p.n_lo, a + 8 * k, bh + 2 * k, DESC_HI, p.idesc_hi, acc);
                        }
                        umma_commit_pair(empty_bar(stage));
                        advance();
                    }
                    umma_commit_pair(pv_done);
                    AT_STAMP(tcount == 0, 1, j, 3);
                };
                AT_TIMED(w_qfull, mbar_wait_spin(q_full, tcount & 1u));              // both CTAs' Q tiles have landed
                issue_qk(g, nkb == 1);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) issue_qk(g + (uint32_t)j + 1u, j + 2 == nkb);
                    issue_pv(j, g + (uint32_t)j);
                }
                g += (uint32_t)nkb;
            }
            if (ZVX_DBG_PTR(p) && blockIdx.x == 0) {
                long long* d = ZVX_DBG_PTR(p);
                d[2] = clock64() - t_start; d[3] = w_sempty; d[4] = w_pfull; d[5] = w_oempty; d[6] = w_full_qk; d[7] = w_full_v;
                d[13] = w_qfull;
            }
        }
        __syncwarp();
    } else {
        softmax_role<true>(p, smem + bar_off + 256, tmem_base, s_full, mapa_u32(s_empty, 0), mapa_u32(p_full, 0), pv_done,
                           mapa_u32(o_empty, 0), rank);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ tile plan
// ws: [0] number of tiles, [4 .. 4 + B) keys per utterance (1 + last unmasked position), then the int4 tile list
// {q tile (pair) index, head, utterance, key blocks}, utterance-major.
__global__ void __launch_bounds__(1024)
attn_plan_kernel(const uint8_t* __restrict__ mask, int mask_ld, int B, int L, int n_head, int rows_per_tile, int skip_q, int* ws) {
    int* kv_len = ws + 4;
    int4* tiles = reinterpret_cast<int4*>(ws + 4 + (B + 3) / 4 * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = (L + 31) / 32;
    for (int b = warp; b < B; b += 32) {
        int last = L;
        if (mask) {
            last = 0;
            const uint8_t* m = mask + (long long)b * mask_ld;
            const int k1 = min(L, (lane + 1) * seg);
            for (int k = lane * seg; k < k1; ++k)
                if (!m[k]) last = k + 1;
            for (int o = 16; o; o >>= 1) last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
        }
        if (lane == 0) kv_len[b] = last;
    }
    __shared__ int s_scan[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int qfull = (L + rows_per_tile - 1) / rows_per_tile;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + (int)threadIdx.x;
        const int kv = b < B ? kv_len[b] : 0;
        const int nq = b < B ? (skip_q ? (kv + rows_per_tile - 1) / rows_per_tile : qfull) : 0;
        const int cnt = nq * n_head;
        s_scan[threadIdx.x] = cnt;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const int v = (int)threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int base = s_base + s_scan[threadIdx.x] - cnt;
        const int nkb = max(1, (kv + KB - 1) / KB);
        for (int i = 0; i < cnt; ++i) tiles[base + i] = make_int4(i % nq, i / nq, b, nkb);
        __syncthreads();
        if (threadIdx.x == 1023) s_base += s_scan[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) ws[0] = s_base;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn at_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    });
    if (!fn) throw Error("cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

// 4-D fp32 view, dim 0 contiguous; TFLOAT32 element type (round-to-nearest in flight), 128B swizzle, zero fill out of bounds
CUtensorMap at_map(const float* base, const long long dims[4], const long long strides_elems[3], int box0, int box1,
                   CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    CUtensorMap m;
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1}, es[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) gd[i] = (cuuint64_t)std::max<long long>(dims[i], 1);
    long long prev = 16;
    for (int i = 0; i < 3; ++i) {
        long long s = strides_elems[i] * 4;
        if (dims[i + 1] <= 1 && s <= 0) s = prev;
        gs[i] = (cuuint64_t)s;
        prev = std::max<long long>(s, 16);
    }
    const CUresult rc = at_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), gd, gs, bx, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof(buf), "attention: cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld %lld box %d %d", (int)rc,
                 dims[0], dims[1], dims[2], dims[3], box0, box1);
        throw Error(buf);
    }
    return m;
}

uint32_t at_idesc(int n, int m = QT) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

int at_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        ZVX_CUDA_CHECK(cudaGetDevice(&dev));
        ZVX_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

// -DZVX_DEBUG with ZVX_ATTN_DBG set: wait-cycle counters and the first tile's time stamps of CTA 0, printed after every launch
long long* g_at_dbg_buf = nullptr;
bool at_dbg_on() { static const bool on = env_set("ZVX_ATTN_DBG"); return on; }
void at_dbg_begin(AttnParams& p, cudaStream_t st) {
    if (!at_dbg_on()) return;
    if (!g_at_dbg_buf) ZVX_CUDA_CHECK(cudaMalloc(&g_at_dbg_buf, 256 * sizeof(long long)));
    ZVX_CUDA_CHECK(cudaMemsetAsync(g_at_dbg_buf, 0, 256 * sizeof(long long), st));
    p.dbg = g_at_dbg_buf;
}
void at_dbg_end(const AttnParams& p, cudaStream_t st) {
    if (!at_dbg_on()) return;
    long long hst[256];
    ZVX_CUDA_CHECK(cudaMemcpyAsync(hst, g_at_dbg_buf, sizeof(hst), cudaMemcpyDeviceToHost, st));
    ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[attn dbg] CTA0 cycles: producer total %lld wait_empty %lld | issuer total %lld wait s_empty %lld p_full %lld "
                    "o_empty %lld full(qk) %lld full(v) %lld q_full %lld | softmax total %lld wait s_full %lld pv_done %lld bar %lld epi %lld\n",
            hst[0], hst[1], hst[2], hst[3], hst[4], hst[5], hst[6], hst[7], hst[13], hst[8], hst[9], hst[10], hst[11], hst[12]);
    const long long t0 = hst[16 + 64];   // issuer: QK(0) issue start
    for (int j = 0; j < std::min(p.nkb, 16); ++j) {
        const long long* a = hst + 16 + j * 4;
        fprintf(stderr, "[attn dbg] blk %2d | load qk %6lld..%6lld v %6lld..%6lld | issue qk %6lld..%6lld pv %6lld..%6lld | softmax s_full %6lld "
                        "s_empty %6lld computed %6lld p_full %6lld\n", j, a[0] - t0, a[1] - t0, a[2] - t0, a[3] - t0, a[64] - t0, a[65] - t0,
                a[66] - t0, a[67] - t0, a[128] - t0, a[129] - t0, a[130] - t0, a[131] - t0);
    }
}

bool at_common_ok(const AttnFusedArgs& a) {
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    return a.dk % 8 == 0 && a.dk >= 16 && a.H % 4 == 0 && a.Lp % 4 == 0 && a.Lp >= a.L && a.L >= 1 && a.B >= 1 && al16(a.qk) &&
           al16(a.vt) && al16(a.out);
}

bool at_single_supported(const AttnFusedArgs& a) {
    const int NV = (a.dk + 15) / 16 * 16;
    return at_common_ok(a) && NV <= 512 - 2 * KB && (NV <= 256 || NV - 128 >= 16);
}

// The pair kernel's plan: value columns rounded up to 32 (tcgen05.mma.cta_group::2 with A from TMEM: N % 32 == 0), split in at
// most two MMAs; Q resident, the rest of the 227 KB as operand ring.
struct PairPlan { int NV, n_lo, n_hi, stage_bytes, stages, q_bytes, narrow_last; size_t smem; };
constexpr int AP_AUX_BYTES = 256 + 6 * QT * 4;   // barriers; max / sum exchange

bool at_pair_plan(const AttnFusedArgs& a, PairPlan& pl) {
    if (!at_common_ok(a)) return false;
    pl.NV = (a.dk + 31) / 32 * 32;
    if (pl.NV > 512 - 2 * KB) return false;
    if (pl.NV <= 256) { pl.n_lo = pl.NV; pl.n_hi = 0; }
    else { pl.n_hi = 128; pl.n_lo = pl.NV - 128; }
    const int kchunks = cdiv(a.dk, CH);
    pl.narrow_last = (a.dk % CH == 8) ? 1 : 0;
    pl.q_bytes = kchunks * Q_CHUNK_BYTES - (pl.narrow_last ? Q_CHUNK_BYTES * 3 / 4 : 0);
    pl.stage_bytes = (int)round_up((long long)std::max(2 * KH_CHUNK_BYTES, (pl.NV / 2) * CH * 4), 1024);
    const long long room = 227LL * 1024 - 1024 - pl.q_bytes - AP_AUX_BYTES;
    pl.stages = (int)std::min<long long>(AP_MAX_STAGES, room / pl.stage_bytes);
    pl.smem = 1024 + (size_t)pl.q_bytes + (size_t)pl.stages * pl.stage_bytes + AP_AUX_BYTES;
    return pl.stages >= 3;
}

// clusters of two CTAs that can be co-resident (1 CTA / SM, the two SMs of a TPC)
int at_pair_clusters(size_t smem) {
    static int n = -1;
    static size_t n_smem = 0;
    if (n < 0 || n_smem != smem) {
        static std::once_flag once;   // the occupancy query counts against the kernel's opted-in shared-memory limit
        std::call_once(once, [] {
            ZVX_CUDA_CHECK(cudaFuncSetAttribute(attn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        });
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(at_num_sms() / 2 * 2));
        cfg.blockDim = dim3(AT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, attn_pair_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
        n = nc; n_smem = smem;
    }
    return n;
}

void at_maps(const AttnFusedArgs& a, int q_rows, int k_rows, int vlo_rows, int vhi_rows, CUtensorMap& mapQ, CUtensorMap& mapK,
             CUtensorMap& mapVlo, CUtensorMap& mapVhi) {
    const long long H2 = 2LL * a.H;
    const long long qdims[4] = {a.dk, a.L, a.n_head, a.B};
    const long long qstr[3] = {H2, a.dk, (long long)a.L * H2};
    mapQ = at_map(a.qk, qdims, qstr, CH, q_rows);
    mapK = at_map(a.qk + a.H, qdims, qstr, CH, k_rows);
    const long long vdims[4] = {a.L, a.dk, a.n_head, a.B};
    const long long vstr[3] = {a.Lp, (long long)a.dk * a.Lp, (long long)a.H * a.Lp};
    mapVlo = at_map(a.vt, vdims, vstr, CH, vlo_rows);
    mapVhi = vhi_rows ? at_map(a.vt, vdims, vstr, CH, vhi_rows) : mapVlo;
}

void at_fill_common(const AttnFusedArgs& a, AttnParams& p) {
    p.B = a.B; p.L = a.L; p.n_head = a.n_head; p.dk = a.dk; p.H = a.H;
    p.nkb = cdiv(a.L, KB);
    p.kchunks = cdiv(a.dk, CH);
    p.last_ksteps = (a.dk - (p.kchunks - 1) * CH) / 8;
    p.sc = 1.4426950408889634f / a.temperature;
    p.out = a.out; p.mask = a.key_mask; p.mask_ld = a.mask_ld;
    p.dbg = nullptr;
    p.dbg_skip = 0;
}

// variant 0: the CTA-pair kernel whenever an utterance has at least two query tiles (a pair whose second tile is empty would do
// the single kernel's work on two SMs)
bool at_use_pair(const AttnFusedArgs& a, PairPlan& pl) {
    if (a.variant == 1) return false;
    if (a.variant == 0 && cdiv(a.L, QT) < 2) return false;
    return at_pair_plan(a, pl) && at_pair_clusters(pl.smem) >= 1;
}

void at_use_plan(const AttnFusedArgs& a, AttnParams& p, int rows_per_tile) {
    if (!a.plan) return;
    ZVX_REQUIRE(a.plan->dev && a.plan->rows_per_tile == rows_per_tile, "attn_fused: the plan was built for the other kernel variant");
    p.ntiles_dev = a.plan->dev;
    p.tiles = reinterpret_cast<const int4*>(a.plan->dev + 4 + (a.B + 3) / 4 * 4);
}

void attn_pair_launch(const AttnFusedArgs& a, const PairPlan& pl, cudaStream_t st) {
    AttnParams p{};
    at_fill_common(a, p);
    p.qtiles = cdiv(cdiv(a.L, QT), 2);                  // pairs of q tiles
    p.num_tiles = p.qtiles * a.n_head * a.B;
    p.NV = pl.NV; p.n_lo = pl.n_lo; p.n_hi = pl.n_hi;
    p.stage_bytes = pl.stage_bytes; p.stages = pl.stages; p.narrow_last = pl.narrow_last; p.q_bytes = pl.q_bytes;
    p.idesc_qk = at_idesc(KB, 2 * QT);
    p.idesc_lo = at_idesc(p.n_lo, 2 * QT);
    p.idesc_hi = at_idesc(p.n_hi ? p.n_hi : 32, 2 * QT);
    at_use_plan(a, p, 2 * QT);
    CUtensorMap mapQ, mapK, mapVlo, mapVhi;
    at_maps(a, QT, KB / 2, p.n_lo / 2, p.n_hi / 2, mapQ, mapK, mapVlo, mapVhi);
    CUtensorMap mapQ8 = mapQ, mapK8 = mapK;
    if (p.narrow_last) {
        const long long H2 = 2LL * a.H;
        const long long qdims[4] = {a.dk, a.L, a.n_head, a.B};
        const long long qstr[3] = {H2, a.dk, (long long)a.L * H2};
        mapQ8 = at_map(a.qk, qdims, qstr, 8, QT, CU_TENSOR_MAP_SWIZZLE_32B);
        mapK8 = at_map(a.qk + a.H, qdims, qstr, 8, KB / 2, CU_TENSOR_MAP_SWIZZLE_32B);
    }
    const int clusters = std::min(p.num_tiles, at_pair_clusters(pl.smem));
    ZVX_REQUIRE(clusters >= 1, "attn_fused: no CTA pair can be resident");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(AT_THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
    at_dbg_begin(p, st);
    ZVX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_pair_kernel, mapQ, mapK, mapVlo, mapVhi, mapQ8, mapK8, p));
    ZVX_POST_LAUNCH();
    at_dbg_end(p, st);
}

}  // namespace

size_t attn_plan_bytes(int B, int L, int n_head) {
    return sizeof(int) * (size_t)(4 + (B + 3) / 4 * 4) + sizeof(int4) * (size_t)B * n_head * cdiv(L, QT);
}

void attn_plan(const AttnFusedArgs& a, bool skip_masked_queries, int* ws, AttnPlan& plan, cudaStream_t st) {
    ZVX_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "attn_plan: workspace must be 16-byte aligned");
    PairPlan pl;
    plan.variant = a.variant;
    plan.rows_per_tile = at_use_pair(a, pl) ? 2 * QT : QT;
    plan.dev = ws;
    attn_plan_kernel<<<1, 1024, 0, st>>>(a.key_mask, a.mask_ld, a.B, a.L, a.n_head, plan.rows_per_tile, skip_masked_queries ? 1 : 0, ws);
    ZVX_POST_LAUNCH();
}

bool attn_fused_supported(const AttnFusedArgs& a) {
    PairPlan pl;
    if (a.variant == 2) return at_pair_plan(a, pl) && at_pair_clusters(pl.smem) >= 1;
    return at_single_supported(a);
}

void attn_fused(const AttnFusedArgs& a, cudaStream_t st) {
    ZVX_REQUIRE(attn_fused_supported(a), "attn_fused: unsupported shape / alignment");
    PairPlan pl;
    if (at_use_pair(a, pl)) {
        ZVX_REQUIRE(at_pair_plan(a, pl), "attn_fused: unsupported shape for the pair kernel");
        attn_pair_launch(a, pl, st);
        return;
    }
    AttnParams p{};
    at_fill_common(a, p);
    p.qtiles = cdiv(a.L, QT);
    p.num_tiles = p.qtiles * a.n_head * a.B;
    p.NV = (a.dk + 15) / 16 * 16;
    if (p.NV <= 256) { p.n_lo = p.NV; p.n_hi = 0; }
    else { p.n_hi = 128; p.n_lo = p.NV - 128; }
    p.stage_bytes = std::max(QK_STAGE_BYTES, (int)round_up((long long)p.NV * CH * 4, 1024));
    p.stages = AT_STAGES;
    p.idesc_qk = at_idesc(KB);
    p.idesc_lo = at_idesc(p.n_lo);
    p.idesc_hi = at_idesc(p.n_hi ? p.n_hi : 16);
    p.dbg_skip = env_int("ZVX_ATTN_EXP", 0);
    at_use_plan(a, p, QT);
    at_dbg_begin(p, st);
    CUtensorMap mapQ, mapK, mapVlo, mapVhi;
    at_maps(a, QT, KB, p.n_lo, p.n_hi, mapQ, mapK, mapVlo, mapVhi);

    const size_t smem = 1024 + (size_t)AT_STAGES * p.stage_bytes + 128 + 6 * QT * 4;   // barriers, max / sum exchange
    static std::once_flag once;
    std::call_once(once, [] {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    ZVX_REQUIRE(smem <= 227 * 1024, "attn_fused: shared memory");
    const int grid = std::min(p.num_tiles, at_num_sms());
    attn_fused_kernel<<<grid, AT_THREADS, smem, st>>>(mapQ, mapK, mapVlo, mapVhi, p);
    ZVX_POST_LAUNCH();
    at_dbg_end(p, st);
}

}  // namespace zvx

namespace { thread_local std::string g_attn_error; }

extern "C" {

int zvx_attention_ex(const float* qk, const float* vt, int64_t vt_pitch, const uint8_t* key_mask, int B, int L, int n_head, int d_k,
                     float temperature, float* out, int variant, int skip_masked_queries, void* workspace, int64_t workspace_bytes,
                     void* stream) {
    try {
        zvx::AttnFusedArgs a;
        a.qk = qk; a.vt = vt; a.out = out; a.key_mask = key_mask; a.mask_ld = L; a.B = B; a.L = L; a.n_head = n_head; a.dk = d_k;
        a.H = n_head * d_k; a.Lp = (int)vt_pitch; a.temperature = temperature; a.variant = variant;
        if (variant < 0 || variant > 2) throw zvx::Error("zvx_attention: variant must be 0, 1 or 2");
        if (!zvx::attn_fused_supported(a)) throw zvx::Error("zvx_attention: unsupported shape / alignment (d_k % 8, pitches % 4, 16-byte bases)");
        zvx::AttnPlan plan;
        if (workspace) {
            if (workspace_bytes < (int64_t)zvx::attn_plan_bytes(B, L, n_head)) throw zvx::Error("zvx_attention: workspace too small (zvx_attention_workspace_bytes)");
            zvx::attn_plan(a, skip_masked_queries != 0, static_cast<int*>(workspace), plan, (cudaStream_t)stream);
            a.plan = &plan;
        } else if (skip_masked_queries) {
            throw zvx::Error("zvx_attention: skip_masked_queries needs a workspace");
        }
        zvx::attn_fused(a, (cudaStream_t)stream);
        return 0;
    } catch (const std::exception& e) {
        g_attn_error = e.what();
        return 1;
    }
}

int64_t zvx_attention_workspace_bytes(int B, int L, int n_head) {
    return (B < 1 || L < 1 || n_head < 1) ? 0 : (int64_t)zvx::attn_plan_bytes(B, L, n_head);
}

int zvx_attention(const float* qk, const float* vt, int64_t vt_pitch, const uint8_t* key_mask, int B, int L, int n_head, int d_k,
                  float temperature, float* out, void* stream) {
    return zvx_attention_ex(qk, vt, vt_pitch, key_mask, B, L, n_head, d_k, temperature, out, 0, 0, nullptr, 0, stream);
}

const char* zvx_attention_last_error(void) { return g_attn_error.c_str(); }

}  // extern "C"
