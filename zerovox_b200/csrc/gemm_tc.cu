// TF32 GEMM / implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM (gemm_tc.cuh states the generalised problem):
//   warp 0      TMA producer : cp.async.bulk.tensor.4d loads of the A tile (128 positions x 32 k, one per filter tap —
//                              the tap is a coordinate shift, the zero padding is TMA out-of-bounds fill) and of the
//                              W tile (BN x 32 k) into a ring of 128B-swizzled shared-memory stages;
//   warp 1      MMA issuer   : one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = BN, K = 8) straight
//                              from shared memory into one of two TMEM accumulators (2 x 256 columns), releases stages
//                              with tcgen05.commit;
//   warps 2..9  epilogue     : tcgen05.ld the finished accumulator (lane = output position), apply bias / ReLU /
//                              folded BatchNorm / residual, store row-major or transposed, hand the TMEM buffer back —
//                              overlapping the next tile's main loop.
// Operands are fp32 in HBM, rounded to TF32 (10-bit mantissa, round-to-nearest) by the TMA unit on their way to
// shared memory; accumulation is fp32.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <cstdlib>
#include <mutex>

namespace zvx {

namespace {

constexpr int BM = 128;                  // output positions per tile (MMA M)
constexpr int BK = 32;                   // fp32 elements per 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int NUM_THREADS = 320;          // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;          // TMEM columns between the two accumulators
constexpr int MAX_STAGES = 16;          // small-N problems are bound by bytes in flight: more, smaller stages
constexpr int SPLIT_PHASE = 8;           // k-steps per accumulation phase of the 3xTF32 mode (32 hi*hi MMAs per chain)
constexpr int LO_OFFSET = 128;           // TMEM column offset of the cross-term accumulator (split mode, BN <= 128)
constexpr int SMEM_LIMIT = 227 * 1024;

struct TcParams {
    int num_tiles, tiles_n, tiles_x, tiles_y;
    int TW, TH, BN;
    int ksx, taps, kchunks, dil, pad_x, pad_y, stride, stride_y;
    int b_batched;
    int stages, stage_bytes, b_tile_bytes, b_tile_stride;
    // x-tap reuse (xr): one activation tile of 128 + (ksx-1)*dil positions per (dy, k-chunk) serves all ksx taps of the row
    int xr, xr_na, xr_a_bytes, xr_a_tx, xr_halo;
    // 2-D tap reuse (xr = 2): ONE activation tile of (TH + halo_y) x (TW + halo_x) positions per k-chunk serves every tap of the
    // filter: accumulator row r <-> tile position (r / row_w, r % row_w) with row_w = TW + halo_x, a tap (dy, dx) is the row
    // shift (dy * row_w + dx) * dil; the halo_x columns of every tile row are computed and discarded by the epilogue.
    int xr_ky, xr_kx;         // activation loads per k-chunk (filter rows; 1 in 2-D mode) and taps served by each
    int xr_wrap8;             // extra descriptor units when the tap index wraps to the next filter row (2-D mode)
    int row_w;                // accumulator rows per tile row (= TW except in 2-D tap-reuse mode)
    int mt, a_bytes;          // MMA M tiles (128 positions each) per CTA tile, bytes of the activation tile
    int nbuf;                 // TMEM accumulator buffers: 2 (epilogue of tile i overlaps the MMAs of tile i+1) or 1 (mt * BN = 512 columns:
                              // two full-width M tiles share every weight tile — half the L2 -> SM weight traffic per MMA, the bound of
                              // the long-K convs — and the epilogue, short against a K = 4752 main loop, runs between tiles)
    int spin;                 // single-thread roles spin on their barriers (small tiles)
    int ug, unit_bytes;       // k-steps grouped per pipeline stage (small-N problems: fewer barrier round trips / commits)
    uint32_t idesc;
    int Wo, Ho, N;
    float* C;
    const float* R;
    long long c_simg, c_sy, c_sx, c_sn;
    long long r_simg, r_sy, r_sx;
    const float* bias;
    const float* scale;
    const float* shift;
    int relu_first, relu_last, vec4;
    float act_slope;          // leaky-ReLU slope applied to the value stored in C (1 = identity)
    float* C2; float slope2;  // optional second output lrelu(v, slope2), same addressing as C
    int acc_mode, acc_init; float acc_scale;   // C = (acc_init ? 0 : C) + v * acc_scale
    long long* dbg;           // optional clock64 trace of CTA 0 (ZVX_GEMM_DBG): [role][event]
    int dbg_skip;             // experiment bits: 1 = issue no MMA, 2 = issue no TMA (MMAs on stale tiles), 4 = empty epilogue
};

// K-major, 128-byte-swizzled operand tile (rows of 32 fp32 = 128 B, 8-row atoms of 1024 B): matrix descriptor.
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units               bits [0,14)
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused: one atom in K)  bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: next 8-row atom          bits [32,46)
    d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)               bits [46,48)
    d |= (uint64_t)2 << 61;                     // layout type: SWIZZLE_128B                    bits [61,64)
    return d;
}

// A 128-byte-swizzled tile can be entered `rows` rows (128 B each) further down by adding rows * 8 to the descriptor's
// address field.  Measured on B200: the swizzle is a function of the absolute shared-memory address bits, so a start
// address that is not aligned to the 1024-byte atom needs NO base-offset field (setting it to (addr >> 7) & 7 produces
// wrong operands).  The x-tap-reuse path below relies on this.

// Waits of the two single-thread roles: small tiles are bound by hand-shake latency (spin on test_wait); with large tiles the
// waits are long and a spinning thread steals issue slots from the epilogue warps on its scheduler (suspending try_wait).
__device__ __forceinline__ void mbar_wait_sel(int spin, uint32_t bar, uint32_t parity) {
    if (spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity);
}

struct Tile { int n0, x0, y0, img; };
__device__ __forceinline__ Tile decode_tile(const TcParams& p, int t) {
    Tile c;
    const int nt = t % p.tiles_n; t /= p.tiles_n;
    const int xt = t % p.tiles_x; t /= p.tiles_x;
    const int yt = t % p.tiles_y;
    c.img = t / p.tiles_y;
    c.n0 = nt * p.BN; c.x0 = xt * p.TW; c.y0 = yt * p.TH;
    return c;
}

__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float epi(const TcParams& p, float x, int n) {
    if (p.bias) x += __ldg(p.bias + n);
    if (p.relu_first) x = fmaxf(x, 0.f);
    if (p.scale) x = fmaf(x, __ldg(p.scale + n), __ldg(p.shift + n));
    return x;
}

// SPLIT (3xTF32, fp32-grade products): each operand arrives as a pair (raw fp32 tile, whose low 13 mantissa bits the
// tensor core ignores = hi; and a TF32-exact tile lo = rn(x - hi) prepared by tf32_split_lo) and every k-step issues
// A_lo*B_hi + A_hi*B_lo + A_hi*B_hi into the same fp32 accumulator; the dropped lo*lo term is ~2^-22 relative.
template <bool GENERAL, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapBlo, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_a_ring = smem_u32(smem);                               // xr mode: activation ring first
    const uint32_t smem_base = smem_a_ring + (uint32_t)(p.xr_na * p.xr_a_bytes);   // operand stages
    const uint32_t bars = smem_base + (uint32_t)(p.stages * p.stage_bytes);
    // barrier i at bars + 8*i : full[stages] | empty[stages] | tmem_full[2] | tmem_empty[2] | tmem base slot
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(p.stages + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (uint32_t)(2 * p.stages + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (uint32_t)(2 * p.stages + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (uint32_t)(2 * p.stages + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem + p.xr_na * p.xr_a_bytes + p.stages * p.stage_bytes + 8 * (2 * p.stages + 4));
    auto afull_bar = [&](int s) { return bars + 8u * (uint32_t)(2 * p.stages + 5 + s); };
    auto aempty_bar = [&](int s) { return bars + 8u * (uint32_t)(2 * p.stages + 5 + p.xr_na + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        if (SPLIT || (p.xr == 1 && p.mt == 2)) prefetch_tmap(&mapAlo);
        if (SPLIT) prefetch_tmap(&mapBlo);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);
        }
        for (int s = 0; s < p.xr_na; ++s) {
            mbar_init(afull_bar(s), 1);
            mbar_init(aempty_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();   // (barriers, TMEM and descriptors are set up: from here on the predecessor's output is read)

    const int ksteps = p.taps * p.kchunks;
    // shared-memory descriptor of the first operand stage: low word (address | LBO), and the word common to all (SBO = 1024 B,
    // version 1, SWIZZLE_128B)
    const uint32_t desc_lo0 = (uint32_t)(sw128_desc(smem_base) & 0xFFFFFFFFull);
    const uint32_t DESC_HI = (uint32_t)(sw128_desc(0) >> 32);
    // SPLIT: the tensor core's fp32 accumulate truncates, a bias that grows with the accumulation chain; the chain is
    // cut every SPLIT_PHASE k-steps and the partial sums are added in registers (round-to-nearest) by the epilogue warps.
    const int phase_len = SPLIT ? SPLIT_PHASE : ksteps;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)(p.a_bytes + p.b_tile_bytes) * (SPLIT ? 2u : 1u);
            if (!SPLIT && p.xr) {
                // x-tap reuse: per (dy, k-chunk) ONE activation tile of 128 + halo positions, then the ksx weight tiles
                int sa = 0;
                uint32_t pa = 0;
                for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                    const Tile c = decode_tile(p, t);
                    if (t + (int)gridDim.x >= p.num_tiles) pdl_trigger();   // last tile of this CTA: the next launch may move in behind it
                    for (int dy = 0; dy < p.xr_ky; ++dy)
                        for (int kc = 0; kc < p.kchunks; ++kc) {
                            mbar_wait_sel(p.spin, aempty_bar(sa), pa ^ 1u);
                            mbar_arrive_expect_tx(afull_bar(sa), (uint32_t)p.xr_a_tx);
                            if (p.xr == 1 && p.mt == 2) {
                                // 256 + halo positions exceed the 256-wide TMA box: the first 128 through the 128-wide map
                                // (passed in mapAlo's slot), the other 128 + halo behind them
                                tma_load_4d(&mapAlo, afull_bar(sa), smem_a_ring + (uint32_t)(sa * p.xr_a_bytes), kc * BK,
                                            c.x0 - p.pad_x, c.y0 + dy * p.dil - p.pad_y, c.img);
                                tma_load_4d(&mapA, afull_bar(sa), smem_a_ring + (uint32_t)(sa * p.xr_a_bytes + A_TILE_BYTES), kc * BK,
                                            c.x0 - p.pad_x + BM, c.y0 + dy * p.dil - p.pad_y, c.img);
                            } else
                            tma_load_4d(&mapA, afull_bar(sa), smem_a_ring + (uint32_t)(sa * p.xr_a_bytes), kc * BK, c.x0 - p.pad_x,
                                        c.y0 + dy * p.dil - p.pad_y, c.img);
                            if (++sa == p.xr_na) { sa = 0; pa ^= 1u; }
                            for (int dx = 0; dx < p.xr_kx; dx += p.ug) {   // p.ug weight tiles (taps) per stage
                                const int ng = min(p.ug, p.xr_kx - dx);
                                mbar_wait_sel(p.spin, empty_bar(stage), phase ^ 1u);
                                const uint32_t fb = full_bar(stage);
                                mbar_arrive_expect_tx(fb, (uint32_t)(p.b_tile_bytes * ng));
                                for (int j = 0; j < ng; ++j)
                                    tma_load_4d(&mapB, fb, smem_base + (uint32_t)(stage * p.stage_bytes + j * p.b_tile_stride), kc * BK,
                                                c.n0, dy * p.xr_kx + dx + j, 0);
                                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                            }
                        }
                }
            } else
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const Tile c = decode_tile(p, t);
                if (t + (int)gridDim.x >= p.num_tiles) pdl_trigger();
                if (ZVX_DBG_PTR(p) && blockIdx.x == 0 && t / (int)gridDim.x < 16) ZVX_DBG_PTR(p)[t / gridDim.x] = clock64();
                long long pwait = 0;
                // (single-thread loop: its instruction latency bounds small tiles -> counters instead of divisions)
                const int cx = c.x0 * p.stride - p.pad_x, cy = c.y0 * p.stride_y - p.pad_y;
                const int bz1 = p.b_batched ? c.y0 : 0, bz2 = p.b_batched ? c.img : 0;
                int kc = 0, dx = 0, dy = 0, tap = 0;
                const int ug = SPLIT ? 1 : p.ug;
                for (int s = 0; s < ksteps; s += ug) {
                    const int ng = SPLIT ? 1 : min(ug, ksteps - s);   // k-steps in this stage
                    const long long w0 = ZVX_DBG_PTR(p) ? clock64() : 0;
                    mbar_wait_sel(p.spin, empty_bar(stage), phase ^ 1u);
                    if (ZVX_DBG_PTR(p)) pwait += clock64() - w0;
                    if (ZVX_DBG_SKIP(p) & 2) { mbar_arrive(full_bar(stage)); if (++stage == p.stages) { stage = 0; phase ^= 1u; } continue; }
                    const uint32_t fb = full_bar(stage);
                    mbar_arrive_expect_tx(fb, tx_bytes * (uint32_t)ng);
                    uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
                    for (int j = 0; j < ng; ++j, sa += (uint32_t)p.unit_bytes) {
                        const int xk = kc * BK, xa = cx + dx * p.dil, ya = cy + dy * p.dil;
                        const int z1 = p.b_batched ? bz1 : tap;
                        tma_load_4d(&mapA, fb, sa, xk, xa, ya, c.img);
                        tma_load_4d(&mapB, fb, sa + (uint32_t)p.a_bytes, xk, c.n0, z1, bz2);
                        if (SPLIT) {
                            const uint32_t sl = sa + (uint32_t)(A_TILE_BYTES + p.b_tile_stride);
                            tma_load_4d(&mapAlo, fb, sl, xk, xa, ya, c.img);
                            tma_load_4d(&mapBlo, fb, sl + A_TILE_BYTES, xk, c.n0, z1, bz2);
                        }
                        if (++kc == p.kchunks) {
                            kc = 0; ++tap;
                            if (++dx == p.ksx) { dx = 0; ++dy; }
                        }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                if (ZVX_DBG_PTR(p) && blockIdx.x == 0 && t / (int)gridDim.x < 16) ZVX_DBG_PTR(p)[48 + t / gridDim.x] = ZVX_DBG_PTR(p)[0] + pwait;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t cnt = 0;   // accumulation phases issued so far (one per tile unless SPLIT)
            if (!SPLIT && p.xr) {
                int sa = 0;
                uint32_t pa = 0;
                for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++cnt) {
                    const int buf = p.nbuf == 2 ? (int)(cnt & 1u) : 0;
                    mbar_wait_sel(p.spin, tempty_bar(buf), ((p.nbuf == 2 ? (cnt >> 1) : cnt) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_STRIDE);
                    uint32_t acc_tap = 0u;   // 0 for the first tap of the tile (fresh accumulators), 1 afterwards
                    // M tiles that lie entirely past the row's end are not computed (a 256-wide last tile of a row costs what a
                    // 128-wide one costs)
                    const int mt_eff = (p.xr == 1 && p.mt == 2 && decode_tile(p, t).x0 + BM >= p.Wo) ? 1 : p.mt;   // (1-D rows only)
                    for (int dy = 0; dy < p.xr_ky; ++dy)
                        for (int kc = 0; kc < p.kchunks; ++kc) {
                            mbar_wait_sel(p.spin, afull_bar(sa), pa);
                            tc_fence_after();
                            // the tap's operand = the tile entered (tap shift) rows further down: one row = 128 B = 8 descriptor units
                            uint32_t alo = (uint32_t)(sw128_desc(smem_a_ring + (uint32_t)(sa * p.xr_a_bytes)) & 0xFFFFFFFFull);
                            int tx = 0;   // tap column inside its filter row (2-D mode: the shift jumps a tile row when it wraps)
                            for (int dx = 0; dx < p.xr_kx; dx += p.ug) {
                                const int ng = min(p.ug, p.xr_kx - dx);
                                mbar_wait_sel(p.spin, full_bar(stage), phase);
                                uint32_t blo = desc_lo0 + (uint32_t)((stage * p.stage_bytes) >> 4);
                                for (int j = 0; j < ng; ++j, blo += (uint32_t)(p.b_tile_stride >> 4)) {
                                    for (int mi = 0; mi < mt_eff; ++mi) {   // M tile mi = the same tile entered 128 rows further down
                                        const uint32_t am = alo + (uint32_t)(mi * BM * 8);
#pragma unroll
                                        for (int k = 0; k < BK / 8; ++k)
                                            umma_tf32_lo(d_tmem + (uint32_t)(mi * p.BN), am + 2 * k, blo + 2 * k, DESC_HI, p.idesc,
                                                         k == 0 ? acc_tap : 1u);
                                    }
                                    acc_tap = 1u;
                                    alo += (uint32_t)(p.dil * 8);
                                    if (++tx == p.ksx) { tx = 0; alo += (uint32_t)p.xr_wrap8; }
                                }
                                umma_commit(empty_bar(stage));
                                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                            }
                            umma_commit(aempty_bar(sa));
                            if (++sa == p.xr_na) { sa = 0; pa ^= 1u; }
                        }
                    umma_commit(tfull_bar(buf));
                }
            } else
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                for (int s0 = 0; s0 < ksteps; s0 += phase_len, ++cnt) {
                    const int buf = p.nbuf == 2 ? (int)(cnt & 1u) : 0;
                    const uint32_t par = (p.nbuf == 2 ? (cnt >> 1) : cnt) & 1u;
                    mbar_wait_sel(p.spin, tempty_bar(buf), par ^ 1u);
                    tc_fence_after();
                    if (ZVX_DBG_PTR(p) && blockIdx.x == 0 && cnt < 16) ZVX_DBG_PTR(p)[16 + cnt] = clock64();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_STRIDE);
                    const int s1 = min(ksteps, s0 + phase_len);
                    long long waited = 0;
                    const int ug = SPLIT ? 1 : p.ug;
                    for (int s = s0; s < s1; s += ug) {
                        const int ng = SPLIT ? 1 : min(ug, s1 - s);
                        const long long w0 = ZVX_DBG_PTR(p) ? clock64() : 0;
                        mbar_wait_sel(p.spin, full_bar(stage), phase);
                        if (ZVX_DBG_PTR(p)) waited += clock64() - w0;
                        // no tcgen05 fence here: the operands were written by the async proxy (TMA) and the mbarrier
                        // completion orders them before the MMA's own async-proxy reads
                        uint32_t alo = desc_lo0 + (uint32_t)((stage * p.stage_bytes) >> 4);   // low descriptor words
                        for (int j = 0; j < ng; ++j, alo += (uint32_t)(p.unit_bytes >> 4)) {
                            const uint32_t blo = alo + (uint32_t)(p.a_bytes >> 4);
                            const uint32_t acc0 = (s + j == s0) ? 0u : 1u;
                            if (SPLIT) {
                                const uint32_t la = alo + (uint32_t)((A_TILE_BYTES + p.b_tile_stride) >> 4);
                                const uint32_t lb = la + (uint32_t)(A_TILE_BYTES >> 4);
#pragma unroll
                                for (int k = 0; k < BK / 8; ++k) {
                                    const uint32_t accum = k == 0 ? acc0 : 1u;
                                    umma_tf32_lo(d_tmem + LO_OFFSET, la + 2 * k, blo + 2 * k, DESC_HI, p.idesc, accum);
                                    umma_tf32_lo(d_tmem + LO_OFFSET, alo + 2 * k, lb + 2 * k, DESC_HI, p.idesc, 1u);
                                    umma_tf32_lo(d_tmem, alo + 2 * k, blo + 2 * k, DESC_HI, p.idesc, accum);
                                }
                            } else if (!(ZVX_DBG_SKIP(p) & 1)) {
                                for (int mi = 0; mi < p.mt; ++mi) {   // the M tiles of the CTA tile share the weight tile
                                    const uint32_t am = alo + (uint32_t)(mi * (A_TILE_BYTES >> 4));
#pragma unroll
                                    for (int k = 0; k < BK / 8; ++k)   // 8 tf32 = 32 bytes = 2 descriptor units per MMA
                                        umma_tf32_lo(d_tmem + (uint32_t)(mi * p.BN), am + 2 * k, blo + 2 * k, DESC_HI, p.idesc,
                                                     k == 0 ? acc0 : 1u);
                                }
                            }
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(tfull_bar(buf));
                    if (ZVX_DBG_PTR(p) && blockIdx.x == 0 && cnt < 16) { ZVX_DBG_PTR(p)[32 + cnt] = clock64(); ZVX_DBG_PTR(p)[64 + cnt] = ZVX_DBG_PTR(p)[0] + waited; }
                }
            }
        }
        __syncwarp();
    } else {
        // ================================================================ epilogue (warps 2..9)
        // Two warps per TMEM lane quarter; they take alternate 16-column chunks of the accumulator.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        uint32_t cnt = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const Tile c = decode_tile(p, t);
            // position of this thread's accumulator row inside M tile mi of the CTA tile
            int r = q * 32 + lane;
            int ty = r / p.row_w, tx = r - ty * p.row_w;
            int y = c.y0 + ty, x = c.x0 + tx;
            bool valid = (y < p.Ho) && (x < p.Wo) && (tx < p.TW);
            long long off = (long long)c.img * p.c_simg + (long long)y * p.c_sy + (long long)x * p.c_sx;
            float* __restrict__ cp = p.C + off;
            const float* __restrict__ rp =
                p.R ? p.R + ((long long)c.img * p.r_simg + (long long)y * p.r_sy + (long long)x * p.r_sx) : nullptr;
            if (SPLIT) {
                // ---- 3xTF32: sum the per-phase partial accumulators (hi*hi and the cross terms) in registers
                float acc[4][16];
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[ci][e] = 0.f;
                for (int s0 = 0; s0 < ksteps; s0 += phase_len, ++cnt) {
                    const int buf = p.nbuf == 2 ? (int)(cnt & 1u) : 0;
                    mbar_wait(tfull_bar(buf), (p.nbuf == 2 ? (cnt >> 1) : cnt) & 1u);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC_STRIDE);
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int c0 = half * 16 + ci * 32;
                        if (c0 < p.BN) {
                            uint32_t v[16], w[16];
                            tmem_ld16(trow + (uint32_t)c0, v);
                            tmem_ld16(trow + (uint32_t)(LO_OFFSET + c0), w);
                            tmem_wait_ld();
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[ci][e] += __uint_as_float(v[e]) + __uint_as_float(w[e]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(buf));
                }
                if (valid) {
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int c0 = half * 16 + ci * 32;
                        const int n = c.n0 + c0;
                        if (c0 >= p.BN || n >= p.N) continue;
                        float o[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            o[e] = 0.f;
                            if (n + e < p.N) {
                                float xv = epi(p, acc[ci][e], n + e);
                                if (rp) xv += rp[(long long)(n + e) * p.c_sn];
                                if (p.relu_last) xv = fmaxf(xv, 0.f);
                                o[e] = xv;
                            }
                        }
                        if (p.vec4 && n + 16 <= p.N) {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                *(reinterpret_cast<float4*>(cp + n) + g) =
                                    make_float4(o[4 * g + 0], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (n + e < p.N && c0 + e < p.BN) cp[(long long)(n + e) * p.c_sn] = o[e];
                        }
                    }
                }
                continue;
            }
            const int buf = p.nbuf == 2 ? (int)(cnt & 1u) : 0;
            const uint32_t par = (p.nbuf == 2 ? (cnt >> 1) : cnt) & 1u;
            ++cnt;
            mbar_wait(tfull_bar(buf), par);
            tc_fence_after();
            const int mt_epi = (p.xr == 1 && p.mt == 2 && c.x0 + BM >= p.Wo) ? 1 : p.mt;   // (M tile never computed: nothing to store)
            for (int mi = 0; mi < mt_epi; ++mi) {
            if (mi > 0) {
                r = mi * BM + q * 32 + lane;
                ty = r / p.row_w; tx = r - ty * p.row_w;
                y = c.y0 + ty; x = c.x0 + tx;
                valid = (y < p.Ho) && (x < p.Wo) && (tx < p.TW);
                off = (long long)c.img * p.c_simg + (long long)y * p.c_sy + (long long)x * p.c_sx;
                cp = p.C + off;
                rp = p.R ? p.R + ((long long)c.img * p.r_simg + (long long)y * p.r_sy + (long long)x * p.r_sx) : nullptr;
            }
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC_STRIDE + mi * p.BN);
            for (int c0 = half * 16; c0 < ((ZVX_DBG_SKIP(p) & 4) ? 0 : p.BN); c0 += 32) {
                uint32_t v[16];
                __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
                tmem_ld16(trow + (uint32_t)c0, v);
                const int n = c.n0 + c0;
                if (p.vec4 && n + 16 <= p.N) {
                    // ---- vector path: every global access of the chunk is issued before the first use
                    float4 bb[4], rr[4], oo[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        bb[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rr[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                        oo[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (p.bias) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) bb[g] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + g);
                    }
                    if (valid && rp) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) rr[g] = *(reinterpret_cast<const float4*>(rp + n) + g);
                    }
                    if (GENERAL && valid && p.acc_mode && !p.acc_init) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) oo[g] = *(reinterpret_cast<const float4*>(cp + n) + g);
                    }
                    tmem_wait_ld();
                    float o[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        o[4 * g + 0] = __uint_as_float(v[4 * g + 0]) + bb[g].x;
                        o[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + bb[g].y;
                        o[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + bb[g].z;
                        o[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + bb[g].w;
                    }
                    if (p.relu_first) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
                    }
                    if (GENERAL && p.scale) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n) + g);
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + n) + g);
                            o[4 * g + 0] = fmaf(o[4 * g + 0], sc.x, sh.x); o[4 * g + 1] = fmaf(o[4 * g + 1], sc.y, sh.y);
                            o[4 * g + 2] = fmaf(o[4 * g + 2], sc.z, sh.z); o[4 * g + 3] = fmaf(o[4 * g + 3], sc.w, sh.w);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        o[4 * g + 0] += rr[g].x; o[4 * g + 1] += rr[g].y; o[4 * g + 2] += rr[g].z; o[4 * g + 3] += rr[g].w;
                    }
                    if (p.relu_last) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
                    }
                    if (GENERAL && p.acc_mode) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            o[4 * g + 0] = fmaf(o[4 * g + 0], p.acc_scale, oo[g].x);
                            o[4 * g + 1] = fmaf(o[4 * g + 1], p.acc_scale, oo[g].y);
                            o[4 * g + 2] = fmaf(o[4 * g + 2], p.acc_scale, oo[g].z);
                            o[4 * g + 3] = fmaf(o[4 * g + 3], p.acc_scale, oo[g].w);
                        }
                    }
                    if (valid) {
                        if (GENERAL) {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                *(reinterpret_cast<float4*>(cp + n) + g) =
                                    make_float4(lrelu_f(o[4 * g + 0], p.act_slope), lrelu_f(o[4 * g + 1], p.act_slope),
                                                lrelu_f(o[4 * g + 2], p.act_slope), lrelu_f(o[4 * g + 3], p.act_slope));
                            if (p.C2) {
#pragma unroll
                                for (int g = 0; g < 4; ++g)
                                    *(reinterpret_cast<float4*>(p.C2 + off + n) + g) =
                                        make_float4(lrelu_f(o[4 * g + 0], p.slope2), lrelu_f(o[4 * g + 1], p.slope2),
                                                    lrelu_f(o[4 * g + 2], p.slope2), lrelu_f(o[4 * g + 3], p.slope2));
                            }
                        } else {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                *(reinterpret_cast<float4*>(cp + n) + g) =
                                    make_float4(o[4 * g + 0], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                        }
                    }
                } else {
                    // ---- element path: column tails, unaligned or transposed (c_sn != 1) outputs
                    tmem_wait_ld();
                    if (!valid) continue;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        if (c0 + e < p.BN && n + e < p.N) {
                            float xv = epi(p, __uint_as_float(v[e]), n + e);
                            const long long a = (long long)(n + e) * p.c_sn;
                            if (rp) xv += rp[a];
                            if (p.relu_last) xv = fmaxf(xv, 0.f);
                            if (GENERAL) {
                                if (p.acc_mode) xv = fmaf(xv, p.acc_scale, p.acc_init ? 0.f : cp[a]);
                                cp[a] = lrelu_f(xv, p.act_slope);
                                if (p.C2) p.C2[off + a] = lrelu_f(xv, p.slope2);
                            } else {
                                cp[a] = xv;
                            }
                        }
                    }
                }
            }
            }   // mi
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    if (!fn) throw Error("cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

// 4-D fp32 tensor map, dim 0 contiguous, 128-byte swizzle, zero fill out of bounds.
// estride: traversal step of dims 1 / 2 (strided convolutions): the box spans box[i]*estride elements, every
// estride-th one is fetched, so the shared-memory tile keeps box[i] rows.
CUtensorMap make_map(const float* base, const long long dims[4], const long long strides_elems[3], const int box[4],
                     bool raw_f32 = false, int estride_x = 1, int estride_y = 1) {
    CUtensorMap m;
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4] = {1, (cuuint32_t)estride_x, (cuuint32_t)estride_y, 1};
    long long prev = 16;
    for (int i = 0; i < 4; ++i) {
        gd[i] = (cuuint64_t)std::max<long long>(dims[i], 1);
        bx[i] = (cuuint32_t)box[i] * es[i];
    }
    for (int i = 0; i < 3; ++i) {
        long long s = strides_elems[i] * 4;
        if (dims[i + 1] <= 1 && s <= 0) s = prev;   // unit dims: any valid stride
        gs[i] = (cuuint64_t)s;
        prev = std::max<long long>(s, 16);
    }
    // TFLOAT32 element type: the TMA unit rounds fp32 -> tf32 (nearest) in flight.  With plain FLOAT32 the tensor core
    // truncates the low 13 mantissa bits, a systematic -7e-4 relative bias per contraction (measured, tools/diag_tf32.py;
    // ZVX_TMAP_F32=1 restores that behaviour for the experiment).
    static const bool f32_env = env_set("ZVX_TMAP_F32");
    const bool f32_type = f32_env || raw_f32;   // split mode: hi = the tensor core's own truncation of the raw bits
    const CUresult rc = encode_fn()(&m, f32_type ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<float*>(base), gd, gs, bx, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof(buf),
                 "cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld %lld strides(B) %lld %lld %lld box %d %d %d %d",
                 (int)rc, dims[0], dims[1], dims[2], dims[3], (long long)gs[0], (long long)gs[1], (long long)gs[2], box[0],
                 box[1], box[2], box[3]);
        throw Error(buf);
    }
    return m;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool mult4(long long v) { return (v & 3) == 0; }

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        ZVX_CUDA_CHECK(cudaGetDevice(&dev));
        ZVX_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

}  // namespace

namespace {
// lo = rn_tf32(x - trunc_tf32(x)): the part of x the tensor core drops when it reads raw fp32 bits as TF32
__global__ void tf32_split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, long long n4) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    auto f = [](float a) { return rn_tf32(a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u)); };
    lo[i] = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
}
}  // namespace

void tf32_split_lo(const float* x, float* lo, long long n, cudaStream_t st) {
    if (n <= 0) return;
    ZVX_REQUIRE(aligned16(x) && aligned16(lo), "tf32_split_lo: unaligned");
    const long long n4 = (n + 3) / 4;   // buffers are padded to 16 bytes by the workspace / weight uploader
    launch_k(tf32_split_lo_kernel, dim3(cdiv(n4, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(lo), n4);
    ZVX_POST_LAUNCH();
}

bool gemm_tc_supported(const TcGemmArgs& a) {
    if (!a.A || !a.W || !a.C) return false;
    if (a.K < 8 || a.N < 8 || a.Wo < 1 || a.Ho < 1 || a.IMG < 1) return false;
    if (!aligned16(a.A) || !aligned16(a.W)) return false;
    if (!mult4(a.a_sx) || !mult4(a.a_sy) || !mult4(a.a_simg) || !mult4(a.w_sn) || !mult4(a.w_s1) || !mult4(a.w_s2))
        return false;
    if (a.a_sx <= 0 || a.w_sn <= 0) return false;
    return true;
}

void gemm_tc(const TcGemmArgs& a, cudaStream_t st) {
    ZVX_REQUIRE(gemm_tc_supported(a), "gemm_tc: operand layout not supported by the TMA path");
    const bool split = a.A_lo != nullptr || a.W_lo != nullptr;
    ZVX_REQUIRE(!split || (a.A_lo && a.W_lo && aligned16(a.A_lo) && aligned16(a.W_lo)), "gemm_tc: split mode needs both lo operands");
    ZVX_REQUIRE(!split || (!a.acc_mode && a.act_slope == 1.f && !a.C2), "gemm_tc: split mode has the plain epilogue only");
    TcParams p{};
    // column tile: multiple of 16, <= 256, least padding over a few tile counts
    {
        const int bn_max = split ? 128 : 256;   // split stages hold four operand tiles
        const int t0 = cdiv(a.N, bn_max);
        long long bw = -1;
        for (int tn = t0; tn <= t0 + 4; ++tn) {
            const int bn = (int)std::min<long long>(bn_max, round_up(cdiv(a.N, tn), 16));
            const int tiles = cdiv(a.N, bn);
            const long long waste = (long long)tiles * bn - a.N;
            if (bw < 0 || waste < bw) { bw = waste; p.BN = bn; p.tiles_n = tiles; }
        }
    }
    // x-tap reuse: the ksx taps of a filter row read one activation tile of 128 + (ksx-1)*dil positions through
    // row-shifted descriptors instead of ksx separate TMA loads (the activation re-fetch is what bounds small-N convs)
    static const bool no_xr = env_set("ZVX_NO_XR");
    const int halo = (a.ksx - 1) * a.dil;
    // (measured: pays off for filter rows of >= 5 taps — FFN k = 9, HiFi-GAN k = 7 / 11; for 3-tap rows the forced
    // 128 x 1 tile shape costs more in padding than the saved fetches)
    static const int xr_mink = env_int("ZVX_XR_MINK", 5);
    // narrow 1-D convs are hand-shake bound: there even a 3-tap row gains from one activation load per k-chunk (A/B: ZVX_XR1_K3)
    // (measured: vocoder stage 8.21 -> 8.18 ms, profiles/r01_ab_conv1d_tap_reuse_two_m_tiles_and_k3.jsonl)
    static const int xr1_k3 = env_int("ZVX_XR1_K3", 1);
    const bool xr_narrow = xr1_k3 && a.ksy == 1 && a.ksx >= 3 && a.N <= 128;
    const bool xr = !no_xr && !split && (a.ksx >= xr_mink || xr_narrow) && a.stride == 1 && !a.b_batched && halo <= 120;
    // 2-D tap reuse (ksy > 1, unit stride): one (TH + halo_y) x (TW + halo_x) activation tile per k-chunk serves all taps; the
    // activation fetch drops from taps x 16 KB to ~25 KB per k-chunk at the price of halo_x discarded columns per tile row.
    // Measured on configs[1] (profiles/r01_ab_conv2d_tap_reuse.jsonl): speaker net 4.21 -> 3.81 ms with N <= 128 (32 / 64 / 128
    // channel 3x3 convs); ZVX_XR2=0 switches it off, =2 forces it on problems below the size threshold (tests).
    static const int xr2_env = env_int("ZVX_XR2", 1);
    static const int xr2_maxn = env_int("ZVX_XR2_MAXN", 128);
    const int halo_y = (a.ksy - 1) * a.dil;
    bool xr2 = xr2_env && !split && !xr && a.ksy > 1 && a.ksx > 1 && a.stride == 1 && !a.b_batched && a.N <= xr2_maxn;
    int xr2_tw = 0, xr2_th = 0, xr2_roww = 0, xr2_mt = 1, xr2_na = 3;
    if (xr2) {
        // tile = mt * 128 accumulator rows = (mt * th) tile rows of roww positions, (roww - halo_x) of them kept.  Fewest tiles
        // first (ties: the narrower tile row, smaller halo buffer); two M tiles per CTA tile (they share the weight stage, the
        // activation load and every hand-shake) unless that pads the map by more than 4 %.  ZVX_XR2_MT=1 keeps one M tile.
        // Measured (profiles/r01_ab_conv2d_tap_reuse_two_m_tiles.jsonl): speaker net 4.32 -> 3.79 ms with N <= 128.
        static const int xr2_mtmax = env_int("ZVX_XR2_MT", 2);
        const int wstride = (int)round_up(p.BN * BK * 4, 1024);
        const int wstage = std::max(1, std::min(a.ksx * a.ksy, (48 * 1024) / wstride)) * wstride;
        long long best_mt[3] = {-1, -1, -1};
        int sel[3][4] = {};
        static const int xr2_mt_maxn = env_int("ZVX_XR2_MT_MAXN", 128);
        for (int mt = 1; mt <= std::min(2, xr2_mtmax); ++mt) {
            if (mt * p.BN > ACC_STRIDE || (mt > 1 && a.N > xr2_mt_maxn)) continue;
            for (int roww = 8; roww <= 128; roww <<= 1) {
                const int tw = roww - halo, th = BM / roww;
                if (tw < 1 || mt * th + halo_y > 256) continue;
                const long long abytes = round_up((long long)(mt * BM + halo_y * roww + halo) * BK * 4, 1024);
                int na = 3;
                if ((SMEM_LIMIT - 2048 - na * abytes) / wstage < 2) na = 2;
                if ((SMEM_LIMIT - 2048 - na * abytes) / wstage < 2) continue;
                const long long tiles = (long long)cdiv(a.Wo, tw) * cdiv(a.Ho, mt * th) * mt;   // in 128-row units
                if (best_mt[mt] < 0 || tiles < best_mt[mt]) {
                    best_mt[mt] = tiles; sel[mt][0] = tw; sel[mt][1] = th; sel[mt][2] = roww; sel[mt][3] = na;
                }
            }
        }
        int mt = 1;
        if (best_mt[2] >= 0 && (best_mt[1] < 0 || best_mt[2] * 100 <= best_mt[1] * 104) &&
            best_mt[2] * a.IMG * p.tiles_n >= 4LL * num_sms())
            mt = 2;
        const long long best2 = best_mt[mt];
        xr2_tw = sel[mt][0]; xr2_th = sel[mt][1]; xr2_roww = sel[mt][2]; xr2_na = sel[mt][3]; xr2_mt = mt;
        // small problems keep the padding-free tiles of the general path
        if (best2 < 0 || (xr2_env < 2 && best2 * a.IMG * cdiv(a.N, 256) < 2LL * num_sms())) xr2 = false;   // ZVX_XR2=2: always (tests)
    }
    // tile shape: TH x TW = 128 * mt positions, least padding first, wider rows on ties.  Two M tiles per CTA tile for narrow
    // outputs (N <= 64): they share every weight tile and, above all, every stage hand-shake and per-tile overhead.
    static const bool no_mt2 = env_set("ZVX_NO_MT2");
    const long long positions = (long long)a.IMG * a.Ho * a.Wo;
    static const int mt2_n = env_int("ZVX_MT2_N", 64);
    p.mt = (!no_mt2 && !split && !xr && !xr2 && !a.b_batched && a.N <= mt2_n && positions >= 2LL * 256 * num_sms()) ? 2 : 1;
    long long best = -1;
    if (xr) {
        // filter-row tap reuse: 128 x 1 tiles, or 256 x 1 (two M tiles behind one haloed activation buffer and one weight stage)
        // for narrow outputs when the longer tiles pad the rows by < 4 % and still fill the GPU.  ZVX_XR1_MT=1: one M tile.
        // (measured: vocoder stage 8.30 -> 8.19 ms)
        static const int xr1_mtmax = env_int("ZVX_XR1_MT", 2);
        const long long t1 = cdiv(a.Wo, BM), t2 = 2LL * cdiv(a.Wo, 2 * BM);
        if (xr1_mtmax >= 2 && 2 * p.BN <= ACC_STRIDE && a.N <= 128 && t2 * 100 <= t1 * 104 &&
            t2 * a.Ho * a.IMG * p.tiles_n >= 4LL * num_sms())
            p.mt = 2;
        // long-K, full-width convs (decoder FFN k = 9: K = 9 * 528, N = 1024) are bound by the L2 -> SM fetch of the weight tiles
        // (32 KB per tap per k-chunk): two M tiles behind one weight stage halve it.  The two 256-column accumulators take the
        // whole TMEM (single-buffered).  Row tails cost nothing extra: an M tile entirely past the row's end is skipped.
        static const int xr1_wide = env_int("ZVX_XR1_WIDE", 1);
        if (xr1_wide && p.mt == 1 && 2 * p.BN <= TMEM_COLS && (long long)a.K * a.ksx >= 2048 &&
            (long long)cdiv(a.Wo, 2 * BM) * a.Ho * a.IMG * p.tiles_n >= 2LL * num_sms())
            p.mt = 2;
        p.TW = BM * p.mt; p.TH = 1; best = 0;
    }
    if (xr2) { p.TW = xr2_tw; p.TH = xr2_th * xr2_mt; p.mt = xr2_mt; best = 0; }
    for (int pass = 0; pass < 2 && best < 0; ++pass) {
        const int bm = BM * p.mt;
        for (int tw = std::min(bm, 256); tw >= ((a.b_batched || xr) ? 128 : 8); tw >>= 1) {   // per-(y,img) W operands / xr: one y per tile
            const int th = bm / tw;
            if (th * a.stride > 256 || tw * a.stride > 256) continue;
            const long long padded = round_up(a.Wo, tw) * round_up(a.Ho, th);
            if (best < 0 || padded < best) { best = padded; p.TW = tw; p.TH = th; }
        }
        // two M tiles must not cost more than a few percent of extra padding
        if (p.mt == 2) {
            long long best1 = -1;
            for (int tw = 128; tw >= 8; tw >>= 1) {
                const long long padded = round_up(a.Wo, tw) * round_up(a.Ho, BM / tw);
                if (best1 < 0 || padded < best1) best1 = padded;
            }
            if (best < 0 || best * 100 > best1 * 104) { p.mt = 1; best = -1; }
        }
    }
    p.tiles_x = cdiv(a.Wo, p.TW);
    p.tiles_y = cdiv(a.Ho, p.TH);
    const long long nt = (long long)a.IMG * p.tiles_y * p.tiles_x * p.tiles_n;
    ZVX_REQUIRE(nt < (1LL << 30), "gemm_tc: too many tiles");
    p.num_tiles = (int)nt;
    p.ksx = a.ksx; p.taps = a.ksx * a.ksy; p.kchunks = cdiv(a.K, BK); p.dil = a.dil; p.pad_x = a.pad_x; p.pad_y = a.pad_y;
    p.b_batched = a.b_batched;
    p.b_tile_bytes = p.BN * BK * 4;
    p.b_tile_stride = (int)round_up(p.b_tile_bytes, 1024);
    p.a_bytes = p.mt * A_TILE_BYTES;
    p.stage_bytes = (p.a_bytes + p.b_tile_stride) * (split ? 2 : 1);
    p.unit_bytes = p.stage_bytes;
    p.ug = 1;
    p.spin = (p.BN <= 64) ? 1 : 0;
    {   // small stages: group k-steps so that one barrier round trip / commit covers ~40 KB of operands
        static const int ug_env = env_int("ZVX_GEMM_UG", 0);
        const int ksteps = a.ksx * a.ksy * cdiv(a.K, BK);
        // measured per configs[1] step: 1 k-step per stage 22.5 ms, 2: 21.3, 3: 20.9, 4: 21.0 (>= 3 stages must remain); an
        // ablated run (no TMA, no MMA, empty epilogue) still costs ~650 cycles per stage hand-shake, tcgen05.commit included
        int ug = ug_env > 0 ? ug_env : 3;
        ug = std::min(ug, std::min(4, ksteps));
        while (ug > 1 && (SMEM_LIMIT - 2048) / (ug * p.unit_bytes) < 3) --ug;
        if (!split && !xr && !xr2 && ug > 1) { p.ug = ug; p.stage_bytes = ug * p.unit_bytes; }
    }
    p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - 2048) / p.stage_bytes);
    p.row_w = p.TW;
    p.nbuf = (p.mt * p.BN > ACC_STRIDE) ? 1 : 2;
    if (xr || xr2) {
        // full-width two-M-tile rows (FFN k = 9): an activation buffer lasts ksx x 1024 tensor cycles, two of them are plenty, and
        // the third one's 34 KB buy a fourth 32 KB weight stage (A/B on one box: decoder 5.50 -> 5.42 ms)
        p.xr = xr ? 1 : 2; p.xr_na = xr ? ((p.mt == 2 && p.BN > 128) ? 2 : 3) : xr2_na; p.xr_halo = halo;
        if (xr) {
            p.xr_a_tx = (p.mt * BM + halo) * BK * 4;
            p.xr_a_bytes = (int)round_up(p.xr_a_tx, 1024);
            p.xr_ky = a.ksy; p.xr_kx = a.ksx; p.xr_wrap8 = 0;
        } else {
            p.row_w = xr2_roww;
            p.xr_a_tx = xr2_roww * (p.TH + halo_y) * BK * 4;                     // what the TMA box delivers
            // + halo_x rows that only the discarded accumulator rows of the last taps read (never written: any bits will do)
            p.xr_a_bytes = (int)round_up((p.mt * BM + halo_y * xr2_roww + halo) * BK * 4, 1024);
            p.xr_ky = 1; p.xr_kx = a.ksx * a.ksy; p.xr_wrap8 = (xr2_roww - a.ksx) * a.dil * 8;
        }
        // the operand stages hold weight tiles only, several taps per stage when they are small (stage hand-shakes are costly)
        p.ug = std::max(1, std::min(p.xr_kx, (48 * 1024) / p.b_tile_stride));
        p.stage_bytes = p.ug * p.b_tile_stride;
        p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - 2048 - p.xr_na * p.xr_a_bytes) / p.stage_bytes);
        if (p.stages < 2 && p.xr_na > 2) {
            p.xr_na = 2;
            p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - 2048 - p.xr_na * p.xr_a_bytes) / p.stage_bytes);
        }
    }
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3, M >> 4
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    p.Wo = a.Wo; p.Ho = a.Ho; p.N = a.N;
    p.C = a.C; p.R = a.R; p.c_simg = a.c_simg; p.c_sy = a.c_sy; p.c_sx = a.c_sx; p.c_sn = a.c_sn;
    const bool rs = a.r_sx != 0;   // residual strides default to the output's
    p.r_simg = rs ? a.r_simg : a.c_simg; p.r_sy = rs ? a.r_sy : a.c_sy; p.r_sx = rs ? a.r_sx : a.c_sx;
    p.bias = a.bias; p.scale = a.scale; p.shift = a.shift; p.relu_first = a.relu_first; p.relu_last = a.relu_last;
    p.act_slope = a.act_slope; p.C2 = a.C2; p.slope2 = a.slope2;
    p.acc_mode = a.acc_mode; p.acc_init = a.acc_init; p.acc_scale = a.acc_scale;
    ZVX_REQUIRE(!a.scale || a.shift, "gemm_tc: scale needs shift");
    p.vec4 = (a.c_sn == 1) && mult4(a.c_simg) && mult4(a.c_sy) && mult4(a.c_sx) && aligned16(a.C) &&
             (!a.R || (aligned16(a.R) && mult4(p.r_simg) && mult4(p.r_sy) && mult4(p.r_sx))) && (!a.C2 || aligned16(a.C2)) &&
             (!a.bias || aligned16(a.bias)) && (!a.scale || (aligned16(a.scale) && aligned16(a.shift)));

    const long long adims[4] = {a.K, a.Wi, a.Hi, a.IMG};
    const long long astr[3] = {a.a_sx, a.a_sy, a.a_simg};
    const int abox[4] = {BK, xr ? BM + halo : (xr2 ? xr2_roww : p.TW), xr2 ? p.TH + halo_y : p.TH, 1};
    const long long wdims[4] = {a.K, a.N, a.Z1, a.Z2};
    const long long wstr[3] = {a.w_sn, a.w_s1, a.w_s2};
    const int wbox[4] = {BK, p.BN, 1, 1};
    ZVX_REQUIRE(a.stride >= 1 && a.stride <= 2 && p.TW * a.stride <= 256 && p.TH * a.stride <= 256, "gemm_tc: unsupported stride");
    const int sy = (a.ksy > 1 || a.Hi != a.Ho) ? a.stride : 1;   // Conv1d rows (y = utterance) are never strided
    p.stride = a.stride; p.stride_y = sy;
    const CUtensorMap mapA = make_map(a.A, adims, astr, abox, split, a.stride, sy);
    const CUtensorMap mapB = make_map(a.W, wdims, wstr, wbox, split);
    const int abox128[4] = {BK, BM, 1, 1};   // xr + two M tiles: the first 128 positions of the activation buffer
    const CUtensorMap mapAlo = split ? make_map(a.A_lo, adims, astr, abox, false, a.stride, sy)
                               : ((xr && p.mt == 2) ? make_map(a.A, adims, astr, abox128, false, a.stride, sy) : mapA);
    const CUtensorMap mapBlo = split ? make_map(a.W_lo, wdims, wstr, wbox) : mapB;

    int smem = p.xr_na * p.xr_a_bytes + p.stages * p.stage_bytes + 8 * (2 * p.stages + 5 + 2 * p.xr_na) + 16 + 1024;
    smem = std::max(smem, 117 * 1024);   // one CTA per SM: every CTA allocates all 512 TMEM columns
    static bool attr = false;
    if (!attr) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr = true;
    }
    // > 116 KB keeps the kernel at one CTA per SM (each CTA allocates all 512 TMEM columns)
    ZVX_REQUIRE(p.stages >= 2 && smem <= SMEM_LIMIT && smem > 116 * 1024, "gemm_tc: shared-memory plan out of range");
#ifdef ZVX_DEBUG
    static const char* dbg_env = getenv("ZVX_GEMM_DBG");   // "N": trace launches whose N equals that value
    static long long* dbg_buf = nullptr;
    const bool dbg = dbg_env && atoi(dbg_env) == a.N && !split;
    if (dbg) {
        if (!dbg_buf) ZVX_CUDA_CHECK(cudaMalloc(&dbg_buf, 80 * sizeof(long long)));
        ZVX_CUDA_CHECK(cudaMemsetAsync(dbg_buf, 0, 80 * sizeof(long long), st));
        p.dbg = dbg_buf;
        p.dbg_skip = getenv("ZVX_GEMM_SKIP") ? atoi(getenv("ZVX_GEMM_SKIP")) : 0;
    }
#endif
    const int grid = std::min(p.num_tiles, num_sms());
    const bool general = a.scale || a.acc_mode || a.act_slope != 1.f || a.C2;
    if (split) launch_k(gemm_tc_kernel<true, true>, dim3(grid), dim3(NUM_THREADS), smem, st, mapA, mapB, mapAlo, mapBlo, p);
    else if (general) launch_k(gemm_tc_kernel<true, false>, dim3(grid), dim3(NUM_THREADS), smem, st, mapA, mapB, mapAlo, mapBlo, p);
    else launch_k(gemm_tc_kernel<false, false>, dim3(grid), dim3(NUM_THREADS), smem, st, mapA, mapB, mapAlo, mapBlo, p);
    ZVX_POST_LAUNCH();
#ifdef ZVX_DEBUG
    if (dbg) {
        long long h[80];
        ZVX_CUDA_CHECK(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
        ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
        fprintf(stderr, "[gemm dbg] N=%d K=%d taps=%d BN=%d tiles=%d stages=%d xr=%d\n", a.N, a.K, p.taps, p.BN, p.num_tiles, p.stages, p.xr);
        const char* names[5] = {"producer tile start", "mma tile start", "mma tile committed", "producer wait empty", "mma wait on TMA data"};
        for (int r = 0; r < 5; ++r) {
            fprintf(stderr, "  %-20s:", names[r]);
            for (int i = 0; i < 10 && h[16 * r + i]; ++i) fprintf(stderr, " %lld", h[16 * r + i] - h[0]);
            fprintf(stderr, "\n");
        }
    }
#endif
}

// Smallest row count that goes to the tensor-core kernel.  A single short utterance has M = T phonemes: on the fp32 FMA
// kernel its few CTAs take 50 - 200 us per contraction, a (mostly empty) 128-row tensor-core tile a fraction of that.
// Measured with the demo flow (46 phonemes, profiles/r01_demo_host_profile.log): threshold 64 -> 16 takes zvx_encode from
// 3.56 to 1.48 ms and the whole sentence from 5.48 to 3.35 ms; all GPU parity tests pass with either value.
int tc_min_rows() {
    static const int v = env_int("ZVX_TC_MIN_M", 16);
    return v;
}

bool gemm_tc_from(const GemmArgs& g, TcGemmArgs* o) {
    if (g.b_kn || g.nz != 1) return false;
    TcGemmArgs a;
    a.A = g.A; a.K = g.K; a.W = g.W; a.N = g.N; a.w_sn = g.ldw; a.Z1 = g.taps; a.w_s1 = g.w_tap_stride; a.Z2 = 1;
    a.C = g.C; a.R = g.R; a.bias = g.bias; a.scale = g.scale; a.shift = g.shift;
    a.relu_first = g.relu_first; a.relu_last = g.relu_last; a.c_sn = 1;
    if (g.post_scale != 1.f) { a.acc_mode = 1; a.acc_init = 1; a.acc_scale = g.post_scale; }
    if (g.R) a.r_sx = g.ldr;
    if (g.mode == ROW_PLAIN) {
        if (g.taps != 1) return false;
        a.Wi = a.Wo = g.M; a.a_sx = g.lda; a.c_sx = g.ldc;
    } else if (g.mode == ROW_CONV1D) {
        if (g.Lin != g.Lout || g.M % g.Lout != 0 || g.stride != 1) return false;
        a.Wi = a.Wo = g.Lout; a.Hi = a.Ho = g.M / g.Lout;
        a.a_sx = g.lda; a.a_sy = (long long)g.Lin * g.lda;
        a.c_sx = g.ldc; a.c_sy = (long long)g.Lout * g.ldc; a.r_sy = (long long)g.Lout * g.ldr;
        a.ksx = g.taps; a.ksy = 1; a.dil = g.dil; a.pad_x = g.pad; a.pad_y = 0;
    } else {
        if (g.taps != g.ksize * g.ksize || g.stride < 1 || g.stride > 2) return false;
        if (g.Ho != (g.Hi + 2 * g.pad - g.ksize) / g.stride + 1 || g.Wo != (g.Wi + 2 * g.pad - g.ksize) / g.stride + 1) return false;
        a.Wi = g.Wi; a.Wo = g.Wo; a.Hi = g.Hi; a.Ho = g.Ho; a.IMG = g.M / (g.Ho * g.Wo); a.stride = g.stride;
        a.a_sx = g.lda; a.a_sy = (long long)g.Wi * g.lda; a.a_simg = (long long)g.Hi * g.Wi * g.lda;
        a.c_sx = g.ldc; a.c_sy = (long long)g.Wo * g.ldc; a.c_simg = (long long)g.Ho * g.Wo * g.ldc;
        a.r_sy = (long long)g.Wo * g.ldr; a.r_simg = (long long)g.Ho * g.Wo * g.ldr;
        a.ksx = a.ksy = g.ksize; a.dil = 1; a.pad_x = a.pad_y = g.pad;
    }
    *o = a;
    if (g.M < tc_min_rows()) return false;   // tiny problems stay on the fp32 FMA kernel
    return gemm_tc_supported(a);
}

}  // namespace zvx
