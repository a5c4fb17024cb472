// "Row-shift" convolution on the tensor cores for small / medium channel counts (C_in, C_out in {32, 64}; sm_100a).
//
// The generic TMA implicit GEMM (gemm_tc.cu) re-fetches the activation tile for every filter tap and, with N = C_out
// <= 64, spends most of its time on that L2 traffic (a tcgen05.mma costs the same issue slot whatever N is).  Here the
// activation tile (with its halo) is staged ONCE per CTA in the no-swizzle K-major layout [cq][row][4 floats], where a
// filter tap is nothing but a row offset of the operand descriptor:
//   Conv1d (HiFi-GAN stage with 64 channels): row r <-> sample t0 - h + r, tap j reads rows i + j*dil;
//   Conv2d 3x3 (ResNetSE34V2 32- and 64-channel stages): the (TH+2) x (TW+2) halo tile is stored line after line, so tap
//   (dy, dx) reads rows i + dy*(TW+2) + dx — a 2-D stencil as a linear shift (the 2 pad columns of every line produce
//   junk accumulator rows that are never stored).
// Accumulators of all M tiles (128 rows each) of the CTA tile live in TMEM at once (MT * N <= 512 columns), so the loop
// runs tap-outermost and the per-tap weight slab [cq][N][4] streams through a small ring of `cp.async.bulk` stages, each
// used by every M tile.  leaky-ReLU of the input (HiFi-GAN) and the TF32 rounding happen on the way into shared memory;
// the epilogue (8 warps) applies bias / ReLU / folded BatchNorm / residual / MRF-mean accumulation and stores
// channel-last rows.
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace zvx {

namespace {

constexpr int EPI_THREADS = 256;
constexpr int NT = EPI_THREADS + 32;
constexpr int WSTAGES = 3;

struct RsPlan {
    int mode;              // 0 Conv1d, 1 Conv2d 3x3
    int C, N, CQ, taps;
    int MT;                // M tiles per CTA
    int rows_in, Rp;       // staged rows (needed / allocated, bank-padded)
    int TT;                // 1-D: outputs per CTA
    int TH, TW, LW;        // 2-D: tile lines / positions per line / staged line width (TW + 2)
    int tiles_x, tiles_y;
    int shift[16];         // row shift of every tap
    int tap_bytes;
    uint32_t offA, offW, offBar;
    uint32_t idesc;
    int tmem_cols;
    int smem_bytes;
};

__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ uint64_t nosw_desc(uint32_t saddr, uint32_t lbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo16 & 0x3FFFu) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(NT) conv_rs_kernel(const ConvRsArgs a, const RsPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sb = smem_u32(smem);
    const uint32_t sA = sb + p.offA;
    // barriers: w_full[WSTAGES] | w_empty[WSTAGES] | done
    const uint32_t bars = sb + p.offBar;
    auto w_full = [&](int i) { return bars + 8u * (uint32_t)i; };
    auto w_empty = [&](int i) { return bars + 8u * (uint32_t)(WSTAGES + i); };
    const uint32_t done_bar = bars + 8u * (2 * WSTAGES);
    const uint32_t slot = done_bar + 8;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.offBar + 8 * (2 * WSTAGES + 1));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int CQ = p.CQ, Rp = p.Rp;

    // tile origin
    int t0 = 0, y0 = 0, x0 = 0;
    if (p.mode == 0) t0 = blockIdx.x * p.TT;
    else { y0 = (blockIdx.x / p.tiles_x) * p.TH; x0 = (blockIdx.x % p.tiles_x) * p.TW; }

    if (tid == 0) {
        for (int i = 0; i < WSTAGES; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == EPI_THREADS / 32) tmem_alloc(slot, (uint32_t)p.tmem_cols);

    if (warp < EPI_THREADS / 32) {
        // ---- activation tile: global (channel-last) -> lrelu -> TF32 -> [cq][row][4]; out-of-range = zero padding
        const float* __restrict__ xb = a.x + (long long)b * a.x_bs;
        const int total = p.rows_in * CQ;
        for (int base = 0; base < total; base += EPI_THREADS * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * EPI_THREADS + tid;
                const int row = idx / CQ, cq = idx - row * CQ;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total) {
                    if (p.mode == 0) {
                        const int t = t0 - p.shift[(p.taps - 1) / 2] + row;   // shift of the centre tap = one-sided halo
                        if (t >= 0 && t < a.T) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * a.C) + cq);
                    } else {
                        const int ly = row / p.LW, lx = row - ly * p.LW;
                        const int y = y0 - 1 + ly, x = x0 - 1 + lx;
                        if (y >= 0 && y < a.H && x >= 0 && x < a.W)
                            v[u] = __ldg(reinterpret_cast<const float4*>(xb + ((long long)y * a.W + x) * a.C) + cq);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * EPI_THREADS + tid;
                if (idx < total) {
                    const int row = idx / CQ, cq = idx - row * CQ;
                    float4 o;
                    o.x = rna_tf32(lrelu(v[u].x, a.in_slope)); o.y = rna_tf32(lrelu(v[u].y, a.in_slope));
                    o.z = rna_tf32(lrelu(v[u].z, a.in_slope)); o.w = rna_tf32(lrelu(v[u].w, a.in_slope));
                    st_shared_v4(sA + (uint32_t)((cq * Rp + row) * 16), o);
                }
            }
        }
        // rows past the staged ones feed junk accumulator rows only: keep them finite
        for (int idx = tid; idx < (Rp - p.rows_in) * CQ; idx += EPI_THREADS) {
            const int row = p.rows_in + idx / CQ, cq = idx % CQ;
            st_shared_v4(sA + (uint32_t)((cq * Rp + row) * 16), make_float4(0.f, 0.f, 0.f, 0.f));
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;

    if (warp == EPI_THREADS / 32) {
        // ================================================================ weight streamer + MMA issuer (one thread)
        if (lane == 0) {
            auto load_tap = [&](int t) {
                const int st = t % WSTAGES;
                mbar_arrive_expect_tx(w_full(st), (uint32_t)p.tap_bytes);
                bulk_load_1d(sb + p.offW + (uint32_t)(st * p.tap_bytes), a.w + (long long)t * (p.tap_bytes / 4), (uint32_t)p.tap_bytes,
                             w_full(st));
            };
            for (int t = 0; t < WSTAGES && t < p.taps; ++t) load_tap(t);
            const uint64_t da0 = nosw_desc(sA, (uint32_t)Rp);
            for (int t = 0; t < p.taps; ++t) {
                const int st = t % WSTAGES;
                mbar_wait_spin(w_full(st), (uint32_t)((t / WSTAGES) & 1));
                tc_fence_after();
                const uint64_t db0 = nosw_desc(sb + p.offW + (uint32_t)(st * p.tap_bytes), (uint32_t)p.N);
                const int sh = p.shift[t];
                for (int m = 0; m < p.MT; ++m) {
                    uint64_t da = da0 + (uint64_t)(m * 128 + sh);
                    uint64_t db = db0;
                    for (int pp = 0; pp < p.C / 8; ++pp) {
                        umma_tf32(tmem_base + (uint32_t)(m * p.N), da, db, p.idesc, (t | pp) ? 1u : 0u);
                        da += (uint64_t)(2 * Rp);
                        db += (uint64_t)(2 * p.N);
                    }
                }
                umma_commit(w_empty(st));
                // refill the stage of the PREVIOUS tap (its MMAs are ahead of this tap's in the pipe)
                if (t >= 1 && t - 1 + WSTAGES < p.taps) {
                    mbar_wait_spin(w_empty((t - 1) % WSTAGES), (uint32_t)(((t - 1) / WSTAGES) & 1));
                    load_tap(t - 1 + WSTAGES);
                }
            }
            umma_commit(done_bar);
        }
        __syncwarp();
    } else {
        // ================================================================ epilogue (warps 0..7)
        const int q = warp & 3, g = warp >> 2;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int m = g; m < p.MT; m += 2) {
            const int i = m * 128 + q * 32 + lane;
            bool valid;
            long long off;
            if (p.mode == 0) {
                const int t = t0 + i;
                valid = (i < p.TT) && (t < a.T);
                off = (long long)b * a.y_bs + (long long)t * a.N;
            } else {
                const int ly = i / p.LW, lx = i - ly * p.LW;
                const int y = y0 + ly, x = x0 + lx;
                valid = (ly < p.TH) && (lx < p.TW) && (y < a.H) && (x < a.W);
                off = (long long)b * a.y_bs + ((long long)y * a.W + x) * a.N;
            }
            for (int c0 = 0; c0 < p.N; c0 += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(trow + (uint32_t)(m * p.N + c0), v);
                tmem_wait_ld();
                if (!valid) continue;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const int n = c0 + 4 * gq;
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float xv = __uint_as_float(v[4 * gq + e]);
                        if (a.bias) xv += __ldg(a.bias + n + e);
                        if (a.relu_first) xv = fmaxf(xv, 0.f);
                        if (a.scale) xv = fmaf(xv, __ldg(a.scale + n + e), __ldg(a.shift + n + e));
                        o[e] = xv;
                    }
                    if (a.R) {
                        const float4 r = *reinterpret_cast<const float4*>(a.R + (long long)b * a.r_bs + (off - (long long)b * a.y_bs) + n);
                        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
                    }
                    float* yp = a.y + off + n;
                    if (a.acc_mode) {
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!a.acc_init) s = *reinterpret_cast<const float4*>(yp);
                        o[0] = fmaf(o[0], a.acc_scale, s.x); o[1] = fmaf(o[1], a.acc_scale, s.y);
                        o[2] = fmaf(o[2], a.acc_scale, s.z); o[3] = fmaf(o[3], a.acc_scale, s.w);
                    }
                    *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                    if (a.y2)
                        *reinterpret_cast<float4*>(a.y2 + off + n) =
                            make_float4(lrelu(o[0], a.slope2), lrelu(o[1], a.slope2), lrelu(o[2], a.slope2), lrelu(o[3], a.slope2));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EPI_THREADS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

int pad_rows(int rows, int CQ) {
    const int want = (8 / CQ) & 7;   // chunk planes rows*16 B apart: spread a quarter-warp's 8 stores over all banks
    int r = rows;
    while ((r & 7) != want) ++r;
    return r;
}

bool make_plan(const ConvRsArgs& a, RsPlan* out) {
    RsPlan p{};
    p.mode = a.mode; p.C = a.C; p.N = a.N; p.CQ = a.C / 4;
    if (!(a.C == 32 || a.C == 64) || !(a.N == 32 || a.N == 64)) return false;
    p.taps = a.mode == 0 ? a.k : 9;
    if (p.taps < 1 || p.taps > 16 || (a.mode == 0 && (a.k & 1) == 0)) return false;
    p.tap_bytes = p.CQ * a.N * 16;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int w_bytes = WSTAGES * p.tap_bytes;
    // two co-resident CTAs per SM (their load / MMA / epilogue phases overlap) whenever a useful tile fits in half the SM
    static const int kb = getenv("ZVX_RS_KB") ? atoi(getenv("ZVX_RS_KB")) : 110;
    const int budget = kb * 1024 - w_bytes - 256;
    const int row_bytes = a.C * 4;
    bool found = false;
    if (a.mode == 0) {
        const int h = (a.k - 1) / 2 * a.dil;
        for (int j = 0; j < a.k; ++j) p.shift[j] = j * a.dil;
        for (int mt = 512 / a.N; mt >= 1 && !found; --mt) {
            if (mt > 8) continue;
            const int rows = 128 * mt + 2 * h;
            const int rp = pad_rows(rows, p.CQ);
            if ((long long)rp * row_bytes > budget) continue;
            p.MT = mt; p.TT = 128 * mt; p.rows_in = rows; p.Rp = rp;
            p.tiles_x = cdiv(a.T, p.TT); p.tiles_y = 1;
            found = true;
        }
        (void)h;
    } else {
        // tile search: maximise useful outputs per computed accumulator row and per staged row
        double best = 0.0;
        for (int th = 1; th <= 8; ++th)
            for (int tw = 16; tw <= 512; tw += 2) {
                if (tw > a.W + 1 && tw > 16) break;
                const int lw = tw + 2;
                const int acc_rows = th * lw;
                const int mt = cdiv(acc_rows, 128);
                if (mt * a.N > 512 || mt > 8) continue;
                const int rows = (th + 2) * lw;
                const int need = std::max(rows, 128 * mt + 2 * lw + 2);
                const int rp = pad_rows(need, p.CQ);
                if ((long long)rp * row_bytes > budget) continue;
                const int tx = cdiv(a.W, tw), ty = cdiv(a.H, th);
                const double useful = (double)a.W * a.H;
                const double cost = (double)tx * ty * (mt * 128.0 + 0.35 * rows);   // MMA rows + (cheaper) staged rows
                const double score = useful / cost;
                if (score > best) {
                    best = score;
                    p.MT = mt; p.TH = th; p.TW = tw; p.LW = lw; p.rows_in = rows; p.Rp = rp; p.tiles_x = tx; p.tiles_y = ty;
                    found = true;
                }
            }
        if (found)
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx) p.shift[dy * 3 + dx] = dy * p.LW + dx;
    }
    if (!found) return false;
    int cols = 32;
    while (cols < p.MT * a.N) cols <<= 1;
    p.tmem_cols = cols;
    uint32_t o = 0;
    p.offA = o; o += (uint32_t)(p.Rp * row_bytes);
    o = (uint32_t)round_up(o, 128);
    p.offW = o; o += (uint32_t)w_bytes;
    p.offBar = o; o += 8 * (2 * WSTAGES + 1) + 16;
    p.smem_bytes = (int)o + 128;
    if (p.smem_bytes > 227 * 1024) return false;
    *out = p;
    return true;
}

}  // namespace

bool conv_rs_supported(const ConvRsArgs& a) {
    RsPlan p;
    if (!a.x || !a.w || !a.y || a.B < 1 || a.B > 65535) return false;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.y) & 15) || (a.x_bs & 3) || (a.y_bs & 3)) return false;
    if (a.R && ((reinterpret_cast<uintptr_t>(a.R) & 15) || (a.r_bs & 3))) return false;
    if (a.mode == 0 ? (a.T < 1) : (a.H < 1 || a.W < 1)) return false;
    return make_plan(a, &p);
}

void conv_rs(const ConvRsArgs& a, cudaStream_t st) {
    RsPlan p;
    ZVX_REQUIRE(conv_rs_supported(a) && make_plan(a, &p), "conv_rs: unsupported problem");
    static int attr_done = 0;
    if (!attr_done) {
        ZVX_CUDA_CHECK(cudaFuncSetAttribute(conv_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = 1;
    }
    dim3 grid(p.tiles_x * p.tiles_y, a.B);
    conv_rs_kernel<<<grid, NT, p.smem_bytes, st>>>(a, p);
    ZVX_POST_LAUNCH();
}

}  // namespace zvx
