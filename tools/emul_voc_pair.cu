// Host emulation of voc_pair_kernel (csrc/voc_pair.cu): runs the REAL plan / schedule / weight-image builders and replays
// the kernel's data movement (residue-major layouts, descriptor offsets, trimmed-N MMAs, epilogue scatter, halo bounds) in
// double precision on the CPU against a direct ResBlock1, so indexing errors are found without a GPU.
//   nvcc -std=c++17 -O1 -o build/emul_voc_pair tools/emul_voc_pair.cu && build/emul_voc_pair
#include "../zerovox_b200/csrc/voc_pair.cu"
namespace zvx { long long g_launches = 0; int g_pdl = 0; }
#include <cmath>
#include <random>
using namespace zvx;

static double lrelu_d(double v, double s) { return v > 0 ? v : v * s; }

static int layout_row_h(int tau, int d, int P, int S, int G) {
    if (tau < 0) return -1;
    const int v = tau / d, rho = tau - v * d;
    int lg = 0; while ((1 << lg) < P) ++lg;
    const int n = (v >> lg) * d + rho;
    return n < 128 ? (v & (P - 1)) * S + G + n : -1;
}

int run_case(int C, int k, std::vector<int> dils, int T) {
    const int CQ = C / 4, P = 128 / C, NPH = P / 2, nd = (int)dils.size();
    std::mt19937 rng(C * 131 + k * 7 + T);
    std::normal_distribution<double> N01(0.0, 1.0);
    VocResArgs a; a.C = C; a.k = k; a.T = T; a.B = 1; a.nsteps = 0;
    std::vector<std::vector<float>> W, Bv, img;
    for (int i = 0; i < nd; ++i) for (int h = 0; h < 2; ++h) {
        std::vector<float> w((size_t)C * C * k), b((size_t)C);
        for (auto& v : w) v = (float)(N01(rng) / std::sqrt((double)C * k));
        for (auto& v : b) v = (float)(0.1 * N01(rng));
        W.push_back(w); Bv.push_back(b); img.push_back(voc_pair_pack_weight(w.data(), b.data(), C, k));
        a.steps[a.nsteps].dil = h ? 1 : dils[i]; a.steps[a.nsteps].kind = h; a.nsteps++;
    }
    static PairPlan p;
    if (!make_plan(a, &p)) { printf("C=%d k=%d: no plan\n", C, k); return 1; }
    printf("C=%2d k=%2d dils=", C, k); for (int d : dils) printf("%d,", d);
    printf(" P=%d S=%d G=%d Rtot=%d lo=%d TT=%d ZC=%d nwbuf=%d ctas=%d smem=%d mmas/step=%d\n", P, p.S, p.G, p.Rtot, p.lo, p.TT, p.ZC, p.nwbuf,
           p.ctas, p.smem_bytes, p.sched_off[1]);
    std::vector<double> x((size_t)T * C), ref, cur;
    for (auto& v : x) v = N01(rng);
    // direct reference (the images hold TF32-rounded weights: use them through their own rounding for an exact comparison)
    auto rn = [](float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x0FFFu + ((u >> 13) & 1u)) & ~0x1FFFu; memcpy(&v, &u, 4); return (double)v; };
    cur = x;
    for (int i = 0; i < nd; ++i) {
        std::vector<double> xt((size_t)T * C), y((size_t)T * C);
        for (int h = 0; h < 2; ++h) {
            const int d = h ? 1 : dils[i], c = (k - 1) / 2;
            const std::vector<double>& in = h ? xt : cur;
            std::vector<double> o((size_t)T * C);
            for (int t = 0; t < T; ++t) for (int co = 0; co < C; ++co) {
                double acc = Bv[2 * i + h][co];
                for (int j = 0; j < k; ++j) { const int ti = t + (j - c) * d; if (ti < 0 || ti >= T) continue;
                    for (int ci = 0; ci < C; ++ci) acc += rn(W[2 * i + h][((size_t)co * C + ci) * k + j]) * lrelu_d(in[(size_t)ti * C + ci], 0.1); }
                o[(size_t)t * C + co] = acc;
            }
            if (h == 0) xt = o; else y = o;
        }
        for (size_t e = 0; e < cur.size(); ++e) cur[e] += y[e];
    }
    ref = cur;
    // emulated kernel
    std::vector<double> out((size_t)T * C, 1e30);
    const int tiles = (T + p.TT - 1) / p.TT, S = p.S, G = p.G, Rtot = p.Rtot;
    for (int tile = 0; tile < tiles; ++tile) {
        const int tbase = tile * p.TT - p.lo;
        std::vector<double> A((size_t)CQ * Rtot * 4, 0.0), Z(256 * 4, 0.0);
        std::vector<double> xo((size_t)128 * P * C, 0.0);   // residual stream by (n, r, ch)
        auto Aat = [&](int cq, int row) -> double* { return &A[((size_t)cq * Rtot + row) * 4]; };
        for (int tau = 0; tau < p.R; ++tau) {
            const int t = tbase + tau;
            for (int ch = 0; ch < C; ++ch) {
                const double v = (t >= 0 && t < T) ? x[(size_t)t * C + ch] : 0.0;
                xo[((size_t)(tau / P) * P + tau % P) * C + ch] = v;
            }
        }
        for (int tau = 0; tau < p.R; ++tau) {
            const int row = layout_row_h(tau, p.ld[0], P, S, G);
            if (row < 0) continue;
            for (int ch = 0; ch < C; ++ch) Aat(ch / 4, row)[ch & 3] = lrelu_d(xo[((size_t)(tau / P) * P + tau % P) * C + ch], 0.1);
        }
        for (int s = 0; s < p.nsteps; ++s) {
            std::vector<double> D((size_t)128 * 128, 1e30);
            const std::vector<float>& wi = img[s];
            {   // the step's first MMA: ones (1, 1, 0, 0) x bias block rows (hi, lo, 0, 0)
                const size_t bb = (size_t)CQ * p.ZC * 4;
                for (int n = 0; n < 128; ++n) for (int j = 0; j < 128; ++j)
                    D[(size_t)n * 128 + j] = (double)wi[bb + (size_t)j * 4] + (double)wi[bb + (size_t)j * 4 + 1];
            }
            for (int i = p.sched_off[s]; i < p.sched_off[s + 1]; ++i) {
                const uint4 e4 = p.sched[i];
                const int ncols = (int)((e4.z >> 17) & 0x3F) << 3, col = (int)e4.w;
                if ((int)(e4.x >> 16) != Rtot || (int)(e4.y >> 16) != p.ZC) { printf("bad LBO\n"); return 1; }
                const int ao = e4.x & 0xFFFF, bo = e4.y & 0xFFFF;
                for (int n = 0; n < 128; ++n) for (int j = 0; j < ncols; ++j) {
                    double sum = 0;
                    for (int kk = 0; kk < 8; ++kk) {
                        const size_t ar = (size_t)ao + n + (kk / 4) * Rtot, br = (size_t)bo + j + (kk / 4) * p.ZC;
                        if (ar * 4 + 3 >= A.size() || br * 4 + 3 >= wi.size()) { printf("operand read out of range (step %d)\n", s); return 1; }
                        sum += A[ar * 4 + (kk & 3)] * (double)wi[br * 4 + (kk & 3)];
                    }
                    D[(size_t)n * 128 + col + j] += sum;
                }
            }
            const int kind = a.steps[s].kind, ds = p.ld[s];
            const bool last = s == p.nsteps - 1;
            const int dn = last ? 1 : p.ld[s + 1];
            for (int n = 0; n < 128; ++n) for (int r = 0; r < P; ++r) {
                int tau;
                if (ds == 1) tau = P * n + r; else { const int qs = n / ds, rho = n - qs * ds; tau = ds * (qs * P + r) + rho; }
                const int t = tbase + tau;
                const bool inside = t >= 0 && t < T;
                for (int ch = 0; ch < C; ++ch) {
                    const double c4 = D[(size_t)n * 128 + r * C + ch];
                    if (kind == 0) {
                        const int row = (tau >> p.lgP) < 128 ? (tau & (P - 1)) * S + G + (tau >> p.lgP) : -1;
                        if (row >= 0) Aat(ch / 4, row)[ch & 3] = inside ? lrelu_d(c4, 0.1) : 0.0;
                    } else {
                        double& xr = xo[((size_t)n * P + r) * C + ch];
                        if (!last) {
                            if (inside) xr += c4;
                            const int row = layout_row_h(tau, dn, P, S, G);
                            if (row >= 0) Aat(ch / 4, row)[ch & 3] = lrelu_d(xr, 0.1);
                        } else if (inside && tau >= p.lo && tau < p.lo + p.TT) out[(size_t)t * C + ch] = xr + c4;
                    }
                }
            }
        }
    }
    double worst = 0;
    for (size_t e = 0; e < out.size(); ++e) worst = std::max(worst, std::fabs(out[e] - ref[e]));
    printf("   max |emulated - direct| = %.3e %s\n", worst, worst < 2e-6 ? "ok" : "FAIL");
    return worst < 2e-6 ? 0 : 1;
}

int main() {
    int bad = 0;
    for (int C : {8, 16, 32})
        for (int k : {3, 7, 11}) bad += run_case(C, k, {1, 3, 5}, C == 8 ? 4500 : (C == 16 ? 2300 : 1100));
    bad += run_case(16, 5, {2, 6}, 1500);
    bad += run_case(32, 7, {1}, 300);
    bad += run_case(8, 3, {4, 1, 7}, 3000);
    printf(bad ? "FAILED\n" : "all ok\n");
    return bad;
}
