// TF32 tensor-core GEMM / implicit-GEMM convolution on tcgen05 (sm_100a): TMA-fed, TMEM accumulators.
#pragma once
#include "common.cuh"

namespace zvx {

// Generalised problem of the tcgen05 kernel.
//
//   out[img, y, x, n] = epilogue( sum_{dy,dx} sum_k A[img, y*s + dy*dil - pad_y, x*s + dx*dil - pad_x, k] * W[z1, z2][n, k] )
//
// A is a 4-D view (K contiguous) addressed through one TMA tensor map; positions outside [0,Hi)x[0,Wi) read as
// zero (TMA out-of-bounds fill = the convolutions' zero padding).  W is a 4-D view (K contiguous) with
//   (z1, z2) = (tap, 0)   tap = dy*ksx + dx      for weights shared by all tiles (Linear / Conv1d / Conv2d), or
//   (z1, z2) = (y, img)                          for per-(y,img) operands (attention: y = head, img = utterance).
// An output tile is TH x TW = 128 positions of one image (M of the MMA) by BN columns.
// Plain GEMM: Hi = Ho = IMG = 1, Wi = Wo = M.   Conv1d over [B, L, C]: Hi = Ho = B, Wi = Wo = L, ksy = 1.
// epilogue: v = acc + bias[n]; relu_first; v = v*scale[n] + shift[n]; v += R[...]; relu_last
// out / R element offset: img*c_simg + y*c_sy + x*c_sx + n*c_sn  (c_sn = 1: row-major; c_sx = 1: transposed store).
struct TcGemmArgs {
    const float* A = nullptr;
    int K = 0, Wi = 1, Hi = 1, IMG = 1;
    long long a_sx = 0, a_sy = 0, a_simg = 0;   // element strides of the x / y / img dims (K stride = 1)
    const float* W = nullptr;
    int N = 0, Z1 = 1, Z2 = 1;
    long long w_sn = 0, w_s1 = 0, w_s2 = 0;     // element strides of the n / z1 / z2 dims
    int b_batched = 0;
    // 3xTF32 split mode (fp32-grade products): TF32-exact low parts of A and W (tf32_split_lo), same layout / strides
    const float* A_lo = nullptr;
    const float* W_lo = nullptr;
    int Wo = 1, Ho = 1;
    int ksx = 1, ksy = 1, dil = 1, pad_x = 0, pad_y = 0;
    int stride = 1;                             // output stride of a Conv2d (1 or 2; TMA element strides), x and y alike
    float* C = nullptr;
    long long c_simg = 0, c_sy = 0, c_sx = 0, c_sn = 1;
    const float* R = nullptr;
    long long r_simg = 0, r_sy = 0, r_sx = 0;   // residual strides (all 0: same as the output's); n stride = c_sn
    const float* bias = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    int relu_first = 0, relu_last = 0;
    // vocoder extras, applied after the above:  v = v*acc_scale + (acc_init ? 0 : C)  when acc_mode;
    // C = lrelu(v, act_slope); C2 = lrelu(v, slope2) (optional second output, addressed like C)
    float act_slope = 1.f;
    float* C2 = nullptr; float slope2 = 1.f;
    int acc_mode = 0, acc_init = 0; float acc_scale = 1.f;
    double flops() const { return 2.0 * IMG * Ho * Wo * (double)N * K * ksx * ksy; }   // algorithmic (split issues 3x)
};

// Constraints of the TMA descriptors (16-byte aligned base pointers and strides) and of the tile shapes.
bool gemm_tc_supported(const TcGemmArgs& a);
void gemm_tc(const TcGemmArgs& a, cudaStream_t st);

// lo[i] = rn_tf32(x[i] - trunc_tf32(x[i])) for i < n (n rounded up to a multiple of 4; both buffers 16-byte aligned and
// padded accordingly).
void tf32_split_lo(const float* x, float* lo, long long n, cudaStream_t st);

// Conversion from the SIMT kernel's argument block (ROW_PLAIN / ROW_CONV1D / ROW_CONV2D with stride 1, nz == 1).
bool gemm_tc_from(const GemmArgs& g, TcGemmArgs* out);
int tc_min_rows();   // smallest M that goes to the tensor-core kernel (ZVX_TC_MIN_M)

}  // namespace zvx
