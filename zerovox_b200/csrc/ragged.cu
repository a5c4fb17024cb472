// Ragged <-> padded copies at the multi-GPU boundary (zerovox_b200/parallel.py).
//
// ZeroVox.forward returns zero-padded batches (wav [B, L*hop], mel [B, n_mels, L]; model.py:260-306) of which the
// consumer keeps wav[i][:mel_len[i]*hop] and mel[i][:, :mel_len[i]] (utils/export_hifigan.py:138-151).  A shard therefore
// packs only those valid parts into one contiguous buffer before the NCCL gather, and rank 0 may expand the gathered
// buffer back to the padded tensors.  Both directions are pure HBM copies: utterance b occupies rows * len[b] * unit
// elements of the packed buffer starting at off[b], laid out [rows][len[b] * unit].
#include "../../include/zerovox_b200.h"
#include "common.cuh"

#include <algorithm>
#include <string>

namespace zvx {
namespace {

thread_local std::string g_ragged_error;

// PACK: padded -> packed.  !PACK: packed -> padded, elements past the utterance's length zero-filled when zero_tail.
template <bool PACK, int VEC>
__global__ void __launch_bounds__(256) ragged_copy_kernel(float* __restrict__ padded, float* __restrict__ packed,
                                                          const int64_t* __restrict__ lens, const int64_t* __restrict__ offs,
                                                          long long b_stride, long long row_stride, int rows, int unit,
                                                          long long max_units, int zero_tail) {
    const int b = blockIdx.z, r = blockIdx.y;
    long long n = lens[b];
    n = (n < 0 ? 0 : (n > max_units ? max_units : n)) * unit;           // valid elements of this row
    const long long width = PACK ? n : (zero_tail ? max_units * unit : n);
    float* prow = padded + (long long)b * b_stride + (long long)r * row_stride;
    float* qrow = packed + offs[b] + (long long)r * n;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < width; i += (long long)gridDim.x * blockDim.x * VEC) {
        if (VEC == 4) {
            if (PACK) *reinterpret_cast<float4*>(qrow + i) = *reinterpret_cast<const float4*>(prow + i);
            else *reinterpret_cast<float4*>(prow + i) = (i < n) ? *reinterpret_cast<const float4*>(qrow + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            if (PACK) qrow[i] = prow[i];
            else prow[i] = (i < n) ? qrow[i] : 0.f;
        }
    }
}

int run(bool pack, float* padded, float* packed, const int64_t* lens, const int64_t* offs, long long b_stride,
        long long row_stride, int rows, int unit, int B, long long max_units, int zero_tail, cudaStream_t st) {
    try {
        ZVX_REQUIRE(padded && packed && lens && offs, "zvx_ragged: null pointer");
        ZVX_REQUIRE(rows >= 1 && unit >= 1 && B >= 0 && max_units >= 0 && rows <= 65535 && B <= 65535, "zvx_ragged: bad sizes");
        if (B == 0 || max_units == 0) return 0;
        // 128-bit accesses when every row start stays 16-byte aligned on both sides (the waveform: unit = hop)
        const bool vec = (unit % 4 == 0) && (b_stride % 4 == 0) && (row_stride % 4 == 0 || rows == 1) &&
                         ((reinterpret_cast<uintptr_t>(padded) | reinterpret_cast<uintptr_t>(packed)) & 15) == 0;
        const long long width = max_units * unit;
        const int per = 256 * (vec ? 4 : 1);
        // a few CTAs per row, capped so that the grid stays a small multiple of the SM count for large batches
        int gx = (int)std::min<long long>((width + per - 1) / per, std::max<long long>(1, (8LL * 148) / ((long long)B * rows) + 1));
        dim3 grid(gx, rows, B);
        if (pack) {
            if (vec) ragged_copy_kernel<true, 4><<<grid, 256, 0, st>>>(padded, packed, lens, offs, b_stride, row_stride, rows, unit, max_units, 0);
            else ragged_copy_kernel<true, 1><<<grid, 256, 0, st>>>(padded, packed, lens, offs, b_stride, row_stride, rows, unit, max_units, 0);
        } else {
            if (vec) ragged_copy_kernel<false, 4><<<grid, 256, 0, st>>>(padded, packed, lens, offs, b_stride, row_stride, rows, unit, max_units, zero_tail);
            else ragged_copy_kernel<false, 1><<<grid, 256, 0, st>>>(padded, packed, lens, offs, b_stride, row_stride, rows, unit, max_units, zero_tail);
        }
        ZVX_POST_LAUNCH();
        return 0;
    } catch (const std::exception& e) {
        g_ragged_error = e.what();
        return 1;
    }
}

}  // namespace
}  // namespace zvx

extern "C" {

int zvx_ragged_pack(const float* padded, int64_t b_stride, int64_t row_stride, int rows, int unit, int B, int64_t max_units,
                    const int64_t* lens, const int64_t* offs, float* packed, void* stream) {
    return zvx::run(true, const_cast<float*>(padded), packed, lens, offs, b_stride, row_stride, rows, unit, B, max_units, 0,
                    (cudaStream_t)stream);
}

int zvx_ragged_unpack(const float* packed, const int64_t* lens, const int64_t* offs, int rows, int unit, int B,
                      int64_t max_units, float* padded, int64_t b_stride, int64_t row_stride, int zero_tail, void* stream) {
    return zvx::run(false, padded, const_cast<float*>(packed), lens, offs, b_stride, row_stride, rows, unit, B, max_units,
                    zero_tail, (cudaStream_t)stream);
}

const char* zvx_ragged_last_error(void) { return zvx::g_ragged_error.c_str(); }

}  // extern "C"
