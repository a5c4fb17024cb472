// Speaker-prompt front-end: silence trim + log-mel spectrogram, the step immediately before
// `_spkemb` in ZeroVoxTTS.speaker_embed (zerovox/tts/synthesize.py:123-143), which the reference runs on the CPU through
// librosa (zerovox/tts/mels.py:356-394).  One fused kernel turns waveform frames into log-mel rows in the [B, T_ref, n_mels]
// layout zvx_spkemb consumes: reflect padding by index arithmetic, periodic Hann window, a 1024-point real FFT done as a
// 512-point complex FFT (three radix-8 passes through shared memory, 64 threads per frame), magnitudes, the sparse Slaney
// mel filterbank, log(clip(., 1e-5)) and the per-frame spectral energy — the waveform is read once, nothing intermediate
// touches HBM.  Independent of the model weights, so it has its own small handle (zvx_frontend).
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/zerovox_b200.h"
#include "common.cuh"

namespace zvx {

constexpr int FE_NFFT = 1024;          // real FFT length (every reference config: fft_size 1024)
constexpr int FE_TPF = 64;             // threads per frame: one radix-8 butterfly each per pass
constexpr int FE_FPC = 4;              // frames per CTA
constexpr int FE_BUF = 8 * 72;         // float2 slots of the exchange buffer (padded strides 68 / 72+9, see passes)
constexpr int FE_MAXMEL = 256;

struct MelParams {
    const float2* tw512;     // W_512^j = exp(-2*pi*i*j/512), j < 512
    const float2* tw1024;    // W_1024^k, k <= 512
    const float* window;     // [1024] periodic Hann, win_length centred in fft_size
    const int* f_start;      // per mel filter: first bin with a non-zero weight
    const int* f_count;      // number of consecutive non-zero bins
    const int* f_off;        // offset of its weights in f_w
    const float* f_w;
    int num_mels, hop, pad;
    float clip;
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// Forward 8-point DFT in registers, natural order in and out (decimation in frequency: one radix-2 split, two DFT-4).
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    float2 a4 = d0;
    float2 a5 = make_float2((d1.x + d1.y) * h, (d1.y - d1.x) * h);      // * W8^1
    float2 a6 = mul_mi(d2);                                              // * W8^2
    float2 a7 = make_float2((d3.y - d3.x) * h, -(d3.x + d3.y) * h);     // * W8^3
    {   // even outputs: DFT-4 of a0..a3
        float2 c0 = cadd(a0, a2), c2 = csub(a0, a2), c1 = cadd(a1, a3), c3 = mul_mi(csub(a1, a3));
        v[0] = cadd(c0, c1); v[4] = csub(c0, c1); v[2] = cadd(c2, c3); v[6] = csub(c2, c3);
    }
    {   // odd outputs: DFT-4 of a4..a7
        float2 c0 = cadd(a4, a6), c2 = csub(a4, a6), c1 = cadd(a5, a7), c3 = mul_mi(csub(a5, a7));
        v[1] = cadd(c0, c1); v[5] = csub(c0, c1); v[3] = cadd(c2, c3); v[7] = csub(c2, c3);
    }
}

// number of STFT frames of a waveform of `len` samples after reflect padding by `pad` on both sides (center=False)
__host__ __device__ __forceinline__ long long fe_num_frames(long long len, int hop, int pad) {
    long long padded = len + 2LL * pad;
    return (len > pad && padded >= FE_NFFT) ? 1 + (padded - FE_NFFT) / hop : 0;
}

__global__ void __launch_bounds__(FE_TPF* FE_FPC)
mel_spectrogram_kernel(const float* __restrict__ wav, long long n_stride, const long long* __restrict__ wav_start,
                       const long long* __restrict__ wav_len, int n_frames, MelParams p, float* __restrict__ mel,
                       float* __restrict__ energy) {
    __shared__ float2 bufs[FE_FPC][FE_BUF];
    __shared__ float mags[FE_FPC][FE_NFFT / 2 + 8];
    __shared__ float esum[FE_FPC][2];

    const int fl = threadIdx.x / FE_TPF, t = threadIdx.x % FE_TPF;
    const int b = blockIdx.y;
    const int f = blockIdx.x * FE_FPC + fl;
    // the window is clamped to the row: a bad (start, len) pair can shorten the output but never read outside the buffer
    const long long start = wav_start ? min(max(wav_start[b], 0LL), n_stride) : 0;
    const long long len = wav_len ? min(max(wav_len[b], 0LL), n_stride - start) : n_stride - start;
    const bool valid = f < n_frames && f < fe_num_frames(len, p.hop, p.pad);
    float2* buf = bufs[fl];
    float* mag = mags[fl];
    const float* x = wav + (long long)b * n_stride + start;

    float2 v[8];
    // ---- pass 1: thread r = n mod 64 takes z[64*n2 + r], n2 = 0..7 (z[c] = x[2c] + i*x[2c+1], windowed), DFT over n2,
    //      twiddle W_512^(r*k0); buf[k0*68 + r]
    {
        const long long base = (long long)f * p.hop - p.pad;
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            const int c = 64 * n2 + t;
            float s0 = 0.f, s1 = 0.f;
            if (valid) {
                long long i0 = base + 2 * c, i1 = i0 + 1;
                i0 = i0 < 0 ? -i0 : (i0 >= len ? 2 * (len - 1) - i0 : i0);      // np.pad(mode='reflect'), mels.py:384-385
                i1 = i1 < 0 ? -i1 : (i1 >= len ? 2 * (len - 1) - i1 : i1);
                s0 = __ldg(x + i0);
                s1 = __ldg(x + i1);
            }
            const float2 w = __ldg(reinterpret_cast<const float2*>(p.window) + c);
            v[n2] = make_float2(s0 * w.x, s1 * w.y);
        }
        dft8(v);
        buf[t] = v[0];
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) buf[k0 * 68 + t] = cmul(v[k0], __ldg(p.tw512 + t * k0));
    }
    __syncthreads();
    // ---- pass 2: thread (k0, n0) takes r = 8*n1 + n0, DFT over n1, twiddle W_64^(n0*k1); buf[k1*72 + k0*9 + n0]
    {
        const int k0 = t >> 3, n0 = t & 7;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) v[n1] = buf[k0 * 68 + 8 * n1 + n0];
        __syncthreads();
        dft8(v);
        buf[k0 * 9 + n0] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) buf[k1 * 72 + k0 * 9 + n0] = cmul(v[k1], __ldg(p.tw512 + 8 * n0 * k1));
    }
    __syncthreads();
    // ---- pass 3: thread (k1, k0) takes n0 = 0..7, DFT over n0 -> Z[k0 + 8*k1 + 64*k2] = Z[t + 64*k2]
    {
        const int k1 = t >> 3, k0 = t & 7;
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) v[n0] = buf[k1 * 72 + k0 * 9 + n0];
        __syncthreads();
        dft8(v);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) buf[t + 64 * k2] = v[k2];
    }
    __syncthreads();
    // ---- real-input split: X[k] = E[k] + W_1024^k * O[k], E/O from Z[k] and conj(Z[512-k]); magnitudes (np.abs, mels.py:389)
    float e2 = 0.f;
    for (int k = t; k <= FE_NFFT / 2; k += FE_TPF) {
        const float2 zk = buf[k & 511];
        float2 zc = buf[(512 - k) & 511];
        zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = csub(zk, zc);
        const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);
        const float2 xk = cadd(e, cmul(__ldg(p.tw1024 + k), o));
        const float m2 = xk.x * xk.x + xk.y * xk.y;
        mag[k] = sqrtf(m2);
        e2 += m2;
    }
    e2 = warp_sum(e2);
    if ((t & 31) == 0) esum[fl][t >> 5] = e2;
    __syncthreads();
    if (f >= n_frames) return;
    // ---- mel filterbank (np.dot(mel_basis, magnitudes), mels.py:391), log(clip) (mels.py:350-351), energy (mels.py:394)
    float* out = mel + ((long long)b * n_frames + f) * p.num_mels;
    for (int m = t; m < p.num_mels; m += FE_TPF) {
        float s = 0.f;
        if (valid) {
            const int j0 = p.f_start[m], cnt = p.f_count[m];
            const float* w = p.f_w + p.f_off[m];
            for (int j = 0; j < cnt; ++j) s = fmaf(__ldg(w + j), mag[j0 + j], s);
            s = logf(fmaxf(s, p.clip));
        }
        out[m] = s;
    }
    if (energy && t == 0) energy[(long long)b * n_frames + f] = valid ? sqrtf(esum[fl][0] + esum[fl][1]) : 0.f;
}

// ---- librosa.effects.trim (synthesize.py:126): frame RMS over zero-padded centred frames ...
__global__ void __launch_bounds__(256)
frame_rms_kernel(const float* __restrict__ wav, long long n_stride, const long long* __restrict__ wav_len,
                 int n_frames_max, int frame_length, int hop, float* __restrict__ rms, unsigned* __restrict__ rmax_bits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const long long f = (long long)blockIdx.x * 8 + warp;
    const long long len = wav_len ? min(max(wav_len[b], 0LL), n_stride) : n_stride;
    if (f >= n_frames_max || f > len / hop) return;                  // 1 + len // hop frames
    const float* x = wav + (long long)b * n_stride;
    const long long lo = f * hop - frame_length / 2;
    float s = 0.f;
    for (int j = lane; j < frame_length; j += 32) {
        const long long i = lo + j;
        const float v = (i >= 0 && i < len) ? __ldg(x + i) : 0.f;
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
        const float r = sqrtf(s / (float)frame_length);
        rms[(long long)b * n_frames_max + f] = r;
        atomicMax(rmax_bits + b, __float_as_uint(r));                // r >= 0: the bit patterns order like the values
    }
}

// ... dB relative to the loudest frame; first / last frame above -top_db -> [start, end) in samples.
__global__ void __launch_bounds__(256)
trim_bounds_kernel(const float* __restrict__ rms, const unsigned* __restrict__ rmax_bits,
                   const long long* __restrict__ wav_len, long long n_stride, int n_frames_max, int hop, float top_db,
                   long long* __restrict__ start_out, long long* __restrict__ len_out) {
    __shared__ int s_first, s_last;
    const int b = blockIdx.x;
    const long long len = wav_len ? min(max(wav_len[b], 0LL), n_stride) : n_stride;
    const long long nfr = min((long long)n_frames_max, 1 + len / hop);
    if (threadIdx.x == 0) { s_first = 0x7fffffff; s_last = -1; }
    __syncthreads();
    const float amin = 1e-10f;
    const float ref = __uint_as_float(rmax_bits[b]);
    const float ref_db = 10.f * log10f(fmaxf(amin, ref * ref));
    int first = 0x7fffffff, last = -1;
    for (long long f = threadIdx.x; f < nfr; f += blockDim.x) {
        const float r = rms[(long long)b * n_frames_max + f];
        const float db = 10.f * log10f(fmaxf(amin, r * r)) - ref_db;
        if (db > -top_db) { first = min(first, (int)f); last = max(last, (int)f); }
    }
    atomicMin(&s_first, first);
    atomicMax(&s_last, last);
    __syncthreads();
    if (threadIdx.x == 0) {
        long long st = 0, en = 0;
        if (s_last >= 0) {
            st = (long long)s_first * hop;
            en = min(len, (long long)(s_last + 1) * hop);
        }
        start_out[b] = st;
        len_out[b] = en - st;
    }
}

// ------------------------------------------------------------------------------------------------ resampler
// Band-limited sample-rate conversion of the speaker prompt (what `librosa.load(path, sr=sampling_rate)` does in front of the
// path, synthesize.py:113-121: the packaged prompts are 24 kHz, the models 22.05 kHz).  One thread per output sample m:
//   t = m * down / up (input samples),  y[m] = sum_j w[phase][j] * x[floor(t) - KH + 1 + j],   phase = (m * down) mod up,
// w = the Kaiser-windowed sinc of resampy's published "kaiser_best" design (64 zero crossings, beta 14.7697, roll-off
// 0.9476), time-scaled by min(1, sr_out / sr_in) and tabulated exactly (no table interpolation) per rational phase.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long x_stride, const long long* __restrict__ x_len,
                                                       const float* __restrict__ tab, int up, int down, int taps, int kh,
                                                       float* __restrict__ y, long long y_stride, long long n_out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (m >= n_out) return;
    const long long n_in = x_len ? min(max(x_len[b], 0LL), x_stride) : x_stride;
    const long long num = m * down, i0 = num / up;
    const int phase = (int)(num - i0 * up);
    const float* __restrict__ w = tab + (long long)phase * taps;
    const float* __restrict__ xb = x + (long long)b * x_stride;
    const long long k0 = i0 - kh + 1;
    float acc = 0.f;
    for (int j = 0; j < taps; ++j) {
        const long long k = k0 + j;
        if (k >= 0 && k < n_in) acc = fmaf(__ldg(w + j), __ldg(xb + k), acc);
    }
    y[(long long)b * y_stride + m] = acc;
}

struct ResampleTable { int up = 1, down = 1, taps = 0, kh = 0; float* tab = nullptr; };

static double bessel_i0(double x) {   // power series, converges fast for the argument range of a Kaiser window
    double s = 1.0, term = 1.0;
    for (int k = 1; k < 200; ++k) {
        term *= (x / (2.0 * k)) * (x / (2.0 * k));
        s += term;
        if (term < 1e-18 * s) break;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------ host side
static std::string g_fe_create_error;

struct Frontend {
    zvx_mel_config cfg;
    int device = 0;
    std::string err;
    MelParams p{};
    std::vector<void*> owned;
    float* rms = nullptr; size_t rms_cap = 0;
    unsigned* rmax = nullptr; size_t rmax_cap = 0;
    std::map<std::pair<int, int>, ResampleTable> rs_tabs;   // per (sr_in, sr_out), built on first use

    template <typename T> T* upload(const std::vector<T>& h) {
        T* d = nullptr;
        ZVX_CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(1, h.size()) * sizeof(T)));
        owned.push_back(d);
        if (!h.empty()) ZVX_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        return d;
    }

    // librosa.hz_to_mel / mel_to_hz (htk=False): linear below 1 kHz, logarithmic above
    static double hz_to_mel(double f) {
        const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
        return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
    }
    static double mel_to_hz(double m) {
        const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
        return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
    }

    Frontend(const zvx_mel_config& c, int dev) : cfg(c), device(dev) {
        ZVX_REQUIRE(c.abi_version == ZVX_ABI_VERSION, "zvx_mel_config.abi_version mismatch");
        ZVX_REQUIRE(c.fft_size == FE_NFFT, "mel front-end: fft_size must be 1024 (every reference config)");
        ZVX_REQUIRE(c.win_length >= 2 && c.win_length <= c.fft_size, "mel front-end: 2 <= win_length <= fft_size");
        ZVX_REQUIRE(c.hop_size >= 1 && c.hop_size <= c.fft_size && (c.fft_size - c.hop_size) % 2 == 0,
                    "mel front-end: hop_size must be <= fft_size with (fft_size - hop_size) even");
        ZVX_REQUIRE(c.num_mels >= 1 && c.num_mels <= FE_MAXMEL, "mel front-end: 1 <= num_mels <= 256");
        ZVX_REQUIRE(c.sampling_rate > 0 && c.fmin >= 0 && c.fmax > c.fmin && c.fmax <= 0.5f * c.sampling_rate,
                    "mel front-end: need 0 <= fmin < fmax <= sampling_rate / 2");
        int ndev = 0;
        ZVX_CUDA_CHECK(cudaGetDeviceCount(&ndev));
        ZVX_REQUIRE(dev >= 0 && dev < ndev, "mel front-end: no such CUDA device (there is no CPU path)");
        ZVX_CUDA_CHECK(cudaSetDevice(dev));

        const double PI = 3.14159265358979323846;
        std::vector<float2> tw512(512), tw1024(513);
        for (int j = 0; j < 512; ++j) tw512[j] = make_float2((float)std::cos(2 * PI * j / 512), (float)-std::sin(2 * PI * j / 512));
        for (int k = 0; k <= 512; ++k) tw1024[k] = make_float2((float)std::cos(2 * PI * k / 1024), (float)-std::sin(2 * PI * k / 1024));
        // scipy.signal.get_window('hann', win_length, fftbins=True), centred in fft_size (librosa.util.pad_center)
        std::vector<float> win(FE_NFFT, 0.f);
        const int lpad = (c.fft_size - c.win_length) / 2;
        for (int n = 0; n < c.win_length; ++n) win[lpad + n] = (float)(0.5 - 0.5 * std::cos(2 * PI * n / c.win_length));
        // librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): Slaney scale, triangular, Slaney area normalisation
        const int nb = FE_NFFT / 2 + 1, nm = c.num_mels;
        std::vector<double> mel_f(nm + 2);
        const double m_lo = hz_to_mel(c.fmin), m_hi = hz_to_mel(c.fmax);
        const double m_step = (m_hi - m_lo) / (nm + 1);                               // np.linspace arithmetic
        for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(i == nm + 1 ? m_hi : i * m_step + m_lo);
        std::vector<int> f_start(nm), f_count(nm), f_off(nm);
        std::vector<float> f_w;
        for (int i = 0; i < nm; ++i) {
            const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
            int first = -1, last = -2;
            std::vector<float> row(nb);
            for (int k = 0; k < nb; ++k) {
                const double fk = (double)k * c.sampling_rate / c.fft_size;
                const double lower = (fk - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
                const double upper = (mel_f[i + 2] - fk) / (mel_f[i + 2] - mel_f[i + 1]);
                const float tri = (float)std::max(0.0, std::min(lower, upper));      // stored as float32 ...
                row[k] = (float)((double)tri * enorm);                                // ... then scaled in double
                if (row[k] != 0.f) { if (first < 0) first = k; last = k; }
            }
            f_start[i] = first < 0 ? 0 : first;
            f_count[i] = first < 0 ? 0 : last - first + 1;
            f_off[i] = (int)f_w.size();
            for (int k = f_start[i]; k < f_start[i] + f_count[i]; ++k) f_w.push_back(row[k]);
        }
        p.tw512 = upload(tw512);
        p.tw1024 = upload(tw1024);
        p.window = upload(win);
        p.f_start = upload(f_start);
        p.f_count = upload(f_count);
        p.f_off = upload(f_off);
        p.f_w = upload(f_w);
        p.num_mels = nm;
        p.hop = c.hop_size;
        p.pad = (c.fft_size - c.hop_size) / 2;
        p.clip = c.clip_val > 0.f ? c.clip_val : 1e-5f;
    }

    ~Frontend() {
        for (void* d : owned) cudaFree(d);
        cudaFree(rms);
        cudaFree(rmax);
    }

    void mel(const float* wav, int B, long long n_stride, const long long* wav_start, const long long* wav_len,
             int n_frames, float* mel_out, float* energy, cudaStream_t st) {
        ZVX_REQUIRE(B >= 1 && B <= 65535 && n_stride >= 1 && n_frames >= 0, "zvx_mel_spectrogram: bad sizes");
        if (n_frames == 0) return;                    // input shorter than one frame: nothing to write
        ZVX_REQUIRE(wav && mel_out, "zvx_mel_spectrogram: null pointer");
        ZVX_CUDA_CHECK(cudaSetDevice(device));
        dim3 grid(cdiv(n_frames, FE_FPC), B);
        mel_spectrogram_kernel<<<grid, FE_TPF * FE_FPC, 0, st>>>(wav, n_stride, wav_start, wav_len, n_frames, p, mel_out, energy);
        ZVX_POST_LAUNCH();
    }

    const ResampleTable& resample_table(int sr_in, int sr_out) {
        auto it = rs_tabs.find({sr_in, sr_out});
        if (it != rs_tabs.end()) return it->second;
        ZVX_REQUIRE(sr_in >= 1000 && sr_out >= 1000 && sr_in <= 768000 && sr_out <= 768000, "zvx_resample: sampling rate out of range");
        int a = sr_in, b = sr_out;
        while (b) { const int t = a % b; a = b; b = t; }
        ResampleTable r;
        r.up = sr_out / a; r.down = sr_in / a;
        ZVX_REQUIRE(r.up <= 4096, "zvx_resample: rate ratio needs more than 4096 filter phases");
        const double num_zeros = 64.0, beta = 14.769656459379492, rolloff = 0.9475937167399596;
        const double scale = std::min(1.0, (double)sr_out / sr_in);
        r.kh = (int)std::ceil(num_zeros / scale);
        r.taps = 2 * r.kh;
        std::vector<float> h((size_t)r.up * r.taps);
        const double i0b = bessel_i0(beta), pi = 3.14159265358979323846;
        for (int ph = 0; ph < r.up; ++ph)
            for (int j = 0; j < r.taps; ++j) {
                const double u = ((double)ph / r.up + r.kh - 1 - j) * scale;   // (t - k) * scale, k = floor(t) - kh + 1 + j
                double w = 0.0;
                if (std::fabs(u) < num_zeros) {
                    const double xs = pi * rolloff * u, sinc = std::fabs(xs) < 1e-12 ? 1.0 : std::sin(xs) / xs;
                    const double q = u / num_zeros;
                    w = scale * rolloff * sinc * bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - q * q))) / i0b;
                }
                h[(size_t)ph * r.taps + j] = (float)w;
            }
        r.tab = upload(h);
        return rs_tabs.emplace(std::make_pair(sr_in, sr_out), r).first->second;
    }

    void resample(const float* in, int B, long long n_in_stride, const long long* len_in, int sr_in, int sr_out, float* out,
                  long long n_out_stride, long long n_out, cudaStream_t st) {
        ZVX_REQUIRE(in && out, "zvx_resample: null pointer");
        ZVX_REQUIRE(B >= 1 && B <= 65535 && n_in_stride >= 1 && n_out >= 0 && n_out <= n_out_stride, "zvx_resample: bad sizes");
        ZVX_CUDA_CHECK(cudaSetDevice(device));
        if (n_out == 0) return;
        const ResampleTable& r = resample_table(sr_in, sr_out);
        resample_kernel<<<dim3(cdiv(n_out, 256), B), 256, 0, st>>>(in, n_in_stride, len_in, r.tab, r.up, r.down, r.taps, r.kh, out,
                                                                    n_out_stride, n_out);
        ZVX_POST_LAUNCH();
    }

    void trim(const float* wav, int B, long long n_stride, const long long* wav_len, float top_db, int frame_length,
              int hop, long long* start, long long* len, long long* host_out, cudaStream_t st) {
        ZVX_REQUIRE(wav && start && len, "zvx_trim_silence: null pointer");
        ZVX_REQUIRE(B >= 1 && B <= 65535 && n_stride >= 1, "zvx_trim_silence: bad sizes");
        ZVX_REQUIRE(frame_length >= 1 && hop >= 1, "zvx_trim_silence: frame_length and hop_length must be positive");
        ZVX_REQUIRE(n_stride / hop < (1LL << 30), "zvx_trim_silence: waveform too long");
        ZVX_CUDA_CHECK(cudaSetDevice(device));
        const int nfm = (int)(1 + n_stride / hop);
        const size_t need = (size_t)B * nfm;
        if (need > rms_cap) {          // grows only when a longer prompt arrives; steady state allocates nothing
            ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(rms);
            rms = nullptr; rms_cap = 0;
            ZVX_CUDA_CHECK(cudaMalloc(&rms, need * sizeof(float)));
            rms_cap = need;
        }
        if ((size_t)B > rmax_cap) {
            ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(rmax);
            rmax = nullptr; rmax_cap = 0;
            ZVX_CUDA_CHECK(cudaMalloc(&rmax, B * sizeof(unsigned)));
            rmax_cap = B;
        }
        ZVX_CUDA_CHECK(cudaMemsetAsync(rmax, 0, B * sizeof(unsigned), st));
        frame_rms_kernel<<<dim3(cdiv(nfm, 8), B), 256, 0, st>>>(wav, n_stride, wav_len, nfm, frame_length, hop, rms, rmax);
        ZVX_POST_LAUNCH();
        trim_bounds_kernel<<<B, 256, 0, st>>>(rms, rmax, wav_len, n_stride, nfm, hop, top_db, start, len);
        ZVX_POST_LAUNCH();
        if (host_out) {
            ZVX_CUDA_CHECK(cudaMemcpyAsync(host_out, start, B * sizeof(long long), cudaMemcpyDeviceToHost, st));
            ZVX_CUDA_CHECK(cudaMemcpyAsync(host_out + B, len, B * sizeof(long long), cudaMemcpyDeviceToHost, st));
            ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
        }
    }
};

}  // namespace zvx

struct zvx_frontend {
    std::unique_ptr<zvx::Frontend> fe;
};

#define ZVX_FE_GUARD(h, ...)                                                   \
    if (!(h) || !(h)->fe) return -1;                                           \
    try {                                                                      \
        __VA_ARGS__;                                                           \
        return 0;                                                              \
    } catch (const std::exception& e) {                                        \
        (h)->fe->err = e.what();                                               \
        return 1;                                                              \
    } catch (...) {                                                            \
        (h)->fe->err = "unknown error";                                        \
        return 2;                                                              \
    }

extern "C" {

int zvx_frontend_create(const zvx_mel_config* cfg, int device, zvx_frontend** out) {
    if (!cfg || !out) {
        zvx::g_fe_create_error = "zvx_frontend_create: null argument";
        return -1;
    }
    try {
        auto* h = new zvx_frontend();
        h->fe.reset(new zvx::Frontend(*cfg, device));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        zvx::g_fe_create_error = e.what();
        return 1;
    } catch (...) {
        zvx::g_fe_create_error = "unknown error";
        return 2;
    }
}

void zvx_frontend_destroy(zvx_frontend* h) { delete h; }

const char* zvx_frontend_last_error(const zvx_frontend* h) {
    if (!h || !h->fe) return zvx::g_fe_create_error.c_str();
    return h->fe->err.c_str();
}

int64_t zvx_mel_num_frames(const zvx_frontend* h, int64_t n_samples) {
    if (!h || !h->fe) return -1;
    return zvx::fe_num_frames(n_samples, h->fe->p.hop, h->fe->p.pad);
}

int zvx_trim_silence(zvx_frontend* h, const float* wav, int B, int64_t n_stride, const int64_t* wav_len, float top_db,
                     int frame_length, int hop_length, int64_t* start, int64_t* len, int64_t* start_len_host,
                     void* stream) {
    ZVX_FE_GUARD(h, h->fe->trim(wav, B, n_stride, (const long long*)wav_len, top_db, frame_length, hop_length,
                                (long long*)start, (long long*)len, (long long*)start_len_host, (cudaStream_t)stream));
}

int64_t zvx_resample_num_samples(int64_t n_in, int sr_in, int sr_out) {
    if (n_in <= 0 || sr_in <= 0 || sr_out <= 0) return 0;
    return (n_in * (int64_t)sr_out + sr_in - 1) / sr_in;   // ceil(n * ratio), as librosa.resample sizes its output
}

int zvx_resample(zvx_frontend* h, const float* wav_in, int B, int64_t n_in_stride, const int64_t* len_in, int sr_in, int sr_out,
                 float* wav_out, int64_t n_out_stride, int64_t n_out, void* stream) {
    ZVX_FE_GUARD(h, h->fe->resample(wav_in, B, n_in_stride, (const long long*)len_in, sr_in, sr_out, wav_out, n_out_stride, n_out,
                                    (cudaStream_t)stream));
}

int zvx_mel_spectrogram(zvx_frontend* h, const float* wav, int B, int64_t n_stride, const int64_t* wav_start,
                        const int64_t* wav_len, int n_frames, float* mel_BTC, float* energy, void* stream) {
    ZVX_FE_GUARD(h, h->fe->mel(wav, B, n_stride, (const long long*)wav_start, (const long long*)wav_len, n_frames,
                               mel_BTC, energy, (cudaStream_t)stream));
}

}  // extern "C"
