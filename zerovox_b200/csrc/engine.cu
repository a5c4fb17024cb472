// zerovox_b200 engine: weight packing, workspace, stage orchestration and the C ABI (include/zerovox_b200.h).
//
// Data layout in HBM
//   acoustic model / speaker net : channel-last  ([B, T, C] / [B, H, W, C]) so that every Linear, Conv1d and Conv2d
//                                  is one (implicit) GEMM with K contiguous;
//   vocoder                      : channel-first ([B, C, T], time contiguous), the reference's own layout;
//   weights                      : packed once at zvx_finalize_weights (tap-major [tap][N][K] for convs, fused QKV,
//                                  stacked SCLN affine, eval-BatchNorm folded to scale/shift).
#include "../../include/zerovox_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "gemm_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace zvx {

namespace {

std::string g_create_error;

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

// Growable block workspace: pointers are stable within a call; the same call sequence re-uses the same blocks, so
// steady state performs no cudaMalloc.  A request that fits no remaining block first RELEASES every block behind the cursor
// (they are too small for this call sequence and would otherwise be stranded until zvx_destroy: device memory of a
// long-running server with growing shapes stays bounded by the live high-water mark) and then allocates one new block.
class Workspace {
  public:
    ~Workspace() { release(); }
    void reset() { cur_ = 0; off_ = 0; }
    template <typename T>
    T* get(long long n) {
        size_t bytes = (size_t)round_up(std::max<long long>(n, 1) * (long long)sizeof(T), 256);
        while (cur_ < blocks_.size() && off_ + bytes > blocks_[cur_].size) {
            if (off_ == 0) {
                // an EMPTY block that is too small: nothing of this call lives in it -> free it instead of skipping it
                ZVX_CUDA_CHECK(cudaDeviceSynchronize());   // earlier calls' kernels may still read it (rare path: shapes grew)
                cudaFree(blocks_[cur_].p);
                blocks_.erase(blocks_.begin() + (long)cur_);
                continue;
            }
            ++cur_;
            off_ = 0;
        }
        if (cur_ == blocks_.size()) {
            Block b;
            b.size = std::max(bytes, (size_t)256 << 20);
            ZVX_CUDA_CHECK(cudaMalloc(&b.p, b.size));
            blocks_.push_back(b);
            off_ = 0;
        }
        T* p = reinterpret_cast<T*>(static_cast<char*>(blocks_[cur_].p) + off_);
        off_ += bytes;
        return p;
    }
    int64_t bytes() const {
        int64_t t = 0;
        for (auto& b : blocks_) t += (int64_t)b.size;
        return t;
    }
    void release() {
        for (auto& b : blocks_) cudaFree(b.p);
        blocks_.clear();
        reset();
    }

  private:
    struct Block { void* p = nullptr; size_t size = 0; };
    std::vector<Block> blocks_;
    size_t cur_ = 0, off_ = 0;
};

struct FFTLayer {
    float *wqkv = nullptr, *bqkv = nullptr, *wfc = nullptr, *bfc = nullptr;
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;  // LayerNorm variant
};

struct VarPredictor {
    float *wc1 = nullptr, *bc1 = nullptr, *ln1_g = nullptr, *ln1_b = nullptr;
    float *wc2 = nullptr, *bc2 = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    float *wlin = nullptr, *blin = nullptr;
};

struct SEBlock {
    int inpl = 0, planes = 0, stride = 1, red = 0;
    float *w1 = nullptr, *bn1_s = nullptr, *bn1_b = nullptr;
    float *w2 = nullptr, *bn2_s = nullptr, *bn2_b = nullptr;
    float *se_w1 = nullptr, *se_b1 = nullptr, *se_w2 = nullptr, *se_b2 = nullptr;
    float *wd = nullptr, *bnd_s = nullptr, *bnd_b = nullptr;  // downsample (nullable)
};

struct STConv { float* w = nullptr; float* b = nullptr; int cin = 0, cout = 0, k = 1; };   // weight-norm folded, tap-major
struct STBlock {                  // ResBlk1d (affine InstanceNorm) or AdainResBlk1d (styletts.py:11-69, 95-139)
    STConv c1, c2, sc;            // sc.w == nullptr: identity shortcut
    float *n1_g = nullptr, *n1_b = nullptr, *n2_g = nullptr, *n2_b = nullptr;   // affine IN weight / bias
    float *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;   // AdaIN fc (2C x style)
    int cin = 0, cout = 0, cmid = 0;   // cmid: channels between conv1 and conv2
};

struct HGConv { float* w = nullptr; float* b = nullptr; int cin = 0, cout = 0, k = 1, dil = 1; };

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline numbers).
// Classes: see ZVX_PROF_* in the header.
struct Profiler {
    struct Rec { int cls; cudaEvent_t a, b; double flops, bytes; };
    bool on = false;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t ev() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        ZVX_CUDA_CHECK(cudaEventCreate(&e));
        return e;
    }
    void begin(int cls, double flops, double bytes, cudaStream_t st) {
        if (!on) return;
        Rec r{cls, ev(), ev(), flops, bytes};
        ZVX_CUDA_CHECK(cudaEventRecord(r.a, st));
        recs.push_back(r);
    }
    void end(cudaStream_t st) {
        if (!on) return;
        ZVX_CUDA_CHECK(cudaEventRecord(recs.back().b, st));
    }
    // Synchronises; returns totals for one class and recycles nothing (call reset() afterwards).
    void read(int cls, double* ms, int64_t* launches, double* flops, double* bytes) {
        *ms = 0; *launches = 0; *flops = 0; *bytes = 0;
        for (auto& r : recs) {
            if (r.cls != cls) continue;
            ZVX_CUDA_CHECK(cudaEventSynchronize(r.b));
            float t = 0.f;
            ZVX_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
            *ms += t; *launches += 1; *flops += r.flops; *bytes += r.bytes;
        }
    }
    void reset() {
        for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
        recs.clear();
    }
    ~Profiler() {
        reset();
        for (auto e : pool) cudaEventDestroy(e);
    }
};

}  // namespace

class Engine {
  public:
    Engine(const zvx_config& c, int device) : cfg(c), dev(device) {
        ZVX_REQUIRE(c.abi_version == ZVX_ABI_VERSION, "zvx_config.abi_version mismatch");
        int n = 0;
        ZVX_CUDA_CHECK(cudaGetDeviceCount(&n));
        ZVX_REQUIRE(n > 0 && device >= 0 && device < n, "no usable CUDA device (this engine has no CPU path)");
        ZVX_CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop{};
        ZVX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        ZVX_REQUIRE(prop.major == 10, "zerovox_b200 is built for sm_100a (Blackwell B200) only");
        num_sms = prop.multiProcessorCount;
        H = c.emb_dim + c.punct_emb_dim;
        ZVX_REQUIRE(H % 4 == 0 && H <= 1024, "hidden size must be a multiple of 4 and <= 1024");
        ZVX_REQUIRE(c.enc_heads > 0 && H % c.enc_heads == 0 && (H / c.enc_heads) % 4 == 0, "bad encoder head split");
        ZVX_REQUIRE(c.decoder_kind == 0 || c.decoder_kind == 1, "decoder_kind must be 0 (fastspeech2) or 1 (styletts)");
        ZVX_REQUIRE(c.dec_heads > 0 && H % c.dec_heads == 0 && (H / c.dec_heads) % 4 == 0, "bad decoder head split");
        ZVX_REQUIRE(c.vp_filter_size % 4 == 0 && c.conv_filter_size % 4 == 0, "filter sizes must be multiples of 4");
        ZVX_REQUIRE(c.n_mels % 8 == 0, "n_mels must be a multiple of 8");
        ZVX_REQUIRE(c.hg_resblock == 1 || c.hg_resblock == 2, "hg_resblock must be 1 or 2");
        ZVX_REQUIRE(c.hg_num_upsamples >= 1 && c.hg_num_upsamples <= ZVX_MAX_UPSAMPLES, "hg_num_upsamples");
        ZVX_REQUIRE(c.hg_num_kernels >= 1 && c.hg_num_kernels <= ZVX_MAX_RESBLOCK_KERNELS, "hg_num_kernels");
        ZVX_REQUIRE(c.hg_num_dilations >= 1 && c.hg_num_dilations <= ZVX_MAX_DILATIONS, "hg_num_dilations");
        long long hop = 1;
        for (int i = 0; i < c.hg_num_upsamples; ++i) hop *= c.hg_upsample_rates[i];
        ZVX_REQUIRE(hop == c.hop_length, "product of upsample rates must equal hop_length");
        ZVX_CUDA_CHECK(cudaMallocHost(&pinned_len, sizeof(int64_t) * kMaxBatch));
    }

    ~Engine() {
        cudaSetDevice(dev);
        if (side_stream) { cudaStreamSynchronize(side_stream); cudaStreamDestroy(side_stream); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_vp_x) cudaEventDestroy(ev_vp_x);
        if (ev_vp_done) cudaEventDestroy(ev_vp_done);
        for (void* p : owned) cudaFree(p);
        if (pinned_len) cudaFreeHost(pinned_len);
    }

    // ------------------------------------------------------------------------------------------ weights
    int set_weight(const char* key, const void* data, const int64_t* shape, int ndim) {
        ZVX_REQUIRE(key && data && ndim >= 0 && ndim <= 8, "zvx_set_weight: bad arguments");
        ZVX_CUDA_CHECK(cudaSetDevice(dev));
        HostTensor t;
        t.shape.assign(shape, shape + ndim);
        const int64_t n = t.numel();
        t.data.resize((size_t)n);
        cudaPointerAttributes attr{};
        cudaError_t e = cudaPointerGetAttributes(&attr, data);
        if (e != cudaSuccess) {
            cudaGetLastError();
            attr.type = cudaMemoryTypeUnregistered;
        }
        if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
            ZVX_CUDA_CHECK(cudaMemcpy(t.data.data(), data, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
        } else {
            std::memcpy(t.data.data(), data, (size_t)n * sizeof(float));
        }
        raw[key] = std::move(t);
        finalized = false;
        return 0;
    }

    const HostTensor& W(const std::string& key) {
        auto it = raw.find(key);
        if (it == raw.end()) throw Error("missing weight: " + key);
        return it->second;
    }
    const HostTensor& W(const std::string& key, std::initializer_list<int64_t> shape) {
        const HostTensor& t = W(key);
        if (t.shape != std::vector<int64_t>(shape)) throw Error("weight has unexpected shape: " + key);
        return t;
    }

    float* upload(const std::vector<float>& v) {
        float* p = nullptr;
        ZVX_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(v.size(), 4) * sizeof(float)));
        owned.push_back(p);
        if (!v.empty()) ZVX_CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return p;
    }
    float* upload(const HostTensor& t) { return upload(t.data); }

    // Conv weight [N][C][k] -> tap-major [k][N][C]
    static std::vector<float> tap_major(const HostTensor& t) {
        const int64_t N = t.shape[0], C = t.shape[1];
        int64_t k = 1;
        for (size_t i = 2; i < t.shape.size(); ++i) k *= t.shape[i];
        std::vector<float> o((size_t)(N * C * k));
        for (int64_t n = 0; n < N; ++n)
            for (int64_t c = 0; c < C; ++c)
                for (int64_t j = 0; j < k; ++j) o[(size_t)((j * N + n) * C + c)] = t.data[(size_t)((n * C + c) * k + j)];
        return o;
    }

    void bn_fold(const std::string& p, int C, float** scale, float** shift) {
        const HostTensor &w = W(p + ".weight", {C}), &b = W(p + ".bias", {C});
        const HostTensor &m = W(p + ".running_mean", {C}), &v = W(p + ".running_var", {C});
        std::vector<float> s((size_t)C), sh((size_t)C);
        for (int i = 0; i < C; ++i) {
            const float inv = 1.0f / std::sqrt(v.data[i] + 1e-5f);
            s[i] = w.data[i] * inv;
            sh[i] = b.data[i] - m.data[i] * s[i];
        }
        *scale = upload(s);
        *shift = upload(sh);
    }

    void pack_fft(const std::string& prefix, int n_layers, bool scln, std::vector<FFTLayer>& out,
                  std::vector<float>* scln_stack, bool exact) {
        const int DI = cfg.conv_filter_size, k1 = cfg.conv_kernel_size[0], k2 = cfg.conv_kernel_size[1];
        out.resize((size_t)n_layers);
        for (int i = 0; i < n_layers; ++i) {
            const std::string p = prefix + ".layer_stack." + std::to_string(i);
            FFTLayer& L = out[(size_t)i];
            std::vector<float> wqkv, bqkv;
            for (const char* nm : {"w_qs", "w_ks", "w_vs"}) {
                const HostTensor& w = W(p + ".slf_attn." + nm + ".weight", {H, H});
                const HostTensor& b = W(p + ".slf_attn." + nm + ".bias", {H});
                wqkv.insert(wqkv.end(), w.data.begin(), w.data.end());
                bqkv.insert(bqkv.end(), b.data.begin(), b.data.end());
            }
            L.wqkv = upload(wqkv);
            L.bqkv = upload(bqkv);
            L.wfc = upload(W(p + ".slf_attn.fc.weight", {H, H}));
            L.bfc = upload(W(p + ".slf_attn.fc.bias", {H}));
            L.w1 = upload(tap_major(W(p + ".pos_ffn.w_1.weight", {DI, H, k1})));
            L.b1 = upload(W(p + ".pos_ffn.w_1.bias", {DI}));
            L.w2 = upload(tap_major(W(p + ".pos_ffn.w_2.weight", {H, DI, k2})));
            L.b2 = upload(W(p + ".pos_ffn.w_2.bias", {H}));
            if (exact) {
                register_lo(L.wqkv, (size_t)3 * H * H);
                register_lo(L.wfc, (size_t)H * H);
                register_lo(L.w1, (size_t)DI * H * k1);
                register_lo(L.w2, (size_t)H * DI * k2);
            }
            if (scln) {
                for (const char* sub : {"slf_attn", "pos_ffn"}) {
                    const HostTensor& a = W(p + "." + sub + ".layer_norm.affine_layer.linear.weight", {2 * H, H});
                    scln_stack->insert(scln_stack->end(), a.data.begin(), a.data.end());
                }
            } else {
                L.ln1_g = upload(W(p + ".slf_attn.layer_norm.weight", {H}));
                L.ln1_b = upload(W(p + ".slf_attn.layer_norm.bias", {H}));
                L.ln2_g = upload(W(p + ".pos_ffn.layer_norm.weight", {H}));
                L.ln2_b = upload(W(p + ".pos_ffn.layer_norm.bias", {H}));
            }
        }
    }

    static std::vector<float> sinusoid_table(int rows, int d, int first_row = 0) {
        // fs2.py:17-37, float64 arithmetic then cast
        std::vector<float> t((size_t)rows * d);
        for (int r = 0; r < rows; ++r)
            for (int j = 0; j < d; ++j) {
                const double ang = (double)(first_row + r) / std::pow(10000.0, 2.0 * (double)(j / 2) / (double)d);
                t[(size_t)r * d + j] = (float)((j & 1) ? std::cos(ang) : std::sin(ang));
            }
        return t;
    }

    // Position table: rows from the state_dict parameter, extended by the formula when longer inputs arrive
    // (the reference recomputes the table in eval when L > max_len: fs2.py:287-294, 383-388).
    const float* pos_table(bool decoder, int rows) {
        auto& cache = decoder ? dec_pos : enc_pos;
        auto& cache_rows = decoder ? dec_pos_rows : enc_pos_rows;
        const int param_rows = (decoder ? cfg.max_mel_len : cfg.max_txt_len) + 1;
        // reference semantics: parameter rows if L <= max_len else a freshly computed table
        const bool need_formula = rows > param_rows - 1;
        if (!need_formula) return decoder ? dec_pos_param : enc_pos_param;
        if (rows > cache_rows) {
            std::vector<float> t = sinusoid_table(rows, H);
            cache = upload(t);  // old tables stay owned until destroy (rare growth)
            cache_rows = rows;
        }
        return cache;
    }

    // Each section (encoder + variance adaptor, decoder, speaker net, vocoder) is packed independently so that a
    // stand-alone Generator / ResNetSE34V2 (get_meldec, model.py:86-118; synthesize.py:141) works without the rest.
    int finalize() {
        ZVX_CUDA_CHECK(cudaSetDevice(dev));
        for (void* p : owned) cudaFree(p);
        owned.clear();
        w_lo.clear();
        enc_pos = dec_pos = nullptr;
        enc_pos_rows = dec_pos_rows = 0;
        int ready = 0;
        auto section = [&](int idx, void (Engine::*fn)()) {
            try {
                (this->*fn)();
                sec_ready[idx] = true;
                sec_err[idx].clear();
                ++ready;
            } catch (const std::exception& e) {
                sec_ready[idx] = false;
                sec_err[idx] = e.what();
            }
        };
        section(SEC_ENC, &Engine::pack_encoder);
        section(SEC_DEC, &Engine::pack_decoder);
        section(SEC_SPK, &Engine::pack_spknet);
        section(SEC_VOC, &Engine::pack_vocoder);
        ZVX_CUDA_CHECK(cudaDeviceSynchronize());   // weight uploads / low-part splits ran on the default stream
        finalized = true;
        if (ready == 0) throw Error("zvx_finalize_weights: no complete section; first problem: " + sec_err[SEC_ENC]);
        return 0;
    }

    void pack_encoder() {
        const int E = cfg.emb_dim, P = cfg.punct_emb_dim;
        // encoder
        const std::string e = "_phoneme_encoder._encoder";
        enc_pos_param = upload(W(e + ".position_enc", {1, cfg.max_txt_len + 1, H}));
        phon_emb = upload(W(e + ".src_word_emb.weight", {cfg.num_phones + 1, E}));
        punct_emb = upload(W(e + ".punct_embed.weight", {cfg.num_puncts + 1, P}));
        pack_fft(e, cfg.enc_layers, false, enc, nullptr, /*exact=*/true);
        // variance adaptor
        const std::string va = "_phoneme_encoder._variance_adaptor";
        const int F = cfg.vp_filter_size, K = cfg.vp_kernel_size;
        int vi = 0;
        for (const char* nm : {"duration", "pitch", "energy"}) {
            const std::string p = va + "." + nm + "_predictor";
            VarPredictor& v = vp[vi++];
            v.wc1 = upload(tap_major(W(p + ".conv_layer.conv1d_1.conv.weight", {F, H, K})));
            v.bc1 = upload(W(p + ".conv_layer.conv1d_1.conv.bias", {F}));
            v.ln1_g = upload(W(p + ".conv_layer.layer_norm_1.weight", {F}));
            v.ln1_b = upload(W(p + ".conv_layer.layer_norm_1.bias", {F}));
            v.wc2 = upload(tap_major(W(p + ".conv_layer.conv1d_2.conv.weight", {F, F, K})));
            v.bc2 = upload(W(p + ".conv_layer.conv1d_2.conv.bias", {F}));
            v.ln2_g = upload(W(p + ".conv_layer.layer_norm_2.weight", {F}));
            v.ln2_b = upload(W(p + ".conv_layer.layer_norm_2.bias", {F}));
            register_lo(v.wc1, (size_t)F * H * K);
            register_lo(v.wc2, (size_t)F * F * K);
            v.wlin = upload(W(p + ".linear_layer.weight", {1, F}));
            v.blin = upload(W(p + ".linear_layer.bias", {1}));
        }
        pitch_emb = upload(W(va + ".pitch_embedding.weight", {cfg.ve_n_bins, H}));
        energy_emb = upload(W(va + ".energy_embedding.weight", {cfg.ve_n_bins, H}));
    }

    // weight_norm(dim=0): w = g * v / ||v|| per output channel (styletts.py:28-34), then tap-major [k][N][C]
    STConv pack_wn_conv(const std::string& p, int cout, int cin, int k, bool bias) {
        const HostTensor& v = W(p + ".weight_v", {cout, cin, k});
        const HostTensor& g = W(p + ".weight_g", {cout, 1, 1});
        HostTensor t;
        t.shape = v.shape;
        t.data.resize(v.data.size());
        const size_t per = (size_t)cin * k;
        for (int n = 0; n < cout; ++n) {
            double ss = 0.0;
            for (size_t i = 0; i < per; ++i) ss += (double)v.data[n * per + i] * v.data[n * per + i];
            const float scale = g.data[(size_t)n] / (float)std::sqrt(ss);
            for (size_t i = 0; i < per; ++i) t.data[n * per + i] = v.data[n * per + i] * scale;
        }
        STConv c;
        c.w = upload(tap_major(t));
        c.b = bias ? upload(W(p + ".bias", {cout})) : nullptr;
        c.cin = cin; c.cout = cout; c.k = k;
        return c;
    }

    void pack_decoder_styletts() {
        const std::string d = "_mel_decoder";
        const int RD = 64, BN = 2 * H;   // residual_dim, bottleneck (model.py:238-242; styletts.py:148)
        st_enc.clear();
        st_dec.clear();
        const int enc_dims[2][2] = {{H, BN}, {BN, BN}};
        for (int i = 0; i < 2; ++i) {
            const std::string p = d + ".encode." + std::to_string(i);
            const int ci = enc_dims[i][0], co = enc_dims[i][1];
            STBlock b;
            b.cin = ci; b.cout = co; b.cmid = ci;
            b.c1 = pack_wn_conv(p + ".conv1", ci, ci, 3, true);
            b.c2 = pack_wn_conv(p + ".conv2", co, ci, 3, true);
            b.n1_g = upload(W(p + ".norm1.weight", {ci})); b.n1_b = upload(W(p + ".norm1.bias", {ci}));
            b.n2_g = upload(W(p + ".norm2.weight", {ci})); b.n2_b = upload(W(p + ".norm2.bias", {ci}));
            if (ci != co) b.sc = pack_wn_conv(p + ".conv1x1", co, ci, 1, false);
            st_enc.push_back(b);
        }
        const int dec_dims[5][2] = {{BN + RD, BN}, {BN + RD, BN}, {BN + RD, H}, {H, H}, {H, H}};
        for (int i = 0; i < 5; ++i) {
            const std::string p = d + ".decode." + std::to_string(i);
            const int ci = dec_dims[i][0], co = dec_dims[i][1];
            STBlock b;
            b.cin = ci; b.cout = co; b.cmid = co;
            b.c1 = pack_wn_conv(p + ".conv1", co, ci, 3, true);
            b.c2 = pack_wn_conv(p + ".conv2", co, co, 3, true);
            b.fc1_w = upload(W(p + ".norm1.fc.weight", {2 * ci, H})); b.fc1_b = upload(W(p + ".norm1.fc.bias", {2 * ci}));
            b.fc2_w = upload(W(p + ".norm2.fc.weight", {2 * co, H})); b.fc2_b = upload(W(p + ".norm2.fc.bias", {2 * co}));
            if (ci != co) b.sc = pack_wn_conv(p + ".conv1x1", co, ci, 1, false);
            st_dec.push_back(b);
        }
        st_asr = pack_wn_conv(d + ".asr_res.0", RD, H, 1, true);
        st_asr_g = upload(W(d + ".asr_res.1.weight", {RD}));
        st_asr_b = upload(W(d + ".asr_res.1.bias", {RD}));
        st_out = pack_wn_conv(d + ".to_out.0", cfg.n_mels, H, 1, true);
    }

    void pack_decoder() {
        if (cfg.decoder_kind == 1) return pack_decoder_styletts();
        const std::string d = "_mel_decoder";
        dec_pos_param = upload(W(d + ".position_enc", {1, cfg.max_mel_len + 1, H}));
        std::vector<float> scln_stack;
        pack_fft(d, cfg.dec_layers, cfg.dec_scln != 0, dec, &scln_stack, /*exact=*/false);
        if (cfg.dec_scln) scln_w = upload(scln_stack);
        mel_w = upload(W(d + ".mel_linear.weight", {cfg.n_mels, H}));
        mel_b = upload(W(d + ".mel_linear.bias", {cfg.n_mels}));
    }

    void pack_spknet() {
        const std::string s = "_spkemb";
        const int* nf = cfg.resnet_num_filters;
        {
            const HostTensor& w = W(s + ".conv1.weight", {nf[0], 1, 3, 3});
            std::vector<float> t((size_t)9 * nf[0]);
            for (int c = 0; c < nf[0]; ++c)
                for (int j = 0; j < 9; ++j) t[(size_t)j * nf[0] + c] = w.data[(size_t)c * 9 + j];
            stem_w = upload(t);
            stem_b = upload(W(s + ".conv1.bias", {nf[0]}));
            bn_fold(s + ".bn1", nf[0], &stem_s, &stem_sh);
        }
        se_blocks.clear();
        int inpl = nf[0];
        for (int li = 0; li < 4; ++li) {
            const int planes = nf[li];
            ZVX_REQUIRE(planes % 8 == 0, "resnet_num_filters must be multiples of 8");
            for (int bi = 0; bi < cfg.resnet_layers[li]; ++bi) {
                const std::string p = s + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
                SEBlock b;
                b.inpl = inpl;
                b.planes = planes;
                b.stride = (li > 0 && bi == 0) ? 2 : 1;
                b.red = planes / 8;
                b.w1 = upload(tap_major(W(p + ".conv1.weight", {planes, inpl, 3, 3})));
                bn_fold(p + ".bn1", planes, &b.bn1_s, &b.bn1_b);
                b.w2 = upload(tap_major(W(p + ".conv2.weight", {planes, planes, 3, 3})));
                bn_fold(p + ".bn2", planes, &b.bn2_s, &b.bn2_b);
                b.se_w1 = upload(W(p + ".se.fc.0.weight", {b.red, planes}));
                b.se_b1 = upload(W(p + ".se.fc.0.bias", {b.red}));
                b.se_w2 = upload(W(p + ".se.fc.2.weight", {planes, b.red}));
                b.se_b2 = upload(W(p + ".se.fc.2.bias", {planes}));
                if (b.stride != 1 || inpl != planes) {
                    b.wd = upload(W(p + ".downsample.0.weight", {planes, inpl, 1, 1}));
                    bn_fold(p + ".downsample.1", planes, &b.bnd_s, &b.bnd_b);
                }
                se_blocks.push_back(b);
                inpl = planes;
            }
        }
        spk_D = nf[3] * (cfg.n_mels / 8);
        att_w0 = upload(W(s + ".attention.0.weight", {128, spk_D, 1}));
        att_b0 = upload(W(s + ".attention.0.bias", {128}));
        bn_fold(s + ".attention.2", 128, &att_bn_s, &att_bn_b);
        att_w3 = upload(W(s + ".attention.3.weight", {spk_D, 128, 1}));
        att_b3 = upload(W(s + ".attention.3.bias", {spk_D}));
        const int out_dim = cfg.resnet_encoder_type == 1 ? 2 * spk_D : spk_D;
        spk_fc_w = upload(W(s + ".fc.weight", {H, out_dim}));
        spk_fc_b = upload(W(s + ".fc.bias", {H}));
    }

    void pack_vocoder() {
        // vocoder: conv weights [Cout][Cin][k] -> [Cin][k][Cout]; transposed convs [Cin][Cout][k] -> [Cin][k][Cout]
        auto pack_conv = [&](const std::string& key, int cout, int cin, int k, int dil) {
            const HostTensor& w = W("_meldec." + key + ".weight", {cout, cin, k});
            std::vector<float> t((size_t)cout * cin * k);
            for (int co = 0; co < cout; ++co)
                for (int ci = 0; ci < cin; ++ci)
                    for (int j = 0; j < k; ++j)
                        t[((size_t)ci * k + j) * cout + co] = w.data[((size_t)co * cin + ci) * k + j];
            HGConv c;
            c.w = upload(t);
            c.b = upload(W("_meldec." + key + ".bias", {cout}));
            c.cin = cin; c.cout = cout; c.k = k; c.dil = dil;
            return c;
        };
        const int C0 = cfg.hg_upsample_initial_channel;
        hg_pre = pack_conv("conv_pre", C0, cfg.n_mels, 7, 1);
        hg_ups.clear();
        hg_c1.clear();
        hg_c2.clear();
        int ch = C0;
        for (int i = 0; i < cfg.hg_num_upsamples; ++i) {
            const int cin = C0 >> i, cout = C0 >> (i + 1), k = cfg.hg_upsample_kernel_sizes[i];
            ZVX_REQUIRE(cout >= 1, "too many upsample stages for upsample_initial_channel");
            const HostTensor& w = W("_meldec.ups." + std::to_string(i) + ".weight", {cin, cout, k});
            std::vector<float> t((size_t)cin * cout * k);
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co)
                    for (int j = 0; j < k; ++j)
                        t[((size_t)ci * k + j) * cout + co] = w.data[((size_t)ci * cout + co) * k + j];
            HGConv c;
            c.w = upload(t);
            c.b = upload(W("_meldec.ups." + std::to_string(i) + ".bias", {cout}));
            c.cin = cin; c.cout = cout; c.k = k; c.dil = 1;
            hg_ups.push_back(c);
            ch = cout;
            for (int j = 0; j < cfg.hg_num_kernels; ++j) {
                const std::string p = "resblocks." + std::to_string(i * cfg.hg_num_kernels + j);
                const int rk = cfg.hg_resblock_kernel_sizes[j];
                for (int di = 0; di < cfg.hg_num_dilations; ++di) {
                    const int dl = cfg.hg_resblock_dilation_sizes[j][di];
                    if (cfg.hg_resblock == 1) {
                        hg_c1.push_back(pack_conv(p + ".convs1." + std::to_string(di), ch, ch, rk, dl));
                        hg_c2.push_back(pack_conv(p + ".convs2." + std::to_string(di), ch, ch, rk, 1));
                    } else {
                        hg_c1.push_back(pack_conv(p + ".convs." + std::to_string(di), ch, ch, rk, dl));
                    }
                }
            }
        }
        hg_post = pack_conv("conv_post", 1, ch, 7, 1);
        pack_vocoder_tc();
    }

    // Tensor-core (channel-last) image of the vocoder weights, used when tensor_core_policy != 0 and every stage has a
    // channel count one of the tcgen05 kernels covers: C in {8,16,32} -> fused resblock kernels (voc_poly.cu, voc_res.cu); C >= 64 (and a
    // multiple of 4) -> generic TMA-fed implicit GEMM (gemm_tc.cu).
    void pack_vocoder_tc() {
        hg_tc_ok = false;
        hg_stage_kind.clear(); hg_ups_tc_w.clear(); hg_ups_tc_b.clear(); hg_c1_tc.clear(); hg_c2_tc.clear();
        hg_c1_poly.clear(); hg_c2_poly.clear(); hg_c1_pair.clear(); hg_c2_pair.clear();
        const int C0 = cfg.hg_upsample_initial_channel, nk = cfg.hg_num_kernels, nd = cfg.hg_num_dilations;
        if (C0 % 4 != 0 || C0 < 8 || cfg.n_mels % 4 != 0) return;
        for (int i = 0; i < cfg.hg_num_upsamples; ++i) {
            const int ch = C0 >> (i + 1), u = cfg.hg_upsample_rates[i];
            if (cfg.hg_upsample_kernel_sizes[i] != 2 * u || (u & 1) || (C0 >> i) % 4 != 0 || (C0 >> i) < 8) return;
            int kind = 0;
            if (ch == 8 || ch == 16 || ch == 32) {
                kind = 1;
                for (int j = 0; j < nk; ++j)
                    if (!voc_resblock_supported(ch, cfg.hg_resblock_kernel_sizes[j], cfg.hg_resblock_dilation_sizes[j], nd,
                                                cfg.hg_resblock == 1))
                        kind = 0;
            } else if (ch >= 64 && ch % 4 == 0) {
                kind = 2;
            }
            if (kind == 0) return;
            hg_stage_kind.push_back(kind);
        }
        hg_pre_tc = upload(tap_major(W("_meldec.conv_pre.weight", {C0, cfg.n_mels, 7})));
        size_t ci = 0;
        for (int i = 0; i < cfg.hg_num_upsamples; ++i) {
            const int cin = C0 >> i, cout = C0 >> (i + 1), u = cfg.hg_upsample_rates[i], k = 2 * u;
            // ConvTranspose1d as a 2-tap GEMM over input positions q: out[q*u + r - p, co] = W[ci,co,r] x[q] + W[ci,co,r+u] x[q-1]
            // tap 0 pairs with x[q-1], tap 1 with x[q]; GEMM column n = r*cout + co.
            const HostTensor& w = W("_meldec.ups." + std::to_string(i) + ".weight", {cin, cout, k});
            const HostTensor& b = W("_meldec.ups." + std::to_string(i) + ".bias", {cout});
            std::vector<float> wg((size_t)2 * u * cout * cin), bg((size_t)u * cout);
            for (int r = 0; r < u; ++r)
                for (int co = 0; co < cout; ++co) {
                    bg[(size_t)r * cout + co] = b.data[(size_t)co];
                    for (int c = 0; c < cin; ++c) {
                        wg[((size_t)0 * u * cout + (size_t)r * cout + co) * cin + c] = w.data[((size_t)c * cout + co) * k + r + u];
                        wg[((size_t)1 * u * cout + (size_t)r * cout + co) * cin + c] = w.data[((size_t)c * cout + co) * k + r];
                    }
                }
            hg_ups_tc_w.push_back(upload(wg));
            hg_ups_tc_b.push_back(upload(bg));
            for (int j = 0; j < nk; ++j) {
                const std::string p = "resblocks." + std::to_string(i * nk + j);
                const int rk = cfg.hg_resblock_kernel_sizes[j];
                for (int di = 0; di < nd; ++di, ++ci) {
                    const int dl = cfg.hg_resblock_dilation_sizes[j][di];
                    const bool pair_ok = cfg.hg_resblock == 1 && hg_stage_kind[(size_t)i] == 1 &&
                                         voc_pair_supported(cout, rk, cfg.hg_resblock_dilation_sizes[j], nd);
                    auto pack = [&](const std::string& key, int dil, std::vector<float*>& poly, std::vector<float*>& pairw) {
                        const HostTensor& t = W("_meldec." + key + ".weight", {cout, cout, rk});
                        if (hg_stage_kind[(size_t)i] != 1) {
                            poly.push_back(nullptr);
                            pairw.push_back(nullptr);
                            return upload(tap_major(t));
                        }
                        pairw.push_back(pair_ok ? upload(voc_pair_pack_weight(t.data.data(), W("_meldec." + key + ".bias", {cout}).data.data(),
                                                                              cout, rk)) : nullptr);
                        poly.push_back(upload(voc_poly_pack_weight(t.data.data(), cout, rk, dil)));
                        return upload(voc_pack_weight(t.data.data(), cout, cout, rk));
                    };
                    if (cfg.hg_resblock == 1) {
                        hg_c1_tc.push_back(pack(p + ".convs1." + std::to_string(di), dl, hg_c1_poly, hg_c1_pair));
                        hg_c2_tc.push_back(pack(p + ".convs2." + std::to_string(di), 1, hg_c2_poly, hg_c2_pair));
                    } else {
                        hg_c1_tc.push_back(pack(p + ".convs." + std::to_string(di), dl, hg_c1_poly, hg_c1_pair));
                    }
                }
            }
        }
        {
            const int ch = hg_post.cin;
            if (ch % 4 != 0) return;
            const HostTensor& w = W("_meldec.conv_post.weight", {1, ch, 7});
            std::vector<float> t((size_t)7 * ch);
            for (int c = 0; c < ch; ++c)
                for (int j = 0; j < 7; ++j) t[(size_t)j * ch + c] = w.data[(size_t)c * 7 + j];
            hg_post_tc = upload(t);
        }
        hg_tc_ok = true;
    }

    // ------------------------------------------------------------------------------------------ helpers
    // Precision classes of a contraction.  P_EXACT: fp32-grade products (encoder, variance predictors — their outputs
    // are rounded to integers / buckets): 3xTF32 split on the tensor cores when policy != 0, else the fp32 FMA kernel.
    // P_TF32: TF32 tensor cores when policy != 0 (decoder, vocoder, speaker net).
    enum { P_EXACT = 0, P_TF32 = 1 };
    bool tc_enabled(int prec) const { return cfg.tensor_core_policy != 0 && (prec == P_TF32 || split_on); }

    // TF32-exact low part of a packed weight (3xTF32 split), created once at finalize.
    void register_lo(const float* w, size_t n) {
        if (n % 4 != 0 || cfg.tensor_core_policy == 0) return;
        float* lo = nullptr;
        ZVX_CUDA_CHECK(cudaMalloc(&lo, n * sizeof(float)));
        owned.push_back(lo);
        tf32_split_lo(w, lo, (long long)n, 0);
        w_lo[w] = lo;
    }
    const float* lo_of(const float* w) const {
        auto it = w_lo.find(w);
        return it == w_lo.end() ? nullptr : it->second;
    }

    static long long a_extent(const GemmArgs& a) {
        long long rows = a.M;
        if (a.mode == ROW_CONV1D) rows = (long long)(a.M / a.Lout) * a.Lin;
        else if (a.mode == ROW_CONV2D) rows = (long long)(a.M / (a.Ho * a.Wo)) * a.Hi * a.Wi;
        return (rows - 1) * a.lda + a.K;
    }

    void gemm(const GemmArgs& a, int prec, cudaStream_t st) {
        TcGemmArgs t;
        bool tc = tc_enabled(prec) && gemm_tc_from(a, &t);
        bool split = false;
        if (tc && prec == P_EXACT) {
            const float* wlo = lo_of(a.W);
            const long long n = a_extent(a);
            if (wlo && lo_buf && n <= lo_cap) {
                tf32_split_lo(a.A, lo_buf, n, st);
                t.A_lo = lo_buf;
                t.W_lo = wlo;
                split = true;
            } else {
                tc = false;
            }
        }
        const double flops = 2.0 * a.M * a.N * a.K * a.taps * a.nz;
        const double bytes = 4.0 * a.nz * ((double)a.M * a.K + (double)a.N * a.K * a.taps + (double)a.M * a.N);
        prof.begin(split ? ZVX_PROF_GEMM_TC3 : tc ? ZVX_PROF_GEMM_TC : ZVX_PROF_GEMM_FP32, flops, bytes, st);
        if (tc) gemm_tc(t, st); else gemm_simt(a, st);
        prof.end(st);
    }

    void gemm_tc_prof(const TcGemmArgs& t, cudaStream_t st) {
        const double pos = (double)t.IMG * t.Ho * t.Wo;
        prof.begin(t.A_lo ? ZVX_PROF_GEMM_TC3 : ZVX_PROF_GEMM_TC, t.flops(),
                   4.0 * (pos * t.K + pos * t.N + (double)t.N * t.K * t.ksx * t.ksy), st);
        gemm_tc(t, st);
        prof.end(st);
    }

    void vconv(const Conv1dArgs& c, cudaStream_t st) {
        prof.begin(ZVX_PROF_VOC_CONV, 2.0 * c.B * c.T * c.Cin * c.Cout * c.k,
                   4.0 * c.B * c.T * (c.Cin + c.Cout * (1 + (c.res ? 1 : 0) + (c.acc && !c.acc_init ? 1 : 0))), st);
        conv1d_cf(c, st);
        prof.end(st);
    }

    void linear(const float* x, int M, int K, const float* w, const float* b, int N, float* y, int tc,
                cudaStream_t st, const float* R = nullptr, int relu_first = 0, const float* scale = nullptr,
                const float* shift = nullptr) {
        GemmArgs a;
        a.A = x; a.lda = K; a.W = w; a.ldw = K; a.C = y; a.ldc = N; a.bias = b; a.M = M; a.N = N; a.K = K;
        a.R = R; a.ldr = N; a.relu_first = relu_first; a.scale = scale; a.shift = shift;
        gemm(a, tc, st);
    }

    void conv1d_cl(const float* x, int B, int L, int Cin, const float* w, const float* b, int Cout, int k, int pad,
                   float* y, int tc, cudaStream_t st, const float* R = nullptr, int relu_first = 0) {
        GemmArgs a;
        a.A = x; a.lda = Cin; a.W = w; a.ldw = Cin; a.w_tap_stride = (long long)Cout * Cin; a.C = y; a.ldc = Cout;
        a.bias = b; a.M = B * L; a.N = Cout; a.K = Cin; a.taps = k; a.R = R; a.ldr = Cout; a.relu_first = relu_first;
        if (k > 1) {
            a.mode = ROW_CONV1D; a.Lout = L; a.Lin = L; a.stride = 1; a.pad = pad; a.dil = 1;
        }
        gemm(a, tc, st);
    }

    struct FFTScratch {
        float *qkv = nullptr, *att = nullptr, *y = nullptr, *S = nullptr, *h1 = nullptr;
        float *lo_qkv = nullptr, *lo_S = nullptr;   // 3xTF32 split: low parts of the attention operands
        long long qkv_elems = 0, S_elems = 0;
        int Lq_max = 0, ldS = 0;
        AttnPlan plan;                              // fused attention: tile list of this mask (one per stack of FFT blocks)
        bool has_plan = false;
    };

    // Tile list of the fused attention kernel for a stack of FFT blocks that share `mask` (attn_fused.cu attn_plan): masked key
    // blocks and the query tiles of masked positions are left out — fft_block zero-fills masked rows after each layer norm.
    void plan_attention(FFTScratch& s, int B, int L, int n_head, const uint8_t* mask, cudaStream_t st) {
        AttnFusedArgs fa;
        fa.key_mask = mask; fa.mask_ld = L; fa.B = B; fa.L = L; fa.n_head = n_head; fa.dk = H / n_head; fa.H = H;
        fa.Lp = (int)round_up(L, 4); fa.variant = kAttentionVariant;
        if (!kFusedAttention || !mask || H % n_head) return;
        int* w = ws.get<int>((long long)(attn_plan_bytes(B, L, n_head) / sizeof(int)));
        attn_plan(fa, /*skip_masked_queries=*/true, w, s.plan, st);
        s.has_plan = true;
    }

    FFTScratch fft_scratch(int B, int L, int n_head, int prec) {
        FFTScratch s;
        const long long rows = (long long)B * L;
        const int DI = cfg.conv_filter_size;
        s.qkv_elems = rows * 3 * H + 4LL * B * H;   // + slack: tensor-core path keeps V transposed, rows padded to 4
        s.qkv = ws.get<float>(s.qkv_elems);
        s.att = ws.get<float>(rows * H);
        s.y = ws.get<float>(rows * H);
        // attention is chunked over query rows so that the score matrix stays below kScoreBytes
        s.ldS = (int)round_up(L, 4);
        const long long max_q = kScoreBytes / ((long long)B * n_head * s.ldS * (long long)sizeof(float));
        s.Lq_max = (int)std::max<long long>(1, std::min<long long>(L, max_q));
        s.S_elems = (long long)B * n_head * s.Lq_max * s.ldS;
        // the score matrix only exists on the unfused paths (3xTF32 attention, fused_attention = 0, shapes the fused kernel rejects)
        AttnFusedArgs fa;
        fa.B = B; fa.L = L; fa.n_head = n_head; fa.dk = H / n_head; fa.H = H; fa.Lp = (int)round_up(L, 4); fa.variant = kAttentionVariant;
        const bool fused = prec == P_TF32 && tc_enabled(prec) && kFusedAttention && H % n_head == 0 && (H / n_head) % 4 == 0 &&
                           (long long)B * L >= tc_min_rows() && attn_fused_supported(fa);
        s.S = fused ? nullptr : ws.get<float>(s.S_elems);
        if (prec == P_EXACT && tc_enabled(prec)) {
            s.lo_qkv = ws.get<float>(s.qkv_elems);
            s.lo_S = ws.get<float>(s.S_elems);
        }
        // position-wise feed-forward hidden: re-use the qkv buffer when it is large enough
        s.h1 = (3 * H >= DI) ? s.qkv : ws.get<float>(rows * DI);
        return s;
    }

    // FFTBlock.forward (fs2.py:221-230) in place on x [B, L, H].
    void fft_block(float* x, int B, int L, int n_head, const FFTLayer& ly, const uint8_t* mask, bool scln,
                   const float* gb1, const float* gb2, int gb_ld, int tc, const FFTScratch& sc, cudaStream_t st) {
        const int dk = H / n_head, DI = cfg.conv_filter_size;
        const int k1 = cfg.conv_kernel_size[0], k2 = cfg.conv_kernel_size[1];
        const long long rows = (long long)B * L;
        float *qkv = sc.qkv, *att = sc.att, *y = sc.y, *S = sc.S;
        const int ldS = sc.ldS, Lq_max = sc.Lq_max;
        const int nz = B * n_head;
        const float temperature = (float)std::pow((double)dk, 0.5);  // np.power(d_k, 0.5), fs2.py:122
        static const int tc_mask = env_int("ZVX_TC_MASK", 15);   // bit0: tensor-core attention (debug builds only)
        const bool split = (tc == P_EXACT);
        const float* wqkv_lo = split ? lo_of(ly.wqkv) : nullptr;
        const bool tc_attn = tc_enabled(tc) && (tc_mask & 1) && (dk % 4 == 0) && rows >= tc_min_rows() &&
                             (!split || (wqkv_lo && sc.lo_qkv && lo_buf && rows * H <= lo_cap));
        if (tc_attn) {
            // tcgen05 path: [Q|K] row-major [rows, 2H]; V written transposed per utterance, Vt[b][c][t] (row pitch Lp),
            // so that both attention contractions read K-major operands through TMA.  Query rows are processed in
            // chunks of Lq_max so that the score matrix stays inside the workspace budget (long-form inputs).
            const int Lp = (int)round_up(L, 4);
            float* qk = qkv;
            float* vt = qkv + rows * 2 * H;
            linear(x, (int)rows, H, ly.wqkv, ly.bqkv, 2 * H, qk, tc, st);
            TcGemmArgs v;
            v.A = x; v.K = H; v.Wi = v.Wo = L; v.Hi = v.Ho = B; v.a_sx = H; v.a_sy = (long long)L * H;
            v.W = ly.wqkv + (long long)2 * H * H; v.N = H; v.w_sn = H; v.bias = ly.bqkv + 2 * H;
            v.C = vt; v.c_sy = (long long)H * Lp; v.c_sx = 1; v.c_sn = Lp;
            if (split) {
                tf32_split_lo(x, lo_buf, rows * H, st);
                v.A_lo = lo_buf; v.W_lo = wqkv_lo + (long long)2 * H * H;
            }
            gemm_tc_prof(v, st);
            if (split) tf32_split_lo(qkv, sc.lo_qkv, rows * 2 * H + (long long)B * H * Lp, st);
            // TF32 policy: ONE kernel for S = QK^T, the masked softmax and PV — the score matrix never leaves the SM
            // (attn_fused.cu).  The 3xTF32 policy (encoder, T <= a few hundred) keeps the three-kernel path below.
            AttnFusedArgs fa;
            fa.qk = qk; fa.vt = vt; fa.out = att; fa.key_mask = mask; fa.mask_ld = L; fa.B = B; fa.L = L; fa.n_head = n_head;
            fa.dk = dk; fa.H = H; fa.Lp = Lp; fa.temperature = temperature; fa.variant = kAttentionVariant;
            fa.plan = sc.has_plan ? &sc.plan : nullptr;
            if (!split && kFusedAttention && attn_fused_supported(fa)) {
                prof.begin(ZVX_PROF_GEMM_TC, fa.flops(), fa.bytes(), st);
                attn_fused(fa, st);
                prof.end(st);
            } else {
            ZVX_REQUIRE(S, "fft_block: no score workspace was planned for the unfused attention path");
            // Utterances are independent, so the three attention kernels run over slices of the batch whose score
            // matrices (plus their Q / K / V rows) fit in L2: the softmax and the PV product then read what the kernel
            // before them wrote from L2 instead of DRAM (the whole-batch score tensor is 172 MB at configs[1], moved
            // four times).  ZVX_ATTN_SLICE_BYTES = 0 restores one slice.
            static const long long slice_bytes = env_ll("ZVX_ATTN_SLICE_BYTES", kAttnSliceBytes);
            int Bs = B;
            if (slice_bytes > 0 && !split) {
                const long long per_utt = (long long)n_head * std::min(Lq_max, L) * ldS * (long long)sizeof(float);
                Bs = (int)std::max<long long>(1, std::min<long long>(B, slice_bytes / per_utt));
                if (Bs * 2 > B) Bs = B;                       // a single uneven split gains nothing
                else Bs = cdiv(B, cdiv(B, Bs));               // even slices
            }
            for (int b0 = 0; b0 < B; b0 += Bs) {
            const int Bn = std::min(Bs, B - b0);
            const long long r0 = (long long)b0 * L;          // first row of the slice in the [rows, .] buffers
            const uint8_t* mask_s = mask ? mask + r0 : nullptr;
            for (int q0 = 0; q0 < L; q0 += Lq_max) {
                const int Lq = std::min(Lq_max, L - q0);
                TcGemmArgs sq;   // S[b,h,q,j] = <Q[b,q0+q,h,:], K[b,j,h,:]>
                sq.A = qk + (r0 + q0) * 2 * H; sq.K = dk; sq.Wi = sq.Wo = Lq; sq.Hi = sq.Ho = n_head; sq.IMG = Bn;
                sq.a_sx = 2 * H; sq.a_sy = dk; sq.a_simg = (long long)L * 2 * H;
                sq.W = qk + r0 * 2 * H + H; sq.N = L; sq.Z1 = n_head; sq.Z2 = Bn; sq.w_sn = 2 * H; sq.w_s1 = dk; sq.w_s2 = (long long)L * 2 * H;
                sq.b_batched = 1;
                sq.C = S; sq.c_simg = (long long)n_head * Lq * ldS; sq.c_sy = (long long)Lq * ldS; sq.c_sx = ldS; sq.c_sn = 1;
                if (split) { sq.A_lo = sc.lo_qkv + (r0 + q0) * 2 * H; sq.W_lo = sc.lo_qkv + r0 * 2 * H + H; }
                gemm_tc_prof(sq, st);
                attn_softmax(S, Bn * n_head, n_head, Lq, L, ldS, mask_s, L, temperature, st);
                if (split) tf32_split_lo(S, sc.lo_S, (long long)Bn * n_head * Lq * ldS, st);
                TcGemmArgs pv;   // att[b,q0+q,h*dk + c] = sum_j P[b,h,q,j] * Vt[b, h*dk + c, j]
                pv.A = S; pv.K = L; pv.Wi = pv.Wo = Lq; pv.Hi = pv.Ho = n_head; pv.IMG = Bn;
                pv.a_sx = ldS; pv.a_sy = (long long)Lq * ldS; pv.a_simg = (long long)n_head * Lq * ldS;
                pv.W = vt + (long long)b0 * H * Lp; pv.N = dk; pv.Z1 = n_head; pv.Z2 = Bn; pv.w_sn = Lp; pv.w_s1 = (long long)dk * Lp; pv.w_s2 = (long long)H * Lp;
                pv.b_batched = 1;
                pv.C = att + (r0 + q0) * H; pv.c_simg = (long long)L * H; pv.c_sy = dk; pv.c_sx = H; pv.c_sn = 1;
                if (split) { pv.A_lo = sc.lo_S; pv.W_lo = sc.lo_qkv + rows * 2 * H + (long long)b0 * H * Lp; }
                gemm_tc_prof(pv, st);
            }
            }
            }
        } else {
        ZVX_REQUIRE(S, "fft_block: no score workspace was planned for the unfused attention path");
        linear(x, (int)rows, H, ly.wqkv, ly.bqkv, 3 * H, qkv, tc, st);
        for (int q0 = 0; q0 < L; q0 += Lq_max) {
            const int Lq = std::min(Lq_max, L - q0);
            GemmArgs s;
            s.A = qkv + (long long)q0 * 3 * H; s.lda = 3 * H; s.W = qkv + H; s.ldw = 3 * H; s.C = S; s.ldc = ldS;
            s.M = Lq; s.N = L; s.K = dk; s.nz = nz; s.nzh = n_head;
            s.sA_b = (long long)L * 3 * H; s.sA_h = dk; s.sW_b = (long long)L * 3 * H; s.sW_h = dk;
            s.sC_b = (long long)n_head * Lq * ldS; s.sC_h = (long long)Lq * ldS;
            gemm(s, tc, st);
            attn_softmax(S, nz, n_head, Lq, L, ldS, mask, L, temperature, st);
            GemmArgs o;
            o.A = S; o.lda = ldS; o.W = qkv + 2 * H; o.ldw = 3 * H; o.b_kn = 1; o.C = att + (long long)q0 * H; o.ldc = H;
            o.M = Lq; o.N = dk; o.K = L; o.nz = nz; o.nzh = n_head;
            o.sA_b = (long long)n_head * Lq * ldS; o.sA_h = (long long)Lq * ldS;
            o.sW_b = (long long)L * 3 * H; o.sW_h = dk; o.sC_b = (long long)L * H; o.sC_h = dk;
            gemm(o, tc, st);
        }
        }
        // the residual adds of fs2.py:158-162 / 205-208 are fused into the normalisation kernel (same fp32 add)
        linear(att, (int)rows, H, ly.wfc, ly.bfc, H, y, tc, st);
        NormArgs n;
        n.x = y; n.res = x; n.out = x; n.rows = (int)rows; n.C = H; n.rows_per_batch = L; n.mask = mask;
        if (scln) { n.scln = 1; n.gb = gb1; n.gb_ld = gb_ld; n.eps = 1e-8f; }
        else { n.gamma = ly.ln1_g; n.beta = ly.ln1_b; n.eps = 1e-5f; }
        layer_norm(n, st);
        // position-wise feed-forward (fs2.py:196-209)
        float* h1 = sc.h1;
        conv1d_cl(x, B, L, H, ly.w1, ly.b1, DI, k1, (k1 - 1) / 2, h1, tc, st, nullptr, /*relu_first=*/1);
        conv1d_cl(h1, B, L, DI, ly.w2, ly.b2, H, k2, (k2 - 1) / 2, y, tc, st);
        n.x = y; n.res = x; n.out = x;
        if (scln) n.gb = gb2; else { n.gamma = ly.ln2_g; n.beta = ly.ln2_b; }
        layer_norm(n, st);
    }

    void variance_predictor(const float* x, int B, int T, const VarPredictor& v, const uint8_t* mask, float* out,
                            cudaStream_t st, cudaEvent_t x_consumed = nullptr) {
        const int F = cfg.vp_filter_size, K = cfg.vp_kernel_size;
        const long long rows = (long long)B * T;
        float* c1 = ws.get<float>(rows * F);
        float* c2 = ws.get<float>(rows * F);
        conv1d_cl(x, B, T, H, v.wc1, v.bc1, F, K, (K - 1) / 2, c1, P_EXACT, st, nullptr, 1);
        if (x_consumed) ZVX_CUDA_CHECK(cudaEventRecord(x_consumed, st));   // (the only read of x)
        NormArgs n;
        n.x = c1; n.out = c1; n.rows = (int)rows; n.C = F; n.gamma = v.ln1_g; n.beta = v.ln1_b; n.eps = 1e-5f;
        layer_norm(n, st);
        conv1d_cl(c1, B, T, F, v.wc2, v.bc2, F, K, 1, c2, P_EXACT, st, nullptr, 1);  // padding=1 (fs2.py:543)
        n.x = c2; n.out = nullptr; n.gamma = v.ln2_g; n.beta = v.ln2_b;
        n.dot_w = v.wlin; n.dot_b = v.blin; n.dot_out = out; n.mask = mask;
        layer_norm(n, st);
    }

    void check_ready(int sec) {
        ZVX_REQUIRE(finalized, "weights not finalized: call zvx_finalize_weights first");
        if (!sec_ready[sec]) throw Error("weights for this stage are incomplete: " + sec_err[sec]);
        ZVX_CUDA_CHECK(cudaSetDevice(dev));
    }

    // ------------------------------------------------------------------------------------------ stages
    int spkemb(const float* ref_mel, int B, int T, float* style, cudaStream_t st) {
        check_ready(SEC_SPK);
        ZVX_REQUIRE(B >= 0 && T >= 8, "zvx_spkemb: need T_ref >= 8 frames");
        ws_spk.reset();   // the speaker net owns a workspace: it may run next to the encoder (spkemb_encode)
        const int M = cfg.n_mels;
        const int* nf = cfg.resnet_num_filters;
        const int tc = P_TF32;
        float* x0 = ws_spk.get<float>((long long)B * M * T);
        instance_norm_time(ref_mel, B, T, M, x0, st);
        long long big = (long long)B * M * T * nf[0];
        float* buf[4];
        for (int i = 0; i < 4; ++i) buf[i] = ws_spk.get<float>(big);
        float* gate = ws_spk.get<float>((long long)B * 1024);
        int* ticket = ws_spk.get<int>(B);   // se_squeeze_excite: one ticket counter per utterance, self-resetting
        ZVX_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(int) * (size_t)B, st));
        float* x = buf[0];
        stem_conv3x3(x0, stem_w, stem_b, stem_s, stem_sh, B, M, T, nf[0], x, st);
        int Hh = M, Ww = T;
        int xi = 0;
        for (const SEBlock& b : se_blocks) {
            ZVX_REQUIRE(b.planes <= 1024, "resnet filters > 1024");
            const int Ho = (Hh - 1) / b.stride + 1, Wo = (Ww - 1) / b.stride + 1;
            float* t1 = buf[(xi + 1) & 3];
            float* t2 = buf[(xi + 2) & 3];
            float* rs = buf[(xi + 3) & 3];
            GemmArgs c;
            c.A = x; c.lda = b.inpl; c.W = b.w1; c.ldw = b.inpl; c.w_tap_stride = (long long)b.planes * b.inpl;
            c.C = t1; c.ldc = b.planes; c.M = B * Ho * Wo; c.N = b.planes; c.K = b.inpl; c.taps = 9;
            c.mode = ROW_CONV2D; c.Ho = Ho; c.Wo = Wo; c.Hi = Hh; c.Wi = Ww; c.ksize = 3; c.stride = b.stride; c.pad = 1;
            c.relu_first = 1; c.scale = b.bn1_s; c.shift = b.bn1_b;
            gemm(c, tc, st);
            GemmArgs c2;
            c2.A = t1; c2.lda = b.planes; c2.W = b.w2; c2.ldw = b.planes; c2.w_tap_stride = (long long)b.planes * b.planes;
            c2.C = t2; c2.ldc = b.planes; c2.M = B * Ho * Wo; c2.N = b.planes; c2.K = b.planes; c2.taps = 9;
            c2.mode = ROW_CONV2D; c2.Ho = Ho; c2.Wo = Wo; c2.Hi = Ho; c2.Wi = Wo; c2.ksize = 3; c2.stride = 1; c2.pad = 1;
            c2.scale = b.bn2_s; c2.shift = b.bn2_b;
            gemm(c2, tc, st);
            const int S = hw_mean_splits(B, Ho * Wo);
            float* pooled = ws_spk.get<float>((long long)B * S * b.planes);
            se_squeeze_excite(t2, B, Ho * Wo, b.planes, S, pooled, ticket, b.se_w1, b.se_b1, b.se_w2, b.se_b2, b.red, gate, st);
            const float* res = x;
            if (b.wd) {
                GemmArgs dn;
                dn.A = x; dn.lda = b.inpl; dn.W = b.wd; dn.ldw = b.inpl; dn.C = rs; dn.ldc = b.planes;
                dn.M = B * Ho * Wo; dn.N = b.planes; dn.K = b.inpl; dn.taps = 1;
                dn.mode = ROW_CONV2D; dn.Ho = Ho; dn.Wo = Wo; dn.Hi = Hh; dn.Wi = Ww; dn.ksize = 1; dn.stride = b.stride; dn.pad = 0;
                dn.scale = b.bnd_s; dn.shift = b.bnd_b;
                gemm(dn, tc, st);
                res = rs;
            }
            se_scale_add_relu(t2, gate, res, B, Ho * Wo, b.planes, t1, st);  // t1 is free again
            x = t1;
            xi = (xi + 1) & 3;
            Hh = Ho;
            Ww = Wo;
        }
        // attentive statistics pooling head (ResNetSE34V2.py:196-208)
        const int C3 = nf[3];
        ZVX_REQUIRE(Hh * C3 == spk_D, "speaker net: unexpected feature-map height");
        float* flat = ws_spk.get<float>((long long)B * Ww * spk_D);
        float* a1 = ws_spk.get<float>((long long)B * Ww * 128);
        float* lg = ws_spk.get<float>((long long)B * Ww * spk_D);
        const int asp = cfg.resnet_encoder_type == 1;
        float* stats = ws_spk.get<float>((long long)B * spk_D * 2);
        spk_flatten(x, B, Hh, Ww, C3, flat, st);
        linear(flat, B * Ww, spk_D, att_w0, att_b0, 128, a1, tc, st, nullptr, 1, att_bn_s, att_bn_b);
        linear(a1, B * Ww, 128, att_w3, att_b3, spk_D, lg, tc, st);
        attentive_pool(flat, lg, B, Ww, spk_D, asp, stats, st);
        linear(stats, B, asp ? 2 * spk_D : spk_D, spk_fc_w, spk_fc_b, H, style, P_EXACT, st);
        l2_normalize(style, B, H, st);
        return 0;
    }

    int encode(const int32_t* phoneme, const int32_t* puncts, const uint8_t* mask, const float* style,
               const int32_t* forced, int B, int T, float* pitch, float* energy, float* log_dur, int32_t* dur,
               int64_t* mel_len, float* xprime, int64_t* mel_len_host, int* L_max_out, cudaStream_t st,
               cudaEvent_t style_ready = nullptr) {
        check_ready(SEC_ENC);
        ZVX_REQUIRE(B >= 1 && T >= 1, "zvx_encode: empty batch");
        ZVX_REQUIRE(B <= kMaxBatch, "zvx_encode: batch too large");
        ws.reset();
        float* x = xprime;
        embed_posenc(phoneme, puncts, phon_emb, punct_emb, pos_table(false, T), B, T, cfg.emb_dim, cfg.punct_emb_dim, x, st);
        if (tc_enabled(P_EXACT)) {   // scratch for the low part of a split GEMM's A operand
            lo_cap = (long long)B * T * std::max(std::max(3 * H, cfg.conv_filter_size), cfg.vp_filter_size);
            lo_buf = ws.get<float>(lo_cap);
        }
        const FFTScratch sc = fft_scratch(B, T, cfg.enc_heads, P_EXACT);
        for (int i = 0; i < cfg.enc_layers; ++i)
            fft_block(x, B, T, cfg.enc_heads, enc[(size_t)i], mask, false, nullptr, nullptr, 0, P_EXACT, sc, st);
        if (style_ready) ZVX_CUDA_CHECK(cudaStreamWaitEvent(st, style_ready, 0));   // (spkemb_encode: produced on the side stream)
        add_batch_vector(x, style, B, T, H, st);  // all positions, padded ones too (fs2.py:740-741)
        if (style_ready && side_stream) {
            // The duration and the pitch predictor read the same x (fs2.py:667-671) and are chains of one-wave-or-less launches
            // (B*T / 128 x 2 tiles): the duration predictor runs on the side stream — idle again once the style vector is there —
            // with its own split scratch; the pitch embedding may only be added to x once the duration predictor's first conv
            // has read it.
            if (!ev_vp_x) {
                ZVX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_vp_x, cudaEventDisableTiming));
                ZVX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_vp_done, cudaEventDisableTiming));
            }
            ZVX_CUDA_CHECK(cudaEventRecord(ev_fork, st));
            ZVX_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
            float* lo_main = lo_buf;
            lo_buf = lo_main ? ws.get<float>(lo_cap) : nullptr;
            variance_predictor(x, B, T, vp[0], mask, log_dur, side_stream, ev_vp_x);
            ZVX_CUDA_CHECK(cudaEventRecord(ev_vp_done, side_stream));
            lo_buf = lo_main;
            variance_predictor(x, B, T, vp[1], mask, pitch, st);
            ZVX_CUDA_CHECK(cudaStreamWaitEvent(st, ev_vp_x, 0));
        } else {
            variance_predictor(x, B, T, vp[0], mask, log_dur, st);
            variance_predictor(x, B, T, vp[1], mask, pitch, st);
        }
        bucket_embed_add(x, pitch, pitch_emb, B * T, H, cfg.ve_n_bins, nullptr, st);
        variance_predictor(x, B, T, vp[2], mask, energy, st);
        bucket_embed_add(x, energy, energy_emb, B * T, H, cfg.ve_n_bins, nullptr, st);
        if (style_ready && side_stream) ZVX_CUDA_CHECK(cudaStreamWaitEvent(st, ev_vp_done, 0));
        duration_round(log_dur, forced, dur, B * T, st);
        int32_t* cum = ws.get<int32_t>((long long)B * T);
        duration_scan(dur, B, T, cum, mel_len, st);
        lo_buf = nullptr;
        lo_cap = 0;
        if (L_max_out) {
            ZVX_CUDA_CHECK(cudaMemcpyAsync(pinned_len, mel_len, sizeof(int64_t) * B, cudaMemcpyDeviceToHost, st));
            ZVX_CUDA_CHECK(cudaStreamSynchronize(st));  // the one inherent sync (the reference: B*T + 1)
            int64_t mx = 0;
            for (int b = 0; b < B; ++b) mx = std::max(mx, pinned_len[b]);
            ZVX_REQUIRE(mx < (1LL << 30), "zvx_encode: implausible mel length");
            *L_max_out = (int)mx;
            if (mel_len_host) std::memcpy(mel_len_host, pinned_len, sizeof(int64_t) * B);
        }
        return 0;
    }

    // ZeroVox.forward's first two module calls as ONE call (model.py:263-265): the speaker net and the encoder's FFT blocks do not
    // depend on each other — the style vector enters after the last encoder layer (fs2.py:740-741) — so the speaker net is enqueued
    // on an engine-owned side stream and the caller's stream waits for it right before that addition.  The encoder's kernels are
    // small (B*T rows: one wave or less of tiles, launch-latency class); next to the speaker net's they fill SMs that would idle.
    int spkemb_encode(const float* ref_mel, int T_ref, float* style, const int32_t* phoneme, const int32_t* puncts,
                      const uint8_t* mask, const int32_t* forced, int B, int T, float* pitch, float* energy, float* log_dur,
                      int32_t* dur, int64_t* mel_len, float* xprime, int64_t* mel_len_host, int* L_max_out, cudaStream_t st) {
        check_ready(SEC_SPK);
        check_ready(SEC_ENC);
        if (prof.on) {   // the per-launch profiler times kernels one at a time: keep everything in the caller's stream order
            spkemb(ref_mel, B, T_ref, style, st);
            return encode(phoneme, puncts, mask, style, forced, B, T, pitch, energy, log_dur, dur, mel_len, xprime, mel_len_host,
                          L_max_out, st);
        }
        if (!side_stream) {
            ZVX_CUDA_CHECK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
            ZVX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            ZVX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        ZVX_CUDA_CHECK(cudaEventRecord(ev_fork, st));                     // the inputs are ready in the caller's stream order
        ZVX_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
        spkemb(ref_mel, B, T_ref, style, side_stream);
        ZVX_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
        return encode(phoneme, puncts, mask, style, forced, B, T, pitch, energy, log_dur, dur, mel_len, xprime, mel_len_host,
                      L_max_out, st, ev_join);
    }

    int length_regulate(const float* xprime, const int32_t* dur, int B, int T, int frame0, int L_max, float* features,
                        int32_t* src_index, cudaStream_t st) {
        ZVX_CUDA_CHECK(cudaSetDevice(dev));
        ZVX_REQUIRE(B >= 1 && T >= 1 && L_max >= 0 && frame0 >= 0 && B <= 65535, "zvx_length_regulate: bad sizes");
        ws.reset();
        int32_t* cum = ws.get<int32_t>((long long)B * T);
        duration_scan(dur, B, T, cum, nullptr, st);
        length_regulate_gather(xprime, cum, B, T, H, frame0, L_max, features, src_index, st);
        return 0;
    }

    int decode(const float* features, const uint8_t* mask, const int64_t* mel_len, const float* style, int B, int L,
               int zero_padded_mel, float* mel_BLC, float* mel_BCL, cudaStream_t st) {
        check_ready(SEC_DEC);
        ZVX_REQUIRE(B >= 1 && L >= 1, "zvx_decode: empty batch");
        ZVX_REQUIRE(mask || mel_len, "zvx_decode: need mask or mel_len");
        ws.reset();
        if (cfg.decoder_kind == 1) return decode_styletts(features, mask, mel_len, style, B, L, zero_padded_mel, mel_BLC, mel_BCL, st);
        const int tc = P_TF32;
        if (!mask) {
            uint8_t* m = ws.get<uint8_t>((long long)B * L);
            mask_from_lengths(mel_len, B, L, m, st);
            mask = m;
        }
        const int nscln = cfg.dec_layers * 2;
        float* gb = nullptr;
        if (cfg.dec_scln) {
            gb = ws.get<float>((long long)B * nscln * 2 * H);
            linear(style, B, H, scln_w, nullptr, nscln * 2 * H, gb, P_EXACT, st);
        }
        float* x = ws.get<float>((long long)B * L * H);
        add_posenc(features, pos_table(true, L), B, L, H, x, st);
        FFTScratch sc = fft_scratch(B, L, cfg.dec_heads, tc);
        plan_attention(sc, B, L, cfg.dec_heads, mask, st);
        for (int i = 0; i < cfg.dec_layers; ++i) {
            const float* gb1 = gb ? gb + (long long)(2 * i) * 2 * H : nullptr;
            const float* gb2 = gb ? gb + (long long)(2 * i + 1) * 2 * H : nullptr;
            fft_block(x, B, L, cfg.dec_heads, dec[(size_t)i], mask, cfg.dec_scln != 0, gb1, gb2, nscln * 2 * H, tc, sc, st);
        }
        float* mel = mel_BLC ? mel_BLC : ws.get<float>((long long)B * L * cfg.n_mels);
        linear(x, B * L, H, mel_w, mel_b, cfg.n_mels, mel, tc, st);
        if (mel_BCL || zero_padded_mel)
            transpose_mel(mel, mask, zero_padded_mel, B, L, cfg.n_mels, mel_BCL, zero_padded_mel ? mel : nullptr, st);
        return 0;
    }

    // Conv1d over channel-last [B, L, *] with explicit leading dimensions (StyleTTS blocks write into concat buffers).
    void st_conv(const float* x, int ldx, const STConv& c, float* y, int ldy, int B, int L, const float* R, int ldr,
                 float post_scale, cudaStream_t st) {
        GemmArgs a;
        a.A = x; a.lda = ldx; a.W = c.w; a.ldw = c.cin; a.w_tap_stride = (long long)c.cout * c.cin; a.C = y; a.ldc = ldy;
        a.bias = c.b; a.M = B * L; a.N = c.cout; a.K = c.cin; a.taps = c.k; a.R = R; a.ldr = ldr; a.post_scale = post_scale;
        if (c.k > 1) { a.mode = ROW_CONV1D; a.Lout = L; a.Lin = L; a.stride = 1; a.pad = (c.k - 1) / 2; a.dil = 1; }
        gemm(a, P_TF32, st);
    }

    // StyleTTSDecoder.forward (styletts.py:181-205).  `mask` is ignored by the reference decoder (InstanceNorm statistics
    // run over all L frames, padded ones included); it only drives the mel zero-fill of ZeroVox.forward (model.py:283-285).
    int decode_styletts(const float* features, const uint8_t* mask, const int64_t* mel_len, const float* style, int B, int L,
                        int zero_padded_mel, float* mel_BLC, float* mel_BCL, cudaStream_t st) {
        const int RD = 64, BN = 2 * H, CC = BN + RD;
        const long long rows = (long long)B * L;
        const float inv_sqrt2 = (float)(1.0 / std::sqrt(2.0));
        if (!mask && (mel_BCL || zero_padded_mel)) {
            uint8_t* m = ws.get<uint8_t>(rows);
            mask_from_lengths(mel_len, B, L, m, st);
            mask = m;
        }
        float* catA = ws.get<float>(rows * CC);   // [x | asr_res] concat buffers (ld = CC), ping-pong
        float* catB = ws.get<float>(rows * CC);
        float* t1 = ws.get<float>(rows * CC);     // normalised + activated conv operand
        float* t2 = ws.get<float>(rows * BN);     // conv1 output
        float* scb = ws.get<float>(rows * BN);    // learned shortcut
        float* mean = ws.get<float>((long long)B * CC);
        float* rstd = ws.get<float>((long long)B * CC);
        float* hbuf = ws.get<float>((long long)B * 2 * CC);
        float* xa = ws.get<float>(rows * H);
        float* xb = ws.get<float>(rows * H);

        auto block = [&](const STBlock& bl, const float* x, int ldx, float* y, int ldy, bool adain) {
            auto norm_act = [&](const float* src, int lds, int C, const float* g, const float* b, const float* fw, const float* fb) {
                instnorm_stats(src, B, L, C, lds, 1e-5f, mean, rstd, st);
                if (adain) {
                    linear(style, B, H, fw, fb, 2 * C, hbuf, P_EXACT, st);
                    instnorm_apply(src, lds, mean, rstd, hbuf, hbuf + C, 2 * C, 1.f, 0.2f, B, L, C, t1, C, st);
                } else {
                    instnorm_apply(src, lds, mean, rstd, g, b, 0, 0.f, 0.2f, B, L, C, t1, C, st);
                }
            };
            norm_act(x, ldx, bl.cin, bl.n1_g, bl.n1_b, bl.fc1_w, bl.fc1_b);
            st_conv(t1, bl.cin, bl.c1, t2, bl.cmid, B, L, nullptr, 0, 1.f, st);
            norm_act(t2, bl.cmid, bl.cmid, bl.n2_g, bl.n2_b, bl.fc2_w, bl.fc2_b);
            const float* R = x;
            int ldr = ldx;
            if (bl.sc.w) {
                st_conv(x, ldx, bl.sc, scb, bl.cout, B, L, nullptr, 0, 1.f, st);
                R = scb; ldr = bl.cout;
            }
            st_conv(t1, bl.cmid, bl.c2, y, ldy, B, L, R, ldr, inv_sqrt2, st);   // (residual + shortcut) / sqrt(2)
        };

        block(st_enc[0], features, H, catA, CC, false);
        block(st_enc[1], catA, CC, catB, CC, false);
        {   // asr_res = InstanceNorm_affine(conv1x1(enc_seq)) -> tail columns of both concat buffers
            st_conv(features, H, st_asr, t2, RD, B, L, nullptr, 0, 1.f, st);
            instnorm_stats(t2, B, L, RD, RD, 1e-5f, mean, rstd, st);
            instnorm_apply(t2, RD, mean, rstd, st_asr_g, st_asr_b, 0, 0.f, 1.f, B, L, RD, catA + BN, CC, st);
            instnorm_apply(t2, RD, mean, rstd, st_asr_g, st_asr_b, 0, 0.f, 1.f, B, L, RD, catB + BN, CC, st);
        }
        block(st_dec[0], catB, CC, catA, CC, true);
        block(st_dec[1], catA, CC, catB, CC, true);
        block(st_dec[2], catB, CC, xa, H, true);      // the "upsample" block: the residual concat stops after it
        block(st_dec[3], xa, H, xb, H, true);
        block(st_dec[4], xb, H, xa, H, true);
        float* mel = mel_BLC ? mel_BLC : ws.get<float>(rows * cfg.n_mels);
        st_conv(xa, H, st_out, mel, cfg.n_mels, B, L, nullptr, 0, 1.f, st);
        if (mel_BCL || zero_padded_mel)
            transpose_mel(mel, mask, zero_padded_mel, B, L, cfg.n_mels, mel_BCL, zero_padded_mel ? mel : nullptr, st);
        return 0;
    }

    struct View { float* p; long long bs; };

    // Channel-last tensor-core vocoder (gemm_tc.cu + voc_poly.cu / voc_res.cu); same arithmetic graph as vocode_simt below.
    int vocode_tc(const float* mel_BCL, int B, int L, float* wav, cudaStream_t st) {
        const int C0 = cfg.hg_upsample_initial_channel, M = cfg.n_mels;
        const int nk = cfg.hg_num_kernels, nd = cfg.hg_num_dilations, nu = cfg.hg_num_upsamples;
        long long big = std::max((long long)B * L * C0, (long long)B * L * M), t = L;
        for (int i = 0; i < nu; ++i) {
            t *= cfg.hg_upsample_rates[i];
            // the transposed-conv GEMM writes (T_in + 1) * u = T_out + u rows per utterance
            big = std::max(big, (long long)B * (t + cfg.hg_upsample_rates[i]) * (C0 >> (i + 1)));
        }
        float* bX = ws.get<float>(big);    // stage output (MRF accumulator), raw
        float* bXA = ws.get<float>(big);   // its leaky-ReLU'd copy (upsampler operand)
        float* bY = ws.get<float>(big);    // upsampler output, raw
        float* bYA = ws.get<float>(big);
        float* bRA = ws.get<float>(big);
        float* bRAa = ws.get<float>(big);
        float* bRB = ws.get<float>(big);
        float* bRBa = ws.get<float>(big);
        float* bT = ws.get<float>(big);
        float* mel = ws.get<float>((long long)B * L * M);
        transpose_mel(mel_BCL, nullptr, 0, B, /*rows=*/M, /*cols=*/L, mel, nullptr, st);   // [B,M,L] -> [B,L,M]
        {   // conv_pre, stored leaky-ReLU'd: its only consumer is the first upsampler
            TcGemmArgs g;
            g.A = mel; g.K = M; g.Wi = g.Wo = L; g.Hi = g.Ho = B; g.a_sx = M; g.a_sy = (long long)L * M;
            g.W = hg_pre_tc; g.N = C0; g.w_sn = M; g.Z1 = 7; g.w_s1 = (long long)C0 * M; g.ksx = 7; g.pad_x = 3;
            g.bias = hg_pre.b; g.C = bXA; g.c_sx = C0; g.c_sy = (long long)L * C0; g.act_slope = 0.1f;
            gemm_tc_prof(g, st);
        }
        View xa{bXA, (long long)L * C0};
        int T = L;
        size_t ci = 0;
        for (int i = 0; i < nu; ++i) {
            const int cin = C0 >> i, ch = C0 >> (i + 1), u = cfg.hg_upsample_rates[i], pd = u / 2;
            const int kind = hg_stage_kind[(size_t)i];
            const int Tout = T * u;
            const long long ybs = (long long)(Tout + u) * ch;
            {   // ConvTranspose1d: 2-tap GEMM over the T+1 input positions, N = u*ch columns = u consecutive output samples
                TcGemmArgs g;
                g.A = xa.p; g.K = cin; g.Wi = T; g.Wo = T + 1; g.Hi = g.Ho = B; g.a_sx = cin; g.a_sy = xa.bs;
                g.W = hg_ups_tc_w[(size_t)i]; g.N = u * ch; g.w_sn = cin; g.Z1 = 2; g.w_s1 = (long long)u * ch * cin;
                g.ksx = 2; g.pad_x = 1; g.bias = hg_ups_tc_b[(size_t)i];
                g.C = bY; g.c_sx = (long long)u * ch; g.c_sy = ybs;
                if (kind == 2) { g.C2 = bYA; g.slope2 = 0.1f; }
                gemm_tc_prof(g, st);
            }
            T = Tout;
            const View y{bY + (long long)pd * ch, ybs}, ya{bYA + (long long)pd * ch, ybs};
            const long long bs = (long long)T * ch;
            const bool more = (i + 1 < nu);
            for (int j = 0; j < nk; ++j) {
                const int rk = cfg.hg_resblock_kernel_sizes[j];
                const bool pair = (cfg.hg_resblock == 1);
                if (kind == 1) {
                    // whole resblock in one launch; MRF mean accumulated in bX, the last one emits the next operand
                    VocResArgs a;
                    a.x = y.p; a.x_bs = y.bs; a.B = B; a.T = T; a.C = ch; a.k = rk;
                    for (int di = 0; di < nd; ++di, ++ci) {
                        const int dl = cfg.hg_resblock_dilation_sizes[j][di];
                        if (pair) {
                            a.steps[a.nsteps].w = hg_c1_tc[ci]; a.steps[a.nsteps].w_poly = hg_c1_poly[ci];
                            a.steps[a.nsteps].w_pair = hg_c1_pair[ci];
                            a.steps[a.nsteps].b = hg_c1[ci].b; a.steps[a.nsteps].dil = dl; a.steps[a.nsteps++].kind = 0;
                            a.steps[a.nsteps].w = hg_c2_tc[ci]; a.steps[a.nsteps].w_poly = hg_c2_poly[ci];
                            a.steps[a.nsteps].w_pair = hg_c2_pair[ci];
                            a.steps[a.nsteps].b = hg_c2[ci].b; a.steps[a.nsteps].dil = 1; a.steps[a.nsteps++].kind = 1;
                        } else {
                            a.steps[a.nsteps].w = hg_c1_tc[ci]; a.steps[a.nsteps].w_poly = hg_c1_poly[ci];
                            a.steps[a.nsteps].b = hg_c1[ci].b; a.steps[a.nsteps].dil = dl; a.steps[a.nsteps++].kind = 1;
                        }
                    }
                    if (j > 0) { a.acc_in = bX; a.acc_in_bs = bs; }
                    a.out_scale = 1.f / (float)nk;
                    const bool emit_act = (j == nk - 1) && more;
                    a.out = emit_act ? bXA : bX; a.out_bs = bs; a.out_slope = emit_act ? 0.1f : 1.f;
                    prof.begin(ZVX_PROF_VOC_TC, 2.0 * B * T * ch * ch * rk * a.nsteps, 4.0 * B * T * ch * (j > 0 ? 3.0 : 2.0), st);
                    static const bool no_poly = env_set("ZVX_NO_POLY");   // A/B switch (debug builds): voc_res.cu only
                    static const bool dbg_times = env_set("ZVX_VOC_DBG");  // debug builds: print one CTA's phase timestamps
                    long long* dbg = nullptr;
                    if (dbg_times) { dbg = ws.get<long long>(64); ZVX_CUDA_CHECK(cudaMemsetAsync(dbg, 0, 64 * 8, st)); a.dbg = dbg; }
                    static const bool no_pair = env_set("ZVX_NO_PAIR");   // A/B switch (debug builds): skip voc_pair.cu
                    if (no_pair || !pair || !voc_pair_tc(a, st))
                        if (no_poly || !voc_poly_tc(a, st)) voc_resblock_tc(a, st);
                    prof.end(st);
                    if (dbg) {
                        long long h[64];
                        ZVX_CUDA_CHECK(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
                        ZVX_CUDA_CHECK(cudaStreamSynchronize(st));
                        fprintf(stderr, "[voc dbg] C=%d k=%d T=%d:", ch, rk, T);
                        for (int i = 1; i < 32 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
                        fprintf(stderr, " | issuer:");
                        for (int i = 32; i < 64 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
                        fprintf(stderr, "\n");
                    }
                    continue;
                }
                View r = y, ra = ya;
                for (int di = 0; di < nd; ++di, ++ci) {
                    const bool last = (di == nd - 1);
                    const bool first_buf = (r.p != bRA);
                    const View rn{first_buf ? bRA : bRB, bs}, rna{first_buf ? bRAa : bRBa, bs};
                    const int dl = cfg.hg_resblock_dilation_sizes[j][di];
                    {
                        TcGemmArgs g;   // conv over the leaky-ReLU'd copy
                        g.K = ch; g.Wi = g.Wo = T; g.Hi = g.Ho = B; g.a_sx = ch; g.N = ch; g.w_sn = ch; g.Z1 = rk;
                        g.w_s1 = (long long)ch * ch; g.ksx = rk; g.c_sx = ch;
                        if (pair) {
                            g.A = ra.p; g.a_sy = ra.bs; g.W = hg_c1_tc[ci]; g.bias = hg_c1[ci].b; g.dil = dl;
                            g.pad_x = (rk - 1) / 2 * dl; g.C = bT; g.c_sy = bs; g.act_slope = 0.1f;
                            gemm_tc_prof(g, st);
                            g.A = bT; g.a_sy = bs; g.W = hg_c2_tc[ci]; g.bias = hg_c2[ci].b; g.dil = 1; g.pad_x = (rk - 1) / 2;
                            g.act_slope = 1.f;
                        } else {
                            g.A = ra.p; g.a_sy = ra.bs; g.W = hg_c1_tc[ci]; g.bias = hg_c1[ci].b; g.dil = dl;
                            g.pad_x = (rk - 1) / 2 * dl;
                        }
                        g.R = r.p; g.r_sx = ch; g.r_sy = r.bs;
                        if (last) {
                            g.C = bX; g.c_sy = bs; g.acc_mode = 1; g.acc_init = (j == 0); g.acc_scale = 1.f / (float)nk;
                            if (j == nk - 1 && more) { g.C2 = bXA; g.slope2 = 0.1f; }
                        } else {
                            g.C = rn.p; g.c_sy = rn.bs; g.C2 = rna.p; g.slope2 = 0.1f;
                        }
                        gemm_tc_prof(g, st);
                    }
                    r = rn; ra = rna;
                }
            }
            xa = View{bXA, bs};
        }
        const int chl = hg_post.cin;
        prof.begin(ZVX_PROF_VOC_CONV, 2.0 * B * T * chl * 7, 4.0 * B * T * (chl + 1), st);
        conv_post_cl(bX, (long long)T * chl, hg_post_tc, hg_post.b, B, T, chl, 7, 0.01f, wav, st);  // F.leaky_relu default slope (hifigan.py:126)
        prof.end(st);
        return 0;
    }

    int vocode(const float* mel, int B, int L, float* wav, cudaStream_t st) {
        check_ready(SEC_VOC);
        ZVX_REQUIRE(B >= 1 && L >= 1 && B <= 65535, "zvx_vocode: bad sizes");
        ws.reset();
        if (cfg.tensor_core_policy != 0 && hg_tc_ok) return vocode_tc(mel, B, L, wav, st);
        const int C0 = cfg.hg_upsample_initial_channel;
        const int nk = cfg.hg_num_kernels, nd = cfg.hg_num_dilations;
        // largest activation: max over stages of C * T
        long long big = (long long)C0 * L, t = L;
        for (int i = 0; i < cfg.hg_num_upsamples; ++i) {
            t *= cfg.hg_upsample_rates[i];
            big = std::max(big, (long long)(C0 >> (i + 1)) * t);
        }
        big *= B;
        float* x = ws.get<float>(big);
        float* y = ws.get<float>(big);
        float* ra = ws.get<float>(big);
        float* rb = ws.get<float>(big);
        float* tmp = ws.get<float>(big);
        Conv1dArgs a;
        a.x = mel; a.w = hg_pre.w; a.bias = hg_pre.b; a.B = B; a.Cin = cfg.n_mels; a.Cout = C0; a.T = L; a.k = 7;
        a.dil = 1; a.in_slope = 1.f; a.out = x;
        vconv(a, st);
        int T = L;
        size_t ci = 0;
        for (int i = 0; i < cfg.hg_num_upsamples; ++i) {
            const HGConv& up = hg_ups[(size_t)i];
            prof.begin(ZVX_PROF_VOC_UPSAMPLE, 2.0 * B * T * up.cin * up.cout * up.k,
                       4.0 * B * T * (up.cin + (double)up.cout * cfg.hg_upsample_rates[i]), st);
            conv_transpose1d_cf(x, up.w, up.b, B, up.cin, up.cout, T, up.k, cfg.hg_upsample_rates[i], 0.1f, y, st);
            prof.end(st);
            T *= cfg.hg_upsample_rates[i];
            const int ch = up.cout;
            for (int j = 0; j < nk; ++j) {
                const float* r = y;
                for (int di = 0; di < nd; ++di, ++ci) {
                    const bool last = (di == nd - 1);
                    float* rn = (r == ra) ? rb : ra;
                    Conv1dArgs c;
                    c.B = B; c.Cin = ch; c.Cout = ch; c.T = T; c.in_slope = 0.1f;
                    if (cfg.hg_resblock == 1) {
                        const HGConv &c1 = hg_c1[ci], &c2 = hg_c2[ci];
                        c.x = r; c.w = c1.w; c.bias = c1.b; c.k = c1.k; c.dil = c1.dil; c.out = tmp;
                        vconv(c, st);
                        c.x = tmp; c.w = c2.w; c.bias = c2.b; c.k = c2.k; c.dil = 1; c.res = r;
                    } else {
                        const HGConv& c1 = hg_c1[ci];
                        c.x = r; c.w = c1.w; c.bias = c1.b; c.k = c1.k; c.dil = c1.dil; c.res = r;
                    }
                    if (last) {  // xs (+)= resblock(x); x = xs / num_kernels   (hifigan.py:119-125)
                        c.out = nullptr; c.acc = x; c.acc_init = (j == 0); c.acc_scale = 1.f / (float)nk;
                    } else {
                        c.out = rn;
                    }
                    vconv(c, st);
                    r = rn;
                }
            }
        }
        Conv1dArgs p;
        p.x = x; p.w = hg_post.w; p.bias = hg_post.b; p.B = B; p.Cin = hg_post.cin; p.Cout = 1; p.T = T; p.k = 7;
        p.dil = 1; p.in_slope = 0.01f; p.tanh_out = 1; p.out = wav;  // F.leaky_relu default slope (hifigan.py:126)
        vconv(p, st);
        return 0;
    }

    int set_option(const char* name, int64_t value) {
        ZVX_REQUIRE(name, "zvx_set_option: null name");
        const std::string n(name);
        if (n == "score_workspace_bytes") {
            ZVX_REQUIRE(value >= (1 << 20), "zvx_set_option: score_workspace_bytes must be >= 1 MiB");
            kScoreBytes = value;
            return 0;
        }
        if (n == "fused_attention") {
            ZVX_REQUIRE(value >= 0 && value <= 3, "zvx_set_option: fused_attention must be 0..3");
            kFusedAttention = value != 0;
            kAttentionVariant = value >= 2 ? (int)value - 1 : 0;
            return 0;
        }
        if (n == "pdl") {   // programmatic dependent launch of the hot kernels (common.cuh); process-wide like the launch counter
            ZVX_REQUIRE(value == 0 || value == 1, "zvx_set_option: pdl must be 0 or 1");
            g_pdl = (int)value;
            return 0;
        }
        throw Error("zvx_set_option: unknown option '" + n + "'");
    }

    static constexpr int kMaxBatch = 65536;
    // attention-score workspace budget: longer sequences are processed in chunks of query rows (exact)
    long long kScoreBytes = 4LL << 30;   // zvx_set_option("score_workspace_bytes")
    bool kFusedAttention = true;         // zvx_set_option("fused_attention"): 0 = QK^T / softmax / PV as three kernels
    int kAttentionVariant = 0;           // AttnFusedArgs::variant (0 = the kernel chooses)
    // attention batch-slice budget (score bytes per slice); 0 = whole batch per launch.  Default set by measurement.
    static constexpr long long kAttnSliceBytes = 0;

    zvx_config cfg;
    int dev = 0;
    int num_sms = 0;
    int H = 0;
    std::string err;
    std::map<std::string, HostTensor> raw;
    std::vector<void*> owned;
    bool finalized = false;
    enum { SEC_ENC = 0, SEC_DEC = 1, SEC_SPK = 2, SEC_VOC = 3 };
    bool sec_ready[4] = {false, false, false, false};
    std::string sec_err[4];
    Workspace ws;
    Workspace ws_spk;                    // the speaker net's own (spkemb_encode runs it next to the encoder)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_vp_x = nullptr, ev_vp_done = nullptr;
    Profiler prof;
    bool split_on = !env_set("ZVX_NO_SPLIT");   // 3xTF32 for P_EXACT contractions (switchable in debug builds only)
    std::map<const float*, const float*> w_lo;
    float* lo_buf = nullptr;
    long long lo_cap = 0;
    int64_t* pinned_len = nullptr;

    float *enc_pos_param = nullptr, *dec_pos_param = nullptr, *enc_pos = nullptr, *dec_pos = nullptr;
    int enc_pos_rows = 0, dec_pos_rows = 0;
    float *phon_emb = nullptr, *punct_emb = nullptr, *pitch_emb = nullptr, *energy_emb = nullptr;
    std::vector<FFTLayer> enc, dec;
    VarPredictor vp[3];
    float *scln_w = nullptr, *mel_w = nullptr, *mel_b = nullptr;
    float *stem_w = nullptr, *stem_b = nullptr, *stem_s = nullptr, *stem_sh = nullptr;
    std::vector<SEBlock> se_blocks;
    int spk_D = 0;
    float *att_w0 = nullptr, *att_b0 = nullptr, *att_bn_s = nullptr, *att_bn_b = nullptr, *att_w3 = nullptr,
          *att_b3 = nullptr, *spk_fc_w = nullptr, *spk_fc_b = nullptr;
    std::vector<STBlock> st_enc, st_dec;
    STConv st_asr, st_out;
    float *st_asr_g = nullptr, *st_asr_b = nullptr;
    HGConv hg_pre, hg_post;
    std::vector<HGConv> hg_ups, hg_c1, hg_c2;
    bool hg_tc_ok = false;
    std::vector<int> hg_stage_kind;   // per upsample stage: 1 = fused pair kernel, 2 = generic tcgen05 implicit GEMM
    float *hg_pre_tc = nullptr, *hg_post_tc = nullptr;
    std::vector<float*> hg_ups_tc_w, hg_ups_tc_b, hg_c1_tc, hg_c2_tc, hg_c1_poly, hg_c2_poly, hg_c1_pair, hg_c2_pair;
};

}  // namespace zvx

// -------------------------------------------------------------------------------------------------- C ABI
struct zvx_handle {
    std::unique_ptr<zvx::Engine> eng;
};

#define ZVX_GUARD(h, ...)                                                      \
    if (!(h) || !(h)->eng) return -1;                                          \
    try {                                                                      \
        __VA_ARGS__;                                                           \
    } catch (const std::exception& e) {                                        \
        (h)->eng->err = e.what();                                              \
        return 1;                                                              \
    } catch (...) {                                                            \
        (h)->eng->err = "unknown error";                                       \
        return 2;                                                              \
    }

extern "C" {

int zvx_abi_version(void) { return ZVX_ABI_VERSION; }

int zvx_create(const zvx_config* cfg, int device, zvx_handle** out) {
    if (!cfg || !out) {
        zvx::g_create_error = "zvx_create: null argument";
        return -1;
    }
    try {
        auto* h = new zvx_handle();
        h->eng.reset(new zvx::Engine(*cfg, device));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        zvx::g_create_error = e.what();
        return 1;
    } catch (...) {
        zvx::g_create_error = "unknown error";
        return 2;
    }
}

void zvx_destroy(zvx_handle* h) { delete h; }

const char* zvx_last_error(const zvx_handle* h) {
    if (!h || !h->eng) return zvx::g_create_error.c_str();
    return h->eng->err.c_str();
}

int zvx_set_weight(zvx_handle* h, const char* key, const void* data, const int64_t* shape, int ndim) {
    ZVX_GUARD(h, return h->eng->set_weight(key, data, shape, ndim));
}

int zvx_finalize_weights(zvx_handle* h) { ZVX_GUARD(h, return h->eng->finalize()); }

int zvx_spkemb(zvx_handle* h, const float* ref_mel, int B, int T_ref, float* style, void* stream) {
    ZVX_GUARD(h, return h->eng->spkemb(ref_mel, B, T_ref, style, (cudaStream_t)stream));
}

int zvx_spkemb_encode(zvx_handle* h, const float* ref_mel, int T_ref, float* style, const int32_t* phoneme, const int32_t* puncts,
                      const uint8_t* phoneme_mask, const int32_t* forced_dur, int B, int T, float* pitch, float* energy,
                      float* log_dur, int32_t* dur_rounded, int64_t* mel_len, float* xprime, int64_t* mel_len_host,
                      int* L_max_out, void* stream) {
    ZVX_GUARD(h, return h->eng->spkemb_encode(ref_mel, T_ref, style, phoneme, puncts, phoneme_mask, forced_dur, B, T, pitch,
                                              energy, log_dur, dur_rounded, mel_len, xprime, mel_len_host, L_max_out,
                                              (cudaStream_t)stream));
}

int zvx_encode(zvx_handle* h, const int32_t* phoneme, const int32_t* puncts, const uint8_t* phoneme_mask,
               const float* style, const int32_t* forced_dur, int B, int T, float* pitch, float* energy,
               float* log_dur, int32_t* dur_rounded, int64_t* mel_len, float* xprime, int64_t* mel_len_host,
               int* L_max_out, void* stream) {
    ZVX_GUARD(h, return h->eng->encode(phoneme, puncts, phoneme_mask, style, forced_dur, B, T, pitch, energy,
                                       log_dur, dur_rounded, mel_len, xprime, mel_len_host, L_max_out,
                                       (cudaStream_t)stream));
}

int zvx_length_regulate(zvx_handle* h, const float* xprime, const int32_t* dur, int B, int T, int L_max,
                        float* features, int32_t* src_index, void* stream) {
    ZVX_GUARD(h, return h->eng->length_regulate(xprime, dur, B, T, 0, L_max, features, src_index, (cudaStream_t)stream));
}

int zvx_length_regulate_chunk(zvx_handle* h, const float* xprime, const int32_t* dur, int B, int T, int frame0,
                              int n_frames, float* features, int32_t* src_index, void* stream) {
    ZVX_GUARD(h, return h->eng->length_regulate(xprime, dur, B, T, frame0, n_frames, features, src_index,
                                                (cudaStream_t)stream));
}

int zvx_decode(zvx_handle* h, const float* features, const uint8_t* mask, const int64_t* mel_len,
               const float* style, int B, int L, int zero_padded_mel, float* mel_BLC, float* mel_BCL, void* stream) {
    ZVX_GUARD(h, return h->eng->decode(features, mask, mel_len, style, B, L, zero_padded_mel, mel_BLC, mel_BCL,
                                       (cudaStream_t)stream));
}

int zvx_vocode(zvx_handle* h, const float* mel_BCL, int B, int L, float* wav, void* stream) {
    ZVX_GUARD(h, return h->eng->vocode(mel_BCL, B, L, wav, (cudaStream_t)stream));
}

int zvx_profile_enable(zvx_handle* h, int on) {
    ZVX_GUARD(h, { h->eng->prof.reset(); h->eng->prof.on = (on != 0); return 0; });
}

int zvx_profile_read(zvx_handle* h, int kernel_class, double* ms, int64_t* launches, double* flops, double* bytes) {
    ZVX_GUARD(h, {
        ZVX_REQUIRE(ms && launches && flops && bytes, "zvx_profile_read: null output");
        h->eng->prof.read(kernel_class, ms, launches, flops, bytes);
        return 0;
    });
}

int zvx_debug_gemm(zvx_handle* h, const zvx_gemm_desc* d, int use_tc, void* stream) {
    ZVX_GUARD(h, {
        ZVX_REQUIRE(d && d->A && d->W && d->C, "zvx_debug_gemm: null operand");
        ZVX_CUDA_CHECK(cudaSetDevice(h->eng->dev));
        zvx::GemmArgs a;
        a.A = d->A; a.lda = d->lda; a.W = d->W; a.ldw = d->ldw; a.w_tap_stride = (long long)d->N * d->ldw;
        a.C = d->C; a.ldc = d->ldc; a.bias = d->bias; a.scale = d->scale; a.shift = d->shift; a.R = d->R; a.ldr = d->ldc;
        a.M = d->M; a.N = d->N; a.K = d->K; a.taps = d->taps; a.relu_first = d->relu_first; a.relu_last = d->relu_last;
        if (d->mode == 1) {
            a.mode = zvx::ROW_CONV1D; a.Lout = a.Lin = d->L; a.stride = 1; a.pad = d->pad; a.dil = d->dil;
        } else if (d->mode == 2) {
            a.mode = zvx::ROW_CONV2D; a.Hi = d->Hh; a.Wi = d->Ww; a.ksize = d->ksize; a.stride = d->stride > 1 ? d->stride : 1;
            a.pad = d->pad;
            a.Ho = (a.Hi + 2 * a.pad - a.ksize) / a.stride + 1;
            a.Wo = (a.Wi + 2 * a.pad - a.ksize) / a.stride + 1;
        }
        if (use_tc) {
            zvx::TcGemmArgs t;
            ZVX_REQUIRE(zvx::gemm_tc_from(a, &t), "zvx_debug_gemm: layout not supported by the tcgen05 path");
            float *alo = nullptr, *wlo = nullptr;
            if (use_tc == 2) {   // 3xTF32 split: prepare the low parts (test hook only: allocates)
                const long long na = zvx::Engine::a_extent(a), nw = (long long)a.taps * a.N * a.ldw;
                ZVX_REQUIRE(na % 4 == 0 && nw % 4 == 0, "zvx_debug_gemm: split mode needs sizes that are multiples of 4");
                ZVX_CUDA_CHECK(cudaMalloc(&alo, (size_t)na * sizeof(float)));
                ZVX_CUDA_CHECK(cudaMalloc(&wlo, (size_t)nw * sizeof(float)));
                zvx::tf32_split_lo(a.A, alo, na, (cudaStream_t)stream);
                zvx::tf32_split_lo(a.W, wlo, nw, (cudaStream_t)stream);
                t.A_lo = alo; t.W_lo = wlo;
            }
            zvx::gemm_tc(t, (cudaStream_t)stream);
            if (alo) {
                ZVX_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
                cudaFree(alo);
                cudaFree(wlo);
            }
        } else {
            zvx::gemm_simt(a, (cudaStream_t)stream);
        }
        return 0;
    });
}

int zvx_set_option(zvx_handle* h, const char* name, int64_t value) { ZVX_GUARD(h, return h->eng->set_option(name, value)); }

int64_t zvx_workspace_bytes(const zvx_handle* h) { return (h && h->eng) ? h->eng->ws.bytes() + h->eng->ws_spk.bytes() : 0; }

int64_t zvx_launch_count(const zvx_handle* h) {
    (void)h;
    return zvx::g_launches;
}

}  // extern "C"
