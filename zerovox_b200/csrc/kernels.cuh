// Launchers for the non-GEMM kernels (HBM-bound gather / normalise / elementwise work).
#pragma once
#include "common.cuh"

#include <vector>

namespace zvx {

// ---- FastSpeech2 acoustic model (kernels_fs2.cu) ---------------------------------------------------
// K1: out[b,t,:] = cat(phon_emb[phoneme[b,t]], punct_emb[puncts[b,t]]) + pos[t]      (fs2.py:372-392)
void embed_posenc(const int32_t* phoneme, const int32_t* puncts, const float* phon_emb, const float* punct_emb,
                  const float* pos, int B, int T, int E, int P, float* out, cudaStream_t st);

// K6: LayerNorm (eps inside sqrt, biased variance) or SCLN (unbiased std, sigma + eps; fs2.py:76-90).
struct NormArgs {
    const float* x = nullptr;     // [rows, C]
    const float* res = nullptr;   // optional residual [rows, C] added to x before normalising (fs2.py:158-162, 205-208)
    float* out = nullptr;         // [rows, C] (may alias x); ignored when dot_w != nullptr
    int rows = 0, C = 0;
    int rows_per_batch = 1;       // L: batch index of a row = row / L
    int scln = 0;
    const float* gamma = nullptr; // LN weight [C]
    const float* beta = nullptr;  // LN bias [C]
    const float* gb = nullptr;    // SCLN: per-batch affine rows, bias at gb[b*gb_ld + 0..C), gain at +C..2C
    int gb_ld = 0;
    float eps = 1e-5f;
    const uint8_t* mask = nullptr;  // [rows], 1 = padding -> output row (or dot) is 0   (fs2.py:225,228,560-561)
    const float* dot_w = nullptr;   // optional fused Linear(C -> 1): dot_out[row] = <norm(x), dot_w> + dot_b[0]
    const float* dot_b = nullptr;
    float* dot_out = nullptr;
};
void layer_norm(const NormArgs& a, cudaStream_t st);

// K4 (softmax part): in-place over S[z][q][0..L) with z = b*nh + h:  softmax_j( S/temperature, key mask -> -inf ).
// Columns [L, ldS) are zeroed so S can feed a K-padded GEMM.
void attn_softmax(float* S, int nz, int nh, int Lq, int L, int ldS, const uint8_t* key_mask /*[B,L]*/, int mask_ld,
                  float temperature, cudaStream_t st);

// x[b,t,:] += v[b,:]   (fs2.py:740-741)
void add_batch_vector(float* x, const float* v, int B, int T, int C, cudaStream_t st);

// K9: x[b,t,:] += table[clamp(rint(pred[b,t]*(n_bins-1)), 0, n_bins-1), :]   (fs2.py:639, 649, 668, 672)
void bucket_embed_add(float* x, const float* pred, const float* table, int rows, int C, int n_bins,
                      int32_t* bucket_out /*nullable*/, cudaStream_t st);

// K10: dur[b,t] = forced ? forced[b,t] : clamp(rint(exp(log_d) - 1), 0)      (fs2.py:674-681)
void duration_round(const float* log_d, const int32_t* forced, int32_t* dur, int n, cudaStream_t st);

// K11a: inclusive scan of max(dur,0) along T; total -> mel_len (int64).      (fs2.py:447-455)
void duration_scan(const int32_t* dur, int B, int T, int32_t* cum, int64_t* mel_len, cudaStream_t st);

// K11b: gather of the frames [frame0, frame0 + L_max): features[b,f-frame0,:] = x[b, idx, :], idx = upper_bound(cum[b], f);
// zero rows past mel_len. (fs2.py:403-459)
void length_regulate_gather(const float* x, const int32_t* cum, int B, int T, int C, int frame0, int L_max,
                            float* features, int32_t* src_index /*nullable*/, cudaStream_t st);

// K12: out[b,l,:] = x[b,l,:] + pos[l,:]      (fs2.py:287-304)
void add_posenc(const float* x, const float* pos, int B, int L, int C, float* out, cudaStream_t st);

// mask[b,l] = l >= mel_len[b]      (model.py:269-273, fs2.py:565-573)
void mask_from_lengths(const int64_t* mel_len, int B, int L, uint8_t* mask, cudaStream_t st);

// [B,L,C] -> [B,C,L], rows with mask==1 zeroed when zero_masked (model.py:283-285); also optionally fixes up the
// [B,L,C] copy in place.
void transpose_mel(const float* mel_BLC, const uint8_t* mask, int zero_masked, int B, int L, int C, float* mel_BCL,
                   float* mel_BLC_inplace /*nullable*/, cudaStream_t st);

// ---- StyleTTS decoder (styletts.py): InstanceNorm1d over time, channel-last [B, L, ld] ---------------
void instnorm_stats(const float* x, int B, int L, int C, int ld, float eps, float* mean, float* rstd, cudaStream_t st);
// y = lrelu( (g_add + g[b*g_bs + c]) * (x - mean) * rstd + bt[b*g_bs + c], slope )   (affine IN: g_bs = 0, g_add = 0;
// AdaIN: g = h, bt = h + C, g_bs = 2C, g_add = 1)
void instnorm_apply(const float* x, int ld, const float* mean, const float* rstd, const float* g, const float* bt,
                    int g_bs, float g_add, float slope, int B, int L, int C, float* y, int ldy, cudaStream_t st);

// ---- HiFi-GAN (kernels_hifigan.cu), channel-first [B, C, T] ------------------------------------------
struct Conv1dArgs {
    const float* x = nullptr;   // [B, Cin, T]
    const float* w = nullptr;   // packed [Cin][k][Cout]
    const float* bias = nullptr;  // [Cout]
    int B = 0, Cin = 0, Cout = 0, T = 0, k = 1, dil = 1;  // 'same' padding (k*d - d)/2      (hifigan.py:22-23)
    float in_slope = 1.f;       // leaky-ReLU slope applied to x on load (1 = identity)
    const float* res = nullptr; // optional residual [B, Cout, T] added to the result
    float* out = nullptr;       // optional plain output
    float* acc = nullptr;       // optional accumulator: acc = (acc_init ? 0 : acc) + result * acc_scale
    int acc_init = 0;
    float acc_scale = 1.f;
    int tanh_out = 0;
};
void conv1d_cf(const Conv1dArgs& a, cudaStream_t st);

// ConvTranspose1d(stride u, kernel k, padding (k-u)/2) with leaky-ReLU on load   (hifigan.py:99-102, 117-118)
// x [B,Cin,T] -> out [B,Cout,T*u]; w packed [Cin][k][Cout].
void conv_transpose1d_cf(const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int T, int k,
                         int u, float in_slope, float* out, cudaStream_t st);

// ---- HiFi-GAN on the tensor cores, channel-last [B, T, C] (voc_res.cu, voc_poly.cu) ---------------------------------
// [Cout][Cin][k] (PyTorch Conv1d / flattened Conv2d) -> shared-memory image [k][Cin/4][max(Cout,16)][4], TF32-rounded
std::vector<float> voc_pack_weight(const float* w, int cout, int cin, int k);

// Whole MRF residual block fused (voc_res.cu): ResBlock1 = [conv_{k,d} (kind 0), conv_{k,1} (kind 1)] per dilation,
// ResBlock2 = [conv_{k,d} (kind 1)] per dilation; kind 1 steps add the residual.  Result:
//   out = lrelu( resblock(x) * out_scale + (acc_in ? acc_in : 0), out_slope )      (acc_in may alias out)
struct VocResArgs {
    static constexpr int MAX_STEPS = 8;
    struct Step {
        const float* w = nullptr;        // packed image of voc_pack_weight (voc_res.cu)
        const float* w_poly = nullptr;   // packed image of voc_poly_pack_weight (voc_poly.cu), optional
        const float* w_pair = nullptr;   // packed image of voc_pair_pack_weight (voc_pair.cu), optional
        const float* b = nullptr; int dil = 1; int kind = 1;
    };
    const float* x = nullptr; long long x_bs = 0;      // raw input [B][T][C], batch stride in elements
    int B = 0, T = 0, C = 0, k = 1, nsteps = 0;
    Step steps[MAX_STEPS];                             // weights in the packed image of voc_pack_weight
    float in_slope = 0.1f, mid_slope = 0.1f;
    const float* acc_in = nullptr; long long acc_in_bs = 0;
    float out_scale = 1.f, out_slope = 1.f;
    float* out = nullptr; long long out_bs = 0;
    const void* sched = nullptr;                       // voc_poly.cu: device copy of the MMA schedule (set by voc_poly_tc)
    long long* dbg = nullptr;                          // optional: clock64 checkpoints of one CTA (tools/voc_phase_times.py)
};
bool voc_resblock_supported(int C, int k, const int* dils, int nd, bool pair);
void voc_resblock_tc(const VocResArgs& a, cudaStream_t st);
// Polyphase formulation (voc_poly.cu): P = 128/C output samples per accumulator row, Toeplitz-expanded weights.
// voc_poly_tc returns false (nothing launched) when the shape is outside its plan; callers then use voc_resblock_tc.
bool voc_poly_supported(int C, int k, const int* dils, int nd, bool pair);
bool voc_poly_tc(const VocResArgs& a, cudaStream_t st);
std::vector<float> voc_poly_pack_weight(const float* w, int C, int k, int dil);
// ResBlock1 in the residue-major / trimmed-Toeplitz formulation, two CTAs per SM where they fit (voc_pair.cu): the first
// choice for C in {8, 16, 32}; returns false (nothing launched) outside its plan, callers then fall back to voc_poly_tc.
bool voc_pair_supported(int C, int k, const int* dils, int nd);
bool voc_pair_tc(const VocResArgs& a, cudaStream_t st);
std::vector<float> voc_pair_pack_weight(const float* w, const float* bias, int C, int k);

// conv_post on channel-last input: wav[b,t] = tanh(bias + sum_{j,c} w[j][c] * lrelu(x[b, t+j-(k-1)/2, c], slope))
void conv_post_cl(const float* x, long long x_bs, const float* w /*[k][C]*/, const float* bias, int B, int T, int C,
                  int k, float slope, float* wav, cudaStream_t st);

// ---- fused attention (attn_fused.cu) ------------------------------------------------------------------
// out[b, q, h*dk + c] = sum_j softmax_j(<Q[b,q,h,:], K[b,j,h,:]> / temperature | key j not masked) V[b,j,h,c]   (fs2.py:101-163)
// qk: [B*L, 2H] row-major, Q in columns [0, H), K in [H, 2H), head h in columns h*dk..; vt: V transposed per utterance,
// vt[(b*H + h*dk + c) * Lp + j]; key_mask (optional): key_mask[b*mask_ld + j] != 0 masks key j of utterance b.
// Device-built tile list of one (mask, shape): key blocks that are entirely masked are never visited (exact: they contribute
// exp(-inf) = 0), and with skip_masked_queries the query tiles past an utterance's last unmasked position are left out as well —
// their output rows are NOT written; the FFT block zero-fills them after the layer norm (fs2.py:226, 229).  The list is what
// balances a ragged batch over the SMs.  One plan serves every layer that shares the mask.
struct AttnPlan { const int* dev = nullptr; int rows_per_tile = 0; int variant = 0; };
struct AttnFusedArgs {
    const float* qk = nullptr; const float* vt = nullptr; float* out = nullptr;
    const uint8_t* key_mask = nullptr; int mask_ld = 0;
    int B = 0, L = 0, n_head = 1, dk = 0, H = 0, Lp = 0;
    float temperature = 1.f;
    int variant = 0;   // 0 = choose; 1 = single-CTA kernel; 2 = CTA-pair kernel (cta_group::2, Q resident)
    const AttnPlan* plan = nullptr;
    double flops() const { return 4.0 * B * n_head * (double)L * L * dk; }
    double bytes() const { return 4.0 * 4.0 * B * (double)L * H; }
};
size_t attn_plan_bytes(int B, int L, int n_head);
// ws: attn_plan_bytes() bytes, 16-byte aligned; uses a.key_mask / mask_ld / B / L / n_head / dk / variant
void attn_plan(const AttnFusedArgs& a, bool skip_masked_queries, int* ws, AttnPlan& plan, cudaStream_t st);
bool attn_fused_supported(const AttnFusedArgs& a);
void attn_fused(const AttnFusedArgs& a, cudaStream_t st);

// ---- ResNetSE34V2 (kernels_spk.cu), channel-last [B, H, W, C] ---------------------------------------
// InstanceNorm1d over time (no affine, biased var, eps 1e-5): ref_mel [B,T,n_mels] -> out [B, n_mels(H), T(W)]
void instance_norm_time(const float* ref_mel, int B, int T, int n_mels, float* out, cudaStream_t st);
// stem: Conv2d(1->C, 3x3, pad 1, bias) -> ReLU -> BN(scale, shift); in [B,H,W] -> out [B,H,W,C]
void stem_conv3x3(const float* in, const float* w /*[9][C]*/, const float* bias, const float* scale,
                  const float* shift, int B, int H, int W, int C, float* out, cudaStream_t st);
// SE squeeze + excitation (ResNetSE34V2.py:52-67): S = hw_mean_splits(B, HW) partial sums over HW -> partial [B, S, C]; the block
// of an utterance that finishes last forms p = (sum_s partial[b,s,:]) / HW and y = sigmoid(W2 relu(W1 p + b1) + b2), W1 [R,C],
// W2 [C,R].  ticket [B] ints, zero before the first launch, left zero.
int hw_mean_splits(int B, int HW);
void se_squeeze_excite(const float* x, int B, int HW, int C, int S, float* partial, int* ticket, const float* w1,
                       const float* b1, const float* w2, const float* b2, int R, float* y, cudaStream_t st);
// out = relu(x * y[b,c] + res)
void se_scale_add_relu(const float* x, const float* y, const float* res, int B, int HW, int C, float* out,
                       cudaStream_t st);
// [B, Hh, W, C] -> [B, W, C*Hh] with feature index c*Hh + h  (x.reshape(B, -1, W), ResNetSE34V2.py:196)
void spk_flatten(const float* x, int B, int Hh, int W, int C, float* out, cudaStream_t st);
// softmax over time of logits [B,W,D], then attentive statistics (mu, sigma) of feat [B,W,D] -> out [B, 2D] (ASP)
// or weighted mean only -> out [B, D] (SAP)       (ResNetSE34V2.py:198-204)
void attentive_pool(const float* feat, const float* logits, int B, int W, int D, int asp, float* out,
                    cudaStream_t st);
// x[b,:] /= max(||x[b,:]||_2, 1e-12)       (F.normalize, ResNetSE34V2.py:207-208)
void l2_normalize(float* x, int B, int C, cudaStream_t st);

}  // namespace zvx
