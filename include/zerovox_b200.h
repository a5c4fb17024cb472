/*
 * zerovox_b200 — C ABI of the B200-native ZeroVOX inference engine (libzerovox_b200.so).
 *
 * The reference (gooofy/zerovox @ 56a4316) has no FFI: its boundary is the Python module API
 * of zerovox/tts/model.py and its sub-modules.  Each entry point below replaces the eval-mode
 * forward of one of those modules; the Python shims in zerovox_b200/tts/ keep the reference
 * signatures and state_dict keys and call these symbols through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every tensor pointer is a DEVICE pointer on the handle's device unless it says "host";
 *     all memory is caller-owned, fp32 row-major contiguous unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); work is enqueued on it
 *     and the call returns without synchronising, except where stated;
 *   - return 0 on success, non-zero on error; the message is available from zvx_last_error().
 *     No C++ exception ever crosses the ABI;
 *   - a handle is bound to one device and is not thread-safe (the reference is single-threaded).
 *   - there is no CPU path: every entry point fails if no CUDA device is usable.
 */
#ifndef ZEROVOX_B200_H
#define ZEROVOX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZVX_MAX_UPSAMPLES 8
#define ZVX_MAX_RESBLOCK_KERNELS 8
#define ZVX_MAX_DILATIONS 4

#define ZVX_ABI_VERSION 1

typedef struct zvx_handle zvx_handle;

/* Hyper-parameters: ZeroVox.__init__ kwargs (zerovox/tts/model.py:159-201; values from
 * configs/tts_medium.yaml:16-51) plus the HiFi-GAN config.json fields read by
 * Generator.__init__ (zerovox/tts/hifigan.py:93-110). */
typedef struct zvx_config {
    int32_t abi_version;          /* must be ZVX_ABI_VERSION */
    int32_t num_phones;           /* Symbols.num_phones (embedding has +1 rows, fs2.py:350) */
    int32_t num_puncts;           /* Symbols.num_puncts incl. _NP_ (embedding has +1 rows, fs2.py:354) */
    int32_t emb_dim;              /* 512 */
    int32_t punct_emb_dim;        /* 16  -> hidden = emb_dim + punct_emb_dim = 528 */
    int32_t max_txt_len;          /* 512 */
    int32_t max_mel_len;          /* 1750 */
    int32_t enc_layers;           /* 4 */
    int32_t enc_heads;            /* 2 */
    int32_t vp_filter_size;       /* 256 */
    int32_t vp_kernel_size;       /* 3 */
    int32_t ve_n_bins;            /* 256 */
    int32_t decoder_kind;         /* 0 = fastspeech2 (FFT blocks + SCLN); 1 = styletts (InstanceNorm / AdaIN conv stacks, styletts.py:181-205) */
    int32_t dec_layers;           /* 6 */
    int32_t dec_heads;            /* 2 */
    int32_t conv_filter_size;     /* 1024 */
    int32_t conv_kernel_size[2];  /* {9, 1} */
    int32_t dec_scln;             /* 1 */
    int32_t resnet_layers[4];     /* {3,4,6,3} */
    int32_t resnet_num_filters[4];/* {32,64,128,256} */
    int32_t resnet_encoder_type;  /* 0 = SAP, 1 = ASP */
    int32_t n_mels;               /* 80 */
    int32_t hop_length;           /* 256 (= product of upsample rates) */
    /* HiFi-GAN generator */
    int32_t hg_resblock;          /* 1 or 2 */
    int32_t hg_num_upsamples;
    int32_t hg_upsample_rates[ZVX_MAX_UPSAMPLES];
    int32_t hg_upsample_kernel_sizes[ZVX_MAX_UPSAMPLES];
    int32_t hg_upsample_initial_channel;
    int32_t hg_num_kernels;
    int32_t hg_resblock_kernel_sizes[ZVX_MAX_RESBLOCK_KERNELS];
    int32_t hg_num_dilations;
    int32_t hg_resblock_dilation_sizes[ZVX_MAX_RESBLOCK_KERNELS][ZVX_MAX_DILATIONS];
    /* numerics policy: 0 = every contraction in fp32 FMA; 1 = decoder / vocoder / speaker-net
     * contractions on TF32 tensor cores (tcgen05), encoder + variance predictors in 3xTF32 split
     * arithmetic on the tensor cores (fp32-grade products, fp32 accumulation). */
    int32_t tensor_core_policy;
    int32_t reserved[7];
} zvx_config;

int zvx_abi_version(void);

/* Lifetime.  zvx_create replaces ZeroVox.__init__ + .to(device) (model.py:159-257). */
int  zvx_create(const zvx_config* cfg, int device, zvx_handle** out);
void zvx_destroy(zvx_handle* h);
/* Message of the last failing call on this handle (h may be NULL for zvx_create failures). */
const char* zvx_last_error(const zvx_handle* h);

/* Weights.  One call per reference state_dict entry (load_state_dict, synthesize.py:78-88;
 * get_meldec, model.py:86-118 — vocoder keys in post-remove_weight_norm form, prefixed
 * "_meldec.").  `data` may be a host or a device pointer (fp32, contiguous, `shape[ndim]`);
 * it is copied.  Unknown keys return an error code > 0 but are otherwise harmless
 * (the reference loads with strict=False). */
int zvx_set_weight(zvx_handle* h, const char* state_dict_key, const void* data,
                   const int64_t* shape, int ndim);
/* Packs weights for the kernels (QKV fusion, tap-major conv weights, eval BatchNorm folded
 * to scale/shift, SCLN affine stack, position tables) and uploads them.  Fails listing the
 * first missing key. */
int zvx_finalize_weights(zvx_handle* h);

/* ResNetSE34V2.forward (zerovox/tts/ResNetSE34V2.py:176-212).
 * ref_mel [B, T_ref, n_mels] -> style [B, hidden] (unit L2 norm). */
int zvx_spkemb(zvx_handle* h, const float* ref_mel, int B, int T_ref, float* style, void* stream);

/* FS2Encoder.forward up to (not including) the LengthRegulator (fs2.py:732-765, 370-401,
 * 652-681).  phoneme/puncts int32 [B,T]; phoneme_mask uint8 [B,T] (1 = padding) or NULL;
 * style [B,hidden]; forced_dur int32 [B,T] or NULL (force_duration, fs2.py:745).
 * Outputs: pitch, energy, log_dur fp32 [B,T]; dur_rounded int32 [B,T]; mel_len int64 [B];
 * xprime fp32 [B,T,hidden] (features + pitch/energy embeddings, the LengthRegulator input).
 * If L_max_out != NULL the call synchronises `stream` ONCE and returns max(mel_len) there and,
 * if mel_len_host != NULL (host, int64 [B]), the per-utterance lengths — the reference pays
 * B*T + 1 such syncs (fs2.py:451, model.py:325). */
int zvx_encode(zvx_handle* h, const int32_t* phoneme, const int32_t* puncts,
               const uint8_t* phoneme_mask, const float* style, const int32_t* forced_dur,
               int B, int T, float* pitch, float* energy, float* log_dur, int32_t* dur_rounded,
               int64_t* mel_len, float* xprime, int64_t* mel_len_host, int* L_max_out, void* stream);

/* zvx_spkemb followed by zvx_encode as ONE call — ZeroVox.forward's `style_embed = self._spkemb(x["ref_mel"])` and
 * `self._phoneme_encoder(x, style_embed, ...)` (model.py:263-265).  Same results as the two calls; the speaker net is enqueued on
 * an engine-owned side stream and joins the caller's stream right before the style vector is first needed (after the encoder's
 * FFT blocks, fs2.py:740-741), so the two independent kernel sequences share the GPU.  ref_mel fp32 [B, T_ref, n_mels]; style out
 * fp32 [B, hidden] (valid in `stream` order after the call, like every other output); the rest as zvx_encode. */
int zvx_spkemb_encode(zvx_handle* h, const float* ref_mel, int T_ref, float* style, const int32_t* phoneme, const int32_t* puncts,
                      const uint8_t* phoneme_mask, const int32_t* forced_dur, int B, int T, float* pitch, float* energy,
                      float* log_dur, int32_t* dur_rounded, int64_t* mel_len, float* xprime, int64_t* mel_len_host,
                      int* L_max_out, void* stream);

/* LengthRegulator.forward + pad (fs2.py:403-459): features[b, f, :] = xprime[b, i, :] where i is
 * the phoneme whose run covers frame f, zero for f >= mel_len[b].  src_index (nullable) receives
 * i (or -1) — bit-exact integer arithmetic. */
int zvx_length_regulate(zvx_handle* h, const float* xprime, const int32_t* dur, int B, int T,
                        int L_max, float* features, int32_t* src_index, void* stream);

/* The same gather restricted to the frames [frame0, frame0 + n_frames) (long-form inputs: the features of one chunk
 * without materialising the whole sequence): features [B, n_frames, hidden], src_index [B, n_frames]. */
int zvx_length_regulate_chunk(zvx_handle* h, const float* xprime, const int32_t* dur, int B, int T,
                              int frame0, int n_frames, float* features, int32_t* src_index, void* stream);

/* FS2Decoder.forward (fs2.py:281-315) + the mel masking of ZeroVox.forward (model.py:283-285).
 * features [B,L,hidden]; mask uint8 [B,L] (1 = padding) or NULL to derive it from mel_len
 * (model.py:269-273); style [B,hidden]; zero_padded_mel != 0 applies masked_fill(mask, 0) to the
 * mel.  Outputs (either may be NULL): mel_BLC [B,L,n_mels], mel_BCL [B,n_mels,L]. */
int zvx_decode(zvx_handle* h, const float* features, const uint8_t* mask, const int64_t* mel_len,
               const float* style, int B, int L, int zero_padded_mel, float* mel_BLC,
               float* mel_BCL, void* stream);

/* hifigan.Generator.forward (hifigan.py:114-130).  mel_BCL [B,n_mels,L] -> wav [B, L*hop]. */
int zvx_vocode(zvx_handle* h, const float* mel_BCL, int B, int L, float* wav, void* stream);

/* Per-kernel-class device timing (CUDA events on the launching stream around every launch of the class),
 * used by bench.py for the roofline figures.  Enable, run any stage calls, then read (synchronises).
 * flops / bytes are the ALGORITHMIC work of the recorded launches (2*M*N*K*taps; operand + result bytes). */
#define ZVX_PROF_GEMM_FP32    0   /* fp32 FMA GEMM / implicit conv (encoder, variance predictors, bring-up) */
#define ZVX_PROF_GEMM_TC      1   /* tcgen05 TF32 GEMM / implicit conv */
#define ZVX_PROF_VOC_CONV     2   /* HiFi-GAN dilated Conv1d (+ fused lrelu / residual / MRF mean / tanh) */
#define ZVX_PROF_VOC_UPSAMPLE 3   /* HiFi-GAN polyphase ConvTranspose1d */
#define ZVX_PROF_VOC_TC       4   /* HiFi-GAN fused ResBlock (+ MRF mean) on tcgen05, C = 8/16/32 */
#define ZVX_PROF_GEMM_TC3     5   /* tcgen05 3xTF32 split GEMM (fp32-grade products: encoder, variance predictors) */
#define ZVX_PROF_NUM_CLASSES  6
int zvx_profile_enable(zvx_handle* h, int on);
int zvx_profile_read(zvx_handle* h, int kernel_class, double* ms, int64_t* launches, double* flops,
                     double* bytes);

/* Kernel-level test hook for the two contraction kernels every Linear / Conv1d / Conv2d of the acoustic
 * model and speaker net lowers to (nn.Linear, nn.Conv1d, nn.Conv2d call sites of fs2.py:143-162, 198-202,
 * 537-552 and ResNetSE34V2.py:81-92, 184-186), channel-last operands:
 *   C[m, n] = epilogue( sum_tap sum_k A[rowmap(m, tap), k] * W[tap][n, k] )
 *   mode 0 plain GEMM; mode 1 Conv1d over [M/L, L, K] ('same' padding given by pad, dilation dil);
 *   mode 2 Conv2d ksize x ksize over [IMG, Hh, Ww, K] (stride 1 or 2, padding pad), M = IMG*Ho*Wo
 *   epilogue: + bias[n]; relu_first; * scale[n] + shift[n]; + R[m, n]; relu_last.
 * use_tc = 0: fp32 FMA kernel; 1: tcgen05 TF32 kernel (fails if the layout is not TMA-addressable);
 * 2: tcgen05 3xTF32 split (fp32-grade products; the low parts of A and W are prepared internally). */
typedef struct zvx_gemm_desc {
    const float* A; const float* W; float* C;
    const float* bias; const float* scale; const float* shift; const float* R;
    int32_t M, N, K, taps, mode, L, Hh, Ww, ksize, pad, dil, relu_first, relu_last, lda, ldw, ldc;
    int32_t stride;   /* mode 2 only: 0/1 = unit stride, 2 = stride-2 Conv2d (Hh, Ww are the INPUT sizes) */
} zvx_gemm_desc;
int zvx_debug_gemm(zvx_handle* h, const zvx_gemm_desc* d, int use_tc, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Callers either side of the path.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Speaker-prompt front-end: what ZeroVoxTTS.speaker_embed (zerovox/tts/synthesize.py:123-143) does before `_spkemb`,
 * i.e. librosa.effects.trim(top_db=40) and get_mel_from_wav (zerovox/tts/mels.py:356-394), on the GPU.  Independent of
 * the model weights, hence its own handle.  Audio arguments of ZeroVoxTTS.__init__ (synthesize.py:48-60; values from
 * configs/tts_medium.yaml:3-12). */
typedef struct zvx_frontend zvx_frontend;
typedef struct zvx_mel_config {
    int32_t abi_version;    /* must be ZVX_ABI_VERSION */
    int32_t sampling_rate;  /* 22050 */
    int32_t fft_size;       /* 1024 (the only size built) */
    int32_t hop_size;       /* 256 */
    int32_t win_length;     /* 1024 (<= fft_size; centred, as librosa pads it) */
    int32_t num_mels;       /* 80 */
    float   fmin;           /* 0 */
    float   fmax;           /* 8000 */
    float   clip_val;       /* 1e-5 (dynamic_range_compression_numpy, mels.py:350); <= 0 selects 1e-5 */
    int32_t reserved[7];
} zvx_mel_config;

int  zvx_frontend_create(const zvx_mel_config* cfg, int device, zvx_frontend** out);
void zvx_frontend_destroy(zvx_frontend* h);
const char* zvx_frontend_last_error(const zvx_frontend* h);   /* h may be NULL for create failures */

/* Frames get_mel_from_wav yields for n_samples of audio: 1 + (n + 2*pad - fft_size) / hop with pad = (fft_size-hop)/2
 * (reflect padding needs n > pad; 0 otherwise). */
int64_t zvx_mel_num_frames(const zvx_frontend* h, int64_t n_samples);

/* librosa.effects.trim(wav, top_db) with its defaults frame_length = 2048, hop_length = 512 (synthesize.py:126).
 * wav fp32 [B, n_stride]; wav_len int64 [B] (device, NULL = n_stride each).  Writes the kept window per utterance to
 * start / len (device int64 [B]: samples [start, start + len)).  If start_len_host != NULL (host int64 [2*B]: B starts
 * then B lengths) the call synchronises `stream` once and returns them there, so the caller can size the mel. */
int zvx_trim_silence(zvx_frontend* h, const float* wav, int B, int64_t n_stride, const int64_t* wav_len, float top_db,
                     int frame_length, int hop_length, int64_t* start, int64_t* len, int64_t* start_len_host,
                     void* stream);

/* Sample-rate conversion of the prompt: the `sr=` argument of librosa.load in ZeroVoxTTS.get_speakerref (synthesize.py:113-121;
 * the packaged prompts are 24 kHz, the models 22.05 kHz).  Band-limited polyphase resampling with a Kaiser-windowed sinc
 * (resampy "kaiser_best" design: 64 zero crossings, beta 14.7697, roll-off 0.9476), exact taps per rational phase.  librosa's
 * default (soxr_hq) is a different filter of the same class: parity with it is a stated tolerance, not bit-exactness
 * (DESIGN.md section 2).  wav_in fp32 [B, n_in_stride], len_in device int64 [B] or NULL (= n_in_stride each);
 * wav_out fp32 [B, n_out_stride], the first n_out samples of every row are written (n_out = zvx_resample_num_samples of the
 * longest row; rows are zero-extended past their own length). */
int64_t zvx_resample_num_samples(int64_t n_in, int sr_in, int sr_out);
int zvx_resample(zvx_frontend* h, const float* wav_in, int B, int64_t n_in_stride, const int64_t* len_in, int sr_in, int sr_out,
                 float* wav_out, int64_t n_out_stride, int64_t n_out, void* stream);

/* get_mel_from_wav (mels.py:356-394), fused: reflect pad, Hann STFT magnitude, Slaney mel filterbank, log(clip), energy.
 * wav fp32 [B, n_stride]; wav_start / wav_len device int64 [B] select the window of each row (NULL = 0 / the rest of the
 * row) — the outputs of zvx_trim_silence plug in directly.  mel_BTC fp32 [B, n_frames, num_mels] (the reference's `spec`
 * transposed: exactly the `_spkemb` / zvx_spkemb input layout); energy fp32 [B, n_frames] or NULL.  Rows beyond an
 * utterance's own frame count are zero-filled.  Windows are clamped to the row: a bad (start, len) pair shortens the
 * output, it never reads outside the buffer. */
int zvx_mel_spectrogram(zvx_frontend* h, const float* wav, int B, int64_t n_stride, const int64_t* wav_start,
                        const int64_t* wav_len, int n_frames, float* mel_BTC, float* energy, void* stream);

/* Tokeniser + padding collator — host memory only, no device work.
 * Symbols (zerovox/tts/symbols.py:2-49): vocabularies given as UTF-8 strings of single code points
 * (configs/tts_medium.yaml:20-21); punct id 0 is the reserved '_NP_'. */
typedef struct zvx_symbols zvx_symbols;
int  zvx_symbols_create(const char* phones_utf8, const char* puncts_utf8, zvx_symbols** out);
void zvx_symbols_destroy(zvx_symbols* s);
const char* zvx_symbols_last_error(const zvx_symbols* s);
int  zvx_symbols_num_phones(const zvx_symbols* s);   /* Symbols.num_phones */
int  zvx_symbols_num_puncts(const zvx_symbols* s);   /* Symbols.num_puncts (includes '_NP_') */

/* ZeroVoxTTS.transcript2phonemids (synthesize.py:145-190).  Returns the number of phones; at most `capacity` entries
 * are written to phone_ids / punct_ids (host int32), so a return value > capacity means "call again with that much".
 * Negative = error (-2: a blank met a vocabulary without ' ', where the reference raises KeyError). */
int zvx_transcript2phonemids(zvx_symbols* s, const char* transcript_utf8, int32_t* phone_ids, int32_t* punct_ids,
                             int capacity);

/* collate_fn's padding (zerovox/tts/data.py:56-60, 82-83; get_mask_from_lengths fs2.py:565-573): B host sequences of
 * lens[b] <= T ids -> phoneme / puncts int32 [B, T] zero-padded, phoneme_mask uint8 [B, T] (1 = padding; nullable):
 * the zvx_encode inputs. */
int zvx_collate(const int32_t* const* phone_seqs, const int32_t* const* punct_seqs, const int32_t* lens, int B, int T,
                int32_t* phoneme, int32_t* puncts, uint8_t* phoneme_mask);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU boundary: ragged <-> padded copies around the single NCCL scatter / gather.
 * ZeroVox.forward returns zero-padded batches (model.py:260-306) of which the consumer keeps wav[i][:mel_len[i]*hop]
 * and mel[i][:, :mel_len[i]] (utils/export_hifigan.py:138-151): a shard packs only those valid parts before the gather.
 * Utterance b occupies rows * lens[b] * unit elements of `packed` starting at element offs[b], laid out
 * [rows][lens[b] * unit]; in `padded` its row r starts at b * b_stride + r * row_stride (elements).  lens / offs are
 * DEVICE int64 [B]; lens are clamped to [0, max_units].  When unit % 4 == 0 the offsets must be multiples of 4 elements
 * and the bases 16-byte aligned for the 128-bit path (else a scalar path runs).  Work is enqueued on `stream` of the
 * CURRENT device; no handle is involved.  waveform: rows = 1, unit = hop; mel [B, n_mels, L]: rows = n_mels, unit = 1.
 * ------------------------------------------------------------------------------------------------------------------ */
int zvx_ragged_pack(const float* padded, int64_t b_stride, int64_t row_stride, int rows, int unit, int B,
                    int64_t max_units, const int64_t* lens, const int64_t* offs, float* packed, void* stream);
/* The inverse; with zero_tail != 0 the elements [lens[b]*unit, max_units*unit) of every padded row are zero-filled. */
int zvx_ragged_unpack(const float* packed, const int64_t* lens, const int64_t* offs, int rows, int unit, int B,
                      int64_t max_units, float* padded, int64_t b_stride, int64_t row_stride, int zero_tail,
                      void* stream);
const char* zvx_ragged_last_error(void);

/* Fused scaled-dot-product attention of the FFT blocks (fs2.py:101-163: bmm, / temperature, masked_fill(-inf), softmax, bmm)
 * as ONE tcgen05 kernel; the engine's decoder calls the same kernel.  qk fp32 [B*L, 2*n_head*d_k] row-major (Q in the first
 * n_head*d_k columns, K in the rest; head h in columns h*d_k..), vt = V transposed per utterance: vt[(b*n_head*d_k + h*d_k + c)
 * * vt_pitch + j]; key_mask (optional, uint8 [B, L]): non-zero = key j of utterance b is masked; out fp32 [B*L, n_head*d_k].
 * d_k % 8 == 0, d_k <= 288, vt_pitch % 4 == 0, 16-byte aligned pointers.  Products in TF32, softmax and accumulation in fp32. */
int zvx_attention(const float* qk, const float* vt, int64_t vt_pitch, const uint8_t* key_mask, int B, int L, int n_head, int d_k,
                  float temperature, float* out, void* stream);
/* The same with the kernel chosen by the caller: variant 0 = as zvx_attention (the CTA-pair kernel when an utterance has at least
 * two 128-row query tiles); 1 = single-CTA kernel (tcgen05.mma.cta_group::1, operands re-streamed per key block); 2 = CTA-pair
 * kernel (two SMs of a TPC run every product as one tcgen05.mma.cta_group::2, M = 256; Q resident in shared memory, each SM
 * fetches half of every K / V block).  Both compute the same function (tests/test_gpu_attention.py runs every case on both).
 * workspace (optional, device memory, 16-byte aligned, zvx_attention_workspace_bytes(B, L, n_head) bytes): the tile list is then
 * built on the device from key_mask — key blocks past an utterance's last unmasked key are not visited (exact), a ragged batch
 * is balanced over the SMs — and with skip_masked_queries != 0 the query rows past that position are NOT computed and NOT
 * written (the FFT block zero-fills masked positions right after, fs2.py:226-229; the engine's decoder runs this way). */
int zvx_attention_ex(const float* qk, const float* vt, int64_t vt_pitch, const uint8_t* key_mask, int B, int L, int n_head, int d_k,
                     float temperature, float* out, int variant, int skip_masked_queries, void* workspace, int64_t workspace_bytes,
                     void* stream);
int64_t zvx_attention_workspace_bytes(int B, int L, int n_head);
const char* zvx_attention_last_error(void);

/* Runtime options of a handle (the release library reads no environment variables).
 *   "score_workspace_bytes": budget of the attention-score workspace; longer inputs are processed in chunks of query rows
 *                            (exact: the softmax is per row).  Default 4 GiB.  Only used when the fused kernel is off.
 *   "fused_attention":       1 (default) = the TF32 policy's attention runs as one kernel (zvx_attention); 0 = QK^T, softmax
 *                            and PV as three kernels with the score matrix in HBM / L2 (A/B and debugging); 2 / 3 = one kernel,
 *                            always the single-CTA / the CTA-pair variant (zvx_attention_ex variants 1 / 2).
 *   "pdl":                   1 = the hot kernels are launched with programmatic stream serialisation: the set-up of launch
 *                            N+1 (barriers, TMEM, shared-memory clearing) overlaps the tail of launch N; every kernel waits
 *                            for its predecessor's completion before its first global access, so results are those of plain
 *                            stream order.  0 (default) = plain launches; measured difference on configs[1]: +-0.5 %.
 *                            Process-wide. */
int zvx_set_option(zvx_handle* h, const char* name, int64_t value);

/* Workspace control: bytes of engine-owned scratch currently reserved on the device. */
int64_t zvx_workspace_bytes(const zvx_handle* h);

/* Number of kernels this library has launched through this handle since creation
 * (bench.py's gpu_launches). */
int64_t zvx_launch_count(const zvx_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ZEROVOX_B200_H */
