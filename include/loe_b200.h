/*
 * loe_b200.h -- C ABI of the B200-native hot path of loe_speech_recognition.
 *
 * The reference (loeeeee/CS-304-Speech-Recognition-Code) is pure Python and has no
 * FFI / plugin interface; its boundary is the Python package API.  This header is the
 * flat C boundary that sits directly under those classes: every entry point names the
 * reference function(s) it replaces (paths relative to src/loe_speech_recognition/).
 * The reference-side binding (ctypes) is shown in INTEGRATION.md and implemented in
 * cs-304-speech-recognition-code_b200/loe_speech_recognition/_native.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - every *_dev pointer is DEVICE memory on the current CUDA device; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream and never synchronise or allocate.
 *   - no ownership transfer: the caller owns every buffer.
 *   - return value: LOE_OK or an error code mapped 1:1 to the Python exception the
 *     reference raises at the same API point (see loe_status).  loe_last_error()
 *     returns a thread-local message for the last failing call.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns
 *     LOE_ERR_CUDA.
 */
#ifndef LOE_B200_H
#define LOE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum loe_status {
    LOE_OK = 0,
    LOE_ERR_CUDA = 1,        /* RuntimeError: CUDA failure (message has the cudaError string)      */
    LOE_ERR_VALUE = 2,       /* ValueError: bad argument (e.g. < 9 frames for the delta filter,
                                mfcc.py:39 -> scipy savgol_filter)                                    */
    LOE_ERR_OVERFLOW = 3,    /* OverflowError: more than 128 trellis positions; the reference's
                                tracer/path are int8 (hidden_markov_model.py:175, 498-501)          */
    LOE_ERR_DIM = 4,         /* AssertionError: feature dimension mismatch (hidden_markov_model.py:76-77) */
    LOE_ERR_UNSUPPORTED = 5  /* NotImplementedError: configuration outside the built kernels       */
} loe_status;

#define LOE_ABI_VERSION 1
#define LOE_MAX_POS 128          /* trellis positions per utterance (int8 path, see above)          */
#define LOE_N_MELS 40
#define LOE_N_FFT 320
#define LOE_HOP 160
#define LOE_N_BINS 161
#define LOE_PCM_F32 0            /* float32 samples at int16 scale (what the reference passes)       */
#define LOE_PCM_S16 1            /* int16 samples as stored in a WAV file: half the PCIe / HBM bytes  */
#define LOE_MEL_NA_MAX 32        /* mel lane table: max iterations of round A / round B             */
#define LOE_MEL_NB_MAX 16

/* trellis position flags (loe_viterbi_dev / loe_align_dev) */
#define LOE_POS_INIT 1           /* receives logpdf(x_0) + self-loop at t = 0                        */
#define LOE_POS_START 2          /* word start of the loop grammar (cross-word rule applies)         */
#define LOE_POS_END 4            /* termination candidate / word end                                 */

int loe_abi_version(void);
const char* loe_last_error(void);
/* number of CUDA devices visible, or a negative loe_status */
int loe_device_count(void);

/* --------------------------------------------------------------------------------------
 * MFCC front end.  Replaces MFCC.__post_init__ / MFCC.normalize_mfccs / MFCC.batch
 * (mfcc.py:24-44, 50-69, 71-84) and the librosa calls behind them (melspectrogram
 * n_fft=320 hop=160 periodic Hann centre-padded, slaney mel 40 bands, power_to_db with the
 * per-utterance maximum as reference and top_db=80, DCT-II ortho 13 ceps, Savitzky-Golay
 * delta / delta-delta of width 9, per-frame normalisation of the static block).
 *
 *   pcm_dev      [total_samples] PCM, utterances back to back; pcm_format = LOE_PCM_F32 (float32 at
 *                int16 scale, the dtype the reference feeds librosa, ti_digits.py:133) or LOE_PCM_S16
 *                (raw int16 as read by scipy.io.wavfile; converted on load, bit-identical results)
 *   pcm_off_dev  [n_utt+1] int64 sample offsets
 *   frm_off_dev  [n_utt+1] int64 frame offsets, frames(u) = 1 + samples(u)/160
 *   max_frames   max_u frames(u)            min_frames  min_u frames(u) (must be >= 9)
 *   mel_bin_dev [(na+nb)*32] int32, mel_w_dev [(na+nb)*32] float32: the slaney filterbank as a
 *                lane-balanced table.  Entry (it, lane) adds mel_w * power[mel_bin] to a partial sum:
 *                iterations [0, na): lane l accumulates filter l (filters 0..31);
 *                iterations [na, na+nb): lanes 4q..4q+3 share filter 32+q (non-zero j of that filter
 *                goes to lane 4q + j%4, iteration na + j/4).  Unused entries have weight 0.  A lane's
 *                bins must be consecutive (round A: bin(it) = bin(0) + it; round B: bin(it) = bin(0) + 4 it):
 *                the kernel reads only mel_bin of each lane's first entry.
 *   mel_ws_dev   [total_frames*40] float32 workspace (mel energies), 16-byte aligned
 *   utt_max_dev  [n_utt] float32 workspace (per-utterance mel maximum)
 *   feat_dev     [total_frames*39] float32 out, row-major (frame, coefficient): the
 *                transposed (T,39) layout MFCC.batch hands to the HMM code
 * -------------------------------------------------------------------------------------- */
int loe_mfcc_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                 int n_utt, int64_t total_frames, int max_frames, int min_frames,
                 const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                 float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream);

/* The two kernels of loe_mfcc_dev separately (profiling / timing): phases bit 0 = PCM -> mel energies
 * + utterance maxima, bit 1 = mel -> features.  loe_mfcc_dev == phases 3. */
int loe_mfcc_phase_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                       int n_utt, int64_t total_frames, int max_frames, int min_frames,
                       const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                       float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream, int phases);

/* loe_mfcc_phase_dev with the A-operand image of the 3xFP16 emission kernel as an additional output of phase 2 (see
 * loe_emission_h16_img_dev); feat_dev may be NULL when only the image is wanted (the decode path). */
int loe_mfcc_img_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                     int n_utt, int64_t total_frames, int max_frames, int min_frames,
                     const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                     float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* a_img_dev, float* inv2_dev, void* stream, int phases);

/* --------------------------------------------------------------------------------------
 * Parameterised MFCC front end (csrc/mfcc_ex.cu).  The reference hard-wires one parameter set (mfcc.py:31-34), which
 * loe_mfcc_dev serves; this entry point is the same pipeline (mfcc.py:24-44) with the stages made parameters, for
 * BASELINE.json configs[3] (25 ms / 10 ms frames, 512-point FFT, Hamming, pre-emphasis 0.97, 40 mel, 13 ceps + deltas,
 * cepstral mean normalisation).  No live reference call site: parity is against the restated oracle only.
 *
 *   frames(u) = 1 + samples(u) / hop, frame t covers samples [hop t - n_fft/2, hop t + n_fft/2), zeros outside
 *   y[i] = x[i] - preemph * x[i-1]  (x[-1] := x[0]);  window_dev [n_fft] (a shorter window zero-padded to n_fft, centred)
 *   filter m = sum_i mel_w_dev[m * mel_pitch + i] * power[mel_start_dev[m] + i], i < mel_len_dev[m]
 *   log_mode LOE_LOG_DB: 10 log10(max(1e-10, mel)) - 10 log10(max over the utterance), floored at -80 (power_to_db
 *            ref=np.max);  LOE_LOG_LN: ln(max(1e-10, mel))
 *   dct_dev [n_ceps * n_mels] (row k = coefficient k);  deltas: Savitzky-Golay width 9 on the raw cepstra
 *   norm_mode (static block only): NONE, FRAME ((c - mean_k c) / (std_k c + 1e-8) inside each frame, mfcc.py:62-66),
 *            CMN (c - mean over the utterance's frames), CMVN (CMN / (std over the frames + 1e-8))
 *   workspaces: mel_ws_dev [total_frames * n_mels], ceps_ws_dev [total_frames * n_ceps],
 *               utt_stat_dev [n_utt * 2 * n_ceps]  (float32)
 *   feat_dev [total_frames * 3 * n_ceps] out, row-major (frame, coefficient)
 * -------------------------------------------------------------------------------------- */
#define LOE_LOG_DB 0
#define LOE_LOG_LN 1
#define LOE_NORM_NONE 0
#define LOE_NORM_FRAME 1
#define LOE_NORM_CMN 2
#define LOE_NORM_CMVN 3
typedef struct loe_mfcc_config {
    int32_t n_fft;       /* power of two, 64 .. 1024 */
    int32_t hop;
    int32_t n_mels;      /* <= 64 */
    int32_t n_ceps;      /* <= 16 */
    int32_t log_mode;    /* LOE_LOG_* */
    int32_t norm_mode;   /* LOE_NORM_* */
    float preemph;       /* 0 = none */
    float reserved;
} loe_mfcc_config;
int loe_mfcc_ex_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                    int n_utt, int64_t total_frames, int max_frames, int min_frames, const loe_mfcc_config* cfg,
                    const float* window_dev, const int32_t* mel_start_dev, const int32_t* mel_len_dev,
                    const float* mel_w_dev, int mel_pitch, const float* dct_dev, float* mel_ws_dev, float* ceps_ws_dev,
                    float* utt_stat_dev, float* feat_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * Gaussian emission scoring.  Replaces MultivariateNormal.log_pdf
 * (hidden_markov_model.py:46-48 -> scipy multivariate_normal_frozen.logpdf) for every
 * (frame, state) pair at once:
 *     out[f*ld_out + s] = cst[s] - 0.5 * | (x_f - mean_s) . U_s |^2
 * with U_s = V diag(lambda^-1/2) (scipy _PSD.U) and cst = -0.5*(D log 2pi + log_pdet).
 *
 *   feat_dev [n_frames*dim] float32;  mean_dev [n_states*dim];  u_dev [n_states*dim*dim]
 *   (row i, column j at u[(s*dim+i)*dim+j]);  cst_dev [n_states]
 *   precision: 0 = float32 SIMT, 1 = float64 SIMT (mean/u/cst are then double arrays)
 * -------------------------------------------------------------------------------------- */
int loe_emission_dev(const float* feat_dev, int64_t n_frames, int dim,
                     const void* mean_dev, const void* u_dev, const void* cst_dev, int n_states,
                     float* out_dev, int ld_out, int precision, void* stream);

/* Same result on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in TMEM; dim == 39).
 * The model is pre-packed on the host into the shared-memory image the kernel keeps resident:
 * states are grouped in tiles of 6, tile t owns floats [t*19200, (t+1)*19200) of b_packed:
 *     b_packed[t*19200 + h*9600 + kc*960 + n*4 + q] = part_h( W_s[4*kc + q][j] ),
 *     s = 6*t + n/40, j = n%40, h = 0 (TF32-rounded value) / 1 (residual),
 *     W_s = [ U_s ; -mean_s . U_s ] (40 x 39, column 39 and states >= n_states are zero)
 * cst_pad_dev has loe_emission_tc_tiles(n_states)*6 entries (zero padded). */
int loe_emission_tc_tiles(int n_states);
int loe_emission_tc_dev(const float* feat_dev, int64_t n_frames, int dim, const float* b_packed_dev,
                        const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream);

/* 3xFP16 variant of loe_emission_tc_dev (csrc/emission_h16.cu): the operands are split into two
 * binary16 parts (the same 22 significant bits as the TF32 pair) and run as kind::f16 MMAs, 8 per
 * 128-frame tile against 15, six of them over 48 to 192 of the 240 columns -- more than twice the throughput at the
 * same measured accuracy.
 * b_packed_dev: per tile of 6 states loe_emission_h16_tile_bytes() (= 57600) bytes of binary16,
 *   [chunk c (15)][n = (j/8)*48 + state_local*8 + j%8 (240)][q (8)] with W_s[k][j], k = 8*(c % 5) + q:
 *   chunks 0-4 = fp16(W), chunks 5-9 = fp16(W - fp16(W)), chunks 10-14 = chunks 0-4 again.
 *   W_s = [ R_s^T ; -mean_s . R_s^T ] with U_s^T = Q R_s (any W with W W^T = U_s U_s^T scores alike):
 *   W_s MUST be lower triangular in its first 39 rows (W_s[k][j] = 0 for j > k) -- the kernel multiplies the
 *   features 8c .. 8c+7 with the columns j < 8 (c + 1) only -- and column 39 and states >= n_states zero.
 * Domain: every |W| must be below 32768 (the host packer checks and the caller uses the TF32 image
 * otherwise); feature rows of any magnitude are accepted (rows reaching 2^15 are rescaled by an exact
 * power of two inside the kernel).  cst_pad_dev as for loe_emission_tc_dev. */
int loe_emission_h16_tile_bytes(void);
int loe_emission_h16_dev(const float* feat_dev, int64_t n_frames, int dim, const void* b_packed_dev,
                         const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream);
/* The same scores from the PRE-SPLIT operand (decode path): loe_mfcc_img_dev makes the cepstrum kernel write every feature
 * row as the binary16 hi / lo A operand (tile-major image: per 128 frames 20 480 bytes = [chunk (10: hi 0-4, lo 5-9)][row (128)]
 * [8 halfs], rows scaled by 2^-e when they reach 2^15) plus the row scales inv2 = 4^e; the emission kernel then bulk-copies
 * tile after tile into its A stages and has no producer work.  Bit-identical to loe_emission_h16_dev on the same features.
 *   a_img_dev: loe_emission_h16_img_bytes(total_frames) bytes, 16-byte aligned;  inv2_dev: ceil(total_frames / 128) * 128 floats */
int64_t loe_emission_h16_img_bytes(int64_t n_frames);
int loe_emission_h16_img_dev(const void* a_img_dev, const float* inv2_dev, int64_t n_frames, const void* b_packed_dev,
                             const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream);
/* Several models in ONE launch (batched training, where the reference trains its word models one after the other,
 * hidden_markov_model.py:294-318): segment i scores the frames [seg_begin[i], seg_end[i]) of feat_dev with the
 * seg_states[i] (<= 12 = max_states bound) states whose image starts at tile seg_tile[i] of b_packed_dev / cst_pad_dev,
 * into the columns from seg_col[i] of out_dev.  active_dev (may be NULL): segments whose flag is not 1 are skipped on the
 * device (the M-step freezes converged models there, loe_mstep_dev).  All seg_* arrays are device arrays. */
int loe_emission_h16_multi_dev(const float* feat_dev, int dim, const void* b_packed_dev, const float* cst_pad_dev, int n_seg,
                               const int64_t* seg_begin_dev, const int64_t* seg_end_dev, const int32_t* seg_tile_dev,
                               const int32_t* seg_states_dev, const int32_t* seg_col_dev, const int32_t* active_dev,
                               int max_states, float* out_dev, int ld_out, void* stream);

/* The multi-model launch from pre-split images (training: the features do not change between iterations, so their A
 * operand is built ONCE by loe_h16_image_dev and every iteration bulk-copies it).  Segment i owns ceil(frames_i / 128)
 * image tiles from tile seg_img_tile[i] (rows past its end are zero) and the inv2 rows from 128 * seg_img_tile[i];
 * n_img_tiles = sum of the segments' tiles; a_img_dev: n_img_tiles * 20 480 bytes, inv2_dev: n_img_tiles * 128 floats. */
int loe_h16_image_dev(const float* feat_dev, int dim, int n_seg, const int64_t* seg_begin_dev, const int64_t* seg_end_dev,
                      const int32_t* seg_img_tile_dev, int n_img_tiles, void* a_img_dev, float* inv2_dev, void* stream);
int loe_emission_h16_multi_img_dev(const void* a_img_dev, const float* inv2_dev, const void* b_packed_dev, const float* cst_pad_dev,
                                   int n_seg, const int64_t* seg_begin_dev, const int64_t* seg_end_dev,
                                   const int32_t* seg_img_tile_dev, const int32_t* seg_tile_dev, const int32_t* seg_states_dev,
                                   const int32_t* seg_col_dev, const int32_t* active_dev, int max_states, float* out_dev,
                                   int ld_out, void* stream);

/* --------------------------------------------------------------------------------------
 * Diagonal-covariance Gaussian-mixture emission scoring (csrc/emission_gmm.cu; BASELINE.json north_star kernel (2),
 * configs[0] extension set and configs[4]).  The live reference scores one full-covariance Gaussian per state
 * (hidden_markov_model.py:20-48); the semantics of the mixture follow its abandoned deprecated/gaussian_mixture_model.py:157-162:
 *     out[f*ld_out + s] = log sum_m exp( cst[s,m] - 0.5 * sum_k (x_fk - mean[s,m,k])^2 * inv_var[s,m,k] )
 *     cst[s,m] = log w[s,m] - 0.5 * (D log 2pi + sum_k log var[s,m,k])
 * No live reference call site: parity is against the restated oracle (oracle/gmm.py) only.
 *
 * loe_emission_gmm_dev: SIMT.  mean_dev / inv_var_dev [n_states*n_mix*dim], cst_dev [n_states*n_mix]; precision 0 =
 *   float32 arrays, 1 = float64 arrays (the exact mode).  dim == 39, n_mix <= 16.
 * loe_emission_gmm_tc_dev: the same result as ONE dense contraction on the tcgen05 tensor cores,
 *     y[f, n] = [z^2 (39), 1, z (39), 0] . B[:, n],  z = (x - shift) * iscale,  n = state_local * MP + mixture
 *   (MP = n_mix padded to a power of two; a column tile holds 240 / MP states), 3 x binary16 split operands with fp32
 *   accumulation in TMEM, log-sum-exp over the mixtures in the epilogue thread of the frame.
 *   b_packed_dev: per tile loe_emission_gmm_tile_bytes() (= 76800) bytes of binary16, [chunk c (20)][n (240)][q (8)]:
 *     chunks 0-9 = fp16(B[8c + q][n]), chunks 10-19 = fp16(B - fp16(B)) of rows 8(c-10) + q;
 *     B[k < 39] = -t_k^2 / (2 var_k), B[39] = cst - 0.5 sum (mean - shift)^2 / var, B[40 + k] = t_k (mean_k - shift_k) / var_k,
 *     B[79] = 0; unused columns zero.
 *   shift_scale_dev [tiles * 80] float32: per tile shift[40] then iscale[40] = 1 / t (t_k powers of two).
 *   mean32_dev / inv_var32_dev / cst32_dev: the float32 arrays of loe_emission_gmm_dev -- frame rows with |z| >= 128 or
 *     non-finite values are evaluated from them by plain float32 arithmetic inside the same kernel.
 * -------------------------------------------------------------------------------------- */
int loe_emission_gmm_tile_bytes(void);
int loe_emission_gmm_tiles(int n_states, int n_mix);
int loe_emission_gmm_dev(const float* feat_dev, int64_t n_frames, int dim, const void* mean_dev, const void* inv_var_dev,
                         const void* cst_dev, int n_states, int n_mix, float* out_dev, int ld_out, int precision, void* stream);
int loe_emission_gmm_tc_dev(const float* feat_dev, int64_t n_frames, int dim, const void* b_packed_dev,
                            const float* shift_scale_dev, const float* mean32_dev, const float* inv_var32_dev,
                            const float* cst32_dev, int n_states, int n_mix, float* out_dev, int ld_out, void* stream);

/* --------------------------------------------------------------------------------------
 * Viterbi + backtrace, one CTA per utterance.  Replaces HiddenMarkovModel._viterbi /
 * _viterbi_static (hidden_markov_model.py:80-91, 160-208), HiddenMarkovModelInference.
 * _viterbi / _viterbi_static (:463-581) and the forced alignment of
 * HiddenMarkovModelMultiWord (:591), bit for bit (float32 recursion, lowest-index
 * tie-breaking, "all -inf -> back-pointer 0", off-by-one backtrace, float64 word-penalty mode).
 *
 * A trellis is a list of positions; trellis k owns positions [tr_off[k], tr_off[k+1]).
 *   col   [n_pos_total] int32   emission column of the position in `scores`
 *   band  [n_pos_total*3] float32 log-transition into p from p, p-1, p-2 (-inf = not allowed)
 *   flags [n_pos_total] uint8   LOE_POS_* bits
 *   utt_tr_dev [n_utt] int32 trellis of each utterance (NULL: all use trellis 0)
 *   loop: 0 = left-to-right only; 1 = digit-loop grammar (START positions additionally take
 *         max over END positions + penalty, ties -> lowest END, self loop last)
 *   penalty / penalty_f64: word-transition log penalty; penalty_f64 != 0 reproduces the
 *         reference when the attribute is an np.float64 (the default np.log(0.005))
 *   scores_dev [total_frames*ld] float32 from loe_emission_dev
 *   path_dev   [total_frames] int8 out (position index local to the trellis; -1 when T == 1)
 *   end_scores_dev [n_utt*max_ends] float32 out (may be NULL), END positions in order
 *   best_dev   [n_utt] int32 out: index (among END positions) of the best end
 *   best_score_dev [n_utt] float32 out
 *   bp_ws_dev  workspace of total_frames*LOE_MAX_POS bytes, only used when the back-pointers
 *              of the longest utterance do not fit in shared memory (may be NULL otherwise;
 *              loe_viterbi_bp_fits tells)
 *   word_dev / word_lo_dev / skip_label / words_dev / max_words / count_dev: optional fused
 *              label decoding with the semantics of loe_labels_dev (words_dev == NULL: off)
 * -------------------------------------------------------------------------------------- */
int loe_viterbi_bp_fits(int max_frames, int max_pos);
int loe_viterbi_dev(const float* scores_dev, int ld, const int64_t* frm_off_dev, int n_utt, int max_frames,
                    const int32_t* tr_off_dev, const int32_t* col_dev, const float* band_dev,
                    const uint8_t* flags_dev, int max_pos, const int32_t* utt_tr_dev,
                    int loop, double penalty, int penalty_f64,
                    int8_t* path_dev, float* end_scores_dev, int max_ends,
                    int32_t* best_dev, float* best_score_dev, uint8_t* bp_ws_dev,
                    const int32_t* word_dev, const int32_t* word_lo_dev, int skip_label,
                    int8_t* words_dev, int max_words, int32_t* count_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * State path -> word sequence.  Replaces ModelBoundary.get_labels / append_to_labels
 * (model_boundary.py:107-147) for a whole batch: run-length compress the path, emit a word
 * when the path leaves the current word's state range or re-enters its first state from its
 * last state (repeated word); words whose label id equals skip_label (silence) are dropped.
 *   words_dev [n_utt*max_words] int8 out: label ids (word_dev values) in order
 *   count_dev [n_utt] int32 out: number of words (may exceed max_words: the caller re-derives
 *             those utterances on the host), or -1 when the path holds a negative state
 *             (T == 1: the reference raises there)
 * -------------------------------------------------------------------------------------- */
int loe_labels_dev(const int8_t* path_dev, const int64_t* frm_off_dev, int n_utt,
                   const int32_t* tr_off_dev, const int32_t* word_dev, const int32_t* word_lo_dev,
                   const int32_t* utt_tr_dev, int skip_label,
                   int8_t* words_dev, int max_words, int32_t* count_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * Energy-hysteresis silence stripper (the step before MFCC in the reference's silence-model training,
 * scripts/project5_train_no_empty.py:18-31).  Replaces SignalSeparation._remove_empty / detect_speech
 * (signal_separation.py:103-164) bit for bit (NumPy's float32 pairwise summation order for the frame
 * energies, float64 thresholds = high/low * max|x|, the hysteresis state machine and its quirks).
 *   efrm_off_dev [n_utt+1] int64: offsets into energy/noise; utterance u has samples/frame_size full
 *                frames plus ONE trailing partial frame (possibly empty)
 *   energy_dev [total] float32 out: mean |x| per frame (NaN for an empty trailing frame)
 *   noise_dev  [total] uint8 out: 1 = frame the reference appends to its noise list
 *   seg_dev    [n_utt*4] int32 out: {done, start, end, n_frames}: the stripped signal is frames
 *                [start, end); done = 0 where the reference raises FailToProcess (speech never ended)
 *   max_dev    [n_utt] float32 out: max |x|
 * -------------------------------------------------------------------------------------- */
int loe_silence_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, int n_utt,
                    int frame_size, double high, double low, int max_silence_frames,
                    const int64_t* efrm_off_dev, float* energy_dev, uint8_t* noise_dev, int32_t* seg_dev,
                    float* max_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * Time-synchronous template DTW.  Replaces DynamicTimeWarping.search (dynamic_time_wrapping.py:66-116)
 * for a batch of samples against one concatenated template set, bit for bit (float32 local distance in
 * NumPy's summation order, float64 accumulated cost, beam pruning, and the reference's row-sharing /
 * wrap-around / read-one-row-early quirks -- see csrc/dtw.cu).
 *   seq_dev [n_rows*dim] template features back to back; word w owns frames [starts[w], starts[w]+lens[w])
 *   row_start_dev [n_rows+1]: for cost row i >= 1 the start of the word that holds frame i-1;
 *   row_is_boundary_dev [n_rows+1]: 1 where i == starts[w] for a word w > 0
 *   samp_dev [total_sample_frames*dim], samp_off_dev [n_samples+1] int64
 *   dist_dev [n_samples*n_words] float64 out; best_idx_dev / best_dist_dev [n_samples] out (first minimum)
 *   cost_out_dev / path_out_dev: optional (n_rows+1) x (L+1) matrices of sample 0 (may be NULL)
 * -------------------------------------------------------------------------------------- */
int loe_dtw_dev(const float* seq_dev, int n_rows, int dim, const int32_t* row_start_dev,
                const int32_t* row_is_boundary_dev, const int32_t* starts_dev, const int32_t* lens_dev,
                int n_words, const float* samp_dev, const int64_t* samp_off_dev, int n_samples,
                int pruning, double pruning_factor, double* dist_dev, int32_t* best_idx_dev,
                double* best_dist_dev, double* cost_out_dev, int8_t* path_out_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * Segmental K-means sufficient statistics.  Replaces Signal.order_by_state,
 * SortedSignals.order_by_state / .transition_probabilities (signal.py:23-47, 68-91), the
 * accumulation half of HiddenMarkovModelTrainable._update_middleware_parameters
 * (hidden_markov_model.py:320-350) and HiddenMarkovModelMultiWord._remux_path_and_signal
 * (:602-636).
 *
 * loe_align_dev: turns alignments into per-frame bucket ids (global state of the word model
 * the frame is credited to, 0xFFFF = not credited) and transition counts.
 *   word_dev [n_pos_total] int32 word-instance label id of each trellis position (positions
 *            of one word instance are contiguous); word_lo_dev [n_pos_total] int32 first
 *            position (local to the trellis) of the instance the position belongs to
 *   remux: 0 = the whole utterance is one Signal (isolated training);
 *          1 = cut where the label changes, re-base, and DROP the final piece (:614-636)
 *   bucket_dev [total_frames] uint16 out;  counts_dev [n_glob*n_glob] int32, accumulated
 *            (caller zeroes): counts[g_from*n_glob + g_to]
 * loe_kmeans_dev: per bucket g:  stats[g*stride + 0] = N,  [1..D] = sum(x - shift_g),
 *   then the upper triangle (row-major, i <= j) of sum (x-shift_g)(x-shift_g)^T, all float64,
 *   stride = 1 + D + D(D+1)/2.  Deterministic (fixed chunking and reduction order).
 *   shift_dev [n_glob*dim] float32 (typically the previous means);  part_ws_dev workspace of
 *   loe_kmeans_ws_doubles(total_frames, n_glob, dim) doubles.
 * -------------------------------------------------------------------------------------- */
int loe_align_dev(const int8_t* path_dev, const int64_t* frm_off_dev, int n_utt,
                  const int32_t* tr_off_dev, const int32_t* col_dev, const int32_t* word_dev,
                  const int32_t* word_lo_dev, const int32_t* utt_tr_dev, int remux, int n_glob,
                  uint16_t* bucket_dev, int32_t* counts_dev, void* stream);
int64_t loe_kmeans_ws_doubles(int64_t total_frames, int n_glob, int dim);
int loe_kmeans_dev(const float* feat_dev, const uint16_t* bucket_dev, int64_t total_frames, int dim,
                   int n_glob, const float* shift_dev, double* part_ws_dev, double* stats_dev, void* stream);

/* --------------------------------------------------------------------------------------
 * Segmental K-means M-step on the device (csrc/mstep.cu).  Replaces, for every word model of a batched training run,
 * HiddenMarkovModelTrainable._update_middleware_parameters (hidden_markov_model.py:320-350: means, the means-only
 * convergence test np.allclose(new, old) BEFORE covariances / transitions are touched, np.cov (N-1) + 1e-3 I, transition
 * counts / row sum) and _update_inference_weights (:283-292: log-transitions, the Gaussians' whitening data), from the
 * statistics of loe_kmeans_dev (shifted by the current means) and the counts of loe_align_dev.  Word w owns the global
 * states [word_first[w], word_first[w] + word_n[w]) and the 6-state image tiles from word_tile[w].
 *   means32_dev [n_glob*39] in/out (float32 model means: shift of the statistics on entry, new means where updated)
 *   cov32_dev [n_glob*39*39], counts_applied_dev [n_glob*n_glob]: written for updated words only (a converged word keeps
 *       the previous values, like the reference)
 *   band_dev [n_glob*3]: log-transition into state g from g, g-1, g-2 (the per-word trellises of loe_viterbi_dev)
 *   b_h16_dev / cst_pad_dev: the 3xFP16 image of loe_emission_h16_dev, rewritten in place for updated words
 *   active_dev [n_words] in/out: 1 = training, 0 = converged (frozen), -1 = failed (empty state), -2 = parked (suspect
 *       covariance: no image was written; the caller repairs the image and sets the word back to 1)
 *   updated_dev [n_words] scratch;  status_dev [n_words] out: LOE_MSTEP_* bits of THIS call
 * The float32 means / covariances / transition probabilities equal the host M-step's bit for bit given the same
 * statistics; the whitening matrix is the inverse Cholesky factor of the (reversed) covariance instead of scipy's
 * eigenvector form (same quadratic form).  LOE_MSTEP_SUSPECT: the covariance is not finite, fails a pivot test or its
 * whitening matrix leaves the binary16 range -- the caller then applies scipy's own test on the host.
 * -------------------------------------------------------------------------------------- */
#define LOE_MSTEP_UPDATED 1
#define LOE_MSTEP_CONVERGED 2
#define LOE_MSTEP_MEAN_FAIL 4
#define LOE_MSTEP_SUSPECT 8
int loe_mstep_dev(const double* stats_dev, const int32_t* counts_dev, int n_glob, int n_words,
                  const int32_t* state_word_dev, const int32_t* word_first_dev, const int32_t* word_n_dev,
                  const int32_t* word_tile_dev, float* means32_dev, float* cov32_dev, int32_t* counts_applied_dev,
                  float* band_dev, void* b_h16_dev, float* cst_pad_dev, int32_t* active_dev, int32_t* updated_dev,
                  int32_t* status_dev, int dim, void* stream);

/* --------------------------------------------------------------------------------------
 * Host-buffer decoder: the whole hot path behind ONE call that takes HOST memory and needs no
 * torch / Python on the caller's side.  Replaces, for a batch of utterances,
 *     [HiddenMarkovModelInference.predict(MFCC(sig, sr).feature_vector) for sig in signals]
 * (mfcc.py:24-44 + hidden_markov_model.py:458-581 + model_boundary.py:107-147).
 *
 * loe_decoder_create uploads the model tables once (mel lane tables as loe_mfcc_dev reads them,
 * the tensor-core Gaussian image as loe_emission_tc_dev reads it, ONE loop-grammar trellis as
 * loe_viterbi_dev reads it; all pointers are HOST pointers) and owns two streams plus a device
 * workspace that grows on demand and is reused by later calls.
 *
 * loe_decoder_decode_host: pcm_host holds the utterances back to back (float32 or int16, see
 * pcm_format), utterance i = samples [sample_off[i], sample_off[i+1]).  The batch is cut into
 * n_chunks chunks of whole utterances (n_chunks <= 0: about 64 MB each, at most 64); the copy of
 * chunk c+1 overlaps the kernels of chunk c (pinned pcm_host makes that copy asynchronous; pageable
 * memory works, without the overlap).  Outputs (HOST): words [n_utt*max_words] int8 word ids,
 * count [n_utt] int32 (count > max_words or < 0: decode that utterance's path on the host, as
 * loe_labels_dev documents), optional best_score [n_utt] float32 and path [total_frames] int8.
 * The call returns after the results have landed.  One call at a time per decoder.
 * -------------------------------------------------------------------------------------- */
int loe_decoder_create(int device, const int32_t* mel_bin_host, const float* mel_w_host, int mel_na, int mel_nb,
                       const float* b_packed_host, const float* cst_pad_host, int n_states,
                       int n_pos, const int32_t* col_host, const float* band_host, const uint8_t* flags_host,
                       const int32_t* word_host, const int32_t* word_lo_host, void** decoder_out);
int loe_decoder_decode_host(void* decoder, const void* pcm_host, int pcm_format, const int64_t* sample_off_host, int n_utt,
                            double penalty, int penalty_f64, int skip_label, int max_words, int n_chunks,
                            int8_t* words_host, int32_t* count_host, float* best_score_host, int8_t* path_host);
/* Optional: the 3xFP16 Gaussian image of the same model (HOST pointer, layout of loe_emission_h16_dev,
 * ceil(n_states / 6) * loe_emission_h16_tile_bytes() bytes).  When set, decode scores with
 * loe_emission_h16_dev instead of loe_emission_tc_dev -- the choice the Python package makes for models
 * whose whitening matrices fit the binary16 range.  NULL removes it again. */
int loe_decoder_set_h16(void* decoder, const void* b_h16_host);
void loe_decoder_destroy(void* decoder);
/* float32 PCM that holds int16 values (the reference converts WAV samples to float32 on the host,
 * ti_digits.py:85-139) can be narrowed back to int16 by worker threads inside loe_decoder_decode_host before it
 * crosses PCIe: every sample is verified (converted back and compared), a chunk with one inexact sample
 * travels as float32, and the features are bit-identical either way.
 *   workers: hardware threads / LOCAL_WORLD_SIZE (the ranks sharing the box, as torchrun exports it), at most 32,
 *            bound to the CPUs next to the GPU (sysfs local_cpulist) when readable; LOE_B200_NARROW_THREADS
 *            overrides the count (0 = never narrow), LOE_B200_NARROW_PIN=0 leaves them floating.
 *   loe_decoder_set_narrow(mode): LOE_NARROW_ON / LOE_NARROW_OFF make it the caller's explicit choice (also
 *            LOE_B200_NARROW=on|off in the environment when the decoder is created).  LOE_NARROW_AUTO (default):
 *            the decoder narrows while it samples, per chunk, its conversion rate (host clock) and the wire rate
 *            of its host->device copies (CUDA events); after 6 chunks of >= 2^20 samples each it keeps narrowing
 *            iff median conversion GB/s (float32 bytes) > 1.25 x median copy GB/s -- a pipelined chunk period gets shorter
 *            when the conversion beats the copy -- re-examines that on every later call (medians of the last 64 chunks)
 *            and switches off, for the decoder's lifetime, once the margin falls below 1.05.
 *   loe_decoder_stats: out[0..LOE_DECODER_STATS) = {mode, narrowing on (1) / off (0) / still sampling (-1),
 *            median conversion GB/s (float32 bytes), median copy GB/s (wire bytes), worker threads, CPUs the workers
 *            are bound to (0 = floating), PCM bytes of the last call as the caller holds them, bytes that crossed
 *            PCIe host->device in the last call, chunks of the last call, chunks of it that travelled as int16}.
 * loe_pcm_narrow_host: the same conversion as a stand-alone call (single thread): writes dst_host[i] =
 * (int16) src_host[i] and returns 1 if every sample was exact, 0 otherwise (dst contents then unspecified).
 * loe_decoder_narrow_rate: median conversion GB/s measured so far (negative: narrowing is off, 0: no sample yet). */
#define LOE_NARROW_AUTO (-1)
#define LOE_NARROW_OFF 0
#define LOE_NARROW_ON 1
#define LOE_DECODER_STATS 10
int loe_decoder_set_narrow(void* decoder, int mode);
int loe_decoder_stats(void* decoder, double* out, int n);
int loe_pcm_narrow_host(const float* src_host, int16_t* dst_host, int64_t n_samples);
double loe_decoder_narrow_rate(void* decoder);
/* page-locked host buffers for pcm_host (cudaHostAlloc / cudaFreeHost) */
int loe_host_alloc(void** ptr_out, size_t bytes);
int loe_host_free(void* ptr);

/* 64-bit content fingerprint of n_blocks host memory blocks (no CUDA call).  The Python layer keeps the device copy of a
 * model next to the reference's own objects (scipy frozen distributions + the dict-backed transition table,
 * hidden_markov_model.py:20-48, transition_probability.py:11-40) and asks before every single-utterance call whether they
 * were edited in place: the reference reads them afresh on every predict (hidden_markov_model.py:481-531). */
uint64_t loe_host_fingerprint(const void* const* blocks, const int64_t* n_bytes, int n_blocks);

/* Word-id tables (loe_viterbi_dev / loe_labels_dev / loe_decoder_decode_host) -> text, on the host: the "".join(labels) of
 * HiddenMarkovModelInference.predict (hidden_markov_model.py:458-461, model_boundary.py:141-147) for a whole batch of
 * single-character labels.  Utterance i writes its first count[i] labels and then sep; a count outside [0, max_words]
 * writes sep alone (the caller decodes those utterances from the state path).  Returns the bytes written
 * (<= n_utt * (max_words + 1)). */
int64_t loe_labels_text_host(const int8_t* words_host, const int32_t* count_host, int n_utt, int max_words,
                             const char* label_chars, int n_labels, char sep, char* out_host);

#ifdef __cplusplus
}
#endif
#endif /* LOE_B200_H */
