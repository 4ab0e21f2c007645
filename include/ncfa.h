/*
 * ncfa.h — C ABI of the B200-native nightcore front-end (libncfa.so).
 *
 * The reference (Tealdragon204/nightcore-to-flac-analyzer) is pure Python and has no FFI; its
 * boundary for this path is the Python module surface.  These entry points are
 * what the drop-in `nightcore_analyzer` Python package binds through ctypes, one per kernel
 * family; each cites the reference interface it replaces (paths under
 * /root/reference/nightcore_analyzer/).
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; h_* is host memory;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises or allocates;
 *   - workspaces are caller-provided; query the size with the paired *_workspace_bytes();
 *   - return 0 on success, a negative NCFA_E_* code otherwise; ncfa_last_error() gives the
 *     (thread-local) message.  No exceptions cross the boundary.
 *   - a "segment" is a contiguous run of samples inside one device audio buffer:
 *     d_seg_off[i] (int64 sample offset) and d_seg_len[i] (int32 samples).  Analysis windows
 *     (io.py:94-110) and whole tracks (tempo.py:158) are both segments.
 */
#ifndef NCFA_H
#define NCFA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCFA_OK 0
#define NCFA_E_INVALID (-1)   /* bad argument */
#define NCFA_E_CUDA (-2)      /* CUDA runtime error (no device, launch failure, ...) */
#define NCFA_E_WORKSPACE (-3) /* workspace too small */
#define NCFA_E_OVERFLOW (-4)  /* internal capacity exceeded (reported, never silent) */

#define NCFA_N_FFT 2048
#define NCFA_N_MELS 128

int ncfa_version(void);
const char *ncfa_last_error(void);

/* Optional per-kernel timing for bench.py: while enabled, every kernel launch is bracketed by CUDA
 * events on its stream; ncfa_profile_report() waits for them and writes "name,launches,total_ms"
 * lines into buf, then forgets them. */
void ncfa_profile_enable(int on);
int ncfa_profile_report(char *buf, size_t cap);

/* Build and upload the constant tables (Hann, FFT twiddles, Slaney mel bank for `sr`) on the
 * current device.  Idempotent and thread-safe; the other entry points call it lazily. */
int ncfa_init_tables(int sr);

/* ---- io.py:38-40,94-110  _rms_db / slice_windows --------------------------------------------
 * d_meansq[i] = mean(x^2) of segment i accumulated in float64 (host finishes 20·log10(sqrt)). */
int ncfa_window_energy(const float *d_audio, const int64_t *d_seg_off, const int32_t *d_seg_len, int n_seg,
                       double *d_meansq, void *stream);

/* ---- io.py:58-79 strip_silence → librosa.effects.trim; xcorr.py:210 librosa.feature.rms -------
 * d_rms[f] = sqrt(mean(frame^2)), centred frames of `frame_length`, zero padded, hop `hop`;
 * n_frames = 1 + n/hop.  float32 like librosa for float32 input. */
int ncfa_rms_frames(const float *d_audio, int64_t n, int frame_length, int hop, float *d_rms, void *stream);

/* ---- tempo.py:44,158  librosa.onset.onset_strength(y, sr, hop_length) -----------------------------
 * For each segment: STFT(2048, hop, periodic Hann, centred, zero pad) → |.|^2 → Slaney mel(128)
 * → 10·log10(max(1e-10,.)) → clamp to (max over the segment − 80 dB) → positive first difference
 * → mean over mels → left-pad 1 + 1024/hop zeros → n_frames = 1 + len/hop values.
 * d_onset_off[i] is the float offset of segment i's envelope inside d_onset.
 * Workspace: log-mel spectrogram of every segment + per-segment maxima. */
size_t ncfa_onset_workspace_bytes(int n_seg, int max_seg_len, int hop);
int ncfa_onset_strength_batched(const float *d_audio, const int64_t *d_seg_off, const int32_t *d_seg_len,
                                int n_seg, int max_seg_len, int hop, int sr, float *d_onset,
                                const int64_t *d_onset_off, void *d_workspace, size_t workspace_bytes,
                                void *stream);

/* ---- tempo.py:63-66  librosa.feature.tempo(onset_envelope, sr, hop_length, start_bpm) ----------
 * Mean over frames of the inf-normalised Hann-windowed autocorrelation tempogram
 * (win_length = floor(8·sr/hop)), log-normal prior around d_start_bpm[i] (std 1 octave), lags
 * with bpm >= 320 masked; d_lag[i] = argmax lag (bpm = 60·sr/(hop·lag)); 0 for an all-zero
 * envelope (beat_track returns (0.0, [])). */
size_t ncfa_tempo_workspace_bytes(int n_seg, int max_env_len, int win_length);
int ncfa_tempo_lag_batched(const float *d_onset, const int64_t *d_onset_off, const int32_t *d_env_len, int n_seg,
                           int max_env_len, int hop, int sr, const double *d_start_bpm, int32_t *d_lag,
                           void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- tempo.py:45-49,159-164  librosa.beat.beat_track(onset_envelope, sr, hop_length, start_bpm) --
 * Dynamic-programming beat tracker with frames-per-beat = d_lag[i] <= max_lag (tightness 100, trim=True;
 * max_lag = the tempogram win_length floor(8·sr/hop) bounds every lag ncfa_tempo_lag_batched returns).
 * d_beats holds max_beats int32 frames per segment (ascending); d_n_beats[i] the count
 * (NCFA_E_OVERFLOW is flagged in d_n_beats[i] = -1 if max_beats is too small). */
size_t ncfa_beat_workspace_bytes(int n_seg, int max_env_len, int max_lag);
int ncfa_beat_track_batched(const float *d_onset, const int64_t *d_onset_off, const int32_t *d_env_len, int n_seg,
                            int max_env_len, const int32_t *d_lag, int max_lag, int32_t *d_beats, int max_beats,
                            int32_t *d_n_beats, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- consensus.py:243-267 _bootstrap_ratio, :270-312 compute_ibi_ratio, pitch.py:143-150 ---------
 * Job j resamples a_j (drawn FIRST) then b_j with numpy's Generator(PCG64).choice semantics
 * (32-bit Lemire bounded draws, low half of each 64-bit output first, high half buffered
 * across calls), n_boot times, starting from the PCG64 state {state_hi,state_lo,inc_hi,inc_lo}.
 *   boot[i] = median(a*) / median(b*)      (b_len = 0: boot[i] = median(a*))
 *   d_out[3j+0] = median(a)/median(b) (or median(a)),  [3j+1], [3j+2] = np.percentile(boot, q_lo/q_hi)
 * consensus._bootstrap_ratio(nc, src): a = nc, b = src.  compute_ibi_ratio(nc, src): a = src, b = nc.
 * d_boot (optional, n_jobs·n_boot doubles) receives the bootstrap replicates;
 * d_idx (optional, only when n_jobs == 1) receives every bounded draw in order (int32). */
size_t ncfa_bootstrap_workspace_bytes(int n_jobs, int max_a, int max_b, int n_boot);
int ncfa_bootstrap_ratio_batched(const double *d_a, const int64_t *d_a_off, const int32_t *d_a_len,
                                 const double *d_b, const int64_t *d_b_off, const int32_t *d_b_len, int n_jobs,
                                 int max_a, int max_b, int n_boot, const uint64_t h_pcg_state[4], double q_lo,
                                 double q_hi, double *d_out, double *d_boot, int32_t *d_idx, void *d_workspace,
                                 size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NCFA_H */
