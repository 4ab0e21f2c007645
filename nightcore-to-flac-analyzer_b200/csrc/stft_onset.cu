// Framing → Hann → real FFT-2048 → |.|^2 → Slaney mel(128) → dB  (kernel 1)
// and top_db clamp → positive flux → mean over mels → onset envelope (kernel 2).
//
// Replaces librosa.onset.onset_strength as called at tempo.py:44 (hop 512, one 10 s window per
// segment) and tempo.py:158 (hop 64, whole track per segment).  SURVEY.md Appendix A.1/A.2.
//
// Kernel 1 layout: one CTA of 4 warps owns a tile of FT consecutive frames of one segment.  The
// overlapping frames are staged once in shared memory; each warp then transforms whole frames:
// the 2048 real samples are packed as 1024 complex values, lane n2 holds z[32·n1+n2] in
// registers, runs a 32-point FFT over n1, multiplies by W_1024^(n2·k1), transposes through a
// padded shared-memory tile, runs the second 32-point FFT, and un-packs the real spectrum with
// one shuffle per complex value.  The power spectrum goes through shared memory into the
// sparse mel projection (each lane owns 4 bands).
#include "stft_core.cuh"

namespace ncfa {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kTileMax = 6144;  // samples staged per CTA

struct OnsetSmem {
    float tile[kTileMax];
    float hann[2048];
    float2 tw[1024];
    float2 scr[kWarps][32 * kScrStride];
    float melw[2048];
    int mel_start[NCFA_N_MELS + 1];
    int mel_bin0[NCFA_N_MELS];
    float wmax[kWarps];
};

__host__ __device__ inline int onset_frames_per_tile(int hop) {
    int ft = (kTileMax - 2048) / hop + 1;
    ft &= ~3;
    return ft < 4 ? 4 : (ft > 64 ? 64 : ft);
}

__global__ void __launch_bounds__(kThreads) stft_logmel_kernel(const float *__restrict__ audio,
                                                               const int64_t *__restrict__ seg_off,
                                                               const int32_t *__restrict__ seg_len, int hop,
                                                               int frames_per_tile, int frame_stride, Tables tb,
                                                               float *__restrict__ S, unsigned *__restrict__ seg_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OnsetSmem &sm = *reinterpret_cast<OnsetSmem *>(smem_raw);
    const int seg = blockIdx.y;
    const int len = seg_len[seg];
    const int n_frames = 1 + len / hop;
    const int f0 = blockIdx.x * frames_per_tile;
    if (f0 >= n_frames) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *src = audio + seg_off[seg];

    // ---- stage constants and the sample tile
    for (int i = tid; i < 2048; i += kThreads) sm.hann[i] = tb.hann[i];
    for (int i = tid; i < 1024; i += kThreads) sm.tw[i] = tb.tw1024[i];
    for (int i = tid; i < tb.mel_nnz; i += kThreads) sm.melw[i] = tb.mel_w[i];
    for (int i = tid; i <= NCFA_N_MELS; i += kThreads) sm.mel_start[i] = tb.mel_start[i];
    for (int i = tid; i < NCFA_N_MELS; i += kThreads) sm.mel_bin0[i] = tb.mel_bin0[i];
    const int nf_tile = min(frames_per_tile, n_frames - f0);
    const int tile_n = (nf_tile - 1) * hop + 2048;
    const int64_t pos0 = (int64_t)f0 * hop - 1024;  // sample index of tile[0]
    for (int i = tid; i < tile_n; i += kThreads) {
        int64_t p = pos0 + i;
        sm.tile[i] = (p >= 0 && p < len) ? __ldg(src + p) : 0.0f;
    }
    __syncthreads();

    const cf twl = cf{tb.tw2048[lane].x, tb.tw2048[lane].y};
    float2 *scr = sm.scr[warp];
    float *pw = reinterpret_cast<float *>(scr);
    float vmax = -INFINITY;

    for (int fl = warp; fl < nf_tile; fl += kWarps) {
        warp_power_spectrum(sm.tile + fl * hop, sm.hann, sm.tw, scr, twl, lane);

        // sparse mel projection: lane owns bands lane, 63-lane, 64+lane, 127-lane
        const int frame = f0 + fl;
        float *Sout = S + ((size_t)seg * frame_stride + frame) * NCFA_N_MELS;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int band = (q == 0) ? lane : (q == 1) ? 63 - lane : (q == 2) ? 64 + lane : 127 - lane;
            const int s = sm.mel_start[band], e = sm.mel_start[band + 1];
            const float *pk = pw + sm.mel_bin0[band];
            float acc = 0.0f;
            for (int i = s; i < e; ++i) acc = fmaf(sm.melw[i], pk[i - s], acc);
            float db = 10.0f * log10f(fmaxf(1e-10f, acc));
            Sout[band] = db;
            vmax = fmaxf(vmax, db);
        }
        __syncwarp();
    }
    vmax = warp_max(vmax);
    if (lane == 0) sm.wmax[warp] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = sm.wmax[0];
        for (int w = 1; w < kWarps; ++w) m = fmaxf(m, sm.wmax[w]);
        atomicMax(seg_max + seg, float_to_ordered(m));
    }
}

// onset[j] = 0 for j < pad;  else mean_m relu(clamp(S[j-pad+1][m]) - clamp(S[j-pad][m]))
__global__ void __launch_bounds__(256) flux_kernel(const float *__restrict__ S, const unsigned *__restrict__ seg_max,
                                                   const int32_t *__restrict__ seg_len, int hop, int frame_stride,
                                                   int pad, float *__restrict__ onset,
                                                   const int64_t *__restrict__ onset_off) {
    const int seg = blockIdx.y;
    const int n_frames = 1 + seg_len[seg] / hop;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= n_frames) return;
    float *out = onset + onset_off[seg];
    if (j < pad) {
        if (lane == 0) out[j] = 0.0f;
        return;
    }
    const float floor_db = ordered_to_float(seg_max[seg]) - 80.0f;
    const float4 *row0 = reinterpret_cast<const float4 *>(S + ((size_t)seg * frame_stride + (j - pad)) * NCFA_N_MELS);
    const float4 *row1 = row0 + NCFA_N_MELS / 4;
    float4 a = row0[lane], b = row1[lane];
    float s = fmaxf(0.0f, fmaxf(b.x, floor_db) - fmaxf(a.x, floor_db));
    s += fmaxf(0.0f, fmaxf(b.y, floor_db) - fmaxf(a.y, floor_db));
    s += fmaxf(0.0f, fmaxf(b.z, floor_db) - fmaxf(a.z, floor_db));
    s += fmaxf(0.0f, fmaxf(b.w, floor_db) - fmaxf(a.w, floor_db));
    s = warp_sum(s);
    if (lane == 0) out[j] = s * (1.0f / NCFA_N_MELS);
}

}  // namespace ncfa

using namespace ncfa;

extern "C" size_t ncfa_onset_workspace_bytes(int n_seg, int max_seg_len, int hop) {
    if (n_seg <= 0 || hop <= 0 || max_seg_len < 0) return 0;
    size_t frames = 1 + (size_t)max_seg_len / hop;
    return align_up((size_t)n_seg * frames * NCFA_N_MELS * sizeof(float), 256) + align_up((size_t)n_seg * 4, 256);
}

extern "C" int ncfa_onset_strength_batched(const float *d_audio, const int64_t *d_seg_off, const int32_t *d_seg_len,
                                           int n_seg, int max_seg_len, int hop, int sr, float *d_onset,
                                           const int64_t *d_onset_off, void *d_workspace, size_t workspace_bytes,
                                           void *stream) {
    NCFA_REQUIRE(n_seg >= 0 && n_seg <= 65535, "n_seg must be in [0, 65535] per call");
    if (n_seg == 0) return NCFA_OK;
    NCFA_REQUIRE(d_audio && d_seg_off && d_seg_len && d_onset && d_onset_off && d_workspace, "null pointer");
    NCFA_REQUIRE(hop >= 16 && hop <= 1024 && (hop % 2) == 0, "hop must be even and in [16, 1024]");
    NCFA_REQUIRE(max_seg_len >= 0, "max_seg_len");
    if (workspace_bytes < ncfa_onset_workspace_bytes(n_seg, max_seg_len, hop)) {
        set_error("onset workspace too small: %zu < %zu", workspace_bytes,
                  ncfa_onset_workspace_bytes(n_seg, max_seg_len, hop));
        return NCFA_E_WORKSPACE;
    }
    Tables tb;
    int rc = get_tables(sr, &tb);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int frames = 1 + max_seg_len / hop;
    float *S = (float *)d_workspace;
    unsigned *seg_max = (unsigned *)((char *)d_workspace + align_up((size_t)n_seg * frames * NCFA_N_MELS * 4, 256));
    NCFA_CUDA_OK(cudaMemsetAsync(seg_max, 0, (size_t)n_seg * 4, st));
    const int ft = onset_frames_per_tile(hop);
    static bool attr_done = false;
    if (!attr_done) {
        NCFA_CUDA_OK(cudaFuncSetAttribute(stft_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(OnsetSmem)));
        attr_done = true;
    }
    dim3 g1((frames + ft - 1) / ft, n_seg);
    {
        ProfScope _p("stft_logmel_kernel", st);
        stft_logmel_kernel<<<g1, kThreads, sizeof(OnsetSmem), st>>>(d_audio, d_seg_off, d_seg_len, hop, ft, frames, tb, S,
                                                                seg_max);
    }
    NCFA_LAUNCH_OK("stft_logmel_kernel");
    const int pad = 1 + NCFA_N_FFT / (2 * hop);
    dim3 g2((frames + 7) / 8, n_seg);
    {
        ProfScope _p("flux_kernel", st);
        flux_kernel<<<g2, 256, 0, st>>>(S, seg_max, d_seg_len, hop, frames, pad, d_onset, d_onset_off);
    }
    NCFA_LAUNCH_OK("flux_kernel");
    return NCFA_OK;
}
