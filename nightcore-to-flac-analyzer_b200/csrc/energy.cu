// Energy reductions: per-window mean square in float64 (io.py:38-40 _rms_db, used by
// slice_windows io.py:94-110 and energy_gate io.py:115-126) and framed RMS
// (librosa.feature.rms / effects.trim: io.py:76, xcorr.py:210-211).
#include "ncfa_common.cuh"

namespace ncfa {

__global__ void __launch_bounds__(256) window_energy_kernel(const float *__restrict__ audio,
                                                            const int64_t *__restrict__ seg_off,
                                                            const int32_t *__restrict__ seg_len,
                                                            double *__restrict__ meansq) {
    __shared__ double sh[8];
    const int seg = blockIdx.x;
    const int n = seg_len[seg];
    const float *x = audio + seg_off[seg];
    double acc = 0.0;
    // 4 independent accumulators per thread keep the FP64 pipe busy behind the loads
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = threadIdx.x;
    for (; i + 768 < n; i += 1024) {
        double v0 = (double)__ldg(x + i), v1 = (double)__ldg(x + i + 256), v2 = (double)__ldg(x + i + 512),
               v3 = (double)__ldg(x + i + 768);
        a0 = fma(v0, v0, a0);
        a1 = fma(v1, v1, a1);
        a2 = fma(v2, v2, a2);
        a3 = fma(v3, v3, a3);
    }
    for (; i < n; i += 256) {
        double v = (double)__ldg(x + i);
        a0 = fma(v, v, a0);
    }
    acc = (a0 + a1) + (a2 + a3);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        meansq[seg] = n > 0 ? s / (double)n : 0.0;
    }
}

// one warp per frame; frame f covers samples [f·hop − L/2, f·hop + L/2), zeros outside [0, n)
__global__ void __launch_bounds__(256) rms_frames_kernel(const float *__restrict__ x, int64_t n, int frame_length,
                                                         int hop, int64_t n_frames, float *__restrict__ out) {
    const int64_t f = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    const int lane = threadIdx.x & 31;
    const int64_t s0 = f * hop - frame_length / 2;
    double acc = 0.0;
    for (int j = lane; j < frame_length; j += 32) {
        int64_t p = s0 + j;
        if (p >= 0 && p < n) {
            double v = (double)__ldg(x + p);
            acc = fma(v, v, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[f] = (float)sqrt(acc / (double)frame_length);
}

}  // namespace ncfa

using namespace ncfa;

extern "C" int ncfa_window_energy(const float *d_audio, const int64_t *d_seg_off, const int32_t *d_seg_len, int n_seg,
                                  double *d_meansq, void *stream) {
    NCFA_REQUIRE(n_seg >= 0, "n_seg");
    if (n_seg == 0) return NCFA_OK;
    NCFA_REQUIRE(d_audio && d_seg_off && d_seg_len && d_meansq, "null pointer");
    {
        ProfScope _p("window_energy_kernel", (cudaStream_t)stream);
        window_energy_kernel<<<n_seg, 256, 0, (cudaStream_t)stream>>>(d_audio, d_seg_off, d_seg_len, d_meansq);
    }
    NCFA_LAUNCH_OK("window_energy_kernel");
    return NCFA_OK;
}

extern "C" int ncfa_rms_frames(const float *d_audio, int64_t n, int frame_length, int hop, float *d_rms, void *stream) {
    NCFA_REQUIRE(d_audio && d_rms, "null pointer");
    NCFA_REQUIRE(n >= 0 && frame_length > 0 && hop > 0, "n/frame_length/hop");
    const int64_t n_frames = 1 + n / hop;
    NCFA_REQUIRE((n_frames + 7) / 8 < 2147483647LL, "too many frames");
    {
        ProfScope _p("rms_frames_kernel", (cudaStream_t)stream);
        rms_frames_kernel<<<(unsigned)((n_frames + 7) / 8), 256, 0, (cudaStream_t)stream>>>(d_audio, n, frame_length, hop,
                                                                                       n_frames, d_rms);
    }
    NCFA_LAUNCH_OK("rms_frames_kernel");
    return NCFA_OK;
}
