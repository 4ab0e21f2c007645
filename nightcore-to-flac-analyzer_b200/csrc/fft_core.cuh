// In-register 32-point complex FFT (radix-2 decimation in frequency, fully unrolled, compile-time
// twiddles).  The 1024-point transform of one STFT frame is two of these per lane with a
// twiddle + transpose between them (stft_onset.cu); the file is also compiled as plain host C++
// by tests/test_fft_core.py to check the butterfly network against numpy.
#pragma once

#if defined(__CUDACC__)
#define NCFA_HD __host__ __device__ __forceinline__
#else
#define NCFA_HD inline
#endif

namespace ncfa {

struct cf {
    float x, y;
};

NCFA_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
NCFA_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
NCFA_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
NCFA_HD cf cconj(cf a) { return cf{a.x, -a.y}; }

// bit reversal of a 5-bit index: after fft32_dif, X[k] lives in v[br5(k)]
NCFA_HD constexpr int br5(int k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// d · W_32^t for a compile-time t in [0,16);  W_32 = exp(-2πi/32)
template <int T>
NCFA_HD cf mul_w32(cf d) {
    constexpr float C[16] = {1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f,
                             0.0f,
                             -0.19509032201612826785f,
                             -0.38268343236508977173f,
                             -0.55557023301960222474f,
                             -0.70710678118654752440f,
                             -0.83146961230254523708f,
                             -0.92387953251128675613f,
                             -0.98078528040323044913f};
    constexpr float S[16] = {0.0f,
                             0.19509032201612826785f,
                             0.38268343236508977173f,
                             0.55557023301960222474f,
                             0.70710678118654752440f,
                             0.83146961230254523708f,
                             0.92387953251128675613f,
                             0.98078528040323044913f,
                             1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f};
    if constexpr (T == 0) {
        return d;
    } else if constexpr (T == 8) {  // · (-i)
        return cf{d.y, -d.x};
    } else if constexpr (T == 4) {  // · (1 - i)/√2
        return cf{(d.x + d.y) * C[4], (d.y - d.x) * C[4]};
    } else if constexpr (T == 12) {  // · (-1 - i)/√2
        return cf{(d.y - d.x) * C[4], -(d.x + d.y) * C[4]};
    } else {
        // (x + iy)(c - is) = (xc + ys) + i(yc - xs)
        return cf{d.x * C[T] + d.y * S[T], d.y * C[T] - d.x * S[T]};
    }
}

template <int N, int BASE, int J>
struct DifStage {
    static NCFA_HD void run(cf (&v)[32]) {
        cf a = v[BASE + J], b = v[BASE + J + N / 2];
        v[BASE + J] = cadd(a, b);
        v[BASE + J + N / 2] = mul_w32<J * (32 / N)>(csub(a, b));
        if constexpr (J + 1 < N / 2) DifStage<N, BASE, J + 1>::run(v);
    }
};

template <int N, int BASE>
NCFA_HD void fft_dif(cf (&v)[32]) {
    if constexpr (N >= 2) {
        DifStage<N, BASE, 0>::run(v);
        fft_dif<N / 2, BASE>(v);
        fft_dif<N / 2, BASE + N / 2>(v);
    }
}

// in place; output bin k is v[br5(k)]
NCFA_HD void fft32_dif(cf (&v)[32]) { fft_dif<32, 0>(v); }

}  // namespace ncfa
