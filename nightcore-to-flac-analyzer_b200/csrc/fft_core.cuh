// In-register 32-point complex FFT (radix-2 decimation in frequency, fully unrolled, compile-time
// twiddles).  The 1024-point transform of one STFT frame is two of these per lane with a
// twiddle + transpose between them (stft_core.cuh).
//
// On the device every complex value is one aligned register pair and the arithmetic is written with
// Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2, sm_100+): a complex add is ONE instruction
// and a complex multiply TWO — ptxas folds the half swaps, sign patterns and scalar broadcasts into
// operand modifiers (R.F32x2.LO_HI.NP, R.F32) — which halves the instruction count of a kernel that is
// issue bound (profiles/r1aq).  The same header compiles as plain host C++ (scalar code with the same
// products and roundings, FMA contraction aside) for tests/test_fft_core.py.
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

#if defined(__CUDA_ARCH__)
NCFA_HD float2 f2(cf a) { return make_float2(a.x, a.y); }
NCFA_HD cf c2(float2 a) { return cf{a.x, a.y}; }
NCFA_HD cf cadd(cf a, cf b) { return c2(__fadd2_rn(f2(a), f2(b))); }
NCFA_HD cf csub(cf a, cf b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
// a + conj(b), a − conj(b)
NCFA_HD cf cadd_conj(cf a, cf b) { return c2(__fadd2_rn(f2(a), make_float2(b.x, -b.y))); }
NCFA_HD cf csub_conj(cf a, cf b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, b.y))); }
// a · b:  (ax·bx, ay·bx) then + (ay·(−by), ax·by)
NCFA_HD cf cmul(cf a, cf b) {
    const float2 t = __fmul2_rn(f2(a), make_float2(b.x, b.x));
    return c2(__ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), t));
}
// a · (c − i·s):  (ax·c + ay·s, ay·c − ax·s)
NCFA_HD cf cmul_cs(cf a, float c, float s) {
    const float2 t = __fmul2_rn(f2(a), make_float2(c, c));
    return c2(__ffma2_rn(make_float2(a.y, a.x), make_float2(s, -s), t));
}
NCFA_HD cf cscale(cf a, float s) { return c2(__fmul2_rn(f2(a), make_float2(s, s))); }
NCFA_HD cf cmul_elem(cf a, cf b) { return c2(__fmul2_rn(f2(a), f2(b))); }  // (ax·bx, ay·by)
#else
NCFA_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
NCFA_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
NCFA_HD cf cadd_conj(cf a, cf b) { return cf{a.x + b.x, a.y - b.y}; }
NCFA_HD cf csub_conj(cf a, cf b) { return cf{a.x - b.x, a.y + b.y}; }
NCFA_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y}; }
NCFA_HD cf cmul_cs(cf a, float c, float s) { return cf{a.x * c + a.y * s, a.y * c - a.x * s}; }
NCFA_HD cf cscale(cf a, float s) { return cf{a.x * s, a.y * s}; }
NCFA_HD cf cmul_elem(cf a, cf b) { return cf{a.x * b.x, a.y * b.y}; }
#endif
NCFA_HD cf cconj(cf a) { return cf{a.x, -a.y}; }
// a · (−i)
NCFA_HD cf cmul_mi(cf a) { return cf{a.y, -a.x}; }

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
        return cmul_mi(d);
    } else if constexpr (T == 4) {  // · (1 - i)/√2 = (d + d·(−i))/√2
        return cscale(cadd(d, cmul_mi(d)), C[4]);
    } else if constexpr (T == 12) {  // · (-1 - i)/√2 = (d·(−i) − d)/√2
        return cscale(csub(cmul_mi(d), d), C[4]);
    } else {
        return cmul_cs(d, C[T], S[T]);
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
