// In-register radix-2 DIF FFTs of compile-time size (4..32 complex points).
//
// Every loop is fully unrolled and every twiddle is a compile-time literal, so a transform compiles to
// a straight line of FADD/FMUL/FFMA on registers (immediate-operand FFMA where a twiddle is involved);
// trivial twiddles (1, -i, (1-i)/sqrt2, (-1-i)/sqrt2) are special-cased at compile time.
//
// Output is left in BIT-REVERSED order: after fft_dif<N>(re, im) the value X[k] sits in element
// bitrev<N>(k).  Callers index with the constexpr bitrev, which costs nothing once unrolled.
//
// The functions are __host__ __device__ so that tests/host/fft_host_check.cpp can run the very same
// arithmetic (and the 16x32 four-step decomposition built on it) on the CPU build box.
#pragma once

#if defined(__CUDACC__)
#define SIR_HD __host__ __device__ __forceinline__
#else
#define SIR_HD inline
#endif

namespace sir {

// cos(2*pi*k/32), sin(2*pi*k/32) for k = 0..15 (quarter-wave symmetric, written out for clarity).
constexpr float kCos32[16] = {
    1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
    0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
    0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
    -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
constexpr float kSin32[16] = {
    0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
    0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
    1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
    0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};

template <int N>
SIR_HD constexpr int bitrev(int i) {
    int r = 0;
    for (int b = 1; b < N; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// (re, im) *= exp(-2*pi*i * J / M) with J, M compile-time; M divides 32, 0 <= J < M/2.
template <int J, int M>
SIR_HD void mul_twiddle(float& re, float& im) {
    constexpr int K = J * (32 / M);          // index into the 32-point table
    if constexpr (K == 0) {
        // multiply by 1
    } else if constexpr (K == 8) {           // -i : (a + ib)(-i) = b - ia
        const float t = re;
        re = im;
        im = -t;
    } else if constexpr (K == 4) {           // (1 - i)/sqrt2
        const float c = 0.70710678118654752440f;
        const float a = re, b = im;
        re = (a + b) * c;
        im = (b - a) * c;
    } else if constexpr (K == 12) {          // (-1 - i)/sqrt2
        const float c = 0.70710678118654752440f;
        const float a = re, b = im;
        re = (b - a) * c;
        im = -(a + b) * c;
    } else {                                 // w = c - i s
        constexpr float c = kCos32[K];
        constexpr float s = kSin32[K];
        const float a = re, b = im;
        re = a * c + b * s;
        im = b * c - a * s;
    }
}

template <int N, int HALF, int BASE, int J>
struct DifButterflies {
    SIR_HD static void run(float (&re)[N], float (&im)[N]) {
        if constexpr (J < HALF) {
            const float ar = re[BASE + J], ai = im[BASE + J];
            const float br = re[BASE + J + HALF], bi = im[BASE + J + HALF];
            re[BASE + J] = ar + br;
            im[BASE + J] = ai + bi;
            float dr = ar - br, di = ai - bi;
            mul_twiddle<J, 2 * HALF>(dr, di);
            re[BASE + J + HALF] = dr;
            im[BASE + J + HALF] = di;
            DifButterflies<N, HALF, BASE, J + 1>::run(re, im);
        }
    }
};

template <int N, int HALF, int BASE>
struct DifGroups {
    SIR_HD static void run(float (&re)[N], float (&im)[N]) {
        if constexpr (BASE < N) {
            DifButterflies<N, HALF, BASE, 0>::run(re, im);
            DifGroups<N, HALF, BASE + 2 * HALF>::run(re, im);
        }
    }
};

template <int N, int HALF>
struct DifStages {
    SIR_HD static void run(float (&re)[N], float (&im)[N]) {
        if constexpr (HALF >= 1) {
            DifGroups<N, HALF, 0>::run(re, im);
            DifStages<N, HALF / 2>::run(re, im);
        }
    }
};

// Forward DFT, X[k] = sum_n x[n] exp(-2*pi*i*n*k/N); result element bitrev<N>(k) holds X[k].
template <int N>
SIR_HD void fft_dif(float (&re)[N], float (&im)[N]) {
    static_assert(N == 4 || N == 8 || N == 16 || N == 32, "supported sizes");
    DifStages<N, N / 2>::run(re, im);
}

}  // namespace sir
