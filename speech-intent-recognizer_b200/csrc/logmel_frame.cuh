// Per-frame arithmetic of the fused log-mel frontend: one 1024-sample frame is transformed by 16 lanes
// (half a warp) as a 512-point complex FFT of the even/odd-packed real frame, followed by the real-FFT
// post-pass, the power spectrum and the sparse triangular mel projection.
//
// Replaces, for one frame, the reference's chain:
//   torch.stft(n_fft 1024, hann, onesided)  TA:functional/functional.py:123-134
//   spec.abs().pow(2)                       TA:functional/functional.py:141-144
//   MelScale matmul with fb[513, n_mels]    TA:transforms/_transforms.py:417
//   AmplitudeToDB 10*log10(clamp(x,1e-10))  TA:functional/functional.py:390-391
//
// Decomposition (n = l + 16 j, k = k2 + 32 k1, l,k1 in [0,16), j,k2 in [0,32)):
//   Z[k2 + 32 k1] = sum_l W16^(l k1) * ( W512^(l k2) * sum_j z[l + 16 j] W32^(j k2) )
//   phase A: lane l   : 32-point FFT over j in registers, twiddle W512^(l k2), transpose through smem
//   phase B: lane q   : two 16-point FFTs over l (k2 = q and q + 16), Z stored in natural order
//   phase C: lane q   : post-pass for k = q + 16 m (and its mirror 512 - k): X[k] = Xe[k] + W1024^k Xo[k]
//   phase D: lane q   : mel bands {q, 31-q, 32+q, 63-q, ...} as short dot products over contiguous bins
//
// All phases are __host__ __device__ and take the lane index explicitly so tests/host/fft_host_check.cpp can
// emulate the 16 lanes on the CPU.  Between phases the caller synchronises the 16 lanes (__syncwarp).
#pragma once

#include "fft_regs.cuh"

namespace sir {

constexpr int kNfft = 1024;          // the only frame size the CUDA path implements (all reference configs)
constexpr int kHop = 512;
constexpr int kBins = 513;
constexpr int kMaxMels = 128;
constexpr int kRowPad = 33;          // transposition row stride (floats): conflict-free for both access orders
constexpr int kFrameScratch = 1072;  // floats per frame: 2*16*33 = 1056 transposition / 1024 Z / 513 P, +16 so the
                                     // second half-warp's buffer is shifted by 16 banks

struct alignas(8) F2 {      // 64-bit shared-memory accesses on the device, plain struct on the host
    float x, y;
};
struct alignas(16) F4 {
    float x, y, z, w;
};

// Device-resident constant tables, built on the host in double precision (frontend.cu: build_tables).
struct FrontendTables {
    const float* window;      // [1024] periodic Hann
    const float* tw512;       // [32][16][2]  (cos, -sin) of 2*pi*l*k2/512, index k2*16 + l
    const float* tw1024;      // [257][2]     (cos, sin)  of 2*pi*k/1024
    const int* mel_start;     // [n_mels] first bin of the band's run, rounded down to a multiple of 4
    const int* mel_count;     // [n_mels] number of taps (a multiple of 4; zero weights pad the run)
    const int* mel_offset;    // [n_mels] offset of the first tap in mel_weight (a multiple of 4)
    const float* mel_weight;  // taps, already multiplied by 0.25 (the post-pass leaves 4*|X|^2); 16-byte aligned
};

// ---- phase A ------------------------------------------------------------------------------------------------
// `load(n)` returns the real samples (2n, 2n+1) of the frame as an F2; scr_re/scr_im: 16*33 floats each.
// tw512 is stored [k2][l] so that the 16 lanes read consecutive 8-byte words (conflict-free).
template <typename Loader>
SIR_HD void frame_phase_a(int l, const Loader& load, const float* __restrict__ window,
                          const float* __restrict__ tw512, float* __restrict__ scr_re, float* __restrict__ scr_im) {
    float re[32], im[32];
    const F2* __restrict__ w2 = reinterpret_cast<const F2*>(window);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = l + 16 * j;            // complex index; samples 2n, 2n+1
        const F2 v = load(n), w = w2[n];
        re[j] = v.x * w.x;
        im[j] = v.y * w.y;
    }
    fft_dif<32>(re, im);
    const F2* __restrict__ tw = reinterpret_cast<const F2*>(tw512) + l;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        const float yr = re[bitrev<32>(k2)], yi = im[bitrev<32>(k2)];
        const F2 t = tw[k2 * 16];            // (cos, -sin)(2 pi l k2 / 512)
        scr_re[l * kRowPad + k2] = yr * t.x - yi * t.y;
        scr_im[l * kRowPad + k2] = yr * t.y + yi * t.x;
    }
}

// Loader over 1024 contiguous, 8-byte aligned samples (interior frames read global memory directly).
struct ContiguousFrame {
    const F2* base;
    SIR_HD F2 operator()(int n) const { return base[n]; }
};

// Loader that resolves torch.stft's reflect padding: sample i of the utterance for i in [first, first+1024),
// mirrored at 0 and at L-1 (TA:functional/functional.py:123-134, pad_mode="reflect").
// Sample types the frontend ingests: fp32 in [-1, 1], or 16-bit PCM scaled by 1/32768 exactly like torchaudio.load's
// normalisation of a PCM16 file (scripts/precompute_features.py:47) - the conversion is exact in fp32.
SIR_HD float sample_to_float(float v) { return v; }
SIR_HD float sample_to_float(short v) { return (float)v * (1.0f / 32768.0f); }

template <typename T = float>
struct ReflectFrame {
    const T* row;
    int first, L;
    SIR_HD float at(int i) const {
        // branch-free: the index is clamped and the load is unconditional, so the 128 loads of a frame are all in
        // flight together (a guarded load per sample serialised them: one L2 round trip each)
        const int r = i < 0 ? -i : (i >= L ? 2 * (L - 1) - i : i);
        const int rc = r < 0 ? 0 : (r >= L ? L - 1 : r);
        const float v = sample_to_float(row[rc]);
        return r == rc ? v : 0.f;
    }
    SIR_HD F2 operator()(int n) const { return F2{at(first + 2 * n), at(first + 2 * n + 1)}; }
};

// ---- phase B ------------------------------------------------------------------------------------------------
// Split in load / compute+store so the caller can put a lane barrier between them (Z overwrites the
// transposition buffer).
struct PhaseBRegs {
    float re0[16], im0[16], re1[16], im1[16];
};

SIR_HD void frame_phase_b_load(int q, const float* __restrict__ scr_re, const float* __restrict__ scr_im,
                               PhaseBRegs& r) {
#pragma unroll
    for (int l = 0; l < 16; ++l) {
        r.re0[l] = scr_re[l * kRowPad + q];
        r.im0[l] = scr_im[l * kRowPad + q];
        r.re1[l] = scr_re[l * kRowPad + q + 16];
        r.im1[l] = scr_im[l * kRowPad + q + 16];
    }
}

SIR_HD void frame_phase_b_store(int q, PhaseBRegs& r, float* __restrict__ z /* [512][2] */) {
    fft_dif<16>(r.re0, r.im0);
    fft_dif<16>(r.re1, r.im1);
    F2* __restrict__ z2 = reinterpret_cast<F2*>(z);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const int k = q + 32 * k1;
        z2[k] = F2{r.re0[bitrev<16>(k1)], r.im0[bitrev<16>(k1)]};
        z2[k + 16] = F2{r.re1[bitrev<16>(k1)], r.im1[bitrev<16>(k1)]};
    }
}

// ---- phase C ------------------------------------------------------------------------------------------------
struct PhaseCRegs {
    float pk[16], pm[16], p256;
};

SIR_HD void post_pass_pair(float ar, float ai, float br, float bi, float c, float s, float& p_k, float& p_m) {
    // 2E = A + conj(B), 2D = A - conj(B), 2O = -i 2D, 2T = (c - i s) 2O ; returns 4|E+T|^2 and 4|E-T|^2
    const float er = ar + br, ei = ai - bi;
    const float dr = ar - br, di = ai + bi;
    const float tr = c * di - s * dr;
    const float ti = -(c * dr) - s * di;
    const float xr = er + tr, xi = ei + ti;
    const float yr = er - tr, yi = ei - ti;
    p_k = xr * xr + xi * xi;
    p_m = yr * yr + yi * yi;
}

SIR_HD void frame_phase_c_compute(int q, const float* __restrict__ z, const float* __restrict__ tw1024, PhaseCRegs& r) {
    const F2* __restrict__ z2 = reinterpret_cast<const F2*>(z);
    const F2* __restrict__ tw = reinterpret_cast<const F2*>(tw1024);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int k = q + 16 * m;
        const F2 a = z2[k], b = z2[(512 - k) & 511], w = tw[k];
        post_pass_pair(a.x, a.y, b.x, b.y, w.x, w.y, r.pk[m], r.pm[m]);
    }
    float unused;
    const F2 a = z2[256], w = tw[256];
    post_pass_pair(a.x, a.y, a.x, a.y, w.x, w.y, r.p256, unused);
}

SIR_HD void frame_phase_c_store(int q, const PhaseCRegs& r, float* __restrict__ p /* [513] */) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int k = q + 16 * m;
        p[k] = r.pk[m];
        p[512 - k] = r.pm[m];
    }
    if (q == 0) {
        p[256] = r.p256;
        p[513] = p[514] = p[515] = 0.f;      // the 4-aligned mel runs may touch these (with zero weights)
    }
}

// ---- phase D ------------------------------------------------------------------------------------------------
// Lane q owns mel bands 16 j + q (j even) and 16 j + 15 - q (j odd): narrow low bands pair with wide high
// bands, so the 16 lanes carry nearly equal tap counts (the HTK triangles span 3..41 bins).
SIR_HD int mel_of_lane(int q, int j) { return (j & 1) ? 16 * j + 15 - q : 16 * j + q; }

// `p` (the frame's power array) and t.mel_weight must be 16-byte aligned: both runs are read as 4-float vectors.
SIR_HD float mel_band_power(int m, const float* __restrict__ p, const FrontendTables& t) {
    const int n4 = t.mel_count[m] >> 2;
    const F4* __restrict__ w4 = reinterpret_cast<const F4*>(t.mel_weight + t.mel_offset[m]);
    const F4* __restrict__ p4 = reinterpret_cast<const F4*>(p + t.mel_start[m]);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = 0; i < n4; ++i) {
        const F4 w = w4[i], x = p4[i];
        a0 += w.x * x.x;
        a1 += w.y * x.y;
        a2 += w.z * x.z;
        a3 += w.w * x.w;
    }
    return (a0 + a1) + (a2 + a3);
}

}  // namespace sir
