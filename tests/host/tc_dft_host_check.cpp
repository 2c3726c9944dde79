// CPU emulation of the tensor-core DFT frontend's arithmetic (frontend_tc.cu): TF32 (hi, lo) operand splits as the tensor
// core reads them (the 13 low mantissa bits of an fp32 operand are ignored), the two matrix stages with the operand IMAGES
// the kernel copies into shared memory (decoded through the same swizzle function), the twiddle step and the mirrored bin
// map - against a direct double-precision DFT.
// Built and run by tests/test_host_logic.py with g++.  Bar: per trial, the worst mel-band error (dB) stays within 8x of what an
// fp32 radix-2 FFT of the same frame makes (pure tones 100+ dB above the noise floor are hard for fp32 itself), or 1e-4 dB.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../speech-intent-recognizer_b200/csrc/frontend_tables.h"
#include "../../speech-intent-recognizer_b200/csrc/frontend_tc_tables.h"

using namespace sir;
using namespace sir::fetc;

static void split(float v, float& hi, float& lo) {       // what the tensor core reads of v, and of what that left behind
    hi = tf32_trunc(v);
    lo = tf32_trunc(v - hi);
}

int main() {
    const HostTcTables tc = build_tc_tables();
    const HostFrontendTables ft = build_frontend_tables(16000, 64);
    // decode the operand images back into (hi, lo) matrices through the swizzle map
    std::vector<float> b1h(32 * 32), b1l(32 * 32), b2h(64 * 64), b2l(64 * 64);
    for (int k = 0; k < 32; ++k)
        for (int o = 0; o < 32; ++o) {
            b1h[k * 32 + o] = tc.b1_img[sw128_offset(o, k) / 4];
            b1l[k * 32 + o] = tc.b1_img[1024 + sw128_offset(o, k) / 4];
        }
    for (int k = 0; k < 64; ++k)
        for (int n = 0; n < 64; ++n) {
            b2h[k * 64 + n] = tc.b2_img[(k >> 5) * 2048 + sw128_offset(n, k & 31) / 4];
            b2l[k * 64 + n] = tc.b2_img[4096 + (k >> 5) * 2048 + sw128_offset(n, k & 31) / 4];
        }
    // every bin of the one-sided spectrum is produced exactly once
    std::vector<int> hits(513, 0);
    for (int k1 = 0; k1 <= 16; ++k1)
        for (int k2 = 0; k2 < 32; ++k2) {
            const int k = power_bin(k1, k2);
            if (k >= 0) {
                if (k > 512) { printf("FAIL bin map out of range\n"); return 1; }
                ++hits[k];
            }
        }
    for (int k = 0; k <= 512; ++k)
        if (hits[k] != 1) { printf("FAIL bin %d produced %d times\n", k, hits[k]); return 1; }

    const double pi = 3.14159265358979323846;
    double worst_db = 0, worst_rel = 0;
    unsigned s = 777u;
    for (int trial = 0; trial < 6; ++trial) {
        std::vector<float> x(1024);
        double lp = 0;
        for (int n = 0; n < 1024; ++n) {
            s = s * 1664525u + 1013904223u;
            const float u = (float)((s >> 8) & 0xFFFF) / 65536.f - 0.5f;
            lp = 0.98 * lp + u;                                          // 1/f-ish
            float v = 0.f;
            switch (trial) {
                case 0: v = 0.3f * u; break;
                case 1: v = 0.4f * std::sin(0.05f * n) + 1e-3f * u; break;
                case 2: v = 1e-4f * (float)lp; break;                    // near-silence
                case 3: v = 0.05f * (float)lp; break;                    // speech-like spectrum, 50-60 dB of dynamic range
                case 4: v = n < 512 ? 0.5f * (float)lp * 0.1f : 1e-4f * u; break;   // loud half, silent half
                default: v = 0.9f * std::sin(2.0 * pi * 37.3 * n / 1024.0) + 2e-5f * u; break;
            }
            x[n] = v;
        }
        // reference: double DFT of the fp32-windowed frame
        std::vector<double> pref(513);
        for (int k = 0; k <= 512; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 1024; ++n) {
                const double v = (double)x[n] * (double)ft.window[n];
                re += v * std::cos(2 * pi * k * n / 1024.0);
                im -= v * std::sin(2 * pi * k * n / 1024.0);
            }
            pref[k] = re * re + im * im;
        }
        // kernel arithmetic
        std::vector<float> ah(1024), al(1024);
        for (int n = 0; n < 1024; ++n) split(x[n] * ft.window[n], ah[n], al[n]);
        std::vector<float> P(513, -1.f);
        std::vector<float> a2h(17 * 64), a2l(17 * 64);
        for (int n2 = 0; n2 < 32; ++n2) {
            float y[32];
            for (int o = 0; o < 32; ++o) {
                float d0 = 0.f;                                           // one accumulator, three passes
                for (int n1 = 0; n1 < 32; ++n1) d0 += ah[32 * n1 + n2] * b1h[n1 * 32 + o];
                for (int n1 = 0; n1 < 32; ++n1) d0 += ah[32 * n1 + n2] * b1l[n1 * 32 + o];
                for (int n1 = 0; n1 < 32; ++n1) d0 += al[32 * n1 + n2] * b1h[n1 * 32 + o];
                y[o] = d0;
            }
            for (int k1 = 0; k1 <= 16; ++k1) {
                float re, im;
                if (k1 == 0) { re = y[0]; im = 0.f; }
                else {
                    const float c = tc.twiddle[(n2 * 16 + k1 - 1) * 2], d = tc.twiddle[(n2 * 16 + k1 - 1) * 2 + 1];
                    if (k1 == 16) { re = y[1] * c; im = y[1] * d; }
                    else {
                        const float a = y[2 * k1], b = y[2 * k1 + 1];
                        re = a * c - b * d;
                        im = a * d + b * c;
                    }
                }
                split(re, a2h[k1 * 64 + 2 * n2], a2l[k1 * 64 + 2 * n2]);
                split(im, a2h[k1 * 64 + 2 * n2 + 1], a2l[k1 * 64 + 2 * n2 + 1]);
            }
        }
        for (int k1 = 0; k1 <= 16; ++k1)
            for (int k2 = 0; k2 < 32; ++k2) {
                float xr[2];
                for (int cp = 0; cp < 2; ++cp) {
                    const int nu = 2 * k2 + cp;
                    float d0 = 0.f;
                    for (int kap = 0; kap < 64; ++kap) d0 += a2h[k1 * 64 + kap] * b2h[kap * 64 + nu];
                    for (int kap = 0; kap < 64; ++kap) d0 += a2h[k1 * 64 + kap] * b2l[kap * 64 + nu];
                    for (int kap = 0; kap < 64; ++kap) d0 += a2l[k1 * 64 + kap] * b2h[kap * 64 + nu];
                    xr[cp] = d0;
                }
                const int k = power_bin(k1, k2);
                if (k >= 0) P[k] = xr[0] * xr[0] + xr[1] * xr[1];
            }
        // fp32 iterative radix-2 FFT of the same windowed frame (what an fp32 library FFT does)
        std::vector<float> fr(1024), fi(1024, 0.f);
        for (int n = 0; n < 1024; ++n) { int r = 0; for (int b = 0; b < 10; ++b) if (n >> b & 1) r |= 1 << (9 - b); fr[r] = x[n] * ft.window[n]; }
        for (int len = 2; len <= 1024; len <<= 1)
            for (int i = 0; i < 1024; i += len)
                for (int j = 0; j < len / 2; ++j) {
                    const float c = (float)std::cos(2 * pi * j / len), sn = (float)-std::sin(2 * pi * j / len);
                    const float ur = fr[i + j], ui = fi[i + j], vr = fr[i + j + len / 2] * c - fi[i + j + len / 2] * sn,
                                vi = fr[i + j + len / 2] * sn + fi[i + j + len / 2] * c;
                    fr[i + j] = ur + vr; fi[i + j] = ui + vi; fr[i + j + len / 2] = ur - vr; fi[i + j + len / 2] = ui - vi;
                }
        double w32 = 0;
        for (int mband = 0; mband < 64; ++mband) {
            double a = 0, r = 0;
            for (int i = 0; i < ft.mel_count[mband]; ++i) {
                const int k = ft.mel_start[mband] + i;
                if (k > 512) continue;
                const double w = 4.0 * (double)ft.mel_weight[ft.mel_offset[mband] + i];
                a += w * (double)(fr[k] * fr[k] + fi[k] * fi[k]);
                r += w * pref[k];
            }
            w32 = std::fmax(w32, std::fabs(10 * std::log10(std::fmax(a, 1e-10)) - 10 * std::log10(std::fmax(r, 1e-10))));
        }
        double wtc = 0;
        for (int mband = 0; mband < 64; ++mband) {
            double a = 0, r = 0;
            for (int i = 0; i < ft.mel_count[mband]; ++i) {
                const int k = ft.mel_start[mband] + i;
                if (k > 512) continue;
                const double w = 4.0 * (double)ft.mel_weight[ft.mel_offset[mband] + i];
                a += w * (double)P[k];
                r += w * pref[k];
            }
            wtc = std::fmax(wtc, std::fabs(10 * std::log10(std::fmax(a, 1e-10)) - 10 * std::log10(std::fmax(r, 1e-10))));
        }
        printf("trial %d: fp32 FFT worst band error %.3e dB, tc %.3e dB\n", trial, w32, wtc);
        if (!(wtc < std::fmax(1e-4, 8.0 * w32))) { printf("FAIL: the split-precision DFT is much less accurate than an fp32 FFT\n"); return 1; }
        double pmax = 0;
        for (int k = 0; k <= 512; ++k) pmax = std::fmax(pmax, pref[k]);
        for (int k = 0; k <= 512; ++k) worst_rel = std::fmax(worst_rel, std::fabs(P[k] - pref[k]) / (pref[k] + 1e-9 * pmax));
        for (int mband = 0; mband < 64; ++mband) {
            double a = 0, r = 0;
            for (int i = 0; i < ft.mel_count[mband]; ++i) {
                const int k = ft.mel_start[mband] + i;
                if (k > 512) continue;
                const double w = 4.0 * (double)ft.mel_weight[ft.mel_offset[mband] + i];
                a += w * (double)P[k];
                r += w * pref[k];
            }
            const double da = 10 * std::log10(std::fmax(a, 1e-10)), dr = 10 * std::log10(std::fmax(r, 1e-10));
            worst_db = std::fmax(worst_db, std::fabs(da - dr));
        }
    }
    printf("tc dft: worst bin error %.3e (relative, floor 1e-9 of the peak), worst mel band error %.3e dB\n", worst_rel, worst_db);
    printf("OK\n");
    return 0;
}
