// Host-side construction (double precision, rounded once to fp32) of the frontend's constant tables.
// Plain C++ so both frontend.cu and the CPU check in tests/host/ can include it.
//
// What they restate:
//   window    torch.hann_window(1024, periodic=True)                TA:transforms/_transforms.py:70,86
//   mel bank  melscale_fbanks(513, 0, sr/2, n_mels, sr, None, htk)  TA:functional/functional.py:518-590
// The filterbank is kept SPARSE: band m is non-zero on one contiguous run of FFT bins (3..41 bins at the
// reference settings; 1,000 non-zeros of 32,832), stored CSR-style, pre-multiplied by 0.25 because the
// post-pass leaves 4|X|^2 (logmel_frame.cuh).  Every run is widened with zero weights to start and end on a
// multiple of 4 bins (mel_start, mel_count, mel_offset are multiples of 4; bins up to 515 may be referenced, the
// power array is zero there), so the kernel reads weights and powers as aligned 16-byte vectors.
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

namespace sir {

struct HostFrontendTables {
    std::vector<float> window;      // [1024]
    std::vector<float> tw512;       // [32*16*2]  (cos, -sin)(2 pi l k2 / 512), index k2*16 + l
    std::vector<float> tw1024;      // [257*2]    (cos,  sin)(2 pi k / 1024)
    std::vector<int32_t> mel_start, mel_count, mel_offset;
    std::vector<float> mel_weight;
};

inline HostFrontendTables build_frontend_tables(int sample_rate, int n_mels) {
    const double pi = 3.14159265358979323846;
    HostFrontendTables t;
    t.window.resize(1024);
    for (int n = 0; n < 1024; ++n) t.window[n] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * n / 1024.0));
    t.tw512.resize(16 * 32 * 2);
    for (int l = 0; l < 16; ++l)
        for (int k2 = 0; k2 < 32; ++k2) {
            const double a = 2.0 * pi * (double)(l * k2) / 512.0;
            t.tw512[(k2 * 16 + l) * 2] = (float)std::cos(a);
            t.tw512[(k2 * 16 + l) * 2 + 1] = (float)(-std::sin(a));
        }
    t.tw1024.resize(257 * 2);
    for (int k = 0; k <= 256; ++k) {
        const double a = 2.0 * pi * (double)k / 1024.0;
        t.tw1024[2 * k] = (float)std::cos(a);
        t.tw1024[2 * k + 1] = (float)std::sin(a);
    }
    // HTK mel triangles, f_min 0, f_max sample_rate/2, norm None.
    const int n_freqs = 513;
    const double f_max = (double)(sample_rate / 2);
    const double m_min = 2595.0 * std::log10(1.0 + 0.0 / 700.0);
    const double m_max = 2595.0 * std::log10(1.0 + f_max / 700.0);
    std::vector<double> f_pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) {
        const double m = m_min + (m_max - m_min) * (double)i / (double)(n_mels + 1);
        f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    }
    t.mel_start.assign(n_mels, 0);
    t.mel_count.assign(n_mels, 0);
    t.mel_offset.assign(n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
        t.mel_offset[m] = (int32_t)t.mel_weight.size();
        int first = -1, last = -2;
        std::vector<float> w(n_freqs, 0.f);
        for (int k = 0; k < n_freqs; ++k) {
            const double f = f_max * (double)k / (double)(n_freqs - 1);
            const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
            const double v = std::fmax(0.0, std::fmin(down, up));
            w[k] = (float)v;
            if (w[k] > 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        if (first < 0) {            // empty band (possible for very large n_mels): one zero tap
            first = 0;
            last = 0;
        }
        const int first4 = first & ~3, end4 = (last + 1 + 3) & ~3;       // end4 <= 516
        t.mel_start[m] = first4;
        t.mel_count[m] = end4 - first4;
        for (int k = first4; k < end4; ++k) t.mel_weight.push_back(k < n_freqs ? 0.25f * w[k] : 0.f);
    }
    return t;
}

}  // namespace sir
