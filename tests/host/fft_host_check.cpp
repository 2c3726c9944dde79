// CPU emulation of the frontend's per-frame pipeline (logmel_frame.cuh) - 16 "lanes" run phase by phase -
// checked against a direct double-precision DFT.  Built and run by tests/test_host_logic.py with g++.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../speech-intent-recognizer_b200/csrc/logmel_frame.cuh"
#include "../../speech-intent-recognizer_b200/csrc/frontend_tables.h"

using namespace sir;

int main(int argc, char** argv) {
    const int n_mels = argc > 1 ? atoi(argv[1]) : 64;
    HostFrontendTables ht = build_frontend_tables(16000, n_mels);
    FrontendTables t{ht.window.data(), ht.tw512.data(), ht.tw1024.data(), ht.mel_start.data(),
                     ht.mel_count.data(), ht.mel_offset.data(), ht.mel_weight.data()};
    double worst_p = 0, worst_mel = 0;
    unsigned s = 12345u;
    for (int trial = 0; trial < 4; ++trial) {
        std::vector<float> frame(1024);
        for (int n = 0; n < 1024; ++n) {
            s = s * 1664525u + 1013904223u;
            float u = (float)((s >> 8) & 0xFFFF) / 65536.f - 0.5f;
            frame[n] = (trial == 3 ? 1e-4f : 0.3f) * u + (trial >= 1 ? 0.4f * std::sin(0.05f * n * (trial + 1)) : 0.f);
        }
        std::vector<float> scratch_store(kFrameScratch + 4, 0.f);      // 16-byte aligned view, like the kernel's scratch
        float* scratch_ptr = scratch_store.data();
        while (reinterpret_cast<uintptr_t>(scratch_ptr) & 15u) ++scratch_ptr;
        struct { float* p; float* data() { return p; } float& operator[](size_t i) { return p[i]; } } scratch{scratch_ptr};
        float* scr_re = scratch.data();
        float* scr_im = scratch.data() + 16 * kRowPad;
        if (trial & 1) {   // exercise both loaders: contiguous, and reflect with the frame inside the signal
            ContiguousFrame ld{reinterpret_cast<const F2*>(frame.data())};
            for (int l = 0; l < 16; ++l) frame_phase_a(l, ld, t.window, t.tw512, scr_re, scr_im);
        } else {
            ReflectFrame<float> ld{frame.data(), 0, 1024};
            for (int l = 0; l < 16; ++l) frame_phase_a(l, ld, t.window, t.tw512, scr_re, scr_im);
        }
        std::vector<PhaseBRegs> rb(16);
        for (int q = 0; q < 16; ++q) frame_phase_b_load(q, scr_re, scr_im, rb[q]);
        for (int q = 0; q < 16; ++q) frame_phase_b_store(q, rb[q], scratch.data());
        std::vector<PhaseCRegs> rc(16);
        for (int q = 0; q < 16; ++q) frame_phase_c_compute(q, scratch.data(), t.tw1024, rc[q]);
        for (int q = 0; q < 16; ++q) frame_phase_c_store(q, rc[q], scratch.data());
        // reference: direct DFT in double of the windowed frame
        std::vector<double> P(516, 0.0);
        double pmax = 0;
        for (int k = 0; k <= 512; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 1024; ++n) {
                double x = (double)frame[n] * (double)ht.window[n];
                double a = -2.0 * M_PI * (double)((n * k) % 1024) / 1024.0;
                re += x * std::cos(a);
                im += x * std::sin(a);
            }
            P[k] = re * re + im * im;
            pmax = std::fmax(pmax, P[k]);
        }
        for (int k = 0; k <= 512; ++k) worst_p = std::fmax(worst_p, std::fabs(0.25 * scratch[k] - P[k]) / pmax);
        int covered = 0;
        std::vector<int> seen(n_mels, 0);
        for (int q = 0; q < 16; ++q)
            for (int j = 0; 16 * j < n_mels; ++j) {
                int m = mel_of_lane(q, j);
                if (m >= n_mels) continue;
                seen[m]++;
                covered++;
                double want = 0;
                for (int i = 0; i < ht.mel_count[m]; ++i)
                    want += 4.0 * (double)ht.mel_weight[ht.mel_offset[m] + i] * P[ht.mel_start[m] + i];
                double got = mel_band_power(m, scratch.data(), t);
                // compare in dB, the domain the reference normalises in
                double e = std::fabs(10 * std::log10(std::fmax(got, 1e-10)) - 10 * std::log10(std::fmax(want, 1e-10)));
                worst_mel = std::fmax(worst_mel, e);
            }
        for (int m = 0; m < n_mels; ++m)
            if (seen[m] != 1) { printf("FAIL mel %d covered %d times\n", m, seen[m]); return 1; }
    }
    size_t taps = ht.mel_weight.size();
    printf("n_mels %d taps %zu worst_power_rel %.3e worst_mel_db %.3e\n", n_mels, taps, worst_p, worst_mel);
    return (worst_p < 2e-6 && worst_mel < 2e-2) ? 0 : 1;
}
