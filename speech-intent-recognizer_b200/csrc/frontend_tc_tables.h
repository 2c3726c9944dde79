// Constant tables of the tensor-core DFT frontend (frontend_tc.cu), built on the host in double precision.
// Plain C++ (no CUDA headers) so that tests/host/tc_dft_host_check.cpp can include it and emulate the kernel's
// arithmetic on the CPU against a direct DFT.
//
// The 1024-point DFT of a windowed real frame xw[n] is computed in two matrix stages, n = 32 n1 + n2, k = k1 + 32 k2:
//   stage 1   Y[k1][n2]  = sum_n1 xw[32 n1 + n2] W32^(n1 k1)            k1 = 0..16 only (real input: Y[32-k1] = conj Y[k1])
//   twiddle   Y'[k1][n2] = Y[k1][n2] W1024^(n2 k1)                      (CUDA cores, between the stages)
//   stage 2   X[k1 + 32 k2] = sum_n2 Y'[k1][n2] W32^(n2 k2)             k2 = 0..31
// Bins k = k1 + 32 k2 with k mod 32 <= 16 come out directly; the others are mirrors, |X[k]| = |X[1024 - k]|.
// Both stages run on tcgen05 as TF32 (hi, lo) split products with fp32 accumulation:
//   x.B ~= x_hi.B_hi + x_hi.B_lo + x_lo.B_hi          (the dropped lo.lo term is ~2^-22 relative)
// hi = the fp32 value with its 13 low mantissa bits cleared - exactly what the tensor core reads of an fp32 operand
// (measured: tools/tf32_probe.cu, the low bits are ignored) -, lo = x - hi (exact in fp32; the tensor core keeps its 11
// leading bits).  TF32 has the exponent range of fp32: no per-frame scaling, no subnormal corner (the fp16 split needed
// a power-of-two scale per frame, i.e. a maximum over the windowed samples).  The constant matrices are split with round
// to nearest on the host.  The B operands are stored K-major with the 128-byte swizzle of the UMMA descriptors (rows of
// 32 four-byte elements), as the exact shared-memory image the kernel copies in.
//
// Stage-1 outputs are the 32 REAL numbers that describe Y[0..16] (Im Y[0] = Im Y[16] = 0):
//   o = 0: Re Y[0];  o = 1: Re Y[16];  o = 2j: Re Y[j];  o = 2j + 1: Im Y[j]   (j = 1..15)
// Stage-2 operand rows are (frame, k1) with K index kappa = 2 n2 + c (c = 0 re, 1 im) and outputs nu = 2 k2 + c'.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace sir {
namespace fetc {

constexpr int kTileFrames = 14;          // frames per work item: two self-contained 128-row UMMA tiles of 7 frames x 17 stage-2 rows
constexpr int kPStride = 532;            // floats per frame of the power buffer: 16-byte aligned rows whose 16-byte chunk index
                                         // advances by 5 (mod 8) per frame - lanes = frames read LDS.128 without bank conflicts

// Byte offset of 4-byte element (row, k), k = 0..31, inside a K-major SWIZZLE_128B operand tile whose rows are 128 bytes:
// 8-row atoms of 1024 bytes, the 16-byte chunk index XOR-ed with the row index inside the atom.
inline uint32_t sw128_offset(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + (k & 3) * 4);
}

// TF32 pieces of a double: hi = nearest value with a 10-bit mantissa, lo = the same of the remainder
inline float tf32_round(double v) {
    float f = (float)v;
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0xFFFu + ((u >> 13) & 1u);                                          // round to nearest even on bit 13
    u &= 0xFFFFE000u;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float tf32_trunc(float f) {                                           // what the tensor core reads of an fp32 operand
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u &= 0xFFFFE000u;
    std::memcpy(&f, &u, 4);
    return f;
}

struct HostTcTables {
    std::vector<float> b1_img;           // stage-1 operand image: [hi | lo] x 32 rows (o) x 128 bytes (K = n1): 2048 floats
    std::vector<float> b2_img;           // stage-2 operand image: [hi K 0..31 | hi K 32..63 | lo K 0..31 | lo K 32..63] x 64 rows (nu)
                                         // x 128 bytes: 8192 floats
    std::vector<float> win_img;          // Hann window as [8][32 lanes][4]: lane's w[32 n1 + lane], n1 = 4c .. 4c + 3
    std::vector<float> twiddle;          // [32 n2][16 (k1 = 1..16)][2]: (cos, -sin)(2 pi n2 k1 / 1024)
    std::vector<float> b1, b2;           // the unsplit fp32 matrices [K][N] (for the host emulation)
};

inline double stage1_entry(int n1, int o) {
    const double pi = 3.14159265358979323846;
    if (o == 0) return 1.0;
    if (o == 1) return (n1 & 1) ? -1.0 : 1.0;
    const int j = o >> 1;
    const double a = 2.0 * pi * (double)((n1 * j) & 31) / 32.0;
    return (o & 1) ? -std::sin(a) : std::cos(a);
}
inline double stage2_entry(int kappa, int nu) {
    const double pi = 3.14159265358979323846;
    const int n2 = kappa >> 1, c = kappa & 1, k2 = nu >> 1, cp = nu & 1;
    const double a = 2.0 * pi * (double)((n2 * k2) & 31) / 32.0;
    if (c == 0 && cp == 0) return std::cos(a);
    if (c == 1 && cp == 0) return std::sin(a);
    if (c == 0 && cp == 1) return -std::sin(a);
    return std::cos(a);
}

inline HostTcTables build_tc_tables() {
    const double pi = 3.14159265358979323846;
    HostTcTables t;
    t.b1_img.assign(2 * 32 * 32, 0.f);
    t.b2_img.assign(4 * 64 * 32, 0.f);
    t.b1.assign(32 * 32, 0.f);
    t.b2.assign(64 * 64, 0.f);
    for (int n1 = 0; n1 < 32; ++n1)
        for (int o = 0; o < 32; ++o) {
            const float v = (float)stage1_entry(n1, o);
            t.b1[n1 * 32 + o] = v;
            const float hi = tf32_round(stage1_entry(n1, o));
            const float lo = tf32_round(stage1_entry(n1, o) - (double)hi);
            t.b1_img[sw128_offset(o, n1) / 4] = hi;
            t.b1_img[1024 + sw128_offset(o, n1) / 4] = lo;
        }
    for (int kappa = 0; kappa < 64; ++kappa)
        for (int nu = 0; nu < 64; ++nu) {
            const float v = (float)stage2_entry(kappa, nu);
            t.b2[kappa * 64 + nu] = v;
            const float hi = tf32_round(stage2_entry(kappa, nu));
            const float lo = tf32_round(stage2_entry(kappa, nu) - (double)hi);
            t.b2_img[(kappa >> 5) * 2048 + sw128_offset(nu, kappa & 31) / 4] = hi;
            t.b2_img[4096 + (kappa >> 5) * 2048 + sw128_offset(nu, kappa & 31) / 4] = lo;
        }
    t.twiddle.resize(32 * 16 * 2);
    for (int n2 = 0; n2 < 32; ++n2)
        for (int k1 = 1; k1 <= 16; ++k1) {
            const double a = 2.0 * pi * (double)(n2 * k1) / 1024.0;
            t.twiddle[(n2 * 16 + (k1 - 1)) * 2] = (float)std::cos(a);
            t.twiddle[(n2 * 16 + (k1 - 1)) * 2 + 1] = (float)(-std::sin(a));
        }
    return t;
}

// Where the power of stage-2 output (k1, k2) lands in the one-sided spectrum, or -1 if that output is redundant.
inline int power_bin(int k1, int k2) {
    if (k1 == 0) return k2 <= 16 ? 32 * k2 : -1;
    if (k1 == 16) return k2 < 16 ? 16 + 32 * k2 : -1;
    return k2 < 16 ? k1 + 32 * k2 : 1024 - (k1 + 32 * k2);
}

}  // namespace fetc
}  // namespace sir
