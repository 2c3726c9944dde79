// Launch parameters shared by the two frontend kernels (frontend.cu: CUDA-core FFT; frontend_tc.cu: tensor-core DFT).
#pragma once

#include <cstdint>

#include "logmel_frame.cuh"

namespace sir {

// Device tables of the tensor-core DFT kernel (frontend_tc_tables.h builds them on the host).
struct TcDeviceTables {
    const float* b1_img;         // 8 KB   stage-1 operand image (tf32 pieces, swizzled: hi tile then lo tile)
    const float* b2_img;         // 32 KB  stage-2 operand image (hi K 0..31, hi K 32..63, lo K 0..31, lo K 32..63)
    const float* win_img;        // 4 KB   Hann window as [8][32 lanes][4]: lane's w[32 n1 + lane], n1 = 4c .. 4c + 3
    const float* twiddle;        // [32][16][2] (cos, -sin)(2 pi n2 k1 / 1024), k1 = 1..16
    const float* mel_weight;     // the sparse filterbank taps WITHOUT the 0.25 of the CUDA-core post-pass, the runs of a band
                                 // quad (4q .. 4q+3) zero-padded to the same length; 1536 floats, zero behind the taps
    const int32_t* mel_start;    // [n_mels] first bin / padded tap count / offset into mel_weight (multiples of 4)
    const int32_t* mel_count;
    const int32_t* mel_offset;
};

// One (utterance, 8-frame group) work item's contribution to the utterance statistics, written to global memory by
// the CTA that processed the item and merged in group order by the CTA that finishes the utterance.
struct alignas(32) ItemPartial {
    double s1, s2;             // sums of (v - shift) and (v - shift)^2 over the item's values
    float shift, vmax;
    int n, pad;                // number of values
};

struct FrontendParams {
    const void* wave;          // fp32 or int16 samples (template parameter of the kernel)
    int64_t wave_stride;
    const int32_t* lengths;
    int n_samples;
    int max_samples;
    int n_mels;
    int mode;
    int out_frames;
    float* out;
    const int32_t* masks;
    int32_t* status;
    FrontendTables tables;
    int mel_weight_count;
    // SIR_OUT_MFCC: dB values are staged in db_stage [batch][n_mels][stage_frames]; after the per-utterance maximum is
    // known each frame is clamped at max - top_db and projected with dct [n_mels][n_mfcc] (ortho DCT-II)
    float* db_stage;
    int stage_frames;
    const float* dct;
    int n_mfcc;
    float top_db;
    // work items: utterance i / groups_max, 8-frame group i % groups_max; per-item partial statistics and the
    // per-utterance count of finished items (zero between launches)
    int batch;
    int groups_max;
    struct ItemPartial* partials;
    int* counters;
    unsigned long long* work_counter;   // ticket counter (never reset) and the first ticket of this launch
    unsigned long long work_base;
    TcDeviceTables tc;                  // tensor-core DFT kernel only (frontend_tc.cu)
};


}  // namespace sir
