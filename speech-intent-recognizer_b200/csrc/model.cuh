// Shared definition of the classifier handle (sir_model): flat-parameter offsets, device-side repacked weights,
// the evaluation workspace and the activations the training forward keeps for the backward pass.
//
// The flat parameter buffer has the order of utils/synth.py:state_dict_spec, i.e. the reference's state_dict
// (models/models.py:10-39): conv/bn 1..3 (weight, bn.weight, bn.bias, running_mean, running_var), gru l0,
// l0_reverse, l1, l1_reverse (weight_ih, weight_hh, bias_ih, bias_hh), attention, fc.
#pragma once

#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {

struct FlatOffsets {
    int64_t conv_w[3], bn_g[3], bn_b[3], bn_m[3], bn_v[3];
    int64_t wih[2][2], whh[2][2], bih[2][2], bhh[2][2];   // [layer][direction]
    int64_t att_w, att_b, fc_w, fc_b, total;
};

inline FlatOffsets make_offsets(int num_classes, int n_mels) {
    FlatOffsets o{};
    const int64_t gin = 128 * (n_mels / 8);
    const int cin[3] = {1, 32, 64}, cout[3] = {32, 64, 128};
    int64_t p = 0;
    for (int l = 0; l < 3; ++l) {
        o.conv_w[l] = p;
        p += (int64_t)cout[l] * cin[l] * 9;
        o.bn_g[l] = p;
        p += cout[l];
        o.bn_b[l] = p;
        p += cout[l];
        o.bn_m[l] = p;
        p += cout[l];
        o.bn_v[l] = p;
        p += cout[l];
    }
    for (int l = 0; l < 2; ++l) {
        const int64_t in_sz = l == 0 ? gin : 512;
        for (int d = 0; d < 2; ++d) {
            o.wih[l][d] = p;
            p += 768 * in_sz;
            o.whh[l][d] = p;
            p += 768 * 256;
            o.bih[l][d] = p;
            p += 768;
            o.bhh[l][d] = p;
            p += 768;
        }
    }
    o.att_w = p;
    p += 512;
    o.att_b = p;
    p += 1;
    o.fc_w = p;
    p += (int64_t)num_classes * 512;
    o.fc_b = p;
    p += num_classes;
    o.total = p;
    return o;
}

// Activations of one training forward, kept for sir_model_backward (all inside sir_model::train_ws).
struct TrainSaved {
    int B = 0, H = 0, W = 0;            // input features [B, H = n_mels, W = frames]
    const float* feat = nullptr;        // caller's features (must stay alive until the backward)
    float *z1, *z2, *z3;                // raw conv outputs, fp32 NHWC
    float *a1f, *a2f, *ginf;            // pooled activations fp32 (conv wgrad / W_ih grad operands)
    __half *a1_hi, *a1_lo, *a2_hi, *a2_lo, *gin_hi, *gin_lo;
    float* stats[3];                    // per layer [mean(C), invstd(C)]
    float *gi[2], *y[2];                // gate pre-activations of the input projection, layer outputs
    float* y0d;                         // layer-0 output after dropout (layer-1 input), fp32
    __half *y0d_hi, *y0d_lo, *ytmp_hi, *ytmp_lo;
    uint8_t* keep;                      // dropout keep mask [B*T*512]
    // backward scratch
    float *dy, *dx, *ds, *ctx, *dgi, *dgh, *gh, *hprevf, *dz, *dact, *wg_partial;
    __half *hprev_hi, *hprev_lo, *dz_hi, *dz_lo;
    // K-major fp16 (hi, lo) operands of the GRU gradient GEMMs (train.cu: operand_prep_kernel)
    __half *gT_hi, *gT_lo, *gs_hi, *gs_lo, *ghT_hi, *ghT_lo, *xT_hi, *xT_lo, *hT_hi, *hT_lo, *wT_hi, *wT_lo;
    double* bn_acc;                     // [3 layers][2 (fwd, bwd)][2*128] accumulators, zero between uses
    float* amax;                        // max |dz| per data-gradient convolution + the inverse split scales
};

}  // namespace sir

struct sir_model {
    int num_classes = 31, n_mels = 64, gru_in = 1024;
    bool loaded = false;       // eval weights (BN folded) are current
    sir::FlatOffsets off{};
    sir::DeviceBuffer flat;      // the model's own copy of the flat parameters (eval path)
    sir::DeviceBuffer packed;    // repacked fp32 parameters
    sir::DeviceBuffer halves;    // fp16 (hi, lo) operands of the tensor-core contractions
    // eval activations: one workspace per stream the handle is called on, so that batches enqueued on different
    // streams (IntentPipeline slots, bench.py's alternating steps) run concurrently on the GPU - the latency-bound
    // GRU recurrence of one batch then overlaps the frontend / conv stack of the next
    static constexpr int kMaxStreams = 32;
    sir::DeviceBuffer work[kMaxStreams];
    sir::DeviceBuffer ticket_buf[kMaxStreams];   // one tile-ticket counter per stream (its launches are stream-ordered)
    sir::tc::TicketSource ticket_src[kMaxStreams];
    void* work_stream[kMaxStreams] = {};
    int work_used = 0;
    sir::DeviceBuffer train_ws;  // training activations + backward scratch
    // fp32 pointers into `packed`
    float *w1 = nullptr, *sh1 = nullptr, *sh2 = nullptr, *sh3 = nullptr;
    float *bih[2] = {nullptr, nullptr}, *bhh[2] = {nullptr, nullptr}, *bhh_perm[2] = {nullptr, nullptr};
    float *att_w = nullptr, *att_b = nullptr, *fc_w = nullptr, *fc_b = nullptr;
    // fp16 pointers into `halves`: conv weights [tap][C_out][C_in], their flipped transposes [tap][C_in][C_out]
    // (data gradient), W_ih [1536][K], recurrent weights as per-CTA UMMA tiles [2 dirs][8 ranks][96][256]
    __half *w2_hi = nullptr, *w2_lo = nullptr, *w3_hi = nullptr, *w3_lo = nullptr;
    __half *w2t_hi = nullptr, *w2t_lo = nullptr, *w3t_hi = nullptr, *w3t_lo = nullptr;
    __half *wih_hi[2] = {nullptr, nullptr}, *wih_lo[2] = {nullptr, nullptr};
    __half *whh_hi[2] = {nullptr, nullptr}, *whh_lo[2] = {nullptr, nullptr};
    sir::TrainSaved ts;
    bool have_saved = false;
    const void* train_state = nullptr;       // device TrainState (train.cu): the dropout offset comes from it when set
    int num_sms = 148;
};

namespace sir {

namespace tc {
int tc_gemm_nt(const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo, const float* bias,
               float* C, int M, int N, int K, cudaStream_t st, const char* name, TicketSource* tickets = nullptr,
               const GemmOutput* out = nullptr);
template <int CIN, int COUT>
int tc_conv_wgrad(const __half* dz_hi, const __half* dz_lo, const __half* a_hi, const __half* a_lo, float* partial, int B, int H,
                  int W, int max_chunks, int* chunks_out, cudaStream_t st, const char* name);
template <int CIN, int COUT>
int tc_conv3x3(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
               __half* out_hi, __half* out_lo, float* raw_out, int B, int H, int W, int out_whc, cudaStream_t st,
               const char* name);
template <int CIN, int COUT>
int tc_conv3x3_persistent(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
                          __half* out_hi, __half* out_lo, int B, int H, int W, int num_sms, cudaStream_t st, const char* name,
                          TicketSource* tickets = nullptr);
template <int CIN, int COUT>
int tc_conv3x3_stream(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
                      __half* out_hi, __half* out_lo, int B, int H, int W, int out_whc, int num_sms, cudaStream_t st,
                      const char* name, TicketSource* tickets = nullptr);
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box);
// conv1 (C_in = 1) + folded BN + ReLU + 2x2 max-pool on tcgen05 (conv1_tc.cu): feat [B,H,W] fp32 -> [B,H/2,W/2,32] fp16 (hi, lo)
int conv1_tc(const float* feat, const float* w1, const float* shift1, __half* out_hi, __half* out_lo, int B, int H, int W,
             int num_sms, cudaStream_t st);
int gru_layer_tc(const __half* w_hi, const __half* w_lo, const float* gi, const float* bhh, float* y,
                 __half* y_hi, __half* y_lo, int B, int T, cudaStream_t st);
}  // namespace tc

// classifier.cu
int model_repack(sir_model* m, const float* d_flat, bool fold_bn, float bn_eps, cudaStream_t st);
int launch_attention_fc(const sir_model* m, const float* y, float* logits, int B, int T, cudaStream_t st);

}  // namespace sir
