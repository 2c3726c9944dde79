// CNNAudioGRU forward (eval mode) for sm_100a - fp32 CUDA-core kernels.
//
// Replaces models/models.py:41-68 of the reference:
//   conv{1,2,3} 3x3 s1 p1 (no bias) + BatchNorm2d(eval, folded) + ReLU + MaxPool2d(2)   :50-52
//   permute(0,3,1,2).view(b, w, c*h)  (feature index = c * H/8 + h)                      :55-57
//   2-layer bidirectional GRU(hidden 256), gate order r,z,n, h0 = 0                      :60
//   softmax-attention pooling over time + Linear(512 -> num_classes)                     :63-67
//
// Data layout in HBM: activations are channels-last (NHWC) so that a 3x3 tap is a contiguous CIN vector
// (the K dimension of the implicit GEMM); conv3's epilogue writes the GRU input [B, T/8, 128*H/8] directly,
// so the reference's permute+contiguous copy (K9) never happens.
//
// conv2, conv3 and the two GRU input projections (88 % of the FLOPs) run on the tcgen05 tensor cores
// (gemm_tc.cu) as 3-pass fp16 hi/lo contractions; every producer therefore writes its activations as an
// fp16 (hi, lo) pair - the same 4 bytes per value as fp32.  conv1 (K = 9), the GRU recurrence and the
// attention/fc head stay on the fp32 CUDA cores.
#include <cmath>
#include <cstring>
#include <vector>

#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {

namespace tc {
int tc_gemm_nt(const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo, const float* bias,
               float* C, int M, int N, int K, cudaStream_t st, const char* name);
template <int CIN, int COUT>
int tc_conv3x3(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
               __half* out_hi, __half* out_lo, int B, int H, int W, int out_whc, cudaStream_t st, const char* name);
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box);
int gru_layer_tc(const CUtensorMap& tm_w_hi, const CUtensorMap& tm_w_lo, const float* gi, const float* bhh, float* y,
                 __half* y_hi, __half* y_lo, int B, int T, cudaStream_t st);
}  // namespace tc

// ---------------------------------------------------------------------------------------------------------
// conv1 (C_in = 1): one thread per pooled pixel, all 32 output channels; output NHWC [B, H/2, W/2, 32] as
// an fp16 (hi, lo) pair - the A operand of conv2's implicit GEMM.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv1_bn_relu_pool_kernel(const float* __restrict__ feat,
                                                                 const float* __restrict__ w1,      // [32][9] folded
                                                                 const float* __restrict__ shift1,  // [32]
                                                                 __half* __restrict__ out_hi,
                                                                 __half* __restrict__ out_lo, int H, int W) {
    __shared__ float s_w[32 * 9];
    __shared__ float s_shift[32];
    for (int i = threadIdx.x; i < 288; i += 128) s_w[i] = w1[i];
    if (threadIdx.x < 32) s_shift[threadIdx.x] = shift1[threadIdx.x];
    __syncthreads();
    const int H2 = H / 2, W2 = W / 2;
    const int pix = blockIdx.x * 128 + threadIdx.x;
    if (pix >= H2 * W2) return;
    const int b = blockIdx.y;
    const int h2 = pix / W2, w2 = pix - h2 * W2;
    const float* __restrict__ img = feat + (int64_t)b * H * W;
    float patch[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gh = 2 * h2 - 1 + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gw = 2 * w2 - 1 + c;
            patch[r][c] = (gh >= 0 && gh < H && gw >= 0 && gw < W) ? __ldg(img + (int64_t)gh * W + gw) : 0.f;
        }
    }
    const int64_t opix = (((int64_t)b * H2 + h2) * W2 + w2) * 32;
    uint2* __restrict__ dst_hi = reinterpret_cast<uint2*>(out_hi + opix);
    uint2* __restrict__ dst_lo = reinterpret_cast<uint2*>(out_lo + opix);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        float res[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const float* w = s_w + (c4 * 4 + cc) * 9;
            float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float wv = w[kh * 3 + kw];
                    o00 = fmaf(wv, patch[kh][kw], o00);
                    o01 = fmaf(wv, patch[kh][kw + 1], o01);
                    o10 = fmaf(wv, patch[kh + 1][kw], o10);
                    o11 = fmaf(wv, patch[kh + 1][kw + 1], o11);
                }
            res[cc] = fmaxf(fmaxf(fmaxf(o00, o01), fmaxf(o10, o11)) + s_shift[c4 * 4 + cc], 0.f);
        }
        __half h[4], l[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tc::split_f16(res[cc], h[cc], l[cc]);
        dst_hi[c4] = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                                (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
        dst_lo[c4] = make_uint2((uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
                                (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
    }
}

// ---------------------------------------------------------------------------------------------------------
// attention-softmax pooling over time + classifier head; one CTA per utterance.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_fc_kernel(const float* __restrict__ y,      // [B, T, 512]
                                                           const float* __restrict__ att_w,  // [512]
                                                           float att_b, const float* __restrict__ fc_w,  // [C][512]
                                                           const float* __restrict__ fc_b, float* __restrict__ logits,
                                                           int T, int C) {
    extern __shared__ float sm[];
    float* score = sm;            // [T]
    float* ctx = sm + T;          // [512]
    __shared__ float s_red[2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ yb = y + (int64_t)b * T * 512;
    for (int t = warp; t < T; t += 4) {
        float s = 0.f;
        for (int k = lane; k < 512; k += 32) s = fmaf(yb[(int64_t)t * 512 + k], __ldg(att_w + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) score[t] = s + att_b;
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, score[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) sum += expf(score[t] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            s_red[0] = mx;
            s_red[1] = 1.f / sum;
        }
    }
    __syncthreads();
    const float mx = s_red[0], inv = s_red[1];
    for (int k = tid; k < 512; k += 128) {
        float c = 0.f;
        for (int t = 0; t < T; ++t) c = fmaf(yb[(int64_t)t * 512 + k], expf(score[t] - mx) * inv, c);
        ctx[k] = c;
    }
    __syncthreads();
    for (int c = warp; c < C; c += 4) {
        float s = 0.f;
        for (int k = lane; k < 512; k += 32) s = fmaf(ctx[k], __ldg(fc_w + (int64_t)c * 512 + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) logits[(int64_t)b * C + c] = s + __ldg(fc_b + c);
    }
}

}  // namespace sir

// ---- C ABI ------------------------------------------------------------------------------------------------
using namespace sir;

struct sir_model {
    int num_classes = 31, n_mels = 64, gru_in = 1024;
    bool loaded = false;
    float att_b = 0.f;
    DeviceBuffer weights;        // repacked fp32 parameters
    DeviceBuffer weights_h;      // fp16 (hi, lo) operands of the tensor-core contractions
    DeviceBuffer work;           // activations
    // fp32 pointers into `weights`
    float *w1 = nullptr, *sh1 = nullptr, *sh2 = nullptr, *sh3 = nullptr;
    float *bih[2] = {nullptr, nullptr}, *bhh[2] = {nullptr, nullptr};
    float *att_w = nullptr, *fc_w = nullptr, *fc_b = nullptr;
    // fp16 pointers into `weights_h`: conv weights [tap][C_out][C_in] (BN scale folded), W_ih [1536][K]
    __half *w2_hi = nullptr, *w2_lo = nullptr, *w3_hi = nullptr, *w3_lo = nullptr;
    __half *wih_hi[2] = {nullptr, nullptr}, *wih_lo[2] = {nullptr, nullptr};
    // recurrent weights as per-CTA UMMA tiles [2 dirs][8 ranks][96][256] (gru_tc.cu) + their TMA maps
    __half *whh_hi[2] = {nullptr, nullptr}, *whh_lo[2] = {nullptr, nullptr};
    CUtensorMap tm_whh_hi[2], tm_whh_lo[2];
};

static int64_t model_weight_count(int num_classes, int n_mels) {
    const int64_t gin = 128 * (n_mels / 8);
    int64_t n = 32 * 9 + 4 * 32 + 64 * 32 * 9 + 4 * 64 + 128 * 64 * 9 + 4 * 128;
    n += 2 * (768 * gin + 768 * 256 + 768 + 768);
    n += 2 * (768 * 512 + 768 * 256 + 768 + 768);
    n += 512 + 1 + (int64_t)num_classes * 512 + num_classes;
    return n;
}

extern "C" int sir_model_create(sir_model** out, int num_classes, int n_mels) {
    if (!out) return fail(SIR_ERR_INVALID, "sir_model_create: out is NULL");
    *out = nullptr;
    if (num_classes < 1 || num_classes > 4096) return fail(SIR_ERR_INVALID, "bad num_classes %d", num_classes);
    if (n_mels < 8 || n_mels % 8 != 0 || n_mels > 128)
        return fail(SIR_ERR_UNSUPPORTED, "n_mels must be a multiple of 8 in [8,128] (got %d)", n_mels);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(SIR_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
    sir_model* m = new sir_model();
    m->num_classes = num_classes;
    m->n_mels = n_mels;
    m->gru_in = 128 * (n_mels / 8);
    *out = m;
    return SIR_OK;
}

extern "C" void sir_model_destroy(sir_model* m) {
    if (!m) return;
    m->weights.release();
    m->weights_h.release();
    m->work.release();
    delete m;
}

extern "C" int64_t sir_model_weight_count(const sir_model* m) {
    return m ? model_weight_count(m->num_classes, m->n_mels) : 0;
}

namespace {

struct HalfPack {                       // host staging of the fp16 (hi, lo) arrays, 128-byte aligned sections
    std::vector<__half> data;
    size_t add(const std::vector<float>& v, bool lo) {
        while (data.size() % 64) data.push_back(__float2half_rn(0.f));
        const size_t off = data.size();
        for (float x : v) {
            x = std::fmin(std::fmax(x, -65504.f), 65504.f);
            const __half h = __float2half_rn(x);
            data.push_back(lo ? __float2half_rn(x - __half2float(h)) : h);
        }
        return off;
    }
};

}  // namespace

extern "C" int sir_model_load_weights(sir_model* m, const float* weights, int64_t count, float bn_eps, void* stream) {
    if (!m || !weights) return fail(SIR_ERR_INVALID, "sir_model_load_weights: NULL argument");
    const int64_t want = model_weight_count(m->num_classes, m->n_mels);
    if (count != want)
        return fail(SIR_ERR_INVALID, "sir_model_load_weights: expected %lld floats, got %lld", (long long)want,
                    (long long)count);
    cudaStream_t st = (cudaStream_t)stream;
    SIR_CUDA(cudaStreamSynchronize(st));
    std::vector<float> h((size_t)count);
    SIR_CUDA(cudaMemcpy(h.data(), weights, (size_t)count * sizeof(float), cudaMemcpyDefault));
    const int gin = m->gru_in, C = m->num_classes, H8 = m->n_mels / 8;
    const float* p = h.data();               // walk the flat buffer in state_dict_spec order
    auto take = [&](int64_t n) {
        const float* r = p;
        p += n;
        return r;
    };
    const int cin[3] = {1, 32, 64}, cout[3] = {32, 64, 128};
    std::vector<float> packed;
    HalfPack hp;
    auto al4 = [&]() {
        while (packed.size() % 4) packed.push_back(0.f);
    };
    size_t off_w1 = 0, off_s[3], off_cw_hi[3] = {0, 0, 0}, off_cw_lo[3] = {0, 0, 0};
    for (int l = 0; l < 3; ++l) {
        const float* w = take((int64_t)cout[l] * cin[l] * 9);
        const float *g = take(cout[l]), *bt = take(cout[l]), *mu = take(cout[l]), *var = take(cout[l]);
        std::vector<double> scale(cout[l]);
        for (int o = 0; o < cout[l]; ++o) scale[o] = (double)g[o] / std::sqrt((double)var[o] + (double)bn_eps);
        if (l == 0) {                                 // [32][9] fp32
            al4();
            off_w1 = packed.size();
            for (int o = 0; o < 32; ++o)
                for (int k = 0; k < 9; ++k) packed.push_back((float)((double)w[o * 9 + k] * scale[o]));
        } else {                                      // [tap][cout][cin] -> fp16 hi/lo, K (= cin) contiguous
            std::vector<float> wt((size_t)9 * cout[l] * cin[l]);
            for (int k = 0; k < 9; ++k)
                for (int o = 0; o < cout[l]; ++o)
                    for (int i = 0; i < cin[l]; ++i)
                        wt[((size_t)k * cout[l] + o) * cin[l] + i] =
                            (float)((double)w[((int64_t)o * cin[l] + i) * 9 + k] * scale[o]);
            off_cw_hi[l] = hp.add(wt, false);
            off_cw_lo[l] = hp.add(wt, true);
        }
        al4();
        off_s[l] = packed.size();
        for (int o = 0; o < cout[l]; ++o) packed.push_back((float)((double)bt[o] - (double)mu[o] * scale[o]));
    }
    size_t off_bih[2], off_bhh[2], off_wih_hi[2], off_wih_lo[2], off_whh_hi[2], off_whh_lo[2];
    for (int l = 0; l < 2; ++l) {
        const int in_sz = l == 0 ? gin : 512;
        const float *wih[2], *whh[2], *bih[2], *bhh[2];
        for (int d = 0; d < 2; ++d) {
            wih[d] = take((int64_t)768 * in_sz);
            whh[d] = take((int64_t)768 * 256);
            bih[d] = take(768);
            bhh[d] = take(768);
        }
        // both directions stacked along N; layer 0's columns are re-ordered from the reference's
        // c * H8 + h (models/models.py:55-57) to h * 128 + c, the channels-last order conv3 writes
        std::vector<float> wcat((size_t)1536 * in_sz);
        for (int d = 0; d < 2; ++d)
            for (int n = 0; n < 768; ++n)
                for (int f = 0; f < in_sz; ++f) {
                    int fp = f;
                    if (l == 0) {
                        const int c = f / H8, hh = f % H8;
                        fp = hh * 128 + c;
                    }
                    wcat[((size_t)d * 768 + n) * in_sz + fp] = wih[d][(size_t)n * in_sz + f];
                }
        off_wih_hi[l] = hp.add(wcat, false);
        off_wih_lo[l] = hp.add(wcat, true);
        al4();
        off_bih[l] = packed.size();
        for (int d = 0; d < 2; ++d) packed.insert(packed.end(), bih[d], bih[d] + 768);
        // recurrent weights: per (direction, cluster rank) a 96-row tile, row = gate*32 + local unit
        std::vector<float> wt((size_t)2 * 8 * 96 * 256, 0.f);
        for (int d = 0; d < 2; ++d)
            for (int r = 0; r < 8; ++r)
                for (int g = 0; g < 3; ++g)
                    for (int u = 0; u < 32; ++u)
                        std::memcpy(&wt[(((size_t)d * 8 + r) * 96 + g * 32 + u) * 256],
                                    whh[d] + (size_t)(g * 256 + r * 32 + u) * 256, 256 * sizeof(float));
        off_whh_hi[l] = hp.add(wt, false);
        off_whh_lo[l] = hp.add(wt, true);
        off_bhh[l] = packed.size();
        for (int d = 0; d < 2; ++d) packed.insert(packed.end(), bhh[d], bhh[d] + 768);
    }
    const float* aw = take(512);
    const float* ab = take(1);
    const float* fw = take((int64_t)C * 512);
    const float* fb = take(C);
    al4();
    const size_t off_att = packed.size();
    packed.insert(packed.end(), aw, aw + 512);
    const size_t off_fcw = packed.size();
    packed.insert(packed.end(), fw, fw + (int64_t)C * 512);
    const size_t off_fcb = packed.size();
    packed.insert(packed.end(), fb, fb + C);
    m->att_b = ab[0];
    int rc = m->weights.reserve(packed.size() * sizeof(float));
    if (rc != SIR_OK) return rc;
    rc = m->weights_h.reserve(hp.data.size() * sizeof(__half));
    if (rc != SIR_OK) return rc;
    SIR_CUDA(cudaMemcpy(m->weights.ptr, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(m->weights_h.ptr, hp.data.data(), hp.data.size() * sizeof(__half), cudaMemcpyHostToDevice));
    float* base = (float*)m->weights.ptr;
    __half* hb = (__half*)m->weights_h.ptr;
    m->w1 = base + off_w1;
    m->sh1 = base + off_s[0];
    m->sh2 = base + off_s[1];
    m->sh3 = base + off_s[2];
    m->w2_hi = hb + off_cw_hi[1];
    m->w2_lo = hb + off_cw_lo[1];
    m->w3_hi = hb + off_cw_hi[2];
    m->w3_lo = hb + off_cw_lo[2];
    for (int l = 0; l < 2; ++l) {
        m->wih_hi[l] = hb + off_wih_hi[l];
        m->wih_lo[l] = hb + off_wih_lo[l];
        m->bih[l] = base + off_bih[l];
        m->bhh[l] = base + off_bhh[l];
        m->whh_hi[l] = hb + off_whh_hi[l];
        m->whh_lo[l] = hb + off_whh_lo[l];
        const uint64_t wd[2] = {256, 2 * 8 * 96};
        const uint32_t wb[2] = {64, 96};
        if ((rc = tc::make_tmap(&m->tm_whh_hi[l], m->whh_hi[l], 2, wd, wb))) return rc;
        if ((rc = tc::make_tmap(&m->tm_whh_lo[l], m->whh_lo[l], 2, wd, wb))) return rc;
    }
    m->att_w = base + off_att;
    m->fc_w = base + off_fcw;
    m->fc_b = base + off_fcb;
    m->loaded = true;
    return SIR_OK;
}

namespace sir {

constexpr int kModelChunk = 336;   // utterances per pass through the workspace: 2 x 7 GRU clusters of 48 = one wave

struct Workspace {
    __half *act1_hi, *act1_lo, *act2_hi, *act2_lo, *gin_hi, *gin_lo, *y0_hi, *y0_lo;
    float *gi, *y0, *y1;
};

static size_t carve(Workspace& w, uint8_t* base, int B, int H, int W, int gin) {
    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, Tg = W4 / 2;
    size_t off = 0;
    auto next = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    const size_t n1 = (size_t)B * H2 * W2 * 32, n2 = (size_t)B * H4 * W4 * 64, n3 = (size_t)B * Tg * gin,
                 ny = (size_t)B * Tg * 512;
    w.act1_hi = (__half*)next(n1 * 2);
    w.act1_lo = (__half*)next(n1 * 2);
    w.act2_hi = (__half*)next(n2 * 2);
    w.act2_lo = (__half*)next(n2 * 2);
    w.gin_hi = (__half*)next(n3 * 2);
    w.gin_lo = (__half*)next(n3 * 2);
    w.y0_hi = (__half*)next(ny * 2);
    w.y0_lo = (__half*)next(ny * 2);
    w.gi = (float*)next((size_t)B * Tg * 1536 * 4);
    w.y0 = (float*)next(ny * 4);
    w.y1 = (float*)next(ny * 4);
    return off;
}

int model_forward_chunk(sir_model* m, const Workspace& ws, const float* feat, int B, int W, float* logits,
                        cudaStream_t st) {
    const int H = m->n_mels;
    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, Tg = W4 / 2;
    int rc;
    {
        dim3 grid((unsigned)((H2 * W2 + 127) / 128), (unsigned)B);
        ProfScope ps("conv1_bn_relu_pool", st);
        conv1_bn_relu_pool_kernel<<<grid, 128, 0, st>>>(feat, m->w1, m->sh1, ws.act1_hi, ws.act1_lo, H, W);
        SIR_CHECK_LAUNCH("conv1_bn_relu_pool_kernel");
    }
    if ((rc = tc::tc_conv3x3<32, 64>(ws.act1_hi, ws.act1_lo, m->w2_hi, m->w2_lo, m->sh2, ws.act2_hi, ws.act2_lo, B, H2,
                                     W2, 0, st, "conv2_bn_relu_pool")))
        return rc;
    // conv3 writes [B][T/8][H/8][128]: the GRU input, time-major with channels-last features
    if ((rc = tc::tc_conv3x3<64, 128>(ws.act2_hi, ws.act2_lo, m->w3_hi, m->w3_lo, m->sh3, ws.gin_hi, ws.gin_lo, B, H4,
                                      W4, 1, st, "conv3_bn_relu_pool")))
        return rc;
    const __half *x_hi = ws.gin_hi, *x_lo = ws.gin_lo;
    int in_sz = m->gru_in;
    float* ys[2] = {ws.y0, ws.y1};
    for (int l = 0; l < 2; ++l) {
        const int M = B * Tg;
        if ((rc = tc::tc_gemm_nt(x_hi, x_lo, m->wih_hi[l], m->wih_lo[l], m->bih[l], ws.gi, M, 1536, in_sz, st,
                                 l == 0 ? "gru_l0_input_gemm" : "gru_l1_input_gemm")))
            return rc;
        {
            ProfScope ps(l == 0 ? "gru_l0_recurrence" : "gru_l1_recurrence", st);
            if ((rc = tc::gru_layer_tc(m->tm_whh_hi[l], m->tm_whh_lo[l], ws.gi, m->bhh[l], ys[l],
                                       l == 0 ? ws.y0_hi : nullptr, l == 0 ? ws.y0_lo : nullptr, B, Tg, st)))
                return rc;
        }
        x_hi = ws.y0_hi;
        x_lo = ws.y0_lo;
        in_sz = 512;
    }
    {
        const size_t smem = (size_t)(Tg + 512) * sizeof(float);
        ProfScope ps("attention_fc", st);
        attention_fc_kernel<<<(unsigned)B, 128, smem, st>>>(ws.y1, m->att_w, m->att_b, m->fc_w, m->fc_b, logits, Tg,
                                                             m->num_classes);
        SIR_CHECK_LAUNCH("attention_fc_kernel");
    }
    return SIR_OK;
}

}  // namespace sir

static int model_prepare(sir_model* m, int batch, int n_frames, Workspace& ws, int& chunk) {
    if (!m->loaded) return fail(SIR_ERR_INVALID, "sir_model_forward: weights not loaded");
    if (n_frames < 8) return fail(SIR_ERR_INVALID, "sir_model_forward: n_frames must be >= 8 (got %d)", n_frames);
    chunk = batch < kModelChunk ? batch : kModelChunk;
    Workspace probe;
    const size_t need = carve(probe, nullptr, chunk, m->n_mels, n_frames, m->gru_in);
    int rc = m->work.reserve(need);
    if (rc != SIR_OK) return rc;
    carve(ws, (uint8_t*)m->work.ptr, chunk, m->n_mels, n_frames, m->gru_in);
    return SIR_OK;
}

extern "C" int sir_model_forward(sir_model* m, const float* d_features, int batch, int n_frames, float* d_logits,
                                 void* stream) {
    if (!m || !d_features || !d_logits) return fail(SIR_ERR_INVALID, "sir_model_forward: NULL argument");
    if (batch < 0) return fail(SIR_ERR_INVALID, "sir_model_forward: negative batch");
    if (batch == 0) return SIR_OK;
    Workspace ws;
    int chunk = 0;
    int rc = model_prepare(m, batch, n_frames, ws, chunk);
    if (rc != SIR_OK) return rc;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = batch - b0 < chunk ? batch - b0 : chunk;
        rc = model_forward_chunk(m, ws, d_features + (int64_t)b0 * m->n_mels * n_frames, nb, n_frames,
                                 d_logits + (int64_t)b0 * m->num_classes, (cudaStream_t)stream);
        if (rc != SIR_OK) return rc;
    }
    return SIR_OK;
}

extern "C" int sir_pipeline_forward(sir_frontend* fe, sir_model* m, const float* d_wave, int64_t wave_stride,
                                    const int32_t* d_lengths, int n_samples, int batch, int max_samples,
                                    int out_frames, float* d_features, float* d_logits, void* stream) {
    if (!fe || !m || !d_wave || !d_logits) return fail(SIR_ERR_INVALID, "sir_pipeline_forward: NULL argument");
    if (batch <= 0) return batch == 0 ? SIR_OK : fail(SIR_ERR_INVALID, "sir_pipeline_forward: negative batch");
    if (!d_features) return fail(SIR_ERR_INVALID, "sir_pipeline_forward: d_features is required in this build");
    int rc = sir_frontend_forward(fe, d_wave, wave_stride, d_lengths, n_samples, batch, max_samples,
                                  SIR_OUT_LOGMEL_NORM, out_frames, d_features, nullptr, nullptr, stream);
    if (rc != SIR_OK) return rc;
    return sir_model_forward(m, d_features, batch, out_frames, d_logits, stream);
}
