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
#include <cstdlib>
#include <cstring>
#include <vector>

#include "model.cuh"

namespace sir {

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
__global__ void __launch_bounds__(256) attention_fc_kernel(const float* __restrict__ y,      // [B, T, 512]
                                                           const float* __restrict__ att_w,  // [512]
                                                           const float* __restrict__ att_b_ptr,
                                                           const float* __restrict__ fc_w,  // [C][512]
                                                           const float* __restrict__ fc_b, float* __restrict__ logits,
                                                           int T, int C) {
    extern __shared__ __align__(16) float sm[];
    float* ctx = sm;              // [512]
    float* score = sm + 512;      // [T] scores, then softmax weights
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float att_b = __ldg(att_b_ptr);
    const float4* __restrict__ yb4 = reinterpret_cast<const float4*>(y + (int64_t)b * T * 512);
    const float4* __restrict__ aw4 = reinterpret_cast<const float4*>(att_w);
    float4 aw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) aw[j] = __ldg(aw4 + lane + 32 * j);
    for (int t = warp; t < T; t += 8) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = yb4[t * 128 + lane + 32 * j];
            s = fmaf(v.x, aw[j].x, s);
            s = fmaf(v.y, aw[j].y, s);
            s = fmaf(v.z, aw[j].z, s);
            s = fmaf(v.w, aw[j].w, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) score[t] = s + att_b;
    }
    __syncthreads();
    if (warp == 0) {                                        // softmax over time (models/models.py:63)
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, score[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) sum += expf(score[t] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        for (int t = lane; t < T; t += 32) score[t] = expf(score[t] - mx) * inv;
    }
    __syncthreads();
    {                                                       // context = sum_t w_t y_t: 2 features per thread
        const float2* __restrict__ yb2 = reinterpret_cast<const float2*>(yb4);
        float2 c = make_float2(0.f, 0.f);
        constexpr int kU = 5;                               // loads of five time steps in flight; the sum keeps the order of t
        for (int t0 = 0; t0 < T; t0 += kU) {
            float2 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = t0 + u < T ? yb2[(t0 + u) * 256 + tid] : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (t0 + u < T) {
                    const float w = score[t0 + u];
                    c.x = fmaf(v[u].x, w, c.x);
                    c.y = fmaf(v[u].y, w, c.y);
                }
            }
        }
        reinterpret_cast<float2*>(ctx)[tid] = c;
    }
    __syncthreads();
    const float4* __restrict__ ctx4 = reinterpret_cast<const float4*>(ctx);
    for (int c = warp; c < C; c += 8) {
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(fc_w + (int64_t)c * 512);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = ctx4[lane + 32 * j], w = __ldg(w4 + lane + 32 * j);
            s = fmaf(v.x, w.x, s);
            s = fmaf(v.y, w.y, s);
            s = fmaf(v.z, w.z, s);
            s = fmaf(v.w, w.w, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) logits[(int64_t)b * C + c] = s + __ldg(fc_b + c);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Device-side weight repack: flat state_dict-order fp32 parameters -> the layouts the kernels consume.  One
// launch; runs once per load in eval mode (BatchNorm folded into the conv weights + a shift) and once per
// step in training mode (raw conv weights; BatchNorm is applied with batch statistics by its own kernels).
// ---------------------------------------------------------------------------------------------------------
struct RepackParams {
    const float* flat;
    FlatOffsets off;
    int fold;
    float eps;
    int C, gin, H8;
    float *w1, *sh[3], *bih[2], *bhh[2], *bhh_perm[2], *att_w, *att_b, *fc_w, *fc_b;
    __half *cw_hi[2], *cw_lo[2], *cwt_hi[2], *cwt_lo[2], *wih_hi[2], *wih_lo[2], *whh_hi[2], *whh_lo[2];
    int64_t job_end[16];
    int n_jobs;
};

__device__ __forceinline__ double bn_scale(const RepackParams& p, int l, int o) {
    return (double)p.flat[p.off.bn_g[l] + o] / sqrt((double)p.flat[p.off.bn_v[l] + o] + (double)p.eps);
}

__device__ __forceinline__ void store_split(__half* hi, __half* lo, int64_t i, float x) {
    __half h, l;
    tc::split_f16(x, h, l);
    hi[i] = h;
    lo[i] = l;
}

__global__ void __launch_bounds__(256) repack_weights_kernel(const RepackParams p) {
    const int64_t total = p.job_end[p.n_jobs - 1];
    const int cin[3] = {1, 32, 64}, cout[3] = {32, 64, 128};
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        int job = 0;
        while (g >= p.job_end[job]) ++job;
        const int64_t i = g - (job ? p.job_end[job - 1] : 0);
        const float* f = p.flat;
        switch (job) {
            case 0: {                                   // conv1 [32][9]
                const int o = (int)(i / 9);
                const float w = f[p.off.conv_w[0] + i];
                p.w1[i] = p.fold ? (float)((double)w * bn_scale(p, 0, o)) : w;
                break;
            }
            case 1: {                                   // BN shifts (eval): beta - mean * scale
                int l = 0, o = (int)i;
                while (o >= cout[l]) o -= cout[l++];
                p.sh[l][o] = (float)((double)f[p.off.bn_b[l] + o] - (double)f[p.off.bn_m[l] + o] * bn_scale(p, l, o));
                break;
            }
            case 2:
            case 3: {                                   // conv2 / conv3 [tap][C_out][C_in]
                const int l = job - 1, ci_n = cin[l], co_n = cout[l];
                const int ci = (int)(i % ci_n), co = (int)((i / ci_n) % co_n), tap = (int)(i / ((int64_t)ci_n * co_n));
                const float w = f[p.off.conv_w[l] + ((int64_t)co * ci_n + ci) * 9 + tap];
                store_split(p.cw_hi[l - 1], p.cw_lo[l - 1], i, p.fold ? (float)((double)w * bn_scale(p, l, co)) : w);
                break;
            }
            case 4:
            case 5: {                                   // data-gradient weights [tap'][C_in][C_out], tap' = 8 - tap
                const int l = job - 3, ci_n = cin[l], co_n = cout[l];
                const int co = (int)(i % co_n), ci = (int)((i / co_n) % ci_n), tap = (int)(i / ((int64_t)ci_n * co_n));
                const float w = f[p.off.conv_w[l] + ((int64_t)co * ci_n + ci) * 9 + (8 - tap)];
                store_split(p.cwt_hi[l - 1], p.cwt_lo[l - 1], i, w);
                break;
            }
            case 6:
            case 7: {                                   // W_ih, both directions stacked along N; layer 0's columns
                const int l = job - 6;                  // go from c * H8 + h (models.py:55-57) to h * 128 + c
                const int in_sz = l == 0 ? p.gin : 512;
                const int fp = (int)(i % in_sz);
                const int n = (int)((i / in_sz) % 768), d = (int)(i / ((int64_t)in_sz * 768));
                int fsrc = fp;
                if (l == 0) {
                    const int hh = fp / 128, c = fp % 128;
                    fsrc = c * p.H8 + hh;
                }
                store_split(p.wih_hi[l], p.wih_lo[l], i, f[p.off.wih[l][d] + (int64_t)n * in_sz + fsrc]);
                break;
            }
            case 8:
            case 9: {                                   // W_hh as per-(direction, cluster rank) tiles, row = gate*32 + unit
                const int l = job - 8;
                const int k = (int)(i % 256), row = (int)((i / 256) % 96), r = (int)((i / (256 * 96)) % 8),
                          d = (int)(i / (256 * 96 * 8));
                const int gate = row / 32, u = row % 32;
                store_split(p.whh_hi[l], p.whh_lo[l], i, f[p.off.whh[l][d] + (int64_t)(gate * 256 + r * 32 + u) * 256 + k]);
                break;
            }
            case 10: {                                  // biases [layer][dir][768] (+ b_hh in the tile order above)
                const int n = (int)(i % 768), d = (int)((i / 768) % 2), l = (int)(i / 1536);
                p.bih[l][d * 768 + n] = f[p.off.bih[l][d] + n];
                const float bh = f[p.off.bhh[l][d] + n];
                p.bhh[l][d * 768 + n] = bh;
                const int gate = n / 256, r = (n % 256) / 32, u = n % 32;
                p.bhh_perm[l][d * 768 + r * 96 + gate * 32 + u] = bh;
                break;
            }
            default: {                                  // attention + fc
                if (i < 512) p.att_w[i] = f[p.off.att_w + i];
                else if (i == 512) p.att_b[0] = f[p.off.att_b];
                else if (i < 513 + (int64_t)p.C * 512) p.fc_w[i - 513] = f[p.off.fc_w + i - 513];
                else p.fc_b[i - 513 - (int64_t)p.C * 512] = f[p.off.fc_b + i - 513 - (int64_t)p.C * 512];
                break;
            }
        }
    }
}

int model_repack(sir_model* m, const float* d_flat, bool fold_bn, float bn_eps, cudaStream_t st) {
    RepackParams p{};
    p.flat = d_flat;
    p.off = m->off;
    p.fold = fold_bn ? 1 : 0;
    p.eps = bn_eps;
    p.C = m->num_classes;
    p.gin = m->gru_in;
    p.H8 = m->n_mels / 8;
    p.w1 = m->w1;
    p.sh[0] = m->sh1;
    p.sh[1] = m->sh2;
    p.sh[2] = m->sh3;
    p.cw_hi[0] = m->w2_hi;
    p.cw_lo[0] = m->w2_lo;
    p.cw_hi[1] = m->w3_hi;
    p.cw_lo[1] = m->w3_lo;
    p.cwt_hi[0] = m->w2t_hi;
    p.cwt_lo[0] = m->w2t_lo;
    p.cwt_hi[1] = m->w3t_hi;
    p.cwt_lo[1] = m->w3t_lo;
    for (int l = 0; l < 2; ++l) {
        p.bih[l] = m->bih[l];
        p.bhh[l] = m->bhh[l];
        p.bhh_perm[l] = m->bhh_perm[l];
        p.wih_hi[l] = m->wih_hi[l];
        p.wih_lo[l] = m->wih_lo[l];
        p.whh_hi[l] = m->whh_hi[l];
        p.whh_lo[l] = m->whh_lo[l];
    }
    p.att_w = m->att_w;
    p.att_b = m->att_b;
    p.fc_w = m->fc_w;
    p.fc_b = m->fc_b;
    const int64_t counts[12] = {288, 224, 9 * 64 * 32, 9 * 128 * 64, fold_bn ? 0 : 9 * 64 * 32, fold_bn ? 0 : 9 * 128 * 64,
                                (int64_t)1536 * m->gru_in, 1536 * 512, 2 * 8 * 96 * 256, 2 * 8 * 96 * 256, 2 * 1536,
                                513 + (int64_t)m->num_classes * 513};
    int64_t acc = 0;
    for (int j = 0; j < 12; ++j) {
        acc += counts[j];
        p.job_end[j] = acc;
    }
    p.n_jobs = 12;
    repack_weights_kernel<<<148 * 8, 256, 0, st>>>(p);
    SIR_CHECK_LAUNCH("repack_weights_kernel");
    return SIR_OK;
}

int launch_attention_fc(const sir_model* m, const float* y, float* logits, int B, int T, cudaStream_t st) {
    const size_t smem = (size_t)(T + 512) * sizeof(float);
    ProfScope ps("attention_fc", st);
    attention_fc_kernel<<<(unsigned)B, 256, smem, st>>>(y, m->att_w, m->att_b, m->fc_w, m->fc_b, logits, T, m->num_classes);
    SIR_CHECK_LAUNCH("attention_fc_kernel");
    return SIR_OK;
}

}  // namespace sir

// ---- C ABI ------------------------------------------------------------------------------------------------
using namespace sir;

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
    m->off = make_offsets(num_classes, n_mels);
    if (cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || m->num_sms < 1) m->num_sms = 148;
    // carve the repacked-weight buffers (sizes depend on the architecture only)
    const int gin = m->gru_in, C = num_classes;
    size_t nf = 0, nh = 0;
    auto f32 = [&](size_t n) {
        const size_t o = nf;
        nf += (n + 3) & ~(size_t)3;
        return o;
    };
    auto f16 = [&](size_t n) {
        const size_t o = nh;
        nh += (n + 63) & ~(size_t)63;           // 128-byte aligned sections (TMA global addresses)
        return o;
    };
    const size_t o_w1 = f32(288), o_s1 = f32(32), o_s2 = f32(64), o_s3 = f32(128);
    size_t o_bih[2], o_bhh[2], o_bhp[2];
    for (int l = 0; l < 2; ++l) {
        o_bih[l] = f32(1536);
        o_bhh[l] = f32(1536);
        o_bhp[l] = f32(1536);
    }
    const size_t o_aw = f32(512), o_ab = f32(1), o_fw = f32((size_t)C * 512), o_fb = f32(C);
    const size_t h_w2h = f16(9 * 64 * 32), h_w2l = f16(9 * 64 * 32), h_w3h = f16(9 * 128 * 64), h_w3l = f16(9 * 128 * 64);
    const size_t h_w2th = f16(9 * 64 * 32), h_w2tl = f16(9 * 64 * 32), h_w3th = f16(9 * 128 * 64),
                 h_w3tl = f16(9 * 128 * 64);
    size_t h_ih[2][2], h_hh[2][2];
    for (int l = 0; l < 2; ++l) {
        const size_t n = (size_t)1536 * (l == 0 ? gin : 512);
        h_ih[l][0] = f16(n);
        h_ih[l][1] = f16(n);
        h_hh[l][0] = f16(2 * 8 * 96 * 256);
        h_hh[l][1] = f16(2 * 8 * 96 * 256);
    }
    int rc = m->packed.reserve(nf * sizeof(float));
    if (rc == SIR_OK) rc = m->halves.reserve(nh * sizeof(__half) + 256);
    if (rc != SIR_OK) {
        sir_model_destroy(m);
        return rc;
    }
    float* base = (float*)m->packed.ptr;
    __half* hb = (__half*)m->halves.ptr;
    m->w1 = base + o_w1;
    m->sh1 = base + o_s1;
    m->sh2 = base + o_s2;
    m->sh3 = base + o_s3;
    m->att_w = base + o_aw;
    m->att_b = base + o_ab;
    m->fc_w = base + o_fw;
    m->fc_b = base + o_fb;
    m->w2_hi = hb + h_w2h;
    m->w2_lo = hb + h_w2l;
    m->w3_hi = hb + h_w3h;
    m->w3_lo = hb + h_w3l;
    m->w2t_hi = hb + h_w2th;
    m->w2t_lo = hb + h_w2tl;
    m->w3t_hi = hb + h_w3th;
    m->w3t_lo = hb + h_w3tl;
    for (int l = 0; l < 2; ++l) {
        m->bih[l] = base + o_bih[l];
        m->bhh[l] = base + o_bhh[l];
        m->bhh_perm[l] = base + o_bhp[l];
        m->wih_hi[l] = hb + h_ih[l][0];
        m->wih_lo[l] = hb + h_ih[l][1];
        m->whh_hi[l] = hb + h_hh[l][0];
        m->whh_lo[l] = hb + h_hh[l][1];
    }
    *out = m;
    return SIR_OK;
}

extern "C" void sir_model_destroy(sir_model* m) {
    if (!m) return;
    m->flat.release();
    m->packed.release();
    m->halves.release();
    for (auto& w : m->work) w.release();
    for (auto& w : m->ticket_buf) w.release();
    m->train_ws.release();
    delete m;
}

extern "C" int64_t sir_model_weight_count(const sir_model* m) { return m ? m->off.total : 0; }

extern "C" int sir_model_load_weights(sir_model* m, const float* weights, int64_t count, float bn_eps, void* stream) {
    if (!m || !weights) return fail(SIR_ERR_INVALID, "sir_model_load_weights: NULL argument");
    if (count != m->off.total)
        return fail(SIR_ERR_INVALID, "sir_model_load_weights: expected %lld floats, got %lld", (long long)m->off.total,
                    (long long)count);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = m->flat.reserve((size_t)count * sizeof(float));
    if (rc != SIR_OK) return rc;
    // host or device source; ordered on `st` with the repack and with later forwards
    SIR_CUDA(cudaMemcpyAsync(m->flat.ptr, weights, (size_t)count * sizeof(float), cudaMemcpyDefault, st));
    if ((rc = model_repack(m, (const float*)m->flat.ptr, true, bn_eps, st))) return rc;
    m->loaded = true;
    return SIR_OK;
}

namespace sir {

constexpr int kModelChunk = 336;   // utterances per pass through the workspace: 2 x 7 GRU clusters of 48 = one wave

struct Workspace {
    __half *act1_hi, *act1_lo, *act2_hi, *act2_lo, *gin_hi, *gin_lo, *y0_hi, *y0_lo;
    float *gi, *y0, *y1;
    tc::TicketSource* tickets = nullptr;    // the stream's tile-ticket counter (dynamic tile order of the persistent kernels)
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

// conv stack for utterances [first, first + count) of a workspace carved for a larger batch: the host entry
// that overlaps the H2D copy of the next sub-batch with the conv stack of the previous one uses this directly.
int model_forward_convs(sir_model* m, const Workspace& ws, const float* feat, int first, int count, int W, cudaStream_t st) {
    const int H = m->n_mels;
    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, Tg = W4 / 2;
    const size_t o1 = (size_t)first * H2 * W2 * 32, o2 = (size_t)first * H4 * W4 * 64, o3 = (size_t)first * Tg * m->gru_in;
    int rc;
    static const bool conv1_cuda_cores = [] {                // SIR_CONV1_KERNEL=cuda: round 1's fp32 kernel (A/B measurements)
        const char* v = getenv("SIR_CONV1_KERNEL");
        return v && (v[0] == 'c' || v[0] == 'C');
    }();
    if (conv1_cuda_cores) {
        dim3 grid((unsigned)((H2 * W2 + 127) / 128), (unsigned)count);
        ProfScope ps("conv1_bn_relu_pool", st);
        conv1_bn_relu_pool_kernel<<<grid, 128, 0, st>>>(feat, m->w1, m->sh1, ws.act1_hi + o1, ws.act1_lo + o1, H, W);
        SIR_CHECK_LAUNCH("conv1_bn_relu_pool_kernel");
    } else if ((rc = tc::conv1_tc(feat, m->w1, m->sh1, ws.act1_hi + o1, ws.act1_lo + o1, count, H, W, m->num_sms, st))) {
        return rc;
    }
    if ((rc = tc::tc_conv3x3_persistent<32, 64>(ws.act1_hi + o1, ws.act1_lo + o1, m->w2_hi, m->w2_lo, m->sh2, ws.act2_hi + o2,
                                                ws.act2_lo + o2, count, H2, W2, m->num_sms, st, "conv2_bn_relu_pool", ws.tickets)))
        return rc;
    // conv3 writes [B][T/8][H/8][128]: the GRU input, time-major with channels-last features
    return tc::tc_conv3x3_stream<64, 128>(ws.act2_hi + o2, ws.act2_lo + o2, m->w3_hi, m->w3_lo, m->sh3, ws.gin_hi + o3,
                                          ws.gin_lo + o3, count, H4, W4, 1, m->num_sms, st, "conv3_bn_relu_pool", ws.tickets);
}

// 2-layer bidirectional GRU + attention pooling + fc over the B utterances whose GRU input is in the workspace.
int model_forward_head(sir_model* m, const Workspace& ws, int B, int W, float* logits, cudaStream_t st) {
    const int Tg = W / 8;
    int rc;
    const __half *x_hi = ws.gin_hi, *x_lo = ws.gin_lo;
    int in_sz = m->gru_in;
    float* ys[2] = {ws.y0, ws.y1};
    for (int l = 0; l < 2; ++l) {
        const int M = B * Tg;
        if ((rc = tc::tc_gemm_nt(x_hi, x_lo, m->wih_hi[l], m->wih_lo[l], m->bih[l], ws.gi, M, 1536, in_sz, st,
                                 l == 0 ? "gru_l0_input_gemm" : "gru_l1_input_gemm", ws.tickets)))
            return rc;
        {
            ProfScope ps(l == 0 ? "gru_l0_recurrence" : "gru_l1_recurrence", st);
            if ((rc = tc::gru_layer_tc(m->whh_hi[l], m->whh_lo[l], ws.gi, m->bhh[l], ys[l],
                                       l == 0 ? ws.y0_hi : nullptr, l == 0 ? ws.y0_lo : nullptr, B, Tg, st)))
                return rc;
        }
        x_hi = ws.y0_hi;
        x_lo = ws.y0_lo;
        in_sz = 512;
    }
    return launch_attention_fc(m, ws.y1, logits, B, Tg, st);
}

int model_forward_chunk(sir_model* m, const Workspace& ws, const float* feat, int B, int W, float* logits,
                        cudaStream_t st) {
    int rc = model_forward_convs(m, ws, feat, 0, B, W, st);
    if (rc != SIR_OK) return rc;
    return model_forward_head(m, ws, B, W, logits, st);
}

}  // namespace sir

static int model_prepare(sir_model* m, int batch, int n_frames, Workspace& ws, int& chunk, void* stream) {   // (re)carves the workspace
    if (!m->loaded) return fail(SIR_ERR_INVALID, "sir_model_forward: weights not loaded");
    if (n_frames < 8) return fail(SIR_ERR_INVALID, "sir_model_forward: n_frames must be >= 8 (got %d)", n_frames);
    chunk = batch < kModelChunk ? batch : kModelChunk;
    Workspace probe;
    const size_t need = carve(probe, nullptr, chunk, m->n_mels, n_frames, m->gru_in);
    int slot = 0;
    while (slot < m->work_used && m->work_stream[slot] != stream) ++slot;
    if (slot == m->work_used) {
        if (slot == sir_model::kMaxStreams)
            return fail(SIR_ERR_UNSUPPORTED, "sir_model_forward: one handle serves at most %d streams; create another handle",
                        sir_model::kMaxStreams);
        m->work_stream[slot] = stream;
        ++m->work_used;
    }
    int rc = m->work[slot].reserve(need);
    if (rc != SIR_OK) return rc;
    carve(ws, (uint8_t*)m->work[slot].ptr, chunk, m->n_mels, n_frames, m->gru_in);
    if (!m->ticket_buf[slot].ptr) {
        if ((rc = m->ticket_buf[slot].reserve(sizeof(unsigned long long))) != SIR_OK) return rc;
        SIR_CUDA(cudaMemsetAsync(m->ticket_buf[slot].ptr, 0, sizeof(unsigned long long), (cudaStream_t)stream));
        m->ticket_src[slot] = tc::TicketSource{(unsigned long long*)m->ticket_buf[slot].ptr, 0};
    }
    ws.tickets = &m->ticket_src[slot];
    return SIR_OK;
}

extern "C" int sir_model_forward(sir_model* m, const float* d_features, int batch, int n_frames, float* d_logits,
                                 void* stream) {
    if (!m || !d_features || !d_logits) return fail(SIR_ERR_INVALID, "sir_model_forward: NULL argument");
    if (batch < 0) return fail(SIR_ERR_INVALID, "sir_model_forward: negative batch");
    if (batch == 0) return SIR_OK;
    Workspace ws;
    int chunk = 0;
    int rc = model_prepare(m, batch, n_frames, ws, chunk, stream);
    if (rc != SIR_OK) return rc;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = batch - b0 < chunk ? batch - b0 : chunk;
        rc = model_forward_chunk(m, ws, d_features + (int64_t)b0 * m->n_mels * n_frames, nb, n_frames,
                                 d_logits + (int64_t)b0 * m->num_classes, (cudaStream_t)stream);
        if (rc != SIR_OK) return rc;
    }
    return SIR_OK;
}

extern "C" int sir_model_forward_convs(sir_model* m, const float* d_features, int batch_total, int first, int count,
                                       int n_frames, void* stream) {
    if (!m || !d_features) return fail(SIR_ERR_INVALID, "sir_model_forward_convs: NULL argument");
    if (batch_total < 1 || batch_total > kModelChunk)
        return fail(SIR_ERR_UNSUPPORTED, "sir_model_forward_convs: batch_total must be in [1, %d] (got %d)", kModelChunk,
                    batch_total);
    if (first < 0 || count < 0 || first + count > batch_total)
        return fail(SIR_ERR_INVALID, "sir_model_forward_convs: [%d, %d) is not inside the batch of %d", first, first + count,
                    batch_total);
    if (count == 0) return SIR_OK;
    Workspace ws;
    int chunk = 0;
    int rc = model_prepare(m, batch_total, n_frames, ws, chunk, stream);
    if (rc != SIR_OK) return rc;
    return model_forward_convs(m, ws, d_features, first, count, n_frames, (cudaStream_t)stream);
}

extern "C" int sir_model_forward_head(sir_model* m, int batch_total, int n_frames, float* d_logits, void* stream) {
    if (!m || !d_logits) return fail(SIR_ERR_INVALID, "sir_model_forward_head: NULL argument");
    if (batch_total < 1 || batch_total > kModelChunk)
        return fail(SIR_ERR_UNSUPPORTED, "sir_model_forward_head: batch_total must be in [1, %d] (got %d)", kModelChunk,
                    batch_total);
    Workspace ws;
    int chunk = 0;
    int rc = model_prepare(m, batch_total, n_frames, ws, chunk, stream);
    if (rc != SIR_OK) return rc;
    return model_forward_head(m, ws, batch_total, n_frames, d_logits, (cudaStream_t)stream);
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

// ---------------------------------------------------------------------------------------------------------
// Evaluation head: softmax, arg-max, confidence, top-k and (optionally) accuracy / confusion counts on the
// device - the step right after the logits in scripts/test_model.py:121-156 (predict, get_top_predictions) and
// scripts/evaluate.py:79-98 (arg-max, accuracy_score, confusion_matrix).  One warp per utterance.
//   pred      = argmax(logits)           (first maximum, torch.argmax)
//   conf      = softmax(logits)[pred]
//   top-k     = argsort(probs)[::-1][:k]; among exactly equal probabilities the HIGHER index comes first (what a
//               stable argsort reversed gives; numpy's default introsort leaves that order unspecified)
// ---------------------------------------------------------------------------------------------------------
namespace sir {

constexpr int kMaxTopK = 8;

__global__ void __launch_bounds__(128) predict_kernel(const float* __restrict__ logits, int B, int C, int k,
                                                      const int64_t* __restrict__ labels, int32_t* __restrict__ pred,
                                                      float* __restrict__ conf, int32_t* __restrict__ topk_idx,
                                                      float* __restrict__ topk_prob, unsigned long long* __restrict__ confusion,
                                                      unsigned long long* __restrict__ correct) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* row = logits + (int64_t)b * C;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float v = row[c];
        if (v > mx) {
            mx = v;
            arg = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > mx || (ov == mx && oa < arg)) {
            mx = ov;
            arg = oa;
        }
    }
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    if (lane == 0) {
        if (pred) pred[b] = arg;
        if (conf) conf[b] = expf(row[arg] - mx) * inv;
        if (labels) {
            const int y = (int)labels[b];
            if (correct && y == arg) atomicAdd(correct, 1ull);
            if (confusion && y >= 0 && y < C) atomicAdd(confusion + (int64_t)y * C + arg, 1ull);
        }
    }
    if (k > 0 && topk_idx) {
        int chosen[kMaxTopK];
        for (int j = 0; j < k; ++j) {
            float best = -1.f;
            int bi = -1;
            for (int c = lane; c < C; c += 32) {
                bool taken = false;
                for (int t = 0; t < j; ++t) taken |= chosen[t] == c;
                if (taken) continue;
                const float pv = expf(row[c] - mx) * inv;
                if (pv > best || (pv == best && c > bi)) {
                    best = pv;
                    bi = c;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi > bi)) {
                    best = ov;
                    bi = oi;
                }
            }
            chosen[j] = bi;
            if (lane == 0) {
                topk_idx[(int64_t)b * k + j] = bi;
                if (topk_prob) topk_prob[(int64_t)b * k + j] = best;
            }
        }
    }
}

}  // namespace sir

extern "C" int sir_predict(const float* d_logits, int batch, int num_classes, int k, const int64_t* d_labels, int32_t* d_pred,
                           float* d_conf, int32_t* d_topk_idx, float* d_topk_prob, int64_t* d_confusion, int64_t* d_correct,
                           void* stream) {
    if (!d_logits || batch < 0 || num_classes < 1) return fail(SIR_ERR_INVALID, "sir_predict: bad arguments");
    if (k < 0 || k > kMaxTopK || k > num_classes) return fail(SIR_ERR_INVALID, "sir_predict: k must be in [0, min(%d, C)]", kMaxTopK);
    if (batch == 0) return SIR_OK;
    predict_kernel<<<(batch + 3) / 4, 128, 0, (cudaStream_t)stream>>>(d_logits, batch, num_classes, k, d_labels, d_pred, d_conf,
                                                                     d_topk_idx, d_topk_prob, (unsigned long long*)d_confusion,
                                                                     (unsigned long long*)d_correct);
    SIR_CHECK_LAUNCH("predict_kernel");
    return SIR_OK;
}
