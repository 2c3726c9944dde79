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
// These are the parity-reference kernels of the library (every op in fp32 FMA, fp32 accumulate); the
// tensor-core (tcgen05) contractions in gemm_tc.cu replace conv2/conv3/GRU-input GEMMs when enabled.
#include <cmath>
#include <cstring>
#include <vector>

#include "sir_common.cuh"

namespace sir {

// ---------------------------------------------------------------------------------------------------------
// conv1 (C_in = 1): one thread per pooled pixel, all 32 output channels; output NHWC [B, H/2, W/2, 32].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv1_bn_relu_pool_kernel(const float* __restrict__ feat,
                                                                 const float* __restrict__ w1,      // [32][9] folded
                                                                 const float* __restrict__ shift1,  // [32]
                                                                 float* __restrict__ out, int H, int W) {
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
    float4* __restrict__ dst = reinterpret_cast<float4*>(out + (((int64_t)b * H2 + h2) * W2 + w2) * 32);
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
        dst[c4] = make_float4(res[0], res[1], res[2], res[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// conv2 / conv3: direct 3x3 convolution as an implicit GEMM on CUDA cores.
// CTA tile: 8 x 16 conv outputs x all COUT; 256 threads = 16 pixel groups (2x4 pixels) x 16 channel groups.
// ---------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool GRU_LAYOUT>
__global__ void __launch_bounds__(256) conv3x3_bn_relu_pool_kernel(const float* __restrict__ in,     // [B,H,W,CIN]
                                                                   const float* __restrict__ wt,     // [9][CIN][COUT]
                                                                   const float* __restrict__ shift,  // [COUT]
                                                                   float* __restrict__ out, int H, int W) {
    constexpr int PSTRIDE = CIN + 4;                 // padded pixel stride: neighbouring pixel groups hit other banks
    constexpr int CPT = COUT / 16;                   // channels per thread (4 or 8)
    constexpr int NV = CPT / 4;
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                             // [10][18][PSTRIDE]
    float* wtap = smem + 10 * 18 * PSTRIDE;          // [CIN][COUT]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int h0 = blockIdx.y * 8, w0 = blockIdx.x * 16;
    const int H2 = H / 2, W2 = W / 2;

    for (int idx = tid; idx < 180 * (CIN / 4); idx += 256) {
        const int pi = idx / (CIN / 4), c4 = idx - pi * (CIN / 4);
        const int r = pi / 18, c = pi - r * 18;
        const int gh = h0 - 1 + r, gw = w0 - 1 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gh >= 0 && gh < H && gw >= 0 && gw < W)
            v = __ldg(reinterpret_cast<const float4*>(in + (((int64_t)b * H + gh) * W + gw) * CIN) + c4);
        *reinterpret_cast<float4*>(patch + pi * PSTRIDE + 4 * c4) = v;
    }

    const int pg = tid >> 4, cgp = tid & 15;
    const int pr = pg >> 2, pc = pg & 3;
    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
        __syncthreads();                             // previous tap's weights consumed (and patch visible)
        {
            const float4* __restrict__ src = reinterpret_cast<const float4*>(wt + (int64_t)tap * CIN * COUT);
            for (int idx = tid; idx < CIN * COUT / 4; idx += 256) reinterpret_cast<float4*>(wtap)[idx] = __ldg(src + idx);
        }
        __syncthreads();
        const int kh = tap / 3, kw = tap - kh * 3;
#pragma unroll 2
        for (int c4 = 0; c4 < CIN / 4; ++c4) {
            float4 a[8];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    a[r * 4 + c] = *reinterpret_cast<const float4*>(
                        patch + ((pr * 2 + r + kh) * 18 + (pc * 4 + c + kw)) * PSTRIDE + 4 * c4);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                float w[CPT];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 wv =
                        *reinterpret_cast<const float4*>(wtap + (4 * c4 + ci) * COUT + 64 * v + 4 * cgp);
                    w[4 * v] = wv.x;
                    w[4 * v + 1] = wv.y;
                    w[4 * v + 2] = wv.z;
                    w[4 * v + 3] = wv.w;
                }
#pragma unroll
                for (int px = 0; px < 8; ++px) {
                    const float av = ci == 0 ? a[px].x : (ci == 1 ? a[px].y : (ci == 2 ? a[px].z : a[px].w));
#pragma unroll
                    for (int j = 0; j < CPT; ++j) acc[px][j] = fmaf(av, w[j], acc[px][j]);
                }
            }
        }
    }

    // epilogue: + shift, ReLU, 2x2 max-pool (the thread's 2x4 pixels hold two complete windows)
    const int h2 = h0 / 2 + pr;
#pragma unroll
    for (int win = 0; win < 2; ++win) {
        const int w2 = w0 / 2 + pc * 2 + win;
        if (h2 >= H2 || w2 >= W2) continue;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float res[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int j = 4 * v + cc;
                const float m = fmaxf(fmaxf(acc[2 * win][j], acc[2 * win + 1][j]),
                                      fmaxf(acc[4 + 2 * win][j], acc[4 + 2 * win + 1][j]));
                res[cc] = fmaxf(m + __ldg(shift + 64 * v + 4 * cgp + cc), 0.f);
            }
            const int ch = 64 * v + 4 * cgp;
            if constexpr (GRU_LAYOUT) {              // out[b][w2][ch*H2 + h2]   (models/models.py:55-57)
                float* dst = out + ((int64_t)b * W2 + w2) * (COUT * H2) + h2;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) dst[(int64_t)(ch + cc) * H2] = res[cc];
            } else {                                 // NHWC
                *reinterpret_cast<float4*>(out + (((int64_t)b * H2 + h2) * W2 + w2) * COUT + ch) =
                    make_float4(res[0], res[1], res[2], res[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (GRU input projections, both directions stacked along N)
// 128x128 tile, BK = 8, 256 threads, 8x8 outputs per thread.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_nt_bias_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ C,
                                                           int M, int N, int K) {
    __shared__ __align__(16) float As[2][8][128 + 4];
    __shared__ __align__(16) float Bs[2][8][128 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    const int lrow = tid >> 1, lk = (tid & 1) * 4;       // each thread loads one float4 of A and one of W per k-tile
    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int arow = m0 + lrow;
    const float* __restrict__ ap = A + (int64_t)(arow < M ? arow : M - 1) * K + lk;
    const float* __restrict__ bp = Wt + (int64_t)(n0 + lrow) * K + lk;
    float4 av = __ldg(reinterpret_cast<const float4*>(ap));
    float4 bv = __ldg(reinterpret_cast<const float4*>(bp));
    const int nk = K / 8;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        As[buf][lk + 0][lrow] = av.x;
        As[buf][lk + 1][lrow] = av.y;
        As[buf][lk + 2][lrow] = av.z;
        As[buf][lk + 3][lrow] = av.w;
        Bs[buf][lk + 0][lrow] = bv.x;
        Bs[buf][lk + 1][lrow] = bv.y;
        Bs[buf][lk + 2][lrow] = bv.z;
        Bs[buf][lk + 3][lrow] = bv.w;
        __syncthreads();
        if (kt + 1 < nk) {
            av = __ldg(reinterpret_cast<const float4*>(ap + (kt + 1) * 8));
            bv = __ldg(reinterpret_cast<const float4*>(bp + (kt + 1) * 8));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        // the next iteration writes the other buffer; one barrier per k-tile is enough with two buffers
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
#pragma unroll
        for (int jv = 0; jv < 2; ++jv) {
            const int n = n0 + jv * 64 + tx * 4;
            const float4 bsv = __ldg(reinterpret_cast<const float4*>(bias + n));
            *reinterpret_cast<float4*>(C + (int64_t)m * N + n) =
                make_float4(acc[i][jv * 4 + 0] + bsv.x, acc[i][jv * 4 + 1] + bsv.y, acc[i][jv * 4 + 2] + bsv.z,
                            acc[i][jv * 4 + 3] + bsv.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// GRU recurrence of one layer, both directions, all T steps in ONE launch.
//
//   r = s(gi_r + W_hr h + b_hr)  z = s(gi_z + W_hz h + b_hz)  n = tanh(gi_n + r * (W_hn h + b_hn))
//   h' = (1 - z) n + z h                                             (torch.nn.GRU; gi already holds b_i*)
//
// A thread-block CLUSTER of 8 CTAs owns one (direction, slice of 32 utterances); CTA r of the cluster keeps
// the recurrent weights of hidden units [32 r, 32 r + 32) - 3 gates x 32 units x 256 k fp32 = 96 KB -
// resident in shared memory for all steps.  Per step every CTA loads the slice's previous hidden state
// (32 x 256 fp32, written by its 7 peers to the layer output y, still in L2) into shared memory, does its
// 96 x 256 by 256 x 32 product on the fp32 pipe, applies the gates and writes its 32 x 32 block of h' to y;
// one cluster barrier (release/acquire) per step orders the exchange.  No grid-wide synchronisation, no
// per-step launch.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGruCluster = 8;
constexpr int kGruUnits = 32;        // hidden units per CTA
constexpr int kGruBatch = 32;        // utterances per cluster
constexpr int kGruHStride = 256 + 4; // padded row of the staged hidden state (floats)
constexpr size_t kGruSmemBytes = (size_t)(3 * 64 * kGruUnits * 4 + kGruBatch * kGruHStride) * sizeof(float);

__global__ void __cluster_dims__(kGruCluster, 1, 1) __launch_bounds__(256, 1)
    gru_layer_kernel(const float* __restrict__ gi,   // [B*T, 1536]
                     const float* __restrict__ whh,  // [2][768][256]
                     const float* __restrict__ bhh,  // [2][768]
                     float* __restrict__ y,          // [B, T, 512]
                     int B, int T) {
    extern __shared__ __align__(16) float gsm[];
    float* Wsh = gsm;                                   // [3][64][32][4]
    float* Hs = gsm + 3 * 64 * kGruUnits * 4;           // [32][260]
    const int tid = threadIdx.x, unit = tid & 31, bg = tid >> 5;
    const int rank = blockIdx.x % kGruCluster;          // == cluster rank for a 1-D cluster
    const int slice = blockIdx.x / kGruCluster;
    const int dir = blockIdx.y;
    const int j0 = rank * kGruUnits, b0 = slice * kGruBatch;
    const float* __restrict__ W = whh + (int64_t)dir * 768 * 256;

    for (int idx = tid; idx < 3 * kGruUnits * 256; idx += 256) {
        const int k = idx & 255, u = (idx >> 8) & 31, g = idx >> 13;
        Wsh[((g * 64 + (k >> 2)) * kGruUnits + u) * 4 + (k & 3)] = __ldg(W + (int64_t)(g * 256 + j0 + u) * 256 + k);
    }
    const float br = __ldg(bhh + dir * 768 + j0 + unit), bz = __ldg(bhh + dir * 768 + 256 + j0 + unit),
                bn = __ldg(bhh + dir * 768 + 512 + j0 + unit);
    float hprev[4] = {0.f, 0.f, 0.f, 0.f};              // this thread's own (unit, 4 utterances) state
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        // gate pre-activations from the input projection: issue the loads before the matrix product
        float g_in[3][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int bb = b0 + bg * 4 + u;
            const float* __restrict__ gp = gi + ((int64_t)(bb < B ? bb : B - 1) * T + t) * 1536 + dir * 768 + j0 + unit;
            g_in[0][u] = __ldg(gp);
            g_in[1][u] = __ldg(gp + 256);
            g_in[2][u] = __ldg(gp + 512);
        }
        float acc[3][4];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[g][u] = 0.f;
        if (s > 0) {
            const int tp = dir == 0 ? t - 1 : t + 1;
            // stage h_{s-1} of the whole slice: 32 rows x 1 KB, written by the cluster's CTAs in the last step
            for (int idx = tid; idx < kGruBatch * 64; idx += 256) {
                const int row = idx >> 6, c4 = idx & 63;
                const int bb = b0 + row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bb < B) v = __ldcg(reinterpret_cast<const float4*>(y + ((int64_t)bb * T + tp) * 512 + dir * 256) + c4);
                *reinterpret_cast<float4*>(Hs + row * kGruHStride + 4 * c4) = v;
            }
            __syncthreads();
            const float4* __restrict__ w4 = reinterpret_cast<const float4*>(Wsh) + unit;
            const float* __restrict__ hrow = Hs + (bg * 4) * kGruHStride;
#pragma unroll 4
            for (int k4 = 0; k4 < 64; ++k4) {
                const float4 w0 = w4[(0 * 64 + k4) * kGruUnits], w1 = w4[(1 * 64 + k4) * kGruUnits],
                             w2 = w4[(2 * 64 + k4) * kGruUnits];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 h = *reinterpret_cast<const float4*>(hrow + u * kGruHStride + 4 * k4);
                    acc[0][u] = fmaf(w0.w, h.w, fmaf(w0.z, h.z, fmaf(w0.y, h.y, fmaf(w0.x, h.x, acc[0][u]))));
                    acc[1][u] = fmaf(w1.w, h.w, fmaf(w1.z, h.z, fmaf(w1.y, h.y, fmaf(w1.x, h.x, acc[1][u]))));
                    acc[2][u] = fmaf(w2.w, h.w, fmaf(w2.z, h.z, fmaf(w2.y, h.y, fmaf(w2.x, h.x, acc[2][u]))));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int bb = b0 + bg * 4 + u;
            const float r = 1.f / (1.f + expf(-(g_in[0][u] + acc[0][u] + br)));
            const float z = 1.f / (1.f + expf(-(g_in[1][u] + acc[1][u] + bz)));
            const float n = tanhf(g_in[2][u] + r * (acc[2][u] + bn));
            const float hn = (1.f - z) * n + z * hprev[u];
            hprev[u] = hn;
            if (bb < B) y[((int64_t)bb * T + t) * 512 + dir * 256 + j0 + unit] = hn;
        }
        // publish this step's h' to the 7 peers (and make sure nobody still reads Hs) before the next step
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
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
    DeviceBuffer weights;        // repacked parameters
    DeviceBuffer work;           // activations
    // device pointers into `weights`
    float *w1 = nullptr, *sh1 = nullptr, *w2 = nullptr, *sh2 = nullptr, *w3 = nullptr, *sh3 = nullptr;
    float *wih[2] = {nullptr, nullptr}, *bih[2] = {nullptr, nullptr}, *whh[2] = {nullptr, nullptr},
          *bhh[2] = {nullptr, nullptr};
    float *att_w = nullptr, *fc_w = nullptr, *fc_b = nullptr;
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
    m->work.release();
    delete m;
}

extern "C" int64_t sir_model_weight_count(const sir_model* m) {
    return m ? model_weight_count(m->num_classes, m->n_mels) : 0;
}

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
    const int gin = m->gru_in, C = m->num_classes;
    // walk the flat buffer in state_dict_spec order
    const float* p = h.data();
    auto take = [&](int64_t n) {
        const float* r = p;
        p += n;
        return r;
    };
    const int cin[3] = {1, 32, 64}, cout[3] = {32, 64, 128};
    std::vector<float> packed;
    auto al4 = [&]() {
        while (packed.size() % 4) packed.push_back(0.f);
    };
    size_t off_w[3], off_s[3];
    for (int l = 0; l < 3; ++l) {
        const float* w = take((int64_t)cout[l] * cin[l] * 9);
        const float *g = take(cout[l]), *bt = take(cout[l]), *mu = take(cout[l]), *var = take(cout[l]);
        std::vector<double> scale(cout[l]);
        for (int o = 0; o < cout[l]; ++o) scale[o] = (double)g[o] / std::sqrt((double)var[o] + (double)bn_eps);
        al4();
        off_w[l] = packed.size();
        if (l == 0) {                                 // [32][9]
            for (int o = 0; o < 32; ++o)
                for (int k = 0; k < 9; ++k) packed.push_back((float)((double)w[o * 9 + k] * scale[o]));
        } else {                                      // [tap][cin][cout]
            for (int k = 0; k < 9; ++k)
                for (int i = 0; i < cin[l]; ++i)
                    for (int o = 0; o < cout[l]; ++o)
                        packed.push_back((float)((double)w[((int64_t)o * cin[l] + i) * 9 + k] * scale[o]));
        }
        al4();
        off_s[l] = packed.size();
        for (int o = 0; o < cout[l]; ++o) packed.push_back((float)((double)bt[o] - (double)mu[o] * scale[o]));
    }
    size_t off_wih[2], off_bih[2], off_whh[2], off_bhh[2];
    for (int l = 0; l < 2; ++l) {
        const int in_sz = l == 0 ? gin : 512;
        const float *wih[2], *whh[2], *bih[2], *bhh[2];
        for (int d = 0; d < 2; ++d) {
            wih[d] = take((int64_t)768 * in_sz);
            whh[d] = take((int64_t)768 * 256);
            bih[d] = take(768);
            bhh[d] = take(768);
        }
        al4();
        off_wih[l] = packed.size();
        for (int d = 0; d < 2; ++d) packed.insert(packed.end(), wih[d], wih[d] + (int64_t)768 * in_sz);
        off_bih[l] = packed.size();
        for (int d = 0; d < 2; ++d) packed.insert(packed.end(), bih[d], bih[d] + 768);
        off_whh[l] = packed.size();
        for (int d = 0; d < 2; ++d) packed.insert(packed.end(), whh[d], whh[d] + (int64_t)768 * 256);
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
    SIR_CUDA(cudaMemcpy(m->weights.ptr, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    float* base = (float*)m->weights.ptr;
    m->w1 = base + off_w[0];
    m->sh1 = base + off_s[0];
    m->w2 = base + off_w[1];
    m->sh2 = base + off_s[1];
    m->w3 = base + off_w[2];
    m->sh3 = base + off_s[2];
    for (int l = 0; l < 2; ++l) {
        m->wih[l] = base + off_wih[l];
        m->bih[l] = base + off_bih[l];
        m->whh[l] = base + off_whh[l];
        m->bhh[l] = base + off_bhh[l];
    }
    m->att_w = base + off_att;
    m->fc_w = base + off_fcw;
    m->fc_b = base + off_fcb;
    m->loaded = true;
    return SIR_OK;
}

namespace sir {

constexpr int kModelChunk = 512;   // utterances per pass through the workspace

struct Workspace {
    float *act1, *act2, *gru_in, *gi, *y0, *y1, *h;
};

static size_t carve(Workspace& w, float* base, int B, int H, int W, int gin) {
    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, Tg = W4 / 2;
    size_t off = 0;
    auto next = [&](size_t n) {
        float* p = base ? base + off : nullptr;
        off += (n + 63) & ~(size_t)63;
        return p;
    };
    w.act1 = next((size_t)B * H2 * W2 * 32);
    w.act2 = next((size_t)B * H4 * W4 * 64);
    w.gru_in = next((size_t)B * Tg * gin);
    w.gi = next((size_t)B * Tg * 1536);
    w.y0 = next((size_t)B * Tg * 512);
    w.y1 = next((size_t)B * Tg * 512);
    w.h = next((size_t)2 * 2 * B * 256);
    return off;
}

int model_forward_chunk(sir_model* m, const Workspace& ws, const float* feat, int B, int W, float* logits,
                        cudaStream_t st) {
    const int H = m->n_mels;
    const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2, Tg = W4 / 2;
    {
        dim3 grid((unsigned)((H2 * W2 + 127) / 128), (unsigned)B);
        ProfScope ps("conv1_bn_relu_pool", st);
        conv1_bn_relu_pool_kernel<<<grid, 128, 0, st>>>(feat, m->w1, m->sh1, ws.act1, H, W);
        SIR_CHECK_LAUNCH("conv1_bn_relu_pool_kernel");
    }
    {
        constexpr size_t smem = (size_t)(180 * (32 + 4) + 32 * 64) * sizeof(float);
        dim3 grid((unsigned)((W2 + 15) / 16), (unsigned)((H2 + 7) / 8), (unsigned)B);
        ProfScope ps("conv2_bn_relu_pool", st);
        conv3x3_bn_relu_pool_kernel<32, 64, false><<<grid, 256, smem, st>>>(ws.act1, m->w2, m->sh2, ws.act2, H2, W2);
        SIR_CHECK_LAUNCH("conv3x3_bn_relu_pool_kernel<32,64>");
    }
    {
        constexpr size_t smem = (size_t)(180 * (64 + 4) + 64 * 128) * sizeof(float);
        dim3 grid((unsigned)((W4 + 15) / 16), (unsigned)((H4 + 7) / 8), (unsigned)B);
        ProfScope ps("conv3_bn_relu_pool", st);
        conv3x3_bn_relu_pool_kernel<64, 128, true><<<grid, 256, smem, st>>>(ws.act2, m->w3, m->sh3, ws.gru_in, H4, W4);
        SIR_CHECK_LAUNCH("conv3x3_bn_relu_pool_kernel<64,128>");
    }
    const float* x = ws.gru_in;
    int in_sz = m->gru_in;
    float* ys[2] = {ws.y0, ws.y1};
    for (int l = 0; l < 2; ++l) {
        const int M = B * Tg;
        dim3 ggrid(1536 / 128, (unsigned)((M + 127) / 128));
        {
            ProfScope ps(l == 0 ? "gru_l0_input_gemm" : "gru_l1_input_gemm", st);
            gemm_nt_bias_kernel<<<ggrid, 256, 0, st>>>(x, m->wih[l], m->bih[l], ws.gi, M, 1536, in_sz);
        }
        SIR_CHECK_LAUNCH("gemm_nt_bias_kernel");
        ProfScope ps(l == 0 ? "gru_l0_recurrence" : "gru_l1_recurrence", st);
        {
            dim3 rgrid((unsigned)(kGruCluster * ((B + kGruBatch - 1) / kGruBatch)), 2);
            gru_layer_kernel<<<rgrid, 256, kGruSmemBytes, st>>>(ws.gi, m->whh[l], m->bhh[l], ys[l], B, Tg);
            SIR_CHECK_LAUNCH("gru_layer_kernel");
        }
        x = ys[l];
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
    static bool attr_done = false;
    if (!attr_done) {
        SIR_CUDA(cudaFuncSetAttribute(conv3x3_bn_relu_pool_kernel<64, 128, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((180 * 68 + 64 * 128) * sizeof(float))));
        SIR_CUDA(cudaFuncSetAttribute(conv3x3_bn_relu_pool_kernel<32, 64, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((180 * 36 + 32 * 64) * sizeof(float))));
        SIR_CUDA(cudaFuncSetAttribute(gru_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kGruSmemBytes));
        attr_done = true;
    }
    chunk = batch < kModelChunk ? batch : kModelChunk;
    Workspace probe;
    const size_t need = carve(probe, nullptr, chunk, m->n_mels, n_frames, m->gru_in) * sizeof(float);
    int rc = m->work.reserve(need);
    if (rc != SIR_OK) return rc;
    carve(ws, (float*)m->work.ptr, chunk, m->n_mels, n_frames, m->gru_in);
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
    if (d_features) {
        int rc = sir_frontend_forward(fe, d_wave, wave_stride, d_lengths, n_samples, batch, max_samples,
                                      SIR_OUT_LOGMEL_NORM, out_frames, d_features, nullptr, nullptr, stream);
        if (rc != SIR_OK) return rc;
        return sir_model_forward(m, d_features, batch, out_frames, d_logits, stream);
    }
    return fail(SIR_ERR_INVALID, "sir_pipeline_forward: d_features is required in this build");
}
