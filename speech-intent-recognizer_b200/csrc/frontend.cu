// Fused log-mel frontend for sm_100a: waveform -> reflect-padded framing -> Hann -> 1024-point real FFT ->
// |X|^2 -> sparse HTK mel -> 10 log10 -> per-utterance (x-mean)/(std+1e-5) -> SpecAugment bands -> pad/trim.
//
// Replaces (per utterance) scripts/precompute_features.py:59-73, scripts/dataset.py:105-113,160-176 of the
// reference and the torchaudio calls behind them with ONE kernel launch for a batch.
//
// Mapping
//   * one thread-block CLUSTER per utterance (1, 2, 4 or 8 CTAs, picked so the grid covers the 148 SMs a
//     few times over even for small batches); the cluster's CTAs take 8-frame groups round-robin;
//   * each HALF-WARP owns one frame and reads its 1024 samples straight from global memory (128 contiguous
//     bytes per half-warp request; the 50 % overlap with the neighbouring frame, held by the other half of
//     the same warp, is served by L1/L2, so HBM sees every sample once); only the first and last frame of an
//     utterance take the scalar path that resolves torch.stft's reflect padding;
//   * per frame: 512-point complex FFT as 32-point x 16-point register FFTs with one
//     shared-memory transposition (logmel_frame.cuh), real-FFT post-pass, power, sparse mel taps, log10;
//   * un-normalised values go to the output through an 8-frame shared tile (32-byte row segments), the
//     per-utterance mean / unbiased std are combined across the cluster through distributed shared memory
//     (shifted sums + Chan's merge), then each CTA re-reads its own L2-resident values, normalises, applies
//     the mask bands and writes the zero padding.  HBM sees each sample once and each feature once.
//
// Roofline: HBM-bound by intent (4 L + 4 n_mels T bytes per utterance), but at ~19 k fp32 instructions per
// frame the kernel sits at the fp32-issue ridge; see DESIGN.md for the arithmetic.
#include <cooperative_groups.h>

#include <cmath>
#include <vector>

#include "frontend_tables.h"
#include "logmel_frame.cuh"
#include "philox.cuh"
#include "sir_common.cuh"

namespace cg = cooperative_groups;

namespace sir {

thread_local char g_error[512] = "";
std::atomic<int64_t> g_launches{0};
bool g_profile = false;

struct ProfRecord {
    const char* name;
    cudaEvent_t begin, end;
};
static std::vector<ProfRecord> g_prof_records;

void prof_mark(const char* name, cudaStream_t st, bool begin) {
    if (begin) {
        ProfRecord r{name, nullptr, nullptr};
        cudaEventCreate(&r.begin);
        cudaEventCreate(&r.end);
        cudaEventRecord(r.begin, st);
        g_prof_records.push_back(r);
    } else {
        for (size_t i = g_prof_records.size(); i-- > 0;)
            if (g_prof_records[i].name == name) {
                cudaEventRecord(g_prof_records[i].end, st);
                break;
            }
    }
}

constexpr int kFeWarps = 4;
constexpr int kFeThreads = kFeWarps * 32;
constexpr int kGroupFrames = 2 * kFeWarps;                 // one frame per half-warp
constexpr int kTileStride = kGroupFrames + 1;              // padded to dodge bank conflicts
constexpr int kMelWeightCap = 1280;

// shared-memory carve-up (float offsets)
constexpr int kOffWindow = 0;
constexpr int kOffTw512 = kOffWindow + 1024;
constexpr int kOffTw1024 = kOffTw512 + 1024;
constexpr int kOffMelStart = kOffTw1024 + 516;
constexpr int kOffMelCount = kOffMelStart + kMaxMels;
constexpr int kOffMelOffset = kOffMelCount + kMaxMels;
constexpr int kOffMelWeight = kOffMelOffset + kMaxMels;
constexpr int kOffScratch = kOffMelWeight + kMelWeightCap;
constexpr int kOffTile = kOffScratch + kGroupFrames * kFrameScratch;
constexpr int kOffReduce = kOffTile + kMaxMels * kTileStride;
constexpr int kFeSmemFloats = kOffReduce + 64;
constexpr size_t kFeSmemBytes = (size_t)kFeSmemFloats * sizeof(float);
static_assert(kOffScratch % 4 == 0 && kOffTile % 4 == 0, "16-byte alignment of vector regions");
#ifndef SIR_FE_MIN_CTAS
#define SIR_FE_MIN_CTAS 3
#endif
static_assert(SIR_FE_MIN_CTAS * (kFeSmemBytes + 1024) <= 232448, "resident CTAs per SM");

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
};

// Interior frames read their 1024 samples straight from global memory: lane l takes the 8-byte words
// l + 16 j, so a half-warp covers 128 contiguous bytes per request, and the 50 % overlap with the neighbouring
// frame (the other half of the same warp) is served by L1/L2 - HBM still sees every sample once.
template <typename T>
struct GlobalFrame;
template <>
struct GlobalFrame<float> {
    const float2* base;
    __device__ __forceinline__ explicit GlobalFrame(const float* p) : base(reinterpret_cast<const float2*>(p)) {}
    __device__ __forceinline__ F2 operator()(int n) const {
        const float2 v = __ldg(base + n);
        return F2{v.x, v.y};
    }
    static constexpr uintptr_t kAlignMask = 7u;
};
template <>
struct GlobalFrame<short> {                       // PCM16: one 4-byte word per sample pair
    const short2* base;
    __device__ __forceinline__ explicit GlobalFrame(const short* p) : base(reinterpret_cast<const short2*>(p)) {}
    __device__ __forceinline__ F2 operator()(int n) const {
        const short2 v = __ldg(base + n);
        return F2{sample_to_float(v.x), sample_to_float(v.y)};
    }
    static constexpr uintptr_t kAlignMask = 3u;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename SampleT>
__global__ void __launch_bounds__(kFeThreads, SIR_FE_MIN_CTAS) logmel_frontend_kernel(const FrontendParams p) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize;
    const int tid = threadIdx.x;

    // ---- constants into shared memory ------------------------------------------------------------------
    for (int i = tid; i < 1024; i += kFeThreads) {
        smem[kOffWindow + i] = p.tables.window[i];
        smem[kOffTw512 + i] = p.tables.tw512[i];
    }
    for (int i = tid; i < 514; i += kFeThreads) smem[kOffTw1024 + i] = p.tables.tw1024[i];
    int* s_mel_start = reinterpret_cast<int*>(smem + kOffMelStart);
    int* s_mel_count = reinterpret_cast<int*>(smem + kOffMelCount);
    int* s_mel_offset = reinterpret_cast<int*>(smem + kOffMelOffset);
    for (int i = tid; i < p.n_mels; i += kFeThreads) {
        s_mel_start[i] = p.tables.mel_start[i];
        s_mel_count[i] = p.tables.mel_count[i];
        s_mel_offset[i] = p.tables.mel_offset[i];
    }
    for (int i = tid; i < p.mel_weight_count; i += kFeThreads) smem[kOffMelWeight + i] = p.tables.mel_weight[i];
    const FrontendTables st{smem + kOffWindow, smem + kOffTw512, smem + kOffTw1024, s_mel_start,
                            s_mel_count,       s_mel_offset,     smem + kOffMelWeight};

    // ---- this utterance --------------------------------------------------------------------------------
    int L = p.lengths ? min(p.lengths[b], p.n_samples) : p.n_samples;
    if (p.max_samples > 0) L = min(L, p.max_samples);
    const bool valid = L > kNfft / 2;                       // reflect padding needs L > 512
    const int T = valid ? 1 + L / kHop : 0;
    const int n_groups = (T + kGroupFrames - 1) / kGroupFrames;
    const SampleT* __restrict__ row = static_cast<const SampleT*>(p.wave) + (int64_t)b * p.wave_stride;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & GlobalFrame<SampleT>::kAlignMask) == 0);
    const bool mfcc = p.mode == SIR_OUT_MFCC;
    // where the group loop writes its (un-normalised) values: the output itself, or the dB staging buffer for MFCC
    const int row_stride = mfcc ? p.stage_frames : p.out_frames;
    float* __restrict__ out = mfcc ? p.db_stage + (int64_t)b * p.n_mels * p.stage_frames
                                   : p.out + (int64_t)b * p.n_mels * p.out_frames;
    const bool out_vec = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (p.status && crank == 0 && tid == 0) p.status[b] = valid ? 0 : 1;

    float* tile = smem + kOffTile;
    float* red = smem + kOffReduce;
    const int lane = tid & 31, warp = tid >> 5, half = lane >> 4, q = lane & 15;
    const int slot = 2 * warp + half;
    float* scr = smem + kOffScratch + slot * kFrameScratch;

    float shift = 0.f, s1 = 0.f, s2 = 0.f, vmax = -INFINITY;
    bool have_shift = false;
    __syncthreads();                                        // tables visible

    for (int g = crank; g < n_groups; g += csize) {
        const int t0 = g * kGroupFrames;
        const int t = t0 + slot;
        const bool active = t < T;
        if (active) {
            // frame t covers original samples [512 t - 512, 512 t + 512)
            const int first = t * kHop - kNfft / 2;
            if (vec_ok && first >= 0 && first + kNfft <= L) {
                const GlobalFrame<SampleT> ld(row + first);
                frame_phase_a(q, ld, st.window, st.tw512, scr, scr + 16 * kRowPad);
            } else {
                const ReflectFrame<SampleT> ld{row, first, L};
                frame_phase_a(q, ld, st.window, st.tw512, scr, scr + 16 * kRowPad);
            }
        }
        __syncwarp();
        {
            PhaseBRegs rb;
            if (active) frame_phase_b_load(q, scr, scr + 16 * kRowPad, rb);
            __syncwarp();
            if (active) frame_phase_b_store(q, rb, scr);
        }
        __syncwarp();
        {
            PhaseCRegs rc;
            if (active) frame_phase_c_compute(q, scr, st.tw1024, rc);
            __syncwarp();
            if (active) frame_phase_c_store(q, rc, scr);
        }
        __syncwarp();
        if (active) {
            for (int j = 0; 16 * j < p.n_mels; ++j) {
                const int m = mel_of_lane(q, j);
                if (m < p.n_mels) {
                    float v = mel_band_power(m, scr, st);
                    if (p.mode != SIR_OUT_MEL_POWER) v = 10.0f * log10f(fmaxf(v, 1e-10f));
                    tile[m * kTileStride + slot] = v;
                }
            }
        }
        __syncthreads();                                    // tile complete

        // tile -> global (8 consecutive frames of one mel row = one 32-byte segment) + statistics
        if (!have_shift) {
            shift = tile[0];
            have_shift = true;
        }
        const int nslots = min(kGroupFrames, T - t0);
        if (out_vec && nslots == kGroupFrames && t0 + kGroupFrames <= row_stride) {
            for (int idx = tid; idx < p.n_mels * 2; idx += kFeThreads) {
                const int m = idx >> 1, h4 = (idx & 1) * 4;
                const float* tp = tile + m * kTileStride + h4;
                const float4 v = make_float4(tp[0], tp[1], tp[2], tp[3]);
                *reinterpret_cast<float4*>(out + (int64_t)m * row_stride + t0 + h4) = v;
                vmax = fmaxf(vmax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                const float d0 = v.x - shift, d1 = v.y - shift, d2 = v.z - shift, d3 = v.w - shift;
                s1 += (d0 + d1) + (d2 + d3);
                s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        } else {
            for (int idx = tid; idx < p.n_mels * kGroupFrames; idx += kFeThreads) {
                const int m = idx >> 3, s = idx & 7;
                if (s < nslots) {
                    const float v = tile[m * kTileStride + s];
                    if (t0 + s < row_stride) out[(int64_t)m * row_stride + t0 + s] = v;
                    vmax = fmaxf(vmax, v);
                    const float d = v - shift;
                    s1 += d;
                    s2 += d * d;
                }
            }
        }
        __syncthreads();                                    // tile consumed before the next group overwrites it
    }

    if ((p.mode != SIR_OUT_LOGMEL_NORM && !mfcc) || !valid) {
        // zero padding (and whole rows of invalid utterances); rows are split over the cluster
        const int first = valid ? min(T, p.out_frames) : 0;
        const int rows = mfcc ? p.n_mfcc : p.n_mels;
        float* __restrict__ o = p.out + (int64_t)b * rows * p.out_frames;
        for (int m = crank; m < rows; m += csize)
            for (int t = first + tid; t < p.out_frames; t += kFeThreads) o[(int64_t)m * p.out_frames + t] = 0.f;
        return;                                             // uniform over the whole cluster
    }
    if (mfcc) {
        // ---- MFCC: per-utterance max over the cluster -> clamp at max - top_db -> DCT of every frame ------------
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 16));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 8));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 4));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 2));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 1));
        __syncthreads();
        if (lane == 0) red[warp] = vmax;
        __syncthreads();
        if (tid == 0) {
            float mx = red[0];
            for (int w = 1; w < kFeWarps; ++w) mx = fmaxf(mx, red[w]);
            red[16] = mx;
        }
        float* s_dct = smem + kOffScratch;                  // the frame scratch is free now: [n_mels][n_mfcc]
        for (int i = tid; i < p.n_mels * p.n_mfcc; i += kFeThreads) s_dct[i] = p.dct[i];
        cluster.sync();
        float mx = -INFINITY;
        for (int r = 0; r < csize; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(red + 16, r));
        const float floor_db = p.top_db > 0.f ? mx - p.top_db : -INFINITY;
        float* __restrict__ o = p.out + (int64_t)b * p.n_mfcc * p.out_frames;
        for (int g = crank; g < n_groups; g += csize) {
            const int t0 = g * kGroupFrames;
            const int nslots = min(min(kGroupFrames, T - t0), p.out_frames - t0);
            for (int idx = tid; idx < p.n_mfcc * kGroupFrames; idx += kFeThreads) {
                const int c = idx >> 3, sl = idx & 7;
                if (sl < nslots) {
                    float acc = 0.f;
                    for (int m = 0; m < p.n_mels; ++m)
                        acc = fmaf(fmaxf(__ldcg(out + (int64_t)m * row_stride + t0 + sl), floor_db), s_dct[m * p.n_mfcc + c], acc);
                    o[(int64_t)c * p.out_frames + t0 + sl] = acc;
                }
            }
        }
        for (int c = crank; c < p.n_mfcc; c += csize)
            for (int t = T + tid; t < p.out_frames; t += kFeThreads) o[(int64_t)c * p.out_frames + t] = 0.f;
        cluster.sync();                                     // keep every CTA's shared memory alive for its peers
        return;
    }

    // ---- per-utterance statistics: CTA partial (n, mean, M2) -> cluster merge over DSMEM ----------------
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    __syncthreads();
    if (lane == 0) {
        red[warp] = s1;
        red[8 + warp] = s2;
    }
    __syncthreads();
    if (tid == 0) {
        int my_frames = 0;
        for (int g = crank; g < n_groups; g += csize) my_frames += min(kGroupFrames, T - g * kGroupFrames);
        const double n = (double)my_frames * (double)p.n_mels;
        double S1 = 0, S2 = 0;
        for (int w = 0; w < kFeWarps; ++w) {
            S1 += (double)red[w];
            S2 += (double)red[8 + w];
        }
        double mean = 0, m2 = 0;
        if (n > 0) {
            mean = (double)shift + S1 / n;
            m2 = fmax(S2 - S1 * S1 / n, 0.0);
        }
        double* part = reinterpret_cast<double*>(red + 16);
        part[0] = n;
        part[1] = mean;
        part[2] = m2;
    }
    cluster.sync();
    {
        double n = 0, mean = 0, m2 = 0;                     // every thread merges in rank order: deterministic
        for (int r = 0; r < csize; ++r) {
            const double* part = reinterpret_cast<const double*>(cluster.map_shared_rank(red + 16, r));
            const double nb = part[0], mb = part[1], m2b = part[2];
            if (nb > 0) {
                const double nt = n + nb, delta = mb - mean;
                mean += delta * nb / nt;
                m2 += m2b + delta * delta * n * nb / nt;
                n = nt;
            }
        }
        const float fmean = (float)mean;
        const float inv = (float)(1.0 / (sqrt(m2 / (n - 1.0)) + 1e-5));
        int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
        if (p.masks) {
            mt0 = p.masks[4 * b + 0];
            mt1 = p.masks[4 * b + 1];
            mf0 = p.masks[4 * b + 2];
            mf1 = p.masks[4 * b + 3];
        }
        // normalise this CTA's own frames (still L2-resident; ld.cg so no stale L1 line is read)
        for (int g = crank; g < n_groups; g += csize) {
            const int t0 = g * kGroupFrames;
            const int nslots = min(min(kGroupFrames, T - t0), p.out_frames - t0);
            for (int idx = tid; idx < p.n_mels * kGroupFrames; idx += kFeThreads) {
                const int m = idx >> 3, s = idx & 7;
                if (s < nslots) {
                    float* addr = out + (int64_t)m * p.out_frames + t0 + s;
                    float v = (__ldcg(addr) - fmean) * inv;
                    const int t = t0 + s;
                    if ((t >= mt0 && t < mt1) || (m >= mf0 && m < mf1)) v = 0.f;
                    *addr = v;
                }
            }
        }
        for (int m = crank; m < p.n_mels; m += csize)
            for (int t = T + tid; t < p.out_frames; t += kFeThreads) out[(int64_t)m * p.out_frames + t] = 0.f;
    }
    cluster.sync();                                         // keep every CTA's shared memory alive for its peers
}

// ---- small companions ----------------------------------------------------------------------------------
__global__ void amplitude_to_db_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = 10.0f * log10f(fmaxf(in[i], 1e-10f));
}

// One thread per utterance: the draws of scripts/dataset.py:105,166-171 + torchaudio mask_along_axis.
__global__ void specaugment_sample_kernel(uint64_t seed, uint64_t first_index, int batch, int n_mels, int n_frames,
                                          const int32_t* __restrict__ frames, float augment_prob, int tparam,
                                          int fparam, int32_t* __restrict__ masks) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    uint32_t a[4], c[4];
    philox4x32_10(seed, first_index + (uint64_t)b, 0u, a);
    philox4x32_10(seed, first_index + (uint64_t)b, 1u, c);
    const float gate_aug = u01(a[0]);
    const float u[6] = {u01(a[1]), u01(a[2]), u01(a[3]), u01(c[0]), u01(c[1]), u01(c[2])};
    const int T = frames ? frames[b] : n_frames;
    int m[4] = {0, 0, 0, 0};
    if (gate_aug < augment_prob) {
        const int param[2] = {tparam, fparam};
        const int size[2] = {T, n_mels};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (u[3 * k] < 0.5f && param[k] >= 1) {
                const float value = __fmul_rn(u[3 * k + 1], (float)param[k]);
                const float min_value = __fmul_rn(u[3 * k + 2], __fsub_rn((float)size[k], value));
                const int start = (int)min_value;
                m[2 * k] = start;
                m[2 * k + 1] = start + (int)value;
            }
        }
    }
    reinterpret_cast<int4*>(masks)[b] = make_int4(m[0], m[1], m[2], m[3]);
}

__global__ void features_finalize_kernel(const float* __restrict__ in, int n_mels, int in_frames,
                                         const int32_t* __restrict__ frames, const int32_t* __restrict__ masks,
                                         int out_frames, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int T = frames ? min(frames[b], in_frames) : in_frames;
    int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
    if (masks) {
        mt0 = masks[4 * b];
        mt1 = masks[4 * b + 1];
        mf0 = masks[4 * b + 2];
        mf1 = masks[4 * b + 3];
    }
    const int total = n_mels * out_frames;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int m = i / out_frames, t = i - m * out_frames;
        float v = 0.f;
        if (t < T) {
            v = in[((int64_t)b * n_mels + m) * in_frames + t];
            if ((t >= mt0 && t < mt1) || (m >= mf0 && m < mf1)) v = 0.f;
        }
        out[((int64_t)b * n_mels + m) * out_frames + t] = v;
    }
}

}  // namespace sir

// ---- C ABI ------------------------------------------------------------------------------------------------
using namespace sir;

struct sir_frontend {
    int device = 0;
    int sample_rate = 16000, n_mels = 64;
    int num_sms = 148;
    int mel_weight_count = 0;
    DeviceBuffer tables;
    FrontendTables dev{};
    DeviceBuffer db_stage;       // MFCC: staged dB values
    DeviceBuffer dct;            // MFCC: [n_mels][n_mfcc] ortho DCT-II
    int dct_n_mfcc = 0;
};

extern "C" const char* sir_last_error(void) { return g_error; }
extern "C" int sir_version(void) { return 100; }
extern "C" int64_t sir_launch_count(void) { return g_launches.load(); }

extern "C" void sir_profile_enable(int on) { g_profile = on != 0; }

// Aggregates the recorded stages by name: names are written as a ';'-separated list, ms[i] / calls[i] hold the
// total device time and number of occurrences.  Clears the records.  Synchronises the device.
extern "C" int sir_profile_read(char* names, int names_cap, float* ms, int* calls, int max_stages) {
    cudaDeviceSynchronize();
    std::vector<const char*> order;
    std::vector<float> total;
    std::vector<int> count;
    for (auto& r : g_prof_records) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.begin, r.end);
        cudaEventDestroy(r.begin);
        cudaEventDestroy(r.end);
        size_t i = 0;
        for (; i < order.size(); ++i)
            if (order[i] == r.name) break;
        if (i == order.size()) {
            order.push_back(r.name);
            total.push_back(0.f);
            count.push_back(0);
        }
        total[i] += t;
        count[i] += 1;
    }
    g_prof_records.clear();
    int n = 0, pos = 0;
    if (names && names_cap > 0) names[0] = 0;
    for (size_t i = 0; i < order.size() && n < max_stages; ++i, ++n) {
        ms[n] = total[i];
        calls[n] = count[i];
        if (names) pos += snprintf(names + pos, pos < names_cap ? names_cap - pos : 0, "%s%s", n ? ";" : "", order[i]);
    }
    return n;
}

extern "C" int sir_frontend_create(sir_frontend** out, int sample_rate, int n_mels, int n_fft, int hop_length) {
    if (!out) return fail(SIR_ERR_INVALID, "sir_frontend_create: out is NULL");
    *out = nullptr;
    if (n_fft != kNfft || hop_length != kHop)
        return fail(SIR_ERR_UNSUPPORTED, "only n_fft=1024 / hop_length=512 are implemented (got %d / %d)", n_fft,
                    hop_length);
    if (n_mels < 1 || n_mels > kMaxMels) return fail(SIR_ERR_UNSUPPORTED, "n_mels must be in [1,%d]", kMaxMels);
    if (sample_rate < 2) return fail(SIR_ERR_INVALID, "bad sample_rate %d", sample_rate);
    HostFrontendTables h = build_frontend_tables(sample_rate, n_mels);
    if ((int)h.mel_weight.size() > kMelWeightCap)
        return fail(SIR_ERR_UNSUPPORTED, "filterbank has %zu taps (cap %d)", h.mel_weight.size(), kMelWeightCap);
    sir_frontend* fe = new sir_frontend();
    fe->sample_rate = sample_rate;
    fe->n_mels = n_mels;
    fe->mel_weight_count = (int)h.mel_weight.size();
    cudaError_t e = cudaGetDevice(&fe->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&fe->num_sms, cudaDevAttrMultiProcessorCount, fe->device);
    if (e != cudaSuccess) {
        delete fe;
        return fail(SIR_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
    }
    // one allocation, 16-byte aligned sections
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_win = 0, o_512 = al(o_win + 1024 * 4), o_1024 = al(o_512 + 1024 * 4),
                 o_ms = al(o_1024 + 514 * 4), o_mc = al(o_ms + n_mels * 4), o_mo = al(o_mc + n_mels * 4),
                 o_mw = al(o_mo + n_mels * 4), total = al(o_mw + h.mel_weight.size() * 4);
    int rc = fe->tables.reserve(total);
    if (rc != SIR_OK) {
        delete fe;
        return rc;
    }
    char* base = (char*)fe->tables.ptr;
    struct Up {
        size_t off;
        const void* src;
        size_t bytes;
    } ups[] = {{o_win, h.window.data(), 1024 * 4},          {o_512, h.tw512.data(), 1024 * 4},
               {o_1024, h.tw1024.data(), 514 * 4},          {o_ms, h.mel_start.data(), (size_t)n_mels * 4},
               {o_mc, h.mel_count.data(), (size_t)n_mels * 4}, {o_mo, h.mel_offset.data(), (size_t)n_mels * 4},
               {o_mw, h.mel_weight.data(), h.mel_weight.size() * 4}};
    for (auto& u : ups) {
        e = cudaMemcpy(base + u.off, u.src, u.bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            fe->tables.release();
            delete fe;
            return fail(SIR_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(e));
        }
    }
    fe->dev = FrontendTables{(const float*)(base + o_win), (const float*)(base + o_512), (const float*)(base + o_1024),
                             (const int*)(base + o_ms),    (const int*)(base + o_mc),    (const int*)(base + o_mo),
                             (const float*)(base + o_mw)};
    e = cudaFuncSetAttribute(logmel_frontend_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFeSmemBytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(logmel_frontend_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFeSmemBytes);
    if (e != cudaSuccess) {
        fe->tables.release();
        delete fe;
        return fail(SIR_ERR_CUDA, "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
    }
    *out = fe;
    return SIR_OK;
}

extern "C" void sir_frontend_destroy(sir_frontend* fe) {
    if (!fe) return;
    fe->tables.release();
    fe->db_stage.release();
    fe->dct.release();
    delete fe;
}

static int frontend_launch(sir_frontend* fe, const void* d_wave, bool pcm16, int64_t wave_stride, const int32_t* d_lengths,
                           int n_samples, int batch, int max_samples, int mode, int out_frames, float* d_out,
                           const int32_t* d_masks, int32_t* d_status, void* stream, int n_mfcc = 0, float top_db = 0.f) {
    if (!fe || !d_wave || !d_out) return fail(SIR_ERR_INVALID, "sir_frontend_forward: NULL handle or buffer");
    if (batch < 0 || n_samples < 0 || out_frames < 1 || wave_stride < n_samples)
        return fail(SIR_ERR_INVALID, "sir_frontend_forward: bad sizes (batch %d, n_samples %d, out_frames %d)", batch,
                    n_samples, out_frames);
    if (mode < SIR_OUT_MEL_POWER || mode > SIR_OUT_MFCC) return fail(SIR_ERR_INVALID, "bad mode %d", mode);
    if (d_masks && mode != SIR_OUT_LOGMEL_NORM)
        return fail(SIR_ERR_INVALID, "mask bands are applied to normalised features only");
    if (batch == 0) return SIR_OK;
    FrontendParams p{};
    p.wave = d_wave;
    p.wave_stride = wave_stride;
    p.lengths = d_lengths;
    p.n_samples = n_samples;
    p.max_samples = max_samples;
    p.n_mels = fe->n_mels;
    p.mode = mode;
    p.out_frames = out_frames;
    p.out = d_out;
    p.masks = d_masks;
    p.status = d_status;
    p.tables = fe->dev;
    p.mel_weight_count = fe->mel_weight_count;
    if (mode == SIR_OUT_MFCC) {
        if (n_mfcc < 1 || n_mfcc > fe->n_mels || fe->n_mels * n_mfcc > kGroupFrames * kFrameScratch)
            return fail(SIR_ERR_INVALID, "sir_frontend_mfcc: n_mfcc must be in [1, n_mels] (got %d)", n_mfcc);
        if (fe->dct_n_mfcc != n_mfcc) {                      // create_dct(n_mfcc, n_mels, "ortho"), TA:functional.py:636-665
            const double pi = 3.14159265358979323846;
            std::vector<float> h((size_t)fe->n_mels * n_mfcc);
            for (int m = 0; m < fe->n_mels; ++m)
                for (int c = 0; c < n_mfcc; ++c) {
                    double v = std::cos(pi / fe->n_mels * (m + 0.5) * c) * std::sqrt(2.0 / fe->n_mels);
                    if (c == 0) v *= 1.0 / std::sqrt(2.0);
                    h[(size_t)m * n_mfcc + c] = (float)v;
                }
            int rc = fe->dct.reserve(h.size() * sizeof(float));
            if (rc != SIR_OK) return rc;
            SIR_CUDA(cudaMemcpyAsync(fe->dct.ptr, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
            SIR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));      // h goes out of scope
            fe->dct_n_mfcc = n_mfcc;
        }
        const int eff0 = max_samples > 0 && max_samples < n_samples ? max_samples : n_samples;
        p.stage_frames = ((1 + eff0 / kHop) + 3) & ~3;
        int rc = fe->db_stage.reserve((size_t)batch * fe->n_mels * p.stage_frames * sizeof(float));
        if (rc != SIR_OK) return rc;
        p.db_stage = (float*)fe->db_stage.ptr;
        p.dct = (const float*)fe->dct.ptr;
        p.n_mfcc = n_mfcc;
        p.top_db = top_db;
    }
    // cluster size: enough CTAs for ~2 waves of 3 CTAs/SM, never more CTAs than 8-frame groups
    int eff = max_samples > 0 && max_samples < n_samples ? max_samples : n_samples;
    const int groups = (1 + eff / kHop + kGroupFrames - 1) / kGroupFrames;
    int csize = 1;
    while (csize < 8 && (int64_t)batch * csize < (int64_t)fe->num_sms * 6 && csize * 2 <= groups) csize *= 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(batch * csize));
    cfg.blockDim = dim3(kFeThreads);
    cfg.dynamicSmemBytes = kFeSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
        ProfScope ps("logmel_frontend_kernel", cfg.stream);
        if (pcm16)
            SIR_CUDA(cudaLaunchKernelEx(&cfg, logmel_frontend_kernel<short>, p));
        else
            SIR_CUDA(cudaLaunchKernelEx(&cfg, logmel_frontend_kernel<float>, p));
    }
    SIR_CHECK_LAUNCH("logmel_frontend_kernel");
    return SIR_OK;
}

extern "C" int sir_frontend_forward(sir_frontend* fe, const float* d_wave, int64_t wave_stride,
                                    const int32_t* d_lengths, int n_samples, int batch, int max_samples, int mode,
                                    int out_frames, float* d_out, const int32_t* d_masks, int32_t* d_status,
                                    void* stream) {
    return frontend_launch(fe, d_wave, false, wave_stride, d_lengths, n_samples, batch, max_samples, mode, out_frames, d_out,
                           d_masks, d_status, stream);
}

extern "C" int sir_frontend_mfcc(sir_frontend* fe, const float* d_wave, int64_t wave_stride, const int32_t* d_lengths,
                                 int n_samples, int batch, int max_samples, int n_mfcc, float top_db, int out_frames,
                                 float* d_out, int32_t* d_status, void* stream) {
    return frontend_launch(fe, d_wave, false, wave_stride, d_lengths, n_samples, batch, max_samples, SIR_OUT_MFCC, out_frames,
                           d_out, nullptr, d_status, stream, n_mfcc, top_db);
}

// y[n] = x[n] - coeff * x[n-1], y[0] = x[0] per row (torchaudio.functional.preemphasis, TA:functional.py:2426)
namespace sir {
__global__ void preemphasis_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t stride, int n, float coeff) {
    const float* row = in + (int64_t)blockIdx.y * stride;
    float* orow = out + (int64_t)blockIdx.y * stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        orow[i] = i > 0 ? row[i] - coeff * row[i - 1] : row[i];
}
}  // namespace sir

extern "C" int sir_preemphasis(const float* d_in, float* d_out, int64_t stride, int n_samples, int batch, float coeff,
                               void* stream) {
    if (!d_in || !d_out || d_in == d_out || n_samples < 0 || batch < 0 || stride < n_samples)
        return fail(SIR_ERR_INVALID, "sir_preemphasis: bad arguments (in-place is not supported)");
    if (batch == 0 || n_samples == 0) return SIR_OK;
    if (batch > 65535) return fail(SIR_ERR_UNSUPPORTED, "sir_preemphasis: batch > 65535");
    dim3 grid((unsigned)((n_samples + 1023) / 1024 < 64 ? (n_samples + 1023) / 1024 : 64), (unsigned)batch);
    preemphasis_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, stride, n_samples, coeff);
    SIR_CHECK_LAUNCH("preemphasis_kernel");
    return SIR_OK;
}

extern "C" int sir_frontend_forward_pcm16(sir_frontend* fe, const int16_t* d_pcm, int64_t wave_stride,
                                          const int32_t* d_lengths, int n_samples, int batch, int max_samples, int mode,
                                          int out_frames, float* d_out, const int32_t* d_masks, int32_t* d_status,
                                          void* stream) {
    return frontend_launch(fe, d_pcm, true, wave_stride, d_lengths, n_samples, batch, max_samples, mode, out_frames, d_out,
                           d_masks, d_status, stream);
}

extern "C" int sir_amplitude_to_db(const float* d_in, float* d_out, int64_t n, void* stream) {
    if (!d_in || !d_out || n < 0) return fail(SIR_ERR_INVALID, "sir_amplitude_to_db: bad arguments");
    if (n == 0) return SIR_OK;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    amplitude_to_db_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n);
    SIR_CHECK_LAUNCH("amplitude_to_db_kernel");
    return SIR_OK;
}

extern "C" int sir_specaugment_sample(uint64_t seed, uint64_t first_index, int batch, int n_mels, int n_frames,
                                      const int32_t* d_frames, float augment_prob, int time_mask_param,
                                      int freq_mask_param, int32_t* d_masks, void* stream) {
    if (!d_masks || batch < 0) return fail(SIR_ERR_INVALID, "sir_specaugment_sample: bad arguments");
    if ((reinterpret_cast<uintptr_t>(d_masks) & 15u) != 0)
        return fail(SIR_ERR_INVALID, "sir_specaugment_sample: d_masks must be 16-byte aligned");
    if (batch == 0) return SIR_OK;
    specaugment_sample_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        seed, first_index, batch, n_mels, n_frames, d_frames, augment_prob, time_mask_param, freq_mask_param, d_masks);
    SIR_CHECK_LAUNCH("specaugment_sample_kernel");
    return SIR_OK;
}

extern "C" int sir_features_finalize(const float* d_in, int batch, int n_mels, int in_frames,
                                     const int32_t* d_frames, const int32_t* d_masks, int out_frames, float* d_out,
                                     void* stream) {
    if (!d_in || !d_out || batch < 0 || n_mels < 1 || in_frames < 0 || out_frames < 1)
        return fail(SIR_ERR_INVALID, "sir_features_finalize: bad arguments");
    if (batch == 0) return SIR_OK;
    if (batch > 65535) return fail(SIR_ERR_UNSUPPORTED, "sir_features_finalize: batch > 65535");
    const int total = n_mels * out_frames;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)batch);
    features_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, n_mels, in_frames, d_frames, d_masks,
                                                                     out_frames, d_out);
    SIR_CHECK_LAUNCH("features_finalize_kernel");
    return SIR_OK;
}

// ---- polyphase sinc resampler ---------------------------------------------------------------------------------
// torchaudio.transforms.Resample(orig, new) (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) - the step the
// reference applies before the frontend when a file is not 16 kHz (scripts/precompute_features.py:54-56,
// scripts/test_model.py:69-72; the TTS clips are 24 kHz).  out[m * new + p] = sum_k kernel[p][k] x_pad[m * orig + k],
// x_pad = x padded by (width, width + orig) zeros; one thread per output sample, the kernel table in shared memory.
namespace sir {
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ in, int64_t in_stride, const int32_t* __restrict__ in_len,
                                                       int n_in, const float* __restrict__ taps, int n_taps, int width, int orig,
                                                       int nw, int staged, float* __restrict__ out, int64_t out_stride,
                                                       int n_out_cap, int32_t* __restrict__ out_len) {
    extern __shared__ float s_taps_buf[];
    if (staged) {                                          // small tap tables live in shared memory, big ones in L2
        for (int i = threadIdx.x; i < nw * n_taps; i += blockDim.x) s_taps_buf[i] = taps[i];
        __syncthreads();
    }
    const float* __restrict__ s_taps = staged ? s_taps_buf : taps;
    const int b = blockIdx.y;
    const int L = in_len ? min(in_len[b], n_in) : n_in;
    const int n_out = (int)(((int64_t)nw * L + orig - 1) / orig);              // ceil(new * L / orig)
    if (out_len && blockIdx.x == 0 && threadIdx.x == 0) out_len[b] = n_out;
    const float* __restrict__ row = in + (int64_t)b * in_stride;
    float* __restrict__ orow = out + (int64_t)b * out_stride;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out_cap; o += gridDim.x * blockDim.x) {
        float acc = 0.f;
        if (o < n_out) {
            const int m = o / nw, p = o - m * nw;
            const int base = m * orig - width;
            const float* __restrict__ k = s_taps + p * n_taps;
            for (int j = 0; j < n_taps; ++j) {
                const int i = base + j;
                if (i >= 0 && i < L) acc = fmaf(k[j], __ldg(row + i), acc);
            }
        }
        orow[o] = acc;
    }
}
}  // namespace sir

struct sir_resampler {
    int orig = 1, nw = 1, width = 0, n_taps = 0;
    DeviceBuffer taps;
};

extern "C" int sir_resampler_create(sir_resampler** out, int orig_freq, int new_freq, int lowpass_filter_width, double rolloff) {
    if (!out) return fail(SIR_ERR_INVALID, "sir_resampler_create: out is NULL");
    *out = nullptr;
    if (orig_freq < 1 || new_freq < 1 || lowpass_filter_width < 1 || !(rolloff > 0.0 && rolloff <= 1.0))
        return fail(SIR_ERR_INVALID, "sir_resampler_create: bad arguments");
    int a = orig_freq, b2 = new_freq;
    while (b2) {
        const int t = a % b2;
        a = b2;
        b2 = t;
    }
    const int orig = orig_freq / a, nw = new_freq / a;
    const double pi = 3.14159265358979323846;
    const double base_freq = (orig < nw ? orig : nw) * rolloff;
    const int width = (int)std::ceil(lowpass_filter_width * orig / base_freq);
    const int n_taps = 2 * width + orig;
    if ((size_t)nw * n_taps > ((size_t)1 << 26))
        return fail(SIR_ERR_UNSUPPORTED, "sir_resampler_create: %d x %d taps", nw, n_taps);
    std::vector<float> h((size_t)nw * n_taps);
    for (int p = 0; p < nw; ++p)
        for (int j = 0; j < n_taps; ++j) {
            double t = ((double)(-p) / nw + (double)(j - width) / orig) * base_freq;
            t = t < -lowpass_filter_width ? -lowpass_filter_width : (t > lowpass_filter_width ? lowpass_filter_width : t);
            const double c = std::cos(t * pi / lowpass_filter_width / 2);
            const double window = c * c;
            const double tp = t * pi;
            const double sinc = tp == 0.0 ? 1.0 : std::sin(tp) / tp;
            h[(size_t)p * n_taps + j] = (float)(sinc * window * (base_freq / orig));
        }
    sir_resampler* r = new sir_resampler();
    r->orig = orig;
    r->nw = nw;
    r->width = width;
    r->n_taps = n_taps;
    int rc = r->taps.reserve(h.size() * sizeof(float));
    if (rc == SIR_OK && cudaMemcpy(r->taps.ptr, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail(SIR_ERR_CUDA, "sir_resampler_create: tap upload failed");
    if (rc != SIR_OK) {
        r->taps.release();
        delete r;
        return rc;
    }
    if ((size_t)nw * n_taps * sizeof(float) > 48 * 1024 && (size_t)nw * n_taps * sizeof(float) <= 160 * 1024)
        cudaFuncSetAttribute(sir::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)nw * n_taps * sizeof(float)));
    *out = r;
    return SIR_OK;
}

extern "C" void sir_resampler_destroy(sir_resampler* r) {
    if (!r) return;
    r->taps.release();
    delete r;
}

extern "C" int64_t sir_resampler_output_length(const sir_resampler* r, int64_t n_in) {
    return r ? (n_in * r->nw + r->orig - 1) / r->orig : 0;
}

extern "C" int sir_resampler_forward(sir_resampler* r, const float* d_in, int64_t in_stride, const int32_t* d_in_lengths, int n_in,
                                     int batch, float* d_out, int64_t out_stride, int n_out, int32_t* d_out_lengths, void* stream) {
    if (!r || !d_in || !d_out || batch < 0 || n_in < 0 || n_out < 0 || in_stride < n_in || out_stride < n_out)
        return fail(SIR_ERR_INVALID, "sir_resampler_forward: bad arguments");
    if (batch == 0 || n_out == 0) return SIR_OK;
    if (batch > 65535) return fail(SIR_ERR_UNSUPPORTED, "sir_resampler_forward: batch > 65535");
    size_t smem = (size_t)r->nw * r->n_taps * sizeof(float);
    const bool staged = smem <= 160 * 1024;                 // small rational ratios (24 k -> 16 k: 2 x 23 taps) live in smem
    if (!staged) smem = 0;                                  // e.g. 22.05 k -> 16 k: 320 x 459 taps stay in global / L2
    dim3 grid((unsigned)((n_out + 255) / 256 < 148 ? (n_out + 255) / 256 : 148), (unsigned)batch);
    sir::resample_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(d_in, in_stride, d_in_lengths, n_in, (const float*)r->taps.ptr,
                                                                    r->n_taps, r->width, r->orig, r->nw, staged ? 1 : 0, d_out, out_stride,
                                                                    n_out, d_out_lengths);
    SIR_CHECK_LAUNCH("resample_kernel");
    return SIR_OK;
}
