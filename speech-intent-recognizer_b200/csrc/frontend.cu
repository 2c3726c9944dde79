// Fused log-mel frontend for sm_100a: waveform -> reflect-padded framing -> Hann -> 1024-point real FFT ->
// |X|^2 -> sparse HTK mel -> 10 log10 -> per-utterance (x-mean)/(std+1e-5) -> SpecAugment bands -> pad/trim.
//
// Replaces (per utterance) scripts/precompute_features.py:59-73, scripts/dataset.py:105-113,160-176 of the
// reference and the torchaudio calls behind them with ONE kernel launch for a batch.
//
// Mapping
//   * work item = (utterance, group of 8 frames); a persistent grid (3 CTAs per SM) draws items from a ticket
//     counter, so every SM is busy until the last item whatever the batch size;
//   * each HALF-WARP owns one frame and reads its 1024 samples straight from global memory (128 contiguous
//     bytes per half-warp request; the 50 % overlap with the neighbouring frame, held by the other half of
//     the same warp, is served by L1/L2, so HBM sees every sample once); only the first and last frame of an
//     utterance take the scalar path that resolves torch.stft's reflect padding;
//   * per frame: 512-point complex FFT as 32-point x 16-point register FFTs with one
//     shared-memory transposition (logmel_frame.cuh), real-FFT post-pass, power, sparse mel taps, log10;
//   * un-normalised values go to the output through an 8-frame shared tile (32-byte row segments) and the item's
//     partial statistics (shifted sums) to a small global array; the CTA whose item completes the utterance (atomic
//     count) merges the partials in group order (Chan's formula, fp64), re-reads the utterance's L2-resident values,
//     normalises, applies the mask bands and writes the zero padding.  HBM sees each sample once and each feature once.
//
// Roofline: HBM-bound by intent (4 L + 4 n_mels T bytes per utterance), but at ~19 k fp32 instructions per
// frame the kernel sits at the fp32-issue ridge; see DESIGN.md for the arithmetic.
#include <cmath>
#include <vector>

#include "frontend_params.cuh"
#include "frontend_tables.h"
#include "frontend_tc.h"
#include "logmel_frame.cuh"
#include "philox.cuh"
#include "sir_common.cuh"

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
constexpr int kMelWeightCap = 1536;              // taps incl. the zero padding of the 4-aligned runs

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
static_assert(kOffScratch % 4 == 0 && kOffTile % 4 == 0 && kOffTw512 % 4 == 0 && kOffMelWeight % 4 == 0 &&
                  kMelWeightCap == 3 * 4 * 128 && kMaxMels <= 128,
              "16-byte alignment of vector regions; table staging shape");
#ifndef SIR_FE_MIN_CTAS
#define SIR_FE_MIN_CTAS 3
#endif
static_assert(SIR_FE_MIN_CTAS * (kFeSmemBytes + 1024) <= 232448, "resident CTAs per SM");

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

// Persistent grid (SIR_FE_MIN_CTAS CTAs per SM) over work items: item i is the 8-frame group i % groups_max of
// utterance i / groups_max.  CTAs DRAW items from a global ticket counter (the next ticket is fetched while the
// current item computes), so the SMs finish within one item of each other whatever the batch size and whatever an
// item costs (a cluster per utterance left 42 % of the machine idle at 256 utterances: 512 CTAs of 6 groups on 444
// slots; a static round-robin let the CTAs that lag - reflect-padded first/last groups, finisher work - collect
// every later finisher job too).  The CTA whose item completes an utterance - an atomic counter per utterance,
// reset by that CTA for the next launch - merges the item partials in GROUP order (deterministic whichever CTA it
// is), re-reads the utterance's still-L2-resident values, normalises, masks and pads them.
// Tickets: the counter is never reset; every CTA draws (its items + 1) tickets, so a launch consumes exactly
// total_items + gridDim.x of them and the host passes the first ticket of the launch (work_base).
template <typename SampleT>
__global__ void __launch_bounds__(kFeThreads, SIR_FE_MIN_CTAS) logmel_frontend_kernel(const FrontendParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;

    // ---- constants into shared memory (once per CTA): all loads of a thread are issued before its first store -----
    int* s_mel_start = reinterpret_cast<int*>(smem + kOffMelStart);
    int* s_mel_count = reinterpret_cast<int*>(smem + kOffMelCount);
    int* s_mel_offset = reinterpret_cast<int*>(smem + kOffMelOffset);
    {
        static_assert(kFeThreads == 128, "table staging below assumes 128 threads");
        const float4* gw = reinterpret_cast<const float4*>(p.tables.window);      // 256 float4 each (16-byte aligned
        const float4* gt = reinterpret_cast<const float4*>(p.tables.tw512);       //  sections, sir_frontend_create)
        const float4* gm = reinterpret_cast<const float4*>(p.tables.mel_weight);
        const int mel4 = (p.mel_weight_count + 3) >> 2;                           // <= kMelWeightCap / 4 = 3 * 128
        const float4 w0 = __ldg(gw + tid), w1 = __ldg(gw + tid + 128), t0 = __ldg(gt + tid), t1 = __ldg(gt + tid + 128);
        float4 m4[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            m4[k] = tid + 128 * k < mel4 ? __ldg(gm + tid + 128 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
        float tw[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) tw[k] = tid + 128 * k < 514 ? __ldg(p.tables.tw1024 + tid + 128 * k) : 0.f;
        int ms = 0, mc = 0, mo = 0;
        if (tid < p.n_mels) {
            ms = __ldg(p.tables.mel_start + tid);
            mc = __ldg(p.tables.mel_count + tid);
            mo = __ldg(p.tables.mel_offset + tid);
        }
        reinterpret_cast<float4*>(smem + kOffWindow)[tid] = w0;
        reinterpret_cast<float4*>(smem + kOffWindow)[tid + 128] = w1;
        reinterpret_cast<float4*>(smem + kOffTw512)[tid] = t0;
        reinterpret_cast<float4*>(smem + kOffTw512)[tid + 128] = t1;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (tid + 128 * k < mel4) reinterpret_cast<float4*>(smem + kOffMelWeight)[tid + 128 * k] = m4[k];
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (tid + 128 * k < 514) smem[kOffTw1024 + tid + 128 * k] = tw[k];
        if (tid < p.n_mels) {
            s_mel_start[tid] = ms;
            s_mel_count[tid] = mc;
            s_mel_offset[tid] = mo;
        }
    }
    const FrontendTables st{smem + kOffWindow, smem + kOffTw512, smem + kOffTw1024, s_mel_start,
                            s_mel_count,       s_mel_offset,     smem + kOffMelWeight};

    float* tile = smem + kOffTile;
    float* red = smem + kOffReduce;
    int* s_flag = reinterpret_cast<int*>(red + 32);
    const int lane = tid & 31, warp = tid >> 5, half = lane >> 4, q = lane & 15;
    const int slot = 2 * warp + half;
    float* scr = smem + kOffScratch + slot * kFrameScratch;
    const bool mfcc = p.mode == SIR_OUT_MFCC;
    const bool needs_finish = mfcc || p.mode == SIR_OUT_LOGMEL_NORM;
    const int out_rows = mfcc ? p.n_mfcc : p.n_mels;
    const int row_stride = mfcc ? p.stage_frames : p.out_frames;
    const int64_t total_items = (int64_t)p.batch * p.groups_max;
    long long* s_item = reinterpret_cast<long long*>(red + 40);

    auto process_item = [&](const int64_t item) {
        const int b = (int)(item / p.groups_max), g = (int)(item - (int64_t)b * p.groups_max);
        int L = p.lengths ? min(p.lengths[b], p.n_samples) : p.n_samples;
        if (p.max_samples > 0) L = min(L, p.max_samples);
        const bool valid = L > kNfft / 2;                   // reflect padding needs L > 512
        const int T = valid ? 1 + L / kHop : 0;
        const int n_groups = valid ? (T + kGroupFrames - 1) / kGroupFrames : 1;   // an invalid utterance: one zero-fill item
        if (g >= n_groups) return;                          // uniform over the CTA
        float* __restrict__ final_out = p.out + (int64_t)b * out_rows * p.out_frames;
        if (g == 0 && tid == 0 && p.status) p.status[b] = valid ? 0 : 1;
        if (!valid) {
            for (int i = tid; i < out_rows * p.out_frames; i += kFeThreads) final_out[i] = 0.f;
            return;
        }
        const SampleT* __restrict__ row = static_cast<const SampleT*>(p.wave) + (int64_t)b * p.wave_stride;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & GlobalFrame<SampleT>::kAlignMask) == 0);
        // where the item writes its (un-normalised) values: the output itself, or the dB staging buffer for MFCC
        float* __restrict__ out = mfcc ? p.db_stage + (int64_t)b * p.n_mels * p.stage_frames : final_out;
        const bool out_vec = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);

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
                    // 10 log10(x) = (10 log10 2) lg2(x): lg2.approx is good to ~1e-7 relative here, i.e. ~1e-6 dB
                    if (p.mode != SIR_OUT_MEL_POWER) v = 3.01029995663981195f * __log2f(fmaxf(v, 1e-10f));
                    tile[m * kTileStride + slot] = v;
                }
            }
        }
        __syncthreads();                                    // tile complete

        // tile -> global (8 consecutive frames of one mel row = one 32-byte segment) + the item's statistics
        const float shift = tile[0];                        // fp32 sums of deviations from a nearby value stay short
        float s1 = 0.f, s2 = 0.f, vmax = -INFINITY;
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
        if (!needs_finish) {
            if (g == 0)                                     // zero padding behind the last frame, all rows
                for (int m = 0; m < p.n_mels; ++m)
                    for (int tt = T + tid; tt < p.out_frames; tt += kFeThreads) out[(int64_t)m * p.out_frames + tt] = 0.f;
            return;                                         // (the item loop ends every item with a barrier)
        }

        // ---- the item's partial (n, mean, M2, max) -> global; count the item; the last one finishes the utterance ----
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (lane == 0) {
            red[warp] = s1;
            red[8 + warp] = s2;
            red[16 + warp] = vmax;
        }
        __syncthreads();                                    // every thread's feature stores happen before thread 0's fence
        if (tid == 0) {
            ItemPartial part;
            part.s1 = ((double)red[0] + (double)red[1]) + ((double)red[2] + (double)red[3]);
            part.s2 = ((double)red[8] + (double)red[9]) + ((double)red[10] + (double)red[11]);
            part.shift = shift;
            part.vmax = fmaxf(fmaxf(red[16], red[17]), fmaxf(red[18], red[19]));
            part.n = nslots * p.n_mels;
            part.pad = 0;
            p.partials[item] = part;
            __threadfence();                                // release: the CTA's stores (bar.sync above) and the partial
            const bool last = atomicAdd(p.counters + b, 1) == n_groups - 1;
            if (last) __threadfence();                      // acquire: the other items' stores; also drops this SM's L1 lines
            *s_flag = last;
        }
        __syncthreads();
        if (!*s_flag) return;                               // uniform over the CTA

        // ---- finisher: every item of utterance b is in global memory ------------------------------------------
        if (warp == 0) {
            if (lane == 0) p.counters[b] = 0;               // ready for the next launch
            const ItemPartial* parts = p.partials + (int64_t)b * p.groups_max;
            double n = 0, mean = 0, m2 = 0;
            float mx = -INFINITY;
            for (int base = 0; base < n_groups; base += 32) {
                // lane i turns item base + i into (n, mean, M2); then all lanes merge the items in GROUP order
                // (deterministic whichever CTA finishes) with one fp64 division per item
                double ni = 0, mi = 0, m2i = 0;
                if (base + lane < n_groups) {
                    const double2 a = __ldcg(reinterpret_cast<const double2*>(parts + base + lane));       // (s1, s2)
                    const float2 c = __ldcg(reinterpret_cast<const float2*>(parts + base + lane) + 2);     // (shift, vmax)
                    const int cnt = __ldcg(reinterpret_cast<const int*>(parts + base + lane) + 6);
                    ni = (double)cnt;
                    const double r = a.x / ni;
                    mi = (double)c.x + r;
                    m2i = fmax(a.y - a.x * r, 0.0);
                    mx = fmaxf(mx, c.y);
                }
                const int cnt_items = min(32, n_groups - base);
                for (int gg = 0; gg < cnt_items; ++gg) {
                    const double nb = __shfl_sync(0xffffffffu, ni, gg), mb = __shfl_sync(0xffffffffu, mi, gg),
                                 m2b = __shfl_sync(0xffffffffu, m2i, gg);
                    const double nt = n + nb, delta = mb - mean, qd = delta * nb / nt;
                    mean += qd;
                    m2 += m2b + delta * qd * n;
                    n = nt;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) {
                red[33] = (float)mean;
                red[34] = (float)(1.0 / (sqrt(m2 / (n - 1.0)) + 1e-5));
                red[35] = mx;
            }
        }
        if (mfcc) {
            float* s_dct = smem + kOffScratch;              // the frame scratch is free: [n_mels][n_mfcc]
            for (int i = tid; i < p.n_mels * p.n_mfcc; i += kFeThreads) s_dct[i] = p.dct[i];
            __syncthreads();
            const float floor_db = p.top_db > 0.f ? red[35] - p.top_db : -INFINITY;
            const int Tn = min(T, p.out_frames);
            for (int idx = tid; idx < p.n_mfcc * Tn; idx += kFeThreads) {
                const int c = idx / Tn, tt = idx - c * Tn;
                float acc = 0.f;
                for (int m = 0; m < p.n_mels; ++m)
                    acc = fmaf(fmaxf(__ldcg(out + (int64_t)m * row_stride + tt), floor_db), s_dct[m * p.n_mfcc + c], acc);
                final_out[(int64_t)c * p.out_frames + tt] = acc;
            }
            for (int c = 0; c < p.n_mfcc; ++c)
                for (int tt = T + tid; tt < p.out_frames; tt += kFeThreads) final_out[(int64_t)c * p.out_frames + tt] = 0.f;
            return;                                         // (the item loop's barrier protects s_dct / the scratch)
        }
        __syncthreads();
        const float fmean = red[33], inv = red[34];
        int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
        if (p.masks) {
            mt0 = p.masks[4 * b + 0];
            mt1 = p.masks[4 * b + 1];
            mf0 = p.masks[4 * b + 2];
            mf1 = p.masks[4 * b + 3];
        }
        // normalise + mask + pad the whole utterance.  Its values are still L2-resident; plain loads are safe: thread 0's
        // fence after the counting atomic is the acquire and dropped this SM's L1 lines.  Loads of a batch are issued
        // before its first store (a load -> store chain per element would pay one L2 round trip per iteration: the
        // stores alias the loads as far as the compiler knows); (row, column) advance by kFeThreads elements per step
        // without a division per element.
        constexpr int kU = 8;
        const int cols = out_vec ? p.out_frames >> 2 : p.out_frames;          // float4 chunks or floats per row
        const int dq = kFeThreads / cols, dr = kFeThreads % cols;
        int m_ld = tid / cols, c_ld = tid % cols;
        while (m_ld < p.n_mels) {
            int mm[kU], cc[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                mm[u] = m_ld;
                cc[u] = c_ld;
                m_ld += dq;
                c_ld += dr;
                if (c_ld >= cols) {
                    c_ld -= cols;
                    ++m_ld;
                }
            }
            if (out_vec) {
                float4 x[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (mm[u] < p.n_mels && 4 * cc[u] < T)
                        x[u] = *reinterpret_cast<const float4*>(out + (int64_t)mm[u] * p.out_frames + 4 * cc[u]);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    if (mm[u] < p.n_mels) {
                        const int m = mm[u], tt = 4 * cc[u];
                        const bool fm = m >= mf0 && m < mf1;
                        float4 v;
                        v.x = (tt >= T || fm || (tt >= mt0 && tt < mt1)) ? 0.f : (x[u].x - fmean) * inv;
                        v.y = (tt + 1 >= T || fm || (tt + 1 >= mt0 && tt + 1 < mt1)) ? 0.f : (x[u].y - fmean) * inv;
                        v.z = (tt + 2 >= T || fm || (tt + 2 >= mt0 && tt + 2 < mt1)) ? 0.f : (x[u].z - fmean) * inv;
                        v.w = (tt + 3 >= T || fm || (tt + 3 >= mt0 && tt + 3 < mt1)) ? 0.f : (x[u].w - fmean) * inv;
                        *reinterpret_cast<float4*>(out + (int64_t)m * p.out_frames + tt) = v;
                    }
                }
            } else {
                float x[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    x[u] = 0.f;
                    if (mm[u] < p.n_mels && cc[u] < T) x[u] = out[(int64_t)mm[u] * p.out_frames + cc[u]];
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    if (mm[u] < p.n_mels) {
                        const int m = mm[u], tt = cc[u];
                        const bool masked = (tt >= mt0 && tt < mt1) || (m >= mf0 && m < mf1);
                        out[(int64_t)m * p.out_frames + tt] = (tt >= T || masked) ? 0.f : (x[u] - fmean) * inv;
                    }
                }
            }
        }
    };

    // ---- item loop: draw a ticket, fetch the next one while this item computes ---------------------------------
    long long fetched = 0;
    if (tid == 0) fetched = (long long)(atomicAdd(p.work_counter, 1ULL) - p.work_base);
    for (;;) {
        if (tid == 0) *s_item = fetched;
        __syncthreads();                                    // also: tables visible (first pass), previous item fully done
        const long long item = *s_item;
        if (item < 0 || item >= total_items) break;         // (negative: the host's ticket base is ahead of the counter)
        if (tid == 0) fetched = (long long)(atomicAdd(p.work_counter, 1ULL) - p.work_base);
        process_item(item);
        __syncthreads();                                    // tile / red[] / scratch / s_item consumed
    }
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

// Per-stream scratch of a frontend handle (launches on different streams may overlap on the GPU).
struct FrontendWorkspace {
    void* stream = nullptr;
    DeviceBuffer partials;       // [batch][groups_max] ItemPartial
    DeviceBuffer counters;       // [batch] finished items per utterance; the finishing CTA resets its entry
    DeviceBuffer db_stage;       // MFCC: staged dB values
    DeviceBuffer tickets;        // one 64-bit ticket counter
    unsigned long long next_ticket = 0;
};

struct sir_frontend {
    int device = 0;
    int sample_rate = 16000, n_mels = 64;
    int num_sms = 148;
    int mel_weight_count = 0;
    DeviceBuffer tables;
    FrontendTables dev{};
    static constexpr int kMaxStreams = 32;
    FrontendWorkspace work[kMaxStreams];
    int work_used = 0;
    DeviceBuffer dct;            // MFCC: [n_mels][n_mfcc] ortho DCT-II
    int dct_n_mfcc = 0;
    DeviceBuffer tc_tables;      // tensor-core DFT kernel: operand images, twiddles, unscaled mel taps
    TcDeviceTables tc_dev{};
};

// Which kernel serves sir_frontend_forward*: the tensor-core DFT (default) or the CUDA-core FFT of round 1
// (SIR_FRONTEND_KERNEL=cuda in the environment; kept for A/B measurements - both are CUDA, there is no CPU path).
static bool use_tc_kernel() {
    static const bool tc = [] {
        const char* v = getenv("SIR_FRONTEND_KERNEL");
        return !(v && (v[0] == 'c' || v[0] == 'C'));
    }();
    return tc;
}

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
    rc = frontend_tc_upload_tables(fe->tc_tables, fe->tc_dev, sample_rate, n_mels);
    if (rc != SIR_OK) {
        fe->tables.release();
        fe->tc_tables.release();
        delete fe;
        return rc;
    }
    *out = fe;
    return SIR_OK;
}

extern "C" void sir_frontend_destroy(sir_frontend* fe) {
    if (!fe) return;
    fe->tables.release();
    fe->tc_tables.release();
    for (auto& w : fe->work) {
        w.partials.release();
        w.counters.release();
        w.db_stage.release();
        w.tickets.release();
    }
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
    int slot = 0;
    while (slot < fe->work_used && fe->work[slot].stream != stream) ++slot;
    if (slot == fe->work_used) {
        if (slot == sir_frontend::kMaxStreams)
            return fail(SIR_ERR_UNSUPPORTED, "sir_frontend_forward: one handle serves at most %d streams; create another handle",
                        sir_frontend::kMaxStreams);
        fe->work[slot].stream = stream;
        ++fe->work_used;
    }
    FrontendWorkspace& ws = fe->work[slot];
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
    const bool tc_kernel = use_tc_kernel();
    p.tc = fe->tc_dev;
    if (mode == SIR_OUT_MFCC) {
        if (n_mfcc < 1 || n_mfcc > fe->n_mels || fe->n_mels * n_mfcc > 7936)
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
        int rc = ws.db_stage.reserve((size_t)batch * fe->n_mels * p.stage_frames * sizeof(float));
        if (rc != SIR_OK) return rc;
        p.db_stage = (float*)ws.db_stage.ptr;
        p.dct = (const float*)fe->dct.ptr;
        p.n_mfcc = n_mfcc;
        p.top_db = top_db;
    }
    // work items: (utterance, 8-frame group); a persistent grid of SIR_FE_MIN_CTAS CTAs per SM takes them round-robin
    int eff = max_samples > 0 && max_samples < n_samples ? max_samples : n_samples;
    const int groups = eff > kNfft / 2 ? (tc_kernel ? frontend_tc_groups(1 + eff / kHop) : (1 + eff / kHop + kGroupFrames - 1) / kGroupFrames) : 1;
    const int64_t items = (int64_t)batch * groups;
    if (items >= ((int64_t)1 << 31)) return fail(SIR_ERR_UNSUPPORTED, "sir_frontend_forward: %lld work items in one launch", (long long)items);
    p.batch = batch;
    p.groups_max = groups;
    if (mode == SIR_OUT_LOGMEL_NORM || mode == SIR_OUT_MFCC) {
        int rc = ws.partials.reserve((size_t)items * sizeof(ItemPartial));
        if (rc != SIR_OK) return rc;
        if ((size_t)batch * sizeof(int) > ws.counters.bytes) {
            if ((rc = ws.counters.reserve((size_t)batch * sizeof(int))) != SIR_OK) return rc;
            SIR_CUDA(cudaMemsetAsync(ws.counters.ptr, 0, ws.counters.bytes, (cudaStream_t)stream));
        }
        p.partials = (ItemPartial*)ws.partials.ptr;
        p.counters = (int*)ws.counters.ptr;
    }
    const int64_t slots = tc_kernel ? (int64_t)fe->num_sms : (int64_t)fe->num_sms * SIR_FE_MIN_CTAS;
    const int64_t grid = items < slots ? items : slots;
    if (!ws.tickets.ptr) {
        int rc = ws.tickets.reserve(sizeof(unsigned long long));
        if (rc != SIR_OK) return rc;
        SIR_CUDA(cudaMemsetAsync(ws.tickets.ptr, 0, sizeof(unsigned long long), (cudaStream_t)stream));
        ws.next_ticket = 0;
    }
    p.work_counter = (unsigned long long*)ws.tickets.ptr;
    p.work_base = ws.next_ticket;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kFeThreads);
    cfg.dynamicSmemBytes = kFeSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    {
        ProfScope ps("logmel_frontend_kernel", cfg.stream);
        if (tc_kernel) {
            const int rc = frontend_tc_launch(p, pcm16, fe->num_sms, cfg.stream);
            if (rc != SIR_OK) return rc;
        } else if (pcm16)
            SIR_CUDA(cudaLaunchKernelEx(&cfg, logmel_frontend_kernel<short>, p));
        else
            SIR_CUDA(cudaLaunchKernelEx(&cfg, logmel_frontend_kernel<float>, p));
    }
    SIR_CHECK_LAUNCH("logmel_frontend_kernel");
    // tickets a launch consumes (a failed launch: none): items + 1 per CTA, items + 2 per CTA for the tensor-core kernel
    ws.next_ticket += tc_kernel ? (unsigned long long)frontend_tc_tickets(items, grid) : (unsigned long long)(items + grid);
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
