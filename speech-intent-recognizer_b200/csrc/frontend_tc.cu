// Fused log-mel frontend on the sm_100a TENSOR CORES: the 1024-point real DFT of every frame is two matrix stages on
// tcgen05 (32 x 32 Cooley-Tukey, fp16 (hi, lo) split operands, fp32 accumulators in tensor memory); CUDA cores only
// window / scale / split the samples, apply the inter-stage twiddles, square, and run the sparse mel projection.
//
// Replaces (per utterance) scripts/precompute_features.py:59-73, scripts/dataset.py:105-113,160-176 of the reference and
// the torchaudio calls behind them, like frontend.cu, whose CUDA-core FFT it supersedes: that kernel
// needs ~1,360 warp instructions per frame and is issue / latency bound at 8-11 % of the HBM roofline; here the butterflies
// are MMAs and ~550 warp instructions per frame remain.  Numerics: frontend_tc_tables.h, tests/host/tc_dft_host_check.cpp
// (within ~2x of an fp32 FFT's own rounding error).
//
// One persistent CTA per SM, 13 warps, work item = 15 consecutive frames of one utterance (15 x 17 stage-2 rows = 255 = two
// 128-row UMMA tiles), drawn from a ticket counter:
//   warps 0-3   "A": load samples (lane = n2, one coalesced 128-byte request per 32 samples; a 512-sample block is loaded
//                once and serves the two frames that overlap it), per-frame power-of-two scale to max|x| in [1, 2), Hann
//                window, fp16 (hi, lo) split, write the stage-1 operand rows (frame, n2) x K = n1 (hi | lo in one 128-byte
//                swizzled row).  Warp w fills sub-tile w = frames 4w..4w+3.
//   warp 4      issues the MMAs: stage 1 per sub-tile: D1 = A_hi [B1_hi; B1_lo] (N = 64) + A_lo B1_hi (N = 32);
//                stage 2 per 128-row tile: D2 = A2_hi [B2_hi; B2_lo] (N = 128) + A2_lo B2_hi (N = 64).  Stage 2 of item i is
//                issued after stage 1 of item i+1, so the two stages of consecutive items overlap.
//   warps 5-8   "C": read D1 (lane = n2, 32 real numbers = Y[0..16]), multiply by the twiddles W1024^(n2 k1), split, and
//                write the stage-2 operand rows (frame, k1) x K = (n2, re/im): for a fixed k1 the 32 lanes write 32
//                consecutive words of one row - the transposition between the stages costs no bank conflict.
//   warps 9-12  "D/E": read D2 (lane = (frame, k1), 32 complex bins k1 + 32 k2), |X|^2 into the one-sided power spectrum
//                (the bins with k mod 32 > 16 are mirrors), sparse mel taps, dB, tile -> global, the item's partial
//                statistics, and - for the item that completes an utterance - the normalisation pass (as in frontend.cu).
// Every hand-off is an mbarrier; every wait is bounded and traps.
#include <cuda_fp16.h>

#include <vector>

#include "frontend_params.cuh"
#include "frontend_tables.h"
#include "frontend_tc.h"
#include "frontend_tc_tables.h"
#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {
namespace fetc {

using namespace tc;

constexpr int kNumWarps = 13;
constexpr int kThreads = kNumWarps * 32;
constexpr int kWarpMma = 4, kWarpD0 = 9;       // warps 0-3: A, 4: MMA, 5-8: C, 9-12: D/E
constexpr int kRing = 8;                       // item slots between the ticket drawer and the other warps
constexpr int kMelWeightCap = 1536;

struct ItemSlot {
    long long item;                            // < 0: no more work
    int b, t0, nfr, T, L, n_groups;
    float inv2[16];                            // per frame: 1 / scale^2 (undoes the power-of-two scaling on the power)
};
struct Control {
    uint64_t ring_full[kRing], ring_empty[kRing], sc_full[kRing];
    uint64_t a1_full[4], a1_empty[4], d1_full[4], d1_empty[4];
    uint64_t a2_full[2], a2_empty, d2_full[2], d2_empty[2];
    ItemSlot slot[kRing];
    uint32_t tmem_base;
    int unit_counter;
    int flag;
    int pad;
    float red[48];
};

// shared-memory carve-up (bytes from the 1024-aligned base)
constexpr uint32_t kOffA1 = 0;                                   // 4 sub-tiles x 128 rows x 128 B (hi | lo halves of K)
constexpr uint32_t kOffA2Hi = 65536, kOffA2Lo = 98304;           // 256 rows x 128 B each
constexpr uint32_t kOffB1 = 131072;                              // 64 rows x 128 B
constexpr uint32_t kOffB2 = 139264;                              // 128 rows x 128 B
constexpr uint32_t kOffP = 155648;                               // 15 x 528 floats
constexpr uint32_t kPBytes = 31744;
constexpr uint32_t kOffMelW = kOffP + kPBytes;                   // 1536 floats
constexpr uint32_t kOffMelIdx = kOffMelW + kMelWeightCap * 4;    // start / count / offset: 3 x 128 ints
constexpr uint32_t kOffTile = kOffMelIdx + 3 * kMaxMels * 4;     // [n_mels][16] floats
constexpr uint32_t kOffWin = kOffTile + kMaxMels * 16 * 4;       // Hann window as [8][32 lanes][4]: lane's w[32 n1 + lane], n1 = 4c..4c+3
constexpr uint32_t kOffCtl = kOffWin + 4096;
constexpr uint32_t kSmemBytes = kOffCtl + ((sizeof(Control) + 127) & ~127u) + 1024;   // + slack for the 1024-byte alignment
static_assert(kTileFrames * kPStride * 4 <= kPBytes, "power buffer");
static_assert(kSmemBytes <= 232448, "shared memory per CTA");
static_assert(kOffA2Hi % 1024 == 0 && kOffA2Lo % 1024 == 0 && kOffB1 % 1024 == 0 && kOffB2 % 1024 == 0, "swizzle atoms");

__device__ __forceinline__ void group_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four D/E warps
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (a, b) -> fp16 pair of the values and fp16 pair of what the rounding dropped
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(a - f.x, b - f.y));
}

// One 512-sample block (samples [512 jb, 512 jb + 512) of the reflect-padded utterance, jb >= -1): lane takes the samples
// 32 i + lane, so every load of the warp is one contiguous 128-byte (fp32) / 64-byte (PCM16) request.  Loads only: the
// caller issues the next block's loads before it works on the current frame (the loads then fly during ~200 instructions).
template <typename SampleT>
__device__ __forceinline__ void load_block(const SampleT* __restrict__ row, int L, int jb, int lane, float (&x)[16]) {
    const int n0 = jb * kHop + lane;
    if (jb >= 0 && (jb + 1) * kHop <= L) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = sample_to_float(__ldg(row + n0 + 32 * i));
    } else {                                                       // torch.stft's reflect padding at both ends
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int n = n0 + 32 * i;
            n = n < 0 ? -n : n;
            n = n >= L ? 2 * (L - 1) - n : n;
            n = n < 0 ? 0 : n;
            x[i] = sample_to_float(__ldg(row + n));
        }
    }
}

// One frame = blocks (first: n1 0..15, second: n1 16..31): Hann window, power-of-two scale that puts the frame's largest
// WINDOWED value in [1, 2) (the window can take a loud frame edge down by 100 dB: scaling by the raw maximum would leave the
// operand in the fp16 subnormals there), fp16 (hi, lo) split, store row `r` of the stage-1 operand sub-tile.
// Returns 1 / scale^2.
__device__ __forceinline__ float store_frame(const float (&first)[16], const float (&second)[16], const float4* __restrict__ s_win,
                                             uint32_t tile_addr, int r) {
    float t[32];
    float m = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 w = s_win[c * 32];                          // (this lane's) w[32 n1 + lane], n1 = 4c .. 4c+3
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n1 = 4 * c + e;
            t[n1] = (n1 < 16 ? first[n1] : second[n1 - 16]) * wv[e];
            m = fmaxf(m, fabsf(t[n1]));
        }
    }
    uint32_t eb = __reduce_max_sync(0xffffffffu, __float_as_uint(m)) >> 23;
    eb = eb < 65u ? 65u : (eb > 187u ? 187u : eb);
    const float scale = __uint_as_float((254u - eb) << 23), inv = __uint_as_float(eb << 23);
    const uint32_t row_addr = tile_addr + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const int sw = r & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n1 = 8 * c + 2 * e;
            split_pair(t[n1] * scale, t[n1 + 1] * scale, hi[e], lo[e]);         // the scaling is exact
        }
        st_shared_v4(row_addr + (uint32_t)((c ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);       // K = n1        (hi half)
        st_shared_v4(row_addr + (uint32_t)(((4 + c) ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]); // K = 32 + n1   (lo half)
    }
    return inv * inv;
}

template <typename SampleT>
__global__ void __launch_bounds__(kThreads, 1) logmel_frontend_tc_kernel(const FrontendParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Control* ctl = reinterpret_cast<Control*>(smem + kOffCtl);
    float* s_P = reinterpret_cast<float*>(smem + kOffP);
    float* s_melw = reinterpret_cast<float*>(smem + kOffMelW);
    int* s_mel_start = reinterpret_cast<int*>(smem + kOffMelIdx);
    int* s_mel_count = s_mel_start + kMaxMels;
    int* s_mel_offset = s_mel_count + kMaxMels;
    float* s_tile = reinterpret_cast<float*>(smem + kOffTile);
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const uint32_t sbase = smem_u32(smem);

    // ---- one-time setup -------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) {
            mbar_init(&ctl->ring_full[i], 1);
            mbar_init(&ctl->ring_empty[i], 12);                  // A warps 1-3, MMA warp, 4 C warps, 4 D/E warps
            mbar_init(&ctl->sc_full[i], 4);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&ctl->a1_full[i], 1);
            mbar_init(&ctl->a1_empty[i], 1);
            mbar_init(&ctl->d1_full[i], 1);
            mbar_init(&ctl->d1_empty[i], 4);
        }
        mbar_init(&ctl->a2_full[0], 8);                          // frames 0..7: two sub-tiles x four C warps
        mbar_init(&ctl->a2_full[1], 16);                         // every frame has its k1 = 16 row in the second tile
        mbar_init(&ctl->a2_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ctl->d2_full[i], 1);
            mbar_init(&ctl->d2_empty[i], 4);
        }
        ctl->unit_counter = 0;
        fence_barrier_init();
    }
    if (warp == kWarpMma) tmem_alloc<512>(&ctl->tmem_base);
    {
        const uint4* g1 = reinterpret_cast<const uint4*>(p.tc.b1_img);
        const uint4* g2 = reinterpret_cast<const uint4*>(p.tc.b2_img);
        uint4* s1 = reinterpret_cast<uint4*>(smem + kOffB1);
        uint4* s2 = reinterpret_cast<uint4*>(smem + kOffB2);
        for (int i = tid; i < 512; i += kThreads) s1[i] = __ldg(g1 + i);
        for (int i = tid; i < 1024; i += kThreads) s2[i] = __ldg(g2 + i);
        for (int i = tid; i < kMelWeightCap; i += kThreads) s_melw[i] = i < p.mel_weight_count ? __ldg(p.tc.mel_weight + i) : 0.f;
        for (int i = tid; i < p.n_mels; i += kThreads) {
            s_mel_start[i] = __ldg(p.tables.mel_start + i);
            s_mel_count[i] = __ldg(p.tables.mel_count + i);
            s_mel_offset[i] = __ldg(p.tables.mel_offset + i);
        }
        for (int i = tid; i < 1024; i += kThreads) {               // window[32 n1 + l] -> [n1 / 4][l][n1 % 4]
            const int n1 = i >> 5, l = i & 31;
            reinterpret_cast<float*>(smem + kOffWin)[((n1 >> 2) * 32 + l) * 4 + (n1 & 3)] = __ldg(p.tables.window + i);
        }
        for (int i = tid; i < (int)(kPBytes / 4); i += kThreads) s_P[i] = 0.f;       // bins 513..527 stay zero for good
        fence_proxy_async();                                     // the operand images are read by the tensor core
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    const int64_t total_items = (int64_t)p.batch * p.groups_max;

    if (warp < 4) {
        // =============================== A: samples -> stage-1 operand ===============================================
        const float4* s_win = reinterpret_cast<const float4*>(smem + kOffWin) + lane;
        // ticket pipeline of the drawing lane: i0 / L0 = this iteration's item and its length, i1 = the next item
        long long i0 = -1, i1 = -1;
        int L0 = 0;
        auto length_of = [&](long long item) -> int {
            if (item < 0 || item >= total_items) return 0;
            const int b = (int)(item / p.groups_max);
            int L = p.lengths ? min(__ldg(p.lengths + b), p.n_samples) : p.n_samples;
            if (p.max_samples > 0) L = min(L, p.max_samples);
            return L;
        };
        auto draw = [&]() -> long long { return (long long)(atomicAdd(p.work_counter, 1ULL) - p.work_base); };
        if (warp == 0 && lane == 0) {
            i0 = draw();
            i1 = draw();
            L0 = length_of(i0);
        }
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            const uint32_t rph = (it / kRing) & 1u;
            if (warp == 0) {
                if (lane == 0) {
                    mbar_wait(&ctl->ring_empty[rs], rph ^ 1u);
                    ItemSlot& sl = ctl->slot[rs];
                    for (;;) {                                   // skip tickets beyond an utterance's last group (ragged batches)
                        if (i0 < 0 || i0 >= total_items) {
                            sl.item = -1;
                            break;
                        }
                        const int b = (int)(i0 / p.groups_max), g = (int)(i0 - (long long)b * p.groups_max);
                        const bool valid = L0 > kNfft / 2;
                        const int T = valid ? 1 + L0 / kHop : 0;
                        const int n_groups = valid ? (T + kTileFrames - 1) / kTileFrames : 1;
                        if (g < n_groups) {
                            sl.item = i0;
                            sl.b = b;
                            sl.t0 = g * kTileFrames;
                            sl.nfr = valid ? min(kTileFrames, T - g * kTileFrames) : 0;
                            sl.T = T;
                            sl.L = L0;
                            sl.n_groups = n_groups;
                            break;
                        }
                        const int L1 = length_of(i1);
                        i0 = i1;
                        L0 = L1;
                        i1 = draw();
                    }
                    mbar_arrive(&ctl->ring_full[rs]);
                    if (sl.item >= 0) {                          // advance: the next length load and the next draw are in
                        const int L1 = length_of(i1);            // flight while this item is processed
                        i0 = i1;
                        L0 = L1;
                        i1 = draw();
                    }
                }
                __syncwarp();
            }
            mbar_wait(&ctl->ring_full[rs], rph);
            ItemSlot& sl = ctl->slot[rs];
            if (sl.item < 0) break;
            const int L = sl.L, t0 = sl.t0, nfr = sl.nfr;
            const SampleT* __restrict__ row = static_cast<const SampleT*>(p.wave) + (int64_t)sl.b * p.wave_stride;
            mbar_wait(&ctl->a1_empty[warp], (it & 1u) ^ 1u);     // stage 1 of the previous item has read this sub-tile
            const int f0 = 4 * warp;
            if (f0 < nfr) {
                const uint32_t tile_addr = sbase + kOffA1 + (uint32_t)warp * 16384u;
                // blocks t0 + f0 - 1 .. t0 + f0 + 3 in three rotating register buffers: frame j = (block j, block j + 1), and
                // block j + 2 is requested before frame j is processed
                float x0[16], x1[16], x2[16];
                const int jb = t0 + f0 - 1, nf = min(4, nfr - f0);
                load_block(row, L, jb, lane, x0);
                load_block(row, L, jb + 1, lane, x1);
                if (nf > 1) load_block(row, L, jb + 2, lane, x2);
                float i2 = store_frame(x0, x1, s_win, tile_addr, lane);
                if (lane == 0) sl.inv2[f0] = i2;
                if (nf > 1) {
                    if (nf > 2) load_block(row, L, jb + 3, lane, x0);
                    i2 = store_frame(x1, x2, s_win, tile_addr, 32 + lane);
                    if (lane == 0) sl.inv2[f0 + 1] = i2;
                }
                if (nf > 2) {
                    if (nf > 3) load_block(row, L, jb + 4, lane, x1);
                    i2 = store_frame(x2, x0, s_win, tile_addr, 64 + lane);
                    if (lane == 0) sl.inv2[f0 + 2] = i2;
                }
                if (nf > 3) {
                    i2 = store_frame(x0, x1, s_win, tile_addr, 96 + lane);
                    if (lane == 0) sl.inv2[f0 + 3] = i2;
                }
                fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ctl->a1_full[warp]);
                mbar_arrive(&ctl->sc_full[rs]);
                if (warp != 0) mbar_arrive(&ctl->ring_empty[rs]);
            }
        }
    } else if (warp == kWarpMma) {
        // =============================== MMA issue ===================================================================
        constexpr uint32_t id1a = make_idesc_f16(128, 64), id1b = make_idesc_f16(128, 32);
        constexpr uint32_t id2a = make_idesc_f16(128, 128), id2b = make_idesc_f16(128, 64);
        const uint64_t b1 = make_kmajor_desc<128>(sbase + kOffB1), b2 = make_kmajor_desc<128>(sbase + kOffB2);
        bool have_prev = false;
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            mbar_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const bool more = ctl->slot[rs].item >= 0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
            if (more) {
                for (int s = 0; s < 4; ++s) {
                    mbar_wait(&ctl->a1_full[s], it & 1u);
                    mbar_wait(&ctl->d1_empty[s], (it & 1u) ^ 1u);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a = make_kmajor_desc<128>(sbase + kOffA1 + (uint32_t)s * 16384u);
                        const uint32_t d = tmem_base + (uint32_t)s * 64u;
                        umma_f16(d, desc_advance_k(a, 0), desc_advance_k(b1, 0), id1a, 0u);      // hi . [B_hi; B_lo]
                        umma_f16(d, desc_advance_k(a, 16), desc_advance_k(b1, 16), id1a, 1u);
                        umma_f16(d, desc_advance_k(a, 32), desc_advance_k(b1, 0), id1b, 1u);     // lo . B_hi
                        umma_f16(d, desc_advance_k(a, 48), desc_advance_k(b1, 16), id1b, 1u);
                        umma_commit(&ctl->a1_empty[s]);
                        umma_commit(&ctl->d1_full[s]);
                    }
                    __syncwarp();
                }
            }
            if (have_prev) {                                     // stage 2 of the previous item
                const uint32_t pit = it - 1;
                for (int m = 0; m < 2; ++m) {
                    mbar_wait(&ctl->a2_full[m], pit & 1u);
                    mbar_wait(&ctl->d2_empty[m], (pit & 1u) ^ 1u);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t ah = make_kmajor_desc<128>(sbase + kOffA2Hi + (uint32_t)m * 16384u);
                        const uint64_t al = make_kmajor_desc<128>(sbase + kOffA2Lo + (uint32_t)m * 16384u);
                        const uint32_t d = tmem_base + 256u + (uint32_t)m * 128u;
#pragma unroll
                        for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(ah, k), desc_advance_k(b2, k), id2a, k ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(al, k), desc_advance_k(b2, k), id2b, 1u);
                        umma_commit(&ctl->d2_full[m]);
                        if (m == 1) umma_commit(&ctl->a2_empty);
                    }
                    __syncwarp();
                }
            }
            if (!more) break;
            have_prev = true;
        }
    } else if (warp < kWarpD0) {
        // =============================== C: D1 -> twiddle -> stage-2 operand ========================================
        const int q = warp & 3;                                  // TMEM lane quadrant = frame slot inside a sub-tile
        float2 tw[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tw[k] = __ldg(reinterpret_cast<const float2*>(p.tc.twiddle) + lane * 16 + k);
        uint32_t lane_off[8];                                    // swizzled position of this lane's (re, im) word in a row
#pragma unroll
        for (int c = 0; c < 8; ++c) lane_off[c] = (uint32_t)((((lane >> 2) ^ c) << 4) | ((lane & 3) << 2));
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            mbar_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const long long item = ctl->slot[rs].item;
            const int nfr = ctl->slot[rs].nfr;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
            if (item < 0) break;
            bool a2_ready = false;
            for (int s = 0; s < 4; ++s) {
                const int f = 4 * s + q;
                const bool valid = f < nfr;
                mbar_wait(&ctl->d1_full[s], it & 1u);
                tc_fence_after();
                float y[32];
                if (valid) {
                    float u[32];
                    const uint32_t trow = tmem_base + (uint32_t)s * 64u + ((uint32_t)(q * 32) << 16);
                    tmem_ld_32x32(trow, y);                      // hi.hi + lo.hi
                    tmem_ld_32x32(trow + 32, u);                 // hi.lo
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] += u[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d1_empty[s]);
                if (!a2_ready) {                                 // stage 2 of the previous item has read the operand
                    mbar_wait(&ctl->a2_empty, (it & 1u) ^ 1u);
                    a2_ready = true;
                }
                if (valid) {
                    const uint32_t hi_base = sbase + kOffA2Hi + (uint32_t)f * 2048u;          // rows 16 f + k1
#pragma unroll
                    for (int k1 = 0; k1 < 16; ++k1) {
                        float re, im;
                        if (k1 == 0) {
                            re = y[0];
                            im = 0.f;
                        } else {
                            const float a = y[2 * k1], b = y[2 * k1 + 1], c = tw[k1 - 1].x, d = tw[k1 - 1].y;
                            re = a * c - b * d;
                            im = fmaf(a, d, b * c);
                        }
                        uint32_t hi, lo;
                        split_pair(re, im, hi, lo);
                        const uint32_t addr = hi_base + (uint32_t)k1 * 128u + lane_off[k1 & 7];
                        st_shared_b32(addr, hi);
                        st_shared_b32(addr + (kOffA2Lo - kOffA2Hi), lo);
                    }
                    {                                            // k1 = 16 (real Y): row 240 + f
                        uint32_t hi, lo;
                        split_pair(y[1] * tw[15].x, y[1] * tw[15].y, hi, lo);
                        const uint32_t addr = sbase + kOffA2Hi + (uint32_t)(240 + f) * 128u +
                                              (uint32_t)((((lane >> 2) ^ (f & 7)) << 4) | ((lane & 3) << 2));
                        st_shared_b32(addr, hi);
                        st_shared_b32(addr + (kOffA2Lo - kOffA2Hi), lo);
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) {
                    if (f < 8) mbar_arrive(&ctl->a2_full[0]);
                    mbar_arrive(&ctl->a2_full[1]);
                }
            }
        }
    } else {
        // =============================== D/E: D2 -> power -> mel -> dB -> statistics -> finisher ====================
        const int q = warp & 3;
        const int dt = (warp - kWarpD0) * 32 + lane;             // 0..127
        constexpr int kET = 128;
        float* red = ctl->red;
        const bool mfcc = p.mode == SIR_OUT_MFCC;
        const bool needs_finish = mfcc || p.mode == SIR_OUT_LOGMEL_NORM;
        const int out_rows = mfcc ? p.n_mfcc : p.n_mels;
        const int row_stride = mfcc ? p.stage_frames : p.out_frames;
        const int n_units = 3 * ((p.n_mels + 31) >> 5);
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            mbar_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const ItemSlot& sl = ctl->slot[rs];
            const long long item = sl.item;
            if (item < 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
                break;
            }
            const int b = sl.b, t0 = sl.t0, nfr = sl.nfr, T = sl.T, n_groups = sl.n_groups;
            // ---- power spectrum of the item's frames ------------------------------------------------------------
            for (int m = 0; m < 2; ++m) {
                mbar_wait(&ctl->d2_full[m], it & 1u);
                tc_fence_after();
                const int R = 128 * m + 32 * q + lane;
                const int f = R < 240 ? R >> 4 : R - 240, k1 = R < 240 ? R & 15 : 16;
                const bool valid = f < nfr && R != 255;
                if (__any_sync(0xffffffffu, valid)) {
                    const uint32_t trow = tmem_base + 256u + (uint32_t)m * 128u + ((uint32_t)(q * 32) << 16);
                    float* Pf = s_P + f * kPStride;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float v[32], u[32];
                        tmem_ld_32x32(trow + 32 * j, v);
                        tmem_ld_32x32(trow + 64 + 32 * j, u);
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float re = v[2 * i] + u[2 * i], im = v[2 * i + 1] + u[2 * i + 1];
                                const float pw = fmaf(re, re, im * im);
                                if (j == 0) {
                                    Pf[k1 + 32 * i] = pw;                                   // k2 = i: bin k1 + 32 k2
                                } else if (k1 == 0) {
                                    if (i == 0) Pf[512] = pw;                               // k2 = 16: the Nyquist bin
                                } else if (k1 != 16) {
                                    Pf[512 - k1 - 32 * i] = pw;                             // mirror: 1024 - (k1 + 32 (16 + i))
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d2_empty[m]);
            }
            mbar_wait(&ctl->sc_full[rs], (it / kRing) & 1u);     // the A warps' per-frame scales
            group_barrier();                                     // P complete

            const bool valid_utt = T > 0;
            float* __restrict__ final_out = p.out + (int64_t)b * out_rows * p.out_frames;
            float* __restrict__ out = mfcc ? p.db_stage + (int64_t)b * p.n_mels * p.stage_frames : final_out;
            if (t0 == 0 && dt == 0 && p.status) p.status[b] = valid_utt ? 0 : 1;
            bool last = false;
            if (!valid_utt) {
                for (int i = dt; i < out_rows * p.out_frames; i += kET) final_out[i] = 0.f;
            } else {
                // ---- sparse mel taps: units = (32 bands, 5 frames), most expensive bands first ------------------------
                for (;;) {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&ctl->unit_counter, 1);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= n_units) break;
                    const int bg = (n_units / 3) - 1 - u / 3, f_lo = 5 * (u % 3);
                    if (f_lo >= nfr) continue;
                    const int band = 32 * bg + lane;
                    const bool active = band < p.n_mels;
                    const int n4 = active ? s_mel_count[band] >> 2 : 0;
                    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(s_melw + (active ? s_mel_offset[band] : 0));
                    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(s_P + f_lo * kPStride + (active ? s_mel_start[band] : 0));
                    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    const int nf = min(5, nfr - f_lo);
                    for (int i = 0; i < n4; ++i) {
                        const float4 w = w4[i];
#pragma unroll
                        for (int ff = 0; ff < 5; ++ff) {
                            if (ff < nf) {
                                const float4 x = p4[ff * (kPStride / 4) + i];
                                acc[ff] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[ff]))));
                            }
                        }
                    }
                    if (active) {
#pragma unroll
                        for (int ff = 0; ff < 5; ++ff) {
                            if (ff < nf) {
                                float v = acc[ff] * sl.inv2[f_lo + ff];
                                // 10 log10(x) = (10 log10 2) lg2(x): lg2.approx is good to ~1e-7 relative here, i.e. ~1e-6 dB
                                if (p.mode != SIR_OUT_MEL_POWER) v = 3.01029995663981195f * __log2f(fmaxf(v, 1e-10f));
                                s_tile[band * 16 + f_lo + ff] = v;
                            }
                        }
                    }
                }
                group_barrier();                                 // tile complete
                if (dt == 0) ctl->unit_counter = 0;              // (next use is behind the next item's barriers)

                // ---- tile -> global + the item's statistics ----------------------------------------------------------
                const float shift = s_tile[0];
                float s1 = 0.f, s2 = 0.f, vmax = -INFINITY;
                for (int idx = dt; idx < p.n_mels * 16; idx += kET) {
                    const int mrow = idx >> 4, s = idx & 15;
                    if (s < nfr) {
                        const float v = s_tile[idx];
                        if (t0 + s < row_stride) out[(int64_t)mrow * row_stride + t0 + s] = v;
                        vmax = fmaxf(vmax, v);
                        const float d = v - shift;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                }
                if (!needs_finish) {
                    if (t0 == 0)                                 // zero padding behind the last frame, all rows
                        for (int mrow = 0; mrow < p.n_mels; ++mrow)
                            for (int tt = T + dt; tt < p.out_frames; tt += kET) out[(int64_t)mrow * p.out_frames + tt] = 0.f;
                } else {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                    }
                    const int w = warp - kWarpD0;
                    if (lane == 0) {
                        red[w] = s1;
                        red[8 + w] = s2;
                        red[16 + w] = vmax;
                    }
                    group_barrier();                             // every thread's feature stores happen before the fence below
                    if (dt == 0) {
                        ItemPartial part;
                        part.s1 = ((double)red[0] + (double)red[1]) + ((double)red[2] + (double)red[3]);
                        part.s2 = ((double)red[8] + (double)red[9]) + ((double)red[10] + (double)red[11]);
                        part.shift = shift;
                        part.vmax = fmaxf(fmaxf(red[16], red[17]), fmaxf(red[18], red[19]));
                        part.n = nfr * p.n_mels;
                        part.pad = 0;
                        p.partials[item] = part;
                        __threadfence();                         // release: the group's stores (barrier above) and the partial
                        const bool l = atomicAdd(p.counters + b, 1) == n_groups - 1;
                        if (l) __threadfence();                  // acquire: the other items' stores
                        ctl->flag = l;
                    }
                    group_barrier();
                    last = ctl->flag != 0;
                }
            }
            if (last) {
                // ---- finisher: every item of utterance b is in global memory (same merge as frontend.cu) -------------
                if (warp == kWarpD0) {
                    if (lane == 0) p.counters[b] = 0;            // ready for the next launch
                    const ItemPartial* parts = p.partials + (int64_t)b * p.groups_max;
                    double n = 0, mean = 0, m2 = 0;
                    float mx = -INFINITY;
                    for (int base = 0; base < n_groups; base += 32) {
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
                        for (int gg = 0; gg < cnt_items; ++gg) {     // merge in GROUP order: deterministic whichever CTA finishes
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
                    float* s_dct = s_P;                          // the power buffer is free between items: [n_mels][n_mfcc]
                    for (int i = dt; i < p.n_mels * p.n_mfcc; i += kET) s_dct[i] = p.dct[i];
                    group_barrier();
                    const float floor_db = p.top_db > 0.f ? red[35] - p.top_db : -INFINITY;
                    const int Tn = min(T, p.out_frames);
                    for (int idx = dt; idx < p.n_mfcc * Tn; idx += kET) {
                        const int c = idx / Tn, tt = idx - c * Tn;
                        float acc = 0.f;
                        for (int mrow = 0; mrow < p.n_mels; ++mrow)
                            acc = fmaf(fmaxf(__ldcg(out + (int64_t)mrow * row_stride + tt), floor_db), s_dct[mrow * p.n_mfcc + c], acc);
                        final_out[(int64_t)c * p.out_frames + tt] = acc;
                    }
                    for (int c = 0; c < p.n_mfcc; ++c)
                        for (int tt = T + dt; tt < p.out_frames; tt += kET) final_out[(int64_t)c * p.out_frames + tt] = 0.f;
                    group_barrier();
                    for (int i = dt; i < (int)(kPBytes / 4); i += kET) s_P[i] = 0.f;     // the pad bins must read as zero again
                } else {
                    group_barrier();
                    const float fmean = red[33], inv = red[34];
                    int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
                    if (p.masks) {
                        mt0 = p.masks[4 * b + 0];
                        mt1 = p.masks[4 * b + 1];
                        mf0 = p.masks[4 * b + 2];
                        mf1 = p.masks[4 * b + 3];
                    }
                    // normalise + mask + pad the whole utterance; its values are still L2-resident (plain loads are safe
                    // behind the acquire fence).  Loads of a batch are issued before its first store.
                    const bool out_vec = (p.out_frames % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
                    constexpr int kU = 8;
                    const int cols = out_vec ? p.out_frames >> 2 : p.out_frames;
                    const int dq = kET / cols, dr = kET % cols;
                    int m_ld = dt / cols, c_ld = dt % cols;
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
                                    const int mrow = mm[u], tt = 4 * cc[u];
                                    const bool fm = mrow >= mf0 && mrow < mf1;
                                    float4 v;
                                    v.x = (tt >= T || fm || (tt >= mt0 && tt < mt1)) ? 0.f : (x[u].x - fmean) * inv;
                                    v.y = (tt + 1 >= T || fm || (tt + 1 >= mt0 && tt + 1 < mt1)) ? 0.f : (x[u].y - fmean) * inv;
                                    v.z = (tt + 2 >= T || fm || (tt + 2 >= mt0 && tt + 2 < mt1)) ? 0.f : (x[u].z - fmean) * inv;
                                    v.w = (tt + 3 >= T || fm || (tt + 3 >= mt0 && tt + 3 < mt1)) ? 0.f : (x[u].w - fmean) * inv;
                                    *reinterpret_cast<float4*>(out + (int64_t)mrow * p.out_frames + tt) = v;
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
                                    const int mrow = mm[u], tt = cc[u];
                                    const bool masked = (tt >= mt0 && tt < mt1) || (mrow >= mf0 && mrow < mf1);
                                    out[(int64_t)mrow * p.out_frames + tt] = (tt >= T || masked) ? 0.f : (x[u] - fmean) * inv;
                                }
                            }
                        }
                    }
                }
            }
            group_barrier();                                     // P / tile / red / flag consumed; slot data no longer needed
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace fetc

// ---- host side ------------------------------------------------------------------------------------------------------
int frontend_tc_upload_tables(DeviceBuffer& buf, TcDeviceTables& dev, int sample_rate, int n_mels) {
    const fetc::HostTcTables t = fetc::build_tc_tables();
    HostFrontendTables ft = build_frontend_tables(sample_rate, n_mels);
    for (auto& w : ft.mel_weight) w *= 4.0f;                     // frontend_tables.h folds the CUDA-core post-pass's 0.25 in
    const size_t o_b1 = 0, o_b2 = o_b1 + t.b1_img.size() * 2, o_tw = o_b2 + t.b2_img.size() * 2,
                 o_mw = o_tw + t.twiddle.size() * 4, total = o_mw + ft.mel_weight.size() * 4;
    int rc = buf.reserve(total);
    if (rc != SIR_OK) return rc;
    char* base = static_cast<char*>(buf.ptr);
    SIR_CUDA(cudaMemcpy(base + o_b1, t.b1_img.data(), t.b1_img.size() * 2, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_b2, t.b2_img.data(), t.b2_img.size() * 2, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_tw, t.twiddle.data(), t.twiddle.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_mw, ft.mel_weight.data(), ft.mel_weight.size() * 4, cudaMemcpyHostToDevice));
    dev.b1_img = reinterpret_cast<const uint16_t*>(base + o_b1);
    dev.b2_img = reinterpret_cast<const uint16_t*>(base + o_b2);
    dev.twiddle = reinterpret_cast<const float*>(base + o_tw);
    dev.mel_weight = reinterpret_cast<const float*>(base + o_mw);
    return SIR_OK;
}

int frontend_tc_groups(int n_frames) { return (n_frames + fetc::kTileFrames - 1) / fetc::kTileFrames; }
long long frontend_tc_tickets(long long items, long long grid) { return items + 2 * grid; }   // every CTA draws its items + 2

int frontend_tc_launch(const FrontendParams& p, bool pcm16, int num_sms, cudaStream_t stream) {
    SIR_SMEM_OPTIN(fetc::logmel_frontend_tc_kernel<float>, fetc::kSmemBytes);
    SIR_SMEM_OPTIN(fetc::logmel_frontend_tc_kernel<short>, fetc::kSmemBytes);
    const long long items = (long long)p.batch * p.groups_max;
    const int grid = (int)(items < num_sms ? items : num_sms);
    if (pcm16)
        fetc::logmel_frontend_tc_kernel<short><<<grid, fetc::kThreads, fetc::kSmemBytes, stream>>>(p);
    else
        fetc::logmel_frontend_tc_kernel<float><<<grid, fetc::kThreads, fetc::kSmemBytes, stream>>>(p);
    return SIR_OK;
}

}  // namespace sir
