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
// One persistent CTA per SM, 16 warps, work item = 15 consecutive frames of one utterance (15 x 17 stage-2 rows = 255 = two
// 128-row UMMA tiles), drawn from a ticket counter:
//   warps 0-3   "A": load samples (lane = n2, one coalesced 128-byte request per 32 samples; a 512-sample block is loaded
//                once and serves the two frames that overlap it), per-frame power-of-two scale to max|x| in [1, 2), Hann
//                window, fp16 (hi, lo) split, write the stage-1 operand rows (frame, n2) x K = n1 (hi | lo in one 128-byte
//                swizzled row).  Warp w fills sub-tile w = frames 4w..4w+3.
//   warp 4      issues the MMAs: stage 1 per sub-tile: D1 = A_hi [B1_hi; B1_lo] (N = 64) + A_lo B1_hi (N = 32);
//                stage 2 per 128-row tile: D2 = A2_hi [B2_hi; B2_lo] (N = 128) + A2_lo B2_hi (N = 64).  The warp polls the
//                barriers of both stages and issues whichever is ready, so stage 1 of later items never queues behind a
//                stage 2 that waits for its consumers.
//   warps 5-8   "C": read D1 (lane = n2, 32 real numbers = Y[0..16]), multiply by the twiddles W1024^(n2 k1), split, and
//                write the stage-2 operand rows (frame, k1) x K = (n2, re/im): for a fixed k1 the 32 lanes write 32
//                consecutive words of one row - the transposition between the stages costs no bank conflict.
//   warps 9-12  "D/E": read D2 (lane = (frame, k1), 32 complex bins k1 + 32 k2), |X|^2 into the one-sided power spectrum
//                (the bins with k mod 32 > 16 are mirrors), sparse mel taps, dB -> one of two [n_mels][16] tiles in shared
//                memory.  These warps never touch global memory.
//   warps 13-15 "F": tile -> global, the item's partial statistics, the counting atomic and - for the item that completes
//                an utterance - the normalisation pass (as in frontend.cu).  The fence / atomic / L2 round trips of this
//                group (~2 us per item) run beside the next item's arithmetic instead of in front of it.
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

#ifndef SIR_FE_SLEEP
#define SIR_FE_SLEEP 32
#endif
#ifndef SIR_FE_PREFETCH
#define SIR_FE_PREFETCH 1
#endif
#ifndef SIR_FE_S1_FIRST
#define SIR_FE_S1_FIRST 1
#endif
constexpr int kNumWarps = 16;
constexpr int kThreads = kNumWarps * 32;
constexpr int kWarpMma = 4, kWarpD0 = 9, kWarpF0 = 13;       // warps 0-3: A, 4: MMA, 5-8: C, 9-12: D/E, 13-15: F
constexpr int kFThreads = 96;
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
    uint64_t tile_full[2], tile_empty[2];
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
constexpr uint32_t kPBytes = 32256;
constexpr uint32_t kOffMelW = kOffP + kPBytes;                   // 1536 floats
constexpr uint32_t kOffMelIdx = kOffMelW + kMelWeightCap * 4;    // start / count / offset: 3 x 128 ints
constexpr uint32_t kOffTile = kOffMelIdx + 3 * kMaxMels * 4;     // two [n_mels][16] float tiles (D/E -> F)
constexpr uint32_t kTileFloats = kMaxMels * 16;
constexpr uint32_t kOffWin = kOffTile + 2 * kTileFloats * 4;       // Hann window as [8][32 lanes][4]: lane's w[32 n1 + lane], n1 = 4c..4c+3
constexpr uint32_t kOffCtl = kOffWin + 4096;
constexpr uint32_t kSmemBytes = kOffCtl + ((sizeof(Control) + 127) & ~127u) + 1024;   // + slack for the 1024-byte alignment
static_assert(kTileFrames * kPStride * 4 <= kPBytes, "power buffer");
static_assert(kSmemBytes <= 232448, "shared memory per CTA");
static_assert(kOffA2Hi % 1024 == 0 && kOffA2Lo % 1024 == 0 && kOffB1 % 1024 == 0 && kOffB2 % 1024 == 0, "swizzle atoms");

// mbarrier wait for the pipeline hand-offs: try_wait with a suspend-time hint, so a waiting warp sleeps in hardware until the
// phase completes (or 20 us pass) instead of re-issuing the probe - with 16 warps of 5 roles on one SM the plain spin loops
// were a third of all issued instructions.  Bounded: a protocol bug traps.
__device__ __forceinline__ void pipe_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        __nanosleep(SIR_FE_SLEEP);                               // (the hint alone still re-issued the probe every ~100 cycles)
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

__device__ __forceinline__ void group_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four D/E warps
__device__ __forceinline__ void f_barrier() { asm volatile("bar.sync 2, 96;" ::: "memory"); }        // the three F warps
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
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS,
    // not generic loads - an integer round trip made every shared access of this kernel a generic one)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
            mbar_init(&ctl->ring_empty[i], 15);                  // A warps 1-3, MMA warp, 4 C warps, 4 D/E warps, 3 F warps
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
            mbar_init(&ctl->tile_full[i], 4);
            mbar_init(&ctl->tile_empty[i], 3);
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
        for (int i = tid; i < kMelWeightCap; i += kThreads) s_melw[i] = __ldg(p.tc.mel_weight + i);   // (zero behind the taps)
        for (int i = tid; i < p.n_mels; i += kThreads) {
            s_mel_start[i] = __ldg(p.tc.mel_start + i);
            s_mel_count[i] = __ldg(p.tc.mel_count + i);
            s_mel_offset[i] = __ldg(p.tc.mel_offset + i);
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
    const int64_t total_items = (int64_t)p.batch * p.groups_max;        // (the host refuses launches with >= 2^31 items)

    if (warp < 4) {
        // =============================== A: samples -> stage-1 operand ===============================================
        const float4* s_win = reinterpret_cast<const float4*>(smem + kOffWin) + lane;
        // ticket pipeline of the drawing lane: i0 / L0 = this iteration's item and its length, i1 = the next item
        // (32-bit item indices: one launch has fewer than 2^31 items; a ticket outside [0, total) means "no more work")
        const int total_i = (int)total_items;
        int i0 = -1, i1 = -1;
        int L0 = 0;
        auto length_of = [&](int item) -> int {
            if (item < 0) return 0;
            const int b = item / p.groups_max;
            int L = p.lengths ? min(__ldg(p.lengths + b), p.n_samples) : p.n_samples;
            if (p.max_samples > 0) L = min(L, p.max_samples);
            return L;
        };
        auto draw = [&]() -> int {
            const unsigned long long t = atomicAdd(p.work_counter, 1ULL) - p.work_base;      // wraps to huge if the base is ahead
            return t < (unsigned long long)total_i ? (int)t : -1;
        };
        // Request the samples of a coming item into L2 (one bulk-prefetch instruction for its 16 blocks): the A warps hold
        // only 8 KB of loads in flight per SM, which at HBM latency is ~1 TB/s for the whole chip; at L2 latency it is enough.
        auto prefetch_item = [&](int item, int Li) {
            if (item < 0 || Li <= kNfft / 2) return;
            const int b = item / p.groups_max, g = item - b * p.groups_max;
            const int n_lo = max(0, (g * kTileFrames - 1) * kHop), n_hi = min(Li, (g * kTileFrames + kTileFrames) * kHop);
            if (n_hi <= n_lo) return;
            const char* base = reinterpret_cast<const char*>(static_cast<const SampleT*>(p.wave) + (int64_t)b * p.wave_stride);
            uintptr_t lo = reinterpret_cast<uintptr_t>(base + (size_t)n_lo * sizeof(SampleT));
            uintptr_t hi = reinterpret_cast<uintptr_t>(base + (size_t)n_hi * sizeof(SampleT));
            lo = (lo + 15) & ~(uintptr_t)15;
            hi &= ~(uintptr_t)15;
#if SIR_FE_PREFETCH == 1
            if (hi > lo) prefetch_l2_bulk(reinterpret_cast<const void*>(lo), (uint32_t)(hi - lo));
#elif SIR_FE_PREFETCH == 2
            for (uintptr_t a = lo; a < hi; a += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
#endif
        };
        if (warp == 0 && lane == 0) {
            i0 = draw();
            i1 = draw();
            L0 = length_of(i0);
            prefetch_item(i0, L0);
        }
        uint32_t pub = 0;                                        // slots published so far (drawing lane)
        bool pub_end = false;
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            const uint32_t rph = (it / kRing) & 1u;
            if (warp == 0) {
                if (lane == 0) {
                    // publish ONE ITEM AHEAD: the other warps never wait for this warp to finish its own frames first
                    while (!pub_end && pub <= it + 1) {
                        const int ps = pub % kRing;
                        pipe_wait(&ctl->ring_empty[ps], ((pub / kRing) & 1u) ^ 1u);
                        ItemSlot& sl = ctl->slot[ps];
                        for (;;) {                               // skip tickets beyond an utterance's last group (ragged batches)
                            if (i0 < 0) {
                                sl.item = -1;
                                pub_end = true;
                                break;
                            }
                            const int b = i0 / p.groups_max, g = i0 - b * p.groups_max;
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
                        mbar_arrive(&ctl->ring_full[ps]);
                        ++pub;
                        if (!pub_end) {                          // advance: the next item's samples are requested into L2, the
                            const int L1 = length_of(i1);        // draw after it is in flight while this item is processed
                            prefetch_item(i1, L1);
                            i0 = i1;
                            L0 = L1;
                            i1 = draw();
                        }
                    }
                }
                __syncwarp();
            }
            pipe_wait(&ctl->ring_full[rs], rph);
            ItemSlot& sl = ctl->slot[rs];
            if (sl.item < 0) break;
            const int L = sl.L, t0 = sl.t0, nfr = sl.nfr;
            const SampleT* __restrict__ row = static_cast<const SampleT*>(p.wave) + (int64_t)sl.b * p.wave_stride;
            const int f0 = 4 * warp;
            if (f0 < nfr) {
                const uint32_t tile_addr = sbase + kOffA1 + (uint32_t)warp * 16384u;
                // blocks t0 + f0 - 1 .. t0 + f0 + 3: frame j = (block j, block j + 1); block j + 2 is requested before frame j
                // is processed
                float xa[16], xb[16], xc[16];
                const int jb = t0 + f0 - 1, nf = min(4, nfr - f0);
                // step s requests block jb + s and - from s = 2 on - turns blocks (s - 2, s - 1) into frame s - 2: one copy of
                // the load code and one of the frame code (rolled: the inlined copies thrashed the instruction cache), and
                // every block is requested two frames before it is needed
#pragma unroll 1
                for (int step = 0; step < nf + 2; ++step) {
                    if (step <= nf) load_block(row, L, jb + step, lane, xc);
                    if (step == 2) pipe_wait(&ctl->a1_empty[warp], (it & 1u) ^ 1u);   // stage 1 of the previous item has read this sub-tile
                    if (step >= 2) {
                        const float i2 = store_frame(xa, xb, s_win, tile_addr, 32 * (step - 2) + lane);
                        if (lane == 0) sl.inv2[f0 + step - 2] = i2;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        xa[i] = xb[i];
                        xb[i] = xc[i];
                    }
                }
                fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core
            } else {
                pipe_wait(&ctl->a1_empty[warp], (it & 1u) ^ 1u); // (keeps the barrier phases in step)
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
        // Two streams of work, polled in turn: stage 1 of item i1 (sub-tile s1) and stage 2 of item i2 (row tile m2), i2
        // trailing i1.  All 32 lanes walk the loop (uniform), one elected lane issues.
        constexpr uint32_t id1a = make_idesc_f16(128, 64), id1b = make_idesc_f16(128, 32);
        constexpr uint32_t id2a = make_idesc_f16(128, 128), id2b = make_idesc_f16(128, 64);
        const uint64_t b1 = make_kmajor_desc<128>(sbase + kOffB1), b2 = make_kmajor_desc<128>(sbase + kOffB2);
        uint32_t i1 = 0, s1 = 0, i2 = 0, m2 = 0, idle = 0;
        bool have_slot = false, end1 = false;
        for (;;) {
            bool progressed = false;
            // stage 1 first: its four MMAs are short and release an A warp; a stage 2 issued ahead of them (the tensor pipe
            // executes in order) would keep that warp waiting for ~1,000 cycles
#if SIR_FE_S1_FIRST
            if (!end1) {
                if (!have_slot) {
                    const int rs = i1 % kRing;
                    if (mbar_test_wait(&ctl->ring_full[rs], (i1 / kRing) & 1u)) {
                        const bool more = ctl->slot[rs].item >= 0;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
                        if (more) have_slot = true;
                        else end1 = true;
                        progressed = true;
                    }
                }
                if (have_slot && mbar_test_wait(&ctl->a1_full[s1], i1 & 1u) && mbar_test_wait(&ctl->d1_empty[s1], (i1 & 1u) ^ 1u)) {
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a = make_kmajor_desc<128>(sbase + kOffA1 + s1 * 16384u);
                        const uint32_t d = tmem_base + s1 * 64u;
                        umma_f16(d, desc_advance_k(a, 0), desc_advance_k(b1, 0), id1a, 0u);      // hi . [B_hi; B_lo]
                        umma_f16(d, desc_advance_k(a, 16), desc_advance_k(b1, 16), id1a, 1u);
                        umma_f16(d, desc_advance_k(a, 32), desc_advance_k(b1, 0), id1b, 1u);     // lo . B_hi
                        umma_f16(d, desc_advance_k(a, 48), desc_advance_k(b1, 16), id1b, 1u);
                        umma_commit(&ctl->a1_empty[s1]);
                        umma_commit(&ctl->d1_full[s1]);
                    }
                    __syncwarp();
                    if (++s1 == 4) {
                        s1 = 0;
                        ++i1;
                        have_slot = false;
                    }
                    progressed = true;
                }
            }
            if (i2 < i1 && mbar_test_wait(&ctl->a2_full[m2], i2 & 1u) && mbar_test_wait(&ctl->d2_empty[m2], (i2 & 1u) ^ 1u)) {
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t ah = make_kmajor_desc<128>(sbase + kOffA2Hi + m2 * 16384u);
                    const uint64_t al = make_kmajor_desc<128>(sbase + kOffA2Lo + m2 * 16384u);
                    const uint32_t d = tmem_base + 256u + m2 * 128u;
#pragma unroll
                    for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(ah, k), desc_advance_k(b2, k), id2a, k ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(al, k), desc_advance_k(b2, k), id2b, 1u);
                    umma_commit(&ctl->d2_full[m2]);
                    if (m2 == 1) umma_commit(&ctl->a2_empty);
                }
                __syncwarp();
                if (++m2 == 2) {
                    m2 = 0;
                    ++i2;
                }
                progressed = true;
            }
#else
            if (i2 < i1 && mbar_test_wait(&ctl->a2_full[m2], i2 & 1u) && mbar_test_wait(&ctl->d2_empty[m2], (i2 & 1u) ^ 1u)) {
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t ah = make_kmajor_desc<128>(sbase + kOffA2Hi + m2 * 16384u);
                    const uint64_t al = make_kmajor_desc<128>(sbase + kOffA2Lo + m2 * 16384u);
                    const uint32_t d = tmem_base + 256u + m2 * 128u;
#pragma unroll
                    for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(ah, k), desc_advance_k(b2, k), id2a, k ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 64; k += 16) umma_f16(d, desc_advance_k(al, k), desc_advance_k(b2, k), id2b, 1u);
                    umma_commit(&ctl->d2_full[m2]);
                    if (m2 == 1) umma_commit(&ctl->a2_empty);
                }
                __syncwarp();
                if (++m2 == 2) {
                    m2 = 0;
                    ++i2;
                }
                progressed = true;
            }
            if (!end1) {
                if (!have_slot) {
                    const int rs = i1 % kRing;
                    if (mbar_test_wait(&ctl->ring_full[rs], (i1 / kRing) & 1u)) {
                        const bool more = ctl->slot[rs].item >= 0;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
                        if (more) have_slot = true;
                        else end1 = true;
                        progressed = true;
                    }
                }
                if (have_slot && mbar_test_wait(&ctl->a1_full[s1], i1 & 1u) && mbar_test_wait(&ctl->d1_empty[s1], (i1 & 1u) ^ 1u)) {
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a = make_kmajor_desc<128>(sbase + kOffA1 + s1 * 16384u);
                        const uint32_t d = tmem_base + s1 * 64u;
                        umma_f16(d, desc_advance_k(a, 0), desc_advance_k(b1, 0), id1a, 0u);      // hi . [B_hi; B_lo]
                        umma_f16(d, desc_advance_k(a, 16), desc_advance_k(b1, 16), id1a, 1u);
                        umma_f16(d, desc_advance_k(a, 32), desc_advance_k(b1, 0), id1b, 1u);     // lo . B_hi
                        umma_f16(d, desc_advance_k(a, 48), desc_advance_k(b1, 16), id1b, 1u);
                        umma_commit(&ctl->a1_empty[s1]);
                        umma_commit(&ctl->d1_full[s1]);
                    }
                    __syncwarp();
                    if (++s1 == 4) {
                        s1 = 0;
                        ++i1;
                        have_slot = false;
                    }
                    progressed = true;
                }
            }
#endif
            if (end1 && i2 == i1) break;
            if (progressed) {
                idle = 0;
            } else {
                __nanosleep(40);                                 // nothing ready: leave the issue slots to the working warps
                if (++idle > (1u << 24)) __trap();               // a protocol bug must surface as an error, never as a hang
            }
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
            pipe_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const long long item = ctl->slot[rs].item;
            const int nfr = ctl->slot[rs].nfr;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
            if (item < 0) break;
            bool a2_ready = false;
            for (int s = 0; s < 4; ++s) {
                const int f = 4 * s + q;
                const bool valid = f < nfr;
                pipe_wait(&ctl->d1_full[s], it & 1u);
                tc_fence_after();
                float y[32];
                if (valid) {
                    float u[32];
                    const uint32_t trow = tmem_base + (uint32_t)s * 64u + ((uint32_t)(q * 32) << 16);
                    tmem_ld_32x32_pair(trow, trow + 32, y, u);   // (hi.hi + lo.hi), (hi.lo)
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] += u[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d1_empty[s]);
                if (!a2_ready) {                                 // stage 2 of the previous item has read the operand
                    pipe_wait(&ctl->a2_empty, (it & 1u) ^ 1u);
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
    } else if (warp < kWarpF0) {
        // =============================== D/E: D2 -> power -> mel -> dB -> tile (shared memory only) ==================
        const int q = warp & 3;
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            pipe_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const ItemSlot& sl = ctl->slot[rs];
            if (sl.item < 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
                break;
            }
            const int nfr = sl.nfr;
            // ---- power spectrum of the item's frames ------------------------------------------------------------
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                pipe_wait(&ctl->d2_full[m], it & 1u);
                tc_fence_after();
                const int R = 128 * m + 32 * q + lane;
                const int f = R < 240 ? R >> 4 : R - 240, k1 = R < 240 ? R & 15 : 16;
                const bool valid = f < nfr && R != 255;
                if (__any_sync(0xffffffffu, valid)) {
                    const uint32_t trow = tmem_base + 256u + (uint32_t)m * 128u + ((uint32_t)(q * 32) << 16);
                    float* Pf = s_P + f * kPStride;
#pragma unroll 1
                    for (int j = 0; j < 2; ++j) {                // (rolled: code size)
                        // where this lane's 16 bins of column chunk j go: base + step * i for i < count (no branch in the loop)
                        //   j = 0 (k2 = i):        bin k1 + 32 i
                        //   j = 1 (k2 = 16 + i):   mirror 1024 - (k1 + 32 k2) = 512 - k1 - 32 i  (k1 = 1..15);
                        //                          k1 = 0: only k2 = 16, the Nyquist bin 512;  k1 = 16: nothing
                        int base, step, count;
                        if (j == 0) {
                            base = k1; step = 32; count = 16;
                        } else {
                            base = 512 - k1; step = -32; count = k1 == 0 ? 1 : (k1 == 16 ? 0 : 16);
                        }
                        if (!valid) count = 0;
                        float v[32], u[32];
                        tmem_ld_32x32_pair(trow + 32 * j, trow + 64 + 32 * j, v, u);
                        float* dst = Pf + base;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float re = v[2 * i] + u[2 * i], im = v[2 * i + 1] + u[2 * i + 1];
                            const float pw = fmaf(re, re, im * im);
                            if (i < count) dst[step * i] = pw;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d2_empty[m]);
            }
            pipe_wait(&ctl->sc_full[rs], (it / kRing) & 1u);     // the A warps' per-frame scales
            const uint32_t buf = it & 1u;
            float* tile = s_tile + buf * kTileFloats;
            pipe_wait(&ctl->tile_empty[buf], ((it >> 1) & 1u) ^ 1u);   // the F warps have drained this tile
            group_barrier();                                     // P complete
            // ---- sparse mel taps: lane = (band, frame).  The 16 lanes of a half-warp read the same four taps (broadcast)
            // and their own frame's four bins (rows 532 floats apart: conflict-free LDS.128).  A warp takes QUADS of
            // neighbouring bands (4q..4q+3; the host pads their tap runs to one length), two bands per lane, the loads of
            // tap group i + 1 in flight while group i is accumulated; quads go round-robin over the four warps.
            if (nfr > 0) {
                const int w = warp - kWarpD0, h = lane >> 4, f = lane & 15;
                const int n_quads = (p.n_mels + 3) >> 2;
                const float* Pf = s_P + (f < nfr ? f : 0) * kPStride;
                const float i2 = sl.inv2[f < nfr ? f : 0];
                for (int pq = w; pq < n_quads; pq += 4) {
                    const int band0 = 4 * pq + h, band1 = band0 + 2;
                    const int bq = min(band0, p.n_mels - 1), br = min(band1, p.n_mels - 1);
                    const int n4 = s_mel_count[4 * pq] >> 2;     // the same for the whole quad
                    const float4* __restrict__ wa = reinterpret_cast<const float4*>(s_melw + s_mel_offset[bq]);
                    const float4* __restrict__ wb = reinterpret_cast<const float4*>(s_melw + s_mel_offset[br]);
                    const float4* __restrict__ pa = reinterpret_cast<const float4*>(Pf + s_mel_start[bq]);
                    const float4* __restrict__ pb = reinterpret_cast<const float4*>(Pf + s_mel_start[br]);
                    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
                    float4 wva = wa[0], xa = pa[0], wvb = wb[0], xb = pb[0];
                    for (int i = 0; i < n4; ++i) {
                        const int nx = i + 1 < n4 ? i + 1 : i;
                        const float4 nwa = wa[nx], nxa = pa[nx], nwb = wb[nx], nxb = pb[nx];
                        a0 = fmaf(wva.x, xa.x, fmaf(wva.z, xa.z, a0));
                        a1 = fmaf(wva.y, xa.y, fmaf(wva.w, xa.w, a1));
                        b0 = fmaf(wvb.x, xb.x, fmaf(wvb.z, xb.z, b0));
                        b1 = fmaf(wvb.y, xb.y, fmaf(wvb.w, xb.w, b1));
                        wva = nwa; xa = nxa; wvb = nwb; xb = nxb;
                    }
                    if (f < nfr) {
                        float va = (a0 + a1) * i2, vb = (b0 + b1) * i2;
                        // 10 log10(x) = (10 log10 2) lg2(x): lg2.approx is good to ~1e-7 relative here, i.e. ~1e-6 dB
                        if (p.mode != SIR_OUT_MEL_POWER) {
                            va = 3.01029995663981195f * __log2f(fmaxf(va, 1e-10f));
                            vb = 3.01029995663981195f * __log2f(fmaxf(vb, 1e-10f));
                        }
                        if (band0 < p.n_mels) tile[band0 * 16 + f] = va;
                        if (band1 < p.n_mels) tile[band1 * 16 + f] = vb;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->tile_full[buf]);    // release: this warp's tile entries
            group_barrier();                                     // every warp is done with P
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
        }
    } else {
        // =============================== F: tile -> global, statistics, counting atomic, finisher ====================
        const int ft = (warp - kWarpF0) * 32 + lane;             // 0..95
        constexpr int kET = kFThreads;
        float* red = ctl->red;
        const bool mfcc = p.mode == SIR_OUT_MFCC;
        const bool needs_finish = mfcc || p.mode == SIR_OUT_LOGMEL_NORM;
        const int out_rows = mfcc ? p.n_mfcc : p.n_mels;
        const int row_stride = mfcc ? p.stage_frames : p.out_frames;
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            pipe_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const ItemSlot& sl = ctl->slot[rs];
            const long long item = sl.item;
            if (item < 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
                break;
            }
            const int b = sl.b, t0 = sl.t0, nfr = sl.nfr, T = sl.T;
            const uint32_t buf = it & 1u;
            const float* tile = s_tile + buf * kTileFloats;
            pipe_wait(&ctl->tile_full[buf], (it >> 1) & 1u);

            const bool valid_utt = T > 0;
            float* __restrict__ final_out = p.out + (int64_t)b * out_rows * p.out_frames;
            float* __restrict__ out = mfcc ? p.db_stage + (int64_t)b * p.n_mels * p.stage_frames : final_out;
            if (t0 == 0 && ft == 0 && p.status) p.status[b] = valid_utt ? 0 : 1;
            if (!valid_utt) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->tile_empty[buf]);
                for (int i = ft; i < out_rows * p.out_frames; i += kET) final_out[i] = 0.f;
            } else {
                // ---- tile -> global + the item's statistics ----------------------------------------------------------
                const float shift = tile[0];
                float s1 = 0.f, s2 = 0.f, vmax = -INFINITY;
                for (int idx = ft; idx < p.n_mels * 16; idx += kET) {
                    const int mrow = idx >> 4, s = idx & 15;
                    if (s < nfr) {
                        const float v = tile[idx];
                        if (t0 + s < row_stride) out[(int64_t)mrow * row_stride + t0 + s] = v;
                        vmax = fmaxf(vmax, v);
                        const float d = v - shift;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->tile_empty[buf]);        // the D/E warps may refill this tile
                if (!needs_finish) {
                    if (t0 == 0)                                 // zero padding behind the last frame, all rows
                        for (int mrow = 0; mrow < p.n_mels; ++mrow)
                            for (int tt = T + ft; tt < p.out_frames; tt += kET) out[(int64_t)mrow * p.out_frames + tt] = 0.f;
                } else {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                    }
                    const int w = warp - kWarpF0;
                    if (lane == 0) {
                        red[w] = s1;
                        red[8 + w] = s2;
                        red[16 + w] = vmax;
                    }
                    f_barrier();
                    if (ft == 0) {                               // the item's partial statistics; frontend_finish_kernel merges them
                        ItemPartial part;
                        part.s1 = ((double)red[0] + (double)red[1]) + (double)red[2];
                        part.s2 = ((double)red[8] + (double)red[9]) + (double)red[10];
                        part.shift = shift;
                        part.vmax = fmaxf(fmaxf(red[16], red[17]), red[18]);
                        part.n = nfr * p.n_mels;
                        part.pad = 0;
                        p.partials[item] = part;
                    }
                }
            }
            f_barrier();                                         // red consumed; slot data no longer needed
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

// Second launch of the NORM / MFCC modes: one CTA per utterance merges the item partials in GROUP order (Chan's formula,
// fp64: deterministic), re-reads the utterance's un-normalised dB values (L2-resident for all but the largest batches),
// normalises, applies the SpecAugment bands and writes the zero padding; MFCC: floor at max - top_db, ortho DCT-II.
// (The CUDA-core kernel does this inside the main launch, behind a fence + counting atomic per item; with one CTA per SM
// that latency chain - ~2 us per item - sat in front of the next item's arithmetic.)
constexpr int kFinThreads = 128;
__global__ void __launch_bounds__(kFinThreads) frontend_finish_kernel(const FrontendParams p) {
    __shared__ float red[8];
    extern __shared__ float s_dct[];
    // gridDim.y CTAs share an utterance (each merges the partials itself - a few dozen flops - and takes a slice of the rows)
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.y, parts = gridDim.y;
    int L = p.lengths ? min(p.lengths[b], p.n_samples) : p.n_samples;
    if (p.max_samples > 0) L = min(L, p.max_samples);
    if (L <= kNfft / 2) return;                                  // invalid utterance: zero-filled by the main kernel
    const int T = 1 + L / kHop, n_groups = (T + kTileFrames - 1) / kTileFrames;
    const bool mfcc = p.mode == SIR_OUT_MFCC;
    const int out_rows = mfcc ? p.n_mfcc : p.n_mels;
    const int row_stride = mfcc ? p.stage_frames : p.out_frames;
    float* __restrict__ final_out = p.out + (int64_t)b * out_rows * p.out_frames;
    float* __restrict__ out = mfcc ? p.db_stage + (int64_t)b * p.n_mels * p.stage_frames : final_out;
    if (mfcc)
        for (int i = tid; i < p.n_mels * p.n_mfcc; i += kFinThreads) s_dct[i] = p.dct[i];
    if (warp == 0) {
        const ItemPartial* parts = p.partials + (int64_t)b * p.groups_max;
        double n = 0, mean = 0, m2 = 0;
        float mx = -INFINITY;
        for (int base = 0; base < n_groups; base += 32) {
            double ni = 0, mi = 0, m2i = 0;
            if (base + lane < n_groups) {
                const ItemPartial part = parts[base + lane];
                ni = (double)part.n;
                const double r = part.s1 / ni;
                mi = (double)part.shift + r;
                m2i = fmax(part.s2 - part.s1 * r, 0.0);
                mx = fmaxf(mx, part.vmax);
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
            red[0] = (float)mean;
            red[1] = (float)(1.0 / (sqrt(m2 / (n - 1.0)) + 1e-5));
            red[2] = mx;
        }
    }
    __syncthreads();
    if (mfcc) {
        const float floor_db = p.top_db > 0.f ? red[2] - p.top_db : -INFINITY;
        const int Tn = min(T, p.out_frames);
        for (int idx = part * kFinThreads + tid; idx < p.n_mfcc * Tn; idx += parts * kFinThreads) {
            const int c = idx / Tn, tt = idx - c * Tn;
            float acc = 0.f;
            for (int mrow = 0; mrow < p.n_mels; ++mrow)
                acc = fmaf(fmaxf(out[(int64_t)mrow * row_stride + tt], floor_db), s_dct[mrow * p.n_mfcc + c], acc);
            final_out[(int64_t)c * p.out_frames + tt] = acc;
        }
        for (int c = part; c < p.n_mfcc; c += parts)
            for (int tt = T + tid; tt < p.out_frames; tt += kFinThreads) final_out[(int64_t)c * p.out_frames + tt] = 0.f;
        return;
    }
    const float fmean = red[0], inv = red[1];
    int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
    if (p.masks) {
        mt0 = p.masks[4 * b + 0];
        mt1 = p.masks[4 * b + 1];
        mf0 = p.masks[4 * b + 2];
        mf1 = p.masks[4 * b + 3];
    }
    const bool out_vec = (p.out_frames % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (out_vec) {
        const int cols = p.out_frames >> 2, all = p.n_mels * cols;
        const int per = (all + parts - 1) / parts, total = min(all, (part + 1) * per);
        constexpr int kU = 4;
        for (int i0 = part * per + tid; i0 < total; i0 += kU * kFinThreads) {
            float4 x[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {                      // the loads of a batch are issued before its first store
                const int i = i0 + u * kFinThreads;
                x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int mrow = i / cols, tt = 4 * (i - mrow * cols);
                    if (tt < T) x[u] = *reinterpret_cast<const float4*>(out + (int64_t)mrow * p.out_frames + tt);
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * kFinThreads;
                if (i < total) {
                    const int mrow = i / cols, tt = 4 * (i - mrow * cols);
                    const bool fm = mrow >= mf0 && mrow < mf1;
                    float4 v;
                    v.x = (tt >= T || fm || (tt >= mt0 && tt < mt1)) ? 0.f : (x[u].x - fmean) * inv;
                    v.y = (tt + 1 >= T || fm || (tt + 1 >= mt0 && tt + 1 < mt1)) ? 0.f : (x[u].y - fmean) * inv;
                    v.z = (tt + 2 >= T || fm || (tt + 2 >= mt0 && tt + 2 < mt1)) ? 0.f : (x[u].z - fmean) * inv;
                    v.w = (tt + 3 >= T || fm || (tt + 3 >= mt0 && tt + 3 < mt1)) ? 0.f : (x[u].w - fmean) * inv;
                    *reinterpret_cast<float4*>(out + (int64_t)mrow * p.out_frames + tt) = v;
                }
            }
        }
    } else {
        const int total = p.n_mels * p.out_frames;
        for (int i = part * kFinThreads + tid; i < total; i += parts * kFinThreads) {
            const int mrow = i / p.out_frames, tt = i - mrow * p.out_frames;
            const bool masked = (tt >= mt0 && tt < mt1) || (mrow >= mf0 && mrow < mf1);
            const float x = tt < T ? out[(int64_t)mrow * p.out_frames + tt] : 0.f;
            out[(int64_t)mrow * p.out_frames + tt] = (tt >= T || masked) ? 0.f : (x - fmean) * inv;
        }
    }
}

}  // namespace fetc

// ---- host side ------------------------------------------------------------------------------------------------------
int frontend_tc_upload_tables(DeviceBuffer& buf, TcDeviceTables& dev, int sample_rate, int n_mels) {
    const fetc::HostTcTables t = fetc::build_tc_tables();
    const HostFrontendTables ft = build_frontend_tables(sample_rate, n_mels);
    // The kernel walks the taps of a QUAD of neighbouring bands (4q .. 4q + 3) in lock step: their runs are padded with zero
    // weights to the longest one (the extra taps multiply finite power bins, the rows of the power buffer are zero behind
    // bin 512).  frontend_tables.h folds the CUDA-core post-pass's 0.25 into its weights: undone here.
    std::vector<int32_t> m_start(n_mels), m_count(n_mels), m_offset(n_mels);
    std::vector<float> m_weight;
    for (int m = 0; m < n_mels; ++m) {
        int n = 0;
        for (int o = m & ~3; o < (m & ~3) + 4 && o < n_mels; ++o) n = ft.mel_count[o] > n ? ft.mel_count[o] : n;
        if (ft.mel_start[m] + n > fetc::kPStride) return fail(SIR_ERR_UNSUPPORTED, "mel band %d: padded run leaves the power row", m);
        m_start[m] = ft.mel_start[m];
        m_count[m] = n;
        m_offset[m] = (int32_t)m_weight.size();
        for (int i = 0; i < n; ++i) m_weight.push_back(i < ft.mel_count[m] ? 4.0f * ft.mel_weight[ft.mel_offset[m] + i] : 0.f);
    }
    if ((int)m_weight.size() > fetc::kMelWeightCap)
        return fail(SIR_ERR_UNSUPPORTED, "filterbank has %zu padded taps (cap %d)", m_weight.size(), fetc::kMelWeightCap);
    const size_t o_b1 = 0, o_b2 = o_b1 + t.b1_img.size() * 2, o_tw = o_b2 + t.b2_img.size() * 2,
                 o_mw = o_tw + t.twiddle.size() * 4, o_ms = o_mw + fetc::kMelWeightCap * 4, o_mc = o_ms + n_mels * 4,
                 o_mo = o_mc + n_mels * 4, total = o_mo + n_mels * 4;
    int rc = buf.reserve(total);
    if (rc != SIR_OK) return rc;
    char* base = static_cast<char*>(buf.ptr);
    SIR_CUDA(cudaMemcpy(base + o_b1, t.b1_img.data(), t.b1_img.size() * 2, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_b2, t.b2_img.data(), t.b2_img.size() * 2, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_tw, t.twiddle.data(), t.twiddle.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemset(base + o_mw, 0, fetc::kMelWeightCap * 4));
    SIR_CUDA(cudaMemcpy(base + o_mw, m_weight.data(), m_weight.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_ms, m_start.data(), n_mels * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_mc, m_count.data(), n_mels * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_mo, m_offset.data(), n_mels * 4, cudaMemcpyHostToDevice));
    dev.b1_img = reinterpret_cast<const uint16_t*>(base + o_b1);
    dev.b2_img = reinterpret_cast<const uint16_t*>(base + o_b2);
    dev.twiddle = reinterpret_cast<const float*>(base + o_tw);
    dev.mel_weight = reinterpret_cast<const float*>(base + o_mw);
    dev.mel_start = reinterpret_cast<const int32_t*>(base + o_ms);
    dev.mel_count = reinterpret_cast<const int32_t*>(base + o_mc);
    dev.mel_offset = reinterpret_cast<const int32_t*>(base + o_mo);
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
    if (p.mode == SIR_OUT_LOGMEL_NORM || p.mode == SIR_OUT_MFCC) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(SIR_ERR_CUDA, "launch of logmel_frontend_tc_kernel failed: %s", cudaGetErrorString(e));
        count_launch();
        const size_t dct_bytes = p.mode == SIR_OUT_MFCC ? (size_t)p.n_mels * p.n_mfcc * sizeof(float) : 0;
        // small batches: several CTAs per utterance, so that the pass is not one wave of under-occupied SMs
        const int parts = p.batch >= 2048 ? 1 : (p.batch >= 512 ? 2 : 4);
        fetc::frontend_finish_kernel<<<dim3((unsigned)p.batch, (unsigned)parts), fetc::kFinThreads, dct_bytes, stream>>>(p);
    }
    return SIR_OK;
}

}  // namespace sir
