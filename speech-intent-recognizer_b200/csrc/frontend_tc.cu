// Fused log-mel frontend on the sm_100a TENSOR CORES: the 1024-point real DFT of every frame is two matrix stages on
// tcgen05 (32 x 32 Cooley-Tukey, TF32 (hi, lo) split operands, fp32 accumulators in tensor memory); CUDA cores only
// window / split the samples, apply the inter-stage twiddles, square, and run the sparse mel projection.
//
// Replaces (per utterance) scripts/precompute_features.py:59-73, scripts/dataset.py:105-113,160-176 of the reference and
// the torchaudio calls behind them, like frontend.cu, whose CUDA-core FFT it supersedes: that kernel
// needs ~1,360 warp instructions per frame and is issue / latency bound at 8-11 % of the HBM roofline; here the butterflies
// are MMAs.  Numerics: frontend_tc_tables.h, tests/host/tc_dft_host_check.cpp (within ~5x of an fp32 FFT's own rounding
// error).  TF32 pieces need no scaling (fp32 exponent range) and the split is a mask and a subtraction; the first version
// of this kernel used fp16 pieces, which cost a maximum over every windowed frame, a power-of-two rescale and three
// conversions per value on the CUDA cores - the resource this kernel is bound by.
//
// One persistent CTA per SM, 16 warps, work item = 14 consecutive frames of one utterance = two self-contained 128-row UMMA
// tiles of 7 frames x 17 stage-2 rows (a true double buffer between the twiddle warps and the stage-2 MMAs), drawn from a
// ticket counter.  The kernel is bound by the shared-memory data pipe (profiles/r2_summary.md):
//   warps 0-3   "A": load samples (lane = n2, one coalesced 128-byte request per 32 samples; a 512-sample block is loaded
//                once and serves the two frames that overlap it), Hann window, TF32 (hi, lo) split, and write the stage-1
//                operand rows (frame, n2) x K = n1 STRAIGHT INTO TENSOR MEMORY (tcgen05.st: lane = row, 32 columns hi,
//                32 columns lo).  Warp w takes frames 4w..4w+3 and owns lane quadrant w of every sub-tile: frame 4w + s is
//                quadrant w of sub-tile s, so sub-tile s is complete after every warp's s-th frame.
//   warp 4      issues the MMAs: stage 1 per sub-tile, A from tensor memory: D1 = A_hi B1_hi + A_hi B1_lo + A_lo B1_hi (12
//                instructions of K = 8, N = 32); stage 2 per 128-row tile, A from shared memory: 24 instructions of K = 8,
//                N = 64.  The warp polls the barriers of both stages and issues whichever is ready.
//   warps 5-8   "C": read D1 (lane = n2, 32 real numbers = Y[0..16]), multiply by the twiddles W1024^(n2 k1), split, and
//                write the stage-2 operand rows (frame, k1) x K = (n2, re/im): for a fixed k1 the 32 lanes write 256
//                consecutive bytes - the transposition between the stages costs no bank conflict.
//   warps 9-12  "D/E": read D2 (lane = (frame, k1), 32 complex bins k1 + 32 k2), |X|^2 into the one-sided power spectrum
//                (the bins with k mod 32 > 16 are mirrors), sparse mel taps, dB -> one of two [n_mels][16] tiles in shared
//                memory.  These warps never touch global memory.
//   warps 13-14 "F": tile -> global and the item's partial statistics (frontend_finish_kernel merges them).
//   warp 15     draws the tickets and publishes the items to the other warps (three ahead of the A warps).
// Every hand-off is an mbarrier; every wait is bounded and traps.
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
#ifndef SIR_FE_CHUNK
#define SIR_FE_CHUNK 2                          // (pass, column block) chunks of stage 2 issued between looks at stage 1: 1, 2, 3 or 6
#endif
// Optional timeline of the first CTAs' first items (build with -DSIR_FE_TRACE; tools/fe_trace.py reads it): clock64 at the
// hand-offs of every role, to see which dependency the pipeline waits for.  Not compiled into the product library.
#ifdef SIR_FE_TRACE
constexpr int kTraceCtas = 4, kTraceItems = 96, kTraceEvents = 32;
__device__ long long g_fe_trace[kTraceCtas][kTraceItems][kTraceEvents];
#define FE_TRACE(EV, IT)                                                                                    \
    do {                                                                                                     \
        if (blockIdx.x < kTraceCtas && (IT) < (uint32_t)kTraceItems && lane == 0) g_fe_trace[blockIdx.x][IT][EV] = clock64(); \
    } while (0)
#else
#define FE_TRACE(EV, IT) do { } while (0)
#endif
constexpr int kNumWarps = 16;
constexpr int kThreads = kNumWarps * 32;
constexpr int kWarpMma = 4, kWarpD0 = 9, kWarpF0 = 13, kWarpPub = 15;   // warps 0-3: A, 4: MMA, 5-8: C, 9-12: D/E, 13-14: F, 15: items
constexpr int kFThreads = 64;
constexpr int kRing = 8;                       // item slots between the ticket drawer and the other warps
constexpr int kMelWeightCap = 1536;

struct ItemSlot {
    long long item;                            // < 0: no more work
    int b, t0, nfr, T, L, n_groups;
};
struct Control {
    uint64_t ring_full[kRing], ring_empty[kRing];
    uint64_t a1_full[4], a1_empty[4], d1_full[4], d1_empty[4];
    uint64_t a2_full[2], a2_empty[2], d2_full[2], d2_empty[2];
    uint64_t tile_full[2], tile_empty[2];
    ItemSlot slot[kRing];
    uint32_t tmem_base;
    volatile uint32_t a_progress;              // the item A warp 0 works on (the publisher stays at most three ahead of it)
    float red[48];
};

// shared-memory carve-up (bytes from the 1024-aligned base).  The stage-1 data operand lives in tensor memory.
constexpr uint32_t kOffA2Hi = 0, kOffA2Lo = 65536;               // each: K 0..31 then K 32..63 (32 KB apart), 256 rows x 128 B
constexpr uint32_t kA2KBlock = 32768;
constexpr uint32_t kOffB1 = 131072;                              // hi tile, lo tile: 32 rows x 128 B each
constexpr uint32_t kOffB2 = 139264;                              // hi K 0..31, hi K 32..63, lo K 0..31, lo K 32..63: 64 rows x 128 B each
constexpr uint32_t kOffP = 172032;                               // 14 x 532 floats
constexpr uint32_t kPBytes = 29824;
constexpr uint32_t kOffMelW = kOffP + kPBytes;                   // 1536 floats
constexpr uint32_t kOffMelIdx = kOffMelW + kMelWeightCap * 4;    // start / count / offset: 3 x 128 ints
constexpr uint32_t kOffTile = kOffMelIdx + 3 * kMaxMels * 4;     // two [n_mels][16] float tiles (D/E -> F)
constexpr uint32_t kTileFloats = kMaxMels * 16;
constexpr uint32_t kOffWin = kOffTile + 2 * kTileFloats * 4;     // Hann window as [8][32 lanes][4]: lane's w[32 n1 + lane], n1 = 4c..4c+3
constexpr uint32_t kOffCtl = kOffWin + 4096;
constexpr uint32_t kSmemBytes = kOffCtl + ((sizeof(Control) + 127) & ~127u) + 1024;   // + slack for the 1024-byte alignment
static_assert(kTileFrames * kPStride * 4 <= kPBytes, "power buffer");
static_assert(kSmemBytes <= 232448, "shared memory per CTA");
static_assert(kOffA2Lo % 1024 == 0 && kOffB1 % 1024 == 0 && kOffB2 % 1024 == 0, "swizzle atoms");
// tensor-memory columns (512): stage-1 accumulators, stage-1 data operand (hi | lo per sub-tile), stage-2 accumulators
constexpr uint32_t kColD1 = 0, kColA1 = 128, kColD2 = 384;

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
__device__ __forceinline__ void f_barrier() { asm volatile("bar.sync 2, 64;" ::: "memory"); }        // the two F warps
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// what the tensor core does not read of an fp32 operand: x minus x with its 13 low mantissa bits cleared (exact)
__device__ __forceinline__ uint32_t tf32_lo_bits(uint32_t x_bits) {
    return __float_as_uint(__uint_as_float(x_bits) - __uint_as_float(x_bits & 0xFFFFE000u));
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

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// Half a frame = one 512-sample block in its role as n1 = 16 h .. 16 h + 15 of the frame: Hann window, then columns
// [16 h, 16 h + 16) of the row (frame, n2 = lane) of the stage-1 operand in tensor memory get the windowed values themselves
// (the tensor core reads their 19 leading bits) and columns 32 + [16 h, 16 h + 16) what it does not read of them.
template <int H>
__device__ __forceinline__ void store_half_frame(const float (&x)[16], const float4* __restrict__ g_win, uint32_t taddr) {
    uint32_t t[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 w = g_win[(4 * H + c) * 32];                // (this lane's) w[32 n1 + lane], n1 = 16 H + 4c .. + 3
        t[4 * c + 0] = __float_as_uint(x[4 * c + 0] * w.x);
        t[4 * c + 1] = __float_as_uint(x[4 * c + 1] * w.y);
        t[4 * c + 2] = __float_as_uint(x[4 * c + 2] * w.z);
        t[4 * c + 3] = __float_as_uint(x[4 * c + 3] * w.w);
    }
    tmem_st_32x16(taddr + 16u * H, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = tf32_lo_bits(t[i]);
    tmem_st_32x16(taddr + 32u + 16u * H, t);
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
            mbar_init(&ctl->ring_empty[i], 15);                  // 4 A warps, MMA warp, 4 C warps, 4 D/E warps, 2 F warps
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&ctl->a1_full[i], 4);                      // every A warp owns one lane quadrant of every sub-tile
            mbar_init(&ctl->a1_empty[i], 1);
            mbar_init(&ctl->d1_full[i], 1);
            mbar_init(&ctl->d1_empty[i], 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ctl->a2_full[i], 7);                      // seven frames per stage-2 tile (one arrival per C warp and frame)
            mbar_init(&ctl->a2_empty[i], 1);
            mbar_init(&ctl->d2_full[i], 1);
            mbar_init(&ctl->d2_empty[i], 4);
            mbar_init(&ctl->tile_full[i], 4);
            mbar_init(&ctl->tile_empty[i], 2);
        }
        ctl->a_progress = 0;
        fence_barrier_init();
    }
    if (warp == kWarpMma) tmem_alloc<512>(&ctl->tmem_base);
    {
        const uint4* g1 = reinterpret_cast<const uint4*>(p.tc.b1_img);
        const uint4* g2 = reinterpret_cast<const uint4*>(p.tc.b2_img);
        uint4* s1 = reinterpret_cast<uint4*>(smem + kOffB1);
        uint4* s2 = reinterpret_cast<uint4*>(smem + kOffB2);
        for (int i = tid; i < 512; i += kThreads) s1[i] = __ldg(g1 + i);
        for (int i = tid; i < 2048; i += kThreads) s2[i] = __ldg(g2 + i);
        for (int i = tid; i < kMelWeightCap; i += kThreads) s_melw[i] = __ldg(p.tc.mel_weight + i);   // (zero behind the taps)
        for (int i = tid; i < p.n_mels; i += kThreads) {
            s_mel_start[i] = __ldg(p.tc.mel_start + i);
            s_mel_count[i] = __ldg(p.tc.mel_count + i);
            s_mel_offset[i] = __ldg(p.tc.mel_offset + i);
        }
        for (int i = tid; i < 256; i += kThreads)
            reinterpret_cast<uint4*>(smem + kOffWin)[i] = __ldg(reinterpret_cast<const uint4*>(p.tc.win_img) + i);
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
        const float4* g_win = reinterpret_cast<const float4*>(smem + kOffWin) + lane;
        struct WarpItem {                                        // what an A warp needs of an item: its row, length, first block, frames
            const SampleT* row;
            int L, jb, nf;
            __device__ __forceinline__ void read(const ItemSlot& sl, const FrontendParams& p, int warp) {
                const bool more = sl.item >= 0;
                row = static_cast<const SampleT*>(p.wave) + (int64_t)(more ? sl.b : 0) * p.wave_stride;
                L = sl.L;
                jb = sl.t0 + 4 * warp - 1;
                nf = more ? min(4, max(sl.nfr - 4 * warp, 0)) : 0;
            }
        };
        WarpItem cur{nullptr, 0, 0, 0}, nxt{nullptr, 0, 0, 0};
        float x0[16], x1[16], x2[16], x3[16], x4[16];
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            const uint32_t rph = (it / kRing) & 1u;
            if (warp == 0 && lane == 0) ctl->a_progress = it;
            pipe_wait(&ctl->ring_full[rs], rph);
            if (ctl->slot[rs].item < 0) break;
            if (warp == 0) FE_TRACE(0, it);
            // This warp's frames 4 warp .. 4 warp + 3 need blocks b0..b4 (frame j = blocks j, j + 1).  The work is BLOCK by block:
            // block k is the upper half (n1 = 16..31) of frame k - 1 and the lower half of frame k, so only the block in hand
            // and the blocks in flight hold registers.  Five buffers with static roles (buffer k = block k of every item); after
            // block k the block TWO positions further in the warp's stream (b2, b3, b4, then b0, b1 of the NEXT item, which is
            // published two ahead) is requested: every load is in flight during two blocks of arithmetic, also across item
            // boundaries.  (A rolled three-buffer rotation - the first version - waited a fifth of the warp's time for the block
            // requested one frame earlier and spent more MOVs rotating buffers than FMULs on the window.)
            if (it == 0) {
                cur.read(ctl->slot[rs], p, warp);
                if (cur.nf > 0) {
                    load_block(cur.row, cur.L, cur.jb + 0, lane, x0);
                    load_block(cur.row, cur.L, cur.jb + 1, lane, x1);
                }
            }
            {                                                    // the item after this one (published two ahead)
                const int ns = (it + 1) % kRing;
                pipe_wait(&ctl->ring_full[ns], ((it + 1) / kRing) & 1u);
                nxt.read(ctl->slot[ns], p, warp);
            }
            if (warp == 0) FE_TRACE(1, it);
            const uint32_t lane_addr = tmem_base + kColA1 + ((uint32_t)(warp * 32) << 16);
            const uint32_t ph = (it & 1u) ^ 1u;
            // block K in buffer XK: closes frame K - 1 (sub-tile K - 1), opens frame K; then the request two positions ahead
#define SIR_FE_BLOCK(K, XK, REQ)                                                                                          \
            if (K >= 1) {                                                                                                \
                if (K - 1 < cur.nf) {                                                                                    \
                    store_half_frame<1>(XK, g_win, lane_addr + (K >= 1 ? K - 1 : 0) * 64u);                                           \
                    tmem_st_wait();                                                                                      \
                }                                                                                                        \
                tc_fence_before();                                                                                       \
                __syncwarp();                                                                                            \
                if (lane == 0) mbar_arrive(&ctl->a1_full[K >= 1 ? K - 1 : 0]);                                                      \
                if (warp == 0) FE_TRACE(3 + 2 * (K - 1), it);                                                            \
            }                                                                                                            \
            if (K <= 3) {                                                                                                \
                pipe_wait(&ctl->a1_empty[K], ph);            /* stage 1 of the previous item has read this sub-tile */   \
                if (warp == 0) FE_TRACE(2 + 2 * K, it);                                                                  \
                tc_fence_after();                                                                                        \
                if (K < cur.nf) store_half_frame<0>(XK, g_win, lane_addr + K * 64u);                                     \
            }                                                                                                            \
            REQ
            SIR_FE_BLOCK(0, x0, if (cur.nf >= 2) load_block(cur.row, cur.L, cur.jb + 2, lane, x2);)
            SIR_FE_BLOCK(1, x1, if (cur.nf >= 3) load_block(cur.row, cur.L, cur.jb + 3, lane, x3);)
            SIR_FE_BLOCK(2, x2, if (cur.nf >= 4) load_block(cur.row, cur.L, cur.jb + 4, lane, x4);)
            SIR_FE_BLOCK(3, x3, if (nxt.nf > 0) load_block(nxt.row, nxt.L, nxt.jb + 0, lane, x0);)
            SIR_FE_BLOCK(4, x4, if (nxt.nf > 0) load_block(nxt.row, nxt.L, nxt.jb + 1, lane, x1);)
#undef SIR_FE_BLOCK
            cur = nxt;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
        }
    } else if (warp == kWarpMma) {
        // =============================== MMA issue ===================================================================
        // Two streams of work, polled in turn: stage 1 of item i1 (sub-tile s1) and stage 2 of item i2 (row tile m2), i2
        // trailing i1.  All 32 lanes walk the loop (uniform), one elected lane issues.
        constexpr uint32_t id1 = make_idesc_tf32(128, 32), id2 = make_idesc_tf32(128, 64);
        const uint64_t b1_hi = make_kmajor_desc<128>(sbase + kOffB1), b1_lo = make_kmajor_desc<128>(sbase + kOffB1 + 4096u);
        uint32_t i1 = 0, s1 = 0, i2 = 0, m2 = 0, c2 = 0, idle = 0;
        bool have_slot = false, end1 = false;
        for (;;) {
            bool progressed = false;
            // stage 1 first: its MMAs are short and release the A warps; a stage 2 issued ahead of them (the tensor pipe
            // executes in order) would keep those warps waiting for ~1,000 cycles
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
                    FE_TRACE(10 + s1, i1);
                    if (elect_one_sync()) {
                        // A from tensor memory (one column per K element, 8 per instruction); K = n1 = 32 in four steps
                        const uint32_t d = tmem_base + kColD1 + s1 * 32u;
                        const uint32_t a_hi = tmem_base + kColA1 + s1 * 64u, a_lo = a_hi + 32u;
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) umma_tf32_ts(d, a_hi + 8u * kk, b1_hi + 2u * kk, id1, kk ? 1u : 0u);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) umma_tf32_ts(d, a_hi + 8u * kk, b1_lo + 2u * kk, id1, 1u);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) umma_tf32_ts(d, a_lo + 8u * kk, b1_hi + 2u * kk, id1, 1u);
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
            // stage 2 of a tile in SIX CHUNKS of four instructions (~200 cycles of the in-order tensor pipe), with a look at
            // stage 1 between them: issued in one go (24 instructions, > 1,100 cycles) it kept the A warps waiting for the
            // stage-1 MMAs queued behind it - a third of their time
            if (c2 > 0 || (mbar_test_wait(&ctl->a2_full[m2], i2 & 1u) && mbar_test_wait(&ctl->d2_empty[m2], (i2 & 1u) ^ 1u))) {
                if (c2 == 0) {
                    tc_fence_after();
                    FE_TRACE(14 + m2, i2);
                }
                if (elect_one_sync()) {
                    // K = (n2, re / im) = 64 = two 128-byte column blocks of four K = 8 steps; chunk = (pass, column block):
                    // hi.hi, hi.lo, lo.hi
                    const uint32_t d = tmem_base + kColD2 + m2 * 64u;
#pragma unroll 1
                    for (uint32_t c = c2; c < c2 + SIR_FE_CHUNK; ++c) {
                        const uint32_t pass = c >> 1, kb = c & 1u;
                        const uint64_t a = make_kmajor_desc<128>(sbase + (pass == 2 ? kOffA2Lo : kOffA2Hi) + kb * kA2KBlock + m2 * 16384u);
                        const uint64_t b = make_kmajor_desc<128>(sbase + kOffB2 + (pass == 1 ? 16384u : 0u) + kb * 8192u);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) umma_tf32(d, a + 2u * kk, b + 2u * kk, id2, (c | kk) ? 1u : 0u);
                    }
                    if (c2 + SIR_FE_CHUNK == 6) {
                        umma_commit(&ctl->d2_full[m2]);
                        umma_commit(&ctl->a2_empty[m2]);
                    }
                }
                __syncwarp();
                c2 += SIR_FE_CHUNK;
                if (c2 == 6) {
                    c2 = 0;
                    if (++m2 == 2) {
                        m2 = 0;
                        ++i2;
                    }
                }
                progressed = true;
            }
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
        const int q = warp & 3;                                  // TMEM lane quadrant: frames 4q .. 4q+3 of the item
        float2 tw[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tw[k] = __ldg(reinterpret_cast<const float2*>(p.tc.twiddle) + lane * 16 + k);
        // this lane's (re, im) pair = K elements 2 n2, 2 n2 + 1 of a stage-2 row: 128-byte column block n2 / 16, 16-byte chunk
        // (n2 % 16) / 2 XOR-ed with the row index inside its 8-row atom, byte (n2 & 1) * 8
        const uint32_t kb_off = (uint32_t)(lane >> 4) * kA2KBlock + (uint32_t)((lane & 1) << 3);
        uint32_t lane_off[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) lane_off[c] = kb_off + (uint32_t)(((((lane & 15) >> 1) ^ c)) << 4);
        for (uint32_t it = 0;; ++it) {
            const int rs = it % kRing;
            pipe_wait(&ctl->ring_full[rs], (it / kRing) & 1u);
            const long long item = ctl->slot[rs].item;
            const int nfr = ctl->slot[rs].nfr;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
            if (item < 0) break;
            for (int s = 0; s < 4; ++s) {
                // sub-tile s holds frame 4 q + s in quadrant q.  Its 17 stage-2 rows go to one of two SELF-CONTAINED 7-frame
                // tiles, in the order the sub-tiles complete: tile 0 = sub-tile 0 and quadrants 0..2 of sub-tile 1, tile 1 =
                // the rest (frames 14, 15 do not exist).  Position t in the tile: rows 16 t + k1 (k1 < 16) and 112 + t (k1 = 16).
                // The two tiles are a true double buffer: stage 2 of one runs while the other is written.
                const int f = 4 * q + s;
                const int tile = s == 0 ? 0 : (s == 1 ? (q == 3 ? 1 : 0) : 1);
                const int t = s == 0 ? q : (s == 1 ? (q == 3 ? 0 : 4 + q) : (s == 2 ? 1 + q : 4 + q));
                const bool valid = f < nfr;
                pipe_wait(&ctl->d1_full[s], it & 1u);
                if (q == 0) FE_TRACE(16 + s, it);
                tc_fence_after();
                float y[32];
                if (valid) tmem_ld_32x32(tmem_base + kColD1 + (uint32_t)s * 32u + ((uint32_t)(q * 32) << 16), y);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d1_empty[s]);
                if (f >= kTileFrames) continue;                  // (q = 3, s >= 2: no such frame, no stage-2 arrival expected)
                pipe_wait(&ctl->a2_empty[tile], (it & 1u) ^ 1u); // stage 2 of the previous item has read this tile
                if (q == 0 && s == 0) FE_TRACE(20, it);
                if (valid) {
                    const uint32_t tile_base = sbase + kOffA2Hi + (uint32_t)tile * 16384u;
                    const uint32_t hi_base = tile_base + (uint32_t)t * 2048u;                 // rows 16 t + k1
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
                        const uint32_t rb = __float_as_uint(re), ib = __float_as_uint(im);
                        const uint32_t addr = hi_base + (uint32_t)k1 * 128u + lane_off[k1 & 7];
                        st_shared_v2(addr, rb, ib);              // (the tensor core reads the 19 leading bits)
                        st_shared_v2(addr + (kOffA2Lo - kOffA2Hi), tf32_lo_bits(rb), tf32_lo_bits(ib));
                    }
                    {                                            // k1 = 16 (real Y): row 112 + t
                        const uint32_t rb = __float_as_uint(y[1] * tw[15].x), ib = __float_as_uint(y[1] * tw[15].y);
                        const uint32_t addr = tile_base + (uint32_t)(112 + t) * 128u + kb_off + (uint32_t)(((((lane & 15) >> 1) ^ (t & 7))) << 4);
                        st_shared_v2(addr, rb, ib);
                        st_shared_v2(addr + (kOffA2Lo - kOffA2Hi), tf32_lo_bits(rb), tf32_lo_bits(ib));
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->a2_full[tile]);
                if (q == 0 && s == 3) FE_TRACE(21, it);
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
                if (q == 0) FE_TRACE(22 + m, it);
                tc_fence_after();
                // row r of tile m: position t = r / 16 with k1 = r % 16 (r < 112), or t = r - 112 with k1 = 16 (r < 119);
                // tile 0: t < 4 is frame 4 t, else 4 (t - 4) + 1; tile 1: t = 0 is frame 13, t = 1..3 frames 4 (t - 1) + 2,
                // t = 4..6 frames 4 (t - 4) + 3 (the C warps' order of completion)
                const int r = 32 * q + lane;
                const int t = r < 112 ? r >> 4 : r - 112, k1 = r < 112 ? r & 15 : 16;
                const int f = m == 0 ? (t < 4 ? 4 * t : 4 * (t - 4) + 1) : (t == 0 ? 13 : (t < 4 ? 4 * (t - 1) + 2 : 4 * (t - 4) + 3));
                const bool valid = r < 119 && f < nfr;
                if (__any_sync(0xffffffffu, valid)) {
                    const uint32_t trow = tmem_base + kColD2 + (uint32_t)m * 64u + ((uint32_t)(q * 32) << 16);
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
                        float v[32];
                        tmem_ld_32x32(trow + 32 * j, v);
                        float* dst = Pf + base;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float re = v[2 * i], im = v[2 * i + 1];
                            const float pw = fmaf(re, re, im * im);
                            if (i < count) dst[step * i] = pw;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->d2_empty[m]);
            }
            const uint32_t buf = it & 1u;
            float* tile = s_tile + buf * kTileFloats;
            pipe_wait(&ctl->tile_empty[buf], ((it >> 1) & 1u) ^ 1u);   // the F warps have drained this tile
            if (q == 0) FE_TRACE(24, it);
            group_barrier();                                     // P complete
            if (q == 0) FE_TRACE(25, it);
            // ---- sparse mel taps: lane = (band, frame).  The 16 lanes of a half-warp read the same four taps (broadcast)
            // and their own frame's four bins (rows 532 floats apart: conflict-free LDS.128).  A warp takes QUADS of
            // neighbouring bands (4q..4q+3; the host pads their tap runs to one length), two bands per lane, the loads of
            // tap group i + 1 in flight while group i is accumulated.
            if (nfr > 0) {
                const int w = warp - kWarpD0, h = lane >> 4, f = lane & 15;
                const int n_quads = (p.n_mels + 3) >> 2;
                const float* Pf = s_P + (f < nfr ? f : 0) * kPStride;
                // quads in SNAKE order over the four warps (w, 7 - w, 8 + w, 15 - w, ...): the tap runs grow with the band index, and
                // with a plain round-robin the warp with the widest bands kept the other three waiting at the barrier below
                for (int j = 0; 4 * j < n_quads; ++j) {
                    const int pq = 4 * j + ((j & 1) ? 3 - w : w);
                    if (pq >= n_quads) continue;
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
                        float va = a0 + a1, vb = b0 + b1;
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
            if (q == 0) FE_TRACE(26, it);
            if (lane == 0) mbar_arrive(&ctl->ring_empty[rs]);
        }
    } else if (warp == kWarpPub) {
        // =============================== work items: draw tickets, publish the items to the other warps (one lane) =====
        // A warp of its own, because the draw (an atomic every CTA hits once per item), the lengths and the integer divisions
        // wait for memory: when A warp 0 did this between its frames, register pressure made ptxas spill buffers whose loads
        // were still pending, and every such spill waited out a memory latency - ~2,800 cycles per item on the pipeline's
        // critical path (timeline: tools/fe_trace.py); inside the MMA warp's polling loop it delayed every MMA instead.
        // i0 / L0 = the next item to publish and its length, r1 = the raw reply of the draw behind it
        // (32-bit item indices: one launch has fewer than 2^31 items; a ticket outside [0, total) means "no more work")
        const int total_i = (int)total_items;
        int i0 = -1;
        int L0 = 0;
        auto length_of = [&](int item) -> int {
            if (item < 0) return 0;
            const int b = item / p.groups_max;
            int L = p.lengths ? min(__ldg(p.lengths + b), p.n_samples) : p.n_samples;
            if (p.max_samples > 0) L = min(L, p.max_samples);
            return L;
        };
        // A draw is an atomic on one global word that every CTA hits once per item: its reply takes 1-2 us.  The RAW reply stays
        // in a register (r1) and is only decoded when the item is needed, one publish later - decoding it right away (the
        // first version) stalled the drawing lane, and with it this warp's frames and everything downstream, for ~2,800 cycles
        // per item (timeline: tools/fe_trace.py).
        auto draw_raw = [&]() -> unsigned long long { return atomicAdd(p.work_counter, 1ULL); };
        auto decode = [&](unsigned long long raw) -> int {
            const unsigned long long t = raw - p.work_base;      // wraps to huge if the base is ahead
            return t < (unsigned long long)total_i ? (int)t : -1;
        };
        unsigned long long r1 = 0;                               // raw reply of the draw behind i0
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
        if (lane == 0) {
            i0 = decode(draw_raw());
            r1 = draw_raw();
            L0 = length_of(i0);
            prefetch_item(i0, L0);
        }
        if (lane == 0) {
            for (uint32_t pub = 0;;) {
                for (uint32_t spin = 0; pub > ctl->a_progress + 3u; ++spin) {     // at most three items ahead of the A warps
                    __nanosleep(64);
                    if (spin > (1u << 26)) __trap();
                }
                const int ps = pub % kRing;
                pipe_wait(&ctl->ring_empty[ps], ((pub / kRing) & 1u) ^ 1u);
                ItemSlot& sl = ctl->slot[ps];
                bool end = false;
                for (;;) {                                       // skip tickets beyond an utterance's last group (ragged batches)
                    if (i0 < 0) {
                        sl.item = -1;
                        end = true;
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
                    i0 = decode(r1);
                    L0 = length_of(i0);
                    r1 = draw_raw();
                }
                mbar_arrive(&ctl->ring_full[ps]);
                ++pub;
                if (end) break;
                i0 = decode(r1);                                 // advance: the next item's samples are requested into L2, the
                L0 = length_of(i0);                              // draw after it is in flight until the next publish
                prefetch_item(i0, L0);
                r1 = draw_raw();
            }
        }
    } else {
        // =============================== F: tile -> global, statistics, counting atomic, finisher ====================
        const int ft = (warp - kWarpF0) * 32 + lane;             // 0..63
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
            if (warp == kWarpF0) FE_TRACE(27, it);

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
                        part.s1 = (double)red[0] + (double)red[1];
                        part.s2 = (double)red[8] + (double)red[9];
                        part.shift = shift;
                        part.vmax = fmaxf(red[16], red[17]);
                        part.n = nfr * p.n_mels;
                        part.pad = 0;
                        p.partials[item] = part;
                    }
                }
            }
            f_barrier();                                         // red consumed; slot data no longer needed
            if (warp == kWarpF0) FE_TRACE(28, it);
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
    // Hann window in the order the A warps read it: [n1 / 4][lane = n2][n1 % 4] (one 16-byte load per four taps)
    std::vector<float> win_img(1024);
    for (int i = 0; i < 1024; ++i) {
        const int n1 = i >> 5, l = i & 31;
        win_img[((n1 >> 2) * 32 + l) * 4 + (n1 & 3)] = ft.window[i];
    }
    const size_t o_b1 = 0, o_b2 = o_b1 + t.b1_img.size() * 4, o_tw = o_b2 + t.b2_img.size() * 4,
                 o_win = o_tw + t.twiddle.size() * 4, o_mw = o_win + win_img.size() * 4, o_ms = o_mw + fetc::kMelWeightCap * 4,
                 o_mc = o_ms + n_mels * 4, o_mo = o_mc + n_mels * 4, total = o_mo + n_mels * 4;
    int rc = buf.reserve(total);
    if (rc != SIR_OK) return rc;
    char* base = static_cast<char*>(buf.ptr);
    SIR_CUDA(cudaMemcpy(base + o_b1, t.b1_img.data(), t.b1_img.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_b2, t.b2_img.data(), t.b2_img.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_tw, t.twiddle.data(), t.twiddle.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_win, win_img.data(), win_img.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemset(base + o_mw, 0, fetc::kMelWeightCap * 4));
    SIR_CUDA(cudaMemcpy(base + o_mw, m_weight.data(), m_weight.size() * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_ms, m_start.data(), n_mels * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_mc, m_count.data(), n_mels * 4, cudaMemcpyHostToDevice));
    SIR_CUDA(cudaMemcpy(base + o_mo, m_offset.data(), n_mels * 4, cudaMemcpyHostToDevice));
    dev.b1_img = reinterpret_cast<const float*>(base + o_b1);
    dev.b2_img = reinterpret_cast<const float*>(base + o_b2);
    dev.twiddle = reinterpret_cast<const float*>(base + o_tw);
    dev.win_img = reinterpret_cast<const float*>(base + o_win);
    dev.mel_weight = reinterpret_cast<const float*>(base + o_mw);
    dev.mel_start = reinterpret_cast<const int32_t*>(base + o_ms);
    dev.mel_count = reinterpret_cast<const int32_t*>(base + o_mc);
    dev.mel_offset = reinterpret_cast<const int32_t*>(base + o_mo);
    return SIR_OK;
}

#ifdef SIR_FE_TRACE
extern "C" int sir_debug_fe_trace(long long* host, int count) {
    const size_t n = sizeof(long long) * fetc::kTraceCtas * fetc::kTraceItems * fetc::kTraceEvents;
    if ((size_t)count * sizeof(long long) < n) return -1;
    return cudaMemcpyFromSymbol(host, fetc::g_fe_trace, n) == cudaSuccess ? 0 : -2;
}
#endif

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
