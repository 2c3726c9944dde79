// Persistent implicit-GEMM 3x3 convolution + BN shift + ReLU + 2x2 max-pool for the layers whose weights fit in
// shared memory next to the activation ring (conv2: 32 -> 64 channels, 72 KB of fp16 hi/lo weights).
//
// Replaces models/models.py:51 (conv2 + bn2 + relu + pool) of the reference in eval mode; same arithmetic as
// tc_contract_kernel<CONV> in gemm_tc.cu (3-pass fp16 hi/lo split, fp32 accumulation in TMEM), different schedule:
//
//   * one CTA per SM walks over output tiles (16 x 8 pixels = 128 UMMA rows) round-robin;
//   * the nine weight taps are loaded ONCE per CTA and stay resident in shared memory;
//   * per tile only THREE activation loads are issued, one per horizontal tap offset kw: the TMA box
//     (C, 16, 10, 1) carries the vertical halo, and the three vertical taps read it through UMMA descriptors
//     advanced by 16 rows (row = y * 16 + x, so a vertical shift is a 1024-byte offset - swizzle-atom aligned);
//     L2 -> SM traffic per tile drops from 9 x (A + B) tiles to 3 x 1.25 A tiles (measured on the non-persistent
//     kernel: 1.6 GB per launch at batch 256, L2-bandwidth bound);
//   * two TMEM accumulators: the epilogue of tile i (pool by warp shuffles, shift, ReLU, fp16 split, channels-last
//     store) overlaps the MMAs of tile i+1.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane
// quadrant = warp % 4).  Every mbarrier wait is bounded (tc_common.cuh) - a protocol bug traps instead of hanging.
#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {
namespace tc {

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

constexpr int kCpThreads = 192;           // conv3 (stream kernel): TMA producer, MMA issuer, 4 epilogue warps
constexpr int kCp2Threads = 224;          // conv2 (persistent kernel): + a warp that collects the barriers of a tile

// Optional timeline of the MMA-issuing warp of CTA 0 (build with -DSIR_CONV_TRACE; tools/conv_trace.py reads it): clock64 per
// tile after the ring / accumulator waits and, per horizontal tap, after the stage wait and after its MMAs + commit are issued.
// Measured (conv2, 256 utterances): 2,880 cycles per tile = 3 x ~740 issuing (the thread is held at tcgen05.mma while the
// pipe's short queue is full) + ~100 per barrier wait although the phases are complete + 240 for the next tile: 660 cycles
// per tile with a draining tensor pipe.  Probing the next barriers between the first and the remaining MMAs of a batch
// made it slower (0.102 ms instead of 0.089: the queue is too short to cover a second elect / sync region).
#ifdef SIR_CONV_TRACE
__device__ long long g_conv_trace[64][8];
#define CONV_TRACE(T, E) do { if (blockIdx.x == 0 && (T) < 64u && lane == 0) g_conv_trace[T][E] = clock64(); } while (0)
#else
#define CONV_TRACE(T, E) do { } while (0)
#endif

template <int CIN, int COUT, int STAGES>
struct CpLayout {
    static constexpr int kRowBytes = CIN * 2;                       // one pixel's channels = one swizzle row (64 B)
    static constexpr int kTapBytes = COUT * kRowBytes;              // one weight tap, one of (hi, lo)
    static constexpr int kWBytes = 9 * 2 * kTapBytes;               // all taps, hi + lo
    static constexpr int kABytes = 160 * kRowBytes;                 // 16 x 10 pixel halo box, one of (hi, lo)
    static constexpr int kStageBytes = 2 * kABytes;
    static constexpr int kOffA = kWBytes;
    static constexpr int kOffBar = kOffA + STAGES * kStageBytes;
    static constexpr int kSmemBytes = kOffBar + (2 * STAGES + 5 + 2) * 8 + (int)sizeof(TileRing) + 32 + 1024;
    static_assert(kTapBytes % 1024 == 0 && kABytes % 1024 == 0, "operand tiles keep 1024-byte alignment");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct CpParams {
    TileTickets tickets;
    int B, H, W, tiles_x, tiles_y, num_tiles;
    const float* shift;
    __half* out_hi;
    __half* out_lo;
};

template <int CIN, int COUT, int STAGES>
__global__ void __launch_bounds__(kCp2Threads, 1)
    conv3x3_persistent_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                              const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                              const CpParams p) {
    using L = CpLayout<CIN, COUT, STAGES>;
    constexpr int kSwizzle = L::kRowBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* empty = full + STAGES;
    uint64_t* w_full = empty + STAGES;
    uint64_t* acc_full = w_full + 1;       // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2]
    uint64_t* go = acc_empty + 2;          // [2] "everything tile lt needs is there" (helper warp -> MMA warp)
    TileRing* ring = reinterpret_cast<TileRing*>(go + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + 1);
    volatile int* go_tile = reinterpret_cast<volatile int*>(tmem_slot + 2);      // [2] the tile behind go[lt & 1] (-1: none left)
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a_hi);
        prefetch_tmap(&tm_a_lo);
        prefetch_tmap(&tm_w_hi);
        prefetch_tmap(&tm_w_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(w_full, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);           // one arrival per epilogue warp
            mbar_init(&go[a], 1);
        }
        ring_init(ring, 5);                            // helper warp + 4 epilogue warps
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<4 * COUT>(tmem_slot);     // 2 accumulators x [hi.hi + lo.hi | hi.lo]
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            // resident weights: tap t at smem + t * 2 * kTapBytes (hi), + kTapBytes (lo)
            mbar_arrive_expect_tx(w_full, L::kWBytes);
            for (int t = 0; t < 9; ++t) {
                tma_load_2d(smem + t * 2 * L::kTapBytes, &tm_w_hi, w_full, 0, t * COUT);
                tma_load_2d(smem + t * 2 * L::kTapBytes + L::kTapBytes, &tm_w_lo, w_full, 0, t * COUT);
            }
            uint32_t it = 0;
            TileProducer sched(ring, p.tickets, p.num_tiles);
            for (int tile = sched.pop(); tile >= 0; tile = sched.pop()) {
                const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
                const int y0 = (r / p.tiles_x) * 8, x0 = (r % p.tiles_x) * 16;
                for (int kw = 0; kw < 3; ++kw, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1u) ^ 1u);
                    uint8_t* st = smem + L::kOffA + s * L::kStageBytes;
                    mbar_arrive_expect_tx(&full[s], L::kStageBytes);
                    tma_load_4d(st, &tm_a_hi, &full[s], 0, x0 + kw - 1, y0 - 1, img);
                    tma_load_4d(st + L::kABytes, &tm_a_lo, &full[s], 0, x0 + kw - 1, y0 - 1, img);
                }
            }
        }
    } else if (warp == 1) {
        // all 32 lanes walk the loop (uniform control flow and operands); one elected lane issues
        // Per k-slice: A_hi x [W_hi; W_lo] (the two weight halves of a tap are adjacent in shared memory: one N = 2 COUT
        // operand) and A_lo x W_hi (N = COUT) - the activation tile is fetched twice instead of three times per k-slice
        // (operand fetch from shared memory bounds this kernel).  Columns [0, COUT) collect hi.hi + lo.hi, columns
        // [COUT, 2 COUT) hi.lo; the epilogue adds them.
        constexpr uint32_t idesc = make_idesc_f16(128, COUT), idesc2 = make_idesc_f16(128, 2 * COUT);
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint32_t sbase = smem_u32(smem);
        // The issuing thread is held at a tcgen05.mma while the pipe's short queue is full, so whatever it does between two
        // batches runs while the queue drains: with a wait per activation stage, the accumulator wait and the tile ring that
        // was 660 of 2,880 cycles per tile (CONV_TRACE above).  Warp 6 now collects everything a tile needs and raises ONE
        // barrier (go) per tile; this warp waits for it and issues the tile's 36 MMAs in one go.
        uint32_t s = 0;
        for (uint32_t lt = 0;; ++lt) {
            const uint32_t acc = lt & 1u;
            CONV_TRACE(lt, 0);
            mbar_wait(&go[acc], (lt >> 1) & 1u);
            if (go_tile[acc] < 0) break;
            tc_fence_after();                                        // (the epilogue's reads of this accumulator are done)
            CONV_TRACE(lt, 1);
            const uint32_t d_tmem = tmem_base + acc * 2 * COUT;
            if (elect_one_sync()) {
                uint32_t sk = s;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t a_base = sbase + L::kOffA + sk * L::kStageBytes;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint32_t a_off = kh * 16 * L::kRowBytes;      // vertical tap = 16 rows further down the halo box
                        const uint64_t a_hi = make_kmajor_desc<kSwizzle>(a_base + a_off);
                        const uint64_t a_lo = make_kmajor_desc<kSwizzle>(a_base + L::kABytes + a_off);
                        const uint32_t w_base = sbase + (kh * 3 + kw) * 2 * L::kTapBytes;
                        const uint64_t b_hi = make_kmajor_desc<kSwizzle>(w_base);      // W_hi rows, W_lo rows right behind
#pragma unroll
                        for (int k = 0; k < CIN; k += 16) {
                            umma_f16(d_tmem, desc_advance_k(a_hi, k), desc_advance_k(b_hi, k), idesc2, (kw | kh | k) ? 1u : 0u);
                            umma_f16(d_tmem, desc_advance_k(a_lo, k), desc_advance_k(b_hi, k), idesc, 1u);
                        }
                    }
                    umma_commit(&empty[sk]);
                    sk = sk + 1 == STAGES ? 0 : sk + 1;
                }
                umma_commit(&acc_full[acc]);
            }
            __syncwarp();
            CONV_TRACE(lt, 7);
            s = (s + 3) % STAGES;
        }
    } else if (warp == 6) {
        // ---- helper: per tile, wait for its three activation stages and for the accumulator, then raise go ------------------
        uint32_t rt = 0, s = 0, sph = 0;
        for (uint32_t lt = 0;; ++lt) {
            const uint32_t acc = lt & 1u;
            const int tile = ring_next(ring, rt, lane);
            if (tile >= 0) {
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_wait(&full[s], sph);
                    if (++s == STAGES) {
                        s = 0;
                        sph ^= 1u;
                    }
                }
            }
            // the epilogue has drained this accumulator - which also says that the MMA warp has consumed go[acc] of tile lt - 2
            // (waited for even behind the last tile: a second arrival on a barrier whose previous phase nobody has looked at
            // yet would make that phase indistinguishable from the one before it)
            mbar_wait(&acc_empty[acc], ((lt >> 1) & 1u) ^ 1u);
            if (lane == 0) {
                go_tile[acc] = tile;
                mbar_arrive(&go[acc]);
            }
            __syncwarp();
            if (tile < 0) break;
        }
    } else {
        // ---- epilogue warps: TMEM lane quadrant q; rows of the quadrant = pixel rows 2q, 2q+1 of the tile ----------
        const int q = warp & 3;
        const int H2 = p.H / 2, W2 = p.W / 2;
        uint32_t lt = 0, rt = 0;
        for (int tile = ring_next(ring, rt, lane); tile >= 0; tile = ring_next(ring, rt, lane), ++lt) {
            const uint32_t acc = lt & 1u;
            const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
            const int y0 = (r / p.tiles_x) * 8, x0 = (r % p.tiles_x) * 16;
            mbar_wait(&acc_full[acc], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + acc * 2 * COUT + ((uint32_t)(q * 32) << 16);
            const int y2 = y0 / 2 + q, x2 = x0 / 2 + ((lane & 15) >> 1);
            const bool inside = y2 < H2 && x2 < W2;
            const int64_t pix = ((int64_t)img * H2 + y2) * W2 + x2;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 32) {
                float v[32], v2[32], o[8];
                tmem_ld_32x32(trow + c, v);
                tmem_ld_32x32(trow + COUT + c, v2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += v2[i];
                const int ch = c + pool2x2_split_channels(v, lane, o);
                if (inside) shift_relu_split_store8(o, p.shift + ch, p.out_hi + pix * COUT + ch, p.out_lo + pix * COUT + ch);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);             // this warp's quadrant of the accumulator is free
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<4 * COUT>(tmem_base);
    }
}

#ifdef SIR_CONV_TRACE
extern "C" int sir_debug_conv_trace(long long* host, int count) {
    if (count < 64 * 8) return -1;
    return cudaMemcpyFromSymbol(host, g_conv_trace, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -2;
}
#endif

// conv (3x3, s1, p1) + shift + ReLU + 2x2 max-pool, channels-last fp16 hi/lo in and out; weights [9][COUT][CIN].
template <int CIN, int COUT>
int tc_conv3x3_persistent(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
                          __half* out_hi, __half* out_lo, int B, int H, int W, int num_sms, cudaStream_t st, const char* name,
                          TicketSource* tickets) {
    constexpr int STAGES = 6;
    using L = CpLayout<CIN, COUT, STAGES>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    const uint64_t adims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint32_t abox[4] = {(uint32_t)CIN, 16, 10, 1};
    const uint64_t wdims[2] = {(uint64_t)CIN, (uint64_t)(9 * COUT)};
    const uint32_t wbox[2] = {(uint32_t)CIN, (uint32_t)COUT};
    int rc;
    if ((rc = make_tmap(&ta_hi, in_hi, 4, adims, abox)) || (rc = make_tmap(&ta_lo, in_lo, 4, adims, abox)) ||
        (rc = make_tmap(&tw_hi, w_hi, 2, wdims, wbox)) || (rc = make_tmap(&tw_lo, w_lo, 2, wdims, wbox)))
        return rc;
    CpParams p{};
    p.B = B;
    p.H = H;
    p.W = W;
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 7) / 8;
    p.num_tiles = B * p.tiles_x * p.tiles_y;
    p.shift = shift;
    p.out_hi = out_hi;
    p.out_lo = out_lo;
    auto kern = conv3x3_persistent_kernel<CIN, COUT, STAGES>;
    SIR_SMEM_OPTIN(kern, L::kSmemBytes);
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    p.tickets = tickets ? tickets->first() : TileTickets{nullptr, 0};
    {
        ProfScope ps(name, st);
        kern<<<grid, kCp2Threads, L::kSmemBytes, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
    }
    SIR_CHECK_LAUNCH(name);
    if (tickets) tickets->consumed(p.num_tiles, grid);
    return SIR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Streamed-weight variant for conv3 (64 -> 128 channels: 288 KB of fp16 hi/lo weights do not fit in shared memory).
// Same persistent schedule, but the tile is 8 x 16 pixels (W = 50 -> 7 tiles of 8 instead of 4 of 16: 12 % fewer
// tiles; a vertical tap is still a swizzle-atom aligned shift, 8 rows x 128 B = 1024 B), activations come through a
// 2-deep ring of per-kw halo boxes (C, 8, 18, 1) and the weight taps through a 4-deep ring (one tap = 32 KB hi+lo).
// The epilogue can write the GRU-input order [B][W/2][H/2][C] (models/models.py:55-57 folded into the store).
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct CsLayout {
    static constexpr int kASt = 2, kBSt = 4;
    static constexpr int kRowBytes = CIN * 2;                       // 128 B
    static constexpr int kTapBytes = COUT * kRowBytes;              // one weight tap, one of (hi, lo): 16 KB
    static constexpr int kABytes = 144 * kRowBytes;                 // 8 x 18 pixel halo box, one of (hi, lo): 18 KB
    static constexpr int kOffB = kASt * 2 * kABytes;
    static constexpr int kOffBar = kOffB + kBSt * 2 * kTapBytes;
    static constexpr int kSmemBytes = kOffBar + (2 * kASt + 2 * kBSt + 4) * 8 + (int)sizeof(TileRing) + 16 + 1024;
    static_assert(kRowBytes == 128, "built for 64 input channels (128-byte swizzle rows)");
    static_assert(kTapBytes % 1024 == 0 && kABytes % 1024 == 0, "operand tiles keep 1024-byte alignment");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct CsParams {
    TileTickets tickets;
    int B, H, W, tiles_x, tiles_y, num_tiles, out_whc;
    const float* shift;
    __half* out_hi;
    __half* out_lo;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kCpThreads, 1)
    conv3x3_stream_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                          const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                          const CsParams p) {
    using L = CsLayout<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* a_empty = a_full + L::kASt;
    uint64_t* b_full = a_empty + L::kASt;
    uint64_t* b_empty = b_full + L::kBSt;
    uint64_t* acc_full = b_empty + L::kBSt;   // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    TileRing* ring = reinterpret_cast<TileRing*>(acc_empty + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + 1);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a_hi);
        prefetch_tmap(&tm_a_lo);
        prefetch_tmap(&tm_w_hi);
        prefetch_tmap(&tm_w_lo);
        for (int s = 0; s < L::kASt; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < L::kBSt; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        ring_init(ring, 5);                            // MMA warp + 4 epilogue warps
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<4 * COUT>(tmem_slot);     // 2 accumulators x [hi.hi + lo.hi | hi.lo]
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;
            TileProducer sched(ring, p.tickets, p.num_tiles);
            for (int tile = sched.pop(); tile >= 0; tile = sched.pop()) {
                const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
                const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
                for (int kw = 0; kw < 3; ++kw, ++ia) {
                    const int sa = ia % L::kASt;
                    mbar_wait(&a_empty[sa], ((ia / L::kASt) & 1u) ^ 1u);
                    uint8_t* sta = smem + sa * 2 * L::kABytes;
                    mbar_arrive_expect_tx(&a_full[sa], 2 * L::kABytes);
                    tma_load_4d(sta, &tm_a_hi, &a_full[sa], 0, x0 + kw - 1, y0 - 1, img);
                    tma_load_4d(sta + L::kABytes, &tm_a_lo, &a_full[sa], 0, x0 + kw - 1, y0 - 1, img);
                    for (int kh = 0; kh < 3; ++kh, ++ib) {
                        const int sb = ib % L::kBSt;
                        mbar_wait(&b_empty[sb], ((ib / L::kBSt) & 1u) ^ 1u);
                        uint8_t* stb = smem + L::kOffB + sb * 2 * L::kTapBytes;
                        mbar_arrive_expect_tx(&b_full[sb], 2 * L::kTapBytes);
                        tma_load_2d(stb, &tm_w_hi, &b_full[sb], 0, (kh * 3 + kw) * COUT);
                        tma_load_2d(stb + L::kTapBytes, &tm_w_lo, &b_full[sb], 0, (kh * 3 + kw) * COUT);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // A_hi x [W_hi; W_lo] (N = 2 COUT, the halves of a tap are adjacent in the ring slot) + A_lo x W_hi, as in conv2
        constexpr uint32_t idesc = make_idesc_f16(128, COUT), idesc2 = make_idesc_f16(128, 2 * COUT);
        const uint32_t sbase = smem_u32(smem);
        uint32_t ia = 0, ib = 0, lt = 0, rt = 0;
        for (int tile = ring_next(ring, rt, lane); tile >= 0; tile = ring_next(ring, rt, lane), ++lt) {
            const uint32_t acc = lt & 1u;
            mbar_wait(&acc_empty[acc], ((lt >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 2 * COUT;
            for (int kw = 0; kw < 3; ++kw, ++ia) {
                const int sa = ia % L::kASt;
                mbar_wait(&a_full[sa], (ia / L::kASt) & 1u);
                tc_fence_after();
                const uint32_t a_base = sbase + sa * 2 * L::kABytes;
                for (int kh = 0; kh < 3; ++kh, ++ib) {
                    const int sb = ib % L::kBSt;
                    mbar_wait(&b_full[sb], (ib / L::kBSt) & 1u);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a_off = kh * 8 * L::kRowBytes;      // vertical tap = 8 rows (one 1024-byte atom) further down
                        const uint64_t a_hi = make_kmajor_desc<128>(a_base + a_off);
                        const uint64_t a_lo = make_kmajor_desc<128>(a_base + L::kABytes + a_off);
                        const uint32_t w_base = sbase + L::kOffB + sb * 2 * L::kTapBytes;
                        const uint64_t b_hi = make_kmajor_desc<128>(w_base);          // W_hi rows, W_lo rows right behind
#pragma unroll
                        for (int k = 0; k < CIN; k += 16) {
                            umma_f16(d_tmem, desc_advance_k(a_hi, k), desc_advance_k(b_hi, k), idesc2, (kw | kh | k) ? 1u : 0u);
                            umma_f16(d_tmem, desc_advance_k(a_lo, k), desc_advance_k(b_hi, k), idesc, 1u);
                        }
                        umma_commit(&b_empty[sb]);
                        if (kh == 2) umma_commit(&a_empty[sa]);
                        if (kh == 2 && kw == 2) umma_commit(&acc_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // epilogue: quadrant q holds tile rows y = 4q .. 4q+3 (lane = dy * 8 + x); pooling partners are lanes ^1 and ^8
        const int q = warp & 3;
        const int H2 = p.H / 2, W2 = p.W / 2;
        uint32_t lt = 0, rt = 0;
        for (int tile = ring_next(ring, rt, lane); tile >= 0; tile = ring_next(ring, rt, lane), ++lt) {
            const uint32_t acc = lt & 1u;
            const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
            const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
            mbar_wait(&acc_full[acc], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + acc * 2 * COUT + ((uint32_t)(q * 32) << 16);
            const int y2 = y0 / 2 + 2 * q + (lane >> 4), x2 = x0 / 2 + ((lane & 7) >> 1);
            const bool inside = y2 < H2 && x2 < W2;
            const int64_t pix = p.out_whc ? ((int64_t)img * W2 + x2) * H2 + y2 : ((int64_t)img * H2 + y2) * W2 + x2;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 32) {
                float v[32], v2[32], o[8];
                tmem_ld_32x32(trow + c, v);
                tmem_ld_32x32(trow + COUT + c, v2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += v2[i];
                const int ch = c + pool2x2_split_channels<8>(v, lane, o);
                if (inside) shift_relu_split_store8(o, p.shift + ch, p.out_hi + pix * COUT + ch, p.out_lo + pix * COUT + ch);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<4 * COUT>(tmem_base);
    }
}

template <int CIN, int COUT>
int tc_conv3x3_stream(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
                      __half* out_hi, __half* out_lo, int B, int H, int W, int out_whc, int num_sms, cudaStream_t st,
                      const char* name, TicketSource* tickets) {
    using L = CsLayout<CIN, COUT>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    const uint64_t adims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint32_t abox[4] = {(uint32_t)CIN, 8, 18, 1};
    const uint64_t wdims[2] = {(uint64_t)CIN, (uint64_t)(9 * COUT)};
    const uint32_t wbox[2] = {(uint32_t)CIN, (uint32_t)COUT};
    int rc;
    if ((rc = make_tmap(&ta_hi, in_hi, 4, adims, abox)) || (rc = make_tmap(&ta_lo, in_lo, 4, adims, abox)) ||
        (rc = make_tmap(&tw_hi, w_hi, 2, wdims, wbox)) || (rc = make_tmap(&tw_lo, w_lo, 2, wdims, wbox)))
        return rc;
    CsParams p{};
    p.B = B;
    p.H = H;
    p.W = W;
    p.tiles_x = (W + 7) / 8;
    p.tiles_y = (H + 15) / 16;
    p.num_tiles = B * p.tiles_x * p.tiles_y;
    p.out_whc = out_whc;
    p.shift = shift;
    p.out_hi = out_hi;
    p.out_lo = out_lo;
    auto kern = conv3x3_stream_kernel<CIN, COUT>;
    SIR_SMEM_OPTIN(kern, L::kSmemBytes);
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    p.tickets = tickets ? tickets->first() : TileTickets{nullptr, 0};
    {
        ProfScope ps(name, st);
        kern<<<grid, kCpThreads, L::kSmemBytes, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
    }
    SIR_CHECK_LAUNCH(name);
    if (tickets) tickets->consumed(p.num_tiles, grid);
    return SIR_OK;
}

template int tc_conv3x3_stream<64, 128>(const __half*, const __half*, const __half*, const __half*, const float*, __half*,
                                        __half*, int, int, int, int, int, cudaStream_t, const char*, TicketSource*);

template int tc_conv3x3_persistent<32, 64>(const __half*, const __half*, const __half*, const __half*, const float*, __half*,
                                           __half*, int, int, int, int, cudaStream_t, const char*, TicketSource*);

}  // namespace tc
}  // namespace sir
