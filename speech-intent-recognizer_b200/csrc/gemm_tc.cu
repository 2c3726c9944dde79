// tcgen05 contractions of the classifier: GRU input projections (plain GEMM) and conv2/conv3 (implicit GEMM),
// each as a 3-pass fp16 hi/lo split with fp32 accumulation in TMEM (tc_common.cuh explains the numerics).
//
// Replaces the dense contractions behind models/models.py:51-52 (conv2/conv3 + BN + ReLU + pool) and :60
// (nn.GRU's W_ih x for all time steps) - 88 % of the classifier's 400.6 MFLOP per utterance.
//
// One CTA computes one 128-row output tile; 6 warps:
//   warp 0   : TMA producer - per k-block, four bulk-tensor loads (A_hi, A_lo, B_hi, B_lo) into a
//              STAGES-deep shared-memory ring, completion on a `full` mbarrier (expect_tx);
//   warp 1   : allocates TMEM, one elected lane issues tcgen05.mma (128 x BLOCK_N x 16, kind::f16) - three
//              per 16-wide k-slice: hi.hi, hi.lo, lo.hi - and tcgen05.commit's the stage's `empty` mbarrier;
//   warps 2-5: epilogue - tcgen05.ld the accumulator (TMEM lane = tile row, column = output channel),
//              GEMM : + bias -> fp32 [M, N];
//              CONV : 2x2 max-pool by warp shuffles (the tile is 16 x 8 pixels, so a warp's 32 rows are a
//                     2 x 16 pixel patch and pooling partners are lanes ^1 and ^16), + BN shift, ReLU,
//                     split to fp16 hi/lo, channels-last store - directly the next contraction's A operand.
//
// Implicit GEMM: the activation is a 4-D TMA tensor (C, W, H, B); tap (kh, kw) of the 3x3 stencil is the box
// (C, 16, 8, 1) at (0, x0+kw-1, y0+kh-1, b) - the halo and the image border come for free from TMA's
// out-of-bounds zero fill, and the box lands in shared memory exactly as the K-major, swizzled UMMA operand
// (128 rows of C fp16).  K-blocks = 9 taps; weights are stored [tap][C_out][C_in].
#include <cstdio>
#include <cstring>

#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {
namespace tc {

constexpr int kTcThreads = 192;
enum { kModeGemm = 0, kModeConv = 1, kModeConvRaw = 2 };   // ConvRaw: no pool / shift / ReLU, fp32 NHWC output

struct TcParams {
    int num_kblocks;
    // GEMM
    int M, N;
    const float* bias;      // nullable
    float* C;
    float* C2;              // rows >= m_split go to C2 + (m - m_split) * N (two parameter blocks out of one launch)
    int m_split;
    const float* out_scale; // nullable: ONE device float the product is multiplied by (inverse of an operand's split scale)
    // CONV (input dims H x W, output pooled H/2 x W/2)
    int H, W, tiles_x;
    const float* shift;
    __half* out_hi;
    __half* out_lo;
    int out_whc;        // 0: [B][H2][W2][C]   1: [B][W2][H2][C]  (the GRU input order: time-major, then mel, then channel)
    int kc;             // CONV: k-blocks per tap (C_in / BLOCK_K)
    float* raw_out;     // CONV_RAW: fp32 [B][H][W][BLOCK_N]
};

template <int MODE, int BLOCK_N, int BLOCK_K, int STAGES>
struct TcLayout {
    static constexpr int kSwizzle = BLOCK_K * 2;
    static constexpr int kABytes = 128 * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kBarrierOffset = STAGES * kStageBytes;
    static constexpr int kSmemBytes = kBarrierOffset + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
    static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "tiles must keep 1024-byte alignment");
};

template <int MODE, int BLOCK_N, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_contract_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                       const TcParams p) {
    using L = TcLayout<MODE, BLOCK_N, BLOCK_K, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int m0 = 0, n0 = 0, img = 0, x0 = 0, y0 = 0;
    if constexpr (MODE == kModeGemm) {
        n0 = blockIdx.x * BLOCK_N;
        m0 = blockIdx.y * 128;
    } else {
        img = blockIdx.y;
        y0 = (blockIdx.x / p.tiles_x) * 8;
        x0 = (blockIdx.x % p.tiles_x) * 16;
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a_hi);
        prefetch_tmap(&tm_a_lo);
        prefetch_tmap(&tm_b_hi);
        prefetch_tmap(&tm_b_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BLOCK_N>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                uint8_t* st = smem + s * L::kStageBytes;
                mbar_arrive_expect_tx(&full[s], L::kStageBytes);
                if constexpr (MODE == kModeGemm) {
                    tma_load_2d(st, &tm_a_hi, &full[s], kb * BLOCK_K, m0);
                    tma_load_2d(st + L::kABytes, &tm_a_lo, &full[s], kb * BLOCK_K, m0);
                    tma_load_2d(st + 2 * L::kABytes, &tm_b_hi, &full[s], kb * BLOCK_K, n0);
                    tma_load_2d(st + 2 * L::kABytes + L::kBBytes, &tm_b_lo, &full[s], kb * BLOCK_K, n0);
                } else {
                    const int tap = kb / p.kc, c0 = (kb - tap * p.kc) * BLOCK_K;
                    const int kh = tap / 3, kw = tap - kh * 3;
                    tma_load_4d(st, &tm_a_hi, &full[s], c0, x0 + kw - 1, y0 + kh - 1, img);
                    tma_load_4d(st + L::kABytes, &tm_a_lo, &full[s], c0, x0 + kw - 1, y0 + kh - 1, img);
                    tma_load_2d(st + 2 * L::kABytes, &tm_b_hi, &full[s], c0, tap * BLOCK_N);
                    tma_load_2d(st + 2 * L::kABytes + L::kBBytes, &tm_b_lo, &full[s], c0, tap * BLOCK_N);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N);
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * L::kStageBytes);
                const uint64_t a_hi = make_kmajor_desc<L::kSwizzle>(base);
                const uint64_t a_lo = make_kmajor_desc<L::kSwizzle>(base + L::kABytes);
                const uint64_t b_hi = make_kmajor_desc<L::kSwizzle>(base + 2 * L::kABytes);
                const uint64_t b_lo = make_kmajor_desc<L::kSwizzle>(base + 2 * L::kABytes + L::kBBytes);
#pragma unroll
                for (int k = 0; k < BLOCK_K; k += 16) {
                    umma_f16(tmem_base, desc_advance_k(a_hi, k), desc_advance_k(b_hi, k), idesc, (kb | k) ? 1u : 0u);
                    umma_f16(tmem_base, desc_advance_k(a_hi, k), desc_advance_k(b_lo, k), idesc, 1u);
                    umma_f16(tmem_base, desc_advance_k(a_lo, k), desc_advance_k(b_hi, k), idesc, 1u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) ------------------------------------------
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if constexpr (MODE == kModeGemm) {
            const int m = m0 + q * 32 + lane;
            const float os = p.out_scale ? __ldg(p.out_scale) : 1.f;
            float* row = m < p.m_split ? p.C + (int64_t)m * p.N : p.C2 + (int64_t)(m - p.m_split) * p.N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                float v[32];
                tmem_ld_32x32(trow + c, v);
                if (m < p.M) {
                    float4* dst = reinterpret_cast<float4*>(row + n0 + c);
                    if (p.bias) {
                        const float4* bs = reinterpret_cast<const float4*>(p.bias + n0 + c);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b4 = __ldg(bs + i);
                            dst[i] = make_float4(fmaf(v[4 * i], os, b4.x), fmaf(v[4 * i + 1], os, b4.y),
                                                 fmaf(v[4 * i + 2], os, b4.z), fmaf(v[4 * i + 3], os, b4.w));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            dst[i] = make_float4(v[4 * i] * os, v[4 * i + 1] * os, v[4 * i + 2] * os, v[4 * i + 3] * os);
                    }
                }
            }
        } else if constexpr (MODE == kModeConvRaw) {
            // rows of this warp: dy = 2q + (lane >> 4), x = lane & 15; every thread stores its pixel's channels
            const int y = y0 + 2 * q + (lane >> 4), x = x0 + (lane & 15);
            const bool inside = y < p.H && x < p.W;
            float* dst_row = p.raw_out + (((int64_t)img * p.H + y) * p.W + x) * BLOCK_N;
            const float os = p.shift ? __ldg(p.shift) : 1.f;        // raw mode: shift[0] is the output scale (device scalar)
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                float v[32];
                tmem_ld_32x32(trow + c, v);
                if (inside) {
                    float4* dst = reinterpret_cast<float4*>(dst_row + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(v[4 * i] * os, v[4 * i + 1] * os, v[4 * i + 2] * os, v[4 * i + 3] * os);
                }
            }
        } else {
            // rows of this warp: dy = 2q + (lane >> 4), x = lane & 15  ->  one 2 x 16 pixel patch
            const int H2 = p.H / 2, W2 = p.W / 2;
            const int y2 = y0 / 2 + q, x2 = x0 / 2 + ((lane & 15) >> 1);
            const bool inside = y2 < H2 && x2 < W2;
            const int64_t pix = p.out_whc ? ((int64_t)img * W2 + x2) * H2 + y2 : ((int64_t)img * H2 + y2) * W2 + x2;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                float v[32], o[8];
                tmem_ld_32x32(trow + c, v);
                const int ch = c + pool2x2_split_channels(v, lane, o);
                if (inside)
                    shift_relu_split_store8(o, p.shift + ch, p.out_hi + pix * BLOCK_N + ch, p.out_lo + pix * BLOCK_N + ch);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BLOCK_N>(tmem_base);
    }
}

__global__ void split_f16_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                                 int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __half h, l;
        split_f16(in[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 tensor of `rank` dims (innermost first), box of the same rank, swizzle = inner box bytes (64 or 128).
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(SIR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4], gstride[3];
    cuuint32_t bdim[4], estride[4];
    uint64_t stride = 2;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estride[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    const uint32_t inner_bytes = box[0] * 2;
    const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_NONE) return fail(SIR_ERR_INVALID, "make_tmap: inner box must be 64 or 128 bytes");
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim,
                    estride, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SIR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SIR_OK;
}

template <int MODE, int BLOCK_N, int BLOCK_K, int STAGES>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                     const TcParams& p, dim3 grid, cudaStream_t st, const char* name) {
    using L = TcLayout<MODE, BLOCK_N, BLOCK_K, STAGES>;
    auto kern = tc_contract_kernel<MODE, BLOCK_N, BLOCK_K, STAGES>;
    SIR_SMEM_OPTIN(kern, L::kSmemBytes);
    {
        ProfScope ps(name, st);
        kern<<<grid, kTcThreads, L::kSmemBytes, st>>>(a_hi, a_lo, b_hi, b_lo, p);
    }
    SIR_CHECK_LAUNCH(name);
    return SIR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent GEMM for the GRU input projections C[M, N] = X[M, K] W[N, K]^T + bias (M = B*T rows, N = 1536,
// K = 1024 / 512), computed TRANSPOSED: the weight rows are the UMMA M dimension (128 per tile), the activation rows
// the UMMA N dimension in tiles of BN, and BN is picked per shape so that the tiles fill whole waves of the 148 SMs:
// with 6400 activation rows (256 utterances x 200 frames / 8: the bench's shape) the usual 128 x 128 / 128 x 256
// tilings give 600 / 300 tiles = 4.05 / 2.03 waves (measured: a third of the launch runs 4 % of the SMs) while
// 12 x ceil(6400 / 176) = 444 tiles are exactly 3; with 9472 rows (256 x 296 frames) BN = 176 would give 648 tiles =
// 4.4 waves (5 rounds of 176 columns) and BN = 256 exactly 3 rounds of 256 (- 13 % tensor time).  gp_pick_bn()
// minimises rounds x BN over the instantiated widths (the bench's shapes, 6400 and 12800 rows, both keep 176).
// One CTA per SM walks its tiles; two TMEM accumulators let the epilogue of tile i (bias add, transposed fp32 store:
// 32 lanes = 32 consecutive output columns) overlap the MMAs of tile i+1; barrier setup, TMEM allocation and pipeline
// fill are paid once per CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct GpLayout {
    static constexpr int kBK = 32, kBN = BN;                       // K-blocks of 32 = 64-byte swizzle rows
    static constexpr int kABytes = 128 * kBK * 2;                  // weight tile, one of (hi, lo): 8 KB
    static constexpr int kBBytes = kBN * kBK * 2;                  // activation tile: 10 - 16 KB
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // 36 - 48 KB
    static constexpr int kStages = 5 * kStageBytes <= 200 * 1024 ? 5 : 4;
    static constexpr int kOffBar = kStages * kStageBytes;
    static constexpr int kSmemBytes = kOffBar + (2 * kStages + 4) * 8 + (int)sizeof(TileRing) + 16 + 1024;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128; two accumulators of <= 256 columns");
    static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "operand tiles keep 1024-byte alignment");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct GpParams {
    TileTickets tickets;
    int M, N, num_kblocks, tiles_w, tiles_x, num_tiles;   // tiles_w: weight-row tiles (N / 128), tiles_x: activation-row tiles
    const float* bias;
    float* C;
    float* C2;
    int m_split;
    const float* out_scale;
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
    gemm_persistent_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                           const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                           const GpParams p) {
    using L = GpLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* empty = full + L::kStages;
    uint64_t* acc_full = empty + L::kStages;   // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]
    TileRing* ring = reinterpret_cast<TileRing*>(acc_empty + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + 1);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_w_hi);
        prefetch_tmap(&tm_w_lo);
        prefetch_tmap(&tm_x_hi);
        prefetch_tmap(&tm_x_lo);
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        ring_init(ring, 5);                            // MMA warp + 4 epilogue warps
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kAccStride = 256;       // second accumulator at column 256

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            TileProducer sched(ring, p.tickets, p.num_tiles);
            for (int tile = sched.pop(); tile >= 0; tile = sched.pop()) {
                const int w0 = (tile % p.tiles_w) * 128, x0 = (tile / p.tiles_w) * L::kBN;
                for (int kb = 0; kb < p.num_kblocks; ++kb, ++it) {
                    const int s = it % L::kStages;
                    mbar_wait(&empty[s], ((it / L::kStages) & 1u) ^ 1u);
                    uint8_t* st = smem + s * L::kStageBytes;
                    mbar_arrive_expect_tx(&full[s], L::kStageBytes);
                    tma_load_2d(st, &tm_w_hi, &full[s], kb * L::kBK, w0);
                    tma_load_2d(st + L::kABytes, &tm_w_lo, &full[s], kb * L::kBK, w0);
                    tma_load_2d(st + 2 * L::kABytes, &tm_x_hi, &full[s], kb * L::kBK, x0);
                    tma_load_2d(st + 2 * L::kABytes + L::kBBytes, &tm_x_lo, &full[s], kb * L::kBK, x0);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, L::kBN);
        const uint32_t sbase = smem_u32(smem);
        uint32_t it = 0, lt = 0, rt = 0;
        for (int tile = ring_next(ring, rt, lane); tile >= 0; tile = ring_next(ring, rt, lane), ++lt) {
            const uint32_t acc = lt & 1u;
            mbar_wait(&acc_empty[acc], ((lt >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            for (int kb = 0; kb < p.num_kblocks; ++kb, ++it) {
                const int s = it % L::kStages;
                mbar_wait(&full[s], (it / L::kStages) & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t base = sbase + s * L::kStageBytes;
                    const uint64_t a_hi = make_kmajor_desc<2 * L::kBK>(base);
                    const uint64_t a_lo = make_kmajor_desc<2 * L::kBK>(base + L::kABytes);
                    const uint64_t b_hi = make_kmajor_desc<2 * L::kBK>(base + 2 * L::kABytes);
                    const uint64_t b_lo = make_kmajor_desc<2 * L::kBK>(base + 2 * L::kABytes + L::kBBytes);
#pragma unroll
                    for (int k = 0; k < L::kBK; k += 16) {
                        umma_f16(d_tmem, desc_advance_k(a_hi, k), desc_advance_k(b_hi, k), idesc, (kb | k) ? 1u : 0u);
                        umma_f16(d_tmem, desc_advance_k(a_hi, k), desc_advance_k(b_lo, k), idesc, 1u);
                        umma_f16(d_tmem, desc_advance_k(a_lo, k), desc_advance_k(b_hi, k), idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == p.num_kblocks - 1) umma_commit(&acc_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // TMEM lane = weight row (output column n), TMEM column = activation row (output row m)
        const int q = warp & 3;
        uint32_t lt = 0, rt = 0;
        for (int tile = ring_next(ring, rt, lane); tile >= 0; tile = ring_next(ring, rt, lane), ++lt) {
            const uint32_t acc = lt & 1u;
            const int w0 = (tile % p.tiles_w) * 128, x0 = (tile / p.tiles_w) * L::kBN;
            mbar_wait(&acc_full[acc], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);
            const int n = w0 + q * 32 + lane;
            const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
            const float os = p.out_scale ? __ldg(p.out_scale) : 1.f;
#pragma unroll 1
            for (int c = 0; c < L::kBN; c += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(trow + c)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // 16 consecutive output rows m0 .. m0 + 15: ONE base pointer per chunk unless the chunk straddles the split between
                // the two output blocks (the per-element select + 64-bit multiply that served the split made this epilogue - and
                // with it the inference GEMMs - 20-50 % slower)
                const int m0 = x0 + c;
                if (m0 >= p.m_split || m0 + 16 <= p.m_split) {
                    float* __restrict__ dst = (m0 < p.m_split ? p.C + (int64_t)m0 * p.N : p.C2 + (int64_t)(m0 - p.m_split) * p.N) + n;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (m0 + i < p.M) dst[(int64_t)i * p.N] = fmaf(__uint_as_float(r[i]), os, bias);   // 32 lanes -> 128 contiguous bytes
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = m0 + i;
                        if (m < p.M) {
                            float* dst = m < p.m_split ? p.C + (int64_t)m * p.N : p.C2 + (int64_t)(m - p.m_split) * p.N;
                            dst[n] = fmaf(__uint_as_float(r[i]), os, bias);
                        }
                    }
                }
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
        tmem_dealloc<512>(tmem_base);
    }
}

template <int BN>
static int gp_launch(const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo, const float* bias, float* C,
                     int M, int N, int K, cudaStream_t st, const char* name, TicketSource* tickets, const GemmOutput* out) {
    using L = GpLayout<BN>;
    CUtensorMap tw_hi, tw_lo, tx_hi, tx_lo;
    const uint64_t xdims[2] = {(uint64_t)K, (uint64_t)M}, wdims[2] = {(uint64_t)K, (uint64_t)N};
    const uint32_t wbox[2] = {L::kBK, 128}, xbox[2] = {L::kBK, L::kBN};
    int rc;
    if ((rc = make_tmap(&tw_hi, w_hi, 2, wdims, wbox)) || (rc = make_tmap(&tw_lo, w_lo, 2, wdims, wbox)) ||
        (rc = make_tmap(&tx_hi, a_hi, 2, xdims, xbox)) || (rc = make_tmap(&tx_lo, a_lo, 2, xdims, xbox)))
        return rc;
    GpParams p{};
    p.M = M;
    p.N = N;
    p.num_kblocks = K / L::kBK;
    p.tiles_w = N / 128;
    p.tiles_x = (M + L::kBN - 1) / L::kBN;
    p.num_tiles = p.tiles_w * p.tiles_x;
    p.bias = bias;
    p.C = C;
    p.C2 = out && out->C2 ? out->C2 : C;
    p.m_split = out && out->C2 ? out->m_split : M;
    p.out_scale = out ? out->out_scale : nullptr;
    SIR_SMEM_OPTIN(gemm_persistent_kernel<BN>, L::kSmemBytes);
    const int num_sms = device_sm_count();
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    p.tickets = tickets ? tickets->first() : TileTickets{nullptr, 0};
    {
        ProfScope ps(name, st);
        gemm_persistent_kernel<BN><<<grid, kTcThreads, L::kSmemBytes, st>>>(tw_hi, tw_lo, tx_hi, tx_lo, p);
    }
    SIR_CHECK_LAUNCH(name);
    if (tickets) tickets->consumed(p.num_tiles, grid);
    return SIR_OK;
}

// Activation-tile width for M activation rows and N / 128 weight tiles on `sms` SMs: the width whose tiles need the
// least tensor time, rounds x BN (every round costs one tile of BN columns on the busiest SM); ties go to the wider tile
// (fewer tiles, fewer re-reads of the weight rows).
constexpr int kGpWidths[] = {160, 176, 208, 256};
static int gp_pick_bn(int M, int tiles_w, int sms) {
    int best = kGpWidths[0];
    int64_t best_cost = INT64_MAX;
    for (int bn : kGpWidths) {
        const int64_t tiles = (int64_t)tiles_w * ((M + bn - 1) / bn);
        const int64_t cost = ((tiles + sms - 1) / sms) * bn;
        if (cost <= best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

static bool gp_worthwhile(int M, int N) {   // enough tiles to be worth a persistent launch
    return (int64_t)((M + 175) / 176) * (N / 128) >= 32;
}

static int tc_gemm_nt_persistent(const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo, const float* bias,
                                 float* C, int M, int N, int K, cudaStream_t st, const char* name, TicketSource* tickets,
                                 const GemmOutput* out) {
    switch (gp_pick_bn(M, N / 128, device_sm_count())) {
        case 160: return gp_launch<160>(a_hi, a_lo, w_hi, w_lo, bias, C, M, N, K, st, name, tickets, out);
        case 176: return gp_launch<176>(a_hi, a_lo, w_hi, w_lo, bias, C, M, N, K, st, name, tickets, out);
        case 208: return gp_launch<208>(a_hi, a_lo, w_hi, w_lo, bias, C, M, N, K, st, name, tickets, out);
        default: return gp_launch<256>(a_hi, a_lo, w_hi, w_lo, bias, C, M, N, K, st, name, tickets, out);
    }
}

// C[M,N] = out_scale * A[M,K] W[N,K]^T + bias ; operands as fp16 hi/lo pairs.  N % 128 == 0, K % 64 == 0.
int tc_gemm_nt(const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo, const float* bias,
               float* C, int M, int N, int K, cudaStream_t st, const char* name, TicketSource* tickets, const GemmOutput* out) {
    if (N % 128 || K % 64) return fail(SIR_ERR_INVALID, "tc_gemm_nt: N %% 128 and K %% 64 required (N %d, K %d)", N, K);
    if (gp_worthwhile(M, N))
        return tc_gemm_nt_persistent(a_hi, a_lo, w_hi, w_lo, bias, C, M, N, K, st, name, tickets, out);
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const uint64_t adims[2] = {(uint64_t)K, (uint64_t)M}, bdims[2] = {(uint64_t)K, (uint64_t)N};
    const uint32_t abox[2] = {64, 128}, bbox[2] = {64, 128};
    int rc;
    if ((rc = make_tmap(&ta_hi, a_hi, 2, adims, abox)) || (rc = make_tmap(&ta_lo, a_lo, 2, adims, abox)) ||
        (rc = make_tmap(&tb_hi, w_hi, 2, bdims, bbox)) || (rc = make_tmap(&tb_lo, w_lo, 2, bdims, bbox)))
        return rc;
    TcParams p{};
    p.num_kblocks = K / 64;
    p.M = M;
    p.N = N;
    p.bias = bias;
    p.C = C;
    p.C2 = out && out->C2 ? out->C2 : C;
    p.m_split = out && out->C2 ? out->m_split : M;
    p.out_scale = out ? out->out_scale : nullptr;
    dim3 grid((unsigned)(N / 128), (unsigned)((M + 127) / 128));
    return launch_tc<kModeGemm, 128, 64, 3>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, st, name);
}

// 3x3 conv (s1, p1) on channels-last fp16 hi/lo activations [B,H,W,CIN]; weights [9][COUT][CIN] hi/lo.
//   raw_out == nullptr : + shift + ReLU + 2x2 max-pool -> [B,H/2,W/2,COUT] (or [B,W/2,H/2,COUT]) hi/lo  (eval forward)
//   raw_out != nullptr : plain convolution -> fp32 [B,H,W,COUT]  (training forward before BatchNorm, and the
//                        data gradient of a convolution = the same stencil with flipped, transposed weights);
//                        `shift`, if given, points at ONE device float the output is multiplied by (the inverse
//                        of the power-of-two scale applied to the gradient operand before its fp16 split)
template <int CIN, int COUT>
int tc_conv3x3(const __half* in_hi, const __half* in_lo, const __half* w_hi, const __half* w_lo, const float* shift,
               __half* out_hi, __half* out_lo, float* raw_out, int B, int H, int W, int out_whc, cudaStream_t st,
               const char* name) {
    constexpr int BK = CIN < 64 ? CIN : 64;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const uint64_t adims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint32_t abox[4] = {(uint32_t)BK, 16, 8, 1};
    const uint64_t bdims[2] = {(uint64_t)CIN, (uint64_t)(9 * COUT)};
    const uint32_t bbox[2] = {(uint32_t)BK, (uint32_t)COUT};
    int rc;
    if ((rc = make_tmap(&ta_hi, in_hi, 4, adims, abox)) || (rc = make_tmap(&ta_lo, in_lo, 4, adims, abox)) ||
        (rc = make_tmap(&tb_hi, w_hi, 2, bdims, bbox)) || (rc = make_tmap(&tb_lo, w_lo, 2, bdims, bbox)))
        return rc;
    TcParams p{};
    p.kc = CIN / BK;
    p.num_kblocks = 9 * p.kc;
    p.H = H;
    p.W = W;
    p.tiles_x = (W + 15) / 16;
    p.shift = shift;
    p.out_hi = out_hi;
    p.out_lo = out_lo;
    p.out_whc = out_whc;
    p.raw_out = raw_out;
    const int tiles_y = (H + 7) / 8;
    dim3 grid((unsigned)(p.tiles_x * tiles_y), (unsigned)B);
    if (raw_out) return launch_tc<kModeConvRaw, COUT, BK, (BK == 32 ? 4 : 3)>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, st, name);
    if constexpr (CIN <= 64 && COUT >= 64)
        return launch_tc<kModeConv, COUT, BK, (BK == 32 ? 4 : 3)>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, st, name);
    else
        return fail(SIR_ERR_UNSUPPORTED, "tc_conv3x3<%d,%d>: pooled epilogue not instantiated", CIN, COUT);
}

#define SIR_INST_CONV(CIN, COUT)                                                                                       \
    template int tc_conv3x3<CIN, COUT>(const __half*, const __half*, const __half*, const __half*, const float*,      \
                                       __half*, __half*, float*, int, int, int, int, cudaStream_t, const char*);
SIR_INST_CONV(32, 64)     // conv2 forward
SIR_INST_CONV(64, 128)    // conv3 forward
SIR_INST_CONV(128, 64)    // conv3 data gradient
SIR_INST_CONV(64, 32)     // conv2 data gradient
#undef SIR_INST_CONV

int split_f16_async(const float* in, __half* hi, __half* lo, int64_t n, cudaStream_t st) {
    if (n <= 0) return SIR_OK;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    split_f16_kernel<<<blocks, 256, 0, st>>>(in, hi, lo, n);
    SIR_CHECK_LAUNCH("split_f16_kernel");
    return SIR_OK;
}

}  // namespace tc
}  // namespace sir

// ---- C ABI: the split-precision tensor-core GEMM as a stand-alone operator --------------------------------------
using namespace sir;

extern "C" int sir_gemm_tile_width(int M, int N, int sms) {
    if (M < 1 || N < 128 || N % 128 || sms < 1 || !sir::tc::gp_worthwhile(M, N)) return 0;
    return sir::tc::gp_pick_bn(M, N / 128, sms);
}

extern "C" int sir_gemm_nt_split_f16(const float* d_a, const float* d_w, const float* d_bias, float* d_c, int M, int N,
                                     int K, void* stream) {
    if (!d_a || !d_w || !d_bias || !d_c || M < 1) return fail(SIR_ERR_INVALID, "sir_gemm_nt_split_f16: bad arguments");
    static DeviceBuffer scratch;
    const size_t na = (size_t)M * K, nw = (size_t)N * K;
    int rc = scratch.reserve((na + nw) * 2 * sizeof(__half) + 1024);
    if (rc != SIR_OK) return rc;
    __half* a_hi = (__half*)scratch.ptr;
    __half* a_lo = a_hi + na;
    __half* w_hi = a_lo + na;
    __half* w_lo = w_hi + nw;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = tc::split_f16_async(d_a, a_hi, a_lo, (int64_t)na, st))) return rc;
    if ((rc = tc::split_f16_async(d_w, w_hi, w_lo, (int64_t)nw, st))) return rc;
    return tc::tc_gemm_nt(a_hi, a_lo, w_hi, w_lo, d_bias, d_c, M, N, K, st, "gemm_nt_split_f16", nullptr, nullptr);
}

// 3x3 convolution (stride 1, zero padding 1, no bias) on channels-last fp32 tensors through the same implicit-GEMM
// kernel the classifier uses; exported so that tests can check every instantiation against torch's conv2d.
extern "C" int sir_conv3x3_nhwc_split_f16(const float* d_in, const float* d_w, float* d_out, int B, int H, int W, int cin,
                                          int cout, void* stream) {
    if (!d_in || !d_w || !d_out || B < 1 || H < 1 || W < 1) return fail(SIR_ERR_INVALID, "sir_conv3x3_nhwc_split_f16: bad arguments");
    static DeviceBuffer scratch;
    const size_t na = (size_t)B * H * W * cin, nw = (size_t)9 * cout * cin;
    const size_t na_pad = (na + 63) & ~(size_t)63, nw_pad = (nw + 63) & ~(size_t)63;
    int rc = scratch.reserve((2 * na_pad + 2 * nw_pad) * sizeof(__half) + 1024);
    if (rc != SIR_OK) return rc;
    __half* a_hi = (__half*)scratch.ptr;
    __half* a_lo = a_hi + na_pad;
    __half* w_hi = a_lo + na_pad;
    __half* w_lo = w_hi + nw_pad;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = tc::split_f16_async(d_in, a_hi, a_lo, (int64_t)na, st))) return rc;
    if ((rc = tc::split_f16_async(d_w, w_hi, w_lo, (int64_t)nw, st))) return rc;
    if (cin == 32 && cout == 64)
        return tc::tc_conv3x3<32, 64>(a_hi, a_lo, w_hi, w_lo, nullptr, nullptr, nullptr, d_out, B, H, W, 0, st, "conv3x3_32_64");
    if (cin == 64 && cout == 128)
        return tc::tc_conv3x3<64, 128>(a_hi, a_lo, w_hi, w_lo, nullptr, nullptr, nullptr, d_out, B, H, W, 0, st, "conv3x3_64_128");
    if (cin == 128 && cout == 64)
        return tc::tc_conv3x3<128, 64>(a_hi, a_lo, w_hi, w_lo, nullptr, nullptr, nullptr, d_out, B, H, W, 0, st, "conv3x3_128_64");
    if (cin == 64 && cout == 32)
        return tc::tc_conv3x3<64, 32>(a_hi, a_lo, w_hi, w_lo, nullptr, nullptr, nullptr, d_out, B, H, W, 0, st, "conv3x3_64_32");
    return fail(SIR_ERR_UNSUPPORTED, "sir_conv3x3_nhwc_split_f16: (C_in, C_out) = (%d, %d) is not instantiated", cin, cout);
}
