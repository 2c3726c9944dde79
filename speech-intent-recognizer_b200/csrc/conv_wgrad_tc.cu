// Weight gradient of the 3x3 convolutions (conv2, conv3) on tcgen05.
//
// The autograd of models/models.py:51-52 (nn.Conv2d inside conv_block) produces, per tap (kh, kw),
//     dW[co][ci][kh][kw] = sum over pixels of dz[pixel][co] * a[pixel shifted by the tap][ci]
// a GEMM whose reduction dimension is the PIXEL index.  Both tensors live channels-last ([B][H][W][C], fp16 hi/lo pairs
// that the data-gradient convolution and the forward pass already use), i.e. with the GEMM's M (co) and N (ci) contiguous
// and K (pixels) strided: the operands are fed to the tensor cores MN-major, exactly as TMA delivers a (C, 16, 4, 1) box -
// 64 pixel rows of C channels, 128-byte (C = 64) or 64-byte (C = 32) swizzle.  No transposed copy is made.
//
// grid (chunks, 3): CTA (c, kh) walks the 64-pixel tiles of chunk c and accumulates the three taps (kh, 0..2) in tensor
// memory: per tile TMA loads dz once and the activation tile three times (shifted by kw - 1; the halo and the zero padding
// of the convolution are TMA's out-of-bounds fill), then 3 taps x 4 k-steps x 3 split passes (hi.hi, hi.lo, lo.hi) MMAs of
// 128 x CIN x 16.  The accumulators are written ONCE per CTA: partial[chunk][tap][co][ci]; wgrad_reduce_kernel (train.cu)
// adds the chunks in a fixed order and undoes the power-of-two scale dz carries.
// COUT = 64 runs as M = 128 as well: the second 64-row block of the A descriptor points at whatever follows in shared
// memory, and accumulator rows 64..127 are never read.
#include <cstdint>

#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {
namespace tc {

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

constexpr int kWgThreads = 192;
constexpr int kWgPix = 64;            // pixels per tile = 4 image rows x 16 columns = the K block

template <int CIN, int COUT>
struct WgLayout {
    static constexpr int kDzBytes = kWgPix * COUT * 2;                 // one of (hi, lo); COUT = 128: two 64-channel boxes
    static constexpr int kABytes = kWgPix * CIN * 2;                   // one shifted activation tile, one of (hi, lo)
    static constexpr int kStageBytes = 2 * kDzBytes + 6 * kABytes;     // conv2: 40 KB, conv3: 80 KB
    static constexpr int kStages = CIN == 32 ? 4 : 2;
    static constexpr int kTmemCols = 3 * CIN <= 128 ? 128 : 256;
    static constexpr int kOffBar = kStages * kStageBytes;
    static constexpr int kSmemBytes = kOffBar + (2 * kStages + 1) * 8 + 16 + 1024;
    static_assert(kDzBytes % 1024 == 0 && kABytes % 1024 == 0, "operand tiles keep 1024-byte alignment");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// MN-major operand tile as TMA writes it: K rows (pixels) of SWIZZLE bytes (the contiguous M/N run), 8-row swizzle atoms
// every 8 * SWIZZLE bytes (stride byte offset); a second block of M/N elements `lbo_bytes` further (leading byte offset).
template <int SWIZZLE>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    static_assert(SWIZZLE == 64 || SWIZZLE == 128, "swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(((8u * SWIZZLE) >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SWIZZLE == 128 ? 2 : 4) << 61;
    return d;
}

struct WgParams {
    int tiles_x, tiles_y, tiles_per_img, num_tiles, tiles_per_chunk;
    float* partial;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kWgThreads, 1)
    conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dz_hi, const __grid_constant__ CUtensorMap tm_dz_lo,
                         const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                         const WgParams p) {
    using L = WgLayout<CIN, COUT>;
    constexpr int ASW = CIN * 2;                     // swizzle span of the activation operand: 64 or 128 bytes
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* empty = full + L::kStages;
    uint64_t* tmem_full = empty + L::kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int chunk = blockIdx.x, kh = blockIdx.y;
    const int t0 = chunk * p.tiles_per_chunk;
    const int t1 = t0 + p.tiles_per_chunk < p.num_tiles ? t0 + p.tiles_per_chunk : p.num_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_dz_hi);
        prefetch_tmap(&tm_dz_lo);
        prefetch_tmap(&tm_a_hi);
        prefetch_tmap(&tm_a_lo);
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<L::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = t0; t < t1; ++t, ++it) {
                const int img = t / p.tiles_per_img, r = t - img * p.tiles_per_img;
                const int y0 = (r / p.tiles_x) * 4, x0 = (r % p.tiles_x) * 16;
                const int s = it % L::kStages;
                mbar_wait(&empty[s], ((it / L::kStages) & 1u) ^ 1u);
                uint8_t* st = smem + s * L::kStageBytes;
                mbar_arrive_expect_tx(&full[s], L::kStageBytes);
#pragma unroll
                for (int j = 0; j < COUT / 64; ++j) {
                    tma_load_4d(st + j * 8192, &tm_dz_hi, &full[s], j * 64, x0, y0, img);
                    tma_load_4d(st + L::kDzBytes + j * 8192, &tm_dz_lo, &full[s], j * 64, x0, y0, img);
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    tma_load_4d(st + 2 * L::kDzBytes + (2 * kw) * L::kABytes, &tm_a_hi, &full[s], 0, x0 + kw - 1, y0 + kh - 1, img);
                    tma_load_4d(st + 2 * L::kDzBytes + (2 * kw + 1) * L::kABytes, &tm_a_lo, &full[s], 0, x0 + kw - 1, y0 + kh - 1,
                                img);
                }
            }
        }
    } else if (warp == 1) {
        // both operands MN-major: bits 15 (A) and 16 (B) of the instruction descriptor
        constexpr uint32_t idesc = make_idesc_f16(128, CIN) | (1u << 15) | (1u << 16);
        const uint32_t sbase = smem_u32(smem);
        uint32_t it = 0;
        for (int t = t0; t < t1; ++t, ++it) {
            const int s = it % L::kStages;
            mbar_wait(&full[s], (it / L::kStages) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t base = sbase + s * L::kStageBytes;
                const uint64_t dz_hi = make_mnmajor_desc<128>(base, 8192), dz_lo = make_mnmajor_desc<128>(base + L::kDzBytes, 8192);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t ab = base + 2 * L::kDzBytes + 2 * kw * L::kABytes;
                    const uint64_t a_hi = make_mnmajor_desc<ASW>(ab, L::kABytes), a_lo = make_mnmajor_desc<ASW>(ab + L::kABytes, L::kABytes);
                    const uint32_t d_tmem = tmem_base + kw * CIN;
#pragma unroll
                    for (int ks = 0; ks < kWgPix / 16; ++ks) {
                        // 16 pixel rows further: two 8-row atoms of each operand
                        const uint64_t da = (uint64_t)((ks * 16 * 128) >> 4), db = (uint64_t)((ks * 16 * ASW) >> 4);
                        umma_f16(d_tmem, dz_hi + da, a_hi + db, idesc, (it | (uint32_t)ks) ? 1u : 0u);
                        umma_f16(d_tmem, dz_hi + da, a_lo + db, idesc, 1u);
                        umma_f16(d_tmem, dz_lo + da, a_hi + db, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);
                if (t == t1 - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int co = q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if (q * 32 < COUT) {                          // warp-uniform: accumulator rows beyond COUT are padding
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                float* dst = p.partial + (((int64_t)chunk * 9 + kh * 3 + kw) * COUT + co) * CIN;
#pragma unroll 1
                for (int c = 0; c < CIN; c += 32) {
                    float v[32];
                    tmem_ld_32x32(trow + kw * CIN + c, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        reinterpret_cast<float4*>(dst + c)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<L::kTmemCols>(tmem_base);
    }
}

// partial[chunk][tap][COUT][CIN] for chunk < *chunks_out (<= max_chunks); every chunk owns at least one tile.
template <int CIN, int COUT>
int tc_conv_wgrad(const __half* dz_hi, const __half* dz_lo, const __half* a_hi, const __half* a_lo, float* partial, int B, int H,
                  int W, int max_chunks, int* chunks_out, cudaStream_t st, const char* name) {
    using L = WgLayout<CIN, COUT>;
    CUtensorMap tdz_hi, tdz_lo, ta_hi, ta_lo;
    const uint64_t zdims[4] = {(uint64_t)COUT, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t adims[4] = {(uint64_t)CIN, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint32_t zbox[4] = {64, 16, 4, 1}, abox[4] = {(uint32_t)CIN, 16, 4, 1};
    int rc;
    if ((rc = make_tmap(&tdz_hi, dz_hi, 4, zdims, zbox)) || (rc = make_tmap(&tdz_lo, dz_lo, 4, zdims, zbox)) ||
        (rc = make_tmap(&ta_hi, a_hi, 4, adims, abox)) || (rc = make_tmap(&ta_lo, a_lo, 4, adims, abox)))
        return rc;
    WgParams p{};
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 3) / 4;
    p.tiles_per_img = p.tiles_x * p.tiles_y;
    p.num_tiles = p.tiles_per_img * B;
    p.tiles_per_chunk = (p.num_tiles + max_chunks - 1) / max_chunks;
    const int chunks = (p.num_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.partial = partial;
    auto kern = conv_wgrad_tc_kernel<CIN, COUT>;
    SIR_SMEM_OPTIN(kern, L::kSmemBytes);
    {
        ProfScope ps(name, st);
        kern<<<dim3((unsigned)chunks, 3), kWgThreads, L::kSmemBytes, st>>>(tdz_hi, tdz_lo, ta_hi, ta_lo, p);
    }
    SIR_CHECK_LAUNCH(name);
    *chunks_out = chunks;
    return SIR_OK;
}

template int tc_conv_wgrad<32, 64>(const __half*, const __half*, const __half*, const __half*, float*, int, int, int, int, int*,
                                   cudaStream_t, const char*);
template int tc_conv_wgrad<64, 128>(const __half*, const __half*, const __half*, const __half*, float*, int, int, int, int, int*,
                                    cudaStream_t, const char*);

}  // namespace tc
}  // namespace sir
