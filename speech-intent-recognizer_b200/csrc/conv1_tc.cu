// conv1 (C_in = 1, 3x3, s1 p1, no bias) + folded BatchNorm + ReLU + 2x2 max-pool on the sm_100a tensor cores.
//
// Replaces models/models.py:50 of the reference (conv1/bn1/relu/pool) and round 1's fp32 CUDA-core kernel
// (one thread per pooled pixel, 1,152 FMAs each: FMA-pipe bound at 0.070 ms for 256 utterances).  With one input channel
// the convolution is an implicit GEMM with K = 9: every thread gathers the 3x3 neighbourhood of ONE output pixel, splits
// it into fp16 (hi, lo) halves and writes one K = 16 row of the A operand (9 taps + 7 zeros); three N = 32 MMAs
//     D[128 pixels x 32 channels] = A_hi W_hi + A_hi W_lo + A_lo W_hi
// do the 288 multiply-adds per pixel, and the epilogue is conv2's: a warp's TMEM lane quadrant is a 2 x 16 pixel patch,
// pooled with shuffles while the lanes split the channels, + shift, ReLU, fp16 (hi, lo) split, one 16-byte store each.
//
// Structure (as the frontend): a CTA is four independent warpgroups; a warpgroup takes 8 x 16-pixel tiles in a static
// round-robin (no tickets: the launch is safe inside a CUDA graph) through build -> barrier -> MMA -> mbarrier -> epilogue;
// two CTAs per SM, so eight tiles are in different phases on every SM.
#include "model.cuh"
#include "tc_common.cuh"

namespace sir {
namespace tc {

constexpr int kC1Wgs = 4;
constexpr int kC1Threads = kC1Wgs * 128;
constexpr uint32_t kC1OffW = kC1Wgs * 16384;                 // weight operand image: 64 rows x 128 B (rows 0..31 hi, 32..63 lo)
constexpr uint32_t kC1OffBar = kC1OffW + 8192;
constexpr uint32_t kC1SmemBytes = kC1OffBar + 64 + 1024;     // + slack for the 1024-byte alignment

__device__ __forceinline__ void c1_wg_barrier(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void c1_mma_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        __nanosleep(100);
        if (spin > (1u << 21)) __trap();                     // a protocol bug must surface as an error, never as a hang
    }
}
__device__ __forceinline__ void c1_st_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void c1_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kC1Threads, 2) conv1_tc_kernel(const float* __restrict__ feat,     // [B, H, W] fp32
                                                                 const float* __restrict__ w1,       // [32][9] BN-folded
                                                                 const float* __restrict__ shift1,   // [32]
                                                                 __half* __restrict__ out_hi,        // [B, H/2, W/2, 32]
                                                                 __half* __restrict__ out_lo, int B, int H, int W) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_idx();
    const int g = warp >> 2, q = warp & 3, wt = tid & 127;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kC1OffBar) + g;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kC1OffBar + 32);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_addr = sbase + (uint32_t)g * 16384u;

    if (wt == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<128>(s_tmem);
    // weight operand: row n < 32 = hi halves of channel n's 9 taps at K = 0..8 (zeros up to K = 15), row 32 + n = the lo halves
    for (int i = tid; i < 2048; i += kC1Threads) reinterpret_cast<uint32_t*>(smem + kC1OffW)[i] = 0u;
    __syncthreads();
    if (tid < 32) {
        const int n = tid;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = 2 * e < 9 ? __ldg(w1 + n * 9 + 2 * e) : 0.f, b = 2 * e + 1 < 9 ? __ldg(w1 + n * 9 + 2 * e + 1) : 0.f;
            c1_split_pair(a, b, hi[e], lo[e]);
        }
        const uint32_t r_hi = sbase + kC1OffW + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
        const uint32_t r_lo = r_hi + 4096u;                  // row 32 + n: four 8-row atoms further
        const int sw = n & 7;
        c1_st_v4(r_hi + (uint32_t)((0 ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
        c1_st_v4(r_hi + (uint32_t)((1 ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
        c1_st_v4(r_lo + (uint32_t)((0 ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
        c1_st_v4(r_lo + (uint32_t)((1 ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_wg = *s_tmem + (uint32_t)g * 32u;
    const uint32_t tmem_row = tmem_wg + ((uint32_t)(q * 32) << 16);

    const int H2 = H / 2, W2 = W / 2;
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8, tiles_img = tiles_x * tiles_y;
    const int num_tiles = B * tiles_img;
    constexpr uint32_t idesc = make_idesc_f16(128, 32);
    const uint64_t w_hi = make_kmajor_desc<128>(sbase + kC1OffW), w_lo = make_kmajor_desc<128>(sbase + kC1OffW + 4096u);
    const uint64_t a_desc = make_kmajor_desc<128>(a_addr);
    const int py = wt >> 4, px = wt & 15;                    // this thread's pixel inside the 8 x 16 tile = row wt of the operand
    const uint32_t row_addr = a_addr + (uint32_t)((wt >> 3) * 1024 + (wt & 7) * 128);
    const int sw = wt & 7;

    // tile coordinates are advanced incrementally (three integer divisions per tile were a tenth of the kernel)
    const int stride = gridDim.x * kC1Wgs;
    const int d_img = stride / tiles_img, d_r = stride - d_img * tiles_img, d_ty = d_r / tiles_x, d_tx = d_r - d_ty * tiles_x;
    int tile = blockIdx.x * kC1Wgs + g;
    int img = tile / tiles_img, ty = (tile - img * tiles_img) / tiles_x, tx = (tile - img * tiles_img) - ty * tiles_x;
    uint32_t it = 0;
    for (; tile < num_tiles; tile += stride, ++it) {
        const int y0 = ty * 8, x0 = tx * 16;
        // ---- build: the 3x3 neighbourhood of pixel (y0 + py, x0 + px), zero outside the image ----------------------------
        {
            const float* __restrict__ src = feat + (int64_t)img * H * W;
            const int gy = y0 + py, gx = x0 + px;
            float v[10];
            if (y0 > 0 && y0 + 8 < H && x0 > 0 && x0 + 16 < W) {          // interior tile (warpgroup-uniform): no bounds checks
                const float* __restrict__ c = src + (int64_t)(gy - 1) * W + (gx - 1);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) v[kh * 3 + kw] = __ldg(c + kh * W + kw);
            } else {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int yy = gy + kh - 1, xx = gx + kw - 1;
                        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
                        v[kh * 3 + kw] = in ? __ldg(src + (int64_t)(in ? yy : 0) * W + (in ? xx : 0)) : 0.f;
                    }
            }
            v[9] = 0.f;
            uint32_t hi[5], lo[5];
#pragma unroll
            for (int e = 0; e < 5; ++e) c1_split_pair(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
            c1_st_v4(row_addr + (uint32_t)((0 ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);      // K 0..7   (hi)
            c1_st_v4(row_addr + (uint32_t)((1 ^ sw) << 4), hi[4], 0u, 0u, 0u);               // K 8..15  (hi)
            c1_st_v4(row_addr + (uint32_t)((2 ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);      // K 16..23 (lo)
            c1_st_v4(row_addr + (uint32_t)((3 ^ sw) << 4), lo[4], 0u, 0u, 0u);               // K 24..31 (lo)
        }
        fence_proxy_async();                                 // generic-proxy stores -> visible to the tensor core
        tc_fence_before();                                   // (the previous tile's accumulator reads are done)
        c1_wg_barrier(g);
        if (q == 0) {
            tc_fence_after();
            if (elect_one_sync()) {
                umma_f16(tmem_wg, a_desc, w_hi, idesc, 0u);                              // hi . W_hi
                umma_f16(tmem_wg, a_desc, w_lo, idesc, 1u);                              // hi . W_lo
                umma_f16(tmem_wg, desc_advance_k(a_desc, 16), w_hi, idesc, 1u);          // lo . W_hi
                umma_commit(bar);
            }
            __syncwarp();
        }
        c1_mma_wait(bar, it & 1u);
        tc_fence_after();
        // ---- epilogue: the quadrant's 2 x 16 pixels -> pooled 1 x 8, 8 channels per lane ----------------------------------
        {
            float v[32], o[8];
            tmem_ld_32x32(tmem_row, v);
            const int ch = pool2x2_split_channels(v, lane, o);
            const int y2 = y0 / 2 + q, x2 = x0 / 2 + ((lane & 15) >> 1);
            if (y2 < H2 && x2 < W2) {
                const int64_t pix = ((int64_t)img * H2 + y2) * W2 + x2;
                shift_relu_split_store8(o, shift1 + ch, out_hi + pix * 32 + ch, out_lo + pix * 32 + ch);
            }
        }
        tx += d_tx;
        if (tx >= tiles_x) {
            tx -= tiles_x;
            ++ty;
        }
        ty += d_ty;
        if (ty >= tiles_y) {
            ty -= tiles_y;
            ++img;
        }
        img += d_img;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<128>(*s_tmem);
    }
}

int conv1_tc(const float* feat, const float* w1, const float* shift1, __half* out_hi, __half* out_lo, int B, int H, int W,
             int num_sms, cudaStream_t st) {
    if (B < 1) return SIR_OK;
    if ((H & 1) || (W & 1)) return fail(SIR_ERR_INVALID, "conv1: H and W must be even (got %d x %d)", H, W);
    SIR_SMEM_OPTIN(conv1_tc_kernel, kC1SmemBytes);
    const int64_t tiles = (int64_t)B * ((W + 15) / 16) * ((H + 7) / 8);
    const int64_t ctas = (tiles + kC1Wgs - 1) / kC1Wgs;
    const int grid = (int)(ctas < 2 * num_sms ? ctas : 2 * num_sms);
    {
        ProfScope ps("conv1_bn_relu_pool", st);
        conv1_tc_kernel<<<grid, kC1Threads, kC1SmemBytes, st>>>(feat, w1, shift1, out_hi, out_lo, B, H, W);
    }
    SIR_CHECK_LAUNCH("conv1_tc_kernel");
    return SIR_OK;
}

}  // namespace tc
}  // namespace sir
