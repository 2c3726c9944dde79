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


// ---------------------------------------------------------------------------------------------------------------
// Pool-aligned variant (the one the model runs).  The kernel above is bound by instruction issue, and most of its
// instructions are the 2 x 2 max-pool across lanes (24 shuffles + 48 selects per thread) and the (hi, lo) split of
// every input value nine times over.  Here a thread owns one POOLED pixel: it loads the 4 x 4 input patch behind it,
// splits the 16 values once, and writes four operand rows - the 3 x 3 neighbourhoods of its four window members -
// into four operand matrices (member (dy, dx) = matrix 2 dy + dx, 128 rows x 64 B, SWIZZLE_64B).  Four groups of
// three N = 32 MMAs put the members' results into four column groups of the SAME tensor-memory lane, so the pool is
// three max instructions per channel in registers, and the thread finishes its pixel alone: + shift, ReLU, (hi, lo)
// split, two 32-byte stores per 16 channels.  16.2 M warp instructions per 256-utterance launch instead of 42 M.
// A warpgroup works on a patch of (128 / PW) x PW pooled pixels; 4 warpgroups x 128 accumulator columns fill tensor
// memory, so one CTA per SM.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kC1pAWg = 4u * 128u * 64u;                // a warpgroup's four member operands: 32 KB
constexpr uint32_t kC1pOffW = kC1Wgs * kC1pAWg;              // weight operand image as above (64 rows x 128 B)
constexpr uint32_t kC1pOffShift = kC1pOffW + 8192;           // 32 floats
constexpr uint32_t kC1pOffBar = kC1pOffShift + 128;
constexpr uint32_t kC1pSmemBytes = kC1pOffBar + 64 + 1024;

__device__ __forceinline__ uint32_t c1_pack_sat(float first, float second) {      // {second : first} as fp16, saturating
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
    return r;
}
__device__ __forceinline__ void c1_split_pair_sat(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = c1_pack_sat(a, b);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = c1_pack_sat(a - f.x, b - f.y);
}
// element c of a patch row held as two packed words -> 16-bit halves picked into one word: {E(w2, c2) : E(w1, c1)}
template <int C1, int C2>
__device__ __forceinline__ uint32_t c1_pick(const uint32_t (&r1)[2], const uint32_t (&r2)[2]) {
    constexpr uint32_t sel = (C1 & 1 ? 0x32u : 0x10u) | ((C2 & 1 ? 0x76u : 0x54u) << 8);
    return __byte_perm(r1[C1 >> 1], r2[C2 >> 1], sel);
}
template <int C1>
__device__ __forceinline__ uint32_t c1_pick_last(const uint32_t (&r1)[2]) {       // {0 : E(w1, c1)}
    return C1 & 1 ? r1[C1 >> 1] >> 16 : r1[C1 >> 1] & 0xFFFFu;
}
// the five packed words (9 taps + a zero) of window member (DY, DX) from the thread's 4 x 4 patch
template <int DY, int DX>
__device__ __forceinline__ void c1_member_row(const uint32_t (&w)[4][2], uint32_t (&o)[5]) {
    o[0] = c1_pick<DX, DX + 1>(w[DY], w[DY]);
    o[1] = c1_pick<DX + 2, DX>(w[DY], w[DY + 1]);
    o[2] = c1_pick<DX + 1, DX + 2>(w[DY + 1], w[DY + 1]);
    o[3] = c1_pick<DX, DX + 1>(w[DY + 2], w[DY + 2]);
    o[4] = c1_pick_last<DX + 2>(w[DY + 2]);
}
template <int DY, int DX>
__device__ __forceinline__ void c1_store_member(uint32_t row_addr, int sw, const uint32_t (&h)[4][2], const uint32_t (&l)[4][2]) {
    uint32_t a[5], b[5];
    c1_member_row<DY, DX>(h, a);
    c1_member_row<DY, DX>(l, b);
    const uint32_t base = row_addr + (uint32_t)(2 * DY + DX) * 8192u;
    c1_st_v4(base + (uint32_t)((0 ^ sw) << 4), a[0], a[1], a[2], a[3]);      // K 0..7   (hi)
    c1_st_v4(base + (uint32_t)((1 ^ sw) << 4), a[4], 0u, 0u, 0u);            // K 8..15  (hi)
    c1_st_v4(base + (uint32_t)((2 ^ sw) << 4), b[0], b[1], b[2], b[3]);      // K 16..23 (lo)
    c1_st_v4(base + (uint32_t)((3 ^ sw) << 4), b[4], 0u, 0u, 0u);            // K 24..31 (lo)
}
__device__ __forceinline__ void c1_tmem_ld16(uint32_t taddr, float (&v)[16]) {    // no wait: several loads share one
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void c1_st_global_v8(void* dst, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

template <int PW>
__global__ void __launch_bounds__(kC1Threads, 1) conv1_pool_tc_kernel(const float* __restrict__ feat,     // [B, H, W] fp32
                                                                      const float* __restrict__ w1,       // [32][9] BN-folded
                                                                      const float* __restrict__ shift1,   // [32]
                                                                      __half* __restrict__ out_hi,        // [B, H/2, W/2, 32]
                                                                      __half* __restrict__ out_lo, int B, int H, int W) {
    constexpr int PH = 128 / PW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const int g = warp >> 2, q = warp & 3, wt = tid & 127;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kC1pOffBar) + g;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kC1pOffBar + 32);
    float* s_shift = reinterpret_cast<float*>(smem + kC1pOffShift);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_addr = sbase + (uint32_t)g * kC1pAWg;

    if (wt == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(s_tmem);
    for (int i = tid; i < 2048; i += kC1Threads) reinterpret_cast<uint32_t*>(smem + kC1pOffW)[i] = 0u;
    __syncthreads();
    if (tid < 32) {                                          // weight operand image, as in the kernel above
        const int n = tid;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = 2 * e < 9 ? __ldg(w1 + n * 9 + 2 * e) : 0.f, b = 2 * e + 1 < 9 ? __ldg(w1 + n * 9 + 2 * e + 1) : 0.f;
            c1_split_pair(a, b, hi[e], lo[e]);
        }
        const uint32_t r_hi = sbase + kC1pOffW + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
        const uint32_t r_lo = r_hi + 4096u;
        const int sw = n & 7;
        c1_st_v4(r_hi + (uint32_t)((0 ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
        c1_st_v4(r_hi + (uint32_t)((1 ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
        c1_st_v4(r_lo + (uint32_t)((0 ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
        c1_st_v4(r_lo + (uint32_t)((1 ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
        s_shift[n] = __ldg(shift1 + n);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_wg = *s_tmem + (uint32_t)g * 128u;
    const uint32_t tmem_row = tmem_wg + ((uint32_t)(q * 32) << 16);

    const int H2 = H / 2, W2 = W / 2;
    const int tiles_x = (W2 + PW - 1) / PW, tiles_y = (H2 + PH - 1) / PH, tiles_img = tiles_x * tiles_y;
    const int num_tiles = B * tiles_img;
    constexpr uint32_t idesc = make_idesc_f16(128, 32);
    const uint64_t w_hi = make_kmajor_desc<128>(sbase + kC1pOffW), w_lo = make_kmajor_desc<128>(sbase + kC1pOffW + 4096u);
    const uint64_t a_desc = make_kmajor_desc<64>(a_addr);
    const int py = wt / PW, px = wt % PW;                    // this thread's pooled pixel inside the patch = row wt of the operands
    const uint32_t row_addr = a_addr + (uint32_t)((wt >> 3) * 512 + (wt & 7) * 64);
    const int sw = (wt >> 1) & 3;                            // SWIZZLE_64B: 16-byte chunk index ^ address bits [7, 9)

    const int stride = gridDim.x * kC1Wgs;
    const int d_img = stride / tiles_img, d_r = stride - d_img * tiles_img, d_ty = d_r / tiles_x, d_tx = d_r - d_ty * tiles_x;
    int tile = blockIdx.x * kC1Wgs + g;
    int img = tile / tiles_img, ty = (tile - img * tiles_img) / tiles_x, tx = (tile - img * tiles_img) - ty * tiles_x;
    // the 4 x 4 input patch behind pooled pixel (Y2, X2) of image `img` (rows 2 Y2 - 1 .., columns 2 X2 - 1 ..), zero outside
    auto load_patch = [&](int img_, int ty_, int tx_, float (&v)[16]) {
        const int Y2 = ty_ * PH + py, X2 = tx_ * PW + px;
        const bool in_img = Y2 < H2 && X2 < W2;
        const float* __restrict__ src = feat + (int64_t)img_ * H * W;
        const int r0 = 2 * Y2 - 1, c0 = 2 * X2 - 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i;
            const bool rv = in_img && r >= 0 && r < H;
            const float* __restrict__ rowp = src + (int64_t)(rv ? r : 0) * W + (rv ? c0 : 0);
            v[4 * i + 0] = rv && c0 >= 0 ? __ldg(rowp) : 0.f;
            v[4 * i + 1] = rv ? __ldg(rowp + 1) : 0.f;
            v[4 * i + 2] = rv ? __ldg(rowp + 2) : 0.f;
            v[4 * i + 3] = rv && c0 + 3 < W ? __ldg(rowp + 3) : 0.f;
        }
    };
    float patch[16];
    if (tile < num_tiles) load_patch(img, ty, tx, patch);
    uint32_t it = 0;
    for (; tile < num_tiles; tile += stride, ++it) {
        const int Y2 = ty * PH + py, X2 = tx * PW + px;
        const bool inside = Y2 < H2 && X2 < W2;
        const int64_t pix = ((int64_t)img * H2 + Y2) * W2 + X2;
        // ---- build: split the patch once, four operand rows -----------------------------------------------------------
        {
            uint32_t h[4][2], l[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c1_split_pair_sat(patch[4 * i], patch[4 * i + 1], h[i][0], l[i][0]);
                c1_split_pair_sat(patch[4 * i + 2], patch[4 * i + 3], h[i][1], l[i][1]);
            }
            c1_store_member<0, 0>(row_addr, sw, h, l);
            c1_store_member<0, 1>(row_addr, sw, h, l);
            c1_store_member<1, 0>(row_addr, sw, h, l);
            c1_store_member<1, 1>(row_addr, sw, h, l);
        }
        fence_proxy_async();                                 // generic-proxy stores -> visible to the tensor core
        tc_fence_before();                                   // (the previous tile's accumulator reads are done)
        c1_wg_barrier(g);
        if (q == 0) {
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t a_j = a_desc + (uint64_t)((j * 8192u) >> 4);
                    umma_f16(tmem_wg + j * 32, a_j, w_hi, idesc, 0u);                          // hi . W_hi
                    umma_f16(tmem_wg + j * 32, a_j, w_lo, idesc, 1u);                          // hi . W_lo
                    umma_f16(tmem_wg + j * 32, desc_advance_k(a_j, 16), w_hi, idesc, 1u);      // lo . W_hi
                }
                umma_commit(bar);
            }
            __syncwarp();
        }
        // the next tile's patch is loaded while this tile's MMAs run and its epilogue waits for tensor memory
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
        if (tile + stride < num_tiles) load_patch(img, ty, tx, patch);
        c1_mma_wait(bar, it & 1u);
        tc_fence_after();
        // ---- epilogue: max over the four members (column groups), + shift, ReLU, (hi, lo), 16 channels at a time --------
        {
#pragma unroll
            for (int cc = 0; cc < 32; cc += 16) {
                float v0[16], v1[16], v2[16], v3[16];
                c1_tmem_ld16(tmem_row + cc, v0);
                c1_tmem_ld16(tmem_row + 32 + cc, v1);
                c1_tmem_ld16(tmem_row + 64 + cc, v2);
                c1_tmem_ld16(tmem_row + 96 + cc, v3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t hw[8], lw[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 sh = *reinterpret_cast<const float2*>(s_shift + cc + i);
                    const float a = fmaxf(fmaxf(fmaxf(v0[i], v1[i]), fmaxf(v2[i], v3[i])) + sh.x, 0.f);
                    const float b = fmaxf(fmaxf(fmaxf(v0[i + 1], v1[i + 1]), fmaxf(v2[i + 1], v3[i + 1])) + sh.y, 0.f);
                    c1_split_pair_sat(a, b, hw[i >> 1], lw[i >> 1]);
                }
                if (inside) {
                    c1_st_global_v8(out_hi + pix * 32 + cc, hw);
                    c1_st_global_v8(out_lo + pix * 32 + cc, lw);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(*s_tmem);
    }
}

template <int PW>
static int conv1_pool_launch(const float* feat, const float* w1, const float* shift1, __half* out_hi, __half* out_lo, int B,
                             int H, int W, int num_sms, cudaStream_t st) {
    constexpr int PH = 128 / PW;
    auto kern = conv1_pool_tc_kernel<PW>;
    SIR_SMEM_OPTIN(kern, kC1pSmemBytes);
    const int64_t tiles = (int64_t)B * ((W / 2 + PW - 1) / PW) * ((H / 2 + PH - 1) / PH);
    const int64_t ctas = (tiles + kC1Wgs - 1) / kC1Wgs;
    const int grid = (int)(ctas < num_sms ? ctas : num_sms);
    {
        ProfScope ps("conv1_bn_relu_pool", st);
        kern<<<grid, kC1Threads, kC1pSmemBytes, st>>>(feat, w1, shift1, out_hi, out_lo, B, H, W);
    }
    SIR_CHECK_LAUNCH("conv1_pool_tc_kernel");
    return SIR_OK;
}

int conv1_tc(const float* feat, const float* w1, const float* shift1, __half* out_hi, __half* out_lo, int B, int H, int W,
             int num_sms, cudaStream_t st) {
    if (B < 1) return SIR_OK;
    if ((H & 1) || (W & 1)) return fail(SIR_ERR_INVALID, "conv1: H and W must be even (got %d x %d)", H, W);
    static const bool per_pixel = [] {                       // SIR_CONV1_KERNEL=pixel: the one-thread-per-input-pixel kernel (A/B)
        const char* v = getenv("SIR_CONV1_KERNEL");
        return v && (v[0] == 'p' || v[0] == 'P');
    }();
    if (!per_pixel) {
        // pooled patch of 16 x 8 or 8 x 16 per warpgroup: whichever covers the pooled image with fewer patches
        const int H2 = H / 2, W2 = W / 2;
        const int n8 = ((H2 + 15) / 16) * ((W2 + 7) / 8), n16 = ((H2 + 7) / 8) * ((W2 + 15) / 16);
        return n8 <= n16 ? conv1_pool_launch<8>(feat, w1, shift1, out_hi, out_lo, B, H, W, num_sms, st)
                         : conv1_pool_launch<16>(feat, w1, shift1, out_hi, out_lo, B, H, W, num_sms, st);
    }
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
