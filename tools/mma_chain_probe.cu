// Probe (scratch, gpurun): how long do chains of small tcgen05.mma kind::tf32 instructions take when they accumulate into ONE
// tensor-memory accumulator, compared with the same instructions alternating between two accumulators?  (Is a chain of
// dependent N = 64 MMAs bound by the MMA latency or by its issue / operand-fetch rate?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/mma_chain_probe tools/mma_chain_probe.cu
#include <cstdio>
#include <vector>
#include "../speech-intent-recognizer_b200/csrc/tc_common.cuh"
using namespace sir::tc;

__global__ void __launch_bounds__(192) probe(long long* out, const uint8_t* __restrict__ gsrc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint64_t stage_bar[3];
    __shared__ uint64_t copy_bar;
    __shared__ volatile int copy_stop, copy_on, ld_on, ld4_on;
    __shared__ long long copied;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 192) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)(i & 255);
    if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 3; ++i) mbar_init(&stage_bar[i], 1); mbar_init(&copy_bar, 1); copy_stop = 0; copy_on = 0; ld_on = 0; ld4_on = 0; copied = 0; fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base_s;
    uint32_t phase = 0;
    if (warp == 0) {
        const uint64_t a = make_kmajor_desc<128>(smem_u32(smem)), b = make_kmajor_desc<128>(smem_u32(smem) + 32768u);
        for (int variant = 0; variant < 18; ++variant) {
            if (variant == 15) { ld_on = 0; ld4_on = 1; __threadfence_block(); __nanosleep(2000); }  // ... reader on the ISSUER's scheduler (warp 4)
            if (variant == 14) { copy_on = 0; ld_on = 1; __threadfence_block(); __nanosleep(2000); }  // ... with the other warps reading tensor memory
            if (variant == 13) { copy_on = 1; __threadfence_block(); __nanosleep(2000); }     // conv2's tile again, with bulk copies landing
            for (int rep = 0; rep < 3; ++rep) {
                __syncwarp();
                const long long t0 = clock64();
                if (variant >= 16) {
                    // conv2's issue loop as written: three batches of 12 MMAs (one per horizontal tap), each behind a
                    // tcgen05.fence::after_thread_sync (16) and closed by a tcgen05.commit that releases the stage (16, 17)
                    const uint64_t a64 = make_kmajor_desc<64>(smem_u32(smem)), b64 = make_kmajor_desc<64>(smem_u32(smem) + 32768u);
                    for (uint32_t kw = 0; kw < 3; ++kw) {
                        if (variant == 16) tc_fence_after();
                        if (elect_one_sync()) {
                            for (uint32_t i = 6 * kw; i < 6 * kw + 6; ++i) {
                                umma_f16(tm, a64 + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 128), i ? 1u : 0u);
                                umma_f16(tm, a64 + 640u + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 64), 1u);
                            }
                            umma_commit(&stage_bar[kw]);
                            if (kw == 2) umma_commit(&bar);
                        }
                        __syncwarp();
                    }
                } else if (elect_one_sync()) {
                    if (variant == 0) {            // 24 x (128 x 64 x 8) SS into one accumulator
                        for (uint32_t i = 0; i < 24; ++i) umma_tf32(tm, a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 64), i ? 1u : 0u);
                    } else if (variant == 1) {     // the same, alternating between two accumulators
                        for (uint32_t i = 0; i < 24; ++i) umma_tf32(tm + 64u * (i & 1), a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 64), i > 1 ? 1u : 0u);
                    } else if (variant == 2) {     // four accumulators
                        for (uint32_t i = 0; i < 24; ++i) umma_tf32(tm + 64u * (i & 3), a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 64), i > 3 ? 1u : 0u);
                    } else if (variant == 3) {     // 8 x N = 128 + 8 x N = 64 (stacked B), one accumulator
                        for (uint32_t i = 0; i < 8; ++i) umma_tf32(tm, a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 128), i ? 1u : 0u);
                        for (uint32_t i = 0; i < 8; ++i) umma_tf32(tm, a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 64), 1u);
                    } else if (variant == 4) {     // 12 x (128 x 32 x 8) TS, one accumulator (stage 1)
                        for (uint32_t i = 0; i < 12; ++i) umma_tf32_ts(tm, tm + 256u + 8u * (i & 3), b + 2u * (i & 3), make_idesc_tf32(128, 32), i ? 1u : 0u);
                    } else if (variant == 5) {     // 12 x TS over two accumulators
                        for (uint32_t i = 0; i < 12; ++i) umma_tf32_ts(tm + 32u * (i & 1), tm + 256u + 8u * (i & 3), b + 2u * (i & 3), make_idesc_tf32(128, 32), i > 1 ? 1u : 0u);
                    } else if (variant == 6) {     // 12 x fp16 (128 x 64 x 16) SS, one accumulator (what fp16 pieces would need)
                        for (uint32_t i = 0; i < 12; ++i) umma_f16(tm, a + 2u * (i & 3), b + 2u * (i & 3), make_idesc_f16(128, 64), i ? 1u : 0u);
                    } else if (variant == 8) {     // 32 x fp16 (128 x 128 x 16) SS, operands in SWIZZLE_128B rows (conv3 / GEMM layout)
                        for (uint32_t i = 0; i < 32; ++i) umma_f16(tm, a + 2u * (i & 3), b + 2u * (i & 3), make_idesc_f16(128, 128), i ? 1u : 0u);
                    } else if (variant == 9) {     // the same with 64-byte rows (SWIZZLE_64B: conv2's C_in = 32 layout), two K slices per row
                        const uint64_t a64 = make_kmajor_desc<64>(smem_u32(smem)), b64 = make_kmajor_desc<64>(smem_u32(smem) + 32768u);
                        for (uint32_t i = 0; i < 32; ++i) umma_f16(tm, a64 + 2u * (i & 1), b64 + 2u * (i & 1), make_idesc_f16(128, 128), i ? 1u : 0u);
                    } else if (variant == 10) {    // the same with 32-byte rows (SWIZZLE_32B): one K = 16 slice per row, slices in separate tiles
                        uint64_t a32 = 0, b32 = 0;
                        a32 |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
                        b32 |= (uint64_t)(((smem_u32(smem) + 32768u) >> 4) & 0x3FFF) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
                        for (uint32_t i = 0; i < 32; ++i) umma_f16(tm, a32 + (4096u >> 4) * (i & 1), b32 + (4096u >> 4) * (i & 1), make_idesc_f16(128, 128), i ? 1u : 0u);
                    } else if (variant >= 11 && variant <= 15) {   // conv2's tile: 18 x (N = 128 then N = 64) on 64-byte rows; 12: grouped by shape
                        const uint64_t a64 = make_kmajor_desc<64>(smem_u32(smem)), b64 = make_kmajor_desc<64>(smem_u32(smem) + 32768u);
                        if (variant != 12) {
                            for (uint32_t i = 0; i < 18; ++i) {
                                umma_f16(tm, a64 + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 128), i ? 1u : 0u);
                                umma_f16(tm, a64 + 640u + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 64), 1u);
                            }
                        } else {
                            for (uint32_t i = 0; i < 18; ++i) umma_f16(tm, a64 + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 128), i ? 1u : 0u);
                            for (uint32_t i = 0; i < 18; ++i) umma_f16(tm, a64 + 640u + 64u * (i % 3) + 2u * (i & 1), b64 + 512u * (i % 4) + 2u * (i & 1), make_idesc_f16(128, 64), 1u);
                        }
                    } else {                       // 48 x (128 x 64 x 8) SS into one accumulator: twice variant 0 (slope)
                        for (uint32_t i = 0; i < 48; ++i) umma_tf32(tm, a + 2u * (i & 3) + 1024u * ((i >> 2) & 1), b + 2u * (i & 3), make_idesc_tf32(128, 64), i ? 1u : 0u);
                    }
                    umma_commit(&bar);
                }
                __syncwarp();
                const long long t1 = clock64();
                mbar_wait(&bar, phase);
                phase ^= 1u;
                const long long t2 = clock64();
                if (tid == 0) { out[(variant * 3 + rep) * 2] = t1 - t0; out[(variant * 3 + rep) * 2 + 1] = t2 - t0; }
                tc_fence_after();
            }
        }
    }
    if (warp == 0 && tid == 0) { copy_stop = 1; __threadfence_block(); }
    if (warp == 4) {
        // the same reads + the pooling shuffles of conv2's epilogue, from a warp on the MMA-issuing warp's own scheduler
        // (warp 4 and warp 0 share sub-partition 0, as conv2's MMA warp 1 shares sub-partition 1 with epilogue warp 5)
        float acc = 0.f;
        long long n = 0;
        while (!copy_stop) {
            if (!ld4_on) { __nanosleep(200); continue; }
            for (int c = 0; c < 128; c += 32) {
                float v[32];
                tmem_ld_32x32(tm + 256u + c, v);
                for (int i = 0; i < 32; ++i) acc += fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
            }
            ++n;
        }
        if (acc == 12345.678f) out[0] = 0;
        if ((tid & 31) == 0) out[18 * 6 + 2] = n;
    }
    if (warp == 2 || warp == 3) {
        // what conv2's epilogue warps do while the MMAs of the next tile run: tcgen05.ld of the OTHER accumulator (columns
        // 256..383 of this warp's lane quadrant), 4 x 32 columns per pass, back to back
        float acc = 0.f;
        long long n = 0;
        while (!copy_stop) {
            if (!ld_on) { __nanosleep(200); continue; }
            for (int c = 0; c < 128; c += 32) {
                float v[32];
                tmem_ld_32x32(tm + 256u + c + ((uint32_t)(warp * 32) << 16), v);
                for (int i = 0; i < 32; ++i) acc += v[i];
            }
            ++n;
        }
        if (acc == 12345.678f) out[0] = 0;
        if (warp == 2 && (tid & 31) == 0) out[18 * 6 + 1] = n;
    }
    if (warp == 1 && (tid & 31) == 0) {
        // TMA-style traffic INTO shared memory while the MMAs of variant 13 read their operands: 20 KB bulk copies from
        // global memory (L2-resident after the first pass), one in flight, like conv2's activation ring
        uint32_t ph = 0;
        long long n = 0;
        while (!copy_stop) {
            if (!copy_on) { __nanosleep(200); continue; }
            mbar_arrive_expect_tx(&copy_bar, 20480u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + 45056u),
                         "l"(gsrc + (n % 8) * 20480), "r"(20480u), "r"(smem_u32(&copy_bar))
                         : "memory");
            mbar_wait(&copy_bar, ph);
            ph ^= 1u;
            ++n;
        }
        copied = n;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) out[18 * 6] = copied;
    if (warp == 0) tmem_dealloc<512>(tm);
}

int main() {
    long long* d;
    cudaMalloc(&d, (18 * 3 * 2 + 3) * sizeof(long long));
    uint8_t* gsrc;
    cudaMalloc(&gsrc, 8 * 20480);
    cudaMemset(gsrc, 1, 8 * 20480);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    probe<<<1, 192, 70000>>>(d, gsrc);
    printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<long long> h(111);
    cudaMemcpy(h.data(), d, 111 * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[18] = {"24 SS tf32 N=64, one accumulator", "24 SS tf32 N=64, two accumulators", "24 SS tf32 N=64, four accumulators",
                            "8 x N=128 + 8 x N=64 SS tf32, one accumulator", "12 TS tf32 N=32, one accumulator", "12 TS tf32 N=32, two accumulators",
                            "12 SS fp16 N=64 K=16, one accumulator", "48 SS tf32 N=64, one accumulator",
                            "32 SS fp16 N=128 K=16, 128-byte rows (SWIZZLE_128B)", "32 SS fp16 N=128 K=16, 64-byte rows (SWIZZLE_64B)",
                            "32 SS fp16 N=128 K=16, 32-byte rows (SWIZZLE_32B)", "conv2 tile: 18 x (N=128, N=64) alternating",
                            "conv2 tile: 18 x N=128 then 18 x N=64", "conv2 tile alternating + 20 KB bulk copies into smem", "conv2 tile alternating + two warps reading tensor memory",
                            "conv2 tile alternating + reads and shuffles on the issuer's scheduler",
                            "conv2 tile in 3 batches: fence + 12 MMAs + commit each", "conv2 tile in 3 batches: 12 MMAs + commit each (no fence)"};
    for (int v = 0; v < 18; ++v) printf("%-48s issue %5lld cycles, done %5lld cycles (third run)\n", names[v], h[(v * 3 + 2) * 2], h[(v * 3 + 2) * 2 + 1]);
    printf("bulk copies of 20 KB completed while variant 13 ran (3 timed runs + gaps): %lld\n", h[108]);
    printf("passes of 128 tensor-memory columns read by warp 2 while variant 14 ran: %lld\n", h[109]);
    printf("passes by warp 4 while variant 15 ran: %lld\n", h[110]);
    return 0;
}
