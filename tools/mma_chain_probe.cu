// Probe (scratch, gpurun): how long do chains of small tcgen05.mma kind::tf32 instructions take when they accumulate into ONE
// tensor-memory accumulator, compared with the same instructions alternating between two accumulators?  (Is a chain of
// dependent N = 64 MMAs bound by the MMA latency or by its issue / operand-fetch rate?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/mma_chain_probe tools/mma_chain_probe.cu
#include <cstdio>
#include <vector>
#include "../speech-intent-recognizer_b200/csrc/tc_common.cuh"
using namespace sir::tc;

__global__ void __launch_bounds__(128) probe(long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)(i & 255);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base_s;
    uint32_t phase = 0;
    if (warp == 0) {
        const uint64_t a = make_kmajor_desc<128>(smem_u32(smem)), b = make_kmajor_desc<128>(smem_u32(smem) + 32768u);
        for (int variant = 0; variant < 11; ++variant) {
            for (int rep = 0; rep < 3; ++rep) {
                __syncwarp();
                const long long t0 = clock64();
                if (elect_one_sync()) {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

int main() {
    long long* d;
    cudaMalloc(&d, 11 * 3 * 2 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    probe<<<1, 128, 70000>>>(d);
    printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<long long> h(66);
    cudaMemcpy(h.data(), d, 66 * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[11] = {"24 SS tf32 N=64, one accumulator", "24 SS tf32 N=64, two accumulators", "24 SS tf32 N=64, four accumulators",
                            "8 x N=128 + 8 x N=64 SS tf32, one accumulator", "12 TS tf32 N=32, one accumulator", "12 TS tf32 N=32, two accumulators",
                            "12 SS fp16 N=64 K=16, one accumulator", "48 SS tf32 N=64, one accumulator",
                            "32 SS fp16 N=128 K=16, 128-byte rows (SWIZZLE_128B)", "32 SS fp16 N=128 K=16, 64-byte rows (SWIZZLE_64B)",
                            "32 SS fp16 N=128 K=16, 32-byte rows (SWIZZLE_32B)"};
    for (int v = 0; v < 11; ++v) printf("%-48s issue %5lld cycles, done %5lld cycles (third run)\n", names[v], h[(v * 3 + 2) * 2], h[(v * 3 + 2) * 2 + 1]);
    return 0;
}
