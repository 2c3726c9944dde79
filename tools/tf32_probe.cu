// Probe of tcgen05.mma kind::tf32 on sm_100a (scratch, run under gpurun): what the frontend's tf32 formulation relies on.
//   1. SS mode: A [128 x 32] and B [32 x 32] tf32, K-major, SWIZZLE_128B rows of 128 bytes, four K = 8 dispatches.
//   2. TS mode: the same A written to tensor memory with tcgen05.st.32x32b (lane = row, one 32-bit column per element).
//   3. What the tensor core does with the 13 low mantissa bits of an fp32 operand (ignored = truncation, or rounded).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tf32_probe tools/tf32_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../speech-intent-recognizer_b200/csrc/tc_common.cuh"

using namespace sir::tc;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

__device__ __forceinline__ uint32_t sw128(int row, int k_elem4) {       // byte offset of 4-byte element k of `row`
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k_elem4 >> 2) ^ (row & 7)) & 7) << 4) + (k_elem4 & 3) * 4);
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D_ss,
                                             float* __restrict__ D_ts) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                 // 128 rows x 128 B
    uint8_t* sB = smem + 16384;         // 32 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) *reinterpret_cast<float*>(sA + sw128(i >> 5, i & 31)) = A[i];
    for (int i = tid; i < 32 * 32; i += 128) *reinterpret_cast<float*>(sB + sw128(i >> 5, i & 31)) = B[i];    // B[n][k]
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<128>(&tmem_base_s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base_s;
    constexpr uint32_t idesc = make_idesc_tf32(128, 32);
    // ---- SS
    if (warp == 0 && elect_one_sync()) {
        const uint64_t a = make_kmajor_desc<128>(smem_u32(sA)), b = make_kmajor_desc<128>(smem_u32(sB));
        for (int kk = 0; kk < 4; ++kk) umma_tf32(tm, a + 2 * kk, b + 2 * kk, idesc, kk ? 1u : 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        float v[32];
        tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), v);
        for (int i = 0; i < 32; ++i) D_ss[(warp * 32 + lane) * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    // ---- TS: A into tensor-memory columns [64, 96)
    {
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(A[(warp * 32 + lane) * 32 + i]);
        tmem_st_32x32(tm + 64u + ((uint32_t)(warp * 32) << 16), r);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one_sync()) {
        const uint64_t b = make_kmajor_desc<128>(smem_u32(sB));
        for (int kk = 0; kk < 4; ++kk) umma_tf32_ts(tm + 32u, tm + 64u + 8u * kk, b + 2 * kk, idesc, kk ? 1u : 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    tc_fence_after();
    {
        float v[32];
        tmem_ld_32x32(tm + 32u + ((uint32_t)(warp * 32) << 16), v);
        for (int i = 0; i < 32; ++i) D_ts[(warp * 32 + lane) * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}

static float to_tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::vector<float> A(128 * 32), B(32 * 32);
    unsigned s = 7u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.f - 0.5f; };
    for (auto& v : A) v = to_tf32_trunc(rnd());
    for (auto& v : B) v = to_tf32_trunc(rnd());
    // rounding probe: row 0 of A = (1 + 2^-11 + 2^-12) e_0 (low bits below the tf32 grid 2^-10), B[n][0] = 1 for n = 0
    for (int k = 0; k < 32; ++k) A[k] = 0.f;
    A[0] = 1.0f + ldexpf(1.f, -11) + ldexpf(1.f, -12);      // trunc -> 1.0 ; round-to-nearest -> 1 + 2^-10
    for (int k = 0; k < 32; ++k) B[0 * 32 + k] = 0.f;
    B[0] = 1.0f;
    float *dA, *dB, *dS, *dT;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dS, 128 * 32 * 4); cudaMalloc(&dT, 128 * 32 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 128 * 32 * 4); cudaMemset(dT, 0, 128 * 32 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<1, 128, 32768>>>(dA, dB, dS, dT);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> S(128 * 32), T(128 * 32);
    cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost);
    double es = 0, et = 0;
    for (int m = 1; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * (double)B[n * 32 + k];
            es = fmax(es, fabs(ref - S[m * 32 + n]));
            et = fmax(et, fabs(ref - T[m * 32 + n]));
        }
    printf("SS max abs err %.3e   TS max abs err %.3e  (exact tf32 inputs: expect ~1e-7)\n", es, et);
    printf("rounding probe: SS D[0][0] = %.10f  TS D[0][0] = %.10f   (1.0 = truncation, %.10f = round to nearest)\n", S[0], T[0],
           1.0 + ldexp(1.0, -10));
    return (es < 1e-5 && et < 1e-5) ? 0 : 1;
}
