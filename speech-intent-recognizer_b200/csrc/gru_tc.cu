// GRU recurrence of one layer (both directions, all T steps, ONE launch) with the recurrent product on tcgen05.
//
//   r = s(gi_r + W_hr h + b_hr)  z = s(gi_z + W_hz h + b_hz)  n = tanh(gi_n + r * (W_hn h + b_hn))
//   h' = (1 - z) n + z h            torch.nn.GRU, models/models.py:60 of the reference (gi already holds b_i*)
//
// A thread-block CLUSTER of 8 CTAs owns one (direction, slice of NB utterances).  CTA r keeps the recurrent
// weights of hidden units [32 r, 32 r + 32) - 96 gate rows as fp16 (hi, lo) pairs, 96 KB - resident in TENSOR
// MEMORY for all steps (written once with tcgen05.st: lane = gate row, 128 columns each for hi and lo; rows 96..127
// are zero): the MMAs take A from TMEM and fetch only the hidden state from shared memory, and the CTA's shared
// memory footprint no longer includes them.  The hidden state
// never leaves the chip between steps: it lives as the fp16 (hi, lo) B operand [NB x 256] in every CTA's
// shared memory (128-byte swizzled K-major, double buffered).  Per step:
//   1. one thread issues 48 tcgen05.mma (128 x NB x 16; hi.hi, hi.lo, lo.hi over K = 256) into a TMEM
//      accumulator D[gate row, utterance];
//   2. warps 0/1/2 read the r/z/n rows back (TMEM lane = gate row) and transpose them through shared memory;
//   3. every thread updates 8 hidden units of one utterance, writes h' to the layer output y (fp32, plus the
//      fp16 pair the next layer's input GEMM consumes) and PUSHES the fp16 (hi, lo) of its 8 units - one
//      16-byte chunk each - into the next-step B operand of all 8 CTAs through distributed shared memory;
//   4. the pushes are st.async stores that complete their bytes on the `h_full` mbarrier of the CTA they land in; the
//      MMA-issuing warp of each CTA posts the expected byte count, waits (acquire, cluster scope) on its own `h_full`,
//      fences towards the async proxy and issues the next step's MMAs.  No cluster-wide barrier, no fence and no
//      arrive on the pushing warps: a CTA only waits for the data it needs.
// No grid-wide synchronisation, no per-step launch, no L2 round trip on the recurrence's critical path.
// Slices of 64 utterances (batches > 144) run the chain-pipelined variant further down (gru_layer_pp_kernel).
#include "sir_common.cuh"
#include "tc_common.cuh"

namespace sir {
namespace tc {

constexpr int kGtCluster = 8;
constexpr int kGtUnits = 32;                  // hidden units per CTA
constexpr int kGtThreads = 256;
constexpr int kGtWRows = 96;                  // 3 gates x 32 units
constexpr uint32_t kGtColWlo = 128;           // TMEM columns: W_hi [0, 128), W_lo [128, 256), accumulator at 256
constexpr uint32_t kGtColAcc = 256;

template <int NB>
struct GtLayout {
    static constexpr int kHBytes = NB * 256 * 2;                 // one of (hi, lo) of one buffer: 4 K-blocks of NB rows
    static constexpr int kOffH = 0;                              // [2 buffers][hi, lo]
    static constexpr int kOffS = kOffH + 4 * kHBytes;            // gate pre-activations [3][32 units][NB + 1] fp32
    static constexpr int kSStride = NB + 1;
    static constexpr int kOffBar = (kOffS + 3 * 32 * kSStride * 4 + 15) & ~15;
    static constexpr int kSmemBytes = kOffBar + 64 + 1024;
    static_assert(kHBytes % 1024 == 0, "B operand K-blocks must stay 1024-byte aligned");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
// Asynchronous 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on an mbarrier of
// that CTA when the data has landed: the store and its completion signal travel together, the issuing thread neither
// fences nor waits (a release-arrive after plain st.shared::cluster costs a MEMBAR.ALL.GPU per warp and step).
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
                 : "memory");
}
// Gate non-linearities on the SFU: ex2.approx + rcp.approx (relative error ~1e-6, far inside the 1e-3 logit bar)
// instead of expf / tanhf / IEEE division, which cost ~1.2 us of the ~7 us per time step of the first version
// (measured by switching them off).  tanh(x) = 1 - 2 / (1 + e^{2x}) saturates correctly for large |x|.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

template <int NB>
__global__ void __cluster_dims__(kGtCluster, 1, 1) __launch_bounds__(kGtThreads, 1)
    gru_layer_tc_kernel(const __half* __restrict__ w_hi,               // [2 dirs][8 ranks][96 rows][256] fp16
                        const __half* __restrict__ w_lo,
                        const float* __restrict__ gi,                  // [B*T, 1536]
                        const float* __restrict__ bhh,                 // [2][768]
                        float* __restrict__ y,                         // [B, T, 512]
                        __half* __restrict__ y_hi, __half* __restrict__ y_lo, int B, int T) {
    using L = GtLayout<NB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_gate = reinterpret_cast<float*>(smem + L::kOffS);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* h_full = mma_done + 1;          // [2]: operand buffer b holds the complete hidden state of the next step
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_full + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % kGtCluster;
    const int slice = blockIdx.x / kGtCluster;
    const int dir = blockIdx.y;
    const int j0 = rank * kGtUnits, b0 = slice * NB;

    if (tid == 0) {
        mbar_init(mma_done, 1);
        mbar_init(&h_full[0], 1);                               // the issuer's arrive.expect_tx; the pushes complete the bytes
        mbar_init(&h_full[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    // resident weights -> tensor memory: thread t < 128 owns TMEM lane t = gate row t of this CTA (zeros beyond 96)
    if (warp < 4) {
        const int row = tid;
        const size_t src = ((size_t)(dir * kGtCluster + rank) * kGtWRows + row) * 256;      // halves
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const uint4* g = reinterpret_cast<const uint4*>((part ? w_lo : w_hi) + src);
#pragma unroll 1
            for (int c = 0; c < 4; c += 2) {                 // 2 x 32 columns per pass: 16 loads of 16 bytes in flight
                uint32_t r[2][32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint4 v = row < kGtWRows ? __ldg(g + c * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
                    r[i >> 3][4 * (i & 7) + 0] = v.x;
                    r[i >> 3][4 * (i & 7) + 1] = v.y;
                    r[i >> 3][4 * (i & 7) + 2] = v.z;
                    r[i >> 3][4 * (i & 7) + 3] = v.w;
                }
                tmem_st_32x32(lane_addr + (part ? kGtColWlo : 0u) + (uint32_t)(c * 32), r[0]);
                tmem_st_32x32(lane_addr + (part ? kGtColWlo : 0u) + (uint32_t)(c * 32 + 32), r[1]);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // update role: thread -> (utterance i, group of 8 hidden units ug)
    const bool updater = tid < 4 * NB;
    const int ui = tid >> 2, ug = tid & 3;
    const int ubb = b0 + ui;
    const bool uvalid = updater && ubb < B;
    float b_r[8], b_z[8], b_n[8], hprev[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int u = j0 + 8 * ug + e;
        b_r[e] = __ldg(bhh + dir * 768 + u);
        b_z[e] = __ldg(bhh + dir * 768 + 256 + u);
        b_n[e] = __ldg(bhh + dir * 768 + 512 + u);
        hprev[e] = 0.f;
    }
    // where this thread's 16-byte chunk (8 units of utterance ui) sits inside a swizzled B-operand buffer
    uint32_t chunk_off = 0;
    {
        const int k = j0 + 8 * ug, kb = k >> 6, chunk = (k & 63) >> 3;
        chunk_off = (uint32_t)(kb * (NB * 128) + (ui >> 3) * 1024 + (ui & 7) * 128 + ((chunk ^ (ui & 7)) << 4));
    }
    // everybody's barriers/TMEM are set up before any peer may push into this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");

    constexpr uint32_t idesc = make_idesc_f16(128, NB);
    constexpr uint32_t kStateBytes = NB * 256 * 2 * 2;                      // the slice's hidden state, hi + lo: what 8 CTAs push
    const bool issue_warp = uniform_warp_idx() == kGtThreads / 32 - 1;      // provably warp-uniform
    if (issue_warp) {
        if (T > 1 && elect_one_sync()) mbar_arrive_expect_tx(&h_full[1], kStateBytes);      // the pushes of step 0
        __syncwarp();
    }

    for (int s = 0; s < T; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        const int cur = s & 1, nxt = cur ^ 1;
        // gate pre-activations of the input projection (consumed after the MMA: the loads overlap it)
        float4 gin[3][2];
        if (updater) {
            const float4* gp = reinterpret_cast<const float4*>(
                gi + ((int64_t)(uvalid ? ubb : B - 1) * T + t) * 1536 + dir * 768 + j0 + 8 * ug);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                gin[g][0] = __ldg(gp + g * 64);
                gin[g][1] = __ldg(gp + g * 64 + 1);
            }
        }
        if (s > 0) {
            // (1) D[128 gate rows, NB utterances] = W_slice[128, 256] . h^T   (h: buffer `cur`)
            if (issue_warp) {                      // the last warp has no update work: the MMAs of step s start the
                                                   // moment the state is complete; all lanes walk, one elected lane issues
                mbar_wait_cluster(&h_full[cur], (uint32_t)((s - 1) >> 1) & 1u);     // all 8 CTAs' pushes of step s-1 landed
                fence_proxy_async_all();             // the pushes were generic-proxy writes; the MMAs read through the async proxy
                tc_fence_after();
                if (elect_one_sync()) {
                    if (s + 1 < T) mbar_arrive_expect_tx(&h_full[nxt], kStateBytes);          // the pushes of step s
                    const uint32_t hb = sbase + L::kOffH + cur * 2 * L::kHBytes;
                    const uint32_t d_acc = tmem_base + kGtColAcc;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t b_hi = make_kmajor_desc<128>(hb + kb * (NB * 128));
                        const uint64_t b_lo = make_kmajor_desc<128>(hb + L::kHBytes + kb * (NB * 128));
#pragma unroll
                        for (int k = 0; k < 64; k += 16) {
                            const uint32_t a_hi = tmem_base + (uint32_t)((kb * 64 + k) >> 1);      // 2 halves per column
                            const uint32_t a_lo = a_hi + kGtColWlo;
                            umma_f16_ts(d_acc, a_hi, desc_advance_k(b_hi, k), idesc, (kb | k) ? 1u : 0u);
                            umma_f16_ts(d_acc, a_hi, desc_advance_k(b_lo, k), idesc, 1u);
                            umma_f16_ts(d_acc, a_lo, desc_advance_k(b_hi, k), idesc, 1u);
                        }
                    }
                    umma_commit(mma_done);
                }
                __syncwarp();
            }
            // (2) accumulator rows -> shared memory, transposed to [gate][unit][utterance]
            if (warp < 3) {
                mbar_wait(mma_done, (uint32_t)(s - 1) & 1u);
                tc_fence_after();
                uint32_t r[NB];
#pragma unroll
                for (int c = 0; c < NB; c += 16) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[c + 0]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]),
                          "=r"(r[c + 6]), "=r"(r[c + 7]), "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]),
                          "=r"(r[c + 12]), "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                        : "r"(tmem_base + kGtColAcc + ((uint32_t)(warp * 32) << 16) + c)
                        : "memory");
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");      // one wait for all column chunks
#pragma unroll
                for (int i = 0; i < NB; ++i) s_gate[(warp * 32 + lane) * L::kSStride + i] = __uint_as_float(r[i]);
                tc_fence_before();
            }
            __syncthreads();
        }

        // (3) state update for 8 units of one utterance, output, and push of the next-step operand
        if (updater) {
            const float gr[8] = {gin[0][0].x, gin[0][0].y, gin[0][0].z, gin[0][0].w,
                                 gin[0][1].x, gin[0][1].y, gin[0][1].z, gin[0][1].w};
            const float gz[8] = {gin[1][0].x, gin[1][0].y, gin[1][0].z, gin[1][0].w,
                                 gin[1][1].x, gin[1][1].y, gin[1][1].z, gin[1][1].w};
            const float gn[8] = {gin[2][0].x, gin[2][0].y, gin[2][0].z, gin[2][0].w,
                                 gin[2][1].x, gin[2][1].y, gin[2][1].z, gin[2][1].w};
            float hn[8];
            uint32_t hi2[4], lo2[4];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float ar = 0.f, az = 0.f, an = 0.f;
                if (s > 0) {
                    const int u = 8 * ug + e;
                    ar = s_gate[(0 * 32 + u) * L::kSStride + ui];
                    az = s_gate[(1 * 32 + u) * L::kSStride + ui];
                    an = s_gate[(2 * 32 + u) * L::kSStride + ui];
                }
                const float r = sigmoid_f(gr[e] + ar + b_r[e]);
                const float z = sigmoid_f(gz[e] + az + b_z[e]);
                const float n = tanh_f(gn[e] + r * (an + b_n[e]));
                hn[e] = (1.f - z) * n + z * hprev[e];
                hprev[e] = hn[e];
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                __half h0, l0, h1, l1;
                split_f16(hn[e], h0, l0);
                split_f16(hn[e + 1], h1, l1);
                hi2[e >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                lo2[e >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            const uint4 vhi = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
            const uint4 vlo = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
            if (s + 1 < T) {
                const uint32_t dst = sbase + L::kOffH + nxt * 2 * L::kHBytes + chunk_off;
                const uint32_t bar = smem_u32(&h_full[nxt]);
#pragma unroll
                for (int c = 0; c < kGtCluster; ++c) {       // (4) each store completes its 16 bytes on h_full of CTA c
                    const uint32_t ra = map_to_cta(dst, (uint32_t)c), rb = map_to_cta(bar, (uint32_t)c);
                    st_async_v4(ra, vhi, rb);
                    st_async_v4(ra + L::kHBytes, vlo, rb);
                }
            }
            if (uvalid) {
                const int64_t o = ((int64_t)ubb * T + t) * 512 + dir * 256 + j0 + 8 * ug;
                float4* yo = reinterpret_cast<float4*>(y + o);
                yo[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                yo[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                if (y_hi) {
                    *reinterpret_cast<uint4*>(y_hi + o) = vhi;
                    *reinterpret_cast<uint4*>(y_lo + o) = vlo;
                }
            }
        }
    }
    // no CTA leaves while a peer might still signal it
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Chunk-pipelined variant for large slices: the NB = 32 * NCH utterances of a cluster are NCH independent chains of
// 32 utterances.  Each chain has its own accumulator columns, its own gate staging buffer, its own `mma_done` and
// `h_full` barriers and its own four warps (read-back by the first three - TMEM lane quadrants 0..2 - then the update
// of 32 utterances x 32 units, the DSMEM pushes and the release-arrives); a ninth/fifth warp only issues MMAs.  A
// chain's MMAs of step s+1 need nothing but that chain's rows of the hidden state, so while one chain is in its
// update / push phase (CUDA cores, DSMEM) the other one's MMAs run on the tensor core: the phases of a step overlap
// instead of adding up.  Nothing inside the loop synchronises more than the 128 threads of a chain (named barrier).
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
struct PpLayout {
    static constexpr int kNB = 32 * NCH;
    static constexpr int kThreads = 128 * NCH + 32;
    static constexpr int kHBytes = kNB * 256 * 2;                // one of (hi, lo) of one buffer: 4 K-blocks of NB rows
    static constexpr int kOffH = 0;                              // [2 buffers][hi, lo]
    static constexpr int kGateStride = 33;
    static constexpr int kGateBytes = 3 * 32 * kGateStride * 4;  // per chain: [3 gates][32 units][32 utterances + 1]
    static constexpr int kOffS = kOffH + 4 * kHBytes;
    static constexpr int kOffBar = (kOffS + NCH * kGateBytes + 15) & ~15;
    static constexpr int kSmemBytes = kOffBar + 8 * (3 * NCH) + 16 + 1024;
    static_assert(kHBytes % 1024 == 0, "B operand K-blocks must stay 1024-byte aligned");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int NCH>
__global__ void __cluster_dims__(kGtCluster, 1, 1) __launch_bounds__(PpLayout<NCH>::kThreads, 1)
    gru_layer_pp_kernel(const __half* __restrict__ w_hi,               // [2 dirs][8 ranks][96 rows][256] fp16
                        const __half* __restrict__ w_lo,
                        const float* __restrict__ gi,                  // [B*T, 1536]
                        const float* __restrict__ bhh,                 // [2][768]
                        float* __restrict__ y,                         // [B, T, 512]
                        __half* __restrict__ y_hi, __half* __restrict__ y_lo, int B, int T) {
    using L = PpLayout<NCH>;
    constexpr int NB = L::kNB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + L::kOffBar);      // [NCH]
    uint64_t* h_full = mma_done + NCH;                                         // [2 buffers][NCH]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_full + 2 * NCH);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % kGtCluster;
    const int slice = blockIdx.x / kGtCluster;
    const int dir = blockIdx.y;
    const int j0 = rank * kGtUnits, b0 = slice * NB;
    const bool issuer = uniform_warp_idx() == 4 * NCH;       // provably warp-uniform: the MMA operands stay in uniform registers

    if (tid == 0) {
        for (int c = 0; c < NCH; ++c) {
            mbar_init(&mma_done[c], 1);
            mbar_init(&h_full[c], 1);                        // the issuer's arrive.expect_tx; the pushes complete the bytes
            mbar_init(&h_full[NCH + c], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    // resident weights -> tensor memory: thread t < 128 owns TMEM lane t = gate row t of this CTA (zeros beyond 96)
    if (warp < 4) {
        const int row = tid;
        const size_t src = ((size_t)(dir * kGtCluster + rank) * kGtWRows + row) * 256;      // halves
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const uint4* g = reinterpret_cast<const uint4*>((part ? w_lo : w_hi) + src);
#pragma unroll 1
            for (int c = 0; c < 4; c += 2) {                 // 2 x 32 columns per pass: 16 loads of 16 bytes in flight
                uint32_t r[2][32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint4 v = row < kGtWRows ? __ldg(g + c * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
                    r[i >> 3][4 * (i & 7) + 0] = v.x;
                    r[i >> 3][4 * (i & 7) + 1] = v.y;
                    r[i >> 3][4 * (i & 7) + 2] = v.z;
                    r[i >> 3][4 * (i & 7) + 3] = v.w;
                }
                tmem_st_32x32(lane_addr + (part ? kGtColWlo : 0u) + (uint32_t)(c * 32), r[0]);
                tmem_st_32x32(lane_addr + (part ? kGtColWlo : 0u) + (uint32_t)(c * 32 + 32), r[1]);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // everybody's barriers/TMEM are set up before any peer may push into this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");

    if (issuer) {
        // ---- MMA issue: chain after chain, step after step; each chain's MMAs start when ITS rows are complete.  All
        // 32 lanes walk the loop (uniform control flow and operands), one elected lane issues.
        constexpr uint32_t idesc = make_idesc_f16(128, 32);
        constexpr uint32_t kChainBytes = 32 * 256 * 2 * 2;   // a chain's rows of the hidden state, hi + lo: what 8 CTAs push
        if (T > 1 && elect_one_sync())
            for (int c = 0; c < NCH; ++c) mbar_arrive_expect_tx(&h_full[NCH + c], kChainBytes);   // the pushes of step 0
        __syncwarp();
        for (int s = 1; s < T; ++s) {
            const int cur = s & 1;
            const uint32_t hb = sbase + L::kOffH + cur * 2 * L::kHBytes;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                mbar_wait_cluster(&h_full[cur * NCH + c], (uint32_t)((s - 1) >> 1) & 1u);   // the chain's pushes of step s-1
                fence_proxy_async_all();                     // the pushes were generic-proxy writes; the MMAs read through the async proxy
                tc_fence_after();
                if (elect_one_sync()) {
                    if (s + 1 < T) mbar_arrive_expect_tx(&h_full[(cur ^ 1) * NCH + c], kChainBytes);   // the pushes of step s
                    const uint32_t d_acc = tmem_base + kGtColAcc + (uint32_t)(c * 32);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t b_hi = make_kmajor_desc<128>(hb + kb * (NB * 128) + c * (32 * 128));
                        const uint64_t b_lo = make_kmajor_desc<128>(hb + L::kHBytes + kb * (NB * 128) + c * (32 * 128));
#pragma unroll
                        for (int k = 0; k < 64; k += 16) {
                            const uint32_t a_hi = tmem_base + (uint32_t)((kb * 64 + k) >> 1);      // 2 halves per column
                            const uint32_t a_lo = a_hi + kGtColWlo;
                            umma_f16_ts(d_acc, a_hi, desc_advance_k(b_hi, k), idesc, (kb | k) ? 1u : 0u);
                            umma_f16_ts(d_acc, a_hi, desc_advance_k(b_lo, k), idesc, 1u);
                            umma_f16_ts(d_acc, a_lo, desc_advance_k(b_hi, k), idesc, 1u);
                        }
                    }
                    umma_commit(&mma_done[c]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- one chain: 4 warps, thread -> (utterance ui of the slice, group of 8 hidden units ug) -----------------
        const int chain = warp >> 2, wq = warp & 3;
        const int ui = tid >> 2, ug = tid & 3, uc = ui & 31;
        const int ubb = b0 + ui;
        const bool uvalid = ubb < B;
        float* s_gate = reinterpret_cast<float*>(smem + L::kOffS + chain * L::kGateBytes);
        float b_r[8], b_z[8], b_n[8], hprev[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int u = j0 + 8 * ug + e;
            b_r[e] = __ldg(bhh + dir * 768 + u);
            b_z[e] = __ldg(bhh + dir * 768 + 256 + u);
            b_n[e] = __ldg(bhh + dir * 768 + 512 + u);
            hprev[e] = 0.f;
        }
        // where this thread's 16-byte chunk (8 units of utterance ui) sits inside a swizzled B-operand buffer
        uint32_t chunk_off = 0;
        {
            const int k = j0 + 8 * ug, kb = k >> 6, chunk = (k & 63) >> 3;
            chunk_off = (uint32_t)(kb * (NB * 128) + (ui >> 3) * 1024 + (ui & 7) * 128 + ((chunk ^ (ui & 7)) << 4));
        }
        // gate pre-activations of the input projection, requested ONE STEP AHEAD: the loads of step s + 1 fly during the whole
        // of step s (issued at the top of the step they belong to, they were still in flight when the MMAs had finished)
        const float* gbase = gi + (int64_t)(uvalid ? ubb : B - 1) * T * 1536 + dir * 768 + j0 + 8 * ug;
        float4 gnext[3][2];
        {
            const float4* gp = reinterpret_cast<const float4*>(gbase + (int64_t)(dir == 0 ? 0 : T - 1) * 1536);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                gnext[g][0] = __ldg(gp + g * 64);
                gnext[g][1] = __ldg(gp + g * 64 + 1);
            }
        }
        for (int s = 0; s < T; ++s) {
            const int t = dir == 0 ? s : T - 1 - s;
            const int nxt = (s & 1) ^ 1;
            float4 gin[3][2];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                gin[g][0] = gnext[g][0];
                gin[g][1] = gnext[g][1];
            }
            if (s + 1 < T) {
                const float4* gp = reinterpret_cast<const float4*>(gbase + (int64_t)(dir == 0 ? s + 1 : T - 2 - s) * 1536);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    gnext[g][0] = __ldg(gp + g * 64);
                    gnext[g][1] = __ldg(gp + g * 64 + 1);
                }
            }
            if (s > 0) {
                // accumulator rows of this chain -> its staging buffer, transposed to [gate][unit][utterance]
                if (wq < 3) {
                    mbar_wait(&mma_done[chain], (uint32_t)(s - 1) & 1u);
                    tc_fence_after();
                    float v[32];
                    tmem_ld_32x32(tmem_base + kGtColAcc + (uint32_t)(chain * 32) + ((uint32_t)(wq * 32) << 16), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) s_gate[(wq * 32 + lane) * L::kGateStride + i] = v[i];
                    tc_fence_before();
                }
                named_barrier_sync(1 + chain, 128);
            }
            const float gr[8] = {gin[0][0].x, gin[0][0].y, gin[0][0].z, gin[0][0].w,
                                 gin[0][1].x, gin[0][1].y, gin[0][1].z, gin[0][1].w};
            const float gz[8] = {gin[1][0].x, gin[1][0].y, gin[1][0].z, gin[1][0].w,
                                 gin[1][1].x, gin[1][1].y, gin[1][1].z, gin[1][1].w};
            const float gn[8] = {gin[2][0].x, gin[2][0].y, gin[2][0].z, gin[2][0].w,
                                 gin[2][1].x, gin[2][1].y, gin[2][1].z, gin[2][1].w};
            float hn[8];
            uint32_t hi2[4], lo2[4];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float ar = 0.f, az = 0.f, an = 0.f;
                if (s > 0) {
                    const int u = 8 * ug + e;
                    ar = s_gate[(0 * 32 + u) * L::kGateStride + uc];
                    az = s_gate[(1 * 32 + u) * L::kGateStride + uc];
                    an = s_gate[(2 * 32 + u) * L::kGateStride + uc];
                }
                const float r = sigmoid_f(gr[e] + ar + b_r[e]);
                const float z = sigmoid_f(gz[e] + az + b_z[e]);
                const float n = tanh_f(gn[e] + r * (an + b_n[e]));
                hn[e] = (1.f - z) * n + z * hprev[e];
                hprev[e] = hn[e];
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                __half h0, l0, h1, l1;
                split_f16(hn[e], h0, l0);
                split_f16(hn[e + 1], h1, l1);
                hi2[e >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                lo2[e >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            const uint4 vhi = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
            const uint4 vlo = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
            if (s + 1 < T) {
                const uint32_t dst = sbase + L::kOffH + nxt * 2 * L::kHBytes + chunk_off;
                const uint32_t bar = smem_u32(&h_full[nxt * NCH + chain]);
#pragma unroll
                for (int c = 0; c < kGtCluster; ++c) {       // each store completes its 16 bytes on the CHAIN's barrier of CTA c
                    const uint32_t ra = map_to_cta(dst, (uint32_t)c), rb = map_to_cta(bar, (uint32_t)c);
                    st_async_v4(ra, vhi, rb);
                    st_async_v4(ra + L::kHBytes, vlo, rb);
                }
            }
            if (uvalid) {
                const int64_t o = ((int64_t)ubb * T + t) * 512 + dir * 256 + j0 + 8 * ug;
                float4* yo = reinterpret_cast<float4*>(y + o);
                yo[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                yo[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                if (y_hi) {
                    *reinterpret_cast<uint4*>(y_hi + o) = vhi;
                    *reinterpret_cast<uint4*>(y_lo + o) = vlo;
                }
            }
        }
    }
    // no CTA leaves while a peer might still signal it
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int NCH>
static int launch_gru_pp(const __half* w_hi, const __half* w_lo, const float* gi, const float* bhh, float* y,
                         __half* y_hi, __half* y_lo, int B, int T, cudaStream_t st) {
    using L = PpLayout<NCH>;
    SIR_SMEM_OPTIN(gru_layer_pp_kernel<NCH>, L::kSmemBytes);
    dim3 grid((unsigned)(kGtCluster * ((B + L::kNB - 1) / L::kNB)), 2);
    gru_layer_pp_kernel<NCH><<<grid, L::kThreads, L::kSmemBytes, st>>>(w_hi, w_lo, gi, bhh, y, y_hi, y_lo, B, T);
    SIR_CHECK_LAUNCH("gru_layer_pp_kernel");
    return SIR_OK;
}

template <int NB>
static int launch_gru(const __half* w_hi, const __half* w_lo, const float* gi, const float* bhh, float* y,
                      __half* y_hi, __half* y_lo, int B, int T, cudaStream_t st) {
    using L = GtLayout<NB>;
    SIR_SMEM_OPTIN(gru_layer_tc_kernel<NB>, L::kSmemBytes);
    dim3 grid((unsigned)(kGtCluster * ((B + NB - 1) / NB)), 2);
    gru_layer_tc_kernel<NB><<<grid, kGtThreads, L::kSmemBytes, st>>>(w_hi, w_lo, gi, bhh, y, y_hi, y_lo, B, T);
    SIR_CHECK_LAUNCH("gru_layer_tc_kernel");
    return SIR_OK;
}

// Utterances per cluster.  Up to 112 utterances: slices of 16 (the smallest UMMA N), both directions of the whole
// batch in one wave of <= 14 clusters - the shortest step for a batch that is served alone (training at batch 16,
// single-file inference).  Beyond that: slices of 64 as two 32-utterance chains (gru_layer_pp_kernel) - 8 SMs per 64
// utterances (256 utterances: 64 SMs; the model chunks its batch to 336 = one wave of 2 x 6 clusters), a step of 3.6 us,
// and the SMs it leaves take the frontend / conv stack of the next batch, which runs on another stream.
constexpr int kGtMaxClustersPerWave = 15;     // co-resident 8-CTA clusters on a B200 (measured launch__cluster_max_active)

int gru_layer_tc(const __half* w_hi, const __half* w_lo, const float* gi, const float* bhh, float* y,
                 __half* y_hi, __half* y_lo, int B, int T, cudaStream_t st) {
    if (2 * ((B + 15) / 16) <= kGtMaxClustersPerWave) return launch_gru<16>(w_hi, w_lo, gi, bhh, y, y_hi, y_lo, B, T, st);
    return launch_gru_pp<2>(w_hi, w_lo, gi, bhh, y, y_hi, y_lo, B, T, st);
}

}  // namespace tc
}  // namespace sir
