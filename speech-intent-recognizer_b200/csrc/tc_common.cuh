// Raw-PTX building blocks for the sm_100a tensor-core path: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (TMEM allocation, UMMA issue/commit, TMEM loads), UMMA shared-memory / instruction descriptors, and the
// fp16 hi/lo operand split used for near-fp32 contractions on the f16 tensor pipe.
//
// Numerics: every contraction of the classifier is computed as
//     A.B  ~=  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi ,   x_hi = fp16(x),  x_lo = fp16(x - x_hi)
// with fp32 accumulation in TMEM.  The dropped A_lo.B_lo term is ~2^-22 relative, so logits stay within
// ~1e-4 of the fp32 reference (a single fp16/bf16/tf32 pass misses the 1e-3 parity bar by 10x - measured on
// the CPU restatement).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace sir {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp-uniform role dispatch ------------------------------------------------------------------------------
// tcgen05.mma / commit and TMA take warp-uniform operands.  Inside `if (lane == 0)` the compiler cannot prove
// uniformity and wraps every issue in an R2UR + vote "waterfall" loop (~10 SASS instructions per MMA, measured to
// bound the issue rate of the single MMA thread).  Making the warp index provably uniform (shuffle from lane 0) and
// electing the issuing lane with elect.sync right at the issue site keeps descriptors in uniform registers.
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all state spaces: also orders generic-proxy stores into a PEER CTA's shared memory (st.shared::cluster)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe of a phase (a warp that serves several barriers polls them in turn).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Two 32-column TMEM loads in flight, one wait (the second load does not queue behind the first one's wait).
__device__ __forceinline__ void tmem_ld_32x32_pair(uint32_t taddr0, uint32_t taddr1, float (&v)[32], float (&u)[32]);
// Cluster-scope variants: the waiter acquires what threads of PEER CTAs released with mbar_arrive_remote.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait_cluster(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
// `remote_bar` is a shared::cluster address (mapa) of an mbarrier in a peer CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a CUDA error (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}

// ---- dynamic tile scheduler of the persistent kernels --------------------------------------------------------
// The producer thread of a CTA DRAWS tile indices from a global ticket counter (the next draw is in flight while the
// current tile loads) and publishes them through a small shared-memory ring that every consumer warp (MMA issuer,
// epilogue warps) reads in the same order.  With a static round-robin a CTA that starts late - its SM was busy with
// another stream's kernel, e.g. the GRU recurrence of the previous batch - still owns 1/grid of the tiles and the
// SMs that started on time idle until it is done; with tickets the CTAs that run take the tiles.
// The counter is never reset: every producer draws (its tiles + 1) tickets, so a launch consumes exactly
// num_tiles + gridDim.x of them and the host passes the launch's first ticket.  counter == nullptr: static order.
struct TileTickets {
    unsigned long long* counter;
    unsigned long long base;
};
// Host view of a ticket counter: launches that share one must be stream-ordered.  first() is the launch's first
// ticket; consumed() is called once the launch has been accepted (a failed launch draws nothing, so the host's
// idea of the counter must not move either).
// Optional output routing of tc_gemm_nt: rows >= m_split are written to C2 (two parameter blocks from one launch) and the
// product is multiplied by *out_scale (a device float: the inverse of the power-of-two scale of a gradient operand).
struct GemmOutput {
    float* C2 = nullptr;
    int m_split = 0;
    const float* out_scale = nullptr;
};

struct TicketSource {
    unsigned long long* dev = nullptr;
    unsigned long long next = 0;
    TileTickets first() const { return TileTickets{dev, next}; }
    void consumed(long long num_tiles, long long grid) { next += (unsigned long long)(num_tiles + grid); }
};
constexpr int kRingDepth = 4;
struct TileRing {
    uint64_t full[kRingDepth], empty[kRingDepth];
    int tile[kRingDepth];
    int pad[2];
};
__device__ __forceinline__ void ring_init(TileRing* r, uint32_t consumer_warps) {      // one thread, before fence_barrier_init
    for (int i = 0; i < kRingDepth; ++i) {
        mbar_init(&r->full[i], 1);
        mbar_init(&r->empty[i], consumer_warps);
    }
}
struct TileProducer {                                       // lives in the registers of the producer thread
    TileRing* r;
    TileTickets tk;
    int num_tiles;
    uint32_t it = 0, drawn = 0;
    long long next = 0;
    __device__ __forceinline__ long long draw() {
        const long long t = tk.counter ? (long long)(atomicAdd(tk.counter, 1ULL) - tk.base)
                                       : (long long)blockIdx.x + (long long)drawn * gridDim.x;
        ++drawn;
        return t;
    }
    __device__ __forceinline__ TileProducer(TileRing* ring, const TileTickets& t, int n) : r(ring), tk(t), num_tiles(n) {
        next = draw();
    }
    // the next tile of this CTA (-1: none left), already published to the consumers
    __device__ __forceinline__ int pop() {
        const int tile = next < (long long)num_tiles ? (int)next : -1;
        const int slot = it % kRingDepth;
        mbar_wait(&r->empty[slot], ((it / kRingDepth) & 1u) ^ 1u);
        r->tile[slot] = tile;
        mbar_arrive(&r->full[slot]);                        // release: the tile index is visible to whoever acquires `full`
        ++it;
        if (tile >= 0) next = draw();                       // in flight while this tile's loads are issued
        return tile;
    }
};
// consumer side: every lane of the warp calls it (uniform), lane 0 frees the slot
__device__ __forceinline__ int ring_next(TileRing* r, uint32_t& it, int lane) {
    const int slot = it % kRingDepth;
    mbar_wait(&r->full[slot], (it / kRingDepth) & 1u);
    const int tile = r->tile[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&r->empty[slot]);
    ++it;
    return tile;
}

// ---- TMA -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {        // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in TENSOR MEMORY: lane = row of A, 32-bit column c holds elements (2c, 2c+1) of the row;
// one instruction (K = 16) reads 8 columns starting at `tmem_a`.  Only B is fetched from shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32: A and B hold fp32 bit patterns of which the tensor core reads sign, exponent and the 10 leading mantissa bits
// (the 13 low bits are ignored - measured, tools/tf32_probe.cu); K = 8 per instruction (32 bytes of a K-major row).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A in tensor memory: lane = row, one 32-bit column per element; one instruction reads 8 columns starting at `tmem_a`.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: every thread of the warp writes 32 consecutive 32-bit columns of ITS lane (warp quadrant)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// mbarrier arrives when all previously issued UMMAs of this thread have completed (implies fence::before).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = TMEM lane of the warp's quadrant).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_pair(uint32_t taddr0, uint32_t taddr1, float (&v)[32], float (&u)[32]) {
    uint32_t a[32], b[32];
    tmem_ld_32x32_issue(taddr0, a);
    tmem_ld_32x32_issue(taddr1, b);
    // the loaded registers are operands of the wait (and of an empty statement behind it): no use can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]) :: "memory");
    asm volatile("" : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]), "+r"(b[16]), "+r"(b[17]), "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]), "+r"(b[24]), "+r"(b[25]), "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31]) :: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(a[i]);
        u[i] = __uint_as_float(b[i]);
    }
}

// ---- descriptors -------------------------------------------------------------------------------------------
// K-major operand tile, rows of SWIZZLE bytes (64 or 128) = one swizzle atom along K, 8-row atoms stacked
// along M/N every 8*SWIZZLE bytes.  Bits: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
template <int SWIZZLE>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    static_assert(SWIZZLE == 64 || SWIZZLE == 128, "swizzle span");
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(((8u * SWIZZLE) >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SWIZZLE == 128 ? 2 : 4) << 61;
    return d;
}
// Advancing by k fp16 elements inside the swizzle atom moves the start address by 2k bytes.
__device__ __forceinline__ uint64_t desc_advance_k(uint64_t desc, uint32_t k_elems) {
    return desc + (uint64_t)((k_elems * 2u) >> 4);
}
// kind::f16 instruction descriptor: D = F32 [4,6)=1, A/B = F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::tf32 instruction descriptor: D = F32, A/B = TF32 (2) at [7,10) / [10,13), both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- fp16 hi/lo split ---------------------------------------------------------------------------------------
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);      // fp16 range; activations and weights of this model are O(1..100)
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// ---- pooled conv epilogue -----------------------------------------------------------------------------------
// One warp holds a 2 x 16 (or 4 x 8) pixel patch of a conv tile (lane = dy * width + x) and, per lane, 32 consecutive
// output channels v[0..32) of its pixel.  2x2 max-pool partners are lanes ^1 (x) and ^16 (or ^8) (y).  Instead of reducing all
// 32 channels in all four lanes, the lanes of a pooling group split the channels while they reduce: after the x
// step each lane owns 16 channels, after the y step 8 - 24 shuffles per lane instead of 64, and every lane ends
// with 8 DISTINCT pooled channels: + shift, ReLU, fp16 hi/lo split, one 16-byte store each to hi and lo.
// Returns the first channel (relative to the chunk) this lane owns.
template <int YBIT = 16>     // lane bit that separates the two pixel rows of a pooling window (tile width 16 or 8)
__device__ __forceinline__ int pool2x2_split_channels(const float (&v)[32], int lane, float (&out)[8]) {
    const bool sx = lane & 1, sy = (lane & YBIT) != 0;
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float keep = sx ? v[i + 16] : v[i];
        const float send = sx ? v[i] : v[i + 16];
        m[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = sy ? m[j + 8] : m[j];
        const float send = sy ? m[j] : m[j + 8];
        out[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, YBIT));
    }
    return (sx ? 16 : 0) + (sy ? 8 : 0);
}

// out[j] + shift[j] -> ReLU -> fp16 (hi, lo) -> dst_hi[0..8), dst_lo[0..8)  (16-byte aligned destinations)
__device__ __forceinline__ void shift_relu_split_store8(const float (&o)[8], const float* __restrict__ shift8,
                                                        __half* __restrict__ dst_hi, __half* __restrict__ dst_lo) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(shift8) + 1);
    const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t hi2[4], lo2[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const float a = fminf(fmaxf(o[j] + sh[j], 0.f), 65504.f), b = fminf(fmaxf(o[j + 1] + sh[j + 1], 0.f), 65504.f);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        hi2[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        lo2[j >> 1] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(dst_hi) = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
    *reinterpret_cast<uint4*>(dst_lo) = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
}

}  // namespace tc
}  // namespace sir
