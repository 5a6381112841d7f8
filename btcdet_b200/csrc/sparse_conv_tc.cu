// tcgen05 (5th-gen tensor core) gather-GEMM tile for the sparse-conv layers.
//
// Same output-stationary formulation as sparse_conv.cu (128 output rows per tile, the K offsets walked in ascending
// order, every output row written once, no atomics), with the products on the tensor cores and the accumulator in
// TMEM.  The reduction axis is the flattened (offset, input channel) index cut into stages of 32:
//
//   * producer warps gather the neighbour rows named by nbr_out[o][k] (index tile TMA-staged in shared memory) with
//     cp.async (8 lanes per row, zero fill for missing neighbours), read them back one row per thread, split every
//     fp32 value into a tf32 "hi" part (low 13 mantissa bits cleared) and the exact fp32 remainder "lo", and write the
//     A operand straight into tensor memory with tcgen05.st;
//   * the matching weight slice comes pre-packed (btc_sparse_conv_tc_pack) as the shared-memory image of the K-major,
//     128-byte-swizzled hi/lo tiles; a dedicated loader warp fetches it with one cp.async.bulk per stage (mbarrier
//     complete_tx) into its own ring, NB stages ahead of the MMAs;
//   * one elected lane issues tcgen05.mma.kind::tf32 (M=128, N, K=8), A from TMEM, B from shared memory — 3xTF32:
//     D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi — fp32-class accuracy (the 1e-4 parity bar) from the tf32 pipe;
//     tcgen05.commit releases the stage / publishes the accumulator through mbarriers;
//   * epilogue warps read the (double-buffered) accumulator back with tcgen05.ld, apply bias / folded-BN affine /
//     ReLU and write the output rows while the next tile's main loop runs.
#include <stdlib.h>
#include <atomic>
#include <cuda_bf16.h>
#include "common.cuh"

namespace btc {

constexpr int TC_BM = 128;      // output rows per CTA == UMMA M
constexpr int TC_KC = 32;       // input channels per stage (32 tf32 = 128 B = one swizzle row)

// ---- raw PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // one non-blocking probe
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// raw shared-window addresses (computed once per role: &bar -> cvta + cluster-window math is not free in a hot loop)
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
// explicit shared-space loads: pointers derived from the aligned dynamic-smem base lose their address space and
// compile to generic LD.E (long-scoreboard latency, plus padding instructions in front of every LDGSTS)
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128).
// Rows are 128 B; 8-row groups (1024 B) are SBO apart; the tile base is 1024-B aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;              // LBO (unused for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;    // SBO = 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}

// K-major, SWIZZLE_64B descriptor (bf16 tiles of 32 reduction elements: rows of 64 B, 8-row atoms 512 B apart, layout = 4)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// byte offset of element (row r, float j) inside a K-major SW128 tile
__host__ __device__ __forceinline__ int sw128_offset(int r, int j) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((j >> 2) ^ (r & 7)) & 7) << 4) + (j & 3) * 4;
}

// ---- weight pre-pack ------------------------------------------------------------------------------
// The reduction axis is the flattened (offset k, input channel ci) index e = k*c_in + ci, cut into chunks of 32:
// thin layers pack several offsets into one MMA stage (c_in = 4: eight offsets per stage, 4 stages per tile).
// packed[chunk] = { B_hi image [N x 128 B] , B_lo image [N x 128 B] },  B[n][j] = W[e/c_in][e%c_in][n], e = 32*chunk + j
__global__ void tc_pack_weight_kernel(const float* __restrict__ w, int K, int c_in, int c_out, int N,
                                      float* __restrict__ packed) {
    const int E = K * c_in;
    const int nchunk = (E + TC_KC - 1) / TC_KC;
    const int64_t total = (int64_t)nchunk * N * TC_KC;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(t % TC_KC);
        int64_t r = t / TC_KC;
        int n = (int)(r % N);
        int cc = (int)(r / N);
        int e = cc * TC_KC + j;
        float v = (e < E && n < c_out) ? w[(int64_t)e * c_out + n] : 0.f;   // w[k][ci][n] is e-major already
        float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        float lo = v - hi;
        char* base = (char*)packed + (int64_t)cc * (2 * N * 128);
        *(float*)(base + sw128_offset(n, j)) = hi;
        *(float*)(base + N * 128 + sw128_offset(n, j)) = lo;
    }
}

// byte offset of bf16 element (row r, element j of 32) inside a K-major SWIZZLE_64B tile: rows of 64 B, 8-row atoms of
// 512 B, the 16-byte chunk index XORed with bits [7,9) of the byte address (cute Swizzle<2,4,3>) = (r >> 1) & 3
__host__ __device__ __forceinline__ int sw64_offset_bf16(int r, int j) {
    return (r >> 3) * 512 + (r & 7) * 64 + ((((j >> 3) ^ (r >> 1)) & 3) << 4) + (j & 7) * 2;
}

// Split-format weights: the reduction axis e = k*c_in + ci in chunks of 32 (one MMA stage); packed[chunk] = { B_hi image
// [N x 64 B] of bf16, B_lo image }, hi = bf16_rn(w), lo = bf16_rn(w - hi) — 64-byte rows, so a stage fetches exactly the
// 2 * N * 64 bytes it multiplies (a 128-byte-row tile of 64 elements was fetched whole by both of its stages).
__global__ void tc_pack_weight_split_kernel(const float* __restrict__ w, int K, int c_in, int c_out, int N,
                                            unsigned short* __restrict__ packed) {
    const int E = K * c_in;
    const int nchunk = (E + TC_KC - 1) / TC_KC;
    const int64_t total = (int64_t)nchunk * N * TC_KC;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % TC_KC);
        const int64_t r = t / TC_KC;
        const int n = (int)(r % N);
        const int cc = (int)(r / N);
        const int e = cc * TC_KC + j;
        const float v = (e < E && n < c_out) ? w[(int64_t)e * c_out + n] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        char* base = (char*)packed + (int64_t)cc * (2 * N * 64);
        *(unsigned short*)(base + sw64_offset_bf16(n, j)) = __bfloat16_as_ushort(hi);
        *(unsigned short*)(base + N * 64 + sw64_offset_bf16(n, j)) = __bfloat16_as_ushort(lo);
    }
}

// fp32 rows <-> split rows (per 32 channels [32 x bf16 hi | 32 x bf16 lo]); c % 32 == 0.  One thread per channel pair.
__global__ void features_to_split_kernel(const float* __restrict__ in, int n_cap, const int* __restrict__ n_dev, int c,
                                         uint32_t* __restrict__ out) {
    const int64_t work = (int64_t)live_count(n_cap, n_dev) * (c / 2);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / (c / 2);
        const int pr = (int)(t % (c / 2));                 // channel pair (2 pr, 2 pr + 1)
        const float2 x = *reinterpret_cast<const float2*>(in + row * c + 2 * pr);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x.y - __bfloat162float(h1));
        uint32_t* blk = out + row * c + (pr >> 4) * 32;    // 32 words per 32-channel block
        blk[pr & 15] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        blk[16 + (pr & 15)] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
}
__global__ void features_from_split_kernel(const uint32_t* __restrict__ in, int n_cap, const int* __restrict__ n_dev, int c,
                                           float* __restrict__ out) {
    const int64_t work = (int64_t)live_count(n_cap, n_dev) * (c / 2);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / (c / 2);
        const int pr = (int)(t % (c / 2));
        const uint32_t* blk = in + row * c + (pr >> 4) * 32;
        const uint32_t h = blk[pr & 15], l = blk[16 + (pr & 15)];
        float2 x;
        x.x = __uint_as_float(h << 16) + __uint_as_float(l << 16);
        x.y = __uint_as_float(h & 0xFFFF0000u) + __uint_as_float(l & 0xFFFF0000u);
        *reinterpret_cast<float2*>(out + row * c + 2 * pr) = x;
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------
// Persistent, warp-specialised: one CTA per SM walks 128-row output tiles handed out by a tile scheduler.
//   producers  : NPW warps in NPW/4 groups (4 warps = 128 rows = the four TMEM lane quarters; thread == output row).
//                Group g owns the stages whose global index is congruent to g modulo the group count.  Per stage:
//                cp.async gather of the neighbour rows (8 lanes per row, zero fill, TcDepth stages in flight) into a
//                swizzled staging slot, read back one row per thread, hi/lo split, tcgen05.st STRAIGHT INTO TENSOR
//                MEMORY — the MMA reads A from TMEM (".ts" form), so shared memory only carries the weight tiles.
//                (An SS-form 3xTF32 step reads 18 KB of smem per 8-deep k-step and is smem-bandwidth bound at ~70
//                cycles per MMA.)
//   MMA issuer : one warp in warp-uniform control flow, one ELECTed lane issues; two stages per wait/fence/elect trip;
//                one tcgen05.commit per stage (or per commit group) frees the A slot and the weight slot together;
//   epilogue   : four warps (TMEM lane quarter = warp % 4), overlapping the next tile's main loop through a
//                double-buffered TMEM accumulator; the tile id reaches them through s_epi_tile;
//   index loader / tile scheduler : TMA-stages each tile's [128 x K] block of the neighbour table one tile ahead, builds
//                the tile's list of active 32-element chunks (block skipping) and fetches tiles statically or from a
//                global counter; a negative tile id is the end marker every role leaves on;
//   weight loader : one cp.async.bulk per stage into the weight ring, as soon as the slot's previous MMAs have retired.
// TMEM map (512 columns): [0, 2*ACC) two accumulators (ACC = N, or 2N with concatenated B) | then per stage s 64 columns:
// [+0, +32) A_hi, [+32, +64) A_lo.  The protocol is modelled in tests/test_tc_protocol_model.py.
// Warp roles for NPW producer warps (8 or 16): [0, NPW) producers in NPW/4 groups, NPW = MMA issuer, NPW+1..NPW+4
// epilogue (TMEM lane quarter = warp % 4 covers 1,2,3,0), NPW+5 = index loader + tile scheduler, NPW+6 = weight loader,
// the rest (to a multiple of four warps) idle until the final barrier.
template <int NPW> struct TcRoles {
    static constexpr int kGroups = NPW / 4;
    static constexpr int kMma = NPW, kEpi0 = NPW + 1, kIdx = NPW + 5, kBld = NPW + 6;
    static constexpr int kWarps = ((NPW + 7 + 3) / 4) * 4;
    static constexpr int kThreads = kWarps * 32;
};
// Operand rings.  A: 64 TMEM columns (hi + lo) per stage next to the two accumulators; weights: one packed hi/lo tile
// (2 * N * 128 B) per stage in shared memory, fetched by the loader warp the moment the MMAs that read the slot retire.
// Both rings share the slot index, the phase and ONE tcgen05.commit per stage (empty_bar), so their depth is the same:
// six stages at N = 32 (448 TMEM columns, 48 KB of weight tiles), four at N = 64 / 128 (64 / 128 KB of weight tiles).
// Round-1 timing diagnostics (tools/step_breakdown.py --diag) on the way here: separate, deeper rings (6 x A, 8 x B) and
// a second commit per stage were within 1 % of this; dropping the weight copies altogether changes < 1 %.
// Split format: a stage's weight tile is half the bytes (64-byte rows) and its A operand half the TMEM columns (16 hi +
// 16 lo), so the same shared memory / tensor memory holds rings twice as deep: 12 stages at N = 32, 8 at N = 64 / 128.
template <int N, bool SPLIT> struct TcAStages { static constexpr int value = (N <= 32 ? 6 : 4) * (SPLIT ? 2 : 1); };
template <int N, bool SPLIT> struct TcBStages { static constexpr int value = TcAStages<N, SPLIT>::value; };
template <int N, int NPW> struct TcDepth { static constexpr int value = (N > 64 || NPW > 8) ? 2 : 4; };   // cp.async gather stages in flight per producer warp (smem budget)

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// bf16 operands (A: 16 values = 8 TMEM columns per lane, two per column, even element in the low half)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// warp-converged election of one lane (all 32 lanes must execute this)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Operand formats (round 2).  SIN / SOUT select the "split" feature format on the input / output side:
//   fp32 format  : rows of C floats; the producers split every value into tf32 hi / lo in registers and the k-step is
//                  three kind::tf32 MMAs of K = 8 (3xTF32) — 12 MMAs per 32-element stage;
//   split format : the SAME 4 C bytes per row, but every value is stored by the producing layer's epilogue as two
//                  bfloat16: x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi) (|x - hi - lo| <= 2^-17 |x|).  Per
//                  32 channels the row holds [32 x hi | 32 x lo] (64 B + 64 B), so a stage's gather is byte-for-byte
//                  the copy it was, the read-back registers ARE the packed A operands (no ALU work in the producers,
//                  half the tcgen05.st) and the k-step is three kind::f16 (bf16) MMAs of K = 16: A_lo B_hi + A_hi B_lo
//                  + A_hi B_hi — 6 MMAs per stage.  Weights are packed per 32-element chunk as bf16 hi / lo tiles with
//                  64-byte rows (SWIZZLE_64B), so a stage fetches exactly the bytes it multiplies.
// Both meet the 1e-4 parity bar (split: ~2^-16 per product, measured in tests/test_parity_gpu.py).  The experimental
// variants of round 1 (concatenated [B_hi|B_lo] MMAs, commit groups, programmatic dependent launch) were measured on
// hardware at the start of round 2 (profiles/r2_battery.json: no gain / slower) and removed.
template <int N, int NPW, bool SIN, bool SOUT>
__global__ void __launch_bounds__(TcRoles<NPW>::kThreads, 1)
conv_fwd_tc_kernel(const float* __restrict__ feat_in, const int* __restrict__ table,
                   const float* __restrict__ packed_w, const float* __restrict__ bias,
                   const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                   float* __restrict__ feat_out, const int* __restrict__ out_rows, int n_cap,
                   const int* __restrict__ n_dev, int K, int c_in, int c_out, int* __restrict__ tile_ctr, int diag,
                   unsigned long long* __restrict__ trace, const unsigned long long* __restrict__ tile_mask,
                   const int* __restrict__ tile_order, int tiles_cap) {
    constexpr int STAGES = TcAStages<N, SIN>::value;
    constexpr int TC_DEPTH = TcDepth<N, NPW>::value;
    using Roles = TcRoles<NPW>;
    constexpr int G = Roles::kGroups;               // producer groups; group g feeds the stages with gi % G == g
    constexpr int TC_PRODUCER_WARPS = NPW, TC_MMA_WARP = Roles::kMma, TC_IDX_WARP = Roles::kIdx, TC_BLD_WARP = Roles::kBld;
    constexpr int NB = TcBStages<N, SIN>::value;    // weight-tile ring depth
    static_assert(NB == STAGES, "the A ring and the weight ring share slot index, phase and the commit");
    static_assert(G <= STAGES, "a group advances by G stages and may wrap the ring at most once per step");
    constexpr int B_BYTES = N * (SIN ? 64 : 128); // one B tile (hi or lo): K-major SW128 (tf32) / SW64 (bf16), 32 elements per row
    constexpr int STAGE_BYTES = 2 * B_BYTES;
    constexpr uint32_t A_STRIDE = SIN ? 32u : 64u; // TMEM columns of one A stage (hi + lo)
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t ACC_COLS = N;              // TMEM columns of one accumulator buffer
    constexpr uint32_t A_COL0 = 2 * ACC_COLS;     // first A-operand column
    // instruction descriptor: D=f32 (1<<4), A / B format at bits 7 / 10 (tf32 = 2, bf16 = 1), K-major both, N>>3 at bit 17,
    // M>>4 at bit 24
    constexpr uint32_t FMT = SIN ? 1u : 2u;
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    static_assert(2 * ACC_COLS + A_STRIDE * STAGES <= 512, "TMEM budget");

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* stages = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1 KB aligned
    unsigned char* a_stage = stages + NB * STAGE_BYTES;                 // [8 warps][TC_DEPTH][32 rows x 128 B]
    int* nbr_s = (int*)(a_stage + TC_PRODUCER_WARPS * TC_DEPTH * 4096);     // [2][TC_BM * K]

    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2], nbr_full[2], nbr_empty[2],
        list_full[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_cnt[2];                        // active reduction chunks of the tile in each index buffer
    __shared__ int s_tile[2];                       // tile id in each index buffer (-1: no more tiles for this CTA)
    __shared__ int s_epi_tile[2];                   // tile id behind each accumulator buffer (MMA issuer -> epilogue)
    __shared__ int s_cls_start[66];                 // heaviest-first prefix of the cost-class counts (tile_order lookup)
    // epilogue constants per output column, staged once per CTA: y = (acc + bias) * scale + shift (absent -> 0 / 1 / 0);
    // a __ldg per column and row in the epilogue loop cost ~140 cycles per column (exposed L1 latency, round-2 timeline)
    __shared__ __align__(16) float s_ep_bias[N], s_ep_scale[N], s_ep_shift[N];

    const int n = live_count(n_cap, n_dev);
    if ((int)blockIdx.x * TC_BM >= n) {           // no tile for this CTA (whole CTA leaves before any barrier)
        if (tile_ctr && threadIdx.x == 0 && atomicAdd(tile_ctr + 1, 1) == (int)gridDim.x - 1) {   // see the scheduler warp
            atomicExch(tile_ctr, 0);
            atomicExch(tile_ctr + 1, 0);
        }
        return;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Timeline diagnostics (btc_sparse_conv_tc_trace): 32 u64 slots per CTA, see tools/tc_timeline.py for the legend.
    unsigned long long* tr = trace ? trace + (size_t)blockIdx.x * 32 : nullptr;
    if (tr && tid == 0) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        tr[0] = g;
        tr[1] = (unsigned long long)clock64();
    }
    const int num_tiles = (n + TC_BM - 1) / TC_BM;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int E = K * c_in;                         // flattened (offset, channel) reduction length
    const int T = (E + TC_KC - 1) / TC_KC;          // reduction chunks (32 elements each) of a full tile
    // Block skipping: a chunk whose offsets have no valid neighbour in any of the tile's 128 rows contributes exact
    // zeros, so neither the gather nor the MMAs are issued for it.  The index loader publishes, per tile, the ordered
    // list of active chunks; producers, weight loads and the MMA issuer all walk that list (stage counts per tile vary).
    unsigned short* s_list = reinterpret_cast<unsigned short*>(nbr_s + 2 * TC_BM * K);   // [2][T]

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 128 + 1);      // the 128 threads of one producer group + the weight loader (+ its tx bytes)
            mbar_init(&empty_bar[s], 1);           // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 2);           // the issuer's early arrive (publishes the tile id) + one tcgen05.commit
            mbar_init(&tmem_empty[b], 128);        // the 128 epilogue threads
            mbar_init(&nbr_full[b], 1);            // the index loader (+ tx bytes)
            mbar_init(&nbr_empty[b], NPW * 32 + 2);   // every producer thread + the MMA issuer + the weight loader
            mbar_init(&list_full[b], 1);           // the index loader, after it built the chunk list
        }
        fence_mbar_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc(&s_tmem, TMEM_COLS);
    for (int c = tid; c < N; c += Roles::kThreads) {
        const bool in = c < c_out;
        s_ep_bias[c] = (bias && in) ? __ldg(bias + c) : 0.f;
        s_ep_scale[c] = (scale && in) ? __ldg(scale + c) : 1.f;
        s_ep_shift[c] = (scale && in) ? __ldg(shift + c) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (tr && tid == 0) tr[2] = (unsigned long long)clock64();

    if (warp < TC_PRODUCER_WARPS) {
        // ================= producers =================
        // Each warp owns 32 output rows of the tile end to end.  Global -> shared: cp.async, 8 lanes per row so a
        // warp request touches 4 cache lines (a row-per-thread gather would touch 32 and serialise in L1);
        // invalid neighbours are zero-filled by the copy itself.  Shared -> registers: one row per thread (the
        // layout tcgen05.st wants), conflict-free through a 16-byte-chunk XOR swizzle.  TC_DEPTH stages in flight.
        const int group = warp >> 2, quarter = warp & 3;
        const int sub = lane >> 3, q = lane & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t abuf = smem_u32(a_stage + warp * (TC_DEPTH * 4096));
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t nbr_s32 = smem_u32(nbr_s), list_s32 = smem_u32(s_list);
        // Lane constants (the issue path is instruction-bound: keep the per-copy address math to a few ops).
        // Row r = g*4 + sub of the warp's 32 rows goes to r*128 + ((q ^ (r & 7)) << 4); (r & 7) = (g & 1)*4 + sub.
        const uint32_t dst_even = sub * 128 + ((q ^ sub) << 4);
        const uint32_t dst_odd = sub * 128 + ((q ^ (4 + sub)) << 4);
        const uint32_t rowbytes = (uint32_t)c_in * 4u;
        const int strideK4 = 4 * K;                 // index-tile stride between rows r and r + 4
        const char* fbase = reinterpret_cast<const char*>(feat_in);
        // Issue-side iterator over the global stage sequence (tile, position in the tile's active-chunk list); this
        // group owns the stages whose global index is congruent to `group` modulo G.
        const uint32_t magic = 0xFFFFFFFFu / (uint32_t)c_in + 1u;   // e / c_in == umulhi(e, magic) for e * c_in < 2^32
        int it_tile = 0, it_pos = group, cur_tile = -1, cur_cnt = 0, rows_left = 0;
        bool it_done = false;
        int inflight = 0;
        // Position the iterator on this group's next stage.  Every tile's index buffer is acquired and released exactly
        // once per thread.  A warp that still has gathers in flight must never BLOCK on a later tile's list: the MMA
        // issuer may be waiting for exactly those stages before it can release the buffer the list needs (groups that
        // own no stage in two consecutive sparse tiles would deadlock) — so with may_block == false an unpublished
        // list makes this return false and the caller drains a stage first.
        auto locate = [&](bool may_block) -> bool {
            while (true) {
                if (cur_tile != it_tile) {
                    uint64_t* bar = &list_full[it_tile & 1];
                    const uint32_t par = (uint32_t)(it_tile >> 1) & 1u;
                    if (!may_block && !__any_sync(0xffffffffu, mbar_test(bar, par))) return false;
                    mbar_wait(bar, par);
                    const int tile = s_tile[it_tile & 1];
                    if (tile < 0) { it_done = true; return false; }   // the scheduler's end marker
                    cur_tile = it_tile;
                    cur_cnt = s_cnt[it_tile & 1];
                    rows_left = n - (tile * TC_BM + quarter * 32);
                }
                if (it_pos < cur_cnt) return true;
                it_pos -= cur_cnt;
                ++it_tile;
                mbar_arrive(&nbr_empty[cur_tile & 1]);      // done with this tile's index block
                cur_tile = -1;
            }
        };
        auto issue = [&](int slot) {
            const int buf = it_tile & 1;
            const uint32_t chunk = lds_u16(list_s32 + 2u * (uint32_t)(buf * T + it_pos));
            // this lane's 4-float piece covers flattened elements e .. e+3 -> offset i_k = e / c_in, channel e % c_in
            // (c_in % 4 == 0, so a piece never straddles two offsets)
            // fp32 rows: piece q = floats e .. e+3.  Split rows (c_in % 32 == 0): the chunk is one 128-byte block
            // [32 hi | 32 lo] of its offset's row and piece q is its q-th 16 bytes.
            const uint32_t e = SIN ? chunk * TC_KC : chunk * TC_KC + (uint32_t)q * 4u;
            const int i_k = (int)__umulhi(e, magic);
            const uint32_t colbytes = SIN ? (e - (uint32_t)i_k * (uint32_t)c_in) * 4u + (uint32_t)q * 16u
                                          : (e - (uint32_t)i_k * (uint32_t)c_in) * 4u;
            const bool col_ok = i_k < K;            // beyond the end of the reduction axis: zero fill
            const uint32_t nb = nbr_s32 + 4u * (uint32_t)(buf * TC_BM * K + (quarter * 32 + sub) * K + (col_ok ? i_k : 0));
            const uint32_t dbase = abuf + (uint32_t)slot * 4096u;
            if (!(diag & 1)) {                      // (timing diagnostics: bit 0 drops the gather)
                // all eight index loads first: volatile asm keeps program order, and one LDS latency in front of every
                // copy was the largest stall of this warp's dependent chain (round-1 source-level profile)
                int src[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) src[g] = lds_i32(nb + 4u * (uint32_t)(g * strideK4));
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const bool ok = col_ok && (g * 4 + sub) < rows_left && src[g] >= 0;
                    const uint32_t off = ok ? (uint32_t)src[g] * rowbytes + colbytes : 0u;
                    const uint32_t dst = dbase + (uint32_t)(g * 512) + ((g & 1) ? dst_odd : dst_even);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(fbase + off), "r"(ok ? 16u : 0u) : "memory");
                }
            }
            cp_async_commit();
            ++inflight;
            it_pos += G;
        };
        const uint32_t rd_base = abuf + (uint32_t)lane * 128u;
        const uint32_t x7 = (uint32_t)(lane & 7);
        int s = group % STAGES, ph = 0, slot = 0, islot = 0;   // consume-side stage slot / phase, staging slots
        bool located = false;
        const bool tr_me = tr && warp == 0 && lane == 0;
        bool tr_first_issue = true, tr_first_arrive = true;
        long long tr_wait_empty = 0, tr_p_issue = 0, tr_p_cpwait = 0, tr_p_store = 0, tr_tp = 0;
        while (true) {
            if (tr_me) tr_tp = clock64();
            // issue ahead: up to TC_DEPTH gathers in flight, never blocking while some are
            while (!it_done && inflight < TC_DEPTH) {
                if (!located) located = locate(inflight == 0);
                if (!located) break;
                if (tr_me && tr_first_issue) { tr[4] = (unsigned long long)clock64(); tr_first_issue = false; }
                issue(islot);
                if (++islot == TC_DEPTH) islot = 0;
                located = false;
            }
            if (inflight == 0) break;              // iterator exhausted and everything drained
            if (tr_me) { const long long t = clock64(); tr_p_issue += t - tr_tp; tr_tp = t; }
            --inflight;                            // = committed groups allowed to stay pending
            if (inflight == 0) cp_async_wait<0>();
            else if (inflight == 1) cp_async_wait<1>();
            else if (inflight == 2) cp_async_wait<2>();
            else cp_async_wait<3>();
            __syncwarp();
            float4 v[8];
            if (!(diag & 2)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t a = rd_base + (uint32_t)slot * 4096u + (((uint32_t)i ^ x7) << 4);
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(a));
            }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();                          // everyone has read the slot before it is refilled
            long long tw0 = 0;
            if (tr_me) { tw0 = clock64(); tr_p_cpwait += tw0 - tr_tp; }
            mbar_wait_a(empty0 + 8u * (uint32_t)s, (uint32_t)(ph ^ 1));
            if (tr_me) tr_wait_empty += clock64() - tw0;
            tc_fence_after();
            // hi = low 13 mantissa bits cleared (exact tf32), lo = exact fp32 remainder; written in 16-column halves
            // to keep the live register set small (the 16-warp variant runs the producers at 96 registers)
            if (SIN) {
                // split rows: pieces 0..3 are the packed bf16 hi operand (32 values = 16 columns), 4..7 the lo operand
                if (!(diag & 2)) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t w[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            w[4 * i + 0] = __float_as_uint(v[4 * h + i].x);
                            w[4 * i + 1] = __float_as_uint(v[4 * h + i].y);
                            w[4 * i + 2] = __float_as_uint(v[4 * h + i].z);
                            w[4 * i + 3] = __float_as_uint(v[4 * h + i].w);
                        }
                        tmem_st16(lane_base + A_COL0 + (uint32_t)s * A_STRIDE + 16u * (uint32_t)h, w);
                    }
                }
            } else
            if (!(diag & 2))                        // (bit 1 drops the smem read-back, the hi/lo split and the TMEM stores)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[4 * i + 0] = __float_as_uint(v[4 * h + i].x) & 0xFFFFE000u;
                    w[4 * i + 1] = __float_as_uint(v[4 * h + i].y) & 0xFFFFE000u;
                    w[4 * i + 2] = __float_as_uint(v[4 * h + i].z) & 0xFFFFE000u;
                    w[4 * i + 3] = __float_as_uint(v[4 * h + i].w) & 0xFFFFE000u;
                }
                tmem_st16(lane_base + A_COL0 + (uint32_t)(s * 64 + 16 * h), w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w[4 * i + 0] = __float_as_uint(v[4 * h + i].x - __uint_as_float(w[4 * i + 0]));
                    w[4 * i + 1] = __float_as_uint(v[4 * h + i].y - __uint_as_float(w[4 * i + 1]));
                    w[4 * i + 2] = __float_as_uint(v[4 * h + i].z - __uint_as_float(w[4 * i + 2]));
                    w[4 * i + 3] = __float_as_uint(v[4 * h + i].w - __uint_as_float(w[4 * i + 3]));
                }
                tmem_st16(lane_base + A_COL0 + (uint32_t)(s * 64 + 32 + 16 * h), w);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_a(full0 + 8u * (uint32_t)s);
            if (tr_me && tr_first_arrive) { tr[5] = (unsigned long long)clock64(); tr_first_arrive = false; }
            if (tr_me) tr_p_store += clock64() - tw0;
            // advance the consume-side counters by two global stages
            s += G; if (s >= STAGES) { s -= STAGES; ph ^= 1; }
            if (++slot == TC_DEPTH) slot = 0;
        }
        if (cur_tile >= 0) mbar_arrive(&nbr_empty[cur_tile & 1]);   // (not reached: the iterator releases as it leaves)
        if (tr_me) {
            tr[14] = (unsigned long long)tr_wait_empty;
            tr[22] = (unsigned long long)tr_p_issue;      // locate (incl. blocking on chunk lists) + gather issue
            tr[23] = (unsigned long long)tr_p_cpwait;     // cp.async wait + shared read-back
            tr[24] = (unsigned long long)tr_p_store;      // wait for the free A slot + TMEM store + arrive
        }
    } else if (warp == TC_MMA_WARP) {
        // ================= MMA issuer =================
        // The whole warp runs the loop in warp-uniform control flow and one ELECTed lane issues: a divergent
        // `if (lane == 0)` makes the compiler wrap every UTCHMMA in a serialisation loop (~100 cycles per MMA
        // instead of the 32-cycle M*N/256 floor measured with tools/mma_rate.cu).
        // The issue loop is a dependent chain on one thread: stage slot / phase are running counters, barrier addresses
        // and the descriptor words are hoisted, and a stage's descriptors differ from the base only by an add on the
        // 14-bit start-address field (no carry: shared addresses < 256 KB).
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint64_t desc0 = SIN ? make_desc_sw64(smem_u32(stages)) : make_desc_sw128(smem_u32(stages));
        const uint64_t desc_hi64 = desc0 & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = (uint32_t)desc0;
        uint32_t s = 0, ph = 0;   // ring slot / phase (A operand in TMEM and weight tile in smem share both)
        long long tr_wait_full = 0, tr_wait_list = 0, tr_wait_acc = 0;
        unsigned long long tr_stages = 0;
        bool tr_first = true;
        for (int tl = 0;; ++tl) {
            const int buf = tl & 1;
            long long tw0 = 0;
            if (tr) tw0 = clock64();
            mbar_wait(&list_full[buf], (tl >> 1) & 1);
            if (tr) { const long long t = clock64(); tr_wait_list += t - tw0; tw0 = t; }
            const int tile = s_tile[buf];
            const int cnt = s_cnt[buf];
            __syncwarp();
            mbar_wait(&tmem_empty[buf], ((tl >> 1) & 1) ^ 1);    // epilogue drained this accumulator (and read its tile id)
            if (tr) tr_wait_acc += clock64() - tw0;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * ACC_COLS;
            if (elect_one()) {
                s_epi_tile[buf] = tile;
                mbar_arrive(&tmem_full[buf]);           // release: the tile id is visible once the phase completes
                if (tile < 0) mbar_arrive(&tmem_full[buf]);          // end marker: no MMAs, complete the phase now
                else mbar_arrive(&nbr_empty[buf]);
            }
            __syncwarp();
            if (tile < 0) break;
            // one stage = 12 tf32 (6 bf16) MMAs + the commit that frees its A and weight slots
            auto issue_stage = [&](uint32_t sa, bool first) {
                const uint32_t a_hi = tmem_base + A_COL0 + sa * A_STRIDE;
                const uint32_t a_lo = a_hi + (SIN ? 16u : 32u);
                const uint32_t dl = desc_lo0 + sa * (uint32_t)(STAGE_BYTES >> 4);
                constexpr int KSTEPS = SIN ? TC_KC / 16 : TC_KC / 8;   // UMMA_K = 16 bf16 / 8 tf32 = 8 TMEM columns, 32 B of B
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                    const uint64_t db_hi = desc_hi64 | (uint64_t)(dl + 2u * (uint32_t)kk);
                    const uint32_t acc = (first && kk == 0) ? 0u : 1u;
                    const uint64_t db_lo = desc_hi64 | (uint64_t)(dl + (uint32_t)(B_BYTES >> 4) + 2u * (uint32_t)kk);
                    if (SIN) {
                        umma_f16_ts(d_tmem, a_lo + kk * 8, db_hi, IDESC, acc);    // small terms first
                        if (!(diag & 4)) {
                            umma_f16_ts(d_tmem, a_hi + kk * 8, db_lo, IDESC, 1u);
                            umma_f16_ts(d_tmem, a_hi + kk * 8, db_hi, IDESC, 1u);
                        }
                    } else {
                        umma_tf32_ts(d_tmem, a_lo + kk * 8, db_hi, IDESC, acc);   // small terms first
                        if (!(diag & 4)) {            // (bit 2: one MMA per k-step instead of three)
                            umma_tf32_ts(d_tmem, a_hi + kk * 8, db_lo, IDESC, 1u);
                            umma_tf32_ts(d_tmem, a_hi + kk * 8, db_hi, IDESC, 1u);
                        }
                    }
                }
                umma_commit_a(empty0 + 8u * sa);   // frees the A and weight stages once the MMAs retire
            };
            // Two stages per trip where the list allows (diag bit 4 forces one): the wait -> fence -> elect -> issue ->
            // reconverge sequence has a fixed latency that a 12-MMA stage does not cover.
            int j = 0;
            if (!(diag & 16))
            for (; j + 1 < cnt; j += 2) {
                uint32_t s1 = s + 1, ph1 = ph;
                if (s1 == (uint32_t)STAGES) { s1 = 0; ph1 ^= 1u; }
                long long tw1 = 0;
                if (tr) tw1 = clock64();
                mbar_wait_a(full0 + 8u * s, ph);          // A operand in TMEM and weight tile in smem (one barrier)
                mbar_wait_a(full0 + 8u * s1, ph1);
                if (tr) {
                    const long long t = clock64();
                    tr_wait_full += t - tw1;
                    tr_stages += 2;
                    if (tr_first && lane == 0) { tr[6] = (unsigned long long)t; }
                    tr_first = false;
                }
                tc_fence_after();
                if (elect_one()) {
                    issue_stage(s, j == 0);
                    issue_stage(s1, false);
                }
                __syncwarp();
                s = s1 + 1; ph = ph1;
                if (s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
            }
            for (; j < cnt; ++j) {
                long long tw1 = 0;
                if (tr) tw1 = clock64();
                mbar_wait_a(full0 + 8u * s, ph);         // A operand in TMEM + weight tile landed
                if (tr) {
                    const long long t = clock64();
                    tr_wait_full += t - tw1;
                    tr_stages += 1;
                    if (tr_first && lane == 0) { tr[6] = (unsigned long long)t; }
                    tr_first = false;
                }
                tc_fence_after();
                if (elect_one()) issue_stage(s, j == 0);
                __syncwarp();
                if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(&tmem_full[buf]);   // accumulator complete -> epilogue
            __syncwarp();
            if (tr && lane == 0) tr[7] = (unsigned long long)clock64();
        }
        if (tr && lane == 0) {
            tr[12] = (unsigned long long)tr_wait_full;
            tr[13] = tr_stages;
            tr[3] = (unsigned long long)tr_wait_list;     // waiting for the index loader's chunk lists
            tr[15] = (unsigned long long)tr_wait_acc;      // waiting for the epilogue to drain an accumulator
        }
    } else if (warp == TC_BLD_WARP) {
        // ================= weight loader: one cp.async.bulk per stage, NB stages ahead of the MMAs =================
        if (lane == 0) {
            const uint32_t list_s32 = smem_u32(s_list);
            uint32_t sb = 0, pb = 0;
            for (int tl = 0;; ++tl) {
                const int buf = tl & 1;
                mbar_wait(&list_full[buf], (tl >> 1) & 1);
                if (s_tile[buf] < 0) break;
                const int cnt = s_cnt[buf];
                for (int j = 0; j < cnt; ++j) {
                    const uint32_t chunk = lds_u16(list_s32 + 2u * (uint32_t)(buf * T + j));
                    mbar_wait(&empty_bar[sb], pb ^ 1u);   // the MMAs that read this slot have retired
                    tc_fence_after();
                    if (!(diag & 8)) {                         // (timing diagnostics: bit 3 drops the weight-tile copy)
                        mbar_expect_tx(&full_bar[sb], 2 * B_BYTES);
                        bulk_copy_g2s(stages + sb * STAGE_BYTES, (const char*)packed_w + (int64_t)chunk * (2 * B_BYTES),
                                      2 * B_BYTES, &full_bar[sb]);
                    }
                    mbar_arrive(&full_bar[sb]);
                    if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1u; }
                }
                mbar_arrive(&nbr_empty[buf]);                  // done with this tile's chunk list
            }
        }
    } else if (warp == TC_IDX_WARP) {
        // ================= index loader: TMA-stage each tile's neighbour block, publish its active chunks ==========
        // Tile scheduler.  Static: tile = blockIdx.x + tl * gridDim.x.  Dynamic (tile_ctr != null): the first tile is
        // blockIdx.x, the following ones come from a global counter — tiles cost between a few and all T chunks, so a
        // fixed round-robin leaves SMs idle ~20 % of a layer.  tile_ctr[0] = positions handed out beyond
        // the first P = min(grid, tiles), tile_ctr[1] = CTAs that are done claiming; the last CTA to report zeroes both.
        const int P = num_tiles < (int)gridDim.x ? num_tiles : (int)gridDim.x;
        const uint32_t nbr_s32 = smem_u32(nbr_s);
        // Dynamic hand-out in guided batches (lane 0).  One global counter served a tile per atomicAdd: on the thin layers
        // (~1700 tiles in ~40 us) 148 CTAs on one address run into the L2 atomic unit's per-address rate (B300_MICROARCH
        // "L2-atom multi-CTA": ~27 cycles per op, microseconds of latency under contention), so a claim takes
        // clamp(remaining / (4 P), 1, 4) consecutive positions of the heaviest-first sequence — the batches shrink to one
        // tile towards the end, where balance matters.  While plenty of tiles remain (> 4 per CTA) the next claim is issued
        // as soon as the local queue runs dry (two tiles ahead of the one being staged); in the last rounds only when the
        // next tile is actually needed, so that no CTA sits on a claimed tile while others idle.  A reply is only looked
        // at after the current tile's list has been published.
        bool exhausted = false;       // the counter ran past the last position: no more claims
        bool claim_inflight = false;
        int claim_raw = 0, claim_n = 0, q_pos = 0, q_end = 0, rem_est = num_tiles - P;   // local queue [q_pos, q_end)
        auto claim = [&]() {
            if (!tile_ctr || exhausted || claim_inflight) return;
            int b = rem_est / (4 * P);
            b = b < 1 ? 1 : (b > 4 ? 4 : b);
            claim_n = b;
            claim_raw = atomicAdd(tile_ctr, b);
            claim_inflight = true;
        };
        auto settle = [&]() {          // look at the reply of the claim in flight (stalls until it is back)
            if (!claim_inflight) return;
            claim_inflight = false;
            const int first = P + claim_raw;
            rem_est = num_tiles - first - claim_n;
            if (first >= num_tiles) { exhausted = true; return; }
            q_pos = first;
            q_end = first + claim_n < num_tiles ? first + claim_n : num_tiles;
            if (q_end == num_tiles) exhausted = true;
        };
        if (lane == 0) claim();                          // lands while the prefix below is built
        if (tile_order) {
            // s_cls_start[r] = number of tiles in classes heavier than class (64 - r); s_cls_start[65] = all tiles
            int c0 = __ldg(tile_order + (64 - lane)), c1 = __ldg(tile_order + (32 - lane >= 0 ? 32 - lane : 0));
            if (lane > 32) c1 = 0;
            const int i0 = warp_inclusive_scan(c0);
            const int t0 = __shfl_sync(0xffffffffu, i0, 31);
            const int i1 = warp_inclusive_scan(c1) + t0;
            s_cls_start[lane] = i0 - c0;                       // r = lane        (classes 64 .. 33)
            s_cls_start[32 + lane] = i1 - c1;                  // r = 32 + lane   (classes 32 .. 1)
            if (lane == 0) s_cls_start[64] = 0;                // fixed below
            __syncwarp();
            if (lane == 0) {
                const int c_last = __ldg(tile_order + 0);      // class 0 (cannot hold live tiles of a well-formed table)
                s_cls_start[64] = s_cls_start[63] + __ldg(tile_order + 1);
                s_cls_start[65] = s_cls_start[64] + c_last;
            }
            __syncwarp();
        }
        long long tr_idx_wait = 0, tr_idx_fetch = 0, tr_idx_copy = 0, tr_idx_list = 0, tr_t = 0;
        auto static_pos = [&](int j) -> int { return j < my_tiles ? (int)blockIdx.x + j * (int)gridDim.x : -1; };
        // heaviest-first hand-out (btc_rulebook_tile_meta): position in the sequence -> (cost class, slot) -> tile id, so
        // that the last tiles of a launch are its cheapest and the CTAs finish together
        auto order_lookup = [&](int pos) -> int {
            if (pos < 0 || !tile_order) return pos;
            int lo_r = 0, hi_r = 65;                    // largest r with s_cls_start[r] <= pos
            while (hi_r - lo_r > 1) {
                const int mid = (lo_r + hi_r) >> 1;
                if (s_cls_start[mid] <= pos) lo_r = mid; else hi_r = mid;
            }
            return __ldg(tile_order + 65 + (int64_t)(64 - lo_r) * tiles_cap + (pos - s_cls_start[lo_r]));
        };
        int tile_cur = -1;
        if (lane == 0) tile_cur = order_lookup((int)blockIdx.x);    // iteration 0 (static and dynamic alike)
        for (int tl = 0;; ++tl) {
            const int buf = tl & 1;
            if (tr) tr_t = clock64();
            if (lane == 0) mbar_wait(&nbr_empty[buf], ((tl >> 1) & 1) ^ 1);
            if (tr && lane == 0) { const long long t = clock64(); tr_idx_wait += t - tr_t; tr_t = t; }
            const int tile = __shfl_sync(0xffffffffu, tile_cur, 0);
            if (tile < 0) {                            // publish the end marker and leave
                if (lane == 0) {
                    s_tile[buf] = -1;
                    s_cnt[buf] = 0;
                    mbar_arrive(&list_full[buf]);
                    if (tr) {
                        tr[11] = (unsigned long long)tl;
                        tr[16] = (unsigned long long)tr_idx_wait;
                        tr[17] = (unsigned long long)tr_idx_fetch;
                        tr[18] = (unsigned long long)tr_idx_copy;
                        tr[19] = (unsigned long long)tr_idx_list;
                    }
                }
                break;
            }
            const int row0 = tile * TC_BM;
            int* dst = nbr_s + buf * TC_BM * K;
            unsigned long long m = 0;
            int tile_nxt = -1;
            bool have_nxt = false;
            if (lane == 0) {
                const int rows = n_cap - row0 < TC_BM ? n_cap - row0 : TC_BM;
                const uint32_t bytes = (uint32_t)rows * K * 4, bulk = bytes & ~15u;
                const int* src = table + (int64_t)row0 * K;
                if (bulk) {
                    mbar_expect_tx(&nbr_full[buf], bulk);
                    bulk_copy_g2s(dst, src, bulk, &nbr_full[buf]);
                }
                if (tile_mask) m = __ldg(tile_mask + tile);
                // next iteration's tile id: from the local queue when it holds one (its order entry then has this whole
                // iteration to arrive), else from the claim in flight — looked at below, after this tile's list is out
                if (!tile_ctr) {
                    tile_nxt = order_lookup(static_pos(tl + 1));
                    have_nxt = true;
                } else if (q_pos < q_end) {
                    tile_nxt = order_lookup(q_pos++);
                    have_nxt = true;
                    if (q_pos == q_end && rem_est > 4 * P) claim();   // the queue just ran dry: claim the next batch now
                } else {
                    claim();
                }
                for (uint32_t e = bulk / 4; e < bytes / 4; ++e) dst[e] = __ldg(src + e);   // < 16-byte tail
                mbar_arrive(&nbr_full[buf]);
            }
            __syncwarp();
            if (tr && lane == 0) { const long long t = clock64(); tr_idx_fetch += t - tr_t; tr_t = t; }
            // offsets with at least one valid neighbour among the tile's live rows: precomputed with the rulebook
            // (btc_rulebook_tile_meta; the chunk list is then built while the index block is still in flight), else scanned
            // here from the staged index tile (~2.5 us per tile on one warp)
            // Thin layers (c_in <= 16: a 32-element chunk spans two or more offsets, and a 128-row tile of a LiDAR scene
            // touches practically all of them) take every chunk as active when no mask is given: skipping is only an
            // optimisation, and neither the mask launch (22 us on the serial front of the step) nor the scan pays there.
            const bool dense = !tile_mask && c_in <= 16;
            if (tile_mask) {
                m = __shfl_sync(0xffffffffu, m, 0);
            } else if (dense) {
                m = ~0ull;
            } else {
                mbar_wait(&nbr_full[buf], (tl >> 1) & 1);
                const int rows_live = n - row0 < TC_BM ? n - row0 : TC_BM;
                for (int r = lane; r < rows_live; r += 32) {
                    const uint32_t rp = nbr_s32 + 4u * (uint32_t)(buf * TC_BM * K + r * K);
                    for (int k = 0; k < K; ++k) m |= (unsigned long long)(lds_i32(rp + 4u * (uint32_t)k) >= 0) << k;
                }
                const uint32_t m_lo = __reduce_or_sync(0xffffffffu, (uint32_t)m);
                const uint32_t m_hi = __reduce_or_sync(0xffffffffu, (uint32_t)(m >> 32));
                m = (unsigned long long)m_lo | ((unsigned long long)m_hi << 32);
            }
            int cnt = 0;
            for (int c0 = 0; c0 < T; c0 += 32) {
                const int c = c0 + lane;
                bool act = false;
                if (c < T) {
                    const int k_lo = (c * TC_KC) / c_in;
                    int k_hi = (c * TC_KC + TC_KC - 1) / c_in;
                    if (k_hi > K - 1) k_hi = K - 1;
                    const int span = k_hi - k_lo + 1;
                    const unsigned long long bits = (span >= 64 ? ~0ull : ((1ull << span) - 1ull)) << k_lo;
                    act = (m & bits) != 0ull;
                }
                const unsigned b = __ballot_sync(0xffffffffu, act);
                if (act) s_list[buf * T + cnt + __popc(b & ((1u << lane) - 1u))] = (unsigned short)c;
                cnt += __popc(b);
            }
            if (cnt == 0) {                          // cannot happen for a well-formed rulebook; keep the pipeline alive
                if (lane == 0) s_list[buf * T] = 0;
                cnt = 1;
            }
            if (tr && lane == 0) { const long long t = clock64(); tr_idx_list += t - tr_t; tr_t = t; }
            if (tile_mask || dense) mbar_wait(&nbr_full[buf], (tl >> 1) & 1);     // the index block has landed
            if (lane == 0) { s_cnt[buf] = cnt; s_tile[buf] = tile; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&list_full[buf]);
            if (tr && lane == 0) { const long long t = clock64(); tr_idx_copy += t - tr_t; tr_t = t; }
            if (lane == 0 && !have_nxt) {              // the queue was dry: take the first tile of the claim in flight
                settle();
                if (q_pos < q_end) {
                    tile_nxt = order_lookup(q_pos++);
                    if (q_pos == q_end && rem_est > 4 * P) claim();
                }
            }
            tile_cur = tile_nxt;
        }
        // every CTA of the launch reports here after its last claim; the last one re-arms the counter pair for the next
        // launch that is handed this slot
        if (lane == 0 && tile_ctr) {
            settle();
            __threadfence();
            if (atomicAdd(tile_ctr + 1, 1) == (int)gridDim.x - 1) {
                atomicExch(tile_ctr, 0);
                atomicExch(tile_ctr + 1, 0);
            }
        }
    } else if (warp >= Roles::kEpi0 && warp < Roles::kEpi0 + 4) {
        // ================= epilogue =================
        const int quarter = warp & 3;                // TMEM lanes [32*quarter, 32*quarter+32)
        const bool tr_epi = tr && warp == Roles::kEpi0 && lane == 0;
        long long tr_epi_wait = 0, tr_epi_busy = 0, tr_te = 0;
        for (int tl = 0;; ++tl) {
            const int buf = tl & 1;
            if (tr_epi) tr_te = clock64();
            mbar_wait(&tmem_full[buf], (tl >> 1) & 1);
            if (tr_epi) { const long long t = clock64(); tr_epi_wait += t - tr_te; tr_te = t; }
            tc_fence_after();
            const int tile = s_epi_tile[buf];
            if (tile < 0) break;
            const int slot_row = tile * TC_BM + quarter * 32 + lane;
            // sorted rulebooks (btc_rulebook_sort_rows) process rows in mask order and scatter to the original rows
            const int row = slot_row < n ? (out_rows ? __ldg(out_rows + slot_row) : slot_row) : n;
            float* dst = feat_out + (int64_t)row * c_out;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                if (c0 >= c_out) break;                       // padded accumulator columns (c_out < N): nothing to write
                uint32_t acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * ACC_COLS + (uint32_t)c0, acc);
                if (row < n) {
                float x[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {              // broadcast vector loads of the column constants
                    const float4 b4 = *reinterpret_cast<const float4*>(s_ep_bias + c0 + 4 * j4);
                    const float4 s4 = *reinterpret_cast<const float4*>(s_ep_scale + c0 + 4 * j4);
                    const float4 h4 = *reinterpret_cast<const float4*>(s_ep_shift + c0 + 4 * j4);
                    x[4 * j4 + 0] = (__uint_as_float(acc[4 * j4 + 0]) + b4.x) * s4.x + h4.x;
                    x[4 * j4 + 1] = (__uint_as_float(acc[4 * j4 + 1]) + b4.y) * s4.y + h4.y;
                    x[4 * j4 + 2] = (__uint_as_float(acc[4 * j4 + 2]) + b4.z) * s4.z + h4.z;
                    x[4 * j4 + 3] = (__uint_as_float(acc[4 * j4 + 3]) + b4.w) * s4.w + h4.w;
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (SOUT) {
                    // split format out: 16 channels -> 8 packed bf16 hi words + 8 lo words of the row's 128-byte block
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) {
                        const float x0 = x[2 * pp], x1 = x[2 * pp + 1];
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                        const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                        const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                        hi[pp] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[pp] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    char* blk = reinterpret_cast<char*>(dst) + (c0 >> 5) * 128 + ((c0 >> 4) & 1) * 32;
                    *reinterpret_cast<uint4*>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(blk + 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4*>(blk + 64) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(blk + 80) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const int col = c0 + j4 * 4;
                        if (col + 3 < c_out) {
                            *reinterpret_cast<float4*>(dst + col) = make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (col + jj < c_out) dst[col + jj] = x[4 * j4 + jj];
                        }
                    }
                }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);           // this accumulator buffer may be overwritten
            if (tr_epi) {
                const long long t = clock64();
                tr_epi_busy += t - tr_te;
                tr[8] = (unsigned long long)t;
                tr[20] = (unsigned long long)tr_epi_wait;
                tr[21] = (unsigned long long)tr_epi_busy;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
    if (tr && tid == 0) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        tr[9] = (unsigned long long)clock64();
        tr[10] = g;
    }
}

// ---- per-tile metadata of a neighbour table (btc_rulebook_tile_meta) ------------------------------------------------
// Masks and cost classes in ONE launch, no CTA waiting for another: every live CTA computes the mask of its tile, takes a
// slot in the bucket of its cost class (number of active offsets, 0..64) with one atomic and stores its tile id there.
// `tile_order` = [0, 65): class counts (zeroed by the caller), [65, 65 + 65 * tiles_cap): the buckets.  The conv kernel's
// scheduler turns a position of the heaviest-first sequence into a tile id with a 65-entry prefix (tc_order_lookup).
__global__ void __launch_bounds__(128) tile_meta_kernel(const int* __restrict__ table, int n_cap, const int* __restrict__ n_dev,
                                                        int K, unsigned long long* __restrict__ tile_mask,
                                                        int* __restrict__ tile_order, int tiles_cap) {
    __shared__ unsigned long long s_m[4];
    const int n = live_count(n_cap, n_dev);
    for (int tile = blockIdx.x; tile * TC_BM < n; tile += gridDim.x) {     // grid-stride over the LIVE tiles
        const int row0 = tile * TC_BM;
        unsigned long long m = 0;
        const int rows = n - row0 < TC_BM ? n - row0 : TC_BM;
        const int total = rows * K;
        const int* src = table + (int64_t)row0 * K;
        // coalesced sweep of the tile's [rows x K] block in batches of 16 independent loads per thread (a load whose
        // value is tested right away costs a full L2 round trip each)
        for (int i0 = threadIdx.x; i0 < total; i0 += 128 * 16) {
            int v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = i0 + 128 * j < total ? __ldg(src + i0 + 128 * j) : -1;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (v[j] >= 0) m |= 1ull << ((i0 + 128 * j) % K);
        }
        const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)m), hi = __reduce_or_sync(0xffffffffu, (uint32_t)(m >> 32));
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = (unsigned long long)lo | ((unsigned long long)hi << 32);
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long mm = s_m[0] | s_m[1] | s_m[2] | s_m[3];
            tile_mask[tile] = mm;
            if (tile_order) {
                const int c = __popcll(mm);
                const int pos = atomicAdd(tile_order + c, 1);
                tile_order[65 + (int64_t)c * tiles_cap + pos] = tile;
            }
        }
        __syncthreads();
    }
}

// Tile counters of the dynamic scheduler: zero at module load, every launch leaves its counter pair at zero again (see
// the scheduler warp).  Launches rotate through the slots, so kernels that overlap on different streams (or a captured
// graph and eager launches) do not share one unless kTcCtrSlots launches are in flight at once.
constexpr int kTcCtrSlots = 1024;
__device__ int g_tc_tile_ctr[2 * kTcCtrSlots];   // per slot: {next position, CTAs finished}
static int g_tc_npw = 0, g_tc_dyn = -1, g_tc_diag = 0, g_tc_grid = kNumSM;
static unsigned long long* g_tc_trace = nullptr;   // timeline diagnostics buffer (device, 32 u64 per CTA) or null

static int* next_tile_counter() {
    static int* base[64] = {nullptr};            // (benign race: every thread computes the same address)
    static std::atomic<unsigned> next{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!base[dev]) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_tc_tile_ctr) != cudaSuccess) return nullptr;
        base[dev] = (int*)p;
    }
    return base[dev] + 2 * (next.fetch_add(1, std::memory_order_relaxed) % kTcCtrSlots);
}

// dynamic shared memory of one CTA: weight ring + cp.async staging + two index tiles + two chunk lists + alignment slack
static size_t tc_smem_bytes(int N, int npw, int K, int c_in) {
    const int stage_bytes = 2 * N * 128;
    const int nb = N <= 32 ? 6 : 4, depth = (N > 64 || npw > 8) ? 2 : 4;
    const int T = (K * c_in + TC_KC - 1) / TC_KC;
    return (size_t)nb * stage_bytes + (size_t)npw * depth * 4096 + (size_t)2 * TC_BM * K * sizeof(int) +
           (size_t)((4 * T + 15) & ~15) + 1024 + 16;
}
constexpr size_t kTcMaxSmem = 227 * 1024;

template <int N, int NPW, bool SIN, bool SOUT>
static int launch_tc_npw(const float* feat_in, const int* table, const float* packed_w, const float* bias,
                         const float* scale, const float* shift, int relu, float* feat_out, const int* out_rows, int n_cap,
                         const int* n_dev, int K, int c_in, int c_out, cudaStream_t st,
                         const unsigned long long* tile_mask, const int* tile_order) {
    static_assert(TcBStages<N, false>::value == (N <= 32 ? 6 : 4) && TcBStages<N, true>::value == 2 * TcBStages<N, false>::value &&
                      TcDepth<N, NPW>::value == ((N > 64 || NPW > 8) ? 2 : 4),
                  "tc_smem_bytes mirrors these (split format: twice the stages of half the bytes)");
    const size_t smem = tc_smem_bytes(N, NPW, K, c_in);
    auto kern = conv_fwd_tc_kernel<N, NPW, SIN, SOUT>;
    // opt in to > 48 KB dynamic smem: the attribute is per device / context, so it is tracked per device (not a stream op)
    static size_t attr_set[64] = {0};
    int dev = 0;
    BTC_CUDA(cudaGetDevice(&dev), "tc device");
    if (dev < 0 || dev >= 64 || attr_set[dev] < smem) {
        BTC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tc smem attr");
        if (dev >= 0 && dev < 64) attr_set[dev] = smem;
    }
    int tiles = (n_cap + TC_BM - 1) / TC_BM;
    dim3 grid(tiles < g_tc_grid ? tiles : g_tc_grid);    // persistent: one CTA per SM (or fewer: btc_sparse_conv_tc_grid)
    int* ctr = nullptr;
    if (g_tc_dyn) {
        ctr = next_tile_counter();
        if (!ctr) return set_error(BTC_E_CUDA, "btc_sparse_conv_fwd_tc: tile counter symbol not available", cudaGetLastError());
    }
    kern<<<grid, TcRoles<NPW>::kThreads, smem, st>>>(feat_in, table, packed_w, bias, scale, shift, relu, feat_out, out_rows,
                                                     n_cap, n_dev, K, c_in, c_out, ctr, g_tc_diag, g_tc_trace, tile_mask,
                                                     g_tc_dyn ? tile_order : nullptr, tiles);
    BTC_CHECK_LAUNCH("conv_fwd_tc");
    return BTC_OK;
}

// Tile variant knobs (A/B measurements and tests; defaults are the measured-best ones).  16 producer warps (four
// groups) when shared memory allows (N <= 64), else 8.  Dynamic tile scheduling (global counter) instead of a fixed
// round-robin.  Environment overrides at first use: BTC_TC_NPW=8, BTC_TC_DYN=0; btc_sparse_conv_tc_config() changes them
// at run time.
static void tc_config_init() {
    if (g_tc_npw == 0) {
        const char* e = getenv("BTC_TC_NPW");
        g_tc_npw = (e && atoi(e) == 8) ? 8 : 16;
    }
    if (g_tc_dyn < 0) {
        const char* e = getenv("BTC_TC_DYN");
        g_tc_dyn = (e && atoi(e) == 0) ? 0 : 1;
    }
}

template <int N, bool SIN, bool SOUT>
static int launch_tc(const float* feat_in, const int* table, const float* packed_w, const float* bias,
                     const float* scale, const float* shift, int relu, float* feat_out, const int* out_rows, int n_cap,
                     const int* n_dev, int K, int c_in, int c_out, cudaStream_t st,
                     const unsigned long long* tile_mask, const int* tile_order) {
    tc_config_init();
    constexpr int NS = N <= 64 ? N : 64;   // instantiation guard for the N <= 64 only variant
#define BTC_TC_ARGS feat_in, table, packed_w, bias, scale, shift, relu, feat_out, out_rows, n_cap, n_dev, K, c_in, c_out, st, tile_mask, tile_order
    if (N <= 64 && g_tc_npw == 16) return launch_tc_npw<NS, 16, SIN, SOUT>(BTC_TC_ARGS);
    return launch_tc_npw<N, 8, SIN, SOUT>(BTC_TC_ARGS);
#undef BTC_TC_ARGS
}

static int tc_padded_n(int c_out) {
    if (c_out <= 32) return 32;
    if (c_out <= 64) return 64;
    if (c_out <= 128) return 128;
    return 0;
}

}  // namespace btc

using namespace btc;

extern "C" {

int btc_sparse_conv_tc_config(int producer_warps, int concat_b, int dynamic_tiles) {
    tc_config_init();
    if (concat_b > 0)
        return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_tc_config: the concatenated [B_hi|B_lo] variant was measured (no gain) and removed", cudaSuccess);
    if (producer_warps >= 0) {
        if (producer_warps != 8 && producer_warps != 16) return badarg("btc_sparse_conv_tc_config: producer_warps must be 8 or 16");
        g_tc_npw = producer_warps;
    }
    if (dynamic_tiles >= 0) g_tc_dyn = dynamic_tiles ? 1 : 0;
    return BTC_OK;
}

int btc_sparse_conv_tc_grid(int max_ctas) {
    if (max_ctas < 1 || max_ctas > kNumSM) return badarg("btc_sparse_conv_tc_grid: max_ctas must be in [1, 148]");
    g_tc_grid = max_ctas;
    return BTC_OK;
}

int btc_sparse_conv_tc_diag(int mask) {
    g_tc_diag = mask;
    return BTC_OK;
}

int btc_sparse_conv_tc_trace(void* trace_u64) {
    g_tc_trace = (unsigned long long*)trace_u64;   // device buffer of 148 * 32 u64, or null to switch the timeline off
    return BTC_OK;
}

int btc_sparse_conv_tc_supported(int K, int c_in, int c_out) {
    // c_out: a multiple of 4 (vector stores of 16-byte aligned row pieces) or below 4 (scalar stores: the occupancy head's
    // 32 -> 2 and 32 -> 3 sub-manifold convolutions, occ_head_3D.py:25-31)
    if (!(K >= 1 && K <= 64 && c_in >= 4 && c_in % 4 == 0 && c_out >= 1 && (c_out % 4 == 0 || c_out < 4) &&
          tc_padded_n(c_out) != 0 && (int64_t)K * c_in >= 32))
        return 0;
    const int N = tc_padded_n(c_out);   // the index tiles of large kernels (K > 33) do not fit next to the rings
    return tc_smem_bytes(N, N <= 64 ? 16 : 8, K, c_in) <= kTcMaxSmem && tc_smem_bytes(N, 8, K, c_in) <= kTcMaxSmem ? 1 : 0;
}

int64_t btc_sparse_conv_tc_packed_bytes(int K, int c_in, int c_out) {
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return BTC_E_UNSUPPORTED;
    int N = tc_padded_n(c_out);
    int nchunk = (K * c_in + TC_KC - 1) / TC_KC;
    return (int64_t)nchunk * 2 * N * 128;
}

int btc_sparse_conv_tc_pack(const float* weight, int K, int c_in, int c_out, void* packed, void* stream) {
    if (!weight || !packed) return badarg("btc_sparse_conv_tc_pack: null argument");
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_tc_pack: shape not supported", cudaSuccess);
    int N = tc_padded_n(c_out);
    int nchunk = (K * c_in + TC_KC - 1) / TC_KC;
    int64_t total = (int64_t)nchunk * N * TC_KC;
    tc_pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, c_in, c_out, N, (float*)packed);
    BTC_CHECK_LAUNCH("tc_pack_weight");
    return BTC_OK;
}

// ---- split (bf16 hi / lo) format ------------------------------------------------------------------------------------
int btc_sparse_conv_tc_split_supported(int K, int c_in, int c_out, int in_split, int out_split) {
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return 0;
    if (in_split && c_in % 32 != 0) return 0;
    if (out_split && c_out % 32 != 0) return 0;
    return 1;
}

int64_t btc_sparse_conv_tc_split_packed_bytes(int K, int c_in, int c_out) {
    if (!btc_sparse_conv_tc_split_supported(K, c_in, c_out, 1, 0)) return BTC_E_UNSUPPORTED;
    const int N = tc_padded_n(c_out);
    const int nchunk = (K * c_in + TC_KC - 1) / TC_KC;
    return (int64_t)nchunk * 2 * N * 64;
}

int btc_sparse_conv_tc_pack_split(const float* weight, int K, int c_in, int c_out, void* packed, void* stream) {
    if (!weight || !packed) return badarg("btc_sparse_conv_tc_pack_split: null argument");
    if (!btc_sparse_conv_tc_split_supported(K, c_in, c_out, 1, 0))
        return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_tc_pack_split: shape not supported", cudaSuccess);
    const int N = tc_padded_n(c_out);
    const int nchunk = (K * c_in + TC_KC - 1) / TC_KC;
    const int64_t total = (int64_t)nchunk * N * TC_KC;
    tc_pack_weight_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, c_in, c_out, N,
                                                                                       (unsigned short*)packed);
    BTC_CHECK_LAUNCH("tc_pack_weight_split");
    return BTC_OK;
}

int btc_features_to_split(const float* feat, int n_cap, const int* n_dev, int c, void* out, void* stream) {
    if (n_cap < 0 || c < 32 || c % 32 != 0) return badarg("btc_features_to_split: channels must be a multiple of 32");
    if (n_cap == 0) return BTC_OK;
    if (!feat || !out) return badarg("btc_features_to_split: null argument");
    features_to_split_kernel<<<grid_for((int64_t)n_cap * c / 2, 256), 256, 0, (cudaStream_t)stream>>>(feat, n_cap, n_dev, c,
                                                                                                     (uint32_t*)out);
    BTC_CHECK_LAUNCH("features_to_split");
    return BTC_OK;
}

int btc_features_from_split(const void* feat_split, int n_cap, const int* n_dev, int c, float* out, void* stream) {
    if (n_cap < 0 || c < 32 || c % 32 != 0) return badarg("btc_features_from_split: channels must be a multiple of 32");
    if (n_cap == 0) return BTC_OK;
    if (!feat_split || !out) return badarg("btc_features_from_split: null argument");
    features_from_split_kernel<<<grid_for((int64_t)n_cap * c / 2, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)feat_split, n_cap, n_dev, c, out);
    BTC_CHECK_LAUNCH("features_from_split");
    return BTC_OK;
}

static int fwd_tc(const char* who, const float* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                  const float* scale, const float* shift, int relu, float* feat_out, const int* out_rows, int n_out_cap,
                  const int* n_out_dev, int K, int c_in, int c_out, void* stream,
                  const unsigned long long* tile_mask = nullptr, const int* tile_order = nullptr, int in_split = 0,
                  int out_split = 0) {
    (void)who;
    if ((scale == nullptr) != (shift == nullptr)) return badarg("btc_sparse_conv_fwd_tc: scale/shift must come together");
    if (!btc_sparse_conv_tc_split_supported(K, c_in, c_out, in_split, out_split))
        return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_fwd_tc: shape not supported", cudaSuccess);
    if (n_out_cap <= 0) return BTC_OK;   // empty output (null data pointers of 0-row tensors are fine)
    if (!nbr_out || !packed_weight || !feat_out) return badarg("btc_sparse_conv_fwd_tc: null argument");
    if (!feat_in) return badarg("btc_sparse_conv_fwd_tc: null feat_in");
    if (((uintptr_t)feat_in & 15) || ((uintptr_t)feat_out & 15) || ((uintptr_t)packed_weight & 15))
        return badarg("btc_sparse_conv_fwd_tc: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const float* pw = (const float*)packed_weight;
#define BTC_TC_CALL(NN, SI, SO) \
    launch_tc<NN, SI, SO>(feat_in, nbr_out, pw, bias, scale, shift, relu, feat_out, out_rows, n_out_cap, n_out_dev, K, c_in, c_out, st, tile_mask, tile_order)
#define BTC_TC_FMT(NN)                                              \
    do {                                                            \
        if (in_split && out_split) return BTC_TC_CALL(NN, true, true);   \
        if (in_split) return BTC_TC_CALL(NN, true, false);          \
        if (out_split) return BTC_TC_CALL(NN, false, true);         \
        return BTC_TC_CALL(NN, false, false);                       \
    } while (0)
    switch (tc_padded_n(c_out)) {
        case 32: BTC_TC_FMT(32);
        case 64: BTC_TC_FMT(64);
        case 128: BTC_TC_FMT(128);
    }
#undef BTC_TC_FMT
#undef BTC_TC_CALL
    return BTC_E_UNSUPPORTED;
}

int btc_sparse_conv_fwd_tc(const float* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                           const float* scale, const float* shift, int relu, float* feat_out, int n_out_cap,
                           const int* n_out_dev, int K, int c_in, int c_out, void* stream) {
    return fwd_tc("btc_sparse_conv_fwd_tc", feat_in, nbr_out, packed_weight, bias, scale, shift, relu, feat_out, nullptr,
                  n_out_cap, n_out_dev, K, c_in, c_out, stream);
}

int btc_sparse_conv_fwd_tc_meta(const float* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                                const float* scale, const float* shift, int relu, float* feat_out, int n_out_cap,
                                const int* n_out_dev, int K, int c_in, int c_out, const uint64_t* tile_mask,
                                const int* tile_order, void* stream) {
    return fwd_tc("btc_sparse_conv_fwd_tc_meta", feat_in, nbr_out, packed_weight, bias, scale, shift, relu, feat_out, nullptr,
                  n_out_cap, n_out_dev, K, c_in, c_out, stream, (const unsigned long long*)tile_mask, tile_order);
}

int btc_sparse_conv_fwd_tc_split(const void* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                                 const float* scale, const float* shift, int relu, void* feat_out, int n_out_cap,
                                 const int* n_out_dev, int K, int c_in, int c_out, int in_split, int out_split,
                                 const uint64_t* tile_mask, const int* tile_order, void* stream) {
    return fwd_tc("btc_sparse_conv_fwd_tc_split", (const float*)feat_in, nbr_out, packed_weight, bias, scale, shift, relu,
                  (float*)feat_out, nullptr, n_out_cap, n_out_dev, K, c_in, c_out, stream, (const unsigned long long*)tile_mask,
                  tile_order, in_split ? 1 : 0, out_split ? 1 : 0);
}

int64_t btc_rulebook_tile_order_ints(int n_out_cap) {
    const int64_t tiles = (n_out_cap + TC_BM - 1) / TC_BM;
    return 65 + 65 * (tiles > 0 ? tiles : 1);
}

int btc_rulebook_tile_meta(const int* nbr_out, int n_out_cap, const int* n_out_dev, int K, uint64_t* tile_mask,
                           int* tile_order, void* stream) {
    if (K < 1 || K > 64 || n_out_cap < 0) return badarg("btc_rulebook_tile_meta: bad sizes");
    if (n_out_cap == 0) return BTC_OK;
    if (!nbr_out || !tile_mask) return badarg("btc_rulebook_tile_meta: null argument");
    const int tiles = (n_out_cap + TC_BM - 1) / TC_BM;
    cudaStream_t st = (cudaStream_t)stream;
    if (tile_order) BTC_CUDA(cudaMemsetAsync(tile_order, 0, 65 * sizeof(int), st), "tile_meta memset");
    const int grid = tiles < kNumSM * 16 ? tiles : kNumSM * 16;
    tile_meta_kernel<<<grid, 128, 0, st>>>(nbr_out, n_out_cap, n_out_dev, K, (unsigned long long*)tile_mask, tile_order, tiles);
    BTC_CHECK_LAUNCH("tile_meta");
    return BTC_OK;
}

int btc_sparse_conv_fwd_tc_rows(const float* feat_in, const int* nbr_sorted, const int* out_rows, const void* packed_weight,
                                const float* bias, const float* scale, const float* shift, int relu, float* feat_out,
                                int n_out_cap, const int* n_out_dev, int K, int c_in, int c_out, void* stream) {
    if (!out_rows) return badarg("btc_sparse_conv_fwd_tc_rows: null out_rows");
    return fwd_tc("btc_sparse_conv_fwd_tc_rows", feat_in, nbr_sorted, packed_weight, bias, scale, shift, relu, feat_out, out_rows,
                  n_out_cap, n_out_dev, K, c_in, c_out, stream);
}

}  // extern "C"
