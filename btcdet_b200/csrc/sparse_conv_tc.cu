// tcgen05 (5th-gen tensor core) gather-GEMM tile for the wide sparse-conv layers.
//
// Same output-stationary formulation as sparse_conv.cu (one CTA = 128 output rows, all K
// offsets walked in ascending order, each output row written once), but the
// [128 x 32] x [32 x N] products run on the tensor cores with the accumulator in TMEM:
//
//   * 4 producer warps gather the neighbour rows named by nbr_out[o][k] from global/L2 with
//     16-byte loads, split every fp32 value into a tf32 "hi" part (low 13 mantissa bits cleared)
//     and the exact fp32 remainder "lo", and store both as K-major, 128-byte-swizzled UMMA
//     operand tiles in shared memory (the gather cannot be a TMA tile: rows are arbitrary);
//   * the matching weight slice W[k][c0:c0+32][:] comes pre-packed (btc_sparse_conv_tc_pack) as
//     the exact shared-memory image of the K-major swizzled hi/lo tiles, so one elected thread
//     fetches it with a single cp.async.bulk (TMA 1-D bulk copy, complete_tx on the stage mbarrier);
//   * one elected thread of warp 4 issues tcgen05.mma.kind::tf32 (M=128, N, K=8) — 3xTF32:
//     D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi — fp32-class accuracy (the 1e-4 parity bar) from the
//     tf32 pipe; tcgen05.commit releases the smem stage / publishes the accumulator (mbarriers);
//   * the producer warps then read the accumulator back with tcgen05.ld (32 lanes x 16 columns per
//     instruction), apply bias / folded-BN affine / ReLU and write the output rows.
//
// Shared memory per stage: A_hi + A_lo (2 x 16 KB) + B_hi + B_lo (2 x N x 128 B).
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

// byte offset of element (row r, float j) inside a K-major SW128 tile
__host__ __device__ __forceinline__ int sw128_offset(int r, int j) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((j >> 2) ^ (r & 7)) & 7) << 4) + (j & 3) * 4;
}

// ---- weight pre-pack ------------------------------------------------------------------------------
// packed[k][chunk] = { B_hi image [N x 128 B] , B_lo image [N x 128 B] }, B = W[k]^T (N x Cin, K-major)
__global__ void tc_pack_weight_kernel(const float* __restrict__ w, int K, int c_in, int c_out, int N,
                                      float* __restrict__ packed) {
    const int nchunk = (c_in + TC_KC - 1) / TC_KC;
    const int64_t total = (int64_t)K * nchunk * N * TC_KC;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(t % TC_KC);
        int64_t r = t / TC_KC;
        int n = (int)(r % N);
        r /= N;
        int cc = (int)(r % nchunk);
        int k = (int)(r / nchunk);
        int ci = cc * TC_KC + j;
        float v = (ci < c_in && n < c_out) ? w[((int64_t)k * c_in + ci) * c_out + n] : 0.f;
        float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        float lo = v - hi;
        char* base = (char*)packed + ((int64_t)k * nchunk + cc) * (2 * N * 128);
        *(float*)(base + sw128_offset(n, j)) = hi;
        *(float*)(base + N * 128 + sw128_offset(n, j)) = lo;
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------
// Persistent, warp-specialised: one CTA per SM loops over 128-row output tiles.
//   warps 0-7  : two producer groups (4 warps each); group g fills the stages with (global stage index % 2 == g),
//                prefetching the next stage's gather into registers before it stores the current one;
//   warp  8    : MMA issuer (one elected lane);
//   warps 9-12 : epilogue (TMEM lane quarter = warp % 4), overlapping the next tile's main loop through a
//                double-buffered TMEM accumulator (2 x N columns).
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_MMA_WARP = 8;
constexpr int TC_EPI_WARP0 = 9;
constexpr int TC_PERSIST_THREADS = 13 * 32;

template <int N, int STAGES>
__global__ void __launch_bounds__(TC_PERSIST_THREADS, 1)
conv_fwd_tc_kernel(const float* __restrict__ feat_in, const int* __restrict__ table, int mirror,
                   const float* __restrict__ packed_w, const float* __restrict__ bias,
                   const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                   float* __restrict__ feat_out, int n_cap, const int* __restrict__ n_dev, int K, int c_in,
                   int c_out) {
    constexpr int A_BYTES = TC_BM * 128;          // one A tile (hi or lo)
    constexpr int B_BYTES = N * 128;              // one B tile (hi or lo)
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * N < 32 ? 32 : 2 * N;   // double-buffered accumulator
    // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* stages = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1 KB aligned

    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t s_tmem;

    const int n = live_count(n_cap, n_dev);
    if ((int)blockIdx.x * TC_BM >= n) return;     // no tile for this CTA (whole CTA leaves before any barrier)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_tiles = (n + TC_BM - 1) / TC_BM;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nchunk = (c_in + TC_KC - 1) / TC_KC;
    const int T = K * nchunk;                      // stages per tile
    const int total_stages = my_tiles * T;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 128);          // the 128 threads of one producer group (+ bulk-copy tx bytes)
            mbar_init(&empty_bar[s], 1);           // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);           // one tcgen05.commit
            mbar_init(&tmem_empty[b], 128);        // the 128 epilogue threads
        }
        fence_mbar_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc(&s_tmem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < TC_PRODUCER_WARPS) {
        // ================= producers =================
        const int group = warp >> 2, wg = warp & 3;
        const int sub = lane >> 3;       // row within the 4-row group handled per load instruction
        const int q = lane & 7;          // 16-byte chunk within the 128-byte row
        float4 cur[8], nxt[8];
        auto gather = [&](int gi, float4 (&v)[8]) {
            const int tl = gi / T, j = gi - tl * T;
            const int k = j / nchunk, cc = j - k * nchunk;
            const int row0 = ((int)blockIdx.x + tl * (int)gridDim.x) * TC_BM;
            const int c = cc * TC_KC + q * 4;
            const int kk = mirror ? K - 1 - k : k;
            int src[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int r = row0 + wg * 32 + g * 4 + sub;
                src[g] = r < n ? __ldg(table + (int64_t)r * K + kk) : -1;
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                v[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src[g] >= 0 && c < c_in) v[g] = __ldg(reinterpret_cast<const float4*>(feat_in + (int64_t)src[g] * c_in + c));
            }
        };
        int gi = group;
        if (gi < total_stages) gather(gi, cur);
        for (; gi < total_stages; gi += 2) {
            if (gi + 2 < total_stages) gather(gi + 2, nxt);   // next stage's loads fly while this one is stored
            const int s = gi % STAGES;
            const uint32_t ph = (gi / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            unsigned char* st = stages + s * STAGE_BYTES;
            if ((tid & 127) == 0) {
                const int j = gi % T;
                mbar_expect_tx(&full_bar[s], 2 * B_BYTES);
                bulk_copy_g2s(st + 2 * A_BYTES, (const char*)packed_w + (int64_t)j * (2 * B_BYTES), 2 * B_BYTES, &full_bar[s]);
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int r = wg * 32 + g * 4 + sub;
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(cur[g].x) & 0xFFFFE000u);
                hi.y = __uint_as_float(__float_as_uint(cur[g].y) & 0xFFFFE000u);
                hi.z = __uint_as_float(__float_as_uint(cur[g].z) & 0xFFFFE000u);
                hi.w = __uint_as_float(__float_as_uint(cur[g].w) & 0xFFFFE000u);
                lo.x = cur[g].x - hi.x;
                lo.y = cur[g].y - hi.y;
                lo.z = cur[g].z - hi.z;
                lo.w = cur[g].w - hi.w;
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(st + off) = hi;
                *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
            }
            fence_proxy_async();                    // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&full_bar[s]);
#pragma unroll
            for (int g = 0; g < 8; ++g) cur[g] = nxt[g];
        }
    } else if (warp == TC_MMA_WARP) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int gi = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int buf = tl & 1;
                mbar_wait(&tmem_empty[buf], ((tl >> 1) & 1) ^ 1);    // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                for (int j = 0; j < T; ++j, ++gi) {
                    const int s = gi % STAGES;
                    const uint32_t ph = (gi / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stages + s * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES;
                    const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TC_KC / 8; ++kk) {   // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
                        const uint64_t da_hi = make_desc_sw128(a_hi + kk * 32), da_lo = make_desc_sw128(a_lo + kk * 32);
                        const uint64_t db_hi = make_desc_sw128(b_hi + kk * 32), db_lo = make_desc_sw128(b_lo + kk * 32);
                        umma_tf32(d_tmem, da_lo, db_hi, IDESC, (j | kk) != 0);   // small terms first
                        umma_tf32(d_tmem, da_hi, db_lo, IDESC, 1u);
                        umma_tf32(d_tmem, da_hi, db_hi, IDESC, 1u);
                    }
                    umma_commit(&empty_bar[s]);      // frees the stage once the MMAs above retire
                }
                umma_commit(&tmem_full[buf]);        // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        const int quarter = warp & 3;                // TMEM lanes [32*quarter, 32*quarter+32)
        for (int tl = 0; tl < my_tiles; ++tl) {
            const int buf = tl & 1;
            mbar_wait(&tmem_full[buf], (tl >> 1) & 1);
            tc_fence_after();
            const int row = ((int)blockIdx.x + tl * (int)gridDim.x) * TC_BM + quarter * 32 + lane;
            float* dst = feat_out + (int64_t)row * c_out;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N + c0), acc);
                if (row < n && c0 < c_out) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float o[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int col = c0 + j4 * 4 + jj;
                            float x = __uint_as_float(acc[j4 * 4 + jj]);
                            if (col < c_out) {
                                if (bias) x += __ldg(bias + col);
                                if (scale) x = x * __ldg(scale + col) + __ldg(shift + col);
                                if (relu) x = fmaxf(x, 0.f);
                            }
                            o[jj] = x;
                        }
                        const int col = c0 + j4 * 4;
                        if (col + 3 < c_out) {
                            *reinterpret_cast<float4*>(dst + col) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (col + jj < c_out) dst[col + jj] = o[jj];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);           // this accumulator buffer may be overwritten
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int N, int STAGES>
static int launch_tc(const float* feat_in, const int* table, int mirror, const float* packed_w, const float* bias,
                     const float* scale, const float* shift, int relu, float* feat_out, int n_cap, const int* n_dev,
                     int K, int c_in, int c_out, cudaStream_t st) {
    constexpr int STAGE_BYTES = 2 * TC_BM * 128 + 2 * N * 128;
    size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    auto kern = conv_fwd_tc_kernel<N, STAGES>;
    static size_t attr_set = 0;   // opt in to > 48 KB dynamic smem once per instantiation (not a stream op)
    if (attr_set < smem) {
        BTC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tc smem attr");
        attr_set = smem;
    }
    int tiles = (n_cap + TC_BM - 1) / TC_BM;
    dim3 grid(tiles < kNumSM ? tiles : kNumSM);    // persistent: one CTA per SM
    kern<<<grid, TC_PERSIST_THREADS, smem, st>>>(feat_in, table, mirror, packed_w, bias, scale, shift, relu, feat_out,
                                                 n_cap, n_dev, K, c_in, c_out);
    BTC_CHECK_LAUNCH("conv_fwd_tc");
    return BTC_OK;
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

int btc_sparse_conv_tc_supported(int K, int c_in, int c_out) {
    return (K >= 1 && K <= 64 && c_in >= 16 && c_in % 4 == 0 && c_out >= 16 && c_out % 4 == 0 && tc_padded_n(c_out) != 0) ? 1 : 0;
}

int64_t btc_sparse_conv_tc_packed_bytes(int K, int c_in, int c_out) {
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return BTC_E_UNSUPPORTED;
    int N = tc_padded_n(c_out);
    int nchunk = (c_in + TC_KC - 1) / TC_KC;
    return (int64_t)K * nchunk * 2 * N * 128;
}

int btc_sparse_conv_tc_pack(const float* weight, int K, int c_in, int c_out, void* packed, void* stream) {
    if (!weight || !packed) return badarg("btc_sparse_conv_tc_pack: null argument");
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_tc_pack: shape not supported", cudaSuccess);
    int N = tc_padded_n(c_out);
    int nchunk = (c_in + TC_KC - 1) / TC_KC;
    int64_t total = (int64_t)K * nchunk * N * TC_KC;
    tc_pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, K, c_in, c_out, N, (float*)packed);
    BTC_CHECK_LAUNCH("tc_pack_weight");
    return BTC_OK;
}

int btc_sparse_conv_fwd_tc(const float* feat_in, const int* nbr_out, const void* packed_weight, const float* bias,
                           const float* scale, const float* shift, int relu, float* feat_out, int n_out_cap,
                           const int* n_out_dev, int K, int c_in, int c_out, void* stream) {
    if (!nbr_out || !packed_weight || !feat_out) return badarg("btc_sparse_conv_fwd_tc: null argument");
    if ((scale == nullptr) != (shift == nullptr)) return badarg("btc_sparse_conv_fwd_tc: scale/shift must come together");
    if (!btc_sparse_conv_tc_supported(K, c_in, c_out)) return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_fwd_tc: shape not supported", cudaSuccess);
    if (n_out_cap <= 0) return BTC_OK;
    if (!feat_in) return badarg("btc_sparse_conv_fwd_tc: null feat_in");
    if (((uintptr_t)feat_in & 15) || ((uintptr_t)feat_out & 15) || ((uintptr_t)packed_weight & 15))
        return badarg("btc_sparse_conv_fwd_tc: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const float* pw = (const float*)packed_weight;
    switch (tc_padded_n(c_out)) {
        case 32: return launch_tc<32, 4>(feat_in, nbr_out, 0, pw, bias, scale, shift, relu, feat_out, n_out_cap, n_out_dev, K, c_in, c_out, st);
        case 64: return launch_tc<64, 4>(feat_in, nbr_out, 0, pw, bias, scale, shift, relu, feat_out, n_out_cap, n_out_dev, K, c_in, c_out, st);
        case 128: return launch_tc<128, 3>(feat_in, nbr_out, 0, pw, bias, scale, shift, relu, feat_out, n_out_cap, n_out_dev, K, c_in, c_out, st);
    }
    return BTC_E_UNSUPPORTED;
}

}  // extern "C"
