// Micro-benchmark: issue-to-retire cost of tcgen05.mma (kind::tf32 / kind::f16) on sm_100a as a
// function of N and of the accumulator dependency pattern.  One CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
// Modes 6 / 7 (added at the end of round 1, not run on hardware yet) probe the hypothesis of DESIGN.md §8.1(0):
//   6: elect-uniform TS issue with a tcgen05.commit after every `group` MMAs (rotating mbarriers nobody waits on):
//      if a commit drains the MMA pipeline, cycles/MMA rises as `group` shrinks;
//   7: mode 6 while warps 1-3 keep writing other TMEM columns with tcgen05.st (the producers' traffic).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>  // 0 tf32, 1 f16(bf16)
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = SS one accumulator, 1 = SS two accumulators alternating, 2 = TS one accumulator, 3 = TS two accumulators
template <int KIND>
__global__ void __launch_bounds__(128) bench(int N, int mode, int iters, long long* out, int group = 12) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t ring[8];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i % 7);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_stop = 0;
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tm = s_tmem;
    if (mode == 8 || mode == 9) {
        // SS-form probe (round 2): A and B both K-major SW128 tiles in shared memory, elect-uniform issue.
        //   N2 == 0: every MMA is A(128 x K) x B(N x K).   N2 > 0 ("CAT" 3-term pattern of a split-precision product):
        //   alternate A_hi x [B_hi|B_lo] (width N) and A_lo x B_hi (width N2) -- two different A tiles.
        // mode 9: warps 1-3 stream st.shared.v4 into another region meanwhile (stand-in for the cp.async / TMA fill traffic).
        const int N2 = group;   // re-used parameter
        if (threadIdx.x < 32) {
            uint32_t fmt = KIND == 0 ? 2u : 1u;
            uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((N2 ? N2 : N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint64_t da = make_desc_sw128(smem_u32(smem)), da2 = make_desc_sw128(smem_u32(smem + 16384));
            uint64_t db = make_desc_sw128(smem_u32(smem + 32768));
            long long t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    uint32_t pred = 0;
                    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
                    if (pred) {
                        if (N2 && (u & 1)) mma_ss<KIND>(tm, da2 + ((u >> 1) & 3) * 2, db + ((u >> 1) & 3) * 2, idesc2, 1u);
                        else mma_ss<KIND>(tm, da + ((u >> 1) & 3) * 2, db + ((u >> 1) & 3) * 2, idesc, 1u);
                    }
                }
            }
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            } while (!done);
            long long t1 = clock64();
            if (threadIdx.x == 0) { out[0] = t1 - t0; s_stop = 1; }
        } else if (mode == 9) {
            const uint32_t dst = smem_u32(smem + 65536) + (threadIdx.x - 32) * 16;
            long long n = 0;
            while (!s_stop) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst + r * 1536), "r"(threadIdx.x) : "memory");
                ++n;
            }
            if (threadIdx.x == 32) out[1] = n * 8 * 96 * 16;   // bytes written by the three warps
        }
    } else
    if (mode == 6 || mode == 7) {
        if (threadIdx.x < 32) {
            uint32_t fmt = KIND == 0 ? 2u : 1u;
            uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint64_t db = make_desc_sw128(smem_u32(smem + 32768));
            long long t0 = clock64();
            int since = 0, slot = 0;
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t pred = 0;
                    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
                    if (pred) mma_ts<KIND>(tm, tm + 384 + (u & 3) * 8, db + (u & 3) * 2, idesc, 1u);
                }
                since += 4;
                if (since >= group) {
                    since = 0;
                    uint32_t pred = 0;
                    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
                    if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ring[slot])) : "memory");
                    slot = (slot + 1) & 7;
                }
            }
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            } while (!done);
            long long t1 = clock64();
            if (threadIdx.x == 0) { out[0] = t1 - t0; s_stop = 1; }
        } else if (mode == 7) {
            // warps 1-3: keep storing 32 columns of their TMEM lane quarter (columns 256..287, not touched by the MMAs)
            const uint32_t lane_base = tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 256u;
            while (!s_stop) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(lane_base), "r"(threadIdx.x) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(lane_base + 16u), "r"(threadIdx.x) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
    } else
    if (mode == 5) {
        // warp-uniform issue loop: the whole warp runs the loop, one ELECTed lane issues (CUTLASS style)
        if (threadIdx.x < 32) {
            uint32_t fmt = KIND == 0 ? 2u : 1u;
            uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint64_t db = make_desc_sw128(smem_u32(smem + 32768));
            long long t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    uint32_t pred = 0;
                    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
                    if (pred) mma_ts<KIND>(tm, tm + 384 + (u & 3) * 8, db + (u & 3) * 2, idesc, 1u);
                }
            }
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            } while (!done);
            long long t1 = clock64();
            if (threadIdx.x == 0) out[0] = t1 - t0;
        }
    } else
    if (threadIdx.x == 0) {
        uint32_t fmt = KIND == 0 ? 2u : 1u;  // tf32 / bf16
        uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t da = make_desc_sw128(smem_u32(smem)), db = make_desc_sw128(smem_u32(smem + 32768));
        long long t0 = clock64();
        if (mode < 4) {
        for (int i = 0; i < iters; ++i) {
            uint32_t d = tm + ((mode & 1) ? (uint32_t)((i & 1) * 256) : 0u);
            uint32_t kk = (i & 3) * 2;   // walk the 4 k-steps of a 128-byte row
            if (mode < 2) mma_ss<KIND>(d, da + kk, db + kk, idesc, 1u);
            else mma_ts<KIND>(d, tm + 384 + (i & 3) * 8, db + kk, idesc, 1u);
        }
        } else {   // mode 4: TS, fully unrolled x8, constant operands (pure issue-rate probe)
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma_ts<KIND>(tm, tm + 384 + (u & 3) * 8, db + (u & 3) * 2, idesc, 1u);
        }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
    size_t smem = 65536 + 16384 + 1024;   // A tiles at 0 / 16 KB, B tile at 32 KB (up to 256 rows), writer scratch at 64 KB
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 4096;
    const char* names[6] = {"SS 1acc", "SS 2acc", "TS 1acc", "TS 2acc", "TS unrolled", "TS elect-uniform"};
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {64, 128, 192, 256})
            for (int mode = 4; mode < 6; ++mode) {
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    if (kind == 0) bench<0><<<1, 128, smem>>>(N, mode, iters, d_out);
                    else bench<1><<<1, 128, smem>>>(N, mode, iters, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
                }
                double cyc = (double)h / iters;
                int K = kind == 0 ? 8 : 16;
                printf("%s M=128 N=%3d K=%2d %s: %7.1f cycles/MMA  -> %7.0f MAC/clk\n", kind == 0 ? "tf32" : "bf16", N, K, names[mode], cyc, 128.0 * N * K / cyc);
            }
    // commit-drain probe (modes 6 / 7), tf32 only
    for (int N : {32, 64})
        for (int mode = 6; mode < 8; ++mode)
            for (int group : {4, 12, 24, 48, 4096}) {
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    bench<0><<<1, 128, smem>>>(N, mode, iters, d_out, group);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
                }
                printf("tf32 M=128 N=%3d K= 8 TS elect-uniform, commit every %4d MMAs%s: %7.1f cycles/MMA\n", N, group,
                       mode == 7 ? " + tcgen05.st traffic" : "", (double)h / iters);
            }
    // SS-form probe (round 2): cycles per MMA with both operands in shared memory, with and without concurrent smem stores
    for (int kind = 0; kind < 2; ++kind)
        for (int mode = 8; mode < 10; ++mode)
            for (int cfg = 0; cfg < 6; ++cfg) {
                const int Ns[6] = {32, 64, 128, 256, 128, 64}, N2s[6] = {0, 0, 0, 0, 64, 32};
                long long h[2] = {0, 0};
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemset(d_out, 0, 16);
                    if (kind == 0) bench<0><<<1, 128, smem>>>(Ns[cfg], mode, iters, d_out, N2s[cfg]);
                    else bench<1><<<1, 128, smem>>>(Ns[cfg], mode, iters, d_out, N2s[cfg]);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
                }
                double cyc = (double)h[0] / iters;
                const int K = kind == 0 ? 8 : 16, eb = kind == 0 ? 4 : 2;
                // operand bytes read from shared memory per MMA (average over the alternating pattern)
                double bytes = N2s[cfg] ? 0.5 * ((128 + Ns[cfg]) + (128 + N2s[cfg])) * K * eb : (128.0 + Ns[cfg]) * K * eb;
                printf("%s SS M=128 N=%3d%s K=%2d%s: %7.1f cycles/MMA, operand reads %6.1f B/clk", kind == 0 ? "tf32" : "bf16", Ns[cfg],
                       N2s[cfg] ? (N2s[cfg] == 64 ? "/64 alternating" : "/32 alternating") : "", K,
                       mode == 9 ? " + concurrent st.shared" : "", cyc, bytes / cyc);
                if (mode == 9) printf(", stores %6.1f B/clk", (double)h[1] / (double)h[0]);
                printf("\n");
            }
    return 0;
}
