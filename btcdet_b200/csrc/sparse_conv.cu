// Sparse convolution arithmetic — replaces spconv 1.2.1 ops.indice_conv / indice_conv_backward
// (src/spconv/spconv_ops.cc: for each of the K offsets {gather rows -> torch::mm (cuBLAS SGEMM)
// -> scatter-add}; reordering.cu gather/scatter kernels), i.e. 3*K launches and
// 2*P*(Cin+Cout)*4 bytes of gather/scatter buffer traffic per layer.
//
// Here: ONE output-stationary gather-GEMM kernel per layer.  A CTA owns BM output rows x BN
// output channels; it walks the K offsets in ascending order (the summation order of the
// reference's Native algorithm), gathers the neighbour rows named by nbr_out[o][k] straight
// into shared memory (16-byte vector loads, zero fill for missing neighbours), multiplies by
// W[k] staged in shared memory and accumulates in registers.  Every output row is written
// exactly once with the bias / folded-BatchNorm affine / ReLU epilogue fused: no scatter, no
// atomics, bit-reproducible.  Offsets with no neighbour in the tile are skipped.
// This file holds the fp32 FFMA tile (exact fp32 products/sums, used for parity and for thin
// layers); sparse_conv_tc.cu holds the tcgen05 tensor-core tile for wide layers.
#include "common.cuh"

namespace btc {

constexpr int CK = 16;  // input-channel chunk staged per step

template <int BM, int BN, int TM, int TN, bool VEC_A, bool VEC_W>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_fwd_ffma_kernel(const float* __restrict__ feat_in, const int* __restrict__ table, int mirror,
                     const float* __restrict__ weight, const float* __restrict__ bias,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                     float* __restrict__ feat_out, int n_cap, const int* __restrict__ n_dev, int K, int c_in,
                     int c_out) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int AS = BM + 4;  // row stride of the transposed A tile (keeps float4 alignment)
    static_assert(TM % 4 == 0 && TN % 4 == 0, "thread tile must be float4 friendly");
    static_assert((BM * CK / 4) % NT == 0, "A tile must divide across threads");
    constexpr int A_LD = (BM * CK / 4) / NT;                       // float4 loads of A per thread
    constexpr int W_LD = (CK * BN / 4 + NT - 1) / NT;              // float4 loads of W per thread

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                // [2][CK][AS]
    float* Ws = As + 2 * CK * AS;                                  // [2][CK][BN]
    int* nbr_s = reinterpret_cast<int*>(Ws + 2 * CK * BN);         // [BM][K]
    int* klist = nbr_s + BM * K;                                   // [K]
    __shared__ int s_nk;
    __shared__ unsigned s_kmask[8];                                // K <= 256

    const int n = live_count(n_cap, n_dev);
    const int col0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    // persistent over row tiles: the grid is sized by the SM count, not by the (capacity) row count
    for (int row0 = blockIdx.x * BM; row0 < n; row0 += gridDim.x * BM) {
    __syncthreads();   // previous tile's readers of nbr_s / klist / smem tiles are done
    if (tid < 8) s_kmask[tid] = 0u;
    __syncthreads();
    // neighbour tile (coalesced: the BM x K block is contiguous in the table)
    for (int e = tid; e < BM * K; e += NT) {
        int r = e / K, k = e - r * K;
        int v = -1;
        if (row0 + r < n) v = __ldg(table + (int64_t)(row0 + r) * K + (mirror ? K - 1 - k : k));
        nbr_s[e] = v;
        if (v >= 0) atomicOr(&s_kmask[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    if (tid == 0) {
        int nk = 0;
        for (int k = 0; k < K; ++k)
            if (s_kmask[k >> 5] & (1u << (k & 31))) klist[nk++] = k;
        s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk;
    const int nchunk = (c_in + CK - 1) / CK;
    const int T = nk * nchunk;

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 a_reg[A_LD];
    float4 w_reg[W_LD];

    auto load_regs = [&](int it) {
        const int k = klist[it / nchunk];
        const int c0 = (it % nchunk) * CK;
#pragma unroll
        for (int j = 0; j < A_LD; ++j) {
            int idx = tid + j * NT;
            int r = idx / (CK / 4), q = idx % (CK / 4);
            int src = nbr_s[r * K + k];
            int c = c0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src >= 0) {
                const float* p = feat_in + (int64_t)src * c_in + c;
                if (VEC_A && c + 3 < c_in) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (c < c_in) v.x = __ldg(p);
                    if (c + 1 < c_in) v.y = __ldg(p + 1);
                    if (c + 2 < c_in) v.z = __ldg(p + 2);
                    if (c + 3 < c_in) v.w = __ldg(p + 3);
                }
            }
            a_reg[j] = v;
        }
#pragma unroll
        for (int j = 0; j < W_LD; ++j) {
            int idx = tid + j * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < CK * BN / 4) {
                int cc = idx / (BN / 4), q = idx % (BN / 4);
                int c = c0 + cc, nn = col0 + q * 4;
                if (c < c_in) {
                    const float* p = weight + ((int64_t)k * c_in + c) * c_out + nn;
                    if (VEC_W && nn + 3 < c_out) {
                        v = __ldg(reinterpret_cast<const float4*>(p));
                    } else {
                        if (nn < c_out) v.x = __ldg(p);
                        if (nn + 1 < c_out) v.y = __ldg(p + 1);
                        if (nn + 2 < c_out) v.z = __ldg(p + 2);
                        if (nn + 3 < c_out) v.w = __ldg(p + 3);
                    }
                }
            }
            w_reg[j] = v;
        }
    };
    auto store_smem = [&](int buf) {
        float* a = As + buf * CK * AS;
        float* w = Ws + buf * CK * BN;
#pragma unroll
        for (int j = 0; j < A_LD; ++j) {
            int idx = tid + j * NT;
            int r = idx / (CK / 4), q = idx % (CK / 4);
            a[(q * 4 + 0) * AS + r] = a_reg[j].x;
            a[(q * 4 + 1) * AS + r] = a_reg[j].y;
            a[(q * 4 + 2) * AS + r] = a_reg[j].z;
            a[(q * 4 + 3) * AS + r] = a_reg[j].w;
        }
#pragma unroll
        for (int j = 0; j < W_LD; ++j) {
            int idx = tid + j * NT;
            if (idx < CK * BN / 4) reinterpret_cast<float4*>(w)[idx] = w_reg[j];
        }
    };

    if (T > 0) {
        load_regs(0);
        store_smem(0);
    }
    __syncthreads();
    for (int it = 0; it < T; ++it) {
        if (it + 1 < T) load_regs(it + 1);
        const float* a = As + (it & 1) * CK * AS + ty * TM;
        const float* w = Ws + (it & 1) * CK * BN + tx * TN;
#pragma unroll
        for (int c = 0; c < CK; ++c) {
            float av[TM], wv[TN];
#pragma unroll
            for (int i = 0; i < TM / 4; ++i) {
                float4 t = *reinterpret_cast<const float4*>(a + c * AS + i * 4);
                av[i * 4 + 0] = t.x; av[i * 4 + 1] = t.y; av[i * 4 + 2] = t.z; av[i * 4 + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN / 4; ++j) {
                float4 t = *reinterpret_cast<const float4*>(w + c * BN + j * 4);
                wv[j * 4 + 0] = t.x; wv[j * 4 + 1] = t.y; wv[j * 4 + 2] = t.z; wv[j * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (it + 1 < T) store_smem((it + 1) & 1);
        __syncthreads();
    }

    // epilogue: bias -> affine -> relu -> store (each output row written exactly once)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int row = row0 + ty * TM + i;
        if (row >= n) continue;
        float* dst = feat_out + (int64_t)row * c_out;
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
            float v[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                int col = col0 + tx * TN + j4 * 4 + jj;
                float x = acc[i][j4 * 4 + jj];
                if (col < c_out) {
                    if (bias) x += __ldg(bias + col);
                    if (scale) x = x * __ldg(scale + col) + __ldg(shift + col);
                    if (relu) x = fmaxf(x, 0.f);
                }
                v[jj] = x;
            }
            int col = col0 + tx * TN + j4 * 4;
            if (VEC_W && col + 3 < c_out) {
                *reinterpret_cast<float4*>(dst + col) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (col + jj < c_out) dst[col + jj] = v[jj];
            }
        }
    }
    }  // row tiles
}

template <int BM, int BN, int TM, int TN>
static int launch_fwd(const float* feat_in, const int* table, int mirror, const float* weight, const float* bias,
                      const float* scale, const float* shift, int relu, float* feat_out, int n_cap, const int* n_dev,
                      int K, int c_in, int c_out, cudaStream_t st) {
    constexpr int NT = (BM / TM) * (BN / TN);
    size_t smem = (size_t)(2 * CK * (BM + 4) + 2 * CK * BN) * sizeof(float) + (size_t)(BM * K + K) * sizeof(int);
    int tiles = (n_cap + BM - 1) / BM;
    dim3 grid(tiles < 8 * kNumSM ? tiles : 8 * kNumSM, (c_out + BN - 1) / BN);
    const bool va = (c_in % 4 == 0) && (((uintptr_t)feat_in & 15) == 0);
    const bool vw = (c_out % 4 == 0) && (((uintptr_t)weight & 15) == 0) && (((uintptr_t)feat_out & 15) == 0);
#define BTC_LAUNCH(VA, VW)                                                                                         \
    do {                                                                                                           \
        auto kern = conv_fwd_ffma_kernel<BM, BN, TM, TN, VA, VW>;                                                  \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        kern<<<grid, NT, smem, st>>>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap,     \
                                     n_dev, K, c_in, c_out);                                                       \
    } while (0)
    if (va && vw) BTC_LAUNCH(true, true);
    else if (va) BTC_LAUNCH(true, false);
    else if (vw) BTC_LAUNCH(false, true);
    else BTC_LAUNCH(false, false);
#undef BTC_LAUNCH
    BTC_CHECK_LAUNCH("conv_fwd_ffma");
    return BTC_OK;
}

int conv_fwd_ffma(const float* feat_in, const int* table, int mirror, const float* weight, const float* bias,
                  const float* scale, const float* shift, int relu, float* feat_out, int n_cap, const int* n_dev, int K,
                  int c_in, int c_out, cudaStream_t st) {
    if (n_cap <= 0) return BTC_OK;
    // Small problems under-fill 148 SMs with 128-row tiles: use 64-row tiles then.
    const bool small = ((int64_t)(n_cap + 127) / 128) * ((c_out + 63) / 64) < 2 * kNumSM;
    if (c_out <= 16) {
        return small ? launch_fwd<64, 16, 4, 4>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st)
                     : launch_fwd<128, 16, 4, 4>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st);
    } else if (c_out <= 32) {
        return small ? launch_fwd<64, 32, 4, 4>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st)
                     : launch_fwd<128, 32, 8, 4>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st);
    } else {
        return small ? launch_fwd<64, 64, 4, 8>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st)
                     : launch_fwd<128, 64, 8, 8>(feat_in, table, mirror, weight, bias, scale, shift, relu, feat_out, n_cap, n_dev, K, c_in, c_out, st);
    }
}

// ---- backward ---------------------------------------------------------------------------------
__global__ void transpose_weight_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int c_in, int c_out) {
    const int64_t total = (int64_t)K * c_in * c_out;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int co = (int)(t % c_out);
        int64_t r = t / c_out;
        int ci = (int)(r % c_in);
        int k = (int)(r / c_in);
        wt[((int64_t)k * c_out + co) * c_in + ci] = w[t];
    }
}

// dW[k] += A_k^T (x) dO over a slab of output rows.  grid (row slabs, K, channel tiles).
// CTA tile: 64 (c_in) x 64 (c_out), 256 threads, 4x4 per thread, rows staged 32 at a time.
constexpr int BW_ROWS = 32;
constexpr int BW_T = 64;
__global__ void __launch_bounds__(256)
conv_bwd_weight_kernel(const float* __restrict__ feat_in, const float* __restrict__ d_out,
                       const int* __restrict__ nbr_out, float* __restrict__ d_weight, int n_cap,
                       const int* __restrict__ n_dev, int rows_per_slab, int K, int c_in, int c_out) {
    __shared__ float As[BW_ROWS][BW_T + 4];
    __shared__ float Ds[BW_ROWS][BW_T + 4];
    __shared__ int s_src[BW_ROWS];
    const int n = live_count(n_cap, n_dev);
    const int k = blockIdx.y;
    const int tiles_co = (c_out + BW_T - 1) / BW_T;
    const int ci0 = (blockIdx.z / tiles_co) * BW_T, co0 = (blockIdx.z % tiles_co) * BW_T;
    const int r_begin = blockIdx.x * rows_per_slab;
    int r_end = r_begin + rows_per_slab;
    if (r_end > n) r_end = n;
    const int tid = threadIdx.x;
    const int ti = tid / 16, tj = tid % 16;  // 16 x 16 threads, each 4 (ci) x 4 (co)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += BW_ROWS) {
        int src_t = -1;
        if (tid < BW_ROWS) {
            int r = r0 + tid;
            src_t = r < r_end ? __ldg(nbr_out + (int64_t)r * K + k) : -1;
            s_src[tid] = src_t;
        }
        // barrier + block-wide "any valid row for this offset in the slab?"
        if (!__syncthreads_or(src_t >= 0)) continue;
        for (int e = tid; e < BW_ROWS * BW_T; e += 256) {
            int rr = e / BW_T, c = e % BW_T;
            int src = s_src[rr];
            float a = 0.f, d = 0.f;
            if (src >= 0) {
                if (ci0 + c < c_in) a = __ldg(feat_in + (int64_t)src * c_in + ci0 + c);
                if (co0 + c < c_out) d = __ldg(d_out + (int64_t)(r0 + rr) * c_out + co0 + c);
            }
            As[rr][c] = a;
            Ds[rr][c] = d;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < BW_ROWS; ++rr) {
            float4 a = *reinterpret_cast<const float4*>(&As[rr][ti * 4]);
            float4 d = *reinterpret_cast<const float4*>(&Ds[rr][tj * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = ci0 + ti * 4 + i;
        if (ci >= c_in) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = co0 + tj * 4 + j;
            if (co < c_out && acc[i][j] != 0.f) atomicAdd(d_weight + ((int64_t)k * c_in + ci) * c_out + co, acc[i][j]);
        }
    }
}

__global__ void col_sum_kernel(const float* __restrict__ x, int n_cap, const int* __restrict__ n_dev, int c,
                               float* __restrict__ out) {
    // grid.x = row slabs; each thread owns one channel (c <= blockDim.x) for its slab
    const int n = live_count(n_cap, n_dev);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = 0.f;
        for (int r = blockIdx.x; r < n; r += gridDim.x) s += __ldg(x + (int64_t)r * c + ch);
        atomicAdd(out + ch, s);
    }
}

}  // namespace btc

using namespace btc;

extern "C" {

int btc_sparse_conv_fwd(const float* feat_in, const int* nbr_out, const float* weight, const float* bias,
                        const float* scale, const float* shift, int relu, float* feat_out, int n_out_cap,
                        const int* n_out_dev, int K, int c_in, int c_out, int algo, void* stream) {
    if (K < 1 || K > 256 || c_in < 1 || c_out < 1 || n_out_cap < 0) return badarg("btc_sparse_conv_fwd: bad sizes");
    if (n_out_cap == 0) return BTC_OK;   // empty output (torch hands out null data pointers for 0-row tensors): nothing to do
    if (!nbr_out || !weight || !feat_out) return badarg("btc_sparse_conv_fwd: null argument");
    if ((scale == nullptr) != (shift == nullptr)) return badarg("btc_sparse_conv_fwd: scale/shift must come together");
    if (n_out_cap > 0 && !feat_in) return badarg("btc_sparse_conv_fwd: null feat_in");
    cudaStream_t st = (cudaStream_t)stream;
    if (algo != 0 && algo != 1)
        return set_error(BTC_E_UNSUPPORTED, "btc_sparse_conv_fwd: tensor-core tiles take pre-packed weights, use btc_sparse_conv_fwd_tc", cudaSuccess);
    return conv_fwd_ffma(feat_in, nbr_out, 0, weight, bias, scale, shift, relu, feat_out, n_out_cap, n_out_dev, K, c_in,
                         c_out, st);
}

int64_t btc_sparse_conv_bwd_workspace_bytes(int K, int c_in, int c_out) {
    return align_up((int64_t)K * c_in * c_out * 4, 256);
}

int btc_sparse_conv_bwd_data(const float* d_out, const int* table, int mirror, const float* weight, float* d_in,
                             int n_in_cap, const int* n_in_dev, int K, int c_in, int c_out, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    if (n_in_cap <= 0) return BTC_OK;    // no input rows: no gradient rows
    if (!table || !weight || !d_in || !workspace) return badarg("btc_sparse_conv_bwd_data: null argument");
    if (workspace_bytes < btc_sparse_conv_bwd_workspace_bytes(K, c_in, c_out)) return badarg("btc_sparse_conv_bwd_data: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* wt = (float*)workspace;
    int64_t total = (int64_t)K * c_in * c_out;
    transpose_weight_kernel<<<grid_for(total, 256), 256, 0, st>>>(weight, wt, K, c_in, c_out);
    BTC_CHECK_LAUNCH("transpose_weight");
    // d_in = gather-GEMM over the input-major table with W^T: "c_in" of that product is c_out.
    return conv_fwd_ffma(d_out, table, mirror, wt, nullptr, nullptr, nullptr, 0, d_in, n_in_cap, n_in_dev, K, c_out, c_in, st);
}

int btc_sparse_conv_bwd_weight(const float* feat_in, const float* d_out, const int* nbr_out, float* d_weight,
                               float* d_bias, int n_out_cap, const int* n_out_dev, int K, int c_in, int c_out,
                               void* stream) {
    if (!d_weight || (n_out_cap > 0 && !nbr_out)) return badarg("btc_sparse_conv_bwd_weight: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    BTC_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)K * c_in * c_out * 4, st), "bwd_weight memset");
    if (d_bias) BTC_CUDA(cudaMemsetAsync(d_bias, 0, (size_t)c_out * 4, st), "bwd_bias memset");
    if (n_out_cap <= 0) return BTC_OK;
    if (!feat_in || !d_out) return badarg("btc_sparse_conv_bwd_weight: null features");
    int tiles = ((c_in + BW_T - 1) / BW_T) * ((c_out + BW_T - 1) / BW_T);
    // enough slabs to fill the machine: K * tiles CTAs per slab
    int slabs = (4 * kNumSM + K * tiles - 1) / (K * tiles);
    int max_slabs = (n_out_cap + BW_ROWS - 1) / BW_ROWS;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    int rows_per_slab = ((n_out_cap + slabs - 1) / slabs + BW_ROWS - 1) / BW_ROWS * BW_ROWS;
    slabs = (n_out_cap + rows_per_slab - 1) / rows_per_slab;
    dim3 grid(slabs, K, tiles);
    conv_bwd_weight_kernel<<<grid, 256, 0, st>>>(feat_in, d_out, nbr_out, d_weight, n_out_cap, n_out_dev, rows_per_slab, K,
                                                 c_in, c_out);
    if (d_bias) col_sum_kernel<<<kNumSM, 128, 0, st>>>(d_out, n_out_cap, n_out_dev, c_out, d_bias);
    BTC_CHECK_LAUNCH("conv_bwd_weight");
    return BTC_OK;
}

}  // extern "C"
