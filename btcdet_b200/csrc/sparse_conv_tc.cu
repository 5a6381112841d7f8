// tcgen05 (5th-gen tensor core) gather-GEMM tile for wide sparse-conv layers.
// Placeholder until the UMMA tile lands: reports "unsupported" so the dispatcher uses the
// fp32 FFMA tile.
#include "common.cuh"

namespace btc {

int conv_fwd_tc(const float*, const int*, int, const float*, const float*, const float*, const float*, int, float*, int,
                const int*, int, int, int, cudaStream_t) {
    return BTC_E_UNSUPPORTED;
}

}  // namespace btc
