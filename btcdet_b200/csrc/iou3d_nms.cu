// Rotated bird's-eye-view overlap / IoU and rotated NMS.
//
// Replaces the reference's btcdet/ops/iou3d_nms extension (iou3d_nms_kernel.cu:236-413 kernels, iou3d_nms.cpp:103-140 host
// side: the N x N/64 suppression mask is copied to the host and scanned there).  Same results, different construction:
//   * overlap of two rotated rectangles: rectangle A is moved into B's frame (B becomes an axis-aligned box centred at the
//     origin), its quadrilateral is clipped against B's four half-planes (Sutherland-Hodgman, at most 8 vertices)
//     and the area comes from the shoelace sum — no edge-pair intersection tests, no angular vertex sort;
//   * NMS: 64 x 64 blocks of the suppression matrix as bit masks (one thread per row of a block, the column block's
//     boxes staged in shared memory), then the greedy scan ON THE DEVICE by one warp (the removed-set lives in shared
//     memory, lanes OR the kept row's mask words in parallel) — the keep list and its length never need the N^2/8-byte
//     device->host copy, and the call is stream-ordered / graph-capturable.
// Boxes are (x, y, z, dx, dy, dz, heading) rows of 7 floats, heading = rotation about z, as in the reference.
#include "common.cuh"

namespace btc {

struct Rect {
    float cx, cy, hx, hy, c, s;      // centre, half extents, cos / sin of the heading
};

__device__ __forceinline__ Rect load_rect(const float* b) {
    Rect r;
    r.cx = b[0]; r.cy = b[1]; r.hx = 0.5f * b[3]; r.hy = 0.5f * b[4];
    sincosf(b[6], &r.s, &r.c);
    return r;
}

// clip a convex polygon (n <= 8 vertices) against the half-plane  SIGN * coord(AXIS) <= lim  (Sutherland-Hodgman step:
// keep inside vertices, add the crossing point of every edge that changes side)
template <int AXIS, int SIGN>
__device__ __forceinline__ int clip_halfplane(float (&px)[8], float (&py)[8], int n, float lim) {
    float qx[8], qy[8];
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const float ax = px[i], ay = py[i], bx = px[j], by = py[j];
        const float da = (AXIS == 0 ? ax : ay) * SIGN - lim;     // <= 0: inside
        const float db = (AXIS == 0 ? bx : by) * SIGN - lim;
        const bool ina = da <= 0.f, inb = db <= 0.f;
        if (ina && m < 8) { qx[m] = ax; qy[m] = ay; ++m; }
        if (ina != inb && m < 8) {
            const float t = da / (da - db);
            qx[m] = ax + t * (bx - ax);
            qy[m] = ay + t * (by - ay);
            ++m;
        }
    }
    for (int i = 0; i < m; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
    return m;
}

// area of the intersection of two rotated rectangles
__device__ float rect_overlap(const Rect& a, const Rect& b) {
    // A's centre and axes in B's frame
    const float dx = a.cx - b.cx, dy = a.cy - b.cy;
    const float ox = dx * b.c + dy * b.s, oy = -dx * b.s + dy * b.c;
    const float c = a.c * b.c + a.s * b.s, s = a.s * b.c - a.c * b.s;       // cos / sin of (heading_a - heading_b)
    // quick reject on the bounding circles
    const float ra = sqrtf(a.hx * a.hx + a.hy * a.hy), rb = sqrtf(b.hx * b.hx + b.hy * b.hy);
    if (ox * ox + oy * oy > (ra + rb) * (ra + rb)) return 0.f;
    const float ux = a.hx * c, uy = a.hx * s, vx = -a.hy * s, vy = a.hy * c;
    float px[8], py[8];
    px[0] = ox - ux - vx; py[0] = oy - uy - vy;
    px[1] = ox + ux - vx; py[1] = oy + uy - vy;
    px[2] = ox + ux + vx; py[2] = oy + uy + vy;
    px[3] = ox - ux + vx; py[3] = oy - uy + vy;
    int n = 4;
    n = clip_halfplane<0, 1>(px, py, n, b.hx);
    if (n < 3) return 0.f;
    n = clip_halfplane<0, -1>(px, py, n, b.hx);
    if (n < 3) return 0.f;
    n = clip_halfplane<1, 1>(px, py, n, b.hy);
    if (n < 3) return 0.f;
    n = clip_halfplane<1, -1>(px, py, n, b.hy);
    if (n < 3) return 0.f;
    float twice = 0.f;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        twice += px[i] * py[j] - px[j] * py[i];
    }
    return 0.5f * fabsf(twice);
}

__device__ __forceinline__ float iou_rotated(const float* pa, const float* pb) {
    const Rect a = load_rect(pa), b = load_rect(pb);
    const float sa = pa[3] * pa[4], sb = pb[3] * pb[4];
    const float ov = rect_overlap(a, b);
    return ov / fmaxf(sa + sb - ov, 1e-8f);
}

// axis-aligned variant of nms_normal (iou3d_nms_kernel.cu:224-234): headings ignored
__device__ __forceinline__ float iou_axis_aligned(const float* pa, const float* pb) {
    const float l = fmaxf(pa[0] - 0.5f * pa[3], pb[0] - 0.5f * pb[3]), r = fminf(pa[0] + 0.5f * pa[3], pb[0] + 0.5f * pb[3]);
    const float t = fmaxf(pa[1] - 0.5f * pa[4], pb[1] - 0.5f * pb[4]), d = fminf(pa[1] + 0.5f * pa[4], pb[1] + 0.5f * pb[4]);
    const float ov = fmaxf(r - l, 0.f) * fmaxf(d - t, 0.f);
    return ov / fmaxf(pa[3] * pa[4] + pb[3] * pb[4] - ov, 1e-8f);
}

__global__ void boxes_bev_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, int mode,
                                 float* __restrict__ out) {
    const int64_t total = (int64_t)n * m;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const float* pa = a + (t / m) * 7;
        const float* pb = b + (t % m) * 7;
        out[t] = mode == 1 ? rect_overlap(load_rect(pa), load_rect(pb)) : iou_rotated(pa, pb);
    }
}

// suppression matrix in 64 x 64 blocks: bit j of mask[i][cb] = IoU(box i, box cb * 64 + j) > thresh, only for j > i
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, int n, float thresh, int normal,
                                                      unsigned long long* __restrict__ mask, int nblk) {
    __shared__ float s_box[64 * 7];
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;                                    // below the diagonal: never read by the scan
    const int ncol = min(n - cb * 64, 64);
    for (int t = threadIdx.x; t < ncol * 7; t += 64) s_box[t] = boxes[(int64_t)cb * 64 * 7 + t];
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    float mine[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) mine[t] = boxes[(int64_t)i * 7 + t];
    unsigned long long bits = 0;
    const int j0 = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = j0; j < ncol; ++j) {
        const float v = normal ? iou_axis_aligned(mine, s_box + j * 7) : iou_rotated(mine, s_box + j * 7);
        if (v > thresh) bits |= 1ull << j;
    }
    mask[(int64_t)i * nblk + cb] = bits;
}

// greedy scan by one warp: keep box i unless an earlier kept box suppressed it
__global__ void __launch_bounds__(32) nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int nblk,
                                                      long long* __restrict__ keep, int* __restrict__ num_out) {
    extern __shared__ unsigned long long s_removed[];       // [nblk]
    const int lane = threadIdx.x;
    for (int w = lane; w < nblk; w += 32) s_removed[w] = 0ull;
    __syncwarp();
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const unsigned long long word = s_removed[i >> 6];
        if (!((word >> (i & 63)) & 1ull)) {                 // warp-uniform
            if (lane == 0) keep[cnt] = i;
            ++cnt;
            for (int w = (i >> 6) + lane; w < nblk; w += 32) s_removed[w] |= mask[(int64_t)i * nblk + w];
            __syncwarp();
        }
    }
    if (lane == 0) *num_out = cnt;
}

}  // namespace btc

using namespace btc;

extern "C" {

int btc_boxes_bev(const float* boxes_a, int n, const float* boxes_b, int m, int mode, float* out, void* stream) {
    if (n < 0 || m < 0 || (mode != 0 && mode != 1)) return badarg("btc_boxes_bev: bad sizes / mode");
    if (n == 0 || m == 0) return BTC_OK;
    if (!boxes_a || !boxes_b || !out) return badarg("btc_boxes_bev: null argument");
    boxes_bev_kernel<<<grid_for((int64_t)n * m, 128), 128, 0, (cudaStream_t)stream>>>(boxes_a, n, boxes_b, m, mode, out);
    BTC_CHECK_LAUNCH("boxes_bev");
    return BTC_OK;
}

int64_t btc_nms_workspace_bytes(int n) {
    if (n < 0) return BTC_E_BADARG;
    const int64_t nblk = (n + 63) / 64;
    return align_up((int64_t)(n > 0 ? n : 1) * nblk * 8, 256);
}

int btc_nms(const float* boxes, int n, float thresh, int normal, long long* keep, int* num_out, void* workspace,
            int64_t workspace_bytes, void* stream) {
    if (n < 0 || !num_out) return badarg("btc_nms: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        BTC_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int), st), "nms memset");
        return BTC_OK;
    }
    if (!boxes || !keep || !workspace) return badarg("btc_nms: null argument");
    if (workspace_bytes < btc_nms_workspace_bytes(n)) return badarg("btc_nms: workspace too small");
    const int nblk = (n + 63) / 64;
    if ((size_t)nblk * 8 > 48 * 1024) return set_error(BTC_E_UNSUPPORTED, "btc_nms: more than 393216 boxes", cudaSuccess);
    unsigned long long* mask = (unsigned long long*)workspace;
    BTC_CUDA(cudaMemsetAsync(mask, 0, (size_t)n * nblk * 8, st), "nms memset mask");
    nms_mask_kernel<<<dim3(nblk, nblk), 64, 0, st>>>(boxes, n, thresh, normal ? 1 : 0, mask, nblk);
    nms_scan_kernel<<<1, 32, (size_t)nblk * 8, st>>>(mask, n, nblk, keep, num_out);
    BTC_CHECK_LAUNCH("nms");
    return BTC_OK;
}

}  // extern "C"
