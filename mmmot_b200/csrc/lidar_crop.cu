// Per-detection LiDAR cropping on the GPU (the step immediately before the hot path).
// Replaces the per-box host loop of reference point_cloud/preprocess.py:72-81 (remove_points_outside_boxes
// -> box_np_ops.points_in_rbbox -> geometry._points_in_convex_polygon_3d_jit, geometry.py:96-114).
//
// The six inward-facing plane equations of every rotated box are prepared on the host exactly as the
// reference's numpy code computes them (mmmot_b200/lidar_crop.py); the kernels evaluate the reference's
// membership predicate  sign = x*nx + y*ny + z*nz + d ; inside <=> sign < 0 for all 6 planes  with the same
// operation order and NO fused multiply-add, in the precision of the plane equations: float64 in the reference's
// real pipeline (box_camera_to_lidar promotes the boxes to float64, box_np_ops.py:584-589, so numba evaluates the
// predicate in float64 on the float32 points), float32 when the caller hands float32 boxes.  Membership is
// bit-identical to the reference in both cases.
// Output = the packed per-detection point list + CSR offsets that mmmot_pointnet_fwd consumes; point order
// inside a detection is the scene order (stable compaction); an empty box yields one all-zero point
// (preprocess.py:78-79).
#include "common.cuh"

namespace {

constexpr int kTile = 256;   // points per CTA

__device__ __forceinline__ bool inside_box(const float* __restrict__ pl, float x, float y, float z) {
  // pl: 6 planes x (nx, ny, nz, d)
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float s = __fadd_rn(__fmul_rn(x, pl[4 * k]), __fmul_rn(y, pl[4 * k + 1]));
    s = __fadd_rn(s, __fmul_rn(z, pl[4 * k + 2]));
    s = __fadd_rn(s, pl[4 * k + 3]);
    if (s >= 0.f) return false;
  }
  return true;
}
__device__ __forceinline__ bool inside_box(const double* __restrict__ pl, float xf, float yf, float zf) {
  const double x = xf, y = yf, z = zf;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double s = __dadd_rn(__dmul_rn(x, pl[4 * k]), __dmul_rn(y, pl[4 * k + 1]));
    s = __dadd_rn(s, __dmul_rn(z, pl[4 * k + 2]));
    s = __dadd_rn(s, pl[4 * k + 3]);
    if (s >= 0.0) return false;
  }
  return true;
}

// counts[b][tile] = number of points of the tile inside box b
template <typename T>
__global__ void __launch_bounds__(kTile) crop_count_kernel(const float* __restrict__ pts, int stride, int P,
                                                           const T* __restrict__ planes, int tiles,
                                                           int* __restrict__ counts) {
  __shared__ T pl[24];
  const int b = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < 24) pl[threadIdx.x] = planes[b * 24 + threadIdx.x];
  __syncthreads();
  const int p = tile * kTile + threadIdx.x;
  bool in = false;
  if (p < P) in = inside_box(pl, pts[(long)p * stride], pts[(long)p * stride + 1], pts[(long)p * stride + 2]);
  const int c = __syncthreads_count(in);
  if (threadIdx.x == 0) counts[b * tiles + tile] = c;
}

// per box: exclusive scan of its tile counts (in place) + total; one CTA per box
__global__ void __launch_bounds__(256) crop_scan_tiles_kernel(int* __restrict__ counts, int tiles,
                                                              int* __restrict__ totals) {
  __shared__ int carry;
  __shared__ int wsum[8];
  const int b = blockIdx.x;
  int* row = counts + (long)b * tiles;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < tiles; t0 += 256) {
    const int t = t0 + threadIdx.x;
    const int v = t < tiles ? row[t] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); w++) woff += wsum[w];
    const int incl = x + woff + carry;
    if (t < tiles) row[t] = incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[b] = carry;
}

// split[b+1] = split[b] + max(total_b, 1)  (an empty box keeps one zero point); single thread, n is small
__global__ void crop_scan_boxes_kernel(const int* __restrict__ totals, int n, int* __restrict__ split) {
  if (threadIdx.x || blockIdx.x) return;
  int acc = 0;
  split[0] = 0;
  for (int b = 0; b < n; b++) {
    acc += totals[b] > 0 ? totals[b] : 1;
    split[b + 1] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kTile) crop_scatter_kernel(const float* __restrict__ pts, int stride, int P,
                                                             const T* __restrict__ planes, int tiles,
                                                             const int* __restrict__ tile_off,
                                                             const int* __restrict__ split, int out_c,
                                                             float* __restrict__ out) {
  __shared__ T pl[24];
  __shared__ int wcnt[kTile / 32];
  const int b = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < 24) pl[threadIdx.x] = planes[b * 24 + threadIdx.x];
  __syncthreads();
  const int p = tile * kTile + threadIdx.x;
  bool in = false;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (p < P) {
    for (int c = 0; c < stride && c < 4; c++) v[c] = pts[(long)p * stride + c];
    in = inside_box(pl, v[0], v[1], v[2]);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, in);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wcnt[warp] = __popc(bal);
  __syncthreads();
  int off = 0;
  for (int w = 0; w < warp; w++) off += wcnt[w];
  if (in) {
    const long dst = (long)split[b] + tile_off[b * tiles + tile] + off + __popc(bal & ((1u << lane) - 1));
    for (int c = 0; c < out_c; c++) out[dst * out_c + c] = v[c];
  }
}

// boxes with no point inside: one all-zero point
__global__ void crop_fill_empty_kernel(const int* __restrict__ totals, const int* __restrict__ split, int n,
                                       int out_c, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n || totals[b] > 0) return;
  for (int c = 0; c < out_c; c++) out[(long)split[b] * out_c + c] = 0.f;
}

}  // namespace

extern "C" size_t mmmot_crop_workspace(int n_points, int n_boxes) {
  const size_t tiles = (size_t)mm_cdiv(n_points, kTile);
  return mm_align(tiles * n_boxes * sizeof(int)) + mm_align((size_t)n_boxes * sizeof(int));
}

extern "C" int mmmot_crop_count(const float* points, int n_points, int stride, const void* planes, int planes_f64,
                                int n_boxes, int* split, void* workspace, size_t workspace_bytes, void* stream) {
  if (!points || !planes || !split || !workspace || n_points <= 0 || n_boxes <= 0 || stride < 3) return MMMOT_E_ARG;
  if (workspace_bytes < mmmot_crop_workspace(n_points, n_boxes)) return MMMOT_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = mm_cdiv(n_points, kTile);
  int* counts = (int*)workspace;
  int* totals = (int*)((char*)workspace + mm_align((size_t)tiles * n_boxes * sizeof(int)));
  if (planes_f64)
    crop_count_kernel<double><<<dim3(tiles, n_boxes), kTile, 0, st>>>(points, stride, n_points, (const double*)planes, tiles, counts);
  else
    crop_count_kernel<float><<<dim3(tiles, n_boxes), kTile, 0, st>>>(points, stride, n_points, (const float*)planes, tiles, counts);
  MM_LAUNCH_CHECK();
  crop_scan_tiles_kernel<<<n_boxes, 256, 0, st>>>(counts, tiles, totals);
  MM_LAUNCH_CHECK();
  crop_scan_boxes_kernel<<<1, 32, 0, st>>>(totals, n_boxes, split);
  MM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mmmot_crop_scatter(const float* points, int n_points, int stride, const void* planes, int planes_f64,
                                  int n_boxes, const int* split, int out_channels, float* out_points, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!points || !planes || !split || !out_points || !workspace || n_points <= 0 || n_boxes <= 0) return MMMOT_E_ARG;
  if (out_channels < 3 || out_channels > 4 || out_channels > stride) return MMMOT_E_ARG;
  if (workspace_bytes < mmmot_crop_workspace(n_points, n_boxes)) return MMMOT_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = mm_cdiv(n_points, kTile);
  const int* tile_off = (const int*)workspace;
  const int* totals = (const int*)((const char*)workspace + mm_align((size_t)tiles * n_boxes * sizeof(int)));
  if (planes_f64)
    crop_scatter_kernel<double><<<dim3(tiles, n_boxes), kTile, 0, st>>>(points, stride, n_points, (const double*)planes, tiles,
                                                                        tile_off, split, out_channels, out_points);
  else
    crop_scatter_kernel<float><<<dim3(tiles, n_boxes), kTile, 0, st>>>(points, stride, n_points, (const float*)planes, tiles,
                                                                       tile_off, split, out_channels, out_points);
  MM_LAUNCH_CHECK();
  crop_fill_empty_kernel<<<mm_cdiv(n_boxes, 128), 128, 0, st>>>(totals, split, n_boxes, out_channels, out_points);
  MM_LAUNCH_CHECK();
  return 0;
}
