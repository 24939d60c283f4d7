// Exact solver for the 2-frame association integer programme.
// Replaces reference solvers.py:9-138 (ortools_solve: CBC MIP built variable by variable in Python).
//
// For two frames the programme (variables solvers.py:17-30, objective :31-49, flow constraints
// :83-111) is totally unimodular and equals a rectangular assignment problem:
//   prev det j  : inactive (0) | active, track ends  u_j = det_j + new_j + end_j | linked to k
//   next det k  : inactive (0) | active, track starts v_k = det_k + end_k + new_k | linked from j
//   link (j,k)  : c_jk = (det_j + new_j) + (det_k + end_k) + link_jk
//   maximise  sum_j U_j + sum_k V_k + sum_{linked} (c_jk - U_j - V_k),  U = max(u,0), V = max(v,0)
// i.e. max-weight bipartite matching with weights w_jk = c_jk - U_j - V_k and a zero-weight private
// "stay unmatched" column per row.  Solved with the shortest-augmenting-path Hungarian method in
// fp64 — one WARP per frame-pair, columns strided over the 32 lanes, arg-min by warp shuffles, no
// block barrier.  Deterministic tie rule: smallest column index wins; a detection whose best
// unmatched value is exactly 0 stays inactive.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;

__host__ __device__ inline size_t lp_warp_bytes(int n, int m) {
  size_t nc = (size_t)n + m + 1;
  size_t d = (nc * 2 + (n + 1) + n + m) * sizeof(double);  // v, minv, u, A, B
  size_t i = nc * 2 * sizeof(int);                          // p, way
  size_t b = (nc + 7) / 8 * 8;                              // used
  return (d + i + b + 15) / 16 * 16;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) lp_assign_kernel(
    const float* __restrict__ det, long det_stride, const float* __restrict__ link, long link_stride,
    const float* __restrict__ new_s, long new_stride, const float* __restrict__ end_s, long end_stride,
    int pairs, int n, int m, float* __restrict__ a_det, float* __restrict__ a_link,
    float* __restrict__ a_new, float* __restrict__ a_end, int* __restrict__ match) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = blockIdx.x * kWarpsPerCta + warp;
  if (pair >= pairs) return;
  const int nc = n + m;  // columns 1..m real, m+1..m+n private dummies (column 0 = sentinel)
  unsigned char* base = smem + (size_t)warp * lp_warp_bytes(n, m);
  double* v = (double*)base;
  double* minv = v + nc + 1;
  double* u = minv + nc + 1;
  double* A = u + n + 1;
  double* B = A + n;
  int* p = (int*)(B + m);
  int* way = p + nc + 1;
  unsigned char* used = (unsigned char*)(way + nc + 1);

  const float* d = det + (long)pair * det_stride;
  const float* ns = new_s + (long)pair * new_stride;
  const float* es = end_s + (long)pair * end_stride;
  const float* lk = link + (long)pair * link_stride;
  const double INF = INFINITY;

  // A_j = a_j - U_j, B_k = b_k - V_k; matched weight w_jk = A_j + B_k + link_jk
  for (int j = lane; j < n; j += 32) {
    double a = (double)d[j] + (double)ns[j];
    double uu = a + (double)es[j];
    A[j] = a - fmax(uu, 0.0);
    u[j + 1] = 0.0;
  }
  for (int k = lane; k < m; k += 32) {
    double b = (double)d[n + k] + (double)es[n + k];
    double vv = b + (double)ns[n + k];
    B[k] = b - fmax(vv, 0.0);
  }
  for (int j = lane; j <= nc; j += 32) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
  if (lane == 0) u[0] = 0.0;
  __syncwarp();

  for (int i = 1; i <= n; i++) {
    if (lane == 0) p[0] = i;
    for (int j = lane; j <= nc; j += 32) { minv[j] = INF; used[j] = 0; }
    __syncwarp();
    int j0 = 0;
    // every pass marks one more column used, so nc + 1 passes bound the search (guard against hangs)
    for (int pass = 0; pass <= nc; pass++) {
      if (lane == 0) used[j0] = 1;
      __syncwarp();
      const int i0 = p[j0];
      const double ui = u[i0], ai = A[i0 - 1];
      const float* lrow = lk + (long)(i0 - 1) * m;
      double best = INF;
      int bj = 0x7fffffff;
      for (int j = 1 + lane; j <= nc; j += 32) {
        if (used[j]) continue;
        double cost;
        if (j <= m) cost = -(ai + B[j - 1] + (double)lrow[j - 1]);
        else cost = (j - m == i0) ? 0.0 : INF;
        double cur = cost - ui - v[j];
        double mv = minv[j];
        if (cur < mv) { mv = cur; minv[j] = cur; way[j] = j0; }
        if (mv < best) { best = mv; bj = j; }   // ascending j per lane: first minimum kept
      }
      // the scan above wrote minv[j] / way[j] from lane (j-1)%32; the update below touches minv[j] from lane j%32
      __syncwarp();
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
      }
      const double delta = best;
      for (int j = lane; j <= nc; j += 32) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      __syncwarp();
      j0 = bj;
      if (p[j0] == 0) break;
    }
    // every lane has read p[j0] before lane 0 rewrites p[] below (racecheck: read at the loop exit vs the augmenting write)
    __syncwarp();
    // augment along the alternating path (serial, short)
    if (lane == 0) {
      int j = j0;
      while (j) { int j1 = way[j]; p[j] = p[j1]; j = j1; }
    }
    __syncwarp();
  }

  // ---- write the 0/1 solution in ortools_solve's layout (solvers.py:115-138) ----
  const int L = n + m;
  float* od = a_det + (long)pair * L;
  float* on = a_new + (long)pair * L;
  float* oe = a_end + (long)pair * L;
  float* ol = a_link + (long)pair * n * m;
  int* om = match + (long)pair * n;
  // next-frame detections: default unmatched
  for (int k = lane; k < m; k += 32) {
    double vv = (double)d[n + k] + (double)es[n + k] + (double)ns[n + k];
    float act = vv > 0.0 ? 1.f : 0.f;
    od[n + k] = act; on[n + k] = act; oe[n + k] = act;
  }
  for (int j = lane; j < n; j += 32) {
    double uu = (double)d[j] + (double)ns[j] + (double)es[j];
    float act = uu > 0.0 ? 1.f : 0.f;
    od[j] = act; on[j] = act; oe[j] = act;
    om[j] = -1;
  }
  __syncwarp();
  for (int k = 1 + lane; k <= m; k += 32) {
    int r = p[k];
    if (r > 0) {  // prev det r-1 linked to next det k-1
      int j = r - 1;
      od[j] = 1.f; on[j] = 1.f; oe[j] = 0.f; om[j] = k - 1;
      od[n + k - 1] = 1.f; on[n + k - 1] = 0.f; oe[n + k - 1] = 1.f;
      ol[(long)j * m + (k - 1)] = 1.f;
    }
  }
}

}  // namespace

extern "C" size_t mmmot_lp_workspace(int pairs, int n, int m) {
  (void)pairs; (void)n; (void)m;
  return 256;  // all solver state lives in shared memory; kept non-zero so callers can always pass a buffer
}

extern "C" int mmmot_lp_assign(const float* det, long det_stride, const float* link, long link_stride,
                               const float* new_s, long new_stride, const float* end_s, long end_stride,
                               int pairs, int n, int m, float* a_det, float* a_link, float* a_new,
                               float* a_end, int* match, void* workspace, size_t workspace_bytes,
                               void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!det || !link || !new_s || !end_s || !a_det || !a_link || !a_new || !a_end || !match) return MMMOT_E_ARG;
  if (pairs <= 0 || n <= 0 || m <= 0) return MMMOT_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = lp_warp_bytes(n, m) * kWarpsPerCta;
  if (smem > 227 * 1024) return MMMOT_E_SHAPE;
  MM_CUDA(cudaFuncSetAttribute(lp_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // size varies: set per call
  MM_CUDA(cudaMemsetAsync(a_link, 0, (size_t)pairs * n * m * sizeof(float), st));
  const bool timed = mm_timing_on();
  if (timed) mm_timing_begin(st, MM_T_LP, 0.0, 4.0 * pairs * ((double)n * m * 2 + 7.0 * (n + m)));
  lp_assign_kernel<<<mm_cdiv(pairs, kWarpsPerCta), kWarpsPerCta * 32, smem, st>>>(
      det, det_stride, link, link_stride, new_s, new_stride, end_s, end_stride, pairs, n, m, a_det, a_link,
      a_new, a_end, match);
  MM_LAUNCH_CHECK();
  if (timed) mm_timing_end(st);
  return 0;
}
