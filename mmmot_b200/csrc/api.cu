// Library-wide entry points of libmmmot_sm100a.so (see include/mmmot_b200.h).
#include <atomic>

#include "common.cuh"

static std::atomic<unsigned long long> g_launches{0};

void mm_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" unsigned long long mmmot_launch_count(void) { return g_launches.load(); }

extern "C" int mmmot_abi_version(void) { return MMMOT_ABI_VERSION; }

extern "C" int mmmot_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MM_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Optional per-launch timing of the dominant kernel (the 3x3-conv contraction), used by bench.py
// for the roofline figure: CUDA events recorded on the launching stream around every launch
// while enabled; mmmot_timing_collect() synchronises on the events and returns the totals.
#include <mutex>
#include <vector>

namespace {
std::mutex g_tmu;
bool g_timing = false;
struct Span { cudaEvent_t a, b; double flop; };
std::vector<Span> g_spans;
}  // namespace

bool mm_timing_on() { return g_timing; }

void mm_timing_begin(cudaStream_t st, double flop) {
  std::lock_guard<std::mutex> l(g_tmu);
  Span s;
  cudaEventCreate(&s.a);
  cudaEventCreate(&s.b);
  s.flop = flop;
  cudaEventRecord(s.a, st);
  g_spans.push_back(s);
}

void mm_timing_end(cudaStream_t st) {
  std::lock_guard<std::mutex> l(g_tmu);
  if (!g_spans.empty()) cudaEventRecord(g_spans.back().b, st);
}

extern "C" int mmmot_timing_enable(int on) {
  std::lock_guard<std::mutex> l(g_tmu);
  g_timing = on != 0;
  return 0;
}

extern "C" int mmmot_timing_collect(double* total_ms, double* total_flop, long* launches) {
  std::lock_guard<std::mutex> l(g_tmu);
  double ms = 0.0, fl = 0.0;
  long n = 0;
  for (auto& s : g_spans) {
    float t = 0.f;
    cudaError_t e = cudaEventSynchronize(s.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, s.a, s.b);
    if (e != cudaSuccess) return (int)e;
    ms += t; fl += s.flop; n++;
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  g_spans.clear();
  if (total_ms) *total_ms = ms;
  if (total_flop) *total_flop = fl;
  if (launches) *launches = n;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Engine selection + single-contraction test hook.
#include "gemm_tc.cuh"

namespace { int g_engine = 0; int g_dbg = 0; int g_kseg = 36; }
int mm_kseg_chunks() { return g_kseg; }
extern "C" int mmmot_set_kseg(int chunks) { if (chunks < 0) return MMMOT_E_ARG; g_kseg = chunks; return 0; }
int mm_debug_flags() { return g_dbg; }
extern "C" int mmmot_set_debug(int flags) { g_dbg = flags; return 0; }

int mm_engine() { return g_engine; }

extern "C" int mmmot_set_engine(int engine) {
  if (engine < 0 || engine > 2) return MMMOT_E_ARG;
  g_engine = engine;
  return 0;
}

extern "C" int mmmot_debug_linear(const float* Wt, const void* Wp, float wp_scale, const float* bias, const float* X,
                                  float* Y, int M, int K, int S, int engine, void* stream) {
  if (!Wt || !X || !Y || M <= 0 || K <= 0 || S <= 0) return MMMOT_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  GemmP p = gemm_defaults();
  p.Wt = Wt; p.ldw = M; p.bias = bias; p.M = M; p.K = K;
  p.S = S;
  p.X = X; p.x_ks = S;
  p.Y = Y; p.y_ms = S;
  if (engine == 2) {
    p.tiles_per_group = mm_cdiv(S, tc::BN); p.num_tiles = p.tiles_per_group;
    return gemm_tc_launch<XM_DIRECT>(p, (const uint4*)Wp, wp_scale, st);
  }
  p.tiles_per_group = mm_cdiv(S, 128); p.num_tiles = p.tiles_per_group;
  return gemm_simt_launch<XM_DIRECT>(p, st);
}

// ---------------------------------------------------------------------------------------------
// Test hooks for the TMA-fed tcgen05 engine (planar FP16 hi/lo channels-last operands).
#include "gemm_tma.cuh"

// Y[rows][M] fp32 (channels-last) = X W^T + bias ; X given as planes Xhi[rows][K], Xlo = Xhi + rows*K
extern "C" int mmmot_debug_linear_planar(const void* Wp, float wp_scale, const float* bias, const void* Xhi,
                                         float* Y, int M, int K, long rows, void* stream) {
  if (!Wp || !Xhi || !Y || M <= 0 || K <= 0 || rows <= 0) return MMMOT_E_ARG;
  GemmP p = gemm_defaults();
  p.bias = bias; p.M = M; p.K = K;
  p.S = (int)rows; p.tiles_per_group = mm_cdiv(rows, tc::BN); p.num_tiles = p.tiles_per_group;
  p.Y = Y; p.y_ms = M;
  return gemm_tma_launch_mat(p, (const uint4*)Wp, wp_scale, (const __half*)Xhi, rows * (long)K, rows, K, tc::OUT_CL, 0,
                             (cudaStream_t)stream);
}

// 3x3 conv + bias + ReLU on planar FP16 NHWC: X planes [2][n][H][W][C] -> Y planes [2][n][H][W][M]
extern "C" int mmmot_debug_conv_planar(const void* Wp, float wp_scale, const float* bias, const void* Xhi, void* Yhi,
                                       int n_img, int H, int W, int C, int M, float* kseg_scratch, void* stream) {
  if (!Wp || !Xhi || !Yhi) return MMMOT_E_ARG;
  GemmP p = gemm_defaults();
  p.bias = bias; p.M = M; p.relu = 1;
  return gemm_tma_launch_conv(p, (const uint4*)Wp, wp_scale, (const __half*)Xhi, (long)n_img * H * W * C, n_img, H, W, C,
                              (__half*)Yhi, (long)n_img * H * W * M, (cudaStream_t)stream, kseg_scratch);
}
