// Per-detection image crop-and-resize on the GPU (the image-side step before the hot path).
// Replaces, per detection, reference dataset/test_seq_dataset.py:212-218
//     img.crop((x1, y1, x2, y2)).resize((224, 224), Image.BILINEAR)
// followed by utils/build_util.py:137-142 (ToTensor, Normalize; Resize/CenterCrop are identities at 224), and writes
// the fp32 [n][3][S][S] tensor TrackingNet.forward takes as `dets`.
//
// The resize is Pillow's 8-bit two-pass resampler (triangle filter widened by the down-scale factor); it is
// reproduced bit-exactly: coefficients in IEEE double with the same operation order and no fused multiply-add,
// 22-bit fixed point, integer accumulation, clip to 8 bits after the horizontal and after the vertical pass.
// Crop pixels outside the image are 0 (PIL pads with black).  Normalisation uses IEEE fp32 division like torch.
#include "common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kMaxTaps = 1024;   // upper bound on the caller's tap stride (ceil(scale)*2+1, scale = crop side / output side)

struct NormP { float mean[3], stdv[3]; };

// one thread per (detection, axis, output index): tap range + fixed-point weights
__global__ void resize_coeff_kernel(const int* __restrict__ boxes, int n_det, int S, int taps,
                                    int2* __restrict__ bounds, int* __restrict__ kk) {
  const int d = blockIdx.y >> 1, axis = blockIdx.y & 1;
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= S) return;
  const int in_size = axis ? boxes[4 * d + 3] - boxes[4 * d + 1] : boxes[4 * d + 2] - boxes[4 * d];
  const double scale = (double)((float)in_size - 0.0f) / (double)S;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;             // triangle filter support 1.0
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, scale));
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > taps) xmax = taps;                   // host sizes `taps` from the boxes; never taken
  auto weight = [&](int x) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (a < 0.0) a = -a;
    return a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
  };
  double ww = 0.0;
  for (int x = 0; x < xmax; x++) ww = __dadd_rn(ww, weight(x));
  int* k = kk + ((long)(d * 2 + axis) * S + xx) * taps;
  for (int x = 0; x < xmax; x++) {
    const double w = weight(x);
    const double v = ww != 0.0 ? __ddiv_rn(w, ww) : w;
    k[x] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(v, (double)(1 << kPrecisionBits))));
  }
  bounds[(long)(d * 2 + axis) * S + xx] = make_int2(xmin, xmax);
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: tmp[row_off[d] + y][xx][c], y over the crop's rows
__global__ void resize_h_kernel(const unsigned char* __restrict__ img, int H, int W, const int* __restrict__ boxes,
                                const long long* __restrict__ row_off, int S, int taps,
                                const int2* __restrict__ bounds, const int* __restrict__ kk,
                                unsigned char* __restrict__ tmp) {
  const int d = blockIdx.y;
  const int x1 = boxes[4 * d], y1 = boxes[4 * d + 1], y2 = boxes[4 * d + 3];
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ch = y2 - y1;
  if (idx >= (long)ch * S) return;
  const int y = (int)(idx / S), xx = (int)(idx - (long)y * S);
  const int2 b = bounds[(long)(d * 2) * S + xx];
  const int* k = kk + ((long)(d * 2) * S + xx) * taps;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const int iy = y1 + y;
  if (iy >= 0 && iy < H) {
    const unsigned char* rowp = img + (long)iy * W * 3;
    for (int x = 0; x < b.y; x++) {
      const int ix = x1 + b.x + x;
      if (ix >= 0 && ix < W) {
        const int c = k[x];
        s0 += rowp[ix * 3] * c; s1 += rowp[ix * 3 + 1] * c; s2 += rowp[ix * 3 + 2] * c;
      }
    }
  }
  unsigned char* o = tmp + ((row_off[d] + y) * S + xx) * 3;
  o[0] = (unsigned char)clip8(s0); o[1] = (unsigned char)clip8(s1); o[2] = (unsigned char)clip8(s2);
}

// vertical pass + ToTensor + Normalize: out[d][c][yy][xx]
__global__ void resize_v_kernel(const unsigned char* __restrict__ tmp, const long long* __restrict__ row_off, int S,
                                int taps, const int2* __restrict__ bounds, const int* __restrict__ kk, NormP nm,
                                float* __restrict__ out) {
  const int d = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int yy = idx / S, xx = idx - yy * S;
  const int2 b = bounds[(long)(d * 2 + 1) * S + yy];
  const int* k = kk + ((long)(d * 2 + 1) * S + yy) * taps;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const unsigned char* src = tmp + ((row_off[d] + b.x) * S + xx) * 3;
  for (int y = 0; y < b.y; y++) {
    const int c = k[y];
    s0 += src[0] * c; s1 += src[1] * c; s2 += src[2] * c;
    src += (long)S * 3;
  }
  const int v[3] = {clip8(s0), clip8(s1), clip8(s2)};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float t = __fdiv_rn((float)v[c], 255.f);
    out[(((long)d * 3 + c) * S + yy) * S + xx] = __fdiv_rn(__fsub_rn(t, nm.mean[c]), nm.stdv[c]);
  }
}

struct CrWs { int2* bounds; int* kk; unsigned char* tmp; };
CrWs carve_cr(MmArena& a, int n_det, long total_rows, int S, int taps) {
  CrWs w;
  w.bounds = a.take<int2>((size_t)n_det * 2 * S);
  w.kk = a.take<int>((size_t)n_det * 2 * S * taps);
  w.tmp = a.take<unsigned char>((size_t)total_rows * S * 3);
  return w;
}

}  // namespace

extern "C" int mmmot_crop_resize_max_taps(void) { return kMaxTaps; }

extern "C" size_t mmmot_crop_resize_workspace(int n_det, long total_rows, int out_size, int taps) {
  MmArena a(nullptr, 0);
  carve_cr(a, n_det, total_rows, out_size, taps);
  return a.off;
}

extern "C" int mmmot_crop_resize(const unsigned char* image, int img_h, int img_w, const int* boxes,
                                 const long long* row_off, int n_det, long total_rows, int max_crop_h, int out_size,
                                 int taps, const float* mean_std, float* out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (!image || !boxes || !row_off || !mean_std || !out || !workspace) return MMMOT_E_ARG;
  if (n_det <= 0 || img_h <= 0 || img_w <= 0 || out_size <= 0 || max_crop_h <= 0 || total_rows <= 0) return MMMOT_E_SHAPE;
  if (taps < 3 || taps > kMaxTaps) return MMMOT_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  MmArena ar(workspace, workspace_bytes);
  CrWs w = carve_cr(ar, n_det, total_rows, out_size, taps);
  if (!ar.ok()) return MMMOT_E_WORKSPACE;
  NormP nm;
  for (int c = 0; c < 3; c++) { nm.mean[c] = mean_std[c]; nm.stdv[c] = mean_std[3 + c]; }
  const int S = out_size;
  resize_coeff_kernel<<<dim3(mm_cdiv(S, 128), 2 * n_det), 128, 0, st>>>(boxes, n_det, S, taps, w.bounds, w.kk);
  MM_LAUNCH_CHECK();
  resize_h_kernel<<<dim3(mm_cdiv((long)max_crop_h * S, 256), n_det), 256, 0, st>>>(image, img_h, img_w, boxes, row_off, S,
                                                                                   taps, w.bounds, w.kk, w.tmp);
  MM_LAUNCH_CHECK();
  resize_v_kernel<<<dim3(mm_cdiv(S * S, 256), n_det), 256, 0, st>>>(w.tmp, row_off, S, taps, w.bounds, w.kk, nm, out);
  MM_LAUNCH_CHECK();
  return 0;
}
