// Fusion modules A/B/C and the detection-score branch.
// Replaces reference modules/fusion_net.py:31-42 (C), :62-70 (B), :85-92 (A) and
// modules/tracking_net.py:149-163 (determine_det, eval) with w_det from :92-100.
// GroupNorm(D,D) here normalises each channel over the L detections of one frame-pair.
#include "gemm_simt.cuh"
#include "norm_ops.cuh"

namespace {

// stack2[pair][c][l] from the pre-norm linear outputs.
//   A: GN(Yp)                      B: GN(Yp) + GN(Yi)
//   C: (s(Gp)*GN(Yp) + s(Gi)*GN(Yi)) / (s(Gp) + s(Gi))
__global__ void fusion_combine_kernel(int arch, const float* __restrict__ yp, const float* __restrict__ yi,
                                      const float* __restrict__ gp, const float* __restrict__ gi,
                                      const float* __restrict__ scp, const float* __restrict__ shp,
                                      const float* __restrict__ sci, const float* __restrict__ shi, int pairs,
                                      int L, float* __restrict__ feats) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)pairs * 512 * L) return;
  int l = (int)(idx % L);
  long t = idx / L;
  int c = (int)(t % 512), pair = (int)(t / 512);
  int gc = pair * 512 + c;
  float v = fmaf(yp[idx], scp[gc], shp[gc]);
  if (arch != MMMOT_FUSION_A) {
    float u = fmaf(yi[idx], sci[gc], shi[gc]);
    if (arch == MMMOT_FUSION_B) {
      v = v + u;
    } else {
      float a = mm_sigmoid(gp[idx]), b = mm_sigmoid(gi[idx]);
      v = (a * v + b * u) / (a + b);
    }
  }
  feats[(((long)pair * 3 + 2) * 512 + c) * L + l] = v;
}

// det_scores[g][l] = s - [s < thr],  s = sigmoid(a) if 'cls' in score_arch else a,  a = w3 . h2[g][:, l] + b3
// (tracking_net.py:153-162; the threshold step is the eval branch only)
__global__ void det_score_kernel(const float* __restrict__ h2, const float* __restrict__ w3,
                                 const float* __restrict__ b3, int flags, float thr, int G, int L,
                                 float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * L) return;
  int g = idx / L, l = idx - g * L;
  const float* col = h2 + (long)g * 256 * L + l;
  float a = b3[0];
  for (int c = 0; c < 256; c++) a = fmaf(w3[c], col[(long)c * L], a);
  float s = (flags & MMMOT_SCORE_SIGMOID) ? mm_sigmoid(a) : a;
  out[idx] = ((flags & MMMOT_SCORE_THRESHOLD) && s < thr) ? s - 1.0f : s;
}

struct FdWs {
  float *yp, *yi, *gp, *gi, *scp, *shp, *sci, *shi, *h1, *h2;
  double* stats;
  double2* part;
};
FdWs carve(MmArena& a, int pairs, int L) {
  FdWs w;
  size_t n = (size_t)pairs * 512 * L;
  w.yp = a.take<float>(n); w.yi = a.take<float>(n); w.gp = a.take<float>(n); w.gi = a.take<float>(n);
  w.scp = a.take<float>((size_t)pairs * 512); w.shp = a.take<float>((size_t)pairs * 512);
  w.sci = a.take<float>((size_t)pairs * 512); w.shi = a.take<float>((size_t)pairs * 512);
  w.h1 = a.take<float>(3 * n);
  w.h2 = a.take<float>(3 * n / 2);
  w.stats = a.take<double>((size_t)pairs * 512 * 2);
  w.part = a.take<double2>((size_t)pairs * mm_cdiv(L, 128) * 512);
  return w;
}

}  // namespace

extern "C" size_t mmmot_fusion_det_workspace(int pairs, int L) {
  MmArena a(nullptr, 0);
  carve(a, pairs, L);
  return a.off;
}

extern "C" int mmmot_fusion_det_fwd(const mmmot_weights* wts, int fusion_arch, int score_flags, float neg_threshold,
                                    int pairs, int L, float* feats, float* det_scores, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!wts || !feats || !det_scores || !workspace || pairs <= 0 || L <= 0) return MMMOT_E_ARG;
  if (fusion_arch < MMMOT_FUSION_A || fusion_arch > MMMOT_FUSION_C) return MMMOT_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  MmArena ar(workspace, workspace_bytes);
  FdWs w = carve(ar, pairs, L);
  if (!ar.ok()) return MMMOT_E_WORKSPACE;
  const int tpg = mm_cdiv(L, 128);
  const long fs = 3L * 512 * L;  // floats per pair in feats

  // linear (+ optional GroupNorm statistics) on one modality: X = feats[pair][stack]
  auto linear = [&](int wt, int wb, int K, int stack, float* Y, bool stats, int gw, int gb, float* sc,
                    float* sh) -> int {
    GemmP p = gemm_defaults();
    p.Wt = wts->w[wt]; p.bias = wts->w[wb]; p.ldw = 512; p.M = 512; p.K = K;
    p.S = L; p.tiles_per_group = tpg; p.num_tiles = tpg * pairs;
    p.X = feats + (long)stack * 512 * L; p.x_gs = fs; p.x_ks = L;
    p.Y = Y; p.y_gs = 512L * L; p.y_ms = L;
    p.part = stats ? w.part : nullptr;
    MM_TRY(gemm_simt_launch<XM_DIRECT>(p, st));
    if (stats) MM_TRY(stats_reduce(w.part, 512, pairs, tpg, nullptr, w.stats, st));
    if (stats) MM_TRY(gn_finalize(w.stats, wts->w[gw], wts->w[gb], nullptr, L, pairs, 512, 1, sc, sh, st));
    return 0;
  };

  if (fusion_arch == MMMOT_FUSION_A) {
    // input_w: D x 2D on the concatenation [image; points] = stacks 0 and 1, contiguous in feats
    MM_TRY(linear(MMMOT_W_FU_WPT, MMMOT_W_FU_BP, 1024, 0, w.yp, true, MMMOT_W_FU_GPW, MMMOT_W_FU_GPB, w.scp, w.shp));
  } else {
    // NB the reference applies input_p / gate_p to stack 0 (image): names are swapped there.
    MM_TRY(linear(MMMOT_W_FU_WPT, MMMOT_W_FU_BP, 512, 0, w.yp, true, MMMOT_W_FU_GPW, MMMOT_W_FU_GPB, w.scp, w.shp));
    MM_TRY(linear(MMMOT_W_FU_WIT, MMMOT_W_FU_BI, 512, 1, w.yi, true, MMMOT_W_FU_GIW, MMMOT_W_FU_GIB, w.sci, w.shi));
    if (fusion_arch == MMMOT_FUSION_C) {
      MM_TRY(linear(MMMOT_W_FU_GATE_PT, MMMOT_W_FU_GATE_PB, 512, 0, w.gp, false, 0, 0, nullptr, nullptr));
      MM_TRY(linear(MMMOT_W_FU_GATE_IT, MMMOT_W_FU_GATE_IB, 512, 1, w.gi, false, 0, 0, nullptr, nullptr));
    }
  }
  fusion_combine_kernel<<<mm_cdiv((long)pairs * 512 * L, 256), 256, 0, st>>>(
      fusion_arch, w.yp, w.yi, w.gp, w.gi, w.scp, w.shp, w.sci, w.shi, pairs, L, feats);
  MM_LAUNCH_CHECK();

  // w_det on all three stacks: groups g = pair*3 + stack, BN(eval) folded, ReLU in the epilogue
  const int G = pairs * 3;
  {
    GemmP p = gemm_defaults();
    p.Wt = wts->w[MMMOT_W_WD_W1T]; p.bias = wts->w[MMMOT_W_WD_B1]; p.ldw = 512; p.M = 512; p.K = 512;
    p.S = L; p.tiles_per_group = tpg; p.num_tiles = tpg * G;
    p.X = feats; p.x_gs = 512L * L; p.x_ks = L;
    p.Y = w.h1; p.y_gs = 512L * L; p.y_ms = L;
    p.relu = 1;
    MM_TRY(gemm_simt_launch<XM_DIRECT>(p, st));
    p.Wt = wts->w[MMMOT_W_WD_W2T]; p.bias = wts->w[MMMOT_W_WD_B2]; p.ldw = 256; p.M = 256;
    p.X = w.h1;
    p.Y = w.h2; p.y_gs = 256L * L;
    MM_TRY(gemm_simt_launch<XM_DIRECT>(p, st));
  }
  det_score_kernel<<<mm_cdiv((long)G * L, 128), 128, 0, st>>>(w.h2, wts->w[MMMOT_W_WD_W3],
                                                              wts->w[MMMOT_W_WD_B3], score_flags, neg_threshold, G, L,
                                                              det_scores);
  MM_LAUNCH_CHECK();
  return 0;
}
