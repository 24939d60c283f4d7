// Small kernels of the tensor-core (channels-last) pipeline: activations between tcgen05 contractions
// are stored [row][channel], fp32 before the GroupNorm statistics are known and packed FP16 (hi|lo)
// after normalisation.
#pragma once
#include "gemm_tc.cuh"

// out[row][c] = split(relu(in[row][c]*sc[g][c] + sh[g][c])),  g = seg ? seg[row] / L : row / rows_per_group.
// GroupNorm + ReLU of the producer layer applied once per element, emitted as the packed FP16 (hi|lo)
// words the next contraction's operand producers copy straight into shared memory.
static __global__ void norm_split_kernel(const float* __restrict__ in, long ldi, const float* __restrict__ sc,
                                         const float* __restrict__ sh, int C, long rows, int rows_per_group,
                                         const int* __restrict__ seg, int L, uint32_t* __restrict__ out, long ldo) {
  const int c4n = C >> 2;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c4n) return;
  const long row = idx / c4n;
  const int c = (int)(idx - row * c4n) * 4;
  const int g = seg ? seg[row] / L : (int)(row / rows_per_group);
  const float4 x = *reinterpret_cast<const float4*>(in + row * ldi + c);
  const float4 a = *reinterpret_cast<const float4*>(sc + (long)g * C + c);
  const float4 b = *reinterpret_cast<const float4*>(sh + (long)g * C + c);
  uint4 o;
  o.x = tc::pack_split_f16(fmaxf(fmaf(x.x, a.x, b.x), 0.f));
  o.y = tc::pack_split_f16(fmaxf(fmaf(x.y, a.y, b.y), 0.f));
  o.z = tc::pack_split_f16(fmaxf(fmaf(x.z, a.z, b.z), 0.f));
  o.w = tc::pack_split_f16(fmaxf(fmaf(x.w, a.w, b.w), 0.f));
  *reinterpret_cast<uint4*>(out + row * ldo + c) = o;
}

static inline int norm_split(const float* in, long ldi, const float* sc, const float* sh, int C, long rows,
                             int rows_per_group, const int* seg, int L, uint32_t* out, long ldo, cudaStream_t st) {
  norm_split_kernel<<<mm_cdiv(rows * (C / 4), 256), 256, 0, st>>>(in, ldi, sc, sh, C, rows, rows_per_group, seg, L,
                                                                  out, ldo);
  MM_LAUNCH_CHECK();
  return 0;
}
