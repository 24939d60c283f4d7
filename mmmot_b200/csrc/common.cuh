// Shared helpers for libmmmot_sm100a.so (sm_100a only; no torch headers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmmot_b200.h"

#define MM_CUDA(x)                                    \
  do {                                                \
    cudaError_t _e = (x);                             \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

#define MM_LAUNCH_CHECK()                             \
  do {                                                \
    mm_count_launch();                                \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

#define MM_TRY(x)                                     \
  do {                                                \
    int _r = (x);                                     \
    if (_r != 0) return _r;                           \
  } while (0)

void mm_count_launch();
int mm_debug_flags();
int mm_kseg_chunks();  // K-segment length (in 32-wide chunks) of the tcgen05 conv engine, 0 = off (mmmot_set_kseg)
int mm_engine();  // 0 auto, 1 FP32 FFMA engine, 2 tcgen05 engine (mmmot_set_engine)
bool mm_timing_on();
void mm_timing_begin(cudaStream_t st, double flop);
void mm_timing_end(cudaStream_t st);

static inline size_t mm_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int mm_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// Bump allocator over the caller-provided workspace.
struct MmArena {
  char* base;
  size_t cap, off;
  bool dry;  // dry run: only measure
  MmArena(void* p, size_t c) : base((char*)p), cap(c), off(0), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = mm_align(n * sizeof(T));
    char* r = dry ? nullptr : base + off;
    off += bytes;
    return (T*)r;
  }
  bool ok() const { return dry || off <= cap; }
};

__device__ __forceinline__ float mm_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
