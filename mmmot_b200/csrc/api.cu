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
