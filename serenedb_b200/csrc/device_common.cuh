// device_common.cuh -- small device helpers shared by the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sdbg {

constexpr uint32_t kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

// Streaming 16-byte load that does not allocate in L1 (column scans touch every byte once).
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// Read-only 16-byte load that stays cacheable (posting payloads and descriptors are re-read by
// neighbouring windows and by other queries of a batch).
__device__ __forceinline__ uint4 ld_ro_v4(const uint4* p) { return __ldg(p); }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, v, o);
    if (lane >= uint32_t(o)) v += t;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Sortable top-k key: positive fp32 scores order like their bit patterns; the low half holds the
// complemented global doc ordinal so that, among equal scores, the smaller ordinal wins a
// "larger key" comparison. Canonical order = (score desc, segment asc, doc asc).
__device__ __host__ __forceinline__ unsigned long long make_key(float score, uint32_t ordinal) {
#ifdef __CUDA_ARCH__
  const uint32_t bits = __float_as_uint(score);
#else
  uint32_t bits; memcpy(&bits, &score, 4);
#endif
  return (static_cast<unsigned long long>(bits) << 32) | static_cast<unsigned long long>(~ordinal);
}

// In-place bitonic sort (descending) of n = power-of-two 64-bit keys in shared memory by the
// whole CTA. Ends with a barrier.
__device__ __forceinline__ void block_sort_desc(unsigned long long* s, uint32_t n) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t p = i ^ j;
        if (p > i) {
          const unsigned long long a = s[i], b = s[p];
          const bool desc = (i & k) == 0;
          if ((a < b) == desc) { s[i] = b; s[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace sdbg
