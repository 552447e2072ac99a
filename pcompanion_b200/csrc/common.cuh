// Shared helpers for the pcompanion_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pcompanion_b200.h"

namespace pc {

void set_error(const char* fmt, ...);

#define PC_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      ::pc::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)

#define PC_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::pc::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PC_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define PC_LAUNCH_CHECK() PC_CUDA(cudaPeekAtLastError())

static inline cudaStream_t as_stream(pc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

// true the first time it is called for the CURRENT device with this flag array (per-kernel, per-device one-time setup such
// as cudaFuncSetAttribute: a process may drive several GPUs)
static inline bool first_use_on_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }

// sum over the GROUP consecutive lanes that share a head (GROUP = 32 / heads); every lane gets it
template <int GROUP>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
  return x;
}

// streaming / read-only 128-bit loads
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// counter-based keep mask for attention dropout: three 32-bit multiply-xorshift rounds over (seed, dst), src, head
// (the 64-bit murmur finaliser used first cost ~35 of the ~160 instructions per edge of the src-major backward)
__device__ __forceinline__ uint32_t mix_hash(uint64_t seed, uint32_t dst, uint32_t src, uint32_t head) {
  uint32_t x = dst * 0x9E3779B1u + uint32_t(seed);
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= src * 0x85EBCA77u + uint32_t(seed >> 32);
  x ^= x >> 13;
  x *= 0x297A2D39u;
  x += head * 0xC2B2AE3Du;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  return x;
}
// Keep bits of one edge for all H heads (bit h = kept): the (seed, dst, src) rounds are shared, only the last round depends
// on the head.  The GAT kernels call this ONCE per edge in one lane (the lane that loaded the edge's id) and hand the bits
// to the arithmetic by shuffle, instead of evaluating the hash in every lane of every head for every edge.
template <int H>
__device__ __forceinline__ uint32_t keep_bits(uint64_t seed, uint32_t dst, uint32_t src, uint32_t drop_threshold) {
  uint32_t x = dst * 0x9E3779B1u + uint32_t(seed);
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= src * 0x85EBCA77u + uint32_t(seed >> 32);
  x ^= x >> 13;
  x *= 0x297A2D39u;
  uint32_t bits = 0;
#pragma unroll
  for (uint32_t h = 0; h < uint32_t(H); ++h) {
    uint32_t y = x + h * 0xC2B2AE3Du;
    y ^= y >> 16;
    y *= 0x7FEB352Du;
    y ^= y >> 15;
    bits |= uint32_t(y >= drop_threshold) << h;
  }
  return bits;
}
// returns 0 (dropped) or 1/(1-p)
__device__ __forceinline__ float keep_scale(uint64_t seed, uint32_t dst, uint32_t src, uint32_t head,
                                            uint32_t drop_threshold, float inv_keep) {
  return mix_hash(seed, dst, src, head) >= drop_threshold ? inv_keep : 0.f;
}

}  // namespace pc
