// Shared device/host helpers for libmmnas_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MMNAS_OK 0
#define MMNAS_ERR_ARG (-1)        // bad argument (null pointer, unsupported size)
#define MMNAS_ERR_UNSUPPORTED (-2)
#define MMNAS_ERR_CUDA (-3)       // a CUDA runtime/driver call failed; see mmnas_last_error()

void mmnas_set_error(const char* msg);

#define MMNAS_CHECK_ARG(cond, msg)                      \
  do {                                                  \
    if (!(cond)) {                                      \
      mmnas_set_error(msg);                             \
      return MMNAS_ERR_ARG;                             \
    }                                                   \
  } while (0)

#define MMNAS_CUDA(call)                                \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) {                           \
      mmnas_set_error(cudaGetErrorString(e__));         \
      return MMNAS_ERR_CUDA;                            \
    }                                                   \
  } while (0)

#define MMNAS_LAUNCH_CHECK() MMNAS_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Counter-based dropout RNG.  key = f(seed, step, site, call); one 64-bit mix per element.
// rng_state is a device array {seed, step}; null or p == 0 disables dropout.
// ---------------------------------------------------------------------------------------------
struct DropCfg {
  const unsigned long long* state;  // device {seed, step} or null
  unsigned long long salt;          // site id and per-call salt, folded on the host
  unsigned int thresh;              // drop when rand16 < thresh  (thresh = round(p * 65536))
  float scale;                      // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint64_t mmnas_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t drop_key(const DropCfg& d) {
  return mmnas_mix64(d.state[0] ^ mmnas_mix64(d.state[1] * 0xD1B54A32D192ED03ull + d.salt));
}

// keep-multiplier (0 or scale) for element idx
__device__ __forceinline__ float drop_mult(uint64_t key, uint64_t idx, unsigned thresh, float scale) {
  uint64_t r = mmnas_mix64(key ^ ((idx >> 2) * 0x9E3779B97F4A7C15ull));
  unsigned v = (unsigned)(r >> (16 * (idx & 3))) & 0xFFFFu;
  return v < thresh ? 0.f : scale;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel launched through mmnas_launch() may start while its
// predecessor in the stream is still draining: its prologue (barrier init, TMEM allocation, descriptor
// prefetch, constant setup) overlaps the predecessor's tail.  Contract inside such a kernel: every thread
// executes pdl_wait() before its first global-memory access (read OR write); pdl_launch() tells the scheduler
// that the next kernel may be scheduled once all CTAs of this grid have issued it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool mmnas_pdl_enabled();
void mmnas_count_launch();      // process-wide count of kernels launched by this library (mmnas_launch_count())

template <typename... KArgs, typename... Args>
inline cudaError_t mmnas_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mmnas_pdl_enabled() ? 1 : 0;
  mmnas_count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
