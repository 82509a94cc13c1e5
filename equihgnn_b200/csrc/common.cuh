// Shared device helpers for libeqh_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_set>

#include "eqh_b200.h"

#ifndef __CUDA_ARCH__
#define EQH_HOST_ONLY 1
#endif

namespace eqh {

constexpr int kSMs = 148;           // B200: 2 dies x 74 SMs
constexpr int kPartialRows = 296;   // persistent kernels with per-CTA partials launch 2 CTAs / SM
constexpr unsigned kFull = 0xffffffffu;

#define EQH_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return EQH_ERR_ARG; \
  } while (0)

#define EQH_LAUNCH_RC()                                  \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    return e__ == cudaSuccess ? EQH_OK : (int)e__;       \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// The step is ~150 small kernels back to back: the launch latency and the prologue of kernel N+1 are overlapped with
// the tail of kernel N.  Every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_prologue(): it first lets ITS dependents be scheduled (their CTAs then park in
// their own griddepcontrol.wait) and then waits until the grid it depends on has completed and flushed its
// memory.  Nothing touches global memory before the wait, so ordering is exactly that of a plain stream; the
// instructions are no-ops when a kernel is launched without the attribute (EQH_PDL=0, or a foreign predecessor).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
#ifndef EQH_PDL_NO_EARLY_TRIGGER
  pdl_launch_dependents();
#endif
  pdl_wait();
}

#ifndef __CUDA_ARCH__
static inline bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("EQH_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}
#endif

// Every kernel of the library asks for the MAXIMUM shared-memory carve-out, also those that use none.  An SM changes its
// L1 / shared-memory split only when it is empty: a CTA of a small kernel launched under the driver's default (small)
// carve-out keeps every 194-200 KB GEMM / edge-kernel CTA of a parallel graph branch off that SM until it has drained, and
// vice versa (measured in the step timeline: the P|Q GEMM started only when the neighbour search next to it retired its
// first CTAs, row kernels next to a grouped weight-gradient launch took 19 us instead of 4).  With one split everywhere
// the branches of the captured step really share SMs.  EQH_SMEM_CARVEOUT=default restores the driver's choice.
static inline void prefer_max_smem(const void* kernel) {
#ifndef __CUDA_ARCH__
  static const bool on = []() {
    const char* e = getenv("EQH_SMEM_CARVEOUT");
    return !(e && e[0] == 'd');
  }();
  if (!on) return;
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.insert(kernel).second)
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#endif
}

template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  prefer_max_smem(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifndef __CUDA_ARCH__
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
#endif
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through cudaGetLastError()
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One feature row of C = 32*V floats spread over a warp.  Lane l owns, for each chunk q,
// the W consecutive floats starting at column (q*32 + l)*W, W = min(V,4): every warp-wide access
// is a fully coalesced run of 32*W*4 bytes (512 B for W=4 -> 128-bit loads).
template <int V>
struct Row {
  static constexpr int W = (V >= 4) ? 4 : V;
  static constexpr int Q = V / W;
  float v[V];

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float* p = row + (q * 32 + lane) * W;
      if constexpr (W == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
      } else if constexpr (W == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[q * 2 + 0] = t.x; v[q * 2 + 1] = t.y;
      } else {
        v[q] = __ldg(p);
      }
    }
  }
  // same layout from shared (or any non-read-only) memory
  __device__ __forceinline__ void load_plain(const float* row, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float* p = row + (q * 32 + lane) * W;
      if constexpr (W == 4) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
      } else if constexpr (W == 2) {
        float2 t = *reinterpret_cast<const float2*>(p);
        v[q * 2 + 0] = t.x; v[q * 2 + 1] = t.y;
      } else {
        v[q] = *p;
      }
    }
  }
  // asynchronous copy of a row into shared memory, each lane the chunks it will read back with load_plain
  __device__ __forceinline__ static void copy_async(float* dst, const float* __restrict__ row, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (q * 32 + lane) * W);
      const float* p = row + (q * 32 + lane) * W;
      if constexpr (W == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p) : "memory");
      else if constexpr (W == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(p) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(p) : "memory");
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ row, int lane) const {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float* p = row + (q * 32 + lane) * W;
      if constexpr (W == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else if constexpr (W == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[q * 2 + 0], v[q * 2 + 1]);
      } else {
        *p = v[q];
      }
    }
  }
  __device__ __forceinline__ float sum() const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += v[i];
    return s;
  }
};

// 1 / max(1, rows in segment `row`) from a CSR rowptr; 1 when rowptr is NULL
__device__ __forceinline__ float inv_count(const int32_t* __restrict__ rowptr, int32_t row) {
  if (!rowptr) return 1.f;
  int32_t c = __ldg(rowptr + row + 1) - __ldg(rowptr + row);
  return 1.f / (float)max(c, 1);
}

// Dispatch on V = C/32 for the supported widths C in {32,64,128,256,512}.
#define EQH_DISPATCH_V(C, ...)                                   \
  do {                                                           \
    switch ((C) / 32) {                                          \
      case 1: { constexpr int V = 1; __VA_ARGS__; } break;       \
      case 2: { constexpr int V = 2; __VA_ARGS__; } break;       \
      case 4: { constexpr int V = 4; __VA_ARGS__; } break;       \
      case 8: { constexpr int V = 8; __VA_ARGS__; } break;       \
      case 16: { constexpr int V = 16; __VA_ARGS__; } break;     \
      default: return EQH_ERR_UNSUPPORTED;                       \
    }                                                            \
  } while (0)

static inline bool supported_C(int64_t C) { return C == 32 || C == 64 || C == 128 || C == 256 || C == 512; }

// fp32 parity path: accurate expf and IEEE division (the tolerance is 1e-5 relative)
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
// d/dx silu(x) = s + x*s*(1-s), s = sigmoid(x)
__device__ __forceinline__ float dsilu_f(float x) {
  float s = 1.f / (1.f + expf(-x));
  return s * (1.f + x * (1.f - s));
}

}  // namespace eqh
