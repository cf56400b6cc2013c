// Test-only: run simple CUDA kernels of transformers_b200/csrc on the HOST, one std::thread per CUDA thread, so that their
// index arithmetic, warp shuffles and block barriers can be checked in the GPU-less container (tests/test_kernels_emulated_cpu.py).
// Blocks run one after the other; __shared__ becomes a static (shared by the threads of the running block); a shuffle is an
// exchange through a per-warp slot array between two warp barriers, so -- like on the device with a full mask -- all 32
// lanes must reach it (a divergent shuffle deadlocks here and the test times out).  Nothing in the product includes this.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace emu {
struct Idx {
  unsigned x = 0, y = 0, z = 0;
};
struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<std::array<uint64_t, 32>> slots;  // one exchange array per warp
};
inline thread_local Idx t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_ctx = nullptr;

template <typename F>
void launch(dim3 grid, dim3 block, F&& kernel_call) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        BlockCtx ctx;
        ctx.block_bar = std::make_unique<std::barrier<>>(nthreads);
        ctx.slots.resize(nwarps);
        for (unsigned w = 0; w < nwarps; ++w) {
          const unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
          ctx.warp_bar.push_back(std::make_unique<std::barrier<>>(lanes));
        }
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          ts.emplace_back([&, t] {
            t_threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            t_blockIdx = {bx, by, bz};
            t_blockDim = {block.x, block.y, block.z};
            t_gridDim = {grid.x, grid.y, grid.z};
            t_ctx = &ctx;
            kernel_call();
          });
        for (auto& th : ts) th.join();
      }
}

inline unsigned linear_tid() { return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z); }

template <typename T>
T shfl_xor(T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const unsigned tid = linear_tid(), w = tid / 32, lane = tid % 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  t_ctx->slots[w][lane] = bits;
  t_ctx->warp_bar[w]->arrive_and_wait();
  const uint64_t got = t_ctx->slots[w][lane ^ static_cast<unsigned>(o)];
  t_ctx->warp_bar[w]->arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

inline void __syncthreads() { emu::t_ctx->block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_ctx->warp_bar[emu::linear_tid() / 32]->arrive_and_wait(); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
  return emu::shfl_xor(v, o);
}
template <typename T>
inline T __ldg(const T* p) {
  return *p;
}
template <typename T>
inline T __ldcs(const T* p) {
  return *p;
}
template <typename T>
inline void __stcs(T* p, T v) {
  *p = v;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __frcp_rn(float x) { return 1.0f / x; }
#include <mutex>
namespace emu {
inline std::mutex g_atomic_mutex;
}
inline int atomicExch(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const int old = *p;
  *p = v;
  return old;
}
inline int atomicAdd(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const int old = *p;
  *p = old + v;
  return old;
}
inline __nv_bfloat162 atomicAdd(__nv_bfloat162* p, __nv_bfloat162 v) {
  std::lock_guard<std::mutex> g(emu::g_atomic_mutex);
  const __nv_bfloat162 old = *p;
  const float2 a = __bfloat1622float2(old), b = __bfloat1622float2(v);
  *p = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
  return old;
}
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return std::exp(x); }

namespace b200 {
inline float fast_exp2(float x) { return std::exp2(x); }
inline float fast_tanh(float x) { return std::tanh(x); }
inline int num_sms() { return 148; }
inline void set_last_error(const char*, ...) {}
}  // namespace b200
#define B200_OK 0
#define B200_ERR_INVALID (-22)
#define B200_ERR_CUDA (-5)
#define B200_REQUIRE(cond, ...) \
  do {                          \
    if (!(cond)) return -22;    \
  } while (0)
#define B200_CHECK_CUDA(expr) \
  do {                        \
  } while (0)
