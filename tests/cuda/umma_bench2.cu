// Interference probe: MMA issue rate (SS N=64 / TS N=128 mix as in attn_bwd_dkdv) while W other warps hammer TMEM with
// tcgen05.ld (x32) / tcgen05.st, or shared memory with LDS.  Reports cycles per MMA and TMEM-load bytes per cycle.
#include "ptx.cuh"
#include <cstdio>
using namespace b200;

template <int MODE>  // 0: no side traffic, 1: tcgen05.ld, 2: tcgen05.ld + st, 3: LDS.128 broadcast
__global__ void __launch_bounds__(640, 1) bench(long long* out, int iters, int nwarps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); done = 0; }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t id64 = make_idesc_bf16(128, 64, 0, 0), id128 = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        for (int k = 0; k < 16; ++k)
          umma_ss(tm + 256 + (k >> 3) * 64, make_smem_desc(a + (k & 3) * 32, 16, 1024, SWZ_128B), make_smem_desc(b + (k & 3) * 32, 16, 1024, SWZ_128B), id64, 1);
        for (int k = 0; k < 8; ++k)
          umma_ts(tm + 384, tm + (k & 3) * 8, make_smem_desc(b + (k & 3) * 2048, 8192, 1024, SWZ_128B), id128, 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      out[blockIdx.x * 4 + 0] = t1 - t0;
      done = 1;
    }
  } else if (warp - 1 < nwarps && MODE > 0) {
    const uint32_t tl = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    long long n = 0; uint32_t acc = 0;
    long long t0 = clock64();
    const float4* sp = reinterpret_cast<const float4*>(smem + 32768);
    while (!done) {
      if (MODE == 3) {
        float4 v = sp[(n & 63)]; acc += __float_as_uint(v.x) + __float_as_uint(v.w);
      } else {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tl + 128 + ((warp >> 2) & 1) * 32, r);
        tmem_ld_wait();
        acc += r[0] ^ r[31];
        if (MODE == 2) { uint32_t w[16]; for (int e = 0; e < 16; ++e) w[e] = r[e]; tmem_st_32x32b_x16(tl + 192 + ((warp >> 2) & 1) * 16, w); tmem_st_wait(); }
      }
      ++n;
    }
    long long t1 = clock64();
    if (lane == 0) { atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 1], (unsigned long long)n); out[blockIdx.x * 4 + 2] = t1 - t0; out[blockIdx.x * 4 + 3] = acc; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int MODE> void run(const char* name, long long* d, int sms, int nwarps) {
  const int iters = 400, smem = 65536 + 65536 + 1024;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d, 0, 1024 * 8);
  bench<MODE><<<sms, 640, smem>>>(d, iters, nwarps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double per_iter = double(h[0]) / iters;
  double ld_bytes_per_cyc = MODE == 3 ? 0 : double(h[1]) * 32 * 32 * 4 / double(h[2] ? h[2] : 1);
  printf("%-22s side warps %2d : %7.1f cycles per 16xSS64+8xTS128 (ideal 1280), side ops %lld, tmem-ld %.0f B/cyc/SM  %s\n", name, nwarps, per_iter, h[1], ld_bytes_per_cyc,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; cudaMalloc(&d, 1024 * 8);
  run<0>("no side traffic", d, sms, 0);
  for (int w : {4, 8, 16}) run<1>("tcgen05.ld x32", d, sms, w);
  for (int w : {4, 8, 16}) run<2>("tcgen05.ld+st", d, sms, w);
  for (int w : {8, 16}) run<3>("LDS.128 broadcast", d, sms, w);
  return 0;
}
