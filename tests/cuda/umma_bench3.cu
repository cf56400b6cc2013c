// Does tcgen05.commit cost tensor-pipe time?  Cycles per MMA with a commit every C MMAs (C = 0: only one at the end).
#include "ptx.cuh"
#include <cstdio>
using namespace b200;

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters, int every, int nbar) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8];
  __shared__ uint64_t fin;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&fin, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B), db = make_smem_desc(smem_u32(smem + 65536), 16, 1024, SWZ_128B);
    long long t0 = clock64();
    int c = 0, b = 0;
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      if (TS) umma_ts(tm + 256, tm + k * 8, desc_advance(db, k * 32), idesc, 1);
      else umma_ss(tm + 256, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, 1);
      if (every && ++c == every) { c = 0; umma_commit(&bar[b]); b = (b + 1) % nbar; }
    }
    umma_commit(&fin);
    mbar_wait(&fin, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int TS> void run(long long* d, int sms, int every, int nbar) {
  const int iters = 16384, smem = 65536 + 65536 + 1024;
  cudaFuncSetAttribute(bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, TS><<<sms, 128, smem>>>(d, 1024, every, nbar);
  bench<N, TS><<<sms, 128, smem>>>(d, iters, every, nbar);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("%s N=%3d commit every %2d MMAs (%d barriers): %6.1f cycles/MMA (ideal %d)  %s\n", TS ? "TS" : "SS", N, every, nbar, avg / iters, N < 128 ? 48 : N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; cudaMalloc(&d, 256 * 8);
  for (int ev : {0, 16, 8, 4, 2, 1}) run<256, 0>(d, sms, ev, 4);
  for (int ev : {0, 16, 8, 4}) run<64, 0>(d, sms, ev, 4);
  for (int ev : {0, 8, 4}) run<128, 1>(d, sms, ev, 4);
  run<256, 0>(d, sms, 4, 1);
  return 0;
}
