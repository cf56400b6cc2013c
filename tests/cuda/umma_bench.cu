// Micro-benchmark: sustained cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) for the operand modes the
// attention / GEMM kernels use.  One CTA per SM, one thread issues `iters` back-to-back MMAs, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I transformers_b200/csrc tests/cuda/umma_bench.cu -o tests/cuda/bin/umma_bench
#include "ptx.cuh"
#include <cstdio>
using namespace b200;

template <int N, int TS, int B_MN, int A_MN>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      const uint64_t db = B_MN ? make_smem_desc(b + k * 2048, 16384, 1024, SWZ_128B) : make_smem_desc(b + k * 32, 16, 1024, SWZ_128B);
      if (TS) umma_ts(tm + 256, tm + k * 8, db, idesc, 1);
      else {
        const uint64_t da = A_MN ? make_smem_desc(a + k * 2048, 8192, 1024, SWZ_128B) : make_smem_desc(a + k * 32, 16, 1024, SWZ_128B);
        umma_ss(tm + 256, da, db, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int TS, int B_MN, int A_MN>
void run(const char* name, long long* d, int sms) {
  const int iters = 20000, smem = 65536 + 65536 + 1024;
  cudaFuncSetAttribute(bench<N, TS, B_MN, A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, TS, B_MN, A_MN><<<sms, 128, smem>>>(d, 1000);
  bench<N, TS, B_MN, A_MN><<<sms, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("%-28s N=%3d : %7.1f cycles/MMA (ideal %3d)  %s\n", name, N, avg / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; cudaMalloc(&d, 256 * sizeof(long long));
  run<64, 0, 0, 0>("SS A-K  B-K", d, sms);   run<128, 0, 0, 0>("SS A-K  B-K", d, sms);   run<256, 0, 0, 0>("SS A-K  B-K", d, sms);
  run<64, 0, 1, 0>("SS A-K  B-MN", d, sms);  run<128, 0, 1, 0>("SS A-K  B-MN", d, sms);  run<256, 0, 1, 0>("SS A-K  B-MN", d, sms);
  run<128, 0, 1, 1>("SS A-MN B-MN", d, sms); run<256, 0, 1, 1>("SS A-MN B-MN", d, sms);
  run<64, 1, 0, 0>("TS A-tmem B-K", d, sms); run<128, 1, 0, 0>("TS A-tmem B-K", d, sms); run<256, 1, 0, 0>("TS A-tmem B-K", d, sms);
  run<64, 1, 1, 0>("TS A-tmem B-MN", d, sms); run<128, 1, 1, 0>("TS A-tmem B-MN", d, sms); run<256, 1, 1, 0>("TS A-tmem B-MN", d, sms);
  return 0;
}
