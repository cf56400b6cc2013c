// Collectives of the tensor-parallel path done by our own kernels over NVLink peer memory instead of NCCL
// (the two reductions per block of the reference's tp_plan, distributed/tensor_parallel.py:320-328).
//
// Every rank writes the partial output of its rowwise GEMM into a buffer that all peers have mapped (symmetric
// allocation: same size on every rank, base pointers exchanged once).  After a device-side barrier each rank PULLS the rows
// it owns from all peers and sums them in fp32 in rank order -- reduce-scatter as one kernel whose loads are the NVLink
// transfer, with the residual add fused into the same pass (the reference does GEMM -> all-reduce -> `residual + x` as
// three passes over [T, H]).  NVSwitch gives every GPU full bandwidth to every peer, so the direct one-hop pull moves
// (N-1)/N of one partial per rank, the minimum for a reduce-scatter, with no intermediate staging copies.
//
// Algorithmic bytes per output element: 2 B x world (one local + world-1 remote reads) + 2 B residual + 2 B write.
#include <cuda_bf16.h>

#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

namespace b200 {

constexpr int PEER_MAX_WORLD = 16;

struct PeerPtrs {
  const __nv_bfloat16* p[PEER_MAX_WORLD];
};

__device__ __forceinline__ void peer_unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// out[i] = bf16( residual[i] + sum_{s=0..world-1} peer[s][offset + i] ), i in [0, n); n % 8 == 0, 16-byte aligned.
// Loads of the `world` sources for one vector are issued back to back (independent) before the adds: world x 16 B in
// flight per thread, grid-stride so a few hundred CTAs keep the NVLink ports busy.
template <int WORLD>
__global__ void __launch_bounds__(256) pull_reduce_kernel(PeerPtrs src, int64_t offset, int64_t n8,
                                                          const uint4* __restrict__ residual, uint4* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 v[WORLD];
#pragma unroll
    for (int s = 0; s < WORLD; ++s) v[s] = reinterpret_cast<const uint4*>(src.p[s] + offset)[i];  // plain (coherent) loads: peer data
    float acc[8];
    if (residual) {
      peer_unpack8(__ldg(residual + i), acc);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    }
#pragma unroll
    for (int s = 0; s < WORLD; ++s) {
      float f[8];
      peer_unpack8(v[s], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    out[i] = o;
  }
}

#ifndef B200_HOST_EMU
template <int WORLD>
static int launch_pull_reduce(const PeerPtrs& src, int64_t offset, int64_t n8, const void* residual, void* out,
                              cudaStream_t stream) {
  int64_t blocks = (n8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms() > 0 ? num_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  pull_reduce_kernel<WORLD><<<static_cast<int>(blocks), 256, 0, stream>>>(src, offset, n8, reinterpret_cast<const uint4*>(residual),
                                                                        reinterpret_cast<uint4*>(out));
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

#endif  // B200_HOST_EMU

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

// peer_ptrs: HOST array of `world` device pointers (this rank's own buffer included, in rank order), each the base of that
// rank's partial-sum buffer; offset_elems: first element of the rows this rank owns; residual may be NULL.
extern "C" int b200_pull_reduce_bf16(const void* const* peer_ptrs, int world, int64_t offset_elems, int64_t n_elems,
                                     const void* residual, void* out, cudaStream_t stream) {
  B200_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD, "pull_reduce: world size %d not supported (1..%d)", world, PEER_MAX_WORLD);
  B200_REQUIRE(n_elems >= 0 && n_elems % 8 == 0 && offset_elems % 8 == 0, "pull_reduce: sizes must be multiples of 8 elements");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "pull_reduce: out / residual must be 16B aligned");
  if (n_elems == 0) return B200_OK;
  PeerPtrs src;
  for (int s = 0; s < world; ++s) {
    B200_REQUIRE(peer_ptrs[s] != nullptr && (reinterpret_cast<uintptr_t>(peer_ptrs[s]) & 15) == 0,
                 "pull_reduce: peer buffer %d is null or not 16B aligned", s);
    src.p[s] = reinterpret_cast<const __nv_bfloat16*>(peer_ptrs[s]);
  }
  for (int s = world; s < PEER_MAX_WORLD; ++s) src.p[s] = nullptr;
  const int64_t n8 = n_elems / 8;
  switch (world) {
    case 1: return launch_pull_reduce<1>(src, offset_elems, n8, residual, out, stream);
    case 2: return launch_pull_reduce<2>(src, offset_elems, n8, residual, out, stream);
    case 4: return launch_pull_reduce<4>(src, offset_elems, n8, residual, out, stream);
    case 8: return launch_pull_reduce<8>(src, offset_elems, n8, residual, out, stream);
    default:
      set_last_error("pull_reduce: world size %d not instantiated (1, 2, 4, 8)", world);
      return B200_ERR_INVALID;
  }
}
#endif  // B200_HOST_EMU
