// bf16 GEMM on tcgen05 tensor cores for sm_100a:  D[M,N] (+)= sum_k A(m,k) * B(n,k), fp32 accumulation in TMEM.
//
// Replaces the cuBLAS GEMMs behind nn.Linear on the reference hot path
// (models/llama/modeling_llama.py:174-176 MLP, :254-256 q/k/v, :280 o_proj, :480 lm_head) and their autograd
// backward (dgrad: dX = dY * W, wgrad: dW = dY^T * X), so three operand layouts are needed:
//   forward  Y = X W^T   : A = X  [M,K] K-major,   B = W  [N,K] K-major
//   dgrad    dX = dY W   : A = dY [M,K'] K-major,  B = W  stored [K',N'] -> "MN-major" B
//   wgrad    dW = dY^T X : A = dY stored [K',M'] -> MN-major A,  B = X stored [K',N'] -> MN-major B
//
// Structure (one persistent CTA per SM, 192 threads, warp specialised):
//   warp 0   TMA producer : cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring, mbarrier expect_tx
//   warp 1   MMA issuer   : one thread issues tcgen05.mma (128x256x16, cta_group::1) per k-step; tcgen05.commit
//                           releases smem stages and publishes the accumulator; owns the TMEM allocation (512 cols =
//                           two 128x256 fp32 accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-5 epilogue    : tcgen05.ld 32 lanes x 32 columns -> registers -> bf16 -> 16-byte global stores
// Tile order is grouped along M so a wave of CTAs shares A/B tiles through the 126 MB L2.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 256 + 1024;  // ring + barriers + alignment slack

struct GemmParams {
  __nv_bfloat16* C;
  int M, N, K, ldc;
  int accumulate;  // C += D (bf16 read-modify-write) instead of C = D
  int group_m;     // M-tiles per rasterisation group
  // smem-descriptor byte offsets (leading / stride dimension) per operand; see make_smem_desc
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
};

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int& tm, int& tn) {
    const int group_size = p.group_m * num_n;
    const int group = tile / group_size;
    const int first_m = group * p.group_m;
    const int gsz = min(p.group_m, num_m - first_m);
    const int in_group = tile - group * group_size;
    tm = first_m + in_group % gsz;
    tn = in_group / gsz;
  };

  if (warp == 0) {
    if (elect_one()) {  // elect.sync: the compiler keeps UTCHMMA / UTMALDG operands in uniform registers (no per-op ELECT loop)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        tile_coords(tile, tm, tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (A_MN == 0) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, tm * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sA + j * (64 * BK * 2), &tmA, &full_bar[stage], tm * BM + j * 64, kb * BK);
          }
          if (B_MN == 0) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, tn * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * (64 * BK * 2), &tmB, &full_bar[stage], tn * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync: the compiler keeps UTCHMMA / UTMALDG operands in uniform registers (no per-op ELECT loop)
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      // byte advance of the descriptor start address per UMMA_K = 16 step
      constexpr uint32_t a_kstep = A_MN ? 16 * 128 : 32;
      constexpr uint32_t b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc(sA + k * a_kstep, p.a_lbo, p.a_sbo, SWZ_128B);
            const uint64_t db = make_smem_desc(sB + k * b_kstep, p.b_lbo, p.b_sbo, SWZ_128B);
            umma_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int tm, tn;
      tile_coords(tile, tm, tn);
      const int buf = it & 1;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
      __nv_bfloat16* crow = p.C + static_cast<size_t>(row) * p.ldc + tn * BN;
      const int ncols = min(BN, p.N - tn * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 >= ncols) break;   // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (row < p.M) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int col = c * 32 + v * 8;
            if (col + 8 <= ncols) {
              uint4* dst = reinterpret_cast<uint4*>(crow + col);
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[v * 8 + e]);
              if (p.accumulate) {
                uint4 old = *dst;
                const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 of = __bfloat1622float2(o2[e]);
                  f[2 * e] += of.x;
                  f[2 * e + 1] += of.y;
                }
              }
              uint4 o;
              o.x = pack_bf16(f[0], f[1]);
              o.y = pack_bf16(f[2], f[3]);
              o.z = pack_bf16(f[4], f[5]);
              o.w = pack_bf16(f[6], f[7]);
              *dst = o;
            } else {
              for (int e = 0; e < 8; ++e) {
                if (col + e < ncols) {
                  float f = __uint_as_float(r[v * 8 + e]);
                  if (p.accumulate) f += __bfloat162float(crow[col + e]);
                  crow[col + e] = __float2bfloat16_rn(f);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int A_MN, int B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05<A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    attr_set = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int sms = num_sms();
  if (sms <= 0) {
    set_last_error("no CUDA device");
    return B200_ERR_NODEV;
  }
  const int grid = num_tiles < sms ? num_tiles : sms;
  kern<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(tmA, tmB, p);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200

// C-ABI.  a_mn / b_mn: 0 = operand stored [rows = M or N, cols = K] (K-major); 1 = stored [rows = K, cols = M or N].
// lda / ldb / ldc are row strides in elements.  1-CTA kernel (128 x 256 tiles): used for M <= 128 (one tile tall).
extern "C" int b200_gemm_bf16_1sm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                                  int a_mn, int b_mn, int accumulate, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0, "gemm: C must be 16B aligned, ldc %% 8 == 0");
  B200_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn)
    rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, BK, BM);
  else
    rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn)
    rc = make_tmap_2d_bf16(&tmB, B, N, K, ldb, BK, BN);
  else
    rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, 64, BK);
  if (rc) return rc;

  GemmParams p;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.accumulate = accumulate;
  p.group_m = 16;
  // K-major, 128B swizzle: 8-row groups are 1024 B apart (SBO); LBO unused (one swizzle atom spans the 64-wide K tile).
  // MN-major, 128B swizzle: 64-element MN chunks are 64*BK*2 = 8192 B apart (LBO); 8-row K groups 1024 B apart (SBO).
  const uint32_t k_lbo = 16, k_sbo = 1024, mn_lbo = 64 * BK * 2, mn_sbo = 1024;
  p.a_lbo = a_mn ? mn_lbo : k_lbo;
  p.a_sbo = a_mn ? mn_sbo : k_sbo;
  p.b_lbo = b_mn ? mn_lbo : k_lbo;
  p.b_sbo = b_mn ? mn_sbo : k_sbo;
  if (!a_mn && !b_mn) return launch_gemm<0, 0>(tmA, tmB, p, stream);
  if (!a_mn && b_mn) return launch_gemm<0, 1>(tmA, tmB, p, stream);
  if (a_mn && b_mn) return launch_gemm<1, 1>(tmA, tmB, p, stream);
  return launch_gemm<1, 0>(tmA, tmB, p, stream);
}

extern "C" int b200_gemm_bf16_2sm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                                  int a_mn, int b_mn, int accumulate, cudaStream_t stream);

// Dispatcher.  Measured on the full Llama-3-8B step under the 1 kW power cap (profiles/README.md): the CTA-pair kernel
// (gemm2.cu, 256x256 tiles, 32 KB/stage/SM) sustains 1453 TF/s vs 1351 TF/s for the 1-CTA 128x256 kernel, so it is the
// default for anything taller than one tile; shapes of one tile or less use the 1-CTA kernel.
extern "C" int b200_gemv_bf16(const void* x, const void* W, void* y, int M, int N, int K, int ldx, int ldw, int ldy,
                              cudaStream_t stream);

extern "C" int b200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                              int a_mn, int b_mn, int accumulate, cudaStream_t stream) {
  // decode rows (M <= 4): stream the weights once with the CUDA-core kernel of gemv.cu (4.2-6.1 TB/s of weight stream
  // measured) instead of a 1/128-full tensor-core tile
  if (M >= 1 && M <= 4 && !a_mn && !b_mn && !accumulate && K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0)
    return b200_gemv_bf16(A, B, C, M, N, K, lda, ldb, ldc, stream);
  if (M > 128 && N > 64)
    return b200_gemm_bf16_2sm(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, accumulate, stream);
  return b200_gemm_bf16_1sm(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, accumulate, stream);
}
