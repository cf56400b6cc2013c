// bf16 GEMM on CTA pairs (tcgen05 cta_group::2): one 256x256 output tile per 2-CTA cluster.
//
// Same math / operand layouts / epilogue as gemm.cu (D[M,N] (+)= sum_k A(m,k) B(n,k), K- or MN-major operands), but the
// two CTAs of a cluster (same TPC) cooperate on UMMA_M = 256: each CTA stages its 128 rows of A and its 128-row half of
// B, so every SM loads 32 KB instead of 48 KB per 64-wide k-block (less L2->SM traffic, less smem read per MMA) and six
// pipeline stages fit.  The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 for both SMs;
// tcgen05.commit multicasts stage-release / accumulator-ready arrivals to both CTAs; the peer CTA's TMA loads complete
// on the leader's mbarrier; the peer's epilogue warps signal "accumulator drained" on the leader's barrier remotely.
#include "act.cuh"
#include "common.cuh"
#include "ptx.cuh"


namespace b200 {

constexpr int G2_BM = 256;        // per cluster
constexpr int G2_BN = 256;
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;   // per CTA: 128 rows of A
constexpr int G2_B_BYTES = 128 * G2_BK * 2;   // per CTA: 128 of the 256 B rows
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;  // 32 KB
constexpr int G2_THREADS = 192;
constexpr int G2_EPI_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 staging buffers x (32 rows x 128 B), 128B-swizzled for the TMA store
constexpr int G2_MAX_GROUPS = 64;            // experts per grouped launch
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + G2_EPI_BYTES + 256 + 4 * (G2_MAX_GROUPS + 1) + 1024;

struct Gemm2Params {
  __nv_bfloat16* C;
  int M, N, K, ldc;
  int accumulate;
  int group_m;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  // SCATTER variant only: row block r (rows_per_owner rows) of the output goes to scatter_maps[r] (tensor maps in global
  // memory, one per destination buffer -- peer-mapped memory of rank r); m_rot rotates the tile order so that the blocks
  // of the other ranks are produced (and travel over NVLink) first, the own block last
  union {
    const CUtensorMap* scatter_maps;
    const int* offsets;  // GROUPED variant (below); the two variants never combine, the fields share storage
  };
  union {
    int rows_per_owner;
    int groups;
  };
  int m_rot;
  // GLU variant only: the weight rows are block-interleaved (256-row groups = 128 gate rows + 128 up rows), so every
  // 256-column output tile holds 128 gate columns and the matching 128 up columns; the epilogue stores them (tmC, the
  // [M, 2I] interleaved gate|up matrix the backward needs) AND act(gate) * up (tmH, [M, I])
  int gelu;
  // GROUPED variant only (Mixtral experts, integrations/moe.py:377-478): the M rows of A / C are cut into `groups` row
  // ranges [offsets[g], offsets[g+1]) read from DEVICE memory (the routing kernels wrote them: no host sync), range g is
  // multiplied by B's g-th [N, K] (or [K, N]) matrix.  Tiles never straddle two ranges: every range starts its own 256-row
  // tiles, the rows a tile computes beyond its range are not stored.  (`offsets` / `groups`: the unions above.)
};

constexpr int G2_MODE_PLAIN = 0, G2_MODE_SCATTER = 2, G2_MODE_GLU = 3, G2_MODE_GROUPED = 4;

// (A soft lock-step of the persistent clusters at tile boundaries -- to keep co-running tiles walking K together and cut
// the DRAM re-reads -- was measured in round 2: no step-time gain on the Llama-3-8B shapes, so it is gone;
// profiles/r02_call3a_gemm_tuning.md.)
template <int A_MN, int B_MN, int MODE>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_bf16_tcgen05_2sm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmH, Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + G2_STAGES * G2_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + G2_EPI_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] (used in the leader CTA only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  int* g_tile_start = reinterpret_cast<int*>(tmem_slot + 2);  // GROUPED: first tile index of every row range, [groups + 1]
  constexpr bool GROUPED = MODE == G2_MODE_GROUPED;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (p.M + G2_BM - 1) / G2_BM;
  const int num_n = (p.N + G2_BN - 1) / G2_BN;
  const int num_tiles_dense = num_m * num_n;
  const int num_kb = (p.K + G2_BK - 1) / G2_BK;

  if (threadIdx.x == 0) {
    if constexpr (GROUPED) {
      int acc = 0;
      for (int g = 0; g < p.groups; ++g) {
        g_tile_start[g] = acc;
        acc += ((p.offsets[g + 1] - p.offsets[g] + G2_BM - 1) / G2_BM) * num_n;
      }
      g_tile_start[p.groups] = acc;
    }
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 256);  // 128 epilogue threads of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr bool SCATTER = MODE == G2_MODE_SCATTER;
  constexpr bool GLU = MODE == G2_MODE_GLU;
  const int num_tiles = GROUPED ? g_tile_start[p.groups] : num_tiles_dense;
  // GROUPED tile -> (row range g, first row of the tile, n tile, end of the range); `g` only moves forward (tiles ascend)
  auto group_tile = [&](int tile, int& g, int& row0, int& tn, int& row_end) {
    while (tile >= g_tile_start[g + 1]) ++g;
    const int lo = p.offsets[g];
    row_end = p.offsets[g + 1];
    const int num_m_g = (row_end - lo + G2_BM - 1) / G2_BM;
    const int local = tile - g_tile_start[g];
    row0 = lo + (local % num_m_g) * G2_BM;   // m fastest: consecutive tiles share the range's B panel
    tn = local / num_m_g;
  };

  auto tile_coords = [&](int tile, int& tm, int& tn) {
    const int group_size = p.group_m * num_n;
    const int group = tile / group_size;
    const int first_m = group * p.group_m;
    const int gsz = min(p.group_m, num_m - first_m);
    const int in_group = tile - group * group_size;
    tm = first_m + in_group % gsz;
    tn = in_group / gsz;
    if constexpr (SCATTER) {
      tm += p.m_rot;
      if (tm >= num_m) tm -= num_m;
    }
  };

  if (warp == 0) {
    if (elect_one()) {  // elect.sync: the compiler keeps UTCHMMA / UTMALDG operands in uniform registers (no per-op ELECT loop)
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] int grp = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int tm, tn, m0, n0;
        [[maybe_unused]] int b_k0 = 0;  // GROUPED with an MN-major B: the range's [K, N] matrix starts at row g * K
        if constexpr (GROUPED) {
          int row0, row_end;
          group_tile(tile, grp, row0, tn, row_end);
          m0 = row0 + rank * 128;
          n0 = tn * G2_BN + rank * 128 + (B_MN ? 0 : grp * p.N);
          b_k0 = B_MN ? grp * p.K : 0;
        } else {
          tile_coords(tile, tm, tn);
          m0 = tm * G2_BM + rank * 128;
          n0 = tn * G2_BN + rank * 128;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * G2_STAGE_BYTES;
          uint8_t* sB = sA + G2_A_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);  // both CTAs' bytes land on this barrier
          if (A_MN == 0) {
            tma_load_2d_2sm(sA, &tmA, leader_full, kb * G2_BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sA + j * (64 * G2_BK * 2), &tmA, leader_full, m0 + j * 64, kb * G2_BK);
          }
          if (B_MN == 0) {
            tma_load_2d_2sm(sB, &tmB, leader_full, kb * G2_BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_2d_2sm(sB + j * (64 * G2_BK * 2), &tmB, leader_full, n0 + j * 64, kb * G2_BK + (GROUPED ? b_k0 : 0));
          }
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(G2_BM, G2_BN, A_MN, B_MN);
      constexpr uint32_t a_kstep = A_MN ? 16 * 128 : 32;
      constexpr uint32_t b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * G2_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t sB = sA + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) {
            const uint64_t da = make_smem_desc(sA + k * a_kstep, p.a_lbo, p.a_sbo, SWZ_128B);
            const uint64_t db = make_smem_desc(sB + k * b_kstep, p.b_lbo, p.b_sbo, SWZ_128B);
            umma_ss_2sm(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 0x3);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull_bar[buf], 0x3);
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    [[maybe_unused]] int grp = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      int tm = 0, tn;
      [[maybe_unused]] int g_row0 = 0, g_row_end = 0;
      if constexpr (GROUPED) {
        group_tile(tile, grp, g_row0, tn, g_row_end);
      } else {
        tile_coords(tile, tm, tn);
      }
      const int buf = it & 1;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * G2_BN;
      const int ncols = min(G2_BN, p.N - tn * G2_BN);
      if constexpr (GLU) {
        // tile columns [0, 128) = gate, [128, 256) = up of the same 128 MLP columns (N % 256 == 0 is required).  Per 64-column
        // chunk: gate -> tmC, up -> tmC, act(gate) * up -> tmH, three TMA stores rotating over the two staging buffers
        uint8_t* stage = epi_smem + (warp - 2) * 8192;
        const int row0 = tm * G2_BM + rank * 128 + q * 32;
        int nstore = 0;
        auto stage_and_store = [&](const uint32_t (&pk)[32], const CUtensorMap* map, int col) {
          uint8_t* sbuf = stage + (nstore & 1) * 4096;
          if (nstore >= 2) {  // the store issued two stores ago must have read this buffer
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          uint8_t* srow = sbuf + lane * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            uint4 o;
            o.x = pk[v * 4 + 0];
            o.y = pk[v * 4 + 1];
            o.z = pk[v * 4 + 2];
            o.w = pk[v * 4 + 3];
            *reinterpret_cast<uint4*>(srow + ((v ^ (lane & 7)) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(map, sbuf, col, row0);
            tma_store_commit();
          }
          ++nstore;
        };
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t g0[32], g1[32], u0[32], u1[32];
          tmem_ld_32x32b_x32(taddr + c * 64, g0);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, g1);
          tmem_ld_32x32b_x32(taddr + 128 + c * 64, u0);
          tmem_ld_32x32b_x32(taddr + 128 + c * 64 + 32, u1);
          tmem_ld_wait();
          uint32_t pg[32], pu[32], ph[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const uint32_t* rg = e < 16 ? g0 : g1;
            const uint32_t* ru = e < 16 ? u0 : u1;
            const int i = (e & 15) * 2;
            pg[e] = pack_bf16(__uint_as_float(rg[i]), __uint_as_float(rg[i + 1]));
            pu[e] = pack_bf16(__uint_as_float(ru[i]), __uint_as_float(ru[i + 1]));
            // the activation sees the bf16-rounded projections, exactly like the stand-alone GLU kernel reading them back
            const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&pg[e]);
            const __nv_bfloat162 ub = *reinterpret_cast<const __nv_bfloat162*>(&pu[e]);
            const float2 gf = __bfloat1622float2(gb), uf = __bfloat1622float2(ub);
            ph[e] = pack_bf16(glu_value(gf.x, uf.x, p.gelu), glu_value(gf.y, uf.y, p.gelu));
          }
          stage_and_store(pg, &tmC, tn * G2_BN + c * 64);
          stage_and_store(pu, &tmC, tn * G2_BN + 128 + c * 64);
          stage_and_store(ph, &tmH, tn * (G2_BN / 2) + c * 64);
        }
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[buf]), 0));  // accumulator drained: tell the leader
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      } else {
        // TMEM -> registers -> bf16 -> 128B-swizzled smem tile (32 rows x 64 cols) -> one TMA store per chunk: full-line
        // writes, no per-thread global stores, M / N tails clipped by the tensor map.  accumulate: the same tile goes out
        // as a TMA reduce-add (C += tile, bf16 adds performed at L2) -- no read-modify-write through the SM
        uint8_t* stage = epi_smem + (warp - 2) * 8192;
        const int row0 = (GROUPED ? g_row0 : tm * G2_BM) + rank * 128 + q * 32;
        // GROUPED: a strip that crosses the end of its row range stores its valid rows with plain vector stores (the rows
        // beyond belong to the next range and are computed by that range's own tile); a strip wholly beyond stores nothing
        [[maybe_unused]] const bool strip_full = !GROUPED || row0 + 32 <= g_row_end;
        [[maybe_unused]] const bool lane_valid = !GROUPED || row0 + lane < g_row_end;
#pragma unroll 1
        for (int c = 0; c < G2_BN / 64; ++c) {
          if (c * 64 >= ncols) break;
          uint8_t* sbuf = stage + (c & 1) * 4096;
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + c * 64, r0);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
          if (c >= 2) {  // the buffer is being reused: the TMA store issued two chunks ago must have read it
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          tmem_ld_wait();
          uint8_t* srow = sbuf + lane * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint32_t* r = v < 4 ? r0 : r1;
            const int e = (v & 3) * 8;
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[e + 0]), __uint_as_float(r[e + 1]));
            o.y = pack_bf16(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
            o.z = pack_bf16(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
            o.w = pack_bf16(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
            if constexpr (GROUPED) {
              if (!strip_full) {
                if (lane_valid)
                  *reinterpret_cast<uint4*>(p.C + static_cast<size_t>(row0 + lane) * p.ldc + tn * G2_BN + c * 64 + v * 8) = o;
                continue;
              }
            }
            *reinterpret_cast<uint4*>(srow + ((v ^ (lane & 7)) << 4)) = o;  // 128B swizzle: 16-byte chunk index ^ (row % 8)
          }
          if constexpr (GROUPED) {
            if (!strip_full) continue;  // warp-uniform
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if constexpr (SCATTER) {  // the 32-row strip lives in exactly one owner's buffer (rows_per_owner % 256 == 0)
              const int owner = row0 / p.rows_per_owner;
              tma_store_2d(p.scatter_maps + owner, sbuf, tn * G2_BN + c * 64, row0 - owner * p.rows_per_owner);
            } else if (p.accumulate) {
              tma_reduce_add_2d(&tmC, sbuf, tn * G2_BN + c * 64, row0);
            } else {
              tma_store_2d(&tmC, sbuf, tn * G2_BN + c * 64, row0);
            }
            tma_store_commit();
          }
        }
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[buf]), 0));  // accumulator drained: tell the leader
        if (lane == 0) tma_store_wait_read<0>();  // staging buffers free before the next tile
        __syncwarp();
      }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait<0>();  // all output tiles written before the CTA retires
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int A_MN, int B_MN, int MODE>
static int launch_gemm2_v(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmH,
                          Gemm2Params p, cudaStream_t stream, int tiles_override = 0) {
  auto kern = gemm_bf16_tcgen05_2sm<A_MN, B_MN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    attr_set = true;
  }
  const int num_tiles = tiles_override > 0 ? tiles_override : ((p.M + G2_BM - 1) / G2_BM) * ((p.N + G2_BN - 1) / G2_BN);
  int sms = num_sms();
  if (sms <= 0) {
    set_last_error("no CUDA device");
    return B200_ERR_NODEV;
  }
  int clusters = sms / 2;
  if (num_tiles < clusters) clusters = num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmH, p));
  return B200_OK;
}

template <int A_MN, int B_MN>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Gemm2Params& p,
                        cudaStream_t stream) {
  if (p.scatter_maps) return launch_gemm2_v<A_MN, B_MN, G2_MODE_SCATTER>(tmA, tmB, tmC, tmC, p, stream);
  return launch_gemm2_v<A_MN, B_MN, G2_MODE_PLAIN>(tmA, tmB, tmC, tmC, p, stream);
}

}  // namespace b200

static int gemm2_run(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                     int b_mn, int accumulate, const CUtensorMap* scatter_maps, int rows_per_owner, int m_rot,
                     cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0, "gemm: C must be 16B aligned, ldc %% 8 == 0");
  B200_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements");
  CUtensorMap tmA, tmB;
  int rc;
  // per-CTA boxes: 128 rows of A / B (K-major) or 64-column x 64-k boxes (MN-major)
  if (!a_mn)
    rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, G2_BK, 128);
  else
    rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, 64, G2_BK);
  if (rc) return rc;
  if (!b_mn)
    rc = make_tmap_2d_bf16(&tmB, B, N, K, ldb, G2_BK, 128);
  else
    rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, 64, G2_BK);
  if (rc) return rc;
  CUtensorMap tmC;  // output tile store: 64-column x 32-row boxes, 128B swizzle (N tails clipped by TMA)
  rc = make_tmap_2d_bf16(&tmC, C, M, N, ldc, 64, 32);
  if (rc) return rc;
  Gemm2Params p;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.accumulate = accumulate;
  p.scatter_maps = scatter_maps;
  p.rows_per_owner = rows_per_owner;
  p.m_rot = m_rot;
  p.gelu = 0;
  p.group_m = 8;  // M tiles per rasterisation group (wave footprint ~ 8 x 9.25 tiles of 256x256); 4 and 16 measured no better
  const uint32_t k_lbo = 16, k_sbo = 1024, mn_lbo = 64 * G2_BK * 2, mn_sbo = 1024;
  p.a_lbo = a_mn ? mn_lbo : k_lbo;
  p.a_sbo = a_mn ? mn_sbo : k_sbo;
  p.b_lbo = b_mn ? mn_lbo : k_lbo;
  p.b_sbo = b_mn ? mn_sbo : k_sbo;
  if (!a_mn && !b_mn) return launch_gemm2<0, 0>(tmA, tmB, tmC, p, stream);
  if (!a_mn && b_mn) return launch_gemm2<0, 1>(tmA, tmB, tmC, p, stream);
  if (a_mn && b_mn) return launch_gemm2<1, 1>(tmA, tmB, tmC, p, stream);
  return launch_gemm2<1, 0>(tmA, tmB, tmC, p, stream);
}

// Same contract as b200_gemm_bf16 (gemm.cu); requires M > 128 to be worthwhile.  Called by the dispatcher there.
extern "C" int b200_gemm_bf16_2sm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                                  int a_mn, int b_mn, int accumulate, cudaStream_t stream) {
  return gemm2_run(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, accumulate, nullptr, 0, 0, stream);
}

// Gate|up projection with the gated activation in the epilogue (LlamaMLP.forward models/llama/modeling_llama.py:174-176,
// first two thirds): W is the [2I, K] BLOCK-INTERLEAVED gate / up weight (256-row groups: 128 gate_proj rows followed by the
// matching 128 up_proj rows); gu [M, 2I] receives the projections in the same interleaved column order (kept for the
// backward), h [M, I] = bf16(bf16(act(gate)) * up).  Bit-identical to b200_gemm_bf16 + b200_glu_fwd.  Requires M > 128
// (CTA-pair kernel) and I % 128 == 0.
extern "C" int b200_gemm_glu_bf16(const void* A, const void* W, void* gu, void* h, int M, int I, int K, int lda, int ldw, int ldgu,
                                  int ldh, int gelu, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(M > 128 && I > 0 && K > 0 && I % 128 == 0, "gemm_glu: needs M > 128 and I %% 128 == 0 (M=%d I=%d K=%d)", M, I, K);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(gu) & 15) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0 && ldgu % 8 == 0 &&
                   ldh % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0,
               "gemm_glu: outputs must be 16B aligned, leading dimensions multiples of 8");
  CUtensorMap tmA, tmB, tmC, tmH;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, G2_BK, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, W, 2 * I, K, ldw, G2_BK, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmC, gu, M, 2 * I, ldgu, 64, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmH, h, M, I, ldh, 64, 32))) return rc;
  Gemm2Params p;
  p.C = reinterpret_cast<__nv_bfloat16*>(gu);
  p.M = M;
  p.N = 2 * I;
  p.K = K;
  p.ldc = ldgu;
  p.accumulate = 0;
  p.group_m = 8;
  p.a_lbo = p.b_lbo = 16;
  p.a_sbo = p.b_sbo = 1024;
  p.scatter_maps = nullptr;
  p.rows_per_owner = 0;
  p.m_rot = 0;
  p.gelu = gelu & 1;
  return launch_gemm2_v<0, 0, G2_MODE_GLU>(tmA, tmB, tmC, tmH, p, stream);
}

// Grouped GEMM for mixture-of-experts blocks (MixtralExperts.forward models/mixtral/modeling_mixtral.py:69-93 /
// grouped_mm_experts_forward integrations/moe.py:377-478): rows [offsets[g], offsets[g+1]) of A [M_total, K] -- the tokens
// routed to expert g, sorted by expert -- are multiplied by expert g's matrix, B + g * N * K ([N, K] K-major; or [K, N] when
// b_mn, the dgrad layout) into the same rows of C [M_total, N].  `offsets` is an int32[groups + 1] array in DEVICE memory
// (written by b200_moe_route): no host synchronisation, one launch for all experts.  N % 64 == 0; b_mn needs K % 64 == 0.
extern "C" int b200_gemm_bf16_grouped(const void* A, const void* B, void* C, const int* offsets, int groups, int M_total, int N,
                                      int K, int lda, int ldb, int ldc, int b_mn, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(groups >= 1 && groups <= G2_MAX_GROUPS, "gemm_grouped: %d groups (1..%d)", groups, G2_MAX_GROUPS);
  B200_REQUIRE(M_total >= 0 && N > 0 && K > 0 && N % 64 == 0 && (!b_mn || K % 64 == 0),
               "gemm_grouped: needs N %% 64 == 0 (and K %% 64 == 0 for the dgrad layout), got M=%d N=%d K=%d", M_total, N, K);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && offsets != nullptr,
               "gemm_grouped: C must be 16B aligned, leading dimensions multiples of 8, offsets non-null");
  if (M_total == 0) return B200_OK;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, M_total, K, lda, G2_BK, 128))) return rc;
  if (!b_mn)
    rc = make_tmap_2d_bf16(&tmB, B, static_cast<uint64_t>(groups) * N, K, ldb, G2_BK, 128);
  else
    rc = make_tmap_2d_bf16(&tmB, B, static_cast<uint64_t>(groups) * K, N, ldb, 64, G2_BK);
  if (rc) return rc;
  if ((rc = make_tmap_2d_bf16(&tmC, C, M_total, N, ldc, 64, 32))) return rc;
  Gemm2Params p;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.M = M_total;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.accumulate = 0;
  p.group_m = 8;
  const uint32_t k_lbo = 16, k_sbo = 1024, mn_lbo = 64 * G2_BK * 2, mn_sbo = 1024;
  p.a_lbo = k_lbo;
  p.a_sbo = k_sbo;
  p.b_lbo = b_mn ? mn_lbo : k_lbo;
  p.b_sbo = b_mn ? mn_sbo : k_sbo;
  p.m_rot = 0;
  p.gelu = 0;
  p.offsets = offsets;
  p.groups = groups;
  // the tile count lives on the device; size the persistent grid for the worst case (every range adds one partial m tile)
  const int max_tiles = ((M_total + G2_BM - 1) / G2_BM + groups) * ((N + G2_BN - 1) / G2_BN);
  return b_mn ? launch_gemm2_v<0, 1, G2_MODE_GROUPED>(tmA, tmB, tmC, tmC, p, stream, max_tiles)
              : launch_gemm2_v<0, 0, G2_MODE_GROUPED>(tmA, tmB, tmC, tmC, p, stream, max_tiles);
}

// GEMM whose epilogue is the first half of a reduce-scatter: D = A B^T as b200_gemm_bf16, but row block r of the output
// (rows [r * M / world, (r + 1) * M / world)) is TMA-stored straight into dest_ptrs[r] -- a [M / world, N] bf16 buffer with
// leading dimension ldc, for r != rank a slot in rank r's peer-mapped memory, so the partial sums cross NVLink tile by tile
// while the tensor cores work on the following tiles (own block last).  The caller then barriers and sums its `world` slots
// (b200_pull_reduce_bf16 on local memory).  dest_ptrs: HOST array; M / world must be a multiple of 256.
extern "C" int b200_gemm_bf16_scatter(const void* A, const void* B, void* const* dest_ptrs, int world, int rank, int M, int N,
                                      int K, int lda, int ldb, int ldc, int a_mn, int b_mn, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "gemm_scatter: bad world=%d rank=%d", world, rank);
  B200_REQUIRE(M > 0 && M % world == 0 && (M / world) % G2_BM == 0, "gemm_scatter: M=%d / world=%d must be a multiple of %d rows",
               M, world, G2_BM);
  const int rows = M / world;
  // tensor maps of the destination buffers live in device memory; they only depend on (pointers, shape), which repeat
  // every other call (double-buffered slots), so they are built and uploaded once per distinct destination set
  struct Entry {
    void* ptr[16];
    int world, rows, N, ldc;
    CUtensorMap* dev;
  };
  static Entry cache[32];
  static int n_cache = 0;
  const CUtensorMap* dev_maps = nullptr;
  for (int i = 0; i < n_cache && !dev_maps; ++i) {
    const Entry& e = cache[i];
    bool same = e.world == world && e.rows == rows && e.N == N && e.ldc == ldc;
    for (int r = 0; same && r < world; ++r) same = e.ptr[r] == dest_ptrs[r];
    if (same) dev_maps = e.dev;
  }
  if (!dev_maps) {
    B200_REQUIRE(n_cache < 32, "gemm_scatter: more than 32 distinct destination sets");
    Entry& e = cache[n_cache];
    CUtensorMap host_maps[16];
    for (int r = 0; r < world; ++r) {
      B200_REQUIRE(dest_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(dest_ptrs[r]) & 15) == 0 && ldc % 8 == 0,
                   "gemm_scatter: destination %d must be 16B aligned, ldc %% 8 == 0", r);
      const int rc = make_tmap_2d_bf16(&host_maps[r], dest_ptrs[r], rows, N, ldc, 64, 32);
      if (rc) return rc;
      e.ptr[r] = dest_ptrs[r];
    }
    e.world = world, e.rows = rows, e.N = N, e.ldc = ldc;
    B200_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&e.dev), sizeof(CUtensorMap) * 16));
    B200_CHECK_CUDA(cudaMemcpy(e.dev, host_maps, sizeof(CUtensorMap) * world, cudaMemcpyHostToDevice));
    dev_maps = e.dev;
    ++n_cache;
  }
  const int m_rot = ((rank + 1) % world) * (rows / G2_BM);
  // C is only used for the (unused) plain output tensor map: point it at the own block
  return gemm2_run(A, B, dest_ptrs[rank], M, N, K, lda, ldb, ldc, a_mn, b_mn, 0, dev_maps, rows, m_rot, stream);
}
