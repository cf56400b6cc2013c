// Flash attention forward on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces the attention core the reference dispatches through its registry
// (eager_attention_forward models/llama/modeling_llama.py:191-213, sdpa_attention_forward
// integrations/sdpa_attention.py:79-170): out = softmax_fp32(q k^T * scaling [softcap] + mask) v, GQA by indexing
// kv_head = q_head / n_rep (never materialising repeat_kv, modeling_llama.py:179-188), causal / sliding-window /
// per-batch kv range masks computed from indices (masking_utils.py:76-101), never materialised.
//
// One CTA = one 128-row query tile of one (batch, q head); two CTAs are co-resident per SM (96 KB smem, 256 TMEM
// columns each) so one CTA's tensor-core work overlaps the other's softmax.  192 threads, warp specialised:
//   warp 0    TMA producer: Q once, then K_j / V_j tiles (128 x D, 128B swizzle) through 4-D tensor maps over the
//             caller's strided [B, S, h, D] storage (no repacking: works on the packed QKV projection buffer)
//   warp 1    MMA issuer:   S = Q K_j^T  (SS, K-major x K-major, fp32 accum in TMEM cols [0,128))
//                           O += P_j V_j (TS: A = P from TMEM (bf16, aliases S cols [0,64)), B = V MN-major in smem)
//   warps 2-5 softmax:      one thread per query row: tcgen05.ld S row -> scale / softcap / mask -> online softmax with
//                           lazy rescaling of O (only when the running max grows by > 2^8) -> P as bf16 back to TMEM
//   epilogue (warps 2-5):   O / l -> bf16 -> global [B, Sq, Hq, D]; LSE (natural log) for the backward pass.
#include "attn_common.cuh"

#include <stdlib.h>

namespace b200 {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 128;
constexpr int ATT_THREADS = 192;  // softmax warps 0-3, TMA warp 4, MMA warp 5 (highest warp id: favoured by the issue arbiter)
constexpr int ATT_TMA_WARP = 4, ATT_MMA_WARP = 5;

struct AttnFwdParams {
  __nv_bfloat16* O;
  float* lse;  // [B, Hq, lse_stride]
  int lse_stride;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  int B, Hq, Hkv, Sq, Skv;
  float scale;    // softmax scaling (head_dim^-0.5)
  float softcap;  // 0 = off
  int causal;
  int window;  // 0 = off; otherwise kv_idx > q_idx - window
  const int* kv_start;  // optional [B]: first valid kv index (left padding)
  const int* kv_end;    // optional [B]: one past last valid kv index (right padding)
};

struct KvRange {
  int lo, hi;        // valid kv index range [lo, hi) for this (batch, q tile)
  int t_lo, t_hi;    // kv tile range [t_lo, t_hi)
};

__device__ __forceinline__ KvRange kv_range(const AttnFwdParams& p, int b, int q0) {
  KvRange r;
  const int off = p.Skv - p.Sq;  // bottom-right aligned causal (q row i sits at kv position i + off)
  int lo = p.kv_start ? p.kv_start[b] : 0;
  int hi = p.kv_end ? p.kv_end[b] : p.Skv;
  hi = min(hi, p.Skv);
  lo = max(lo, 0);
  const int q_last = min(q0 + ATT_BM, p.Sq) - 1 + off;
  if (p.causal) hi = min(hi, q_last + 1);
  if (p.window > 0) lo = max(lo, q0 + off - p.window + 1);
  r.lo = lo;
  r.hi = hi;
  r.t_lo = lo / ATT_BN;
  r.t_hi = hi > lo ? (hi + ATT_BN - 1) / ATT_BN : r.t_lo;
  return r;
}

template <int D, bool SOFTCAP, bool ILP>
__global__ void __launch_bounds__(ATT_THREADS, D <= 128 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, AttnFwdParams p) {
  constexpr int DCH = D / 64;                 // 64-wide (128 B) column chunks per row
  constexpr int TILE_BYTES = 128 * D * 2;     // one 128 x D bf16 tile
  constexpr int CHUNK_BYTES = 128 * 128;      // one 128-row x 64-col chunk
  constexpr uint32_t TMEM_COLS = D <= 128 ? 256 : 512;  // S (128) + O (D) fp32 columns; head_dim 256 (Gemma-2) takes the whole TMEM
  constexpr uint32_t S_COL = 0, P_COL = 0, O_COL = 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heavy (late) query tiles first; neighbouring CTAs are neighbouring heads (GQA groups share K/V through L2)
  const int num_q_tiles = (p.Sq + ATT_BM - 1) / ATT_BM;
  const int bh_count = p.B * p.Hq;
  const int qt = num_q_tiles - 1 - blockIdx.x / bh_count;
  const int bh = blockIdx.x % bh_count;
  const int b = bh / p.Hq;
  const int h = bh % p.Hq;
  const int hkv = h / (p.Hq / p.Hkv);
  const int q0 = qt * ATT_BM;
  const KvRange kr = kv_range(p, b, q0);
  const int n_iter = kr.t_hi - kr.t_lo;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == ATT_MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == ATT_TMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + c * CHUNK_BYTES, &tmQ, q_full, c * 64, q0, h, b);
      for (int it = 0; it < n_iter; ++it) {
        const int kv0 = (kr.t_lo + it) * ATT_BN;
        mbar_wait(k_empty, (it & 1) ^ 1);
        mbar_expect_tx(k_full, TILE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sK + c * CHUNK_BYTES, &tmK, k_full, c * 64, kv0, hkv, b);
        mbar_wait(v_empty, (it & 1) ^ 1);
        mbar_expect_tx(v_full, TILE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sV + c * CHUNK_BYTES, &tmV, v_full, c * 64, kv0, hkv, b);
      }
    }
  } else if (warp == ATT_MMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, D, 0, 1);
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, SWZ_128B), dK = make_smem_desc(smem_u32(sK), 16, 1024, SWZ_128B);
      const uint64_t dV = make_smem_desc(smem_u32(sV), CHUNK_BYTES, 1024, SWZ_128B);  // V: MN-major B (LBO = chunk pitch)
      mbar_wait(q_full, 0);
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(k_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * CHUNK_BYTES + (kk % 4) * 32;
          umma_ss(tmem_base + S_COL, desc_advance(dQ, off), desc_advance(dK, off), idesc_s, kk != 0);
        }
        umma_commit(k_empty);
        umma_commit(s_full);
        mbar_wait(v_full, it & 1);
        mbar_wait(p_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          // V is the MN-major B operand: 64-column chunks CHUNK_BYTES apart (LBO), 8-row groups 1024 B apart (SBO)
          umma_ts(tmem_base + O_COL, tmem_base + P_COL + kk * 8, desc_advance(dV, kk * 2048), idesc_pv, (it | kk) != 0);
        }
        umma_commit(v_empty);
        if (it == n_iter - 1) umma_commit(o_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qrow = q0 + row;
    const int qpos = qrow + (p.Skv - p.Sq);
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    // exponent multiplier: p = exp2(x * c2 - m * c2); with softcap the score is cap * tanh(s * scale / cap)
    const float c2 = SOFTCAP ? p.softcap * LOG2E : p.scale * LOG2E;
    const float pre = SOFTCAP ? p.scale / p.softcap : 1.0f;
    float m_ref = -INFINITY, l = 0.f;

    for (int it = 0; it < n_iter; ++it) {
      const int kv0 = (kr.t_lo + it) * ATT_BN;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float s[ATT_BN];
      if constexpr (ILP) {
        // all four 32-column loads in flight, one wait (the baseline pays the TMEM read latency four times in a row)
        uint32_t r[ATT_BN / 32][32];
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) tmem_ld_32x32b_x32(tlane + S_COL + c * 32, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) s[c * 32 + e] = __uint_as_float(r[c][e]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tlane + S_COL + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) s[c * 32 + e] = __uint_as_float(r[e]);
        }
      }
      if (SOFTCAP) {
#pragma unroll
        for (int e = 0; e < ATT_BN; ++e) s[e] = fast_tanh(s[e] * pre);
      }
      // masking is only evaluated on boundary tiles (block-uniform test)
      const int tile_hi = kv0 + ATT_BN;  // exclusive
      const bool need_mask = (tile_hi > kr.hi) || (kv0 < kr.lo) ||
                             (p.causal && tile_hi - 1 > q0 + (p.Skv - p.Sq)) ||
                             (p.window > 0 && kv0 <= q0 + ATT_BM - 1 + (p.Skv - p.Sq) - p.window);
      if (need_mask) {
        int hi = kr.hi, lo = kr.lo;
        if (p.causal) hi = min(hi, qpos + 1);
        if (p.window > 0) lo = max(lo, qpos - p.window + 1);
#pragma unroll
        for (int e = 0; e < ATT_BN; ++e) {
          const int col = kv0 + e;
          if (col >= hi || col < lo) s[e] = -INFINITY;
        }
      }
      float m_tile;
      if constexpr (ILP) {
        // four independent max chains of depth 32 instead of one of depth 128 (two softmax warps per scheduler cannot
        // hide a 128-deep dependent chain)
        float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int e = 4; e < ATT_BN; e += 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m4[j], s[e + j]);
        }
        m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
        m_tile = s[0];
#pragma unroll
        for (int e = 1; e < ATT_BN; ++e) m_tile = fmaxf(m_tile, s[e]);
      }
      const float m_cand = fmaxf(m_ref, m_tile);
      const bool need = (m_cand - m_ref) * c2 > 8.0f;  // lazy rescale threshold (2^8 headroom in fp32 / bf16 P)
      float alpha = 1.0f;
      if (need) {
        alpha = fast_exp2((m_ref - m_cand) * c2);
        l *= alpha;
        m_ref = m_cand;
      }
      if (it > 0 && __any_sync(0xffffffffu, need)) {
        // S_it complete implies PV_{it-1} complete (in-order tcgen05 pipe, commit covers all earlier MMAs): O is stable
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tlane + O_COL + c * 32, r);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            w0[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            w1[e] = __float_as_uint(__uint_as_float(r[16 + e]) * alpha);
          }
          tmem_st_32x32b_x16(tlane + O_COL + c * 32, w0);
          tmem_st_32x32b_x16(tlane + O_COL + c * 32 + 16, w1);
        }
      }
      const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c2;
      float lsum = 0.f;
      if constexpr (ILP) {
        float la[ATT_BN / 32], lb[ATT_BN / 32];  // eight independent partial row sums (depth 16 each)
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t pk[16];
          la[c] = 0.f;
          lb[c] = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = fast_exp2(fmaf(s[c * 32 + 2 * e], c2, -mc));
            const float p1 = fast_exp2(fmaf(s[c * 32 + 2 * e + 1], c2, -mc));
            la[c] += p0;
            lb[c] += p1;
            pk[e] = pack_bf16(p0, p1);
          }
          tmem_st_32x32b_x16(tlane + P_COL + c * 16, pk);
        }
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) lsum += la[c] + lb[c];
      } else {
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = fast_exp2(fmaf(s[c * 32 + 2 * e], c2, -mc));
            const float p1 = fast_exp2(fmaf(s[c * 32 + 2 * e + 1], c2, -mc));
            lsum += p0 + p1;
            pk[e] = pack_bf16(p0, p1);
          }
          tmem_st_32x32b_x16(tlane + P_COL + c * 16, pk);
        }
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // epilogue
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    if (n_iter > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    if (qrow < p.Sq) {
      if (p.lse) {
        const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c2;
        p.lse[(static_cast<size_t>(b) * p.Hq + h) * p.lse_stride + qrow] =
            l > 0.f ? mc * 0.6931471805599453f + logf(l) : -INFINITY;
      }
    }
    __nv_bfloat16* orow = p.O + b * p.o_batch_stride + static_cast<int64_t>(qrow) * p.o_row_stride + h * p.o_head_stride;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      if (n_iter > 0) {
        tmem_ld_32x32b_x32(tlane + O_COL + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0;
      }
      if (qrow < p.Sq) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[v * 8 + 0]) * inv_l, __uint_as_float(r[v * 8 + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(r[v * 8 + 2]) * inv_l, __uint_as_float(r[v * 8 + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(r[v * 8 + 4]) * inv_l, __uint_as_float(r[v * 8 + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(r[v * 8 + 6]) * inv_l, __uint_as_float(r[v * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int D, bool SOFTCAP, bool ILP>
static int launch_attn_fwd_v(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdParams& p,
                             cudaStream_t stream) {
  auto kern = attn_fwd_kernel<D, SOFTCAP, ILP>;
  constexpr int smem = 3 * 128 * D * 2 + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int num_q_tiles = (p.Sq + ATT_BM - 1) / ATT_BM;
  const int grid = num_q_tiles * p.B * p.Hq;
  kern<<<grid, ATT_THREADS, smem, stream>>>(tq, tk, tv, p);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------ two-tile ping-pong
// attn_fwd2_kernel: one CTA = TWO adjacent 128-row query tiles (A, B) of one (batch, q head), one CTA per SM, all 512 TMEM
// columns: S_A | S_B | O_A | O_B.  The single MMA thread interleaves the two tiles
//     ... PV_A(j)  S_A(j+1)  PV_B(j)  S_B(j+1)  PV_A(j+1) ...
// so the softmax of tile A's next S runs while the tensor pipe works on tile B and vice versa (the one-tile kernel above
// leaves that overlap to chance between two co-resident CTAs and measured 45 % tensor-pipe activity, profiles/README.md).
// K/V tiles are double-buffered and shared by both q tiles (loaded once per CTA: half the K/V smem fills per q row).
//   warps 0-3  softmax of tile A (TMEM lane quarter = warp % 4, one thread per query row)
//   warps 4-7  softmax of tile B
//   warp  8    TMA producer (Q_A, Q_B once; K_j, V_j through two stages)
//   warp  9    MMA issuer
constexpr int ATT2_THREADS = 320;
constexpr int ATT2_TMA_WARP = 8, ATT2_MMA_WARP = 9;

template <int D, bool SOFTCAP>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, AttnFwdParams p) {
  static_assert(D == 64 || D == 128, "two-tile kernel: head_dim 64 or 128");
  constexpr int DCH = D / 64;
  constexpr int TILE_BYTES = 128 * D * 2;
  constexpr int CHUNK_BYTES = 128 * 128;
  constexpr uint32_t S_COL0 = 0, O_COL0 = 256;   // group g: S at g * 128, O at 256 + g * D

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                      // [2][TILE_BYTES]
  uint8_t* sK = smem + 2 * TILE_BYTES;     // [2][TILE_BYTES]
  uint8_t* sV = smem + 4 * TILE_BYTES;     // [2][TILE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per q tile
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* o_full = bars + 13;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_pairs = (p.Sq + 2 * ATT_BM - 1) / (2 * ATT_BM);
  const int bh_count = p.B * p.Hq;
  const int qp = num_pairs - 1 - blockIdx.x / bh_count;   // heavy (late) query tiles first
  const int bh = blockIdx.x % bh_count;
  const int b = bh / p.Hq;
  const int h = bh % p.Hq;
  const int hkv = h / (p.Hq / p.Hkv);
  const int q0 = qp * 2 * ATT_BM;
  // kv tile range of each q tile; tile B may lie entirely beyond Sq (odd number of 128-row tiles)
  KvRange kr[2];
  kr[0] = kv_range(p, b, q0);
  kr[1] = kv_range(p, b, q0 + ATT_BM);
  if (q0 + ATT_BM >= p.Sq) kr[1].t_hi = kr[1].t_lo;
  const int n0 = kr[0].t_hi - kr[0].t_lo, n1 = kr[1].t_hi - kr[1].t_lo;
  int t_lo, t_hi;
  if (n0 > 0 && n1 > 0) {
    t_lo = min(kr[0].t_lo, kr[1].t_lo);
    t_hi = max(kr[0].t_hi, kr[1].t_hi);
  } else if (n0 > 0) {
    t_lo = kr[0].t_lo, t_hi = kr[0].t_hi;
  } else {
    t_lo = kr[1].t_lo, t_hi = kr[1].t_hi;
  }
  const int n_iter = (n0 > 0 || n1 > 0) ? t_hi - t_lo : 0;
  auto active = [&](int g, int i) { return t_lo + i >= kr[g].t_lo && t_lo + i < kr[g].t_hi; };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == ATT2_MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == ATT2_TMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      const int q_tiles = (q0 + ATT_BM < p.Sq) ? 2 : 1;
      mbar_expect_tx(q_full, q_tiles * TILE_BYTES);
      for (int g = 0; g < q_tiles; ++g) {
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + g * TILE_BYTES + c * CHUNK_BYTES, &tmQ, q_full, c * 64, q0 + g * ATT_BM, h, b);
      }
      for (int i = 0; i < n_iter; ++i) {
        const int st = i & 1;
        const int kv0 = (t_lo + i) * ATT_BN;
        mbar_wait(&k_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sK + st * TILE_BYTES + c * CHUNK_BYTES, &tmK, &k_full[st], c * 64, kv0, hkv, b);
        mbar_wait(&v_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sV + st * TILE_BYTES + c * CHUNK_BYTES, &tmV, &v_full[st], c * 64, kv0, hkv, b);
      }
    }
  } else if (warp == ATT2_MMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, D, 0, 1);
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, SWZ_128B), dK = make_smem_desc(smem_u32(sK), 16, 1024, SWZ_128B);
      const uint64_t dV = make_smem_desc(smem_u32(sV), CHUNK_BYTES, 1024, SWZ_128B);  // V: MN-major B (LBO = chunk pitch)
      int cnt[2] = {0, 0};   // PV MMAs issued per q tile (= barrier phase counters of that tile)
      auto issue_s = [&](int g, int i) {   // S_g = Q_g K_i^T
        const int st = i & 1;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk / 4) * CHUNK_BYTES + (kk % 4) * 32;
          umma_ss(tmem_base + S_COL0 + g * 128, desc_advance(dQ, g * TILE_BYTES + off), desc_advance(dK, st * TILE_BYTES + off),
                  idesc_s, kk != 0);
        }
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int i) {  // O_g += P_g V_i
        const int st = i & 1;
        mbar_wait(&p_full[g], cnt[g] & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          umma_ts(tmem_base + O_COL0 + g * D, tmem_base + S_COL0 + g * 128 + kk * 8, desc_advance(dV, st * TILE_BYTES + kk * 2048),
                  idesc_pv, (cnt[g] | kk) != 0);
        }
        ++cnt[g];
        if (cnt[g] == (g == 0 ? n0 : n1)) umma_commit(&o_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (active(0, 0)) issue_s(0, 0);
      if (active(1, 0)) issue_s(1, 0);
      umma_commit(&k_empty[0]);
      for (int i = 0; i < n_iter; ++i) {
        const int st = i & 1;
        const bool more = i + 1 < n_iter;
        mbar_wait(&v_full[st], (i >> 1) & 1);
        if (more) mbar_wait(&k_full[st ^ 1], ((i + 1) >> 1) & 1);
        tc_fence_after();
        if (active(0, i)) issue_pv(0, i);
        if (more && active(0, i + 1)) issue_s(0, i + 1);
        if (active(1, i)) issue_pv(1, i);
        umma_commit(&v_empty[st]);
        if (more) {
          if (active(1, i + 1)) issue_s(1, i + 1);
          umma_commit(&k_empty[st ^ 1]);
        }
      }
    }
  } else {
    const int g = warp >> 2;          // q tile of this warp
    const int q = warp & 3;           // TMEM lane quarter
    const int row = q * 32 + lane;
    const int qt0 = q0 + g * ATT_BM;
    const int qrow = qt0 + row;
    const int qpos = qrow + (p.Skv - p.Sq);
    const KvRange& kg = kr[g];
    const int n_g = g == 0 ? n0 : n1;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_col = S_COL0 + g * 128, o_col = O_COL0 + g * D;
    const float LOG2E = 1.4426950408889634f;
    const float c2 = SOFTCAP ? p.softcap * LOG2E : p.scale * LOG2E;
    const float pre = SOFTCAP ? p.scale / p.softcap : 1.0f;
    float m_ref = -INFINITY, l = 0.f;

    for (int it = 0; it < n_g; ++it) {
      const int kv0 = (kg.t_lo + it) * ATT_BN;
      mbar_wait(&s_full[g], it & 1);
      tc_fence_after();
      float s[ATT_BN];
      {
        uint32_t r[ATT_BN / 32][32];
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) tmem_ld_32x32b_x32(tlane + s_col + c * 32, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) s[c * 32 + e] = __uint_as_float(r[c][e]);
        }
      }
      if (SOFTCAP) {
#pragma unroll
        for (int e = 0; e < ATT_BN; ++e) s[e] = fast_tanh(s[e] * pre);
      }
      const int tile_hi = kv0 + ATT_BN;
      const bool need_mask = (tile_hi > kg.hi) || (kv0 < kg.lo) || (p.causal && tile_hi - 1 > qt0 + (p.Skv - p.Sq)) ||
                             (p.window > 0 && kv0 <= qt0 + ATT_BM - 1 + (p.Skv - p.Sq) - p.window);
      if (need_mask) {
        int hi = kg.hi, lo = kg.lo;
        if (p.causal) hi = min(hi, qpos + 1);
        if (p.window > 0) lo = max(lo, qpos - p.window + 1);
#pragma unroll
        for (int e = 0; e < ATT_BN; ++e) {
          const int col = kv0 + e;
          if (col >= hi || col < lo) s[e] = -INFINITY;
        }
      }
      float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int e = 4; e < ATT_BN; e += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m4[j], s[e + j]);
      }
      const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float m_cand = fmaxf(m_ref, m_tile);
      const bool need = (m_cand - m_ref) * c2 > 8.0f;  // lazy rescale threshold (2^8 headroom in fp32 / bf16 P)
      float alpha = 1.0f;
      if (need) {
        alpha = fast_exp2((m_ref - m_cand) * c2);
        l *= alpha;
        m_ref = m_cand;
      }
      if (it > 0 && __any_sync(0xffffffffu, need)) {
        // S_g(it) complete implies PV_g(it-1) complete (in-order tcgen05 pipe); the next PV_g waits for this warp's arrive
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tlane + o_col + c * 32, r);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            w0[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            w1[e] = __float_as_uint(__uint_as_float(r[16 + e]) * alpha);
          }
          tmem_st_32x32b_x16(tlane + o_col + c * 32, w0);
          tmem_st_32x32b_x16(tlane + o_col + c * 32 + 16, w1);
        }
      }
      const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c2;
      float la[ATT_BN / 32], lb[ATT_BN / 32];
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t pk[16];
        la[c] = 0.f;
        lb[c] = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = fast_exp2(fmaf(s[c * 32 + 2 * e], c2, -mc));
          const float p1 = fast_exp2(fmaf(s[c * 32 + 2 * e + 1], c2, -mc));
          la[c] += p0;
          lb[c] += p1;
          pk[e] = pack_bf16(p0, p1);
        }
        tmem_st_32x32b_x16(tlane + s_col + c * 16, pk);   // P (bf16) aliases the first 64 columns of S_g
      }
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) lsum += la[c] + lb[c];
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
    }

    // epilogue of this q tile
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    if (n_g > 0) {
      mbar_wait(&o_full[g], 0);
      tc_fence_after();
    }
    if (qrow < p.Sq) {
      if (p.lse) {
        const float mc = (m_ref == -INFINITY) ? 0.f : m_ref * c2;
        p.lse[(static_cast<size_t>(b) * p.Hq + h) * p.lse_stride + qrow] = l > 0.f ? mc * 0.6931471805599453f + logf(l) : -INFINITY;
      }
    }
    __nv_bfloat16* orow = p.O + b * p.o_batch_stride + static_cast<int64_t>(qrow) * p.o_row_stride + h * p.o_head_stride;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      if (n_g > 0) {
        tmem_ld_32x32b_x32(tlane + o_col + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0;
      }
      if (qrow < p.Sq) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[v * 8 + 0]) * inv_l, __uint_as_float(r[v * 8 + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(r[v * 8 + 2]) * inv_l, __uint_as_float(r[v * 8 + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(r[v * 8 + 4]) * inv_l, __uint_as_float(r[v * 8 + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(r[v * 8 + 6]) * inv_l, __uint_as_float(r[v * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT2_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D, bool SOFTCAP>
static int launch_attn_fwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdParams& p,
                            cudaStream_t stream) {
  auto kern = attn_fwd2_kernel<D, SOFTCAP>;
  constexpr int smem = 6 * 128 * D * 2 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int num_pairs = (p.Sq + 2 * ATT_BM - 1) / (2 * ATT_BM);
  kern<<<num_pairs * p.B * p.Hq, ATT2_THREADS, smem, stream>>>(tq, tk, tv, p);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

template <int D, bool SOFTCAP>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdParams& p,
                           cudaStream_t stream) {
  // softmax stage with batched TMEM loads and split max / sum dependency chains (measured 0.578 vs 0.587 ms for the
  // one-chain variant at B4 S4096 32/8 heads, profiles/r02_call1_validation.md); the one-chain variant is gone
  return launch_attn_fwd_v<D, SOFTCAP, true>(tq, tk, tv, p, stream);
}

}  // namespace b200

// q [B, Sq, Hq, D], k/v [B, Skv, Hkv, D] and out [B, Sq, Hq, D] as strided views (strides in elements; last dim
// contiguous).  lse: fp32 [B, Hq, lse_stride] or NULL (lse_stride >= Sq).  kv_start / kv_end: optional int32 [B] valid kv ranges (padding).
static int attn_fwd_run(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, int B,
                        int Sq, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_rs, int64_t q_hs, int64_t k_bs,
                        int64_t k_rs, int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs,
                        int64_t o_rs, int64_t o_hs, float scale, float softcap, int causal, int window,
                        const int* kv_start, const int* kv_end, bool one_tile_kernel, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(D == 64 || D == 128 || D == 256, "attn_fwd: head_dim %d not supported (64, 128 or 256)", D);
  B200_REQUIRE(Hkv > 0 && Hq % Hkv == 0, "attn_fwd: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  B200_REQUIRE(o_rs % 8 == 0 && o_hs % 8 == 0 && o_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "attn_fwd: output must be 16B aligned with strides multiple of 8");
  if (B == 0 || Sq == 0) return B200_OK;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap(&tq, q, D, Sq, Hq, B, q_bs, q_rs, q_hs, ATT_BM))) return rc;
  if ((rc = make_qkv_tmap(&tk, k, D, Skv, Hkv, B, k_bs, k_rs, k_hs, ATT_BN))) return rc;
  if ((rc = make_qkv_tmap(&tv, v, D, Skv, Hkv, B, v_bs, v_rs, v_hs, ATT_BN))) return rc;
  AttnFwdParams p;
  p.O = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.lse_stride = lse_stride;
  p.o_batch_stride = o_bs;
  p.o_row_stride = o_rs;
  p.o_head_stride = o_hs;
  p.B = B;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.Sq = Sq;
  p.Skv = Skv;
  p.scale = scale;
  p.softcap = softcap;
  p.causal = causal;
  p.window = window;
  p.kv_start = kv_start;
  p.kv_end = kv_end;
  const bool sc = softcap > 0.f;
  if (D == 256) return sc ? launch_attn_fwd<256, true>(tq, tk, tv, p, stream) : launch_attn_fwd<256, false>(tq, tk, tv, p, stream);
  // more than one 128-row q tile: the two-tile ping-pong kernel (head_dim 256 needs all of TMEM for one tile)
  if (Sq > ATT_BM && !one_tile_kernel) {
    if (D == 128) return sc ? launch_attn_fwd2<128, true>(tq, tk, tv, p, stream) : launch_attn_fwd2<128, false>(tq, tk, tv, p, stream);
    return sc ? launch_attn_fwd2<64, true>(tq, tk, tv, p, stream) : launch_attn_fwd2<64, false>(tq, tk, tv, p, stream);
  }
  if (D == 128) return sc ? launch_attn_fwd<128, true>(tq, tk, tv, p, stream) : launch_attn_fwd<128, false>(tq, tk, tv, p, stream);
  return sc ? launch_attn_fwd<64, true>(tq, tk, tv, p, stream) : launch_attn_fwd<64, false>(tq, tk, tv, p, stream);
}

extern "C" int b200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, int B,
                             int Sq, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_rs, int64_t q_hs, int64_t k_bs,
                             int64_t k_rs, int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs,
                             int64_t o_rs, int64_t o_hs, float scale, float softcap, int causal, int window,
                             const int* kv_start, const int* kv_end, cudaStream_t stream) {
  return attn_fwd_run(q, k, v, out, lse, lse_stride, B, Sq, Skv, Hq, Hkv, D, q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs,
                      o_bs, o_rs, o_hs, scale, softcap, causal, window, kv_start, kv_end, false, stream);
}

// The one-tile kernel (one 128-row q tile per CTA, two CTAs per SM) for any shape: what b200_attn_fwd uses for Sq <= 128 and
// head_dim 256; exported so the two kernels can be timed and cross-checked against each other.
extern "C" int b200_attn_fwd_1tile(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, int B,
                                   int Sq, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_rs, int64_t q_hs,
                                   int64_t k_bs, int64_t k_rs, int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                                   int64_t o_bs, int64_t o_rs, int64_t o_hs, float scale, float softcap, int causal, int window,
                                   const int* kv_start, const int* kv_end, cudaStream_t stream) {
  return attn_fwd_run(q, k, v, out, lse, lse_stride, B, Sq, Skv, Hq, Hkv, D, q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs,
                      o_bs, o_rs, o_hs, scale, softcap, causal, window, kv_start, kv_end, true, stream);
}
