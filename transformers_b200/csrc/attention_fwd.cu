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

// (Round 2 measured a two-tile variant of this kernel -- one CTA per SM owning two adjacent q tiles, S_A | S_B | O_A | O_B in
// TMEM, the MMA thread interleaving  PV_A(j) S_A(j+1) PV_B(j) S_B(j+1)  so that one tile's softmax always runs under the other
// tile's MMAs: bit-identical results, 0.587 ms against this kernel's 0.574 ms at B4 S4096 32/8 heads and 10 % slower inside the
// power-capped training step (profiles/r02_attn_two_tile_vs_one_tile.txt).  The overlap was never the limit: with one thread
// per row both tiles' softmax phases run at the same time and share the SM's 16 MUFU lanes, ~3800 cycles per tile against
// 1048 cycles of MMA.  The variant is gone; the next lever is the exponential itself.  ex2.approx.ftz.bf16x2 is not it: nvcc 12.9
// lowers it to TWO MUFU.EX2.BF16 (one per half) on sm_100a, so the MUFU load is unchanged; what remains is moving part of the
// exponentials to an FMA-pipe polynomial with packed f32x2 arithmetic.)
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
extern "C" int b200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, int B,
                        int Sq, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_rs, int64_t q_hs, int64_t k_bs,
                        int64_t k_rs, int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs,
                        int64_t o_rs, int64_t o_hs, float scale, float softcap, int causal, int window,
                        const int* kv_start, const int* kv_end, cudaStream_t stream) {
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
  if (D == 128) return sc ? launch_attn_fwd<128, true>(tq, tk, tv, p, stream) : launch_attn_fwd<128, false>(tq, tk, tv, p, stream);
  return sc ? launch_attn_fwd<64, true>(tq, tk, tv, p, stream) : launch_attn_fwd<64, false>(tq, tk, tv, p, stream);
}

