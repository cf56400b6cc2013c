// Flash attention backward on tcgen05 / TMEM / TMA for sm_100a (autograd of the reference attention core,
// models/llama/modeling_llama.py:191-213), as two atomics-free, deterministic kernels plus a small preprocess:
//
//   attn_bwd_prep   delta[b,h,i] = sum_d dO*O (fp32), lse2 = lse * log2(e) (+inf for fully masked rows)
//   attn_bwd_dkdv   one CTA per (batch, kv head, 128-row kv tile); loops over the GQA group's q heads and 64-row q tiles:
//                     S^T = K Q^T, dP^T = V dO^T            (SS MMAs, M = kv rows, N = 64 q rows, double-buffered TMEM)
//                     P^T = exp2(S^T c - lse2), dS^T = P^T (dP^T - delta)   (one thread per kv row; written to TMEM, bf16)
//                     dV += P^T dO, dK += dS^T Q             (TS MMAs: A from TMEM, B = dO / Q tiles read MN-major)
//   attn_bwd_dq     one CTA per (batch, q head, 128-row q tile); loops over 64-row kv tiles:
//                     S = Q K^T, dP = dO V^T;  dS = P (dP - delta);  dQ += dS K   (A = dS from TMEM, B = K MN-major)
// Recomputing S / dP in both kernels costs 7 instead of 5 MMAs per tile pair but needs no dQ atomics, no smem
// transposes (the transposed operands are produced directly in TMEM by swapping the MMA roles) and is bit-reproducible.
#include "attn_common.cuh"

namespace b200 {

constexpr int BWD_THREADS = 576;  // 16 compute warps (per TMEM lane quarter: four 16-column quarters of each tile) + TMA warp + MMA warp.
constexpr int BWD_TMA_WARP = 16;  // The single-thread MMA issuer sits in the HIGHEST warp id: the SM sub-partition arbiter favours
constexpr int BWD_MMA_WARP = 17;  // high warp ids, and a starved issuer idles the tensor pipe (measured: 2075 vs 1280 cycles / tile).
constexpr float kLog2e = 1.4426950408889634f;

struct AttnBwdParams {
  AttnMask mask;
  int B, Hq, Hkv;
  float scale, softcap;
  const float* lse2;   // [B, Hq, lse_stride]
  const float* delta;  // [B, Hq, lse_stride]
  int lse_stride;
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int64_t dq_bs, dq_rs, dq_hs, dk_bs, dk_rs, dk_hs, dv_bs, dv_rs, dv_hs;
};

// ------------------------------------------------------------------------------------------------ preprocess
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                     const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ lse2,
                                     int B, int Hq, int Sq, int D, int64_t o_bs, int64_t o_rs, int64_t o_hs,
                                     int64_t do_bs, int64_t do_rs, int64_t do_hs, int lse_stride) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = B * Hq * lse_stride;
  if (warp_global >= total) return;
  const int i = warp_global % lse_stride;
  const int h = (warp_global / lse_stride) % Hq;
  const int b = warp_global / (lse_stride * Hq);
  float acc = 0.f;
  if (i < Sq) {
    const __nv_bfloat16* orow = o + b * o_bs + static_cast<int64_t>(i) * o_rs + h * o_hs;
    const __nv_bfloat16* drow = dout + b * do_bs + static_cast<int64_t>(i) * do_rs + h * do_hs;
    for (int c = lane * 2; c < D; c += 64) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(orow + c));
      const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(drow + c));
      acc += a.x * d.x + a.y * d.y;
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
  if (lane == 0) {
    delta[warp_global] = acc;
    float l = INFINITY;
    if (i < Sq) {
      const float v = lse[(static_cast<size_t>(b) * Hq + h) * lse_stride + i];
      l = (v == -INFINITY) ? INFINITY : v * kLog2e;
    }
    lse2[warp_global] = l;
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
template <int D, bool SOFTCAP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     AttnBwdParams p) {
  constexpr int DCH = D / 64;
  constexpr int KV_TILE = 128, Q_TILE = 64;
  constexpr int KV_BYTES = KV_TILE * D * 2;       // K or V tile
  constexpr int KV_CHUNK = KV_TILE * 128;         // 128 rows x 64 cols
  constexpr int Q_BYTES = Q_TILE * D * 2;         // Q or dO tile
  constexpr int Q_CHUNK = Q_TILE * 128;           // 64 rows x 64 cols
  constexpr uint32_t ST_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 384;  // ST/DP: 2 buffers x 64 cols each
  constexpr int NST = 4;  // smem stages of the streamed Q / dO tiles (TMA latency ~ more than one iteration of MMAs)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;            // [NST][Q_BYTES]
  uint8_t* sdO = sQ + NST * Q_BYTES;      // [NST][Q_BYTES]
  float* sLse = reinterpret_cast<float*>(sdO + NST * Q_BYTES);  // [NST][64]
  float* sDelta = sLse + NST * Q_TILE;                           // [NST][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + NST * Q_TILE);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;            // [NST]
  uint64_t* q_empty = q_full + NST;       // [NST]
  uint64_t* sdp_full = q_empty + NST;     // [2]  S^T / dP^T accumulators ready
  uint64_t* pds_full = sdp_full + 2;      // [2]  P^T / dS^T written to TMEM
  uint64_t* acc_full = pds_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AttnMask& mk = p.mask;
  const int n_rep = p.Hq / p.Hkv;
  const int num_kv_tiles = (mk.Skv + KV_TILE - 1) / KV_TILE;
  // early kv tiles are attended by the most q tiles under a causal mask: schedule them first
  const int bh_count = p.B * p.Hkv;
  const int jt = blockIdx.x / bh_count;
  const int bh = blockIdx.x % bh_count;
  const int hkv = bh % p.Hkv;
  const int b = bh / p.Hkv;
  const int kv0 = jt * KV_TILE;
  const int off = mk.Skv - mk.Sq;
  const int kv_lo = max(mk.kv_start ? mk.kv_start[b] : 0, 0);
  const int kv_hi = min(mk.kv_end ? mk.kv_end[b] : mk.Skv, mk.Skv);
  // q-tile range that can attend to this kv tile
  int q_first = 0, q_last = mk.Sq - 1;
  if (mk.causal) q_first = max(q_first, kv0 - off);
  if (mk.window > 0) q_last = min(q_last, kv0 + KV_TILE - 1 + mk.window - 1 - off);
  const bool any_kv = (kv0 < kv_hi) && (kv0 + KV_TILE > kv_lo);
  const int qt_lo = q_first / Q_TILE;
  const int qt_hi = (any_kv && q_last >= q_first) ? q_last / Q_TILE + 1 : qt_lo;
  const int n_qt = qt_hi - qt_lo;
  const int n_iter = n_qt * n_rep;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pds_full[i], 512);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == BWD_MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == BWD_TMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      mbar_expect_tx(kv_full, 2 * KV_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tma_load_4d(sK + c * KV_CHUNK, &tmK, kv_full, c * 64, kv0, hkv, b);
        tma_load_4d(sV + c * KV_CHUNK, &tmV, kv_full, c * 64, kv0, hkv, b);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % NST;
        const int hq = hkv * n_rep + it / n_qt;
        const int q0 = (qt_lo + it % n_qt) * Q_TILE;
        mbar_wait(&q_empty[st], ((it / NST) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * Q_BYTES + 2 * Q_TILE * 4);
#pragma unroll
        for (int c = 0; c < DCH; ++c) {
          tma_load_4d(sQ + st * Q_BYTES + c * Q_CHUNK, &tmQ, &q_full[st], c * 64, q0, hq, b);
          tma_load_4d(sdO + st * Q_BYTES + c * Q_CHUNK, &tmdO, &q_full[st], c * 64, q0, hq, b);
        }
        const size_t row_base = (static_cast<size_t>(b) * p.Hq + hq) * p.lse_stride + q0;
        bulk_load_1d(sLse + st * Q_TILE, p.lse2 + row_base, Q_TILE * 4, &q_full[st]);
        bulk_load_1d(sDelta + st * Q_TILE, p.delta + row_base, Q_TILE * 4, &q_full[st]);
      }
    }
  } else if (warp == BWD_MMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(KV_TILE, Q_TILE, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(KV_TILE, D, 0, 1);
      // descriptors are built once; per MMA only the start-address field advances (keeps the issuing thread short)
      const uint64_t dK_k = make_smem_desc(smem_u32(sK), 16, 1024, SWZ_128B), dV_k = make_smem_desc(smem_u32(sV), 16, 1024, SWZ_128B);
      const uint64_t dQ_k = make_smem_desc(smem_u32(sQ), 16, 1024, SWZ_128B), ddO_k = make_smem_desc(smem_u32(sdO), 16, 1024, SWZ_128B);
      const uint64_t dQ_mn = make_smem_desc(smem_u32(sQ), Q_CHUNK, 1024, SWZ_128B), ddO_mn = make_smem_desc(smem_u32(sdO), Q_CHUNK, 1024, SWZ_128B);
      auto issue_sdp = [&](int it) {
        const int buf = it & 1;
        const int st = it % NST;
        mbar_wait(&q_full[st], (it / NST) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t oa = (kk / 4) * KV_CHUNK + (kk % 4) * 32;
          const uint32_t ob = st * Q_BYTES + (kk / 4) * Q_CHUNK + (kk % 4) * 32;
          umma_ss(tmem_base + ST_COL + buf * 64, desc_advance(dK_k, oa), desc_advance(dQ_k, ob), idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t oa = (kk / 4) * KV_CHUNK + (kk % 4) * 32;
          const uint32_t ob = st * Q_BYTES + (kk / 4) * Q_CHUNK + (kk % 4) * 32;
          umma_ss(tmem_base + DP_COL + buf * 64, desc_advance(dV_k, oa), desc_advance(ddO_k, ob), idesc_s, kk != 0);
        }
        umma_commit(&sdp_full[buf]);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        const int st = it % NST;
        if (it + 1 < n_iter) issue_sdp(it + 1);
        mbar_wait(&pds_full[buf], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < Q_TILE / 16; ++kk) {
          // dO / Q as MN-major B: 64-col chunks Q_CHUNK apart (LBO), 8-row groups 1024 B apart (SBO), 16 rows per k-step
          umma_ts(tmem_base + DV_COL, tmem_base + ST_COL + buf * 64 + kk * 16,
                  desc_advance(ddO_mn, st * Q_BYTES + kk * 2048), idesc_acc, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < Q_TILE / 16; ++kk) {
          umma_ts(tmem_base + DK_COL, tmem_base + DP_COL + buf * 64 + kk * 16,
                  desc_advance(dQ_mn, st * Q_BYTES + kk * 2048), idesc_acc, (it | kk) != 0);
        }
        umma_commit(&q_empty[st]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int qd = warp & 3;    // TMEM lane quarter: hardware lets warp w touch lanes 32*(w%4).. only
    const int cq = warp >> 2;   // which 16-column quarter of each 64-column tile this warp owns (all 16 warps work on every
                                // tile: the per-tile latency of this stage bounds the pipeline with only two TMEM buffers)
    const int row = qd * 32 + lane;
    const int kvpos = kv0 + row;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const float c2 = SOFTCAP ? p.softcap * kLog2e : p.scale * kLog2e;
    const float pre = SOFTCAP ? p.scale / p.softcap : 1.0f;
    const bool row_valid = kvpos >= kv_lo && kvpos < kv_hi;
    // q rows this kv row is visible to: causal kvpos <= q + off; window kvpos > q + off - window; q < Sq
    const int q_min = row_valid ? (mk.causal ? kvpos - off : 0) : 0x7fffffff;
    const int q_max = min(mk.Sq, mk.window > 0 ? kvpos - off + mk.window : 0x7fffffff);
    const int c0 = cq * 16;

    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const int q0 = (qt_lo + it % n_qt) * Q_TILE;
      const int st = it % NST;
      mbar_wait(&q_full[st], (it / NST) & 1);    // lse2 / delta rows landed (same barrier as Q / dO)
      mbar_wait(&sdp_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t rs[16], rd[16];
      tmem_ld_32x32b_x16(tlane + ST_COL + buf * 64 + c0, rs);
      tmem_ld_32x32b_x16(tlane + DP_COL + buf * 64 + c0, rd);
      // block-uniform: does this (kv tile, q tile) pair need element masks?
      const bool need_mask = (kv0 < kv_lo) || (kv0 + KV_TILE > kv_hi) || (q0 + Q_TILE > mk.Sq) ||
                             (mk.causal && kv0 + KV_TILE - 1 > q0 + off) ||
                             (mk.window > 0 && kv0 <= q0 + Q_TILE - 1 + off - mk.window);
      const float4* lrow4 = reinterpret_cast<const float4*>(sLse + st * Q_TILE + c0);
      const float4* drow4 = reinterpret_cast<const float4*>(sDelta + st * Q_TILE + c0);
      float lv[16], dl[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 a = lrow4[v], d4 = drow4[v];
        lv[v * 4 + 0] = a.x; lv[v * 4 + 1] = a.y; lv[v * 4 + 2] = a.z; lv[v * 4 + 3] = a.w;
        dl[v * 4 + 0] = d4.x; dl[v * 4 + 1] = d4.y; dl[v * 4 + 2] = d4.z; dl[v * 4 + 3] = d4.w;
      }
      tmem_ld_wait();
      float pv[16], dv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float x = __uint_as_float(rs[e]);
        float t = 0.f;
        if (SOFTCAP) {
          t = fast_tanh(x * pre);
          x = t;
        }
        const float pe = fast_exp2(fmaf(x, c2, -lv[e]));
        float de = pe * (__uint_as_float(rd[e]) - dl[e]);
        if (SOFTCAP) de *= (1.0f - t * t);
        pv[e] = pe;
        dv[e] = de;
      }
      if (need_mask) {  // block-uniform branch; inside: selects only (q window [q_min, q_max) visible to this kv row)
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int qi = q0 + c0 + e;
          const bool ok = (qi >= q_min) & (qi < q_max);
          pv[e] = ok ? pv[e] : 0.f;
          dv[e] = ok ? dv[e] : 0.f;
        }
      }
      uint32_t pp[8], pd[8];
#pragma unroll
      for (int e2 = 0; e2 < 8; ++e2) {
        pp[e2] = pack_bf16(pv[2 * e2], pv[2 * e2 + 1]);
        pd[e2] = pack_bf16(dv[2 * e2], dv[2 * e2 + 1]);
      }
      // P^T / dS^T (bf16 pairs) overwrite the first 8 of this warp's own 16 S^T / dP^T columns: k-step cq of the TS MMAs
      tmem_st_32x32b_x8(tlane + ST_COL + buf * 64 + c0, pp);
      tmem_st_32x32b_x8(tlane + DP_COL + buf * 64 + c0, pd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_full[buf]);
    }

    if (n_iter > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const bool store = kvpos < mk.Skv;
    __nv_bfloat16* dvrow = p.dv + b * p.dv_bs + static_cast<int64_t>(kvpos) * p.dv_rs + hkv * p.dv_hs;
    __nv_bfloat16* dkrow = p.dk + b * p.dk_bs + static_cast<int64_t>(kvpos) * p.dk_rs + hkv * p.dk_hs;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float mul = which == 0 ? 1.0f : p.scale;
      __nv_bfloat16* dst = which == 0 ? dvrow : dkrow;
      {
        const int c = cq;   // one 32-column chunk of the D-wide accumulator per warp
        if (c >= D / 32) continue;
        uint32_t r[32];
        if (n_iter > 0) {
          tmem_ld_32x32b_x32(tlane + (which == 0 ? DV_COL : DK_COL) + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0;
        }
        if (store) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[v * 8 + 0]) * mul, __uint_as_float(r[v * 8 + 1]) * mul);
            o.y = pack_bf16(__uint_as_float(r[v * 8 + 2]) * mul, __uint_as_float(r[v * 8 + 3]) * mul);
            o.z = pack_bf16(__uint_as_float(r[v * 8 + 4]) * mul, __uint_as_float(r[v * 8 + 5]) * mul);
            o.w = pack_bf16(__uint_as_float(r[v * 8 + 6]) * mul, __uint_as_float(r[v * 8 + 7]) * mul);
            *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BWD_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ dQ
template <int D, bool SOFTCAP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   AttnBwdParams p) {
  constexpr int DCH = D / 64;
  constexpr int Q_TILE = 128, KV_TILE = 64;
  constexpr int Q_BYTES = Q_TILE * D * 2;
  constexpr int Q_CHUNK = Q_TILE * 128;
  constexpr int KV_BYTES = KV_TILE * D * 2;
  constexpr int KV_CHUNK = KV_TILE * 128;
  constexpr uint32_t S_COL = 0, DP_COL = 128, DQ_COL = 256;
  constexpr int NST = 4;  // smem stages of the streamed K / V tiles

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + Q_BYTES;
  uint8_t* sK = sdO + Q_BYTES;         // [NST][KV_BYTES]
  uint8_t* sV = sK + NST * KV_BYTES;   // [NST][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NST * KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;           // [NST]
  uint64_t* kv_empty = kv_full + NST;     // [NST]
  uint64_t* sdp_full = kv_empty + NST;    // [2]
  uint64_t* ds_full = sdp_full + 2;       // [2]
  uint64_t* acc_full = ds_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AttnMask& mk = p.mask;
  const int num_q_tiles = (mk.Sq + Q_TILE - 1) / Q_TILE;
  const int bh_count = p.B * p.Hq;
  const int qt = num_q_tiles - 1 - blockIdx.x / bh_count;
  const int bh = blockIdx.x % bh_count;
  const int b = bh / p.Hq;
  const int h = bh % p.Hq;
  const int hkv = h / (p.Hq / p.Hkv);
  const int q0 = qt * Q_TILE;
  const int off = mk.Skv - mk.Sq;
  int lo = max(mk.kv_start ? mk.kv_start[b] : 0, 0);
  int hi = min(mk.kv_end ? mk.kv_end[b] : mk.Skv, mk.Skv);
  const int q_last = min(q0 + Q_TILE, mk.Sq) - 1 + off;
  if (mk.causal) hi = min(hi, q_last + 1);
  if (mk.window > 0) lo = max(lo, q0 + off - mk.window + 1);
  const int t_lo = lo / KV_TILE;
  const int t_hi = hi > lo ? (hi + KV_TILE - 1) / KV_TILE : t_lo;
  const int n_iter = t_hi - t_lo;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&ds_full[i], 512);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == BWD_MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == BWD_TMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tma_load_4d(sQ + c * Q_CHUNK, &tmQ, q_full, c * 64, q0, h, b);
        tma_load_4d(sdO + c * Q_CHUNK, &tmdO, q_full, c * 64, q0, h, b);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % NST;
        const int kv0 = (t_lo + it) * KV_TILE;
        mbar_wait(&kv_empty[st], ((it / NST) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * KV_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) {
          tma_load_4d(sK + st * KV_BYTES + c * KV_CHUNK, &tmK, &kv_full[st], c * 64, kv0, hkv, b);
          tma_load_4d(sV + st * KV_BYTES + c * KV_CHUNK, &tmV, &kv_full[st], c * 64, kv0, hkv, b);
        }
      }
    }
  } else if (warp == BWD_MMA_WARP) {
    if (n_iter > 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(Q_TILE, KV_TILE, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(Q_TILE, D, 0, 1);
      const uint64_t dQ_k = make_smem_desc(smem_u32(sQ), 16, 1024, SWZ_128B), ddO_k = make_smem_desc(smem_u32(sdO), 16, 1024, SWZ_128B);
      const uint64_t dK_k = make_smem_desc(smem_u32(sK), 16, 1024, SWZ_128B), dV_k = make_smem_desc(smem_u32(sV), 16, 1024, SWZ_128B);
      const uint64_t dK_mn = make_smem_desc(smem_u32(sK), KV_CHUNK, 1024, SWZ_128B);
      auto issue_sdp = [&](int it) {
        const int buf = it & 1;
        const int st = it % NST;
        mbar_wait(&kv_full[st], (it / NST) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t oa = (kk / 4) * Q_CHUNK + (kk % 4) * 32;
          const uint32_t ob = st * KV_BYTES + (kk / 4) * KV_CHUNK + (kk % 4) * 32;
          umma_ss(tmem_base + S_COL + buf * 64, desc_advance(dQ_k, oa), desc_advance(dK_k, ob), idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t oa = (kk / 4) * Q_CHUNK + (kk % 4) * 32;
          const uint32_t ob = st * KV_BYTES + (kk / 4) * KV_CHUNK + (kk % 4) * 32;
          umma_ss(tmem_base + DP_COL + buf * 64, desc_advance(ddO_k, oa), desc_advance(dV_k, ob), idesc_s, kk != 0);
        }
        umma_commit(&sdp_full[buf]);
      };
      mbar_wait(q_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_iter) issue_sdp(it + 1);
        mbar_wait(&ds_full[buf], (it >> 1) & 1);
        tc_fence_after();
        const int st = it % NST;
#pragma unroll
        for (int kk = 0; kk < KV_TILE / 16; ++kk) {
          umma_ts(tmem_base + DQ_COL, tmem_base + DP_COL + buf * 64 + kk * 16,
                  desc_advance(dK_mn, st * KV_BYTES + kk * 2048), idesc_acc, (it | kk) != 0);
        }
        umma_commit(&kv_empty[st]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int qd = warp & 3;
    const int cq = warp >> 2;   // 16-column quarter of each 64-column tile
    const int row = qd * 32 + lane;
    const int qrow = q0 + row;
    const int qpos = qrow + off;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const float c2 = SOFTCAP ? p.softcap * kLog2e : p.scale * kLog2e;
    const float pre = SOFTCAP ? p.scale / p.softcap : 1.0f;
    float my_lse2 = INFINITY, my_delta = 0.f;
    if (qrow < mk.Sq) {
      const size_t idx = (static_cast<size_t>(b) * p.Hq + h) * p.lse_stride + qrow;
      my_lse2 = p.lse2[idx];
      my_delta = p.delta[idx];
    }
    int r_hi = min(mk.kv_end ? mk.kv_end[b] : mk.Skv, mk.Skv), r_lo = max(mk.kv_start ? mk.kv_start[b] : 0, 0);
    if (mk.causal) r_hi = min(r_hi, qpos + 1);
    if (mk.window > 0) r_lo = max(r_lo, qpos - mk.window + 1);
    const int c0 = cq * 16;

    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const int kv0 = (t_lo + it) * KV_TILE;
      mbar_wait(&sdp_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t rs[16], rd[16];
      tmem_ld_32x32b_x16(tlane + S_COL + buf * 64 + c0, rs);
      tmem_ld_32x32b_x16(tlane + DP_COL + buf * 64 + c0, rd);
      const bool need_mask = (kv0 + KV_TILE > hi) || (kv0 < lo) || (mk.causal && kv0 + KV_TILE - 1 > q0 + off) ||
                             (mk.window > 0 && kv0 <= q0 + Q_TILE - 1 + off - mk.window);
      tmem_ld_wait();
      float dv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float x = __uint_as_float(rs[e]);
        float t = 0.f;
        if (SOFTCAP) {
          t = fast_tanh(x * pre);
          x = t;
        }
        const float pe = fast_exp2(fmaf(x, c2, -my_lse2));
        float de = pe * (__uint_as_float(rd[e]) - my_delta);
        if (SOFTCAP) de *= (1.0f - t * t);
        dv[e] = de;
      }
      if (need_mask) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = kv0 + c0 + e;
          const bool ok = (col >= r_lo) & (col < r_hi);
          dv[e] = ok ? dv[e] : 0.f;
        }
      }
      uint32_t pd[8];
#pragma unroll
      for (int e2 = 0; e2 < 8; ++e2) pd[e2] = pack_bf16(dv[2 * e2], dv[2 * e2 + 1]);
      tmem_st_32x32b_x8(tlane + DP_COL + buf * 64 + c0, pd);   // inside this warp's own dP columns: k-step cq of dQ += dS K
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[buf]);
    }

    if (n_iter > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    __nv_bfloat16* dst = p.dq + b * p.dq_bs + static_cast<int64_t>(qrow) * p.dq_rs + h * p.dq_hs;
    for (int c = cq; c < D / 32; c += 4) {   // one 32-column chunk of dQ per warp
      uint32_t r[32];
      if (n_iter > 0) {
        tmem_ld_32x32b_x32(tlane + DQ_COL + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0;
      }
      if (qrow < mk.Sq) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[v * 8 + 0]) * p.scale, __uint_as_float(r[v * 8 + 1]) * p.scale);
          o.y = pack_bf16(__uint_as_float(r[v * 8 + 2]) * p.scale, __uint_as_float(r[v * 8 + 3]) * p.scale);
          o.z = pack_bf16(__uint_as_float(r[v * 8 + 4]) * p.scale, __uint_as_float(r[v * 8 + 5]) * p.scale);
          o.w = pack_bf16(__uint_as_float(r[v * 8 + 6]) * p.scale, __uint_as_float(r[v * 8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BWD_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D, bool SOFTCAP>
static int launch_bwd(const CUtensorMap& tq64, const CUtensorMap& tk128, const CUtensorMap& tv128,
                      const CUtensorMap& tdo64, const CUtensorMap& tq128, const CUtensorMap& tk64,
                      const CUtensorMap& tv64, const CUtensorMap& tdo128, const AttnBwdParams& p, cudaStream_t stream) {
  {
    auto kern = attn_bwd_dkdv_kernel<D, SOFTCAP>;
    constexpr int smem = 2 * 128 * D * 2 + 8 * 64 * D * 2 + 8 * 64 * 4 + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    const int grid = ((p.mask.Skv + 127) / 128) * p.B * p.Hkv;
    kern<<<grid, BWD_THREADS, smem, stream>>>(tq64, tk128, tv128, tdo64, p);
    B200_CHECK_CUDA(cudaGetLastError());
  }
  {
    auto kern = attn_bwd_dq_kernel<D, SOFTCAP>;
    constexpr int smem = 2 * 128 * D * 2 + 8 * 64 * D * 2 + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    const int grid = ((p.mask.Sq + 127) / 128) * p.B * p.Hq;
    kern<<<grid, BWD_THREADS, smem, stream>>>(tq128, tk64, tv64, tdo128, p);
    B200_CHECK_CUDA(cudaGetLastError());
  }
  return B200_OK;
}

}  // namespace b200

// Workspace: fp32 [2 * B * Hq * lse_stride] (delta, lse2).  lse is the forward's output with the same lse_stride
// (a multiple of 128 and >= Sq).  All tensors are strided [B, S, h, D] views (strides in elements).
extern "C" int b200_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                             const float* lse, void* dq, void* dk, void* dv, float* workspace, int B, int Sq, int Skv,
                             int Hq, int Hkv, int D, int lse_stride, const int64_t* strides /* 8 tensors x (bs, rs, hs):
                             q, k, v, out, dout, dq, dk, dv */, float scale, float softcap, int causal, int window,
                             const int* kv_start, const int* kv_end, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(D == 64 || D == 128, "attn_bwd: head_dim %d not supported (64 or 128)", D);
  B200_REQUIRE(Hkv > 0 && Hq % Hkv == 0, "attn_bwd: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  B200_REQUIRE(lse_stride >= Sq && lse_stride % 128 == 0, "attn_bwd: lse_stride %d must be >= Sq and a multiple of 128", lse_stride);
  if (B == 0 || Sq == 0 || Skv == 0) return B200_OK;
  const int64_t* sq = strides + 0;
  const int64_t* sk = strides + 3;
  const int64_t* sv = strides + 6;
  const int64_t* so = strides + 9;
  const int64_t* sdo = strides + 12;
  const int64_t* sdq = strides + 15;
  const int64_t* sdk = strides + 18;
  const int64_t* sdv = strides + 21;
  float* delta = workspace;
  float* lse2 = workspace + static_cast<size_t>(B) * Hq * lse_stride;
  {
    const int total_warps = B * Hq * lse_stride;
    const int threads = 256;
    const int grid = (total_warps * 32 + threads - 1) / threads;
    attn_bwd_prep_kernel<<<grid, threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta, lse2, B,
        Hq, Sq, D, so[0], so[1], so[2], sdo[0], sdo[1], sdo[2], lse_stride);
    B200_CHECK_CUDA(cudaGetLastError());
  }
  CUtensorMap tq64, tq128, tk64, tk128, tv64, tv128, tdo64, tdo128;
  int rc;
  if ((rc = make_qkv_tmap(&tq64, q, D, Sq, Hq, B, sq[0], sq[1], sq[2], 64))) return rc;
  if ((rc = make_qkv_tmap(&tq128, q, D, Sq, Hq, B, sq[0], sq[1], sq[2], 128))) return rc;
  if ((rc = make_qkv_tmap(&tdo64, dout, D, Sq, Hq, B, sdo[0], sdo[1], sdo[2], 64))) return rc;
  if ((rc = make_qkv_tmap(&tdo128, dout, D, Sq, Hq, B, sdo[0], sdo[1], sdo[2], 128))) return rc;
  if ((rc = make_qkv_tmap(&tk64, k, D, Skv, Hkv, B, sk[0], sk[1], sk[2], 64))) return rc;
  if ((rc = make_qkv_tmap(&tk128, k, D, Skv, Hkv, B, sk[0], sk[1], sk[2], 128))) return rc;
  if ((rc = make_qkv_tmap(&tv64, v, D, Skv, Hkv, B, sv[0], sv[1], sv[2], 64))) return rc;
  if ((rc = make_qkv_tmap(&tv128, v, D, Skv, Hkv, B, sv[0], sv[1], sv[2], 128))) return rc;
  AttnBwdParams p;
  p.mask.Sq = Sq;
  p.mask.Skv = Skv;
  p.mask.causal = causal;
  p.mask.window = window;
  p.mask.kv_start = kv_start;
  p.mask.kv_end = kv_end;
  p.B = B;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.scale = scale;
  p.softcap = softcap;
  p.lse2 = lse2;
  p.delta = delta;
  p.lse_stride = lse_stride;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dq_bs = sdq[0]; p.dq_rs = sdq[1]; p.dq_hs = sdq[2];
  p.dk_bs = sdk[0]; p.dk_rs = sdk[1]; p.dk_hs = sdk[2];
  p.dv_bs = sdv[0]; p.dv_rs = sdv[1]; p.dv_hs = sdv[2];
  const bool sc = softcap > 0.f;
  if (D == 128)
    return sc ? launch_bwd<128, true>(tq64, tk128, tv128, tdo64, tq128, tk64, tv64, tdo128, p, stream)
              : launch_bwd<128, false>(tq64, tk128, tv128, tdo64, tq128, tk64, tv64, tdo128, p, stream);
  return sc ? launch_bwd<64, true>(tq64, tk128, tv128, tdo64, tq128, tk64, tv64, tdo128, p, stream)
            : launch_bwd<64, false>(tq64, tk128, tv128, tdo64, tq128, tk64, tv64, tdo128, p, stream);
}
