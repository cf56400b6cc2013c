// Decode-step attention (q_len == 1 over the whole KV cache): generate()'s inner loop, LlamaAttention.forward with a cache
// (models/llama/modeling_llama.py:243-281; eager_attention_forward :191-213 on [B, H, 1, D] x [B, Hkv, ctx, D]).
//
// One query row per head against ctx keys is a stream over the cache: 2 * ctx * D * 2 bytes per kv head against
// 4 * ctx * D * G FLOPs (G = q heads per kv head) -- HBM-bound, so CUDA cores, no tensor-core tile (a 128-row q tile would
// be 1/128 full) and, unlike the prefill kernel, the context is SPLIT across CTAs so that B * Hkv * nsplit CTAs cover the
// 148 SMs even at batch 1 (flash-decoding style).  Every K / V byte is read exactly once and shared by the G q heads of
// its kv head (GQA by indexing, repeat_kv :179-188 never materialised).
//
//   kernel 1 (grid nsplit x Hkv x B, 4 warps): each warp walks rows of its split, 256 / D rows at a time; a lane holds one
//            16-byte piece of the row; scores for the G heads by partial dot + group shuffle reduction; online softmax in
//            fp32 (exp2 domain, optional softcap); lane-local slice of the G output accumulators; block-level merge
//            through shared memory; partial (m, l, O[D]) per (b, q head, split) to the workspace.
//   kernel 2 (grid Hq x B): merges the nsplit partials, writes bf16 out and the natural-log lse.
//
// Algorithmic bytes: B * Hkv * ctx * D * 4 (K and V once) + workspace 2 * B * Hq * nsplit * (D + 2) * 4.
#ifndef B200_HOST_EMU
#include "attn_common.cuh"
#endif

namespace b200 {

constexpr int DEC_WARPS = 4;
constexpr int DEC_MAX_G = 8;

struct DecodeParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  float* lse;   // optional [B, Hq, lse_stride] (entry 0)
  float* ws;    // [B, Hq, nsplit, D + 2] fp32: m, l, O[D]
  int B, Skv, Hq, Hkv, nsplit, lse_stride;
  int64_t q_bs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs, o_bs, o_hs;
  float scale, softcap;
  int window;
  const int *kv_start, *kv_end;
};

__device__ __forceinline__ void dec_unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int D, int G>
__global__ void __launch_bounds__(DEC_WARPS * 32) decode_attn_split_kernel(DecodeParams p) {
  constexpr int LPR = D / 8;          // lanes per row (one 16-byte piece each)
  constexpr int RPI = 32 / LPR;       // rows per warp iteration
  constexpr float LOG2E = 1.4426950408889634f;
  __shared__ float q_s[G][D];                               // q * (scale or scale / softcap), fp32
  __shared__ float m_s[DEC_WARPS * RPI][G], l_s[DEC_WARPS * RPI][G];
  __shared__ float o_s[DEC_WARPS * RPI][G][D];              // <= 4 * 1 * 8 * 256 * 4 B = 32 KB (D = 256), 4*4*8*64*4 = 32 KB (D = 64)

  const int split = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, piece = lane % LPR;

  // valid kv range of this batch row: padding range, sliding window of the (single) query at position Skv - 1
  int lo = p.kv_start ? p.kv_start[b] : 0;
  int hi = p.kv_end ? p.kv_end[b] : p.Skv;
  if (hi > p.Skv) hi = p.Skv;
  if (p.window > 0 && lo < p.Skv - p.window) lo = p.Skv - p.window;  // kv_idx > q_pos - window, q_pos = Skv - 1
  if (lo < 0) lo = 0;
  const int span = hi > lo ? hi - lo : 0;
  const int per = (span + p.nsplit - 1) / p.nsplit;
  const int r0 = lo + split * per;
  const int r1 = min(hi, r0 + per);

  const float pre = p.softcap > 0.f ? p.scale / p.softcap : p.scale * LOG2E;  // softcap: s = cap * tanh(q.k * scale / cap)
  for (int i = threadIdx.x; i < G * D; i += blockDim.x) {
    const int g = i / D, d = i % D;
    q_s[g][d] = __bfloat162float(p.q[b * p.q_bs + static_cast<int64_t>(hkv * G + g) * p.q_hs + d]) * pre;
  }
  __syncthreads();

  float m[G], l[G], acc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[g][e] = 0.f;
  }
  const __nv_bfloat16* kbase = p.k + b * p.k_bs + static_cast<int64_t>(hkv) * p.k_hs;
  const __nv_bfloat16* vbase = p.v + b * p.v_bs + static_cast<int64_t>(hkv) * p.v_hs;
  // the trip count is uniform per warp (all 32 lanes execute the shuffles); lane groups whose row falls beyond the split
  // only skip the loads and the state update
  for (int rb = r0 + warp * RPI; rb < r1; rb += DEC_WARPS * RPI) {
    const int r = rb + sub;
    const bool valid = r < r1;
    float kf[8], vf[8];
    if (valid) {
      dec_unpack8(*reinterpret_cast<const uint4*>(kbase + static_cast<int64_t>(r) * p.k_rs + piece * 8), kf);
      dec_unpack8(*reinterpret_cast<const uint4*>(vbase + static_cast<int64_t>(r) * p.v_rs + piece * 8), vf);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) kf[e] = vf[e] = 0.f;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(kf[e], q_s[g][piece * 8 + e], s);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // the LPR lanes of a row are contiguous
      if (valid) {
        if (p.softcap > 0.f) s = p.softcap * LOG2E * fast_tanh(s);
        const float mn = fmaxf(m[g], s);
        const float corr = fast_exp2(m[g] - mn);  // first row: exp2(-inf) = 0
        const float pr = fast_exp2(s - mn);
        l[g] = l[g] * corr + pr;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[g][e] = fmaf(pr, vf[e], acc[g][e] * corr);
        m[g] = mn;
      }
    }
  }
  const int slot = warp * RPI + sub;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (piece == 0) {
      m_s[slot][g] = m[g];
      l_s[slot][g] = l[g];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) o_s[slot][g][piece * 8 + e] = acc[g][e];
  }
  __syncthreads();
  // merge the DEC_WARPS * RPI partial softmax states; thread t handles (g, d) pairs
  constexpr int NS = DEC_WARPS * RPI;
  for (int i = threadIdx.x; i < G * D; i += blockDim.x) {
    const int g = i / D, d = i % D;
    float mm = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) mm = fmaxf(mm, m_s[s2][g]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) {
      const float w = (m_s[s2][g] == -INFINITY) ? 0.f : fast_exp2(m_s[s2][g] - mm);
      ll += l_s[s2][g] * w;
      oo += o_s[s2][g][d] * w;
    }
    float* dst = p.ws + ((static_cast<int64_t>(b) * p.Hq + hkv * G + g) * p.nsplit + split) * (D + 2);
    if (d == 0) {
      dst[0] = mm;
      dst[1] = ll;
    }
    dst[2 + d] = oo;
  }
}

template <int D>
__global__ void decode_attn_combine_kernel(DecodeParams p) {
  const int hq = blockIdx.x, b = blockIdx.y;
  const float* src = p.ws + (static_cast<int64_t>(b) * p.Hq + hq) * p.nsplit * (D + 2);
  float mm = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) mm = fmaxf(mm, src[s * (D + 2)]);
  float ll = 0.f;
  for (int s = 0; s < p.nsplit; ++s) {
    const float ms = src[s * (D + 2)];
    ll += (ms == -INFINITY) ? 0.f : src[s * (D + 2) + 1] * fast_exp2(ms - mm);
  }
  const float inv = ll > 0.f ? 1.f / ll : 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float oo = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float ms = src[s * (D + 2)];
      if (ms != -INFINITY) oo += src[s * (D + 2) + 2 + d] * fast_exp2(ms - mm);
    }
    p.out[b * p.o_bs + static_cast<int64_t>(hq) * p.o_hs + d] = __float2bfloat16_rn(oo * inv);
  }
  if (threadIdx.x == 0 && p.lse)
    p.lse[(static_cast<int64_t>(b) * p.Hq + hq) * p.lse_stride] = ll > 0.f ? mm * 0.6931471805599453f + logf(ll) : -INFINITY;
}

#ifndef B200_HOST_EMU
template <int D, int G>
static int launch_decode(const DecodeParams& p, cudaStream_t stream) {
  decode_attn_split_kernel<D, G><<<dim3(p.nsplit, p.Hkv, p.B), DEC_WARPS * 32, 0, stream>>>(p);
  B200_CHECK_CUDA(cudaGetLastError());
  decode_attn_combine_kernel<D><<<dim3(p.Hq, p.B), D < 128 ? D : 128, 0, stream>>>(p);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

template <int D>
static int dispatch_g(const DecodeParams& p, int G, cudaStream_t stream) {
  switch (G) {
    case 1: return launch_decode<D, 1>(p, stream);
    case 2: return launch_decode<D, 2>(p, stream);
    case 4: return launch_decode<D, 4>(p, stream);
    case 8: return launch_decode<D, 8>(p, stream);
    default:
      set_last_error("attn_decode: %d q heads per kv head not instantiated (1, 2, 4, 8)", G);
      return B200_ERR_INVALID;
  }
}

#endif  // B200_HOST_EMU

}  // namespace b200

// number of context splits b200_attn_decode will use (workspace = B * Hq * nsplit * (D + 2) floats)
extern "C" int b200_attn_decode_splits(int B, int Hkv, int Skv) {
  int sms = b200::num_sms();
  if (sms <= 0) sms = 148;
  const int ctas = B * Hkv > 0 ? B * Hkv : 1;
  int n = (2 * sms + ctas - 1) / ctas;
  const int max_by_len = (Skv + 255) / 256;
  if (n > max_by_len) n = max_by_len;
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  return n;
}

#ifndef B200_HOST_EMU
// q [B, 1, Hq, D] (q_bs, q_hs strides), k / v [B, Skv, Hkv, D] strided (batch, row, head), out [B, 1, Hq, D]; lse optional.
extern "C" int b200_attn_decode(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride,
                                float* workspace, int B, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_hs,
                                int64_t k_bs, int64_t k_rs, int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs,
                                int64_t o_bs, int64_t o_hs, float scale, float softcap, int window, const int* kv_start,
                                const int* kv_end, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(D == 64 || D == 128 || D == 256, "attn_decode: head_dim %d not supported (64, 128 or 256)", D);
  B200_REQUIRE(Hkv > 0 && Hq % Hkv == 0, "attn_decode: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  B200_REQUIRE(k_rs % 8 == 0 && v_rs % 8 == 0 && k_hs % 8 == 0 && v_hs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
               "attn_decode: k / v must be 16B aligned with strides multiple of 8");
  if (B == 0 || Skv == 0) return B200_OK;
  DecodeParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.ws = workspace;
  p.B = B;
  p.Skv = Skv;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.nsplit = b200_attn_decode_splits(B, Hkv, Skv);
  p.lse_stride = lse_stride;
  p.q_bs = q_bs, p.q_hs = q_hs, p.k_bs = k_bs, p.k_rs = k_rs, p.k_hs = k_hs, p.v_bs = v_bs, p.v_rs = v_rs, p.v_hs = v_hs;
  p.o_bs = o_bs, p.o_hs = o_hs;
  p.scale = scale;
  p.softcap = softcap;
  p.window = window;
  p.kv_start = kv_start;
  p.kv_end = kv_end;
  const int G = Hq / Hkv;
  if (D == 256) return dispatch_g<256>(p, G, stream);
  if (D == 128) return dispatch_g<128>(p, G, stream);
  return dispatch_g<64>(p, G, stream);
}
#endif  // B200_HOST_EMU
