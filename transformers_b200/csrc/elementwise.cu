// HBM-bound kernels of the decoder hot path: embedding gather / scatter-add, RMSNorm fwd/bwd (Llama and Gemma
// variants), RoPE fwd/bwd, SwiGLU / GeGLU fwd/bwd, residual add, causal-LM cross-entropy fwd/bwd.
// All are streaming kernels: 16-byte vectorised coalesced accesses, fp32 math, bf16 I/O, one pass over HBM where the
// algorithm allows.  Rounding points follow the reference eager path.
#ifndef B200_HOST_EMU  // tests/emu executes the kernels below on the host (launchers and inline PTX excluded)
#include "common.cuh"
#endif

#include <cuda_bf16.h>
#include <math.h>

#include "act.cuh"

namespace b200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------ embedding
// reference: nn.Embedding in LlamaModel.forward (models/llama/modeling_llama.py:381); Gemma2 scaled variant
// (models/gemma2/modeling_gemma2.py:338-349) multiplies by bf16(sqrt(hidden)) in bf16.
__global__ void embedding_fwd_kernel(const int64_t* __restrict__ ids, const uint4* __restrict__ W, uint4* __restrict__ out,
                                     int T, int H8, int V, float scale, int has_scale, int* __restrict__ err) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= T) return;
  const int lane = threadIdx.x & 31;
  const int64_t id = ids[row];
  if (id < 0 || id >= V) {
    if (lane == 0) atomicExch(err, 1);
    return;
  }
  const uint4* src = W + static_cast<size_t>(id) * H8;
  uint4* dst = out + static_cast<size_t>(row) * H8;
  for (int c = lane; c < H8; c += 32) {
    uint4 v = __ldg(src + c);
    if (has_scale) {
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= scale;
      v = pack8(f);
    }
    dst[c] = v;
  }
}

// dW[ids[t], :] += dY[t, :] (* scale); bf16x2 atomics (duplicates are rare for LM batches).
__global__ void embedding_bwd_kernel(const int64_t* __restrict__ ids, const __nv_bfloat162* __restrict__ dY,
                                     __nv_bfloat162* __restrict__ dW, int T, int H2, int V, int64_t padding_idx,
                                     float scale, int has_scale) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= T) return;
  const int lane = threadIdx.x & 31;
  const int64_t id = ids[row];
  if (id < 0 || id >= V || id == padding_idx) return;
  const __nv_bfloat162* src = dY + static_cast<size_t>(row) * H2;
  __nv_bfloat162* dst = dW + static_cast<size_t>(id) * H2;
  for (int c = lane; c < H2; c += 32) {
    __nv_bfloat162 v = src[c];
    if (has_scale) {
      float2 f = __bfloat1622float2(v);
      v = __floats2bfloat162_rn(f.x * scale, f.y * scale);
    }
    atomicAdd(dst + c, v);
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// reference: LlamaRMSNorm.forward models/llama/modeling_llama.py:62-67  -> y = w * bf16(x32 * rsqrt(mean(x32^2)+eps))
//            Gemma2RMSNorm       models/gemma2/modeling_gemma2.py:55-63 -> y = bf16(x32 * rsqrt(..) * (1 + f32(w)))
// One warp per row; the row stays in registers between the reduction and the scaling pass (H <= 8192), so HBM sees one
// read and one write.  Optional fused residual add: r = bf16(x + res) is normalised and also written out.
constexpr int NORM_MAX_CHUNKS = 32;  // 32 chunks * 32 lanes * 8 elems = 8192

template <bool GEMMA, bool ADD, int NCH>
__global__ void __launch_bounds__(128)
rmsnorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res_in, const uint4* __restrict__ w,
                   uint4* __restrict__ res_out, uint4* __restrict__ y, float* __restrict__ rstd_out, int T, int H8,
                   float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  const int lane = threadIdx.x & 31;
  const size_t base = static_cast<size_t>(row) * H8;
  uint4 xv[NCH];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < H8) {
      uint4 v = x[base + c];
      float f[8];
      unpack8(v, f);
      if (ADD) {
        float g[8];
        unpack8(res_in[base + c], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = bf16_round(f[e] + g[e]);
        v = pack8(f);
        res_out[base + c] = v;
      }
      xv[i] = v;
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / static_cast<float>(H8 * 8) + eps);
  if (lane == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < H8) {
      float f[8], wf[8];
      unpack8(xv[i], f);
      unpack8(__ldg(w + c), wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (GEMMA)
          f[e] = f[e] * rstd * (1.0f + wf[e]);
        else
          f[e] = wf[e] * bf16_round(f[e] * rstd);
      }
      y[base + c] = pack8(f);
    }
  }
}

// Backward.  Each thread owns 8 consecutive columns (blockDim = H/8 rounded up to a warp multiple); the block walks rows
// ROWS at a time, reducing sum(g * xhat) per row across the block, and keeps its dW partial sums in registers.
//   g = dy * w (Llama) or dy * (1 + w) (Gemma);  xhat = x * rstd;  dx = rstd * (g - xhat * mean(g * xhat))
//   dw = sum_t dy * bf16(xhat) (Llama) / dy * xhat (Gemma)      -> fp32 partials [gridDim.x, H], reduced by a 2nd kernel
template <bool GEMMA, int ROWS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1024 / MAXT)
rmsnorm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ w,
                                   const float* __restrict__ rstd, uint4* __restrict__ dx, float* __restrict__ dw_partial,
                                   int T, int H8) {
  __shared__ float red[32][ROWS];
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const bool active = tid < H8;
  float wf[8];
  if (active) {
    unpack8(__ldg(w + tid), wf);
    if (GEMMA) {
#pragma unroll
      for (int e = 0; e < 8; ++e) wf[e] += 1.0f;
    }
  }
  float dwacc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dwacc[e] = 0.f;
  const float inv_h = 1.0f / static_cast<float>(H8 * 8);

  // software pipeline: the next iteration's rows are requested before this iteration's block reduction
  uint4 dyv[ROWS], xv[ROWS], dyn[ROWS], xn[ROWS];
  float rsv[ROWS], rsn[ROWS];
  auto load_rows = [&](int r0, uint4 (&dd)[ROWS], uint4 (&xx)[ROWS], float (&rr)[ROWS]) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      if (active && row < T) {
        dd[r] = __ldcs(dy + static_cast<size_t>(row) * H8 + tid);
        xx[r] = __ldcs(x + static_cast<size_t>(row) * H8 + tid);
        rr[r] = rstd[row];
      }
    }
  };
  load_rows(blockIdx.x * ROWS, dyv, xv, rsv);
  for (int r0 = blockIdx.x * ROWS; r0 < T; r0 += gridDim.x * ROWS) {
    const int rnext = r0 + gridDim.x * ROWS;
    if (rnext < T) load_rows(rnext, dyn, xn, rsn);
    float g[ROWS][8], xh[ROWS][8], part[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      part[r] = 0.f;
      const int row = r0 + r;
      if (active && row < T) {
        float dyf[8], xf[8];
        unpack8(dyv[r], dyf);
        unpack8(xv[r], xf);
        const float rs = rsv[r];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[r][e] = xf[e] * rs;
          g[r][e] = dyf[e] * wf[e];
          part[r] += g[r][e] * xh[r][e];
          dwacc[e] += dyf[e] * (GEMMA ? xh[r][e] : bf16_round(xh[r][e]));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) g[r][e] = xh[r][e] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) part[r] = warp_sum(part[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) red[wid][r] = part[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float s = 0.f;
      for (int k = 0; k < nwarps; ++k) s += red[k][r];
      part[r] = s * inv_h;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      if (active && row < T) {
        const float rs = rsv[r];
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (g[r][e] - xh[r][e] * part[r]);
        dx[static_cast<size_t>(row) * H8 + tid] = pack8(o);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      dyv[r] = dyn[r];
      xv[r] = xn[r];
      rsv[r] = rsn[r];
    }
  }
  if (active) {
    float* dst = dw_partial + static_cast<size_t>(blockIdx.x) * H8 * 8 + tid * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[e] = dwacc[e];
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, int nparts,
                                       int H, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[static_cast<size_t>(p) * H + c];
  if (accumulate) s += __bfloat162float(out[c]);
  out[c] = __float2bfloat16_rn(s);
}

// ------------------------------------------------------------------------------------------------ RoPE
// reference: rotate_half / apply_rotary_pos_emb models/llama/modeling_llama.py:130-160 (half-split layout):
//   q' = bf16( bf16(q*cos) + bf16(rot(q)*sin) ),  rot(x) = cat(-x[D/2:], x[:D/2])
// In place on the packed projection buffer qkv[T, n_heads_total*D]; the first n_rot heads (q heads then k heads) are
// rotated.  cos/sin are the bf16 tables the reference's rotary_emb returns, [Bc, S, D] with Bc in {1, B}.
// Backward (BWD): dq = bf16( dq'*cos + rot^T(dq'*sin) ), rot^T(y) = cat(y[D/2:], -y[:D/2]).
template <bool BWD>
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ cos_t,
                            const __nv_bfloat16* __restrict__ sin_t, int T, int S, int n_rot, int D, int row_stride,
                            int cos_batch_stride) {
  // blockIdx.x = token row; threads walk (head, 16-byte chunk of the first half); no per-element integer division
  const int half8 = D / 16;  // uint4 chunks per half
  const int t = blockIdx.x;
  const int b = t / S, s_pos = t - b * S;
  __nv_bfloat16* row = qkv + static_cast<size_t>(t) * row_stride;
  const __nv_bfloat16* cb = cos_t + static_cast<size_t>(b) * cos_batch_stride + static_cast<size_t>(s_pos) * D;
  const __nv_bfloat16* sb = sin_t + static_cast<size_t>(b) * cos_batch_stride + static_cast<size_t>(s_pos) * D;
  for (int i = threadIdx.x; i < n_rot * half8; i += blockDim.x) {
    const int h = i / half8, c = i - h * half8;
    __nv_bfloat16* base = row + h * D;
    uint4* p1 = reinterpret_cast<uint4*>(base) + c;
    uint4* p2 = reinterpret_cast<uint4*>(base + D / 2) + c;
    float x1[8], x2[8], c1[8], c2[8], s1[8], s2[8], o1[8], o2[8];
    unpack8(*p1, x1);
    unpack8(*p2, x2);
    unpack8(__ldg(reinterpret_cast<const uint4*>(cb) + c), c1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(cb + D / 2) + c), c2);
    unpack8(__ldg(reinterpret_cast<const uint4*>(sb) + c), s1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(sb + D / 2) + c), s2);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!BWD) {
        o1[e] = bf16_round(x1[e] * c1[e]) + bf16_round(-x2[e] * s1[e]);
        o2[e] = bf16_round(x2[e] * c2[e]) + bf16_round(x1[e] * s2[e]);
      } else {
        o1[e] = x1[e] * c1[e] + x2[e] * s2[e];
        o2[e] = x2[e] * c2[e] - x1[e] * s1[e];
      }
    }
    *p1 = pack8(o1);
    *p2 = pack8(o2);
  }
}

// ------------------------------------------------------------------------------------------------ gated MLP activation
// reference: LlamaMLP.forward models/llama/modeling_llama.py:174-176: h = bf16( bf16(act(g)) * u )
//   act = silu (activations.py:92-103) or gelu(approximate="tanh") (activations.py:30-49, Gemma)
// The activation itself lives in act.cuh (shared with the GEMM's GLU epilogue).  The first version used IEEE division and
// recomputed the exponential for the gradient, which made glu_bwd ALU-bound (~50 instructions per element, 0.46 ms of pure
// issue per call at the Llama-3-8B shape).
// Column layout: plain (gate and up are separate [T, I] matrices with a common row pitch) or BLOCK-INTERLEAVED
// (ilv = 1: one [T, 2I] matrix whose 256-column groups hold 128 gate columns followed by the matching 128 up columns -- the
// layout the GLU-epilogue GEMM produces, so that a 256-wide output tile carries both halves of its columns).
// 2-D launch: blockIdx.y walks GLU_ROWS token rows, threads walk 16-byte column chunks -> no integer division per element
// (a 64-bit div/mod per chunk made the first version ALU-bound: ncu sm__throughput 67 % at 4 TB/s).
constexpr int GLU_ROWS = 4;

__global__ void __launch_bounds__(256)
glu_fwd_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up,
               __nv_bfloat16* __restrict__ out, int T, int I8, int ld_gu, int ld_out, int gelu, int ilv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * GLU_ROWS;
  if (c >= I8) return;
  const int cg = ilv ? ((c >> 4) << 5) + (c & 15) : c;  // 16-byte chunk index of gate / up column chunk c (16 chunks = 128 columns)
  uint4 gv[GLU_ROWS], uv[GLU_ROWS];
#pragma unroll
  for (int j = 0; j < GLU_ROWS; ++j) {
    const int t = t0 + j;
    if (t < T) {
      gv[j] = __ldg(reinterpret_cast<const uint4*>(gate + static_cast<size_t>(t) * ld_gu) + cg);
      uv[j] = __ldg(reinterpret_cast<const uint4*>(up + static_cast<size_t>(t) * ld_gu) + cg);
    }
  }
#pragma unroll
  for (int j = 0; j < GLU_ROWS; ++j) {
    const int t = t0 + j;
    if (t < T) {
      float g[8], u[8], o[8];
      unpack8(gv[j], g);
      unpack8(uv[j], u);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = glu_value(g[e], u[e], gelu);
      *(reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * ld_out) + c) = pack8(o);
    }
  }
}

// dgate, dup written to dgu (same layout as gate/up).
constexpr int GLU_BWD_ROWS = 2;
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ gate,
               const __nv_bfloat16* __restrict__ up, __nv_bfloat16* __restrict__ dgate,
               __nv_bfloat16* __restrict__ dup, int T, int I8, int ld_dh, int ld_gu, int ld_dgu, int gelu, int ilv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * GLU_BWD_ROWS;
  if (c >= I8) return;
  const int cg = ilv ? ((c >> 4) << 5) + (c & 15) : c;
  uint4 dv[GLU_BWD_ROWS], gv[GLU_BWD_ROWS], uv[GLU_BWD_ROWS];
#pragma unroll
  for (int j = 0; j < GLU_BWD_ROWS; ++j) {
    const int t = t0 + j;
    if (t < T) {
      dv[j] = __ldg(reinterpret_cast<const uint4*>(dh + static_cast<size_t>(t) * ld_dh) + c);
      gv[j] = __ldg(reinterpret_cast<const uint4*>(gate + static_cast<size_t>(t) * ld_gu) + cg);
      uv[j] = __ldg(reinterpret_cast<const uint4*>(up + static_cast<size_t>(t) * ld_gu) + cg);
    }
  }
#pragma unroll
  for (int j = 0; j < GLU_BWD_ROWS; ++j) {
    const int t = t0 + j;
    if (t < T) {
      float d[8], g[8], u[8], dg[8], du[8];
      unpack8(dv[j], d);
      unpack8(gv[j], g);
      unpack8(uv[j], u);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a, da;
        act_fwd_grad(g[e], gelu, a, da);
        du[e] = d[e] * bf16_round(a);
        dg[e] = bf16_round(d[e] * u[e]) * da;
      }
      *(reinterpret_cast<uint4*>(dgate + static_cast<size_t>(t) * ld_dgu) + cg) = pack8(dg);
      *(reinterpret_cast<uint4*>(dup + static_cast<size_t>(t) * ld_dgu) + cg) = pack8(du);
    }
  }
}

// ------------------------------------------------------------------------------------------------ residual add
__global__ void add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, size_t n8) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float x[8], y[8];
  unpack8(a[idx], x);
  unpack8(b[idx], y);
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] += y[e];
  out[idx] = pack8(x);
}

// ------------------------------------------------------------------------------------------------ causal-LM loss
// reference: ForCausalLMLoss loss/loss_utils.py:48-70 + fixed_cross_entropy :32-45: logits.float(), labels shifted left
// by one inside each sequence (last position -> ignore_index), mean over valid targets (or sum / num_items).
// One CTA per token row: online log-sum-exp over the vocabulary (one read of the row), loss_row = lse - logit[target].
__global__ void __launch_bounds__(1024)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ lse_out,
              float* __restrict__ loss_rows, int T, int S, int V, int ld, int shift, int64_t ignore_index) {
  __shared__ float sm[32], ss[32];
  const int row = blockIdx.x;
  const __nv_bfloat16* lrow = logits + static_cast<size_t>(row) * ld;
  float m = -INFINITY, s = 0.f;
  const int V8 = V / 8;
  for (int c = threadIdx.x; c < V8; c += blockDim.x) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(lrow) + c), f);
    float cm = f[0];
#pragma unroll
    for (int e = 0; e < 8; ++e) cm = fmaxf(cm, f[e]);
    const float nm = fmaxf(m, cm);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += __expf(f[e] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  for (int c = V8 * 8 + threadIdx.x; c < V; c += blockDim.x) {
    const float f = __bfloat162float(lrow[c]);
    const float nm = fmaxf(m, f);
    s = s * __expf(m - nm) + __expf(f - nm);
    m = nm;
  }
  // block reduce (m, s)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float wm = warp_max(m);
  float wsum = warp_sum(m == -INFINITY ? 0.f : s * __expf(m - wm));
  if (lane == 0) {
    sm[wid] = wm;
    ss[wid] = wsum;
  }
  __syncthreads();
  if (wid == 0) {
    float mm = lane < nw ? sm[lane] : -INFINITY;
    float s2 = lane < nw ? ss[lane] : 0.f;
    const float gm = warp_max(mm);
    const float gs = warp_sum(mm == -INFINITY ? 0.f : s2 * __expf(mm - gm));
    if (lane == 0) {
      const float lse = gm + logf(gs);
      lse_out[row] = lse;
      const int spos = row % S;
      int64_t tgt = ignore_index;
      if (shift) {
        if (spos + 1 < S) tgt = labels[row + 1];
      } else {
        tgt = labels[row];
      }
      float l = 0.f;
      if (tgt != ignore_index && tgt >= 0 && tgt < V) {
        l = lse - __bfloat162float(lrow[tgt]);
      }
      loss_rows[row] = l;
    }
  }
}

// loss = sum(loss_rows) / denom, denom = #valid targets (or num_items when > 0).  Single block, deterministic order.
__global__ void ce_reduce_kernel(const float* __restrict__ loss_rows, const int64_t* __restrict__ labels,
                                 float* __restrict__ loss_out, float* __restrict__ denom_out, int T, int S, int shift,
                                 int64_t ignore_index, float num_items) {
  __shared__ float s1[32], s2[32];
  float a = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < T; r += blockDim.x) {
    a += loss_rows[r];
    int64_t tgt = ignore_index;
    if (shift) {
      if ((r % S) + 1 < S) tgt = labels[r + 1];
    } else {
      tgt = labels[r];
    }
    n += (tgt != ignore_index) ? 1.f : 0.f;
  }
  a = warp_sum(a);
  n = warp_sum(n);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s1[wid] = a;
    s2[wid] = n;
  }
  __syncthreads();
  if (wid == 0) {
    a = lane < (blockDim.x >> 5) ? s1[lane] : 0.f;
    n = lane < (blockDim.x >> 5) ? s2[lane] : 0.f;
    a = warp_sum(a);
    n = warp_sum(n);
    if (lane == 0) {
      const float d = num_items > 0.f ? num_items : n;
      *denom_out = d;
      *loss_out = a / d;
    }
  }
}

// dlogits[row, v] = (softmax(row)[v] - [v == target]) * dloss / denom   (zero row when the target is ignored)
__global__ void __launch_bounds__(1024)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ labels,
              const float* __restrict__ lse, const float* __restrict__ dloss, const float* __restrict__ denom,
              __nv_bfloat16* __restrict__ dlogits, int T, int S, int V, int ld, int ld_out, int shift,
              int64_t ignore_index) {
  const int row = blockIdx.x;
  const int spos = row % S;
  int64_t tgt = ignore_index;
  if (shift) {
    if (spos + 1 < S) tgt = labels[row + 1];
  } else {
    tgt = labels[row];
  }
  const bool valid = (tgt != ignore_index && tgt >= 0 && tgt < V);
  const float scale = valid ? (*dloss) / (*denom) : 0.f;
  const float l = lse[row];
  const __nv_bfloat16* lrow = logits + static_cast<size_t>(row) * ld;
  __nv_bfloat16* drow = dlogits + static_cast<size_t>(row) * ld_out;
  const int V8 = V / 8;
  for (int c = threadIdx.x; c < V8; c += blockDim.x) {
    float f[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(lrow) + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float p = __expf(f[e] - l);
      if (c * 8 + e == tgt) p -= 1.f;
      o[e] = p * scale;
    }
    *(reinterpret_cast<uint4*>(drow) + c) = pack8(o);
  }
  for (int c = V8 * 8 + threadIdx.x; c < V; c += blockDim.x) {
    float p = __expf(__bfloat162float(lrow[c]) - l);
    if (c == tgt) p -= 1.f;
    drow[c] = __float2bfloat16_rn(p * scale);
  }
}

static inline int ceil_div(size_t a, size_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

#define B200_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int b200_embedding_fwd(const int64_t* ids, const void* weight, void* out, int T, int H, int V, float scale,
                                  int has_scale, int* err_flag, cudaStream_t stream) {
  B200_REQUIRE(T >= 0 && H > 0 && H % 8 == 0, "embedding_fwd: H=%d must be a positive multiple of 8", H);
  B200_REQUIRE(B200_ALIGNED16(weight) && B200_ALIGNED16(out), "embedding_fwd: pointers must be 16B aligned");
  if (T == 0) return B200_OK;
  embedding_fwd_kernel<<<ceil_div(T, 8), 256, 0, stream>>>(ids, reinterpret_cast<const uint4*>(weight),
                                                           reinterpret_cast<uint4*>(out), T, H / 8, V, scale, has_scale,
                                                           err_flag);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_embedding_bwd(const int64_t* ids, const void* dout, void* dweight, int T, int H, int V,
                                  int64_t padding_idx, float scale, int has_scale, cudaStream_t stream) {
  B200_REQUIRE(T >= 0 && H > 0 && H % 2 == 0, "embedding_bwd: H=%d must be even", H);
  if (T == 0) return B200_OK;
  embedding_bwd_kernel<<<ceil_div(T, 8), 256, 0, stream>>>(ids, reinterpret_cast<const __nv_bfloat162*>(dout),
                                                           reinterpret_cast<__nv_bfloat162*>(dweight), T, H / 2, V,
                                                           padding_idx, scale, has_scale);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// y = rmsnorm(x [+ res_in]) ; when res_in != NULL also writes res_out = bf16(x + res_in).  rstd_out may be NULL.
extern "C" int b200_rmsnorm_fwd(const void* x, const void* res_in, const void* weight, void* res_out, void* y,
                                float* rstd_out, int T, int H, float eps, int gemma, cudaStream_t stream) {
  B200_REQUIRE(H > 0 && H % 8 == 0 && H <= NORM_MAX_CHUNKS * 256, "rmsnorm_fwd: H=%d must be a multiple of 8 and <= %d",
               H, NORM_MAX_CHUNKS * 256);
  B200_REQUIRE(B200_ALIGNED16(x) && B200_ALIGNED16(y) && B200_ALIGNED16(weight), "rmsnorm_fwd: pointers must be 16B aligned");
  if (T == 0) return B200_OK;
  const int grid = ceil_div(T, 4);
  const uint4 *xp = reinterpret_cast<const uint4*>(x), *rp = reinterpret_cast<const uint4*>(res_in),
              *wp = reinterpret_cast<const uint4*>(weight);
  uint4 *rop = reinterpret_cast<uint4*>(res_out), *yp = reinterpret_cast<uint4*>(y);
  if (res_in) B200_REQUIRE(res_out != nullptr, "rmsnorm_fwd: res_out required with res_in");
#define B200_NORM_LAUNCH(G, A, N) \
  rmsnorm_fwd_kernel<G, A, N><<<grid, 128, 0, stream>>>(xp, rp, wp, rop, yp, rstd_out, T, H / 8, eps)
#define B200_NORM_DISPATCH(N)                       \
  do {                                              \
    if (res_in) {                                   \
      if (gemma) B200_NORM_LAUNCH(true, true, N);   \
      else B200_NORM_LAUNCH(false, true, N);        \
    } else {                                        \
      if (gemma) B200_NORM_LAUNCH(true, false, N);  \
      else B200_NORM_LAUNCH(false, false, N);       \
    }                                               \
  } while (0)
  if (H <= 2048) B200_NORM_DISPATCH(8);
  else if (H <= 4096) B200_NORM_DISPATCH(16);
  else B200_NORM_DISPATCH(32);
#undef B200_NORM_DISPATCH
#undef B200_NORM_LAUNCH
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// workspace: fp32 [b200_rmsnorm_bwd_workspace_rows() * H]
extern "C" int b200_rmsnorm_bwd_workspace_rows(void) { return 2 * (num_sms() > 0 ? num_sms() : 148); }

extern "C" int b200_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, void* dx,
                                void* dweight, float* workspace, int T, int H, int gemma, int accumulate_dw,
                                cudaStream_t stream) {
  B200_REQUIRE(H > 0 && H % 8 == 0 && H <= 8192, "rmsnorm_bwd: H=%d must be a multiple of 8 and <= 8192", H);
  if (T == 0) return B200_OK;
  constexpr int ROWS = 1;
  const int threads = ((H / 8 + 31) / 32) * 32;
  int grid = b200_rmsnorm_bwd_workspace_rows();
  if (grid > ceil_div(T, ROWS)) grid = ceil_div(T, ROWS);
  const uint4 *dyp = reinterpret_cast<const uint4*>(dy), *xp = reinterpret_cast<const uint4*>(x),
              *wp = reinterpret_cast<const uint4*>(weight);
  uint4* dxp = reinterpret_cast<uint4*>(dx);
  if (threads <= 512) {
    if (gemma) rmsnorm_bwd_kernel<true, ROWS, 512><<<grid, threads, 0, stream>>>(dyp, xp, wp, rstd, dxp, workspace, T, H / 8);
    else rmsnorm_bwd_kernel<false, ROWS, 512><<<grid, threads, 0, stream>>>(dyp, xp, wp, rstd, dxp, workspace, T, H / 8);
  } else {
    if (gemma) rmsnorm_bwd_kernel<true, ROWS, 1024><<<grid, threads, 0, stream>>>(dyp, xp, wp, rstd, dxp, workspace, T, H / 8);
    else rmsnorm_bwd_kernel<false, ROWS, 1024><<<grid, threads, 0, stream>>>(dyp, xp, wp, rstd, dxp, workspace, T, H / 8);
  }
  B200_CHECK_CUDA(cudaGetLastError());
  reduce_partials_kernel<<<ceil_div(H, 256), 256, 0, stream>>>(workspace, reinterpret_cast<__nv_bfloat16*>(dweight),
                                                               grid, H, accumulate_dw);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// In-place rotary embedding on the first n_rot heads of qkv[T, row_stride]; backward when bwd != 0.
extern "C" int b200_rope(void* qkv, const void* cos_t, const void* sin_t, int B, int S, int n_rot, int D, int row_stride,
                         int cos_batch, int bwd, cudaStream_t stream) {
  B200_REQUIRE(D % 16 == 0 && row_stride % 8 == 0, "rope: head_dim %d must be a multiple of 16", D);
  B200_REQUIRE(cos_batch == 1 || cos_batch == B, "rope: cos/sin batch must be 1 or B");
  const int T = B * S;
  const size_t total = static_cast<size_t>(T) * n_rot * (D / 16);
  if (total == 0) return B200_OK;
  const int cbs = cos_batch == 1 ? 0 : S * D;
  int threads = n_rot * (D / 16);
  threads = threads > 512 ? 512 : ((threads + 31) / 32) * 32;
  if (bwd)
    rope_kernel<true><<<T, threads, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(cos_t),
        reinterpret_cast<const __nv_bfloat16*>(sin_t), T, S, n_rot, D, row_stride, cbs);
  else
    rope_kernel<false><<<T, threads, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(cos_t),
        reinterpret_cast<const __nv_bfloat16*>(sin_t), T, S, n_rot, D, row_stride, cbs);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_glu_fwd(const void* gate, const void* up, void* out, int T, int I, int ld_gu, int ld_out, int gelu,
                            cudaStream_t stream) {
  B200_REQUIRE(I % 8 == 0 && ld_gu % 8 == 0 && ld_out % 8 == 0, "glu_fwd: I=%d must be a multiple of 8", I);
  B200_REQUIRE(!(gelu & 2) || I % 128 == 0, "glu_fwd: the block-interleaved layout needs I=%d to be a multiple of 128", I);
  if (T == 0 || I == 0) return B200_OK;
  glu_fwd_kernel<<<dim3(ceil_div(I / 8, 256), ceil_div(T, GLU_ROWS)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(gate), reinterpret_cast<const __nv_bfloat16*>(up),
      reinterpret_cast<__nv_bfloat16*>(out), T, I / 8, ld_gu, ld_out, gelu & 1, (gelu >> 1) & 1);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_glu_bwd(const void* dh, const void* gate, const void* up, void* dgate, void* dup, int T, int I,
                            int ld_dh, int ld_gu, int ld_dgu, int gelu, cudaStream_t stream) {
  B200_REQUIRE(I % 8 == 0 && ld_gu % 8 == 0 && ld_dh % 8 == 0 && ld_dgu % 8 == 0, "glu_bwd: I=%d must be a multiple of 8", I);
  B200_REQUIRE(!(gelu & 2) || I % 128 == 0, "glu_bwd: the block-interleaved layout needs I=%d to be a multiple of 128", I);
  if (T == 0 || I == 0) return B200_OK;
  glu_bwd_kernel<<<dim3(ceil_div(I / 8, 256), ceil_div(T, GLU_BWD_ROWS)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dh), reinterpret_cast<const __nv_bfloat16*>(gate),
      reinterpret_cast<const __nv_bfloat16*>(up), reinterpret_cast<__nv_bfloat16*>(dgate),
      reinterpret_cast<__nv_bfloat16*>(dup), T, I / 8, ld_dh, ld_gu, ld_dgu, gelu & 1, (gelu >> 1) & 1);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t stream) {
  B200_REQUIRE(n % 8 == 0, "add: n must be a multiple of 8");
  if (n == 0) return B200_OK;
  add_kernel<<<ceil_div(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(a),
                                                       reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out),
                                                       static_cast<size_t>(n / 8));
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

// lse[T], loss_rows[T], loss_out[1], denom_out[1] are fp32 device buffers owned by the caller.
extern "C" int b200_ce_fwd(const void* logits, const int64_t* labels, float* lse, float* loss_rows, float* loss_out,
                           float* denom_out, int B, int S, int V, int ld, int shift, int64_t ignore_index,
                           float num_items, cudaStream_t stream) {
  B200_REQUIRE(ld % 8 == 0 && B200_ALIGNED16(logits), "ce_fwd: logits rows must be 16B aligned");
  const int T = B * S;
  if (T == 0) return B200_OK;
  ce_fwd_kernel<<<T, 1024, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), labels, lse, loss_rows, T, S, V,
                                        ld, shift, ignore_index);
  B200_CHECK_CUDA(cudaGetLastError());
  ce_reduce_kernel<<<1, 1024, 0, stream>>>(loss_rows, labels, loss_out, denom_out, T, S, shift, ignore_index, num_items);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dloss,
                           const float* denom, void* dlogits, int B, int S, int V, int ld, int ld_out, int shift,
                           int64_t ignore_index, cudaStream_t stream) {
  B200_REQUIRE(ld % 8 == 0 && ld_out % 8 == 0, "ce_bwd: rows must be 16B aligned");
  const int T = B * S;
  if (T == 0) return B200_OK;
  ce_bwd_kernel<<<T, 1024, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), labels, lse, dloss, denom,
                                        reinterpret_cast<__nv_bfloat16*>(dlogits), T, S, V, ld, ld_out, shift,
                                        ignore_index);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
