// Host harness around the kernels of optim.cu / peer.cu / gemv.cu / attention_decode.cu (see cuda_emu.h): plain C entry points
// on host arrays, loaded with ctypes by tests/test_kernels_emulated_cpu.py.
#define B200_HOST_EMU 1
#include "cuda_emu.h"

#include "optim.cu"
#include "peer.cu"
#include "gemv.cu"
#include "attention_decode.cu"
#include "ce_sharded.cu"
#include "elementwise.cu"
#include "kvcache.cu"
#include "moe.cu"
#include "rope_table.cu"

using namespace b200;

extern "C" int emu_adamw(const int64_t* table, const int32_t* chunks, int n_chunks, int state_fp32, float lr, float beta1,
                         float beta2, float eps, float wd, float bc1, float bc2_sqrt, const float* grad_scale) {
  AdamArgs a{lr, beta1, beta2, eps, wd, lr / bc1, 1.f / bc2_sqrt};
  const int2* ch = reinterpret_cast<const int2*>(chunks);
  const dim3 g(n_chunks), b(OPT_THREADS);
  switch (state_fp32 & 3) {  // bit 0: fp32 moments, bit 1: fp32 master parameters (like b200_adamw_step)
    case 0: emu::launch(g, b, [&] { adamw_multi_kernel<__nv_bfloat16, false>(table, ch, a, grad_scale); }); break;
    case 1: emu::launch(g, b, [&] { adamw_multi_kernel<float, false>(table, ch, a, grad_scale); }); break;
    case 2: emu::launch(g, b, [&] { adamw_multi_kernel<__nv_bfloat16, true>(table, ch, a, grad_scale); }); break;
    default: emu::launch(g, b, [&] { adamw_multi_kernel<float, true>(table, ch, a, grad_scale); }); break;
  }
  return 0;
}

extern "C" int emu_grad_norm(const int64_t* table, const int32_t* chunks, int n_chunks, float* partial, float max_norm, float* out2) {
  const int2* ch = reinterpret_cast<const int2*>(chunks);
  emu::launch(dim3(n_chunks), dim3(OPT_THREADS), [&] { grad_sq_norm_kernel(table, ch, partial); });
  emu::launch(dim3(1), dim3(1024), [&] { grad_norm_finish_kernel(partial, n_chunks, max_norm, out2); });
  return 0;
}

extern "C" int emu_grad_scale(const int64_t* table, const int32_t* chunks, int n_chunks, const float* coef) {
  const int2* ch = reinterpret_cast<const int2*>(chunks);
  emu::launch(dim3(n_chunks), dim3(OPT_THREADS), [&] { grad_scale_kernel(table, ch, coef); });
  return 0;
}

extern "C" int emu_pull_reduce(const void* const* ptrs, int world, int64_t offset, int64_t n, const void* residual, void* out,
                               int blocks) {
  PeerPtrs src;
  for (int s = 0; s < PEER_MAX_WORLD; ++s) src.p[s] = s < world ? reinterpret_cast<const __nv_bfloat16*>(ptrs[s]) : nullptr;
  const uint4* res = reinterpret_cast<const uint4*>(residual);
  uint4* o = reinterpret_cast<uint4*>(out);
  const int64_t n8 = n / 8;
  auto go = [&](auto kern) { emu::launch(dim3(blocks), dim3(256), [&] { kern(src, offset, n8, res, o); }); };
  switch (world) {
    case 1: go(pull_reduce_kernel<1>); break;
    case 2: go(pull_reduce_kernel<2>); break;
    case 4: go(pull_reduce_kernel<4>); break;
    case 8: go(pull_reduce_kernel<8>); break;
    default: return -22;
  }
  return 0;
}

extern "C" int emu_gemv(const void* x, const void* W, void* y, int M, int N, int K, int ldx, int ldw, int ldy) {
  const auto* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const auto* wp = reinterpret_cast<const __nv_bfloat16*>(W);
  auto* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const dim3 grid((N + GEMV_WARPS - 1) / GEMV_WARPS), block(GEMV_WARPS * 32);
  switch (M) {
    case 1: emu::launch(grid, block, [&] { gemv_bf16_kernel<1>(xp, wp, yp, N, K, ldx, ldw, ldy); }); break;
    case 2: emu::launch(grid, block, [&] { gemv_bf16_kernel<2>(xp, wp, yp, N, K, ldx, ldw, ldy); }); break;
    case 3: emu::launch(grid, block, [&] { gemv_bf16_kernel<3>(xp, wp, yp, N, K, ldx, ldw, ldy); }); break;
    case 4: emu::launch(grid, block, [&] { gemv_bf16_kernel<4>(xp, wp, yp, N, K, ldx, ldw, ldy); }); break;
    default: return -22;
  }
  return 0;
}

template <int D, int G>
static void run_decode(const DecodeParams& p) {
  emu::launch(dim3(p.nsplit, p.Hkv, p.B), dim3(DEC_WARPS * 32), [&] { decode_attn_split_kernel<D, G>(p); });
  emu::launch(dim3(p.Hq, p.B), dim3(D < 128 ? D : 128), [&] { decode_attn_combine_kernel<D>(p); });
}

extern "C" int emu_attn_decode(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, float* ws, int B,
                               int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_hs, int64_t k_bs, int64_t k_rs,
                               int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs, int64_t o_hs, float scale,
                               float softcap, int window, const int* kv_start, const int* kv_end, int nsplit) {
  DecodeParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse, p.ws = ws, p.B = B, p.Skv = Skv, p.Hq = Hq, p.Hkv = Hkv, p.nsplit = nsplit, p.lse_stride = lse_stride;
  p.q_bs = q_bs, p.q_hs = q_hs, p.k_bs = k_bs, p.k_rs = k_rs, p.k_hs = k_hs, p.v_bs = v_bs, p.v_rs = v_rs, p.v_hs = v_hs;
  p.o_bs = o_bs, p.o_hs = o_hs, p.scale = scale, p.softcap = softcap, p.window = window, p.kv_start = kv_start, p.kv_end = kv_end;
  const int G = Hq / Hkv;
#define CASE(DD, GG) \
  if (D == DD && G == GG) { run_decode<DD, GG>(p); return 0; }
  CASE(64, 1) CASE(64, 2) CASE(64, 4) CASE(64, 8) CASE(128, 1) CASE(128, 2) CASE(128, 4) CASE(128, 8) CASE(256, 1) CASE(256, 2)
  CASE(256, 4) CASE(256, 8)
#undef CASE
  return -22;
}

extern "C" int emu_ce_bwd_sharded(const void* logits, const int64_t* target_local, const float* lse, const float* row_scale,
                                  void* dlogits, int T, int V, int ld, int ld_out) {
  emu::launch(dim3(T), dim3(1024), [&] {
    ce_bwd_sharded_kernel(reinterpret_cast<const __nv_bfloat16*>(logits), target_local, lse, row_scale,
                          reinterpret_cast<__nv_bfloat16*>(dlogits), V, ld, ld_out);
  });
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// The GPU-validated CUDA-core kernels (elementwise.cu, kvcache.cu, moe.cu) with the launch configurations of their C-ABI
// launchers restated here (the launchers themselves contain <<<>>> and stay device-only).
static inline int cdiv(size_t a, size_t b) { return static_cast<int>((a + b - 1) / b); }
using bf = __nv_bfloat16;

extern "C" int emu_embedding_fwd(const int64_t* ids, const void* w, void* out, int T, int H, int V, float scale, int has_scale, int* err) {
  emu::launch(dim3(cdiv(T, 8)), dim3(256), [&] {
    embedding_fwd_kernel(ids, reinterpret_cast<const uint4*>(w), reinterpret_cast<uint4*>(out), T, H / 8, V, scale, has_scale, err);
  });
  return 0;
}
extern "C" int emu_embedding_bwd(const int64_t* ids, const void* dout, void* dw, int T, int H, int V, int64_t pad, float scale, int has_scale) {
  emu::launch(dim3(cdiv(T, 8)), dim3(256), [&] {
    embedding_bwd_kernel(ids, reinterpret_cast<const __nv_bfloat162*>(dout), reinterpret_cast<__nv_bfloat162*>(dw), T, H / 2, V, pad, scale, has_scale);
  });
  return 0;
}
extern "C" int emu_rmsnorm_fwd(const void* x, const void* res_in, const void* w, void* res_out, void* y, float* rstd, int T, int H, float eps, int gemma) {
  const uint4 *xp = reinterpret_cast<const uint4*>(x), *rp = reinterpret_cast<const uint4*>(res_in), *wp = reinterpret_cast<const uint4*>(w);
  uint4 *rop = reinterpret_cast<uint4*>(res_out), *yp = reinterpret_cast<uint4*>(y);
  const dim3 grid(cdiv(T, 4)), block(128);
#define NL(G, A, N) emu::launch(grid, block, [&] { rmsnorm_fwd_kernel<G, A, N>(xp, rp, wp, rop, yp, rstd, T, H / 8, eps); })
#define ND(N) do { if (res_in) { if (gemma) NL(true, true, N); else NL(false, true, N); } else { if (gemma) NL(true, false, N); else NL(false, false, N); } } while (0)
  if (H <= 2048) ND(8); else if (H <= 4096) ND(16); else ND(32);
#undef ND
#undef NL
  return 0;
}
extern "C" int emu_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, void* dw, float* ws, int T, int H, int gemma, int acc) {
  const int threads = ((H / 8 + 31) / 32) * 32;
  int grid = 2 * 148;
  if (grid > T) grid = T;
  const uint4 *dyp = reinterpret_cast<const uint4*>(dy), *xp = reinterpret_cast<const uint4*>(x), *wp = reinterpret_cast<const uint4*>(w);
  uint4* dxp = reinterpret_cast<uint4*>(dx);
  if (threads <= 512) {
    if (gemma) emu::launch(dim3(grid), dim3(threads), [&] { rmsnorm_bwd_kernel<true, 1, 512>(dyp, xp, wp, rstd, dxp, ws, T, H / 8); });
    else emu::launch(dim3(grid), dim3(threads), [&] { rmsnorm_bwd_kernel<false, 1, 512>(dyp, xp, wp, rstd, dxp, ws, T, H / 8); });
  } else {
    if (gemma) emu::launch(dim3(grid), dim3(threads), [&] { rmsnorm_bwd_kernel<true, 1, 1024>(dyp, xp, wp, rstd, dxp, ws, T, H / 8); });
    else emu::launch(dim3(grid), dim3(threads), [&] { rmsnorm_bwd_kernel<false, 1, 1024>(dyp, xp, wp, rstd, dxp, ws, T, H / 8); });
  }
  emu::launch(dim3(cdiv(H, 256)), dim3(256), [&] { reduce_partials_kernel(ws, reinterpret_cast<bf*>(dw), grid, H, acc); });
  return 0;
}
extern "C" int emu_rope(void* qkv, const void* c, const void* s, int B, int S, int n_rot, int D, int row_stride, int cos_batch, int bwd) {
  const int T = B * S, cbs = cos_batch == 1 ? 0 : S * D;
  int threads = n_rot * (D / 16);
  threads = threads > 512 ? 512 : ((threads + 31) / 32) * 32;
  auto *q = reinterpret_cast<bf*>(qkv);
  auto *cp = reinterpret_cast<const bf*>(c), *sp = reinterpret_cast<const bf*>(s);
  if (bwd) emu::launch(dim3(T), dim3(threads), [&] { rope_kernel<true>(q, cp, sp, T, S, n_rot, D, row_stride, cbs); });
  else emu::launch(dim3(T), dim3(threads), [&] { rope_kernel<false>(q, cp, sp, T, S, n_rot, D, row_stride, cbs); });
  return 0;
}
extern "C" int emu_rope_table(const float* inv_freq, const int64_t* pos, void* c, void* s, int rows, int D, float scaling) {
  const int half = D / 2;
  emu::launch(dim3(cdiv(rows * half, 256)), dim3(256), [&] {
    rope_table_kernel(inv_freq, pos, reinterpret_cast<bf*>(c), reinterpret_cast<bf*>(s), rows, half, scaling);
  });
  return 0;
}
extern "C" int emu_glu_fwd(const void* g, const void* u, void* out, int T, int I, int ld_gu, int ld_out, int gelu) {
  emu::launch(dim3(cdiv(I / 8, 256), cdiv(T, GLU_ROWS)), dim3(256), [&] {
    glu_fwd_kernel(reinterpret_cast<const bf*>(g), reinterpret_cast<const bf*>(u), reinterpret_cast<bf*>(out), T, I / 8, ld_gu, ld_out, gelu & 1, (gelu >> 1) & 1);
  });
  return 0;
}
extern "C" int emu_glu_bwd(const void* dh, const void* g, const void* u, void* dg, void* du, int T, int I, int ld_dh, int ld_gu, int ld_dgu, int gelu) {
  emu::launch(dim3(cdiv(I / 8, 256), cdiv(T, GLU_BWD_ROWS)), dim3(256), [&] {
    glu_bwd_kernel(reinterpret_cast<const bf*>(dh), reinterpret_cast<const bf*>(g), reinterpret_cast<const bf*>(u), reinterpret_cast<bf*>(dg),
                   reinterpret_cast<bf*>(du), T, I / 8, ld_dh, ld_gu, ld_dgu, gelu & 1, (gelu >> 1) & 1);
  });
  return 0;
}
extern "C" int emu_add(const void* a, const void* b, void* out, int64_t n) {
  emu::launch(dim3(cdiv(n / 8, 256)), dim3(256), [&] {
    add_kernel(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), static_cast<size_t>(n / 8));
  });
  return 0;
}
extern "C" int emu_ce_fwd(const void* logits, const int64_t* labels, float* lse, float* rows, float* loss, float* denom, int B, int S, int V, int ld,
                          int shift, int64_t ignore, float num_items) {
  const int T = B * S;
  emu::launch(dim3(T), dim3(1024), [&] { ce_fwd_kernel(reinterpret_cast<const bf*>(logits), labels, lse, rows, T, S, V, ld, shift, ignore); });
  emu::launch(dim3(1), dim3(1024), [&] { ce_reduce_kernel(rows, labels, loss, denom, T, S, shift, ignore, num_items); });
  return 0;
}
extern "C" int emu_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dloss, const float* denom, void* dl, int B, int S,
                          int V, int ld, int ld_out, int shift, int64_t ignore) {
  const int T = B * S;
  emu::launch(dim3(T), dim3(1024), [&] {
    ce_bwd_kernel(reinterpret_cast<const bf*>(logits), labels, lse, dloss, denom, reinterpret_cast<bf*>(dl), T, S, V, ld, ld_out, shift, ignore);
  });
  return 0;
}
extern "C" int emu_kv_append(const void* kn, const void* vn, void* kc, void* vc, int B, int H, int q, int D, int64_t ks_b, int64_t ks_h, int64_t ks_r,
                             int64_t vs_b, int64_t vs_h, int64_t vs_r, int64_t cs_b, int64_t cs_h, int64_t cs_r, int offset) {
  const int total = B * H * q;
  emu::launch(dim3((total * 32 + 255) / 256), dim3(256), [&] {
    kv_append_kernel(reinterpret_cast<const bf*>(kn), reinterpret_cast<const bf*>(vn), reinterpret_cast<bf*>(kc), reinterpret_cast<bf*>(vc), B, H, q,
                     D / 8, ks_b, ks_h, ks_r, vs_b, vs_h, vs_r, cs_b, cs_h, cs_r, offset);
  });
  return 0;
}
extern "C" int emu_moe_route(const int64_t* idx, int* counts, int* offsets, int* cursor, int* slot, int* tok, int T, int topk, int E) {
  const int n = T * topk;
  int grid = (n + 255) / 256;
  if (grid > 1024) grid = 1024;
  emu::launch(dim3(grid), dim3(256), [&] { moe_count_kernel(idx, counts, n, E); });
  emu::launch(dim3(1), dim3(32), [&] { moe_scan_kernel(counts, offsets, cursor, E); });
  emu::launch(dim3((n + 255) / 256), dim3(256), [&] { moe_scatter_kernel(idx, offsets, cursor, slot, tok, n, topk, E); });
  return 0;
}
extern "C" int emu_moe_gather(const void* x, const int* tok, void* xs, int n, int H) {
  emu::launch(dim3((n * 32 + 255) / 256), dim3(256), [&] { moe_gather_kernel(reinterpret_cast<const uint4*>(x), tok, reinterpret_cast<uint4*>(xs), n, H / 8); });
  return 0;
}
extern "C" int emu_moe_combine(const void* ys, const int* slot, const float* w, void* out, int T, int topk, int H) {
  emu::launch(dim3((T * 32 + 255) / 256), dim3(256), [&] {
    moe_combine_kernel(reinterpret_cast<const uint4*>(ys), slot, w, reinterpret_cast<uint4*>(out), T, topk, H / 8);
  });
  return 0;
}
