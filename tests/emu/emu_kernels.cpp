// Host harness around the kernels of optim.cu / peer.cu / gemv.cu / attention_decode.cu (see cuda_emu.h): plain C entry points
// on host arrays, loaded with ctypes by tests/test_kernels_emulated_cpu.py.
#define B200_HOST_EMU 1
#include "cuda_emu.h"

#include "optim.cu"
#include "peer.cu"
#include "gemv.cu"
#include "attention_decode.cu"
#include "ce_sharded.cu"

using namespace b200;

extern "C" int emu_adamw(const int64_t* table, const int32_t* chunks, int n_chunks, int state_fp32, float lr, float beta1,
                         float beta2, float eps, float wd, float bc1, float bc2_sqrt, const float* grad_scale) {
  AdamArgs a{lr, beta1, beta2, eps, wd, lr / bc1, 1.f / bc2_sqrt};
  const int2* ch = reinterpret_cast<const int2*>(chunks);
  if (state_fp32)
    emu::launch(dim3(n_chunks), dim3(OPT_THREADS), [&] { adamw_multi_kernel<float>(table, ch, a, grad_scale); });
  else
    emu::launch(dim3(n_chunks), dim3(OPT_THREADS), [&] { adamw_multi_kernel<__nv_bfloat16>(table, ch, a, grad_scale); });
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
