// K1, large gate, at ranks above one TMEM bucket (96 < max(r, rg) <= 192) -- the rank the reference's T5 scripts ship
// (README.md:253, scripts/image-text/T5-VL-PET-large.sh:46-57: adapter_down_dim 192 = 4 heads x 48, adapter_gating_down_dim
// 192).  Math: my_transformers/modeling_t5.py:777-824, 359-409 / my_transformers/modeling_bart.py:1145-1155, 1195-1209,
// 1256-1260; oracle/pet_oracle.py gated_pet_fwd / _bwd.
//
// The fused kernels keep A, P (and in the backward dz, dq) for ONE rank bucket in TMEM: 2 x 96 fp32 columns next to the
// packed operands and the U/T chunk fill all 512 columns, so r = 192 does not fit one tile pass.  Both branches are sums
// over rank halves, though:
//     y1 = kappa x2 + alpha (z_a Wu_a^T + z_b Wu_b^T + bu),   T = q_a Gu_a^T + q_b Gu_b^T + gbu
// with z_h = gelu_new(x2 Wd_h^T + bd_h) (the down projection splits by rows, gelu is element-wise), so this path COMPOSES
// the module from the ungated tcgen05 kernels (vlpet_k1_sm100.cu / vlpet_k1_bwd_sm100.cu, GATED = false), one launch per
// rank half and branch, plus one element-wise gate kernel:
//   forward   y1a = k x2 + a(U_a + bu) -> y1 = y1a + a U_b -> ta = T_a + gbu -> T = ta + T_b -> out = x1 + s D(y1 (*|+) sig(T))
//   backward  y1, T again (nothing is saved by the forward) -> dy1 = dh (*G), dT = dh (y1) G (1 - G)          (element-wise)
//             -> dx2 = k dy1 + da_a Wd_a (+ weight gradients of half a) -> dx2 += da_b Wd_b (in place, half b)
//             -> dx1 = dout + dp_a Gd_a -> dx1 += dp_b Gd_b
// y1 and T cross HBM in bf16 between the launches -- exactly what the reference does under torch.autocast, where every
// nn.Linear returns bf16 -- so the parity bar for this path is pinned against the reference run in bf16
// (tests/test_gpu_parity.py::test_k1_wide_rank_*).  Cost: ~16 [M, d] passes forward instead of 3; still 12x faster than the
// CUDA-core path these ranks ran on before.  Column slices of Wu / Gu are copied into contiguous temporaries with
// cudaMemcpy2DAsync (147 KB each); the weight gradients of a half go straight into the column slice of the real dWu / dGu
// (row pitch argument of the weight-gradient GEMM).
#include <cstring>

#include "vlpet_common.cuh"

namespace vlpet {
namespace {

constexpr int HALF_R = 96;      // rank bucket of the fused backward

struct GateArgs {
  int64_t n8;                   // number of 8-element groups (M * d / 8)
  int add_gate;
  float s, inv_keep;
  uint32_t thr16;
  uint64_t seed;
  const uint64_t* seed_dev;
  const __nv_bfloat16 *x1, *y1, *t, *dout;
  __nv_bfloat16 *out, *dy1, *dt;
};

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(u[e] << 16); v[2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t u[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    u[e] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

// out = x1 + s * dropout(y1 (*|+) sigmoid(T)); 8 elements per thread, grid-stride
__global__ void __launch_bounds__(256) wide_gate_fwd_kernel(const GateArgs a) {
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? __ldg(a.seed_dev) : 0ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n8; i += (int64_t)gridDim.x * blockDim.x) {
    float x1[8], y1[8], t[8], m[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.x1) + i), x1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.y1) + i), y1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.t) + i), t);
    drop_scale8(seed, a.thr16, a.inv_keep, i * 8, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = sigmoid_f(t[e]);
      const float h = a.add_gate ? y1[e] + g : y1[e] * g;
      o[e] = x1[e] + a.s * m[e] * h;
    }
    reinterpret_cast<uint4*>(a.out)[i] = pack8(o);
  }
}

// dh = s * mask * dout;  mul gate: dy1 = dh G, dT = dh y1 G (1 - G);  add gate: dy1 = dh, dT = dh G (1 - G)
__global__ void __launch_bounds__(256) wide_gate_bwd_kernel(const GateArgs a) {
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? __ldg(a.seed_dev) : 0ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n8; i += (int64_t)gridDim.x * blockDim.x) {
    float go[8], y1[8], t[8], m[8], dy1[8], dt[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.dout) + i), go);
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.y1) + i), y1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(a.t) + i), t);
    drop_scale8(seed, a.thr16, a.inv_keep, i * 8, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = sigmoid_f(t[e]);
      const float dh = a.s * m[e] * go[e];
      const float gg = g * (1.f - g);
      dy1[e] = a.add_gate ? dh : dh * g;
      dt[e] = a.add_gate ? dh * gg : dh * y1[e] * gg;
    }
    reinterpret_cast<uint4*>(a.dy1)[i] = pack8(dy1);
    reinterpret_cast<uint4*>(a.dt)[i] = pack8(dt);
  }
}

struct WideWs {
  __nv_bfloat16 *ya, *y1, *ta, *t;           // [M, d] each; the backward reuses ya / ta for dy1 / dT
  __nv_bfloat16 *wu[2], *gu[2];              // contiguous column halves of Wu / Gu: [d, r_h]
  __nv_bfloat16* zero_d;                     // [d] zeros: the up-projection bias of a second half
  void* sub;                                 // workspace of the ungated tcgen05 backward
  size_t sub_bytes, bytes;
  int ra, rb, ga, gb;                        // rank halves of the adapter / gate branch (rb, gb may be 0)
};

WideWs carve_wide(const VlpetK1Desc& D, bool bwd, void* ws) {
  WideWs w;
  memset(&w, 0, sizeof(w));
  w.ra = D.r < HALF_R ? D.r : HALF_R; w.rb = D.r - w.ra;
  w.ga = D.rg < HALF_R ? D.rg : HALF_R; w.gb = D.rg - w.ga;
  Arena a(ws, (size_t)-1);
  const size_t md = (size_t)D.M * D.d;
  w.ya = a.take<__nv_bfloat16>(md); w.y1 = a.take<__nv_bfloat16>(md);
  w.ta = a.take<__nv_bfloat16>(md); w.t = a.take<__nv_bfloat16>(md);
  w.wu[0] = a.take<__nv_bfloat16>((size_t)D.d * w.ra); w.wu[1] = a.take<__nv_bfloat16>((size_t)D.d * (w.rb ? w.rb : 8));
  w.gu[0] = a.take<__nv_bfloat16>((size_t)D.d * w.ga); w.gu[1] = a.take<__nv_bfloat16>((size_t)D.d * (w.gb ? w.gb : 8));
  w.zero_d = a.take<__nv_bfloat16>((size_t)D.d);
  if (bwd) {
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = HALF_R; K2.dtype = D.dtype;
    w.sub_bytes = fused_k2_bwd_ws(K2);
    w.sub = a.take<char>(w.sub_bytes);
  }
  w.bytes = a.off;
  return w;
}

VlpetK1Desc ungated(const VlpetK1Desc& D, int r, float alpha, float kappa) {
  VlpetK1Desc K = D;
  K.gate = VLPET_GATE_NONE; K.r = r; K.rg = 0; K.add_gate = 0; K.s = 1.0f; K.alpha = alpha; K.kappa = kappa; K.p_drop = 0.f;
  K.seed = 0; K.seed_dev = nullptr; K.impl = VLPET_IMPL_AUTO;
  return K;
}

// column slice [:, c0 : c0 + n) of a row-major [d, r] bf16 matrix -> contiguous [d, n]
int slice_cols(__nv_bfloat16* dst, const void* src, int d, int r, int c0, int n, cudaStream_t st) {
  VLPET_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)n * 2, static_cast<const __nv_bfloat16*>(src) + c0, (size_t)r * 2, (size_t)n * 2,
                                  (size_t)d, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// res = base_term + alpha * sum over the rank halves of Up_h(gelu_new(Down_h(in))) + bias, through the ungated fused forward:
//   first half : res_a = in_res + 1 * (kap0 * in + alpha (U_a + bias))      (in_res, kap0 chosen by the caller)
//   second half: res   = res_a  + 1 * (0 * in    + alpha (U_b + 0))
int branch_fwd(const VlpetK1Desc& D, int r, int ra, int rb, float alpha, float kap0, const void* in_res, const void* in,
               const void* Wd, const void* bd, const void* Wu, const void* bias, __nv_bfloat16* const* wu_h, const __nv_bfloat16* zero_d,
               __nv_bfloat16* tmp, __nv_bfloat16* res, cudaStream_t st) {
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  const void* wu0 = Wu;
  if (rb) {
    VLPET_TRY(slice_cols(wu_h[0], Wu, D.d, r, 0, ra, st));
    VLPET_TRY(slice_cols(wu_h[1], Wu, D.d, r, ra, rb, st));
    wu0 = wu_h[0];
  }
  P.Wd = Wd; P.bd = bd; P.Wu = wu0; P.bu = bias;
  VLPET_TRY(fused_k1_fwd(ungated(D, ra, alpha, kap0), in_res, in, P, rb ? tmp : res, nullptr, 0, st));
  if (!rb) return 0;
  P.Wd = static_cast<const __nv_bfloat16*>(Wd) + (size_t)ra * D.d; P.bd = static_cast<const __nv_bfloat16*>(bd) + ra;
  P.Wu = wu_h[1]; P.bu = zero_d;
  return fused_k1_fwd(ungated(D, rb, alpha, 0.f), tmp, in, P, res, nullptr, 0, st);
}

// y1 and T of the whole module (shared by forward and backward)
int wide_recompute(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, const WideWs& W, cudaStream_t st) {
  VLPET_CUDA_OK(cudaMemsetAsync(W.zero_d, 0, (size_t)D.d * 2, st));
  // y1 = x2 + ((kappa - 1) x2 + alpha (U + bu))
  VLPET_TRY(branch_fwd(D, D.r, W.ra, W.rb, D.alpha, D.kappa - 1.0f, x2, x2, w.Wd, w.bd, w.Wu, w.bu, W.wu, W.zero_d, W.ya, W.y1, st));
  // T = x1 + (-1 x1 + 1 (T' + gbu)): the ungated kernel always adds its residual input, so it is cancelled (to within one
  // fp32 rounding of |x1|, far below the bf16 storage of T)
  return branch_fwd(D, D.rg, W.ga, W.gb, 1.0f, -1.0f, x1, x1, w.Gd, w.gbd, w.Gu, w.gbu, W.gu, W.zero_d, W.ta, W.t, st);
}

GateArgs gate_args(const VlpetK1Desc& D) {
  GateArgs a;
  memset(&a, 0, sizeof(a));
  a.n8 = D.M * (int64_t)D.d / 8;
  a.add_gate = D.add_gate; a.s = D.s;
  a.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  a.inv_keep = a.thr16 ? 1.0f / (1.0f - (float)a.thr16 / 65536.0f) : 1.0f;
  a.seed = D.seed; a.seed_dev = D.seed_dev;
  return a;
}
unsigned gate_blocks(int64_t n8, int sms) {
  int64_t b = (n8 + 255) / 256;
  const int64_t cap = (int64_t)sms * 8;
  return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// ---- the adapter branch alone at 96 < r <= 192: y1 for the row-wise gate kernels (vlpet_rows.cu, COMPOSED mode) ------------
namespace {
struct AdWs { __nv_bfloat16 *ya, *wu[2], *zero_d; void* sub; size_t sub_bytes, bytes; int ra, rb; };
AdWs carve_adapter(const VlpetK1Desc& D, bool bwd, void* ws) {
  AdWs w;
  memset(&w, 0, sizeof(w));
  w.ra = D.r < HALF_R ? D.r : HALF_R; w.rb = D.r - w.ra;
  Arena a(ws, (size_t)-1);
  w.ya = a.take<__nv_bfloat16>((size_t)D.M * D.d);
  w.wu[0] = a.take<__nv_bfloat16>((size_t)D.d * w.ra); w.wu[1] = a.take<__nv_bfloat16>((size_t)D.d * (w.rb ? w.rb : 8));
  w.zero_d = a.take<__nv_bfloat16>((size_t)D.d);
  if (bwd) {
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = HALF_R; K2.dtype = D.dtype;
    w.sub_bytes = fused_k2_bwd_ws(K2);
    w.sub = a.take<char>(w.sub_bytes);
  }
  w.bytes = a.off;
  return w;
}
}  // namespace

bool wide_adapter_supported(const VlpetK1Desc& D, bool bwd) {
  if (D.dtype != VLPET_BF16 || D.M <= 0 || D.r <= HALF_R || D.r > 2 * HALF_R || D.r % 8 != 0 || device_sm_count() <= 0) return false;
  const int halves[2] = {HALF_R, D.r - HALF_R};
  for (int i = 0; i < 2; ++i) {
    if (!fused_k1_fwd_supported(ungated(D, halves[i], 1.f, 0.f))) return false;
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = halves[i]; K2.dtype = D.dtype;
    if (bwd && !fused_k2_supported(K2)) return false;
  }
  return true;
}
size_t wide_adapter_ws(const VlpetK1Desc& D, bool bwd) { return carve_adapter(D, bwd, nullptr).bytes; }

// y1 = kappa x2 + alpha (Up(gelu_new(Down x2)) + bu)
int wide_adapter_fwd(const VlpetK1Desc& D, const void* x2, const VlpetK1Params& w, void* y1, void* ws, size_t ws_bytes, cudaStream_t st) {
  AdWs W = carve_adapter(D, false, ws);
  if (!ws || ws_bytes < W.bytes) return fail(VLPET_E_WORKSPACE, "adapter(wide): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  VLPET_CUDA_OK(cudaMemsetAsync(W.zero_d, 0, (size_t)D.d * 2, st));
  return branch_fwd(D, D.r, W.ra, W.rb, D.alpha, D.kappa - 1.0f, x2, x2, w.Wd, w.bd, w.Wu, w.bu, W.wu, W.zero_d, W.ya,
                    static_cast<__nv_bfloat16*>(y1), st);
}

// dx2 = kappa dy1 + (alpha (dy1 Wu) gelu_new'(a)) Wd and the adapter's parameter gradients; the column halves of Wu must
// still be in the workspace of a wide_adapter_fwd call on the same stream (the row-wise backward recomputes y1 first)
int wide_adapter_bwd(const VlpetK1Desc& D, const void* x2, const void* dy1, const VlpetK1Params& w, void* dx2, const VlpetK1Grads& G,
                     void* ws_fwd, void* ws, size_t ws_bytes, cudaStream_t st) {
  AdWs F = carve_adapter(D, false, ws_fwd);
  AdWs W = carve_adapter(D, true, ws);
  if (!ws || ws_bytes < W.bytes || !ws_fwd) return fail(VLPET_E_WORKSPACE, "adapter_bwd(wide): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  for (int h = 0; h < 2; ++h) {
    const int c0 = h ? W.ra : 0, rh = h ? W.rb : W.ra;
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = rh; K2.dtype = D.dtype; K2.sf = D.alpha;
    VlpetK2Params P2;
    P2.Wd = static_cast<const __nv_bfloat16*>(w.Wd) + (size_t)c0 * D.d; P2.bd = static_cast<const __nv_bfloat16*>(w.bd) + c0;
    P2.Wu = F.wu[h]; P2.bu = h ? static_cast<const void*>(F.zero_d) : w.bu;
    VlpetK2Grads G2;
    G2.dWd = G.dWd ? G.dWd + (size_t)c0 * D.d : nullptr; G2.dbd = G.dbd ? G.dbd + c0 : nullptr;
    G2.dWu = G.dWu ? G.dWu + c0 : nullptr; G2.dbu = h ? nullptr : G.dbu;
    if (G2.dbu && !G2.dWu) return fail(VLPET_E_BADARG, "adapter_bwd(wide): a bias gradient needs its weight gradient buffer");
    BwdExtras ex;
    ex.kap_src = h ? dx2 : nullptr; ex.ldo_wu = D.r;
    VLPET_TRY(fused_k2_bwd_ex(K2, h ? 1.0f : D.kappa, x2, dy1, P2, dx2, G2, ex, W.sub, W.sub_bytes, st));
  }
  return 0;
}

bool wide_k1_supported(const VlpetK1Desc& D, bool bwd) {
  if (D.dtype != VLPET_BF16 || D.gate != VLPET_GATE_LARGE || D.M <= 0 || D.d % 8 != 0) return false;
  const int rmax = D.r > D.rg ? D.r : D.rg;
  if (rmax <= HALF_R || rmax > 2 * HALF_R || D.r % 8 != 0 || D.rg % 8 != 0 || D.r < 8 || D.rg < 8) return false;
  if (device_sm_count() <= 0) return false;
  const int halves[4] = {D.r < HALF_R ? D.r : HALF_R, D.r - HALF_R, D.rg < HALF_R ? D.rg : HALF_R, D.rg - HALF_R};
  for (int i = 0; i < 4; ++i) {
    if (halves[i] <= 0) continue;
    if (!fused_k1_fwd_supported(ungated(D, halves[i], 1.f, 0.f))) return false;
    if (bwd) {
      VlpetK2Desc K2;
      memset(&K2, 0, sizeof(K2));
      K2.M = D.M; K2.d = D.d; K2.r = halves[i]; K2.dtype = D.dtype;
      if (!fused_k2_supported(K2)) return false;
    }
  }
  return true;
}
size_t wide_k1_fwd_ws(const VlpetK1Desc& D) { return carve_wide(D, false, nullptr).bytes; }
size_t wide_k1_bwd_ws(const VlpetK1Desc& D) { return carve_wide(D, true, nullptr).bytes; }

int wide_k1_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  WideWs W = carve_wide(D, false, ws);
  if (!ws || ws_bytes < W.bytes) return fail(VLPET_E_WORKSPACE, "k1_fwd(wide): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  VLPET_TRY(wide_recompute(D, x1, x2, w, W, st));
  GateArgs a = gate_args(D);
  a.x1 = static_cast<const __nv_bfloat16*>(x1); a.y1 = W.y1; a.t = W.t; a.out = static_cast<__nv_bfloat16*>(out);
  wide_gate_fwd_kernel<<<gate_blocks(a.n8, device_sm_count()), 256, 0, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}

int wide_k1_bwd(const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w, void* dx1, void* dx2,
                const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  WideWs W = carve_wide(D, true, ws);
  if (!ws || ws_bytes < W.bytes) return fail(VLPET_E_WORKSPACE, "k1_bwd(wide): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  VLPET_TRY(wide_recompute(D, x1, x2, w, W, st));
  GateArgs a = gate_args(D);
  a.dout = static_cast<const __nv_bfloat16*>(dout); a.y1 = W.y1; a.t = W.t;
  a.dy1 = W.ya; a.dt = W.ta;                       // the first-half temporaries are dead by now
  wide_gate_bwd_kernel<<<gate_blocks(a.n8, device_sm_count()), 256, 0, st>>>(a);
  VLPET_LAUNCH_OK();
  // one ungated backward per branch and rank half: activation gradient = kap * K + d_h W_h, weight gradients of the half
  auto half_bwd = [&](int r_tot, int c0, int rh, float alpha, float kap, const void* kap_src, const void* in, const void* dgrad,
                      const void* Wd, const void* bd, const __nv_bfloat16* wu_h, const void* bias, void* dxo, float* dWd, float* dbd,
                      float* dWu, float* dbu) -> int {
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = rh; K2.dtype = D.dtype; K2.sf = alpha;
    VlpetK2Params P2;
    P2.Wd = static_cast<const __nv_bfloat16*>(Wd) + (size_t)c0 * D.d; P2.bd = static_cast<const __nv_bfloat16*>(bd) + c0;
    P2.Wu = wu_h; P2.bu = bias;
    VlpetK2Grads G2;
    G2.dWd = dWd ? dWd + (size_t)c0 * D.d : nullptr; G2.dbd = dbd ? dbd + c0 : nullptr;
    G2.dWu = dWu ? dWu + c0 : nullptr; G2.dbu = dbu;
    if (G2.dbu && !G2.dWu) return fail(VLPET_E_BADARG, "k1_bwd(wide): a bias gradient needs its weight gradient buffer");
    BwdExtras ex;
    ex.kap_src = kap_src; ex.ldo_wu = r_tot;
    return fused_k2_bwd_ex(K2, kap, in, dgrad, P2, dxo, G2, ex, W.sub, W.sub_bytes, st);
  };
  const __nv_bfloat16* wu0 = W.rb ? W.wu[0] : static_cast<const __nv_bfloat16*>(w.Wu);
  const __nv_bfloat16* gu0 = W.gb ? W.gu[0] : static_cast<const __nv_bfloat16*>(w.Gu);
  // adapter branch (input x2, upstream gradient dy1): dx2 = kappa dy1 + da_a Wd_a, then += da_b Wd_b
  VLPET_TRY(half_bwd(D.r, 0, W.ra, D.alpha, D.kappa, nullptr, x2, a.dy1, w.Wd, w.bd, wu0, w.bu, dx2, G.dWd, G.dbd, G.dWu, G.dbu));
  if (W.rb)
    VLPET_TRY(half_bwd(D.r, W.ra, W.rb, D.alpha, 1.0f, dx2, x2, a.dy1, w.Wd, w.bd, W.wu[1], W.zero_d, dx2, G.dWd, G.dbd, G.dWu, nullptr));
  // gate branch (input x1, upstream gradient dT): dx1 = dout + dp_a Gd_a, then += dp_b Gd_b
  VLPET_TRY(half_bwd(D.rg, 0, W.ga, 1.0f, 1.0f, dout, x1, a.dt, w.Gd, w.gbd, gu0, w.gbu, dx1, G.dGd, G.dgbd, G.dGu, G.dgbu));
  if (W.gb)
    VLPET_TRY(half_bwd(D.rg, W.ga, W.gb, 1.0f, 1.0f, dx1, x1, a.dt, w.Gd, w.gbd, W.gu[1], W.zero_d, dx1, G.dGd, G.dgbd, G.dGu, nullptr));
  return 0;
}

}  // namespace vlpet
