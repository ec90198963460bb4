// LayerNorm that follows every encoder PET site (my_transformers/modeling_bart.py:1260-1261, 1376-1377
// `self.self_attn_layer_norm(hidden_states)` / `self.final_layer_norm(hidden_states)`), forward and backward, for the
// training configuration: bf16 (or fp32) activations, fp32 affine parameters (they are trainable under
// --unfreeze_encoder_layer_norms), fp32 statistics.
//
// Pure HBM-bound row kernels: one warp per row, the row lives in registers (d <= 1024, d % 256 == 0), persistent grid.
//   fwd: reads x once, writes y and (mean, rstd)                      algorithmic bytes / row: 2 d e (+8)
//   bwd: reads x, dy once, writes dx; dgamma / dbeta are accumulated in registers over all rows a warp walks and
//        reduced once per block through shared memory + one fp32 atomic per block and column      3 d e / row
// (torch's GammaBetaBackward kernel alone took 108 us per call at M = 28 000; this backward is one launch.)
#include "vlpet_common.cuh"

namespace vlpet {
namespace {

template <int VPL>  // uint4 (8 bf16) vectors per lane: d = 256 * VPL
__global__ void __launch_bounds__(256) ln_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ mean, float* __restrict__ rstd, int64_t M,
                                                          float eps) {
  constexpr int D = 256 * VPL;
  const int lane = threadIdx.x % 32;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
  // Rows are latency-bound (load -> two warp reductions -> store): the next row's loads are issued before this row's
  // reductions, so every warp keeps a row in flight while it computes.
  uint4 cur[VPL];
  if (warp0 < M) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = *reinterpret_cast<const uint4*>(x + warp0 * D + (i * 32 + lane) * 8);
  }
  for (int64_t row = warp0; row < M; row += nwarps) {
    uint4 nxt[VPL];
    const int64_t nrow = row + nwarps;
    if (nrow < M) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) nxt[i] = *reinterpret_cast<const uint4*>(x + nrow * D + (i * 32 + lane) * 8);
    }
    float v[VPL * 8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint4 q = cur[i];
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[i * 8 + 2 * e] = __uint_as_float(u[e] << 16);
        v[i * 8 + 2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u);
        s += v[i * 8 + 2 * e] + v[i * 8 + 2 * e + 1];
      }
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL * 8; ++i) { v[i] -= mu; ss += v[i] * v[i]; }
    const float rs = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c) + 1);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 t = __floats2bfloat162_rn(fmaf(v[i * 8 + 2 * e] * rs, ww[2 * e], bb[2 * e]),
                                                 fmaf(v[i * 8 + 2 * e + 1] * rs, ww[2 * e + 1], bb[2 * e + 1]));
        o[e] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(y + row * D + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                          const float* __restrict__ w, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx,
                                                          float* __restrict__ dw, float* __restrict__ db, int64_t M) {
  constexpr int D = 256 * VPL;
  __shared__ float red[8][32 * 8 + 8];   // [warp][lane * 8 + e] staging for the block-level column reduction
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x / 32) + wid;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
  float ww[VPL * 8], aw[VPL * 8], ab[VPL * 8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 8;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c) + 1);
    ww[i * 8 + 0] = w0.x; ww[i * 8 + 1] = w0.y; ww[i * 8 + 2] = w0.z; ww[i * 8 + 3] = w0.w;
    ww[i * 8 + 4] = w1.x; ww[i * 8 + 5] = w1.y; ww[i * 8 + 6] = w1.z; ww[i * 8 + 7] = w1.w;
  }
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { aw[i] = 0.f; ab[i] = 0.f; }
  for (int64_t row = warp0; row < M; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float xh[VPL * 8], g[VPL * 8];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      const uint4 qx = *reinterpret_cast<const uint4*>(x + row * D + c);
      const uint4 qd = *reinterpret_cast<const uint4*>(dy + row * D + c);
      const uint32_t ux[4] = {qx.x, qx.y, qx.z, qx.w}, ud[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = (e & 1) ? __uint_as_float(ux[e >> 1] & 0xffff0000u) : __uint_as_float(ux[e >> 1] << 16);
        const float dv = (e & 1) ? __uint_as_float(ud[e >> 1] & 0xffff0000u) : __uint_as_float(ud[e >> 1] << 16);
        const float h = (xv - mu) * rs;
        xh[i * 8 + e] = h;
        ab[i * 8 + e] += dv;
        aw[i * 8 + e] += dv * h;
        const float gv = dv * ww[i * 8 + e];
        g[i * 8 + e] = gv;
        sg += gv;
        sgx += gv * h;
      }
    }
    const float mg = warp_sum(sg) * (1.0f / D), mgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 t = __floats2bfloat162_rn(rs * (g[i * 8 + 2 * e] - mg - xh[i * 8 + 2 * e] * mgx),
                                                 rs * (g[i * 8 + 2 * e + 1] - mg - xh[i * 8 + 2 * e + 1] * mgx));
        o[e] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(dx + row * D + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (!dw && !db) return;
  // block-level reduction of the per-warp column partials, then one atomic per block and column
  for (int pass = 0; pass < 2; ++pass) {
    float* acc = pass ? ab : aw;
    float* dst = pass ? db : dw;
    if (!dst) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) red[wid][lane * 8 + e] = acc[i * 8 + e];
      __syncthreads();
      // 256 threads <-> the 256 columns of this slab
      const int t = threadIdx.x;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += red[k][t];
      atomicAdd(dst + i * 256 + t, sum);
    }
  }
}

// ---- y = LayerNorm(res + dropout_p(h)): the post-LN residual step of the frozen decoder blocks in one pass ---------------------
// (my_transformers/modeling_bart.py:1663-1665, 1683-1685, 1697-1699: F.dropout -> residual add -> LayerNorm = three kernels and
// four [M, d] round trips per sublayer forward, 18 times per step; at the 8-GPU per-rank batch these are ~5 us launches that do
// not shrink).  xs = res + dropout(h) is kept (bf16) for the backward, the mask is the counter-based stream (regenerated).
template <int VPL>
__global__ void __launch_bounds__(256) dal_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ res,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ xs,
                                                           float* __restrict__ mean, float* __restrict__ rstd, int64_t M, float eps,
                                                           uint32_t thr16, float inv_keep, uint64_t seed, const uint64_t* seed_dev) {
  constexpr int D = 256 * VPL;
  const int lane = threadIdx.x % 32;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
  const uint64_t seed_eff = seed + ((thr16 && seed_dev) ? __ldg(seed_dev) : 0ull);
  for (int64_t row = warp0; row < M; row += nwarps) {
    float v[VPL * 8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      const uint4 qh = *reinterpret_cast<const uint4*>(h + row * D + c);
      const uint4 qr = *reinterpret_cast<const uint4*>(res + row * D + c);
      const uint32_t uh[4] = {qh.x, qh.y, qh.z, qh.w}, ur[4] = {qr.x, qr.y, qr.z, qr.w};
      float m[8];
      drop_scale8(seed_eff, thr16, inv_keep, row * D + c, m);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // x = res + dropout(h), rounded to bf16 exactly as the unfused sequence stores it (statistics are taken on the stored value)
        const float h0 = __uint_as_float(uh[e] << 16) * m[2 * e], h1 = __uint_as_float(uh[e] & 0xffff0000u) * m[2 * e + 1];
        __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(ur[e] << 16) + h0, __uint_as_float(ur[e] & 0xffff0000u) + h1);
        o[e] = *reinterpret_cast<uint32_t*>(&t);
        v[i * 8 + 2 * e] = __uint_as_float(o[e] << 16);
        v[i * 8 + 2 * e + 1] = __uint_as_float(o[e] & 0xffff0000u);
        s += v[i * 8 + 2 * e] + v[i * 8 + 2 * e + 1];
      }
      *reinterpret_cast<uint4*>(xs + row * D + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL * 8; ++i) { v[i] -= mu; ss += v[i] * v[i]; }
    const float rs = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c) + 1);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 t = __floats2bfloat162_rn(fmaf(v[i * 8 + 2 * e] * rs, ww[2 * e], bb[2 * e]),
                                                 fmaf(v[i * 8 + 2 * e + 1] * rs, ww[2 * e + 1], bb[2 * e + 1]));
        o[e] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(y + row * D + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// backward: dres = LayerNorm backward of dy at xs, dh = dres * mask / (1 - p); dw / db accumulated (may be null: frozen LayerNorm)
template <int VPL>
__global__ void __launch_bounds__(256) dal_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           const float* __restrict__ w, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dres,
                                                           __nv_bfloat16* __restrict__ dh, float* __restrict__ dw, float* __restrict__ db,
                                                           int64_t M, uint32_t thr16, float inv_keep, uint64_t seed,
                                                           const uint64_t* seed_dev) {
  constexpr int D = 256 * VPL;
  __shared__ float red[8][32 * 8 + 8];
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x / 32) + wid;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x / 32);
  const uint64_t seed_eff = seed + ((thr16 && seed_dev) ? __ldg(seed_dev) : 0ull);
  const bool want_param = dw || db;
  float ww[VPL * 8], aw[VPL * 8], ab[VPL * 8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 8;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c) + 1);
    ww[i * 8 + 0] = w0.x; ww[i * 8 + 1] = w0.y; ww[i * 8 + 2] = w0.z; ww[i * 8 + 3] = w0.w;
    ww[i * 8 + 4] = w1.x; ww[i * 8 + 5] = w1.y; ww[i * 8 + 6] = w1.z; ww[i * 8 + 7] = w1.w;
  }
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { aw[i] = 0.f; ab[i] = 0.f; }
  for (int64_t row = warp0; row < M; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float xh[VPL * 8], g[VPL * 8];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      const uint4 qx = *reinterpret_cast<const uint4*>(x + row * D + c);
      const uint4 qd = *reinterpret_cast<const uint4*>(dy + row * D + c);
      const uint32_t ux[4] = {qx.x, qx.y, qx.z, qx.w}, ud[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = (e & 1) ? __uint_as_float(ux[e >> 1] & 0xffff0000u) : __uint_as_float(ux[e >> 1] << 16);
        const float dv = (e & 1) ? __uint_as_float(ud[e >> 1] & 0xffff0000u) : __uint_as_float(ud[e >> 1] << 16);
        const float hh = (xv - mu) * rs;
        xh[i * 8 + e] = hh;
        if (want_param) { ab[i * 8 + e] += dv; aw[i * 8 + e] += dv * hh; }
        const float gv = dv * ww[i * 8 + e];
        g[i * 8 + e] = gv;
        sg += gv;
        sgx += gv * hh;
      }
    }
    const float mg = warp_sum(sg) * (1.0f / D), mgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 8;
      float m[8];
      drop_scale8(seed_eff, thr16, inv_keep, row * D + c, m);
      uint32_t o[4], oh[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = rs * (g[i * 8 + 2 * e] - mg - xh[i * 8 + 2 * e] * mgx);
        const float d1 = rs * (g[i * 8 + 2 * e + 1] - mg - xh[i * 8 + 2 * e + 1] * mgx);
        __nv_bfloat162 t = __floats2bfloat162_rn(d0, d1);
        o[e] = *reinterpret_cast<uint32_t*>(&t);
        // the unfused backward scales the bf16-stored dx: do the same
        __nv_bfloat162 th = __floats2bfloat162_rn(__uint_as_float(o[e] << 16) * m[2 * e], __uint_as_float(o[e] & 0xffff0000u) * m[2 * e + 1]);
        oh[e] = *reinterpret_cast<uint32_t*>(&th);
      }
      *reinterpret_cast<uint4*>(dres + row * D + c) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(dh + row * D + c) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    }
  }
  if (!want_param) return;
  for (int pass = 0; pass < 2; ++pass) {
    float* acc = pass ? ab : aw;
    float* dst = pass ? db : dw;
    if (!dst) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) red[wid][lane * 8 + e] = acc[i * 8 + e];
      __syncthreads();
      const int t = threadIdx.x;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += red[k][t];
      atomicAdd(dst + i * 256 + t, sum);
    }
  }
}

}  // namespace

bool layernorm_supported(int d, int dtype) { return dtype == VLPET_BF16 && d % 256 == 0 && d >= 256 && d <= 1024; }

int layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M, int d,
                  float eps, cudaStream_t st) {
  int64_t blocks = (M + 7) / 8;
  if (blocks > 148 * 3) blocks = 148 * 3;   // all resident (3 blocks of 256 threads per SM): every warp pipelines its rows
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  switch (d / 256) {
    case 1: ln_fwd_bf16_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(xb, w, b, yb, mean, rstd, M, eps); break;
    case 2: ln_fwd_bf16_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(xb, w, b, yb, mean, rstd, M, eps); break;
    case 3: ln_fwd_bf16_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(xb, w, b, yb, mean, rstd, M, eps); break;
    case 4: ln_fwd_bf16_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(xb, w, b, yb, mean, rstd, M, eps); break;
    default: return fail(VLPET_E_UNSUPPORTED, "layernorm: d=%d", d);
  }
  VLPET_LAUNCH_OK();
  return 0;
}

int layernorm_bwd(const void* x, const void* dy, const float* w, const float* mean, const float* rstd, void* dx, float* dw,
                  float* db, int64_t M, int d, cudaStream_t st) {
  int64_t blocks = (M + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;   // few blocks: each ends with d atomics per parameter
  const __nv_bfloat16 *xb = static_cast<const __nv_bfloat16*>(x), *db16 = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx);
  switch (d / 256) {
    case 1: ln_bwd_bf16_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(xb, db16, w, mean, rstd, dxb, dw, db, M); break;
    case 2: ln_bwd_bf16_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(xb, db16, w, mean, rstd, dxb, dw, db, M); break;
    case 3: ln_bwd_bf16_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(xb, db16, w, mean, rstd, dxb, dw, db, M); break;
    case 4: ln_bwd_bf16_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(xb, db16, w, mean, rstd, dxb, dw, db, M); break;
    default: return fail(VLPET_E_UNSUPPORTED, "layernorm: d=%d", d);
  }
  VLPET_LAUNCH_OK();
  return 0;
}

int dropout_add_layernorm_fwd(const void* h, const void* res, const float* w, const float* b, void* y, void* xs, float* mean,
                              float* rstd, int64_t M, int d, float eps, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                              cudaStream_t st) {
  int64_t blocks = (M + 7) / 8;
  if (blocks > 148 * 3) blocks = 148 * 3;
  const uint32_t thr16 = p_drop > 0.f ? drop_thr16(p_drop) : 0u;
  const float inv_keep = thr16 ? 1.0f / (1.0f - (float)thr16 / 65536.0f) : 1.0f;
  const __nv_bfloat16 *hb = static_cast<const __nv_bfloat16*>(h), *rb = static_cast<const __nv_bfloat16*>(res);
  __nv_bfloat16 *yb = static_cast<__nv_bfloat16*>(y), *xb = static_cast<__nv_bfloat16*>(xs);
#define VLPET_DAL_F(V) dal_fwd_bf16_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(hb, rb, w, b, yb, xb, mean, rstd, M, eps, thr16, inv_keep, seed, seed_dev)
  switch (d / 256) {
    case 1: VLPET_DAL_F(1); break;
    case 2: VLPET_DAL_F(2); break;
    case 3: VLPET_DAL_F(3); break;
    case 4: VLPET_DAL_F(4); break;
    default: return fail(VLPET_E_UNSUPPORTED, "dropout_add_layernorm: d=%d", d);
  }
#undef VLPET_DAL_F
  VLPET_LAUNCH_OK();
  return 0;
}

int dropout_add_layernorm_bwd(const void* xs, const void* dy, const float* w, const float* mean, const float* rstd, void* dres, void* dh,
                              float* dw, float* db, int64_t M, int d, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                              cudaStream_t st) {
  int64_t blocks = (M + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;
  const uint32_t thr16 = p_drop > 0.f ? drop_thr16(p_drop) : 0u;
  const float inv_keep = thr16 ? 1.0f / (1.0f - (float)thr16 / 65536.0f) : 1.0f;
  const __nv_bfloat16 *xb = static_cast<const __nv_bfloat16*>(xs), *db16 = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16 *drb = static_cast<__nv_bfloat16*>(dres), *dhb = static_cast<__nv_bfloat16*>(dh);
#define VLPET_DAL_B(V) dal_bwd_bf16_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(xb, db16, w, mean, rstd, drb, dhb, dw, db, M, thr16, inv_keep, seed, seed_dev)
  switch (d / 256) {
    case 1: VLPET_DAL_B(1); break;
    case 2: VLPET_DAL_B(2); break;
    case 3: VLPET_DAL_B(3); break;
    case 4: VLPET_DAL_B(4); break;
    default: return fail(VLPET_E_UNSUPPORTED, "dropout_add_layernorm: d=%d", d);
  }
#undef VLPET_DAL_B
  VLPET_LAUNCH_OK();
  return 0;
}

}  // namespace vlpet
