// K1 for the granularities whose gate is not a pair of GEMMs -- middleX (N x 1), middleY (1 x d), small (1 x 1), and the
// ungated form -- and for ranks too small for a tensor-core tile (r <= 16, e.g. BASELINE config 4: VL-PET-small, r = 4).
// Reference: my_transformers/modeling_bart.py:1210-1231 (gates), 1145-1155 (adapter), 1256-1260 (scale, dropout,
// residual); T5 twins my_transformers/modeling_t5.py:777-824, 359-409.  Math: oracle/pet_oracle.py gated_pet_fwd / _bwd.
//
// These gates are ROW-WISE (a dot product over d per token, a per-column vector, a per-sample mean of per-token scalars):
// HBM-bound vector work, no GEMM.  One warp owns a token row: lane l holds the 8 elements [256 ch + 8 l, +8) of every
// 256-wide chunk as one 16-byte load -- coalesced 512-byte warp accesses, everything else in registers, row reductions by
// shuffles.  Two modes:
//   INLINE (r <= 16): the adapter y1 = kappa x2 + alpha (gelu_new(x2 Wd^T + bd) Wu^T + bu) is evaluated in the same kernel
//     from Wd / Wu^T staged once per CTA in shared memory (r dot products of length d per token: vector FMAs) -- ONE launch
//     forward; the backward writes du / z / da for the token-contracted weight-gradient GEMM (vlpet_wgrad_sm100.cu).
//   COMPOSED (r a multiple of 8 up to the tensor-core buckets): y1 comes from the fused tcgen05 kernel in its ungated form
//     (vlpet_k1_sm100.cu), this kernel applies the gate; backward: gate kernel -> dy1, then the ungated tcgen05 backward.
// The small gate needs the per-sample mean of sigmoid over the sequence before any output can be formed: a first pass
// (same kernel, PASS 0) accumulates it (forward: gate values, backward: dG) with one atomicAdd per token.
#include <cstring>

#include "vlpet_common.cuh"

namespace vlpet {
namespace {

constexpr int RW_THREADS = 256;   // 8 warps = 8 rows in flight per CTA
constexpr int RW_MAXR = 16;

struct RowsArgs {
  int64_t M;
  int L, d, r, r8, gate, add_gate, pass, has_y1;
  float s, alpha, kappa, inv_keep;
  uint32_t thr16;
  uint64_t seed;
  const uint64_t* seed_dev;
  const __nv_bfloat16 *x1, *x2, *y1in, *dout;
  __nv_bfloat16 *out, *dx1, *dx2, *dy1, *du;           // du: [M, d] scratch for the weight-gradient GEMM (INLINE backward)
  __nv_bfloat16 *zs, *das;                             // [M, pz] scratch (z | 1 | 0.., da)
  int pz;
  const __nv_bfloat16 *Wd, *bd, *Wu, *bu, *gw, *gb, *gz;
  float *gmean, *dgsum;                                // [B] per-sample gate mean / dG sum (small gate)
  float *dgw, *dgb, *dgz;                              // fp32 gate gradients, accumulated into
  float* dbd;                                          // INLINE backward: fp32 down-projection bias gradient, accumulated into
};

__device__ __forceinline__ float bf2f(uint32_t v, int hi) { return __uint_as_float(hi ? (v & 0xffff0000u) : (v << 16)); }
__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <int NCH>
__device__ __forceinline__ void load_row(const __nv_bfloat16* row, int lane, float (&v)[NCH * 8]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + ch * 256 + lane * 8));
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[ch * 8 + 2 * e] = bf2f(u[e], 0); v[ch * 8 + 2 * e + 1] = bf2f(u[e], 1); }
  }
}
template <int NCH>
__device__ __forceinline__ void store_row(__nv_bfloat16* row, int lane, const float (&v)[NCH * 8]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    uint4 q;
    q.x = pack_bf(v[ch * 8 + 0], v[ch * 8 + 1]); q.y = pack_bf(v[ch * 8 + 2], v[ch * 8 + 3]);
    q.z = pack_bf(v[ch * 8 + 4], v[ch * 8 + 5]); q.w = pack_bf(v[ch * 8 + 6], v[ch * 8 + 7]);
    *reinterpret_cast<uint4*>(row + ch * 256 + lane * 8) = q;
  }
}
// 8 consecutive bf16 of a shared-memory row -> floats
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&w)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { w[2 * e] = bf2f(u[e], 0); w[2 * e + 1] = bf2f(u[e], 1); }
}

// Shared state of one row: forward quantities the backward needs again.
template <int NCH>
struct RowState {
  float x1[NCH * 8], y1[NCH * 8];
  float a[RW_MAXR], z[RW_MAXR];   // INLINE only
  float g;                        // middleX: sigmoid of the row's dot; small: the sample's mean gate
  float sg;                       // small: this token's sigmoid
};

// y1 (INLINE: from x2 and the staged weights; COMPOSED: loaded), then the row's gate scalars.
template <int NCH, int GATE>
__device__ __forceinline__ void row_forward(const RowsArgs& p, int64_t row, int lane, const __nv_bfloat16* sWd, const __nv_bfloat16* sWuT,
                                            const float* gwr, float gbias, RowState<NCH>& S, float (&x2)[NCH * 8]) {
  constexpr int NE = NCH * 8;
  load_row<NCH>(p.x1 + row * p.d, lane, S.x1);
  if (p.has_y1) {
    load_row<NCH>(p.y1in + row * p.d, lane, S.y1);
  } else {
    load_row<NCH>(p.x2 + row * p.d, lane, x2);
#pragma unroll
    for (int j = 0; j < RW_MAXR; ++j) {
      if (j >= p.r) break;
      float acc = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float w[8];
        lds8(sWd + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(x2[ch * 8 + e], w[e], acc);
      }
      acc = warp_sum(acc) + __bfloat162float(p.bd[j]);
      S.a[j] = acc;
      S.z[j] = gelu_new_f(acc);
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.bu + ch * 256 + lane * 8));
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { S.y1[ch * 8 + 2 * e] = bf2f(u[e], 0); S.y1[ch * 8 + 2 * e + 1] = bf2f(u[e], 1); }
    }
#pragma unroll
    for (int j = 0; j < RW_MAXR; ++j) {
      if (j >= p.r) break;
      const float zj = S.z[j];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float w[8];
        lds8(sWuT + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) S.y1[ch * 8 + e] = fmaf(zj, w[e], S.y1[ch * 8 + e]);
      }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) S.y1[e] = p.kappa * x2[e] + p.alpha * S.y1[e];
  }
  S.g = 1.f;
  S.sg = 0.f;
  if (GATE == VLPET_GATE_MIDDLE_X) {
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) t = fmaf(S.x1[e] + S.y1[e], gwr[e], t);
    S.g = sigmoid_f(warp_sum(t) + gbias);
  } else if (GATE == VLPET_GATE_SMALL) {
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) t = fmaf(S.x1[e], gwr[e], fmaf(S.y1[e], gwr[NE + e], t));
    S.sg = sigmoid_f(warp_sum(t) + gbias);
    if (p.pass != 0) S.g = p.gmean[row / p.L];
  }
}

// per-lane copies of the gate's row vector(s): middleX gw[d], middleY gz[d], small gw[2d] (x1 half, then y1 half)
template <int NCH, int GATE>
__device__ __forceinline__ void load_gate_vec(const RowsArgs& p, int lane, float (&gwr)[2 * NCH * 8], float& gbias) {
  constexpr int NE = NCH * 8;
  gbias = 0.f;
#pragma unroll
  for (int e = 0; e < 2 * NE; ++e) gwr[e] = 0.f;
  const __nv_bfloat16* v = GATE == VLPET_GATE_MIDDLE_Y ? p.gz : p.gw;
  if (GATE == VLPET_GATE_NONE) return;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gwr[ch * 8 + e] = __bfloat162float(v[ch * 256 + lane * 8 + e]);
      if (GATE == VLPET_GATE_SMALL) gwr[NE + ch * 8 + e] = __bfloat162float(v[p.d + ch * 256 + lane * 8 + e]);
    }
  if (GATE != VLPET_GATE_MIDDLE_Y) gbias = __bfloat162float(p.gb[0]);
}

__device__ __forceinline__ void stage_weights(const RowsArgs& p, __nv_bfloat16* sWd, __nv_bfloat16* sWuT) {
  if (p.has_y1) return;
  const int n = p.r * p.d;
  for (int i = threadIdx.x; i < n; i += RW_THREADS) {
    sWd[i] = p.Wd[i];
    const int j = i / p.d, c = i % p.d;
    sWuT[i] = p.Wu[(size_t)c * p.r + j];      // Wu is [d, r]: transposed into [r, d] rows
  }
  __syncthreads();
}

template <int NCH, int GATE>
__global__ void __launch_bounds__(RW_THREADS) rows_fwd_kernel(const RowsArgs p) {
  constexpr int NE = NCH * 8;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sWd = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sWuT = sWd + (size_t)p.r * p.d;
  stage_weights(p, sWd, sWuT);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gwr[2 * NE], gbias;
  load_gate_vec<NCH, GATE>(p, lane, gwr, gbias);
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * (RW_THREADS / 32);
  for (int64_t row = (int64_t)blockIdx.x * (RW_THREADS / 32) + warp; row < p.M; row += nw) {
    RowState<NCH> S;
    float x2[NE];
    row_forward<NCH, GATE>(p, row, lane, sWd, sWuT, gwr, gbias, S, x2);
    if (GATE == VLPET_GATE_SMALL && p.pass == 0) {
      if (lane == 0) atomicAdd(p.gmean + row / p.L, S.sg / (float)p.L);
      continue;
    }
    float o[NE];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float m[8];
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + ch * 256 + lane * 8, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = ch * 8 + e;
        float h;
        if (GATE == VLPET_GATE_MIDDLE_Y) h = p.add_gate ? S.y1[i] + 1.f + gwr[i] : S.y1[i] * (1.f + gwr[i]);
        else if (GATE == VLPET_GATE_NONE) h = S.y1[i];
        else h = p.add_gate ? S.y1[i] + S.g : S.y1[i] * S.g;
        o[i] = S.x1[i] + p.s * m[e] * h;
      }
    }
    store_row<NCH>(p.out + row * p.d, lane, o);
  }
}

template <int NCH, int GATE>
__global__ void __launch_bounds__(RW_THREADS) rows_bwd_kernel(const RowsArgs p) {
  constexpr int NE = NCH * 8;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sWd = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sWuT = sWd + (size_t)p.r * p.d;
  float* sred = reinterpret_cast<float*>(smem + (p.has_y1 ? 0 : (size_t)2 * p.r * p.d * 2));   // [2 d + 1] block reduction of gate grads
  stage_weights(p, sWd, sWuT);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gwr[2 * NE], gbias;
  load_gate_vec<NCH, GATE>(p, lane, gwr, gbias);
  float acc_g[2 * NE], acc_b = 0.f;     // gate-parameter gradients of the rows this warp handles
#pragma unroll
  for (int e = 0; e < 2 * NE; ++e) acc_g[e] = 0.f;
  float acc_da[RW_MAXR];                // dbd = sum over rows of da, kept in fp32 (the bf16 da scratch only feeds the GEMM)
#pragma unroll
  for (int j = 0; j < RW_MAXR; ++j) acc_da[j] = 0.f;
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * (RW_THREADS / 32);
  for (int64_t row = (int64_t)blockIdx.x * (RW_THREADS / 32) + warp; row < p.M; row += nw) {
    RowState<NCH> S;
    float x2[NE], dh[NE], dy1[NE];
    row_forward<NCH, GATE>(p, row, lane, sWd, sWuT, gwr, gbias, S, x2);
    float dx1[NE];
    load_row<NCH>(p.dout + row * p.d, lane, dx1);           // dx1 starts as dout
    float dgs = 0.f;                                        // sum_c dG contribution of this row
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float m[8];
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + ch * 256 + lane * 8, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = ch * 8 + e;
        dh[i] = p.s * m[e] * dx1[i];
        dgs += p.add_gate ? dh[i] : dh[i] * S.y1[i];
      }
    }
    if (GATE == VLPET_GATE_SMALL && p.pass == 0) {          // first pass: dG of the sample = sum over its tokens and columns
      dgs = warp_sum(dgs);
      if (lane == 0) atomicAdd(p.dgsum + row / p.L, dgs);
      continue;
    }
    if (GATE == VLPET_GATE_NONE) {
#pragma unroll
      for (int i = 0; i < NE; ++i) dy1[i] = dh[i];
    } else if (GATE == VLPET_GATE_MIDDLE_Y) {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        dy1[i] = p.add_gate ? dh[i] : dh[i] * (1.f + gwr[i]);
        acc_g[i] += p.add_gate ? dh[i] : dh[i] * S.y1[i];
      }
    } else {
      // middleX: dt = dG g (1 - g) with dG the row sum; small: dt = dG_sample / L * sg (1 - sg), gate value = the sample mean
      float dt;
      if (GATE == VLPET_GATE_MIDDLE_X) dt = warp_sum(dgs) * S.g * (1.f - S.g);
      else dt = p.dgsum[row / p.L] / (float)p.L * S.sg * (1.f - S.sg);
      acc_b += dt;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const float base = p.add_gate ? dh[i] : dh[i] * S.g;
        if (GATE == VLPET_GATE_MIDDLE_X) {
          acc_g[i] += dt * (S.x1[i] + S.y1[i]);
          dy1[i] = base + dt * gwr[i];
          dx1[i] += dt * gwr[i];
        } else {
          acc_g[i] += dt * S.x1[i];
          acc_g[NE + i] += dt * S.y1[i];
          dx1[i] += dt * gwr[i];
          dy1[i] = base + dt * gwr[NE + i];
        }
      }
    }
    store_row<NCH>(p.dx1 + row * p.d, lane, dx1);
    if (p.has_y1) {
      store_row<NCH>(p.dy1 + row * p.d, lane, dy1);          // the ungated tcgen05 backward takes it from here
      continue;
    }
    // ---- INLINE adapter backward: du = alpha dy1, dz = du Wu, da = dz gelu'(a), dx2 = kappa dy1 + da Wd
    float du[NE], dx2[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) { du[i] = p.alpha * dy1[i]; dx2[i] = p.kappa * dy1[i]; }
    __nv_bfloat16* zrow = p.zs + row * p.pz;
    __nv_bfloat16* darow = p.das + row * p.pz;
#pragma unroll
    for (int j = 0; j < RW_MAXR; ++j) {
      if (j >= p.r) break;
      float dz = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float w[8];
        lds8(sWuT + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) dz = fmaf(du[ch * 8 + e], w[e], dz);
      }
      const float da = warp_sum(dz) * gelu_new_grad_f(S.a[j]);
      acc_da[j] += da;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float w[8];
        lds8(sWd + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) dx2[ch * 8 + e] = fmaf(da, w[e], dx2[ch * 8 + e]);
      }
      if (lane == 0) { zrow[j] = __float2bfloat16_rn(S.z[j]); darow[j] = __float2bfloat16_rn(da); }
    }
    if (lane == 0) {      // zero padding up to the GEMM's rank r8, then the ones column (its bias-gradient trick) at index r8
      for (int j = p.r; j < p.pz; ++j) { zrow[j] = __float2bfloat16_rn(j == p.r8 ? 1.f : 0.f); darow[j] = __float2bfloat16_rn(0.f); }
    }
    store_row<NCH>(p.dx2 + row * p.d, lane, dx2);
    store_row<NCH>(p.du + row * p.d, lane, du);
  }
  if (!p.has_y1 && p.dbd && lane == 0 && !(GATE == VLPET_GATE_SMALL && p.pass == 0)) {
#pragma unroll
    for (int j = 0; j < RW_MAXR; ++j)
      if (j < p.r) atomicAdd(p.dbd + j, acc_da[j]);
  }
  // ---- gate-parameter gradients: lanes own distinct columns; reduce over the CTA's warps in shared memory, then one
  //      atomicAdd per column and CTA
  if (GATE == VLPET_GATE_NONE || (GATE == VLPET_GATE_SMALL && p.pass == 0)) return;
  const int nvec = (GATE == VLPET_GATE_SMALL ? 2 : 1) * p.d;
  __syncthreads();                                           // staged weights no longer needed (sred may overlap nothing, kept simple)
  for (int i = threadIdx.x; i < nvec + 1; i += RW_THREADS) sred[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sred[ch * 256 + lane * 8 + e], acc_g[ch * 8 + e]);
      if (GATE == VLPET_GATE_SMALL) atomicAdd(&sred[p.d + ch * 256 + lane * 8 + e], acc_g[NE + ch * 8 + e]);
    }
  if (lane == 0 && GATE != VLPET_GATE_MIDDLE_Y) atomicAdd(&sred[nvec], acc_b);
  __syncthreads();
  float* gvec = GATE == VLPET_GATE_MIDDLE_Y ? p.dgz : p.dgw;
  if (gvec)
    for (int i = threadIdx.x; i < nvec; i += RW_THREADS) atomicAdd(gvec + i, sred[i]);
  if (GATE != VLPET_GATE_MIDDLE_Y && p.dgb && threadIdx.x == 0) atomicAdd(p.dgb, sred[nvec]);
}

// add the first r columns / rows of the zero-padded weight-gradient buffers (rank padded to 8 for the GEMM) into the real ones
__global__ void unpad_add_kernel(const float* __restrict__ pWu, const float* __restrict__ pWd, float* dWu, float* dWd, int d, int r, int r8) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d * r) return;
  if (dWu) { const int c = i / r, j = i % r; dWu[i] += pWu[c * r8 + j]; }
  if (dWd) dWd[i] += pWd[i];                       // [r8, d] row-major: the first r rows are the first r*d elements
}

size_t smem_for(const RowsArgs& a, bool bwd) {
  size_t s = a.has_y1 ? 0 : (size_t)2 * a.r * a.d * 2;
  if (bwd) s += (size_t)(2 * a.d + 1) * 4 + 16;
  return s;
}

template <int NCH>
int launch_rows(bool bwd, const RowsArgs& a, int sms, cudaStream_t st) {
  const size_t smem = smem_for(a, bwd);
  int64_t blocks = (a.M + (RW_THREADS / 32) - 1) / (RW_THREADS / 32);
  if (blocks > 2 * sms) blocks = 2 * sms;
#define VLPET_ROWS_CASE(G)                                                                                          \
  case G: {                                                                                                          \
    auto kf = rows_fwd_kernel<NCH, G>;                                                                               \
    auto kb = rows_bwd_kernel<NCH, G>;                                                                               \
    if (smem > 48 * 1024) {                                                                                          \
      if (bwd) VLPET_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
      else VLPET_CUDA_OK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    }                                                                                                                \
    if (bwd) kb<<<(unsigned)blocks, RW_THREADS, smem, st>>>(a);                                                      \
    else kf<<<(unsigned)blocks, RW_THREADS, smem, st>>>(a);                                                          \
    break;                                                                                                           \
  }
  switch (a.gate) {
    VLPET_ROWS_CASE(VLPET_GATE_NONE)
    VLPET_ROWS_CASE(VLPET_GATE_MIDDLE_X)
    VLPET_ROWS_CASE(VLPET_GATE_MIDDLE_Y)
    VLPET_ROWS_CASE(VLPET_GATE_SMALL)
    default: return fail(VLPET_E_UNSUPPORTED, "rows: gate %d", a.gate);
  }
#undef VLPET_ROWS_CASE
  VLPET_LAUNCH_OK();
  return 0;
}
int launch_rows_d(bool bwd, const RowsArgs& a, int sms, cudaStream_t st) {
  switch (a.d / 256) {
    case 1: return launch_rows<1>(bwd, a, sms, st);
    case 2: return launch_rows<2>(bwd, a, sms, st);
    case 3: return launch_rows<3>(bwd, a, sms, st);
  }
  return fail(VLPET_E_UNSUPPORTED, "rows: d = %d", a.d);
}

bool inline_mode(const VlpetK1Desc& D) { return D.r <= RW_MAXR; }
VlpetK1Desc adapter_desc(const VlpetK1Desc& D) {   // y1 = x2 + 1 * ((kappa - 1) x2 + alpha (Up(gelu_new(Down x2)) + bu)) through the ungated fused kernel
  VlpetK1Desc K = D;
  K.gate = VLPET_GATE_NONE; K.rg = 0; K.add_gate = 0; K.s = 1.0f; K.kappa = D.kappa - 1.0f; K.p_drop = 0.f; K.seed = 0;
  K.seed_dev = nullptr; K.impl = VLPET_IMPL_AUTO;
  return K;
}
int r8_of(int r) { return r < 8 ? 8 : (r + 7) / 8 * 8; }

struct RowsWs {
  __nv_bfloat16 *y1, *dy1, *du, *zs, *das;
  float *gmean, *dgsum, *pWu, *pWd;
  void* sub;          // workspace of the composed tcgen05 backward
  void* wfwd;         // r > 96: workspace of the two-half adapter forward (vlpet_wide.cu)
  size_t sub_bytes, wfwd_bytes, bytes;
  int pz, r8;
};
bool wide_mode(const VlpetK1Desc& D) { return D.r > 96; }   // the adapter runs as two rank halves (vlpet_wide.cu)
RowsWs carve_rows(const VlpetK1Desc& D, bool bwd, void* ws) {
  RowsWs w;
  memset(&w, 0, sizeof(w));
  Arena a(ws, (size_t)-1);
  const int64_t B = D.gate == VLPET_GATE_SMALL ? D.M / D.L : 0;
  const bool inl = inline_mode(D);
  w.r8 = r8_of(D.r);
  w.pz = w.r8 + 8;
  if (B) { w.gmean = a.take<float>((size_t)B); if (bwd) w.dgsum = a.take<float>((size_t)B); }
  if (!inl) w.y1 = a.take<__nv_bfloat16>((size_t)D.M * D.d);
  if (!inl && wide_mode(D)) {
    w.wfwd_bytes = wide_adapter_ws(D, false);
    w.wfwd = a.take<char>(w.wfwd_bytes);
  }
  if (bwd) {
    if (inl) {
      w.du = a.take<__nv_bfloat16>((size_t)D.M * D.d);
      w.zs = a.take<__nv_bfloat16>((size_t)D.M * w.pz);
      w.das = a.take<__nv_bfloat16>((size_t)D.M * w.pz);
      if (w.r8 != D.r) { w.pWu = a.take<float>((size_t)D.d * w.r8); w.pWd = a.take<float>((size_t)D.d * w.r8); }
    } else {
      w.dy1 = a.take<__nv_bfloat16>((size_t)D.M * D.d);
      VlpetK2Desc K2;
      memset(&K2, 0, sizeof(K2));
      K2.M = D.M; K2.d = D.d; K2.r = D.r; K2.dtype = D.dtype;
      w.sub_bytes = wide_mode(D) ? wide_adapter_ws(D, true) : fused_k2_bwd_ws(K2);
      w.sub = a.take<char>(w.sub_bytes);
    }
  }
  w.bytes = a.off;
  return w;
}

RowsArgs base_args(const VlpetK1Desc& D, const VlpetK1Params& w) {
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  a.M = D.M; a.L = D.L > 0 ? D.L : 1; a.d = D.d; a.r = D.r; a.gate = D.gate; a.add_gate = D.add_gate; a.pass = 1;
  a.s = D.s; a.alpha = D.alpha; a.kappa = D.kappa;
  a.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  a.inv_keep = a.thr16 ? 1.0f / (1.0f - (float)a.thr16 / 65536.0f) : 1.0f;
  a.seed = D.seed; a.seed_dev = D.seed_dev;
  a.Wd = static_cast<const __nv_bfloat16*>(w.Wd); a.bd = static_cast<const __nv_bfloat16*>(w.bd);
  a.Wu = static_cast<const __nv_bfloat16*>(w.Wu); a.bu = static_cast<const __nv_bfloat16*>(w.bu);
  a.gw = static_cast<const __nv_bfloat16*>(w.gw); a.gb = static_cast<const __nv_bfloat16*>(w.gb);
  a.gz = static_cast<const __nv_bfloat16*>(w.gz);
  return a;
}

}  // namespace

// ---- entry points ------------------------------------------------------------------------------------------------
bool rows_k1_supported(const VlpetK1Desc& D, bool bwd) {
  if (D.dtype != VLPET_BF16 || D.gate == VLPET_GATE_LARGE) return false;
  if (D.d % 256 != 0 || D.d < 256 || D.d > 768 || D.M <= 0) return false;
  if (D.gate == VLPET_GATE_SMALL && (D.L <= 0 || D.M % D.L != 0)) return false;
  if (device_sm_count() <= 0) return false;
  if (inline_mode(D)) return !bwd || wgrad_sm100_supported(D.d, r8_of(D.r));
  if (D.gate == VLPET_GATE_NONE) return false;      // the ungated form at tensor-core ranks IS the fused kernel (K2 form)
  if (wide_mode(D)) return wide_adapter_supported(D, bwd);
  VlpetK1Desc A = adapter_desc(D);
  if (!fused_k1_fwd_supported(A)) return false;
  if (bwd) {
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = D.r; K2.dtype = D.dtype;
    if (!fused_k2_supported(K2)) return false;
  }
  return true;
}
size_t rows_k1_fwd_ws(const VlpetK1Desc& D) { return carve_rows(D, false, nullptr).bytes; }
size_t rows_k1_bwd_ws(const VlpetK1Desc& D) { return carve_rows(D, true, nullptr).bytes; }

int rows_k1_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  RowsWs W = carve_rows(D, false, ws);
  if (W.bytes && (!ws || ws_bytes < W.bytes)) return fail(VLPET_E_WORKSPACE, "k1_fwd(rows): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  const int sms = device_sm_count();
  RowsArgs a = base_args(D, w);
  a.x1 = static_cast<const __nv_bfloat16*>(x1); a.x2 = static_cast<const __nv_bfloat16*>(x2);
  a.out = static_cast<__nv_bfloat16*>(out);
  if (!inline_mode(D)) {
    VlpetK1Params P;
    memset(&P, 0, sizeof(P));
    P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
    if (wide_mode(D)) VLPET_TRY(wide_adapter_fwd(D, x2, P, W.y1, W.wfwd, W.wfwd_bytes, st));
    else VLPET_TRY(fused_k1_fwd(adapter_desc(D), x2, x2, P, W.y1, nullptr, 0, st));
    a.has_y1 = 1; a.y1in = W.y1;
  }
  if (D.gate == VLPET_GATE_SMALL) {
    VLPET_CUDA_OK(cudaMemsetAsync(W.gmean, 0, (size_t)(D.M / D.L) * sizeof(float), st));
    a.gmean = W.gmean;
    a.pass = 0;
    VLPET_TRY(launch_rows_d(false, a, sms, st));
    a.pass = 1;
  }
  return launch_rows_d(false, a, sms, st);
}

int rows_k1_bwd(const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w, void* dx1, void* dx2,
                const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  RowsWs W = carve_rows(D, true, ws);
  if (!ws || ws_bytes < W.bytes) return fail(VLPET_E_WORKSPACE, "k1_bwd(rows): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  const int sms = device_sm_count();
  const bool inl = inline_mode(D);
  RowsArgs a = base_args(D, w);
  a.x1 = static_cast<const __nv_bfloat16*>(x1); a.x2 = static_cast<const __nv_bfloat16*>(x2);
  a.dout = static_cast<const __nv_bfloat16*>(dout);
  a.dx1 = static_cast<__nv_bfloat16*>(dx1); a.dx2 = static_cast<__nv_bfloat16*>(dx2);
  a.dgw = G.dgw; a.dgb = G.dgb; a.dgz = G.dgz; a.dbd = G.dbd;
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  if (!inl) {
    if (wide_mode(D)) VLPET_TRY(wide_adapter_fwd(D, x2, P, W.y1, W.wfwd, W.wfwd_bytes, st));
    else VLPET_TRY(fused_k1_fwd(adapter_desc(D), x2, x2, P, W.y1, nullptr, 0, st));     // y1 again: nothing is saved by the forward
    a.has_y1 = 1; a.y1in = W.y1; a.dy1 = W.dy1;
  } else {
    a.du = W.du; a.zs = W.zs; a.das = W.das; a.pz = W.pz; a.r8 = W.r8;
  }
  if (D.gate == VLPET_GATE_SMALL) {
    const size_t nb = (size_t)(D.M / D.L) * sizeof(float);
    VLPET_CUDA_OK(cudaMemsetAsync(W.gmean, 0, nb, st));
    VLPET_CUDA_OK(cudaMemsetAsync(W.dgsum, 0, nb, st));
    a.gmean = W.gmean; a.dgsum = W.dgsum;
    a.pass = 0;
    VLPET_TRY(launch_rows_d(false, a, sms, st));       // forward pass 0: the per-sample gate mean
    a.pass = 1;
    RowsArgs b0 = a;
    b0.pass = 0;
    VLPET_TRY(launch_rows_d(true, b0, sms, st));       // backward pass 0: dG per sample
  }
  VLPET_TRY(launch_rows_d(true, a, sms, st));
  if (!inl) {
    // adapter backward through the ungated tcgen05 kernels with dout := dy1: dx2 = kappa dy1 + (alpha (dy1 Wu) gelu') Wd
    if (wide_mode(D)) {
      VlpetK1Grads Ga;
      memset(&Ga, 0, sizeof(Ga));
      Ga.dWd = G.dWd; Ga.dbd = G.dbd; Ga.dWu = G.dWu; Ga.dbu = G.dbu;
      return wide_adapter_bwd(D, x2, W.dy1, P, dx2, Ga, W.wfwd, W.sub, W.sub_bytes, st);
    }
    VlpetK2Desc K2;
    memset(&K2, 0, sizeof(K2));
    K2.M = D.M; K2.d = D.d; K2.r = D.r; K2.dtype = D.dtype; K2.sf = D.alpha;
    VlpetK2Params P2;
    P2.Wd = w.Wd; P2.bd = w.bd; P2.Wu = w.Wu; P2.bu = w.bu;
    VlpetK2Grads G2;
    G2.dWd = G.dWd; G2.dbd = G.dbd; G2.dWu = G.dWu; G2.dbu = G.dbu;
    return fused_k2_bwd_kappa(K2, D.kappa, x2, W.dy1, P2, dx2, G2, W.sub, W.sub_bytes, st);
  }
  // INLINE: dWu = du^T z (+ dbu through the ones column), dWd = (x2^T da)^T, dbd = column sums of da
  const int r8 = W.r8;
  float *oWu = G.dWu, *oWd = G.dWd;
  if (r8 != D.r) {
    VLPET_CUDA_OK(cudaMemsetAsync(W.pWu, 0, (size_t)D.d * r8 * sizeof(float), st));
    VLPET_CUDA_OK(cudaMemsetAsync(W.pWd, 0, (size_t)D.d * r8 * sizeof(float), st));
    oWu = G.dWu ? W.pWu : nullptr;
    oWd = G.dWd ? W.pWd : nullptr;
  }
  const void* A[2]; const void* Bm[2]; int64_t lda[2], ldb[2]; int nbv[2], tr[2]; float* out[2]; float* bias[2]; float sc[2];
  int k = 0;
  if (oWu || G.dbu) {
    if (!oWu) return fail(VLPET_E_BADARG, "k1_bwd(rows): a bias gradient needs its weight gradient buffer");
    A[k] = W.du; lda[k] = D.d; Bm[k] = W.zs; ldb[k] = W.pz; nbv[k] = r8 + 1; tr[k] = 0; out[k] = oWu; bias[k] = G.dbu; sc[k] = 1.f; ++k;
  }
  if (oWd) {
    A[k] = x2; lda[k] = D.d; Bm[k] = W.das; ldb[k] = W.pz; nbv[k] = r8; tr[k] = 1; out[k] = oWd; bias[k] = nullptr; sc[k] = 1.f; ++k;
  }
  if (k) VLPET_TRY(wgrad_sm100(k, A, lda, Bm, ldb, nbv, out, bias, sc, tr, D.M, D.d, r8, sms, st));
  if (r8 != D.r && (G.dWu || G.dWd)) {
    const int n = D.d * D.r;
    unpad_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(W.pWu, W.pWd, G.dWu, G.dWd, D.d, D.r, r8);
    VLPET_LAUNCH_OK();
  }
  return 0;
}

}  // namespace vlpet
