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

constexpr int RW_THREADS = 256;   // 8 warps = 8 rows in flight per CTA, 2 CTAs per SM
constexpr int RW_WARPS = RW_THREADS / 32;
constexpr int RW_MAXR = 16;
constexpr int RW_NS = 2;          // row stages per warp: the row being processed + the next one in flight

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

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 8 consecutive bf16 at a 16-byte aligned shared-memory address -> floats (one LDS.128 + 8 ALU)
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&w)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { w[2 * e] = __uint_as_float(u[e] << 16); w[2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u); }
}
__device__ __forceinline__ void stg8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf(v[0], v[1]); q.y = pack_bf(v[2], v[3]); q.z = pack_bf(v[4], v[5]); q.w = pack_bf(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8], float acc) {
#pragma unroll
  for (int e = 0; e < 8; ++e) acc = fmaf(a[e], b[e], acc);
  return acc;
}
// all N sums at once: N independent shuffle chains per step instead of N serial 5-step reductions
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Shared-memory plan of one CTA (dynamic):
//   [Wd r x d][Wu^T r x d]   bf16, INLINE only (staged once per CTA)
//   [bu d][gv 2d]            bf16: up-projection bias (INLINE), gate vector(s) (middleX gw | middleY gz | small gw_x, gw_y)
//   [GW 16]                  fp32: <gate vector of the y1 half, Wu^T_j>  (INLINE backward of middleX / small)
//   rows: RW_WARPS x RW_NS x NT x d bf16 -- every warp owns RW_NS row stages of NT tensors (x1, x2 | y1 [, dout]); a lane
//         copies exactly the 16-byte pieces it later reads (cp.async, no cross-lane traffic, no barrier), so the bytes in
//         flight do not depend on the register budget: 8 warps x 2 CTAs x one 3-4.6 KB row ahead = 48-74 KB per SM.
//   The block reduction of the gate gradients at the end of the backward reuses the row region.
struct SmemPlan {
  __nv_bfloat16 *Wd, *WuT, *bu, *gv;
  float* GW;
  __nv_bfloat16* rows;
};
__host__ __device__ inline size_t rows_smem_bytes(int d, int r_inline, int nt) {
  size_t s = (size_t)2 * r_inline * d * 2 + (size_t)3 * d * 2 + RW_MAXR * 4;
  s = (s + 15) / 16 * 16;
  return s + (size_t)RW_WARPS * RW_NS * nt * d * 2;
}
__device__ __forceinline__ SmemPlan carve_smem(uint8_t* smem, int d, int r_inline) {
  SmemPlan P;
  P.Wd = reinterpret_cast<__nv_bfloat16*>(smem);
  P.WuT = P.Wd + (size_t)r_inline * d;
  P.bu = P.WuT + (size_t)r_inline * d;
  P.gv = P.bu + d;
  P.GW = reinterpret_cast<float*>(P.gv + 2 * d);
  size_t off = (size_t)2 * r_inline * d * 2 + (size_t)3 * d * 2 + RW_MAXR * 4;
  off = (off + 15) / 16 * 16;
  P.rows = reinterpret_cast<__nv_bfloat16*>(smem + off);
  return P;
}

// stage the per-CTA constants; RMAX = 0: COMPOSED mode (no adapter weights)
template <int RMAX, int GATE>
__device__ __forceinline__ void stage_constants(const RowsArgs& p, const SmemPlan& S, bool bwd) {
  if constexpr (RMAX > 0) {
    const int n = p.r * p.d;
    for (int i = threadIdx.x; i < n; i += RW_THREADS) {
      S.Wd[i] = p.Wd[i];
      const int j = i / p.d, c = i % p.d;
      S.WuT[i] = p.Wu[(size_t)c * p.r + j];      // Wu is [d, r]: transposed into [r, d] rows
    }
    for (int i = threadIdx.x; i < p.d; i += RW_THREADS) S.bu[i] = p.bu[i];
  }
  if (GATE == VLPET_GATE_MIDDLE_X) for (int i = threadIdx.x; i < p.d; i += RW_THREADS) S.gv[i] = p.gw[i];
  if (GATE == VLPET_GATE_MIDDLE_Y) for (int i = threadIdx.x; i < p.d; i += RW_THREADS) S.gv[i] = p.gz[i];
  if (GATE == VLPET_GATE_SMALL) for (int i = threadIdx.x; i < 2 * p.d; i += RW_THREADS) S.gv[i] = p.gw[i];
  __syncthreads();
  if (RMAX > 0 && bwd && (GATE == VLPET_GATE_MIDDLE_X || GATE == VLPET_GATE_SMALL)) {
    // GW_j = sum_c gv_y[c] Wu^T[j][c]: the part of dz_j that is proportional to the row's gate scalar gradient dt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const __nv_bfloat16* gy = S.gv + (GATE == VLPET_GATE_SMALL ? p.d : 0);
    for (int j = warp; j < p.r; j += RW_WARPS) {
      float acc = 0.f;
      for (int c = lane; c < p.d; c += 32) acc = fmaf(__bfloat162float(gy[c]), __bfloat162float(S.WuT[(size_t)j * p.d + c]), acc);
      acc = warp_sum(acc);
      if (lane == 0) S.GW[j] = acc;
    }
    __syncthreads();
  }
}

// issue the asynchronous copy of one row (NT tensors) into a stage; every lane copies its own NCH pieces per tensor
// (SKIP_X1: the backward of the gates without a per-row scalar -- none, middleY -- never reads x1)
template <int NCH, int NT, bool SKIP_X1 = false>
__device__ __forceinline__ void issue_row(const RowsArgs& p, int64_t row, int lane, __nv_bfloat16* stage) {
  const __nv_bfloat16* src[3] = {p.x1 + row * p.d, (p.has_y1 ? p.y1in : p.x2) + row * p.d, NT > 2 ? p.dout + row * p.d : nullptr};
#pragma unroll
  for (int t = SKIP_X1 ? 1 : 0; t < NT; ++t)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) cp_async16(stage + (size_t)t * p.d + ch * 256 + lane * 8, src[t] + ch * 256 + lane * 8);
}

// Forward quantities of one row: y1 (kept in registers), the adapter pre-activations a_j / z_j (INLINE), the gate scalars.
template <int NCH, int RMAX>
struct RowFwd {
  float y1[NCH * 8];
  float a[RMAX > 0 ? RMAX : 1], z[RMAX > 0 ? RMAX : 1];
  float g, sg;          // middleX: g = sigmoid of the row's dot; small: sg = this token's sigmoid, g = the sample's mean gate
};

template <int NCH, int GATE, int RMAX>
__device__ __forceinline__ void row_forward(const RowsArgs& p, const SmemPlan& S, const __nv_bfloat16* sx1, const __nv_bfloat16* sx2,
                                            int lane, float gbias, RowFwd<NCH, RMAX>& F) {
  if constexpr (RMAX > 0) {
    // ---- sweep A: a_j = <x2, Wd_j> + bd_j for all ranks, one batched reduction
    float acc[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float x2c[8];
      lds8(sx2 + ch * 256 + lane * 8, x2c);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        if (j < p.r) {
          float w[8];
          lds8(S.Wd + (size_t)j * p.d + ch * 256 + lane * 8, w);
          acc[j] = dot8(x2c, w, acc[j]);
        }
      }
    }
    warp_sum_n<RMAX>(acc);
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      F.a[j] = acc[j] + (j < p.r ? __bfloat162float(p.bd[j]) : 0.f);
      F.z[j] = j < p.r ? gelu_new_f(F.a[j]) : 0.f;
    }
  }
  // ---- sweep B: y1 chunks (INLINE: kappa x2 + alpha (sum_j z_j Wu^T_j + bu); COMPOSED: loaded) and the gate's dot product
  float t = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    float y[8];
    if constexpr (RMAX > 0) {
      float x2c[8];
      lds8(sx2 + ch * 256 + lane * 8, x2c);
      lds8(S.bu + ch * 256 + lane * 8, y);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        if (j < p.r) {
          float w[8];
          lds8(S.WuT + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = fmaf(F.z[j], w[e], y[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = p.kappa * x2c[e] + p.alpha * y[e];
    } else {
      lds8(sx2 + ch * 256 + lane * 8, y);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) F.y1[ch * 8 + e] = y[e];
    if (GATE == VLPET_GATE_MIDDLE_X || GATE == VLPET_GATE_SMALL) {
      float x1c[8], gx[8];
      lds8(sx1 + ch * 256 + lane * 8, x1c);
      lds8(S.gv + ch * 256 + lane * 8, gx);
      if (GATE == VLPET_GATE_MIDDLE_X) {
#pragma unroll
        for (int e = 0; e < 8; ++e) t = fmaf(x1c[e] + y[e], gx[e], t);
      } else {
        float gy[8];
        lds8(S.gv + p.d + ch * 256 + lane * 8, gy);
#pragma unroll
        for (int e = 0; e < 8; ++e) t = fmaf(x1c[e], gx[e], fmaf(y[e], gy[e], t));
      }
    }
  }
  F.g = 1.f;
  F.sg = 0.f;
  if (GATE == VLPET_GATE_MIDDLE_X) F.g = sigmoid_f(warp_sum(t) + gbias);
  if (GATE == VLPET_GATE_SMALL) F.sg = sigmoid_f(warp_sum(t) + gbias);
}

template <int NCH, int GATE, int RMAX>
__global__ void __launch_bounds__(RW_THREADS, 2) rows_fwd_kernel(const RowsArgs p) {
  constexpr int NT = 2;
  extern __shared__ __align__(16) uint8_t smem[];
  const SmemPlan S = carve_smem(smem, p.d, RMAX > 0 ? p.r : 0);
  stage_constants<RMAX, GATE>(p, S, false);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float gbias = (GATE == VLPET_GATE_MIDDLE_X || GATE == VLPET_GATE_SMALL) ? __bfloat162float(p.gb[0]) : 0.f;
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * RW_WARPS;
  __nv_bfloat16* const wrows = S.rows + (size_t)warp * RW_NS * NT * p.d;
  int64_t row = (int64_t)blockIdx.x * RW_WARPS + warp;
  if (row < p.M) issue_row<NCH, NT>(p, row, lane, wrows);
  cp_async_commit();
  for (int it = 0; row < p.M; ++it, row += nw) {
    if (row + nw < p.M) issue_row<NCH, NT>(p, row + nw, lane, wrows + (size_t)((it + 1) & 1) * NT * p.d);
    cp_async_commit();
    cp_async_wait1();
    const __nv_bfloat16* sx1 = wrows + (size_t)(it & 1) * NT * p.d;
    const __nv_bfloat16* sx2 = sx1 + p.d;
    RowFwd<NCH, RMAX> F;
    row_forward<NCH, GATE, RMAX>(p, S, sx1, sx2, lane, gbias, F);
    if (GATE == VLPET_GATE_SMALL) {
      if (p.pass == 0) {
        if (lane == 0) atomicAdd(p.gmean + row / p.L, F.sg / (float)p.L);
        continue;
      }
      F.g = p.gmean[row / p.L];
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float x1c[8], m[8], o[8], gz[8];
      lds8(sx1 + ch * 256 + lane * 8, x1c);
      if (GATE == VLPET_GATE_MIDDLE_Y) lds8(S.gv + ch * 256 + lane * 8, gz);
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + ch * 256 + lane * 8, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = F.y1[ch * 8 + e];
        float h;
        if (GATE == VLPET_GATE_MIDDLE_Y) h = p.add_gate ? y + 1.f + gz[e] : y * (1.f + gz[e]);
        else if (GATE == VLPET_GATE_NONE) h = y;
        else h = p.add_gate ? y + F.g : y * F.g;
        o[e] = x1c[e] + p.s * m[e] * h;
      }
      stg8(p.out + row * p.d + ch * 256 + lane * 8, o);
    }
  }
}

// Backward.  pass 0 (small gate only): the per-sample reductions both directions need before any output can be formed --
// the gate mean (sum of the tokens' sigmoids) AND dG (sum over tokens and columns of dh (y1)) -- in ONE sweep over x1, x2, dout.
template <int NCH, int GATE, int RMAX>
__global__ void __launch_bounds__(RW_THREADS, 2) rows_bwd_kernel(const RowsArgs p) {
  constexpr int NT = 3;
  constexpr int NE = NCH * 8;
  constexpr bool ROWGATE = GATE == VLPET_GATE_MIDDLE_X || GATE == VLPET_GATE_SMALL;
  extern __shared__ __align__(16) uint8_t smem[];
  const SmemPlan S = carve_smem(smem, p.d, RMAX > 0 ? p.r : 0);
  stage_constants<RMAX, GATE>(p, S, true);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float gbias = ROWGATE ? __bfloat162float(p.gb[0]) : 0.f;
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * RW_WARPS;
  __nv_bfloat16* const wrows = S.rows + (size_t)warp * RW_NS * NT * p.d;
  // gate-parameter gradients of the rows this warp handles (lanes own distinct columns): middleX / middleY one vector, small two
  float acc_g[GATE == VLPET_GATE_SMALL ? 2 * NE : (GATE == VLPET_GATE_NONE ? 1 : NE)];
#pragma unroll
  for (int e = 0; e < (int)(sizeof(acc_g) / sizeof(float)); ++e) acc_g[e] = 0.f;
  float acc_b = 0.f;
  float acc_da[RMAX > 0 ? RMAX : 1];      // dbd = sum over rows of da, kept in fp32 (the bf16 da scratch only feeds the GEMM)
#pragma unroll
  for (int j = 0; j < (RMAX > 0 ? RMAX : 1); ++j) acc_da[j] = 0.f;

  int64_t row = (int64_t)blockIdx.x * RW_WARPS + warp;
  if (row < p.M) issue_row<NCH, NT, !ROWGATE>(p, row, lane, wrows);
  cp_async_commit();
  for (int it = 0; row < p.M; ++it, row += nw) {
    if (row + nw < p.M) issue_row<NCH, NT, !ROWGATE>(p, row + nw, lane, wrows + (size_t)((it + 1) & 1) * NT * p.d);
    cp_async_commit();
    cp_async_wait1();
    const __nv_bfloat16* sx1 = wrows + (size_t)(it & 1) * NT * p.d;
    const __nv_bfloat16* sx2 = sx1 + p.d;
    const __nv_bfloat16* sdo = sx2 + p.d;
    RowFwd<NCH, RMAX> F;
    row_forward<NCH, GATE, RMAX>(p, S, sx1, sx2, lane, gbias, F);
    if (GATE == VLPET_GATE_SMALL && p.pass != 0) F.g = p.gmean[row / p.L];
    // ---- sweep C: dh = s m dout; the row's dG contribution; the dt-independent part of dz_j (INLINE)
    //      base = dy1 without the gate-scalar term: dh (add gate) | dh g (middleX / small) | dh (1 + gz) (middleY) | dh (none)
    uint32_t mbits = 0;
    float dgs = 0.f;
    float dzb[RMAX > 0 ? RMAX : 1];
#pragma unroll
    for (int j = 0; j < (RMAX > 0 ? RMAX : 1); ++j) dzb[j] = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float doc[8], m[8], base[8], gz[8];
      lds8(sdo + ch * 256 + lane * 8, doc);
      if (GATE == VLPET_GATE_MIDDLE_Y) lds8(S.gv + ch * 256 + lane * 8, gz);
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + ch * 256 + lane * 8, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (m[e] != 0.f) mbits |= 1u << (ch * 8 + e);
        const float dh = p.s * m[e] * doc[e];
        dgs += p.add_gate ? dh : dh * F.y1[ch * 8 + e];
        if (p.add_gate || GATE == VLPET_GATE_NONE) base[e] = dh;
        else if (GATE == VLPET_GATE_MIDDLE_Y) base[e] = dh * (1.f + gz[e]);
        else base[e] = dh * F.g;
      }
      if constexpr (RMAX > 0) {
        if (!(GATE == VLPET_GATE_SMALL && p.pass == 0)) {
#pragma unroll
        for (int j = 0; j < RMAX; ++j) {
          if (j < p.r) {
            float w[8];
            lds8(S.WuT + (size_t)j * p.d + ch * 256 + lane * 8, w);
            dzb[j] = dot8(base, w, dzb[j]);
          }
        }
        }
      }
    }
    if (GATE == VLPET_GATE_SMALL && p.pass == 0) {
      dgs = warp_sum(dgs);
      if (lane == 0) {
        atomicAdd(p.gmean + row / p.L, F.sg / (float)p.L);
        atomicAdd(p.dgsum + row / p.L, dgs);
      }
      continue;
    }
    // ---- the row's gate-scalar gradient: middleX dt = dG g (1 - g) with dG the row sum; small dt = dG_sample / L * sg (1 - sg)
    float dt = 0.f;
    if (GATE == VLPET_GATE_MIDDLE_X) dt = warp_sum(dgs) * F.g * (1.f - F.g);
    if (GATE == VLPET_GATE_SMALL) dt = p.dgsum[row / p.L] / (float)p.L * F.sg * (1.f - F.sg);
    acc_b += dt;
    // ---- INLINE adapter: dz_j = alpha (dzb_j + dt GW_j), da_j = dz_j gelu_new'(a_j); z / da rows for the weight-gradient GEMM
    float da[RMAX > 0 ? RMAX : 1];
    if constexpr (RMAX > 0) {
      warp_sum_n<RMAX>(dzb);
      __nv_bfloat16* zrow = p.zs + row * p.pz;
      __nv_bfloat16* darow = p.das + row * p.pz;
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        const float dz = p.alpha * (dzb[j] + (ROWGATE ? dt * S.GW[j < p.r ? j : 0] : 0.f));
        da[j] = j < p.r ? dz * gelu_new_grad_f(F.a[j]) : 0.f;
        acc_da[j] += da[j];
      }
      if (lane < p.pz) {     // columns [0, r): z / da; column r8: the ones column (bias-gradient trick of the GEMM); the rest zero
        float zv = 0.f, dv = 0.f;
#pragma unroll
        for (int j = 0; j < RMAX; ++j)
          if (lane == j) { zv = F.z[j]; dv = da[j]; }
        if (lane >= p.r) { zv = lane == p.r8 ? 1.f : 0.f; dv = 0.f; }
        zrow[lane] = __float2bfloat16_rn(zv);
        darow[lane] = __float2bfloat16_rn(dv);
      }
    }
    // ---- sweep D: dy1 = base + dt gv_y, dx1 = dout + dt gv_x, gate-parameter gradients; COMPOSED: dy1 out; INLINE: du, dx2 out
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float doc[8], x1c[8], gx[8], gy[8], dy1[8], dx1[8];
      lds8(sdo + ch * 256 + lane * 8, doc);
      if (GATE != VLPET_GATE_NONE) lds8(S.gv + ch * 256 + lane * 8, gx);
      if (GATE == VLPET_GATE_SMALL) lds8(S.gv + p.d + ch * 256 + lane * 8, gy);
      if (ROWGATE) lds8(sx1 + ch * 256 + lane * 8, x1c);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = ch * 8 + e;
        const float dh = ((mbits >> i) & 1u) ? p.s * p.inv_keep * doc[e] : 0.f;
        const float y = F.y1[i];
        float base;
        if (p.add_gate || GATE == VLPET_GATE_NONE) base = dh;
        else if (GATE == VLPET_GATE_MIDDLE_Y) base = dh * (1.f + gx[e]);
        else base = dh * F.g;
        dx1[e] = doc[e];
        dy1[e] = base;
        if (GATE == VLPET_GATE_MIDDLE_Y) acc_g[i] += p.add_gate ? dh : dh * y;
        if (GATE == VLPET_GATE_MIDDLE_X) {
          acc_g[i] += dt * (x1c[e] + y);
          dy1[e] = base + dt * gx[e];
          dx1[e] = doc[e] + dt * gx[e];
        }
        if (GATE == VLPET_GATE_SMALL) {
          acc_g[i] += dt * x1c[e];
          acc_g[NE + i] += dt * y;
          dy1[e] = base + dt * gy[e];
          dx1[e] = doc[e] + dt * gx[e];
        }
      }
      if (p.dx1) stg8(p.dx1 + row * p.d + ch * 256 + lane * 8, dx1);   // (K2 through the ungated form: dx1 = dout, not written)
      if constexpr (RMAX == 0) {
        stg8(p.dy1 + row * p.d + ch * 256 + lane * 8, dy1);      // the ungated tcgen05 backward takes it from here
      } else {
        float du[8], dx2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { du[e] = p.alpha * dy1[e]; dx2[e] = p.kappa * dy1[e]; }
#pragma unroll
        for (int j = 0; j < RMAX; ++j) {
          if (j < p.r) {
            float w[8];
            lds8(S.Wd + (size_t)j * p.d + ch * 256 + lane * 8, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) dx2[e] = fmaf(da[j], w[e], dx2[e]);
          }
        }
        stg8(p.dx2 + row * p.d + ch * 256 + lane * 8, dx2);
        stg8(p.du + row * p.d + ch * 256 + lane * 8, du);
      }
    }
  }
  if (GATE == VLPET_GATE_SMALL && p.pass == 0) return;
  if constexpr (RMAX > 0) {
    if (p.dbd && lane == 0) {
#pragma unroll
      for (int j = 0; j < RMAX; ++j)
        if (j < p.r) atomicAdd(p.dbd + j, acc_da[j]);
    }
  }
  // ---- gate-parameter gradients: lanes own distinct columns; reduce over the CTA's warps in shared memory (the row region
  //      is idle now), then one atomicAdd per column and CTA
  if (GATE == VLPET_GATE_NONE) return;
  const int nvec = (GATE == VLPET_GATE_SMALL ? 2 : 1) * p.d;
  float* sred = reinterpret_cast<float*>(S.rows);
  __syncthreads();
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

// ---- large gate at ranks too small for a tensor-core tile (r, rg <= 16; e.g. r = 4 in the SURVEY 8(d) sweep) ---------------
// G = sigmoid(gelu_new(x1 Gd^T + gbd) Gu^T + gbu) is a second skinny adapter on x1 (my_transformers/modeling_bart.py:1195-1209):
// the same row-local structure, two branches.  The backward writes du / dT and z / q / da / dp rows for the weight-gradient GEMM.
struct LargeArgs {
  RowsArgs a;                                 // adapter branch + common fields (gate / gw / gz unused)
  int rg, rg8, pq;
  const __nv_bfloat16 *Gd, *gbd, *Gu, *gbu;
  __nv_bfloat16 *dt, *qs, *dps;               // [M, d] dT scratch, [M, pq] q | 1 | 0.. and dp rows
  float* dgbd;
};
struct LargePlan { __nv_bfloat16 *Wd, *WuT, *Gd, *GuT, *bu, *gbu, *rows; };
__host__ __device__ inline size_t large_smem_bytes(int d, int r, int rg, int nt) {
  size_t s = (size_t)2 * (r + rg) * d * 2 + (size_t)2 * d * 2;
  s = (s + 15) / 16 * 16;
  return s + (size_t)RW_WARPS * RW_NS * nt * d * 2;
}
__device__ __forceinline__ LargePlan carve_large(uint8_t* smem, int d, int r, int rg) {
  LargePlan P;
  P.Wd = reinterpret_cast<__nv_bfloat16*>(smem);
  P.WuT = P.Wd + (size_t)r * d;
  P.Gd = P.WuT + (size_t)r * d;
  P.GuT = P.Gd + (size_t)rg * d;
  P.bu = P.GuT + (size_t)rg * d;
  P.gbu = P.bu + d;
  size_t off = (size_t)2 * (r + rg) * d * 2 + (size_t)2 * d * 2;
  off = (off + 15) / 16 * 16;
  P.rows = reinterpret_cast<__nv_bfloat16*>(smem + off);
  return P;
}
__device__ __forceinline__ void stage_large(const LargeArgs& q, const LargePlan& S) {
  const RowsArgs& p = q.a;
  for (int i = threadIdx.x; i < p.r * p.d; i += RW_THREADS) {
    S.Wd[i] = p.Wd[i];
    S.WuT[i] = p.Wu[(size_t)(i % p.d) * p.r + i / p.d];
  }
  for (int i = threadIdx.x; i < q.rg * p.d; i += RW_THREADS) {
    S.Gd[i] = q.Gd[i];
    S.GuT[i] = q.Gu[(size_t)(i % p.d) * q.rg + i / p.d];
  }
  for (int i = threadIdx.x; i < p.d; i += RW_THREADS) { S.bu[i] = p.bu[i]; S.gbu[i] = q.gbu[i]; }
  __syncthreads();
}
// acc[j] += <row chunk, W_j chunk> over all chunks of a row held in shared memory
template <int NCH, int RMAX>
__device__ __forceinline__ void branch_dots(const __nv_bfloat16* srow, const __nv_bfloat16* sW, int r, int d, int lane, float (&acc)[RMAX]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    float xc[8];
    lds8(srow + ch * 256 + lane * 8, xc);
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      if (j < r) {
        float w[8];
        lds8(sW + (size_t)j * d + ch * 256 + lane * 8, w);
        acc[j] = dot8(xc, w, acc[j]);
      }
    }
  }
}
// y = bias_c + sum_j z_j W^T_j,c for one 8-element chunk piece
template <int RMAX>
__device__ __forceinline__ void up_piece(const __nv_bfloat16* sWT, const __nv_bfloat16* sbias, const float (&z)[RMAX], int r, int d, int off,
                                         float (&y)[8]) {
  lds8(sbias + off, y);
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    if (j < r) {
      float w[8];
      lds8(sWT + (size_t)j * d + off, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaf(z[j], w[e], y[e]);
    }
  }
}
// forward quantities of a row: y1 and G per element (registers), pre-activations of both branches
template <int NCH, int RMAX>
struct LargeFwd { float y1[NCH * 8], G[NCH * 8], a[RMAX], z[RMAX], pa[RMAX], q[RMAX]; };
template <int NCH, int RMAX>
__device__ __forceinline__ void large_forward(const LargeArgs& q, const LargePlan& S, const __nv_bfloat16* sx1, const __nv_bfloat16* sx2,
                                              int lane, LargeFwd<NCH, RMAX>& F) {
  const RowsArgs& p = q.a;
  float acc[2 * RMAX];
#pragma unroll
  for (int j = 0; j < 2 * RMAX; ++j) acc[j] = 0.f;
  {
    float (&aa)[RMAX] = *reinterpret_cast<float (*)[RMAX]>(&acc[0]);
    float (&pp)[RMAX] = *reinterpret_cast<float (*)[RMAX]>(&acc[RMAX]);
    branch_dots<NCH, RMAX>(sx2, S.Wd, p.r, p.d, lane, aa);
    branch_dots<NCH, RMAX>(sx1, S.Gd, q.rg, p.d, lane, pp);
  }
  warp_sum_n<2 * RMAX>(acc);
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    F.a[j] = acc[j] + (j < p.r ? __bfloat162float(p.bd[j]) : 0.f);
    F.z[j] = j < p.r ? gelu_new_f(F.a[j]) : 0.f;
    F.pa[j] = acc[RMAX + j] + (j < q.rg ? __bfloat162float(q.gbd[j]) : 0.f);
    F.q[j] = j < q.rg ? gelu_new_f(F.pa[j]) : 0.f;
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int off = ch * 256 + lane * 8;
    float x2c[8], y[8], t[8];
    lds8(sx2 + off, x2c);
    up_piece<RMAX>(S.WuT, S.bu, F.z, p.r, p.d, off, y);
    up_piece<RMAX>(S.GuT, S.gbu, F.q, q.rg, p.d, off, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      F.y1[ch * 8 + e] = p.kappa * x2c[e] + p.alpha * y[e];
      F.G[ch * 8 + e] = sigmoid_f(t[e]);
    }
  }
}

template <int NCH, int RMAX>
__global__ void __launch_bounds__(RW_THREADS, 2) rows_large_fwd_kernel(const LargeArgs q) {
  constexpr int NT = 2;
  const RowsArgs& p = q.a;
  extern __shared__ __align__(16) uint8_t smem[];
  const LargePlan S = carve_large(smem, p.d, p.r, q.rg);
  stage_large(q, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * RW_WARPS;
  __nv_bfloat16* const wrows = S.rows + (size_t)warp * RW_NS * NT * p.d;
  int64_t row = (int64_t)blockIdx.x * RW_WARPS + warp;
  if (row < p.M) issue_row<NCH, NT>(p, row, lane, wrows);
  cp_async_commit();
  for (int it = 0; row < p.M; ++it, row += nw) {
    if (row + nw < p.M) issue_row<NCH, NT>(p, row + nw, lane, wrows + (size_t)((it + 1) & 1) * NT * p.d);
    cp_async_commit();
    cp_async_wait1();
    const __nv_bfloat16* sx1 = wrows + (size_t)(it & 1) * NT * p.d;
    LargeFwd<NCH, RMAX> F;
    large_forward<NCH, RMAX>(q, S, sx1, sx1 + p.d, lane, F);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float x1c[8], m[8], o[8];
      lds8(sx1 + ch * 256 + lane * 8, x1c);
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + ch * 256 + lane * 8, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = F.y1[ch * 8 + e], g = F.G[ch * 8 + e];
        o[e] = x1c[e] + p.s * m[e] * (p.add_gate ? y + g : y * g);
      }
      stg8(p.out + row * p.d + ch * 256 + lane * 8, o);
    }
  }
}

template <int NCH, int RMAX>
__global__ void __launch_bounds__(RW_THREADS, 2) rows_large_bwd_kernel(const LargeArgs q) {
  constexpr int NT = 3;
  const RowsArgs& p = q.a;
  extern __shared__ __align__(16) uint8_t smem[];
  const LargePlan S = carve_large(smem, p.d, p.r, q.rg);
  stage_large(q, S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
  const int64_t nw = (int64_t)gridDim.x * RW_WARPS;
  __nv_bfloat16* const wrows = S.rows + (size_t)warp * RW_NS * NT * p.d;
  float acc_da[RMAX], acc_dp[RMAX];       // dbd / dgbd in fp32
#pragma unroll
  for (int j = 0; j < RMAX; ++j) acc_da[j] = acc_dp[j] = 0.f;
  int64_t row = (int64_t)blockIdx.x * RW_WARPS + warp;
  if (row < p.M) issue_row<NCH, NT>(p, row, lane, wrows);
  cp_async_commit();
  for (int it = 0; row < p.M; ++it, row += nw) {
    if (row + nw < p.M) issue_row<NCH, NT>(p, row + nw, lane, wrows + (size_t)((it + 1) & 1) * NT * p.d);
    cp_async_commit();
    cp_async_wait1();
    const __nv_bfloat16* sx1 = wrows + (size_t)(it & 1) * NT * p.d;
    const __nv_bfloat16* sx2 = sx1 + p.d;
    const __nv_bfloat16* sdo = sx2 + p.d;
    LargeFwd<NCH, RMAX> F;
    large_forward<NCH, RMAX>(q, S, sx1, sx2, lane, F);
    // ---- sweep C: dh = s m dout; dy1 = dh G (add gate: dh), dT = dh y1 G (1 - G) (add gate: dh G (1 - G)); du / dT rows out;
    //      dz_j = alpha <dy1, Wu^T_j>, dq_j = <dT, Gu^T_j>.  F.y1 / F.G are overwritten by dy1 / dT (sweep D needs only those).
    float dzq[2 * RMAX];
#pragma unroll
    for (int j = 0; j < 2 * RMAX; ++j) dzq[j] = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int off = ch * 256 + lane * 8;
      float doc[8], m[8], dy1[8], dT[8], du[8];
      lds8(sdo + off, doc);
      drop_scale8(seed_eff, p.thr16, p.inv_keep, row * p.d + off, m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dh = p.s * m[e] * doc[e], y = F.y1[ch * 8 + e], g = F.G[ch * 8 + e], gg = g * (1.f - g);
        dy1[e] = p.add_gate ? dh : dh * g;
        dT[e] = p.add_gate ? dh * gg : dh * y * gg;
        du[e] = p.alpha * dy1[e];
        F.y1[ch * 8 + e] = dy1[e];
        F.G[ch * 8 + e] = dT[e];
      }
      stg8(p.du + row * p.d + off, du);
      stg8(q.dt + row * p.d + off, dT);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        float w[8];
        if (j < p.r) { lds8(S.WuT + (size_t)j * p.d + off, w); dzq[j] = dot8(du, w, dzq[j]); }
        if (j < q.rg) { lds8(S.GuT + (size_t)j * p.d + off, w); dzq[RMAX + j] = dot8(dT, w, dzq[RMAX + j]); }
      }
    }
    warp_sum_n<2 * RMAX>(dzq);
    float da[RMAX], dp[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      da[j] = j < p.r ? dzq[j] * gelu_new_grad_f(F.a[j]) : 0.f;
      dp[j] = j < q.rg ? dzq[RMAX + j] * gelu_new_grad_f(F.pa[j]) : 0.f;
      acc_da[j] += da[j];
      acc_dp[j] += dp[j];
    }
    {   // z | 1 | 0.. and da rows (pitch pz), q | 1 | 0.. and dp rows (pitch pq): the ones column sits at index r8 / rg8
      float zv = 0.f, dv = 0.f, qv = 0.f, pv = 0.f;
#pragma unroll
      for (int j = 0; j < RMAX; ++j)
        if (lane == j) { zv = F.z[j]; dv = da[j]; qv = F.q[j]; pv = dp[j]; }
      if (lane < p.pz) {
        if (lane >= p.r) { zv = lane == p.r8 ? 1.f : 0.f; dv = 0.f; }
        p.zs[row * p.pz + lane] = __float2bfloat16_rn(zv);
        p.das[row * p.pz + lane] = __float2bfloat16_rn(dv);
      }
      if (lane < q.pq) {
        if (lane >= q.rg) { qv = lane == q.rg8 ? 1.f : 0.f; pv = 0.f; }
        q.qs[row * q.pq + lane] = __float2bfloat16_rn(qv);
        q.dps[row * q.pq + lane] = __float2bfloat16_rn(pv);
      }
    }
    // ---- sweep D: dx2 = kappa dy1 + sum_j da_j Wd_j, dx1 = dout + sum_j dp_j Gd_j
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int off = ch * 256 + lane * 8;
      float dx1[8], dx2[8];
      lds8(sdo + off, dx1);
#pragma unroll
      for (int e = 0; e < 8; ++e) dx2[e] = p.kappa * F.y1[ch * 8 + e];
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        float w[8];
        if (j < p.r) {
          lds8(S.Wd + (size_t)j * p.d + off, w);
#pragma unroll
          for (int e = 0; e < 8; ++e) dx2[e] = fmaf(da[j], w[e], dx2[e]);
        }
        if (j < q.rg) {
          lds8(S.Gd + (size_t)j * p.d + off, w);
#pragma unroll
          for (int e = 0; e < 8; ++e) dx1[e] = fmaf(dp[j], w[e], dx1[e]);
        }
      }
      stg8(p.dx1 + row * p.d + off, dx1);
      stg8(p.dx2 + row * p.d + off, dx2);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      if (p.dbd && j < p.r) atomicAdd(p.dbd + j, acc_da[j]);
      if (q.dgbd && j < q.rg) atomicAdd(q.dgbd + j, acc_dp[j]);
    }
  }
}

// add the first r columns / rows of the zero-padded weight-gradient buffers (rank padded to 8 for the GEMM) into the real ones
__global__ void unpad_add_kernel(const float* __restrict__ pWu, const float* __restrict__ pWd, float* dWu, float* dWd, int d, int r, int r8) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d * r) return;
  if (dWu) { const int c = i / r, j = i % r; dWu[i] += pWu[c * r8 + j]; }
  if (dWd) dWd[i] += pWd[i];                       // [r8, d] row-major: the first r rows are the first r*d elements
}

// rank bucket of the INLINE kernels (compile-time loop bound; the loops skip j >= r): 0 = COMPOSED (y1 is an input)
int rmax_of(const RowsArgs& a) { return a.has_y1 ? 0 : (a.r <= 4 ? 4 : RW_MAXR); }

template <int NCH, int GATE, int RMAX>
int launch_rows_k(bool bwd, const RowsArgs& a, int sms, cudaStream_t st) {
  const size_t smem = rows_smem_bytes(a.d, RMAX > 0 ? a.r : 0, bwd ? 3 : 2);
  int64_t blocks = (a.M + RW_WARPS - 1) / RW_WARPS;
  if (blocks > 2 * sms) blocks = 2 * sms;            // two resident CTAs per SM (launch bounds), every warp walks rows
  auto kf = rows_fwd_kernel<NCH, GATE, RMAX>;
  auto kb = rows_bwd_kernel<NCH, GATE, RMAX>;
  static int attr_f[64] = {0}, attr_b[64] = {0};
  if (bwd) VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kb), attr_b, (int)smem));
  else VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kf), attr_f, (int)smem));
  if (bwd) kb<<<(unsigned)blocks, RW_THREADS, smem, st>>>(a);
  else kf<<<(unsigned)blocks, RW_THREADS, smem, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}
template <int NCH, int GATE>
int launch_rows_g(bool bwd, const RowsArgs& a, int sms, cudaStream_t st) {
  switch (rmax_of(a)) {
    case 0: return launch_rows_k<NCH, GATE, 0>(bwd, a, sms, st);
    case 4: return launch_rows_k<NCH, GATE, 4>(bwd, a, sms, st);
    default: return launch_rows_k<NCH, GATE, RW_MAXR>(bwd, a, sms, st);
  }
}
template <int NCH>
int launch_rows(bool bwd, const RowsArgs& a, int sms, cudaStream_t st) {
  switch (a.gate) {
    case VLPET_GATE_NONE: return launch_rows_g<NCH, VLPET_GATE_NONE>(bwd, a, sms, st);
    case VLPET_GATE_MIDDLE_X: return launch_rows_g<NCH, VLPET_GATE_MIDDLE_X>(bwd, a, sms, st);
    case VLPET_GATE_MIDDLE_Y: return launch_rows_g<NCH, VLPET_GATE_MIDDLE_Y>(bwd, a, sms, st);
    case VLPET_GATE_SMALL: return launch_rows_g<NCH, VLPET_GATE_SMALL>(bwd, a, sms, st);
  }
  return fail(VLPET_E_UNSUPPORTED, "rows: gate %d", a.gate);
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


// ---- large gate at small ranks: host side -------------------------------------------------------------------------------------
struct LargeWs {
  __nv_bfloat16 *du, *dt, *zs, *das, *qs, *dps;
  float *pWu, *pWd, *pGu, *pGd;
  size_t bytes;
  int r8, rg8, pz, pq;
};
LargeWs carve_large_ws(const VlpetK1Desc& D, bool bwd, void* ws) {
  LargeWs w;
  memset(&w, 0, sizeof(w));
  w.r8 = r8_of(D.r); w.rg8 = r8_of(D.rg); w.pz = w.r8 + 8; w.pq = w.rg8 + 8;
  Arena a(ws, (size_t)-1);
  if (bwd) {
    w.du = a.take<__nv_bfloat16>((size_t)D.M * D.d); w.dt = a.take<__nv_bfloat16>((size_t)D.M * D.d);
    w.zs = a.take<__nv_bfloat16>((size_t)D.M * w.pz); w.das = a.take<__nv_bfloat16>((size_t)D.M * w.pz);
    w.qs = a.take<__nv_bfloat16>((size_t)D.M * w.pq); w.dps = a.take<__nv_bfloat16>((size_t)D.M * w.pq);
    if (w.r8 != D.r) { w.pWu = a.take<float>((size_t)D.d * w.r8); w.pWd = a.take<float>((size_t)D.d * w.r8); }
    if (w.rg8 != D.rg) { w.pGu = a.take<float>((size_t)D.d * w.rg8); w.pGd = a.take<float>((size_t)D.d * w.rg8); }
  }
  w.bytes = a.off;
  return w;
}
bool large_rows_supported(const VlpetK1Desc& D, bool bwd) {
  if (D.dtype != VLPET_BF16 || D.gate != VLPET_GATE_LARGE || D.r > RW_MAXR || D.rg > RW_MAXR || D.r < 1 || D.rg < 1) return false;
  if (D.d % 256 != 0 || D.d < 256 || D.d > 768 || D.M <= 0 || device_sm_count() <= 0) return false;
  if (large_smem_bytes(D.d, D.r, D.rg, 3) > 200 * 1024) return false;
  return !bwd || (wgrad_sm100_supported(D.d, r8_of(D.r)) && wgrad_sm100_supported(D.d, r8_of(D.rg)));
}
LargeArgs large_args(const VlpetK1Desc& D, const VlpetK1Params& w) {
  LargeArgs q;
  memset(&q, 0, sizeof(q));
  q.a = base_args(D, w);
  q.rg = D.rg; q.rg8 = r8_of(D.rg); q.pq = q.rg8 + 8;
  q.a.r8 = r8_of(D.r); q.a.pz = q.a.r8 + 8;
  q.Gd = static_cast<const __nv_bfloat16*>(w.Gd); q.gbd = static_cast<const __nv_bfloat16*>(w.gbd);
  q.Gu = static_cast<const __nv_bfloat16*>(w.Gu); q.gbu = static_cast<const __nv_bfloat16*>(w.gbu);
  return q;
}
template <int NCH, int RMAX>
int launch_large_k(bool bwd, const LargeArgs& q, int sms, cudaStream_t st) {
  const size_t smem = large_smem_bytes(q.a.d, q.a.r, q.rg, bwd ? 3 : 2);
  int64_t blocks = (q.a.M + RW_WARPS - 1) / RW_WARPS;
  if (blocks > 2 * sms) blocks = 2 * sms;
  auto kf = rows_large_fwd_kernel<NCH, RMAX>;
  auto kb = rows_large_bwd_kernel<NCH, RMAX>;
  static int attr_f[64] = {0}, attr_b[64] = {0};
  if (bwd) VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kb), attr_b, (int)smem));
  else VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kf), attr_f, (int)smem));
  if (bwd) kb<<<(unsigned)blocks, RW_THREADS, smem, st>>>(q);
  else kf<<<(unsigned)blocks, RW_THREADS, smem, st>>>(q);
  VLPET_LAUNCH_OK();
  return 0;
}
int launch_large(bool bwd, const LargeArgs& q, int sms, cudaStream_t st) {
  const int rm = (q.a.r > q.rg ? q.a.r : q.rg) <= 4 ? 4 : RW_MAXR;
  switch (q.a.d / 256) {
    case 1: return rm == 4 ? launch_large_k<1, 4>(bwd, q, sms, st) : launch_large_k<1, RW_MAXR>(bwd, q, sms, st);
    case 2: return rm == 4 ? launch_large_k<2, 4>(bwd, q, sms, st) : launch_large_k<2, RW_MAXR>(bwd, q, sms, st);
    case 3: return rm == 4 ? launch_large_k<3, 4>(bwd, q, sms, st) : launch_large_k<3, RW_MAXR>(bwd, q, sms, st);
  }
  return fail(VLPET_E_UNSUPPORTED, "rows(large): d = %d", q.a.d);
}
int large_rows_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, cudaStream_t st) {
  LargeArgs q = large_args(D, w);
  q.a.x1 = static_cast<const __nv_bfloat16*>(x1); q.a.x2 = static_cast<const __nv_bfloat16*>(x2);
  q.a.out = static_cast<__nv_bfloat16*>(out);
  return launch_large(false, q, device_sm_count(), st);
}
int large_rows_bwd(const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w, void* dx1, void* dx2,
                   const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  LargeWs W = carve_large_ws(D, true, ws);
  if (!ws || ws_bytes < W.bytes) return fail(VLPET_E_WORKSPACE, "k1_bwd(rows, large): workspace %zu < %zu bytes", ws_bytes, W.bytes);
  const int sms = device_sm_count();
  LargeArgs q = large_args(D, w);
  q.a.x1 = static_cast<const __nv_bfloat16*>(x1); q.a.x2 = static_cast<const __nv_bfloat16*>(x2);
  q.a.dout = static_cast<const __nv_bfloat16*>(dout);
  q.a.dx1 = static_cast<__nv_bfloat16*>(dx1); q.a.dx2 = static_cast<__nv_bfloat16*>(dx2);
  q.a.du = W.du; q.dt = W.dt; q.a.zs = W.zs; q.a.das = W.das; q.qs = W.qs; q.dps = W.dps;
  q.a.dbd = G.dbd; q.dgbd = G.dgbd;
  VLPET_TRY(launch_large(true, q, sms, st));
  // weight gradients: dWu = du^T z (+dbu), dWd = (x2^T da)^T for the adapter branch; dGu = dT^T q (+dgbu), dGd = (x1^T dp)^T for the gate
  for (int br = 0; br < 2; ++br) {
    const int r = br ? D.rg : D.r, r8 = br ? W.rg8 : W.r8, pitch = br ? W.pq : W.pz;
    float* gWu = br ? G.dGu : G.dWu; float* gWd = br ? G.dGd : G.dWd; float* gbu = br ? G.dgbu : G.dbu;
    float* pWu = br ? W.pGu : W.pWu; float* pWd = br ? W.pGd : W.pWd;
    float *oWu = gWu, *oWd = gWd;
    if (r8 != r) {
      VLPET_CUDA_OK(cudaMemsetAsync(pWu, 0, (size_t)D.d * r8 * sizeof(float), st));
      VLPET_CUDA_OK(cudaMemsetAsync(pWd, 0, (size_t)D.d * r8 * sizeof(float), st));
      oWu = gWu ? pWu : nullptr; oWd = gWd ? pWd : nullptr;
    }
    const void* A[2]; const void* Bm[2]; int64_t lda[2], ldb[2]; int nbv[2], tr[2]; float* out[2]; float* bias[2]; float sc[2];
    int k = 0;
    if (oWu || gbu) {
      if (!oWu) return fail(VLPET_E_BADARG, "k1_bwd(rows, large): a bias gradient needs its weight gradient buffer");
      A[k] = br ? W.dt : W.du; lda[k] = D.d; Bm[k] = br ? W.qs : W.zs; ldb[k] = pitch; nbv[k] = r8 + 1; tr[k] = 0; out[k] = oWu; bias[k] = gbu; sc[k] = 1.f; ++k;
    }
    if (oWd) {
      A[k] = br ? x1 : x2; lda[k] = D.d; Bm[k] = br ? W.dps : W.das; ldb[k] = pitch; nbv[k] = r8; tr[k] = 1; out[k] = oWd; bias[k] = nullptr; sc[k] = 1.f; ++k;
    }
    if (k) VLPET_TRY(wgrad_sm100(k, A, lda, Bm, ldb, nbv, out, bias, sc, tr, D.M, D.d, r8, sms, st));
    if (r8 != r && (gWu || gWd)) {
      const int n = D.d * r;
      unpad_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(pWu, pWd, gWu, gWd, D.d, r, r8);
      VLPET_LAUNCH_OK();
    }
  }
  return 0;
}

}  // namespace

// ---- entry points ------------------------------------------------------------------------------------------------
bool rows_k1_supported(const VlpetK1Desc& D, bool bwd) {
  if (D.gate == VLPET_GATE_LARGE) return large_rows_supported(D, bwd);
  if (D.dtype != VLPET_BF16) return false;
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
size_t rows_k1_fwd_ws(const VlpetK1Desc& D) {
  return D.gate == VLPET_GATE_LARGE ? carve_large_ws(D, false, nullptr).bytes : carve_rows(D, false, nullptr).bytes;
}
size_t rows_k1_bwd_ws(const VlpetK1Desc& D) {
  return D.gate == VLPET_GATE_LARGE ? carve_large_ws(D, true, nullptr).bytes : carve_rows(D, true, nullptr).bytes;
}

int rows_k1_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  if (D.gate == VLPET_GATE_LARGE) return large_rows_fwd(D, x1, x2, w, out, st);
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
  if (D.gate == VLPET_GATE_LARGE) return large_rows_bwd(D, x1, x2, dout, w, dx1, dx2, G, ws, ws_bytes, st);
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
    RowsArgs b0 = a;
    b0.pass = 0;
    VLPET_TRY(launch_rows_d(true, b0, sms, st));       // backward pass 0: the per-sample gate mean and dG in one sweep
    a.pass = 1;
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
