// Short-sequence attention of the frozen BART blocks (my_transformers/modeling_bart.py:143-280: bmm -> softmax -> dropout -> bmm),
// forward and backward, ONE CTA per (batch, head).  SURVEY §8 f-3: at the sequence lengths of the VL-PET workloads (36-72
// visual + 4-20 text tokens, 5-40 target tokens) the library flash kernels are overhead-bound (64..128-wide tiles, one
// CTA per 64 queries, online softmax): 59 us forward / 119 us backward per call, 24 % of a training step.  Here the
// whole [L x L] score tile of a head lives in shared memory: S = Q K^T, softmax, dropout, O = P V (and the five products
// of the backward) are single-tile warp-level tensor-core products (wmma m16n16k16 bf16 -> fp32; the tiles are far too
// small for a tcgen05 pipeline), row softmax by one warp per row.  head_dim = 64, L <= 128, bf16, no padding mask
// (optionally causal); anything else stays on torch SDPA.  Dropout is the counter-based stream of K1 (drop_scale): the
// backward regenerates the mask, nothing but the row log-sum-exp is saved.
// STATUS: parity-green (tests/test_gpu_parity.py::test_short_attention_*), but measured slower than torch's memory-efficient
// SDPA on B200 (146 vs 59 us forward, 295 vs 119 us backward per call at B = 300..500, H = 12, L = 56): the shared-memory
// round trips of S / P and the one-warp-per-row softmax dominate.  The host model keeps torch SDPA (host/vlbart.py
// SHORT_ATTENTION = False); a register-resident (FA2-style fragment softmax) version is the follow-up.
#include <mma.h>

#include "vlpet_common.cuh"

namespace vlpet {
namespace {
using namespace nvcuda;

constexpr int HD = 64;        // head dimension
constexpr int QP = HD + 8;    // row pitch (elements) of the Q/K/V/dO tiles in shared memory

template <int LP> struct AttnCfg {
  static constexpr int SP = LP + 4;                 // fp32 score pitch
  static constexpr int PP = LP + 8;                 // bf16 probability pitch
  static constexpr int THREADS = LP == 64 ? 128 : 256;
  static constexpr int TILE = LP * QP * 2;          // bytes of one [LP x 64] bf16 tile
  static constexpr int S_BYTES = LP * SP * 4;
  static constexpr int P_BYTES = LP * PP * 2;
  static constexpr int FWD_SMEM = 3 * TILE + S_BYTES + P_BYTES;
  static constexpr int BWD_SMEM = 4 * TILE + S_BYTES + 2 * P_BYTES + LP * 4;
};

struct AttnArgs {
  const __nv_bfloat16 *q, *k, *v;
  int64_t q_rs, k_rs, v_rs;          // row strides (elements) of q / k / v; head h starts h*64 elements into a row
  __nv_bfloat16* out;                // [B, Lq, H*64]
  float* lse;                        // [B, H, Lq]
  const __nv_bfloat16 *o, *dout;     // backward: forward output and its gradient, [B, Lq, H*64]
  __nv_bfloat16 *dq, *dk, *dv;       // [B, L, H*64]
  int B, H, Lq, Lk, causal;
  float scale;
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t thr16;
  float inv_keep;
};

// [rows x 64] bf16 tile of head h -> shared memory (zero rows beyond `rows`)
template <int LP, int THREADS>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t rs, int rows) {
  for (int i = threadIdx.x; i < LP * 8; i += THREADS) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (r < rows) q = *reinterpret_cast<const uint4*>(src + (int64_t)r * rs + c);
    *reinterpret_cast<uint4*>(dst + r * QP + c) = q;
  }
}
// fp32 [rows x 64] tile in shared memory (pitch `pitch`) -> bf16 global rows of head h, scaled
template <int THREADS>
__device__ __forceinline__ void store_tile(__nv_bfloat16* dst, int64_t rs, const float* src, int pitch, int rows, float sc) {
  for (int i = threadIdx.x; i < rows * 8; i += THREADS) {
    const int r = i >> 3, c = (i & 7) * 8;
    const float* s = src + r * pitch + c;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(sc * s[0], sc * s[1]), t1 = __floats2bfloat162_rn(sc * s[2], sc * s[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(sc * s[4], sc * s[5]), t3 = __floats2bfloat162_rn(sc * s[6], sc * s[7]);
    uint4 q;
    q.x = *reinterpret_cast<uint32_t*>(&t0); q.y = *reinterpret_cast<uint32_t*>(&t1);
    q.z = *reinterpret_cast<uint32_t*>(&t2); q.w = *reinterpret_cast<uint32_t*>(&t3);
    *reinterpret_cast<uint4*>(dst + (int64_t)r * rs + c) = q;
  }
}

typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> FragA;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> FragAT;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> FragB;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> FragBT;
typedef wmma::fragment<wmma::accumulator, 16, 16, 16, float> FragC;

// C[M x N] (fp32, pitch ldc) = A[M x K] * B[K x N]; A row-major (TA = false) or stored transposed; B likewise.
// Tiles are distributed round-robin over the warps of the CTA.
template <bool TA, bool TB, int NWARPS>
__device__ __forceinline__ void tile_gemm(float* C, int ldc, const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int ldb,
                                          int M, int N, int K) {
  const int warp = threadIdx.x >> 5;
  const int tm = M / 16, tn = N / 16;
  for (int t = warp; t < tm * tn; t += NWARPS) {
    const int mi = t / tn, ni = t % tn;
    FragC acc;
    wmma::fill_fragment(acc, 0.f);
    for (int kk = 0; kk < K / 16; ++kk) {
      if (TA) {
        FragAT a;
        wmma::load_matrix_sync(a, A + kk * 16 * lda + mi * 16, lda);       // A^T stored: element (m, k) at [k][m]
        if (TB) { FragBT b; wmma::load_matrix_sync(b, Bm + ni * 16 * ldb + kk * 16, ldb); wmma::mma_sync(acc, a, b, acc); }
        else { FragB b; wmma::load_matrix_sync(b, Bm + kk * 16 * ldb + ni * 16, ldb); wmma::mma_sync(acc, a, b, acc); }
      } else {
        FragA a;
        wmma::load_matrix_sync(a, A + mi * 16 * lda + kk * 16, lda);
        if (TB) { FragBT b; wmma::load_matrix_sync(b, Bm + ni * 16 * ldb + kk * 16, ldb); wmma::mma_sync(acc, a, b, acc); }   // B^T stored: (k, n) at [n][k]
        else { FragB b; wmma::load_matrix_sync(b, Bm + kk * 16 * ldb + ni * 16, ldb); wmma::mma_sync(acc, a, b, acc); }
      }
    }
    wmma::store_matrix_sync(C + mi * 16 * ldc + ni * 16, acc, ldc, wmma::mem_row_major);
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int LP>
__global__ void __launch_bounds__(AttnCfg<LP>::THREADS) attn_fwd_kernel(const AttnArgs a) {
  using C = AttnCfg<LP>;
  constexpr int NW = C::THREADS / 32;
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + LP * QP;
  __nv_bfloat16* sV = sK + LP * QP;
  float* sS = reinterpret_cast<float*>(smem + 3 * C::TILE);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(smem + 3 * C::TILE + C::S_BYTES);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? *a.seed_dev : 0ull);
  load_tile<LP, C::THREADS>(sQ, a.q + (int64_t)b * a.Lq * a.q_rs + h * HD, a.q_rs, a.Lq);
  load_tile<LP, C::THREADS>(sK, a.k + (int64_t)b * a.Lk * a.k_rs + h * HD, a.k_rs, a.Lk);
  load_tile<LP, C::THREADS>(sV, a.v + (int64_t)b * a.Lk * a.v_rs + h * HD, a.v_rs, a.Lk);
  __syncthreads();
  tile_gemm<false, true, NW>(sS, C::SP, sQ, QP, sK, QP, LP, LP, HD);                 // S = Q K^T
  __syncthreads();
  const float sl2 = a.scale * 1.4426950408889634f;
  for (int i = warp; i < LP; i += NW) {                                               // one warp per query row
    float s[LP / 32];
    float m = -3.0e38f;
#pragma unroll
    for (int t = 0; t < LP / 32; ++t) {
      const int j = lane + 32 * t;
      const bool ok = i < a.Lq && j < a.Lk && (!a.causal || j <= i);
      s[t] = ok ? sS[i * C::SP + j] * sl2 : -3.0e38f;
      m = fmaxf(m, s[t]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < LP / 32; ++t) { s[t] = (s[t] > -1.0e38f) ? exp2f(s[t] - m) : 0.f; sum += s[t]; }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    const int64_t base = (((int64_t)b * a.H + h) * a.Lq + i) * a.Lk;
#pragma unroll
    for (int t = 0; t < LP / 32; ++t) {
      const int j = lane + 32 * t;
      float pv = s[t] * inv;
      if (a.thr16 && pv != 0.f) pv *= drop_scale(seed, a.thr16, a.inv_keep, base + j);
      sP[i * C::PP + j] = __float2bfloat16_rn(pv);
    }
    if (lane == 0 && i < a.Lq) a.lse[((int64_t)b * a.H + h) * a.Lq + i] = (m + log2f(sum)) * 0.6931471805599453f;   // natural log of sum exp(scale*s)
  }
  __syncthreads();
  tile_gemm<false, false, NW>(sS, C::SP, sP, C::PP, sV, QP, LP, HD, LP);              // O = P V  (fp32 into the score buffer)
  __syncthreads();
  store_tile<C::THREADS>(a.out + (int64_t)b * a.Lq * (a.H * HD) + h * HD, (int64_t)a.H * HD, sS, C::SP, a.Lq, 1.0f);
}

template <int LP>
__global__ void __launch_bounds__(AttnCfg<LP>::THREADS) attn_bwd_kernel(const AttnArgs a) {
  using C = AttnCfg<LP>;
  constexpr int NW = C::THREADS / 32;
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + LP * QP;
  __nv_bfloat16* sV = sK + LP * QP;
  __nv_bfloat16* sdO = sV + LP * QP;
  float* sS = reinterpret_cast<float*>(smem + 4 * C::TILE);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(smem + 4 * C::TILE + C::S_BYTES);          // P (not dropped)
  __nv_bfloat16* sX = sP + LP * C::PP;                                                            // P dropped, later dS
  float* sD = reinterpret_cast<float*>(smem + 4 * C::TILE + C::S_BYTES + 2 * C::P_BYTES);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? *a.seed_dev : 0ull);
  const int64_t orow = (int64_t)a.H * HD;
  load_tile<LP, C::THREADS>(sQ, a.q + (int64_t)b * a.Lq * a.q_rs + h * HD, a.q_rs, a.Lq);
  load_tile<LP, C::THREADS>(sK, a.k + (int64_t)b * a.Lk * a.k_rs + h * HD, a.k_rs, a.Lk);
  load_tile<LP, C::THREADS>(sV, a.v + (int64_t)b * a.Lk * a.v_rs + h * HD, a.v_rs, a.Lk);
  load_tile<LP, C::THREADS>(sdO, a.dout + (int64_t)b * a.Lq * orow + h * HD, orow, a.Lq);
  for (int i = warp; i < LP; i += NW) {                                               // D_i = sum_d dO[i,d] O[i,d]
    float d = 0.f;
    if (i < a.Lq) {
      const int64_t off = ((int64_t)b * a.Lq + i) * orow + h * HD + lane * 2;
      const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(a.o + off);
      const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(a.dout + off);
      d = __bfloat162float(o2.x) * __bfloat162float(g2.x) + __bfloat162float(o2.y) * __bfloat162float(g2.y);
    }
    d = warp_sum(d);
    if (lane == 0) sD[i] = d;
  }
  __syncthreads();
  tile_gemm<false, true, NW>(sS, C::SP, sQ, QP, sK, QP, LP, LP, HD);                 // S = Q K^T
  __syncthreads();
  const float sl2 = a.scale * 1.4426950408889634f;
  for (int i = warp; i < LP; i += NW) {                                               // P = exp(scale*S - lse), P dropped
    const float l2 = (i < a.Lq) ? a.lse[((int64_t)b * a.H + h) * a.Lq + i] * 1.4426950408889634f : 0.f;
    const int64_t base = (((int64_t)b * a.H + h) * a.Lq + i) * a.Lk;
#pragma unroll
    for (int t = 0; t < LP / 32; ++t) {
      const int j = lane + 32 * t;
      const bool ok = i < a.Lq && j < a.Lk && (!a.causal || j <= i);
      const float pv = ok ? exp2f(sS[i * C::SP + j] * sl2 - l2) : 0.f;
      float pd = pv;
      if (a.thr16 && pv != 0.f) pd *= drop_scale(seed, a.thr16, a.inv_keep, base + j);
      sP[i * C::PP + j] = __float2bfloat16_rn(pv);
      sX[i * C::PP + j] = __float2bfloat16_rn(pd);
    }
  }
  __syncthreads();
  tile_gemm<true, false, NW>(sS, C::SP, sX, C::PP, sdO, QP, LP, HD, LP);              // dV = Pd^T dO
  __syncthreads();
  store_tile<C::THREADS>(a.dv + (int64_t)b * a.Lk * orow + h * HD, orow, sS, C::SP, a.Lk, 1.0f);
  __syncthreads();
  tile_gemm<false, true, NW>(sS, C::SP, sdO, QP, sV, QP, LP, LP, HD);                 // dP = dO V^T
  __syncthreads();
  for (int i = warp; i < LP; i += NW) {                                               // dS = P (dP*mask - D) (scale applied at the stores)
    const float di = sD[i];
    const int64_t base = (((int64_t)b * a.H + h) * a.Lq + i) * a.Lk;
#pragma unroll
    for (int t = 0; t < LP / 32; ++t) {
      const int j = lane + 32 * t;
      const float pv = __bfloat162float(sP[i * C::PP + j]);
      float dp = sS[i * C::SP + j];
      if (a.thr16 && pv != 0.f) dp *= drop_scale(seed, a.thr16, a.inv_keep, base + j);
      sX[i * C::PP + j] = __float2bfloat16_rn(pv * (dp - di));
    }
  }
  __syncthreads();
  tile_gemm<false, false, NW>(sS, C::SP, sX, C::PP, sK, QP, LP, HD, LP);              // dQ = dS K
  __syncthreads();
  store_tile<C::THREADS>(a.dq + (int64_t)b * a.Lq * orow + h * HD, orow, sS, C::SP, a.Lq, a.scale);
  __syncthreads();
  tile_gemm<true, false, NW>(sS, C::SP, sX, C::PP, sQ, QP, LP, HD, LP);               // dK = dS^T Q
  __syncthreads();
  store_tile<C::THREADS>(a.dk + (int64_t)b * a.Lk * orow + h * HD, orow, sS, C::SP, a.Lk, a.scale);
}

template <int LP>
int launch_attn(bool bwd, const AttnArgs& a, cudaStream_t st) {
  using C = AttnCfg<LP>;
  static bool set_f = false, set_b = false;
  if (!bwd && !set_f) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::FWD_SMEM));
    set_f = true;
  }
  if (bwd && !set_b) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::BWD_SMEM));
    set_b = true;
  }
  const unsigned grid = (unsigned)(a.B * a.H);
  if (bwd) attn_bwd_kernel<LP><<<grid, C::THREADS, C::BWD_SMEM, st>>>(a);
  else attn_fwd_kernel<LP><<<grid, C::THREADS, C::FWD_SMEM, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}

int check_attn(const AttnArgs& a) {
  if (!a.q || !a.k || !a.v || a.B <= 0 || a.H <= 0 || a.Lq <= 0 || a.Lk <= 0) return fail(VLPET_E_BADARG, "attn: bad arguments");
  if (a.Lq > 128 || a.Lk > 128) return fail(VLPET_E_UNSUPPORTED, "attn: sequence lengths up to 128 (Lq=%d Lk=%d)", a.Lq, a.Lk);
  if ((a.q_rs | a.k_rs | a.v_rs) % 8 != 0 || !aligned16(a.q) || !aligned16(a.k) || !aligned16(a.v))
    return fail(VLPET_E_ALIGN, "attn: q/k/v rows must be 16-byte aligned");
  if ((int64_t)a.B * a.H > 0x7fffffff) return fail(VLPET_E_UNSUPPORTED, "attn: too many heads");
  return 0;
}

}  // namespace

int attn_run(bool bwd, const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, void* out, float* lse,
             const void* o, const void* dout, void* dq, void* dk, void* dv, int B, int H, int Lq, int Lk, int causal, float p_drop,
             uint64_t seed, const uint64_t* seed_dev, cudaStream_t st) {
  AttnArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k); a.v = static_cast<const __nv_bfloat16*>(v);
  a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs;
  a.out = static_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.o = static_cast<const __nv_bfloat16*>(o); a.dout = static_cast<const __nv_bfloat16*>(dout);
  a.dq = static_cast<__nv_bfloat16*>(dq); a.dk = static_cast<__nv_bfloat16*>(dk); a.dv = static_cast<__nv_bfloat16*>(dv);
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal;
  a.scale = 0.125f;   // 1 / sqrt(64)
  a.seed = seed; a.seed_dev = seed_dev;
  a.thr16 = p_drop > 0.f ? drop_thr16(p_drop) : 0u;
  a.inv_keep = a.thr16 ? 1.0f / (1.0f - (float)a.thr16 / 65536.0f) : 1.0f;
  VLPET_TRY(check_attn(a));
  if (!lse) return fail(VLPET_E_BADARG, "attn: lse missing");
  if (bwd && (!o || !dout || !dq || !dk || !dv || !aligned16(o) || !aligned16(dout) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv)))
    return fail(VLPET_E_BADARG, "attn_bwd: bad arguments");
  if (!bwd && (!out || !aligned16(out))) return fail(VLPET_E_BADARG, "attn_fwd: bad output");
  const int L = Lq > Lk ? Lq : Lk;
  return L <= 64 ? launch_attn<64>(bwd, a, st) : launch_attn<128>(bwd, a, st);
}

}  // namespace vlpet
