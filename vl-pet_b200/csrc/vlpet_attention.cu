// Short-sequence attention of the frozen BART blocks (my_transformers/modeling_bart.py:143-280: bmm -> softmax -> dropout -> bmm),
// forward and backward, ONE CTA per (batch, head).  SURVEY §8 f-3: at the sequence lengths of the VL-PET workloads (36-72
// visual + 4-20 text tokens, 5-40 target tokens) the library flash kernels are overhead-bound (64..128-wide tiles, one
// CTA per 64 queries, online softmax): 59 us forward / 119 us backward per call, 24 % of a training step.  Here the
// whole [L x L] score tile of a head lives in shared memory: S = Q K^T, softmax, dropout, O = P V (and the five products
// of the backward) are single-tile warp-level tensor-core products (wmma m16n16k16 bf16 -> fp32; the tiles are far too
// small for a tcgen05 pipeline), row softmax by one warp per row.  head_dim = 64, L <= 128, bf16, no padding mask
// (optionally causal); anything else stays on torch SDPA.  Dropout is the counter-based stream of K1 (drop_scale): the
// backward regenerates the mask, nothing but the row log-sum-exp is saved.
// STATUS (tools/attn_probe.py, B200): L <= 64 runs the register-resident kernels further down (attn2_*: mma.sync fragments,
// softmax on the fragments, no shared-memory round trip of S / P): 56 vs 95 us forward and 187 vs 314 us forward + backward
// per encoder call (B = 300, H = 12, L = 56) against torch SDPA, 139 vs 328 us for the causal decoder call.  65 <= L <= 128
// runs the shared-memory (wmma) kernels below, which are parity-green but SLOWER than SDPA (583 vs 194 us at L = 92): the host
// model only routes L <= 64 here.
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
  __nv_bfloat16 *dq, *dk, *dv;       // rows of head-interleaved gradients, row strides below
  int64_t dq_rs, dk_rs, dv_rs;       // elements; H*64 for separate [B, L, H*64] tensors, 3*H*64 for the thirds of a fused buffer
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
  store_tile<C::THREADS>(a.dv + (int64_t)b * a.Lk * a.dv_rs + h * HD, a.dv_rs, sS, C::SP, a.Lk, 1.0f);
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
  store_tile<C::THREADS>(a.dq + (int64_t)b * a.Lq * a.dq_rs + h * HD, a.dq_rs, sS, C::SP, a.Lq, a.scale);
  __syncthreads();
  tile_gemm<true, false, NW>(sS, C::SP, sX, C::PP, sQ, QP, LP, HD, LP);               // dK = dS^T Q
  __syncthreads();
  store_tile<C::THREADS>(a.dk + (int64_t)b * a.Lk * a.dk_rs + h * HD, a.dk_rs, sS, C::SP, a.Lk, a.scale);
}

// =====================================================================================================================
// L <= 64: register-resident version.  One CTA (4 warps) per (batch, head); warp w owns query rows 16w..16w+15.  The
// score / probability tile of a warp never leaves its registers: S = Q K^T and dP = dO V^T are mma.sync m16n8k16
// accumulator fragments, the row softmax runs on the fragments (quad shuffles), and P (dS) is re-packed in registers into
// the A operand of the next product (accumulator pair -> bf16x2, the FlashAttention-2 trick).  Only the two products of
// the backward that contract over queries (dV = Pd^T dO, dK = dS^T Q) go through shared memory once.
// Dropout: one 32-bit hash per pair of adjacent key columns of a row (its own counter-based stream, regenerated by the
// backward).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {     // lowbias32
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// per-row key of the dropout stream, and 2 x 16 random bits for key columns (c, c+1), c even, of that row
__device__ __forceinline__ uint32_t attn_rowkey(uint32_t seed_lo, uint32_t seed_hi, uint32_t rowid) {
  return seed_lo ^ mix32(rowid * 0x9E3779B1u + seed_hi);
}
__device__ __forceinline__ uint32_t attn_bits(uint32_t rowkey, uint32_t c) { return mix32(rowkey ^ ((c >> 1) * 0x85EBCA6Bu)); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
constexpr int L2P = 64;                       // padded sequence length of the register-resident kernels
constexpr int T2 = L2P * QP * 2;              // bytes of one [64 x 64] bf16 tile (pitch QP)

// A fragments (m16 x k16, 4 k-steps over 64 columns) of rows r0..r0+15 of a row-major [64][QP] tile
__device__ __forceinline__ void load_a_rows(uint32_t tile, int r0, int lane, uint32_t (&a)[4][4]) {
  const int id = lane >> 3, r = lane & 7;
  const uint32_t base = tile + (uint32_t)((r0 + (id & 1) * 8 + r) * QP + (id >> 1) * 8) * 2u;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(base + ks * 32, a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}
// acc[16 x 64] += A[16 x 64] * T^T where T is a row-major [64 n][64 k] tile (k contiguous): scores Q K^T, dP = dO V^T
__device__ __forceinline__ void gemm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
  const int id = lane >> 3, r = lane & 7;
  const uint32_t base = tile + (uint32_t)(((id >> 1) * 8 + r) * QP + (id & 1) * 8) * 2u;
#pragma unroll
  for (int np = 0; np < 4; ++np) {            // two n-tiles (16 rows of T) per ldmatrix.x4
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(base + (uint32_t)(np * 16 * QP) * 2u + ks * 32, b0, b1, b2, b3);
      mma16816(acc[2 * np], a[ks], b0, b1);
      mma16816(acc[2 * np + 1], a[ks], b2, b3);
    }
  }
}
// acc[16 x 64] += P[16 x 64 k] * T where T is a row-major [64 k][64 n] tile (n contiguous): O = P V, dQ = dS K
__device__ __forceinline__ void gemm_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
  const int id = lane >> 3, r = lane & 7;
  const uint32_t base = tile + (uint32_t)(((id & 1) * 8 + r) * QP + (id >> 1) * 8) * 2u;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(base + (uint32_t)(ks * 16 * QP) * 2u + np * 32, b0, b1, b2, b3);
      mma16816(acc[2 * np], a[ks], b0, b1);
      mma16816(acc[2 * np + 1], a[ks], b2, b3);
    }
  }
}
// accumulator fragments [16 x 64] -> A fragments (bf16) of the same [16 x 64] matrix
__device__ __forceinline__ void acc_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// fragment rows of this thread: row0 = 16w + g, row1 = row0 + 8; columns of element (n, e): 8n + 2t + (e & 1), e < 2 -> row0
// writes the thread's two rows of a [16 x 64] accumulator tile as bf16 (4-byte stores)
__device__ __forceinline__ void store_frag_rows(__nv_bfloat16* dst, int64_t rs, const float (&c)[8][4], int row0, int nrows, int t, float sc) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (row0 < nrows) *reinterpret_cast<uint32_t*>(dst + (int64_t)row0 * rs + n * 8 + 2 * t) = pack_bf2(sc * c[n][0], sc * c[n][1]);
    if (row0 + 8 < nrows) *reinterpret_cast<uint32_t*>(dst + (int64_t)(row0 + 8) * rs + n * 8 + 2 * t) = pack_bf2(sc * c[n][2], sc * c[n][3]);
  }
}

__global__ void __launch_bounds__(128) attn2_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + L2P * QP;
  __nv_bfloat16* sV = sK + L2P * QP;
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(sQ), uK = uQ + T2, uV = uK + T2;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? *a.seed_dev : 0ull);
  load_tile<L2P, 128>(sQ, a.q + (int64_t)b * a.Lq * a.q_rs + h * HD, a.q_rs, a.Lq);
  load_tile<L2P, 128>(sK, a.k + (int64_t)b * a.Lk * a.k_rs + h * HD, a.k_rs, a.Lk);
  load_tile<L2P, 128>(sV, a.v + (int64_t)b * a.Lk * a.v_rs + h * HD, a.v_rs, a.Lk);
  __syncthreads();
  if (warp * 16 >= a.Lq) return;                       // no query rows for this warp
  uint32_t qa[4][4];
  load_a_rows(uQ, warp * 16, lane, qa);
  float s[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
  gemm_nt(s, qa, uK, lane);                            // S = Q K^T
  const float sl2 = a.scale * 1.4426950408889634f;
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = n * 8 + 2 * t + (e & 1), row = (e < 2) ? row0 : row1;
      const bool ok = col < a.Lk && (!a.causal || col <= row);
      s[n][e] = ok ? s[n][e] * sl2 : -3.0e38f;
      if (e < 2) m0 = fmaxf(m0, s[n][e]); else m1 = fmaxf(m1, s[n][e]);
    }
  m0 = quad_max(m0); m1 = quad_max(m1);
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float mm = (e < 2) ? m0 : m1;
      const float pv = (s[n][e] > -1.0e38f) ? ex2_approx(s[n][e] - mm) : 0.f;
      s[n][e] = pv;
      if (e < 2) sum0 += pv; else sum1 += pv;
    }
  sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
  const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
  const int64_t bh = (int64_t)b * a.H + h;
  if (t == 0) {
    if (row0 < a.Lq) a.lse[bh * a.Lq + row0] = (m0 + log2f(sum0)) * 0.6931471805599453f;
    if (row1 < a.Lq) a.lse[bh * a.Lq + row1] = (m1 + log2f(sum1)) * 0.6931471805599453f;
  }
  const float k0 = inv0 * a.inv_keep, k1 = inv1 * a.inv_keep;   // inv_keep == 1 without dropout
  const uint32_t rk0 = attn_rowkey((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * a.Lq + row0));
  const uint32_t rk1 = attn_rowkey((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * a.Lq + row1));
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    float f0 = k0, f1 = k0, f2 = k1, f3 = k1;
    if (a.thr16) {
      const uint32_t c = (uint32_t)(n * 8 + 2 * t);
      const uint32_t r0b = attn_bits(rk0, c), r1b = attn_bits(rk1, c);
      if ((r0b & 0xffffu) < a.thr16) f0 = 0.f;
      if ((r0b >> 16) < a.thr16) f1 = 0.f;
      if ((r1b & 0xffffu) < a.thr16) f2 = 0.f;
      if ((r1b >> 16) < a.thr16) f3 = 0.f;
    }
    s[n][0] *= f0; s[n][1] *= f1; s[n][2] *= f2; s[n][3] *= f3;
  }
  uint32_t pa[4][4];
  acc_to_a(s, pa);
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  gemm_nn(o, pa, uV, lane);                            // O = P V
  store_frag_rows(a.out + (int64_t)b * a.Lq * (a.H * HD) + h * HD, (int64_t)a.H * HD, o, row0, a.Lq, t, 1.0f);
}

__global__ void __launch_bounds__(128) attn2_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + L2P * QP;
  __nv_bfloat16* sV = sK + L2P * QP;
  __nv_bfloat16* sdO = sV + L2P * QP;
  __nv_bfloat16* sPd = sdO + L2P * QP;                 // P dropped  [q][key]
  __nv_bfloat16* sdS = sPd + L2P * QP;                 // dS         [q][key]
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(sQ), uK = uQ + T2, uV = uK + T2, udO = uV + T2, uPd = udO + T2, udS = uPd + T2;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint64_t seed = a.seed + ((a.thr16 && a.seed_dev) ? *a.seed_dev : 0ull);
  const int64_t orow = (int64_t)a.H * HD;
  load_tile<L2P, 128>(sQ, a.q + (int64_t)b * a.Lq * a.q_rs + h * HD, a.q_rs, a.Lq);
  load_tile<L2P, 128>(sK, a.k + (int64_t)b * a.Lk * a.k_rs + h * HD, a.k_rs, a.Lk);
  load_tile<L2P, 128>(sV, a.v + (int64_t)b * a.Lk * a.v_rs + h * HD, a.v_rs, a.Lk);
  load_tile<L2P, 128>(sdO, a.dout + (int64_t)b * a.Lq * orow + h * HD, orow, a.Lq);
  __syncthreads();
  const int row0 = warp * 16 + g, row1 = row0 + 8;
  const int64_t bh = (int64_t)b * a.H + h;
  const float sl2 = a.scale * 1.4426950408889634f;
  {
    uint32_t qa[4][4], da[4][4];
    load_a_rows(uQ, warp * 16, lane, qa);
    load_a_rows(udO, warp * 16, lane, da);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
    gemm_nt(s, qa, uK, lane);                          // S = Q K^T
    gemm_nt(dp, da, uV, lane);                         // dP = dO V^T
    const float l0 = (row0 < a.Lq) ? a.lse[bh * a.Lq + row0] * 1.4426950408889634f : 0.f;
    const float l1 = (row1 < a.Lq) ? a.lse[bh * a.Lq + row1] * 1.4426950408889634f : 0.f;
    const uint32_t rk0 = attn_rowkey((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * a.Lq + row0));
    const uint32_t rk1 = attn_rowkey((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * a.Lq + row1));
    float d0 = 0.f, d1 = 0.f;
    float pd[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float f[4] = {a.inv_keep, a.inv_keep, a.inv_keep, a.inv_keep};
      if (a.thr16) {
        const uint32_t c = (uint32_t)(n * 8 + 2 * t);
        const uint32_t r0b = attn_bits(rk0, c), r1b = attn_bits(rk1, c);
        if ((r0b & 0xffffu) < a.thr16) f[0] = 0.f;
        if ((r0b >> 16) < a.thr16) f[1] = 0.f;
        if ((r1b & 0xffffu) < a.thr16) f[2] = 0.f;
        if ((r1b >> 16) < a.thr16) f[3] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = n * 8 + 2 * t + (e & 1), row = (e < 2) ? row0 : row1;
        const bool ok = row < a.Lq && col < a.Lk && (!a.causal || col <= row);
        const float pv = ok ? ex2_approx(s[n][e] * sl2 - ((e < 2) ? l0 : l1)) : 0.f;
        s[n][e] = pv;                                  // P
        pd[n][e] = pv * f[e];                          // P dropped (scaled)
        dp[n][e] *= f[e];                              // dP through the dropout
        if (e < 2) d0 += pv * dp[n][e]; else d1 += pv * dp[n][e];
      }
    }
    d0 = quad_sum(d0); d1 = quad_sum(d1);              // D_i = sum_j P_ij dP_ij  (= rowsum(dO . O))
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] *= dp[n][0] - d0; s[n][1] *= dp[n][1] - d0;          // dS (unscaled)
      s[n][2] *= dp[n][2] - d1; s[n][3] *= dp[n][3] - d1;
      const int c = n * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(sPd + row0 * QP + c) = pack_bf2(pd[n][0], pd[n][1]);
      *reinterpret_cast<uint32_t*>(sPd + row1 * QP + c) = pack_bf2(pd[n][2], pd[n][3]);
      *reinterpret_cast<uint32_t*>(sdS + row0 * QP + c) = pack_bf2(s[n][0], s[n][1]);
      *reinterpret_cast<uint32_t*>(sdS + row1 * QP + c) = pack_bf2(s[n][2], s[n][3]);
    }
    uint32_t sa[4][4];
    acc_to_a(s, sa);
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
    gemm_nn(dq, sa, uK, lane);                         // dQ = dS K
    store_frag_rows(a.dq + (int64_t)b * a.Lq * a.dq_rs + h * HD, a.dq_rs, dq, row0, a.Lq, t, a.scale);
  }
  __syncthreads();
  // dV = Pd^T dO and dK = dS^T Q: warp w owns key rows 16w..16w+15, contraction over all 64 queries
  {
    const int id = lane >> 3, r = lane & 7;
    // A^T fragments from a [q][key] tile: matrix id -> q rows (id >> 1) * 8, key columns 16w + (id & 1) * 8
    const uint32_t offA = (uint32_t)(((id >> 1) * 8 + r) * QP + warp * 16 + (id & 1) * 8) * 2u;
    uint32_t pa[4][4], sa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4_t(uPd + offA + (uint32_t)(ks * 16 * QP) * 2u, pa[ks][0], pa[ks][1], pa[ks][2], pa[ks][3]);
      ldsm_x4_t(udS + offA + (uint32_t)(ks * 16 * QP) * 2u, sa[ks][0], sa[ks][1], sa[ks][2], sa[ks][3]);
    }
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; }
    gemm_nn(dv, pa, udO, lane);                        // dV[key][dim] = sum_q Pd[q][key] dO[q][dim]
    gemm_nn(dk, sa, uQ, lane);                         // dK[key][dim] = sum_q dS[q][key] Q[q][dim]
    store_frag_rows(a.dv + (int64_t)b * a.Lk * a.dv_rs + h * HD, a.dv_rs, dv, warp * 16 + g, a.Lk, t, 1.0f);
    store_frag_rows(a.dk + (int64_t)b * a.Lk * a.dk_rs + h * HD, a.dk_rs, dk, warp * 16 + g, a.Lk, t, a.scale);
  }
}

int launch_attn2(bool bwd, const AttnArgs& a, cudaStream_t st) {
  static int set_b[64] = {0};      // per device ordinal: the attribute applies to the current device's copy of the kernel
  constexpr int FWD_SMEM = 3 * T2, BWD_SMEM = 6 * T2;
  if (bwd) VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(attn2_bwd_kernel), set_b, BWD_SMEM));
  const unsigned grid = (unsigned)(a.B * a.H);
  if (bwd) attn2_bwd_kernel<<<grid, 128, BWD_SMEM, st>>>(a);
  else attn2_fwd_kernel<<<grid, 128, FWD_SMEM, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}

template <int LP>
int launch_attn(bool bwd, const AttnArgs& a, cudaStream_t st) {
  using C = AttnCfg<LP>;
  static int set_f[64] = {0}, set_b[64] = {0};   // per device ordinal (one process may drive several GPUs)
  if (!bwd) VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd_kernel<LP>), set_f, C::FWD_SMEM));
  else VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_kernel<LP>), set_b, C::BWD_SMEM));
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
             const void* o, const void* dout, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs, int64_t dv_rs, int B, int H,
             int Lq, int Lk, int causal, float p_drop, uint64_t seed, const uint64_t* seed_dev, cudaStream_t st) {
  AttnArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k); a.v = static_cast<const __nv_bfloat16*>(v);
  a.q_rs = q_rs; a.k_rs = k_rs; a.v_rs = v_rs;
  a.out = static_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.o = static_cast<const __nv_bfloat16*>(o); a.dout = static_cast<const __nv_bfloat16*>(dout);
  a.dq = static_cast<__nv_bfloat16*>(dq); a.dk = static_cast<__nv_bfloat16*>(dk); a.dv = static_cast<__nv_bfloat16*>(dv);
  a.dq_rs = dq_rs; a.dk_rs = dk_rs; a.dv_rs = dv_rs;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.causal = causal;
  a.scale = 0.125f;   // 1 / sqrt(64)
  a.seed = seed; a.seed_dev = seed_dev;
  a.thr16 = p_drop > 0.f ? drop_thr16(p_drop) : 0u;
  a.inv_keep = a.thr16 ? 1.0f / (1.0f - (float)a.thr16 / 65536.0f) : 1.0f;
  VLPET_TRY(check_attn(a));
  if (!lse) return fail(VLPET_E_BADARG, "attn: lse missing");
  if (bwd && (!o || !dout || !dq || !dk || !dv || !aligned16(o) || !aligned16(dout) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv) ||
              dq_rs < (int64_t)H * HD || dk_rs < (int64_t)H * HD || dv_rs < (int64_t)H * HD || (dq_rs | dk_rs | dv_rs) % 8 != 0))
    return fail(VLPET_E_BADARG, "attn_bwd: bad arguments");
  if (!bwd && (!out || !aligned16(out))) return fail(VLPET_E_BADARG, "attn_fwd: bad output");
  const int L = Lq > Lk ? Lq : Lk;
  return L <= 64 ? launch_attn2(bwd, a, st) : launch_attn<128>(bwd, a, st);
}

}  // namespace vlpet
