// K1 backward, fused, sm_100a ("B1"): activation gradients of the granularity-controlled PET module, large gate, in ONE
// launch; the four token-contracted weight-gradient GEMMs run right after it in vlpet_wgrad_sm100.cu ("B2") from the
// small intermediates this kernel leaves behind.  Math: SURVEY Appendix A / oracle gated_pet_bwd
// (my_transformers/modeling_bart.py:1145-1155, 1195-1209, 1256-1260 differentiated).
//
// Persistent CTAs, one per SM, each walks 128-token tiles.  Per tile (nothing is saved by the forward: everything is
// recomputed from x1 / x2):
//   phase 1  (K = d)        A = x2 Wd^T, P = x1 Gd^T                                  tcgen05, fp32 accum in TMEM (kept)
//   epi 1                   z = gelu_new(A+bd), q = gelu_new(P+gbd)                  -> swizzled smem (bf16) + scratch
//   phase 2  (64-col chunks) U_c = z Wu_c^T, T_c = q Gu_c^T                           tcgen05
//   epi 2                   y1 = k x2 + a(U+bu), G = sig(T+gbu), dh = s m dout,
//                           du = a dh G, dt = dh y1 G(1-G)   (add-gate: du = a dh, dt = dh G(1-G))
//                                                                                     -> in place over x2_c / dout_c in smem
//            MMA            dz += du_c Wu_c, dq += dt_c Gu_c   (B operands MN-major from the SAME smem tiles of Wu_c/Gu_c)
//            store          du_c, dt_c -> scratch (TMA store) for the weight-gradient GEMMs
//   epi 3                   da = dz gelu_new'(A+bd), dp = dq gelu_new'(P+gbd)        -> smem (over z) + scratch; dbd, dgbd
//   phase 3  (64-col chunks) T_c again (only G is needed; recompute beats keeping [128 x 768] gates),
//                           DX2_c = da Wd[:,c], DX1_c = dp Gd[:,c]                    tcgen05 (B MN-major from Wd_c/Gd_c tiles)
//   epi 4                   dx2 = k dh G + DX2, dx1 = dout + DX1                      -> smem -> TMA store
// Warp roles: 0 TMA producer (activations), 3 TMA producer (weights), 1 MMA issuer (TMEM owner), 2 TMA-store issuer,
// 4..19 epilogue (warp%4 = TMEM lane quarter; cg = (warp-4)/4: adapter | gate branch and column half in epi 1/3, 16 of the
// chunk's 64 columns in epi 2/4).  Epilogue arithmetic is on packed fp32 pairs (FFMA2); the pre-scaled fp32 bias tables
// live in the unused 64-byte halves of the swizzled dp block (BCfg::OFF_TAB).
#include <cstdlib>
#include <type_traits>

#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {
int make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                  uint32_t box_cols, bool weight);
namespace {

// -DVLPET_B1_FAST: role loops under elect.sync (the fast issue path under investigation: profiles/r2_sanitizer_*)
#ifdef VLPET_B1_FAST
#define VLPET_B1_ISSUER ptx::elect_one()
#else
#define VLPET_B1_ISSUER (lane == 0)
#endif
constexpr int TILE_M = 128;
constexpr int CH = 64;
constexpr int SX = 2;
constexpr int SW = 2;
constexpr int XCH_BYTES = TILE_M * CH * 2;  // 16 KB
constexpr int NUM_THREADS = 640;   // 4 role warps + 16 epilogue warps
constexpr int EPI_THREADS = 512;
constexpr int TM_A = 0, TM_P = 96, TM_DZ = 192, TM_DQ = 288, TM_UT = 384;  // phase 1-2 TMEM columns
constexpr int TM_ACC = 128, ACC_STRIDE = 192;                              // phase 3: {T, DX2, DX1} x 2 buffers

template <int R>
struct BCfg {
  static constexpr int KB = (R + 63) / 64;
  static constexpr int WA_BYTES = R * CH * 2;        // [R x 64] chunk of Wd / Gd
  static constexpr int WB_BYTES = KB * CH * CH * 2;  // [64 x KB*64] chunk of Wu / Gu
  static constexpr int W3 = WB_BYTES + 2 * WA_BYTES;
  static constexpr int W12 = (2 * WA_BYTES > 2 * WB_BYTES) ? 2 * WA_BYTES : 2 * WB_BYTES;
  static constexpr int WSLOT = W3 > W12 ? W3 : W12;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_W = OFF_X + SX * 2 * XCH_BYTES;
  static constexpr int OFF_Z0 = OFF_W + SW * WSLOT;                      // z (later da), columns 0..63
  static constexpr int OFF_Q0 = OFF_Z0 + XCH_BYTES;                      // q, columns 0..63
  static constexpr int OFF_ZQ1 = OFF_Q0 + XCH_BYTES;                     // columns 64..95: z/da in bytes 0..63 of each row, q in 64..127
  static constexpr int OFF_DP0 = OFF_ZQ1 + (KB == 2 ? XCH_BYTES : 0);
  static constexpr int OFF_DP1 = OFF_DP0 + XCH_BYTES;
  // fp32 tables alpha*bu[d] | 0.5*gbu[d] for epilogues 2 / 4, 16 columns (64 B) per entry.  With KB == 2 the dp block of
  // columns 64..95 only occupies one 64-byte half of each 128-byte row (which half depends on the row's swizzle phase);
  // entry gi lives in the OTHER half of row gi of that block -- shared memory is full otherwise.  KB == 1: own block.
  static constexpr int OFF_TAB = OFF_DP1;
  static constexpr int OFF_BAR = OFF_DP1 + XCH_BYTES;
  static constexpr int MAX_D = 1024;                                             // 2 * d / 16 table entries <= 128 rows
  static constexpr int SMEM_BYTES = OFF_BAR + 1024 + 2 * R * 4 + 256 + 1024;   // barriers | fp32 bd, gbd | slack + alignment
  static_assert(R <= 96 && R % 16 == 0, "fused backward covers ranks up to 96");
  static_assert(WSLOT % 1024 == 0 && WA_BYTES % 1024 == 0, "swizzle atoms must stay 1024-byte aligned");
};

// Optional phase-timestamp trace (tools/trace_k1.py --bwd): thread 128 of every CTA stamps %globaltimer at the phase
// boundaries of its first tile into [cta][128] slots.
static unsigned long long* g_trace_b = nullptr;   // host copy; travels to the kernel as BParams::trace
__device__ __forceinline__ unsigned long long gtimer_b() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VLPET_TRACE_B(slot)                                                                                               \
  do {                                                                                                                    \
    if (p.trace && threadIdx.x == 128 && tile == blockIdx.x && (slot) < 128) p.trace[blockIdx.x * 128 + (slot)] = gtimer_b(); \
  } while (0)

struct BParams {
  int64_t M;
  int d, r, rg;
  int add_gate;
  float s, alpha, kappa;
  const __nv_bfloat16 *bd, *bu, *gbd, *gbu;
  __nv_bfloat16 *zs, *qs, *das, *dps;   // scratch [M, pz] / [M, pq]
  int pz, pq;                           // scratch row pitches (elements)
  float *dbd, *dgbd;                    // fp32 bias gradients (accumulated into) or nullptr
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t thr16;
  float inv_keep;
  unsigned long long* trace;            // developer hook (tools/trace_k1_bwd.py), normally null
  uint32_t* dbg;                        // host-mapped trap record (ptx::mbar_wait_dbg)
};

enum { B_XFULL = 0, B_XEMPTY = B_XFULL + SX, B_WFULL = B_XEMPTY + SX, B_WEMPTY = B_WFULL + SW, B_APFULL = B_WEMPTY + SW,
       B_ZQFULL, B_UTFULL, B_UTEMPTY, B_DUDT, B_DZQDONE = B_DUDT + SX, B_DZFULL = B_DZQDONE + SX, B_DAPFULL,
       B_ACCFULL, B_ACCEMPTY = B_ACCFULL + 2, B_OUTRDY = B_ACCEMPTY + 2, B_COUNT = B_OUTRDY + SX };

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
using namespace ptx;   // f2 helpers (packed fp32 pairs)
// gelu_new(v) and gelu_new'(v) for a pair: 0.5 v (1 + th), th = tanh(c (v + 0.044715 v^3))
__device__ __forceinline__ void gelu_new_both2(f2 v, f2& g, f2& dg) {
  const float c = 0.7978845608028654f, ck = 0.7978845608028654f * 0.044715f;
  const f2 c2 = mk2(c, c), ck2 = mk2(ck, ck), ck3 = mk2(3.0f * ck, 3.0f * ck), half = mk2(0.5f, 0.5f), one = mk2(1.0f, 1.0f);
  const f2 v2 = mul2(v, v);
  const f2 th = tanh2(mul2(v, fma2(ck2, v2, c2)));
  const f2 hv = mul2(half, v);
  g = fma2(hv, th, hv);
  const f2 omt = fma2(mul2(th, mk2(-1.0f, -1.0f)), th, one);          // 1 - th^2
  dg = fma2(mul2(hv, omt), fma2(ck3, v2, c2), fma2(half, th, half));
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// Column sums over the 32 lanes of a warp: lane l enters with its row's 32 values, leaves with sum_rows column l in v[0].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// GATED = false is the ungated form used for the decoder value parallel adapter (K2): out = x1 + alpha*(Up(gelu_new(Down x2)))
// -> dx2 = (alpha * (dout Wu) * gelu_new'(A)) Wd, no gate branch, no U/T recompute, no du/dt scratch (du = alpha*dout).
// Same over 16 columns: every lane enters with its row's 16 values; lanes l and l^16 end up with sum_rows column (l & 15).
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int R, bool GATED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                    const __grid_constant__ CUtensorMap tm_dout, const __grid_constant__ CUtensorMap tm_dx1,
                    const __grid_constant__ CUtensorMap tm_dx2, const __grid_constant__ CUtensorMap tm_du,
                    const __grid_constant__ CUtensorMap tm_dt, const __grid_constant__ CUtensorMap tm_wd,
                    const __grid_constant__ CUtensorMap tm_gd, const __grid_constant__ CUtensorMap tm_wu,
                    const __grid_constant__ CUtensorMap tm_gu, const BParams p) {
  using C = BCfg<R>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int nkc = p.d / CH;
  const int64_t num_tiles = (p.M + TILE_M - 1) / TILE_M;
  const bool mulgate = p.add_gate == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SX; ++i) {
      ptx::mbar_init(bar(B_XFULL + i), 1); ptx::mbar_init(bar(B_XEMPTY + i), 1);
      ptx::mbar_init(bar(B_DUDT + i), EPI_THREADS); ptx::mbar_init(bar(B_DZQDONE + i), 1);
      ptx::mbar_init(bar(B_OUTRDY + i), EPI_THREADS);
    }
    for (int i = 0; i < SW; ++i) { ptx::mbar_init(bar(B_WFULL + i), 1); ptx::mbar_init(bar(B_WEMPTY + i), 1); }
    ptx::mbar_init(bar(B_APFULL), 1);
    ptx::mbar_init(bar(B_ZQFULL), EPI_THREADS);
    ptx::mbar_init(bar(B_UTFULL), 1);
    ptx::mbar_init(bar(B_UTEMPTY), EPI_THREADS);
    ptx::mbar_init(bar(B_DZFULL), 1);
    ptx::mbar_init(bar(B_DAPFULL), EPI_THREADS);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar(B_ACCFULL + i), 1); ptx::mbar_init(bar(B_ACCEMPTY + i), EPI_THREADS); }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_dout); }
  if (warp == 3 && lane == 0) {
    ptx::prefetch_tmap(&tm_wd); ptx::prefetch_tmap(&tm_gd); ptx::prefetch_tmap(&tm_wu); ptx::prefetch_tmap(&tm_gu);
  }
  if (warp == 2 && lane == 0) {
    ptx::prefetch_tmap(&tm_dx1); ptx::prefetch_tmap(&tm_dx2); ptx::prefetch_tmap(&tm_du); ptx::prefetch_tmap(&tm_dt);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // K-major A-operand descriptor start of the k-th 16-element K step of z/da (which=0), q (1), dp (2)
  auto small_a = [&](int which, int ks) -> uint32_t {
    const uint32_t b0 = smem_base + (which == 0 ? C::OFF_Z0 : (which == 1 ? C::OFF_Q0 : C::OFF_DP0));
    if (ks < 4) return b0 + (uint32_t)ks * 32u;
    const uint32_t b1 = smem_base + (which == 2 ? C::OFF_DP1 : C::OFF_ZQ1) + (which == 1 ? 64u : 0u);
    return b1 + (uint32_t)(ks - 4) * 32u;
  };

  if (warp == 0) {
    // ===================================== TMA producer: activations =====================================
    if (VLPET_B1_ISSUER) {
      uint32_t xi = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = (int)(tile * TILE_M);
        for (int ph = 0; ph < 3; ++ph) {
          for (int c = 0; c < nkc; ++c, ++xi) {
            const uint32_t sx = xi % SX;
            ptx::mbar_wait_dbg(bar(B_XEMPTY + sx), ((xi / SX) & 1) ^ 1, p.dbg, __LINE__);
            const uint32_t xdst = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
            if (ph == 0) {
              ptx::mbar_arrive_expect_tx(bar(B_XFULL + sx), (GATED ? 2 : 1) * XCH_BYTES);
              if (GATED) ptx::tma_load_2d(xdst, &tm_x1, c * CH, row0, bar(B_XFULL + sx));
              ptx::tma_load_2d(xdst + XCH_BYTES, &tm_x2, c * CH, row0, bar(B_XFULL + sx));
            } else if (ph == 1) {
              ptx::mbar_arrive_expect_tx(bar(B_XFULL + sx), (GATED ? 2 : 1) * XCH_BYTES);
              if (GATED) {
                ptx::tma_load_2d(xdst, &tm_x2, c * CH, row0, bar(B_XFULL + sx));
                ptx::tma_load_2d(xdst + XCH_BYTES, &tm_dout, c * CH, row0, bar(B_XFULL + sx));
              } else {
                ptx::tma_load_2d(xdst, &tm_dout, c * CH, row0, bar(B_XFULL + sx));   // du = alpha*dout: consumed as is
              }
            } else if (GATED) {
              ptx::mbar_arrive_expect_tx(bar(B_XFULL + sx), XCH_BYTES);
              ptx::tma_load_2d(xdst, &tm_dout, c * CH, row0, bar(B_XFULL + sx));
            } else {
              ptx::mbar_arrive(bar(B_XFULL + sx));   // the stage is only the staging buffer of dx2_c
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA producer: weights (always L2 hits) =====================================
    if (VLPET_B1_ISSUER) {
      uint32_t wi = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int ph = 0; ph < 3; ++ph) {
          for (int c = 0; c < nkc; ++c, ++wi) {
            const uint32_t sw = wi % SW;
            ptx::mbar_wait_dbg(bar(B_WEMPTY + sw), ((wi / SW) & 1) ^ 1, p.dbg, __LINE__);
            const uint32_t wdst = smem_base + C::OFF_W + sw * C::WSLOT;
            if (ph == 0) {
              ptx::mbar_arrive_expect_tx(bar(B_WFULL + sw), (GATED ? 2 : 1) * C::WA_BYTES);
              ptx::tma_load_2d(wdst, &tm_wd, c * CH, 0, bar(B_WFULL + sw));
              if (GATED) ptx::tma_load_2d(wdst + C::WA_BYTES, &tm_gd, c * CH, 0, bar(B_WFULL + sw));
            } else if (ph == 1) {
              ptx::mbar_arrive_expect_tx(bar(B_WFULL + sw), (GATED ? 2 : 1) * C::WB_BYTES);
#pragma unroll
              for (int kb = 0; kb < C::KB; ++kb) {
                ptx::tma_load_2d(wdst + kb * (CH * CH * 2), &tm_wu, kb * CH, c * CH, bar(B_WFULL + sw));
                if (GATED)
                  ptx::tma_load_2d(wdst + C::WB_BYTES + kb * (CH * CH * 2), &tm_gu, kb * CH, c * CH, bar(B_WFULL + sw));
              }
            } else {
              ptx::mbar_arrive_expect_tx(bar(B_WFULL + sw), ((GATED && mulgate) ? C::WB_BYTES : 0) + (GATED ? 2 : 1) * C::WA_BYTES);
              if (GATED && mulgate) {
#pragma unroll
                for (int kb = 0; kb < C::KB; ++kb)
                  ptx::tma_load_2d(wdst + kb * (CH * CH * 2), &tm_gu, kb * CH, c * CH, bar(B_WFULL + sw));
              }
              ptx::tma_load_2d(wdst + C::WB_BYTES, &tm_wd, c * CH, 0, bar(B_WFULL + sw));
              if (GATED) ptx::tma_load_2d(wdst + C::WB_BYTES + C::WA_BYTES, &tm_gd, c * CH, 0, bar(B_WFULL + sw));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (VLPET_B1_ISSUER) {   // NOT elect_one() by default: with the fast issue path B1 faults sporadically
      constexpr uint32_t IDESC_AP = ptx::umma_idesc_bf16_m128(R);                       // A/P: N = R, K-major x K-major
      constexpr uint32_t IDESC_UT = ptx::umma_idesc_bf16_m128(CH);                      // U/T: N = 64
      constexpr uint32_t IDESC_DZ = ptx::umma_idesc_bf16_m128_major(R, 0u, 1u);         // dz/dq: B MN-major, N = R
      constexpr uint32_t IDESC_DX = ptx::umma_idesc_bf16_m128_major(CH, 0u, 1u);        // DX: B MN-major, N = 64
      uint32_t xi = 0, wi = 0, ui = 0, p2i = 0, ai = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        // the phase-3 accumulators of the previous tile overlap P / DZ / DQ / UT: wait until the epilogue drained them
        if (ai > 0) {
          for (uint32_t b = 0; b < 2; ++b) {
            const uint32_t nb = (ai + 1 - b) >> 1;
            if (nb > 0) ptx::mbar_wait_dbg(bar(B_ACCEMPTY + b), (nb - 1) & 1, p.dbg, __LINE__);
          }
          ptx::tc_fence_after();
        }
        // ---- phase 1
        for (int c = 0; c < nkc; ++c, ++xi, ++wi) {
          const uint32_t sx = xi % SX, sw = wi % SW;
          ptx::mbar_wait_dbg(bar(B_XFULL + sx), (xi / SX) & 1, p.dbg, __LINE__);
          ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t x1s = smem_base + C::OFF_X + sx * (2 * XCH_BYTES), x2s = x1s + XCH_BYTES;
          const uint32_t wds = smem_base + C::OFF_W + sw * C::WSLOT, gds = wds + C::WA_BYTES;
#pragma unroll
          for (int ks = 0; ks < CH / 16; ++ks) {
            const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
            ptx::umma_bf16_ss(tmem_base + TM_A, ptx::umma_desc_kmajor_sw128(x2s + ks * 32),
                              ptx::umma_desc_kmajor_sw128(wds + ks * 32), IDESC_AP, acc);
            if (GATED)
              ptx::umma_bf16_ss(tmem_base + TM_P, ptx::umma_desc_kmajor_sw128(x1s + ks * 32),
                                ptx::umma_desc_kmajor_sw128(gds + ks * 32), IDESC_AP, acc);
          }
          ptx::umma_commit(bar(B_XEMPTY + sx));
          ptx::umma_commit(bar(B_WEMPTY + sw));
        }
        ptx::umma_commit(bar(B_APFULL));
        // ---- phase 2
        if (!GATED) {
          for (int c = 0; c < nkc; ++c, ++xi, ++wi) {
            const uint32_t sx = xi % SX, sw = wi % SW;
            ptx::mbar_wait_dbg(bar(B_XFULL + sx), (xi / SX) & 1, p.dbg, __LINE__);
            ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
            ptx::tc_fence_after();
            const uint32_t dos = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
            const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT;
#pragma unroll
            for (int ks = 0; ks < CH / 16; ++ks)
              ptx::umma_bf16_ss(tmem_base + TM_DZ, ptx::umma_desc_kmajor_sw128(dos + ks * 32),
                                ptx::umma_desc_mnmajor_sw128(wus + ks * 2048, CH * CH * 2), IDESC_DZ, (c > 0 || ks > 0) ? 1u : 0u);
            ptx::umma_commit(bar(B_XEMPTY + sx));
            ptx::umma_commit(bar(B_WEMPTY + sw));
          }
          ptx::umma_commit(bar(B_DZFULL));
        }
        if (GATED) {
        ptx::mbar_wait_dbg(bar(B_ZQFULL), ti & 1, p.dbg, __LINE__);
        ptx::tc_fence_after();
        const uint32_t xi2 = xi, wi2 = wi, p2i0 = p2i;
        auto issue_dzdq = [&](int cc) {
          const uint32_t sx = (xi2 + cc) % SX, sw = (wi2 + cc) % SW, k = (p2i0 + cc) % SX;
          ptx::mbar_wait_dbg(bar(B_DUDT + k), ((p2i0 + cc) / SX) & 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t dus = smem_base + C::OFF_X + sx * (2 * XCH_BYTES), dts = dus + XCH_BYTES;
          const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT, gus = wus + C::WB_BYTES;
#pragma unroll
          for (int ks = 0; ks < CH / 16; ++ks) {
            const uint32_t acc = (cc > 0 || ks > 0) ? 1u : 0u;
            ptx::umma_bf16_ss(tmem_base + TM_DZ, ptx::umma_desc_kmajor_sw128(dus + ks * 32),
                              ptx::umma_desc_mnmajor_sw128(wus + ks * 2048, CH * CH * 2), IDESC_DZ, acc);
            ptx::umma_bf16_ss(tmem_base + TM_DQ, ptx::umma_desc_kmajor_sw128(dts + ks * 32),
                              ptx::umma_desc_mnmajor_sw128(gus + ks * 2048, CH * CH * 2), IDESC_DZ, acc);
          }
          ptx::umma_commit(bar(B_DZQDONE + k));
          ptx::umma_commit(bar(B_WEMPTY + sw));
        };
        for (int c = 0; c < nkc; ++c, ++xi, ++wi, ++ui, ++p2i) {
          const uint32_t sw = wi % SW;
          ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
          ptx::mbar_wait_dbg(bar(B_UTEMPTY), (ui & 1) ^ 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT, gus = wus + C::WB_BYTES;
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks) {
            const uint32_t kb = ks / 4, kin = ks % 4;
            ptx::umma_bf16_ss(tmem_base + TM_UT, ptx::umma_desc_kmajor_sw128(small_a(0, ks)),
                              ptx::umma_desc_kmajor_sw128(wus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
            ptx::umma_bf16_ss(tmem_base + TM_UT + CH, ptx::umma_desc_kmajor_sw128(small_a(1, ks)),
                              ptx::umma_desc_kmajor_sw128(gus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
          }
          ptx::umma_commit(bar(B_UTFULL));
          if (c > 0) issue_dzdq(c - 1);
        }
        issue_dzdq(nkc - 1);
        ptx::umma_commit(bar(B_DZFULL));
        }
        // ---- phase 3
        ptx::mbar_wait_dbg(bar(B_DAPFULL), ti & 1, p.dbg, __LINE__);
        ptx::tc_fence_after();
        for (int c = 0; c < nkc; ++c, ++xi, ++wi, ++ai) {
          const uint32_t sw = wi % SW, b = ai & 1;
          ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
          ptx::mbar_wait_dbg(bar(B_ACCEMPTY + b), ((ai >> 1) & 1) ^ 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t gus = smem_base + C::OFF_W + sw * C::WSLOT, wds = gus + C::WB_BYTES, gds = wds + C::WA_BYTES;
          const uint32_t tacc = tmem_base + TM_ACC + b * ACC_STRIDE;
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks) {
            const uint32_t kb = ks / 4, kin = ks % 4;
            if (GATED && mulgate)
              ptx::umma_bf16_ss(tacc, ptx::umma_desc_kmajor_sw128(small_a(1, ks)),
                                ptx::umma_desc_kmajor_sw128(gus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
            ptx::umma_bf16_ss(tacc + CH, ptx::umma_desc_kmajor_sw128(small_a(0, ks)),
                              ptx::umma_desc_mnmajor_sw128(wds + ks * 2048, C::WA_BYTES), IDESC_DX, ks > 0);
            if (GATED)
              ptx::umma_bf16_ss(tacc + 2 * CH, ptx::umma_desc_kmajor_sw128(small_a(2, ks)),
                                ptx::umma_desc_mnmajor_sw128(gds + ks * 2048, C::WA_BYTES), IDESC_DX, ks > 0);
          }
          ptx::umma_commit(bar(B_WEMPTY + sw));
          ptx::umma_commit(bar(B_ACCFULL + b));
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== TMA store issuer =====================================
    if (VLPET_B1_ISSUER) {
      uint32_t xi = 0, p2i = 0, p3i = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = (int)(tile * TILE_M);
        xi += nkc;  // phase-1 stages are released by the MMA warp
        if (!GATED) xi += nkc;  // ungated: phase-2 stages are released by the MMA warp as well
        for (int c = 0; GATED && c < nkc; ++c, ++xi, ++p2i) {
          const uint32_t sx = xi % SX, k = p2i % SX;
          ptx::mbar_wait_dbg(bar(B_DUDT + k), (p2i / SX) & 1, p.dbg, __LINE__);
          const uint32_t src = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
          ptx::tma_store_2d(&tm_du, src, c * CH, row0);
          ptx::tma_store_2d(&tm_dt, src + XCH_BYTES, c * CH, row0);
          ptx::tma_store_commit();
          ptx::mbar_wait_dbg(bar(B_DZQDONE + k), (p2i / SX) & 1, p.dbg, __LINE__);   // dz/dq MMAs finished reading du_c / dt_c
          ptx::tma_store_wait_read0();
          ptx::mbar_arrive(bar(B_XEMPTY + sx));
        }
        for (int c = 0; c < nkc; ++c, ++xi, ++p3i) {
          const uint32_t sx = xi % SX, k = p3i % SX;
          ptx::mbar_wait_dbg(bar(B_OUTRDY + k), (p3i / SX) & 1, p.dbg, __LINE__);
          const uint32_t src = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
          if (GATED) ptx::tma_store_2d(&tm_dx1, src, c * CH, row0);
          ptx::tma_store_2d(&tm_dx2, src + XCH_BYTES, c * CH, row0);
          ptx::tma_store_commit();
          ptx::tma_store_wait_read0();
          ptx::mbar_arrive(bar(B_XEMPTY + sx));
        }
      }
      ptx::tma_store_wait_all0();
    }
  } else if (warp >= 4) {
    // ===================================== epilogue warps =====================================
    // 16 warps: warp%4 = TMEM lane quarter, cg = (warp-4)/4.  Epilogues 1 / 3: branch = cg/2 (adapter | gate), column half =
    // cg%2; epilogues 2 / 4: columns [cg*16, cg*16+16) of the 64-column chunk.
    const int quarter = warp % 4;
    const int cg = (warp - 4) / 4;
    const int branch = cg >> 1;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
    constexpr int HALF = R / 2;                // columns per warp in epilogues 1 / 3 (a multiple of 16)
    const int jbeg = (cg & 1) * HALF;
    // fp32 copies of the down-projection biases in shared memory (broadcast reads instead of scalar global loads)
    const uint32_t sb_base = bar_base + 1024;   // [2][R] floats behind the barrier block
    for (int i = threadIdx.x - 128; i < 2 * R; i += EPI_THREADS) {
      const int br = i / R, j = i % R;
      const int rr = br ? p.rg : p.r;
      const __nv_bfloat16* src = br ? p.gbd : p.bd;
      const float v = (j < rr && (GATED || br == 0)) ? __bfloat162float(src[j]) : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb_base + 4u * (uint32_t)i), "f"(v) : "memory");
    }
    // alpha*bu and 0.5*gbu, 16 columns per 64-byte entry (layout: BCfg::OFF_TAB)
    auto tab_addr = [&](int gi) -> uint32_t { return smem_base + C::OFF_TAB + (uint32_t)gi * 128u + ((gi & 4) ? 0u : 64u); };
    const int ngrp = p.d / 16;
    if (GATED) {
      for (int i = threadIdx.x - 128; i < 2 * p.d; i += EPI_THREADS) {
        const int col = i < p.d ? i : i - p.d;
        const float v = i < p.d ? p.alpha * __bfloat162float(p.bu[col]) : 0.5f * __bfloat162float(p.gbu[col]);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tab_addr(i >> 4) + 4u * (uint32_t)(i & 15)), "f"(v) : "memory");
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // epilogue warps only
    const f2 half2 = mk2(0.5f, 0.5f), kappa2 = mk2(p.kappa, p.kappa), alpha2 = mk2(p.alpha, p.alpha);
    const f2 s2 = mk2(p.s, p.s);
    const float s_keep = p.s * p.inv_keep;
    // smem address of the 16-byte group holding columns [k, k+8) of this thread's row in z/da (which=0), q (1), dp (2)
    auto small_addr = [&](int which, int k) -> uint32_t {
      if (k < 64) {
        const uint32_t b0 = smem_base + (which == 0 ? C::OFF_Z0 : (which == 1 ? C::OFF_Q0 : C::OFF_DP0));
        return b0 + (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ swz) << 4);
      }
      const uint32_t b1 = smem_base + (which == 2 ? C::OFF_DP1 : C::OFF_ZQ1);
      const uint32_t c16 = (((uint32_t)k - 64u) >> 3) + (which == 1 ? 4u : 0u);
      return b1 + (uint32_t)row * 128u + ((c16 ^ swz) << 4);
    };
    uint32_t xi = 0, ui = 0, p2i = 0, ai = 0, ti = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const int64_t grow = tile * TILE_M + row;
      const bool row_ok = grow < p.M;
      // ---- epilogue 1 (APFULL also implies that every MMA of the previous tile, which read z/q/da/dp, has completed)
      VLPET_TRACE_B(0);
      ptx::mbar_wait_dbg(bar(B_APFULL), ti & 1, p.dbg, __LINE__);
      VLPET_TRACE_B(1);
      ptx::tc_fence_after();
      if (GATED || branch == 0) {
        const uint32_t tsrc = lane_addr + (branch ? TM_P : TM_A);
        const int rr = branch ? p.rg : p.r;
        __nv_bfloat16* srow = (branch ? p.qs + grow * p.pq : p.zs + grow * p.pz);
#pragma unroll
        for (int jj = 0; jj < HALF; jj += 16) {
          const int j0 = jbeg + jj;
          uint32_t v[16], dgv[16];
          ptx::tmem_ld_32x32b_x16(tsrc + j0, v);
          f2 bias[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) lds_f2x2(sb_base + 4u * (uint32_t)(branch * R + j0 + e * 4), bias[2 * e], bias[2 * e + 1]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f2 z2, d2;
              gelu_new_both2(add2(mk2u(v[g * 8 + e * 2], v[g * 8 + e * 2 + 1]), bias[g * 4 + e]), z2, d2);
              o[e] = pack2(z2);
              un2u(d2, dgv[g * 8 + e * 2], dgv[g * 8 + e * 2 + 1]);
            }
            const int k = j0 + g * 8;
            sts128(small_addr(branch, k), o);
            if (row_ok && k < rr) *reinterpret_cast<uint4*>(srow + k) = make_uint4(o[0], o[1], o[2], o[3]);
          }
          ptx::tmem_st_32x32b_x16(tsrc + j0, dgv);   // gelu_new'(pre-activation) replaces the pre-activation in TMEM: epilogue 3 only multiplies
        }
        ptx::tmem_st_wait();
        if (row_ok && (cg & 1) == 1) {  // ones column (bias-gradient trick of the weight-gradient GEMM) + zero pad up to the pitch
          const uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(srow + rr) = one;
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(B_ZQFULL));
      VLPET_TRACE_B(2);
      xi += nkc;
      // ---- epilogue 2, per 64-column chunk: du, dt
      if (!GATED) xi += nkc;
      for (int c = 0; GATED && c < nkc; ++c, ++xi, ++ui, ++p2i) {
        const uint32_t sx = xi % SX;
        ptx::mbar_wait_dbg(bar(B_UTFULL), ui & 1, p.dbg, __LINE__);
        VLPET_TRACE_B(3 + 3 * c);
        ptx::tc_fence_after();
        uint32_t u[16], t[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + TM_UT + cg * 16, u);
        ptx::tmem_ld_32x32b_x16(lane_addr + TM_UT + CH + cg * 16, t);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(B_UTEMPTY));
        ptx::mbar_wait_dbg(bar(B_XFULL + sx), (xi / SX) & 1, p.dbg, __LINE__);
        VLPET_TRACE_B(4 + 3 * c);
        const uint32_t x2row = smem_base + C::OFF_X + sx * (2 * XCH_BYTES) + (uint32_t)row * 128u;
        const uint32_t dorow = x2row + XCH_BYTES;
        const int col0 = c * CH + cg * 16;
        const int64_t idx0 = grow * p.d + col0;
        const uint32_t tabu = tab_addr(c * 4 + cg), thgb = tab_addr(ngrp + c * 4 + cg);
        auto group2 = [&](auto drop_tag, int g) {
          constexpr bool DROP = decltype(drop_tag)::value;
          const uint32_t off = (((uint32_t)(cg * 2 + g)) ^ swz) << 4;
          uint32_t xv[4], dv[4], ou[4], ot[4];
          lds128(x2row + off, xv);
          lds128(dorow + off, dv);
          f2 abu[4], hgb[4];
          lds_f2x2(tabu + g * 32, abu[0], abu[1]);
          lds_f2x2(tabu + g * 32 + 16, abu[2], abu[3]);
          lds_f2x2(thgb + g * 32, hgb[0], hgb[1]);
          lds_f2x2(thgb + g * 32 + 16, hgb[2], hgb[3]);
          uint64_t h0 = 0, h1 = 0;
          if (DROP) {
            h0 = drop_hash4(seed_eff, (uint64_t)(idx0 + g * 8) >> 2);
            h1 = drop_hash4(seed_eff, ((uint64_t)(idx0 + g * 8) >> 2) + 1);
          }
          const f2 q2 = mk2(0.25f, 0.25f), nq2 = mk2(-0.25f, -0.25f);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            f2 sce = s2;                                                          // s * dropout mask of this pair
            if (DROP) {
              const uint32_t two = (uint32_t)((e >> 1 ? h1 : h0) >> (32 * (e & 1)));
              sce = mk2(((two & 0xffffu) >= p.thr16) ? s_keep : 0.f, ((two >> 16) >= p.thr16) ? s_keep : 0.f);
            }
            const f2 dh = mul2(sce, bf2_to_f2(dv[e]));                            // dh = s m dout
            const f2 th = tanh2(fma2(half2, mk2u(t[j], t[j + 1]), hgb[e]));      // G = 0.5 + 0.5 th
            const f2 gg = fma2(nq2, mul2(th, th), q2);                            // G (1 - G) = 0.25 (1 - th^2)
            f2 du, dt;
            if (mulgate) {
              f2 y1 = fma2(kappa2, bf2_to_f2(xv[e]), abu[e]);                     // kappa x2 + alpha bu
              y1 = fma2(alpha2, mk2u(u[j], u[j + 1]), y1);                        // + alpha U
              du = mul2(mul2(alpha2, dh), fma2(half2, th, half2));                // alpha dh G
              dt = mul2(mul2(dh, y1), gg);
            } else {
              du = mul2(alpha2, dh);
              dt = mul2(dh, gg);
            }
            ou[e] = pack2(du);
            ot[e] = pack2(dt);
          }
          sts128(x2row + off, ou);
          sts128(dorow + off, ot);
        };
        if (p.thr16) { group2(std::true_type{}, 0); group2(std::true_type{}, 1); }
        else { group2(std::false_type{}, 0); group2(std::false_type{}, 1); }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(bar(B_DUDT + (p2i % SX)));
        VLPET_TRACE_B(5 + 3 * c);
      }
      // ---- epilogue 3: da = dz * gelu_new'(A + bd) (branch 0), dp = dq * gelu_new'(P + gbd) (branch 1)
      VLPET_TRACE_B(40);
      ptx::mbar_wait_dbg(bar(B_DZFULL), ti & 1, p.dbg, __LINE__);
      VLPET_TRACE_B(41);
      ptx::tc_fence_after();
      if (GATED || branch == 0) {
        const float dzscale = GATED ? 1.0f : p.alpha;   // ungated: du = alpha*dout was fed unscaled
        const uint32_t tpre = lane_addr + (branch ? TM_P : TM_A);
        const uint32_t tdz = lane_addr + (branch ? TM_DQ : TM_DZ);
        const int rr = branch ? p.rg : p.r;
        __nv_bfloat16* srow = (branch ? p.dps + grow * p.pq : p.das + grow * p.pz);
#pragma unroll
        for (int jj = 0; jj < HALF; jj += 16) {
          const int j0 = jbeg + jj;
          uint32_t a[16], dz[16];
          ptx::tmem_ld_32x32b_x16(tpre + j0, a);     // gelu_new'(A + bd), stored by epilogue 1
          ptx::tmem_ld_32x32b_x16(tdz + j0, dz);
          ptx::tmem_ld_wait();
          const f2 dzs2 = mk2(dzscale, dzscale);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = g * 8 + e * 2;
              const f2 da = mul2(mk2u(dz[j], dz[j + 1]), mk2u(a[j], a[j + 1]));
              o[e] = pack2(GATED ? da : mul2(dzs2, da));
            }
            const int k = j0 + g * 8;
            sts128(small_addr(branch ? 2 : 0, k), o);
            if (row_ok && k < rr) *reinterpret_cast<uint4*>(srow + k) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(B_DAPFULL));
      VLPET_TRACE_B(42);
      // ---- epilogue 4, per 64-column chunk: dx1, dx2
      for (int c = 0; c < nkc; ++c, ++xi, ++ai) {
        const uint32_t sx = xi % SX, b = ai & 1;
        ptx::mbar_wait_dbg(bar(B_ACCFULL + b), (ai >> 1) & 1, p.dbg, __LINE__);
        VLPET_TRACE_B(43 + 3 * c);
        ptx::tc_fence_after();
        const uint32_t tacc = lane_addr + TM_ACC + b * ACC_STRIDE + cg * 16;
        uint32_t t[16], g2[16], g1[16];
        if (GATED && mulgate) ptx::tmem_ld_32x32b_x16(tacc, t);
        ptx::tmem_ld_32x32b_x16(tacc + CH, g2);
        if (GATED) ptx::tmem_ld_32x32b_x16(tacc + 2 * CH, g1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        // The XFULL wait comes BEFORE the ACCEMPTY arrive: the MMA warp waits XFULL itself for the phase-1 entries of the
        // next tile, and a parity wait is only meaningful once the previous phase of that barrier has completed.  With the
        // arrive first, the MMA warp could reach the next tile while this chunk's dout load was still in flight, see
        // "previous-previous phase complete" as "complete", consume a stale slot and release it a second time
        // (profiles/r2_b1_fault_rootcause.md: the sporadic launch failure of round 1).
        ptx::mbar_wait_dbg(bar(B_XFULL + sx), (xi / SX) & 1, p.dbg, __LINE__);
        ptx::mbar_arrive(bar(B_ACCEMPTY + b));
        VLPET_TRACE_B(44 + 3 * c);
        const uint32_t dorow = smem_base + C::OFF_X + sx * (2 * XCH_BYTES) + (uint32_t)row * 128u;
        const uint32_t o2row = dorow + XCH_BYTES;
        const int col0 = c * CH + cg * 16;
        const int64_t idx0 = grow * p.d + col0;
        const uint32_t thgb = tab_addr(ngrp + c * 4 + cg);
        auto group4 = [&](auto drop_tag, int g) {
          constexpr bool DROP = decltype(drop_tag)::value;
          const uint32_t off = (((uint32_t)(cg * 2 + g)) ^ swz) << 4;
          uint32_t dv[4] = {0, 0, 0, 0}, o1[4], o2[4];
          if (GATED) lds128(dorow + off, dv);
          f2 hgb[4] = {0, 0, 0, 0};
          if (GATED && mulgate) {
            lds_f2x2(thgb + g * 32, hgb[0], hgb[1]);
            lds_f2x2(thgb + g * 32 + 16, hgb[2], hgb[3]);
          }
          uint64_t h0 = 0, h1 = 0;
          if (DROP) {
            h0 = drop_hash4(seed_eff, (uint64_t)(idx0 + g * 8) >> 2);
            h1 = drop_hash4(seed_eff, ((uint64_t)(idx0 + g * 8) >> 2) + 1);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            const f2 g2p = mk2u(g2[j], g2[j + 1]);
            if (GATED) {
              const f2 dof = bf2_to_f2(dv[e]);
              f2 sce = s2;
              if (DROP) {
                const uint32_t two = (uint32_t)((e >> 1 ? h1 : h0) >> (32 * (e & 1)));
                sce = mk2(((two & 0xffffu) >= p.thr16) ? s_keep : 0.f, ((two >> 16) >= p.thr16) ? s_keep : 0.f);
              }
              f2 dy1 = mul2(sce, dof);                                               // dh = s m dout
              if (mulgate) {
                const f2 th = tanh2(fma2(half2, mk2u(t[j], t[j + 1]), hgb[e]));
                dy1 = mul2(dy1, fma2(half2, th, half2));                             // dh G
              }
              o2[e] = pack2(fma2(kappa2, dy1, g2p));                                 // dx2 = kappa dy1 + da Wd
              o1[e] = pack2(add2(dof, mk2u(g1[j], g1[j + 1])));                      // dx1 = dout + dp Gd
            } else {
              o2[e] = pack2(g2p);
              o1[e] = 0u;
            }
          }
          if (GATED) sts128(dorow + off, o1);
          sts128(o2row + off, o2);
        };
        if (p.thr16) { group4(std::true_type{}, 0); group4(std::true_type{}, 1); }
        else { group4(std::false_type{}, 0); group4(std::false_type{}, 1); }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(bar(B_OUTRDY + (ai % SX)));
        VLPET_TRACE_B(45 + 3 * c);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// dbd / dgbd = column sums of the da / dp scratch ([M, pitch] bf16, first ncols columns).  Kept out of the tile kernel: a
// warp-shuffle transpose-reduce there cost ~5 us per tile on its critical path (tools/trace_k1_bwd.py).
__global__ void __launch_bounds__(128) colsum_scratch_kernel(const __nv_bfloat16* __restrict__ A, int pitch, int ncols,
                                                             int64_t M, int rows_per_block, float* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= ncols) return;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  for (int64_t m = m0; m < m1; ++m) acc += __bfloat162float(A[m * pitch + c]);
  atomicAdd(out + c, acc);
}
int launch_colsum_scratch(const __nv_bfloat16* A, int pitch, int ncols, int64_t M, float* out, cudaStream_t st) {
  if (!out) return 0;
  int64_t rpb = (M + 295) / 296;
  if (rpb < 32) rpb = 32;
  const int64_t blocks = (M + rpb - 1) / rpb;
  colsum_scratch_kernel<<<(unsigned)blocks, 128, 0, st>>>(A, pitch, ncols, M, (int)rpb, out);
  VLPET_LAUNCH_OK();
  return 0;
}

struct Scratch {
  __nv_bfloat16 *zs, *qs, *das, *dps, *dus, *dts;
  int pz, pq;
  size_t bytes;
};
Scratch carve(int64_t M, int d, int r, int rg, bool gated, void* ws) {
  Scratch s;
  s.pz = (r + 1 + 7) / 8 * 8;
  s.pq = (rg + 1 + 7) / 8 * 8;
  Arena a(ws, (size_t)-1);
  s.zs = a.take<__nv_bfloat16>((size_t)M * s.pz);
  s.das = a.take<__nv_bfloat16>((size_t)M * s.pz);
  s.qs = s.dps = s.dus = s.dts = nullptr;
  if (gated) {
    s.qs = a.take<__nv_bfloat16>((size_t)M * s.pq);
    s.dps = a.take<__nv_bfloat16>((size_t)M * s.pq);
    s.dus = a.take<__nv_bfloat16>((size_t)M * d);
    s.dts = a.take<__nv_bfloat16>((size_t)M * d);
  }
  s.bytes = a.off;
  return s;
}

template <int R, bool GATED>
int launch(const VlpetK1Desc& D, const CUtensorMap* m, const BParams& p, int sms, cudaStream_t st) {
  using C = BCfg<R>;
  static bool attr_set = false;
  if (!attr_set) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(k1_bwd_sm100_kernel<R, GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = (D.M + TILE_M - 1) / TILE_M;
  const int grid = (int)(tiles < sms ? tiles : sms);
  k1_bwd_sm100_kernel<R, GATED><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9],
                                                                  m[10], p);
  VLPET_LAUNCH_OK();
  return 0;
}

int pick_R2(int r, int rg) {
  const int m = r > rg ? r : rg;
  return m <= 32 ? 32 : (m <= 64 ? 64 : (m <= 96 ? 96 : 0));
}

// shared driver of the gated (K1, large gate) and ungated (K2) backward
int run_bwd(bool gated, const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w,
            void* dx1, void* dx2, const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rg = gated ? D.rg : D.r;
  Scratch s = carve(D.M, D.d, D.r, rg, gated, ws);
  if (!ws || ws_bytes < s.bytes) return fail(VLPET_E_WORKSPACE, "bwd(fused): workspace %zu < %zu bytes", ws_bytes, s.bytes);
  const int R = pick_R2(D.r, rg);
  const int sms = device_sm_count();
  CUtensorMap m[11];
  const uint64_t M = (uint64_t)D.M, d = (uint64_t)D.d;
  VLPET_TRY(make_map_bf16(&m[0], gated ? x1 : x2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[1], x2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[2], dout, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[3], gated ? dx1 : dx2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[4], dx2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[5], gated ? (void*)s.dus : dx2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[6], gated ? (void*)s.dts : dx2, M, d, d, TILE_M, CH, false));
  VLPET_TRY(make_map_bf16(&m[7], w.Wd, (uint64_t)D.r, d, d, (uint32_t)R, CH, true));
  VLPET_TRY(make_map_bf16(&m[8], gated ? w.Gd : w.Wd, (uint64_t)rg, d, d, (uint32_t)R, CH, true));
  VLPET_TRY(make_map_bf16(&m[9], w.Wu, d, (uint64_t)D.r, (uint64_t)D.r, CH, CH, true));
  VLPET_TRY(make_map_bf16(&m[10], gated ? w.Gu : w.Wu, d, (uint64_t)rg, (uint64_t)rg, CH, CH, true));
  BParams p;
  p.M = D.M; p.d = D.d; p.r = D.r; p.rg = rg; p.add_gate = D.add_gate;
  p.s = D.s; p.alpha = D.alpha; p.kappa = D.kappa;
  p.bd = static_cast<const __nv_bfloat16*>(w.bd); p.bu = static_cast<const __nv_bfloat16*>(w.bu);
  p.gbd = static_cast<const __nv_bfloat16*>(gated ? w.gbd : w.bd); p.gbu = static_cast<const __nv_bfloat16*>(gated ? w.gbu : w.bu);
  p.zs = s.zs; p.qs = s.qs; p.das = s.das; p.dps = s.dps; p.pz = s.pz; p.pq = s.pq;
  p.dbd = G.dbd; p.dgbd = gated ? G.dgbd : nullptr;
  p.seed = D.seed;
  p.seed_dev = D.seed_dev;
  p.trace = g_trace_b;
  p.dbg = trap_buffer_dev();
  p.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  p.inv_keep = p.thr16 ? 1.0f / (1.0f - (float)p.thr16 / 65536.0f) : 1.0f;
  // developer hook (tools/run_sanitizers.sh): VLPET_DEBUG_BWD_PARTS bit 0 = run the tile kernel, bit 1 = column sums,
  // bit 2 = weight-gradient GEMM (default: all)
  static const int parts = []() { const char* e = getenv("VLPET_DEBUG_BWD_PARTS"); return e ? atoi(e) : 7; }();
  int rc = 0;
  if (!(parts & 1)) {
  } else if (gated) {
    switch (R) {
      case 32: rc = launch<32, true>(D, m, p, sms, st); break;
      case 64: rc = launch<64, true>(D, m, p, sms, st); break;
      case 96: rc = launch<96, true>(D, m, p, sms, st); break;
      default: return fail(VLPET_E_UNSUPPORTED, "bwd(fused): unsupported rank");
    }
  } else {
    switch (R) {
      case 32: rc = launch<32, false>(D, m, p, sms, st); break;
      case 64: rc = launch<64, false>(D, m, p, sms, st); break;
      case 96: rc = launch<96, false>(D, m, p, sms, st); break;
      default: return fail(VLPET_E_UNSUPPORTED, "bwd(fused): unsupported rank");
    }
  }
  if (rc) return rc;
  if (parts & 2) {
    VLPET_TRY(launch_colsum_scratch(s.das, s.pz, D.r, D.M, G.dbd, st));
    if (gated) VLPET_TRY(launch_colsum_scratch(s.dps, s.pq, rg, D.M, G.dgbd, st));
  }
  if (!(parts & 4)) return 0;
  // ---- weight gradients: dWu = du^T z (+dbu), dGu = dt^T q (+dgbu), dWd = (x2^T da)^T, dGd = (x1^T dp)^T
  //      (ungated: du = alpha*dout, so A = dout with scale alpha)
  const void* A[4]; const void* B[4]; int64_t lda[4], ldb[4]; int nbv[4], tr[4]; float* out[4]; float* bias[4]; float sc[4];
  auto run = [&](const int* which, int n, int nout) -> int {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      const int q = which[i];
      float* o = q == 0 ? G.dWu : (q == 1 ? G.dGu : (q == 2 ? G.dWd : G.dGd));
      float* b = q == 0 ? G.dbu : (q == 1 ? G.dgbu : nullptr);
      if (!o && !b) continue;
      if (!o) return fail(VLPET_E_BADARG, "bwd(fused): a bias gradient needs its weight gradient buffer");
      A[k] = q == 0 ? (gated ? (const void*)s.dus : dout) : (q == 1 ? (const void*)s.dts : (q == 2 ? x2 : x1));
      lda[k] = D.d;
      B[k] = q == 0 ? s.zs : (q == 1 ? s.qs : (q == 2 ? s.das : s.dps));
      ldb[k] = (q == 0 || q == 2) ? s.pz : s.pq;
      nbv[k] = (q < 2) ? nout + 1 : nout;
      tr[k] = q >= 2;
      out[k] = o; bias[k] = b; sc[k] = (q == 0 && !gated) ? D.alpha : 1.0f;
      ++k;
    }
    if (k == 0) return 0;
    return wgrad_sm100(k, A, lda, B, ldb, nbv, out, bias, sc, tr, D.M, D.d, nout, sms, st);
  };
  if (!gated) {
    const int ad[2] = {0, 2};
    return run(ad, 2, D.r);
  }
  if (D.r == D.rg) {
    const int all[4] = {0, 1, 2, 3};
    return run(all, 4, D.r);
  }
  const int ad[2] = {0, 2}, ga[2] = {1, 3};
  VLPET_TRY(run(ad, 2, D.r));
  return run(ga, 2, D.rg);
}

}  // namespace

int set_k1_bwd_trace(unsigned long long* dev_buf) {
  g_trace_b = dev_buf;
  return 0;
}

bool fused_k1_bwd_supported(const VlpetK1Desc& D) {
  if (D.dtype != VLPET_BF16 || D.gate != VLPET_GATE_LARGE) return false;
  if (D.d % 128 != 0 || D.d < 128 || D.d > 1024) return false;   // d <= 1024: the fp32 bias tables (BCfg::OFF_TAB)
  if (D.r % 8 != 0 || D.rg % 8 != 0 || D.r < 8 || D.rg < 8 || pick_R2(D.r, D.rg) == 0) return false;
  if (D.M <= 0 || D.M > (int64_t)0x7fffff00) return false;
  return device_sm_count() > 0 && wgrad_sm100_supported(D.d, D.r) && wgrad_sm100_supported(D.d, D.rg);
}

size_t fused_k1_bwd_ws(const VlpetK1Desc& D) { return carve(D.M, D.d, D.r, D.rg, true, nullptr).bytes; }

int fused_k1_bwd(const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w, void* dx1,
                 void* dx2, const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || !aligned16(w.Gd) || !aligned16(w.Gu) || !aligned16(w.bu) || !aligned16(w.gbu))
    return fail(VLPET_E_ALIGN, "k1_bwd(fused): weights must be 16-byte aligned");
  return run_bwd(true, D, x1, x2, dout, w, dx1, dx2, G, ws, ws_bytes, st);
}

// ---- K2 (decoder value parallel adapter) through the same kernels, ungated ---------------------------------------
static VlpetK1Desc k2_as_k1(const VlpetK2Desc& D) {
  VlpetK1Desc K;
  memset(&K, 0, sizeof(K));
  K.M = D.M; K.d = D.d; K.r = D.r; K.rg = 0; K.gate = VLPET_GATE_NONE; K.dtype = D.dtype; K.impl = D.impl;
  K.s = 1.0f; K.alpha = D.sf; K.kappa = 0.0f;     // out = y + 1 * (0 * kv + sf * (Up(gelu_new(Down kv)) + bu))
  return K;
}

bool fused_k2_supported(const VlpetK2Desc& D) {
  if (D.dtype != VLPET_BF16) return false;
  if (D.d % 128 != 0 || D.d < 128 || D.r % 8 != 0 || D.r < 8 || pick_R2(D.r, D.r) == 0) return false;
  if (D.M <= 0 || D.M > (int64_t)0x7fffff00) return false;
  return device_sm_count() > 0 && wgrad_sm100_supported(D.d, D.r) && fused_k1_fwd_supported(k2_as_k1(D));
}

size_t fused_k2_bwd_ws(const VlpetK2Desc& D) { return carve(D.M, D.d, D.r, D.r, false, nullptr).bytes; }

int fused_k2_fwd(const VlpetK2Desc& D, const void* kv, const void* y, const VlpetK2Params& w, void* out, cudaStream_t st) {
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  return fused_k1_fwd(k2_as_k1(D), y, kv, P, out, nullptr, 0, st);
}

int fused_k2_bwd(const VlpetK2Desc& D, const void* kv, const void* dout, const VlpetK2Params& w, void* dkv,
                 const VlpetK2Grads& g, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || !aligned16(w.bu))
    return fail(VLPET_E_ALIGN, "k2_bwd(fused): weights must be 16-byte aligned");
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  VlpetK1Grads G;
  memset(&G, 0, sizeof(G));
  G.dWd = g.dWd; G.dbd = g.dbd; G.dWu = g.dWu; G.dbu = g.dbu;
  return run_bwd(false, k2_as_k1(D), nullptr, kv, dout, P, nullptr, dkv, G, ws, ws_bytes, st);
}

}  // namespace vlpet
