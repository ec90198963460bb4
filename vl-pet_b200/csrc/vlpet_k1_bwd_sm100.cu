// K1 backward, fused, sm_100a ("B1"): activation gradients of the granularity-controlled PET module, large gate, in ONE
// launch; the four token-contracted weight-gradient GEMMs run right after it in vlpet_wgrad_sm100.cu ("B2") from the
// intermediates this kernel leaves behind.  Math: SURVEY Appendix A / oracle gated_pet_bwd
// (my_transformers/modeling_bart.py:1145-1155, 1195-1209, 1256-1260 differentiated).
//
// Persistent CTAs, one per SM, each walks 128-token tiles.  Per tile (nothing is saved by the forward: everything is
// recomputed from x1 / x2):
//   phase 1  (K = d)        A = x2 Wd^T, P = x1 Gd^T                                  tcgen05 SS, fp32 accum in TMEM
//   epi 1                   z = gelu_new(A+bd), q = gelu_new(P+gbd)  -> packed bf16 IN PLACE over A / P in TMEM (the A operand
//                           of every later MMA that contracts over r: tcgen05.mma TS form) + scratch for B2;
//                           gelu_new'(.) -> 16-bit fixed point next to it (8 + 8 columns per 16 pre-activations)
//   phase 2  (64-col chunks) U_c = z Wu_c^T, T_c = q Gu_c^T                           tcgen05 TS
//   epi 2                   y1 = k x2 + a(U+bu), G = sig(T+gbu), dh = s m dout,
//                           du = a dh G, dt = dh y1 G(1-G)   (add-gate: du = a dh, dt = dh G(1-G))
//                                                                                     -> in place over x2_c / dout_c in smem
//            MMA            dz += du_c Wu_c, dq += dt_c Gu_c   (SS; B operands MN-major from the SAME smem tiles of Wu_c/Gu_c)
//            store          du_c, dt_c -> scratch (TMA store) for the weight-gradient GEMMs
//   epi 3                   da = dz gelu_new'(A+bd), dp = dq gelu_new'(P+gbd)        -> packed bf16 over the gelu' slots in TMEM
//                                                                                        + scratch
//   phase 3  (64-col chunks) T_c again (only G is needed; recompute beats keeping [128 x 768] gates),
//                           DX2_c = da Wd[:,c], DX1_c = dp Gd[:,c]                    tcgen05 TS (B MN-major from Wd_c/Gd_c tiles)
//   epi 4                   dx2 = k dh G + DX2, dx1 = dout + DX1                      -> smem -> TMA store
// Round-2 changes against the first version: z / q / da / dp never touch shared memory (80 KB and ~1.5 MB
// of shared-memory port traffic per tile freed: the kernel is port-bound), which pays for 3-stage activation and weight
// rings; role loops run under elect.sync (uniform-register issue path); and the barrier-phase hazard behind the sporadic
// launch failure of round 1 is closed (see epilogue 4).
// Tile split (small M, the strong-scaling case: fewer tiles than SMs): a thread-block cluster of nsplit CTAs shares a tile.
// Every CTA runs phase 1 and epilogue 1 in full (redundant, L2 hits), but only nkc/nsplit of the column chunks of phases 2
// and 3; the partial dz / dq sums meet in an L2-resident exchange buffer between phase 2 and epilogue 3 (plain stores, one
// cluster-scope mbarrier round, summed in rank order: deterministic).  One tile then costs phase 1 + 1/nsplit of the rest.
// Warp roles: 0 TMA producer (activations), 3 TMA producer (weights), 1 MMA issuer (TMEM owner), 2 TMA-store issuer,
// 4..19 epilogue (warp%4 = TMEM lane quarter; cg = (warp-4)/4: adapter | gate branch and column half in epi 1/3, 16 of the
// chunk's 64 columns in epi 2/4).  Epilogue arithmetic is on packed fp32 pairs (FFMA2).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {
int make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                  uint32_t box_cols, bool weight);
namespace {

constexpr int TILE_M = 128;
constexpr int CH = 64;
constexpr int SX = 3;              // activation ring stages (2 x [128 x 64] bf16 each)
constexpr int SW = 3;              // weight ring stages
constexpr int XCH_BYTES = TILE_M * CH * 2;  // 16 KB
constexpr int NUM_THREADS = 640;   // 4 role warps + 16 epilogue warps
constexpr int EPI_THREADS = 512;
constexpr int SMEM_LIMIT = 232448;

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
  static constexpr int OFF_BD = OFF_W + SW * WSLOT;          // fp32 bd[R] | gbd[R]
  static constexpr int OFF_DSUM = OFF_BD + 2 * R * 4;        // fp32 column sums of da | dp over this CTA's tiles (dbd, dgbd)
  static constexpr int OFF_BAR = OFF_DSUM + 2 * R * 4;
  static constexpr int OFF_TAB = OFF_BAR + 512;              // fp32 alpha*bu[d] | 0.5*gbu[d]
  static constexpr int smem_bytes(int d) { return OFF_TAB + 2 * d * 4 + 1024; }   // + slack for the manual 1024-B alignment
  // optional: the dropout bits of a tile row, u16 [d / 64][512 epilogue threads], filled while phase 1 streams (see "dropout bits")
  static constexpr int mask_bytes(int d) { return (d / CH) * EPI_THREADS * 2; }
  // TMEM columns.  Phases 1-2: A | P (in place after epilogue 1: per 16 columns, 8 of packed z/q then 8 of gelu' fixed point,
  // later packed da/dp) | dz | dq | U_c T_c.  Phase 3 reuses dz.. for {T_c, DX2_c, DX1_c}.
  static constexpr int TM_A = 0, TM_P = R, TM_DZ = 2 * R, TM_DQ = 3 * R, TM_UT = 4 * R;
  static constexpr int TM_T = 2 * R, TM_DX2 = 2 * R + CH, TM_DX1 = 2 * R + 2 * CH;
  static_assert(R <= 96 && R % 32 == 0, "fused backward covers rank buckets 32 / 64 / 96");
  static_assert(TM_UT + 2 * CH <= 512 && TM_DX1 + CH <= 512, "TMEM budget");
  static_assert(WSLOT % 1024 == 0 && WA_BYTES % 1024 == 0, "swizzle atoms must stay 1024-byte aligned");
};

// Optional phase-timestamp trace (tools/trace_k1.py --bwd): thread 128 of every CTA stamps %globaltimer at the phase
// boundaries of its first tile into [cta][128] slots.
static unsigned long long* g_trace_b = nullptr;   // host copy; travels to the kernel as BParams::trace
static int g_bwd_parts = []() { const char* e = getenv("VLPET_DEBUG_BWD_PARTS"); return e ? atoi(e) : 7; }();
__device__ __forceinline__ unsigned long long gtimer_b() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef VLPET_TRACE   // developer builds only (tools/_variant.py): the stamps cost the issue-bound epilogues ~30 instructions per chunk
#define VLPET_TRACE_B(slot)                                                                                               \
  do {                                                                                                                    \
    if (p.trace && threadIdx.x == 128 && ti == 0 && (slot) < 128) p.trace[blockIdx.x * 128 + (slot)] = gtimer_b();       \
  } while (0)
#else
#define VLPET_TRACE_B(slot) do { } while (0)
#endif

struct BParams {
  int64_t M;
  int d, r, rg;
  int add_gate;
  float s, alpha, kappa;
  const __nv_bfloat16 *bd, *bu, *gbd, *gbu;
  __nv_bfloat16 *zs, *qs, *das, *dps;   // scratch [M, pz] / [M, pq]
  float *dbd, *dgbd;                    // fp32 bias gradients of the down projections (accumulated into) or nullptr
  int pz, pq;                           // scratch row pitches (elements)
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t thr16;
  float inv_keep;
  int premask;                          // the shared-memory table of dropout bits exists (it fits for d <= 768)
  int64_t tile_begin, tile_end;         // tiles [tile_begin, tile_end) belong to this launch
  int nsplit;                           // > 1: a cluster of nsplit CTAs shares every tile (small M), see "tile split" below
  float* xchg;                          // [tiles][nsplit][2R][128] fp32: partial dz | dq of the CTAs of a cluster
  unsigned long long* trace;            // developer hook (tools/trace_k1_bwd.py), normally null
  uint32_t* dbg;                        // host-mapped trap record (ptx::mbar_wait_dbg)
};

enum { B_XFULL = 0, B_XEMPTY = B_XFULL + SX, B_WFULL = B_XEMPTY + SX, B_WEMPTY = B_WFULL + SW, B_APFULL = B_WEMPTY + SW,
       B_ZQFULL, B_UTFULL, B_UTEMPTY, B_DUDT, B_DZQDONE = B_DUDT + SX, B_DZFULL = B_DZQDONE + SX, B_DAPFULL,
       B_ACCFULL, B_ACCEMPTY, B_XCHG, B_OUTRDY, B_COUNT = B_OUTRDY + SX };
static_assert(8 * B_COUNT + 8 <= 512, "barriers + the TMEM address slot must fit the 512-byte barrier block");

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
// gelu_new' lies in [-0.13, 1.13]; it waits in TMEM between epilogues 1 and 3 as 16-bit fixed point, two per column:
// u = round(40000 g + 10000), taken from the mantissa of (2^23 + u).  Step 2.5e-5 absolute -- a bf16 pair would cost
// 2^-9 relative on every da / dp, on top of their own rounding.
__device__ __forceinline__ uint32_t enc_fix2(f2 g) {
  const f2 t = fma2(g, mk2(40000.0f, 40000.0f), mk2(8398608.0f, 8398608.0f));
  uint32_t lo, hi;
  un2u(t, lo, hi);
  return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ f2 dec_fix2(uint32_t v) {
  const f2 f = mk2u(__byte_perm(v, 0x4B000000u, 0x7610), __byte_perm(v, 0x4B000000u, 0x7632));
  return fma2(add2(f, mk2(-8388608.0f, -8388608.0f)), mk2(2.5e-5f, 2.5e-5f), mk2(-0.25f, -0.25f));
}
// Column sums over the 32 lanes of a warp (= 32 token rows), 16 columns: every lane enters with its row's 16 values; lanes l and
// l ^ 16 leave with the sum over rows of column (l & 15).  31 shuffles instead of 80 for 16 butterflies.
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
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// GATED = false is the ungated form used for the decoder value parallel adapter (K2): out = x1 + alpha*(Up(gelu_new(Down x2)))
// -> dx2 = (alpha * (dout Wu) * gelu_new'(A)) Wd, no gate branch, no U/T recompute, no du/dt scratch (du = alpha*dout).
// MUL = the gate multiplies (default) / false = the add-gate ablation flag (a compile-time switch: as a run-time one it
// cost ~40 of the ~530 instructions epilogue 2 issues per thread and chunk, and the epilogues are issue-bound).
// SPLIT = the tile-split form (a cluster shares a tile): a compile-time switch as well -- as a run-time one the exchange
// code sat in epilogue 3 as ~100 predicated-off instructions per 16 columns.
template <int R, bool GATED, bool MUL, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                    const __grid_constant__ CUtensorMap tm_dout, const __grid_constant__ CUtensorMap tm_dx1,
                    const __grid_constant__ CUtensorMap tm_dx2, const __grid_constant__ CUtensorMap tm_du,
                    const __grid_constant__ CUtensorMap tm_dt, const __grid_constant__ CUtensorMap tm_wd,
                    const __grid_constant__ CUtensorMap tm_gd, const __grid_constant__ CUtensorMap tm_wu,
                    const __grid_constant__ CUtensorMap tm_gu, const __grid_constant__ CUtensorMap tm_kap, const BParams p) {
  using C = BCfg<R>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int nkc = p.d / CH;
  const int64_t num_tiles = p.tile_end;
  constexpr bool mulgate = MUL;
  const int nsplit = SPLIT ? p.nsplit : 1;
  const uint32_t crank = SPLIT ? ptx::cluster_ctarank() : 0u;
  const int cps = nkc / nsplit;                       // chunks of phases 2 / 3 this CTA owns: [cb, ce)
  const int cb = (int)crank * cps, ce = cb + cps;
  const int64_t tile0 = p.tile_begin + blockIdx.x / nsplit, tstride = gridDim.x / nsplit;

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
    ptx::mbar_init(bar(B_ACCFULL), 1);
    ptx::mbar_init(bar(B_ACCEMPTY), EPI_THREADS);
    ptx::mbar_init(bar(B_XCHG), (uint32_t)nsplit);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_dout); ptx::prefetch_tmap(&tm_kap); }
  if (warp == 3 && lane == 0) {
    ptx::prefetch_tmap(&tm_wd); ptx::prefetch_tmap(&tm_gd); ptx::prefetch_tmap(&tm_wu); ptx::prefetch_tmap(&tm_gu);
  }
  if (warp == 2 && lane == 0) {
    ptx::prefetch_tmap(&tm_dx1); ptx::prefetch_tmap(&tm_dx2); ptx::prefetch_tmap(&tm_du); ptx::prefetch_tmap(&tm_dt);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  {  // biases -> fp32 in shared memory, pre-multiplied so that the epilogues fold them into FMAs they issue anyway
    float* sbd = reinterpret_cast<float*>(smem_gen + C::OFF_BD);
    for (int i = threadIdx.x; i < 2 * R; i += NUM_THREADS) {
      const int br = i / R, j = i % R;
      const int rr = br ? p.rg : p.r;
      const __nv_bfloat16* src = br ? p.gbd : p.bd;
      sbd[i] = (j < rr && (GATED || br == 0)) ? __bfloat162float(src[j]) : 0.f;
    }
    float* sds = reinterpret_cast<float*>(smem_gen + C::OFF_DSUM);
    for (int i = threadIdx.x; i < 2 * R; i += NUM_THREADS) sds[i] = 0.f;
    float* stab = reinterpret_cast<float*>(smem_gen + C::OFF_TAB);
    if (GATED) {
      for (int i = threadIdx.x; i < p.d; i += NUM_THREADS) {
        stab[i] = p.alpha * __bfloat162float(p.bu[i]);
        stab[p.d + i] = 0.5f * __bfloat162float(p.gbu[i]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (SPLIT) ptx::cluster_sync_all();   // the peers' exchange barriers are initialised before any remote arrive
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================== TMA producer: activations =====================================
    // x1 / x2 are read again by the weight-gradient GEMM right after this kernel, dout again in phase 3: default L2 policy,
    // except the last read of dout.
    if (ptx::elect_one()) {
      uint32_t xi = 0;
      for (int64_t tile = tile0; tile < num_tiles; tile += tstride) {
        const int row0 = (int)(tile * TILE_M);
        for (int ph = 0; ph < 3; ++ph) {
          for (int c = ph ? cb : 0; c < (ph ? ce : nkc); ++c, ++xi) {
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
              ptx::tma_load_2d_hint(xdst, &tm_dout, c * CH, row0, bar(B_XFULL + sx), ptx::L2_EVICT_FIRST);
            } else if (p.kappa != 0.f) {             // ungated with an x2 scale: dx2 = kappa K + da Wd; K = dout again, or
              ptx::mbar_arrive_expect_tx(bar(B_XFULL + sx), XCH_BYTES);   // another [M, d] tensor (BwdExtras::kap_src)
              ptx::tma_load_2d_hint(xdst, &tm_kap, c * CH, row0, bar(B_XFULL + sx), ptx::L2_EVICT_FIRST);
            } else {
              ptx::mbar_arrive(bar(B_XFULL + sx));   // the stage is only the staging buffer of dx2_c
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA producer: weights (always L2 hits) =====================================
    if (ptx::elect_one()) {
      uint32_t wi = 0;
      for (int64_t tile = tile0; tile < num_tiles; tile += tstride) {
        for (int ph = 0; ph < 3; ++ph) {
          for (int c = ph ? cb : 0; c < (ph ? ce : nkc); ++c, ++wi) {
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
    if (ptx::elect_one()) {
      constexpr uint32_t IDESC_AP = ptx::umma_idesc_bf16_m128(R);                       // A/P: N = R, K-major x K-major
      constexpr uint32_t IDESC_UT = ptx::umma_idesc_bf16_m128(CH);                      // U/T: N = 64
      constexpr uint32_t IDESC_DZ = ptx::umma_idesc_bf16_m128_major(R, 0u, 1u);         // dz/dq: B MN-major, N = R
      constexpr uint32_t IDESC_DX = ptx::umma_idesc_bf16_m128_major(CH, 0u, 1u);        // DX: B MN-major, N = 64
      uint32_t xi = 0, wi = 0, ui = 0, p2i = 0, ai = 0, ti = 0;
      for (int64_t tile = tile0; tile < num_tiles; tile += tstride, ++ti) {
        // Ordering, not a TMEM hazard: the phase-3 XFULL phases of the previous tile were observed by the EPILOGUE only, and
        // it arrives on ACCEMPTY after having seen them.  A parity wait on XFULL below is meaningful only once the previous
        // phase of that barrier has completed (see epilogue 4).
        if (ai > 0) ptx::mbar_wait_dbg(bar(B_ACCEMPTY), (ai - 1) & 1, p.dbg, __LINE__);
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
            ptx::umma_bf16_ss(tmem_base + C::TM_A, ptx::umma_desc_kmajor_sw128(x2s + ks * 32),
                              ptx::umma_desc_kmajor_sw128(wds + ks * 32), IDESC_AP, acc);
            if (GATED)
              ptx::umma_bf16_ss(tmem_base + C::TM_P, ptx::umma_desc_kmajor_sw128(x1s + ks * 32),
                                ptx::umma_desc_kmajor_sw128(gds + ks * 32), IDESC_AP, acc);
          }
          ptx::umma_commit(bar(B_XEMPTY + sx));
          ptx::umma_commit(bar(B_WEMPTY + sw));
        }
        ptx::umma_commit(bar(B_APFULL));
        // ---- phase 2
        if (!GATED) {
          // dz += dout_c Wu_c straight from the ring.  The accumulator columns are free: epilogue 3 of the previous tile
          // (their last reader) precedes every phase-3 accumulator hand-over this thread has already waited for.
          for (int c = cb; c < ce; ++c, ++xi, ++wi) {
            const uint32_t sx = xi % SX, sw = wi % SW;
            ptx::mbar_wait_dbg(bar(B_XFULL + sx), (xi / SX) & 1, p.dbg, __LINE__);
            ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
            ptx::tc_fence_after();
            const uint32_t dos = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
            const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT;
#pragma unroll
            for (int ks = 0; ks < CH / 16; ++ks)
              ptx::umma_bf16_ss(tmem_base + C::TM_DZ, ptx::umma_desc_kmajor_sw128(dos + ks * 32),
                                ptx::umma_desc_mnmajor_sw128(wus + ks * 2048, CH * CH * 2), IDESC_DZ, (c > cb || ks > 0) ? 1u : 0u);
            ptx::umma_commit(bar(B_XEMPTY + sx));
            ptx::umma_commit(bar(B_WEMPTY + sw));
          }
          ptx::umma_commit(bar(B_DZFULL));
        } else {
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
              ptx::umma_bf16_ss(tmem_base + C::TM_DZ, ptx::umma_desc_kmajor_sw128(dus + ks * 32),
                                ptx::umma_desc_mnmajor_sw128(wus + ks * 2048, CH * CH * 2), IDESC_DZ, acc);
              ptx::umma_bf16_ss(tmem_base + C::TM_DQ, ptx::umma_desc_kmajor_sw128(dts + ks * 32),
                                ptx::umma_desc_mnmajor_sw128(gus + ks * 2048, CH * CH * 2), IDESC_DZ, acc);
            }
            ptx::umma_commit(bar(B_DZQDONE + k));
            ptx::umma_commit(bar(B_WEMPTY + sw));
          };
          for (int c = cb; c < ce; ++c, ++xi, ++wi, ++ui, ++p2i) {
            const uint32_t sw = wi % SW;
            ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
            ptx::mbar_wait_dbg(bar(B_UTEMPTY), (ui & 1) ^ 1, p.dbg, __LINE__);
            ptx::tc_fence_after();
            const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT, gus = wus + C::WB_BYTES;
#pragma unroll
            for (int ks = 0; ks < R / 16; ++ks) {   // z / q: A operand from TMEM, 8 packed columns per K step at stride 16
              const uint32_t kb = ks / 4, kin = ks % 4;
              ptx::umma_bf16_ts(tmem_base + C::TM_UT, tmem_base + C::TM_A + 16 * ks,
                                ptx::umma_desc_kmajor_sw128(wus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
              ptx::umma_bf16_ts(tmem_base + C::TM_UT + CH, tmem_base + C::TM_P + 16 * ks,
                                ptx::umma_desc_kmajor_sw128(gus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
            }
            ptx::umma_commit(bar(B_UTFULL));
            if (c > cb) issue_dzdq(c - 1 - cb);
          }
          issue_dzdq(cps - 1);
          ptx::umma_commit(bar(B_DZFULL));
        }
        // ---- phase 3
        ptx::mbar_wait_dbg(bar(B_DAPFULL), ti & 1, p.dbg, __LINE__);
        ptx::tc_fence_after();
        for (int c = cb; c < ce; ++c, ++xi, ++wi, ++ai) {
          const uint32_t sw = wi % SW;
          ptx::mbar_wait_dbg(bar(B_WFULL + sw), (wi / SW) & 1, p.dbg, __LINE__);
          ptx::mbar_wait_dbg(bar(B_ACCEMPTY), (ai & 1) ^ 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t gus = smem_base + C::OFF_W + sw * C::WSLOT, wds = gus + C::WB_BYTES, gds = wds + C::WA_BYTES;
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks) {
            const uint32_t kb = ks / 4, kin = ks % 4;
            if (GATED && mulgate)
              ptx::umma_bf16_ts(tmem_base + C::TM_T, tmem_base + C::TM_P + 16 * ks,
                                ptx::umma_desc_kmajor_sw128(gus + kb * (CH * CH * 2) + kin * 32), IDESC_UT, ks > 0);
            ptx::umma_bf16_ts(tmem_base + C::TM_DX2, tmem_base + C::TM_A + 16 * ks + 8,
                              ptx::umma_desc_mnmajor_sw128(wds + ks * 2048, C::WA_BYTES), IDESC_DX, ks > 0);
            if (GATED)
              ptx::umma_bf16_ts(tmem_base + C::TM_DX1, tmem_base + C::TM_P + 16 * ks + 8,
                                ptx::umma_desc_mnmajor_sw128(gds + ks * 2048, C::WA_BYTES), IDESC_DX, ks > 0);
          }
          ptx::umma_commit(bar(B_WEMPTY + sw));
          ptx::umma_commit(bar(B_ACCFULL));
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== TMA store issuer =====================================
    if (ptx::elect_one()) {
      uint32_t xi = 0, p2i = 0, p3i = 0;
      for (int64_t tile = tile0; tile < num_tiles; tile += tstride) {
        const int row0 = (int)(tile * TILE_M);
        xi += nkc;  // phase-1 stages are released by the MMA warp
        if (!GATED) xi += cps;  // ungated: phase-2 stages are released by the MMA warp as well
        for (int c = cb; GATED && c < ce; ++c, ++xi, ++p2i) {
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
        for (int c = cb; c < ce; ++c, ++xi, ++p3i) {
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
    const uint32_t sb_base = smem_base + C::OFF_BD;     // fp32 [2][R]
    const uint32_t tab_base = smem_base + C::OFF_TAB;   // fp32 alpha*bu[d] | 0.5*gbu[d]
    const f2 half2 = mk2(0.5f, 0.5f), kappa2 = mk2(p.kappa, p.kappa), alpha2 = mk2(p.alpha, p.alpha);
    const f2 halpha2 = mk2(0.5f * p.alpha, 0.5f * p.alpha);
    const f2 s2 = mk2(p.s, p.s);
    const float s_keep = p.s * p.inv_keep;
    uint32_t ui = 0, ai = 0, ti = 0;
    // dbd / dgbd: this warp's running fp32 column sums of da / dp over all tiles of the CTA (lane l < 16 <-> column
    // jbeg + 16 k + l), reduced over the four lane quarters once, after the tile loop.  (A shared-memory float atomicAdd per
    // tile is a CAS spin loop under 4-way contention: it cost epilogue 3 ~5 us per tile.)
    float dsum_acc[HALF / 16];
#pragma unroll
    for (int k = 0; k < HALF / 16; ++k) dsum_acc[k] = 0.f;
    // Ring positions are advanced incrementally: with 3-stage rings, index % 3 and index / 3 in every chunk cost the
    // (issue-bound) epilogues a multiply-high sequence each.  xs / xph = stage and parity of the activation-ring entry
    // this role looks at next; k2 / k3 = stage of the DUDT / OUTRDY barrier rings.
    uint32_t xs = 0, xph = 0, k2 = 0, k3 = 0;
    auto x_adv = [&]() { if (++xs == SX) { xs = 0; xph ^= 1u; } };
    auto x_skip = [&](uint32_t n) { const uint32_t t = xs + n; xph ^= (t / SX) & 1u; xs = t % SX; };
    // dropout mode of epilogues 2 / 4: 0 none, 1 hash in the loop, 2 bits from the shared-memory table
    const int drop_mode = (GATED && p.thr16) ? (p.premask ? 2 : 1) : 0;
    const uint32_t mask_base = tab_base + 8u * (uint32_t)p.d + 2u * (uint32_t)(threadIdx.x - 128);   // u16 [chunk][thread]
    for (int64_t tile = tile0; tile < num_tiles; tile += tstride, ++ti) {
      const int64_t grow = tile * TILE_M + row;
      const bool row_ok = grow < p.M;
      // ---- dropout bits: the 16 keep bits of (this row, this thread's 16 columns) for every chunk this CTA owns, formed
      //      here -- while phase 1 streams x1 / x2 and the epilogue warps have nothing to do -- instead of hashing inside
      //      epilogues 2 and 4, which are issue-bound (the hashes were 136 of ~530 instructions per chunk there).
      //      Bit 4h + k = element k of hash h, the order the in-loop hashing uses.  Written and read by the same thread.
      if (drop_mode == 2) {
        for (int c = cb; c < ce; ++c) {
          const uint64_t i4 = (uint64_t)(grow * p.d + c * CH + cg * 16) >> 2;
          uint32_t bits = 0;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const uint64_t hv = drop_hash4(seed_eff, i4 + h);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              bits |= (((uint32_t)(hv >> (16 * k)) & 0xffffu) >= p.thr16) ? (1u << (4 * h + k)) : 0u;
          }
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(mask_base + (uint32_t)(c - cb) * (EPI_THREADS * 2)), "h"((uint16_t)bits) : "memory");
        }
      }
      // ---- epilogue 1 (APFULL also implies that every MMA of the previous tile, which read q / da / dp, has completed)
      VLPET_TRACE_B(0);
      ptx::mbar_wait_dbg(bar(B_APFULL), ti & 1, p.dbg, __LINE__);
      VLPET_TRACE_B(1);
      ptx::tc_fence_after();
      if (GATED || branch == 0) {
        const uint32_t tsrc = lane_addr + (branch ? C::TM_P : C::TM_A);
        const int rr = branch ? p.rg : p.r;
        __nv_bfloat16* srow = (branch ? p.qs + grow * p.pq : p.zs + grow * p.pz);
#pragma unroll
        for (int jj = 0; jj < HALF; jj += 16) {
          const int j0 = jbeg + jj;
          uint32_t v[16], o[16];
          ptx::tmem_ld_32x32b_x16(tsrc + j0, v);
          f2 bias[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) lds_f2x2(sb_base + 4u * (uint32_t)(branch * R + j0 + e * 4), bias[2 * e], bias[2 * e + 1]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f2 z2, d2;
            gelu_new_both2(add2(mk2u(v[e * 2], v[e * 2 + 1]), bias[e]), z2, d2);
            o[e] = pack2(z2);
            o[8 + e] = enc_fix2(d2);
          }
          // in place: columns j0 .. j0+7 <- packed z (K step j0/16 of the TS-form A operand), j0+8 .. j0+15 <- gelu'
          ptx::tmem_st_32x32b_x16(tsrc + j0, o);
          if (row_ok && crank == 0) {
            if (j0 < rr) *reinterpret_cast<uint4*>(srow + j0) = make_uint4(o[0], o[1], o[2], o[3]);
            if (j0 + 8 < rr) *reinterpret_cast<uint4*>(srow + j0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        ptx::tmem_st_wait();
        if (row_ok && crank == 0 && (cg & 1) == 1) {  // ones column (bias-gradient trick of the weight-gradient GEMM) + zero pad up to the pitch
          const uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(srow + rr) = one;
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(B_ZQFULL));
      VLPET_TRACE_B(2);
      x_skip((uint32_t)nkc + (GATED ? 0u : (uint32_t)cps));
      // ---- epilogue 2, per 64-column chunk: du, dt
      for (int c = cb; GATED && c < ce; ++c, ++ui) {
        const uint32_t sx = xs;
        ptx::mbar_wait_dbg(bar(B_UTFULL), ui & 1, p.dbg, __LINE__);
        VLPET_TRACE_B(3 + 3 * (c - cb));
        ptx::tc_fence_after();
        uint32_t u[16], t[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + C::TM_UT + cg * 16, u);
        ptx::tmem_ld_32x32b_x16(lane_addr + C::TM_UT + CH + cg * 16, t);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(B_UTEMPTY));
        ptx::mbar_wait_dbg(bar(B_XFULL + sx), xph, p.dbg, __LINE__);
        VLPET_TRACE_B(4 + 3 * (c - cb));
        const uint32_t x2row = smem_base + C::OFF_X + sx * (2 * XCH_BYTES) + (uint32_t)row * 128u;
        const uint32_t dorow = x2row + XCH_BYTES;
        const int col0 = c * CH + cg * 16;
        const int64_t idx0 = grow * p.d + col0;
        const uint32_t tabu = tab_base + 4u * (uint32_t)col0, thgb = tabu + 4u * (uint32_t)p.d;
        uint32_t mbits = 0;
        if (drop_mode == 2) {
          uint16_t mb;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(mb) : "r"(mask_base + (uint32_t)(c - cb) * (EPI_THREADS * 2)));
          mbits = mb;
        }
        auto group2 = [&](auto drop_tag, int g) {
          constexpr int DROP = decltype(drop_tag)::value;
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
          if (DROP == 1) {
            h0 = drop_hash4(seed_eff, (uint64_t)(idx0 + g * 8) >> 2);
            h1 = drop_hash4(seed_eff, ((uint64_t)(idx0 + g * 8) >> 2) + 1);
          }
          const f2 q2 = mk2(0.25f, 0.25f), nq2 = mk2(-0.25f, -0.25f);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            f2 sce = s2;                                                          // s * dropout mask of this pair
            if (DROP == 1) {
              const uint32_t two = (uint32_t)((e >> 1 ? h1 : h0) >> (32 * (e & 1)));
              sce = mk2(((two & 0xffffu) >= p.thr16) ? s_keep : 0.f, ((two >> 16) >= p.thr16) ? s_keep : 0.f);
            } else if (DROP == 2) {
              sce = mk2((mbits & (1u << j)) ? s_keep : 0.f, (mbits & (2u << j)) ? s_keep : 0.f);
            }
            const f2 dh = mul2(sce, bf2_to_f2(dv[e]));                            // dh = s m dout
            const f2 th = tanh2(fma2(half2, mk2u(t[j], t[j + 1]), hgb[e]));      // G = 0.5 + 0.5 th
            const f2 gg = fma2(nq2, mul2(th, th), q2);                            // G (1 - G) = 0.25 (1 - th^2)
            f2 du, dt;
            if (mulgate) {
              f2 y1 = fma2(kappa2, bf2_to_f2(xv[e]), abu[e]);                     // kappa x2 + alpha bu
              y1 = fma2(alpha2, mk2u(u[j], u[j + 1]), y1);                        // + alpha U
              const f2 h5 = mul2(halpha2, dh);
              du = fma2(h5, th, h5);                                              // alpha dh G = (0.5 alpha dh)(1 + th)
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
        using Tag0 = std::integral_constant<int, 0>; using Tag1 = std::integral_constant<int, 1>; using Tag2 = std::integral_constant<int, 2>;
        if (drop_mode == 2) { group2(Tag2{}, 0); group2(Tag2{}, 1); }
        else if (drop_mode == 1) { group2(Tag1{}, 0); group2(Tag1{}, 1); }
        else { group2(Tag0{}, 0); group2(Tag0{}, 1); }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(bar(B_DUDT + k2));
        if (++k2 == SX) k2 = 0;
        x_adv();
        VLPET_TRACE_B(5 + 3 * (c - cb));
      }
      // ---- epilogue 3: da = dz * gelu_new'(A + bd) (branch 0), dp = dq * gelu_new'(P + gbd) (branch 1)
      VLPET_TRACE_B(40);
      ptx::mbar_wait_dbg(bar(B_DZFULL), ti & 1, p.dbg, __LINE__);
      VLPET_TRACE_B(41);
      ptx::tc_fence_after();
      if (SPLIT) {
        // ---- tile split: publish this CTA's partial dz | dq, meet the other CTAs of the cluster
        if (GATED || branch == 0) {
          const uint32_t tdz = lane_addr + (branch ? C::TM_DQ : C::TM_DZ);
          // [tile][rank][column][row]: the 32 lanes of a warp (32 rows) write one 128-byte line per column -- a row-major
          // slot made every warp instruction touch 32 lines (~40 us per exchange in the L1 wavefront queue)
          float* mine = p.xchg + (((size_t)(tile - p.tile_begin) * p.nsplit + crank) * (2 * R) + branch * R + jbeg) * TILE_M + row;
#pragma unroll
          for (int jj = 0; jj < HALF; jj += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(tdz + jbeg + jj, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) __stcg(mine + (size_t)(jj + e) * TILE_M, __uint_as_float(v[e]));
          }
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // epilogue warps only
        if (threadIdx.x == 128) {
          for (uint32_t pr = 0; pr < (uint32_t)p.nsplit; ++pr) ptx::mbar_arrive_cluster_release(ptx::mapa(bar(B_XCHG), pr));
        }
        ptx::mbar_wait_cluster_dbg(bar(B_XCHG), ti & 1, p.dbg, __LINE__);
        __threadfence();
      }
      if (GATED || branch == 0) {
        const f2 dzs2 = GATED ? mk2(1.0f, 1.0f) : alpha2;   // ungated: du = alpha*dout was fed unscaled
        const uint32_t tpre = lane_addr + (branch ? C::TM_P : C::TM_A);
        const uint32_t tdz = lane_addr + (branch ? C::TM_DQ : C::TM_DZ);
        const int rr = branch ? p.rg : p.r;
        __nv_bfloat16* srow = (branch ? p.dps + grow * p.pq : p.das + grow * p.pz);
#pragma unroll
        for (int jj = 0; jj < HALF; jj += 16) {
          const int j0 = jbeg + jj;
          uint32_t gq[8], dz[16], o[8];
          ptx::tmem_ld_32x32b_x8(tpre + j0 + 8, gq);   // gelu_new'(pre-activation), stored by epilogue 1
          ptx::tmem_ld_32x32b_x16(tdz + j0, dz);
          ptx::tmem_ld_wait();
          if (SPLIT) {   // sum of the partials in rank order (every CTA of the cluster gets the same bits)
            // two peers' partials are requested per round (two L2 round trips per 16 columns instead of one per peer; four
            // at once spilled); the CTA's own slot is read back like the others, so the code is the same for every rank and
            // the sum is formed in rank order ((p0 + p1) + p2) + p3 on every CTA
            float acc[16];
#pragma unroll
            for (uint32_t pr = 0; pr < 4; pr += 2) {
              float pv[2][16];
#pragma unroll
              for (uint32_t q = 0; q < 2; ++q) {
                const float* peer = p.xchg + (((size_t)(tile - p.tile_begin) * p.nsplit + pr + q) * (2 * R) + branch * R + j0) * TILE_M + row;
#pragma unroll
                for (int e = 0; e < 16; ++e) pv[q][e] = (pr + q) < (uint32_t)p.nsplit ? __ldcg(peer + (size_t)e * TILE_M) : 0.f;
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) acc[e] = pr == 0 ? pv[0][e] + pv[1][e] : (acc[e] + pv[0][e]) + pv[1][e];
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) dz[e] = __float_as_uint(acc[e]);
          }
          float cs[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const f2 da2 = mul2(mul2(mk2u(dz[2 * e], dz[2 * e + 1]), dec_fix2(gq[e])), dzs2);
            o[e] = pack2(da2);
            un2(da2, cs[2 * e], cs[2 * e + 1]);
          }
          ptx::tmem_st_32x32b_x8(tpre + j0 + 8, o);     // packed da / dp: K step j0/16 of the phase-3 A operands
          // dbd / dgbd: fp32 column sums of da / dp (rows beyond M contribute exact zeros: their dout is zero-filled)
          dsum_acc[jj / 16] += warp_colsum16(cs, lane);
          if (row_ok && crank == 0) {
            if (j0 < rr) *reinterpret_cast<uint4*>(srow + j0) = make_uint4(o[0], o[1], o[2], o[3]);
            if (j0 + 8 < rr) *reinterpret_cast<uint4*>(srow + j0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(B_DAPFULL));
      VLPET_TRACE_B(42);
      // ---- epilogue 4, per 64-column chunk: dx1, dx2
      for (int c = cb; c < ce; ++c, ++ai) {
        const uint32_t sx = xs;
        ptx::mbar_wait_dbg(bar(B_ACCFULL), ai & 1, p.dbg, __LINE__);
        VLPET_TRACE_B(43 + 3 * (c - cb));
        ptx::tc_fence_after();
        uint32_t t[16], g2[16], g1[16];
        if (GATED && mulgate) ptx::tmem_ld_32x32b_x16(lane_addr + C::TM_T + cg * 16, t);
        ptx::tmem_ld_32x32b_x16(lane_addr + C::TM_DX2 + cg * 16, g2);
        if (GATED) ptx::tmem_ld_32x32b_x16(lane_addr + C::TM_DX1 + cg * 16, g1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        // The XFULL wait comes BEFORE the ACCEMPTY arrive.  The MMA warp waits XFULL itself for the phase-1 entries of the
        // next tile, and a parity wait is only meaningful once the previous phase of that barrier has completed.  With the
        // arrive first (round 1), the MMA warp could reach the next tile while this chunk's dout load was still in flight,
        // take "the phase before the previous one is complete" for "complete", consume a stale slot and release it a second
        // time: the sporadic launch failure of round 1 (profiles/r2_b1_fault_rootcause.md).
        ptx::mbar_wait_dbg(bar(B_XFULL + sx), xph, p.dbg, __LINE__);
        ptx::mbar_arrive(bar(B_ACCEMPTY));
        VLPET_TRACE_B(44 + 3 * (c - cb));
        const uint32_t dorow = smem_base + C::OFF_X + sx * (2 * XCH_BYTES) + (uint32_t)row * 128u;
        const uint32_t o2row = dorow + XCH_BYTES;
        const int col0 = c * CH + cg * 16;
        const int64_t idx0 = grow * p.d + col0;
        const uint32_t thgb = tab_base + 4u * (uint32_t)(p.d + col0);
        uint32_t mbits = 0;
        if (drop_mode == 2) {
          uint16_t mb;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(mb) : "r"(mask_base + (uint32_t)(c - cb) * (EPI_THREADS * 2)));
          mbits = mb;
        }
        auto group4 = [&](auto drop_tag, int g) {
          constexpr int DROP = decltype(drop_tag)::value;
          const uint32_t off = (((uint32_t)(cg * 2 + g)) ^ swz) << 4;
          uint32_t dv[4] = {0, 0, 0, 0}, o1[4], o2[4];
          const bool ungated_kappa = !GATED && p.kappa != 0.f;
          if (GATED || ungated_kappa) lds128(dorow + off, dv);
          f2 hgb[4] = {0, 0, 0, 0};
          if (GATED && mulgate) {
            lds_f2x2(thgb + g * 32, hgb[0], hgb[1]);
            lds_f2x2(thgb + g * 32 + 16, hgb[2], hgb[3]);
          }
          uint64_t h0 = 0, h1 = 0;
          if (DROP == 1) {
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
              if (DROP == 1) {
                const uint32_t two = (uint32_t)((e >> 1 ? h1 : h0) >> (32 * (e & 1)));
                sce = mk2(((two & 0xffffu) >= p.thr16) ? s_keep : 0.f, ((two >> 16) >= p.thr16) ? s_keep : 0.f);
              } else if (DROP == 2) {
                sce = mk2((mbits & (1u << j)) ? s_keep : 0.f, (mbits & (2u << j)) ? s_keep : 0.f);
              }
              f2 dy1 = mul2(sce, dof);                                               // dh = s m dout
              if (mulgate) {
                const f2 th = tanh2(fma2(half2, mk2u(t[j], t[j + 1]), hgb[e]));
                dy1 = mul2(dy1, fma2(half2, th, half2));                             // dh G
              }
              o2[e] = pack2(fma2(kappa2, dy1, g2p));                                 // dx2 = kappa dy1 + da Wd
              o1[e] = pack2(add2(dof, mk2u(g1[j], g1[j + 1])));                      // dx1 = dout + dp Gd
            } else {
              o2[e] = pack2(ungated_kappa ? fma2(kappa2, bf2_to_f2(dv[e]), g2p) : g2p);
              o1[e] = 0u;
            }
          }
          if (GATED) sts128(dorow + off, o1);
          sts128(o2row + off, o2);
        };
        using Tag0 = std::integral_constant<int, 0>; using Tag1 = std::integral_constant<int, 1>; using Tag2 = std::integral_constant<int, 2>;
        if (drop_mode == 2) { group4(Tag2{}, 0); group4(Tag2{}, 1); }
        else if (drop_mode == 1) { group4(Tag1{}, 0); group4(Tag1{}, 1); }
        else { group4(Tag0{}, 0); group4(Tag0{}, 1); }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(bar(B_OUTRDY + k3));
        if (++k3 == SX) k3 = 0;
        x_adv();
        VLPET_TRACE_B(45 + 3 * (c - cb));
      }
    }
    // The four lane quarters take turns adding their sums into the [2R] block (within a quarter every column has exactly one
    // owner lane): four named barriers among the epilogue warps, once per launch, no atomics.
    for (int qq = 0; qq < 4; ++qq) {
      if (quarter == qq && (GATED || branch == 0) && lane < 16) {
        float* s1 = reinterpret_cast<float*>(smem_gen + C::OFF_DSUM) + branch * R + jbeg + lane;
#pragma unroll
        for (int k = 0; k < HALF / 16; ++k) s1[16 * k] += dsum_acc[k];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  {  // this CTA's share of the down-projection bias gradients
    const float* sds = reinterpret_cast<const float*>(smem_gen + C::OFF_DSUM);
    for (int i = threadIdx.x; i < 2 * R; i += NUM_THREADS) {
      const int br = i / R, j = i % R;
      float* dst = br ? p.dgbd : p.dbd;
      if (dst && crank == 0 && j < (br ? p.rg : p.r) && sds[i] != 0.f) atomicAdd(dst + j, sds[i]);
    }
  }
  if (SPLIT) ptx::cluster_sync_all();   // no CTA exits while a peer can still arrive on its exchange barrier
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// dbd / dgbd = column sums of the da / dp scratch ([M, pitch] bf16, first ncols columns), both in ONE launch.  Kept out of
// the tile kernel: a warp-shuffle transpose-reduce there cost ~5 us per tile on its critical path (tools/trace_k1_bwd.py).
// The scratch is read as a flat stream of 16-byte groups: a thread keeps its group-of-8-columns (pitch/8 groups per row
// divide the thread stride), so every load is a coalesced uint4 and the reduction over rows happens in registers.
struct ColsumArgs {
  const __nv_bfloat16* A[2];
  float* out[2];
  int pitch[2], ncols[2];
  int64_t M;
};
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS) colsum_scratch_kernel(const ColsumArgs a) {
  const int w = blockIdx.y;
  const __nv_bfloat16* A = a.A[w];
  float* out = a.out[w];
  if (!A || !out) return;
  const int gpr = a.pitch[w] / 8;                   // 16-byte groups per row
  const int rows_per_pass = CS_THREADS / gpr;       // rows one pass of the block covers
  const int t = threadIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  __shared__ float red[256];                        // [column] partial sums (pitch <= 256)
  for (int i = t; i < 256; i += CS_THREADS) red[i] = 0.f;
  __syncthreads();
  if (t < rows_per_pass * gpr) {
    const int g = t % gpr, r0 = t / gpr;
    const uint4* base = reinterpret_cast<const uint4*>(A);
    for (int64_t row = (int64_t)blockIdx.x * rows_per_pass + r0; row < a.M; row += (int64_t)gridDim.x * rows_per_pass) {
      const uint4 v = __ldg(base + row * gpr + g);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(u[e] << 16);
        acc[2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (g * 8 + e < a.ncols[w]) atomicAdd(&red[g * 8 + e], acc[e]);
  }
  __syncthreads();
  for (int c = t; c < a.ncols[w]; c += CS_THREADS) atomicAdd(out + c, red[c]);
}
int launch_colsum_scratch(const __nv_bfloat16* A0, int pitch0, int ncols0, float* out0, const __nv_bfloat16* A1, int pitch1,
                          int ncols1, float* out1, int64_t M, int sms, cudaStream_t st) {
  if (!out0 && !out1) return 0;
  ColsumArgs a;
  a.A[0] = out0 ? A0 : nullptr; a.out[0] = out0; a.pitch[0] = pitch0; a.ncols[0] = ncols0;
  a.A[1] = out1 ? A1 : nullptr; a.out[1] = out1; a.pitch[1] = pitch1 > 0 ? pitch1 : 8; a.ncols[1] = ncols1;
  a.M = M;
  const int rpp = CS_THREADS / (pitch0 / 8);
  int64_t bx = (M + (int64_t)rpp * 8 - 1) / ((int64_t)rpp * 8);   // >= 8 passes per block
  if (bx > 2 * sms) bx = 2 * sms;
  if (bx < 1) bx = 1;
  colsum_scratch_kernel<<<dim3((unsigned)bx, 2), CS_THREADS, 0, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}

// Dynamic shared memory of a launch: the layout of BCfg plus, when it still fits the 227 KB limit (d <= 768 at R = 96),
// the table of dropout bits (BParams::premask).
template <int R>
int bwd_smem(int d, bool gated, int* premask) {
  int smem = BCfg<R>::smem_bytes(d);
  const bool fits = gated && smem + BCfg<R>::mask_bytes(d) <= SMEM_LIMIT;
  if (fits) smem += BCfg<R>::mask_bytes(d);
  if (premask) *premask = fits ? 1 : 0;
  return smem;
}

// ---- tile split: how many CTAs share a tile -----------------------------------------------------------------------
// The largest divisor of nkc (<= 8, the portable cluster size) for which every tile still gets its own cluster in ONE wave.
// Cluster capacity is asked from the driver once per device and cluster size (GPCs strand SMs for some sizes).
template <int R, bool GATED>
int max_clusters(int ns, int d) {
  static int cache[64][9];   // shared memory is 213-225 KB for every supported d: one CTA per SM either way
  static bool init = false;
  if (!init) { memset(cache, 0, sizeof(cache)); init = true; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cache[dev][ns] != 0) return cache[dev][ns] > 0 ? cache[dev][ns] : 0;
  auto kern = k1_bwd_sm100_kernel<R, GATED, true, true>;      // (the add-gate instantiation has the same footprint)
  const int smem = bwd_smem<R>(d, GATED, nullptr);
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ns; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(ns * 16)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  } else {
    cudaGetLastError();
  }
  cache[dev][ns] = n > 0 ? n : -1;
  return n;
}
int g_bwd_split = []() { const char* e = getenv("VLPET_K1_BWD_SPLIT"); return e ? atoi(e) : -1; }();   // -1 auto, 1 never
template <int R, bool GATED>
int pick_split(int64_t tiles, int d, int sms) {
  if (g_bwd_split == 1 || tiles * 2 > sms) return 1;
  const int nkc = d / CH;
  for (int ns = 4; ns >= 2; --ns) {   // 6 measured no faster than 4 (43 us at M = 2128): the exchange grows with the cluster
    if (nkc % ns != 0 || (g_bwd_split > 1 && ns != g_bwd_split)) continue;
    if (tiles <= max_clusters<R, GATED>(ns, d)) return ns;
  }
  return 1;
}
int pick_R2(int r, int rg);
int pick_split_rt(bool gated, int R, int64_t tiles, int d, int sms) {
  switch (R) {
    case 32: return gated ? pick_split<32, true>(tiles, d, sms) : pick_split<32, false>(tiles, d, sms);
    case 64: return gated ? pick_split<64, true>(tiles, d, sms) : pick_split<64, false>(tiles, d, sms);
    case 96: return gated ? pick_split<96, true>(tiles, d, sms) : pick_split<96, false>(tiles, d, sms);
  }
  return 1;
}

struct Scratch {
  __nv_bfloat16 *zs, *qs, *das, *dps, *dus, *dts;
  float* xchg;
  int pz, pq, nsplit;
  int64_t main_tiles, split_tiles;   // launch plan: tiles [0, main_tiles) one CTA each (persistent), then split_tiles tiles with
                                     // nsplit CTAs each (small M: everything; large M: the last, partial wave)
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
  const int64_t tiles = (M + TILE_M - 1) / TILE_M;
  const int R = pick_R2(r, rg);
  const int sms = device_sm_count();
  s.main_tiles = tiles; s.split_tiles = 0; s.nsplit = 1;
  if (R && sms > 0) {
    if (tiles * 2 <= sms) {                       // small M: every tile is shared by a cluster
      s.nsplit = pick_split_rt(gated, R, tiles, d, sms);
      if (s.nsplit > 1) { s.main_tiles = 0; s.split_tiles = tiles; }
    } else if (tiles > sms && tiles % sms != 0 && (tiles % sms) * 2 <= sms) {
      // large M: the last wave is partial (750 tiles on 148 SMs: 10 tiles would cost a sixth full round) -> its tiles
      // go to a second launch, split over clusters
      const int64_t rem = tiles % sms;
      const int ns = pick_split_rt(gated, R, rem, d, sms);
      if (ns > 1) { s.nsplit = ns; s.main_tiles = tiles - rem; s.split_tiles = rem; }
    }
  }
  s.xchg = s.nsplit > 1 ? a.take<float>((size_t)s.split_tiles * s.nsplit * TILE_M * 2 * R) : nullptr;
  s.bytes = a.off;
  return s;
}

template <int R, bool GATED, bool MUL>
int launch(const VlpetK1Desc& D, const CUtensorMap* m, const BParams& p0, int sms, cudaStream_t st) {
  BParams p = p0;
  const int smem = bwd_smem<R>(D.d, GATED, &p.premask);
  auto kern = k1_bwd_sm100_kernel<R, GATED, MUL, false>;
  auto kern_split = k1_bwd_sm100_kernel<R, GATED, MUL, true>;
  static int attr_set[64] = {0};   // per device: cudaFuncSetAttribute applies to the current device's copy of the kernel
  int dev = 0;
  VLPET_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || attr_set[dev] < smem) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VLPET_CUDA_OK(cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = smem;
  }
  const int64_t tiles = p.tile_end - p.tile_begin;
  if (tiles <= 0) return 0;
  if (p.nsplit > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.nsplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(tiles * p.nsplit)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    VLPET_CUDA_OK(cudaLaunchKernelEx(&cfg, kern_split, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], p));
    count_launch();
    return 0;
  }
  const int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, NUM_THREADS, smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], p);
  VLPET_LAUNCH_OK();
  return 0;
}

int pick_R2(int r, int rg) {
  const int m = r > rg ? r : rg;
  return m <= 32 ? 32 : (m <= 64 ? 64 : (m <= 96 ? 96 : 0));
}

// shared driver of the gated (K1, large gate) and ungated (K2) backward
int run_bwd(bool gated, const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w,
            void* dx1, void* dx2, const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st,
            const BwdExtras* ex = nullptr) {
  const int rg = gated ? D.rg : D.r;
  Scratch s = carve(D.M, D.d, D.r, rg, gated, ws);
  if (!ws || ws_bytes < s.bytes) return fail(VLPET_E_WORKSPACE, "bwd(fused): workspace %zu < %zu bytes", ws_bytes, s.bytes);
  const int R = pick_R2(D.r, rg);
  const int sms = device_sm_count();
  CUtensorMap m[12];
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
  VLPET_TRY(make_map_bf16(&m[11], (ex && ex->kap_src) ? ex->kap_src : dout, M, d, d, TILE_M, CH, false));
  BParams p;
  p.M = D.M; p.d = D.d; p.r = D.r; p.rg = rg; p.add_gate = D.add_gate;
  p.s = D.s; p.alpha = D.alpha; p.kappa = D.kappa;
  p.bd = static_cast<const __nv_bfloat16*>(w.bd); p.bu = static_cast<const __nv_bfloat16*>(w.bu);
  p.gbd = static_cast<const __nv_bfloat16*>(gated ? w.gbd : w.bd); p.gbu = static_cast<const __nv_bfloat16*>(gated ? w.gbu : w.bu);
  p.zs = s.zs; p.qs = s.qs; p.das = s.das; p.dps = s.dps; p.pz = s.pz; p.pq = s.pq;
  p.nsplit = s.nsplit; p.xchg = s.xchg;
  p.dbd = G.dbd; p.dgbd = gated ? G.dgbd : nullptr;
  p.seed = D.seed;
  p.seed_dev = D.seed_dev;
  p.trace = g_trace_b;
  p.dbg = trap_buffer_dev();
  p.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  p.inv_keep = p.thr16 ? 1.0f / (1.0f - (float)p.thr16 / 65536.0f) : 1.0f;
  // developer hook (tools/run_sanitizers.sh): VLPET_DEBUG_BWD_PARTS bit 0 = run the tile kernel, bit 1 = column sums,
  // bit 2 = weight-gradient GEMM (default: all)
  const int parts = g_bwd_parts;
  int rc = 0;
  auto launch_range = [&](int64_t t0, int64_t t1, int ns) -> int {
    BParams q = p;
    q.tile_begin = t0; q.tile_end = t1; q.nsplit = ns;
    if (gated && D.add_gate == 0) {
      switch (R) {
        case 32: return launch<32, true, true>(D, m, q, sms, st);
        case 64: return launch<64, true, true>(D, m, q, sms, st);
        case 96: return launch<96, true, true>(D, m, q, sms, st);
      }
    } else if (gated) {
      switch (R) {
        case 32: return launch<32, true, false>(D, m, q, sms, st);
        case 64: return launch<64, true, false>(D, m, q, sms, st);
        case 96: return launch<96, true, false>(D, m, q, sms, st);
      }
    } else {
      switch (R) {
        case 32: return launch<32, false, true>(D, m, q, sms, st);
        case 64: return launch<64, false, true>(D, m, q, sms, st);
        case 96: return launch<96, false, true>(D, m, q, sms, st);
      }
    }
    return fail(VLPET_E_UNSUPPORTED, "bwd(fused): unsupported rank");
  };
  if (parts & 1) {
    rc = launch_range(0, s.main_tiles, 1);
    if (!rc) rc = launch_range(s.main_tiles, s.main_tiles + s.split_tiles, s.nsplit);
  }
  if (rc) return rc;
  // (round 1 ran two column-sum launches over the da / dp scratch here; the sums are now formed in epilogue 3 of the tile kernel)
  if (!(parts & 4)) return 0;
  // ---- weight gradients: dWu = du^T z (+dbu), dGu = dt^T q (+dgbu), dWd = (x2^T da)^T, dGd = (x1^T dp)^T
  //      (ungated: du = alpha*dout, so A = dout with scale alpha)
  const void* A[4]; const void* B[4]; int64_t lda[4], ldb[4], ldo[4]; int nbv[4], tr[4]; float* out[4]; float* bias[4]; float sc[4];
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
      ldo[k] = (q == 0 && ex && ex->ldo_wu > 0) ? ex->ldo_wu : 0;
      ++k;
    }
    if (k == 0) return 0;
    return wgrad_sm100(k, A, lda, B, ldb, nbv, out, bias, sc, tr, D.M, D.d, nout, sms, st, ldo);
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

int colsum_bf16(const void* A, int pitch, int ncols, float* out, int64_t M, int sms, cudaStream_t st) {
  return launch_colsum_scratch(static_cast<const __nv_bfloat16*>(A), pitch, ncols, out, nullptr, 0, 0, nullptr, M, sms, st);
}
int set_k1_bwd_parts(int parts) {
  g_bwd_parts = parts;
  return 0;
}
int set_k1_bwd_trace(unsigned long long* dev_buf) {
  g_trace_b = dev_buf;
  return 0;
}

bool fused_k1_bwd_supported(const VlpetK1Desc& D) {
  if (D.dtype != VLPET_BF16 || D.gate != VLPET_GATE_LARGE) return false;
  if (D.d % 128 != 0 || D.d < 128 || BCfg<96>::smem_bytes(D.d) > SMEM_LIMIT) return false;   // the fp32 bias tables must fit
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

int fused_k2_bwd_kappa(const VlpetK2Desc& D, float kappa, const void* kv, const void* dout, const VlpetK2Params& w, void* dkv,
                       const VlpetK2Grads& g, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || !aligned16(w.bu))
    return fail(VLPET_E_ALIGN, "k2_bwd(fused): weights must be 16-byte aligned");
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  VlpetK1Grads G;
  memset(&G, 0, sizeof(G));
  G.dWd = g.dWd; G.dbd = g.dbd; G.dWu = g.dWu; G.dbu = g.dbu;
  VlpetK1Desc K = k2_as_k1(D);
  K.kappa = kappa;
  return run_bwd(false, K, nullptr, kv, dout, P, nullptr, dkv, G, ws, ws_bytes, st);
}

int fused_k2_bwd_ex(const VlpetK2Desc& D, float kappa, const void* kv, const void* dout, const VlpetK2Params& w, void* dkv,
                    const VlpetK2Grads& g, const BwdExtras& ex, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || !aligned16(w.bu))
    return fail(VLPET_E_ALIGN, "k2_bwd(fused): weights must be 16-byte aligned");
  if (ex.kap_src && !aligned16(ex.kap_src)) return fail(VLPET_E_ALIGN, "k2_bwd(fused): kap_src must be 16-byte aligned");
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  VlpetK1Grads G;
  memset(&G, 0, sizeof(G));
  G.dWd = g.dWd; G.dbd = g.dbd; G.dWu = g.dWu; G.dbu = g.dbu;
  VlpetK1Desc K = k2_as_k1(D);
  K.kappa = kappa;
  return run_bwd(false, K, nullptr, kv, dout, P, nullptr, dkv, G, ws, ws_bytes, st, &ex);
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
