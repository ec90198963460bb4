// K1 forward, fused, sm_100a: the granularity-controlled PET module with the "large" gate in ONE launch.
//
//   out = x1 + s * ((kappa*x2 + alpha*(gelu_new(x2 Wd^T + bd) Wu^T + bu)) (*|+) sigmoid(gelu_new(x1 Gd^T + gbd) Gu^T + gbu))
//
// (my_transformers/modeling_bart.py:1145-1155, 1195-1209, 1256-1260; T5: my_transformers/modeling_t5.py:777-824, 359-409)
//
// Design.  Persistent CTAs, one per SM, each walks 128-token tiles.  Per tile:
//   phase A  (contraction over d):   A[128,R] = x2 Wd^T,  P[128,R] = x1 Gd^T        tcgen05.mma, fp32 accum in TMEM
//   epilogue 1:                      z = gelu_new(A + bd), q = gelu_new(P + gbd)    TMEM -> regs -> bf16 -> swizzled smem
//   phase B  (per 64-column chunk):  U = z Wu_n^T, T = q Gu_n^T                     tcgen05.mma, double-buffered TMEM
//   epilogue 2:                      out_n = x1_n + s*(kappa*x2_n + alpha*(U+bu)) (*|+) sigmoid(T+gbu)   -> smem -> TMA store
// Warp roles: warp 0 = TMA producer (activations), warp 3 = TMA producer (weights), warp 1 = MMA issuer (+ TMEM owner),
// warp 2 = TMA-store issuer,
// warps 4..11 = epilogue (two warpgroups; warp%4 selects the TMEM lane quarter).
// Two smem rings fed by TMA: an x-ring (x1/x2 64-column chunks, used as MMA operands in phase A and as the residual
// inputs + output staging in phase B) and a w-ring (weight chunks, always L2 hits).  Weights are padded to R rows /
// columns by TMA out-of-bounds zero fill, so any r, rg <= R that is a multiple of 8 runs on the same instantiation.
#include <mutex>
#include <type_traits>

#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with a row pitch of `pitch_elems` elements; box = [box_rows x box_cols],
// 128-byte swizzle (box_cols * 2 bytes must be <= 128), out-of-bounds elements read as zero / are not written.
// Descriptor cache: encoding a tensor map costs ~1 us on the host and a backward call needs 19 of them; the same
// (pointer, shape) tuples recur every step (weights always, activations whenever the allocator reuses addresses).
struct MapKey {
  const void* base;
  uint64_t rows, cols, pitch;
  uint32_t box_rows, box_cols, weight;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols && weight == o.weight;
  }
};
struct MapSlot { MapKey key; CUtensorMap map; bool valid; };
static MapSlot g_map_cache[256];
static std::mutex g_map_mutex;

static int encode_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                           uint32_t box_rows, uint32_t box_cols, bool weight);

int make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                  uint32_t box_cols, bool weight) {
  const MapKey k{base, rows, cols, pitch_elems, box_rows, box_cols, weight ? 1u : 0u};
  uint64_t h = reinterpret_cast<uint64_t>(base) * 0x9E3779B97F4A7C15ull ^ (rows * 1315423911ull) ^ (cols << 20) ^ (box_rows << 8) ^ box_cols;
  h ^= h >> 29;
  MapSlot& slot = g_map_cache[h & 255];
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    if (slot.valid && slot.key == k) {
      *m = slot.map;
      return 0;
    }
  }
  VLPET_TRY(encode_map_bf16(m, base, rows, cols, pitch_elems, box_rows, box_cols, weight));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  slot.key = k; slot.map = *m; slot.valid = true;
  return 0;
}

static int encode_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                           uint32_t box_rows, uint32_t box_cols, bool weight) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VLPET_E_NODEVICE, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   weight ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VLPET_E_BADARG, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box=%ux%u",
                                     (int)r, (unsigned long long)rows, (unsigned long long)cols,
                                     (unsigned long long)pitch_elems, box_rows, box_cols);
  return 0;
}

namespace {

constexpr int TILE_M = 128;
constexpr int CH = 64;  // chunk width in elements: 64 bf16 = one 128-byte swizzle row
constexpr int SW = 2;   // w-ring stages
constexpr int XCH_BYTES = TILE_M * CH * 2;  // 16 KB: one [128 x 64] bf16 chunk
constexpr int NUM_THREADS = 640;   // 4 role warps + 16 epilogue warps
constexpr int EPI_THREADS = 512;
constexpr int TMEM_COLS = 512;
constexpr int TM_A = 0, TM_P = 128, TM_UT = 256;  // TMEM column offsets
constexpr int SMEM_LIMIT = 232448;                // 227 KB opt-in maximum per CTA

template <int R>
struct Cfg {
  static constexpr int SX = 4;                         // x-ring stages
  static constexpr int KB = (R + 63) / 64;             // 64-wide K blocks of the phase-B weight chunks
  static constexpr int WA_BYTES = R * CH * 2;          // one [R x 64] weight chunk (phase A)
  static constexpr int WB_BYTES = KB * CH * CH * 2;    // one [64 x (KB*64)] weight chunk (phase B)
  static constexpr int WSLOT = (2 * WA_BYTES > 2 * WB_BYTES) ? 2 * WA_BYTES : 2 * WB_BYTES;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_W = OFF_X + SX * 2 * XCH_BYTES;
  static constexpr int OFF_BD = OFF_W + SW * WSLOT;    // fp32 bd[128], gbd[128] (zero padded)
  static constexpr int OFF_BAR = OFF_BD + 2 * 128 * 4;
  static constexpr int OFF_BU = OFF_BAR + 256;         // fp32 alpha*bu[d], 0.5*gbu[d]
  static constexpr int smem_bytes(int d) { return OFF_BU + 2 * d * 4 + 1024; }  // + slack for the manual 1024-B alignment
  // z = gelu_new(A + bd) and q = gelu_new(P + gbd) never touch shared memory: epilogue 1 packs them to bf16 and stores
  // them back into TMEM over the accumulator columns it has just read, and phase B feeds them to tcgen05.mma as the
  // A operand straight from TMEM.  Each epilogue warp owns HALF = R/2 accumulator columns and writes its HALF/2 packed
  // columns at the start of its own range, so K step ks (16 values = 8 columns) of z / q starts at this column:
  static constexpr int HALF = R / 2;
  __host__ __device__ static constexpr uint32_t zq_col(int ks) {
    return (16 * ks < HALF) ? (uint32_t)(8 * ks) : (uint32_t)(HALF + (16 * ks - HALF) / 2);
  }
};

// Optional phase-timestamp trace (tools/trace_k1.py): when non-null, thread 128 (epilogue warp 4, lane 0) of every CTA
// appends %globaltimer values at the phase boundaries of its tiles: [cta][64] slots.
static unsigned long long* g_trace = nullptr;   // host copy; travels to the kernel as Params::trace (constant bank: free when off)
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VLPET_TRACE(slot)                                                                      \
  do {                                                                                         \
    if (p.trace && threadIdx.x == 128 && w == blockIdx.x && (slot) < 64) p.trace[blockIdx.x * 64 + (slot)] = gtimer(); \
  } while (0)

struct Params {
  int64_t M;
  int d, r, rg;
  int add_gate;
  int64_t full_tiles; // tiles [0, full_tiles) are one work item each; every later tile is split over nsplit work items
  int nsplit;         // split tiles (the last, partial wave of a large M, or all tiles of a small M): nsplit CTAs each redo
                      // phase A and own nkc/nsplit of the phase-B column chunks -- trades redundant L2 reads for a
                      // shorter critical path
  float s, alpha, kappa;
  const __nv_bfloat16 *bd, *bu, *gbd, *gbu;
  uint64_t seed;      // dropout stream (vlpet_common.cuh: drop_hash4)
  const uint64_t* seed_dev;  // optional device scalar added to seed (CUDA-graph replays)
  uint32_t thr16;     // 0 = no dropout
  float inv_keep;
  unsigned long long* trace;  // developer hook (tools/trace_k1.py), normally null
};

// barrier slots (8 bytes each) inside the barrier block
constexpr int SXM = 4;  // barrier slots are laid out for the deepest x-ring
enum { B_XFULL = 0, B_XEMPTY = B_XFULL + SXM, B_WFULL = B_XEMPTY + SXM, B_WEMPTY = B_WFULL + SW, B_APFULL = B_WEMPTY + SW,
       B_ZQFULL, B_UTFULL, B_UTEMPTY = B_UTFULL + 2, B_OUTRDY = B_UTEMPTY + 2, B_COUNT = B_OUTRDY + SXM };

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
using namespace ptx;   // f2 helpers (packed fp32 pairs)
// gelu_new(v) = 0.5 v (1 + tanh(c (v + 0.044715 v^3)))
__device__ __forceinline__ f2 gelu_new2(f2 v) {
  const f2 ck = mk2(0.7978845608028654f * 0.044715f, 0.7978845608028654f * 0.044715f);
  const f2 c = mk2(0.7978845608028654f, 0.7978845608028654f), half = mk2(0.5f, 0.5f);
  const f2 t = tanh2(mul2(v, fma2(ck, mul2(v, v), c)));
  const f2 hv = mul2(half, v);
  return fma2(hv, t, hv);
}
template <int R, bool GATED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_wd,
                    const __grid_constant__ CUtensorMap tm_gd, const __grid_constant__ CUtensorMap tm_wu,
                    const __grid_constant__ CUtensorMap tm_gu, const Params p) {
  using C = Cfg<R>;
  constexpr int SX = C::SX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));   // generic-space view of the aligned base
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;  // 4 bytes: TMEM base address written by tcgen05.alloc

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int nkc = p.d / CH;  // chunks along d (phase A: K chunks; phase B: N chunks)
  const int64_t num_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int64_t num_items = p.full_tiles + (num_tiles - p.full_tiles) * p.nsplit;
  const int cps = nkc / p.nsplit;   // phase-B chunks per split work item (nsplit divides nkc)
#define WORK_ITEM()                                                    \
  int64_t tile = w;                                                    \
  int cb = 0, ce = nkc;                                                \
  if (w >= p.full_tiles) {                                             \
    const int64_t t_ = w - p.full_tiles;                               \
    tile = p.full_tiles + t_ / p.nsplit;                               \
    cb = (int)(t_ % p.nsplit) * cps;                                   \
    ce = cb + cps;                                                     \
  }                                                                    \
  (void)tile; (void)cb; (void)ce

  if (threadIdx.x == 0) {
    for (int i = 0; i < SX; ++i) { ptx::mbar_init(bar(B_XFULL + i), 1); ptx::mbar_init(bar(B_XEMPTY + i), 1); ptx::mbar_init(bar(B_OUTRDY + i), EPI_THREADS); }
    for (int i = 0; i < SW; ++i) { ptx::mbar_init(bar(B_WFULL + i), 1); ptx::mbar_init(bar(B_WEMPTY + i), 1); }
    ptx::mbar_init(bar(B_APFULL), 1);
    ptx::mbar_init(bar(B_ZQFULL), EPI_THREADS);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar(B_UTFULL + i), 1); ptx::mbar_init(bar(B_UTEMPTY + i), EPI_THREADS); }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_out); }
  if (warp == 3 && lane == 0) {
    ptx::prefetch_tmap(&tm_wd); ptx::prefetch_tmap(&tm_gd); ptx::prefetch_tmap(&tm_wu); ptx::prefetch_tmap(&tm_gu);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  {  // biases -> fp32 in shared memory, pre-multiplied so that the epilogues fold them into FMAs they issue anyway:
     // bd / gbd (zero padded to 128), alpha*bu, 0.5*gbu
    float* sbd = reinterpret_cast<float*>(smem_gen + C::OFF_BD);
    float* sbu = reinterpret_cast<float*>(smem_gen + C::OFF_BU);
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) {
      const int j = i & 127;
      float v = 0.f;
      if (i < 128) { if (j < p.r) v = __bfloat162float(p.bd[j]); }
      else if (GATED) { if (j < p.rg) v = __bfloat162float(p.gbd[j]); }
      sbd[i] = v;
    }
    for (int i = threadIdx.x; i < p.d; i += NUM_THREADS) {
      sbu[i] = p.alpha * __bfloat162float(p.bu[i]);
      sbu[p.d + i] = GATED ? 0.5f * __bfloat162float(p.gbu[i]) : 0.f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================== TMA producer: activations =====================================
    if (lane == 0) {
      uint32_t xi = 0;
      for (int64_t w = blockIdx.x; w < num_items; w += gridDim.x) {
        WORK_ITEM();
        const int row0 = (int)(tile * TILE_M);
        for (int ph = 0; ph < 2; ++ph) {
          for (int c = (ph ? cb : 0); c < (ph ? ce : nkc); ++c, ++xi) {
            const uint32_t sx = xi % SX;
            ptx::mbar_wait(bar(B_XEMPTY + sx), ((xi / SX) & 1) ^ 1);
            const uint32_t xdst = smem_base + C::OFF_X + sx * (2 * XCH_BYTES);
            ptx::mbar_arrive_expect_tx(bar(B_XFULL + sx), 2 * XCH_BYTES);
            ptx::tma_load_2d(xdst, &tm_x1, c * CH, row0, bar(B_XFULL + sx));
            ptx::tma_load_2d(xdst + XCH_BYTES, &tm_x2, c * CH, row0, bar(B_XFULL + sx));
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA producer: weights (always L2 hits) =====================================
    // its own thread, so that the weight ring runs ahead independently of the activation ring's (later) releases
    if (lane == 0) {
      uint32_t wi = 0;
      for (int64_t w = blockIdx.x; w < num_items; w += gridDim.x) {
        WORK_ITEM();
        for (int ph = 0; ph < 2; ++ph) {
          for (int c = (ph ? cb : 0); c < (ph ? ce : nkc); ++c, ++wi) {
            const uint32_t sw = wi % SW;
            ptx::mbar_wait(bar(B_WEMPTY + sw), ((wi / SW) & 1) ^ 1);
            const uint32_t wdst = smem_base + C::OFF_W + sw * C::WSLOT;
            if (ph == 0) {
              ptx::mbar_arrive_expect_tx(bar(B_WFULL + sw), (GATED ? 2 : 1) * C::WA_BYTES);
              ptx::tma_load_2d(wdst, &tm_wd, c * CH, 0, bar(B_WFULL + sw));
              if (GATED) ptx::tma_load_2d(wdst + C::WA_BYTES, &tm_gd, c * CH, 0, bar(B_WFULL + sw));
            } else {
              ptx::mbar_arrive_expect_tx(bar(B_WFULL + sw), (GATED ? 2 : 1) * C::WB_BYTES);
#pragma unroll
              for (int kb = 0; kb < C::KB; ++kb) {
                ptx::tma_load_2d(wdst + kb * (CH * CH * 2), &tm_wu, kb * CH, c * CH, bar(B_WFULL + sw));
                if (GATED)
                  ptx::tma_load_2d(wdst + C::WB_BYTES + kb * (CH * CH * 2), &tm_gu, kb * CH, c * CH, bar(B_WFULL + sw));
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t IDESC_A = ptx::umma_idesc_bf16_m128(R);
      constexpr uint32_t IDESC_B = ptx::umma_idesc_bf16_m128(CH);
      uint32_t xi = 0, wi = 0, ui = 0, ti = 0;
      for (int64_t w = blockIdx.x; w < num_items; w += gridDim.x, ++ti) {
        WORK_ITEM();
        // ---- phase A: A += x2_c Wd_c^T ; P += x1_c Gd_c^T
        for (int c = 0; c < nkc; ++c, ++xi, ++wi) {
          const uint32_t sx = xi % SX, sw = wi % SW;
          ptx::mbar_wait(bar(B_XFULL + sx), (xi / SX) & 1);
          ptx::mbar_wait(bar(B_WFULL + sw), (wi / SW) & 1);
          ptx::tc_fence_after();
          const uint32_t x1s = smem_base + C::OFF_X + sx * (2 * XCH_BYTES), x2s = x1s + XCH_BYTES;
          const uint32_t wds = smem_base + C::OFF_W + sw * C::WSLOT, gds = wds + C::WA_BYTES;
#pragma unroll
          for (int ks = 0; ks < CH / 16; ++ks) {
            const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
            ptx::umma_bf16_ss(tmem_base + TM_A, ptx::umma_desc_kmajor_sw128(x2s + ks * 32),
                              ptx::umma_desc_kmajor_sw128(wds + ks * 32), IDESC_A, acc);
            if (GATED)
              ptx::umma_bf16_ss(tmem_base + TM_P, ptx::umma_desc_kmajor_sw128(x1s + ks * 32),
                                ptx::umma_desc_kmajor_sw128(gds + ks * 32), IDESC_A, acc);
          }
          ptx::umma_commit(bar(B_XEMPTY + sx));
          ptx::umma_commit(bar(B_WEMPTY + sw));
        }
        ptx::umma_commit(bar(B_APFULL));
        // ---- phase B: U_n = z Wu_n^T ; T_n = q Gu_n^T
        ptx::mbar_wait(bar(B_ZQFULL), ti & 1);
        ptx::tc_fence_after();
        for (int c = cb; c < ce; ++c, ++xi, ++wi, ++ui) {
          const uint32_t sw = wi % SW, ub = ui & 1;
          ptx::mbar_wait(bar(B_WFULL + sw), (wi / SW) & 1);
          ptx::mbar_wait(bar(B_UTEMPTY + ub), ((ui >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t wus = smem_base + C::OFF_W + sw * C::WSLOT, gus = wus + C::WB_BYTES;
          const uint32_t tU = tmem_base + TM_UT + ub * 128, tT = tU + 64;
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks) {
            const int kb = ks / 4, kin = ks % 4;
            ptx::umma_bf16_ts(tU, tmem_base + TM_A + C::zq_col(ks),
                              ptx::umma_desc_kmajor_sw128(wus + kb * (CH * CH * 2) + kin * 32), IDESC_B, ks > 0);
            if (GATED)
              ptx::umma_bf16_ts(tT, tmem_base + TM_P + C::zq_col(ks),
                                ptx::umma_desc_kmajor_sw128(gus + kb * (CH * CH * 2) + kin * 32), IDESC_B, ks > 0);
          }
          ptx::umma_commit(bar(B_WEMPTY + sw));
          ptx::umma_commit(bar(B_UTFULL + ub));
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== TMA store issuer =====================================
    if (lane == 0) {
      uint32_t xi = 0, oi = 0;
      for (int64_t w = blockIdx.x; w < num_items; w += gridDim.x) {
        WORK_ITEM();
        const int row0 = (int)(tile * TILE_M);
        xi += nkc;  // phase A steps of the x-ring are consumed by the MMA warp
        for (int c = cb; c < ce; ++c, ++xi, ++oi) {
          const uint32_t sx = xi % SX, so = oi % SX;
          ptx::mbar_wait(bar(B_OUTRDY + so), (oi / SX) & 1);
          ptx::tma_store_2d(&tm_out, smem_base + C::OFF_X + sx * (2 * XCH_BYTES), c * CH, row0);
          ptx::tma_store_commit();
          ptx::tma_store_wait_read0();
          ptx::mbar_arrive(bar(B_XEMPTY + sx));
        }
      }
      ptx::tma_store_wait_all0();
    }
  } else if (warp >= 4) {
    // ===================================== epilogue warps =====================================
    // 16 warps: warp%4 = TMEM lane quarter (hardware rule), cg = (warp-4)/4 = column group.  Four warps per scheduler
    // keep the issue slots busy while tcgen05.ld / ld.shared / MUFU latencies are in flight.  All arithmetic is on
    // packed fp32 pairs (FFMA2), biases come pre-scaled from shared memory: ~8 issue slots per output element.
    const int quarter = warp % 4;            // TMEM lanes [32*quarter, 32*quarter+32)
    const int cg = (warp - 4) / 4;           // epilogue 1: branch = cg/2 (z | q), half of its columns = cg%2; epilogue 2: 16 of 64 columns
    const int row = quarter * 32 + lane;     // row inside the 128-token tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    const uint64_t seed_eff = p.seed + ((p.thr16 && p.seed_dev) ? __ldg(p.seed_dev) : 0ull);
    const f2 kappa2 = mk2(p.kappa, p.kappa), alpha2 = mk2(p.alpha, p.alpha), half2 = mk2(0.5f, 0.5f);
    const f2 s2 = mk2(p.s, p.s), sh2 = mk2(0.5f * p.s, 0.5f * p.s);
    const float s_keep = p.s * p.inv_keep;
    uint32_t xi = 0, ui = 0, oi = 0, ti = 0;
    for (int64_t w = blockIdx.x; w < num_items; w += gridDim.x, ++ti) {
        WORK_ITEM();
      // ---- epilogue 1: z = gelu_new(A + bd) (cg 0,1) / q = gelu_new(P + gbd) (cg 2,3) -> packed bf16, back into TMEM
      VLPET_TRACE(0);
      ptx::mbar_wait(bar(B_APFULL), ti & 1);
      VLPET_TRACE(1);
      ptx::tc_fence_after();
      const int branch = cg >> 1;
      if (GATED || branch == 0) {
        const uint32_t tsrc = lane_addr + (branch ? TM_P : TM_A);
        const uint32_t sbias = smem_base + C::OFF_BD + (uint32_t)branch * 512u;
        constexpr int HALF = C::HALF;          // columns per warp (R % 32 == 0 -> a multiple of 16)
        const int jbeg = (cg & 1) * HALF;
#pragma unroll
        for (int jj = 0; jj < HALF; jj += 16) {
          const int j0 = jbeg + jj;
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(tsrc + j0, v);
          f2 bias[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) lds_f2x2(sbias + (uint32_t)(j0 + e * 4) * 4u, bias[2 * e], bias[2 * e + 1]);
          ptx::tmem_ld_wait();
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = pack2(gelu_new2(add2(mk2u(v[e * 2], v[e * 2 + 1]), bias[e])));
          ptx::tmem_st_32x32b_x8(tsrc + jbeg + jj / 2, o);   // packed bf16 over columns this warp has already read
        }
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(B_ZQFULL));
      VLPET_TRACE(2);
      xi += nkc;
      // ---- epilogue 2, per 64-column chunk: this warp owns columns [cg*16, cg*16+16) of the chunk
      for (int c = cb; c < ce; ++c, ++xi, ++ui, ++oi) {
        const uint32_t sx = xi % SX, ub = ui & 1, so = oi % SX;
        const int col0 = c * CH + cg * 16;  // first of this thread's 16 output columns
        ptx::mbar_wait(bar(B_UTFULL + ub), (ui >> 1) & 1);
        VLPET_TRACE(3 + 4 * (c - cb));
        ptx::tc_fence_after();
        uint32_t u[16], t[16];
        const uint32_t tU = lane_addr + TM_UT + ub * 128 + cg * 16;
        ptx::tmem_ld_32x32b_x16(tU, u);
        if (GATED) ptx::tmem_ld_32x32b_x16(tU + 64, t);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(B_UTEMPTY + ub));  // accumulators are in registers: the MMA warp may overwrite them
        VLPET_TRACE(4 + 4 * (c - cb));
        ptx::mbar_wait(bar(B_XFULL + sx), (xi / SX) & 1);
        VLPET_TRACE(5 + 4 * (c - cb));
        const uint32_t x1row = smem_base + C::OFF_X + sx * (2 * XCH_BYTES) + (uint32_t)row * 128u;
        const uint32_t x2row = x1row + XCH_BYTES;
        const uint32_t sabu = smem_base + C::OFF_BU + (uint32_t)col0 * 4u;   // alpha*bu for this thread's columns
        const uint32_t shgb = sabu + (uint32_t)p.d * 4u;                      // 0.5*gbu
        const int64_t idx0 = ((int64_t)tile * TILE_M + row) * p.d + col0;  // flat element index (dropout stream)
        auto group = [&](auto drop_tag, int g) {
          constexpr bool DROP = decltype(drop_tag)::value;
          const uint32_t off = (((uint32_t)(cg * 2 + g)) ^ swz) << 4;
          uint32_t a[4], b[4], o[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(x1row + off));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(x2row + off));
          f2 abu[4], hgb[4];
          lds_f2x2(sabu + g * 32, abu[0], abu[1]);
          lds_f2x2(sabu + g * 32 + 16, abu[2], abu[3]);
          if (GATED) {
            lds_f2x2(shgb + g * 32, hgb[0], hgb[1]);
            lds_f2x2(shgb + g * 32 + 16, hgb[2], hgb[3]);
          }
          f2 sc[4] = {s2, s2, s2, s2}, sh[4] = {sh2, sh2, sh2, sh2};   // s * dropout mask (and half of it), per element
          if (DROP) {
            const uint64_t h0 = drop_hash4(seed_eff, (uint64_t)(idx0 + g * 8) >> 2);
            const uint64_t h1 = drop_hash4(seed_eff, ((uint64_t)(idx0 + g * 8) >> 2) + 1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t two = (uint32_t)((e >> 1 ? h1 : h0) >> (32 * (e & 1)));  // 16 bits for column j, 16 for j+1
              sc[e] = mk2(((two & 0xffffu) >= p.thr16) ? s_keep : 0.f, ((two >> 16) >= p.thr16) ? s_keep : 0.f);
              sh[e] = mul2(sc[e], half2);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            const f2 x1p = bf2_to_f2(a[e]);
            f2 y = fma2(kappa2, bf2_to_f2(b[e]), abu[e]);          // kappa*x2 + alpha*bu
            y = fma2(alpha2, mk2u(u[j], u[j + 1]), y);             // + alpha*U
            f2 res;
            if (GATED) {
              // G = sigmoid(T + gbu) = 0.5 + 0.5*tanh(0.5*(T + gbu))
              const f2 th = tanh2(fma2(half2, mk2u(t[j], t[j + 1]), hgb[e]));
              if (!p.add_gate) {        // out = x1 + sc*y*G = (x1 + hs) + hs*th,  hs = 0.5*sc*y
                const f2 hs = mul2(sh[e], y);
                res = fma2(hs, th, add2(x1p, hs));
              } else {                  // out = x1 + sc*(y + G) = (x1 + sc*y + 0.5*sc) + 0.5*sc*th
                res = fma2(sh[e], th, add2(fma2(sc[e], y, x1p), sh[e]));
              }
            } else {
              res = fma2(sc[e], y, x1p);
            }
            o[e] = pack2(res);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(x1row + off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        };
        if (p.thr16) { group(std::true_type{}, 0); group(std::true_type{}, 1); }
        else { group(std::false_type{}, 0); group(std::false_type{}, 1); }
        ptx::fence_proxy_async_smem();  // out chunk (in the x1 slot) is read by the TMA store
        ptx::mbar_arrive(bar(B_OUTRDY + so));
        VLPET_TRACE(6 + 4 * (c - cb));
      }
      if (p.trace && threadIdx.x == 128 && ti < 12) p.trace[blockIdx.x * 64 + 52 + ti] = gtimer();   // end of work item ti
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool weight) {
  return make_map_bf16(m, base, rows, cols, cols, box_rows, (uint32_t)CH, weight);
}

int pick_R(const VlpetK1Desc& D) {
  int m = (D.gate == VLPET_GATE_LARGE && D.rg > D.r) ? D.rg : D.r;
  if (m <= 32) return 32;
  if (m <= 64) return 64;
  if (m <= 96) return 96;
  if (m <= 128) return 128;
  return 0;
}

struct DevInfo { int ok, sms, major; };
const DevInfo& dev_info() {
  static DevInfo di = []() {
    DevInfo d{0, 0, 0};
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      d.ok = 1; d.sms = p.multiProcessorCount; d.major = p.major;
    }
    return d;
  }();
  return di;
}

template <int R, bool GATED>
int launch(const VlpetK1Desc& D, const CUtensorMap* maps, const Params& p, cudaStream_t st) {
  using C = Cfg<R>;
  const int smem = C::smem_bytes(D.d);
  static int attr_set = 0;
  if (attr_set < smem) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(k1_fwd_sm100_kernel<R, GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = smem;
  }
  const int64_t tiles = (D.M + TILE_M - 1) / TILE_M;
  const int64_t items = p.full_tiles + (tiles - p.full_tiles) * p.nsplit;
  int grid = (int)(items < dev_info().sms ? items : dev_info().sms);
  k1_fwd_sm100_kernel<R, GATED><<<grid, NUM_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5],
                                                                 maps[6], p);
  VLPET_LAUNCH_OK();
  return 0;
}

template <int R>
int launch_r(const VlpetK1Desc& D, const CUtensorMap* maps, const Params& p, bool gated, cudaStream_t st) {
  return gated ? launch<R, true>(D, maps, p, st) : launch<R, false>(D, maps, p, st);
}

int smem_need(int R, int d) {
  switch (R) {
    case 32: return Cfg<32>::smem_bytes(d);
    case 64: return Cfg<64>::smem_bytes(d);
    case 96: return Cfg<96>::smem_bytes(d);
    case 128: return Cfg<128>::smem_bytes(d);
  }
  return 1 << 30;
}

}  // namespace

int set_k1_trace(unsigned long long* dev_buf) {
  g_trace = dev_buf;
  return 0;
}

bool fused_k1_fwd_supported(const VlpetK1Desc& D) {
  if (D.dtype != VLPET_BF16 || (D.gate != VLPET_GATE_LARGE && D.gate != VLPET_GATE_NONE)) return false;
  if (D.d % CH != 0 || D.d < CH) return false;
  const bool gated = D.gate == VLPET_GATE_LARGE;
  if (D.r % 8 != 0 || (gated && D.rg % 8 != 0) || pick_R(D) == 0) return false;
  if (D.M <= 0 || D.M > (int64_t)0x7fffff00) return false;
  if (smem_need(pick_R(D), D.d) > SMEM_LIMIT) return false;   // the fp32 bias tables (8 bytes per column of d) must fit
  const DevInfo& di = dev_info();
  return di.ok && di.major == 10;
}

size_t fused_k1_fwd_ws(const VlpetK1Desc&) { return 0; }

int fused_k1_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, void*, size_t,
                 cudaStream_t st) {
  const bool gated = D.gate == VLPET_GATE_LARGE;
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || (gated && (!aligned16(w.Gd) || !aligned16(w.Gu))))
    return fail(VLPET_E_ALIGN, "k1_fwd(fused): weights must be 16-byte aligned");
  const int R = pick_R(D);
  CUtensorMap maps[7];
  VLPET_TRY(make_map(&maps[0], x1, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[1], x2, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[2], out, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[3], w.Wd, (uint64_t)D.r, (uint64_t)D.d, (uint32_t)R, true));
  VLPET_TRY(make_map(&maps[4], gated ? w.Gd : w.Wd, (uint64_t)(gated ? D.rg : D.r), (uint64_t)D.d, (uint32_t)R, true));
  VLPET_TRY(make_map(&maps[5], w.Wu, (uint64_t)D.d, (uint64_t)D.r, CH, true));
  VLPET_TRY(make_map(&maps[6], gated ? w.Gu : w.Wu, (uint64_t)D.d, (uint64_t)(gated ? D.rg : D.r), CH, true));
  Params p;
  p.M = D.M; p.d = D.d; p.r = D.r; p.rg = gated ? D.rg : D.r; p.add_gate = D.add_gate;
  p.s = D.s; p.alpha = D.alpha; p.kappa = D.kappa;
  p.bd = static_cast<const __nv_bfloat16*>(w.bd); p.bu = static_cast<const __nv_bfloat16*>(w.bu);
  p.gbd = static_cast<const __nv_bfloat16*>(gated ? w.gbd : w.bd); p.gbu = static_cast<const __nv_bfloat16*>(gated ? w.gbu : w.bu);
  p.seed = D.seed;
  p.seed_dev = D.seed_dev;
  p.trace = g_trace;
  {  // Whole waves of tiles run one tile per CTA.  The tiles of the last, partial wave (all tiles when M is small) are
     // split over several CTAs each (largest divisor of nkc that still fits the wave): the tail costs one phase A plus
     // nkc/nsplit chunks instead of a full tile time.
    const int nkc = D.d / CH, sms = dev_info().sms;
    const int64_t tiles = (D.M + TILE_M - 1) / TILE_M;
    const int64_t rem = tiles % sms;
    p.full_tiles = tiles - rem;
    p.nsplit = 1;
    for (int ns = 2; ns <= nkc; ++ns)
      if (nkc % ns == 0 && rem * ns <= sms) p.nsplit = ns;
    if (p.nsplit == 1) p.full_tiles = tiles;
  }
  p.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  p.inv_keep = p.thr16 ? 1.0f / (1.0f - (float)p.thr16 / 65536.0f) : 1.0f;
  switch (R) {
    case 32: return launch_r<32>(D, maps, p, gated, st);
    case 64: return launch_r<64>(D, maps, p, gated, st);
    case 96: return launch_r<96>(D, maps, p, gated, st);
    case 128: return launch_r<128>(D, maps, p, gated, st);
  }
  return fail(VLPET_E_UNSUPPORTED, "k1_fwd(fused): unsupported rank");
}

}  // namespace vlpet
