// K1 forward, fused, sm_100a: the granularity-controlled PET module with the "large" gate in ONE launch.
//
//   out = x1 + s * ((kappa*x2 + alpha*(gelu_new(x2 Wd^T + bd) Wu^T + bu)) (*|+) sigmoid(gelu_new(x1 Gd^T + gbd) Gu^T + gbu))
//
// (my_transformers/modeling_bart.py:1145-1155, 1195-1209, 1256-1260; T5: my_transformers/modeling_t5.py:777-824, 359-409)
//
// Design.  Persistent CTAs, one per SM, each walks 128-token tiles.  Per tile:
//   phase A  (contraction over d):   A[128,R] = x2 Wd^T,  P[128,R] = x1 Gd^T        tcgen05.mma, fp32 accum in TMEM
//   epilogue 1:                      z = gelu_new(A + bd), q = gelu_new(P + gbd)    TMEM -> regs -> packed bf16 -> TMEM
//   phase B  (per 64-column chunk):  U = z Wu_n^T, T = q Gu_n^T                     tcgen05.mma, A operand from TMEM
//   epilogue 2:                      out_n = x1_n + s*(kappa*x2_n + alpha*(U+bu)) (*|+) sigmoid(T+gbu)   -> smem -> TMA store
// Software pipeline across tiles: phase A of tile i+1 (the HBM stream) is interleaved step by step with phase B of tile i
// (L2 re-reads + stores), so that HBM, the tensor pipe and the epilogue warps are busy at the same time instead of all
// CTAs alternating in lockstep between an HBM-bound and an epilogue-bound half (tools/trace_k1.py showed exactly that).
// Warp roles: warp 0 = TMA producer of the phase-A operands (x chunks + Wd/Gd chunks), warp 3 = TMA producer of the
// phase-B weights (Wu/Gu chunks), warp 1 = MMA issuer (+ TMEM owner; issues phase-A steps and phase-B chunks in
// whatever order their inputs become ready), warp 2 = phase-B x manager (TMA load of the
// residual chunks, TMA store of the finished out chunk from the same slot), warps 4..19 = epilogue (warp%4 = TMEM lane
// quarter).  Shared memory: XA ring (2 x 32 KB), XB ring (2 x 32 KB), WA ring (3 x [R x 64]), WB ring (3 x [64 x 128]),
// fp32 bias tables.  Weights are padded to R rows / columns by TMA out-of-bounds zero fill, so any r, rg <= R that is a
// multiple of 8 runs on the same instantiation.  The kernel is shared-memory-bandwidth bound (TMA writes + UMMA operand
// reads + epilogue LDS/STS all go through the 128 B/clk port); keeping z/q in TMEM removed a quarter of that traffic.
#include <cstdlib>
#include <cstring>
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
// 128-byte swizzle (64-byte swizzle when the box is 32 columns wide), out-of-bounds elements read as zero / are not written.
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
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
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
constexpr int XCH_BYTES = TILE_M * CH * 2;  // 16 KB: one [128 x 64] bf16 chunk
constexpr int NUM_THREADS = 640;   // 4 role warps + 16 epilogue warps
constexpr int EPI_THREADS = 512;
constexpr int TMEM_COLS = 512;
// TMEM columns: fp32 accumulators A, P | packed bf16 z, q (2 values per column) | U, T of one 64-column chunk
constexpr int TM_A = 0, TM_P = 128, TM_Z = 256, TM_Q = 320, TM_UT = 384;
constexpr int SMEM_LIMIT = 232448;                // 227 KB opt-in maximum per CTA

template <int R, int NCTA = 1>
struct Cfg {
  // NCTA = 2: CTA pairs (cta_group::2).  One UMMA spans the two 128-token tiles of a pair (M = 256) and every CTA stages
  // only HALF of each weight chunk (its N/2 rows of the B operand): half the weight bytes through the shared-memory port
  // -- the kernel's bottleneck -- and room for a third XB stage.
  static constexpr int NXA = 2;                        // XA ring stages (phase-A operands of the NEXT tile: the HBM stream)
  static constexpr int NXB = (NCTA == 2) ? 3 : 2;      // XB ring stages (phase-B residual inputs + out staging); a third
                                                       // XA stage instead measured slower (132 vs 128 us)
  // Weight rings hold 4 slots = two full steps (a step consumes one chunk of each of the two branches): with 3 slots the
  // second matrix of every step was requested only after the previous step's MMAs had retired -- one exposed L2 latency
  // (~1 us) per step (tools/trace_k1.py).
  static constexpr int NWA = (R == 128) ? 2 : 3;       // WA ring slots (one [R x 64] chunk of Wd or Gd each); phase A is not
                                                       // the critical path, the MMA warp never blocks on it (ready_a)
  static constexpr int NWB = (R == 128) ? 2 : 4;       // WB ring slots (one [64 x R] chunk of Wu or Gu each)
  static constexpr int KBF = R / 64;                   // full 64-wide K blocks of a phase-B weight chunk (128-byte swizzle)
  static constexpr int REM = R % 64;                   // 0 or 32: a last 32-wide K block (64-byte swizzle, half the bytes)
  static constexpr int WA_ROWS = R / NCTA;             // rows of a [R x 64] phase-A weight chunk staged by this CTA
  static constexpr int WB_ROWS = CH / NCTA;            // rows of a [64 x R] phase-B weight chunk staged by this CTA
  static constexpr int WA_BYTES = WA_ROWS * CH * 2;
  static constexpr int WB_BLK = WB_ROWS * CH * 2;      // one 64-wide K block of the staged rows
  static constexpr int WB_BYTES = KBF * WB_BLK + (REM ? WB_ROWS * 32 * 2 : 0);
  static constexpr int OFF_XA = 0;
  static constexpr int OFF_XB = OFF_XA + NXA * 2 * XCH_BYTES;
  static constexpr int OFF_WA = OFF_XB + NXB * 2 * XCH_BYTES;
  static constexpr int OFF_WB = OFF_WA + NWA * WA_BYTES;
  static constexpr int OFF_BD = OFF_WB + NWB * WB_BYTES;   // fp32 bd[128], gbd[128] (zero padded)
  static constexpr int OFF_BAR = OFF_BD + 2 * 128 * 4;
  static constexpr int OFF_BU = OFF_BAR + 512;         // fp32 alpha*bu[d], 0.5*gbu[d] (the barrier block is 512 bytes)
  static constexpr int smem_bytes(int d) { return OFF_BU + 2 * d * 4 + 1024; }  // + slack for the manual 1024-B alignment
  static constexpr int HALF = R / 2;                   // accumulator columns per epilogue-1 warp
  static_assert(REM == 0 || REM == 32, "rank buckets are multiples of 32");
  static_assert(WA_BYTES % 1024 == 0 && WB_BYTES % 1024 == 0, "swizzle atoms must stay 1024-byte aligned");
};

// Optional phase-timestamp trace (tools/trace_k1.py): when non-null, thread 128 (epilogue warp 4, lane 0) of every CTA
// stamps %globaltimer at the phase boundaries of its first work item and at the end of every work item; the MMA thread and
// the phase-B x manager stamp their issue times of the first item: [cta][256] slots.
static unsigned long long* g_trace = nullptr;   // host copy; travels to the kernel as Params::trace (constant bank: free when off)
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef VLPET_TRACE   // developer builds only (tools/_variant.py)
#define VLPET_TR(slot)                                                             \
  do {                                                                             \
    if (p.trace) p.trace[blockIdx.x * 256 + (slot)] = gtimer();                    \
  } while (0)
#define VLPET_TRACE(slot)                                                                      \
  do {                                                                                         \
    if (p.trace && threadIdx.x == 128 && ti == 0 && (slot) < 52) p.trace[blockIdx.x * 256 + (slot)] = gtimer(); \
  } while (0)
#else
#define VLPET_TR(slot) do { } while (0)
#define VLPET_TRACE(slot) do { } while (0)
#endif

struct Params {
  int64_t M;
  int d, r, rg;
  int add_gate;
  int64_t num_units;  // work units: 128-token tiles (NCTA = 1) or pairs of adjacent tiles (NCTA = 2)
  int64_t full_tiles; // units [0, full_tiles) are one work item each; every later unit is split over nsplit work items
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
  uint32_t* dbg;              // host-mapped trap record (ptx::mbar_wait_dbg)
};

// barrier slots (8 bytes each) inside the barrier block
enum { B_XAFULL = 0, B_XAEMPTY = B_XAFULL + 3, B_XBFULL = B_XAEMPTY + 3, B_OUTRDY = B_XBFULL + 3,
       B_WAFULL = B_OUTRDY + 3, B_WAEMPTY = B_WAFULL + 4, B_WBFULL = B_WAEMPTY + 4, B_WBEMPTY = B_WBFULL + 4,
       B_APFULL = B_WBEMPTY + 4, B_ZQFULL, B_UTFULL, B_UTEMPTY, B_COUNT };
static_assert(8 * B_COUNT + 8 <= 512, "barriers + the TMEM address slot must fit the 512-byte barrier block");

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
// The work items of one CTA, in order: w = blockIdx.x, + gridDim.x, ...  (tile, [cb, ce) = its phase-B chunks)
struct ItemCursor {
  int64_t w, num_items, full_tiles, tile;
  int nsplit, cps, nkc, stride, cb, ce;
  __device__ __forceinline__ void init(const Params& p, int nkc_, int64_t num_items_, int ncta = 1) {
    w = blockIdx.x / ncta; num_items = num_items_; full_tiles = p.full_tiles; nsplit = p.nsplit; nkc = nkc_;
    cps = nkc_ / p.nsplit; stride = gridDim.x / ncta;
    decode();
  }
  __device__ __forceinline__ bool valid() const { return w < num_items; }
  __device__ __forceinline__ void decode() {
    tile = w; cb = 0; ce = nkc;
    if (w >= full_tiles) {
      const int64_t t = w - full_tiles;
      tile = full_tiles + t / nsplit;
      cb = (int)(t % nsplit) * cps;
      ce = cb + cps;
    }
  }
  __device__ __forceinline__ void next() { w += stride; decode(); }
};

template <int R, bool GATED, int NCTA>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_wd,
                    const __grid_constant__ CUtensorMap tm_gd, const __grid_constant__ CUtensorMap tm_wu,
                    const __grid_constant__ CUtensorMap tm_gu, const __grid_constant__ CUtensorMap tm_wu_t,
                    const __grid_constant__ CUtensorMap tm_gu_t, const Params p) {
  using C = Cfg<R, NCTA>;
  constexpr int NWA = C::NWA, NWB = C::NWB, NXA = C::NXA, NXB = C::NXB;
  const uint32_t rank = (NCTA == 2) ? ptx::cluster_ctarank() : 0u;   // position in the CTA pair; rank 0 issues the MMAs
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));   // generic-space view of the aligned base
  const uint32_t bar_base = smem_base + C::OFF_BAR;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bar_base + 8u * B_COUNT;  // 4 bytes: TMEM base address written by tcgen05.alloc

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int nkc = p.d / CH;  // chunks along d (phase A: K chunks; phase B: N chunks)
  const int64_t num_items = p.full_tiles + (p.num_units - p.full_tiles) * p.nsplit;
  // barrier `i` of the pair leader as a shared::cluster address (the leader's MMA thread waits on ITS barriers only)
  auto lbar = [&](int i) { return (NCTA == 2) ? ptx::mapa(bar(i), 0u) : bar(i); };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NXA; ++i) { ptx::mbar_init(bar(B_XAFULL + i), 1); ptx::mbar_init(bar(B_XAEMPTY + i), 1); }
    for (int i = 0; i < NXB; ++i) { ptx::mbar_init(bar(B_XBFULL + i), 1); ptx::mbar_init(bar(B_OUTRDY + i), EPI_THREADS); }
    for (int i = 0; i < NWA; ++i) { ptx::mbar_init(bar(B_WAFULL + i), 1); ptx::mbar_init(bar(B_WAEMPTY + i), 1); }
    for (int i = 0; i < NWB; ++i) { ptx::mbar_init(bar(B_WBFULL + i), 1); ptx::mbar_init(bar(B_WBEMPTY + i), 1); }
    ptx::mbar_init(bar(B_APFULL), 1);
    ptx::mbar_init(bar(B_ZQFULL), EPI_THREADS * NCTA);    // the epilogue threads of BOTH CTAs arrive at the leader's copy
    ptx::mbar_init(bar(B_UTFULL), 1);
    ptx::mbar_init(bar(B_UTEMPTY), EPI_THREADS * NCTA);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); }
  if (warp == 2 && lane == 0) { ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_out); }
  if (warp == 3 && lane == 0) {
    ptx::prefetch_tmap(&tm_wu); ptx::prefetch_tmap(&tm_gu);
    if (C::REM) { ptx::prefetch_tmap(&tm_wu_t); ptx::prefetch_tmap(&tm_gu_t); }
  }
  if (warp == 0 && lane == 1) { ptx::prefetch_tmap(&tm_wd); ptx::prefetch_tmap(&tm_gd); }
  if (warp == 1) {
    if (NCTA == 2) ptx::tmem_alloc_cg2(tmem_slot, TMEM_COLS); else ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  }
  {  // biases -> fp32 in shared memory, pre-multiplied so that the epilogues fold them into FMAs they issue anyway:
     // bd / gbd (zero padded to 128), alpha*bu, 0.5*gbu
    float* sbd = reinterpret_cast<float*>(smem_gen + C::OFF_BD);
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) {
      const int j = i & 127;
      float v = 0.f;
      if (i < 128) { if (j < p.r) v = __bfloat162float(p.bd[j]); }
      else if (GATED) { if (j < p.rg) v = __bfloat162float(p.gbd[j]); }
      sbd[i] = v;
    }
    float* sbu = reinterpret_cast<float*>(smem_gen + C::OFF_BU);
    for (int i = threadIdx.x; i < p.d; i += NUM_THREADS) {
      sbu[i] = p.alpha * __bfloat162float(p.bu[i]);
      sbu[p.d + i] = GATED ? 0.5f * __bfloat162float(p.gbu[i]) : 0.f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================== TMA producer: phase-A x chunks (HBM stream) =====================================
    // (every role loop is run by ONE elected thread, see the MMA warp)
    if (ptx::elect_one()) {
      ItemCursor it; it.init(p, nkc, num_items, NCTA);
      uint32_t n = 0, na = 0;
      for (; it.valid(); it.next()) {
        const int row0 = (int)((it.tile * NCTA + rank) * TILE_M);
        for (int k = 0; k < nkc; ++k, ++n) {
          const uint32_t sl = n % NXA;
          ptx::mbar_wait_dbg(bar(B_XAEMPTY + sl), ((n / NXA) & 1) ^ 1, p.dbg, __LINE__);
          const uint32_t dst = smem_base + C::OFF_XA + sl * (2 * XCH_BYTES);
          // (pairs: the leader expects the bytes of both CTAs; every load signals the LEADER's barrier)
          if (rank == 0) ptx::mbar_arrive_expect_tx(bar(B_XAFULL + sl), NCTA * (GATED ? 2 : 1) * XCH_BYTES);
          // first touch of data that phase B reads again one tile later: ask L2 to keep it
          if (NCTA == 2) {
            if (GATED) ptx::tma_load_2d_hint_cg2(dst, &tm_x1, k * CH, row0, lbar(B_XAFULL + sl), ptx::L2_EVICT_LAST);
            ptx::tma_load_2d_hint_cg2(dst + XCH_BYTES, &tm_x2, k * CH, row0, lbar(B_XAFULL + sl), ptx::L2_EVICT_LAST);
          } else {
            if (GATED) ptx::tma_load_2d_hint(dst, &tm_x1, k * CH, row0, bar(B_XAFULL + sl), ptx::L2_EVICT_LAST);
            ptx::tma_load_2d_hint(dst + XCH_BYTES, &tm_x2, k * CH, row0, bar(B_XAFULL + sl), ptx::L2_EVICT_LAST);
          }
          // the weight chunks the same MMA step consumes: Wd_k, Gd_k (always L2 hits)
#pragma unroll
          for (int m = 0; m < (GATED ? 2 : 1); ++m, ++na) {
            const uint32_t sw = na % NWA;
            ptx::mbar_wait_dbg(bar(B_WAEMPTY + sw), ((na / NWA) & 1) ^ 1, p.dbg, __LINE__);
            if (rank == 0) ptx::mbar_arrive_expect_tx(bar(B_WAFULL + sw), NCTA * C::WA_BYTES);
            if (NCTA == 2)   // this CTA's half of the rows
              ptx::tma_load_2d_hint_cg2(smem_base + C::OFF_WA + sw * C::WA_BYTES, m ? &tm_gd : &tm_wd, k * CH, (int)rank * C::WA_ROWS,
                                        lbar(B_WAFULL + sw), ptx::L2_EVICT_LAST);
            else
              ptx::tma_load_2d_hint(smem_base + C::OFF_WA + sw * C::WA_BYTES, m ? &tm_gd : &tm_wd, k * CH, 0, bar(B_WAFULL + sw),
                                    ptx::L2_EVICT_LAST);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA producer: phase-B weights Wu_c, Gu_c (always L2 hits) ====================
    if (ptx::elect_one()) {
      uint32_t nb = 0;
      ItemCursor it; it.init(p, nkc, num_items, NCTA);
      for (; it.valid(); it.next()) {
        for (int c = it.cb; c < it.ce; ++c) {
          const int wrow = c * CH + (int)rank * C::WB_ROWS;   // first row of Wu / Gu this CTA stages
#pragma unroll
          for (int m = 0; m < (GATED ? 2 : 1); ++m, ++nb) {
            const uint32_t sl = nb % NWB;
            ptx::mbar_wait_dbg(bar(B_WBEMPTY + sl), ((nb / NWB) & 1) ^ 1, p.dbg, __LINE__);
            if (rank == 0) ptx::mbar_arrive_expect_tx(bar(B_WBFULL + sl), NCTA * C::WB_BYTES);
            const uint32_t wdst = smem_base + C::OFF_WB + sl * C::WB_BYTES;
#pragma unroll
            for (int kb = 0; kb < C::KBF; ++kb) {
              if (NCTA == 2) ptx::tma_load_2d_hint_cg2(wdst + kb * C::WB_BLK, m ? &tm_gu : &tm_wu, kb * CH, wrow, lbar(B_WBFULL + sl), ptx::L2_EVICT_LAST);
              else ptx::tma_load_2d_hint(wdst + kb * C::WB_BLK, m ? &tm_gu : &tm_wu, kb * CH, wrow, bar(B_WBFULL + sl), ptx::L2_EVICT_LAST);
            }
            if (C::REM) {
              if (NCTA == 2) ptx::tma_load_2d_hint_cg2(wdst + C::KBF * C::WB_BLK, m ? &tm_gu_t : &tm_wu_t, C::KBF * CH, wrow, lbar(B_WBFULL + sl), ptx::L2_EVICT_LAST);
              else ptx::tma_load_2d_hint(wdst + C::KBF * C::WB_BLK, m ? &tm_gu_t : &tm_wu_t, C::KBF * CH, wrow, bar(B_WBFULL + sl), ptx::L2_EVICT_LAST);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // ONE thread chosen by elect.sync runs the whole loop (the CUTLASS idiom).  Operands of tcgen05.mma / tcgen05.commit /
    // cp.async.bulk.tensor live in uniform registers: under `if (lane == 0)` ptxas wrapped EVERY issue in an elect /
    // R2UR.BROADCAST / branch sequence (~15 instructions, > 100 ns per MMA measured with tools/trace_k1.py -- the whole
    // kernel was issue-bound); under an elect predicate it keeps the descriptors in uniform registers.
    if (rank == 0 && ptx::elect_one()) {   // pairs: the leader CTA issues for both SMs
      constexpr uint32_t IDESC_A = ptx::umma_idesc_bf16(128 * NCTA, R);
      constexpr uint32_t IDESC_B = ptx::umma_idesc_bf16(128 * NCTA, CH);
      auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if (NCTA == 2) ptx::umma_bf16_ss_cg2(d, a, b, idesc, acc); else ptx::umma_bf16_ss(d, a, b, idesc, acc);
      };
      auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if (NCTA == 2) ptx::umma_bf16_ts_cg2(d, a, b, idesc, acc); else ptx::umma_bf16_ts(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t b) {   // pairs: arrives on the barrier at this offset in BOTH CTAs
        if (NCTA == 2) ptx::umma_commit_cg2(b); else ptx::umma_commit(b);
      };
      uint32_t nxa = 0, na = 0, nb = 0, ui = 0, ti = 0;
      // Phase A of the next item and phase B of this item are issued in whatever order their inputs become ready
      // (non-blocking barrier tests): the HBM stream is not paced by the epilogue and vice versa.
      auto ready_a = [&]() -> bool {
        bool r = ptx::mbar_test(bar(B_XAFULL + nxa % NXA), (nxa / NXA) & 1) && ptx::mbar_test(bar(B_WAFULL + na % NWA), (na / NWA) & 1);
        if (GATED) r = r && ptx::mbar_test(bar(B_WAFULL + (na + 1) % NWA), ((na + 1) / NWA) & 1);
        return r;
      };
      auto ready_b = [&]() -> bool {
        bool r = ptx::mbar_test(bar(B_UTEMPTY), (ui & 1) ^ 1) && ptx::mbar_test(bar(B_WBFULL + nb % NWB), (nb / NWB) & 1);
        if (GATED) r = r && ptx::mbar_test(bar(B_WBFULL + (nb + 1) % NWB), ((nb + 1) / NWB) & 1);
        return r;
      };
      // one phase-A step: A += x2_k Wd_k^T ; P += x1_k Gd_k^T
      auto step_a = [&](int k) {
        const uint32_t sx = nxa % NXA;
        ptx::mbar_wait_dbg(bar(B_XAFULL + sx), (nxa / NXA) & 1, p.dbg, __LINE__);
        const uint32_t x1s = smem_base + C::OFF_XA + sx * (2 * XCH_BYTES), x2s = x1s + XCH_BYTES;
#pragma unroll
        for (int m = 0; m < (GATED ? 2 : 1); ++m, ++na) {
          const uint32_t sl = na % NWA;
          ptx::mbar_wait_dbg(bar(B_WAFULL + sl), (na / NWA) & 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t ws = smem_base + C::OFF_WA + sl * C::WA_BYTES;
          const uint64_t adesc = ptx::umma_desc_kmajor_sw128(m ? x1s : x2s), bdesc = ptx::umma_desc_kmajor_sw128(ws);
#pragma unroll
          for (int ks = 0; ks < CH / 16; ++ks)   // +32 bytes per K step = +2 in the descriptor's 16-byte address units
            mma_ss(tmem_base + (m ? TM_P : TM_A), adesc + 2 * ks, bdesc + 2 * ks, IDESC_A, (k > 0 || ks > 0) ? 1u : 0u);
          commit(bar(B_WAEMPTY + sl));
          if (m == (GATED ? 1 : 0)) commit(bar(B_XAEMPTY + sx));
        }
        ++nxa;
      };
      // one phase-B chunk: U = z Wu_c^T ; T = q Gu_c^T  (z, q: A operand from TMEM)
      auto chunk_b = [&]() {
        ptx::mbar_wait_dbg(bar(B_UTEMPTY), (ui & 1) ^ 1, p.dbg, __LINE__);
#pragma unroll
        for (int m = 0; m < (GATED ? 2 : 1); ++m, ++nb) {
          const uint32_t sl = nb % NWB;
          ptx::mbar_wait_dbg(bar(B_WBFULL + sl), (nb / NWB) & 1, p.dbg, __LINE__);
          ptx::tc_fence_after();
          const uint32_t ws = smem_base + C::OFF_WB + sl * C::WB_BYTES;
          const uint64_t d128 = ptx::umma_desc_kmajor_sw128(ws), d64 = ptx::umma_desc_kmajor_sw64(ws + C::KBF * C::WB_BLK);
          const uint32_t ta = tmem_base + (m ? TM_Q : TM_Z), td = tmem_base + TM_UT + (m ? CH : 0);
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks) {
            const int kb = ks / 4, kin = ks % 4;
            const uint64_t bdesc = (kb < C::KBF) ? d128 + (uint64_t)(kb * (C::WB_BLK / 16) + 2 * kin) : d64 + (uint64_t)(2 * kin);
            mma_ts(td, ta + 8 * ks, bdesc, IDESC_B, ks > 0);
          }
          commit(bar(B_WBEMPTY + sl));
          if (m == (GATED ? 1 : 0)) commit(bar(B_UTFULL));
        }
        ++ui;
      };
      ItemCursor it; it.init(p, nkc, num_items, NCTA);
      if (it.valid()) {
        for (int k = 0; k < nkc; ++k) { step_a(k); if (k < 12) VLPET_TR(100 + k); }
        commit(bar(B_APFULL));
      }
      for (; it.valid(); it.next(), ++ti) {
        const int nB = it.ce - it.cb;
        const int nA = (it.w + it.stride < num_items) ? nkc : 0;
        // z/q of this item are in TMEM, and A/P have been drained: phase B of this item and phase A of the next may run
        ptx::mbar_wait_dbg(bar(B_ZQFULL), ti & 1, p.dbg, __LINE__);
        ptx::tc_fence_after();
        int kb = 0, ka = 0;
        uint32_t spins = 0;
        while (kb < nB || ka < nA) {
          bool did = false;
          if (kb < nB && ready_b()) { chunk_b(); if (ti == 0 && kb < 12) VLPET_TR(64 + 2 * kb); ++kb; did = true; }
          if (ka < nA && ready_a()) {
            step_a(ka);
            if (ti == 0 && ka < 12) VLPET_TR(65 + 2 * ka);
            if (++ka == nA) commit(bar(B_APFULL));
            did = true;
          }
          if (did) spins = 0;
          else if (++spins > (1u << 24)) __trap();
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== phase-B x manager: TMA load of x1_c / x2_c, TMA store of out_c ==============
    // The residual chunks of ALL work items form one stream of entries; entry e lives in XB slot e % NXB.  This thread
    // frees a slot itself (its store has read it), so it can refill it right away: no empty barrier.
    if (ptx::elect_one()) {
      ItemCursor ld; ld.init(p, nkc, num_items, NCTA);
      ItemCursor st = ld;
      int lc = ld.valid() ? ld.cb : 0, sc = lc;
      uint32_t e_st = 0;
      auto load = [&](uint32_t sl) {
        const uint32_t dst = smem_base + C::OFF_XB + sl * (2 * XCH_BYTES);
        const int row0 = (int)((ld.tile * NCTA + rank) * TILE_M);
        ptx::mbar_arrive_expect_tx(bar(B_XBFULL + sl), 2 * XCH_BYTES);
        ptx::tma_load_2d_hint(dst, &tm_x1, lc * CH, row0, bar(B_XBFULL + sl), ptx::L2_EVICT_FIRST);   // last use
        ptx::tma_load_2d_hint(dst + XCH_BYTES, &tm_x2, lc * CH, row0, bar(B_XBFULL + sl), ptx::L2_EVICT_FIRST);
        if (++lc == ld.ce) { ld.next(); lc = ld.cb; }
      };
      for (uint32_t i = 0; i < NXB && ld.valid(); ++i) load(i);
      for (; st.valid(); ++e_st) {
        const uint32_t sl = e_st % NXB;
        ptx::mbar_wait_dbg(bar(B_OUTRDY + sl), (e_st / NXB) & 1, p.dbg, __LINE__);
        if (e_st < 12) VLPET_TR(128 + 2 * e_st);
        ptx::tma_store_2d_hint(&tm_out, smem_base + C::OFF_XB + sl * (2 * XCH_BYTES), sc * CH, (int)((st.tile * NCTA + rank) * TILE_M),
                               ptx::L2_EVICT_FIRST);   // never re-read by this kernel
        ptx::tma_store_commit();
        ptx::tma_store_wait_read0();
        if (e_st < 12) VLPET_TR(129 + 2 * e_st);
        if (++sc == st.ce) { st.next(); sc = st.cb; }
        if (ld.valid()) load(sl);
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
    uint32_t e = 0, ui = 0, ti = 0;
    const uint32_t zqfull = lbar(B_ZQFULL), utempty = lbar(B_UTEMPTY);
    ItemCursor it; it.init(p, nkc, num_items, NCTA);
    for (; it.valid(); it.next(), ++ti) {
      const int64_t tile = it.tile * NCTA + rank;
      const int cb = it.cb, ce = it.ce;
      // ---- epilogue 1: z = gelu_new(A + bd) (cg 0,1) / q = gelu_new(P + gbd) (cg 2,3) -> packed bf16 into TMEM
      VLPET_TRACE(0);
      ptx::mbar_wait_dbg(bar(B_APFULL), ti & 1, p.dbg, __LINE__);
      VLPET_TRACE(1);
      ptx::tc_fence_after();
      const int branch = cg >> 1;
      if (GATED || branch == 0) {
        const uint32_t tsrc = lane_addr + (branch ? TM_P : TM_A);
        const uint32_t tdst = lane_addr + (branch ? TM_Q : TM_Z);
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
          for (int q = 0; q < 4; ++q) lds_f2x2(sbias + (uint32_t)(j0 + q * 4) * 4u, bias[2 * q], bias[2 * q + 1]);
          ptx::tmem_ld_wait();
          uint32_t o[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o[q] = pack2(gelu_new2(add2(mk2u(v[q * 2], v[q * 2 + 1]), bias[q])));
          ptx::tmem_st_32x32b_x8(tdst + j0 / 2, o);
        }
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      if (NCTA == 2 && rank != 0) ptx::mbar_arrive_cluster(zqfull); else ptx::mbar_arrive(bar(B_ZQFULL));   // z/q in place AND A/P drained
      VLPET_TRACE(2);
      // ---- epilogue 2, per 64-column chunk: this warp owns columns [cg*16, cg*16+16) of the chunk
      for (int c = cb; c < ce; ++c, ++ui, ++e) {
        const uint32_t sl = e % NXB;
        const int col0 = c * CH + cg * 16;  // first of this thread's 16 output columns
        ptx::mbar_wait_dbg(bar(B_UTFULL), ui & 1, p.dbg, __LINE__);
        VLPET_TRACE(3 + 4 * (c - cb));
        ptx::tc_fence_after();
        uint32_t u[16], t[16];
        const uint32_t tU = lane_addr + TM_UT + cg * 16;
        ptx::tmem_ld_32x32b_x16(tU, u);
        if (GATED) ptx::tmem_ld_32x32b_x16(tU + CH, t);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        if (NCTA == 2 && rank != 0) ptx::mbar_arrive_cluster(utempty); else ptx::mbar_arrive(bar(B_UTEMPTY));  // accumulators are in registers
        VLPET_TRACE(4 + 4 * (c - cb));
        ptx::mbar_wait_dbg(bar(B_XBFULL + sl), (e / NXB) & 1, p.dbg, __LINE__);
        VLPET_TRACE(5 + 4 * (c - cb));
        const uint32_t x1row = smem_base + C::OFF_XB + sl * (2 * XCH_BYTES) + (uint32_t)row * 128u;
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
        ptx::mbar_arrive(bar(B_OUTRDY + sl));
        VLPET_TRACE(6 + 4 * (c - cb));
      }
      if (p.trace && threadIdx.x == 128 && ti < 12) p.trace[blockIdx.x * 256 + 52 + ti] = gtimer();   // end of work item ti
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();   // no CTA of a pair exits (or frees TMEM) while its peer can still signal it
  if (warp == 1) {
    ptx::tc_fence_after();
    if (NCTA == 2) ptx::tmem_dealloc_cg2(tmem_base, TMEM_COLS); else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
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
DevInfo dev_info() {   // of the CURRENT device
  const int sms = device_sm_count();
  return DevInfo{sms != 0 ? 1 : 0, sms > 0 ? sms : 0, sms > 0 ? 10 : 0};
}

template <int R, bool GATED, int NCTA>
int launch(const VlpetK1Desc& D, const CUtensorMap* maps, Params p, cudaStream_t st) {
  using C = Cfg<R, NCTA>;
  const int smem = C::smem_bytes(D.d);
  auto kern = k1_fwd_sm100_kernel<R, GATED, NCTA>;
  static int attr_set[64] = {0};
  VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kern), attr_set, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  int slots = dev_info().sms;      // concurrently resident work units: CTAs, or CTA pairs
  if (NCTA == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs_dev[64];  // a GPC with an odd number of SMs strands one SM: ask the driver how many pairs fit at once
    static bool mp_init = false;
    if (!mp_init) { for (int i = 0; i < 64; ++i) max_pairs_dev[i] = -1; mp_init = true; }
    int dev_ = 0;
    cudaGetDevice(&dev_);
    int& max_pairs = max_pairs_dev[(dev_ >= 0 && dev_ < 64) ? dev_ : 0];
    if (max_pairs < 0) {
      cfg.gridDim = dim3((unsigned)(dev_info().sms / 2 * 2));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 0; }
      max_pairs = n;
    }
    if (max_pairs < 1) return fail(VLPET_E_UNSUPPORTED, "k1_fwd(fused): no room for a CTA pair");
    slots = max_pairs;
  }
  {  // Whole waves run one unit (tile / tile pair) per slot.  The units of the last, partial wave (all of them when M is
     // small) are split over several slots each (largest divisor of nkc that still fits the wave): the tail costs one
     // phase A plus nkc/nsplit chunks instead of a full tile time.
    const int nkc = D.d / CH;
    const int64_t tiles = (D.M + TILE_M - 1) / TILE_M;
    p.num_units = (tiles + NCTA - 1) / NCTA;
    const int64_t rem = p.num_units % slots;
    p.full_tiles = p.num_units - rem;
    p.nsplit = 1;
    for (int ns = 2; ns <= nkc; ++ns)
      if (nkc % ns == 0 && rem * ns <= slots) p.nsplit = ns;
    if (p.nsplit == 1) p.full_tiles = p.num_units;
  }
  const int64_t items = p.full_tiles + (p.num_units - p.full_tiles) * p.nsplit;
  cfg.gridDim = dim3((unsigned)(NCTA * (items < slots ? items : slots)));
  VLPET_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], maps[8], p));
  count_launch();
  return 0;
}

template <int R>
int launch_r(const VlpetK1Desc& D, const CUtensorMap* maps, const Params& p, bool gated, int ncta, cudaStream_t st) {
  if (ncta == 2) return gated ? launch<R, true, 2>(D, maps, p, st) : launch<R, false, 2>(D, maps, p, st);
  return gated ? launch<R, true, 1>(D, maps, p, st) : launch<R, false, 1>(D, maps, p, st);
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

// CTA pairs pay off when there are several tiles per SM (weights are the largest share of the shared-memory traffic:
// 128 vs 132 us at M = 96 000); with less than two waves the single-CTA kernel's finer tail split wins.
// g_pairs_mode: -1 auto, 0 never, 1 whenever there are at least two tiles (vlpet_debug_set_k1_pairs / VLPET_K1_PAIRS).
int g_pairs_mode = []() { const char* e = getenv("VLPET_K1_PAIRS"); return e ? atoi(e) : -1; }();
bool use_pairs(const VlpetK1Desc& D, int R) {
  if (R == 128 || g_pairs_mode == 0) return false;
  const int64_t tiles = (D.M + TILE_M - 1) / TILE_M;
  if (g_pairs_mode == 1) return tiles >= 2;
  return tiles >= 2 * (int64_t)dev_info().sms;
}

}  // namespace

int set_k1_trace(unsigned long long* dev_buf) {
  g_trace = dev_buf;
  return 0;
}
int set_k1_pairs(int mode) {
  g_pairs_mode = mode;
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
  if (!aligned16(w.Wd) || !aligned16(w.Wu) || !aligned16(w.bu) ||
      (gated && (!aligned16(w.Gd) || !aligned16(w.Gu) || !aligned16(w.gbu))))
    return fail(VLPET_E_ALIGN, "k1_fwd(fused): weights must be 16-byte aligned");
  const int R = pick_R(D);
  const int ncta = use_pairs(D, R) ? 2 : 1;
  const uint32_t wa_rows = (uint32_t)(R / ncta), wb_rows = (uint32_t)(CH / ncta);   // rows of a weight chunk one CTA stages
  CUtensorMap maps[9];
  VLPET_TRY(make_map(&maps[0], x1, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[1], x2, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[2], out, (uint64_t)D.M, (uint64_t)D.d, TILE_M, false));
  VLPET_TRY(make_map(&maps[3], w.Wd, (uint64_t)D.r, (uint64_t)D.d, wa_rows, true));
  VLPET_TRY(make_map(&maps[4], gated ? w.Gd : w.Wd, (uint64_t)(gated ? D.rg : D.r), (uint64_t)D.d, wa_rows, true));
  VLPET_TRY(make_map(&maps[5], w.Wu, (uint64_t)D.d, (uint64_t)D.r, wb_rows, true));
  VLPET_TRY(make_map(&maps[6], gated ? w.Gu : w.Wu, (uint64_t)D.d, (uint64_t)(gated ? D.rg : D.r), wb_rows, true));
  // 32-column boxes (64-byte swizzle) for the last, half-wide K block of the phase-B weight chunks (R % 64 == 32)
  VLPET_TRY(make_map_bf16(&maps[7], w.Wu, (uint64_t)D.d, (uint64_t)D.r, (uint64_t)D.r, wb_rows, 32, true));
  VLPET_TRY(make_map_bf16(&maps[8], gated ? w.Gu : w.Wu, (uint64_t)D.d, (uint64_t)(gated ? D.rg : D.r),
                          (uint64_t)(gated ? D.rg : D.r), wb_rows, 32, true));
  Params p;
  p.M = D.M; p.d = D.d; p.r = D.r; p.rg = gated ? D.rg : D.r; p.add_gate = D.add_gate;
  p.s = D.s; p.alpha = D.alpha; p.kappa = D.kappa;
  p.bd = static_cast<const __nv_bfloat16*>(w.bd); p.bu = static_cast<const __nv_bfloat16*>(w.bu);
  p.gbd = static_cast<const __nv_bfloat16*>(gated ? w.gbd : w.bd); p.gbu = static_cast<const __nv_bfloat16*>(gated ? w.gbu : w.bu);
  p.seed = D.seed;
  p.seed_dev = D.seed_dev;
  p.trace = g_trace;
  p.dbg = trap_buffer_dev();
  p.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  p.inv_keep = p.thr16 ? 1.0f / (1.0f - (float)p.thr16 / 65536.0f) : 1.0f;
  switch (R) {
    case 32: return launch_r<32>(D, maps, p, gated, ncta, st);
    case 64: return launch_r<64>(D, maps, p, gated, ncta, st);
    case 96: return launch_r<96>(D, maps, p, gated, ncta, st);
    case 128: return launch_r<128>(D, maps, p, gated, 1, st);
  }
  return fail(VLPET_E_UNSUPPORTED, "k1_fwd(fused): unsupported rank");
}

}  // namespace vlpet
