// Token-contracted weight-gradient GEMM, sm_100a (tcgen05 + TMA), used by the fused K1 / K2 backward:
//
//     D_p[c, n] += scale_p * sum_tok A_p[tok, c] * B_p[tok, n]        p = 0 .. npairs-1   (one launch)
//
// A_p is a [Mtok, d] bf16 activation-like tensor (x1, x2, du, dt), B_p a skinny [Mtok, nb] bf16 tensor (z, q, da, dp,
// optionally followed by a column of ones so that the extra output column is sum_tok A_p = the bias gradient).
// These are the dW GEMMs of SURVEY Appendix A: dWu = du^T z, dGu = dt^T q, dWd = da^T x2, dGd = dp^T x1 (the last two
// are produced transposed, D = x^T da, and scattered into the [r, d] layout on the way out).
//
// Why a separate kernel (DESIGN.md "K1-bwd"): the four dW accumulators are 4 x [768 x 96] fp32 = 1.18 MB, far more
// than one SM can keep on chip (TMEM is 256 KB), and they contract over ALL tokens -- so a token-tile-major kernel
// would have to flush 1.18 MB of partials per 128-token tile (885 MB of L2 reductions at M = 96 000).  Here the
// accumulators are STATIONARY instead: a CTA owns (pair, 128 output rows) in TMEM ([128 x N] fp32) for the whole
// launch, streams 64-token slices of A and B through a 6-stage TMA ring, and flushes once at the end.
//
// Both operands are consumed MN-major straight from the row-major global layout ([tok][c] / [tok][n], contraction
// index = row), i.e. no transposed copies exist anywhere.
#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {
namespace {

constexpr int KT = 64;                    // tokens per ring stage
constexpr int SLAB = 128;                 // output rows (d columns) per accumulator = MMA M
constexpr int SPC = 1;                    // slabs per CTA: one -- the final flush (fp32 reductions into the gradient buffers) is
                                          // the per-launch floor and scales with the rows a CTA owns (3 slabs: 31 us at any M)
constexpr int BOX_BYTES = KT * 64 * 2;    // one [64 tok x 64 cols] bf16 box = 8 KB
constexpr int A_STAGE = SPC * 2 * BOX_BYTES;   // 48 KB
constexpr int B_STAGE = 2 * BOX_BYTES;         // 16 KB
constexpr int STAGE = A_STAGE + B_STAGE;       // 64 KB
constexpr int NST = 6;
constexpr int NACC = 4;                   // accumulators per slab, used round-robin by the steps and summed at the flush: the
                                          // fp32 accumulation chain of one TMEM accumulator stays 4x shorter (rounding)
constexpr int SMEM_BYTES = NST * STAGE + 256 + 1024;
constexpr int THREADS = 192;

struct alignas(64) WgradMaps {
  CUtensorMap a[4];
  CUtensorMap b[4];
};
struct WgradArgs {
  float* out[4];        // [d, nout] (transposed == 0, row pitch ldo floats) or [nout, d] (transposed == 1), fp32, accumulated into
  int64_t ldo[4];       // row pitch of a non-transposed output (a column slice of a wider matrix: rank halves, vlpet_wide.cu)
  float* bias[4];       // [d] or nullptr: receives column `nout` of D (the ones-column trick)
  float scale[4];
  int transposed[4];
  int64_t Mtok;
  int d, nout, NB;      // NB = MMA N (multiple of 16, >= nout + (bias ? 1 : 0))
  int ncolgroups;       // ceil(d / (SPC*SLAB))
  uint32_t* dbg;        // host-mapped trap record (ptx::mbar_wait_dbg)
};

__global__ void __launch_bounds__(THREADS, 1)
wgrad_sm100_kernel(const __grid_constant__ WgradMaps maps, const WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + NST * STAGE;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };   // FULL[NST], EMPTY[NST], DONE
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST + 1);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int pair = blockIdx.x / p.ncolgroups, cg = blockIdx.x % p.ncolgroups;
  const int nslab_total = p.d / SLAB;
  const int slab0 = cg * SPC;
  const int nslab = (nslab_total - slab0) < SPC ? (nslab_total - slab0) : SPC;
  const int64_t nsteps = (p.Mtok + KT - 1) / KT;
  const int64_t my_steps = (nsteps > blockIdx.y) ? (nsteps - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(bar(i), 1); ptx::mbar_init(bar(NST + i), 1); }
    ptx::mbar_init(bar(2 * NST), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&maps.a[pair]); ptx::prefetch_tmap(&maps.b[pair]); }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 128 * SPC * NACC);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (ptx::elect_one()) {
      for (int64_t i = 0; i < my_steps; ++i) {
        const uint32_t s = (uint32_t)(i % NST);
        const int tok0 = (int)((blockIdx.y + i * gridDim.y) * KT);
        ptx::mbar_wait_dbg(bar(NST + s), (uint32_t)((i / NST) & 1) ^ 1, p.dbg, __LINE__);
        const uint32_t dst = smem_base + s * STAGE;
        ptx::mbar_arrive_expect_tx(bar(s), (uint32_t)(nslab * 2 * BOX_BYTES + B_STAGE));
        for (int sl = 0; sl < nslab; ++sl) {
          ptx::tma_load_2d(dst + (sl * 2) * BOX_BYTES, &maps.a[pair], (slab0 + sl) * SLAB, tok0, bar(s));
          ptx::tma_load_2d(dst + (sl * 2 + 1) * BOX_BYTES, &maps.a[pair], (slab0 + sl) * SLAB + 64, tok0, bar(s));
        }
        ptx::tma_load_2d(dst + A_STAGE, &maps.b[pair], 0, tok0, bar(s));
        ptx::tma_load_2d(dst + A_STAGE + BOX_BYTES, &maps.b[pair], 64, tok0, bar(s));
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_bf16_m128_major((uint32_t)p.NB, 1u, 1u);
      for (int64_t i = 0; i < my_steps; ++i) {
        const uint32_t s = (uint32_t)(i % NST);
        ptx::mbar_wait_dbg(bar(s), (uint32_t)((i / NST) & 1), p.dbg, __LINE__);
        ptx::tc_fence_after();
        const uint32_t a0 = smem_base + s * STAGE, b0 = a0 + A_STAGE;
        for (int sl = 0; sl < nslab; ++sl) {
#pragma unroll
          for (int ks = 0; ks < KT / 16; ++ks) {
            ptx::umma_bf16_ss(tmem_base + (uint32_t)((sl * NACC + (int)(i % NACC)) * 128),
                              ptx::umma_desc_mnmajor_sw128(a0 + (sl * 2) * BOX_BYTES + ks * 2048, BOX_BYTES),
                              ptx::umma_desc_mnmajor_sw128(b0 + ks * 2048, BOX_BYTES), idesc, (i >= NACC || ks > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(bar(NST + s));
      }
      ptx::umma_commit(bar(2 * NST));
    }
  } else if (my_steps > 0) {
    // ---- final flush: TMEM -> (transpose through shared memory) -> coalesced fp32 vector reductions
    // A thread owns one accumulator row c, but the gradient buffers want contiguous runs: [c][n] rows of nout floats, or
    // for the transposed outputs (dWd, dGd) [n][c] rows of d floats.  The ring memory is free once every MMA has
    // completed, so each slab is staged there in the output's orientation and then written with red.global.add.v4.f32,
    // 512 contiguous bytes per warp instruction (scalar atomics took ~19 us per CTA for the transposed outputs).
    const int quarter = warp % 4;
    const int ew = warp - 2;                 // 0..3
    const int row = quarter * 32 + lane;
    ptx::mbar_wait_dbg(bar(2 * NST), 0, p.dbg, __LINE__);
    ptx::tc_fence_after();
    float* out = p.out[pair];
    const int64_t ldo = p.ldo[pair];
    float* bias = p.bias[pair];
    const float sc = p.scale[pair];
    const int transposed = p.transposed[pair];
    float* stage = reinterpret_cast<float*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)));   // [128][STG] or [NB][128]
    constexpr int STG = 132;                 // row pitch (floats) of the non-transposed staging: 16-byte aligned, conflict-light
    for (int sl = 0; sl < nslab; ++sl) {
      const int c0 = (slab0 + sl) * SLAB;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * NACC * 128);
      const int nacc = my_steps < NACC ? (int)my_steps : NACC;   // accumulators that were written at all
      for (int n0 = 0; n0 < p.NB; n0 += 16) {
        // all round-robin accumulators of this column block are requested before ONE wait: at small M (one or two steps per
        // CTA) the flush is most of the kernel, and four serial tcgen05.ld round trips per block were most of the flush
        uint32_t v[16], w[NACC - 1][16];
        ptx::tmem_ld_32x32b_x16(taddr + n0, v);
#pragma unroll
        for (int a = 1; a < NACC; ++a)
          if (a < nacc) ptx::tmem_ld_32x32b_x16(taddr + a * 128 + n0, w[a - 1]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int a = 1; a < NACC; ++a) {
          if (a < nacc) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(w[a - 1][e]));
          }
        }
        if (transposed) {
#pragma unroll
          for (int e = 0; e < 16; ++e) stage[(n0 + e) * SLAB + row] = sc * __uint_as_float(v[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(stage + row * STG + n0 + e) =
                make_float4(sc * __uint_as_float(v[e]), sc * __uint_as_float(v[e + 1]), sc * __uint_as_float(v[e + 2]),
                            sc * __uint_as_float(v[e + 3]));
        }
        if (bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (n0 + e == p.nout) atomicAdd(bias + c0 + row, sc * __uint_as_float(v[e]));
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (transposed) {
        // out[n][c0 .. c0+127]: one warp per n, lane l -> 4 consecutive c
        for (int n = ew; n < p.nout; n += 4) {
          const float4 q = *reinterpret_cast<const float4*>(stage + n * SLAB + lane * 4);
          ptx::red_add_v4(out + (size_t)n * p.d + c0 + lane * 4, q.x, q.y, q.z, q.w);
        }
      } else if ((p.nout & 3) == 0) {
        // out[c][0 .. nout): one warp per row, lanes cover the row in float4 steps
        const int nv = p.nout / 4;
        for (int r = ew; r < SLAB; r += 4)
          for (int j = lane; j < nv; j += 32) {
            const float4 q = *reinterpret_cast<const float4*>(stage + r * STG + j * 4);
            ptx::red_add_v4(out + (size_t)(c0 + r) * ldo + j * 4, q.x, q.y, q.z, q.w);
          }
      } else {
        for (int r = ew; r < SLAB; r += 4)
          for (int j = lane; j < p.nout; j += 32) atomicAdd(out + (size_t)(c0 + r) * ldo + j, stage[r * STG + j]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128 * SPC * NACC);
  }
}

}  // namespace

int make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                  uint32_t box_cols, bool weight);

bool wgrad_sm100_supported(int d, int nout) {
  return d % SLAB == 0 && d >= SLAB && nout >= 8 && nout + 1 <= 128;
}

// pairs: A [Mtok, d] (row pitch lda elems), B [Mtok, nb_valid] (row pitch ldb elems; nb_valid = nout or nout + 1 with the
// ones column), out / bias fp32 accumulated into.
int wgrad_sm100(int npairs, const void* const* A, const int64_t* lda, const void* const* B, const int64_t* ldb,
                const int* nb_valid, float* const* out, float* const* bias, const float* scale, const int* transposed,
                int64_t Mtok, int d, int nout, int sm_count, cudaStream_t st, const int64_t* ldo) {
  if (npairs < 1 || npairs > 4) return fail(VLPET_E_BADARG, "wgrad: 1..4 pairs");
  if (!wgrad_sm100_supported(d, nout)) return fail(VLPET_E_UNSUPPORTED, "wgrad: d=%d nout=%d", d, nout);
  WgradMaps maps;
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  int nbmax = nout;
  for (int i = 0; i < npairs; ++i) {
    VLPET_TRY(make_map_bf16(&maps.a[i], A[i], (uint64_t)Mtok, (uint64_t)d, (uint64_t)lda[i], KT, 64, false));
    VLPET_TRY(make_map_bf16(&maps.b[i], B[i], (uint64_t)Mtok, (uint64_t)nb_valid[i], (uint64_t)ldb[i], KT, 64, false));
    a.out[i] = out[i]; a.bias[i] = bias[i]; a.scale[i] = scale[i]; a.transposed[i] = transposed[i];
    a.ldo[i] = (ldo && ldo[i] > 0) ? ldo[i] : nout;
    if (!transposed[i] && (a.ldo[i] & 3) != 0 && (nout & 3) == 0)
      return fail(VLPET_E_ALIGN, "wgrad: output pitch %lld must be a multiple of 4 floats", (long long)a.ldo[i]);
    if (nb_valid[i] > nbmax) nbmax = nb_valid[i];
  }
  for (int i = npairs; i < 4; ++i) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; }
  a.Mtok = Mtok; a.d = d; a.nout = nout;
  a.NB = (nbmax + 15) / 16 * 16;
  a.ncolgroups = (d / SLAB + SPC - 1) / SPC;
  a.dbg = trap_buffer_dev();
  static int attr_set[64] = {0};
  VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(wgrad_sm100_kernel), attr_set, SMEM_BYTES));
  // Token groups: one CTA per (pair, slab, token group); every group ends with a flush of its [128 x nout] partial sums.
  const int gx = npairs * a.ncolgroups;
  const int64_t nsteps = (Mtok + KT - 1) / KT;
  int gy = sm_count / gx;
  if (gy > nsteps) gy = (int)nsteps;
  if (gy < 1) gy = 1;
  wgrad_sm100_kernel<<<dim3(gx, gy), THREADS, SMEM_BYTES, st>>>(maps, a);
  VLPET_LAUNCH_OK();
  return 0;
}

}  // namespace vlpet
