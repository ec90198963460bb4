// Dense projection GEMM for the visual projection (K3), sm_100a:   C[M, N] (fp32) = A[M, K] W[N, K]^T + bias[N]
// (the nn.Linear(feat_dim, d_model) of VisualEmbedding.feat_embedding, src/modeling_bart.py:90, 157; bf16 operands,
// fp32 accumulate in TMEM).  The LayerNorm / order-embedding epilogue stays in the row kernel of vlpet_generic.cu,
// which needs the pre-norm projection in fp32 anyway (it is the tensor saved for the backward).
//
// Persistent CTAs walk 128 x 256 output tiles (n fastest, so the three n-tiles of one row block re-read A from L2).
// 4-stage TMA ring of 64-wide K chunks, one MMA-issuer thread, TMEM accumulators double-buffered (2 x 256 columns) so
// the epilogue of tile i overlaps the MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA and
// masked on store, so any M, N, K (K % 8 == 0 for the 16-byte row pitch) works.
#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {
int make_map_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                  uint32_t box_cols, bool weight);
namespace {

constexpr int BM = 128, BN = 256, BK = 64, NST = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = NST * STAGE + 256 + 1024;
constexpr int THREADS = 192;

struct GemmParams {
  int64_t M;
  int N, K;
  float* C;
  int64_t ldc;
  const __nv_bfloat16* bias;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + NST * STAGE;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };   // FULL[NST], EMPTY[NST], ACCFULL[2], ACCEMPTY[2]
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST + 4);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nk = (p.K + BK - 1) / BK;
  const int ntn = (p.N + BN - 1) / BN;
  const int64_t ntiles = ((p.M + BM - 1) / BM) * ntn;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(bar(i), 1); ptx::mbar_init(bar(NST + i), 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar(2 * NST + i), 1); ptx::mbar_init(bar(2 * NST + 2 + i), 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tm_a); ptx::prefetch_tmap(&tm_w); }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (int)(tile / ntn) * BM, n0 = (int)(tile % ntn) * BN;
        for (int k = 0; k < nk; ++k, ++it) {
          const uint32_t s = it % NST;
          ptx::mbar_wait(bar(NST + s), ((it / NST) & 1) ^ 1);
          const uint32_t dst = smem_base + s * STAGE;
          ptx::mbar_arrive_expect_tx(bar(s), STAGE);
          ptx::tma_load_2d(dst, &tm_a, k * BK, m0, bar(s));
          ptx::tma_load_2d(dst + A_BYTES, &tm_w, k * BK, n0, bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t IDESC = ptx::umma_idesc_bf16_m128(BN);
      uint32_t it = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const uint32_t b = ti & 1;
        ptx::mbar_wait(bar(2 * NST + 2 + b), ((ti >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        for (int k = 0; k < nk; ++k, ++it) {
          const uint32_t s = it % NST;
          ptx::mbar_wait(bar(s), (it / NST) & 1);
          ptx::tc_fence_after();
          const uint32_t as = smem_base + s * STAGE, ws = as + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)
            ptx::umma_bf16_ss(tmem_base + b * BN, ptx::umma_desc_kmajor_sw128(as + ks * 32),
                              ptx::umma_desc_kmajor_sw128(ws + ks * 32), IDESC, (k > 0 || ks > 0) ? 1u : 0u);
          ptx::umma_commit(bar(NST + s));
        }
        ptx::umma_commit(bar(2 * NST + b));
      }
    }
  } else {
    const int quarter = warp % 4;
    const int row = quarter * 32 + lane;
    uint32_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const uint32_t b = ti & 1;
      const int64_t m = (tile / ntn) * BM + row;
      const int n0 = (int)(tile % ntn) * BN;
      ptx::mbar_wait(bar(2 * NST + b), (ti >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + b * BN;
#pragma unroll 1
      for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + j0, v);
        ptx::tmem_ld_wait();
        if (m < p.M) {
          float* crow = p.C + m * p.ldc + n0 + j0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int n = n0 + j0 + g * 4;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = __uint_as_float(v[g * 4 + e]) + ((p.bias && n + e < p.N) ? __bfloat162float(p.bias[n + e]) : 0.f);
            if (n + 3 < p.N) {
              *reinterpret_cast<float4*>(crow + g * 4) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < p.N) crow[g * 4 + e] = o[e];
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(2 * NST + 2 + b));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool gemm_sm100_supported(int64_t M, int N, int K, int64_t ldc) {
  return M > 0 && N >= 8 && K >= 8 && K % 8 == 0 && ldc % 4 == 0 && device_sm_count() > 0;
}

// C[M,N] fp32 (row pitch ldc) = A[M,K] bf16 (pitch lda) * W[N,K]^T bf16 (pitch ldw) + bias[N] (bf16, may be null)
int gemm_sm100(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, float* C, int64_t ldc, int64_t M,
               int N, int K, cudaStream_t st) {
  if (!aligned16(A) || !aligned16(W) || !aligned16(C) || (lda & 7) || (ldw & 7))
    return fail(VLPET_E_ALIGN, "gemm: operands must be 16-byte aligned with pitches that are multiples of 8");
  CUtensorMap ta, tw;
  VLPET_TRY(make_map_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK, false));
  VLPET_TRY(make_map_bf16(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN, BK, true));
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = static_cast<const __nv_bfloat16*>(bias);
  static int attr_set[64] = {0};
  VLPET_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(gemm_sm100_kernel), attr_set, SMEM_BYTES));
  const int64_t ntiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int sms = device_sm_count();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  gemm_sm100_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ta, tw, p);
  VLPET_LAUNCH_OK();
  return 0;
}

}  // namespace vlpet
