// Shape-generic CUDA-core implementation of the PET hot path (any d, r, gate, fp32 or bf16 I/O).
//
// This is the correctness-first path (fp32 accumulate everywhere, exact tanhf/expf) that serves
//   * the fp32 configuration (BASELINE config 1; 1e-5 parity bar, which rules out bf16/tf32 tensor cores),
//   * shapes none of the bf16 paths covers (d not a multiple of 128 / 256, ranks above 192; under VLPET_IMPL_AUTO everything
//     else runs the tcgen05 kernels, the row-wise kernels of vlpet_rows.cu or the rank halves of vlpet_wide.cu), and
//   * the on-device cross-check of the fused kernels in tests (VLPET_IMPL_GENERIC).
// K3 (visual projection) keeps its row kernel and the one-pass reductions of its backward here; its GEMMs are tcgen05.
// It runs the op sequence of SURVEY Appendix A as a handful of launches over fp32 workspace intermediates:
// one strided tile GEMM (fp32 accumulate, fused bias / gelu_new / residual epilogue, split-K for the
// token-contracted weight gradients) plus small row-wise kernels for the gates.
//
// Reference lines restated: my_transformers/modeling_bart.py:1145-1155,1195-1231,1256-1260 (K1),
// adapters/adapter_controller.py:131-162 + adapters/adapter_modeling.py:55-61 (K2),
// src/modeling_bart.py:143-192 and src/modeling_t5.py:124-174 (K3).
#include "vlpet_common.cuh"

namespace vlpet {
namespace {

// ------------------------------------------------------------------------------------------------------------
// Strided tile GEMM:  C[m,n] (+)= alpha * (sum_k A(m,k) B(k,n) + bias[n]) -> act -> + addend_scale*addend[m,n]
// ------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const void* A = nullptr;
  int64_t sam = 0, sak = 0;
  int a_bf16 = 0;
  const void* B = nullptr;
  int64_t sbk = 0, sbn = 0;
  int b_bf16 = 0;
  void* C = nullptr;
  int64_t ldc = 0;
  int c_bf16 = 0;
  int64_t M = 0;
  int N = 0;
  int64_t K = 0;
  float alpha = 1.f;
  const void* bias = nullptr;
  int bias_bf16 = 0;
  float* pre = nullptr;  // [M,N] fp32, value before the activation
  int act = 0;           // 1 = gelu_new
  const void* addend = nullptr;
  int addend_bf16 = 0;
  float addend_scale = 0.f;
  int accumulate = 0;  // fp32 C only
  int splitk = 1;
};

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int64_t kchunk = (g.K + g.splitk - 1) / g.splitk;
  kchunk = (kchunk + BK - 1) / BK * BK;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = (kbeg + kchunk < g.K) ? (kbeg + kchunk) : g.K;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (g.sak == 1) { mm = idx / BK; kk = idx % BK; } else { kk = idx / BM; mm = idx % BM; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = ld_as_float(g.A, gm * g.sam + gk * g.sak, g.a_bf16);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int nn, kk;
      if (g.sbn == 1) { kk = idx / BN; nn = idx % BN; } else { nn = idx / BK; kk = idx % BK; }
      int64_t gk = k0 + kk;
      int gn = n0 + nn;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = ld_as_float(g.B, gk * g.sbk + (int64_t)gn * g.sbn, g.b_bf16);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tm + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tn + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + tm + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tn + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += ld_as_float(g.bias, gn, g.bias_bf16);
      v *= g.alpha;
      if (g.pre) g.pre[gm * g.N + gn] = v;
      if (g.act == 1) v = gelu_new_f(v);
      int64_t ci = gm * g.ldc + gn;
      if (g.addend) v += g.addend_scale * ld_as_float(g.addend, ci, g.addend_bf16);
      if (g.accumulate) {
        float* c = static_cast<float*>(g.C) + ci;
        if (g.splitk > 1) atomicAdd(c, v); else *c += v;
      } else {
        st_from_float(g.C, ci, v, g.c_bf16);
      }
    }
  }
}

int launch_gemm(GemmArgs g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.K <= 0) g.K = 0;
  if (g.splitk < 1) g.splitk = 1;
  if (g.splitk > 1 && (!g.accumulate || g.bias || g.act || g.addend || g.pre || g.c_bf16))
    return fail(VLPET_E_BADARG, "internal: split-K GEMM needs a pure fp32 accumulate epilogue");
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN), (unsigned)g.splitk);
  gemm_kernel<<<grid, 256, 0, st>>>(g);
  VLPET_LAUNCH_OK();
  return 0;
}

// Y[M,N] = act(alpha*(X[M,K] W[N,K]^T + b))      (nn.Linear)
GemmArgs linear_nt(const void* X, int x_bf16, const void* W, const void* b, int w_bf16, void* Y, int y_bf16,
                   int64_t M, int N, int K) {
  GemmArgs g;
  g.A = X; g.sam = K; g.sak = 1; g.a_bf16 = x_bf16;
  g.B = W; g.sbk = 1; g.sbn = K; g.b_bf16 = w_bf16;
  g.C = Y; g.ldc = N; g.c_bf16 = y_bf16;
  g.M = M; g.N = N; g.K = K;
  g.bias = b; g.bias_bf16 = w_bf16;
  return g;
}
// Y[M,K] = X[M,N] W[N,K]                          (grad wrt the Linear's input)
GemmArgs linear_nn(const void* X, int x_bf16, const void* W, int w_bf16, void* Y, int y_bf16, int64_t M, int N,
                   int K) {
  GemmArgs g;
  g.A = X; g.sam = N; g.sak = 1; g.a_bf16 = x_bf16;
  g.B = W; g.sbk = K; g.sbn = 1; g.b_bf16 = w_bf16;
  g.C = Y; g.ldc = K; g.c_bf16 = y_bf16;
  g.M = M; g.N = K; g.K = N;
  return g;
}
// dW[N1,N2] += alpha * sum_t A[t,N1] B[t,N2]      (weight gradient, contraction over tokens, split-K)
int launch_wgrad(const void* A, int a_bf16, int N1, const void* B, int b_bf16, int N2, int64_t tokens, float* dW,
                 float alpha, cudaStream_t st) {
  if (!dW) return 0;
  GemmArgs g;
  g.A = A; g.sam = 1; g.sak = N1; g.a_bf16 = a_bf16;
  g.B = B; g.sbk = N2; g.sbn = 1; g.b_bf16 = b_bf16;
  g.C = dW; g.ldc = N2; g.c_bf16 = 0;
  g.M = N1; g.N = N2; g.K = tokens;
  g.alpha = alpha;
  g.accumulate = 1;
  int64_t tiles = ((N1 + BM - 1) / BM) * (int64_t)((N2 + BN - 1) / BN);
  int64_t want = (1184 + tiles - 1) / tiles;            // ~8 CTAs per SM
  int64_t maxsplit = (tokens + 4 * BK - 1) / (4 * BK);  // at least 64 tokens per split
  int64_t sk = want < maxsplit ? want : maxsplit;
  if (sk < 1) sk = 1;
  if (sk > 65535) sk = 65535;
  g.splitk = (int)sk;
  return launch_gemm(g, st);
}

// ------------------------------------------------------------------------------------------------------------
// Small row-wise / elementwise kernels
// ------------------------------------------------------------------------------------------------------------
// out[c] += scale * sum_m A[m,c] * (B ? B[m,c] : 1) * (ids ? ids[m]==match : 1)
struct Drop {  // dropout stream of one K1 call (thr16 == 0: identity)
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float inv_keep = 1.f;
  const uint64_t* seed_dev = nullptr;
  __device__ __forceinline__ uint64_t eff() const { return seed + ((thr16 && seed_dev) ? *seed_dev : 0ull); }
};
inline Drop make_drop(const VlpetK1Desc& D) {
  Drop x;
  x.seed = D.seed;
  x.seed_dev = D.seed_dev;
  x.thr16 = D.p_drop > 0.f ? drop_thr16(D.p_drop) : 0u;
  x.inv_keep = x.thr16 ? 1.0f / (1.0f - (float)x.thr16 / 65536.0f) : 1.0f;
  return x;
}

__global__ void colsum_kernel(const void* A, int a_bf16, const void* B, int b_bf16, const int64_t* ids,
                              int64_t match, int64_t M, int C, int rows_per_block, float scale, float* out, Drop dr) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int64_t m0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  for (int64_t m = m0; m < m1; ++m) {
    if (ids && ids[m] != match) continue;
    float v = ld_as_float(A, m * C + c, a_bf16);
    if (B) v *= ld_as_float(B, m * C + c, b_bf16);
    acc += v * drop_scale(dr.eff(), dr.thr16, dr.inv_keep, m * C + c);
  }
  atomicAdd(out + c, acc * scale);
}
int launch_colsum(const void* A, int a_bf16, const void* B, int b_bf16, const int64_t* ids, int64_t match, int64_t M,
                  int C, float scale, float* out, cudaStream_t st, Drop dr = Drop()) {
  if (!out || M <= 0 || C <= 0) return 0;
  int threads = C >= 128 ? 128 : 32;
  int gx = (C + threads - 1) / threads;
  int64_t gy = (1184 + gx - 1) / gx;
  int64_t rpb = (M + gy - 1) / gy;
  if (rpb < 16) rpb = 16;
  gy = (M + rpb - 1) / rpb;
  colsum_kernel<<<dim3(gx, (unsigned)gy), threads, 0, st>>>(A, a_bf16, B, b_bf16, ids, match, M, C, (int)rpb, scale,
                                                            out, dr);
  VLPET_LAUNCH_OK();
  return 0;
}

// U <- kappa*x2 + alpha*U   (y1, modeling_bart.py:1155 / modeling_t5.py:789-795)
__global__ void y1_kernel(const void* x2, int bf16, float* U, int64_t n, float kappa, float alpha) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) U[i] = kappa * ld_as_float(x2, i, bf16) + alpha * U[i];
}
// rowgate[m] = sigmoid(sum_c x1[m,c] w1[c] + y1[m,c] w2[c] + b)   (middle_x: w1=w2=gw; small: gw[:d], gw[d:])
__global__ void rowgate_kernel(const void* x1, int bf16, const float* Y1, const void* w1, const void* w2,
                               const void* b, int64_t M, int d, float* rowgate) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= M) return;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32)
    acc += ld_as_float(x1, row * d + c, bf16) * ld_as_float(w1, c, bf16) + Y1[row * d + c] * ld_as_float(w2, c, bf16);
  acc = warp_sum(acc);
  if (lane == 0) rowgate[row] = sigmoid_f(acc + ld_as_float(b, 0, bf16));
}
// per-sample mean over L (small gate, modeling_bart.py:1214): Gb[b] = mean_l sg[b*L+l]
__global__ void sample_mean_kernel(const float* sg, int64_t B, int L, float* Gb) {
  int64_t b = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (b >= B) return;
  float acc = 0.f;
  for (int l = lane; l < L; l += 32) acc += sg[b * L + l];
  acc = warp_sum(acc);
  if (lane == 0) Gb[b] = acc / (float)L;
}
// out = x1 + s * gate(y1)
__global__ void k1_out_kernel(const void* x1, int bf16, const float* Y1, const float* Tg, const float* rowG, int L,
                              const void* gz, int gate, int add_gate, float s, int64_t M, int d, void* out, Drop dr) {
  int64_t n = M * d;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t m = i / d;
    int c = (int)(i - m * d);
    float y1 = Y1[i], h;
    if (gate == VLPET_GATE_LARGE) {
      float G = sigmoid_f(Tg[i]);
      h = add_gate ? y1 + G : y1 * G;
    } else if (gate == VLPET_GATE_MIDDLE_X) {
      float G = rowG[m];
      h = add_gate ? y1 + G : y1 * G;
    } else if (gate == VLPET_GATE_SMALL) {
      float G = rowG[m / L];
      h = add_gate ? y1 + G : y1 * G;
    } else if (gate == VLPET_GATE_MIDDLE_Y) {
      float z = ld_as_float(gz, c, bf16);
      h = add_gate ? y1 + 1.0f + z : y1 + y1 * z;
    } else {
      h = y1;
    }
    st_from_float(out, i, ld_as_float(x1, i, bf16) + s * h * drop_scale(dr.eff(), dr.thr16, dr.inv_keep, i), bf16);
  }
}
// large / none / middle_y backward, elementwise, in place:  Y1 <- dy1,  Tg <- dt (large);  dx1 <- dout (non-large)
__global__ void k1_bwd_elem_kernel(const void* dout, int bf16, float* Y1, float* Tg, const void* gz, int gate,
                                   int add_gate, float s, int64_t M, int d, void* dx1, Drop dr) {
  int64_t n = M * d;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float go = ld_as_float(dout, i, bf16);
    float dh = s * go * drop_scale(dr.eff(), dr.thr16, dr.inv_keep, i);
    if (gate == VLPET_GATE_LARGE) {
      float G = sigmoid_f(Tg[i]);
      float y1 = Y1[i];
      float dy1 = add_gate ? dh : dh * G;
      float dG = add_gate ? dh : dh * y1;
      Y1[i] = dy1;
      Tg[i] = dG * G * (1.0f - G);
    } else if (gate == VLPET_GATE_MIDDLE_Y) {
      int c = (int)(i % d);
      Y1[i] = add_gate ? dh : dh * (1.0f + ld_as_float(gz, c, bf16));
      st_from_float(dx1, i, go, bf16);
    } else {
      Y1[i] = dh;
      st_from_float(dx1, i, go, bf16);
    }
  }
}
// rs[m] = sum_c dh[m,c] * (add ? 1 : y1[m,c])
__global__ void rowsum_dh_kernel(const void* dout, int bf16, const float* Y1, int add_gate, float s, int64_t M, int d,
                                 float* rs, Drop dr) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= M) return;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) {
    float dh = s * ld_as_float(dout, row * d + c, bf16) * drop_scale(dr.eff(), dr.thr16, dr.inv_keep, row * d + c);
    acc += add_gate ? dh : dh * Y1[row * d + c];
  }
  acc = warp_sum(acc);
  if (lane == 0) rs[row] = acc;
}
// middle_x: dtrow[m] = rs[m] G(1-G), G = rowgate[m]
// small   : dtrow[m] = (sum_l rs[b,l]) / L * sg(1-sg), sg = rowgate[m]     (one warp per sample)
__global__ void rowgate_bwd_kernel(const float* rs, const float* rowgate, int gate, int64_t M, int L, float* dtrow) {
  if (gate == VLPET_GATE_MIDDLE_X) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
      float G = rowgate[i];
      dtrow[i] = rs[i] * G * (1.0f - G);
    }
  } else {
    int64_t b = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (b >= M / L) return;
    float acc = 0.f;
    for (int l = lane; l < L; l += 32) acc += rs[b * L + l];
    acc = warp_sum(acc) / (float)L;
    for (int l = lane; l < L; l += 32) {
      float sg = rowgate[b * L + l];
      dtrow[b * L + l] = acc * sg * (1.0f - sg);
    }
  }
}
// row-gate backward, elementwise:  Y1 <- dy1 = (add ? dh : dh*G) + dtrow[m] w2[c];  dx1 = dout + dtrow[m] w1[c]
__global__ void k1_bwd_rowgate_elem_kernel(const void* dout, int bf16, float* Y1, const float* G, int L_or_1,
                                           const float* dtrow, const void* w1, const void* w2, int add_gate, float s,
                                           int64_t M, int d, void* dx1, Drop dr) {
  int64_t n = M * d;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t m = i / d;
    int c = (int)(i - m * d);
    float go = ld_as_float(dout, i, bf16);
    float dh = s * go * drop_scale(dr.eff(), dr.thr16, dr.inv_keep, i);
    float g = G[m / L_or_1];
    float dt = dtrow[m];
    Y1[i] = (add_gate ? dh : dh * g) + dt * ld_as_float(w2, c, bf16);
    st_from_float(dx1, i, go + dt * ld_as_float(w1, c, bf16), bf16);
  }
}
// D <- D * gelu_new'(pre)
__global__ void mul_gelu_grad_kernel(float* D, const float* pre, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) D[i] *= gelu_new_grad_f(pre[i]);
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}
inline unsigned warp_rows_blocks(int64_t rows) { return (unsigned)((rows + 7) / 8); }  // 256 threads = 8 rows

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------------------
static bool k1_row_gate(int gate) { return gate == VLPET_GATE_MIDDLE_X || gate == VLPET_GATE_SMALL; }

size_t generic_k1_fwd_ws(const VlpetK1Desc& D) {
  Arena a(nullptr, 0);
  a.take<float>(D.M * D.r);
  a.take<float>(D.M * D.d);
  if (D.gate == VLPET_GATE_LARGE) {
    a.take<float>(D.M * D.rg);
    a.take<float>(D.M * D.d);
  }
  if (k1_row_gate(D.gate)) {
    a.take<float>(D.M);
    a.take<float>(D.gate == VLPET_GATE_SMALL ? D.M / D.L : 1);
  }
  return a.off;
}

int generic_k1_fwd(const VlpetK1Desc& D, const void* x1, const void* x2, const VlpetK1Params& w, void* out, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  const int d = D.d, r = D.r, rg = D.rg;
  Arena a(ws, ws_bytes);
  float* Z = a.take<float>(M * r);
  float* U = a.take<float>(M * d);
  float *Q = nullptr, *Tg = nullptr, *rowgate = nullptr, *Gb = nullptr;
  if (D.gate == VLPET_GATE_LARGE) {
    Q = a.take<float>(M * rg);
    Tg = a.take<float>(M * d);
  }
  if (k1_row_gate(D.gate)) {
    rowgate = a.take<float>(M);
    Gb = a.take<float>(D.gate == VLPET_GATE_SMALL ? M / D.L : 1);
  }
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k1_fwd: workspace %zu < %zu bytes", ws_bytes, a.off);

  GemmArgs g = linear_nt(x2, bf, w.Wd, w.bd, bf, Z, 0, M, r, d);
  g.act = 1;
  VLPET_TRY(launch_gemm(g, st));
  VLPET_TRY(launch_gemm(linear_nt(Z, 0, w.Wu, w.bu, bf, U, 0, M, d, r), st));
  y1_kernel<<<ew_blocks(M * d), 256, 0, st>>>(x2, bf, U, M * d, D.kappa, D.alpha);
  VLPET_LAUNCH_OK();
  const float* rowG = nullptr;
  if (D.gate == VLPET_GATE_LARGE) {
    g = linear_nt(x1, bf, w.Gd, w.gbd, bf, Q, 0, M, rg, d);
    g.act = 1;
    VLPET_TRY(launch_gemm(g, st));
    VLPET_TRY(launch_gemm(linear_nt(Q, 0, w.Gu, w.gbu, bf, Tg, 0, M, d, rg), st));
  } else if (k1_row_gate(D.gate)) {
    const char* gw = static_cast<const char*>(w.gw);
    const void* w2 = D.gate == VLPET_GATE_SMALL ? gw + (size_t)d * esize(D.dtype) : gw;
    rowgate_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(x1, bf, U, w.gw, w2, w.gb, M, d, rowgate);
    VLPET_LAUNCH_OK();
    rowG = rowgate;
    if (D.gate == VLPET_GATE_SMALL) {
      sample_mean_kernel<<<warp_rows_blocks(M / D.L), 256, 0, st>>>(rowgate, M / D.L, D.L, Gb);
      VLPET_LAUNCH_OK();
      rowG = Gb;
    }
  }
  k1_out_kernel<<<ew_blocks(M * d), 256, 0, st>>>(x1, bf, U, Tg, rowG, D.L > 0 ? D.L : 1, w.gz, D.gate, D.add_gate,
                                                 D.s, M, d, out, make_drop(D));
  VLPET_LAUNCH_OK();
  return 0;
}

size_t generic_k1_bwd_ws(const VlpetK1Desc& D) {
  Arena a(nullptr, 0);
  a.take<float>(D.M * D.r);  // Apre
  a.take<float>(D.M * D.r);  // Z
  a.take<float>(D.M * D.d);  // U / Y1 / DY1
  a.take<float>(D.M * D.r);  // DZ
  if (D.gate == VLPET_GATE_LARGE) {
    a.take<float>(D.M * D.rg);  // Ppre
    a.take<float>(D.M * D.rg);  // Q
    a.take<float>(D.M * D.d);   // T / DT
    a.take<float>(D.M * D.rg);  // DQ
  }
  if (k1_row_gate(D.gate)) {
    a.take<float>(D.M);
    a.take<float>(D.M);
    a.take<float>(D.M);
    a.take<float>(D.gate == VLPET_GATE_SMALL ? D.M / D.L : 1);
  }
  return a.off;
}

int generic_k1_bwd(const VlpetK1Desc& D, const void* x1, const void* x2, const void* dout, const VlpetK1Params& w,
                   void* dx1, void* dx2, const VlpetK1Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  const int d = D.d, r = D.r, rg = D.rg;
  Arena a(ws, ws_bytes);
  float* Apre = a.take<float>(M * r);
  float* Z = a.take<float>(M * r);
  float* Y1 = a.take<float>(M * d);
  float* DZ = a.take<float>(M * r);
  float *Ppre = nullptr, *Q = nullptr, *Tg = nullptr, *DQ = nullptr, *rowgate = nullptr, *rs = nullptr,
        *dtrow = nullptr, *Gb = nullptr;
  if (D.gate == VLPET_GATE_LARGE) {
    Ppre = a.take<float>(M * rg);
    Q = a.take<float>(M * rg);
    Tg = a.take<float>(M * d);
    DQ = a.take<float>(M * rg);
  }
  if (k1_row_gate(D.gate)) {
    rowgate = a.take<float>(M);
    rs = a.take<float>(M);
    dtrow = a.take<float>(M);
    Gb = a.take<float>(D.gate == VLPET_GATE_SMALL ? M / D.L : 1);
  }
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k1_bwd: workspace %zu < %zu bytes", ws_bytes, a.off);

  // ---- recompute the forward intermediates
  GemmArgs g = linear_nt(x2, bf, w.Wd, w.bd, bf, Z, 0, M, r, d);
  g.act = 1;
  g.pre = Apre;
  VLPET_TRY(launch_gemm(g, st));
  VLPET_TRY(launch_gemm(linear_nt(Z, 0, w.Wu, w.bu, bf, Y1, 0, M, d, r), st));
  y1_kernel<<<ew_blocks(M * d), 256, 0, st>>>(x2, bf, Y1, M * d, D.kappa, D.alpha);
  VLPET_LAUNCH_OK();

  if (D.gate == VLPET_GATE_LARGE) {
    g = linear_nt(x1, bf, w.Gd, w.gbd, bf, Q, 0, M, rg, d);
    g.act = 1;
    g.pre = Ppre;
    VLPET_TRY(launch_gemm(g, st));
    VLPET_TRY(launch_gemm(linear_nt(Q, 0, w.Gu, w.gbu, bf, Tg, 0, M, d, rg), st));
    k1_bwd_elem_kernel<<<ew_blocks(M * d), 256, 0, st>>>(dout, bf, Y1, Tg, nullptr, D.gate, D.add_gate, D.s, M, d,
                                                        nullptr, make_drop(D));
    VLPET_LAUNCH_OK();
  } else if (k1_row_gate(D.gate)) {
    const bool small = D.gate == VLPET_GATE_SMALL;
    const char* gw = static_cast<const char*>(w.gw);
    const void* w2 = small ? gw + (size_t)d * esize(D.dtype) : gw;
    rowgate_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(x1, bf, Y1, w.gw, w2, w.gb, M, d, rowgate);
    VLPET_LAUNCH_OK();
    const float* Gval = rowgate;
    if (small) {
      sample_mean_kernel<<<warp_rows_blocks(M / D.L), 256, 0, st>>>(rowgate, M / D.L, D.L, Gb);
      VLPET_LAUNCH_OK();
      Gval = Gb;
    }
    rowsum_dh_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(dout, bf, Y1, D.add_gate, D.s, M, d, rs, make_drop(D));
    VLPET_LAUNCH_OK();
    if (small)
      rowgate_bwd_kernel<<<warp_rows_blocks(M / D.L), 256, 0, st>>>(rs, rowgate, D.gate, M, D.L, dtrow);
    else
      rowgate_bwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(rs, rowgate, D.gate, M, 1, dtrow);
    VLPET_LAUNCH_OK();
    // gate parameter grads (need y1, so before Y1 is overwritten): dgw = dtrow^T [x1 (;|+) y1], dgb = sum dtrow
    VLPET_TRY(launch_wgrad(dtrow, 0, 1, x1, bf, d, M, G.dgw, 1.f, st));
    VLPET_TRY(launch_wgrad(dtrow, 0, 1, Y1, 0, d, M, G.dgw ? G.dgw + (small ? d : 0) : nullptr, 1.f, st));
    VLPET_TRY(launch_colsum(dtrow, 0, nullptr, 0, nullptr, 0, M, 1, 1.f, G.dgb, st));
    k1_bwd_rowgate_elem_kernel<<<ew_blocks(M * d), 256, 0, st>>>(dout, bf, Y1, Gval, small ? D.L : 1, dtrow, w.gw, w2,
                                                                D.add_gate, D.s, M, d, dx1, make_drop(D));
    VLPET_LAUNCH_OK();
  } else {
    if (D.gate == VLPET_GATE_MIDDLE_Y) {  // dgz = sum_m dh * (add ? 1 : y1)
      VLPET_TRY(launch_colsum(dout, bf, D.add_gate ? nullptr : Y1, 0, nullptr, 0, M, d, D.s, G.dgz, st, make_drop(D)));
    }
    k1_bwd_elem_kernel<<<ew_blocks(M * d), 256, 0, st>>>(dout, bf, Y1, nullptr, w.gz, D.gate, D.add_gate, D.s, M, d,
                                                        dx1, make_drop(D));
    VLPET_LAUNCH_OK();
  }
  // ---- adapter branch: Y1 now holds dy1
  VLPET_TRY(launch_wgrad(Y1, 0, d, Z, 0, r, M, G.dWu, D.alpha, st));
  VLPET_TRY(launch_colsum(Y1, 0, nullptr, 0, nullptr, 0, M, d, D.alpha, G.dbu, st));
  g = linear_nn(Y1, 0, w.Wu, bf, DZ, 0, M, d, r);
  g.alpha = D.alpha;
  VLPET_TRY(launch_gemm(g, st));
  mul_gelu_grad_kernel<<<ew_blocks(M * r), 256, 0, st>>>(DZ, Apre, M * r);
  VLPET_LAUNCH_OK();
  VLPET_TRY(launch_wgrad(DZ, 0, r, x2, bf, d, M, G.dWd, 1.f, st));
  VLPET_TRY(launch_colsum(DZ, 0, nullptr, 0, nullptr, 0, M, r, 1.f, G.dbd, st));
  g = linear_nn(DZ, 0, w.Wd, bf, dx2, bf, M, r, d);
  g.addend = Y1;
  g.addend_bf16 = 0;
  g.addend_scale = D.kappa;
  VLPET_TRY(launch_gemm(g, st));
  // ---- gate branch (large): Tg holds dt
  if (D.gate == VLPET_GATE_LARGE) {
    VLPET_TRY(launch_wgrad(Tg, 0, d, Q, 0, rg, M, G.dGu, 1.f, st));
    VLPET_TRY(launch_colsum(Tg, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dgbu, st));
    VLPET_TRY(launch_gemm(linear_nn(Tg, 0, w.Gu, bf, DQ, 0, M, d, rg), st));
    mul_gelu_grad_kernel<<<ew_blocks(M * rg), 256, 0, st>>>(DQ, Ppre, M * rg);
    VLPET_LAUNCH_OK();
    VLPET_TRY(launch_wgrad(DQ, 0, rg, x1, bf, d, M, G.dGd, 1.f, st));
    VLPET_TRY(launch_colsum(DQ, 0, nullptr, 0, nullptr, 0, M, rg, 1.f, G.dgbd, st));
    g = linear_nn(DQ, 0, w.Gd, bf, dx1, bf, M, rg, d);
    g.addend = dout;
    g.addend_bf16 = bf;
    g.addend_scale = 1.f;
    VLPET_TRY(launch_gemm(g, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------------------
size_t generic_k2_fwd_ws(const VlpetK2Desc& D) { return align_up((size_t)D.M * D.r * 4, 256); }
size_t generic_k2_bwd_ws(const VlpetK2Desc& D) { return 3 * align_up((size_t)D.M * D.r * 4, 256); }

int generic_k2_fwd(const VlpetK2Desc& D, const void* kv, const void* y, const VlpetK2Params& w, void* out, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  Arena a(ws, ws_bytes);
  float* Z = a.take<float>(D.M * D.r);
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k2_fwd: workspace %zu < %zu bytes", ws_bytes, a.off);
  GemmArgs g = linear_nt(kv, bf, w.Wd, w.bd, bf, Z, 0, D.M, D.r, D.d);
  g.act = 1;
  VLPET_TRY(launch_gemm(g, st));
  g = linear_nt(Z, 0, w.Wu, w.bu, bf, out, bf, D.M, D.d, D.r);
  g.alpha = D.sf;
  g.addend = y;
  g.addend_bf16 = bf;
  g.addend_scale = 1.f;
  return launch_gemm(g, st);
}

int generic_k2_bwd(const VlpetK2Desc& D, const void* kv, const void* dout, const VlpetK2Params& w, void* dkv,
                   const VlpetK2Grads& G, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  const int d = D.d, r = D.r;
  Arena a(ws, ws_bytes);
  float* Apre = a.take<float>(M * r);
  float* Z = a.take<float>(M * r);
  float* DZ = a.take<float>(M * r);
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k2_bwd: workspace %zu < %zu bytes", ws_bytes, a.off);
  GemmArgs g = linear_nt(kv, bf, w.Wd, w.bd, bf, Z, 0, M, r, d);
  g.act = 1;
  g.pre = Apre;
  VLPET_TRY(launch_gemm(g, st));
  VLPET_TRY(launch_wgrad(dout, bf, d, Z, 0, r, M, G.dWu, D.sf, st));
  VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, nullptr, 0, M, d, D.sf, G.dbu, st));
  g = linear_nn(dout, bf, w.Wu, bf, DZ, 0, M, d, r);
  g.alpha = D.sf;
  VLPET_TRY(launch_gemm(g, st));
  mul_gelu_grad_kernel<<<ew_blocks(M * r), 256, 0, st>>>(DZ, Apre, M * r);
  VLPET_LAUNCH_OK();
  VLPET_TRY(launch_wgrad(DZ, 0, r, kv, bf, d, M, G.dWd, 1.f, st));
  VLPET_TRY(launch_colsum(DZ, 0, nullptr, 0, nullptr, 0, M, r, 1.f, G.dbd, st));
  if (dkv) VLPET_TRY(launch_gemm(linear_nn(DZ, 0, w.Wd, bf, dkv, bf, M, r, d), st));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// K3: visual projection
// ------------------------------------------------------------------------------------------------------------
namespace {
// one warp per visual token: out = norm(f) + norm([pos,area] Wp^T + bp) + E_img[img] + E_obj[V-1-obj]
// (src/modeling_bart.py:157-190).  In backward mode (dout != nullptr) it instead writes df, da (grads wrt the two
// pre-norm projections), xh_f, xh_a (normalised values) and pos5 for the weight-gradient GEMMs.
__global__ void visproj_row_kernel(const float* F, const void* pos, int bf16, const int64_t* img_ids,
                                   const int64_t* obj_ids, VlpetK3Params w, int64_t M, int N, int d, int V, int rms,
                                   float eps, void* out, const void* dout, float* dF, float* dA, float* XF, float* XA,
                                   float* P5, __nv_bfloat16* dFb) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= M) return;
  float p5[5];
#pragma unroll
  for (int i = 0; i < 4; ++i) p5[i] = ld_as_float(pos, row * 4 + i, bf16);
  p5[4] = (p5[3] - p5[2]) * (p5[1] - p5[0]);  // area = height * width, modeling_bart.py:129-141
  // pass 1: means, pass 2: centred second moments (rms: mean := 0)
  float sf = 0.f, sa = 0.f;
  for (int c = lane; c < d; c += 32) {
    float a = ld_as_float(w.bp, c, bf16);
#pragma unroll
    for (int i = 0; i < 5; ++i) a += p5[i] * ld_as_float(w.Wp, (int64_t)c * 5 + i, bf16);
    sf += F[row * d + c];
    sa += a;
  }
  float mf = rms ? 0.f : warp_sum(sf) / d, ma = rms ? 0.f : warp_sum(sa) / d;
  float sff = 0.f, saa = 0.f;
  for (int c = lane; c < d; c += 32) {
    float f = F[row * d + c] - mf;
    float a = ld_as_float(w.bp, c, bf16);
#pragma unroll
    for (int i = 0; i < 5; ++i) a += p5[i] * ld_as_float(w.Wp, (int64_t)c * 5 + i, bf16);
    a -= ma;
    sff += f * f;
    saa += a * a;
  }
  float rf = 1.0f / sqrtf(warp_sum(sff) / d + eps), ra = 1.0f / sqrtf(warp_sum(saa) / d + eps);
  if (dout == nullptr) {
    int64_t img = img_ids ? img_ids[row] : 0;
    int64_t obj = obj_ids ? obj_ids[row] : (row % N);
    int64_t orow = (int64_t)V - obj - 1;
    for (int c = lane; c < d; c += 32) {
      float f = F[row * d + c];
      float a = ld_as_float(w.bp, c, bf16);
#pragma unroll
      for (int i = 0; i < 5; ++i) a += p5[i] * ld_as_float(w.Wp, (int64_t)c * 5 + i, bf16);
      float v = (f - mf) * rf * ld_as_float(w.ln_f_w, c, bf16) + (a - ma) * ra * ld_as_float(w.ln_p_w, c, bf16);
      if (!rms) v += ld_as_float(w.ln_f_b, c, bf16) + ld_as_float(w.ln_p_b, c, bf16);
      v += ld_as_float(w.E_img, img * d + c, bf16) + ld_as_float(w.E_obj, orow * d + c, bf16);
      st_from_float(out, row * d + c, v, bf16);
    }
    return;
  }
  // backward of the two norms: dx = rstd * (g - mean(g) - xh * mean(g*xh)),  g = dout * weight   (rms: no mean(g))
  float gf = 0.f, gfx = 0.f, ga = 0.f, gax = 0.f;
  for (int c = lane; c < d; c += 32) {
    float f = F[row * d + c];
    float a = ld_as_float(w.bp, c, bf16);
#pragma unroll
    for (int i = 0; i < 5; ++i) a += p5[i] * ld_as_float(w.Wp, (int64_t)c * 5 + i, bf16);
    float go = ld_as_float(dout, row * d + c, bf16);
    float xf = (f - mf) * rf, xa = (a - ma) * ra;
    float g1 = go * ld_as_float(w.ln_f_w, c, bf16), g2 = go * ld_as_float(w.ln_p_w, c, bf16);
    gf += g1; gfx += g1 * xf; ga += g2; gax += g2 * xa;
  }
  gf = warp_sum(gf) / d; gfx = warp_sum(gfx) / d; ga = warp_sum(ga) / d; gax = warp_sum(gax) / d;
  if (rms) { gf = 0.f; ga = 0.f; }
  for (int c = lane; c < d; c += 32) {
    float f = F[row * d + c];
    float a = ld_as_float(w.bp, c, bf16);
#pragma unroll
    for (int i = 0; i < 5; ++i) a += p5[i] * ld_as_float(w.Wp, (int64_t)c * 5 + i, bf16);
    float go = ld_as_float(dout, row * d + c, bf16);
    float xf = (f - mf) * rf, xa = (a - ma) * ra;
    float g1 = go * ld_as_float(w.ln_f_w, c, bf16), g2 = go * ld_as_float(w.ln_p_w, c, bf16);
    const float dfv = rf * (g1 - gf - xf * gfx);
    dF[row * d + c] = dfv;
    if (dFb) dFb[row * d + c] = __float2bfloat16_rn(dfv);   // bf16 copy: operand of the tensor-core dWf GEMM
    dA[row * d + c] = ra * (g2 - ga - xa * gax);
    XF[row * d + c] = xf;
    XA[row * d + c] = xa;
  }
  if (lane < 5) P5[row * 5 + lane] = p5[lane];
}
}  // namespace

size_t generic_k3_fwd_ws(const VlpetK3Desc&) { return 0; }
// tensor-core paths of K3 (bf16 only): the feat projection GEMM (vlpet_gemm_sm100.cu) and dWf through the
// token-contracted weight-gradient GEMM (vlpet_wgrad_sm100.cu) in column blocks of `k3_wgrad_block(d)` outputs
static int k3_wgrad_block(int d) { return d % 96 == 0 ? 96 : (d % 64 == 0 ? 64 : 0); }
static bool k3_tc_fwd(const VlpetK3Desc& D) {
  return D.dtype == VLPET_BF16 && D.impl != VLPET_IMPL_GENERIC && gemm_sm100_supported(D.M, D.d, D.F, D.d);
}
static bool k3_tc_bwd(const VlpetK3Desc& D) {
  return D.dtype == VLPET_BF16 && D.impl != VLPET_IMPL_GENERIC && k3_wgrad_block(D.d) != 0 && D.F % 8 == 0 &&
         wgrad_sm100_supported(D.F, k3_wgrad_block(D.d)) && device_sm_count() > 0;
}
size_t generic_k3_bwd_ws(const VlpetK3Desc& D) {
  return 4 * align_up((size_t)D.M * D.d * 4, 256) + align_up((size_t)D.M * 5 * 4, 256) +
         (k3_tc_bwd(D) ? align_up((size_t)D.M * D.d * 2, 256) : 0);
}

int generic_k3_fwd(const VlpetK3Desc& D, const void* feats, const void* pos, const int64_t* img_ids,
                   const int64_t* obj_ids, const VlpetK3Params& w, void* out, float* save, void*, size_t,
                   cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  if (k3_tc_fwd(D) && aligned16(feats) && aligned16(w.Wf) && aligned16(save))
    VLPET_TRY(gemm_sm100(feats, D.F, w.Wf, D.F, w.bf, save, D.d, D.M, D.d, D.F, st));
  else
    VLPET_TRY(launch_gemm(linear_nt(feats, bf, w.Wf, w.bf, bf, save, 0, D.M, D.d, D.F), st));
  visproj_row_kernel<<<warp_rows_blocks(D.M), 256, 0, st>>>(save, pos, bf, img_ids, obj_ids, w, D.M, D.N, D.d, D.V,
                                                           D.rms, D.eps, out, nullptr, nullptr, nullptr, nullptr,
                                                           nullptr, nullptr, nullptr);
  VLPET_LAUNCH_OK();
  return 0;
}

// Every token-contracted [M, d] reduction of the K3 backward in ONE pass over dout, XF, XA, dF, dA (round 1 ran a column-sum or
// CUDA-core GEMM launch per output: nine launches, each a full pass):
//   dln_f_w = sum dout XF, dln_p_w = sum dout XA, dln_f_b = dln_p_b = sum dout, dbf = sum dF, dbp = sum dA,
//   dE_img[i] = sum_{img == i} dout, dWp[c, k] = sum dA[., c] P5[., k]   (k < 5: the box features of the position projection)
// A thread owns a column, a block a slab of rows; partial sums leave through one atomicAdd per output element and block.
struct K3SumArgs {
  const void* dout;
  int bf16, n_img, d;
  const float *XF, *XA, *dF, *dA, *P5;
  const int64_t* img_ids;
  int64_t M;
  int rows_per_block;
  float *dln_f_w, *dln_p_w, *dln_f_b, *dln_p_b, *dbf, *dbp, *dE_img, *dWp;
};
constexpr int K3S_MAX_IMG = 4;
__global__ void __launch_bounds__(128) k3_sums_kernel(const K3SumArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.d) return;
  int64_t m0 = (int64_t)blockIdx.y * a.rows_per_block, m1 = m0 + a.rows_per_block;
  if (m1 > a.M) m1 = a.M;
  float s_fw = 0.f, s_pw = 0.f, s_do = 0.f, s_df = 0.f, s_da = 0.f, s_img[K3S_MAX_IMG], s_wp[5];
#pragma unroll
  for (int i = 0; i < K3S_MAX_IMG; ++i) s_img[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) s_wp[k] = 0.f;
  for (int64_t m = m0; m < m1; ++m) {
    const int64_t i = m * a.d + c;
    const float go = ld_as_float(a.dout, i, a.bf16), da = a.dA[i];
    s_fw = fmaf(go, a.XF[i], s_fw);
    s_pw = fmaf(go, a.XA[i], s_pw);
    s_do += go;
    s_df += a.dF[i];
    s_da += da;
    const int img = a.img_ids ? (int)a.img_ids[m] : 0;
#pragma unroll
    for (int q = 0; q < K3S_MAX_IMG; ++q)
      if (img == q) s_img[q] += go;
#pragma unroll
    for (int k = 0; k < 5; ++k) s_wp[k] = fmaf(da, __ldg(a.P5 + m * 5 + k), s_wp[k]);
  }
  if (a.dln_f_w) atomicAdd(a.dln_f_w + c, s_fw);
  if (a.dln_p_w) atomicAdd(a.dln_p_w + c, s_pw);
  if (a.dln_f_b) atomicAdd(a.dln_f_b + c, s_do);
  if (a.dln_p_b) atomicAdd(a.dln_p_b + c, s_do);
  if (a.dbf) atomicAdd(a.dbf + c, s_df);
  if (a.dbp) atomicAdd(a.dbp + c, s_da);
  if (a.dE_img)
    for (int q = 0; q < a.n_img && q < K3S_MAX_IMG; ++q)
      if (s_img[q] != 0.f) atomicAdd(a.dE_img + (size_t)q * a.d + c, s_img[q]);
  if (a.dWp)
#pragma unroll
    for (int k = 0; k < 5; ++k) atomicAdd(a.dWp + (size_t)c * 5 + k, s_wp[k]);
}
int launch_k3_sums(K3SumArgs a, cudaStream_t st) {
  const int gx = (a.d + 127) / 128;
  int64_t gy = (1184 + gx - 1) / gx;
  int64_t rpb = (a.M + gy - 1) / gy;
  if (rpb < 16) rpb = 16;
  gy = (a.M + rpb - 1) / rpb;
  a.rows_per_block = (int)rpb;
  k3_sums_kernel<<<dim3(gx, (unsigned)gy), 128, 0, st>>>(a);
  VLPET_LAUNCH_OK();
  return 0;
}

int generic_k3_bwd(const VlpetK3Desc& D, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                   const VlpetK3Params& w, const float* save, void* dfeats, const VlpetK3Grads& G, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  const int d = D.d;
  Arena a(ws, ws_bytes);
  float* dF = a.take<float>(M * d);
  float* dA = a.take<float>(M * d);
  float* XF = a.take<float>(M * d);
  float* XA = a.take<float>(M * d);
  float* P5 = a.take<float>(M * 5);
  const bool tc = k3_tc_bwd(D) && aligned16(feats) && G.dWf && aligned16(G.dWf);
  __nv_bfloat16* dFb = tc ? a.take<__nv_bfloat16>(M * d) : nullptr;
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k3_bwd: workspace %zu < %zu bytes", ws_bytes, a.off);
  visproj_row_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(save, pos, bf, img_ids, nullptr, w, M, D.N, d, D.V, D.rms,
                                                         D.eps, nullptr, dout, dF, dA, XF, XA, P5, dFb);
  VLPET_LAUNCH_OK();
  const bool one_pass = D.n_img <= K3S_MAX_IMG;
  if (one_pass) {
    K3SumArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.dout = dout; sa.bf16 = bf; sa.n_img = (img_ids == nullptr) ? 1 : D.n_img; sa.d = d; sa.M = M;
    sa.XF = XF; sa.XA = XA; sa.dF = dF; sa.dA = dA; sa.P5 = P5; sa.img_ids = img_ids;
    sa.dln_f_w = G.dln_f_w; sa.dln_p_w = G.dln_p_w;
    sa.dln_f_b = D.rms ? nullptr : G.dln_f_b; sa.dln_p_b = D.rms ? nullptr : G.dln_p_b;
    sa.dbf = G.dbf; sa.dbp = G.dbp; sa.dE_img = G.dE_img; sa.dWp = G.dWp;
    VLPET_TRY(launch_k3_sums(sa, st));
  } else {
  VLPET_TRY(launch_colsum(dout, bf, XF, 0, nullptr, 0, M, d, 1.f, G.dln_f_w, st));
  VLPET_TRY(launch_colsum(dout, bf, XA, 0, nullptr, 0, M, d, 1.f, G.dln_p_w, st));
  if (!D.rms) {
    VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, nullptr, 0, M, d, 1.f, G.dln_f_b, st));
    VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, nullptr, 0, M, d, 1.f, G.dln_p_b, st));
  }
  }
  if (tc) {
    // dWf[n, f] = sum_tok dF[tok, n] feats[tok, f], produced transposed (rows = f) in column blocks of nb outputs
    const int nb = k3_wgrad_block(d), nblk = d / nb, sms = device_sm_count();
    for (int b0 = 0; b0 < nblk; b0 += 4) {
      const int np = (nblk - b0) < 4 ? (nblk - b0) : 4;
      const void* A[4]; const void* B[4]; int64_t lda[4], ldb[4]; int nbv[4], tr[4]; float* out[4]; float* bias[4]; float sc[4];
      for (int i = 0; i < np; ++i) {
        A[i] = feats; lda[i] = D.F;
        B[i] = dFb + (size_t)(b0 + i) * nb; ldb[i] = d; nbv[i] = nb; tr[i] = 1;
        out[i] = G.dWf + (size_t)(b0 + i) * nb * D.F; bias[i] = nullptr; sc[i] = 1.0f;
      }
      VLPET_TRY(wgrad_sm100(np, A, lda, B, ldb, nbv, out, bias, sc, tr, M, D.F, nb, sms, st));
    }
  } else {
    VLPET_TRY(launch_wgrad(dF, 0, d, feats, bf, D.F, M, G.dWf, 1.f, st));
  }
  if (!one_pass) {
  VLPET_TRY(launch_colsum(dF, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dbf, st));
  VLPET_TRY(launch_wgrad(dA, 0, d, P5, 0, 5, M, G.dWp, 1.f, st));
  VLPET_TRY(launch_colsum(dA, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dbp, st));
  }
  if (G.dE_img && !one_pass) {
    for (int i = 0; i < D.n_img; ++i) {
      if (img_ids == nullptr && i > 0) break;  // default ids are all 0
      VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, img_ids, i, M, d, 1.f, G.dE_img + (size_t)i * d, st));
    }
  }
  if (dfeats) VLPET_TRY(launch_gemm(linear_nn(dF, 0, w.Wf, bf, dfeats, bf, M, d, D.F), st));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// K3-LR: PET-shaped visual projector (src/modeling_bart.py:263-334)
// ------------------------------------------------------------------------------------------------------------
namespace {
// E <- U * G  |  U + U * G   with G = sigmoid(T)      (forward; E may alias U)
__global__ void lr_gate_fwd_kernel(const float* U, const float* T, float* E, int64_t n, int residual) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float g = sigmoid_f(T[i]), u = U[i];
    E[i] = residual ? u + u * g : u * g;
  }
}
// U <- dU = dE * (G | 1 + G),  T <- dT = dE * U * G (1 - G)      (backward, in place)
__global__ void lr_gate_bwd_kernel(const float* dE, float* U, float* T, int64_t n, int residual) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float g = sigmoid_f(T[i]), u = U[i], de = dE[i];
    U[i] = residual ? de * (1.0f + g) : de * g;
    T[i] = de * u * g * (1.0f - g);
  }
}
VlpetK3Params lr_row_params(const VlpetK3LRParams& w) {
  VlpetK3Params k;
  memset(&k, 0, sizeof(k));
  k.ln_f_w = w.ln_f_w; k.ln_f_b = w.ln_f_b; k.Wp = w.Wp; k.bp = w.bp; k.ln_p_w = w.ln_p_w; k.ln_p_b = w.ln_p_b;
  k.E_img = w.E_img; k.E_obj = w.E_obj;
  return k;
}
}  // namespace

size_t generic_k3lr_fwd_ws(const VlpetK3LRDesc& D) {
  Arena a(nullptr, 0);
  a.take<float>(D.M * D.r);
  if (D.gated) { a.take<float>(D.M * D.rg); a.take<float>(D.M * D.d); }
  return a.off;
}
size_t generic_k3lr_bwd_ws(const VlpetK3LRDesc& D) {
  Arena a(nullptr, 0);
  a.take<float>(D.M * D.r); a.take<float>(D.M * D.r); a.take<float>(D.M * D.d); a.take<float>(D.M * D.r);   // Apre Z U DZ
  if (D.gated) { a.take<float>(D.M * D.rg); a.take<float>(D.M * D.rg); a.take<float>(D.M * D.d); a.take<float>(D.M * D.rg); }
  for (int i = 0; i < 4; ++i) a.take<float>(D.M * D.d);   // dE dApos XF XA
  a.take<float>(D.M * 5);
  return a.off;
}

int generic_k3lr_fwd(const VlpetK3LRDesc& D, const void* feats, const void* pos, const int64_t* img_ids,
                     const int64_t* obj_ids, const VlpetK3LRParams& w, void* out, float* save, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  Arena a(ws, ws_bytes);
  float* Z = a.take<float>(M * D.r);
  float *Q = nullptr, *T = nullptr;
  if (D.gated) { Q = a.take<float>(M * D.rg); T = a.take<float>(M * D.d); }
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k3lr_fwd: workspace %zu < %zu bytes", ws_bytes, a.off);
  GemmArgs g = linear_nt(feats, bf, w.Wd, w.bd, bf, Z, 0, M, D.r, D.F);
  g.act = 1;
  VLPET_TRY(launch_gemm(g, st));
  VLPET_TRY(launch_gemm(linear_nt(Z, 0, w.Wu, w.bu, bf, save, 0, M, D.d, D.r), st));
  if (D.gated) {
    g = linear_nt(feats, bf, w.Gd, w.gbd, bf, Q, 0, M, D.rg, D.F);
    g.act = 1;
    VLPET_TRY(launch_gemm(g, st));
    VLPET_TRY(launch_gemm(linear_nt(Q, 0, w.Gu, w.gbu, bf, T, 0, M, D.d, D.rg), st));
    lr_gate_fwd_kernel<<<ew_blocks(M * D.d), 256, 0, st>>>(save, T, save, M * D.d, D.residual);
    VLPET_LAUNCH_OK();
  }
  visproj_row_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(save, pos, bf, img_ids, obj_ids, lr_row_params(w), M, D.N, D.d, D.V, 0,
                                                         D.eps, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                         nullptr);
  VLPET_LAUNCH_OK();
  return 0;
}

int generic_k3lr_bwd(const VlpetK3LRDesc& D, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                     const VlpetK3LRParams& w, const float* save, const VlpetK3LRGrads& G, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  const int bf = D.dtype == VLPET_BF16;
  const int64_t M = D.M;
  const int d = D.d, r = D.r, rg = D.rg;
  Arena a(ws, ws_bytes);
  float* Apre = a.take<float>(M * r);
  float* Z = a.take<float>(M * r);
  float* U = a.take<float>(M * d);
  float* DZ = a.take<float>(M * r);
  float *Ppre = nullptr, *Q = nullptr, *T = nullptr, *DQ = nullptr;
  if (D.gated) { Ppre = a.take<float>(M * rg); Q = a.take<float>(M * rg); T = a.take<float>(M * d); DQ = a.take<float>(M * rg); }
  float* dE = a.take<float>(M * d);
  float* dA = a.take<float>(M * d);
  float* XF = a.take<float>(M * d);
  float* XA = a.take<float>(M * d);
  float* P5 = a.take<float>(M * 5);
  if (!a.ok()) return fail(VLPET_E_WORKSPACE, "k3lr_bwd: workspace %zu < %zu bytes", ws_bytes, a.off);
  // LayerNorm / position / order-embedding part: identical to K3 with `save` = e
  visproj_row_kernel<<<warp_rows_blocks(M), 256, 0, st>>>(save, pos, bf, img_ids, nullptr, lr_row_params(w), M, D.N, d, D.V, 0,
                                                         D.eps, nullptr, dout, dE, dA, XF, XA, P5, nullptr);
  VLPET_LAUNCH_OK();
  VLPET_TRY(launch_colsum(dout, bf, XF, 0, nullptr, 0, M, d, 1.f, G.dln_f_w, st));
  VLPET_TRY(launch_colsum(dout, bf, XA, 0, nullptr, 0, M, d, 1.f, G.dln_p_w, st));
  VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, nullptr, 0, M, d, 1.f, G.dln_f_b, st));
  VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, nullptr, 0, M, d, 1.f, G.dln_p_b, st));
  VLPET_TRY(launch_wgrad(dA, 0, d, P5, 0, 5, M, G.dWp, 1.f, st));
  VLPET_TRY(launch_colsum(dA, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dbp, st));
  if (G.dE_img) {
    for (int i = 0; i < D.n_img; ++i) {
      if (img_ids == nullptr && i > 0) break;
      VLPET_TRY(launch_colsum(dout, bf, nullptr, 0, img_ids, i, M, d, 1.f, G.dE_img + (size_t)i * d, st));
    }
  }
  // recompute the projector
  GemmArgs g = linear_nt(feats, bf, w.Wd, w.bd, bf, Z, 0, M, r, D.F);
  g.act = 1;
  g.pre = Apre;
  VLPET_TRY(launch_gemm(g, st));
  VLPET_TRY(launch_gemm(linear_nt(Z, 0, w.Wu, w.bu, bf, U, 0, M, d, r), st));
  const float* dU = dE;
  if (D.gated) {
    g = linear_nt(feats, bf, w.Gd, w.gbd, bf, Q, 0, M, rg, D.F);
    g.act = 1;
    g.pre = Ppre;
    VLPET_TRY(launch_gemm(g, st));
    VLPET_TRY(launch_gemm(linear_nt(Q, 0, w.Gu, w.gbu, bf, T, 0, M, d, rg), st));
    lr_gate_bwd_kernel<<<ew_blocks(M * d), 256, 0, st>>>(dE, U, T, M * d, D.residual);
    VLPET_LAUNCH_OK();
    dU = U;
    VLPET_TRY(launch_wgrad(T, 0, d, Q, 0, rg, M, G.dGu, 1.f, st));
    VLPET_TRY(launch_colsum(T, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dgbu, st));
    VLPET_TRY(launch_gemm(linear_nn(T, 0, w.Gu, bf, DQ, 0, M, d, rg), st));
    mul_gelu_grad_kernel<<<ew_blocks(M * rg), 256, 0, st>>>(DQ, Ppre, M * rg);
    VLPET_LAUNCH_OK();
    VLPET_TRY(launch_wgrad(DQ, 0, rg, feats, bf, D.F, M, G.dGd, 1.f, st));
    VLPET_TRY(launch_colsum(DQ, 0, nullptr, 0, nullptr, 0, M, rg, 1.f, G.dgbd, st));
  }
  VLPET_TRY(launch_wgrad(dU, 0, d, Z, 0, r, M, G.dWu, 1.f, st));
  VLPET_TRY(launch_colsum(dU, 0, nullptr, 0, nullptr, 0, M, d, 1.f, G.dbu, st));
  VLPET_TRY(launch_gemm(linear_nn(dU, 0, w.Wu, bf, DZ, 0, M, d, r), st));
  mul_gelu_grad_kernel<<<ew_blocks(M * r), 256, 0, st>>>(DZ, Apre, M * r);
  VLPET_LAUNCH_OK();
  VLPET_TRY(launch_wgrad(DZ, 0, r, feats, bf, D.F, M, G.dWd, 1.f, st));
  VLPET_TRY(launch_colsum(DZ, 0, nullptr, 0, nullptr, 0, M, r, 1.f, G.dbd, st));
  return 0;
}

}  // namespace vlpet
