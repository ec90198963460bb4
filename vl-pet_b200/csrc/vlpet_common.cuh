// Shared helpers for libvlpet.so (sm_100a only).  No torch types anywhere in this library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vlpet.h"

namespace vlpet {

// ---- error plumbing ------------------------------------------------------------------------------------
char* tls_error_buffer();  // 512 bytes, thread local (vlpet_api.cu)
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}
extern std::atomic<uint64_t> g_launches;
// 16 x uint32 of host-mapped pinned memory (device view / host view; nullptr if the allocation failed): a kernel whose
// barrier wait times out writes its identity there before it traps, see ptx::mbar_wait_dbg
uint32_t* trap_buffer_dev();
const uint32_t* trap_buffer_host();
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define VLPET_CUDA_OK(expr)                                                                           \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return ::vlpet::fail((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                                 \
  } while (0)
#define VLPET_LAUNCH_OK()                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess)                                                                                  \
      return ::vlpet::fail((int)_e, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                                       \
    ::vlpet::count_launch();                                                                                \
  } while (0)
#define VLPET_TRY(expr)     \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device's copy of a kernel: remember per device
// ordinal what was set (one process may drive several GPUs).  `slot` is one zero-initialised int[64] per kernel.
inline int ensure_dyn_smem(const void* kern, int (&slot)[64], int bytes) {
  int dev = 0;
  VLPET_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || slot[dev] < bytes) {
    VLPET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) slot[dev] = bytes;
  }
  return 0;
}

inline size_t esize(int dtype) { return dtype == VLPET_BF16 ? 2 : 4; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace (256-byte granules)
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    char* r = base ? base + off : nullptr;
    off += bytes;
    return reinterpret_cast<T*>(r);
  }
  bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

// ---- device math ---------------------------------------------------------------------------------------
// gelu_new = transformers.activations.NewGELUActivation: 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
__device__ __forceinline__ float gelu_new_f(float t) {
  const float c = 0.7978845608028654f, k = 0.044715f;
  return 0.5f * t * (1.0f + tanhf(c * (t + k * t * t * t)));
}
__device__ __forceinline__ float gelu_new_grad_f(float t) {
  const float c = 0.7978845608028654f, k = 0.044715f;
  float th = tanhf(c * (t + k * t * t * t));
  return 0.5f * (1.0f + th) + 0.5f * t * (1.0f - th * th) * c * (1.0f + 3.0f * k * t * t);
}
__device__ __forceinline__ float sigmoid_f(float t) { return 1.0f / (1.0f + expf(-t)); }

// Counter-based dropout: one 64-bit hash per 4 consecutive elements, 16 bits per element.
// keep(idx) = bits16 >= thr16 with thr16 = round(p * 65536); kept values are scaled by 1/(1-p).
// The hash is a 5-round Philox-2x32-style network (32x32->64 multiply, xor with the round key and the other word, swap):
// one IMAD.WIDE + one 3-input LOP3 + one IADD per round, 15 instructions against ~25 for the splitmix64 of round 1 (two
// 64-bit multiplies = 6 IMADs + carries).  Counter = idx4, key = seed: consecutive seeds (CUDA-graph replays bump the
// device seed by one) and consecutive counters give uncorrelated keep bits (checked offline over 4 M elements per seed:
// keep fraction within 3e-4 of 1-p, lag-1..768 and cross-seed correlations < 1.2e-3 = noise).
__device__ __forceinline__ uint64_t drop_hash4(uint64_t seed, uint64_t idx4) {
  uint32_t c0 = (uint32_t)idx4, c1 = (uint32_t)(idx4 >> 32) ^ (uint32_t)(seed >> 32);
  uint32_t k = (uint32_t)seed;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint64_t p = (uint64_t)c0 * 0xD256D193u;
    c0 = (uint32_t)(p >> 32) ^ k ^ c1;
    c1 = (uint32_t)p;
    k += 0x9E3779B9u;
  }
  return ((uint64_t)c1 << 32) | c0;
}
__host__ __device__ __forceinline__ uint32_t drop_thr16(float p) {
  float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
}
// multiplicative mask value for element idx: 0 or inv_keep
__device__ __forceinline__ float drop_scale(uint64_t seed, uint32_t thr16, float inv_keep, int64_t idx) {
  if (thr16 == 0) return 1.0f;
  uint64_t h = drop_hash4(seed, (uint64_t)idx >> 2);
  uint32_t bits = (uint32_t)(h >> (16 * ((uint32_t)idx & 3u))) & 0xffffu;
  return bits >= thr16 ? inv_keep : 0.0f;
}

__device__ __forceinline__ float ld_as_float(const void* p, int64_t i, bool bf16) {
  return bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]) : static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_from_float(void* p, int64_t i, float v, bool bf16) {
  if (bf16)
    static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else
    static_cast<float*>(p)[i] = v;
}
// the same mask for 8 consecutive elements starting at idx (a multiple of 8): two hashes instead of eight
__device__ __forceinline__ void drop_scale8(uint64_t seed, uint32_t thr16, float inv_keep, int64_t idx, float (&m)[8]) {
  if (thr16 == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = 1.0f;
    return;
  }
  const uint64_t h0 = drop_hash4(seed, (uint64_t)idx >> 2), h1 = drop_hash4(seed, ((uint64_t)idx >> 2) + 1);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t bits = (uint32_t)((e < 4 ? h0 : h1) >> (16 * (e & 3))) & 0xffffu;
    m[e] = bits >= thr16 ? inv_keep : 0.0f;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- generic (CUDA-core) building blocks, vlpet_generic.cu --------------------------------------------------
int generic_k1_fwd(const VlpetK1Desc&, const void* x1, const void* x2, const VlpetK1Params&, void* out, void* ws,
                   size_t ws_bytes, cudaStream_t);
int generic_k1_bwd(const VlpetK1Desc&, const void* x1, const void* x2, const void* dout, const VlpetK1Params&,
                   void* dx1, void* dx2, const VlpetK1Grads&, void* ws, size_t ws_bytes, cudaStream_t);
size_t generic_k1_fwd_ws(const VlpetK1Desc&);
size_t generic_k1_bwd_ws(const VlpetK1Desc&);
int generic_k2_fwd(const VlpetK2Desc&, const void* kv, const void* y, const VlpetK2Params&, void* out, void* ws,
                   size_t ws_bytes, cudaStream_t);
int generic_k2_bwd(const VlpetK2Desc&, const void* kv, const void* dout, const VlpetK2Params&, void* dkv,
                   const VlpetK2Grads&, void* ws, size_t ws_bytes, cudaStream_t);
size_t generic_k2_fwd_ws(const VlpetK2Desc&);
size_t generic_k2_bwd_ws(const VlpetK2Desc&);
int generic_k3_fwd(const VlpetK3Desc&, const void* feats, const void* pos, const int64_t* img_ids,
                   const int64_t* obj_ids, const VlpetK3Params&, void* out, float* save, void* ws, size_t ws_bytes,
                   cudaStream_t);
int generic_k3_bwd(const VlpetK3Desc&, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                   const VlpetK3Params&, const float* save, void* dfeats, const VlpetK3Grads&, void* ws,
                   size_t ws_bytes, cudaStream_t);
size_t generic_k3lr_fwd_ws(const VlpetK3LRDesc&);
size_t generic_k3lr_bwd_ws(const VlpetK3LRDesc&);
int generic_k3lr_fwd(const VlpetK3LRDesc&, const void* feats, const void* pos, const int64_t* img_ids, const int64_t* obj_ids,
                     const VlpetK3LRParams&, void* out, float* save, void* ws, size_t ws_bytes, cudaStream_t);
int generic_k3lr_bwd(const VlpetK3LRDesc&, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                     const VlpetK3LRParams&, const float* save, const VlpetK3LRGrads&, void* ws, size_t ws_bytes, cudaStream_t);
size_t generic_k3_fwd_ws(const VlpetK3Desc&);
size_t generic_k3_bwd_ws(const VlpetK3Desc&);

// ---- fused sm_100a kernels (tcgen05 + TMA), vlpet_k1_sm100.cu ----------------------------------------------
bool fused_k1_fwd_supported(const VlpetK1Desc&);
size_t fused_k1_fwd_ws(const VlpetK1Desc&);
int fused_k1_fwd(const VlpetK1Desc&, const void* x1, const void* x2, const VlpetK1Params&, void* out, void* ws,
                 size_t ws_bytes, cudaStream_t);

// fused K1 backward (activation gradients, vlpet_k1_bwd_sm100.cu) + the weight-gradient GEMMs below
bool fused_k1_bwd_supported(const VlpetK1Desc&);
size_t fused_k1_bwd_ws(const VlpetK1Desc&);
int fused_k1_bwd(const VlpetK1Desc&, const void* x1, const void* x2, const void* dout, const VlpetK1Params&, void* dx1,
                 void* dx2, const VlpetK1Grads&, void* ws, size_t ws_bytes, cudaStream_t);

// K2 through the same fused kernels (ungated form)
bool fused_k2_supported(const VlpetK2Desc&);
size_t fused_k2_bwd_ws(const VlpetK2Desc&);
int fused_k2_fwd(const VlpetK2Desc&, const void* kv, const void* y, const VlpetK2Params&, void* out, cudaStream_t);
int fused_k2_bwd(const VlpetK2Desc&, const void* kv, const void* dout, const VlpetK2Params&, void* dkv, const VlpetK2Grads&,
                 void* ws, size_t ws_bytes, cudaStream_t);

// same with the x2 scale of the K1 form: dkv = kappa * dout + adapter-path gradient (the row-wise gate path, vlpet_rows.cu)
int fused_k2_bwd_kappa(const VlpetK2Desc&, float kappa, const void* kv, const void* dout, const VlpetK2Params&, void* dkv,
                       const VlpetK2Grads&, void* ws, size_t ws_bytes, cudaStream_t);
// The ungated backward as a building block of composed paths (vlpet_wide.cu: ranks above one TMEM bucket as two rank
// halves): kap_src = the [M, d] tensor the kappa term of dkv is read from (nullptr: dout itself; dkv for "accumulate into
// the previous half's result", in place), ldo_wu = row pitch in floats of the dWu buffer (0: r; a column slice of a wider
// [d, r_total] gradient otherwise).
struct BwdExtras {
  const void* kap_src;
  int64_t ldo_wu;
};
int fused_k2_bwd_ex(const VlpetK2Desc&, float kappa, const void* kv, const void* dout, const VlpetK2Params&, void* dkv,
                    const VlpetK2Grads&, const BwdExtras&, void* ws, size_t ws_bytes, cudaStream_t);
// column sums of a [M, pitch] bf16 matrix (first ncols columns) accumulated into out[ncols] (vlpet_k1_bwd_sm100.cu)
int colsum_bf16(const void* A, int pitch, int ncols, float* out, int64_t M, int sms, cudaStream_t st);

// ---- row-wise K1 (middleX / middleY / small gates, ungated form, ranks <= 16), vlpet_rows.cu -----------------------
bool rows_k1_supported(const VlpetK1Desc&, bool bwd);
size_t rows_k1_fwd_ws(const VlpetK1Desc&);
size_t rows_k1_bwd_ws(const VlpetK1Desc&);
int rows_k1_fwd(const VlpetK1Desc&, const void* x1, const void* x2, const VlpetK1Params&, void* out, void* ws, size_t ws_bytes,
                cudaStream_t);
int rows_k1_bwd(const VlpetK1Desc&, const void* x1, const void* x2, const void* dout, const VlpetK1Params&, void* dx1, void* dx2,
                const VlpetK1Grads&, void* ws, size_t ws_bytes, cudaStream_t);

// ---- K1, large gate, 96 < max(r, rg) <= 192: composed from the ungated fused kernels per rank half, vlpet_wide.cu -------
bool wide_k1_supported(const VlpetK1Desc&, bool bwd);
size_t wide_k1_fwd_ws(const VlpetK1Desc&);
size_t wide_k1_bwd_ws(const VlpetK1Desc&);
int wide_k1_fwd(const VlpetK1Desc&, const void* x1, const void* x2, const VlpetK1Params&, void* out, void* ws, size_t ws_bytes,
                cudaStream_t);
int wide_k1_bwd(const VlpetK1Desc&, const void* x1, const void* x2, const void* dout, const VlpetK1Params&, void* dx1, void* dx2,
                const VlpetK1Grads&, void* ws, size_t ws_bytes, cudaStream_t);

// the adapter branch alone at 96 < r <= 192 (y1 = kappa x2 + alpha (Up(gelu_new(Down x2)) + bu)) for the row-wise gates
bool wide_adapter_supported(const VlpetK1Desc&, bool bwd);
size_t wide_adapter_ws(const VlpetK1Desc&, bool bwd);
int wide_adapter_fwd(const VlpetK1Desc&, const void* x2, const VlpetK1Params&, void* y1, void* ws, size_t ws_bytes, cudaStream_t);
int wide_adapter_bwd(const VlpetK1Desc&, const void* x2, const void* dy1, const VlpetK1Params&, void* dx2, const VlpetK1Grads&,
                     void* ws_fwd, void* ws, size_t ws_bytes, cudaStream_t);

// ---- token-contracted weight-gradient GEMM (tcgen05), vlpet_wgrad_sm100.cu -----------------------------------
bool wgrad_sm100_supported(int d, int nout);
int wgrad_sm100(int npairs, const void* const* A, const int64_t* lda, const void* const* B, const int64_t* ldb,
                const int* nb_valid, float* const* out, float* const* bias, const float* scale, const int* transposed,
                int64_t Mtok, int d, int nout, int sm_count, cudaStream_t st, const int64_t* ldo = nullptr);
int device_sm_count();
// LayerNorm behind the PET sites, vlpet_layernorm.cu
bool layernorm_supported(int d, int dtype);
int layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M, int d,
                  float eps, cudaStream_t st);
int layernorm_bwd(const void* x, const void* dy, const float* w, const float* mean, const float* rstd, void* dx, float* dw,
                  float* db, int64_t M, int d, cudaStream_t st);
int dropout_add_layernorm_fwd(const void* h, const void* res, const float* w, const float* b, void* y, void* xs, float* mean,
                              float* rstd, int64_t M, int d, float eps, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                              cudaStream_t st);
int dropout_add_layernorm_bwd(const void* xs, const void* dy, const float* w, const float* mean, const float* rstd, void* dres, void* dh,
                              float* dw, float* db, int64_t M, int d, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                              cudaStream_t st);
// short-sequence attention, vlpet_attention.cu
int attn_run(bool bwd, const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, void* out, float* lse,
             const void* o, const void* dout, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs, int64_t dv_rs, int B, int H,
             int Lq, int Lk, int causal, float p_drop, uint64_t seed, const uint64_t* seed_dev, cudaStream_t st);
int set_k1_trace(unsigned long long* dev_buf);   // developer hook, tools/trace_k1.py
int set_k1_bwd_trace(unsigned long long* dev_buf);
int set_k1_bwd_parts(int parts);                 // developer hook: bit 0 tile kernel, bit 1 column sums, bit 2 weight-gradient GEMM
int set_k1_pairs(int mode);                       // developer hook: -1 auto, 0 single CTAs, 1 CTA pairs (cta_group::2)
// dense projection GEMM (tcgen05), vlpet_gemm_sm100.cu: C fp32 = A bf16 * W^T bf16 + bias
bool gemm_sm100_supported(int64_t M, int N, int K, int64_t ldc);
int gemm_sm100(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, float* C, int64_t ldc, int64_t M,
               int N, int K, cudaStream_t st);

}  // namespace vlpet
