// C-ABI entry points of libvlpet.so (declared in include/vlpet.h): argument validation, dispatch between the
// fused sm_100a kernels and the generic CUDA-core kernels, and the flat-bucket helpers (cast / AdamW / sumsq).
#include <cstring>

#include "sm100_ptx.cuh"
#include "vlpet_common.cuh"

namespace vlpet {
std::atomic<uint64_t> g_launches{0};
char* tls_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

static uint32_t *g_trap_host = nullptr, *g_trap_dev = nullptr;
uint32_t* trap_buffer_dev() {
  static bool once = []() {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return false; }
    memset(h, 0, 64);
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaGetLastError(); return false; }
    g_trap_host = static_cast<uint32_t*>(h);
    g_trap_dev = static_cast<uint32_t*>(d);
    return true;
  }();
  (void)once;
  return g_trap_dev;
}
const uint32_t* trap_buffer_host() { return g_trap_host; }

int device_sm_count() {   // SM count of the CURRENT device (cached per ordinal); -1 if it is not a compute-capability-10 part
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && cache[dev] != 0) return cache[dev];
  int sms = 0, major = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return 0;
  const int v = major == 10 ? sms : -1;
  if (dev >= 0 && dev < 64) cache[dev] = v;
  return v;
}

namespace {
int check_k1(const VlpetK1Desc* D, const VlpetK1Params* w) {
  if (!D || !w) return fail(VLPET_E_BADARG, "k1: null desc/params");
  if (D->M <= 0 || D->d <= 0 || D->r <= 0) return fail(VLPET_E_BADARG, "k1: M, d, r must be positive");
  if (D->dtype != VLPET_F32 && D->dtype != VLPET_BF16) return fail(VLPET_E_BADARG, "k1: bad dtype %d", D->dtype);
  if (D->gate < VLPET_GATE_NONE || D->gate > VLPET_GATE_SMALL) return fail(VLPET_E_BADARG, "k1: bad gate %d", D->gate);
  if (!w->Wd || !w->bd || !w->Wu || !w->bu) return fail(VLPET_E_BADARG, "k1: adapter weights missing");
  if (D->gate == VLPET_GATE_LARGE && (D->rg <= 0 || !w->Gd || !w->gbd || !w->Gu || !w->gbu))
    return fail(VLPET_E_BADARG, "k1: large gate needs rg > 0 and Gd/gbd/Gu/gbu");
  if ((D->gate == VLPET_GATE_MIDDLE_X || D->gate == VLPET_GATE_SMALL) && (!w->gw || !w->gb))
    return fail(VLPET_E_BADARG, "k1: middle_x/small gate needs gw/gb");
  if (D->gate == VLPET_GATE_MIDDLE_Y && !w->gz) return fail(VLPET_E_BADARG, "k1: middle_y gate needs gz");
  if (!(D->p_drop >= 0.f && D->p_drop < 1.f)) return fail(VLPET_E_BADARG, "k1: p_drop must be in [0,1)");
  if (D->gate == VLPET_GATE_SMALL && (D->L <= 0 || D->M % D->L != 0))
    return fail(VLPET_E_BADARG, "k1: small gate needs L > 0 dividing M");
  return 0;
}
bool use_fused_fwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && fused_k1_fwd_supported(D); }
bool use_fused_bwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && fused_k1_bwd_supported(D); }
bool use_rows_fwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && !use_fused_fwd(D) && rows_k1_supported(D, false); }
bool use_rows_bwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && !use_fused_bwd(D) && rows_k1_supported(D, true); }
bool use_wide_fwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && !use_fused_fwd(D) && !use_rows_fwd(D) && wide_k1_supported(D, false); }
bool use_wide_bwd(const VlpetK1Desc& D) { return D.impl != VLPET_IMPL_GENERIC && !use_fused_bwd(D) && !use_rows_bwd(D) && wide_k1_supported(D, true); }
}  // namespace
}  // namespace vlpet

using namespace vlpet;

extern "C" {

int vlpet_version(void) { return VLPET_VERSION; }
const char* vlpet_last_error(void) { return tls_error_buffer(); }
uint64_t vlpet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vlpet_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  VLPET_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  VLPET_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

// ---- K1 ----------------------------------------------------------------------------------------------------
int vlpet_k1_fwd_is_fused(const VlpetK1Desc* D) {
  return !D ? 0 : (use_fused_fwd(*D) ? 1 : (use_rows_fwd(*D) ? 2 : (use_wide_fwd(*D) ? 3 : 0)));
}
int vlpet_k1_bwd_is_fused(const VlpetK1Desc* D) {
  return !D ? 0 : (use_fused_bwd(*D) ? 1 : (use_rows_bwd(*D) ? 2 : (use_wide_bwd(*D) ? 3 : 0)));
}

size_t vlpet_k1_fwd_workspace_bytes(const VlpetK1Desc* D) {
  if (!D) return 0;
  return use_fused_fwd(*D) ? fused_k1_fwd_ws(*D)
                           : (use_rows_fwd(*D) ? rows_k1_fwd_ws(*D) : (use_wide_fwd(*D) ? wide_k1_fwd_ws(*D) : generic_k1_fwd_ws(*D)));
}
size_t vlpet_k1_bwd_workspace_bytes(const VlpetK1Desc* D) {
  if (!D) return 0;
  return use_fused_bwd(*D) ? fused_k1_bwd_ws(*D)
                           : (use_rows_bwd(*D) ? rows_k1_bwd_ws(*D) : (use_wide_bwd(*D) ? wide_k1_bwd_ws(*D) : generic_k1_bwd_ws(*D)));
}

int vlpet_k1_fwd(const VlpetK1Desc* D, const void* x1, const void* x2, const VlpetK1Params* w, void* out, void* ws,
                 size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k1(D, w));
  if (!x1 || !x2 || !out) return fail(VLPET_E_BADARG, "k1_fwd: null activation pointer");
  if (!aligned16(x1) || !aligned16(x2) || !aligned16(out)) return fail(VLPET_E_ALIGN, "k1_fwd: activations must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_fused_fwd(*D)) return fused_k1_fwd(*D, x1, x2, *w, out, ws, ws_bytes, st);
  if (use_rows_fwd(*D)) return rows_k1_fwd(*D, x1, x2, *w, out, ws, ws_bytes, st);
  if (use_wide_fwd(*D)) return wide_k1_fwd(*D, x1, x2, *w, out, ws, ws_bytes, st);
  if (D->impl == VLPET_IMPL_FUSED)
    return fail(VLPET_E_UNSUPPORTED, "k1_fwd: fused kernel does not cover d=%d r=%d rg=%d gate=%d dtype=%d", D->d, D->r,
                D->rg, D->gate, D->dtype);
  return generic_k1_fwd(*D, x1, x2, *w, out, ws, ws_bytes, st);
}

int vlpet_k1_bwd(const VlpetK1Desc* D, const void* x1, const void* x2, const void* dout, const VlpetK1Params* w,
                 void* dx1, void* dx2, const VlpetK1Grads* g, void* ws, size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k1(D, w));
  if (!x1 || !x2 || !dout || !dx1 || !dx2 || !g) return fail(VLPET_E_BADARG, "k1_bwd: null pointer");
  if (!aligned16(x1) || !aligned16(x2) || !aligned16(dout) || !aligned16(dx1) || !aligned16(dx2))
    return fail(VLPET_E_ALIGN, "k1_bwd: activations must be 16-byte aligned");
  if (use_fused_bwd(*D))
    return fused_k1_bwd(*D, x1, x2, dout, *w, dx1, dx2, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (use_rows_bwd(*D))
    return rows_k1_bwd(*D, x1, x2, dout, *w, dx1, dx2, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (use_wide_bwd(*D))
    return wide_k1_bwd(*D, x1, x2, dout, *w, dx1, dx2, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (D->impl == VLPET_IMPL_FUSED)
    return fail(VLPET_E_UNSUPPORTED, "k1_bwd: fused kernel does not cover d=%d r=%d rg=%d gate=%d dtype=%d", D->d, D->r,
                D->rg, D->gate, D->dtype);
  return generic_k1_bwd(*D, x1, x2, dout, *w, dx1, dx2, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- K3-LR ---------------------------------------------------------------------------------------------------
static int check_k3lr(const VlpetK3LRDesc* D, const VlpetK3LRParams* w) {
  if (!D || !w) return fail(VLPET_E_BADARG, "k3lr: null desc/params");
  if (D->M <= 0 || D->d <= 0 || D->F <= 0 || D->N <= 0 || D->V <= 0 || D->n_img <= 0 || D->r <= 0 || (D->gated && D->rg <= 0))
    return fail(VLPET_E_BADARG, "k3lr: sizes must be positive");
  if (D->dtype != VLPET_F32 && D->dtype != VLPET_BF16) return fail(VLPET_E_BADARG, "k3lr: bad dtype %d", D->dtype);
  if (!w->Wd || !w->bd || !w->Wu || !w->bu || !w->ln_f_w || !w->ln_f_b || !w->Wp || !w->bp || !w->ln_p_w || !w->ln_p_b ||
      !w->E_img || !w->E_obj || (D->gated && (!w->Gd || !w->gbd || !w->Gu || !w->gbu)))
    return fail(VLPET_E_BADARG, "k3lr: weights missing");
  if (D->impl == VLPET_IMPL_FUSED) return fail(VLPET_E_UNSUPPORTED, "k3lr: only the generic CUDA path exists");
  return 0;
}
size_t vlpet_k3lr_fwd_workspace_bytes(const VlpetK3LRDesc* D) { return D ? generic_k3lr_fwd_ws(*D) : 0; }
size_t vlpet_k3lr_bwd_workspace_bytes(const VlpetK3LRDesc* D) { return D ? generic_k3lr_bwd_ws(*D) : 0; }
int vlpet_k3lr_fwd(const VlpetK3LRDesc* D, const void* feats, const void* pos, const int64_t* img_ids, const int64_t* obj_ids,
                   const VlpetK3LRParams* w, void* out, float* save, void* ws, size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k3lr(D, w));
  if (!feats || !pos || !out || !save) return fail(VLPET_E_BADARG, "k3lr_fwd: null pointer");
  return generic_k3lr_fwd(*D, feats, pos, img_ids, obj_ids, *w, out, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int vlpet_k3lr_bwd(const VlpetK3LRDesc* D, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                   const VlpetK3LRParams* w, const float* save, const VlpetK3LRGrads* g, void* ws, size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k3lr(D, w));
  if (!feats || !pos || !dout || !save || !g) return fail(VLPET_E_BADARG, "k3lr_bwd: null pointer");
  return generic_k3lr_bwd(*D, feats, pos, img_ids, dout, *w, save, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- LayerNorm behind the PET sites --------------------------------------------------------------------------
int vlpet_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M, int32_t d,
                        float eps, int32_t dtype, void* stream) {
  if (!x || !w || !b || !y || !mean || !rstd || M <= 0) return fail(VLPET_E_BADARG, "layernorm_fwd: bad arguments");
  if (!layernorm_supported(d, dtype)) return fail(VLPET_E_UNSUPPORTED, "layernorm_fwd: needs bf16, d %% 256 == 0, d <= 1024 (d=%d)", d);
  if (!aligned16(x) || !aligned16(y) || !aligned16(w) || !aligned16(b)) return fail(VLPET_E_ALIGN, "layernorm_fwd: misaligned");
  return layernorm_fwd(x, w, b, y, mean, rstd, M, d, eps, static_cast<cudaStream_t>(stream));
}
int vlpet_layernorm_bwd(const void* x, const void* dy, const float* w, const float* mean, const float* rstd, void* dx, float* dw,
                        float* db, int64_t M, int32_t d, int32_t dtype, void* stream) {
  if (!x || !dy || !w || !mean || !rstd || !dx || M <= 0) return fail(VLPET_E_BADARG, "layernorm_bwd: bad arguments");
  if (!layernorm_supported(d, dtype)) return fail(VLPET_E_UNSUPPORTED, "layernorm_bwd: needs bf16, d %% 256 == 0, d <= 1024 (d=%d)", d);
  if (!aligned16(x) || !aligned16(dy) || !aligned16(dx) || !aligned16(w)) return fail(VLPET_E_ALIGN, "layernorm_bwd: misaligned");
  return layernorm_bwd(x, dy, w, mean, rstd, dx, dw, db, M, d, static_cast<cudaStream_t>(stream));
}

int vlpet_dropout_add_layernorm_fwd(const void* h, const void* res, const float* w, const float* b, void* y, void* xs, float* mean,
                                    float* rstd, int64_t M, int32_t d, float eps, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                    void* stream) {
  if (!h || !res || !w || !b || !y || !xs || !mean || !rstd || M <= 0) return fail(VLPET_E_BADARG, "dropout_add_layernorm_fwd: bad arguments");
  if (!(p_drop >= 0.f && p_drop < 1.f)) return fail(VLPET_E_BADARG, "dropout_add_layernorm_fwd: p_drop must be in [0,1)");
  if (!layernorm_supported(d, VLPET_BF16)) return fail(VLPET_E_UNSUPPORTED, "dropout_add_layernorm_fwd: needs d %% 256 == 0, d <= 1024 (d=%d)", d);
  if (!aligned16(h) || !aligned16(res) || !aligned16(y) || !aligned16(xs) || !aligned16(w) || !aligned16(b))
    return fail(VLPET_E_ALIGN, "dropout_add_layernorm_fwd: misaligned");
  return dropout_add_layernorm_fwd(h, res, w, b, y, xs, mean, rstd, M, d, eps, p_drop, seed, seed_dev, static_cast<cudaStream_t>(stream));
}
int vlpet_dropout_add_layernorm_bwd(const void* xs, const void* dy, const float* w, const float* mean, const float* rstd, void* dres,
                                    void* dh, float* dw, float* db, int64_t M, int32_t d, float p_drop, uint64_t seed,
                                    const uint64_t* seed_dev, void* stream) {
  if (!xs || !dy || !w || !mean || !rstd || !dres || !dh || M <= 0) return fail(VLPET_E_BADARG, "dropout_add_layernorm_bwd: bad arguments");
  if (!layernorm_supported(d, VLPET_BF16)) return fail(VLPET_E_UNSUPPORTED, "dropout_add_layernorm_bwd: needs d %% 256 == 0, d <= 1024 (d=%d)", d);
  if (!aligned16(xs) || !aligned16(dy) || !aligned16(dres) || !aligned16(dh) || !aligned16(w))
    return fail(VLPET_E_ALIGN, "dropout_add_layernorm_bwd: misaligned");
  return dropout_add_layernorm_bwd(xs, dy, w, mean, rstd, dres, dh, dw, db, M, d, p_drop, seed, seed_dev, static_cast<cudaStream_t>(stream));
}

// developer hook: a kernel that waits on a barrier nobody completes -- proves that the trap record reaches the host
namespace vlpet {
__global__ void selftest_trap_kernel(uint32_t* dbg) {
  __shared__ uint64_t mbar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) ptx::mbar_wait_dbg(b, 0, dbg, 4242);
}
}  // namespace vlpet
__attribute__((visibility("default"))) int vlpet_debug_selftest_trap(void) {
  vlpet::selftest_trap_kernel<<<1, 32>>>(vlpet::trap_buffer_dev());
  return (int)cudaDeviceSynchronize();
}
// developer hook: who timed out?  {source line, block, thread, parity, barrier address} of the last barrier-wait trap
__attribute__((visibility("default"))) int vlpet_debug_last_trap(uint32_t* out5) {
  const uint32_t* h = vlpet::trap_buffer_host();
  if (!h || !out5) return 1;
  for (int i = 0; i < 5; ++i) out5[i] = h[i];
  return 0;
}
// developer hook (not part of include/vlpet.h): phase timestamps of the fused K1 forward, see tools/trace_k1.py
__attribute__((visibility("default"))) int vlpet_debug_set_k1_trace(void* dev_buf) {
  return vlpet::set_k1_trace(static_cast<unsigned long long*>(dev_buf));
}
// developer hook: force (1) / forbid (0) / auto-select (-1) the CTA-pair (cta_group::2) variant of the fused K1 forward
__attribute__((visibility("default"))) int vlpet_debug_set_k1_pairs(int mode) { return vlpet::set_k1_pairs(mode); }
__attribute__((visibility("default"))) int vlpet_debug_set_k1_bwd_parts(int parts) { return vlpet::set_k1_bwd_parts(parts); }
__attribute__((visibility("default"))) int vlpet_debug_set_k1_bwd_trace(void* dev_buf) {
  return vlpet::set_k1_bwd_trace(static_cast<unsigned long long*>(dev_buf));
}

// ---- weight-gradient GEMM ------------------------------------------------------------------------------------
int vlpet_wgrad_bf16(const VlpetWgradPair* pairs, int32_t npairs, int64_t Mtok, int32_t d, int32_t nout, void* stream) {
  if (!pairs || npairs < 1 || npairs > 4 || Mtok <= 0) return fail(VLPET_E_BADARG, "wgrad: bad arguments");
  const void *A[4], *B[4];
  int64_t lda[4], ldb[4];
  int nbv[4], tr[4];
  float *out[4], *bias[4], sc[4];
  for (int i = 0; i < npairs; ++i) {
    const VlpetWgradPair& q = pairs[i];
    if (!q.A || !q.B || !q.out) return fail(VLPET_E_BADARG, "wgrad: null pointer in pair %d", i);
    if (!aligned16(q.A) || !aligned16(q.B) || !aligned16(q.out) || (q.lda & 7) || (q.ldb & 7))
      return fail(VLPET_E_ALIGN, "wgrad: operands must be 16-byte aligned with pitches that are multiples of 8");
    if (q.nb_valid != nout && q.nb_valid != nout + 1) return fail(VLPET_E_BADARG, "wgrad: nb_valid must be nout or nout+1");
    if (q.bias && q.nb_valid != nout + 1) return fail(VLPET_E_BADARG, "wgrad: bias needs the ones column (nb_valid = nout+1)");
    A[i] = q.A; B[i] = q.B; lda[i] = q.lda; ldb[i] = q.ldb; nbv[i] = q.nb_valid; tr[i] = q.transposed;
    out[i] = q.out; bias[i] = q.bias; sc[i] = q.scale;
  }
  int sms = device_sm_count();
  if (sms <= 0) return fail(VLPET_E_NODEVICE, "wgrad: no CUDA device");
  return wgrad_sm100(npairs, A, lda, B, ldb, nbv, out, bias, sc, tr, Mtok, d, nout, sms, static_cast<cudaStream_t>(stream));
}

// ---- K2 ----------------------------------------------------------------------------------------------------
static int check_k2(const VlpetK2Desc* D, const VlpetK2Params* w) {
  if (!D || !w) return fail(VLPET_E_BADARG, "k2: null desc/params");
  if (D->M <= 0 || D->d <= 0 || D->r <= 0) return fail(VLPET_E_BADARG, "k2: M, d, r must be positive");
  if (D->dtype != VLPET_F32 && D->dtype != VLPET_BF16) return fail(VLPET_E_BADARG, "k2: bad dtype %d", D->dtype);
  if (!w->Wd || !w->bd || !w->Wu || !w->bu) return fail(VLPET_E_BADARG, "k2: weights missing");
  return 0;
}
static bool use_fused_k2(const VlpetK2Desc& D) { return D.impl != VLPET_IMPL_GENERIC && fused_k2_supported(D); }
// ranks too small for a tensor-core tile (r <= 16, e.g. BASELINE config 4: r = 4): the ungated form of the row-wise K1
// kernels (csrc/vlpet_rows.cu, one launch forward): out = y + 1 * (0 * kv + sf * (Up(gelu_new(Down kv)) + bu))
static VlpetK1Desc k2_rows_desc(const VlpetK2Desc& D) {
  VlpetK1Desc K;
  memset(&K, 0, sizeof(K));
  K.M = D.M; K.L = 1; K.d = D.d; K.r = D.r; K.rg = 0; K.gate = VLPET_GATE_NONE; K.dtype = D.dtype; K.impl = D.impl;
  K.s = 1.0f; K.alpha = D.sf; K.kappa = 0.0f;
  return K;
}
static bool use_rows_k2(const VlpetK2Desc& D, bool bwd) {
  return D.impl != VLPET_IMPL_GENERIC && !use_fused_k2(D) && D.r <= 16 && rows_k1_supported(k2_rows_desc(D), bwd);
}
static VlpetK1Params k2_as_k1_params(const VlpetK2Params& w) {
  VlpetK1Params P;
  memset(&P, 0, sizeof(P));
  P.Wd = w.Wd; P.bd = w.bd; P.Wu = w.Wu; P.bu = w.bu;
  return P;
}
size_t vlpet_k2_fwd_workspace_bytes(const VlpetK2Desc* D) {
  if (!D) return 0;
  const size_t a = generic_k2_fwd_ws(*D), b = use_rows_k2(*D, false) ? rows_k1_fwd_ws(k2_rows_desc(*D)) : 0;
  return a > b ? a : b;
}
size_t vlpet_k2_bwd_workspace_bytes(const VlpetK2Desc* D) {
  if (!D) return 0;
  const size_t a = generic_k2_bwd_ws(*D), b = use_fused_k2(*D) ? fused_k2_bwd_ws(*D) : 0;
  const size_t c = use_rows_k2(*D, true) ? rows_k1_bwd_ws(k2_rows_desc(*D)) : 0;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);   // dkv == NULL falls back to the generic path even when the shape qualifies
}
int vlpet_k2_is_fused(const VlpetK2Desc* D) { return !D ? 0 : (use_fused_k2(*D) ? 1 : (use_rows_k2(*D, true) ? 2 : 0)); }
int vlpet_k2_fwd(const VlpetK2Desc* D, const void* kv, const void* y, const VlpetK2Params* w, void* out, void* ws,
                 size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k2(D, w));
  if (!kv || !out) return fail(VLPET_E_BADARG, "k2_fwd: null activation pointer"); /* y may be NULL: no residual */
  if (y && use_fused_k2(*D) && aligned16(kv) && aligned16(y) && aligned16(out))
    return fused_k2_fwd(*D, kv, y, *w, out, static_cast<cudaStream_t>(stream));
  if (y && use_rows_k2(*D, false) && aligned16(kv) && aligned16(y) && aligned16(out))
    return rows_k1_fwd(k2_rows_desc(*D), y, kv, k2_as_k1_params(*w), out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (D->impl == VLPET_IMPL_FUSED) return fail(VLPET_E_UNSUPPORTED, "k2_fwd: fused kernel needs bf16, y != NULL, d %% 128 == 0, r %% 8 == 0, r <= 96");
  return generic_k2_fwd(*D, kv, y, *w, out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int vlpet_k2_bwd(const VlpetK2Desc* D, const void* kv, const void* dout, const VlpetK2Params* w, void* dkv,
                 const VlpetK2Grads* g, void* ws, size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k2(D, w));
  if (!kv || !dout || !g) return fail(VLPET_E_BADARG, "k2_bwd: null pointer");
  if (dkv && use_fused_k2(*D) && aligned16(kv) && aligned16(dout) && aligned16(dkv))
    return fused_k2_bwd(*D, kv, dout, *w, dkv, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (dkv && use_rows_k2(*D, true) && aligned16(kv) && aligned16(dout) && aligned16(dkv)) {
    VlpetK1Grads G;
    memset(&G, 0, sizeof(G));
    G.dWd = g->dWd; G.dbd = g->dbd; G.dWu = g->dWu; G.dbu = g->dbu;
    // x1 (the residual y) is not needed by the ungated backward and its gradient is dout itself: no dx1 output
    return rows_k1_bwd(k2_rows_desc(*D), kv, kv, dout, k2_as_k1_params(*w), nullptr, dkv, G, ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
  }
  if (D->impl == VLPET_IMPL_FUSED) return fail(VLPET_E_UNSUPPORTED, "k2_bwd: fused kernel needs bf16, dkv != NULL, d %% 128 == 0, r %% 8 == 0, r <= 96");
  return generic_k2_bwd(*D, kv, dout, *w, dkv, *g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- K3 ----------------------------------------------------------------------------------------------------
static int check_k3(const VlpetK3Desc* D, const VlpetK3Params* w) {
  if (!D || !w) return fail(VLPET_E_BADARG, "k3: null desc/params");
  if (D->M <= 0 || D->d <= 0 || D->F <= 0 || D->N <= 0 || D->V <= 0 || D->n_img <= 0)
    return fail(VLPET_E_BADARG, "k3: M, N, F, d, V, n_img must be positive");
  if (D->dtype != VLPET_F32 && D->dtype != VLPET_BF16) return fail(VLPET_E_BADARG, "k3: bad dtype %d", D->dtype);
  if (!w->Wf || !w->bf || !w->ln_f_w || !w->Wp || !w->bp || !w->ln_p_w || !w->E_img || !w->E_obj)
    return fail(VLPET_E_BADARG, "k3: weights missing");
  if (!D->rms && (!w->ln_f_b || !w->ln_p_b)) return fail(VLPET_E_BADARG, "k3: LayerNorm biases missing");
  return 0;
}
size_t vlpet_k3_fwd_workspace_bytes(const VlpetK3Desc* D) { return D ? generic_k3_fwd_ws(*D) : 0; }
size_t vlpet_k3_bwd_workspace_bytes(const VlpetK3Desc* D) { return D ? generic_k3_bwd_ws(*D) : 0; }
size_t vlpet_k3_save_floats(const VlpetK3Desc* D) { return D ? (size_t)D->M * D->d : 0; }
int vlpet_k3_fwd(const VlpetK3Desc* D, const void* feats, const void* pos, const int64_t* img_ids,
                 const int64_t* obj_ids, const VlpetK3Params* w, void* out, float* save, void* ws, size_t ws_bytes,
                 void* stream) {
  VLPET_TRY(check_k3(D, w));
  if (!feats || !pos || !out || !save) return fail(VLPET_E_BADARG, "k3_fwd: null pointer");
  return generic_k3_fwd(*D, feats, pos, img_ids, obj_ids, *w, out, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int vlpet_k3_bwd(const VlpetK3Desc* D, const void* feats, const void* pos, const int64_t* img_ids, const void* dout,
                 const VlpetK3Params* w, const float* save, void* dfeats, const VlpetK3Grads* g, void* ws,
                 size_t ws_bytes, void* stream) {
  VLPET_TRY(check_k3(D, w));
  if (!feats || !pos || !dout || !save || !g) return fail(VLPET_E_BADARG, "k3_bwd: null pointer");
  return generic_k3_bwd(*D, feats, pos, img_ids, dout, *w, save, dfeats, *g, ws, ws_bytes,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---- flat-bucket helpers -----------------------------------------------------------------------------------
namespace vlpet {
namespace flat {
// y = dropout(gelu_erf(x)) / dx = dy * mask/(1-p) * gelu_erf'(x); 8 bf16 per thread, mask from drop_hash4.
// The kernel was bound by erff() (~30 instructions per element: 2.3 TB/s); now Phi(v) comes from the Abramowitz-Stegun
// 7.1.26 rational form (|error| < 1.5e-7, far inside bf16), evaluated on packed fp32 pairs, and shares its exponential
// with the pdf of the backward:  E = exp(-v^2/2), t = 1/(1 + p|v|/sqrt2), h = E * t*(a1+t*(a2+...))/2,
// Phi(v) = 1/2 + copysign(1/2 - h, v), phi(v) = E/sqrt(2 pi).
template <bool BWD>
__global__ void __launch_bounds__(256) gelu_dropout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    __nv_bfloat16* __restrict__ out, int64_t nvec, uint64_t seed, const uint64_t* seed_dev,
                                    uint32_t thr16, float inv_keep) {
  using namespace ptx;
  const uint64_t s = seed + ((thr16 && seed_dev) ? *seed_dev : 0ull);
  const f2 kp = mk2(0.3275911f * 0.7071067811865476f, 0.3275911f * 0.7071067811865476f), one = mk2(1.f, 1.f);
  const f2 a1 = mk2(0.5f * 0.254829592f, 0.5f * 0.254829592f), a2 = mk2(0.5f * -0.284496736f, 0.5f * -0.284496736f),
           a3 = mk2(0.5f * 1.421413741f, 0.5f * 1.421413741f), a4 = mk2(0.5f * -1.453152027f, 0.5f * -1.453152027f),
           a5 = mk2(0.5f * 1.061405429f, 0.5f * 1.061405429f);
  const f2 ce = mk2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f), half = mk2(0.5f, 0.5f), mone = mk2(-1.f, -1.f);
  const f2 cpdf = mk2(0.3989422804014327f, 0.3989422804014327f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    uint32_t g[4] = {0, 0, 0, 0};
    if (BWD) {
      const uint4 qd = *reinterpret_cast<const uint4*>(dy + i * 8);
      g[0] = qd.x; g[1] = qd.y; g[2] = qd.z; g[3] = qd.w;
    }
    uint64_t h[2] = {0, 0};
    if (thr16) { h[0] = drop_hash4(s, (uint64_t)i * 2); h[1] = drop_hash4(s, (uint64_t)i * 2 + 1); }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t vlo = u[e] << 16, vhi = u[e] & 0xffff0000u;
      const f2 v = mk2u(vlo, vhi), av = mk2u(vlo & 0x7fffffffu, vhi & 0x7fffffffu);
      float t0, t1, e0, e1;
      un2(fma2(kp, av, one), t0, t1);
      un2(mul2(mul2(v, v), ce), e0, e1);
      float r0, r1, x0, x1;                               // MUFU.RCP / MUFU.EX2: 1-2 ulp, far inside the bf16 result
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x0) : "f"(e0));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x1) : "f"(e1));
      const f2 t = mk2(r0, r1), E = mk2(x0, x1);
      f2 poly = fma2(a5, t, a4);
      poly = fma2(poly, t, a3);
      poly = fma2(poly, t, a2);
      poly = fma2(poly, t, a1);
      const f2 hh = mul2(mul2(poly, t), E);               // (1 - erf(|v|/sqrt2)) / 2
      uint32_t c0, c1;
      un2u(fma2(hh, mone, half), c0, c1);                 // 1/2 - h  >= 0
      const f2 cdf = add2(half, mk2u(c0 | (vlo & 0x80000000u), c1 | (vhi & 0x80000000u)));
      f2 m = one;
      if (thr16) {
        const uint32_t two = (uint32_t)(h[e >> 1] >> (32 * (e & 1)));
        m = mk2(((two & 0xffffu) >= thr16) ? inv_keep : 0.f, ((two >> 16) >= thr16) ? inv_keep : 0.f);
      }
      f2 r;
      if (BWD) {
        const f2 d = bf2_to_f2(g[e]);
        r = mul2(mul2(d, m), fma2(mul2(v, cpdf), E, cdf));   // dy * mask * (Phi + v phi)
      } else {
        r = mul2(mul2(m, v), cdf);
      }
      o[e] = pack2(r);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Token cross-entropy over bf16 logits: one CTA per row, 8 logits per thread and iteration, online (max, sum) per
// thread, block reduction through shared memory.  exp2 with the max folded in: sum_j 2^((x_j - m) log2e).
constexpr int CE_THREADS = 256;
__device__ __forceinline__ void ce_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  s = s * exp2f((m - mm) * 1.4426950408889634f) + s2 * exp2f((m2 - mm) * 1.4426950408889634f);
  m = mm;
}
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ labels, float* __restrict__ loss,
                                                            float* __restrict__ lse, int ncols, int64_t ignore_index) {
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* x = logits + row * ld;
  float m = -3.0e38f, s = 0.f;
  for (int i = threadIdx.x; i < ncols / 8; i += CE_THREADS) {
    const uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(u[e] << 16); v[2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u); }
    float lm = v[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) lm = fmaxf(lm, v[e]);
    const float mm = fmaxf(m, lm);
    float acc = s * exp2f((m - mm) * 1.4426950408889634f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += exp2f((v[e] - mm) * 1.4426950408889634f);
    m = mm; s = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ce_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
  __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w) ce_merge(m, s, sm[w], ss[w]);
    const float l = m + logf(s);
    const int64_t lab = labels[row];
    lse[row] = l;
    loss[row] = (lab == ignore_index) ? 0.f : l - __bfloat162float(x[lab]);
  }
}
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ labels, const float* __restrict__ lse,
                                                            const float* __restrict__ dloss, __nv_bfloat16* __restrict__ dlogits,
                                                            int ncols, int64_t ignore_index) {
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* x = logits + row * ld;
  __nv_bfloat16* dx = dlogits + row * ld;
  const int64_t lab = labels[row];
  const float g = (lab == ignore_index) ? 0.f : dloss[row];
  const float l2 = lse[row] * 1.4426950408889634f;
  for (int i = threadIdx.x; i < ncols / 8; i += CE_THREADS) {
    const uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = i * 8 + 2 * e;
      float p0 = exp2f(fmaf(__uint_as_float(u[e] << 16), 1.4426950408889634f, -l2));
      float p1 = exp2f(fmaf(__uint_as_float(u[e] & 0xffff0000u), 1.4426950408889634f, -l2));
      if (j == lab) p0 -= 1.f;
      if (j + 1 == lab) p1 -= 1.f;
      __nv_bfloat162 t = __floats2bfloat162_rn(g * p0, g * p1);
      o[e] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one thread = 8 consecutive channels of one output cell; windows per PyTorch's adaptive rule
template <typename TI, typename TO>
__global__ void grid_maxpool_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t nimg, int g, int o, int F) {
  const int64_t nvec = (int64_t)F / 8;
  const int64_t total = nimg * o * o * nvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i % nvec;
    int64_t t = i / nvec;
    const int ox = (int)(t % o); t /= o;
    const int oy = (int)(t % o);
    const int64_t img = t / o;
    const int y0 = (oy * g) / o, y1 = ((oy + 1) * g + o - 1) / o;
    const int x0 = (ox * g) / o, x1 = ((ox + 1) * g + o - 1) / o;
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const TI* src = in + ((img * g * g + (int64_t)y * g + x) * F + v * 8);
        if constexpr (sizeof(TI) == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
          m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
        } else {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
          const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            m[2 * e] = fmaxf(m[2 * e], __uint_as_float(w[e] << 16));
            m[2 * e + 1] = fmaxf(m[2 * e + 1], __uint_as_float(w[e] & 0xffff0000u));
          }
        }
      }
    TO* dst = out + (((img * o + oy) * o + ox) * (int64_t)F + v * 8);
    if constexpr (sizeof(TO) == 4) {
      reinterpret_cast<float4*>(dst)[0] = make_float4(m[0], m[1], m[2], m[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(m[4], m[5], m[6], m[7]);
    } else {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 t2 = __floats2bfloat162_rn(m[2 * e], m[2 * e + 1]);
        w[e] = *reinterpret_cast<uint32_t*>(&t2);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
  // tail (n % 4) handled by the first threads of the grid
  int64_t tail0 = n / 4 * 4;
  int64_t t = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __float2bfloat16_rn(src[t]);
}

// AdamW as transformers.optimization.AdamW (used at trainer_base.py:633): m,v update; p -= lr*sqrt(bc2)/bc1 * m/(sqrt(v)+eps);
// then decoupled decay p -= lr*wd*p.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, const uint8_t* __restrict__ wd_mask, int64_t n, float lr, float b1,
                             float b2, float eps, float wd, float step_size, const float* __restrict__ gscale,
                             __nv_bfloat16* __restrict__ shadow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float gs = gscale ? *gscale : 1.0f;
  for (; i < n; i += stride) {
    float gi = g[i] * gs;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    float pi = p[i];
    pi -= step_size * mi / (sqrtf(vi) + eps);
    float w = wd_mask ? (wd_mask[i >> 10] ? wd : 0.0f) : wd;
    pi -= lr * w * pi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, const uint8_t* __restrict__ wd_mask, int64_t n,
                                 const float* __restrict__ hyper, float b1, float b2, float eps, float wd,
                                 const float* __restrict__ gscale, __nv_bfloat16* __restrict__ shadow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float gs = gscale ? *gscale : 1.0f;
  const float lr = hyper[0], step_size = hyper[1];
  for (; i < n; i += stride) {
    float gi = g[i] * gs;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    float pi = p[i];
    pi -= step_size * mi / (sqrtf(vi) + eps);
    float w = wd_mask ? (wd_mask[i >> 10] ? wd : 0.0f) : wd;
    pi -= lr * w * pi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i < n; i += stride) acc += x[i] * x[i];
  acc = warp_sum(acc);
  __shared__ float s[8];
  if (threadIdx.x % 32 == 0) s[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < blockDim.x / 32 ? s[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}
inline int flat_blocks(int64_t n, int per_thread) {
  int64_t b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}
}  // namespace flat
}  // namespace vlpet
using namespace vlpet::flat;

extern "C" {
static int gelu_dropout_launch(bool bwd, const void* x, const void* dy, void* out, int64_t n, float p_drop, uint64_t seed,
                               const uint64_t* seed_dev, void* stream) {
  if (!x || !out || (bwd && !dy) || n <= 0 || !(p_drop >= 0.f && p_drop < 1.f)) return fail(VLPET_E_BADARG, "gelu_dropout: bad arguments");
  if (n % 8 != 0) return fail(VLPET_E_UNSUPPORTED, "gelu_dropout: n must be a multiple of 8");
  if (!aligned16(x) || !aligned16(out) || (bwd && !aligned16(dy))) return fail(VLPET_E_ALIGN, "gelu_dropout: misaligned");
  const uint32_t thr16 = p_drop > 0.f ? drop_thr16(p_drop) : 0u;
  const float inv_keep = thr16 ? 1.0f / (1.0f - (float)thr16 / 65536.0f) : 1.0f;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bwd)
    gelu_dropout_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
                                                               static_cast<__nv_bfloat16*>(out), nvec, seed, seed_dev, thr16, inv_keep);
  else
    gelu_dropout_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), nullptr,
                                                                static_cast<__nv_bfloat16*>(out), nvec, seed, seed_dev, thr16, inv_keep);
  VLPET_LAUNCH_OK();
  return 0;
}
int vlpet_gelu_dropout_fwd(const void* x, void* y, int64_t n, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  return gelu_dropout_launch(false, x, nullptr, y, n, p_drop, seed, seed_dev, stream);
}
int vlpet_gelu_dropout_bwd(const void* x, const void* dy, void* dx, int64_t n, float p_drop, uint64_t seed,
                           const uint64_t* seed_dev, void* stream) {
  return gelu_dropout_launch(true, x, dy, dx, n, p_drop, seed, seed_dev, stream);
}

int vlpet_attn_fwd(const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, void* out, float* lse,
                   int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t causal, float p_drop, uint64_t seed,
                   const uint64_t* seed_dev, void* stream) {
  return attn_run(false, q, k, v, q_rs, k_rs, v_rs, out, lse, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, B, H, Lq, Lk,
                  causal, p_drop, seed, seed_dev, static_cast<cudaStream_t>(stream));
}
int vlpet_attn_bwd(const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, const void* out,
                   const void* dout, const float* lse, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs, int64_t dv_rs,
                   int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t causal, float p_drop, uint64_t seed,
                   const uint64_t* seed_dev, void* stream) {
  return attn_run(true, q, k, v, q_rs, k_rs, v_rs, nullptr, const_cast<float*>(lse), out, dout, dq, dk, dv, dq_rs, dk_rs, dv_rs, B, H,
                  Lq, Lk, causal, p_drop, seed, seed_dev, static_cast<cudaStream_t>(stream));
}

static int check_ce(const void* logits, int64_t ld, const void* labels, int64_t rows, int32_t ncols) {
  if (!logits || !labels || rows <= 0 || ncols <= 0 || ld < ncols) return fail(VLPET_E_BADARG, "ce: bad arguments");
  if (ncols % 8 != 0 || ld % 8 != 0) return fail(VLPET_E_UNSUPPORTED, "ce: ncols and ld must be multiples of 8 (ncols=%d)", ncols);
  if (!aligned16(logits)) return fail(VLPET_E_ALIGN, "ce: logits must be 16-byte aligned");
  if (rows > 0x7fffffff) return fail(VLPET_E_UNSUPPORTED, "ce: too many rows");
  return 0;
}
int vlpet_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, float* loss, float* lse, int64_t rows, int32_t ncols,
                 int64_t ignore_index, void* stream) {
  VLPET_TRY(check_ce(logits, ld, labels, rows, ncols));
  if (!loss || !lse) return fail(VLPET_E_BADARG, "ce_fwd: null output");
  flat::ce_fwd_kernel<<<(unsigned)rows, flat::CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), ld, labels, loss, lse, ncols, ignore_index);
  VLPET_LAUNCH_OK();
  return 0;
}
int vlpet_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* dloss, void* dlogits,
                 int64_t rows, int32_t ncols, int64_t ignore_index, void* stream) {
  VLPET_TRY(check_ce(logits, ld, labels, rows, ncols));
  if (!lse || !dloss || !dlogits || !aligned16(dlogits)) return fail(VLPET_E_BADARG, "ce_bwd: bad arguments");
  flat::ce_bwd_kernel<<<(unsigned)rows, flat::CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), ld, labels, lse, dloss, static_cast<__nv_bfloat16*>(dlogits), ncols, ignore_index);
  VLPET_LAUNCH_OK();
  return 0;
}

int vlpet_grid_maxpool(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t nimg, int32_t g, int32_t o,
                       int32_t F, void* stream) {
  if (!in || !out || nimg <= 0 || g <= 0 || o <= 0 || o > g || F <= 0) return fail(VLPET_E_BADARG, "grid_maxpool: bad arguments");
  if (F % 8 != 0) return fail(VLPET_E_UNSUPPORTED, "grid_maxpool: F must be a multiple of 8");
  if (!aligned16(in) || !aligned16(out)) return fail(VLPET_E_ALIGN, "grid_maxpool: buffers must be 16-byte aligned");
  if ((in_dtype != VLPET_F32 && in_dtype != VLPET_BF16) || (out_dtype != VLPET_F32 && out_dtype != VLPET_BF16))
    return fail(VLPET_E_BADARG, "grid_maxpool: bad dtype");
  const int64_t total = nimg * o * o * (F / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == VLPET_F32 && out_dtype == VLPET_BF16)
    grid_maxpool_kernel<float, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), nimg, g, o, F);
  else if (in_dtype == VLPET_F32)
    grid_maxpool_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(in), static_cast<float*>(out), nimg, g, o, F);
  else if (out_dtype == VLPET_BF16)
    grid_maxpool_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), nimg, g, o, F);
  else
    grid_maxpool_kernel<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), nimg, g, o, F);
  VLPET_LAUNCH_OK();
  return 0;
}

int vlpet_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return fail(VLPET_E_BADARG, "cast: bad arguments");
  if (n == 0) return 0;
  if (!aligned16(src) || (reinterpret_cast<uintptr_t>(dst) & 7u)) return fail(VLPET_E_ALIGN, "cast: misaligned buffers");
  cast_kernel<<<flat_blocks(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  VLPET_LAUNCH_OK();
  return 0;
}

int vlpet_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* wd_mask,
                     int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     const float* grad_scale_dev, void* bf16_shadow, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return fail(VLPET_E_BADARG, "adamw: bad arguments");
  if (n == 0) return 0;
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  float step_size = (float)(lr * sqrt(bc2) / bc1);
  adamw_kernel<<<flat_blocks(n, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, wd_mask, n, lr, beta1, beta2, eps, weight_decay, step_size, grad_scale_dev,
      static_cast<__nv_bfloat16*>(bf16_shadow));
  VLPET_LAUNCH_OK();
  return 0;
}

int vlpet_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* wd_mask,
                         int64_t n, const float* hyper_dev, float beta1, float beta2, float eps, float weight_decay,
                         const float* grad_scale_dev, void* bf16_shadow, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper_dev || n < 0) return fail(VLPET_E_BADARG, "adamw_dev: bad arguments");
  if (n == 0) return 0;
  adamw_dev_kernel<<<flat_blocks(n, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, wd_mask, n, hyper_dev, beta1, beta2, eps, weight_decay, grad_scale_dev,
      static_cast<__nv_bfloat16*>(bf16_shadow));
  VLPET_LAUNCH_OK();
  return 0;
}

int vlpet_sumsq(const float* x, int64_t n, float* out_dev, void* stream) {
  if (!x || !out_dev || n < 0) return fail(VLPET_E_BADARG, "sumsq: bad arguments");
  if (n == 0) return 0;
  sumsq_kernel<<<flat_blocks(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out_dev);
  VLPET_LAUNCH_OK();
  return 0;
}
}  // extern "C"
