/*
 * vlpet.h — C ABI of libvlpet.so: the B200-native (sm_100a) PET hot path of VL-PET.
 *
 * The reference (HenryHZY/VL-PET) is pure PyTorch: there is no native interface to mirror, so each entry
 * point below cites the reference Python lines whose op sequence it replaces (paths relative to
 * /root/reference/src).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer to a contiguous row-major tensor, 16-byte aligned.
 *  - nn.Linear weights are [out_features, in_features] row-major exactly as PyTorch stores them.
 *  - All calls are asynchronous on the caller's stream (`stream` is a cudaStream_t passed as void*),
 *    allocate nothing and keep no state besides an immutable per-process function-pointer cache.
 *    Scratch memory is caller-owned: query *_workspace_bytes first.
 *  - Return value: 0 = ok, >0 = cudaError_t, <0 = VLPET_E_* argument error.  vlpet_last_error() returns a
 *    thread-local human-readable message for the last non-zero return.  Nothing throws across the boundary.
 *  - Weight gradients are fp32 and are ACCUMULATED into (`+=`) the caller's buffers, so they can alias
 *    views of one flat all-reduce bucket.  Activation gradients are overwritten.
 *  - There is no CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef VLPET_H_
#define VLPET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VLPET_API __attribute__((visibility("default")))
#else
#define VLPET_API
#endif

#define VLPET_VERSION 100 /* major*100 + minor */

/* activation / weight element type of a call (weights, biases and activations share it) */
enum { VLPET_F32 = 0, VLPET_BF16 = 1 };

/* granularity-control matrix G (paper Fig. 2; my_transformers/modeling_bart.py:1195-1231) */
enum {
  VLPET_GATE_NONE = 0,     /* no gate: h = y1                                                           */
  VLPET_GATE_LARGE = 1,    /* N x d : sigmoid(GUp(gelu_new(GDown x1)))            modeling_bart.py:1195-1209 */
  VLPET_GATE_MIDDLE_X = 2, /* N x 1 : sigmoid(Linear(d,1)(x1 + y1))               modeling_bart.py:1219-1226 */
  VLPET_GATE_MIDDLE_Y = 3, /* 1 x d : h = y1 + y1*z (no sigmoid)                  modeling_bart.py:1227-1231 */
  VLPET_GATE_SMALL = 4     /* 1 x 1 : mean_L sigmoid(Linear(2d,1)([x1;y1]))       modeling_bart.py:1210-1218 */
};

/* error codes (negative returns) */
enum {
  VLPET_E_BADARG = -1,     /* null pointer / bad enum / non-positive size */
  VLPET_E_UNSUPPORTED = -2,/* shape outside what the kernels implement (never silently falls back to CPU) */
  VLPET_E_WORKSPACE = -3,  /* workspace too small */
  VLPET_E_NODEVICE = -4,   /* no sm_100 device / driver entry point missing */
  VLPET_E_ALIGN = -5       /* pointer not 16-byte aligned */
};

/* which implementation a call may use */
enum {
  VLPET_IMPL_AUTO = 0,     /* fused tcgen05/TMA kernel when the shape qualifies, else the generic kernels */
  VLPET_IMPL_GENERIC = 1,  /* shape-generic CUDA-core kernels (any d, r, gate, dtype)                     */
  VLPET_IMPL_FUSED = 2     /* fused sm_100a kernel only; VLPET_E_UNSUPPORTED if the shape does not qualify */
};

/* ---- K1: granularity-controlled PET module -------------------------------------------------------------
 * Replaces the inline op sequence of BartEncoderLayer.forward (my_transformers/modeling_bart.py:1145-1155
 * adapter, 1195-1231 gate, 1256-1260 scale+residual; FFN twin 1268-1278,1317-1325,1372-1376) and of
 * T5LayerSelfAttention.forward / T5LayerFF.forward (my_transformers/modeling_t5.py:777-824, 359-409):
 *
 *     y1  = kappa*x2 + alpha*(gelu_new(x2 Wd^T + bd) Wu^T + bu)
 *     out = x1 + dropout_p(s * gate(x1, y1))
 *
 * The multi-head down projection is passed as its row-concatenation Wd [r,d].            */
typedef struct VlpetK1Desc {
  int64_t M;        /* tokens = B*L                                                           */
  int32_t L;        /* sequence length; required (M % L == 0) for VLPET_GATE_SMALL, else may be 0 */
  int32_t d;        /* d_model                                                                */
  int32_t r;        /* adapter rank  (adapter_down_dim)                                       */
  int32_t rg;       /* gate rank     (adapter_gating_down_dim); VLPET_GATE_LARGE only         */
  int32_t gate;     /* VLPET_GATE_*                                                           */
  int32_t add_gate; /* --use_encoder_adapter_gating_add: h = y1 + G instead of y1 * G         */
  int32_t dtype;    /* VLPET_F32 | VLPET_BF16                                                 */
  int32_t impl;     /* VLPET_IMPL_*                                                           */
  float s;          /* encoder_gating_scaling_factor (1 when the flag is off)                 */
  float alpha;      /* encoder_adapter_scaling_factor                                         */
  float kappa;      /* encoder_x2_scaling_factor                                              */
  float p_drop;     /* dropout between gate and residual (modeling_bart.py:1259); 0 = identity (eval / parity) */
  uint64_t seed;    /* dropout stream: keep(m,c) is a pure function of (seed, m*d+c), so the backward regenerates
                       the forward mask from the same seed; the mask is NOT torch's Philox stream            */
  const uint64_t* seed_dev; /* optional DEVICE scalar added to `seed` when the kernel starts: lets a CUDA graph that
                       captured this call draw a fresh mask on every replay (bump it on the stream between replays;
                       forward and backward of one step must see the same value)                              */
} VlpetK1Desc;

typedef struct VlpetK1Params { /* dtype = desc.dtype; unused members NULL */
  const void* Wd;  /* [r,d]   attn_adapter_multihead_down.{h}.weight row-concatenated */
  const void* bd;  /* [r]                                                             */
  const void* Wu;  /* [d,r]   attn_adapter_multihead_up.weight                        */
  const void* bu;  /* [d]                                                             */
  const void* Gd;  /* [rg,d]  encoder_attn_adapter_gating_large_x_down.weight         */
  const void* gbd; /* [rg]                                                            */
  const void* Gu;  /* [d,rg]  encoder_attn_adapter_gating_large_x_up.weight           */
  const void* gbu; /* [d]                                                             */
  const void* gw;  /* [d] middle_x | [2d] small : gating Linear(.,1).weight           */
  const void* gb;  /* [1]                                                             */
  const void* gz;  /* [d] middle_y parameter                                          */
} VlpetK1Params;

typedef struct VlpetK1Grads { /* fp32, accumulated into; members mirror VlpetK1Params; NULL = skip */
  float *dWd, *dbd, *dWu, *dbu, *dGd, *dgbd, *dGu, *dgbu, *dgw, *dgb, *dgz;
} VlpetK1Grads;

VLPET_API size_t vlpet_k1_fwd_workspace_bytes(const VlpetK1Desc* desc);
VLPET_API size_t vlpet_k1_bwd_workspace_bytes(const VlpetK1Desc* desc);
VLPET_API int vlpet_k1_fwd(const VlpetK1Desc* desc, const void* x1, const void* x2, const VlpetK1Params* w, void* out,
                 void* workspace, size_t workspace_bytes, void* stream);
/* Recomputes the forward intermediates from x1/x2 (nothing is saved by the forward). */
VLPET_API int vlpet_k1_bwd(const VlpetK1Desc* desc, const void* x1, const void* x2, const void* dout,
                 const VlpetK1Params* w, void* dx1, void* dx2, const VlpetK1Grads* g, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Which path vlpet_k1_fwd / _bwd take for this desc under VLPET_IMPL_AUTO: 1 = the fused tcgen05 kernels (large gate, ungated
 * form), 2 = the row-wise gate kernels (middleX / middleY / small gates -- composed with the tcgen05 adapter kernel at tensor-core
 * ranks -- and ranks <= 16 in one launch; csrc/vlpet_rows.cu), 3 = large gate with 96 < max(r, rg) <= 192 (the rank of the
 * reference's T5 scripts) composed from the ungated tcgen05 kernels per rank half + one element-wise gate kernel
 * (csrc/vlpet_wide.cu), 0 = the generic CUDA-core path (fp32, odd shapes) */
VLPET_API int vlpet_k1_fwd_is_fused(const VlpetK1Desc* desc);
VLPET_API int vlpet_k1_bwd_is_fused(const VlpetK1Desc* desc);

/* ---- K2: decoder cross-attention value parallel adapter -----------------------------------------------
 * Replaces AdapterController.forward(inputs, task, y) + Adapter.forward (adapters/adapter_controller.py:131-162,
 * adapters/adapter_modeling.py:55-61) as called at my_transformers/modeling_bart.py:427-430 and
 * my_transformers/modeling_t5.py:600-603:     out = y + sf * (gelu_new(kv Wd^T + bd) Wu^T + bu)
 * (pass y = kv for the non-parallel residual of adapter_controller.py:160-161).                     */
typedef struct VlpetK2Desc {
  int64_t M;
  int32_t d, r;
  int32_t dtype, impl;
  float sf; /* scaling_factor (1 when use_scaling_factor is off) */
} VlpetK2Desc;
typedef struct VlpetK2Params { const void *Wd, *bd, *Wu, *bu; } VlpetK2Params;
typedef struct VlpetK2Grads { float *dWd, *dbd, *dWu, *dbu; } VlpetK2Grads;

VLPET_API size_t vlpet_k2_fwd_workspace_bytes(const VlpetK2Desc* desc);
VLPET_API size_t vlpet_k2_bwd_workspace_bytes(const VlpetK2Desc* desc);
VLPET_API int vlpet_k2_fwd(const VlpetK2Desc* desc, const void* kv, const void* y, const VlpetK2Params* w, void* out,
                 void* workspace, size_t workspace_bytes, void* stream);
/* 1 if vlpet_k2_fwd (with y != NULL) / vlpet_k2_bwd (with dkv != NULL) run the fused tcgen05 kernels for this desc */
VLPET_API int vlpet_k2_is_fused(const VlpetK2Desc* desc);
/* dkv receives ONLY the adapter-path gradient (autograd adds the frozen k/v_proj paths); dy == dout. */
VLPET_API int vlpet_k2_bwd(const VlpetK2Desc* desc, const void* kv, const void* dout, const VlpetK2Params* w, void* dkv,
                 const VlpetK2Grads* g, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3: visual projection -----------------------------------------------------------------------------
 * Replaces VisualEmbedding.forward (src/modeling_bart.py:143-192; T5: src/modeling_t5.py:124-174):
 *     out = LN(feats Wf^T + bf) + LN([pos,area] Wp^T + bp) + E_img[img_ids] + E_obj[V-1-obj_ids]
 * rms = 1 selects T5LayerNorm (my_transformers/modeling_t5.py:235-252: no mean, no bias).           */
typedef struct VlpetK3Desc {
  int64_t M;      /* B*N visual tokens                                                     */
  int32_t N;      /* boxes per sample (ids default to 0 / arange(N) when the id pointers are NULL) */
  int32_t F;      /* feat_dim                                                              */
  int32_t d;
  int32_t V;      /* rows of E_obj (vocabulary size)                                       */
  int32_t n_img;  /* rows of E_img                                                         */
  int32_t rms;
  int32_t dtype, impl;
  float eps;
} VlpetK3Desc;
typedef struct VlpetK3Params {
  const void *Wf, *bf, *ln_f_w, *ln_f_b; /* [d,F],[d],[d],[d] (ln_*_b NULL when rms)     */
  const void *Wp, *bp, *ln_p_w, *ln_p_b; /* [d,5],[d],[d],[d]                            */
  const void *E_img, *E_obj;             /* [n_img,d], [V,d]                             */
} VlpetK3Params;
typedef struct VlpetK3Grads {
  float *dWf, *dbf, *dln_f_w, *dln_f_b, *dWp, *dbp, *dln_p_w, *dln_p_b, *dE_img;
} VlpetK3Grads;

VLPET_API size_t vlpet_k3_fwd_workspace_bytes(const VlpetK3Desc* desc);
VLPET_API size_t vlpet_k3_bwd_workspace_bytes(const VlpetK3Desc* desc);
/* pos [M,4] (x1,x2,y1,y2); img_ids / obj_ids int64 [M] or NULL.
 * save (fp32, >= vlpet_k3_save_floats(desc) floats) keeps the pre-norm projections for the backward. */
VLPET_API size_t vlpet_k3_save_floats(const VlpetK3Desc* desc);
VLPET_API int vlpet_k3_fwd(const VlpetK3Desc* desc, const void* feats, const void* pos, const int64_t* img_ids,
                 const int64_t* obj_ids, const VlpetK3Params* w, void* out, float* save, void* workspace,
                 size_t workspace_bytes, void* stream);
VLPET_API int vlpet_k3_bwd(const VlpetK3Desc* desc, const void* feats, const void* pos, const int64_t* img_ids,
                 const void* dout, const VlpetK3Params* w, const float* save, void* dfeats /* may be NULL */,
                 const VlpetK3Grads* g, void* workspace, size_t workspace_bytes, void* stream);

/* ---- LayerNorm behind the encoder PET sites ---------------------------------------------
 * The nn.LayerNorm the reference applies to the K1 output (my_transformers/modeling_bart.py:1260-1261, 1376-1377), for the
 * training configuration: bf16 activations x / y / dy / dx [M, d], fp32 affine parameters (trainable under
 * --unfreeze_encoder_layer_norms) and fp32 row statistics mean / rstd [M] saved by the forward.  dw / db are
 * accumulated into (may be NULL for a frozen LayerNorm).  Requires d % 256 == 0, 256 <= d <= 1024.        */
VLPET_API int vlpet_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M,
                        int32_t d, float eps, int32_t dtype, void* stream);
VLPET_API int vlpet_layernorm_bwd(const void* x, const void* dy, const float* w, const float* mean, const float* rstd, void* dx,
                        float* dw, float* db, int64_t M, int32_t d, int32_t dtype, void* stream);

/* y = LayerNorm(res + dropout_p(h)) in one pass: the post-LN residual step of the frozen decoder blocks -- `F.dropout` ->
 * `residual +` -> `LayerNorm` (my_transformers/modeling_bart.py:1663-1665, 1683-1685, 1697-1699), three kernels per sublayer in
 * the reference.  bf16 activations, fp32 affine parameters and statistics; xs receives res + dropout(h) (bf16, rounded as the
 * unfused sequence stores it) for the backward; the mask is the counter-based stream (seed + *seed_dev), regenerated by the
 * backward: dres = LayerNorm backward, dh = dres * mask / (1 - p); dw / db accumulated into (NULL: frozen LayerNorm).      */
VLPET_API int vlpet_dropout_add_layernorm_fwd(const void* h, const void* res, const float* w, const float* b, void* y, void* xs,
                        float* mean, float* rstd, int64_t M, int32_t d, float eps, float p_drop, uint64_t seed,
                        const uint64_t* seed_dev, void* stream);
VLPET_API int vlpet_dropout_add_layernorm_bwd(const void* xs, const void* dy, const float* w, const float* mean, const float* rstd,
                        void* dres, void* dh, float* dw, float* db, int64_t M, int32_t d, float p_drop, uint64_t seed,
                        const uint64_t* seed_dev, void* stream);

/* ---- FFN activation of the frozen blocks around the PET sites ---------------------------
 * y = dropout_p(gelu(x)) in one pass and its backward dx = dy * mask/(1-p) * gelu'(x) in one pass, replacing the
 * reference's `activation_fn(fc1(h))` + `F.dropout(.., p=activation_dropout)` pair (my_transformers/modeling_bart.py:
 * 1264-1266) and the two autograd kernels behind it.  gelu is the exact erf form (ACT2FN["gelu"]); the dropout mask is the
 * counter-based stream of K1 (never stored: the backward regenerates it from seed + *seed_dev).  bf16, n % 8 == 0.     */
VLPET_API int vlpet_gelu_dropout_fwd(const void* x, void* y, int64_t n, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                           void* stream);
VLPET_API int vlpet_gelu_dropout_bwd(const void* x, const void* dy, void* dx, int64_t n, float p_drop, uint64_t seed,
                           const uint64_t* seed_dev, void* stream);

/* ---- short-sequence attention of the frozen blocks -------------------------------------
 * out = dropout_p(softmax(q k^T / 8 [+ causal mask])) v per (batch, head), head_dim = 64, bf16, Lq, Lk <= 128, no padding
 * mask -- the bmm / softmax / dropout / bmm sequence of BartAttention.forward (my_transformers/modeling_bart.py:143-280)
 * at the sequence lengths of the VL-PET workloads, where the library flash kernels are overhead-bound.  One CTA per
 * (batch, head), the whole score tile in shared memory.  q / k / v: row i of batch b at base + (b*L + i)*rs elements, head h
 * at +64h inside the row (so the three thirds of a fused [B, L, 3*H*64] projection, or three [B, L, H*64] tensors).
 * out / dout: [B, L, H*64] contiguous; dq / dk / dv: rows with strides dq_rs / dk_rs / dv_rs elements (H*64 for separate
 * tensors, 3*H*64 for the thirds of one fused [B, L, 3*H*64] gradient buffer); lse: [B, H, Lq] fp32 (saved for the backward).
 * The dropout mask is a counter-based stream (seed + *seed_dev), regenerated by the backward.                          */
VLPET_API int vlpet_attn_fwd(const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, void* out,
                   float* lse, int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t causal, float p_drop, uint64_t seed,
                   const uint64_t* seed_dev, void* stream);
VLPET_API int vlpet_attn_bwd(const void* q, const void* k, const void* v, int64_t q_rs, int64_t k_rs, int64_t v_rs, const void* out,
                   const void* dout, const float* lse, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs, int64_t dv_rs,
                   int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t causal, float p_drop, uint64_t seed,
                   const uint64_t* seed_dev, void* stream);

/* ---- token cross-entropy of the LM head (frozen decoder output) -------------------------
 * loss[i] = logsumexp_j(logits[i, j]) - logits[i, labels[i]] in fp32 straight from the bf16 logits (0, and a zero gradient
 * row, where labels[i] == ignore_index), replacing `lm_logits.float()` + `CrossEntropyLoss(ignore_index=-100,
 * reduction='none')` of the reference (src/modeling_bart.py:1585-1586): for a caption batch the logits are [10 000 x 50 465],
 * and the fp32 copy + softmax forward/backward moved ~10 GB per step.  Forward: one pass, online max/sum per row, also
 * writes lse[i] for the backward.  Backward: dlogits[i, j] = dloss[i] * (exp(logits[i, j] - lse[i]) - [j == labels[i]]) as
 * bf16 (may alias `logits`: every element is read once before it is written by the same thread).
 * logits / dlogits: [rows, ncols] bf16 with row pitch `ld` elements; ncols % 8 == 0, ld % 8 == 0, 16-byte aligned.   */
VLPET_API int vlpet_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, float* loss, float* lse, int64_t rows,
                 int32_t ncols, int64_t ignore_index, void* stream);
VLPET_API int vlpet_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* dloss,
                 void* dlogits, int64_t rows, int32_t ncols, int64_t ignore_index, void* stream);

/* ---- CLIP-grid downsample feeding K3 --------------------------------------------------------------------
 * Replaces Downsample.downsample_inputs (src/modeling_bart.py:566-583: permute -> [B, F, g, g] -> AdaptiveMaxPool2d((o, o))
 * -> permute back) for the pre-extracted grid features of the VL-PET scripts (--n_boxes 36 --downsample: 7x7 -> 6x6),
 * fused with the cast to the compute dtype: in [nimg, g*g, F] (fp32 or bf16) -> out [nimg, o*o, F] (fp32 or bf16).
 * Max-pooling commutes with the (monotonic) bf16 rounding, so the result equals pooling in fp32 and casting after.
 * Window of output cell i along one axis: [floor(i*g/o), ceil((i+1)*g/o)) -- PyTorch's adaptive pooling rule.
 * F must be a multiple of 8.  Features are inputs: there is no backward.                                    */
VLPET_API int vlpet_grid_maxpool(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t nimg, int32_t g,
                       int32_t o, int32_t F, void* stream);

/* ---- K3-LR: the PET-shaped visual projector (flag --use_lowrank_visual_projector) -----------------------
 * Replaces LowRankVisualEmbedding.forward (src/modeling_bart.py:263-334):
 *     e   = Up(gelu_new(Down feats))                      Down = row-concatenated multi-head Linear(F, r/h)
 *     e   = e * G   |   e + e * G                         G = sigmoid(GUp(gelu_new(GDown feats)))   (gated / residual flags)
 *     out = LN(e) + LN([pos,area] Wp^T + bp) + E_img[img_ids] + E_obj[V-1-obj_ids]
 * Shape-generic CUDA-core implementation (no shipped VL-PET script enables the flag).  `save` as for K3.       */
typedef struct VlpetK3LRDesc {
  int64_t M;
  int32_t N, F, d, r, rg, V, n_img;
  int32_t gated;    /* use_visual_projector_gating_large_x_lowrank                */
  int32_t residual; /* use_visual_projector_residual_connection: e + e*G          */
  int32_t dtype, impl;
  float eps;
} VlpetK3LRDesc;
typedef struct VlpetK3LRParams {
  const void *Wd, *bd, *Wu, *bu, *Gd, *gbd, *Gu, *gbu; /* [r,F],[r],[d,r],[d],[rg,F],[rg],[d,rg],[d] */
  const void *ln_f_w, *ln_f_b, *Wp, *bp, *ln_p_w, *ln_p_b, *E_img, *E_obj;
} VlpetK3LRParams;
typedef struct VlpetK3LRGrads {
  float *dWd, *dbd, *dWu, *dbu, *dGd, *dgbd, *dGu, *dgbu, *dln_f_w, *dln_f_b, *dWp, *dbp, *dln_p_w, *dln_p_b, *dE_img;
} VlpetK3LRGrads;
VLPET_API size_t vlpet_k3lr_fwd_workspace_bytes(const VlpetK3LRDesc* desc);
VLPET_API size_t vlpet_k3lr_bwd_workspace_bytes(const VlpetK3LRDesc* desc);
VLPET_API int vlpet_k3lr_fwd(const VlpetK3LRDesc* desc, const void* feats, const void* pos, const int64_t* img_ids,
                   const int64_t* obj_ids, const VlpetK3LRParams* w, void* out, float* save /* M*d floats */,
                   void* workspace, size_t workspace_bytes, void* stream);
VLPET_API int vlpet_k3lr_bwd(const VlpetK3LRDesc* desc, const void* feats, const void* pos, const int64_t* img_ids,
                   const void* dout, const VlpetK3LRParams* w, const float* save, const VlpetK3LRGrads* g, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- token-contracted weight-gradient GEMM (building block of the fused backward) -----------------------
 * out_p[c, n] += scale_p * sum_tok A_p[tok, c] * B_p[tok, n]   (transposed_p: out_p[n, c] instead), up to 4 pairs per
 * launch, bf16 operands, fp32 accumulate on the tensor cores (tcgen05), fp32 reductions into out.  These are the dW
 * GEMMs the reference's autograd runs as separate addmm calls for every nn.Linear of a PET site
 * (my_transformers/modeling_bart.py:1045-1056).  If nb_valid == nout + 1, column nout of B must be
 * all ones and bias_p[c] += scale_p * sum_tok A_p[tok, c].  Requires d % 128 == 0, 8 <= nout <= 127.            */
typedef struct VlpetWgradPair {
  const void* A;     /* [Mtok, d] bf16, row pitch lda elements (multiple of 8)            */
  int64_t lda;
  const void* B;     /* [Mtok, nb_valid] bf16, row pitch ldb elements (multiple of 8)     */
  int64_t ldb;
  int32_t nb_valid;  /* nout or nout + 1                                                  */
  int32_t transposed;
  float* out;        /* fp32 [d, nout] or [nout, d], accumulated into                     */
  float* bias;       /* fp32 [d] or NULL                                                  */
  float scale;
} VlpetWgradPair;
VLPET_API int vlpet_wgrad_bf16(const VlpetWgradPair* pairs, int32_t npairs, int64_t Mtok, int32_t d, int32_t nout, void* stream);

/* ---- PET parameter/gradient bucket helpers (rows a9 / C1 of SURVEY §8) -------------------------------- */
/* dst_bf16[i] = bf16(src_f32[i]) : one launch refreshes the bf16 shadow of the whole flat PET bucket.    */
VLPET_API int vlpet_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* Fused AdamW on a flat fp32 bucket (transformers.optimization.AdamW semantics used at trainer_base.py:633:
 * decoupled weight decay, bias correction on); wd_mask (uint8 per 1024-element block, may be NULL = all
 * decayed) selects the no-decay group ("bias", "LayerNorm.weight": trainer_base.py:640-660).
 * grad_scale multiplies the gradient first (1/world_size and/or the clip factor).                        */
VLPET_API int vlpet_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* wd_mask,
                     int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     const float* grad_scale_dev /* device scalar or NULL */, void* bf16_shadow /* or NULL */,
                     void* stream);
/* Same step with the step-dependent scalars read from DEVICE memory, so the launch can live in a CUDA graph:
 * hyper_dev[0] = lr, hyper_dev[1] = lr * sqrt(1 - beta2^t) / (1 - beta1^t).                                   */
VLPET_API int vlpet_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* wd_mask,
                         int64_t n, const float* hyper_dev, float beta1, float beta2, float eps, float weight_decay,
                         const float* grad_scale_dev, void* bf16_shadow, void* stream);
/* sum of squares of a flat fp32 buffer accumulated into *out_dev (device scalar; caller zeroes it)       */
VLPET_API int vlpet_sumsq(const float* x, int64_t n, float* out_dev, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------------- */
VLPET_API int vlpet_version(void);
VLPET_API const char* vlpet_last_error(void);
/* number of kernels this library has launched in this process (all streams); for bench.py's gpu_launches */
VLPET_API uint64_t vlpet_launch_count(void);
/* device info the host side needs: returns 0 and fills sm_count / cc_major / cc_minor */
VLPET_API int vlpet_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

#ifdef __cplusplus
}
#endif
#endif /* VLPET_H_ */
