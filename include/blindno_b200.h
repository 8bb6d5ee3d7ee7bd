/*
 * blindno_b200.h -- C ABI of the B200 (sm_100a) NIO-FNO hot path.
 *
 * The reference (yl602019618/Reconstruction-of-PDE-without-Time-Label) is pure
 * Python and has no FFI of its own: its "operator interface" for this path is
 * the forward() of a handful of nn.Modules.  Each entry point below replaces
 * the body of one of them (file:line relative to the reference tree) and is
 * what a ctypes / cffi / pybind stub on the reference side binds (see
 * INTEGRATION.md).  Plain pointers and sizes only, no torch types, no
 * exceptions across the boundary.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all tensors are dense fp32; "complex" means interleaved (re, im) pairs,
 *     which is both torch.complex64 and the reference's [..., 2] float layout;
 *   - activations inside an FNO are channels-first and zero padded:
 *     [images, width, Hp, Wp] (1-D: Hp = 1);
 *   - inputs are borrowed, outputs and workspaces are caller-owned;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and
 *     returns without synchronising;
 *   - return value 0 = ok, negative = BdnStatus; bdn_last_error() gives text.
 *   - re-entrant: the only shared state is the per-device plan cache (DFT
 *     tables), guarded by a mutex.
 */
#ifndef BLINDNO_B200_H
#define BLINDNO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDN_MAX_LAYERS 8
#define BDN_ABI_VERSION 2

typedef enum BdnStatus {
  BDN_OK = 0,
  BDN_ERR_INVALID = -1,     /* bad shape / null pointer / unsupported size      */
  BDN_ERR_CUDA = -2,        /* a CUDA runtime call failed                         */
  BDN_ERR_WORKSPACE = -3,   /* workspace too small                                */
  BDN_ERR_UNSUPPORTED = -4  /* valid request this build cannot serve              */
} BdnStatus;

/* Arithmetic of the DFT GEMMs.
 *   FP32   = CUDA-core FFMA kernels.
 *   TF32X3 = tcgen05 tensor cores, every operand split into a TF32 high part and the TF32-rounded
 *            remainder, three MMAs per K step (lo*hi + hi*lo + hi*hi) accumulated in fp32 in tensor
 *            memory: fp32-level accuracy, meets the 1e-5 bound.  The default of the 2-D nets.
 *   TF32   = one MMA per K step (bound 2e-3 on outputs, 1e-2 on gradients).
 * 2-D nets run the fused tensor-core layer kernels (all four DFT GEMMs of a layer on tcgen05) when the
 * shape fits their shared-memory plan; bdn_fno_layer_path() tells.  A shape that does not fit runs the
 * FFMA layer kernels (W-forward stage alone on tcgen05) and says so once on stderr.  1-D nets: the
 * W-forward stage only. */
typedef enum BdnPrecision {
  BDN_PREC_FP32 = 0,
  BDN_PREC_TF32 = 1,
  BDN_PREC_TF32X3 = 2
} BdnPrecision;

/* ---------------------------------------------------------------------------
 * Shape of one spectral convolution.
 *   2-D: SpectralConv2d.forward  2d_FPE/FNOModules.py:156-178 (+ compl_mul2d :141-154)
 *   1-D: SpectralConv1d.forward  1d_FPE/FNOModules.py:47-59   (hp = 1, m1 = 0, DC bin * 0.5)
 * ------------------------------------------------------------------------- */
typedef struct BdnSpectralShape {
  int32_t ndim;      /* 1 or 2                                                    */
  int32_t images;    /* B' = batch (x bag size when snapshots are folded in)      */
  int32_t c_in;      /* input channels                                            */
  int32_t c_out;     /* output channels                                           */
  int32_t hp, wp;    /* transformed (already padded) extents; hp = 1 in 1-D       */
  int32_t m1, m2;    /* kept modes: rows {0..m1-1} u {hp-m1..hp-1}, cols 0..m2-1;
                        1-D: m1 = 0 and m2 = modes1                               */
  int32_t prec;      /* BdnPrecision                                              */
} BdnSpectralShape;

/* y = spectral(x).  w1/w2: [c_in, c_out, m1, m2] complex (w2 NULL in 1-D, where
 * w1 is [c_in, c_out, m2]).  xs_saved (optional, may be NULL): receives the kept
 * spectrum of x, [images, c_in, K, m2] complex with K = 2*m1 (1-D: 1), which
 * bdn_spectral_backward needs.  ws: bdn_spectral_workspace_bytes() bytes. */
size_t bdn_spectral_workspace_bytes(const BdnSpectralShape* s);
int bdn_spectral_forward(const BdnSpectralShape* s, const float* x, const float* w1, const float* w2,
                         float* y, float* xs_saved, void* ws, size_t ws_bytes, void* stream);
/* gx (may be NULL), gw1, gw2 are OVERWRITTEN (not accumulated). */
int bdn_spectral_backward(const BdnSpectralShape* s, const float* gy, const float* xs_saved,
                          const float* w1, const float* w2, float* gx, float* gw1, float* gw2,
                          void* ws, size_t ws_bytes, void* stream);

/* One stage on its own, for stage-level parity tests and kernel benchmarks: the pruned forward
 * DFT along W of `rows` rows (the rfft of 1d_FPE/FNOModules.py:50 / the W pass of rfft2,
 * 2d_FPE/FNOModules.py:163, kept bins only).  x: [rows, wp]; out: [rows, m2] complex.
 * act != 0 applies the exact GELU on load.  prec = BDN_PREC_TF32 runs the tcgen05 kernel. */
int bdn_stage_wfwd(int32_t hp, int32_t wp, int32_t m1, int32_t m2, int32_t rows, const float* x, float* out,
                   int32_t act, int32_t prec, void* stream);

/* ---------------------------------------------------------------------------
 * A whole FNO net: FNO1d.forward 1d_FPE/FNOModules.py:99-122,
 *                  FNO2d.forward 2d_FPE/FNOModules.py:218-240.
 * lift fc0 -> zero pad -> n_layers x [spectral + 1x1 conv (+ exact GELU except
 * last)] -> crop -> fc1 -> GELU -> fc2.
 * ------------------------------------------------------------------------- */
typedef struct BdnFnoShape {
  int32_t ndim;            /* 1 or 2                                              */
  int32_t images;          /* B'                                                  */
  int32_t c_in;            /* fc0 input features                                  */
  int32_t width;           /* channel width C                                     */
  int32_t c_out;           /* fc2 output features (FNO2d: always 1, Q3)           */
  int32_t hidden;          /* fc1 output features (128 in the reference)          */
  int32_t n_layers;        /* <= BDN_MAX_LAYERS                                   */
  int32_t h, w;            /* unpadded grid (1-D: h = 1, w = N)                   */
  int32_t hp, wp;          /* padded grid: h + round(h/4), w + round(w/4)         */
  int32_t out_h, out_w;    /* cropped grid the projection runs on (Q4: the
                              reference crops H by the W pad and W by the H pad)  */
  int32_t m1, m2;          /* kept modes (1-D: m1 = 0)                            */
  int32_t prec;            /* BdnPrecision                                        */
} BdnFnoShape;

typedef struct BdnFnoParams {      /* names = the reference state_dict keys       */
  const float* fc0_w;              /* fc0.weight [width, c_in]                    */
  const float* fc0_b;              /* fc0.bias   [width]                          */
  const float* conv_w[BDN_MAX_LAYERS];   /* conv_list.k.weight [width, width(,1,1)] */
  const float* conv_b[BDN_MAX_LAYERS];   /* conv_list.k.bias   [width]              */
  const float* spec_w1[BDN_MAX_LAYERS];  /* spectral_list.k.weights1 (complex)      */
  const float* spec_w2[BDN_MAX_LAYERS];  /* spectral_list.k.weights2 (2-D only)     */
  const float* fc1_w;              /* fc1.weight [hidden, width]                  */
  const float* fc1_b;              /* fc1.bias   [hidden]                         */
  const float* fc2_w;              /* fc2.weight [c_out, hidden]                  */
  const float* fc2_b;              /* fc2.bias   [c_out]                          */
} BdnFnoParams;

typedef struct BdnFnoGrads {       /* same shapes as BdnFnoParams; ACCUMULATED into
                                      (+=), so the caller zeroes them (this is what
                                      lets weight-grad kernels write straight into a
                                      flat all-reduce buffer)                        */
  float* fc0_w; float* fc0_b;
  float* conv_w[BDN_MAX_LAYERS]; float* conv_b[BDN_MAX_LAYERS];
  float* spec_w1[BDN_MAX_LAYERS]; float* spec_w2[BDN_MAX_LAYERS];
  float* fc1_w; float* fc1_b; float* fc2_w; float* fc2_b;
} BdnFnoGrads;

/* Where the lift reads its input from.
 *   x_cl != NULL : channels-last [images, h, w, c_in]               (FNO heads)
 *   x_cl == NULL : the NIO-FNO per-snapshot input, never materialised:
 *                  image (b, l) = concat(bags[b, idx[l]], grid), c_in = 1 + grid_dim
 *                  (NIOFP2D_FNO.forward 2d_FPE/NIOModules.py:548-560,
 *                   NIOFP_FNO.forward   1d_FPE/NIOModules.py:124-135)            */
typedef struct BdnLiftInput {
  const float* x_cl;
  const float* bags;       /* [n_bags, bag_len, h, w]                             */
  const int32_t* idx;      /* [n_keep] snapshot indices, NULL = identity          */
  const float* grid;       /* [h, w, grid_dim]                                    */
  int32_t n_bags, bag_len, n_keep, grid_dim;   /* images == n_bags * n_keep       */
} BdnLiftInput;

/* Buffers kept from forward to backward (caller-owned, sizes below). */
size_t bdn_fno_act_floats(const BdnFnoShape* s);    /* z: (n_layers+1) x [images,width,hp,wp] */
size_t bdn_fno_spec_floats(const BdnFnoShape* s);   /* xs: n_layers x [images,width,K,m2,2]
                                                        (+ n_layers x [m2,K,width,width,2]: mode-major
                                                        weight copy, few-image 2-D nets only)      */
size_t bdn_fno_workspace_bytes(const BdnFnoShape* s);

/* out: [images, out_h, out_w, c_out] channels-last.  z_saved / xs_saved may be
 * NULL for inference (then ws must be bdn_fno_workspace_bytes() + room for two
 * activations, which bdn_fno_workspace_bytes already includes). */
int bdn_fno_forward(const BdnFnoShape* s, const BdnFnoParams* p, const BdnLiftInput* in,
                    float* out, float* z_saved, float* xs_saved,
                    void* ws, size_t ws_bytes, void* stream);

/* g_out: [images, out_h, out_w, c_out]; or, when pooled_g != 0, the gradient of
 * the bag mean [n_bags, out_h, out_w, c_out] which every snapshot of the bag
 * receives scaled by 1/n_keep (backward of the mean in bag_pool_lift).
 * gx_cl: [images, h, w, c_in] or NULL (the raw bags need no gradient). */
int bdn_fno_backward(const BdnFnoShape* s, const BdnFnoParams* p, const BdnLiftInput* in,
                     const float* g_out, int32_t pooled_g, int32_t n_keep,
                     const float* z_saved, const float* xs_saved,
                     const BdnFnoGrads* grads, float* gx_cl,
                     void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Single stages of an FNO net (what the torch custom ops fno_lift_pad / fno_layer{1,2}d /
 * fno_project bind).  They take the BdnFnoShape of the net they belong to (n_layers is only
 * range-checked) and run the same kernels bdn_fno_forward / bdn_fno_backward chain.
 * Activations are PRE-activation tensors z_k [images, width, hp, wp]; a layer applies the
 * exact GELU to its input on load when act_in != 0, so the reference's
 *   x_{k+1} = gelu(spectral(x_k) + conv1x1(x_k))        2d_FPE/FNOModules.py:226-232
 *                                                        1d_FPE/FNOModules.py:108-114
 * is  z_{k+1} = layer(z_k, act_in = (k > 0)),  x_k = gelu(z_k) for k > 0,  x_0 = z_0.
 * All g_* gradient pointers are ACCUMULATED into (+=): the caller zeroes them.
 * ------------------------------------------------------------------------- */
/* lift: fc0 + channels-first + zero pad   FNOModules.py:103-106 (1-D), :219-224 (2-D) */
int bdn_stage_lift_forward(const BdnFnoShape* s, const float* fc0_w, const float* fc0_b,
                           const BdnLiftInput* in, float* z0, void* stream);
int bdn_stage_lift_backward(const BdnFnoShape* s, const float* fc0_w, const float* fc0_b,
                            const BdnLiftInput* in, const float* gz0,
                            float* g_fc0_w, float* g_fc0_b, float* gx_cl /* may be NULL */, void* stream);
/* one layer body.  xs_saved (may be NULL in forward): [images, width, K, m2] complex. */
size_t bdn_stage_layer_workspace_bytes(const BdnFnoShape* s);
int bdn_stage_layer_forward(const BdnFnoShape* s, const float* z_in, int32_t act_in,
                            const float* spec_w1, const float* spec_w2 /* NULL in 1-D */,
                            const float* conv_w, const float* conv_b, float* z_out, float* xs_saved,
                            void* ws, size_t ws_bytes, void* stream);
int bdn_stage_layer_backward(const BdnFnoShape* s, const float* gz_out, const float* z_in, int32_t act_in,
                             const float* xs_saved, const float* spec_w1, const float* spec_w2,
                             const float* conv_w, float* gz_in /* overwritten */,
                             float* g_spec_w1, float* g_spec_w2, float* g_conv_w, float* g_conv_b,
                             void* ws, size_t ws_bytes, void* stream);
/* Which kernels serve the spectral layers of this net (FNO2d.forward 2d_FPE/FNOModules.py:226-232):
 * 1 = fused tensor-core layer kernels (tc_p / tc_q_fwd / tc_q_bwd), 0 = FFMA layer kernels,
 * negative = BdnStatus.  Depends on ndim, images, width, hp, wp, m1, m2 and prec only. */
int bdn_fno_layer_path(const BdnFnoShape* s);
/* crop + fc1 + exact GELU + fc2   FNOModules.py:116-121 (1-D), :234-239 (2-D).
 * z: the last layer's output [images, width, hp, wp]; out: [images, out_h, out_w, c_out].
 * backward: gz [images, width, hp, wp] is overwritten (zero outside the crop); pooled_g as in
 * bdn_fno_backward. */
int bdn_stage_project_forward(const BdnFnoShape* s, const float* z, const float* fc1_w, const float* fc1_b,
                              const float* fc2_w, const float* fc2_b, float* out, void* stream);
int bdn_stage_project_backward(const BdnFnoShape* s, const float* z, const float* fc1_w, const float* fc1_b,
                               const float* fc2_w, const float* fc2_b, const float* g_out,
                               int32_t pooled_g, int32_t n_keep, float* gz,
                               float* g_fc1_w, float* g_fc1_b, float* g_fc2_w, float* g_fc2_b, void* stream);

/* ---------------------------------------------------------------------------
 * Bag mean + lift: fc0([grid, mean_l s_l]) with fc0 detached
 *   2d_FPE/NIOModules.py:564-575, 1d_FPE/NIOModules.py:139-149,
 *   1d_GPE/NIOModules.py:209-219.
 * s: [n_bags, n_keep, npix] per-snapshot scalars; grid: [npix, grid_dim];
 * w0: [width, grid_dim + 1]; b0: [width]; out: [n_bags, npix, width].
 * backward: gs[b, l, p] = (1/n_keep) * sum_j w0[j, grid_dim] * g[b, p, j]
 * is produced in POOLED form gpool [n_bags, npix] (feed it to
 * bdn_fno_backward with pooled_g = 1).
 * ------------------------------------------------------------------------- */
int bdn_bag_pool_lift_forward(const float* s, const float* grid, const float* w0, const float* b0,
                              float* out, int32_t n_bags, int32_t n_keep, int32_t npix,
                              int32_t grid_dim, int32_t width, void* stream);
int bdn_bag_pool_lift_backward(const float* g, const float* w0, float* gpool,
                               int32_t n_bags, int32_t npix, int32_t grid_dim, int32_t width,
                               void* stream);

/* ---------------------------------------------------------------------------
 * NIO tail: bag mean of the branch coefficients + DeepONet contraction + detached lift in one kernel
 *   DeepOnetNoBiasOrg.forward  1d_GPE/DeepONetModules.py:142-151   (branch(u) @ trunk(x)^T + b0) / sqrt(p)
 *   followed by the bag mean + fc0 lift of NIOFP_schrodinger.forward 1d_GPE/NIOModules.py:209-219
 *   (NIOFP2D.forward 2d_FPE/NIOModules.py:64-76, NIOFP.forward 1d_FPE/NIOModules.py:70-80).
 * By linearity the bag mean is taken on the coefficients, so the reference's [B, L, n_points] tensor is never formed.
 * w: [n_bags, n_keep, p] branch coefficients; basis: [npix, p] trunk output; b0: scalar (device);
 * grid: [npix, grid_dim]; fc0_w: [width, grid_dim + 1]; fc0_b: [width]; out: [n_bags, npix, width];
 * wbar_saved: [n_bags, p] (the pooled coefficients, needed by backward).
 * backward: g [n_bags, npix, width] -> g_w [n_bags, n_keep, p], g_basis [npix, p], g_b0 [1] (all OVERWRITTEN);
 * g_wbar_ws: [n_bags, p] scratch.  fc0 is detached in the reference (.data): it receives no gradient.
 * ------------------------------------------------------------------------- */
int bdn_nio_tail_forward(const float* w, const float* basis, const float* b0, const float* grid, const float* fc0_w,
                         const float* fc0_b, float* out, float* wbar_saved, int32_t n_bags, int32_t n_keep, int32_t p,
                         int32_t npix, int32_t grid_dim, int32_t width, void* stream);
int bdn_nio_tail_backward(const float* g, const float* basis, const float* wbar_saved, const float* fc0_w, float* g_w,
                          float* g_basis, float* g_b0, float* g_wbar_ws, int32_t n_bags, int32_t n_keep, int32_t p,
                          int32_t npix, int32_t grid_dim, int32_t width, void* stream);

/* ---------------------------------------------------------------------------
 * Bag attention + bag mean of the BlinDNO models, without [n_keep, dim] intermediates:
 *   TemporalSelfAttention.forward  2d_FPE/NIOModules.py:1063-1083  softmax(X X^T / sqrt(dim)) X + X, LayerNorm(dim)
 *   followed by .mean(dim=1)       2d_FPE/NIOModules.py:1153-1170  (1d_FPE/NIOModules.py, 1d_GPE/NIOModules.py alike)
 * x: [n_bags, n_keep, dim] tokens (flattened feature maps), n_keep <= 128; ln_w, ln_b: [dim] (LayerNorm affine);
 * out: [n_bags, dim]; saved: bdn_bag_attention_saved_floats() floats (row statistics, centered Gram matrix, attention
 * weights: what backward needs besides x).  backward: g [n_bags, dim] -> g_x [n_bags, n_keep, dim] (OVERWRITTEN),
 * g_ln_w_per_bag [n_bags, dim] (OVERWRITTEN; the caller sums over bags; d loss / d ln_b is the sum of g over bags);
 * ws: bdn_bag_attention_workspace_floats() floats.
 * ------------------------------------------------------------------------- */
size_t bdn_bag_attention_saved_floats(int32_t n_bags, int32_t n_keep);
size_t bdn_bag_attention_workspace_floats(int32_t n_bags, int32_t n_keep);
int bdn_bag_attention_mean_forward(const float* x, const float* ln_w, const float* ln_b, float* out, float* saved,
                                   int32_t n_bags, int32_t n_keep, int32_t dim, float eps, void* stream);
int bdn_bag_attention_mean_backward(const float* x, const float* g, const float* ln_w, const float* saved, float* g_x,
                                    float* g_ln_w_per_bag, float* ws, int32_t n_bags, int32_t n_keep, int32_t dim,
                                    void* stream);

/* ---------------------------------------------------------------------------
 * Training loss in one launch: criterion(model(inputs, grid), outputs) with criterion = torch.nn.MSELoss()
 *   2d_FPE/train_fno.py:116,146-147 (1d_FPE/train_fno.py, 1d_GPE/train_nio_GPE.py alike); the model's
 *   torch.cat of its head outputs (2d_FPE/NIOModules.py:577-581) is folded into the addressing.
 * outs: host array of n_heads (<= 4) device pointers, head k = [npix, c]; target: [npix, n_heads * c];
 * loss: one float (device); g_outs: null, or n_heads device pointers that receive d loss / d outs[k] =
 * (2 / (npix * n_heads * c)) * (outs[k] - target_k) in the same launch (what backward gives for grad_loss = 1: the
 * train step's case); scratch: 65 * 4 bytes of device memory, zero before the FIRST call (every call leaves
 * its counter word zero again, so CUDA-graph replays need no reset).  The sum is deterministic.
 * backward: g_outs[k] = (2 / (npix * n_heads * c)) * grad_loss[0] * (outs[k] - target_k)   (OVERWRITTEN)
 * ------------------------------------------------------------------------- */
int bdn_mse_heads_forward(const float* const* outs, int32_t n_heads, int32_t c, int64_t npix, const float* target,
                          float* loss, float* const* g_outs, void* scratch, void* stream);
int bdn_mse_heads_backward(const float* const* outs, int32_t n_heads, int32_t c, int64_t npix, const float* target,
                           const float* grad_loss, float* const* g_outs, void* stream);

/* ---------------------------------------------------------------------------
 * Optimiser step fused over a flat fp32 buffer (torch.optim.Adam semantics,
 * 2d_FPE/train_fno.py:117; eps added after the bias-corrected sqrt, no
 * weight decay / amsgrad).  step = 1-based step count.
 * ------------------------------------------------------------------------- */
int bdn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                  float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                  void* stream);

/* ---------------------------------------------------------------------------
 * misc
 * ------------------------------------------------------------------------- */
int bdn_abi_version(void);
const char* bdn_last_error(void);          /* thread-local, valid until the next failing call */
int bdn_pad_amount(int n);                 /* int(round(n / 4)) with Python's banker's rounding */
/* Build (and cache) the DFT tables of a shape now.  The first use of a shape allocates and copies synchronously,
 * which a CUDA-graph capture does not survive: compute entry points refuse (BDN_ERR_UNSUPPORTED text in
 * bdn_last_error) a first use under capture; call this, or the entry point once eagerly, beforehand. */
int bdn_prepare_plan(int32_t ndim, int32_t hp, int32_t wp, int32_t m1, int32_t m2);
int64_t bdn_kernel_launches(void);         /* kernels launched by this library so far (process-wide) */
int bdn_device_sm_count(void);
/* Per-kernel device timing for bench.py's roofline: between begin and end every kernel this
 * library launches is bracketed by a CUDA event pair on its stream (launches into a stream that is being captured
 * into a CUDA graph are not timed).  Every launch is also an NVTX range named after the kernel.  end synchronises those
 * events and writes a JSON object {"kernel/tag": {"launches": n, "ms": total}, ...} into buf
 * (truncated to cap); returns the number of bytes the full text needs. */
int bdn_profile_begin(void);
long bdn_profile_end(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BLINDNO_B200_H */
