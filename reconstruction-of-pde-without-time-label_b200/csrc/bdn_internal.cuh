// Internal declarations shared by the kernel translation units and the C ABI.
// Nothing here is part of the public boundary (include/blindno_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace bdn {

// ---------------------------------------------------------------------------
// DFT tables for one (hp, wp, m1, m2) shape, built once per device in fp64 and
// rounded to fp32.  theta = 2*pi*l*w/wp, phi = 2*pi*kk*h/hp with kk the kept
// row frequency (kk = k for k < m1, hp - 2*m1 + k otherwise).
// ---------------------------------------------------------------------------
struct Plan {
  int ndim, hp, wp, m1, m2;
  int K;        // 2*m1 (1-D: 1)
  int F;        // m1 + 1 row frequencies 0..m1: the kept rows come in conjugate pairs (k = f, k = -f) that share cos / sin
  int Fp;       // F rounded up to 4 (table row pitch of t_hf)
  int hp8;      // hp rounded up to 8 (table row pitch of t_fh)
  int wp4;      // wp rounded up to 4 (row pitch of t_lw_cos / t_lw_sin)
  float2* t_wl;      // [wp][m2]   (cos theta, sin theta)      forward W transform
  float2* t_wl_half; // [wl_nh4][wl_m2p] the same for w = 0..wp/2 only (rows / modes zero padded to multiples of 4): the
  int wl_nh, wl_nh4, wl_m2p;   // folded W-forward kernel pairs x[w] with x[wp - w], which share cos and differ in the sign of sin
  float* t_lw_cos;   // [m2][wp4]  cos theta                   inverse W transform
  float* t_lw_sin;   // [m2][wp4]  sin theta
  float2* t_hf;      // [hp][Fp]   (cos phi, sin phi), phi = 2*pi*f*h/hp     forward H transform (h-major)
  float2* t_fh;      // [F][hp8]   (cos phi, sin phi)                          inverse H transform (f-major)
  float* col_fwd;    // [m2]  c_l / (hp*wp)   Hermitian doubling + irfft normalisation
  float* col_dc;     // [m2]  1-D: 0.5 at l = 0 (reference halves the DC bin), else 1
  // TF32 tensor-core path: the W tables as K-major GEMM operands (see tc_gemm.cu)
  float* tc_fwd_b;   // [ceil(wp/32)][2*m2 rounded up to 16][32] fp32, 128-byte-swizzled smem image (or null)
  float* tc_inv_b;   // reserved for the inverse transform
  // fused tensor-core layer (tc_layer.cu): the four DFT operands F1..F4 as (hi | lo) no-swizzle K-major smem images
  float* tcl_f1; float* tcl_f1r; float* tcl_f2; float* tcl_f3; float* tcl_f4;   // shared-memory images
  float* tcl_t1; float* tcl_t3; float* tcl_t4;                   // tensor-memory tables of the operands that can be MMA operand A
};

const Plan* get_plan(int ndim, int hp, int wp, int m1, int m2);   // nullptr on failure

void count_launch(int n = 1);
int set_error(int code, const char* fmt, ...);

// Brackets one kernel launch: counts it, and when profiling is on (bdn_profile_begin) records a
// CUDA event pair on the launch stream so bench.py can report per-kernel device time.
struct LaunchScope {
  LaunchScope(const char* name, cudaStream_t st, int tag = -1);   // tag (e.g. channel width) is appended to the name
  ~LaunchScope();
  const char* name;
  int tag;
  cudaStream_t st;
  cudaEvent_t e0;
  bool on;
};

// ---------------------------------------------------------------------------
// kernel launchers (spectral.cu)
// ---------------------------------------------------------------------------
// rows x wp real -> rows x m2 complex; act != 0 applies exact GELU on load.
// prec = BDN_PREC_TF32 routes eligible calls (no activation on load, shape fits) to the tcgen05 kernel.
// returns true when the tcgen05 kernel ran (false: the FFMA kernel)
bool launch_wfwd(const Plan* pl, const float* x, float2* out, int rows, int act, cudaStream_t st, int prec = 0);
bool wfwd_uses_tensor_cores(const Plan* pl, const float* x, int rows, int prec);

// 2-D middle stage for one pass over `images` images:
//   in  [images, ca, hp, m2] complex  --H fwd, *pre--> spec_out [images, ca, K, m2] (if non-null)
//   --mix with W (fwd: W[a][b]; bwd: conj(W[b][a]))--> --H inv, *post--> out [images, cb, hp, m2]
// wt (optional): mode-major copy of (w1, w2) made by launch_spec_weights_mode_major.
void launch_core2d(const Plan* pl, const float2* in, float2* out, float2* spec_out,
                   const float2* w1, const float2* w2, int images, int ci_layer, int co_layer,
                   bool bwd, cudaStream_t st, const float2* wt = nullptr);
void launch_spec_weights_mode_major(const Plan* pl, const float* const* w1, const float* const* w2, int n_layers,
                                    int ci, int co, float2* wt, cudaStream_t st);
// 1-D middle stage (no H transform): in [images, ca, m2] -> out [images, cb, m2]
void launch_mix1d(const Plan* pl, const float2* in, float2* out, float2* spec_out, const float2* w,
                  int images, int ci_layer, int co_layer, bool bwd, cudaStream_t st);
// gw[i,o,k,l] += sum_b conj(xs[b,i,k,l]) * gys[b,o,k,l]   (split into w1 / w2 halves in 2-D)
void launch_gw_reduce(const Plan* pl, const float2* xs, const float2* gys, float2* gw1, float2* gw2,
                      int images, int ci, int co, cudaStream_t st);

// Fused 1-D layer (fused1d.cu): W-forward DFT + mix + inverse DFT + 1x1 conv epilogue of one layer in one launch.
//   fwd: out = z_out, spec = xs_saved (may be null);  bwd: out = gz_in, spec = gys (input of launch_gw_reduce),
//   g_pw_w / g_pw_b accumulated.  Returns false (nothing launched) when unsupported: use the three-kernel path.
bool launch_layer1d(const Plan* pl, bool bwd, const float* z_in, const float* g_out, float* out, float2* spec,
                    const float2* w, const float* pw_w, const float* pw_b, float* g_pw_w, float* g_pw_b, int images, int c,
                    int act_in, cudaStream_t st);

// ---------------------------------------------------------------------------
// tensor-core path (tc_gemm.cu)
// ---------------------------------------------------------------------------
}  // namespace bdn
#include <vector>
namespace bdn {
int tc_n_pad(int m2);
int tc_kch(int wp);
void tc_build_b_image(int wp, int m2, std::vector<float>& img);     // host image of the swizzled DFT operand
bool tc_wfwd_supported(const Plan* pl, const float* x, bool split);
// act: exact GELU on load; split: 3xTF32 (hi/lo operands, fp32-level accuracy) instead of plain TF32
bool launch_wfwd_tc(const Plan* pl, const float* x, float2* out, int rows, int act, bool split, cudaStream_t st);

// ---------------------------------------------------------------------------
// fused tensor-core layer (tc_layer.cu): kernel P (planes -> kept spectrum) and kernel Q (kept spectrum -> planes
// with the layer epilogue).  prec: 1 = TF32, 2 = 3xTF32.  The launchers return false (nothing launched) when the
// shape does not fit; tcl_supported says so beforehand.
// ---------------------------------------------------------------------------
void tcl_build_tables(Plan* pl);
bool tcl_supported(const Plan* pl, int images, int C);
bool launch_tcl_p(const Plan* pl, const float* x, float* a_out, float2* spec_out, const float* pre, int images, int C,
                  int act, int prec, cudaStream_t st);
bool launch_tcl_q(const Plan* pl, bool bwd, const float2* xin, const float2* w1, const float2* w2, const float* a_in,
                  const float* zin, float* out, const float* pw_w, const float* pw_b, float* g_pw_w, float* g_pw_b,
                  const float* post, int images, int C, int act_in, int prec, cudaStream_t st);

enum WinvMode { WINV_PLAIN = 0, WINV_LAYER_FWD = 1, WINV_LAYER_BWD = 2 };
struct WinvArgs {
  const float2* z;     // [images, c, hp, m2] complex (already column-scaled)
  float* y;            // PLAIN / FWD: output [images, c, hp, wp]; BWD: gz_in
  const float* a;      // FWD: layer input z_in (activation applied on load if act_in)
                       // BWD: gz_out
  const float* zin;    // BWD: layer input pre-activation z_in
  const float* pw_w;   // [c, c] 1x1 conv weight
  const float* pw_b;   // [c]    (FWD only)
  float* g_pw_w;       // BWD: += sum gz_out[o] * act(z_in)[i]
  float* g_pw_b;       // BWD: += sum gz_out[o]
  int images, c, act_in;
};
void launch_winv(const Plan* pl, int mode, const WinvArgs& a, cudaStream_t st);

// ---------------------------------------------------------------------------
// kernel launchers (pointwise.cu)
// ---------------------------------------------------------------------------
struct LiftArgs {
  const float* x_cl; const float* bags; const int32_t* idx; const float* grid;
  int n_bags, bag_len, n_keep, grid_dim;
  const float* w0; const float* b0;
  int images, c_in, width, h, w, hp, wp;
};
void launch_lift(const LiftArgs& a, float* z0, cudaStream_t st);
void launch_lift_bwd(const LiftArgs& a, const float* gz0, float* g_w0, float* g_b0, float* gx_cl,
                     cudaStream_t st);

struct ProjArgs {
  const float* z;       // [images, width, hp, wp]
  const float* w1; const float* b1; const float* w2; const float* b2;
  int images, width, hidden, c_out, hp, wp, out_h, out_w;
};
void launch_project(const ProjArgs& a, float* out, cudaStream_t st);
// gz is fully overwritten (zero outside the cropped window)
void launch_project_bwd(const ProjArgs& a, const float* g_out, int pooled_g, int n_keep, float* gz,
                        float* g_w1, float* g_b1, float* g_w2, float* g_b2, cudaStream_t st);

void launch_pool_lift(const float* s, const float* grid, const float* w0, const float* b0, float* out,
                      int n_bags, int n_keep, int npix, int grid_dim, int width, cudaStream_t st);
void launch_pool_lift_bwd(const float* g, const float* w0, float* gpool, int n_bags, int npix,
                          int grid_dim, int width, cudaStream_t st);
// NIO tail (K6): bag mean of the branch coefficients + DeepONet contraction + detached lift, and its backward
void launch_nio_tail(const float* w, const float* basis, const float* b0, const float* grid, const float* w0, const float* fb,
                     float* out, float* wbar, int n_bags, int L, int p, int npix, int gd, int width, cudaStream_t st);
void launch_nio_tail_bwd(const float* g, const float* basis, const float* wbar, const float* w0, float* g_wbar, float* g_basis,
                         float* g_b0, float* g_w, int n_bags, int L, int p, int npix, int gd, int width, cudaStream_t st);
// bag attention + bag mean (bagattn.cu): TemporalSelfAttention + .mean(dim=1) of the BlinDNO models
size_t bagattn_saved_floats(int n_bags, int L);
size_t bagattn_backward_ws_floats(int n_bags, int L);
void launch_bagattn_forward(const float* x, const float* gamma, const float* beta, float* out, float* saved, int n_bags, int L,
                            int D, float eps, cudaStream_t st);
void launch_bagattn_backward(const float* x, const float* g, const float* gamma, const float* saved, float* dx, float* dgamma,
                             float* ws, int n_bags, int L, int D, cudaStream_t st);
// MSE over the concatenated head outputs without the concatenation (forward: deterministic two-level sum; backward)
constexpr int MSE_MAX_HEADS = 4;
struct MseHeadsArgs {
  const float* out[MSE_MAX_HEADS]; float* g[MSE_MAX_HEADS]; const float* target;
  int n_heads, c; long npix;
};
int mse_heads_blocks(long npix);
void launch_mse_heads(const MseHeadsArgs& a, float* loss, float* partial, unsigned int* counter, cudaStream_t st);
void launch_mse_heads_bwd(const MseHeadsArgs& a, const float* grad_loss, cudaStream_t st);
void launch_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                 float eps, int step, float grad_scale, cudaStream_t st);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
// Exact (erf) GELU, F.gelu's default (FNOModules.py:113-114, :231-232), evaluated with ONE ex2 and
// ONE rcp.  Phi(x) = 1 - q (x >= 0) or q (x < 0) with q = phi(|x|) * (b1 t + ... + b5 t^5), t = 1 / (1 + p |x|):
// Abramowitz-Stegun 26.2.17 (the normal-CDF form of 7.1.26; |error| <= 7.5e-8 on the CDF, measured max |gelu error|
// 4.2e-7 over [-12, 12] against 1.2e-6 for torch's own fp32 gelu; tests/test_cabi_cpu.py restates the formula in
// NumPy fp32 and checks that bound).  The density phi(x) = exp(-x^2/2)/sqrt(2 pi) comes straight out of the ex2
// (its constant is folded into the exponent: log2(1/sqrt(2 pi)) = -1.3257...), so the derivative
// gelu'(x) = Phi + x phi costs one more FMA: that is what makes the recompute-in-backward projection affordable.
// 12 FMA-pipe instructions + 2 MUFU for (Phi, phi).
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.2316418882663604f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(x * x, -0.72134752044448170368f, -1.3257480647361592f)));
  float p = 1.3302745f;
  p = fmaf(p, t, -1.8212559f);
  p = fmaf(p, t, 1.7814779f);
  p = fmaf(p, t, -0.35656378f);
  p = fmaf(p, t, 0.31938154f);
  const float q = p * t * e;
  cdf = x >= 0.f ? 1.0f - q : q;
  pdf = e;
}
// gelu(x) = x Phi(x) without forming Phi: x (1 - q) for x >= 0 and x q for x < 0 are both max(x, 0) - |x| q
// (one FMNMX + one FFMA instead of a compare, a predicated subtraction and a multiply: 13 instead of 14 instructions)
__device__ __forceinline__ float gelu_fast(float x) {
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.2316418882663604f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(x * x, -0.72134752044448170368f, -1.3257480647361592f)));
  float p = 1.3302745f;
  p = fmaf(p, t, -1.8212559f);
  p = fmaf(p, t, 1.7814779f);
  p = fmaf(p, t, -0.35656378f);
  p = fmaf(p, t, 0.31938154f);
  const float q = p * t * e;
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_exact(float x) { return gelu_fast(x); }
__device__ __forceinline__ float gelu_grad(float x) {
  float cdf, pdf;
  gelu_cdf_pdf(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// ---------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D, SASS UBLKCP) helpers: global -> shared staging without a
// register round trip, completion signalled on an mbarrier by transaction bytes.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  A step is a chain of ~90 short dependent kernels; with the launch
// attribute below the next kernel's blocks are scheduled while the previous kernel drains, and wait in
// pdl_wait() until that kernel has completed and flushed its memory.  Every kernel of this library calls
// pdl_launch_dependents() at its very top and pdl_wait() before it touches anything a previous kernel
// may have written (loads of the constant DFT tables are hoisted above the wait).  Both are no-ops for
// launches without the attribute.  Opt-in with BDN_PDL=1 in the environment (off by default: it measured
// neutral once the step is replayed from a CUDA graph, 1.402 ms vs 1.381 ms per step).
// ---------------------------------------------------------------------------
// Build with -DBDN_PDL_LATE=1 (BDN_NVCC_EXTRA in the environment of build.py) to move the trigger from the top of every
// kernel to a point late in its work (pdl_trigger_late): with the trigger at the top a whole chain of future kernels
// pre-launches and their waiting blocks hold shared memory (one head alone: 252 -> 317 us).
#ifdef BDN_PDL_LATE
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_trigger_late() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger_late() {}
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();      // BDN_PDL: 0 off, 1 every launch, 2 every launch except those of few-image nets (pdl_few_images)
extern thread_local int pdl_few_images;      // set by the FNO entry points while they launch a few-image net (the heads)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace bdn
