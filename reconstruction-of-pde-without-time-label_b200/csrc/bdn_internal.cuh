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
  int Kp;       // K rounded up to 8 (table row pitch of t_hk)
  int hp8;      // hp rounded up to 8 (table row pitch of t_kh)
  int wp4;      // wp rounded up to 4 (row pitch of t_lw_cos / t_lw_sin)
  float2* t_wl;      // [wp][m2]   (cos theta, sin theta)      forward W transform
  float* t_lw_cos;   // [m2][wp4]  cos theta                   inverse W transform
  float* t_lw_sin;   // [m2][wp4]  sin theta
  float2* t_hk;      // [hp][Kp]   (cos phi, sin phi)          forward H transform
  float2* t_kh;      // [K][hp8]   (cos phi, sin phi)          inverse H transform
  float* col_fwd;    // [m2]  c_l / (hp*wp)   Hermitian doubling + irfft normalisation
  float* col_dc;     // [m2]  1-D: 0.5 at l = 0 (reference halves the DC bin), else 1
  // TF32 tensor-core path: the W tables as K-major GEMM operands (see tc_gemm.cu)
  float* tc_fwd_b;   // [n_pad = 2*m2 rounded up to 16][k_pad = wp rounded up to 8]
  float* tc_inv_b;   // [n_pad = wp rounded up to 16][k_pad = 2*m2 rounded up to 8]
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
void launch_wfwd(const Plan* pl, const float* x, float2* out, int rows, int act, cudaStream_t st);

// 2-D middle stage for one pass over `images` images:
//   in  [images, ca, hp, m2] complex  --H fwd, *pre--> spec_out [images, ca, K, m2] (if non-null)
//   --mix with W (fwd: W[a][b]; bwd: conj(W[b][a]))--> --H inv, *post--> out [images, cb, hp, m2]
void launch_core2d(const Plan* pl, const float2* in, float2* out, float2* spec_out,
                   const float2* w1, const float2* w2, int images, int ci_layer, int co_layer,
                   bool bwd, cudaStream_t st);
// 1-D middle stage (no H transform): in [images, ca, m2] -> out [images, cb, m2]
void launch_mix1d(const Plan* pl, const float2* in, float2* out, float2* spec_out, const float2* w,
                  int images, int ci_layer, int co_layer, bool bwd, cudaStream_t st);
// gw[i,o,k,l] += sum_b conj(xs[b,i,k,l]) * gys[b,o,k,l]   (split into w1 / w2 halves in 2-D)
void launch_gw_reduce(const Plan* pl, const float2* xs, const float2* gys, float2* gw1, float2* gw2,
                      int images, int ci, int co, cudaStream_t st);

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
void launch_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                 float eps, int step, float grad_scale, cudaStream_t st);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float gelu_exact(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace bdn
