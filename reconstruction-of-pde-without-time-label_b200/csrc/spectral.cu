// Pruned-DFT stages of the FNO spectral convolution, fp32 CUDA-core path (the 1e-5 parity
// mode).  One spectral layer is W-forward -> [H-forward -> channel mix -> H-inverse] -> W-inverse
// with the 1x1 conv / bias / GELU-gradient epilogue fused into the W-inverse store.  Only the kept
// modes are ever computed; the zero-filled spectrum of the reference is never materialised.
//
// Replaces (reference file:line): SpectralConv2d.forward 2d_FPE/FNOModules.py:156-178,
// compl_mul2d :141-154, SpectralConv1d.forward 1d_FPE/FNOModules.py:47-59 and the layer body
// FNO2d.forward :226-232 / FNO1d.forward :108-114.
#include "bdn_internal.cuh"

#include <cstdlib>

#include <cstdio>
#include <mutex>
#include <set>
#include <tuple>

namespace bdn {

// ===========================================================================
// W-forward: out[r, l] = sum_w act(x[r, w]) * (cos - i sin)(theta_lw),   r over (image, chan, h)
// Block = 32*RT rows; lane <-> RT consecutive rows, warp <-> 4 consecutive modes.
// ===========================================================================
template <int RT>
__global__ void __launch_bounds__(512) wfwd_generic_kernel(const float* __restrict__ x, float2* __restrict__ out,
                                                   const float2* __restrict__ t_wl, int rows, int wp, int m2,
                                                   int act, int wc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LT = 4;
  constexpr int BR = 32 * RT;
  constexpr int PITCH = BR + 4;
  extern __shared__ __align__(16) float smem[];
  const int nlg = blockDim.x >> 5;
  const int m2p = nlg * LT;
  float* xs = smem;                                             // [wc][PITCH]
  float2* ts = reinterpret_cast<float2*>(smem + (size_t)wc * PITCH);  // [wc][m2p]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * BR;

  float ar[RT][LT], ai[RT][LT];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int j = 0; j < LT; ++j) ar[r][j] = ai[r][j] = 0.f;

  for (int w0 = 0; w0 < wp; w0 += wc) {
    const int wcur = min(wc, wp - w0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < BR * wcur; idx += blockDim.x) {
      const int r = idx / wcur, w = idx - r * wcur;
      const int row = row0 + r;
      float v = row < rows ? __ldg(x + (size_t)row * wp + w0 + w) : 0.f;
      if (act) v = gelu_exact(v);
      xs[w * PITCH + r] = v;
    }
    for (int idx = threadIdx.x; idx < wcur * m2p; idx += blockDim.x) {
      const int w = idx / m2p, l = idx - w * m2p;
      ts[idx] = l < m2 ? __ldg(t_wl + (size_t)(w0 + w) * m2 + l) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int w = 0; w < wcur; ++w) {
      float xv[RT];
      if constexpr (RT == 4) {
        const float4 v = *reinterpret_cast<const float4*>(xs + w * PITCH + lane * 4);
        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      } else if constexpr (RT == 2) {
        const float2 v = *reinterpret_cast<const float2*>(xs + w * PITCH + lane * 2);
        xv[0] = v.x; xv[1] = v.y;
      } else {
        xv[0] = xs[w * PITCH + lane];
      }
      const float4 t01 = *reinterpret_cast<const float4*>(ts + w * m2p + warp * LT);
      const float4 t23 = *reinterpret_cast<const float4*>(ts + w * m2p + warp * LT + 2);
      const float tc[4] = {t01.x, t01.z, t23.x, t23.z};
      const float tsn[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int j = 0; j < LT; ++j) {
          ar[r][j] = fmaf(xv[r], tc[j], ar[r][j]);
          ai[r][j] = fmaf(-xv[r], tsn[j], ai[r][j]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = row0 + lane * RT + r;
    if (row >= rows) continue;
#pragma unroll
    for (int j = 0; j < LT; ++j) {
      const int l = warp * LT + j;
      if (l < m2) out[(size_t)row * m2 + l] = make_float2(ar[r][j], ai[r][j]);
    }
  }
}

static int pick_wchunk(int wp, int per_w_bytes, int budget) {
  int n = 1;
  while (ceil_div(wp, n) * per_w_bytes > budget) ++n;
  return ceil_div(wp, n);
}

static void launch_wfwd_generic(const Plan* pl, const float* x, float2* out, int rows, int act, cudaStream_t st) {
  const int m2 = pl->m2, wp = pl->wp;
  const int nlg = ceil_div(m2, 4);
  const int sms = 148;
  int rt = 4;
  if (rows < 128 * 2 * sms) rt = 2;
  if (rows < 64 * 2 * sms) rt = 1;
  const int br = 32 * rt;
  const int per_w = (br + 4) * 4 + nlg * 4 * 8;
  const int wc = pick_wchunk(wp, per_w, 96 * 1024);
  const size_t smem = (size_t)wc * per_w;
  dim3 grid(ceil_div(rows, br)), block(32 * nlg);
#define BDN_WFWD(RT)                                                                                      \
  {                                                                                                       \
    cudaFuncSetAttribute(wfwd_generic_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
    launch_k(wfwd_generic_kernel<RT>, dim3(grid), dim3(block), smem, st, x, out, pl->t_wl, rows, wp, m2, act, wc);          \
  }
  if (rt == 4) BDN_WFWD(4) else if (rt == 2) BDN_WFWD(2) else BDN_WFWD(1)
#undef BDN_WFWD
}

// ---------------------------------------------------------------------------
// W-forward, pipelined: persistent blocks, the x tile (BR consecutive rows = one contiguous span of
// HBM) is staged by per-row bulk async copies into a double-buffered shared tile whose completion an
// mbarrier tracks; the DFT table stays in shared memory for the block's lifetime.  Warp = (row group,
// group of 4 modes), lane = RT rows.  x is read with 128-bit shared loads along w; the row pitch is
// chosen so that the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups.
// ---------------------------------------------------------------------------
struct WfwdParams {
  const float* x; float2* out; const float2* t_wl;
  int rows, wp, m2, act, pitch, nrg, nmg, m2p, ntiles;
};

template <int RT>
__global__ void __launch_bounds__(256) wfwd_pipe_kernel(const WfwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int BR = 32 * RT * p.nrg;
  const int tile_floats = BR * p.pitch;
  float* stage0 = smem;
  float* stage1 = smem + tile_floats;
  float2* ts = reinterpret_cast<float2*>(smem + 2 * tile_floats);     // [wp][m2p]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ts + (size_t)p.wp * p.m2p);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int wp = p.wp, m2 = p.m2, m2p = p.m2p;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init_fence();
  }
  for (int i = tid; i < wp * m2p; i += nt) {      // constant plan data: staged before the dependency wait
    const int w = i / m2p, l = i - w * m2p;
    ts[i] = l < m2 ? __ldg(p.t_wl + (size_t)w * m2 + l) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  pdl_wait();

  const uint32_t row_bytes = (uint32_t)wp * 4u;
  auto issue = [&](int tile, int stage) {   // called by all lanes of warp 0
    const int row0 = tile * BR;
    const int nrows = min(BR, p.rows - row0);
    float* dst = stage ? stage1 : stage0;
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bars[stage], (uint32_t)nrows * row_bytes);
      // the tile is one contiguous span of HBM: a single bulk copy when no row padding is needed
      if (p.pitch == wp) bulk_g2s(dst, p.x + (size_t)row0 * wp, (uint32_t)nrows * row_bytes, &bars[stage]);
    }
    __syncwarp();
    if (p.pitch != wp)
      for (int r = lane; r < nrows; r += 32)
        bulk_g2s(dst + r * p.pitch, p.x + (size_t)(row0 + r) * wp, row_bytes, &bars[stage]);
  };
  if (warp == 0 && (int)blockIdx.x < p.ntiles) issue(blockIdx.x, 0);

  const int rg = warp % p.nrg, mg0 = warp / p.nrg, mgstep = (nt >> 5) / p.nrg;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const int next = tile + gridDim.x;
    if (warp == 0 && next < p.ntiles) issue(next, stage ^ 1);
    if (next >= p.ntiles) pdl_trigger_late();
    mbar_wait(&bars[stage], (it >> 1) & 1);
    float* xs = stage ? stage1 : stage0;
    const int row0 = tile * BR;
    if (p.act) {
      const int nrows = min(BR, p.rows - row0);
      const int w4n = wp >> 2;
      for (int i = tid; i < nrows * w4n; i += nt) {
        const int r = i / w4n, q = i - r * w4n;
        float4* ptr = reinterpret_cast<float4*>(xs + r * p.pitch) + q;
        float4 v = *ptr;
        v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
        *ptr = v;
      }
      __syncthreads();
    }
    for (int mg = mg0; mg < p.nmg; mg += mgstep) {
      float ar[RT][4], ai[RT][4];
#pragma unroll
      for (int j = 0; j < RT; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) ar[j][m] = ai[j][m] = 0.f;
      const float* xrow = xs + (rg * 32 * RT + lane) * p.pitch;
      const float4* tb = reinterpret_cast<const float4*>(ts + mg * 4);
      const int tpitch4 = m2p >> 1;   // float4 per table row
#pragma unroll 2
      for (int w4 = 0; w4 < (wp >> 2); ++w4) {
        float xv[RT][4];
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(xrow + j * 32 * p.pitch + 4 * w4);
          xv[j][0] = v.x; xv[j][1] = v.y; xv[j][2] = v.z; xv[j][3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t01 = tb[(4 * w4 + i) * tpitch4];
          const float4 t23 = tb[(4 * w4 + i) * tpitch4 + 1];
          const float tc[4] = {t01.x, t01.z, t23.x, t23.z};
          const float tsn[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
          for (int j = 0; j < RT; ++j)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              ar[j][m] = fmaf(xv[j][i], tc[m], ar[j][m]);
              ai[j][m] = fmaf(-xv[j][i], tsn[m], ai[j][m]);
            }
        }
      }
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int row = row0 + rg * 32 * RT + j * 32 + lane;
        if (row >= p.rows) continue;
        float2* o = p.out + (size_t)row * m2 + mg * 4;
        if ((m2 & 3) == 0) {
          reinterpret_cast<float4*>(o)[0] = make_float4(ar[j][0], ai[j][0], ar[j][1], ai[j][1]);
          reinterpret_cast<float4*>(o)[1] = make_float4(ar[j][2], ai[j][2], ar[j][3], ai[j][3]);
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (mg * 4 + m < m2) o[m] = make_float2(ar[j][m], ai[j][m]);
        }
      }
    }
    fence_proxy_async();   // generic-proxy accesses to this stage are ordered before the async refill
    __syncthreads();       // everyone is done with this stage before it is refilled
  }
}

// ---------------------------------------------------------------------------
// W-forward, folded: cos(theta_{l, wp-w}) = cos(theta_{l, w}) and sin(theta_{l, wp-w}) = -sin(theta_{l, w}), so with
//   e[w] = x[w] + x[wp-w],  o[w] = x[w] - x[wp-w]     (w = 1 .. ceil(wp/2)-1;  e = x, o = 0 at w = 0 and w = wp/2)
// the pruned DFT is  Re X[l] = sum_{w <= wp/2} e[w] cos,  Im X[l] = -sum o[w] sin: half the multiply-adds and half the
// table.  Same thread mapping as wfwd_pipe_kernel (warp = (row group, 4 modes), lane = RT rows, 128-bit row-strided
// shared loads).  The raw tile lands by one bulk copy; a fold pass (which also applies the GELU of layers > 0) writes
// the (e | o) tile; the raw buffer is then free, so the next tile's copy runs under this tile's multiply-adds.
// ---------------------------------------------------------------------------
struct WfoldParams {
  const float* x; float2* out; const float2* t_half;
  int rows, wp, m2, act, nh, nh4, fpitch, nrg, nmg, m2p, ntiles;
};

template <int RT>
__global__ void __launch_bounds__(256) wfwd_fold_kernel(const WfoldParams p) {
  extern __shared__ __align__(16) float smem[];
  const int BR = 32 * RT * p.nrg;
  const int wp = p.wp, m2 = p.m2, m2p = p.m2p, nh = p.nh, nh4 = p.nh4, fpitch = p.fpitch;
  float* raw = smem;                                          // [BR][wp]
  float* F = raw + BR * wp;                                   // [BR][fpitch]: e at 0 .. nh4-1, o at nh4 .. 2*nh4-1
  float2* ts = reinterpret_cast<float2*>(F + BR * fpitch);    // [nh4][m2p]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ts + (size_t)nh4 * m2p);   // [0] raw tile, [1] table
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init_fence();
    const uint32_t tb = (uint32_t)(nh4 * m2p) * 8u;           // constant plan data: staged before the dependency wait
    mbar_expect_tx(&bars[1], tb);
    bulk_g2s(ts, p.t_half, tb, &bars[1]);
  }
  __syncthreads();
  pdl_wait();
  auto issue = [&](int tile) {   // thread 0
    const int row0 = tile * BR;
    const uint32_t bytes = (uint32_t)min(BR, p.rows - row0) * (uint32_t)wp * 4u;
    fence_proxy_async();
    mbar_expect_tx(&bars[0], bytes);
    bulk_g2s(raw, p.x + (size_t)row0 * wp, bytes, &bars[0]);
  };
  if (tid == 0 && (int)blockIdx.x < p.ntiles) issue(blockIdx.x);
  mbar_wait(&bars[1], 0);

  const int rg = warp % p.nrg, mg0 = warp / p.nrg, mgstep = (nt >> 5) / p.nrg;
  const float inv_nh4 = 1.0f / (float)nh4;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    mbar_wait(&bars[0], it & 1);
    const int row0 = tile * BR, nrows = min(BR, p.rows - row0);
    for (int idx = tid; idx < BR * nh4; idx += nt) {
      const int r = __float2int_rz(((float)idx + 0.5f) * inv_nh4), j = idx - r * nh4;     // exact for these ranges
      float e = 0.f, o = 0.f;
      if (r < nrows && j < nh) {
        float a = raw[r * wp + j];
        if (p.act) a = gelu_fast(a);
        e = a;
        const int jm = wp - j;
        if (j > 0 && jm > j) {
          float b = raw[r * wp + jm];
          if (p.act) b = gelu_fast(b);
          e = a + b;
          o = a - b;
        }
      }
      F[r * fpitch + j] = e;
      F[r * fpitch + nh4 + j] = o;
    }
    fence_proxy_async();   // generic-proxy reads of the raw tile are ordered before the async refill
    __syncthreads();       // the raw tile is free, the folded tile complete
    const int next = tile + gridDim.x;
    if (tid == 0 && next < p.ntiles) issue(next);
    if (next >= p.ntiles) pdl_trigger_late();      // the block's last tile: only its multiply-adds and stores remain

    for (int mg = mg0; mg < p.nmg; mg += mgstep) {
      float ar[RT][4], ai[RT][4];
#pragma unroll
      for (int j = 0; j < RT; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) ar[j][m] = ai[j][m] = 0.f;
      const float* xrow = F + (rg * 32 * RT + lane) * fpitch;
      const float4* tb = reinterpret_cast<const float4*>(ts + mg * 4);
      const int tpitch4 = m2p >> 1;   // float4 per table row
#pragma unroll 2
      for (int w4 = 0; w4 < (nh4 >> 2); ++w4) {
        float ev[RT][4], ov[RT][4];
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(xrow + j * 32 * fpitch + 4 * w4);
          const float4 u = *reinterpret_cast<const float4*>(xrow + j * 32 * fpitch + nh4 + 4 * w4);
          ev[j][0] = v.x; ev[j][1] = v.y; ev[j][2] = v.z; ev[j][3] = v.w;
          ov[j][0] = u.x; ov[j][1] = u.y; ov[j][2] = u.z; ov[j][3] = u.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t01 = tb[(4 * w4 + i) * tpitch4];
          const float4 t23 = tb[(4 * w4 + i) * tpitch4 + 1];
          const float tc[4] = {t01.x, t01.z, t23.x, t23.z};
          const float tsn[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
          for (int j = 0; j < RT; ++j)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              ar[j][m] = fmaf(ev[j][i], tc[m], ar[j][m]);
              ai[j][m] = fmaf(-ov[j][i], tsn[m], ai[j][m]);
            }
        }
      }
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int row = row0 + rg * 32 * RT + j * 32 + lane;
        if (row >= p.rows) continue;
        float2* o = p.out + (size_t)row * m2 + mg * 4;
        if ((m2 & 3) == 0) {
          reinterpret_cast<float4*>(o)[0] = make_float4(ar[j][0], ai[j][0], ar[j][1], ai[j][1]);
          reinterpret_cast<float4*>(o)[1] = make_float4(ar[j][2], ai[j][2], ar[j][3], ai[j][3]);
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (mg * 4 + m < m2) o[m] = make_float2(ar[j][m], ai[j][m]);
        }
      }
    }
    __syncthreads();       // the folded tile is free before the next fold pass
  }
}

bool wfwd_uses_tensor_cores(const Plan* pl, const float* x, int rows, int prec) {
  return (prec == 1 || prec == 2) && rows >= 128 && tc_wfwd_supported(pl, x, prec == 2);
}

bool launch_wfwd(const Plan* pl, const float* x, float2* out, int rows, int act, cudaStream_t st, int prec) {
  // prec 1 = TF32 tensor cores (2e-3 mode); prec 2 = 3xTF32 tensor cores (operands split into TF32 high and
  // low parts, three MMAs per K step: fp32-level accuracy).  A shape that does not fit the tensor-core kernel's
  // shared-memory budget runs the FFMA kernel below: the return value says which (true = tcgen05), the profile
  // tag names the kernel, and the first such fallback of a shape is reported on stderr.
  // fp32 mode, many rows, few modes, GELU on load (layers > 0 of the per-snapshot net): the tcgen05 kernel with 3xTF32
  // operands (fp32-level accuracy) is the faster one -- r2d: 23.8 vs 27.0 us per launch; without the GELU it is not
  // (21.8 vs 21.0 us) and the FFMA kernel stays.  BDN_WFWD_TC_AUTO=0 switches this off, 2 extends it to the plain case.
  static const int tc_auto = [] { const char* e = getenv("BDN_WFWD_TC_AUTO"); return e ? atoi(e) : 1; }();
  if (tc_auto && (act || tc_auto == 2) && prec == 0 && rows >= 50000 && pl->m2 < 20 && tc_wfwd_supported(pl, x, true) &&
      launch_wfwd_tc(pl, x, out, rows, act, true, st))
    return true;
  if (wfwd_uses_tensor_cores(pl, x, rows, prec) && launch_wfwd_tc(pl, x, out, rows, act, prec == 2, st)) return true;
  if (prec == 1 || prec == 2) {
    static std::mutex mu;
    static std::set<std::tuple<int, int, int>> seen;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.insert(std::make_tuple(pl->wp, pl->m2, prec)).second)
      fprintf(stderr, "blindno_b200: W-forward DFT (wp=%d, modes=%d, rows=%d) does not fit the tcgen05 kernel in %s mode; "
                      "running the fp32 FFMA kernel\n", pl->wp, pl->m2, rows, prec == 2 ? "3xTF32" : "TF32");
  }
  LaunchScope scope(act ? "wfwd_gelu" : "wfwd", st, pl->m2);
  const int m2 = pl->m2, wp = pl->wp;
  if ((wp & 3) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) {   // bulk copies need 16-byte rows
    launch_wfwd_generic(pl, x, out, rows, act, st);
    return false;
  }
  static const int fold_knob = [] { const char* e = getenv("BDN_WFWD_FOLD"); return e ? atoi(e) : 1; }();   // (tuning knob)
  // folding pays where the multiply-adds dominate the fold pass: measured (r2r) 13 % faster at 32 modes (the heads),
  // 10 % slower at 12 modes (the per-snapshot net: 24 FMAs per element become 12, the fold pass adds ~8 instructions)
  if (fold_knob && pl->t_wl_half != nullptr && (m2 >= 20 || fold_knob == 2)) {
    WfoldParams f;
    f.x = x; f.out = out; f.t_half = pl->t_wl_half; f.rows = rows; f.wp = wp; f.m2 = m2; f.act = act;
    f.nh = pl->wl_nh; f.nh4 = pl->wl_nh4; f.m2p = pl->wl_m2p;
    f.nmg = f.m2p >> 2;
    f.fpitch = 2 * f.nh4 + ((((2 * f.nh4) >> 2) & 1) ? 0 : 4);      // (pitch / 4) odd: conflict-free 128-bit row-strided loads
    f.nrg = f.nmg <= 2 ? 4 : (f.nmg <= 4 ? 2 : 1);
    const int nwarps = f.nrg * (f.nmg < 8 / f.nrg ? f.nmg : 8 / f.nrg);
    auto smem_of = [&](int r) {
      return (size_t)32 * r * f.nrg * (wp + f.fpitch) * 4 + (size_t)f.nh4 * f.m2p * 8 + 16;
    };
    int rt = 4;
    while (rt > 1 && (ceil_div(rows, 32 * rt * f.nrg) < 2 * 148 || smem_of(rt) > 100 * 1024)) rt >>= 1;
    if (smem_of(rt) <= 200 * 1024) {
      f.ntiles = ceil_div(rows, 32 * rt * f.nrg);
      const size_t smem = smem_of(rt);
      const int per_sm = (int)((220 * 1024) / (smem + 1024));
      const int cap = 148 * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
      const int grid = f.ntiles < cap ? f.ntiles : cap;
#define BDN_WFOLD(RT)                                                                                    \
  {                                                                                                      \
    cudaFuncSetAttribute(wfwd_fold_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    launch_k(wfwd_fold_kernel<RT>, dim3(grid), dim3(32 * nwarps), smem, st, f);                          \
  }
      if (rt == 4) BDN_WFOLD(4) else if (rt == 2) BDN_WFOLD(2) else BDN_WFOLD(1)
#undef BDN_WFOLD
      return false;
    }
  }
  WfwdParams p;
  p.x = x; p.out = out; p.t_wl = pl->t_wl; p.rows = rows; p.wp = wp; p.m2 = m2; p.act = act;
  p.nmg = ceil_div(m2, 4);
  p.m2p = p.nmg * 4;
  p.pitch = ((wp >> 2) & 1) ? wp : wp + 4;       // (pitch / 4) odd: conflict-free 128-bit row-strided loads
  p.nrg = p.nmg <= 2 ? 4 : (p.nmg <= 4 ? 2 : 1);
  const int nwarps = p.nrg * (p.nmg < 8 / p.nrg ? p.nmg : 8 / p.nrg);
  int rt = 4;
  auto smem_of = [&](int r) {
    return (size_t)2 * 32 * r * p.nrg * p.pitch * 4 + (size_t)wp * p.m2p * 8 + 16;
  };
  while (rt > 1 && (ceil_div(rows, 32 * rt * p.nrg) < 2 * 148 || smem_of(rt) > 100 * 1024)) rt >>= 1;
  if (smem_of(rt) > 200 * 1024) {
    launch_wfwd_generic(pl, x, out, rows, act, st);
    return false;
  }
  const int BR = 32 * rt * p.nrg;
  p.ntiles = ceil_div(rows, BR);
  const size_t smem = smem_of(rt);
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int cap = 148 * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
  const int grid = p.ntiles < cap ? p.ntiles : cap;
#define BDN_WFWD(RT)                                                                                   \
  {                                                                                                    \
    cudaFuncSetAttribute(wfwd_pipe_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    launch_k(wfwd_pipe_kernel<RT>, dim3(grid), dim3(32 * nwarps), smem, st, p);                                          \
  }
  if (rt == 4) BDN_WFWD(4) else if (rt == 2) BDN_WFWD(2) else BDN_WFWD(1)
#undef BDN_WFWD
  return false;
}

// ===========================================================================
// 2-D middle stage: H-forward -> mix -> H-inverse for one image and TL mode columns.
//
// The kept rows are k = 0..m1-1 (row frequency +k) and k = m1..2*m1-1 (frequency k - 2*m1 = -m1..-1): rows f and
// K - f are a conjugate pair that shares cos / sin of phi = 2*pi*f*h/hp.  Both transforms work per FREQUENCY
// f = 0..m1 (F = m1 + 1 of them) instead of per kept row, which halves their multiply-adds:
//   forward   A_f = sum_h x_h cos, B_f = sum_h x_h sin  (4 FMA per h for the pair);  X[f] = A - iB,  X[K-f] = A + iB
//   inverse   z_h = sum_f P_f cos + Q_f sin             (4 FMA per f for the pair);  P = Y[f] + Y[K-f],  Q = i (Y[f] - Y[K-f])
// (f = 0 has only row 0 and f = m1 only row m1: the missing partner counts as zero.)
// ===========================================================================
struct CoreParams {
  const float2* in; float2* out; float2* spec_out;
  const float2* w1; const float2* w2;
  const float2* wt;      // optional mode-major copy [m2][K][ci][co] (few-image regime, see spec_weights_mode_major)
  const float2* t_hf; const float2* t_fh;
  const float* pre; const float* post;
  int ca, cb, co_layer, hp, hp8, m1, m2, K, F, Fp, TL;
  int tables_global;     // large hp x F: the DFT tables do not fit shared memory and are read through L1/L2
  int wstage;            // core2d_kernel with TL = 1: the block's column of wt is staged in shared memory
  int onetab;            // core2d_kernel: the inverse transform reads the h-major table too (t_fh is not staged)
};

// G = frequencies (phase 1) / spatial rows (phase 3) accumulated per work item.  Small G gives more
// items per block (the few-image heads need that to keep 256 threads busy), large G more FMAs per
// shared/L1 load (the many-image per-snapshot net).
template <int G>
__device__ __forceinline__ void pacc_freqs(const float2 x, const float2* __restrict__ t, float (&ar)[G], float (&ai)[G],
                                           float (&br)[G], float (&bi)[G]) {
  if constexpr (G == 1) {
    const float2 cs = *t;
    ar[0] = fmaf(x.x, cs.x, ar[0]); ai[0] = fmaf(x.y, cs.x, ai[0]);
    br[0] = fmaf(x.x, cs.y, br[0]); bi[0] = fmaf(x.y, cs.y, bi[0]);
  } else {
    const float4* t4 = reinterpret_cast<const float4*>(t);
#pragma unroll
    for (int q = 0; q < G / 2; ++q) {
      const float4 cs = t4[q];
      ar[2 * q] = fmaf(x.x, cs.x, ar[2 * q]);         ai[2 * q] = fmaf(x.y, cs.x, ai[2 * q]);
      br[2 * q] = fmaf(x.x, cs.y, br[2 * q]);         bi[2 * q] = fmaf(x.y, cs.y, bi[2 * q]);
      ar[2 * q + 1] = fmaf(x.x, cs.z, ar[2 * q + 1]); ai[2 * q + 1] = fmaf(x.y, cs.z, ai[2 * q + 1]);
      br[2 * q + 1] = fmaf(x.x, cs.w, br[2 * q + 1]); bi[2 * q + 1] = fmaf(x.y, cs.w, bi[2 * q + 1]);
    }
  }
}

// pq = (P.re, P.im, Q.re, Q.im) of one frequency; z_h += P cos + Q sin for G rows h
template <int G>
__device__ __forceinline__ void pacc_rows(const float4 pq, const float2* __restrict__ t, float (&re)[G], float (&im)[G]) {
  if constexpr (G == 1) {
    const float2 cs = *t;
    re[0] = fmaf(pq.x, cs.x, fmaf(pq.z, cs.y, re[0]));
    im[0] = fmaf(pq.y, cs.x, fmaf(pq.w, cs.y, im[0]));
  } else {
    const float4* t4 = reinterpret_cast<const float4*>(t);
#pragma unroll
    for (int q = 0; q < G / 2; ++q) {
      const float4 cs = t4[q];
      re[2 * q] = fmaf(pq.x, cs.x, fmaf(pq.z, cs.y, re[2 * q]));
      im[2 * q] = fmaf(pq.y, cs.x, fmaf(pq.w, cs.y, im[2 * q]));
      re[2 * q + 1] = fmaf(pq.x, cs.z, fmaf(pq.z, cs.w, re[2 * q + 1]));
      im[2 * q + 1] = fmaf(pq.y, cs.z, fmaf(pq.w, cs.w, im[2 * q + 1]));
    }
  }
}

// rows of the conjugate pair of frequency f: X[f] = A - iB (f < m1), X[K-f] = A + iB (f > 0), scaled
__device__ __forceinline__ float2 pair_lo(float ar, float ai, float br, float bi, float sc) {
  return make_float2((ar + bi) * sc, (ai - br) * sc);
}
__device__ __forceinline__ float2 pair_hi(float ar, float ai, float br, float bi, float sc) {
  return make_float2((ar - bi) * sc, (ai + br) * sc);
}
// (P, Q) of the pair (Y[f], Y[K-f]); a missing partner is zero
__device__ __forceinline__ float4 pair_pq(const float2 y1, const float2 y2) {
  return make_float4(y1.x + y2.x, y1.y + y2.y, y2.y - y1.y, y1.x - y2.x);
}

// The three inner loops of the middle stage, on running pointers (every stride is a run-time value: indexing by
// h * pitch costs an IMAD + LEA per load, ncu source view) and inlined once per address space so that the shared
// memory tables are read with LDS (a pointer that may be global or shared compiles to generic LD).
template <int G>
__device__ __forceinline__ void h_forward_freqs(const float2* __restrict__ x, int xpitch, const float2* __restrict__ t, int tpitch,
                                                int hp, float (&ar)[G], float (&ai)[G], float (&br)[G], float (&bi)[G]) {
#pragma unroll 4
  for (int h = 0; h < hp; ++h) {
    pacc_freqs<G>(*x, t, ar, ai, br, bi);
    x += xpitch;
    t += tpitch;
  }
}

template <int G>
__device__ __forceinline__ void h_inverse_rows(const float4* __restrict__ pq, int ppitch, const float2* __restrict__ t, int tpitch,
                                               int F, float (&re)[G], float (&im)[G]) {
#pragma unroll 4
  for (int f = 0; f < F; ++f) {
    pacc_rows<G>(*pq, t, re, im);
    pq += ppitch;
    t += tpitch;
  }
}

// the same from the h-major table: row j of the item is t[j] (one 8-byte load per row instead of 16 bytes per row pair)
template <int G>
__device__ __forceinline__ void h_inverse_rows_hmajor(const float4* __restrict__ pq, int ppitch, const float2* const (&t)[G],
                                                      int F, float (&re)[G], float (&im)[G]) {
#pragma unroll 4
  for (int f = 0; f < F; ++f) {
    const float4 v = *pq;
    pq += ppitch;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const float2 cs = t[j][f];
      re[j] = fmaf(v.x, cs.x, fmaf(v.z, cs.y, re[j]));
      im[j] = fmaf(v.y, cs.x, fmaf(v.w, cs.y, im[j]));
    }
  }
}

// y = sum_a x_a W[a][b] (fwd) or sum_a x_a conj(W[b][a]) (bwd): x at stride xstride, W at stride wstride (float2 units)
template <bool BWD>
__device__ __forceinline__ float2 mix_row(const float2* __restrict__ x, int xstride, const float2* __restrict__ w, int wstride,
                                          int ca) {
  float yr = 0.f, yi = 0.f;
#pragma unroll 6
  for (int a = 0; a < ca; ++a) {
    const float2 xv = *x, wv = *w;
    x += xstride;
    w += wstride;
    if (!BWD) {
      yr = fmaf(xv.x, wv.x, fmaf(-xv.y, wv.y, yr));
      yi = fmaf(xv.x, wv.y, fmaf(xv.y, wv.x, yi));
    } else {
      yr = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, yr));
      yi = fmaf(xv.y, wv.x, fmaf(-xv.x, wv.y, yi));
    }
  }
  return make_float2(yr, yi);
}

template <bool BWD, int G1, int G3>   // frequencies per phase-1 item, spatial rows per phase-3 item
__global__ void __launch_bounds__(1024) core2d_kernel(const CoreParams p) {
  extern __shared__ __align__(16) float smem[];
  const int TL = p.TL, Pa = p.ca * TL, Pb = p.cb * TL, K = p.K, F = p.F, hp = p.hp, m2 = p.m2;
  const int nA = (max(hp * Pa, 2 * F * Pb) + 1) & ~1;
  float2* bufA = reinterpret_cast<float2*>(smem);            // phase 0/1: x[h][pa]; phase 2/3: (P, Q)[f][pb] as float4
  float4* bufPQ = reinterpret_cast<float4*>(smem);
  float2* bufX = bufA + nA;
  float2* tab = bufX + K * Pa + ((K * Pa) & 1);
  const bool tg = p.tables_global != 0;
  float2* tab_hf = tab;                                    // [hp][Fp], 16-byte aligned rows
  float2* tab_fh = tab + hp * p.Fp;                        // [F][hp8]  (not with onetab)
  const int fh_n = p.onetab ? 0 : F * p.hp8;
  float2* wcol = tab + (tg ? 0 : hp * p.Fp + fh_n);        // [K][ci][co]: this block's column of the mode-major weights
  const int wcol_n = p.wstage ? K * p.ca * p.cb : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wcol + wcol_n);   // [0] t_hf, [1] t_fh, [2] weights
  const int l0 = blockIdx.x * TL, b = blockIdx.y;
  const int tid = threadIdx.x, nt = blockDim.x;

  // The two DFT tables live in shared memory for the block's lifetime (they are re-read by every item; from L1/L2 the
  // inner loops were latency-bound: ncu long-scoreboard 20 cycles per issue), and with one mode column per block so
  // does the column's slice of the mode-major weights (the mix was a chain of dependent L2 reads).  Each is contiguous
  // in HBM: three bulk async copies, each with its own barrier, in flight while phase 0 stages the image; a phase
  // waits only for the operand it reads.
  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init_fence();
    if (!tg) {
      const uint32_t b1 = (uint32_t)(hp * p.Fp) * 8u, b2 = (uint32_t)(F * p.hp8) * 8u;
      mbar_expect_tx(&bars[0], b1);
      bulk_g2s(tab_hf, p.t_hf, b1, &bars[0]);
      if (!p.onetab) {
        mbar_expect_tx(&bars[1], b2);
        bulk_g2s(tab_fh, p.t_fh, b2, &bars[1]);
      }
    }
  }
  pdl_wait();      // the tables above are constant plan data; everything below reads the previous kernel's output
  if (tid == 0 && p.wstage) {     // (the mode-major weights are written by a kernel earlier in the step)
    const uint32_t bw = (uint32_t)wcol_n * 8u;
    mbar_expect_tx(&bars[2], bw);
    bulk_g2s(wcol, p.wt + (size_t)l0 * wcol_n, bw, &bars[2]);
  }
  // phase 0: stage the image's TL columns, all channels: bufA[h][a*TL + lt]
  if (TL == 1) {
    const float inv_hp = 1.0f / (float)hp;
    for (int idx = tid; idx < p.ca * hp; idx += nt) {
      const int a = __float2int_rz(((float)idx + 0.5f) * inv_hp), h = idx - a * hp;      // exact for these ranges
      bufA[h * Pa + a] = __ldg(p.in + ((size_t)(b * p.ca + a) * hp + h) * m2 + l0);
    }
  } else {
    for (int idx = tid; idx < p.ca * hp * TL; idx += nt) {
      const int lt = idx % TL, h = (idx / TL) % hp, a = idx / (TL * hp);
      const int l = l0 + lt;
      bufA[h * Pa + a * TL + lt] =
          l < m2 ? __ldg(p.in + ((size_t)(b * p.ca + a) * hp + h) * m2 + l) : make_float2(0.f, 0.f);
    }
  }
  __syncthreads();          // also publishes the mbarrier inits to the waiting threads
  if (!tg) mbar_wait(&bars[0], 0);

  // phase 1: X[k][pa] = pre[l] * sum_h x[h][pa] * e^{-i phi_kh}, G1 frequencies (row pairs) per item
  constexpr int G = G1;
  const int nfg = (F + G - 1) / G;
  for (int idx = tid; idx < Pa * nfg; idx += nt) {
    const int fg = idx / Pa, pa = idx - fg * Pa;
    float ar[G], ai[G], br[G], bi[G];
#pragma unroll
    for (int j = 0; j < G; ++j) ar[j] = ai[j] = br[j] = bi[j] = 0.f;
    if (tg) h_forward_freqs<G>(bufA + pa, Pa, p.t_hf + fg * G, p.Fp, hp, ar, ai, br, bi);
    else h_forward_freqs<G>(bufA + pa, Pa, tab_hf + fg * G, p.Fp, hp, ar, ai, br, bi);
    const int a = pa / TL, lt = pa - a * TL, l = l0 + lt;
    const float sc = l < m2 ? __ldg(p.pre + l) : 0.f;
    float2* so = (p.spec_out != nullptr && l < m2) ? p.spec_out + (size_t)(b * p.ca + a) * K * m2 + l : nullptr;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int f = fg * G + j;
      if (f < p.m1) {
        const float2 v = pair_lo(ar[j], ai[j], br[j], bi[j], sc);
        bufX[f * Pa + pa] = v;
        if (so != nullptr) so[(size_t)f * m2] = v;
      }
      if (f > 0 && f <= p.m1) {
        const float2 v = pair_hi(ar[j], ai[j], br[j], bi[j], sc);
        bufX[(K - f) * Pa + pa] = v;
        if (so != nullptr) so[(size_t)(K - f) * m2] = v;
      }
    }
  }
  __syncthreads();
  pdl_trigger_late();

  // phase 2: per-mode channel mix of both rows of a frequency -> (P, Q).
  // fwd: y_b = sum_a x_a W[a][b]; bwd: y_b = sum_a x_a conj(W[b][a])
  if (p.wt != nullptr) {
    // mode-major weights: the K*ci*co coefficients of one mode column are contiguous, so a block that owns
    // few columns reads them with full sectors (the parameter layout has the mode index innermost: one
    // 8-byte element per 32-byte sector for a single column)
    const int ci = BWD ? p.cb : p.ca, co = p.co_layer;
    const int wa = BWD ? 1 : co, wb = BWD ? co : 1;        // strides of the summed / the produced channel
    if (p.wstage) mbar_wait(&bars[2], 0);
    for (int idx = tid; idx < F * Pb; idx += nt) {
      const int bc = idx % p.cb, f = (idx / p.cb) % F, lt = idx / (p.cb * F);
      const int l = l0 + lt;
      float2 y[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (l < m2) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (r == 0 ? f >= p.m1 : f == 0) continue;
          const int k = r == 0 ? f : K - f;
          if (p.wstage) y[r] = mix_row<BWD>(bufX + k * Pa + lt, TL, wcol + (size_t)k * ci * co + bc * wb, wa, p.ca);
          else y[r] = mix_row<BWD>(bufX + k * Pa + lt, TL, p.wt + ((size_t)l * K + k) * ci * co + bc * wb, wa, p.ca);
        }
      }
      bufPQ[f * Pb + bc * TL + lt] = pair_pq(y[0], y[1]);
    }
  } else {
    const size_t cstride = (size_t)p.m1 * m2;
    const int wa = (int)((BWD ? 1 : p.co_layer) * cstride), wb = (int)((BWD ? p.co_layer : 1) * cstride);
    for (int idx = tid; idx < F * Pb; idx += nt) {
      const int lt = idx % TL, f = (idx / TL) % F, bc = idx / (TL * F);
      const int l = l0 + lt;
      float2 y[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (l < m2) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (r == 0 ? f >= p.m1 : f == 0) continue;
          const int k = r == 0 ? f : K - f;
          const float2* wsel = r == 0 ? p.w1 : p.w2;           // row k >= m1 is row k - m1 of weights2
          const size_t mode_off = (size_t)(r == 0 ? f : p.m1 - f) * m2 + l;
          y[r] = mix_row<BWD>(bufX + k * Pa + lt, TL, wsel + (size_t)bc * wb + mode_off, wa, p.ca);
        }
      }
      bufPQ[f * Pb + bc * TL + lt] = pair_pq(y[0], y[1]);
    }
  }
  __syncthreads();
  if (!tg && !p.onetab) mbar_wait(&bars[1], 0);

  // phase 3: Z[h][pb] = post[l] * sum_k y[k][pb] * e^{+i phi_kh} = post[l] * sum_f P cos + Q sin, G3 rows per item
  const int nhg = (hp + G3 - 1) / G3;
  for (int idx = tid; idx < Pb * nhg; idx += nt) {
    const int hg = idx / Pb, pb = idx - hg * Pb;
    float re[G3], im[G3];
#pragma unroll
    for (int j = 0; j < G3; ++j) re[j] = im[j] = 0.f;
    if (tg) {
      h_inverse_rows<G3>(bufPQ + pb, Pb, p.t_fh + hg * G3, p.hp8, F, re, im);
    } else if (p.onetab) {
      const float2* rows[G3];
#pragma unroll
      for (int j = 0; j < G3; ++j) rows[j] = tab_hf + min(hg * G3 + j, hp - 1) * p.Fp;
      h_inverse_rows_hmajor<G3>(bufPQ + pb, Pb, rows, F, re, im);
    } else {
      h_inverse_rows<G3>(bufPQ + pb, Pb, tab_fh + hg * G3, p.hp8, F, re, im);
    }
    const int bc = pb / TL, lt = pb - bc * TL, l = l0 + lt;
    if (l >= m2) continue;
    const float sc = __ldg(p.post + l);
    float2* dst = p.out + ((size_t)(b * p.cb + bc) * hp + hg * G3) * m2 + l;
#pragma unroll
    for (int j = 0; j < G3; ++j)
      if (hg * G3 + j < hp) dst[(size_t)j * m2] = make_float2(re[j] * sc, im[j] * sc);
  }
}

// ---------------------------------------------------------------------------
// 2-D middle stage, streaming variant for the many-image regime (the per-snapshot net).  Persistent
// blocks; a block owns whole images: the W-transformed image [ca][hp][m2] is ONE contiguous span of HBM,
// brought in by a single bulk async copy into a double-buffered shared tile (the copy of image i+1 is
// in flight while image i is transformed); the two DFT tables and nothing else are staged once per
// block.  Phases as in core2d_kernel.
// ---------------------------------------------------------------------------
template <bool BWD, int G1, int G3>
__global__ void __launch_bounds__(1024) core2d_stream_kernel(const CoreParams p, int images) {
  extern __shared__ __align__(16) float smem[];
  const int K = p.K, F = p.F, hp = p.hp, m2 = p.m2, Pa = p.ca * m2, Pb = p.cb * m2;
  const int in_elems = p.ca * hp * m2;                      // float2 per image (even: m2*hp*ca*... padded below)
  const int in_pad = (in_elems + 1) & ~1;
  float2* in0 = reinterpret_cast<float2*>(smem);
  float2* in1 = in0 + in_pad;
  float2* bufX = in1 + in_pad;                              // [K][Pa]
  float4* bufPQ = reinterpret_cast<float4*>(bufX + ((K * Pa + 1) & ~1));   // [F][Pb] (P, Q)
  float2* s_hf = reinterpret_cast<float2*>(bufPQ + F * Pb); // [hp][Fp]
  float2* s_fh = s_hf + hp * p.Fp;                          // [F][hp8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_fh + F * p.hp8);   // [0],[1]: image stages, [2]: tables
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t img_bytes = (uint32_t)in_elems * 8u;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init_fence();
    const uint32_t b1 = (uint32_t)(hp * p.Fp) * 8u, b2 = (uint32_t)(F * p.hp8) * 8u;
    mbar_expect_tx(&bars[2], b1 + b2);
    bulk_g2s(s_hf, p.t_hf, b1, &bars[2]);
    bulk_g2s(s_fh, p.t_fh, b2, &bars[2]);
  }
  pdl_wait();
  __syncthreads();
  auto issue = [&](int img, int stage) {     // thread 0 only
    fence_proxy_async();
    mbar_expect_tx(&bars[stage], img_bytes);
    bulk_g2s(stage ? in1 : in0, p.in + (size_t)img * in_elems, img_bytes, &bars[stage]);
  };
  if (tid == 0 && (int)blockIdx.x < images) issue(blockIdx.x, 0);
  mbar_wait(&bars[2], 0);

  int it = 0;
  for (int b = blockIdx.x; b < images; b += gridDim.x, ++it) {
    const int stage = it & 1;
    if (tid == 0 && b + (int)gridDim.x < images) issue(b + gridDim.x, stage ^ 1);
    mbar_wait(&bars[stage], (it >> 1) & 1);
    if (b + (int)gridDim.x >= images) pdl_trigger_late();
    const float2* xin = stage ? in1 : in0;                  // [a][h][l]

    // phase 1: X[k][(a,l)] = pre[l] * sum_h x[a][h][l] * e^{-i phi_kh}, by frequency (row pairs)
    const int nfg = (F + G1 - 1) / G1;
    for (int idx = tid; idx < Pa * nfg; idx += nt) {
      const int pa = idx % Pa, fg = idx / Pa;
      const int a = pa / m2, l = pa - a * m2;
      float ar[G1], ai[G1], br[G1], bi[G1];
#pragma unroll
      for (int j = 0; j < G1; ++j) ar[j] = ai[j] = br[j] = bi[j] = 0.f;
      const float2* trow = s_hf + fg * G1;
      const float2* xcol = xin + (size_t)a * hp * m2 + l;
#pragma unroll 4
      for (int h = 0; h < hp; ++h) pacc_freqs<G1>(xcol[h * m2], trow + (size_t)h * p.Fp, ar, ai, br, bi);
      const float sc = __ldg(p.pre + l);
      float2* so = p.spec_out != nullptr ? p.spec_out + (size_t)(b * p.ca + a) * K * m2 + l : nullptr;
#pragma unroll
      for (int j = 0; j < G1; ++j) {
        const int f = fg * G1 + j;
        if (f < p.m1) {
          const float2 v = pair_lo(ar[j], ai[j], br[j], bi[j], sc);
          bufX[f * Pa + pa] = v;
          if (so != nullptr) so[(size_t)f * m2] = v;
        }
        if (f > 0 && f <= p.m1) {
          const float2 v = pair_hi(ar[j], ai[j], br[j], bi[j], sc);
          bufX[(K - f) * Pa + pa] = v;
          if (so != nullptr) so[(size_t)(K - f) * m2] = v;
        }
      }
    }
    __syncthreads();

    // phase 2: per-mode channel mix of both rows of a frequency -> (P, Q)
    for (int idx = tid; idx < F * Pb; idx += nt) {
      const int l = idx % m2, f = (idx / m2) % F, bc = idx / (m2 * F);
      const size_t cstride = (size_t)p.m1 * m2;
      float2 y[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 0 ? f >= p.m1 : f == 0) continue;
        const int k = r == 0 ? f : K - f;
        const float2* wsel = r == 0 ? p.w1 : p.w2;             // row k >= m1 is row k - m1 of weights2
        const size_t mode_off = (size_t)(r == 0 ? f : p.m1 - f) * m2 + l;
        float yr = 0.f, yi = 0.f;
#pragma unroll 4
        for (int a = 0; a < p.ca; ++a) {
          const float2 x = bufX[k * Pa + a * m2 + l];
          if (!BWD) {
            const float2 w = __ldg(wsel + (size_t)(a * p.co_layer + bc) * cstride + mode_off);
            yr = fmaf(x.x, w.x, fmaf(-x.y, w.y, yr));
            yi = fmaf(x.x, w.y, fmaf(x.y, w.x, yi));
          } else {
            const float2 w = __ldg(wsel + (size_t)(bc * p.co_layer + a) * cstride + mode_off);
            yr = fmaf(x.x, w.x, fmaf(x.y, w.y, yr));
            yi = fmaf(x.y, w.x, fmaf(-x.x, w.y, yi));
          }
        }
        y[r] = make_float2(yr, yi);
      }
      bufPQ[f * Pb + bc * m2 + l] = pair_pq(y[0], y[1]);
    }
    __syncthreads();

    // phase 3: Z[(bc)][h][l] = post[l] * sum_f P cos + Q sin
    const int nhg = (hp + G3 - 1) / G3;
    for (int idx = tid; idx < Pb * nhg; idx += nt) {
      const int pb = idx % Pb, hg = idx / Pb;
      float re[G3], im[G3];
#pragma unroll
      for (int j = 0; j < G3; ++j) re[j] = im[j] = 0.f;
      const float2* trow = s_fh + hg * G3;
#pragma unroll 4
      for (int f = 0; f < F; ++f) pacc_rows<G3>(bufPQ[f * Pb + pb], trow + (size_t)f * p.hp8, re, im);
      const int bc = pb / m2, l = pb - bc * m2;
      const float sc = __ldg(p.post + l);
#pragma unroll
      for (int j = 0; j < G3; ++j) {
        const int h = hg * G3 + j;
        if (h < hp) p.out[((size_t)(b * p.cb + bc) * hp + h) * m2 + l] = make_float2(re[j] * sc, im[j] * sc);
      }
    }
    fence_proxy_async();
    __syncthreads();          // bufX / bufPQ and this image stage are free again
  }
}

template <bool BWD, int G1, int G3>
static void launch_core2d_stream_t(const CoreParams& p, int images, int grid, int threads, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(core2d_stream_kernel<BWD, G1, G3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  launch_k(core2d_stream_kernel<BWD, G1, G3>, dim3(grid), dim3(threads), smem, st, p, images);
}

template <bool BWD, int G1, int G3>
static void launch_core2d_t(const CoreParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(core2d_kernel<BWD, G1, G3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  launch_k(core2d_kernel<BWD, G1, G3>, dim3(grid), dim3(threads), smem, st, p);
}

// Mode-major copy of the spectral weights of up to BDN_MAX_LAYERS layers: wt[layer][l][k][i][o] (k over
// the 2*m1 kept rows, weights1 then weights2) from the parameter layout [i][o][m1][m2] (x2 tensors).
struct WtParams {
  const float2* w1[8]; const float2* w2[8];
  float2* wt; int ci, co, m1, m2, n_layers;
};

__global__ void spec_weights_mode_major_kernel(const WtParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int K = 2 * p.m1;
  const long per_layer = (long)p.m2 * K * p.ci * p.co;
  const int layer = blockIdx.y;
  const float2* w1 = p.w1[layer];
  const float2* w2 = p.w2[layer];
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < per_layer; idx += (long)gridDim.x * blockDim.x) {
    const int o = idx % p.co, i = (idx / p.co) % p.ci, k = (idx / ((long)p.co * p.ci)) % K,
              l = idx / ((long)p.co * p.ci * K);
    const bool lo = k < p.m1;
    const int kk = lo ? k : k - p.m1;
    p.wt[layer * per_layer + idx] = __ldg((lo ? w1 : w2) + ((size_t)(i * p.co + o) * p.m1 + kk) * p.m2 + l);
  }
}

void launch_spec_weights_mode_major(const Plan* pl, const float* const* w1, const float* const* w2, int n_layers,
                                    int ci, int co, float2* wt, cudaStream_t st) {
  LaunchScope scope("spec_w_mode_major", st, co);
  WtParams p;
  for (int k = 0; k < 8; ++k) {
    p.w1[k] = k < n_layers ? reinterpret_cast<const float2*>(w1[k]) : nullptr;
    p.w2[k] = k < n_layers ? reinterpret_cast<const float2*>(w2[k]) : nullptr;
  }
  p.wt = wt; p.ci = ci; p.co = co; p.m1 = pl->m1; p.m2 = pl->m2; p.n_layers = n_layers;
  const long per_layer = (long)pl->m2 * pl->K * ci * co;
  dim3 grid((unsigned)((per_layer + 255) / 256 < 148 * 8 ? (per_layer + 255) / 256 : 148 * 8), n_layers);
  launch_k(spec_weights_mode_major_kernel, dim3(grid), dim3(256), 0, st, p);
}

void launch_core2d(const Plan* pl, const float2* in, float2* out, float2* spec_out, const float2* w1,
                   const float2* w2, int images, int ci_layer, int co_layer, bool bwd, cudaStream_t st,
                   const float2* wt) {
  LaunchScope scope(bwd ? "core2d_bwd" : "core2d_fwd", st, co_layer);
  CoreParams p;
  p.in = in; p.out = out; p.spec_out = spec_out; p.w1 = w1; p.w2 = w2; p.wt = wt;
  p.t_hf = pl->t_hf; p.t_fh = pl->t_fh;
  p.pre = bwd ? pl->col_fwd : pl->col_dc;
  p.post = bwd ? pl->col_dc : pl->col_fwd;
  p.ca = bwd ? co_layer : ci_layer;
  p.cb = bwd ? ci_layer : co_layer;
  p.co_layer = co_layer;
  p.hp = pl->hp; p.hp8 = pl->hp8; p.m1 = pl->m1; p.m2 = pl->m2; p.K = pl->K; p.F = pl->F; p.Fp = pl->Fp;
  const size_t table_f2 = (size_t)pl->hp * pl->Fp + (size_t)pl->F * pl->hp8;
  p.tables_global = table_f2 * sizeof(float2) > 150 * 1024;    // e.g. 320 x 128: 337 KB of tables
  // many images: persistent streaming kernel, whole images per block (G1 = 2 frequencies, G3 = 8 rows per item)
  {
    const size_t in_pad = ((size_t)p.ca * pl->hp * pl->m2 + 1) & ~(size_t)1;
    const size_t nX = ((size_t)pl->K * p.ca * pl->m2 + 1) & ~(size_t)1, nPQ = 2 * (size_t)pl->F * p.cb * pl->m2;
    const size_t smem_s = (2 * in_pad + nX + nPQ + (size_t)pl->hp * pl->Fp + (size_t)pl->F * pl->hp8) * sizeof(float2) + 32;
    const bool aligned = (((size_t)p.ca * pl->hp * pl->m2 * 8) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (images >= 148 && smem_s <= 110 * 1024 && aligned) {
      p.TL = pl->m2;
      p.wstage = 0; p.onetab = 0;
      const int Pa_s = p.ca * pl->m2, Pb_s = p.cb * pl->m2;
      const int items1 = Pa_s * ceil_div(pl->F, 2), items3 = Pb_s * ceil_div(pl->hp, 8);
      int threads = items1 > items3 ? items1 : items3;
      threads = ceil_div(threads, ceil_div(threads, 1024));
      threads = (threads + 31) & ~31;
      if (threads < 128) threads = 128;
      const int per_sm = (int)((220 * 1024) / (smem_s + 1024)) < 2048 / threads ? (int)((220 * 1024) / (smem_s + 1024)) : 2048 / threads;
      const int cap = 148 * (per_sm < 1 ? 1 : per_sm);
      const int grid_s = images < cap ? images : cap;
      if (bwd) launch_core2d_stream_t<true, 2, 8>(p, images, grid_s, threads, smem_s, st);
      else launch_core2d_stream_t<false, 2, 8>(p, images, grid_s, threads, smem_s, st);
      return;
    }
  }
  // TL mode columns per block: the widest column tile that still leaves >= 2 blocks per SM and fits
  // shared memory (wide tiles read the W-transformed image with full 32-byte sectors)
  static const int onetab_knob = [] { const char* e = getenv("BDN_CORE_ONETAB"); return e ? atoi(e) : 1; }();   // (tuning knob)
  auto smem_of = [&](int t) {
    const size_t a0 = (size_t)pl->hp * p.ca, a1 = 2 * (size_t)pl->F * p.cb;
    size_t nA = (a0 > a1 ? a0 : a1) * t;
    nA = (nA + 1) & ~(size_t)1;
    const size_t nX = (size_t)pl->K * p.ca * t;
    return (nA + nX + (nX & 1) + (p.tables_global ? 0 : table_f2)) * sizeof(float2) + 32;
  };
  int tl = 1;
  for (int parts = 1; parts <= pl->m2; ++parts) {
    const int cand = ceil_div(pl->m2, parts);
    if (smem_of(cand) > 100 * 1024) continue;
    if ((long)images * ceil_div(pl->m2, cand) >= 2 * 148 || cand == 1) { tl = cand; break; }
  }
  p.TL = tl;
  size_t smem = smem_of(tl);
  // one mode column per block: its K*ci*co slice of the mode-major weights is contiguous -> staged by one bulk copy
  const size_t wcol_bytes = (size_t)pl->K * p.ca * p.cb * sizeof(float2);
  static const int wstage_knob = [] { const char* e = getenv("BDN_CORE_WSTAGE"); return e ? atoi(e) : 1; }();   // (tuning knob)
  p.wstage = wstage_knob && wt != nullptr && tl == 1 && smem + wcol_bytes <= 200 * 1024 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0;
  if (p.wstage) smem += wcol_bytes;
  // with the weights staged a block holds > 113 KB: dropping the f-major table (the inverse transform then reads the
  // h-major one) lets two blocks -- one of each output head, which run side by side -- share an SM
  p.onetab = onetab_knob && p.wstage && !p.tables_global && smem > 113 * 1024;
  if (p.onetab) smem -= (size_t)pl->F * pl->hp8 * sizeof(float2);
  dim3 grid(ceil_div(pl->m2, tl), images);
  // Work items: phase 1 has Pa * ceil(F / G1), phase 3 has Pb * ceil(hp / G3).  With few images (the
  // heads) every block should run as many threads as it has items (latency-bound); with many images
  // (the per-snapshot net) larger G gives more FMAs per shared-memory load.
  const int Pa = p.ca * tl, Pb = p.cb * tl;
  int g1 = 1, g3 = 1;
  const int target = 256;
  for (int cand = 4; cand >= 1; cand >>= 1)
    if (Pa * ceil_div(pl->F, cand) >= target || cand == 1) { g1 = cand; break; }
  for (int cand = 8; cand >= 1; cand >>= 1)
    if (Pb * ceil_div(pl->hp, cand) >= target || cand == 1) { g3 = cand; break; }
  const int items1 = Pa * ceil_div(pl->F, g1), items3 = Pb * ceil_div(pl->hp, g3);
  const int most = items1 > items3 ? items1 : items3;
  int threads = ceil_div(most, ceil_div(most, 1024));      // balanced rounds when one block cannot hold all items
  threads = (threads + 31) & ~31;
  if (threads < 128) threads = 128;
  if (threads > 1024) threads = 1024;
#define BDN_CORE3(B, A1)                                                              \
  {                                                                                   \
    if (g3 == 8) launch_core2d_t<B, A1, 8>(p, grid, threads, smem, st);               \
    else if (g3 == 4) launch_core2d_t<B, A1, 4>(p, grid, threads, smem, st);          \
    else if (g3 == 2) launch_core2d_t<B, A1, 2>(p, grid, threads, smem, st);          \
    else launch_core2d_t<B, A1, 1>(p, grid, threads, smem, st);                       \
  }
#define BDN_CORE(B)                                                                   \
  {                                                                                   \
    if (g1 == 4) BDN_CORE3(B, 4) else if (g1 == 2) BDN_CORE3(B, 2) else BDN_CORE3(B, 1) \
  }
  if (bwd) BDN_CORE(true) else BDN_CORE(false)
#undef BDN_CORE
#undef BDN_CORE3
}

// ===========================================================================
// 1-D middle stage: column scale, save, per-mode mix, column scale.
// ===========================================================================
template <bool BWD>
__global__ void mix1d_kernel(const float2* __restrict__ in, float2* __restrict__ out, float2* __restrict__ spec_out,
                             const float2* __restrict__ w, const float* __restrict__ pre,
                             const float* __restrict__ post, int images, int ca, int cb, int co_layer, int m2) {
  pdl_launch_dependents();
  pdl_wait();
  const int cmax = ca > cb ? ca : cb;
  const long total = (long)images * cmax * m2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int l = idx % m2, c = (idx / m2) % cmax, b = idx / ((long)m2 * cmax);
    const float ps = __ldg(pre + l);
    if (spec_out != nullptr && c < ca) {
      const float2 x = __ldg(in + ((size_t)b * ca + c) * m2 + l);
      spec_out[((size_t)b * ca + c) * m2 + l] = make_float2(x.x * ps, x.y * ps);
    }
    if (c < cb) {
      float yr = 0.f, yi = 0.f;
      for (int a = 0; a < ca; ++a) {
        float2 x = __ldg(in + ((size_t)b * ca + a) * m2 + l);
        x.x *= ps; x.y *= ps;
        if (!BWD) {
          const float2 wv = __ldg(w + (size_t)(a * co_layer + c) * m2 + l);
          yr = fmaf(x.x, wv.x, fmaf(-x.y, wv.y, yr));
          yi = fmaf(x.x, wv.y, fmaf(x.y, wv.x, yi));
        } else {
          const float2 wv = __ldg(w + (size_t)(c * co_layer + a) * m2 + l);
          yr = fmaf(x.x, wv.x, fmaf(x.y, wv.y, yr));
          yi = fmaf(x.y, wv.x, fmaf(-x.x, wv.y, yi));
        }
      }
      const float qs = __ldg(post + l);
      out[((size_t)b * cb + c) * m2 + l] = make_float2(yr * qs, yi * qs);
    }
  }
}

void launch_mix1d(const Plan* pl, const float2* in, float2* out, float2* spec_out, const float2* w, int images,
                  int ci_layer, int co_layer, bool bwd, cudaStream_t st) {
  LaunchScope scope(bwd ? "mix1d_bwd" : "mix1d_fwd", st, co_layer);
  const int ca = bwd ? co_layer : ci_layer, cb = bwd ? ci_layer : co_layer;
  const int cmax = ca > cb ? ca : cb;
  const long total = (long)images * cmax * pl->m2;
  const int block = 256;
  const int grid = (int)((total + block - 1) / block);
  if (bwd)
    launch_k(mix1d_kernel<true>, dim3(grid), dim3(block), 0, st, in, out, spec_out, w, pl->col_fwd, pl->col_dc, images, ca, cb,
                                               co_layer, pl->m2);
  else
    launch_k(mix1d_kernel<false>, dim3(grid), dim3(block), 0, st, in, out, spec_out, w, pl->col_dc, pl->col_fwd, images, ca, cb,
                                                co_layer, pl->m2);
}

// ===========================================================================
// spectral weight gradient: gw[i,o,k,l] += sum_b conj(xs[b,i,k,l]) * gys[b,o,k,l]
// ===========================================================================
__global__ void gw_reduce_kernel(const float2* __restrict__ xs, const float2* __restrict__ gys, float2* gw1,
                                 float2* gw2, int images, int ci, int co, int K, int m1, int m2, int bchunk) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)ci * co * K * m2;
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = idx % m2, k = (idx / m2) % K, o = (idx / ((long)m2 * K)) % co, i = idx / ((long)m2 * K * co);
  const int b0 = blockIdx.y * bchunk, b1 = min(images, b0 + bchunk);
  float re = 0.f, im = 0.f;
  const size_t xoff = ((size_t)i * K + k) * m2 + l, goff = ((size_t)o * K + k) * m2 + l;
  const size_t xstride = (size_t)ci * K * m2, gstride = (size_t)co * K * m2;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float2 x = __ldg(xs + b * xstride + xoff);
    const float2 g = __ldg(gys + b * gstride + goff);
    re = fmaf(x.x, g.x, fmaf(x.y, g.y, re));
    im = fmaf(x.x, g.y, fmaf(-x.y, g.x, im));
  }
  const bool lo = (m1 == 0) || k < m1;
  const int kk = lo ? k : k - m1;
  const int mrows = m1 == 0 ? 1 : m1;
  float* dst = reinterpret_cast<float*>((lo ? gw1 : gw2) + ((size_t)(i * co + o) * mrows + kk) * m2 + l);
  if (gridDim.y == 1) {
    dst[0] += re; dst[1] += im;
  } else {
    atomicAdd(dst, re); atomicAdd(dst + 1, im);
  }
}

// Register-tiled form for channel counts that are multiples of 4: a thread owns one mode (k, l) and a 4 x 4 block of
// (input, output) channel pairs, so an image costs 8 loads for 16 complex multiply-adds (the plain kernel: 2 loads for
// one, every spectrum value re-read co times through L1/L2 -- 80 us per launch for the heads at batch 32).
__global__ void __launch_bounds__(128) gw_reduce_tiled_kernel(const float2* __restrict__ xs, const float2* __restrict__ gys,
                                                              float2* gw1, float2* gw2, int images, int ci, int co, int K,
                                                              int m1, int m2, int bchunk) {
  // block = 32 mode tiles (threadIdx.x) x 4 image sub-chunks (threadIdx.y): the sub-chunks' partial sums are joined in
  // shared memory, so a block issues one (atomic) update per output instead of four
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float part[4][32][33];
  const int nog = co >> 2, nig = ci >> 2;
  const long total = (long)nig * nog * K * m2;
  const long idx = blockIdx.x * 32L + threadIdx.x;
  const bool live = idx < total;
  const long id = live ? idx : 0;
  const int l = id % m2, k = (id / m2) % K, og = (id / ((long)m2 * K)) % nog, ig = id / ((long)m2 * K * nog);
  const int sub = threadIdx.y;
  const int b0 = blockIdx.y * bchunk, b1 = min(images, b0 + bchunk);
  const size_t plane = (size_t)K * m2;
  const float2* xp = xs + ((size_t)(b0 + sub) * ci + ig * 4) * plane + (size_t)k * m2 + l;
  const float2* gp = gys + ((size_t)(b0 + sub) * co + og * 4) * plane + (size_t)k * m2 + l;
  float re[4][4], im[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) re[a][c] = im[a][c] = 0.f;
  if (live) {
#pragma unroll 2
    for (int b = b0 + sub; b < b1; b += 4) {
      float2 x[4], g[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        x[a] = __ldg(xp + a * plane);
        g[a] = __ldg(gp + a * plane);
      }
      xp += 4 * (size_t)ci * plane;
      gp += 4 * (size_t)co * plane;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          re[a][c] = fmaf(x[a].x, g[c].x, fmaf(x[a].y, g[c].y, re[a][c]));
          im[a][c] = fmaf(x[a].x, g[c].y, fmaf(-x[a].y, g[c].x, im[a][c]));
        }
    }
  }
  pdl_trigger_late();
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      part[sub][threadIdx.x][(a * 4 + c) * 2] = re[a][c];
      part[sub][threadIdx.x][(a * 4 + c) * 2 + 1] = im[a][c];
    }
  __syncthreads();
  if (!live) return;
  const bool lo = (m1 == 0) || k < m1;
  const int kk = lo ? k : k - m1;
  const int mrows = m1 == 0 ? 1 : m1;
  float2* base = lo ? gw1 : gw2;
#pragma unroll
  for (int q = 0; q < 4; ++q) {            // this thread finishes (a, c) pairs sub * 4 + q
    const int pr = sub * 4 + q, a = pr >> 2, c = pr & 3;
    const float vr = (part[0][threadIdx.x][2 * pr] + part[1][threadIdx.x][2 * pr]) +
                     (part[2][threadIdx.x][2 * pr] + part[3][threadIdx.x][2 * pr]);
    const float vi = (part[0][threadIdx.x][2 * pr + 1] + part[1][threadIdx.x][2 * pr + 1]) +
                     (part[2][threadIdx.x][2 * pr + 1] + part[3][threadIdx.x][2 * pr + 1]);
    float* dst = reinterpret_cast<float*>(base + ((size_t)((ig * 4 + a) * co + og * 4 + c) * mrows + kk) * m2 + l);
    if (gridDim.y == 1) {
      dst[0] += vr; dst[1] += vi;
    } else {
      atomicAdd(dst, vr); atomicAdd(dst + 1, vi);
    }
  }
}

void launch_gw_reduce(const Plan* pl, const float2* xs, const float2* gys, float2* gw1, float2* gw2, int images,
                      int ci, int co, cudaStream_t st) {
  LaunchScope scope("gw_reduce", st, co);
  static const int tiled_knob = [] { const char* e = getenv("BDN_GW_TILED"); return e ? atoi(e) : 1; }();   // (tuning knob)
  if (tiled_knob && (ci & 3) == 0 && (co & 3) == 0) {
    const long total = (long)(ci >> 2) * (co >> 2) * pl->K * pl->m2;
    const int gx = (int)((total + 31) / 32);
    // split the image loop further over blocks (atomic accumulation) when the modes alone do not fill the GPU:
    // >= 16 images per block (4 per sub-chunk)
    int chunks = 1;
    while (chunks < images && (long)gx * chunks < 2 * 148 && images / (chunks * 2) >= 16) chunks *= 2;
    const int bchunk = ceil_div(images, chunks);
    dim3 grid(gx, ceil_div(images, bchunk));
    launch_k(gw_reduce_tiled_kernel, dim3(grid), dim3(32, 4), 0, st, xs, gys, gw1, gw2, images, ci, co, pl->K, pl->m1, pl->m2,
             bchunk);
    return;
  }
  const long total = (long)ci * co * pl->K * pl->m2;
  const int block = 128;
  const int gx = (int)((total + block - 1) / block);
  // split the image loop when there are too few modes to fill the GPU
  int chunks = 1;
  while (chunks < images && (long)gx * chunks < 4 * 148 && images / (chunks * 2) >= 8) chunks *= 2;
  const int bchunk = ceil_div(images, chunks);
  dim3 grid(gx, ceil_div(images, bchunk));
  launch_k(gw_reduce_kernel, dim3(grid), dim3(block), 0, st, xs, gys, gw1, gw2, images, ci, co, pl->K, pl->m1, pl->m2, bchunk);
}

// ===========================================================================
// W-inverse with fused epilogue.
//   PLAIN      y = winv(z)
//   LAYER_FWD  z_out = winv(z) + W_pw * act(z_in) + b
//   LAYER_BWD  gz_in = (winv(gz~) + W_pw^T * gz_out) * act'(z_in);  gW_pw += gz_out x act(z_in); gb += gz_out
// A block owns HT consecutive "lines" (image, h) for all channels and one chunk of w.
// ===========================================================================
struct WinvParams {
  const float2* z; float* y; const float* a; const float* zin;
  const float* pw_w; const float* pw_b; float* g_pw_w; float* g_pw_b;
  const float* t_cos; const float* t_sin;
  int lines, c, hp, wp, wp4, m2, act_in, HT, WCH;
  int table_bulk, bar_off;      // tables by bulk copy (WCH == wp4); float offset of the mbarrier in shared memory
};

template <int MODE, int CG>   // CG = channels accumulated per work item (4, 2 or 1)
__global__ void __launch_bounds__(256) winv_kernel(const WinvParams p) {
  extern __shared__ __align__(16) float smem[];
  const int c = p.c, cpad = (c + 3) & ~3, HT = p.HT, WCH = p.WCH, m2 = p.m2, hp = p.hp, wp = p.wp;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int line0 = blockIdx.x * HT, wc0 = blockIdx.y * WCH;
  const int npx = HT * WCH;

  float* tc = smem;                                        // [m2][WCH]
  float* tsn = tc + m2 * WCH;                              // [m2][WCH]
  float2* zs = reinterpret_cast<float2*>(tsn + m2 * WCH);  // [cpad][HT][m2]
  float* as = reinterpret_cast<float*>(zs + cpad * HT * m2);  // [c][HT][WCH]       (MODE >= 1)
  float* pws = as + (MODE >= 1 ? c * npx : 0);             // [cpad][c] + bias[cpad]
  float* zact = pws + (MODE >= 1 ? cpad * c + cpad : 0);   // [c][HT][WCH] act(z_in)   (MODE == 2)
  float* gsm = zact + (MODE == 2 ? c * npx : 0);           // [c][HT][WCH] act'(z_in)  (MODE == 2)
  uint64_t* tbar = reinterpret_cast<uint64_t*>(smem + p.bar_off);   // completion of the table copies (bulk path)

  // Staging.  Every index split below is a multiply by a precomputed reciprocal (exact for these
  // ranges): ncu showed the integer divisions of the element-wise staging loops to be 45 % of this
  // kernel's instructions.  Activations move as float4.
  const int nwq = WCH >> 2;
  const float inv_nwq = 1.0f / (float)nwq, inv_ht = 1.0f / (float)HT, inv_hp = 1.0f / (float)hp,
              inv_m2 = 1.0f / (float)m2;
  auto fdiv = [](int n, float inv) { return __float2int_rz(((float)n + 0.5f) * inv); };
  pdl_launch_dependents();
  // constant plan data, staged before the dependency wait: when the block covers whole table rows (one w chunk) the
  // two tables are contiguous -> two bulk async copies that land while the spectrum and the activations are staged
  const bool tbulk = p.table_bulk != 0;
  if (tbulk && tid == 0) {
    mbar_init(tbar, 1);
    mbar_init_fence();
    const uint32_t tb = (uint32_t)(m2 * WCH) * 4u;
    mbar_expect_tx(tbar, 2u * tb);
    bulk_g2s(tc, p.t_cos, tb, tbar);
    bulk_g2s(tsn, p.t_sin, tb, tbar);
  }
  for (int idx = tid; idx < (tbulk ? 0 : m2 * nwq); idx += nt) {
    const int l = fdiv(idx, inv_nwq), q = idx - l * nwq;
    const int w = wc0 + 4 * q;     // wp4 is a multiple of 4: a float4 is inside the table row or fully outside
    float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), sv = cv;
    if (w < p.wp4) {
      cv = __ldg(reinterpret_cast<const float4*>(p.t_cos + (size_t)l * p.wp4 + w));
      sv = __ldg(reinterpret_cast<const float4*>(p.t_sin + (size_t)l * p.wp4 + w));
    }
    reinterpret_cast<float4*>(tc)[idx] = cv;
    reinterpret_cast<float4*>(tsn)[idx] = sv;
  }
  pdl_wait();
  for (int idx = tid; idx < cpad * HT * m2; idx += nt) {
    const int row = fdiv(idx, inv_m2), l = idx - row * m2;
    const int ch = fdiv(row, inv_ht), hh = row - ch * HT;
    const int line = line0 + hh;
    float2 v = make_float2(0.f, 0.f);
    if (ch < c && line < p.lines) {
      const int b = fdiv(line, inv_hp), h = line - b * hp;
      v = __ldg(p.z + ((size_t)(b * c + ch) * hp + h) * m2 + l);
    }
    zs[idx] = v;
  }
  if (MODE >= 1) {
    const bool vec = (wp & 3) == 0;
    for (int idx = tid; idx < c * HT * nwq; idx += nt) {
      const int row = fdiv(idx, inv_nwq), q = idx - row * nwq;
      const int ch = fdiv(row, inv_ht), hh = row - ch * HT;
      const int line = line0 + hh, wg = wc0 + 4 * q;
      float v[4] = {0.f, 0.f, 0.f, 0.f}, za[4] = {0.f, 0.f, 0.f, 0.f}, gr[4] = {0.f, 0.f, 0.f, 0.f};
      if (line < p.lines && wg < wp) {
        const int b = fdiv(line, inv_hp), h = line - b * hp;
        const size_t off = ((size_t)(b * c + ch) * hp + h) * wp + wg;
        if (vec) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.a + off));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          if (MODE == 2) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(p.zin + off));
            za[0] = u.x; za[1] = u.y; za[2] = u.z; za[3] = u.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (wg + j < wp) {
              v[j] = __ldg(p.a + off + j);
              if (MODE == 2) za[j] = __ldg(p.zin + off + j);
            }
        }
        if (MODE == 1 && p.act_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
        }
        if (MODE == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (p.act_in) {
              float cdf, pdf;
              gelu_cdf_pdf(za[j], cdf, pdf);
              gr[j] = fmaf(za[j], pdf, cdf);
              za[j] *= cdf;
            } else {
              gr[j] = 1.0f;
            }
          }
        }
      }
      reinterpret_cast<float4*>(as)[idx] = make_float4(v[0], v[1], v[2], v[3]);
      if (MODE == 2) {
        reinterpret_cast<float4*>(zact)[idx] = make_float4(za[0], za[1], za[2], za[3]);
        reinterpret_cast<float4*>(gsm)[idx] = make_float4(gr[0], gr[1], gr[2], gr[3]);
      }
    }
    for (int idx = tid; idx < cpad * c; idx += nt) {
      const int r = idx / c, q = idx - r * c;   // r: channel this pass produces, q: channel it consumes
      float v = 0.f;
      if (r < c) v = MODE == 1 ? __ldg(p.pw_w + r * c + q) : __ldg(p.pw_w + q * c + r);
      pws[idx] = v;
    }
    for (int idx = tid; idx < cpad; idx += nt)
      pws[cpad * c + idx] = (MODE == 1 && idx < c) ? __ldg(p.pw_b + idx) : 0.f;
  }
  __syncthreads();          // (also publishes the mbarrier init)
  pdl_trigger_late();
  if (tbulk) mbar_wait(tbar, 0);

  const int nwg = WCH >> 2, nog = cpad / CG;
  for (int idx = tid; idx < nog * HT * nwg; idx += nt) {
    const int irow = fdiv(idx, inv_nwq), wg = idx - irow * nwg;
    const int og = fdiv(irow, inv_ht), hh = irow - og * HT;
    float acc[CG][4];
#pragma unroll
    for (int r = 0; r < CG; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    const float2* zrow = zs + ((og * CG) * HT + hh) * m2;
    const int zpitch = HT * m2;
#pragma unroll 2
    for (int l = 0; l < m2; ++l) {
      const float4 cs = *reinterpret_cast<const float4*>(tc + l * WCH + wg * 4);
      const float4 sn = *reinterpret_cast<const float4*>(tsn + l * WCH + wg * 4);
#pragma unroll
      for (int r = 0; r < CG; ++r) {
        const float2 zv = zrow[r * zpitch + l];
        acc[r][0] = fmaf(zv.x, cs.x, fmaf(-zv.y, sn.x, acc[r][0]));
        acc[r][1] = fmaf(zv.x, cs.y, fmaf(-zv.y, sn.y, acc[r][1]));
        acc[r][2] = fmaf(zv.x, cs.z, fmaf(-zv.y, sn.z, acc[r][2]));
        acc[r][3] = fmaf(zv.x, cs.w, fmaf(-zv.y, sn.w, acc[r][3]));
      }
    }
    if (MODE >= 1) {
      for (int q = 0; q < c; ++q) {
        const float4 av = *reinterpret_cast<const float4*>(as + (q * HT + hh) * WCH + wg * 4);
#pragma unroll
        for (int r = 0; r < CG; ++r) {
          const float wv = pws[(og * CG + r) * c + q];
          acc[r][0] = fmaf(wv, av.x, acc[r][0]);
          acc[r][1] = fmaf(wv, av.y, acc[r][1]);
          acc[r][2] = fmaf(wv, av.z, acc[r][2]);
          acc[r][3] = fmaf(wv, av.w, acc[r][3]);
        }
      }
    }
    const int line = line0 + hh;
    if (line >= p.lines) continue;
    const int b = fdiv(line, inv_hp), h = line - b * hp;
    const int w0 = wc0 + wg * 4;
#pragma unroll
    for (int r = 0; r < CG; ++r) {
      const int ch = og * CG + r;
      if (ch >= c) continue;
      const size_t off = ((size_t)(b * c + ch) * hp + h) * wp + w0;
      float v[4] = {acc[r][0], acc[r][1], acc[r][2], acc[r][3]};
      if (MODE == 1) {
        const float bias = pws[cpad * c + ch];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += bias;
      }
      if (MODE == 2) {
        const float4 gq = *reinterpret_cast<const float4*>(gsm + (ch * HT + hh) * WCH + wg * 4);
        v[0] *= gq.x; v[1] *= gq.y; v[2] *= gq.z; v[3] *= gq.w;
      }
      if ((wp & 3) == 0 && w0 + 3 < wp) {
        *reinterpret_cast<float4*>(p.y + off) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (w0 + j < wp) p.y[off + j] = v[j];
      }
    }
  }

  if (MODE == 2) {
    // 1x1-conv weight / bias gradients over this tile.  Thread = (pair (o, i), pixel slice): 128-bit
    // loads along the pixels (row-strided, conflict-free for the padded widths used), partial sums
    // joined in shared memory, then one flush per block with 128-bit atomics (4x fewer L2 atomic
    // operations on these few, heavily contended addresses).
    const int npair = c * c + c, npair4 = (npair + 3) & ~3;
    float* part = gsm + c * npx;     // [c*c + c], padded to a multiple of 4
    for (int q = tid; q < npair4; q += nt) part[q] = 0.f;
    __syncthreads();
    const int nsl = nt / npair > 0 ? nt / npair : 1;           // pixel slices (threads per pair)
    const int nq = npx >> 2;                                   // float4 chunks per channel plane (WCH % 4 == 0)
    for (int item = tid; item < npair * nsl; item += nt) {
      const int pair = item % npair, sl = item / npair;
      const int q0 = (int)((long)sl * nq / nsl), q1 = (int)((long)(sl + 1) * nq / nsl);
      float s = 0.f;
      if (pair < c * c) {
        const int o = pair / c, i = pair - o * c;
        const float4* go = reinterpret_cast<const float4*>(as + o * npx);
        const float4* ai = reinterpret_cast<const float4*>(zact + i * npx);
        for (int q = q0; q < q1; ++q) {
          const float4 g4 = go[q], a4 = ai[q];
          s = fmaf(g4.x, a4.x, fmaf(g4.y, a4.y, fmaf(g4.z, a4.z, fmaf(g4.w, a4.w, s))));
        }
      } else {
        const float4* go = reinterpret_cast<const float4*>(as + (pair - c * c) * npx);
        for (int q = q0; q < q1; ++q) {
          const float4 g4 = go[q];
          s += (g4.x + g4.y) + (g4.z + g4.w);
        }
      }
      if (nsl > 1) atomicAdd(part + pair, s); else part[pair] = s;
    }
    __syncthreads();
    const bool vec_ok = ((c * c) & 3) == 0 && (c & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.g_pw_w) | reinterpret_cast<uintptr_t>(p.g_pw_b)) & 15) == 0;
    if (vec_ok) {
      for (int q = tid; q < npair >> 2; q += nt) {
        const float4 v = reinterpret_cast<const float4*>(part)[q];
        float* dst = 4 * q < c * c ? p.g_pw_w + 4 * q : p.g_pw_b + (4 * q - c * c);
        atomicAdd(reinterpret_cast<float4*>(dst), v);
      }
    } else {
      for (int q = tid; q < npair; q += nt) atomicAdd(q < c * c ? p.g_pw_w + q : p.g_pw_b + (q - c * c), part[q]);
    }
  }
}

template <int MODE, int CG>
static void launch_winv_t(const WinvParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(winv_kernel<MODE, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  launch_k(winv_kernel<MODE, CG>, dim3(grid), dim3(256), smem, st, p);
}

void launch_winv(const Plan* pl, int mode, const WinvArgs& a, cudaStream_t st) {
  LaunchScope scope(mode == WINV_PLAIN ? "winv_plain" : (mode == WINV_LAYER_FWD ? "winv_layer_fwd" : "winv_layer_bwd"), st, a.c);
  WinvParams p;
  p.z = a.z; p.y = a.y; p.a = a.a; p.zin = a.zin; p.pw_w = a.pw_w; p.pw_b = a.pw_b;
  p.g_pw_w = a.g_pw_w; p.g_pw_b = a.g_pw_b;
  p.t_cos = pl->t_lw_cos; p.t_sin = pl->t_lw_sin;
  p.lines = a.images * pl->hp; p.c = a.c; p.hp = pl->hp; p.wp = pl->wp; p.wp4 = pl->wp4; p.m2 = pl->m2;
  p.act_in = a.act_in;
  const int c = a.c, cpad = (c + 3) & ~3;
  auto smem_of = [&](int ht, int wch) {
    size_t f = 2 * (size_t)pl->m2 * wch + 2 * (size_t)cpad * ht * pl->m2;
    if (mode >= 1) f += (size_t)c * ht * wch + cpad * c + cpad;
    if (mode == 2) f += 2 * (size_t)c * ht * wch + ((c * c + c + 3) & ~3);
    return f * sizeof(float);
  };
  // (lines per block, channels per item): a work item is CG channels x 4 pixels of one line.  Pick the
  // pair that keeps the 256 threads busiest, counting the FMA density of the inner loop
  // (8*CG FMAs per 2 + CG shared loads), among tilings that leave >= 2 blocks per SM when possible.
  int wch = pl->wp4, nch = 1;
  while (smem_of(1, wch) > 64 * 1024 && wch > 4) {
    ++nch;
    wch = (ceil_div(pl->wp4, nch) + 3) & ~3;
  }
  const int nwg = wch >> 2;
  int best_ht = 1, best_cg = 1;
  double best = -1.0;
  const long want_blocks = 2 * 148;
  for (int pass = 0; pass < 2 && best < 0.0; ++pass)
    for (int cg = 4; cg >= 1; cg >>= 1)
      for (int ht = 1; ht <= 16; ++ht) {
        if (smem_of(ht, wch) > 64 * 1024) break;
        const long blocks = (long)ceil_div(p.lines, ht) * nch;
        if (pass == 0 && blocks < want_blocks) break;
        const int items = (cpad / cg) * ht * nwg;
        const double util = (double)items / (ceil_div(items, 256) * 256.0);
        const double dens = 8.0 * cg / (8.0 * cg + 2.0 + cg);
        const double score = util * dens * (pass == 1 && ht > 1 ? 0.0 : 1.0);
        if (score > best) { best = score; best_ht = ht; best_cg = cg; }
      }
  {   // (tuning knobs for the many-image regime: "ht,cg" applied when lines >= 4096)
    static const char* knob = getenv("BDN_WINV_TILE");
    int kh = 0, kc = 0;
    if (knob && p.lines >= 4096 && sscanf(knob, "%d,%d", &kh, &kc) == 2 && kh >= 1 && (kc == 1 || kc == 2 || kc == 4) &&
        smem_of(kh, wch) <= 64 * 1024) {
      best_ht = kh; best_cg = kc;
    }
  }
  p.HT = best_ht; p.WCH = wch;
  size_t smem = smem_of(best_ht, wch);
  smem = (smem + 15) & ~(size_t)15;
  p.bar_off = (int)(smem / sizeof(float));
  p.table_bulk = nch == 1 && wch == pl->wp4;
  smem += 16;
  dim3 grid(ceil_div(p.lines, best_ht), ceil_div(pl->wp4, wch));
#define BDN_WINV(M)                                                        \
  {                                                                        \
    if (best_cg == 4) launch_winv_t<M, 4>(p, grid, smem, st);              \
    else if (best_cg == 2) launch_winv_t<M, 2>(p, grid, smem, st);         \
    else launch_winv_t<M, 1>(p, grid, smem, st);                           \
  }
  if (mode == WINV_PLAIN) BDN_WINV(0) else if (mode == WINV_LAYER_FWD) BDN_WINV(1) else BDN_WINV(2)
#undef BDN_WINV
}

}  // namespace bdn
