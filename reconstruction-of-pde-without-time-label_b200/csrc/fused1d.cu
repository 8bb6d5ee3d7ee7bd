// Fused 1-D FNO layer (SpectralConv1d + 1x1 conv + bias, GELU on load; 1d_FPE/FNOModules.py:47-59, :108-114).
//
// A 1-D image is tiny (C x Np floats: 12 KB at C = 30, Np = 100), so the three-kernel layer of the 2-D path
// (W-forward DFT -> per-mode mix -> inverse DFT + epilogue) is pure launch latency here.  One kernel per layer
// and direction keeps the whole image group in shared memory:
//
//   forward : a = act(z_in) -> X = pre * DFT_m(a) (saved for backward) -> Y = post * mix(X, W)
//             -> z_out = iDFT(Y) + W_pw a + b
//   backward: G = pre' * DFT_m(g) (written out: the spectral weight gradient is reduced over images by
//             gw_reduce) -> GZ = post' * mix(G, conj W^T) -> gz_in = (iDFT(GZ) + W_pw^T g) * act'(z_in),
//             1x1-conv weight / bias gradients reduced per block and flushed with atomics
//
// Same arithmetic, tables and column scales as wfwd / mix1d / winv (spectral.cu); a block works on groups of IB
// images so that small-width nets (the per-snapshot FNO_input, C = 4) still fill its 256 threads.
#include "bdn_internal.cuh"

namespace bdn {

struct Layer1dParams {
  const float* z_in;      // [images, c, wp] pre-activation input
  const float* g_out;     // bwd: gradient of the layer output [images, c, wp]
  float* out;             // fwd: z_out; bwd: gz_in
  float2* spec;           // fwd: xs_saved (may be null); bwd: gys [images, c, m2] (for gw_reduce)
  const float2* w;        // [c, c, m2] complex spectral weights
  const float* pw_w; const float* pw_b;     // 1x1 conv [c, c], [c]
  float* g_pw_w; float* g_pw_b;             // bwd accumulators (+=)
  const float* t_cos; const float* t_sin;   // [m2][wp4]
  const float* pre; const float* post;      // [m2] column scales (fwd: col_dc, col_fwd; bwd: col_fwd, col_dc)
  int images, c, wp, wp4, tp, m2, act_in, ib;
  int split;              // > 1 (only with ib == 1): SPLIT blocks share an image, each owning a slice of the OUTPUT channels
                          // in the mix / inverse / weight-gradient phases (staging and the forward DFT are repeated)
};

__device__ __forceinline__ int fdiv_i(int n, float inv) { return __float2int_rz(((float)n + 0.5f) * inv); }

template <bool BWD>
__global__ void __launch_bounds__(256) layer1d_kernel(const Layer1dParams p) {
  extern __shared__ __align__(16) float smem[];
  const int c = p.c, wp = p.wp, wp4 = p.wp4, tp = p.tp, m2 = p.m2, IB = p.ib;
  const int rows = IB * c, nq = wp4 >> 2, tq = tp >> 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  // layout (floats): tables 2*m2*tp | a rows*tp | (bwd: g rows*tp, gp rows*tp) | X rows*m2*2 | Y rows*m2*2 | pw c*c | pb c
  //                  | pre m2 | post m2 | (bwd: acc c*c + c)
  float* tc = smem;
  float* ts = tc + m2 * tp;
  float* a = ts + m2 * tp;
  float* g = a + rows * tp;
  float* gp = g + (BWD ? rows * tp : 0);
  float2* X = reinterpret_cast<float2*>(gp + (BWD ? rows * tp : 0));
  float2* Y = X + rows * m2;
  float* pw = reinterpret_cast<float*>(Y + rows * m2);
  float* pb = pw + c * c;
  float* pre = pb + ((c + 3) & ~3);
  float* post = pre + ((m2 + 3) & ~3);
  float* acc = post + ((m2 + 3) & ~3);          // [c*c + c] (BWD)

  const float inv_tp = 1.0f / (float)tp, inv_m2 = 1.0f / (float)m2, inv_c = 1.0f / (float)c, inv_nq = 1.0f / (float)nq;
  pdl_launch_dependents();
  for (int i = tid; i < m2 * tp; i += nt) {        // constant plan data
    const int l = fdiv_i(i, inv_tp), w = i - l * tp;
    tc[i] = w < wp4 ? __ldg(p.t_cos + (size_t)l * wp4 + w) : 0.f;
    ts[i] = w < wp4 ? __ldg(p.t_sin + (size_t)l * wp4 + w) : 0.f;
  }
  for (int i = tid; i < m2; i += nt) { pre[i] = __ldg(p.pre + i); post[i] = __ldg(p.post + i); }
  pdl_wait();
  for (int i = tid; i < c * c; i += nt) pw[i] = __ldg(p.pw_w + i);
  if (!BWD) for (int i = tid; i < c; i += nt) pb[i] = __ldg(p.pw_b + i);
  if (BWD) for (int i = tid; i < c * c + c; i += nt) acc[i] = 0.f;
  __syncthreads();

  const int ngroups = (p.images + IB - 1) / IB;
  const int part = blockIdx.x % p.split, nblk = gridDim.x / p.split;
  // output rows this block owns in phases C-E (all rows unless the image is split over several blocks)
  const int o_lo = p.split > 1 ? part * c / p.split : 0, o_hi = p.split > 1 ? (part + 1) * c / p.split : rows;
  const int orows = o_hi - o_lo;
  for (int grp = blockIdx.x / p.split; grp < ngroups; grp += nblk) {
    const int b0 = grp * IB;
    const int live = min(IB, p.images - b0) * c;       // rows of this group that exist
    // ---- A: stage rows (row r = (image b0 + r / c, channel r % c)); GELU / GELU' on load
    for (int i = tid; i < rows * tp; i += nt) {
      const int r = fdiv_i(i, inv_tp), w = i - r * tp;
      float zv = 0.f, gv = 0.f;
      const bool in = r < live && w < wp;
      if (in) {
        zv = __ldg(p.z_in + ((size_t)b0 * c + r) * wp + w);
        if (BWD) gv = __ldg(p.g_out + ((size_t)b0 * c + r) * wp + w);
      }
      if (BWD) {
        float d = in ? 1.0f : 0.f;
        if (p.act_in) {
          float cdf, pdf;
          gelu_cdf_pdf(zv, cdf, pdf);
          d = in ? fmaf(zv, pdf, cdf) : 0.f;
          zv *= cdf;
        }
        g[i] = gv; gp[i] = d;
      } else if (p.act_in) {
        zv = gelu_fast(zv);
      }
      a[i] = zv;
    }
    __syncthreads();
    // ---- B: pruned forward DFT of the rows (fwd: of a; bwd: of g), times the pre scale
    const float* src = BWD ? g : a;
    for (int i = tid; i < rows * m2; i += nt) {
      const int r = fdiv_i(i, inv_m2), l = i - r * m2;
      const float4* xr = reinterpret_cast<const float4*>(src + r * tp);
      const float4* cr = reinterpret_cast<const float4*>(tc + l * tp);
      const float4* sr = reinterpret_cast<const float4*>(ts + l * tp);
      float re0 = 0.f, re1 = 0.f, im0 = 0.f, im1 = 0.f;
#pragma unroll 5
      for (int q = 0; q < nq; ++q) {
        const float4 x = xr[q], c4 = cr[q], s4 = sr[q];
        re0 = fmaf(x.x, c4.x, fmaf(x.y, c4.y, re0));
        re1 = fmaf(x.z, c4.z, fmaf(x.w, c4.w, re1));
        im0 = fmaf(x.x, s4.x, fmaf(x.y, s4.y, im0));
        im1 = fmaf(x.z, s4.z, fmaf(x.w, s4.w, im1));
      }
      const float sc = pre[l];
      const float2 v = make_float2((re0 + re1) * sc, -(im0 + im1) * sc);
      X[i] = v;
      if (p.spec != nullptr && r < live && part == 0) p.spec[((size_t)b0 * c + r) * m2 + l] = v;
    }
    __syncthreads();
    // ---- C: per-mode channel mix.  fwd: y_o = sum_a x_a W[a][o]; bwd: y_i = sum_o x_o conj(W[i][o])
    for (int i = tid; i < orows * m2; i += nt) {
      const int r = o_lo + fdiv_i(i, inv_m2), l = i - (r - o_lo) * m2;
      const int ib = fdiv_i(r, inv_c), o = r - ib * c;
      const float2* xin = X + ib * c * m2 + l;
      float yr = 0.f, yi = 0.f;
#pragma unroll 6
      for (int k = 0; k < c; ++k) {
        const float2 x = xin[k * m2];
        if (!BWD) {
          const float2 wv = __ldg(p.w + (size_t)(k * c + o) * m2 + l);
          yr = fmaf(x.x, wv.x, fmaf(-x.y, wv.y, yr));
          yi = fmaf(x.x, wv.y, fmaf(x.y, wv.x, yi));
        } else {
          const float2 wv = __ldg(p.w + (size_t)(o * c + k) * m2 + l);
          yr = fmaf(x.x, wv.x, fmaf(x.y, wv.y, yr));
          yi = fmaf(x.y, wv.x, fmaf(-x.x, wv.y, yi));
        }
      }
      const float sc = post[l];
      Y[r * m2 + l] = make_float2(yr * sc, yi * sc);
    }
    __syncthreads();
    // ---- D: inverse DFT + 1x1 conv (+ bias | * GELU'), 4 pixels per item
    for (int i = tid; i < orows * nq; i += nt) {
      const int r = o_lo + fdiv_i(i, inv_nq), q = i - (r - o_lo) * nq;
      if (r >= live) continue;
      const int ib = fdiv_i(r, inv_c), o = r - ib * c;
      float v0, v1, v2, v3;
      v0 = v1 = v2 = v3 = BWD ? 0.f : pb[o];
      const float2* yrow = Y + r * m2;
      const float4* cq = reinterpret_cast<const float4*>(tc) + q;
      const float4* sq = reinterpret_cast<const float4*>(ts) + q;
#pragma unroll 4
      for (int l = 0; l < m2; ++l) {
        const float2 y = yrow[l];
        const float4 c4 = cq[l * tq], s4 = sq[l * tq];
        v0 = fmaf(y.x, c4.x, fmaf(-y.y, s4.x, v0));
        v1 = fmaf(y.x, c4.y, fmaf(-y.y, s4.y, v1));
        v2 = fmaf(y.x, c4.z, fmaf(-y.y, s4.z, v2));
        v3 = fmaf(y.x, c4.w, fmaf(-y.y, s4.w, v3));
      }
      const float4* arow = reinterpret_cast<const float4*>((BWD ? g : a) + ib * c * tp) + q;
#pragma unroll 6
      for (int k = 0; k < c; ++k) {
        const float4 av = arow[k * tq];
        const float wv = BWD ? pw[k * c + o] : pw[o * c + k];
        v0 = fmaf(wv, av.x, v0); v1 = fmaf(wv, av.y, v1); v2 = fmaf(wv, av.z, v2); v3 = fmaf(wv, av.w, v3);
      }
      if (BWD) {
        const float4 d = *(reinterpret_cast<const float4*>(gp + r * tp) + q);
        v0 *= d.x; v1 *= d.y; v2 *= d.z; v3 *= d.w;
      }
      float* dst = p.out + ((size_t)b0 * c + r) * wp + 4 * q;
      const float v[4] = {v0, v1, v2, v3};
      if ((wp & 3) == 0 && 4 * q + 3 < wp) {
        *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * q + j < wp) dst[j] = v[j];
      }
    }
    if (BWD) {
      // ---- E: 1x1-conv weight / bias gradient partials of this group: pair (o, i) x pixel slices
      // (a split image: this block owns the pairs whose first index o lies in [o_lo, o_hi); rows == c there)
      const int oc = p.split > 1 ? orows : c, ob = p.split > 1 ? o_lo : 0;
      const int npair = oc * c + oc;
      const int nsl = nt / npair > 0 ? nt / npair : 1;
      for (int item = tid; item < npair * nsl; item += nt) {
        const int lp = item % npair, sl = item / npair;
        const int pair = lp < oc * c ? (ob + lp / c) * c + lp % c : c * c + ob + (lp - oc * c);
        float s = 0.f;
        for (int ib = 0; ib * c < live; ++ib) {
          const int q0 = (int)((long)sl * nq / nsl), q1 = (int)((long)(sl + 1) * nq / nsl);
          if (pair < c * c) {
            const int o = pair / c, k = pair - o * c;
            const float4* go = reinterpret_cast<const float4*>(g + (ib * c + o) * tp);
            const float4* ak = reinterpret_cast<const float4*>(a + (ib * c + k) * tp);
            for (int q = q0; q < q1; ++q) {
              const float4 g4 = go[q], a4 = ak[q];
              s = fmaf(g4.x, a4.x, fmaf(g4.y, a4.y, fmaf(g4.z, a4.z, fmaf(g4.w, a4.w, s))));
            }
          } else {
            const float4* go = reinterpret_cast<const float4*>(g + (ib * c + pair - c * c) * tp);
            for (int q = q0; q < q1; ++q) {
              const float4 g4 = go[q];
              s += (g4.x + g4.y) + (g4.z + g4.w);
            }
          }
        }
        atomicAdd(acc + pair, s);
      }
    }
    __syncthreads();       // the next group overwrites a / g / X / Y
  }
  if (BWD) {
    for (int i = tid; i < c * c + c; i += nt) atomicAdd(i < c * c ? p.g_pw_w + i : p.g_pw_b + (i - c * c), acc[i]);
  }
}

static size_t layer1d_smem(int c, int tp, int m2, int ib, bool bwd) {
  const int rows = ib * c;
  size_t f = 2 * (size_t)m2 * tp + (size_t)rows * tp * (bwd ? 3 : 1) + 4 * (size_t)rows * m2 + (size_t)c * c + ((c + 3) & ~3) +
             2 * ((m2 + 3) & ~3);
  if (bwd) f += c * c + c;
  return f * sizeof(float);
}

static bool fused1d_enabled() {
  static const bool on = [] {
    const char* e = getenv("BDN_FUSED1D");      // BDN_FUSED1D=0: the three-kernel layer (A/B measurements)
    return !(e && e[0] == '0');
  }();
  return on;
}

// One whole 1-D layer in one launch.  Returns false (nothing launched) when the shape does not fit shared memory.
bool launch_layer1d(const Plan* pl, bool bwd, const float* z_in, const float* g_out, float* out, float2* spec,
                    const float2* w, const float* pw_w, const float* pw_b, float* g_pw_w, float* g_pw_b, int images, int c,
                    int act_in, cudaStream_t st) {
  if (!fused1d_enabled() || pl->ndim != 1 || images <= 0) return false;
  const int m2 = pl->m2, wp4 = pl->wp4;
  const int tp = ((wp4 >> 2) & 1) ? wp4 : wp4 + 4;      // odd number of float4 per table row: conflict-free across modes
  int ib = 256 / (c * m2);                              // fill the 256 threads in the DFT / mix phases
  ib = ib < 1 ? 1 : (ib > 16 ? 16 : ib);
  if (ib > images) ib = images;
  while (ib > 1 && layer1d_smem(c, tp, m2, ib, bwd) > 96 * 1024) --ib;
  const size_t smem = layer1d_smem(c, tp, m2, ib, bwd);
  if (smem > 200 * 1024) return false;
  LaunchScope scope(bwd ? "layer1d_bwd" : "layer1d_fwd", st, c);
  Layer1dParams p;
  p.z_in = z_in; p.g_out = g_out; p.out = out; p.spec = spec; p.w = w; p.pw_w = pw_w; p.pw_b = pw_b;
  p.g_pw_w = g_pw_w; p.g_pw_b = g_pw_b; p.t_cos = pl->t_lw_cos; p.t_sin = pl->t_lw_sin;
  p.pre = bwd ? pl->col_fwd : pl->col_dc;
  p.post = bwd ? pl->col_dc : pl->col_fwd;
  p.images = images; p.c = c; p.wp = pl->wp; p.wp4 = wp4; p.tp = tp; p.m2 = m2; p.act_in = act_in; p.ib = ib;
  const int ngroups = ceil_div(images, ib);
  const int per_sm = (int)((220 * 1024) / (smem + 1024)) < 8 ? (int)((220 * 1024) / (smem + 1024)) : 8;
  const int cap = 148 * (per_sm < 1 ? 1 : per_sm);
  // few wide images (the heads: 32 images x 30 channels): several blocks per image so that the machine is not 3/4 idle
  int split = 1;
  if (ib == 1 && c >= 8) {
    split = 148 / ngroups;
    split = split < 1 ? 1 : (split > 4 ? 4 : split);
    if (split > c / 4) split = c / 4 > 0 ? c / 4 : 1;
  }
  p.split = split;
  int gblocks = cap / split > 0 ? cap / split : 1;
  if (gblocks > ngroups) gblocks = ngroups;
  const int grid = gblocks * split;
  if (bwd) {
    cudaFuncSetAttribute(layer1d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_k(layer1d_kernel<true>, dim3(grid), dim3(256), smem, st, p);
  } else {
    cudaFuncSetAttribute(layer1d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_k(layer1d_kernel<false>, dim3(grid), dim3(256), smem, st, p);
  }
  return true;
}

}  // namespace bdn
