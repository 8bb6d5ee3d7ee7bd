// Bag attention + bag mean of the BlinDNO models in a handful of small launches per U-Net level, without ever forming
// an [L, D] intermediate:
//
//   TemporalSelfAttention.forward  2d_FPE/NIOModules.py:1063-1083   (tokens = flattened feature maps, no projections)
//       S = X X^T / sqrt(D),  A = softmax(S),  O = A X + X,  Y = LayerNorm_D(O)
//   followed by  .mean(dim=1)      2d_FPE/NIOModules.py:1153-1170   (PermInvUNet_attn.forward: h_att.mean, skip_att.mean)
//
// X is [L, D] per bag (L <= 128 snapshots, D = C*H*W up to a few thousand).  Everything between the two passes over X
// lives in L x L matrices: with M = A + I, m_l = mean_d X_ld and the CENTERED Gram matrix Gc = (X - m 1^T)(X - m 1^T)^T
//   S        = (Gc + D m m^T) / sqrt(D)
//   mu_l     = (M m)_l                       row means of O
//   var_l    = (M Gc M^T)_ll / D             row variances of O (a sum of squares: no cancellation)
//   r_l      = 1 / sqrt(var_l + eps)
//   out_d    = gamma_d / L * (sum_l' v_l' X_l'd - c) + beta_d,     v = M^T r,  c = sum_l r_l mu_l
// so the forward is: row means -> centered Gram (one pass over X) -> L x L algebra (one block per bag) -> weighted
// column sum (one pass over X).  Backward, with h_d = g_d gamma_d / L, hbar = mean_d h, w' = X h:
//   p_l   = r_l / D * ((M w')_l - mu_l sum_d h_d),   a_l = r_l^2 p_l,   w = w' - hbar D m
//   dA    = r w^T - diag(a) (M Gc),   dS = A .* (dA - rowsum(A .* dA))
//   dX    = K X + v (h - hbar)^T + z 1^T,   K = -M^T diag(a) M + (dS + dS^T) / sqrt(D),   z = M^T (a .* mu)
// i.e. one pass for w' (and the LayerNorm parameter gradients), L x L algebra, one pass for dX.
#include "bdn_internal.cuh"

#include <cmath>

namespace bdn {

constexpr int BA_MAXL = 128;
constexpr int BA_DC = 64;        // columns of X per shared-memory chunk

// ---------------------------------------------------------------------------
// row means: m[b, l] = mean_d x[b, l, d]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_rowmean_kernel(const float* __restrict__ x, float* __restrict__ m, int D) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const float* row = x + (size_t)blockIdx.x * D;
  float s = 0.f;
#pragma unroll 8
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += __ldg(row + d);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    m[blockIdx.x] = t / (float)D;
  }
}

// ---------------------------------------------------------------------------
// centered Gram matrix: Gc[b, l, l'] += sum_d (x_ld - m_l)(x_l'd - m_l')   (Gc zeroed by the caller)
// grid (blocks per bag, bags); a block loops over its column chunks, a thread owns 4 x 4 tiles of (l, l')
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_gram_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                      float* __restrict__ gc, int L, int LP, int D) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* xt = smem;                       // [BA_DC][LP]  centered, transposed: 4 consecutive l are one float4
  float* ms = xt + BA_DC * LP;            // [LP]
  const int b = blockIdx.y, tid = threadIdx.x;
  const float* xb = x + (size_t)b * L * D;
  for (int l = tid; l < LP; l += blockDim.x) ms[l] = l < L ? __ldg(m + b * L + l) : 0.f;
  const int nt4 = LP >> 2, ntiles = nt4 * nt4;
  constexpr int MAXT = (BA_MAXL / 4) * (BA_MAXL / 4) / 256;      // tiles per thread at L = 128
  float acc[MAXT][4][4];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  const int nchunk = (D + BA_DC - 1) / BA_DC;
  for (int ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    __syncthreads();                      // ms staged / previous chunk consumed
    const int d0 = ch * BA_DC;
    for (int i = tid; i < LP * BA_DC; i += blockDim.x) {
      const int l = i / BA_DC, k = i - l * BA_DC;
      const int d = d0 + k;
      xt[k * LP + l] = (l < L && d < D) ? __ldg(xb + (size_t)l * D + d) - ms[l] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int tile = tid + t * 256;
      if (tile < ntiles) {
        const int ti = tile / nt4, tj = tile - ti * nt4;
#pragma unroll 4
        for (int k = 0; k < BA_DC; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(xt + k * LP + 4 * ti);
          const float4 c = *reinterpret_cast<const float4*>(xt + k * LP + 4 * tj);
          const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = fmaf(av[i], cv[j], acc[t][i][j]);
        }
      }
    }
  }
  float* g = gc + (size_t)b * L * L;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    const int tile = tid + t * 256;
    if (tile < ntiles) {
      const int ti = tile / nt4, tj = tile - ti * nt4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int l = 4 * ti + i, lp = 4 * tj + j;
          if (l < L && lp < L) atomicAdd(g + l * L + lp, acc[t][i][j]);
        }
    }
  }
}

// ---------------------------------------------------------------------------
// L x L algebra of the forward.  The rows of A, mu and r are independent of each other: grid (row groups, bags), a
// block handles rows blockIdx.x, blockIdx.x + gridDim.x, ... (one warp per row) with the whole centered Gram matrix
// in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_small_fwd_kernel(const float* __restrict__ gc, const float* __restrict__ m,
                                                           float* __restrict__ a_out, float* __restrict__ r_out,
                                                           float* __restrict__ mu_out, int L, int LP, int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* G = smem;                 // [L][LP] centered Gram
  float* ms = G + L * LP;          // [LP]
  float* Mrow = ms + LP;           // [warps][LP]: row l of A, then of A + I
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float* g = gc + (size_t)b * L * L;
  for (int i = tid; i < L * L; i += blockDim.x) {
    const int l = i / L, lp = i - l * L;
    G[l * LP + lp] = __ldg(g + i);
  }
  for (int l = tid; l < LP; l += blockDim.x) ms[l] = l < L ? __ldg(m + b * L + l) : 0.f;
  __syncthreads();
  const float inv_sqrt_d = rsqrtf((float)D), fd = (float)D;
  float* M = Mrow + warp * LP;
  for (int l = blockIdx.x * nw + warp; l < L; l += gridDim.x * nw) {
    // softmax of row l, M = A + I, mu = M m
    float mx = -INFINITY;
    for (int lp = lane; lp < L; lp += 32) {
      const float sv = (G[l * LP + lp] + fd * ms[l] * ms[lp]) * inv_sqrt_d;
      M[lp] = sv;
      mx = fmaxf(mx, sv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int lp = lane; lp < L; lp += 32) {
      const float e = expf(M[lp] - mx);
      M[lp] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float mu = 0.f;
    for (int lp = lane; lp < L; lp += 32) {
      const float av = M[lp] * inv;
      a_out[((size_t)b * L + l) * L + lp] = av;
      const float mm = av + (lp == l ? 1.0f : 0.f);
      M[lp] = mm;
      mu = fmaf(mm, ms[lp], mu);
    }
    mu = warp_sum(mu);
    __syncwarp();
    // var_l = (M Gc M^T)_ll / D: t_j = sum_k M_lk Gc_kj (lane owns columns j), then the dot with M_l.
    float q = 0.f;
    for (int j = lane; j < L; j += 32) {
      float t = 0.f;
      for (int k = 0; k < L; ++k) t = fmaf(M[k], G[k * LP + j], t);
      q = fmaf(t, M[j], q);
    }
    q = warp_sum(q);
    if (lane == 0) {
      r_out[b * L + l] = rsqrtf(fmaxf(q, 0.f) / fd + eps);
      mu_out[b * L + l] = mu;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// out[b, d] = gamma_d / L * (sum_l v_l x_ld - c) + beta_d,  v = M^T r,  c = sum_l r_l mu_l  (every block forms v and c
// from A, r, mu: L^2 multiply-adds; block (0, b) also saves them for the backward)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_out_kernel(const float* __restrict__ x, const float* __restrict__ a_in,
                                                     const float* __restrict__ r, const float* __restrict__ mu,
                                                     float* __restrict__ v_out, float* __restrict__ c_out,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* __restrict__ out, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float vs[BA_MAXL], rs[BA_MAXL], cs;
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int l = tid; l < L; l += blockDim.x) rs[l] = __ldg(r + b * L + l);
  __syncthreads();
  for (int lp = tid; lp < L; lp += blockDim.x) {
    float v = rs[lp];                                  // the + I of M = A + I
    const float* col = a_in + (size_t)b * L * L + lp;
#pragma unroll 8
    for (int l = 0; l < L; ++l) v = fmaf(rs[l], __ldg(col + (size_t)l * L), v);      // (independent loads: keep 8 in flight)
    vs[lp] = v;
    if (blockIdx.x == 0) v_out[b * L + lp] = v;
  }
  if (tid < 32) {
    float c = 0.f;
    for (int l = tid; l < L; l += 32) c = fmaf(rs[l], __ldg(mu + b * L + l), c);
    c = warp_sum(c);
    if (tid == 0) {
      cs = c;
      if (blockIdx.x == 0) c_out[b] = c;
    }
  }
  __syncthreads();
  const int d = blockIdx.x * blockDim.x + tid;
  if (d >= D) return;
  const float* xb = x + (size_t)b * L * D + d;
  float s = 0.f;
#pragma unroll 16
  for (int l = 0; l < L; ++l) s = fmaf(vs[l], __ldg(xb + (size_t)l * D), s);      // (few blocks: keep 16 loads in flight)
  out[(size_t)b * D + d] = fmaf(__ldg(gamma + d) / (float)L, s - cs, __ldg(beta + d));
}

// ---------------------------------------------------------------------------
// backward pass 1: w'[b, l] += sum_d x_ld h_d, hsum[b] += sum_d h_d, dgamma[b, d] = g_d u_d   (w', hsum zeroed by the caller)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         const float* __restrict__ v, const float* __restrict__ c,
                                                         const float* __restrict__ gamma, float* __restrict__ wp,
                                                         float* __restrict__ hsum, float* __restrict__ dgamma, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float vs[BA_MAXL], wacc[BA_MAXL];
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  for (int l = threadIdx.x; l < L; l += blockDim.x) { vs[l] = __ldg(v + b * L + l); wacc[l] = 0.f; }
  __syncthreads();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = d < D;
  const float* xb = x + (size_t)b * L * D + (live ? d : 0);
  const float gd = live ? __ldg(g + (size_t)b * D + d) : 0.f;
  const float h = live ? gd * __ldg(gamma + d) / (float)L : 0.f;
  float s = 0.f;
  // 32 rows at a time: a lane holds its column's 32 products, a butterfly of 31 shuffles leaves the total of row
  // l0 + lane on every lane (a warp_sum per row would be 160 shuffles)
  for (int l0 = 0; l0 < L; l0 += 32) {
    float pr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int l = l0 + i;
      const float xv = (live && l < L) ? __ldg(xb + (size_t)l * D) : 0.f;
      s = fmaf(l < L ? vs[l] : 0.f, xv, s);
      pr[i] = xv * h;
    }
    int n = 32;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
      n >>= 1;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < n) {
          const float send = up ? pr[i] : pr[i + n];
          const float keep = up ? pr[i + n] : pr[i];
          pr[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
    }
    if (l0 + lane < L) atomicAdd(&wacc[l0 + lane], pr[0]);
  }
  if (live) dgamma[(size_t)b * D + d] = gd * (s - __ldg(c + b)) / (float)L;
  const float hs = warp_sum(h);
  if (lane == 0) atomicAdd(hsum + b, hs);
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) atomicAdd(wp + b * L + l, wacc[l]);
}

// ---------------------------------------------------------------------------
// L x L algebra of the backward, part 1 (rows independent: grid (row groups, bags), one warp per row):
//   a_l = r_l^3 / D ((M w')_l - mu_l hsum),  T_l. = (M Gc)_l.,  dA_l. = r_l w^T - a_l T_l.,  dS_l. = A_l. .* (dA_l. - <A_l., dA_l.>)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_small_bwd_rows_kernel(const float* __restrict__ gc, const float* __restrict__ a_in,
                                                                const float* __restrict__ m, const float* __restrict__ r,
                                                                const float* __restrict__ mu, const float* __restrict__ wp,
                                                                const float* __restrict__ hsum, float* __restrict__ ds_out,
                                                                float* __restrict__ as_out, int L, int LP, int D) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* G = smem;                 // [L][LP] centered Gram
  float* wps = G + L * LP;         // [LP] w'
  float* ws = wps + LP;            // [LP] w = w' - hbar D m
  float* Mrow = ws + LP;           // [warps][LP]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float fd = (float)D;
  for (int i = tid; i < L * L; i += blockDim.x) {
    const int l = i / L, lp = i - l * L;
    G[l * LP + lp] = __ldg(gc + (size_t)b * L * L + i);
  }
  const float hs = __ldg(hsum + b), hbar = hs / fd;
  for (int l = tid; l < L; l += blockDim.x) {
    const float w = __ldg(wp + b * L + l);
    wps[l] = w;
    ws[l] = w - hbar * fd * __ldg(m + b * L + l);
  }
  __syncthreads();
  float* M = Mrow + warp * LP;
  for (int l = blockIdx.x * nw + warp; l < L; l += gridDim.x * nw) {
    float s = 0.f;
    for (int k = lane; k < L; k += 32) {
      const float mm = __ldg(a_in + ((size_t)b * L + l) * L + k) + (k == l ? 1.0f : 0.f);
      M[k] = mm;
      s = fmaf(mm, wps[k], s);
    }
    s = warp_sum(s);
    const float rl = __ldg(r + b * L + l);
    const float al = rl * rl * rl / fd * (s - __ldg(mu + b * L + l) * hs);
    if (lane == 0) as_out[b * L + l] = al;
    __syncwarp();
    float dav[BA_MAXL / 32];
    float dot = 0.f;
#pragma unroll
    for (int jj = 0; jj < BA_MAXL / 32; ++jj) {
      const int j = lane + 32 * jj;
      dav[jj] = 0.f;
      if (j < L) {
        float t = 0.f;
        for (int k = 0; k < L; ++k) t = fmaf(M[k], G[k * LP + j], t);
        const float da = rl * ws[j] - al * t;
        dav[jj] = da;
        dot = fmaf(M[j] - (j == l ? 1.0f : 0.f), da, dot);
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int jj = 0; jj < BA_MAXL / 32; ++jj) {
      const int j = lane + 32 * jj;
      if (j < L) ds_out[((size_t)b * L + l) * L + j] = (M[j] - (j == l ? 1.0f : 0.f)) * (dav[jj] - dot);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// part 2: K_ij = -sum_l a_l M_li M_lj + (dS_ij + dS_ji) / sqrt(D), stored transposed (kt[j][i] = K_ij) for the dX pass;
// z_i = sum_l M_li a_l mu_l.  grid (row groups, bags): a block handles rows i = blockIdx.x, blockIdx.x + gridDim.x, ...
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_small_bwd_k_kernel(const float* __restrict__ a_in, const float* __restrict__ ds,
                                                             const float* __restrict__ as_in, const float* __restrict__ mu,
                                                             const float* __restrict__ hsum, float* __restrict__ kt_out,
                                                             float* __restrict__ z_out, float* __restrict__ hbar_out, int L,
                                                             int LP, int D) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* M = smem;                 // [L][LP] A + I
  float* as = M + L * LP;          // [LP]
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < L * L; i += blockDim.x) {
    const int l = i / L, lp = i - l * L;
    M[l * LP + lp] = __ldg(a_in + (size_t)b * L * L + i) + (l == lp ? 1.0f : 0.f);
  }
  for (int l = tid; l < L; l += blockDim.x) as[l] = __ldg(as_in + b * L + l);
  __syncthreads();
  const float inv_sqrt_d = rsqrtf((float)D);
  const float* dsb = ds + (size_t)b * L * L;
  const int nrows = (L - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // rows blockIdx.x, + gridDim.x, ...
  for (int idx = tid; idx < nrows * L; idx += blockDim.x) {
    const int ri = idx / L, j = idx - ri * L, i = blockIdx.x + ri * gridDim.x;
    float s = 0.f;
#pragma unroll 4
    for (int l = 0; l < L; ++l) s = fmaf(as[l] * M[l * LP + i], M[l * LP + j], s);
    kt_out[((size_t)b * L + j) * L + i] = (__ldg(dsb + (size_t)i * L + j) + __ldg(dsb + (size_t)j * L + i)) * inv_sqrt_d - s;
  }
  for (int ri = tid >> 5; ri < nrows; ri += (int)(blockDim.x >> 5)) {
    const int i = blockIdx.x + ri * gridDim.x, lane = tid & 31;
    float z = 0.f;
    for (int l = lane; l < L; l += 32) z = fmaf(M[l * LP + i], as[l] * __ldg(mu + b * L + l), z);
    z = warp_sum(z);
    if (lane == 0) z_out[b * L + i] = z;
  }
  if (blockIdx.x == 0 && tid == 0) hbar_out[b] = __ldg(hsum + b) / (float)D;
}

// ---------------------------------------------------------------------------
// backward pass 2: dx[b, l, d] = sum_l' K_ll' x_l'd + v_l (h_d - hbar) + z_l
// block = 64 columns x 4 row groups; K^T and the X chunk in shared memory
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_dx_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                    const float* __restrict__ gamma, const float* __restrict__ kt,
                                                    const float* __restrict__ v, const float* __restrict__ z,
                                                    const float* __restrict__ hbar, float* __restrict__ dx, int L, int LP, int D) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  float* KT = smem;                // [L][LP]: KT[l'][l] = K_ll'
  float* xs = KT + L * LP;         // [L][BA_DC]
  float* vs = xs + L * BA_DC;      // [LP]
  float* zs = vs + LP;             // [LP]
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < L * LP; i += blockDim.x) {
    const int lp = i / LP, l = i - lp * LP;
    KT[i] = l < L ? __ldg(kt + ((size_t)b * L + lp) * L + l) : 0.f;
  }
  for (int l = tid; l < LP; l += blockDim.x) {
    vs[l] = l < L ? __ldg(v + b * L + l) : 0.f;
    zs[l] = l < L ? __ldg(z + b * L + l) : 0.f;
  }
  const int d0 = blockIdx.x * BA_DC;
  const float* xb = x + (size_t)b * L * D;
  for (int i = tid; i < L * BA_DC; i += blockDim.x) {
    const int l = i / BA_DC, k = i - l * BA_DC;
    xs[i] = d0 + k < D ? __ldg(xb + (size_t)l * D + d0 + k) : 0.f;
  }
  __syncthreads();
  const int k = tid & (BA_DC - 1), grp = tid / BA_DC;        // column, row group (4 groups)
  const int d = d0 + k;
  const float hd = d < D ? __ldg(g + (size_t)b * D + d) * __ldg(gamma + d) / (float)L - __ldg(hbar + b) : 0.f;
  const int nt4 = LP >> 2;
  for (int t = grp; t < nt4; t += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int lp = 0; lp < L; ++lp) {
      const float4 kk = *reinterpret_cast<const float4*>(KT + lp * LP + 4 * t);
      const float xv = xs[lp * BA_DC + k];
      acc[0] = fmaf(kk.x, xv, acc[0]);
      acc[1] = fmaf(kk.y, xv, acc[1]);
      acc[2] = fmaf(kk.z, xv, acc[2]);
      acc[3] = fmaf(kk.w, xv, acc[3]);
    }
    if (d < D) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int l = 4 * t + i;
        if (l < L) dx[((size_t)b * L + l) * D + d] = acc[i] + fmaf(vs[l], hd, zs[l]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int ba_lp(int L) { return (L + 3) & ~3; }

size_t bagattn_saved_floats(int n_bags, int L) {      // m, r, mu, v [L] each, c [1 padded to 4], Gc, A [L, L] per bag
  return (size_t)n_bags * ((size_t)4 * L + 4 + 2 * (size_t)L * L);
}

void launch_bagattn_forward(const float* x, const float* gamma, const float* beta, float* out, float* saved, int n_bags, int L,
                            int D, float eps, cudaStream_t st) {
  const int LP = ba_lp(L);
  float* m = saved;
  float* r = m + (size_t)n_bags * L;
  float* mu = r + (size_t)n_bags * L;
  float* v = mu + (size_t)n_bags * L;
  float* c = v + (size_t)n_bags * L;
  float* gc = c + (size_t)n_bags * 4;
  float* a = gc + (size_t)n_bags * L * L;
  cudaMemsetAsync(gc, 0, (size_t)n_bags * L * L * sizeof(float), st);
  {
    LaunchScope scope("bagattn_rowmean", st);
    launch_k(ba_rowmean_kernel, dim3(n_bags * L), dim3(256), 0, st, x, m, D);
  }
  {
    LaunchScope scope("bagattn_gram", st);
    const int nchunk = ceil_div(D, BA_DC);
    const int per_bag = nchunk < 16 ? nchunk : 16;
    const size_t smem = (size_t)(BA_DC * LP + LP) * sizeof(float);
    cudaFuncSetAttribute(ba_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    launch_k(ba_gram_kernel, dim3(per_bag, n_bags), dim3(256), smem, st, x, (const float*)m, gc, L, LP, D);
  }
  const int rgroups = 8;           // row groups of the L x L kernels: 8 blocks of 8 warps per bag
  {
    LaunchScope scope("bagattn_small_fwd", st);
    const size_t smem = (size_t)(L * LP + LP + 8 * LP) * sizeof(float);
    cudaFuncSetAttribute(ba_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    launch_k(ba_small_fwd_kernel, dim3(rgroups, n_bags), dim3(256), smem, st, (const float*)gc, (const float*)m, a, r, mu, L, LP, D,
             eps);
  }
  {
    LaunchScope scope("bagattn_out", st);
    launch_k(ba_out_kernel, dim3(ceil_div(D, 256), n_bags), dim3(256), 0, st, x, (const float*)a, (const float*)r, (const float*)mu, v,
             c, gamma, beta, out, L, D);
  }
}

size_t bagattn_backward_ws_floats(int n_bags, int L) {   // w', z, a [L], hsum, hbar (padded), K^T, dS [L, L] per bag
  return (size_t)n_bags * ((size_t)3 * L + 8 + 2 * (size_t)L * L);
}

void launch_bagattn_backward(const float* x, const float* g, const float* gamma, const float* saved, float* dx, float* dgamma,
                             float* ws, int n_bags, int L, int D, cudaStream_t st) {
  const int LP = ba_lp(L);
  const float* m = saved;
  const float* r = m + (size_t)n_bags * L;
  const float* mu = r + (size_t)n_bags * L;
  const float* v = mu + (size_t)n_bags * L;
  const float* c = v + (size_t)n_bags * L;
  const float* gc = c + (size_t)n_bags * 4;
  const float* a = gc + (size_t)n_bags * L * L;
  float* wp = ws;
  float* z = wp + (size_t)n_bags * L;
  float* hsum = z + (size_t)n_bags * L;
  float* hbar = hsum + (size_t)n_bags * 4;
  float* as = hbar + (size_t)n_bags * 4;
  float* kt = as + (size_t)n_bags * L;
  float* ds = kt + (size_t)n_bags * L * L;
  cudaMemsetAsync(ws, 0, ((size_t)n_bags * 2 * L + (size_t)n_bags * 8) * sizeof(float), st);
  {
    LaunchScope scope("bagattn_bwd_vec", st);
    launch_k(ba_bwd_vec_kernel, dim3(ceil_div(D, 256), n_bags), dim3(256), 0, st, x, g, v, c, gamma, wp, hsum, dgamma, L, D);
  }
  const int rgroups = 8;
  {
    LaunchScope scope("bagattn_small_bwd_rows", st);
    const size_t smem = (size_t)(L * LP + 2 * LP + 8 * LP) * sizeof(float);
    cudaFuncSetAttribute(ba_small_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    launch_k(ba_small_bwd_rows_kernel, dim3(rgroups, n_bags), dim3(256), smem, st, gc, a, m, r, mu, (const float*)wp,
             (const float*)hsum, ds, as, L, LP, D);
  }
  {
    LaunchScope scope("bagattn_small_bwd_k", st);
    const size_t smem = (size_t)(L * LP + LP) * sizeof(float);
    cudaFuncSetAttribute(ba_small_bwd_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    launch_k(ba_small_bwd_k_kernel, dim3(rgroups * 2, n_bags), dim3(256), smem, st, a, (const float*)ds, (const float*)as, mu,
             (const float*)hsum, kt, z, hbar, L, LP, D);
  }
  {
    LaunchScope scope("bagattn_dx", st);
    const size_t smem = (size_t)(L * LP + L * BA_DC + 2 * LP) * sizeof(float);
    cudaFuncSetAttribute(ba_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    launch_k(ba_dx_kernel, dim3(ceil_div(D, BA_DC), n_bags), dim3(256), smem, st, x, g, gamma, (const float*)kt, v, (const float*)z,
             (const float*)hbar, dx, L, LP, D);
  }
}

}  // namespace bdn
