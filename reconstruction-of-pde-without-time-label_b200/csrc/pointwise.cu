// Per-pixel stages of the FNO nets and the snapshot-bag pool, fp32.
//   lift (+ zero pad)            FNO2d.forward 2d_FPE/FNOModules.py:219-224, FNO1d.forward :103-106
//   crop + fc1 -> GELU -> fc2    FNO2d.forward :234-239, FNO1d.forward :116-121
//   bag mean + detached fc0      NIOFP2D_FNO.forward 2d_FPE/NIOModules.py:564-575
// The per-snapshot input concat(snapshot, grid) of NIO-FNO (2d_FPE/NIOModules.py:555-560) and the
// bag gather x[:, idx] (:548-551) are folded into the lift's addressing; the [B*L, n, n, 128] hidden
// tensor of the projection lives in registers only and is recomputed in backward.
#include "bdn_internal.cuh"

#include <cstdlib>

namespace bdn {

// ===========================================================================
// lift: z0[img, c, hp, wp] = (h < H && w < W) ? b0[c] + sum_i W0[c][i] * in_i : 0
// ===========================================================================
struct LiftParams {
  const float* x_cl; const float* bags; const int32_t* idx; const float* grid;
  int n_keep, bag_len, grid_dim;
  const float* w0; const float* b0;
  int images, c_in, width, h, w, hp, wp;
};


template <int CP, int PX>   // CP = width rounded up to 4 (accumulators live in registers), PX = pixels per thread
__global__ void __launch_bounds__(256) lift_kernel(const LiftParams p, float* __restrict__ z0) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float smem[];
  float* ws = smem;                        // [c_in][CP]  (transposed: one input feature's column of W0)
  float* bs = smem + p.c_in * CP;          // [CP]
  for (int i = threadIdx.x; i < p.c_in * CP; i += blockDim.x) {
    const int ci = i / CP, c = i - ci * CP;
    ws[i] = c < p.width ? __ldg(p.w0 + c * p.c_in + ci) : 0.f;
  }
  for (int i = threadIdx.x; i < CP; i += blockDim.x) bs[i] = i < p.width ? __ldg(p.b0 + i) : 0.f;
  __syncthreads();
  // blockIdx.y = image (block-uniform: the bag / snapshot lookup is scalar work), x = tiles of the padded plane
  const int plane = p.hp * p.wp;
  const float inv_wp = 1.0f / (float)p.wp;
  for (int img = blockIdx.y; img < p.images; img += gridDim.y) {
    const float* src0 = nullptr;           // bags form: the snapshot's plane
    if (p.x_cl == nullptr) {
      const int b = img / p.n_keep, l = img - b * p.n_keep;
      const int snap = p.idx != nullptr ? __ldg(p.idx + l) : l;
      src0 = p.bags + ((size_t)b * p.bag_len + snap) * p.h * p.w;
    }
    for (int q0 = blockIdx.x * blockDim.x * PX + threadIdx.x; q0 < plane; q0 += gridDim.x * blockDim.x * PX) {
      float acc[PX][CP];
      int pix[PX];
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const int q = q0 + u * blockDim.x;
        const int hh = __float2int_rz(((float)q + 0.5f) * inv_wp), ww = q - hh * p.wp;
        pix[u] = (q < plane && hh < p.h && ww < p.w) ? hh * p.w + ww : -1;
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[u][c] = pix[u] >= 0 ? bs[c] : 0.f;
      }
      for (int i = 0; i < p.c_in; ++i) {
        float v[PX];
#pragma unroll
        for (int u = 0; u < PX; ++u) {
          v[u] = 0.f;
          if (pix[u] >= 0) {
            if (p.x_cl != nullptr) v[u] = __ldg(p.x_cl + ((size_t)img * p.h * p.w + pix[u]) * p.c_in + i);
            else v[u] = i == 0 ? __ldg(src0 + pix[u]) : __ldg(p.grid + (size_t)pix[u] * p.grid_dim + (i - 1));
          }
        }
#pragma unroll
        for (int c4 = 0; c4 < CP; c4 += 4) {
          const float4 w = *reinterpret_cast<const float4*>(ws + i * CP + c4);
#pragma unroll
          for (int u = 0; u < PX; ++u) {
            acc[u][c4] = fmaf(w.x, v[u], acc[u][c4]);
            acc[u][c4 + 1] = fmaf(w.y, v[u], acc[u][c4 + 1]);
            acc[u][c4 + 2] = fmaf(w.z, v[u], acc[u][c4 + 2]);
            acc[u][c4 + 3] = fmaf(w.w, v[u], acc[u][c4 + 3]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const int q = q0 + u * blockDim.x;
        if (q >= plane) continue;
        float* dst = z0 + (size_t)img * p.width * plane + q;
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < p.width) dst[(size_t)c * plane] = acc[u][c];
      }
    }
  }
}

// Bags form of the per-snapshot net (width <= 4, grid_dim 1 or 2): a thread owns 4 consecutive padded columns of one row
// and writes one float4 per channel; the generic kernel above spends ~700 instructions per 4 pixels on its run-time
// loops over input features (ncu r1n: 7.7 M warp-instructions for 240 images, 16 us for 28 MB of stores).
template <int GD>
__global__ void __launch_bounds__(256) lift_bags4_kernel(const LiftParams p, float* __restrict__ z0) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float ws[(GD + 1) * 4], bs[4];       // ws[i * 4 + c] = W0[c][i]
  if (threadIdx.x < (GD + 1) * 4) {
    const int i = threadIdx.x >> 2, c = threadIdx.x & 3;
    ws[threadIdx.x] = c < p.width ? __ldg(p.w0 + c * (GD + 1) + i) : 0.f;
  }
  if (threadIdx.x < 4) bs[threadIdx.x] = (int)threadIdx.x < p.width ? __ldg(p.b0 + threadIdx.x) : 0.f;
  __syncthreads();
  const int wq = p.wp >> 2, per_img = p.hp * wq;
  const float inv_wq = 1.0f / (float)wq;
  const size_t plane = (size_t)p.hp * p.wp;
  for (int img = blockIdx.y; img < p.images; img += gridDim.y) {
    const int b = img / p.n_keep, l = img - b * p.n_keep;
    const int snap = p.idx != nullptr ? __ldg(p.idx + l) : l;
    const float* src0 = p.bags + ((size_t)b * p.bag_len + snap) * p.h * p.w;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_img; t += gridDim.x * blockDim.x) {
      const int hh = __float2int_rz(((float)t + 0.5f) * inv_wq), ww = (t - hh * wq) * 4;     // exact for these ranges
      float o[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[c][j] = 0.f;
      if (hh < p.h && ww < p.w) {
        const int pix0 = hh * p.w + ww;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ww + j < p.w) {
            const float v = __ldg(src0 + pix0 + j);
            float g[GD];
#pragma unroll
            for (int d = 0; d < GD; ++d) g[d] = __ldg(p.grid + (size_t)(pix0 + j) * GD + d);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float acc = fmaf(ws[c], v, bs[c]);
#pragma unroll
              for (int d = 0; d < GD; ++d) acc = fmaf(ws[(1 + d) * 4 + c], g[d], acc);
              o[c][j] = acc;
            }
          }
        }
      }
      float* dst = z0 + (size_t)img * p.width * plane + (size_t)hh * p.wp + ww;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < p.width) *reinterpret_cast<float4*>(dst + c * plane) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
    }
  }
}

static LiftParams make_lift_params(const LiftArgs& a) {
  LiftParams p;
  p.x_cl = a.x_cl; p.bags = a.bags; p.idx = a.idx; p.grid = a.grid;
  p.n_keep = a.n_keep; p.bag_len = a.bag_len; p.grid_dim = a.grid_dim;
  p.w0 = a.w0; p.b0 = a.b0;
  p.images = a.images; p.c_in = a.c_in; p.width = a.width; p.h = a.h; p.w = a.w; p.hp = a.hp; p.wp = a.wp;
  return p;
}

void launch_lift(const LiftArgs& a, float* z0, cudaStream_t st) {
  LaunchScope scope("lift", st, a.width);
  const LiftParams p = make_lift_params(a);
  const int plane = a.hp * a.wp;
  const int block = 256;
  static const int bags4_knob = [] { const char* e = getenv("BDN_LIFT_BAGS4"); return e ? atoi(e) : 1; }();   // (tuning knob)
  if (bags4_knob && a.x_cl == nullptr && a.width <= 4 && (a.wp & 3) == 0 && (a.grid_dim == 1 || a.grid_dim == 2) &&
      a.c_in == a.grid_dim + 1 && (reinterpret_cast<uintptr_t>(z0) & 15) == 0) {
    const int per_img = a.hp * (a.wp >> 2);
    dim3 grid((per_img + block - 1) / block, a.images < 65535 ? a.images : 65535);
    if (a.grid_dim == 1) launch_k(lift_bags4_kernel<1>, grid, dim3(block), 0, st, p, z0);
    else launch_k(lift_bags4_kernel<2>, grid, dim3(block), 0, st, p, z0);
    return;
  }
  // few pixels in all (the heads: 4 images): one pixel per thread, 4x the blocks (24 blocks were latency-bound at 9 us)
  const int px = (long)a.images * plane < 148L * block * 4 ? 1 : 4;
  int gx = (plane + px * block - 1) / (px * block);
  // 1-D nets have tiny planes and many images: fold several images' worth of blocks only through grid.y
  dim3 grid(gx, a.images < 65535 ? a.images : 65535);
  const int cp = (a.width + 3) & ~3;
  const size_t smem = (size_t)(a.c_in * cp + cp) * sizeof(float);
#define BDN_LIFT(CPV)                                                                   \
  {                                                                                     \
    if (px == 1) launch_k(lift_kernel<CPV, 1>, grid, dim3(block), smem, st, p, z0);     \
    else launch_k(lift_kernel<CPV, 4>, grid, dim3(block), smem, st, p, z0);             \
  }
  switch (cp) {
    case 4: BDN_LIFT(4) break;
    case 8: BDN_LIFT(8) break;
    case 12: BDN_LIFT(12) break;
    case 16: BDN_LIFT(16) break;
    case 20: BDN_LIFT(20) break;
    case 24: BDN_LIFT(24) break;
    case 28: BDN_LIFT(28) break;
    default: BDN_LIFT(32) break;
  }
#undef BDN_LIFT
}

// lift backward: g_w0[c][i] += sum gz0[c] * in_i, g_b0[c] += sum gz0[c], gx_cl[.., i] = sum_c W0[c][i] gz0[c]
// A block stages a tile of TP pixels (gz0 for all channels, the inputs; every thread loads), computes
// gx from shared memory, and reduces the outer product with 128-bit row-strided shared loads
// (pitch TP + 4: conflict-free).  TP = 64 when there are few pixels (the heads), 256 otherwise.
template <int TP>
__global__ void __launch_bounds__(256) lift_bwd_kernel(const LiftParams p, const float* __restrict__ gz0,
                                                       float* g_w0, float* g_b0, float* __restrict__ gx_cl,
                                                       int tiles_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PT = TP + 4;
  extern __shared__ __align__(16) float smem[];
  const int C = p.width, CI = p.c_in;
  float* gs = smem;                  // [C][PT]
  float* xs = gs + C * PT;           // [CI][PT]
  float* ws = xs + CI * PT;          // [C][CI]
  float* acc = ws + ((C * CI + 3) & ~3);   // [C*CI + C] (+ pad)
  long* s_goff = reinterpret_cast<long*>(acc + ((C * CI + C + 3) & ~3) + 4);   // [TP] offset of the pixel in a gz plane, -1 = dead
  int* s_img = reinterpret_cast<int*>(s_goff + TP);   // [TP] image of the pixel
  int* s_pix = s_img + TP;                            // [TP] pixel index inside the unpadded image
  const int npair = C * CI + C;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < C * CI; i += nt) ws[i] = __ldg(p.w0 + i);
  for (int i = tid; i < ((npair + 3) & ~3); i += nt) acc[i] = 0.f;
  const int plane = p.hp * p.wp, hw = p.h * p.w;
  const long total = (long)p.images * hw;
  const int sub = nt / npair > 0 ? nt / npair : 1;   // pixel sub-slices per pair
  const float inv_tp = 1.0f / (float)TP, inv_ci = 1.0f / (float)CI;
  auto fdiv = [](int n, float inv) { return __float2int_rz(((float)n + 0.5f) * inv); };
  for (int it = 0; it < tiles_per_block; ++it) {
    const long t0 = ((long)blockIdx.x * tiles_per_block + it) * TP;
    if (t0 >= total) break;
    __syncthreads();
    if (tid < TP) {
      const long t = t0 + tid;
      long off = -1;
      int img = 0, pix = 0;
      if (t < total) {
        img = t / hw; pix = t - (long)img * hw;
        const int hh = pix / p.w, ww = pix - hh * p.w;
        off = (long)img * C * plane + hh * p.wp + ww;
      }
      s_goff[tid] = off;
      s_img[tid] = img;
      s_pix[tid] = pix;
    }
    __syncthreads();
    for (int idx = tid; idx < C * TP; idx += nt) {
      const int c = fdiv(idx, inv_tp), tp = idx - c * TP;
      const long off = s_goff[tp];
      gs[c * PT + tp] = off >= 0 ? __ldg(gz0 + off + (size_t)c * plane) : 0.f;
    }
    for (int idx = tid; idx < CI * TP; idx += nt) {
      const int tp = fdiv(idx, inv_ci), i = idx - tp * CI;
      const long t = t0 + tp;
      float v = 0.f;
      if (t < total) {
        if (p.x_cl != nullptr) {
          v = __ldg(p.x_cl + (size_t)t * CI + i);
        } else {
          const int img = s_img[tp], pix = s_pix[tp];
          if (i == 0) {
            const int b = img / p.n_keep, l = img - b * p.n_keep;
            const int snap = p.idx != nullptr ? __ldg(p.idx + l) : l;
            v = __ldg(p.bags + ((size_t)b * p.bag_len + snap) * hw + pix);
          } else {
            v = __ldg(p.grid + (size_t)pix * p.grid_dim + (i - 1));
          }
        }
      }
      xs[i * PT + tp] = v;
    }
    __syncthreads();
    if (gx_cl != nullptr) {
      for (int idx = tid; idx < CI * TP; idx += nt) {
        const int tp = fdiv(idx, inv_ci), i = idx - tp * CI;
        const long t = t0 + tp;
        if (t >= total) continue;
        float gx = 0.f;
        for (int c = 0; c < C; ++c) gx = fmaf(ws[c * CI + i], gs[c * PT + tp], gx);
        gx_cl[(size_t)t * CI + i] = gx;
      }
    }
    for (int item = tid; item < npair * sub; item += nt) {
      const int pair = item % npair, sl = item / npair;
      const int lo = sl * (TP / 4) / sub, hi = (sl + 1) * (TP / 4) / sub;   // float4 chunks
      float s = 0.f;
      if (pair < C * CI) {
        const float4* g = reinterpret_cast<const float4*>(gs + (pair / CI) * PT);
        const float4* x = reinterpret_cast<const float4*>(xs + (pair % CI) * PT);
        for (int q = lo; q < hi; ++q) {
          const float4 a4 = g[q], b4 = x[q];
          s = fmaf(a4.x, b4.x, fmaf(a4.y, b4.y, fmaf(a4.z, b4.z, fmaf(a4.w, b4.w, s))));
        }
      } else {
        const float4* g = reinterpret_cast<const float4*>(gs + (pair - C * CI) * PT);
        for (int q = lo; q < hi; ++q) {
          const float4 a4 = g[q];
          s += (a4.x + a4.y) + (a4.z + a4.w);
        }
      }
      if (sub > 1) atomicAdd(acc + pair, s); else acc[pair] += s;
    }
  }
  __syncthreads();
  const bool vec_ok = ((C * CI) & 3) == 0 && (C & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(g_w0) | reinterpret_cast<uintptr_t>(g_b0)) & 15) == 0;
  if (vec_ok) {
    for (int q = tid; q < npair >> 2; q += nt) {
      float* dst = 4 * q < C * CI ? g_w0 + 4 * q : g_b0 + (4 * q - C * CI);
      atomicAdd(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(acc)[q]);
    }
  } else {
    for (int i = tid; i < npair; i += nt) atomicAdd(i < C * CI ? g_w0 + i : g_b0 + (i - C * CI), acc[i]);
  }
}

// lift backward, narrow nets without an input gradient (the per-snapshot net: width 4, 2-3 input features):
// the C*(CI+1) sums live in registers, one thread walks pixels of one image (grid.y = image, so the bag /
// snapshot lookup is block-uniform), then warp shuffle -> shared -> one atomic per sum per block.
template <int C, int CI>
__global__ void __launch_bounds__(256) lift_bwd_small_kernel(const LiftParams p, const float* __restrict__ gz0,
                                                             float* g_w0, float* g_b0) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][C * (CI + 1)];
  const int plane = p.hp * p.wp, hw = p.h * p.w;
  const float inv_w = 1.0f / (float)p.w;
  float acc[C][CI + 1];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i <= CI; ++i) acc[c][i] = 0.f;
  for (int img = blockIdx.y; img < p.images; img += gridDim.y) {
    const float* src0 = nullptr;
    if (p.x_cl == nullptr) {
      const int b = img / p.n_keep, l = img - b * p.n_keep;
      const int snap = p.idx != nullptr ? __ldg(p.idx + l) : l;
      src0 = p.bags + ((size_t)b * p.bag_len + snap) * hw;
    }
    const float* g = gz0 + (size_t)img * C * plane;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
      const int hh = __float2int_rz(((float)pix + 0.5f) * inv_w), ww = pix - hh * p.w;
      float in[CI];
      if (p.x_cl != nullptr) {
#pragma unroll
        for (int i = 0; i < CI; ++i) in[i] = __ldg(p.x_cl + ((size_t)img * hw + pix) * CI + i);
      } else {
        in[0] = __ldg(src0 + pix);
#pragma unroll
        for (int d = 1; d < CI; ++d) in[d] = __ldg(p.grid + (size_t)pix * (CI - 1) + (d - 1));
      }
      const int off = hh * p.wp + ww;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float gv = __ldg(g + (size_t)c * plane + off);
#pragma unroll
        for (int i = 0; i < CI; ++i) acc[c][i] = fmaf(gv, in[i], acc[c][i]);
        acc[c][CI] += gv;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i <= CI; ++i) {
      const float v = warp_sum(acc[c][i]);
      if (lane == 0) red[warp][c * (CI + 1) + i] = v;
    }
  __syncthreads();
  if (threadIdx.x < C * (CI + 1)) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    const int c = threadIdx.x / (CI + 1), i = threadIdx.x - c * (CI + 1);
    if (c < p.width) {
      if (i < CI) atomicAdd(g_w0 + c * CI + i, v);
      else atomicAdd(g_b0 + c, v);
    }
  }
}

void launch_lift_bwd(const LiftArgs& a, const float* gz0, float* g_w0, float* g_b0, float* gx_cl,
                     cudaStream_t st) {
  LaunchScope scope("lift_bwd", st, a.width);
  const LiftParams p = make_lift_params(a);
  const long total = (long)a.images * a.h * a.w;
  if (gx_cl == nullptr && a.width == 4 && (a.c_in == 2 || a.c_in == 3) && (a.x_cl != nullptr || a.grid_dim == a.c_in - 1)) {
    const int hw = a.h * a.w;
    int gx = ceil_div(hw, 256 * 4);                     // >= 4 pixels per thread per image
    if (gx < 1) gx = 1;
    int gy = a.images;
    while ((long)gx * gy > 148L * 16 && gy > 1) gy = (gy + 1) / 2;   // a few blocks per SM, several images per block
    dim3 grid(gx, gy);
    if (a.c_in == 3) launch_k(lift_bwd_small_kernel<4, 3>, grid, dim3(256), 0, st, p, gz0, g_w0, g_b0);
    else launch_k(lift_bwd_small_kernel<4, 2>, grid, dim3(256), 0, st, p, gz0, g_w0, g_b0);
    return;
  }
  const bool few = total < 148L * 256 * 2;
  const int tp = few ? 64 : 256;
  const int tiles = (int)((total + tp - 1) / tp);
  int tpb = 1;
  while (ceil_div(tiles, tpb) > 4 * 148) ++tpb;
  const int grid = ceil_div(tiles, tpb);
  const size_t smem = (size_t)((a.width + a.c_in) * (tp + 4) + ((a.width * a.c_in + 3) & ~3) +
                               ((a.width * a.c_in + a.width + 3) & ~3) + 4) * sizeof(float) +
                      (size_t)tp * (sizeof(long) + 2 * sizeof(int)) + 16;
  if (few) {
    cudaFuncSetAttribute(lift_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    launch_k(lift_bwd_kernel<64>, dim3(grid), dim3(256), smem, st, p, gz0, g_w0, g_b0, gx_cl, tpb);
  } else {
    cudaFuncSetAttribute(lift_bwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    launch_k(lift_bwd_kernel<256>, dim3(grid), dim3(256), smem, st, p, gz0, g_w0, g_b0, gx_cl, tpb);
  }
}

// ===========================================================================
// projection: out[pix, o] = b2[o] + sum_j W2[o][j] * gelu(b1[j] + sum_c W1[j][c] * z[c, pix])
//
// Thread mapping: a warp holds 32/JS pixels x JS hidden-unit slices (lane = pixel_slot * JS + slice);
// a thread owns PP pixels (tile-strided) and the hidden units j = slice, slice + JS, ...  JS = 1 is
// the many-pixel regime (per-snapshot net: one thread per pixel, everything in registers); JS = 8
// spreads the 128 hidden units of a pixel over 8 lanes so that the few-image heads still fill the
// machine.  The [pixels, hidden] tensor exists only in registers, forward and backward.
// ===========================================================================
constexpr int PROJ_MAX_OUT = 4;
constexpr int PROJ_THREADS = 256;

template <int CP>
__device__ __forceinline__ void load_w1_row(const float* w1s, int j, float (&w)[CP]) {
#pragma unroll
  for (int c4 = 0; c4 < CP; c4 += 4) {
    const float4 v = *reinterpret_cast<const float4*>(w1s + j * CP + c4);
    w[c4] = v.x; w[c4 + 1] = v.y; w[c4 + 2] = v.z; w[c4 + 3] = v.w;
  }
}

template <int CP, int NOUT, int JS, int PP>   // CP = width rounded up to 4; NOUT = 1 or PROJ_MAX_OUT (runtime c_out)
__global__ void __launch_bounds__(PROJ_THREADS) project_kernel(const ProjArgs a, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int hidden = a.hidden, nout = NOUT == 1 ? 1 : a.c_out;
  float* w1s = smem;                               // [hidden][CP]
  float* b1s = w1s + hidden * CP;                  // [hidden]
  float* w2s = b1s + hidden;                       // [nout][hidden]
  for (int i = threadIdx.x; i < hidden * CP; i += blockDim.x) {
    const int j = i / CP, c = i - j * CP;
    w1s[i] = c < a.width ? __ldg(a.w1 + j * a.width + c) : 0.f;
  }
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) b1s[i] = __ldg(a.b1 + i);
  for (int i = threadIdx.x; i < nout * hidden; i += blockDim.x) w2s[i] = __ldg(a.w2 + i);
  __syncthreads();

  constexpr int SLOTS = PROJ_THREADS / JS;         // pixels per tile row
  constexpr int TILE = SLOTS * PP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = lane % JS, slot = warp * (32 / JS) + lane / JS;
  const int opix = a.out_h * a.out_w, plane = a.hp * a.wp;
  const long total = (long)a.images * opix;
  const long ntiles = (total + TILE - 1) / TILE;

  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tile + gridDim.x >= ntiles) pdl_trigger_late();      // the block's last tile
    float z[PP][CP], o[PP][NOUT];
    long t[PP];
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
      t[pp] = tile * TILE + (long)pp * SLOTS + slot;
#pragma unroll
      for (int k = 0; k < NOUT; ++k) o[pp][k] = 0.f;
      if (t[pp] < total) {
        const int img = t[pp] / opix, q = t[pp] - (long)img * opix;
        const int oh = q / a.out_w, ow = q - oh * a.out_w;
        const float* zp = a.z + (size_t)img * a.width * plane + oh * a.wp + ow;
#pragma unroll
        for (int c = 0; c < CP; ++c) z[pp][c] = c < a.width ? __ldg(zp + (size_t)c * plane) : 0.f;
      } else {
#pragma unroll
        for (int c = 0; c < CP; ++c) z[pp][c] = 0.f;
      }
    }
#pragma unroll 2
    for (int j = slice; j < hidden; j += JS) {
      float w1r[CP], w2r[NOUT];
      load_w1_row<CP>(w1s, j, w1r);
      const float b1j = b1s[j];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) w2r[k] = k < nout ? w2s[k * hidden + j] : 0.f;
#pragma unroll
      for (int pp = 0; pp < PP; ++pp) {
        float h = b1j;
#pragma unroll
        for (int c = 0; c < CP; ++c) h = fmaf(w1r[c], z[pp][c], h);
        const float g = gelu_fast(h);
#pragma unroll
        for (int k = 0; k < NOUT; ++k) o[pp][k] = fmaf(w2r[k], g, o[pp][k]);
      }
    }
#pragma unroll
    for (int pp = 0; pp < PP; ++pp)
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
#pragma unroll
        for (int off = JS >> 1; off > 0; off >>= 1) o[pp][k] += __shfl_xor_sync(0xffffffffu, o[pp][k], off);
        if (slice == 0 && t[pp] < total && k < nout) out[(size_t)t[pp] * nout + k] = o[pp][k] + __ldg(a.b2 + k);
      }
  }
}

template <typename F>
static bool dispatch_cp(int width, F&& f) {
  const int cp = (width + 3) & ~3;
  switch (cp) {
    case 4: f(std::integral_constant<int, 4>()); return true;
    case 8: f(std::integral_constant<int, 8>()); return true;
    case 12: f(std::integral_constant<int, 12>()); return true;
    case 16: f(std::integral_constant<int, 16>()); return true;
    case 20: f(std::integral_constant<int, 20>()); return true;
    case 24: f(std::integral_constant<int, 24>()); return true;
    case 28: f(std::integral_constant<int, 28>()); return true;
    case 32: f(std::integral_constant<int, 32>()); return true;
    default: return false;
  }
}

// many pixels -> one thread per pixel; few pixels -> 8 lanes per pixel
static bool proj_many_pixels(long total) { return total >= 148L * PROJ_THREADS * 4; }

template <int CP, int NOUT, int JS, int PP>
static void launch_project_t(const ProjArgs& a, float* out, long total, cudaStream_t st) {
  constexpr int TILE = PROJ_THREADS / JS * PP;
  const long ntiles = (total + TILE - 1) / TILE;
  static const long fcap = [] { const char* e = getenv("BDN_PROJ_FWD_CAP"); return e ? atol(e) : 148L * 8; }();   // (tuning knob)
  const int grid = (int)(ntiles < fcap ? ntiles : fcap);
  const size_t smem = (size_t)(a.hidden * CP + a.hidden + a.c_out * a.hidden) * sizeof(float);
  cudaFuncSetAttribute(project_kernel<CP, NOUT, JS, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);   // hidden = 512, width 32, 4 outputs: 76 KB
  launch_k(project_kernel<CP, NOUT, JS, PP>, dim3(grid), dim3(PROJ_THREADS), smem, st, a, out);
}

void launch_project(const ProjArgs& a, float* out, cudaStream_t st) {
  LaunchScope scope("project", st, a.width);
  const long total = (long)a.images * a.out_h * a.out_w;
  const bool many = proj_many_pixels(total);
  static const int ppf = [] { const char* e = getenv("BDN_PROJ_FWD_PP"); return e ? atoi(e) : 4; }();   // (tuning knob: 2, 4, 8)
  dispatch_cp(a.width, [&](auto cp) {
    constexpr int CP = decltype(cp)::value;
    constexpr int PPM = CP <= 8 ? 4 : 2;
    if (a.c_out == 1) {
      if (many && CP <= 8 && ppf == 2) launch_project_t<CP, 1, 1, 2>(a, out, total, st);
      else if (many && CP <= 8 && ppf == 8) launch_project_t<CP, 1, 1, 8>(a, out, total, st);
      else if (many) launch_project_t<CP, 1, 1, PPM>(a, out, total, st);
      else launch_project_t<CP, 1, 8, 1>(a, out, total, st);
    } else {
      if (many) launch_project_t<CP, PROJ_MAX_OUT, 1, 2>(a, out, total, st);
      else launch_project_t<CP, PROJ_MAX_OUT, 8, 1>(a, out, total, st);
    }
  });
}

template <int CP>
__host__ __device__ constexpr int proj_nvp() { return CP + 2 <= 8 ? 8 : (CP + 2 <= 16 ? 16 : 32); }
template <int CP, int NOUT, int JS>
__host__ __device__ constexpr bool proj_treduce() { return NOUT == 1 && JS == 1 && CP <= 12; }
// few-pixel regime (JS = 8): the 4 lanes that share a hidden-unit slice transpose-reduce their CP + 1 + NOUT partial
// sums in 2 butterfly steps (0.75 shuffles per value instead of 2) and add them to warp-private accumulators
template <int CP, int NOUT, int JS>
__host__ __device__ constexpr bool proj_qreduce() { return JS == 8; }
template <int CP, int NOUT>
__host__ __device__ constexpr int proj_nvq() { return (CP + 1 + NOUT + 3) & ~3; }
template <int CP, int NOUT, int JS>
__host__ __device__ constexpr int proj_acc_pitch() {
  return proj_treduce<CP, NOUT, JS>() ? proj_nvp<CP>() : (proj_qreduce<CP, NOUT, JS>() ? proj_nvq<CP, NOUT>() : 0);
}

// projection backward.  Same thread mapping; for every hidden unit the pre-activation is
// recomputed, gz accumulates in registers (reduced over the JS slices at the end), and the weight
// gradient partials are reduced over the lanes that share a slice, then added to per-block shared
// accumulators that are flushed once per block with coalesced global atomics.
template <int CP, int NOUT, int JS, int PP>
__global__ void __launch_bounds__(PROJ_THREADS) project_bwd_kernel(const ProjArgs a, const float* __restrict__ g_out,
                                                                   int pooled_g, int n_keep, float* __restrict__ gz,
                                                                   float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                                                   int wred) {   // warp-private accumulators fit shared memory
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem[];
  const int hidden = a.hidden, nout = NOUT == 1 ? 1 : a.c_out;
  float* w1s = smem;                       // [hidden][CP]
  float* b1s = w1s + hidden * CP;          // [hidden]
  float* w2s = b1s + hidden;               // [nout][hidden]
  float* aw1 = w2s + nout * hidden;        // [hidden][CP]  accumulators
  float* ab1 = aw1 + hidden * CP;          // [hidden]
  float* aw2 = ab1 + hidden;               // [nout][hidden]
  float* ab2 = aw2 + nout * hidden;        // [PROJ_MAX_OUT]
  constexpr bool TREDUCE = proj_treduce<CP, NOUT, JS>();
  const bool QREDUCE = proj_qreduce<CP, NOUT, JS>() && wred != 0;
  constexpr int NVP = TREDUCE ? proj_nvp<CP>() : proj_nvq<CP, NOUT>();
  float* wacc = ab2 + PROJ_MAX_OUT;        // [warps][hidden][NVP] warp-private accumulators (TREDUCE / QREDUCE)
  if (TREDUCE || QREDUCE)
    for (int i = threadIdx.x; i < (PROJ_THREADS / 32) * hidden * NVP; i += blockDim.x) wacc[i] = 0.f;
  for (int i = threadIdx.x; i < hidden * CP; i += blockDim.x) {
    const int j = i / CP, c = i - j * CP;
    w1s[i] = c < a.width ? __ldg(a.w1 + j * a.width + c) : 0.f;
    aw1[i] = 0.f;
  }
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) { b1s[i] = __ldg(a.b1 + i); ab1[i] = 0.f; }
  for (int i = threadIdx.x; i < nout * hidden; i += blockDim.x) { w2s[i] = __ldg(a.w2 + i); aw2[i] = 0.f; }
  if (threadIdx.x < PROJ_MAX_OUT) ab2[threadIdx.x] = 0.f;
  __syncthreads();

  constexpr int SLOTS = PROJ_THREADS / JS;
  constexpr int TILE = SLOTS * PP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = lane % JS, slot = warp * (32 / JS) + lane / JS;
  const int opix = a.out_h * a.out_w, plane = a.hp * a.wp;
  const long total = (long)a.images * opix;
  const long ntiles = (total + TILE - 1) / TILE;
  const float gscale = pooled_g ? 1.0f / (float)n_keep : 1.0f;

  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tile + gridDim.x >= ntiles) pdl_trigger_late();      // the block's last tile
    float z[PP][CP], g[PP][NOUT], gzr[PP][CP];
    size_t zoff[PP];
    bool live[PP];
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
      const long t = tile * TILE + (long)pp * SLOTS + slot;
      live[pp] = t < total;
      zoff[pp] = 0;
#pragma unroll
      for (int c = 0; c < CP; ++c) { z[pp][c] = 0.f; gzr[pp][c] = 0.f; }
#pragma unroll
      for (int k = 0; k < NOUT; ++k) g[pp][k] = 0.f;
      if (live[pp]) {
        const int img = t / opix, q = t - (long)img * opix;
        const int oh = q / a.out_w, ow = q - oh * a.out_w;
        zoff[pp] = (size_t)img * a.width * plane + oh * a.wp + ow;
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < a.width) z[pp][c] = __ldg(a.z + zoff[pp] + (size_t)c * plane);
        const size_t goff = pooled_g ? ((size_t)(img / n_keep) * opix + q) * nout : (size_t)t * nout;
#pragma unroll
        for (int k = 0; k < NOUT; ++k)
          if (k < nout) g[pp][k] = __ldg(g_out + goff + k) * gscale;
      }
    }
    // fc2 bias gradient: every pixel once (slice 0 holds it)
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      float s = 0.f;
#pragma unroll
      for (int pp = 0; pp < PP; ++pp) s += g[pp][k];
      if (slice != 0) s = 0.f;
      s = warp_sum(s);
      if (lane == 0 && k < nout) atomicAdd(ab2 + k, s);
    }
#pragma unroll 1
    for (int j = slice; j < hidden; j += JS) {
      float w1r[CP], w2r[NOUT];
      load_w1_row<CP>(w1s, j, w1r);
      const float b1j = b1s[j];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) w2r[k] = k < nout ? w2s[k * hidden + j] : 0.f;
      float sw1[CP], sb1 = 0.f, sw2[NOUT];
#pragma unroll
      for (int c = 0; c < CP; ++c) sw1[c] = 0.f;
#pragma unroll
      for (int k = 0; k < NOUT; ++k) sw2[k] = 0.f;
#pragma unroll
      for (int pp = 0; pp < PP; ++pp) {
        float h = b1j;
#pragma unroll
        for (int c = 0; c < CP; ++c) h = fmaf(w1r[c], z[pp][c], h);
        float cdf, pdf;
        gelu_cdf_pdf(h, cdf, pdf);
        const float gel = h * cdf;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < NOUT; ++k) { d = fmaf(g[pp][k], w2r[k], d); sw2[k] = fmaf(g[pp][k], gel, sw2[k]); }
        const float e = d * fmaf(h, pdf, cdf);
        sb1 += e;
#pragma unroll
        for (int c = 0; c < CP; ++c) { gzr[pp][c] = fmaf(e, w1r[c], gzr[pp][c]); sw1[c] = fmaf(e, z[pp][c], sw1[c]); }
      }
      if constexpr (TREDUCE) {
        // Transpose-reduce: the NV = CP + 2 per-lane partials (padded to NVP, a power of two) are summed
        // over the warp so that lane (v * 32 / NVP) ends up with the total of value v: NVP - 1 + log2(32 / NVP)
        // shuffles instead of 5 * NV, and the totals are added to warp-private shared accumulators (no atomics).
        float val[NVP];
#pragma unroll
        for (int c = 0; c < CP; ++c) val[c] = sw1[c];
        val[CP] = sb1;
        val[CP + 1] = sw2[0];
#pragma unroll
        for (int v = CP + 2; v < NVP; ++v) val[v] = 0.f;
        int n = NVP;
#pragma unroll
        for (int off = 16; off >= 32 / NVP; off >>= 1) {
          const bool up = (lane & off) != 0;
          n >>= 1;
#pragma unroll
          for (int i = 0; i < NVP / 2; ++i) {
            if (i < n) {
              const float send = up ? val[i] : val[i + n];
              const float keep = up ? val[i + n] : val[i];
              val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        }
#pragma unroll
        for (int off = 16 / NVP; off > 0; off >>= 1) val[0] += __shfl_xor_sync(0xffffffffu, val[0], off);
        if ((lane & (32 / NVP - 1)) == 0) {
          float* dst = wacc + ((size_t)warp * hidden + j) * NVP + lane / (32 / NVP);
          *dst += val[0];
        }
      } else if (QREDUCE) {
        // lanes {s, s+8, s+16, s+24} hold the partials of hidden unit j (slice s): after the two steps lane
        // (bit4, bit3) owns the totals of quarter 2*bit4 + bit3 of the value vector [sw1[CP], sb1, sw2[NOUT], 0..]
        float val[NVP];
#pragma unroll
        for (int c = 0; c < CP; ++c) val[c] = sw1[c];
        val[CP] = sb1;
#pragma unroll
        for (int k = 0; k < NOUT; ++k) val[CP + 1 + k] = sw2[k];
#pragma unroll
        for (int v = CP + 1 + NOUT; v < NVP; ++v) val[v] = 0.f;
        int base = 0;
#pragma unroll
        for (int step = 0; step < 2; ++step) {
          const int off = 16 >> step;
          const int n = NVP >> (step + 1);
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < NVP / 2; ++i) {
            if (i < n) {
              const float send = up ? val[i] : val[i + n];
              const float keep = up ? val[i + n] : val[i];
              val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          if (up) base += n;
        }
        float* dst = wacc + ((size_t)warp * hidden + j) * NVP + base;
#pragma unroll
        for (int i = 0; i < NVP / 4; ++i) dst[i] += val[i];
      } else {
        // reduce over the lanes that share this slice (xor offsets >= JS), then one lane per slice adds
#pragma unroll
        for (int off = 16; off >= JS; off >>= 1) {
#pragma unroll
          for (int c = 0; c < CP; ++c) sw1[c] += __shfl_xor_sync(0xffffffffu, sw1[c], off);
          sb1 += __shfl_xor_sync(0xffffffffu, sb1, off);
#pragma unroll
          for (int k = 0; k < NOUT; ++k) sw2[k] += __shfl_xor_sync(0xffffffffu, sw2[k], off);
        }
        if (lane < JS) {
#pragma unroll
          for (int c = 0; c < CP; ++c) atomicAdd(aw1 + j * CP + c, sw1[c]);
          atomicAdd(ab1 + j, sb1);
#pragma unroll
          for (int k = 0; k < NOUT; ++k)
            if (k < nout) atomicAdd(aw2 + k * hidden + j, sw2[k]);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
#pragma unroll
      for (int c = 0; c < CP; ++c) {
#pragma unroll
        for (int off = JS >> 1; off > 0; off >>= 1) gzr[pp][c] += __shfl_xor_sync(0xffffffffu, gzr[pp][c], off);
        if (slice == 0 && live[pp] && c < a.width) gz[zoff[pp] + (size_t)c * plane] = gzr[pp][c];
      }
    }
  }
  __syncthreads();
  if (TREDUCE || QREDUCE) {
    constexpr int NV = CP + 1 + NOUT;      // (TREDUCE has NOUT == 1)
    for (int i = threadIdx.x; i < hidden * NV; i += blockDim.x) {
      const int j = i / NV, v = i - j * NV;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PROJ_THREADS / 32; ++w) s += wacc[((size_t)w * hidden + j) * NVP + v];
      if (v < CP) aw1[j * CP + v] = s;
      else if (v == CP) ab1[j] = s;
      else if (v - CP - 1 < nout) aw2[(v - CP - 1) * hidden + j] = s;
    }
    __syncthreads();
  }
  // flush: 128-bit atomics where the layout allows (4x fewer L2 atomic operations on contended lines)
  const bool h4 = (hidden & 3) == 0;       // keeps every shared accumulator array 16-byte aligned
  auto vec_ok = [h4](const void* ptr, int n) { return h4 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (n & 3) == 0; };
  if (a.width == CP && vec_ok(g_w1, hidden * CP)) {
    for (int i = threadIdx.x; i < (hidden * CP) >> 2; i += blockDim.x)
      atomicAdd(reinterpret_cast<float4*>(g_w1) + i, reinterpret_cast<const float4*>(aw1)[i]);
  } else {
    for (int i = threadIdx.x; i < hidden * a.width; i += blockDim.x) {
      const int j = i / a.width, c = i - j * a.width;
      atomicAdd(g_w1 + i, aw1[j * CP + c]);
    }
  }
  if (vec_ok(g_b1, hidden)) {
    for (int i = threadIdx.x; i < hidden >> 2; i += blockDim.x)
      atomicAdd(reinterpret_cast<float4*>(g_b1) + i, reinterpret_cast<const float4*>(ab1)[i]);
  } else {
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) atomicAdd(g_b1 + i, ab1[i]);
  }
  if (vec_ok(g_w2, nout * hidden)) {
    for (int i = threadIdx.x; i < (nout * hidden) >> 2; i += blockDim.x)
      atomicAdd(reinterpret_cast<float4*>(g_w2) + i, reinterpret_cast<const float4*>(aw2)[i]);
  } else {
    for (int i = threadIdx.x; i < nout * hidden; i += blockDim.x) atomicAdd(g_w2 + i, aw2[i]);
  }
  if (threadIdx.x < nout) atomicAdd(g_b2 + threadIdx.x, ab2[threadIdx.x]);
}

template <int CP, int NOUT, int JS, int PP>
static void launch_project_bwd_t(const ProjArgs& a, const float* g_out, int pooled_g, int n_keep, float* gz,
                                 float* g_w1, float* g_b1, float* g_w2, float* g_b2, long total, cudaStream_t st) {
  constexpr int TILE = PROJ_THREADS / JS * PP;
  const long ntiles = (total + TILE - 1) / TILE;
  static const long cap8 = [] { const char* e = getenv("BDN_PROJ_BWD_CAP8"); return e ? atol(e) : 148L * 2; }();   // (tuning knob; r2k: 4 blocks per SM 88.5, 2 per SM 78.8, 1 per SM 87.4 us per step)
  const long cap = JS == 1 ? 148L * 2 : cap8;
  const long rounds = (ntiles + cap - 1) / cap;
  const int grid = (int)((ntiles + rounds - 1) / rounds);      // balanced: every block runs `rounds` tiles
  size_t smem = (size_t)(2 * (a.hidden * CP + a.hidden + a.c_out * a.hidden) + PROJ_MAX_OUT) * sizeof(float);
  const size_t acc = (size_t)(PROJ_THREADS / 32) * a.hidden * proj_acc_pitch<CP, NOUT, JS>() * sizeof(float);
  const int wred = smem + acc <= 220 * 1024;      // (always true for the reference's hidden = 128)
  if (wred || proj_treduce<CP, NOUT, JS>()) smem += acc;
  cudaFuncSetAttribute(project_bwd_kernel<CP, NOUT, JS, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  launch_k(project_bwd_kernel<CP, NOUT, JS, PP>, dim3(grid), dim3(PROJ_THREADS), smem, st, a, g_out, pooled_g, n_keep, gz, g_w1, g_b1,
                                                                        g_w2, g_b2, wred);
}

void launch_project_bwd(const ProjArgs& a, const float* g_out, int pooled_g, int n_keep, float* gz, float* g_w1,
                        float* g_b1, float* g_w2, float* g_b2, cudaStream_t st) {
  LaunchScope scope("project_bwd", st, a.width);
  const size_t act_bytes = (size_t)a.images * a.width * a.hp * a.wp * sizeof(float);
  cudaMemsetAsync(gz, 0, act_bytes, st);
  const long total = (long)a.images * a.out_h * a.out_w;
  const bool many = proj_many_pixels(total);
  // few pixels: pixels per thread between two reductions of the weight-gradient partials (tuning knob)
  // (r2x, heads: 1 pixel 38.2 us per launch, 2 pixels 27.1 us)
  static const int pp8 = [] { const char* e = getenv("BDN_PROJ_BWD_PP8"); return e ? atoi(e) : 2; }();
  dispatch_cp(a.width, [&](auto cp) {
    constexpr int CP = decltype(cp)::value;
    constexpr int PPM = CP <= 4 ? 8 : (CP <= 12 ? 4 : 2);
    if (a.c_out == 1) {
      if (many) launch_project_bwd_t<CP, 1, 1, PPM>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
      else if (pp8 == 2) launch_project_bwd_t<CP, 1, 8, 2>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
      else if (pp8 == 4) launch_project_bwd_t<CP, 1, 8, 4>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
      else launch_project_bwd_t<CP, 1, 8, 1>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
    } else {
      if (many) launch_project_bwd_t<CP, PROJ_MAX_OUT, 1, 2>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
      else launch_project_bwd_t<CP, PROJ_MAX_OUT, 8, 1>(a, g_out, pooled_g, n_keep, gz, g_w1, g_b1, g_w2, g_b2, total, st);
    }
  });
}

// ===========================================================================
// bag mean + detached lift
// ===========================================================================
__global__ void pool_lift_kernel(const float* __restrict__ s, const float* __restrict__ grid,
                                 const float* __restrict__ w0, const float* __restrict__ b0,
                                 float* __restrict__ out, int n_bags, int n_keep, int npix, int gd, int width) {
  pdl_launch_dependents();
  pdl_wait();
  // 8 lanes share one pixel and split the bag; 4 pixels per warp (each lane group reads 4 consecutive pixels' worth
  // of 32-byte sectors together)
  const long total = (long)n_bags * npix;
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 2, pl = lane & 3;          // sub: bag slice 0..7, pl: pixel within the warp's group of 4
  const long t = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + pl;
  float acc = 0.f;
  int b = 0, pix = 0;
  if (t < total) {
    b = t / npix; pix = t - (long)b * npix;
    const float* sp = s + (size_t)b * n_keep * npix + pix;
    for (int l = sub; l < n_keep; l += 8) acc += __ldg(sp + (size_t)l * npix);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (t >= total) return;
  const float mean = acc / (float)n_keep;
  for (int j = sub; j < width; j += 8) {
    float v = __ldg(b0 + j);
    for (int d = 0; d < gd; ++d) v = fmaf(__ldg(w0 + j * (gd + 1) + d), __ldg(grid + (size_t)pix * gd + d), v);
    out[(size_t)t * width + j] = fmaf(__ldg(w0 + j * (gd + 1) + gd), mean, v);
  }
}

void launch_pool_lift(const float* s, const float* grid, const float* w0, const float* b0, float* out, int n_bags,
                      int n_keep, int npix, int grid_dim, int width, cudaStream_t st) {
  LaunchScope scope("pool_lift", st);
  const long total = (long)n_bags * npix;
  const int block = 128;                              // 4 warps x 4 pixels
  const long pix_per_block = (block / 32) * 4;
  launch_k(pool_lift_kernel, dim3((unsigned)((total + pix_per_block - 1) / pix_per_block)), dim3(block), 0, st, s, grid, w0,
           b0, out, n_bags, n_keep, npix, grid_dim, width);
}

__global__ void pool_lift_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w0,
                                     float* __restrict__ gpool, long total, int gd, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  float acc = 0.f;
  for (int j = 0; j < width; ++j) acc = fmaf(__ldg(w0 + j * (gd + 1) + gd), __ldg(g + (size_t)t * width + j), acc);
  gpool[t] = acc;
}

void launch_pool_lift_bwd(const float* g, const float* w0, float* gpool, int n_bags, int npix, int grid_dim,
                          int width, cudaStream_t st) {
  LaunchScope scope("pool_lift_bwd", st);
  const long total = (long)n_bags * npix;
  const int block = 128;
  launch_k(pool_lift_bwd_kernel, dim3((int)((total + block - 1) / block)), dim3(block), 0, st, g, w0, gpool, total, grid_dim, width);
}

// ===========================================================================
// NIO tail (K6): bag mean of the branch coefficients -> DeepONet contraction with the trunk basis -> (+ b0) / sqrt(p)
// -> detached fc0 lift, one kernel.  DeepOnetNoBiasOrg.forward */DeepONetModules.py:142-151 followed by the bag mean +
// lift of NIOFP_schrodinger.forward 1d_GPE/NIOModules.py:209-219 (NIOFP2D.forward 2d_FPE/NIOModules.py:64-76).  By
// linearity the mean over the bag is taken on the [B, L, p] coefficients, so the [B, L, n_points] DeepONet output of
// the reference is never formed.
//   out[b, x, j] = fc0_b[j] + sum_d W0[j, d] grid[x, d] + W0[j, gd] * (sum_q wbar[b, q] basis[x, q] + b0) / sqrt(p)
// ===========================================================================
constexpr int TAIL_MAXP = 256;

__global__ void __launch_bounds__(128) nio_tail_fwd_kernel(const float* __restrict__ w, const float* __restrict__ basis,
                                                           const float* __restrict__ b0, const float* __restrict__ grid,
                                                           const float* __restrict__ w0, const float* __restrict__ fb,
                                                           float* __restrict__ out, float* __restrict__ wbar_out, int L, int p,
                                                           int npix, int gd, int width) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float wbar[TAIL_MAXP];
  const int b = blockIdx.y;
  for (int q = threadIdx.x; q < p; q += blockDim.x) {
    const float* wp = w + (size_t)b * L * p + q;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += __ldg(wp + (size_t)l * p);
    s /= (float)L;
    wbar[q] = s;
    if (blockIdx.x == 0) wbar_out[b * p + q] = s;
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= npix) return;
  float s = __ldg(b0);
  const float* bx = basis + (size_t)x * p;
  for (int q = 0; q < p; ++q) s = fmaf(wbar[q], __ldg(bx + q), s);
  s *= rsqrtf((float)p);
  float* o = out + ((size_t)b * npix + x) * width;
  for (int j = 0; j < width; ++j) {
    float v = __ldg(fb + j);
    for (int d = 0; d < gd; ++d) v = fmaf(__ldg(w0 + j * (gd + 1) + d), __ldg(grid + (size_t)x * gd + d), v);
    o[j] = fmaf(__ldg(w0 + j * (gd + 1) + gd), s, v);
  }
}

// backward: gs[b, x] = sum_j W0[j, gd] g[b, x, j] / sqrt(p);  g_wbar[b, q] += sum_x gs basis[x, q];
//           g_basis[x, q] += sum_b gs wbar[b, q];  g_b0 += sum gs          (all three zeroed by the caller)
__global__ void __launch_bounds__(128) nio_tail_bwd_kernel(const float* __restrict__ g, const float* __restrict__ basis,
                                                           const float* __restrict__ wbar, const float* __restrict__ w0,
                                                           float* g_wbar, float* g_basis, float* g_b0, int p, int npix, int gd,
                                                           int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  float gs = 0.f;
  if (x < npix) {
    const float* gp = g + ((size_t)b * npix + x) * width;
    for (int j = 0; j < width; ++j) gs = fmaf(__ldg(w0 + j * (gd + 1) + gd), __ldg(gp + j), gs);
    gs *= rsqrtf((float)p);
  }
  for (int q = 0; q < p; ++q) {
    const float bv = x < npix ? __ldg(basis + (size_t)x * p + q) : 0.f;
    if (x < npix) atomicAdd(g_basis + (size_t)x * p + q, gs * __ldg(wbar + b * p + q));
    const float t = warp_sum(gs * bv);
    if (lane == 0) atomicAdd(g_wbar + b * p + q, t);
  }
  const float t = warp_sum(gs);
  if (lane == 0) atomicAdd(g_b0, t);
}

// g_w[b, l, q] = g_wbar[b, q] / L  (backward of the bag mean)
__global__ void nio_tail_expand_kernel(const float* __restrict__ g_wbar, float* __restrict__ g_w, int n_bags, int L, int p) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)n_bags * L * p;
  const float inv = 1.0f / (float)L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int q = i % p, b = i / ((long)L * p);
    g_w[i] = __ldg(g_wbar + b * p + q) * inv;
  }
}

void launch_nio_tail(const float* w, const float* basis, const float* b0, const float* grid, const float* w0, const float* fb,
                     float* out, float* wbar, int n_bags, int L, int p, int npix, int gd, int width, cudaStream_t st) {
  LaunchScope scope("nio_tail", st);
  launch_k(nio_tail_fwd_kernel, dim3(ceil_div(npix, 128), n_bags), dim3(128), 0, st, w, basis, b0, grid, w0, fb, out, wbar, L, p,
           npix, gd, width);
}

void launch_nio_tail_bwd(const float* g, const float* basis, const float* wbar, const float* w0, float* g_wbar, float* g_basis,
                         float* g_b0, float* g_w, int n_bags, int L, int p, int npix, int gd, int width, cudaStream_t st) {
  {
    LaunchScope scope("nio_tail_bwd", st);
    launch_k(nio_tail_bwd_kernel, dim3(ceil_div(npix, 128), n_bags), dim3(128), 0, st, g, basis, wbar, w0, g_wbar, g_basis, g_b0,
             p, npix, gd, width);
  }
  LaunchScope scope("nio_tail_expand", st);
  const long total = (long)n_bags * L * p;
  launch_k(nio_tail_expand_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const float*)g_wbar, g_w, n_bags, L, p);
}

// ===========================================================================
// Training loss of the bag models in one launch: MSE between the concatenated head outputs and the target
// (criterion = torch.nn.MSELoss() on model(inputs, grid), 2d_FPE/train_fno.py:116,146-147), without forming the
// concatenation.  Deterministic: per-block partial sums, the last block to finish adds them in block order.
//   outs[k]: [npix, c] (head k), target: [npix, n_heads * c], loss = mean((cat_k outs[k] - target)^2)
// backward: g[k] = (2 / total) * grad_loss * (outs[k] - target_k)
// ===========================================================================
constexpr int MSE_BLOCKS = 64;

// Thread = pixel, static loops over the heads (constant indices into the pointer arrays of the argument struct: a
// run-time head index would copy the struct to local memory) -- no index divisions.  32-bit arithmetic: the entry
// point bounds the element count.
template <bool WITH_G>
__device__ __forceinline__ float mse_pixel(const MseHeadsArgs& a, int pix, int c, int C, float gscale) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MSE_MAX_HEADS; ++k) {
    if (k < a.n_heads) {
      for (int j = 0; j < c; ++j) {
        const float d = __ldg(a.out[k] + pix * c + j) - __ldg(a.target + pix * C + k * c + j);
        s = fmaf(d, d, s);
        if (WITH_G) a.g[k][pix * c + j] = gscale * d;
      }
    }
  }
  return s;
}

__global__ void __launch_bounds__(1024) mse_heads_fwd_kernel(const MseHeadsArgs a, float* __restrict__ loss,
                                                             float* __restrict__ partial, unsigned int* counter) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  __shared__ bool last;
  const int C = a.n_heads * a.c, c = a.c, npix = (int)a.npix;
  const float total = (float)npix * (float)C;
  const bool with_g = a.g[0] != nullptr;          // also write d loss / d outs for grad_loss = 1 (the train step's case)
  const float gscale = 2.0f / total;
  float s = 0.f;
  const int stride = gridDim.x * blockDim.x;
  if (c == 1) {
    // one channel per head (every model of the reference): 8 pixels per pass, all loads before the first store -- the
    // gradient stores may alias the inputs as far as the compiler knows, so a store between two loads serialises them
    // (the straightforward loop ran 15 dependent round trips to L2: 19 us for 30 k elements)
    constexpr int U = 8;
    for (int pix0 = blockIdx.x * blockDim.x + threadIdx.x; pix0 < npix; pix0 += U * stride) {
      float d[U][MSE_MAX_HEADS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * stride;
#pragma unroll
        for (int k = 0; k < MSE_MAX_HEADS; ++k)
          d[u][k] = (pix < npix && k < a.n_heads) ? __ldg(a.out[k] + pix) - __ldg(a.target + pix * C + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * stride;
#pragma unroll
        for (int k = 0; k < MSE_MAX_HEADS; ++k) {
          s = fmaf(d[u][k], d[u][k], s);
          if (with_g && pix < npix && k < a.n_heads) a.g[k][pix] = gscale * d[u][k];
        }
      }
    }
  } else if (with_g) {
#pragma unroll 4
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += stride) s += mse_pixel<true>(a, pix, c, C, gscale);
  } else {
#pragma unroll 4
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += stride) s += mse_pixel<false>(a, pix, c, C, gscale);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      if (gridDim.x == 1) {
        *loss = t / total;
        last = false;
      } else {
        partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
      }
    }
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    float t = 0.f;                                 // fixed order: lane l adds partials l, l + 32, ...; then the warp tree
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += *reinterpret_cast<volatile float*>(partial + b);
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      *loss = t / total;
      *counter = 0u;            // ready for the next launch (graph replays included)
    }
  }
}

__global__ void mse_heads_bwd_kernel(const MseHeadsArgs a, const float* __restrict__ grad_loss) {
  pdl_launch_dependents();
  pdl_wait();
  const int C = a.n_heads * a.c, c = a.c, npix = (int)a.npix;
  const float scale = 2.0f / ((float)npix * (float)C) * __ldg(grad_loss);
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < MSE_MAX_HEADS; ++k)
      if (k < a.n_heads)
        for (int j = 0; j < c; ++j)
          a.g[k][pix * c + j] = scale * (__ldg(a.out[k] + pix * c + j) - __ldg(a.target + pix * C + k * c + j));
  }
}

int mse_heads_blocks(long npix) {
  static const long per_block = [] { const char* e = getenv("BDN_MSE_PIX_PER_BLOCK"); return e ? atol(e) : 1024L; }();   // (tuning knob)
  const long b = (npix + per_block - 1) / per_block;      // one pixel per thread: a single round of loads, then two-level sum
  return (int)(b < 1 ? 1 : (b > MSE_BLOCKS ? MSE_BLOCKS : b));
}

void launch_mse_heads(const MseHeadsArgs& a, float* loss, float* partial, unsigned int* counter, cudaStream_t st) {
  LaunchScope scope("mse_heads", st);
  launch_k(mse_heads_fwd_kernel, dim3(mse_heads_blocks(a.npix)), dim3(1024), 0, st, a, loss, partial, counter);
}

void launch_mse_heads_bwd(const MseHeadsArgs& a, const float* grad_loss, cudaStream_t st) {
  LaunchScope scope("mse_heads_bwd", st);
  const long b = (a.npix + 255) / 256;
  launch_k(mse_heads_bwd_kernel, dim3((unsigned)(b > 148 * 4 ? 148 * 4 : b)), dim3(256), 0, st, a, grad_loss);
}

// ===========================================================================
// Adam over a flat buffer (torch.optim.Adam defaults: no weight decay, no amsgrad)
// ===========================================================================
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float grad_scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

void launch_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                 int step, float grad_scale, cudaStream_t st) {
  LaunchScope scope("adam", st);
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const int block = 256;
  size_t blocks = (n + block - 1) / block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(adam_kernel, dim3((int)blocks), dim3(block), 0, st, p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
}

}  // namespace bdn
