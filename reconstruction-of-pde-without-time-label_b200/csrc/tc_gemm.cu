// W-forward pruned DFT on the 5th-generation tensor cores (tcgen05), TF32 operands, fp32 accumulation
// in tensor memory.  This is the BDN_PREC_TF32 mode of stage K1:
//
//     out[r, n] = sum_w x[r, w] * B[n, w]        r over (image, channel, h) rows, n = 2*l + {re, im}
//     B[2l, w] = cos(2 pi l w / Wp),  B[2l+1, w] = -sin(2 pi l w / Wp)      (only the kept modes l < m2)
//
// i.e. a [rows x Wp] x [Wp x 2*m2] GEMM whose B operand is the mode-limited DFT matrix.  It replaces
// torch.fft.rfft/rfft2's W pass (2d_FPE/FNOModules.py:163, 1d_FPE/FNOModules.py:50) for the kept bins.
//
// Structure (one persistent CTA per SM, 6 warps):
//   warp 0      TMA producer: per 128-row tile, ceil(Wp/32) boxes [32 floats x 128 rows] of x through a
//               CUtensorMap with 128-byte swizzle (out-of-range columns / rows are zero filled), two stages,
//               completion by transaction bytes on an mbarrier
//   warp 1      allocates tensor memory, issues tcgen05.mma.kind::tf32 (M = 128, N = 2*m2 padded to 16,
//               K = 8 per instruction) from one elected thread; tcgen05.commit releases the smem stage and
//               publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> 128-bit global stores;
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
// The DFT matrix is staged once per CTA (already in the swizzled K-major image the MMA descriptor expects).
#include "bdn_internal.cuh"

#include <cuda.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace bdn {

constexpr int TC_BM = 128;         // rows per tile = UMMA M
constexpr int TC_KC = 32;          // fp32 per 128-byte swizzle row
constexpr int TC_STAGES = 2;       // smem stages of the x tile
constexpr int TC_ATILE = TC_BM * 128;   // bytes of one [128 x 32] fp32 box

// ---------------------------------------------------------------------------
// host: tensor map (driver entry point fetched through the runtime, no -lcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      cudaGetLastError();
  });
  return fn;
}

// ---------------------------------------------------------------------------
// device helpers (PTX)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused with swizzle)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8 rows x 128 B         [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // layout: SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcWfwdParams {
  float* out;             // [rows][nreal]
  const float* b_image;   // [kch][n_pad][32] fp32, already in the 128-byte-swizzled smem image
                          // (split mode: the TF32-truncated high part, followed by the low part)
  int rows, nreal, n_pad, kch, ntiles;
  int act;                // exact GELU applied to the tile in shared memory before the MMAs
  int split;              // 3xTF32: x = hi + lo, B = hi + lo, D = lo*hi + hi*lo + hi*hi (fp32-level accuracy)
  uint32_t tmem_cols;     // power of two >= 2 * n_pad
};

// instruction descriptor: D = fp32, A = B = tf32, both K-major, N = n_pad, M = 128
__host__ __device__ inline uint32_t tc_idesc(int n_pad) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_pad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

constexpr int TC_XFORM_THREADS = 384;   // 12 transform warps (GELU / hi-lo split of a 48 KB tile per ~1 us)

__global__ void __launch_bounds__(192 + TC_XFORM_THREADS, 1) tc_wfwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const TcWfwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [A stages (hi | lo)][B image (hi | lo)][barriers][tmem slot]
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_tile_bytes = p.kch * TC_ATILE;                   // one copy of the x tile
  const int a_stage_bytes = a_tile_bytes * (p.split ? 2 : 1);  // split: low parts behind the high parts
  unsigned char* a_smem = base;
  unsigned char* b_smem = a_smem + TC_STAGES * a_stage_bytes;               // multiple of 1024
  const int b_part_bytes = p.kch * p.n_pad * 128;
  const int b_bytes = b_part_bytes * (p.split ? 2 : 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + ((b_bytes + 1023) & ~1023));
  uint64_t* full = bars;                 // [TC_STAGES]  TMA -> MMA
  uint64_t* empty = bars + TC_STAGES;    // [TC_STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * TC_STAGES;       // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * TC_STAGES + 2;  // [2] epilogue -> MMA
  uint64_t* bbar = bars + 2 * TC_STAGES + 4;    // B image landed
  uint64_t* xfull = bars + 2 * TC_STAGES + 5;   // [TC_STAGES] transform warps -> MMA (GELU / hi-lo split done)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 5);
  const bool transform = p.act || p.split;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&xfull[s], TC_XFORM_THREADS); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
    mbar_init(bbar, 1);
    mbar_init_fence();
    mbar_expect_tx(bbar, (uint32_t)b_bytes);
    bulk_g2s(b_smem, p.b_image, (uint32_t)b_bytes, bbar);
  }
  if (warp == 1) {   // one warp allocates tensor memory and later frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // barriers, tensor memory and the (constant) DFT operand are set up; x is the previous kernel's output

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)a_tile_bytes);
        for (int kc = 0; kc < p.kch; ++kc)
          tma_load_2d(a_smem + s * a_stage_bytes + kc * TC_ATILE, &tmap_x, &full[s], kc * TC_KC, tile * TC_BM);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = tc_idesc(p.n_pad);
      mbar_wait(bbar, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
        const int acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        mbar_wait(transform ? &xfull[s] : &full[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(a_smem + s * a_stage_bytes), b0 = smem_u32(b_smem);
        const uint32_t d = tmem_base + (uint32_t)(acc * p.n_pad);
        for (int kc = 0; kc < p.kch; ++kc)
#pragma unroll
          for (int k = 0; k < TC_KC / 8; ++k) {
            const uint32_t ao = a0 + kc * TC_ATILE + k * 32, bo = b0 + kc * p.n_pad * 128 + k * 32;
            if (p.split) {     // small cross terms first, then the main term
              umma_tf32(d, umma_desc_sw128(ao + a_tile_bytes), umma_desc_sw128(bo), idesc, (uint32_t)((kc | k) != 0));
              umma_tf32(d, umma_desc_sw128(ao), umma_desc_sw128(bo + b_part_bytes), idesc, 1u);
              umma_tf32(d, umma_desc_sw128(ao), umma_desc_sw128(bo), idesc, 1u);
            } else {
              umma_tf32(d, umma_desc_sw128(ao), umma_desc_sw128(bo), idesc, (uint32_t)((kc | k) != 0));
            }
          }
        tc_commit(&empty[s]);      // the x stage may be refilled once these MMAs have read it
        tc_commit(&tfull[acc]);    // ... and the accumulator is complete
      }
    }
  } else if (warp >= 6) {
    // ---------------- transform warps 6..17: exact GELU and / or the hi-lo split, in place in shared memory ----------------
    const int t = threadIdx.x - 192;
    const int n4 = a_tile_bytes >> 4;                      // float4 per tile copy
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
      mbar_wait(&full[s], ph);
      float4* hi = reinterpret_cast<float4*>(a_smem + s * a_stage_bytes);
      float4* lo = reinterpret_cast<float4*>(a_smem + s * a_stage_bytes + a_tile_bytes);
      for (int i = t; i < n4; i += TC_XFORM_THREADS) {
        float4 v = hi[i];
        if (p.act) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
        if (p.split) {
          float4 h;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          v = h;
        }
        hi[i] = v;
      }
      fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(&xfull[s]);
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quadrant = warp % 4 ----------------
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const int row = tile * TC_BM + quad * 32 + lane;
      float* orow = p.out + (size_t)row * p.nreal;
      for (int c0 = 0; c0 < p.n_pad; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.n_pad + c0), v);
        if (c0 + 32 >= p.n_pad) {       // last chunk read: the accumulator stage can be reused
          tc_fence_before();
          mbar_arrive(&tempty[acc]);
        }
        if (row < p.rows) {
          if ((p.nreal & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + 4 * q < p.nreal)
                *reinterpret_cast<float4*>(orow + c0 + 4 * q) =
                    make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                __uint_as_float(v[4 * q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c0 + q < p.nreal) orow[c0 + q] = __uint_as_float(v[q]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int tc_n_pad(int m2) { return ((2 * m2) + 15) & ~15; }
int tc_kch(int wp) { return (wp + TC_KC - 1) / TC_KC; }

// The B operand image for one plan: [kch][n_pad rows][128 bytes], 16-byte chunks XOR-swizzled by (row % 8).
void tc_build_b_image(int wp, int m2, std::vector<float>& img) {
  // [hi part | lo part]: hi = value truncated to TF32 (13 low mantissa bits cleared), lo = value - hi.
  // The plain TF32 mode reads only the first part (the tensor core ignores the low bits anyway).
  const int n_pad = tc_n_pad(m2), kch = tc_kch(wp);
  const size_t part = (size_t)kch * n_pad * 32;
  img.assign(2 * part, 0.f);
  const double two_pi = 6.283185307179586476925286766559;
  for (int n = 0; n < 2 * m2; ++n) {
    const int l = n >> 1;
    for (int w = 0; w < wp; ++w) {
      const double th = two_pi * (double)(((long long)l * w) % wp) / (double)wp;
      const float v = (n & 1) ? (float)(-std::sin(th)) : (float)std::cos(th);
      const int kc = w / TC_KC, k = w % TC_KC;
      const size_t off = (size_t)kc * n_pad * 32 + (size_t)n * 32 + (size_t)((((k >> 2) ^ (n & 7)) << 2) + (k & 3));
      uint32_t bits;
      memcpy(&bits, &v, 4);
      bits &= 0xFFFFE000u;
      float hi;
      memcpy(&hi, &bits, 4);
      img[off] = hi;
      img[part + off] = v - hi;
    }
  }
}

static size_t tc_smem_bytes(int wp, int m2, bool split) {
  const int kch = tc_kch(wp), n_pad = tc_n_pad(m2), f = split ? 2 : 1;
  return 1024 + (size_t)TC_STAGES * kch * TC_ATILE * f + (((size_t)kch * n_pad * 128 * f + 1023) & ~(size_t)1023) + 256;
}

bool tc_wfwd_supported(const Plan* pl, const float* x, bool split) {
  return pl->tc_fwd_b != nullptr && (pl->wp & 3) == 0 && tc_n_pad(pl->m2) <= 256 &&
         tc_smem_bytes(pl->wp, pl->m2, split) <= 224 * 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
         encode_fn() != nullptr;
}

// returns false if the tensor map could not be encoded (caller falls back to the fp32 kernel)
bool launch_wfwd_tc(const Plan* pl, const float* x, float2* out, int rows, int act, bool split, cudaStream_t st) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)pl->wp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pl->wp * sizeof(float)};
  const cuuint32_t box[2] = {TC_KC, TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;

  LaunchScope scope(split ? (act ? "wfwd_tc3x_gelu" : "wfwd_tc3x") : (act ? "wfwd_tc_gelu" : "wfwd_tc"), st, pl->m2);
  TcWfwdParams p;
  p.out = reinterpret_cast<float*>(out);
  p.b_image = pl->tc_fwd_b;
  p.rows = rows; p.nreal = 2 * pl->m2; p.n_pad = tc_n_pad(pl->m2); p.kch = tc_kch(pl->wp);
  p.ntiles = ceil_div(rows, TC_BM);
  p.act = act; p.split = split ? 1 : 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.n_pad)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = tc_smem_bytes(pl->wp, pl->m2, split);
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.ntiles < sms ? p.ntiles : sms;
  cudaFuncSetAttribute(tc_wfwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k(tc_wfwd_kernel, dim3(grid), dim3((act || split) ? 192 + TC_XFORM_THREADS : 192), smem, st, tmap, p);
  return true;
}

}  // namespace bdn
