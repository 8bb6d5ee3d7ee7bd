// One 2-D FNO spectral layer on the 5th-generation tensor cores (tcgen05), all four pruned-DFT GEMMs:
//
//   kernel P ("forward half")   z planes --[act, hi/lo split]--> S1: W-forward DFT  --> S2: H-forward DFT --> kept spectrum
//   kernel Q ("inverse half")   kept spectrum --per-mode channel mix--> S3: H-inverse DFT --> S4: W-inverse DFT
//                               --> epilogue (1x1 conv + bias, or the GELU' / 1x1-conv-gradient backward epilogue)
//
// Replaces (reference file:line) torch.fft.rfft2 / compl_mul2d / irfft2 of SpectralConv2d.forward
// (2d_FPE/FNOModules.py:141-178) and the layer body of FNO2d.forward (:226-232) for the kept modes only.
//
// Every GEMM is D[128 x N] (fp32, tensor memory) = A[128 x K] * B[N x K]^T with TF32 operands read from shared
// memory, issued by one thread (tcgen05.mma.cta_group::1.kind::tf32).  BDN_PREC_TF32X3 (the default parity mode)
// splits every operand into a TF32 high part and the TF32-rounded remainder and accumulates
// lo*hi + hi*lo + hi*hi in tensor memory: fp32-level accuracy (the 1e-5 bound); BDN_PREC_TF32 issues hi*hi only.
//
//   S1  D1[h, (l,re|im)]          = sum_w  act(z)[h, w]        * F1[(l,re|im), w]     per channel plane
//   S2  D2[(c,l,re|im), (k,c|s)]  = sum_h  X1[(c,l,re|im), h]  * F2[(k,c|s), h]       complex product finished by a
//                                                                                     lane-pair exchange (re <-> im rows)
//   mix Y[o,k,l] = sum_i X[i,k,l] W[i,o,k,l]   (CUDA cores, between the two halves: K = C is tiny)
//   S3  D3[(o,l,re|im), (h,c|s)]  = sum_k  Y[(o,l,re|im), k]   * F3[(h,c|s), k]
//   S4  D4[(o,h), w]              = sum_(l,re|im) Z[(o,h), (l,re|im)] * F4[w, (l,re|im)]
//
// Operands live in shared memory in the no-swizzle K-major canonical layout (8-row x 16-byte core matrices):
// element (r, k) at (k/4)*LBO + r*16 + (k%4)*4 bytes, LBO = rows*16 + 16 (the +16 staggers the K slabs over the
// banks so that both row-wise and column-wise writers are conflict free).  The constant DFT operands F1..F4 are
// built once per plan in fp64, rounded, split and stored in exactly that image, so one bulk async copy (UBLKCP)
// stages each of them.  The data operands are written by the CTA's threads (the "epilogue" of stage n is the
// operand producer of stage n+1): tcgen05.ld -> registers -> hi/lo split -> st.shared -> fence.proxy.async.
//
// One CTA per SM, 512 threads: every thread takes part in the CUDA-core phases, thread 0 issues the MMAs, stage
// completion is a tcgen05.commit on an mbarrier.  A work item is (image, group of cg channels); the many-image
// per-snapshot net (C = 4) takes whole images (cg = C), the few-image output heads one channel plane per CTA.
#include "bdn_internal.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace bdn {
namespace tcl {

constexpr int NT = 512;            // threads per CTA
constexpr int NW = NT / 32;        // 16 warps: quadrant = warp & 3 (the TMEM lanes a warp may read), group = warp >> 2
constexpr int NG = NW / 4;
constexpr int MAXB = 8;            // transform: register-prefetched float4 per warp per plane (hp, wp <= 128)

static inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int ns_lbo(int rows) { return rows * 16 + 16; }
// bytes of one (hi or lo) part of an operand tile, with slack for the rows an M = 128 instruction reads past `rows`
static inline uint32_t ns_part_bytes(int rows, int K) { return (uint32_t)((K / 4) * ns_lbo(rows) + 2048); }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ns_off(int r, int k, int lbo) { return (uint32_t)((k >> 2) * lbo + r * 16 + (k & 3) * 4); }
// v = hi + lo with hi the nearest TF32 value and lo the nearest TF32 value of the (exact) remainder: both parts
// are exact tensor-core operands, the representation error is 2^-24 |v| and unbiased (truncation, which is what the
// tensor core does to an fp32 operand on its own, would leave a one-sided 2^-22 |v|).
__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = rn_tf32(v);
  lo = rn_tf32(v - hi);
}
__device__ __forceinline__ void st_split(unsigned char* hi_base, uint32_t part, uint32_t off, float v) {
  float hi, lo;
  split_tf32(v, hi, lo);
  *reinterpret_cast<float*>(hi_base + off) = hi;
  *reinterpret_cast<float*>(hi_base + part + off) = lo;
}
// K-major, no swizzle: leading byte offset = distance of the two core matrices an instruction reads along K,
// stride byte offset = distance of consecutive 8-row groups (128 bytes here).
// (Checked on the B200 by exchanging the two fields: the other reading faults.)
__device__ __forceinline__ uint64_t ns_desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(128u >> 4) << 32;
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
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
// instruction descriptor: D = fp32, A = B = tf32, both K-major, M = 128
__host__ __device__ inline uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// D (+)= A * B^T, K = 8 per instruction; 3 passes (lo*hi, hi*lo, hi*hi) or hi*hi only.  One thread.
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_part, uint32_t lbo_a, uint32_t b_hi,
                                           uint32_t b_part, uint32_t lbo_b, int ksteps, int passes, uint32_t idesc) {
  uint32_t acc = 0;
  for (int pass = (passes == 3 ? 0 : 2); pass < 3; ++pass) {
    const uint32_t a = pass == 0 ? a_hi + a_part : a_hi;
    const uint32_t b = pass == 1 ? b_hi + b_part : b_hi;
    for (int ks = 0; ks < ksteps; ++ks) {
      umma_tf32(d_tmem, ns_desc(a + 2u * ks * lbo_a, lbo_a), ns_desc(b + 2u * ks * lbo_b, lbo_b), idesc, acc);
      acc = 1;
    }
  }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// ===========================================================================
// kernel P: planes -> kept spectrum
// ===========================================================================
struct PParams {
  const float* x;          // [nitems * cg planes][hp][wp]
  float* a_out;            // act(x) planes (the 1x1 conv input of kernel Q), or null
  float2* spec_out;        // [images, C, K, m2]
  const float* pre;        // [m2] column scale applied to the spectrum
  const float* f1; const float* f2;       // operand images (hi | lo) in global memory
  int nitems, C, cg, hp, wp, m2, K;       // K = 2 * m1 kept rows
  int act, passes;
  int N1, K1, N2, K2;
  int lboA1, lboF1, lboA2, lboF2;
  uint32_t partA1, partA2, partF1, partF2;
  uint32_t offF1, offF2, offA1[2], offA2, offBar;
  int nbuf;
  uint32_t tmem_cols, d2col;
};

__global__ void __launch_bounds__(NT, 1) p_kernel(const PParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);   // [0] tables, [1] stage done, [2], [3] A1 buffers free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_init_fence();
    mbar_expect_tx(&bars[0], 2u * p.partF1 + 2u * p.partF2);
    bulk_g2s(smem + p.offF1, p.f1, 2u * p.partF1, &bars[0]);
    bulk_g2s(smem + p.offF2, p.f2, 2u * p.partF2, &bars[0]);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t sbase = smem_u32(smem);

  const int hp = p.hp, wp = p.wp, m2 = p.m2, cg = p.cg;
  const int nq = wp >> 2, nslab1 = p.K1 >> 2;
  const int nkb = (nslab1 + 3) >> 2, nrb = (hp + 7) >> 3, nblk = nrb * nkb;
  const size_t plane_elems = (size_t)hp * wp;
  const int E = cg * m2;

  // the planes this CTA transforms, in order: n -> item blockIdx.x + (n / cg) * gridDim.x, plane n % cg of it
  const int my_items = (int)blockIdx.x < p.nitems ? (p.nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_planes = my_items * cg;
  auto plane_of = [&](int n) { return ((size_t)blockIdx.x + (size_t)(n / cg) * gridDim.x) * cg + (n % cg); };

  pdl_wait();
  float4 nxt[MAXB];
  auto prefetch = [&](int n) {
    const float* src = p.x + plane_of(n) * plane_elems;
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
      const int blk = warp + j * NW;
      nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blk < nblk) {
        const int rb = blk / nkb, kb = blk - rb * nkb;
        const int r = rb * 8 + (lane & 7), kq = kb * 4 + (lane >> 3);
        if (r < hp && kq < nq) nxt[j] = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * wp) + kq);
      }
    }
  };
  if (my_planes > 0) prefetch(0);
  mbar_wait(&bars[0], 0);      // tables landed

  uint32_t ph_main = 0, ph_buf[2] = {0, 0};
  const uint32_t idesc1 = idesc_n(p.N1), idesc2 = idesc_n(p.N2);
  for (int n = 0; n < my_planes; ++n) {
    const int pl = n % cg;
    const int buf = p.nbuf == 2 ? (n & 1) : 0;
    if (n >= p.nbuf) {          // the MMAs that read this buffer (plane n - nbuf) are complete
      mbar_wait(&bars[2 + buf], ph_buf[buf]);
      ph_buf[buf] ^= 1;
    }
    // ---- transform: act, hi/lo split, operand layout; side output act(x) ----
    unsigned char* a1 = smem + p.offA1[buf];
    float* aout = p.a_out ? p.a_out + plane_of(n) * plane_elems : nullptr;
    float4 cur[MAXB];
#pragma unroll
    for (int j = 0; j < MAXB; ++j) cur[j] = nxt[j];
    if (n + 1 < my_planes) prefetch(n + 1);
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
      const int blk = warp + j * NW;
      if (blk < nblk) {
        const int rb = blk / nkb, kb = blk - rb * nkb;
        const int r = rb * 8 + (lane & 7), kq = kb * 4 + (lane >> 3);
        if (kq < nslab1) {        // rows >= hp of the last row block are written too (zeros): finite garbage rows
          float4 v = cur[j];
          if (p.act) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
          if (aout != nullptr && r < hp && kq < nq) *reinterpret_cast<float4*>(aout + (size_t)r * wp + 4 * kq) = v;
          float4 hi, lo;
          split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
          const uint32_t off = (uint32_t)(kq * p.lboA1 + r * 16);
          *reinterpret_cast<float4*>(a1 + off) = hi;
          *reinterpret_cast<float4*>(a1 + p.partA1 + off) = lo;
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    const bool last = pl == cg - 1;
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base + (uint32_t)(pl * p.N1), sbase + p.offA1[buf], p.partA1, p.lboA1, sbase + p.offF1, p.partF1,
                 p.lboF1, p.K1 >> 3, p.passes, idesc1);
      tc_commit(&bars[2 + buf]);
      if (last) tc_commit(&bars[1]);
    }
    if (!last) continue;

    // ---- E1: D1 -> A2[(plane, l, re|im) rows][h columns] ----
    const size_t item = (size_t)blockIdx.x + (size_t)(n / cg) * gridDim.x;
    mbar_wait(&bars[1], ph_main);
    ph_main ^= 1;
    tc_fence_after();
    {
      unsigned char* a2 = smem + p.offA2;
      const int h = quad * 32 + lane;
      const int nch = p.N1 >> 4;
      for (int u = grp; u < cg * nch; u += NG) {
        const int upl = u / nch, ch = u - upl * nch;
        float v[16];
        tmem_ld16(lane_base + (uint32_t)(upl * p.N1 + ch * 16), v);
        if (h < p.K2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nn = ch * 16 + j;
            if (nn < 2 * m2) {
              const int e = upl * m2 + (nn >> 1);
              const int rho = ((e >> 4) << 5) + ((nn & 1) << 4) + (e & 15);
              st_split(a2, p.partA2, ns_off(rho, h, p.lboA2), h < hp ? v[j] : 0.f);
            }
          }
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base + p.d2col, sbase + p.offA2, p.partA2, p.lboA2, sbase + p.offF2, p.partF2, p.lboF2, p.K2 >> 3,
                 p.passes, idesc2);
      tc_commit(&bars[1]);
    }
    // ---- E2: finish the complex product (lane pair re <-> im), scale, store the kept spectrum ----
    mbar_wait(&bars[1], ph_main);
    ph_main ^= 1;
    tc_fence_after();
    {
      const int e = quad * 16 + (lane & 15), reim = lane >> 4;
      const bool valid = e < E;
      const int upl = valid ? e / m2 : 0, l = valid ? e - upl * m2 : 0;
      const float sc = valid ? __ldg(p.pre + l) : 0.f;
      const size_t plane = item * cg + upl;          // = img * C + channel
      float* dst = reinterpret_cast<float*>(p.spec_out + plane * p.K * m2 + l) + reim;
      const int nch = p.N2 >> 4;
      for (int ch = grp; ch < nch; ch += NG) {
        float v[16];
        tmem_ld16(lane_base + p.d2col + (uint32_t)(ch * 16), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ps = __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 16);
          const float val = reim ? v[2 * j] - ps : v[2 * j] + ps;
          const int k = ch * 8 + j;
          if (valid && k < p.K) dst[(size_t)k * m2 * 2] = val * sc;
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // D1 / D2 / A2 are free for the next item
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ===========================================================================
// kernel Q: kept spectrum -> planes
// ===========================================================================
struct QParams {
  const float2* xin;       // [images, C, K, m2]: X (forward) or GY (backward)
  const float2* w1; const float2* w2;    // [C, C, m1, m2] complex
  const float* a_in;       // forward: 1x1-conv input planes act(z_in); backward: gz_out planes
  const float* zin;        // backward: z_in planes (pre-activation)
  float* out;              // forward: z_out; backward: gz_in
  const float* pw_w; const float* pw_b;  // [C, C], [C]
  float* g_pw_w; float* g_pw_b;          // backward: accumulated
  const float* post;       // [m2]
  const float* f3; const float* f4;
  int nitems, C, cg, hp, wp, m1, m2, K, act_in, passes;
  int N3, K3, N4, K4, lboA3, lboF3, lboA4, lboF4;
  uint32_t partA3, partA4, partF3, partF4;
  uint32_t offF3, offF4, offA3, offA4, offPw, offBar;
  int HB, ntiles;
  uint32_t tmem_cols, d4col;
};

template <bool BWD>
__global__ void __launch_bounds__(NT, 1) q_kernel(const QParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);   // [0] tables, [1] stage done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* pws = reinterpret_cast<float*>(smem + p.offPw);           // [C*C] weights, [C] bias, then [C*C + C] gradient sums
  const int C = p.C, cg = p.cg, hp = p.hp, wp = p.wp, m1 = p.m1, m2 = p.m2, K = p.K;
  float* gacc = pws + C * C + C;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init_fence();
    mbar_expect_tx(&bars[0], 2u * p.partF3 + 2u * p.partF4);
    bulk_g2s(smem + p.offF3, p.f3, 2u * p.partF3, &bars[0]);
    bulk_g2s(smem + p.offF4, p.f4, 2u * p.partF4, &bars[0]);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  // K padding of the data operands (columns K..K3-1 of A3, 2*m2..K4-1 of A4) is zeroed once: the producers only
  // ever write valid columns.  Whole slabs from the first one that holds a padded column.
  {
    const int rows3 = (p.lboA3 - 16) >> 4, rows4 = (p.lboA4 - 16) >> 4;
    for (int part = 0; part < 2; ++part) {
      for (int s = K >> 2; s < (p.K3 >> 2); ++s)
        for (int i = tid; i < rows3 * 4; i += NT)
          reinterpret_cast<float*>(smem + p.offA3 + part * p.partA3 + s * p.lboA3)[i] = 0.f;
      for (int s = (2 * m2) >> 2; s < (p.K4 >> 2); ++s)
        for (int i = tid; i < rows4 * 4; i += NT)
          reinterpret_cast<float*>(smem + p.offA4 + part * p.partA4 + s * p.lboA4)[i] = 0.f;
    }
  }
  for (int i = tid; i < C * C; i += NT) pws[i] = __ldg(p.pw_w + i);
  for (int i = tid; i < C; i += NT) pws[C * C + i] = (!BWD && p.pw_b != nullptr) ? __ldg(p.pw_b + i) : 0.f;
  if (BWD)
    for (int i = tid; i < C * C + C; i += NT) gacc[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t sbase = smem_u32(smem);
  const int E = cg * m2;
  const size_t plane_elems = (size_t)hp * wp;
  const uint32_t idesc3 = idesc_n(p.N3), idesc4 = idesc_n(p.N4);
  unsigned char* a3 = smem + p.offA3;
  unsigned char* a4 = smem + p.offA4;
  uint32_t ph = 0;

  pdl_wait();
  mbar_wait(&bars[0], 0);

  for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    const int img = (item * cg) / C, o0 = (item * cg) - img * C;
    // ---- per-mode channel mix -> A3[(o, l, re|im) rows][k columns] ----
    {
      const size_t cstride = (size_t)m1 * m2;
      const float2* xb = p.xin + (size_t)img * C * K * m2;
      for (int u = tid; u < cg * K * m2; u += NT) {
        const int l = u % m2, k = (u / m2) % K, ol = u / (m2 * K);
        const bool lo = k < m1;
        const float2* wsel = lo ? p.w1 : p.w2;
        const size_t mode_off = (size_t)(lo ? k : k - m1) * m2 + l;
        const int o = o0 + ol;
        const float2* xp = xb + (size_t)k * m2 + l;
        float yr = 0.f, yi = 0.f;
#pragma unroll 4
        for (int a = 0; a < C; ++a) {
          const float2 x = __ldg(xp + (size_t)a * K * m2);
          if (!BWD) {
            const float2 w = __ldg(wsel + (size_t)(a * C + o) * cstride + mode_off);
            yr = fmaf(x.x, w.x, fmaf(-x.y, w.y, yr));
            yi = fmaf(x.x, w.y, fmaf(x.y, w.x, yi));
          } else {
            const float2 w = __ldg(wsel + (size_t)(o * C + a) * cstride + mode_off);
            yr = fmaf(x.x, w.x, fmaf(x.y, w.y, yr));
            yi = fmaf(x.y, w.x, fmaf(-x.x, w.y, yi));
          }
        }
        const int e = ol * m2 + l;
        const int rho = ((e >> 4) << 5) + (e & 15);
        st_split(a3, p.partA3, ns_off(rho, k, p.lboA3), yr);
        st_split(a3, p.partA3, ns_off(rho + 16, k, p.lboA3), yi);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base, sbase + p.offA3, p.partA3, p.lboA3, sbase + p.offF3, p.partF3, p.lboF3, p.K3 >> 3, p.passes,
                 idesc3);
      tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph);
    ph ^= 1;
    tc_fence_after();
    // ---- E3: finish the complex product, scale, -> A4[(o, h) rows][(l, re|im) columns] ----
    {
      const int e = quad * 16 + (lane & 15), reim = lane >> 4;
      const bool valid = e < E;
      const int ol = valid ? e / m2 : 0, l = valid ? e - ol * m2 : 0;
      const float sc = valid ? __ldg(p.post + l) : 0.f;
      const int col = 2 * l + reim;
      const int nch = p.N3 >> 4;
      for (int ch = grp; ch < nch; ch += NG) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)(ch * 16), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ps = __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 16);
          const float val = (reim ? v[2 * j] + ps : v[2 * j] - ps) * sc;
          const int h = ch * 8 + j;
          if (valid && h < hp) {
            const int t = h / p.HB;
            const int r4 = t * 128 + ol * p.HB + (h - t * p.HB);
            st_split(a4, p.partA4, ns_off(r4, col, p.lboA4), val);
          }
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int t = 0; t < p.ntiles; ++t)
        issue_gemm(tmem_base + p.d4col + (uint32_t)(t * p.N4), sbase + p.offA4 + (uint32_t)t * 2048u, p.partA4, p.lboA4,
                   sbase + p.offF4, p.partF4, p.lboF4, p.K4 >> 3, p.passes, idesc4);
      tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph);
    ph ^= 1;
    tc_fence_after();
    // ---- E4: epilogue on the rows (o, h) of every tile ----
    {
      const int r = quad * 32 + lane;
      const int ol = r / p.HB, hh = r - ol * p.HB;      // HB >= 32: ol is uniform over a warp
      const int nch = p.N4 >> 4;
      const int ch_own = o0 + ol;                        // the channel this row belongs to
      const bool row_ok = ol < cg;
      for (int u = grp; u < p.ntiles * nch; u += NG) {
        const int t = u / nch, ch = u - t * nch;
        const int h = t * p.HB + hh, w0 = ch * 16;
        float v[16];
        tmem_ld16(lane_base + p.d4col + (uint32_t)(t * p.N4 + ch * 16), v);
        const bool valid = row_ok && hh < p.HB && h < hp && w0 < wp;
        const int nv4 = valid ? (min(16, wp - w0) >> 2) : 0;
        const size_t pix = (size_t)h * wp + w0;
        if (!BWD) {
          if (valid) {
            const float bias = pws[C * C + ch_own];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += bias;
            const float* ab = p.a_in + (size_t)img * C * plane_elems + pix;
#pragma unroll 2
            for (int i = 0; i < C; ++i) {
              const float wv = pws[ch_own * C + i];
              const float4* ap = reinterpret_cast<const float4*>(ab + (size_t)i * plane_elems);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nv4) {
                  const float4 a = __ldg(ap + q);
                  v[4 * q] = fmaf(wv, a.x, v[4 * q]);
                  v[4 * q + 1] = fmaf(wv, a.y, v[4 * q + 1]);
                  v[4 * q + 2] = fmaf(wv, a.z, v[4 * q + 2]);
                  v[4 * q + 3] = fmaf(wv, a.w, v[4 * q + 3]);
                }
            }
            float4* op = reinterpret_cast<float4*>(p.out + ((size_t)img * C + ch_own) * plane_elems + pix);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nv4) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else {
          // gz_in[i] = (winv + sum_o W[o][i] g[o]) * act'(z_in[i]);  gW[o][i] += g[o] . act(z_in[i]);  gb[i] += sum g[i]
          float a[16], dact[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[j] = 0.f; dact[j] = 0.f; }
          if (valid) {
            const float4* zp = reinterpret_cast<const float4*>(p.zin + ((size_t)img * C + ch_own) * plane_elems + pix);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nv4) {
                const float4 z4 = __ldg(zp + q);
                const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  if (p.act_in) {
                    float cdf, pdf;
                    gelu_cdf_pdf(zz[jj], cdf, pdf);
                    a[4 * q + jj] = zz[jj] * cdf;
                    dact[4 * q + jj] = fmaf(zz[jj], pdf, cdf);
                  } else {
                    a[4 * q + jj] = zz[jj];
                    dact[4 * q + jj] = 1.0f;
                  }
                }
              }
          }
          const float* gb = p.a_in + (size_t)img * C * plane_elems + pix;
          for (int o = 0; o < C; ++o) {
            float dot = 0.f, gsum = 0.f;
            if (valid) {
              const float wv = pws[o * C + ch_own];
              const float4* gp = reinterpret_cast<const float4*>(gb + (size_t)o * plane_elems);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nv4) {
                  const float4 g = __ldg(gp + q);
                  v[4 * q] = fmaf(wv, g.x, v[4 * q]);
                  v[4 * q + 1] = fmaf(wv, g.y, v[4 * q + 1]);
                  v[4 * q + 2] = fmaf(wv, g.z, v[4 * q + 2]);
                  v[4 * q + 3] = fmaf(wv, g.w, v[4 * q + 3]);
                  dot = fmaf(g.x, a[4 * q], fmaf(g.y, a[4 * q + 1], fmaf(g.z, a[4 * q + 2], fmaf(g.w, a[4 * q + 3], dot))));
                  gsum += (g.x + g.y) + (g.z + g.w);
                }
            }
            dot = warp_sum(dot);
            if (lane == 0 && row_ok) atomicAdd(gacc + o * C + ch_own, dot);
            if (o == ch_own) {      // warp-uniform (ch_own is)
              gsum = warp_sum(gsum);
              if (lane == 0 && row_ok) atomicAdd(gacc + C * C + ch_own, gsum);
            }
          }
          if (valid) {
            float4* op = reinterpret_cast<float4*>(p.out + ((size_t)img * C + ch_own) * plane_elems + pix);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nv4)
                op[q] = make_float4(v[4 * q] * dact[4 * q], v[4 * q + 1] * dact[4 * q + 1], v[4 * q + 2] * dact[4 * q + 2],
                                    v[4 * q + 3] * dact[4 * q + 3]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // D3 / D4 / A3 / A4 are free for the next item
  }

  if (BWD) {
    __syncthreads();
    for (int i = tid; i < C * C + C; i += NT) {
      const float s = gacc[i];
      if (s != 0.f) atomicAdd(i < C * C ? p.g_pw_w + i : p.g_pw_b + (i - C * C), s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ===========================================================================
// host: operand images
// ===========================================================================
static float host_rn_tf32(float v) {      // nearest TF32 value, ties away from zero (cvt.rna.tf32.f32)
  uint32_t bits;
  memcpy(&bits, &v, 4);
  bits = (bits + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &bits, 4);
  return r;
}

template <typename F>
static void build_image(std::vector<float>& img, int rows, int K, F&& value) {
  const int lbo = ns_lbo(rows);
  const size_t part = ns_part_bytes(rows, K) / 4;     // floats
  img.assign(2 * part, 0.f);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const float v = (float)value(r, k);
      const float hi = host_rn_tf32(v);
      const size_t off = ((size_t)(k >> 2) * lbo + (size_t)r * 16 + (size_t)(k & 3) * 4) / 4;
      img[off] = hi;
      img[part + off] = host_rn_tf32(v - hi);
    }
}

static float* upload_f(const std::vector<float>& v) {
  float* d = nullptr;
  if (cudaMalloc(&d, v.size() * sizeof(float)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  return d;
}

}  // namespace tcl

using namespace tcl;

// Built by get_plan for 2-D shapes that can ever fit (hp, wp <= 128); a failed upload only disables the path.
void tcl_build_tables(Plan* pl) {
  pl->tcl_f1 = pl->tcl_f2 = pl->tcl_f3 = pl->tcl_f4 = nullptr;
  if (pl->ndim != 2 || pl->hp > 128 || pl->wp > 128 || (pl->wp & 3) != 0) return;
  const int hp = pl->hp, wp = pl->wp, m1 = pl->m1, m2 = pl->m2, K = pl->K;
  const double two_pi = 6.283185307179586476925286766559;
  auto theta = [&](int l, int w) { return two_pi * (double)(((long long)l * w) % wp) / (double)wp; };
  auto phi = [&](int k, int h) {
    const int kk = k < m1 ? k : hp - 2 * m1 + k;
    return two_pi * (double)(((long long)kk * h) % hp) / (double)hp;
  };
  const int N1 = pad_to(2 * m2, 16), K1 = pad_to(wp, 8);
  const int N2 = pad_to(2 * K, 16), K2 = pad_to(hp, 8);
  const int N3 = pad_to(2 * hp, 16), K3 = pad_to(K, 8);
  const int N4 = pad_to(wp, 16), K4 = pad_to(2 * m2, 8);
  std::vector<float> img;
  build_image(img, N1, K1, [&](int n, int w) {
    if (n >= 2 * m2 || w >= wp) return 0.0;
    return (n & 1) ? -std::sin(theta(n >> 1, w)) : std::cos(theta(n >> 1, w));
  });
  pl->tcl_f1 = upload_f(img);
  build_image(img, N2, K2, [&](int n, int h) {
    if (n >= 2 * K || h >= hp) return 0.0;
    return (n & 1) ? std::sin(phi(n >> 1, h)) : std::cos(phi(n >> 1, h));
  });
  pl->tcl_f2 = upload_f(img);
  build_image(img, N3, K3, [&](int n, int k) {
    if (n >= 2 * hp || k >= K) return 0.0;
    return (n & 1) ? std::sin(phi(k, n >> 1)) : std::cos(phi(k, n >> 1));
  });
  pl->tcl_f3 = upload_f(img);
  build_image(img, N4, K4, [&](int w, int c) {
    if (w >= wp || c >= 2 * m2) return 0.0;
    return (c & 1) ? -std::sin(theta(c >> 1, w)) : std::cos(theta(c >> 1, w));
  });
  pl->tcl_f4 = upload_f(img);
  if (!pl->tcl_f1 || !pl->tcl_f2 || !pl->tcl_f3 || !pl->tcl_f4) {
    cudaGetLastError();
    pl->tcl_f1 = pl->tcl_f2 = pl->tcl_f3 = pl->tcl_f4 = nullptr;
  }
}

// ===========================================================================
// host: shared-memory plans and launchers
// ===========================================================================
static const size_t TCL_SMEM_MAX = 227 * 1024;

static int sm_count() {
  static const int n = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
  }();
  return n;
}

// channels per work item: as many as fit the 128 accumulator rows of S2 / S3, fewer when there are too few items to
// give every SM one
static int pick_cg(int images, int C, int m2) {
  int cg = 4;
  while (cg > 1 && (C % cg != 0 || cg * 2 * m2 > 128)) cg >>= 1;
  if (cg * 2 * m2 > 128) return 0;
  while (cg > 1 && (long)images * C / cg < sm_count()) cg >>= 1;
  return cg;
}

static uint32_t pow2_cols(int need) {
  uint32_t c = 32;
  while ((int)c < need) c <<= 1;
  return c;
}

static bool plan_p(const Plan* pl, int images, int C, int passes, PParams& p, size_t& smem) {
  if (!pl->tcl_f1) return false;
  const int cg = pick_cg(images, C, pl->m2);
  if (cg == 0) return false;
  p.C = C; p.cg = cg; p.hp = pl->hp; p.wp = pl->wp; p.m2 = pl->m2; p.K = pl->K;
  p.nitems = images * C / cg;
  p.N1 = pad_to(2 * pl->m2, 16); p.K1 = pad_to(pl->wp, 8);
  p.N2 = pad_to(2 * pl->K, 16); p.K2 = pad_to(pl->hp, 8);
  if (p.N2 > 256 || cg * p.N1 + p.N2 > 512 || p.K2 > 128) return false;
  const int nrb = (pl->hp + 7) / 8, nkb = (p.K1 / 4 + 3) / 4;
  if (nrb * nkb > NW * MAXB) return false;
  const int rowsA1 = pad_to(pl->hp, 8), rowsA2 = pad_to(((cg * pl->m2 + 15) / 16) * 32, 8);
  p.lboA1 = ns_lbo(rowsA1); p.lboF1 = ns_lbo(p.N1); p.lboA2 = ns_lbo(rowsA2); p.lboF2 = ns_lbo(p.N2);
  p.partA1 = ns_part_bytes(rowsA1, p.K1); p.partF1 = ns_part_bytes(p.N1, p.K1);
  p.partA2 = ns_part_bytes(rowsA2, p.K2); p.partF2 = ns_part_bytes(p.N2, p.K2);
  p.tmem_cols = pow2_cols(cg * p.N1 + p.N2);
  p.d2col = (uint32_t)(cg * p.N1);
  p.passes = passes;
  // layouts tried in order: two A1 buffers + own A2; one A1 buffer + own A2; A2 overlaid on the single A1 buffer
  for (int variant = 0; variant < 3; ++variant) {
    const int nbuf = variant == 0 ? 2 : 1;
    const bool overlay = variant == 2;
    uint32_t off = 0;
    p.offF1 = off; off += 2 * p.partF1;
    p.offF2 = off; off += 2 * p.partF2;
    p.offA1[0] = off;
    const uint32_t a1 = 2 * p.partA1, a2 = 2 * p.partA2;
    if (overlay) {
      p.offA2 = off; p.offA1[1] = off;
      off += a1 > a2 ? a1 : a2;
    } else {
      off += a1;
      p.offA1[1] = nbuf == 2 ? off : p.offA1[0];
      if (nbuf == 2) off += a1;
      p.offA2 = off; off += a2;
    }
    p.offBar = off; off += 64;
    p.nbuf = nbuf;
    if (off <= TCL_SMEM_MAX) { smem = off; return true; }
  }
  return false;
}

static bool plan_q(const Plan* pl, int images, int C, int passes, QParams& p, size_t& smem) {
  if (!pl->tcl_f3) return false;
  const int cg = pick_cg(images, C, pl->m2);
  if (cg == 0) return false;
  p.C = C; p.cg = cg; p.hp = pl->hp; p.wp = pl->wp; p.m1 = pl->m1; p.m2 = pl->m2; p.K = pl->K;
  p.nitems = images * C / cg;
  p.N3 = pad_to(2 * pl->hp, 16); p.K3 = pad_to(pl->K, 8);
  p.N4 = pad_to(pl->wp, 16); p.K4 = pad_to(2 * pl->m2, 8);
  p.HB = 128 / cg;
  p.ntiles = (pl->hp + p.HB - 1) / p.HB;
  if (p.N3 > 256 || p.N3 + p.ntiles * p.N4 > 512) return false;
  const int rowsA3 = pad_to(((cg * pl->m2 + 15) / 16) * 32, 8);
  const int last_rows = pl->hp - (p.ntiles - 1) * p.HB;                       // h rows in the last tile
  const int rowsA4 = pad_to((p.ntiles - 1) * 128 + (cg - 1) * p.HB + last_rows, 8);
  p.lboA3 = ns_lbo(rowsA3); p.lboF3 = ns_lbo(p.N3); p.lboA4 = ns_lbo(rowsA4); p.lboF4 = ns_lbo(p.N4);
  p.partA3 = ns_part_bytes(rowsA3, p.K3); p.partF3 = ns_part_bytes(p.N3, p.K3);
  p.partA4 = ns_part_bytes(rowsA4, p.K4) + (uint32_t)(p.ntiles - 1) * 2048u;
  p.partF4 = ns_part_bytes(p.N4, p.K4);
  p.tmem_cols = pow2_cols(p.N3 + p.ntiles * p.N4);
  p.d4col = (uint32_t)p.N3;
  p.passes = passes;
  uint32_t off = 0;
  p.offF3 = off; off += 2 * p.partF3;
  p.offF4 = off; off += 2 * p.partF4;
  p.offA3 = off; off += 2 * p.partA3;
  p.offA4 = off; off += 2 * p.partA4;
  p.offPw = off; off += (uint32_t)pad_to((2 * (C * C + C)) * 4, 16);
  p.offBar = off; off += 64;
  smem = off;
  return off <= TCL_SMEM_MAX;
}

bool tcl_supported(const Plan* pl, int images, int C) {
  if (pl == nullptr || pl->ndim != 2 || images < 1) return false;
  PParams pp;
  QParams qp;
  size_t s1 = 0, s2 = 0;
  return plan_p(pl, images, C, 3, pp, s1) && plan_q(pl, images, C, 3, qp, s2);
}

bool launch_tcl_p(const Plan* pl, const float* x, float* a_out, float2* spec_out, const float* pre, int images, int C,
                  int act, int prec, cudaStream_t st) {
  PParams p;
  size_t smem = 0;
  if (!plan_p(pl, images, C, prec == 2 ? 3 : 1, p, smem)) return false;
  LaunchScope scope(act ? "tc_p_gelu" : "tc_p", st, C);
  p.x = x; p.a_out = a_out; p.spec_out = spec_out; p.pre = pre;
  p.f1 = pl->tcl_f1; p.f2 = pl->tcl_f2;
  p.act = act;
  const int grid = p.nitems < sm_count() ? p.nitems : sm_count();
  cudaFuncSetAttribute(p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCL_SMEM_MAX);
  launch_k(p_kernel, dim3(grid), dim3(NT), smem, st, p);
  return true;
}

bool launch_tcl_q(const Plan* pl, bool bwd, const float2* xin, const float2* w1, const float2* w2, const float* a_in,
                  const float* zin, float* out, const float* pw_w, const float* pw_b, float* g_pw_w, float* g_pw_b,
                  const float* post, int images, int C, int act_in, int prec, cudaStream_t st) {
  QParams p;
  size_t smem = 0;
  if (!plan_q(pl, images, C, prec == 2 ? 3 : 1, p, smem)) return false;
  LaunchScope scope(bwd ? "tc_q_bwd" : "tc_q_fwd", st, C);
  p.xin = xin; p.w1 = w1; p.w2 = w2; p.a_in = a_in; p.zin = zin; p.out = out;
  p.pw_w = pw_w; p.pw_b = pw_b; p.g_pw_w = g_pw_w; p.g_pw_b = g_pw_b; p.post = post;
  p.f3 = pl->tcl_f3; p.f4 = pl->tcl_f4;
  p.act_in = act_in;
  const int grid = p.nitems < sm_count() ? p.nitems : sm_count();
  if (bwd) {
    cudaFuncSetAttribute(q_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCL_SMEM_MAX);
    launch_k(q_kernel<true>, dim3(grid), dim3(NT), smem, st, p);
  } else {
    cudaFuncSetAttribute(q_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCL_SMEM_MAX);
    launch_k(q_kernel<false>, dim3(grid), dim3(NT), smem, st, p);
  }
  return true;
}

}  // namespace bdn
