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
// memory (tcgen05.mma.cta_group::1.kind::tf32).  BDN_PREC_TF32X3 (the default parity mode) splits every operand
// into its nearest TF32 value and the TF32-rounded remainder and accumulates lo*hi + hi*lo + hi*hi in tensor
// memory: fp32-level accuracy (the 1e-5 bound); BDN_PREC_TF32 issues hi*hi only.  The orientation of every stage
// is chosen so that the thread that owns an accumulator row (a TMEM lane) writes a CONTIGUOUS run of the next
// operand's K axis (128-bit shared stores):
//
//   S1  D1[(l,re|im), h]            = sum_w  F1[(l,re|im), w] * act(z)[h, w]            per channel plane
//   S2  D2[(c,l,re|im), (k,c|s)]    = sum_h  X1[(c,l,re|im), h] * F2[(k,c|s), h]        complex product finished by a
//                                                                                       lane-pair exchange (re <-> im rows)
//   mix Y[o,k,l] = sum_i X[i,k,l] W[i,o,k,l]   (CUDA cores, between the two halves: the contraction length is C)
//   S3  D3[(h,c|s), (o,l,re|im)]    = sum_k  F3[(h,c|s), k] * Y[(o,l,re|im), k]         lane-pair exchange (cos <-> sin rows)
//   S4  D4[(o,h), w]                = sum_(re|im,l) Z[(o,h), (re|im,l)] * F4[w, (re|im,l)]
//
// Operands live in shared memory in the no-swizzle K-major canonical layout (8-row x 16-byte core matrices):
// element (r, k) at (k/4)*LBO + r*16 + (k%4)*4 bytes, LBO = rows*16 + 16 (the +16 staggers the K slabs over the
// banks).  The constant DFT operands F1..F4 are built once per plan in fp64, rounded, split and stored in exactly
// that image, so one bulk async copy (UBLKCP) stages each of them.  The data operands are written by the CTA's
// worker threads: the "epilogue" of stage n is the operand producer of stage n+1 (tcgen05.ld -> registers ->
// hi/lo split -> st.shared.v4 -> fence.proxy.async).
//
// One CTA per SM: 12 worker warps do the CUDA-core phases, a 13th warp issues the MMAs.  Workers hand a finished
// operand to the issuer with a named barrier (bar.arrive, they do not wait), the issuer signals completion with
// tcgen05.commit on an mbarrier.  A work item is (image, group of cg channels): the many-image per-snapshot net
// (C = 4) takes whole images (cg = C), the few-image output heads one channel plane per CTA.
#include "bdn_internal.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace bdn {
namespace tcl {

constexpr int NWK = 384;           // worker threads
constexpr int NW = NWK / 32;       // 12 worker warps: quadrant = warp & 3 (the TMEM lanes a warp may read), group = warp >> 2
constexpr int NG = NW / 4;
constexpr int NB = NWK + 32;       // named-barrier participants: the workers + the MMA issuer warp (warp NW)
constexpr int NT = 512;            // launched threads: warps are allocated in fours, and 16 warps leave 128 registers
                                   // per thread (17 would be costed as 20: 96); warps 13..15 only hold the CTA open
constexpr int CB = 4;              // channels per batch of global loads in the mix / the epilogue

static inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int ns_lbo(int rows) { return rows * 16 + 16; }
// bytes of one (hi or lo) part of an operand: `tiles` row tiles of 128 are read by M = 128 instructions, the last one
// possibly past `rows`
static inline uint32_t ns_part_bytes(int rows, int K, int tiles = 1) {
  return (uint32_t)((K / 4) * ns_lbo(rows) + tiles * 2048);
}

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
__device__ __forceinline__ void st_split4(unsigned char* hi_base, uint32_t part, uint32_t off, float a, float b, float c,
                                          float d) {
  float4 hi, lo;
  split_tf32(a, hi.x, lo.x); split_tf32(b, hi.y, lo.y); split_tf32(c, hi.z, lo.z); split_tf32(d, hi.w, lo.w);
  *reinterpret_cast<float4*>(hi_base + off) = hi;
  *reinterpret_cast<float4*>(hi_base + part + off) = lo;
}
__device__ __forceinline__ void st_split1(unsigned char* hi_base, uint32_t part, uint32_t off, float v) {
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

// N (4 or 8) consecutive columns of this thread's TMEM lane, without the wait (tmem_ld_wait before the values are used)
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]) {
  static_assert(N == 4 || N == 8, "4 or 8 columns");
  uint32_t r[N];
  if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  }
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D (+)= A * B^T, K = 8 per instruction; 3 passes (lo*hi, hi*lo, hi*hi) or hi*hi only.  One thread.
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_part, uint32_t lbo_a, uint32_t b_hi,
                                           uint32_t b_part, uint32_t lbo_b, int ksteps, int passes, uint32_t idesc) {
  uint32_t acc = 0;
  const uint32_t astep = (2u * lbo_a) >> 4, bstep = (2u * lbo_b) >> 4;     // descriptor address units of 16 bytes
  for (int pass = (passes == 3 ? 0 : 2); pass < 3; ++pass) {
    uint64_t ad = ns_desc(pass == 0 ? a_hi + a_part : a_hi, lbo_a);
    uint64_t bd = ns_desc(pass == 1 ? b_hi + b_part : b_hi, lbo_b);
    for (int ks = 0; ks < ksteps; ++ks) {
      umma_tf32(d_tmem, ad, bd, idesc, acc);
      ad += astep;             // the 14-bit address field cannot carry: shared memory is < 256 KB
      bd += bstep;
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
// named barriers: workers arrive (and go on), the issuer warp waits
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(NB) : "memory"); }
__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NB) : "memory"); }

#define TCL_STAMP(i)                                                                                       \
  do {                                                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) p.dbg[(i)] = clock64();       \
  } while (0)
#define TCL_ISTAMP(i)      /* the issuer's lane 0 */                                                        \
  do {                                                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == NWK && (i) < 64) p.dbg[(i)] = clock64();     \
  } while (0)

// A from tensor memory ("TS"): the constant DFT operand sits in TMEM (128 lanes x K columns, hi part and lo part),
// so an instruction only fetches its B rows from shared memory.  With both operands in shared memory every K = 8
// step re-reads 128 rows x 32 bytes of A: ~100 cycles per instruction measured on the B200 whatever N is.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with the three passes kept in three accumulators `dstride` columns apart and interleaved per K step
// (three independent accumulation chains in flight; the consumer adds them up).
__device__ __forceinline__ void issue_gemm_split(uint32_t d_tmem, uint32_t dstride, uint32_t a_hi, uint32_t a_part,
                                                 uint32_t lbo_a, uint32_t b_hi, uint32_t b_part, uint32_t lbo_b, int ksteps,
                                                 uint32_t idesc) {
  const uint32_t astep = (2u * lbo_a) >> 4, bstep = (2u * lbo_b) >> 4;
  uint64_t ah = ns_desc(a_hi, lbo_a), al = ns_desc(a_hi + a_part, lbo_a);
  uint64_t bh = ns_desc(b_hi, lbo_b), bl = ns_desc(b_hi + b_part, lbo_b);
  uint32_t acc = 0;
  for (int ks = 0; ks < ksteps; ++ks) {
    umma_tf32(d_tmem, al, bh, idesc, acc);
    umma_tf32(d_tmem + dstride, ah, bl, idesc, acc);
    umma_tf32(d_tmem + 2u * dstride, ah, bh, idesc, acc);
    ah += astep; al += astep; bh += bstep; bl += bstep;
    acc = 1;
  }
}
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_part,
                                              uint32_t lbo_b, int ksteps, int passes, uint32_t idesc) {
  uint32_t acc = 0;
  const uint32_t bstep = (2u * lbo_b) >> 4;
  for (int pass = (passes == 3 ? 0 : 2); pass < 3; ++pass) {
    uint32_t a = pass == 0 ? a_lo : a_hi;
    uint64_t bd = ns_desc(pass == 1 ? b_hi + b_part : b_hi, lbo_b);
    for (int ks = 0; ks < ksteps; ++ks) {
      umma_tf32_ts(d_tmem, a, bd, idesc, acc);
      a += 8;                  // 8 columns = the K extent of one instruction
      bd += bstep;
      acc = 1;
    }
  }
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A constant operand into tensor memory: tab = [parts = 2 (hi, lo)][tiles][128 rows][K] fp32 in global memory, landed
// at columns col0 + (part * tiles + tile) * K.  Every worker warp writes the lanes of its quadrant.
__device__ __forceinline__ void load_table_tmem(uint32_t tmem_base, uint32_t col0, const float* tab, int tiles, int K,
                                                int quad, int grp, int lane) {
  const int nch = (K + 15) >> 4;          // K is a multiple of 8: the last chunk is 16 or 8 columns wide
  for (int pt = 0; pt < 2 * tiles; ++pt) {
    const float* row = tab + ((size_t)pt * 128 + quad * 32 + lane) * K;
    for (int ch = grp; ch < nch; ch += NG) {
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch * 16 + 4 * q < K) t = __ldg(reinterpret_cast<const float4*>(row + ch * 16) + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
      const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + col0 + (uint32_t)(pt * K + ch * 16);
      if (ch * 16 + 16 <= K) tmem_st16(ta, v);
      else tmem_st8(ta, v);
    }
  }
  tmem_st_wait();
}

// ===========================================================================
// kernel P: planes -> kept spectrum
// ===========================================================================
struct PParams {
  const float* x;          // [nitems * cg planes][hp][wp]
  float* a_out;            // act(x) planes (the 1x1 conv input of kernel Q), or null
  float2* spec_out;        // [images, C, K, m2]
  const float* pre;        // [m2] column scale applied to the spectrum
  const float* f1; const float* f2;       // shared-memory operand images (hi | lo); f1 only without TS
  const float* t1;         // F1 for tensor memory: [2][128][K1]
  int nitems, C, cg, hp, wp, m2, K;       // K = 2 * m1 kept rows
  int act, passes;
  int N1, K1, N2, K2;      // S1: N1 = hp padded to 16 columns, K1 = wp padded to 8; S2: N2 = 2K padded, K2 = hp padded to 8
  int lboA1, lboF1, lboA2, lboF2;
  uint32_t partA1, partA2, partF1, partF2;
  uint32_t offF1, offF2, offA1[2], offA2, offBar;
  int nbuf;
  int ts1, rep1;           // F1 in tensor memory; copies of F1's rows over the 128 lanes (4, 2 or 1)
  int split1;              // S1's three passes in three accumulators (cg * N1 columns apart), summed by E1
  uint32_t tmem_cols, colF1, colD;
  long long* dbg;          // phase time stamps of block 0 (diagnostic, BDN_TCL_DBG=1), or null
};

template <int MAXB>   // float4 per worker thread per plane
__global__ void __launch_bounds__(NT, 1) p_kernel(const PParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);   // [0] tables, [1] stage done, [2], [3] A1 buffers free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  pdl_launch_dependents();
  if (tid == NWK) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_init_fence();
    mbar_expect_tx(&bars[0], (p.ts1 ? 0u : 2u * p.partF1) + 2u * p.partF2);
    if (!p.ts1) bulk_g2s(smem + p.offF1, p.f1, 2u * p.partF1, &bars[0]);
    bulk_g2s(smem + p.offF2, p.f2, 2u * p.partF2, &bars[0]);
  }
  if (warp == NW) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int hp = p.hp, wp = p.wp, m2 = p.m2, cg = p.cg;
  if (p.ts1 && warp < NW) load_table_tmem(tmem_base, p.colF1, p.t1, 1, p.K1, warp & 3, warp >> 2, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // the planes this CTA transforms, in order: n -> item blockIdx.x + (n / cg) * gridDim.x, plane n % cg of it
  const int my_items = (int)blockIdx.x < p.nitems ? (p.nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_planes = my_items * cg;
  pdl_wait();

  if (warp == NW) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc1 = idesc_n(p.N1), idesc2 = idesc_n(p.N2);
    if (lane == 0) mbar_wait(&bars[0], 0);      // tables landed
    __syncwarp();
    for (int n = 0; n < my_planes; ++n) {
      const int pl = n % cg, buf = p.nbuf == 2 ? (n & 1) : 0;
      nb_sync(2 + (n & 1));                     // A1[buf] of plane n is written and fenced
      tc_fence_after();
      if (n < 4) TCL_ISTAMP(32 + 2 * n);
      if (lane == 0) {
        const uint32_t d = tmem_base + p.colD + (uint32_t)(pl * p.N1);
        if (p.split1)
          issue_gemm_split(d, (uint32_t)(cg * p.N1), sbase + p.offF1, p.partF1, p.lboF1, sbase + p.offA1[buf], p.partA1, p.lboA1,
                           p.K1 >> 3, idesc1);
        else if (p.ts1)
          issue_gemm_ts(d, tmem_base + p.colF1, tmem_base + p.colF1 + (uint32_t)p.K1, sbase + p.offA1[buf], p.partA1, p.lboA1,
                        p.K1 >> 3, p.passes, idesc1);
        else
          issue_gemm(d, sbase + p.offF1, p.partF1, p.lboF1, sbase + p.offA1[buf], p.partA1, p.lboA1, p.K1 >> 3, p.passes,
                     idesc1);
        tc_commit(&bars[2 + buf]);
        if (pl == cg - 1) tc_commit(&bars[1]);
      }
      if (n < 4) TCL_ISTAMP(33 + 2 * n);
      __syncwarp();
      if (pl == cg - 1) {
        nb_sync(4);                             // A2 of this item is written and fenced
        tc_fence_after();
        if (n < cg) TCL_ISTAMP(40);
        if (lane == 0) {
          issue_gemm(tmem_base + p.colD, sbase + p.offA2, p.partA2, p.lboA2, sbase + p.offF2, p.partF2, p.lboF2, p.K2 >> 3,
                     p.passes, idesc2);
          tc_commit(&bars[1]);
        }
        if (n < cg) TCL_ISTAMP(41);
        __syncwarp();
      }
    }
  } else if (warp < NW) {
    // ------------------------------- workers -------------------------------
    const int quad = warp & 3, grp = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + p.colD;
    const int nq = wp >> 2, nslab1 = p.K1 >> 2;
    const int nkb = (nslab1 + 3) >> 2, nrb = (p.N1 + 7) >> 3, nblk = nrb * nkb;
    const size_t plane_elems = (size_t)hp * wp;
    const int E = cg * m2;
    auto plane_of = [&](int n) { return ((size_t)blockIdx.x + (size_t)(n / cg) * gridDim.x) * cg + (n % cg); };
    // a thread's float4s of a plane (same for every plane): 8 rows x 4 K-slabs per warp instruction, which is
    // conflict free for the shared stores and reads 64-byte runs of global memory
    int goff[MAXB], soff[MAXB];
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
      const int blk = warp + j * NW;
      goff[j] = -1; soff[j] = -1;
      if (blk < nblk) {
        const int rb = blk / nkb, kb = blk - rb * nkb;
        const int r = rb * 8 + (lane & 7), kq = kb * 4 + (lane >> 3);
        if (kq < nslab1 && r < p.N1) soff[j] = kq * p.lboA1 + r * 16;      // rows hp..N1-1 and the K padding are zeroed
        if (r < hp && kq < nq) goff[j] = r * wp + kq * 4;
      }
    }
    float4 nxt[MAXB];
    auto prefetch = [&](int n) {
      const float* src = p.x + plane_of(n) * plane_elems;
#pragma unroll
      for (int j = 0; j < MAXB; ++j)
        nxt[j] = goff[j] >= 0 ? __ldg(reinterpret_cast<const float4*>(src + goff[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // E1 work split: the accumulator rows (l, re|im) are replicated rep1 times over the lanes; copy r serves the
    // units (plane, 16-column chunk) with (plane + chunk) % rep1 == r, so that every quadrant (= SM sub-partition)
    // has its share
    const int qpr = 4 / p.rep1;
    const int my_rep = quad / qpr;
    const int nn = (quad % qpr) * 32 + lane;              // accumulator row within the copy
    const bool e1_warp = (quad % qpr) * 32 < 2 * m2;
    TCL_STAMP(0);
    if (my_planes > 0) prefetch(0);
    uint32_t ph_main = 0, ph_buf[2] = {0, 0};
    for (int n = 0; n < my_planes; ++n) {
      const int pl = n % cg;
      const int buf = p.nbuf == 2 ? (n & 1) : 0;
      if (n < 4) TCL_STAMP(2 + 4 * n);
      if (n >= p.nbuf) {          // the MMAs that read this buffer (plane n - nbuf) are complete
        mbar_wait(&bars[2 + buf], ph_buf[buf]);
        ph_buf[buf] ^= 1;
      }
      if (n < 4) TCL_STAMP(3 + 4 * n);
      // ---- transform: act, hi/lo split, operand layout; side output act(x) ----
      unsigned char* a1 = smem + p.offA1[buf];
      float* aout = p.a_out ? p.a_out + plane_of(n) * plane_elems : nullptr;
      float4 cur[MAXB];
#pragma unroll
      for (int j = 0; j < MAXB; ++j) cur[j] = nxt[j];
      if (n + 1 < my_planes) prefetch(n + 1);
#pragma unroll
      for (int j = 0; j < MAXB; ++j) {
        if (soff[j] >= 0) {
          float4 v = cur[j];
          if (p.act) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
          if (aout != nullptr && goff[j] >= 0) *reinterpret_cast<float4*>(aout + goff[j]) = v;
          st_split4(a1, p.partA1, (uint32_t)soff[j], v.x, v.y, v.z, v.w);
        }
      }
      fence_proxy_async();
      if (n < 4) TCL_STAMP(4 + 4 * n);
      nb_arrive(2 + (n & 1));
      if (pl != cg - 1) continue;

      // ---- E1: D1[(l,re|im), h] of every plane -> A2[(plane, l, re|im) rows][h columns] ----
      const size_t item = (size_t)blockIdx.x + (size_t)(n / cg) * gridDim.x;
      mbar_wait(&bars[1], ph_main);
      ph_main ^= 1;
      tc_fence_after();
      if (n < cg) TCL_STAMP(20);
      if (e1_warp) {
        unsigned char* a2 = smem + p.offA2;
        const int nch = p.N1 >> 4;
        // units (plane, chunk) with (plane + chunk) % rep1 == my_rep, dealt round robin to the NG warps of a quadrant
        int cnt = 0;
        for (int upl = 0; upl < cg; ++upl)
        for (int ch = (my_rep + p.rep1 - upl % p.rep1) % p.rep1; ch < nch; ch += p.rep1) {
          if (cnt++ % NG != grp) continue;
          float v[16];
          if (n < cg && cnt == 1) TCL_STAMP(27);
          tmem_ld16(lane_base + (uint32_t)(upl * p.N1 + ch * 16), v);
          if (n < cg && cnt == 1) TCL_STAMP(28);
          if (p.split1) {
            float v2[16], v3[16];
            tmem_ld16(lane_base + (uint32_t)(cg * p.N1 + upl * p.N1 + ch * 16), v2);
            tmem_ld16(lane_base + (uint32_t)(2 * cg * p.N1 + upl * p.N1 + ch * 16), v3);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] + v2[j]) + v3[j];
          }
          if (nn < 2 * m2) {
            const int e = upl * m2 + (nn >> 1);
            const int rho = ((e >> 4) << 5) + ((nn & 1) << 4) + (e & 15);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int h0 = ch * 16 + 4 * q;
              if (h0 < p.K2)
                st_split4(a2, p.partA2, ns_off(rho, h0, p.lboA2), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
          if (n < cg && cnt == 1) TCL_STAMP(29);
        }
        if (n < cg) TCL_STAMP(30);
      }
      tc_fence_before();
      if (n < cg) TCL_STAMP(31);
      fence_proxy_async();
      if (n < cg) TCL_STAMP(21);
      nb_arrive(4);
      // ---- E2: finish the complex product (lane pair re <-> im), scale, store the kept spectrum ----
      mbar_wait(&bars[1], ph_main);
      ph_main ^= 1;
      tc_fence_after();
      if (n < cg) TCL_STAMP(23);
      {
        const int e = quad * 16 + (lane & 15), reim = lane >> 4;
        const bool valid = e < E;
        const int upl = valid ? e / m2 : 0, l = valid ? e - upl * m2 : 0;
        const float sc = valid ? __ldg(p.pre + l) : 0.f;
        const size_t plane = item * cg + upl;          // = img * C + channel
        float* dst = reinterpret_cast<float*>(p.spec_out + plane * p.K * m2 + l) + reim;
        const int nch = p.N2 >> 4;
        if (quad * 16 < E) {
          for (int ch = grp; ch < nch; ch += NG) {
            float v[16];
            tmem_ld16(lane_base + (uint32_t)(ch * 16), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float ps = __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 16);
              const float val = reim ? v[2 * j] - ps : v[2 * j] + ps;
              const int k = ch * 8 + j;
              if (valid && k < p.K) dst[(size_t)k * m2 * 2] = val * sc;
            }
          }
        }
      }
      tc_fence_before();     // ordered before the next arrive: the next item's MMAs overwrite these accumulators
      if (n < cg) TCL_STAMP(24);
    }
    TCL_STAMP(26);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NW) {
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
  const float* f3; const float* f4;      // shared-memory operand images (without TS)
  const float* t3; const float* t4;      // the same operands for tensor memory: [2][tiles][128][K]
  int nitems, C, cg, hp, wp, m1, m2, K, act_in, passes;
  int N3, K3, N4, K4;      // S3: N3 = 2*cg*m2 padded to 16 columns, K3 = K padded to 8; S4: N4 = wp padded to 16, K4 = 2*m2 padded to 8
  int hb_shift, ntiles4;   // S4 row tile = cg blocks of HB = 128 / cg rows of h
  int npb;                 // planes per staging batch of the epilogue's input planes (two buffers)
  int lboA3, lboF3, lboA4, lboF4;
  uint32_t partA3, partA4, partF3, partF4;
  uint32_t offF3, offF4, offA3, offA4, offPw, offPost, offStg, offBar;
  int ntiles3;
  int ts3;
  uint32_t tmem_cols, colF3, colD;
  long long* dbg;
};

template <bool BWD, bool STREAM>
__global__ void __launch_bounds__(NT, 1) q_kernel(const QParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);   // [0] tables, [1] stage done, [2], [3] staging buffers full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* pws = reinterpret_cast<float*>(smem + p.offPw);           // [C*C] weights, [C] bias, then [C*C + C] gradient sums
  float* posts = reinterpret_cast<float*>(smem + p.offPost);       // [m2]
  const int C = p.C, cg = p.cg, hp = p.hp, wp = p.wp, m1 = p.m1, m2 = p.m2, K = p.K;
  float* gacc = pws + C * C + C;

  pdl_launch_dependents();
  if (tid == NWK) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_init_fence();
    mbar_expect_tx(&bars[0], (p.ts3 ? 0u : 2u * p.partF3) + 2u * p.partF4);
    if (!p.ts3) bulk_g2s(smem + p.offF3, p.f3, 2u * p.partF3, &bars[0]);
    bulk_g2s(smem + p.offF4, p.f4, 2u * p.partF4, &bars[0]);
  }
  if (warp == NW) tmem_alloc(tmem_slot, p.tmem_cols);
  if (tid < NWK) {
    // K padding of A4 (columns 2*m2..K4-1) is zeroed once: its producer only writes the kept modes.  Whole slabs
    // from the first one that holds a padded column.  (A3's producer writes its own padding.)
    const int rows4 = (p.lboA4 - 16) >> 4;
    for (int part = 0; part < 2; ++part)
      for (int s = (2 * m2) >> 2; s < (p.K4 >> 2); ++s)
        for (int i = tid; i < rows4 * 4; i += NWK)
          reinterpret_cast<float*>(smem + p.offA4 + part * p.partA4 + s * p.lboA4)[i] = 0.f;
    for (int i = tid; i < C * C; i += NWK) pws[i] = __ldg(p.pw_w + i);
    for (int i = tid; i < C; i += NWK) pws[C * C + i] = (!BWD && p.pw_b != nullptr) ? __ldg(p.pw_b + i) : 0.f;
    for (int i = tid; i < m2; i += NWK) posts[i] = __ldg(p.post + i);
    if (BWD)
      for (int i = tid; i < C * C + C; i += NWK) gacc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  if (warp < NW) {
    if (p.ts3) load_table_tmem(tmem_base, p.colF3, p.t3, p.ntiles3, p.K3, warp & 3, warp >> 2, lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();

  if (warp == NW) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc3 = idesc_n(p.N3), idesc4 = idesc_n(p.N4);
    if (lane == 0) mbar_wait(&bars[0], 0);
    __syncwarp();
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      nb_sync(2);                               // A3 (the mixed spectrum) is written and fenced
      tc_fence_after();
      if (item == (int)blockIdx.x) TCL_ISTAMP(32);
      if (lane == 0) {
        for (int t = 0; t < p.ntiles3; ++t) {
          const uint32_t d = tmem_base + p.colD + (uint32_t)(t * p.N3);
          if (p.ts3)
            issue_gemm_ts(d, tmem_base + p.colF3 + (uint32_t)(t * p.K3), tmem_base + p.colF3 + (uint32_t)((p.ntiles3 + t) * p.K3),
                          sbase + p.offA3, p.partA3, p.lboA3, p.K3 >> 3, p.passes, idesc3);
          else
            issue_gemm(d, sbase + p.offF3 + (uint32_t)t * 2048u, p.partF3, p.lboF3, sbase + p.offA3, p.partA3, p.lboA3,
                       p.K3 >> 3, p.passes, idesc3);
        }
        tc_commit(&bars[1]);
      }
      if (item == (int)blockIdx.x) TCL_ISTAMP(33);
      __syncwarp();
      nb_sync(3);                               // A4 is written and fenced, D3 has been read
      tc_fence_after();
      if (item == (int)blockIdx.x) TCL_ISTAMP(34);
      if (lane == 0) {
        for (int t = 0; t < p.ntiles4; ++t)
          issue_gemm(tmem_base + p.colD + (uint32_t)(t * p.N4), sbase + p.offA4 + (uint32_t)t * 2048u, p.partA4, p.lboA4,
                     sbase + p.offF4, p.partF4, p.lboF4, p.K4 >> 3, p.passes, idesc4);
        tc_commit(&bars[1]);
      }
      if (item == (int)blockIdx.x) TCL_ISTAMP(35);
      __syncwarp();
    }
  } else if (warp < NW) {
    // ------------------------------- workers -------------------------------
    const int quad = warp & 3, grp = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + p.colD;
    const int E = cg * m2;
    const size_t plane_elems = (size_t)hp * wp;
    unsigned char* a3 = smem + p.offA3;
    unsigned char* a4 = smem + p.offA4;
    const int HB = 1 << p.hb_shift;
    uint32_t ph = 0, ph_stg[2] = {0, 0};
    // Staging of the epilogue's input planes (forward: the 1x1 conv input act(z_in), backward: gz_out): batches of
    // npb planes of the item's image, two buffers, brought in by bulk async copies that are issued as early as the
    // buffers are free -- the first two batches of an item while its spectrum is still being mixed.
    const int nbatch = (C + p.npb - 1) / p.npb;
    const uint32_t stg_bytes = (uint32_t)(p.npb * plane_elems * sizeof(float));
    int gbase = 0;                             // batches staged for the items before this one: batch b is in buffer (gbase + b) & 1
    auto stage = [&](int img_, int b, int gb) {        // one thread; gb = the batch's running number
      const int i0 = b * p.npb, n = min(p.npb, C - i0);
      const uint32_t bytes = (uint32_t)(n * plane_elems * sizeof(float));
      fence_proxy_async();
      mbar_expect_tx(&bars[2 + (gb & 1)], bytes);
      bulk_g2s(smem + p.offStg + (gb & 1) * stg_bytes, p.a_in + ((size_t)img_ * C + i0) * plane_elems, bytes, &bars[2 + (gb & 1)]);
    };
    TCL_STAMP(0);
    if (tid == 0 && (int)blockIdx.x < p.nitems) {
      const int img0 = ((int)blockIdx.x * cg) / C;
      stage(img0, 0, 0);
      if (nbatch > 1) stage(img0, 1, 1);
    }

    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int img = (item * cg) / C, o0 = (item * cg) - img * C;
      const bool first = item == (int)blockIdx.x;
      if (first) TCL_STAMP(2);
      // ---- per-mode channel mix -> A3[(o, l, re|im) rows][k columns]; a unit is 4 consecutive kept rows k ----
      {
        const size_t cstride = (size_t)m1 * m2;
        const size_t xstride = (size_t)K * m2;
        const float2* xb = p.xin + (size_t)img * C * xstride;
        const int nkq = p.K3 >> 2;
        for (int u = tid; u < cg * nkq * m2; u += NWK) {
          const int l = u % m2, kq = (u / m2) % nkq, ol = u / (m2 * nkq);
          const int o = o0 + ol;
          float yr[4] = {0.f, 0.f, 0.f, 0.f}, yi[4] = {0.f, 0.f, 0.f, 0.f};
          size_t xoff[4], woff[4];
          const float2* wsel[4];
          bool kv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kq * 4 + j;
            kv[j] = k < K;
            const bool lo = k < m1;
            wsel[j] = lo ? p.w1 : p.w2;
            xoff[j] = (size_t)(kv[j] ? k : 0) * m2 + l;
            woff[j] = (size_t)(kv[j] ? (lo ? k : k - m1) : 0) * m2 + l;
          }
          for (int a0 = 0; a0 < C; a0 += CB) {
            float2 xv[CB][4], wv[CB][4];
#pragma unroll
            for (int aa = 0; aa < CB; ++aa) {
              const int a = a0 + aa < C ? a0 + aa : C - 1;
              const size_t wch = (size_t)(BWD ? o * C + a : a * C + o) * cstride;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                xv[aa][j] = __ldg(xb + (size_t)a * xstride + xoff[j]);
                wv[aa][j] = __ldg(wsel[j] + wch + woff[j]);
              }
            }
#pragma unroll
            for (int aa = 0; aa < CB; ++aa) {
              if (a0 + aa < C) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 x = xv[aa][j], w = wv[aa][j];
                  if (!BWD) {
                    yr[j] = fmaf(x.x, w.x, fmaf(-x.y, w.y, yr[j]));
                    yi[j] = fmaf(x.x, w.y, fmaf(x.y, w.x, yi[j]));
                  } else {
                    yr[j] = fmaf(x.x, w.x, fmaf(x.y, w.y, yr[j]));
                    yi[j] = fmaf(x.y, w.x, fmaf(-x.x, w.y, yi[j]));
                  }
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (!kv[j]) { yr[j] = 0.f; yi[j] = 0.f; }
          const int e = ol * m2 + l;
          st_split4(a3, p.partA3, ns_off(2 * e, kq * 4, p.lboA3), yr[0], yr[1], yr[2], yr[3]);
          st_split4(a3, p.partA3, ns_off(2 * e + 1, kq * 4, p.lboA3), yi[0], yi[1], yi[2], yi[3]);
        }
      }
      fence_proxy_async();
      if (first) TCL_STAMP(3);
      nb_arrive(2);
      mbar_wait(&bars[1], ph);
      ph ^= 1;
      tc_fence_after();
      if (first) TCL_STAMP(5);
      // ---- E3: D3[(h, cos|sin), (o, l, re|im)] -> finish the complex product, scale -> A4[(o, h)][(re|im, l)] ----
      // rows of a tile of 64 h: quadrant = h & 3, lane = (cos|sin) * 16 + (h >> 2): every quadrant has its share
      {
        const int cs = lane >> 4;
        const int hl = ((lane & 15) << 2) | quad;
        const int nch = p.N3 >> 4;
        for (int u = grp; u < p.ntiles3 * nch; u += NG) {
          const int t = u / nch, ch = u - t * nch;
          const int h = t * 64 + hl;
          if (t * 64 + quad >= hp) continue;            // warp-uniform: no row of this quadrant in the tile
          float v[16];
          tmem_ld16(lane_base + (uint32_t)(t * p.N3 + ch * 16), v);
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ps = __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 16);
            z[j] = cs ? v[2 * j] + ps : v[2 * j] - ps;
          }
          if (h < hp) {
            const int r4h = (h >> p.hb_shift) * 128 + (h & (HB - 1));       // row of (channel 0, h) in A4
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
              const int e0 = ch * 8 + g4 * 4;
              if (e0 < E) {
                const int ol = e0 / m2, l0 = e0 - ol * m2;
                if ((m2 & 3) == 0) {
                  const float4 sc = *reinterpret_cast<const float4*>(posts + l0);
                  st_split4(a4, p.partA4, ns_off(r4h + (ol << p.hb_shift), cs * m2 + l0, p.lboA4), z[g4 * 4] * sc.x, z[g4 * 4 + 1] * sc.y,
                            z[g4 * 4 + 2] * sc.z, z[g4 * 4 + 3] * sc.w);
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int e = e0 + j;
                    if (e < E) {
                      const int olj = e / m2, lj = e - olj * m2;
                      st_split1(a4, p.partA4, ns_off(r4h + (olj << p.hb_shift), cs * m2 + lj, p.lboA4), z[g4 * 4 + j] * posts[lj]);
                    }
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      if (first) TCL_STAMP(6);
      nb_arrive(3);
      mbar_wait(&bars[1], ph);
      ph ^= 1;
      tc_fence_after();
      if (first) TCL_STAMP(8);
      // ---- E4: epilogue on the accumulator rows (o, h) of every tile; a unit is 16 pixels of one row.  The input
      // planes come from the staging buffers: a lane reads its own row with 128-bit shared loads (conflict free for
      // wp / 4 odd), which replaces global loads that touched 32 different cache lines per instruction.
      // STREAM = false: the two buffers hold all C planes (C <= 2 * npb), units are worked one after the other.
      // STREAM = true (the heads, C = 12): the planes stream through the buffers batch by batch and a warp keeps
      // its (at most MAXU) units' accumulators in registers meanwhile. ----
      {
        const int r = quad * 32 + lane;
        const int ol = r >> p.hb_shift, hh = r & (HB - 1);      // HB >= 32: ol is uniform over a warp
        const int nch = p.N4 >> 4;
        const int nunits = p.ntiles4 * nch;
        const int ch_own = o0 + ol;                              // the channel this row belongs to
        constexpr int MAXU = STREAM ? 2 : 1;
        float acc[MAXU][16], a[BWD ? MAXU : 1][16];
        // geometry of unit u for this lane
        auto unit_geom = [&](int u, int& h, int& w0, int& nv4) {
          const int t = u / nch, ch = u - t * nch;
          h = t * HB + hh; w0 = ch * 16;
          const bool valid = h < hp && w0 < wp;
          nv4 = valid ? (min(16, wp - w0) >> 2) : 0;
          if (!valid) { h = 0; w0 = 0; }
          return (uint32_t)(t * p.N4 + ch * 16);
        };
        auto load_z = [&](int h, int w0, int nv4, float (&av)[16], float (&dv)[16]) {     // act(z_in), act'(z_in) of the own channel
          const float4* zp = reinterpret_cast<const float4*>(p.zin + ((size_t)img * C + ch_own) * plane_elems + (size_t)h * wp + w0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 z4 = q < nv4 ? __ldg(zp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if (p.act_in) {
                float cdf, pdf;
                gelu_cdf_pdf(zz[jj], cdf, pdf);
                av[4 * q + jj] = zz[jj] * cdf;
                dv[4 * q + jj] = fmaf(zz[jj], pdf, cdf);
              } else {
                av[4 * q + jj] = zz[jj];
                dv[4 * q + jj] = 1.0f;
              }
            }
          }
        };
        // one staged plane (channel i) into a unit's accumulators
        auto plane_fma = [&](const float* rowp, int i, int nv4, float (&ac)[16], const float (&av)[16]) {
          const float4* ap = reinterpret_cast<const float4*>(rowp);
          float4 x4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) x4[q] = q < nv4 ? ap[q] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float wv = BWD ? pws[i * C + ch_own] : pws[ch_own * C + i];
          float dot = 0.f, gsum = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ac[4 * q] = fmaf(wv, x4[q].x, ac[4 * q]);
            ac[4 * q + 1] = fmaf(wv, x4[q].y, ac[4 * q + 1]);
            ac[4 * q + 2] = fmaf(wv, x4[q].z, ac[4 * q + 2]);
            ac[4 * q + 3] = fmaf(wv, x4[q].w, ac[4 * q + 3]);
            if (BWD) {
              dot = fmaf(x4[q].x, av[4 * q], fmaf(x4[q].y, av[4 * q + 1], fmaf(x4[q].z, av[4 * q + 2], fmaf(x4[q].w, av[4 * q + 3], dot))));
              gsum += (x4[q].x + x4[q].y) + (x4[q].z + x4[q].w);
            }
          }
          if (BWD) {       // gW[o = i][own] += g[o] . act(z_in[own]);  gb[own] += sum g[own]
            dot = warp_sum(dot);
            if (lane == 0) atomicAdd(gacc + i * C + ch_own, dot);
            if (i == ch_own) {                                   // warp-uniform
              gsum = warp_sum(gsum);
              if (lane == 0) atomicAdd(gacc + C * C + ch_own, gsum);
            }
          }
        };
        auto store_unit = [&](int h, int w0, int nv4, const float (&ac)[16], const float (&dv)[16]) {
          float4* op = reinterpret_cast<float4*>(p.out + ((size_t)img * C + ch_own) * plane_elems + (size_t)h * wp + w0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < nv4) {
              if (!BWD) op[q] = make_float4(ac[4 * q], ac[4 * q + 1], ac[4 * q + 2], ac[4 * q + 3]);
              else op[q] = make_float4(ac[4 * q] * dv[4 * q], ac[4 * q + 1] * dv[4 * q + 1], ac[4 * q + 2] * dv[4 * q + 2], ac[4 * q + 3] * dv[4 * q + 3]);
            }
        };
        auto refill = [&](int b) {      // after every worker has read buffer b & 1: batch b + 2 of this item, or a first batch of the next
          asm volatile("bar.sync 5, %0;" ::"r"(NWK) : "memory");
          if (tid == 0) {
            if (b + 2 < nbatch) {
              stage(img, b + 2, gbase + b + 2);
            } else {
              const int nitem = item + (int)gridDim.x, nb0 = b + 2 - nbatch;
              if (nitem < p.nitems && nb0 < nbatch) stage((nitem * cg) / C, nb0, gbase + b + 2);
            }
          }
        };
        const float bias = BWD ? 0.f : pws[C * C + ch_own];
        if (!STREAM) {
          for (int b = 0; b < nbatch; ++b) {
            const int buf = (gbase + b) & 1;
            mbar_wait(&bars[2 + buf], ph_stg[buf]);
            ph_stg[buf] ^= 1;
          }
          for (int u = grp; u < nunits; u += NG) {
            int h, w0, nv4;
            const uint32_t col = unit_geom(u, h, w0, nv4);
            float dv[16];
            if (BWD) load_z(h, w0, nv4, a[0], dv);
            tmem_ld16(lane_base + col, acc[0]);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[0][j] += bias;
            for (int i = 0; i < C; ++i) {
              const int b = i / p.npb;
              const float* sp = reinterpret_cast<const float*>(smem + p.offStg + ((gbase + b) & 1) * stg_bytes) + (size_t)(i - b * p.npb) * plane_elems;
              plane_fma(sp + (size_t)h * wp + w0, i, nv4, acc[0], a[0]);
            }
            store_unit(h, w0, nv4, acc[0], dv);
          }
          // both buffers are free once every worker is through: bring in the next item's planes
          asm volatile("bar.sync 5, %0;" ::"r"(NWK) : "memory");
          if (tid == 0) {
            const int nitem = item + (int)gridDim.x;
            if (nitem < p.nitems)
              for (int b = 0; b < nbatch; ++b) stage((nitem * cg) / C, b, gbase + nbatch + b);
          }
        } else {
          int hs[MAXU], w0s[MAXU], nv4s[MAXU];
#pragma unroll
          for (int ui = 0; ui < MAXU; ++ui) {
            const int u = grp + ui * NG;
            hs[ui] = 0; w0s[ui] = 0; nv4s[ui] = 0;
            if (u < nunits) {
              const uint32_t col = unit_geom(u, hs[ui], w0s[ui], nv4s[ui]);
              if (BWD) {
                float dv[16];
                load_z(hs[ui], w0s[ui], nv4s[ui], a[BWD ? ui : 0], dv);
              }
              tmem_ld16(lane_base + col, acc[ui]);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[ui][j] += bias;
            }
          }
          for (int b = 0; b < nbatch; ++b) {
            const int buf = (gbase + b) & 1;
            mbar_wait(&bars[2 + buf], ph_stg[buf]);
            ph_stg[buf] ^= 1;
            const float* sp = reinterpret_cast<const float*>(smem + p.offStg + buf * stg_bytes);
            const int i0 = b * p.npb, nb = min(p.npb, C - i0);
#pragma unroll
            for (int ui = 0; ui < MAXU; ++ui)
              if (grp + ui * NG < nunits)
                for (int ii = 0; ii < nb; ++ii)
                  plane_fma(sp + (size_t)ii * plane_elems + (size_t)hs[ui] * wp + w0s[ui], i0 + ii, nv4s[ui], acc[ui], a[BWD ? ui : 0]);
            refill(b);
          }
#pragma unroll
          for (int ui = 0; ui < MAXU; ++ui)
            if (grp + ui * NG < nunits) {
              float av[16], dv[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) dv[j] = 1.f;
              if (BWD) load_z(hs[ui], w0s[ui], nv4s[ui], av, dv);     // act'(z_in) again rather than 16 more live registers
              store_unit(hs[ui], w0s[ui], nv4s[ui], acc[ui], dv);
            }
        }
      }
      gbase += nbatch;
      tc_fence_before();     // ordered before the next arrive: the next item's MMAs overwrite these accumulators
      if (first) TCL_STAMP(9);
    }
    TCL_STAMP(11);
  }

  tc_fence_before();
  __syncthreads();
  if (BWD) {
    for (int i = tid; i < C * C + C; i += NT) {
      const float s = gacc[i];
      if (s != 0.f) atomicAdd(i < C * C ? p.g_pw_w + i : p.g_pw_b + (i - C * C), s);
    }
  }
  if (warp == NW) {
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

// shared-memory image (hi | lo) of an operand with `rows` rows
template <typename F>
static void build_image(std::vector<float>& img, int rows, int K, int tiles, F&& value) {
  const int lbo = ns_lbo(rows);
  const size_t part = ns_part_bytes(rows, K, tiles) / 4;     // floats
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
// tensor-memory form [2 (hi, lo)][tiles][128][K] of an operand whose row r of tile t is value(t * 128 + r, k)
template <typename F>
static void build_tmem_table(std::vector<float>& tab, int tiles, int K, F&& value) {
  const size_t part = (size_t)tiles * 128 * K;
  tab.assign(2 * part, 0.f);
  for (int r = 0; r < tiles * 128; ++r)
    for (int k = 0; k < K; ++k) {
      const float v = (float)value(r, k);
      const float hi = host_rn_tf32(v);
      tab[(size_t)r * K + k] = hi;
      tab[part + (size_t)r * K + k] = host_rn_tf32(v - hi);
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

// operand dimensions that do not depend on the channel grouping
struct Dims { int N1, K1, rowsF1, rep1, N2, K2, K3, tiles3, N4, K4; };
static Dims dims_of(const Plan* pl) {
  Dims d;
  d.N1 = pad_to(pl->hp, 16); d.K1 = pad_to(pl->wp, 8);
  d.rowsF1 = pad_to(2 * pl->m2, 8);
  d.rep1 = 2 * pl->m2 <= 32 ? 4 : (2 * pl->m2 <= 64 ? 2 : 1);
  d.N2 = pad_to(2 * pl->K, 16); d.K2 = pad_to(pl->hp, 8);
  d.K3 = pad_to(pl->K, 8);
  d.tiles3 = (pl->hp + 63) / 64;
  d.N4 = pad_to(pl->wp, 16); d.K4 = pad_to(2 * pl->m2, 8);
  return d;
}

}  // namespace tcl

using namespace tcl;

// Built by get_plan for 2-D shapes that can ever fit (hp, wp <= 128); a failed upload only disables the path.
void tcl_build_tables(Plan* pl) {
  pl->tcl_f1 = pl->tcl_f1r = pl->tcl_f2 = pl->tcl_f3 = pl->tcl_f4 = pl->tcl_t1 = pl->tcl_t3 = pl->tcl_t4 = nullptr;
  if (pl->ndim != 2 || pl->hp > 128 || pl->wp > 128 || (pl->wp & 3) != 0 || 2 * pl->m2 > 128) return;
  const int hp = pl->hp, wp = pl->wp, m1 = pl->m1, m2 = pl->m2, K = pl->K;
  const double two_pi = 6.283185307179586476925286766559;
  auto theta = [&](int l, int w) { return two_pi * (double)(((long long)l * w) % wp) / (double)wp; };
  auto phi = [&](int k, int h) {
    const int kk = k < m1 ? k : hp - 2 * m1 + k;
    return two_pi * (double)(((long long)kk * h) % hp) / (double)hp;
  };
  const Dims d = dims_of(pl);
  // F1 (A operand of S1): rows (l, re|im), K = w
  auto f1 = [&](int n, int w) {
    if (n >= 2 * m2 || w >= wp) return 0.0;
    return (n & 1) ? -std::sin(theta(n >> 1, w)) : std::cos(theta(n >> 1, w));
  };
  // F3 (A operand of S3): a tile is 64 rows h x (cos, sin): quadrant = h & 3, lane = (cos|sin) * 16 + (h >> 2), K = k
  auto f3 = [&](int r, int k) {
    const int t = r >> 7, quad = (r >> 5) & 3, lane = r & 31;
    const int h = t * 64 + (((lane & 15) << 2) | quad), cs = lane >> 4;
    if (h >= hp || k >= K) return 0.0;
    return cs ? std::sin(phi(k, h)) : std::cos(phi(k, h));
  };
  // F4 (B operand of S4): rows w, K = (re|im) * m2 + l
  auto f4 = [&](int w, int c) {
    if (w >= wp || c >= 2 * m2) return 0.0;
    const int reim = c / m2, l = c - reim * m2;
    return reim ? -std::sin(theta(l, w)) : std::cos(theta(l, w));
  };
  std::vector<float> img;
  build_image(img, d.rowsF1, d.K1, 1, f1);
  pl->tcl_f1 = upload_f(img);
  build_image(img, 128, d.K1, 1, [&](int r, int w) { return f1(r % (128 / d.rep1), w); });
  pl->tcl_f1r = upload_f(img);      // rows replicated rep1 times: every TMEM quadrant gets a copy of the S1 result
  // F2 (B operand of S2): rows (k, cos|sin), K = h
  build_image(img, d.N2, d.K2, 1, [&](int n, int h) {
    if (n >= 2 * K || h >= hp) return 0.0;
    return (n & 1) ? std::sin(phi(n >> 1, h)) : std::cos(phi(n >> 1, h));
  });
  pl->tcl_f2 = upload_f(img);
  build_image(img, d.tiles3 * 128, d.K3, d.tiles3, f3);
  pl->tcl_f3 = upload_f(img);
  build_image(img, d.N4, d.K4, 1, f4);
  pl->tcl_f4 = upload_f(img);
  const int rows_per_copy = 128 / d.rep1;
  build_tmem_table(img, 1, d.K1, [&](int r, int w) { return f1(r % rows_per_copy, w); });
  pl->tcl_t1 = upload_f(img);
  build_tmem_table(img, d.tiles3, d.K3, f3);
  pl->tcl_t3 = upload_f(img);
  pl->tcl_t4 = pl->tcl_t3;      // (S4's constant operand is the B operand: always from shared memory)
  if (!pl->tcl_f1 || !pl->tcl_f1r || !pl->tcl_f2 || !pl->tcl_f3 || !pl->tcl_f4 || !pl->tcl_t1 || !pl->tcl_t3 || !pl->tcl_t4) {
    cudaGetLastError();
    pl->tcl_f1 = pl->tcl_f1r = pl->tcl_f2 = pl->tcl_f3 = pl->tcl_f4 = pl->tcl_t1 = pl->tcl_t3 = pl->tcl_t4 = nullptr;
  }
}

// ===========================================================================
// host: shared-memory / tensor-memory plans and launchers
// ===========================================================================
static const size_t TCL_SMEM_MAX = 227 * 1024;

static bool ts_enabled() {
  static const bool on = [] {
    const char* e = getenv("BDN_TCL_TS");       // BDN_TCL_TS=0: keep every operand in shared memory (diagnostic)
    return !(e && e[0] == '0');
  }();
  return on;
}

// BDN_TCL_DBG=1: every launch is followed by a synchronisation and a dump of block 0's phase time stamps (cycles
// since the first stamp) on stderr.  Diagnostic only.
static long long* dbg_buffer() {
  static long long* buf = [] {
    const char* e = getenv("BDN_TCL_DBG");
    long long* d = nullptr;
    if (e && e[0] == '1' && cudaMalloc(&d, 64 * sizeof(long long)) != cudaSuccess) d = nullptr;
    if (d) cudaMemset(d, 0, 64 * sizeof(long long));
    return d;
  }();
  return buf;
}
static void dbg_dump(const char* name, int C, int nitems, int cg, size_t smem, cudaStream_t st) {
  long long* d = dbg_buffer();
  if (!d) return;
  long long h[64];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  fprintf(stderr, "[tcl] %s C=%d items=%d cg=%d smem=%zu :", name, C, nitems, cg, smem);
  for (int i = 0; i < 44; ++i)
    if (h[i]) fprintf(stderr, " %d:%lld", i, h[i] - h[0]);
  fprintf(stderr, "\n");
  cudaMemset(d, 0, sizeof(h));
}

static int sm_count() {
  static const int n = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
  }();
  return n;
}

// channels per work item: as many as fit the 128 accumulator rows of S2 (and the 256 columns of S3), fewer when there
// are too few items to give every SM one
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

static bool plan_p(const Plan* pl, int images, int C, int passes, PParams& p, size_t& smem, int& maxb) {
  if (!pl->tcl_f1) return false;
  const int cg = pick_cg(images, C, pl->m2);
  if (cg == 0) return false;
  const Dims d = dims_of(pl);
  p.C = C; p.cg = cg; p.hp = pl->hp; p.wp = pl->wp; p.m2 = pl->m2; p.K = pl->K;
  p.nitems = images * C / cg;
  p.N1 = d.N1; p.K1 = d.K1; p.N2 = d.N2; p.K2 = d.K2;
  const int dcols = cg * p.N1 > p.N2 ? cg * p.N1 : p.N2;           // D2 reuses D1's columns (read before S2 is issued)
  if (p.N2 > 256 || p.N1 > 256) return false;
  {
    // F1 from tensor memory measured no faster than from shared memory (an instruction costs ~100 cycles either way),
    // so it stays opt-in: BDN_TCL_TS1=1
    const char* e = getenv("BDN_TCL_TS1");
    p.ts1 = e && e[0] == '1' && ts_enabled() && 2 * p.K1 + dcols <= 512;
  }
  if (!p.ts1 && dcols > 512) return false;
  {
    const char* e = getenv("BDN_TCL_SPLIT");     // experiment: independent accumulation chains in S1
    p.split1 = e && e[0] == '1' && passes == 3 && 3 * cg * p.N1 <= 512 && p.N2 <= 3 * cg * p.N1;
    if (p.split1) p.ts1 = 0;
  }
  p.rep1 = p.ts1 ? d.rep1 : 1;
  int rowsF1 = d.rowsF1;
  p.colF1 = 0;
  p.colD = p.ts1 ? (uint32_t)(2 * p.K1) : 0u;
  p.tmem_cols = pow2_cols((int)p.colD + (p.split1 ? 3 * cg * p.N1 : dcols));
  const int nblk = ((p.N1 + 7) / 8) * ((p.K1 / 4 + 3) / 4);
  if (nblk > NW * 11) return false;
  maxb = nblk <= NW * 5 ? 5 : (nblk <= NW * 7 ? 7 : (nblk <= NW * 9 ? 9 : 11));
  const int rowsA1 = p.N1, rowsA2 = pad_to(((cg * pl->m2 + 15) / 16) * 32, 8);
  p.lboA1 = ns_lbo(rowsA1); p.lboA2 = ns_lbo(rowsA2); p.lboF2 = ns_lbo(p.N2);
  p.partA1 = ns_part_bytes(rowsA1, p.K1);
  p.partA2 = ns_part_bytes(rowsA2, p.K2); p.partF2 = ns_part_bytes(p.N2, p.K2);
  p.passes = passes;
  // F1 from shared memory: first with its rows replicated over the 128 accumulator lanes (E1 then runs on every SM
  // sub-partition instead of the one that owns lanes 0..31), then compact.
  // A-operand layouts tried in order: two A1 buffers + own A2; one A1 buffer + own A2; A2 overlaid on the single A1 buffer
  for (int variant = 0; variant < 6; ++variant) {
    const bool replicate = variant < 3 && !p.ts1 && d.rep1 > 1 && !p.split1;
    if (variant < 3 && !replicate) continue;
    if (!p.ts1) {
      rowsF1 = replicate ? 128 : d.rowsF1;
      p.rep1 = replicate ? d.rep1 : 1;
    }
    p.lboF1 = ns_lbo(rowsF1);
    p.partF1 = ns_part_bytes(rowsF1, p.K1);
    const int nbuf = variant % 3 == 0 ? 2 : 1;
    const bool overlay = variant % 3 == 2;
    uint32_t off = 0;
    p.offF1 = off; if (!p.ts1) off += 2 * p.partF1;
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
  const Dims d = dims_of(pl);
  p.C = C; p.cg = cg; p.hp = pl->hp; p.wp = pl->wp; p.m1 = pl->m1; p.m2 = pl->m2; p.K = pl->K;
  p.nitems = images * C / cg;
  p.N3 = pad_to(2 * cg * pl->m2, 16); p.K3 = d.K3; p.N4 = d.N4; p.K4 = d.K4;
  p.ntiles3 = d.tiles3;
  p.hb_shift = cg == 4 ? 5 : (cg == 2 ? 6 : 7);
  const int HB = 128 / cg;
  p.ntiles4 = (pl->hp + HB - 1) / HB;
  const int dcols = p.ntiles3 * p.N3 > p.ntiles4 * p.N4 ? p.ntiles3 * p.N3 : p.ntiles4 * p.N4;       // D4 reuses D3's columns
  if (p.N3 > 256 || p.N4 > 256 || dcols > 512) return false;
  const int c3 = 2 * p.ntiles3 * p.K3;
  p.ts3 = ts_enabled() && c3 + dcols <= 512;
  p.colF3 = 0;
  p.colD = p.ts3 ? (uint32_t)c3 : 0u;
  p.tmem_cols = pow2_cols((int)p.colD + dcols);
  const int rowsA3 = p.N3;
  const int last_rows = pl->hp - (p.ntiles4 - 1) * HB;                        // h rows in the last tile
  const int rowsA4 = pad_to((p.ntiles4 - 1) * 128 + (cg - 1) * HB + last_rows, 8);
  p.lboA3 = ns_lbo(rowsA3); p.lboF3 = ns_lbo(d.tiles3 * 128); p.lboA4 = ns_lbo(rowsA4); p.lboF4 = ns_lbo(p.N4);
  p.partA3 = ns_part_bytes(rowsA3, p.K3); p.partF3 = ns_part_bytes(d.tiles3 * 128, p.K3, d.tiles3);
  p.partA4 = ns_part_bytes(rowsA4, p.K4, p.ntiles4);
  p.partF4 = ns_part_bytes(p.N4, p.K4);
  p.passes = passes;
  // epilogue staging: two buffers of npb planes; all C planes resident when they fit (npb = C / 2), else 2 per batch
  const size_t plane_bytes = (size_t)pl->hp * pl->wp * sizeof(float);
  uint32_t off = 0;
  p.offF3 = off; if (!p.ts3) off += 2 * p.partF3;
  p.offF4 = off; off += 2 * p.partF4;
  p.offA3 = off; off += 2 * p.partA3;
  p.offA4 = off; off += 2 * p.partA4;
  p.offPw = off; off += (uint32_t)pad_to((2 * (C * C + C)) * 4, 16);
  p.offPost = off; off += (uint32_t)pad_to(pl->m2 * 4, 16);
  p.offBar = off; off += 64;
  p.offStg = off;
  if (off > TCL_SMEM_MAX) return false;
  const int fit = (int)((TCL_SMEM_MAX - off) / (2 * plane_bytes));           // planes per buffer that fit
  if (fit < 1) return false;
  const int half = (C + 1) / 2;
  p.npb = fit >= half ? half : (fit >= 2 ? 2 : 1);
  const int nbatch = (C + p.npb - 1) / p.npb;
  // streaming keeps at most 2 units' accumulators per warp
  if (nbatch > 2 && p.ntiles4 * (p.N4 / 16) > 2 * NG) return false;
  off += (uint32_t)(2 * p.npb * plane_bytes);
  smem = off;
  return off <= TCL_SMEM_MAX;
}

bool tcl_supported(const Plan* pl, int images, int C) {
  if (pl == nullptr || pl->ndim != 2 || images < 1) return false;
  PParams pp;
  QParams qp;
  size_t s1 = 0, s2 = 0;
  int maxb = 0;
  return plan_p(pl, images, C, 3, pp, s1, maxb) && plan_q(pl, images, C, 3, qp, s2);
}

bool launch_tcl_p(const Plan* pl, const float* x, float* a_out, float2* spec_out, const float* pre, int images, int C,
                  int act, int prec, cudaStream_t st) {
  PParams p;
  size_t smem = 0;
  int maxb = 0;
  if (!plan_p(pl, images, C, prec == 2 ? 3 : 1, p, smem, maxb)) return false;
  LaunchScope scope(act ? "tc_p_gelu" : "tc_p", st, C);
  p.x = x; p.a_out = a_out; p.spec_out = spec_out; p.pre = pre;
  p.f1 = (!p.ts1 && p.rep1 > 1) ? pl->tcl_f1r : pl->tcl_f1; p.f2 = pl->tcl_f2; p.t1 = pl->tcl_t1;
  p.act = act;
  p.dbg = dbg_buffer();
  const int grid = p.nitems < sm_count() ? p.nitems : sm_count();
#define BDN_TCL_P(MB)                                                                                      \
  {                                                                                                        \
    cudaFuncSetAttribute(p_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCL_SMEM_MAX);    \
    launch_k(p_kernel<MB>, dim3(grid), dim3(NT), smem, st, p);                                             \
  }
  if (maxb == 5) BDN_TCL_P(5) else if (maxb == 7) BDN_TCL_P(7) else if (maxb == 9) BDN_TCL_P(9) else BDN_TCL_P(11)
#undef BDN_TCL_P
  dbg_dump(act ? "p_gelu" : "p", C, p.nitems, p.cg, smem, st);
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
  p.f3 = pl->tcl_f3; p.f4 = pl->tcl_f4; p.t3 = pl->tcl_t3; p.t4 = pl->tcl_t4;
  p.act_in = act_in;
  p.dbg = dbg_buffer();
  const int grid = p.nitems < sm_count() ? p.nitems : sm_count();
  const bool stream = (C + p.npb - 1) / p.npb > 2;
#define BDN_TCL_Q(B, S)                                                                                      \
  {                                                                                                          \
    cudaFuncSetAttribute(q_kernel<B, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCL_SMEM_MAX);    \
    launch_k(q_kernel<B, S>, dim3(grid), dim3(NT), smem, st, p);                                             \
  }
  if (bwd) { if (stream) BDN_TCL_Q(true, true) else BDN_TCL_Q(true, false) }
  else { if (stream) BDN_TCL_Q(false, true) else BDN_TCL_Q(false, false) }
#undef BDN_TCL_Q
  dbg_dump(bwd ? "q_bwd" : "q_fwd", C, p.nitems, p.cg, smem, st);
  return true;
}

}  // namespace bdn
