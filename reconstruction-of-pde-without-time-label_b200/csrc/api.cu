// C ABI of libblindno_b200 (include/blindno_b200.h): plan cache, argument checking and the
// host-side sequencing of the kernels of one FNO net / one spectral convolution.
#include "../../include/blindno_b200.h"
#include "bdn_internal.cuh"

#include <nvtx3/nvToolsExt.h>      // header-only (NVTX 3): ranges show up in Nsight Systems / Compute, no library to link

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

namespace bdn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

thread_local int pdl_few_images = 0;

bool pdl_enabled() {
  static const int mode = [] {
    const char* e = getenv("BDN_PDL");      // opt-in (profiles/README.md)
    return e ? atoi(e) : 0;
  }();
  return mode == 1 || (mode == 2 && !pdl_few_images);
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------------------
// per-kernel device timing (bench.py roofline): event pairs around every launch while enabled
// ---------------------------------------------------------------------------
struct ProfRecord { std::string name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfRecord> g_prof;

// Event pairs cannot be recorded into a stream that is being captured into a CUDA graph (the events would belong to
// the graph and could never be read back): profiling is skipped for launches under capture.
static bool stream_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st != cudaStreamCaptureStatusNone;
}

LaunchScope::LaunchScope(const char* n, cudaStream_t s, int t) : name(n), tag(t), st(s), e0(nullptr), on(false) {
  nvtxRangePushA(n);           // one NVTX range per kernel launch of this library (a no-op without a tool attached)
  if (g_prof_on.load(std::memory_order_relaxed) && !stream_capturing(s)) {
    on = cudaEventCreate(&e0) == cudaSuccess && cudaEventRecord(e0, st) == cudaSuccess;
  }
}

LaunchScope::~LaunchScope() {
  nvtxRangePop();
  count_launch(1);
  if (!on) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess || cudaEventRecord(e1, st) != cudaSuccess) return;
  std::string full(name);
  if (tag >= 0) full += "/" + std::to_string(tag);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back({full, e0, e1});
}

static int check_cuda(const char* where) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(BDN_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
  }
  return BDN_OK;
}

// ---------------------------------------------------------------------------
// plan cache
// ---------------------------------------------------------------------------
static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int, int, int, int>, Plan*> g_plans;

template <typename T>
static T* upload(const std::vector<T>& v) {
  if (v.empty()) return nullptr;
  T* d = nullptr;
  if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  return d;
}

static int round_up(int a, int m) { return (a + m - 1) / m * m; }

const Plan* get_plan(int ndim, int hp, int wp, int m1, int m2) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  const auto key = std::make_tuple(dev, ndim, hp, wp, m1, m2);
  std::lock_guard<std::mutex> lock(g_plan_mu);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) return it->second;

  Plan* pl = new Plan();
  pl->ndim = ndim; pl->hp = hp; pl->wp = wp; pl->m1 = m1; pl->m2 = m2;
  pl->K = ndim == 2 ? 2 * m1 : 1;
  pl->F = ndim == 2 ? m1 + 1 : 1;
  pl->Fp = round_up(pl->F, 4);
  pl->hp8 = round_up(hp, 8);
  pl->wp4 = round_up(wp, 4);
  const double two_pi = 6.283185307179586476925286766559;

  std::vector<float2> t_wl((size_t)wp * m2);
  std::vector<float> t_cos((size_t)m2 * pl->wp4, 0.f), t_sin((size_t)m2 * pl->wp4, 0.f);
  for (int w = 0; w < wp; ++w)
    for (int l = 0; l < m2; ++l) {
      const double th = two_pi * (double)(((long long)l * w) % wp) / (double)wp;
      const float c = (float)std::cos(th), s = (float)std::sin(th);
      t_wl[(size_t)w * m2 + l] = make_float2(c, s);
      t_cos[(size_t)l * pl->wp4 + w] = c;
      t_sin[(size_t)l * pl->wp4 + w] = s;
    }
  pl->t_wl = upload(t_wl);
  pl->wl_nh = wp / 2 + 1;
  pl->wl_nh4 = round_up(pl->wl_nh, 4);
  pl->wl_m2p = round_up(m2, 4);
  {
    std::vector<float2> half((size_t)pl->wl_nh4 * pl->wl_m2p, make_float2(0.f, 0.f));
    for (int w = 0; w < pl->wl_nh; ++w)
      for (int l = 0; l < m2; ++l) half[(size_t)w * pl->wl_m2p + l] = t_wl[(size_t)w * m2 + l];
    pl->t_wl_half = upload(half);      // optional: without it the unfolded kernel runs
  }
  pl->t_lw_cos = upload(t_cos);
  pl->t_lw_sin = upload(t_sin);

  pl->t_hf = nullptr; pl->t_fh = nullptr;
  if (ndim == 2) {
    // kept rows k = 0..m1-1 (frequency +k) and k = m1..2*m1-1 (frequency k - 2*m1 = -m1..-1): rows f and 2*m1 - f are a
    // conjugate pair, so one (cos, sin) per frequency f = 0..m1 serves both (f = 0 and f = m1 have one row each)
    std::vector<float2> t_hf((size_t)hp * pl->Fp, make_float2(0.f, 0.f));
    std::vector<float2> t_fh((size_t)pl->F * pl->hp8, make_float2(0.f, 0.f));
    for (int f = 0; f < pl->F; ++f)
      for (int h = 0; h < hp; ++h) {
        const double ph = two_pi * (double)(((long long)f * h) % hp) / (double)hp;
        const float2 v = make_float2((float)std::cos(ph), (float)std::sin(ph));
        t_hf[(size_t)h * pl->Fp + f] = v;
        t_fh[(size_t)f * pl->hp8 + h] = v;
      }
    pl->t_hf = upload(t_hf);
    pl->t_fh = upload(t_fh);
  }
  std::vector<float> col_fwd(m2), col_dc(m2, 1.f);
  for (int l = 0; l < m2; ++l) {
    const bool self_conj = l == 0 || (wp % 2 == 0 && l == wp / 2);
    col_fwd[l] = (float)((self_conj ? 1.0 : 2.0) / ((double)hp * (double)wp));
  }
  if (ndim == 1) col_dc[0] = 0.5f;
  pl->col_fwd = upload(col_fwd);
  pl->col_dc = upload(col_dc);
  pl->tc_fwd_b = nullptr; pl->tc_inv_b = nullptr;
  if ((wp & 3) == 0 && tc_n_pad(m2) <= 256) {
    std::vector<float> img;
    tc_build_b_image(wp, m2, img);
    pl->tc_fwd_b = upload(img);      // optional: a failed upload only disables the tensor-core path
  }

  tcl_build_tables(pl);              // optional as well (2-D, hp and wp <= 128)

  if (!pl->t_wl || !pl->t_lw_cos || !pl->t_lw_sin || !pl->col_fwd || !pl->col_dc ||
      (ndim == 2 && (!pl->t_hf || !pl->t_fh))) {
    set_error(BDN_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete pl;
    return nullptr;
  }
  g_plans[key] = pl;
  return pl;
}

static bool plan_cached(int ndim, int hp, int wp, int m1, int m2) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  return g_plans.count(std::make_tuple(dev, ndim, hp, wp, m1, m2)) != 0;
}

// The entry points' way to a plan.  Building one allocates and copies synchronously; inside a CUDA-graph capture that
// would invalidate the capture with an opaque error, so a first use of a shape under capture is refused with advice.
static const Plan* plan_for(int ndim, int hp, int wp, int m1, int m2, void* stream) {
  if (!plan_cached(ndim, hp, wp, m1, m2) && stream_capturing((cudaStream_t)stream)) {
    set_error(BDN_ERR_UNSUPPORTED,
              "first use of shape ndim=%d %dx%d modes %dx%d inside a CUDA-graph capture: run the call once eagerly, or "
              "call bdn_prepare_plan, before capturing", ndim, hp, wp, m1, m2);
    return nullptr;
  }
  return get_plan(ndim, hp, wp, m1, m2);
}

// ---------------------------------------------------------------------------
// shape helpers
// ---------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

static int check_modes(int ndim, int hp, int wp, int m1, int m2) {
  if (ndim != 1 && ndim != 2) return set_error(BDN_ERR_INVALID, "ndim must be 1 or 2 (got %d)", ndim);
  if (wp < 1 || hp < 1 || m2 < 1) return set_error(BDN_ERR_INVALID, "empty transform %dx%d modes %d", hp, wp, m2);
  if (m2 > wp / 2 + 1) return set_error(BDN_ERR_INVALID, "modes2=%d exceeds wp/2+1=%d", m2, wp / 2 + 1);
  if (ndim == 2 && (m1 < 1 || 2 * m1 > hp))
    return set_error(BDN_ERR_INVALID, "2*modes1=%d exceeds hp=%d (row blocks would overlap)", 2 * m1, hp);
  if (ndim == 1 && (hp != 1 || m1 != 0)) return set_error(BDN_ERR_INVALID, "1-D needs hp=1, m1=0");
  if (m2 > 64) return set_error(BDN_ERR_UNSUPPORTED, "modes2=%d > 64 not built", m2);
  return BDN_OK;
}

struct Carver {
  char* p; size_t left;
  void* take(size_t bytes) {
    bytes = align_up(bytes);
    if (bytes > left) return nullptr;
    void* r = p; p += bytes; left -= bytes;
    return r;
  }
};

static size_t spec1_bytes(int images, int c, int hp, int m2) { return (size_t)images * c * hp * m2 * sizeof(float2); }
static size_t kspec_bytes(int images, int c, int K, int m2) { return (size_t)images * c * K * m2 * sizeof(float2); }

}  // namespace bdn

using namespace bdn;

extern "C" {

int bdn_abi_version(void) { return BDN_ABI_VERSION; }
const char* bdn_last_error(void) { return g_err; }
int64_t bdn_kernel_launches(void) { return (int64_t)g_launches.load(); }

int bdn_pad_amount(int n) {
  // int(round(n * 0.25)) with round-half-to-even: n = 4q + r
  const int q = n / 4, r = n % 4;
  if (r < 2) return q;
  if (r > 2) return q + 1;
  return (q % 2 == 0) ? q : q + 1;   // exactly .5 -> nearest even
}

int bdn_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on.store(true);
  return BDN_OK;
}

long bdn_profile_end(char* buf, size_t cap) {
  g_prof_on.store(false);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1;
      a.second += ms;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    js += tmp;
    first = false;
  }
  js += "}";
  if (buf && cap > 0) {
    const size_t n = js.size() < cap - 1 ? js.size() : cap - 1;
    memcpy(buf, js.data(), n);
    buf[n] = 0;
  }
  return (long)js.size() + 1;
}

int bdn_prepare_plan(int32_t ndim, int32_t hp, int32_t wp, int32_t m1, int32_t m2) {
  int rc = check_modes(ndim, hp, wp, m1, m2);
  if (rc != BDN_OK) return rc;
  return get_plan(ndim, hp, wp, m1, m2) ? BDN_OK : BDN_ERR_CUDA;
}

int bdn_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}

// ---------------------------------------------------------------------------
// one spectral convolution
// ---------------------------------------------------------------------------
static int check_spectral(const BdnSpectralShape* s) {
  if (!s) return set_error(BDN_ERR_INVALID, "null shape");
  if (s->images < 0 || s->c_in < 1 || s->c_out < 1) return set_error(BDN_ERR_INVALID, "bad channel/batch counts");
  return check_modes(s->ndim, s->hp, s->wp, s->m1, s->m2);
}

size_t bdn_spectral_workspace_bytes(const BdnSpectralShape* s) {
  if (check_spectral(s) != BDN_OK) return 0;
  const int K = s->ndim == 2 ? 2 * s->m1 : 1;
  return align_up(spec1_bytes(s->images, s->c_in, s->hp, s->m2)) +
         align_up(spec1_bytes(s->images, s->c_out, s->hp, s->m2)) +
         align_up(kspec_bytes(s->images, s->c_out, K, s->m2)) + 256;
}

int bdn_spectral_forward(const BdnSpectralShape* s, const float* x, const float* w1, const float* w2, float* y,
                         float* xs_saved, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_spectral(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!x || !w1 || !y || !ws || (s->ndim == 2 && !w2)) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if (s->prec != BDN_PREC_FP32) return set_error(BDN_ERR_UNSUPPORTED, "standalone spectral op is fp32 only");
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv{(char*)ws, ws_bytes};
  float2* X1 = (float2*)cv.take(spec1_bytes(s->images, s->c_in, s->hp, s->m2));
  float2* Z = (float2*)cv.take(spec1_bytes(s->images, s->c_out, s->hp, s->m2));
  if (!X1 || !Z) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  launch_wfwd(pl, x, X1, s->images * s->c_in * s->hp, 0, st);
  if (s->ndim == 2)
    launch_core2d(pl, X1, Z, (float2*)xs_saved, (const float2*)w1, (const float2*)w2, s->images, s->c_in, s->c_out,
                  false, st);
  else
    launch_mix1d(pl, X1, Z, (float2*)xs_saved, (const float2*)w1, s->images, s->c_in, s->c_out, false, st);
  WinvArgs wa{};
  wa.z = Z; wa.y = y; wa.images = s->images; wa.c = s->c_out;
  launch_winv(pl, WINV_PLAIN, wa, st);
  return check_cuda("bdn_spectral_forward");
}

int bdn_spectral_backward(const BdnSpectralShape* s, const float* gy, const float* xs_saved, const float* w1,
                          const float* w2, float* gx, float* gw1, float* gw2, void* ws, size_t ws_bytes,
                          void* stream) {
  int rc = check_spectral(s);
  if (rc != BDN_OK) return rc;
  if (!gy || !xs_saved || !w1 || !gw1 || !ws || (s->ndim == 2 && (!w2 || !gw2)))
    return set_error(BDN_ERR_INVALID, "null pointer argument");
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int K = pl->K, m1r = s->ndim == 2 ? s->m1 : 1;
  const size_t wbytes = (size_t)s->c_in * s->c_out * m1r * s->m2 * sizeof(float2);
  cudaMemsetAsync(gw1, 0, wbytes, st);
  if (s->ndim == 2) cudaMemsetAsync(gw2, 0, wbytes, st);
  if (s->images == 0) return check_cuda("bdn_spectral_backward");
  Carver cv{(char*)ws, ws_bytes};
  float2* GZ = (float2*)cv.take(spec1_bytes(s->images, s->c_in, s->hp, s->m2));
  float2* G1 = (float2*)cv.take(spec1_bytes(s->images, s->c_out, s->hp, s->m2));
  float2* GY = (float2*)cv.take(kspec_bytes(s->images, s->c_out, K, s->m2));
  if (!GZ || !G1 || !GY) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  launch_wfwd(pl, gy, G1, s->images * s->c_out * s->hp, 0, st);
  if (s->ndim == 2)
    launch_core2d(pl, G1, GZ, GY, (const float2*)w1, (const float2*)w2, s->images, s->c_in, s->c_out, true, st);
  else
    launch_mix1d(pl, G1, GZ, GY, (const float2*)w1, s->images, s->c_in, s->c_out, true, st);
  launch_gw_reduce(pl, (const float2*)xs_saved, GY, (float2*)gw1, (float2*)gw2, s->images, s->c_in, s->c_out, st);
  if (gx) {
    WinvArgs wa{};
    wa.z = GZ; wa.y = gx; wa.images = s->images; wa.c = s->c_in;
    launch_winv(pl, WINV_PLAIN, wa, st);
  }
  return check_cuda("bdn_spectral_backward");
}

int bdn_stage_wfwd(int32_t hp, int32_t wp, int32_t m1, int32_t m2, int32_t rows, const float* x, float* out,
                   int32_t act, int32_t prec, void* stream) {
  const int ndim = hp > 1 ? 2 : 1;
  int rc = check_modes(ndim, hp, wp, m1, m2);
  if (rc != BDN_OK) return rc;
  if (!x || !out || rows < 0) return set_error(BDN_ERR_INVALID, "bad arguments");
  if (rows == 0) return BDN_OK;
  const Plan* pl = plan_for(ndim, hp, wp, m1, m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  // a single stage asked for in a tensor-core mode either runs on the tensor cores or fails: no silent change of
  // the arithmetic (the whole-net calls fall back per shape and say so, see launch_wfwd / use_tc_layer)
  if (prec != BDN_PREC_FP32 && !wfwd_uses_tensor_cores(pl, x, rows, prec))
    return set_error(BDN_ERR_UNSUPPORTED, "W-forward DFT wp=%d modes=%d rows=%d does not fit the tcgen05 kernel in mode %d", wp,
                     m2, rows, prec);
  launch_wfwd(pl, x, (float2*)out, rows, act, (cudaStream_t)stream, prec);
  return check_cuda("bdn_stage_wfwd");
}

// ---------------------------------------------------------------------------
// a whole FNO net
// ---------------------------------------------------------------------------
static int check_fno(const BdnFnoShape* s) {
  if (!s) return set_error(BDN_ERR_INVALID, "null shape");
  if (s->images < 0 || s->c_in < 1 || s->width < 1 || s->c_out < 1 || s->hidden < 1)
    return set_error(BDN_ERR_INVALID, "bad channel/batch counts");
  if (s->n_layers < 1 || s->n_layers > BDN_MAX_LAYERS)
    return set_error(BDN_ERR_INVALID, "n_layers=%d outside 1..%d", s->n_layers, BDN_MAX_LAYERS);
  if (s->width > 32 || s->c_in > 32) return set_error(BDN_ERR_UNSUPPORTED, "width/c_in > 32 not built");
  if (s->c_out > 4) return set_error(BDN_ERR_UNSUPPORTED, "c_out > 4 not built");
  if (s->hidden > 512) return set_error(BDN_ERR_UNSUPPORTED, "hidden > 512 not built");
  if (s->h < 1 || s->w < 1 || s->hp < s->h || s->wp < s->w) return set_error(BDN_ERR_INVALID, "bad grid/pad extents");
  if (s->out_h < 1 || s->out_w < 1 || s->out_h > s->hp || s->out_w > s->wp)
    return set_error(BDN_ERR_INVALID, "bad cropped extents %dx%d", s->out_h, s->out_w);
  if (s->ndim == 1 && s->h != 1) return set_error(BDN_ERR_INVALID, "1-D needs h=1");
  // the pointwise kernels split flat pixel indices with a float reciprocal (exact while the plane has < 2^22 pixels)
  if ((long)s->hp * s->wp >= (1L << 22)) return set_error(BDN_ERR_UNSUPPORTED, "padded plane %dx%d has >= 2^22 pixels", s->hp, s->wp);
  return check_modes(s->ndim, s->hp, s->wp, s->m1, s->m2);
}

static size_t act_floats1(const BdnFnoShape* s) { return (size_t)s->images * s->width * s->hp * s->wp; }
static size_t kspec_floats1(const BdnFnoShape* s) {
  const int K = s->ndim == 2 ? 2 * s->m1 : 1;
  return (size_t)s->images * s->width * K * s->m2 * 2;
}

size_t bdn_fno_act_floats(const BdnFnoShape* s) {
  return check_fno(s) == BDN_OK ? (size_t)(s->n_layers + 1) * act_floats1(s) : 0;
}
// Few-image nets (the output heads) keep a mode-major copy of their spectral weights next to the saved
// spectra: core2d blocks then own a single mode column each and read its coefficients contiguously.
static bool use_mode_major(const BdnFnoShape* s) { return s->ndim == 2 && (long)s->images * s->m2 < 4L * 2 * 148; }
static size_t wt_floats1(const BdnFnoShape* s) { return (size_t)s->m2 * 2 * s->m1 * s->width * s->width * 2; }

size_t bdn_fno_spec_floats(const BdnFnoShape* s) {
  if (check_fno(s) != BDN_OK) return 0;
  return (size_t)s->n_layers * kspec_floats1(s) + (use_mode_major(s) ? (size_t)s->n_layers * wt_floats1(s) : 0);
}
size_t bdn_fno_workspace_bytes(const BdnFnoShape* s) {
  if (check_fno(s) != BDN_OK) return 0;
  return 2 * align_up(spec1_bytes(s->images, s->width, s->hp, s->m2)) + 3 * align_up(act_floats1(s) * sizeof(float)) +
         align_up(kspec_floats1(s) * sizeof(float)) +
         (use_mode_major(s) ? align_up((size_t)s->n_layers * wt_floats1(s) * sizeof(float)) : 0) + 256;
}

// The fused tensor-core layer kernels (tc_layer.cu) serve 2-D nets in the TF32 / 3xTF32 modes when the shape fits
// their shared-memory plan.  A tensor-core mode that cannot be served says so once on stderr and runs the FFMA
// kernels (bdn_fno_layer_path reports which path a shape takes; the profile tags name the kernels that ran).
static bool use_tc_layer(const BdnFnoShape* s, const Plan* pl) {
  if (s->ndim != 2 || s->prec == BDN_PREC_FP32) return false;
  if (tcl_supported(pl, s->images, s->width)) return true;
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, int, int, int>, bool> seen;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(s->images, s->width, s->hp, s->wp, s->m1, s->m2);
  if (!seen[key]) {
    seen[key] = true;
    fprintf(stderr,
            "blindno_b200: tensor-core layer path not available for images=%d width=%d grid=%dx%d modes=%dx%d; "
            "running the fp32 FFMA kernels for this shape\n",
            s->images, s->width, s->hp, s->wp, s->m1, s->m2);
  }
  return false;
}

// A side stream per (device, caller stream slot) for work that depends on the parameters only (the mode-major copy of
// the heads' spectral weights): forked from the caller's stream by an event and joined before its first consumer, so
// it runs next to the lift and the first W transform instead of in front of them.  Under graph capture the fork / join
// become parallel branches.  The pool is created outside capture (first eager call); until then the work stays in
// the caller's stream.
struct SideCtx { cudaStream_t side; cudaEvent_t fork, join; };
static std::mutex g_side_mu;
static std::map<std::pair<int, int>, SideCtx> g_side;

static bool side_ctx(cudaStream_t st, SideCtx* out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  const int slot = (int)((reinterpret_cast<uintptr_t>(st) >> 4) % 8);
  std::lock_guard<std::mutex> lock(g_side_mu);
  auto it = g_side.find({dev, slot});
  if (it == g_side.end()) {
    if (stream_capturing(st)) return false;
    SideCtx c{};
    if (cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    it = g_side.emplace(std::make_pair(dev, slot), c).first;
  }
  *out = it->second;
  return true;
}

static LiftArgs make_lift(const BdnFnoShape* s, const BdnFnoParams* p, const BdnLiftInput* in) {
  LiftArgs a{};
  a.x_cl = in->x_cl; a.bags = in->bags; a.idx = in->idx; a.grid = in->grid;
  a.n_bags = in->n_bags; a.bag_len = in->bag_len; a.n_keep = in->n_keep; a.grid_dim = in->grid_dim;
  a.w0 = p->fc0_w; a.b0 = p->fc0_b;
  a.images = s->images; a.c_in = s->c_in; a.width = s->width;
  a.h = s->h; a.w = s->w; a.hp = s->hp; a.wp = s->wp;
  return a;
}

static int check_lift_input(const BdnFnoShape* s, const BdnLiftInput* in) {
  if (!in) return set_error(BDN_ERR_INVALID, "null lift input");
  if (in->x_cl) return BDN_OK;
  if (!in->bags || !in->grid) return set_error(BDN_ERR_INVALID, "lift input: need x_cl or bags+grid");
  if (in->n_bags * in->n_keep != s->images)
    return set_error(BDN_ERR_INVALID, "n_bags*n_keep=%d != images=%d", in->n_bags * in->n_keep, s->images);
  if (1 + in->grid_dim != s->c_in) return set_error(BDN_ERR_INVALID, "1+grid_dim=%d != c_in=%d", 1 + in->grid_dim, s->c_in);
  if (in->idx == nullptr && in->n_keep != in->bag_len)
    return set_error(BDN_ERR_INVALID, "idx is null but n_keep=%d != bag_len=%d", in->n_keep, in->bag_len);
  return BDN_OK;
}

static ProjArgs make_proj(const BdnFnoShape* s, const BdnFnoParams* p, const float* z) {
  ProjArgs a{};
  a.z = z; a.w1 = p->fc1_w; a.b1 = p->fc1_b; a.w2 = p->fc2_w; a.b2 = p->fc2_b;
  a.images = s->images; a.width = s->width; a.hidden = s->hidden; a.c_out = s->c_out;
  a.hp = s->hp; a.wp = s->wp; a.out_h = s->out_h; a.out_w = s->out_w;
  return a;
}

int bdn_fno_forward(const BdnFnoShape* s, const BdnFnoParams* p, const BdnLiftInput* in, float* out, float* z_saved,
                    float* xs_saved, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!p || !out || !ws) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if ((rc = check_lift_input(s, in)) != BDN_OK) return rc;
  if (s->prec != BDN_PREC_FP32 && s->prec != BDN_PREC_TF32 && s->prec != BDN_PREC_TF32X3)
    return set_error(BDN_ERR_INVALID, "bad precision");
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t act = act_floats1(s), ksp = kspec_floats1(s);
  Carver cv{(char*)ws, ws_bytes};
  float2* X1 = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float2* Z = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float* zping[2] = {nullptr, nullptr};
  if (!z_saved) {
    zping[0] = (float*)cv.take(act * sizeof(float));
    zping[1] = (float*)cv.take(act * sizeof(float));
  }
  if (!X1 || !Z || (!z_saved && (!zping[0] || !zping[1])))
    return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  auto zbuf = [&](int k) { return z_saved ? z_saved + (size_t)k * act : zping[k & 1]; };
  const bool tc = use_tc_layer(s, pl);
  struct FewGuard { int prev; FewGuard(bool few) : prev(pdl_few_images) { pdl_few_images = few; } ~FewGuard() { pdl_few_images = prev; } }
      few_guard(use_mode_major(s));
  float* abuf = nullptr;            // act(z_k) planes: written by kernel P, read by kernel Q's 1x1 conv
  if (tc && s->n_layers > 1 && !(abuf = (float*)cv.take(act * sizeof(float))))
    return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  float2* wt = nullptr;
  cudaEvent_t wt_join = nullptr;     // the mode-major copy runs on a side stream; joined before the first core2d
  if (use_mode_major(s) && !tc) {
    wt = xs_saved ? (float2*)(xs_saved + (size_t)s->n_layers * ksp)
                  : (float2*)cv.take((size_t)s->n_layers * wt_floats1(s) * sizeof(float));
    if (!wt) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
    SideCtx sc{};
    if (side_ctx(st, &sc) && cudaEventRecord(sc.fork, st) == cudaSuccess &&
        cudaStreamWaitEvent(sc.side, sc.fork, 0) == cudaSuccess) {
      launch_spec_weights_mode_major(pl, p->spec_w1, p->spec_w2, s->n_layers, s->width, s->width, wt, sc.side);
      cudaEventRecord(sc.join, sc.side);
      wt_join = sc.join;
    } else {
      cudaGetLastError();
      launch_spec_weights_mode_major(pl, p->spec_w1, p->spec_w2, s->n_layers, s->width, s->width, wt, st);
    }
  }
  const size_t wt1 = wt_floats1(s) / 2;   // float2 per layer

  launch_lift(make_lift(s, p, in), zbuf(0), st);
  const int rows = s->images * s->width * s->hp;
  for (int k = 0; k < s->n_layers; ++k) {
    const int act_in = k > 0;
    float2* xs_k = xs_saved ? (float2*)(xs_saved + (size_t)k * ksp) : nullptr;
    if (tc) {
      float2* xsp = xs_k ? xs_k : X1;
      launch_tcl_p(pl, zbuf(k), act_in ? abuf : nullptr, xsp, pl->col_dc, s->images, s->width, act_in, s->prec, st);
      launch_tcl_q(pl, false, xsp, (const float2*)p->spec_w1[k], (const float2*)p->spec_w2[k], act_in ? abuf : zbuf(k),
                   nullptr, zbuf(k + 1), p->conv_w[k], p->conv_b[k], nullptr, nullptr, pl->col_fwd, s->images, s->width,
                   act_in, s->prec, st);
      continue;
    }
    if (s->ndim == 1 && s->prec == BDN_PREC_FP32 &&
        launch_layer1d(pl, false, zbuf(k), nullptr, zbuf(k + 1), xs_k, (const float2*)p->spec_w1[k], p->conv_w[k],
                       p->conv_b[k], nullptr, nullptr, s->images, s->width, act_in, st))
      continue;
    launch_wfwd(pl, zbuf(k), X1, rows, act_in, st, s->prec);
    if (wt_join != nullptr) {
      cudaStreamWaitEvent(st, wt_join, 0);
      wt_join = nullptr;
    }
    if (s->ndim == 2)
      launch_core2d(pl, X1, Z, xs_k, (const float2*)p->spec_w1[k], (const float2*)p->spec_w2[k], s->images, s->width,
                    s->width, false, st, wt ? wt + (size_t)k * wt1 : nullptr);
    else
      launch_mix1d(pl, X1, Z, xs_k, (const float2*)p->spec_w1[k], s->images, s->width, s->width, false, st);
    WinvArgs wa{};
    wa.z = Z; wa.y = zbuf(k + 1); wa.a = zbuf(k); wa.pw_w = p->conv_w[k]; wa.pw_b = p->conv_b[k];
    wa.images = s->images; wa.c = s->width; wa.act_in = act_in;
    launch_winv(pl, WINV_LAYER_FWD, wa, st);
  }
  if (wt_join != nullptr) cudaStreamWaitEvent(st, wt_join, 0);
  launch_project(make_proj(s, p, zbuf(s->n_layers)), out, st);
  return check_cuda("bdn_fno_forward");
}

int bdn_fno_backward(const BdnFnoShape* s, const BdnFnoParams* p, const BdnLiftInput* in, const float* g_out,
                     int32_t pooled_g, int32_t n_keep, const float* z_saved, const float* xs_saved,
                     const BdnFnoGrads* g, float* gx_cl, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!p || !g_out || !z_saved || !xs_saved || !g || !ws) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if ((rc = check_lift_input(s, in)) != BDN_OK) return rc;
  if (pooled_g && (n_keep < 1 || s->images % n_keep != 0))
    return set_error(BDN_ERR_INVALID, "pooled gradient: images=%d not a multiple of n_keep=%d", s->images, n_keep);
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t act = act_floats1(s), ksp = kspec_floats1(s);
  Carver cv{(char*)ws, ws_bytes};
  float2* G1 = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float2* GZ = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float* gz[2];
  gz[0] = (float*)cv.take(act * sizeof(float));
  gz[1] = (float*)cv.take(act * sizeof(float));
  float2* GY = (float2*)cv.take(ksp * sizeof(float));
  if (!G1 || !GZ || !gz[0] || !gz[1] || !GY) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);

  const bool tc = use_tc_layer(s, pl);
  struct FewGuard { int prev; FewGuard(bool few) : prev(pdl_few_images) { pdl_few_images = few; } ~FewGuard() { pdl_few_images = prev; } }
      few_guard(use_mode_major(s));
  const float2* wt = use_mode_major(s) && !tc ? (const float2*)(xs_saved + (size_t)s->n_layers * ksp) : nullptr;
  const size_t wt1 = wt_floats1(s) / 2;
  int cur = 0;
  SideCtx sc{};
  const bool side_ok = use_mode_major(s) && !tc && side_ctx(st, &sc);
  cudaEvent_t gw_join = nullptr;
  launch_project_bwd(make_proj(s, p, z_saved + (size_t)s->n_layers * act), g_out, pooled_g, n_keep, gz[cur], g->fc1_w,
                     g->fc1_b, g->fc2_w, g->fc2_b, st);
  const int rows = s->images * s->width * s->hp;
  for (int k = s->n_layers - 1; k >= 0; --k) {
    if (tc) {
      launch_tcl_p(pl, gz[cur], nullptr, GY, pl->col_fwd, s->images, s->width, 0, s->prec, st);
      launch_gw_reduce(pl, (const float2*)(xs_saved + (size_t)k * ksp), GY, (float2*)g->spec_w1[k],
                       (float2*)g->spec_w2[k], s->images, s->width, s->width, st);
      launch_tcl_q(pl, true, GY, (const float2*)p->spec_w1[k], (const float2*)p->spec_w2[k], gz[cur],
                   z_saved + (size_t)k * act, gz[cur ^ 1], p->conv_w[k], nullptr, g->conv_w[k], g->conv_b[k], pl->col_dc,
                   s->images, s->width, k > 0, s->prec, st);
      cur ^= 1;
      continue;
    }
    if (s->ndim == 1 && s->prec == BDN_PREC_FP32 &&
        launch_layer1d(pl, true, z_saved + (size_t)k * act, gz[cur], gz[cur ^ 1], GY, (const float2*)p->spec_w1[k],
                       p->conv_w[k], nullptr, g->conv_w[k], g->conv_b[k], s->images, s->width, k > 0, st)) {
      launch_gw_reduce(pl, (const float2*)(xs_saved + (size_t)k * ksp), GY, (float2*)g->spec_w1[k],
                       (float2*)g->spec_w2[k], s->images, s->width, s->width, st);
      cur ^= 1;
      continue;
    }
    launch_wfwd(pl, gz[cur], G1, rows, 0, st, s->prec);
    if (gw_join != nullptr) {          // the previous layer's weight-gradient reduction still reads GY
      cudaStreamWaitEvent(st, gw_join, 0);
      gw_join = nullptr;
    }
    if (s->ndim == 2)
      launch_core2d(pl, G1, GZ, GY, (const float2*)p->spec_w1[k], (const float2*)p->spec_w2[k], s->images, s->width,
                    s->width, true, st, wt ? wt + (size_t)k * wt1 : nullptr);
    else
      launch_mix1d(pl, G1, GZ, GY, (const float2*)p->spec_w1[k], s->images, s->width, s->width, true, st);
    // few images (the heads): neither the weight-gradient reduction nor the inverse transform fills the machine, so the
    // reduction runs on the side stream next to the inverse transform and the next layer's W transform
    if (side_ok && cudaEventRecord(sc.fork, st) == cudaSuccess && cudaStreamWaitEvent(sc.side, sc.fork, 0) == cudaSuccess) {
      launch_gw_reduce(pl, (const float2*)(xs_saved + (size_t)k * ksp), GY, (float2*)g->spec_w1[k],
                       (float2*)g->spec_w2[k], s->images, s->width, s->width, sc.side);
      cudaEventRecord(sc.join, sc.side);
      gw_join = sc.join;
    } else {
      launch_gw_reduce(pl, (const float2*)(xs_saved + (size_t)k * ksp), GY, (float2*)g->spec_w1[k],
                       (float2*)g->spec_w2[k], s->images, s->width, s->width, st);
    }
    WinvArgs wa{};
    wa.z = GZ; wa.y = gz[cur ^ 1]; wa.a = gz[cur]; wa.zin = z_saved + (size_t)k * act;
    wa.pw_w = p->conv_w[k]; wa.g_pw_w = g->conv_w[k]; wa.g_pw_b = g->conv_b[k];
    wa.images = s->images; wa.c = s->width; wa.act_in = k > 0;
    launch_winv(pl, WINV_LAYER_BWD, wa, st);
    cur ^= 1;
  }
  launch_lift_bwd(make_lift(s, p, in), gz[cur], g->fc0_w, g->fc0_b, gx_cl, st);
  if (gw_join != nullptr) cudaStreamWaitEvent(st, gw_join, 0);
  return check_cuda("bdn_fno_backward");
}

// ---------------------------------------------------------------------------
// single stages of an FNO net (the custom-op layer exposes them one by one; the whole-net calls
// above run exactly these launches back to back)
// ---------------------------------------------------------------------------
static BdnFnoParams lift_only_params(const float* fc0_w, const float* fc0_b) {
  BdnFnoParams p{};
  p.fc0_w = fc0_w; p.fc0_b = fc0_b;
  return p;
}

int bdn_stage_lift_forward(const BdnFnoShape* s, const float* fc0_w, const float* fc0_b, const BdnLiftInput* in,
                           float* z0, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!fc0_w || !fc0_b || !z0) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if ((rc = check_lift_input(s, in)) != BDN_OK) return rc;
  const BdnFnoParams p = lift_only_params(fc0_w, fc0_b);
  launch_lift(make_lift(s, &p, in), z0, (cudaStream_t)stream);
  return check_cuda("bdn_stage_lift_forward");
}

int bdn_stage_lift_backward(const BdnFnoShape* s, const float* fc0_w, const float* fc0_b, const BdnLiftInput* in,
                            const float* gz0, float* g_fc0_w, float* g_fc0_b, float* gx_cl, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!fc0_w || !fc0_b || !gz0 || !g_fc0_w || !g_fc0_b) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if ((rc = check_lift_input(s, in)) != BDN_OK) return rc;
  const BdnFnoParams p = lift_only_params(fc0_w, fc0_b);
  launch_lift_bwd(make_lift(s, &p, in), gz0, g_fc0_w, g_fc0_b, gx_cl, (cudaStream_t)stream);
  return check_cuda("bdn_stage_lift_backward");
}

size_t bdn_stage_layer_workspace_bytes(const BdnFnoShape* s) {
  if (check_fno(s) != BDN_OK) return 0;
  return 2 * align_up(spec1_bytes(s->images, s->width, s->hp, s->m2)) + align_up(kspec_floats1(s) * sizeof(float)) +
         align_up(act_floats1(s) * sizeof(float)) + 256;
}

int bdn_fno_layer_path(const BdnFnoShape* s) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->ndim != 2 || s->prec == BDN_PREC_FP32) return 0;
  const Plan* pl = get_plan(s->ndim, s->hp, s->wp, s->m1, s->m2);
  if (!pl) return BDN_ERR_CUDA;
  return tcl_supported(pl, s->images, s->width) ? 1 : 0;
}

int bdn_stage_layer_forward(const BdnFnoShape* s, const float* z_in, int32_t act_in, const float* spec_w1,
                            const float* spec_w2, const float* conv_w, const float* conv_b, float* z_out,
                            float* xs_saved, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!z_in || !spec_w1 || (s->ndim == 2 && !spec_w2) || !conv_w || !conv_b || !z_out || !ws)
    return set_error(BDN_ERR_INVALID, "null pointer argument");
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv{(char*)ws, ws_bytes};
  float2* X1 = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float2* Z = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  if (!X1 || !Z) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  if (use_tc_layer(s, pl)) {
    float* abuf = nullptr;
    if (act_in && !(abuf = (float*)cv.take(act_floats1(s) * sizeof(float))))
      return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
    float2* xsp = xs_saved ? (float2*)xs_saved : X1;
    launch_tcl_p(pl, z_in, abuf, xsp, pl->col_dc, s->images, s->width, act_in != 0, s->prec, st);
    launch_tcl_q(pl, false, xsp, (const float2*)spec_w1, (const float2*)spec_w2, act_in ? abuf : z_in, nullptr, z_out,
                 conv_w, conv_b, nullptr, nullptr, pl->col_fwd, s->images, s->width, act_in != 0, s->prec, st);
    return check_cuda("bdn_stage_layer_forward");
  }
  if (s->ndim == 1 && s->prec == BDN_PREC_FP32 &&
      launch_layer1d(pl, false, z_in, nullptr, z_out, (float2*)xs_saved, (const float2*)spec_w1, conv_w, conv_b, nullptr,
                     nullptr, s->images, s->width, act_in != 0, st))
    return check_cuda("bdn_stage_layer_forward");
  launch_wfwd(pl, z_in, X1, s->images * s->width * s->hp, act_in != 0, st, s->prec);
  if (s->ndim == 2)
    launch_core2d(pl, X1, Z, (float2*)xs_saved, (const float2*)spec_w1, (const float2*)spec_w2, s->images, s->width,
                  s->width, false, st);
  else
    launch_mix1d(pl, X1, Z, (float2*)xs_saved, (const float2*)spec_w1, s->images, s->width, s->width, false, st);
  WinvArgs wa{};
  wa.z = Z; wa.y = z_out; wa.a = z_in; wa.pw_w = conv_w; wa.pw_b = conv_b;
  wa.images = s->images; wa.c = s->width; wa.act_in = act_in != 0;
  launch_winv(pl, WINV_LAYER_FWD, wa, st);
  return check_cuda("bdn_stage_layer_forward");
}

int bdn_stage_layer_backward(const BdnFnoShape* s, const float* gz_out, const float* z_in, int32_t act_in,
                             const float* xs_saved, const float* spec_w1, const float* spec_w2, const float* conv_w,
                             float* gz_in, float* g_spec_w1, float* g_spec_w2, float* g_conv_w, float* g_conv_b,
                             void* ws, size_t ws_bytes, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!gz_out || !z_in || !xs_saved || !spec_w1 || (s->ndim == 2 && (!spec_w2 || !g_spec_w2)) || !conv_w || !gz_in ||
      !g_spec_w1 || !g_conv_w || !g_conv_b || !ws)
    return set_error(BDN_ERR_INVALID, "null pointer argument");
  const Plan* pl = plan_for(s->ndim, s->hp, s->wp, s->m1, s->m2, stream);
  if (!pl) return BDN_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv{(char*)ws, ws_bytes};
  float2* G1 = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float2* GZ = (float2*)cv.take(spec1_bytes(s->images, s->width, s->hp, s->m2));
  float2* GY = (float2*)cv.take(kspec_floats1(s) * sizeof(float));
  if (!G1 || !GZ || !GY) return set_error(BDN_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  if (use_tc_layer(s, pl)) {
    launch_tcl_p(pl, gz_out, nullptr, GY, pl->col_fwd, s->images, s->width, 0, s->prec, st);
    launch_gw_reduce(pl, (const float2*)xs_saved, GY, (float2*)g_spec_w1, (float2*)g_spec_w2, s->images, s->width,
                     s->width, st);
    launch_tcl_q(pl, true, GY, (const float2*)spec_w1, (const float2*)spec_w2, gz_out, z_in, gz_in, conv_w, nullptr,
                 g_conv_w, g_conv_b, pl->col_dc, s->images, s->width, act_in != 0, s->prec, st);
    return check_cuda("bdn_stage_layer_backward");
  }
  if (s->ndim == 1 && s->prec == BDN_PREC_FP32 &&
      launch_layer1d(pl, true, z_in, gz_out, gz_in, GY, (const float2*)spec_w1, conv_w, nullptr, g_conv_w, g_conv_b,
                     s->images, s->width, act_in != 0, st)) {
    launch_gw_reduce(pl, (const float2*)xs_saved, GY, (float2*)g_spec_w1, (float2*)g_spec_w2, s->images, s->width,
                     s->width, st);
    return check_cuda("bdn_stage_layer_backward");
  }
  launch_wfwd(pl, gz_out, G1, s->images * s->width * s->hp, 0, st, s->prec);
  if (s->ndim == 2)
    launch_core2d(pl, G1, GZ, GY, (const float2*)spec_w1, (const float2*)spec_w2, s->images, s->width, s->width, true, st);
  else
    launch_mix1d(pl, G1, GZ, GY, (const float2*)spec_w1, s->images, s->width, s->width, true, st);
  launch_gw_reduce(pl, (const float2*)xs_saved, GY, (float2*)g_spec_w1, (float2*)g_spec_w2, s->images, s->width,
                   s->width, st);
  WinvArgs wa{};
  wa.z = GZ; wa.y = gz_in; wa.a = gz_out; wa.zin = z_in;
  wa.pw_w = conv_w; wa.g_pw_w = g_conv_w; wa.g_pw_b = g_conv_b;
  wa.images = s->images; wa.c = s->width; wa.act_in = act_in != 0;
  launch_winv(pl, WINV_LAYER_BWD, wa, st);
  return check_cuda("bdn_stage_layer_backward");
}

static BdnFnoParams proj_only_params(const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b) {
  BdnFnoParams p{};
  p.fc1_w = fc1_w; p.fc1_b = fc1_b; p.fc2_w = fc2_w; p.fc2_b = fc2_b;
  return p;
}

int bdn_stage_project_forward(const BdnFnoShape* s, const float* z, const float* fc1_w, const float* fc1_b,
                              const float* fc2_w, const float* fc2_b, float* out, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!z || !fc1_w || !fc1_b || !fc2_w || !fc2_b || !out) return set_error(BDN_ERR_INVALID, "null pointer argument");
  const BdnFnoParams p = proj_only_params(fc1_w, fc1_b, fc2_w, fc2_b);
  launch_project(make_proj(s, &p, z), out, (cudaStream_t)stream);
  return check_cuda("bdn_stage_project_forward");
}

int bdn_stage_project_backward(const BdnFnoShape* s, const float* z, const float* fc1_w, const float* fc1_b,
                               const float* fc2_w, const float* fc2_b, const float* g_out, int32_t pooled_g,
                               int32_t n_keep, float* gz, float* g_fc1_w, float* g_fc1_b, float* g_fc2_w,
                               float* g_fc2_b, void* stream) {
  int rc = check_fno(s);
  if (rc != BDN_OK) return rc;
  if (s->images == 0) return BDN_OK;
  if (!z || !fc1_w || !fc1_b || !fc2_w || !fc2_b || !g_out || !gz || !g_fc1_w || !g_fc1_b || !g_fc2_w || !g_fc2_b)
    return set_error(BDN_ERR_INVALID, "null pointer argument");
  if (pooled_g && (n_keep < 1 || s->images % n_keep != 0))
    return set_error(BDN_ERR_INVALID, "pooled gradient: images=%d not a multiple of n_keep=%d", s->images, n_keep);
  const BdnFnoParams p = proj_only_params(fc1_w, fc1_b, fc2_w, fc2_b);
  launch_project_bwd(make_proj(s, &p, z), g_out, pooled_g, n_keep, gz, g_fc1_w, g_fc1_b, g_fc2_w, g_fc2_b,
                     (cudaStream_t)stream);
  return check_cuda("bdn_stage_project_backward");
}

// ---------------------------------------------------------------------------
// bag pool, optimiser
// ---------------------------------------------------------------------------
int bdn_bag_pool_lift_forward(const float* s, const float* grid, const float* w0, const float* b0, float* out,
                              int32_t n_bags, int32_t n_keep, int32_t npix, int32_t grid_dim, int32_t width,
                              void* stream) {
  if (n_bags < 0 || n_keep < 1 || npix < 1 || grid_dim < 1 || width < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (n_bags == 0) return BDN_OK;          // an empty batch has empty (null) buffers: nothing to do
  if (!s || !grid || !w0 || !b0 || !out) return set_error(BDN_ERR_INVALID, "null pointer argument");
  launch_pool_lift(s, grid, w0, b0, out, n_bags, n_keep, npix, grid_dim, width, (cudaStream_t)stream);
  return check_cuda("bdn_bag_pool_lift_forward");
}

int bdn_bag_pool_lift_backward(const float* g, const float* w0, float* gpool, int32_t n_bags, int32_t npix,
                               int32_t grid_dim, int32_t width, void* stream) {
  if (n_bags < 0 || npix < 1 || grid_dim < 1 || width < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (n_bags == 0) return BDN_OK;
  if (!g || !w0 || !gpool) return set_error(BDN_ERR_INVALID, "null pointer argument");
  launch_pool_lift_bwd(g, w0, gpool, n_bags, npix, grid_dim, width, (cudaStream_t)stream);
  return check_cuda("bdn_bag_pool_lift_backward");
}

int bdn_nio_tail_forward(const float* w, const float* basis, const float* b0, const float* grid, const float* fc0_w,
                         const float* fc0_b, float* out, float* wbar_saved, int32_t n_bags, int32_t n_keep, int32_t p,
                         int32_t npix, int32_t grid_dim, int32_t width, void* stream) {
  if (n_bags < 0 || n_keep < 1 || p < 1 || npix < 1 || grid_dim < 1 || width < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (p > 256) return set_error(BDN_ERR_UNSUPPORTED, "n_basis=%d > 256 not built", p);
  if (n_bags == 0) return BDN_OK;
  if (!w || !basis || !b0 || !grid || !fc0_w || !fc0_b || !out || !wbar_saved) return set_error(BDN_ERR_INVALID, "null pointer argument");
  launch_nio_tail(w, basis, b0, grid, fc0_w, fc0_b, out, wbar_saved, n_bags, n_keep, p, npix, grid_dim, width, (cudaStream_t)stream);
  return check_cuda("bdn_nio_tail_forward");
}

int bdn_nio_tail_backward(const float* g, const float* basis, const float* wbar_saved, const float* fc0_w, float* g_w,
                          float* g_basis, float* g_b0, float* g_wbar_ws, int32_t n_bags, int32_t n_keep, int32_t p, int32_t npix,
                          int32_t grid_dim, int32_t width, void* stream) {
  if (n_bags < 0 || n_keep < 1 || p < 1 || npix < 1 || grid_dim < 1 || width < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (p > 256) return set_error(BDN_ERR_UNSUPPORTED, "n_basis=%d > 256 not built", p);
  if (!g_basis || !g_b0) return set_error(BDN_ERR_INVALID, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(g_basis, 0, (size_t)npix * p * sizeof(float), st);
  cudaMemsetAsync(g_b0, 0, sizeof(float), st);
  if (n_bags == 0) return check_cuda("bdn_nio_tail_backward");
  if (!g || !basis || !wbar_saved || !fc0_w || !g_w || !g_wbar_ws) return set_error(BDN_ERR_INVALID, "null pointer argument");
  cudaMemsetAsync(g_wbar_ws, 0, (size_t)n_bags * p * sizeof(float), st);
  launch_nio_tail_bwd(g, basis, wbar_saved, fc0_w, g_wbar_ws, g_basis, g_b0, g_w, n_bags, n_keep, p, npix, grid_dim, width, st);
  return check_cuda("bdn_nio_tail_backward");
}

size_t bdn_bag_attention_saved_floats(int32_t n_bags, int32_t n_keep) {
  return (n_bags < 1 || n_keep < 1 || n_keep > 128) ? 0 : bagattn_saved_floats(n_bags, n_keep);
}
size_t bdn_bag_attention_workspace_floats(int32_t n_bags, int32_t n_keep) {
  return (n_bags < 1 || n_keep < 1 || n_keep > 128) ? 0 : bagattn_backward_ws_floats(n_bags, n_keep);
}

static int check_bag_attention(int32_t n_bags, int32_t n_keep, int32_t dim) {
  if (n_bags < 0 || n_keep < 1 || dim < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (n_keep > 128) return set_error(BDN_ERR_UNSUPPORTED, "bag of %d snapshots > 128 not built", n_keep);
  return BDN_OK;
}

int bdn_bag_attention_mean_forward(const float* x, const float* ln_w, const float* ln_b, float* out, float* saved,
                                   int32_t n_bags, int32_t n_keep, int32_t dim, float eps, void* stream) {
  int rc = check_bag_attention(n_bags, n_keep, dim);
  if (rc != BDN_OK) return rc;
  if (n_bags == 0) return BDN_OK;
  if (!x || !ln_w || !ln_b || !out || !saved) return set_error(BDN_ERR_INVALID, "null pointer argument");
  launch_bagattn_forward(x, ln_w, ln_b, out, saved, n_bags, n_keep, dim, eps, (cudaStream_t)stream);
  return check_cuda("bdn_bag_attention_mean_forward");
}

int bdn_bag_attention_mean_backward(const float* x, const float* g, const float* ln_w, const float* saved, float* g_x,
                                    float* g_ln_w_per_bag, float* ws, int32_t n_bags, int32_t n_keep, int32_t dim,
                                    void* stream) {
  int rc = check_bag_attention(n_bags, n_keep, dim);
  if (rc != BDN_OK) return rc;
  if (n_bags == 0) return BDN_OK;
  if (!x || !g || !ln_w || !saved || !g_x || !g_ln_w_per_bag || !ws) return set_error(BDN_ERR_INVALID, "null pointer argument");
  launch_bagattn_backward(x, g, ln_w, saved, g_x, g_ln_w_per_bag, ws, n_bags, n_keep, dim, (cudaStream_t)stream);
  return check_cuda("bdn_bag_attention_mean_backward");
}

int bdn_mse_heads_forward(const float* const* outs, int32_t n_heads, int32_t c, int64_t npix, const float* target,
                          float* loss, float* const* g_outs, void* scratch, void* stream) {
  if (n_heads < 1 || n_heads > MSE_MAX_HEADS || c < 1 || npix < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (npix * n_heads * c >= (1LL << 31)) return set_error(BDN_ERR_UNSUPPORTED, "more than 2^31 output elements");
  if (!outs || !target || !loss || !scratch) return set_error(BDN_ERR_INVALID, "null pointer argument");
  MseHeadsArgs a{};
  for (int k = 0; k < n_heads; ++k) {
    if (!outs[k] || (g_outs && !g_outs[k])) return set_error(BDN_ERR_INVALID, "null head tensor");
    a.out[k] = outs[k];
    a.g[k] = g_outs ? g_outs[k] : nullptr;
  }
  a.target = target; a.n_heads = n_heads; a.c = c; a.npix = npix;
  // scratch: [0] the block counter (zero before the first launch; every launch leaves it zero), [1 .. 64] partial sums
  launch_mse_heads(a, loss, (float*)scratch + 1, (unsigned int*)scratch, (cudaStream_t)stream);
  return check_cuda("bdn_mse_heads_forward");
}

int bdn_mse_heads_backward(const float* const* outs, int32_t n_heads, int32_t c, int64_t npix, const float* target,
                           const float* grad_loss, float* const* g_outs, void* stream) {
  if (n_heads < 1 || n_heads > MSE_MAX_HEADS || c < 1 || npix < 1) return set_error(BDN_ERR_INVALID, "bad sizes");
  if (npix * n_heads * c >= (1LL << 31)) return set_error(BDN_ERR_UNSUPPORTED, "more than 2^31 output elements");
  if (!outs || !target || !grad_loss || !g_outs) return set_error(BDN_ERR_INVALID, "null pointer argument");
  MseHeadsArgs a{};
  for (int k = 0; k < n_heads; ++k) {
    if (!outs[k] || !g_outs[k]) return set_error(BDN_ERR_INVALID, "null head tensor");
    a.out[k] = outs[k];
    a.g[k] = g_outs[k];
  }
  a.target = target; a.n_heads = n_heads; a.c = c; a.npix = npix;
  launch_mse_heads_bwd(a, grad_loss, (cudaStream_t)stream);
  return check_cuda("bdn_mse_heads_backward");
}

int bdn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return set_error(BDN_ERR_INVALID, "null pointer argument");
  if (step < 1) return set_error(BDN_ERR_INVALID, "step must be >= 1");
  if (n == 0) return BDN_OK;
  launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
  return check_cuda("bdn_adam_step");
}

}  // extern "C"
