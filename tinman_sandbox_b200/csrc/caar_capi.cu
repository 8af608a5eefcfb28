// caar_capi.cu — the C-ABI of include/caar_b200.h over the CUDA kernels. No CPU fallback: every
// compute entry point needs a CUDA device and fails with CAAR_ERR_CUDA otherwise.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "caar_device.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU_TRY(expr)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(e_ == cudaErrorMemoryAllocation ? CAAR_ERR_NOMEM : CAAR_ERR_CUDA,           \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

size_t field_count(const caar_dims& d, int f) {
  const size_t E = (size_t)d.nelem, L = (size_t)d.nlev, P = 16;
  switch (f) {
    case 0: case 1: return E * P * 4;
    case 2: case 3: case 4: case 5: case 9: return E * P;
    case 6: case 8: return E * d.ntl * L * P;
    case 7: return E * d.ntl * L * P * 2;
    case 10: return E * d.qsize_d * 2 * L * P;
    case 11: return E * (L + 1) * P;
    case 12: case 13: case 14: return E * L * P;
    case 15: return E * L * P * 2;
  }
  return 0;
}

double* const* as_table(const caar_arrays* a) { return reinterpret_cast<double* const*>(a); }

}  // namespace

struct caar_handle_s {
  caar_dims dims;
  int device;
  cudaStream_t own_stream, stream;
  double* dev[CAAR_NUM_FIELDS];
  bool params_set;
  caar_constants c;
  double dvv[16], ps0, hyai0;
  double* partial;    // [nelem][3] device
  double* out3;       // [3] device
  double* out3_host;  // [3] pinned
  cudaEvent_t ev0, ev1;
  long long launches;
  caar::TmaMaps* tma;  // TMA descriptors of the device mirrors (nlev with a TMA-pipelined kernel only)
};

static_assert(sizeof(caar_arrays) == CAAR_NUM_FIELDS * sizeof(double*), "caar_arrays must be 16 pointers");

namespace {

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int validate_control(const caar_handle_s* h, const caar_control* ctl) {
  const caar_dims& d = h->dims;
  if (!ctl) return fail(CAAR_ERR_INVALID, "control is null");
  if (ctl->nets < 0 || ctl->nete > d.nelem || ctl->nets > ctl->nete)
    return fail(CAAR_ERR_INVALID, "element range [%d,%d) outside [0,%d)", ctl->nets, ctl->nete, d.nelem);
  const int tl[3] = {ctl->n0, ctl->np1, ctl->nm1};
  for (int i = 0; i < 3; ++i)
    if (tl[i] < 0 || tl[i] >= d.ntl) return fail(CAAR_ERR_INVALID, "time level %d outside [0,%d)", tl[i], d.ntl);
  if (ctl->qn0 < -1 || ctl->qn0 >= 2) return fail(CAAR_ERR_INVALID, "qn0=%d must be -1, 0 or 1", ctl->qn0);
  return CAAR_OK;
}

caar::KernelArgs make_args(const caar_handle_s* h, const caar_control* ctl) {
  caar::KernelArgs a;
  std::memset(&a, 0, sizeof a);
  a.D = h->dev[0]; a.Dinv = h->dev[1]; a.fcor = h->dev[2]; a.spheremp = h->dev[3];
  a.metdet = h->dev[4]; a.rmetdet = h->dev[5]; a.dp3d = h->dev[6]; a.v = h->dev[7];
  a.T = h->dev[8]; a.phis = h->dev[9]; a.Qdp = h->dev[10]; a.eta_dot_dpdn = h->dev[11];
  a.omega_p = h->dev[12]; a.phi = h->dev[13]; a.pecnd = h->dev[14]; a.vn0 = h->dev[15];
  a.nelem = h->dims.nelem; a.nlev = h->dims.nlev; a.qsize_d = h->dims.qsize_d; a.ntl = h->dims.ntl;
  if (ctl) {
    a.nets = ctl->nets; a.nete = ctl->nete; a.n0 = ctl->n0; a.np1 = ctl->np1; a.nm1 = ctl->nm1;
    a.qn0 = ctl->qn0; a.dt2 = ctl->dt2;
  }
  a.rrearth = h->c.rrearth; a.eta_ave_w = h->c.eta_ave_w; a.Rwv = h->c.Rwater_vapor; a.Rgas = h->c.Rgas;
  a.kappa = h->c.kappa; a.hyai0 = h->hyai0; a.ps0 = h->ps0;
  for (int i = 0; i < 16; ++i) a.dvv[i] = h->dvv[i];
  a.tma = h->tma;
  return a;
}

}  // namespace

extern "C" {

const char* caar_last_error(void) { return g_err; }
const char* caar_version(void) { return "caar_b200 0.1 (sm_100a)"; }

size_t caar_field_count(const caar_dims* dims, int f) {
  if (!dims || f < 0 || f >= CAAR_NUM_FIELDS) return 0;
  return field_count(*dims, f);
}

int caar_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(CAAR_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -CAAR_ERR_CUDA;
  }
  return n;
}

int caar_create(caar_handle* out, const caar_dims* dims, int device) {
  if (!out || !dims) return fail(CAAR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (dims->np != CAAR_NP) return fail(CAAR_ERR_UNSUPPORTED, "np=%d: only np=4 is implemented (reference NP=4)", dims->np);
  if (dims->nelem < 1 || dims->nlev < 2 || dims->qsize_d < 1 || dims->ntl < 1)
    return fail(CAAR_ERR_INVALID, "bad dims nelem=%d nlev=%d qsize_d=%d ntl=%d", dims->nelem, dims->nlev,
                dims->qsize_d, dims->ntl);
  if (caar::strict_smem_bytes(dims->nlev) > 200 * 1024 && !caar::fused_supports(dims->nlev))
    return fail(CAAR_ERR_UNSUPPORTED, "nlev=%d not covered by the compiled kernels", dims->nlev);
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(CAAR_ERR_CUDA, "device %d not available (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CAAR_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  caar_handle_s* h = new (std::nothrow) caar_handle_s();
  if (!h) return fail(CAAR_ERR_NOMEM, "host allocation failed");
  std::memset(h, 0, sizeof *h);
  h->dims = *dims;
  h->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  for (int f = 0; f < CAAR_NUM_FIELDS && e == cudaSuccess; ++f) {
    const size_t bytes = field_count(*dims, f) * sizeof(double);
    e = cudaMalloc(&h->dev[f], bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->dev[f], 0, bytes, h->own_stream);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->partial, (size_t)dims->nelem * 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->out3, 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost(&h->out3_host, 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
  if (e != cudaSuccess) {
    const int code = fail(e == cudaErrorMemoryAllocation ? CAAR_ERR_NOMEM : CAAR_ERR_CUDA, "caar_create: %s",
                          cudaGetErrorString(e));
    char keep[sizeof g_err];
    std::memcpy(keep, g_err, sizeof keep);
    caar_destroy(h);
    std::memcpy(g_err, keep, sizeof keep);
    cudaGetLastError();
    return code;
  }
  h->stream = h->own_stream;
  if (dims->nlev == 72 || dims->nlev == 128) {
    h->tma = new (std::nothrow) caar::TmaMaps();
    char msg[256] = "host allocation failed";
    if (!h->tma || caar::build_tma_maps(h->tma, make_args(h, nullptr), msg, sizeof msg)) {
      caar_destroy(h);
      return fail(CAAR_ERR_CUDA, "caar_create: %s", msg);
    }
  }
  *out = h;
  return CAAR_OK;
}

int caar_destroy(caar_handle h) {
  if (!h) return CAAR_OK;
  DeviceGuard guard(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  for (int f = 0; f < CAAR_NUM_FIELDS; ++f)
    if (h->dev[f]) cudaFree(h->dev[f]);
  if (h->partial) cudaFree(h->partial);
  if (h->out3) cudaFree(h->out3);
  if (h->out3_host) cudaFreeHost(h->out3_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h->tma;
  delete h;
  return CAAR_OK;
}

int caar_set_params(caar_handle h, const caar_constants* c, const double dvv[16], double ps0, const double* hyai) {
  if (!h || !c || !dvv || !hyai) return fail(CAAR_ERR_INVALID, "null argument");
  h->c = *c;
  std::memcpy(h->dvv, dvv, sizeof h->dvv);
  h->ps0 = ps0;
  h->hyai0 = hyai[0];
  h->params_set = true;
  return CAAR_OK;
}

int caar_set_stream(caar_handle h, void* cuda_stream) {
  if (!h) return fail(CAAR_ERR_INVALID, "null handle");
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return CAAR_OK;
}

static int copy_fields(caar_handle h, const caar_arrays* host, unsigned mask, bool to_device) {
  if (!h || !host) return fail(CAAR_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  double* const* tab = as_table(host);
  for (int f = 0; f < CAAR_NUM_FIELDS; ++f) {
    if (!(mask & (1u << f))) continue;
    if (!tab[f]) return fail(CAAR_ERR_INVALID, "host pointer of field %d is null", f);
    const size_t bytes = field_count(h->dims, f) * sizeof(double);
    if (to_device)
      CU_TRY(cudaMemcpyAsync(h->dev[f], tab[f], bytes, cudaMemcpyHostToDevice, h->stream));
    else
      CU_TRY(cudaMemcpyAsync(tab[f], h->dev[f], bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  CU_TRY(cudaStreamSynchronize(h->stream));
  return CAAR_OK;
}

int caar_upload(caar_handle h, const caar_arrays* host, unsigned mask) { return copy_fields(h, host, mask, true); }
int caar_download(caar_handle h, const caar_arrays* host, unsigned mask) { return copy_fields(h, host, mask, false); }

int caar_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(CAAR_ERR_INVALID, "null argument");
  CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return CAAR_OK;
}

int caar_host_unregister(void* ptr) {
  if (!ptr) return fail(CAAR_ERR_INVALID, "null argument");
  CU_TRY(cudaHostUnregister(ptr));
  return CAAR_OK;
}

int caar_device_arrays(caar_handle h, caar_arrays* out) {
  if (!h || !out) return fail(CAAR_ERR_INVALID, "null argument");
  std::memcpy(out, h->dev, sizeof h->dev);
  return CAAR_OK;
}

int caar_run(caar_handle h, const caar_control* ctl, int nsteps, int mode) {
  if (!h) return fail(CAAR_ERR_INVALID, "null handle");
  if (!h->params_set) return fail(CAAR_ERR_STATE, "caar_set_params must be called before caar_run");
  if (int rc = validate_control(h, ctl)) return rc;
  if (nsteps < 0) return fail(CAAR_ERR_INVALID, "nsteps=%d", nsteps);
  if (mode != CAAR_MODE_FAST && mode != CAAR_MODE_STRICT) return fail(CAAR_ERR_INVALID, "mode=%d", mode);
  DeviceGuard guard(h->device);
  const caar::KernelArgs a = make_args(h, ctl);
  const bool fast = (mode == CAAR_MODE_FAST) && caar::fused_supports(h->dims.nlev);
  if (!fast && caar::strict_smem_bytes(h->dims.nlev) > 200 * 1024)
    return fail(CAAR_ERR_UNSUPPORTED, "nlev=%d too large for the strict kernel", h->dims.nlev);
  if (ctl->nete == ctl->nets) return CAAR_OK;
  for (int s = 0; s < nsteps; ++s) {
    CU_TRY(fast ? caar::launch_fused(a, h->stream) : caar::launch_strict(a, h->stream));
    ++h->launches;
  }
  return CAAR_OK;
}

int caar_sync(caar_handle h) {
  if (!h) return fail(CAAR_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  CU_TRY(cudaStreamSynchronize(h->stream));
  return CAAR_OK;
}

long long caar_launch_count(caar_handle h) { return h ? h->launches : 0; }

int caar_timer_start(caar_handle h) {
  if (!h) return fail(CAAR_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  return CAAR_OK;
}

int caar_timer_stop(caar_handle h, float* ms) {
  if (!h || !ms) return fail(CAAR_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  CU_TRY(cudaEventRecord(h->ev1, h->stream));
  CU_TRY(cudaEventSynchronize(h->ev1));
  CU_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return CAAR_OK;
}

int caar_norms(caar_handle h, int tl, int nets, int nete, double sumsq[3]) {
  if (!h || !sumsq) return fail(CAAR_ERR_INVALID, "null argument");
  if (tl < 0 || tl >= h->dims.ntl) return fail(CAAR_ERR_INVALID, "time level %d", tl);
  if (nets < 0 || nete > h->dims.nelem || nets > nete) return fail(CAAR_ERR_INVALID, "element range [%d,%d)", nets, nete);
  DeviceGuard guard(h->device);
  const caar::KernelArgs a = make_args(h, nullptr);
  CU_TRY(caar::launch_norms(a, tl, nets, nete, h->partial, h->out3, h->stream));
  h->launches += (nete > nets) ? 2 : 1;
  CU_TRY(cudaMemcpyAsync(h->out3_host, h->out3, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 3; ++i) sumsq[i] = h->out3_host[i];
  return CAAR_OK;
}

int caar_compute_and_apply_rhs_host(const caar_dims* dims, const caar_arrays* host, const caar_control* ctl,
                                    const caar_constants* c, const double dvv[16], double ps0,
                                    const double* hyai, int device, int mode) {
  caar_handle h = nullptr;
  int rc = caar_create(&h, dims, device);
  if (rc) return rc;
  rc = caar_set_params(h, c, dvv, ps0, hyai);
  if (!rc) rc = caar_upload(h, host, CAAR_F_ALL);
  if (!rc) rc = caar_run(h, ctl, 1, mode);
  if (!rc) rc = caar_download(h, host, CAAR_F_MUTATED);
  char keep[sizeof g_err];
  std::memcpy(keep, g_err, sizeof keep);
  caar_destroy(h);
  std::memcpy(g_err, keep, sizeof keep);
  return rc;
}

int caar_saxpby_device(double a, double b, double* x_dev, const double* y_dev, size_t n, void* cuda_stream) {
  if (!x_dev || !y_dev) return fail(CAAR_ERR_INVALID, "null argument");
  CU_TRY(caar::launch_saxpby(a, b, x_dev, y_dev, n, static_cast<cudaStream_t>(cuda_stream)));
  return CAAR_OK;
}

int caar_saxpby_host(double a, double b, double* x, const double* y, size_t n, int sweeps, int device) {
  if (!x || !y) return fail(CAAR_ERR_INVALID, "null argument");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(CAAR_ERR_CUDA, "device %d not available", device);
  DeviceGuard guard(device);
  double *dx = nullptr, *dy = nullptr;
  CU_TRY(cudaMalloc(&dx, n * sizeof(double)));
  cudaError_t e = cudaMalloc(&dy, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dy, y, n * sizeof(double), cudaMemcpyHostToDevice);
  for (int s = 0; s < sweeps && e == cudaSuccess; ++s) e = caar::launch_saxpby(a, b, dx, dy, n, 0);
  if (e == cudaSuccess) e = cudaMemcpy(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dy);
  if (e != cudaSuccess) return fail(CAAR_ERR_CUDA, "caar_saxpby_host: %s", cudaGetErrorString(e));
  return CAAR_OK;
}

}  // extern "C"
