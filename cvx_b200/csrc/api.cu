// C ABI of libcvxb (include/cvxb.h): handle lifetime, parameters, and seam B -- the per-step linear
// algebra entry points on caller-owned column-major matrices (KKTSystem.solve, choleskySolve,
// ruizEquilibrate, regularizedCholesky, triangularSolve, SymmetricLinearSystem.solve).
// Host arrays are staged into padded device buffers; with CVXB_FLAG_DEVICE_PTRS the arrays are
// device pointers (copied device-to-device only when their alignment / leading dimension does not
// satisfy the kernels' 16-byte requirement).
#include <cstdarg>
#include <vector>
#include "kkt.cuh"

namespace cvxb {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// -------------------------------------------------------------------------------- staging helpers
struct Staged {
  double* d = nullptr;
  int ld = 0;
  bool owned = false;
  ~Staged() { if (owned && d) cudaFree(d); }
};

static bool aligned_ok(const double* p, int ld) { return (((uintptr_t)p & 15) == 0) && ((ld & 1) == 0); }

// rows x cols column-major source (host or device) -> padded device matrix
int stage_in(Handle& h, int rows, int cols, const double* src, int lds, Staged& out) {
  if (rows <= 0 || cols <= 0) { out.d = nullptr; out.ld = pad_ld(rows); return CVXB_OK; }
  if (!src) { set_last_error("null matrix argument"); return CVXB_EINVAL; }
  if (lds < rows) { set_last_error("leading dimension %d < rows %d", lds, rows); return CVXB_EDIM; }
  bool dev = (h.flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  if (dev && aligned_ok(src, lds)) { out.d = const_cast<double*>(src); out.ld = lds; out.owned = false; return CVXB_OK; }
  out.ld = pad_ld(rows);
  CVXB_CUDA_OK(cudaMalloc((void**)&out.d, (size_t)out.ld * cols * sizeof(double)));
  out.owned = true;
  if (out.ld != rows) CVXB_CUDA_OK(cudaMemsetAsync(out.d, 0, (size_t)out.ld * cols * sizeof(double), h.stream));
  CVXB_CUDA_OK(cudaMemcpy2DAsync(out.d, (size_t)out.ld * sizeof(double), src, (size_t)lds * sizeof(double),
                                 (size_t)rows * sizeof(double), cols, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 h.stream));
  return CVXB_OK;
}
int stage_out_alloc(Handle& h, int rows, int cols, Staged& out) {
  (void)h;
  out.ld = pad_ld(rows);
  CVXB_CUDA_OK(cudaMalloc((void**)&out.d, (size_t)out.ld * (cols > 0 ? cols : 1) * sizeof(double)));
  out.owned = true;
  return CVXB_OK;
}
int copy_out(Handle& h, int rows, int cols, const Staged& s, double* dst, int ldd) {
  if (rows <= 0 || cols <= 0 || !dst) return CVXB_OK;
  bool dev = (h.flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)ldd * sizeof(double), s.d, (size_t)s.ld * sizeof(double),
                                 (size_t)rows * sizeof(double), cols, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                 h.stream));
  return CVXB_OK;
}

static KktWork* cached_work(Handle& h, int n, int p, int* status) {
  KktWork* W = (KktWork*)h.kkt_cache;
  if (W && (W->n != n || W->p != p)) { kkt_work_free(*W); delete W; W = nullptr; h.kkt_cache = nullptr; }
  if (!W) {
    W = new KktWork();
    *status = kkt_work_alloc(h, *W, n, p);
    if (*status != CVXB_OK) { kkt_work_free(*W); delete W; return nullptr; }
    h.kkt_cache = W;
  }
  *status = CVXB_OK;
  return W;
}

int prof_begin(Handle& h, int id) {
  if (!h.prof_on || h.capturing || id <= 0 || id >= PROF_COUNT) return CVXB_OK;
  Handle::ProfRange& R = h.prof_range[id];
  if (R.used + 2 > R.ev.size())
    for (int i = 0; i < 64; ++i) {
      cudaEvent_t e;
      CVXB_CUDA_OK(cudaEventCreate(&e));
      R.ev.push_back(e);
    }
  CVXB_CUDA_OK(cudaEventRecord(R.ev[R.used], h.stream));
  return CVXB_OK;
}
int prof_end(Handle& h, int id, double work) {
  if (!h.prof_on || h.capturing || id <= 0 || id >= PROF_COUNT) return CVXB_OK;
  Handle::ProfRange& R = h.prof_range[id];
  if (R.used + 2 > R.ev.size()) return CVXB_OK;
  CVXB_CUDA_OK(cudaEventRecord(R.ev[R.used + 1], h.stream));
  R.used += 2;
  R.work += work;
  return CVXB_OK;
}

// KKTData.reduced (KKTData.scala:68-91): column j is "null" when ||H(:,j)|| + ||A(:,j)|| == 0
__global__ void null_column_kernel(int n, int p, const double* __restrict__ H, int ldh, const double* __restrict__ A, int lda,
                                   int* __restrict__ keep) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { double v = H[(size_t)j * ldh + i]; s = fma(v, v, s); }
  for (int i = threadIdx.x; i < p; i += blockDim.x) { double v = A[(size_t)j * lda + i]; s = fma(v, v, s); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) keep[j] = red[0] > 0.0 ? 1 : 0;
}
// reduced system: Hr = H(I,I), Ar = A(:,I), gr = g(I)
__global__ void gather_reduced_kernel(int nr, int p, const int* __restrict__ I, const double* __restrict__ H, int ldh,
                                      const double* __restrict__ A, int lda, const double* __restrict__ g,
                                      double* __restrict__ Hr, int ldhr, double* __restrict__ Ar, int ldar,
                                      double* __restrict__ gr) {
  const int jr = blockIdx.x, j = I[jr];
  for (int ir = threadIdx.x; ir < nr; ir += blockDim.x) Hr[(size_t)jr * ldhr + ir] = H[(size_t)j * ldh + I[ir]];
  for (int i = threadIdx.x; i < p; i += blockDim.x) Ar[(size_t)jr * ldar + i] = A[(size_t)j * lda + i];
  if (threadIdx.x == 0) gr[jr] = g[j];
}
// KKTData.paddVector (KKTData.scala:105-127): zeros at the eliminated coordinates
__global__ void pad_vector_kernel(int nr, const int* __restrict__ I, const double* __restrict__ xr, double* __restrict__ x) {
  int ir = blockIdx.x * blockDim.x + threadIdx.x;
  if (ir < nr) x[I[ir]] = xr[ir];
}

__global__ void check_symmetric_kernel(int n, const double* __restrict__ Q, int ldq, double* out) {
  // ||Q - Q'||_F^2 partial per block -> atomic-free: one block per column, then summed by block 0 later
  __shared__ double red[256];
  int j = blockIdx.x;
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double d = Q[(size_t)j * ldq + i] - Q[(size_t)i * ldq + j];
    s = fma(d, d, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[j] = red[0];
}

}  // namespace cvxb

using namespace cvxb;

#define CHECK_HANDLE(h)                                                     \
  if (!(h)) { cvxb::set_last_error("null handle"); return CVXB_EINVAL; }   \
  cvxb::DeviceGuard _guard((h)->device)

extern "C" {

const char* cvxb_last_error(void) { return cvxb::g_err; }
const char* cvxb_version(void) { return "cvxb 0.1 (sm_100a, FP64 DMMA)"; }

int cvxb_default_params(cvxb_params* p) {
  if (!p) return CVXB_EINVAL;
  p->maxIter = 1000; p->alpha = 0.04; p->beta = 0.8; p->tolSolver = 1e-8; p->tolEqSolve = 1e-1;
  p->tolFeas = 1e-7; p->delta = 1e-6; p->mu = 10.0; p->t0 = 1.0; p->ruizMaxSweeps = 20; p->ruizTol = 1e-6;
  p->cholRegDelta = 1e-10; p->cholMinDiag = 1e-7; p->newtonRegDelta = 1e-9; p->phase1EqTol = 1e-6;
  p->pdStepFraction = 0.99; p->bugCompat = 0; p->stepLimit = 0;
  return CVXB_OK;
}

int cvxb_create(int device, void* stream, unsigned flags, cvxb_handle* out) {
  if (!out) return CVXB_EINVAL;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cvxb::set_last_error("cvxb_create: no CUDA device visible (this library has no CPU path)");
    return CVXB_ECUDA;
  }
  if (device < 0 || device >= count) { cvxb::set_last_error("cvxb_create: bad device %d", device); return CVXB_EINVAL; }
  cvxb::DeviceGuard _guard(device);
  cudaDeviceProp prop;
  CVXB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    cvxb::set_last_error("cvxb_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                         prop.minor);
    return CVXB_ECUDA;
  }
  cvxb_handle h = new cvxb_handle_s();
  h->device = device;
  h->flags = flags;
  h->sm_count = prop.multiProcessorCount;
  if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
  else {
    // the handle's own stream carries the critical chain of the blocked factorisations (leaf -> panel -> look-ahead
    // update); give it the highest priority so its CTAs are dispatched ahead of the bulk updates queued on stream2
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (getenv("CVXB_NO_PRIO")) hi = 0;
    CVXB_CUDA_OK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));
    h->own_stream = true;
  }
  CVXB_CUDA_OK(cudaMalloc((void**)&h->d_scal, NSCAL * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&h->d_flag, NFLAG * sizeof(int)));
  CVXB_CUDA_OK(cudaMalloc((void**)&h->d_part, PART_DOUBLES * sizeof(double)));
  CVXB_CUDA_OK(cudaMalloc((void**)&h->d_ticket, 16 * sizeof(unsigned) + 64 * sizeof(unsigned long long)));   // + Ruiz rho slots
  CVXB_CUDA_OK(cudaMemset(h->d_scal, 0, NSCAL * sizeof(double)));
  CVXB_CUDA_OK(cudaMemset(h->d_flag, 0, NFLAG * sizeof(int)));
  CVXB_CUDA_OK(cudaMemset(h->d_ticket, 0, 16 * sizeof(unsigned) + 64 * sizeof(unsigned long long)));
  {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (coop && !getenv("CVXB_NO_WAVEFRONT")) CVXB_CUDA_OK(cudaMalloc((void**)&h->wave_ready, 1024 * sizeof(int)));
    if (coop && !getenv("CVXB_NO_STREAMK")) {
      CVXB_CUDA_OK(cudaMalloc((void**)&h->sk_ws, (size_t)h->sm_count * 128 * 128 * sizeof(double)));
      CVXB_CUDA_OK(cudaMalloc((void**)&h->sk_flags, (size_t)h->sm_count * sizeof(int)));
      CVXB_CUDA_OK(cudaMemset(h->sk_flags, 0, (size_t)h->sm_count * sizeof(int)));
    }
  }
  CVXB_CUDA_OK(cudaMalloc((void**)&h->d_prof, 4 * sizeof(unsigned long long)));
  CVXB_CUDA_OK(cudaMemset(h->d_prof, 0, 4 * sizeof(unsigned long long)));
  CVXB_CUDA_OK(cudaMallocHost((void**)&h->h_scal, NSCAL * sizeof(double)));
  CVXB_CUDA_OK(cudaMallocHost((void**)&h->h_flag, NFLAG * sizeof(int)));
  CVXB_CUDA_OK(cudaEventCreate(&h->ev0));
  CVXB_CUDA_OK(cudaEventCreate(&h->ev1));
  {   // keep freed problem memory in the device's default pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      cudaGetLastError();
    }
  }
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    // three lanes: the handle's stream (critical chain, highest priority), stream2 (the look-ahead schedule's own bulk
    // updates, which the chain waits for two leaves later) and stream3 (bulk updates of the tile-DAG schedule, lowest)
    const int mid = (lo - 1 >= hi) ? lo - 1 : lo;
    CVXB_CUDA_OK(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, mid));
    if (!getenv("CVXB_NO_DAG")) {
      CVXB_CUDA_OK(cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, lo));
      for (int i = 0; i < 3; ++i) CVXB_CUDA_OK(cudaStreamCreateWithPriority(&h->side[i], cudaStreamNonBlocking, lo));
      CVXB_CUDA_OK(cudaMalloc((void**)&h->d_part3, DAG_LANES * PART3_DOUBLES * sizeof(double)));
      for (int i = 0; i < 130; ++i) {
        cudaEvent_t e;
        CVXB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->dag_events.push_back(e);
      }
      if (getenv("CVXB_DAG_BLOCK") || getenv("CVXB_DAG_RESERVE")) h->dag_auto = false;
      if (const char* e = getenv("CVXB_DAG_BLOCK")) h->dag_block = atoi(e) / NB * NB;
      if (const char* e = getenv("CVXB_DAG_MIN_N")) h->dag_min_n = atoi(e);
      if (const char* e = getenv("CVXB_DAG_RESERVE")) h->dag_reserve = atoi(e);
    }
  }
  for (int i = 0; i < 600; ++i) {
    cudaEvent_t e;
    CVXB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->la_events.push_back(e);
  }
  CVXB_CUDA_OK(cudaEventCreate(&h->gev0));
  CVXB_CUDA_OK(cudaEventCreate(&h->gev1));
  if (const char* e = getenv("CVXB_NO_GRAPHS")) h->use_graphs = (e[0] == '0' || e[0] == 0) ? 1 : 0;
  if (const char* e = getenv("CVXB_NO_LOOP")) h->use_loop = (e[0] == '0' || e[0] == 0) ? 1 : 0;
  if (!h->use_graphs) h->use_loop = 0;
  CVXB_TRY(gemm_dmma_init());
  CVXB_TRY(gemm_dmma_build_tile_orders(*h));
  CVXB_TRY(factor_init());
  *out = h;
  return CVXB_OK;
}

int cvxb_destroy(cvxb_handle h) {
  if (!h) return CVXB_OK;
  cvxb::DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->kkt_cache) { KktWork* W = (KktWork*)h->kkt_cache; kkt_work_free(*W); delete W; }
  cudaFree(h->sk_ws); cudaFree(h->sk_flags); cudaFree(h->tile_order);
  cudaFree(h->d_prof);
  cudaFree(h->wave_ready); cudaFree(h->d_scal); cudaFree(h->d_flag); cudaFree(h->d_part); cudaFree(h->d_ticket);
  cudaFreeHost(h->h_scal); cudaFreeHost(h->h_flag);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->gev0); cudaEventDestroy(h->gev1);
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  for (auto& R : h->prof_range) for (cudaEvent_t e : R.ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->la_events) cudaEventDestroy(e);
  for (cudaEvent_t e : h->dag_events) cudaEventDestroy(e);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream3) cudaStreamDestroy(h->stream3);
  for (int i = 0; i < 3; ++i) if (h->side[i]) cudaStreamDestroy(h->side[i]);
  cudaFree(h->d_part3);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return CVXB_OK;
}

int cvxb_synchronize(cvxb_handle h) {
  CHECK_HANDLE(h);
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

long long cvxb_launch_count(cvxb_handle h) { return h ? h->launches : 0; }
long long cvxb_status_read_count(cvxb_handle h) { return h ? h->status_reads : 0; }

int cvxb_profile_enable(cvxb_handle h, int on) {
  CHECK_HANDLE(h);
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->prof_on = on;
  h->prof_used = 0;
  h->prof_flops = 0.0;
  h->prof_ms_graph = 0.0;
  h->prof_launches_graph = 0;
  CVXB_CUDA_OK(cudaMemsetAsync(h->d_prof, 0, 4 * sizeof(unsigned long long), h->stream));
  for (auto& R : h->prof_range) { R.used = 0; R.work = 0.0; }
  return CVXB_OK;
}

int cvxb_profile_read_range(cvxb_handle h, int range, long long* count, double* ms_total, double* work_total) {
  CHECK_HANDLE(h);
  if (range == PROF_HESSIAN) return cvxb_profile_read(h, count, ms_total, work_total);
  if (range < 0 || range >= PROF_COUNT) { cvxb::set_last_error("cvxb_profile_read_range: unknown range %d", range); return CVXB_EINVAL; }
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  Handle::ProfRange& R = h->prof_range[range];
  double ms = 0;
  for (size_t i = 0; i + 1 < R.used; i += 2) {
    float t = 0;
    CVXB_CUDA_OK(cudaEventElapsedTime(&t, R.ev[i], R.ev[i + 1]));
    ms += t;
  }
  if (count) *count = (long long)(R.used / 2);
  if (ms_total) *ms_total = ms;
  if (work_total) *work_total = R.work;
  return CVXB_OK;
}

int cvxb_profile_read(cvxb_handle h, long long* launches, double* ms_total, double* flops_total) {
  CHECK_HANDLE(h);
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  double ms = 0;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    float t = 0;
    CVXB_CUDA_OK(cudaEventElapsedTime(&t, h->prof_events[i], h->prof_events[i + 1]));
    ms += t;
  }
  unsigned long long dp[4] = {0, 0, 0, 0};       // SYRKs timed inside device-driven stage loops
  CVXB_CUDA_OK(cudaMemcpy(dp, h->d_prof, sizeof(dp), cudaMemcpyDeviceToHost));
  if (launches) *launches = (long long)(h->prof_used / 2) + h->prof_launches_graph + (long long)dp[2];
  if (ms_total) *ms_total = ms + h->prof_ms_graph + (double)dp[1] * 1e-6;
  if (flops_total) *flops_total = h->prof_flops;
  return CVXB_OK;
}

// ------------------------------------------------------------------------------------ seam B
int cvxb_kkt_solve(cvxb_handle h, int n, int p, const double* H, int ldh, const double* A, int lda, const double* q,
                   const double* b, double tol, double* x, double* w, cvxb_kkt_info* info) {
  CHECK_HANDLE(h);
  if (n < 1 || p < 1) { cvxb::set_last_error("cvxb_kkt_solve: need n >= 1 and p >= 1 (got %d, %d)", n, p); return CVXB_EDIM; }
  if (!x || !w) { cvxb::set_last_error("cvxb_kkt_solve: null output"); return CVXB_EINVAL; }
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, p, &st);
  if (!W) return st;
  Staged dH, dA, dq, db, dx, dw;
  CVXB_TRY(stage_in(*h, n, n, H, ldh, dH));
  CVXB_TRY(stage_in(*h, p, n, A, lda, dA));
  CVXB_TRY(stage_in(*h, n, 1, q, n, dq));
  CVXB_TRY(stage_in(*h, p, 1, b, p, db));
  CVXB_TRY(stage_out_alloc(*h, n, 1, dx));
  CVXB_TRY(stage_out_alloc(*h, p, 1, dw));
  st = kkt_solve_device(*h, *W, P, dH.d, dH.ld, dA.d, dA.ld, dq.d, db.d, tol, dx.d, dw.d, info);
  if (st != CVXB_OK) return st;
  CVXB_TRY(copy_out(*h, n, 1, dx, x, n));
  CVXB_TRY(copy_out(*h, p, 1, dw, w, p));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_kkt_solve_reduced(cvxb_handle h, int n, int p, const double* H, int ldh, const double* A, int lda, const double* g,
                           const double* r, double tol, double* x, double* w, int* null_indices, int* n_null,
                           cvxb_kkt_info* info) {
  // KKTData(H,A,g,r).reduced -> KKTSystem(rH,rA,rg,r).solve -> KKTData.paddVector   (KKTData.scala:68-127, KktTest.scala:52-104)
  CHECK_HANDLE(h);
  if (n < 1 || p < 1) { cvxb::set_last_error("cvxb_kkt_solve_reduced: need n >= 1 and p >= 1 (got %d, %d)", n, p); return CVXB_EDIM; }
  if (!x || !w || !g) { cvxb::set_last_error("cvxb_kkt_solve_reduced: null argument"); return CVXB_EINVAL; }
  if (h->flags & CVXB_FLAG_DEVICE_PTRS) { cvxb::set_last_error("cvxb_kkt_solve_reduced takes host pointers"); return CVXB_EINVAL; }
  cvxb_params P;
  cvxb_default_params(&P);
  Staged dH, dA, dg, db;
  CVXB_TRY(stage_in(*h, n, n, H, ldh, dH));
  CVXB_TRY(stage_in(*h, p, n, A, lda, dA));
  CVXB_TRY(stage_in(*h, n, 1, g, n, dg));
  CVXB_TRY(stage_in(*h, p, 1, r, p, db));
  int* d_keep = nullptr;
  CVXB_CUDA_OK(cudaMalloc((void**)&d_keep, sizeof(int) * 2 * (size_t)n));
  struct Free { int* p; ~Free() { cudaFree(p); } } guard{d_keep};
  int* d_idx = d_keep + n;
  CVXB_LAUNCH(*h, null_column_kernel, n, 256, 0, n, p, dH.d, dH.ld, dA.d, dA.ld, d_keep);
  std::vector<int> keep((size_t)n), idx;
  CVXB_CUDA_OK(cudaMemcpyAsync(keep.data(), d_keep, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  int nn = 0;
  for (int j = 0; j < n; ++j) {
    if (keep[j]) { idx.push_back(j); continue; }
    if (!(fabs(g[j]) < 1e-15)) {
      cvxb::set_last_error("Unsolvable KKT system, row %d is zero with nonzero right hand side.", j);
      return CVXB_EUNSOLVABLE;
    }
    if (null_indices) null_indices[nn] = j;
    ++nn;
  }
  if (n_null) *n_null = nn;
  const int nr = (int)idx.size();
  if (nr < 1) { cvxb::set_last_error("cvxb_kkt_solve_reduced: every row of the system is zero"); return CVXB_EUNSOLVABLE; }
  int st;
  KktWork* W = cached_work(*h, nr, p, &st);
  if (!W) return st;
  Staged dx, dw;
  CVXB_TRY(stage_out_alloc(*h, nr, 1, dx));
  CVXB_TRY(stage_out_alloc(*h, p, 1, dw));
  if (nn == 0) {
    st = kkt_solve_device(*h, *W, P, dH.d, dH.ld, dA.d, dA.ld, dg.d, db.d, tol, dx.d, dw.d, info);
    if (st != CVXB_OK) return st;
    CVXB_TRY(copy_out(*h, n, 1, dx, x, n));
  } else {
    Staged rH, rA, rg, xp;
    CVXB_TRY(stage_out_alloc(*h, nr, nr, rH));
    CVXB_TRY(stage_out_alloc(*h, p, nr, rA));
    CVXB_TRY(stage_out_alloc(*h, nr, 1, rg));
    CVXB_TRY(stage_out_alloc(*h, n, 1, xp));
    CVXB_CUDA_OK(cudaMemsetAsync(rH.d, 0, sizeof(double) * (size_t)rH.ld * nr, h->stream));
    CVXB_CUDA_OK(cudaMemsetAsync(rA.d, 0, sizeof(double) * (size_t)rA.ld * nr, h->stream));
    CVXB_CUDA_OK(cudaMemsetAsync(rg.d, 0, sizeof(double) * (size_t)rg.ld, h->stream));
    CVXB_CUDA_OK(cudaMemsetAsync(xp.d, 0, sizeof(double) * (size_t)xp.ld, h->stream));
    CVXB_CUDA_OK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, h->stream));
    CVXB_LAUNCH(*h, gather_reduced_kernel, nr, 256, 0, nr, p, d_idx, dH.d, dH.ld, dA.d, dA.ld, dg.d, rH.d, rH.ld, rA.d, rA.ld, rg.d);
    st = kkt_solve_device(*h, *W, P, rH.d, rH.ld, rA.d, rA.ld, rg.d, db.d, tol, dx.d, dw.d, info);
    if (st != CVXB_OK) return st;
    CVXB_LAUNCH(*h, pad_vector_kernel, (nr + 255) / 256, 256, 0, nr, d_idx, dx.d, xp.d);
    CVXB_TRY(copy_out(*h, n, 1, xp, x, n));
  }
  CVXB_TRY(copy_out(*h, p, 1, dw, w, p));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

struct cvxb_solution_space_s : cvxb::SolutionSpaceDev {};

int cvxb_solution_space_create(cvxb_handle h, int p, int n, const double* A, int lda, const double* b,
                               cvxb_solution_space* out) {
  CHECK_HANDLE(h);
  if (!A || !b || !out) { cvxb::set_last_error("cvxb_solution_space_create: null argument"); return CVXB_EINVAL; }
  if (!(p >= 1 && p < n)) { cvxb::set_last_error("SolutionSpace: need 1 <= A.rows < A.cols (got %d x %d)", p, n); return CVXB_EDIM; }
  Staged dA, db;
  CVXB_TRY(stage_in(*h, p, n, A, lda, dA));
  CVXB_TRY(stage_in(*h, p, 1, b, p, db));
  SolutionSpaceDev* S = nullptr;
  CVXB_TRY(solution_space_build(*h, p, n, dA.d, dA.ld, db.d, &S));
  *out = (cvxb_solution_space)S;
  return CVXB_OK;
}

int cvxb_solution_space_from_basis(cvxb_handle h, int n, int k, const double* z0, const double* F, int ldf,
                                   cvxb_solution_space* out) {
  CHECK_HANDLE(h);
  if (!z0 || !F || !out) { cvxb::set_last_error("cvxb_solution_space_from_basis: null argument"); return CVXB_EINVAL; }
  if (!(k >= 1 && k <= n)) { cvxb::set_last_error("affineTransformed: need 1 <= F.cols <= F.rows (got %d x %d)", n, k); return CVXB_EDIM; }
  if (ldf < n) { cvxb::set_last_error("leading dimension %d < rows %d", ldf, n); return CVXB_EDIM; }
  SolutionSpaceDev* S = new SolutionSpaceDev();
  S->n = n; S->p = n - k; S->ldq = pad_ld(n);
  S->device = h->device; S->stream = h->stream;
  const bool dev = (h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  cudaError_t e = cudaMalloc((void**)&S->Fext, sizeof(double) * (size_t)S->ldq * k);
  if (e == cudaSuccess) e = cudaMalloc((void**)&S->z0, sizeof(double) * (size_t)S->ldq);
  if (e == cudaSuccess) e = cudaMalloc((void**)&S->tmp, sizeof(double) * 2 * (size_t)S->ldq);
  if (e == cudaSuccess) e = cudaMemsetAsync(S->Fext, 0, sizeof(double) * (size_t)S->ldq * k, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(S->z0, 0, sizeof(double) * (size_t)S->ldq, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(S->Fext, sizeof(double) * (size_t)S->ldq, F, sizeof(double) * (size_t)ldf, sizeof(double) * (size_t)n, k,
                          kind, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S->z0, z0, sizeof(double) * (size_t)n, kind, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    cvxb::set_last_error("CUDA error %s in cvxb_solution_space_from_basis", cudaGetErrorString(e));
    solution_space_free(S);
    return CVXB_ECUDA;
  }
  *out = (cvxb_solution_space)S;
  return CVXB_OK;
}

int cvxb_solution_space_destroy(cvxb_solution_space space) {
  if (space) {
    SolutionSpaceDev* S = (SolutionSpaceDev*)space;
    cvxb::DeviceGuard _guard(S->device);
    if (S->stream) cudaStreamSynchronize(S->stream); else cudaDeviceSynchronize();
    solution_space_free(S);
  }
  return CVXB_OK;
}

int cvxb_solution_space_get(cvxb_handle h, cvxb_solution_space space, double* z0, double* F, int ldf) {
  CHECK_HANDLE(h);
  if (!space) { cvxb::set_last_error("null solution space"); return CVXB_EINVAL; }
  SolutionSpaceDev* S = (SolutionSpaceDev*)space;
  if (F && ldf < S->n) { cvxb::set_last_error("leading dimension %d < rows %d", ldf, S->n); return CVXB_EDIM; }
  Staged sz, sf;
  sz.d = S->z0; sz.ld = pad_ld(S->n);
  sf.d = S->F(); sf.ld = S->ldq;
  CVXB_TRY(copy_out(*h, S->n, 1, sz, z0, S->n));
  CVXB_TRY(copy_out(*h, S->n, S->k(), sf, F, ldf));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_solution_space_parameter(cvxb_handle h, cvxb_solution_space space, const double* x, double* u) {
  CHECK_HANDLE(h);
  if (!space || !x || !u) { cvxb::set_last_error("cvxb_solution_space_parameter: null argument"); return CVXB_EINVAL; }
  SolutionSpaceDev* S = (SolutionSpaceDev*)space;
  Staged dx, du;
  CVXB_TRY(stage_in(*h, S->n, 1, x, S->n, dx));
  CVXB_TRY(stage_out_alloc(*h, S->k(), 1, du));
  CVXB_TRY(solution_space_parameter(*h, S, dx.d, du.d));
  CVXB_TRY(copy_out(*h, S->k(), 1, du, u, S->k()));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_solution_space_map(cvxb_handle h, cvxb_solution_space space, const double* u, double* x) {
  CHECK_HANDLE(h);
  if (!space || !x || !u) { cvxb::set_last_error("cvxb_solution_space_map: null argument"); return CVXB_EINVAL; }
  SolutionSpaceDev* S = (SolutionSpaceDev*)space;
  Staged dx, du;
  CVXB_TRY(stage_in(*h, S->k(), 1, u, S->k(), du));
  CVXB_TRY(stage_out_alloc(*h, S->n, 1, dx));
  CVXB_TRY(solution_space_map(*h, S, du.d, dx.d));
  CVXB_TRY(copy_out(*h, S->n, 1, dx, x, S->n));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_solve_underdetermined(cvxb_handle h, int p, int n, const double* A, int lda, const double* b, double* z0,
                               double* F, int ldf) {
  cvxb_solution_space sp = nullptr;
  CVXB_TRY(cvxb_solution_space_create(h, p, n, A, lda, b, &sp));
  int st = cvxb_solution_space_get(h, sp, z0, F, ldf);
  cvxb_solution_space_destroy(sp);
  return st;
}

int cvxb_cholesky_solve(cvxb_handle h, int n, const double* H, int ldh, const double* b, double tol, double* x,
                        cvxb_kkt_info* info) {
  CHECK_HANDLE(h);
  if (n < 1) { cvxb::set_last_error("cvxb_cholesky_solve: n = %d", n); return CVXB_EDIM; }
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, 0, &st);
  if (!W) return st;
  Staged dH, db, dx;
  CVXB_TRY(stage_in(*h, n, n, H, ldh, dH));
  CVXB_TRY(stage_in(*h, n, 1, b, n, db));
  CVXB_TRY(stage_out_alloc(*h, n, 1, dx));
  st = chol_solve_device(*h, *W, P, dH.d, dH.ld, db.d, 1.0, tol, dx.d, info);
  if (st != CVXB_OK) return st;
  CVXB_TRY(copy_out(*h, n, 1, dx, x, n));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

// ---- seam B with pinned, double-buffered staging (general objectives: the Hessian is assembled on the host) ----------
// The caller fills one pinned buffer with a block of columns of H while the previous block is still on its way to the
// device on a copy stream of its own; the solves then run on the device-resident matrix.
struct cvxb_stage_s {
  cvxb_handle_s* h = nullptr;
  int n = 0, p = 0, ldn = 0, ldp = 0, block_cols = 0;
  double *dH = nullptr, *dA = nullptr, *dvec = nullptr;     // device: H (ldn x n), A (ldp x n), vectors q | b | x | w
  double* pin[2] = {nullptr, nullptr};                      // pinned host buffers, n x block_cols each (leading dimension n)
  cudaEvent_t done[2] = {nullptr, nullptr};                 // buffer i may be overwritten once done[i] has completed
  cudaEvent_t all = nullptr;
  cudaStream_t copy = nullptr;
  long long pushed_cols = 0;
  KktWork W;
};

static void stage_free(cvxb_stage_s* S) {
  if (!S) return;
  cudaFree(S->dH); cudaFree(S->dA); cudaFree(S->dvec);
  for (int i = 0; i < 2; ++i) { if (S->pin[i]) cudaFreeHost(S->pin[i]); if (S->done[i]) cudaEventDestroy(S->done[i]); }
  if (S->all) cudaEventDestroy(S->all);
  if (S->copy) cudaStreamDestroy(S->copy);
  kkt_work_free(S->W);
  delete S;
}

int cvxb_stage_create(cvxb_handle h, int n, int p, int block_cols, cvxb_stage* out) {
  CHECK_HANDLE(h);
  if (!out) { cvxb::set_last_error("cvxb_stage_create: null argument"); return CVXB_EINVAL; }
  if (n < 1 || p < 0) { cvxb::set_last_error("cvxb_stage_create: need n >= 1, p >= 0 (got %d, %d)", n, p); return CVXB_EDIM; }
  if (block_cols < 1) block_cols = n < 256 ? n : 256;
  if (block_cols > n) block_cols = n;
  cvxb_stage_s* S = new cvxb_stage_s();
  S->h = h; S->n = n; S->p = p; S->ldn = pad_ld(n); S->ldp = pad_ld(p); S->block_cols = block_cols;
  cudaError_t e = cudaMalloc((void**)&S->dH, sizeof(double) * (size_t)S->ldn * n);
  if (e == cudaSuccess) e = cudaMemsetAsync(S->dH, 0, sizeof(double) * (size_t)S->ldn * n, h->stream);
  if (e == cudaSuccess && p > 0) e = cudaMalloc((void**)&S->dA, sizeof(double) * (size_t)S->ldp * n);
  if (e == cudaSuccess && p > 0) e = cudaMemsetAsync(S->dA, 0, sizeof(double) * (size_t)S->ldp * n, h->stream);
  if (e == cudaSuccess) e = cudaMalloc((void**)&S->dvec, sizeof(double) * (2 * (size_t)S->ldn + 2 * (size_t)S->ldp));
  if (e == cudaSuccess) e = cudaMemsetAsync(S->dvec, 0, sizeof(double) * (2 * (size_t)S->ldn + 2 * (size_t)S->ldp), h->stream);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMallocHost((void**)&S->pin[i], sizeof(double) * (size_t)n * block_cols);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S->done[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S->all, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&S->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  int st = e == cudaSuccess ? kkt_work_alloc(*h, S->W, n, p) : CVXB_ECUDA;
  if (st != CVXB_OK) {
    if (e != cudaSuccess) cvxb::set_last_error("CUDA error %s in cvxb_stage_create", cudaGetErrorString(e));
    stage_free(S);
    return st;
  }
  *out = S;
  return CVXB_OK;
}

int cvxb_stage_destroy(cvxb_stage S) {
  if (!S) return CVXB_OK;
  cvxb::DeviceGuard guard(S->h->device);
  cudaStreamSynchronize(S->copy);
  cudaStreamSynchronize(S->h->stream);
  stage_free(S);
  return CVXB_OK;
}

int cvxb_stage_buffer(cvxb_stage S, int which, double** host_buffer, int* block_cols) {
  if (!S || which < 0 || which > 1 || !host_buffer) { cvxb::set_last_error("cvxb_stage_buffer: bad argument"); return CVXB_EINVAL; }
  *host_buffer = S->pin[which];
  if (block_cols) *block_cols = S->block_cols;
  return CVXB_OK;
}

int cvxb_stage_wait(cvxb_stage S, int which) {
  if (!S || which < 0 || which > 1) { cvxb::set_last_error("cvxb_stage_wait: bad argument"); return CVXB_EINVAL; }
  cvxb::DeviceGuard guard(S->h->device);
  CVXB_CUDA_OK(cudaEventSynchronize(S->done[which]));
  return CVXB_OK;
}

int cvxb_stage_push(cvxb_stage S, int which, int col0, int ncols) {
  if (!S || which < 0 || which > 1) { cvxb::set_last_error("cvxb_stage_push: bad argument"); return CVXB_EINVAL; }
  if (col0 < 0 || ncols < 1 || ncols > S->block_cols || col0 + ncols > S->n) {
    cvxb::set_last_error("cvxb_stage_push: columns %d..%d outside 0..%d or more than %d at once", col0, col0 + ncols, S->n, S->block_cols);
    return CVXB_EDIM;
  }
  cvxb::DeviceGuard guard(S->h->device);
  if (S->pushed_cols == 0)      // first block of a new matrix: the previous solve on the handle's stream must have finished with H
    CVXB_CUDA_OK(cudaStreamSynchronize(S->h->stream));
  CVXB_CUDA_OK(cudaMemcpy2DAsync(S->dH + (size_t)col0 * S->ldn, sizeof(double) * (size_t)S->ldn, S->pin[which],
                                 sizeof(double) * (size_t)S->n, sizeof(double) * (size_t)S->n, ncols, cudaMemcpyHostToDevice,
                                 S->copy));
  CVXB_CUDA_OK(cudaEventRecord(S->done[which], S->copy));
  S->pushed_cols += ncols;
  return CVXB_OK;
}

// the solves wait for the copy stream on the device (no host stall), then run on the handle's stream
static int stage_join(cvxb_stage_s* S) {
  CVXB_CUDA_OK(cudaEventRecord(S->all, S->copy));
  CVXB_CUDA_OK(cudaStreamWaitEvent(S->h->stream, S->all, 0));
  S->pushed_cols = 0;
  return CVXB_OK;
}

int cvxb_stage_cholesky_solve(cvxb_stage S, const double* b, double tol, double* x, cvxb_kkt_info* info) {
  if (!S || !b || !x) { cvxb::set_last_error("cvxb_stage_cholesky_solve: null argument"); return CVXB_EINVAL; }
  cvxb_handle h = S->h;
  cvxb::DeviceGuard guard(h->device);
  cvxb_params P;
  cvxb_default_params(&P);
  double *db = S->dvec, *dx = S->dvec + S->ldn;
  CVXB_TRY(stage_join(S));
  CVXB_CUDA_OK(cudaMemcpyAsync(db, b, sizeof(double) * (size_t)S->n, cudaMemcpyHostToDevice, h->stream));
  int st = chol_solve_device(*h, S->W, P, S->dH, S->ldn, db, 1.0, tol, dx, info);
  if (st != CVXB_OK) return st;
  CVXB_CUDA_OK(cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)S->n, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_stage_set_equalities(cvxb_stage S, const double* A, int lda) {
  if (!S || !A || S->p < 1) { cvxb::set_last_error("cvxb_stage_set_equalities: no equalities in this stage"); return CVXB_EINVAL; }
  if (lda < S->p) { cvxb::set_last_error("leading dimension %d < rows %d", lda, S->p); return CVXB_EDIM; }
  cvxb::DeviceGuard guard(S->h->device);
  CVXB_CUDA_OK(cudaMemcpy2DAsync(S->dA, sizeof(double) * (size_t)S->ldp, A, sizeof(double) * (size_t)lda,
                                 sizeof(double) * (size_t)S->p, S->n, cudaMemcpyHostToDevice, S->h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(S->h->stream));
  return CVXB_OK;
}

int cvxb_stage_kkt_solve(cvxb_stage S, const double* q, const double* b, double tol, double* x, double* w,
                         cvxb_kkt_info* info) {
  if (!S || !q || !b || !x || !w) { cvxb::set_last_error("cvxb_stage_kkt_solve: null argument"); return CVXB_EINVAL; }
  if (S->p < 1) { cvxb::set_last_error("cvxb_stage_kkt_solve: stage created with p = 0"); return CVXB_EDIM; }
  cvxb_handle h = S->h;
  cvxb::DeviceGuard guard(h->device);
  cvxb_params P;
  cvxb_default_params(&P);
  double *dq = S->dvec, *dx = S->dvec + S->ldn, *db = S->dvec + 2 * (size_t)S->ldn, *dw = db + S->ldp;
  CVXB_TRY(stage_join(S));
  CVXB_CUDA_OK(cudaMemcpyAsync(dq, q, sizeof(double) * (size_t)S->n, cudaMemcpyHostToDevice, h->stream));
  CVXB_CUDA_OK(cudaMemcpyAsync(db, b, sizeof(double) * (size_t)S->p, cudaMemcpyHostToDevice, h->stream));
  int st = kkt_solve_device(*h, S->W, P, S->dH, S->ldn, S->dA, S->ldp, dq, db, tol, dx, dw, info);
  if (st != CVXB_OK) return st;
  CVXB_CUDA_OK(cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)S->n, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaMemcpyAsync(w, dw, sizeof(double) * (size_t)S->p, cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_kkt_solve_with_chol_factor(cvxb_handle h, int n, int p, const double* L, int ldl, const double* A, int lda,
                                    const double* q, const double* b, double tol, double* x, double* w,
                                    cvxb_kkt_info* info) {
  // KKTSystem.solveWithCholFactor (KKTSystem.scala:99-167) with a caller-supplied factor: the same
  // block elimination as kkt_enqueue, minus equilibration and factorisation.
  CHECK_HANDLE(h);
  if (n < 1 || p < 1) { cvxb::set_last_error("solveWithCholFactor: need n >= 1 and p >= 1"); return CVXB_EDIM; }
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, p, &st);
  if (!W) return st;
  Staged dL, dA, dq, db, dx, dw;
  CVXB_TRY(stage_in(*h, n, n, L, ldl, dL));
  CVXB_TRY(stage_in(*h, p, n, A, lda, dA));
  CVXB_TRY(stage_in(*h, n, 1, q, n, dq));
  CVXB_TRY(stage_in(*h, p, 1, b, p, db));
  CVXB_TRY(stage_out_alloc(*h, n, 1, dx));
  CVXB_TRY(stage_out_alloc(*h, p, 1, dw));
  // one attempt with the caller's factor; LinSolveException when the residual test or the Schur Cholesky fails
  CVXB_TRY(kkt_enqueue(*h, *W, P, dL.d, dL.ld, dA.d, dA.ld, dq.d, db.d, tol, true, true, dx.d, dw.d, true));
  CVXB_TRY(fetch_status(*h));
  fill_info(*h, info, 0, 0);
  if (h->h_flag[F_ZERO_DIAG]) { cvxb::set_last_error("solveWithCholFactor: zero on the diagonal of L"); return CVXB_ELINSOLVE; }
  if (h->h_flag[F_BAD]) {
    cvxb::set_last_error("solveWithCholFactor: Error in solution exceeds tolerance (err1 %.3g, err2 %.3g, Schur info %d)",
                         h->h_scal[S_ERR1], h->h_scal[S_ERR2], h->h_flag[F_CHOL_S]);
    return CVXB_ELINSOLVE;
  }
  CVXB_TRY(copy_out(*h, n, 1, dx, x, n));
  CVXB_TRY(copy_out(*h, p, 1, dw, w, p));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_ruiz_equilibrate(cvxb_handle h, int n, const double* H, int ldh, double* d, double* Q, int ldq, int* sweeps) {
  CHECK_HANDLE(h);
  if (n < 1) return CVXB_EDIM;
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, 0, &st);
  if (!W) return st;
  Staged dH, dQ;
  CVXB_TRY(stage_in(*h, n, n, H, ldh, dH));
  CVXB_TRY(ruiz_equilibrate(*h, n, dH.d, dH.ld, W->dr, W->colsq, P.ruizMaxSweeps, P.ruizTol, W->L, (size_t)W->ldn * n));
  if (Q) {
    CVXB_TRY(stage_out_alloc(*h, n, n, dQ));
    CVXB_TRY(scaled_full(*h, n, dH.d, dH.ld, W->dr, dQ.d, dQ.ld));
    CVXB_TRY(copy_out(*h, n, n, dQ, Q, ldq));
  }
  if (d) { Staged sd; sd.d = W->dr; sd.ld = W->ldn; CVXB_TRY(copy_out(*h, n, 1, sd, d, n)); }
  CVXB_TRY(fetch_status(*h));
  if (sweeps) *sweeps = h->h_flag[F_RUIZ_SWEEPS];
  return CVXB_OK;
}

int cvxb_regularized_cholesky(cvxb_handle h, int n, const double* Q, int ldq, double* L, int ldl, cvxb_kkt_info* info) {
  CHECK_HANDLE(h);
  if (n < 1) return CVXB_EDIM;
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, 0, &st);
  if (!W) return st;
  Staged dQ;
  CVXB_TRY(stage_in(*h, n, n, Q, ldq, dQ));
  int reg = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CVXB_TRY(scaled_lower(*h, n, dQ.d, dQ.ld, nullptr, attempt ? P.cholRegDelta : 0.0, W->L, W->ldn));
    CVXB_TRY(potrf_lower(*h, n, W->L, W->ldn, W->invD, F_CHOL_H, S_MINDIAG_H));
    CVXB_TRY(fetch_status(*h));
    bool fail = h->h_flag[F_CHOL_H] != 0;
    if (attempt == 0 && (fail || !(h->h_scal[S_MINDIAG_H] > P.cholMinDiag))) { reg = 1; continue; }
    fill_info(*h, info, 0, reg);
    if (fail) {
      cvxb::set_last_error("regularizedCholesky: not positive definite at column %d", h->h_flag[F_CHOL_H]);
      return CVXB_ELINSOLVE;
    }
    break;
  }
  Staged sl; sl.d = W->L; sl.ld = W->ldn;
  CVXB_TRY(copy_out(*h, n, n, sl, L, ldl));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

int cvxb_triangular_solve(cvxb_handle h, char uplo, int n, int nrhs, const double* T, int ldt, double* B, int ldb) {
  CHECK_HANDLE(h);
  if (n < 1 || nrhs < 1) return CVXB_EDIM;
  if (uplo != 'L' && uplo != 'U') { cvxb::set_last_error("triangularSolve: uplo must be 'L' or 'U'"); return CVXB_EINVAL; }
  int st;
  KktWork* W = cached_work(*h, n, 0, &st);
  if (!W) return st;
  Staged dT, dB;
  CVXB_TRY(stage_in(*h, n, n, T, ldt, dT));
  // private padded copy of B (it is overwritten)
  CVXB_TRY(stage_out_alloc(*h, n, nrhs, dB));
  bool dev = (h->flags & CVXB_FLAG_DEVICE_PTRS) != 0;
  CVXB_CUDA_OK(cudaMemcpy2DAsync(dB.d, (size_t)dB.ld * sizeof(double), B, (size_t)ldb * sizeof(double),
                                 (size_t)n * sizeof(double), nrhs, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 h->stream));
  if (uplo == 'L') {
    CVXB_TRY(scaled_lower(*h, n, dT.d, dT.ld, nullptr, 0.0, W->L, W->ldn));
  } else {
    // U x = b with U upper  <=>  (U')' x = b: store L = U' (lower) and solve with the transposed kernel
    CVXB_TRY(transpose_scale(*h, n, n, dT.d, dT.ld, nullptr, W->L, W->ldn));
    CVXB_TRY(scaled_lower(*h, n, W->L, W->ldn, nullptr, 0.0, W->L, W->ldn));
  }
  CVXB_TRY(invert_diag_blocks(*h, n, W->L, W->ldn, W->invD));
  CVXB_TRY(trsm_lower(*h, n, nrhs, W->L, W->ldn, W->invD, dB.d, dB.ld, uplo == 'U'));
  CVXB_TRY(copy_out(*h, n, nrhs, dB, B, ldb));
  CVXB_TRY(fetch_status(*h));
  if (h->h_flag[F_ZERO_DIAG]) {
    cvxb::set_last_error("triangularSolve: singular triangular matrix (zero on the diagonal)");
    return CVXB_ELINSOLVE;
  }
  return CVXB_OK;
}

int cvxb_symmetric_solve(cvxb_handle h, int n, const double* Hm, int ldh, const double* r, double tol, double* x,
                         cvxb_kkt_info* info) {
  // SymmetricLinearSystem (SymmetricLinearSystem.scala:15-56): Ruiz at construction, then
  // choleskySolve(Q, d o r) (which equilibrates again, defect D6), x = d o u.
  CHECK_HANDLE(h);
  if (n < 1) return CVXB_EDIM;
  cvxb_params P;
  cvxb_default_params(&P);
  int st;
  KktWork* W = cached_work(*h, n, 0, &st);
  if (!W) return st;
  Staged dH, dr, dx, dQ, du;
  CVXB_TRY(stage_in(*h, n, n, Hm, ldh, dH));
  CVXB_TRY(stage_in(*h, n, 1, r, n, dr));
  CVXB_TRY(stage_out_alloc(*h, n, 1, dx));
  CVXB_TRY(stage_out_alloc(*h, n, 1, du));
  CVXB_TRY(stage_out_alloc(*h, n, n, dQ));
  CVXB_TRY(ruiz_equilibrate(*h, n, dH.d, dH.ld, W->dr2, W->colsq, P.ruizMaxSweeps, P.ruizTol, W->L, (size_t)W->ldn * n));
  CVXB_TRY(scaled_full(*h, n, dH.d, dH.ld, W->dr2, dQ.d, dQ.ld));
  // checkSymmetric(Q, 1e-13)  (SymmetricLinearSystem.scala:28)
  CVXB_LAUNCH(*h, check_symmetric_kernel, n, 256, 0, n, dQ.d, dQ.ld, W->t3);
  std::vector<double> colerr(n);
  CVXB_CUDA_OK(cudaMemcpyAsync(colerr.data(), W->t3, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  double asym = 0;
  for (double v : colerr) asym += v;
  const bool symmetric = std::sqrt(asym) < 1e-13;
  // s = d o r ; u = choleskySolve(Q, s) ; x = d o u
  Staged ds;
  CVXB_TRY(stage_out_alloc(*h, n, 1, ds));
  CVXB_TRY(scale_rows(*h, n, 1, dr.d, pad_ld(n), W->dr2, ds.d, ds.ld, false));
  if (!symmetric) {
    st = svd_solve_device(*h, n, dQ.d, dQ.ld, ds.d, 1.0, tol, du.d, nullptr, false);     // svdSolve branch (:29)
    if (info) { memset(info, 0, sizeof(*info)); info->path = 3; }
  } else {
    st = chol_solve_device(*h, *W, P, dQ.d, dQ.ld, ds.d, 1.0, tol, du.d, info);
    if (st == CVXB_ELINSOLVE) {                                                    // symSolve fallback (:33)
      st = svd_solve_device(*h, n, dQ.d, dQ.ld, ds.d, 1.0, tol, du.d, nullptr, true);
      if (info) info->path = 2;
    }
  }
  if (st != CVXB_OK) return st;
  CVXB_TRY(scale_rows(*h, n, 1, du.d, du.ld, W->dr2, dx.d, dx.ld, false));
  CVXB_TRY(copy_out(*h, n, 1, dx, x, n));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

// ------------------------------------------------------------------------- test / bench helpers
int cvxb_test_dgemm(cvxb_handle h, int a_kc, int b_kc, int M, int N, int K, double alpha, const double* A, int lda,
                    const double* B, int ldb, double beta, double* C, int ldc, int tri) {
  CHECK_HANDLE(h);
  Staged dA, dB, dC;
  // A is stored (K-contiguous ? M x K with K fastest : K x M ... ) as a column-major array with `lda`:
  //   a_kc: lda >= K, M columns;  else lda >= M, K columns
  CVXB_TRY(stage_in(*h, a_kc ? K : M, a_kc ? M : K, A, lda, dA));
  CVXB_TRY(stage_in(*h, b_kc ? K : N, b_kc ? N : K, B, ldb, dB));
  CVXB_TRY(stage_in(*h, M, N, C, ldc, dC));
  Staged dC2;
  double* cptr = dC.d;
  int cld = dC.ld;
  if (!dC.owned) {   // never write into the caller's array in place during staging
    CVXB_TRY(stage_out_alloc(*h, M, N, dC2));
    CVXB_TRY(copy_matrix(*h, M, N, dC.d, dC.ld, dC2.d, dC2.ld));
    cptr = dC2.d; cld = dC2.ld;
  }
  GemmArgs g{M, N, K, dA.d, dA.ld, a_kc != 0, dB.d, dB.ld, b_kc != 0, cptr, cld, alpha, beta, tri};
  g.streamk = tri != 0;      // as the solver's big SYRKs do (taken only when the tile grid leaves a partial wave)
  CVXB_TRY(gemm_dmma(*h, g));
  Staged so; so.d = cptr; so.ld = cld;
  CVXB_TRY(copy_out(*h, M, N, so, C, ldc));
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  return CVXB_OK;
}

__global__ void fill_random_kernel(size_t count, double* a, unsigned long long seed, double lo, double hi) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    a[i] = lo + (hi - lo) * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
}

__global__ void copy_kernel(size_t count, const double2* __restrict__ a, double2* __restrict__ b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) b[i] = a[i];
}

int cvxb_debug_leaf_clocks(long long* out, int reset) { return cvxb::leaf_clocks(out, reset != 0); }

int cvxb_debug_dag_blocks(int n, int dag_block, int* starts, int cap) { return cvxb::dag_block_starts(n, dag_block, starts, cap); }

int cvxb_debug_set_schedule(cvxb_handle h, int dag_block, int dag_min_n, int dag_reserve) {
  CHECK_HANDLE(h);
  CVXB_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (dag_block < 0 && dag_min_n < 0 && dag_reserve < 0) {      // all three "keep": back to the library's defaults
    h->dag_block = 2048; h->dag_min_n = 5120; h->dag_reserve = 8; h->dag_auto = true;
    return CVXB_OK;
  }
  h->dag_auto = false;                                          // explicit settings are taken as given
  if (dag_block >= 0) h->dag_block = dag_block / NB * NB;       // 0 switches the tile-DAG schedule off
  if (dag_min_n >= 0) h->dag_min_n = dag_min_n;
  if (dag_reserve >= 0) h->dag_reserve = dag_reserve < h->sm_count ? dag_reserve : h->sm_count - 1;
  return CVXB_OK;
}

int cvxb_bench_kernel(cvxb_handle h, int which, int n, int k, int reps, double* ms_per_launch,
                      double* flops_or_bytes_per_launch) {
  CHECK_HANDLE(h);
  if (reps < 1) reps = 1;
  Handle& H = *h;
  if (which == 0) {
    double ms, fl;
    CVXB_TRY(dmma_peak_probe(H, n > 0 ? n : 20000, &ms, &fl));
    *ms_per_launch = ms; *flops_or_bytes_per_launch = fl;
    return CVXB_OK;
  }
  int ldk = pad_ld(k), ldn = pad_ld(n);
  double *G = nullptr, *C = nullptr, *invD = nullptr;
  float t = 0;
  if (which == 1 || which == 2 || which == 4) {
    size_t gcount = which == 1 ? (size_t)ldk * n : (size_t)ldn * k;
    if (which == 4) gcount = (size_t)n * k;
    CVXB_CUDA_OK(cudaMalloc((void**)&G, gcount * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&C, (which == 4 ? gcount : (size_t)ldn * n) * sizeof(double)));
    fill_random_kernel<<<1024, 256, 0, H.stream>>>(gcount, G, 12345ull, -1.0, 1.0);
    if (which != 4) fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldn * n, C, 999ull, -1.0, 1.0);
    for (int r = -1; r < reps; ++r) {
      if (r == 0) CVXB_CUDA_OK(cudaEventRecord(H.ev0, H.stream));
      if (which == 1) {        // Hessian-assembly SYRK  H = G'G, G k x n (K contiguous)
        GemmArgs g{n, n, k, G, ldk, true, G, ldk, true, C, ldn, 1.0, 0.0, 2};
        g.streamk = true;
        CVXB_TRY(gemm_dmma(H, g));
      } else if (which == 2) { // Cholesky trailing update  C -= A A', A n x k (M contiguous), lower
        GemmArgs g{n, n, k, G, ldn, false, G, ldn, false, C, ldn, -1.0, 1.0, 1};
        g.streamk = true;
        CVXB_TRY(gemm_dmma(H, g));
      } else {
        copy_kernel<<<H.sm_count * 8, 512, 0, H.stream>>>(gcount / 2, (const double2*)G, (double2*)C);
        H.launches++;
      }
    }
    CVXB_CUDA_OK(cudaEventRecord(H.ev1, H.stream));
    CVXB_CUDA_OK(cudaEventSynchronize(H.ev1));
    CVXB_CUDA_OK(cudaEventElapsedTime(&t, H.ev0, H.ev1));
    *ms_per_launch = t / reps;
    *flops_or_bytes_per_launch = which == 4 ? 16.0 * (double)gcount : (double)k * n * ((double)n + 1.0);
  } else if (which == 3) {
    // blocked Cholesky of a diagonally dominant SPD matrix (restored from a pristine copy each rep)
    CVXB_CUDA_OK(cudaMalloc((void**)&G, (size_t)ldn * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&C, (size_t)ldn * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&invD, (size_t)((n + NB - 1) / NB) * NB * NB * sizeof(double)));
    fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldn * n, G, 777ull, -1.0, 1.0);
    CVXB_TRY(add_diag(H, n, (double)n + 1.0, G, ldn));
    double total = 0;
    for (int r = -1; r < reps; ++r) {
      CVXB_TRY(copy_matrix(H, n, n, G, ldn, C, ldn));
      CVXB_CUDA_OK(cudaEventRecord(H.ev0, H.stream));
      CVXB_TRY(potrf_lower(H, n, C, ldn, invD, F_CHOL_H, S_MINDIAG_H));
      CVXB_CUDA_OK(cudaEventRecord(H.ev1, H.stream));
      CVXB_CUDA_OK(cudaEventSynchronize(H.ev1));
      CVXB_CUDA_OK(cudaEventElapsedTime(&t, H.ev0, H.ev1));
      if (r >= 0) total += t;
    }
    *ms_per_launch = total / reps;
    *flops_or_bytes_per_launch = (double)n * n * n / 3.0;
  } else if (which == 9) {
    // blocked Cholesky of an n x n matrix with the forward substitution of k right-hand sides riding along
    // (the factor_h_with_trsm range of the KKT step: C4 n = 8192, k = 2049)
    double* Bm = nullptr;
    CVXB_CUDA_OK(cudaMalloc((void**)&G, (size_t)ldn * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&C, (size_t)ldn * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&Bm, (size_t)ldn * (k > 0 ? k : 1) * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&invD, (size_t)((n + NB - 1) / NB) * NB * NB * sizeof(double)));
    fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldn * n, G, 777ull, -1.0, 1.0);
    CVXB_TRY(add_diag(H, n, (double)n + 1.0, G, ldn));
    double total = 0;
    for (int r = -1; r < reps; ++r) {
      CVXB_TRY(copy_matrix(H, n, n, G, ldn, C, ldn));
      fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldn * (k > 0 ? k : 1), Bm, 31ull, -1.0, 1.0);
      CVXB_CUDA_OK(cudaEventRecord(H.ev0, H.stream));
      CVXB_TRY(potrf_lower_rhs(H, n, C, ldn, invD, F_CHOL_H, S_MINDIAG_H, Bm, ldn, k));
      CVXB_CUDA_OK(cudaEventRecord(H.ev1, H.stream));
      CVXB_CUDA_OK(cudaEventSynchronize(H.ev1));
      CVXB_CUDA_OK(cudaEventElapsedTime(&t, H.ev0, H.ev1));
      if (r >= 0) total += t;
    }
    cudaFree(Bm);
    *ms_per_launch = total / reps;
    *flops_or_bytes_per_launch = (double)n * n * n / 3.0 + (double)n * n * k;
  } else if (which == 5 || which == 6) {
    // 5: forward + backward single-RHS solves with a factor of size n; 6: Ruiz equilibration (20 enqueued sweeps)
    CVXB_CUDA_OK(cudaMalloc((void**)&G, (size_t)ldn * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&C, (size_t)ldn * 4 * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&invD, (size_t)((n + NB - 1) / NB) * NB * NB * sizeof(double)));
    fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldn * n, G, 777ull, -1.0, 1.0);
    fill_random_kernel<<<64, 256, 0, H.stream>>>((size_t)ldn * 4, C, 5ull, -1.0, 1.0);
    double total = 0;
    if (which == 5) {
      CVXB_TRY(add_diag(H, n, (double)n + 1.0, G, ldn));
      CVXB_TRY(potrf_lower(H, n, G, ldn, invD, F_CHOL_H, S_MINDIAG_H));
    }
    for (int r = -1; r < reps; ++r) {
      CVXB_CUDA_OK(cudaEventRecord(H.ev0, H.stream));
      if (which == 5) {
        CVXB_TRY(trsm_lower(H, n, 1, G, ldn, invD, C, ldn, false));
        CVXB_TRY(trsm_lower(H, n, 1, G, ldn, invD, C, ldn, true));
      } else {
        CVXB_TRY(ruiz_equilibrate(H, n, G, ldn, C, C + ldn, 20, 1e-6, invD, (size_t)((n + NB - 1) / NB) * NB * NB));
      }
      CVXB_CUDA_OK(cudaEventRecord(H.ev1, H.stream));
      CVXB_CUDA_OK(cudaEventSynchronize(H.ev1));
      CVXB_CUDA_OK(cudaEventElapsedTime(&t, H.ev0, H.ev1));
      if (r >= 0) total += t;
    }
    *ms_per_launch = total / reps;
    *flops_or_bytes_per_launch = which == 5 ? 8.0 * n * n : 8.0 * n * n * 20;
  } else if (which == 7 || which == 8) {
    // 7: gemv_n  y = G x (G k x n: slacks, line-search direction);  8: gemv_t  y = G' w (gradient)  -- HBM-bound, 8kn bytes
    CVXB_CUDA_OK(cudaMalloc((void**)&G, (size_t)ldk * n * sizeof(double)));
    CVXB_CUDA_OK(cudaMalloc((void**)&C, (size_t)(ldk + ldn) * 2 * sizeof(double)));
    fill_random_kernel<<<1024, 256, 0, H.stream>>>((size_t)ldk * n, G, 4242ull, -1.0, 1.0);
    fill_random_kernel<<<64, 256, 0, H.stream>>>((size_t)(ldk + ldn) * 2, C, 7ull, -1.0, 1.0);
    double* xv = C;                       // length max(k, n)
    double* yv = C + (ldk + ldn);
    for (int r = -1; r < reps; ++r) {
      if (r == 0) CVXB_CUDA_OK(cudaEventRecord(H.ev0, H.stream));
      if (which == 7) CVXB_TRY(gemv_n(H, k, n, 1.0, G, ldk, xv, 0.0, yv));
      else CVXB_TRY(gemv_t(H, k, n, 1.0, G, ldk, xv, 0.0, yv));
    }
    CVXB_CUDA_OK(cudaEventRecord(H.ev1, H.stream));
    CVXB_CUDA_OK(cudaEventSynchronize(H.ev1));
    CVXB_CUDA_OK(cudaEventElapsedTime(&t, H.ev0, H.ev1));
    *ms_per_launch = t / reps;
    *flops_or_bytes_per_launch = 8.0 * (double)k * n;
  } else {
    cvxb::set_last_error("cvxb_bench_kernel: unknown kernel %d", which);
    return CVXB_EINVAL;
  }
  cudaFree(G); cudaFree(C); cudaFree(invD);
  return CVXB_OK;
}

}  // extern "C"
