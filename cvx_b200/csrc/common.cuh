// cvx_b200 -- shared declarations for the CUDA side of the Newton/KKT hot path.
// sm_100a only.  All matrices are column-major FP64 with padded leading dimensions
// (multiples of 16 doubles = 128 B) so every 16-byte cp.async / vector access is aligned.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "../../include/cvxb.h"
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost a pointer test when no tool is attached

namespace cvxb {

#define CVXB_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cvxb::set_last_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                \
                           cudaGetErrorString(_e));                                          \
      return CVXB_ECUDA;                                                                     \
    }                                                                                        \
  } while (0)

#define CVXB_TRY(expr)                                                                       \
  do {                                                                                       \
    int _s = (expr);                                                                         \
    if (_s != CVXB_OK) return _s;                                                            \
  } while (0)

void set_last_error(const char* fmt, ...);

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int pad_ld(int rows) { return round_up(rows < 1 ? 1 : rows, 16); }

constexpr int NB = 128;          // diagonal-block size of the blocked factorisations / solves

}  // namespace cvxb

// One handle = one device + one stream (include/cvxb.h "Threading").
struct cvxb_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  unsigned flags = 0;
  long long launches = 0;
  long long status_reads = 0;     // host round trips: copies of the status block followed by a stream synchronisation
  int sm_count = 148;
  // device scalar / flag blocks shared by all kernels of this handle, and their pinned mirrors
  double* d_scal = nullptr;   // cvxb::NSCAL doubles
  int* d_flag = nullptr;      // cvxb::NFLAG ints
  double* h_scal = nullptr;
  int* h_flag = nullptr;
  double* d_part = nullptr;   // partial-sum scratch for split reductions (PART_DOUBLES doubles)
  unsigned* d_ticket = nullptr;  // last-block-done counters
  cudaStream_t stream2 = nullptr;            // look-ahead stream of the blocked Cholesky
  std::vector<cudaEvent_t> la_events;        // fork / join events of the look-ahead schedule
  unsigned* tile_order = nullptr;            // L2-aware tile orders of the triangular stream-K grids (gemm_dmma.cu)
  std::vector<size_t> tile_order_off;        // offset of the table for tm tiles per dimension, (size_t)-1 = none
  double* sk_ws = nullptr;       // stream-K partial tiles: one 128x128 slot per SM
  int* sk_flags = nullptr;       // "slot c holds a partial" flags (self-cleaning)
  int* wave_ready = nullptr;     // per-block "y_k published" flags of the wavefront triangular solves
  void* kkt_cache = nullptr;     // cvxb::KktWork of the last seam-B call (re-used when (n,p) repeat)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // optional per-launch timing of the dominant kernel (Hessian-assembly SYRK), bench.py roofline
  // CUDA-graph replay of the fixed per-step launch sequence
  bool capturing = false;
  double capture_flops = 0.0;                   // algorithmic flops of the timed SYRK inside the step being captured
  cudaEvent_t gev0 = nullptr, gev1 = nullptr;   // external events around the SYRK inside a captured step
  double prof_ms_graph = 0.0;
  long long prof_launches_graph = 0;
  int use_graphs = 1;
  int use_loop = 1;                             // centering stages driven from the device (WHILE node); 0 = one graph launch per step
  bool capture_plain = false;                   // capturing into a WHILE body: no event-record nodes allowed there
  unsigned long long* d_prof = nullptr;         // [0] start stamp, [1] summed ns, [2] count: SYRK timing inside WHILE bodies (%globaltimer)
  int prof_on = 0;
  std::vector<cudaEvent_t> prof_events;   // start/stop pairs (range PROF_HESSIAN)
  size_t prof_used = 0;
  double prof_flops = 0.0;
  // further timed ranges of one Newton step (cvxb_profile_read_range; plain launches only, not inside a captured step)
  struct ProfRange {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    double work = 0.0;
  };
  ProfRange prof_range[8];
  // tile-DAG schedule of the big factorisations (factor.cu: potrf_dag): a third stream carries the bulk updates through
  // persistent GEMM grids that leave `sk_reserve` SMs to the kernels of the concurrent critical chain
  cudaStream_t stream3 = nullptr;
  double* d_part3 = nullptr;                 // scratch of the bulk lanes' triangular-solve leaves (DAG_LANES x PART3_DOUBLES)
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // side lanes: latency-bound triangular-solve recursions run beside each other
  size_t part_cap = (size_t)1 << 22;         // doubles usable at d_part (switched together with d_part)
  std::vector<cudaEvent_t> dag_events;
  int sk_reserve = 0;                        // != 0 only while bulk work is being enqueued
  bool in_dag = false;                       // potrf_dag is enqueuing (its diagonal blocks use the look-ahead / recursive schedules)
  int dag_block = 2048, dag_min_n = 5120, dag_reserve = 8;   // cvxb_debug_set_schedule
  bool dag_auto = true;                      // potrf_dag may narrow the blocks / the reserve for narrow right-hand sides
};

namespace cvxb {

typedef cvxb_handle_s Handle;

// bump allocator over ONE cudaMalloc: a problem's ~60 buffers cost one driver call instead of sixty
struct Arena {
  char* base = nullptr;
  size_t size = 0, used = 0;
  void* take(size_t bytes) {
    size_t a = (used + 255) & ~(size_t)255;
    if (!base || a + bytes > size) return nullptr;
    used = a + bytes;
    return base + a;
  }
};
constexpr int NSCAL = 128, NFLAG = 64;

// every extern "C" entry point switches to the handle's device and restores the caller's on return
struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// NVTX range for the scope (SURVEY.md section 5 "tracing"): solves, outer stages, Newton steps and the factorisation show
// up as nested ranges in Nsight Systems / Compute timelines
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// timed ranges (cvxb_profile_read_range): CUDA events on the handle's stream around a piece of the step
enum ProfId {
  PROF_HESSIAN = 0,      // Hessian-assembly SYRK G' diag(w) G (kept in prof_events for the captured-step path)
  PROF_CHOL_TRAIL = 1,   // Cholesky trailing updates A22 -= A21 A21' of the recursive levels (NT SYRK, lower)
  PROF_FACTOR = 2,       // Ruiz-scaled Cholesky of H with the forward substitution of [DA', Dq] riding along
  PROF_SCHUR = 3,        // Schur complement S = Y'Y (TN SYRK, mirrored)
  PROF_RUIZ = 4,         // ruizEquilibrate(H)
  PROF_GEMV = 5,         // the HBM-bound GEMVs over G of one step (bytes instead of flops)
  PROF_LOOKAHEAD = 6,    // chain-bound phases of the factorisation: the look-ahead schedule on diagonal blocks <= 2560..4608 columns
  PROF_TRSM_RIGHT = 7,   // A21 := A21 L11^-T at the recursive levels
  PROF_COUNT = 8
};
int prof_begin(cvxb_handle_s& h, int id);
int prof_end(cvxb_handle_s& h, int id, double work);
constexpr size_t PART_DOUBLES = (size_t)1 << 22;   // 32 MiB of split-K partials for gemv_n
constexpr size_t PART3_DOUBLES = (size_t)1 << 21;  // scratch per bulk lane (out-of-place leaf solves of up to 15872 rows)
constexpr int DAG_LANES = 4;                       // stream3 + three side lanes

// count + launch + error check
#define CVXB_LAUNCH(h, kernel, grid, block, smem, ...)                                       \
  do {                                                                                       \
    kernel<<<(grid), (block), (smem), (h).stream>>>(__VA_ARGS__);                            \
    (h).launches++;                                                                          \
    CVXB_CUDA_OK(cudaGetLastError());                                                        \
  } while (0)

// ---- device scalar slots (h.d_scal) ------------------------------------------------------------
enum Scal {
  S_RUIZ_RHO = 0, S_MINDIAG_H, S_MINDIAG_S, S_ERR1, S_ERR2, S_NORM_Q, S_NORM_B,
  S_FVAL,        // barrier function value at x
  S_F0,          // objective value f0(x)
  S_LOGSUM, S_NORMGRAD, S_EQNORM, S_Q /* d.y */, S_ND, S_STEP, S_TRUST, S_HNORM, S_T,
  S_C1, S_C2,    // quadratic-objective line coefficients
  S_TMP0, S_TMP1, S_TMP2, S_TMP3,
  S_PD_GAP, S_PD_RNORM, S_PD_SMAX, S_PD_RDUAL, S_PD_EQGAP, S_PD_T, S_OBJ,
  S_MINSLACK,
  // device-driven centering stage (solver.cu: stage loop as a CUDA-graph WHILE node): loop state the host seeds and reads back
  S_L_ND, S_L_NORMGRAD, S_L_EQGAP, S_L_TOL,
  S_COUNT
};
// ---- device flag slots (h.d_flag) --------------------------------------------------------------
enum Flag {
  F_RUIZ_DONE = 0, F_RUIZ_SWEEPS, F_CHOL_H /* 0 or 1-based failing column */, F_CHOL_S, F_INFEAS /* slack<=0 */,
  F_LS_STATUS /* 0 ok, 1 set-backtrack failed, 2 armijo failed, 3 not feasible in value */, F_LS_TRIALS,
  F_STEP_TAKEN, F_BAD /* any failure upstream: gates the x update */, F_ZERO_DIAG,
  F_PD_LS_FAIL, F_PD_NOTNEG, F_PD_LAMNEG, F_ITER0 /* first Newton step of an unconstrained stage */,
  F_WAVE_ABORT /* a wavefront solve gave up waiting (never expected) */,
  // stage-loop state (see S_L_*): Newton iterations counted as the reference counts them, steps executed, line-search
  // trials, remaining step budget, and why the loop stopped
  F_L_ITER, F_L_EXEC, F_L_TRIALS, F_L_BUDGET, F_L_LIMITED, F_L_MAXITER, F_L_MODE, F_L_REASON,
  F_COUNT
};

// ---------------------------------------------------------------- GEMM (gemm_dmma.cu)
// C[MxN] (col-major, ldc) = alpha * op(A) * op(B) + beta * C on FP64 DMMA tensor cores.
//   a_kc: A(m,k) = A[m*lda + k]   (K contiguous) else A[k*lda + m] (M contiguous)
//   b_kc: B(k,n) = B[n*ldb + k]   (K contiguous) else B[k*ldb + n] (N contiguous)
//   tri:  0 = all tiles; 1 = only tiles with bm >= bn are computed and only elements m >= n are
//         written (lower triangle); 2 = like 1 and the strict lower part is mirrored to the
//         upper triangle (exactly symmetric result).  Needs M == N.
// Requirements: lda/ldb even, A/B 16-byte aligned, tile origins even.
struct GemmArgs {
  int M, N, K;
  const double* A; int lda; bool a_kc;
  const double* B; int ldb; bool b_kc;
  double* C; int ldc;
  double alpha, beta;
  int tri;
  int tile = 0;   // 0 = choose 128 / 64 / 32 by grid size; in-place callers (C aliases A or B) must pin 128
  bool lower_only = false;   // rectangular C whose rows and columns share an origin: never write elements with m < n
  bool streamk = false;      // tri != 0, 128x128 tiles, handle's own stream: persistent stream-K grid (no partial last wave)
  int tri_skip = 0;          // tri != 0 on the bulk stream: leave out the leading tri_skip x tri_skip block of tiles (done elsewhere)
};
int gemm_dmma(Handle& h, const GemmArgs& g);
int gemm_dmma_on(Handle& h, const GemmArgs& g, cudaStream_t st);   // same, on another stream of the handle
// same on the handle's stream, launched as a programmatic dependent of the kernel enqueued just before it (its
// prologue overlaps that kernel's tail); only for launches with no other operation between the two
int gemm_dmma_pdl(Handle& h, const GemmArgs& g);
// gemm_dmma bracketed by CUDA events when the handle's profiling is on (flops = algorithmic flops)
int gemm_dmma_timed(Handle& h, const GemmArgs& g, double flops);
int gemm_dmma_init();   // sets the dynamic-smem attribute on all instantiations (once per device)
int gemm_dmma_build_tile_orders(Handle& h);   // per handle: device tables of the L2-aware tile orders
int dmma_peak_probe(Handle& h, int iters, double* ms, double* flops);   // register-only DMMA issue rate

// ---------------------------------------------------------------- BLAS-2 (blas2.cu)
// y[m] = alpha * A x + beta * y      A m x n col-major
int gemv_n(Handle& h, int m, int n, double alpha, const double* A, int lda, const double* x, double beta, double* y);
// y[n] = alpha * A' x + beta * y     A m x n col-major, x length m
int gemv_t(Handle& h, int m, int n, double alpha, const double* A, int lda, const double* x, double beta, double* y);
// Gs(i,j) = s_i * G(i,j)             m x n
int scale_rows(Handle& h, int m, int n, const double* G, int ldg, const double* s, double* Gs, int ldgs, bool sqrt_of_s);
// C = alpha * A (+ diag)              n x n (objective Hessian prefill); A may be NULL (zero)
int fill_matrix(Handle& h, int n, double alpha, const double* A, int lda, const double* diag_num, double diag_scale,
                double* C, int ldc, const double* mul_dev = nullptr,   // alpha, diag_scale *= *mul_dev when given
                double diag_exp = -1.0);   // diagonal += diag_scale * |diag_num|^diag_exp  (-1: KL's 1/x; p-2: p-norm)
// Bt(i,j) = s_i * A(j,i)   (n x p from p x n), s may be NULL
int transpose_scale(Handle& h, int p, int n, const double* A, int lda, const double* s, double* Bt, int ldbt);
// C (n x n) += alpha * I
int add_diag(Handle& h, int n, double alpha, double* C, int ldc);
int copy_matrix(Handle& h, int m, int n, const double* A, int lda, double* B, int ldb);

// ---------------------------------------------------------------- factorisations (factor.cu)
// MatrixUtils.ruizEquilibrate: leaves d in `d` (n), sweeps in flag F_RUIZ_SWEEPS.  No host sync.
// big_scratch (optional): >= 2 * ceil(n/128) * pad_ld(n) doubles -- enables the symmetric-half sweeps for n >= 4096
int ruiz_equilibrate(Handle& h, int n, const double* Hm, int ldh, double* d, double* colnorm_scratch,
                     int max_sweeps, double tol, double* big_scratch = nullptr, size_t big_doubles = 0);
// L := lower((d d') o H) + delta*I, zeros above the diagonal
int scaled_lower(Handle& h, int n, const double* Hm, int ldh, const double* d, double delta, double* L, int ldl);
// full Q = (d d') o H
int scaled_full(Handle& h, int n, const double* Hm, int ldh, const double* d, double* Q, int ldq);
// in-place blocked Cholesky of the lower triangle; inverse diagonal blocks to invD (ceil(n/NB) * NB*NB);
// failure column -> d_flag[flag_slot] (first failure wins), min diag -> d_scal[mindiag_slot]
int potrf_lower(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot);
// same, and B (n x r) := L^-1 B computed along the way (the forward substitution overlaps the factorisation)
int potrf_lower_rhs(Handle& h, int n, double* A, int lda, double* invD, int flag_slot, int mindiag_slot, double* B, int ldb,
                    int r);
// B (n x r) := L^-1 B   /  L^-T B, using the inverse diagonal blocks
int trsm_lower(Handle& h, int n, int r, const double* L, int ldl, const double* invD, double* B, int ldb, bool trans);
// inverse diagonal blocks of a given lower-triangular matrix (for cvxb_triangular_solve and
// solveWithCholFactor); zero diagonal -> F_ZERO_DIAG
int factor_init();   // kernel attributes (once per process, outside any stream capture)
int dag_block_starts(int n, int nbk, int* out, int cap);   // block starts of the tile-DAG schedule (+ n as the last entry)
int leaf_clocks(long long* out, bool reset);   // per-phase clock64 sums of the leaf kernel (debug builds)
int invert_diag_blocks(Handle& h, int n, const double* L, int ldl, double* invD);

}  // namespace cvxb
