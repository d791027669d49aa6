// FP64 GEMM / SYRK on the DMMA tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, the only FP64
// tensor instruction sm_100a has: tcgen05 has no f64 kind).  This one mainloop serves every dense
// contraction of the Newton step:
//   * Hessian assembly      H  = Gs' Gs            (TN, tri=2)   BarrierSolver.scala:303-315
//   * Cholesky trailing     A22 -= L21 L21'        (NT, tri=1)   MatrixUtils.scala:452-461 (dpotrf)
//   * panel / TRSM updates  B2  -= L21 Y1          (NN)          KKTSystem.scala:116-124 (dtrtrs)
//   * Schur complement      S  = Y' Y              (TN, tri=2)   KKTSystem.scala:126-139
//
// Three CTA tile shapes share one templated mainloop (k-slab 32, 3-stage cp.async ring; 16 x 4 stages measured 3% slower):
//   128x128, 8 warps (2x4), warp tile 64x32 = 8x4 DMMA tiles: each k4-step issues 12 LDS.64 for 32 DMMA
//            (shared-memory pipe ~19% busy, the tensor pipe is the limiter) -- the big contractions;
//    64x64,  4 warps (2x2), warp tile 32x32;   32x32, 4 warps (2x2), warp tile 16x16 -- the small GEMMs
//            of the recursive Cholesky / TRSM levels, where a 128x128 grid would leave most of the 148
//            SMs idle and a single CTA would walk the whole K range alone.
// The big SYRKs on the handle's own stream run as ONE persistent CTA per SM with stream-K splitting of the tiles that
// would otherwise form a partial last wave (gemm_dmma_streamk_kernel: deterministic fix-up, bitwise symmetric).
// Operands are staged global->shared with 16-byte cp.async (zero-fill predication at the edges), each thread's
// chunk pattern computed once (SlabLoader) and the copies issued from inside the DMMA stream;
// padded shared layouts make every fragment load bank-conflict free:
//   K-contiguous operand: [R][BK+4] doubles  -> bank = 8*g + 2*t   (g = lane/4, t = lane%4)
//   M-contiguous operand: [BK][R+4] doubles  -> bank = 8*t + 2*g
#include "common.cuh"

namespace cvxb {

namespace {

#ifndef CVXB_BK
#define CVXB_BK 32
#endif
#ifndef CVXB_STAGES
#define CVXB_STAGES 3
#endif
constexpr int BK = CVXB_BK;
// cp.async ring depth per tile shape (a 4-stage ring for the 32x32 tile -- all four slabs of a rank-128 GEMM in
// flight from the prologue on -- measured no gain on the Cholesky chain)
template <int BM> struct StagesFor { static constexpr int value = CVXB_STAGES; };
constexpr int KC_LD = BK + 4;               // K-contiguous tile row stride (doubles)

template <int R> struct TileElems { static constexpr int value = (R * KC_LD > BK * (R + 4)) ? R * KC_LD : BK * (R + 4); };

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Per-thread state for staging one operand's R x BK slabs.  The chunk pattern of a thread is the same in every
// slab, so the global pointer, the row predicate and the shared-memory offset are computed ONCE; per slab a chunk
// costs one 64-bit add, one select and the cp.async itself, and the chunks are issued one or two at a time from
// inside the DMMA stream (see the main loop) instead of as a ~250-instruction block behind the barrier.
//   KC: element (r,k) at src[r*ld + k]; chunk `it` = row tid/(BK/2) + it*ROWS_PER_IT, columns kc..kc+1
//   MC: element (r,k) at src[k*ld + r]; chunk `it` = k-row tid/(R/2) + it*ROWS_PER_IT, rows rc..rc+1
// Out-of-range chunks are issued with src-size 0 (pure zero fill: the source address is never dereferenced).
template <bool KC, int R, int NTHR>
struct SlabLoader {
  static constexpr int CHUNKS_PER_LINE = KC ? BK / 2 : R / 2;
  static constexpr int ROWS_PER_IT = NTHR / CHUNKS_PER_LINE;
  static constexpr int ITERS = (R * BK / 2) / NTHR;
  static constexpr int LINE_LD = KC ? KC_LD : R + 4;
  const double* p;        // this thread's chunk 0 of the NEXT slab to load
  long long it_stride;    // elements between consecutive chunks of this thread
  unsigned soff;          // byte offset of chunk 0 inside a stage's tile
  unsigned mask;          // KC: bit it = row of chunk it is inside the matrix
  int a;                  // KC: kc (k offset inside the slab);  MC: k-row of chunk 0 inside the slab
  int rbytes;             // MC: bytes of this thread's row pair inside the matrix (0 / 8 / 16)
  int ld;

  __device__ __forceinline__ void init(const double* src, int ld_, int rows, int r0, int tid) {
    ld = ld_;
    const int line = tid / CHUNKS_PER_LINE, c2 = (tid % CHUNKS_PER_LINE) * 2;
    soff = (unsigned)((line * LINE_LD + c2) * (int)sizeof(double));
    if (KC) {
      a = c2;
      mask = 0;
#pragma unroll
      for (int it = 0; it < ITERS; ++it)
        if (r0 + line + it * ROWS_PER_IT < rows) mask |= 1u << it;
      p = src + (size_t)(r0 + line) * ld + c2;
      it_stride = (long long)ROWS_PER_IT * ld;
      rbytes = 16;
    } else {
      a = line;
      mask = 0;
      int rem = (rows - (r0 + c2)) * 8;
      rbytes = rem < 0 ? 0 : (rem > 16 ? 16 : rem);
      p = src + (size_t)line * ld + r0 + c2;
      it_stride = (long long)ROWS_PER_IT * ld;
    }
  }
  // chunk `it` of the slab whose first k index leaves `krem` = K - k0 valid k's
  __device__ __forceinline__ void issue(unsigned stage_base, int it, int krem) const {
    int bytes;
    if (KC) {
      int kb = (krem - a) * 8;
      kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
      bytes = ((mask >> it) & 1u) ? kb : 0;
    } else {
      bytes = (a + it * ROWS_PER_IT < krem) ? rbytes : 0;
    }
    unsigned dst = stage_base + soff + (unsigned)(it * ROWS_PER_IT * LINE_LD * (int)sizeof(double));
    const double* g = p + (long long)it * it_stride;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(g), "r"(bytes));
  }
  __device__ __forceinline__ void advance() { p += KC ? (long long)BK : (long long)BK * ld; }
};

// linear id -> (bm, bn) with bn <= bm, row-major over the lower triangle of tiles
__device__ __forceinline__ void tri_tile(int id, int& bm, int& bn) {
  int r = (int)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
  while ((long long)(r + 1) * (r + 2) / 2 <= id) ++r;
  while ((long long)r * (r + 1) / 2 > id) --r;
  bm = r;
  bn = id - r * (r + 1) / 2;
}

// One CTA tile: the k-slabs [kt0, kt1) of C(m0.., n0..) accumulated into acc (which the caller zeroes).
template <int BM, int BN, int WARPS_M, int WARPS_N, bool A_KC, bool B_KC>
struct CtaTile {
  static constexpr int NTHR = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int MT = WM / 8, NTL = WN / 8;
  static constexpr int A_ELEMS = TileElems<BM>::value, B_ELEMS = TileElems<BN>::value;
  static constexpr int A_MC_LD = BM + 4, B_MC_LD = BN + 4;
  static constexpr int STAGES = StagesFor<BM>::value;

  __device__ static __forceinline__ void zero(double (&acc)[MT][NTL][2]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  __device__ static __forceinline__ void mainloop(double (&acc)[MT][NTL][2], double* smem, int M, int N, int K,
                                                  const double* __restrict__ A, int lda, const double* __restrict__ B,
                                                  int ldb, int m0, int n0, int kt0, int kt1) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int KT = kt1 - kt0;
    const int Krem0 = K - kt0 * BK;               // valid k's from the first slab of this range on
    constexpr unsigned STAGE_BYTES = (unsigned)((A_ELEMS + B_ELEMS) * sizeof(double));
    constexpr unsigned B_OFF = (unsigned)(A_ELEMS * sizeof(double));
    const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
    auto stageA = [&](int s) { return smem + (size_t)s * (A_ELEMS + B_ELEMS); };
    auto stageB = [&](int s) { return smem + (size_t)s * (A_ELEMS + B_ELEMS) + A_ELEMS; };
    typedef SlabLoader<A_KC, BM, NTHR> LoadA;
    typedef SlabLoader<B_KC, BN, NTHR> LoadB;
    LoadA la;
    LoadB lb;
    la.init(A + (A_KC ? (size_t)kt0 * BK : (size_t)kt0 * BK * lda), lda, M, m0, tid);
    lb.init(B + (B_KC ? (size_t)kt0 * BK : (size_t)kt0 * BK * ldb), ldb, N, n0, tid);
    constexpr int NCHUNK = LoadA::ITERS + LoadB::ITERS;       // cp.async per thread per slab
    constexpr int KSTEPS = BK / 4;
    constexpr int PER_STEP = (NCHUNK + KSTEPS - 1) / KSTEPS;  // issued behind each k4-step's DMMAs
    auto issue_chunk = [&](unsigned sbase, int c, int krem) {
      if (c < LoadA::ITERS) la.issue(sbase, c, krem);
      else lb.issue(sbase + B_OFF, c - LoadA::ITERS, krem);
    };

    // prologue: slabs 0 .. STAGES-2 of the range
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < KT) {
        const int krem = Krem0 - s * BK;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) issue_chunk(smem_u32 + s * STAGE_BYTES, c, krem);
        la.advance();
        lb.advance();
      }
      cp_async_commit();
    }

    int cs = 0, ls = STAGES - 1;       // stage being consumed / stage being refilled
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      // slab kt+STAGES-1 goes into the stage every warp finished reading before the barrier above; its cp.asyncs
      // are spread over the k4-steps so that they issue in the shadow of the DMMA pipe
      const int krem = Krem0 - (kt + STAGES - 1) * BK;
      const bool more = kt + STAGES - 1 < KT;
      const unsigned lbase = smem_u32 + (unsigned)ls * STAGE_BYTES;
      const double* As = stageA(cs);
      const double* Bs = stageB(cs);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int kk = ks * 4;
        double a[MT], b[NTL];
#pragma unroll
        for (int i = 0; i < MT; ++i)
          a[i] = A_KC ? As[(wm0 + i * 8 + g) * KC_LD + kk + t] : As[(kk + t) * A_MC_LD + wm0 + i * 8 + g];
#pragma unroll
        for (int j = 0; j < NTL; ++j)
          b[j] = B_KC ? Bs[(wn0 + j * 8 + g) * KC_LD + kk + t] : Bs[(kk + t) * B_MC_LD + wn0 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
#pragma unroll
          for (int j = 0; j < NTL; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          if (i == MT / 2 - 1 || (MT == 1 && i == 0)) {
            if (more) {
#pragma unroll
              for (int c = ks * PER_STEP; c < (ks + 1) * PER_STEP && c < NCHUNK; ++c) issue_chunk(lbase, c, krem);
            }
          }
        }
      }
      if (more) {
        la.advance();
        lb.advance();
      }
      cp_async_commit();
      cs = cs + 1 == STAGES ? 0 : cs + 1;
      ls = ls + 1 == STAGES ? 0 : ls + 1;
    }
    cp_async_wait<0>();
  }

  // thread (g,t) of DMMA tile (i,j) holds C(m, n), C(m, n+1), m = ..+g, n = ..+2t.
  // With beta != 0 the old values of C are fetched in batches of IB row-tiles (IB*NTL*2 predicated loads in flight)
  // BEFORE any of them is used: an element-by-element load -> fma -> store loop is one dependent L2 round trip per
  // element (64 per thread in the 128x128 tile: ~50 us per tile, more than the whole mainloop of a rank-128 update).
  template <int IB = (MT >= 4 ? 4 : (MT >= 2 ? 2 : 1))>
  __device__ static __forceinline__ void epilogue(const double (&acc)[MT][NTL][2], int M, int N, double* __restrict__ C,
                                                  int ldc, double alpha, double beta, int tri, bool diag, int m0, int n0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const bool use_c = beta != 0.0;
#pragma unroll
    for (int ib = 0; ib < MT; ib += IB) {
      double cv[IB][NTL][2];
#pragma unroll
      for (int ii = 0; ii < IB; ++ii) {
        const int m = m0 + wm0 + (ib + ii) * 8 + g;
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = n0 + wn0 + j * 8 + 2 * t + e;
            const bool ok = use_c && m < M && n < N && !(diag && n > m);
            cv[ii][j][e] = ok ? C[(size_t)n * ldc + m] : 0.0;
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < IB; ++ii) {
        const int m = m0 + wm0 + (ib + ii) * 8 + g;
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = n0 + wn0 + j * 8 + 2 * t + e;
            if (m >= M || n >= N || (diag && n > m)) continue;
            double v = alpha * acc[ib + ii][j][e];
            if (use_c) v += beta * cv[ii][j][e];
            C[(size_t)n * ldc + m] = v;
            if (tri == 2 && n != m) C[(size_t)m * ldc + n] = v;
          }
        }
      }
    }
  }
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, (BM >= 128 ? 1 : 2))
gemm_dmma_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                 double* __restrict__ C, int ldc, double alpha, double beta, int tri, int tiles_m, int lower_only) {
  typedef CtaTile<BM, BN, WARPS_M, WARPS_N, A_KC, B_KC> T;
  extern __shared__ __align__(16) double smem[];
  int bm, bn;
  if (tri) {
    tri_tile(blockIdx.x, bm, bn);
  } else {
    bm = blockIdx.x % tiles_m;
    bn = blockIdx.x / tiles_m;
  }
  const int m0 = bm * BM, n0 = bn * BN;
  // programmatic dependent launch (no-ops for a plain launch): everything above overlapped the previous kernel's
  // tail; its results are visible after the wait.  Our own dependents may start launching once the mainloop is done.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (lower_only && m0 + BM <= n0) return;      // tile entirely above the diagonal
  double acc[T::MT][T::NTL][2];
  T::zero(acc);
  T::mainloop(acc, smem, M, N, K, A, lda, B, ldb, m0, n0, 0, (K + BK - 1) / BK);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  T::epilogue(acc, M, N, C, ldc, alpha, beta, tri, (tri && (bm == bn)) || lower_only, m0, n0);
}

// Stream-K variant of the 128x128 triangular (SYRK) grid: ONE persistent CTA per SM.  With T tiles and G CTAs the
// plain grid runs ceil(T/G) waves and leaves the last one partly empty (C2: 136 tiles on 148 SMs; C4: 2080 =
// 14 waves + 8 tiles).  Here the tiles of the full waves are dealt out whole (CTA i takes tiles i, i+G, ...) and the
// k-slabs of the remaining T%G tiles (of G + T%G tiles when that would leave shares shorter than 8 slabs, or when there is
// no full wave) are cut into G equal contiguous shares.  A share starts with the tail
// (or a middle piece) of a tile -- written as a partial to the CTA's workspace slot and flagged -- and ends with the
// head of another tile, whose CTA is that tile's owner: it adds the partials of CTAs i+1, i+2, ... in index order
// (fixed order: results are deterministic and tri=2 stays bitwise symmetric) and runs the epilogue.  The partials
// an owner needs were produced at the START of the other CTAs' shares, so nobody waits long; co-residency is
// guaranteed by the cooperative launch.
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 1)
gemm_dmma_streamk_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B,
                         int ldb, double* __restrict__ C, int ldc, double alpha, double beta, int tri, int tiles_total,
                         int tiles_whole, double* __restrict__ ws, int* flags, int* abort_flag,
                         const unsigned* __restrict__ order) {
  typedef CtaTile<128, 128, 2, 4, A_KC, B_KC> T;
  constexpr int PER_THREAD = T::MT * T::NTL * 2;       // 64 accumulators
  extern __shared__ __align__(16) double smem[];
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int KT = (K + BK - 1) / BK;
  double acc[T::MT][T::NTL][2];
  // whole tiles
  // tile id -> (bm, bn): through the L2-aware order table when there is one (supertiles of 12 x 12 tiles: the ~148 tiles in
  // flight at any time then share 24 operand panels instead of ~67), else row-major over the lower triangle
  auto tile_of = [&](int id, int& bm, int& bn) {
    if (order) {
      const unsigned v = __ldg(order + id);
      bm = (int)(v >> 16);
      bn = (int)(v & 0xffffu);
    } else {
      tri_tile(id, bm, bn);
    }
  };
  for (int id = cta; id < tiles_whole; id += G) {
    int bm, bn;
    tile_of(id, bm, bn);
    T::zero(acc);
    __syncthreads();          // every warp is done with the previous tile's stages
    T::mainloop(acc, smem, M, N, K, A, lda, B, ldb, bm * 128, bn * 128, 0, KT);
    T::template epilogue<2>(acc, M, N, C, ldc, alpha, beta, tri, bm == bn, bm * 128, bn * 128);
  }
  // this CTA's share of the split tiles' slabs
  const long long W = (long long)(tiles_total - tiles_whole) * KT;
  long long pos = W * cta / G;
  const long long end = W * (cta + 1) / G;
  while (pos < end) {
    const int tl = (int)(pos / KT);
    const int kb = (int)(pos - (long long)tl * KT);
    const int ke = (end - pos < KT - kb) ? kb + (int)(end - pos) : KT;
    int bm, bn;
    tile_of(tiles_whole + tl, bm, bn);
    T::zero(acc);
    __syncthreads();
    T::mainloop(acc, smem, M, N, K, A, lda, B, ldb, bm * 128, bn * 128, kb, ke);
    if (kb > 0) {
      // not the owner (only ever the first piece of a share): partial -> slot of this CTA, then the flag
      double* slot = ws + (size_t)cta * (256 * PER_THREAD);
#pragma unroll
      for (int i = 0; i < T::MT; ++i)
#pragma unroll
        for (int j = 0; j < T::NTL; ++j) {
          slot[((i * T::NTL + j) * 2 + 0) * 256 + tid] = acc[i][j][0];
          slot[((i * T::NTL + j) * 2 + 1) * 256 + tid] = acc[i][j][1];
        }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(flags + cta, 1);
    } else {
      if (ke < KT) {
        // owner of a split tile: the rest comes from the CTAs whose shares start inside this tile
        const long long tile_end = (long long)(tl + 1) * KT;
        for (int c = cta + 1; c < G && W * c / G < tile_end; ++c) {
          int ok = 1;
          if (tid == 0) {
            // bounded (~0.3 s): a partial that never arrives means the grid was not co-resident after all -- never
            // expected under a cooperative launch; flag it (the host then drops the stream-K path) instead of hanging
            int spins = 0;
            while (atomicAdd(flags + c, 0) == 0) {
              __nanosleep(64);
              if (++spins > (1 << 22) || ((spins & 1023) == 0 && *(volatile int*)abort_flag)) {
                *(volatile int*)abort_flag = 1;
                ok = 0;
                break;
              }
            }
            __threadfence();
          }
          if (!__syncthreads_and(ok)) break;
          const double* slot = ws + (size_t)c * (256 * PER_THREAD);
#pragma unroll
          for (int i = 0; i < T::MT; ++i)
#pragma unroll
            for (int j = 0; j < T::NTL; ++j) {
              acc[i][j][0] += __ldcg(slot + ((i * T::NTL + j) * 2 + 0) * 256 + tid);
              acc[i][j][1] += __ldcg(slot + ((i * T::NTL + j) * 2 + 1) * 256 + tid);
            }
          __syncthreads();
          if (tid == 0) atomicExch(flags + c, 0);      // consumed: clean for the next launch
        }
      }
      T::template epilogue<2>(acc, M, N, C, ldc, alpha, beta, tri, bm == bn, bm * 128, bn * 128);   // (4-row batches spill here)
    }
    pos += ke - kb;
  }
}

// Persistent tile loop on a RESTRICTED grid: G = (SMs - reserve) CTAs of one per SM (216 KB of shared memory each) walk
// the 128x128 tiles id0 + cta, id0 + cta + G, ... of one GEMM.  The tile-DAG schedule of the big factorisations
// (factor.cu: potrf_dag) runs its bulk updates through this kernel on a low-priority stream: the grid never takes more
// than G SMs, so the kernels of the concurrent critical chain (leaf -> panel solve -> look-ahead update of the next
// diagonal block) always find a free SM instead of waiting a whole tile time (0.13-0.27 ms at K = 1024..2048) behind a
// full wave of bulk tiles.  No inter-CTA dependency (no cooperative launch needed); `id0` lets a triangular update skip
// the leading tile rows the critical path has already updated (ids are row-major over the lower triangle).
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 1)
gemm_dmma_persist_kernel(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                         double* __restrict__ C, int ldc, double alpha, double beta, int tri, int tiles_m, int id0, int id1,
                         int lower_only) {
  typedef CtaTile<128, 128, 2, 4, A_KC, B_KC> T;
  extern __shared__ __align__(16) double smem[];
  const int KT = (K + BK - 1) / BK;
  double acc[T::MT][T::NTL][2];
  for (int id = id0 + (int)blockIdx.x; id < id1; id += (int)gridDim.x) {
    int bm, bn;
    if (tri) {
      tri_tile(id, bm, bn);
    } else {
      bm = id % tiles_m;
      bn = id / tiles_m;
    }
    const int m0 = bm * 128, n0 = bn * 128;
    if (lower_only && m0 + 128 <= n0) continue;
    T::zero(acc);
    __syncthreads();          // every warp is done with the previous tile's stages
    T::mainloop(acc, smem, M, N, K, A, lda, B, ldb, m0, n0, 0, KT);
    T::template epilogue<2>(acc, M, N, C, ldc, alpha, beta, tri, (tri && bm == bn) || lower_only, m0, n0);
  }
}

template <int BM, int WARPS_M, int WARPS_N>
constexpr int smem_bytes() { return StagesFor<BM>::value * 2 * TileElems<BM>::value * (int)sizeof(double); }

thread_local cudaStream_t g_gemm_stream = nullptr;     // set by gemm_dmma_on for the duration of one call
thread_local bool g_gemm_pdl = false;                  // set by gemm_dmma_pdl: launch as a programmatic dependent

template <int BM, int WARPS_M, int WARPS_N, bool A_KC, bool B_KC>
int launch(Handle& h, const GemmArgs& g) {
  cudaStream_t st = g_gemm_stream ? g_gemm_stream : h.stream;
  int tm = (g.M + BM - 1) / BM, tn = (g.N + BM - 1) / BM;
  long long grid = g.tri ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn;
  if (grid <= 0) return CVXB_OK;
  if (g_gemm_pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(WARPS_M * WARPS_N * 32);
    cfg.dynamicSmemBytes = smem_bytes<BM, WARPS_M, WARPS_N>();
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CVXB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, A_KC, B_KC>, g.M, g.N, g.K, g.A, g.lda,
                                    g.B, g.ldb, g.C, g.ldc, g.alpha, g.beta, g.tri, tm, g.lower_only ? 1 : 0));
    h.launches++;
    return CVXB_OK;
  }
  gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, A_KC, B_KC>
      <<<(unsigned)grid, WARPS_M * WARPS_N * 32, smem_bytes<BM, WARPS_M, WARPS_N>(), st>>>(
          g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, g.C, g.ldc, g.alpha, g.beta, g.tri, tm, g.lower_only ? 1 : 0);
  h.launches++;
  CVXB_CUDA_OK(cudaGetLastError());
  return CVXB_OK;
}

template <int BM, int WARPS_M, int WARPS_N>
int launch_layout(Handle& h, const GemmArgs& g) {
  if (g.a_kc && g.b_kc) return launch<BM, WARPS_M, WARPS_N, true, true>(h, g);
  if (!g.a_kc && g.b_kc) return launch<BM, WARPS_M, WARPS_N, false, true>(h, g);
  if (!g.a_kc && !g.b_kc) return launch<BM, WARPS_M, WARPS_N, false, false>(h, g);
  return launch<BM, WARPS_M, WARPS_N, true, false>(h, g);
}

template <int BM, int WARPS_M, int WARPS_N>
int set_attr() {
  const int b = smem_bytes<BM, WARPS_M, WARPS_N>();
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel<BM, BM, WARPS_M, WARPS_N, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
  return CVXB_OK;
}

constexpr int NT = 256;

// Register-only DMMA issue-rate probe: every warp of every SM issues independent DMMA chains.
// Used to calibrate the FP64 tensor peak the roofline fractions are quoted against.
__global__ void __launch_bounds__(NT, 1) dmma_peak_kernel(int iters, double* sink) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

}  // namespace

__global__ void prof_stamp_kernel(int which, unsigned long long* slots) {
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (which == 0) slots[0] = now;
  else { slots[1] += now - slots[0]; slots[2] += 1ull; }
}

int dmma_peak_probe(Handle& h, int iters, double* ms, double* flops) {
  cudaEvent_t e0, e1;
  CVXB_CUDA_OK(cudaEventCreate(&e0));
  CVXB_CUDA_OK(cudaEventCreate(&e1));
  int grid = h.sm_count * 2;
  dmma_peak_kernel<<<grid, NT, 0, h.stream>>>(iters / 10 + 1, h.d_scal + S_TMP3);
  CVXB_CUDA_OK(cudaEventRecord(e0, h.stream));
  dmma_peak_kernel<<<grid, NT, 0, h.stream>>>(iters, h.d_scal + S_TMP3);
  CVXB_CUDA_OK(cudaEventRecord(e1, h.stream));
  h.launches += 2;
  CVXB_CUDA_OK(cudaEventSynchronize(e1));
  float t = 0;
  CVXB_CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
  *ms = t;
  *flops = (double)grid * (NT / 32) * (double)iters * 16.0 * 512.0;   // m8n8k4 = 256 FMA = 512 flop
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return CVXB_OK;
}

// L2-aware tile orders for the triangular stream-K grids, one per tile count per dimension tm = 13..256 (n <= 32768):
// supertiles of 12 x 12 tiles, row-major over the lower triangle of supertiles, row-major inside.  With the plain
// row-major order the ~148 tiles in flight span two to three tile rows, i.e. up to 64 + 3 operand panels -- at the C4
// Hessian SYRK (K = 16384: 16.8 MB per panel, far beyond what L2 can keep across waves) all of G was re-read in every one
// of the 14 waves (ncu: 13.2 GB of DRAM reads for a 1.07 GB operand).  A 12 x 12 supertile needs 24 panels.
int gemm_dmma_build_tile_orders(Handle& h) {
  if (getenv("CVXB_NO_TILE_ORDER")) return CVXB_OK;
  constexpr int S = 12, TM_MAX = 256;
  std::vector<unsigned> table;
  h.tile_order_off.assign(TM_MAX + 1, (size_t)-1);
  for (int tm = S + 1; tm <= TM_MAX; ++tm) {
    h.tile_order_off[tm] = table.size();
    const int SR = (tm + S - 1) / S;
    for (int sr = 0; sr < SR; ++sr)
      for (int sc = 0; sc <= sr; ++sc)
        for (int bm = sr * S; bm < tm && bm < (sr + 1) * S; ++bm)
          for (int bn = sc * S; bn <= bm && bn < (sc + 1) * S; ++bn) table.push_back(((unsigned)bm << 16) | (unsigned)bn);
    if (table.size() - h.tile_order_off[tm] != (size_t)tm * (tm + 1) / 2) {
      set_last_error("gemm_dmma_build_tile_orders: internal error at tm = %d", tm);
      return CVXB_EINVAL;
    }
  }
  CVXB_CUDA_OK(cudaMalloc((void**)&h.tile_order, table.size() * sizeof(unsigned)));
  CVXB_CUDA_OK(cudaMemcpy(h.tile_order, table.data(), table.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  return CVXB_OK;
}

int gemm_dmma_init() {
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_streamk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem_bytes<128, 2, 4>()));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_streamk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem_bytes<128, 2, 4>()));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_persist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128, 2, 4>()));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_persist_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128, 2, 4>()));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_persist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128, 2, 4>()));
  CVXB_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_persist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128, 2, 4>()));
  CVXB_TRY((set_attr<128, 2, 4>()));
  CVXB_TRY((set_attr<64, 2, 2>()));
  CVXB_TRY((set_attr<32, 2, 2>()));
  return CVXB_OK;
}

int gemm_dmma_timed(Handle& h, const GemmArgs& g, double flops) {
  if (h.capturing && h.capture_plain) {
    // WHILE body: event-record nodes are not allowed there; two one-thread kernels stamp %globaltimer around the SYRK
    // and accumulate on the device (cvxb_profile_read adds the sums)
    prof_stamp_kernel<<<1, 1, 0, h.stream>>>(0, h.d_prof);
    int st = gemm_dmma(h, g);
    prof_stamp_kernel<<<1, 1, 0, h.stream>>>(1, h.d_prof);
    h.launches += 2;
    h.capture_flops += flops;
    return st;
  }
  if (h.capturing) {   // inside a captured Newton step: external event nodes, read back after each replay
    CVXB_CUDA_OK(cudaEventRecordWithFlags(h.gev0, h.stream, cudaEventRecordExternal));
    int st = gemm_dmma(h, g);
    CVXB_CUDA_OK(cudaEventRecordWithFlags(h.gev1, h.stream, cudaEventRecordExternal));
    h.capture_flops += flops;
    return st;
  }
  if (!h.prof_on) return gemm_dmma(h, g);
  if (h.prof_used + 2 > h.prof_events.size()) {
    for (int i = 0; i < 64; ++i) {
      cudaEvent_t e;
      CVXB_CUDA_OK(cudaEventCreate(&e));
      h.prof_events.push_back(e);
    }
  }
  CVXB_CUDA_OK(cudaEventRecord(h.prof_events[h.prof_used], h.stream));
  int st = gemm_dmma(h, g);
  CVXB_CUDA_OK(cudaEventRecord(h.prof_events[h.prof_used + 1], h.stream));
  h.prof_used += 2;
  h.prof_flops += flops;
  return st;
}

int gemm_dmma_pdl(Handle& h, const GemmArgs& g) {
  g_gemm_pdl = true;
  int r = gemm_dmma(h, g);
  g_gemm_pdl = false;
  return r;
}

int gemm_dmma_on(Handle& h, const GemmArgs& g, cudaStream_t st) {
  g_gemm_stream = st;
  int r = gemm_dmma(h, g);
  g_gemm_stream = nullptr;
  return r;
}

int gemm_dmma(Handle& h, const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0) return CVXB_OK;
  if ((g.lda & 1) || (g.ldb & 1) || ((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15)) {
    set_last_error("gemm_dmma: operands must be 16-byte aligned with even leading dimensions");
    return CVXB_EINVAL;
  }
  if (g.tri && g.M != g.N) {
    set_last_error("gemm_dmma: triangular mode needs M == N");
    return CVXB_EINVAL;
  }
  // tile shape: the largest one that still fills the machine (in-place callers pin 128x128)
  auto ntiles = [&](int b) {
    long long tm = (g.M + b - 1) / b, tn = (g.N + b - 1) / b;
    return g.tri ? tm * (tm + 1) / 2 : tm * tn;
  };
  const long long want = (long long)h.sm_count * 3 / 4;
  int tile = g.tile;
  if (tile == 0) tile = ntiles(128) >= want ? 128 : (ntiles(64) >= want ? 64 : 32);
  // short contractions (rank-128 / rank-256 updates): the 64x64 tile runs two CTAs per SM, so one CTA's epilogue (read,
  // update and write of C: as long as the mainloop itself at K = 128) overlaps the other's mainloop
  static const int smallk = getenv("CVXB_SMALLK") ? atoi(getenv("CVXB_SMALLK")) : 256;
  if (g.tile == 0 && tile == 128 && g.K <= smallk && g.beta != 0.0) tile = 64;
  if (tile == 128 && g.streamk && g.tri && h.sk_ws && g_gemm_stream == nullptr && g.a_kc == g.b_kc) {
    // stream-K for the big SYRKs on the handle's own stream (see gemm_dmma_streamk_kernel)
    const long long T = ntiles(128);
    const int G = h.sm_count, KT = (g.K + BK - 1) / BK;
    const int r = (int)(T % G);
    if (r != 0 && KT >= 16 && T >= G / 2 && T < (1ll << 30)) {
      // tiles cut along K: only the r tiles of the partial last wave when their K-shares are long enough (>= 8 slabs per
      // CTA), else one more full wave with them.  Cutting few tiles keeps the CTAs of every full wave aligned in k, so
      // the panels they share are read from HBM once per wave; a cut wave's CTAs sit at staggered k offsets and re-read
      // each panel (ncu at C4: 5 of the 10.7 GB of DRAM reads came from a cut wave of 156 tiles).
      static const bool short_tail = getenv("CVXB_SK_LONG_TAIL") == nullptr;
      const long long tail = (short_tail && T >= G && (long long)r * KT / G >= 8) ? r : (long long)G + r;
      const int whole = T >= tail ? (int)(T - tail) : 0;
      int M = g.M, N = g.N, K = g.K, lda = g.lda, ldb = g.ldb, ldc = g.ldc, tri = g.tri, tt = (int)T, tw = whole;
      const double *A = g.A, *B = g.B;
      double* C = g.C;
      double alpha = g.alpha, beta = g.beta;
      double* ws = h.sk_ws;
      int* fl = h.sk_flags;
      int* ab = h.d_flag + F_WAVE_ABORT;
      const int tm = (g.M + 127) / 128;
      const unsigned* ord = (h.tile_order && tm < (int)h.tile_order_off.size() && h.tile_order_off[tm] != (size_t)-1)
                                ? h.tile_order + h.tile_order_off[tm] : nullptr;
      void* args[] = {&M, &N, &K, &A, &lda, &B, &ldb, &C, &ldc, &alpha, &beta, &tri, &tt, &tw, &ws, &fl, &ab, &ord};
      const void* fn = g.a_kc ? (const void*)gemm_dmma_streamk_kernel<true, true>
                              : (const void*)gemm_dmma_streamk_kernel<false, false>;
      cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(256), args, smem_bytes<128, 2, 4>(), h.stream);
      if (e == cudaSuccess) {
        h.launches++;
        return CVXB_OK;
      }
      cudaGetLastError();       // not launchable cooperatively here: the plain grid below
    }
  }
  if (h.sk_reserve > 0 && g.tile == 0 && (tile == 128 || (g.tri && g.tri_skip > 0))) {
    // bulk stream of the tile-DAG factorisation: persistent loop on (SMs - reserve) CTAs when the GEMM is long enough to
    // keep the chain's kernels waiting (at least half a wave of tiles; shorter ones finish within a chain step anyway)
    const int tm = (g.M + 127) / 128;
    const long long T = ntiles(128);
    long long id0 = 0;
    if (g.tri && g.tri_skip > 0) id0 = (long long)g.tri_skip * (g.tri_skip + 1) / 2;
    const int G = h.sm_count - h.sk_reserve;
    if (G > 0 && T > id0 && (T - id0 >= G / 2 || g.tri_skip > 0) && T < (1ll << 30)) {
      cudaStream_t st = g_gemm_stream ? g_gemm_stream : h.stream;
      const int grid = (int)(T - id0 < G ? T - id0 : G);
      const int lo = g.lower_only ? 1 : 0;
#define CVXB_PERSIST(AK, BKC)                                                                                              \
  gemm_dmma_persist_kernel<AK, BKC><<<grid, 256, smem_bytes<128, 2, 4>(), st>>>(g.M, g.N, g.K, g.A, g.lda, g.B, g.ldb, g.C, \
                                                                               g.ldc, g.alpha, g.beta, g.tri, tm, (int)id0, \
                                                                               (int)T, lo)
      if (g.a_kc && g.b_kc) CVXB_PERSIST(true, true);
      else if (!g.a_kc && g.b_kc) CVXB_PERSIST(false, true);
      else if (!g.a_kc && !g.b_kc) CVXB_PERSIST(false, false);
      else CVXB_PERSIST(true, false);
#undef CVXB_PERSIST
      h.launches++;
      CVXB_CUDA_OK(cudaGetLastError());
      return CVXB_OK;
    }
  }
  if (g.tri && g.tri_skip > 0) {
    if ((long long)g.tri_skip * 128 >= g.M) return CVXB_OK;      // nothing left below the skipped block
    set_last_error("gemm_dmma: tri_skip is only served on the bulk stream (persistent 128x128 grid)");
    return CVXB_EINVAL;
  }
  if (tile == 128) return launch_layout<128, 2, 4>(h, g);    // (a 16-warp 4x4 layout measured 7% slower)
  if (tile == 64) return launch_layout<64, 2, 2>(h, g);
  return launch_layout<32, 2, 2>(h, g);
}

}  // namespace cvxb
