package cvx

/** JNI binding of libcvxb (include/cvxb.h) through jni/cvxb_jni.c.  JDK 1.8 / Scala 2.11 compatible
  * (no Panama).  One handle = one CUDA device + stream; not thread-safe, use one per thread.
  *
  * Status: the build image has no JVM, so this file is not compiled there.  What IS checked in the repository's
  * tests: the C shim compiles warning-free against the JNI subset it uses, exports exactly the @native methods
  * declared below with the same argument counts (tests/test_boundary_cpu.py), and every one of them is executed
  * on the GPU through a fake JNIEnv (tests/test_boundary_gpu.py), exception classes and constructors included.
  */
object CvxbNative {

  System.loadLibrary("cvxb_jni")

  @native def create(device: Int): Long
  @native def destroy(handle: Long): Unit

  // ---- seam B: per-step linear algebra on Breeze storage (data, offset, majorStride) ----------------------------
  /** KKTSystem.solve (KKTSystem.scala:43-66); info = (path, regularized, ruizSweeps, cholInfo). */
  @native def kktSolve(handle: Long, n: Int, p: Int, H: Array[Double], hOff: Int, ldh: Int,
                       A: Array[Double], aOff: Int, lda: Int, q: Array[Double], b: Array[Double], tol: Double,
                       x: Array[Double], w: Array[Double], info: Array[Int]): Unit

  /** KKTSystem.solveWithCholFactor (KKTSystem.scala:99-167). */
  @native def kktSolveWithCholFactor(handle: Long, n: Int, p: Int, L: Array[Double], lOff: Int, ldl: Int,
                                     A: Array[Double], aOff: Int, lda: Int, q: Array[Double], b: Array[Double],
                                     tol: Double, x: Array[Double], w: Array[Double]): Unit

  /** MatrixUtils.choleskySolve (MatrixUtils.scala:468-516). */
  @native def choleskySolve(handle: Long, n: Int, H: Array[Double], hOff: Int, ldh: Int, b: Array[Double],
                            tol: Double, x: Array[Double]): Unit

  /** SymmetricLinearSystem.solve (SymmetricLinearSystem.scala:15-56). */
  @native def symmetricSolve(handle: Long, n: Int, H: Array[Double], hOff: Int, ldh: Int, r: Array[Double],
                             tol: Double, x: Array[Double]): Unit

  /** MatrixUtils.solveUnderdetermined / SolutionSpace: z0 (n), F (n x (n-p), column-major). */
  @native def solveUnderdetermined(handle: Long, p: Int, n: Int, A: Array[Double], aOff: Int, lda: Int,
                                   b: Array[Double], z0: Array[Double], F: Array[Double]): Unit

  /** KKTData.reduced -> KKTSystem.solve -> KKTData.paddVector; returns the number of eliminated indices. */
  @native def kktSolveReduced(handle: Long, n: Int, p: Int, H: Array[Double], ldh: Int, A: Array[Double], lda: Int,
                              g: Array[Double], r: Array[Double], tol: Double, x: Array[Double], w: Array[Double],
                              nullIdx: Array[Int]): Int

  // ---- seam A: device-resident problems ---------------------------------------------------------------------------
  /** kind: 0 linear, 1 quadratic, 2 KL, 4 p-norm (objPow).  Quadratic constraints: mq packed n x n matrices qP,
    * qA (n x mq column-major), qR, qUb.  Null arrays = absent. */
  @native def problemCreate(handle: Long, n: Int, m: Int, p: Int, kind: Int, objA: Array[Double], objR: Double,
                            objP: Array[Double], objPow: Double, G: Array[Double], gR: Array[Double], ub: Array[Double],
                            A: Array[Double], b: Array[Double], xFeasible: Array[Double], xDefined: Array[Double],
                            mq: Int, qP: Array[Double], qA: Array[Double], qR: Array[Double], qUb: Array[Double]): Long
  @native def problemDestroy(problem: Long): Unit

  /** solver: 0 barrier, 1 primal-dual; params = (maxIter, alpha, beta, tolSolver, tolEqSolve, tolFeas, delta[, bugCompat]);
    * stats(16): see jni/cvxb_jni.c stats_out. */
  @native def solve(handle: Long, problem: Long, solver: Int, params: Array[Double], x: Array[Double],
                    lambda: Array[Double], nu: Array[Double], stats: Array[Double]): Unit

  /** ConstraintSet.withFeasiblePoint: xFeasible (n), xs = (x, s) (n + 1). */
  @native def phase1(handle: Long, problem: Long, params: Array[Double], xFeasible: Array[Double], xs: Array[Double],
                     stats: Array[Double]): Unit

  @native def barrierNewtonDirection(handle: Long, problem: Long, params: Array[Double], x: Array[Double], t: Double,
                                     H: Array[Double], g: Array[Double], dx: Array[Double], nu: Array[Double],
                                     info: Array[Int]): Unit

  @native def pdNewtonDirection(handle: Long, problem: Long, params: Array[Double], x: Array[Double],
                                lambda: Array[Double], nu: Array[Double], t: Double, H: Array[Double],
                                dx: Array[Double], dlambda: Array[Double], dnu: Array[Double], info: Array[Int]): Unit

  /** g_i(x) of an uploaded problem's constraints; returns 1 when all are strictly satisfied. */
  @native def constraintValues(handle: Long, problem: Long, x: Array[Double], g: Array[Double]): Int

  // ---- equality elimination x = z0 + F u ----------------------------------------------------------------------------
  @native def solutionSpaceCreate(handle: Long, p: Int, n: Int, A: Array[Double], aOff: Int, lda: Int, b: Array[Double]): Long
  @native def solutionSpaceFromBasis(handle: Long, n: Int, k: Int, z0: Array[Double], F: Array[Double], fOff: Int, ldf: Int): Long
  @native def solutionSpaceDestroy(space: Long): Unit
  @native def solutionSpaceMap(handle: Long, space: Long, u: Array[Double], x: Array[Double]): Unit
  @native def problemReduce(handle: Long, problem: Long, space: Long, params: Array[Double]): Long

  // ---- batched small problems (n <= 64, m <= 128, p in {0,1}); returns device milliseconds ---------------------------
  @native def batchSolve(handle: Long, B: Int, n: Int, m: Int, p: Int, objective: Array[Int], pcount: Array[Int],
                         objA: Array[Double], objR: Array[Double], objP: Array[Double], G: Array[Double],
                         ub: Array[Double], A: Array[Double], b: Array[Double], x0: Array[Double], params: Array[Double],
                         x: Array[Double], status: Array[Int], newtonSteps: Array[Int], outerStages: Array[Int],
                         objectiveOut: Array[Double], dualityGap: Array[Double], equalityGap: Array[Double],
                         phase1: Array[Int], phase1NewtonSteps: Array[Int]): Double
  // phase1 (or null): 1 = x0 of that problem is only ConstraintSet.pointWhereDefined, run phase_I_Analysis first
  // (ConstraintSet.scala:326-395, 556-575; needs n <= 63, m + 2p <= 128); phase1NewtonSteps (or null) receives its steps

  lazy val defaultHandle: Long = create(0)

  def paramsArray(pars: SolverParams): Array[Double] =
    Array(pars.maxIter.toDouble, pars.alpha, pars.beta, pars.tolSolver, pars.tolEqSolve, pars.tolFeas, pars.delta)

  /** Drop-in body for SolutionSpace's `sol` member (SolutionSpace.scala:24):
    * {{{ val sol = CvxbNative.solutionSpace(A, b) }}} */
  def solutionSpace(A: breeze.linalg.DenseMatrix[Double], b: breeze.linalg.DenseVector[Double])
      : (breeze.linalg.DenseVector[Double], breeze.linalg.DenseMatrix[Double]) = {
    val Ad = if (A.isTranspose) A.copy else A
    val (p, n) = (Ad.rows, Ad.cols)
    val z0 = new Array[Double](n)
    val F = new Array[Double](n * (n - p))
    solveUnderdetermined(defaultHandle, p, n, Ad.data, Ad.offset, Ad.majorStride, b.toArray, z0, F)
    (breeze.linalg.DenseVector(z0), new breeze.linalg.DenseMatrix(n, n - p, F))
  }
}

/** Thrown by the shim for CVXB_EINFEASIBLE; GpuSolver rethrows it as InfeasibleProblemException with a report. */
class CvxbInfeasibleException(msg: String) extends Exception(msg)
