package cvx

/** JNI binding of libcvxb (include/cvxb.h) through jni/cvxb_jni.c.  JDK 1.8 / Scala 2.11 compatible
  * (no Panama).  One handle = one CUDA device + stream; not thread-safe, use one per thread.
  *
  * UNVERIFIED: written against the reference's sources, but never compiled -- the build image has no
  * JVM.  The numerics behind every native method are tested through the same C ABI from Python.
  */
object CvxbNative {

  System.loadLibrary("cvxb_jni")

  @native def create(device: Int): Long
  @native def destroy(handle: Long): Unit

  @native def kktSolve(handle: Long, n: Int, p: Int, H: Array[Double], hOff: Int, ldh: Int,
                       A: Array[Double], aOff: Int, lda: Int, q: Array[Double], b: Array[Double], tol: Double,
                       x: Array[Double], w: Array[Double], info: Array[Int]): Unit

  @native def choleskySolve(handle: Long, n: Int, H: Array[Double], hOff: Int, ldh: Int, b: Array[Double],
                            tol: Double, x: Array[Double]): Unit

  @native def problemCreate(handle: Long, n: Int, m: Int, p: Int, kind: Int, objA: Array[Double], objR: Double,
                            objP: Array[Double], G: Array[Double], gR: Array[Double], ub: Array[Double],
                            A: Array[Double], b: Array[Double], xFeasible: Array[Double],
                            xDefined: Array[Double]): Long
  @native def problemDestroy(problem: Long): Unit

  /** solver: 0 barrier, 1 primal-dual; params = (maxIter, alpha, beta, tolSolver, tolEqSolve, tolFeas, delta). */
  @native def solve(handle: Long, problem: Long, solver: Int, params: Array[Double], x: Array[Double],
                    lambda: Array[Double], nu: Array[Double], stats: Array[Double]): Unit

  /** MatrixUtils.solveUnderdetermined / SolutionSpace: z0 (n), F (n x (n-p), column-major). */
  @native def solveUnderdetermined(handle: Long, p: Int, n: Int, A: Array[Double], aOff: Int, lda: Int,
                                   b: Array[Double], z0: Array[Double], F: Array[Double]): Unit

  /** KKTData.reduced -> KKTSystem.solve -> KKTData.paddVector; returns the number of eliminated indices. */
  @native def kktSolveReduced(handle: Long, n: Int, p: Int, H: Array[Double], ldh: Int, A: Array[Double], lda: Int,
                              g: Array[Double], r: Array[Double], tol: Double, x: Array[Double], w: Array[Double],
                              nullIdx: Array[Int]): Int

  /** g_i(x) of an uploaded problem's constraints; returns 1 when all are strictly satisfied. */
  @native def constraintValues(handle: Long, problem: Long, x: Array[Double], g: Array[Double]): Int

  lazy val defaultHandle: Long = create(0)

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
