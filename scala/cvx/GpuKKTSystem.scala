package cvx

import breeze.linalg.{DenseMatrix, DenseVector}

/** Drop-in for KKTSystem.solve (KKTSystem.scala:43-66), seam B: the same constructor asserts, the same
  * (x, w) result, LinSolveException / UnsolvableSystemException on failure; the linear algebra runs in
  * libcvxb (Ruiz equilibration, Cholesky, Schur complement, fallback H + A'A) on the GPU.
  *
  * To switch the reference over, replace the body of KKTSystem.solve by
  *   GpuKKTSystem(M, A, q, b).solve(delta, logger, tol, debugLevel)
  * Failures arrive as the reference's own exception types: LinSolveException (built by the shim with its real
  * 4-argument constructor, matrices null: they stay on the device) and UnsolvableSystemException.
  * Not compiled in this repository (no JVM in the build image); see CvxbNative for what is checked.
  */
class GpuKKTSystem(val H: DenseMatrix[Double], val A: DenseMatrix[Double],
                   val q: DenseVector[Double], val b: DenseVector[Double]) {

  val n: Int = H.cols
  val p: Int = A.rows
  assert(n == H.rows, "Matrix M not square: n=M.cols=" + n + ", M.rows=" + H.rows)
  assert(A.cols == n, "Dimension mismatch A.cols=" + A.cols + " not equal to n=M.cols=" + n)

  /** column-major copy with unit stride when the matrix is a transposed / strided view */
  private def dense(M: DenseMatrix[Double]): DenseMatrix[Double] =
    if (!M.isTranspose && M.offset == 0 && M.majorStride == M.rows) M else M.copy

  def solve(delta: Double, logger: Logger, tol: Double, debugLevel: Int): (DenseVector[Double], DenseVector[Double]) = {
    val Hd = dense(H); val Ad = dense(A)
    val x = new Array[Double](n); val w = new Array[Double](p)
    val info = new Array[Int](4)
    CvxbNative.kktSolve(CvxbNative.defaultHandle, n, p, Hd.data, Hd.offset, Hd.majorStride, Ad.data, Ad.offset,
                        Ad.majorStride, q.toArray, b.toArray, tol, x, w, info)
    if (debugLevel > 1) logger.println("GpuKKTSystem: path " + info(0) + ", regularized " + info(1) + ", Ruiz sweeps " + info(2))
    (DenseVector(x), DenseVector(w))
  }
}

object GpuKKTSystem {
  def apply(H: DenseMatrix[Double], A: DenseMatrix[Double], q: DenseVector[Double], b: DenseVector[Double]) =
    new GpuKKTSystem(H, A, q, b)
}
