package cvx

import breeze.linalg.{DenseMatrix, DenseVector}

/** Drop-in Solver (Solver.scala:29-33) for the closed-form problem families, seam A: the problem is
  * uploaded once and the whole solve -- phase I included -- runs device resident in libcvxb.
  *
  * Recognised families (everything else must keep using the CPU solvers, or seam B per step):
  *   objective   LinearObjectiveFunction(r, a) | QuadraticObjectiveFunction(r, a, P) | Dist_KL objective
  *   constraints every Constraint of the set is a LinearConstraint(r, a, ub)
  *   equalities  EqualityConstraint(A, b) or none
  *
  * Wiring: in OptimizationProblem.apply (OptimizationProblem.scala:147-155) choose
  *   GpuSolver.forProblem(objF, ineqs, eqs, solverType, pars, logger).getOrElse(<the existing solver>)
  * UNVERIFIED (no JVM in the build image).
  */
class GpuSolver(val dim: Int, kind: Int, objA: Array[Double], objR: Double, objP: Array[Double],
                G: DenseMatrix[Double], gR: Array[Double], ub: Array[Double], eqs: Option[EqualityConstraint],
                xFeasible: Option[DenseVector[Double]], xDefined: DenseVector[Double], solverType: String,
                pars: SolverParams, logger: Logger) extends Solver {

  private val m = G.rows
  private val p = eqs.map(_.A.rows).getOrElse(0)
  private val handle = CvxbNative.defaultHandle
  private val problem = CvxbNative.problemCreate(handle, dim, m, p, kind, objA, objR, objP, G.copy.data, gR, ub,
    eqs.map(_.A.copy.data).orNull, eqs.map(_.b.toArray).orNull, xFeasible.map(_.toArray).orNull, xDefined.toArray)

  def startingPoint: DenseVector[Double] = xFeasible.getOrElse(xDefined)

  def solve(debugLevel: Int = 0): Solution = {
    val x = new Array[Double](dim); val lam = new Array[Double](m); val nu = new Array[Double](math.max(p, 1))
    val stats = new Array[Double](12)
    val prm = Array(pars.maxIter.toDouble, pars.alpha, pars.beta, pars.tolSolver, pars.tolEqSolve, pars.tolFeas, pars.delta)
    try CvxbNative.solve(handle, problem, if (solverType == "BR") 0 else 1, prm, x, lam, nu, stats)
    catch { case e: CvxbInfeasibleException => throw new IllegalStateException(e.getMessage) /* InfeasibleProblemException(report, tol) */ }
    val has = stats(7).toInt
    def opt(bit: Int, v: Double): Option[Double] = if ((has & bit) != 0) Some(v) else None
    Solution(DenseVector(x),
      if ((has & 32) != 0) Some(DenseVector(lam)) else None, if ((has & 64) != 0) Some(DenseVector(nu.take(p))) else None,
      opt(1, stats(0)), opt(2, stats(1)), opt(4, stats(2)), opt(8, stats(3)), opt(16, stats(4)),
      stats(5).toInt, stats(6) != 0.0)
  }

  /** the terminationCriterion of the reference's solveSpecial is fixed by the solver type on the device */
  def solveSpecial(terminationCriterion: (OptimizationState) => Boolean, debugLevel: Int = 0): Solution = solve(debugLevel)

  override def finalize(): Unit = CvxbNative.problemDestroy(problem)
}

object GpuSolver {

  /** Some(solver) when every piece of the problem belongs to a closed-form family, else None. */
  def forProblem(objF: ObjectiveFunction, cnts: ConstraintSet, eqs: Option[EqualityConstraint], solverType: String,
                 pars: SolverParams, logger: Logger): Option[GpuSolver] = {
    val n = cnts.dim
    val lin = cnts.constraints.collect { case c: LinearConstraint => c }
    if (lin.length != cnts.constraints.length) return None
    val G = DenseMatrix.zeros[Double](lin.length, n)
    lin.zipWithIndex.foreach { case (c, i) => G(i, ::) := c.a.t }
    val gR = lin.map(_.r).toArray
    val ub = lin.map(_.ub).toArray
    val (kind, a, r, pm) = objF match {
      case f: LinearObjectiveFunction => (0, f.a.toArray, f.r, null)
      case f: QuadraticObjectiveFunction => (1, f.a.toArray, f.r, f.P.copy.data)
      case f if f.getClass.getName.contains("Dist_KL") => (2, null, 0.0, null)
      case _ => return None
    }
    val feasible = cnts match { case c: FeasiblePoint => Some(c.feasiblePoint); case _ => None }
    Some(new GpuSolver(n, kind, a, r, pm, G, gR, ub, eqs, feasible, cnts.pointWhereDefined, solverType, pars, logger))
  }
}
