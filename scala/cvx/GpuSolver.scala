package cvx

import breeze.linalg.{DenseMatrix, DenseVector, norm}

/** Drop-in Solver (Solver.scala:24-47) for the closed-form problem families, seam A: the problem is
  * uploaded once and the whole solve -- phase I included -- runs device resident in libcvxb.
  *
  * Recognised families (everything else must keep using the CPU solvers, or seam B per step):
  *   objective   LinearObjectiveFunction(r, a) | QuadraticObjectiveFunction(r, a, P) | Dist_KL objective |
  *               ObjectiveFunctions.p_norm_p (by hint: it is an anonymous class)
  *   constraints LinearConstraint(r, a, ub) and QuadraticConstraint(r, a, P, ub), linear ones first (the order of the
  *               multipliers lambda in the returned Solution)
  *   equalities  EqualityConstraint(A, b) or none
  *
  * Wiring: in OptimizationProblem.apply (OptimizationProblem.scala:147-155) choose
  *   GpuSolver.forProblem(objF, ineqs, eqs, solverType, pars, logger).getOrElse(<the existing solver>)
  * Not compiled in this repository (no JVM in the build image); see CvxbNative for what is checked.
  */
class GpuSolver private (handle: Long, problem: Long, val problemDim: Int, m: Int, p: Int, solverType: String,
                         val pars: SolverParams, logger: Logger, start: DenseVector[Double],
                         report: (DenseVector[Double]) => FeasibilityReport,
                         space: Long = 0L) extends Solver {

  def startingPoint: DenseVector[Double] = start

  def solve(debugLevel: Int = 0): Solution = {
    val x = new Array[Double](problemDim); val lam = new Array[Double](math.max(m, 1)); val nu = new Array[Double](math.max(p, 1))
    val stats = new Array[Double](16)
    try CvxbNative.solve(handle, problem, if (solverType == "BR") 0 else 1, CvxbNative.paramsArray(pars), x, lam, nu, stats)
    catch {
      // CVXB_EINFEASIBLE: phase I ended with s >= tol; the reference throws InfeasibleProblemException(report, tol)
      // (InfeasibleProblemException.scala:6, ConstraintSet.scala:566-571) with the report of the point found
      case e: CvxbInfeasibleException => throw new InfeasibleProblemException(report(DenseVector(x)), pars.tolSolver)
    }
    val has = stats(7).toInt
    def opt(bit: Int, v: Double): Option[Double] = if ((has & bit) != 0) Some(v) else None
    Solution(DenseVector(x),
      if ((has & 32) != 0) Some(DenseVector(lam.take(m))) else None, if ((has & 64) != 0) Some(DenseVector(nu.take(p))) else None,
      opt(1, stats(0)), opt(2, stats(1)), opt(4, stats(2)), opt(8, stats(3)), opt(16, stats(4)),
      stats(5).toInt, stats(6) != 0.0)
  }

  /** The device solvers implement exactly the two termination criteria the reference ever passes
    * (CvxUtils.standardTerminationCriterion, CvxUtils.scala:61-70, and phase_I_TerminationCriterion, :78-87, which
    * cvxb_phase1 uses internally).  A caller-supplied criterion is honoured by checking it on the state of the
    * returned solution; if it is not met there the GPU path cannot serve the call and says so. */
  def solveSpecial(terminationCriterion: (OptimizationState) => Boolean, debugLevel: Int = 0): Solution = {
    val sol = solve(debugLevel)
    val state = OptimizationState(sol.normGrad, sol.newtonDecrement, sol.dualityGap, sol.equalityGap, None,
                                  sol.normDualResidual)
    if (!terminationCriterion(state) && !sol.maxedOut)
      throw new UnsupportedOperationException(
        "GpuSolver.solveSpecial: the device loop stops on the standard criterion; this criterion is not met at its solution")
    sol
  }

  /** Solver.affineTransformed (Solver.scala:33-46; BarrierSolver.scala:209-233): the same problem in the variable u of
    * x = z0 + F u, solved on the device (cvxb_problem_reduce); the Solution is reported in u as in the reference. */
  def affineTransformed(z0: DenseVector[Double], F: DenseMatrix[Double], u0: DenseVector[Double]): Solver = {
    assert(norm(startingPoint - (z0 + F * u0)) < pars.tolEqSolve, "\nu0 does not map to x0 under the variable transform.\n")
    val Fd = if (!F.isTranspose) F else F.copy
    val sp = CvxbNative.solutionSpaceFromBasis(handle, F.rows, F.cols, z0.toArray, Fd.data, Fd.offset, Fd.majorStride)
    val red = CvxbNative.problemReduce(handle, problem, sp, CvxbNative.paramsArray(pars))
    new GpuSolver(handle, red, F.cols, m, 0, solverType, pars, logger, u0, u => report(z0 + F * u), sp)
  }

  override def finalize(): Unit = {
    CvxbNative.problemDestroy(problem)
    if (space != 0L) CvxbNative.solutionSpaceDestroy(space)
  }
}

object GpuSolver {

  private def packed(M: DenseMatrix[Double]): Array[Double] =
    if (!M.isTranspose && M.offset == 0 && M.majorStride == M.rows) M.data else M.copy.data

  /** Some(solver) when every piece of the problem belongs to a closed-form family, else None.
    * pNorm: Some(p) when objF is ObjectiveFunctions.p_norm_p(dim, p) (an anonymous class, not recognisable by type). */
  def forProblem(objF: ObjectiveFunction, cnts: ConstraintSet, eqs: Option[EqualityConstraint], solverType: String,
                 pars: SolverParams, logger: Logger, pNorm: Option[Double] = None): Option[GpuSolver] = {
    val n = cnts.dim
    val lin = cnts.constraints.collect { case c: LinearConstraint => c }
    val quad = cnts.constraints.collect { case c: QuadraticConstraint => c }
    if (lin.length + quad.length != cnts.constraints.length) return None
    val G = DenseMatrix.zeros[Double](lin.length, n)
    lin.zipWithIndex.foreach { case (c, i) => G(i, ::) := c.a.t }
    val gR = lin.map(_.r).toArray
    val ub = lin.map(_.ub).toArray
    val mq = quad.length
    val qP = if (mq == 0) null else quad.flatMap(c => packed(c.P)).toArray
    val qA = if (mq == 0) null else quad.flatMap(_.a.toArray).toArray
    val qR = if (mq == 0) null else quad.map(_.r).toArray
    val qUb = if (mq == 0) null else quad.map(_.ub).toArray
    val (kind, a, r, pm, pw) = objF match {
      case f: LinearObjectiveFunction => (0, f.a.toArray, f.r, null, 2.0)
      case f: QuadraticObjectiveFunction => (1, f.a.toArray, f.r, packed(f.P), 2.0)
      case f if f.getClass.getName.startsWith("cvx.Dist_KL$") => (2, null, 0.0, null, 2.0)
      case _ if pNorm.isDefined => (4, null, 0.0, null, pNorm.get)
      case _ => return None
    }
    val feasible = cnts match { case c: FeasiblePoint => Some(c.feasiblePoint); case _ => None }
    val handle = CvxbNative.defaultHandle
    val p = eqs.map(_.A.rows).getOrElse(0)
    val problem = CvxbNative.problemCreate(handle, n, lin.length, p, kind, a, r, pm, pw, G.data, gR, ub,
      eqs.map(e => packed(e.A)).orNull, eqs.map(_.b.toArray).orNull, feasible.map(_.toArray).orNull,
      cnts.pointWhereDefined.toArray, mq, qP, qA, qR, qUb)
    def report(x0: DenseVector[Double]): FeasibilityReport = {
      val s = cnts.constraints.map(c => c.valueAt(x0) - c.ub).max
      FeasibilityReport(x0, DenseVector(s), s < 0, cnts, eqs.map(e => norm(e.A * x0 - e.b)))
    }
    Some(new GpuSolver(handle, problem, n, lin.length + mq, p, solverType, pars, logger,
                       feasible.getOrElse(cnts.pointWhereDefined), report))
  }
}
