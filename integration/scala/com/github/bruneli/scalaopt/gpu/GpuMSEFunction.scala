/*
 * GpuMSEFunction.scala — MSEFunction (core/.../function/MSEFunction.scala:12-76) answered by the CUDA kernels.
 * Usage is unchanged:   LevenbergMarquardt.minimize(GpuMSEFunction(GaussExpBkg, gpuData), p0)
 *                       BFGS.minimize(GpuMSEFunction(Logistic, gpuData), w0)
 *                       LevenbergMarquardt.minimize(GpuMSEFunction(MlpCe, irisOnGpu, decay = 0.05), w0)
 * UNVERIFIED SOURCE (no JVM here); C++ twin: scalaopt_b200/host/gpu.hpp (compiled, tested on B200), which this file follows
 * function by function: fused K1 pass with a memo on the bit pattern of p, speculative K2 / TSQR pass at Levenberg-Marquardt
 * trial points, and the later trial points of a StrongWolfe search scored in batches by K3 (scoreRay).  The twin keeps the last
 * 8 evaluated points, this file the last one (+ 8 losses).
 */
package com.github.bruneli.scalaopt.gpu

import com.github.bruneli.scalaopt.core._
import com.github.bruneli.scalaopt.core.SeqDataSetConverter._
import com.github.bruneli.scalaopt.core.function.MSEFunction
import com.github.bruneli.scalaopt.core.linalg.AugmentedRow
import com.github.bruneli.scalaopt.core.variable._

/**
 * @param useTsqr   take the (n+1)-row system from so_eval_qr (Householder on the rows: the condition number of J is not
 *                  squared) instead of K2 + Cholesky
 * @param speculate Levenberg-Marquardt: score the first trial point after a Jacobian with a K2 / TSQR pass whose sums serve
 *                  the Jacobian request that follows an accepted step (one pass per iteration instead of two)
 * @param decay     FFNeuralNetworkTrainer's weight decay (std-apps/.../nnet/FFNeuralNetworkTrainer.scala:57-69,147-153)
 * @param lineBatch step sizes scored per K3 pass once a line search needs a second trial point: the requested one and the
 *                  continuation of the bracket schedule alpha <- c3 alpha (linesearch/StrongWolfe.scala:114)
 */
case class GpuMSEFunction(family: Int, gpuData: GpuDataSet, useTsqr: Boolean = false, speculate: Boolean = true,
                          decay: Double = 0.0, lineBatch: Int = 3) extends MSEFunction {

  override val data: DataSet[DataPoint] = gpuData

  private def raw(p: UnconstrainedVariablesType): Array[Double] = p.force.coordinates
  private def check(rc: Int): Unit = if (rc != 0) throw new RuntimeException(NativeLib.lastError(gpuData.ctx))
  private def m: Double = gpuData.size.toDouble

  /** loss = lossScale * r^T r / m where the loss is a function of r^T r (0: cross-entropy families) */
  private val lossScale: Double = family match {
    case NativeLib.MlpMse => 1.0                               // FFNeuralNetwork.scala:78-80, not halved
    case NativeLib.Logistic | NativeLib.MlpCe | NativeLib.SoftMax => 0.0
    case _ => 0.5                                              // MSEFunction.scala:34
  }
  private def decayLoss(x: Array[Double]): Double = if (decay > 0.0) 0.5 * decay * x.map(v => v * v).sum / m else 0.0

  // memo on the bit pattern of p: LineSearchPoint re-asks the same point up to 4 times (variable/Point.scala:48-63)
  private var lastP: Array[Double] = Array.empty
  private var lastOut: Array[Double] = Array.empty      // (loss, gradient) incl. the decay terms
  private var lossOnly = false                           // Levenberg-Marquardt only asks for losses between Jacobians
  private var trialsSinceJacobian = 0
  private var specP: Array[Double] = Array.empty         // trial point whose K2 / TSQR sums are kept
  private var specSystem: Array[Double] = Array.empty
  private var lossMemo = scala.collection.immutable.ListMap.empty[Seq[Double], Double]   // insertion-ordered: takeRight keeps the newest

  private def lossGrad(p: UnconstrainedVariablesType): Array[Double] = {
    val x = raw(p)
    if (!java.util.Arrays.equals(x, lastP)) {
      val out = new Array[Double](1 + x.length)
      check(NativeLib.evalLossGrad(gpuData.handle, family, x, out, true))   // K1: one fused pass
      if (decay > 0.0) { out(0) += decayLoss(x); for (i <- x.indices) out(1 + i) += x(i) * decay / m }
      lastP = x.clone(); lastOut = out
      if (onRayBeta(x).isDefined) rayTrials += 1
    }
    lastOut
  }

  // ---- the ray a line search walks (set by dirder): later trial points on it are scored in batches by K3 ----
  private var rayOrigin: Array[Double] = Array.empty
  private var rayD: Array[Double] = Array.empty
  private var rayTrials = 0                              // distinct points evaluated on the current ray
  private var line = Vector.empty[(Double, Double, Double)]   // (beta, phi, phi') of the points scored so far, decay included

  private def dot(a: Array[Double], b: Array[Double]): Double = a.indices.map(i => a(i) * b(i)).sum
  private def sameDir(d: Array[Double]): Boolean = rayD.nonEmpty && java.util.Arrays.equals(d, rayD)
  private def noteRay(x: Array[Double], d: Array[Double]): Unit =
    if (!sameDir(d)) { rayOrigin = x.clone(); rayD = d.clone(); rayTrials = 1; line = Vector.empty }

  /** beta with x = origin + beta d, to a few ulp of the point's magnitude */
  private def onRayBeta(x: Array[Double]): Option[Double] =
    if (rayD.isEmpty || x.length != rayD.length) None
    else {
      val dd = dot(rayD, rayD)
      if (!(dd > 0.0)) None
      else {
        val diff = x.indices.map(i => x(i) - rayOrigin(i)).toArray
        val beta = dot(diff, rayD) / dd
        val err = x.indices.map(i => math.abs(diff(i) - beta * rayD(i))).max
        val scale = x.indices.map(i => math.abs(x(i)) + math.abs(beta * rayD(i))).max
        if (err > 64.0 * 2.220446049250313e-16 * math.max(scale, 1e-300)) None else Some(beta)
      }
    }

  private def findOnRay(x: Array[Double]): Option[(Double, Double, Double)] =
    if (line.isEmpty) None
    else onRayBeta(x).flatMap(beta => line.find(le => math.abs(le._1 - beta) <= 1e-12 * math.max(1.0, math.abs(beta))))

  /** one K3 pass: this step size and the next lineBatch - 1 of the bracket schedule beta <- 2 beta + 1 */
  private def scoreRay(beta: Double): (Double, Double, Double) = {
    val al = Iterator.iterate(beta)(b => 2.0 * b + 1.0).take(if (rayTrials >= 1) lineBatch else 1).toArray
    val out = new Array[Double](2 * al.length)           // phi[0, k), phi'[k, 2k)
    check(NativeLib.evalLine(gpuData.handle, family, rayOrigin, rayD, al, out))
    val entries = al.indices.map { i =>
      // the device scores the data term; the trainer's weight decay belongs to every point of the ray as well
      val xa = rayOrigin.indices.map(j => rayOrigin(j) + rayD(j) * al(i)).toArray
      val (penPhi, penDphi) = if (decay > 0.0) (decayLoss(xa), decay * dot(xa, rayD) / m) else (0.0, 0.0)
      (al(i), out(i) + penPhi, out(al.length + i) + penDphi)
    }
    line = line ++ entries
    rayTrials += 1
    entries.head
  }

  def apply(p: UnconstrainedVariablesType, x: InputsType): OutputsType =
    throw new UnsupportedOperationException("per-row model evaluation stays on the device")
  def residual(p: UnconstrainedVariablesType, xy: DataPoint): Double =
    throw new UnsupportedOperationException("per-row residuals stay on the device")
  def jacobianAndResidual(p: UnconstrainedVariablesType, xy: DataPoint): (InputsType, Output) =
    throw new UnsupportedOperationException("per-row Jacobians stay on the device")

  /** sum r^2/2 / m, exactly MSEFunction.apply (MSEFunction.scala:33-36); cross-entropy families: the trainer's loss */
  override def apply(p: UnconstrainedVariablesType): Double = {
    val x = raw(p)
    lossMemo.get(x.toSeq) match {
      case Some(l) => l
      case None if findOnRay(x).isDefined => findOnRay(x).get._2          // scored by an earlier K3 pass
      case None if onRayBeta(x).isDefined => scoreRay(onRayBeta(x).get)._2
      case None =>
        val l =
          if (lossOnly && speculate && lossScale > 0.0 && trialsSinceJacobian == 0 && (useTsqr || x.length <= 16)) {
            trialsSinceJacobian += 1
            val (sys, rtr) = system(x)                    // K2 or TSQR pass; kept for the Jacobian request that follows
            specP = x.clone(); specSystem = sys
            rtr * lossScale / m + decayLoss(x)
          } else if (lossOnly) {
            trialsSinceJacobian += 1
            val out = new Array[Double](1)
            check(NativeLib.evalLossGrad(gpuData.handle, family, x, out, false))
            out(0) + decayLoss(x)
          } else lossGrad(p)(0)
        lossMemo = lossMemo.updated(x.toSeq, l).takeRight(8)
        l
    }
  }

  override def gradient(p: UnconstrainedVariablesType): UnconstrainedVariablesType =
    new UnconstrainedVariables(lossGrad(p).drop(1))

  /** gradient(p) dot d (ObjectiveFunctionWithGradient.scala:37-39); also tells the objective which ray the search walks */
  override def dirder(p: UnconstrainedVariablesType, d: UnconstrainedVariablesType): Double = {
    val x = raw(p)
    val dd = raw(d)
    if (java.util.Arrays.equals(x, lastP)) { noteRay(x, dd); dot(lastOut.drop(1), dd) }
    else if (sameDir(dd) && findOnRay(x).isDefined) findOnRay(x).get._3
    else { val g = lossGrad(p).drop(1); noteRay(x, dd); dot(g, dd) }
  }

  override def dirHessian(p: UnconstrainedVariablesType, d: UnconstrainedVariablesType): UnconstrainedVariablesType = {
    val out = new Array[Double](p.length)
    check(NativeLib.evalHessVec(gpuData.handle, family, raw(p), raw(d), out))   // K4 (exact product)
    // the trainer's dirHessian differences gradients that contain w * decay / m: + decay * d / m
    if (decay > 0.0) { val dd = raw(d); for (i <- out.indices) out(i) += decay * dd(i) / m }
    new UnconstrainedVariables(out)
  }

  /** rows of the compressed system, flattened (n+1) x (n+1), and r^T r */
  private def system(x: Array[Double]): (Array[Double], Double) = {
    val n = x.length
    if (useTsqr) {
      val r = new Array[Double]((n + 1) * (n + 1))
      check(NativeLib.evalQr(gpuData.handle, family, x, r))                      // Householder TSQR on the rows
      (r, (0 to n).map(i => r(i * (n + 1) + n)).map(v => v * v).sum)
    } else {
      val out = new Array[Double](n * n + n + 1)
      check(NativeLib.evalNormalEq(gpuData.handle, family, x, out))              // K2
      val jtj = out.take(n * n); val jtr = out.slice(n * n, n * n + n); val rtr = out(n * n + n)
      // the decay rows sqrt(decay) e_i | 0 add decay to the Gram's diagonal: folded in BEFORE the factorisation, J alone
      // is often rank-deficient for a network while [J; sqrt(decay) I] never is
      if (decay > 0.0) for (i <- 0 until n) jtj(i * n + i) += decay
      val r = Array.ofDim[Double](n, n)
      val q = new Array[Double](n)
      for (i <- 0 until n) {
        var dsum = jtj(i * n + i); for (k <- 0 until i) dsum -= r(k)(i) * r(k)(i)
        if (dsum > 0.0) {
          val rii = math.sqrt(dsum); r(i)(i) = rii
          for (j <- i + 1 until n) { var s = jtj(i * n + j); for (k <- 0 until i) s -= r(k)(i) * r(k)(j); r(i)(j) = s / rii }
          var s = jtr(i); for (k <- 0 until i) s -= r(k)(i) * q(k); q(i) = s / rii
        }
      }
      val rho = math.sqrt(math.max(0.0, rtr - q.map(v => v * v).sum))
      val flat = new Array[Double]((n + 1) * (n + 1))
      for (i <- 0 until n) { for (j <- 0 until n) flat(i * (n + 1) + j) = r(i)(j); flat(i * (n + 1) + n) = q(i) }
      flat(n * (n + 1) + n) = rho
      (flat, rtr)
    }
  }

  /** (n+1)-row system with the same R, QtB, column norms and ||b|| as the m-row one (DESIGN.md section 1) */
  def jacobianAndResidualsMatrix(p: UnconstrainedVariablesType): DataSet[AugmentedRow] = {
    val x = raw(p)
    val n = x.length
    lossOnly = true
    trialsSinceJacobian = 0
    val (sys, rtr) = if (java.util.Arrays.equals(x, specP)) (specSystem, Double.NaN) else system(x)
    if (lossScale > 0.0 && !rtr.isNaN) lossMemo = lossMemo.updated(x.toSeq, rtr * lossScale / m + decayLoss(x))
    val rows = (0 to n).map(i => AugmentedRow(new Inputs(sys.slice(i * (n + 1), i * (n + 1) + n)), Output(sys(i * (n + 1) + n)), i.toLong))
    // TSQR route: the triangle comes from the rows, the decay rows are appended as the trainer does (Trainer.scala:147-153)
    val penalty = if (useTsqr && decay > 0.0) (0 until n).map { i =>
      val e = new Array[Double](n); e(i) = math.sqrt(decay); AugmentedRow(new Inputs(e), Output(0.0), (n + 1 + i).toLong)
    } else Seq.empty
    rows ++ penalty
  }
}
