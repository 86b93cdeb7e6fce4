/*
 * GpuMSEFunction.scala — MSEFunction (core/.../function/MSEFunction.scala:12-76) answered by the CUDA kernels.
 * Usage is unchanged:   LevenbergMarquardt.minimize(GpuMSEFunction(GaussExpBkg, gpuData), p0)
 *                       BFGS.minimize(GpuMSEFunction(Logistic, gpuData), w0)
 *                       LevenbergMarquardt.minimize(GpuMSEFunction(MlpCe, irisOnGpu, decay = 0.05), w0)
 * UNVERIFIED SOURCE (no JVM here); C++ twin: scalaopt_b200/host/gpu.hpp (compiled, tested on B200).  What the twin has
 * and this file does not: batching of the later trial points of a line search into one K3 pass (score_ray) — here every
 * LineSearchPoint is one fused K1 pass, memoised.
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
 */
case class GpuMSEFunction(family: Int, gpuData: GpuDataSet, useTsqr: Boolean = false, speculate: Boolean = true,
                          decay: Double = 0.0) extends MSEFunction {

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
    }
    lastOut
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

  override def dirder(p: UnconstrainedVariablesType, d: UnconstrainedVariablesType): Double = gradient(p) dot d

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
