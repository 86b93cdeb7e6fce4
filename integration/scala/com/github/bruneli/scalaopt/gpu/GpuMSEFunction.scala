/*
 * GpuMSEFunction.scala — MSEFunction (core/.../function/MSEFunction.scala:12-76) answered by the CUDA kernels.
 * Usage is unchanged:   LevenbergMarquardt.minimize(GpuMSEFunction(GaussExpBkg, gpuData), p0)
 *                       BFGS.minimize(GpuMSEFunction(Logistic, gpuData), w0)
 * UNVERIFIED SOURCE (no JVM here); C++ twin: scalaopt_b200/host/gpu.hpp (compiled, tested on B200).
 */
package com.github.bruneli.scalaopt.gpu

import com.github.bruneli.scalaopt.core._
import com.github.bruneli.scalaopt.core.SeqDataSetConverter._
import com.github.bruneli.scalaopt.core.function.MSEFunction
import com.github.bruneli.scalaopt.core.linalg.AugmentedRow
import com.github.bruneli.scalaopt.core.variable._

case class GpuMSEFunction(family: Int, gpuData: GpuDataSet) extends MSEFunction {

  override val data: DataSet[DataPoint] = gpuData

  private def raw(p: UnconstrainedVariablesType): Array[Double] = p.force.coordinates
  private def check(rc: Int): Unit = if (rc != 0) throw new RuntimeException(NativeLib.lastError(gpuData.ctx))

  // memo on the bit pattern of p: LineSearchPoint re-asks the same point up to 4 times (variable/Point.scala:48-63)
  private var lastP: Array[Double] = Array.empty
  private var lastOut: Array[Double] = Array.empty
  private def lossGrad(p: UnconstrainedVariablesType): Array[Double] = {
    val x = raw(p)
    if (!java.util.Arrays.equals(x, lastP)) {
      val out = new Array[Double](1 + x.length)
      check(NativeLib.evalLossGrad(gpuData.handle, family, x, out, true))   // K1: one fused pass
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

  /** sum r^2/2 / m, exactly MSEFunction.apply (MSEFunction.scala:33-36) */
  override def apply(p: UnconstrainedVariablesType): Double = lossGrad(p)(0)

  override def gradient(p: UnconstrainedVariablesType): UnconstrainedVariablesType =
    new UnconstrainedVariables(lossGrad(p).drop(1))

  override def dirder(p: UnconstrainedVariablesType, d: UnconstrainedVariablesType): Double = gradient(p) dot d

  override def dirHessian(p: UnconstrainedVariablesType, d: UnconstrainedVariablesType): UnconstrainedVariablesType = {
    val out = new Array[Double](p.length)
    check(NativeLib.evalHessVec(gpuData.handle, family, raw(p), raw(d), out))   // K4
    new UnconstrainedVariables(out)
  }

  /** (n+1)-row system with the same R, QtB, column norms and ||b|| as the m-row one (DESIGN.md §1) */
  def jacobianAndResidualsMatrix(p: UnconstrainedVariablesType): DataSet[AugmentedRow] = {
    val n = p.length
    val out = new Array[Double](n * n + n + 1)
    check(NativeLib.evalNormalEq(gpuData.handle, family, raw(p), out))          // K2
    val (jtj, jtr, rtr) = (out.take(n * n), out.slice(n * n, n * n + n), out(n * n + n))
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
    val rows = (0 until n).map(i => AugmentedRow(new Inputs(r(i)), Output(q(i)), i.toLong)) :+
      AugmentedRow(new Inputs(new Array[Double](n)), Output(rho), n.toLong)
    rows
  }
}
