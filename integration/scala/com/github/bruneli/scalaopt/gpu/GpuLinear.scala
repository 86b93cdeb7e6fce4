/*
 * GpuLinear.scala — linear.lm (core/.../linear/package.scala:54-63) for a GpuDataSet.
 * The reference builds m AugmentedRows ((1, x) | y) with zipWithIndex.map and hands them to QR(ab, n): 2n+1 passes over the
 * data set.  Here the m rows never leave HBM: at p = 0 the linear family's Jacobian/residual rows are s_i (-a_i | y_i) with
 * s_i = sign(y_i), so the (n+1)-row system GpuMSEFunction.jacobianAndResidualsMatrix returns carries the Gram of (-a | y)
 * and the reference's own QR(...).solution on it is -beta.  One pass (K2, or the on-device Householder TSQR with useTsqr).
 * UNVERIFIED SOURCE (no JVM here); C++ twin: gpu_lm in scalaopt_b200/host/gpu.hpp (tests/test_gpu_host.py, LinearSpec replay).
 */
package com.github.bruneli.scalaopt.gpu

import com.github.bruneli.scalaopt.core._
import com.github.bruneli.scalaopt.core.linalg.QR
import com.github.bruneli.scalaopt.core.variable.UnconstrainedVariables

import scala.util.Try

object GpuLinear {

  /** Same contract as linear.lm(data, addOrigin): a successful solution beta of X beta = y (least squares) or a failure */
  def lm(data: GpuDataSet, addOrigin: Boolean = true, useTsqr: Boolean = false): Try[UnconstrainedVariablesType] =
    Try {
      val n = if (addOrigin) data.f + 1 else data.f
      val family = if (addOrigin) NativeLib.Linear else NativeLib.LinearNoIntercept
      val zero = new UnconstrainedVariables(new Array[Double](n))
      val system = GpuMSEFunction(family, data, useTsqr = useTsqr).jacobianAndResidualsMatrix(zero)
      new UnconstrainedVariables(QR(system, n).solution.force.coordinates.map(v => -v))
    }
}
