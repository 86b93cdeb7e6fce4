/*
 * NativeLib.scala — JNI declarations for libscalaopt_cuda.so (include/scalaopt_cuda.h).
 * UNVERIFIED SOURCE: this image has no JVM / scalac / jni.h; the executable stand-in for this caller is
 * scalaopt_b200/native.py (ctypes), which issues exactly the same C-ABI calls.
 */
package com.github.bruneli.scalaopt.gpu

object NativeLib {
  System.loadLibrary("scalaopt_jni") // integration/jni/scalaopt_jni.c, links libscalaopt_cuda.so

  @native def init(ngpus: Int, deviceIds: Array[Int]): Long
  @native def shutdown(ctx: Long): Unit
  @native def lastError(ctx: Long): String
  @native def datasetCreate(ctx: Long, m: Long, f: Int): Long
  /** y has ny outputs per observation, row-major (so_dataset_create_multi; network families only) */
  @native def datasetCreateMulti(ctx: Long, m: Long, f: Int, ny: Int): Long
  @native def datasetAppend(ds: Long, x: Array[Double], y: Array[Double], rows: Long): Int
  /** direct ByteBuffers over pinned memory: full PCIe speed, no JVM-side copy, no GC pause */
  @native def datasetAppendDirect(ds: Long, x: java.nio.ByteBuffer, y: java.nio.ByteBuffer, rows: Long): Int
  @native def datasetGenerate(ds: Long, family: Int, pTrue: Array[Double], seed: Long, sigma: Double, rowOffset: Long): Int
  @native def datasetRead(ds: Long, rowBegin: Long, rows: Long, x: Array[Double], y: Array[Double]): Int
  @native def datasetSize(ds: Long): Long
  @native def datasetFree(ds: Long): Unit
  /** out = (loss, grad_0 .. grad_{n-1}); wantGrad = false leaves the gradient untouched */
  @native def evalLossGrad(ds: Long, family: Int, p: Array[Double], out: Array[Double], wantGrad: Boolean): Int
  /** out = (JtJ row-major n*n, Jtr n, rtr) */
  @native def evalNormalEq(ds: Long, family: Int, p: Array[Double], out: Array[Double]): Int
  /** out = (phi_0..phi_{k-1}, dphi_0..dphi_{k-1}) */
  @native def evalLine(ds: Long, family: Int, p: Array[Double], d: Array[Double], alpha: Array[Double], out: Array[Double]): Int
  @native def evalHessVec(ds: Long, family: Int, p: Array[Double], d: Array[Double], out: Array[Double]): Int
  /** out: (n+1)*(n+1) doubles, row-major upper triangular R factor of [J | r] (so_eval_qr: on-device Householder TSQR) */
  @native def evalQr(ds: Long, family: Int, p: Array[Double], out: Array[Double]): Int
  /** out(0) = sum -2 log pdf(p, x_i) (stats/package.scala:97-103) */
  @native def evalNll(ds: Long, pdf: Int, p: Array[Double], consts: Array[Double], out: Array[Double]): Int
  @native def datasetLoadRaw(ds: Long, xPath: String, yPath: String, rows: Long, fileRowOffset: Long): Int

  val Linear = 0; val LinearNoIntercept = 1; val Logistic = 2; val ExpDecay = 3
  val Gauss = 4; val GaussExpBkg = 5; val Poly = 6; val MlpMse = 7; val MlpCe = 8; val SoftMax = 9
  val PdfGaussExpBkg = 0
}
