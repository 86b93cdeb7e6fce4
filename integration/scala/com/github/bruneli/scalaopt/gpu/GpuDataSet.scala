/*
 * GpuDataSet.scala — a DataSet[DataPoint] whose observations live in the HBM of 1-8 B200s.
 * Drop-in for SeqDataSet (core/.../SeqDataSetConverter.scala:28-61) and SparkDataSet
 * (spark-apps/.../SparkDataSetConverter.scala:32-81).  UNVERIFIED SOURCE (no JVM here); the C++ twin that IS
 * compiled and tested is scalaopt_b200/host/gpu.hpp.
 */
package com.github.bruneli.scalaopt.gpu

import com.github.bruneli.scalaopt.core._
import com.github.bruneli.scalaopt.core.SeqDataSetConverter._
import com.github.bruneli.scalaopt.core.variable.{DataPoint, Inputs, Output}

import scala.reflect.ClassTag

class GpuDataSet(val ctx: Long, val handle: Long, val f: Int) extends DataSet[DataPoint] {

  private val collectLimit = 1L << 20

  private def check(rc: Int): Unit =
    if (rc != 0) throw new RuntimeException(NativeLib.lastError(ctx))

  override def size: Long = NativeLib.datasetSize(handle)

  override def head: DataPoint = window(0, 1).head

  /** rows [begin, begin + rows) streamed back to the host */
  def window(begin: Long, rows: Long): Seq[DataPoint] = {
    val x = new Array[Double]((rows * f).toInt)
    val y = new Array[Double](rows.toInt)
    check(NativeLib.datasetRead(handle, begin, rows, x, y))
    (0 until rows.toInt).map(i => DataPoint(new Inputs(x.slice(i * f, (i + 1) * f)), Output(y(i))))
  }

  override def collect(): Seq[DataPoint] = {
    if (size > collectLimit)
      throw new UnsupportedOperationException("GpuDataSet.collect: too large to materialise on the host")
    window(0, size)
  }

  // Closures cannot run on the device and there is no CPU evaluation path: bounded streaming or an exception.
  override def aggregate[B: ClassTag](z: => B)(seqop: (B, DataPoint) => B, combop: (B, B) => B): B =
    collect().aggregate(z)(seqop, combop)
  override def filter(p: DataPoint => Boolean): DataSet[DataPoint] = collect() filter p
  override def flatMap(fn: DataPoint => TraversableOnce[DataPoint]): DataSet[DataPoint] = collect() flatMap fn
  override def foreach(fn: DataPoint => Unit): Unit = collect() foreach fn
  override def map[B: ClassTag](fn: DataPoint => B): DataSet[B] = collect() map fn
  override def reduce(op: (DataPoint, DataPoint) => DataPoint): DataPoint = collect() reduce op
  override def zip[B: ClassTag](that: DataSet[B]): DataSet[(DataPoint, B)] = collect().zip(that.collect())
  override def zipWithIndex: DataSet[(DataPoint, Long)] = collect().zipWithIndex.map { case (a, i) => (a, i.toLong) }

  override def ++(that: DataSet[DataPoint]): DataSet[DataPoint] = {
    val extra = that.collect()
    val out = GpuDataSet.allocate(ctx, size + extra.size, f)
    var b = 0L
    while (b < size) {
      val rows = math.min(collectLimit, size - b)
      out.append(window(b, rows)); b += rows
    }
    out.append(extra)
    out
  }

  def append(points: Seq[DataPoint]): Unit = {
    val x = points.flatMap(_.x.force.coordinates).toArray
    val y = points.map(_.y.coordinate(0)).toArray
    check(NativeLib.datasetAppend(handle, x, y, points.size))
  }

  def free(): Unit = NativeLib.datasetFree(handle)
}

object GpuDataSet {
  def allocate(ctx: Long, m: Long, f: Int): GpuDataSet = new GpuDataSet(ctx, NativeLib.datasetCreate(ctx, m, f), f)

  /** SeqDataSet / SparkDataSet -> GpuDataSet */
  def from(ctx: Long, data: DataSet[DataPoint]): GpuDataSet = {
    val f = data.head.x.length
    val out = allocate(ctx, data.size, f)
    out.append(data.collect())
    out
  }
}
