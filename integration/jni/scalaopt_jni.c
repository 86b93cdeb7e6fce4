/*
 * scalaopt_jni.c — JNI glue between com.github.bruneli.scalaopt.gpu.NativeLib and libscalaopt_cuda.so.
 * Compiled only where a JDK is present (this image has none); type-checked against include/scalaopt_cuda.h with a stand-in
 * <jni.h> by tests/test_abi.py, never run.
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include scalaopt_jni.c \
 *       -L../../scalaopt_b200 -lscalaopt_cuda -o libscalaopt_jni.so
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#include <jni.h>
#include "scalaopt_cuda.h"

#define JFN(name) Java_com_github_bruneli_scalaopt_gpu_NativeLib_00024_##name

JNIEXPORT jlong JNICALL JFN(init)(JNIEnv* e, jobject o, jint ngpus, jintArray ids) {
  so_ctx* ctx = 0;
  jint* p = ids ? (*e)->GetIntArrayElements(e, ids, 0) : 0;
  int rc = so_init(ngpus, (const int*)p, &ctx);
  if (p) (*e)->ReleaseIntArrayElements(e, ids, p, JNI_ABORT);
  return rc == SO_OK ? (jlong)(intptr_t)ctx : 0;
}
JNIEXPORT void JNICALL JFN(shutdown)(JNIEnv* e, jobject o, jlong ctx) { so_shutdown((so_ctx*)(intptr_t)ctx); }
JNIEXPORT jstring JNICALL JFN(lastError)(JNIEnv* e, jobject o, jlong ctx) {
  return (*e)->NewStringUTF(e, so_last_error((so_ctx*)(intptr_t)ctx));
}
JNIEXPORT jlong JNICALL JFN(datasetCreate)(JNIEnv* e, jobject o, jlong ctx, jlong m, jint f) {
  so_dataset* ds = 0;
  return so_dataset_create((so_ctx*)(intptr_t)ctx, m, f, &ds) == SO_OK ? (jlong)(intptr_t)ds : 0;
}
/* Bulk transfers do NOT hold a JNI critical section: a multi-GB host-to-device copy inside GetPrimitiveArrayCritical would
 * stall the JVM's garbage collector for its whole duration.  Get/ReleaseDoubleArrayElements may copy; callers that ingest
 * at PCIe speed use datasetAppendDirect with direct ByteBuffers over so_host_alloc'ed (pinned) memory. */
JNIEXPORT jint JNICALL JFN(datasetAppend)(JNIEnv* e, jobject o, jlong ds, jdoubleArray x, jdoubleArray y, jlong rows) {
  jdouble* px = (*e)->GetDoubleArrayElements(e, x, 0);
  jdouble* py = (*e)->GetDoubleArrayElements(e, y, 0);
  int rc = so_dataset_append((so_dataset*)(intptr_t)ds, px, py, rows);
  (*e)->ReleaseDoubleArrayElements(e, y, py, JNI_ABORT);
  (*e)->ReleaseDoubleArrayElements(e, x, px, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(datasetAppendDirect)(JNIEnv* e, jobject o, jlong ds, jobject xBuf, jobject yBuf, jlong rows) {
  const double* px = (const double*)(*e)->GetDirectBufferAddress(e, xBuf);
  const double* py = (const double*)(*e)->GetDirectBufferAddress(e, yBuf);
  if (!px || !py) return SO_EINVAL;
  return so_dataset_append((so_dataset*)(intptr_t)ds, px, py, rows);
}
JNIEXPORT jlong JNICALL JFN(datasetCreateMulti)(JNIEnv* e, jobject o, jlong ctx, jlong m, jint f, jint ny) {
  so_dataset* ds = 0;
  return so_dataset_create_multi((so_ctx*)(intptr_t)ctx, m, f, ny, &ds) == SO_OK ? (jlong)(intptr_t)ds : 0;
}
JNIEXPORT jint JNICALL JFN(datasetRead)(JNIEnv* e, jobject o, jlong ds, jlong b, jlong rows, jdoubleArray x, jdoubleArray y) {
  jdouble* px = (*e)->GetDoubleArrayElements(e, x, 0);
  jdouble* py = (*e)->GetDoubleArrayElements(e, y, 0);
  int rc = so_dataset_read((so_dataset*)(intptr_t)ds, b, rows, px, py);
  (*e)->ReleaseDoubleArrayElements(e, y, py, 0);
  (*e)->ReleaseDoubleArrayElements(e, x, px, 0);
  return rc;
}
JNIEXPORT jlong JNICALL JFN(datasetSize)(JNIEnv* e, jobject o, jlong ds) {
  int64_t m = -1; so_dataset_size((so_dataset*)(intptr_t)ds, &m); return m;
}
/* Evaluations: parameter vectors and results are a few doubles (n, n*n + n + 1, 2k, ...), so Get/ReleaseDoubleArrayElements —
 * which may copy — costs nothing against a pass over the data, and unlike GetPrimitiveArrayCritical it may be held across a
 * blocking call: so_eval_* blocks for the duration of the kernel (milliseconds to seconds, and in a multi-rank context until
 * the slowest rank has posted its sums), which must not happen inside a JNI critical section (the GC would be held off). */
JNIEXPORT void JNICALL JFN(datasetFree)(JNIEnv* e, jobject o, jlong ds) { so_dataset_free((so_dataset*)(intptr_t)ds); }
JNIEXPORT jint JNICALL JFN(evalLossGrad)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jdoubleArray out, jboolean g) {
  jsize n = (*e)->GetArrayLength(e, p);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_loss_grad((so_dataset*)(intptr_t)ds, fam, pp, n, po, g ? po + 1 : 0);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(evalNormalEq)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jdoubleArray out) {
  jsize n = (*e)->GetArrayLength(e, p);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_normal_eq((so_dataset*)(intptr_t)ds, fam, pp, n, po, po + n * n, po + n * n + n);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(evalLine)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jdoubleArray d, jdoubleArray a, jdoubleArray out) {
  jsize n = (*e)->GetArrayLength(e, p), k = (*e)->GetArrayLength(e, a);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* pd = (*e)->GetDoubleArrayElements(e, d, 0);
  double* pa = (*e)->GetDoubleArrayElements(e, a, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_line((so_dataset*)(intptr_t)ds, fam, pp, pd, n, pa, k, po, po + k);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, a, pa, JNI_ABORT);
  (*e)->ReleaseDoubleArrayElements(e, d, pd, JNI_ABORT);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(evalHessVec)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jdoubleArray d, jdoubleArray out) {
  jsize n = (*e)->GetArrayLength(e, p);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* pd = (*e)->GetDoubleArrayElements(e, d, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_hess_vec((so_dataset*)(intptr_t)ds, fam, pp, pd, n, po);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, d, pd, JNI_ABORT);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(datasetGenerate)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jlong seed, jdouble sigma, jlong off) {
  jsize n = (*e)->GetArrayLength(e, p);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  int rc = so_dataset_generate((so_dataset*)(intptr_t)ds, fam, pp, n, (uint64_t)seed, sigma, off);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(evalQr)(JNIEnv* e, jobject o, jlong ds, jint fam, jdoubleArray p, jdoubleArray out) {
  jsize n = (*e)->GetArrayLength(e, p);                  /* out: (n+1)*(n+1) doubles, the R factor of [J | r] */
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_qr((so_dataset*)(intptr_t)ds, fam, pp, n, po);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(evalNll)(JNIEnv* e, jobject o, jlong ds, jint pdf, jdoubleArray p, jdoubleArray consts, jdoubleArray out) {
  jsize n = (*e)->GetArrayLength(e, p), nc = (*e)->GetArrayLength(e, consts);
  double* pp = (*e)->GetDoubleArrayElements(e, p, 0);
  double* pc = (*e)->GetDoubleArrayElements(e, consts, 0);
  double* po = (*e)->GetDoubleArrayElements(e, out, 0);
  int rc = so_eval_nll((so_dataset*)(intptr_t)ds, pdf, pp, n, pc, nc, po);
  (*e)->ReleaseDoubleArrayElements(e, out, po, 0);
  (*e)->ReleaseDoubleArrayElements(e, consts, pc, JNI_ABORT);
  (*e)->ReleaseDoubleArrayElements(e, p, pp, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL JFN(datasetLoadRaw)(JNIEnv* e, jobject o, jlong ds, jstring xPath, jstring yPath, jlong rows, jlong fileRowOffset) {
  const char* xp = (*e)->GetStringUTFChars(e, xPath, 0);
  const char* yp = (*e)->GetStringUTFChars(e, yPath, 0);
  int rc = so_dataset_load_raw((so_dataset*)(intptr_t)ds, xp, yp, rows, fileRowOffset);
  (*e)->ReleaseStringUTFChars(e, yPath, yp);
  (*e)->ReleaseStringUTFChars(e, xPath, xp);
  return rc;
}
#endif
#endif
