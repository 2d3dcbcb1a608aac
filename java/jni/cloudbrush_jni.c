/* cloudbrush_jni.c -- the JNI side of java/Brush/GpuOverlap.java: one thin function per entry
 * point of include/cloudbrush_gpu.h.  Build (on a box with a JDK):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       cloudbrush_jni.c -L../../cloudbrush_b200 -lcloudbrush_gpu -o libcloudbrush_jni.so
 * Not compiled in this repository's image (no jni.h). */
#include <jni.h>
#include <stdint.h>

#include "cloudbrush_gpu.h"

#define CTX(h) ((cb_ctx*)(intptr_t)(h))

JNIEXPORT jlong JNICALL Java_Brush_GpuOverlap_nCreate(JNIEnv* env, jclass cls, jint k, jint readlen, jlong up,
                                                      jlong low, jfloat maj, jfloat pwmn, jint device, jint rank,
                                                      jint world) {
  cb_config cfg;
  cb_ctx* ctx = NULL;
  cb_default_config(&cfg);
  cfg.k = k; cfg.readlen = readlen; cfg.up_kmer = up; cfg.low_kmer = low;
  cfg.majority = maj; cfg.pwm_n = pwmn; cfg.device = device;
  cfg.rank = rank; cfg.world = world;  /* one JVM per GPU: shard `rank` of `world` (0 / 1 on a single GPU) */
  return cb_create(&cfg, &ctx) == CB_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL Java_Brush_GpuOverlap_nDestroy(JNIEnv* env, jclass cls, jlong h) { cb_destroy(CTX(h)); }

JNIEXPORT jstring JNICALL Java_Brush_GpuOverlap_nLastError(JNIEnv* env, jclass cls, jlong h) {
  return (*env)->NewStringUTF(env, cb_last_error(CTX(h)));
}

JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nLoadSfa(JNIEnv* env, jclass cls, jlong h, jstring path) {
  const char* p = (*env)->GetStringUTFChars(env, path, NULL);
  int rc = cb_load_sfa(CTX(h), p);
  (*env)->ReleaseStringUTFChars(env, path, p);
  return rc;
}

JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nPreprocess(JNIEnv* env, jclass cls, jlong h) { return cb_preprocess(CTX(h)); }
JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nBuildOverlap(JNIEnv* env, jclass cls, jlong h) { return cb_build_overlap(CTX(h)); }
JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nStringGraph(JNIEnv* env, jclass cls, jlong h) { return cb_string_graph(CTX(h)); }

JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nWriteNodes(JNIEnv* env, jclass cls, jlong h, jint ck, jstring dir) {
  const char* p = (*env)->GetStringUTFChars(env, dir, NULL);
  int rc = cb_write_nodes(CTX(h), ck, p);
  (*env)->ReleaseStringUTFChars(env, dir, p);
  return rc;
}

JNIEXPORT jlong JNICALL Java_Brush_GpuOverlap_nCounter(JNIEnv* env, jclass cls, jlong h, jstring name) {
  const char* p = (*env)->GetStringUTFChars(env, name, NULL);
  int64_t v = 0;
  int rc = cb_get_counter(CTX(h), p, &v);
  (*env)->ReleaseStringUTFChars(env, name, p);
  return rc == CB_OK ? (jlong)v : INT64_MIN;
}

JNIEXPORT jdouble JNICALL Java_Brush_GpuOverlap_nTimeMs(JNIEnv* env, jclass cls, jlong h, jstring what) {
  const char* p = (*env)->GetStringUTFChars(env, what, NULL);
  double ms = -1.0;
  cb_get_time_ms(CTX(h), p, &ms);
  (*env)->ReleaseStringUTFChars(env, what, p);
  return ms;
}

/* multi-GPU launchers (one JVM per GPU): rank 0 obtains the NCCL id, the launcher distributes the
 * 128 bytes over its own channel, every rank initialises the library's native transport */
JNIEXPORT jbyteArray JNICALL Java_Brush_GpuOverlap_nNcclUniqueId(JNIEnv* env, jclass cls) {
  jbyte id[CB_NCCL_ID_BYTES];
  if (cb_nccl_get_unique_id(id) != CB_OK) return NULL; /* the reason: cb_nccl_last_error() */
  jbyteArray out = (*env)->NewByteArray(env, CB_NCCL_ID_BYTES);
  (*env)->SetByteArrayRegion(env, out, 0, CB_NCCL_ID_BYTES, id);
  return out;
}

JNIEXPORT jstring JNICALL Java_Brush_GpuOverlap_nNcclLastError(JNIEnv* env, jclass cls) {
  return (*env)->NewStringUTF(env, cb_nccl_last_error());
}

JNIEXPORT jint JNICALL Java_Brush_GpuOverlap_nInitNccl(JNIEnv* env, jclass cls, jlong h, jbyteArray id) {
  jbyte buf[CB_NCCL_ID_BYTES];
  (*env)->GetByteArrayRegion(env, id, 0, CB_NCCL_ID_BYTES, buf);
  return cb_init_nccl(CTX(h), buf);
}
