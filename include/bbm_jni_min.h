/*
 * bbm_jni_min.h — the minimal slice of the JNI ABI that BBMap's native plug-in uses.
 *
 * The reference's native library (reference: jni/MultiStateAligner11tsJNI.c:707-812,
 * jni/BandedAlignerJNI.c:588-757) touches exactly three JNIEnv functions:
 * GetArrayLength, GetPrimitiveArrayCritical and ReleasePrimitiveArrayCritical.
 * This header declares a JNIEnv whose function table has those three entries at
 * the slots the JNI specification assigns them (171, 222, 223; slots 0-3 are
 * reserved), so a library built against it is call-compatible with a real JVM
 * without needing a JDK at build time (there is none in this image).
 *
 * It is used by (a) libbbmapcuda.so's Java_align2_* entry points and
 * (b) the test-only oracle build of the reference C (oracle/jni_stub/jni.h
 * forwards here).  Unverified against a live JVM (no JVM in the build image).
 */
#ifndef BBM_JNI_MIN_H
#define BBM_JNI_MIN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int8_t   jbyte;
typedef int16_t  jshort;
typedef int32_t  jint;
typedef int64_t  jlong;
typedef float    jfloat;
typedef double   jdouble;
typedef uint8_t  jboolean;
typedef jint     jsize;

typedef void*    jobject;
typedef jobject  jclass;
typedef jobject  jarray;
typedef jarray   jbyteArray;
typedef jarray   jshortArray;
typedef jarray   jintArray;
typedef jarray   jlongArray;
typedef jarray   jfloatArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

#define BBM_JNI_SLOT_GetArrayLength                 171
#define BBM_JNI_SLOT_GetPrimitiveArrayCritical      222
#define BBM_JNI_SLOT_ReleasePrimitiveArrayCritical  223
#define BBM_JNI_NUM_SLOTS                           235

struct JNINativeInterface_ {
    void* slots_0_170[171];
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);                                      /* 171 */
    void* slots_172_221[50];
    void* (*GetPrimitiveArrayCritical)(JNIEnv* env, jarray array, jboolean* isCopy);         /* 222 */
    void  (*ReleasePrimitiveArrayCritical)(JNIEnv* env, jarray array, void* carray, jint mode); /* 223 */
    void* slots_224_234[11];
};

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_FALSE  0
#define JNI_TRUE   1
#define JNI_COMMIT 1
#define JNI_ABORT  2

#ifdef __cplusplus
}
#endif
#endif /* BBM_JNI_MIN_H */
