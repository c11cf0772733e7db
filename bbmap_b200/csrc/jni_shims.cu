// jni_shims.cu — the JNI entry points of the reference's native library, re-implemented on top of the CUDA path.
//
// Same exported symbols and signatures as reference jni/MultiStateAligner11tsJNI.c:707-812
// (header jni/align2_MultiStateAligner11tsJNI.h:165-174), so `-Djava.library.path=` can point at a directory holding
// libbbmapcuda.so under the name the loader expects (System.loadLibrary("bbtoolsjni"),
// current/align2/MultiStateAligner11tsJNI.java:11-14) — see INTEGRATION.md.  Built against include/bbm_jni_min.h (the three
// JNIEnv slots the reference uses, at their specified indices); unverified against a live JVM (none in this image).
//
// One context PER CALLING THREAD on device $BBM_DEVICE (default 0), created at the thread's first call: the reference's entry points are re-entrant
// and BBMap keeps one MSA per mapping thread (AbstractMapThread.java:133-136), so concurrent Java threads must not serialise on a shared context.
// This is still the compatibility path (one alignment per call, ~150 us each); throughput callers use bbm_map_batch_* / bbm_msa_batch_*.
// Contexts are not destroyed at thread exit (mapping threads live as long as the run; tearing CUDA objects down from a dying thread's TLS destructor
// after the runtime has begun to unload is not safe).
#include <cstdlib>
#include <cstdio>
#include "../../include/bbm_jni_min.h"
#include "../../include/bbmap_cuda.h"

static thread_local bbm_ctx* t_ctx = nullptr;
static thread_local bool t_tried = false;

static bbm_ctx* default_ctx() {
    if (!t_tried) {
        t_tried = true;
        const char* d = getenv("BBM_DEVICE");
        const int rc = bbm_init(d ? atoi(d) : 0, &t_ctx);
        if (rc) { fprintf(stderr, "libbbmapcuda: %s\n", bbm_last_error()); t_ctx = nullptr; }
    }
    return t_ctx;
}

extern "C" bbm_ctx* bbm_default_ctx(void) { return default_ctx(); }

extern "C" JNIEXPORT void JNICALL Java_align2_MultiStateAligner11tsJNI_fillUnlimitedJNI(
    JNIEnv* env, jobject obj, jbyteArray read, jbyteArray ref, jint refStartLoc, jint refEndLoc, jintArray result,
    jlongArray iterationsUnlimited, jintArray packed, jintArray POINTSoff_SUB_ARRAY, jintArray POINTSoff_INS_ARRAY,
    jint maxRows, jint maxColumns) {
    (void)obj; (void)POINTSoff_SUB_ARRAY; (void)POINTSoff_INS_ARRAY;   // tables are the fixed 11ts constants on the device
    bbm_ctx* c = default_ctx();
    const jsize rlen = (*env)->GetArrayLength(env, read), reflen = (*env)->GetArrayLength(env, ref);
    jint* jpacked = (jint*)(*env)->GetPrimitiveArrayCritical(env, packed, nullptr);
    jbyte* jread = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, read, nullptr);
    jbyte* jref = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, ref, nullptr);
    jint* jresult = (jint*)(*env)->GetPrimitiveArrayCritical(env, result, nullptr);
    jlong* jit = (jlong*)(*env)->GetPrimitiveArrayCritical(env, iterationsUnlimited, nullptr);
    int rc = c ? bbm_fillUnlimited(c, jread, jref, rlen, reflen, refStartLoc, refEndLoc, jresult, (int64_t*)jit, jpacked, maxRows, maxColumns)
               : BBM_E_NODEVICE;
    if (rc) { fprintf(stderr, "libbbmapcuda: fillUnlimitedJNI failed (%d): %s\n", rc, bbm_last_error()); jresult[0] = -1; }
    (*env)->ReleasePrimitiveArrayCritical(env, result, jresult, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, iterationsUnlimited, jit, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, read, jread, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, ref, jref, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, packed, jpacked, 0);
}

extern "C" JNIEXPORT void JNICALL Java_align2_MultiStateAligner11tsJNI_fillLimitedXJNI(
    JNIEnv* env, jobject obj, jbyteArray read, jbyteArray ref, jint refStartLoc, jint refEndLoc, jint minScore, jintArray result,
    jlongArray iterationsLimited, jintArray packed, jintArray POINTSoff_SUB_ARRAY, jintArray POINTSoff_INS_ARRAY, jint maxRows,
    jint maxColumns, jint bandwidth, jfloat bandwidthRatio, jintArray vertLimit, jintArray horizLimit, jbyteArray baseToNumber,
    jintArray POINTSoff_INS_ARRAY_C) {
    (void)obj; (void)POINTSoff_SUB_ARRAY; (void)POINTSoff_INS_ARRAY; (void)baseToNumber; (void)POINTSoff_INS_ARRAY_C;
    bbm_ctx* c = default_ctx();
    const jsize rlen = (*env)->GetArrayLength(env, read), reflen = (*env)->GetArrayLength(env, ref);
    jint* jpacked = (jint*)(*env)->GetPrimitiveArrayCritical(env, packed, nullptr);
    jbyte* jread = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, read, nullptr);
    jbyte* jref = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, ref, nullptr);
    jint* jresult = (jint*)(*env)->GetPrimitiveArrayCritical(env, result, nullptr);
    jlong* jit = (jlong*)(*env)->GetPrimitiveArrayCritical(env, iterationsLimited, nullptr);
    jint* jvl = (jint*)(*env)->GetPrimitiveArrayCritical(env, vertLimit, nullptr);
    jint* jhl = (jint*)(*env)->GetPrimitiveArrayCritical(env, horizLimit, nullptr);
    int rc = c ? bbm_fillLimitedX(c, jread, jref, rlen, reflen, refStartLoc, refEndLoc, minScore, jresult, (int64_t*)jit, jpacked,
                                  maxRows, maxColumns, bandwidth, bandwidthRatio, jvl, jhl)
               : BBM_E_NODEVICE;
    if (rc) { fprintf(stderr, "libbbmapcuda: fillLimitedXJNI failed (%d): %s\n", rc, bbm_last_error()); jresult[4] = 1; }
    (*env)->ReleasePrimitiveArrayCritical(env, result, jresult, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, iterationsLimited, jit, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, read, jread, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, ref, jref, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, packed, jpacked, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, vertLimit, jvl, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, horizLimit, jhl, 0);
}

// ---- BandedAligner (reference jni/BandedAlignerJNI.c:588-757; header jni/align2_BandedAlignerJNI.h:17-41) ----
static jint banded_jni(JNIEnv* env, jbyteArray query, jbyteArray ref, jint qstart, jint rstart, jint maxEdits, jboolean exact,
                       jint maxWidth, jintArray returnVals, int dir) {
    bbm_ctx* c = default_ctx();
    const jint rlen = (*env)->GetArrayLength(env, ref), qlen = (*env)->GetArrayLength(env, query);
    jbyte* jref = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, ref, nullptr);
    jbyte* jquery = (jbyte*)(*env)->GetPrimitiveArrayCritical(env, query, nullptr);
    jint* jrv = (jint*)(*env)->GetPrimitiveArrayCritical(env, returnVals, nullptr);
    bbm_band_task t; t.query_off = 0; t.ref_off = 0; t.query_len = qlen; t.ref_len = rlen; t.qstart = qstart; t.rstart = rstart;
    t.max_edits = maxEdits; t.max_width = maxWidth; t.exact = exact ? 1 : 0; t.dir = dir;
    bbm_band_out o = {};
    int rc = c ? bbm_banded_batch_host(c, jquery, qlen, jref, rlen, &t, &o, 1) : BBM_E_NODEVICE;
    if (rc == 0 && o.status != 0) rc = o.status;
    if (rc) fprintf(stderr, "libbbmapcuda: BandedAlignerJNI failed (%d): %s\n", rc, bbm_last_error());
    for (int k = 0; k < 5; ++k) jrv[k] = o.rv[k];
    (*env)->ReleasePrimitiveArrayCritical(env, ref, jref, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, query, jquery, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, returnVals, jrv, 0);
    return o.edits;
}
extern "C" JNIEXPORT jint JNICALL Java_align2_BandedAlignerJNI_alignForwardJNI(JNIEnv* env, jobject obj, jbyteArray query, jbyteArray ref,
    jint qstart, jint rstart, jint maxEdits, jboolean exact, jint maxWidth, jbyteArray baseToNumber, jintArray returnVals) {
    (void)obj; (void)baseToNumber; return banded_jni(env, query, ref, qstart, rstart, maxEdits, exact, maxWidth, returnVals, BBM_DIR_FORWARD);
}
extern "C" JNIEXPORT jint JNICALL Java_align2_BandedAlignerJNI_alignForwardRCJNI(JNIEnv* env, jobject obj, jbyteArray query, jbyteArray ref,
    jint qstart, jint rstart, jint maxEdits, jboolean exact, jint maxWidth, jbyteArray baseToNumber, jbyteArray baseToComplementExtended, jintArray returnVals) {
    (void)obj; (void)baseToNumber; (void)baseToComplementExtended; return banded_jni(env, query, ref, qstart, rstart, maxEdits, exact, maxWidth, returnVals, BBM_DIR_FORWARD_RC);
}
extern "C" JNIEXPORT jint JNICALL Java_align2_BandedAlignerJNI_alignReverseJNI(JNIEnv* env, jobject obj, jbyteArray query, jbyteArray ref,
    jint qstart, jint rstart, jint maxEdits, jboolean exact, jint maxWidth, jbyteArray baseToNumber, jintArray returnVals) {
    (void)obj; (void)baseToNumber; return banded_jni(env, query, ref, qstart, rstart, maxEdits, exact, maxWidth, returnVals, BBM_DIR_REVERSE);
}
extern "C" JNIEXPORT jint JNICALL Java_align2_BandedAlignerJNI_alignReverseRCJNI(JNIEnv* env, jobject obj, jbyteArray query, jbyteArray ref,
    jint qstart, jint rstart, jint maxEdits, jboolean exact, jint maxWidth, jbyteArray baseToNumber, jbyteArray baseToComplementExtended, jintArray returnVals) {
    (void)obj; (void)baseToNumber; (void)baseToComplementExtended; return banded_jni(env, query, ref, qstart, rstart, maxEdits, exact, maxWidth, returnVals, BBM_DIR_REVERSE_RC);
}
