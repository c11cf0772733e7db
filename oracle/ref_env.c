/*
 * ref_env.c — TEST INFRASTRUCTURE ONLY.
 * A fake JNIEnv (arrays are {data,len} records) so that the reference's exported
 * JNI entry points (reference jni/MultiStateAligner11tsJNI.c:707-812,
 * jni/BandedAlignerJNI.c:588-757) can be driven from C / ctypes without a JVM.
 * Linked together with the reference's own, unmodified C files into
 * oracle/_ref/libbbref.so (see oracle/Makefile).  The same driver is also used
 * by tests against the product library's Java_align2_* twins.
 */
#include <stdlib.h>
#include <string.h>
#include "../include/bbm_jni_min.h"

typedef struct { void* data; jsize len; int pins; } fake_array;

static jsize fe_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return ((fake_array*)a)->len; }
static void* fe_GetCritical(JNIEnv* env, jarray a, jboolean* isCopy) { (void)env; if (isCopy) *isCopy = 0; ((fake_array*)a)->pins++; return ((fake_array*)a)->data; }
static void fe_ReleaseCritical(JNIEnv* env, jarray a, void* p, jint mode) { (void)env; (void)p; (void)mode; ((fake_array*)a)->pins--; }

static struct JNINativeInterface_ g_table;
static const struct JNINativeInterface_* g_env_ptr = NULL;

JNIEnv* fake_jni_env(void) {
    if (!g_env_ptr) {
        memset(&g_table, 0, sizeof(g_table));
        g_table.GetArrayLength = fe_GetArrayLength;
        g_table.GetPrimitiveArrayCritical = fe_GetCritical;
        g_table.ReleasePrimitiveArrayCritical = fe_ReleaseCritical;
        g_env_ptr = &g_table;
    }
    return (JNIEnv*)&g_env_ptr;
}

typedef void (*fillU_jni_fn)(JNIEnv*, jobject, jbyteArray, jbyteArray, jint, jint, jintArray, jlongArray, jintArray, jintArray, jintArray, jint, jint);
typedef void (*fillL_jni_fn)(JNIEnv*, jobject, jbyteArray, jbyteArray, jint, jint, jint, jintArray, jlongArray, jintArray, jintArray, jintArray,
                             jint, jint, jint, jfloat, jintArray, jintArray, jbyteArray, jintArray);
typedef jint (*band_jni_fn)(JNIEnv*, jobject, jbyteArray, jbyteArray, jint, jint, jint, jboolean, jint, jbyteArray, jintArray);
typedef jint (*bandrc_jni_fn)(JNIEnv*, jobject, jbyteArray, jbyteArray, jint, jint, jint, jboolean, jint, jbyteArray, jbyteArray, jintArray);

#define FA(name, ptr, n) fake_array name = { (void*)(ptr), (jsize)(n), 0 }

/* Drive any implementation of the fillUnlimitedJNI entry point (fn = its address). */
int fake_call_fillUnlimitedJNI(void* fn, jbyte* read, jint rlen, jbyte* ref, jint reflen, jint refStartLoc, jint refEndLoc,
                               jint* result4, jlong* iterations1, jint* packed, jint packed_len, jint* sub, jint* ins, jint tabLen,
                               jint maxRows, jint maxColumns) {
    FA(aread, read, rlen); FA(aref, ref, reflen); FA(ares, result4, 4); FA(ait, iterations1, 1);
    FA(apk, packed, packed_len); FA(asub, sub, tabLen); FA(ains, ins, tabLen);
    ((fillU_jni_fn)fn)(fake_jni_env(), NULL, &aread, &aref, refStartLoc, refEndLoc, &ares, &ait, &apk, &asub, &ains, maxRows, maxColumns);
    return aread.pins | aref.pins | ares.pins | ait.pins | apk.pins | asub.pins | ains.pins; /* 0 = every pin released */
}

int fake_call_fillLimitedXJNI(void* fn, jbyte* read, jint rlen, jbyte* ref, jint reflen, jint refStartLoc, jint refEndLoc, jint minScore,
                              jint* result5, jlong* iterations1, jint* packed, jint packed_len, jint* sub, jint* ins, jint tabLen,
                              jint maxRows, jint maxColumns, jint bandwidth, jfloat bandwidthRatio,
                              jint* vertLimit, jint* horizLimit, jbyte* baseToNumber, jint* insC) {
    FA(aread, read, rlen); FA(aref, ref, reflen); FA(ares, result5, 5); FA(ait, iterations1, 1);
    FA(apk, packed, packed_len); FA(asub, sub, tabLen); FA(ains, ins, tabLen);
    FA(avl, vertLimit, maxRows + 1); FA(ahl, horizLimit, maxColumns + 1); FA(ab2n, baseToNumber, 128); FA(ainsc, insC, tabLen);
    ((fillL_jni_fn)fn)(fake_jni_env(), NULL, &aread, &aref, refStartLoc, refEndLoc, minScore, &ares, &ait, &apk, &asub, &ains,
                       maxRows, maxColumns, bandwidth, bandwidthRatio, &avl, &ahl, &ab2n, &ainsc);
    return aread.pins | aref.pins | ares.pins | ait.pins | apk.pins | asub.pins | ains.pins | avl.pins | ahl.pins | ab2n.pins | ainsc.pins;
}

jint fake_call_bandedJNI(void* fn, jbyte* query, jint qlen, jbyte* ref, jint reflen, jint qstart, jint rstart, jint maxEdits,
                         jboolean exact, jint maxWidth, jbyte* baseToNumber, jint* returnVals5) {
    FA(aq, query, qlen); FA(ar, ref, reflen); FA(ab, baseToNumber, 128); FA(arv, returnVals5, 5);
    return ((band_jni_fn)fn)(fake_jni_env(), NULL, &aq, &ar, qstart, rstart, maxEdits, exact, maxWidth, &ab, &arv);
}

jint fake_call_bandedRCJNI(void* fn, jbyte* query, jint qlen, jbyte* ref, jint reflen, jint qstart, jint rstart, jint maxEdits,
                           jboolean exact, jint maxWidth, jbyte* baseToNumber, jbyte* baseToComplementExtended, jint* returnVals5) {
    FA(aq, query, qlen); FA(ar, ref, reflen); FA(ab, baseToNumber, 128); FA(ac, baseToComplementExtended, 128); FA(arv, returnVals5, 5);
    return ((bandrc_jni_fn)fn)(fake_jni_env(), NULL, &aq, &ar, qstart, rstart, maxEdits, exact, maxWidth, &ab, &ac, &arv);
}
