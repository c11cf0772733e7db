/* TEST INFRASTRUCTURE ONLY.  Stand-in for <jni.h> so that the reference's own
 * jni/*.c files compile unmodified from /root/reference (no JDK in this image).
 * Forwards to the repo's minimal JNI ABI header. */
#ifndef ORACLE_JNI_STUB_H
#define ORACLE_JNI_STUB_H
#include "../../include/bbm_jni_min.h"
#endif
