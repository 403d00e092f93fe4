/* Minimal stand-in for <jni.h>: only what java/jni/ldagpu_jni.c uses.  There is no JDK in the build image; this lets
 * tests/test_host_logic.py type-check the JNI stub against include/ldagpu.h with gcc -fsyntax-only. */
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef double jdouble; typedef unsigned char jboolean; typedef int jsize;
typedef void* jobject; typedef jobject jclass; typedef jobject jarray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jdoubleArray; typedef jobject jthrowable;
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_; typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ { jclass (*FindClass)(JNIEnv*, const char*); jint (*ThrowNew)(JNIEnv*, jclass, const char*);
 jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*); jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*); jdouble* (*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*);
 void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint); void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint); void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
 jsize (*GetArrayLength)(JNIEnv*, jarray); };
