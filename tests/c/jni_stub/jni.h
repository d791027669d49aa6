/* Minimal stand-in for the JDK's <jni.h>: ONLY the types and the JNIEnv entries jni/cvxb_jni.c uses, with the
 * signatures of the JNI specification (Java SE 8, chapter 4 "JNI Functions").  TEST INFRASTRUCTURE: the build image
 * has no JDK; this header lets the CPU test run compile the shim with -Wall -Werror, and tests/c/fake_jvm.c implements
 * the same entries over plain C arrays so the GPU test run can execute it.  In a real build the JDK's jni.h is used
 * instead (the member ORDER of the function table below is not the JVM's; nothing here may be linked against a JVM). */
#ifndef CVXB_TEST_JNI_STUB_H
#define CVXB_TEST_JNI_STUB_H

#include <stdarg.h>
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef unsigned char jboolean;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jdoubleArray;
typedef jarray jintArray;
struct _jmethodID;
typedef struct _jmethodID* jmethodID;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jmethodID (*GetMethodID)(JNIEnv* env, jclass clazz, const char* name, const char* sig);
  jobject (*NewObject)(JNIEnv* env, jclass clazz, jmethodID methodID, ...);
  jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
  jint (*Throw)(JNIEnv* env, jthrowable obj);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jboolean (*ExceptionCheck)(JNIEnv* env);
  void (*ExceptionClear)(JNIEnv* env);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  jdouble* (*GetDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jboolean* isCopy);
  void (*ReleaseDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jdouble* elems, jint mode);
  jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
  void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
  void (*GetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, jdouble* buf);
  void (*SetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, const jdouble* buf);
  void (*SetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, const jint* buf);
};

#endif
