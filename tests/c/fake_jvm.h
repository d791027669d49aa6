/* A JNIEnv over plain C memory: just enough of a JVM to EXECUTE jni/cvxb_jni.c in the GPU test run (no JDK in the
 * image).  TEST INFRASTRUCTURE.  Arrays are heap blocks handed out as COPIES by Get<Type>ArrayElements and written
 * back on release unless the mode is JNI_ABORT -- the strictest behaviour the specification allows, so a shim that
 * forgets a release or releases with the wrong mode fails the tests.  Classes are looked up in a table that mirrors
 * the reference's exception classes and their real constructors. */
#ifndef CVXB_FAKE_JVM_H
#define CVXB_FAKE_JVM_H
#include <jni.h>

typedef struct fake_array { int is_int; jsize len; void* data; int outstanding; } fake_array;

JNIEnv* fake_env(void);
jdoubleArray fake_new_double_array(jsize len, const double* init);
jintArray fake_new_int_array(jsize len, const int* init);
double* fake_doubles(jdoubleArray a);      /* the array's own storage */
int* fake_ints(jintArray a);
void fake_free_array(jarray a);
/* pending exception: class name ("" when none), constructor signature used, message */
const char* fake_pending_class(void);
const char* fake_pending_ctor(void);
const char* fake_pending_message(void);
void fake_clear_pending(void);
int fake_outstanding_arrays(void);         /* Get...Elements without a matching Release */
#endif
