/* See fake_jvm.h.  The class table below restates which constructors the reference's classes have:
 *   cvx.LinSolveException(DenseMatrix, DenseVector, DenseMatrix, String)   LinSolveException.scala:11-17 (no (String) ctor)
 *   cvx.UnsolvableSystemException(String), cvx.LineSearchFailedException(String), cvx.CvxbInfeasibleException(String)
 *   java.lang.AssertionError(Object), RuntimeException / IllegalArgumentException / UnsupportedOperationException(String) */
#include "fake_jvm.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct fake_class { const char* name; const char* ctors[3]; } fake_class;
static const fake_class g_classes[] = {
    {"cvx/LinSolveException", {"(Lbreeze/linalg/DenseMatrix;Lbreeze/linalg/DenseVector;Lbreeze/linalg/DenseMatrix;Ljava/lang/String;)V", 0, 0}},
    {"cvx/UnsolvableSystemException", {"(Ljava/lang/String;)V", 0, 0}},
    {"cvx/LineSearchFailedException", {"(Ljava/lang/String;)V", 0, 0}},
    {"cvx/CvxbInfeasibleException", {"(Ljava/lang/String;)V", 0, 0}},
    {"java/lang/AssertionError", {"(Ljava/lang/Object;)V", "()V", 0}},
    {"java/lang/RuntimeException", {"(Ljava/lang/String;)V", "()V", 0}},
    {"java/lang/IllegalArgumentException", {"(Ljava/lang/String;)V", "()V", 0}},
    {"java/lang/UnsupportedOperationException", {"(Ljava/lang/String;)V", "()V", 0}},
};
enum { NCLASSES = sizeof(g_classes) / sizeof(g_classes[0]) };

typedef struct fake_string { char text[1024]; } fake_string;
typedef struct fake_method { const fake_class* cls; const char* sig; } fake_method;
typedef struct fake_exception { const fake_class* cls; const char* sig; char msg[1024]; } fake_exception;

static char g_pending_class[128] = "", g_pending_ctor[256] = "", g_pending_msg[1024] = "";
static int g_outstanding = 0;
static fake_method g_methods[64];
static int g_nmethods = 0;

static void set_pending(const char* cls, const char* ctor, const char* msg) {
  snprintf(g_pending_class, sizeof g_pending_class, "%s", cls);
  snprintf(g_pending_ctor, sizeof g_pending_ctor, "%s", ctor);
  snprintf(g_pending_msg, sizeof g_pending_msg, "%s", msg ? msg : "");
}

static jclass f_FindClass(JNIEnv* env, const char* name) {
  (void)env;
  for (int i = 0; i < NCLASSES; ++i)
    if (!strcmp(g_classes[i].name, name)) return (jclass)&g_classes[i];
  set_pending("java/lang/NoClassDefFoundError", "", name);
  return 0;
}
static jmethodID f_GetMethodID(JNIEnv* env, jclass clazz, const char* name, const char* sig) {
  (void)env;
  const fake_class* c = (const fake_class*)clazz;
  if (c && !strcmp(name, "<init>"))
    for (int k = 0; k < 3 && c->ctors[k]; ++k)
      if (!strcmp(c->ctors[k], sig)) {
        fake_method* m = &g_methods[g_nmethods++ % 64];
        m->cls = c; m->sig = c->ctors[k];
        return (jmethodID)m;
      }
  set_pending("java/lang/NoSuchMethodError", sig, name);
  return 0;
}
static jstring f_NewStringUTF(JNIEnv* env, const char* utf) {
  (void)env;
  fake_string* s = (fake_string*)calloc(1, sizeof *s);
  snprintf(s->text, sizeof s->text, "%s", utf ? utf : "");
  return (jstring)s;
}
static jobject f_NewObject(JNIEnv* env, jclass clazz, jmethodID methodID, ...) {
  (void)env;
  const fake_method* m = (const fake_method*)methodID;
  if (!m || m->cls != (const fake_class*)clazz) { set_pending("java/lang/NoSuchMethodError", "", "NewObject"); return 0; }
  /* the message is the LAST argument of every constructor in the table; count the arguments from the signature */
  int nargs = 0;
  for (const char* p = m->sig + 1; *p && *p != ')'; ++p)
    if (*p == 'L') { ++nargs; while (*p && *p != ';') ++p; }
  va_list ap;
  va_start(ap, methodID);
  jobject last = 0;
  for (int i = 0; i < nargs; ++i) last = va_arg(ap, jobject);
  va_end(ap);
  fake_exception* e = (fake_exception*)calloc(1, sizeof *e);
  e->cls = m->cls; e->sig = m->sig;
  if (last) snprintf(e->msg, sizeof e->msg, "%s", ((fake_string*)last)->text);
  return (jobject)e;
}
static jint f_Throw(JNIEnv* env, jthrowable obj) {
  (void)env;
  fake_exception* e = (fake_exception*)obj;
  if (!e) return -1;
  set_pending(e->cls->name, e->sig, e->msg);
  return 0;
}
static jint f_ThrowNew(JNIEnv* env, jclass clazz, const char* msg) {
  (void)env;
  const fake_class* c = (const fake_class*)clazz;
  if (!c) return -1;
  for (int k = 0; k < 3 && c->ctors[k]; ++k)
    if (!strcmp(c->ctors[k], "(Ljava/lang/String;)V")) { set_pending(c->name, c->ctors[k], msg); return 0; }
  set_pending("java/lang/NoSuchMethodError", "(Ljava/lang/String;)V", c->name);     /* what a JVM does: an Error */
  return -1;
}
static jboolean f_ExceptionCheck(JNIEnv* env) { (void)env; return g_pending_class[0] != 0; }
static void f_ExceptionClear(JNIEnv* env) { (void)env; g_pending_class[0] = g_pending_ctor[0] = g_pending_msg[0] = 0; }
static jsize f_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return ((fake_array*)a)->len; }

static void* get_elems(jarray a, size_t esz) {
  fake_array* fa = (fake_array*)a;
  void* copy = malloc((size_t)(fa->len > 0 ? fa->len : 1) * esz);
  memcpy(copy, fa->data, (size_t)fa->len * esz);
  fa->outstanding++;
  g_outstanding++;
  return copy;
}
static void release_elems(jarray a, void* elems, jint mode, size_t esz) {
  fake_array* fa = (fake_array*)a;
  if (mode != JNI_ABORT) memcpy(fa->data, elems, (size_t)fa->len * esz);
  if (mode != JNI_COMMIT) { free(elems); fa->outstanding--; g_outstanding--; }
}
static jdouble* f_GetDoubleArrayElements(JNIEnv* env, jdoubleArray a, jboolean* isCopy) {
  (void)env;
  if (isCopy) *isCopy = 1;
  return (jdouble*)get_elems(a, sizeof(jdouble));
}
static void f_ReleaseDoubleArrayElements(JNIEnv* env, jdoubleArray a, jdouble* e, jint mode) { (void)env; release_elems(a, e, mode, sizeof(jdouble)); }
static jint* f_GetIntArrayElements(JNIEnv* env, jintArray a, jboolean* isCopy) {
  (void)env;
  if (isCopy) *isCopy = 1;
  return (jint*)get_elems(a, sizeof(jint));
}
static void f_ReleaseIntArrayElements(JNIEnv* env, jintArray a, jint* e, jint mode) { (void)env; release_elems(a, e, mode, sizeof(jint)); }
static void f_GetDoubleArrayRegion(JNIEnv* env, jdoubleArray a, jsize start, jsize len, jdouble* buf) {
  (void)env;
  memcpy(buf, (double*)((fake_array*)a)->data + start, (size_t)len * sizeof(double));
}
static void f_SetDoubleArrayRegion(JNIEnv* env, jdoubleArray a, jsize start, jsize len, const jdouble* buf) {
  (void)env;
  memcpy((double*)((fake_array*)a)->data + start, buf, (size_t)len * sizeof(double));
}
static void f_SetIntArrayRegion(JNIEnv* env, jintArray a, jsize start, jsize len, const jint* buf) {
  (void)env;
  memcpy((int*)((fake_array*)a)->data + start, buf, (size_t)len * sizeof(int));
}

static const struct JNINativeInterface_ g_table = {
    f_FindClass, f_GetMethodID, f_NewObject, f_NewStringUTF, f_Throw, f_ThrowNew, f_ExceptionCheck, f_ExceptionClear,
    f_GetArrayLength, f_GetDoubleArrayElements, f_ReleaseDoubleArrayElements, f_GetIntArrayElements,
    f_ReleaseIntArrayElements, f_GetDoubleArrayRegion, f_SetDoubleArrayRegion, f_SetIntArrayRegion};
static JNIEnv g_env = &g_table;

JNIEnv* fake_env(void) { return &g_env; }
static jarray new_array(int is_int, jsize len, const void* init) {
  fake_array* a = (fake_array*)calloc(1, sizeof *a);
  size_t esz = is_int ? sizeof(int) : sizeof(double);
  a->is_int = is_int; a->len = len; a->data = calloc((size_t)(len > 0 ? len : 1), esz);
  if (init) memcpy(a->data, init, (size_t)len * esz);
  return (jarray)a;
}
jdoubleArray fake_new_double_array(jsize len, const double* init) { return new_array(0, len, init); }
jintArray fake_new_int_array(jsize len, const int* init) { return new_array(1, len, init); }
double* fake_doubles(jdoubleArray a) { return (double*)((fake_array*)a)->data; }
int* fake_ints(jintArray a) { return (int*)((fake_array*)a)->data; }
void fake_free_array(jarray a) { if (a) { free(((fake_array*)a)->data); free(a); } }
const char* fake_pending_class(void) { return g_pending_class; }
const char* fake_pending_ctor(void) { return g_pending_ctor; }
const char* fake_pending_message(void) { return g_pending_msg; }
void fake_clear_pending(void) { g_pending_class[0] = g_pending_ctor[0] = g_pending_msg[0] = 0; }
int fake_outstanding_arrays(void) { return g_outstanding; }
