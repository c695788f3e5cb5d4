/* Stand-in for <jni.h> — TEST INFRASTRUCTURE.  The build image has no JDK, so the JNI glue
 * (integration/java/jni/lps_b200_jni.c) cannot be compiled against the real header here.  This header
 * declares the JNI types and the JNIEnv function table entries the glue uses, with the real names and
 * signatures (JNI specification, "Array Operations": Get/ReleasePrimitiveArrayCritical; "String Operations":
 * NewStringUTF), so the glue compiles unmodified; jni_call_order.c implements the table over plain C arrays and
 * calls the Java_lpsolver_LPStateNative_* functions in the order the Java class would. */
#ifndef JNI_MOCK_H
#define JNI_MOCK_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef uint8_t jboolean;

/* a "Java array": length + storage + how many critical borrows are outstanding */
struct mock_array {
  jint length;
  void *data;
  int borrowed;
  int writes_committed;
};
typedef struct mock_array *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;
typedef const char *jstring; /* NewStringUTF hands back a C string in the mock */

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  void *(*GetPrimitiveArrayCritical)(JNIEnv *env, jarray array, jboolean *isCopy);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv *env, jarray array, void *carray, jint mode);
  jstring (*NewStringUTF)(JNIEnv *env, const char *utf);
};
#endif
