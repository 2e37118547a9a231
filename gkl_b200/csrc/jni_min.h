// jni_min.h -- the part of the Java Native Interface this library needs, for building without a JDK.
//
// This image has no JDK (no jni.h anywhere).  When a JDK is present, build with -DGKLB_USE_SYSTEM_JNI
// -I$JAVA_HOME/include -I$JAVA_HOME/include/linux and this file only forwards to <jni.h>.
// Otherwise it declares the JNI primitive types and a JNIEnv whose member functions dispatch through
// the interface function table by slot number.  The slot numbers are those of the "Interface Function
// Table" in the JNI specification (Java SE, chapter 4), which is ABI: every JVM lays the table out this
// way, so code compiled against this header calls the same entries as code compiled against jni.h.
#pragma once

#ifdef GKLB_USE_SYSTEM_JNI
#include <jni.h>
#else
#include <stdarg.h>
#include <stdint.h>

extern "C" {
typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

class _jobject {};
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jbyteArray;
typedef jarray jdoubleArray;
typedef jarray jlongArray;
struct _jfieldID;
typedef _jfieldID* jfieldID;

#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_ERR (-1)
#define JNI_ABORT 2
#define JNI_VERSION_1_6 0x00010006
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

struct JNINativeInterface_ {
  void* fn[240];
};
struct JavaVM_;
typedef JavaVM_ JavaVM;
}

// JNI specification slot numbers used below
enum {
  kJniFindClass = 6, kJniThrowNew = 14, kJniExceptionClear = 17, kJniPushLocalFrame = 19, kJniPopLocalFrame = 20,
  kJniDeleteLocalRef = 23, kJniGetFieldID = 94, kJniGetObjectField = 95, kJniGetArrayLength = 171,
  kJniGetObjectArrayElement = 173, kJniNewDoubleArray = 182, kJniGetByteArrayElements = 184, kJniGetDoubleArrayElements = 190,
  kJniReleaseByteArrayElements = 192, kJniReleaseDoubleArrayElements = 198, kJniGetByteArrayRegion = 200,
  kJniSetDoubleArrayRegion = 214, kJniGetPrimitiveArrayCritical = 222, kJniReleasePrimitiveArrayCritical = 223,
  kJniExceptionCheck = 228
};

struct JNIEnv_ {
  const JNINativeInterface_* functions;
  template <class F> F slot(int i) { return reinterpret_cast<F>(functions->fn[i]); }
  jclass FindClass(const char* name) { return slot<jclass (*)(JNIEnv_*, const char*)>(kJniFindClass)(this, name); }
  jint ThrowNew(jclass c, const char* msg) { return slot<jint (*)(JNIEnv_*, jclass, const char*)>(kJniThrowNew)(this, c, msg); }
  void ExceptionClear() { slot<void (*)(JNIEnv_*)>(kJniExceptionClear)(this); }
  jboolean ExceptionCheck() { return slot<jboolean (*)(JNIEnv_*)>(kJniExceptionCheck)(this); }
  jint PushLocalFrame(jint cap) { return slot<jint (*)(JNIEnv_*, jint)>(kJniPushLocalFrame)(this, cap); }
  jobject PopLocalFrame(jobject r) { return slot<jobject (*)(JNIEnv_*, jobject)>(kJniPopLocalFrame)(this, r); }
  void DeleteLocalRef(jobject o) { slot<void (*)(JNIEnv_*, jobject)>(kJniDeleteLocalRef)(this, o); }
  jfieldID GetFieldID(jclass c, const char* name, const char* sig) {
    return slot<jfieldID (*)(JNIEnv_*, jclass, const char*, const char*)>(kJniGetFieldID)(this, c, name, sig);
  }
  jobject GetObjectField(jobject o, jfieldID f) { return slot<jobject (*)(JNIEnv_*, jobject, jfieldID)>(kJniGetObjectField)(this, o, f); }
  jsize GetArrayLength(jarray a) { return slot<jsize (*)(JNIEnv_*, jarray)>(kJniGetArrayLength)(this, a); }
  jobject GetObjectArrayElement(jobjectArray a, jsize i) {
    return slot<jobject (*)(JNIEnv_*, jobjectArray, jsize)>(kJniGetObjectArrayElement)(this, a, i);
  }
  jbyte* GetByteArrayElements(jbyteArray a, jboolean* is_copy) {
    return slot<jbyte* (*)(JNIEnv_*, jbyteArray, jboolean*)>(kJniGetByteArrayElements)(this, a, is_copy);
  }
  void ReleaseByteArrayElements(jbyteArray a, jbyte* p, jint mode) {
    slot<void (*)(JNIEnv_*, jbyteArray, jbyte*, jint)>(kJniReleaseByteArrayElements)(this, a, p, mode);
  }
  jdouble* GetDoubleArrayElements(jdoubleArray a, jboolean* is_copy) {
    return slot<jdouble* (*)(JNIEnv_*, jdoubleArray, jboolean*)>(kJniGetDoubleArrayElements)(this, a, is_copy);
  }
  void ReleaseDoubleArrayElements(jdoubleArray a, jdouble* p, jint mode) {
    slot<void (*)(JNIEnv_*, jdoubleArray, jdouble*, jint)>(kJniReleaseDoubleArrayElements)(this, a, p, mode);
  }
  jdoubleArray NewDoubleArray(jsize n) { return slot<jdoubleArray (*)(JNIEnv_*, jsize)>(kJniNewDoubleArray)(this, n); }
  void SetDoubleArrayRegion(jdoubleArray a, jsize start, jsize len, const jdouble* buf) {
    slot<void (*)(JNIEnv_*, jdoubleArray, jsize, jsize, const jdouble*)>(kJniSetDoubleArrayRegion)(this, a, start, len, buf);
  }
  void* GetPrimitiveArrayCritical(jarray a, jboolean* is_copy) {
    return slot<void* (*)(JNIEnv_*, jarray, jboolean*)>(kJniGetPrimitiveArrayCritical)(this, a, is_copy);
  }
  void ReleasePrimitiveArrayCritical(jarray a, void* p, jint mode) {
    slot<void (*)(JNIEnv_*, jarray, void*, jint)>(kJniReleasePrimitiveArrayCritical)(this, a, p, mode);
  }
  void GetByteArrayRegion(jbyteArray a, jsize start, jsize len, jbyte* buf) {
    slot<void (*)(JNIEnv_*, jbyteArray, jsize, jsize, jbyte*)>(kJniGetByteArrayRegion)(this, a, start, len, buf);
  }
};
typedef JNIEnv_ JNIEnv;
#endif
