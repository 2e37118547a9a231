// jni_sw.cc -- the JNI face of libgkl_smithwaterman.so: the three symbols GKL's unchanged Java class
// com.intel.gkl.smithwaterman.IntelSmithWaterman binds (IntelSmithWaterman.java:183-186).
//
//   Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_initNative   replaces smithwaterman/IntelSmithWaterman.cc:47-66
//   Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_alignNative  replaces smithwaterman/IntelSmithWaterman.cc:72-121
//   Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_doneNative   replaces smithwaterman/IntelSmithWaterman.cc:128-130
//
// alignNative keeps the reference's conventions: the three arrays are taken with GetPrimitiveArrayCritical and
// released with mode 0, a NULL from any of them throws IllegalArgumentException("Arrays aren't valid.") and returns
// -1, an allocation failure throws OutOfMemoryError("Memory allocation issue") and returns -1, otherwise the
// alignment offset is returned and the CIGAR sits in the caller's byte[] (zero padded; the Java side trims it).
// The device call is made on private copies: a JNI critical section must not block on other JVM work, and the
// CUDA call may.
#include <string.h>

#include <vector>

#include "../../include/gklb_sw.h"
#include "jni_min.h"

namespace {

void sw_throw(JNIEnv* env, const char* cls, const char* msg) {
  if (env->ExceptionCheck()) env->ExceptionClear();
  jclass c = env->FindClass(cls);
  if (c) env->ThrowNew(c, msg);
}

}  // namespace

extern "C" {

JNIEXPORT void JNICALL Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_initNative(JNIEnv* env, jclass cls) {
  (void)cls;
  const int rc = gklb_sw_init();
  // initNative returns void and load() has already answered true (IntelSmithWaterman.java:77-111): a missing device
  // surfaces as an exception here, like an UnsatisfiedLinkError would from JNI_OnLoad
  if (rc) sw_throw(env, "java/lang/RuntimeException", gklb_last_error());
}

JNIEXPORT jint JNICALL Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_alignNative(
    JNIEnv* env, jclass cls, jbyteArray ref, jbyteArray alt, jbyteArray cigar, jint match, jint mismatch, jint open,
    jint extend, jbyte strategy) {
  (void)cls;
  const jint ref_len = env->GetArrayLength(ref), alt_len = env->GetArrayLength(alt), cigar_len = env->GetArrayLength(cigar);
  std::vector<uint8_t> s1((size_t)(ref_len > 0 ? ref_len : 0)), s2((size_t)(alt_len > 0 ? alt_len : 0));
  {
    void* r = env->GetPrimitiveArrayCritical(ref, nullptr);
    void* a = env->GetPrimitiveArrayCritical(alt, nullptr);
    if (!r || !a) {
      if (r) env->ReleasePrimitiveArrayCritical(ref, r, 0);
      if (a) env->ReleasePrimitiveArrayCritical(alt, a, 0);
      sw_throw(env, "java/lang/IllegalArgumentException", "Arrays aren't valid.");
      return -1;
    }
    if (ref_len > 0) memcpy(s1.data(), r, (size_t)ref_len);
    if (alt_len > 0) memcpy(s2.data(), a, (size_t)alt_len);
    env->ReleasePrimitiveArrayCritical(alt, a, 0);
    env->ReleasePrimitiveArrayCritical(ref, r, 0);
  }
  std::vector<char> out((size_t)(cigar_len > 0 ? cigar_len : 1), 0);
  uint32_t count = 0;
  int32_t offset = 0;
  const int rc = gklb_sw_align(match, mismatch, open, extend, s1.data(), s2.data(), ref_len, alt_len, (int32_t)strategy,
                               out.data(), cigar_len, &count, &offset);
  if (rc == GKLB_ERR_OOM) {
    sw_throw(env, "java/lang/OutOfMemoryError", "Memory allocation issue");
    return -1;
  }
  if (rc == GKLB_ERR_INVALID) {
    sw_throw(env, "java/lang/IllegalArgumentException", gklb_last_error());
    return -1;
  }
  if (rc) {
    sw_throw(env, "java/lang/RuntimeException", gklb_last_error());
    return -1;
  }
  void* c = env->GetPrimitiveArrayCritical(cigar, nullptr);
  if (!c) {
    sw_throw(env, "java/lang/IllegalArgumentException", "Arrays aren't valid.");
    return -1;
  }
  memcpy(c, out.data(), (size_t)count);   // runSWOnePairBT writes the string and leaves the rest of the array alone
  env->ReleasePrimitiveArrayCritical(cigar, c, 0);
  return offset;
}

JNIEXPORT void JNICALL Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_doneNative(JNIEnv* env, jclass cls) {
  (void)env; (void)cls;
  gklb_sw_done();
}

}  // extern "C"
