// jni_utils.cc -- libgkl_utils.so: the six symbols GKL's unchanged Java class com.intel.gkl.IntelGKLUtils binds
// (IntelGKLUtils.java:109-114), replacing utils/utils.cc:42-115.
//
// IntelPairHmm.load(), IntelPDHMM.load() and IntelSmithWaterman.load() refuse to go on unless this library loads and
// isAvxSupported() (isAvx2Supported() for Smith-Waterman) answers true (IntelPairHmm.java:66-77,
// IntelSmithWaterman.java:79-88).  On an x86 host GKL's own libgkl_utils.so can stay; this stand-alone replacement
// (plain C++, no CUDA, no OpenMP) exists so that the drop-in also works where GKL's x86-only library does not build
// -- e.g. the Arm host of a Grace-Blackwell node.  The "AVX" questions are answered for what they gate: on x86 by
// CPUID like the reference (common/avx.h:69-132), elsewhere with true, because the kernels behind the gate run on the
// GPU and the loader of each kernel library still refuses to load without one (JNI_OnLoad).
//   get/setFlushToZero: MXCSR.FTZ on x86 (utils.cc:42-66), FPCR.FZ on AArch64.
//   getAvailableOmpThreads: the host threads the process may run on (the reference reports omp_get_max_threads()).
#include <stdint.h>

#if defined(__x86_64__) || defined(__i386__)
#include <cpuid.h>
#include <xmmintrin.h>
#endif
#ifdef __linux__
#include <sched.h>
#endif
#include <thread>

#include "jni_min.h"

namespace {

#if defined(__x86_64__) || defined(__i386__)
uint64_t xgetbv0() {
  uint32_t lo, hi;
  __asm__ volatile("xgetbv" : "=a"(lo), "=d"(hi) : "c"(0));
  return ((uint64_t)hi << 32) | lo;
}
// common/avx.h:69-132: CPUID feature bit and OS support for the register state (XCR0)
bool cpu_has(int level) {  // 1 AVX, 2 AVX2, 3 AVX-512 F+DQ+VL+BW
  unsigned a, b, c, d;
  if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
  const bool osxsave = (c >> 27) & 1, avx = (c >> 28) & 1;
  if (!osxsave || !avx) return false;
  const uint64_t xcr0 = xgetbv0();
  if ((xcr0 & 0x6) != 0x6) return false;
  if (level == 1) return true;
  if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
  if (level == 2) return (b >> 5) & 1;
  const bool f = (b >> 16) & 1, dq = (b >> 17) & 1, bw = (b >> 30) & 1, vl = (b >> 31) & 1;
  return f && dq && bw && vl && (xcr0 & 0xe0) == 0xe0;
}
#endif

}  // namespace

extern "C" {

JNIEXPORT jboolean JNICALL Java_com_intel_gkl_IntelGKLUtils_getFlushToZeroNative(JNIEnv*, jobject) {
#if defined(__x86_64__) || defined(__i386__)
  return _MM_GET_FLUSH_ZERO_MODE() == _MM_FLUSH_ZERO_ON ? 1 : 0;
#elif defined(__aarch64__)
  uint64_t fpcr;
  __asm__ volatile("mrs %0, fpcr" : "=r"(fpcr));
  return (fpcr >> 24) & 1;
#else
  return 0;
#endif
}

JNIEXPORT void JNICALL Java_com_intel_gkl_IntelGKLUtils_setFlushToZeroNative(JNIEnv*, jobject, jboolean value) {
#if defined(__x86_64__) || defined(__i386__)
  _MM_SET_FLUSH_ZERO_MODE(value ? _MM_FLUSH_ZERO_ON : _MM_FLUSH_ZERO_OFF);
#elif defined(__aarch64__)
  uint64_t fpcr;
  __asm__ volatile("mrs %0, fpcr" : "=r"(fpcr));
  fpcr = value ? (fpcr | (1ull << 24)) : (fpcr & ~(1ull << 24));
  __asm__ volatile("msr fpcr, %0" ::"r"(fpcr));
#else
  (void)value;
#endif
}

JNIEXPORT jboolean JNICALL Java_com_intel_gkl_IntelGKLUtils_isAvxSupportedNative(JNIEnv*, jobject) {
#if defined(__x86_64__) || defined(__i386__)
  return cpu_has(1) ? 1 : 0;
#else
  return 1;
#endif
}

JNIEXPORT jboolean JNICALL Java_com_intel_gkl_IntelGKLUtils_isAvx2SupportedNative(JNIEnv*, jobject) {
#if defined(__x86_64__) || defined(__i386__)
  return cpu_has(2) ? 1 : 0;
#else
  return 1;
#endif
}

JNIEXPORT jboolean JNICALL Java_com_intel_gkl_IntelGKLUtils_isAvx512SupportedNative(JNIEnv*, jobject) {
#if defined(__x86_64__) || defined(__i386__)
  return cpu_has(3) ? 1 : 0;
#else
  return 0;  // only logged ("Using CPU-supported AVX-512 instructions", IntelPairHmm.java:92-94)
#endif
}

JNIEXPORT jint JNICALL Java_com_intel_gkl_IntelGKLUtils_getAvailableOmpThreadsNative(JNIEnv*, jobject) {
#ifdef __linux__
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) return (jint)CPU_COUNT(&set);
#endif
  const unsigned n = std::thread::hardware_concurrency();
  return (jint)(n ? n : 1);
}

}  // extern "C"
