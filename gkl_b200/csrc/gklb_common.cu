// gklb_common.cu -- what every library of this package carries: the calling thread's last error message, the device
// probe, and JNI_OnLoad.  libgkl_pairhmm.so holds everything; libgkl_pdhmm.so and libgkl_smithwaterman.so are linked
// from this file plus their own engine and JNI translation units (GKL's loader only accepts its fixed library
// names, NativeLibraryLoader.java:45).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>

#include "../../include/gklb_pairhmm.h"
#include "jni_min.h"

namespace gklb {

namespace {
thread_local std::string t_last_error;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}
const std::string& last_error_string() { return t_last_error; }
void set_last_error(const std::string& s) { t_last_error = s; }

}  // namespace gklb

// used by pdhmm_engine.cu / sw_engine.cu
int gklb_internal_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  gklb::set_last_error(buf);
  return code;
}

extern "C" {

const char* gklb_last_error(void) { return gklb::last_error_string().c_str(); }

const char* gklb_version(void) { return "gkl_b200 0.2 (sm_100a)"; }

int gklb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int i = 0; i < n; i++) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
}

// System.load() fails (UnsatisfiedLinkError -> NativeLibraryLoader.load returns false -> the Java shim's load()
// returns false -> GATK falls back to its own Java implementation) when this machine has no sm_100 GPU: that is the
// reference's own failover path (NativeLibraryLoader.java:114-133, IntelPairHmm.java:66-82); there is no CPU
// implementation inside these libraries.
JNIEXPORT jint JNICALL JNI_OnLoad(JavaVM* vm, void* reserved) {
  (void)vm;
  (void)reserved;
  return gklb_device_count() > 0 ? JNI_VERSION_1_6 : JNI_ERR;
}

}  // extern "C"
