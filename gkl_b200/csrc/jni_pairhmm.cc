// jni_pairhmm.cc -- the JNI face of libgkl_pairhmm.so: exactly the three symbols GKL's unchanged Java
// class com.intel.gkl.pairhmm.IntelPairHmm binds (IntelPairHmm.java:157-164), plus JNI_OnLoad.
//
//   Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative                 replaces pairhmm/IntelPairHmm.cc:55-118
//   Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative   replaces pairhmm/IntelPairHmm.cc:125-181
//   Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative                 replaces pairhmm/IntelPairHmm.cc:189-192
//
// The layer is thin: it caches the six field IDs (pairhmm/JavaData.h:55-62), copies the Java byte[]s of one
// call into the flat arenas of gklb_pairhmm_batch with GetByteArrayRegion (one copy, no pin/release
// bookkeeping, local references deleted as it goes -- GKL's JavaData pins 5R+H arrays and leaks the local
// references, pairhmm/JavaData.h:135-145), pins the output double[] the way GKL does (:147-154) and calls the
// C-ABI.  Errors become Java exceptions by class path like GKL's (IntelPairHmm.cc:64-68,141-145,171-178).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/gklb_pairhmm.h"
#include "jni_min.h"

namespace {

struct FieldIds {
  jfieldID read_bases = nullptr, read_quals = nullptr, ins_gop = nullptr, del_gop = nullptr, gcp = nullptr,
           hap_bases = nullptr;
};
FieldIds g_fid;  // process-global like GKL's static jfieldIDs (JavaData.h:163-176); written by initNative only
std::mutex g_fid_mu;

void throw_java(JNIEnv* env, const char* class_path, const char* msg) {
  env->ExceptionClear();
  jclass c = env->FindClass(class_path);
  if (c) env->ThrowNew(c, msg);
}

void throw_status(JNIEnv* env, int rc) {
  const char* msg = gklb_last_error();
  switch (rc) {
    case GKLB_ERR_OOM: throw_java(env, "java/lang/OutOfMemoryError", msg); break;
    case GKLB_ERR_INVALID: throw_java(env, "java/lang/IllegalArgumentException", msg); break;
    default: throw_java(env, "java/lang/RuntimeException", msg); break;
  }
}

struct Arena {
  std::vector<uint8_t> bytes;
  std::vector<int64_t> off{0};
};

// Appends object[index].field (a byte[]) to the arena.  expect_len >= 0 demands that length
// (GKL assumes the four quality arrays are as long as readBases without checking; reading past a
// shorter one would be undefined behaviour there, here it is an IllegalArgumentException).
// Returns the length, or -1 after throwing.
int append_field(JNIEnv* env, jobjectArray array, int index, jfieldID fid, Arena& a, int expect_len, bool advance) {
  jobject obj = env->GetObjectArrayElement(array, index);
  if (!obj) { throw_java(env, "java/lang/NullPointerException", "null element in input array"); return -1; }
  jbyteArray bytes = (jbyteArray)env->GetObjectField(obj, fid);
  if (!bytes) {
    env->DeleteLocalRef(obj);
    throw_java(env, "java/lang/NullPointerException", "null byte[] in data holder");
    return -1;
  }
  const int len = env->GetArrayLength(bytes);
  if (expect_len >= 0 && len != expect_len) {
    env->DeleteLocalRef(bytes);
    env->DeleteLocalRef(obj);
    throw_java(env, "java/lang/IllegalArgumentException", "per-read arrays differ in length");
    return -1;
  }
  const size_t at = a.bytes.size();
  a.bytes.resize(at + (size_t)len);
  if (len > 0) env->GetByteArrayRegion(bytes, 0, len, reinterpret_cast<jbyte*>(a.bytes.data() + at));
  if (advance) a.off.push_back((int64_t)a.bytes.size());
  env->DeleteLocalRef(bytes);
  env->DeleteLocalRef(obj);
  return len;
}

// Large calls are pipelined: reads are marshalled in blocks, and while block k+1 is being copied out of the Java
// heap, block k is already on the GPU (two engines borrowed from the pool, on the same device, take turns).  In a
// JVM the 5R+H array accesses of this binding cost as much as the kernels once those run at TCUPS rates
// (SURVEY.md 8(f) N1).  The engines come from the global pool (gklb_pairhmm_acquire_engine), so the pipeline runs
// on whatever device the pool was configured with and concurrent Java threads each get their own pair.
int pipeline_block_reads() {
  const char* v = getenv("GKLB_JNI_BLOCK_READS");
  const int n = v ? atoi(v) : 2048;
  return n > 0 ? n : 2048;
}

}  // namespace

extern "C" {

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative(JNIEnv* env, jclass cls,
                                                                            jclass readDataHolder,
                                                                            jclass haplotypeDataHolder,
                                                                            jboolean use_double, jint max_threads) {
  (void)cls;
  FieldIds f;
  struct { jfieldID* dst; jclass c; const char* name; } want[] = {
      {&f.read_bases, readDataHolder, "readBases"},     {&f.read_quals, readDataHolder, "readQuals"},
      {&f.ins_gop, readDataHolder, "insertionGOP"},     {&f.del_gop, readDataHolder, "deletionGOP"},
      {&f.gcp, readDataHolder, "overallGCP"},           {&f.hap_bases, haplotypeDataHolder, "haplotypeBases"}};
  for (auto& w : want) {
    *w.dst = env->GetFieldID(w.c, w.name, "[B");
    if (*w.dst == nullptr) {
      throw_java(env, "java/lang/IllegalArgumentException", "Unable to get field ID");
      return;
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_fid_mu);
    g_fid = f;
  }
  const int rc = gklb_pairhmm_init(use_double ? 1 : 0, (int)max_threads);
  if (rc != GKLB_OK) throw_status(env, rc);
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative(
    JNIEnv* env, jobject obj, jobjectArray readDataArray, jobjectArray haplotypeDataArray,
    jdoubleArray likelihoodArray) {
  (void)obj;
  FieldIds f;
  {
    std::lock_guard<std::mutex> lk(g_fid_mu);
    f = g_fid;
  }
  if (!f.read_bases) { throw_java(env, "java/lang/IllegalStateException", "initNative has not been called"); return; }
  const int n_reads = env->GetArrayLength(readDataArray);
  const int n_haps = env->GetArrayLength(haplotypeDataArray);
  if (n_reads == 0 || n_haps == 0) return;  // GKL's pair loop runs zero iterations

  Arena hap, bases, quals, ins, del, gcp;
  for (int h = 0; h < n_haps; h++)
    if (append_field(env, haplotypeDataArray, h, f.hap_bases, hap, -1, true) < 0) return;

  const int block = pipeline_block_reads();
  if (n_reads >= 2 * block && gklb_pairhmm_devices_in_use() == 1) {
    if ((long long)env->GetArrayLength(likelihoodArray) < (long long)n_reads * n_haps) {
      throw_java(env, "java/lang/IllegalArgumentException", "likelihood array is shorter than reads x haplotypes");
      return;
    }
    gklb_engine* eng[2] = {nullptr, nullptr};
    int rc0 = gklb_pairhmm_acquire_engine(-1, &eng[0]);
    if (rc0 == GKLB_OK) {
      rc0 = gklb_pairhmm_acquire_engine(gklb_engine_device(eng[0]), &eng[1]);
      if (rc0 != GKLB_OK) gklb_pairhmm_release_engine(eng[0]);
    }
    if (rc0 != GKLB_OK) { throw_status(env, rc0); return; }
    jdouble* out = env->GetDoubleArrayElements(likelihoodArray, nullptr);
    if (!out) {
      for (auto* e : eng) gklb_pairhmm_release_engine(e);
      throw_java(env, "java/lang/OutOfMemoryError", "Unable to access jdoubleArray");
      return;
    }
    Arena blk[2][5];
    int rc = GKLB_OK;
    bool thrown = false;
    int k = 0;
    for (int r0 = 0; r0 < n_reads && rc == GKLB_OK && !thrown; r0 += block, k++) {
      const int r1 = r0 + block < n_reads ? r0 + block : n_reads;
      Arena* a = blk[k & 1];
      // the engine that used these arenas two blocks ago must be done with them (pinned staging is per engine,
      // pageable sources are consumed at submit; waiting also bounds the pinned output buffers)
      if (k >= 2) rc = gklb_engine_wait(eng[k & 1]);
      if (rc != GKLB_OK) break;
      for (int i = 0; i < 5; i++) { a[i].bytes.clear(); a[i].off.assign(1, 0); }
      for (int r = r0; r < r1 && !thrown; r++) {
        const int len = append_field(env, readDataArray, r, f.read_bases, a[0], -1, true);
        if (len < 0) { thrown = true; break; }
        if (append_field(env, readDataArray, r, f.ins_gop, a[2], len, false) < 0 ||
            append_field(env, readDataArray, r, f.del_gop, a[3], len, false) < 0 ||
            append_field(env, readDataArray, r, f.gcp, a[4], len, false) < 0 ||
            append_field(env, readDataArray, r, f.read_quals, a[1], len, false) < 0)
          thrown = true;
      }
      if (thrown) break;
      gklb_pairhmm_batch b;
      b.n_reads = r1 - r0;
      b.n_haps = n_haps;
      b.read_off = a[0].off.data();
      b.read_bases = a[0].bytes.data();
      b.read_quals = a[1].bytes.data();
      b.ins_gop = a[2].bytes.data();
      b.del_gop = a[3].bytes.data();
      b.gcp = a[4].bytes.data();
      b.hap_off = hap.off.data();
      b.hap_bases = hap.bytes.data();
      rc = gklb_engine_submit(eng[k & 1], &b, out + (size_t)r0 * n_haps);
    }
    for (auto* e : eng) {
      const int w = gklb_engine_wait(e);
      if (rc == GKLB_OK) rc = w;
      gklb_pairhmm_release_engine(e);
    }
    env->ReleaseDoubleArrayElements(likelihoodArray, out, (rc == GKLB_OK && !thrown) ? 0 : JNI_ABORT);
    if (rc != GKLB_OK && !thrown) throw_status(env, rc);
    return;
  }

  for (int r = 0; r < n_reads; r++) {
    const int len = append_field(env, readDataArray, r, f.read_bases, bases, -1, true);
    if (len < 0) return;
    if (append_field(env, readDataArray, r, f.ins_gop, ins, len, false) < 0) return;
    if (append_field(env, readDataArray, r, f.del_gop, del, len, false) < 0) return;
    if (append_field(env, readDataArray, r, f.gcp, gcp, len, false) < 0) return;
    if (append_field(env, readDataArray, r, f.read_quals, quals, len, false) < 0) return;
  }

  const long long need = (long long)n_reads * n_haps;
  if ((long long)env->GetArrayLength(likelihoodArray) < need) {
    throw_java(env, "java/lang/IllegalArgumentException", "likelihood array is shorter than reads x haplotypes");
    return;
  }
  jdouble* out = env->GetDoubleArrayElements(likelihoodArray, nullptr);
  if (!out) { throw_java(env, "java/lang/OutOfMemoryError", "Unable to access jdoubleArray"); return; }

  gklb_pairhmm_batch b;
  b.n_reads = n_reads;
  b.n_haps = n_haps;
  b.read_off = bases.off.data();
  b.read_bases = bases.bytes.data();
  b.read_quals = quals.bytes.data();
  b.ins_gop = ins.bytes.data();
  b.del_gop = del.bytes.data();
  b.gcp = gcp.bytes.data();
  b.hap_off = hap.off.data();
  b.hap_bases = hap.bytes.data();
  const int rc = gklb_pairhmm_compute(&b, out);
  // mode 0: copy back (if the JVM handed out a copy) and unpin, as GKL's ~JavaData does (JavaData.h:118-125)
  env->ReleaseDoubleArrayElements(likelihoodArray, out, rc == GKLB_OK ? 0 : JNI_ABORT);
  if (rc != GKLB_OK) throw_status(env, rc);
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative(JNIEnv* env, jobject obj) {
  (void)env;
  (void)obj;
  gklb_pairhmm_done();  // GKL's is empty; ours drops a reference, the last one frees the idle engines.  Idempotent.
}

}  // extern "C"
