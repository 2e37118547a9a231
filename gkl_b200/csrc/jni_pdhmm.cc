// jni_pdhmm.cc -- the JNI face of libgkl_pdhmm.so: the four symbols GKL's unchanged Java class
// com.intel.gkl.pdhmm.IntelPDHMM binds (IntelPDHMM.java:206-216).
//
//   Java_com_intel_gkl_pdhmm_IntelPDHMM_initNative                replaces pdhmm/IntelPDHMM.cc:43-60
//   Java_com_intel_gkl_pdhmm_IntelPDHMM_computeLikelihoodsNative  replaces pdhmm/IntelPDHMM.cc:62-137 + pdhmm/JavaData.h:41-242
//   Java_com_intel_gkl_pdhmm_IntelPDHMM_computePDHMMNative        replaces pdhmm/IntelPDHMM.cc:144-244
//   Java_com_intel_gkl_pdhmm_IntelPDHMM_doneNative                replaces pdhmm/IntelPDHMM.cc:246-249
//
// The object API copies every read and haplotype ONCE into max-length-strided operand arrays
// (GetByteArrayRegion, local references deleted as it goes, like pdhmm/JavaData.h:332-403) and lets the kernel
// address the reads x haplotypes cross product; the reference instead expands the cross product into flat
// host batches bounded by maxMemoryInMB (pdhmm/JavaData.h:177-242).
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/gklb_pdhmm.h"
#include "jni_min.h"

namespace {

struct PdFieldIds {
  jfieldID read_bases = nullptr, read_quals = nullptr, ins_gop = nullptr, del_gop = nullptr, gcp = nullptr,
           hap_bases = nullptr, hap_pd = nullptr;
};
PdFieldIds g_fid;
std::mutex g_mu;

void throw_java(JNIEnv* env, const char* cls, const char* msg) {
  env->ExceptionClear();
  jclass c = env->FindClass(cls);
  if (c) env->ThrowNew(c, msg);
}

void throw_status(JNIEnv* env, int rc) {
  // the reference's status -> exception mapping (pdhmm/IntelPDHMM.cc:94-115, pdhmm-common.h:38-42)
  switch (rc) {
    case GKLB_ERR_OOM: throw_java(env, "java/lang/OutOfMemoryError", "Memory allocation issue."); break;
    case GKLB_ERR_INVALID:
      throw_java(env, "java/lang/IllegalArgumentException", "Error while calculating PDHMM. Input arrays aren't valid.");
      break;
    default: throw_java(env, "java/lang/RuntimeException", gklb_last_error()); break;
  }
}

// object[index].field -> dst[index * stride ...]; returns the array length or -1 after throwing
int fetch(JNIEnv* env, jobjectArray array, int index, jfieldID fid, std::vector<int8_t>* bytes) {
  jobject obj = env->GetObjectArrayElement(array, index);
  if (!obj) { throw_java(env, "java/lang/NullPointerException", "null element in input array"); return -1; }
  jbyteArray a = (jbyteArray)env->GetObjectField(obj, fid);
  if (!a) {
    env->DeleteLocalRef(obj);
    throw_java(env, "java/lang/NullPointerException", "null byte[] in data holder");
    return -1;
  }
  const int len = env->GetArrayLength(a);
  bytes->resize((size_t)len);
  if (len > 0) env->GetByteArrayRegion(a, 0, len, reinterpret_cast<jbyte*>(bytes->data()));
  env->DeleteLocalRef(a);
  env->DeleteLocalRef(obj);
  return len;
}

}  // namespace

extern "C" {

JNIEXPORT void JNICALL Java_com_intel_gkl_pdhmm_IntelPDHMM_initNative(JNIEnv* env, jclass cls, jclass readDataHolder,
                                                                      jclass haplotypeDataHolder, jint openMPSetting,
                                                                      jint max_threads, jint avxLevel,
                                                                      jint maxMemoryInMB) {
  (void)cls;
  PdFieldIds f;
  struct { jfieldID* dst; jclass c; const char* name; } want[] = {
      {&f.read_bases, readDataHolder, "readBases"},       {&f.read_quals, readDataHolder, "readQuals"},
      {&f.ins_gop, readDataHolder, "insertionGOP"},       {&f.del_gop, readDataHolder, "deletionGOP"},
      {&f.gcp, readDataHolder, "overallGCP"},             {&f.hap_bases, haplotypeDataHolder, "haplotypeBases"},
      {&f.hap_pd, haplotypeDataHolder, "haplotypePDBases"}};
  for (auto& w : want) {
    *w.dst = env->GetFieldID(w.c, w.name, "[B");
    if (!*w.dst) { throw_java(env, "java/lang/IllegalArgumentException", "Unable to get field ID"); return; }
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_fid = f;
  }
  const int rc = gklb_pdhmm_init((int)openMPSetting, (int)max_threads, (int)avxLevel, (int)maxMemoryInMB);
  if (rc != GKLB_OK) throw_java(env, rc == GKLB_ERR_OOM ? "java/lang/OutOfMemoryError" : "java/lang/IllegalStateException",
                                gklb_last_error());
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pdhmm_IntelPDHMM_computeLikelihoodsNative(JNIEnv* env, jobject obj,
                                                                                    jobjectArray readDataArray,
                                                                                    jobjectArray haplotypeDataArray,
                                                                                    jdoubleArray likelihoodArray) {
  (void)obj;
  PdFieldIds f;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    f = g_fid;
  }
  if (!f.read_bases) { throw_java(env, "java/lang/IllegalStateException", "initNative has not been called"); return; }
  const int n_reads = env->GetArrayLength(readDataArray), n_haps = env->GetArrayLength(haplotypeDataArray);
  if (n_reads == 0 || n_haps == 0) return;
  if ((long long)env->GetArrayLength(likelihoodArray) != (long long)n_reads * n_haps) {
    throw_java(env, "java/lang/IllegalArgumentException", "likelihoodArray length must be reads x haplotypes");
    return;
  }
  // pass 1: copy every array out once, remember lengths
  std::vector<std::vector<int8_t>> hb(n_haps), hp(n_haps), rb(n_reads), rq(n_reads), ri(n_reads), rd(n_reads), rc_(n_reads);
  std::vector<int64_t> hl(n_haps), rl(n_reads);
  int max_hap = 1, max_read = 1;
  for (int h = 0; h < n_haps; h++) {
    const int len = fetch(env, haplotypeDataArray, h, f.hap_bases, &hb[h]);
    if (len < 0) return;
    if (fetch(env, haplotypeDataArray, h, f.hap_pd, &hp[h]) != len) {
      if (!env->ExceptionCheck()) throw_java(env, "java/lang/IllegalArgumentException", "haplotypePDBases length differs");
      return;
    }
    hl[h] = len;
    max_hap = len > max_hap ? len : max_hap;
  }
  for (int r = 0; r < n_reads; r++) {
    const int len = fetch(env, readDataArray, r, f.read_bases, &rb[r]);
    if (len < 0) return;
    std::vector<int8_t>* rest[4] = {&rq[r], &ri[r], &rd[r], &rc_[r]};
    jfieldID fids[4] = {f.read_quals, f.ins_gop, f.del_gop, f.gcp};
    for (int k = 0; k < 4; k++)
      if (fetch(env, readDataArray, r, fids[k], rest[k]) != len) {
        if (!env->ExceptionCheck()) throw_java(env, "java/lang/IllegalArgumentException", "per-read arrays differ in length");
        return;
      }
    rl[r] = len;
    max_read = len > max_read ? len : max_read;
  }
  // pass 2: zero-padded strided operands (the layout of pdhmm/JavaData.h:177-242, once per read / haplotype)
  std::vector<int8_t> H((size_t)n_haps * max_hap, 0), P((size_t)n_haps * max_hap, 0);
  std::vector<int8_t> R[5];
  for (auto& v : R) v.assign((size_t)n_reads * max_read, 0);
  for (int h = 0; h < n_haps; h++) {
    memcpy(H.data() + (size_t)h * max_hap, hb[h].data(), hb[h].size());
    memcpy(P.data() + (size_t)h * max_hap, hp[h].data(), hp[h].size());
  }
  for (int r = 0; r < n_reads; r++) {
    const std::vector<int8_t>* src[5] = {&rb[r], &rq[r], &ri[r], &rd[r], &rc_[r]};
    for (int k = 0; k < 5; k++) memcpy(R[k].data() + (size_t)r * max_read, src[k]->data(), src[k]->size());
  }
  jdouble* out = env->GetDoubleArrayElements(likelihoodArray, nullptr);
  if (!out) { throw_java(env, "java/lang/OutOfMemoryError", "Memory allocation issue."); return; }
  gklb_pdhmm_batch b;
  b.n = 0;
  b.max_hap = max_hap;
  b.max_read = max_read;
  b.hap_bases = H.data();
  b.hap_pdbases = P.data();
  b.read_bases = R[0].data();
  b.read_qual = R[1].data();
  b.read_ins_qual = R[2].data();
  b.read_del_qual = R[3].data();
  b.gcp = R[4].data();
  b.hap_lengths = hl.data();
  b.read_lengths = rl.data();
  const int rc = gklb_pdhmm_compute_cross(&b, n_reads, n_haps, out);
  env->ReleaseDoubleArrayElements(likelihoodArray, out, 0);  // the reference releases with mode 0 on every path
  if (rc != GKLB_OK) throw_status(env, rc);
}

JNIEXPORT jdoubleArray JNICALL Java_com_intel_gkl_pdhmm_IntelPDHMM_computePDHMMNative(
    JNIEnv* env, jobject obj, jbyteArray jhap_bases, jbyteArray jhap_pdbases, jbyteArray jread_bases,
    jbyteArray jread_qual, jbyteArray jread_ins_qual, jbyteArray jread_del_qual, jbyteArray jgcp,
    jlongArray jhap_lengths, jlongArray jread_lengths, jint testcase, jint maxHapLength, jint maxReadLength) {
  (void)obj;
  if (testcase <= 0 || maxHapLength <= 0 || maxReadLength <= 0) {
    throw_java(env, "java/lang/IllegalArgumentException", "batch size and maximum lengths must be greater than 0");
    return nullptr;
  }
  jarray arrays[9] = {jhap_bases, jhap_pdbases, jread_bases, jread_qual, jread_ins_qual, jread_del_qual, jgcp,
                      jhap_lengths, jread_lengths};
  const long long expect[9] = {(long long)testcase * maxHapLength, (long long)testcase * maxHapLength,
                               (long long)testcase * maxReadLength, (long long)testcase * maxReadLength,
                               (long long)testcase * maxReadLength, (long long)testcase * maxReadLength,
                               (long long)testcase * maxReadLength, testcase, testcase};
  for (int i = 0; i < 9; i++) {
    if (!arrays[i]) { throw_java(env, "java/lang/NullPointerException", "input array is null"); return nullptr; }
    if ((long long)env->GetArrayLength(arrays[i]) != expect[i]) {
      throw_java(env, "java/lang/IllegalArgumentException", "input array has the wrong size");
      return nullptr;
    }
  }
  // critical section: only memcpy inside (the reference holds the criticals across the whole computation,
  // pdhmm/IntelPDHMM.cc:160-222, which blocks the garbage collector for its duration)
  std::vector<int8_t> host[7];
  std::vector<int64_t> lens[2];
  for (int i = 0; i < 9; i++) {
    void* p = env->GetPrimitiveArrayCritical(arrays[i], nullptr);
    if (!p) { throw_java(env, "java/lang/OutOfMemoryError", "Memory allocation issue."); return nullptr; }
    if (i < 7) host[i].assign(static_cast<int8_t*>(p), static_cast<int8_t*>(p) + expect[i]);
    else lens[i - 7].assign(static_cast<int64_t*>(p), static_cast<int64_t*>(p) + expect[i]);
    env->ReleasePrimitiveArrayCritical(arrays[i], p, JNI_ABORT);
  }
  gklb_pdhmm_batch b;
  b.n = testcase;
  b.max_hap = maxHapLength;
  b.max_read = maxReadLength;
  b.hap_bases = host[0].data();
  b.hap_pdbases = host[1].data();
  b.read_bases = host[2].data();
  b.read_qual = host[3].data();
  b.read_ins_qual = host[4].data();
  b.read_del_qual = host[5].data();
  b.gcp = host[6].data();
  b.hap_lengths = lens[0].data();
  b.read_lengths = lens[1].data();
  std::vector<double> result((size_t)testcase);
  const int rc = gklb_pdhmm_compute(&b, result.data());
  if (rc != GKLB_OK) { throw_status(env, rc); return nullptr; }
  jdoubleArray out = env->NewDoubleArray(testcase);
  if (!out) { throw_java(env, "java/lang/OutOfMemoryError", "Memory allocation issue."); return nullptr; }
  env->SetDoubleArrayRegion(out, 0, testcase, result.data());
  return out;
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pdhmm_IntelPDHMM_doneNative(JNIEnv* env, jclass cls) {
  (void)env;
  (void)cls;
  gklb_pdhmm_done();
}

}  // extern "C"
