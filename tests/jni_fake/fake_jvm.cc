// fake_jvm.cc -- TEST INFRASTRUCTURE: a minimal in-process stand-in for a JVM, enough to drive the JNI
// exports of libgkl_pairhmm.so the way com.intel.gkl.pairhmm.IntelPairHmm does (there is no JDK in this
// image).  It implements the JNI interface-table slots the library uses, over plain C++ objects, records
// the pending exception, and counts outstanding local references and array pins so that tests can assert
// GKL's ownership conventions.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../gkl_b200/csrc/jni_min.h"

namespace {

enum Kind { K_PLAIN, K_CLASS, K_BYTES, K_DOUBLES, K_HOLDER, K_OBJARR, K_LONGS };
struct Obj : _jobject { Kind kind = K_PLAIN; };
struct ClassObj : Obj { ClassObj() { kind = K_CLASS; } std::string name; std::vector<std::string> fields; };
struct ByteArr : Obj { ByteArr() { kind = K_BYTES; } std::vector<int8_t> data; };
struct DblArr : Obj { DblArr() { kind = K_DOUBLES; } std::vector<double> data; };
struct LongArr : Obj { LongArr() { kind = K_LONGS; } std::vector<int64_t> data; };
struct Holder : Obj { Holder() { kind = K_HOLDER; } ClassObj* cls; std::vector<ByteArr*> vals; };
struct ObjArr : Obj { ObjArr() { kind = K_OBJARR; } std::vector<_jobject*> elems; };

struct Vm {
  JNIEnv env;
  JNINativeInterface_ table;
  std::string exc_class, exc_msg;
  bool exc = false;
  long locals = 0, pins = 0, criticals = 0;
  bool fail_double_pin = false;
  const void* fail_critical_of = nullptr;  // GetPrimitiveArrayCritical of this array returns NULL
  std::vector<ClassObj*> found;
  std::vector<std::pair<double*, DblArr*>> dbl_copies;
};
thread_local Vm* g_vm = nullptr;  // one fake JVM state per thread, like a JNIEnv

jclass fFindClass(JNIEnv*, const char* name) {
  ClassObj* c = new ClassObj;
  c->name = name;
  g_vm->found.push_back(c);
  return c;
}
jint fThrowNew(JNIEnv*, jclass c, const char* msg) {
  g_vm->exc = true;
  g_vm->exc_class = static_cast<ClassObj*>(c)->name;
  g_vm->exc_msg = msg ? msg : "";
  return 0;
}
void fExceptionClear(JNIEnv*) { g_vm->exc = false; }
jboolean fExceptionCheck(JNIEnv*) { return g_vm->exc; }
jint fPushLocalFrame(JNIEnv*, jint) { return 0; }
jobject fPopLocalFrame(JNIEnv*, jobject r) { return r; }
void fDeleteLocalRef(JNIEnv*, jobject o) { if (o) g_vm->locals--; }
jfieldID fGetFieldID(JNIEnv*, jclass c, const char* name, const char* sig) {
  ClassObj* k = static_cast<ClassObj*>(c);
  if (strcmp(sig, "[B") != 0) return nullptr;
  for (size_t i = 0; i < k->fields.size(); i++)
    if (k->fields[i] == name) return reinterpret_cast<jfieldID>(i + 1);
  return nullptr;
}
jobject fGetObjectField(JNIEnv*, jobject o, jfieldID f) {
  Holder* h = static_cast<Holder*>(o);
  ByteArr* a = h->vals[reinterpret_cast<size_t>(f) - 1];
  if (a) g_vm->locals++;
  return a;
}
jsize fGetArrayLength(JNIEnv*, jarray a) {
  Obj* o = static_cast<Obj*>(a);
  if (o->kind == K_OBJARR) return (jsize)static_cast<ObjArr*>(o)->elems.size();
  if (o->kind == K_BYTES) return (jsize)static_cast<ByteArr*>(o)->data.size();
  if (o->kind == K_DOUBLES) return (jsize)static_cast<DblArr*>(o)->data.size();
  if (o->kind == K_LONGS) return (jsize)static_cast<LongArr*>(o)->data.size();
  return 0;
}
jobject fGetObjectArrayElement(JNIEnv*, jobjectArray a, jsize i) {
  _jobject* o = static_cast<ObjArr*>(a)->elems[i];
  if (o) g_vm->locals++;
  return o;
}
jbyte* fGetByteArrayElements(JNIEnv*, jbyteArray a, jboolean* is_copy) {
  if (is_copy) *is_copy = JNI_FALSE;
  g_vm->pins++;
  return static_cast<ByteArr*>(a)->data.data();
}
void fReleaseByteArrayElements(JNIEnv*, jbyteArray, jbyte*, jint) { g_vm->pins--; }
jdouble* fGetDoubleArrayElements(JNIEnv*, jdoubleArray a, jboolean* is_copy) {
  if (g_vm->fail_double_pin) return nullptr;
  DblArr* d = static_cast<DblArr*>(a);
  double* copy = new double[d->data.size() + 1];  // always hand out a COPY: the library must release with mode 0
  memcpy(copy, d->data.data(), sizeof(double) * d->data.size());
  if (is_copy) *is_copy = JNI_TRUE;
  g_vm->pins++;
  g_vm->dbl_copies.push_back({copy, d});
  return copy;
}
void fReleaseDoubleArrayElements(JNIEnv*, jdoubleArray a, jdouble* p, jint mode) {
  DblArr* d = static_cast<DblArr*>(a);
  if (mode != JNI_ABORT) memcpy(d->data.data(), p, sizeof(double) * d->data.size());
  g_vm->pins--;
  delete[] p;
}
void fGetByteArrayRegion(JNIEnv*, jbyteArray a, jsize start, jsize len, jbyte* buf) {
  memcpy(buf, static_cast<ByteArr*>(a)->data.data() + start, (size_t)len);
}

jdoubleArray fNewDoubleArray(JNIEnv*, jsize n) {
  DblArr* d = new DblArr;
  d->data.assign((size_t)n, 0.0);
  g_vm->locals++;
  return d;
}
void fSetDoubleArrayRegion(JNIEnv*, jdoubleArray a, jsize start, jsize len, const jdouble* buf) {
  memcpy(static_cast<DblArr*>(a)->data.data() + start, buf, sizeof(double) * (size_t)len);
}
void* fGetPrimitiveArrayCritical(JNIEnv*, jarray a, jboolean* is_copy) {
  if (is_copy) *is_copy = JNI_FALSE;
  if (a == g_vm->fail_critical_of) return nullptr;
  g_vm->pins++;
  g_vm->criticals++;
  Obj* o = static_cast<Obj*>(a);
  if (o->kind == K_BYTES) return static_cast<ByteArr*>(o)->data.data();
  if (o->kind == K_LONGS) return static_cast<LongArr*>(o)->data.data();
  return static_cast<DblArr*>(o)->data.data();
}
void fReleasePrimitiveArrayCritical(JNIEnv*, jarray, void*, jint) {
  g_vm->pins--;
  g_vm->criticals--;
}

void unimplemented() {
  fprintf(stderr, "fake_jvm: the library called a JNI slot this stand-in does not implement\n");
  abort();
}

void init_vm(Vm* vm) {
  for (auto& f : vm->table.fn) f = reinterpret_cast<void*>(&unimplemented);
  vm->table.fn[kJniFindClass] = (void*)&fFindClass;
  vm->table.fn[kJniThrowNew] = (void*)&fThrowNew;
  vm->table.fn[kJniExceptionClear] = (void*)&fExceptionClear;
  vm->table.fn[kJniExceptionCheck] = (void*)&fExceptionCheck;
  vm->table.fn[kJniPushLocalFrame] = (void*)&fPushLocalFrame;
  vm->table.fn[kJniPopLocalFrame] = (void*)&fPopLocalFrame;
  vm->table.fn[kJniDeleteLocalRef] = (void*)&fDeleteLocalRef;
  vm->table.fn[kJniGetFieldID] = (void*)&fGetFieldID;
  vm->table.fn[kJniGetObjectField] = (void*)&fGetObjectField;
  vm->table.fn[kJniGetArrayLength] = (void*)&fGetArrayLength;
  vm->table.fn[kJniGetObjectArrayElement] = (void*)&fGetObjectArrayElement;
  vm->table.fn[kJniGetByteArrayElements] = (void*)&fGetByteArrayElements;
  vm->table.fn[kJniReleaseByteArrayElements] = (void*)&fReleaseByteArrayElements;
  vm->table.fn[kJniGetDoubleArrayElements] = (void*)&fGetDoubleArrayElements;
  vm->table.fn[kJniReleaseDoubleArrayElements] = (void*)&fReleaseDoubleArrayElements;
  vm->table.fn[kJniGetByteArrayRegion] = (void*)&fGetByteArrayRegion;
  vm->table.fn[kJniNewDoubleArray] = (void*)&fNewDoubleArray;
  vm->table.fn[kJniSetDoubleArrayRegion] = (void*)&fSetDoubleArrayRegion;
  vm->table.fn[kJniGetPrimitiveArrayCritical] = (void*)&fGetPrimitiveArrayCritical;
  vm->table.fn[kJniReleasePrimitiveArrayCritical] = (void*)&fReleasePrimitiveArrayCritical;
  vm->env.functions = &vm->table;
}

ByteArr* make_bytes(const uint8_t* p, int64_t a, int64_t b) {
  ByteArr* r = new ByteArr;
  r->data.assign(reinterpret_cast<const int8_t*>(p) + a, reinterpret_cast<const int8_t*>(p) + b);
  return r;
}

typedef jint (*OnLoadFn)(JavaVM*, void*);
typedef void (*InitFn)(JNIEnv*, jclass, jclass, jclass, jboolean, jint);
typedef void (*ComputeFn)(JNIEnv*, jobject, jobjectArray, jobjectArray, jdoubleArray);
typedef void (*DoneFn)(JNIEnv*, jobject);

}  // namespace

extern "C" {

// JNI_OnLoad result of the library (JNI_VERSION_1_6 with a usable GPU, JNI_ERR without).
int fakejvm_onload(const char* lib_path) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1000;
  OnLoadFn f = (OnLoadFn)dlsym(h, "JNI_OnLoad");
  return f ? f(nullptr, nullptr) : -1001;
}

// fault: 0 none, 1 ReadDataHolder lacks overallGCP, 2 null element in the read array, 3 output array cannot be
// pinned, 4 readQuals shorter than readBases, 5 compute without initNative, 6 output array too short,
// 7 initialise, run done, initialise again and compute (done/re-init cycle)
// Returns 0 (no exception), 1 (exception pending: class/message copied out), <0 harness failure.
int fakejvm_pairhmm(const char* lib_path, int n_reads, int n_haps, const int64_t* read_off, const uint8_t* bases,
                    const uint8_t* quals, const uint8_t* ins, const uint8_t* del, const uint8_t* gcp,
                    const int64_t* hap_off, const uint8_t* haps, int use_double, int fault, double* out,
                    char* exc_class, char* exc_msg, long* leaks) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { snprintf(exc_msg, 255, "%s", dlerror()); return -2; }
  InitFn init = (InitFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative");
  ComputeFn compute = (ComputeFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative");
  DoneFn done = (DoneFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative");
  if (!init || !compute || !done) return -3;

  Vm vm;
  init_vm(&vm);
  g_vm = &vm;
  ClassObj read_cls, hap_cls, self_cls;
  read_cls.name = "org/broadinstitute/gatk/nativebindings/pairhmm/ReadDataHolder";
  read_cls.fields = {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"};
  if (fault == 1) read_cls.fields.pop_back();
  hap_cls.name = "org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder";
  hap_cls.fields = {"haplotypeBases"};
  self_cls.name = "com/intel/gkl/pairhmm/IntelPairHmm";
  Obj self;

  ObjArr reads, hap_arr;
  for (int r = 0; r < n_reads; r++) {
    Holder* o = new Holder;
    o->cls = &read_cls;
    int64_t a = read_off[r], b = read_off[r + 1];
    o->vals = {make_bytes(bases, a, b), make_bytes(quals, a, (fault == 4 && r == 0) ? b - 1 : b), make_bytes(ins, a, b),
               make_bytes(del, a, b), make_bytes(gcp, a, b)};
    reads.elems.push_back(o);
  }
  if (fault == 2 && n_reads > 0) reads.elems[n_reads / 2] = nullptr;
  for (int i = 0; i < n_haps; i++) {
    Holder* o = new Holder;
    o->cls = &hap_cls;
    o->vals = {make_bytes(haps, hap_off[i], hap_off[i + 1])};
    hap_arr.elems.push_back(o);
  }
  DblArr result;
  result.data.assign((size_t)n_reads * n_haps - (fault == 6 ? 1 : 0), -12345.0);
  vm.fail_double_pin = (fault == 3);

  if (fault != 5) init(&vm.env, &self_cls, &read_cls, &hap_cls, (jboolean)use_double, 1);
  if (!vm.exc && fault == 7) {
    done(&vm.env, &self);
    done(&vm.env, &self);
    init(&vm.env, &self_cls, &read_cls, &hap_cls, (jboolean)use_double, 1);
  }
  if (!vm.exc) compute(&vm.env, &self, &reads, &hap_arr, &result);
  const bool exc = vm.exc;
  if (exc) {
    snprintf(exc_class, 255, "%s", vm.exc_class.c_str());
    snprintf(exc_msg, 255, "%s", vm.exc_msg.c_str());
  } else {
    memcpy(out, result.data.data(), sizeof(double) * result.data.size());
  }
  done(&vm.env, &self);
  leaks[0] = vm.locals;
  leaks[1] = vm.pins;
  g_vm = nullptr;
  return exc ? 1 : 0;
}


// Concurrent callers: n_threads threads, each with its own JNIEnv, its own IntelPairHmm "instance" (initNative at
// the start, doneNative at the end) and every n_threads-th region; every region is computed `rounds` times.
// GATK's Spark executors call computeLikelihoods from several threads (IntelPairHmm.java:65 synchronises only load).
// Returns the number of calls that ended with a pending exception (message of the last one copied out).
int fakejvm_pairhmm_mt(const char* lib_path, int n_threads, int n_regions, const int* n_reads, const int* n_haps,
                       const int64_t* const* read_off, const uint8_t* const* bases, const uint8_t* const* quals,
                       const uint8_t* const* ins, const uint8_t* const* del, const uint8_t* const* gcp,
                       const int64_t* const* hap_off, const uint8_t* const* haps, double* const* outs, int rounds,
                       long* leaks, char* exc_msg) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { snprintf(exc_msg, 255, "%s", dlerror()); return -2; }
  InitFn init = (InitFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative");
  ComputeFn compute = (ComputeFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative");
  DoneFn done = (DoneFn)dlsym(h, "Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative");
  if (!init || !compute || !done) return -3;
  std::atomic<int> failures(0);
  std::atomic<long> leaked_locals(0), leaked_pins(0);
  std::vector<std::string> msgs(n_threads);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    th.emplace_back([&, t] {
      Vm vm;
      init_vm(&vm);
      g_vm = &vm;
      ClassObj read_cls, hap_cls, self_cls;
      read_cls.fields = {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"};
      hap_cls.fields = {"haplotypeBases"};
      Obj self;
      init(&vm.env, &self_cls, &read_cls, &hap_cls, JNI_FALSE, 1);
      if (vm.exc) { failures++; msgs[t] = vm.exc_class + ": " + vm.exc_msg; g_vm = nullptr; return; }
      for (int round = 0; round < rounds; round++)
        for (int k = t; k < n_regions; k += n_threads) {
          ObjArr reads, hap_arr;
          for (int r = 0; r < n_reads[k]; r++) {
            Holder* o = new Holder;
            o->cls = &read_cls;
            const int64_t a = read_off[k][r], b = read_off[k][r + 1];
            o->vals = {make_bytes(bases[k], a, b), make_bytes(quals[k], a, b), make_bytes(ins[k], a, b), make_bytes(del[k], a, b),
                       make_bytes(gcp[k], a, b)};
            reads.elems.push_back(o);
          }
          for (int i = 0; i < n_haps[k]; i++) {
            Holder* o = new Holder;
            o->cls = &hap_cls;
            o->vals = {make_bytes(haps[k], hap_off[k][i], hap_off[k][i + 1])};
            hap_arr.elems.push_back(o);
          }
          DblArr result;
          result.data.assign((size_t)n_reads[k] * n_haps[k], -12345.0);
          compute(&vm.env, &self, &reads, &hap_arr, &result);
          if (vm.exc) {
            failures++;
            msgs[t] = vm.exc_class + ": " + vm.exc_msg;
            vm.exc = false;
          } else {
            memcpy(outs[k], result.data.data(), sizeof(double) * result.data.size());
          }
          for (auto* o : reads.elems) { for (auto* v : static_cast<Holder*>(o)->vals) delete v; delete static_cast<Holder*>(o); }
          for (auto* o : hap_arr.elems) { for (auto* v : static_cast<Holder*>(o)->vals) delete v; delete static_cast<Holder*>(o); }
        }
      done(&vm.env, &self);
      leaked_locals += vm.locals;
      leaked_pins += vm.pins;
      g_vm = nullptr;
    });
  }
  for (auto& t : th) t.join();
  leaks[0] = leaked_locals;
  leaks[1] = leaked_pins;
  for (auto& m : msgs)
    if (!m.empty()) snprintf(exc_msg, 255, "%s", m.c_str());
  return failures;
}

// PDHMM binding (com.intel.gkl.pdhmm.IntelPDHMM).  object_api = 0: computePDHMMNative on the flat arrays
// (n pairs).  object_api = 1: computeLikelihoodsNative on n_reads ReadDataHolders x n_haps HaplotypeDataHolders
// built from the strided operands (read r at r * max_read, haplotype h at h * max_hap).
// fault: 0 none, 1 HaplotypeDataHolder lacks haplotypePDBases, 2 flat array of the wrong size
int fakejvm_pdhmm(const char* lib_path, int object_api, long long n, int n_reads, int n_haps, int max_hap, int max_read,
                  const int8_t* hap, const int8_t* pd, const int8_t* rb, const int8_t* rq, const int8_t* ri,
                  const int8_t* rd, const int8_t* rg, const int64_t* hap_len, const int64_t* read_len, int fault,
                  double* out, char* exc_class, char* exc_msg, long* leaks) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { snprintf(exc_msg, 255, "%s", dlerror()); return -2; }
  typedef void (*PInit)(JNIEnv*, jclass, jclass, jclass, jint, jint, jint, jint);
  typedef void (*PLik)(JNIEnv*, jobject, jobjectArray, jobjectArray, jdoubleArray);
  typedef jdoubleArray (*PFlat)(JNIEnv*, jobject, jbyteArray, jbyteArray, jbyteArray, jbyteArray, jbyteArray, jbyteArray,
                                jbyteArray, jlongArray, jlongArray, jint, jint, jint);
  typedef void (*PDone)(JNIEnv*, jclass);
  PInit init = (PInit)dlsym(h, "Java_com_intel_gkl_pdhmm_IntelPDHMM_initNative");
  PLik lik = (PLik)dlsym(h, "Java_com_intel_gkl_pdhmm_IntelPDHMM_computeLikelihoodsNative");
  PFlat flat = (PFlat)dlsym(h, "Java_com_intel_gkl_pdhmm_IntelPDHMM_computePDHMMNative");
  PDone done = (PDone)dlsym(h, "Java_com_intel_gkl_pdhmm_IntelPDHMM_doneNative");
  if (!init || !lik || !flat || !done) return -3;
  Vm vm;
  init_vm(&vm);
  g_vm = &vm;
  ClassObj read_cls, hap_cls, self_cls;
  read_cls.fields = {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"};
  hap_cls.fields = {"haplotypeBases", "haplotypePDBases"};
  if (fault == 1) hap_cls.fields.pop_back();
  Obj self;
  init(&vm.env, &self_cls, &read_cls, &hap_cls, 0, 2, 0, 10);
  long long n_out = 0;
  if (!vm.exc && object_api) {
    ObjArr reads, haps;
    for (int r = 0; r < n_reads; r++) {
      Holder* o = new Holder;
      const int64_t a = (int64_t)r * max_read, b = a + read_len[r];
      o->vals = {make_bytes((const uint8_t*)rb, a, b), make_bytes((const uint8_t*)rq, a, b), make_bytes((const uint8_t*)ri, a, b),
                 make_bytes((const uint8_t*)rd, a, b), make_bytes((const uint8_t*)rg, a, b)};
      reads.elems.push_back(o);
    }
    for (int i = 0; i < n_haps; i++) {
      Holder* o = new Holder;
      const int64_t a = (int64_t)i * max_hap, b = a + hap_len[i];
      o->vals = {make_bytes((const uint8_t*)hap, a, b), make_bytes((const uint8_t*)pd, a, b)};
      haps.elems.push_back(o);
    }
    DblArr result;
    n_out = (long long)n_reads * n_haps;
    result.data.assign((size_t)n_out, -12345.0);
    lik(&vm.env, &self, &reads, &haps, &result);
    if (!vm.exc) memcpy(out, result.data.data(), sizeof(double) * (size_t)n_out);
  } else if (!vm.exc) {
    ByteArr* a[7];
    const int8_t* src[7] = {hap, pd, rb, rq, ri, rd, rg};
    for (int i = 0; i < 7; i++) {
      const long long len = n * (i < 2 ? max_hap : max_read) - ((fault == 2 && i == 3) ? 1 : 0);
      a[i] = make_bytes((const uint8_t*)src[i], 0, len);
    }
    LongArr hl, rl;
    hl.data.assign(hap_len, hap_len + n);
    rl.data.assign(read_len, read_len + n);
    jdoubleArray r = flat(&vm.env, &self, a[0], a[1], a[2], a[3], a[4], a[5], a[6], &hl, &rl, (jint)n, max_hap, max_read);
    if (!vm.exc && r) {
      memcpy(out, static_cast<DblArr*>(r)->data.data(), sizeof(double) * (size_t)n);
      vm.locals--;  // the returned array is handed to the caller
    }
  }
  const bool exc = vm.exc;
  if (exc) {
    snprintf(exc_class, 255, "%s", vm.exc_class.c_str());
    snprintf(exc_msg, 255, "%s", vm.exc_msg.c_str());
  }
  done(&vm.env, &self_cls);
  leaks[0] = vm.locals;
  leaks[1] = vm.pins;
  g_vm = nullptr;
  return exc ? 1 : 0;
}

// IntelSmithWaterman: initNative, one alignNative call, doneNative.  fault: 0 none, 1 the reference array cannot be
// pinned (GetPrimitiveArrayCritical returns NULL).  cigar: `cigar_cap` bytes, zeroed here like a fresh Java byte[].
// Returns 0 (no exception; *offset set), 1 (exception pending), <0 harness failure.
int fakejvm_sw(const char* lib_path, const uint8_t* ref, int ref_len, const uint8_t* alt, int alt_len, int match,
               int mismatch, int open, int extend, int strategy, int fault, char* cigar, int cigar_cap, int* offset,
               char* exc_class, char* exc_msg, long* leaks) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { snprintf(exc_msg, 255, "%s", dlerror()); return -2; }
  typedef void (*SInit)(JNIEnv*, jclass);
  typedef jint (*SAlign)(JNIEnv*, jclass, jbyteArray, jbyteArray, jbyteArray, jint, jint, jint, jint, jbyte);
  SInit init = (SInit)dlsym(h, "Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_initNative");
  SAlign align = (SAlign)dlsym(h, "Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_alignNative");
  SInit done = (SInit)dlsym(h, "Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_doneNative");
  if (!init || !align || !done) return -3;
  Vm vm;
  init_vm(&vm);
  g_vm = &vm;
  ClassObj self_cls;
  init(&vm.env, &self_cls);
  if (!vm.exc) {
    ByteArr* r = make_bytes(ref, 0, ref_len);
    ByteArr* a = make_bytes(alt, 0, alt_len);
    ByteArr* c = new ByteArr;
    c->data.assign((size_t)cigar_cap, 0);
    if (fault == 1) vm.fail_critical_of = r;
    const jint off = align(&vm.env, &self_cls, r, a, c, match, mismatch, open, extend, (jbyte)strategy);
    if (!vm.exc) {
      *offset = off;
      memcpy(cigar, c->data.data(), (size_t)cigar_cap);
    }
    delete r; delete a; delete c;
  }
  const bool exc = vm.exc;
  if (exc) {
    snprintf(exc_class, 255, "%s", vm.exc_class.c_str());
    snprintf(exc_msg, 255, "%s", vm.exc_msg.c_str());
  }
  done(&vm.env, &self_cls);
  leaks[0] = vm.locals;
  leaks[1] = vm.pins;
  g_vm = nullptr;
  return exc ? 1 : 0;
}

}  // extern "C"
