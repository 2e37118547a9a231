/*
 * gklb_pairhmm.h -- C-ABI of the B200-native PairHMM engine (libgkl_pairhmm.so).
 *
 * This is the drop-in boundary for GKL's PairHMM hot path.  Every entry point replaces one
 * piece of GKL's native side; citations are to the reference tree (/root/reference):
 *
 *   gklb_pairhmm_init      <-  Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative
 *                              src/main/native/pairhmm/IntelPairHmm.cc:55-118
 *                              (use_double / max_threads arguments, table construction of
 *                              pairhmm/Context.h:133-189 done once per process)
 *   gklb_pairhmm_compute   <-  Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative
 *                              src/main/native/pairhmm/IntelPairHmm.cc:125-181
 *                              (the flat batch replaces JavaData::getData's std::vector<testcase>,
 *                              pairhmm/JavaData.h:65-111; output index r * n_haps + h, :94-105)
 *   gklb_pairhmm_done      <-  Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative
 *                              src/main/native/pairhmm/IntelPairHmm.cc:189-192
 *
 * The same shared object also exports the three Java_com_intel_gkl_pairhmm_IntelPairHmm_*
 * JNI symbols (gkl_b200/csrc/jni_pairhmm.cc), which marshal Java arrays into a
 * gklb_pairhmm_batch and call the functions below; see INTEGRATION.md.
 *
 * Plain C types only: pointers, sizes, ints.  All functions return a gklb_status; on failure
 * gklb_last_error() holds a message for the calling thread.  There is no CPU fallback: if no
 * sm_100 device is usable every compute call fails with GKLB_ERR_NO_DEVICE.
 */
#ifndef GKLB_PAIRHMM_H
#define GKLB_PAIRHMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GKLB_API __attribute__((visibility("default")))
#else
#define GKLB_API
#endif

/* Status codes.  The JNI layer maps them onto the exception classes GKL throws
 * (IntelPairHmm.cc:64-68,141-145,171-178): OOM -> java/lang/OutOfMemoryError,
 * INVALID -> java/lang/IllegalArgumentException, everything else -> java/lang/RuntimeException. */
typedef enum gklb_status {
  GKLB_OK = 0,
  GKLB_ERR_OOM = 1,        /* host or device allocation failed */
  GKLB_ERR_INVALID = 2,    /* null pointer, negative size, empty sequence, non-monotonic offsets */
  GKLB_ERR_CUDA = 3,       /* CUDA runtime / launch failure */
  GKLB_ERR_NO_DEVICE = 4,  /* no CUDA device of compute capability 10.x */
  GKLB_ERR_STATE = 5       /* engine not initialised, or nothing staged */
} gklb_status;

/* One read x haplotype batch.  Offsets are host pointers in every entry point; the byte arenas
 * are host pointers for gklb_pairhmm_compute / gklb_pairhmm_stage and device pointers for
 * gklb_pairhmm_stage_device.  Read r occupies [read_off[r], read_off[r+1]) of each of the five
 * read arenas; haplotype h occupies [hap_off[h], hap_off[h+1]) of hap_bases.  Bytes are passed
 * exactly as GATK passes them to GKL: ASCII bases, raw (not +33) qualities; the engine applies
 * GKL's own `& 127` and base-code mapping (avx-pairhmm-template.h:134-136,149;
 * pairhmm_common.h:53-66). */
typedef struct gklb_pairhmm_batch {
  int32_t n_reads;
  int32_t n_haps;
  const int64_t* read_off;   /* [n_reads + 1], read_off[0] == 0 */
  const uint8_t* read_bases; /* testcase.rs  */
  const uint8_t* read_quals; /* testcase.q   */
  const uint8_t* ins_gop;    /* testcase.i   */
  const uint8_t* del_gop;    /* testcase.d   */
  const uint8_t* gcp;        /* testcase.c   */
  const int64_t* hap_off;    /* [n_haps + 1], hap_off[0] == 0 */
  const uint8_t* hap_bases;  /* testcase.hap */
} gklb_pairhmm_batch;

/* Counters of the most recent run on an engine (the reference computes
 * m_total_cells and never reports it, pairhmm/JavaData.h:108). */
typedef struct gklb_pairhmm_stats {
  int64_t pairs;            /* n_reads * n_haps */
  int64_t cells;            /* sum over pairs of rslen * haplen */
  int64_t fallback_pairs;   /* pairs whose scaled fp32 sum was < 1e-28f (IntelPairHmm.cc:159): recomputed by the
                               range-extended fp32 rerun, or by the fp64 kernel when its guards say so */
  int32_t kernel_launches;  /* kernels of this library launched by the last run */
  int32_t n_classes;        /* read-length classes (distinct kernel instantiations) used */
  float h2d_ms;             /* cudaEvent-timed phases of the last gklb_pairhmm_compute; 0 if not timed */
  float kernel_ms;
  float d2h_ms;
  float sweep_ms;           /* device time of the forward-sweep task kernels alone in the last run (the
                               dominant kernel; excludes packing and the fp64 rerun of flagged pairs) */
  int32_t sweep_launches;   /* how many launches sweep_ms covers */
  int64_t fp64_pairs;       /* of fallback_pairs, how many went through the fp64 kernel */
} gklb_pairhmm_stats;

typedef struct gklb_engine gklb_engine;

/* ---- process-global surface: what the JNI layer (and GKL's Java shim above it) sees ---- */

/* initNative.  use_double mirrors PairHMMNativeArguments.useDoublePrecision; max_threads is
 * accepted for signature compatibility and ignored (GKL's non-OpenMP library ignores it too,
 * IntelPairHmm.cc:85-89).  Device selection: env GKLB_DEVICE (one device, default 0) or GKLB_DEVICES
 * ("all" or "0,1,...").  The state behind this surface is a pool of engines (engine_global.cu): init and done
 * are reference counted -- GKL's native state is read-only after initNative and its doneNative is empty
 * (IntelPairHmm.cc:41-48,189-192), so several IntelPairHmm instances may initialise and close independently --
 * and a repeated init is cheap: engines are kept.  The precision of the last init wins, as with GKL's
 * g_use_double. */
GKLB_API int gklb_pairhmm_init(int use_double, int max_threads);

/* computeLikelihoodsNative.  Host arenas in, likelihoods[r * n_haps + h] (log10) out; synchronous.
 * n_reads == 0 or n_haps == 0 is a no-op returning GKLB_OK, like GKL.  Thread-safe and concurrent: every call
 * borrows its own engine from the pool, on the least busy configured device (Spark executors call
 * computeLikelihoods from several threads, IntelPairHmm.java:65 synchronises only load()).  With several
 * devices, batches of more than ~4e9 cells per device are sharded over reads inside the call (env GKLB_SHARD =
 * direct | nccl, see engine_global.cu / engine_nccl.cu); in the direct mode a device's share of more than ~2e11
 * cells runs as pieces on two engines in turn, so that the copies of one piece overlap the kernels of the other.
 * The result does not depend on any of this (bit-identical to a single engine's). */
GKLB_API int gklb_pairhmm_compute(const gklb_pairhmm_batch* batch, double* likelihoods);

/* Several regions in one call (what a GATK-side queue that coalesces active regions would call;
 * IntelPairHmm.computeLikelihoods is one synchronous call per region, IntelPairHmm.java:130-147).  All regions
 * are staged with one host->device copy and share the launches: the tasks of every region of a launch group are
 * pulled from one queue, so many small regions fill the GPU like one large batch.  likelihoods[r] receives
 * region r's matrix (double[n_reads * n_haps]); results are bit-identical to one gklb_pairhmm_compute per region.
 * Regions with no reads or no haplotypes are skipped.  Through this (global) entry point the regions are cut into
 * jobs of at most ~192 MB of staging, one job per configured device when every job still holds about a millisecond
 * of work; a region large enough to shard goes through gklb_pairhmm_compute on its own. */
GKLB_API int gklb_pairhmm_compute_multi(const gklb_pairhmm_batch* batches, int n_batches, double* const* likelihoods);

/* doneNative.  Drops one reference; the last one frees the idle engines (device memory, streams, events).
 * Idempotent; a later compute creates engines again. */
GKLB_API int gklb_pairhmm_done(void);

/* Number of devices the global surface was initialised with, engines currently alive in the pool, and the
 * counters of the last compute call that finished (summed over devices; phase times are the maximum over the devices --
 * for a device that ran its share in pieces: the first piece's way in, the span of its kernels, the last piece's way
 * out). */
GKLB_API int gklb_pairhmm_devices_in_use(void);
GKLB_API int gklb_pairhmm_engines_alive(void);
GKLB_API int gklb_pairhmm_last_stats(gklb_pairhmm_stats* out);

/* Borrow an engine of the pool for a sequence of gklb_engine_* calls (the JNI layer pipelines large calls over two
 * engines of one device); device < 0: the least busy configured device.  Must be given back. */
GKLB_API int gklb_pairhmm_acquire_engine(int device, gklb_engine** out);
GKLB_API int gklb_pairhmm_release_engine(gklb_engine* e);

/* ---- explicit engines: one per (host thread | device); used by the benchmark and multi-GPU host ---- */

GKLB_API int gklb_engine_create(gklb_engine** out, int device, int use_double);
GKLB_API int gklb_engine_destroy(gklb_engine* e);
GKLB_API int gklb_engine_device(gklb_engine* e);

/* Run all work of this engine on an existing CUDA stream (a cudaStream_t passed as void*), e.g.
 * the caller's current stream so that caller-side CUDA events bracket the kernels.  NULL restores
 * the engine's own stream. */
GKLB_API int gklb_engine_set_stream(gklb_engine* e, void* cuda_stream);

GKLB_API int gklb_engine_compute(gklb_engine* e, const gklb_pairhmm_batch* batch, double* likelihoods);
GKLB_API int gklb_engine_compute_multi(gklb_engine* e, const gklb_pairhmm_batch* batches, int n_batches,
                                       double* const* likelihoods);

/* Asynchronous form of compute, for callers that have host work to overlap with the GPU (the JNI layer marshals
 * the next block of reads while the previous one is being computed):
 *   submit  validate, plan, queue the copies and kernels on the engine's stream, queue the device->host copy of the
 *           likelihoods into a pinned buffer; returns without waiting.  Pageable input arenas may be reused as soon
 *           as submit returns; pinned ones only after wait.
 *   wait    block until the submitted batch is done and move the likelihoods into the array given to submit.
 * One batch may be in flight per engine. */
GKLB_API int gklb_engine_submit(gklb_engine* e, const gklb_pairhmm_batch* batch, double* likelihoods);
GKLB_API int gklb_engine_wait(gklb_engine* e);

/* Three-phase form of compute, for measuring the device-resident path:
 *   stage         validate, plan (length classes, work units), copy arenas host->device  [async]
 *   stage_device  same, but the six arenas are already device pointers (copied device->device
 *                 into the engine's padded buffers, so callers need not pad or align them)
 *   run           launch the kernels on the engine's stream; results stay in device memory [async]
 *   fetch         copy likelihoods device->host and synchronise
 *   result_device device pointer to the likelihoods (double[n_reads * n_haps]) of the last run
 * run may be called repeatedly on one staged batch. */
GKLB_API int gklb_engine_stage(gklb_engine* e, const gklb_pairhmm_batch* batch);
GKLB_API int gklb_engine_stage_device(gklb_engine* e, const gklb_pairhmm_batch* batch);
GKLB_API int gklb_engine_stage_multi(gklb_engine* e, const gklb_pairhmm_batch* batches, int n_batches);
/* Replace the haplotype bases of the staged batch by device-resident ones of the same lengths (e.g. the
 * buffer an NCCL broadcast just filled): the panel images are rewritten by a kernel on the engine's stream,
 * no host round trip.  [async] */
GKLB_API int gklb_engine_update_haps_device(gklb_engine* e, const void* hap_bases_dev);
GKLB_API int gklb_engine_run(gklb_engine* e);
GKLB_API int gklb_engine_fetch(gklb_engine* e, double* likelihoods);
GKLB_API int gklb_engine_result_device(gklb_engine* e, void** dev_ptr);
/* The result of the last run narrowed on the device to one packed buffer, for hosts that gather results over NVLink
 * (bench.py --gpus N; the in-process NCCL path does the same internally):
 *   float likelihoods[pairs] | uint32 count | uint32 pad | uint32 index[capacity] | (8-aligned) double value[capacity]
 * An unflagged pair's value IS an fp32 widened to double (IntelPairHmm.cc:164), so the fp32 matrix loses nothing; the
 * flagged pairs (fp64 results) travel as (pair index, value) overrides; count > capacity = the list overflowed.  [async] */
GKLB_API int gklb_engine_narrow(gklb_engine* e, unsigned int capacity, void** packed_dev, size_t* packed_bytes);
GKLB_API int gklb_engine_synchronize(gklb_engine* e);

GKLB_API int gklb_engine_stats(gklb_engine* e, gklb_pairhmm_stats* out);
/* Name of the forward-sweep kernel the staged batch's plan launches (for bench reports). */
GKLB_API const char* gklb_engine_sweep_kernel(gklb_engine* e);
/* Text description of the staged job's plan: regions, classes, tiles, launch groups, launches (grid, shared memory). */
GKLB_API int gklb_engine_plan_info(gklb_engine* e, char* buf, int n);

/* Time `iters` back-to-back runs of the staged batch with CUDA events recorded on the engine's
 * stream; returns the mean milliseconds per run in *ms_per_run. */
GKLB_API int gklb_engine_time_runs(gklb_engine* e, int iters, float* ms_per_run);

/* ---- misc ---- */

GKLB_API const char* gklb_last_error(void);
GKLB_API const char* gklb_version(void);
GKLB_API int gklb_device_count(void);

/* Host-built constant tables (Context.h:133-189), exposed so tests can pin them bit-for-bit.
 * which: 0 ph2pr float[128], 1 matchToMatch float[8256], 2 ph2pr double[128], 3 matchToMatch double[8256].
 * Returns a pointer valid for the process lifetime and the element count in *n. */
GKLB_API const void* gklb_pairhmm_table(int which, int* n);

#ifdef __cplusplus
}
#endif
#endif /* GKLB_PAIRHMM_H */
