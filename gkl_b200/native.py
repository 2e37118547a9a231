"""ctypes binding of libgkl_pairhmm.so (the C-ABI of include/gklb_pairhmm.h).

The library holds sm_100a code only.  There is no Python or CPU implementation behind these
calls: if the shared object is missing or no B200-class device is visible, they raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .batch import PairHmmBatch

PKG = Path(__file__).resolve().parent
# GKLB_LIB_DIR: another build of the libraries, e.g. gkl_b200/lib_exp (make EXPERIMENTAL=1: measurement-only kernels)
LIB_DIR = Path(os.environ["GKLB_LIB_DIR"]).resolve() if os.environ.get("GKLB_LIB_DIR") else PKG / "lib"
LIB_PATH = LIB_DIR / "libgkl_pairhmm.so"

OK, ERR_OOM, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE = range(6)


class GklbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gklb error {code}: {msg}")
        self.code = code


class _Batch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_haps", C.c_int32),
                ("read_off", C.c_void_p), ("read_bases", C.c_void_p), ("read_quals", C.c_void_p),
                ("ins_gop", C.c_void_p), ("del_gop", C.c_void_p), ("gcp", C.c_void_p),
                ("hap_off", C.c_void_p), ("hap_bases", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("cells", C.c_int64), ("fallback_pairs", C.c_int64),
                ("kernel_launches", C.c_int32), ("n_classes", C.c_int32),
                ("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
                ("sweep_ms", C.c_float), ("sweep_launches", C.c_int32), ("fp64_pairs", C.c_int64)]


EXPORTS = ["gklb_pairhmm_init", "gklb_pairhmm_compute", "gklb_pairhmm_done", "gklb_pairhmm_devices_in_use",
           "gklb_pairhmm_last_stats", "gklb_pairhmm_engines_alive", "gklb_pairhmm_acquire_engine",
           "gklb_pairhmm_release_engine", "gklb_pairhmm_compute_multi", "gklb_engine_compute_multi", "gklb_engine_device", "gklb_engine_sweep_kernel", "gklb_engine_plan_info", "gklb_engine_stage_multi", "gklb_engine_narrow", "gklb_engine_create",
           "gklb_engine_destroy", "gklb_engine_set_stream", "gklb_engine_compute", "gklb_engine_submit", "gklb_engine_wait", "gklb_engine_stage",
           "gklb_engine_stage_device", "gklb_engine_update_haps_device", "gklb_engine_run", "gklb_engine_fetch", "gklb_engine_result_device",
           "gklb_engine_synchronize", "gklb_engine_stats", "gklb_engine_time_runs", "gklb_last_error",
           "gklb_version", "gklb_device_count", "gklb_pairhmm_table"]

_lib = None


def build(experimental: bool = False) -> None:
    """Compile the library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    cmd = ["make", "-s", "-j", str(min(8, os.cpu_count() or 1)), "-C", str(PKG / "csrc")]
    if experimental:
        cmd.append("EXPERIMENTAL=1")
    subprocess.run(cmd, check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is not built: run `make -C gkl_b200/csrc` "
                                    "(there is no fallback implementation)")
        l = C.CDLL(str(LIB_PATH))
        l.gklb_last_error.restype = C.c_char_p
        l.gklb_version.restype = C.c_char_p
        l.gklb_pairhmm_table.restype = C.c_void_p
        l.gklb_pairhmm_table.argtypes = [C.c_int, C.POINTER(C.c_int)]
        l.gklb_engine_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
        for name in ("gklb_engine_destroy", "gklb_engine_run", "gklb_engine_synchronize"):
            getattr(l, name).argtypes = [C.c_void_p]
        l.gklb_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        l.gklb_engine_compute.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
        l.gklb_engine_submit.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
        l.gklb_engine_wait.argtypes = [C.c_void_p]
        l.gklb_engine_stage.argtypes = [C.c_void_p, C.POINTER(_Batch)]
        l.gklb_engine_stage_device.argtypes = [C.c_void_p, C.POINTER(_Batch)]
        l.gklb_engine_fetch.argtypes = [C.c_void_p, C.c_void_p]
        l.gklb_engine_update_haps_device.argtypes = [C.c_void_p, C.c_void_p]
        l.gklb_engine_result_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.gklb_engine_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        l.gklb_engine_time_runs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        l.gklb_pairhmm_init.argtypes = [C.c_int, C.c_int]
        l.gklb_pairhmm_compute.argtypes = [C.POINTER(_Batch), C.c_void_p]
        l.gklb_pairhmm_last_stats.argtypes = [C.POINTER(Stats)]
        l.gklb_engine_sweep_kernel.restype = C.c_char_p
        l.gklb_engine_sweep_kernel.argtypes = [C.c_void_p]
        l.gklb_engine_device.argtypes = [C.c_void_p]
        l.gklb_engine_stage_multi.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_int]
        l.gklb_engine_narrow.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.gklb_engine_plan_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.gklb_pairhmm_acquire_engine.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        l.gklb_pairhmm_release_engine.argtypes = [C.c_void_p]
        l.gklb_pairhmm_compute_multi.argtypes = [C.POINTER(_Batch), C.c_int, C.POINTER(C.c_void_p)]
        l.gklb_engine_compute_multi.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_int, C.POINTER(C.c_void_p)]
        _lib = l
    return _lib


def _check(rc: int) -> None:
    if rc != OK:
        raise GklbError(rc, lib().gklb_last_error().decode(errors="replace"))


def _ptr(a) -> int:
    """Address of a numpy array, a torch tensor (host or device), or a raw integer address."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def make_batch(b: PairHmmBatch, arenas=None, hap=None) -> _Batch:
    """C struct for a batch.  ``arenas`` (5 objects) / ``hap`` override where the byte arenas are
    read from (e.g. pinned or device tensors); offsets always come from the host arrays."""
    src = arenas if arenas is not None else (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)
    s = _Batch()
    s.n_reads, s.n_haps = b.n_reads, b.n_haps
    s.read_off = _ptr(b.read_off)
    s.read_bases, s.read_quals, s.ins_gop, s.del_gop, s.gcp = (_ptr(x) for x in src)
    s.hap_off = _ptr(b.hap_off)
    s.hap_bases = _ptr(hap if hap is not None else b.hap_bases)
    return s


def tables() -> dict:
    out = {}
    for which, (name, dt) in enumerate((("ph2pr_f", np.float32), ("mm_f", np.float32),
                                        ("ph2pr_d", np.float64), ("mm_d", np.float64))):
        n = C.c_int(0)
        p = lib().gklb_pairhmm_table(which, C.byref(n))
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n.value,)).copy()
    return out


class Engine:
    """One engine = one device + one stream.  Thin wrapper over the gklb_engine_* entry points."""

    def __init__(self, device: int = 0, use_double: bool = False):
        self._h = C.c_void_p()
        _check(lib().gklb_engine_create(C.byref(self._h), device, int(use_double)))
        self._keep = None

    def close(self) -> None:
        if self._h:
            lib().gklb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None) -> None:
        _check(lib().gklb_engine_set_stream(self._h, cuda_stream or 0))

    def compute(self, b: PairHmmBatch, out: np.ndarray | None = None, arenas=None, hap=None, out_ptr=None):
        """Host buffers in, log10 likelihoods[r * n_haps + h] out (synchronous)."""
        b.validate()
        if out is None and out_ptr is None:
            out = np.empty(b.n_reads * b.n_haps, dtype=np.float64)
        s = make_batch(b, arenas, hap)
        _check(lib().gklb_engine_compute(self._h, C.byref(s), out_ptr if out_ptr is not None else _ptr(out)))
        return out

    def compute_multi(self, batches, outs=None):
        """Several regions as one job (gklb_engine_compute_multi): returns the list of likelihood arrays."""
        arr, ptrs, outs = _multi_args(batches, outs)
        _check(lib().gklb_engine_compute_multi(self._h, arr, len(batches), ptrs))
        return outs

    def submit(self, b: PairHmmBatch, out: np.ndarray) -> None:
        """Asynchronous compute: returns once everything is queued; `out` is filled by wait()."""
        b.validate()
        s = make_batch(b)
        self._keep = (b, out)
        _check(lib().gklb_engine_submit(self._h, C.byref(s), _ptr(out)))

    def wait(self) -> None:
        _check(lib().gklb_engine_wait(self._h))

    def stage(self, b: PairHmmBatch, arenas=None, hap=None, device: bool = False) -> None:
        b.validate()
        s = make_batch(b, arenas, hap)
        self._keep = (b, arenas, hap)
        fn = lib().gklb_engine_stage_device if device else lib().gklb_engine_stage
        _check(fn(self._h, C.byref(s)))

    def update_haps_device(self, hap_dev) -> None:
        """New haplotype bases (same lengths) from a device buffer, e.g. the target of an NCCL broadcast."""
        _check(lib().gklb_engine_update_haps_device(self._h, _ptr(hap_dev)))

    def run(self) -> None:
        _check(lib().gklb_engine_run(self._h))

    def fetch(self, n: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(n if n is not None else self.stats().pairs, dtype=np.float64)
        _check(lib().gklb_engine_fetch(self._h, _ptr(out)))
        return out

    def synchronize(self) -> None:
        _check(lib().gklb_engine_synchronize(self._h))

    def result_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib().gklb_engine_result_device(self._h, C.byref(p)))
        return p.value

    def time_runs(self, iters: int) -> float:
        ms = C.c_float(0)
        _check(lib().gklb_engine_time_runs(self._h, iters, C.byref(ms)))
        return ms.value

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().gklb_engine_stats(self._h, C.byref(s)))
        return s

    def narrow(self, capacity: int):
        """(device pointer, bytes) of the last run's result as fp32 matrix + fp64 overrides (gklb_engine_narrow)."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().gklb_engine_narrow(self._h, int(capacity), C.byref(p), C.byref(n)))
        return p.value, n.value

    def plan_info(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        _check(lib().gklb_engine_plan_info(self._h, buf, len(buf)))
        return buf.value.decode()

    def stage_multi(self, batches) -> None:
        arr, _, _ = _multi_args(batches, [np.empty(0)] * len(batches))
        self._keep = (batches, arr)
        _check(lib().gklb_engine_stage_multi(self._h, arr, len(batches)))

    def sweep_kernel(self) -> str:
        """Name of the forward-sweep kernel the staged batch launches."""
        return lib().gklb_engine_sweep_kernel(self._h).decode()


def _multi_args(batches, outs):
    for b in batches:
        b.validate()
    if outs is None:
        outs = [np.empty(b.n_reads * b.n_haps, dtype=np.float64) for b in batches]
    arr = (_Batch * len(batches))(*[make_batch(b) for b in batches])
    ptrs = (C.c_void_p * len(batches))(*[_ptr(o) if o.size else 0 for o in outs])
    return arr, ptrs, outs


def device_count() -> int:
    return int(lib().gklb_device_count())


def global_init(use_double: bool = False, max_threads: int = 1) -> int:
    """The process-global surface the JNI layer uses (gklb_pairhmm_init); returns the number of devices in use."""
    _check(lib().gklb_pairhmm_init(int(use_double), int(max_threads)))
    return int(lib().gklb_pairhmm_devices_in_use())


def global_compute(b: PairHmmBatch, out: np.ndarray | None = None) -> np.ndarray:
    b.validate()
    if out is None:
        out = np.empty(b.n_reads * b.n_haps, dtype=np.float64)
    s = make_batch(b)
    _check(lib().gklb_pairhmm_compute(C.byref(s), _ptr(out)))
    return out


def global_compute_multi(batches, outs=None):
    """gklb_pairhmm_compute_multi: several regions in one call; returns the list of likelihood arrays."""
    arr, ptrs, outs = _multi_args(batches, outs)
    _check(lib().gklb_pairhmm_compute_multi(arr, len(batches), ptrs))
    return outs


def global_stats() -> Stats:
    st = Stats()
    _check(lib().gklb_pairhmm_last_stats(C.byref(st)))
    return st


def global_done() -> None:
    _check(lib().gklb_pairhmm_done())
