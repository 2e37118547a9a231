"""GPU: parity of the CUDA path (through the C-ABI) with the oracle.

Tolerances: BASELINE.json asks for 1e-5 relative on the log10 likelihoods against GKL's AVX fp32
PairHMM (with its fp64 rerun); GKL's own tests hold 1e-5 absolute against the golden file."""
import numpy as np
import pytest

import oracle
from gkl_b200 import (HaplotypeDataHolder, IntelPairHmm, PairHMMNativeArguments, ReadDataHolder, fixtures, native,
                      synth)
from gkl_b200.pairhmm import IllegalArgumentException, NullPointerException

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def checker(b, use_double=False):
    """GKL's own compiled code when oracle/_ref travelled with the repo, else the restatement."""
    t = oracle.host_threads()
    if oracle.ref_available():
        return oracle.ref_pairhmm(b, use_double, threads=t)[0]
    return oracle.port_pairhmm(b, use_double, threads=t)[0]


def rel(a, ref):
    return np.abs(a - ref) / np.maximum(np.abs(ref), 1e-300)


@pytest.fixture(scope="module")
def eng():
    e = native.Engine(0, False)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_d():
    e = native.Engine(0, True)
    yield e
    e.close()


def test_golden_file_fp32_and_fp64(golden_pairhmm, eng, eng_d):
    batches, expected = golden_pairhmm
    got = np.array([eng.compute(b)[0] for b in batches])
    assert np.abs(got - expected).max() <= 1e-5  # PairHmmUnitTest.dataFileTest
    got_d = np.array([eng_d.compute(b)[0] for b in batches])
    assert np.abs(got_d - expected).max() <= 1e-5
    ref = np.array([checker(b)[0] for b in batches])
    assert rel(got, ref).max() <= REL_TOL


def test_simple_test_known_answer_through_operator_interface():
    hmm = IntelPairHmm()
    assert hmm.load()
    hmm.initialize(PairHMMNativeArguments(False, 1))
    q = b"++++"
    out = np.zeros(1)
    hmm.computeLikelihoods([ReadDataHolder(b"ACGT", q, q, q, q)], [HaplotypeDataHolder(b"ACGT")], out)
    assert abs(out[0] - (-6.022797e-01)) <= 1e-5  # PairHmmUnitTest.simpleTest
    with pytest.raises(NullPointerException):
        hmm.computeLikelihoods(None, [HaplotypeDataHolder(b"ACGT")], out)
    with pytest.raises(NullPointerException):
        hmm.computeLikelihoods([ReadDataHolder(b"ACGT", q, q, q, q)], None, out)
    with pytest.raises(NullPointerException):
        hmm.computeLikelihoods([ReadDataHolder(b"ACGT", q, q, q, q)], [HaplotypeDataHolder(b"ACGT")], None)
    with pytest.raises(IllegalArgumentException):
        hmm.computeLikelihoods([ReadDataHolder(b"ACGT", q, q, q, q)], [HaplotypeDataHolder(b"")], out)
    hmm.computeLikelihoods([], [HaplotypeDataHolder(b"ACGT")], out)  # empty input: no-op
    hmm.done()
    hmm.done()  # idempotent
    hmm.initialize(None)  # re-initialise after done
    hmm.computeLikelihoods([ReadDataHolder(b"ACGT", q, q, q, q)], [HaplotypeDataHolder(b"ACGT")], out)
    assert abs(out[0] - (-6.022797e-01)) <= 1e-5
    hmm.done()


def test_config1(eng):
    b = synth.config1()
    assert rel(eng.compute(b), checker(b)).max() <= REL_TOL


@pytest.mark.parametrize("seed,kw", [
    (31, {}),                                                  # ragged lengths, N and unknown bytes
    (32, dict(low_quality=0.1, unrelated=0.3)),                # quals over 0..127, many fp64 reruns
    (33, dict(read_len=(1, 12), hap_len=(1, 12))),             # tiny sequences
    (34, dict(read_len=(250, 700), hap_len=(300, 900))),       # multi-pass reads
])
def test_random_batches_fp32_path_with_fallback(eng, seed, kw):
    b = synth.random_batch(seed, 120, 30, **kw)
    out = eng.compute(b)
    ref = checker(b)
    assert np.all(np.isfinite(out) == np.isfinite(ref))
    ok = np.isfinite(ref)
    assert rel(out[ok], ref[ok]).max() <= REL_TOL
    if kw.get("unrelated"):
        assert eng.stats().fallback_pairs > 0


@pytest.mark.parametrize("seed,kw", [(41, dict(low_quality=0.1, unrelated=0.3)),
                                     (42, dict(read_len=(250, 600), hap_len=(300, 700)))])
def test_random_batches_double_precision_mode(eng_d, seed, kw):
    b = synth.random_batch(seed, 60, 20, **kw)
    out = eng_d.compute(b)
    ref = checker(b, True)
    assert rel(out, ref).max() <= 1e-9


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_small_batches(eng, eng_d, seed):
    rng = np.random.default_rng(1000 + seed)
    n_reads, n_haps = int(rng.integers(1, 41)), int(rng.integers(1, 13))
    rmax, hmax = int(rng.integers(1, 301)), int(rng.integers(1, 501))
    b = synth.random_batch(2000 + seed, n_reads, n_haps, read_len=(1, rmax), hap_len=(1, hmax),
                           low_quality=float(rng.random() * 0.3), n_frac=float(rng.random() * 0.1),
                           unrelated=float(rng.random() * 0.5))
    for e, dbl, tol in ((eng, False, REL_TOL), (eng_d, True, 1e-9)):
        out, ref = e.compute(b), checker(b, dbl)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(out), ok)
        assert rel(out[ok], ref[ok]).max() <= tol, (seed, dbl)


def test_every_length_class_boundary(eng):
    # one read at each class capacity and one past it (32, 33, 40, 41, ... 256, 257, 288, 289, 320, 321, 512, 513)
    rng = np.random.default_rng(7)
    lens = sorted({c + d for c in (32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 288, 320, 512) for d in (0, 1)} | {1, 2})
    hap = synth.ACGT[rng.integers(0, 4, size=300)]
    reads = [hap[:min(L, 300)] if L <= 300 else np.concatenate([hap, synth.ACGT[rng.integers(0, 4, size=L - 300)]])
             for L in lens]
    mk = lambda v: [bytes([v]) * len(r) for r in reads]
    b = fixtures.PairHmmBatch.from_lists([bytes(r) for r in reads], mk(30), mk(40), mk(40), mk(10),
                                         [bytes(hap), bytes(hap[:150]), bytes(hap[40:])])
    out = eng.compute(b)
    ref = checker(b)
    assert rel(out, ref).max() <= REL_TOL


def test_adversarial_gap_open_patterns_stay_within_tolerance(eng):
    """The fp32 kernel keeps the insertion state as X / pMX(row); alternating gap-open penalties of 0 and 127 with a
    zero gap-continuation penalty push that ratio past the fp32 range.  Such pairs must be caught (non-finite sum)
    and recomputed by the fp64 kernel, so the results still match the reference."""
    rng = np.random.default_rng(3)
    hap = synth.ACGT[rng.integers(0, 4, size=260)]
    reads, quals, ins, dele, gcp = [], [], [], [], []
    for k in range(24):
        L = int(rng.integers(60, 200))
        s = int(rng.integers(0, 260 - 60))
        r = hap[s:s + L]
        if len(r) < L:
            r = np.concatenate([r, synth.ACGT[rng.integers(0, 4, size=L - len(r))]])
        reads.append(bytes(r))
        quals.append(bytes(rng.integers(20, 41, size=L).astype(np.uint8)))
        pat = np.where(np.arange(L) % 2 == (k % 2), 0, 127).astype(np.uint8)
        if k % 3 == 0:
            pat = np.where(np.arange(L) < L // 2, 127, 0).astype(np.uint8)
        ins.append(bytes(pat))
        dele.append(bytes(rng.integers(0, 128, size=L).astype(np.uint8)))
        gcp.append(bytes(np.zeros(L, dtype=np.uint8) if k % 2 else rng.integers(0, 12, size=L).astype(np.uint8)))
    b = fixtures.PairHmmBatch.from_lists(reads, quals, ins, dele, gcp, [bytes(hap), bytes(hap[:130]), bytes(hap[70:])])
    out = eng.compute(b)
    ref = checker(b)
    ok = np.isfinite(ref)
    assert np.all(np.isfinite(out[ok]))
    assert rel(out[ok], ref[ok]).max() <= REL_TOL


def test_very_long_reads(eng, eng_d):
    # thousands of rows: many passes, records read from global memory (they would not fit in shared memory)
    rng = np.random.default_rng(8)
    hap = synth.ACGT[rng.integers(0, 4, size=3000)]
    reads = [bytes(hap[100:100 + L]) for L in (2500, 1031, 257)]
    mk = lambda v: [bytes([v]) * len(r) for r in reads]
    b = fixtures.PairHmmBatch.from_lists(reads, mk(35), mk(40), mk(40), mk(10), [bytes(hap), bytes(hap[50:2900])])
    assert rel(eng.compute(b), checker(b)).max() <= REL_TOL
    assert rel(eng_d.compute(b), checker(b, True)).max() <= 1e-9


def test_haplotype_panel_larger_than_shared_memory_is_tiled(eng):
    b = synth.random_batch(51, 24, 800, read_len=(60, 120), hap_len=(250, 450))
    assert rel(eng.compute(b), checker(b)).max() <= REL_TOL


def test_config3_regions(eng):
    for reg in synth.config3(4, seed=3):
        out = eng.compute(reg)
        assert rel(out, checker(reg)).max() <= REL_TOL


def test_config2_full_size_properties(eng):
    """Full BASELINE size (10 000 x 128): compared with the checker on a read sample, plus
    size-independent properties on the whole output."""
    b = synth.config2()
    out = eng.compute(b).reshape(b.n_reads, b.n_haps)
    assert np.all(np.isfinite(out)) and out.max() < 0
    idx = np.arange(0, b.n_reads, 25)
    sample = np.concatenate([checker(b.read_slice(int(r), int(r) + 1)) for r in idx[:120]])
    assert rel(out[idx[:120]].ravel(), sample).max() <= REL_TOL
    # rows are independent of batch composition: recomputing a slice gives identical bits
    again = eng.compute(b.read_slice(1000, 1200)).reshape(200, b.n_haps)
    assert np.array_equal(again, out[1000:1200])
    # permuting haplotypes permutes columns
    perm = np.random.default_rng(0).permutation(b.n_haps)
    haps = [bytes(b.hap_bases[b.hap_off[h]:b.hap_off[h + 1]]) for h in perm]
    off = np.zeros(b.n_haps + 1, dtype=np.int64)
    np.cumsum([len(h) for h in haps], out=off[1:])
    sl = b.read_slice(0, 300)
    pb = fixtures.PairHmmBatch(sl.read_off, sl.read_bases, sl.read_quals, sl.ins_gop, sl.del_gop, sl.gcp, off,
                               np.frombuffer(b"".join(haps), dtype=np.uint8).copy())
    assert np.array_equal(eng.compute(pb).reshape(300, b.n_haps), out[:300][:, perm])


def test_staged_run_fetch_equals_compute(eng):
    b = synth.config2(300, 64)
    a = eng.compute(b)
    eng.stage(b)
    eng.run()
    eng.run()  # re-running a staged batch is allowed
    assert np.array_equal(eng.fetch(b.n_reads * b.n_haps), a)


def test_submit_wait_equals_compute(eng):
    b1, b2 = synth.config2(300, 40, seed=11), synth.config2(200, 24, seed=12)
    want1, want2 = eng.compute(b1), eng.compute(b2)
    other = native.Engine(0, False)
    o1, o2 = np.zeros(b1.n_reads * b1.n_haps), np.zeros(b2.n_reads * b2.n_haps)
    eng.submit(b1, o1)      # two engines in flight at once, as the JNI layer's read-block pipeline uses them
    other.submit(b2, o2)
    with pytest.raises(native.GklbError):
        eng.submit(b2, o2)  # one batch in flight per engine
    for call in (lambda: eng.compute(b2), lambda: eng.stage(b2), lambda: eng.fetch(b1.n_reads * b1.n_haps)):
        with pytest.raises(native.GklbError) as ei:   # ... and nothing may restage or fetch under it
            call()
        assert ei.value.code == native.ERR_STATE
    eng.wait()
    other.wait()
    other.close()
    assert np.array_equal(o1, want1) and np.array_equal(o2, want2)


def test_device_resident_inputs(eng):
    import torch
    b = synth.config2(200, 32)
    dev = [torch.from_numpy(x).cuda() for x in (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)]
    hap = torch.from_numpy(b.hap_bases).cuda()
    eng.stage(b, arenas=dev, hap=hap, device=True)
    eng.run()
    assert np.array_equal(eng.fetch(b.n_reads * b.n_haps), eng.compute(b))


def test_update_haps_device_rewrites_the_panel(eng):
    import torch
    a = synth.config2(100, 24, seed=5)
    rng = np.random.default_rng(9)
    other = a.hap_bases.copy()
    flip = rng.random(len(other)) < 0.05
    other[flip] = synth.ACGT[rng.integers(0, 4, size=int(flip.sum()))]
    b = fixtures.PairHmmBatch(a.read_off, a.read_bases, a.read_quals, a.ins_gop, a.del_gop, a.gcp, a.hap_off, other)
    eng.stage(a)
    eng.update_haps_device(torch.from_numpy(other).cuda())
    eng.run()
    assert np.array_equal(eng.fetch(a.n_reads * a.n_haps), eng.compute(b))


def test_invalid_batches_are_rejected(eng):
    b = synth.config2(4, 4)
    bad = fixtures.PairHmmBatch(b.read_off.copy(), b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp,
                                b.hap_off.copy(), b.hap_bases)
    bad.read_off[2] = bad.read_off[1]  # empty read
    with pytest.raises(native.GklbError) as ei:
        eng.compute(bad)
    assert ei.value.code == native.ERR_INVALID
    assert rel(eng.compute(b), checker(b)).max() <= REL_TOL  # the engine stays usable


def test_in_process_multi_device_sharding(monkeypatch):
    """GKLB_DEVICES: the global (JNI-facing) surface shards big batches over several engines in one process.
    Two engines on device 0 exercise the same code path as two GPUs."""
    import torch
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    monkeypatch.setenv("GKLB_DEVICES", devs)
    assert native.global_init(False, 1) == 2
    try:
        b = synth.config2(2400, 64)  # 4.6e9 cells... above the per-device threshold only when split in two? use a bigger hap count
        big = synth.config2(3000, 128)  # 1.1e10 cells -> two shards
        out = native.global_compute(big)
        st = native.global_stats()
        assert st.pairs == big.n_reads * big.n_haps and st.cells == big.cells()
        ref = checker(big.read_slice(0, 40))
        assert rel(out[:40 * big.n_haps], ref).max() <= REL_TOL
        ref = checker(big.read_slice(2960, 3000))
        assert rel(out[2960 * big.n_haps:], ref).max() <= REL_TOL
        one = native.Engine(0, False)
        assert np.array_equal(one.compute(big), out)  # sharding does not change a single bit
        one.close()
        small = native.global_compute(b.read_slice(0, 50))  # small batches stay on one device
        assert native.global_stats().pairs == 50 * 64 and np.all(np.isfinite(small))
    finally:
        native.global_done()


def test_multi_region_call_is_bit_identical_to_one_call_per_region(eng):
    """gklb_engine_compute_multi: several active regions share the launches (one task queue per launch group);
    every region's matrix must equal what a call of its own returns, bit for bit, and match the reference."""
    regions = synth.config3(6, seed=11) + [synth.random_batch(61, 40, 9, low_quality=0.1, unrelated=0.4),
                                           synth.random_batch(62, 7, 3, read_len=(1, 12), hap_len=(1, 12)),
                                           synth.random_batch(63, 12, 5, read_len=(250, 600), hap_len=(300, 700))]
    single = [eng.compute(r) for r in regions]
    multi = eng.compute_multi(regions)
    for r, a, m in zip(regions, single, multi):
        assert np.array_equal(a, m)
        ref = checker(r)
        ok = np.isfinite(ref)
        assert rel(m[ok], ref[ok]).max() <= REL_TOL
    assert eng.stats().pairs == sum(r.n_reads * r.n_haps for r in regions)
    # the engine stays usable for single calls, and empty regions keep their slot
    empty = fixtures.PairHmmBatch.from_lists([], [], [], [], [], [b"ACGT"])
    outs = eng.compute_multi([regions[0], empty, regions[1]])
    assert np.array_equal(outs[0], single[0]) and outs[1].size == 0 and np.array_equal(outs[2], single[1])


def test_multi_region_call_through_the_global_surface():
    native.global_init(False, 1)
    try:
        regions = synth.config3(40, seed=12)   # more regions than fit one launch group
        outs = native.global_compute_multi(regions)
        one = native.Engine(0, False)
        for r, o in zip(regions, outs):
            assert np.array_equal(one.compute(r), o)
        one.close()
        st = native.global_stats()
        assert st.pairs == sum(r.n_reads * r.n_haps for r in regions)
    finally:
        native.global_done()


def test_multi_region_call_spreads_jobs_over_the_configured_devices(monkeypatch):
    """gklb_pairhmm_compute_multi with GKLB_DEVICES: the regions are balanced into one job per device (each on an
    engine of its own) once there is enough work; a region big enough to shard goes through gklb_pairhmm_compute.
    Two engines on device 0 exercise the same code path as two GPUs."""
    import torch
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    monkeypatch.setenv("GKLB_DEVICES", devs)
    assert native.global_init(False, 1) == 2
    try:
        regions = synth.config3(20, seed=13)                 # 5e9 cells -> two jobs
        empty = fixtures.PairHmmBatch.from_lists([], [], [], [], [], [b"ACGT"])
        big = synth.config2(3000, 128)                       # shards on its own
        job = regions[:7] + [empty] + regions[7:] + [big]
        outs = native.global_compute_multi(job)
        st = native.global_stats()
        assert st.pairs == sum(r.n_reads * r.n_haps for r in job)
        assert st.cells == sum(r.cells() for r in job)
        one = native.Engine(0, False)
        for r, o in zip(job, outs):
            assert np.array_equal(one.compute(r), o) if r.n_reads else o.size == 0
        one.close()
    finally:
        native.global_done()


def test_quality_bytes_above_127_are_masked_like_the_reference(eng, eng_d):
    """avx-pairhmm-template.h:134-136,149: every quality byte is used `& 127`; bytes with the high bit set must give
    the same likelihoods as their low seven bits (checked against GKL's own code on the raw bytes)."""
    b = synth.random_batch(71, 60, 12, low_quality=0.05)
    rng = np.random.default_rng(72)
    for a in (b.read_quals, b.ins_gop, b.del_gop, b.gcp):
        hi = rng.random(a.size) < 0.3
        a[hi] |= 0x80
    for e, dbl, tol in ((eng, False, REL_TOL), (eng_d, True, 1e-9)):
        out, ref = e.compute(b), checker(b, dbl)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(out), ok)
        assert rel(out[ok], ref[ok]).max() <= tol


def test_config4_shape_at_reduced_size_all_pairs(eng):
    """BASELINE configs[3] (150-base reads x 256 haplotypes) at 4 000 reads: every pair against the reference."""
    b = synth.config4(4000, 256)
    out = eng.compute(b)
    ref = checker(b)
    assert np.array_equal(np.isfinite(out), np.isfinite(ref))
    assert rel(out, ref).max() <= REL_TOL
    assert eng.stats().fallback_pairs > 0


def test_nccl_sharding_is_bit_identical_to_direct_sharding(monkeypatch):
    """GKLB_SHARD=nccl (engine_nccl.cu): panel broadcast over NVLink, fp32 slabs + fp64 overrides gathered to GPU 0.
    Needs two GPUs (one NCCL rank per GPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    big = synth.config2(3000, 128)  # 1.1e10 cells -> two shards
    monkeypatch.setenv("GKLB_DEVICES", "0,1")
    outs = {}
    for mode in ("direct", "nccl"):
        monkeypatch.setenv("GKLB_SHARD", mode)
        assert native.global_init(False, 1) == 2
        try:
            outs[mode] = native.global_compute(big)
            st = native.global_stats()
            assert st.pairs == big.n_reads * big.n_haps and st.fallback_pairs > 0
        finally:
            native.global_done()
    assert np.array_equal(outs["direct"], outs["nccl"])
    ref = checker(big.read_slice(0, 64))
    assert rel(outs["nccl"][:64 * big.n_haps], ref).max() <= REL_TOL


def test_global_surface_lifecycle_is_reference_counted():
    """GKL's doneNative is empty and initNative only sets flags (IntelPairHmm.cc:55-118,189-192): several IntelPairHmm
    instances initialise and close independently.  Here init/done are reference counted over a pool of engines."""
    lib = native.lib()
    b = synth.config2(40, 8)
    ref = native.Engine(0, False)
    want = ref.compute(b)
    ref.close()
    assert native.global_init(False, 1) == 1          # instance A
    assert native.global_init(False, 1) == 1          # instance B: cheap, engines are kept
    alive = lib.gklb_pairhmm_engines_alive()
    assert alive >= 1
    native.global_done()                              # A closes
    assert lib.gklb_pairhmm_engines_alive() == alive  # ... B still holds a reference
    assert np.array_equal(native.global_compute(b), want)
    native.global_done()                              # B closes: the idle engines are freed
    assert lib.gklb_pairhmm_engines_alive() == 0
    native.global_done()                              # idempotent
    assert np.array_equal(native.global_compute(b), want)  # a late compute creates an engine again
    native.global_done()
    assert lib.gklb_pairhmm_engines_alive() == 0
    # precision follows the last initialize, as GKL's g_use_double does
    native.global_init(True, 1)
    d = native.Engine(0, True)
    assert np.array_equal(native.global_compute(b), d.compute(b))
    d.close()
    native.global_done()


def test_narrowed_result_reconstructs_the_likelihoods_exactly(eng):
    """gklb_engine_narrow: fp32 matrix + (pair index, fp64) overrides of the rerun pairs must give back every double
    bit for bit (an unflagged pair's value is an fp32 widened to double, IntelPairHmm.cc:164)."""
    import torch
    b = synth.config2(400, 64)
    want = eng.compute(b)
    eng.stage(b)
    eng.run()
    n = b.n_reads * b.n_haps
    cap = 8192
    ptr, nbytes = eng.narrow(cap)
    eng.synchronize()

    class _Dev:
        def __init__(self, p, k):
            self.__cuda_array_interface__ = {"shape": (k,), "typestr": "|u1", "data": (p, False), "version": 3}
    raw = torch.as_tensor(_Dev(ptr, nbytes), device="cuda").cpu().numpy()
    f32 = raw[:4 * n].view(np.float32)
    count = int(raw[4 * n:4 * n + 4].view(np.uint32)[0])
    assert 0 < count <= cap
    idx = raw[4 * n + 8:4 * n + 8 + 4 * cap].view(np.uint32)[:count]
    off_val = (4 * n + 8 + 4 * cap + 7) // 8 * 8
    val = raw[off_val:off_val + 8 * cap].view(np.float64)[:count]
    got = f32.astype(np.float64)
    got[idx] = val
    assert np.array_equal(got, want)
    assert count == int((want < -64.0).sum()) or abs(count - int((want < -64.0).sum())) < 50


def test_big_single_device_call_is_pipelined_in_pieces_and_bit_identical():
    """gklb_pairhmm_compute cuts a shard of more than ~3e11 cells into pieces that two engines on the device take in
    turn (the copies of one piece run under the kernels of the other); nothing but the schedule may change."""
    native.global_init(False, 1)
    try:
        big = synth.config2(42000, 256)   # 3.2e11 cells -> two pieces
        assert big.cells() > 3.0e11
        out = native.global_compute(big)
        st = native.global_stats()
        assert st.pairs == big.n_reads * big.n_haps and st.cells == big.cells()
        one = native.Engine(0, False)
        assert np.array_equal(one.compute(big), out)
        assert one.stats().kernel_launches < st.kernel_launches   # the pieces launched separately
        one.close()
    finally:
        native.global_done()


def test_concurrent_big_calls_share_the_engine_pool_without_deadlock():
    """Three threads, each with a call big enough to want two engines of the device's pool of four: the second engine is
    only taken when free, the first one is waited for; every caller gets the single-engine result."""
    import threading
    native.global_init(False, 1)
    try:
        big = synth.config2(28000, 256)   # 2.1e11 cells -> two pieces
        one = native.Engine(0, False)
        want = one.compute(big)
        one.close()
        outs, errs = [None] * 3, []

        def work(i):
            try:
                outs[i] = native.global_compute(big)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert not any(t.is_alive() for t in th) and not errs
        for o in outs:
            assert np.array_equal(o, want)
    finally:
        native.global_done()


@pytest.mark.parametrize("read_len", [30, 36, 40, 66, 70, 78, 130, 142, 158, 280, 301])
def test_single_class_launches_of_the_twelve_warp_kernels(eng, read_len):
    """Uniform read lengths put a whole batch into one class; classes of up to 10 rows per lane then run as
    k_h2_tasks<G, K, 12> (12 warps per SM) instead of the multi-class kernel the mixed-length tests exercise."""
    b = synth.config2(600, 24, read_len)
    out = eng.compute(b)
    name = eng.sweep_kernel()
    assert name.startswith("k_h2_tasks<") and name.endswith(",12>"), name
    ref = checker(b)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(out), ok)
    assert rel(out[ok], ref[ok]).max() <= REL_TOL


def test_reads_of_257_to_320_rows_take_the_single_pass_classes(eng, eng_d):
    """2 x 300 sequencing: reads of 257..320 rows run as 32 lanes x 9 / 10 rows in one pass, in the fp32 sweep and in
    the fp64 kernels (rerun and useDoublePrecision); mixed with shorter and longer reads they share the multi-class
    launches with the other classes and with the multi-pass class."""
    b = synth.random_batch(91, 90, 14, read_len=(240, 340), hap_len=(260, 420), low_quality=0.05, unrelated=0.3)
    for e, dbl, tol in ((eng, False, REL_TOL), (eng_d, True, 1e-9)):
        out, ref = e.compute(b), checker(b, dbl)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(out), ok)
        assert rel(out[ok], ref[ok]).max() <= tol
    assert eng.stats().fallback_pairs > 0
