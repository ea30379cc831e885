"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: -m gpu."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_rows_equal, random_batch

pytestmark = pytest.mark.gpu

from kmertools_b200 import OligoComputer  # noqa: E402
from kmertools_b200._lib import NORM_CLI, NORM_COUNTS, NORM_PY  # noqa: E402

_cache = {}


def comp(k):
    if k not in _cache:
        _cache[k] = OligoComputer(k)
    return _cache[k]


OPTION_DEFAULTS = (("force_path", 0), ("packed16", 1), ("even_rank", 1), ("dense_odd", 1),
                   ("wave_persistent", 1), ("wave_smem_rank", 1), ("wave_budget_bytes", 96 << 20),
                   ("global_wave_bytes", 64 << 20), ("k7_mid", 1), ("fwd_fold", 1), ("fwd_min_len", 1024),
                   ("bucket", 1), ("bucket_log2_seg", 14), ("bucket_waves", 1), ("long_warps", 0), ("fwd_replicas", 1),
                   ("longest_first", 1), ("bucket_hist_kb", 64), ("k8_long", 1))


def check(k, bases, offsets, mins=True, norm_mode=NORM_CLI, dtype=np.float32, what="", **opts):
    oc = comp(k)
    for key, dflt in OPTION_DEFAULTS:
        oc.set_option(key, opts.get(key, dflt))
    n = len(offsets) - 1
    totals = np.zeros(n, dtype=np.uint64)
    got = oc.vectorise_packed(bases, offsets, norm_mode=norm_mode, mins=mins, dtype=dtype, totals=totals)
    want, wtot = O.vectorise_batch(bases, offsets, k, mins, norm_mode)
    assert np.array_equal(totals, wtot), f"{what}: totals differ"
    assert_rows_equal(got, want, dtype, what)
    for key, dflt in OPTION_DEFAULTS:
        oc.set_option(key, dflt)
    return got


def test_nt4_table_matches_reference_table():
    import ctypes as C
    from kmertools_b200 import _lib
    oc = comp(4)
    buf = (C.c_uint8 * 256)()
    _lib.check(oc._lib.ktb_debug_nt4_table(oc._h, buf))
    assert list(buf) == [O.nt4(b) for b in range(256)]


@pytest.mark.parametrize("fname", ["reads.fa", "reads.fq", "reads.fq.gz"])
def test_golden_files(golden, fname):
    """Config 1: k=4 canonical normalised on the repo fixture, byte-equal text after formatting."""
    seqs = [s for _, s in O.read_fastx(golden / fname)]
    bases, offsets = O.pack(seqs)
    rows = check(4, bases, offsets, dtype=np.float64, what=fname)
    assert O.format_rows(rows, True) == (golden / "expected_fa.kmers").read_bytes()
    cnt = check(4, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32)
    assert O.format_rows(cnt.astype(np.float64), False) == (golden / "expected_fa_batch_unnorm.kmers").read_bytes()


def test_python_api_golden(golden):
    """tests/test_oligo.py of the reference, unmodified logic, against pykmertools drop-in."""
    import pykmertools as kt
    oligo_gen = kt.OligoComputer(4)
    seqs = [s.decode() for _, s in O.read_fastx(golden / "reads.fq")]
    gen = [[round(x, 6) for x in line] for line in oligo_gen.vectorise_batch(seqs)]
    truth = [list(map(float, ln.split())) for ln in (golden / "expected_fa.kmers").read_text().splitlines()]
    assert gen == truth
    assert len(oligo_gen.get_header()) == 136 and len(oligo_gen.get_header(False)) == 256
    assert oligo_gen.get_header() == O.header(4, True)
    one = oligo_gen.vectorise_one(seqs[0])
    assert one == oligo_gen.vectorise_batch(seqs)[0]
    raw = oligo_gen.vectorise_one(seqs[0], True, False)
    assert abs(sum(raw) - 0.5) < 1e-12  # the reference's raw-mode quirk


def test_unit_kats():
    oc = comp(4)
    v = oc.vectorise_batch_array([b"AAAANGAGA"], True, True)
    assert v[0, 0] == 0.5
    u = oc.vectorise_batch_array([b"AAAANGAGA"], False, True)
    assert u[0, 0] == 1.0 and u.sum() == 2.0
    z = oc.vectorise_batch_array([b"ACG", b"", b"NNNNNNNN"], True, True)
    assert not z.any()
    a = oc.vectorise_batch_array([b"acgu", b"ACGT", bytes([0, 1, 2, 3]), b"ACGTRACGT"], False, True)
    assert a[0].tolist() == a[1].tolist() == a[2].tolist() and a[0, 27] == 1 and a[0].sum() == 1
    assert a[3].sum() == 2


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("mins", [True, False])
def test_short_reads_all_k(k, mins):
    rng = np.random.default_rng(100 + k)
    lengths = rng.integers(0, 300, size=700)
    bases, offsets = random_batch(rng, lengths, noise=0.02)
    if k == 8 and not mins:
        lengths = lengths[:64]
        bases, offsets = random_batch(rng, lengths, noise=0.02)
    check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k{k} counts")
    check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float32, what=f"k{k} f32")
    check(k, bases, offsets, mins=mins, norm_mode=NORM_PY, dtype=np.float64, what=f"k{k} f64 py")


@pytest.mark.parametrize("length", [150, 100, 16, 17, 31, 250, 259])
def test_short_kernel_uniform_lengths(length):
    rng = np.random.default_rng(5)
    bases, offsets = random_batch(rng, np.full(3001, length), noise=0.001)
    for k in (3, 4, 5):
        check(k, bases, offsets, dtype=np.float32, what=f"{length}bp k{k}")
        check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"{length}bp k{k} counts")
    check(5, bases, offsets, mins=False, norm_mode=NORM_PY, dtype=np.float64, what=f"{length}bp raw f64")


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_short_kernel_ragged_and_tiny_reads(k):
    """Everything short-eligible: tiny reads (< 16 bp, several per 16-base chunk), empties, N runs,
    homopolymers that push a byte counter to its 255 limit."""
    rng = np.random.default_rng(50 + k)
    lengths = np.r_[rng.integers(0, 40, size=500), rng.integers(0, 255 + k, size=1500), [254 + k] * 40, [0] * 20]
    rng.shuffle(lengths)
    bases, offsets = random_batch(rng, lengths, noise=0.02, n_runs=0.1)
    # homopolymer reads of the maximum short length: one bin reaches 255
    for i in range(0, 2000, 97):
        bases[int(offsets[i]):int(offsets[i + 1])] = ord("A")
    for mins in (True, False):
        check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"ragged k{k} counts")
        check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float32, what=f"ragged k{k} f32")
    check(k, bases, offsets, norm_mode=NORM_PY, dtype=np.float64, what=f"ragged k{k} f64")
    st = comp(k).stats()
    assert st["launches"] >= 1


def test_unaligned_base_pointer_phase():
    """Host path keeps the caller's 16-byte phase: start the batch at every offset 0..15."""
    rng = np.random.default_rng(77)
    bases, offsets = random_batch(rng, np.full(200, 150), noise=0.001)
    for shift in range(16):
        pad = np.frombuffer(b"G" * shift, dtype=np.uint8)
        b2 = np.concatenate([pad, bases])
        o2 = offsets + np.uint64(shift)
        check(5, b2, o2, dtype=np.float32, what=f"phase {shift}")


@pytest.mark.parametrize("k", [3, 4, 5, 6, 7, 8])
def test_medium_and_long_sequences(k):
    rng = np.random.default_rng(200 + k)
    lengths = np.r_[rng.integers(300, 5000, size=40), rng.integers(20000, 120000, size=6), [0, 1, k - 1, k, k + 1]]
    bases, offsets = random_batch(rng, lengths, noise=0.001, n_runs=0.5)
    check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k{k} counts")
    check(k, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"k{k} f32")
    check(k, bases, offsets, mins=False, norm_mode=NORM_CLI, dtype=np.float64, what=f"k{k} raw f64") if k < 8 else None


def test_k8_packed16_variant_with_overflow_fallback():
    """seq_kernel mode 5 (16-bit counters packed in code space): sequences with more than 65535 windows must
    come back through the second launch; a homopolymer drives one counter to its limit."""
    rng = np.random.default_rng(88)
    lengths = np.r_[rng.integers(0, 5000, size=30), [65535 + 7, 65535 + 8, 70000, 200000, 65542]]
    bases, offsets = random_batch(rng, lengths, noise=0.001, n_runs=0.3)
    bases[int(offsets[30]):int(offsets[31])] = ord("A")     # 65535 windows of AAAAAAAA: exactly fits
    bases[int(offsets[31]):int(offsets[32])] = ord("T")     # 65536 windows: must take the fallback
    check(8, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, packed16=1, what="k8 packed counts")
    check(8, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, packed16=1, what="k8 packed f32")
    check(8, bases, offsets, norm_mode=NORM_CLI, dtype=np.float64, packed16=1, what="k8 packed f64")
    # the same arithmetic in both hosts of it: long_kernel MODE_K8 (default) and seq_kernel mode 5
    for dtype, norm in ((np.uint32, NORM_COUNTS), (np.float32, NORM_CLI), (np.float32, NORM_PY)):
        for warps in (0, 4, 8):
            a = check(8, bases, offsets, norm_mode=norm, dtype=dtype, what=f"k8 long_kernel w{warps}", k8_long=1, long_warps=warps)
            b = check(8, bases, offsets, norm_mode=norm, dtype=dtype, what="k8 seq_kernel", k8_long=0)
            assert np.array_equal(a, b)


def test_k8_long_kernel_mixed_lengths():
    """MODE_K8 across step / warp boundaries, short reads that ride along, empty and sub-k sequences, Ns, unaligned
    starts; many sequences so that the look-ahead ring wraps."""
    rng = np.random.default_rng(89)
    lengths = np.r_[np.arange(0, 40), [511, 512, 513, 527, 528, 1000, 4096, 4111, 4112, 10_000, 16_384, 16_385, 33_000],
                    rng.integers(100, 3000, size=400), rng.integers(5000, 12000, size=60)]
    rng.shuffle(lengths)
    bases, offsets = random_batch(rng, lengths, noise=0.004, n_runs=0.2)
    for dtype, norm in ((np.uint32, NORM_COUNTS), (np.float32, NORM_CLI)):
        a = check(8, bases, offsets, norm_mode=norm, dtype=dtype, what="k8 mixed long_kernel", k8_long=1)
        b = check(8, bases, offsets, norm_mode=norm, dtype=dtype, what="k8 mixed seq_kernel", k8_long=0)
        assert np.array_equal(a, b)


def test_alternative_histogram_modes_agree():
    """The non-default histogram layouts stay correct: k=8 through the L2 rank table (mode 2) and k=7 through the
    code-space histogram (mode 1) instead of the in-kernel rank (mode 7) / dense middle-base index (mode 4)."""
    rng = np.random.default_rng(66)
    lengths = np.r_[rng.integers(0, 6000, size=60), [40000, 90000]]
    bases, offsets = random_batch(rng, lengths, noise=0.002, n_runs=0.3)
    check(8, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, packed16=0, what="k8 mode 7")
    check(8, bases, offsets, norm_mode=NORM_CLI, dtype=np.float64, packed16=0, what="k8 mode 7 f64")
    check(8, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, even_rank=0, what="k8 mode 2")
    check(8, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, even_rank=0, what="k8 mode 2 f32")
    check(7, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, dense_odd=0, what="k7 mode 1")
    check(7, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, dense_odd=0, what="k7 mode 1 f32")


@pytest.mark.parametrize("k", [9, 10])
def test_large_k_global_path(k):
    rng = np.random.default_rng(300 + k)
    lengths = np.r_[rng.integers(0, 3000, size=12), [50000]]
    bases, offsets = random_batch(rng, lengths, noise=0.002, n_runs=0.3)
    check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k{k} counts")
    check(k, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"k{k} f32")
    if k == 9:
        check(k, bases, offsets, norm_mode=NORM_CLI, dtype=np.float64, what=f"k{k} f64")


@pytest.mark.parametrize("persistent", [1, 2, 0])
@pytest.mark.parametrize("k,mins", [(9, True), (10, True), (11, True), (8, False)])
def test_global_path_many_waves(k, mins, persistent):
    """Histograms larger than shared memory with a wave budget of a few rows: many waves, ragged last wave,
    empty / shorter-than-k sequences inside a wave, as one cooperative launch and as the multi-launch variant."""
    rng = np.random.default_rng(700 + k)
    lengths = np.r_[rng.integers(0, 4000, size=20), [0, k - 1, k, 30000], rng.integers(500, 9000, size=9)]
    bases, offsets = random_batch(rng, lengths, noise=0.003, n_runs=0.2)
    dim = (4 ** k if not mins else (4 ** k + (4 ** (k // 2) if k % 2 == 0 else 0)) // 2)
    for rows_per_wave in (1, 3, 8):
        budget = 3 * rows_per_wave * dim * 4 + 64
        # 1: cooperative launch, rank from shared-memory tables (k <= 10); 2: rank through the L2 table; 0: multi-launch
        opts = dict(wave_persistent=min(persistent, 1), wave_smem_rank=int(persistent == 1), global_wave_bytes=budget,
                    wave_budget_bytes=budget, bucket=0)
        check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k{k} waves u32", **opts)
        check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float32, what=f"k{k} waves f32", **opts)
    check(k, bases, offsets, mins=mins, norm_mode=NORM_PY, dtype=np.float32, what=f"k{k} py f32", **opts)
    if k == 9:
        check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float64, what=f"k{k} f64", **opts)


def test_global_path_more_items_than_warps():
    """One 3 Mbp contig between short ones at k = 9: a wave holds more 512-base steps than the grid has warps, so
    warps loop over several items, and the step table spans empty and long rows."""
    rng = np.random.default_rng(77)
    lengths = np.r_[[700, 0], [3_000_000], rng.integers(5, 3000, size=6)]
    bases, offsets = random_batch(rng, lengths, noise=0.0005, n_runs=0.05)
    check(9, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k9 long contig u32")
    check(9, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what="k9 long contig f32")
    check(8, bases, offsets, mins=False, norm_mode=NORM_COUNTS, dtype=np.uint32, what="raw k8 long contig u32")


def test_global_path_degenerate_batches():
    """Nothing to count (every sequence shorter than k) and the largest supported k (one 33.5 MB row per wave)."""
    for lengths in ([0, 0, 0], [8, 0, 3]):
        bases, offsets = random_batch(np.random.default_rng(3), lengths)
        check(9, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k9 nothing to count")
        check(9, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what="k9 nothing to count f32")
    rng = np.random.default_rng(12)
    bases, offsets = random_batch(rng, [5000, 11, 12, 40000], noise=0.002)
    check(12, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k12 u32")
    check(12, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what="k12 f32")


@pytest.mark.parametrize("k", [3, 5, 7])
def test_forced_global_path_matches(k):
    rng = np.random.default_rng(400 + k)
    lengths = rng.integers(0, 2000, size=200)
    bases, offsets = random_batch(rng, lengths, noise=0.01)
    check(k, bases, offsets, dtype=np.float32, force_path=1, what="global path")
    check(k, bases, offsets, dtype=np.float32, force_path=2, what="no short kernel")


def test_mixed_short_and_long_groups():
    """Short groups, ineligible groups and long contigs interleaved (length-binned dispatch)."""
    rng = np.random.default_rng(9)
    lengths = np.r_[np.full(64, 150), [100000], np.full(40, 100), rng.integers(0, 3000, size=50), np.full(33, 259)]
    bases, offsets = random_batch(rng, lengths, noise=0.005, n_runs=0.2)
    for k in (4, 5):
        check(k, bases, offsets, dtype=np.float32, what=f"mixed k{k}")
        check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"mixed k{k}")


def test_chunked_host_pipeline_matches():
    rng = np.random.default_rng(11)
    bases, offsets = random_batch(rng, rng.integers(100, 200, size=20000), noise=0.001)
    oc = comp(5)
    oc.set_option("chunk_bytes", 1 << 20)  # ~40 chunks
    try:
        check(5, bases, offsets, dtype=np.float32, what="chunked")
        st = oc.stats()
        assert st["launches"] >= 40 and st["d2h_bytes"] >= 20000 * 512 * 4
    finally:
        oc.set_option("chunk_bytes", 512 << 20)


def test_bad_offsets_are_rejected_in_any_chunk():
    """Offsets are validated chunk by chunk (under the copies of the chunks before): a decreasing offset in a LATER
    chunk must still fail the call, with no copy of this call left in flight, and the handle stays usable."""
    from kmertools_b200._lib import KtbError, KTB_ERR_ARG
    rng = np.random.default_rng(17)
    bases, offsets = random_batch(rng, rng.integers(100, 200, size=20000), noise=0.0)
    oc = comp(5)
    oc.set_option("chunk_bytes", 1 << 20)  # ~40 chunks
    try:
        for at in (0, 7, 10001, 19999):
            bad = offsets.copy()
            if at == 19999:
                bad[at] = bad[-1] + 1        # beyond the end of the buffer
            else:
                bad[at + 1] = bad[at] - 1 if bad[at] else bad[at + 2] + 5   # decreasing
            with pytest.raises(KtbError) as ei:
                oc.vectorise_packed(bases, bad, norm_mode=NORM_CLI, mins=True, dtype=np.float32)
            assert ei.value.code == KTB_ERR_ARG and "non-decreasing" in str(ei.value)
        check(5, bases, offsets, dtype=np.float32, what="after rejected calls")
    finally:
        oc.set_option("chunk_bytes", 512 << 20)


def test_chunked_pipeline_with_rejected_groups():
    """Many small chunks whose kernels could overlap on different streams: the short-read kernel's reject list
    and the work counters are shared by the handle, so consecutive chunks must be chained."""
    rng = np.random.default_rng(13)
    lengths = np.where(rng.random(40000) < 0.1, rng.integers(300, 3000, size=40000), 150)
    bases, offsets = random_batch(rng, lengths, noise=0.001)
    oc = comp(5)
    oc.set_option("chunk_bytes", 1 << 19)  # 256 rows per chunk -> ~160 chunks
    try:
        for _ in range(3):
            check(5, bases, offsets, dtype=np.float32, what="chunked mixed")
    finally:
        oc.set_option("chunk_bytes", 512 << 20)


def test_survey_hashes_via_gpu(golden):
    """sha256 of the CLI text for k=3..7 x {canonical,raw} x {norm,counts}."""
    from tests.test_oracle import SURVEY_KATS
    seqs = [s for _, s in O.read_fastx(golden / "reads.fa")]
    bases, offsets = O.pack(seqs)
    for k, want in SURVEY_KATS.items():
        got = []
        for mins in (True, False):
            for norm in (True, False):
                rows = comp(k).vectorise_packed(bases, offsets, norm_mode=NORM_CLI if norm else NORM_COUNTS,
                                                mins=mins, dtype=np.float64)
                got.append(hashlib.sha256(O.format_rows(rows, norm)).hexdigest()[:16])
        assert tuple(got) == want, k


def test_device_entry_point_with_torch_tensors():
    import torch
    rng = np.random.default_rng(12)
    bases, offsets = random_batch(rng, np.full(4096, 150), noise=0.001)
    oc = comp(5)
    dev = torch.device("cuda:0")
    tb = torch.from_numpy(bases).to(dev)
    to = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    out = torch.empty((4096, 512), dtype=torch.float32, device=dev)
    tot = torch.zeros(4096, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    oc.vectorise_device(tb.data_ptr(), to.data_ptr(), 4096, int(offsets[-1]), out.data_ptr(),
                        d_totals=tot.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    want, wtot = O.vectorise_batch(bases, offsets, 5, True, NORM_CLI)
    assert_rows_equal(out.cpu().numpy(), want, np.float32, "device path")
    assert np.array_equal(tot.cpu().numpy().astype(np.uint64), wtot)
    # size-independent property: every row sums to 1 (or 0)
    s = out.sum(dim=1)
    assert torch.all((s - 1).abs() < 1e-4)


def test_torch_tensor_surface():
    import torch
    oc = comp(4)
    seqs = [b"ACGTTGCANNACGT" * 5, b"", b"acgtacgtacgtaaaaccc", b"GGGTGATGGCCGCTGCCGATGGCGTCAAATCCCACCAAGTTACC"]
    t = oc.vectorise_batch_tensor(seqs)
    assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (4, 136)
    bases, offsets = O.pack(seqs)
    want, _ = O.vectorise_batch(bases, offsets, 4, True, NORM_PY)
    torch.cuda.synchronize()
    assert_rows_equal(t.cpu().numpy(), want, np.float32, "tensor surface")
    c = oc.vectorise_batch_tensor(seqs, norm=False, dtype=torch.int32)
    assert np.array_equal(c.cpu().numpy().astype(np.uint32), O.vectorise_batch(bases, offsets, 4, True, 0)[0].astype(np.uint32))
