"""Parity at (a fraction of) BASELINE.json's sizes through size-independent properties + sampled exact checks.

The oracle is too slow to redo 10 M rows, so the full batch is checked by invariants computed independently
with torch ops on the GPU (number of valid windows per read from a cumulative sum of ambiguous positions,
row sums, checksum of checksums, chunking invariance) and a random sample of rows is compared bit-exactly."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_rows_equal

pytestmark = pytest.mark.gpu


def _valid_windows_torch(bases, n, L, k):
    """windows of length k without an ambiguous base, per read (uniform length L), independent of our kernels"""
    import torch
    valid = torch.zeros(256, dtype=torch.bool, device=bases.device)
    valid[torch.tensor(list(b"ACGTUacgtu\x00\x01\x02\x03"), device=bases.device, dtype=torch.long)] = True
    bad = (~valid[bases.long()]).view(n, L).to(torch.int32)
    c = torch.cumsum(bad, dim=1)
    c = torch.nn.functional.pad(c, (1, 0))
    inwin = c[:, k:] - c[:, :-k]          # ambiguous bases inside each window
    return (inwin == 0).sum(dim=1)


@pytest.mark.parametrize("workload,scale", [("reads150_k5", 0.2), ("reads10k_k7", 0.02), ("reads10k_k8", 0.05),
                                            ("reads100k_k10", 0.05)])
def test_baseline_shapes_by_properties(workload, scale):
    import torch
    import bench
    from kmertools_b200 import OligoComputer
    spec = bench.WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    bases, offsets = bench.make_workload_torch(spec, scale, dev)
    n, L, k = offsets.numel() - 1, int(spec["length"]), spec["k"]
    oc = OligoComputer(k)
    counts = oc.vectorise_tensors(bases, offsets, norm_mode=0, dtype=torch.int32)
    totals = torch.zeros(n, dtype=torch.int64, device=dev)
    oc.vectorise_tensors(bases, offsets, norm_mode=0, dtype=torch.int32, out=counts, totals=totals)
    torch.cuda.synchronize()
    # (1) every valid window is counted exactly once
    want_tot = _valid_windows_torch(bases, n, L, k)
    assert torch.equal(totals, want_tot)
    assert torch.equal(counts.sum(dim=1, dtype=torch.int64), want_tot)
    # (2) checksum of checksums
    assert int(counts.sum(dtype=torch.int64)) == int(want_tot.sum())
    # (3) normalised rows: f32 == counts / total exactly (0 ulp claim) and sum to 1
    rows = oc.vectorise_tensors(bases, offsets, norm_mode=1, dtype=torch.float32)
    ref = (counts.double() / want_tot.clamp(min=1).double().unsqueeze(1)).float()
    assert torch.equal(rows, ref)
    s = rows.sum(dim=1, dtype=torch.float64)
    assert bool(torch.all(((s - 1).abs() < 1e-4) | (want_tot == 0)))
    # (4) idempotence and chunking invariance of the host path on a slice
    m = min(n, 50_000)
    hb = bases[: m * L].cpu().numpy()
    ho = offsets[: m + 1].cpu().numpy().astype(np.uint64)
    a = oc.vectorise_packed(hb, ho, norm_mode=1, dtype=np.float32)
    oc.set_option("chunk_bytes", 8 << 20)
    b = oc.vectorise_packed(hb, ho, norm_mode=1, dtype=np.float32)
    oc.set_option("chunk_bytes", 512 << 20)
    assert np.array_equal(a, b) and np.array_equal(a, rows[:m].cpu().numpy())
    # (5) a random sample of rows, bit-exact against the oracle
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(n, size=min(n, 3000 if L <= 1000 else (300 if L <= 10_000 else 24)), replace=False))
    sb = np.concatenate([bases[int(i) * L:(int(i) + 1) * L].cpu().numpy() for i in pick[:300]])
    so = np.arange(len(pick[:300]) + 1, dtype=np.uint64) * L
    want, _ = O.vectorise_batch(sb, so, k, True, 1)
    assert_rows_equal(rows[torch.from_numpy(pick[:300]).to(dev)].cpu().numpy(), want, np.float32, workload)


def test_contigs_shape_by_properties():
    """Ragged contigs with N runs, IUPAC codes and soft-masking (BASELINE config 4) at reduced count."""
    import torch
    import bench
    from kmertools_b200 import OligoComputer
    spec = bench.WORKLOADS["contigs_k4"]
    dev = torch.device("cuda", 0)
    bases, offsets = bench.make_workload_torch(spec, 0.02, dev)   # 400 contigs, ~30 Mbases
    n = offsets.numel() - 1
    oc = OligoComputer(4)
    totals = torch.zeros(n, dtype=torch.int64, device=dev)
    counts = oc.vectorise_tensors(bases, offsets, norm_mode=0, dtype=torch.int32, totals=totals)
    torch.cuda.synchronize()
    assert torch.equal(counts.sum(dim=1, dtype=torch.int64), totals)
    hb, ho = bases.cpu().numpy(), offsets.cpu().numpy().astype(np.uint64)
    want, wt = O.vectorise_batch(hb, ho, 4, True, 0)
    assert np.array_equal(totals.cpu().numpy().astype(np.uint64), wt)
    assert_rows_equal(counts.cpu().numpy().astype(np.uint32), want, np.uint32, "contigs")
    rows = oc.vectorise_tensors(bases, offsets, norm_mode=1, dtype=torch.float32)
    want_n, _ = O.vectorise_batch(hb, ho, 4, True, 1)
    assert_rows_equal(rows.cpu().numpy(), want_n, np.float32, "contigs norm")
