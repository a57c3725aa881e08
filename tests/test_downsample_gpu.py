"""Device-side DownsampledBatch (pmt_downsample_* through the C-ABI) against a numpy restatement of the same
counter-based decisions, plus the reference's invariants (batch.py:383-439) and quirk Q1."""
import numpy as np
import pytest
import torch

from permutect_b200.data.batch import Batch, DownsampledBatch
from permutect_b200.synthetic import make_wgs_arrays

pytestmark = pytest.mark.gpu
M64 = (1 << 64) - 1


def _hash_uniform(seed, rows):
    z = (seed + 0x9E3779B97F4A7C15 * (rows.astype(object) + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return np.array([(int(v) >> 40) / 16777216.0 for v in z], dtype=np.float32)


def test_downsampled_batch_matches_host_restatement_and_invariants():
    ia, fa, reads = make_wgs_arrays(500, seed=11)
    dev = torch.device("cuda:0")
    parent = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    rng = np.random.default_rng(0)
    rf = torch.from_numpy(rng.uniform(0.2, 1.0, 500).astype(np.float32))
    af = torch.from_numpy(rng.uniform(0.2, 1.0, 500).astype(np.float32))
    seed = 123456789
    import random
    random.seed(5)
    ds = DownsampledBatch(parent, rf, af, seed=seed)
    random.seed(5)
    random_int = random.randint(0, 100)
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = int(ref_off[-1])
    keep_ref = _hash_uniform(seed, np.arange(total_ref)) < np.repeat(rf.numpy(), ref_c)
    keep_alt = _hash_uniform(seed, total_ref + np.arange(int(alt_off[-1]))) < np.repeat(af.numpy(), alt_c)
    keep_alt[alt_off[1:] - (random_int % alt_c) - 1] = True                       # batch.py:418-425
    want_ref = np.add.reduceat(keep_ref.astype(np.int64), ref_off[:-1]) * (ref_c > 0)
    want_alt = np.add.reduceat(keep_alt.astype(np.int64), alt_off[:-1])
    got_ref, got_alt = (t.cpu().numpy() for t in ds.counts())
    np.testing.assert_array_equal(got_ref, want_ref)
    np.testing.assert_array_equal(got_alt, want_alt)
    assert (got_alt >= 1).all() and (got_ref <= ref_c).all() and (got_alt <= alt_c).all()
    n_kept = int(want_ref.sum() + want_alt.sum())
    idx = ds.read_indices[:n_kept].cpu().numpy()
    want_idx = np.concatenate((np.nonzero(keep_ref)[0], np.nonzero(keep_alt)[0]))   # quirk Q1: alt entries un-offset
    np.testing.assert_array_equal(idx, want_idx)
    assert idx[int(want_ref.sum()):].max() < alt_off[-1]
    fixed = DownsampledBatch(parent, rf, af, seed=seed, offset_alt_rows=True)
    # the corrected variant (explicit opt-in) points the alt entries past the ref block
    random.seed(5)
    fixed = DownsampledBatch(parent, rf, af, seed=seed, offset_alt_rows=True)
    np.testing.assert_array_equal(fixed.read_indices[:n_kept].cpu().numpy()[int(want_ref.sum()):],
                                  total_ref + np.nonzero(keep_alt)[0])


def test_downsampled_forward_equals_explicit_indices():
    """The model must give the same result for a device-built DownsampledBatch and for the same indices passed explicitly."""
    from golden_utils import load
    from helpers import model_from_golden
    from permutect_b200.utils.enums import Epoch
    g = load("v040_seed0_b64")
    dev = torch.device("cuda:0")
    model = model_from_golden(g, dev)
    model.set_epoch_type(Epoch.VALID)
    ia, fa, reads = make_wgs_arrays(300, seed=3)
    parent = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    rf, af = torch.full((300,), 0.6), torch.full((300,), 0.7)
    ds = DownsampledBatch(parent, rf, af, seed=99)
    ref_c, alt_c = ds.counts()
    n = int(ref_c.sum() + alt_c.sum())
    explicit = DownsampledBatch(parent, read_indices=ds.read_indices[:n].clone(), ref_counts=ref_c, alt_counts=alt_c)
    with torch.inference_mode():
        a = model.compute_batch_output(ds).logits_b
        b = model.compute_batch_output(explicit).logits_b
    assert torch.equal(a, b)


def test_index_builder_against_the_reference_with_a_forced_keep_mask(monkeypatch):
    """The UNMODIFIED reference's DownsampledBatch (batch.py:383-439) builds counts and gather indices from Bernoulli masks;
    torch's generator cannot be reproduced on the device, so its ``bernoulli_`` is made to return the device kernel's own
    counter-hash decisions and everything downstream -- the forced alt read, the segment sums, nonzero + hstack (quirk Q1
    included) -- is the reference's code.  Counts and indices must be identical."""
    import random
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    reference.load()
    import permutect.data.batch as rb
    import permutect.data.datum as rd
    n = 400
    ia, fa, reads = make_wgs_arrays(n, seed=17)
    ia[::7, 0] = 0                                  # some empty ref sets
    keep_rows = np.ones(len(reads), dtype=bool)
    ref_c, alt_c = make_wgs_arrays(n, seed=17)[0][:, 0].astype(int), ia[:, 1].astype(int)
    ref_off = np.concatenate(([0], np.cumsum(ref_c)))
    for v in range(0, n, 7):
        keep_rows[ref_off[v]:ref_off[v + 1]] = False
    reads = reads[keep_rows]
    ref_c = ia[:, 0].astype(int)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = int(ref_off[-1])
    dev = torch.device("cuda:0")
    parent = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    ref_parent = rb.Batch([rd.Datum(ia[v], fa[v], np.vstack((reads[ref_off[v]:ref_off[v + 1]],
                                                              reads[total_ref + alt_off[v]:total_ref + alt_off[v + 1]])), compressed=True)
                           for v in range(n)])
    rng = np.random.default_rng(3)
    rf = torch.from_numpy(rng.uniform(0.1, 1.0, n).astype(np.float32))
    af = torch.from_numpy(rng.uniform(0.1, 1.0, n).astype(np.float32))
    seed = 987654321
    calls = []

    def forced_bernoulli_(self, p):
        first_row = 0 if not calls else total_ref                      # the ref mask is drawn first, then the alt mask
        calls.append(len(self))
        u = _hash_uniform(seed, first_row + np.arange(len(self)))
        self.copy_(torch.from_numpy((u < p.cpu().numpy()).astype(np.int64)))
        return self

    monkeypatch.setattr(torch.Tensor, "bernoulli_", forced_bernoulli_)
    random.seed(31)
    want = rb.DownsampledBatch(ref_parent, ref_fracs_b=rf, alt_fracs_b=af)
    monkeypatch.undo()
    assert calls == [total_ref, int(alt_off[-1])]
    random.seed(31)
    got = DownsampledBatch(parent, rf, af, seed=seed)
    got_ref, got_alt = (t.cpu() for t in got.counts())
    assert torch.equal(got_ref.int(), want.ref_counts) and torch.equal(got_alt.int(), want.alt_counts)
    n_kept = int(want.ref_counts.sum() + want.alt_counts.sum())
    assert len(want.read_indices) == n_kept
    assert torch.equal(got.read_indices[:n_kept].cpu(), want.read_indices)
    # and the rows the model will read are the reference's (decoded on the device vs the reference's host decode)
    np.testing.assert_array_equal(got.get_reads_re().cpu().numpy(), want.get_reads_re().numpy())
