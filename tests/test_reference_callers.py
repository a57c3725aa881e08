"""INTEGRATION.md Route A, executed: the REFERENCE's own callers run against the drop-in classes.

The unmodified reference comes from ``oracle/_ref`` (``oracle/build_ref.py``).  Only the imports INTEGRATION.md names are
swapped (``DownsampledBatch`` in ``permutect.training.model_training``; the model object and the batches handed in are this
package's); ``Balancer``, ``Downsampler``, ``LossRecorder``, ``Checkpoint``, ``EvaluationMetrics``, ``backpropagate``,
``prefetch_generator`` and ``Datum`` are the reference's.  Reference lines exercised: training/model_training.py:133-201
(train_one_epoch), :203-271 (collect_evaluation_data), tools/filter_variants.py:292-320 (generate_posterior_data),
training/balancer.py:57-117, training/downsampler.py:104-125, training/loss_recorder.py:14-22,
metrics/evaluation_metrics.py:50-66.
"""
import numpy as np
import pytest
import torch

from oracle import reference

pytestmark = pytest.mark.skipif(not reference.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def _ref():
    reference.load()
    import permutect.data.batch as rb
    import permutect.data.datum as rd
    return rb, rd


def _batches(n_batches, size, seed0, sources=1):
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    out = []
    for i in range(n_batches):
        ia, fa, reads = make_wgs_arrays(size, seed=seed0 + i)
        rng = np.random.default_rng(seed0 + 100 + i)
        ia[:, 4] = rng.integers(0, sources, size)
        ia[:, 5] = ia[:, 0] + ia[:, 1] + rng.integers(0, 40, size)        # original depth / alt count
        ia[:, 6] = ia[:, 1] + rng.integers(0, 5, size)
        ia[:, 9] = 20
        ia[:, 10:16] = rng.integers(-32768, 32767, (size, 6))
        out.append(Batch.from_arrays(ia, fa, reads))
    return out


def _reference_batch(batch, rb, rd):
    """The same variants as a reference Batch (collated from reference Datum objects)."""
    ia, fa, reads = batch.int_tensor.numpy(), batch.float_tensor.numpy(), batch.reads.numpy()
    ref_c, alt_c = ia[:, 0].astype(int), ia[:, 1].astype(int)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = ref_off[-1]
    data = []
    for v in range(len(ia)):
        rows = np.vstack((reads[ref_off[v]:ref_off[v + 1]], reads[total_ref + alt_off[v]:total_ref + alt_off[v + 1]]))
        data.append(rd.Datum(ia[v], fa[v], rows, compressed=True))
    return rb.Batch(data)


# ---- CPU: the batch surface the reference's callers use, against the reference's own Batch ----------------------------
def test_batch_indices_match_reference_batch():
    rb, rd = _ref()
    from permutect_b200.data.batch import BatchIndexedTensor
    batch = _batches(1, 257, seed0=11, sources=3)[0]
    want = _reference_batch(batch, rb, rd)
    for original in (False, True):
        a, b = batch.batch_indices(original), want.batch_indices(original)
        for name in ("sources", "labels", "var_types", "ref_count_bins", "alt_count_bins", "flattened_idx"):
            assert torch.equal(getattr(a, name), getattr(b, name).long()), (original, name)
        assert batch.batch_indices(original) is a              # cached (batch.py:69-70)
    idx, ridx = batch.batch_indices(), want.batch_indices()
    g = torch.Generator().manual_seed(0)
    logits = 30 * torch.rand(257, generator=g) - 15
    labels = torch.randint(0, 2, (257,), generator=g)
    sources = torch.randint(0, 3, (257,), generator=g)
    for kw in (dict(), dict(labels=labels), dict(sources=sources), dict(labels=labels, sources=sources)):
        mine, theirs = BatchIndexedTensor.zeros(3, device="cpu"), rb.BatchIndexedTensor.zeros(3, device="cpu")
        values = torch.rand(257, generator=g)
        idx.increment_tensor(mine, values, **kw)
        ridx.increment_tensor(theirs, values, **kw)
        assert torch.equal(torch.Tensor(mine), torch.Tensor(theirs))
        plain = lambda t: t.as_subclass(torch.Tensor)
        assert torch.equal(plain(idx.index_into_tensor(mine, **kw)), plain(ridx.index_into_tensor(theirs, **kw)))
        # the two implementations' objects are interchangeable (the reference's Balancer holds ITS tensors, our indices)
        assert torch.equal(plain(idx.index_into_tensor(theirs, **kw)), plain(ridx.index_into_tensor(mine, **kw)))
    mine6, theirs6 = BatchIndexedTensor.zeros(3, include_logits=True, device="cpu"), rb.BatchIndexedTensor.zeros(3, include_logits=True, device="cpu")
    mine6.record(batch, torch.ones(257), logits=logits, use_original_counts=True)
    theirs6.record(want, torch.ones(257), logits=logits, use_original_counts=True)
    assert torch.equal(torch.Tensor(mine6), torch.Tensor(theirs6))
    for props in ((rb.BatchProperty.LABEL,), (rb.BatchProperty.SOURCE, rb.BatchProperty.ALT_COUNT_BIN), (rb.BatchProperty.LOGIT_BIN,)):
        assert torch.equal(mine6.get_marginal(*props).as_subclass(torch.Tensor), theirs6.get_marginal(*props).as_subclass(torch.Tensor))
    with pytest.raises(AssertionError):
        idx.index_into_tensor(mine6)                           # logits required iff the tensor has a logit axis


def test_batch_accessors_match_reference_batch():
    rb, rd = _ref()
    from permutect_b200.data.batch import BatchProperty
    from permutect_b200.data.datum import Data
    batch = _batches(1, 64, seed0=5, sources=2)[0]
    want = _reference_batch(batch, rb, rd)
    for mine, theirs in zip(Data, rd.Data):
        assert mine.name == theirs.name and mine.idx == theirs.idx and np.dtype(mine.dtype) == np.dtype(theirs.dtype)
    for field in rd.Data:                                       # the reference's own column descriptors are accepted
        if np.dtype(field.dtype) == np.uint32:
            got = batch.get(field)                              # (the reference's vector accessor for these is broken: int(tensor))
            assert got[3] == rd.uint32_from_two_int16s(batch.int_tensor[3, field.idx], batch.int_tensor[3, field.idx + 1])
            continue
        a, b = batch.get(field), want.get(field)
        assert a.dtype == b.dtype and torch.equal(a.nan_to_num(7.0), b.nan_to_num(7.0)), field
    assert np.array_equal(batch.get_int_array_be(), want.get_int_array_be())
    assert np.array_equal(batch.get_float_array_be(), want.get_float_array_be(), equal_nan=True)
    assert torch.equal(batch.get_training_labels(), want.get_training_labels())
    assert torch.equal(batch.get_is_labeled_mask(), want.get_is_labeled_mask())
    assert [(p.name, int(p), p.names_list) for p in BatchProperty] == [(p.name, int(p), p.names_list) for p in rb.BatchProperty]


def test_count_binning_matches_reference():
    _ref()
    import permutect.data.count_binning as rc

    from permutect_b200.data import count_binning as mc
    for name in ("MAX_REF_COUNT", "MIN_ALT_COUNT", "MAX_ALT_COUNT", "MIN_LOGIT", "MAX_LOGIT", "COUNT_BIN_SKIP", "NUM_REF_COUNT_BINS",
                 "NUM_ALT_COUNT_BINS", "NUM_LOGIT_BINS", "ALT_COUNT_BIN_BOUNDS", "REF_COUNT_BIN_BOUNDS"):
        assert getattr(mc, name) == getattr(rc, name), name
    counts = torch.arange(0, 40)
    logits = torch.linspace(-25, 25, 401)
    assert torch.equal(mc.ref_count_bin_indices(counts), rc.ref_count_bin_indices(counts))
    assert torch.equal(mc.alt_count_bin_indices(counts + 1), rc.alt_count_bin_indices(counts + 1))
    assert torch.equal(mc.logit_bin_indices(logits), rc.logit_bin_indices(logits))
    for c in range(1, 16):
        assert mc.round_alt_count_to_bin_center(c) == rc.round_alt_count_to_bin_center(c)
        assert mc.alt_count_bin_name(mc.alt_count_bin_index(c)) == rc.alt_count_bin_name(rc.alt_count_bin_index(c))
    for b in range(mc.NUM_LOGIT_BINS):
        assert mc.logit_bin_name(b) == rc.logit_bin_name(b) and mc.top_of_logit_bin(b) == rc.top_of_logit_bin(b)


# ---- GPU: the reference's loops, end to end, on the drop-in ---------------------------------------------------------------
class _SummaryWriter:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step=None):
        self.scalars[tag] = value

    def add_figure(self, *a, **k): ...
    def add_text(self, *a, **k): ...


@pytest.mark.gpu
def test_reference_train_one_epoch_and_evaluation_run_on_the_drop_in(monkeypatch):
    rb, rd = _ref()
    import permutect.training.model_training as rmt
    from permutect.training.balancer import Balancer
    from permutect.training.checkpoint import Checkpoint
    from permutect.training.downsampler import Downsampler
    from permutect.utils.enums import Epoch as RefEpoch

    import bench
    from permutect_b200.data.batch import DownsampledBatch
    monkeypatch.setattr(rmt, "DownsampledBatch", DownsampledBatch)          # the import INTEGRATION.md Route A swaps
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    num_sources = 2
    model = bench.make_model(dev)
    model.reset_source_predictor(num_sources)
    model.source_predictor.set_adversarial_strength(0.3)
    train, valid = _batches(3, 64, seed0=21, sources=num_sources), _batches(2, 64, seed0=41, sources=num_sources)
    balancer = Balancer(num_sources=num_sources, device=dev).to(device=dev, dtype=torch.float32)
    downsampler = Downsampler(num_sources=num_sources).to(device=dev, dtype=torch.float32)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)            # model_training.py:68-72
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=0.2, patience=5, threshold=0.001, min_lr=1e-5)
    checkpoint = Checkpoint(dev, model, optimizer)
    writer = _SummaryWriter()
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    common = dict(balancer=balancer, checkpoint=checkpoint, device=dev, downsampler=downsampler, epochs_per_evaluation=5,
                  last_epoch=7, model=model, num_sources=num_sources, summary_writer=writer, train_loader=train,
                  train_optimizer=optimizer, train_scheduler=scheduler, valid_loader=valid)
    rmt.train_one_epoch(epoch=1, epoch_type=RefEpoch.TRAIN, is_calibration_epoch=False, **common)
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    changed = [k for k in before if not torch.equal(before[k], after[k])]
    assert len(changed) > 150, len(changed)                                   # every trainable tensor moved
    assert float(balancer.counts_slvra.sum()) == 2 * 3 * 64                   # two draws of three parent batches
    assert any(tag.startswith("semisupervised-loss/TRAIN/LABEL/") for tag in writer.scalars)
    assert all(np.isfinite(v) or np.isnan(v) for v in writer.scalars.values())
    assert checkpoint.best_checkpoint is not None and np.isfinite(checkpoint.best_checkpoint["loss"])

    # a calibration epoch: only the two calibration tensors may move (model_training.py:146-149)
    rmt.train_one_epoch(epoch=2, epoch_type=RefEpoch.TRAIN, is_calibration_epoch=True, **common)
    calibrated = model.state_dict()
    moved = sorted(k for k in after if not torch.equal(after[k], calibrated[k]))
    assert moved == ["feature_clustering.parametrizations.artifact_stdev_k.original",
                     "feature_clustering.parametrizations.nonartifact_stdev_e.original"], moved

    # evaluation passes (what evaluate_model runs before plotting), with the reference's EvaluationMetrics and Datum
    metrics, worst = rmt.collect_evaluation_data(model, num_sources, balancer, downsampler, train, valid, report_worst=True)
    acc = metrics.accuracy_metrics_by_epoch_type
    assert set(int(k) for k in acc) == {0, 1}
    labeled = sum(int((b.int_tensor[:, 2] != 2).sum()) for b in train)
    assert abs(float(torch.Tensor(acc[RefEpoch.TRAIN]).sum()) - 0) >= 0         # tensor exists on the device
    assert sum(q.qsize() for q in worst.values()) > 0
    for (label, _), q in worst.items():
        assert label in (0, 1)
        for confidence, description in q.queue:
            assert confidence > 0 and description.startswith("20:")
    assert labeled > 0


@pytest.mark.gpu
def test_reference_generate_posterior_data_on_the_drop_in():
    rb, rd = _ref()
    import permutect.tools.filter_variants as rfv

    import bench
    from permutect_b200.tools.filter_variants import generate_posterior_arrays
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = bench.make_model(dev)
    from permutect_b200.utils.enums import Epoch
    model.set_epoch_type(Epoch.VALID)
    batches = _batches(2, 96, seed0=61)

    class _Dataset:
        def make_data_loader(self, batch_size, pin_memory=False, num_workers=0):
            return batches

    data = list(rfv.generate_posterior_data(_Dataset(), model, batch_size=96, num_workers=0))     # filter_variants.py:292-320
    assert len(data) == 192 and all(isinstance(d, rd.Datum) for d in data)
    want_int, want_float = zip(*generate_posterior_arrays(batches, model, dev))                   # pmt_pack_posterior
    want_int, want_float = np.vstack(want_int), np.vstack(want_float)
    got_int = np.vstack([d.get_int_array() for d in data])
    got_float = np.vstack([d.get_float_array() for d in data])
    assert got_int.dtype == want_int.dtype and np.array_equal(got_int, want_int)
    assert got_float.dtype == want_float.dtype and np.array_equal(got_float, want_float, equal_nan=True)
    assert (got_int[:, :2] == 0).all()                                                            # counts zeroed
    logits = got_float[:, rd.Data.CACHED_ARTIFACT_LOGIT.idx]
    assert np.array_equal(logits, logits.astype(np.float16).astype(np.float32))                   # fp16-rounded (quirk Q6)


@pytest.mark.gpu
def test_record_embeddings_reference_loop_and_drop_in_agree():
    """artifact_model.py:372-408: the REFERENCE's record_embeddings (its loop, its prefetch_generator, its EmbeddingMetrics
    with the TensorBoard output stubbed) run on this package's model collects the same embeddings and metadata as this
    package's record_embeddings; the embeddings are the set means compute_batch_output reports."""
    rb, rd = _ref()
    import permutect.architecture.artifact_model as ram
    import permutect.metrics.evaluation_metrics as rem

    import bench
    from permutect_b200.architecture.artifact_model import record_embeddings
    from permutect_b200.utils.enums import Epoch
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = bench.make_model(dev)
    model.set_epoch_type(Epoch.VALID)
    batches = _batches(3, 80, seed0=91)

    collected = []

    class _Metrics(rem.EmbeddingMetrics):
        def output_to_summary_writer(self, summary_writer, prefix="", **kwargs):
            collected.append(self)

    class _Loader(list):
        pass

    orig = ram.EmbeddingMetrics
    ram.EmbeddingMetrics = _Metrics
    try:
        ram.record_embeddings(model, _Loader(batches), summary_writer=None)
    finally:
        ram.EmbeddingMetrics = orig
    mine = record_embeddings(model, batches, summary_writer=None)
    assert len(collected) == 2
    with torch.inference_mode():
        want_alt = torch.vstack([model.compute_batch_output(b.copy_to(dev)).features_be.cpu() for b in batches])
    for theirs, ours in zip(collected, mine):
        for name in ("label_metadata", "correct_metadata", "type_metadata", "truncated_count_metadata"):
            assert getattr(theirs, name) == getattr(ours, name), name
        assert torch.equal(torch.vstack(theirs.features), torch.vstack(ours.features))
        assert len(theirs.ref_features) == len(ours.ref_features)
    assert torch.equal(torch.vstack(mine[0].features), want_alt)
    assert len(mine[0].label_metadata) == 240 and set(mine[0].label_metadata) <= {"artifact", "non-artifact", "unlabeled"}
