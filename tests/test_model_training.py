"""The training-loop passes of SURVEY §8 f4 (permutect_b200/training/model_training.py) against a per-variant
restatement of the reference's loops (model_training.py:146-165, 203-271)."""
from queue import PriorityQueue

import numpy as np
import pytest
import torch

from permutect_b200.data.datum import Data
from permutect_b200.training import model_training as mt
from permutect_b200.utils.enums import Epoch, Label


def test_variant_description_helpers():
    # datum.py:41-47: a uint32 is stored as two int16s, base 65535 (sic), each shifted by 32768
    def two_int16s(num):
        return num // 65535 - 32768, num % 65535 - 32768
    for num in (0, 1, 65534, 65535, 123456789, 2_000_000_000):
        assert mt.uint32_from_two_int16s(*two_int16s(num)) == num
    # utils/allele_utils.py: base-5 digits A=1 C=2 G=3 T=4, least significant first
    assert mt.bases5_as_base_string(1 + 2 * 5 + 3 * 25 + 4 * 125) == "ACGT"
    assert mt.bases5_as_base_string(4) == "T" and mt.bases5_as_base_string(0) == ""
    # count_binning.py: alt bins {1-3}, {4-6}, ... -> centres 2, 5, 8, ...
    assert [mt.round_alt_count_to_bin_center(c) for c in (1, 2, 3, 4, 6, 7, 15)] == [2, 2, 2, 5, 5, 8, 14]
    ia = np.zeros(58, np.int16)
    ia[Data.CONTIG.idx] = 20
    ia[10], ia[11] = two_int16s(31_000_123)
    ia[12], ia[13] = two_int16s(2)            # C
    ia[14], ia[15] = two_int16s(4 + 4 * 5)    # TT
    assert mt.describe_variant(ia) == "20:31000123:C->TT"


class _Downsampler:
    def __init__(self, frac):
        self.frac = frac

    def calculate_downsampling_fractions(self, batch):
        f = torch.full((batch.size(),), self.frac)
        return f, f


class _Recorder:
    def __init__(self):
        self.calls = []

    def record_batch(self, epoch_type, batch, logits, weights):
        self.calls.append((epoch_type, batch.int_tensor.cpu().numpy().copy(), batch.counts()[1].cpu().numpy().copy(),
                           logits.cpu().numpy().copy()))

    def record(self, output, losses, batch):
        self.calls.append(float(losses.total_loss.detach()))


@pytest.mark.gpu
def test_epoch_loop_and_evaluation_passes():
    import bench
    from permutect_b200.data.batch import Batch
    from permutect_b200.synthetic import make_wgs_arrays
    from permutect_b200.training.step import make_optimizer
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = bench.make_model(dev)
    opt = make_optimizer(model)
    train = [Batch.from_arrays(*make_wgs_arrays(300, seed=s)) for s in (1, 2)]
    valid = [Batch.from_arrays(*make_wgs_arrays(200, seed=3))]
    before = torch.cat([p.detach().reshape(-1).clone() for p in model.parameters()])
    rec = _Recorder()
    n = mt.run_epoch(model, train, _Downsampler(0.7), Epoch.TRAIN, optimizer=opt, loss_recorder=rec)
    assert n == 4 and len(rec.calls) == 4 and all(np.isfinite(rec.calls))
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert not torch.equal(before, after)
    n = mt.run_epoch(model, valid, _Downsampler(0.7), Epoch.VALID, loss_recorder=rec)
    assert n == 2
    assert torch.equal(after, torch.cat([p.detach().reshape(-1) for p in model.parameters()]))   # no step in a VALID epoch

    metrics = _Recorder()
    _, worst = mt.collect_evaluation_data(model, 1, None, _Downsampler(0.7), train, valid, True, evaluation_metrics=metrics)
    assert [c[0] for c in metrics.calls] == [Epoch.TRAIN] * 6 + [Epoch.VALID] * 3
    # the reference's per-variant loop (model_training.py:229-266) over what was recorded
    want = {}
    for _, ints, alts, logits in metrics.calls:
        for ia, alt, logit in zip(ints, alts, logits.tolist()):
            label = int(ia[Data.LABEL.idx])
            if (label == Label.ARTIFACT and logit < 0) or (label == Label.VARIANT and logit > 0):
                q = want.setdefault((Label(label), mt.round_alt_count_to_bin_center(int(alt))), PriorityQueue(mt.WORST_OFFENDERS_QUEUE_SIZE))
                if q.full() and q.queue[0][0] < abs(logit):
                    q.get()
                if not q.full():
                    q.put((abs(logit), mt.describe_variant(ia)))
    assert set(want) == set(worst) and len(want) > 0
    for key in want:
        assert sorted(want[key].queue) == sorted(worst[key].queue)
