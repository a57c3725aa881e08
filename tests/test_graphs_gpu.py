"""CUDA-graph replay (engine/graphs.py) of the inference call and of the whole optimisation step at the reference tools'
default batch size: identical results to the eager path, batches that do not fit the capture fall back to it."""
import numpy as np
import pytest
import torch

import bench
from permutect_b200.data.batch import Batch, DownsampledBatch
from permutect_b200.engine import library as L
from permutect_b200.engine.graphs import GraphedInference, GraphedTrainStep
from permutect_b200.synthetic import make_wgs_arrays
from permutect_b200.training.step import make_optimizer, train_step
from permutect_b200.utils.enums import Epoch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_graphed_inference_equals_eager(mode):
    dev = torch.device("cuda:0")
    L.set_precision(mode)
    try:
        model = bench.make_model(dev)
        model.set_epoch_type(Epoch.VALID)
        batches = [Batch.from_arrays(*make_wgs_arrays(64, seed=70 + i)) for i in range(4)]
        infer = GraphedInference(model, batches[0])
        for b in batches:
            got = infer(b.pin_memory())
            got = {k: getattr(got, k).clone() for k in ("logits_b", "logits_bk", "features_be", "ref_features_be", "outlier_binary_logits")}
            with torch.inference_mode():
                want = model.compute_batch_output(b.copy_to(dev))
            for k, v in got.items():
                assert torch.equal(v, getattr(want, k)), k
        odd = Batch.from_arrays(*make_wgs_arrays(37, seed=99))          # another size: eager path, same API
        assert infer(odd).logits_b.shape[0] == 37
    finally:
        L.set_precision("fp32")


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_graphed_training_steps_equal_eager_steps(mode):
    dev = torch.device("cuda:0")
    L.set_precision(mode)
    try:
        parent = Batch.from_arrays(*make_wgs_arrays(64, seed=4000)).copy_to(dev)
        frac = torch.full((64,), 0.8, device=dev)
        # the keep decisions also draw from torch's global generator (the reference's randint): build the batches once
        example = DownsampledBatch(parent, frac, frac, seed=1)
        batches = [DownsampledBatch(parent, frac, frac, seed=100 + i) for i in range(6)]

        def run(graphed):
            model = bench.make_model(dev)
            model.set_epoch_type(Epoch.TRAIN)
            opt = make_optimizer(model, learning_rate=1e-3, weight_decay=0.01)
            step = GraphedTrainStep(model, opt, example) if graphed else None
            losses = []
            for batch in batches:
                if graphed:
                    losses.append(float(step(batch)))
                else:
                    losses.append(float(train_step(model, batch, opt)[1].total_loss.detach()))
            return losses, opt.flat.detach().cpu().numpy().copy()

        l_eager, w_eager = run(False)
        l_graph, w_graph = run(True)
        assert l_eager == l_graph
        np.testing.assert_array_equal(w_eager, w_graph)
    finally:
        L.set_precision("fp32")
