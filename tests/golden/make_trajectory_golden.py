"""Golden for N-step training-trajectory parity: the UNMODIFIED reference's training step (compute_batch_output ->
compute_batch_losses -> misc_utils.backpropagate with torch.optim.AdamW as model_training.py:68-72 builds it) run for six
optimiser steps over three fixed batches (no downsampling, so no random numbers), starting from the perturbed two-source
model of ``v040_two_sources.npz``.  Saved: the batches, the loss of every step and the state_dict after the last step.

Runs only in the build container (needs /root/reference):  python tests/golden/make_trajectory_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the import stubs and /root/reference on sys.path)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from permutect.architecture.artifact_model import ArtifactModel  # noqa: E402
from permutect.data.batch import Batch  # noqa: E402
from permutect.misc_utils import backpropagate  # noqa: E402
from permutect.utils.enums import Epoch  # noqa: E402

N_STEPS, N_BATCHES, BATCH = 6, 3, 24
LR, WEIGHT_DECAY = 1e-3, 0.01          # parameters.py defaults used by train_artifact_model


def main():
    cpu = torch.device("cpu")
    # the model of make_golden.py case 4, rebuilt with the same seeds
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    model = ArtifactModel(mg.make_params(mg.V040), 61, 71, 42, device=cpu)
    model.reset_source_predictor(2)
    mg.perturb(model, 4)
    model.source_predictor.set_adversarial_strength(0.4621)
    arrays = {}
    for k, v in model.state_dict().items():
        arrays["sd0/" + k] = v.detach().numpy().copy()
    rng = np.random.default_rng(77)
    batches = []
    for b in range(N_BATCHES):
        data = [mg.make_datum(rng, int(rng.integers(0, 11)), int(rng.integers(1, 16)), int(rng.integers(0, 3)),
                              source=int(rng.integers(0, 2))) for _ in range(BATCH)]
        batch = Batch(data).copy_to(cpu, torch.float32)
        batches.append(batch)
        for k, v in mg.raw_inputs(data, batch).items():
            if k != "decoded_reads":
                arrays[f"in{b}/" + k] = v
    optimizer = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WEIGHT_DECAY)
    model.set_epoch_type(Epoch.TRAIN)
    losses = []
    for step in range(N_STEPS):
        batch = batches[step % N_BATCHES]
        out = model.compute_batch_output(batch)
        ls = model.compute_batch_losses(out, batch)
        backpropagate(optimizer, ls.total_loss, params_to_clip=model.parameters())
        losses.append(float(ls.total_loss))
    arrays["losses"] = np.array(losses, np.float64)
    arrays["meta"] = np.array([N_STEPS, N_BATCHES, BATCH], np.int64)
    for k, v in model.state_dict().items():
        arrays["sd1/" + k] = v.detach().numpy()
    path = os.path.join(HERE, "v040_trajectory.npz")
    np.savez_compressed(path, **arrays)
    moved = max(float(np.abs(arrays["sd1/" + k] - arrays["sd0/" + k]).max()) for k in model.state_dict() if arrays["sd0/" + k].dtype.kind == "f")
    print(f"losses {losses}; largest weight change {moved:.4g}; {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
