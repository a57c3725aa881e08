"""Small driver for ncu captures of the posterior kernels: python profiles/prof_posterior.py [n_records]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
z = np.load(os.path.join(REPO, "tests", "golden", "posterior_model.npz"))
keep = z["int_array"][:, 5] <= 400
reps = n // int(keep.sum()) + 1
ia = np.tile(z["int_array"][keep], (reps, 1))[:n]
fa = np.tile(z["float_array"][keep], (reps, 1))[:n]
dev = torch.device("cuda:0")
model = PosteriorModel(-10.0, -10.0, device=dev)
batch = PosteriorBatch(ia, fa, dev)
for _ in range(2):
    probs = model.posterior_probabilities_bc(batch)
    loss = model.negative_log_evidence(batch)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for _ in range(10):
    model.posterior_probabilities_bc(batch)
ev[1].record()
for _ in range(10):
    model.negative_log_evidence(batch)
ev[2].record()
torch.cuda.synchronize()
print(f"ok n={n} posterior_probabilities_bc {ev[0].elapsed_time(ev[1]) / 10:.3f} ms, negative_log_evidence (E step) "
      f"{ev[1].elapsed_time(ev[2]) / 10:.3f} ms; mean P(somatic)={float(probs[:, 0].mean()):.4f} loss={float(loss.detach()):.4f}")
