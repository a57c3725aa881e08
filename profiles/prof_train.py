"""Small training driver for ncu captures: python profiles/prof_train.py [n_variants] [iters] [precision = tf32x3]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
from permutect_b200.data.batch import Batch, DownsampledBatch  # noqa: E402
from permutect_b200.engine import library as L  # noqa: E402
from permutect_b200.synthetic import make_wgs_arrays  # noqa: E402
from permutect_b200.training.step import make_optimizer, train_step  # noqa: E402
from permutect_b200.utils.enums import Epoch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L.set_precision(sys.argv[3] if len(sys.argv) > 3 else "tf32x3")
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.TRAIN)
opt = make_optimizer(model)
parent = Batch.from_arrays(*make_wgs_arrays(n, seed=3000)).copy_to(dev)
rf = torch.full((n,), 0.65, device=dev)
for i in range(iters):
    batch = DownsampledBatch(parent, rf, rf, seed=i)
    out, losses = train_step(model, batch, opt)
torch.cuda.synchronize()
print("ok", float(losses.total_loss.detach()) / n)
