"""Small forward-only driver for ncu captures: python profiles/prof_forward.py [n_variants] [iters] [precision]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
from permutect_b200.data.batch import Batch  # noqa: E402
from permutect_b200.engine import library as L  # noqa: E402
from permutect_b200.synthetic import make_wgs_arrays  # noqa: E402
from permutect_b200.utils.enums import Epoch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L.set_precision(sys.argv[3] if len(sys.argv) > 3 else "fp32")
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.VALID)
batch = Batch.from_arrays(*make_wgs_arrays(n, seed=3000)).copy_to(dev)
with torch.inference_mode():
    for _ in range(iters):
        out = model.compute_batch_output(batch)
torch.cuda.synchronize()
print("ok", float(out.logits_b.mean()))
