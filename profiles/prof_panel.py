"""High-depth panel sample (BASELINE config 5) through the forward: python profiles/prof_panel.py [n_variants] [iters] [precision]
Times the long-set tensor-core path and, with PMT_LONG_SIMT=1, the FP32 long-set kernel on the same batch."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
from permutect_b200.data.batch import Batch  # noqa: E402
from permutect_b200.engine import library as L  # noqa: E402
from permutect_b200.synthetic import make_panel_arrays  # noqa: E402
from permutect_b200.utils.enums import Epoch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L.set_precision(sys.argv[3] if len(sys.argv) > 3 else "tf32x3")
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.VALID)
ia, fa, reads = make_panel_arrays(n, seed=5000)
batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for simt in ((0, 1) if os.environ.get("PANEL_BOTH", "1") == "1" else (0,)):
    os.environ["PMT_LONG_SIMT"] = str(simt)
    with torch.inference_mode():
        out = model.compute_batch_output(batch)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(iters):
            out = model.compute_batch_output(batch)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    res[simt] = out.logits_b.clone()
    print(f"long sets on {'FP32 SIMT' if simt else 'tensor cores'}: {n} variants, {len(reads)} reads, {ms:.3f} ms, "
          f"{len(reads) / ms / 1e3:.1f} M reads/s, {n / ms:.1f} k variants/s")
if len(res) == 2:
    print("max |logit difference| between the two paths:", float((res[0] - res[1]).abs().max()))
