"""Phase clocks of CTA 0's third tile in reads_backward_kernel (pmt_set_backward_trace):
python profiles/trace_backward.py [n_variants]"""
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

NAMES = {0: "tile built", 1: "recompute: read embedding", 20: "recompute: reducer, rotation, head", 21: "head + rotation backward",
         22: "reducer backward", 23: "concat (d info_seq)", 24: "read embedding backward"}
NAMES.update({200: "  mlp layer: barrier", 201: "  mlp layer: reload input + weight image ready", 202: "  mlp layer: SELU of the block input, prefetch",
              203: "  mlp layer: wgrad_tile", 204: "  mlp layer: bias rowdot", 205: "  mlp layer: dgrad gemm (thread 0 done)",
              300: "    gemm: entered", 301: "    gemm: accumulators initialised", 302: "    gemm: k loop", 303: "    gemm: epilogue"})
for b in range(8):
    NAMES[2 + b] = f"recompute: gated block {b}"
    for q, what in enumerate(["reload x, z; LayerNorms, means, gate", "proj2 wgrad + dgrad", "gate / mean-field / LN2 backward",
                              "proj1 wgrad + dgrad", "LayerNorm backward"]):
        NAMES[100 + 10 * b + q] = f"block {b} backward: {what}"

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.TRAIN)
opt = make_optimizer(model)
parent = Batch.from_arrays(*make_wgs_arrays(n, seed=3000)).copy_to(dev)
rf = torch.full((n,), 0.65, device=dev)
lib = L.load()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
for i in range(3):
    if i == 2:
        lib.pmt_set_backward_trace(buf.data_ptr())
    train_step(model, DownsampledBatch(parent, rf, rf, seed=i), opt)
torch.cuda.synchronize()
lib.pmt_set_backward_trace(None)
t = buf.cpu().tolist()
recs = [(t[1 + 2 * i], t[2 + 2 * i]) for i in range(t[0])]
print(f"{len(recs)} records; tile total {recs[-1][1] - recs[0][1]} cycles")
agg = {}
for (_, c0), (pid, c1) in zip(recs, recs[1:]):
    key = NAMES.get(pid, str(pid))
    if pid >= 100:
        key = "block backward: " + key.split(": ", 1)[1]
    elif 2 <= pid < 20:
        key = "recompute: gated block"
    elif pid >= 200:
        key = "read embedding backward, skip-block layers:" + key
    agg[key] = agg.get(key, 0) + (c1 - c0)
    print(f"{pid:4d} {c1 - c0:9d}  {NAMES.get(pid, '')}")
print("--- summed over blocks")
for k, v in agg.items():
    print(f"{v:9d}  {100 * v / (recs[-1][1] - recs[0][1]):5.1f} %  {k}")
