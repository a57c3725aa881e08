"""Small-batch (64-variant) inference call and training step replayed from CUDA graphs (engine/graphs.py): per-call time
with CUDA events, for an ncu launch list of the kernels inside the graphs.
python profiles/prof_small_batch.py [batch_variants] [calls]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "tests")]
import torch
import bench
from permutect_b200.data.batch import Batch, DownsampledBatch
from permutect_b200.engine import library as L
from permutect_b200.engine.graphs import GraphedInference, GraphedTrainStep
from permutect_b200.synthetic import make_wgs_arrays
from permutect_b200.training.step import make_optimizer
from permutect_b200.utils.enums import Epoch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda:0")
model = bench.make_model(dev)
L.set_precision("tf32x3")
parent = Batch.from_arrays(*make_wgs_arrays(n, seed=4000)).copy_to(dev)

def timed(fn, k):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / k, 1e3 * e0.elapsed_time(e1) / k

model.set_epoch_type(Epoch.VALID)
infer = GraphedInference(model, parent)
w, g = timed(lambda: infer(parent), calls)
print(f"inference graphed: wall {w:.1f} us/call, device {g:.1f} us/call")
w, g = timed(lambda: infer.graph.replay(), calls)
print(f"inference replay only: wall {w:.1f} us/call, device {g:.1f} us/call")
with torch.inference_mode():
    w, g = timed(lambda: model.compute_batch_output(parent), calls)
print(f"inference eager: wall {w:.1f} us/call, device {g:.1f} us/call")

model.set_epoch_type(Epoch.TRAIN)
opt = make_optimizer(model, learning_rate=1e-3, weight_decay=0.01)
frac = torch.full((n,), 0.8, device=dev)
step = GraphedTrainStep(model, opt, DownsampledBatch(parent, frac, frac, seed=1))
seed = [100]
def one():
    seed[0] += 1
    step(DownsampledBatch(parent, frac, frac, seed=seed[0]))
w, g = timed(one, calls)
print(f"train graphed (eager downsample + load + replay): wall {w:.1f} us/step, device {g:.1f} us/step")
w, g = timed(lambda: step.graph.replay(), calls)
print(f"train replay only: wall {w:.1f} us/step, device {g:.1f} us/step")
