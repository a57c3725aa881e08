"""Cycle trace of CTA 0 of reads_forward_tc_kernel (pmt_set_reads_trace): python profiles/trace_reads.py [n_variants] [precision]
Prints, for the second tile of slot 0, every recorded event on one time axis:
  e0/e1 = epilogue warp of slot 0/1, m0/m1 = MMA warp of slot 0/1;
  4xx epilogue of step xx starts, 5xx operand written, 6xx arrived, 7xx accumulator seen; 1xx MMA waits, 2xx issues, 3xx issued."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
L.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32x3")
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.VALID)
batch = Batch.from_arrays(*make_wgs_arrays(n, seed=3000)).copy_to(dev)
lib = L.load()
buf = torch.zeros(4 * 2048, dtype=torch.int64, device=dev)
with torch.inference_mode():
    model.compute_batch_output(batch)
    lib.pmt_set_reads_trace(buf.data_ptr())
    model.compute_batch_output(batch)
    torch.cuda.synchronize()
    lib.pmt_set_reads_trace(None)
t = buf.cpu().view(4, 1024, 2)
names = ["e0", "e1", "m0", "m1"]
starts = [i for i in range(int(buf[2046])) if int(t[0, i, 0]) == 1]
lo, hi = int(t[0, starts[2], 1]), int(t[0, starts[3], 1])
ev = []
for w in range(4):
    for i in range(int(buf[w * 2048 + 2046])):
        c = int(t[w, i, 1])
        if lo <= c <= hi:
            ev.append((c - lo, names[w], int(t[w, i, 0])))
ev.sort()
print(f"tile period of slot 0: {hi - lo} cycles")
print("  ".join(f"{c}:{w}:{e}" for c, w, e in ev))
