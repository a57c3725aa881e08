"""Cycle trace of CTA 0 of hap_cnn_tc_kernel (pmt_set_cnn_trace): python profiles/trace_cnn.py [n_variants] [precision]"""
import ctypes as C
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32x3")
dev = torch.device("cuda:0")
model = bench.make_model(dev)
model.set_epoch_type(Epoch.VALID)
batch = Batch.from_arrays(*make_wgs_arrays(n, seed=3000)).copy_to(dev)
lib = L.load()
lib.pmt_set_cnn_trace.argtypes = [C.c_void_p]
buf = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
with torch.inference_mode():
    model.compute_batch_output(batch)
    lib.pmt_set_cnn_trace(buf.data_ptr())
    model.compute_batch_output(batch)
    torch.cuda.synchronize()
    lib.pmt_set_cnn_trace(None)
t = buf.cpu().view(3, 512, 2)
t0 = int(t[0, 0, 1])
for w, name in enumerate(["epi warp 0", "epi warp 4", "mma warp"]):
    cnt = int(buf[w * 1024 + 1022])
    print(f"--- {name}: {cnt} events (id, cycles since start, delta)")
    prev = t0
    for i in range(min(cnt, 110)):
        ev, clk = int(t[w, i, 0]), int(t[w, i, 1])
        print(f"{ev:5d} {clk - t0:9d} {clk - prev:7d}")
        prev = clk
