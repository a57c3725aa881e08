import sys, random, traceback
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import torch
import test_reference_callers as T
class MP:
    def setattr(self, obj, name, val): setattr(obj, name, val)
from permutect_b200.engine import library as L
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fails = 0
for i in range(n):
    random.seed(1000 + i)
    L.set_precision(mode)
    try:
        T.test_reference_train_one_epoch_and_evaluation_run_on_the_drop_in(MP())
    except Exception as e:
        fails += 1
        print("FAIL seed", 1000 + i, type(e).__name__, str(e)[:300])
        traceback.print_exc(limit=3)
print(mode, "fails", fails, "of", n)
