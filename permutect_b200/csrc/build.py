"""Builds libpermutect_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m permutect_b200.csrc.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["pmt_forward.cu", "pmt_backward.cu", "pmt_loader.cu", "pmt_tc.cu", "pmt_cnn_tc.cu", "pmt_loss.cu", "pmt_optim.cu"]
HEADERS = ["pmt_device.cuh", "pmt_tile.cuh", "pmt_cnn.cuh", "pmt_host.h", "pmt_tc_ptx.cuh", os.path.join("..", "..", "include", "permutect_b200.h")]
OUTPUT = os.path.join(HERE, "libpermutect_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _fingerprint() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in SOURCES + HEADERS:
        with open(os.path.join(HERE, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = OUTPUT + ".stamp"
    fp = _fingerprint()
    if not force and os.path.exists(OUTPUT) and os.path.exists(stamp) and open(stamp).read() == fp:
        return OUTPUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUTPUT] + [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    with open(stamp, "w") as f:
        f.write(fp)
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
