"""Device-side DownsampledBatch construction (reference: permutect/data/batch.py:383-439).

Two kernel launches and one scan, all stream-ordered; unlike the reference's ``torch.nonzero`` path
there is no host synchronisation, so downsampling can be queued ahead of the forward pass.
"""
import random
from typing import Optional

import torch

from permutect_b200.engine import library as L


def downsample_on_device(parent, ref_fracs_b: torch.Tensor, alt_fracs_b: torch.Tensor, offset_alt_rows: bool = False,
                         seed: Optional[int] = None, random_int: Optional[int] = None):
    """Returns (read_indices [N'], new_ref_counts [B], new_alt_counts [B]) as int64 tensors on the batch's device.
    N' is not known on the host: read_indices is allocated at the parent's row count and only its first
    sum(new counts) entries are meaningful (the kernels take totals from the offsets)."""
    lib = L.load()
    ref_off, alt_off = parent.offsets()
    dev = ref_off.device
    if dev.type != "cuda":
        raise RuntimeError("DownsampledBatch is built on the GPU; move the parent batch there first")
    B = parent.size()
    if seed is None:
        seed = random.getrandbits(63)
    if random_int is None:
        random_int = random.randint(0, 100)          # batch.py:418
    rf = ref_fracs_b.to(dev, torch.float32).contiguous()
    af = alt_fracs_b.to(dev, torch.float32).contiguous()
    counts = torch.empty((2, B), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    L.check(lib.pmt_downsample_counts(ref_off.data_ptr(), alt_off.data_ptr(), rf.data_ptr(), af.data_ptr(), B, seed,
                                      random_int, counts[0].data_ptr(), counts[1].data_ptr(), stream))
    new_off = torch.zeros((2, B + 1), dtype=torch.int64, device=dev)
    torch.cumsum(counts[0], dim=0, out=new_off[0, 1:])      # 1-D scans: single-pass device scan
    torch.cumsum(counts[1], dim=0, out=new_off[1, 1:])
    read_indices = torch.zeros(parent.reads.shape[0], dtype=torch.int64, device=dev)   # the tail past N' stays a valid index
    L.check(lib.pmt_downsample_fill(ref_off.data_ptr(), alt_off.data_ptr(), rf.data_ptr(), af.data_ptr(), B, seed,
                                    random_int, new_off[0].data_ptr(), new_off[1].data_ptr(), int(offset_alt_rows),
                                    read_indices.data_ptr(), stream))
    return read_indices, counts[0], counts[1], new_off
