"""The ArtifactModel half of filter_variants (reference: permutect/tools/filter_variants.py:292-320): run the model over
the loader's batches and produce the posterior records (a Datum without reads whose info block is the embedding and
whose CACHED_ARTIFACT_LOGIT is the fp16-rounded artifact logit).

The reference builds every record in a Python loop over variants (one Datum, two ``set`` calls, one ``np.hstack`` per
variant); here the records of a whole batch are packed by one kernel (pmt_pack_posterior) and cross PCIe as two arrays.
Everything downstream of these records (VCF annotation, PosteriorModel, writing the filtered VCF) is outside this
repository's scope."""
import ctypes as C
from typing import Iterable, Iterator, Tuple

import numpy as np
import torch

from permutect_b200.data.batch import Batch
from permutect_b200.data.datum import COMPRESSED_READS_ARRAY_DTYPE, Datum
from permutect_b200.data.prefetch_generator import prefetch_generator
from permutect_b200.engine import library as L


def posterior_arrays_on_device(batch: Batch, logits_b: torch.Tensor, features_be: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(int16 [B, n_int], fp32 [B, 6 + E]) device tensors: the posterior records of the batch (filter_variants.py:302-320)."""
    lib = L.load()
    dev = logits_b.device
    if dev.type != "cuda" or batch.int_tensor.device != dev:
        raise RuntimeError("posterior records are packed on the GPU: batch and model outputs must be on the same CUDA device")
    B, E = batch.size(), features_be.shape[1]
    it, ft = batch.int_tensor, batch.float_tensor
    assert it.dtype == torch.int16 and ft.dtype == torch.float16
    logits_b, features_be = logits_b.detach().contiguous().float(), features_be.detach().contiguous().float()
    int_out = torch.empty((B, it.shape[1]), dtype=torch.int16, device=dev)
    float_out = torch.empty((B, 6 + E), dtype=torch.float32, device=dev)
    L.check(lib.pmt_pack_posterior(it.data_ptr(), it.stride(0), it.shape[1], ft.data_ptr(), ft.stride(0), logits_b.data_ptr(),
                                   features_be.data_ptr(), E, B, int_out.data_ptr(), float_out.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream))
    return int_out, float_out


@torch.inference_mode()
def generate_posterior_arrays(loader: Iterable[Batch], model, device=None) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """Per batch: (int16 [B, n_int], fp32 [B, 6 + E]) host arrays, exactly what MemoryMappedData.from_generator would store
    for the reference's posterior Datum objects (memory_mapped_data.py:319-338).

    The records of batch i return to pinned host memory on a copy stream while batch i + 1 is computed (its inputs are already
    on their way through prefetch_generator), so the arrays of a batch are yielded one iteration late; order and content are
    those of the plain loop."""
    device = torch.device(model._device if device is None else device)
    copy_stream = torch.cuda.Stream(device)
    pending = None
    for batch in prefetch_generator(loader, device):
        output = model.compute_batch_output(batch)
        int_out, float_out = posterior_arrays_on_device(batch, output.logits_b, output.features_be)
        copy_stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(copy_stream):
            int_host = torch.empty(int_out.shape, dtype=int_out.dtype, pin_memory=True).copy_(int_out, non_blocking=True)
            float_host = torch.empty(float_out.shape, dtype=float_out.dtype, pin_memory=True).copy_(float_out, non_blocking=True)
            int_out.record_stream(copy_stream)
            float_out.record_stream(copy_stream)
            done = torch.cuda.Event()
            done.record(copy_stream)
        if pending is not None:
            pending[2].synchronize()
            yield pending[0].numpy(), pending[1].numpy()
        pending = (int_host, float_host, done)
    if pending is not None:
        pending[2].synchronize()
        yield pending[0].numpy(), pending[1].numpy()


def generate_posterior_data(loader: Iterable[Batch], model, device=None) -> Iterator[Datum]:
    """Drop-in for filter_variants.generate_posterior_data (:292-320): yields one Datum per variant."""
    empty_reads = np.zeros((0, 0), dtype=COMPRESSED_READS_ARRAY_DTYPE)
    for int_out, float_out in generate_posterior_arrays(loader, model, device):
        for ia, fa in zip(int_out, float_out):
            yield Datum.from_posterior_record(ia, fa, empty_reads)
