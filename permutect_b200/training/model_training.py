"""The ArtifactModel passes of the reference's training loop (permutect/training/model_training.py): the inner loop of
``train_one_epoch`` (:133-165) and the evaluation passes ``collect_evaluation_data`` (:203-271), which once the training
step is fast are more than half of an epoch.

What changes against the reference: DownsampledBatch is built on the device (no host synchronisation), the model runs
through libpermutect_b200, and the per-variant Python loop that collects the worst offenders (:229-266) only visits the
wrong calls, found by one comparison on the device, instead of every variant of every batch.  The collaborators that are
out of this repository's scope are duck-typed exactly as the reference calls them: ``downsampler`` needs
``calculate_downsampling_fractions(batch)`` (training/downsampler.py), ``balancer`` is passed to
``compute_batch_output`` (training/balancer.py), ``loss_recorder.record(output, losses, batch)`` (metrics/loss_metrics.py),
``evaluation_metrics.record_batch(epoch_type, batch, logits=, weights=)`` (metrics/evaluation_metrics.py).
"""
from collections import defaultdict
from queue import PriorityQueue
from typing import Callable, Iterable, Optional

import numpy as np
import torch

from permutect_b200.data.batch import Batch, DownsampledBatch
from permutect_b200.data.datum import Data
from permutect_b200.data.prefetch_generator import prefetch_generator
from permutect_b200.training import distributed as pdist
from permutect_b200.training.step import backpropagate
from permutect_b200.utils.enums import Epoch, Label

WORST_OFFENDERS_QUEUE_SIZE = 100            # model_training.py:46
NUM_DOWNSAMPLING_ITERATIONS = 2             # model_training.py:153
NUM_EVALUATION_ITERATIONS = 3               # model_training.py:224 ("TODO: magic constant")
_BIGGEST_UINT16, _BIGGEST_INT16 = 65535, 32767   # datum.py:28-29 (the pair is base 65535 in the reference, sic)
_POSITION_IDX, _REF_ALLELE_IDX, _ALT_ALLELE_IDX = 10, 12, 14


def uint32_from_two_int16s(a: int, b: int) -> int:
    """datum.py:45-47."""
    return _BIGGEST_UINT16 * (int(a) + _BIGGEST_INT16 + 1) + (int(b) + _BIGGEST_INT16 + 1)


def bases5_as_base_string(base5: int) -> str:
    """utils/allele_utils.py:77-85."""
    result, remaining = "", base5
    while remaining > 0:
        digit = remaining % 5
        result += "A" if digit == 1 else ("C" if digit == 2 else ("G" if digit == 3 else "T"))
        remaining = (remaining - digit) // 5
    return result


def round_alt_count_to_bin_center(count: int) -> int:
    """data/count_binning.py:69-88 (MIN_ALT_COUNT 1, COUNT_BIN_SKIP 3)."""
    return 1 + 3 * ((count - 1) // 3) + 1


def describe_variant(int_array: np.ndarray) -> str:
    """contig:position:ref->alt as model_training.py:255-263 prints it."""
    return (str(int_array[Data.CONTIG.idx]) + ":" + str(uint32_from_two_int16s(int_array[_POSITION_IDX], int_array[_POSITION_IDX + 1]))
            + ":" + bases5_as_base_string(uint32_from_two_int16s(int_array[_REF_ALLELE_IDX], int_array[_REF_ALLELE_IDX + 1]))
            + "->" + bases5_as_base_string(uint32_from_two_int16s(int_array[_ALT_ALLELE_IDX], int_array[_ALT_ALLELE_IDX + 1])))


def freeze(parameters):
    """misc_utils.py:143-145."""
    for p in parameters:
        p.requires_grad = False


def unfreeze(parameters):
    """misc_utils.py:148-151."""
    for p in parameters:
        if p.dtype.is_floating_point:
            p.requires_grad = True


def run_epoch(model, loader: Iterable[Batch], downsampler, epoch_type: Epoch, optimizer=None, balancer=None,
              loss_recorder=None, process_group=None, on_step: Optional[Callable] = None,
              is_calibration_epoch: bool = False) -> int:
    """The batch loop of train_one_epoch (model_training.py:143-165): two downsampled draws of every parent batch, losses
    recorded, and in a TRAIN epoch one optimiser step per draw.  In a calibration epoch only
    ``model.calibration_parameters()`` train (:146-149).  With a process group (or an initialised default group of more than
    one rank) the gradient is summed across ranks inside the optimiser step and the balancer's counters are kept global
    (training/distributed.py:SyncedBalancer), so every rank holds the same weights and the same balancing state.
    Returns the number of downsampled batches processed."""
    model.set_epoch_type(epoch_type)
    if is_calibration_epoch and epoch_type == Epoch.TRAIN:
        freeze(model.parameters())
        unfreeze(model.calibration_parameters())
    if balancer is not None and (process_group is not None or pdist.is_initialized()) and not isinstance(balancer, pdist.SyncedBalancer):
        balancer = pdist.SyncedBalancer(balancer, process_group)
    device = model._device
    n = 0
    for parent_batch in prefetch_generator(loader, device):
        for _ in range(NUM_DOWNSAMPLING_ITERATIONS):
            ref_fracs_b, alt_fracs_b = downsampler.calculate_downsampling_fractions(parent_batch)
            batch = DownsampledBatch(parent_batch, ref_fracs_b, alt_fracs_b)
            if epoch_type == Epoch.TRAIN:
                output = model.compute_batch_output(batch, balancer)
                losses = model.compute_batch_losses(output, batch)
            else:
                with torch.no_grad():
                    output = model.compute_batch_output(batch, balancer)
                    losses = model.compute_batch_losses(output, batch)
            if loss_recorder is not None:
                loss_recorder.record(output, losses, batch)
            if epoch_type == Epoch.TRAIN:
                backpropagate(optimizer, losses.total_loss, params_to_clip=model.parameters(), process_group=process_group)
            if on_step is not None:
                on_step(batch, output, losses)
            n += 1
    return n


def wrong_calls(batch: Batch, logits_b: torch.Tensor):
    """Indices (device tensor) of the labeled variants the model calls wrongly (model_training.py:238-240)."""
    labels = batch.int_tensor[:, Data.LABEL.idx]
    return torch.nonzero(((labels == int(Label.ARTIFACT)) & (logits_b < 0)) | ((labels == int(Label.VARIANT)) & (logits_b > 0))).view(-1)


@torch.inference_mode()
def collect_evaluation_data(model, num_sources: int, balancer, downsampler, train_loader, valid_loader, report_worst: bool,
                            evaluation_metrics=None):
    """model_training.py:203-271.  ``evaluation_metrics`` is the reference's EvaluationMetrics (or anything with its
    ``record_batch``); the worst-offender queues are keyed and filled exactly as the reference's."""
    worst = defaultdict(lambda: PriorityQueue(WORST_OFFENDERS_QUEUE_SIZE))
    device = model._device
    for epoch_type in (Epoch.TRAIN, Epoch.VALID):
        loader = train_loader if epoch_type == Epoch.TRAIN else valid_loader
        for parent_batch in prefetch_generator(loader, device):
            for _ in range(NUM_EVALUATION_ITERATIONS):
                ref_fracs_b, alt_fracs_b = downsampler.calculate_downsampling_fractions(parent_batch)
                batch = DownsampledBatch(parent_batch, ref_fracs_b, alt_fracs_b)
                output = model.compute_batch_output(batch, balancer)
                if evaluation_metrics is not None:
                    evaluation_metrics.record_batch(epoch_type, batch, logits=output.logits_b, weights=output.weights)
                if report_worst:
                    idx = wrong_calls(batch, output.logits_b)
                    if idx.numel() == 0:
                        continue
                    rows = batch.int_tensor[idx].cpu().numpy()
                    conf = output.logits_b[idx].abs().cpu().tolist()
                    alt_counts = batch.counts()[1][idx].cpu().tolist()
                    for int_array, confidence, alt_count in zip(rows, conf, alt_counts):
                        key = (Label(int(int_array[Data.LABEL.idx])), round_alt_count_to_bin_center(int(alt_count)))
                        pqueue = worst[key]
                        if pqueue.full() and pqueue.queue[0][0] < confidence:
                            pqueue.get()          # discards the least confident bad call
                        if not pqueue.full():
                            pqueue.put((confidence, describe_variant(int_array)))
    return evaluation_metrics, worst
