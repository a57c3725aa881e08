"""One optimisation step, as the reference's training loop performs it
(model_training.py:151-165 with misc_utils.backpropagate :125-129): forward, losses, backward,
clip_grad_norm_(1.0), AdamW.  With a process group the flat gradient is summed across ranks before
clipping, so every rank applies the identical update."""
from typing import Iterable, Optional

import torch
from torch import nn

from permutect_b200.training import distributed as pdist


def make_optimizer(model: nn.Module, learning_rate: float = 1e-3, weight_decay: float = 0.01) -> torch.optim.Optimizer:
    """AdamW as in model_training.py:68-72; the fused (single multi-tensor kernel) variant when on CUDA."""
    params = [p for p in model.parameters()]
    on_cuda = len(params) > 0 and params[0].is_cuda
    return torch.optim.AdamW(params, lr=learning_rate, weight_decay=weight_decay, fused=on_cuda)


def backpropagate(optimizer: torch.optim.Optimizer, loss: torch.Tensor, params_to_clip: Iterable[nn.Parameter] = (),
                  process_group=None):
    """misc_utils.py:125-129 (+ the data-parallel gradient sum)."""
    params_to_clip = list(params_to_clip)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if process_group is not None or pdist.is_initialized():
        pdist.allreduce_gradients(params_to_clip, process_group)
    nn.utils.clip_grad_norm_(params_to_clip, max_norm=1.0)
    optimizer.step()


def train_step(model, batch, optimizer, balancer=None, process_group=None):
    """compute_batch_output -> compute_batch_losses -> backpropagate.  Returns (output, losses)."""
    output = model.compute_batch_output(batch, balancer)
    losses = model.compute_batch_losses(output, batch)
    backpropagate(optimizer, losses.total_loss, params_to_clip=model.parameters(), process_group=process_group)
    return output, losses
