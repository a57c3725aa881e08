"""Parameter containers of the ArtifactModel.

These modules own the trainable tensors under exactly the attribute paths the reference uses, so
``state_dict()`` / ``load_state_dict()`` / ``named_parameters()`` are interchangeable with
``permutect.architecture.*``.  They are constructed from the same torch
building blocks in the same order, so a given ``torch.manual_seed`` yields the same initial weights
as the reference.  They deliberately have NO forward arithmetic: the computation lives in the CUDA
kernels (``permutect_b200/csrc``), which read the materialised tensors through
``permutect_b200.engine.plan``.

Reference files mirrored: mlp.py:8-76, gated_mlp.py:148-276, dna_sequence_convolution.py:29-111,
euclidean_transformation.py:8-23, feature_clustering.py:49-80, exponentially_modified_gaussian.py:58-80,
parameterizations.py:21-112, adversarial.py:6-27.
"""
from math import floor
from typing import List

import torch
from torch import nn
from torch.nn import Parameter
from torch.nn.utils import parametrize
from torch.nn.utils.parametrizations import orthogonal

MIN_BOUND, MAX_BOUND = 0.01, 100.0   # stdev and lambda bounds (feature_clustering.py:49-51; emg.py:14-20)


def _no_forward(self, *args, **kwargs):
    raise RuntimeError(f"{type(self).__name__} is a parameter container; the arithmetic runs in the fused CUDA kernels")


# ---- constraints (parameterizations.py) -------------------------------------------------------
class UnitVector(nn.Module):
    def forward(self, x):
        return x / torch.norm(x, dim=-1, keepdim=True)


class PositiveNumber(nn.Module):
    def forward(self, x):
        return torch.exp(x)

    def right_inverse(self, p):
        return torch.log(p)


class BoundedNumber(nn.Module):
    def __init__(self, min_val: float, max_val: float):
        super().__init__()
        assert min_val <= max_val
        self.min_val, self.max_val, self.size = min_val, max_val, max_val - min_val

    def forward(self, x):
        return self.size * torch.sigmoid(x) + self.min_val

    def right_inverse(self, p):
        return torch.logit((p - self.min_val) / self.size)


class LogWeights(nn.Module):
    def forward(self, x):
        return torch.log_softmax(x, dim=-1)


# ---- MLP (mlp.py) -------------------------------------------------------------------------------
class DenseSkipBlock(nn.Module):
    """x + alpha * g(x), g = num_layers x (SELU -> Linear) at constant width (mlp.py:8-22)."""

    def __init__(self, input_size: int, num_layers: int):
        super().__init__()
        self.mlp = MLP((num_layers + 1) * [input_size], prepend_activation=True)
        self.alpha = Parameter(torch.tensor(0.1))

    forward = _no_forward


class MLP(nn.Module):
    """Linear/SELU stack; a negative width -d is a d-layer DenseSkipBlock (mlp.py:25-76).
    ``_model`` keeps the reference's Sequential indexing (activations occupy a slot)."""

    def __init__(self, layer_sizes: List[int], batch_normalize: bool = False, dropout_p: float = 0,
                 prepend_activation: bool = False):
        super().__init__()
        if batch_normalize or dropout_p > 0:
            raise NotImplementedError("batch_normalize / dropout are not supported by the fused kernels "
                                      "(both default off in the reference, parameters.py:28,138-142)")
        self.layer_sizes = list(layer_sizes)
        layers = [nn.SELU()] if prepend_activation else []
        self._input_dim = input_dim = layer_sizes[0]
        for k, output_dim in enumerate(layer_sizes[1:]):
            if output_dim < 0:
                layers.append(DenseSkipBlock(input_dim, -output_dim))
                continue
            layers.append(nn.Linear(input_dim, output_dim))
            if k < len(layer_sizes) - 2:
                layers.append(nn.SELU())
            input_dim = output_dim
        self._output_dim = input_dim
        self._model = nn.Sequential(*layers)

    def input_dimension(self) -> int:
        return self._input_dim

    def output_dimension(self) -> int:
        return self._output_dim

    forward = _no_forward


# ---- gated ref/alt MLP (gated_mlp.py) ----------------------------------------------------------
class SpatialGatingUnitRefAlt(nn.Module):
    def __init__(self, d_z: int):
        super().__init__()
        self.norm = nn.LayerNorm([d_z // 2])
        self.alpha_ref = Parameter(torch.tensor(0.01))
        self.alpha_alt = Parameter(torch.tensor(0.01))
        self.beta_ref = Parameter(torch.tensor(0.01))
        self.beta_alt = Parameter(torch.tensor(0.01))
        self.gamma = Parameter(torch.tensor(0.01))
        self.ref_regularizer = Parameter(0.1 * torch.ones(d_z // 2))
        self.reg_weight = Parameter(torch.tensor(0.1))
        parametrize.register_parametrization(self, "reg_weight", PositiveNumber())

    forward = _no_forward


class GatedRefAltMLPBlock(nn.Module):
    def __init__(self, d_model: int, d_ffn: int):
        super().__init__()
        self.norm = nn.LayerNorm([d_model])
        self.activation = nn.SELU()
        self.proj1_ref = nn.Linear(d_model, d_ffn)
        self.proj1_alt = nn.Linear(d_model, d_ffn)
        self.sgu = SpatialGatingUnitRefAlt(d_ffn)
        self.proj2_ref = nn.Linear(d_ffn // 2, d_model)
        self.proj2_alt = nn.Linear(d_ffn // 2, d_model)
        self.size = d_model

    forward = _no_forward


class GatedRefAltMLP(nn.Module):
    def __init__(self, d_model: int, d_ffn: int, num_blocks: int):
        super().__init__()
        self.blocks = nn.ModuleList([GatedRefAltMLPBlock(d_model, d_ffn) for _ in range(num_blocks)])
        self.dimension = d_model
        self.d_ffn = d_ffn

    def input_dimension(self) -> int:
        return self.dimension

    def output_dimension(self) -> int:
        return self.dimension

    forward = _no_forward


# ---- haplotype CNN (dna_sequence_convolution.py) ------------------------------------------------
INITIAL_NUM_CHANNELS = 10


def _conv_len(n, kernel_size=1, stride=1, pad=0, dilation=1):
    return floor(((n + 2 * pad - dilation * (kernel_size - 1) - 1) / stride) + 1)


class DNASequenceConvolution(nn.Module):
    """String-configured Conv1d / MaxPool1d / activation / Flatten / Linear stack; one Sequential
    slot per layer string (dna_sequence_convolution.py:57-99).  ``self.layers`` records the parsed
    geometry the kernel planner consumes."""

    def __init__(self, layer_strings, sequence_length: int):
        super().__init__()
        channels, length = INITIAL_NUM_CHANNELS, sequence_length
        modules, self.layers = [], []
        for s in layer_strings:
            tokens = s.split("/")
            kind = tokens[0]
            kw = {k: int(v) for k, v in (t.split("=") for t in tokens[1:])}
            rec = dict(kind=kind, in_ch=channels, in_len=length, **kw)
            if kind == "convolution":
                extra = set(kw) - {"kernel_size", "out_channels", "stride", "padding", "dilation"}
                if extra or kw.get("stride", 1) != 1 or kw.get("padding", 0) != 0 or kw.get("dilation", 1) != 1:
                    raise NotImplementedError(f"convolution options not supported by the fused kernel: {s}")
                modules.append(nn.Conv1d(in_channels=channels, **kw))
                channels, length = kw["out_channels"], _conv_len(length, kw["kernel_size"])
            elif kind == "pool":
                if set(kw) - {"kernel_size", "stride"}:
                    raise NotImplementedError(f"pool options not supported by the fused kernel: {s}")
                assert length > 1, "pooling a length-1 sequence"
                modules.append(nn.MaxPool1d(**kw))
                length = _conv_len(length, kw["kernel_size"], kw.get("stride", kw["kernel_size"]))
            elif kind == "leaky_relu":
                modules.append(nn.LeakyReLU())
            elif kind == "selu":
                modules.append(nn.SELU())
            elif kind == "flatten":
                modules.append(nn.Flatten())
                channels, length = channels * length, 1
            elif kind == "linear":
                assert length == 1, "linear layer before flatten"
                modules.append(nn.Linear(in_features=channels, **kw))
                channels = kw["out_features"]
            elif kind == "batch_norm":
                raise NotImplementedError("batch_norm in the haplotype CNN is not supported by the fused kernel")
            else:
                raise Exception("unsupported layer_type: " + kind)
            rec.update(out_ch=channels, out_len=length)
            self.layers.append(rec)
        assert length == 1, "data have not been flattened"
        self._output_dimension = channels
        self._model = nn.Sequential(*modules)

    def output_dimension(self):
        return self._output_dimension

    forward = _no_forward


# ---- Euclidean transformation (euclidean_transformation.py) --------------------------------------
class EuclideanTransformation(nn.Module):
    def __init__(self, dimension: int):
        super().__init__()
        self.translation_e = Parameter(torch.rand(dimension))
        self.rotation_ee = orthogonal(nn.Linear(dimension, dimension, bias=False))

    forward = _no_forward


# ---- clustering head (feature_clustering.py, exponentially_modified_gaussian.py) ----------------
class ExponentiallyModifiedGaussian(nn.Module):
    def __init__(self, num_distributions: int):
        super().__init__()
        self.mu_k = Parameter(2 * torch.ones(num_distributions))
        self.sigma_k = Parameter(torch.ones(num_distributions))
        parametrize.register_parametrization(self, "sigma_k", BoundedNumber(MIN_BOUND, MAX_BOUND))
        self.lambda_k = Parameter(torch.ones(num_distributions))
        parametrize.register_parametrization(self, "lambda_k", BoundedNumber(MIN_BOUND, MAX_BOUND))

    forward = _no_forward


class FeatureClustering(nn.Module):
    def __init__(self, feature_dimension: int, num_artifact_clusters: int):
        super().__init__()
        self.feature_dim = feature_dimension
        self.num_artifact_clusters = num_artifact_clusters
        self.nonartifact_stdev_e = Parameter(torch.ones(feature_dimension))
        parametrize.register_parametrization(self, "nonartifact_stdev_e", BoundedNumber(MIN_BOUND, MAX_BOUND))
        self.artifact_directions_ke = Parameter(torch.rand(num_artifact_clusters, feature_dimension))
        parametrize.register_parametrization(self, "artifact_directions_ke", UnitVector())
        self.artifact_emg = ExponentiallyModifiedGaussian(num_artifact_clusters)
        self.artifact_stdev_k = Parameter(torch.ones(num_artifact_clusters))
        parametrize.register_parametrization(self, "artifact_stdev_k", BoundedNumber(MIN_BOUND, MAX_BOUND))
        self.log_cluster_weights_k = Parameter(torch.ones(num_artifact_clusters))
        parametrize.register_parametrization(self, "log_cluster_weights_k", LogWeights())

    forward = _no_forward


# ---- adversarial wrapper (adversarial.py, gradient_reversal/) -----------------------------------
class GradientReversal(nn.Module):
    def __init__(self, alpha: float = 1.0):
        super().__init__()
        self.alpha = alpha

    def set_alpha(self, alpha_new):
        self.alpha = alpha_new

    forward = _no_forward


class Adversarial(nn.Module):
    def __init__(self, wrapped_module: nn.Module, adversarial_strength: float = 1.0):
        super().__init__()
        self.wrapped_module = wrapped_module
        self.gradient_reversal = GradientReversal(alpha=adversarial_strength)

    def set_adversarial_strength(self, new_alpha):
        self.gradient_reversal.set_alpha(new_alpha)

    forward = _no_forward
