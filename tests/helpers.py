"""Shared helpers for the test-suite."""
import numpy as np
import torch

from permutect_b200.architecture.artifact_model import ArtifactModel
from permutect_b200.data.batch import Batch, DownsampledBatch
from permutect_b200.parameters import ModelParameters


def params_from_hp(hp: dict) -> ModelParameters:
    return ModelParameters(read_layers=hp["read_layers"],
                           self_attention_hidden_dimension=hp["self_attention_hidden_dimension"],
                           num_self_attention_layers=hp["num_self_attention_layers"], info_layers=hp["info_layers"],
                           aggregation_layers=hp["aggregation_layers"],
                           num_artifact_clusters=hp["num_artifact_clusters"],
                           calibration_layers=hp["calibration_layers"],
                           ref_seq_layers_strings=hp["ref_seq_layer_strings"], dropout_p=hp["dropout_p"],
                           reweighting_range=hp["reweighting_range"], batch_normalize=hp["batch_normalize"])


def model_from_golden(g, device) -> ArtifactModel:
    n_info = g.inputs["info"].shape[1]
    model = ArtifactModel(params_from_hp(g.hp), num_read_features=61, num_info_features=n_info,
                          haplotypes_length=g.inputs["haplotypes"].shape[1], device=device)
    if g.hp["num_sources"] > 1:
        model.reset_source_predictor(g.hp["num_sources"])
    model.source_predictor.set_adversarial_strength(g.hp["source_adversarial_strength"])
    model.load_state_dict(g.sd)
    return model


def batch_from_raw(raw: dict, device=None) -> Batch:
    """Build a product Batch from oracle-style raw arrays (parent batch; read_indices handled by the caller)."""
    B = len(raw["ref_counts"])
    n_hap, n_info = raw["haplotypes"].shape[1], raw["info"].shape[1]
    ia = np.zeros((B, 16 + n_hap), np.int16)
    fa = np.zeros((B, 6 + n_info), np.float16)
    ia[:, 0], ia[:, 1], ia[:, 2] = raw["ref_counts"], raw["alt_counts"], raw["labels"]
    ia[:, 4] = raw.get("sources", np.zeros(B))
    ia[:, 16:] = raw["haplotypes"]
    fa[:, 6:] = raw["info"]
    batch = Batch.from_arrays(ia, fa, np.ascontiguousarray(raw["reads_u8"]))
    return batch if device is None else batch.copy_to(device)


def golden_batch(g, device):
    """Product batch reproducing a golden case (DownsampledBatch when the case carries read_indices)."""
    raw = g.raw()
    idx = raw.get("read_indices")
    if idx is None:
        return batch_from_raw(raw, device)
    # parent counts are not stored in the fixture; rebuild a parent whose rows are the stored reads
    n_rows = len(raw["reads_u8"])
    parent_raw = dict(raw)
    parent_raw["ref_counts"] = np.zeros_like(raw["ref_counts"])
    parent_raw["alt_counts"] = np.zeros_like(raw["alt_counts"])
    parent_raw["alt_counts"][0] = n_rows            # any split summing to n_rows; only the rows matter
    parent = batch_from_raw(parent_raw, device)
    parent.max_rows_per_variant = int((raw["ref_counts"] + raw["alt_counts"]).max())
    return DownsampledBatch(parent, read_indices=torch.from_numpy(idx),
                            ref_counts=torch.from_numpy(raw["ref_counts"].astype(np.int64)),
                            alt_counts=torch.from_numpy(raw["alt_counts"].astype(np.int64)))
