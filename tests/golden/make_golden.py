"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  The reference is imported
with the three import stubs under oracle/ref_stubs (cyvcf2 / intervaltree / matplotlib are absent
here and unused on the ArtifactModel path).  Outputs: tests/golden/<case>.npz with

    hp            JSON string of the hyper-parameters
    sd/<key>      the reference model's state_dict
    in/<name>     raw batch arrays (reads_u8, read_indices, ref_counts, alt_counts, info, haplotypes, labels, sources)
    out/<name>    BatchOutput fields as computed by the reference (train mode)
    loss/<name>   BatchLosses fields
    grad/<key>    parameter gradients left by total_loss.backward()

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "ref_stubs"), "/root/reference"]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from permutect.architecture.artifact_model import ArtifactModel  # noqa: E402
from permutect.data.batch import Batch, DownsampledBatch  # noqa: E402
from permutect.data.datum import Data, Datum  # noqa: E402
from permutect.parameters import ModelParameters  # noqa: E402
from permutect.utils.enums import Epoch  # noqa: E402

V040 = dict(read_layers=[30, -2, -2, -2], self_attention_hidden_dimension=20, num_self_attention_layers=6,
            info_layers=[20, -2, -2, -2], aggregation_layers=[-2, -2, 10], num_artifact_clusters=4,
            calibration_layers=[20, 20, 20, 20, 10],
            ref_seq_layer_strings=['convolution/kernel_size=3/out_channels=32', 'selu', 'pool/kernel_size=2/stride=1',
                                   'convolution/kernel_size=3/out_channels=32', 'selu', 'pool/kernel_size=1',
                                   'convolution/kernel_size=5/out_channels=32', 'selu', 'pool/kernel_size=2',
                                   'convolution/kernel_size=5/out_channels=32', 'selu', 'pool/kernel_size=2',
                                   'flatten', 'linear/out_features=10'],
            dropout_p=0.0, reweighting_range=0.0, batch_normalize=False)
# hyper-parameters of the reference's own training test (permutect/test/tools/test_train_permutect_model.py:19-35)
SMALL = dict(read_layers=[10, 10, 10], self_attention_hidden_dimension=20, num_self_attention_layers=2,
             info_layers=[10, 10], aggregation_layers=[20, 20, 20], num_artifact_clusters=4,
             calibration_layers=[10, 10, 10],
             ref_seq_layer_strings=["convolution/kernel_size=3/out_channels=64", "pool/kernel_size=2", "leaky_relu",
                                    "flatten", "linear/out_features=10"],
             dropout_p=0.0, reweighting_range=0.3, batch_normalize=False)


def make_params(hp):
    return ModelParameters(read_layers=hp["read_layers"],
                           self_attention_hidden_dimension=hp["self_attention_hidden_dimension"],
                           num_self_attention_layers=hp["num_self_attention_layers"], info_layers=hp["info_layers"],
                           aggregation_layers=hp["aggregation_layers"],
                           num_artifact_clusters=hp["num_artifact_clusters"],
                           calibration_layers=hp["calibration_layers"],
                           ref_seq_layers_strings=hp["ref_seq_layer_strings"], dropout_p=hp["dropout_p"],
                           reweighting_range=hp["reweighting_range"], batch_normalize=hp["batch_normalize"])


def make_datum(rng, ref, alt, label, n_info=71, source=0, appendix_a=False):
    ia = np.zeros(16 + 42, np.int16)
    fa = np.zeros(6 + n_info, np.float16)
    ia[0], ia[1], ia[2] = ref, alt, label
    ia[3] = rng.integers(0, 5)
    if not appendix_a:
        ia[4] = source
    ia[16:] = rng.integers(0, 4, 42) if appendix_a else rng.integers(0, 5, 42)
    fa[6:] = rng.standard_normal(n_info).astype(np.float16)
    reads = rng.integers(0, 256, (ref + alt, 12), dtype=np.uint8)
    return Datum(ia, fa, reads, compressed=True)


def raw_inputs(data, batch, read_indices=None):
    ref_rows = [d.get_ref_reads_re() for d in data]
    alt_rows = [d.get_alt_reads_re() for d in data]
    d = {
        "reads_u8": np.vstack(ref_rows + alt_rows),
        "ref_counts": batch.get(Data.REF_COUNT).numpy().astype(np.int32),
        "alt_counts": batch.get(Data.ALT_COUNT).numpy().astype(np.int32),
        "info": batch.get_info_be().numpy().astype(np.float32),
        "haplotypes": batch.get_haplotypes_bs().numpy().astype(np.int16),
        "labels": batch.get(Data.LABEL).numpy().astype(np.int32),
        "sources": batch.get(Data.SOURCE).numpy().astype(np.int32),
        "decoded_reads": batch.reads_re.numpy().astype(np.float32),
    }
    if read_indices is not None:
        d["read_indices"] = read_indices.numpy().astype(np.int64)
    return d


def perturb(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 0:
                sign = 1.0 if torch.rand((), generator=g) < 0.7 else -1.0
                p.copy_(sign * (0.3 + 0.6 * torch.rand((), generator=g)))
            elif "norm." in name:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.05 * torch.randn(p.shape, generator=g))


def run_case(name, hp, model, data, batch, num_sources=1):
    model.set_epoch_type(Epoch.TRAIN)
    model.zero_grad(set_to_none=True)
    out = model.compute_batch_output(batch)
    ls = model.compute_batch_losses(out, batch)
    ls.total_loss.backward()
    arrays = {"hp": np.array(json.dumps(dict(hp, num_sources=num_sources,
                                             source_adversarial_strength=float(model.source_predictor.gradient_reversal.alpha))))}
    for k, v in model.state_dict().items():
        arrays["sd/" + k] = v.detach().numpy()
    read_indices = getattr(batch, "read_indices", None)
    for k, v in raw_inputs(data, batch, read_indices).items():
        arrays["in/" + k] = v
    for k in ("features_be", "ref_features_be", "logits_b", "logits_bk", "artifact_probs_b", "outlier_binary_logits"):
        arrays["out/" + k] = getattr(out, k).detach().numpy()
    for k in ("supervised_losses_b", "unsupervised_losses_b", "alt_count_losses_b", "source_prediction_losses_b",
              "total_losses_b", "total_loss"):
        arrays["loss/" + k] = getattr(ls, k).detach().numpy()
    for k, p in model.named_parameters():
        arrays["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: B={batch.size()} reads={len(batch.get_reads_re())} logits[:5]={out.logits_b[:5].tolist()} "
          f"loss={float(ls.total_loss):.4f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    cpu = torch.device("cpu")

    # 1. SURVEY.md Appendix A recipe (smoke value logits_b[:5] = -13.2103 -4.3585 -15.9768 -15.3913 -18.4226)
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    model = ArtifactModel(make_params(V040), 61, 71, 42, device=cpu)
    data = [make_datum(rng, int(rng.integers(0, 11)), int(rng.integers(1, 16)), int(rng.integers(0, 3)), appendix_a=True)
            for _ in range(64)]
    batch = Batch(data).copy_to(cpu, torch.float32)
    run_case("v040_seed0_b64", V040, model, data, batch)

    # 2. same model, DownsampledBatch from the reference (quirk Q1: alt indices are not offset)
    torch.manual_seed(11)
    import random
    random.seed(11)
    fr = torch.rand(batch.size())
    fa = torch.rand(batch.size())
    ds = DownsampledBatch(batch, ref_fracs_b=fr, alt_fracs_b=fa)
    run_case("v040_seed0_downsampled", V040, model, data, ds)

    # 3. every parameter perturbed away from init; edge cases: empty ref sets, single alt read,
    #    one variant longer than a 128-row tile, indel haplotype code 4
    torch.manual_seed(1)
    rng = np.random.default_rng(1)
    model = ArtifactModel(make_params(V040), 61, 71, 42, device=cpu)
    perturb(model, 2)
    shapes = [(0, 1), (0, 15), (10, 1), (10, 15), (150, 40), (3, 7), (0, 3), (1, 1)]
    shapes += [(int(rng.integers(0, 11)), int(rng.integers(1, 16))) for _ in range(24)]
    data = [make_datum(rng, r, a, int(rng.integers(0, 3))) for r, a in shapes]
    batch = Batch(data).copy_to(cpu, torch.float32)
    run_case("v040_perturbed_edge", V040, model, data, batch)

    # 4. two sources: source predictor MLP [10,-1,-1,2] with gradient reversal strength 0.4621 (epoch 11)
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    model = ArtifactModel(make_params(V040), 61, 71, 42, device=cpu)
    model.reset_source_predictor(2)
    perturb(model, 4)
    model.source_predictor.set_adversarial_strength(0.4621)
    data = [make_datum(rng, int(rng.integers(0, 11)), int(rng.integers(1, 16)), int(rng.integers(0, 3)),
                       source=int(rng.integers(0, 2))) for _ in range(24)]
    batch = Batch(data).copy_to(cpu, torch.float32)
    run_case("v040_two_sources", V040, model, data, batch, num_sources=2)

    # 5. the reference's own small test hyper-parameters (plain Linear stacks, 64-channel CNN with flatten > 1)
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    model = ArtifactModel(make_params(SMALL), 61, 71, 42, device=cpu)
    perturb(model, 6)
    data = [make_datum(rng, int(rng.integers(0, 11)), int(rng.integers(1, 16)), int(rng.integers(0, 3)))
            for _ in range(24)]
    batch = Batch(data).copy_to(cpu, torch.float32)
    run_case("small_hp", SMALL, model, data, batch)


if __name__ == "__main__":
    main()
