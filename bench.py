#!/usr/bin/env python
"""Benchmark of the ArtifactModel hot path (BASELINE.json metric: ArtifactModel variants/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--variants V] [--impl reference]

A step is one pass of the fused forward (compute_batch_output: every BatchOutput field) over one
shard of synthetic WGS-shaped variants (10M variants sharded by variant
over 8 GPUs -> 1.25M variants per GPU; weak scaling, per-GPU shard fixed).  `value` is measured with
the shard resident in HBM; `e2e` goes through the public API with pinned HOST buffers, H2D of the
compressed shard and D2H of the posterior records (int16 record + fp16-rounded logit + embedding, 180 B per variant)
inside the timed region.  `train` (when the backward kernels
are built) is one forward + losses + backward + clip + AdamW step per batch.

`--impl reference` times the UNMODIFIED reference (oracle/_ref, kind "reference"; the oracle port only when the
recipe oracle/build_ref.py has not been run) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

V040 = dict(read_layers=[30, -2, -2, -2], self_attention_hidden_dimension=20, num_self_attention_layers=6,
            info_layers=[20, -2, -2, -2], aggregation_layers=[-2, -2, 10], num_artifact_clusters=4,
            calibration_layers=[20, 20, 20, 20, 10],
            ref_seq_layer_strings=['convolution/kernel_size=3/out_channels=32', 'selu', 'pool/kernel_size=2/stride=1',
                                   'convolution/kernel_size=3/out_channels=32', 'selu', 'pool/kernel_size=1',
                                   'convolution/kernel_size=5/out_channels=32', 'selu', 'pool/kernel_size=2',
                                   'convolution/kernel_size=5/out_channels=32', 'selu', 'pool/kernel_size=2',
                                   'flatten', 'linear/out_features=10'],
            dropout_p=0.0, reweighting_range=0.0, batch_normalize=False, num_sources=1)

# algorithmic work per unit, forward: MAC = 2 FLOP, unpadded dims
FLOP_PER_READ = 66260
FLOP_PER_ALT_READ = 500
FLOP_PER_VARIANT = 286424
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal B200 FP32 pipe
WORKLOAD = "synthetic WGS-scale inference (SURVEY §8d config 3: 10M variants / 8 GPUs)"


def ncu_traffic(precision, n_variants):
    """DRAM bytes of one launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full`
    capture summarised in profiles/ncu_traffic.json (written by profiles/ncu_traffic.py from the .ncu-rep), scaled by variants
    when the capture ran another shard size.  None when no capture of this precision mode is committed."""
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    rec = json.load(open(path)).get(precision)
    if not rec:
        return None, None
    return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * n_variants / rec["variants"], rec.get("source")


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, bf16=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, extra_rows=()):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.rows = list(extra_rows) + self.rows
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


LIBRARY_KERNEL_PREFIXES = ("pmt::", "tc::", "cnntc::", "loss::", "optim::", "expm::", "post::", "void pmt::", "void tc::")


def count_library_launches(fn):
    """Kernels of libpermutect_b200 launched by one call of fn(), counted from a CUPTI trace (torch.profiler) taken OUTSIDE
    every timed region.  Returns (library kernels, all kernels) or (None, None) when CUPTI is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name
                 and "Memset" not in e.name]
        mine = [n for n in names if any(tag in n for tag in ("pmt::", "tc::", "cnntc::", "loss::", "optim::", "expm::", "post::"))]
        return (len(mine), len(names)) if names else (None, None)
    except Exception:   # noqa: BLE001 - no CUPTI on this box: the claim is then absent, not invented
        return None, None


def make_model(device):
    from helpers import params_from_hp
    from permutect_b200.architecture.artifact_model import ArtifactModel
    torch.manual_seed(0)
    model = ArtifactModel(params_from_hp(V040), 61, 71, 42, device=device)
    # move every scalar away from its init so no term is numerically hidden
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() == 0:
                p.copy_(0.05 + 0.1 * torch.rand((), generator=g))
    return model


def oracle_inputs(ia, fa, reads):
    return dict(reads_u8=reads, read_indices=None, ref_counts=ia[:, 0], alt_counts=ia[:, 1],
                info=fa[:, 6:].astype(np.float32), haplotypes=ia[:, 16:], labels=ia[:, 2], sources=ia[:, 4])


def time_oracle(model_sd, n_variants, seed, budget_s=20.0, train=False):
    """CPU restatement of the reference on all host cores, bounded sample.  Returns variants/s."""
    from oracle import artifact_oracle as orc
    from permutect_b200.synthetic import make_wgs_arrays
    torch.set_num_threads(os.cpu_count())
    ia, fa, reads = make_wgs_arrays(n_variants, seed=seed)
    raw = oracle_inputs(ia, fa, reads)
    sd = {k: v.detach().cpu() for k, v in model_sd.items()}
    times = []
    t_end = time.perf_counter() + budget_s
    it = 0
    while True:
        t0 = time.perf_counter()
        if train:
            orc.loss_and_grads(sd, V040, raw, [k for k in sd if sd[k].dtype.is_floating_point and "base" not in k])
        else:
            with torch.no_grad():
                orc.forward(sd, V040, raw)
        dt = time.perf_counter() - t0
        if it > 0:
            times.append(dt)
        it += 1
        if (time.perf_counter() > t_end and len(times) >= 2) or len(times) >= 10:
            break
    return n_variants / float(np.median(times)), len(times)


def reference_setup(state_dict, sizes, seed):
    """The UNMODIFIED reference (oracle/_ref, oracle/build_ref.py) on the CPU: its ArtifactModel with the bench model's weights
    and reference Batch objects (collated from its own Datum class) of the same synthetic distribution, one per size."""
    from oracle import reference
    from permutect_b200.synthetic import make_wgs_arrays
    reference.load()
    import permutect.data.batch as rb
    import permutect.data.datum as rd
    from permutect.architecture.artifact_model import ArtifactModel as RefModel
    from permutect.parameters import ModelParameters as RefParams
    hp = V040
    params = RefParams(read_layers=hp["read_layers"], self_attention_hidden_dimension=hp["self_attention_hidden_dimension"],
                       num_self_attention_layers=hp["num_self_attention_layers"], info_layers=hp["info_layers"],
                       aggregation_layers=hp["aggregation_layers"], num_artifact_clusters=hp["num_artifact_clusters"],
                       calibration_layers=hp["calibration_layers"], ref_seq_layers_strings=hp["ref_seq_layer_strings"],
                       dropout_p=hp["dropout_p"], reweighting_range=hp["reweighting_range"], batch_normalize=hp["batch_normalize"])
    model = RefModel(params, 61, 71, 42, device=torch.device("cpu"))
    model.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()})
    batches = {}
    for size in sizes:
        ia, fa, reads = make_wgs_arrays(size, seed=seed)
        ref_c, alt_c = ia[:, 0].astype(int), ia[:, 1].astype(int)
        ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
        total_ref = ref_off[-1]
        data = [rd.Datum(ia[v], fa[v], np.vstack((reads[ref_off[v]:ref_off[v + 1]],
                                                  reads[total_ref + alt_off[v]:total_ref + alt_off[v + 1]])), compressed=True)
                for v in range(size)]
        batches[size] = rb.Batch(data)
    return model, batches


def time_reference(model, batch, n_variants, train, budget_s, min_steps=2, max_steps=10, warm=1):
    """Median step time of the reference's own code path: compute_batch_output (inference) or compute_batch_output +
    compute_batch_losses + misc_utils.backpropagate (one optimiser step, training/model_training.py:151-165)."""
    from permutect.misc_utils import backpropagate
    from permutect.utils.enums import Epoch as RefEpoch
    torch.set_num_threads(os.cpu_count())
    opt = None
    if train:
        model.set_epoch_type(RefEpoch.TRAIN)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    else:
        model.set_epoch_type(RefEpoch.VALID)
    times = []
    t_end = time.perf_counter() + budget_s
    it = 0
    while True:
        t0 = time.perf_counter()
        if train:
            out = model.compute_batch_output(batch)
            losses = model.compute_batch_losses(out, batch)
            backpropagate(opt, losses.total_loss, params_to_clip=model.parameters())
        else:
            with torch.inference_mode():
                model.compute_batch_output(batch)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
        it += 1
        if (time.perf_counter() > t_end and len(times) >= min_steps) or len(times) >= max_steps:
            break
    med = float(np.median(times))
    return {"variants_per_s": n_variants / med, "ms_per_step": 1e3 * med, "steps": len(times), "batch_variants": n_variants}


def reference_available():
    from oracle import reference
    return reference.available()


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path with every host thread, on bounded samples of the
    workload -- the reference tools' default batch (64 variants: parameters.py:214, filter_variants.py:81) and a large one."""
    if rank != 0:
        return
    model = make_model(torch.device("cpu"))
    steps, warm = max(args.steps, 1), args.warmup
    big = 8192
    if reference_available():
        kind = "reference"
        ref_model, batches = reference_setup(model.state_dict(), [64, big, 2048], seed=3000)
        per_step = max(8.0, 120.0 / (warm + steps))
        infer_big = time_reference(ref_model, batches[big], big, False, budget_s=90.0, min_steps=1, max_steps=steps, warm=max(warm, 1))
        infer_64 = time_reference(ref_model, batches[64], 64, False, budget_s=10.0, max_steps=50, warm=3)
        train_big = time_reference(ref_model, batches[2048], 2048, True, budget_s=30.0, min_steps=1, max_steps=5)
        train_64 = time_reference(ref_model, batches[64], 64, True, budget_s=10.0, max_steps=30, warm=2)
        how = "unmodified reference from oracle/_ref (ArtifactModel.compute_batch_output on torch CPU, all host threads)"
    else:
        kind = "port"
        v, n_it = time_oracle(model.state_dict(), big, seed=3000, budget_s=60.0)
        infer_big = {"variants_per_s": v, "ms_per_step": 1e3 * big / v, "steps": n_it, "batch_variants": big}
        v, n_it = time_oracle(model.state_dict(), 64, seed=3000, budget_s=10.0)
        infer_64 = {"variants_per_s": v, "ms_per_step": 1e3 * 64 / v, "steps": n_it, "batch_variants": 64}
        v, n_it = time_oracle(model.state_dict(), 2048, seed=3001, budget_s=20.0, train=True)
        train_big = {"variants_per_s": v, "ms_per_step": 1e3 * 2048 / v, "steps": n_it, "batch_variants": 2048}
        v, n_it = time_oracle(model.state_dict(), 64, seed=3001, budget_s=10.0, train=True)
        train_64 = {"variants_per_s": v, "ms_per_step": 1e3 * 64 / v, "steps": n_it, "batch_variants": 64}
        how = "oracle port of the reference (oracle/_ref not built)"
    best = infer_big if infer_big["variants_per_s"] >= infer_64["variants_per_s"] else infer_64
    v = best["variants_per_s"]
    sample_txt = (f"{best['batch_variants']} WGS-shaped variants per step (same synthetic distribution as the shard), "
                  f"median of {best['steps']} steps; {how}")
    print(json.dumps({
        "impl": "reference", "metric": "artifact_model_inference_variants_per_sec", "value": v, "unit": "variants/s",
        "n_gpus": args.gpus, "steps": best["steps"], "warmup": warm, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision": "fp32 (torch CPU, all host threads)", "variants_per_gpu": args.variants,
                   "hyperparameters": "artifact-model-v0.4.0", "sample_variants_per_step": best["batch_variants"]},
        "cpu_baseline": {"value": v, "unit": "variants/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample_txt},
        "e2e": {"value": v, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "inference": {"batch_64": infer_64, f"batch_{big}": infer_big},
        "train": {"metric": "artifact_model_training_variants_per_sec", "batch_64": train_64, "batch_2048": train_big,
                  "step": "compute_batch_output + compute_batch_losses + misc_utils.backpropagate (clip 1.0 + AdamW)"},
    }))


def run_training(args, model, ia, fa, reads, dev, world, barrier):
    """every step draws a DownsampledBatch of one resident parent chunk on the device and
    takes one optimiser step.  Returns the `train` record (rank-local timing reduced with MAX by the caller's barrier)."""
    import torch.distributed as dist
    from permutect_b200.data.batch import Batch, DownsampledBatch
    from permutect_b200.engine import function as engine
    from permutect_b200.training.step import make_optimizer, train_step
    from permutect_b200.utils.enums import Epoch

    bt = min(args.train_batch, len(ia))
    n_chunks = max(1, min(8, len(ia) // bt))
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = int(ref_off[-1])
    parents = []
    for c in range(n_chunks):
        v0, v1 = c * bt, (c + 1) * bt
        sub = np.concatenate((reads[ref_off[v0]:ref_off[v1]], reads[total_ref + alt_off[v0]:total_ref + alt_off[v1]]))
        parents.append(Batch.from_arrays(ia[v0:v1], fa[v0:v1], sub).copy_to(dev))
    g = torch.Generator().manual_seed(7)
    fracs = [(0.3 + 0.7 * torch.rand(bt, generator=g), 0.3 + 0.7 * torch.rand(bt, generator=g)) for _ in range(n_chunks)]
    fracs = [(a.to(dev), b.to(dev)) for a, b in fracs]
    model.set_epoch_type(Epoch.TRAIN)
    opt = make_optimizer(model, learning_rate=1e-3, weight_decay=0.01)
    prof = engine.ProfileEvents(dev)

    def step(i, profile=False):
        parent = parents[i % n_chunks]
        batch = DownsampledBatch(parent, fracs[i % n_chunks][0], fracs[i % n_chunks][1], seed=1000 + i)
        out = model.compute_batch_output(batch)
        losses = model.compute_batch_losses(out, batch)
        if profile:
            prof.arm()           # arm only around the backward so the events bracket reads_backward_kernel
        from permutect_b200.training.step import backpropagate
        backpropagate(opt, losses.total_loss, params_to_clip=model.parameters())
        if profile:
            prof.disarm()
        return losses

    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        losses = step(args.warmup + i, profile=True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    mine, total = count_library_launches(lambda: step(args.warmup + args.steps))
    # the data-parallel exchange alone: all-reduce of the flat gradient, CUDA events, max over ranks
    allreduce_us = None
    if world > 1:
        buf = torch.zeros_like(opt.flat_grad)
        for _ in range(5):
            dist.all_reduce(buf)
        barrier()
        ev0.record()
        for _ in range(20):
            dist.all_reduce(buf)
        ev1.record()
        barrier()
        ta = torch.tensor([ev0.elapsed_time(ev1) / 20 * 1e3], device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_us = float(ta.item())
    model.set_epoch_type(Epoch.VALID)
    from permutect_b200.engine import library as pmt_lib
    return {"metric": "artifact_model_training_variants_per_sec", "value": bt * world / (ms_max / 1e3), "unit": "variants/s",
            "ms_per_step": ms_max, "batch_variants_per_gpu": bt, "last_loss_per_variant": float(losses.total_loss.detach()) / bt,
            "backward_kernel_ms": prof.mean_ms(), "gpu_launches_per_step": mine, "all_kernels_per_step": total,
            "gpu_launches": mine * args.steps if mine is not None else None,
            "allreduce_us": allreduce_us, "allreduce_bytes": 4 * opt.flat_grad.numel(),
            "backward": ("read path: tcgen05 (recompute + TF32 data / weight gradient MMAs, pmt_tc_bwd.cu); haplotype CNN: tcgen05 "
                         "recompute with saved activations + warp-level TF32 MMAs (pmt_cnn_bwd.cu)" if pmt_lib.get_precision() != "fp32"
                         else "FP32 SIMT"),
            "step": "device DownsampledBatch + compute_batch_output (pmt_forward_train: the forward keeps the operand panels, no "
                    "recompute in the backward) + compute_batch_losses (fused loss head) + backward + flat grad all-reduce + "
                    "clip(1.0) + AdamW (FlatAdamW)"}


def run_config3_single_gpu(args, model, ia, fa, reads, dev, total_variants=10_000_000, steps=3):
    """BASELINE config 3 as ONE device-resident batch on ONE GPU (the 10 M variants fit a B200 many times over): the
    rank's shard repeated until it holds ``total_variants`` -- the same synthetic read sets eight times, inputs far larger
    than L2 either way.  Reported beside the headline (whose per-GPU shard is what 8 GPUs each get of this config)."""
    from permutect_b200.data.batch import Batch
    k = max(1, total_variants // len(ia))
    total_ref = int(ia[:, 0].astype(np.int64).sum())
    big_reads = np.concatenate([reads[:total_ref]] * k + [reads[total_ref:]] * k)
    batch = Batch.from_arrays(np.tile(ia, (k, 1)), np.tile(fa, (k, 1)), big_reads).copy_to(dev)
    del big_reads
    batch.offsets()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.inference_mode():
        first = model.compute_batch_output(batch).logits_b[:len(ia)].clone()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            out = model.compute_batch_output(batch)
        ev1.record()
        torch.cuda.synchronize()
        # every copy of the shard gets the shard's logits (tiles never straddle the copies' boundaries differently: compare)
        same = bool(torch.equal(out.logits_b[-len(ia):], first))
    ms = ev0.elapsed_time(ev1) / steps
    n = k * len(ia)
    return {"workload": "BASELINE config 3 on one GPU: 10 M WGS-shaped variants in one device-resident batch",
            "variants": n, "reads": int(k * len(reads)), "ms_per_step": ms, "variants_per_s": n / (ms / 1e3), "steps": steps,
            "data": f"the {len(ia)}-variant shard repeated {k}x", "last_copy_equals_first": same}


def run_panel(args, model, dev, world, rank, barrier):
    """BASELINE config 5 (high-depth panel stress, sets of ~2 000 reads): a bounded sample of it per GPU, inference and
    training, through the long-set kernels.  Reported beside the headline, not as it."""
    import torch.distributed as dist
    from permutect_b200.data.batch import Batch, DownsampledBatch
    from permutect_b200.synthetic import make_panel_arrays
    from permutect_b200.training.step import make_optimizer, train_step
    from permutect_b200.utils.enums import Epoch

    n = args.panel_variants
    ia, fa, reads = make_panel_arrays(n, seed=1000 * 5 + rank)
    parent = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, warm, steps):
        for i in range(warm):
            fn(i)
        barrier()
        ev0.record()
        for i in range(steps):
            fn(warm + i)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from permutect_b200.engine import library as pmt_lib
    model.set_epoch_type(Epoch.VALID)
    with torch.inference_mode():
        infer_ms = timed(lambda i: model.compute_batch_output(parent), 2, 3)
        # the same sample through the FP32 long-set kernel (what every mode used before the sets went onto the tile pipeline)
        simt_ms = None
        if pmt_lib.get_precision() != "fp32":
            os.environ["PMT_LONG_SIMT"] = "1"
            try:
                simt_ms = timed(lambda i: model.compute_batch_output(parent), 1, 2)
            finally:
                os.environ.pop("PMT_LONG_SIMT", None)
    # a larger inference sample: 32 768 sets per GPU (a 4 096-set sample repeated 8 times, ~67 M reads)
    big = None
    if args.panel_big_variants > 0:
        base_n = min(4096, args.panel_big_variants)
        k = max(1, args.panel_big_variants // base_n)
        bia, bfa, breads = make_panel_arrays(base_n, seed=1000 * 7 + rank)
        tref = int(bia[:, 0].astype(np.int64).sum())
        big_batch = Batch.from_arrays(np.tile(bia, (k, 1)), np.tile(bfa, (k, 1)),
                                      np.concatenate([breads[:tref]] * k + [breads[tref:]] * k)).copy_to(dev)
        with torch.inference_mode():
            big_ms = timed(lambda i: model.compute_batch_output(big_batch), 1, 3)
        big = {"variants_per_gpu": k * base_n, "reads_per_gpu": k * len(breads), "inference_ms": big_ms,
               "inference_variants_per_s": k * base_n * world / (big_ms / 1e3), "inference_reads_per_s": k * len(breads) * world / (big_ms / 1e3),
               "data": f"a {base_n}-set sample repeated {k}x"}
        del big_batch
        torch.cuda.empty_cache()
    model.set_epoch_type(Epoch.TRAIN)
    opt = make_optimizer(model, learning_rate=1e-3, weight_decay=0.01)
    frac = torch.full((n,), 0.8, device=dev)
    kept = []

    def step(i):
        batch = DownsampledBatch(parent, frac, frac, seed=500 + i)
        train_step(model, batch, opt)
        if not kept:
            kept.append(int(sum(int(c.sum()) for c in batch.counts())))

    train_ms = timed(step, 1, 2)
    model.set_epoch_type(Epoch.VALID)
    return {"workload": "high-depth panel stress (SURVEY \u00a78d config 5), bounded sample", "variants_per_gpu": n,
            "reads_per_gpu": len(reads), "mean_reads_per_variant": len(reads) / n,
            "inference_variants_per_s": n * world / (infer_ms / 1e3), "inference_reads_per_s": len(reads) * world / (infer_ms / 1e3),
            "inference_ms": infer_ms, "train_variants_per_s": n * world / (train_ms / 1e3),
            "train_reads_per_s": kept[0] * world / (train_ms / 1e3), "train_ms": train_ms, "train_reads_per_step": kept[0],
            "inference_ms_fp32_long_kernel": simt_ms,
            "inference_reads_per_s_fp32_long_kernel": len(reads) * world / (simt_ms / 1e3) if simt_ms else None,
            "inference_large_sample": big,
            "kernels": ("inference: reads_forward_tc_kernel<.., LONG> (tcgen05; a set is cut into single-side tiles of 128 reads that are in "
                        "flight together and exchange their mean-field / set-sum partials through global memory); training: "
                        "reads_forward_long_kernel / reads_backward_long_kernel (FP32 SIMT, sets walked in chunks of 128 reads)"
                        if pmt_lib.get_precision() != "fp32" else
                        "reads_forward_long_kernel / reads_backward_long_kernel (FP32 SIMT, sets walked in chunks of 128 reads)")}


def run_small_batch_training(model, dev, batch_variants=64, steps=100):
    """The reference's default training batch (parameters.py:214: 64 variants): the step is bound by the host side
    (autograd over the parametrised tensors, ctypes, launches), not by the kernels."""
    from permutect_b200.data.batch import Batch, DownsampledBatch
    from permutect_b200.synthetic import make_wgs_arrays
    from permutect_b200.training.step import make_optimizer, train_step
    from permutect_b200.utils.enums import Epoch
    model.set_epoch_type(Epoch.TRAIN)
    opt = make_optimizer(model, learning_rate=1e-3, weight_decay=0.01)
    parent = Batch.from_arrays(*make_wgs_arrays(batch_variants, seed=4000)).copy_to(dev)
    frac = torch.full((batch_variants,), 0.8, device=dev)
    for i in range(10):
        train_step(model, DownsampledBatch(parent, frac, frac, seed=i), opt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        train_step(model, DownsampledBatch(parent, frac, frac, seed=100 + i), opt)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    # the same step replayed from a CUDA graph (engine/graphs.py): the downsampling draw stays eager (fresh seed per step)
    from permutect_b200.engine.graphs import GraphedInference, GraphedTrainStep
    graphed = GraphedTrainStep(model, opt, DownsampledBatch(parent, frac, frac, seed=1))
    for i in range(10):
        graphed(DownsampledBatch(parent, frac, frac, seed=i))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        graphed(DownsampledBatch(parent, frac, frac, seed=100 + i))
    torch.cuda.synchronize()
    ms_graph = 1e3 * (time.perf_counter() - t0) / steps
    model.set_epoch_type(Epoch.VALID)
    # inference at the same batch size: eager call and graph replay
    infer = GraphedInference(model, parent)

    def per_call(fn, n=300):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1e6 * (time.perf_counter() - t0) / n

    with torch.inference_mode():
        us_eager = per_call(lambda: model.compute_batch_output(parent))
    us_graph = per_call(lambda: infer(parent))
    return {"batch_variants": batch_variants, "ms_per_step": ms, "variants_per_s": batch_variants / (ms / 1e3), "steps": steps,
            "graphed_ms_per_step": ms_graph, "graphed_variants_per_s": batch_variants / (ms_graph / 1e3),
            "inference_us_per_call": us_eager, "inference_graphed_us_per_call": us_graph,
            "inference_variants_per_s": batch_variants / (min(us_eager, us_graph) / 1e6),
            "timing": "wall clock around the loop with a synchronize on both sides; eager = host-bound, graphed = bound by the "
                      "latency of the kernel chain (one tile walks 24 layers forward, 48 backward)"}


def run_posterior(n, dev):
    """SURVEY §8 f3 (inference half): PosteriorModel.posterior_probabilities_bc over n synthetic posterior records."""
    from permutect_b200.architecture.posterior_model import PosteriorBatch, PosteriorModel
    g = torch.Generator(device=dev).manual_seed(9)
    it = torch.zeros((n, 16 + 42), dtype=torch.int16, device=dev)
    depth = torch.randint(8, 300, (n,), generator=g, device=dev)
    it[:, 3] = torch.randint(0, 5, (n,), generator=g, device=dev)
    it[:, 5] = depth
    it[:, 6] = torch.clamp((depth * torch.rand(n, generator=g, device=dev) * 0.6).long(), min=1)
    it[:, 7] = torch.randint(0, 200, (n,), generator=g, device=dev)
    it[:, 8] = (it[:, 7].float() * 0.03 * torch.rand(n, generator=g, device=dev)).long()
    it[:, 16:] = torch.randint(0, 4, (n, 42), generator=g, device=dev)
    ft = torch.zeros((n, 16), dtype=torch.float32, device=dev)
    ft[:, 0] = -30 * torch.rand(n, generator=g, device=dev)
    ft[:, 1] = -5 * torch.rand(n, generator=g, device=dev)
    ft[:, 2] = 10 ** (-4 * torch.rand(n, generator=g, device=dev) - 0.3)
    ft[:, 3] = 0.05 + 0.45 * torch.rand(n, generator=g, device=dev)
    ft[:, 4] = 0.05 + 0.45 * torch.rand(n, generator=g, device=dev)
    ft[:, 5] = 8 * torch.randn(n, generator=g, device=dev)
    model = PosteriorModel(-10.0, -10.0, device=dev)
    batch = PosteriorBatch(it, ft)
    for _ in range(3):
        model.posterior_probabilities_bc(batch)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(5):
        model.posterior_probabilities_bc(batch)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    return {"what": "PosteriorModel.posterior_probabilities_bc (pmt_posterior_log_posteriors), records resident", "variants": n,
            "ms": ms, "variants_per_s": n / (ms / 1e3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", type=int, default=1_250_000, help="variants per GPU shard")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-batch", type=int, default=65536, help="variants per optimiser step")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--panel-big-variants", type=int, default=32768, help="sets per GPU of the larger panel inference sample (0: skip)")
    ap.add_argument("--no-config3", action="store_true", help="skip the 10 M-variants-on-one-GPU record (N = 1 only)")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="arithmetic of the forward's dense layers: tf32x3 = split-precision TF32 on tcgen05 (fp32 parity, "
                         "logits within 1e-3 of the reference), fp32 = FP32 SIMT, tf32 = plain TF32 (looser, stated tolerance); "
                         "the backward of the tensor-core modes runs on tcgen05 in TF32 (stated gradient tolerance), fp32's on the FP32 pipe")
    ap.add_argument("--e2e-batches", type=int, default=8, help="host batches the shard is delivered in for the e2e leg")
    ap.add_argument("--panel-variants", type=int, default=1024,
                    help="variants per GPU of the high-depth panel sample (config 5, ~2 050 reads each); 0 skips it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from permutect_b200.data.batch import Batch
    from permutect_b200.engine import function as engine
    from permutect_b200.synthetic import make_wgs_arrays
    from permutect_b200.utils.enums import Epoch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = None
    if world > 1:
        # one process per GPU: stay on the socket the GPU hangs off, so the pinned host batches of the e2e leg do not cross
        # the inter-socket link (at N = 1 the process keeps every core: the CPU baseline runs in it)
        from permutect_b200.training.distributed import bind_to_gpu_numa_node
        numa_cores = bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from permutect_b200.engine import library as pmt_lib
    model = make_model(dev)
    model.set_epoch_type(Epoch.VALID)
    pmt_lib.set_precision(args.precision)
    ia, fa, reads = make_wgs_arrays(args.variants, seed=1000 * 3 + rank)
    n_reads = len(reads)
    n_alt = int(ia[:, 1].sum())
    dev_batch = Batch.from_arrays(ia, fa, reads).copy_to(dev)
    dev_batch.offsets()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)

    # ---- device-resident throughput -------------------------------------------------------------------
    with torch.inference_mode():
        for _ in range(args.warmup):
            model.compute_batch_output(dev_batch)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof = engine.ProfileEvents(dev)             # CUDA events around the dominant kernel, same stream
        ev0.record()
        for _ in range(args.steps):
            prof.arm()
            model.compute_batch_output(dev_batch)
        ev1.record()
        barrier()
        sampler.stop()
        step_ms = ev0.elapsed_time(ev1) / args.steps
        read_kernel_ms = prof.mean_ms()
    t = torch.tensor([step_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms_max = float(t.item())
    value = args.variants * world / (step_ms_max / 1e3)
    with torch.inference_mode():
        fwd_launches, _ = count_library_launches(lambda: model.compute_batch_output(dev_batch))

    # ---- end to end through the public API: pinned host batches -> prefetch_generator (H2D on a side stream) ->
    #      compute_batch_output -> posterior records -> D2H; every byte of the shard crosses PCIe inside the timed region ----
    from permutect_b200.data.prefetch_generator import prefetch_generator
    nb = max(1, args.e2e_batches)
    ref_c, alt_c = ia[:, 0].astype(np.int64), ia[:, 1].astype(np.int64)
    ref_off, alt_off = np.concatenate(([0], np.cumsum(ref_c))), np.concatenate(([0], np.cumsum(alt_c)))
    total_ref = int(ref_off[-1])
    bounds = [args.variants * i // nb for i in range(nb + 1)]
    host_batches = []
    for v0, v1 in zip(bounds[:-1], bounds[1:]):
        sub = np.concatenate((reads[ref_off[v0]:ref_off[v1]], reads[total_ref + alt_off[v0]:total_ref + alt_off[v1]]))
        host_batches.append(Batch.from_arrays(ia[v0:v1], fa[v0:v1], sub).pin_memory())
    h2d = sum(b.h2d_bytes() for b in host_batches)
    from permutect_b200.tools.filter_variants import generate_posterior_arrays
    d2h_seen = [0]

    def e2e_passes(k):
        # the call filter_variants makes (tools/filter_variants.py:generate_posterior_arrays): per batch the posterior records
        # (int16 row + fp16-rounded logit + embedding, fp32) come back to host memory.  ONE call over a loader that delivers
        # the shard k times, as the tool makes one call over its whole dataset: the pipeline's fill (first H2D) and drain
        # (last D2H) are inside the timed region once, not once per step.
        loader = (b for _ in range(k) for b in host_batches)
        n = 0
        for int_rec, float_rec in generate_posterior_arrays(loader, model, dev):
            n += int_rec.nbytes + float_rec.nbytes
        d2h_seen[0] = n // k

    with torch.inference_mode():
        e2e_passes(2)
        barrier()
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        ev0.record()
        e2e_passes(args.steps)
        ev1.record()
        barrier()
        e2e_ms = ev0.elapsed_time(ev1) / args.steps
        # ---- the same call fed from the dataset's maps: MemoryMappedBatches cuts each batch out of the (page-cache resident)
        #      int16 / fp16 / compressed-reads arrays in dataset order and stages it into pinned memory with a few copy threads,
        #      INSIDE the timed region; the ref / alt regrouping is a gather-index array written on the device ----
        from permutect_b200.data.reads_dataset import MemoryMappedBatches
        per_variant = ref_c + alt_c
        vv = np.repeat(np.arange(args.variants), per_variant)
        kk = np.arange(int(per_variant.sum())) - np.repeat(np.concatenate(([0], np.cumsum(per_variant)))[:-1], per_variant)
        reads_dataset_order = reads[np.where(kk < ref_c[vv], ref_off[vv] + kk, total_ref + alt_off[vv] + (kk - ref_c[vv]))]
        del vv, kk
        staging_threads = max(2, min(8, (os.cpu_count() or 8) // (2 * max(1, world))))
        maps_loader = MemoryMappedBatches(ia, fa, reads_dataset_order, batch_size=(args.variants + nb - 1) // nb, pin_memory=True,
                                          staging_threads=staging_threads, prefetch=2)
        # and with the dataset loaded into page-locked memory once (load_into_pinned_memory: the load every run starts with, into
        # cudaHostAlloc memory instead of pageable memory): a batch is then three zero-copy slices
        from permutect_b200.data.reads_dataset import load_into_pinned_memory
        reg_loader = MemoryMappedBatches(load_into_pinned_memory(ia), load_into_pinned_memory(fa), load_into_pinned_memory(reads_dataset_order),
                                         batch_size=(args.variants + nb - 1) // nb, pin_memory="register", prefetch=2)

        def maps_passes(k, loader):
            n = 0
            for int_rec, float_rec in generate_posterior_arrays((b for _ in range(k) for b in loader), model, dev):
                n += len(int_rec)
            return n

        assert maps_passes(2, maps_loader) == 2 * args.variants
        barrier()
        ev0.record()
        maps_passes(args.steps, maps_loader)
        ev1.record()
        barrier()
        maps_ms = ev0.elapsed_time(ev1) / args.steps
        assert maps_passes(2, reg_loader) == 2 * args.variants
        barrier()
        ev0.record()
        maps_passes(args.steps, reg_loader)
        ev1.record()
        barrier()
        reg_ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler2.stop(extra_rows=sampler.rows)      # samples of both timed regions (resident steps, e2e pipeline)
    tm = torch.tensor([maps_ms, reg_ms], device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    e2e_maps_value = args.variants * world / (float(tm[0].item()) / 1e3)
    e2e_reg_value = args.variants * world / (float(tm[1].item()) / 1e3)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.variants * world / (float(t.item()) / 1e3)

    # ---- parity at full size (outside every timed region): the tensor-core mode against the FP32 SIMT kernels ----
    parity = None
    if args.precision != "fp32":
        with torch.inference_mode():
            got = model.compute_batch_output(dev_batch).logits_b.clone()
            pmt_lib.set_precision("fp32")
            want = model.compute_batch_output(dev_batch).logits_b
            pmt_lib.set_precision(args.precision)
            diff = (got - want).abs()
            parity = {"against": "FP32 SIMT kernels, same shard", "max_abs_logit_diff": float(diff.max()),
                      "mean_abs_logit_diff": float(diff.mean()), "n_over_1e-3": int((diff > 1e-3).sum()),
                      "sign_flips": int(((got > 0) != (want > 0)).sum()),
                      "fp16_rounding_changes": int((got.half() != want.half()).sum()), "n": int(got.numel())}

    # ---- decision identity against the ORACLE chain (outside every timed region): sign flips, fp16 changes of the cached
    #      logit, posterior call flips on a 65 536-variant sample whose logits straddle 0 (tests/test_decision_identity_gpu.py) ----
    if parity is not None and rank == 0 and not args.no_cpu_baseline:
        from test_decision_identity_gpu import decision_counts, straddling_model
        counts, _ = decision_counts(straddling_model(dev), dev, 65536, seed=9100)
        parity["oracle_chain"] = dict(counts, against="oracle logits -> fp16 -> posterior oracle vs this path -> pmt_pack_posterior -> "
                                                      "pmt_posterior_log_posteriors; bench model with the calibration spread narrowed")
        pmt_lib.set_precision(args.precision)

    # ---- training: downsample -> forward -> losses -> backward -> (all-reduce) -> clip -> AdamW ----------------
    train = None
    if not args.no_train:
        train = run_training(args, model, ia, fa, reads, dev, world, barrier)

    config3 = None
    if world == 1 and not args.no_config3:
        config3 = run_config3_single_gpu(args, model, ia, fa, reads, dev)
        torch.cuda.empty_cache()

    panel, small_batch = None, None
    if args.panel_variants > 0 and not args.no_train:
        panel = run_panel(args, model, dev, world, rank, barrier)
        small_batch = run_small_batch_training(model, dev)     # every rank: the optimiser step all-reduces the gradient
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    traffic, traffic_source = ncu_traffic(args.precision, args.variants)
    flops = FLOP_PER_READ * n_reads + FLOP_PER_ALT_READ * n_alt        # the read kernel's algorithmic work
    achieved = flops / (read_kernel_ms / 1e3) / 1e12 if read_kernel_ms else None
    kernel_name = {"fp32": "reads_forward_kernel (FP32 SIMT)", "tf32x3": "reads_forward_tc_kernel<3> (tcgen05, split TF32x3)",
                   "tf32": "reads_forward_tc_kernel<1> (tcgen05, TF32)"}[args.precision]
    roofline = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved,
                "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"] if achieved else None,
                "peak_source": peaks["source"] + " bf16 dense (sustained)",
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)", "traffic_source": traffic_source,
                "algorithmic_bytes_per_launch": 12 * n_reads + (30 * 4 + 16 + 108) * args.variants,
                "kernel_ms": read_kernel_ms, "kernel_share_of_step": read_kernel_ms / step_ms if read_kernel_ms else None,
                "fp32_fma_peak_tflops_nominal": FP32_FMA_PEAK_TFLOPS,
                "frac_of_fp32_fma_peak": achieved / FP32_FMA_PEAK_TFLOPS if achieved else None,
                "flop_per_launch": flops,
                # what the tensor pipe actually executes for those algorithmic FLOP: zero-padded shapes (61->64, 30->32,
                # 60->64, 20->2x24, 10->16) and, in the split-precision mode, three TF32 MMAs per k-step
                "executed_tensor_flop_per_launch": 2 * 53248 * n_reads * (3 if args.precision == "tf32x3" else 1)
                if args.precision != "fp32" else None,
                "tf32_mma_peak_tflops_probe": 148 * 4096 * 1.965e9 / 1e12}
    result = {
        "metric": "artifact_model_inference_variants_per_sec", "value": value, "unit": "variants/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 split on tcgen05: hi/lo operands, fp32 accumulate)", "tf32": "tf32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "precision": args.precision,
                   "variants_per_gpu": args.variants, "reads_per_gpu": n_reads, "mean_reads_per_variant": n_reads / args.variants,
                   "hyperparameters": "artifact-model-v0.4.0", "timing": "inputs larger than L2 (compressed shard "
                   f"{h2d / 2**20:.0f} MiB), CUDA events, max over ranks"},
        "clocks": clocks, "gpu_launches": fwd_launches * args.steps if fwd_launches is not None else None,
        "gpu_launches_per_step": fwd_launches, "gpu_launches_how": "library kernels of one step counted from a CUPTI trace outside the timed region",
        # e2e: the dataset's arrays -> MemoryMappedBatches (batch cutting inside the timed region) -> prefetch_generator (H2D on a
        # side stream) -> compute_batch_output -> pmt_pack_posterior -> posterior records D2H into pinned host arrays.  The headline
        # is the flow with the dataset page-locked in place when the driver allows it, else the staging ring.
        "e2e": {"value": e2e_reg_value if reg_loader.registered else e2e_maps_value, "unit": "variants/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_seen[0],
                "host_cores_bound": len(numa_cores) if numa_cores else None,
                "passes_per_call": args.steps,
                "call": "one call of tools.filter_variants.generate_posterior_arrays over data.reads_dataset.MemoryMappedBatches("
                        + ("the dataset's arrays loaded into page-locked memory once, every batch three zero-copy slices"
                           if reg_loader.registered else f"staging_threads={staging_threads}: batches staged into a pinned ring by copy threads")
                        + "), reads in dataset order (gather indices written on the device); prefetch_generator H2D on a side stream, "
                          "compute_batch_output, pmt_pack_posterior, posterior records D2H into pinned host arrays",
                "from_registered_dataset": {"value": e2e_reg_value, "unit": "variants/s", "registered": bool(reg_loader.registered)},
                "from_staged_dataset": {"value": e2e_maps_value, "unit": "variants/s", "staging_threads": staging_threads,
                                        "what": "batches copied into a ring of pinned buffers by copy threads (for file-backed maps the "
                                                "driver refuses to page-lock)"},
                "from_pinned_batches": {"value": e2e_value, "unit": "variants/s",
                                        "what": "host batches built and pinned before the timed region (batch order, no gather)"}},
        "roofline": roofline,
    }
    if roofline.get("executed_tensor_flop_per_launch") and read_kernel_ms:
        roofline["executed_tensor_tflops"] = roofline["executed_tensor_flop_per_launch"] / (read_kernel_ms / 1e3) / 1e12
        roofline["frac_of_tf32_mma_peak_executed"] = roofline["executed_tensor_tflops"] / roofline["tf32_mma_peak_tflops_probe"]
    if config3 is not None:
        result["config3_single_gpu"] = config3
    if parity is not None:
        result["parity"] = parity
    if train is not None:
        result["train"] = train
        if result["gpu_launches"] is not None and train.get("gpu_launches") is not None:
            result["gpu_launches"] += train["gpu_launches"]
    if panel is not None:
        result["panel"] = panel
        result["posterior"] = run_posterior(args.variants, dev)
        result["train"]["small_batch"] = small_batch
    if not args.no_cpu_baseline:
        sample = 8192
        if reference_available():
            ref_model, batches = reference_setup(model.state_dict(), [64, sample, 2048], seed=3000)
            big = time_reference(ref_model, batches[sample], sample, False, budget_s=10.0, min_steps=2, max_steps=5)
            b64 = time_reference(ref_model, batches[64], 64, False, budget_s=3.0, max_steps=30, warm=2)
            result["cpu_baseline"] = {"value": big["variants_per_s"], "unit": "variants/s", "cores": torch.get_num_threads(), "kind": "reference",
                                      "sample": f"unmodified reference (oracle/_ref) compute_batch_output on {sample} WGS-shaped variants per "
                                                f"batch, median of {big['steps']} batches", "batch_64": b64}
            if train is not None:
                tb = time_reference(ref_model, batches[2048], 2048, True, budget_s=8.0, min_steps=2, max_steps=4)
                t64 = time_reference(ref_model, batches[64], 64, True, budget_s=3.0, max_steps=20, warm=2)
                result["cpu_baseline"]["train_value"] = tb["variants_per_s"]
                result["cpu_baseline"]["train_sample"] = ("reference compute_batch_output + compute_batch_losses + misc_utils.backpropagate on "
                                                          f"2048 WGS-shaped variants per step, median of {tb['steps']} steps")
                result["cpu_baseline"]["train_batch_64"] = t64
        else:
            v, n_it = time_oracle(model.state_dict(), sample, seed=3000, budget_s=15.0)
            result["cpu_baseline"] = {"value": v, "unit": "variants/s", "cores": torch.get_num_threads(), "kind": "port",
                                      "sample": f"oracle forward on {sample} WGS-shaped variants per batch, median of {n_it} batches"}
            if train is not None:
                vt, n_it = time_oracle(model.state_dict(), 2048, seed=3001, budget_s=10.0, train=True)
                result["cpu_baseline"]["train_value"] = vt
                result["cpu_baseline"]["train_sample"] = (f"oracle forward + losses + autograd backward on 2048 WGS-shaped variants per "
                                                          f"step, median of {n_it} steps")
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
