"""Benchmark of the GTA-attention hot path (BASELINE.json metric: GTA-attention Mtokens/s at the MSN-Hard
gta_so3 token shape; one process per GPU, batch-sharded, no collective in the forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload msn_enc|msn_dec|clevr_enc|clevr_dec|cfg1|sweep*]
    python bench.py --impl reference ...     # the reference's own CPU path (baseline/_ref, else its torch restatement)
    python bench.py --workload train_step_msn|train_step_clevr [--gpus N]    # BASELINE config 5: SRT train step, DDP

A step = one pass of the hot path over one batch of synthetic input: rep construction (once per batch, as the
reference does per encoder forward) + the library call gta_attn_fwd, which picks its pipeline from the shape (one launch
with the K/V rotation done by staging warps of the attention kernel, or staging kernel + attention kernel; `--flags 32`
/ `1024` force one / two launches, `--flags 256` / `512` select the streaming-softmax / spare-P-buffer attention kernels).
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gta_b200.synth import CFG1_A, CLEVR, MSN_SO3, GtaConfig, make_inputs  # noqa: E402

WORKLOADS = {
    # name: (base cfg, Nq, Nk, tq/view, tk/view, cross, per-GPU batch, description)
    "msn_enc": (MSN_SO3, 5, 5, 256, 256, False, 64, "runs/msn/GTA/gta_so3 encoder self-attention, 5 views x 16x16 tokens"),
    "msn_dec": (MSN_SO3, 5, 5, 512, 256, True, 64, "runs/msn/GTA/gta_so3 decoder cross-attention, Tq=2560, Tk=1280"),
    "clevr_enc": (CLEVR, 2, 2, 300, 300, False, 32, "runs/clevrtr/GTA/gta encoder self-attention, 2 views x 15x20 tokens"),
    "clevr_dec": (CLEVR, 3, 2, 853, 300, True, 32, "runs/clevrtr/GTA/gta decoder cross-attention, Tq=2559, Tk=600"),
    "cfg1": (CFG1_A, 2, 2, 1024, 1024, False, 2, "BASELINE config 1: 2 views x 32x32, d=128, 4 heads"),
    # BASELINE config 4: sequence-length sweep, B=1, d=768 (8 heads x 96), N views x 128x128 tokens (L = N*16384).
    # The reference cannot run these (its [B,H,L,L] score matrix would be >= 34 GB); inputs are generated on the device.
    "sweep2": (MSN_SO3, 2, 2, 16384, 16384, False, 1, "seq-len sweep: 2 views x 128x128 tokens, L=32768, d=768"),
    "sweep5": (MSN_SO3, 5, 5, 16384, 16384, False, 1, "seq-len sweep: 5 views x 128x128 tokens, L=81920, d=768"),
    "sweep10": (MSN_SO3, 10, 10, 16384, 16384, False, 1, "seq-len sweep: 10 views x 128x128 tokens, L=163840, d=768"),
    "sweep20": (MSN_SO3, 20, 20, 16384, 16384, False, 1, "seq-len sweep: 20 views x 128x128 tokens, L=327680, d=768"),
}
METRIC = "GTA-attention Mtokens/sec"
UNIT = "Mtokens/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def traffic_bytes(workload, batch, flags):
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed `ncu --set full` capture of this
    very command (profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum; a profiler pass cannot run inside
    the timed region).  Returns (bytes, provenance) or (None, None) when no capture exists for this workload/batch/flags."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        for t in json.load(open(p)):
            if t["workload"] == workload and t["batch_per_gpu"] == batch and t.get("flags", 0) == flags:
                return t["dram_bytes_read"] + t["dram_bytes_write"], "profiles/r02_traffic.json <- " + t["source"]
    except Exception:
        pass
    return None, None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region through NVML (in-process, ~2 ms period;
    falls back to polling nvidia-smi when pynvml is unavailable)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz = index, [], False, None
        self.t_begin, self.t_end = 0.0, float("inf")      # only samples taken inside [t_begin, t_end] are reported
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML indices follow CUDA_VISIBLE_DEVICES only through the UUID; map via torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self.handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    self.handle = h
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    self.samples.append((mhz, mask, time.perf_counter()))
                    time.sleep(0.001)
                else:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in o.strip().split(",")]
                    self.samples.append((float(f[0]), 0, time.perf_counter()))
                    self.max_mhz = float(f[1])
            except Exception:
                time.sleep(0.01)

    def begin(self):
        """The thread is started before the warm-up so that it is already polling; call this when the timed region starts."""
        self.t_begin = time.perf_counter()

    def summary(self):
        self.t_end = time.perf_counter()
        self.stop_flag = True
        self.samples = [x for x in self.samples if self.t_begin <= x[2] <= self.t_end]
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[1]
        reasons = [n for bit, n in self.REASONS.items() if mask & bit]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def cpu_impl():
    """(kind, description): "reference" when the unmodified reference is importable (baseline/_ref on the GPU box,
    /root/reference in the build container), else "port" (oracle/torch_port.py, its ATen restatement)."""
    from baseline import ref_loader
    if ref_loader.available():
        return "reference", ("the UNMODIFIED reference from %s: its own pre_compute_reps (source/encoder.py:183, "
                             "source/decoder.py:247) + multihead_geometric_transform_attention (source/utils/gta.py:92) "
                             "+ AttnFn (source/layers.py:202-211)" % ref_loader.root())
    return "port", "oracle/torch_port.py (ATen restatement of the reference path; baseline/_ref is not installed)"


def cpu_port_step(cfg, inp, tc=0.01):
    """One pass of the reference's CPU path, fp32: the reference itself when available, else its restatement."""
    if cpu_impl()[0] == "reference":
        from oracle import ref_harness as rh
        return rh.ref_gta_attention(cfg, inp, trans_coeff=tc)[0]
    from oracle import torch_port as tp
    return tp.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"], inp["coord_q"],
                            inp["coord_k"], trans_coeff=tc)


def _best_threads(cfg, inp):
    """torch's intra-op pool scales badly on the many small einsums of this path: try a few pool sizes and keep the
    fastest (the count actually used is what `cores` reports)."""
    ncpu = os.cpu_count() or 1
    best = (float("inf"), 1)
    for th in sorted({1, 4, 8, 16, 32, 64, ncpu}):
        if th > ncpu:
            continue
        torch.set_num_threads(th)
        cpu_port_step(cfg, inp)
        t0 = time.perf_counter()
        cpu_port_step(cfg, inp)
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, th)
    torch.set_num_threads(best[1])
    return best[1]


def cpu_baseline(cfg, args_w, budget_s=12.0, batch=2):
    base, nq, nk, tq, tk, cross, _, _ = args_w
    inp = make_inputs(cfg, batch, tq, tk, cross=cross, seed=123)
    inp = {k: (v.contiguous() if k in "qkv" else v) for k, v in inp.items()}
    with torch.no_grad():
        threads = _best_threads(cfg, inp)
        t0, reps = time.perf_counter(), 0
        best = float("inf")
        while reps < 3 or (time.perf_counter() - t0 < budget_s and reps < 50):
            t1 = time.perf_counter()
            cpu_port_step(cfg, inp)
            best = min(best, time.perf_counter() - t1)
            reps += 1
    tokens = batch * nq * tq
    kind, desc = cpu_impl()
    return {"value": tokens / best / 1e6, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(), "kind": kind,
            "sample": f"{desc}, fp32, batch {batch} of the same token shape, best of {reps}, {best*1e3:.1f} ms, "
                      f"thread count chosen as the fastest of a sweep"}


def run_reference(args, wl):
    base, nq, nk, tq, tk, cross, _, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    batch = 2
    inp = make_inputs(cfg, batch, tq, tk, cross=cross, seed=123)
    inp = {k: (v.contiguous() if k in "qkv" else v) for k, v in inp.items()}
    with torch.no_grad():
        threads = _best_threads(cfg, inp)
        for _ in range(args.warmup):
            cpu_port_step(cfg, inp)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_port_step(cfg, inp)
        dt = (time.perf_counter() - t0) / args.steps
    val = batch * nq * tq / dt / 1e6
    kind, desc = cpu_impl()
    sample = (f"{desc}, fp32, {threads} threads (fastest of a sweep; host has {os.cpu_count()} cpus), batch {batch} "
              f"per step")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl, batch),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def informational_legs(cfg, wl, B, dev_bufs, dev_small, views, cross, timed, steps):
    """SURVEY 8(d) informational comparisons on the SAME device-resident inputs, outside `value`: the reference's own eager
    function on the B200 (needs baseline/_ref) and the practical library bar — the reference's PyTorch rep rotation around
    torch.nn.functional.scaled_dot_product_attention.  bf16 q/k/v, reps built by the reference's own pre_compute_reps
    (fp32) once, outside the timed region (as the reference does per forward, not per layer)."""
    from baseline import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref is not installed (python baseline/install_ref.py)"}
    base, nq, nk, tq, tk, _, _, _ = wl
    try:
        ref = ref_loader.load()
        import torch.nn.functional as F
        ek, ck = dev_small["extr_k"], dev_small["coord_k"]
        extras = {"input_transforms": ek, "input_coord": ck.reshape(B, nk, -1, 2)}
        akw = dict(f_dims=dict(cfg.f_dims), so2=cfg.so2, so3=cfg.so3, max_freq_h=cfg.max_freq_h, max_freq_w=cfg.max_freq_w,
                   shared_freqs=cfg.shared_freqs)
        enc = ref.encoder.ImprovedSRTEncoder.__new__(ref.encoder.ImprovedSRTEncoder)
        orig_enc = gta_orig("enc") or ref.encoder.ImprovedSRTEncoder.pre_compute_reps
        orig_enc(enc, akw, extras)
        if cross:
            extras["target_transforms"] = dev_small["extr_q"]
            extras["target_coord"] = dev_small["coord_q"].reshape(B, nq, -1, 2)
            dec = ref.decoder.ImprovedSRTDecoder.__new__(ref.decoder.ImprovedSRTDecoder)
            (gta_orig("dec") or ref.decoder.ImprovedSRTDecoder.pre_compute_reps)(dec, akw, extras)
        scale = cfg.head_dim ** -0.5

        class Eager:            # AttnFn.forward, source/layers.py:207-211
            def __call__(self, q, k, v):
                attn = torch.softmax((q @ k.transpose(-1, -2)) * scale, -1)
                return attn @ v, attn

        class Sdpa:
            def __call__(self, q, k, v):
                return F.scaled_dot_product_attention(q, k, v, scale=scale), None
        q, k, v = views(dev_bufs)
        tc = torch.tensor([0.01], device=q.device)
        fn = gta_orig("attn") or ref.gta.multihead_geometric_transform_attention
        out = {}
        for name, attn_fn in (("reference_eager_gpu", Eager()), ("reference_rotation_plus_sdpa_gpu", Sdpa())):
            def run():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    fn(q, k, v, attn_fn=attn_fn, f_dims=dict(cfg.f_dims), reps=extras, trans_coeff=tc, v_transform=True,
                       euclid=False)
            try:
                for _ in range(2):
                    run()
                ms = timed(run, max(3, steps // 5))
                out[name] = {"ms": ms, "Mtokens_per_s": B * nq * tq / (ms * 1e-3) / 1e6}
            except torch.cuda.OutOfMemoryError:
                out[name] = {"unavailable": "out of memory (the eager path materialises [B,H,Tq,Tk])"}
                torch.cuda.empty_cache()
        out["note"] = ("same device-resident bf16 q/k/v and batch as `value`; reference from %s; reps prebuilt (not timed); "
                       "CUDA events; informational only" % ref_loader.root())
        return out
    except Exception as e:      # never let an informational leg break the bench line
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def gta_orig(which):
    """The reference's original callables when gta_b200.gta.install() has rebound them (it has not in this process unless
    a train_step workload ran)."""
    from gta_b200 import gta as fast
    if which == "attn":
        return fast._original
    return fast._original_reps.get(which)


def run_train_step(args):
    """BASELINE config 5: one training step of the reference's TransformingSRT (built from its own YAML:
    runs/msn/GTA/gta_so3 or runs/clevrtr/GTA/gta) on synthetic batches, global batch 32, AdamW, the config's precision
    (MSN: bf16 autocast), encoder and decoder wrapped in DistributedDataParallel separately (train.py:182-188; the
    collective is DDP's bucketed NCCL gradient all-reduce).  Timed twice on the same weights and batches: with the
    drop-in installed (gta_b200.gta.install()) and with the reference's eager path.  value = samples/s of the installed
    path over all ranks."""
    from baseline import ref_loader, srt_synth
    run = "msn/GTA/gta_so3" if args.workload == "train_step_msn" else "clevrtr/GTA/gta"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not ref_loader.available():
        if rank == 0:
            print(json.dumps({"metric": "SRT train step", "unavailable": "baseline/_ref is not installed"}), flush=True)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    ref = ref_loader.load()
    from gta_b200 import gta as fast
    torch.manual_seed(0)
    model, cfg = srt_synth.build_model(ref, run, dev)          # dropout as configured (0.01)
    mixed = bool(cfg["training"].get("mixed_prec", False))
    gb = 32
    B = args.train_batch or max(1, gb // world)
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model.encoder = DDP(model.encoder, device_ids=[local], output_device=local)
        model.decoder = DDP(model.decoder, device_ids=[local], output_device=local)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    batches = [srt_synth.make_batch(run, B, dev, seed=100 * rank + i) for i in range(2)]
    nparams = sum(p.numel() for p in model.parameters())

    def train_step(i):
        model.train()
        opt.zero_grad(set_to_none=True)
        loss, _ = srt_synth.loss_fn(model, batches[i % len(batches)], mixed)
        loss.backward()
        opt.step()
        return loss

    def timed_steps(n):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            loss = train_step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(loss.detach())

    res = {}
    steps = max(3, min(args.steps, 20))
    for name in ("eager", "installed"):
        if name == "installed":
            fast.install()
        else:
            fast.uninstall()
        for i in range(max(2, min(args.warmup, 3))):
            train_step(i)
        sampler = ClockSampler(local)
        sampler.start()
        sampler.begin()
        res[name] = timed_steps(steps)
        res[name + "_clocks"] = sampler.summary()
    fast.uninstall()
    if rank == 0:
        ms_i, ms_e = res["installed"][0], res["eager"][0]
        line = {"metric": "SRT encoder+decoder train step (samples/sec)", "value": world * B / (ms_i * 1e-3), "unit": "samples/s",
                "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": ms_i, "higher_is_better": True,
                "scaling": "strong" if not args.train_batch else "weak", "vs_baseline": None,
                "dtype": "bf16 autocast" if mixed else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "run": "runs/%s/config.yaml" % run, "global_batch": world * B,
                           "batch_per_gpu": B, "params": nparams, "optimizer": "AdamW",
                           "parallelism": "DDP on encoder and decoder separately (train.py:182-188); NCCL bucketed gradient "
                                          "all-reduce of %.0f MB fp32 per step" % (nparams * 4 / 1e6),
                           "gta_layers": "5 encoder self-attention + 2 decoder cross-attention, forward and backward through "
                                         "gta_attn_fwd / gta_attn_bwd, reps by gta_build_reps"},
                "reference_eager": {"ms_per_step": ms_e, "samples_per_s": world * B / (ms_e * 1e-3), "clocks": res["eager_clocks"]},
                "speedup_vs_reference_eager": ms_e / ms_i, "loss": {"installed": res["installed"][1], "eager": res["eager"][1]},
                "clocks": res["installed_clocks"]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def workload_config(name, wl, batch):
    base, nq, nk, tq, tk, cross, _, desc = wl
    return {"workload": name, "description": desc, "batch_per_gpu": batch, "heads": base["heads"],
            "head_dim": base["head_dim"], "Tq": nq * tq, "Tk": nk * tk, "q_views": nq, "k_views": nk,
            "f_dims": base["f_dims"], "attention": "cross" if cross else "self", "trans_coeff": 0.01,
            "parallelism": "batch-shard, one process per GPU, no collective in the forward"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msn_enc", choices=sorted(WORKLOADS) + ["train_step_msn", "train_step_clevr"])
    ap.add_argument("--train-batch", type=int, default=0, help="train_step workloads: per-GPU batch (default: 32 / world, BASELINE config 5)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--flags", type=int, default=0, help="GTA_FLAG_* bits for gta_attn_fwd (1 = P in TMEM)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="batch chunks of the host-buffer pipeline (e2e leg)")
    ap.add_argument("--no-backward", action="store_true", help="skip the fused-backward leg (`backward` object)")
    ap.add_argument("--bwd-flags", type=int, default=0, help="GTA_FLAG_* bits for the backward leg (4096 = the dK/dV + dQ kernel pair)")
    ap.add_argument("--backward", action="store_true", help="(kept for compatibility; the backward leg is on by default)")
    ap.add_argument("--no-info", action="store_true", help="skip the informational GPU legs (reference eager / SDPA on the B200)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload.startswith("train_step"):
        return run_train_step(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    base, nq, nk, tq, tk, cross, B, desc = wl
    if args.batch:
        B = args.batch
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    from gta_b200 import _lib, ops, shard
    _lib.lib()

    # ---- synthetic inputs, pinned on the host (e2e leg) and resident on the device (value leg)
    host = make_inputs(cfg, B, tq, tk, cross=cross, seed=1000 + rank, dtype=torch.bfloat16)
    self_attn = not cross
    pin = {}
    for k_, v_ in host.items():
        if self_attn and k_ in ("extr_q", "coord_q"):
            continue
        pin[k_] = v_
    # q/k/v are views of shared projection buffers: keep the buffers pinned and re-derive the views
    if self_attn:
        buf = host["q"]._base if host["q"]._base is not None else host["q"]
        while buf._base is not None:
            buf = buf._base
        bufs_host = {"qkv": buf.pin_memory()}
    else:
        bq, bkv = host["q"], host["k"]
        while bq._base is not None:
            bq = bq._base
        while bkv._base is not None:
            bkv = bkv._base
        bufs_host = {"q": bq.pin_memory(), "kv": bkv.pin_memory()}
    small_host = {k_: host[k_].contiguous().pin_memory() for k_ in ("extr_k", "coord_k")}
    if cross:
        small_host.update({k_: host[k_].contiguous().pin_memory() for k_ in ("extr_q", "coord_q")})
    H, D = cfg.heads, cfg.head_dim

    def views(bufs):
        hv = lambda x: x.view(x.shape[0], x.shape[1], -1, D).permute(0, 2, 1, 3)
        if self_attn:
            q, k, v = (hv(t) for t in bufs["qkv"].chunk(3, dim=-1))
        else:
            q = hv(bufs["q"])
            k, v = (hv(t) for t in bufs["kv"].chunk(2, dim=-1))
        return q, k, v

    dev_bufs = {k_: torch.empty_like(v_, device=dev) for k_, v_ in bufs_host.items()}
    dev_small = {k_: torch.empty_like(v_, device=dev) for k_, v_ in small_host.items()}
    out_host = torch.empty(B, nq * tq, H, D, dtype=torch.bfloat16).pin_memory()
    tc = torch.tensor([0.01], device=dev)

    def h2d():
        for k_ in bufs_host:
            dev_bufs[k_].copy_(bufs_host[k_], non_blocking=True)
        for k_ in small_host:
            dev_small[k_].copy_(small_host[k_], non_blocking=True)

    def step(flags=0):
        ek, ck = dev_small["extr_k"], dev_small["coord_k"]
        eq = dev_small.get("extr_q", ek) if cross else ek
        cq = dev_small.get("coord_q", ck) if cross else ck
        reps = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3)
        q, k, v = views(dev_bufs)
        return ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, flags=args.flags | flags), reps

    h2d()
    torch.cuda.synchronize()
    in_bytes = sum(t.numel() * t.element_size() for t in dev_bufs.values())
    _q, _k, _v = views(dev_bufs)
    pipeline = ops.pipeline_of(_q, _k, _v, ops.PackedReps(n_q_views=nq, n_k_views=nk), cfg.f_dims, flags=args.flags)
    two_launch = pipeline != "single launch"
    # kernels per step: build_reps_kernel (1: view tables + both SO(2) tables) + [staging (1)] + attention (1)
    launches_per_step = 1 + (2 if two_launch else 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = shard.max_over_ranks(e0.elapsed_time(e1), device=dev)     # whole-job time = the slowest rank
        barrier()
        return ms / n

    sampler = ClockSampler(local)
    sampler.start()                       # polling starts before the warm-up; only samples of the timed region are kept
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler.begin()
    ms_step = timed(step, args.steps)
    clocks = sampler.summary()

    # ---- the library call alone (reps already built): the dominant launch, roofline numerator
    _, reps = step()
    q, k, v = views(dev_bufs)
    op_only = lambda: ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, flags=args.flags)
    attn_only = lambda: ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc,
                                              flags=(args.flags & ~_lib.GTA_FLAG_SINGLE_LAUNCH) | _lib.GTA_FLAG_TWO_LAUNCH | _lib.GTA_FLAG_SKIP_STAGE)
    stage_only = lambda: ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc,
                                               flags=(args.flags & ~_lib.GTA_FLAG_SINGLE_LAUNCH) | _lib.GTA_FLAG_STAGE_ONLY)
    for _ in range(3):
        op_only()
    ms_op = timed(op_only, args.steps)
    stage_only()                                  # (re)stage the workspace for the attention-only measurement
    for _ in range(3):
        attn_only()
    ms_attn = timed(attn_only, args.steps)
    ms_stage = timed(stage_only, args.steps)

    # ---- end to end through the public API with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        # the repo's host-buffer entry point: batch chunks pipelined over three streams (H2D | compute | D2H)
        from gta_b200.host import HostStagedAttention
        pipe = HostStagedAttention(cfg, bufs_host, small_host, out_host, dev, chunks=args.e2e_chunks)

        def e2e_step():
            pipe.run(trans_coeff=tc, flags=args.flags)
        for _ in range(2):
            e2e_step()
        n_e2e = max(3, args.steps // 2)
        ms_e2e = timed(e2e_step, n_e2e)
        h2d_bytes = in_bytes + sum(t.numel() * t.element_size() for t in dev_small.values())
        e2e = {"value": world * B * nq * tq / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": out_host.numel() * 2, "ms_per_step": ms_e2e, "steps": n_e2e,
               "api": "gta_b200.host.HostStagedAttention.run: %d batch chunks, H2D / compute / D2H on three streams, input copies of consecutive calls back to back" % len(pipe.bounds)}

    bwd = None
    if not args.no_backward:
        out_f, lse_f = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
        dout = torch.randn(out_f.shape, device=dev).to(out_f.dtype)
        bwd_fn = lambda: ops.gta_attention_bwd(dout, q, k, v, out_f, lse_f, reps, cfg.f_dims, trans_coeff=tc, flags=args.bwd_flags)
        for _ in range(3):
            bwd_fn()
        ms_bwd = timed(bwd_fn, max(3, args.steps // 2))
        bwd_flops = 2.5 * 4.0 * B * H * nq * tq * nk * tk * D
        pk_b, _ = peaks()
        fused_bwd = D <= 96 and not (args.bwd_flags & _lib.GTA_FLAG_BWD_SPLIT)
        bwd = {"ms": ms_bwd,
               "kernels": ("K'/V'/Q'/dO' staging + delta + attn_bwd_fused_kernel (dK, dV, bulk-reduced dQ partial sums) + bwd_dq_finish_kernel"
                           if fused_bwd else "K'/V'/Q'/dO' staging + delta + attn_bwd_kernel<dKV> + attn_bwd_kernel<dQ>"),
               "tflops_algorithmic": bwd_flops / (ms_bwd * 1e-3) / 1e12,
               "frac_of_peak": bwd_flops / (ms_bwd * 1e-3) / 1e12 / pk_b["bf16_tflops"],
               "fwd_bwd_Mtokens_per_s": world * B * nq * tq / ((ms_step + ms_bwd) * 1e-3) / 1e6}
        del out_f, lse_f, dout

    info = None
    if not args.no_info and rank == 0:
        info = informational_legs(cfg, wl, B, dev_bufs, dev_small, views, cross, timed, args.steps)

    if rank == 0:
        Tq, Tk = nq * tq, nk * tk
        flops = 4.0 * B * H * Tq * Tk * D
        pk, pk_src = peaks()
        total_s = ms_step * 1e-3 * args.steps
        burst = total_s < 2.0
        peak = pk["bf16_tflops"] if burst else pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        ms_dom = ms_attn if two_launch else ms_op
        achieved = flops / (ms_dom * 1e-3) / 1e12
        if args.flags & _lib.GTA_FLAG_V5_PIPELINE:
            kname = "attn_fwd6_kernel (persistent two-tile pipeline with the spare P buffer; K'/V' staged by rotate_kv_kernel)"
        elif args.flags & _lib.GTA_FLAG_V4_PIPELINE:
            kname = "attn_fwd5_kernel (streaming softmax + epilogue warpgroup; K'/V' staged by rotate_kv_kernel)"
        elif two_launch:
            kname = "attn_fwd3_kernel (persistent two-tile pipeline; K'/V' staged by rotate_kv_kernel)"
        else:
            kname = "attn_fwd4_kernel (ONE launch: K/V rotation by staging warps + persistent two-tile attention pipeline)"
        traffic, traffic_src = traffic_bytes(args.workload, B, args.flags)
        line = {
            "metric": METRIC, "value": shard.job_throughput(B * Tq, world, ms_step) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(args.workload, wl, B),
                           l2="q/k/v inputs %.0f MB per step > 126 MB L2 (no explicit flush needed)" % (in_bytes / 1e6)
                           if in_bytes > 126e6 else "inputs %.0f MB fit L2; not flushed" % (in_bytes / 1e6),
                           p_operand="tmem (TS form of tcgen05.mma: P read from tensor memory)",
                           pipeline=pipeline + (" (chosen by the library from the shape)" if not args.flags & (_lib.GTA_FLAG_SINGLE_LAUNCH | _lib.GTA_FLAG_TWO_LAUNCH) else " (forced by --flags)"), flags=args.flags),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kname, "kernel_ms": ms_dom,
                         "step_frac": flops / (ms_step * 1e-3) / 1e12 / peak,
                         "step_note": "step_frac = the same FLOPs over the whole step (rep construction + every launch)",
                         "two_launch_attention_kernel_ms": ms_attn, "staging_kernel_ms": ms_stage, "library_call_ms": ms_op,
                         "flops_per_launch": flops, "peak_source": pk_src + (" burst" if burst else " sustained")},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if bwd:
            line["backward"] = bwd
        if info:
            line["informational"] = info
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, wl)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
