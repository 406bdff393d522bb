#!/usr/bin/env python
"""Benchmark of the Latent2im hot path: StyleGAN2 latent-walk edited images per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic latents:
z -> mapping -> linear W+ walk -> StyleGAN2-1024 synthesis (fresh per-layer noise) -> images.
N > 1 is launched by torchrun (one rank per GPU); the batch of latents is sharded across ranks,
there is no data-path collective (weak scaling), the timed region is bracketed by a barrier +
torch.cuda.synchronize() and the step time is the MAX over ranks.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` is the same metric
through the public host-buffer API (pinned-host z/alpha -> device, uint8 images -> pinned host
inside the timed region).  `roofline` describes the dominant kernel (the tcgen05 implicit-GEMM
modulated conv) with per-launch durations measured live by CUDA events on the launching stream;
`cpu_baseline` is the CPU oracle port timed on this box's host cores on a bounded sample.
`--impl reference` times the reference's CPU path (the oracle port: the reference itself is
CUDA-only and has no CPU path) for the same metric and config.
"""
import argparse
import ctypes as C
import json
import os
os.environ.setdefault("L2I_ALLOW_RANDOM_INIT", "1")   # synthetic-weight benchmark
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stylegan2_1024_latent_walk_edited_images_per_sec"
UNIT = "images/s"


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32, help="latent samples per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="images in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-layer device timing table here")
    return ap.parse_args()


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9 and parts[0] == str(self.gpu_index):
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _cpu_oracle_images_per_sec(size, n_images, seed=0):
    """The CPU restatement of the reference path (oracle/), fp32, all host threads."""
    import torch
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import synthetic_noise, synthetic_state_dict, synthetic_walk_w, synthetic_z
    from oracle import GeneratorSpec, generator_forward_ref, mapping_ref
    from oracle.walks import walk_linear_ref
    cores = _host_cores()
    torch.set_num_threads(cores)
    spec = GeneratorSpec(size=size)
    shapes = {k: v.shape for k, v in Generator(size, 512, 8).state_dict().items()}
    sd = synthetic_state_dict(shapes, 0)
    walk_w = synthetic_walk_w(1, spec.n_latent, 512, seed=0)
    per_call = min(n_images, 2)

    def one_pass(b, s):
        z = torch.tensor(synthetic_z(b, seed=s), dtype=torch.float32)
        alpha = torch.full((b, 1), 0.5)
        with torch.no_grad():
            w = mapping_ref(sd, z, spec)
            lat = torch.stack(walk_linear_ref([w] * spec.n_latent, alpha, walk_w), 1)
            noise = [torch.randn(b, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(spec.num_layers)]
            return generator_forward_ref(sd, lat, noise, spec)

    one_pass(1, 99)  # warm-up (thread pool, allocator)
    done, t0 = 0, time.perf_counter()
    while done < n_images:
        b = min(per_call, n_images - done)
        one_pass(b, done)
        done += b
    dt = time.perf_counter() - t0
    return done / dt, cores, dt


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_per_step = max(1, args.cpu_sample // 2)
    for _ in range(min(args.warmup, 1)):
        _cpu_oracle_images_per_sec(args.size, 1)
    total, t_total, cores = 0, 0.0, _host_cores()
    for _ in range(args.steps):
        ips, cores, dt = _cpu_oracle_images_per_sec(args.size, n_per_step)
        total += n_per_step
        t_total += dt
    value = total / t_total
    sample = f"{n_per_step} image(s) of the {args.size}px batch-{args.batch} workload per step, fp32, torch CPU threads={cores}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"StyleGAN2 config-f {args.size}px random-init, linear w-walk edit forward "
                               f"(mapping + walk + synthesis), batch {args.batch}/GPU", "size": args.size,
                   "batch_per_gpu": args.batch, "note": "the reference has no CPU path (its ops are CUDA-only); this arm "
                   "times the CPU oracle port of the same algorithm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from latent2im_b200 import _native as nt
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.pipeline import EditPipeline
    from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w, synthetic_z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nt.load()

    size, b = args.size, args.batch
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    gen = load_synthetic(Generator(size, 512, 8), seed=0).to(dev).eval()
    gen.set_native(dtype=dtype, max_batch=b)
    import numpy as np
    np.random.seed(0)
    walk = WalkLinearMultiW(512, gen.log_size - 2, 1, ["Smiling"]).to(dev)
    with torch.no_grad():
        walk.w.copy_(synthetic_walk_w(1, gen.n_latent, 512, seed=0).to(dev))
    pipe = EditPipeline(gen, walk, b, n_attr=1, device=dev)

    def z_shard(step):
        zg = synthetic_z(world * b, seed=step)          # global batch; this rank takes its rows
        return zg[rank * b:(rank + 1) * b]

    alpha_host = torch.linspace(0, 1, b).reshape(b, 1)
    z_dev = torch.tensor(z_shard(0), dtype=torch.float32, device=dev)
    alpha_dev = alpha_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        l0 = nt.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        if after is not None:
            after()          # e.g. make the timing stream wait for the last asynchronous device->host copy
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = nt.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            lt = torch.tensor([launches], device=dev, dtype=torch.int64)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        barrier()
        return ms, launches

    # ---- device-resident arm ------------------------------------------------------------------
    def step_dev(_s):
        pipe.edit_device(z_dev, alpha_dev)

    for s in range(args.warmup):
        step_dev(s)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    ms_dev, launches = timed(step_dev, args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end arm (host buffers through the public API) ---------------------------------
    zs = [z_shard(s) for s in range(max(args.steps, 1))]

    def step_e2e(s):
        pipe.edit(zs[s % len(zs)], alpha_host, sync=False)

    for s in range(min(args.warmup, 3)):
        step_e2e(s)
    ms_e2e, _ = timed(step_e2e, args.steps, after=pipe.join)

    # ---- per-segment device timing of the same step (roofline of the dominant kernel) ---------
    h = gen._handle(dev, b)
    nt.check(h.lib.l2i_generator_set_profiling(h.handle, 1), "set_profiling")
    table = {}
    nprof = 3
    for _ in range(nprof):
        step_dev(0)
        torch.cuda.synchronize()
        n = h.lib.l2i_generator_profile_count(h.handle)
        for i in range(n):
            name = C.create_string_buffer(64)
            kind, ms, fl, by = C.c_int(), C.c_float(), C.c_double(), C.c_double()
            nt.check(h.lib.l2i_generator_profile_entry(h.handle, i, name, 64, C.byref(kind), C.byref(ms), C.byref(fl),
                                                       C.byref(by)), "profile_entry")
            ent = table.setdefault(name.value.decode(), {"kind": kind.value, "ms": 0.0, "flops": fl.value, "bytes": by.value})
            ent["ms"] += ms.value / nprof
    nt.check(h.lib.l2i_generator_set_profiling(h.handle, 0), "set_profiling")
    peaks = _peaks()
    conv = [v for v in table.values() if v["kind"] == 0]
    blur = [v for v in table.values() if v["kind"] == 1]
    conv_ms, conv_fl = sum(v["ms"] for v in conv), sum(v["flops"] for v in conv)
    blur_ms, blur_by = sum(v["ms"] for v in blur), sum(v["bytes"] for v in blur)
    seg_ms = sum(v["ms"] for v in table.values())
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM traffic of the same launches from the committed ncu --set full capture (profiles/), per step like `achieved`
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and args.dtype == "bf16" and size == 1024:
        tj = json.load(open(tpath))
        traffic = tj.get("conv_dram_bytes_per_image", 0.0) * b
        traffic_src = tj.get("source")
    # per-layer roofline: each conv segment against max(tensor time, HBM time) of its ALGORITHMIC flops / bytes
    layers, ideal_ms = [], 0.0
    for name, v in table.items():
        if v["kind"] not in (0, 1) or v["ms"] <= 0:
            continue
        t_tensor = v["flops"] / (peaks["tflops"] * 1e12) * 1e3
        t_hbm = v["bytes"] / (peaks["hbm_gbs"] * 1e9) * 1e3
        ideal = max(t_tensor, t_hbm)
        ideal_ms += ideal
        layers.append({"layer": name, "ms": round(v["ms"], 4), "tflops": round(v["flops"] / v["ms"] / 1e9, 1),
                       "gbs": round(v["bytes"] / v["ms"] / 1e6, 1), "bound": "tensor" if t_tensor >= t_hbm else "hbm",
                       "frac": round(ideal / v["ms"], 3)})
    roofline = {"kernel": "tcgen05 implicit-GEMM modulated convs (conv_tc / conv_tc_halo / conv_tc_quad, all %d launches of a step)" % len(conv)
                if args.dtype == "bf16" else "conv_simt_kernel", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peaks["source"], "launch_ms_total": conv_ms,
                "share_of_step": conv_ms / seg_ms if seg_ms else None, "algorithmic_gflop_per_step": conv_fl / 1e9,
                "speed_of_light_frac": (ideal_ms / (conv_ms + blur_ms)) if (conv_ms + blur_ms) > 0 else None}
    blur_gbs = blur_by / (blur_ms * 1e-3) / 1e9 if blur_ms > 0 else 0.0
    roofline_hbm = {"kernel": "blur_act_kernel (FIR blur + noise + bias + lrelu + next-style scale; layers below 256 px)", "bound": "hbm",
                    "achieved": blur_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": blur_gbs / peaks["hbm_gbs"],
                    "traffic": None, "launch_ms_total": blur_ms, "share_of_step": blur_ms / seg_ms if seg_ms else None}
    if args.profile_json and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
        with open(args.profile_json, "w") as f:
            json.dump({"batch": b, "size": size, "dtype": args.dtype, "segments": table, "peaks": peaks}, f, indent=1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample -------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, dt = _cpu_oracle_images_per_sec(size, args.cpu_sample)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_sample} images of the same {size}px workload, fp32 oracle, {dt:.1f}s"}

    images = world * b * args.steps
    line = {
        "metric": METRIC, "value": images / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"StyleGAN2 config-f {size}px random-init (FFHQ shape), linear w-walk edit forward "
                               f"(mapping + walk + synthesis with fresh noise), batch {b}/GPU", "size": size,
                   "batch_per_gpu": b, "global_batch": world * b, "parallelism": f"dp{world} (latents sharded, no collective)",
                   "l2": "working set (>= 2 GB of activations per step) is larger than the 126 MB L2; no explicit flush"},
        "e2e": {"value": images / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
        "roofline_layers": layers,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = _args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
