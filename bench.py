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
    ap.add_argument("--workload", default="edit", choices=["edit", "train", "panels", "panels_mlp"],
                    help="edit = BASELINE cfg2 (default, the headline); train = cfg3 (walk-training step, all-reduce inside the "
                         "timed region); panels = cfg1 (256 px, batch 4, 10-panel vis_w render); panels_mlp = cfg4 (MLP walk, 5 attributes)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's own GPU path (baseline/_ref) beside ours")
    ap.add_argument("--reg-fp32", action="store_true", help="train / panels: run the stock ResNet-50 regressor in fp32 NCHW (reference arithmetic)")
    a = ap.parse_args()
    if a.workload == "train" and a.batch == 32:
        a.batch = 16                      # BASELINE cfg3: batch 16 / GPU
    if a.workload == "panels":
        a.size, a.batch = (256, 4) if (a.size, a.batch) == (1024, 32) else (a.size, a.batch)
    return a


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


def _ref_gpu_baseline(size, batch, timeout_s=300):
    """The reference's OWN GPU path (its JIT-built upfirdn2d / fused ops + cuDNN grouped convs, fp32) on this GPU, same
    weights / latents: SURVEY 8d's "meaningful beat-it number".  Runs tools/bench_reference_gpu.py --ref-only in a
    subprocess (the staged reference copy baseline/_ref travels with the repo; absent => None)."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "graphs")):
        return {"unavailable": "baseline/_ref not staged"}
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_reference_gpu.py"), "--ref-only", "--size", str(size),
                              "--batch", str(batch), "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=timeout_s)
        rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    except Exception as e:
        return {"unavailable": f"reference GPU run failed: {type(e).__name__}: {str(e)[:120]}"}
    on, off = rec.get("fp32_tf32_on", {}), rec.get("fp32_tf32_off", {})
    return {"value": on.get("images_per_s"), "unit": UNIT, "value_tf32_off": off.get("images_per_s"), "dtype": "fp32 (TF32 convs on = PyTorch default)",
            "what": f"unmodified reference Generator.forward (JIT ops + cuDNN), {size}px batch {batch}, device-resident, no walk / mapping",
            "ops_build_s": rec.get("ops_build_s")}


def _fixture_parity(gen, dev):
    """bf16 path vs the image the UNMODIFIED reference produced on a B200 (tests/golden/ref_gpu_fullsize.npz, unscaled SURVEY 8d
    recipe, 1024 px, batch 1): both PSNR normalisations (peak-to-peak 2 = the north star's [-1, 1] convention; own range =
    amplitude independent).  Restores the benchmark weights afterwards."""
    import numpy as np
    import torch
    path = os.path.join(ROOT, "tests", "golden", "ref_gpu_fullsize.npz")
    if not os.path.exists(path) or gen.size != 1024:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_ref_gpu_fullsize import CASES, CROP, CROPS_1024, fullsize_inputs
    from latent2im_b200.synthetic import load_synthetic, synthetic_z
    z = np.load(path)
    (tag, size, batch, seed), = [c for c in CASES if c[1] == 1024]
    load_synthetic(gen, seed=seed)
    with torch.no_grad():
        w = gen.style(torch.tensor(synthetic_z(batch, 10 + seed), dtype=torch.float32, device=dev))
        lat, noise, _ = fullsize_inputs(size, batch, seed, gen.n_latent, gen.num_layers, w)
        img, _ = gen(lat, input_is_latent=True, noise=noise)
    crops = torch.from_numpy(z[f"{tag}_crops"]).double()
    d = torch.stack([img[:, :, y:y + CROP, x:x + CROP].cpu().double() for (y, x) in CROPS_1024]) - crops
    mse = (d ** 2).mean().item()
    mom = z[f"{tag}_moments"]
    span = float(mom[3].max() - mom[2].min())
    load_synthetic(gen, seed=0)
    return {"psnr_p2p2_db": round(10 * np.log10(4.0 / mse), 2), "psnr_own_range_db": round(10 * np.log10(span ** 2 / mse), 2),
            "image_span": round(span, 2), "recipe": "SURVEY 8d unscaled (unit-variance ToRGB weights), seed 2, 1024 px, batch 1",
            "against": "tests/golden/ref_gpu_fullsize.npz (unmodified reference on B200, fp32, TF32 off), six 128x128 crops"}


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
        # two back-to-back steps, the segment events are those of the second: the host is a whole step ahead of the GPU, as in the
        # timed loop (after a synchronize the first small kernels of a forward wait for their launches and inflate their segments)
        step_dev(0)
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
    traffic, traffic_src, traffic_fir = None, None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and args.dtype == "bf16" and size == 1024:
        tj = json.load(open(tpath))
        traffic = tj.get("conv_dram_bytes_per_image", 0.0) * b
        traffic_fir = tj.get("fir_dram_bytes_per_image", 0.0) * b
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
    roofline = {"kernel": "tcgen05 implicit-GEMM modulated convs (conv_tc / _ares_pair / _vpair / _quad / _uprow, all %d launches of a step)" % len(conv)
                if args.dtype == "bf16" else "conv_simt_kernel", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peaks["source"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                "launch_ms_total": conv_ms,
                "share_of_step": conv_ms / seg_ms if seg_ms else None, "algorithmic_gflop_per_step": conv_fl / 1e9,
                "speed_of_light_frac": (ideal_ms / (conv_ms + blur_ms)) if (conv_ms + blur_ms) > 0 else None}
    blur_gbs = blur_by / (blur_ms * 1e-3) / 1e9 if blur_ms > 0 else 0.0
    roofline_hbm = {"kernel": "fir_tma_kernel (blur_act: FIR blur + noise + bias + lrelu + next-style scale; layers below 256 px)", "bound": "hbm",
                    "achieved": blur_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": blur_gbs / peaks["hbm_gbs"],
                    "traffic": traffic_fir, "traffic_source": traffic_src, "launch_ms_total": blur_ms,
                    "share_of_step": blur_ms / seg_ms if seg_ms else None}
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
        "parity": _fixture_parity(gen, dev) if args.dtype == "bf16" else None,
        "ref_gpu_baseline": None if (args.no_ref_gpu or world > 1) else _ref_gpu_baseline(size, b),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _timed(fn, steps, world, dev, after=None):
    """barrier + synchronize on both sides, CUDA events on the current stream, MAX over ranks; returns total ms."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(s)
    if after is not None:
        after()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    barrier()
    return ms


def _resnet50_regressor(dev, amp):
    """Stock torchvision ResNet-50 with a 40-attribute head (transform_base.py:396-403; random init)."""
    import torch
    import torchvision
    torch.manual_seed(1)
    reg = torchvision.models.resnet50(weights=None)
    reg.fc = torch.nn.Linear(2048, 40)
    reg = torch.nn.Sequential(reg, torch.nn.Sigmoid()).to(dev).eval()
    from latent2im_b200.regressor import fold_batchnorm
    reg = fold_batchnorm(reg, inplace=True)       # frozen eval-mode BN folded into the convs (as TransformGraph.get_reg_module does)
    if amp:
        reg = reg.to(memory_format=torch.channels_last)
    for p_ in reg.parameters():
        p_.requires_grad_(False)
    return reg


def run_train(args):
    """BASELINE cfg3: the train.py walk-training step (G fwd no-grad + R + walk + G fwd/bwd + R fwd/bwd + BCE + all-reduce of
    the walk gradient + Adam), StyleGAN2-1024, batch 16 / GPU, data parallel over latents.  The all-reduce (NCCL) is inside
    the timed region; `e2e` additionally stages z from pinned host memory and reads the loss back every step."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from latent2im_b200 import _native as nt
    from latent2im_b200 import parallel
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.graphs.stylegan_v2_real.transform_base import WalkLinearMultiW
    from latent2im_b200.synthetic import load_synthetic, synthetic_walk_w, synthetic_z
    from latent2im_b200.train_step import WalkTrainer

    world, rank, local_rank, dev = _dist_setup()
    nt.load()
    size, b = args.size, args.batch
    amp = not args.reg_fp32
    gen = load_synthetic(Generator(size, 512, 8), seed=0).to(dev).eval()
    gen.set_native(dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32, max_batch=b)
    reg = _resnet50_regressor(dev, amp)
    np.random.seed(0)
    walk = WalkLinearMultiW(512, gen.log_size - 2, 1, ["Smiling"]).to(dev)
    with torch.no_grad():
        walk.w.copy_(synthetic_walk_w(1, gen.n_latent, 512, seed=0).to(dev))
    trainer = WalkTrainer(gen, walk, reg, [31], lr=1e-4)
    rs = np.random.RandomState(0)            # every rank draws the SAME per-step target (SURVEY 8e equivalence caveat)

    def z_shard(step):
        return synthetic_z(world * b, seed=step)[parallel.shard_rows(world * b, rank, world)]

    z_dev = torch.tensor(z_shard(0), dtype=torch.float32, device=dev)
    target = torch.full((b, 1), float(rs.uniform(0, 1)), device=dev)
    ctx = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp)

    def step_dev(_s):
        with ctx():
            trainer.step(z_dev, target)

    for s_ in range(args.warmup):
        step_dev(s_)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = nt.launch_count()
    ms_dev = _timed(step_dev, args.steps, world, dev)
    launches = nt.launch_count() - l0
    clocks = sampler.stop() if sampler else None

    zs = [z_shard(s_) for s_ in range(args.steps)]
    z_pin = torch.empty(b, 512, dtype=torch.float32).pin_memory()
    t_pin = torch.empty(b, 1, dtype=torch.float32).pin_memory()
    losses = []

    def step_e2e(s_):
        z_pin.copy_(torch.from_numpy(zs[s_ % len(zs)]).to(torch.float32))      # train.py:56 torch.Tensor(z)
        t_pin.fill_(float(rs.uniform(0, 1)))
        zd = z_pin.to(dev, non_blocking=True)
        td = t_pin.to(dev, non_blocking=True)
        with ctx():
            losses.append(trainer.step(zd, td).item())                         # the .item() of train.py's log line

    step_e2e(0)
    ms_e2e = _timed(step_e2e, args.steps, world, dev)

    # parts, timed alone on rank 0's stream (every rank runs them to stay in step)
    n = gen.n_latent
    with torch.no_grad():
        w = gen.style(z_dev)
        lat = w[:, None, :].repeat(1, n, 1)
        ms_fwd = _timed(lambda _s: gen(lat, input_is_latent=True), 5, world, dev) / 5
        img0, _ = gen(lat, input_is_latent=True)
        with ctx():
            ms_reg_fwd = _timed(lambda _s: reg(img0.contiguous(memory_format=torch.channels_last) if amp else img0), 5, world, dev) / 5
    probe = torch.randn(b, 3, size, size, device=dev)

    def fwd_bwd(_s):
        l_ = lat.clone().requires_grad_(True)
        img, _ = gen(l_, input_is_latent=True)
        img.backward(probe)

    fwd_bwd(0)
    ms_fb = _timed(fwd_bwd, 5, world, dev) / 5
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    samples = world * b * args.steps
    g_ms = ms_fwd + ms_fb
    g_flop = 3 * 148.52e9 * b if size == 1024 else None            # fwd (orig) + fwd (edited) + data gradient
    ach = g_flop / (g_ms * 1e-3) / 1e12 if g_flop else None
    line = {
        "metric": "stylegan2_1024_walk_training_samples_per_sec", "value": samples / (ms_dev * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"train.py walk-module step (BASELINE cfg3): StyleGAN2-{size} fwd (no grad) + fwd/bwd + random-init ResNet-50 "
                               f"regressor x2 ({'bf16 autocast, channels_last' if amp else 'fp32 NCHW'}), BCE, linear w-walk, Adam, batch {b}/GPU",
                   "size": size, "batch_per_gpu": b, "global_batch": world * b,
                   "parallelism": f"dp{world} (latents sharded; one NCCL all-reduce of the walk gradient per step)",
                   "l2": "working set (activations kept for the backward, > 10 GB) is larger than the 126 MB L2; no explicit flush"},
        "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": b * 512 * 4 + b * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": losses[-1]},
        "gpu_launches": launches, "clocks": clocks, "allreduce_bytes": trainer.last_allreduce_bytes,
        "parts_ms": {"g_forward_inference": ms_fwd, "g_forward_training_plus_backward": ms_fb, "regressor_forward": ms_reg_fwd,
                     "regressor_and_rest": ms_dev / args.steps - g_ms},
        "roofline": {"kernel": "generator part of the step: inference forward + training forward + data-gradient backward (tcgen05 convs)",
                     "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": ach / peaks["tflops"] if ach else None, "traffic": None, "peak_source": peaks["source"],
                     "algorithmic_gflop_per_step": g_flop / 1e9 if g_flop else None, "launch_ms_total": g_ms},
        "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_panels(args, mlp=False):
    """BASELINE cfg1 (default 256 px, batch 4, 8 samples, 10 panels, linear walk, attribute Smiling) or cfg4 (--workload
    panels_mlp: MLP walk, 5 attributes x 10 panels, 1024 px): the vis_w.py sweep through TransformGraph.apply_alpha - the
    reference's semantics (G(w) and R(G(w)) recomputed for EVERY panel: 2 G + 1 R per edited image, transform_base.py:554-603)
    and the cached variant (--cache_original: G(w), R once per batch).  uint8 panels are copied to the host like
    vis_multi_image_batch_alphas does; PNG encoding is left out."""
    import importlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from latent2im_b200 import _native as nt
    from latent2im_b200 import graphs
    from latent2im_b200.synthetic import load_synthetic, synthetic_z
    from latent2im_b200.utils import util

    world, rank, local_rank, dev = _dist_setup()
    nt.load()
    size, b = args.size, args.batch
    num_panels = 10
    attrs = ["Smiling", "Young", "Male", "Eyeglasses", "Wavy_Hair"] if mlp else ["Smiling"]
    num_samples = b if mlp else 2 * b
    constants = importlib.import_module("latent2im_b200.graphs.stylegan_v2_real.constants")
    constants.resolution, constants.BATCH_SIZE, constants.compute_dtype = size, b, args.dtype
    constants.walk_is_mlp, constants.reg_amp, constants.allow_random_init = mlp, not args.reg_fp32, True
    constants.g_path = constants.reg_path = "/nonexistent (synthetic benchmark weights)"
    names, table = util._read_attr_table(os.path.join(ROOT, "latent2im_b200", "dataset", "attributes_celeba.txt"))
    np.random.seed(0)
    torch.manual_seed(1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = graphs.find_model_using_name("stylegan_v2_real", "face")(lr=1e-4, walk_type="linear", loss="l2", trainEmbed=False,
                                                                     attrList=attrs, attrTable=table, layers=None, stylegan_opts=None)
    load_synthetic(g.module.netG, seed=0)
    z_all = synthetic_z(world * num_samples, seed=0)[rank * num_samples:(rank + 1) * num_samples]
    alphas = np.linspace(0, 1, num_panels)
    hosts = [torch.empty(b, size, size, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def sweep(cache):
        n_img, k = 0, 0
        for start in range(0, num_samples, b):
            zs = z_all[start:start + b]
            for attr in attrs:
                index_ = table[attr] if len(attrs) > 1 else None
                cached = None
                for a_ in alphas:
                    ag = g.scale_test_alpha_for_graph(a_, zs)
                    z = torch.Tensor(zs).to(dev)                                   # host -> device per panel, as the reference
                    im, _, _ = g.apply_alpha({"z": z}, ag, name=attr, index_=index_, cached_original=cached)
                    if cache:
                        cached = g._last_original
                    u8 = torch.empty(im.shape[0], size, size, 3, device=dev, dtype=torch.uint8)
                    nt.check(nt.load().l2i_image_to_uint8(u8.data_ptr(), im.contiguous().data_ptr(), im.shape[0], size, size,
                                                          nt.stream_ptr(dev)), "image_to_uint8")
                    hosts[k & 1][:im.shape[0]].copy_(u8, non_blocking=True)
                    k += 1
                    n_img += im.shape[0]
        return n_img

    for _ in range(max(1, min(args.warmup, 2))):
        n_img = sweep(False)
        sweep(True)
    steps = max(1, min(args.steps, 20))
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = nt.launch_count()
    ms_ref_sem = _timed(lambda _s: sweep(False), steps, world, dev)
    launches = nt.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_cached = _timed(lambda _s: sweep(True), steps, world, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    images = world * n_img * steps
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not mlp:
        cpu = _cpu_panels_baseline(size, b, attrs)
    line = {
        "metric": f"stylegan2_{size}_vis_w_panel_images_per_sec", "value": images / (ms_ref_sem * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms_ref_sem / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": (f"vis_w.py panel render (BASELINE {'cfg4' if mlp else 'cfg1'}): StyleGAN2-{size}, batch {b}, {num_samples} samples/GPU, "
                                f"{len(attrs)} attribute(s) x {num_panels} panels, {'MLP' if mlp else 'linear'} w-walk, reference semantics = "
                                "2 G forwards + 1 ResNet-50 forward per edited image"), "size": size, "batch_per_gpu": b,
                   "images_per_step": world * n_img, "parallelism": f"dp{world} (samples sharded, no collective)",
                   "l2": "one sweep touches more than the 126 MB L2 at 1024 px; at 256 px batch 4 the activations are L2-resident by nature of cfg1"},
        "e2e": {"value": images / (ms_ref_sem * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_img // b * b * 512 * 4,
                "d2h_bytes_per_step": n_img * size * size * 3, "note": "the workload itself is host-buffer to host-buffer (z from numpy per panel, uint8 panels to pinned host)"},
        "cached_original": {"value": images / (ms_cached * 1e-3), "unit": UNIT, "ms_per_step": ms_cached / steps,
                            "what": "G(w) and R(G(w)) computed once per (batch, attribute) instead of once per panel (--cache_original)"},
        "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu,
        "ref_gpu_baseline": None if (args.no_ref_gpu or world > 1) else _ref_gpu_baseline(size, b),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _cpu_panels_baseline(size, b, attrs):
    """cfg1 on the host cores: oracle G (fp32) x 2 + torchvision ResNet-50 x 1 per panel, bounded sample of 2 panels of one batch."""
    import torch
    import torchvision
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    from latent2im_b200.synthetic import synthetic_state_dict, synthetic_walk_w, synthetic_z
    from oracle import GeneratorSpec, generator_forward_ref, mapping_ref
    from oracle.walks import walk_linear_ref
    cores = _host_cores()
    torch.set_num_threads(cores)
    spec = GeneratorSpec(size=size)
    sd = synthetic_state_dict({k: v.shape for k, v in Generator(size, 512, 8).state_dict().items()}, 0)
    walk_w = synthetic_walk_w(len(attrs), spec.n_latent, 512, seed=0)
    torch.manual_seed(1)
    reg = torchvision.models.resnet50(weights=None)
    reg.fc = torch.nn.Linear(2048, 40)
    reg = torch.nn.Sequential(reg, torch.nn.Sigmoid()).eval()
    z = torch.tensor(synthetic_z(b, seed=0), dtype=torch.float32)

    def panel(target):
        with torch.no_grad():
            w = mapping_ref(sd, z, spec)
            noise = lambda: [torch.randn(b, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(spec.num_layers)]
            img0 = generator_forward_ref(sd, w[:, None, :].repeat(1, spec.n_latent, 1), noise(), spec)
            delta = target - reg(img0)[:, [31]]
            lat = torch.stack(walk_linear_ref([w] * spec.n_latent, delta, walk_w), 1)
            return generator_forward_ref(sd, lat, noise(), spec)

    panel(0.5)
    t0 = time.perf_counter()
    for t in (0.0, 1.0):
        panel(t)
    dt = time.perf_counter() - t0
    return {"value": 2 * b / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"2 panels of one batch-{b} {size}px sweep (2 oracle G forwards + 1 torchvision ResNet-50 forward each), fp32, {dt:.1f}s"}


if __name__ == "__main__":
    a = _args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    elif a.workload in ("panels", "panels_mlp"):
        run_panels(a, mlp=a.workload == "panels_mlp")
    else:
        run_ours(a)
