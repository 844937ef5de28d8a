#!/usr/bin/env python
"""bench.py — ray-samples/s through the fused SDF+render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|render] [--rays R]

Default workload = BASELINE.json configs[1]: one TRAINING step on a 1024-ray batch with the config/synthetic.yml networks
and loss: forward (rays -> error-bounded sampler, 5 x 128 sdf-evals/ray -> SDF + grad_x + radiance on the 97 composited
samples/ray -> compositing -> 3R eikonal points) + I2SDFLoss + backward incl. the second-order terms + Adam(eps=1e-15)
+ weight re-pack.  Synthetic rays / targets: W-sharp weights (geometric init seed 0, density.beta =
0.01 so all five sampler rounds run).  value = full-path ray-samples/s = rays * 97 / step time, whole job (all ranks).
`--mode render` times the eval forward render of the same batch instead (also reported as `render` in train mode).

N > 1 (torchrun, one rank per GPU): rays shard across ranks, every rank processes its own 1024-ray batch (weak
scaling); training all-reduces the flat 3.2 MB gradient once per step over NCCL, inference has no collective;
time = max over ranks of the device time of the K steps.

--impl reference: the reference's CPU implementation of the same path.  The reference is Python/PyTorch and cannot
travel to the GPU box, so this arm times the oracle port (oracle/i2sdf_oracle.py, pinned bit-for-bit against the
reference on the golden fixtures) on the host cores, one bounded sample of the workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "full-path ray-samples/sec through SDF+render MLPs"
UNIT = "ray-samples/s"
N_COMPOSITED = 97               # N_samples + 2 + N_samples_extra - 1  (model/network/__init__.py:99-100)
# algorithmic FLOP (2 x MAC, dense layers only) per unit: one sampler sdf evaluation (sdf row of the last
# layer only); one composited ray-sample = SDF fwd + grad_x sweep + radiance (+ light head)
FLOPS = {"synthetic": dict(sdf_eval=918016, ray_sample=2506752),
         "synthetic_light_mask": dict(sdf_eval=2 * (393472 - 65536), ray_sample=1917184)}
FLOP_PER_SDF_EVAL = FLOPS["synthetic"]["sdf_eval"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--config", default="synthetic", choices=["synthetic", "synthetic_light_mask"],
                    help="synthetic (default, BASELINE.json configs[1]/[2]) or synthetic_light_mask (configs[3]: 6x256 SDF, 3x256 "
                         "radiance, light-mask head, light_mask_weight 0.5)")
    ap.add_argument("--cpu-rays", type=int, default=128, help="rays per step of the CPU arms (bounded sample)")
    ap.add_argument("--mode", default="train", choices=["render", "train"],
                    help="train (default, BASELINE.json configs[1]): full training step on a 1024-ray batch (forward + I2SDFLoss + "
                         "backward incl. second order + gradient all-reduce + Adam + weight re-pack); "
                         "render: eval forward render of a 1024-ray batch (configs[2] batch shape)")
    return ap.parse_args()


def build_params(beta=0.01, name="synthetic"):
    """W-sharp: reference geometric init (seed 0) with density.beta = 0.01."""
    from i2sdf_b200 import configs
    from i2sdf_b200.network import I2SDFNetwork
    conf = configs.model_conf(name)
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = I2SDFNetwork(conf)
    with torch.no_grad():
        m.density.beta.fill_(beta)
    return conf, m


# ----------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference path)
# ----------------------------------------------------------------------------------------------------
def pick_cpu_threads(run_once):
    """The reference would run with torch's default thread count (= all cores); on a many-core host these small
    per-op tensors scale badly, so probe a few counts on a tiny sample and keep the fastest (this only ever helps the
    CPU arm)."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        run_once()
        t0 = time.perf_counter()
        run_once()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


# The CPU arms time the oracle PORT (the reference tree does not exist on the GPU box).  Port vs the unmodified reference, same inputs,
# same 8 threads, interleaved repetitions in the build container where both run (tools/port_vs_reference.py): the port
# runs at 1.0-1.1x the reference's speed for the eval render and for the training step, i.e. the CPU baseline is not understated.
PORT_VS_REFERENCE = {"render": "1.0-1.1", "train": "1.0-1.1"}


def cpu_arm(conf, model, rays, steps, warmup):
    from oracle import i2sdf_oracle as orc
    spec = orc.spec_from_model_conf(conf, use_normal=False)
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    probe = orc.synthetic_rays(16, seed=2)
    with torch.no_grad():
        pick_cpu_threads(lambda: orc.render(spec, P, probe, training=False))
    inp = orc.synthetic_rays(rays, seed=1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.render(spec, P, inp, training=False)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return dict(value=rays * N_COMPOSITED * len(times) / total, ms_per_step=1e3 * total / len(times),
                cores=torch.get_num_threads(),
                sample=f"{rays} of the 1024 rays per step (same weights, same ray distribution), eval forward, "
                       f"{len(times)} steps after {warmup} warm-up, torch CPU fp32, best of 8/16/32/64/all host threads "
                       f"(picked {torch.get_num_threads()} of {os.cpu_count()}); oracle port = {PORT_VS_REFERENCE['render']}x the unmodified "
                       f"reference's speed where both run (build container)")


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d.get("bf16_tflops", 1590.0), bf16_tflops_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=d.get("hbm_gbs", 6650.0), source="MEASURED_PEAKS.json (of measured)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="B200_PROFILING.md fallback (of fallback)")


def make_train_gt(R, seed, light=False):
    g = torch.Generator().manual_seed(seed)
    gt = {
        "rgb": torch.rand(R, 3, generator=g),
        "depth": torch.rand(R, generator=g) * 2 + 0.5,
        "depth_mask": torch.ones(R, dtype=torch.bool),
        "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1),
        "normal_mask": torch.ones(R, dtype=torch.bool),
    }
    if light:
        gt["light_mask"] = (torch.rand(R, 1, generator=g) > 0.9).float()
    return gt


def loss_weights(name):
    from i2sdf_b200 import configs
    src = configs.LOSS_SYNTHETIC_LIGHT_MASK if name == "synthetic_light_mask" else configs.LOSS_SYNTHETIC
    keys = ("eikonal_weight", "smooth_weight", "depth_weight", "normal_weight", "bubble_weight", "light_mask_weight")
    return src, {k: v for k, v in src.items() if k in keys}


def cpu_train_arm(conf, model, rays, steps, warmup, name="synthetic"):
    """Reference training step (forward + I2SDFLoss + backward incl. double backward) on the CPU: oracle port."""
    from i2sdf_b200 import configs
    from oracle import i2sdf_oracle as orc
    spec = orc.spec_from_model_conf(conf, use_normal=True)
    P0 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    inp = orc.synthetic_rays(rays, seed=1, train_layout=True)
    light = name == "synthetic_light_mask"
    gt = make_train_gt(rays, 7, light)
    lw = loss_weights(name)[1]

    def tape(n_final_guess=None):
        R = rays
        return {"jitter": torch.rand(R, spec.n_samples_eval), "u_final": torch.rand(R, spec.n_samples),
                "extra_perm": lambda n: torch.randperm(n)[:spec.n_samples_extra], "eik_idx": torch.randint(98, (R,)),
                "eik_uniform": torch.empty(R, 3).uniform_(-spec.bounding_sphere, spec.bounding_sphere),
                "nbr_uniform": torch.empty(R, 3).uniform_(-0.005, 0.005)}

    def one():
        P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
        out = orc.render(spec, P, inp, training=True, tape=tape())
        loss = orc.recon_loss(out, gt, smooth_active=False, **lw)
        loss.backward()

    probe_inp = orc.synthetic_rays(32, seed=2, train_layout=True)

    def probe():
        P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
        R = 32
        tp = {"jitter": torch.rand(R, spec.n_samples_eval), "u_final": torch.rand(R, spec.n_samples),
              "extra_perm": lambda n: torch.randperm(n)[:spec.n_samples_extra], "eik_idx": torch.randint(98, (R,)),
              "eik_uniform": torch.empty(R, 3).uniform_(-3, 3), "nbr_uniform": torch.empty(R, 3).uniform_(-0.005, 0.005)}
        out = orc.render(spec, P, probe_inp, training=True, tape=tp)
        orc.recon_loss(out, make_train_gt(R, 3, light), smooth_active=False, **lw).backward()

    pick_cpu_threads(probe)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=rays * N_COMPOSITED * len(times) / total, ms_per_step=1e3 * total / len(times), cores=torch.get_num_threads(),
                sample=f"{rays} of the 1024 rays per step, full training step (forward + loss + backward; no optimizer), "
                       f"{len(times)} steps after {warmup} warm-up, torch CPU fp32, best of 8/16/32/64/all host threads "
                       f"(picked {torch.get_num_threads()} of {os.cpu_count()}); oracle port = {PORT_VS_REFERENCE['train']}x the unmodified "
                       f"reference's speed where both run (build container)")


def gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from oracle import i2sdf_oracle as orc
    from i2sdf_b200 import configs
    from i2sdf_b200.network import I2SDFLoss
    from i2sdf_b200.parallel import allreduce_gradients
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    train = args.mode == "train"
    conf, model = build_params(name=args.config)
    light = args.config == "synthetic_light_mask"
    flops = FLOPS[args.config]
    if train:
        model.use_normal = True          # trainer sets it from loss.normal_weight (model/trainer/recon.py:34-35)
    cpu_snapshot = {k: v.detach().clone() for k, v in model.state_dict().items()} if rank == 0 else None
    model_gpu = model.to(dev)
    model_gpu.train(train)
    init_state = {k: v.detach().clone() for k, v in model_gpu.state_dict().items()}
    R = args.rays
    # rank-specific rays: the global batch is world * R rays sharded across ranks
    inp_host = {k: v.pin_memory() for k, v in orc.synthetic_rays(R, seed=1 + rank, train_layout=train).items()}
    gt_host = {k: v.pin_memory() for k, v in make_train_gt(R, 7 + rank, light).items()} if train else {}
    inp_dev = {k: v.to(dev) for k, v in inp_host.items()}
    gt_dev = {k: v.to(dev) for k, v in gt_host.items()}
    core = model_gpu._ready_core()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    out_host = {}
    if train:
        loss_fn = I2SDFLoss(**loss_weights(args.config)[0])
        # Adam(lr, eps=1e-15) as model/trainer/recon.py:201-207; i2sdf_b200.optim.Adam is the same update rule and state layout as
        # torch.optim.Adam in ONE launch for all 44 parameter tensors (I2SDF_TORCH_ADAM=1: torch's fused Adam, 4 launches)
        if os.environ.get("I2SDF_TORCH_ADAM") == "1":
            opt = torch.optim.Adam(model_gpu.parameters(), lr=5.0e-4, eps=1e-15, fused=True)
        else:
            from i2sdf_b200.optim import Adam
            opt = Adam(model_gpu.parameters(), lr=5.0e-4, eps=1e-15)
        loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def train_step(inp, gt):
        out = model_gpu(inp)
        loss = loss_fn(out, gt, 0)["loss"]
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            allreduce_gradients(model_gpu.parameters())
        opt.step()
        return loss

    def step_resident():
        if train:
            return train_step(inp_dev, gt_dev)
        return model_gpu(inp_dev)

    def step_e2e():
        d = {k: v.to(dev, non_blocking=True) for k, v in inp_host.items()}
        if train:
            g = {k: v.to(dev, non_blocking=True) for k, v in gt_host.items()}
            loss = train_step(d, g)
            loss_host.copy_(loss.detach(), non_blocking=True)
            return loss
        out = model_gpu(d)
        for k, v in out.items():
            if k not in out_host:
                out_host[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
            out_host[k].copy_(v, non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        if profile:
            core.profile(True)
        for a, b in ev:
            flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        barrier()
        prof = core.profile_read() if profile else None
        if profile:
            core.profile(False)
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return ms, prof

    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    clk = ClockSampler(local_rank)
    clk.start()
    core.rounds_log.clear()
    ms_res, prof = timed(step_resident, args.steps, profile=True)
    rounds_seen = sorted(set(core.rounds_log))
    rounds_total = sum(core.rounds_log)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = clk.stop()
    ms_render = 0.0
    if train:       # also report the forward-render throughput of the same networks (inference: no collective)
        model_gpu.load_state_dict(init_state)          # back to the W-sharp weights (Adam steps changed beta / the surface)
        model_gpu.eval()
        inp_eval = {k: v.to(dev) for k, v in orc.synthetic_rays(R, seed=1 + rank).items()}
        with torch.no_grad():
            for _ in range(3):
                model_gpu(inp_eval)
            ms_render, _ = timed(lambda: model_gpu(inp_eval), args.steps)
        model_gpu.train(True)
    t = torch.tensor([ms_res, ms_e2e, ms_render], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e, ms_render = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        return None
    K = args.steps
    value = world * R * N_COMPOSITED * K / (ms_res * 1e-3)
    e2e_value = world * R * N_COMPOSITED * K / (ms_e2e * 1e-3)
    pk = peaks()
    # dominant kernel: sampler SDF evaluations (73 % of the forward path's FLOPs)
    sdf = prof["sampler_sdf"]
    # launches of rounds the device-side convergence word switched off return at once: only ACTIVE launches count (training: the
    # rounds each timed step really ran, read back by sampler_resolve; eval: the W-sharp weights run all of them)
    sdf_launches = max(rounds_total if (train and rounds_total > 0) else sdf["launches"], 1)
    pts_per_launch = R * 128
    per_launch_ms = sdf["ms"] / sdf_launches
    achieved = pts_per_launch * flops["sdf_eval"] / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"] if core.uses_tensor_cores else 72.0
    launches = sum(v["launches"] for v in prof.values())
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01c_tc_sdf_traffic.json")
    if core.uses_tensor_cores and os.path.exists(tpath) and R == 1024 and not light:
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]      # from the committed ncu --set full capture
    h2d = sum(v.numel() * v.element_size() for v in list(inp_host.values()) + list(gt_host.values()))
    d2h = 4 if train else sum(v.numel() * v.element_size() for v in out_host.values())
    if train:
        workload = ("C2: training step on a 1024-ray batch, config/synthetic.yml networks and loss weights "
                    "(rgb L1 + eikonal + depth + normal/angular; steps < 50k: no bubble/smooth terms): forward "
                    "(error-bounded sampler 5x128 sdf-evals/ray, main pass on 97 samples/ray with saved activations, 3R eikonal "
                    "points) + I2SDFLoss + backward incl. second-order terms (fused tensor-core chain + one weight-gradient launch) "
                    "+ Adam(eps=1e-15) step (one launch) + weight re-pack"
                    + (" + one flat NCCL gradient all-reduce" if world > 1 else ""))
    else:
        workload = ("C2/C3 batch shape: 1024-ray forward render, config/synthetic.yml networks (8x256 SDF + 4x256 radiance), "
                    "eval layout: sampler 5 rounds = 640 sdf-evals/ray + 97 composited samples/ray")
    if light:
        workload = workload.replace("C2: training step", "C4: training step").replace("C2/C3 batch shape", "C4 batch shape") \
            .replace("config/synthetic.yml networks (8x256 SDF + 4x256 radiance)", "config/synthetic_light_mask.yml networks (6x256 SDF + 3x256 radiance + light-mask head)") \
            .replace("config/synthetic.yml networks and loss weights", "config/synthetic_light_mask.yml networks (6x256 SDF, 3x256 radiance, light-mask head) and loss weights (+ light-mask BCE 0.5)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (bf16 hi/lo split products on tcgen05, fp32 accumulate)" if core.uses_tensor_cores else "f32",
        "data": "synthetic",
        "config": {"workload": workload, "mode": args.mode, "network_config": args.config,
                   "weights": ("W-sharp at step 0: reference geometric init (seed 0), density.beta=0.01 (all 5 sampler rounds run); training then "
                               "moves the weights every step (Adam, lr 5e-4) and the packed copies follow, so later steps may converge in fewer "
                               "rounds: see sampler_rounds_in_timed_steps") if train else
                              "W-sharp: reference geometric init (seed 0), density.beta=0.01 so all 5 sampler rounds run",
                   "rays_per_gpu": R, "global_rays": world * R, "samples_per_ray_composited": N_COMPOSITED,
                   "sdf_evals_per_ray": 5 * 128 + N_COMPOSITED,
                   "sampler_rounds_in_timed_steps": rounds_seen if train else "5 (eval: weights fixed)",
                   "parallelism": f"ray-sharded x{world}, " + ("one flat gradient all-reduce per step" if train else "no collective (inference)"),
                   "tensor_cores": {"sampler_sdf": core.uses_tensor_cores, "main_pass": core.uses_tensor_cores_main,
                                    "backward": bool(train and core.fused_main)},
                   "l2": "flushed between timed iterations (256 MB write)", "timing": "CUDA events per step, max over ranks"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "kernel_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
        "kernel_launches_per_step": {k: v["launches"] / K for k, v in prof.items()},
        "roofline": {"bound": "tensor", "kernel": "tc_sdf8_kernel (sampler SDF evaluations)" if core.uses_tensor_cores else "mlp_tile_kernel",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per launch (ncu --set full, profiles/r01c_tc_sdf_traffic.json)",
                     "flop_per_launch": pts_per_launch * flops["sdf_eval"], "ms_per_launch": per_launch_ms,
                     "peak_source": pk["source"] + (", sustained bf16 (kernel timed inside a step)" if core.uses_tensor_cores else "; fp32 FMA peak 148 SM x 128 FMA x 2 x 1.9 GHz"),
                     "note": "achieved counts ALGORITHMIC flops (1 MAC = 2 flop); the kernel issues 3 bf16 MMAs per MAC, so 1/3 of the bf16 peak is this precision scheme's ceiling"},
        "clocks": clocks,
    }
    if train and core.fused_main and prof["weight_grads"]["launches"] > 0 and not light:
        # the two other heavy kernels of the training step, both bounded by HBM: algorithmic bytes = plane slots read / written
        tile = lambda m: (m + 127) // 128                                                     # noqa: E731
        big = lambda m: tile(m) * 131072                                                      # noqa: E731
        m_main, m_eik = R * N_COMPOSITED, 3 * R
        wg_bytes = (39 * big(m_main) + tile(m_main) * (2 * 24576 + 16384)) + (30 * big(m_eik) + tile(m_eik) * 2 * 24576)
        bw_bytes = 53.5 * big(m_main) + 48 * big(m_eik)
        hbm = pk["hbm_gbs"]
        more = []
        for name, key, nbytes, what in (("wgrad_planes_kernel", "weight_grads", wg_bytes, "every (P, X) slot pair read once; dW written by atomics (negligible)"),
                                        ("tc_bwd_kernel", "backward_chain", bw_bytes, "H/Q/C slots read, HD re-read, HD/P/PC/FB slots written")):
            t_ms = prof[key]["ms"] / K
            more.append({"kernel": name, "bound": "hbm", "achieved": nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                         "frac": (nbytes / (t_ms * 1e-3) / 1e9 / hbm) if t_ms > 0 else None, "bytes_per_step": nbytes, "ms_per_step": t_ms,
                         "launches_per_step": prof[key]["launches"] / K, "traffic_model": what})
        line["roofline_more"] = more
    if train and ms_render > 0:
        line["render"] = {"value": world * R * N_COMPOSITED * K / (ms_render * 1e-3), "unit": UNIT, "ms_per_step": ms_render / K,
                          "workload": "eval forward render of the same 1024-ray batch (no backward, no collective), device-resident inputs"}
    model_c = type("S", (), {"state_dict": lambda self: cpu_snapshot})()
    if train:
        cb = cpu_train_arm(conf, model_c, args.cpu_rays, steps=2, warmup=1, name=args.config)
    else:
        cb = cpu_arm(conf, model_c, args.cpu_rays, steps=2, warmup=1)
    line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"]}
    return line


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if rank != 0:
            return
        conf, model = build_params(name=args.config)
        if args.mode == "train":
            cb = cpu_train_arm(conf, model, args.cpu_rays, steps=args.steps, warmup=args.warmup, name=args.config)
        else:
            cb = cpu_arm(conf, model, args.cpu_rays, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"same as --impl ours --mode {args.mode} (1024-ray batch, synthetic.yml, W-sharp); each step is a "
                                       "bounded sample of the batch on the host CPU (see cpu_baseline.sample)", "mode": args.mode},
                "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (i2sdf_b200 has no CPU path)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    line = gpu_arm(args, rank, world, local_rank)
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
