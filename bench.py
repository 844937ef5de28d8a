#!/usr/bin/env python
"""bench.py — ray-samples/s through the fused SDF+render path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|render] [--rays R]

Default workload = BASELINE.json configs[1]: one TRAINING step on a 1024-ray batch with the config/synthetic.yml networks
and loss: forward (rays -> error-bounded sampler, 5 x 128 sdf-evals/ray -> SDF + grad_x + radiance on the 97 composited
samples/ray -> compositing -> 3R eikonal points) + I2SDFLoss + backward incl. the second-order terms + Adam(eps=1e-15)
+ weight re-pack.  Synthetic rays / targets: W-sharp weights (geometric init seed 0, density.beta =
0.01 so all five sampler rounds run).  value = full-path ray-samples/s = rays * 97 / step time, whole job (all ranks).
`--mode render` times the eval forward render of the same batch instead (also reported as `render` in train mode).

N > 1 (torchrun, one rank per GPU): rays shard across ranks.  The headline `value` is WEAK scaling (every rank processes
its own 1024-ray batch); the same line carries `strong_scaling` (ONE 1024-ray batch split over the ranks,
1024 / N rays per GPU).  Training all-reduces the flat 3.2 MB gradient once per step over NCCL (parallel.GradBucket: the
gradients live in one persistent buffer, no flatten / un-flatten), inference has no collective; time = max over ranks of
the device time of the K steps.

--impl reference: the reference's CPU implementation of the same path.  The reference is Python/PyTorch and cannot
travel to the GPU box, so this arm times the oracle port (oracle/i2sdf_oracle.py, pinned bit-for-bit against the
reference on the golden fixtures) on the host cores: the SAME config as the GPU arm (1024 rays per step, same weights, rays and
targets, training step incl. the Adam update).

Algorithmic FLOP model (flop_model below; 2 x MAC of the dense layers only, PE / activations / compositing excluded), synthetic.yml:
  sdf-eval (sampler)      918 016   = 2 x (39x256 + 256x256x2 + 256x217 + 256x256x4 + 256)           sdf row of the last layer only
  ray-sample forward    2 506 752   = 2 x (524 544 SDF incl. 256 feature rows + 459 008 grad_x reverse sweep + 269 824 radiance)
  eikonal point forward 1 836 032   = 2 x (459 008 + 459 008)                                        sdf + grad_x, no features / radiance
  ray-sample backward   4 991 488   = 2 x (2 x 269 824 radiance dX + dW, 458 752 tangent pass, 65 536 feature adjoint,
                                           448 768 adjoint chain, 2 x 458 752 + 65 536 weight gradients)   first + second order
  eikonal point backward 3 650 560  = 2 x (458 752 + 256 + 448 768 + 2 x 458 752)
  training step  = rounds x 128 R x sdf-eval + 97 R x (fwd + bwd) + 3 R x (eik fwd + eik bwd)   = 1.364 TFLOP at R = 1024, 5 rounds
  eval render    = rounds x 128 R x sdf-eval + 97 R x ray-sample forward                          = 0.851 TFLOP at R = 1024, 5 rounds
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


def flop_model(conf):
    """Algorithmic FLOPs (2 x MAC, dense layers only) per unit of work, from the network dims of a model conf (see the module doc)."""
    imp, ren = conf["implicit_network"], conf["rendering_network"]
    ex, ed = 3 + 6 * imp["multires"], 3 + 6 * ren["multires"]
    dims = [ex] + list(imp["dims"])
    skip = list(imp.get("skip_in", ()))
    hid = 0                                     # MACs of the SDF hidden layers (a layer feeding a skip concat is ex narrower)
    for l in range(len(dims) - 1):
        out = dims[l + 1] - (ex if (l + 1) in skip else 0)
        hid += dims[l] * out
    feat = conf["feature_vector_size"]
    sdf_fwd_full = hid + dims[-1] * (1 + feat)              # as the reference executes it (257 output rows)
    sdf_row = hid + dims[-1]                                # sdf row only = the grad_x reverse sweep's MACs too
    cdims = [ed + feat] + list(ren["dims"]) + [3]
    col = sum(a * b for a, b in zip(cdims[:-1], cdims[1:]))
    light = conf.get("light_network")
    lmac = 0
    if light:
        ld = [feat] + list(light["dims"]) + [1]              # an ImplicitNetwork with dims [256, 128, 1] (network/__init__.py:29-32)
        lmac = sum(a * b for a, b in zip(ld[:-1], ld[1:]))
    adj = hid - dims[0] * (dims[1] - (ex if 1 in skip else 0))          # adjoint chain stops in front of layer 0
    return dict(sdf_eval=2 * sdf_row, ray_sample=2 * (sdf_fwd_full + sdf_row + col + lmac), eik_fwd=2 * (sdf_row + sdf_row),
                ray_sample_bwd=2 * (2 * col + hid + feat * dims[-1] + adj + 2 * hid + feat * dims[-1] + 2 * lmac),
                eik_bwd=2 * (hid + dims[-1] + adj + 2 * hid))


def step_flops(fm, R, rounds, train):
    f = rounds * 128 * R * fm["sdf_eval"] + N_COMPOSITED * R * fm["ray_sample"]
    if train:
        f += N_COMPOSITED * R * fm["ray_sample_bwd"] + 3 * R * (fm["eik_fwd"] + fm["eik_bwd"])
    return f


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
    ap.add_argument("--cpu-rays", type=int, default=0, help="rays per step of the CPU arms (0 = --rays: the same config as the GPU arm)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="training: also time the step replayed as ONE CUDA graph (i2sdf_b200.graph.GraphedTrainStep).  auto (default): only for the "
                         "strong-scaling block of a multi-GPU run (1024 / N rays per GPU: issued from Python that step is host-bound); on: for the weak "
                         "block as well; off: never.  Where both ran the faster one is the block's value, the other is reported beside it")
    ap.add_argument("--bubble", type=int, default=0, help="training variant of steps 50k-150k (config/synthetic.yml:22-23): N bubble points per "
                                                          "step through the SDF (bubble_weight 0.5) and the smoothness term switched on")
    ap.add_argument("--grid-res", type=int, default=256, help="--mode grid: resolution of the uniform SDF grid (reference meshes use 100 .. 512)")
    ap.add_argument("--mode", default="train", choices=["render", "train", "grid"],
                    help="train (default, BASELINE.json configs[1]): full training step on a 1024-ray batch (forward + I2SDFLoss + "
                         "backward incl. second order + gradient all-reduce + Adam + weight re-pack); "
                         "render: eval forward render of a 1024-ray batch (configs[2] batch shape); "
                         "grid: SDF of a uniform grid for mesh extraction (SURVEY.md §8(f)-3), reported in points/s - a secondary line, "
                         "not the BASELINE metric")
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


def cpu_arm(conf, model, rays, steps, warmup, full_rays=1024):
    from i2sdf_b200.synthetic import synthetic_rays
    from oracle import i2sdf_oracle as orc
    spec = orc.spec_from_model_conf(conf, use_normal=False)
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    probe = synthetic_rays(16, seed=2)
    with torch.no_grad():
        pick_cpu_threads(lambda: orc.render(spec, P, probe, training=False))
    inp = synthetic_rays(rays, seed=1)
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
                same_config=(rays == full_rays),
                sample=f"{rays} of the {full_rays} rays per step (same weights, same rays), eval forward, "
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
    from i2sdf_b200.synthetic import make_train_gt as mk
    return mk(R, seed, light)


def loss_weights(name):
    from i2sdf_b200 import configs
    src = configs.LOSS_SYNTHETIC_LIGHT_MASK if name == "synthetic_light_mask" else configs.LOSS_SYNTHETIC
    keys = ("eikonal_weight", "smooth_weight", "depth_weight", "normal_weight", "bubble_weight", "light_mask_weight")
    return src, {k: v for k, v in src.items() if k in keys}


def cpu_train_arm(conf, model, rays, steps, warmup, name="synthetic", full_rays=1024):
    """Reference training step (forward + I2SDFLoss + backward incl. double backward + Adam) on the CPU: oracle port."""
    from i2sdf_b200.synthetic import synthetic_rays
    from oracle import i2sdf_oracle as orc
    spec = orc.spec_from_model_conf(conf, use_normal=True)
    P0 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    inp = synthetic_rays(rays, seed=1, train_layout=True)
    light = name == "synthetic_light_mask"
    gt = make_train_gt(rays, 7, light)
    lw = loss_weights(name)[1]

    def tape(n_final_guess=None):
        R = rays
        return {"jitter": torch.rand(R, spec.n_samples_eval), "u_final": torch.rand(R, spec.n_samples),
                "extra_perm": lambda n: torch.randperm(n)[:spec.n_samples_extra], "eik_idx": torch.randint(98, (R,)),
                "eik_uniform": torch.empty(R, 3).uniform_(-spec.bounding_sphere, spec.bounding_sphere),
                "nbr_uniform": torch.empty(R, 3).uniform_(-0.005, 0.005)}

    # the parameters persist and move with Adam(eps=1e-15) (model/trainer/recon.py:201-207), as on the GPU arm
    Pt = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    opt = torch.optim.Adam(list(Pt.values()), lr=5.0e-4, eps=1e-15)

    def one():
        out = orc.render(spec, Pt, inp, training=True, tape=tape())
        loss = orc.recon_loss(out, gt, smooth_active=False, **lw)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    probe_inp = synthetic_rays(32, seed=2, train_layout=True)

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
                same_config=(rays == full_rays),
                sample=f"{rays} of the {full_rays} rays per step, full training step (forward + loss + backward incl. second order + Adam), "
                       f"{len(times)} steps after {warmup} warm-up, torch CPU fp32, best of 8/16/32/64/all host threads "
                       f"(picked {torch.get_num_threads()} of {os.cpu_count()}); oracle port = {PORT_VS_REFERENCE['train']}x the unmodified "
                       f"reference's speed where both run (build container)")


def gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from i2sdf_b200.network import I2SDFLoss
    from i2sdf_b200.parallel import GradBucket
    from i2sdf_b200.synthetic import bubble_points, synthetic_rays
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    train = args.mode == "train"
    conf, model = build_params(name=args.config)
    light = args.config == "synthetic_light_mask"
    fm = flop_model(conf)
    if train:
        model.use_normal = True          # trainer sets it from loss.normal_weight (model/trainer/recon.py:34-35)
    cpu_snapshot = {k: v.detach().clone() for k, v in model.state_dict().items()} if rank == 0 else None
    model_gpu = model.to(dev)
    model_gpu.train(train)
    init_state = {k: v.detach().clone() for k, v in model_gpu.state_dict().items()}
    core = model_gpu._ready_core()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    if train:
        lw = dict(loss_weights(args.config)[0])
        if args.bubble > 0:              # steps 50k-150k of config/synthetic.yml: bubble loss on, smoothness on
            lw.update(bubble_weight=0.5, smooth_weight=lw.get("smooth_weight") or 0.01, smooth_iter=0, min_bubble_iter=0, max_bubble_iter=None)
        loss_fn = I2SDFLoss(**lw)
        loss_step = 10 ** 6 if args.bubble > 0 else 0
        # Adam(lr, eps=1e-15) as model/trainer/recon.py:201-207; i2sdf_b200.optim.Adam is the same update rule and state layout as
        # torch.optim.Adam in ONE launch for all 44 parameter tensors (I2SDF_TORCH_ADAM=1: torch's fused Adam, 4 launches)
        if os.environ.get("I2SDF_TORCH_ADAM") == "1":
            opt = torch.optim.Adam(model_gpu.parameters(), lr=5.0e-4, eps=1e-15, fused=True)
        else:
            from i2sdf_b200.optim import Adam
            opt = Adam(model_gpu.parameters(), lr=5.0e-4, eps=1e-15)
        bucket = GradBucket(model_gpu.parameters())     # every .grad is a view of ONE flat buffer: all-reduced in place, read by Adam
        loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(R, first_ray, global_rays, use_graph=False):
        """K timed steps on this rank's R rays [first_ray, first_ray + R) of a global batch: device-resident and end-to-end."""
        # the global batch is generated identically on every rank, each rank keeps its shard
        full = synthetic_rays(global_rays, seed=1, train_layout=train)
        gt_full = make_train_gt(global_rays, 7, light) if train else {}
        sl = slice(first_ray, first_ray + R)
        if train:
            inp_host = {k: v[sl].contiguous().pin_memory() for k, v in full.items()}
        else:
            inp_host = {"uv": full["uv"][:, sl].contiguous().pin_memory(), "pose": full["pose"].pin_memory(), "intrinsics": full["intrinsics"].pin_memory()}
        gt_host = {k: v[sl].contiguous().pin_memory() for k, v in gt_full.items()}
        if train and args.bubble > 0:
            inp_host["pointcloud"] = bubble_points(args.bubble, seed=11 + rank).pin_memory()
        inp_dev = {k: v.to(dev) for k, v in inp_host.items()}
        gt_dev = {k: v.to(dev) for k, v in gt_host.items()}
        out_host = {}

        def train_step(inp, gt):
            out = model_gpu(inp)
            loss = loss_fn(out, gt, loss_step)["loss"]
            bucket.zero()
            loss.backward()
            bucket.allreduce()               # gathers every .grad into the flat buffer; all-reduces it when world > 1
            opt.step()
            return loss

        def step_resident():
            return train_step(inp_dev, gt_dev) if train else model_gpu(inp_dev)

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

        host = {}

        def timed(fn, steps, profile=False):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            if profile:
                core.profile(True)
            h0 = time.perf_counter()
            for a, b in ev:
                flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
                a.record()
                fn()
                b.record()
            host["enqueue_ms"] = (time.perf_counter() - h0) * 1e3 / steps       # host time to ENQUEUE a step (no device wait except the sampler's round count)
            barrier()
            prof = core.profile_read() if profile else None
            if profile:
                core.profile(False)
            return sum(a.elapsed_time(b) for a, b in ev), prof

        model_gpu.load_state_dict(init_state)          # every measurement starts from the W-sharp weights
        if train:
            opt.state.clear()
        for _ in range(max(args.warmup, 3)):
            step_resident()
            step_e2e()
        core.rounds_log.clear()
        ms_res, prof = timed(step_resident, args.steps, profile=True)
        host_ms = host.get("enqueue_ms")
        rounds = list(core.rounds_log)
        ms_e2e, _ = timed(step_e2e, args.steps)
        h2d = sum(v.numel() * v.element_size() for v in list(inp_host.values()) + list(gt_host.values()))
        d2h = 4 if train else sum(v.numel() * v.element_size() for v in out_host.values())
        ms_ar = None
        if train and world > 1:          # the gradient all-reduce on its own: the bucket, back to back, device time
            ar = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            barrier()
            for a, b in ar:
                a.record()
                dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG)
                b.record()
            barrier()
            ms_ar = sorted(a.elapsed_time(b) for a, b in ar)[len(ar) // 2]
        res = dict(ms_res=ms_res, ms_e2e=ms_e2e, prof=prof, rounds=rounds, h2d=h2d, d2h=d2h, host_ms=host_ms, ms_ar=ms_ar, graph=None)
        if train and use_graph:
            # the same step as ONE CUDA graph (forward + loss + backward + all-reduce + Adam + re-pack captured once, replayed per step)
            from i2sdf_b200.graph import GraphedTrainStep
            gstep = GraphedTrainStep(model_gpu, loss_fn, opt, inp_dev, gt_dev, current_step=loss_step, bucket=bucket, warmup=2)

            def g_resident():
                return gstep(inp_dev, gt_dev)

            def g_e2e():
                loss = gstep(inp_host, gt_host)          # pinned host batches: the H2D copies into the graph's static inputs are part of the step
                loss_host.copy_(loss.detach(), non_blocking=True)
                return loss
            for _ in range(max(args.warmup, 3)):
                g_resident()
                g_e2e()
            core.rounds_log.clear()
            g_ms, _ = timed(g_resident, args.steps)
            g_host = host.get("enqueue_ms")
            g_rounds = list(core.rounds_log)
            g_ms_e2e, _ = timed(g_e2e, args.steps)
            gstep.finish()
            res["graph"] = dict(ms_res=g_ms, ms_e2e=g_ms_e2e, host_ms=g_host, rounds=g_rounds)
            del gstep
        return res

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    R = args.rays
    clk = ClockSampler(local_rank)
    clk.start()
    weak = measure(R, rank * R, world * R, use_graph=(args.graph == "on"))      # weak scaling: R rays per rank, a global batch of world * R
    clocks = clk.stop()
    strong = None
    if world > 1:                                       # strong scaling: ONE R-ray batch split over the ranks
        from i2sdf_b200.parallel import shard_bounds
        lo, hi = shard_bounds(R, rank, world)
        strong = measure(hi - lo, lo, R, use_graph=(args.graph != "off"))
    ms_render = 0.0
    if train:       # also report the forward-render throughput of the same networks (inference: no collective)
        model_gpu.load_state_dict(init_state)          # back to the W-sharp weights (Adam steps changed beta / the surface)
        model_gpu.eval()
        inp_eval = {k: v.to(dev) for k, v in synthetic_rays(R, seed=1 + rank).items()}
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        with torch.no_grad():
            for _ in range(3):
                model_gpu(inp_eval)
            barrier()
            for a, b in ev:
                flush.zero_()
                a.record()
                model_gpu(inp_eval)
                b.record()
            barrier()
        ms_render = sum(a.elapsed_time(b) for a, b in ev)
        model_gpu.train(True)
    vals = reduce_max(weak["ms_res"], weak["ms_e2e"], ms_render, strong["ms_res"] if strong else 0.0, strong["ms_e2e"] if strong else 0.0)
    ms_res, ms_e2e, ms_render, ms_strong, ms_strong_e2e = vals
    gw, gs = weak.get("graph"), (strong.get("graph") if strong else None)
    gvals = reduce_max(gw["ms_res"] if gw else 0.0, gw["ms_e2e"] if gw else 0.0, gs["ms_res"] if gs else 0.0, gs["ms_e2e"] if gs else 0.0)
    if rank != 0:
        return None, None
    K = args.steps
    # where a block ran both ways (every rank did: the flags are the same everywhere) the faster one is its value, the other is kept beside it
    step_api = "eager"
    other = None
    if gw:
        e = {"api": "eager", "ms_per_step": ms_res / K, "e2e_ms_per_step": ms_e2e / K, "host_enqueue_ms_per_step": weak["host_ms"]}
        gr = {"api": "graph", "ms_per_step": gvals[0] / K, "e2e_ms_per_step": gvals[1] / K, "host_enqueue_ms_per_step": gw["host_ms"]}
        if gvals[0] < ms_res:
            step_api, other = "graph", e
            ms_res, ms_e2e = gvals[0], gvals[1]
            weak = dict(weak, host_ms=gw["host_ms"], rounds=gw["rounds"] or weak["rounds"])
        else:
            other = gr
    strong_api, strong_other = "eager", None
    if gs:
        e = {"api": "eager", "ms_per_step": ms_strong / K, "e2e_ms_per_step": ms_strong_e2e / K, "host_enqueue_ms_per_step": strong["host_ms"]}
        gr = {"api": "graph", "ms_per_step": gvals[2] / K, "e2e_ms_per_step": gvals[3] / K, "host_enqueue_ms_per_step": gs["host_ms"]}
        if gvals[2] < ms_strong:
            strong_api, strong_other = "graph", e
            ms_strong, ms_strong_e2e = gvals[2], gvals[3]
            strong = dict(strong, host_ms=gs["host_ms"])
        else:
            strong_other = gr
    prof = weak["prof"]
    value = world * R * N_COMPOSITED * K / (ms_res * 1e-3)
    e2e_value = world * R * N_COMPOSITED * K / (ms_e2e * 1e-3)
    pk = peaks()
    rounds_seen = sorted(set(weak["rounds"]))
    rounds_total = sum(weak["rounds"])
    # dominant kernel: sampler SDF evaluations (44 % of a training step's FLOPs, 71 % of a render's)
    sdf = prof["sampler_sdf"]
    # launches of rounds the device-side convergence word switched off return at once: only ACTIVE launches count (training: the
    # rounds each timed step really ran, read back by sampler_resolve; eval: the W-sharp weights run all of them)
    sdf_launches = max(rounds_total if (train and rounds_total > 0) else sdf["launches"], 1)
    pts_per_launch = R * 128
    per_launch_ms = sdf["ms"] / sdf_launches
    achieved = pts_per_launch * fm["sdf_eval"] / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"] if core.uses_tensor_cores else 72.0
    launches = sum(v["launches"] for v in prof.values())
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_tc_sdf_traffic.json")
    if core.uses_tensor_cores and os.path.exists(tpath) and R == 1024 and not light:
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]      # from the committed ncu --set full capture
    if train:
        workload = ("C2: training step on a 1024-ray batch, config/synthetic.yml networks and loss weights "
                    "(rgb L1 + eikonal + depth + normal/angular; steps < 50k: no bubble/smooth terms): forward "
                    "(error-bounded sampler 5x128 sdf-evals/ray, main pass on 97 samples/ray with saved activations, 3R eikonal "
                    "points) + I2SDFLoss + backward incl. second-order terms (fused tensor-core chain + one weight-gradient launch) "
                    "+ Adam(eps=1e-15) step (one launch) + weight re-pack"
                    + (" + one flat NCCL gradient all-reduce" if world > 1 else ""))
        if args.bubble > 0:
            workload = workload.replace("steps < 50k: no bubble/smooth terms", f"steps 50k-150k: + bubble loss on {args.bubble} surface points + smoothness term")
    else:
        workload = ("C2/C3 batch shape: 1024-ray forward render, config/synthetic.yml networks (8x256 SDF + 4x256 radiance), "
                    "eval layout: sampler 5 rounds = 640 sdf-evals/ray + 97 composited samples/ray")
    if light:
        workload = workload.replace("C2: training step", "C4: training step").replace("C2/C3 batch shape", "C4 batch shape") \
            .replace("config/synthetic.yml networks (8x256 SDF + 4x256 radiance)", "config/synthetic_light_mask.yml networks (6x256 SDF + 3x256 radiance + light-mask head)") \
            .replace("config/synthetic.yml networks and loss weights", "config/synthetic_light_mask.yml networks (6x256 SDF, 3x256 radiance, light-mask head) and loss weights (+ light-mask BCE 0.5)")
    mean_rounds = (rounds_total / len(weak["rounds"])) if (train and weak["rounds"]) else 5.0
    fstep = step_flops(fm, R, mean_rounds, train)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 hi/lo split products on tcgen05, fp32 accumulate in TMEM; backward chain: bf16 hi/lo)" if core.uses_tensor_cores else "f32",
        "data": "synthetic",
        "config": {"workload": workload, "mode": args.mode, "network_config": args.config,
                   "weights": ("W-sharp at step 0: reference geometric init (seed 0), density.beta=0.01 (all 5 sampler rounds run); training then "
                               "moves the weights every step (Adam, lr 5e-4) and the packed copies follow, so later steps may converge in fewer "
                               "rounds: see sampler_rounds_in_timed_steps") if train else
                              "W-sharp: reference geometric init (seed 0), density.beta=0.01 so all 5 sampler rounds run",
                   "rays_per_gpu": R, "global_rays": world * R, "samples_per_ray_composited": N_COMPOSITED,
                   "sdf_evals_per_ray": 5 * 128 + N_COMPOSITED,
                   "sampler_rounds_in_timed_steps": rounds_seen if train else "5 (eval: weights fixed)",
                   "parallelism": f"ray-sharded x{world}, " + ("one flat gradient all-reduce per step (persistent bucket, in place)" if train else "no collective (inference)"),
                   "tensor_cores": {"sampler_sdf": core.uses_tensor_cores, "main_pass": core.uses_tensor_cores_main,
                                    "backward": bool(train and core.fused_main)},
                   "l2": "flushed between timed iterations (256 MB write)", "timing": "CUDA events per step, max over ranks",
                   "step_api": ("i2sdf_b200.graph.GraphedTrainStep: the whole step captured once as a CUDA graph and replayed (one launch per step); "
                                "`kernel_ms_per_step` is taken from the eager run (identical kernels)") if step_api == "graph" else
                               ("I2SDFNetwork.forward (+ I2SDFLoss, backward, GradBucket.allreduce, optim.Adam.step) issued from Python" if train else "I2SDFNetwork.forward (eval)")},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": weak["h2d"], "d2h_bytes_per_step": weak["d2h"]},
        "gpu_launches": launches,
        "kernel_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
        "kernel_launches_per_step": {k: v["launches"] / K for k, v in prof.items()},
        "roofline": {"bound": "tensor", "kernel": "tc_sdf8_kernel (sampler SDF evaluations)" if core.uses_tensor_cores else "mlp_tile_kernel",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per launch (ncu --set full, profiles/r02_tc_sdf_traffic.json)",
                     "flop_per_launch": pts_per_launch * fm["sdf_eval"], "ms_per_launch": per_launch_ms,
                     "peak_source": pk["source"] + (", sustained bf16 (kernel timed inside a step)" if core.uses_tensor_cores else "; fp32 FMA peak 148 SM x 128 FMA x 2 x 1.9 GHz"),
                     "note": "achieved counts ALGORITHMIC flops (1 MAC = 2 flop); the kernel issues 3 fp16 MMAs per MAC (hi*hi + lo*hi + hi*lo), so 1/3 of the 16-bit "
                             "dense peak is this precision scheme's ceiling"},
        # whole-step roofline: the step's algorithmic FLOPs by the formula in this file's doc string / flop_model()
        "roofline_step": {"bound": "tensor", "flop_per_step_per_gpu": fstep, "achieved": fstep / (ms_res / K * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                          "frac": fstep / (ms_res / K * 1e-3) / 1e12 / peak if peak else None, "flop_model": fm, "mean_sampler_rounds": mean_rounds,
                          "formula": ("rounds*128*R*sdf_eval + 97*R*(ray_sample + ray_sample_bwd) + 3*R*(eik_fwd + eik_bwd)" if train
                                      else "rounds*128*R*sdf_eval + 97*R*ray_sample")},
        "clocks": clocks,
        "host_enqueue_ms_per_step": weak["host_ms"],
    }
    if other:
        line["other_step_api"] = other
    if weak["ms_ar"] is not None:
        line["allreduce"] = {"bytes": int(bucket.flat.numel() * 4), "median_ms_alone": weak["ms_ar"],
                             "note": "one NCCL AVG all-reduce of the persistent flat gradient bucket, timed back to back outside the step; inside the step the ranks "
                                     "also wait for the slowest one (the timed value is the max over ranks of every step)"}
    if strong is not None:
        sv = R * N_COMPOSITED * K / (ms_strong * 1e-3)
        line["strong_scaling"] = {
            "value": sv, "unit": UNIT, "ms_per_step": ms_strong / K, "global_rays": R, "rays_per_gpu": R / world,
            "e2e": {"value": R * N_COMPOSITED * K / (ms_strong_e2e * 1e-3), "ms_per_step": ms_strong_e2e / K},
            "workload": "SURVEY.md C5: ONE 1024-ray batch split over the ranks (same weights / rays / targets as N = 1), plain sharding "
                        "(per-shard sampler convergence and loss means; the strict-parity switches add 5 + 1 tiny all-reduces per step)",
            "kernel_ms_per_step": {k: v["ms"] / K for k, v in strong["prof"].items()},
            "host_enqueue_ms_per_step": strong["host_ms"],
            "step_api": strong_api, "other_step_api": strong_other,
            "limiter": "issued from Python (`eager`) the step is host-bound: the kernels of a 1024 / N-ray shard need sum(kernel_ms_per_step) of device "
                       "time, Python needs ~2.7 ms to issue the step's ~60 launches; replayed as a CUDA graph the host is out of the step and what remains is "
                       "launch-to-launch latency of ~60 dependent kernels that each fill at most one partial wave (128 rays x 128 points = 128 tiles on 148 "
                       "SMs at N = 8) plus the all-reduce",
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
        fr = step_flops(fm, R, 5.0, False)
        line["render"] = {"value": world * R * N_COMPOSITED * K / (ms_render * 1e-3), "unit": UNIT, "ms_per_step": ms_render / K,
                          "workload": "eval forward render of the same 1024-ray batch (no backward, no collective), device-resident inputs",
                          "roofline_step": {"flop_per_step_per_gpu": fr, "achieved": fr / (ms_render / K * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                            "frac": fr / (ms_render / K * 1e-3) / 1e12 / peak if peak else None}}
    return line, (conf, cpu_snapshot)


def grid_arm(args, local_rank):
    """SDF of a res^3 uniform grid on [-2, 2]^3 through i2sdf_sdf_grid (points generated on the device, sdf-only chain)."""
    from i2sdf_b200.grid import grid_axes_uniform, grid_points, sdf_grid
    from oracle import i2sdf_oracle as orc
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    conf, model = build_params(name=args.config)
    fm = flop_model(conf)
    m = model.to(dev).eval()
    x, y, z = grid_axes_uniform(args.grid_res)
    n = len(x) * len(y) * len(z)
    for _ in range(max(args.warmup, 3)):
        sdf_grid(m, x, y, z)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    clk = ClockSampler(local_rank)
    clk.start()
    for a, b in ev:
        flush.zero_()
        a.record()
        out = sdf_grid(m, x, y, z)
        b.record()
    torch.cuda.synchronize()
    clocks = clk.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    # end to end: axes from the host, the volume back to the host (what marching cubes consumes)
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host.copy_(sdf_grid(m, x, y, z), non_blocking=True)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    pk = peaks()
    # CPU baseline: the reference's loop over a bounded sample of the grid (first 2^18 points), implicit_network(p)[:, 0] incl. the 256 discarded features
    spec = orc.spec_from_model_conf(conf, use_normal=False)
    P = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    layers = orc.layer_params(P, "implicit_network", spec.n_sdf_layers)
    pts = grid_points(x[:64], y[:64], z[:64])
    with torch.no_grad():
        pick_cpu_threads(lambda: orc.sdf_mlp(spec, layers, pts[:4096]))
        t0 = time.perf_counter()
        ref = orc.sdf_mlp(spec, layers, pts)[0][:, 0]
        dt = time.perf_counter() - t0
    chk = sdf_grid(m, x[:64], y[:64], z[:64]).cpu()
    err = float((chk - ref).abs().max() / ref.abs().max())
    ach = n * fm["sdf_eval"] / (ms * 1e-3) / 1e12
    return {"metric": "SDF grid evaluation for mesh extraction (secondary line; the BASELINE metric is --mode train)", "value": n / (ms * 1e-3), "unit": "points/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 hi/lo split products on tcgen05, fp32 accumulate in TMEM)", "data": "synthetic",
            "config": {"workload": f"sdf of a {args.grid_res}^3 uniform grid on [-2, 2]^3 (utils/plots.py:440-451 point order), points generated on the device, "
                                   f"{args.config}.yml SDF network, W-sharp weights", "grid_points": n, "l2": "flushed between timed iterations"},
            "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 3 * 4 * args.grid_res, "d2h_bytes_per_step": 4 * n},
            "gpu_launches": args.steps,
            "roofline": {"bound": "tensor", "kernel": "tc_sdf8_kernel (grid point source)", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "flop_per_point": fm["sdf_eval"]},
            "clocks": clocks, "parity": {"max_abs_err_over_max_abs_sdf_vs_oracle_on_64^3": err},
            "cpu_baseline": {"value": pts.shape[0] / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"first 64^3 = {pts.shape[0]} grid points through the oracle's ImplicitNetwork.forward (257 outputs, 256 discarded, as "
                                       "model/eval/recon.py:51 does), torch CPU fp32"}}


def add_cpu_baseline(line, args, conf, cpu_snapshot):
    """Rank 0, after the process group is gone (so the other ranks are not spinning in a barrier next to it)."""
    model_c = type("S", (), {"state_dict": lambda self: cpu_snapshot})()
    rays = args.cpu_rays or args.rays
    if args.mode == "train":
        cb = cpu_train_arm(conf, model_c, rays, steps=2, warmup=1, name=args.config, full_rays=args.rays)
    else:
        cb = cpu_arm(conf, model_c, rays, steps=2, warmup=1, full_rays=args.rays)
    line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                            "same_config": cb["same_config"], "ms_per_step": cb["ms_per_step"]}
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
        rays = args.cpu_rays or args.rays
        if args.mode == "train":
            cb = cpu_train_arm(conf, model, rays, steps=args.steps, warmup=args.warmup, name=args.config, full_rays=args.rays)
        else:
            cb = cpu_arm(conf, model, rays, steps=args.steps, warmup=args.warmup, full_rays=args.rays)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"same as --impl ours --mode {args.mode}: {rays}-ray batch, {args.config}.yml networks, W-sharp weights, same rays / targets"
                                       + (", training step incl. the Adam update" if args.mode == "train" else "") + ", on the host CPU (oracle port of the "
                                       "reference's PyTorch path; one process: the reference has no multi-device path, so the CPU arm does not scale with --gpus)",
                           "mode": args.mode, "same_config": cb["same_config"]},
                "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"], "same_config": cb["same_config"]},
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
    if args.mode == "grid":
        if rank == 0:
            print(json.dumps(grid_arm(args, local_rank)), flush=True)
        return
    line, cpu = gpu_arm(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        # the CPU baseline runs with the GPUs idle, at N = 1 only (the same host cores at every N: the multi-GPU lines of a scaling run
        # would just repeat it, ~25 s each)
        if world == 1:
            line = add_cpu_baseline(line, args, *cpu)
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
