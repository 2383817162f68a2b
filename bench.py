#!/usr/bin/env python
"""Benchmark of the GCA+TAM frame-window forward (BASELINE.json configs[1]): 3-frame 1088x1920
windows per second, one process per GPU, windows sharded across ranks (weak scaling, no data-path
collective).

  python bench.py --gpus N --steps K --warmup W             # native sm_100a path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores (unmodified
                                                            # reference from baseline/_ref, else the CPU oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

H, W, S = 1088, 1920, 3
METRIC = "1080p 3-frame windows/sec (GCA+TAM forward)"
UNIT = "windows/s"
GFLOP_PER_WINDOW = 3843.57          # BASELINE.md section 2 (FlopCounterMode on the reference)
FBA_GFLOP_PER_WINDOW = 7111.6       # EvalModel('vmn_fba') at 1088x1920 (223.1 GFLOP @256^2)


JSON_OUT = sys.stdout     # the ONE JSON line goes here; main() points sys.stdout at stderr so that library / reference
                          # chatter (e.g. the reference FBA's "modifying input layer ..." print) cannot pollute it


def bench_config(world):
    """The `config` object of the JSON line -- ONE definition for both arms, so that the driver's same_config check compares
    identical dicts (round-1 verdict: the two arms described the same workload with different strings)."""
    return dict(workload="GCA+TAM forward-only 1080p 3-frame window, batch 1 per GPU (configs[1])",
                frames=S, height=H, width=W, windows_per_step_per_gpu=1,
                weights="calibrated random-init fixture (tests/golden)",
                l2="working set per step (>8 GB of activations) exceeds the 126 MB L2; no flush needed",
                parallelism=f"dp{world} (independent windows per GPU, no collective)")


def window_gflop(h, w):
    """Reference FLOP count of one 3-frame window at h x w (convs scale with pixels,
    GCA attention with pixels^2; anchored on the 256x256 measurement 54.81 + 2.06 GFLOP)."""
    r = (h * w) / 65536.0
    return 54.81 * r + 2.06 * r * r


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 7:
                    continue
                try:
                    sm.append(float(c[0])); mx.append(float(c[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # under load = upper half of the samples (the sampler also sees idle edges)
            out["sm_mhz"] = statistics.median(sorted(sm)[len(sm) // 2:])
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------- CPU arm
_REF_CPU = {}


def cpu_kind():
    """"reference": the UNMODIFIED reference modules from baseline/_ref (baseline/install_ref.py) on the host cores;
    "port": the CPU oracle (oracle/vmn_gca_oracle.py) when that copy is not present."""
    from baseline import ref_env
    return "reference" if ref_env.available() else "port"


def _reference_eval_model_cpu(arch):
    """EvalModel(arch) of the unmodified reference on the CPU, fixture weights loaded (strict=True, pred_test.py:90-94)."""
    if arch not in _REF_CPU:
        import torch
        from baseline import ref_env
        from helpers import fixture_sd, fixture_sd_fba
        ref_env.activate()
        from models.model import EvalModel
        ref = getattr(EvalModel, "_reference", EvalModel)
        m = ref(model=arch, agg_window=7, dilate_kernel=None)
        m.NET.load_state_dict(fixture_sd() if arch == "vmn_gca" else fixture_sd_fba(), strict=True)
        _REF_CPU[arch] = m.eval()
    return _REF_CPU[arch]


def cpu_oracle_time(sample_hw, threads, reps=1):
    """Times the reference's CPU path on one 3-frame window: the unmodified reference modules when baseline/_ref is present
    (cpu_kind() == "reference"), else the CPU oracle port (torch fp32 restatement of the reference)."""
    import torch
    from helpers import fixture_sd
    from tcvom_b200 import synthetic
    torch.set_num_threads(threads)
    h, w = sample_hw
    imgs, tris = synthetic.make_window(h, w, seed=7)
    ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
    if cpu_kind() == "reference":
        m = _reference_eval_model_cpu("vmn_gca")
        run = lambda: m(ti, tt)
    else:
        from oracle import vmn_gca_oracle as O
        sd = fixture_sd()
        run = lambda: O.eval_forward(sd, ti, tt)
    from baseline import ref_env
    ts = []
    with torch.no_grad(), ref_env.on_cpu():
        for _ in range(reps):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
    return ts


def cpu_oracle_time_fba(sample_hw, threads):
    """Times the FBA+TAM CPU oracle (oracle/vmn_fba_oracle.py) on one 3-frame window (seconds)."""
    import torch
    from helpers import fixture_sd_fba
    from oracle import vmn_fba_oracle as OF
    from tcvom_b200 import synthetic
    torch.set_num_threads(threads)
    h, w = sample_hw
    imgs, tris = synthetic.make_window(h, w, seed=7)
    ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
    if cpu_kind() == "reference":
        m = _reference_eval_model_cpu("vmn_fba")
        from baseline import ref_env
        with torch.no_grad(), ref_env.on_cpu():
            t0 = time.perf_counter()
            m(ti, tt)
            return time.perf_counter() - t0
    sd = fixture_sd_fba()
    t0 = time.perf_counter()
    OF.eval_forward(sd, ti, tt)
    return time.perf_counter() - t0


def pick_cpu_sample(threads, budget_s, n_steps):
    """Largest window size whose estimated oracle time fits the budget.  The estimate scales a
    256x480 probe by the reference FLOP model; the returned value is always MEASURED on the
    chosen sample and converted to 1080p-window equivalents by the same FLOP model."""
    probe = min(cpu_oracle_time((256, 480), threads, reps=2))
    rate = window_gflop(256, 480) / probe                      # GFLOP/s on this host
    for hw in ((1088, 1920), (544, 960), (256, 480)):
        if n_steps * window_gflop(*hw) / rate <= budget_s:
            return hw
    return (256, 480)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    hw = pick_cpu_sample(threads, 240.0, args.steps + args.warmup)
    frac = window_gflop(*hw) / window_gflop(H, W)
    for _ in range(args.warmup):
        cpu_oracle_time(hw, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_time(hw, threads)
    dt = time.perf_counter() - t0
    value = args.steps * frac / dt
    kind = cpu_kind()
    world = args.gpus
    sample = (f"one 3-frame {hw[0]}x{hw[1]} window per step = {frac:.4f} of a 1088x1920 window by the "
              f"reference FLOP model (BASELINE.md section 2); " +
              ("unmodified reference modules (baseline/_ref) on the CPU" if kind == "reference" else "CPU oracle port"))
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=bench_config(world),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------- training step (configs[2] / [3])
TRAIN_B, TRAIN_S, TRAIN_HW = 4, 5, 512
TRAIN_GFLOP_PER_SAMPLE = 1310.95        # 5-frame sample, fwd+bwd, 512x512 (reference FlopCounterMode)
LOSS_WEIGHTS = (1.0, 1.0, 1.0, 0.5, 0.25)   # train_ddp.py:61


TRAIN_1080_GFLOP_PER_SAMPLE = 20000.0   # 1088x1920, S=5, fwd 6 661 GFLOP, fwd+bwd ~3x


def run_train_section(args, rank, world, dev, barrier, max_over_ranks, shape=None, freeze=False):
    """Secondary measurement: the native training step (FullModel_VMD fwd + losses + bwd + Adam, train_ddp.py:52-65)
    at BASELINE.json configs[2]'s shape, batch 4 per GPU; with N > 1 under SyncBatchNorm + DistributedDataParallel
    over NCCL exactly like train_ddp.py:270-280 (configs[3]'s recipe).  Reported next to, not instead of, the
    headline forward metric."""
    import torch
    import tcvom_b200
    from tcvom_b200 import _cabi, synthetic
    from helpers import fixture_sd
    # shape = (batch per GPU, frames, H, W, timed steps, reference GFLOP per sample-step, label)
    TRAIN_B, TRAIN_S, TH_, TW_, steps, gflop, label = shape or (4, 5, 512, 512, 5, TRAIN_GFLOP_PER_SAMPLE, "configs[2]")
    torch.cuda.reset_peak_memory_stats(dev)
    # TCV_FREEZE_BACKBONE=1 (tools/train_time.py only): the TAM pre-training mode, train_single_ddp.py:184-185
    freeze = freeze or os.environ.get("TCV_FREEZE_BACKBONE") == "1"
    model = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=None,
                                     **(dict(freeze_backbone=True) if freeze else {}))
    model.NET.load_state_dict(fixture_sd(), strict=True)
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).to(dev)
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], output_device=dev.index,
                                                          find_unused_parameters=True)
    else:
        model = model.to(dev)
    model.train()
    a, fg, bg = (torch.from_numpy(t).float().to(dev)
                 for t in synthetic.make_train_batch(TRAIN_B, TRAIN_S, TH_, TW_, seed=21 + 100 * rank))
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-4)
    torch.manual_seed(1234 + rank)

    def step():
        out = model(a, fg, bg)
        loss = sum(w * o.mean() for w, o in zip(LOSS_WEIGHTS, out[:5]))
        model.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    # The step issues ~1 200 kernel launches from Python and is host-bound at 512x512 (45 ms of issue time for 47 ms of
    # kernels): a full collection of the cyclic garbage collector inside the timed steps (~50 ms over the process's 280 k
    # long-lived objects; measured +11 ms per step averaged over 5 steps whenever one fell into them) is kept out the way
    # training scripts do it -- collect once, then freeze the survivors (gc.freeze: they are never scanned again).
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    n0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier()
    gc.unfreeze()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    samples = world * TRAIN_B
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    what = ("GCA+TAM pre-training step, freeze_backbone=True (train_single_ddp.py:184-185: frozen encoder + decoder head "
            "in eval mode, TAM + decoder tail fwd+bwd+Adam)" if freeze else
            "GCA+TAM train step (L_im+L_tc+L_af fwd+bwd+Adam)")
    out = dict(workload=f"{what} {TH_}x{TW_}, batch {TRAIN_B}/GPU, "
                         f"S={TRAIN_S} ({label}; N>1: SyncBatchNorm + DDP over NCCL, train_ddp.py:270-280)",
                ms_per_step=ms, samples_per_s=samples / (ms / 1e3),
                centre_windows_per_s=samples * (TRAIN_S - 2) / (ms / 1e3),
                algorithmic_tflops=(samples * gflop / (ms / 1e3) / 1e3) if gflop else None, steps=steps, warmup=2,
                gpu_launches_per_step=(_cabi.launch_count() - n0) // steps, loss=float(loss.detach()),
                grad_allreduce_mb=(4 * n_params / 1e6) if world > 1 else 0.0, sync_batchnorm=world > 1,
                peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    # release the step's arena before the next section
    net = model.module.NET if world > 1 else model.NET
    getattr(net, "_train_engines", {}).clear()
    del model, opt, a, fg, bg, loss
    torch.cuda.empty_cache()
    return out


def run_fba_section(args, rank, world, dev, max_over_ranks):
    """Secondary measurement: EvalModel('vmn_fba') forward on one 1088x1920 3-frame window per GPU (BASELINE.json
    configs[4], the second base network behind the same plugin seam).  Inputs resident in HBM, CUDA-graph replay;
    reported next to, not instead of, the headline metric.  A failure on a rank is reported, not raised (the
    collective below must still be entered by every rank)."""
    import torch
    import tcvom_b200
    from tcvom_b200 import synthetic
    from helpers import fixture_sd_fba
    ms_local, info, err = -1.0, {}, None
    try:
        model = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7, dilate_kernel=None)
        model.NET.load_state_dict(fixture_sd_fba(), strict=True)
        model = model.to(dev).eval()
        imgs_np, tris_np = synthetic.make_window(H, W, seed=7 + rank)
        imgs, tris = torch.from_numpy(imgs_np).to(dev), torch.from_numpy(tris_np).to(dev)
        steps = 5
        with torch.no_grad():
            out = model(imgs, tris)
            plan = list(model.NET.engine().plans.values())[0]
            for _ in range(3):
                model.run_plan(plan)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model.run_plan(plan)
            e1.record()
            torch.cuda.synchronize(dev)
            ms_local = e0.elapsed_time(e1) / steps
            kinds = {}
            if rank == 0:
                plan.replay_timed(dev)
                for m, t in zip(plan.meta, plan.replay_timed(dev)):
                    kinds[m["kind"]] = kinds.get(m["kind"], 0.0) + t
            info = dict(gpu_launches_per_window=plan.n_launch, steps=steps, warmup=3,
                        peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                        finite=bool(torch.isfinite(out[0]).all()),
                        breakdown_ms={k: round(v, 3) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])})
        model.NET.engine().plans.clear()
        del plan, model
        import gc
        from tcvom_b200.engine import release_idle_pools
        gc.collect()
        release_idle_pools()
        torch.cuda.empty_cache()
    except Exception as e:                                     # noqa: BLE001 - reported in the JSON line
        err = f"{type(e).__name__}: {e}"
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and err is None:
        # the CPU oracle of this config on the box's host cores, one bounded sample (a half-resolution window costs
        # ~2 s on 16 cores; the conv-only FLOP model scales it to a 1088x1920 window)
        try:
            threads = os.cpu_count() or 1
            hw = (544, 960)
            t = cpu_oracle_time_fba(hw, threads)
            frac = (hw[0] * hw[1]) / float(H * W)
            cpu = dict(value=frac / t, unit=UNIT, cores=threads, kind=cpu_kind(),
                       sample=f"one 3-frame {hw[0]}x{hw[1]} FBA+TAM window ({t:.2f} s) = {frac:.4f} of a 1088x1920 window "
                              f"(convolution FLOPs scale with the pixel count)")
        except Exception as e:                                 # noqa: BLE001
            cpu = dict(error=f"{type(e).__name__}: {e}")
    ms = max_over_ranks(ms_local)
    any_failed = max_over_ranks(1.0 if (err is not None or ms_local < 0) else 0.0) > 0
    if any_failed:
        return dict(workload="FBA+TAM forward 1080p (configs[4])", error=err or "failed on another rank")
    return dict(workload="FBA+TAM forward-only 1080p 3-frame window, batch 1 per GPU (configs[4]; reference FLOP count "
                         f"{FBA_GFLOP_PER_WINDOW:.1f} GFLOP/window, convolutions only)",
                ms_per_window=ms, windows_per_s=world * 1e3 / ms,
                algorithmic_tflops=FBA_GFLOP_PER_WINDOW / ms, cpu_baseline=cpu, **info)


DIM_GFLOP_PER_WINDOW = 7103.1       # convolution FLOPs of EvalModel('vmn_dim') at 1088x1920: 3 x 2028.9 (per frame) + 1016.3 (tail)


INDEX_GFLOP_PER_WINDOW = 874.0      # torch FlopCounterMode over the oracle's EvalModel('vmn_index') forward, scaled to 1088x1920


def run_dim_section(args, rank, world, dev, max_over_ranks, arch="vmn_dim"):
    """Secondary measurement: EvalModel('vmn_dim' | 'vmn_index') forward on one 1088x1920 3-frame window per GPU (SURVEY.md
    section 8 row f4, the third and fourth base network behind the plugin seam).  Inputs resident in HBM, CUDA-graph replay."""
    import gc
    import torch
    import tcvom_b200
    from tcvom_b200 import synthetic
    from tcvom_b200.engine import release_idle_pools
    from helpers import fixture_sd_dim, fixture_sd_index
    label, gflop, fixture = (("DIM+TAM", DIM_GFLOP_PER_WINDOW, fixture_sd_dim) if arch == "vmn_dim" else
                             ("IndexNet+TAM", INDEX_GFLOP_PER_WINDOW, fixture_sd_index))
    ms_local, info, err = -1.0, {}, None
    try:
        torch.cuda.reset_peak_memory_stats(dev)
        model = tcvom_b200.EvalModel(model=arch, agg_window=7, dilate_kernel=None)
        model.NET.load_state_dict(fixture(), strict=True)
        model = model.to(dev).eval()
        imgs_np, tris_np = synthetic.make_window(H, W, seed=7 + rank)
        imgs, tris = torch.from_numpy(imgs_np).to(dev), torch.from_numpy(tris_np).to(dev)
        steps = 5
        with torch.no_grad():
            out = model(imgs, tris)
            plan = list(model.NET.engine().plans.values())[0]
            for _ in range(3):
                model.run_plan(plan)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model.run_plan(plan)
            e1.record()
            torch.cuda.synchronize(dev)
            ms_local = e0.elapsed_time(e1) / steps
            kinds = {}
            if rank == 0:
                plan.replay_timed(dev)
                for m, t in zip(plan.meta, plan.replay_timed(dev)):
                    kinds[m["kind"]] = kinds.get(m["kind"], 0.0) + t
            info = dict(gpu_launches_per_window=plan.n_launch, steps=steps, warmup=3,
                        peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                        finite=bool(torch.isfinite(out).all()),
                        breakdown_ms={k: round(v, 3) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])})
        model.NET.engine().plans.clear()
        del plan, model
        gc.collect()
        release_idle_pools()
        torch.cuda.empty_cache()
    except Exception as e:                                     # noqa: BLE001 - reported in the JSON line
        err = f"{type(e).__name__}: {e}"
    ms = max_over_ranks(ms_local)
    any_failed = max_over_ranks(1.0 if (err is not None or ms_local < 0) else 0.0) > 0
    if any_failed:
        return dict(workload=f"{label} forward 1080p", error=err or "failed on another rank")
    return dict(workload=f"{label} forward-only 1080p 3-frame window, batch 1 per GPU ({arch}, SURVEY 8 f4; "
                         f"{gflop:.1f} GFLOP/window, convolutions only)",
                ms_per_window=ms, windows_per_s=world * 1e3 / ms, algorithmic_tflops=gflop / ms, **info)


# ------------------------------------------------------------------------------------- reference on the same GPU
def run_gpu_eager_section(dev):
    """What the hot path costs TODAY on this GPU without this repo: the reference's PyTorch modules (eager, cuDNN / cuBLAS,
    NCHW fp32) on cuda:0 -- the unmodified reference from baseline/_ref when present ("reference"), else the oracle port
    ("port").  configs[1] forward, configs[2] train step, configs[4] FBA forward; each with TF32 off (fp32-accurate, what
    the 1e-3 contract is stated against) and with TF32 on + cudnn.benchmark (train_ddp.py:189-191 / torch defaults).
    A reported baseline like cpu_baseline: never on the measured path."""
    import torch
    from helpers import fixture_sd, fixture_sd_fba
    from tcvom_b200 import synthetic
    from baseline import ref_env
    kind = "reference" if ref_env.available() else "port"
    out = dict(kind=kind, device=torch.cuda.get_device_name(dev))

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    if kind == "reference":
        ref_env.activate()
        import models.model as RM
        Eval = getattr(RM.EvalModel, "_reference", RM.EvalModel)
        VMD = getattr(RM.FullModel_VMD, "_reference", RM.FullModel_VMD)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for tag, tf32 in (("fp32", False), ("tf32_cudnn_benchmark", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = tf32
            res = {}
            # ---- configs[1]: GCA+TAM forward, one 1088x1920 window
            try:
                imgs, tris = (torch.from_numpy(t).float().to(dev) for t in synthetic.make_window(H, W, seed=7))
                if kind == "reference":
                    m = Eval(model="vmn_gca", agg_window=7, dilate_kernel=None)
                    m.NET.load_state_dict(fixture_sd(), strict=True)
                    m = m.to(dev).eval()
                    fn = lambda: m(imgs, tris)
                else:
                    from oracle import vmn_gca_oracle as O
                    sd = {k: v.to(dev) for k, v in fixture_sd().items()}
                    fn = lambda: O.eval_forward(sd, imgs, tris)
                with torch.no_grad():
                    ms = timed(fn, 2, 5)
                res["forward_1080p"] = dict(ms_per_window=ms, windows_per_s=1e3 / ms)
                del fn
            except Exception as e:                             # noqa: BLE001
                res["forward_1080p"] = dict(error=f"{type(e).__name__}: {e}")
            torch.cuda.empty_cache()
            # ---- configs[2]: train step 512x512, batch 4, S=5 (fwd + losses + bwd + Adam)
            try:
                a, fg, bg = (torch.from_numpy(t).float().to(dev)
                             for t in synthetic.make_train_batch(TRAIN_B, TRAIN_S, TRAIN_HW, TRAIN_HW, seed=21))
                torch.manual_seed(1234)
                if kind == "reference":
                    tm = VMD(model="vmn_gca", agg_window=7)
                    tm.NET.load_state_dict(fixture_sd(), strict=True)
                    tm = tm.to(dev).train()
                    opt = torch.optim.Adam([p for p in tm.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-4)

                    def fn():
                        o = tm(a, fg, bg)
                        loss = sum(w * x.mean() for w, x in zip(LOSS_WEIGHTS, o[:5]))
                        tm.zero_grad()
                        loss.backward()
                        opt.step()
                else:
                    from oracle import vmn_gca_oracle as O
                    sd = {k: v.to(dev).requires_grad_(v.dtype.is_floating_point and not k.endswith(
                        ("_u", "_v", "running_mean", "running_var"))) for k, v in fixture_sd().items()}
                    params = [v for v in sd.values() if v.requires_grad]
                    opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-4)

                    def fn():
                        o = O.full_vmd_forward(sd, a, fg, bg, [3] * TRAIN_B, train=True)
                        loss = sum(w * x.mean() for w, x in zip(LOSS_WEIGHTS, o[:5]))
                        opt.zero_grad()
                        loss.backward()
                        opt.step()
                ms = timed(fn, 1, 2)
                res["train_512"] = dict(ms_per_step=ms, samples_per_s=TRAIN_B * 1e3 / ms)
                del fn, opt
            except Exception as e:                             # noqa: BLE001
                res["train_512"] = dict(error=f"{type(e).__name__}: {e}")
            torch.cuda.empty_cache()
            # ---- configs[4]: FBA+TAM forward, one 1088x1920 window (incl. the reference's host cv2 distance transforms)
            try:
                imgs, tris = (torch.from_numpy(t).float().to(dev) for t in synthetic.make_window(H, W, seed=7))
                if kind == "reference":
                    m = Eval(model="vmn_fba", agg_window=7, dilate_kernel=None)
                    m.NET.load_state_dict(fixture_sd_fba(), strict=True)
                    m = m.to(dev).eval()
                    fn = lambda: m(imgs, tris)
                else:
                    from oracle import vmn_fba_oracle as OF
                    sd = {k: v.to(dev) for k, v in fixture_sd_fba().items()}
                    fn = lambda: OF.eval_forward(sd, imgs, tris)
                with torch.no_grad():
                    ms = timed(fn, 1, 3)
                res["fba_forward_1080p"] = dict(ms_per_window=ms, windows_per_s=1e3 / ms)
                del fn
            except Exception as e:                             # noqa: BLE001
                res["fba_forward_1080p"] = dict(error=f"{type(e).__name__}: {e}")
            torch.cuda.empty_cache()
            out[tag] = res
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


# ------------------------------------------------------------------------------------- native arm
def run_native(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import tcvom_b200
    from tcvom_b200 import _cabi, synthetic
    from helpers import fixture_sd

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None)
    model.NET.load_state_dict(fixture_sd(), strict=True)
    model = model.to(dev).eval()

    imgs_np, tris_np = synthetic.make_window(H, W, seed=7 + rank)
    imgs_h = torch.from_numpy(imgs_np).float().pin_memory()
    tris_h = torch.from_numpy(tris_np).float().pin_memory()
    # end-to-end arm: frames and trimaps cross PCIe as uint8 (what cv2.imread yields, pred_test.py:74-76); the
    # reference API accepts them (EvalModel.preprocess casts with .float(), models/model.py:366-368)
    imgs_u8 = torch.from_numpy(imgs_np).pin_memory()
    tris_u8 = torch.from_numpy(tris_np).pin_memory()
    # pred_test.py:107-110 reads back the CENTRE frame's matte only (`model(imgs, tris).squeeze()[c]...cpu()`); the first /
    # last frame of the returned tensor are zeros by contract (models/model.py:419-421)
    out_h = torch.empty((1, 1, 1, H, W), dtype=torch.float32).pin_memory()

    from tcvom_b200 import dp

    def barrier():
        dp.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        return dp.max_over_ranks(ms, dev)

    with torch.no_grad():
        # ---- record the plan, fill the resident input buffers
        model(imgs_h.to(dev), tris_h.to(dev))
        plan = list(model.NET.engine().plans.values())[0]
        plan.io["imgs"].copy_(imgs_h); plan.io["tris"].copy_(tris_h)
        torch.cuda.synchronize(dev)

        # ---- value: whole hot path, inputs resident in HBM
        for _ in range(max(args.warmup, 3)):
            model.run_plan(plan)
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _cabi.launch_count()
        e0.record()
        for _ in range(args.steps):
            model.run_plan(plan)
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = plan.n_launch * args.steps
        value = world * args.steps / (ms / 1e3)

        # ---- e2e: the user-facing call with HOST buffers (pred_test.py:100-107 sequence)
        # The loop a clip-processing user writes (DataLoader(pin_memory=True) + non_blocking copies): the NEXT window's
        # uint8 frames upload on a side stream while the current window computes.  Every step's upload, its forward and
        # the read-back of its matte are inside the timed region (step 0's upload is ordered after e0).
        main = torch.cuda.current_stream(dev)
        up = torch.cuda.Stream(dev)
        down = torch.cuda.Stream(dev)

        def upload():
            with torch.cuda.stream(up):
                bufs = (imgs_u8.to(dev, non_blocking=True), tris_u8.to(dev, non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(up)
            return bufs, ev

        def e2e_loop(n):
            up.wait_stream(main)                     # step 0's upload starts after everything recorded so far (e0)
            nxt = upload()
            for i in range(n):
                (im_d, tr_d), ev = nxt
                main.wait_event(ev)
                im_d.record_stream(main); tr_d.record_stream(main)
                a = model(im_d, tr_d)                # a fresh tensor per call (the plan's output is cloned)
                if i + 1 < n:
                    nxt = upload()
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(down):        # matte read-back behind the next window's compute
                    down.wait_event(done)
                    out_h.copy_(a[:, S // 2:S // 2 + 1], non_blocking=True)
                    a.record_stream(down)
            main.wait_stream(down)                   # the last read-back ends before e1

        e2e_loop(3)
        barrier()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_value = world * args.steps / (ms_e2e / 1e3)

        # ---- roofline of the dominant kernel: CUDA events around every launch of an eager replay
        roof = None
        if rank == 0:
            plan.replay_timed(dev)
            reps = 3
            acc = [0.0] * len(plan.calls)
            for _ in range(reps):
                for i, t in enumerate(plan.replay_timed(dev)):
                    acc[i] += t / reps
            kinds = {}
            for m, t in zip(plan.meta, acc):
                k = kinds.setdefault(m["kind"], dict(ms=0.0, n=0, flops=0, bytes=0))
                k["ms"] += t; k["n"] += 1; k["flops"] += m["flops"]; k["bytes"] += m["bytes"]
            total = sum(k["ms"] for k in kinds.values())
            top_name, top = max(kinds.items(), key=lambda kv: kv[1]["ms"])
            pk = peaks()
            tensor_kinds = ("gca_scores_gemm", "gca_pv_gemm", "conv_tc", "conv_tc2", "conv_tc2p", "conv_tc3", "gca_scores_gemm_tc", "gca_pv_gemm_tc")
            if top_name in tensor_kinds:
                ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
                roof = dict(bound="tensor", achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s",
                            frac=ach / pk["tf_sustained"], traffic=None,
                            # every MMA of these kernels is issued three times (bf16x3 split):
                            # the tensor pipe does 3x the algorithmic FLOPs, so frac is capped at 1/3
                            mma_issue_multiplier=3, frac_incl_split=3 * ach / pk["tf_sustained"])
            else:
                ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
                roof = dict(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], traffic=None)
            roof.update(kernel=top_name, launches_per_step=top["n"], avg_launch_ms=top["ms"] / top["n"],
                        share_of_step=top["ms"] / total, peak_source=pk["source"],
                        breakdown_ms={k: round(v["ms"], 3) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"])})
            if args.dump_calls:
                with open(args.dump_calls, "w") as f:
                    for i, (m, t) in sorted(enumerate(zip(plan.meta, acc)), key=lambda imt: -imt[1][1]):
                        f.write(json.dumps(dict(idx=i, ms=round(t, 4), **{k: v for k, v in m.items()})) + "\n")
            whole = GFLOP_PER_WINDOW / (ms / args.steps * 1e-3) / 1e3
            roof["whole_step_tflops"] = whole
            roof["whole_step_frac_of_tensor_peak"] = whole / pk["tf_sustained"]
            tr = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tr):
                t = json.load(open(tr)).get(top_name)
                if t:
                    roof["traffic"] = t["dram_bytes_per_launch"]       # ncu dram__bytes_read+write, per launch
                    roof["algorithmic_bytes_per_launch"] = top["bytes"] / top["n"]

    train = train_1080 = pretrain_1080 = None
    if not args.no_train:
        model.NET.engine().plans.clear()           # release the forward plan's activations
        del plan
        import gc
        from tcvom_b200.engine import release_idle_pools
        gc.collect()
        release_idle_pools()
        torch.cuda.empty_cache()
        train = run_train_section(args, rank, world, dev, barrier, max_over_ranks)
        try:
            train_1080 = run_train_section(args, rank, world, dev, barrier, max_over_ranks,
                                           shape=(1, 5, H, W, 3, TRAIN_1080_GFLOP_PER_SAMPLE, "configs[3]"))
        except Exception as e:                                 # noqa: BLE001 - reported in the JSON line
            if world > 1:
                raise                                          # a rank that fails alone would dead-lock the others
            train_1080 = dict(error=f"{type(e).__name__}: {e}")
        if world == 1:
            try:
                pretrain_1080 = run_train_section(args, rank, world, dev, barrier, max_over_ranks, freeze=True,
                                                  shape=(1, 5, H, W, 3, None, "configs[3] shape"))
            except Exception as e:                             # noqa: BLE001 - reported in the JSON line
                pretrain_1080 = dict(error=f"{type(e).__name__}: {e}")

    fba = None
    if not args.no_fba:
        import gc
        from tcvom_b200.engine import release_idle_pools
        model.NET.engine().plans.clear()
        gc.collect()
        release_idle_pools()
        torch.cuda.empty_cache()
        fba = run_fba_section(args, rank, world, dev, max_over_ranks)
    dim = index = None
    if not args.no_dim:
        dim = run_dim_section(args, rank, world, dev, max_over_ranks)
        index = run_dim_section(args, rank, world, dev, max_over_ranks, arch="vmn_index")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    eager = None
    if world == 1 and not args.no_eager_baseline:
        try:
            eager = run_gpu_eager_section(dev)
        except Exception as e:                                 # noqa: BLE001
            eager = dict(error=f"{type(e).__name__}: {e}")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        hw = pick_cpu_sample(threads, 30.0, 1)
        t = min(cpu_oracle_time(hw, threads, reps=1))
        frac = window_gflop(*hw) / window_gflop(H, W)
        cpu = dict(value=frac / t, unit=UNIT, cores=threads, kind=cpu_kind(),
                   sample=f"one 3-frame {hw[0]}x{hw[1]} window ({t:.2f} s) = {frac:.4f} of a 1088x1920 window by the "
                          f"reference FLOP model")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16x3 (split-bf16 storage, fp32 accumulate)", data="synthetic",
                config=bench_config(world),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=imgs_u8.numel() + tris_u8.numel(), input_dtype="uint8",
                         upload="next window's frames copied on a side stream while the current one computes, the matte read "
                                "back on another; every copy inside the timed region",
                         d2h_bytes_per_step=out_h.numel() * 4),
                gpu_launches=launches, roofline=roof, cpu_baseline=cpu, gpu_eager_baseline=eager, train_step=train,
                train_step_1080p=train_1080, pretrain_step_1080p=pretrain_1080, fba_forward=fba, dim_forward=dim, index_forward=index)
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip the reference-on-this-GPU measurement (PyTorch eager, N=1 only)")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    ap.add_argument("--no-fba", action="store_true", help="skip the secondary FBA+TAM forward measurement (configs[4])")
    ap.add_argument("--no-dim", action="store_true", help="skip the secondary DIM+TAM / IndexNet+TAM forward measurements (SURVEY 8 f4)")
    ap.add_argument("--dump-calls", default=None, help="write the per-launch timing table (JSON lines) here")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else -- Python prints of the reference modules AND C-level writes to
    # fd 1 (NCCL prints its version banner there) -- is redirected to stderr
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
