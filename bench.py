#!/usr/bin/env python
"""bench.py -- headline benchmark of the MultiSWAG posterior-predictive hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port)

A STEP is one pass of the hot path over one batch of synthetic input: sample S weight vectors
from the SWAG posterior (K1) and evaluate all S x N_sys (system x weight-sample) pairs with the
fused predictive kernel (K2).  Workload at every N: BASELINE.json configs[1] per GPU -- one SWAG
model, 10,000 synthetic 3-planet systems x 1,000 weight samples = 1e7 evals per GPU per step
(weak scaling: systems are sharded, every rank holds its own 10k; at N>1 the step ends with the
path's single all_gather of the predictions).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_EVAL_V50 = 734_560  # 2*[100*(31*40+40*40+40*20) + (40*40+40*40+40*2)]  (SURVEY 8d, 31 live inputs)
N_SYS, N_SAMP = 10_000, 1_000
FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12  # TFLOP/s at clocks.max.sm (MEASURED_PEAKS.json sm_max_mhz)


def load_stats(seed=0):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"swag_v50_seed{seed}.npz"))
    return z, json.loads(str(z["hparams"])), json.loads(str(z["swa_params"]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except FileNotFoundError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].startswith("Active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def cpu_oracle_arm(steps, warmup, n_sys=10000, n_samp=4):
    """The reference's CPU path (oracle port: same torch ops as spock_reg_model.py, element-wise
    sampler) on a bounded sample of the workload, all host threads."""
    from bnn_chaos_model_b200 import synth
    from oracle import restatement as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    z, hp, sp = load_stats(0)
    spec = R.ModelSpec.from_hparams(hp)
    w_avg, w2_avg, pre_D = (torch.from_numpy(z[k]) for k in ("w_avg", "w2_avg", "pre_D"))
    x = torch.from_numpy(synth.make_systems(n_sys, seed=1))
    g = torch.Generator().manual_seed(0)

    def step():
        for _ in range(n_samp):  # forward_swag_fast: sample_weights + forward, per weight sample
            z1 = torch.randn((1, spec.d), generator=g)
            z2 = torch.randn((sp["K"], 1), generator=g)
            th = R.sample_weights(w_avg, w2_avg, pre_D, sp["K"], 0.5, z1, z2)
            e = torch.randn((n_sys, 40), generator=g)
            R.forward_swag_fast(spec, th, x, e[:, :20], e[:, 20:])

    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    return {"value": n_sys * n_samp / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"{n_sys} systems x {n_samp} weight samples per step (forward_swag_fast incl. sample_weights, "
                      f"oracle/restatement.py on torch CPU fp32, {steps} steps)"}, dt


def cpu_train_arm(B=2000, steps=2):
    """One SWAG-phase training step (noisy forward + loss + KL + autograd backward + clip + SGD) of the oracle
    port on the host cores, same batch shape as BASELINE configs[3]."""
    from bnn_chaos_model_b200 import synth
    from oracle import restatement as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    z, hp, sp = load_stats(0)
    spec = R.ModelSpec.from_hparams(hp)
    x = torch.from_numpy(synth.make_systems(B, seed=2))
    y = torch.from_numpy(synth.make_labels(B, seed=2))
    theta = torch.from_numpy(z["w_avg"]).clone()
    buf = None
    g = torch.Generator().manual_seed(0)

    def step(first):
        nonlocal theta, buf
        th = theta.clone().requires_grad_(True)
        e_in = torch.randn(x.shape, generator=g)
        e = torch.randn((B, 40), generator=g)
        es = torch.randn((B, 40), generator=g)
        total, _ = R.training_loss(spec, th, x, y, e_in, e[:, :20], e[:, 20:], es)
        (grad,) = torch.autograd.grad(total, th)
        theta, buf, _ = R.clip_and_sgd_step(theta, grad, buf, 1e-4, 0.9, 1e-14, 0.1 * spec.d, first)

    step(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(False)
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "unit": "seed-steps/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of 1 seed, batch {B} x 100 x 41 (oracle/restatement.py training_loss + autograd + SGD)"}


def gpu_train_arm(dev, n_seeds=4, B=2000, n_data=8000, iters=10):
    """BASELINE configs[3] per GPU: n_seeds SWAG models trained in one fused call per step (K4)."""
    from bnn_chaos_model_b200 import _lib, synth
    from bnn_chaos_model_b200 import spock_reg_model as S
    from bnn_chaos_model_b200._lib import TrainHParams

    lib = _lib.load()
    z, hp_, sp = load_stats(0)
    m = S.SWAGModel(hp_).init_params(sp).to(dev)
    cfg = m.config(100)
    x = torch.from_numpy(synth.make_systems(n_data, seed=3)).to(dev)
    y = torch.from_numpy(synth.make_labels(n_data, seed=3)).to(dev)
    theta = torch.from_numpy(z["w_avg"]).to(dev)[None].repeat(n_seeds, 1).contiguous()
    mom = torch.zeros_like(theta)
    met = torch.zeros((n_seeds, 8), device=dev)
    ws = torch.empty((lib.bnn_train_workspace_bytes(cfg, B, n_seeds) + 3) // 4, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    idx = torch.stack([torch.randperm(n_data, device=dev, generator=gen)[:B] for _ in range(n_seeds)]).to(torch.int32).contiguous()
    hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                      first_step=1, apply_update=1)

    def step(i):
        hp.first_step = int(i == 0)
        _lib.check(lib.bnn_train_step(cfg, hp, n_seeds, _lib.ptr(theta), _lib.ptr(mom), _lib.ptr(x), _lib.ptr(y),
                                      _lib.ptr(idx), B, None, None, None, 1, i, None, _lib.ptr(met), _lib.ptr(ws),
                                      _lib.current_stream_ptr()))

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        step(3 + i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    flop = 3 * 814_560 * B * n_seeds  # 3 x dense forward FLOPs per system
    tf = flop / (ms * 1e-3) / 1e12
    assert bool((met[:, 6] == 0).all()) and bool(torch.isfinite(met).all())
    return {"value": n_seeds / (ms * 1e-3), "unit": "seed-steps/s", "n_seeds": n_seeds, "batch": B, "ms_per_step": ms,
            "gpu_launches_per_step": 4, "data": "synthetic, 8000 resident systems, per-seed index batches, Philox noise",
            "roofline": {"bound": "fp32_fma", "achieved": tf, "peak": FP32_PEAK_NOMINAL, "unit": "TFLOP/s",
                         "frac": tf / FP32_PEAK_NOMINAL, "flop_per_seed_step": 3 * 814_560 * B}}


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the predict kernel from the committed ncu --set full summary
    (profiles/), per launch of this same workload; None when the summary is absent."""
    path = os.path.join(ROOT, "profiles", "r1_predict_tc4n4_ncu.txt")
    try:
        tot = 0.0
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        return tot or None
    except OSError:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_oracle_arm(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "MultiSWAG (system x weight-sample) evals/s", "value": cb["value"],
        "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 1 SWAG model (v50 seed 0), 10k synthetic systems x 1000 weight "
                               "samples per GPU; this arm times a bounded sample of it on the host CPU"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "train": cpu_train_arm(),
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist

    from bnn_chaos_model_b200 import _lib, synth
    from bnn_chaos_model_b200 import spock_reg_model as S
    from bnn_chaos_model_b200.multiswag import MultiSWAG, gather_system_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    z, hp, sp = load_stats(0)
    m = S.SWAGModel(hp).init_params(sp).to(dev)
    m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
    ens = MultiSWAG([m], device=dev)
    n_sys, n_samp = args.systems, args.samples
    n_total = n_sys * world
    lo = rank * n_sys
    xh = torch.from_numpy(synth.make_systems(n_sys, seed=1000 + rank)).pin_memory()
    x = xh.to(dev)
    cfg = ens.config(100)
    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    k_ev = []  # (start, end) of the predict kernel alone, per timed step

    def step(i, timed):
        _, thp = ens.sample_thetas(n_samp, seed=i)  # K1 + pack (2 launches)
        if timed:
            a, b = ev(), ev()
            a.record(stream)
        out = ens.predict(x, n_samp, seed=i, system_offset=lo, system_major=True, thp=thp)  # K2 (1 launch)
        if timed:
            b.record(stream)
            k_ev.append((a, b))
        if world > 1:
            out = gather_system_shards(out, n_total)
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, False)
    sync()
    clocks = ClockSampler(local_rank)
    clocks.start()
    t0, t1 = ev(), ev()
    t0.record(stream)
    for i in range(args.steps):
        out = step(args.warmup + i, True)
    t1.record(stream)
    sync()
    clk = clocks.stop()
    ms = t0.elapsed_time(t1) / args.steps
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([ms, k_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, k_ms = float(t[0]), float(t[1])
    assert bool(torch.isfinite(out).all())

    # ---- end to end through the public API with HOST buffers (pinned): H2D of x, D2H of the result
    out_h = torch.empty((n_sys, n_samp, 2), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        # the public host-buffer entry: samples theta, pipelines H2D / predict / D2H over chunks of systems
        ens.predict_host(xh, n_samp, seed=i, out_host=out_h, system_offset=lo)

    for i in range(max(1, args.warmup // 2)):
        e2e_step(i)
    sync()
    e0, e1 = ev(), ev()
    e0.record(stream)
    n_e2e = max(1, args.steps // 2)
    for i in range(n_e2e):
        e2e_step(100 + i)
    e1.record(stream)
    sync()
    te = torch.tensor([e0.elapsed_time(e1) / n_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])

    # ---- measured FFMA peak (roofline denominator cross-check), rank 0
    ffma = {}
    sink = torch.zeros(4, device=dev)
    for packed in (1, 0):
        import ctypes

        fl = ctypes.c_int64(0)
        for rep in range(3):
            a, b = ev(), ev()
            a.record(stream)
            _lib.check(lib.bnn_ffma_peak(packed, 4096, _lib.ptr(sink), ctypes.byref(fl), _lib.current_stream_ptr()))
            b.record(stream)
            torch.cuda.synchronize()
            tf = fl.value / (a.elapsed_time(b) * 1e-3) / 1e12
            ffma["f32x2" if packed else "f32"] = max(ffma.get("f32x2" if packed else "f32", 0.0), tf)

    train = gpu_train_arm(dev) if not args.no_train else None
    if train is not None and world > 1:  # seeds are sharded, no traffic: aggregate = sum over ranks, time = max
        tt = torch.tensor([train["ms_per_step"]], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train["ms_per_step"] = float(tt[0])
        train["n_seeds"] *= world
        train["value"] = train["n_seeds"] / (train["ms_per_step"] * 1e-3)
        train["roofline"]["note"] = "per-GPU fraction; value is the aggregate over ranks"

    if rank == 0:
        evals = n_sys * n_samp * world
        achieved = FLOP_PER_EVAL_V50 * n_sys * n_samp / (k_ms * 1e-3) / 1e12
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_oracle_arm(3, 1)
        line = {
            "metric": "MultiSWAG (system x weight-sample) evals/s", "value": evals / (ms * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"BASELINE configs[1] per GPU: 1 SWAG model (v50 seed-0 statistics), {n_sys} synthetic "
                            f"3-planet systems x {n_samp} weight samples, T=100, F=41 (31 live columns)",
                "systems_per_gpu": n_sys, "samples": n_samp, "parallelism": f"systems sharded x{world}, one all_gather",
                "l2": "inputs (164 MB x + 38 MB theta + 80 MB out per step) exceed the 126 MB L2; fresh theta every step",
            },
            "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": xh.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": e2e_ms},
            "gpu_launches": 3 * args.steps,
            "clocks": clk,
            "roofline": {"bound": "fp32_fma", "kernel": "predict (K2)", "achieved": achieved,
                         "peak": FP32_PEAK_NOMINAL, "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_NOMINAL,
                         "peak_source": "148 SM x 128 lanes x 2 x clocks.max.sm 1965 MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                        "no fp32 figure in MEASURED_PEAKS.json",
                         "measured_ffma_peak_tflops": ffma, "kernel_ms": k_ms,
                         "flop_per_eval": FLOP_PER_EVAL_V50, "traffic": ncu_traffic_bytes(),
                         "traffic_source": "profiles/r1_predict_tc4n4_ncu.txt (ncu --set full, same workload, per launch); "
                                           "algorithmic HBM bytes per launch: 0.32e9",
                         "note": "north_star's roofline for this kernel is the FP32 CUDA-core FMA peak; the kernel runs the "
                                 "feature MLP as 3xTF32 on tcgen05 (tensor pipe active 33 %, profiles/)"},
            "cpu_baseline": cb,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--systems", type=int, default=N_SYS)
    ap.add_argument("--samples", type=int, default=N_SAMP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
