#!/usr/bin/env python
"""bench.py -- headline benchmark of the run_classifier hot path (BASELINE.json metric: 1-second 16 kHz clips/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--clips-per-gpu C]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of run_classifier (MFCC + int8 CNN, one fused kernel launch) over one batch of synthetic
int16 clips.  N=1 workload = BASELINE.json configs[1]: batch 65,536 clips, L476 4-label int8 model, 1xB200.
For N>1 every rank owns its own 65,536-clip shard (weak scaling, no collective on the data path; clips are
independent).  `value` is measured with the batch resident in HBM; `e2e` goes through the host-buffer C-ABI call
(eikws_classify_i16_host) with pinned host buffers, H2D of the clips and D2H of the probabilities inside the
timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "clips_per_sec"
UNIT = "1-s 16 kHz clips/s"
MODEL = "l476"
N_SAMPLES = 16000
ALGO_BYTES_PER_CLIP = N_SAMPLES * 2 + 4 * 4  # int16 clip in + 4 float probabilities out = 32,016 B
FALLBACK_HBM_GBS = 6650.0                    # /opt/skills/guides/B200_PROFILING.md fallback


# ---------------------------------------------------------------------------------------------------------------
# CPU reference arm: the unmodified reference (oracle/_ref, built from /root/reference in the build container) on
# all host cores, one forked process per core (the reference is not re-entrant: ei_run_dsp.h:251).
# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, first_clip, n_clips = args
    import numpy as np
    import eikws_pkg
    eikws_pkg.load()
    import eikws_b200.synth as synth
    from oracle_lib import PortOracle, RefOracle
    clips = synth.synth_clips(n_clips, first_clip=first_clip)
    if kind == "reference":
        o = RefOracle(MODEL)
        return o.time_run_classifier_i16(clips)
    o = PortOracle(MODEL)
    t0 = time.perf_counter()
    o.run_classifier_i16(clips)
    return time.perf_counter() - t0


def usable_cores() -> int:
    """host cores this process may actually use: the affinity mask capped by the cgroup CPU quota (cpu.max)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_reference_throughput(clips_per_worker: int, first_clip: int = 0):
    """returns (clips/s aggregate, cores, kind, sample description)"""
    import multiprocessing as mp
    from oracle_lib import have_ref
    kind = "reference" if have_ref(MODEL) else "port"
    cores = usable_cores()
    ctx = mp.get_context("fork")
    jobs = [(kind, first_clip + w * clips_per_worker, clips_per_worker) for w in range(cores)]
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, jobs)  # each worker times only its run_classifier loop (clip synthesis excluded)
    compute = max(times)                     # all workers run concurrently: the slowest one bounds the aggregate
    total = cores * clips_per_worker
    sample = f"{total} clips of the same synthetic stream ({clips_per_worker}/core x {cores} forked processes), {compute:.2f} s"
    return total / compute, cores, kind, sample


# ---------------------------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 5 ms on a host thread, from before the warm-up to the end
    of the timed region; stop(t0, t1) reports the samples that fall inside the timed region (all samples if the region
    was shorter than the sampling period).  Falls back to one `nvidia-smi` query when pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []  # (t, sm_mhz, reasons bitmask)
        self.thread = None
        self.max_mhz = None
        self._stop = False

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop:
                    try:
                        self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t0=None, t1=None):
        self._stop = True
        if self.thread is None:
            return self._smi_once()
        self.thread.join(timeout=2)
        inside = [x for x in self.samples if t0 is not None and t0 <= x[0] <= t1]
        use = inside if inside else self.samples
        sm = sorted(x[1] for x in use)
        mask = 0
        for x in use:
            mask |= x[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": [n for n, b in self.REASONS if mask & b],
                "samples": len(use), "samples_in_timed_region": len(inside), "source": "nvml, 5 ms period"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi after the run (pynvml unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def bind_to_gpu_numa(gpu_index):
    """Pin this rank to the host cores NVML reports as local to its GPU BEFORE the pinned host buffers of the e2e leg are
    allocated (pages are placed on the allocating thread's NUMA node): with 8 ranks streaming 55 GB/s each, buffers on the
    wrong socket turn the PCIe-bound leg into an inter-socket-link-bound one.  Returns a short description for `config`."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(local & allowed)
        if not use:
            return "unchanged (no GPU-local core in the allowed set)"
        os.sched_setaffinity(0, use)
        return f"{len(use)} GPU-local cores (NVML affinity of GPU {idx})"
    except Exception as e:  # plumbing only: never fail the bench over it
        return f"unchanged ({type(e).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=65536)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-clips-per-core", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=MODEL, help="l476 (default, BASELINE configs[1]), gsc12 (config 4), l476f32 (config 5)")
    ap.add_argument("--f32-input", action="store_true", help="feed float32 samples (64,000 B/clip) instead of int16 PCM (config 5)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} does not match WORLD_SIZE {world}")
    n_gpus = world
    which = {"l476": "BASELINE configs[1]", "gsc12": "BASELINE configs[3]", "l476f32": "BASELINE configs[4]"}.get(args.model, "extra model")
    bytes_per_clip = N_SAMPLES * (4 if args.f32_input else 2)
    config = {"workload": f"{which}: batch {args.clips_per_gpu} synthetic 1-s 16 kHz {'float32' if args.f32_input else 'int16'} clips per GPU, "
                          f"model {args.model} (MFCC+CMVN+CNN fused in one kernel), inputs ({args.clips_per_gpu * bytes_per_clip / 1e9:.2f} GB/GPU) larger than L2",
              "model": {"l476": "l476_yes_no (EON-compiled int8, 4 labels)", "gsc12": "synthesised 12-label int8 model (BASELINE config 4)",
                        "l476f32": "float32 twin of l476 (BASELINE config 5)", "l432": "l432 (int8, 3 labels)",
                        "zip6": "third shipped model (Arduino zip, int8, 6 labels, generic op plan)"}[args.model],
              "input": "float32 samples" if args.f32_input else "int16 PCM", "clips_per_gpu": args.clips_per_gpu,
              "sharding": f"{n_gpus} independent shard(s), no collective on the data path"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        steps_done, t_total, clips_total, info = 0, 0.0, 0, None
        for s in range(args.warmup + args.steps):
            thr, cores, kind, sample = cpu_reference_throughput(max(64, args.cpu_clips_per_core // 2), first_clip=s * 1000003)
            if s >= args.warmup:
                steps_done += 1
                n = cores * max(64, args.cpu_clips_per_core // 2)
                t_total += n / thr
                clips_total += n
                info = (cores, kind, sample)
        value = clips_total / t_total
        cores, kind, sample = info
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_total / steps_done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+int8",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": "per step: " + sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ CPU baseline (before CUDA is touched: uses fork)
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        thr, cores, kind, sample = cpu_reference_throughput(args.cpu_clips_per_core)
        cpu = {"value": thr, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if n_gpus > 1 and not os.environ.get("EIKWS_BENCH_NO_BIND"):
        config["host_affinity"] = bind_to_gpu_numa(local_rank)
    import torch
    import torch.distributed as dist
    import eikws_pkg
    eikws = eikws_pkg.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version banner off stdout (ONE JSON line there)
        dist.init_process_group("nccl", device_id=dev)
    imp = eikws.Impulse(args.model, device=local_rank)
    n = args.clips_per_gpu
    clips = imp.synth_clips_device(n, first_clip=rank * n)
    if args.f32_input:  # the float the demo callback would deliver: x / 32768 (exact)
        clips = (clips.to(torch.float32) / 32768.0).contiguous()
    algo_bytes = N_SAMPLES * (4 if args.f32_input else 2) + 4 * imp.label_count
    probs = torch.empty((n, imp.label_count), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        imp.run_classifier_device(clips, out=probs)
    barrier()
    launches0 = imp.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        imp.run_classifier_device(clips, out=probs)
    ev1.record()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = imp.launch_count - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    elapsed_ms = float(t.item())
    value = n_gpus * n * args.steps / (elapsed_ms * 1e-3)

    # ------------------------------------------------------------------ end to end through the host-buffer C ABI
    h_clips = torch.empty((n, N_SAMPLES), dtype=clips.dtype, pin_memory=True)
    h_clips.copy_(clips)
    h_probs = torch.empty((n, imp.label_count), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()
    import ctypes as C
    lib = eikws.load_library()

    def e2e_step():
        fn = lib.eikws_classify_f32_host if args.f32_input else lib.eikws_classify_i16_host
        rc = fn(imp._h, C.c_void_p(h_clips.data_ptr()), n, C.c_void_p(h_probs.data_ptr()))
        assert rc == 0, lib.eikws_last_error()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * n * args.e2e_steps / float(te.item())
    assert torch.equal(h_probs.to(dev), probs), "host-buffer path and device path disagree"

    if rank == 0:
        peak, peak_src = hbm_peak()
        kernel_ms = elapsed_ms / max(launches, 1)
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch from the committed ncu capture of this kernel, scaled to this batch (int16 fused path only)
            if args.model == "l476" and not args.f32_input:
                with open(os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["dram_bytes_per_clip"] * n, tj["source"]
        except Exception:
            pass
        achieved = n * algo_bytes / (kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32+f64 MFCC (bit-exact to the reference), int8 CNN", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * N_SAMPLES * clips.element_size(), "d2h_bytes_per_step": n * imp.label_count * 4,
                        "steps": args.e2e_steps, "api": ("eikws_classify_f32_host" if args.f32_input else "eikws_classify_i16_host") + " (pinned host buffers, 8192-clip chunks on two streams)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "eikws_run_classifier_kernel (one launch per step)",
                             "algo_bytes_per_clip": algo_bytes, "kernel_ms": kernel_ms}}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
