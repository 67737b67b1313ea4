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
# all host cores, one forked process per core (the reference is not re-entrant: ei_run_dsp.h:251).  Every worker
# also returns the probabilities it computed, so the timed baseline doubles as the parity check of the GPU's outputs
# for the same clips (the synthetic stream is a pure function of the clip index).
# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, model, f32, first_clip, n_clips = args
    import numpy as np
    import eikws_pkg
    eikws_pkg.load()
    import eikws_b200.synth as synth
    from oracle_lib import PortOracle, RefOracle
    clips = synth.synth_clips(n_clips, first_clip=first_clip)
    if f32:  # the float the demo callback delivers: x / 32768 (exact)
        clips = (clips.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    if kind == "reference":
        return RefOracle(model).time_run_classifier_all(clips)
    o = PortOracle(model)
    t0 = time.perf_counter()
    probs = o.run_classifier_f32(clips) if f32 else o.run_classifier_i16(clips)
    return time.perf_counter() - t0, probs


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


def cpu_reference_throughput(clips_per_worker: int, first_clip: int = 0, model: str = None, f32: bool = False):
    """returns (clips/s aggregate, cores, kind, sample description, probs [cores * clips_per_worker][labels] of the clips
    first_clip .. first_clip + cores * clips_per_worker - 1 of the synthetic stream)"""
    import multiprocessing as mp
    import numpy as np
    from oracle_lib import have_ref
    model = model or MODEL
    kind = "reference" if have_ref(model) else "port"
    cores = usable_cores()
    ctx = mp.get_context("fork")
    jobs = [(kind, model, f32, first_clip + w * clips_per_worker, clips_per_worker) for w in range(cores)]
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)  # each worker times only its run_classifier loop (clip synthesis excluded)
    compute = max(r[0] for r in res)         # all workers run concurrently: the slowest one bounds the aggregate
    total = cores * clips_per_worker
    sample = f"{total} clips of the same synthetic stream ({clips_per_worker}/core x {cores} forked processes), {compute:.2f} s"
    return total / compute, cores, kind, sample, np.concatenate([r[1] for r in res])


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


WORKLOADS = {  # model -> (BASELINE.json config, description)
    "l476": ("configs[1]", "l476_yes_no (EON-compiled int8, 4 labels)"),
    "gsc12": ("configs[3]", "synthesised 12-label int8 model (BASELINE config 4)"),
    "l476f32": ("configs[4]", "float32 twin of l476 (BASELINE config 5)"),
    "l432": ("extra model", "l432 (int8, 3 labels)"),
    "zip6": ("extra model", "third shipped model (Arduino zip, int8, 6 labels)"),
    "dw3": ("extra model", "l432 graph with a DEPTHWISE_CONV_2D second block (int8, 3 labels)"),
}


def workload_config(model, f32_input, clips_per_gpu, n_gpus):
    which = WORKLOADS[model][0]
    if model == "l476" and n_gpus == 8 and clips_per_gpu * n_gpus == 1048576:
        which = "configs[2]"  # the same model, batch 1,048,576 over 8 GPUs
    bytes_per_clip = N_SAMPLES * (4 if f32_input else 2)
    return {"workload": f"BASELINE {which}: batch {clips_per_gpu} synthetic 1-s 16 kHz {'float32' if f32_input else 'int16'} clips per GPU "
                        f"({clips_per_gpu * n_gpus} in all), model {model} (MFCC+CMVN+CNN on the GPU, one C-ABI call), inputs "
                        f"({clips_per_gpu * bytes_per_clip / 1e9:.2f} GB/GPU) larger than L2",
            "model": WORKLOADS[model][1], "input": "float32 samples" if f32_input else "int16 PCM", "clips_per_gpu": clips_per_gpu,
            "sharding": f"{n_gpus} independent shard(s), no collective on the data path"}


def kernel_counters():
    """per-clip counters of the default kernel from the committed ncu capture (profiles/ncu_kernel_counters.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernel_counters.json")) as f:
            return json.load(f)
    except Exception:
        return None


def compare_with_reference(gpu_probs, ref_probs, f32_model):
    """parity of the first len(ref_probs) clips: bit-exact for an int8 classifier, 1e-5 absolute for the float32 graph
    (north_star's tolerance; the float softmax goes through expf, GPU vs glibc)"""
    import numpy as np
    g = gpu_probs[: ref_probs.shape[0]]
    if f32_model:
        bad = np.any(~(np.abs(g - ref_probs) <= 1e-5), axis=1)
        tol = "1e-5 absolute (float32 graph)"
    else:
        bad = np.any(g != ref_probs, axis=1)
        tol = "bit-exact (int8 classifier)"
    return {"clips": int(ref_probs.shape[0]), "mismatched_clips": int(bad.sum()), "tolerance": tol,
            "max_abs_diff": float(np.max(np.abs(g - ref_probs))) if ref_probs.size else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=0, help="default: 65,536 (configs[1]); 131,072 at 8 GPUs (configs[2]: 1,048,576 clips over the box)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-clips-per-core", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra BASELINE configs (4: gsc12, 5: float32) after the headline")
    ap.add_argument("--also-steps", type=int, default=5)
    ap.add_argument("--model", default=MODEL, help="l476 (default, BASELINE configs[1]), gsc12 (config 4), l476f32 (config 5)")
    ap.add_argument("--f32-input", action="store_true", help="feed float32 samples (64,000 B/clip) instead of int16 PCM (config 5)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} does not match WORLD_SIZE {world}")
    n_gpus = world
    if args.clips_per_gpu <= 0:
        args.clips_per_gpu = 131072 if n_gpus == 8 else 65536
    config = workload_config(args.model, args.f32_input, args.clips_per_gpu, n_gpus)
    f32_model = args.model == "l476f32"

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        steps_done, t_total, clips_total, info = 0, 0.0, 0, None
        per_core = max(64, args.cpu_clips_per_core // 2)
        for s in range(args.warmup + args.steps):
            thr, cores, kind, sample, _ = cpu_reference_throughput(per_core, first_clip=s * 1000003, model=args.model, f32=args.f32_input)
            if s >= args.warmup:
                steps_done += 1
                n = cores * per_core
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

    # ------------------------------------------------------------------ also[]: the other BASELINE configs measured after the headline
    also_specs = []
    if not args.no_also and args.model == MODEL and not args.f32_input:
        also_specs.append({"model": "gsc12", "f32": False, "clips": args.clips_per_gpu, "cpu_per_core": 512})    # config 4
        also_specs.append({"model": "l476f32", "f32": True, "clips": 262144, "cpu_per_core": 256})               # config 5

    # ------------------------------------------------------------------ CPU baseline (before CUDA is touched: uses fork)
    cpu, ref_probs = None, None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        thr, cores, kind, sample, ref_probs = cpu_reference_throughput(args.cpu_clips_per_core, model=args.model, f32=args.f32_input)
        cpu = {"value": thr, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        for sp in also_specs:
            thr, cores, kind, sample, sp["ref_probs"] = cpu_reference_throughput(sp["cpu_per_core"], model=sp["model"], f32=sp["f32"])
            sp["cpu"] = {"value": thr, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    host_affinity = None
    if n_gpus > 1 and not os.environ.get("EIKWS_BENCH_NO_BIND"):
        host_affinity = bind_to_gpu_numa(local_rank)
    import numpy as np
    import torch
    import torch.distributed as dist
    import eikws_pkg
    eikws = eikws_pkg.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version banner off stdout (ONE JSON line there)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = hbm_peak()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    counters = kernel_counters()

    def measure(model, f32_input, n, steps, warmup):
        """device-resident timing of one configuration: returns (result dict, Impulse, clips, probs)"""
        imp = eikws.Impulse(model, device=local_rank)
        clips = imp.synth_clips_device(n, first_clip=rank * n)
        if f32_input:  # the float the demo callback would deliver: x / 32768 (exact)
            clips = (clips.to(torch.float32) / 32768.0).contiguous()
        algo_bytes = N_SAMPLES * (4 if f32_input else 2) + 4 * imp.label_count
        probs = torch.empty((n, imp.label_count), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(warmup):
            imp.run_classifier_device(clips, out=probs)
        barrier()
        # int16 clips + fused int8 classifier run as two kernels per step: CUDA events around each of them, recorded by the library on the
        # launch stream inside the timed region, give the dominant kernel's own duration (no synchronisation until the region ends)
        imp.set_kernel_timing(True)
        launches0 = imp.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
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
        try:
            spec_ms, cep_ms, timed = imp.split_kernel_ms()
        except eikws.EikwsError:
            timed = 0
        imp.set_kernel_timing(False)
        res = {"value": n_gpus * n * steps / (elapsed_ms * 1e-3), "ms_per_step": elapsed_ms / steps, "steps": steps, "warmup": warmup,
               "gpu_launches": int(launches), "clocks": clocks}
        if timed:
            pairs = timed / steps  # kernel pairs per step (batches beyond 65,536 clips run in chunks)
            spec_ms, cep_ms = spec_ms * pairs, cep_ms * pairs
            # two kernels per step (per chunk).  Dominant: eikws_logmel_kernel (PCM in, log-mel records out); the step-level figure keeps SURVEY 8(d)'s
            # 32,016 B per clip over both kernels
            LE_BYTES = 49 * 33 * 4
            in_bytes = N_SAMPLES * (4 if f32_input else 2)
            cep_name = "eikws_cepstral_f32_kernel" if model == "l476f32" else "eikws_cepstral_kernel"
            step_ms = elapsed_ms / steps
            # the dominant kernel of the step and its own algorithmic bytes per clip
            if spec_ms >= cep_ms:
                dom, dom_ms, dom_bytes = "eikws_logmel_kernel", spec_ms, in_bytes + LE_BYTES
                how = f"{in_bytes:,} B of samples read + 6,468 B of log-mel record written per clip"
            else:
                dom, dom_ms, dom_bytes = cep_name, cep_ms, LE_BYTES + 4 * imp.label_count
                how = f"6,468 B of log-mel record read + {4 * imp.label_count} B of probabilities written per clip (a latency- / issue-bound kernel: DCT, the reference's CMVN chains, the CNN)"
            achieved = n * dom_bytes / (dom_ms * 1e-3) / 1e9
            res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                               "peak_source": peak_src, "kernel": dom + " (the dominant of the step's two kernels)",
                               "algo_bytes_per_clip": dom_bytes, "algo_bytes_how": how,
                               "kernel_ms": dom_ms, "kernel_ms_how": f"CUDA events around the kernel on its launch stream, {timed} timed launches in {steps} steps",
                               "kernels": [{"name": "eikws_logmel_kernel", "ms": spec_ms, "share_of_step": spec_ms / step_ms},
                                           {"name": cep_name, "ms": cep_ms, "share_of_step": cep_ms / step_ms}],
                               "step": {"algo_bytes_per_clip": algo_bytes, "achieved": n * algo_bytes / (step_ms * 1e-3) / 1e9,
                                        "frac": n * algo_bytes / (step_ms * 1e-3) / 1e9 / peak, "ms": step_ms}}
        else:
            kernel_ms = elapsed_ms / max(launches, 1)
            achieved = n * algo_bytes / (kernel_ms * 1e-3) / 1e9
            res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                               "peak_source": peak_src, "kernel": "eikws_run_classifier_kernel (one launch per step)",
                               "algo_bytes_per_clip": algo_bytes, "kernel_ms": kernel_ms}
        return res, imp, clips, probs

    head, imp, clips, probs = measure(args.model, args.f32_input, args.clips_per_gpu, args.steps, args.warmup)
    n = args.clips_per_gpu
    parity = None
    if ref_probs is not None:
        parity = dict(against=cpu["kind"], **compare_with_reference(probs.cpu().numpy(), ref_probs, f32_model))

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

    def timed_max(fn, reps):
        """wall time of `reps` calls between barriers, max over ranks"""
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        barrier()
        return float(td.item())

    # ceiling of the e2e leg: nothing but the host->device copy of the same pinned buffer, on all ranks at once
    scratch = torch.empty_like(clips)
    h2d_s = timed_max(lambda: scratch.copy_(h_clips, non_blocking=True), args.e2e_steps)
    del scratch
    h2d_bytes = n * N_SAMPLES * clips.element_size()
    ceiling_gbs = n_gpus * h2d_bytes * args.e2e_steps / h2d_s / 1e9
    e2e_s = timed_max(e2e_step, args.e2e_steps)
    e2e_value = n_gpus * n * args.e2e_steps / e2e_s
    assert torch.equal(h_probs.to(dev), probs), "host-buffer path and device path disagree"
    del h_clips, clips, probs
    imp.close()
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE configs (device-resident timing only)
    also = []
    for sp in also_specs:
        r, imp2, c2, p2 = measure(sp["model"], sp["f32"], sp["clips"], args.also_steps, args.warmup)
        entry = {"config": workload_config(sp["model"], sp["f32"], sp["clips"], n_gpus), "metric": METRIC, "unit": UNIT, **r}
        if "ref_probs" in sp:
            entry["cpu_baseline"] = sp["cpu"]
            entry["parity"] = dict(against=sp["cpu"]["kind"], **compare_with_reference(p2.cpu().numpy(), sp["ref_probs"], sp["model"] == "l476f32"))
        also.append(entry)
        del c2, p2
        imp2.close()
        torch.cuda.empty_cache()

    if rank == 0:
        roof = head["roofline"]
        issue = None
        if counters and args.model == MODEL and not args.f32_input:  # counters of the committed ncu capture of this very kernel
            roof["traffic"] = counters["dram_bytes_per_clip"] * n
            roof["traffic_source"] = counters["source"]
            sm_mhz = (head["clocks"] or {}).get("sm_mhz") or 1965.0
            ceiling = sm_count * 4 * sm_mhz * 1e6 / counters["warp_inst_per_clip"]
            issue = {"inst_per_clip": counters["warp_inst_per_clip"], "ceiling_clips_s": ceiling, "frac": head["value"] / n_gpus / ceiling,
                     "how": f"{sm_count} SMs x 4 schedulers x {sm_mhz:.0f} MHz (median under load) / warp instructions per clip (ncu smsp__inst_executed.sum, "
                            f"{counters['source']}): the path is issue-bound, this is its honest ceiling"}
            if "kernels" in counters:
                issue["kernels"] = counters["kernels"]
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32+f64 MFCC (bit-exact to the reference), " + ("f32 CNN" if f32_model else "int8 CNN"), "data": "synthetic", "config": config,
                "clocks": head["clocks"],
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": n * imp.label_count * 4,
                        "steps": args.e2e_steps, "api": ("eikws_classify_f32_host" if args.f32_input else "eikws_classify_i16_host") + " (pinned host buffers, chunk ring on two streams)",
                        "gbs": n_gpus * h2d_bytes * args.e2e_steps / e2e_s / 1e9, "ceiling_gbs": ceiling_gbs, "frac_of_ceiling": (h2d_s / e2e_s),
                        "ceiling_how": f"cudaMemcpyAsync of the same pinned batch alone, {n_gpus} rank(s) concurrently, max over ranks"},
                "gpu_launches": head["gpu_launches"], "roofline": roof}
        if issue:
            line["issue"] = issue
        if host_affinity:
            line["host_affinity"] = host_affinity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity"] = parity
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
        if parity is not None and parity["mismatched_clips"]:
            raise SystemExit(f"PARITY FAILURE against the {parity['against']}: {parity['mismatched_clips']} of {parity['clips']} clips differ")
        for e in also:
            if e.get("parity", {}).get("mismatched_clips"):
                raise SystemExit(f"PARITY FAILURE ({e['config']['model']}): {e['parity']}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
