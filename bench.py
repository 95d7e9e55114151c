#!/usr/bin/env python
"""bench.py -- headline benchmark of napa-fft3-b200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

A STEP is one pass of the hot path over one batch of synthetic input already resident in HBM.
Default workload is BASELINE.json configs[2] ("c3"), the largest single-GPU configuration and the one the
>= 70 % HBM-roofline target of `north_star` is stated on: 1024 x 2^20 complex doubles, fft :in-order nil followed by
ifft :in-order nil with a shared complex window (the convolution building block).  A short measurement of c2
(4096 x 2^16, the config `metric` is quoted on) and c4 rides along in `other_workloads`.
Metric: FP64 complex FFT GFLOP/s = 5 N log2 N per transform / time (BASELINE.json `metric`).
N > 1: one process per GPU (torchrun), the batch is sharded with no data-path collective
(weak scaling: every GPU gets the full per-GPU batch); time = max over ranks.  The line then also carries a `dist`
record: ONE 2^(27 + log2 N)-point transform spread over the N GPUs (config c5: distributed four-step, NCCL
all-to-all over NVLink), with its time, the exchange alone and a parity figure.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, batch, description)
    "c1": (1 << 10, 1, "fft n=2^10 single transform, :in-order t :scale nil"),
    "c2": (1 << 16, 4096, "4096 x 2^16 complex doubles: windowed-fft (shared Hann, :in-order t) -> ifft :scale :inv"),
    "c3": (1 << 20, 1024, "1024 x 2^20: fft :in-order nil -> ifft :in-order nil :window H (shared complex filter)"),
    "c4": (1 << 14, 8192, "8192 x 2^14 reals: rfft :scale nil -> rifft :scale t"),
    # one transform spread over all ranks (distributed four-step, one all-to-all); strong scaling
    "c5": (1 << 30, 1, "single 2^30-point c2c, distributed four-step, all-to-all over NVLink"),
}


def bench_config(wl):
    """the `config` object of the JSON line: identical for the GPU arm and the reference arm"""
    n, batch, desc = WORKLOADS[wl]
    l2 = ("inputs_exceed_l2 (no flush: every array of a step is far larger than the 126 MB L2)" if wl != "c1" else
          "none (one 16 KiB transform per step: latency-bound by definition, its data never leaves L2)")
    return {"workload": wl, "description": desc, "n": n, "batch_per_gpu": batch, "l2_policy": l2}


# FP64 thread-instructions one step EXECUTES (DADD + DMUL + DFMA, dynamic counts of the round-2 `ncu --set full`
# captures, profiles/r02_ncu_full_*: smsp__inst_executed x the FP64 share of the source-page mix x 32), and the measured
# FP64 issue rate of the chip (profiles/r02_fp64_rate.txt: 63.4 per clock per SM, 18.44e12 per second): the second
# roofline of this double-precision path.  For C3 the FP64-pipe time at peak is 92 % of the HBM time at peak.
FP64_INSTR_PER_STEP = {"c2": 4096 * 65536 * 126.0, "c3": 1024 * 1048576 * 166.0, "c4": 2 * 128e6 * 32.0}
FP64_PEAK_INSTR_PER_S = 18.44e12


def flops_per_step(wl):
    n, batch, _ = WORKLOADS[wl]
    lg = np.log2(n)
    if wl in ("c1", "c5"):
        return 5.0 * n * lg * batch
    if wl == "c4":
        return 2 * 2.5 * n * lg * batch
    return 2 * 5.0 * n * lg * batch


def algorithmic_bytes_per_step(wl):
    """c2c 32 N B per transform (+ shared window once), rfft/rifft 24 N."""
    n, batch, _ = WORKLOADS[wl]
    if wl in ("c1", "c5"):
        return 32.0 * n * batch
    if wl == "c4":
        return 2 * 24.0 * n * batch
    win = 8.0 * n if wl == "c2" else 16.0 * n
    return 2 * 32.0 * n * batch + win


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle (C restatement of Napa-FFT3; SBCL is not available)
# ---------------------------------------------------------------------------------------------
def cpu_oracle_step(wl, items, threads, state):
    """One bounded step of the same workload on the host cores, batch-parallel over `threads`."""
    from oracle import napa_oracle as O
    n = WORKLOADS[wl][0]
    data, win = state["data"], state["win"]

    def work(lo, hi):
        if wl == "c4":
            for i in range(lo, hi):
                O.rifft(O.rfft(data[i]))
            return
        buf = state["work"][lo:hi]
        np.copyto(buf, data[lo:hi])
        if wl == "c1":
            O.c2c_batch(buf, n, 1)
        elif wl == "c2":
            O.c2c_batch(buf, n, 1, True, None, win)
            O.c2c_batch(buf, n, -1, True, True)
        else:
            O.c2c_batch(buf, n, 1, False)
            O.c2c_batch(buf, n, -1, False, True, win)
    threads = max(1, min(threads, items))
    if threads == 1:
        return work(0, items)                  # no thread spawn around a microsecond-scale step (c1)
    per = (items + threads - 1) // threads
    ts = [threading.Thread(target=work, args=(t * per, min(items, (t + 1) * per))) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


def cpu_oracle_state(wl, items):
    from oracle import napa_oracle as O
    n = WORKLOADS[wl][0]
    O.lib().napa_oracle_warm(n)
    rng = np.random.Generator(np.random.Philox(1002))
    if wl == "c4":
        data = rng.uniform(-1, 1, (items, n))
        O.rifft(O.rfft(data[0]))
        win = None
    else:
        data = rng.uniform(-1, 1, (items, n)) + 1j * rng.uniform(-1, 1, (items, n))
        win = O.window_vector("hann", n) if wl == "c2" else (O.fft(data[0], in_order=False) if wl == "c3" else None)
    return {"data": data, "win": win, "work": np.empty_like(data)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    n, batch, desc = WORKLOADS[wl]
    threads = os.cpu_count() or 1
    items = min(batch, (64 if n <= (1 << 16) else 4) * threads)      # ~1 s of host work per step
    threads = max(1, min(threads, items))
    state = cpu_oracle_state(wl, items)
    for _ in range(args.warmup):
        cpu_oracle_step(wl, items, threads, state)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(wl, items, threads, state)
    dt = (time.perf_counter() - t0) / args.steps
    gflops = flops_per_step(wl) * (items / batch) / dt / 1e9
    sample = "%d of %d transforms per step on %d threads (oracle: C restatement of Napa-FFT3, SBCL unavailable)" % (items, batch, threads)
    line = {
        "impl": "reference", "metric": "fp64_complex_fft_gflops", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(wl),
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        return self.summarise(t_begin, t_end)

    def summarise(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t_begin or ts > t_end + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2]); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    # keep stdout to the one JSON line: NCCL writes its version banner / warnings to stdout unless told otherwise
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="napa", choices=["napa", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other_workloads / dist riders")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "napa" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import napa_fft3_b200 as napa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # NCCL's version banner is written through C stdio when the communicator comes up: force that now and
        # flush libc's buffer, so that the JSON line below is the last thing on stdout
        dist.barrier()
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NAPAFFT_DEVICE"] = str(local)
    b = napa.default_binding()                      # CUDA library or bust

    wl = args.workload
    n, batch, desc = WORKLOADS[wl]
    if wl == "c5":
        run_c5(args, torch, dist, napa, b, world, rank, local, dev)
        return
    m = measure_workload(args, torch, dist, napa, b, wl, world, rank, local, dev, args.steps, args.warmup)
    ms_per_step, launches, clocks, parity, win = m["ms_per_step"], m["launches"], m["clocks"], m["parity"], m["win"]
    value = flops_per_step(wl) * world / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (fft_pass_kernel) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = algorithmic_bytes_per_step(wl) / (ms_per_step * 1e-3) / 1e9
    launches_per_step = launches / max(args.steps, 1)
    traffic, traffic_src = None, None
    try:
        # DRAM bytes of one whole step summed over its launches, from this round's `ncu --set full` capture of the same
        # command (tools/profile_workload.sh writes the file next to the summaries it comes from), per launch like `achieved`
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(wl)
        traffic = float(t["dram_bytes_per_step"]) / max(launches_per_step, 1)
        traffic_src = t.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "napa::fft_pass_kernel<LOG2L,KIND,LEAN> (one launch per digit pass; two per 2^14..2^22 transform)",
                "algorithmic_bytes_per_launch": algorithmic_bytes_per_step(wl) / max(launches_per_step, 1),
                "avg_launch_ms": ms_per_step / max(launches_per_step, 1),
                "frac_of_nominal_8000": achieved / 8000.0}
    if wl in FP64_INSTR_PER_STEP:
        t_fp64 = FP64_INSTR_PER_STEP[wl] / FP64_PEAK_INSTR_PER_S * 1e3
        t_hbm = algorithmic_bytes_per_step(wl) / (peak * 1e9) * 1e3
        roofline["fp64_pipe"] = {"instr_per_step": FP64_INSTR_PER_STEP[wl], "peak_instr_per_s": FP64_PEAK_INSTR_PER_S,
                                 "ms_at_peak": t_fp64, "hbm_ms_at_peak": t_hbm, "frac": t_fp64 / ms_per_step,
                                 "note": "second roofline of the path: FP64-pipe time at the measured issue rate "
                                         "(profiles/r02_fp64_rate.txt) next to the HBM time at the measured copy bandwidth"}

    # ---- e2e: the same step through the host-buffer C ABI (pinned host arrays, PCIe inside) ----
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, torch, dist, napa, b, wl, world, dev, win)
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": "GFLOP/s", "error": str(ex)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        items = min(batch, 2 * (os.cpu_count() or 1))
        threads = max(1, min(os.cpu_count() or 1, items))       # threads actually used (1 for the single-transform config)
        st = cpu_oracle_state(wl, items)
        cpu_oracle_step(wl, items, threads, st)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 8.0 and reps < 50:
            cpu_oracle_step(wl, items, threads, st); reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": flops_per_step(wl) * (items / batch) / dt / 1e9, "unit": "GFLOP/s", "cores": threads,
                        "kind": "port",
                        "sample": "%d of %d transforms per step, %d reps, batch-parallel over %d threads; oracle = C "
                                  "restatement of Napa-FFT3 (SBCL unavailable)" % (items, batch, reps, threads)}

    # ---- riders: the other single-GPU configs (short runs) and, for N > 1, the distributed transform ----
    others, dist_rec = None, None
    if not args.no_extras:
        others = {}
        for w2 in ("c2", "c4"):
            if w2 == wl:
                continue
            try:
                m2 = measure_workload(args, torch, dist, napa, b, w2, world, rank, local, dev, max(5, args.steps // 5), 3, clocks=False)
                a2 = algorithmic_bytes_per_step(w2) / (m2["ms_per_step"] * 1e-3) / 1e9
                others[w2] = {"description": WORKLOADS[w2][2], "ms_per_step": m2["ms_per_step"], "steps": m2["steps"],
                              "value": flops_per_step(w2) * world / (m2["ms_per_step"] * 1e-3) / 1e9, "unit": "GFLOP/s",
                              "roofline_frac": a2 / peak, "parity_nre_vs_oracle": m2["parity"]}
            except Exception as ex:
                others[w2] = {"error": str(ex)[:200]}
        if world > 1:
            try:
                lg = 27 + int(np.log2(world))
                dist_rec = measure_c5(args, torch, dist, napa, b, world, rank, local, dev, lg, max(5, args.steps // 5), 3)
            except Exception as ex:
                dist_rec = {"error": str(ex)[:300]}

    if rank == 0:
        line = {
            "metric": "fp64_complex_fft_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(wl),
            "timing": {"l2_policy": "inputs_exceed_l2 (%.1f GiB per array per GPU)" % (16.0 * n * batch / 2**30),
                       "sharding": "batch-sharded, no collective" if world > 1 else "single GPU",
                       "clock": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity_nre_vs_oracle": parity,
            "other_workloads": others, "dist": dist_rec,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_workload(args, torch, dist, napa, b, wl, world, rank, local, dev, steps, warmup, clocks=True):
    """`steps` timed steps of workload `wl` on device-resident synthetic data; the inputs are freed on return."""
    n, batch, desc = WORKLOADS[wl]
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev); g.manual_seed(1002 + rank)

    # ---- synthetic inputs resident in HBM (re, im ~ U[-1,1), test-support.lisp:19-25) ----
    if wl == "c4":
        x = (torch.rand((batch, n), dtype=torch.float64, device=dev, generator=g) * 2 - 1)
        y = torch.empty((batch, n), dtype=torch.complex128, device=dev)
        z = torch.empty((batch, n), dtype=torch.float64, device=dev)
        win = None
    else:
        x = torch.view_as_complex(torch.rand((batch, n, 2), dtype=torch.float64, device=dev, generator=g) * 2 - 1).contiguous()
        y = torch.empty_like(x)
        z = torch.empty_like(x)
        if wl == "c2":
            win = torch.from_numpy(napa.window_vector(napa.hann, n)).to(dev)      # (hann i n), windowing.lisp:10
        elif wl == "c3":
            hw = torch.view_as_complex(torch.rand((n, 2), dtype=torch.float64, device=dev, generator=g) * 2 - 1).contiguous()
            win = torch.empty_like(hw)
            b.c2c_dev(hw.data_ptr(), win.data_ptr(), n, 1, n, 1, 0, 0, None, 0, 0, stream)   # H = fft h :in-order nil
        else:
            win = None
    wp = win.data_ptr() if win is not None else None

    def step():
        if wl == "c1":
            b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, 1, 0, 1, None, 0, 0, stream)
        elif wl == "c2":
            b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, 1, 0, 1, wp, 1, 0, stream)
            b.c2c_dev(y.data_ptr(), z.data_ptr(), n, batch, n, -1, 1, 1, None, 0, 0, stream)
        elif wl == "c3":
            b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, 1, 0, 0, None, 0, 0, stream)
            b.c2c_dev(y.data_ptr(), z.data_ptr(), n, batch, n, -1, 1, 0, wp, 2, 0, stream)
        else:
            b.rfft_dev(x.data_ptr(), y.data_ptr(), n, batch, 0, stream)
            b.rifft_dev(y.data_ptr(), z.data_ptr(), n, batch, 1, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if clocks:
        sampler.start()
        time.sleep(0.3)
    t_warm = time.perf_counter()
    for _ in range(warmup):
        step()
    barrier()
    launches0 = b.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = b.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clk = None
    if clocks:
        clk = sampler.stop(t_begin, t_end)
        clk["window"] = "timed region"
        if not clk.get("samples") or clk["samples"] < 3:
            clk = sampler.summarise(t_warm, t_end)
            clk["window"] = "warm-up + timed region (timed region shorter than 3 samples at 100 ms)"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()

    # ---- quick self-check of the timed path against the oracle (not timed) ----
    parity = None
    if rank == 0:
        from oracle import napa_oracle as O
        i = batch // 2
        if wl == "c4":
            want = O.rifft(O.rfft(x[i].cpu().numpy()))
        elif wl == "c2":
            want = O.ifft(O.fft(x[i].cpu().numpy(), window=win.cpu().numpy()), scale=True)
        elif wl == "c3":
            want = O.ifft(O.fft(x[i].cpu().numpy(), in_order=False), in_order=False, scale=True, window=win.cpu().numpy())
        else:
            want = O.fft(x[i].cpu().numpy())
        got = (y if wl == "c1" else z)[i].cpu().numpy()
        parity = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    win_host = win.cpu() if win is not None else None
    del x, y, z
    torch.cuda.empty_cache()
    return {"ms_per_step": ms / steps, "steps": steps, "launches": launches, "clocks": clk, "parity": parity, "win": win_host}


def measure_c5(args, torch, dist, napa, b, world, rank, local, dev, lg, steps, warmup):
    """ONE 2^lg-point transform over all ranks (distributed four-step, napa_fft3_b200.dist.DistributedFFT over
    napafft_dist_c2c_dev); time = max over ranks.  Parity: the input is a handful of impulses, so every bin is known in
    closed form; a sample of local bins is compared (normwise, against the 1e-13 log2 N tolerance)."""
    from napa_fft3_b200.dist import DistributedFFT
    n = 1 << lg
    f = DistributedFFT(lg, b, dist if world > 1 else None, dev)
    rng = np.random.Generator(np.random.Philox(1005))
    pos = np.unique(np.concatenate([[0, 1, n - 1, n // 2 + 1], rng.integers(0, n, 28)]))
    amp = rng.uniform(-1, 1, len(pos)) + 1j * rng.uniform(-1, 1, len(pos))
    g = torch.Generator(device=dev); g.manual_seed(1005 + rank)
    x = torch.view_as_complex(torch.rand((f.N1, f.B, 2), dtype=torch.float64, device=dev, generator=g) * 2 - 1).contiguous()
    out = torch.empty((f.R, f.N2), dtype=torch.complex128, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        f.forward(x, out)
    barrier()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    launches0 = b.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(steps):
        f.forward(x, out)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = b.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    # the exchange alone (torch.distributed all_to_all_single on buffers of the same size), for the NVLink roofline
    f.exchange_only(2)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a0.record()
    f.exchange_only(steps)
    a1.record()
    barrier()
    a2a_ms = a0.elapsed_time(a1) / steps
    if world > 1:
        t = torch.tensor([ms, a2a_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, a2a_ms = t[0].item(), t[1].item()
    ms_per_step = ms / steps
    # parity (not timed): impulses in, closed-form spectrum on sampled local bins; then the round trip of the random input
    xi = torch.zeros((f.N1, f.B), dtype=torch.complex128, device=dev)
    for p_, a_ in zip(pos, amp):
        j1, j2 = divmod(int(p_), f.N2)
        if f.rank * f.B <= j2 < (f.rank + 1) * f.B:
            xi[j1, j2 - f.rank * f.B] = complex(a_)
    Xi = f.forward(xi)
    sel = rng.integers(0, f.R * f.N2, 400)
    k = (f.rank * f.R + sel // f.N2) + f.N1 * (sel % f.N2)
    ph = (pos[None, :].astype(object) * k[:, None].astype(object)) % n
    want = (amp[None, :] * np.exp(-2j * np.pi * ph.astype(np.float64) / n)).sum(axis=1)
    got = Xi.view(-1)[torch.from_numpy(sel).to(dev)].cpu().numpy()
    nre = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    back = f.inverse(f.forward(x, out))
    rt = (torch.linalg.norm(back - x) / torch.linalg.norm(x)).item()
    if world > 1:
        t = torch.tensor([nre, rt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nre, rt = t[0].item(), t[1].item()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_bytes = 32.0 * n / world                      # compulsory HBM bytes per GPU
    a2a_bytes = f.alltoall_bytes_per_rank()
    rec = {
        "workload": "c5", "description": WORKLOADS["c5"][2], "log2n": lg, "n": n, "N1": f.N1, "N2": f.N2, "n_gpus": world,
        "exchange": ("%s, %d column chunks overlapping the column FFTs (napafft_dist_c2c_dev)" % (
            {0: "?", 1: "single rank: device copy", 2: "ncclSend / ncclRecv on the library's communicator",
             3: "copy-engine transfers into IPC-mapped peer buffers, fenced by NCCL all-reduces"}[int(b.L.napafft_dist_exchange_mode())], f.nch))
        if f.native else "torch.distributed all_to_all_single per chunk",
        "ms_per_step": ms_per_step, "steps": steps, "value": 5.0 * n * lg / (ms_per_step * 1e-3) / 1e9, "unit": "GFLOP/s",
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": hbm_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": hbm_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None,
                     "note": "per-GPU compulsory bytes (32 N / P) over the whole step, all-to-all included"},
        "nvlink": {"alltoall_bytes_per_gpu_each_way": a2a_bytes, "alltoall_alone_ms": a2a_ms,
                   "alltoall_alone_gbs_per_direction": (a2a_bytes / (a2a_ms * 1e-3) / 1e9) if a2a_ms > 0 and world > 1 else None,
                   "whole_transform_gbs_per_direction": (a2a_bytes / (ms_per_step * 1e-3) / 1e9) if world > 1 else None,
                   "peak_measured_ref": 770.0, "peak_nominal": 900.0,
                   "frac_of_measured_alltoall_alone": (a2a_bytes / (a2a_ms * 1e-3) / 1e9 / 770.0) if a2a_ms > 0 and world > 1 else None,
                   "frac_of_measured_whole_transform": (a2a_bytes / (ms_per_step * 1e-3) / 1e9 / 770.0) if world > 1 else None},
        "parity_nre_vs_closed_form": nre, "round_trip_nre": rt, "tolerance": 1e-13 * lg,
    }
    del x, out, xi, Xi, back
    torch.cuda.empty_cache()
    return rec


def run_c5(args, torch, dist, napa, b, world, rank, local, dev):
    """BASELINE config C5 as the main workload: strong scaling (total work fixed)."""
    lg = int(os.environ.get("NAPAFFT_C5_LOG2N", "30"))
    rec = measure_c5(args, torch, dist, napa, b, world, rank, local, dev, lg, args.steps, args.warmup)
    if rank == 0:
        line = {
            "metric": "fp64_complex_fft_gflops", "value": rec["value"], "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(bench_config("c5"), n=rec["n"]),
            "timing": {"l2_policy": "inputs_exceed_l2 (%.1f GiB per GPU)" % (16.0 * rec["n"] / world / 2**30),
                       "layout": "time side: column block of [N1][N2]; freq side: row block k1 of X[k1 + N1*k2]"},
            "clocks": rec["clocks"], "e2e": None, "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"],
            "nvlink": rec["nvlink"], "cpu_baseline": None, "parity_nre_vs_closed_form": rec["parity_nre_vs_closed_form"],
            "round_trip_nre": rec["round_trip_nre"], "dist": rec,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, napa, b, wl, world, dev, win):
    """Same step through the reference-facing host-pointer entry points: inputs and outputs are
    page-locked HOST arrays, every step pays H2D + D2H inside the timed region."""
    n, batch, _ = WORKLOADS[wl]
    rng = np.random.Generator(np.random.Philox(2002))
    if wl == "c4":
        hx = torch.empty((batch, n), dtype=torch.float64, pin_memory=True)
        hy = torch.empty((batch, n), dtype=torch.complex128, pin_memory=True)
        hz = torch.empty((batch, n), dtype=torch.float64, pin_memory=True)
        hx.numpy()[:] = rng.uniform(-1, 1, (batch, n))
    else:
        hx = torch.empty((batch, n), dtype=torch.complex128, pin_memory=True)
        hy = torch.empty((batch, n), dtype=torch.complex128, pin_memory=True)
        hz = hy
        v = hx.numpy().view(np.float64)
        chunk = max(1, (1 << 24) // n)
        for i in range(0, batch, chunk):
            v[i:i + chunk] = rng.uniform(-1, 1, v[i:i + chunk].shape)
    hwin = win.numpy() if win is not None else None
    x, y, z = hx.numpy(), hy.numpy(), hz.numpy()

    def step():
        if wl == "c1":
            napa.fft_batch(x, out=y)
        elif wl == "c2":
            napa.fft_batch(x, window=hwin, out=y)
            napa.fft_batch(y, direction=-1, scale="inv", out=z)
        elif wl == "c3":
            napa.fft_batch(x, in_order=False, out=y)
            napa.fft_batch(y, direction=-1, in_order=False, scale="inv", window=hwin, out=z)
        else:
            napa.rfft_batch(x, out=y)
            napa.rifft_batch(y, out=z)

    step()                                           # warm the staging buffers
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    dt /= args.e2e_steps
    if wl == "c4":
        h2d, d2h = (8 + 16) * n * batch, (16 + 8) * n * batch
    elif wl == "c1":
        h2d, d2h = 16 * n * batch, 16 * n * batch
    else:
        h2d = 2 * 16 * n * batch + (hwin.nbytes if hwin is not None else 0)
        d2h = 2 * 16 * n * batch
    return {"value": flops_per_step(wl) * world / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
            "api": "napafft_c2c / napafft_rfft / napafft_rifft on page-locked host arrays"}


if __name__ == "__main__":
    main()
