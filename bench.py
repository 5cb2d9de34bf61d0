#!/usr/bin/env python3
"""bench.py - the corpus-scan hot path on N B200s of one node (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W             # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement on the host cores

A step = one pass of the hot path (k_plan, k_scan, k_classify, and for N > 1 the single
allreduce of the count table) over one batch of synthetic input: BASELINE config C2, 100 000 files
x 4 KiB per GPU (weak scaling; rank r holds logical files r, r+N, ... of one N*100k corpus).
`value`  = source MB/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
`e2e`    = same metric through the host C-ABI call path (pinned host arena -> H2D -> kernels ->
           allreduce -> D2H of per-file records and count tables) inside the timed region.
The reference ships no scanner, so the reference arm times the repo's own plain-C
restatement (oracle/, kind "port") on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))

import numpy as np  # noqa: E402

SEED_C2 = 0x7053454D0002
FILES_PER_GPU = 100000
FILE_SIZE = 4096
N_GROUPS = 9
METRIC = "source MB/s scanned (tokenise + line-hash + classify + aggregate)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files-per-gpu", type=int, default=FILES_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config(args, n):
    return {"workload": "C2: %d synthetic files x %d B per GPU (seed 0x%X, SURVEY.md section 8d), "
                        "tokenise+hash+classify+aggregate" % (args.files_per_gpu, FILE_SIZE, SEED_C2),
            "files_per_gpu": args.files_per_gpu, "file_bytes": FILE_SIZE, "n_groups": N_GROUPS,
            "global_files": args.files_per_gpu * n, "sharding": "round-robin by file index, no data-path collective",
            "collective": "one allreduce(SUM) of the int64 [n_groups+1][128]+4 count table per step" if n > 1 else "none",
            "l2": "input per GPU (%.0f MB) exceeds the 126 MB L2, no explicit flush" % (args.files_per_gpu * FILE_SIZE / 1e6)}


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_scan_rate(corpus, threads):
    """Oracle (oracle/liborc.so, plain C, kind 'port') over the corpus with `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc  # the one place bench.py executes oracle/: as the measured CPU baseline
    from concurrent.futures import ThreadPoolExecutor
    n = corpus.n_files
    bounds = [n * t // threads for t in range(threads + 1)]

    def part(t):
        a, b = bounds[t], bounds[t + 1]
        if a == b:
            return 0
        off = corpus.off[a:b + 1]
        res = orc.scan(corpus.arena, off, corpus.len[a:b], corpus.ext[a:b], corpus.grp[a:b], corpus.n_groups, events=False)
        return int(res["global_counts"].sum())
    orc.lib()
    t0 = time.perf_counter()
    if threads == 1:
        tot = part(0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            tot = sum(ex.map(part, range(threads)))
    dt = time.perf_counter() - t0
    return corpus.source_bytes / dt / 1e6, n / dt, dt, tot


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tosemscan as ts
    cores = os.cpu_count() or 1
    n = args.files_per_gpu if cores >= 8 else min(args.files_per_gpu, 25000)
    corpus = ts.gen_corpus(SEED_C2, n, 0, FILE_SIZE, first_index=0, index_stride=args.gpus, n_groups=N_GROUPS, pinned=False)
    for _ in range(min(args.warmup, 1)):
        cpu_scan_rate(corpus, cores)
    rates, dts = [], []
    for _ in range(args.steps):
        mb, fps, dt, _ = cpu_scan_rate(corpus, cores)
        rates.append(mb)
        dts.append(dt)
    v = corpus.source_bytes * len(dts) / sum(dts) / 1e6
    sample = "%d of %d files x %d B per step, %d host threads, arena in RAM" % (n, args.files_per_gpu, FILE_SIZE, cores)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config(args, args.gpus),
           "files_per_s": n * len(dts) / sum(dts),
           "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference package ships no scanner; this is the repo's plain-C restatement (oracle/)"}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import tosemscan as ts
    n = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != n:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N > 1)" % (n, world))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if n > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there under NCCL_DEBUG=VERSION,
        # so stdout is pointed at stderr while the communicator comes up
        os.environ["NCCL_DEBUG"] = "WARN"
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    nf = args.files_per_gpu
    corpus = ts.gen_corpus(SEED_C2, nf, 0, FILE_SIZE, first_index=rank, index_stride=n, n_groups=N_GROUPS, pinned=True)
    sc = ts.Scanner(device=local, max_arena_bytes=int(corpus.off[-1]) + 4096, max_files=nf, max_groups=16)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    sc.upload(corpus, sp)
    sc.scan_resident(0, sp)
    ptr, n64 = sc.device_counts()
    # the count table as a torch tensor over the library's device buffer (for the one allreduce)
    counts = torch.empty(0)
    if n > 1:
        import ctypes

        class _Arr:   # __cuda_array_interface__ view, no copy
            def __init__(self, p, m):
                self.__cuda_array_interface__ = {"shape": (m,), "typestr": "<i8", "data": (p, False), "version": 3}
        counts = torch.as_tensor(_Arr(ptr, n64), device=torch.device("cuda", local))

    def step():
        sc.scan_resident(0, sp)
        if n > 1:
            dist.all_reduce(counts)

    def fence():
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    fence()
    sc.kernel_ms_stats(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    fence()
    ms = e0.elapsed_time(e1)
    sums, nscan = sc.kernel_ms_stats(reset=True)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = sc.last_launch_count() * args.steps
    # sanity: the resident result is the corpus' own (guards against timing a no-op)
    res = sc.download(0, sp)
    glob_assert = int(res["global_counts"].sum())
    assert res["totals"][0] > 0 and glob_assert > 0
    # ---- e2e through the host C-ABI path
    ke = args.e2e_steps or min(args.steps, 10)
    h2d = int(corpus.off[-1]) + 4 * (nf + 1) + 4 * nf + nf + 2 * nf
    d2h = 24 * nf + 8 * (N_GROUPS + 1) * 128 + 64

    def e2e_step():
        # tsm_scan: index H2D, arena H2D in 32 MiB slabs overlapped with the scan of earlier slabs,
        # classify/aggregate, D2H of the per-file records and count tables
        out = sc.scan(corpus, 0, sp)
        if n > 1:
            dist.all_reduce(counts)
            out["global_counts_all_ranks"] = counts.cpu()
        return out
    e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    fence()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    if rank == 0:
        src = corpus.source_bytes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_src = None, None
        try:   # one `ncu --set full` capture of this kernel on this workload (tools/ncu_summary.py --traffic)
            tj = json.load(open(os.path.join(ROOT, "profiles", "k_scan_traffic.json")))["k_scan"]
            if nf == FILES_PER_GPU:
                traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        scan_ms = sums[1] / max(nscan, 1)
        achieved = corpus.algorithmic_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        out = {"metric": METRIC, "value": src * n * args.steps / (ms * 1e-3) / 1e6, "unit": "MB/s", "n_gpus": n,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config(args, n),
               "files_per_s": nf * n * args.steps / (ms * 1e-3),
               "roofline": {"bound": "hbm", "kernel": "k_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)",
                            "algorithmic_bytes_per_launch": corpus.algorithmic_bytes,
                            "kernel_ms": {"k_plan": sums[0] / max(nscan, 1), "k_scan": scan_ms,
                                          "k_classify": sums[2] / max(nscan, 1)},
                            "scans_timed": nscan},
               "e2e": {"value": src * n * ke / e2e_s / 1e6, "unit": "MB/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": ke, "ms_per_step": 1e3 * e2e_s / ke,
                       "path": "tsm_scan(pinned host arena): slab-pipelined H2D + kernels + D2H, then the allreduce"},
               "gpu_launches": launches, "clocks": clocks,
               "check": {"lines": int(res["totals"][0]), "assertion_lines": int(res["totals"][1]), "classified": glob_assert}}
        if n == 1 and not args.no_cpu_baseline:
            mb, fps, dt, _ = cpu_scan_rate(corpus, 1)
            out["cpu_baseline"] = {"value": mb, "unit": "MB/s", "cores": 1, "kind": "port", "files_per_s": fps,
                                   "sample": "all %d files x %d B once (%.1f s), oracle/liborc.so -O2, 1 thread, arena in RAM"
                                             % (nf, FILE_SIZE, dt)}
        print(json.dumps(out))
    sc.close()
    if n > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
