#!/usr/bin/env python
"""bench.py -- throughput of the CloudBrush overlap-discovery + transitive-reduction hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config cfg2] [--genome G]

A "step" is one pass of the whole hot path (SFA text -> checkpoint-B edge list) over one batch of
synthetic reads.  At N=1 the workload is BASELINE.json configs[1]
("synthetic E. coli-size (4.6 Mb) random genome, 36 bp reads at 100x, k=21").

  value   reads/s with the ASCII SFA bytes already resident in HBM when the timed region starts
          (device time, CUDA events on the library's stream: preprocess + overlap + string groups)
  e2e     reads/s through the public C-ABI call sequence with HOST buffers: pinned-host SFA ->
          cb_load_sfa_buffer (H2D) -> cb_preprocess/cb_build_overlap/cb_string_graph ->
          cb_export_edges(checkpoint B) + cb_export_nodes (D2H), wall clock
  roofline  the dominant kernel (rs_onesweep of the k-mer tuple sort): algorithmic bytes per launch
          (24 B per tuple: 12 B read + 12 B written) / its mean CUDA-event duration, against the
          measured HBM copy bandwidth of MEASURED_PEAKS.json; roofline_path: SURVEY.md 8(d) B_alg / step
  cpu_baseline  the CPU restatement oracle ("port", 1 core) on a bounded sample of the same workload

`--impl reference` times the reference's CPU algorithm (the oracle port -- the reference itself is
Java/Hadoop and cannot run in this image) on all host cores over bounded samples.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_sec"
UNIT = "reads/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread every ~2 ms
    (the timed region of the default run is ~0.1-0.5 s); `nvidia-smi -lms` as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.sm, self.reasons = [], set()
        self.nvml = None
        self._stop = threading.Event()
        self.mx = None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except Exception:
                pass
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def make_workload(name, genome=None, seed_shift=0, rank=0, world=1, weak=True):
    """SFA shard of `rank`.  weak: the genome grows with the GPU count, reads per GPU stay fixed (the headline
    config); not weak: the named genome is split over the ranks (config 4 is DEFINED on eight GPUs)."""
    from cloudbrush_b200 import synth
    G, L, cov, both, k, seed = synth.CONFIGS[name]
    if genome:
        G = genome
    if weak:
        G *= world
    n_total = int(round(G * cov / L))
    per = synth.N_PARTS // world
    codes = synth.synth_reads(G, L, cov, both, seed + seed_shift, parts=list(range(rank * per, (rank + 1) * per)))
    first_id = 1 + sum(synth.part_sizes(n_total)[:rank * per])
    sfa = synth.sfa_bytes(codes, first_id=first_id)
    return sfa, dict(genome=G, readlen=L, coverage=cov, both_strands=both, k=k, n_reads=int(codes.shape[0]),
                     n_reads_total=n_total)


def cpu_sample_genome(meta, target_reads=250_000):
    g = int(target_reads * meta["readlen"] / meta["coverage"])
    return max(10_000, min(g, meta["genome"]))


def run_oracle_once(sfa_bytes_, k, L):
    from oracle import oracle as O
    t = time.perf_counter()
    r = O.OracleRun(sfa_bytes_, k, L)
    dt = time.perf_counter() - t
    if r.error:
        raise RuntimeError(r.error)
    ea = r.counter("GenReverseEdge:nodes")  # forces nothing; edges counted below
    n_edges_a = len(r.edges(O.CK_A))
    r.close()
    return dt, n_edges_a


# --------------------------------------------------------------------------------------------
# reference arm: the CPU algorithm on all host cores, bounded samples
# --------------------------------------------------------------------------------------------
def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    name = args.config
    from cloudbrush_b200 import synth
    G, L, cov, both, k, seed = synth.CONFIGS[name]
    meta = dict(genome=args.genome or G, readlen=L, coverage=cov, both_strands=both, k=k)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.ref_threads or cores, 32))
    # every step is a bounded sample sized so that the whole (warm-up + steps) run ends within a few minutes:
    # the port does about 14 k reads/s per thread on this shape
    per_step_s = 170.0 / max(1, args.steps + args.warmup)
    sample_reads = args.ref_sample_reads or int(min(250_000, max(20_000, per_step_s * 14_000)))
    sample_g = cpu_sample_genome(meta, target_reads=sample_reads)
    samples = []
    for t in range(threads):  # one independent sample (own genome seed) per host thread
        sfa, m = make_workload(name, genome=sample_g, seed_shift=1000 + t)
        samples.append((bytes(sfa), m))
    n_reads = sum(m["n_reads"] for _, m in samples)

    def step():
        res = [None] * threads

        def work(i):
            res[i] = run_oracle_once(samples[i][0], k, L)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0, sum(r[1] for r in res)

    for _ in range(args.warmup):
        step()
    times, edges = [], 0
    for _ in range(args.steps):
        dt, e = step()
        times.append(dt)
        edges = e
    t = float(np.mean(times))
    value = n_reads / t
    sample_desc = (f"{threads} independent samples (one per host thread) of {samples[0][1]['n_reads']} reads each: "
                   f"{sample_g} bp random genome, {L} bp reads at {cov}x, k={k}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{name}: {meta['genome']} bp random genome, {L} bp reads at {cov}x, k={k} "
                               f"(timed on bounded samples)", "sample": sample_desc},
        "edges_per_sec": edges / t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

PARITY_COUNTERS = ["reads_good", "nodes", "contained", "tuples", "edges_overlap", "edges_chl2", "edges_re", "trans_edge",
                   "edge_removal_1", "edge_removal_2"]


def checkpoint_digest(g):
    """(rows of the three checkpoints, node table, counters) of a finished context"""
    return dict(a=g.edge_rows(1), c=g.edge_rows(2), b=g.edge_rows(3), nodes=g.nodes(),
                counters=[g.counter(x) for x in PARITY_COUNTERS])



def path_alg_bytes(counters, L, k, world):
    """SURVEY.md 8(d) B_alg per GPU from the run's measured counts."""
    N_in, N, T = counters["reads_good"] / world, counters["nodes"] / world, counters["tuples"] / world
    C, E_A = counters["candidates_verified"] / world, counters["edges_overlap"] / world
    E_c, E_B = counters["edges_chl2"] / world, counters["edges_re"] / world
    A_, R_, P_, Pn_ = L + 10, 4 * ((2 * L + 31) // 32), (2 * k + 7) // 8, 4
    return (N_in * (A_ + R_) + N_in * (2 * R_ + 24 + 24 * P_) + N * R_ + 12 * T + T * (8 + 24 * P_) + 12 * T + 12 * C
            + C * (12 + 2 * R_) + 12 * E_A + 24 * E_A + E_A * (8 + 24 * Pn_) + E_A * (12 + R_) + E_c * (12 + R_)
            + 12 * E_B)


def time_other_config(name, steps, warmup, rank, world, local_rank, transport, peak):
    """One more BASELINE config next to the headline: device-resident step time, e2e with host
    buffers, reads/s, edges/s and the whole-path roofline fraction.  Same rules as the main measurement."""
    import torch
    import torch.distributed as dist
    from cloudbrush_b200 import BrushConfig, BrushGpu
    from cloudbrush_b200 import _lib as L_
    sfa_np, meta = make_workload(name, rank=rank, world=world, weak=False)  # the config's own genome, split over the ranks
    k, L = meta["k"], meta["readlen"]
    host = torch.from_numpy(sfa_np).pin_memory()
    dev = host.cuda()
    g = BrushGpu(BrushConfig(k=k, readlen=L, device=local_rank, rank=rank, world=world))
    if world > 1:
        if transport == "nccl":
            g.init_nccl()
        else:
            from cloudbrush_b200.dist import TorchExchange
            g.set_exchange(TorchExchange(device=torch.device("cuda", local_rank)))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        g.borrow_sfa_device(dev.data_ptr(), dev.numel())
        g.run()
        return g.time_ms("preprocess") + g.time_ms("overlap") + g.time_ms("string")

    for _ in range(warmup):
        step()
    sync()
    ms = float(np.mean([step() for _ in range(steps)]))
    sync()
    counters = {c: g.counter(c) for c in ["reads_good", "nodes", "tuples", "candidates_verified", "edges_overlap",
                                          "edges_chl2", "edges_re", "bucket_path"]}
    stage_ms = {t: g.time_ms(t) for t in ["parse", "dedupe", "keygen", "sort_tuples", "join_verify", "cut", "tr"]}
    e2e = []
    for i in range(3):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        g.load_sfa_buffer(host)
        g.run()
        g.edges(L_.CK_RE)
        g.nodes()
        if i:
            e2e.append(time.perf_counter() - t1)
    e2e_ms = float(np.mean(e2e)) * 1e3
    n_reads = float(meta["n_reads"])
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        tot = torch.tensor([n_reads], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        n_reads = float(tot[0])
    g.close()
    del dev, host
    torch.cuda.empty_cache()
    b_alg = path_alg_bytes(counters, L, k, world)
    return {"workload": f"{name}: {meta['genome']} bp random genome, {L} bp reads at {meta['coverage']}x"
                        f"{' both strands' if meta['both_strands'] else ''}, k={k}, {meta['n_reads']} reads per GPU",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "reads_per_sec": n_reads / (ms * 1e-3),
            "edges_per_sec": counters["edges_overlap"] / (ms * 1e-3), "e2e_ms_per_step": e2e_ms,
            "e2e_reads_per_sec": n_reads / (e2e_ms * 1e-3),
            "roofline_path": {"alg_bytes_per_gpu": b_alg, "achieved": b_alg / (ms * 1e-3) / 1e9, "unit": "GB/s",
                              "frac_of_measured_peak": b_alg / (ms * 1e-3) / 1e9 / peak},
            "counters": counters, "stage_ms": stage_ms}


def sharded_parity(name, world, rank, local_rank, transport, genome):
    """N > 1, before the timed loop: a reduced genome of the same shape runs sharded over the SAME transport the
    timed run uses (native NCCL by default) and, on rank 0, on one context; the union of the shards' checkpoints,
    the node table and the counters must be identical.  Returns the "parity" object of the JSON line (rank 0)."""
    import torch
    import torch.distributed as dist
    from cloudbrush_b200 import BrushConfig, BrushGpu, synth
    from cloudbrush_b200.dist import TorchExchange, shard_lines
    buf, m = synth.make_config(name, scale_genome=genome)
    sfa = bytes(buf)
    g = BrushGpu(BrushConfig(k=m["k"], readlen=m["readlen"], device=local_rank, rank=rank, world=world))
    if transport == "nccl":
        g.init_nccl()
    else:
        g.set_exchange(TorchExchange(device=torch.device("cuda", local_rank)))
    g.load_sfa_buffer(shard_lines(sfa, rank, world))
    g.run()
    mine = checkpoint_digest(g)
    g.close()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    res = None
    if rank == 0:
        s1 = BrushGpu(BrushConfig(k=m["k"], readlen=m["readlen"], device=local_rank))
        s1.load_sfa_buffer(sfa)
        s1.run()
        one = checkpoint_digest(s1)
        s1.close()
        bad = []
        for key in ("a", "c", "b"):
            got = np.concatenate([q[key] for q in parts])
            got = got[np.lexsort((got[:, 3], got[:, 2], got[:, 1], got[:, 0]))]
            if not np.array_equal(got, one[key]):
                bad.append(f"checkpoint {key}: {len(got)} vs {len(one[key])} edges")
        if not np.array_equal(np.concatenate([q["nodes"] for q in parts]), one["nodes"]):
            bad.append("node table")
        for r_, q in enumerate(parts):
            if q["counters"] != one["counters"]:
                bad.append(f"counters of rank {r_}")
        res = {"world": world, "ok": not bad, "transport": transport, "against": "one context on rank 0",
               "input": f"{name} shape, {genome} bp genome, {m['n_reads_total']} reads",
               "edges_overlap": int(len(one["a"])), "edges_re": int(len(one["b"])), "nodes": int(len(one["nodes"])),
               "differences": bad}
    flag = [res["ok"] if rank == 0 else None]
    dist.broadcast_object_list(flag, src=0)
    if not flag[0]:
        if rank == 0:
            log("PARITY FAILURE (sharded run differs from one context): " + "; ".join(res["differences"]))
        raise SystemExit(3)
    return res


def oracle_parity(sfa_bytes_, k, L, device):
    """N = 1: the sample the cpu_baseline leg times on the oracle also runs through the C ABI on the GPU, and the
    node set, the three checkpoints and the driver-visible counters are compared."""
    from cloudbrush_b200 import BrushConfig, BrushGpu
    from oracle import oracle as O
    t = time.perf_counter()
    r = O.OracleRun(sfa_bytes_, k, L)
    dt = time.perf_counter() - t
    if r.error:
        raise RuntimeError(r.error)
    g = BrushGpu(BrushConfig(k=k, readlen=L, device=device))
    g.load_sfa_buffer(sfa_bytes_)
    g.run()
    bad = []
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    gn = g.nodes()
    if len(gn) != len(line) or not np.array_equal(gn["line"].astype(np.int64), line) or \
            not np.array_equal(gn["cov"].astype(np.int64), cov.astype(np.int64)):
        bad.append("node table")
    n_edges = {}
    for nm, ock, gck in (("A", O.CK_A, 1), ("chl2", O.CK_CHL2, 2), ("B", O.CK_B, 3)):
        oe, ge = r.edges(ock), g.edge_rows(gck)
        n_edges[nm] = int(len(oe))
        if oe.shape != ge.shape or not np.array_equal(oe, ge):
            bad.append(f"checkpoint {nm}: oracle {len(oe)} gpu {len(ge)}")
    for c_ in ("reads_good", "nodes", "contained", "edge_removal_1", "edge_removal_2", "trans_edge"):
        if max(r.counter(c_), 0) != g.counter(c_):
            bad.append(f"counter {c_}")
    g.close()
    r.close()
    return dt, n_edges["A"], {"world": 1, "ok": not bad, "against": "CPU oracle (oracle/brush_oracle.cpp)",
                              "edges_overlap": n_edges["A"], "edges_re": n_edges["B"], "differences": bad}

def bind_to_gpu_numa(gpu_index):
    """Binds this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers it
    allocates from here on (first touch) sit on the GPU's NUMA node: with eight ranks loading 575 MB each from
    wherever the launcher happened to start them, the host-to-device copy took 23.5 ms instead of 10 (VERDICT r01
    item 9).  Returns the number of CPUs bound to, or None when NVML or the affinity call is not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = gpu_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[gpu_index])
            except Exception:
                idx = gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = ((os.cpu_count() or 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


KERNELS = ["parse_lines", "dedupe_runs", "keygen", "join_verify", "cut", "tr",
           "rs_scatter:dedupe", "rs_upsweep:dedupe", "rs_scatter:tuples", "rs_upsweep:tuples",
           "group_walk", "side_insert", "consistency", "tr_final",
           "ph_hist1", "ph_scatter1", "ph_hist2", "ph_scatter2", "bucket_join",
           "p2p_put", "rs_scatter:partition", "rs_scatter:route"]


def kernel_rooflines(stats, steps, ms, counters, L, k, world, peak, peak_src):
    """Achieved algorithmic GB/s of the main kernels (DESIGN.md section 5 gives the bytes per unit), sorted by
    their share of the step.  The first entry is the dominant kernel (largest ms per step)."""
    T = counters["tuples"] / world
    N = counters["nodes"] / world
    C = counters["candidates_verified"] / world
    R = 4 * ((2 * L + 31) // 32)          # packed read, SURVEY.md 8(d)
    nwo = (L - k + 15) // 16              # 32-bit words of a stored overhang
    slots = counters.get("slots", 0) / world or 2.0 * T
    T_slots = N * (L - k + 1)             # tuple slots keygen writes (windows that emit nothing included)
    alg = {
        "bucket_join": ("group + join + verify + mirror of one hash bucket per CTA", T * (12 + R) + slots * 8 + C * 4 * nwo),
        "ph_scatter1": ("hash partition pass 1 (atomic-reserve scatter)", 24.0 * T),
        "ph_scatter2": ("hash partition pass 2 (atomic-reserve scatter)", 24.0 * T),
        "rs_scatter:tuples": ("one 8/9-bit LSD pass of the flat tuple sort", 24.0 * T_slots),
        "join_verify": ("flat join + verify + mirror", T * (12 + R + 8) + C * (8 + 4 * nwo)),
        "parse_lines": ("SFA line parser / 2-bit packer", counters["reads_good"] / world * (L + 10 + R + 16 + 12 + 1)),
        "consistency": ("per-list overhang nesting test", C * (8 + 4 * nwo)),
        "tr_final": ("TransitiveReduction + EdgeRemoval over the slots", slots * 10),
    }
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    out = []
    for name, (what, b) in alg.items():
        st = stats.get(name)
        if not st or st[1] <= 0:
            continue
        per_launch_ms = st[0] / st[1]
        ach = b / (per_launch_ms * 1e-3) / 1e9
        out.append({"bound": "hbm", "kernel": f"{name} ({what})", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic.get(name, {}).get("dram_bytes_per_launch"),
                    "peak_source": peak_src, "alg_bytes_per_launch": b, "ms_per_launch": per_launch_ms,
                    "launches_per_step": st[1] / steps, "ms_per_step": st[0] / steps, "share_of_step": st[0] / steps / ms})
    out.sort(key=lambda r: -r["ms_per_step"])
    return out


def bench_ours(args):
    import torch
    import torch.distributed as dist
    from cloudbrush_b200 import BrushConfig, BrushGpu
    from cloudbrush_b200 import _lib as L_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bound_cpus = None if os.environ.get("CB_NO_NUMA_BIND") else bind_to_gpu_numa(local_rank)
    log(f"[rank {rank}] host affinity: " + (f"{bound_cpus} CPUs local to GPU {local_rank} (NVML)" if bound_cpus else "unchanged"))
    if world > 1:
        # NCCL reads these when it is first initialised in the process (see cloudbrush_b200/csrc/dist.cu)
        for k_, v_ in (("NCCL_MIN_P2P_NCHANNELS", "64"), ("NCCL_MAX_P2P_NCHANNELS", "64"), ("NCCL_MAX_NCHANNELS", "64")):
            os.environ.setdefault(k_, v_)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.config
    t0 = time.time()
    sfa_np, meta = make_workload(name, genome=args.genome, rank=rank, world=world)
    log(f"[rank {rank}] workload {name}: {meta} ({len(sfa_np)/1e6:.1f} MB SFA) generated in {time.time()-t0:.1f}s")
    k, L = meta["k"], meta["readlen"]
    n_reads = meta["n_reads"]

    host = torch.from_numpy(sfa_np).pin_memory()
    dev = host.cuda(non_blocking=False)
    g = BrushGpu(BrushConfig(k=k, readlen=L, device=local_rank, rank=rank, world=world))
    if world > 1:
        if args.transport == "nccl":  # the library drives NCCL itself (ncclSend/ncclRecv on its own stream)
            g.init_nccl()
        else:  # the cb_exchange callbacks implemented with torch.distributed (cloudbrush_b200/dist.py)
            from cloudbrush_b200.dist import TorchExchange
            exch = TorchExchange(device=torch.device("cuda", local_rank))
            g.set_exchange(exch)

    parity = None
    if world > 1 and not args.no_parity:
        parity = sharded_parity(name, world, rank, local_rank, args.transport, args.parity_genome)

    def step_resident():
        g.borrow_sfa_device(dev.data_ptr(), dev.numel())
        g.run()
        return g.time_ms("preprocess") + g.time_ms("overlap") + g.time_ms("string")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    g.set_profile(True)
    l0 = g.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    w0 = time.perf_counter()
    dev_ms = []
    for _ in range(args.steps):
        dev_ms.append(step_resident())
    barrier()
    wall = (time.perf_counter() - w0) / args.steps
    clocks = sampler.stop()
    launches = g.launch_count() - l0
    stats = {kn: g.kernel_stat(kn) for kn in KERNELS}
    g.set_profile(False)
    ms = float(np.mean(dev_ms))
    counters = {c: g.counter(c) for c in ["reads_good", "nodes", "tuples", "candidates_verified", "edges_overlap",
                                          "edges_chl2", "edges_re", "trans_edge", "edge_removal_1", "slow_lists",
                                          "bucket_path", "slots", "host_collectives"]}
    if world > 1 and args.transport == "nccl":
        counters["p2p_exchange"] = g.counter("p2p_exchange")
    stage_ms = {t: g.time_ms(t) for t in ["parse", "dedupe", "keygen", "sort_tuples", "join_verify",
                                          "cut", "tr"] + (["exchange_tuples", "dd_route", "dd_return", "node_gather",
                                                           "side_route", "tr_votes", "a2a_tuples"] if world > 1 else [])}

    # ---- e2e through the C ABI with host buffers ----
    from cloudbrush_b200.api import EDGE_DTYPE, NODE_DTYPE
    # the caller's result buffers are pinned host memory (sized from the resident run)
    pin_e = torch.empty((g.num_edges(L_.CK_RE) + 1024) * EDGE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    pin_n = torch.empty((counters["nodes"] + 1024) * NODE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out_e = pin_e.numpy().view(EDGE_DTYPE)
    out_n = pin_n.numpy().view(NODE_DTYPE)
    e2e_t, e2e_parts = [], []
    d2h = 0
    # (1) single shot: load (H2D) -> three stage groups -> export (D2H), nothing overlapped
    for i in range(max(1, min(args.steps, 3)) + 1):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        g.load_sfa_buffer(host)
        t2 = time.perf_counter()
        g.run()
        t3 = time.perf_counter()
        eb = g.edges(L_.CK_RE, out=out_e)
        nd = g.nodes(out=out_n)
        t4 = time.perf_counter()
        d2h = eb.nbytes + nd.nbytes
        if i > 0:  # first one warms the export path
            e2e_parts.append((t2 - t1, t3 - t2, t4 - t3, t4 - t1))
    # (2) steady state of a caller that processes batch after batch (cb_prefetch_sfa_buffer): the H2D copy of the
    # NEXT batch runs on the context's copy stream behind the overlap / string-graph stage groups and the export
    # of the current one.  Every timed iteration still contains one full H2D and one D2H.
    # The results leave the same way (cb_export_nodes_begin / cb_export_edges_begin / cb_export_wait): the nodes'
    # D2H copy runs behind the overlap stage group, the edges' behind the NEXT batch's preprocess (two host
    # buffers in turn); the caller waits for batch i-1's edges after batch i's preprocess.  The loop is timed as a
    # whole (a device-wide synchronize between iterations would serialise the copies again): n_it batches in,
    # n_it complete results out, the last export waited for inside the timed region.
    pin_e2 = torch.empty_like(pin_e).pin_memory()
    out_e2 = [out_e, pin_e2.numpy().view(EDGE_DTYPE)]
    n_it = max(4, min(args.steps, 10))
    # Two inputs are prefetched ahead, so the H2D copies run back to back (also during cb_preprocess).
    g.prefetch_sfa_buffer(host)
    for timed in (False, True):
        barrier()
        t1 = time.perf_counter()
        for i in range(n_it if timed else 3):
            g.prefetch_sfa_buffer(host)  # batch i+1; batch i was prefetched one iteration ago
            g.preprocess()
            g.export_wait()  # the previous batch's edges are in host memory now
            nd = g.nodes_begin(out_n)
            g.build_overlap()
            g.string_graph()
            eb = g.edges_begin(L_.CK_RE, out_e2[i & 1])
        g.export_wait()
        torch.cuda.synchronize()  # every H2D started inside the timed region has also finished inside it
        t4 = time.perf_counter()
        if timed:
            e2e_t.append((t4 - t1) / n_it)
    assert len(eb) == g.num_edges(L_.CK_RE)
    e2e_s = float(np.mean(e2e_t))
    e2e_split = [float(np.mean([p[k] for p in e2e_parts])) * 1e3 for k in range(4)]

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
        tot = torch.tensor([n_reads], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_reads, total_edges = float(tot[0]), float(counters["edges_overlap"])  # counters are global already
    else:
        total_reads, total_edges = float(n_reads), float(counters["edges_overlap"])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the other BASELINE configs this GPU count is quoted on: config 3 on one GPU, config 4 on eight
    other = {}
    if not args.no_other_configs and name == "cfg2" and not args.genome:
        for oname in ({1: ["cfg3"], 8: ["cfg4"]}.get(world, [])):
            if g is not None:  # free the headline workload's device memory first
                g.close()
                g = None
                dev = None
                torch.cuda.empty_cache()
            try:
                other[oname] = time_other_config(oname, 3, 2, rank, world, local_rank, args.transport, peak)
            except Exception as e:  # the headline line must not be lost to a side measurement
                other[oname] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # the dominant kernel (largest ms per step) first, the other main kernels after it
        roofs = kernel_rooflines(stats, args.steps, ms, counters, L, k, world, peak, peak_src)
        roof = roofs[0] if roofs else None
        kern = {kn: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                for kn, v in stats.items() if v}
        # ---- whole-path figure of BASELINE.md / SURVEY.md 8(d): B_alg from the run's measured counts ----
        # (the byte budget of a sort-based pipeline; this build replaces its edge sort by slot writes and
        # still reports against it.  C = verified candidates only: rejected ones are never materialised.)
        b_alg = path_alg_bytes(counters, L, k, world)
        path = {"alg_bytes_per_gpu": b_alg, "achieved": b_alg / (ms * 1e-3) / 1e9, "unit": "GB/s",
                "frac_of_measured_peak": b_alg / (ms * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8TBs": b_alg / (ms * 1e-3) / 1e9 / 8000.0,
                "formula": "SURVEY.md 8(d) B_alg (pack+dedupe+keygen+sort+join+verify+mirror+cut+tr), per GPU"}
        # ---- CPU baseline (oracle port, 1 core, bounded sample) ----
        cpu = None
        if not args.no_cpu_baseline:
            sg = cpu_sample_genome(meta, target_reads=args.cpu_sample_reads)
            ssfa, sm = make_workload(name, genome=sg, seed_shift=1000)
            if world == 1 and not args.no_parity:
                dt, _, parity = oracle_parity(bytes(ssfa), k, L, local_rank)
                if not parity["ok"]:
                    log("PARITY FAILURE (CUDA path differs from the oracle): " + "; ".join(parity["differences"]))
                    raise SystemExit(3)
                parity["input"] = f"{sm['n_reads']} reads, {sg} bp genome (the cpu_baseline sample)"
            else:
                dt, _ = run_oracle_once(bytes(ssfa), k, L)
            cpu = {"value": sm["n_reads"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{sm['n_reads']} reads: {sg} bp random genome, {L} bp reads at {meta['coverage']}x, "
                             f"k={k} (oracle/brush_oracle.cpp, single thread, {dt:.1f}s)"}
        line = {
            "metric": METRIC, "value": total_reads / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{name}: {meta['genome']} bp random genome, {L} bp reads at "
                                   f"{meta['coverage']}x{' both strands' if meta['both_strands'] else ''}, k={k}, "
                                   f"{n_reads} reads per GPU",
                       "l2": "inputs larger than L2 (SFA text and tuple arrays exceed 126 MB)" if len(sfa_np) > 126e6
                             else "input smaller than L2 (reduced run)",
                       "host_affinity": (f"each rank bound to the {bound_cpus} CPUs NVML reports as local to its GPU "
                                         "(pinned host buffers on that NUMA node)") if bound_cpus else "unchanged",
                       "sharding": (f"weak scaling: {world} shards of consecutive SFA lines, k-mer groups hash-partitioned, "
                                    f"tuple all-to-all over NCCL ({args.transport} transport), packed reads replicated") if world > 1 else "single GPU"},
            "edges_per_sec": total_edges / (ms * 1e-3),
            "wall_ms_per_step": wall * 1e3,
            "e2e": {"value": total_reads / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(len(sfa_np)),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "how": "steady state through the C ABI with host buffers: cb_prefetch_sfa_buffer (H2D of the next batch "
                           "behind this batch's stage groups, two inputs ahead), cb_preprocess, cb_export_nodes_begin (D2H behind "
                           "the overlap stage group), cb_build_overlap, cb_string_graph, cb_export_edges_begin(B) (D2H "
                           "behind the next batch's preprocess), cb_export_wait; pinned host memory, one full H2D and one "
                           "full D2H per step, the loop of batches timed as a whole",
                    "single_shot": {"ms_per_step": e2e_split[3], "ms_h2d_load": e2e_split[0], "ms_stages": e2e_split[1],
                                    "ms_export_d2h": e2e_split[2], "reads_per_sec": n_reads / (e2e_split[3] * 1e-3)}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "roofline_kernels": roofs[1:],
            "roofline_path": path,
            "cpu_baseline": cpu,
            "parity": parity,
            "other_configs": other,
            "counters": counters,
            "stage_ms": stage_ms,
            "kernels": kern,
        }
        emit(line)
    if g is not None:
        g.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process' original stdout; everything else (NCCL banners, library
    chatter written to fd 1) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--genome", type=int, default=None, help="override the genome length (reduced runs)")
    ap.add_argument("--cpu-sample-reads", type=int, default=250_000)
    ap.add_argument("--ref-sample-reads", type=int, default=0, help="reads per host thread and step (0: sized to the run)")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip config 3 (N=1) / config 4 (N=8)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check that precedes the timing")
    ap.add_argument("--parity-genome", type=int, default=200_000, help="genome length of the N>1 parity input")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "torch"],
                    help="N>1: native NCCL inside the library (default) or the torch.distributed callbacks")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
