#!/usr/bin/env python
"""bench.py -- mapped bases/sec of the seed-chain-align path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (stage-1 MCAS waves + stage-2 remap: sketch, seed lookup, anchor sort,
chaining, extension DP with traceback, host glue) over one batch of synthetic reads drawn from the workload's
distribution.  Workload at N=1: BASELINE.json configs[1] -- 250 Mbp uniform-random reference, ONT-like reads
N50 = 20 kb at 5 % error, -x map-ont -W <top-0.02 % k-mers> k=15 -c.  The 100 k reads of the config are sampled in
batches of --reads reads (fresh reads each step, so nothing is cached between steps; each batch's working set --
read pool, backtrack matrices, 2.6 GB index -- is far larger than L2).

The K timed steps are submitted together (K batches): the library's orchestration lanes pull chunks of reads from
the whole submission, so the steps pipeline instead of draining the GPU at every step boundary.  The timed region is
bracketed by a barrier + device synchronisation on both sides; the warm-up has the same shape (W steps together).

  value : bases/s with the raw reads of all K steps already resident in one HBM pool (wm_bench_upload); CUDA events
          bracket the whole pass (ASCII -> 2-bit codes is inside: it is part of the path)
  e2e   : bases/s through wm_gpu_map_batch with host buffers (staging + H2D of the reads, D2H of every result inside)
  roofline : the DP fill kernel, algorithmic bytes over the time during which a fill kernel was running
          (CUDA events around every launch), vs the measured HBM peak
  cpu_baseline : the real reference (oracle/_ref/winnowmap, SSE4.1, all host cores) on a bounded sample (N = 1 only)

--impl reference times the reference binary itself on the same workload (CPU, all host threads).
Under torchrun (N > 1) every rank maps its own batches (reads shard with no data-path collective): weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")  # before anything loads libgomp: idle workers must not spin (see _lib.py)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

CACHE = os.environ.get("WM_BENCH_CACHE", "/tmp/wm_bench_cache")
REF_LEN = int(os.environ.get("WM_BENCH_REF_LEN", 250_000_000))
N50, ERR, K = 20000, 0.05, 15


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def workload(ref_len):
    """Reference FASTA + -W list (cached on disk), returns (ref_path, w_path, contigs)."""
    import gen_data
    os.makedirs(CACHE, exist_ok=True)
    ref = os.path.join(CACHE, f"ref_{ref_len}.fa")
    wf = os.path.join(CACHE, f"rep_{ref_len}_k{K}.txt")
    t0 = time.time()
    rng = np.random.default_rng(1002)
    contigs = gen_data.make_ref(rng, ref_len, 1, False)  # one contig per 250 Mbp, seed 1000 + cfg
    if not os.path.exists(ref):
        gen_data.write_fasta(ref + ".tmp", contigs)
        os.replace(ref + ".tmp", ref)
    if not os.path.exists(wf):
        n, thr = gen_data.write_top_kmers(wf + ".tmp", contigs, K, 0.9998)
        os.replace(wf + ".tmp", wf)
        log(f"-W list: {n} k-mers above count {thr}")
    log(f"workload ready in {time.time() - t0:.1f}s")
    return ref, wf, contigs


def make_batch(contigs, n_reads, seed):
    import gen_data
    rng = np.random.default_rng(seed)
    return gen_data.make_reads(rng, contigs, n_reads, N50, ERR, min_len=1000)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is polled in-process
    every 20 ms (an nvidia-smi subprocess takes longer than a short timed region); nvidia-smi is the fallback."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, uuid, [], set(), False, None
        self.source = None

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid if isinstance(self.uuid, bytes) else self.uuid.encode())
            except Exception:
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        self.source = "nvml"
        while not self.stop_flag:
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            try:
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for nm, bit in self.REASONS.items():
                if r & bit:
                    self.reasons.add(nm)
            time.sleep(0.02)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def run(self):
        try:
            self._nvml()
        except Exception:
            if not self.stop_flag:
                self._smi()

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "source": self.source}


def run_reference(refbin, ref, wf, reads_fa, threads):
    """Mapping-phase wall time of the reference from its own stderr stamps (main.c:401 -> last map.c:1220 line)."""
    cmd = [refbin, "-t", str(threads), "-c", "-x", "map-ont", "-W", wf, ref, reads_fa]
    p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    t_idx, t_last = None, None
    for ln in p.stderr.splitlines():
        if ln.startswith("[M::main::") and "loaded/built the index" in ln:
            t_idx = float(ln.split("::")[2].split("*")[0])
        if ln.startswith("[M::worker_pipeline::") and "mapped" in ln:
            t_last = float(ln.split("::")[2].split("*")[0])
    if p.returncode != 0 or t_idx is None or t_last is None:
        raise RuntimeError("reference run failed: " + p.stderr[-400:])
    return t_last - t_idx, t_idx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=int(os.environ.get("WM_BENCH_READS", 1500)))
    ap.add_argument("--cpu-reads", type=int, default=int(os.environ.get("WM_BENCH_CPU_READS", 3000)))
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    import gen_data
    cores = os.cpu_count() or 1
    refbin = os.path.join(ROOT, "oracle", "_ref", "winnowmap")

    if a.impl == "reference":
        if rank != 0:
            return
        ref, wf, contigs = workload(REF_LEN)
        vals = []
        tmp = tempfile.mkdtemp(prefix="wm_bench_")
        nb = 0
        for s in range(a.warmup + a.steps):
            recs = make_batch(contigs, a.cpu_reads, 7000 + s)
            fa = os.path.join(tmp, f"r{s}.fa")
            gen_data.write_fasta(fa, recs)
            nb = sum(len(x) for _, x in recs)
            dt, _ = run_reference(refbin, ref, wf, fa, cores)
            if s >= a.warmup:
                vals.append((nb, dt))
        tb, tt = sum(v[0] for v in vals), sum(v[1] for v in vals)
        val = tb / tt
        print(json.dumps({"impl": "reference", "metric": "mapped bases/sec", "value": val, "unit": "bases/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1e3 * tt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int8", "data": "synthetic",
                          "config": {"workload": f"{REF_LEN / 1e6:.0f} Mbp random ref, ONT reads N50=20kb 5% err, -x map-ont -W top-0.02% k=15 -c; "
                                                 f"{a.cpu_reads} reads per step (index build excluded)"},
                          "cpu_baseline": {"value": val, "unit": "bases/s", "cores": cores, "kind": "reference",
                                           "sample": f"{a.cpu_reads} reads ({nb / 1e6:.1f} Mbase) per step, winnowmap -t {cores}"},
                          "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from winnowmap_b200 import lib
    from winnowmap_b200.mapper import MapOpt, Mapper, _setup  # noqa: F401
    L = lib()
    L.wm_prof_get.argtypes = [C.POINTER(C.c_double)]
    L.wm_gpu_map_batch.argtypes = [C.c_void_p, C.POINTER(MapOpt), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
    L.wm_bench_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
    L.wm_bench_map_resident.argtypes = [C.c_void_p, C.POINTER(MapOpt), C.c_int, C.POINTER(C.c_double)]
    L.wm_free_regs.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]

    ref, wf, contigs = workload(REF_LEN) if rank == 0 or world == 1 else (None, None, None)
    if dist is not None:
        dist.barrier()
        if rank != 0:
            ref, wf, contigs = workload(REF_LEN)  # cached files; every rank needs the contigs to draw its own reads
    t0 = time.time()
    # one worker per physical core at N=1; with several ranks per node the ranks share the node's hardware threads
    n_thr = max(1, min(int(os.environ.get("WM_HOST_THREADS", 64)), (cores // 2) // max(1, world)))
    if "WM_THREADS_PER_RANK" in os.environ:
        n_thr = max(1, int(os.environ["WM_THREADS_PER_RANK"]))
    if world == 1:
        mp = Mapper(ref, wf, preset="map-ont", cigar=True, device=local, n_threads=n_thr)
    else:
        # one-time index fan-out: rank 0 builds, one NCCL broadcast (GPU to GPU over NVLink), the others adopt the blob
        from winnowmap_b200 import multi
        import torch
        if rank == 0:
            mp = Mapper(ref, wf, preset="map-ont", cigar=True, device=local, n_threads=n_thr)
            blob = mp.index_blob()
        else:
            blob = None
        tb = time.time()
        blob = multi.broadcast_blob(blob, 0, rank, device=torch.device("cuda", local))
        if rank != 0:
            mp = Mapper(None, None, preset="map-ont", cigar=True, device=local, n_threads=n_thr, blob=blob)
        log(f"rank {rank}: index blob {blob.nbytes / 1e6:.0f} MB broadcast+adopted in {time.time() - tb:.2f}s")
    log(f"rank {rank}: index ready in {time.time() - t0:.1f}s ({mp.stats()['n_keys']:.0f} keys)")

    def pack(recs):
        n = len(recs)
        names = (C.c_char_p * n)(*[nm.encode() for nm, _ in recs])
        seqs_b = [s.tobytes() for _, s in recs]
        seqs = (C.c_char_p * n)(*seqs_b)
        lens = (C.c_int32 * n)(*[len(s) for s in seqs_b])
        return n, names, seqs, lens, seqs_b

    def map_host(recs):
        n, names, seqs, lens, keep = pack(recs)
        n_reg = (C.c_int32 * n)(); regs = (C.c_void_p * n)(); rl = (C.c_int32 * n)(); fg = (C.c_int32 * n)()
        t = time.perf_counter()
        L.wm_gpu_map_batch(mp.ctx, C.byref(mp.mo), n, names, seqs, lens, n_reg, regs, rl, fg, n_thr)
        L.wm_device_synchronize()
        dt = time.perf_counter() - t
        d2h = sum(n_reg) * 80
        L.wm_free_regs(n, n_reg, regs)
        return dt, sum(lens), d2h

    def upload(recs):
        n, names, seqs, lens, keep = pack(recs)
        L.wm_bench_upload(mp.ctx, n, names, seqs, lens)  # raw reads -> one HBM pool (not timed)
        L.wm_device_synchronize()
        return sum(lens)

    def map_uploaded():
        ms = C.c_double()
        L.wm_bench_map_resident(mp.ctx, C.byref(mp.mo), n_thr, C.byref(ms))  # CUDA events bracketing the whole pass
        return ms.value / 1e3

    def map_resident(recs):
        nb = upload(recs)
        return map_uploaded(), nb

    def barrier():
        L.wm_device_synchronize()
        if dist is not None:
            dist.barrier()

    seed0 = 9000 + 1000 * rank
    batches = [make_batch(contigs, a.reads, seed0 + s) for s in range(a.warmup + a.steps)]
    # warm-up in the shape of the timed passes: W steps submitted together, device-resident and through the host API
    # at least as many steps as the timed pass, so that every lane has sized its workspaces for the same chunk size
    n_warm = max(a.warmup, a.steps)
    warm = [r for s in range(n_warm) for r in (batches[s] if s < a.warmup else make_batch(contigs, a.reads, seed0 + 300 + s))]
    if warm:
        map_resident(warm)
        map_host(warm)
    L.wm_prof_enable(1); L.wm_prof_reset()
    L.wm_dump_timers() if os.environ.get("WM_TIMING") else None
    phys, uuid = local, None  # NVML numbers the physical devices: honour CUDA_VISIBLE_DEVICES
    vis = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
    if local < len(vis):
        if vis[local].isdigit():
            phys = int(vis[local])
        elif vis[local].startswith("GPU-"):
            uuid = vis[local]
    sampler = ClockSampler(phys, uuid); sampler.start()
    # the K timed steps are submitted together (K batches of reads_per_step reads, all resident in HBM): the orchestration
    # lanes pull chunks of reads from the whole pool, so the steps pipeline instead of draining the GPU at every step end
    timed = [r for s in range(a.warmup, a.warmup + a.steps) for r in batches[s]]
    bases = upload(timed)
    barrier()
    t_steps = map_uploaded()
    barrier()
    clocks = sampler.result()
    if os.environ.get("WM_TIMING"):
        log(f"timers over {a.steps} timed steps:")
        L.wm_dump_timers()
    prof = (C.c_double * 8)(); L.wm_prof_get(prof)
    L.wm_prof_enable(0)
    # end to end through the host-buffer API (fresh batches)
    e2e_recs = [r for s in range(a.steps) for r in make_batch(contigs, a.reads, seed0 + 500 + s)]
    L.wm_prof_reset()  # zeroes the library's host<->device byte counters
    e2e_t, e2e_b, d2h_b = map_host(e2e_recs)  # K steps in one call: host buffers in, alignment records out
    h2d_step, d2h_step = int(e2e_b / a.steps), int(d2h_b / a.steps)
    try:  # what the library actually copied: reads and job tables in; chains, DP results and CIGARs out
        cp = (C.c_double * 2)()
        L.wm_prof_get_copies(cp)
        if cp[0] > 0:
            h2d_step, d2h_step = int(cp[0] / a.steps), int(cp[1] / a.steps)
    except AttributeError:
        pass
    if dist is not None:
        import torch
        t = torch.tensor([t_steps, e2e_t], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        b = torch.tensor([float(bases), float(e2e_b)], device="cuda"); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        t_steps, e2e_t = t.tolist(); bases, e2e_b = b.tolist()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    # lanes launch their fill kernels concurrently: the denominator is the time during which at least one of them ran
    k_ms = prof[7] if prof[7] > 0 else prof[1]
    ach = prof[3] / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    cpu = None
    try:  # the reference beside it (N = 1 only), bounded sample, all host cores
        if world > 1:
            raise RuntimeError("reported by the N=1 run only")
        recs = make_batch(contigs, a.cpu_reads, 777)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "cpu.fa"); gen_data.write_fasta(fa, recs)
            dt, _ = run_reference(refbin, ref, wf, fa, cores)
        nb = sum(len(x) for _, x in recs)
        cpu = {"value": nb / dt, "unit": "bases/s", "cores": cores, "kind": "reference",
               "sample": f"{a.cpu_reads} reads ({nb / 1e6:.1f} Mbase) of the same distribution, winnowmap -t {cores}, mapping phase only"}
    except Exception as e:  # noqa: BLE001
        cpu = {"value": None, "unit": "bases/s", "cores": cores, "kind": "reference", "sample": f"unavailable: {e}"}
    st = mp.stats()
    print(json.dumps({
        "metric": "mapped bases/sec", "value": bases / t_steps, "unit": "bases/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_steps / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"{REF_LEN / 1e6:.0f} Mbp random ref, ONT reads N50=20kb 5% err, -x map-ont -W top-0.02% k=15 -c (BASELINE configs[1]); "
                               f"{a.reads} fresh reads per step per GPU, working set >> L2", "reads_per_step": a.reads, "host_threads": n_thr,
                   "lanes": int(os.environ.get("WM_LANES", max(2, min(8, n_thr // 8))))},
        "e2e": {"value": e2e_b / e2e_t, "unit": "bases/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step},
        "gpu_launches": int(prof[0]),
        "roofline": {"bound": "hbm", "kernel": "wm_extd2_fill_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": 5.196e9, "traffic_of": "dram read+write of one captured launch of 4.3e9 block cells (ncu --set full, "
                     "profiles/r01_ncu_fill_v3_summary.txt): 1.2 B per algorithmic byte", "of": "measured" if peaks else "fallback", "launches": int(prof[2]), "kernel_ms": k_ms, "kernel_ms_sum_over_launches": prof[1],
                     "block_cells_per_s": prof[4] / (k_ms * 1e-3) if k_ms > 0 else 0.0,
                     "jobs": int(prof[5]), "block_cells": prof[4], "frac_cells_in_16x2_path": prof[6] / prof[4] if prof[4] > 0 else 0.0},
        "cpu_baseline": cpu, "clocks": clocks,
        "breakdown_s": {"seed_chain": st["t_seed"], "dp_rounds": st["t_dp"], "host_glue": st["t_host"]},
    }))


if __name__ == "__main__":
    main()
