#!/usr/bin/env python
"""bench.py -- ICP scans/sec on BASELINE.json's headline configuration (cfg2 replayed as cfg4 shards).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port + verbatim octree)

A "step" registers one batch of B independent OS1-128 scans (131 072 points each, distinct seeds) against the
1 M-point warehouse map with up to 20 ICP iterations each -- exactly what LidarSLAM::Localization does per scan.
`value` = scans/s with the scans already resident in HBM (so_register_batch_device); `e2e` = the same through
so_register_batch with HOST (pinned) scan buffers, H2D of the scans and D2H of the results inside the timed region.
Ranks shard the scan list (no data-path collective); the step ends with the NCCL all-gather of the 7-double poses.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ICP scans/sec (128-beam, 1M-pt map)"
UNIT = "scans/s"
WORKLOAD = "cfg2: OS1-128 synthetic scans (131072 pts) vs 1M-pt local map, 20 ICP iters, planeRes 0.2, all points active"
KNN_BYTES_PER_POINT = 16 + 21 + 80              # unfused build: scan float4 read + 5 positions + flag + 5 neighbour float4 written by k_knn_scan
MATCH_BYTES_PER_POINT = 16 + 25 + 45            # fused k_knn_fit: scan float4 read; 5 positions + d5 + flag; correspondence {n,d} 32 + w 8 + flags 4 + status 1
NCU_DRAM_BYTES_PER_POINT = {"k_knn_scan": 75.1, "k_knn_fit": 63.4}    # profiles/ncu_prof_r1p_metrics.csv: (26.88 MB read + 51.89 MB written) / 1 048 576 points of one
                                                                       # k_knn_scan launch; fused build (gpurun capture r1m): (28.07 + 38.46) MB


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with NVML while the timed region runs.  One sampler per job (rank 0) polls every
    GPU of the job: NVML is initialised BEFORE the warm-up (nvmlInit enumerates all GPUs and holds driver locks for a long
    time -- inside a 40 ms timed region, in N processes at once, it stalled kernel launches), then polls at a low rate."""

    def __init__(self, indices, period: float = 0.02):
        super().__init__(daemon=True)
        self.indices, self.period, self.samples, self.reasons, self.stop_flag, self.max_mhz = list(indices), period, [], set(), False, None
        self.window = None
        self.nv, self.handles = None, []
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.handles = [self._handle(nv, i) for i in self.indices]
            self.max_mhz = min(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in self.handles)
        except Exception as e:      # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    @staticmethod
    def _handle(nv, cuda_index):
        """NVML handle of CUDA device `cuda_index` (CUDA_VISIBLE_DEVICES may renumber): by UUID, else by index."""
        try:
            import torch
            u = str(torch.cuda.get_device_properties(cuda_index).uuid)
            return nv.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
        except Exception:
            return nv.nvmlDeviceGetHandleByIndex(cuda_index)

    def run(self):
        nv = self.nv
        if nv is None:
            return
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        try:
            while not self.stop_flag:
                if self.window is not None:                      # only while a timed region is open
                    mhz = min(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM) for h in self.handles)
                    self.samples.append(mhz)
                    for h in self.handles:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in names.items():
                            if r & bit:
                                self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def open(self):
        self.window = time.perf_counter()

    def close(self):
        self.window = None

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "gpus_sampled": len(self.handles)}


def make_inputs(first_scan: int, n_scans: int):
    """Seeded synthetic cfg2 inputs.  The generated arrays are kept under SO_BENCH_CACHE (default a /tmp
    directory) so that back-to-back runs on one box (N = 1, 2, 4, 8) do not repeat the CPU-side synthesis; the cache holds
    inputs only, never results."""
    from superodom_b200 import synth
    cache = os.environ.get("SO_BENCH_CACHE", "/tmp/superodom_b200_bench_inputs")       # "" disables
    path = os.path.join(cache, f"cfg2_{first_scan}_{n_scans}.npz") if cache else None
    if path and os.path.exists(path):
        try:
            z = np.load(path)
            nn = z["n"]
            offs = np.concatenate([[0], np.cumsum(nn)])
            flat = z["flat"]
            return z["map"], [flat[offs[i]:offs[i + 1]] for i in range(len(nn))], z["priors"], z["truths"]
        except Exception as e:                                 # unreadable / half-written cache: regenerate
            print(f"[bench] input cache {path} unusable ({type(e).__name__}), regenerating", file=sys.stderr)
    scene, map_xyzi = synth.make_map_for("cfg2")
    scans, priors, truths = [], [], []
    for i in range(first_scan, first_scan + n_scans):
        c = synth.make_case_on(scene, map_xyzi, "cfg2", i)
        scans.append(c["scan_xyzi"])
        priors.append(c["pose_prior"])
        truths.append(c["pose_true"])
    if path:
        try:
            os.makedirs(cache, exist_ok=True)
            tmp = path + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, map=map_xyzi, flat=np.concatenate(scans, 0), n=np.array([len(x) for x in scans]), priors=np.stack(priors), truths=np.stack(truths))
            os.replace(tmp, path)                              # atomic: ranks / back-to-back runs may race on the same file
        except Exception as e:                                 # read-only or full /tmp: the cache is only a convenience
            print(f"[bench] input cache not written ({type(e).__name__}: {e})", file=sys.stderr)
    return map_xyzi, scans, np.stack(priors), np.stack(truths)


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path on the host cores: oracle/_ref (reference octree.h compiled
    verbatim + restated fits/solver) when present, else the oracle port.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    sample = max(1, args.ref_scans)
    map_xyzi, scans, priors, truths = make_inputs(0, sample)
    ref = O.has_ref_octree()
    om = O.OracleMap(map_xyzi, ref_octree=ref)
    mode = 2 if ref else 0

    def step():
        for s, p in zip(scans, priors):
            r = om.register(s, p, 0.2, 20, 0, knn_mode=mode, n_threads=cores, skip_map_checks=True)
            assert r.status == 0
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    kind = "reference-octree+port" if ref else "port"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "scans_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} cfg2 scans per step, per-point loop on {cores} threads, k-NN = {kind}"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from superodom_b200 import api, replay

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    t_phase = [time.perf_counter()]

    def phase(label):                                  # wall-clock of the untimed set-up phases, to stderr
        now = time.perf_counter()
        print(f"[bench rank {rank}] {label}: {now - t_phase[0]:.1f} s", file=sys.stderr, flush=True)
        t_phase[0] = now

    # input synthesis (numba ray casting) must not oversubscribe the host when several ranks generate at once
    os.environ.setdefault("NUMBA_NUM_THREADS", str(max(1, min(16, (os.cpu_count() or 8) // max(world, 1)))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line only (NCCL logs there by default)
        dist.init_process_group("nccl", device_id=dev)
    phase("import torch + process group")
    B = args.batch
    n_total = B * world
    b0, b1 = replay.shard_range(n_total, rank, world)
    map_xyzi, scans, priors, truths = make_inputs(b0, b1 - b0)
    phase(f"synthetic inputs ({b1 - b0} scans)")
    n_points = np.array([len(s) for s in scans], np.uint32)
    flat = np.ascontiguousarray(np.concatenate(scans, 0))

    ctx = api.Context(device=local, max_map_points=len(map_xyzi) + 1024, max_scan_points=int(n_points.max()), max_batch=B, plane_res=0.2)
    stream = torch.cuda.Stream(device=dev)             # a real (non-legacy) stream shared by torch events and the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_set_points(map_xyzi)
    d_scans = torch.from_numpy(flat).to(dev)
    h_pinned = torch.from_numpy(flat).pin_memory()
    h_view = h_pinned.numpy()

    def gather(res):
        poses = np.array([list(r.pose) for r in res])
        return replay.gather_poses(poses, n_total, rank, world, dev)

    def step_device():
        res = ctx.register_batch_device(d_scans.data_ptr(), n_points, priors, 20, 0, skip_map_checks=True)
        return res, gather(res)

    def step_host():
        res = ctx.register_batch(h_view, n_points, priors, 20, 0, skip_map_checks=True)
        return res, gather(res)

    def timed(step_fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res, allp = step_fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res, allp

    sampler = ClockSampler(range(world)) if rank == 0 else None      # torchrun on one node: local GPU indices 0..world-1
    if sampler:
        sampler.start()
    phase("context, map upload, pinned staging, NVML init")
    for _ in range(max(args.warmup, 3)):
        res, allp = step_device()
    phase("warm-up steps")
    # correctness guard: the timed thing really registers the scans (cm-level agreement with ground truth)
    err = np.linalg.norm(np.array([list(r.pose) for r in res])[:, :3] - truths[:, :3], axis=1)
    assert all(r.status == 0 for r in res) and err.max() < 0.05, (err.max(), [r.status for r in res])

    ctx.kernel_launches(reset=True)
    if sampler:
        sampler.open()
    ms, res, allp = timed(step_device, args.steps)
    if sampler:
        sampler.close()
    launches = ctx.kernel_launches()
    value = n_total * args.steps / (ms * 1e-3)

    for _ in range(2):
        step_host()
    ctx.bytes_copied(reset=True)
    if sampler:
        sampler.open()
    ms_e2e, _, _ = timed(step_host, args.steps)
    clocks = sampler.result() if sampler else None
    h2d, d2h = ctx.bytes_copied()
    e2e = n_total * args.steps / (ms_e2e * 1e-3)
    phase("timed regions (device-resident + e2e)")

    line = None
    if rank == 0:
        # roofline of the dominant kernel (k_correspond): one profiled step, CUDA events around every launch that has work
        ctx.profile_enable(True)
        for k in range(6):
            ctx.profile_get(k, reset=True)
        pres = ctx.register_batch_device(d_scans.data_ptr(), n_points, priors, 20, 0, skip_map_checks=True)
        ctx.profile_enable(False)
        ms_k, n_k = ctx.profile_get(0)          # k_knn_scan
        ms_f, n_f = ctx.profile_get(4)          # first evaluation of each solve (k_evaluate<PH_CORR> + k_lm_step)
        ms_q, n_q = ctx.profile_get(5)          # k_fit (split build)
        ms_e, n_e = ctx.profile_get(1)          # k_evaluate (+ k_lm_step)
        ms_p, n_p = ctx.profile_get(3)          # scan ordering (keys + radix sort + gather)
        peak, peak_src = _peaks()
        scan_passes = sum(int(r.n_iterations) * int(n) for r, n in zip(pres, n_points))
        # algorithmic bytes of the k-NN kernel: 16 B scan read + 21 B neighbour ids/flag + 80 B neighbour
        # points handed to k_fit, per processed point and ICP iteration, + the map streamed once per launch
        fused = bool(api.build_flags() & 1)
        kname = "k_knn_fit" if fused else "k_knn_scan"
        alg_bytes = scan_passes * (MATCH_BYTES_PER_POINT if fused else KNN_BYTES_PER_POINT) + n_k * len(map_xyzi) * 16
        achieved = alg_bytes / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
        tot = ms_k + ms_q + ms_f + ms_e + ms_p + 1e-12
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": NCU_DRAM_BYTES_PER_POINT[kname] * scan_passes / max(n_k, 1) if NCU_DRAM_BYTES_PER_POINT.get(kname) else None,
                    "traffic_source": f"profiles/ (ncu --set full, dram__bytes_read+write per point of one {kname} launch, scaled to this launch size)",
                    "peak_source": peak_src, "launches_profiled": int(n_k), "avg_launch_ms": ms_k / max(n_k, 1),
                    "algorithmic_bytes_per_launch": alg_bytes / max(n_k, 1),
                    "note": "instruction-issue bound (ncu: 74% issue-active), L1/L2-resident gathers; see DESIGN.md section 4",
                    "first_evaluation": {"launches": int(n_f), "avg_launch_ms": ms_f / max(n_f, 1)},
                    "k_evaluate": {"launches": int(n_e), "avg_launch_ms": ms_e / max(n_e, 1)},
                    "k_fit": {"launches": int(n_q), "avg_launch_ms": ms_q / max(n_q, 1)},
                    "share_of_step": {kname: ms_k / tot, "k_fit": ms_q / tot, "first_evaluation": ms_f / tot, "k_evaluate": ms_e / tot, "scan_ordering": ms_p / tot}}
        # CPU baseline on the host cores: the oracle (reference octree verbatim when oracle/_ref travelled), 1 thread, bounded sample
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            om = O.OracleMap(map_xyzi, ref_octree=O.has_ref_octree())
            mode = 2 if O.has_ref_octree() else 0
            t0 = time.perf_counter()
            ns = 0
            while ns < len(scans) and ns < 8 and (time.perf_counter() - t0 < 12.0 or ns < 2):
                ro = om.register(scans[ns], priors[ns], 0.2, 20, 0, knn_mode=mode, n_threads=1, skip_map_checks=True)
                dpos = np.abs(np.array(ro.pose)[:3] - np.array(res[ns].pose)[:3]).max()
                ns += 1
            dt = time.perf_counter() - t0
            cpu = {"value": ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{ns} of the step's cfg2 scans, whole ICP, 1 thread (the reference's feature loop and Ceres solve are serial); "
                             f"k-NN = {'reference octree.h compiled verbatim' if mode == 2 else 'oracle exact grid'}; last |dpos| vs GPU {dpos:.2e} m"}
        phase("profiling pass + cpu baseline")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "scans_per_step": n_total, "parallelism": f"replay-shard x{world}",
                           "l2": "inputs larger than L2 (scans+correspondences per step >> 126 MB)",
                           "icp_iterations_executed_mean": float(np.mean([r.n_iterations for r in res]))},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d // args.steps), "d2h_bytes_per_step": int(d2h // args.steps),
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="scans per step per GPU")
    ap.add_argument("--ref-scans", type=int, default=2, help="scans per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
