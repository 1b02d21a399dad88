"""Throughput of the cfg2 replay step as a function of the call size and the chunking (run under gpurun).

    python scripts/ab_step.py "64:0,64:4,128:0,128:8,256:0"      # B:SO_CHUNKS pairs (0 = the library's own rule)

For every pair: scans/s device-resident (so_register_batch_device) and end to end (so_register_batch, pinned host scans), 640 scans
per measurement drawn from 256 distinct scans, CUDA-event timed after warm-up.  One JSON line per pair.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from superodom_b200 import api  # noqa: E402

pairs = [tuple(int(x) for x in p.split(":")) for p in (sys.argv[1] if len(sys.argv) > 1 else "64:0,128:0").split(",")]
TOTAL, DISTINCT = 640, 256
map_xyzi, scans, priors, truths = bench.make_inputs(0, DISTINCT)
n_points = np.array([len(s) for s in scans], np.uint32)
offs = np.concatenate([[0], np.cumsum(n_points.astype(np.int64))])
flat = np.ascontiguousarray(np.concatenate(scans, 0))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
d_scans = torch.from_numpy(flat).to(dev)
h_pinned = torch.from_numpy(flat).pin_memory()
h_view = h_pinned.numpy()
for B, chunks in pairs:
    if chunks:
        os.environ["SO_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("SO_CHUNKS", None)
    ctx = api.Context(max_map_points=len(map_xyzi) + 1024, max_scan_points=int(n_points.max()), max_batch=B, plane_res=0.2)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_set_points(map_xyzi)
    calls = [((k * B) % DISTINCT, (k * B) % DISTINCT + B) for k in range(TOTAL // B)]

    def run_dev():
        for a, b in calls:
            r = ctx.register_batch_device(d_scans.data_ptr() + int(offs[a]) * 16, n_points[a:b], priors[a:b], 20, 0, skip_map_checks=True)
        return r

    def run_host():
        for a, b in calls:
            r = ctx.register_batch(h_view[offs[a]:offs[b]], n_points[a:b], priors[a:b], 20, 0, skip_map_checks=True)
        return r

    out = {"B": B, "chunks": chunks, "scans": len(calls) * B}
    for name, fn in (("device", run_dev), ("e2e", run_host)):
        for _ in range(2):
            r = fn()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name + "_scans_per_s"] = len(calls) * B / (float(np.median(ts)) * 1e-3)
        a, b = calls[-1]
        out[name + "_err"] = float(np.abs(np.array([list(x.pose) for x in r])[:, :3] - truths[a:b, :3]).max())
    ctx.close()
    print(json.dumps(out), flush=True)
