#!/usr/bin/env python
"""Kernel timeline of the captured teacher+student step (CUPTI activity records through torch.profiler — nsys is not in
the image): every kernel of `--steps` graph replays with its stream, start and duration, plus a summary: busy time per
stream, wall time of the step, time with >= 2 kernels in flight, idle gaps, the top kernels by time.

    python tools/step_timeline.py [--batch 4] [--hw 800x1344] [--steps 2] [--out gpurun_out/timeline.jsonl]
    torchrun --nproc-per-node 2 tools/step_timeline.py ...        (multi-rank: NCCL kernels show up on their stream)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--hw", default="800x1344")
ap.add_argument("--depth", type=int, default=50)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/timeline.jsonl")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from bench import confident_heads, make_gt  # noqa: E402
from dsl_b200.trainer import DSLEngine  # noqa: E402

H, W = (int(v) for v in a.hw.split("x"))
B = a.batch
eng = DSLEngine(B, H, W, depth=a.depth, seed=0, use_graphs=True)
rng = np.random.RandomState(100 + rank)
img_s = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32))
img_t = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32))
gts, labels, ignores = make_gt(200 + rank, B, H, W)
confident_heads(eng)
eng.set_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
for _ in range(4):
    eng.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.steps):
        eng.step()
    torch.cuda.synchronize()
import tempfile  # noqa: E402
tmp = os.path.join(tempfile.gettempdir(), f"dslb_trace_{os.getpid()}.json")
prof.export_chrome_trace(tmp)
trace = json.load(open(tmp))
os.unlink(tmp)
evs = []
for e in trace.get("traceEvents", []):
    if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
        evs.append(dict(name=e["name"], start_us=float(e["ts"]), dur_us=float(e["dur"]),
                        stream=(e.get("args") or {}).get("stream")))
evs.sort(key=lambda r: r["start_us"])
if not evs:
    print("no CUDA activity records captured")
    sys.exit(1)
t0 = evs[0]["start_us"]
for r in evs:
    r["start_us"] = round(r["start_us"] - t0, 2)
    r["dur_us"] = round(r["dur_us"], 2)
out = a.out if world == 1 else a.out.replace(".jsonl", f".rank{rank}.jsonl")
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
with open(out, "w") as f:
    for r in evs:
        f.write(json.dumps(r) + "\n")
# ---- summary
kern = [r for r in evs if not r["name"].lower().startswith(("memcpy", "memset"))]
end = max(r["start_us"] + r["dur_us"] for r in evs)
pts = sorted([(r["start_us"], 1) for r in evs] + [(r["start_us"] + r["dur_us"], -1) for r in evs])
busy1 = busy2 = 0.0
depth, last = 0, 0.0
for t, d in pts:
    if depth >= 1:
        busy1 += t - last
    if depth >= 2:
        busy2 += t - last
    depth += d
    last = t
by_name = {}
for r in kern:
    k = r["name"][:60]
    by_name.setdefault(k, [0.0, 0])
    by_name[k][0] += r["dur_us"]
    by_name[k][1] += 1
by_stream = {}
for r in evs:
    by_stream.setdefault(r["stream"], 0.0)
    by_stream[r["stream"]] += r["dur_us"]
n = a.steps
print(f"rank {rank}: {len(evs)} GPU activities over {n} step(s); span {end / n:.1f} us/step; some kernel running "
      f"{busy1 / n:.1f} us/step ({100 * busy1 / end:.1f} %), >= 2 in flight {busy2 / n:.1f} us/step; "
      f"sum of durations {sum(r['dur_us'] for r in evs) / n:.1f} us/step")
print("busy us/step per stream:", {str(k): round(v / n, 1) for k, v in sorted(by_stream.items(), key=lambda kv: -kv[1])})
for k, (us, c) in sorted(by_name.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {us / n:9.1f} us/step  x{c / n:6.1f}  {k}")
from dsl_b200 import dist_ops as _D  # noqa: E402
_D.shutdown(eng)
