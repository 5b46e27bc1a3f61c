#!/usr/bin/env python
"""bench.py — images/sec of one DSL teacher+student step (FCOS-R50-FPN, synthetic COCO-shaped 1333x800 -> padded
800x1344, bs 4 per GPU), the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = EMA-teacher forward (no_grad, eval) on B weak images + decode/score gate, student forward + FCOSHead
loss + backward on B strong images, [grad all-reduce], clip-grad-norm + momentum SGD, EMA update, operand repack.
Rank 0 prints ONE JSON line. `value` is timed with the inputs already resident in HBM; `e2e` is the same step driven
through the public API (DSLEngine.set_inputs + step) from PINNED HOST buffers with the losses read back every step.
`roofline` is for the dominant kernel family (the tcgen05 implicit-GEMM conv), timed live with CUDA events on the
launching stream in an instrumented eager pass over the same workload. `cpu_baseline` is the oracle port of the
reference's CPU arithmetic (oracle/cpu_step.py) on a bounded sample, on this box's host cores.
`--impl reference` times that CPU port alone (the reference's own Python needs mmcv, which is neither on the GPU box
nor installable offline) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec (teacher+student step) FCOS-R50 1333x800"
UNIT = "img/s"
B_PER_GPU, H, W = 4, 800, 1344
WORKLOAD = "configs[1]: FCOS-R50-FPN DSL teacher-student, synthetic COCO 1333x800 (padded 800x1344), bs=4/GPU, bf16"
# the other BASELINE.json configs that name a throughput case (secondary lines, committed under profiles/)
WORKLOADS = {
    "configs1": dict(depth=50, batch=4, text=WORKLOAD),
    "configs3": dict(depth=101, batch=2, text="configs[3]: FCOS-R101-FPN DSL teacher-student, synthetic 1333x800 (padded "
                                              "800x1344), bs=2/GPU, bf16"),
    "configs4": dict(depth=50, batch=4, text="configs[4]: FCOS-R50 multi-scale 800-1333 short-edge {640,800}, PatchShuffle "
                                             "p=.5 + flip p=.5 views rendered on the device, variable padded shapes, "
                                             "bs=4/GPU, bf16"),
}


def make_gt(seed, B, h, w, max_gt=20, max_ignore=5):
    """Synthetic COCO-shaped boxes, SURVEY 8(d)'s `_demo_mm_inputs` recipe (cx, cy, bw, bh ~ U(0,1), clipped to the image;
    labels U{0..79}; a few ignore boxes). Returns per-image lists of fp32 (n,4) / int64 (n,) / fp32 (m,4) tensors."""
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    gts, labels, ignores = [], [], []
    for _ in range(B):
        def boxes(n):
            cx, cy, bw, bh = rng.rand(4, n)
            x1 = np.clip((cx * w - w * bw / 2), 0, w)
            y1 = np.clip((cy * h - h * bh / 2), 0, h)
            x2 = np.clip((cx * w + w * bw / 2), 0, w)
            y2 = np.clip((cy * h + h * bh / 2), 0, h)
            return torch.from_numpy(np.stack([x1, y1, x2, y2], 1).astype(np.float32))
        n = int(rng.randint(1, max_gt + 1))
        gts.append(boxes(n))
        labels.append(torch.from_numpy(rng.randint(0, 80, size=n).astype(np.int64)))
        ignores.append(boxes(int(rng.randint(0, max_ignore + 1))))
    return gts, labels, ignores


def confident_heads(eng, bias0=-2.0):
    """Random-init FCOS has conv_cls bias -log(99): every score is 0.01 < score_thr 0.05 and the teacher's decode / NMS /
    pseudo-label kernels would be timed on ZERO candidates. One class gets a trained-like bias so that every step runs
    them on a few thousand gated candidates per image (reported as `cand_counts`)."""
    for net in (eng.student, eng.teacher):
        net.store["bbox_head.conv_cls.bias"][0] = bias0
    eng.student.repack(everything=True)
    eng.teacher.repack(everything=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager-gpu"],
                    help="reference: the reference arithmetic on the host cores; eager-gpu: the same arithmetic under stock "
                         "torch eager + cuDNN on this GPU (SURVEY 8(d)'s GPU comparator; a baseline, not the product)")
    ap.add_argument("--workload", default="configs1", choices=sorted(WORKLOADS),
                    help="configs1 (default) = the BASELINE.json headline config; configs3 = R101 bs 2; configs4 = multi-scale "
                         "+ PatchShuffle, variable padded shapes through the runner's per-shape engine cache")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the workload's)")
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--backbone", default="resnet", choices=["resnet", "rla"],
                    help="rla: RLA_ResNet, the backbone of the shipped DSL configs (not the BASELINE.json config)")
    ap.add_argument("--mix", default="bench", choices=["bench", "literal"],
                    help="literal: the reference's hard-coded per-GPU mix (SURVEY 8d): 2 student images + the half-resolution "
                         "SI copy of the last one (semi_epoch_based_runner.py:186-204) + 1 teacher image "
                         "(unlabel_pred_hook.py:517,543), SI-soft loss on; secondary measurement, not the BASELINE config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-view-bench", action="store_true", help="skip the side measurement of dslb_view_images")
    ap.add_argument("--cpu-sample-hw", default="800x1344", help="HxW of the bounded CPU sample")
    ap.add_argument("--no-ncu-traffic", action="store_true", help="skip the live ncu DRAM-traffic capture of the conv family")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    a.batch = wl["batch"] if a.batch is None else a.batch
    a.depth = wl["depth"] if a.depth is None else a.depth
    a.workload_text = wl["text"] if (a.batch, a.depth) == (wl["batch"], wl["depth"]) else \
        wl["text"] + f" [overridden: depth {a.depth}, bs {a.batch}/GPU]"
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                       reasons=sorted(reasons))
        return out


def ncu_traffic_per_launch(args, timeout=420):
    """DRAM read + write bytes per launch of the conv_igemm family, measured LIVE on the benched libdslb.so: one eager
    step of the same workload (tools/profile_step.py) under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum
    --clock-control none`, in a process of its own after the step has been timed (bytes, not time, are taken from it).
    Falls back to the committed launch list of an earlier build — and says so — when ncu cannot run."""
    import csv
    import shutil
    out = os.path.join(tempfile.gettempdir(), f"dslb_traffic_{os.getpid()}.csv")
    try:
        ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
        cmd = [ncu, "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum",
               "--clock-control", "none", "-k", "regex:conv_igemm", "--csv", "--log-file", out,
               sys.executable, os.path.join(ROOT, "tools", "profile_step.py"), "--batch", str(args.batch), "--depth",
               str(args.depth), "--backbone", args.backbone]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        rows = [ln for ln in open(out) if ln.startswith('"')]
        rd = csv.DictReader(rows)
        per = {}
        for row in rd:
            per.setdefault(row["ID"], 0.0)
            v = float(row["Metric Value"].replace(",", ""))
            unit = row["Metric Unit"].lower()
            per[row["ID"]] += v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        if not per:
            raise RuntimeError((r.stderr or r.stdout)[-200:])
        return dict(bytes=round(sum(per.values()) / len(per)), launches=len(per),
                    source="live: ncu dram__bytes_read.sum + dram__bytes_write.sum over one eager step of this build "
                           "(tools/profile_step.py), per conv_igemm launch")
    except Exception as e:   # noqa: BLE001
        fb = _committed_traffic()
        if fb is not None:
            fb["source"] += f" [live ncu capture failed: {e!r:.120}]"
        return fb
    finally:
        try:
            os.unlink(out)
        except OSError:
            pass


def _committed_traffic():
    path = os.path.join(ROOT, "profiles", "r01d_launches_summary.txt")
    try:
        n_tot, b_tot = 0, 0.0
        for line in open(path):
            if "dslb::conv_igemm" in line:
                f = line.split()
                n, mb = int(f[1]), float(f[3])
                n_tot += n
                b_tot += n * mb * 1e6
        return dict(bytes=round(b_tot / n_tot), source="STALE: profiles/r01d_launches_summary.txt (round-1 build)") \
            if n_tot else None
    except OSError:
        return None


def cpu_baseline(sample_hw, depth, steps=1, warmup=0, batch=1, backbone="resnet"):
    """The reference's CPU arithmetic for the step (oracle port) on a bounded sample of the workload."""
    import torch
    from oracle import cpu_step
    h, w = (int(v) for v in sample_hw.split("x"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    r = cpu_step.time_cpu_steps(batch, h, w, steps=steps, warmup=warmup, depth=depth, threads=cores, backbone=backbone)
    return dict(value=round(r["images_per_sec"], 4), unit=UNIT, cores=r["cores"], kind="port",
                sample=f"{steps} teacher+student step(s) of B={batch} at {h}x{w}, {'RLA_' if backbone == 'rla' else ''}R{depth}, fp32 torch CPU "
                       f"({r['seconds_per_step']:.2f} s/step), warmup {warmup}")


def view_images_side_bench(timeout=120):
    """Side measurement reported under "view_images" (not part of the step, value or e2e): the device-side view pipeline
    kernel dslb_view_images at the benchmark batch shape — CUDA events around the bare launch, algorithmic GB/s against the
    measured HBM peak, the cv2 CPU pipeline timed beside it (tools/view_image_micro.py). Runs in a process of its own
    after the step has been measured, so nothing it does can cost the bench line."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "view_image_micro.py"), "--iters", "30"],
                           capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        if r.returncode != 0:
            return dict(error=(r.stderr or r.stdout)[-300:])
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return dict(error=repr(e))


def eager_gpu_side_bench(args, timeout=180):
    """SURVEY 8(d)'s GPU comparator beside the headline: `bench.py --impl eager-gpu` (the reference arithmetic under stock
    torch eager + cuDNN on this GPU, fp32 and bf16 autocast) in a process of its own after the step has been measured.
    Reported under "gpu_eager_baseline"; a failure there only yields an `error` entry."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "eager-gpu", "--steps", "5", "--warmup",
                            "2", "--batch", str(args.batch), "--depth", str(args.depth), "--backbone", args.backbone],
                           capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        rows = []
        for ln in r.stdout.strip().splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                rows.append({k: d[k] for k in ("dtype", "value", "unit", "ms_per_step", "steps", "losses") if k in d})
        if r.returncode != 0 or not rows:
            return dict(error=(r.stderr or r.stdout)[-300:], rows=rows)
        return dict(impl="torch eager + cuDNN, reference arithmetic (oracle restatement), same GPU", rows=rows)
    except Exception as e:
        return dict(error=repr(e))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    cb = cpu_baseline(args.cpu_sample_hw, args.depth, steps=steps, warmup=warm, batch=args.batch, backbone=args.backbone)
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=round(1000.0 * args.batch / cb["value"], 1), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic", impl="reference",
                config=dict(workload=args.workload_text, note="CPU port of the reference arithmetic, same per-GPU batch, bounded "
                                                    "number of steps; host cores only, no GPU"),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=round(time.time() - t0, 1))
    print(json.dumps(line), flush=True)


def run_eager_gpu(args):
    """SURVEY section 8(d)'s honest GPU comparator: the reference's arithmetic for one teacher+student step (the oracle
    restatement of oracle/cpu_step.py — functional torch + autograd, NCHW) executed by STOCK torch eager + cuDNN on this
    GPU at the benchmark shape; one JSON line per precision (fp32 with cuDNN's TF32 convs, bf16 autocast). Like the
    cpu_baseline leg this is a baseline: nothing of the product runs here. Rank 0 only."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.cpu_step import CpuStep
    assert torch.cuda.is_available(), "--impl eager-gpu needs a CUDA device"
    steps, warm = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))
    for name, amp in (("fp32 (cuDNN, TF32 convs allowed)", None), ("bf16 autocast", torch.bfloat16)):
        cs = CpuStep(args.batch, H, W, depth=args.depth, backbone=args.backbone, device="cuda")
        cs.autocast = amp
        for _ in range(warm):
            losses = cs.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            losses = cs.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps(dict(metric=METRIC, value=round(args.batch / (ms * 1e-3), 2), unit=UNIT, n_gpus=1, steps=steps,
                              warmup=warm, ms_per_step=round(ms, 2), higher_is_better=True, impl="eager-gpu",
                              dtype=name, data="synthetic",
                              config=dict(workload=WORKLOAD, note="reference arithmetic (oracle restatement) under stock "
                                          "torch eager + cuDNN; includes the host syncs of the reference's loss"),
                              losses={k: round(v, 4) for k, v in losses.items()})), flush=True)
        del cs
        torch.cuda.empty_cache()


def run_multiscale(args, world, rank, local):
    """BASELINE configs[4]: multi-scale (short edge 640 / 800, long edge <= 1333) + PatchShuffle p=.5 + flip p=.5, so the
    padded batch shape changes from step to step (SURVEY 8(d) recipe: source aspect drawn per batch from COCO-like
    {4:3, 3:2, 16:9, 3:4} — GroupSampler keeps a batch to one orientation — scale drawn per image). Every step renders
    the strong and weak views from uint8 sources ON THE DEVICE (dslb_view_images), picks the engine of the padded shape
    from the runner's per-shape cache (`SemiEpochBasedRunner._engine_for`: plan + CUDA graph per (B, H, W), shared
    weights / momentum / LR / adathres state) and runs the fused step. `value` times K steps right after W warm-up steps:
    building + capturing the plan of a shape met for the first time is INSIDE the timed region (`captures_in_timed_region`);
    `steady` repeats the same K-step sequence once every shape is cached. `e2e` adds the H2D of the uint8 sources from
    pinned host memory every step and the D2H of the losses."""
    import logging

    import numpy as np
    import torch
    import torch.distributed as dist

    from dsl_b200 import _lib as L
    from dsl_b200 import dist_ops, plugin
    from dsl_b200.geometry import image_view
    from dsl_b200.runner import SemiEpochBasedRunner

    B = args.batch
    model_cfg = dict(
        backbone=dict(type="ResNet", depth=args.depth, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                      norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="caffe"),
        neck=dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                  add_extra_convs="on_output", num_outs=5, relu_before_extra_convs=True),
        bbox_head=dict(type="FCOSHead", num_classes=80, in_channels=256, stacked_convs=4, feat_channels=256,
                       strides=[8, 16, 32, 64, 128], norm_on_bbox=True, centerness_on_reg=True, dcn_on_last_conv=False,
                       center_sampling=True, conv_bias=True, loss_weight=3.0,
                       loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                       loss_bbox=dict(type="GIoULoss", loss_weight=1.0),
                       loss_centerness=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0)),
        test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.6),
                      max_per_img=100))
    model, ema = plugin.FCOS(**model_cfg).cuda(), plugin.FCOS(**model_cfg).cuda()
    with torch.no_grad():
        model.store["bbox_head.conv_cls.bias"][0] = -2.0      # see confident_heads()
    model._dirty()
    ema.load_state_dict(model.state_dict())
    runner = SemiEpochBasedRunner(model, logger=logging.getLogger("bench"), max_epochs=1, ema_model=ema)
    runner.max_cached_shapes = 8
    # uint8 sources of the four aspects (cv2.imread layout), pinned on the host and resident on the device
    rng = np.random.RandomState(300 + rank)
    src_shapes = [(480, 640), (427, 640), (360, 640), (640, 480)]
    h_srcs = {hw: [torch.from_numpy(rng.randint(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)).pin_memory()
                   for _ in range(B)] for hw in src_shapes}
    d_srcs = {hw: [t.cuda() for t in v] for hw, v in h_srcs.items()}
    scales = [(1333, 640), (1333, 800)]
    import random
    mean, std = (103.53, 116.28, 123.675), (1.0, 1.0, 1.0)

    def draw_step(i):
        """the draws of step i, reproducible: (source aspect, strong views, weak views, padded H, W, boxes)"""
        np.random.seed(1000 * (rank + 1) + i)
        random.seed(1000 * (rank + 1) + i)
        hw = src_shapes[np.random.randint(len(src_shapes))]
        sv, wv = [], []
        for _ in range(B):      # Resize(multiscale 'value') -> PatchShuffle(p=.5, place U(.2,.8)) -> RandomFlip(p=.5)
            sc = scales[np.random.randint(len(scales))]
            ps = np.random.rand() < 0.5
            place = 0.2 + 0.6 * np.random.rand()
            mode = random.choice(["flip", "flop"]) if ps else None
            flip = np.random.rand() < 0.5
            sv.append(image_view(hw, sc, ps_mode=mode, ps_place=place, flip=flip)[0])
            wv.append(image_view(hw, sc)[0])    # weak view of the teacher: same scale, test pipeline (no PS / flip)
        up = lambda n: (n + 31) // 32 * 32  # noqa: E731
        Hh, Ww = up(max(v.img_h for v in sv)), up(max(v.img_w for v in sv))
        gts, labels, ignores = make_gt(5000 + i, B, min(v.img_h for v in sv), min(v.img_w for v in sv))
        return hw, sv, wv, Hh, Ww, gts, labels, ignores

    host_out = torch.zeros(4, dtype=torch.float32).pin_memory()
    stats = dict(captures=0)

    def step(i, from_host):
        hw, sv, wv, Hh, Ww, gts, labels, ignores = draw_step(i)
        n_eng = len(runner._engines)
        eng = runner._engine_for(B, Hh, Ww)
        fresh = len(runner._engines) != n_eng or eng.graphs is None
        stats["captures"] += int(fresh)
        if from_host:
            srcs = [t.cuda(non_blocking=True) for t in h_srcs[hw]]     # ~0.9 MB per image instead of 12.9 MB fp32
        else:
            srcs = d_srcs[hw]
        eng.set_images_from_sources(srcs, sv, teacher_srcs=srcs, teacher_views=wv, mean=mean, std=std, to_rgb=False)
        eng.set_inputs(None, gts, labels, ignores)
        losses = eng.step()
        if from_host:
            host_out[:3].copy_(torch.stack([losses["loss_cls"], losses["loss_bbox"], losses["loss_centerness"]]),
                               non_blocking=False)
        return eng

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(i0, n, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(i0, i0 + n):
            eng = step(i, from_host)
        e1.record()
        barrier()
        wall = time.time() - t0
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), eng

    W_, K = max(args.warmup, 3), args.steps
    for i in range(W_):
        step(i, False)
    L.reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stats["captures"] = 0
    ms, wall_ms, eng = timed(W_, K, False)
    cap_timed = stats["captures"]
    clocks = sampler.stop() if rank == 0 else {}
    launches = L.launch_count / K
    ms_steady, _, eng = timed(W_, K, False)           # same sequence again: every shape is cached now
    ms_e2e, _, eng = timed(W_, K, True)
    shapes = sorted(runner._engines)
    torch.cuda.synchronize()
    loss_vals = [float(v) for v in host_out[:3]]
    assert all(np.isfinite(loss_vals)), f"non-finite losses {loss_vals}"
    if rank != 0:
        dist_ops.shutdown(*runner._engines.values())
        return
    pk = peaks()
    # the device clock around the K steps includes the host-side plan builds (the GPU idles meanwhile); wall agrees
    h2d = sum(int(np.prod(hw)) * 3 for hw in src_shapes) / len(src_shapes) * B
    flops = eng.flops_per_step()
    line = dict(metric=METRIC, value=round(B * world * K / (ms * 1e-3), 2), unit=UNIT, n_gpus=world, steps=K, warmup=W_,
                ms_per_step=round(ms / K, 3), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=args.workload_text, global_batch=B * world, per_gpu_batch=B, teacher_batch=B,
                            parallelism=f"dp{world}", l2="working set (activations) >> 126 MB L2; no flush needed",
                            step="views rendered on the device (dslb_view_images), teacher fwd+decode+NMS+pseudo labels, "
                                 "student fwd+loss+bwd, grad allreduce, clip+SGD, EMA, repack; engine per padded shape",
                            shapes=[list(k) for k in shapes], cuda_graph=True,
                            captures_in_timed_region=cap_timed, wall_ms_per_step=round(wall_ms / K, 3)),
                steady=dict(value=round(B * world * K / (ms_steady * 1e-3), 2), unit=UNIT,
                            ms_per_step=round(ms_steady / K, 3), note="same K-step sequence, every shape's plan cached"),
                e2e=dict(value=round(B * world * K / (ms_e2e * 1e-3), 2), unit=UNIT, h2d_bytes_per_step=int(h2d),
                         d2h_bytes_per_step=12, ms_per_step=round(ms_e2e / K, 3),
                         note="uint8 sources H2D from pinned memory every step, plans cached"),
                gpu_launches=int(launches * K), gpu_launches_per_step=int(launches),
                roofline=dict(bound="tensor", kernel="whole step of the last shape's engine (conv_igemm family dominates; "
                                                     "per-family split: the configs[1] line)",
                              achieved=round(flops / (ms_steady / K * 1e-3) / 1e12, 1), peak=pk["tf_burst"],
                              unit="TFLOP/s", frac=round(flops / (ms_steady / K * 1e-3) / 1e12 / pk["tf_burst"], 4),
                              traffic=None, peak_source=pk["src"] + " burst",
                              note="upper bound on the mix: FLOPs of the LAST step's shape over the mean steady step time"),
                cpu_baseline=None, clocks=clocks,
                losses=dict(loss_cls=loss_vals[0], loss_bbox=loss_vals[1], loss_centerness=loss_vals[2]),
                cand_counts=eng.cand_counts.tolist())
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline("640x960", args.depth, steps=2, warmup=1, batch=B, backbone=args.backbone)
    print(json.dumps(line), flush=True)
    dist_ops.shutdown(*runner._engines.values())


def main():
    args = parse()
    if os.environ.get("DSLB_HANG_DUMP"):      # debugging aid: dump every thread's stack if the run is still going then
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["DSLB_HANG_DUMP"]), exit=True)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager-gpu":
        return run_eager_gpu(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    from dsl_b200 import _lib as L
    from dsl_b200 import dist_ops
    from dsl_b200.trainer import DSLEngine

    if args.workload == "configs4":
        return run_multiscale(args, world, rank, local)
    literal = args.mix == "literal"
    B = 2 if literal else args.batch
    tB = 1 if literal else B
    ekw = dict(scale_invariant=True, soft_weight=1.0, soft_warm_up=5000, teacher_B=1) if literal else {}
    eng = DSLEngine(B, H, W, depth=args.depth, seed=0, use_graphs=True, backbone=args.backbone, **ekw)
    rng = np.random.RandomState(100 + rank)
    # synthetic COCO-shaped batch in PINNED host memory (mean-subtracted pixels, caffe normalisation: std 1)
    img_s = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32)).pin_memory()
    img_t = torch.from_numpy((rng.rand(tB, 3, H, W) * 255 - 115).astype(np.float32)).pin_memory()
    gts, labels, ignores = make_gt(200 + rank, B, H, W, max_gt=20, max_ignore=5)
    confident_heads(eng)
    gts = [g.pin_memory() for g in gts]
    labels = [l.pin_memory() for l in labels]
    ignores = [i.pin_memory() for i in ignores]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    eng.set_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
    eng.step()
    torch.cuda.synchronize()
    cand0, det0 = eng.cand_counts.tolist(), eng.post.det_count.tolist()   # teacher post-processing load of the first step
    for _ in range(max(args.warmup, 3) - 1):
        eng.step()
    L.reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(eng.step, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step / 1e3)

    # ---------------------------------------------------------------- end to end from pinned host buffers (`e2e`)
    host_out = torch.zeros(4, dtype=torch.float32).pin_memory()

    def e2e_step():
        # the step consumes the batch prefetched under the previous step; the H2D of the NEXT batch (pinned host ->
        # device staging, every step) is issued on the copy stream before this step's result is read back
        if literal:   # the SI extra image is built by set_inputs (in-stream H2D + one kernel), no prefetch path
            eng.set_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
        losses = eng.step()
        if not literal:
            eng.prefetch_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
        host_out[:3].copy_(torch.stack([losses["loss_cls"], losses["loss_bbox"], losses["loss_centerness"]]),
                           non_blocking=False)  # D2H read of the step's result: synchronises every step

    if not literal:
        eng.prefetch_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    h2d = img_s.numel() * 4 + img_t.numel() * 4 + sum(g.numel() * 4 for g in gts) + \
        sum(l.numel() * 8 for l in labels) + sum(i.numel() * 4 for i in ignores) + 2 * (B + 1) * 4
    e2e = dict(value=round(B * world / (ms_e2e / 1e3), 2), unit=UNIT, h2d_bytes_per_step=int(h2d),
               d2h_bytes_per_step=12, ms_per_step=round(ms_e2e, 3))
    loss_vals = [float(v) for v in host_out[:3]]
    assert all(np.isfinite(loss_vals)), f"non-finite losses {loss_vals}"

    # ---------------------------------------------------------------- roofline of the dominant kernel (conv igemm)
    pk = peaks()
    prof = eng.profile_kernels(steps=3, ridge=pk["tf_sust"] * 1e12 / (pk["hbm"] * 1e9))
    launches = prof["launches_per_step"]
    if rank == 0 and os.environ.get("DSLB_PLAN_TABLE"):
        with open(os.environ["DSLB_PLAN_TABLE"], "w") as f:
            for r in prof["plans"]:
                f.write(json.dumps(r) + "\n")
    ck = prof["conv_igemm"]
    achieved = ck["flops"] / (ck["ms"] * 1e-3) / 1e12 if ck["ms"] > 0 else 0.0
    traffic = None
    if rank == 0 and world == 1 and not args.no_ncu_traffic:
        traffic = ncu_traffic_per_launch(args)
    # which measured peak: the timed region of this run in seconds decides (B200_PROFILING.md: the burst figure for
    # short regions at boost clocks, the sustained one inside a long, power-capped run)
    long_run = ms * 1e-3 >= 5.0
    tf_peak = pk["tf_sust"] if long_run else pk["tf_burst"]
    roofline = dict(bound="tensor", kernel="conv_igemm_kernel + conv_igemm_fast4_kernel (fprop + dgrad implicit GEMM, "
                                           "tcgen05)",
                    achieved=round(achieved, 1), peak=tf_peak, unit="TFLOP/s",
                    frac=round(achieved / tf_peak, 4), frac_of_sustained_peak=round(achieved / pk["tf_sust"], 4),
                    frac_of_burst_peak=round(achieved / pk["tf_burst"], 4),
                    traffic=traffic["bytes"] if traffic else None,
                    traffic_unit="bytes/launch (DRAM read+write)", traffic_source=traffic["source"] if traffic else None,
                    algorithmic_bytes_per_launch=round(sum(r[2] for r in prof.get("rows", [])) / max(ck["n"], 1))
                    if prof.get("rows") else None,
                    peak_source=pk["src"] + (" sustained (timed region >= 5 s)" if long_run else
                                             " burst (timed region %.2f s at boost clocks)" % (ms * 1e-3)),
                    launches_per_step=ck["n"], avg_launch_us=round(1e3 * ck["ms"] / max(ck["n"], 1), 2),
                    flops_per_step=ck["flops"], share_of_step=round(ck["ms"] / prof["step_ms"], 4),
                    share_of="the instrumented EAGER step (%.2f ms: events around every conv launch, teacher branch "
                             "serialised), not the %.2f ms graph step" % (prof["step_ms"], ms_per_step),
                    wgrad=dict(achieved=round(prof["conv_wgrad"]["flops"] / max(prof["conv_wgrad"]["ms"], 1e-9) / 1e9,
                                              1), unit="TFLOP/s", launches_per_step=prof["conv_wgrad"]["n"],
                               share_of_step=round(prof["conv_wgrad"]["ms"] / prof["step_ms"], 4)),
                    head_tower=prof.get("head_tower"),
                    whole_step_tflops=round(eng.flops_per_step() / (ms_per_step * 1e-3) / 1e12, 1))

    try:
        # the same launches split by what can bound them: arithmetic intensity below the ridge (tensor peak / HBM peak)
        # => HBM roofline (algorithmic bytes / time), else tensor roofline
        bb = prof["by_bound"]
        t_, h_ = bb["tensor"], bb["hbm"]
        roofline["by_bound"] = dict(
            ridge_flop_per_byte=round(pk["tf_sust"] * 1e12 / (pk["hbm"] * 1e9), 1),
            tensor=dict(achieved=round(t_["flops"] / max(t_["ms"], 1e-9) / 1e9, 1), peak=pk["tf_sust"], unit="TFLOP/s",
                        frac=round(t_["flops"] / max(t_["ms"], 1e-9) / 1e9 / pk["tf_sust"], 4), launches_per_step=t_["n"],
                        share_of_step=round(t_["ms"] / prof["step_ms"], 4)),
            hbm=dict(achieved=round(h_["bytes"] / max(h_["ms"], 1e-9) / 1e6, 1), peak=pk["hbm"], unit="GB/s",
                     frac=round(h_["bytes"] / max(h_["ms"], 1e-9) / 1e6 / pk["hbm"], 4), launches_per_step=h_["n"],
                     share_of_step=round(h_["ms"] / prof["step_ms"], 4),
                     tflops=round(h_["flops"] / max(h_["ms"], 1e-9) / 1e9, 1)))
    except Exception as e:  # bookkeeping only: never lose the bench line over it
        roofline["by_bound"] = dict(error=repr(e))

    if rank != 0:
        dist_ops.shutdown(eng)
        return

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.cpu_sample_hw, args.depth, steps=2, warmup=1, batch=args.batch, backbone=args.backbone)

    views = eager = None
    if world == 1 and not args.no_view_bench:
        views = view_images_side_bench()
    if world == 1 and not args.no_cpu_baseline and not literal:
        eager = eager_gpu_side_bench(args)

    line = dict(metric=METRIC, value=round(value, 2), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=round(ms_per_step, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=(args.workload_text if args.backbone == "resnet" else WORKLOAD.replace(
                    "configs[1]: FCOS-R50-FPN", "variant of configs[1] with the shipped configs' RLA_ResNet backbone: FCOS-RLA_R50-FPN"))
                    if not literal else WORKLOAD.replace("configs[1]:", "reference-literal mix (SURVEY 8d) of configs[1]:").replace(
                        "bs=4/GPU", "2 student images + half-res SI copy + 1 teacher image per GPU, SI-soft loss on")
                    + ("" if args.backbone == "resnet" else ", RLA_ResNet backbone"),
                            global_batch=B * world, per_gpu_batch=B, teacher_batch=tB,
                            parallelism=f"dp{world}", l2="working set (activations) >> 126 MB L2; no flush needed",
                            step="teacher fwd+decode gate, student fwd+loss+bwd, grad allreduce, clip+SGD, EMA, repack",
                            cuda_graph=True),
                e2e=e2e, gpu_launches=int(launches * args.steps), gpu_launches_per_step=int(launches),
                roofline=roofline, cpu_baseline=cb, clocks=clocks,
                losses=dict(loss_cls=loss_vals[0], loss_bbox=loss_vals[1], loss_centerness=loss_vals[2]),
                cand_counts=dict(first_step=cand0, last_step=eng.cand_counts.tolist(),
                                 note="gated candidates per teacher image entering NMS (random data: the training "
                                      "signal pushes the confident class down as the run proceeds)"),
                det_counts=dict(first_step=det0, last_step=eng.post.det_count.tolist()),
                view_images=views, gpu_eager_baseline=eager)
    print(json.dumps(line), flush=True)
    dist_ops.shutdown(eng)


if __name__ == "__main__":
    main()
