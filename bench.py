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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager-gpu"],
                    help="reference: the reference arithmetic on the host cores; eager-gpu: the same arithmetic under stock "
                         "torch eager + cuDNN on this GPU (SURVEY 8(d)'s GPU comparator; a baseline, not the product)")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="images per GPU (default: the benchmark config)")
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--backbone", default="resnet", choices=["resnet", "rla"],
                    help="rla: RLA_ResNet, the backbone of the shipped DSL configs (not the BASELINE.json config)")
    ap.add_argument("--mix", default="bench", choices=["bench", "literal"],
                    help="literal: the reference's hard-coded per-GPU mix (SURVEY 8d): 2 student images + the half-resolution "
                         "SI copy of the last one (semi_epoch_based_runner.py:186-204) + 1 teacher image "
                         "(unlabel_pred_hook.py:517,543), SI-soft loss on; secondary measurement, not the BASELINE config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-view-bench", action="store_true", help="skip the side measurement of dslb_view_images")
    ap.add_argument("--cpu-sample-hw", default="800x1344", help="HxW of the bounded CPU sample")
    return ap.parse_args()


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


def ncu_traffic_per_launch():
    """DRAM bytes per launch of the conv_igemm family (both kernels) from the committed ncu launch list of
    `tools/profile_step.py` (same workload, `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`),
    summarised by tools/launch_summary.py into profiles/r01d_launches_summary.txt. None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r01d_launches_summary.txt")
    try:
        n_tot, b_tot = 0, 0.0
        for line in open(path):
            if "dslb::conv_igemm" in line:
                f = line.split()
                n, mb = int(f[1]), float(f[3])
                n_tot += n
                b_tot += n * mb * 1e6
        return dict(bytes=round(b_tot / n_tot), source="profiles/r01d_launches_summary.txt (ncu dram__bytes_read+write)") \
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
                config=dict(workload=WORKLOAD, note="CPU port of the reference arithmetic, same per-GPU batch, bounded "
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


def main():
    args = parse()
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
    from dsl_b200.trainer import DSLEngine
    from tests.golden import inputs as GI

    literal = args.mix == "literal"
    B = 2 if literal else args.batch
    tB = 1 if literal else B
    ekw = dict(scale_invariant=True, soft_weight=1.0, soft_warm_up=5000, teacher_B=1) if literal else {}
    eng = DSLEngine(B, H, W, depth=args.depth, seed=0, use_graphs=True, backbone=args.backbone, **ekw)
    rng = np.random.RandomState(100 + rank)
    # synthetic COCO-shaped batch in PINNED host memory (mean-subtracted pixels, caffe normalisation: std 1)
    img_s = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32)).pin_memory()
    img_t = torch.from_numpy((rng.rand(tB, 3, H, W) * 255 - 115).astype(np.float32)).pin_memory()
    gts, labels, ignores = GI.make_gt(200 + rank, B, H, W, max_gt=20, max_ignore=5, with_ignore=True)
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
    for _ in range(max(args.warmup, 3)):
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
    traffic = ncu_traffic_per_launch()
    roofline = dict(bound="tensor", kernel="conv_igemm_kernel + conv_igemm_fast4_kernel (fprop + dgrad implicit GEMM, "
                                           "tcgen05)",
                    achieved=round(achieved, 1), peak=pk["tf_sust"], unit="TFLOP/s",
                    frac=round(achieved / pk["tf_sust"], 4), traffic=traffic["bytes"] if traffic else None,
                    traffic_unit="bytes/launch (DRAM read+write)", traffic_source=traffic["source"] if traffic else None,
                    peak_source=pk["src"] + " sustained",
                    launches_per_step=ck["n"], avg_launch_us=round(1e3 * ck["ms"] / max(ck["n"], 1), 2),
                    flops_per_step=ck["flops"], share_of_step=round(ck["ms"] / prof["step_ms"], 4),
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
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.cpu_sample_hw, args.depth, steps=3, warmup=1, batch=2, backbone=args.backbone)

    views = eager = None
    if world == 1 and not args.no_view_bench:
        views = view_images_side_bench()
    if world == 1 and not args.no_cpu_baseline and not literal:
        eager = eager_gpu_side_bench(args)

    line = dict(metric=METRIC, value=round(value, 2), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=round(ms_per_step, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=(WORKLOAD if args.backbone == "resnet" else WORKLOAD.replace(
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
                view_images=views, gpu_eager_baseline=eager)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
