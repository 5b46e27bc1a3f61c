"""N > 1 host logic on CPU: two gloo ranks exercise the three collectives of the data-parallel step (dist_ops.py) and
check them against what the reference's per-rank code computes (DDP mean of grads; reduce_mean of the two loss
normalisers, mmdet/core/utils/dist_utils.py:63-69; per-key log-var averaging, mmdet/models/detectors/base.py:201-206)."""
import os
import socket
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsl_b200 import dist_ops
    from oracle import fcos_oracle as O
    from tests.golden import inputs as GI
    try:
        assert dist_ops.world_size() == world
        # 1. gradient mean (per-rank seed = seed + rank, as bench.py feeds the ranks)
        g = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(1000, generator=g)
        mine = grad.clone()
        dist_ops.allreduce_mean_(grad)
        gathered = [torch.zeros(1000) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert torch.allclose(grad, torch.stack(gathered).mean(0), atol=1e-7)
        # 1b. bucketed form used by the engine at world > 1: asynchronous mean all-reduces of disjoint slices of the flat
        #     gradient (issued as the backward finishes each bucket) == one all-reduce of the whole buffer
        flat = torch.randn(1000, generator=torch.Generator().manual_seed(200 + rank))
        whole = flat.clone()
        works = [dist_ops.allreduce_mean_async_(flat[lo:hi]) for lo, hi in ((600, 1000), (250, 600), (0, 250))]
        for wk in works:
            wk.wait()
        dist_ops.allreduce_mean_(whole)
        assert torch.equal(flat, whole)
        # 2. packed normalisers: each rank runs the reference's target assignment on ITS images (oracle), the packed
        #    sum over ranks must give the same normalisers as the reference's two reduce_mean calls
        B, H, W = 2, 128, 160
        gts, labels, _ = GI.make_gt(7 + rank, B, H, W, with_ignore=False)
        sizes = GI.level_sizes(H, W)
        pts = O.get_points(sizes, GI.STRIDES)
        lab, tgt = O.get_targets(pts, gts, labels, GI.STRIDES, GI.REGRESS_RANGES, 80, center_sampling=True,
                                 norm_on_bbox=True)
        lab, tgt = torch.cat(lab), torch.cat(tgt)
        pos = (lab >= 0) & (lab < 80)
        num_pos = float(pos.sum())
        sum_ctr = float(O.centerness_target(tgt[pos]).sum()) if num_pos > 0 else 0.0
        counts = torch.tensor([num_pos, sum_ctr], dtype=torch.float64)
        dist_ops.allreduce_sum_(counts)
        norm = dist_ops.normalisers_from_counts(counts, world)
        # reference: reduce_mean(tensor) = all_reduce(tensor / world) (dist_utils.py:63-69), then max(., 1.0) / max(., 1e-6)
        a = torch.tensor(num_pos, dtype=torch.float32).div_(world)
        b = torch.tensor(sum_ctr, dtype=torch.float32).div_(world)
        dist.all_reduce(a)
        dist.all_reduce(b)
        assert abs(float(norm[0]) - max(float(a), 1.0)) <= 1e-6 * max(float(a), 1.0)
        assert abs(float(norm[1]) - max(float(b), 1e-6)) <= 1e-6 * max(float(b), 1e-6)
        # 3. log vars: one packed all-reduce == the reference's per-key all-reduce
        lv = OrderedDict(loss_cls=torch.tensor(1.0 + rank), loss_bbox=torch.tensor(0.5 * (rank + 1)),
                         loss=torch.tensor(3.0 - rank))
        red = dist_ops.reduce_log_vars(lv)
        for k, v in lv.items():
            t = v.clone()
            dist.all_reduce(t.div_(world))
            assert abs(red[k] - t.item()) < 1e-7
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_collectives_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _rla_bucket_worker(rank, world, port, q):
    """Data-parallel backward of the RLA_ResNet backbone plan (executed on tests/emu_lib.py): the gradient leaves in the
    plan's buckets — each all-reduced asynchronously as soon as the ops up to its boundary have run — and must equal a
    full backward followed by ONE all-reduce (the reference's DDP result, mmdet/apis/train.py:88-102), on both ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from dsl_b200 import dist_ops
    from dsl_b200.engine import FCOSNet
    from dsl_b200.params import ParamStore, rla_resnet_spec
    from tests import emu_lib
    from tests.golden import inputs as GI
    try:
        torch.set_num_threads(2)
        B, H, W = 1, 64, 96
        x = GI.make_tensor(np.random.RandomState(300 + rank), B, 3, H, W)      # per-rank images, shared weights
        rng = np.random.RandomState(400 + rank)
        with emu_lib.installed():
            store = ParamStore(rla_resnet_spec(prefix=""), "cpu")
            store.load_state_dict(GI.rla_state_dict(51))
            net = FCOSNet(B, H, W, depth=50, train=True, store=store, device="cpu", parts="backbone", backbone="rla")
            net.img.copy_(x)
            net.forward()
            seeds = [torch.from_numpy(rng.randn(*g.shape).astype(np.float32)).bfloat16() for g in net.gc]

            def seed():
                for g, s in zip(net.gc, seeds):
                    g.copy_(s)

            # bucketed: ops [start, end) of each bucket, then its flat-gradient range goes on the wire
            seed()
            start, works = 0, []
            assert len(net.bwd_buckets) == 2
            for end, lo, hi in net.bwd_buckets:
                net.backward(start=start, end=end)
                works.append(dist_ops.allreduce_mean_async_(net.grad[lo:hi]))
                start = end
            assert start == len(net.bwd_ops)        # nothing is left behind the last bucket
            for wk in works:
                wk.wait()
            bucketed = net.grad.clone()
            # one all-reduce after the whole backward (the gradient maps are consumed in place: seed again)
            seed()
            net.backward()
            local = net.grad.clone()
            dist_ops.allreduce_mean_(net.grad)
            assert torch.equal(bucketed, net.grad)
            both = [torch.zeros_like(local) for _ in range(world)]
            dist.all_gather(both, local)
            assert not torch.equal(both[0], both[1])                       # the ranks really saw different images
            assert torch.allclose(net.grad, (both[0] + both[1]) / 2, rtol=1e-6, atol=1e-9)
            o, n = store.offsets["stages.2.0.bn1.weight"]                   # trainable BatchNorm affines travel with the bucket
            assert float(net.grad[o:o + n].abs().sum()) > 0
            del net
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


def test_two_rank_rla_gradient_buckets_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rla_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
