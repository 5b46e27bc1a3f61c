"""The three collectives of the data-parallel step, one process per GPU over torch.distributed:
  1. grad mean all-reduce          (reference: DDP bucket all-reduce, mmdet/apis/train.py:88-102)
  2. ONE packed 2-scalar all-reduce for num_pos and sum(centerness targets)
                                   (reference: two reduce_mean calls, fcos_head.py:266,274; dist_utils.py:63-69)
  3. packed log-var all-reduce     (reference: one all-reduce + .item() per key, detectors/base.py:201-206)
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _native_avg():
    """NCCL reduces with ncclAvg in the collective itself; gloo has no AVG op (pre-scale there)."""
    try:
        return dist.get_backend() == "nccl"
    except Exception:   # noqa: BLE001
        return False


def allreduce_mean_(t):
    """In place: t <- mean over ranks (what DDP leaves in .grad). On NCCL the 1/world factor rides inside the
    collective (ReduceOp.AVG): no separate pass over the 128 MB gradient."""
    w = world_size()
    if w > 1:
        if _native_avg():
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            t.div_(w)
            dist.all_reduce(t)
    return t


def allreduce_mean_async_(t):
    """Same as allreduce_mean_ but returns the in-flight work handle (None with one rank): call .wait() before using t."""
    w = world_size()
    if w > 1:
        if _native_avg():
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True)
        t.div_(w)
        return dist.all_reduce(t, async_op=True)
    return None


def allreduce_sum_async_(t):
    """In-flight sum all-reduce (None with one rank): .wait() before the result is consumed."""
    return dist.all_reduce(t, async_op=True) if world_size() > 1 else None


def allreduce_sum_(t):
    if world_size() > 1:
        dist.all_reduce(t)
    return t


def normalisers_from_counts(counts_sum, w):
    """(num_pos, sum ctr-targets) summed over ranks -> the reference's normalisers: max(reduce_mean(num_pos), 1.0) and
    max(reduce_mean(sum_ctr), 1e-6) (fcos_head.py:266,273-274). Host restatement of dslb_fcos_norm for the tests."""
    c = counts_sum.to(torch.float64) / float(w)
    return torch.stack([torch.clamp(c[0], min=1.0), torch.clamp(c[1], min=1e-6)]).to(torch.float32)


def reduce_log_vars(log_vars):
    """OrderedDict of 0-dim tensors -> OrderedDict of python floats, world-averaged with ONE all-reduce."""
    keys = list(log_vars.keys())
    vals = torch.stack([log_vars[k].detach().to(torch.float32) for k in keys])
    w = world_size()
    if w > 1:
        vals = vals.clone()
        dist.all_reduce(vals.div_(w))
    return type(log_vars)(zip(keys, vals.tolist()))


def shutdown(*engines, timeout=20.0):
    """End of a multi-rank run. CUDA graphs that captured NCCL work (DSLEngine's one-graph step) must be released BEFORE
    the communicator: ProcessGroupNCCL's destructor otherwise waits forever on work it can no longer see complete
    (measured: both ranks stuck in destroy_process_group). Destroys the graphs, synchronises, barriers, and tears the
    process group down under a watchdog — if NCCL still refuses to die the process exits cleanly instead of hanging."""
    import gc
    import os
    import sys
    import threading
    for e in engines:
        if e is not None:
            e.graphs = None
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not (dist.is_available() and dist.is_initialized()):
        return
    try:
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    except Exception:   # noqa: BLE001
        pass
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout)
    if t.is_alive():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
