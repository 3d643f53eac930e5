"""N-pipelines-on-N-GPUs plumbing: one process per GPU, independent streams, no
collective on the data path.  torch.distributed is used only to line the ranks up and to take the
max over ranks of per-rank timings / sum of per-rank frame counts."""
import os


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None, device=None):
    """returns the torch.distributed module when WORLD_SIZE > 1, else None"""
    rank, world, local = rank_info()
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, **kw)
    return dist


def reduce_max(dist, values, device="cpu"):
    """max over ranks of a list of floats (device timings are reported as the slowest rank)"""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def reduce_sum(dist, values, device="cpu"):
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def stream_seed(rank, frame):
    """distinct synthetic content per (pipeline, frame): each rank is its own video stream"""
    return 1000 * rank + frame


def whole_job_throughput(units_per_rank, world, seconds_max):
    """weak scaling: every rank processes the same number of units; the job's rate is all units
    over the slowest rank's time"""
    return units_per_rank * world / seconds_max
