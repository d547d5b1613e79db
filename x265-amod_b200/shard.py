"""Multi-GPU plumbing for the lookahead: the path shards by STREAM (one independent Lookahead per rank, no
data-path collective), so all that crosses ranks is the barrier and the max-over-ranks step time.  Backend-agnostic
(NCCL on the GPU box, gloo in the CPU tests)."""
import os


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def streams_for_rank(n_streams, rank, world):
    """Independent streams (renditions / sequences) are dealt round-robin; each is owned by exactly one rank."""
    return [s for s in range(n_streams) if s % world == rank]


def init(backend):
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(backend)
    return dist


def reduce_max(dist, value, device="cpu"):
    """max over ranks of a scalar (the step time that bounds the job)"""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, value, device="cpu"):
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(dist, frames_this_rank, ms_this_rank, device="cpu"):
    """whole-job frames/s = frames of all ranks / slowest rank's time"""
    total = reduce_sum(dist, frames_this_rank, device)
    ms = reduce_max(dist, ms_this_rank, device)
    return total / (ms / 1000.0), ms
