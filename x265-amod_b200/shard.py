"""Multi-GPU plumbing for the lookahead.  The path shards at two levels:
  * by STREAM: one independent Lookahead per rank, no data-path collective -- all that crosses ranks is the barrier
    and the max-over-ranks step time (bench.py's default under torchrun);
  * inside ONE stream: the ranks split the searches / estimates by source frame and exchange the stores they wrote
    at the end of every batch (make_exchange: one broadcast per owner, NCCL over NVLink on the GPU box, gloo in the CPU
    tests), so every rank takes the same decisions and rank 0's output is the product.
Backend-agnostic."""
import ctypes as C
import os


class _DevBuf:
    """n bytes of device memory at `ptr` as a CUDA-array-interface object torch can alias"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def make_exchange(dist, exchange_fn_type, cuda):
    """The engine's exchange callback (include/x265cu.h, x265cu_exchange_fn) on top of torch.distributed: for every
    root with bytes to send, broadcast its buffer to all ranks, ordered on the engine's stream."""
    import torch

    def exchange(user, bufs, nbytes, nranks, stream):
        try:
            if cuda:
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    for r in range(nranks):
                        if nbytes[r]:
                            t = torch.as_tensor(_DevBuf(bufs[r], int(nbytes[r])), device="cuda")
                            dist.broadcast(t, src=r)
            else:
                for r in range(nranks):
                    if nbytes[r]:
                        raw = (C.c_ubyte * int(nbytes[r])).from_address(bufs[r])
                        dist.broadcast(torch.frombuffer(raw, dtype=torch.uint8), src=r)
            return 0
        except Exception as e:      # never let an exception cross the C boundary
            import sys
            print("exchange failed: %r" % (e,), file=sys.stderr)
            return 1
    return exchange_fn_type(exchange)


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def streams_for_rank(n_streams, rank, world):
    """Independent streams (renditions / sequences) are dealt round-robin; each is owned by exactly one rank."""
    return [s for s in range(n_streams) if s % world == rank]


def init(backend):
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend == "nccl":
            import torch
            dist.init_process_group(backend, device_id=torch.device("cuda", rank_info()[2]))
        else:
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
