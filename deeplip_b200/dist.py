"""Multi-GPU plumbing: one process per GPU, utterances sharded with no data-path
collective during extraction, ONE all_gather of the fused embeddings before scoring, trial lines
sharded across ranks, scores gathered for the CPU EER step.  The reference has no distributed code
(only single-process nn.DataParallel, unwrapped before extraction: train_fusion.py:318-328).
Device-agnostic: runs on NCCL/CUDA in production and on gloo/CPU tensors in the unit tests.
"""
import os
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_size(n, world):
    return -(-n // world)


def shard_range(n, rank, world):
    """Contiguous block of utterance indices owned by `rank` (equal padded shards)."""
    per = shard_size(n, world)
    return min(rank * per, n), min((rank + 1) * per, n)


def all_gather_rows(local_rows, n_total, rank, world):
    """local_rows: (n_local, D) rows [lo,hi) of a (n_total, D) table -> the full table on every rank.
    Shards are padded to equal size so a single all_gather_into_tensor moves everything."""
    if world == 1:
        return local_rows
    per = shard_size(n_total, world)
    D = local_rows.shape[1]
    send = local_rows
    if local_rows.shape[0] != per:
        send = torch.zeros((per, D), device=local_rows.device, dtype=local_rows.dtype)
        send[:local_rows.shape[0]] = local_rows
    full = torch.empty((per * world, D), device=local_rows.device, dtype=local_rows.dtype)
    dist.all_gather_into_tensor(full, send.contiguous())
    return full[:n_total]


class OverlappedGather:
    """all_gather_rows of one step's rows on a communication stream of its own, so that the collective (and its wait
    for the slowest rank) overlaps the NEXT step's kernels instead of holding the compute stream: with one collective per
    step on the compute stream the ranks run in lockstep, and every step costs the slowest rank's time plus the skew.
    Call it with the rows (valid on the current stream); the result is valid on `.stream` -- consume it there (or after
    `.wait()`, which makes the current stream wait for it).  CUDA only; world == 1 is the identity."""

    def __init__(self, n_total, rank, world, device):
        self.n_total, self.rank, self.world = n_total, rank, world
        self.stream = torch.cuda.Stream(device=device) if world > 1 else None
        self._ev = torch.cuda.Event() if world > 1 else None
        self._done = torch.cuda.Event() if world > 1 else None

    def __call__(self, rows):
        if self.world == 1:
            return rows
        main = torch.cuda.current_stream(rows.device)
        self._ev.record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ev)
            rows.record_stream(self.stream)             # the caching allocator must not recycle it under the collective
            out = all_gather_rows(rows, self.n_total, self.rank, self.world)
            self._done.record(self.stream)
        return out

    def wait(self):
        if self.world > 1:
            torch.cuda.current_stream().wait_event(self._done)


def gather_scores(local_scores, n_total, rank, world):
    """Concatenate per-rank score slices (contiguous trial shards) into the full vector on every rank."""
    if world == 1:
        return local_scores
    per = shard_size(n_total, world)
    send = torch.zeros((per,), device=local_scores.device, dtype=local_scores.dtype)
    send[:local_scores.numel()] = local_scores
    full = torch.empty((per * world,), device=local_scores.device, dtype=local_scores.dtype)
    dist.all_gather_into_tensor(full, send)
    return full[:n_total]


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


_HOST_GROUP = None


def host_barrier():
    """Barrier that BLOCKS on sockets (a gloo group created on first use) instead of spinning: ranks that wait for
    rank 0's CPU work (the EER, the oracle check) in an NCCL barrier each burn a core in cudaStreamSynchronize, and an
    OpenMP team sized to all cores then collapses (measured: the 43-utterance oracle sample 2 s alone, 65 s next to
    three spinning ranks).  Collective: every rank must call it the same number of times."""
    global _HOST_GROUP
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    if dist.get_backend() == 'gloo':
        dist.barrier()
        return
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend='gloo')
    dist.barrier(group=_HOST_GROUP)
