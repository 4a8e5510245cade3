"""One-process-per-GPU plumbing for the batched path (torch.distributed; NCCL on GPUs,
gloo in the CPU tests).  Independent sequences are PARTITIONED over ranks: weights are
replicated, every rank owns the KV cache of its own sequences, and no collective touches the
data path -- only the tokens a driver wants to see globally and the timing reduction cross
ranks.  (RunState is per sequence, llama2.ts:147-163.)"""
import numpy as np


def partition(n_sequences, world, rank):
    """Contiguous, balanced slice [start, start+count) of the global batch owned by `rank`."""
    base, extra = divmod(n_sequences, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value, device="cpu"):
    """MAX-reduce of a scalar (step time): a multi-GPU number is the slowest rank's."""
    d = _dist()
    if d is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def gather_tokens(local_tokens, n_sequences, device="cpu"):
    """All ranks' next tokens in global sequence order (int32[n_sequences])."""
    d = _dist()
    local = np.ascontiguousarray(local_tokens, dtype=np.int32)
    if d is None:
        return local
    import torch
    world, rank = d.get_world_size(), d.get_rank()
    width = max(partition(n_sequences, world, r)[1] for r in range(world))
    buf = torch.full((width,), -1, dtype=torch.int32, device=device)
    buf[:local.size] = torch.from_numpy(local).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    d.all_gather(out, buf)
    parts = [o.cpu().numpy()[:partition(n_sequences, world, r)[1]] for r, o in enumerate(out)]
    return np.concatenate(parts).astype(np.int32)


class ShardedBatch:
    """Drives `step_fn(tokens, pos) -> next_tokens` (one rank's l2b_forward_batch / greedy
    step) over this rank's slice of a global batch and reassembles the global token matrix."""

    def __init__(self, n_sequences, step_fn, device="cpu"):
        d = _dist()
        self.world = d.get_world_size() if d else 1
        self.rank = d.get_rank() if d else 0
        self.n = n_sequences
        self.start, self.count = partition(n_sequences, self.world, self.rank)
        self.step_fn = step_fn
        self.device = device

    def run(self, first_tokens, n_steps):
        """first_tokens: int32[n_sequences] (global).  Returns int32[n_steps, n_sequences]."""
        tok = np.asarray(first_tokens, dtype=np.int32)[self.start:self.start + self.count].copy()
        out = np.empty((n_steps, self.n), dtype=np.int32)
        for s in range(n_steps):
            tok = np.asarray(self.step_fn(tok, np.full(self.count, s, np.int32)), dtype=np.int32)
            out[s] = gather_tokens(tok, self.n, self.device)
        return out


def connect_tp(ctx):
    """Exchange the ranks' handle blobs over the initialised process group and wire the
    tensor-parallel context (l2b_tp_export -> all_gather -> l2b_tp_connect)."""
    d = _dist()
    assert d is not None and d.get_world_size() == ctx.tp_size
    blobs = [None] * ctx.tp_size
    d.all_gather_object(blobs, ctx.tp_export())
    ctx.tp_connect(blobs)
    d.barrier()
