"""Multi-GPU plumbing: how a global self-play population is sharded over ranks and how the
per-rank TD weight deltas are combined.  Pure host logic over torch.distributed, so the same
code runs under NCCL on GPUs and under gloo in the CPU test-suite.

Sharding rule: rank r of R owns slots [r*G/R, (r+1)*G/R); slot s plays global
game ids s, s+G, s+2G, ...; Philox dice are keyed by the global id, so every game is the same
game whatever R is.  The only exchange is one all-reduce(sum) of fp32[25,604] per round.
"""
import torch

from .lib import NPARAMS, NPARAMS_PADDED


def shard(global_games, rank, world):
    """-> (first_id, n_slots, id_stride) for bgx_selfplay_init on this rank."""
    if global_games % world:
        raise ValueError(f"global population {global_games} is not a multiple of world size {world}")
    per = global_games // world
    return rank * per, per, global_games


def allreduce_delta(delta, dist=None):
    """Sum the per-rank weight deltas in place (NCCL on CUDA tensors, gloo on CPU tensors)."""
    if delta.numel() != NPARAMS_PADDED or delta.dtype != torch.float32:
        raise ValueError("delta must be fp32[25604]")
    if dist is None:
        import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    return delta


def torch_nccl_path():
    """The libnccl that torch itself loaded (the wheel bundles its own); None if it cannot be found."""
    import glob
    import os
    import sys
    for base in sys.path:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return sorted(hits)[0]
    return None


class NcclComm:
    """One NCCL communicator made through libbgx's C-ABI (bgx_nccl_*), for bgx_allreduce_delta.  torch.distributed is used
    only to hand rank 0's 128-byte unique id to the other ranks - any transport would do."""

    def __init__(self, engine, dist):
        import ctypes as C
        from . import lib as L
        self._lib, self._eng, self._h = L.load(), engine, C.c_void_p()
        path = torch_nccl_path()
        L.check(self._lib.bgx_nccl_load(path.encode() if path else None))
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            L.check(self._lib.bgx_nccl_unique_id(ident.numpy().ctypes.data))
        ident = ident.cuda() if dist.get_backend() == "nccl" else ident
        dist.broadcast(ident, 0)
        ident = ident.cpu().contiguous()
        L.check(self._lib.bgx_nccl_comm_init(engine._h, world, rank, ident.numpy().ctypes.data, C.byref(self._h)))

    def allreduce_delta(self, delta):
        """Sum fp32[25604] over the ranks in place on the engine's stream (bgx_allreduce_delta)."""
        from . import lib as L
        if delta.numel() != NPARAMS_PADDED or delta.dtype != torch.float32 or not delta.is_cuda:
            raise ValueError("delta must be a CUDA fp32[25604] tensor")
        L.check(self._lib.bgx_allreduce_delta(self._eng._h, self._h, L.ptr(delta)))
        return delta

    def close(self):
        if self._h:
            self._lib.bgx_nccl_comm_destroy(self._h)
            self._h = None


def split_weights(flat):
    """flat fp32[>=25601] in state_dict order -> (W1[128,198], b1[128], w2[1,128], b2[1])"""
    flat = flat[:NPARAMS]
    return flat[:25344].reshape(128, 198), flat[25344:25472], flat[25472:25600].reshape(1, 128), flat[25600:25601]
