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


def split_weights(flat):
    """flat fp32[>=25601] in state_dict order -> (W1[128,198], b1[128], w2[1,128], b2[1])"""
    flat = flat[:NPARAMS]
    return flat[:25344].reshape(128, 198), flat[25344:25472], flat[25472:25600].reshape(1, 128), flat[25600:25601]
