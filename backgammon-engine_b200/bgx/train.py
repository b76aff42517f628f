"""Self-play and TD(lambda) training: the reference's train.py entry points on the compat API,
plus the population trainer that runs whole rounds on the GPU.

The reference's own `play_game` / `apply_td_updates` (train.py:64-121, 124-172) run unmodified on the
compat module and on `TDLGammonModel`; their torch restatement used as a checker lives in tests/ref_td.py.
`play_games_batch` is `play_game` for a whole population on the GPU.
`GpuTrainer` replaces the multiprocessing pool + sequential replay (train.py:305-354, 527-547):
one round = every game of the population played to its end from one weight snapshot
(k_selfplay), every trajectory replayed with exact online TD(lambda) from that snapshot
(k_td_replay), the per-game weight changes summed, all-reduced over the ranks (NCCL) and
applied.  Combining per-game updates by (scaled) summation instead of applying them one after
the other is the one deliberate divergence from the reference.
"""
import numpy as np
import torch

from . import lib as L
from .parallel import NcclComm, allreduce_delta, shard


def play_games_batch(model, n_games, epsilon=0.0, seed=0x5EED2026, device=0, first_id=0):
    """`play_game` for n_games at once on the GPU: one self-play round from the model's current weights.
    Returns a list of (winner, states, total_moves) with the reference's types (train.py:64-121): winner 0/1,
    states = the pre-move encodings float32[198] of every ply, total_moves = len(states) - 1 — ready for
    `apply_td_updates`.  Dice are Philox(seed, ply, first_id + i); opening roll-off as train.py:89-97."""
    eng = model.engine(device)
    eng.selfplay_init(n_games, first_id=first_id, id_stride=n_games, seed=seed, first_mover=L.FIRST_ROLLOFF, traj_cap=2048)
    st = eng.selfplay_round(epsilon)
    if st["truncated"]:
        raise RuntimeError(f"{st['truncated']} games exceeded 2048 plies")
    rec, ply, _ = eng.selfplay_read()
    games = []
    for slot in range(n_games):
        pre, _ = eng.export_trajectory(slot)
        X = eng.encode_host(pre)                                     # _encode_states_np with the mover's flag
        games.append((int(rec[slot, 31]) - 1, [x for x in X], len(pre) - 1))
    return games


class GpuTrainer:
    """Round-structured self-play TD(lambda) on one GPU per process (rank = torch.distributed rank).

    games_per_rank concurrent games; game ids are global (rank * games_per_rank + slot, stride
    world * games_per_rank) so a run is invariant to how the population is sharded.

    The trainer owns a PRIVATE engine: the live weights are the engine's, the torch module is a snapshot that
    `sync_model()` refreshes (save_checkpoint does it by itself), and nothing the module's own engine() does - batched
    make_move, play_games_batch - can disturb the trainer's population or weights.

    lr / lambda follow the reference PER GAME (train.py:538: update_learning_params(games_done + k + 1) before the
    k-th game of the round): the replay kernel looks the schedule of model.py:69-73 up for every game.  What a parallel
    round cannot reproduce is the reference applying the games' updates one after the other; a round applies
    delta_scale x their sum (default: the mean).  One round is ONE step of that size, so the population is a batch
    size: a few hundred games per round train fastest per game played; a 262,144-game round is the
    throughput configuration of BASELINE.json configs[3], not a training recipe - a warning says so."""

    def __init__(self, model, games_per_rank, device=0, seed=0x5EED2026, traj_cap=2048,
                 first_mover=L.FIRST_ROLLOFF, delta_scale=None, schedule="per_game"):
        import warnings
        import torch.distributed as dist
        from .engine import BatchEngine
        if schedule not in ("per_game", "per_round"):
            raise ValueError("schedule is 'per_game' (train.py:538) or 'per_round'")
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.model = model
        self.schedule = schedule
        self.games_per_rank = int(games_per_rank)
        self.global_games = self.games_per_rank * self.world
        if self.global_games > 8192:
            warnings.warn(f"GpuTrainer: {self.global_games:,} games per round means one weight update per {self.global_games:,} games, and the "
                          "reference's lr / lambda schedule (periods 40,000 / 30,000 games) advances by whole periods per update; "
                          "use a few hundred games per round for training and keep rounds this large for throughput runs", stacklevel=2)
        # default: the MEAN of the per-game updates (mini-batch TD); the reference applies them in sequence
        self.delta_scale = (1.0 / self.global_games) if delta_scale is None else float(delta_scale)
        self.eng = BatchEngine(device)
        self.eng.set_weights(*model.weights_np())
        self.eng.set_stream(torch.cuda.current_stream().cuda_stream)
        first_id, n_slots, stride = shard(self.global_games, self.rank, self.world)
        self.eng.selfplay_init(n_slots, first_id=first_id, id_stride=stride, seed=seed, first_mover=first_mover, traj_cap=traj_cap)
        self.delta = torch.zeros(L.NPARAMS_PADDED, dtype=torch.float32, device=torch.device("cuda", device))
        # the exchange goes through the C-ABI (bgx_allreduce_delta on an NCCL communicator made by bgx_nccl_comm_init)
        self.comm = NcclComm(self.eng, self.dist) if self.dist and self.world > 1 and self.dist.get_backend() == "nccl" else None
        self.games_done = 0
        self.rounds = 0

    def round(self, epsilon=0.0):
        """Play, replay, reduce, apply.  Returns the round's statistics (this rank's counts)."""
        if self.rounds:
            self.eng.selfplay_next_round()
        play = self.eng.selfplay_round(epsilon)
        if self.schedule == "per_game":
            td = self.eng.td_replay_scheduled(self.games_done, self.delta)
            self.model.update_learning_params(self.games_done + self.global_games)    # what the reference's model holds after the round
        else:
            self.model.update_learning_params(self.games_done + 1)
            td = self.eng.td_replay(self.model.learning_rate, self.model.lambda_decay, self.delta)
        if self.comm:                                                        # the only cross-GPU traffic: 102,416 B
            self.comm.allreduce_delta(self.delta)
        else:
            allreduce_delta(self.delta, self.dist)
        self.eng.apply_delta(self.delta, self.delta_scale)
        self.games_done += self.global_games
        self.rounds += 1
        self.model._ahead = self
        return {**play, "td_steps": td["td_steps"], "td_sq_error": td["td_sq_error"], "games": self.global_games}

    def sync_model(self):
        """The engine's weights -> the torch module (state_dict, checkpoints, CPU evaluation)."""
        self.model.load_weights_from(self.eng)
        self.model._ahead = None
        return self.model

    def close(self):
        self.sync_model()
        if self.comm:
            self.comm.close()
        self.eng.close()


# ------------------------------------------------------------------ checkpoints (train.py:361-381, 513-515)

def model_compatible(path):
    """True iff the checkpoint loads into the current TDLGammonModel architecture (train.py:361-368)."""
    from .model import TDLGammonModel
    try:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        TDLGammonModel().load_state_dict(sd)
        return True
    except Exception:
        return False


def latest_compatible_model(models_dir):
    """Most recently modified .pth in `models_dir` that fits the 198-128-1 architecture, or None; the load
    test skips the reference's old 3-layer checkpoints (train.py:371-381)."""
    import os
    if not os.path.isdir(models_dir):
        return None
    files = [f for f in os.listdir(models_dir) if f.endswith(".pth")]
    files.sort(key=lambda f: os.path.getmtime(os.path.join(models_dir, f)), reverse=True)
    for f in files:
        if model_compatible(os.path.join(models_dir, f)):
            return f
    return None


def save_checkpoint(model, path):
    """The reference's checkpoint format: the 4-tensor state_dict (train.py:513-515).  If a GpuTrainer holds newer weights
    than the module, they are pulled into the module first."""
    if getattr(model, "_ahead", None) is not None:
        model._ahead.sync_model()
    torch.save(model.state_dict(), path)
