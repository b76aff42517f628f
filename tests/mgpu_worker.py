"""One rank of the multi-GPU correctness run (launched by tests/test_multi_gpu.py through torch.distributed.run).

Every rank trains `rounds` GpuTrainer rounds over its shard of a global population of `games` games; rank 0 saves the
weights after every round and per-round counters.  The same script with WORLD_SIZE=1 is the single-GPU run the others
must reproduce: games are keyed by global ids, so plies, winners and TD steps are IDENTICAL, and the weights agree up to
the order in which fp32 per-game changes are summed (per CTA, per rank, across ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "backgammon-engine_b200"))


def main():
    out, games, rounds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bgx.model import TDLGammonModel
    from bgx.train import GpuTrainer
    torch.manual_seed(5)
    m = TDLGammonModel()
    tr = GpuTrainer(m, games // world, device=local, delta_scale=1.0 / games)
    used_c_abi = tr.comm is not None
    weights, counters, deltas = [], [], []
    for _ in range(rounds):
        st = tr.round()
        cnt = torch.tensor([st["plies"], st["td_steps"], st["p1_wins"], st["games_finished"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cnt)
        counters.append(cnt.cpu().numpy())
        deltas.append(tr.delta.cpu().numpy().copy())
        weights.append(np.concatenate([np.asarray(a).reshape(-1) for a in tr.eng.get_weights()]))
    if int(os.environ.get("RANK", "0")) == 0:
        np.savez(out, weights=np.array(weights), counters=np.array(counters), deltas=np.array(deltas), c_abi=np.int8(used_c_abi), world=np.int32(world))
    tr.eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
