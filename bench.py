"""bench.py — self-play plies/s of the B200-native engine (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm
    python bench.py --impl reference [--steps K] [--warmup W]     # the reference's CPU path

Workload (BASELINE.json configs[2]): greedy 1-ply self-play with the 198-128-1 TD-Gammon net,
65,536 concurrent games per GPU, random-init weights (TDLGammonModel() after torch.manual_seed(0)),
epsilon = 0, Philox dice, first mover g % 2, finished games restart in place.  One STEP = every
game of the population plays PLIES_PER_STEP plies (one k_selfplay launch).

  value      plies/s with the population resident in HBM (CUDA events on the launching stream)
  e2e        plies/s through the C-ABI with HOST buffers: every ply the 65,536 (position, dice) records, ply
             numbers and game ids go pinned-host -> device, one iteration of play_game's loop runs for all of
             them (make_move, is_game_over, setTurn, roll_dice: bgx_play_ply_host_async), the next records and
             the winners come back, and the host restarts finished games
  roofline   the self-play kernel against the measured HBM peak, in the units SURVEY.md §8(d)
             prescribes: 1,696 algorithmic bytes per enumerated afterstate
  cpu_baseline  the reference engine + model.py loop on the host cores (oracle/ref_play.py)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "backgammon-engine_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GAMES_PER_GPU = 65536
PLIES_PER_STEP = 16
SEED = 0x5EED2026
BYTES_PER_AFTERSTATE = 1696          # 792 B X written + 792 B X read + 112 B int32[28] row
FLOP_PER_AFTERSTATE = 50944          # dense 198-128-1 forward, as the reference executes it
METRIC = "selfplay_plies_per_sec"
UNIT = "plies/s"


def workload_config(n_gpus):
    return {"workload": "greedy 1-ply self-play, 198-128-1 net random-init (seed 0), eps=0, Philox dice, "
                        "first mover g%2, in-place restart (BASELINE.json configs[2])",
            "games_per_gpu": GAMES_PER_GPU, "plies_per_step": PLIES_PER_STEP, "n_gpus": n_gpus,
            "parallelism": f"games sharded over {n_gpus} GPU(s), no data-path collective; one process per GPU, pinned to the GPU's local CPUs",
            "l2": "256 MiB scratch written between timed steps (L2 flush); population 2 MiB"}


def init_weights():
    """TDLGammonModel() after torch.manual_seed(0): Xavier-uniform gain 0.1, zero biases (model.py:33-61)."""
    import torch
    torch.manual_seed(0)
    fc1 = torch.nn.Linear(198, 128)
    fc2 = torch.nn.Linear(128, 1)
    for m in (fc1, fc2):
        torch.nn.init.xavier_uniform_(m.weight, gain=0.1)
        torch.nn.init.constant_(m.bias, 0)
    return tuple(t.detach().numpy().copy() for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias))


# ----------------------------------------------------------------------------- reference arm

def peak_json():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """The reference's own CPU implementation of the path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_play
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.cpu_procs or cores))
    w = init_weights()
    # every step is a bounded sample of the workload; the whole run stays within ~3 minutes whatever K is
    args.cpu_seconds = max(1.0, min(args.cpu_seconds, 150.0 / max(args.steps, 1)))
    if ref_play.available():
        kind = "reference"
        for _ in range(args.warmup):
            ref_play.run(w, min(2.0, args.cpu_seconds), procs)
        tot_p = tot_s = tot_t = 0.0
        for _ in range(args.steps):
            r = ref_play.run(w, args.cpu_seconds, procs)
            tot_p += r["plies"]; tot_s += r["sequences"]; tot_t += r["seconds"]
        sample = (f"reference backgammon_env (oracle/_ref, g++ -O2) + model.py make_move loop (numpy encode, torch CPU "
                  f"1 thread/process), {procs} processes x {args.cpu_seconds:g} s per step, greedy self-play from the opening")
    else:
        # the reference did not compile here (no /root/reference at build time): time the oracle port
        kind = "port"
        tot_p, tot_s, tot_t = port_baseline(w, args.cpu_seconds * max(args.steps, 1))
        procs = 1
        sample = f"oracle/bgx_oracle.c greedy self-play, 1 thread, {tot_t:.1f} s"
    value = tot_p / tot_t
    direct = cpp_direct_sample()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "sequences_per_sec": tot_s / tot_t,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample, "enumeration_only": direct},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpp_direct_sample(n=1500):
    """SURVEY 8d's C++-direct comparator on ONE core, enumeration only (no Python, no model): the reference's own
    evaluateTurnSequences (oracle/_ref, ~97 % of it mt19937_64 construction in Game::clone) and the oracle port
    (the same algorithm without that cost = the honest CPU ceiling), on the first n positions of the configs[1] sweep."""
    import numpy as np
    from bgx.synth import make_queries
    from oracle.oracle import Oracle, RefHarness
    q, _ = make_queries(n, seed=20260101)
    out = {"positions": n, "cores": 1}
    try:
        orc = Oracle()
        t0 = time.perf_counter()
        ns, _, _ = orc.turn_summary_batch(q, threads=1)
        out["oracle_port_sequences_per_sec"] = float(ns.sum()) / (time.perf_counter() - t0)
        if RefHarness.available():
            ref = RefHarness()
            tot = np.zeros(1, np.int64)
            import ctypes
            secs = ref.lib.ref_bench_enumerate(np.ascontiguousarray(q[:, :28].astype(np.int32)).reshape(-1), np.ascontiguousarray(q[:, 28]),
                                               np.ascontiguousarray(q[:, 29:31]).reshape(-1), n, tot.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
            out["reference_sequences_per_sec"] = float(tot[0]) / secs
    except Exception as exc:
        out["error"] = str(exc)
    return out


def port_baseline(w, seconds):
    import numpy as np
    from oracle.oracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(1)
    from bgx.synth import START_BOARD
    plies = seqs = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        s = np.zeros(28, np.int32)
        s[:24] = START_BOARD
        pl = 0
        while time.perf_counter() - t0 < seconds:
            d1, d2 = (int(x) for x in rng.integers(1, 7, 2))
            idx, after, v, n = orc.greedy_ply(w, s, pl, d1, d2)
            plies += 1
            seqs += n
            if idx >= 0:
                s = after
            if orc.game_over(s) >= 0:
                break
            pl ^= 1
    return plies, seqs, time.perf_counter() - t0


# ----------------------------------------------------------------------------- our arm

class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def side_legs(eng, dev, peaks, n_positions):
    """The other kernels of the path on BASELINE.json configs[1]'s input (seeded synthetic positions incl. bar
    entry, doubles and bear-off), device-resident, CUDA-event timed inside the library: enumeration only,
    enumeration + evaluation + arg-best, the unfused encoder (the one truly HBM-bound kernel) and the evaluator."""
    import numpy as np
    import torch
    from bgx.synth import make_queries
    q, _ = make_queries(n_positions, seed=20260101)
    qd = torch.from_numpy(q).to(dev)
    n = qd.shape[0]
    n_seq = torch.zeros(n, dtype=torch.int32, device=dev)
    n_uni = torch.zeros(n, dtype=torch.int32, device=dev)
    dig = torch.zeros(n, dtype=torch.int64, device=dev)

    def best_ms(fn, reps=3):
        fn()
        ms = []
        for _ in range(reps):
            fn()
            ms.append(eng.last_kernel_ms())
        return min(ms)

    t_enum = best_ms(lambda: eng.enumerate_summary(qd, n_seq, n_uni, dig)) * 1e-3
    seqs, uniq = int(n_seq.sum().item()), int(n_uni.clamp(min=0).sum().item())
    chosen = torch.zeros((n, 32), dtype=torch.int8, device=dev)
    val = torch.zeros(n, dtype=torch.float32, device=dev)
    t_sel = best_ms(lambda: eng.select_moves(qd, chosen=chosen, value=val)) * 1e-3
    # the materialised list (batched evaluateTurnSequences): every sequence's moves, length and 32-byte state row
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(n_seq.to(torch.int64), 0)
    mv = torch.empty((seqs, 8), dtype=torch.int8, device=dev)
    ln = torch.empty(seqs, dtype=torch.int8, device=dev)
    st = torch.empty((seqs, 32), dtype=torch.int8, device=dev)
    t_wr = best_ms(lambda: eng.enumerate(qd, offs, mv, ln, st)) * 1e-3
    del mv, ln, st
    rows = 4 * n
    qrep = qd.repeat(4, 1)
    X = torch.empty((rows, 198), dtype=torch.float32, device=dev)
    t_enc = best_ms(lambda: eng.encode(qrep, X)) * 1e-3
    enc_gbs = rows * (32 + 792) / t_enc / 1e9
    V = torch.zeros(n, dtype=torch.float32, device=dev)
    t_ev = best_ms(lambda: eng.evaluate(qd, V)) * 1e-3
    return {"positions": n, "sequences": seqs, "unique_afterstates": uniq,
            "enumerate": {"kernel": "k_enumerate_summary", "ms": t_enum * 1e3, "positions_per_sec": n / t_enum,
                          "sequences_per_sec": seqs / t_enum, "unique_afterstates_per_sec": uniq / t_enum},
            "enumerate_materialised": {"kernel": "k_enumerate_write", "ms": t_wr * 1e3, "sequences_per_sec": seqs / t_wr,
                                       "bytes_written_per_sequence": 41, "write_gbs": seqs * 41 / t_wr / 1e9},
            "select": {"kernel": "k_select", "ms": t_sel * 1e3, "positions_per_sec": n / t_sel,
                       "afterstates_enumerated_and_evaluated_per_sec": seqs / t_sel},
            "encode": {"kernel": "k_encode", "rows": rows, "ms": t_enc * 1e3, "rows_per_sec": rows / t_enc,
                       "roofline": {"bound": "hbm", "achieved": enc_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": enc_gbs / peaks["hbm_gbs"], "note": "824 B per row: 32 B record read + 792 B fp32[198] written"}},
            "evaluate": {"kernel": "k_evaluate", "ms": t_ev * 1e3, "rows_per_sec": n / t_ev}}


def pin_to_gpu_cpus(index):
    """Several ranks on one box: keep this rank's host threads (and its pinned buffers, first-touch) on the CPUs NVML lists as
    local to its GPU, so that the per-ply host<->device traffic of the end-to-end leg does not cross sockets.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bgx.engine import BatchEngine
    from bgx.lib import FIRST_PARITY
    from bgx.synth import START_BOARD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = pin_to_gpu_cpus(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    w = init_weights()
    eng = BatchEngine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_weights(*w)
    G = GAMES_PER_GPU
    eng.selfplay_init(G, first_id=rank * G, id_stride=world * G, seed=SEED, first_mover=FIRST_PARITY, traj_cap=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = None
    for _ in range(args.warmup):
        eng.selfplay_step(PLIES_PER_STEP, want_stats=False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, plies, seqs, scored, edges = [], 0, 0, 0, 0
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush between timed iterations (not timed)
        ev[k][0].record(stream)
        st = eng.selfplay_step(PLIES_PER_STEP)      # reads the 64-byte stats block back: one sync per step
        ev[k][1].record(stream)
        kernel_ms.append(eng.last_kernel_ms())
        plies += st["plies"]; seqs += st["sequences"]; scored += st["scored"]; edges += st["tree_edges"]
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    counts = torch.tensor([plies, seqs, scored], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) / 1e3
    plies_all, seqs_all, scored_all = (float(x) for x in counts.tolist())
    value = plies_all / total_s

    # ---- end to end: host-driven self-play through the C-ABI with HOST buffers, one call per ply and part of the
    # population: bgx_play_ply_host_async = one iteration of play_game's loop (train.py:103-121: make_move, is_game_over,
    # setTurn, roll_dice) for a batch of games.  The population is split into three parts, each on its own asynchronous lane
    # (a lane's launch occupies half the SMs, so two are resident at once): while the GPU plays some (H2D of records, next ply
    # numbers and game ids; k_select_order + k_select + k_advance; D2H of the next records and winners) the host handles
    # another (count the ply, restart finished games in place with the next game id).
    e2e_steps = max(2, min(args.steps, 8)) * PLIES_PER_STEP
    from bgx import host as bgx_host
    LANES = 3
    pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    bufs = [pin((G, 32), torch.int8), pin((G, 32), torch.int8)]
    win = pin((G,), torch.int8)
    nxt_ply = pin((G,), torch.int32)
    h_gid_pin = pin((G,), torch.int64)
    rec, h_ply, h_gid = eng.selfplay_read()          # continue the (desynchronised) games of the population from the host
    h_gid_pin[:] = h_gid
    h_gid = h_gid_pin
    parts = [(i * G // LANES, (i + 1) * G // LANES) for i in range(LANES)]
    cur = [0] * LANES
    stride = world * G
    bufs[0][:] = rec
    bgx_host.advance(bufs[0], bufs[0], SEED, h_ply, h_gid)   # dice of the current ply for every game (undo the flip below)
    bufs[0][:, 28] ^= 1
    bufs[0][:, 31] = 0

    def submit(h):
        lo, hi = parts[h]
        nxt_ply[lo:hi] = h_ply[lo:hi] + 1
        eng.play_ply_host_async(h, bufs[cur[h]][lo:hi], nxt_ply[lo:hi], h_gid[lo:hi], bufs[1 - cur[h]][lo:hi], win[lo:hi], dice_seed=SEED)

    def advance(h):
        lo, hi = parts[h]
        eng.wait(h)
        cur[h] ^= 1
        h_ply[lo:hi] += 1
        done = np.flatnonzero(win[lo:hi] >= 0)
        if done.size:                                   # restart in place with the next game id (first mover: id % 2)
            idx = done + lo
            h_gid[idx] += stride
            h_ply[idx] = 0
            fresh = np.zeros((idx.size, 32), np.int8)
            fresh[:, :24] = START_BOARD
            fresh[:, 28] = (h_gid[idx] & 1) ^ 1         # bgx_advance_host flips it and rolls ply 0
            bufs[cur[h]][idx] = bgx_host.advance(fresh, fresh, SEED, h_ply[idx], h_gid[idx])
        return hi - lo

    for h in range(LANES):
        submit(h)
    e2e_plies = 0
    for it in range(-2, e2e_steps):                 # 2 untimed warm-up ply-steps
        if it == 0:
            barrier()                               # drains nothing of ours: the lanes are non-blocking streams
            t0 = time.perf_counter()
        for h in range(LANES):
            n_adv = advance(h)
            submit(h)
            if it >= 0:
                e2e_plies += n_adv
    for h in range(LANES):
        eng.wait(h)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = e2e_plies * world / float(e2e_s.item())
    sampler.stop_flag.set()
    sampler.join(timeout=3)

    legs = None
    if rank == 0 and args.side_positions > 0:
        legs = side_legs(eng, dev, peak_json()[0], args.side_positions)

    # ---- one TD(lambda) round (BASELINE configs[3]/[4]): play to the end, replay, all-reduce, apply
    td = None
    if args.td_games < 0:                            # BASELINE.json configs[3] / configs[4]
        args.td_games = 262144 if world == 1 else (1 << 20) // world
    if args.td_games > 0:
        from bgx.lib import FIRST_ROLLOFF
        from bgx.parallel import allreduce_delta, shard
        first, n_slots, stride = shard(args.td_games * world, rank, world)
        eng.selfplay_init(n_slots, first_id=first, id_stride=stride, seed=SEED + 1, first_mover=FIRST_ROLLOFF, traj_cap=2048)
        delta = torch.zeros(25604, dtype=torch.float32, device=dev)
        allreduce_delta(delta, dist if world > 1 else None)      # untimed: NCCL sets this collective up on first use
        barrier()
        td_sampler = ClockSampler(local)                         # every rank watches its own GPU over the round
        td_sampler.start()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record(stream)
        play = eng.selfplay_round(0.0)
        e1.record(stream)
        tdst = eng.td_replay(0.1, 0.9, delta)
        e2.record(stream)
        allreduce_delta(delta, dist if world > 1 else None)
        eng.apply_delta(delta, 1.0 / (args.td_games * world))
        e3.record(stream)
        barrier()
        td_sampler.stop_flag.set()
        td_sampler.join(timeout=3)
        tcl = td_sampler.summary()
        watts = [float(r[2]) for r in td_sampler.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        lo = torch.tensor([-(tcl["sm_mhz"] or 0.0), max(watts) if watts else 0.0, float("sw_power_cap" in tcl["reasons"]),
                           float(bool(set(tcl["reasons"]) - {"sw_power_cap"}))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MAX)
        ms = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e0.elapsed_time(e3)], dtype=torch.float64, device=dev)
        cnt = torch.tensor([play["plies"], tdst["td_steps"], play["games_finished"], play["truncated"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms, cnt = ms.tolist(), cnt.tolist()
        td = {"workload": ("TD(lambda) self-play training, 262,144 concurrent games on 1 GPU (BASELINE.json configs[3])" if world == 1 and args.td_games == 262144
                           else f"TD(lambda) self-play training sharded over {world} GPU(s), {args.td_games * world:,} games in total (BASELINE.json configs[4])"
                           if world > 1 and args.td_games * world == 1 << 20 else f"TD(lambda) self-play training, {args.td_games:,} games per GPU"),
              "games": int(args.td_games * world), "plies": int(cnt[0]), "td_steps": int(cnt[1]), "games_finished": int(cnt[2]),
              "truncated": int(cnt[3]), "play_ms": ms[0], "td_replay_ms": ms[1], "allreduce_apply_ms": ms[2], "round_ms": ms[3],
              "plies_per_sec_incl_update": cnt[0] / (ms[3] * 1e-3), "td_steps_per_sec": cnt[1] / (ms[1] * 1e-3),
              "replay_ms_this_rank": e1.elapsed_time(e2),
              "clocks": {"sm_mhz_min_over_ranks": -float(lo[0]), "power_w_max_over_ranks": float(lo[1]),
                         "sw_power_cap_on_some_rank": bool(lo[2] > 0), "other_slowdown_on_some_rank": bool(lo[3] > 0),
                         "samples_rank0": tcl["samples"],
                         "note": "each rank samples its own GPU (nvidia-smi, 0.2 s) over the whole round; median SM clock per rank, min over ranks"},
              "note": "one round: every game played to its end from one snapshot (k_selfplay), exact online TD(lambda) replay "
                      "per game (k_td_replay), NCCL all-reduce of fp32[25604], apply; allreduce_apply_ms includes waiting for the "
                      "slowest rank's replay (max over ranks)"}

    if rank == 0:
        peaks, peak_src = peak_json()
        try:
            facts = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_selfplay"]
        except Exception:
            facts = {}
        traffic = facts.get("dram_bytes_per_launch")
        inst_per_ply = facts.get("warp_inst_per_ply")
        k_ms = float(np.mean(kernel_ms))
        seq_per_launch = seqs / args.steps
        scored_per_launch = scored / args.steps
        achieved = seq_per_launch * BYTES_PER_AFTERSTATE / (k_ms * 1e-3) / 1e9
        props = eng.device_props()
        fp32_peak = props["sm_count"] * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        fp32_ach = scored_per_launch * FLOP_PER_AFTERSTATE / (k_ms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(world),
                "sequences_per_sec": seqs_all / total_s, "afterstates_scored_per_sec": scored_all / total_s,
                "legal_afterstates_per_sec": seqs_all / total_s,     # BASELINE.json's second metric: every legal turn sequence's afterstate,
                                                                     # enumerated and evaluated (duplicates and twins scored once, counted all)
                "sequences_per_ply": seqs_all / plies_all, "scored_per_ply": scored_all / plies_all,
                "tree_edges_per_ply_rank0": edges / max(plies, 1), "ply_warps_per_cta": eng.kernel_config(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": G * 44, "d2h_bytes_per_step": G * 33,
                        "note": "one step = one ply for all 65,536 games through bgx_play_ply_host_async (= make_move + is_game_over + "
                                "setTurn + roll_dice of train.py:103-121 for a batch): three thirds of the population on three lanes of half "
                                "the SMs each, pinned host buffers, H2D (records, ply numbers, game ids) + k_select_order + k_select + "
                                f"k_advance + D2H (next records, winners) per ply; the host counts plies and restarts finished games; {e2e_steps} timed ply-steps"},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "kernel": "k_selfplay",
                             "kernel_ms": k_ms, "peak_source": peak_src,
                             "fused_dataflow": {"bytes_per_afterstate": 32, "achieved": seq_per_launch * 32 / (k_ms * 1e-3) / 1e9,
                                                "frac": seq_per_launch * 32 / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                "note": "SURVEY 8d's second figure: an implementation that fuses encode + evaluate but still "
                                                        "writes one packed int8 record (32 B) per enumerated afterstate; this kernel writes none"},
                             "note": "algorithmic bytes = sequences enumerated per launch x 1,696 B: what the materialised dataflow "
                                     "of the reference and of SURVEY 8d moves per enumerated afterstate (792 B of features written, "
                                     "792 B read back, 112 B state row).  The fused kernel never materialises them (`traffic` is its "
                                     "real DRAM traffic per launch), so frac > 1 means it outruns ANY implementation that stages "
                                     "the encodings through HBM; the bound that actually limits it is `roofline_issue`"},
                "roofline_fp32": {"bound": "fp32", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                                  "frac": fp32_ach / fp32_peak,
                                  "note": "afterstates actually scored x 50,944 dense-equivalent FLOP; the kernel skips zero features"},
                "wall_s_timed_region": wall}
        if inst_per_ply:
            issue_peak = props["sm_count"] * 4 * peaks.get("sm_max_mhz", 1965.0) * 1e6
            issue_ach = (plies / args.steps) * inst_per_ply / (k_ms * 1e-3)
            line["roofline_issue"] = {"bound": "warp-instruction issue", "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9,
                                      "unit": "G warp-inst/s", "frac": issue_ach / issue_peak,
                                      "note": f"{inst_per_ply:.0f} warp instructions per ply (smsp__inst_executed.sum of the committed ncu "
                                              "capture / plies of that launch) x live plies/s, against 4 schedulers x SMs x clock; "
                                              "this integer/branch kernel is bound by issue slots and dependent-issue latency, not by HBM"}
        if legs:
            line["side_kernels"] = legs
        if td:
            line["td_round"] = td
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                    "--cpu-seconds", str(args.cpu_seconds)], capture_output=True, text=True, timeout=600)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as exc:                 # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The one JSON line goes to the REAL stdout; everything else any library prints (NCCL's version
    banner, warnings) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU sample length per reference step")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--side-positions", type=int, default=1000000, help="positions of the enumeration/encode side legs (0 = skip)")
    ap.add_argument("--td-games", type=int, default=-1,
                    help="games per GPU in the TD(lambda) round leg (0 = skip; default: 262,144 on one GPU, 2^20 / N on N)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
