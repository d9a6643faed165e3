"""Replica-exchange Monte Carlo fan-out of the reference (utils/replica_exchange_monte_carlo.py:26-271),
re-plumbed for a batched GPU scorer.

The reference forks a fresh `mp.Pool(R)` every global step (:248); each worker reseeds the global `random`
module with its replica index (:227-228, :250) and runs `RE_attempt` sequential Metropolis sub-steps
(:193-208), each ending in one `score_sequence` call.  Here all replicas advance in LOCK STEP: sub-step k
of every replica is proposed first (each replica on its own copy of the `random` state, so the draws are the
ones its worker process would have made), the R mutants are scored by ONE engine call
(energy_scores.score_sequences), then each replica takes its Metropolis decision on its own stream.

Multi-GPU (one process per GPU, torch.distributed): replica r lives on rank r mod world.  A rank only
advances its own replicas; after the sub-steps the ranks all-gather the packed per-replica records (a few
hundred bytes each: scores as float64, sequence and structure as bytes) -- NCCL over NVLink on the GPU
box, gloo in the CPU tests -- and every rank applies the same neighbour swaps with the shared parent RNG
stream, exactly as the parent process does in the reference (:113-173).  Nothing else crosses ranks.

Same names and arguments as the reference: mc_delta, metropolis_score, replica_exchange_attempt,
replica_exchange, single_replica_design, mutate_sequence_re (two optional keyword arguments added).
"""
import math
import random

import numpy as np

from . import energy_scores as es


def metropolis_score(temp, dE, sim_options):
    return math.exp((-sim_options.L / temp) * dE)


def mc_delta(deltaF_o, deltaF_m, T_replica, sim_options, rng=None):
    """accept if not worse, else with probability exp(-L/T * dS); returns (accept, accepted_because_better).
    rng: the generator the uniform number is drawn from (default: the module-level stream, as in the reference)"""
    if deltaF_m <= deltaF_o:
        return True, True
    p = metropolis_score(T_replica, deltaF_m - deltaF_o, sim_options)
    return p > (rng.random() if rng is not None else random.random()), False


def replica_exchange_attempt(T0, T1, dE0, dE1, sim_options):
    if dE1 <= dE0:
        return True, True
    rand_num = random.random()
    p = math.exp(sim_options.L * (1 / T0 - 1 / T1) * (dE0 - dE1))
    return p > rand_num, False


def replica_exchange(seq_score_list, stats_obj, sim_options):
    """Neighbour swaps in temperature order: pairs (1,2),(3,4).. on even global steps, (0,1),(2,3).. on odd ones;
    an accepted swap exchanges the two temperature shelves (reference :113-173).  Uses the caller's RNG stream."""
    n = len(seq_score_list)
    if stats_obj.global_step % 2 == 0:
        pairs = [(i + 1, i + 2) for i in range(0, n - 2, 2)]
    else:
        pairs = [(i, i + 1) for i in range(0, n - 1, 2)]
    ordered = sorted(seq_score_list, key=lambda obj: obj.temp_shelf)
    for a, b in pairs:
        T_i, T_j = ordered[a].temp_shelf, ordered[b].temp_shelf
        accept, better = replica_exchange_attempt(T_i, T_j, ordered[a].scoring_function, ordered[b].scoring_function, sim_options)
        if accept:
            ordered[a].get_temp_shelf(T_j)
            ordered[b].get_temp_shelf(T_i)
            stats_obj.update_acc_re_step()
            if better:
                stats_obj.update_acc_re_better_e()
        else:
            stats_obj.update_rej_re_step()
    return sorted(ordered, key=lambda obj: obj.replica_num), stats_obj


def single_replica_design(sequence_o, nt_list, worker_stats, sim_options, input_file, mutate=None):
    """The reference's sequential inner loop (:176-210), kept for one-replica use and as the yardstick of the
    lock-step loop in the tests.  `mutate(seq_obj, nt_list, sim_options, input_file) -> ScoreSeq`."""
    mutate = mutate or _default_mutate()
    worker_stats.reset_mc_stats()
    for _ in range(sim_options.RE_attempt):
        sequence_m = mutate(sequence_o, nt_list, sim_options, input_file)
        accept, better = mc_delta(sequence_o.scoring_function, sequence_m.scoring_function, sequence_o.temp_shelf, sim_options)
        if accept:
            sequence_o = sequence_m
            worker_stats.update_acc_mc_step()
            if better:
                worker_stats.update_acc_mc_better_e()
        else:
            worker_stats.update_rej_mc_step()
    return sequence_o, worker_stats


def _default_mutate():
    try:
        from utils import sequence_utils as seq_utils  # DesiRNA's own move generator, when running inside DesiRNA
    except Exception:
        from . import sequence_utils as seq_utils       # its mirror here (same draws, utils/sequence_utils.py:926-1136)
    return seq_utils.mutate_sequence


class _Pending:
    """placeholder returned by the move generator while scoring is deferred"""

    def __init__(self, seq):
        self.sequence = seq

    def get_replica_num(self, rep_num):   # mutate_sequence stamps its result (sequence_utils.py:1133-1134)
        pass

    def get_temp_shelf(self, temp):
        pass


def _propose(mutate, seq_obj, nt_list, sim_options, input_file):
    """Run the move generator with scoring deferred; returns the mutant sequence string."""
    with es_deferred() as _:
        res = mutate(seq_obj, nt_list, sim_options, input_file)
    return res if isinstance(res, str) else res.sequence


class es_deferred:
    """Context manager: while active, energy_scores.score_sequence records the sequence instead of folding it
    (DesiRNA's mutate_sequence ends in `return es.score_sequence(...)`, sequence_utils.py:1132)."""

    def __enter__(self):
        self._orig = es.score_sequence
        es.score_sequence = lambda seq, input_file, sim_options: _Pending(seq)
        return self

    def __exit__(self, *exc):
        es.score_sequence = self._orig
        return False


# ------------------------------------------------------------------------------------------ record packing
_NUM_FIELDS = ["scoring_function", "replica_num", "temp_shelf", "sim_step", "edesired_minus_Epf", "Epf", "edesired", "mcc", "mcc_alt",
               "subopt_e", "esubopt_minus_Epf", "sln_Epf", "MFE", "edesired_minus_MFE", "recall", "precision", "edesired2",
               "edesired2_minus_Epf", "ensemble_defect", "oligo_fraction", "oligomer_bonus", "monomer_bonus"]


def pack_records(objs, stride):
    """list of ScoreSeq -> (float64 [n, F], uint8 [n, 2, stride]); NaN marks an absent optional field"""
    num = np.full((len(objs), len(_NUM_FIELDS)), np.nan)
    txt = np.zeros((len(objs), 2, stride), np.uint8)
    for k, o in enumerate(objs):
        for f, name in enumerate(_NUM_FIELDS):
            v = getattr(o, name, None)
            if v is not None:
                num[k, f] = v
        for t, s in enumerate((o.sequence, o.mfe_ss or "")):
            raw = s.encode("ascii")
            txt[k, t, :len(raw)] = np.frombuffer(raw, np.uint8)
    return num, txt


def unpack_records(num, txt):
    out = []
    for k in range(num.shape[0]):
        seq = bytes(txt[k, 0]).rstrip(b"\0").decode("ascii")
        o = es.ScoreSeq(sequence=seq)
        ss = bytes(txt[k, 1]).rstrip(b"\0").decode("ascii")
        o.mfe_ss = ss or None
        for f, name in enumerate(_NUM_FIELDS):
            v = num[k, f]
            if not np.isnan(v):
                setattr(o, name, int(v) if name in ("replica_num", "sim_step") else float(v))
            elif name in ("replica_num", "temp_shelf"):
                setattr(o, name, None)
        out.append(o)
    return out


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def _all_gather_records(local_objs, counts, stride, device):
    """all-gather the packed records of every rank (padded to the largest shard); returns the objects of all ranks,
    rank-major"""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    cap = max(counts)
    num, txt = pack_records(local_objs, stride)
    nbuf = torch.full((cap, num.shape[1]), float("nan"), dtype=torch.float64)
    tbuf = torch.zeros((cap, 2, stride), dtype=torch.uint8)
    nbuf[:len(local_objs)] = torch.from_numpy(num)
    tbuf[:len(local_objs)] = torch.from_numpy(txt)
    nbuf, tbuf = nbuf.to(device), tbuf.to(device)
    nall = [torch.empty_like(nbuf) for _ in range(world)]
    tall = [torch.empty_like(tbuf) for _ in range(world)]
    dist.all_gather(nall, nbuf)
    dist.all_gather(tall, tbuf)
    out = []
    for r in range(world):
        out.append(unpack_records(nall[r][:counts[r]].cpu().numpy(), tall[r][:counts[r]].cpu().numpy()))
    return out


# ------------------------------------------------------------------------------------------ the fan-out
class _ReplicaStreams:
    """One random.Random per replica, seeded with the replica index like the reference's workers (:227-228).  use(r) binds the
    module-level functions of `random` (bound methods of its hidden instance) to replica r's generator, restore() puts the
    parent's back; code that calls random.choice / random.random / ... -- DesiRNA's own move generator included -- then draws from
    that replica's stream."""
    _NAMES = [n for n in random.__all__ if callable(getattr(random._inst, n, None))]

    def __init__(self, replicas):
        self._saved = [(n, getattr(random, n)) for n in self._NAMES]
        self._bound = {}
        self.gen = {}
        for r in replicas:
            g = random.Random()
            g.seed(r)
            self.gen[r] = g
            self._bound[r] = [(n, getattr(g, n)) for n in self._NAMES]

    def use(self, r):
        for n, f in self._bound[r]:
            setattr(random, n, f)

    def restore(self):
        for n, f in self._saved:
            setattr(random, n, f)


def mutate_sequence_re(lst_seq_obj, nt_list, stats_obj, sim_options, input_file, mutate=None, device=None):
    """One global step of the reference's fan-out (:233-271): every replica makes `sim_options.RE_attempt`
    Metropolis sub-steps.  Returns (new list of ScoreSeq in replica order, stats_obj) on every rank.

    mutate: DesiRNA's `sequence_utils.mutate_sequence` (default when importable) or any callable with its signature
            returning either a scored object or the mutant sequence string; its scoring call is deferred and batched.
    device: torch device of the all-gather buffers in distributed runs (default: cuda:LOCAL_RANK with NCCL, cpu with gloo)."""
    mutate = mutate or _default_mutate()
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    R = len(lst_seq_obj)
    mine = [r for r in range(R) if r % world == rank]
    cur = {r: lst_seq_obj[r] for r in mine}
    # every worker of the reference starts each global step from random.seed(replica index)  (:227-228, :250): one generator per
    # replica, switched in under the module-level names the move generator and mc_delta call (random.choice, random.random, ...) --
    # the same draws as random.seed(r) followed by getstate / setstate around every use, without copying the 625-word state
    streams = _ReplicaStreams(mine)
    acc = better = rej = 0
    try:
        for _ in range(sim_options.RE_attempt):
            mutants = []
            for r in mine:
                streams.use(r)
                mutants.append(_propose(mutate, cur[r], nt_list, sim_options, input_file))
            streams.restore()
            scored = es.score_sequences(mutants, input_file, sim_options)
            for r, new in zip(mine, scored):
                old = cur[r]
                # what mutate_sequence copies from the parent record (sequence_utils.py:1133-1134)
                new.get_replica_num(old.replica_num)
                new.get_temp_shelf(old.temp_shelf)
                ok, was_better = mc_delta(old.scoring_function, new.scoring_function, old.temp_shelf, sim_options, rng=streams.gen[r])
                if ok:
                    cur[r] = new
                    acc += 1
                    better += int(was_better)
                else:
                    rej += 1
    finally:
        streams.restore()   # the parent stream, untouched by the replicas, drives replica_exchange only

    if dist:
        import torch
        counts = [len(range(k, R, world)) for k in range(world)]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        stride = max(8, max(len(o.sequence) for o in lst_seq_obj) + 8)
        shards = _all_gather_records([cur[r] for r in mine], counts, stride, device)
        result = [None] * R
        for k in range(world):
            for idx, r in enumerate(range(k, R, world)):
                result[r] = shards[k][idx]
        tot = torch.tensor([acc, better, rej], dtype=torch.int64, device=device)
        dist.all_reduce(tot)
        acc, better, rej = (int(x) for x in tot.cpu())
    else:
        result = [cur[r] for r in range(R)]
    stats_obj.update_step(sim_options.RE_attempt)
    stats_obj.acc_mc_step += acc
    stats_obj.acc_mc_better_e += better
    stats_obj.rej_mc_step += rej
    return result, stats_obj
