"""Host-side mirror of the reference's scoring modules (same names, arguments and return values):

    utils/energy_scores.py            -> desirna_b200.utils.energy_scores
    utils/dimer_multichain_energy.py  -> desirna_b200.utils.dimer_multichain_energy
    utils/sim_score.py                -> desirna_b200.utils.sim_score
    utils/replica_exchange_monte_carlo.py (fan-out part) -> desirna_b200.utils.replica_exchange_monte_carlo
    utils/sequence_utils.py (restraints, start sequences, move generator) -> desirna_b200.utils.sequence_utils
    utils/stats_inputs_outputs.py (input reader, -sf parser, Stats)   -> desirna_b200.utils.stats_inputs_outputs

The device-resident form of the whole loop (many targets x replicas in one batched loop) is desirna_b200.design.

Everything numeric is computed by the B200 engine (desirna_b200.engine, C-ABI include/b200fold.h) through the
ViennaRNA-shaped shim desirna_b200.RNA; nothing here imports ViennaRNA.
"""
