"""Host-side mirror of the reference's scoring modules (same names, arguments and return values):

    utils/energy_scores.py            -> desirna_b200.utils.energy_scores
    utils/dimer_multichain_energy.py  -> desirna_b200.utils.dimer_multichain_energy
    utils/sim_score.py                -> desirna_b200.utils.sim_score
    utils/replica_exchange_monte_carlo.py (fan-out part) -> desirna_b200.utils.replica_exchange_monte_carlo

Everything numeric is computed by the B200 engine (desirna_b200.engine, C-ABI include/b200fold.h) through the
ViennaRNA-shaped shim desirna_b200.RNA; nothing here imports ViennaRNA.
"""
