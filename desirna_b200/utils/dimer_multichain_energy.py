"""Two-strand thermodynamics of the reference (utils/dimer_multichain_energy.py:30-118) on the B200 engine.
Same constants, same signatures; `fc` is a desirna_b200.RNA.fold_compound."""
import numpy as np

from .. import RNA

KB = 0.001987204259
RHO = 55.14        # H2O concentration in mol/L
TEMP = 273.15 + 37
CONC = 1e-3        # RNA concentration in mol/L


def oligo_fraction(seq_dimer, fc):
    """dimer_multichain_energy.py:36-50: equilibrium dimer fraction from FcAB - FA - FB at 1 mM"""
    dimer_ss, pfa, pfb, pfab, dimer_pf = fc.pf_dimer()
    dF = pfab - pfa - pfb
    rhs = CONC / RHO * np.exp(-dF / (KB * TEMP))
    return 1 - (np.sqrt(1 + 4 * rhs) - 1) / (2 * rhs)


def kTlog_oligo_fraction(oligo_frac):
    return -KB * TEMP * np.log(oligo_frac)


def kTlog_monomer_fraction(oligo_frac):
    return -KB * TEMP * np.log(1 - oligo_frac)


def energy_of_oligomer(seq):
    fc = RNA.fold_compound(seq + "&" + seq)
    oligo_structure, mfe_oligo = fc.mfe_dimer()
    return mfe_oligo


def mfe_e_dimer(seq_dimer):
    fc = RNA.fold_compound(seq_dimer)
    mfe_ss_dimer_joint = fc.mfe_dimer()[0]
    seqa, seqb = seq_dimer.split("&")[0], seq_dimer.split("&")[1]
    cofolded = RNA.co_pf_fold(seqa + "&" + seqb)
    e_dimer = cofolded[-1]
    mfe_e_a = RNA.fold(seqa)[1]
    mfe_e_b = RNA.fold(seqb)[1]
    mfe_ss_dimer = mfe_ss_dimer_joint[:len(seqa)] + "&" + mfe_ss_dimer_joint[len(seqa):]
    return e_dimer, mfe_e_a, mfe_e_b, mfe_ss_dimer
