"""Host-side mirror of the pieces of the reference's utils/sequence_utils.py that sit on either side of the scoring
path: sequence restraints -> allowed letters (get_nt_list, :454-525), start sequences (initial_sequence_generator,
:667-763), temperature shelves (get_rep_temps, :811-859) and the move generator (get_mutation_position :926-980,
expand_cases :983-1005, mutate_sequence :1008-1136).

Same names, arguments and random-number consumption as the reference, so that the lock-step replica loop
(utils/replica_exchange_monte_carlo.py here) driven by this module reproduces the reference's trajectories draw for
draw, and so that the device move generator (csrc/bf_design.cu) can be tested against it in distribution.
Not covered (SURVEY.md section 8 marks them out of the hot path): alternative-structure "snake" graphs (:119-431)."""
import random

import numpy as np

from . import energy_scores as es

IUPAC = {"N": "ACGU", "W": "AU", "S": "CG", "M": "AC", "K": "GU", "R": "AG", "Y": "CU", "B": "CGU", "D": "AGU", "H": "ACU",
         "V": "ACG", "A": "A", "C": "C", "G": "G", "U": "U", "&": "&"}
PAIRS_WITH = {"A": "U", "U": "GA", "G": "UC", "C": "G"}        # can_pair (:607-622), in the reference's list order
WATSON_CRICK = {"A": "U", "U": "A", "G": "C", "C": "G"}
BRACKETS = ("()", "[]", "<>", "{}", "Aa", "Bb", "Cc", "Dd", "Ee")
LETTER_BIT = {"A": 1, "C": 2, "G": 4, "U": 8}


def nt_dictionary(nt):
    return list(IUPAC[nt])


def can_pair(nt):
    return list(PAIRS_WITH[nt])


def wc_pair(nt):
    return WATSON_CRICK[nt]


def check_dot_bracket(ss):
    """[[open, close], ...] of every bracket family, family by family (:74-116); ValueError instead of sys.exit()."""
    if any(ch not in "()[]<>{}AaBbCcDdEe.&" for ch in ss):
        raise ValueError("Not allowed characters in structures. Check input file.")
    pairs = []
    for opn, cls in BRACKETS:
        stack = []
        for pos, ch in enumerate(ss):
            if ch == opn:
                stack.append(pos)
            elif ch == cls:
                if not stack:
                    raise ValueError("There is no opening bracket for nt position %d-%s" % (pos + 1, ch))
                pairs.append([stack.pop(), pos])
        if stack:
            raise ValueError("There is no closing bracket for nt position %d-%s" % (stack[-1], ss[stack[-1]]))
    return pairs


class Nucleotide:
    """what get_nt_list keeps per position (:1277-1343)"""

    def __init__(self, number):
        self.number = number
        self.letters = None
        self.pairs_with = None
        self.pair_letters = None
        self.letters_allowed = None
        self.snake = False


def get_nt_list(input_file):
    """Restraints -> per-position allowed letters (:454-525).  A paired position keeps the letters of its own restraint
    that can pair with at least one letter its partner's restraint allows."""
    if getattr(input_file, "graphs", None) is not None or getattr(input_file, "excluded_alt_pairs", None) is not None:
        raise NotImplementedError("alternative-structure graphs are outside the accelerated path")
    nts = []
    for i, ch in enumerate(input_file.seq_restr):
        nt = Nucleotide(i)
        nt.letters = nt_dictionary(ch)
        nts.append(nt)
    for a, b in input_file.pairs:
        for x, y in ((a, b), (b, a)):
            nts[x].pairs_with = y
        for x in (a, b):
            seen = []
            for letter in nts[x].letters:
                for p in can_pair(letter):
                    if p not in seen:
                        seen.append(p)
            nts[x].pair_letters = seen
        nts[a].letters_allowed = [l for l in nts[a].letters if l in nts[b].pair_letters]
        nts[b].letters_allowed = [l for l in nts[b].letters if l in nts[a].pair_letters]
        if not nts[a].letters_allowed or not nts[b].letters_allowed:
            raise ValueError("Wrong restraints in the input file. Nucleotide %d %s, cannot pair with nucleotide %d %s"
                             % (a + 1, nts[a].letters, b + 1, nts[b].letters))
    for nt in nts:
        if nt.letters_allowed is None:
            nt.letters_allowed = nt.letters
    return nts


def allowed_masks(nt_list):
    """letters_allowed as bit masks A=1 C=2 G=4 U=8 (the `allowed` array of bf_design_t)"""
    return np.array([sum(LETTER_BIT[l] for l in nt.letters_allowed) for nt in nt_list], np.uint8)


def allowed_choice(allowed, percs):
    return [percs[nt] for nt in allowed]


def get_rep_temps(sim_options):
    """evenly spaced shelves from T_min to T_max, rounded to 3 decimals; one replica sits at T_max (:811-859)"""
    R = sim_options.replicas
    if R == 1:
        return [sim_options.T_max]
    delta = (sim_options.T_max - sim_options.T_min) / (R - 1)
    temps, t = [], sim_options.T_min
    for _ in range(R):
        temps.append(round(t, 3))
        t += delta
    return temps


def initial_sequence_generator(nt_list, input_file, sim_options):
    """Start sequence from the target (:667-763): A in loops, G (else U) at the first unpaired position after a helix unless
    it is a one-nucleotide bulge, GC pairs where allowed (else AU), random among the allowed letters elsewhere."""
    seq = list(input_file.seq_restr)
    n = len(nt_list)
    for i, nt in enumerate(nt_list):
        if nt.pairs_with is None and "A" in nt.letters:
            seq[i] = "A"
    for i in range(1, n - 1):
        if nt_list[i].pairs_with is None and nt_list[i - 1].pairs_with is not None and nt_list[i + 1].pairs_with is None:
            if "G" in nt_list[i].letters:
                seq[i] = "G"
            elif "U" in nt_list[i].letters:
                seq[i] = "U"
    for a, b in sorted(input_file.pairs):
        la, lb = nt_list[a].letters_allowed, nt_list[b].letters_allowed
        if sim_options.acgu_percentages == "on":
            la.sort()
            lb.sort()
            seq[a] = random.choices(la, weights=allowed_choice(la, sim_options.nt_percentages))[0]
            seq[b] = wc_pair(seq[a])
            continue
        for strong, weak in (("C", "G"), ("A", "U")):
            if all(x in la for x in (strong, weak)) and all(x in lb for x in (strong, weak)):
                seq[a] = random.choice([strong, weak])
                seq[b] = wc_pair(seq[a])
                break
            if strong in la and weak in lb:
                seq[a], seq[b] = strong, weak
                break
            if weak in la and strong in lb:
                seq[a], seq[b] = weak, strong
                break
    for i, ch in enumerate(seq):
        if ch not in "ACGU":
            seq[i] = random.choice(nt_dictionary(ch))
    return "".join(seq)


def generate_initial_list(nt_list, input_file, sim_options):
    """R scored start records, replica r on shelf r (:862-888)"""
    sequence = initial_sequence_generator(nt_list, input_file, sim_options)
    seqs = []
    for _ in range(sim_options.replicas):
        if sim_options.diff_start_replicas == "different":
            sequence = initial_sequence_generator(nt_list, input_file, sim_options)
        seqs.append(sequence)
    out = es.score_sequences(seqs, input_file, sim_options)
    for i, obj in enumerate(out):
        obj.get_replica_num(i + 1)
        obj.get_temp_shelf(sim_options.rep_temps_shelfs[i])
        obj.get_sim_step(0)
    return out


def round_floats(obj):
    """floats and dict values to 3 decimals (:1254-1274); like the reference, a LIST passes through unchanged"""
    if isinstance(obj, float):
        return round(obj, 3)
    if isinstance(obj, dict):
        return {k: round_floats(v) for k, v in obj.items()}
    return obj


def expand_cases(cases, max_value, range_expansion=3):
    out = set()
    for c in cases:
        out.update(v for v in range(c - range_expansion, c + range_expansion + 1) if 0 < v <= max_value)
    return sorted(out)


def targeted_move_probabilities(sim_options):
    """probability of a targeted move per temperature shelf: tm_max on the coldest to tm_min on the hottest (:963)"""
    return [round(float(x), 2) for x in np.linspace(sim_options.tm_max, sim_options.tm_min, num=len(sim_options.rep_temps_shelfs))]


def get_mutation_position(seq_obj, available_positions, sim_options, input_file):
    """Where to mutate (:926-980): with the shelf's probability, near a position whose pairing differs between the target
    and the current MFE structure; otherwise anywhere mutable."""
    if sim_options.point_mutations == "off":
        return random.choice(available_positions)
    query = {tuple(p) for p in check_dot_bracket(seq_obj.mfe_ss)}
    target = input_file.target_pairs_tupl
    avail = set(available_positions)
    wrong = [x for pair in target - query for x in pair if x in avail] + [x for pair in query - target for x in pair if x in avail]
    wrong = list(set(wrong))
    p = targeted_move_probabilities(sim_options)[sim_options.rep_temps_shelfs.index(seq_obj.temp_shelf)]
    if not wrong:
        pool = available_positions
    else:
        near = expand_cases(wrong, len(seq_obj.sequence) - 1)
        pool = random.choices([near, available_positions], weights=[p, 1 - p])[0]
    return random.choice(pool)


def _point_move(sequence_obj, nt_list, sim_options, input_file):
    """one draw of the move (:1023-1083): (mutant string, mutated position, partner position or None)"""
    seq = list(sequence_obj.sequence)
    mutable = [i for i in range(len(seq)) if len(nt_list[i].letters_allowed) != 1]
    pos = get_mutation_position(sequence_obj, mutable, sim_options, input_file)
    nt = nt_list[pos]
    if nt.snake:
        raise NotImplementedError("alternative-structure graphs are outside the accelerated path")
    if nt.pairs_with is None:
        if len(nt.letters_allowed) != 1:
            options = [l for l in nt.letters_allowed if l != seq[pos]] if seq[pos] in nt.letters_allowed else list(nt.letters_allowed)
            seq[pos] = random.choice(options)
        return "".join(seq), pos, None
    options = list(nt.letters_allowed)
    if seq[pos] in options and len(options) != 1:
        options.remove(seq[pos])
    options.sort()
    partner = nt_list[nt.pairs_with]
    weighted = sim_options.acgu_percentages == "on"
    first = random.choices(options, weights=allowed_choice(options, sim_options.nt_percentages))[0] if weighted else random.choice(options)
    second_options = list(set(partner.letters_allowed).intersection(can_pair(first)))
    second = (random.choices(second_options, weights=allowed_choice(second_options, sim_options.nt_percentages))[0] if weighted
              else random.choice(second_options))
    seq[pos], seq[partner.number] = first, second
    return "".join(seq), pos, partner.number


def propose_mutation(sequence_obj, nt_list, sim_options, input_file):
    """The move of mutate_sequence (:1008-1128) up to, not including, the scoring call: returns the mutant string.
    Unpaired position: another allowed letter.  Paired position: another allowed letter, then a partner letter that
    pairs with it (G-U wobbles included).  Homodimer designs keep the two strands identical: with identical target halves
    the strand that changed is copied over the other; with different halves the two changed letters are mirrored onto the
    other strand (:1102-1128)."""
    mutant, pos, partner = _point_move(sequence_obj, nt_list, sim_options, input_file)
    if sim_options.oligo_state != "homodimer":
        return mutant
    half_a, half_b = input_file.sec_struct.split("&")[:2]
    old_a, old_b = sequence_obj.sequence.split("&")[:2]
    new_a, new_b = mutant.split("&")[:2]
    if half_a != half_b:
        if partner is not None:
            lo, hi = sorted((pos, partner))
            A = len(old_a)
            if lo < A < hi:
                in_a, in_b = lo, hi - A - 1        # the reference reads the pair as (strand A, strand B)
                new_a, new_b = (new_a[:in_b] + new_b[in_b] + new_a[in_b + 1:], new_b[:in_a] + new_a[in_a] + new_b[in_a + 1:])
            elif hi < A:
                # Both letters inside strand A.  The reference's index arithmetic goes negative here (Python wrap-around, strings
                # of the wrong length for a pair ending at A-1); the deliberate choice, on host and device alike: the two new
                # letters are mirrored onto the same positions of strand B, which is what keeps the strands identical.
                new_b = "".join(new_a[k] if k in (lo, hi) else ch for k, ch in enumerate(new_b))
            else:
                new_a = "".join(new_b[k] if k in (lo - A - 1, hi - A - 1) else ch for k, ch in enumerate(new_a))
        return new_a + "&" + new_b
    if new_a != old_a:
        return new_a + "&" + new_a
    if new_b != old_b:
        return new_b + "&" + new_b
    return mutant


def mutate_sequence(sequence_obj, nt_list, sim_options, input_file):
    """mutate_sequence (:1008-1136): move, score, stamp replica number and shelf."""
    out = es.score_sequence(propose_mutation(sequence_obj, nt_list, sim_options, input_file), input_file, sim_options)
    out.get_replica_num(sequence_obj.replica_num)
    out.get_temp_shelf(sequence_obj.temp_shelf)
    return out
