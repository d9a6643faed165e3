"""Structure-similarity scores of the reference (utils/sim_score.py:28-147), computed in O(N).

The reference matches brackets with a nested open x close loop (sim_score.py:51-57) and then
counts, per POSITION, whether the partner agrees (sim_score.py:104-119).  The matching it
produces is the ordinary nested one per bracket family, so a stack per family gives the same
pair map.  Values and rounding are the reference's: MCC with eps 1e-5 (:126-127), recall and
precision with +0.001 in the denominator (:137,:147), each rounded to 3 decimals.
"""
import bisect
import functools
import math

OPEN_BRACKETS = {"(": "0", "[": "1", "<": "2", "{": "3", "A": "4", "B": "5", "C": "6", "D": "7", "E": "8"}
CLOSE_BRACKETS = {")": "0", "]": "1", ">": "2", "}": "3", "a": "4", "b": "5", "c": "6", "d": "7", "e": "8"}


@functools.lru_cache(maxsize=8192)
def pairing_positions(s1):
    """dot-bracket string -> {position: partner or -1}; characters that are neither bracket nor
    '.'/'-' get no entry, as in the reference (sim_score.py:41-48).  Memoised (the target recurs in every call, MFE structures
    recur across mutants); the returned dict is shared, callers only read it."""
    pairs = {}
    stacks = {}
    closes = {}
    for i, ch in enumerate(s1):
        if ch in OPEN_BRACKETS:
            stacks.setdefault(OPEN_BRACKETS[ch], []).append(i)
        elif ch in CLOSE_BRACKETS:
            closes.setdefault(CLOSE_BRACKETS[ch], []).append(i)
        elif ch == "." or ch == "-":
            pairs[i] = -1
    # reference order: opens from last to first, each takes the first unused close to its right
    for fam, opens in stacks.items():
        cl = closes.get(fam, [])
        used = [False] * len(cl)
        for o in reversed(opens):
            k = bisect.bisect_right(cl, o)
            while k < len(cl) and used[k]:
                k += 1
            if k < len(cl):
                used[k] = True
                pairs[o] = cl[k]
                pairs[cl[k]] = o
    return dict(sorted(pairs.items()))


@functools.lru_cache(maxsize=8192)
def _confusion(ref_ss, query_ss):
    """(tp, fp, fn, tn) per POSITION as the reference counts them (sim_score.py:104-119); memoised on the two strings"""
    tp = fp = tn = fn = 0
    r, q = pairing_positions(ref_ss), pairing_positions(query_ss)
    for i in range(len(r)):
        if r[i] == q[i]:
            if r[i] != -1:
                tp += 1
            else:
                tn += 1
        elif r[i] == -1:
            fp += 1
        else:
            fn += 1
    return (tp, fp, fn, tn)


class SimScore:
    def __init__(self, ref_ss, query_ss):
        self.ref_ss = ref_ss
        self.query_ss = query_ss

    def find_basepairs(self):
        self.bp_dict_r = pairing_positions(self.ref_ss)
        self.bp_dict_q = pairing_positions(self.query_ss)

    def cofusion_matrix(self):
        self.conf_mat = _confusion(self.ref_ss, self.query_ss)

    def mcc(self):
        tp, fp, fn, tn = self.conf_mat
        if tp == 0 and fp == 0 and fn == 0 and tn != 0:
            numerator, denominator = 1, 1
        else:
            numerator = (tp * tn) - (fp * fn)
            denominator = math.sqrt((tp + fp) * (tp + fn) * (tn + fn) * (tn + fp))
        return round(numerator / (denominator + 0.00001), 3)

    def recall(self):
        tp, fp, fn, tn = self.conf_mat
        return round(tp / (tp + fn + 0.001), 3)

    def precision(self):
        tp, fp, fn, tn = self.conf_mat
        return round(tp / (tp + fp + 0.001), 3)

    def fscore(self):
        makhraj = self.precision() + self.recall()
        if makhraj < 0.001:
            makhraj = 0.001
        return round(2 * (self.precision() * self.recall() / makhraj), 4)
