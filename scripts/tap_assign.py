"""Design aid for the flat tap tables of the interior-loop inner product (bf_fill3.cu): assigns the (loop size s, 5' unpaired u1)
taps of the three decomposable interior-loop families to (slot, lane) so that
  * the lanes of one slot read distinct shared-memory banks (bank = (u1 - s*c) mod M up to a per-cell constant, c = row stride
    mod M; M = 32 for 4-byte MFE entries, 16 per half-warp for 8-byte PF entries), and
  * slots are sorted by loop size, so that the short diagonals (d - 6 < 30) run only the slots they need.
The same greedy runs in bf_fill3.cu on the host; this script prints its quality for every stride residue."""
import sys


def taps(kind):
    if kind == 0:
        return [(s, u1) for s in range(6, 31) for u1 in range(2, s - 1)]
    if kind == 1:
        return [(s, u1) for s in range(4, 31) for u1 in (1, s - 1)]
    return [(s, u1) for s in range(2, 31) for u1 in (0, s)]


def assign(T, c, M, lanes=32, max_conf=1):
    """greedy, s-major: each tap goes to the first open slot where its bank class is still free (in its half for M = 16)"""
    slots = []   # each: list of (s,u1,lane)
    used = []    # per slot: set of (half, bank)
    for (s, u1) in T:
        b = (u1 - s * c) % M
        placed = False
        for k in range(len(slots)):
            if len(slots[k]) >= lanes:
                continue
            halves = (0, 1) if M == 16 else (0,)
            for h in halves:
                cap = lanes // len(halves)
                nh = sum(1 for x in slots[k] if x[2] == h)
                if nh >= cap or (h, b) in used[k]:
                    continue
                slots[k].append((s, u1, h))
                used[k].add((h, b))
                placed = True
                break
            if placed:
                break
        if not placed:
            slots.append([(s, u1, 0)])
            used.append({(0, b)})
    return slots


def cost(slots_by_kind, L):
    """executed slot-loads summed over the pairable-cell-weighted diagonals of a length-L sequence"""
    tot = ideal = 0
    for d in range(4, L):
        smax = min(30, d - 6)
        cells = L - d
        for slots in slots_by_kind:
            n = sum(1 for sl in slots if min(x[0] for x in sl) <= smax)
            tot += cells * n
            ideal += cells * sum(1 for sl in slots for x in sl if x[0] <= smax) / 32.0
    return tot, ideal


if __name__ == "__main__":
    for M in (32, 16):
        print("M", M)
        for c in range(M):
            sk = [assign(taps(k), c, M) for k in range(3)]
            ns = [len(x) for x in sk]
            r = [cost(sk, L) for L in (50, 100, 400)]
            print(c, ns, " ".join("%.2f" % (t / i) for t, i in r))
