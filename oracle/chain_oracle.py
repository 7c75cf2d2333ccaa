"""Mapper chain flags -- ORACLE (test infrastructure only) for what `damapper` encodes in the LAS flags
(START / NEXT / BEST; decoded by DENTIST at dazzler.d:1738-1755, packed into chains at :708-743).
Parity unpinned (DAMAPPER @ b2c9d7fd absent).  Spec: see csrc/pile.cu "mapper chains"."""
import numpy as np

COMP, START, NEXT, BEST = 0x1, 0x4, 0x8, 0x10


def _continues(x, y, max_indel, max_gap):
    if x["aread"] != y["aread"] or x["bread"] != y["bread"] or ((int(x["flags"]) ^ int(y["flags"])) & COMP):
        return False
    if not (y["abpos"] > x["abpos"] and y["aepos"] > x["aepos"] and y["bbpos"] > x["bbpos"] and y["bepos"] > x["bepos"]):
        return False
    ga = int(y["abpos"]) - int(x["aepos"]); gb = int(y["bbpos"]) - int(x["bepos"])
    return abs(ga - gb) <= max_indel and max(abs(ga), abs(gb)) <= max_gap


def mapper_chain_flags(rec, max_indel=1000, max_gap=10000):
    """rec in LAsort order; returns the new flag words."""
    n = len(rec)
    cont = [False] * n
    for i in range(1, n):
        cont[i] = _continues(rec[i - 1], rec[i], max_indel, max_gap)
    best = {}
    i = 0
    while i < n:
        j = i
        score = 0
        while True:
            score += int(rec[j]["aepos"]) - int(rec[j]["abpos"])
            if j + 1 < n and cont[j + 1]:
                j += 1
            else:
                break
        b = int(rec[i]["bread"])
        if b not in best or score > best[b][0]:
            best[b] = (score, i)
        i = j + 1
    out = np.zeros(n, np.uint32)
    for i in range(n):
        f = int(rec[i]["flags"]) & ~(START | NEXT | BEST)
        if cont[i]:
            f |= NEXT
        else:
            f |= START
            if best[int(rec[i]["bread"])][1] == i:
                f |= BEST
        out[i] = f
    return out


def keep_best_chains(rec, flags, n_frac=0.0):
    """damapper's reporting rule (dazzler.d:5920-5923): indices of the records of every read's BEST chain and of the
    chains whose score (sum of A spans) reaches n_frac of the best; n_frac <= 0: best chains only."""
    n = len(rec)
    chains, i = [], 0
    while i < n:
        j = i + 1
        while j < n and (int(flags[j]) & NEXT):
            j += 1
        chains.append((i, j, sum(int(rec[x]["aepos"]) - int(rec[x]["abpos"]) for x in range(i, j))))
        i = j
    best = {int(rec[i]["bread"]): sc for i, j, sc in chains if int(flags[i]) & BEST}
    keep = []
    for i, j, sc in chains:
        if (int(flags[i]) & BEST) or (n_frac > 0 and sc >= n_frac * best[int(rec[i]["bread"])]):
            keep += list(range(i, j))
    return np.array(keep, np.int64)
