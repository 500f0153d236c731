"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the forced aligner's integer fix-ups:
TimestampCorrection (/root/reference/Sources/Qwen3ASR/TimestampCorrection.swift:15-145) and
Qwen3ForcedAligner.findTrailingPlateauStart (/root/reference/Sources/Qwen3ASR/ForcedAligner.swift:191-216).
Pinned by the reference's own unit tests (Tests/Qwen3ASRTests/ForcedAlignerTests.swift:213-259, 441-492)."""
import numpy as np


def lis_positions(arr):                                   # :101-144
    arr = [int(v) for v in arr]
    if not arr:
        return []
    tails, tail_idx, parent = [], [], [-1] * len(arr)
    for i, v in enumerate(arr):
        lo, hi = 0, len(tails)
        while lo < hi:
            mid = (lo + hi) // 2
            if tails[mid] < v:
                lo = mid + 1
            else:
                hi = mid
        if lo == len(tails):
            tails.append(v)
            tail_idx.append(i)
        else:
            tails[lo] = v
            tail_idx[lo] = i
        parent[i] = tail_idx[lo - 1] if lo > 0 else -1
    out, idx = [], tail_idx[-1]
    while idx != -1:
        out.append(idx)
        idx = parent[idx]
    return out[::-1]


def enforce_monotonicity(raw):                            # :15-98
    raw = [int(v) for v in raw]
    if len(raw) <= 1:
        return list(raw)
    pos = lis_positions(raw)
    anchors = [(p, raw[p]) for p in pos]
    if len(anchors) == len(raw):
        return list(raw)
    in_lis = set(pos)
    out = list(raw)
    ai = 0
    for i in range(len(out)):
        if i in in_lis:
            ai = next((k for k, a in enumerate(anchors) if a[0] == i), ai)
            continue
        if ai < len(anchors) and anchors[ai][0] < i:
            prev = anchors[ai]
        elif ai > 0:
            prev = anchors[ai - 1]
        else:
            prev = None
        ni = ai
        while ni < len(anchors) and anchors[ni][0] <= i:
            ni += 1
        nxt = anchors[ni] if ni < len(anchors) else None
        if prev is not None and nxt is not None:
            if nxt[0] - prev[0] <= 3:
                out[i] = prev[1] if (i - prev[0]) <= (nxt[0] - i) else nxt[1]
            else:
                t = np.float32(i - prev[0]) / np.float32(nxt[0] - prev[0])
                out[i] = prev[1] + int(np.float32(t * np.float32(nxt[1] - prev[1])))   # Int(Float): toward zero
        elif prev is not None:
            out[i] = prev[1]
        elif nxt is not None:
            out[i] = nxt[1]
    for i in range(1, len(out)):
        if out[i] < out[i - 1]:
            out[i] = out[i - 1]
    return out


def trailing_plateau_start(start_times, tolerance=0.1, min_size=5):   # ForcedAligner.swift:191-216
    t = np.asarray(start_times, dtype=np.float32)
    n = t.size
    if n <= min_size:
        return n
    plateau = n
    for i in range(n - 1, 0, -1):
        if abs(np.float32(t[i] - t[i - 1])) < np.float32(tolerance):
            plateau = i - 1
        else:
            break
    return plateau if (n - plateau) >= min_size else n
