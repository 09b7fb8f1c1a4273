"""TEST INFRASTRUCTURE ONLY — lane-by-lane Python model of the warp-sequential CCL of csrc/ccl_tiles.cu (cv_ccl_label).

The CUDA kernels are written against this model (same names, same order of operations) and tests/test_ccl_model_cpu.py
holds the model to cv2.connectedComponents (canonical labels: 1 + min linear index of the component), so that the
algorithm — run elements, per-row label propagation, the cross-lane reduction for runs that span several 32-pixel words,
seam unions, the single label-writing pass — is validated on the CPU before it runs on a GPU.

  pass A  one warp per tile of TW = 1024 x TH rows, lane = one 32-pixel word column, rows processed top to bottom:
          a run (maximal horizontal run inside the tile row) takes the label of a run it touches in the row above; a run that
          touches nothing becomes a root (label = its own first pixel); touching two different labels is a union (global
          union-find on the parent entries, which live in the label image at run-start pixels).
          Outputs: bit plane, first[word] = pass A's label of the word's first sub-run (an element of the same component, whether
          that sub-run starts in the word or entered from the left), parent entries for roots and for the second and later
          sub-runs of a word only.
  pass B  unions across tile seams (rows y = k TH, columns x = k TW).
  pass C  every pixel: root by pointer chasing from first[word] (first sub-run of a word) or from the sub-run's own parent
          entry (further sub-runs), label = root + 1.  Written once.
"""
from __future__ import annotations

import numpy as np

TW, LANES = 1024, 32
INF = 0x7FFFFFFF


def run_start(word: int, x: int) -> int:
    """start bit of the run of `word` that contains set bit x"""
    below = (~word) & ((1 << x) - 1) & 0xFFFFFFFF
    return below.bit_length()


def run_len(word: int, s: int) -> int:
    n = 0
    while s + n < 32 and (word >> (s + n)) & 1:
        n += 1
    return n


def sub_runs(word: int):
    """[(start bit, mask)] of the maximal runs inside one word"""
    out, s = [], 0
    while s < 32:
        if (word >> s) & 1:
            n = run_len(word, s)
            out.append((s, ((1 << n) - 1) << s))
            s += n
        else:
            s += 1
    return out


class UF:
    """parent entries in the label image: L[x] = parent + 1 (0 = no entry)"""

    def __init__(self, n):
        self.L = np.zeros(n, np.int64)

    def find(self, x):
        while True:
            y = int(self.L[x]) - 1
            assert y >= 0, "pointer into a pixel that is not a run start"
            if y == x:
                return x
            x = y

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a < b:
            self.L[b] = a + 1
        elif b < a:
            self.L[a] = b + 1


def pass_a_tile(bits, first, uf, W, x0, y0, th, conn, words_per_row):
    """bits[y][word] already filled.  Processes rows y0 .. y0+th-1 of the tile whose first word column is x0 // 32."""
    c0 = x0 // 32
    nl = min(LANES, words_per_row - c0)
    lab_prev = [dict() for _ in range(LANES)]  # per lane: start bit -> label (ancestor pixel index) of the previous row
    up = [0] * LANES
    for y in range(y0, y0 + th):
        w = [int(bits[y][c0 + c]) if c < nl else 0 for c in range(LANES)]
        wl = [0] + w[:-1]
        wr = w[1:] + [0]
        upl = [0] + up[:-1]
        upr = up[1:] + [0]
        cin = [bool((w[c] & 1) and (wl[c] >> 31)) for c in range(LANES)]
        cout = [bool((w[c] >> 31) and (wr[c] & 1)) for c in range(LANES)]
        full = [w[c] == 0xFFFFFFFF for c in range(LANES)]
        brk = [not (full[c] and cin[c]) for c in range(LANES)]  # the run through my bit 31 starts in my word
        origin = [max(s for s in range(c) if brk[s]) if cin[c] else -1 for c in range(LANES)]
        first_row = y == y0
        # ---- step A: candidate label of every word-local sub-run from the row above (unions among several touched labels)
        cand = [dict() for _ in range(LANES)]
        for c in range(LANES):
            for st, rm in sub_runs(w[c]):
                cd = INF
                if not first_row:
                    aw = rm | ((rm << 1) & 0xFFFFFFFF) | (rm >> 1) if conn == 8 else rm
                    ov = aw & up[c]
                    touched = []
                    while ov:
                        u = (ov & -ov).bit_length() - 1
                        us = run_start(up[c], u)
                        ov &= ~(((1 << run_len(up[c], us)) - 1) << us)
                        touched.append(lab_prev[c][us])
                    if conn == 8:
                        if (rm & 1) and (upl[c] >> 31):
                            touched.append(lab_prev[c - 1][run_start(upl[c], 31)])
                        if (rm >> 31) and (upr[c] & 1):
                            touched.append(lab_prev[c + 1][0])
                    for t in touched:
                        if cd == INF:
                            cd = t
                        elif t != cd:
                            uf.union(cd, t)
                            cd = min(cd, t)
                cand[c][st] = cd
        # ---- step B: runs that span words: minimum over the portions, unions between portions that disagree
        lab_cur = [dict() for _ in range(LANES)]
        for c in range(LANES):
            for st, rm in sub_runs(w[c]):
                is_head = st == 0 and cin[c]
                is_tail_org = bool(rm >> 31) and cout[c] and (not is_head or False)
                # the run this portion belongs to: (origin lane, start bit in the origin lane)
                if is_head:
                    o = origin[c]
                else:
                    o = c
                # members: tail portion of lane o (if it continues right) + head portions of lanes with origin == o
                ostart = run_start(w[o], 31) if (is_head or (bool(rm >> 31) and cout[c])) else st
                if not is_head and not (bool(rm >> 31) and cout[c]):
                    members = [(c, st)]
                else:
                    members = [(o, ostart)] + [(k, 0) for k in range(o + 1, LANES) if cin[k] and origin[k] == o]
                vals = [cand[k][s_] for k, s_ in members if cand[k][s_] != INF]
                if vals:
                    m = min(vals)
                    for v in vals:
                        if v != m:
                            uf.union(v, m)
                else:
                    m = y * W + x0 + o * 32 + ostart  # nothing touched: the run is a root, named by its first pixel
                lab_cur[c][st] = m
                # the lane where the run starts writes its parent entry — roots and sub-runs after the first of their word only
                # (UF.find asserts on a pixel without an entry, so a read of a skipped entry fails the model's tests)
                if not is_head:
                    first_st = (w[c] & -w[c]).bit_length() - 1
                    if not vals or st != first_st:
                        uf.L[y * W + x0 + c * 32 + st] = m + 1
            if c < nl and w[c]:
                first[y][c0 + c] = lab_cur[c][(w[c] & -w[c]).bit_length() - 1]
        lab_prev, up = lab_cur, w


def label(mask: np.ndarray, conn: int = 8, th: int = 32) -> np.ndarray:
    """Model of cv_ccl_label on one image.  Returns int32 labels: 1 + min linear index of the component, 0 = background."""
    H, W = mask.shape
    wpr = (W + 31) // 32
    padded = np.zeros((H, wpr * 32), np.uint8)
    padded[:, :W] = mask != 0
    bits = np.packbits(padded.reshape(H, wpr, 32), axis=2, bitorder="little").view(np.uint32).reshape(H, wpr)
    first = np.full((H, wpr), -1, np.int64)
    uf = UF(H * W)
    for y0 in range(0, H, th):
        for x0 in range(0, W, TW):
            pass_a_tile(bits, first, uf, W, x0, y0, min(th, H - y0), conn, wpr)

    def start_of(y, c, bit):
        """label (an element with a parent entry) of the sub-run of word (y, c) containing `bit`: cs_label_of in the kernel"""
        wv = int(bits[y][c])
        st = run_start(wv, bit)
        if st == (wv & -wv).bit_length() - 1:
            return int(first[y][c])
        e = int(uf.L[y * W + c * 32 + st])
        assert e > 0, "sub-run without a parent entry"
        return e - 1

    # ---- pass B: seams
    for y in range(th, H, th):  # horizontal: row y against row y - 1
        for c in range(wpr):
            wv, uv = int(bits[y][c]), int(bits[y - 1][c])
            ul = int(bits[y - 1][c - 1]) >> 31 if c > 0 else 0
            ur = int(bits[y - 1][c + 1]) & 1 if c + 1 < wpr else 0
            for st, rm in sub_runs(wv):
                aw = rm | ((rm << 1) & 0xFFFFFFFF) | (rm >> 1) if conn == 8 else rm
                ov = aw & uv
                while ov:
                    u = (ov & -ov).bit_length() - 1
                    us = run_start(uv, u)
                    ov &= ~(((1 << run_len(uv, us)) - 1) << us)
                    uf.union(start_of(y, c, st), start_of(y - 1, c, us))
                if conn == 8:
                    if (rm & 1) and ul:
                        uf.union(start_of(y, c, st), start_of(y - 1, c - 1, 31))
                    if (rm >> 31) and ur:
                        uf.union(start_of(y, c, st), start_of(y - 1, c + 1, 0))
    for x in range(TW, W, TW):  # vertical: column x against column x - 1
        c = x // 32
        for y in range(H):
            if not (int(bits[y][c]) & 1):
                continue
            me = start_of(y, c, 0)
            if int(bits[y][c - 1]) >> 31:
                uf.union(me, start_of(y, c - 1, 31))
            elif conn == 8:
                if y > 0 and (int(bits[y - 1][c - 1]) >> 31):
                    uf.union(me, start_of(y - 1, c - 1, 31))
                if y + 1 < H and (int(bits[y + 1][c - 1]) >> 31):
                    uf.union(me, start_of(y + 1, c - 1, 31))
        # (the mirrored diagonals — pixel (x-1, y) against (x, y +- 1) — are covered by the horizontal-seam / in-tile logic only
        # when they do not cross the vertical seam; handle them here)
        if conn == 8:
            for y in range(H):
                if not (int(bits[y][c - 1]) >> 31):
                    continue
                me = start_of(y, c - 1, 31)
                if int(bits[y][c]) & 1:
                    continue  # already united through the horizontal neighbour
                if y > 0 and (int(bits[y - 1][c]) & 1):
                    uf.union(me, start_of(y - 1, c, 0))
                if y + 1 < H and (int(bits[y + 1][c]) & 1):
                    uf.union(me, start_of(y + 1, c, 0))
    # ---- pass C
    out = np.zeros((H, W), np.int32)
    for y in range(H):
        for c in range(wpr):
            for k, (st, rm) in enumerate(sub_runs(int(bits[y][c]))):
                root = uf.find(start_of(y, c, st))
                n = run_len(int(bits[y][c]), st)
                x = c * 32 + st
                out[y, x:min(W, x + n)] = root + 1
    return out
