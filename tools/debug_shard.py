import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fractions import Fraction
import nanomod_b200 as nm
det = nm.Detector(0)

def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))

def model_moments(row, mode):
    """mode 'class': 4 accumulators by row index mod 4, combined (c0+c1)+(c2+c3); 'seq': sequential."""
    n = len(row)
    row = [float(x) for x in row]
    if mode == "seq":
        s = 0.0
        for x in row: s += x
        m = s / n
        ss = 0.0
        for x in row:
            d = x - m
            ss = fma(d, d, ss)
        return m, ss / (n - 1)
    s = [0.0] * 4
    for k, x in enumerate(row): s[k & 3] += x
    m = ((s[0] + s[1]) + (s[2] + s[3])) / n
    ss = [0.0] * 4
    for k, x in enumerate(row):
        d = x - m
        ss[k & 3] = fma(d, d, ss[k & 3])
    return m, ((ss[0] + ss[1]) + (ss[2] + ss[3])) / (n - 1)

def model_t(a, b, mode):
    m0, v0 = model_moments(a, mode); m1, v1 = model_moments(b, mode)
    vn0, vn1 = v0 / len(a), v1 / len(b)
    return (m0 - m1) / np.sqrt(vn0 + vn1)

p = nm.synthetic_pileup(6000, 20, 20, drop_frac1=0.01, two_strands=True, poisson=True, clip=(3, 60))
t0 = det.detect(p, nm.DetectOptions(testMethod="ks"))
sl = p.slice_rows(1, p.n_pos)
t1 = det.detect(sl, nm.DetectOptions(testMethod="ks"))
k0 = 1 if (p.off0[1]-p.off0[0] >= 5 and p.off1[1]-p.off1[0] >= 5) else 0
a, b = t0.t_stat[k0:], t1.t_stat
bad = np.nonzero(a != b)[0]
print("mismatch", len(bad), "of", len(a))
cnt = {"full==class": 0, "shifted==class": 0, "neither": 0}
for r in bad[:40]:
    i = t1.row_pos_index[r]
    ra, rb = sl.group(0, i), sl.group(1, i)
    mc = model_t(ra, rb, "class")
    ms = model_t(ra, rb, "seq")
    j = i + 1
    key = "full==class" if a[r] == mc else ("shifted==class" if b[r] == mc else "neither")
    cnt[key] += 1
    if r in bad[:10]:
        print(" r", r, "n", len(ra), len(rb), "shift full", p.off0[j] & 3, p.off1[j] & 3, "shift sl", sl.off0[i] & 3, sl.off1[i] & 3,
              "gpu full", repr(a[r]), "gpu sl", repr(b[r]), "class", repr(mc), "seq", repr(ms), "row%32", (r + k0) % 32, r % 32)
print(cnt)
good = np.nonzero(a == b)[0][:200]
print("good rows matching class model:", sum(a[r] == model_t(sl.group(0, t1.row_pos_index[r]), sl.group(1, t1.row_pos_index[r]), "class") for r in good), "of", len(good))
