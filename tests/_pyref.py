"""Independent pure-Python restatement of the path's *definitions*, for small cases only.

Written separately from oracle/*.c (line-based FASTA parser, big-int arithmetic, and
ProbMinHash3a as the order-free definition  sig[k] = argmin_{(d,i): k_i(d)=k} (i-1+x_i(d))/w_d
rather than Ertl's two-pass loop) so that agreement pins the oracle's control flow, and the
claim the GPU design relies on: the result does not depend on processing order or on when the
early-stop bound is read."""
import math
import struct

M64 = (1 << 64) - 1
AA = "ACDEFGHIKLMNPQRSTVWY"


def splitmix(state):
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return state, z ^ (z >> 31)


def rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


class Xoshiro:
    def __init__(self, seed=None, state=None):
        if state is not None:
            self.s = list(state)
        else:
            self.s = []
            st = seed & M64
            for _ in range(4):
                st, v = splitmix(st)
                self.s.append(v)

    def next(self):
        s = self.s
        r = (rotl((s[0] + s[3]) & M64, 23) + s[0]) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = rotl(s[3], 45)
        return r

    def f64(self):
        bits = (self.next() >> 12) | 0x3FF0000000000000
        return struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0

    def f32(self):
        bits = ((self.next() >> 32) >> 9) | 0x3F800000
        return struct.unpack("<f", struct.pack("<I", bits))[0] - 1.0  # exact in f32

    def usize(self, m):
        zone = M64 - ((M64 - m + 1) % m)
        while True:
            v = self.next() * m
            if (v & M64) <= zone:
                return v >> 64


class Exp01:
    def __init__(self, lam):
        self.lam = lam
        self.c1 = math.expm1(lam) / lam
        self.c2 = math.log(2.0 / (1.0 + math.exp(-lam))) / lam
        self.c3 = (1.0 - math.exp(-lam)) / lam

    def sample(self, rng):
        x = self.c1 * rng.f64()
        if x < 1.0:
            return x
        while True:
            x = rng.f64()
            if x < self.c2:
                return x
            y = 0.5 * rng.f64()
            if y > 1.0 - x:
                x = 1.0 - x
                y = 1.0 - y
            if x <= self.c3 * (1.0 - y):
                return x
            if self.c1 * y <= 1.0 - x:
                return x
            if y * self.c1 * self.lam <= math.expm1(self.lam * (1.0 - x)):
                return x


def parse_fasta(data: bytes, data_t=0, block=False):
    """-> list of code lists.  Line-based: a line starting with '>' opens a record."""
    if not data:
        return [[]] if block else []
    assert data[:1] == b">", "not FASTA"
    recs = []
    cur = None
    for line in data.split(b"\n"):
        if line[:1] == b">":
            cur = [line[1:], []]
            recs.append(cur)
        else:
            cur[1].append(line)
    out = []
    blk = []
    for hdr, lines in recs:
        if b"capsid" in hdr:
            continue
        seq = b"".join(lines)
        if data_t == 0:
            codes = ["ACGT".index(chr(c).upper()) for c in seq if chr(c).upper() in "ACGT"]
        else:
            codes = [AA.index(chr(c)) + 1 for c in seq if chr(c) in AA]
        if block:
            blk.extend(codes)
        elif codes:
            out.append(codes)
    return [blk] if block else out


def kmers(seqs, data_t, k):
    vals = []
    for codes in seqs:
        for i in range(len(codes) - k + 1):
            w = codes[i:i + k]
            if data_t == 0:
                f = 0
                r = 0
                for b in w:
                    f = (f << 2) | b
                for b in reversed(w):
                    r = (r << 2) | (3 - b)
                vals.append(min(f, r))
            else:
                v = 0
                for c in w:
                    v = (v << 5) | c
                vals.append(v)
    return vals


def nohash_seed(v, val_bytes, identity=False):
    if identity:
        return v
    return int.from_bytes(v.to_bytes(val_bytes, "little"), "big")


def probminhash3a_definition(vals, m, val_bytes, identity=False, hcut=None):
    """sig[k] = argmin over all points with h < hcut; asserts hcut covered every slot"""
    counts = {}
    for v in vals:
        counts[v] = counts.get(v, 0) + 1
    e01 = Exp01(math.log(m / (m - 1)))
    best = [(float("inf"), 0)] * m
    if not counts:
        return [0] * m, best
    total = sum(counts.values())
    if hcut is None:
        hcut = 4.0 * (m / total) * (math.log(m) + 8.0)
    for d, w in counts.items():
        rng = Xoshiro(nohash_seed(d, val_bytes, identity))
        winv = 1.0 / float(w)
        i = 1
        while True:
            h0 = winv * float(i - 1)
            if not h0 < hcut:
                break
            x = e01.sample(rng)
            h = winv * x if i == 1 else h0 + winv * x
            if i == 1 and not h < hcut:
                break  # the reference draws k_1 only when h_1 is below the bound
            k = rng.usize(m)
            if h < hcut and (h, d) < best[k]:
                best[k] = (h, d)
            i += 1
    assert max(b[0] for b in best) < hcut, "hcut too small for this input"
    return [b[1] for b in best], best


def optdens_definition(vals, m, f64_draw=False):
    import numpy as np
    large = np.float32(4294967296.0)
    sk = [large] * m
    for v in set(vals):
        h = (v * 0x517CC1B727220A95) & M64
        rng = Xoshiro(h)
        if f64_draw:
            r = np.float32(rng.f64())
        else:
            r = np.float32(rng.f32())
        k = rng.usize(m)
        if r <= sk[k]:
            sk[k] = r
    empty = [s > 1.5 for s in sk]
    if any(empty) and not all(empty):
        filled = list(sk)
        for k in range(m):
            if empty[k]:
                rng = Xoshiro(k)
                while True:
                    j = rng.usize(m)
                    if not empty[j]:
                        filled[k] = sk[j]
                        break
        sk = filled
    return np.array(sk, dtype=np.float32)


def hamming(a, b):
    import numpy as np
    a = np.asarray(a)
    b = np.asarray(b)
    return np.float32(np.float32((a != b).sum()) / np.float32(len(a)))
