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


def expm1_spec(z):
    """the frozen expm1 of the sampler's rejection branch (oracle/rng.c gso_expm1_spec)"""
    r = 1.0
    for k in range(24, 1, -1):
        r = 1.0 + (z / float(k)) * r
    return z * r


def ln_spec(x):
    """the frozen natural logarithm of the SetSketch path (oracle/rng.c gso_ln_spec); Python floats
    are IEEE doubles and every line below is one rounded operation, like the C source"""
    import struct
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    Lg1, Lg2, Lg3, Lg4 = 6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01
    Lg5, Lg6, Lg7 = 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    e = ((bits >> 52) & 0x7FF) - 1023
    m = struct.unpack("<d", struct.pack("<Q", (bits & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000))[0]
    if m > 1.4142135623730951:
        m = m * 0.5
        e += 1
    f = m - 1.0
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6))
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)))
    R = t2 + t1
    hfsq = 0.5 * f * f
    dk = float(e)
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f)


def setsketch_definition(vals, m):
    """SetSketch1 as a plain definition (no early stop): register i = max over the items and their
    points of clamp(floor(1 - log_b x_j), 0, 65535) for the point that the item's lazy Fisher-Yates
    permutation sends to i; b = 1.001, a = 20"""
    lnb = ln_spec(1.001)
    inva = 1.0 / 20.0
    reg = [0] * m
    for v in set(vals):
        rng = Xoshiro((v * 0x517CC1B727220A95) & M64)
        perm = list(range(m))
        x = 0.0
        for j in range(m):
            e = -ln_spec(1.0 - rng.f64())
            x = x + (inva / float(m - j)) * e
            if x > 0.0:
                k = math.floor(1.0 - ln_spec(x) / lnb)
                k = min(max(k, 0), 65535)
            else:
                k = 65535
            if k <= 0:
                break
            r = j + rng.usize(m - j)
            perm[j], perm[r] = perm[r], perm[j]
            if k > reg[perm[j]]:
                reg[perm[j]] = k
    return reg


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
            if y * self.c1 * self.lam <= expm1_spec(self.lam * (1.0 - x)):
                return x


def parse_fasta(data: bytes, data_t=0, block=False):
    """-> list of code lists.  Line-based: a line starting with '>' opens a record."""
    if not data:
        return [[]] if block else []
    if data[:1] == b"@":
        # FASTQ as needletail reads it: four lines per record, the sequence on ONE line
        lines = data.split(b"\n")
        while lines and lines[-1].strip(b"\r") == b"":
            lines.pop()  # trailing terminators
        recs = []
        for r in range(0, len(lines), 4):
            quad = [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines[r:r + 4]]
            assert quad[0][:1] == b"@" and len(quad[0]) > 1, "InvalidStart"
            assert len(quad) >= 3 and quad[2][:1] == b"+", "InvalidSeparator"
            recs.append([quad[0][1:], [quad[1]]])
        return _encode_records(recs, data_t, block)
    assert data[:1] == b">", "not FASTA"
    recs = []
    cur = None
    for line in data.split(b"\n"):
        if line[:1] == b">":
            cur = [line[1:], []]
            recs.append(cur)
        else:
            cur[1].append(line)
    return _encode_records(recs, data_t, block)


def _encode_records(recs, data_t, block):
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


def superminhash_definition(vals, m):
    """Ertl's SuperMinHash as a plain definition, without the a_upper early stop or the lazy
    permutation bookkeeping: every distinct item runs its whole Fisher-Yates permutation
    (j = 0..m-1: r_j, k_j = j + U(m-j), swap) and sig[slot] = min over items of r_j + j."""
    import numpy as np
    sig = [np.float32(4294967296.0)] * m
    for v in set(vals):
        rng = Xoshiro((v * 0x517CC1B727220A95) & M64)
        perm = list(range(m))
        for j in range(m):
            r = np.float32(rng.f32())
            k = j + rng.usize(m - j)
            perm[j], perm[k] = perm[k], perm[j]
            val = np.float32(r + np.float32(j))
            if val < sig[perm[j]]:
                sig[perm[j]] = val
    return np.array(sig, dtype=np.float32)


def superminhash2_definition(vals, m, kt32):
    """SuperMinHash2 as a plain definition (no early stop, no lazy permutation): per slot the fx hash
    of the item with the smallest r_j + j that landed there (identical values: the smaller hash)"""
    best = [(float("inf"), 0xFFFFFFFF if kt32 else M64)] * m
    for v in set(vals):
        hv = ((v & 0xFFFFFFFF) * 0x9E3779B9) & 0xFFFFFFFF if kt32 else (v * 0x517CC1B727220A95) & M64
        rng = Xoshiro(hv)
        perm = list(range(m))
        for j in range(m):
            r = rng.f64()
            k = j + rng.usize(m - j)
            perm[j], perm[k] = perm[k], perm[j]
            cand = (r + float(j), hv)
            if cand < best[perm[j]]:
                best[perm[j]] = cand
    return [b[1] for b in best]


def revdens_target(i, a, m):
    z = (i * 0x9E3779B97F4A7C15 + a * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return (z * m) >> 64


def revoptdens_definition(vals, m):
    """OptDens bins, then reverse densification: round a, every originally non-empty bin i in
    increasing order pushes its value into bin target(i, a) if that bin is still empty"""
    import numpy as np
    large = np.float32(4294967296.0)
    sk = [large] * m
    for v in set(vals):
        rng = Xoshiro((v * 0x517CC1B727220A95) & M64)
        r = np.float32(rng.f32())
        k = rng.usize(m)
        if r <= sk[k]:
            sk[k] = r
    orig = [s <= 1.5 for s in sk]
    left = m - sum(orig)
    a = 0
    while 0 < left < m:
        for i in range(m):
            if orig[i] and left:
                j = revdens_target(i, a, m)
                if sk[j] > 1.5:
                    sk[j] = sk[i]
                    left -= 1
        a += 1
    return np.array(sk, dtype=np.float32)


def hamming(a, b):
    import numpy as np
    a = np.asarray(a)
    b = np.asarray(b)
    return np.float32(np.float32((a != b).sum()) / np.float32(len(a)))


# --------------------------------------------------------------------------- HNSW (small cases)
# Written from the published algorithm (Malkov & Yashunin) and the documented behaviour of Rust's
# std::collections::BinaryHeap, not from oracle/hnsw.c: Python lists, (key, point) tuples, the wave
# semantics of DESIGN.md section 2.  Agreement with the C oracle pins its heap operations, tie
# handling, neighbour selection and wave bookkeeping.
import numpy as _np


class RustHeap:
    """max-heap on key with std BinaryHeap's exact element movements"""

    def __init__(self):
        self.d = []

    def __len__(self):
        return len(self.d)

    def _sift_up(self, start, pos):
        e = self.d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if e[0] <= self.d[parent][0]:
                break
            self.d[pos] = self.d[parent]
            pos = parent
        self.d[pos] = e

    def push(self, key, p):
        self.d.append((key, p))
        self._sift_up(0, len(self.d) - 1)

    def pop(self):
        item = self.d.pop()
        if self.d:
            item, self.d[0] = self.d[0], item
            end, pos, e = len(self.d), 0, self.d[0]
            child = 1
            while child <= max(end - 2, 0) and end >= 2:
                if self.d[child][0] <= self.d[child + 1][0]:
                    child += 1
                self.d[pos] = self.d[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                self.d[pos] = self.d[child]
                pos = child
            self.d[pos] = e
            self._sift_up(0, pos)
        return item

    def peek(self):
        return self.d[0]

    def into_sorted(self):
        d = self.d
        end = len(d)
        while end > 1:
            end -= 1
            d[0], d[end] = d[end], d[0]
            pos, e, child = 0, d[0], 1
            while end >= 2 and child <= end - 2:
                if d[child][0] <= d[child + 1][0]:
                    child += 1
                if e[0] >= d[child][0]:
                    break
                d[pos] = d[child]
                pos = child
                child = 2 * pos + 1
            else:
                if child == end - 1 and e[0] < d[child][0]:
                    d[pos] = d[child]
                    pos = child
            d[pos] = e
        return d


class PyHnsw:
    def __init__(self, M, ef_c, max_layer=16, scale=1.0, extend=True, seed=0x5EED):
        self.M, self.ef_c, self.max_layer, self.extend = M, ef_c, max_layer, extend
        self.scale = scale / math.log(M)
        self.rng = Xoshiro(seed)
        self.pts, self.ids, self.level, self.rank, self.nbrs = [], [], [], [], []
        self.layer_count = [0] * 32
        self.entry = -1

    def dist(self, a, b):
        return _np.float32(int((_np.asarray(a) != _np.asarray(b)).sum())) / _np.float32(len(a))

    def gen_level(self):
        lv = -math.log(self.rng.f64()) * self.scale
        if not lv < self.max_layer:
            return self.rng.usize(self.max_layer)
        return int(math.floor(lv))

    def search_layer(self, q, ep, ef, layer):
        visited = {ep}
        d0 = self.dist(q, self.pts[ep])
        cand, ret = RustHeap(), RustHeap()
        cand.push(-d0, ep)
        ret.push(d0, ep)
        while len(cand):
            c = cand.pop()
            if -c[0] > ret.peek()[0]:
                break
            for e, _ in self.nbrs[c[1]][layer]:
                if e in visited:
                    continue
                visited.add(e)
                ed = self.dist(q, self.pts[e])
                if ed < ret.peek()[0] or len(ret) < ef:
                    cand.push(-ed, e)
                    ret.push(ed, e)
                    if len(ret) > ef:
                        ret.pop()
        return ret

    def select(self, q, cand, nb_asked, extend_asked, layer):
        out = []
        extend = False
        if len(cand) <= nb_asked:
            if not extend_asked:
                while len(cand):
                    k, p = cand.pop()
                    out.append((p, -k))
                return out
            extend = True
        if extend:
            seen = {p for _, p in cand.d}
            new = []
            for _, p in list(cand.d):
                for e, _ in self.nbrs[p][layer]:
                    if e not in seen:
                        seen.add(e)
                        new.append(e)
            for e in new:
                cand.push(-self.dist(q, self.pts[e]), e)
        while len(cand) and len(out) < nb_asked:
            k, p = cand.pop()
            ed = -k
            if all(not (self.dist(self.pts[p], self.pts[s]) <= ed) for s, _ in out):
                out.append((p, ed))
        return out

    def insert_waves(self, sigs, ids, wave_max):
        i, n = 0, len(sigs)
        while i < n:
            if self.entry < 0:
                self._add(sigs[i], ids[i])
                self.entry = 0
                i += 1
                continue
            W = min(max(len(self.pts) // 4, 1), wave_max, n - i)
            first = len(self.pts)
            for t in range(W):
                self._add(sigs[i + t], ids[i + t])
            entry = self.entry
            sels = [self._phase_a(first + t, first, entry) for t in range(W)]
            for t in range(W):
                for l, sel in sels[t].items():
                    self.nbrs[first + t][l] = list(sel)
            for t in range(W):
                np_ = first + t
                for l in sorted(sels[t], reverse=True):
                    for qp, d in sels[t][l]:
                        if qp == np_ or l > self.level[qp]:
                            continue
                        lst = self.nbrs[qp][l]
                        if any(x == np_ for x, _ in lst):
                            continue
                        lst.append((np_, d))
                        lst.sort(key=lambda x: (x[1], x[0]))
                        if len(lst) > (self.M if l > 0 else 2 * self.M):
                            lst.pop()
                if self.level[np_] > self.level[self.entry]:
                    self.entry = np_
            i += W

    def _add(self, sig, pid):
        lv = self.gen_level()
        self.pts.append(_np.array(sig))
        self.ids.append(int(pid))
        self.level.append(lv)
        self.rank.append(self.layer_count[lv])
        self.layer_count[lv] += 1
        self.nbrs.append([[] for _ in range(lv + 1)])

    def _phase_a(self, np_, first, entry):
        q, level = self.pts[np_], self.level[np_]
        ep, lmax = entry, self.level[entry]
        d_ep = self.dist(q, self.pts[ep])
        for l in range(lmax, level, -1):
            ret = self.search_layer(q, ep, 1, l)
            if len(ret):
                d, p = ret.pop()
                if d < d_ep:
                    ep, d_ep = p, d
        sel = {}
        for l in range(min(level, lmax), -1, -1):
            ret = self.search_layer(q, ep, self.ef_c, l)
            for m in range(first, np_):
                if self.level[m] < l:
                    continue
                ed = self.dist(q, self.pts[m])
                if ed < ret.peek()[0] or len(ret) < self.ef_c:
                    ret.push(ed, m)
                    if len(ret) > self.ef_c:
                        ret.pop()
            cand = RustHeap()
            for d, p in ret.d:
                cand.push(-d, p)
            out = self.select(q, cand, 2 * self.M if l == 0 else self.M, l == 0 and self.extend, l)
            out.sort(key=lambda x: (x[1], x[0]))
            sel[l] = out
            for p, _ in out:
                if p < first:
                    ep = p
                    break
        return sel

    def search(self, q, knbn, ef_arg):
        if self.entry < 0:
            return []
        pivot = self.entry
        dte = self.dist(q, self.pts[pivot])
        for layer in range(self.level[pivot], 0, -1):
            newp = pivot
            for e, _ in self.nbrs[pivot][layer]:
                d = self.dist(q, self.pts[e])
                if d < dte:
                    dte, newp = d, e
            pivot = newp
        ef = max(ef_arg, knbn)
        ret = self.search_layer(q, pivot, ef, 0)
        srt = ret.into_sorted()
        return [(self.ids[p], float(d)) for d, p in srt[:min(knbn, ef)]]
