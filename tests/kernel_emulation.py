"""NumPy emulation of the device-side index math of the sm_100a kernels (test infrastructure).

These functions mirror, statement for statement, the address arithmetic of
gpuaudiobench_b200/csrc/direct_fir.cu (swizzle, ring/tile geometry, per-lane block walk) and
upols.cu (Stockham passes, packed-bin real-FFT pre/post passes, FDL-MAC with the {DC, Nyquist}
packing, ring-slot rotation).  They let the CPU test-suite prove the index conventions of
SURVEY.md App. E without a GPU; they are NOT a product path and are never imported outside tests/.
Arithmetic is float64 here: these check indexing, the GPU tests check fp32 values.
"""
import numpy as np

KFIR_WARPS = 8


def swz_chunk(f):
    return f ^ ((f >> 2) & 7)


def swz_float(n):
    return (swz_chunk(n >> 2) << 2) | (n & 3)


def load_block(tile, blk):
    """Mirror of load_block(): 16 floats of logical block `blk` from a swizzled tile."""
    m = blk & 3
    hi = ((blk ^ (blk >> 2)) & 1) << 2
    row = (blk >> 1) << 3
    out = np.empty(16)
    for i in range(4):
        c = row + (hi | (i ^ m))
        out[4 * i:4 * i + 4] = tile[4 * c:4 * c + 4]
    return out


def toeplitz_tile(acc, hv, lo, hi):
    w = np.concatenate([lo, hi])
    for s in range(16):
        for r in range(16):
            acc[r] += hv[s] * w[16 + r - s]


def fir_cta_of_unit(u, U, G):
    return ((u + 1) * G + U - 1) // U - 1


class DirectEmu:
    """Emulates ring_append_kernel + the persistent fir_direct_kernel<A> + fir_finish_mix_kernel."""

    def __init__(self, T, B, L, plan):
        self.T, self.B, self.L = T, B, L
        self.p = plan
        self.cap = plan["cap"]
        self.ring = np.zeros((T, self.cap))
        self.h = np.zeros((T, plan["Lc"] * 16))
        self.pos = 0

    def load_ir(self, ir):
        self.h[:] = 0
        if self.p["CL"] > 1:
            for j in range(self.L):
                self.h[:, swz_float(j)] = ir[:, j]
        else:
            self.h[:, :self.L] = ir  # taps are a broadcast read: stored linear

    def prime(self, hist):
        self.ring[:] = 0
        H = self.L - 1
        for i in range(H):
            self.ring[:, swz_float(self.cap - H + i)] = hist[:, i]
        self.pos = 0

    def _load_taps(self, hs, blk):
        return load_block(hs, blk) if self.p["CL"] > 1 else hs[16 * blk:16 * blk + 16]

    def process(self, x, commit=True):
        p, T, B = self.p, self.T, self.B
        A, CL, SPS, JSb, NS, G, MS, ntiles = p["A"], p["CL"], p["SPS"], p["JSb"], p["NS"], p["G"], p["MS"], p["ntiles"]
        capb, posb = self.cap // 16, self.pos // 16
        partial = np.zeros((MS, T, B))
        written = np.zeros((MS, T, B), dtype=bool)
        OT = A * 16
        U = T * ntiles * NS
        for cta in range(G):
            u_lo, u_hi = cta * U // G, (cta + 1) * U // G
            w, k = divmod(u_lo, NS)
            seg = cta - fir_cta_of_unit(w * NS, U, G)
            assert 0 <= seg < MS
            acc = np.zeros((KFIR_WARPS, 32, 16))
            for it in range(u_hi - u_lo):
                t, ot = divmod(w, ntiles)
                c0 = k * JSb
                qbase = posb + capb + ot * A
                qs = (qbase - c0 - JSb) & ~7
                # producer: history part of the tile from the ring (the current buffer is not in it yet)
                q_hi = min(qbase + A - 1 - c0, posb + capb - 1)
                nblk = max(0, q_hi - qs + 1)  # 0: an upper output tile's first stage has no history part
                assert nblk <= p["xtile_blocks"], (nblk, p["xtile_blocks"])
                src_b = qs % capb
                first = min(nblk, capb - src_b)
                hs = self.h[t, c0 * 16:(c0 + JSb) * 16]
                xs = np.full(p["xtile_blocks"] * 16, np.nan)
                xs[:first * 16] = self.ring[t, src_b * 16:(src_b + first) * 16]
                if first < nblk:
                    xs[first * 16:nblk * 16] = self.ring[t, :(nblk - first) * 16]
                # consumers: current-buffer blocks [cur_lo, cur_hi] straight from d_in, swizzled
                cur_lo, cur_hi = max(qs, posb + capb), qbase + A - 1 - c0
                if cur_lo <= cur_hi:
                    f0 = (cur_lo - qs) * 4
                    for c in range((cur_hi - cur_lo + 1) * 4):
                        src = (cur_lo - posb - capb) * 16 + 4 * c
                        pf = swz_chunk(f0 + c)
                        xs[4 * pf:4 * pf + 4] = x[t, src:src + 4]
                nblk = cur_hi - qs + 1
                for warp in range(KFIR_WARPS):
                    for lane in range(32):
                        a, g = lane & (A - 1), lane // A
                        hb0 = (warp * CL + g) * SPS
                        sb = qbase + a - (c0 + hb0) - qs
                        assert sb - SPS >= 0 and sb < nblk
                        Q = load_block(xs, sb)
                        for q in range(0, SPS, 2):
                            hv = self._load_taps(hs, hb0 + q)
                            P = load_block(xs, sb - q - 1)
                            toeplitz_tile(acc[warp, lane], hv, P, Q)
                            hv = self._load_taps(hs, hb0 + q + 1)
                            Q = load_block(xs, sb - q - 2)
                            toeplitz_tile(acc[warp, lane], hv, Q, P)
                if k + 1 == NS or it + 1 == u_hi - u_lo:  # flush the (CTA, tile) segment
                    row = np.zeros(OT)
                    for warp in range(KFIR_WARPS):
                        for a in range(A):
                            row[a * 16:(a + 1) * 16] += sum(acc[warp, g * A + a] for g in range(CL))
                    assert not written[seg, t, ot * OT:(ot + 1) * OT].any(), "two segments wrote one partial row"
                    partial[seg, t, ot * OT:(ot + 1) * OT] = row
                    written[seg, t, ot * OT:(ot + 1) * OT] = True
                    acc[:] = 0
                    seg = 0
                k += 1
                if k == NS:
                    k, w = 0, w + 1
        if commit:  # fir_finish_mix_kernel appends the consumed buffer
            for n in range(B):
                self.ring[:, swz_float(self.pos + n)] = x[:, n]
            self.pos = (self.pos + B) % self.cap
        assert not np.isnan(partial).any(), "a lane read a tile block nobody staged"
        return partial.sum(axis=0)


# ------------------------------------------------------------------------------------------------
# UPOLS
# ------------------------------------------------------------------------------------------------
def fft_stockham(a, inverse, tw):
    """Mirror of fft_stockham<INV>: complex FFT of a (length M = 2^k) with radix-2/4 Stockham passes."""
    M = a.size
    logM = M.bit_length() - 1
    a = a.astype(np.complex128).copy()
    b = np.empty_like(a)
    Ns = 1
    if logM & 1:
        half = M >> 1
        for j in range(half):
            v0, v1 = a[j], a[j + half]
            b[2 * j] = v0 + v1
            b[2 * j + 1] = v0 - v1
        a, b = b, a
        Ns = 2
    quarter = M >> 2
    rot = 1j if inverse else -1j
    while Ns < M:
        tstride = M // (4 * Ns)
        for j in range(quarter):
            k = j & (Ns - 1)
            v0, v1, v2, v3 = a[j], a[j + quarter], a[j + 2 * quarter], a[j + 3 * quarter]
            if k != 0:
                w1, w2, w3 = tw[k * tstride], tw[2 * k * tstride], tw[3 * k * tstride]
                if inverse:
                    w1, w2, w3 = np.conj(w1), np.conj(w2), np.conj(w3)
                v1, v2, v3 = v1 * w1, v2 * w2, v3 * w3
            t0, t1, t2, d = v0 + v2, v0 - v2, v1 + v3, v1 - v3
            t3 = rot * d
            j0 = ((j - k) << 2) + k
            b[j0], b[j0 + Ns], b[j0 + 2 * Ns], b[j0 + 3 * Ns] = t0 + t2, t1 + t3, t0 - t2, t1 - t3
        a, b = b, a
        Ns <<= 2
    return a


def twiddles(M):
    tw_c = np.exp(-2j * np.pi * np.arange(M) / M)
    tw_r = np.exp(-2j * np.pi * np.arange(M // 2 + 1) / (2 * M))
    return tw_c, tw_r


def rfft_fwd_packed(first, second, scale, tw_c, tw_r):
    """Mirror of rfft_fwd_kernel for one window [first | second] of 2M reals -> M packed bins."""
    M = first.size
    half = M >> 1
    w = np.concatenate([first, second])
    a = w[0::2] + 1j * w[1::2]  # a[n] = (w[2n], w[2n+1]): n < M/2 from first, else second
    z = fft_stockham(a, False, tw_c)
    out = np.zeros(M, dtype=np.complex128)
    for k in range(half + 1):
        A_ = z[k]
        Bc = np.conj(z[(M - k) & (M - 1)])
        E = 0.5 * (A_ + Bc)
        D = 0.5 * (A_ - Bc)
        O = complex(D.imag, -D.real)
        WO = tw_r[k] * O
        Xk = E + WO
        Xmk = complex(E.real - WO.real, -(E.imag - WO.imag))
        if k == 0:
            out[0] = complex(scale * Xk.real, scale * Xmk.real)
        else:
            out[k] = scale * Xk
            if k != half:
                out[M - k] = scale * Xmk
    return out


def mac_packed(H, X, slot0):
    """Mirror of fdl_mac_kernel: H, X [P][M] packed; returns packed Y[M]."""
    P, M = H.shape
    A_ = np.zeros(M)
    B_ = np.zeros(M)
    C_ = np.zeros(M)
    D_ = np.zeros(M)
    for p in range(P):
        sl = slot0 + p
        if sl >= P:
            sl -= P
        h, x = H[p], X[sl]
        A_ += h.real * x.real
        B_ += h.imag * x.imag
        C_ += h.real * x.imag
        D_ += h.imag * x.real
    y = (A_ - B_) + 1j * (C_ + D_)
    y[0] = complex(A_[0], B_[0])
    return y


def irfft_ols_packed(Y, tw_c, tw_r):
    """Mirror of irfft_ols_kernel: packed Y[M] -> the B = M output samples y[B..2B)."""
    M = Y.size
    half = M >> 1
    a = np.zeros(M, dtype=np.complex128)
    for k in range(half + 1):
        mk = (M - k) & (M - 1)
        yk, ym = Y[k], Y[mk]
        if k == 0:
            A_, Bm = complex(yk.real, 0), complex(yk.imag, 0)
        else:
            A_, Bm = yk, ym
        Bc = np.conj(Bm)
        E, D = A_ + Bc, A_ - Bc
        O = np.conj(tw_r[k]) * D
        a[k] = complex(E.real - O.imag, E.imag + O.real)
        if k != 0 and k != half:
            a[M - k] = complex(E.real + O.imag, O.real - E.imag)
    z = fft_stockham(a, True, tw_c)
    out = np.empty(M)
    out[0::2] = z[half:].real
    out[1::2] = z[half:].imag
    return out


class UpolsEmu:
    """Emulates load_ir / prime_history / process of the UPOLS engine for T tracks."""

    def __init__(self, T, B, L):
        self.T, self.B, self.L = T, B, L
        self.P = (L + B - 1) // B
        self.M = B
        self.tw_c, self.tw_r = twiddles(B)
        self.H = np.zeros((T, self.P, B), dtype=np.complex128)
        self.X = np.zeros((T, self.P, B), dtype=np.complex128)
        self.prev = np.zeros((T, B))
        self.m = 0

    def load_ir(self, ir):
        B, P = self.B, self.P
        pad = np.zeros((self.T, P * B))
        pad[:, :self.L] = ir
        zeros = np.zeros(B)
        for t in range(self.T):
            for p in range(P):
                self.H[t, p] = rfft_fwd_packed(pad[t, p * B:(p + 1) * B], zeros, 1.0 / (2 * B), self.tw_c, self.tw_r)
        self.reset()

    def reset(self):
        self.X[:] = 0
        self.prev[:] = 0
        self.m = 0

    def prime(self, hist):
        self.reset()
        B, P, Hn = self.B, self.P, self.L - 1
        row = P * B
        hb = np.zeros((self.T, row))
        hb[:, row - Hn:] = hist
        for j in range(-(P - 1), 0):
            for t in range(self.T):
                self.X[t, -j] = rfft_fwd_packed(hb[t, (P + j - 1) * B:(P + j) * B], hb[t, (P + j) * B:(P + j + 1) * B],
                                                1.0, self.tw_c, self.tw_r)
        self.prev[:] = hb[:, (P - 1) * B:]

    def process(self, x, commit=True):
        P = self.P
        slot0 = (P - (self.m % P)) % P
        y = np.empty((self.T, self.B))
        for t in range(self.T):
            self.X[t, slot0] = rfft_fwd_packed(self.prev[t], x[t], 1.0, self.tw_c, self.tw_r)
            Y = mac_packed(self.H[t], self.X[t], slot0)
            y[t] = irfft_ols_packed(Y, self.tw_c, self.tw_r)
        if commit:
            self.prev[:] = x
            self.m += 1
        return y


# ------------------------------------------------------------------------------------------------
# Tensor-core direct FIR (csrc/tc_toeplitz.cu): byte-level emulation of the operand addressing
# ------------------------------------------------------------------------------------------------
class TcEmu:
    """Mirrors tc_toeplitz_kernel: the band of 4-sample windows (A operand, Hankel via SBO = 128 B / LBO = 64 B),
    the tap images (B operand, rows 16 B apart, planes LBO apart, row block a read A-1-a rows down), the K-major
    no-swizzle core-matrix address rule of the tcgen05 shared-memory descriptor, TMEM accumulation over
    (a, K-step), the pending-output ring, and the FP32-FMA walk over the first B taps that produces the buffer's
    own samples (columns e < A).  float64, no hi/lo split: this checks addressing, not rounding."""
    ROWS, KSTEPS, PLANES = 128, 16, 32

    def __init__(self, T, B, L):
        assert B % 128 == 0
        self.T, self.B, self.L = T, B, L
        self.A = B // 128
        self.C = (L + 127) // 128
        self.NE = self.C + self.A - 1
        self.COLS = min(128, max(16, (self.C - 1 + 15) // 16 * 16))      # MMA N: the tensor core does columns A .. NE-1
        self.NGRP = max(1, (self.C - 1 + self.COLS - 1) // self.COLS)
        self.R = self.COLS + self.A - 1
        self.capP = (128 * self.NE + B - 1) // B * B
        self.pend = np.zeros((T, self.capP))
        self.xprev = np.zeros((T, 128))
        self.ppos = 0
        self.images = None

    def load_ir(self, h):
        # image[grp][S][row][j] = h[128 c + 127 - (4 S + j)], tap column c = N grp + row + 1, zero outside [0, L)
        img = np.zeros((self.T, self.NGRP, self.PLANES, self.R, 4))
        for grp in range(self.NGRP):
            for S in range(self.PLANES):
                for row in range(self.R):
                    c = grp * self.COLS + row + 1
                    for j in range(4):
                        k = 128 * c + 127 - (4 * S + j)
                        if 0 <= c < self.C and 0 <= k < self.L:
                            img[:, grp, S, row, j] = h[:, k]
        self.images = img
        self.hhead = np.zeros((self.T, self.B))
        n = min(self.B, self.L)
        self.hhead[:, :n] = h[:, :n]

    @staticmethod
    def _operand(flat, start, rows, lbo, sbo):
        """rows x 8 operand tile of one MMA read through a K-major no-swizzle descriptor (bytes -> float index)."""
        out = np.empty((rows, 8))
        for r in range(rows):
            for kk in range(8):
                addr = start + (r // 8) * sbo + (r % 8) * 16 + (kk // 4) * lbo + (kk % 4) * 4
                assert addr % 4 == 0 and 0 <= addr // 4 < flat.size, (addr, flat.size)
                out[r, kk] = flat[addr // 4]
        return out

    def process(self, x, commit=True):
        T, B, A = self.T, self.B, self.A
        y = np.full((T, B), np.nan)
        new_pend = self.pend.copy()
        plane_bytes = self.R * 16
        for t in range(T):
            xw = np.concatenate([self.xprev[t], x[t]])          # xw[i] = x[i - 128]
            band = np.empty((B + 124, 4))           # band[g] = x[g-127 .. g-124]; the last window ends at x[B-1]
            for g in range(B + 124):
                band[g] = xw[g + 1:g + 5]
            band = band.ravel()
            # own samples: lane l owns outputs 128 e + 4 l .. + 3 of every row block e, warp w the taps 32 w .. 32 w + 31
            # of every tap column c <= e; window float4 index into xw: 32 + qd - kq, qd = 32 e + l, kq = 32 c + 8 w + i
            hh = self.hhead[t].reshape(-1, 4)
            xw4 = xw.reshape(-1, 4)
            part = np.zeros((4, A, 32, 4))
            for w in range(4):
                for c in range(A):
                    for e in range(c, A):
                        for l in range(32):
                            base = 32 + 32 * (e - c) + l - 8 * w
                            assert base - 8 >= 0 and base < xw4.shape[0]
                            xb = xw4[base]
                            for i in range(8):
                                xa = xw4[base - (i + 1)]
                                win = np.concatenate([xa, xb])    # x[m-4 .. m+3], m = n0 - 4 kq
                                for k in range(4):
                                    part[w, e, l] += hh[32 * c + 8 * w + i, k] * win[4 - k:8 - k]
                                xb = xa
            for e in range(A):
                idx = self.ppos + 128 * e
                if idx >= self.capP:
                    idx -= self.capP
                v = ((part[0, e] + part[1, e]) + part[2, e]) + part[3, e]
                y[t, 128 * e:128 * e + 128] = v.reshape(-1) + self.pend[t, idx:idx + 128]
                new_pend[t, idx:idx + 128] = 0.0
            for grp in range(self.NGRP):
                img = self.images[t, grp].ravel()
                D = np.zeros((self.ROWS, self.COLS))
                for a in range(A):
                    boff = 16 * (A - 1 - a)
                    for q in range(self.KSTEPS):
                        Aop = self._operand(band, 2048 * a + 128 * q, self.ROWS, 64, 128)
                        Bop = self._operand(img, 2 * q * plane_bytes + boff, self.COLS, plane_bytes, 128)
                        D += Aop @ Bop.T
                e0 = A + grp * self.COLS
                idx0 = self.ppos + 128 * e0
                if idx0 >= self.capP:
                    idx0 -= self.capP
                nwrap, ncol = (self.capP - idx0) >> 7, min(self.COLS, self.NE - e0)
                for k in range(self.COLS):
                    if k >= ncol:
                        assert not D[:, k].any(), "a column beyond NE received a contribution"
                        continue
                    idx = idx0 + 128 * k - (0 if k < nwrap else self.capP)   # the kernel's wrap-once rule
                    assert 0 <= idx <= self.capP - 128 and idx == (self.ppos + 128 * (e0 + k)) % self.capP
                    new_pend[t, idx:idx + 128] = D[:, k] + self.pend[t, idx:idx + 128]
            if commit:
                self.xprev[t] = xw[B:B + 128]
        if commit:
            self.pend = new_pend
            self.ppos = (self.ppos + B) % self.capP
        assert not np.isnan(y).any()
        return y


def bus_slice_width(T, Bs):
    """csrc/engine.cu plan_tc: columns per CTA of the column-slice bus (0 = ticket tree)."""
    sl = 4
    while sl * T < Bs:
        sl *= 2
    return sl if sl <= 512 else 0


def bus_slice_emulate(y, gains, B, n_off, Bs, sl):
    """Mirrors bus_slice_reduce (csrc/bus_tree.cuh) for one launch: CTA t sums the columns n_off + t*sl .. + sl of the
    rows y[T][B] over all tracks with 128 threads = (row lanes) x (column quads), lane partials through shared memory
    in two steps.  Returns mix[2][B] with NaN where the launch writes nothing, and asserts every slot is written once."""
    T = y.shape[0]
    nthr = 128
    mix = np.full(2 * B, np.nan)
    Q, RL = sl // 4, nthr // (sl // 4)
    assert Q * RL == nthr and sl * T >= Bs
    for t in range(T):
        c_lo = t * sl
        if c_lo >= Bs:
            continue
        part = np.zeros(nthr * 8)
        for tid in range(nthr):
            rl, q = tid // Q, tid % Q
            col = n_off + c_lo + 4 * q
            L, R = np.zeros(4), np.zeros(4)
            if c_lo + 4 * q < Bs:
                assert col + 4 <= B
                for t0 in range(rl, T, 4 * RL):
                    for j in range(4):
                        row = t0 + j * RL
                        if row < T:
                            L += gains[row, 0] * y[row, col:col + 4]
                            R += gains[row, 1] * y[row, col:col + 4]
            part[(rl * Q + q) * 8:(rl * Q + q) * 8 + 4] = L
            part[(rl * Q + q) * 8 + 4:(rl * Q + q) * 8 + 8] = R
        NO = 8 * Q
        nseg = nthr // NO if NO < nthr else 1
        per = RL // nseg
        assert per * nseg == RL

        def emit(o, total):
            fq, k = o >> 3, o & 7
            if c_lo + 4 * fq >= Bs:
                return
            i = (0 if k < 4 else B) + n_off + c_lo + 4 * fq + (k & 3)
            assert np.isnan(mix[i]), "a bus value written twice"
            mix[i] = total

        if nseg > 1:
            part2 = np.zeros(nseg * NO)
            for tid in range(nthr):
                o, sg = tid % NO, tid // NO
                part2[sg * NO + o] = sum(part[(r * Q + (o >> 3)) * 8 + (o & 7)] for r in range(sg * per, (sg + 1) * per))
            for tid in range(NO):
                emit(tid, sum(part2[g2 * NO + tid] for g2 in range(nseg)))
        else:
            for o in range(NO):
                emit(o, sum(part[(r * Q + (o >> 3)) * 8 + (o & 7)] for r in range(RL)))
    return mix.reshape(2, B)
