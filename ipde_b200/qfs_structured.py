"""
Structured quadrature by fundamental solutions: O(n log n) apply, O(n W log n + n W^2) setup.

Same job as `ipde_b200.qfs.QFS` (the reference's `qfs.*_QFS` objects, call sites
ipde/solvers/internals/modified_helmholtz.py:24-27, poisson.py:23-26, stokes.py:21-24, used at
internals/scalar.py:87-88,106 and internals/vector.py:124-125,150): map single/double layer densities on a closed
curve to point sources on a nearby curve whose field reproduces the layer potential on one side of the curve, and
`u2s`: boundary values -> sources.  The dense version factors the check<-source matrix by an SVD: O(n^3) setup and
O(n^2) bytes per apply, which rules out the boundaries of the 8192^2 / 16384^2 configurations (n_b = 23k .. 65k).

Structure used here.  The curve Z(t) is a trigonometric polynomial (its Fourier series cut at round-off), so it can
be evaluated at complex parameters.  Sources and check points are placed at *complex parameter shifts*
    source_j = Z(s_j - i*sgn*alpha_s*dt),    check_i = Z(t_i + i*sgn*alpha_c*dt),      dt = 2 pi / N
(to first order: normal offsets by alpha * local node spacing).  For the Laplace kernel
log|Z(a) - Z(b)| = log|e^{ia} - e^{ib}| + smooth(a, b): check<-source is a circulant plus a smooth kernel, i.e. in
the Fourier bases of the two curves it is DIAGONAL plus a block coupling only the lowest modes; the other kernels
(K0, stokeslet, double layers, the arc-length weight) add a band whose width is set by the geometry, not by n.
Every operator `out_i = sum_j K(T(t_i), S(s_j)) in_j` between two such curves has the Fourier-space form
    out_hat[k] = sum_p B[p, k] in_hat[k - p],      B[p, k] = DFT_{v -> k} ( DFT_{s -> p} K(T(s + v), S(s)) )
and the p-dependence decays at the rate fixed by the analyticity of the geometry: B is obtained from kernel samples
on J ~ 2W+1 coarse positions s times all offsets v (O(J n) kernel evaluations), never from an n x n matrix.
The density -> source map  Q = E^-1 C  is then computed column block by column block from windows of the banded
system (its inverse decays away from the diagonal as well) and stored as
    a dense block for the lowest modes (|k|, |k'| <= K_L)  +  a band of half-width B_H for all modes.
Per solve: FFT of the densities (length N), banded complex mat-vec, inverse FFT (length M = rho N) -- done by
libipde_b200 (`ipde_banded_apply_z`, csrc/dense.cu) on the device; the NumPy `__call__` below states the same
arithmetic for the oracle.

Setup uses torch tensor ops (on the GPU when there is one): setup-time plumbing, not on the per-solve path.
"""
import numpy as np
import torch

TWO_PI = 2.0*np.pi


class NotResolved(Exception):
    """the geometry or the band structure is not resolved at this N: use the dense QFS"""


def _k0(x):
    return torch.special.modified_bessel_k0(x)


def _k1(x):
    return torch.special.modified_bessel_k1(x)


class _Curve(object):
    """band-limited closed curve Z(t) = sum_k c_k exp(i k t), evaluated at complex parameters"""
    def __init__(self, c, dev, cut=1e-14):
        N = c.size
        ch = np.fft.fft(np.asarray(c, dtype=complex))/N
        k = np.fft.fftfreq(N, 1.0/N)
        big = np.abs(ch) > cut*np.abs(ch).max()
        self.K = int(np.abs(k[big]).max())
        sel = np.abs(k) <= self.K
        if N % 2 == 0 and self.K >= N//2:
            # Nyquist mode present: keep the samples as they are (split it evenly)
            sel = np.ones(N, dtype=bool)
            ch = ch.copy()
            kk = list(k[sel]) + [N//2]
            cc = list(ch)
            cc[N//2] = 0.5*ch[N//2]
            cc.append(0.5*ch[N//2])
            self.k_np, self.ch_np = np.array(kk), np.array(cc)
        else:
            self.k_np, self.ch_np = k[sel].copy(), ch[sel].copy()
        self.N = N
        self.dev = dev
        self.k = torch.from_numpy(self.k_np).to(dev)
        self.ch = torch.from_numpy(self.ch_np).to(dev)
        self.ki = torch.from_numpy(self.k_np.astype(np.int64)).to(dev)

    def grid(self, n, shifts=None, delta=0.0, deriv=0):
        """Z^(deriv)(2 pi j/n + shift_m + i delta) -> (len(shifts), n) complex (or (n,) when shifts is None)"""
        single = shifts is None
        if single:
            shifts = torch.zeros(1, dtype=torch.float64, device=self.dev)
        a = self.ch*torch.exp(-self.k*delta)*(1j*self.k)**deriv
        ph = torch.exp(1j*shifts[:, None]*self.k[None, :])*a[None, :]
        buf = torch.zeros((shifts.numel(), n), dtype=torch.complex128, device=self.dev)
        buf.index_add_(1, self.ki % n, ph)
        out = torch.fft.ifft(buf, dim=1)*n
        return out[0] if single else out

    def at(self, s, delta=0.0, deriv=0):
        a = self.ch*torch.exp(-self.k*delta)*(1j*self.k)**deriv
        return torch.exp(1j*s[:, None]*self.k[None, :]) @ a


# ---------------------------------------------------------------------------------------------------------------
# kernels on complex differences d = target - source; each returns a list of (out_comp, in_comp, values)
class _Scalar(object):
    ncomp = 1
    def __init__(self, kind, k=0.0):
        self.kind, self.k = kind, float(k)          # kind 0 Laplace, 1 modified Helmholtz

    def G(self, d):
        r = d.abs()
        if self.kind == 0:
            return [(0, 0, -torch.log(r)/TWO_PI)]
        return [(0, 0, _k0(self.k*r)/TWO_PI)]

    def dG(self, d, nrm):
        """dG/dn_source * |weight carried by nrm|;  nrm complex (normal times weight)"""
        r = d.abs()
        dn = d.real*nrm.real + d.imag*nrm.imag
        if self.kind == 0:
            return [(0, 0, dn/(r*r)/TWO_PI)]
        return [(0, 0, self.k*_k1(self.k*r)*dn/r/TWO_PI)]


class _Stokes(object):
    ncomp = 2
    kind, k = 2, 0.0

    def G(self, d):
        dx, dy = d.real, d.imag
        r2 = dx*dx + dy*dy
        L = -0.5*torch.log(r2)
        c = 1.0/(4.0*np.pi)
        xy = c*dx*dy/r2
        return [(0, 0, c*(L + dx*dx/r2)), (0, 1, xy), (1, 0, xy), (1, 1, c*(L + dy*dy/r2))]

    def P(self, d):
        """pressure of a point force: (1/2pi) d.f/r^2 (ipde/solvers/internals/stokes_save.py:29-38)"""
        dx, dy = d.real, d.imag
        r2 = dx*dx + dy*dy
        return [(2, 0, dx/r2/TWO_PI), (2, 1, dy/r2/TWO_PI)]

    def dG(self, d, nrm):
        dx, dy = d.real, d.imag
        r2 = dx*dx + dy*dy
        f = (dx*nrm.real + dy*nrm.imag)/(r2*r2)/np.pi
        xy = f*dx*dy
        return [(0, 0, f*dx*dx), (0, 1, xy), (1, 0, xy), (1, 1, f*dy*dy)]


# ---------------------------------------------------------------------------------------------------------------
class _Band(object):
    """B[o, i, p, k]: coefficient of in_hat_i[k - p] in out_hat_o[k];  p in [-W, W] (index p + W), k in [-Ks, Ks]"""
    def __init__(self, data, W, Ks):
        self.data, self.W, self.Ks = data, W, Ks


def _operator_band(ker, layer, tc, t_delta, n_t, sc, s_delta, n_d, J, Ks, chunk_elems=3.0e7):
    """
    Fourier-space band of  out_i = sum_j K(T(t_i), S(s_j)) in_j,  T on the n_t grid, S on the n_d grid (n_t | n_d).
    layer 'c': plain kernel (charges);  's': single layer density (times |S'| ds);  'd': double layer density.
    Returns (band (nc, nc, J, 2Ks+1) with p in fft order, tail): tail = size of the band at |p| ~ J/2 relative to
    its largest entry (aliasing indicator).
    """
    dev = tc.dev
    nc = ker.ncomp
    n_o = nc + 1 if layer == 'cp' else nc               # 'cp': Stokes velocity + pressure of point forces
    kr = torch.arange(-Ks, Ks + 1, device=dev)
    s = TWO_PI*torch.arange(J, dtype=torch.float64, device=dev)/J
    S = sc.at(s, delta=s_delta)
    Sp = sc.at(s, delta=s_delta, deriv=1)
    g1 = torch.zeros((n_o, nc, J, kr.numel()), dtype=torch.complex128, device=dev)
    rows = max(1, int(chunk_elems//n_d))
    w = TWO_PI/n_d
    for m0 in range(0, J, rows):
        m1 = min(J, m0 + rows)
        T = tc.grid(n_d, shifts=s[m0:m1], delta=t_delta)              # T(s_m + v_d)
        d = T - S[m0:m1, None]
        if layer == 'c':
            terms = ker.G(d)
        elif layer == 'cp':
            terms = ker.G(d) + ker.P(d)
        elif layer == 's':
            terms = [(o, i, v*(Sp[m0:m1, None].abs()*w)) for o, i, v in ker.G(d)]
        else:
            terms = ker.dG(d, (-1j*w)*Sp[m0:m1, None])                # |S'| n_S ds
        for o, i, v in terms:
            f = torch.fft.fft(v.to(torch.complex128), dim=1)          # sum_d . exp(-i k v_d)
            g1[o, i, m0:m1] = f[:, kr % n_d]
        del T, d, terms
    band = torch.fft.fft(g1, dim=2)/J                                 # p in fft order
    mag = band.abs().amax(dim=(0, 1, 3))
    q = J//2
    tail = float(mag[max(1, q - max(2, J//16)):q + max(2, J//16) + 1].max()/mag.max())
    return band, tail


def _dense_window(band, J, rows, cols):
    """dense (nc*nr, nc*ncol) matrix of a band (fft p-order, k index = k + Ks) for mode lists rows / cols"""
    nc = band.shape[0]
    Ks = (band.shape[3] - 1)//2
    p = rows[:, None] - cols[None, :]
    ok = p.abs() < J//2
    pi = torch.where(ok, p % J, torch.zeros_like(p))
    ri = (rows + Ks)[:, None].expand_as(p)
    out = torch.zeros((nc, rows.numel(), nc, cols.numel()), dtype=torch.complex128, device=band.device)
    for o in range(nc):
        for i in range(nc):
            out[o, :, i, :] = torch.where(ok, band[o, i][pi, ri], torch.zeros((), dtype=torch.complex128,
                                                                                device=band.device))
    return out.reshape(nc*rows.numel(), nc*cols.numel())


class StructuredOperator(object):
    """
    density (n_in components on N nodes) -> charges (n_out components on M nodes):
        out_hat = Low @ in_hat[low modes]  (+)  band
    Host copies (NumPy) of everything the device apply needs:
      low:   (n_out, n_in, 2KL+1, 2KL+1) complex, rows = output mode k, columns = input mode k'
      band:  (n_out, n_in, 2BH+1, 2Ks+1) complex, band[o,i,p+BH,k+Ks] multiplies in_hat_i[k-p] (zero inside the
             low block and where k-p is not an input mode)
    """
    def __init__(self, N, M, n_in, n_out, KL, BH, Ks, low, band):
        self.N, self.M, self.n_in, self.n_out = N, M, n_in, n_out
        self.KL, self.BH, self.Ks = KL, BH, Ks
        self.low_t, self.band_t = low, band          # torch tensors where they were built (the device, if any)
        self._np = None

    def _host(self):
        if self._np is None:
            self._np = (self.low_t.cpu().numpy(), self.band_t.cpu().numpy())
        return self._np

    @property
    def low(self):
        return self._host()[0]

    @property
    def band(self):
        return self._host()[1]

    def in_modes(self, x):
        """x (n_in, N) real -> (n_in, N+1 or N) complex coefficients of modes -N/2..N/2 (Nyquist split)"""
        N = self.N
        xh = np.fft.fft(x, axis=1)/N
        h = N//2
        out = np.zeros((x.shape[0], 2*h + 1), dtype=complex)
        k = np.fft.fftfreq(N, 1.0/N).astype(int)
        if N % 2 == 0:
            out[:, k[k != -h] + h] = xh[:, k != -h]
            out[:, 0] = 0.5*xh[:, h]
            out[:, 2*h] = 0.5*xh[:, h]
        else:
            out[:, k + h] = xh
        return out

    def apply(self, x):
        """NumPy statement of the device apply; x (n_in, N) -> (n_out, M)"""
        N, M, KL, BH, Ks = self.N, self.M, self.KL, self.BH, self.Ks
        h = N//2
        xin = self.in_modes(np.asarray(x, dtype=float).reshape(self.n_in, N))
        # input modes on the output mode axis (zero outside |k'| <= h)
        xe = np.zeros((self.n_in, 2*Ks + 1 + 2*BH), dtype=complex)          # index k' + Ks + BH
        xe[:, Ks + BH - h:Ks + BH + h + 1] = xin
        oh = np.zeros((self.n_out, 2*Ks + 1), dtype=complex)
        for o in range(self.n_out):
            for i in range(self.n_in):
                for p in range(-BH, BH + 1):
                    # in_hat[k - p] for all k
                    oh[o] += self.band[o, i, p + BH]*xe[i, BH - p:BH - p + 2*Ks + 1]
                if KL >= 0:
                    kl = min(KL, h)
                    oh[o, Ks - KL:Ks + KL + 1] += self.low[o, i][:, KL - kl:KL + kl + 1].dot(xin[i, h - kl:h + kl + 1])
        buf = np.zeros((self.n_out, M), dtype=complex)
        kk = np.arange(-Ks, Ks + 1)
        buf[:, kk % M] = oh
        return (np.fft.ifft(buf, axis=1)*M).real


def _batched_solve(A, B):
    """batched LU solve; on the GPU through cuSOLVER (MAGMA's batched path prints a size warning to stdout for
    windows of this size, which would end up in front of bench.py's JSON line)"""
    if A.is_cuda:
        prev = torch.backends.cuda.preferred_linalg_library()
        torch.backends.cuda.preferred_linalg_library('cusolver')
        try:
            return torch.linalg.solve(A, B)
        finally:
            torch.backends.cuda.preferred_linalg_library(prev)
    return torch.linalg.solve(A, B)


def _solve_deflated(Es, R, null_tol=1e-9):
    """
    Solve Es Y = R when Es may have ONE null direction (Stokes: a uniform normal force on a closed curve generates no
    velocity; the window holding the lowest modes sees it).  The null pair (u, v) comes from two steps of inverse
    iteration on the LU factors; Es + s1 u v^H is then well conditioned and its LU solve equals the pseudo-inverse
    solution up to a multiple of the field-free vector v.  An SVD-based pseudo-inverse would smear cond * eps of
    round-off over every row of the window and defeat the band truncation.
    """
    n = Es.shape[0]
    LU, piv = torch.linalg.lu_factor(Es)
    g = torch.Generator(device='cpu').manual_seed(7)
    def null_vec(adjoint):
        x = torch.randn(n, 1, dtype=torch.float64, generator=g).to(Es.device).to(Es.dtype)
        for _ in range(3):
            x = torch.linalg.lu_solve(LU, piv, x, adjoint=adjoint)
            x = x/torch.linalg.vector_norm(x)
        return x
    v = null_vec(False)
    s1 = float(Es.abs().sum(dim=1).max())
    if float(torch.linalg.vector_norm(Es @ v)) > null_tol*s1:
        return torch.linalg.lu_solve(LU, piv, R)
    u = null_vec(True)
    Y = torch.linalg.solve(Es + s1*(u @ v.conj().T), R)
    # remove the field-free component: the minimum-norm (pseudo-inverse) solution, localised like the rest of E^-1
    return Y - v @ (v.conj().T @ Y)


def _solve_windows(Eb, Cb_list, J, Ks, h_in, col_block, Bw, scale_r, scale_c, pinv_low=False, svd_cut=1e-13):
    """
    X = E^-1 [C_1 | C_2 ...] column block by column block from windows of the banded system.
    Eb: band of E (nc, nc, J, 2Ks+1); Cb_list: bands of the right-hand sides (same layout) or None for the identity.
    Input modes k' in [-h_in, h_in].  Returns (Bx, [X_l]) with X_l (nc, nc, 2Bx+1, 2h_in+1) stored by column:
    X_l[o, i, p + Bx, k' + h_in] = entry (output mode k' + p, input mode k'); zero outside the column's window.
    """
    dev = Eb.device
    nc = Eb.shape[0]
    ncol = 2*h_in + 1
    Bx = Bw + col_block
    outs = [torch.zeros((nc, nc, 2*Bx + 1, ncol), dtype=torch.complex128, device=dev) for _ in Cb_list]
    blocks = []
    a = -h_in
    while a <= h_in:
        b = min(h_in + 1, a + col_block)
        blocks.append((a, b, max(-Ks, a - Bw), min(Ks, b - 1 + Bw)))
        a = b

    def assemble(a, b, lo, hi):
        rows = torch.arange(lo, hi + 1, device=dev)
        cols = torch.arange(a, b, device=dev)
        nr, ncb = rows.numel(), cols.numel()
        sr = scale_r(rows).repeat(nc)
        sc_ = scale_c(rows).repeat(nc)
        Es = _dense_window(Eb, J, rows, rows)*sr[:, None]*sc_[None, :]
        rhs = []
        for Cb in Cb_list:
            if Cb is None:
                C = torch.zeros((nc, nr, nc, ncb), dtype=torch.complex128, device=dev)
                idx = (cols - lo)
                for c in range(nc):
                    C[c, idx, c, torch.arange(ncb, device=dev)] = 1.0
                C = C.reshape(nc*nr, nc*ncb)
            else:
                C = _dense_window(Cb, J, rows, cols)
            rhs.append(C*sr[:, None])
        return Es, torch.cat(rhs, dim=1), sc_, rows, cols

    def scatter(Y, sc_, rows, cols, a, b, lo, hi):
        nr, ncb = rows.numel(), cols.numel()
        Y = Y*sc_[:, None]
        w = nc*ncb
        pidx = rows[:, None] - cols[None, :] + Bx                       # (nr, ncb), in [0, 2Bx]
        cidx = (cols + h_in)[None, :].expand(nr, ncb)
        for l, out in enumerate(outs):
            Yl = Y[:, l*w:(l + 1)*w].reshape(nc, nr, nc, ncb)
            for o in range(nc):
                for i in range(nc):
                    out[o, i][pidx, cidx] = Yl[o, :, i, :]

    # windows of equal shape are solved as one batched LU (most of them: every interior block); the window(s) holding
    # mode 0 of a system with a field-free direction go through the deflated solve one at a time
    groups = {}
    for blk in blocks:
        a, b, lo, hi = blk
        if pinv_low and lo <= 0 <= hi:
            Es, R, sc_, rows, cols = assemble(*blk)
            scatter(_solve_deflated(Es, R), sc_, rows, cols, *blk)
            continue
        groups.setdefault((hi - lo, b - a), []).append(blk)
    for (_, _), blks in groups.items():
        nr = nc*(blks[0][3] - blks[0][2] + 1)
        per = max(1, int(2.0e9//(16*nr*nr)))                            # <= 2 GB of window matrices per batch
        for g0 in range(0, len(blks), per):
            part = blks[g0:g0 + per]
            items = [assemble(*blk) for blk in part]
            Y = _batched_solve(torch.stack([it[0] for it in items]), torch.stack([it[1] for it in items]))
            for yb, it, blk in zip(Y, items, part):
                scatter(yb, it[2], it[3], it[4], *blk)
            del items, Y
    return Bx, outs


def _significant(X, Bx, Ks, att, tol, far):
    """
    X by-column solution (nc, nc, 2Bx+1, ncol); att[k + Ks]: attenuation of source mode k on the way back to the
    curve.  An entry matters if its attenuated size exceeds tol * (largest attenuated entry of its column) and stands
    out of the column's round-off floor.  The floor is measured |k - k'| ~ far away from the diagonal, at the edge
    of the operator bands, where the exact solution has decayed to nothing: the exponentially small high-mode
    entries of the banded operators come out of an FFT with absolute, not relative, accuracy.
    """
    dev = X.device
    ncol = X.shape[3]
    h_in = (ncol - 1)//2
    p = torch.arange(-Bx, Bx + 1, device=dev)
    kp = torch.arange(-h_in, h_in + 1, device=dev)
    k = p[:, None] + kp[None, :]
    ok = k.abs() <= Ks
    a = torch.where(ok, att[(k.clamp(-Ks, Ks) + Ks)], torch.zeros((), dtype=att.dtype, device=dev))
    mag = X.abs().amax(dim=(0, 1))*a                                     # (2Bx+1, ncol)
    ring = (p.abs() >= far) & (p.abs() <= (4*far)//3)
    floor = mag[ring].amax(dim=0, keepdim=True)
    # ... and is not negligible against the largest gain of the whole operator (a column whose exact content is zero,
    # e.g. the double layer of a constant density on the exterior side, is round-off from top to bottom)
    return (mag > torch.maximum(tol*mag.amax(dim=0, keepdim=True), 4.0*floor)) & (mag > tol*mag.max())


def _choose(sig, Bx, Ks, h_in, KL_candidates, bh_cap=None):
    """two-tier storage (dense low block |k|,|k'| <= KL, band half-width BH elsewhere) of least size"""
    dev = sig.device
    p = torch.arange(-Bx, Bx + 1, device=dev)
    kp = torch.arange(-h_in, h_in + 1, device=dev)
    k = p[:, None] + kp[None, :]
    pa = p.abs()[:, None].expand_as(k)
    best = None
    for KL in KL_candidates:
        KL = min(KL, h_in, Bx)
        inlow = (k.abs() <= KL) & (kp.abs()[None, :] <= KL)
        rest = sig & ~inlow
        BH = int(pa[rest].max().item()) if bool(rest.any()) else 0
        if bh_cap is not None:
            BH = min(BH, bh_cap)
        cost = (2*KL + 1)**2 + (2*Ks + 1)*(2*BH + 1)
        if best is None or cost < best[0]:
            best = (cost, KL, BH)
        if KL == min(h_in, Bx):
            break
    return best[1], best[2]


def _fill(X, Bx, Ks, h_in, KL, BH):
    """by-column solution -> (low (nc, nc, 2KL+1 [k], 2KL+1 [k']), band (nc, nc, 2BH+1 [p], 2Ks+1 [k]))"""
    dev = X.device
    nc = X.shape[0]
    kl = torch.arange(-KL, KL + 1, device=dev)
    pl = kl[:, None] - kl[None, :]                                       # (k, k')
    inb = pl.abs() <= Bx
    cl = (kl + h_in)[None, :].expand_as(pl)
    low = (X[:, :, pl.clamp(-Bx, Bx) + Bx, cl]*inb[None, None]).contiguous()
    k = torch.arange(-Ks, Ks + 1, device=dev)
    pb = torch.arange(-BH, BH + 1, device=dev)
    src_k = k[None, :] - pb[:, None]                                     # input mode, (2BH+1, nk)
    ok = (src_k.abs() <= h_in) & ~((k.abs()[None, :] <= KL) & (src_k.abs() <= KL))
    pi = (pb + Bx)[:, None].expand_as(src_k)
    ci = src_k.clamp(-h_in, h_in) + h_in
    band = X[:, :, pi, ci]*ok[None, None]
    # The operators map real densities to real strengths: Q(-k, -k') = conj Q(k, k').  Enforce it exactly -- the device
    # apply packs two real outputs into one complex spectrum and relies on it, and the deflated Stokes solve adds a
    # complex multiple of the (field-free) null vector, whose imaginary part must not reach the other component.
    low = 0.5*(low + low.flip(2, 3).conj())
    band = 0.5*(band + band.flip(2, 3).conj())
    return low.contiguous(), band.contiguous()


class CurveOperator(object):
    """
    charges on the source curve (n_in components, M sources) -> field on target curves (n_out rows of N nodes):
    same two-tier Fourier-space storage as StructuredOperator; the field spectrum (modes |k| <= Ks) is transformed
    back on the M-point grid and every rho-th sample is a target node.
    """
    def __init__(self, N, M, n_in, n_out, KL, BH, Ks, low, band):
        self.N, self.M, self.n_in, self.n_out = N, M, n_in, n_out
        self.KL, self.BH, self.Ks = KL, BH, Ks
        self.low_t, self.band_t = low, band
        self.band_half_t = None                      # (n_out, n_in, 2BH+1, Ks+1): output modes k >= 0 only
        self._np = None

    def _host(self):
        if self._np is None:
            if self.band_t is None:
                # full band from the stored half: B[p, -k] = conj B[-p, k]
                half = self.band_half_t.cpu().numpy()
                neg = np.conj(half[:, :, ::-1, :0:-1])
                band = np.concatenate([neg, half], axis=3)
            else:
                band = self.band_t.cpu().numpy()
            self._np = (self.low_t.cpu().numpy(), band)
        return self._np

    @property
    def low(self):
        return self._host()[0]

    @property
    def band(self):
        return self._host()[1]

    def apply(self, q):
        """NumPy statement of the device apply; q (n_in, M) charges -> (n_out, N) field values"""
        M, KL, BH, Ks = self.M, self.KL, self.BH, self.Ks
        qh = np.fft.fft(np.asarray(q, dtype=float).reshape(self.n_in, M), axis=1)/M
        kk = np.arange(-Ks - BH, Ks + BH + 1)
        ok = np.abs(kk) < M//2
        xe = np.where(ok[None, :], qh[:, kk % M], 0.0)                       # index k' + Ks + BH
        oh = np.zeros((self.n_out, 2*Ks + 1), dtype=complex)
        for o in range(self.n_out):
            for i in range(self.n_in):
                for p in range(-BH, BH + 1):
                    oh[o] += self.band[o, i, p + BH]*xe[i, BH - p:BH - p + 2*Ks + 1]
                oh[o, Ks - KL:Ks + KL + 1] += self.low[o, i].dot(xe[i, Ks + BH - KL:Ks + BH + KL + 1])
        buf = np.zeros((self.n_out, M), dtype=complex)
        buf[:, np.arange(-Ks, Ks + 1) % M] = oh
        return (np.fft.ifft(buf, axis=1)*M).real[:, ::M//self.N]


def source_to_curves(qfs, curves, pressure=False, tol=2e-15, KLc=(32, 64, 128, 256, 512)):
    """
    Field of the point sources of a StructuredQFS object on curves given by N samples each (same parametrisation as
    the QFS curve: its own nodes, the radial levels X(t_j) + r_i n(t_j) of the annulus): the operator of
    ScalarHelper.correct / VectorHelper.correct (ipde/solvers/internals/scalar.py:101,113, vector.py:141,156) as a
    Fourier-banded map, O(n) bytes per level instead of an all-pairs sum or a free-space evaluator per boundary.
    Entries below tol * (largest entry) are at the round-off floor of the sampled bands and are dropped.
    """
    cv, dev = qfs._cv, qfs.dev
    N, M, J, Ks_src = qfs.N, qfs.M, qfs.J, qfs.Ks
    ker = qfs.ker
    nc = ker.ncomp
    Ks = min(Ks_src + J//2, M//2 - 1)
    layer = 'cp' if (pressure and nc == 2) else 'c'
    Bx = J//2 - 1
    k_src = torch.arange(-Ks_src, Ks_src + 1, device=dev)
    p = torch.arange(-Bx, Bx + 1, device=dev)
    kk = p[:, None] + k_src[None, :]                                     # output mode of entry (p, k')
    okk = kk.abs() <= Ks
    def columns(c):
        tc = _Curve(np.asarray(c), dev, cut=1e-15)
        band, _ = _operator_band(ker, layer, tc, 0.0, N, cv, qfs.d_src, M, J, Ks)
        # by-column layout X[o, i, p + Bx, k' + Ks_src] = band[o, i, p, k' + p]
        return band[:, :, (p % J)[:, None].expand_as(kk), kk.clamp(-Ks, Ks) + Ks]*okk[None, None]
    # pass 1: which entries matter (relative to the largest entry over all curves); pass 2: fill the two-tier storage
    mags = [columns(c).abs().amax(dim=(0, 1)) for c in curves]
    top = max(float(m.max()) for m in mags)
    sig = mags[0] > tol*top
    for m in mags[1:]:
        sig |= m > tol*top
    del mags
    KL, BH = _choose(sig, Bx, Ks, Ks_src, list(KLc))
    lows, bands = [], []
    for c in curves:
        lo_, ba_ = _fill(columns(c), Bx, Ks, Ks_src, KL, BH)
        lows.append(lo_)
        bands.append(ba_[:, :, :, Ks:].clone())                         # output modes k >= 0 (what the device reads)
    low = torch.cat(lows, dim=0)
    band_half = torch.cat(bands, dim=0)
    op = CurveOperator(N, M, nc, low.shape[0], KL, BH, Ks, low, None)
    op.band_half_t = band_half
    return op


class _Source(object):
    """point sources: x, y, N, weights (all ones: the strengths returned by the operators are charges)"""
    def __init__(self, c):
        self.c = np.asarray(c)
        self.x = np.ascontiguousarray(self.c.real)
        self.y = np.ascontiguousarray(self.c.imag)
        self.N = self.x.size
        self.weights = np.ones(self.N)
        self.normal_x = np.zeros(self.N)
        self.normal_y = np.zeros(self.N)

    def get_stacked_boundary(self, T=True):
        return np.vstack([self.x, self.y])


class StructuredQFS(object):
    """
    Same constructor meaning as ipde_b200.qfs.QFS.  Attributes: N, M, source, op (densities -> charges), op_u2s.
    `__call__([slp, dlp])` and `u2s(ub)` return charges per source (source.weights == 1).
    """
    structured = True
    on_device = False

    def __init__(self, bdy, interior, slp, dlp, kernel, tol=1e-12, source_upsample_factor=1.0, alpha_s=None,
                 alpha_c=None, rho=None, rho_f=None, u2s=True, b2c=True, operators=True, device=None, J=None):
        from .qfs import default_parameters
        self.bdy, self.interior = bdy, interior
        self.has_slp, self.has_dlp = slp, dlp
        self.kernel = kernel
        self.tol = tol
        a_s, a_c, r, r_f = default_parameters(tol)
        self.alpha_s = a_s if alpha_s is None else alpha_s
        self.alpha_c = a_c if alpha_c is None else alpha_c
        self.rho = int(np.ceil((r if rho is None else rho)*source_upsample_factor))
        self.rho_f = r_f if rho_f is None else rho_f
        self.rho_f = int(np.ceil(self.rho_f/self.rho))*self.rho          # fine grid = multiple of the source grid
        self.want_u2s, self.want_b2c, self.want_operators = u2s, b2c, operators
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.dev = torch.device(device)
        self.ker = _Stokes() if kernel.name == 'stokes' else _Scalar(0 if kernel.name == 'laplace' else 1, kernel.k)
        self._build(J)

    def _build(self, J):
        bdy = self.bdy
        N = bdy.N
        dev = self.dev
        dt = TWO_PI/N
        if bdy.area < 0:
            raise NotResolved('curve must be counter-clockwise')
        cv = _Curve(bdy.c, dev)
        if 4*cv.K > N:
            raise NotResolved('geometry bandwidth %d too close to N/2 = %d' % (cv.K, N//2))
        sgn = 1.0 if self.interior else -1.0
        # complex parameter shift +i*d moves a CCW curve inward; evaluation side inside -> check inward, sources out
        # keep the source curve regular: |Z'(t - i d)| must stay away from zero
        while True:
            d_src, d_chk = -sgn*self.alpha_s*dt, sgn*self.alpha_c*dt
            sp = cv.grid(4*N, delta=d_src, deriv=1).abs()
            sp0 = cv.grid(4*N, deriv=1).abs()
            if float((sp/sp0).min()) > 0.55:
                break
            self.alpha_s *= 0.8
            L = np.log(1.0/max(self.tol, 1e-15))
            self.rho = int(min(8, max(self.rho, np.ceil(1.0 + L/(TWO_PI*(self.alpha_s + self.alpha_c))))))
            self.rho_f = int(np.ceil(self.rho_f/self.rho))*self.rho
            if self.alpha_s < 0.5:
                raise NotResolved('source curve cannot be kept regular')
        M = self.rho*N
        self.N, self.M = N, M
        self.d_src, self.d_chk = d_src, d_chk
        self.source = _Source(cv.grid(M, delta=d_src).cpu().numpy())
        self.check = _Source(cv.grid(M, delta=d_chk).cpu().numpy())
        self.op = self.op_u2s = None
        if not self.want_operators:
            return
        h = N//2
        nyq = 1 if N % 2 == 0 else 0
        h_in = h if nyq else (N - 1)//2
        # how many source modes beyond the density modes can be excited
        room = (M - N)//2 - 2
        att_c = lambda k: torch.exp(self.alpha_c*dt*k.abs().double())
        att_s = lambda k: torch.exp(self.alpha_s*dt*k.abs().double())
        Nf = self.rho_f*N
        Jtry = [J] if J is not None else [j for j in (128, 256, 512, 1024, 2048, 4096) if j <= max(128, 2*N)]
        for Jc in Jtry:
            W = Jc//2
            Ks = h_in + min(room, W)
            Eb, tail = _operator_band(self.ker, 'c', cv, d_chk, M, cv, d_src, M, Jc, Ks)
            if tail < 1e-15 or Jc >= 2*N:
                break
        else:
            raise NotResolved('band not resolved (tail %.1e at J = %d)' % (tail, Jc))
        if tail > 1e-13 and Jc < 2*N:
            raise NotResolved('band not resolved (tail %.1e at J = %d)' % (tail, Jc))
        self.J, self.W = Jc, W
        self._cv, self.Ks, self.h_in = cv, Ks, h_in
        KLc = [32, 64, 128, 256, 512, 1024, 2048]
        att = torch.exp(-self.alpha_s*dt*torch.arange(-Ks, Ks + 1, device=dev).abs().double())
        pinv_low = self.ker.ncomp == 2
        col_block = max(64, min(512, W))
        Bw = min(2*Ks + 1, 2*W)
        self.Bw = Bw
        if self.want_b2c:
            Cb = []
            if self.has_slp:
                Cb.append(_operator_band(self.ker, 's', cv, d_chk, M, cv, 0.0, Nf, Jc, Ks)[0])
            if self.has_dlp:
                Cb.append(_operator_band(self.ker, 'd', cv, d_chk, M, cv, 0.0, Nf, Jc, Ks)[0])
            Xs = _solve_windows(Eb, Cb, Jc, Ks, h_in, col_block, Bw, att_c, att_s, pinv_low)
            del Cb
            self.op = self._pack(Xs, Ks, h_in, att, KLc)
            del Xs
        if self.want_u2s:
            Ebb = _operator_band(self.ker, 'c', cv, 0.0, M, cv, d_src, M, Jc, Ks)[0]
            Xs = _solve_windows(Ebb, [None], Jc, Ks, h_in, col_block, Bw, lambda k: torch.ones_like(k, dtype=torch.float64),
                                att_s, pinv_low)
            self.op_u2s = self._pack(Xs, Ks, h_in, att, KLc)
            del Xs, Ebb

    def _pack(self, sol, Ks, h_in, att, KLc):
        """several by-column solutions (one per layer) -> one StructuredOperator with n_in = nlayers*ncomp"""
        Bx, Xs = sol
        nc = self.ker.ncomp
        far = (3*self.W)//4
        sig = _significant(Xs[0], Bx, Ks, att, 1e-15, far)
        for X in Xs[1:]:
            sig |= _significant(X, Bx, Ks, att, 1e-15, far)
        # the band never needs to be wider than the bands of the operators it is built from (W): what lies beyond
        # is the tail of the windowed inverse, at or below 1e-13 of the column (tests/test_gpu_kernels.py)
        import os
        cap = int(os.environ.get('IPDE_QFS_BH_CAP', self.W))
        KL, BH = _choose(sig, Bx, Ks, h_in, KLc, bh_cap=cap if cap > 0 else None)
        parts = [_fill(X, Bx, Ks, h_in, KL, BH) for X in Xs]
        low = torch.cat([p[0] for p in parts], dim=1)
        band = torch.cat([p[1] for p in parts], dim=1)
        return StructuredOperator(self.N, self.M, nc*len(Xs), nc, KL, BH, Ks, low, band)

    # -- NumPy statements (oracle side / setup-time use)
    def __call__(self, tau_list):
        nc = self.ker.ncomp
        tau = np.concatenate([np.asarray(t, dtype=float).ravel() for t in tau_list]).reshape(-1, self.N)
        return self.op.apply(tau).ravel()

    def u2s(self, ub):
        if self.op_u2s is None:
            raise Exception('this QFS object was built without the u2s operators (u2s=False)')
        return self.op_u2s.apply(np.asarray(ub, dtype=float).reshape(-1, self.N)).ravel()
