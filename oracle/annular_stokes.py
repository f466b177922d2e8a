"""
ORACLE (test infrastructure, CPU, NumPy) -- never imported by the product.

Restatement of the reference's annular Stokes solver:

  AnnularStokesSolver._construct          ipde/annular/stokes.py:107-153  (per-mode (3M-1)^2 complex inverses)
  AnnularStokesSolver._preconditioner     ipde/annular/stokes.py:200-210
  AnnularStokesSolver._apply_optim_real   ipde/annular/stokes.py:321-385  (the operator the reference hands to GMRES)
  AnnularStokesSolver.solve               ipde/annular/stokes.py:519-541
  mfft / mifft / mifftr (Nyquist-dropping theta transforms)   ipde/utilities.py:78-99
on the Nyquist-dropped geometry of ipde/annular/annular.py:52-108 (oracle.annular.ApproximateAnnularGeometry
with drop_nyquist=True).  Unknowns: (u_r, u_t) on M radial nodes, p on M-1, ns = n-1 Fourier modes each.

Pinned: tests/golden/annular_stokes.npz holds inputs and outputs of the reference's own AnnularStokesSolver
(generated in the build container by oracle/make_golden.py); tests/test_oracle.py checks this file against them.
"""
import numpy as np
from .annular import right_gmres


def mfft(f):
    n = f.shape[1]
    n2 = n//2
    fh = np.fft.fft(f, axis=1)
    return np.concatenate([fh[:, :n2], fh[:, n2 + 1:]], axis=1)


def mifft(fh):
    M, ns = fh.shape
    n2 = (ns + 1)//2
    t = np.zeros((M, ns + 1), dtype=complex)
    t[:, :n2] = fh[:, :n2]
    t[:, n2 + 1:] = fh[:, n2:]
    return np.fft.ifft(t, axis=1)


def mifftr(fh):
    return mifft(fh).real


class AnnularStokesSolver(object):
    def __init__(self, AAG, mu=1.0):
        self.AAG, self.mu = AAG, mu
        self.M, self.ns, self.n = AAG.M, AAG.ns, AAG.n
        if self.ns != self.n - 1:
            raise Exception('the annular Stokes solver works on the Nyquist-dropped geometry')
        self.NU = self.M*self.ns
        self.NP = (self.M - 1)*self.ns
        self.NB = 2*self.NU + self.NP
        self._construct()

    def mode_matrices(self):
        """(ns, 3M-1, 3M-1) complex collocation matrices of the concentric-circle operator (stokes.py:127-150)"""
        AAG, CO, M, ns, mu = self.AAG, self.AAG.CO, self.M, self.ns, self.mu
        ap0, ap1 = AAG.approx_psi0, AAG.approx_psi1
        ai1, ai2 = AAG.approx_inv_psi1, AAG.approx_inv_psi2
        ks = AAG.ks
        D01, D12, R01, R12, R02 = CO.D01, CO.D12, CO.R01, CO.R12, CO.R02
        A0 = ai2[:, None]*(D12 @ (ap1[:, None]*D01))
        A1 = R12 @ (ai1[:, None]*R01)
        K = np.zeros((ns, 3*M - 1, 3*M - 1), dtype=complex)
        LL = A0[None] - (ks**2)[:, None, None]*A1[None]
        diag = -LL + (ai2**2)[:, None]*R02
        cross = (2*ai2**2)[:, None]*R02
        ik = 1j*ks[:, None, None]
        K[:, 0:M - 2, 0:M] = mu*diag
        K[:, 0:M - 2, M:2*M] = mu*ik*cross[None]
        K[:, 0:M - 2, 2*M:] = D12
        K[:, M - 2, 0:M] = CO.ibc_dirichlet[0]
        K[:, M - 1, 0:M] = CO.obc_dirichlet[0]
        K[:, M:2*M - 2, 0:M] = -mu*ik*cross[None]
        K[:, M:2*M - 2, M:2*M] = mu*diag
        K[:, M:2*M - 2, 2*M:] = ik*(ai2[:, None]*R12)[None]
        K[:, 2*M - 2, M:2*M] = CO.ibc_dirichlet[0]
        K[:, 2*M - 1, M:2*M] = CO.obc_dirichlet[0]
        K[:, 2*M:, 0:M] = ai1[:, None]*(D01*ap0[None, :])
        K[:, 2*M:, M:2*M] = ik*(ai1[:, None]*R01)[None]
        K[0, 2*M:, 2*M:] += CO.VI1[0]
        return K

    def _construct(self):
        self.KINV = np.linalg.inv(self.mode_matrices())

    def _extract(self, fh):
        M, ns = self.M, self.ns
        return (fh[:self.NU].reshape(M, ns), fh[self.NU:2*self.NU].reshape(M, ns),
                fh[2*self.NU:].reshape(M - 1, ns))

    def preconditioner(self, ffh):
        frh, fth, fph = self._extract(ffh)
        ff = np.vstack([frh, fth, fph]).T                         # (ns, 3M-1)
        out = np.einsum('ijk,ik->ij', self.KINV, ff).T
        M = self.M
        return np.concatenate([out[:M].ravel(), out[M:2*M].ravel(), out[2*M:].ravel()])

    def apply(self, uuh):
        """stokes.py:321-385 (_apply_optim_real)"""
        AAG, RAG, CO, mu = self.AAG, self.RAG, self.AAG.CO, self.mu
        D01, D12, R01, R12, R02, VI1 = CO.D01, CO.D12, CO.R01, CO.R12, CO.R02, CO.VI1
        ibcd, obcd = CO.ibc_dirichlet, CO.obc_dirichlet
        iks = 1j*AAG.ks
        psi0, psi1, ipsi1, ipsi2 = RAG.psi0, RAG.psi1, RAG.inv_psi1, RAG.inv_psi2
        m3r, m3t = RAG.ipsi_DR_ipsi_DT_psi2, RAG.ipsi_DT_ipsi_DR_psi2
        combo1 = 2*RAG.DR_psi2*ipsi2**2
        combo2 = RAG.DR_psi2**2*ipsi2**2
        urh, uth, ph = self._extract(uuh)
        ur, ut, p = mifftr(urh), mifftr(uth), mifftr(ph)
        dur, dut = mifftr(iks*urh), mifftr(iks*uth)
        W1r, W1t = R02 @ ur, R02 @ ut
        t1 = (R02 @ dut)*combo1
        t2 = W1r*combo2
        t3 = W1t*m3r
        t4 = D12 @ p
        ur_tt = R12 @ mifftr(mfft((R01 @ dur)*ipsi1)*iks)
        ur_rr = D12 @ ((D01 @ ur)*psi1)
        frh = mfft(mu*(-(ur_rr + ur_tt)*ipsi2 + t1 + t2 + t3) + t4)
        t1 = (R02 @ dur)*combo1
        t2 = W1t*combo2
        t3 = W1r*m3t
        t4 = (R12 @ mifftr(ph*iks))*ipsi2
        W2 = R01 @ dut
        ut_tt = R12 @ mifftr(mfft(W2*ipsi1)*iks)
        ut_rr = D12 @ ((D01 @ ut)*psi1)
        fth = mfft(mu*(-(ut_rr + ut_tt)*ipsi2 - t1 + t2 - t3) + t4)
        fph = mfft((D01 @ (ur*psi0) + W2)*ipsi1)
        fph[:, 0] += (VI1 @ ph)[0, 0]
        return np.concatenate([frh.ravel(), (ibcd @ urh).ravel(), (obcd @ urh).ravel(),
                               fth.ravel(), (ibcd @ uth).ravel(), (obcd @ uth).ravel(), fph.ravel()])

    def solve(self, RAG, fr, ft, irg, itg, org, otg, tol=1e-8, restart=None, maxiter=None, verbose=False):
        self.RAG = RAG
        CO, M, n = self.AAG.CO, self.M, self.n
        ffr = np.vstack([CO.R02 @ fr, irg[None, :], org[None, :]])
        fft_ = np.vstack([CO.R02 @ ft, itg[None, :], otg[None, :]])
        ffh = np.concatenate([mfft(ffr).ravel(), mfft(fft_).ravel(), np.zeros(self.NP, dtype=complex)])
        res, info, resids = right_gmres(self.apply, ffh, self.preconditioner, tol=tol,
                                        restart=20 if restart is None else restart,
                                        maxiter=100 if maxiter is None else maxiter, verbose=verbose)
        self.iterations_last_call = len(resids)
        urh, uth, ph = self._extract(res)
        return mifft(urh).real, mifft(uth).real, CO.P10 @ mifft(ph).real
