import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, scipy.special as ss, torch
import ipde_b200.device as D
from ipde_b200._lib import lib
from ipde_b200 import ewald as E
kind, kappa, N = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
h = 2.4/N; x0 = -1.2
rng = np.random.default_rng(3)
ns = 300
th = np.sort(rng.uniform(0, 2*np.pi, ns))
sx, sy = 0.8*np.cos(th) + 0.013, 0.75*np.sin(th) - 0.007
q = rng.standard_normal(ns)
xv = x0 + h*np.arange(N)
X, Y = np.meshgrid(xv, xv, indexing='ij')
sel = (X/0.8)**2 + (Y/0.75)**2 < 0.93**2
gix, giy = np.nonzero(sel)
px_, py_ = 0.74*np.cos(th[::2]) + 0.001, 0.7*np.sin(th[::2])
ev = D.FreeSpaceEvaluator(kind, kappa, x0, x0, h, sx, sy, grid_nodes=(gix, giy), grid_out_pos=np.arange(gix.size), points=(px_, py_))
P = ev.params
print('box', ev.px, ev.py, ev.x0, ev.y0, 'ilo', ev.ilo_x, ev.ilo_y, 'L', ev.L, 'ntiles', ev.ntiles)
def buf(which, shape):
    p = ctypes.c_void_p(); n = ctypes.c_int64()
    lib().ipde_ewald_debug_buffer(ev.h_plan, which, ctypes.byref(p), ctypes.byref(n))
    t = torch.empty(n.value, dtype=torch.float64, device='cuda')
    ctypes.CDLL('libcudart.so').cudaMemcpy(ctypes.c_void_p(t.data_ptr()), p, ctypes.c_size_t(8*n.value), 3)
    return t.cpu().numpy().reshape(shape)
out_grid = D.zeros((gix.size,))
qd = D.to_dev(q)
# grid only first
lib().ipde_ewald_apply(ev.h_plan, D.ptr(qd), D.ptr(out_grid), 0, 0, D.stream())
torch.cuda.synchronize()
Sw = buf(0, (ev.px, ev.py)); U = buf(1, (ev.px, ev.py)); Lc = buf(2, (ev.ntiles, 16, 16))
xp = ev.x0 + h*np.arange(ev.px); yp = ev.y0 + h*np.arange(ev.py)
XP, YP = np.meshgrid(xp, yp, indexing='ij')
xi = P.xi
ck = np.exp(-kappa**2/(4*xi*xi)) if kind == 1 else 1.0
Sw_ref = np.zeros_like(Sw)
for s in range(ns):
    dd2 = (XP - sx[s])**2 + (YP - sy[s])**2
    Sw_ref += np.where(dd2 < P.rw**2, q[s]*ck*2*xi*xi/np.pi*np.exp(-2*xi*xi*dd2), 0.0)
print('Sw err', np.abs(Sw - Sw_ref).max(), 'scale', np.abs(Sw_ref).max())
kx = np.fft.fftfreq(ev.px, h/(2*np.pi)); ky = np.fft.fftfreq(ev.py, h/(2*np.pi))
KX, KY = np.meshgrid(kx, ky, indexing='ij'); K2 = KX**2 + KY**2
GT = E.trunc_kernel_hat(kind, kappa, ev.L, np.sqrt(K2))
U_ref = np.fft.ifft2(GT*np.exp(-K2/(8*xi*xi))*np.fft.fft2(Sw_ref)).real
print('U err', np.abs(U - U_ref).max(), 'scale', np.abs(U_ref).max())
G = (lambda r: -np.log(r)/(2*np.pi)) if kind == 0 else (lambda r: ss.k0(kappa*r)/(2*np.pi))
d = np.hypot(X[sel][:, None] - sx, Y[sel][:, None] - sy)
exact = G(d).dot(q)
tab = E.local_tables(kind, kappa, xi, P.rc)
loc_ref = (E.eval_local(tab, d**2)*(d < P.rc)).dot(q)
bix, biy = gix - ev.ilo_x, giy - ev.ilo_y
far_ref = U_ref[bix, biy]
print('ref total err', np.abs(far_ref + loc_ref - exact).max())
print('far err at targets', np.abs(U[bix, biy] - far_ref).max())
got = out_grid.cpu().numpy()
print('local err', np.abs(got - U[bix, biy] - loc_ref).max())
print('grid total err', np.abs(got - exact).max()/np.abs(exact).max())
pts = ev(qd)
d2 = np.hypot(px_[:, None] - sx, py_[:, None] - sy)
ex2 = G(d2).dot(q)
print('points err', np.abs(pts.cpu().numpy() - ex2).max()/np.abs(ex2).max())
tix, tiy, tptr, tsrc, tmap = E.build_tiles(sx, sy, ev.x0, ev.y0, h, ev.px, ev.py, P.rc)
t = 5
ix = tix[t]*16 + np.arange(16); iy = tiy[t]*16 + np.arange(16)
XX, YY = np.meshgrid(ev.x0 + ix*h, ev.y0 + iy*h, indexing='ij')
dd = np.hypot(XX.ravel()[:, None] - sx, YY.ravel()[:, None] - sy)
ref = (np.where((dd < P.rc) & (dd > 0), E.eval_local(tab, np.minimum(dd**2, P.rc**2*0.999999)), 0.0)).dot(q).reshape(16, 16)
print('tile', t, tix[t], tiy[t], 'Lc err', np.abs(Lc[t] - ref).max(), 'scale', np.abs(ref).max())
print(Lc[t][:3, :3]); print(ref[:3, :3])
print('transposed?', np.abs(Lc[t].T - ref).max())
errs = []
for t in range(len(tix)):
    ix = tix[t]*16 + np.arange(16); iy = tiy[t]*16 + np.arange(16)
    XX, YY = np.meshgrid(ev.x0 + ix*h, ev.y0 + iy*h, indexing='ij')
    dd = np.hypot(XX.ravel()[:, None] - sx, YY.ravel()[:, None] - sy)
    ref = (np.where((dd < P.rc) & (dd > 0), E.eval_local(tab, np.minimum(dd**2, P.rc**2*0.999999)), 0.0)).dot(q).reshape(16, 16)
    errs.append(np.abs(Lc[t] - ref).max())
print('max Lc err over tiles', max(errs), int(np.argmax(errs)))
tt = tmap[bix//16, biy//16]
loc_idx = tt*256 + (bix % 16)*16 + (biy % 16)
Lflat = Lc.reshape(-1)
mine = np.where(tt >= 0, Lflat[np.maximum(loc_idx, 0)], 0.0)
print('gathered Lc vs loc_ref', np.abs(mine - loc_ref).max())
print('got - U - mine', np.abs(got - U[bix, biy] - mine).max())
print(got[:4], U[bix, biy][:4], mine[:4], loc_ref[:4])
