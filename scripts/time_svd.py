import torch, time, sys
torch.cuda.set_device(0)
for n in [int(a) for a in sys.argv[1:]]:
    A = torch.randn(n, n, dtype=torch.float64, device='cuda')
    # make it ill-conditioned like the QFS matrices
    U, _ = torch.linalg.qr(A)
    s = torch.logspace(0, -14, n, dtype=torch.float64, device='cuda')
    A = (U*s) @ torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device='cuda'))[0]
    for drv in ['gesvd', 'gesvdj', 'gesvda']:
        torch.cuda.synchronize(); t = time.time()
        try:
            U2, S2, V2 = torch.linalg.svd(A, full_matrices=False, driver=drv)
            torch.cuda.synchronize()
            err = (S2[:n//2] - s[:n//2]).abs().max().item()
            rec = ((U2*S2) @ V2 - A).abs().max().item()
            print(n, drv, '%.2fs' % (time.time() - t), 'sv err %.1e recon %.1e' % (err, rec), flush=True)
        except Exception as e:
            print(n, drv, 'failed', str(e)[:80], flush=True)
    torch.cuda.synchronize(); t = time.time()
    w, Q = torch.linalg.eigh(A.T @ A)
    torch.cuda.synchronize(); print(n, 'eigh(AtA) %.2fs' % (time.time() - t), flush=True)
    torch.cuda.synchronize(); t = time.time()
    Q, R = torch.linalg.qr(A)
    torch.cuda.synchronize(); print(n, 'qr %.2fs' % (time.time() - t), flush=True)
    torch.cuda.synchronize(); t = time.time()
    LU, piv = torch.linalg.lu_factor(A)
    torch.cuda.synchronize(); print(n, 'lu %.2fs' % (time.time() - t), flush=True)
