import torch, numpy as np, time
n = 2364619
a = np.random.rand(n); out = np.empty(n)
d = torch.empty(n, dtype=torch.float64, device='cuda')
h = torch.empty(n, dtype=torch.float64).pin_memory()
def T(fn, reps=20):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t)/reps*1e3
print('threads', torch.get_num_threads())
print('numpy slice copy into pinned  %.2f ms' % T(lambda: h.numpy().__setitem__(slice(None), a)))
print('torch cpu copy into pinned    %.2f ms' % T(lambda: h.copy_(torch.from_numpy(a))))
print('pinned H2D                    %.2f ms' % T(lambda: d.copy_(h, non_blocking=True)))
print('pageable H2D                  %.2f ms' % T(lambda: d.copy_(torch.from_numpy(a))))
print('pinned D2H                    %.2f ms' % T(lambda: h.copy_(d, non_blocking=True)))
print('pageable D2H                  %.2f ms' % T(lambda: torch.from_numpy(out).copy_(d)))
print('torch cpu copy from pinned    %.2f ms' % T(lambda: torch.from_numpy(out).copy_(h)))
print('numpy .copy() of pinned       %.2f ms' % T(lambda: h.numpy().copy()))
print('np.empty(n)                   %.2f ms' % T(lambda: np.empty(n)))
