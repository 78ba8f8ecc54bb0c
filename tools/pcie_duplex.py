import torch, time
n_up, n_dn = 1<<30, 1<<29
hu = torch.empty(n_up, dtype=torch.uint8, pin_memory=True); du = torch.empty(n_up, dtype=torch.uint8, device="cuda")
hd = torch.empty(n_dn, dtype=torch.uint8, pin_memory=True); dd = torch.empty(n_dn, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
a = run(1, 0); b = run(0, 1); c = run(1, 1)
print("H2D 1 GiB alone %.2f ms = %.1f GB/s" % (a*1e3, n_up/a/1e9))
print("D2H 0.5 GiB alone %.2f ms = %.1f GB/s" % (b*1e3, n_dn/b/1e9))
print("both %.2f ms = %.1f GB/s total" % (c*1e3, (n_up+n_dn)/c/1e9))
# down twice (1 GiB each way)
hd2 = torch.empty(n_up, dtype=torch.uint8, pin_memory=True); dd2 = torch.empty(n_up, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0=time.perf_counter()
with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
with torch.cuda.stream(s2): hd2.copy_(dd2, non_blocking=True)
torch.cuda.synchronize(); d=time.perf_counter()-t0
print("1 GiB each way %.2f ms = %.1f GB/s total" % (d*1e3, 2*n_up/d/1e9))
