"""PCIe ceilings of the box: pinned H2D / D2H alone and both at once (run under gpurun)."""
import torch, time
n = 2 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
print("H2D alone %.1f GB/s" % (n / t(h2d) / 1e9))
print("D2H alone %.1f GB/s" % (n / t(d2h) / 1e9))
tb = t(both); print("H2D+D2H concurrent: %.1f GB/s each direction" % (n / tb / 1e9))
# strided 2-D through torch (not cudaMemcpy2D; for reference only): every other 11520-byte row
rows = n // 23040
hv = h_in[:rows * 23040].view(rows, 23040)[:, :11520]
dv = d_in[:rows * 11520].view(rows, 11520)
def h2d2():
    with torch.cuda.stream(s1): dv.copy_(hv, non_blocking=True)
print("H2D 2-D strided rows %.1f GB/s" % (rows * 11520 / t(h2d2) / 1e9))
