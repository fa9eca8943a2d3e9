import time, torch
n = 1 << 28  # 1 GiB of float32
host = torch.empty(n, dtype=torch.float32).pin_memory()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def compute():
    with torch.cuda.stream(s1):
        for _ in range(12): torch.mm(a, b)
def copy():
    with torch.cuda.stream(s2):
        dev.copy_(host, non_blocking=True)
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
compute(); copy(); torch.cuda.synchronize()
tc, tp = t(compute), t(copy)
tb = t(lambda: (compute(), copy()))
print(f"compute {tc:.1f} ms, copy {tp:.1f} ms ({n*4/tp/1e6:.1f} GB/s), both {tb:.1f} ms -> overlap {'yes' if tb < 0.8*(tc+tp) else 'NO'}")
