import torch, time
N, F = 38400, 9472
h = torch.empty(N, F, dtype=torch.float32).pin_memory()
d = torch.empty(N, F, dtype=torch.float32, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
gb = N*F*4/1e9
print("contiguous H2D GB/s", gb / t(lambda: d.copy_(h, non_blocking=True)))
for Fc in (592, 1184, 2368, 4736):
    dd = torch.empty(N, Fc, dtype=torch.float32, device="cuda")
    def f():
        for f0 in range(0, F, Fc): dd.copy_(h[:, f0:f0+Fc], non_blocking=True)
    print("2D chunks of", Fc, "GB/s", gb / t(f))
hT = torch.empty(F, N, dtype=torch.float32).pin_memory()
print("contiguous FN chunks GB/s", gb / t(lambda: [d.view(-1)[:N*1184].copy_(hT[f0:f0+1184].reshape(-1), non_blocking=True) for f0 in range(0, F, 1184)]))
o = torch.empty(F * 1200, dtype=torch.int32, device="cuda"); oh = torch.empty(F*1200, dtype=torch.int32).pin_memory()
print("D2H GB/s", F*1200*4/1e9 / t(lambda: oh.copy_(o, non_blocking=True)))
