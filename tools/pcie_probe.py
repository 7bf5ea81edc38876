"""Device-to-host copy rate of one GPU as a function of the number of copies, alone and with a host-to-device copy running beside it
(design input for the chunk sizes of the host forms, csrc/context.cu):  python tools/pcie_probe.py"""
import torch, time
n = 768 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(544 << 20, dtype=torch.uint8, device="cuda"); h2 = torch.empty(544 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both, chunks):
    torch.cuda.synchronize(); t = time.perf_counter()
    step = n // chunks
    with torch.cuda.stream(s1):
        for c in range(chunks): h[c*step:(c+1)*step].copy_(d[c*step:(c+1)*step], non_blocking=True)
    if both:
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for both in (0, 1):
    for chunks in (1, 32, 128):
        run(both, chunks); ts = [run(both, chunks) for _ in range(5)]
        print(f"D2H {n/1e6:.0f} MB in {chunks} copies, concurrent H2D={both}: {n/min(ts)/1e9:.1f} GB/s")
