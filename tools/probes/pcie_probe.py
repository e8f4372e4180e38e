"""PCIe probe for the e2e leg: pinned H2D / D2H bandwidth alone and concurrently (the bytes of one C4 step: 98 MB up, 81 MB down)."""
import time
import torch
up, down = 97_779_712, 81_002_496
hu = torch.empty(up, dtype=torch.uint8).pin_memory(); du = torch.empty(up, dtype=torch.uint8, device="cuda")
hd = torch.empty(down, dtype=torch.uint8).pin_memory(); dd = torch.empty(down, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_up, do_down, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        if do_up:
            with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
        if do_down:
            with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for _ in range(2): run(True, True, 3)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D alone {a*1e3:.3f} ms ({up/a/1e9:.1f} GB/s) | D2H alone {b*1e3:.3f} ms ({down/b/1e9:.1f} GB/s) | both {c*1e3:.3f} ms (up {up/c/1e9:.1f} + down {down/c/1e9:.1f} GB/s)")
