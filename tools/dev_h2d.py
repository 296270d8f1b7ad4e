"""Raw pinned-host <-> device copy bandwidth on this box (context for the e2e number: frames in, keypoints out)."""
import torch, time
n = 315 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    with torch.cuda.stream(s):
        for _ in range(3): a.copy_(b, non_blocking=True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        for _ in range(10): a.copy_(b, non_blocking=True)
        e1.record(s)
    e1.synchronize()
    print(name, "%.1f GB/s" % (10 * n / e0.elapsed_time(e1) / 1e6))
# both directions at once
s2 = torch.cuda.Stream()
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
e0, e1, f1 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
torch.cuda.synchronize()
e0.record(s)
with torch.cuda.stream(s):
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(s)
with torch.cuda.stream(s2):
    for _ in range(10): h2.copy_(d2, non_blocking=True)
    f1.record(s2)
torch.cuda.synchronize()
print("h2d with concurrent d2h: %.1f GB/s" % (10 * n / e0.elapsed_time(e1) / 1e6))
