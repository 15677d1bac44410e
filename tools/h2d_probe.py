"""Host -> device rate of the end-to-end leg's inputs, measured alone (nothing else on the GPU): the five FPN maps of
one B=16 step (367 MB, pinned) per copy, as bench.py's e2e pass copies them, and the same bytes as one tensor.
bench.py's e2e line moves these bytes every step, so  bytes / this rate  is the floor of its step time."""
import torch

torch.cuda.set_device(0)
B, hws = 16, [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
host = [torch.randn(B, 256, h, w).pin_memory() for h, w in hws]
dev = [torch.empty_like(t, device="cuda") for t in host]
nbytes = sum(t.numel() * 4 for t in host)
big_h = torch.empty(nbytes // 4).pin_memory()
big_d = torch.empty(nbytes // 4, device="cuda")
s = torch.cuda.Stream()


def timed(fn, n=10):
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def five():
    for d, h in zip(dev, host):
        d.copy_(h, non_blocking=True)


t5 = timed(five)
t1 = timed(lambda: big_d.copy_(big_h, non_blocking=True))
print("h2d probe: %d bytes; five tensors %.2f ms = %.1f GB/s; one tensor %.2f ms = %.1f GB/s"
      % (nbytes, t5, nbytes / t5 / 1e6, t1, nbytes / t1 / 1e6))
