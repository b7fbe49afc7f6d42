#!/usr/bin/env python
"""Pinned-memory PCIe copy bandwidth on this box (D2H alone, H2D alone, both at once) -- the ceiling of bench.py's e2e."""
import torch, time
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=8):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def d2h():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
def both():
    d2h(); h2d()
def d2h_chunks():
    with torch.cuda.stream(s1):
        c = 7_620_000
        for o in range(0, n - c, c): h[o:o + c].copy_(d[o:o + c], non_blocking=True)
print("D2H  %.1f GB/s" % (n / run(d2h) / 1e9))
print("H2D  %.1f GB/s" % (n / run(h2d) / 1e9))
t = run(both)
print("both %.1f GB/s each direction" % (n / t / 1e9))
print("D2H in 7.6 MB chunks %.1f GB/s" % ((n // 7_620_000) * 7_620_000 / run(d2h_chunks) / 1e9))
# both directions in slab-sized pieces, like the host-pointer pipeline of wpsgpu_metgrid_flush
def both_chunks():
    ci, co = 4_200_000, 7_620_000
    with torch.cuda.stream(s2):
        for o in range(0, n - ci, ci): d2[o:o + ci].copy_(h2[o:o + ci], non_blocking=True)
    with torch.cuda.stream(s1):
        for o in range(0, n - co, co): h[o:o + co].copy_(d[o:o + co], non_blocking=True)
t = run(both_chunks)
print("both, slab-sized pieces: %.1f GB/s each direction (%.2f ms)" % (n / t / 1e9, t * 1e3))
# same with a compute kernel running on a third stream
x = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
s3 = torch.cuda.Stream()
def both_chunks_compute():
    both_chunks()
    with torch.cuda.stream(s3):
        for _ in range(20): x.mul_(1.0001)
t = run(both_chunks_compute)
print("both + HBM-bound kernels: %.1f GB/s each direction (%.2f ms)" % (n / t / 1e9, t * 1e3))
# two streams per direction, slab-sized pieces
s4, s5 = torch.cuda.Stream(), torch.cuda.Stream()
def both_chunks_2s():
    ci, co = 4_200_000, 7_620_000
    for k, o in enumerate(range(0, n - ci, ci)):
        with torch.cuda.stream(s2 if k % 2 else s4): d2[o:o + ci].copy_(h2[o:o + ci], non_blocking=True)
    for k, o in enumerate(range(0, n - co, co)):
        with torch.cuda.stream(s1 if k % 2 else s5): h[o:o + co].copy_(d[o:o + co], non_blocking=True)
t = run(both_chunks_2s)
print("both, slab-sized pieces, 2 streams per direction: %.1f GB/s each direction (%.2f ms)" % (n / t / 1e9, t * 1e3))
def both_big():
    ci, co = 16 * 4_200_000, 16 * 7_620_000
    with torch.cuda.stream(s2):
        for o in range(0, n - ci + 1, ci): d2[o:o + ci].copy_(h2[o:o + ci], non_blocking=True)
    with torch.cuda.stream(s1):
        for o in range(0, n - co + 1, co): h[o:o + co].copy_(d[o:o + co], non_blocking=True)
t = run(both_big)
print("both, 16-slab pieces: H2D %d MB + D2H %d MB in %.2f ms" % ((n // (16 * 4_200_000)) * 67, (n // (16 * 7_620_000)) * 122, t * 1e3))
