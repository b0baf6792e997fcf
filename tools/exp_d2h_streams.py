#!/usr/bin/env python
"""Device-to-host copy rate of this box with the bytes spread over 1, 2, 4 streams (copy engines), 256 MB pieces, page-locked
target; and the same with host-to-device copies running beside them. Decides whether the library's copier should use more than
one stream. Prints one JSON line."""
import json, time, torch
dev = torch.device("cuda:0")
N = 8 << 30
PIECE = 256 << 20
src = torch.empty(N, dtype=torch.uint8, device=dev)
dst = torch.empty(N, dtype=torch.uint8).pin_memory()
up_src = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
up_dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
res = {}
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, off in enumerate(range(0, N, PIECE)):
            with torch.cuda.stream(streams[k % ns]):
                dst[off:off + PIECE].copy_(src[off:off + PIECE], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    res[f"d2h_{ns}_streams_GBps"] = round(N / dt / 1e9, 1)
# one D2H stream with an H2D stream beside it (what the pipelined decode does)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for k, off in enumerate(range(0, N, PIECE)):
    with torch.cuda.stream(s1):
        dst[off:off + PIECE].copy_(src[off:off + PIECE], non_blocking=True)
    if k % 8 == 0:
        with torch.cuda.stream(s2):
            up_dst.copy_(up_src, non_blocking=True)
torch.cuda.synchronize()
res["d2h_beside_h2d_GBps"] = round(N / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(res))
