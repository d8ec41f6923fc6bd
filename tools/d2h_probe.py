"""Device->host bandwidth of this box for the 201 MB fp32 frame batch: one stream vs the copy split over 2 / 4 streams."""
import torch, time
x = torch.empty(64, 3, 512, 512, device="cuda")
h = torch.empty(64, 3, 512, 512).pin_memory()
def run(nsplit, n=20):
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    xs, hs = x.chunk(nsplit), h.chunk(nsplit)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        for s, a, b in zip(streams, xs, hs):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    print("D2H 201 MB over %d stream(s): %.2f ms = %.1f GB/s" % (nsplit, dt * 1e3, x.numel() * 4 / dt / 1e9), flush=True)
for k in (1, 2, 4, 1):
    run(k)
hp = torch.empty(64, 3, 64, 64).pin_memory(); d = torch.empty(64, 3, 64, 64, device="cuda")
torch.cuda.synchronize(); t0 = time.time()
for _ in range(50): d.copy_(hp, non_blocking=True)
torch.cuda.synchronize(); print("H2D 3 MB: %.1f GB/s" % (hp.numel() * 4 * 50 / (time.time() - t0) / 1e9))
