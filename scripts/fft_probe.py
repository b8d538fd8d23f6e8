"""cuFFT probe on B200: 3-D r2c/c2r and strided 1-D c2c throughput vs transform length (decides the fine-mesh plan)."""
import torch, time, json
dev = "cuda"
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
for N in (256, 280, 288, 300, 304, 320, 324, 336, 360, 384, 560, 576):
    B = 4 if N < 400 else 1
    x = torch.randn(B, N, N, N, device=dev)
    f = torch.fft.rfftn(x, dim=(1, 2, 3))
    r2c = t(lambda: torch.fft.rfftn(x, dim=(1, 2, 3)))
    c2r = t(lambda: torch.fft.irfftn(f, s=(N, N, N), dim=(1, 2, 3)))
    c1 = {}
    for d in (1, 2, 3):
        c1[d] = t(lambda: torch.fft.fft(f, dim=d))
    gb = B * N**3 * 4 / 1e9
    res[N] = dict(batch=B, r2c_ms=r2c / B, c2r_ms=c2r / B, c2c_dim_ms={k: v / B for k, v in c1.items()},
                  r2c_GBs=2 * gb / (r2c * 1e-3), c2c_GBs={k: 2 * B * f[0].numel() * 8 / 1e9 / (v * 1e-3) for k, v in c1.items()})
    print(N, json.dumps(res[N]), flush=True)
    del x, f
