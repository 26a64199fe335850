"""How fast is cuFFT on the padded sizes the reference's default fft_padding = 1.2 produces (9830 = 2*5*983)?"""
import json
import torch


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best, 3)


out = {}
x = torch.randn(2, 9830, 9830, dtype=torch.complex64, device="cuda")
out["fft2 2x9830^2"] = t(lambda: torch.fft.ifft2(x, norm="forward"))
out["fft rows 2x9830x[9830]"] = t(lambda: torch.fft.fft(x, dim=-1))
out["fft cols 2x[9830]x9830"] = t(lambda: torch.fft.fft(x, dim=-2))
y = x.reshape(-1)[:196600 * 983].reshape(196600, 983)
out["fft 196600x[983]"] = t(lambda: torch.fft.fft(y, dim=-1))
y2 = x.reshape(-1)[:19660 * 9830].reshape(19660, 9830)
out["fft 19660x[9830]"] = t(lambda: torch.fft.fft(y2, dim=-1))
del x, y, y2
z = torch.randn(2, 10240, 10240, dtype=torch.complex64, device="cuda")
out["fft2 2x10240^2"] = t(lambda: torch.fft.ifft2(z, norm="forward"))
out["fft rows 2x10240x[10240]"] = t(lambda: torch.fft.fft(z, dim=-1))
zz = z.reshape(-1)[:98304 * 2048].reshape(98304, 2048)
out["fft 98304x[2048]"] = t(lambda: torch.fft.fft(zz, dim=-1))
del z, zz
w = torch.randn(2, 4915, 4915, dtype=torch.complex64, device="cuda")
out["fft2 2x4915^2"] = t(lambda: torch.fft.ifft2(w, norm="forward"))
w = torch.randn(2, 9830, 9830, dtype=torch.complex128, device="cuda")
out["fft2 2x9830^2 c128"] = t(lambda: torch.fft.ifft2(w, norm="forward"))
print(json.dumps(out, indent=1))
