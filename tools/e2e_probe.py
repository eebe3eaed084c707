"""Where does the host-buffer (e2e) path spend its time?  PCIe copy speed vs spz_frame_{compress,decompress}_host."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import sprintz_b200 as spz
from sprintz_b200.synth import ar1_stream
L = spz.lib()
cfgs = {"c2": (1, 8, 1, 1 << 30, 1024), "c3": (2, 16, 1, 4 << 30, 1024)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
esz, D, pred, nbytes, cr = cfgs[name]
rows = nbytes // (esz * D); n = rows * D
x = ar1_stream(rows, D, esz, seed=12345, device="cuda")
xh = torch.empty(n, dtype=x.dtype, pin_memory=True); xh.copy_(x)
d = torch.empty_like(x)
def t(f, k=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
raw = n * esz
print("H2D GB/s", raw / t(lambda: d.copy_(xh, non_blocking=True)) / 1e9)
print("D2H GB/s", raw / t(lambda: xh.copy_(d, non_blocking=True)) / 1e9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
oh = torch.empty(n, dtype=x.dtype, pin_memory=True)
def both():
    with torch.cuda.stream(s1): d.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): oh.copy_(x, non_blocking=True)
print("duplex GB/s per direction", raw / t(both) / 1e9)
del d
fcap = int(L.spz_frame_max_bytes(esz, n, D, cr))
fh = torch.empty(fcap, dtype=torch.uint8, pin_memory=True)
fb = [0]
def comp():
    fb[0] = L.spz_frame_compress_host(pred, esz, xh.data_ptr(), n, D, cr, fh.data_ptr(), fcap)
def dec():
    L.spz_frame_decompress_host(fh.data_ptr(), fb[0], oh.data_ptr(), n)
tc = t(comp, 3); td = t(dec, 3)
print(f"compress_host {tc*1e3:.1f} ms ({raw/tc/1e9:.1f} GB/s raw), frame {fb[0]/1e6:.0f} MB; decompress_host {td*1e3:.1f} ms ({raw/td/1e9:.1f} GB/s raw); e2e {raw/(tc+td)/1e9:.1f} GB/s")
