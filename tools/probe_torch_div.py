"""How does torch CUDA evaluate ``tensor / python_scalar``?  (run under gpurun; output: profiles/r01_torch_div_probe.log)

Compares ``x_cuda / s`` bit for bit with four candidates over 2 M random fp32 values per divisor:
x * fp32(1/fp32(s)) (fp32 reciprocal), x * fp32(1.0/s) (reciprocal in double, rounded once), x / fp32(s) and x / s (IEEE).
Result on torch 2.11.0+cu128 / B200: always and only the second one matches -- the semantics of the library's default
``BDL_DIV_RECIP`` mode (include/bdl.h: bdl_scalars.inv_*).
"""
import numpy as np, torch
torch.manual_seed(0)
dev = torch.device("cuda:0")
n = 2_000_000
x = (torch.rand(n, dtype=torch.float32) * 10 + 1e-6)
xd = x.to(dev)
x64 = x.numpy().astype(np.float64)
def mism(got, want64):
    want = want64.astype(np.float32)
    return int((got.cpu().numpy().view(np.uint32) != want.view(np.uint32)).sum())
for s in [0.1, 1 - 0.9 ** 1, 1 - 0.9 ** 2, 1 - 0.9 ** 3, 1 - 0.99 ** 1, 1 - 0.99 ** 2, 1 - 0.99 ** 3, 0.9 ** 2, 1.3, 1e-3, 3.0, 7.0, 1840 * 3.0, 30000 * 1e3]:
    got = xd / s
    inv_f = np.float32(1.0) / np.float32(s)                 # fp32 reciprocal of the fp32-rounded scalar
    inv_d = np.float32(1.0 / float(s))                      # double reciprocal, then rounded to fp32
    a = mism(got, x64 * np.float64(inv_f))
    b = mism(got, x64 * np.float64(inv_d))
    c = mism(got, x64 / np.float64(np.float32(s)))
    d = mism(got, x64 / float(s))
    print(f"s={s!r:24} inv_f==inv_d: {inv_f == inv_d!s:5}  mismatches: x*fl32(1/fl32(s))={a}  x*fl32(1/s_double)={b}  x/fl32(s)={c}  x/s_double={d}")
