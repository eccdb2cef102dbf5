import numpy as np
rng = np.random.default_rng(0)
def rz32(x):
    # truncate float64 -> float32 toward zero
    f = x.astype(np.float32)
    bad = np.abs(f.astype(np.float64)) > np.abs(x)
    f[bad] = np.nextafter(f[bad], np.float32(0))
    return f
def split16(v):
    h = v.astype(np.float16).astype(np.float32); l = (v - h).astype(np.float16).astype(np.float32); return h, l
def run(K, M=256, N=64, relu=True):
    x = rng.standard_normal((M, K)).astype(np.float32)
    if relu: x = np.maximum(x, 0)
    w = (rng.standard_normal((K, N)) * np.sqrt(2.0 / K)).astype(np.float32)
    exact = x.astype(np.float64) @ w.astype(np.float64)
    xh, xl = split16(x); wh, wl = split16(w)
    # passes: (xh,wh),(xl,wh) interleaved per k-step, then (xh, wl)
    seq = []
    for k0 in range(0, K, 16):
        seq.append((xh[:, k0:k0+16], wh[k0:k0+16]))
        seq.append((xl[:, k0:k0+16], wh[k0:k0+16]))
    for k0 in range(0, K, 16):
        seq.append((xh[:, k0:k0+16], wl[k0:k0+16]))
    acc = np.zeros((M, N), np.float32)
    for a, b in seq:
        acc = rz32(acc.astype(np.float64) + a.astype(np.float64) @ b.astype(np.float64))
    accrn = np.zeros((M, N), np.float32)
    for a, b in seq:
        accrn = (accrn.astype(np.float64) + a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
    L = len(seq)
    rel = (acc.astype(np.float64) - exact) / np.abs(exact)
    sel = np.abs(exact) > 0.3 * np.abs(exact).std()
    # signed bias toward zero: (|acc| - |exact|)/|exact|
    sb = (np.abs(acc.astype(np.float64)) - np.abs(exact)) / np.abs(exact)
    l2 = lambda a: np.sqrt(((a.astype(np.float64) - exact) ** 2).sum() / (exact ** 2).sum())
    # best single scale
    s = (acc.astype(np.float64) * exact).sum() / (acc.astype(np.float64) ** 2).sum()
    print(f"K={K} L={L}: relL2 RZ {l2(acc):.2e}  RN {l2(accrn):.2e}  mean signed bias {sb[sel].mean():.2e} = {sb[sel].mean()/2**-24/L:.3f} ulp24/MMA; best scale-1 {s-1:.2e}; after scale {l2(acc*s):.2e}")
for K in (128, 512, 1152, 2304):
    run(K)
run(1152, relu=False)
