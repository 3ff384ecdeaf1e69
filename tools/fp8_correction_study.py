"""Error study (CPU, numpy): joint GEMM z = act(f+g) . W^T with the BF16X3 split, against a variant whose two
correction terms hi.lo and lo.hi are computed from block-scaled e4m3 operands (what tcgen05 kind::mxf8f6f4 would do).
DESIGN.md section 6, item 0.  Pure numpy emulation of the operand roundings; accumulation in fp64."""
import numpy as np


def bf16_rn(x):
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def e4m3_block(x, block=32):
    """Round to e4m3 (3 mantissa bits, min normal 2^-6, subnormal step 2^-9, max 448) after a power-of-two scale per
    `block` consecutive elements of the last axis (UE8M0 scale, as in the MX formats)."""
    x = np.asarray(x, dtype=np.float64)
    shp = x.shape
    xb = x.reshape(-1, block)
    amax = np.abs(xb).max(axis=1, keepdims=True)
    e = np.ceil(np.log2(np.maximum(amax, 1e-300) / 448.0))
    s = 2.0 ** e
    v = xb / s
    a = np.abs(v)
    ex = np.floor(np.log2(np.maximum(a, 2.0 ** -20)))
    ex = np.maximum(ex, -6.0)                 # subnormals share the exponent of the smallest normal
    step = 2.0 ** (ex - 3)
    q = np.round(a / step) * step
    q = np.minimum(q, 448.0)
    return (np.sign(v) * q * s).reshape(shp)


def study(M=4096, K=640, N=1025, seed=0, trained_scale=1.0):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((M, K)); g = rng.standard_normal((M, K))
    A = np.tanh(f + g)
    W = (rng.random((N, K)) * 2 - 1) / np.sqrt(K) * trained_scale
    z = A @ W.T
    Ah = bf16_rn(A).astype(np.float64); Al = bf16_rn(A - Ah).astype(np.float64)
    Wh = bf16_rn(W).astype(np.float64); Wl = bf16_rn(W - Wh).astype(np.float64)
    out = {}
    out["bf16 (1 issue)"] = Ah @ Wh.T
    out["bf16x3 (3 issues)"] = Ah @ Wh.T + Ah @ Wl.T + Al @ Wh.T
    out["bf16 + 2 x mxfp8 corrections (2 issue-equivalents)"] = (
        Ah @ Wh.T + e4m3_block(Ah) @ e4m3_block(Wl).T + e4m3_block(Al) @ e4m3_block(Wh).T)
    Ah16 = A.astype(np.float16).astype(np.float64); Al16 = A - Ah16
    Wh16 = W.astype(np.float16).astype(np.float64); Wl16 = W - Wh16
    out["fp16 (1 issue)"] = Ah16 @ Wh16.T
    out["fp16 + 2 x mxfp8 corrections (2 issue-equivalents)"] = (
        Ah16 @ Wh16.T + e4m3_block(Ah16) @ e4m3_block(Wl16).T + e4m3_block(Al16) @ e4m3_block(Wh16).T)
    out["fp16 + 1 x mxfp8 correction of A only (1.5)"] = Ah16 @ Wh16.T + e4m3_block(Al16) @ e4m3_block(Wh16).T
    zmax = np.abs(z).max()
    for k, v in out.items():
        err = np.abs(v - z)
        p = np.exp(z - z.max(axis=1, keepdims=True)); p /= p.sum(axis=1, keepdims=True)
        pv = np.exp(v - v.max(axis=1, keepdims=True)); pv /= pv.sum(axis=1, keepdims=True)
        print(f"{k:55s} max|dz|/max|z| = {err.max() / zmax:.2e}   rms = {np.sqrt((err ** 2).mean()) / zmax:.2e}   "
              f"max|dp|/max p = {np.abs(pv - p).max() / p.max():.2e}")


if __name__ == "__main__":
    print("Kaiming-uniform W (the bench's init):")
    study()
    print("W scaled x8 (peaky softmax, closer to a trained joint):")
    study(trained_scale=8.0)
