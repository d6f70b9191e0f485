"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (vrvq_b200/).

numpy restatement of the reference's importance subnet (models/importance_subnet.py:6-45), the upstream producer of
the importance map.  The arithmetic of the reference lives in ATen/oneDNN
(`F.conv1d`, `torch.sin`); this port states the same algorithm in plain array operations:

  fold     w = v * (g / ||v||_2 over dims 1,2)                 torch._weight_norm(v, g, 0)   models/layers.py:17-18
  snake    s = x + 1/(alpha + 1e-9) * sin(alpha * x)^2         models/layers.py:25-31
  conv     y[co,t] = b[co] + sum_{ci,k} w[co,ci,k] s[ci,t+k-1] zero padded, kernel 3, padding 1   importance_subnet.py:18-34
  sigmoid  1 / (1 + exp(-y)) after the last block              importance_subnet.py:43

`dtype=np.float32` follows the reference's precision (reduction order is oneDNN's and unspecified, so agreement is a
tolerance, not bit equality); `dtype=np.float64` gives the exact value both the reference and the CUDA kernel are
measured against.  Pinned against the live reference in tests/test_oracle_vs_reference.py (build container) and against
the committed fixture tests/golden/subnet_*.npz (generated from the reference by tests/golden/make_golden.py).
"""
import numpy as np


def block_keys(prefix, n_blocks):
    """State-dict prefixes of the (snake, conv) pairs in forward order (importance_subnet.py:18-34)."""
    return [(f"{prefix}in_block.0.", f"{prefix}in_block.1.")] + [(f"{prefix}blocks.{i}.0.", f"{prefix}blocks.{i}.1.") for i in range(n_blocks)]


def fold(v, g, dtype):
    v = np.asarray(v, dtype)
    g = np.asarray(g, dtype)
    n = np.sqrt((v * v).sum(axis=(1, 2), keepdims=True))
    return v * (g / n)


def snake(x, alpha, dtype):
    a = np.asarray(alpha, dtype).reshape(1, -1, 1)
    s = np.sin(a * x)
    return x + (dtype(1.0) / (a + dtype(1e-9))) * (s * s)


def conv3(s, w, b):
    B, Cin, T = s.shape
    p = np.zeros((B, Cin, T + 2), s.dtype)
    p[:, :, 1:T + 1] = s
    y = np.zeros((B, w.shape[0], T), s.dtype)
    for k in range(3):
        y += np.einsum("oc,bct->bot", w[:, :, k], p[:, :, k:k + T], optimize=True)
    return y + b.reshape(1, -1, 1)


def importance_subnet(sd, x, prefix="", dtype=np.float32):
    """sd: mapping of numpy arrays with the reference's keys; x [B, d_input, T] -> imp_map [B, 1, T]."""
    n = 0
    while f"{prefix}blocks.{n}.1.weight_v" in sd:
        n += 1
    x = np.asarray(x, dtype)
    for sk, ck in block_keys(prefix, n):
        w = fold(sd[ck + "weight_v"], sd[ck + "weight_g"], dtype)
        x = conv3(snake(x, sd[sk + "alpha"], dtype), w, np.asarray(sd[ck + "bias"], dtype))
    return (dtype(1.0) / (dtype(1.0) + np.exp(-x))).astype(dtype)
