"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of oracle/rvq_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (vrvq_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

from . import build_oracle

CD = 8


class _W(C.Structure):
    _fields_ = [
        ("n_codebooks", C.c_int),
        ("input_dim", C.c_int),
        ("codebook_size", C.c_int),
        ("w_in", C.c_void_p),
        ("b_in", C.c_void_p),
        ("w_out", C.c_void_p),
        ("b_out", C.c_void_p),
        ("cb_raw", C.c_void_p),
        ("cb_nrm", C.c_void_p),
        ("c2", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.OUT
        if build_oracle.needs_build():
            path = build_oracle.build()
        _lib = C.CDLL(path)
        _lib.vrvq_oracle_encode.restype = C.c_int
        _lib.vrvq_oracle_from_codes.restype = C.c_int
        _lib.vrvq_oracle_from_latents.restype = C.c_int
        _lib.vrvq_oracle_num_threads.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_num_threads(n: int):
    lib().vrvq_oracle_set_num_threads(C.c_int(int(n)))


def num_threads() -> int:
    return int(lib().vrvq_oracle_num_threads())


class OracleWeights:
    """Folded RVQ weights.

    w_in [Nq,8,D], b_in [Nq,8], w_out [Nq,D,8], b_out [Nq,D], codebook [Nq,K,8]; all float32.
    The fold W = torch._weight_norm(v, g, 0) (models/layers.py:17-18) is done by the caller with torch.
    """

    def __init__(self, w_in, b_in, w_out, b_out, codebook):
        self.w_in = _f32(w_in)
        self.b_in = _f32(b_in)
        self.w_out = _f32(w_out)
        self.b_out = _f32(b_out)
        self.cb_raw = _f32(codebook)
        self.Nq, cd, self.D = self.w_in.shape
        assert cd == CD, "codebook_dim must be 8"
        self.K = self.cb_raw.shape[1]
        assert self.w_out.shape == (self.Nq, self.D, CD)
        assert self.b_in.shape == (self.Nq, CD) and self.b_out.shape == (self.Nq, self.D)
        assert self.cb_raw.shape == (self.Nq, self.K, CD)
        self.cb_nrm = np.empty_like(self.cb_raw)
        self.c2 = np.empty((self.Nq, self.K), np.float32)
        for s in range(self.Nq):
            lib().vrvq_oracle_prepare_codebook(_p(self.cb_raw[s]), C.c_int(self.K), _p(self.cb_nrm[s]), _p(self.c2[s]))
        self._c = _W(self.Nq, self.D, self.K, _p(self.w_in), _p(self.b_in), _p(self.w_out), _p(self.b_out),
                     _p(self.cb_raw), _p(self.cb_nrm), _p(self.c2))

    @classmethod
    def from_state_dict(cls, sd, prefix=""):
        """Fold a reference-layout state dict (quantizers.{i}.in_proj.weight_g/_v/bias, ...) with torch on the CPU."""
        import torch

        def get(name):
            return sd[prefix + name].detach().to("cpu", torch.float32)

        n = 0
        while f"{prefix}quantizers.{n}.codebook.weight" in sd:
            n += 1
        w_in, b_in, w_out, b_out, cb = [], [], [], [], []
        for i in range(n):
            q = f"quantizers.{i}."
            w_in.append(torch._weight_norm(get(q + "in_proj.weight_v"), get(q + "in_proj.weight_g"), 0)[:, :, 0])
            b_in.append(get(q + "in_proj.bias"))
            w_out.append(torch._weight_norm(get(q + "out_proj.weight_v"), get(q + "out_proj.weight_g"), 0)[:, :, 0])
            b_out.append(get(q + "out_proj.bias"))
            cb.append(get(q + "codebook.weight"))
        st = lambda xs: torch.stack(xs).numpy()
        return cls(st(w_in), st(b_in), st(w_out), st(b_out), st(cb))

    @property
    def cptr(self):
        return C.byref(self._c)


def encode(w: OracleWeights, z, n_quantizers=None, imp_map=None, level=None, want_z_q_is=True, want_gap=True):
    """Eval forward of (VBR)ResidualVectorQuantize.  See vrvq_oracle_encode in rvq_oracle.c.

    z [B,D,T]; imp_map [B,1,T] or [B,T] (VBR masking) or None (CBR masking);
    level: python scalar or array broadcastable to [B]; n_quantizers: stages to run (None = all).
    """
    z = _f32(z)
    B, D, T = z.shape
    assert D == w.D
    n_run = w.Nq if n_quantizers is None else int(n_quantizers)
    lv = None
    stride = 0
    if imp_map is not None:
        imp_map = _f32(imp_map).reshape(B, T)
        lv = _f32(np.asarray(level, dtype=np.float32).reshape(-1))
        assert lv.size in (1, B)
        stride = 0 if lv.size == 1 else 1
    out = {
        "codes": np.zeros((B, n_run, T), np.int64),
        "z_q": np.zeros((B, D, T), np.float32),
        "z_q_is": np.zeros((B, n_run, D, T), np.float32) if want_z_q_is else None,
        "latents": np.zeros((B, CD * n_run, T), np.float32),
        "mask": np.zeros((B, n_run, T), np.float32),
        "loss_pf": np.zeros((B, n_run, T), np.float32),
        "kept": np.zeros((max(n_run, 1),), np.int64),
        "gap": np.zeros((B, n_run, T), np.float32) if want_gap else None,
    }
    loss_sum = C.c_double(0.0)
    rc = lib().vrvq_oracle_encode(
        w.cptr, _p(z), C.c_int(B), C.c_int(T), C.c_int(n_run), _p(imp_map), _p(lv), C.c_int(stride),
        _p(out["codes"]), _p(out["z_q"]), _p(out["z_q_is"]), _p(out["latents"]), _p(out["mask"]),
        _p(out["loss_pf"]), C.byref(loss_sum), _p(out["kept"]), _p(out["gap"]))
    if rc != 0:
        raise RuntimeError(f"vrvq_oracle_encode failed: {rc}")
    out["kept"] = out["kept"][:n_run]
    out["loss_masked_sum"] = loss_sum.value
    # quantize.py:198-199 (CBR) and :422-423 (VBR) both reduce to sum(mask*loss_pf)/(B*T)
    out["commitment_loss"] = loss_sum.value / max(B * T, 1)
    out["codebook_loss"] = out["commitment_loss"]
    return out


def from_codes(w: OracleWeights, codes, want_z_q_is=False):
    codes = np.ascontiguousarray(np.asarray(codes, dtype=np.int64))
    B, n, T = codes.shape
    z_q = np.zeros((B, w.D, T), np.float32)
    z_p = np.zeros((B, CD * n, T), np.float32)
    z_q_is = np.zeros((B, n, w.D, T), np.float32) if want_z_q_is else None
    rc = lib().vrvq_oracle_from_codes(w.cptr, _p(codes), C.c_int(B), C.c_int(T), C.c_int(n), _p(z_q), _p(z_p), _p(z_q_is))
    if rc != 0:
        raise RuntimeError(f"vrvq_oracle_from_codes failed: {rc}")
    return z_q, z_p, z_q_is


def from_latents(w: OracleWeights, latents):
    latents = _f32(latents)
    B, c, T = latents.shape
    n = c // CD
    z_q = np.zeros((B, w.D, T), np.float32)
    z_p = np.zeros((B, CD * n, T), np.float32)
    codes = np.zeros((B, n, T), np.int64)
    rc = lib().vrvq_oracle_from_latents(w.cptr, _p(latents), C.c_int(B), C.c_int(T), C.c_int(n), _p(z_q), _p(z_p), _p(codes))
    if rc != 0:
        raise RuntimeError(f"vrvq_oracle_from_latents failed: {rc}")
    return z_q, z_p, codes


def generate_mask_hard(x, nq):
    x = _f32(x)
    B, T = x.shape[0], x.shape[-1]
    m = np.zeros((B, nq, T), np.float32)
    lib().vrvq_oracle_mask_hard(_p(x.reshape(B, T)), C.c_int(B), C.c_int(T), C.c_int(nq), _p(m))
    return m


def cal_bpf_from_mask(mask, bits_per_codebook):
    mask = _f32(mask)
    B, nq, T = mask.shape
    s = np.zeros((nq,), np.float64)
    lib().vrvq_oracle_mask_sum(_p(mask), C.c_int(B), C.c_int(T), C.c_int(nq), _p(s))
    return float(np.dot(s, np.asarray(bits_per_codebook, dtype=np.float64)) / (B * T))


def audit_code_mismatches(w: OracleWeights, oracle_out, codes_other, eps=1e-5):
    """First-divergence near-tie audit (hard part 4).

    A frame may differ from the oracle from stage s onward only if, at its first differing stage s,
    the other implementation's code is a near-tie under the oracle's own z_e: its distance exceeds
    the oracle's best distance by less than `eps`.  Everything after s in that frame is excused
    because the residuals fork.  Returns (n_bad_frames, n_excused_frames, excused_mask[B,T]).
    """
    co = oracle_out["codes"]
    codes_other = np.asarray(codes_other)
    assert co.shape == codes_other.shape
    B, n, T = co.shape
    diff = co != codes_other
    excused = np.zeros((B, T), bool)
    bad = 0
    bs, ts = np.nonzero(diff.any(axis=1))
    for b, t in zip(bs, ts):
        s = int(np.argmax(diff[b, :, t]))
        z_e = oracle_out["latents"][b, CD * s:CD * s + CD, t].astype(np.float64)
        nrm = max(np.sqrt((z_e * z_e).sum()), 1e-12)
        e = z_e / nrm
        cb = w.cb_nrm[s].astype(np.float64)
        d_or = ((e - cb[co[b, s, t]]) ** 2).sum()
        d_ot = ((e - cb[codes_other[b, s, t]]) ** 2).sum()
        if abs(d_ot - d_or) < eps:
            excused[b, t] = True
        else:
            bad += 1
    return bad, int(excused.sum()), excused
