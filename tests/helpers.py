"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

from oracle import c_oracle
from tests.golden import gen_inputs as gi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances (BASELINE.json north_star): codes/masks bit-exact except audited fp32 near-ties;
# z_q within 1e-5 relative.  "Relative" is taken against the per-frame max-abs of the reference
# tensor (an element-wise relative test is meaningless for values that cancel to ~0; the
# reference itself is not bit-stable across batch shapes, SURVEY.md section 7 hard part 4).
ZQ_RTOL = 1e-5
# Distance gap below which a differing code is an audited near-tie.  The oracle's z_e is a binary64 dot product rounded
# once (<= 6e-8 relative); the kernels' z_e carries up to 1.3e-6 relative error (3xTF32 GEMM, profiles/r1_micro_tc3x.txt;
# an fp32 FMA chain: 4e-6), which moves a cosine distance 2 - 2 e.c by at most ~2.6e-6.  A flip with a larger gap is a bug.
NEAR_TIE_EPS = 3e-6
# Excused near-tie frames allowed per comparison: every GPU run so far printed 0 (profiles/r2_gpu_tests_tail.txt); the bound
# is one frame or 1e-4 of the frames, whichever is larger.
MAX_EXCUSED_FRAC = 1e-4


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def oracle_weights_for(case):
    import torch

    sd = gi.torch_state_dict(gi.make_state_dict(case["seed"], case["Nq"], case["D"], case["K"]))
    return c_oracle.OracleWeights.from_state_dict(sd)


def folded_numpy(case):
    w = oracle_weights_for(case)
    return w


def latents_for(case):
    return gi.make_latents(case["seed"] + 1000, case["B"], case["D"], case["T"], case["sigma"])


def rel_err_per_frame(a, ref, frame_axis=-1):
    """max_d |a-ref| / max_d |ref| per (b,t) frame for tensors shaped [B, ..., T]."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    red = tuple(range(1, a.ndim - 1))
    num = np.abs(a - ref).max(axis=red)
    den = np.abs(ref).max(axis=red)
    return num / np.maximum(den, 1e-30)


def assert_close_frames(a, ref, rtol=ZQ_RTOL, skip=None, what="tensor"):
    err = rel_err_per_frame(a, ref)
    if skip is not None:
        err = np.where(skip, 0.0, err)
    worst = err.max() if err.size else 0.0
    assert worst <= rtol, f"{what}: worst per-frame relative error {worst:.3e} > {rtol:.1e}"
    return worst


def assert_codes_match(w, oracle_out, codes_other, max_excused_frac=MAX_EXCUSED_FRAC, what="codes"):
    """Exact match, except frames whose first differing stage is an audited near-tie."""
    bad, excused, excused_mask = c_oracle.audit_code_mismatches(w, oracle_out, codes_other, eps=NEAR_TIE_EPS)
    nframes = max(excused_mask.size, 1)
    assert bad == 0, f"{what}: {bad} frame(s) differ from the oracle without a near-tie"
    assert excused <= max(1, int(max_excused_frac * nframes)), f"{what}: {excused} excused near-tie frames of {nframes}"
    return excused, excused_mask
