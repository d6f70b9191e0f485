"""Randomised stress of the tensor-core encode kernel against the CUDA-core kernel (shapes, multi-tile CTAs, optional outputs)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

weights = {}
def get_w(D, Nq):
    if (D, Nq) not in weights:
        weights[(D, Nq)] = ops.PackedWeights.from_state_dict(gi.torch_state_dict(gi.make_state_dict(900 + Nq, Nq, D)), "cuda")
    return weights[(D, Nq)]

def run(impl, fn):
    os.environ["VRVQ_ENCODE_IMPL"] = impl
    try:
        return fn()
    finally:
        os.environ.pop("VRVQ_ENCODE_IMPL", None)

def run_cases(seed, n_cases, verbose=True):
  random.seed(seed)
  worst = 0.0
  for it in range(n_cases):
      D = random.choice([1024, 1024, 512, 256])
      Nq = random.choice([8, 8, 8, 5, 3, 1, 9, 12, 28])  # > 8: the grouped instantiation (no z_q_is)
      n_run = Nq if random.random() < 0.7 else random.randint(1, Nq)
      B = random.choice([1, 2, 3, 7, 16, 33])
      T = random.choice([1, 7, 8, 9, 95, 96, 97, 120, 121, 431, 862, 1000, 2049])
      if B * T * D > 40e6:
          B = max(1, int(40e6 // (T * D)))
      vbr = random.random() < 0.6 and n_run == Nq
      zqis = random.random() < 0.5 and Nq <= 8
      if Nq > 8 and B * T > 20000:
          B = max(1, 20000 // T)
      pw = get_w(D, Nq)
      z = torch.randn(B, D, T, device="cuda")
      imp = torch.rand(B, 1, T, device="cuda") if vbr else None
      lvl = random.choice([0.25, 0.5, 1.0, 2.0]) if vbr else None
      call = lambda: ops.rvq_encode(pw, z, n_run, imp, lvl, want_z_q_is=zqis, want_loss_pf=True)
      a, c = run("tc", call), run("cuda", call)
      torch.cuda.synchronize()
      same = (a.codes == c.codes).all(dim=1)
      frac = same.float().mean().item()
      assert frac >= 0.97, (it, D, Nq, n_run, B, T, vbr, zqis, frac)
      assert torch.equal(a.mask, c.mask) and torch.equal(a.kept, c.kept), (it, "mask/kept")
      def rel(x, y):
          num = (x - y).abs().amax(dim=tuple(range(1, x.dim() - 1)))
          den = y.abs().amax(dim=tuple(range(1, y.dim() - 1))).clamp_min(1e-30)
          return torch.where(same, num / den, torch.zeros_like(num)).max().item()
      e = rel(a.z_q, c.z_q)
      if zqis:
          e = max(e, rel(a.z_q_is.reshape(B, -1, T), c.z_q_is.reshape(B, -1, T)))
      e = max(e, rel(a.latents, c.latents))
      worst = max(worst, e)
      assert e <= 1e-5, (it, D, Nq, n_run, B, T, vbr, zqis, e)
      if verbose: print(f"case {it}: D={D} Nq={Nq} n_run={n_run} B={B} T={T} vbr={vbr} zqis={zqis}: same-code frames {frac:.4f}, max rel err {e:.2e}", flush=True)
  return worst


if __name__ == "__main__":
    w = run_cases(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 30)
    print("stress ok, worst rel err", w)
