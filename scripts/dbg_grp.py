import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrvq_b200 import ops
from oracle import c_oracle
from tests import helpers as H
from tests.golden import gen_inputs as gi
Nq, D, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
vbr = len(sys.argv) > 5 and sys.argv[5] == "vbr"
os.environ["VRVQ_ENCODE_IMPL"] = "tc"
sd = gi.torch_state_dict(gi.make_state_dict(100 + Nq, Nq, D))
w = c_oracle.OracleWeights.from_state_dict(sd)
pw = ops.PackedWeights.from_state_dict(sd, "cuda")
print(ops.encode_launch_info(pw, B, T, Nq, "cuda"))
z_np = gi.make_latents(7, B, D, T, 1.0)
imp_np = gi.make_imp_map(8, B, T) if vbr else None
o = c_oracle.encode(w, z_np, None, imp_np, 0.6 if vbr else None, want_z_q_is=False)
out = ops.rvq_encode(pw, torch.from_numpy(z_np).cuda(), None, torch.from_numpy(imp_np).cuda() if vbr else None, 0.6 if vbr else None)
torch.cuda.synchronize()
codes = out.codes.cpu().numpy()
bad, excused, skip = c_oracle.audit_code_mismatches(w, o, codes, eps=3e-6)
agree = (codes == o["codes"]).all(axis=2).mean(axis=0) if False else (codes == o["codes"]).mean(axis=(0, 2))
print("per-stage agreement:", np.round(agree, 4))
print("bad", bad, "excused", excused)
err = H.rel_err_per_frame(out.z_q.cpu().numpy(), o["z_q"])
err = np.where(skip, 0, err)
print("z_q max rel err", err.max(), "latents err", np.where(skip, 0, H.rel_err_per_frame(out.latents.cpu().numpy(), o["latents"])).max())
print("mask eq", np.array_equal(out.mask.cpu().numpy(), o["mask"]), "kept eq", np.array_equal(out.kept.cpu().numpy(), o["kept"]))
