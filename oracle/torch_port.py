"""TEST / BASELINE INFRASTRUCTURE ONLY -- an eager-PyTorch CPU port of the reference RVQ forward.

The reference *is* eager PyTorch (oneDNN 1x1 convs, MKL sgemm for the distance matrix, ATen elementwise
kernels), and its checkout does not travel to the GPU box.  This port issues the same ATen op sequence
per stage (models/quantize.py:66-77, 87-103, 353-365, 389-395, 420-423) on folded weights, so timing it
on the box's host cores is timing the reference's CPU path.  bench.py's cpu_baseline leg and
`--impl reference` arm are the only callers besides tests/; the product path never imports it.
tests/test_oracle_vs_reference.py checks it bit-for-bit against the live reference in the build container.
"""
import torch
import torch.nn.functional as F


class TorchPortWeights:
    def __init__(self, sd, prefix=""):
        n = 0
        while f"{prefix}quantizers.{n}.codebook.weight" in sd:
            n += 1
        self.n = n
        g = lambda k: sd[prefix + k].detach().to("cpu", torch.float32)
        self.v_in = [g(f"quantizers.{i}.in_proj.weight_v") for i in range(n)]
        self.g_in = [g(f"quantizers.{i}.in_proj.weight_g") for i in range(n)]
        self.b_in = [g(f"quantizers.{i}.in_proj.bias") for i in range(n)]
        self.v_out = [g(f"quantizers.{i}.out_proj.weight_v") for i in range(n)]
        self.g_out = [g(f"quantizers.{i}.out_proj.weight_g") for i in range(n)]
        self.b_out = [g(f"quantizers.{i}.out_proj.bias") for i in range(n)]
        self.cb = [g(f"quantizers.{i}.codebook.weight") for i in range(n)]


@torch.no_grad()
def rvq_forward(w: TorchPortWeights, z, n_quantizers=None, imp_map=None, level=None, keep_z_q_is=True):
    """Eval forward.  imp_map given -> VBR masking (all stages run); else CBR with early exit."""
    B, D, T = z.shape
    Nq = w.n
    n_run = Nq if (n_quantizers is None or imp_map is not None) else min(int(n_quantizers), Nq)
    residual = z
    z_q_is, codes, latents, losses = [], [], [], []
    for i in range(n_run):
        # weight_norm pre-hook recomputes the effective weight every forward (models/layers.py:17-18)
        w_in = torch._weight_norm(w.v_in[i], w.g_in[i], 0)
        z_e = F.conv1d(residual, w_in, w.b_in[i])                               # quantize.py:66
        enc = z_e.transpose(1, 2).reshape(B * T, -1)                             # :88
        enc_n = F.normalize(enc)                                                 # :92
        cb_n = F.normalize(w.cb[i])                                              # :93
        dist = enc_n.pow(2).sum(1, keepdim=True) - 2 * enc_n @ cb_n.t() + cb_n.pow(2).sum(1, keepdim=True).t()  # :96-100
        idx = (-dist).max(1)[1].view(B, T)                                       # :101
        z_c = F.embedding(idx, w.cb[i]).transpose(1, 2)                          # :81-85,102
        loss = F.mse_loss(z_e, z_c, reduction="none").mean(1)                    # :69-71 (per frame)
        z_st = z_e + (z_c - z_e)                                                 # :73-75
        w_out = torch._weight_norm(w.v_out[i], w.g_out[i], 0)
        z_q_i = F.conv1d(z_st, w_out, w.b_out[i])                                # :77
        residual = residual - z_q_i                                              # :195 / :360
        z_q_is.append(z_q_i); codes.append(idx); latents.append(z_e); losses.append(loss)
    if imp_map is not None:
        x = imp_map * level * Nq                                                 # :389
        ks = torch.arange(Nq, dtype=torch.float32).view(1, Nq, 1)
        mask = torch.where(x - ks >= 0, torch.ones(()), torch.zeros(())).float().expand(B, Nq, T).contiguous()  # utils.py:55-61
    else:
        mask = torch.ones(B, n_run, T)
    stack = torch.stack(z_q_is, dim=1)                                           # :420
    z_q = torch.sum(stack * mask[:, :n_run, None, :], dim=1)                     # :421
    loss = (torch.stack(losses, dim=1) * mask[:, :n_run]).sum(dim=1).mean()      # :422-423
    return {"z_q": z_q, "z_q_is": stack if keep_z_q_is else None, "codes": torch.stack(codes, dim=1),
            "latents": torch.cat(latents, dim=1), "commitment_loss": loss, "codebook_loss": loss, "mask_imp": mask}
