"""TEST / BASELINE INFRASTRUCTURE ONLY -- an eager-PyTorch CPU port of the reference RVQ forward.

The reference *is* eager PyTorch (oneDNN 1x1 convs, MKL sgemm for the distance matrix, ATen elementwise
kernels), and its checkout does not travel to the GPU box.  This port issues the same ATen op sequence
per stage (models/quantize.py:66-77, 87-103, 353-365, 389-395, 420-423) on folded weights, so timing it
on the box's host cores is timing the reference's CPU path.  bench.py's cpu_baseline leg and
`--impl reference` arm are the only callers besides tests/; the product path never imports it.
tests/test_oracle_vs_reference.py::test_torch_port_is_bit_identical_to_the_live_reference checks it bit-for-bit (codes, z_q,
z_q_is, latents, mask, losses; VBR level sweep and CBR with early exit, B in {4, 16} x T = 862) against the live reference in the
build container.
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
    cbr = imp_map is None
    z_q_cbr, loss_cbr, cbl_cbr = 0, 0, 0
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
        codes.append(idx); latents.append(z_e)
        if cbr:                                                                  # quantize.py:187-199 (eval: mask is all True)
            mask_b = torch.full((B,), fill_value=i) < n_run
            z_q_cbr = z_q_cbr + z_q_i * mask_b[:, None, None]
            loss_cbr = loss_cbr + (F.mse_loss(z_e, z_c, reduction="none").mean([1, 2]) * mask_b).mean()
            # the two losses have equal terms but not equal bits: the operand order decides the memory layout of the
            # elementwise result (z_c is a transposed gather) and with it the order of the mean's reduction
            cbl_cbr = cbl_cbr + (F.mse_loss(z_c, z_e, reduction="none").mean([1, 2]) * mask_b).mean()
        else:
            z_q_is.append(z_q_i); losses.append(loss)
    if cbr:
        return {"z_q": z_q_cbr, "z_q_is": None, "codes": torch.stack(codes, dim=1), "latents": torch.cat(latents, dim=1),
                "commitment_loss": loss_cbr, "codebook_loss": cbl_cbr, "mask_imp": torch.ones(B, n_run, T)}
    x = imp_map * level * Nq                                                     # :389
    ks = torch.arange(Nq, dtype=torch.float32).view(1, Nq, 1)
    mask = torch.where(x - ks >= 0, torch.ones(()), torch.zeros(())).float().expand(B, Nq, T).contiguous()  # utils.py:55-61
    stack = torch.stack(z_q_is, dim=1)                                           # :420
    z_q = torch.sum(stack * mask[:, :n_run, None, :], dim=1)                     # :421
    loss = (torch.stack(losses, dim=1) * mask[:, :n_run]).sum(dim=1).mean()      # :422-423
    return {"z_q": z_q, "z_q_is": stack if keep_z_q_is else None, "codes": torch.stack(codes, dim=1),
            "latents": torch.cat(latents, dim=1), "commitment_loss": loss, "codebook_loss": loss, "mask_imp": mask}
