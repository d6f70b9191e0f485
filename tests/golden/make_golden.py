"""Generates tests/golden/*.npz by running the UNMODIFIED reference (lixinghe1999/VRVQ) on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference quantizer classes are imported as they lie (oracle/ref_import.py); the only
substitution is the importance subnet of the VBR model, an *upstream producer* of the hot path
(SURVEY.md section 2 row 5), which is replaced by a module returning a seeded importance map so that
every mask edge is exercised (a random-init subnet emits a near-constant 0.5).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402
from tests.golden import gen_inputs as gi  # noqa: E402

ROW_STEP = 16  # z_q_is is stored for every 16th channel only (fixture size)


class FixedImportance(torch.nn.Module):
    def __init__(self, imp):
        super().__init__()
        self.imp = imp

    def forward(self, feat):
        return self.imp


def build_reference(ref, case):
    sd = gi.torch_state_dict(gi.make_state_dict(case["seed"], case["Nq"], case["D"], case["K"]))
    if case["kind"] == "vbr":
        m = ref.VBRResidualVectorQuantize(input_dim=case["D"], n_codebooks=case["Nq"], codebook_size=case["K"],
                                          codebook_dim=8, level_min=0.125, level_max=6.0, imp2mask_alpha=2.0)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("imp_subnet.") for k in missing)
    else:
        m = ref.ResidualVectorQuantize(input_dim=case["D"], n_codebooks=case["Nq"], codebook_size=case["K"], codebook_dim=8)
        m.load_state_dict(sd, strict=True)
    return m.eval()


def run_case(ref, name, case):
    torch.set_num_threads(8)
    m = build_reference(ref, case)
    z = torch.from_numpy(gi.make_latents(case["seed"] + 1000, case["B"], case["D"], case["T"], case["sigma"]))
    out = {}
    with torch.no_grad():
        if case["kind"] == "vbr":
            imp = torch.from_numpy(gi.make_imp_map(case["imp_seed"], case["B"], case["T"]))
            m.imp_subnet = FixedImportance(imp)
            for li, level in enumerate(case["levels"]):
                lv = level if not isinstance(level, list) else torch.tensor(level, dtype=torch.float32).view(-1, 1, 1)
                r = m(z, n_quantizers=None, feat_enc=z, level=lv)
                if li == 0:
                    out["codes"] = r["codes"].numpy()
                    out["latents"] = r["latents"].numpy()
                    out["z_q_is_sub"] = r["z_q_is"][:, :, ::ROW_STEP, :].contiguous().numpy()
                else:
                    assert np.array_equal(out["codes"], r["codes"].numpy())
                out[f"z_q_{li}"] = r["z_q"].numpy()
                out[f"mask_{li}"] = r["mask_imp"].numpy()
                out[f"commitment_loss_{li}"] = np.float32(r["commitment_loss"].item())
                out[f"codebook_loss_{li}"] = np.float32(r["codebook_loss"].item())
                out[f"bpf_{li}"] = np.float64(ref.cal_bpf_from_mask(r["mask_imp"], [10] * case["Nq"]))
                # README.md:75-79 / scripts/inference.py:95-100 : the re-mask recipe must agree with the fused mask
                lvl_scaled = lv * case["Nq"]
                mh = ref.generate_mask_hard(imp * lvl_scaled, nq=case["Nq"])
                assert torch.equal(mh, r["mask_imp"]), "generate_mask_ste forward value != generate_mask_hard"
        else:
            for qi, nq in enumerate(case["n_quantizers"]):
                r = m(z, n_quantizers=nq)
                out[f"codes_{qi}"] = r["codes"].numpy()
                out[f"latents_{qi}"] = r["latents"].numpy()
                out[f"z_q_{qi}"] = r["z_q"].numpy()
                out[f"commitment_loss_{qi}"] = np.float32(r["commitment_loss"].item())
                out[f"codebook_loss_{qi}"] = np.float32(r["codebook_loss"].item())
                if nq is None:
                    # decode side (quantize.py:217-285)
                    zq_c, zp_c, _, zqis_c = m.from_codes(r["codes"], return_z_q_is=True)
                    out["from_codes_z_q"] = zq_c.numpy()
                    out["from_codes_z_p"] = zp_c.numpy()
                    out["from_codes_z_q_is_sub"] = zqis_c[:, :, ::ROW_STEP, :].contiguous().numpy()
                    zq_l, zp_l, codes_l = m.from_latents(r["latents"])
                    out["from_latents_z_q"] = zq_l.numpy()
                    out["from_latents_codes"] = codes_l.numpy()
                    assert np.array_equal(zp_l.numpy(), zp_c.numpy())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def mask_cases(ref):
    """generate_mask_hard / cal_bpf_from_mask known answers (models/utils.py:55-73)."""
    rng = np.random.Generator(np.random.PCG64(99))
    x = (rng.uniform(-1.0, 9.5, (4, 1, 57))).astype(np.float32)
    x[0, 0, :10] = np.arange(10, dtype=np.float32) - 1.0  # exact integer thresholds
    xi = rng.integers(1, 9, (3, 1, 5))  # the int64 call site quantize.py:412
    out = {"x": x, "xi": xi}
    for nq in (8, 28):
        mk = ref.generate_mask_hard(torch.from_numpy(x), nq)
        out[f"mask_nq{nq}"] = mk.numpy()
        out[f"bpf_nq{nq}"] = np.float64(ref.cal_bpf_from_mask(mk, [10] * nq))
        out[f"bpf_ragged_nq{nq}"] = np.float64(ref.cal_bpf_from_mask(mk, list(range(1, nq + 1))))
    out["mask_int"] = ref.generate_mask_hard(torch.from_numpy(xi), 8).numpy()
    path = os.path.join(HERE, "mask_utils.npz")
    np.savez_compressed(path, **out)
    print("mask_utils: wrote", path)


def subnet_cases():
    """imp_map of the UNMODIFIED reference ImportanceSubnet (models/importance_subnet.py) on seeded weights and features."""
    if ref_import.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REF_ROOT)
    from models.importance_subnet import ImportanceSubnet

    torch.set_num_threads(8)
    for name, c in gi.SUBNET_CASES.items():
        m = ImportanceSubnet(d_input=c["d_input"], d_feat=c["d_feat"], intermediate_channels=list(c["widths"]), out_channels=1).eval()
        m.load_state_dict(gi.torch_state_dict(gi.make_subnet_state_dict(c["seed"], c["d_input"], c["d_feat"], c["widths"])), strict=True)
        x = torch.from_numpy(gi.make_latents(c["seed"] + 1000, c["B"], c["d_input"], c["T"], c["sigma"]))
        with torch.no_grad():
            imp = m(x)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, imp_map=imp.numpy())
        print(f"{name}: wrote {path}, imp_map in [{imp.min():.3f}, {imp.max():.3f}]")


def dac_model_kwargs(c):
    kw = dict(n_codebooks=c["n_codebooks"], model_type=c["model_type"])
    if c["model_type"] == "VBR":
        kw.update(level_min=0.125, level_max=6.0, imp2mask_alpha=2.0)
    return kw


def dac_cases(ref):
    """DAC_VRVQ.encode(audio, n_quantizers, level) of the UNMODIFIED reference (models/dac_vrvq.py:176-213) on seeded weights
    and audio: the API BASELINE.json's north_star names.  Stored: codes, mask, imp_map, latents, losses in full; z (the
    encoder output the quantizer saw), feat and z_q for every 16th channel."""
    torch.set_num_threads(8)
    for name, c in gi.DAC_CASES.items():
        m = ref.DAC_VRVQ(**dac_model_kwargs(c)).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("decoder.")}
        full = dict(m.state_dict())
        full.update(gi.torch_state_dict(gi.make_dac_state_dict(c["seed"], shapes)))
        m.load_state_dict(full, strict=True)
        x = torch.from_numpy(gi.make_audio(c["seed"] + 1, c["B"], c["samples"]))
        with torch.no_grad():
            xp = m.preprocess(x, 44100)
            z, feat = m.encoder(xp, return_feat=True)
            r = m.encode(xp, c["n_quantizers"], c["level"]) if c["model_type"] == "VBR" else m.encode(xp, c["n_quantizers"])
        out = {"codes": r["codes"].numpy(), "latents": r["latents"].numpy(), "z_sub": z[:, ::ROW_STEP].contiguous().numpy(),
               "feat_sub": feat[:, ::ROW_STEP].contiguous().numpy(), "z_q_sub": r["z_q"][:, ::ROW_STEP].contiguous().numpy(),
               "commitment_loss": np.float32(r["commitment_loss"].item()), "codebook_loss": np.float32(r["codebook_loss"].item()),
               "key_order": np.array(list(shapes.keys()))}
        if c["model_type"] == "VBR":
            out["imp_map"] = r["imp_map"].numpy()
            out["mask_imp"] = r["mask_imp"].numpy()
            out["bpf"] = np.float64(ref.cal_bpf_from_mask(r["mask_imp"], [10] * c["n_codebooks"]))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {path} ({os.path.getsize(path)/1e6:.2f} MB), T={z.shape[-1]}, z std {z.std():.3f}")


def main():
    ref = ref_import.load()
    if "--subnet-only" in sys.argv:
        return subnet_cases()
    if "--dac-only" in sys.argv:
        return dac_cases(ref)
    for name, case in gi.CASES.items():
        run_case(ref, name, case)
    mask_cases(ref)
    subnet_cases()
    dac_cases(ref)


if __name__ == "__main__":
    main()
