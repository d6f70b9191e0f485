"""Secondary measurements (not the bench.py contract): every BASELINE.json config and the companion kernels, one GPU.
Prints one JSON object; used for DESIGN.md / profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vrvq_b200 import ops
from tests.golden import gen_inputs as gi

PEAK = 6544.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e-3


def encode_case(name, B, T, Nq, zqis, vbr=True, D=1024):
    sd = gi.torch_state_dict(gi.make_state_dict(1, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    nbuf = max(2, min(4, int(300e6 // (B * D * T * 4)) + 1))
    zs = [torch.randn(B, D, T, device="cuda") for _ in range(nbuf)]
    imp = torch.rand(B, 1, T, device="cuda") if vbr else None
    out = ops.EncodeOutputs(B, D, T, Nq, "cuda", z_q=True, z_q_is=zqis, latents=True, mask=True)
    i = [0]
    def fn():
        ops.rvq_encode_into(pw, zs[i[0] % nbuf], out, Nq, imp, 0.5 if vbr else None, zero_accum=False); i[0] += 1
    t = timeit(fn, iters=10 if B * T > 200000 else 20)
    bpf = 4 * D * 2 + (4 if vbr else 0) + Nq * (8 + 4 + 32) + (4 * D * Nq if zqis else 0)
    fr = B * T
    return {"case": name, "B": B, "T": T, "Nq": Nq, "z_q_is": zqis, "us": t * 1e6, "Mframes_per_s": fr / t / 1e6,
            "GBps_algorithmic": fr * bpf / t / 1e9, "frac_of_hbm_peak": fr * bpf / t / 1e9 / PEAK, "bytes_per_frame": bpf}


def main():
    res = {"hbm_peak_gbs": PEAK, "encode": [], "companions": []}
    res["encode"].append(encode_case("cfg1 (B=1, 1 s)", 1, 87, 8, True))
    res["encode"].append(encode_case("cfg2 full dict", 16, 862, 8, True))
    res["encode"].append(encode_case("cfg2 without z_q_is", 16, 862, 8, False))
    res["encode"].append(encode_case("cfg2 full dict, T padded to 864", 16, 864, 8, True))
    res["encode"].append(encode_case("cfg3 base_24kbps Nq=28, codes+z_q", 64, 862, 28, False, vbr=False))
    res["encode"].append(encode_case("cfg4 shard (32 of 256 items x 60 s), codes+z_q", 32, 5168, 8, False, vbr=False))
    # companions at cfg2 size
    B, T, Nq, D = 16, 862, 8, 1024
    sd = gi.torch_state_dict(gi.make_state_dict(1, Nq, D))
    pw = ops.PackedWeights.from_state_dict(sd, "cuda")
    z = torch.randn(B, D, T, device="cuda"); imp = torch.rand(B, 1, T, device="cuda")
    enc = ops.rvq_encode(pw, z, None, imp, 1.0, want_z_q_is=True)
    fr = B * T
    t = timeit(lambda: ops.from_codes(pw, enc.codes, want_z_q_is=False))
    b = fr * (8 * Nq + 4 * D + 32 * Nq)
    res["companions"].append({"kernel": "from_codes (z_q + z_p)", "us": t * 1e6, "Mframes_per_s": fr / t / 1e6, "GBps_algorithmic": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / PEAK})
    t = timeit(lambda: ops.remask(enc.z_q_is, imp, 4.0))
    b = fr * (4 * D * Nq + 4 * D + 4 * Nq + 4)
    res["companions"].append({"kernel": "remask (one level of the sweep)", "us": t * 1e6, "Mframes_per_s": fr / t / 1e6, "GBps_algorithmic": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / PEAK})
    t = timeit(lambda: ops.search_latents(pw, enc.latents))
    b = fr * (32 * Nq + 8 * Nq)
    res["companions"].append({"kernel": "search_latents", "us": t * 1e6, "Mframes_per_s": fr / t / 1e6, "GBps_algorithmic": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / PEAK})
    t = timeit(lambda: ops.generate_mask_hard(imp * 8.0, Nq))
    res["companions"].append({"kernel": "generate_mask_hard", "us": t * 1e6})
    t = timeit(lambda: ops.mask_sums(enc.mask))
    res["companions"].append({"kernel": "mask_sums (cal_bpf numerator)", "us": t * 1e6})
    # wire format (section 8(f) row 4): algorithmic bytes per frame = int64 codes + float mask in, uint16 codes + one count out
    from vrvq_b200 import wire
    for Bw, Tw in ((16, 862), (32, 5168)):
        cw = torch.randint(0, 1024, (Bw, Nq, Tw), device="cuda")
        mw = ops.generate_mask_hard(torch.rand(Bw, 1, Tw, device="cuda") * 8.0, Nq)
        b = Bw * Tw * (12 * Nq + 2 * Nq + 1)
        t = timeit(lambda: wire.pack_codes(cw, mw))
        res["companions"].append({"kernel": f"pack_codes B={Bw} T={Tw} (incl. the error-flag read-back)", "us": t * 1e6, "GBps_algorithmic": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / PEAK})
        u16, cnt = wire.pack_codes(cw, mw)
        t = timeit(lambda: wire.unpack_codes(u16, cnt))
        res["companions"].append({"kernel": f"unpack_codes B={Bw} T={Tw} (incl. the error-flag read-back)", "us": t * 1e6, "GBps_algorithmic": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / PEAK})
    # config 5: PyTorch encoder (+subnet) feeding the fused kernel, 4 items x 5 s per GPU
    import vrvq_b200
    torch.manual_seed(0)
    m = vrvq_b200.DAC_VRVQ(n_codebooks=8, model_type="VBR", level_min=0.125, level_max=6.0).cuda().eval()
    x = torch.randn(4, 1, 431 * 512, device="cuda") * 0.1
    with torch.no_grad():
        t_all = timeit(lambda: m.encode(x, None, 1.0), iters=5, warm=2)
        zf = m.encoder(x, return_feat=True)
        t_enc = timeit(lambda: m.encoder(x, return_feat=True), iters=5, warm=2)
        t_sub = timeit(lambda: m.quantizer.imp_subnet(zf[1]), iters=5, warm=2)
    res["cfg5"] = {"B_per_gpu": 4, "T": 431, "encode_ms": t_all * 1e3, "encoder_ms": t_enc * 1e3, "subnet_ms": t_sub * 1e3,
                   "rvq_ms": (t_all - t_enc - t_sub) * 1e3, "frames_per_s_e2e": 4 * 431 / t_all,
                   "note": "conv encoder is PyTorch/cuDNN (upstream producer); importance subnet = six snake_conv3 launches (csrc/subnet.cu); the fused RVQ kernel is the remainder"}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
