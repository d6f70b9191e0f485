"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which runs the real reference)
and by the tests (which re-create the same inputs for the oracle and the CUDA path).

numpy's PCG64 stream is stable across numpy versions, so a fixture only needs the seed.
Weights are laid out exactly like the reference state dict:
quantizers.{i}.in_proj.weight_v [8,D,1], .weight_g [8,1,1], .bias [8],
quantizers.{i}.out_proj.weight_v [D,8,1], .weight_g [D,1,1], .bias [D], quantizers.{i}.codebook.weight [K,8].
Unlike the reference's init (g = ||v||, zero biases) g and the biases are randomised so the
weight-norm fold and both bias paths are exercised.
"""
import numpy as np

CD = 8


def make_state_dict(seed: int, n_codebooks: int, input_dim: int, codebook_size: int = 1024):
    rng = np.random.Generator(np.random.PCG64(seed))
    f32 = np.float32
    sd = {}
    D, K = input_dim, codebook_size
    for i in range(n_codebooks):
        p = f"quantizers.{i}."
        v_in = (rng.uniform(-1.0, 1.0, (CD, D, 1)) / np.sqrt(D)).astype(f32)
        g_in = (np.sqrt((v_in.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
                * rng.uniform(0.5, 1.5, (CD, 1, 1))).astype(f32)
        v_out = (rng.uniform(-1.0, 1.0, (D, CD, 1)) / np.sqrt(CD)).astype(f32)
        g_out = (np.sqrt((v_out.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
                 * rng.uniform(0.5, 1.5, (D, 1, 1))).astype(f32)
        sd[p + "in_proj.weight_g"] = g_in
        sd[p + "in_proj.weight_v"] = v_in
        sd[p + "in_proj.bias"] = rng.normal(0.0, 0.05, (CD,)).astype(f32)
        sd[p + "out_proj.weight_g"] = g_out
        sd[p + "out_proj.weight_v"] = v_out
        sd[p + "out_proj.bias"] = rng.normal(0.0, 0.05, (D,)).astype(f32)
        sd[p + "codebook.weight"] = rng.normal(0.0, 1.0, (K, CD)).astype(f32)
    return sd


def make_latents(seed: int, B: int, D: int, T: int, sigma: float = 1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.normal(0.0, sigma, (B, D, T))).astype(np.float32)


def make_imp_map(seed: int, B: int, T: int):
    """U(0,1) importance map [B,1,T] with a few exact mask-edge values (k / (level*Nq) style) mixed in."""
    rng = np.random.Generator(np.random.PCG64(seed))
    imp = rng.uniform(0.0, 1.0, (B, 1, T)).astype(np.float32)
    edges = np.array([0.0, 0.125, 0.25, 0.5, 0.75, 1.0], np.float32)
    n = min(T, len(edges))
    imp[0, 0, :n] = edges[:n]
    return imp


def torch_state_dict(sd):
    import torch

    return {k: torch.from_numpy(np.array(v)) for k, v in sd.items()}


# name -> parameters of every committed fixture
CASES = {
    # config-1/2 shaped: VBR model, D=1024, Nq=8 (conf/base.yml:9-11), level sweep of configs[1]
    "vbr_d1024_nq8": dict(kind="vbr", seed=11, D=1024, Nq=8, K=1024, B=2, T=87, sigma=1.0,
                          levels=[0.25, 0.5, 1.0], imp_seed=21),
    # per-item tensor level [B,1,1]
    "vbr_tensor_level": dict(kind="vbr", seed=12, D=1024, Nq=8, K=1024, B=3, T=21, sigma=0.7,
                             levels=[[0.3, 1.0, 2.5]], imp_seed=22),
    # CBR class, all stages and early exit (quantize.py:183-184)
    "cbr_d1024_nq8": dict(kind="cbr", seed=13, D=1024, Nq=8, K=1024, B=2, T=45, sigma=1.0, n_quantizers=[None, 3]),
    # conf/base_24kbps.yml:9 (Nq=28)
    "cbr_d1024_nq28": dict(kind="cbr", seed=14, D=1024, Nq=28, K=1024, B=1, T=33, sigma=1.0, n_quantizers=[None]),
    # class defaults of ResidualVectorQuantize (quantize.py:112-119): input_dim=512, n_codebooks=9
    "cbr_d512_nq9": dict(kind="cbr", seed=15, D=512, Nq=9, K=1024, B=1, T=40, sigma=1.0, n_quantizers=[None]),
    # edge shapes: T=1 and T=3, small-magnitude latents (random-init encoder scale)
    "vbr_t1": dict(kind="vbr", seed=16, D=1024, Nq=8, K=1024, B=3, T=1, sigma=0.007, levels=[1.0], imp_seed=26),
    "cbr_t3": dict(kind="cbr", seed=17, D=1024, Nq=8, K=1024, B=2, T=3, sigma=0.007, n_quantizers=[None, 1]),
}


def make_subnet_state_dict(seed: int, d_input: int, d_feat: int, widths=(512, 128, 32, 8), out_channels: int = 1):
    """Importance-subnet parameters with the reference's keys (models/importance_subnet.py:18-34): per block
    `<snake>.alpha [1,Cin,1]`, `<conv>.bias [Cout]`, `.weight_g [Cout,1,1]`, `.weight_v [Cout,Cin,3]`.  Rows of v have
    about unit norm; alpha, g and the biases are randomised so that the Snake and the weight-norm fold are exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f32 = np.float32
    cin = [d_input, d_feat] + list(widths)
    cout = [d_feat] + list(widths) + [out_channels]
    names = [("in_block.0.", "in_block.1.")] + [(f"blocks.{i}.0.", f"blocks.{i}.1.") for i in range(len(widths) + 1)]
    sd = {}
    for (sk, ck), ci, co in zip(names, cin, cout):
        v = (rng.uniform(-1.0, 1.0, (co, ci, 3)) / np.sqrt(ci)).astype(f32)
        g = (np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)) * rng.uniform(0.5, 1.5, (co, 1, 1))).astype(f32)
        sd[sk + "alpha"] = rng.uniform(0.5, 1.5, (1, ci, 1)).astype(f32)
        sd[ck + "bias"] = rng.normal(0.0, 0.05, (co,)).astype(f32)
        sd[ck + "weight_g"] = g
        sd[ck + "weight_v"] = v
    return sd


# importance-subnet fixtures (tests/golden/subnet_*.npz hold the reference's imp_map only; inputs come from the seeds)
SUBNET_CASES = {
    # the shipped architecture (models/quantize.py:317-323: d_input = d_feat = 1024, widths 512/128/32/8), config-1 length
    "subnet_d1024": dict(seed=31, d_input=1024, d_feat=1024, widths=(512, 128, 32, 8), B=2, T=87, sigma=1.0),
    # narrow net, frame counts around the 128-frame tile edge and the degenerate lengths
    "subnet_small_t130": dict(seed=32, d_input=64, d_feat=48, widths=(32, 16, 8, 8), B=3, T=130, sigma=1.5),
    "subnet_small_t1": dict(seed=33, d_input=64, d_feat=48, widths=(32, 16, 8, 8), B=2, T=1, sigma=1.5),
    "subnet_small_t3": dict(seed=34, d_input=64, d_feat=48, widths=(32, 16, 8, 8), B=1, T=3, sigma=1.5),
}


def make_dac_state_dict(seed: int, shapes, imp_gain: float = 6.0, w_scale: float = 1.5):
    """Seeded parameters for a whole (decoder-less) DAC_VRVQ: `shapes` is the ordered {key: shape} of the model's own
    state dict (the reference's and the mirror's key lists are identical, tests/test_oracle_vs_reference.py), so the
    22 M encoder weights never have to be stored.  weight_v ~ U(-1,1) * sqrt(w_scale / fan_in),
    weight_g = ||v|| * U(0.8, 1.2), biases N(0, 0.02), Snake alpha U(0.5, 1.5), codebooks N(0, 1); the last conv of the
    importance subnet gets `imp_gain` x larger g so that the map spreads over (0, 1) instead of hugging 0.5."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f32 = np.float32
    sd = {}
    last_v = None
    for key, shape in shapes.items():
        shape = tuple(shape)
        if key.endswith("weight_v"):
            fan_in = int(np.prod(shape[1:]))
            last_v = (rng.uniform(-1.0, 1.0, shape) * np.sqrt(w_scale / fan_in)).astype(f32)
            sd[key] = last_v
        elif key.endswith("alpha"):
            sd[key] = rng.uniform(0.5, 1.5, shape).astype(f32)
        elif key.endswith("bias"):
            sd[key] = rng.normal(0.0, 0.02, shape).astype(f32)
        elif key.endswith("codebook.weight"):
            sd[key] = rng.normal(0.0, 1.0, shape).astype(f32)
        elif key.endswith("weight_g"):
            sd[key] = None  # filled below (the reference orders g before v)
        else:
            raise KeyError(f"unexpected parameter {key}")
    for key in list(sd):
        if key.endswith("weight_g"):
            v = sd[key[:-1] + "v"].astype(np.float64)
            nrm = np.sqrt((v ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
            gain = imp_gain if key == "quantizer.imp_subnet.blocks.4.1.weight_g" else 1.0
            sd[key] = (nrm * gain * rng.uniform(0.8, 1.2, nrm.shape)).astype(f32)
    return sd


def make_audio(seed: int, B: int, samples: int, scale: float = 0.5):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.normal(0.0, scale, (B, 1, samples))).astype(np.float32)


# DAC_VRVQ.encode fixtures (models/dac_vrvq.py:176-213): BASELINE.json configs[0] -- 1 s of 44.1 kHz mono, level 1 -- and a CBR model
DAC_CASES = {
    "dac_vbr_1s": dict(seed=41, model_type="VBR", n_codebooks=8, B=1, samples=44100, level=1.0, n_quantizers=None),
    "dac_cbr_1s": dict(seed=42, model_type="CBR", n_codebooks=8, B=2, samples=22050, level=None, n_quantizers=5),
}
