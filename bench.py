#!/usr/bin/env python
"""bench.py -- RVQ latent frames/s of the fused sm_100a encode on BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "cfg2"): DAC_VRVQ RVQ + importance-map hard mask, B=16 x 10 s of 44.1 kHz audio
= 16 x 862 latent frames per GPU, D=1024, Nq=8, codebook 1024x8, VBR mode, level cycling over {0.25, 0.5, 1.0};
one "step" = one fused encode of the batch producing the reference's full output dict (codes int64, z_q, z_q_is,
latents, mask_imp) plus the loss sum and the kept-frame counts.  Synthetic latents, seeded random weights.

Keys (see the task contract): value = whole-job frames/s with inputs resident in HBM; e2e = the same through
VBRResidualVectorQuantize.forward with pinned HOST buffers (H2D of z+imp_map and D2H of codes/mask/loss/kept inside
the timed region); roofline = algorithmic bytes / measured launch time against the measured HBM copy peak;
cpu_baseline = the reference's CPU path (the live reference when its checkout is present, else the bit-identical eager-PyTorch
port oracle/torch_port.py) on a bounded sample; e2e_full / e2e_full_dict = e2e with z_q / the whole output dict read back.
Multi-GPU: one process per GPU (torchrun), independent batch shards, no data-path collective ("weak" scaling).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(workload="cfg2", B=16, T=862, D=1024, Nq=8, K=1024, levels=[0.25, 0.5, 1.0])
N_INPUT_BUFFERS = 4  # 4 x 56.5 MB of latents rotate (226 MB > 126 MB L2); each step also writes 570 MB of outputs
CD = 8


def algorithmic_bytes_per_frame(D, Nq, z_q_is=True):
    """read z (+imp), write z_q, int64 codes, f32 mask, latents, optional z_q_is."""
    b = 4 * D + 4 + 4 * D + 8 * Nq + 4 * Nq + 4 * CD * Nq
    if z_q_is:
        b += 4 * D * Nq
    return b


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period_s=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        win = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if len(win) < 3:
            win, window = list(self.samples), "warmup+timed (timed region shorter than 3 samples)"
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "window": window}
        mhz = sorted(s[1] for s in win)
        mask = 0
        for s in win:
            mask |= s[2]
        reasons = [n for b, n in names.items() if mask & b and n != "gpu_idle"]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(win), "window": window}


_ORIG_AFFINITY = None


def bind_to_gpu_numa_node(index):
    """Pin this process (and with it the first-touch placement of the pinned host buffers) to the CPUs NVML reports as
    local to GPU `index`; returns what was found, for the JSON line.  With every GPU of a box on one NUMA node (as on this
    pool) this cannot help the N=8 host->device rate: the ranks then share that node's memory bandwidth."""
    info = {"cpus_total": os.cpu_count()}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            info["numa_node"] = None
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = os.sched_getaffinity(0)
        allowed = sorted(set(cpus) & set(_ORIG_AFFINITY))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound_cpus"] = len(allowed)
    except Exception as e:
        info["affinity_error"] = str(e)[:80]
    return info


def make_state(device):
    import torch

    import vrvq_b200
    from tests.golden import gen_inputs as gi

    sd = gi.torch_state_dict(gi.make_state_dict(0, CFG["Nq"], CFG["D"], CFG["K"]))
    m = vrvq_b200.VBRResidualVectorQuantize(input_dim=CFG["D"], n_codebooks=CFG["Nq"], codebook_size=CFG["K"], codebook_dim=8,
                                            level_min=0.125, level_max=6.0, imp2mask_alpha=2.0)
    m.load_state_dict(sd, strict=False)
    return m.to(device).eval(), sd


def host_inputs(rank, n):
    """Synthetic latents N(0,1) (seed 1234+rank) and importance maps U(0,1) (seed 4321+rank), SURVEY.md 8(d)."""
    import torch

    g = torch.Generator().manual_seed(1234 + rank)
    zs = [torch.randn(CFG["B"], CFG["D"], CFG["T"], generator=g) for _ in range(n)]
    g2 = torch.Generator().manual_seed(4321 + rank)
    imps = [torch.rand(CFG["B"], 1, CFG["T"], generator=g2) for _ in range(n)]
    return zs, imps


def cpu_baseline_run(sd, steps, warmup, sample_B=4):
    """The reference's CPU path on a bounded sample: `sample_B` items x 862 frames of the cfg2 workload per step, all host
    threads.  kind "reference": the UNMODIFIED reference classes imported from its checkout (oracle/ref_import.py; build
    container only).  kind "port": oracle/torch_port.py, the same eager ATen op sequence, pinned bit-for-bit to the live
    reference by tests/test_oracle_vs_reference.py::test_torch_port_is_bit_identical_to_the_live_reference -- what runs on the
    GPU box, where the checkout does not exist.  B = 4 per step keeps the default run short and is the batch size the CPU
    likes best (B = 16 per step is ~2x slower per frame), so ratios against it are conservative."""
    import torch

    from oracle import ref_import, torch_port

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(sample_B, CFG["D"], CFG["T"], generator=g)
    imp = torch.rand(sample_B, 1, CFG["T"], generator=torch.Generator().manual_seed(4321))
    lv = CFG["levels"]
    kind = "port"
    if ref_import.available() and os.environ.get("VRVQ_BENCH_FORCE_PORT") != "1":
        try:
            ref = ref_import.load()
            m = ref.VBRResidualVectorQuantize(input_dim=CFG["D"], n_codebooks=CFG["Nq"], codebook_size=CFG["K"], codebook_dim=8,
                                              level_min=0.125, level_max=6.0, imp2mask_alpha=2.0).eval()
            m.load_state_dict(sd, strict=False)

            class _Fixed(torch.nn.Module):  # the importance map is an input of this workload, as on the GPU arm
                def forward(self, feat):
                    return imp

            m.imp_subnet = _Fixed()

            def fwd(level):
                with torch.no_grad():
                    return m(z, n_quantizers=None, feat_enc=z, level=level)

            kind = "reference"
        except Exception as e:  # fall back to the pinned port, and say so
            sys.stderr.write(f"bench.py: live reference unusable ({e}); timing oracle/torch_port.py\n")
    if kind == "port":
        w = torch_port.TorchPortWeights(sd)

        def fwd(level):
            return torch_port.rvq_forward(w, z, None, imp, level)

    for i in range(warmup):
        fwd(lv[i % 3])
    t0 = time.perf_counter()
    for i in range(steps):
        fwd(lv[i % 3])
    dt = time.perf_counter() - t0
    frames = sample_B * CFG["T"] * steps
    what = ("the unmodified reference VBRResidualVectorQuantize.forward (models/quantize.py:328-443) imported from its checkout"
            if kind == "reference" else "oracle/torch_port.py, the reference's eager ATen op sequence (bit-identical to it, tests/test_oracle_vs_reference.py)")
    return {"value": frames / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind, "sample_B": sample_B,
            "sample": f"{steps} steps x B={sample_B} x T={CFG['T']} frames of the cfg2 workload on the host CPU, fp32, level cycling, "
                      f"imp_map given: {what}", "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from tests.golden import gen_inputs as gi

    sd = gi.torch_state_dict(gi.make_state_dict(0, CFG["Nq"], CFG["D"], CFG["K"]))
    r = cpu_baseline_run(sd, args.steps, args.warmup)
    cfg = config_dict()
    # this arm is ONE CPU process whatever --gpus says: say what it ran, not what the GPU arm runs
    cfg.update({"B_per_step": r["sample_B"], "parallelism": "one CPU process, all host threads (torch intra-op)",
                "l2": "n/a (CPU)", "note": f"bounded sample: B={r['sample_B']} x T={CFG['T']} frames per step (the GPU arm runs B={CFG['B']} per GPU per step); "
                                           "frames/s is per frame, so the arms compare per frame"})
    cfg.pop("B_per_gpu", None)
    line = {"impl": "reference", "metric": "rvq_latent_frames_per_sec", "value": r["value"], "unit": "frames/s", "n_gpus": 0,
            "requested_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "torch": torch.__version__}
    print(json.dumps(line), flush=True)
    return 0


def config_dict():
    return {"workload": CFG["workload"], "B_per_gpu": CFG["B"], "T": CFG["T"], "D": CFG["D"], "n_codebooks": CFG["Nq"],
            "codebook": "1024x8", "mode": "VBR, level cycling 0.25/0.5/1.0, full output dict incl. z_q_is",
            "l2": f"inputs rotate over {N_INPUT_BUFFERS} buffers (226 MB > 126 MB L2); 570 MB of outputs written per step",
            "parallelism": "independent batch shards, one process per GPU, no collective"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_info = bind_to_gpu_numa_node(local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; keep stdout clean for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from vrvq_b200 import _lib, ops

    model, sd = make_state(dev)
    pw = model.packed_weights(dev)
    B, T, D, Nq = CFG["B"], CFG["T"], CFG["D"], CFG["Nq"]
    frames = B * T
    zs_h, imps_h = host_inputs(rank, N_INPUT_BUFFERS)
    zs = [z.to(dev) for z in zs_h]
    imps = [i.to(dev) for i in imps_h]
    out = ops.EncodeOutputs(B, D, T, Nq, dev, z_q=True, z_q_is=True, latents=True, mask=True)
    levels = CFG["levels"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        ops.rvq_encode_into(pw, zs[i % N_INPUT_BUFFERS], out, Nq, imps[i % N_INPUT_BUFFERS], levels[i % 3], zero_accum=False)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    # ---- burst figure (reported beside the headline, never instead of it): 100 launches right after the warm-up, before the power cap
    # of a long back-to-back run pulls the SM clock down (a 4000-launch region runs at ~1800 of 1965 MHz on this pool)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    burst_steps = max(1, min(args.steps, 100))
    b0.record()
    for i in range(burst_steps):
        step(i)
    b1.record()
    barrier()
    burst_ms = b0.elapsed_time(b1)
    # ---- kernel-resident timing: exactly K steps between barrier+sync, CUDA events on the launch stream
    barrier()
    l0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _lib.launch_count - l0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- end-to-end through the public module API with pinned host buffers
    zs_p = [z.pin_memory() for z in zs_h]
    imps_p = [i.pin_memory() for i in imps_h]
    h2d = zs_p[0].numel() * 4 + imps_p[0].numel() * 4
    # Two streams alternate so that step i+1's host->device copy overlaps step i's kernel and read-back, as a
    # streaming caller would run it; every step still copies its own inputs in and its own results out.
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def pinned(shape, dtype):
        return [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(2)]

    h_codes, h_mask, h_small = pinned((B, Nq, T), torch.int64), pinned((B, Nq, T), torch.float32), pinned((Nq + 1,), torch.float64)
    d2h_small = h_codes[0].numel() * 8 + h_mask[0].numel() * 4 + h_small[0].numel() * 8

    def measure_e2e(back, n_steps):
        """back: "codes" = codes + mask + kept counts + loss come back (z_q / z_q_is stay on the device for the decoder);
        "z_q" = additionally the quantised latent z_q; "dict" = the reference's whole output dict incl. z_q_is and latents."""
        h_zq = pinned((B, D, T), torch.float32) if back in ("z_q", "dict") else None
        h_zqis = pinned((B, Nq, D, T), torch.float32) if back == "dict" else None
        h_lat = pinned((B, CD * Nq, T), torch.float32) if back == "dict" else None
        d2h = d2h_small + (h_zq[0].numel() * 4 if h_zq else 0) + (h_zqis[0].numel() * 4 + h_lat[0].numel() * 4 if h_zqis else 0)

        def e2e_step(i):
            k = i % 2
            with torch.cuda.stream(streams[k]):
                z = zs_p[i % N_INPUT_BUFFERS].to(dev, non_blocking=True)
                imp = imps_p[i % N_INPUT_BUFFERS].to(dev, non_blocking=True)
                r = model(z, n_quantizers=None, feat_enc=None, level=levels[i % 3], imp_map=imp)
                h_codes[k].copy_(r["codes"], non_blocking=True)
                h_mask[k].copy_(r["mask_imp"], non_blocking=True)
                h_small[k][:Nq].copy_(r["kept_frames"].to(torch.float64), non_blocking=True)
                h_small[k][Nq:].copy_(r["commitment_loss"].to(torch.float64).reshape(1), non_blocking=True)
                if h_zq is not None:
                    h_zq[k].copy_(r["z_q"], non_blocking=True)
                if h_zqis is not None:
                    h_zqis[k].copy_(r["z_q_is"], non_blocking=True)
                    h_lat[k].copy_(r["latents"], non_blocking=True)

        for i in range(2):
            e2e_step(i)
        barrier()
        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for st in streams:
            st.wait_stream(cur)
        for i in range(n_steps):
            e2e_step(i)
        for st in streams:
            cur.wait_stream(st)
        e1.record(cur)
        barrier()
        return e0.elapsed_time(e1), d2h

    # ---- the reference's real signature, forward(z, None, feat_enc, level): the importance subnet (csrc/subnet.cu) runs inside the call
    def measure_subnet_call(n_steps):
        feat = [torch.randn(B, D, T, generator=torch.Generator().manual_seed(777 + rank + i)).to(dev) for i in range(2)]
        for i in range(2):
            model(zs[i % N_INPUT_BUFFERS], n_quantizers=None, feat_enc=feat[i % 2], level=levels[i % 3])
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_steps):
            model(zs[i % N_INPUT_BUFFERS], n_quantizers=None, feat_enc=feat[i % 2], level=levels[i % 3])
        s1.record()
        barrier()
        return s0.elapsed_time(s1)

    sub_steps = max(4, min(args.steps, 30))
    sub_ms = measure_subnet_call(sub_steps)

    e2e_steps = max(4, min(args.steps, 50))
    e2e_ms, d2h = measure_e2e("codes", e2e_steps)
    zq_steps, dict_steps = max(4, min(args.steps, 30)), max(4, min(args.steps, 8))
    e2e_zq_ms, d2h_zq = measure_e2e("z_q", zq_steps)
    e2e_dict_ms, d2h_dict = measure_e2e("dict", dict_steps)
    sampler.stop()

    # max over ranks
    t = torch.tensor([ms, e2e_ms, e2e_zq_ms, e2e_dict_ms, sub_ms, burst_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e_zq_ms_max, e2e_dict_ms_max, sub_ms_max, burst_ms_max = (float(x) for x in t.tolist())

    if rank == 0:
        bytes_per_launch = algorithmic_bytes_per_frame(D, Nq, True) * frames
        launch_ms = ms / args.steps  # the timed region is K back-to-back launches of the one fused kernel
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        info = ops.encode_launch_info(pw, B, T, Nq, dev, z_q_is=True)
        line = {
            "metric": "rvq_latent_frames_per_sec", "value": world * frames * args.steps / (ms_max * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(CFG["workload"]),
                         "kernel": "rvq_encode_tc_kernel<1024,true> (tcgen05 kind::tf32, TMA-staged latent)" if info["kernel"] == "tc" else "rvq_encode_kernel<1024,1024> (CUDA cores)",
                         "algorithmic_bytes_per_frame": algorithmic_bytes_per_frame(D, Nq, True), "frames_per_launch": frames,
                         "launch_us": launch_ms * 1e3, "peak_source": peak_src, "grid": info["grid"], "block": info["block"],
                         "smem_bytes": info["smem_bytes"],
                         "burst": {"steps": burst_steps, "launch_us": burst_ms_max / burst_steps * 1e3,
                                   "frac": bytes_per_launch / (burst_ms_max / burst_steps * 1e-3) / 1e9 / peak,
                                   "what": "the same launch over the first 100 steps after the warm-up, before the power cap of the long region sets in; "
                                           "`frac` and `value` above are the sustained figures"}},
            "e2e": {"value": world * frames * e2e_steps / (e2e_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms_max / e2e_steps,
                    "api": "VBRResidualVectorQuantize.forward(z, level, imp_map) on pinned host buffers, two alternating CUDA streams; codes, mask, "
                           "kept counts and loss come back, z_q/z_q_is stay on the device (an encoder hands them to the decoder there)"},
            # the same call with more of the result read back: the quantised latent, and the reference's whole output dict
            "e2e_full": {"value": world * frames * zq_steps / (e2e_zq_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h_zq, "steps": zq_steps, "ms_per_step": e2e_zq_ms_max / zq_steps,
                         "what": "e2e + z_q [B,D,T] copied back"},
            "e2e_full_dict": {"value": world * frames * dict_steps / (e2e_dict_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": d2h_dict, "steps": dict_steps, "ms_per_step": e2e_dict_ms_max / dict_steps,
                              "what": "e2e + z_q, z_q_is [B,Nq,D,T] and latents copied back (every key of the reference's dict); PCIe-bound"},
            # device-resident, but through the reference's own signature: the importance map is computed inside the call
            "e2e_subnet": {"value": world * frames * sub_steps / (sub_ms_max * 1e-3), "unit": "frames/s", "steps": sub_steps,
                           "ms_per_step": sub_ms_max / sub_steps,
                           "what": "VBRResidualVectorQuantize.forward(z, None, feat_enc, level) with z and feat_enc resident in HBM: the six "
                                   "importance-subnet launches (blocks 0-2 on tcgen05 3xTF32, fused tail; 9.85 MFLOP per frame) + the fused encode; `value` above takes "
                                   "imp_map as an input (SURVEY.md 8(d))"},
            "host": host_info,
            "gpu_launches": launches * world, "clocks": clocks, "torch": torch.__version__,
        }
        if world == 1 and not args.no_cpu_baseline:
            if _ORIG_AFFINITY is not None:
                os.sched_setaffinity(0, _ORIG_AFFINITY)  # the CPU baseline gets every host core back
            cb = cpu_baseline_run(sd, steps=args.cpu_steps, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        if saved_stdout is not None and rank == 0:
            os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)  # ~0.6 s of back-to-back launches: max-over-ranks jitter stays below 1 %
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=60, help="steps of the bounded CPU-baseline sample (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing hygiene: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps",
               str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
