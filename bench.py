#!/usr/bin/env python
"""Benchmark of the SD1.5 sampler hot path on B200 (metric of BASELINE.json: UNet sampler iterations / s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

1 step ("it") = one sampler iteration of config 2 (SD1.5 1024x1024, dpmpp_2m_cfgpp, bs=1 per GPU): one CFG-batched
(uncond+cond => 2 rows) UNet forward with EPS scaling + CFG combine + solver update.  Full resolution every step
(multiscale off) so every step is the 9.348 TFLOP workload BASELINE.md quotes.  Synthetic seeded weights
(no checkpoints exist offline).  N > 1: one process per GPU (torchrun), each rank owns its own image(s) for the whole
trajectory, no data-path collective (weak scaling); value = all ranks' iterations / max-over-ranks device time.

Beside the contract keys the line carries (N = 1) `secondary`: whole KSampler runs, pipe() end to end, BASELINE config 3
(bs = 32) and config 5 (HiresFix 512 -> 2048 + VAE decode), the REAL reference timed on this GPU through its own stock path
(`gpu_reference`: the comparator north_star names) and the reference's loop driven through the engine seam (`seam`); and
(N > 1) `sharded`: config 3 and config 5 through distributed.sample_sharded on NCCL with an in-run check that the gathered
latents equal what one process computes for the same images.

--impl reference: the CPU arm.  The unmodified reference (snapshotted into baseline/_ref by build(), see
baseline/run_reference.py) runs the same workload through its own sampling stack on all host cores; if the snapshot is
absent the oracle port (oracle/sd15_oracle.py, pinned against the reference's own outputs) is timed instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_TFLOP_1024 = 9.348  # per CFG step bs=1 (B=2), BASELINE.md §2 (measured on the reference module)
UNET_TFLOP_512 = 1.607


def workload_string(size: int, bs: int) -> str:
    """One string for both arms (the driver compares them)."""
    return (f"SD1.5 txt2img {size}x{size} dpmpp_2m_cfgpp bs={bs}/GPU (CFG pair, UNet batch {2 * bs}), "
            "multiscale off (every step full resolution)")


def run_reference_script(extra, timeout):
    """baseline/run_reference.py in a child process (its own CUDA context / torch flags); returns its JSON line or an
    {"unavailable": why} dict.  Never raises: a comparator that cannot run must not take the engine's line down."""
    script = os.path.join(ROOT, "baseline", "run_reference.py")
    try:
        r = subprocess.run([sys.executable, script] + list(extra), capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout} s"}
    for ln in reversed(r.stdout.splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except Exception:
                break
    return {"unavailable": ("rc %d: " % r.returncode) + (r.stderr.strip().splitlines()[-1][:300] if r.stderr.strip() else "no output")}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            have = [k for k in ("bf16_tflops", "bf16_tflops_sustained", "hbm_gbs") if k in d]
            source = "measured" if len(have) == 3 else ("measured (" + ", ".join(have) + "), fallback for the rest" if have else "fallback")
            return dict(bf16=float(d.get("bf16_tflops", 1590.0)), bf16_sus=float(d.get("bf16_tflops_sustained", 1400.0)),
                        hbm=float(d.get("hbm_gbs", 6650.0)), source=source)
        except Exception:
            pass
    return dict(bf16=1590.0, bf16_sus=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_step_time(lat: int, max_seconds: float, max_steps: int, warmup: int):
    """Times oracle CFG steps (2-row UNet forward + lerp + dpmpp_2m update) on the host cores."""
    import torch
    from oracle import sd15_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    sd = O.synth_state_dict(O.unet_param_shapes())
    g = torch.Generator().manual_seed(1234)
    ctx = torch.randn(2, 77, 768, generator=g)
    x = torch.randn(1, 4, lat, lat, generator=g) * 14.6
    tables = O.make_sigma_tables()
    sig = O.calculate_sigmas("karras", 30)

    def step(i, x):
        s = sig[i] * x.new_ones([1])
        den = O.cfg_denoise(lambda a, b, c: O.apply_model(sd, a, b, c, tables), x, s, ctx[1:2], ctx[0:1], 7.0)
        return (sig[i + 1] / sig[i]) * x - torch.expm1(torch.log(sig[i + 1] / sig[i])) * den

    t_start = time.time()
    done_w = 0
    for i in range(warmup):
        if time.time() - t_start > max_seconds * 0.4:
            break
        x = step(i, x)
        done_w += 1
    t0 = time.time()
    n = 0
    while n < max_steps:
        x = step(done_w + n, x)
        n += 1
        if time.time() - t_start > max_seconds:
            break
    dt = time.time() - t0
    return n, dt, done_w, torch.get_num_threads()


def run_reference_arm(args):
    """CPU arm (see module docstring). Rank 0 only; other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(1, args.steps), min(1, max(0, args.warmup))
    ref = run_reference_script(["--device", "cpu", "--size", str(args.size), "--steps", str(K), "--warmup", str(W),
                                "--bs", str(args.bs)], timeout=1500)
    if "unavailable" not in ref:
        v, n, dt, done_w, threads, kind = ref["it_per_s"], ref["steps"], ref["seconds"], ref["warmup"], ref["threads"], "reference"
        sample = (f"{n} timed + {done_w} warm-up full {args.size}x{args.size} sampler steps of the UNMODIFIED reference "
                  f"(baseline/_ref: sampling.ksampler('dpmpp_2m_cfgpp', enable_multiscale=False) + sampling.sample on the CPU, "
                  f"weights {ref['unet_dtype']} cast per call to {ref['manual_cast']}, {ref['attention']}, torch {ref['torch']})")
        dtype = "f32"
    else:
        lat = args.size // 8
        n, dt, done_w, threads = cpu_oracle_step_time(lat, max_seconds=150.0, max_steps=K, warmup=W)
        v, kind = n / dt, "port"
        sample = (f"{n} timed + {done_w} warm-up full {args.size}x{args.size} CFG steps (2-row UNet fwd + CFG + dpmpp_2m update) "
                  f"of the oracle port (torch fp32 CPU), capped at 150 s wall; reference snapshot unavailable: {ref['unavailable']}")
        dtype = "f32"
    line = {
        "impl": "reference", "metric": "it/s (UNet sampler steps/sec)", "value": v, "unit": "it/s", "n_gpus": args.gpus,
        "steps": n, "warmup": done_w, "ms_per_step": 1000.0 * dt / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_string(args.size, args.bs), "sampler": "dpmpp_2m_cfgpp", "scheduler": "karras",
                   "cfg": 7.0, "images_per_gpu": args.bs},
        "cpu_baseline": {"value": v, "unit": "it/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ldn", choices=["ldn", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="image size in pixels (config 2 = 1024)")
    ap.add_argument("--bs", type=int, default=1, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the full-run / pipe() end-to-end secondary numbers")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="N = 1: skip the bs=32 / HiresFix batch configurations")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the unmodified reference on this GPU")
    ap.add_argument("--workload", default="sd15", choices=["sd15", "flux"],
                    help="sd15 = BASELINE config 2 (the contract line); flux = config 4, Flux.1-dev sized DiT steps (1 GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "flux":
        return run_flux(args)

    import torch
    import torch.distributed as dist

    from lightdiffusion_next_b200 import _lib as L
    from lightdiffusion_next_b200.engine import Engine
    from lightdiffusion_next_b200.sampling import SamplerLoop
    from lightdiffusion_next_b200.schedule import calculate_sigmas

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    bs = args.bs
    lat = args.size // 8
    peaks = load_peaks()

    # ---- weights: rank 0 generates the seeded synthetic state dict, NCCL-broadcasts it (north_star: weights broadcast once)
    from lightdiffusion_next_b200.synth import unet_shapes, synth_tensor
    shapes = unet_shapes()
    eng = Engine(max_rows=max(2 * bs, 8), max_h=lat, max_w=lat, max_ctx_tokens=77, use_graph=not args.no_graph, device=dev)
    names = sorted(shapes)
    sd = {}
    for nme in names:
        if rank == 0:
            t = synth_tensor(nme, shapes[nme]).to(dev)
        else:
            t = torch.empty(shapes[nme], dtype=torch.float16, device=dev)
        if world > 1:
            dist.broadcast(t, 0)
        sd[nme] = t
    eng.load_unet(sd)
    del sd
    torch.cuda.empty_cache()

    g = torch.Generator().manual_seed(1234 + rank)
    ctx = torch.randn(2, 77, 768, generator=g)
    ctx_rows = torch.cat([ctx[0:1].expand(bs, -1, -1), ctx[1:2].expand(bs, -1, -1)]).contiguous().to(dev)
    eng.set_context(ctx_rows)
    sig = calculate_sigmas(eng.schedule, "karras", 30)
    t_ = -torch.log(sig)
    ratios = (torch.exp(-t_)[1:] / torch.exp(-t_)[:-1])
    hexp = torch.expm1(-(t_[1:] - t_[:-1]))
    cfg = 7.0
    loop = SamplerLoop(eng, bs, lat, lat)
    x0 = (torch.randn(bs, 4, lat, lat, generator=g) * float(sig[0])).to(dev)
    den = torch.empty_like(x0)
    stream = torch.cuda.current_stream()

    def one_step(i, x):
        j = i % 29  # never the terminal sigma=0 step, so every step is the same work
        du, dc = loop.denoise_pair(x, float(sig[j]))
        eng.cfg_step(x, du, dc, cfg, 0, c0=float(ratios[j]), c1=float(hexp[j]), x_out=loop.x_next, denoised_out=den)
        x, loop.x_next = loop.x_next, x
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing (`value`)
    x = x0.clone()
    for i in range(W):
        x = one_step(i, x)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        x = one_step(i, x)
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(x).all().item())

    # ---------------------------------------------------------------- end-to-end through the host API (`e2e`)
    # every step: pinned-host latent + sigma -> device, UNet CFG step, result latent -> pinned host, sync.
    hx = torch.empty(bs, 4, lat, lat).pin_memory()
    hx.copy_(x0.cpu())
    hout = torch.empty_like(hx).pin_memory()
    dx = torch.empty_like(x0)

    def e2e_step(i):
        dx.copy_(hx, non_blocking=True)
        y = one_step(i, dx)
        hout.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hx.copy_(hout)

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if world > 1:
        tt = torch.tensor([ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    else:
        e2e_ms = e2e_s * 1000.0
    its = world * bs * K / (ms / 1000.0)
    e2e_its = world * bs * K / (e2e_ms / 1000.0)

    if rank == 0:
        # ------------------------------------------------------------ roofline of the dominant kernel
        # L0 self-attention (a9::attn9_tc_kernel: N=16384 tokens, 8 heads, d=40, UNet batch 2): 42 % of step FLOPs.  Called exactly
        # as the UNet program calls it: the [Q | K] buffer with Q pre-scaled by scale * log2(e) and a ones column in every K head
        # slot (folded operands, `causal` bit 1 of the C entry -- include/ldn.h).
        lib = eng.lib
        B2, H, N, d, slot = 2 * bs, 8, lat * lat, 40, 64
        Qb = torch.randn(B2 * N, 2 * H * slot, device=dev)
        Qb.view(B2 * N, 2 * H, slot)[:, :H] *= d ** -0.5 * 1.4426950408889634
        Qb = Qb.bfloat16()
        Qb.view(B2 * N, 2 * H, slot)[:, :, d:] = 0
        Qb.view(B2 * N, 2 * H, slot)[:, H:, d] = 1.0
        Vt = torch.zeros(H * 48, B2 * N, device=dev, dtype=torch.bfloat16)   # 48 rows per head: 40 values, ones row, 7 zero rows
        Vt.view(H, 48, B2 * N)[:, :d] = torch.randn(H, d, B2 * N, device=dev).bfloat16()
        Vt.view(H, 48, B2 * N)[:, d] = 1.0
        Ob = torch.empty(B2 * N, H * d, device=dev, dtype=torch.bfloat16)

        def attn():
            L.check(lib.ldn_attention_bf16(Qb.data_ptr(), 2 * H * slot, Qb.data_ptr() + 2 * H * slot, 2 * H * slot,
                                           Vt.data_ptr(), B2 * N, H * 48, 48, B2, H, N, N, N, d, slot, 2, d ** -0.5,
                                           Ob.data_ptr(), H * d, L.cur_stream()))
        for _ in range(3):
            attn()
        torch.cuda.synchronize()
        a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True)
        reps = 10
        a0.record(stream)
        for _ in range(reps):
            attn()
        a1.record(stream)
        torch.cuda.synchronize()
        attn_ms = a0.elapsed_time(a1) / reps
        attn_flops = 4.0 * B2 * H * N * N * d
        achieved = attn_flops / (attn_ms / 1000.0) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")  # dram__bytes_read + write of one `ncu --set full` capture of this kernel, re-measured per round
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("attn_l0_dram_bytes_per_launch")
            except Exception:
                traffic = None
        hbm_roof = hbm_roofline(eng, dev, bs, lat, peaks)
        step_tflop = (UNET_TFLOP_1024 if args.size == 1024 else UNET_TFLOP_512 if args.size == 512 else None)
        line = {
            "metric": "it/s (UNet sampler steps/sec)", "value": its, "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_string(args.size, bs),
                       "sampler": "dpmpp_2m_cfgpp", "scheduler": "karras", "cfg": cfg, "images_per_gpu": bs,
                       "weights": "seeded synthetic SD1.5 UNet (859.5M params, bf16 in HBM)",
                       "l2": "working set per step (1.72 GB weights + activations) exceeds the 126 MB L2; no flush needed",
                       "cuda_graph": not args.no_graph, "finite": finite},
            "clocks": clk,
            "e2e": {"value": e2e_its, "unit": "it/s", "h2d_bytes_per_step": int(hx.numel() * 4),
                    "d2h_bytes_per_step": int(hout.numel() * 4), "ms_per_step": e2e_ms / K},
            "gpu_launches": int(K * (eng_launches(eng) + 1)),
            "roofline": {"bound": "tensor", "kernel": "a9::attn9_tc_kernel (L0 self-attention, N=%d, d=40, B*H=%d)" % (N, B2 * H),
                         "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                         "traffic": traffic, "peak_source": peaks["source"] + " (burst, kernel timed alone)",
                         "ms_per_launch": attn_ms, "launches_per_step": 5},
        }
        if step_tflop is not None:
            whole = its / world * step_tflop  # `its` counts image-iterations; each is one CFG step of step_tflop
            line["step_roofline"] = {"bound": "tensor", "achieved": whole, "peak": peaks["bf16_sus"], "unit": "TFLOP/s",
                                     "frac": whole / peaks["bf16_sus"], "tflop_per_step": step_tflop,
                                     "peak_source": peaks["source"] + " (sustained)"}
        line["roofline_hbm"] = hbm_roof
    # ---------------------------------------------------------------- sharded product path (every rank takes part)
    sharded = None
    if not args.no_secondary and (world > 1 or not args.no_config3):
        sharded = sharded_metrics(eng, args.size, dev, world, rank)
    if rank == 0:
        if sharded is not None:
            line["sharded" if world > 1 else "batch_configs"] = sharded
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_metrics(eng, args.size, dev)
            if not args.no_gpu_reference:
                line["secondary"].update(gpu_reference_metrics(args.size, its))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.size, lat)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(size: int, lat: int) -> dict:
    """Bounded CPU sample on the box's host cores: one full-size sampler step of the UNMODIFIED reference (baseline/_ref)
    after its own one-step warm-up run; the oracle port if the snapshot is missing."""
    ref = run_reference_script(["--device", "cpu", "--size", str(size), "--steps", "2", "--warmup", "1"], timeout=600)
    if "unavailable" not in ref:
        return {"value": ref["it_per_s"], "unit": "it/s", "cores": ref["threads"], "kind": "reference",
                "sample": f"2 full {size}x{size} sampler steps (+1 warm-up) of the unmodified reference on the CPU "
                          f"(sampling.ksampler + sampling.sample, multiscale off, {ref['unet_dtype']} weights cast per call to "
                          f"{ref['manual_cast']}, torch {ref['torch']}, all host threads)"}
    n, dt, done_w, threads = cpu_oracle_step_time(lat, max_seconds=60.0, max_steps=1, warmup=0)
    return {"value": n / dt, "unit": "it/s", "cores": threads, "kind": "port",
            "sample": f"{n} full {size}x{size} CFG step(s) of the oracle port (torch fp32 CPU, all host threads), no warm-up; "
                      f"reference snapshot unavailable: {ref['unavailable']}"}


def gpu_reference_metrics(size: int, engine_its: float) -> dict:
    """north_star's comparator: the reference's own GPU path (fp16 UNet, torch SDPA) on THIS GPU through its stock sampling
    stack, and the same reference loop with backend.install() (every UNet call answered by the engine through
    model_function_wrapper).  Child processes; 30 steps after a 5-step warm-up run; whole-run wall clock."""
    out = {}
    ref = run_reference_script(["--device", "cuda", "--size", str(size), "--steps", "30", "--warmup", "5"], timeout=900)
    out["gpu_reference"] = ref
    seam = run_reference_script(["--device", "cuda", "--size", str(size), "--steps", "30", "--warmup", "5", "--seam"], timeout=900)
    out["seam"] = seam
    dflt = run_reference_script(["--device", "cuda", "--size", str(size), "--steps", "30", "--warmup", "5", "--default-schedule"],
                                timeout=900)
    out["gpu_reference_default_schedule"] = dflt
    if "it_per_s" in ref:
        out["vs_gpu_reference"] = {"engine_loop": engine_its / ref["it_per_s"],
                                   "through_reference_seam": (seam["it_per_s"] / ref["it_per_s"]) if "it_per_s" in seam else None,
                                   "what": "engine it/s (this line's `value`; and the reference's own loop over the engine seam) "
                                           "/ the unmodified reference's it/s on the same GPU, same workload, multiscale off"}
    return out


def hbm_roofline(eng, dev, bs: int, lat: int, peaks: dict) -> dict:
    """HBM-bound kernels of the step timed alone at their level-0 shape (UNet batch 2*bs, lat x lat, 320 channels):
    achieved = ALGORITHMIC bytes (one read + one write of the tensor) / time, against the measured copy bandwidth."""
    import torch
    from lightdiffusion_next_b200 import _lib as L
    lib = eng.lib
    B2, HW, C = 2 * bs, lat * lat, 320
    x = torch.randn(B2 * HW, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    stream = torch.cuda.current_stream()

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            fn()
        a1.record(stream)
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / reps

    gn_ms = timed(lambda: L.check(lib.ldn_groupnorm_bf16(x.data_ptr(), C, None, 0, B2, HW, 32, 1e-5, gam.data_ptr(), bet.data_ptr(),
                                                         1, y.data_ptr(), L.cur_stream())))
    ln_ms = timed(lambda: L.check(lib.ldn_layernorm_bf16(x.data_ptr(), B2 * HW, C, 1e-5, gam.data_ptr(), bet.data_ptr(), y.data_ptr(),
                                                         L.cur_stream())))
    nbytes = 2.0 * x.numel() * 2
    gn, ln = nbytes / (gn_ms * 1e-3) / 1e9, nbytes / (ln_ms * 1e-3) / 1e9
    # GroupNorm as the UNet runs it behind a ResBlock conv: the conv's epilogue takes the statistics, only the apply
    # kernel follows.  Its cost = (conv + GroupNorm through ldn_conv3x3_groupnorm_bf16) - (the same conv alone).
    import ctypes
    xi = x.view(B2, lat, lat, C)
    w = (torch.randn(C, 3, 3, C, device=dev) / (9 * C) ** 0.5).bfloat16()
    bias = torch.zeros(C, device=dev)
    co = torch.empty_like(xi)
    fused = ctypes.c_int(-1)
    conv_ms = timed(lambda: L.check(lib.ldn_conv3x3_bf16(xi.data_ptr(), w.data_ptr(), B2, lat, lat, C, C, bias.data_ptr(), None, 0, None,
                                                         co.data_ptr(), L.cur_stream())))
    both_ms = timed(lambda: L.check(lib.ldn_conv3x3_groupnorm_bf16(xi.data_ptr(), w.data_ptr(), B2, lat, lat, C, C, bias.data_ptr(), None, 0,
                                                                   None, 1e-5, gam.data_ptr(), bet.data_ptr(), 1, co.data_ptr(),
                                                                   y.data_ptr(), ctypes.addressof(fused), L.cur_stream())))
    inc_ms = max(both_ms - conv_ms, 1e-6)
    after_conv = {"what": "GroupNorm + SiLU behind a 3x3 conv whose epilogue took the statistics (ldn_conv3x3_groupnorm_bf16): "
                          "(conv + GroupNorm) - (conv alone)", "statistics_in_conv_epilogue": bool(fused.value),
                  "conv_ms": conv_ms, "conv_groupnorm_ms": both_ms, "groupnorm_incremental_ms": inc_ms,
                  "achieved": nbytes / (inc_ms * 1e-3) / 1e9, "frac": nbytes / (inc_ms * 1e-3) / 1e9 / peaks["hbm"]}
    return {"bound": "hbm", "kernel": "gn_stats + gn_apply (GroupNorm 32 + SiLU, [%d, %d] bf16)" % (B2 * HW, C), "achieved": gn,
            "after_conv": after_conv,
            "peak": peaks["hbm"], "unit": "GB/s", "frac": gn / peaks["hbm"], "ms_per_launch": gn_ms,
            "algorithmic_bytes": nbytes, "note": "the tensor (%.0f MB) fits the 126 MB L2, back-to-back repeats may hit it" % (x.numel() * 2 / 1e6),
            "layernorm": {"kernel": "layernorm_kernel", "achieved": ln, "frac": ln / peaks["hbm"], "ms_per_launch": ln_ms},
            "peak_source": peaks["source"]}


def sharded_metrics(eng, size: int, dev, world: int, rank: int) -> dict:
    """BASELINE config 3 (bs = 32 in total, 32 / N images per GPU, UNet batch 8 per call) and config 5 (one HiresFix image per
    GPU: 512 -> bislerp x4 latent (ldn_bislerp) -> 2048, 10 steps euler_ancestral_cfgpp / normal / denoise 0.45, VAE decode 2048^2) through
    the product's sharded driver (distributed.sample_sharded: per-rank replay of the full-batch noise, no collective in the
    loop, gather of the final latents on rank 0).  Wall clock bracketed by barrier + device synchronise on every rank, max
    over ranks.  In-run check: rank 0 recomputes the LAST rank's last four images in its own process and compares them with the
    rows it gathered (dpmpp_2m_cfgpp: bit-equal -- the engine is deterministic and batch rows are independent)."""
    import torch
    import torch.distributed as dist
    from lightdiffusion_next_b200 import distributed as D
    from lightdiffusion_next_b200 import sampling as S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    out = {}
    lat = size // 8
    g = torch.Generator().manual_seed(99)
    pos, neg = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    total, steps, per_call = 32, 30, 4
    latent = {"samples": torch.zeros(total, 4, lat, lat)}
    kw = dict(enable_multiscale=False, images_per_call=per_call)
    # warm-up: builds + captures the UNet-batch-8 program (3 steps on the first `world * per_call` images)
    D.sample_sharded(eng, 7, 3, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, {"samples": latent["samples"][:world * per_call]}, **kw)
    barrier()
    t0 = time.perf_counter()
    res = D.sample_sharded(eng, 42, steps, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, latent, **kw)[0]
    barrier()
    dt = tmax(time.perf_counter() - t0)
    ok = None
    if rank == 0:
        full = res["samples"]
        a, b = total - per_call, total  # the last call of the last rank
        chk = S.sample(eng, 42, steps, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, {"samples": latent["samples"][a:b]},
                       noise=S.prepare_noise(latent["samples"], 42)[a:b], enable_multiscale=False,
                       batch_slice=(a, b, total))[0]["samples"]
        ok = bool(torch.equal(chk, full[a:b])) and full.shape[0] == total and bool(torch.isfinite(full).all())
    out["config3"] = {"what": f"SD1.5 {size}x{size} bs=32 sharded over {world} GPU(s) ({total // world} images/GPU, UNet batch {2 * per_call} per call), "
                              f"{steps} steps dpmpp_2m_cfgpp multiscale off, latents gathered on rank 0",
                      "image_it_per_s": total * steps / dt, "seconds": dt, "images": total,
                      "gathered_equals_single_process": ok}
    # ---- config 5: one image per GPU
    if size == 1024:
        n_img = world
        lat5 = {"samples": torch.zeros(n_img, 4, 64, 64)}
        from lightdiffusion_next_b200.latent import latent_upscale
        from lightdiffusion_next_b200.synth import synth_state_dict, vae_decoder_shapes
        eng.load_vae(synth_state_dict(vae_decoder_shapes(), seed=4321))

        def hires_run(seed, steps1, steps2):
            first = D.sample_sharded(eng, seed, steps1, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, lat5, images_per_call=1)[0]
            # LatentUpscale (bislerp, src/Utilities/upscale.py) on the engine's kernel; rank 0 holds the gathered batch
            def upscale(lat):
                return latent_upscale({"samples": lat["samples"].to(dev)}, 2048, 2048, engine=eng)["samples"].cpu()
            if world > 1:
                up = [upscale(first)] if rank == 0 else [None]
                dist.broadcast_object_list(up, 0)
                up = up[0]
            else:
                up = upscale(first)
            second = D.sample_sharded(eng, seed + 1, steps2, 8.0, "euler_ancestral_cfgpp", "normal", pos, neg, {"samples": up},
                                      images_per_call=1, denoise=0.45)[0]
            lo, hi = D.shard_range(n_img, rank, world)
            mine = up[lo:hi] if second is None else second["samples"][lo:hi]
            if world > 1:  # every rank decodes its own image: scatter the final latents back
                mine = D.scatter_rows(second["samples"] if rank == 0 else None, (4, 256, 256), n_img, dev)
            img = eng.vae_decode(mine.to(dev))
            return D.gather_rows(img, n_img)

        hires_run(1, 2, 2)  # warm-up (program build / graph capture at 64^2 and 256^2 latents, VAE 2048^2)
        barrier()
        t0 = time.perf_counter()
        imgs = hires_run(5, 20, 10)
        barrier()
        dt5 = tmax(time.perf_counter() - t0)
        out["config5"] = {"what": f"SD1.5 HiresFix 512 -> 2048 + VAE decode 2048^2, 1 image/GPU on {world} GPU(s): 20 steps dpmpp_2m_cfgpp @512 "
                                  "(reference-default multiscale), bislerp x4 on the device (ldn_bislerp), 10 steps euler_ancestral_cfgpp/normal/denoise 0.45 "
                                  "@2048, decode, images gathered on rank 0",
                          "seconds": dt5, "images_per_s": n_img / dt5,
                          "finite": bool(torch.isfinite(imgs).all()) if rank == 0 else None,
                          "shape": list(imgs.shape) if rank == 0 else None}
    return out


def run_flux(args):
    """BASELINE config 4: Flux.1-dev txt2img 1024x1024 bs=1 on one B200 -- the DiT part of a sampler step as the reference
    runs it: two forwards per step (cond + uncond rows, `disable_cfg1_optimization`, samplers.py:517-520), CONST
    denoising x - v * sigma (sampling.py:100-127), CFG lerp and an Euler update. Seeded synthetic 11.9 G-parameter weights
    generated on the device. Same JSON contract; no CPU arm (the fp32 weights alone would need 48 GB of host memory)."""
    import zlib

    import torch

    from lightdiffusion_next_b200 import _lib as L
    from lightdiffusion_next_b200 import flux as FX
    from lightdiffusion_next_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    cfg = FX.FLUX_DEV
    eng = Engine(max_rows=2, max_h=8, max_w=8, use_graph=not args.no_graph, device=dev)
    batch, nbytes = {}, 0
    for k, shp in FX.flux_shapes(cfg).items():
        g = torch.Generator(device=dev).manual_seed(zlib.crc32(k.encode()) & 0x7FFFFFFF)
        if len(shp) > 1:
            w = torch.randn(shp, generator=g, device=dev, dtype=torch.bfloat16)
            w = (w * ((0.5 if any(t in k for t in ("proj.", "mlp.2", "linear2")) else 1.0) / shp[1] ** 0.5)).to(torch.bfloat16)
        elif k.endswith(".scale"):
            w = 1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)
        else:
            w = 0.02 * torch.randn(shp, generator=g, device=dev)
        batch[k] = w
        nbytes += w.numel() * w.element_size()
        if nbytes > 2 << 30:
            eng.load_weights(4, batch)
            batch, nbytes = {}, 0
            torch.cuda.empty_cache()
    if batch:
        eng.load_weights(4, batch)
    torch.cuda.empty_cache()
    lat, n_txt, W, K = args.size // 8, 256, max(3, args.warmup), args.steps
    g = torch.Generator().manual_seed(1234)
    ctx = torch.randn(2, n_txt, cfg["context_in_dim"], generator=g).to(dev)
    y = torch.randn(2, cfg["vec_in_dim"], generator=g).to(dev)
    guid = torch.full((2,), 3.5, device=dev)
    x0 = torch.randn(1, 16, lat, lat, generator=g).to(dev)
    sig = torch.linspace(1.0, 0.0, 31)
    out, den = torch.empty_like(x0), torch.empty_like(x0)
    stream = torch.cuda.current_stream()

    def one_step(i, x):
        j = i % 29
        s, s_next = float(sig[j]), float(sig[j + 1])
        v = eng.flux_forward(x.expand(2, -1, -1, -1).contiguous(), torch.full((2,), s, device=dev), ctx, y, guid)
        d = x - v * s  # CONST.calculate_denoised, rows: uncond first
        eng.cfg_step(x, d[0:1].contiguous(), d[1:2].contiguous(), 1.0, 1, c0=s_next - s, c1=0.0, c2=s, x_out=out, denoised_out=den)
        return out.clone()

    x = x0.clone()
    for i in range(W):
        x = one_step(i, x)
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        x = one_step(i, x)
    e1.record(stream)
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    hx = torch.empty(1, 16, lat, lat).pin_memory()
    hx.copy_(x0.cpu())
    hout = torch.empty_like(hx).pin_memory()
    dx = torch.empty_like(x0)

    def e2e_step(i):
        dx.copy_(hx, non_blocking=True)
        yv = one_step(i, dx)
        hout.copy_(yv, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hx.copy_(hout)

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # roofline of the dominant non-GEMM kernel: the joint attention (attention6, d = 128) timed alone
    lib, H, N, d = eng.lib, cfg["num_heads"], (lat // 2) ** 2 + n_txt, 128
    Np = (N + 15) // 16 * 16
    QK = torch.randn(Np, 2 * H * d, device=dev).bfloat16()
    Vt = torch.randn(H * d, Np, device=dev).bfloat16()
    Ob = torch.empty(N, H * d, device=dev, dtype=torch.bfloat16)

    def attn():
        L.check(lib.ldn_attention_bf16(QK.data_ptr(), 2 * H * d, QK.data_ptr() + 2 * H * d, 2 * H * d, Vt.data_ptr(), Np, H * d, 0,
                                       1, H, N, N, Np, d, d, 0, d ** -0.5, Ob.data_ptr(), H * d, L.cur_stream()))
    for _ in range(3):
        attn()
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    for _ in range(10):
        attn()
    a1.record(stream)
    torch.cuda.synchronize()
    attn_ms = a0.elapsed_time(a1) / 10
    achieved = 4.0 * H * N * N * d / (attn_ms / 1000.0) / 1e12
    C, M = cfg["hidden_size"], int(cfg["hidden_size"] * cfg["mlp_ratio"])
    Ni = (lat // 2) ** 2
    fwd_tflop = (cfg["depth"] * (2 * N * C * (4 * C + 2 * M) + 4 * N * N * C)
                 + cfg["depth_single_blocks"] * (2 * N * C * (3 * C + M) + 2 * N * (C + M) * C + 4 * N * N * C)) / 1e12
    its = K / (ms / 1000.0)
    line = {
        "metric": "it/s (Flux DiT sampler steps/sec, 2 forwards per step)", "value": its, "unit": "it/s", "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"Flux.1-dev txt2img {args.size}x{args.size} bs=1 (BASELINE config 4): DiT forward x 2 (cond + uncond) + "
                               "CONST denoise + CFG + Euler update; text encoders and VAE not included",
                   "tokens": f"{Ni} image + {n_txt} text", "weights": "seeded synthetic Flux.1-dev layout (11.9 G params, bf16 in HBM)",
                   "l2": "24 GB of weights stream per forward; no flush needed", "cuda_graph": not args.no_graph,
                   "finite": bool(torch.isfinite(x).all().item())},
        "clocks": clk,
        "e2e": {"value": K / e2e_s, "unit": "it/s", "h2d_bytes_per_step": int(hx.numel() * 4),
                "d2h_bytes_per_step": int(hout.numel() * 4), "ms_per_step": e2e_s * 1000.0 / K},
        "gpu_launches": int(K * (2 * 681 + 1)),
        "roofline": {"bound": "tensor", "kernel": f"a6::attn6_tc_kernel<128> (joint attention, N={N}, d=128, H={H})",
                     "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                     "traffic": None, "peak_source": peaks["source"] + " (burst, kernel timed alone)", "ms_per_launch": attn_ms,
                     "launches_per_step": 2 * (cfg["depth"] + cfg["depth_single_blocks"])},
        "step_roofline": {"bound": "tensor", "achieved": its * 2 * fwd_tflop, "peak": peaks["bf16_sus"], "unit": "TFLOP/s",
                          "frac": its * 2 * fwd_tflop / peaks["bf16_sus"], "tflop_per_step": 2 * fwd_tflop,
                          "peak_source": peaks["source"] + " (sustained)"},
        "cpu_baseline": None,
    }
    if not args.no_secondary:
        line["secondary"] = {"flux_image_e2e": flux_image_e2e(eng, dev, args.size)}
    print(json.dumps(line), flush=True)


def flux_image_e2e(eng, dev, size: int) -> dict:
    """BASELINE config 4 end to end, the Flux branch of pipeline() (src/user/pipeline.py:218-275) on one engine:
    CLIP-L pooled vector + T5-XXL states (CLIPTextEncodeFlux) -> 20 steps euler_cfgpp / beta, cfg 1, guidance 3
    (two DiT forwards per step + the sampler's half-resolution calls at steps 2-3) -> 16-channel VAE decode -> image on
    the host.  Seeded synthetic weights of the real sizes (T5-XXL encoder 4.76 G, Flux.1-dev 11.9 G, CLIP-L, VAE), generated
    on the device; wall clock with a device synchronise on both sides, second run (the first builds / captures the programs)."""
    import zlib

    import torch

    from lightdiffusion_next_b200 import t5 as T5H
    from lightdiffusion_next_b200.pipeline import EMPTY_TOKENS, FluxPipeline
    from lightdiffusion_next_b200.synth import clip_shapes, synth_state_dict, vae_decoder_shapes

    t_load = time.perf_counter()
    batch, nbytes = {}, 0
    shapes = T5H.t5_shapes()
    for k, shp in shapes.items():
        g = torch.Generator(device=dev).manual_seed(zlib.crc32(k.encode()) & 0x7FFFFFFF)
        if k == "shared.weight":
            w = torch.randn(shp, generator=g, device=dev, dtype=torch.bfloat16)
        elif k.endswith("relative_attention_bias.weight"):
            w = torch.randn(shp, generator=g, device=dev)
        elif len(shp) > 1:
            w = torch.randn(shp, generator=g, device=dev, dtype=torch.bfloat16) * ((0.125 if ".q." in k else 0.5 if (".o." in k or ".wo." in k) else 1.0) / shp[1] ** 0.5)
        else:
            w = (1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)).float()
        batch[k] = w
        nbytes += w.numel() * w.element_size()
        if nbytes > 2 << 30:
            eng.load_weights(5, batch)
            batch, nbytes = {}, 0
            torch.cuda.empty_cache()
    if batch:
        eng.load_weights(5, batch)
    eng._t5_width = shapes["shared.weight"][1]
    eng.load_clip(synth_state_dict(clip_shapes(), seed=777))
    vsd = {k: v for k, v in synth_state_dict(vae_decoder_shapes(z=16), seed=9753).items() if not k.startswith("post_quant_conv")}
    eng.load_vae(vsd)
    torch.cuda.empty_cache()
    t_load = time.perf_counter() - t_load
    pipe = FluxPipeline(eng)
    clip_tokens = [[(49406, 1.0)] + [(1000 + i, 1.0) for i in range(20)] + [(49407, 1.0)] * 56]
    t5_tokens = [T5H.pad_tokens([100 + 7 * i for i in range(40)])]  # 40 ids + end token, padded to the reference's 256-token minimum
    out = {}
    for run in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cond = pipe.encode(clip_tokens, t5_tokens, guidance=3.0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        lat = pipe.sample(cond, pipe.zero_out(cond), size, size, batch=1, seed=3, steps=20)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        img = pipe.decode(lat)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        out = {"seconds_per_image": t3 - t0, "images_per_s": 1.0 / (t3 - t0), "text_encode_s": t1 - t0, "sample_s": t2 - t1,
               "sampler_it_per_s": 20 / (t2 - t1), "vae_decode_and_d2h_s": t3 - t2, "t5_tokens": len(t5_tokens[0]),
               "shape": list(img.shape), "finite": bool(torch.isfinite(img).all().item())}
    out["what"] = (f"Flux.1-dev {size}x{size} bs=1: CLIP-L + T5-XXL encode -> 20 steps euler_cfgpp / beta (2 DiT forwards per step + the "
                   "half-resolution calls of steps 2-3) -> 16-channel VAE decode -> image to host; synthetic weights of the real sizes")
    out["weight_setup_s"] = t_load
    return out


def secondary_metrics(eng, size: int, dev) -> dict:
    """SURVEY.md 8(d) secondary numbers, all through the public host API and wall-clocked with a device sync:
    whole 30-step KSampler runs with multiscale off and with the reference-default multiscale schedule (which runs
    8 of the 30 steps at half resolution), and end-to-end images/s of the pipe() surface (CLIP encode of two prompts +
    30 sampler steps + VAE decode + image D2H)."""
    import torch
    from lightdiffusion_next_b200 import sampling as S
    from lightdiffusion_next_b200.pipeline import Pipeline, EMPTY_TOKENS
    from lightdiffusion_next_b200.synth import synth_state_dict, vae_decoder_shapes, clip_shapes
    out = {}
    g = torch.Generator().manual_seed(99)
    pos = torch.randn(1, 77, 768, generator=g)
    neg = torch.randn(1, 77, 768, generator=g)
    latent = {"samples": torch.zeros(1, 4, size // 8, size // 8)}
    for name, ms_flag in (("full_run_multiscale_off", False), ("full_run_multiscale_default", True)):
        S.sample(eng, 42, 30, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, latent, enable_multiscale=ms_flag)  # warm-up / graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.sample(eng, 42, 30, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, latent, enable_multiscale=ms_flag)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"it_per_s": 30 / dt, "seconds": dt, "steps": 30}
    eng.load_vae(synth_state_dict(vae_decoder_shapes(), seed=4321))
    eng.load_clip(synth_state_dict(clip_shapes(), seed=777))
    pipe = Pipeline(eng)
    toks = [[(49406, 1.0)] + [(1000 + i, 1.0) for i in range(20)] + [(49407, 1.0)] * 56]
    pipe(toks, width=size, height=size, steps=30, sampler_name="dpmpp_2m_cfgpp", scheduler="karras")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    img = pipe(toks, width=size, height=size, steps=30, sampler_name="dpmpp_2m_cfgpp", scheduler="karras")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["pipe_e2e"] = {"images_per_s": 1.0 / dt, "seconds_per_image": dt, "finite": bool(torch.isfinite(img).all().item()),
                       "what": "CLIP encode (positive + empty negative) + 30 steps dpmpp_2m_cfgpp (reference-default multiscale) "
                               "+ VAE decode + image to host, %dx%d, synthetic weights" % (size, size)}
    return out


def eng_launches(eng) -> int:
    """Kernels per UNet program run, counted by the engine when it built the program."""
    return int(eng.lib.ldn_unet_last_launches(eng.h))


if __name__ == "__main__":
    main()
