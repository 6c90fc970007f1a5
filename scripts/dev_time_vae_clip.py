import sys, os, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from oracle import sd15_oracle as O
eng = Engine(max_rows=2, max_h=128, max_w=128)
eng.load_vae(O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321))
eng.load_clip(O.synth_state_dict(O.clip_param_shapes(), seed=777))
ids = torch.randint(0, 49408, (3, 77))
for _ in range(3): eng.clip_encode(ids)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.clip_encode(ids)
e1.record(); torch.cuda.synchronize()
print(f"clip encode 3x77: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
for hw in (64, 128):
    z = torch.randn(1, 4, hw, hw).cuda()
    for _ in range(2): img = eng.vae_decode(z)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): img = eng.vae_decode(z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tf = {64: 2.515, 128: 10.47}[hw]
    print(f"vae decode {hw*8}^2: {ms:.2f} ms = {tf/ms*1000:.0f} TFLOP/s  finite={torch.isfinite(img).all().item()} mean={img.mean().item():.3f}", flush=True)
sd = dict(O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)); sd.update(O.synth_state_dict(O.vae_encoder_param_shapes(), seed=2468))
eng.load_vae(sd)
for px in (512, 1024):
    img = torch.rand(1, px, px, 3).cuda()
    for _ in range(2): m = eng.vae_encode_moments(img)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): m = eng.vae_encode_moments(img)
    e1.record(); torch.cuda.synchronize()
    print(f"vae encode {px}^2: {e0.elapsed_time(e1)/3:.2f} ms finite={torch.isfinite(m).all().item()}", flush=True)
if os.environ.get("LDN_PROFILE"):
    z = torch.randn(1, 4, 128, 128).cuda(); eng.vae_decode(z)
