"""Config-5 sized pieces on one GPU: VAE decode / encode at 2048^2 (finite + time) and one UNet CFG step at 256x256 latent."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200.synth import synth_state_dict, vae_decoder_shapes, vae_encoder_shapes
eng = Engine(max_rows=2, max_h=256, max_w=256)
sd = dict(synth_state_dict(vae_decoder_shapes(), seed=4321)); sd.update(synth_state_dict(vae_encoder_shapes(), seed=2468))
eng.load_vae(sd)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
z = torch.randn(1, 4, 256, 256).cuda()
for _ in range(2): img = eng.vae_decode(z)
torch.cuda.synchronize(); e0.record(); img = eng.vae_decode(z); e1.record(); torch.cuda.synchronize()
print(f"vae decode 2048^2: {e0.elapsed_time(e1):.1f} ms ({48.5/e0.elapsed_time(e1)*1000:.0f} TFLOP/s) finite={torch.isfinite(img).all().item()} shape={tuple(img.shape)}", flush=True)
px = torch.rand(1, 2048, 2048, 3).cuda()
for _ in range(2): m = eng.vae_encode_moments(px)
torch.cuda.synchronize(); e0.record(); m = eng.vae_encode_moments(px); e1.record(); torch.cuda.synchronize()
print(f"vae encode 2048^2: {e0.elapsed_time(e1):.1f} ms finite={torch.isfinite(m).all().item()} shape={tuple(m.shape)}", flush=True)
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9, "(torch side only)")
