"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): UNet step at 32x32 latent, VAE 16x16, CLIP."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes
from oracle import sd15_oracle as O
eng = Engine(max_rows=2, max_h=32, max_w=32, use_graph=False)
eng.load_unet(synth_state_dict(unet_shapes()))
vsd = dict(O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)); vsd.update(O.synth_state_dict(O.vae_encoder_param_shapes(), seed=2468))
eng.load_vae(vsd)
eng.load_clip(O.synth_state_dict(O.clip_param_shapes(), seed=777))
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, 32, 32, generator=g).cuda(); sigma = torch.tensor([3.0, 0.5]).cuda()
eng.set_context(torch.randn(2, 77, 768, generator=g).cuda())
out = eng.denoise(x, sigma); torch.cuda.synchronize()
img = eng.vae_decode(torch.randn(1, 4, 16, 16, generator=g)); torch.cuda.synchronize()
mom = eng.vae_encode_moments(torch.rand(1, 64, 72, 3, generator=g)); torch.cuda.synchronize()
pen, last = eng.clip_encode(torch.randint(0, 49408, (2, 77), generator=g)); torch.cuda.synchronize()
print("finite", torch.isfinite(out).all().item(), torch.isfinite(img).all().item(), torch.isfinite(mom).all().item(), torch.isfinite(pen).all().item())
