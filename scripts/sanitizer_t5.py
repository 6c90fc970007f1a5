"""Small T5 encode for compute-sanitizer (memcheck / racecheck / synccheck): the tiny configuration of the parity test
(width 128, 2 heads of 64, 3 blocks), 40 and 300 tokens, eager launches.  First thing to run on a GPU in round 2:
  compute-sanitizer --tool memcheck python scripts/sanitizer_t5.py"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lightdiffusion_next_b200.engine import Engine
from test_t5_cpu import t5_tiny_sd
sd, gold = t5_tiny_sd()
eng = Engine(max_rows=2, max_h=16, max_w=16, use_graph=False)
eng.load_t5(sd)
for name in "ab":
    out = eng.t5_encode(gold["ids_" + name]); torch.cuda.synchronize()
    ref = gold["out_" + name]
    print(name, "finite", torch.isfinite(out).all().item(), "rel", float((out.cpu() - ref).norm() / ref.norm()))
