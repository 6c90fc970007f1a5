import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); dev = "cuda"; torch.manual_seed(0)
def rel(a,b): return ((a.float()-b.float()).norm()/b.float().norm()).item()
def rep(name, fn, n=4):
    outs = [fn().clone() for _ in range(n)]
    print(f"{name}: bit-equal={all(torch.equal(outs[0], o) for o in outs)} maxrel={max(rel(o, outs[0]) for o in outs):.3e}", flush=True)
# gemm with in-place residual
for (M,N,K) in ((2048,320,320),(2048,320,1280),(512,1280,5120),(32768,320,320)):
    A = torch.randn(M,K,device=dev).bfloat16(); W=(torch.randn(N,K,device=dev)/K**0.5).bfloat16(); b=torch.randn(N,device=dev)
    X0 = torch.randn(M,N,device=dev).bfloat16()
    def f():
        X = X0.clone()
        L.check(lib.ldn_gemm_bf16(A.data_ptr(),K,K,None,0,0,W.data_ptr(),M,N,b.data_ptr(),None,0,0,X.data_ptr(),N,X.data_ptr(),N,None,0,0,0,0,L.cur_stream()))
        return X
    rep(f"gemm inplace-res {M}x{N}x{K}", f)
# qk gemm with head slots
M,C,d,slot=2048,320,40,64
A = torch.randn(M,C,device=dev).bfloat16(); W=(torch.randn(2*C,C,device=dev)/C**0.5).bfloat16()
def f():
    out = torch.zeros(M, 16*slot, device=dev, dtype=torch.bfloat16)
    L.check(lib.ldn_gemm_bf16(A.data_ptr(),C,C,None,0,0,W.data_ptr(),M,2*C,None,None,0,0,None,0,out.data_ptr(),16*slot,None,0,d,slot,0,L.cur_stream()))
    return out
rep("gemm qk slots", f)
# vt gemm (swapped)
def f():
    out = torch.zeros(C, M, device=dev, dtype=torch.bfloat16)
    L.check(lib.ldn_gemm_bf16(W.data_ptr(),C,C,None,0,0,A.data_ptr(),C,M,None,None,0,0,None,0,out.data_ptr(),M,None,0,0,0,0,L.cur_stream()))
    return out
rep("gemm vt", f)
# conv
for (B,H,Wd,Cin,Cout) in ((2,32,32,320,320),(2,8,8,1280,1280),(2,4,4,2560,1280),(2,64,64,320,320)):
    x = torch.randn(B,H,Wd,Cin,device=dev).bfloat16(); w=(torch.randn(Cout,3,3,Cin,device=dev)/(9*Cin)**0.5).bfloat16(); b=torch.randn(Cout,device=dev)
    def f():
        out = torch.zeros(B,H,Wd,Cout,device=dev,dtype=torch.bfloat16)
        L.check(lib.ldn_conv3x3_bf16(x.data_ptr(),w.data_ptr(),B,H,Wd,Cin,Cout,b.data_ptr(),None,0,None,out.data_ptr(),L.cur_stream()))
        return out
    rep(f"conv {B}x{H}x{Wd} {Cin}->{Cout}", f)
# attention
for (B,Hh,Nq,Nk,d) in ((2,8,1024,1024,40),(2,8,256,256,80),(2,8,64,64,160),(2,8,16,16,160),(2,8,1024,77,40),(2,8,4096,4096,40)):
    slot=(d+63)//64*64; nk_pad=(Nk+127)//128*128 if Nk%8 else Nk
    Qb=torch.zeros(B*Nq,Hh*slot,device=dev,dtype=torch.bfloat16); Kb=torch.zeros(B*nk_pad,Hh*slot,device=dev,dtype=torch.bfloat16)
    Qb.view(B,Nq,Hh,slot)[...,:d]=torch.randn(B,Nq,Hh,d,device=dev).bfloat16(); Kb.view(B,nk_pad,Hh,slot)[:,:Nk,:,:d]=torch.randn(B,Nk,Hh,d,device=dev).bfloat16()
    Vt=torch.zeros(Hh*d,B*nk_pad,device=dev,dtype=torch.bfloat16); Vt.view(Hh,d,B,nk_pad)[...,:Nk]=torch.randn(Hh,d,B,Nk,device=dev).bfloat16()
    def f():
        out=torch.zeros(B*Nq,Hh*d,device=dev,dtype=torch.bfloat16)
        L.check(lib.ldn_attention_bf16(Qb.data_ptr(),Hh*slot,Kb.data_ptr(),Hh*slot,Vt.data_ptr(),B*nk_pad,Hh*d,0,B,Hh,Nq,Nk,nk_pad,d,slot,0,d**-0.5,out.data_ptr(),Hh*d,L.cur_stream()))
        return out
    rep(f"attn B{B} H{Hh} {Nq}x{Nk} d{d}", f)
# groupnorm / layernorm
x0=torch.randn(2,1024,640,device=dev).bfloat16(); x1=torch.randn(2,1024,320,device=dev).bfloat16(); g=torch.randn(960,device=dev); bt=torch.randn(960,device=dev)
def f():
    out=torch.zeros(2,1024,960,device=dev,dtype=torch.bfloat16)
    L.check(lib.ldn_groupnorm_bf16(x0.data_ptr(),640,x1.data_ptr(),320,2,1024,32,1e-5,g.data_ptr(),bt.data_ptr(),1,out.data_ptr(),L.cur_stream()))
    return out
rep("groupnorm concat", f)
x=torch.randn(2048,320,device=dev).bfloat16(); g=torch.randn(320,device=dev); bt=torch.randn(320,device=dev)
def f():
    out=torch.zeros_like(x)
    L.check(lib.ldn_layernorm_bf16(x.data_ptr(),2048,320,1e-5,g.data_ptr(),bt.data_ptr(),out.data_ptr(),L.cur_stream()))
    return out
rep("layernorm", f)
