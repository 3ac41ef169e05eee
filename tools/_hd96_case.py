import sys, torch
sys.path.insert(0, "/root/repo")
from visper_lm_b200 import ops
B,H,KVH,sq,sk,causal,hd = [int(x) for x in sys.argv[1:8]]
causal=bool(causal)
g=torch.Generator().manual_seed(91)
qw, kw = H*hd, KVH*hd
q=torch.randn(B*sq,qw,generator=g).to(torch.bfloat16).cuda()
kv=torch.randn(B*sk,2*kw,generator=g).to(torch.bfloat16).cuda()
do=torch.randn(B*sq,qw,generator=g).to(torch.bfloat16).cuda()
k,v=kv[:,:kw],kv[:,kw:]
o,lse=ops.attn_fwd(q,k,v,B,H,KVH,sq,sk,hd,hd**-0.5,causal)
dq=torch.zeros_like(q); dkv=torch.zeros_like(kv)
try:
    ops.attn_bwd(q,k,v,o,do,lse,dq,dkv[:,:kw],dkv[:,kw:],B,H,KVH,sq,sk,hd,hd**-0.5,causal)
    torch.cuda.synchronize()
    print(sys.argv[1:8], "ok", round(dq.float().abs().mean().item(),4), round(dkv.float().abs().mean().item(),4))
except Exception as e:
    print(sys.argv[1:8], "FAIL", str(e)[:80])
