import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, sys, os
from emdr2_b200 import ops
from emdr2_b200.packed import PackedBatch
DEV="cuda:0"
for (b,s,heads) in [(1,257,1),(1,384,1),(1,385,1),(2,512,1),(4,300,2)]:
    h=heads*64
    lens=np.full(b,s)
    pb=PackedBatch(lens,s,heads,torch.device(DEV))
    g=torch.Generator().manual_seed(1)
    qkv=(torch.randn(pb.T,3*h,generator=g)*0.8).to(torch.bfloat16).to(DEV)
    out=ops.attention_varlen(qkv[:,:h],qkv[:,h:2*h],qkv[:,2*h:],heads,pb.items,pb.n_items,scale=0.125)
    for i,(c0,c1) in enumerate(zip(pb.cu[:-1],pb.cu[1:])):
        q=qkv[c0:c1,:h].float().view(-1,heads,64).permute(1,0,2); k=qkv[c0:c1,h:2*h].float().view(-1,heads,64).permute(1,0,2); v=qkv[c0:c1,2*h:].float().view(-1,heads,64).permute(1,0,2)
        w=(torch.softmax(q@k.transpose(1,2)*0.125,-1)@v).permute(1,0,2).reshape(-1,h)
        err=(out[c0:c1].float()-w).abs().amax(dim=1)
        badrows=(err>0.05).nonzero().flatten().tolist()
        print((b,s,heads),"seq",i,"bad rows",len(badrows), badrows[:6], badrows[-3:], "maxerr",round(err.max().item(),3))
