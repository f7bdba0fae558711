import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitparticlefiltertracker_b200 import ops
dev=torch.device('cuda'); n=4096; H=12; D=768
for T in (197, 170, 145, 101, 65):
    qkv=(torch.randn(n*T,3*D,device=dev)*1.0).to(torch.bfloat16)
    for _ in range(3): ops.attention(qkv,n,H,T)
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.attention(qkv,n,H,T)
    e.record(); torch.cuda.synchronize()
    print('attention T=%d: %.3f ms'%(T, s.elapsed_time(e)/20))
