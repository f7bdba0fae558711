"""CPU emulation of the encoder's roundings with and without the LayerNorm folding (DESIGN.md
section 2): 16-bit operands, fp16 residual stream, fp32 accumulation, against the fp32 oracle.
  plain     LN output rounded to the operand type, W rounded            (the LayerNorm kernels)
  fold      x W'^T with W' = centred gamma (.) W rounded to the operand type, r applied after
  fold_w11  the same with W' kept at fp16 precision
Prints the embedding distance to the oracle and the largest |mean| / std of a stream row per layer
(the quantity that scales the folded weights' residual row-sum error).  Test infrastructure: it
imports the oracle and never touches the GPU.

    python tools/ln_fold_study.py vit_tiny_patch16_224 bf16
"""
import numpy as np, torch, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from oracle.vit import init_vit_weights, VIT_CONFIGS
from oracle.crop import crop_resize
from vitparticlefiltertracker_b200.synthetic import make_clip
torch.manual_seed(0)
model=sys.argv[1] if len(sys.argv)>1 else 'vit_tiny_patch16_224'
opd=torch.float16 if (len(sys.argv)>2 and sys.argv[2]=='fp16') else torch.bfloat16
c=VIT_CONFIGS[model]; d,depth,heads=c['dim'],c['depth'],c['heads']; hd=64
sd=init_vit_weights(model)
frames,gt=make_clip("box",224,224,3,seed=0,target=(48,48))
rng=np.random.default_rng(0); n=12
boxes=np.stack([rng.uniform(60,160,n),rng.uniform(60,160,n),rng.uniform(30,70,n),rng.uniform(30,70,n)],1).astype(np.float32)
crops=torch.from_numpy(crop_resize(frames[1],boxes))
q=lambda t,dt: t.to(dt).float()
def run(mode):
    # mode: 'fp32', 'plain' (LN->round), 'fold', 'fold_w8' (W' rounded bf16)
    x=crops
    if mode=='fp32':
        x=F.conv2d(x,sd["patch_embed.proj.weight"],sd["patch_embed.proj.bias"],stride=16)
    else:
        x=F.conv2d(q(x,opd),q(sd["patch_embed.proj.weight"],opd),sd["patch_embed.proj.bias"],stride=16)
    x=x.flatten(2).transpose(1,2)
    x=torch.cat([sd["cls_token"].expand(n,-1,-1),x],1)+sd["pos_embed"]
    st=lambda t: t if mode=='fp32' else q(t,torch.float16)
    x=st(x)
    ratios=[]
    def lnlin(x,nk,lk):
        g,b=sd[nk+".weight"],sd[nk+".bias"]; W,B=sd[lk+".weight"],sd[lk+".bias"]
        if mode=='fp32': return F.linear(F.layer_norm(x,(d,),g,b,1e-6),W,B)
        if mode=='plain': return F.linear(q(F.layer_norm(x,(d,),g,b,1e-6),opd),q(W,opd),B)
        wf=(W.double()*g.double()[None]); wf=(wf-wf.mean(1,keepdim=True)).float()
        wf=q(wf,opd) if mode=='fold' else q(wf,torch.float16)
        bf=(B.double()+W.double()@b.double()).float()
        mu=x.mean(-1,keepdim=True); var=(x*x).mean(-1,keepdim=True)-mu*mu
        r=torch.rsqrt(var.clamp_min(0)+1e-6)
        ratios.append(float((mu.abs()/var.sqrt()).max()))
        return r*F.linear(x,wf)+bf
    for i in range(depth):
        p=f"blocks.{i}."
        qkv=lnlin(x,p+"norm1",p+"attn.qkv")
        if mode!='fp32': qkv=q(qkv,opd)
        qkv=qkv.reshape(n,197,3,heads,hd).permute(2,0,3,1,4)
        att=((qkv[0]@qkv[1].transpose(-2,-1))/8).softmax(-1)
        if mode!='fp32': att=q(att,opd)
        o=(att@qkv[2]).transpose(1,2).reshape(n,197,d)
        if mode!='fp32': o=q(o,opd)
        x=st(x+F.linear(o,sd[p+"attn.proj.weight"] if mode=='fp32' else q(sd[p+"attn.proj.weight"],opd),sd[p+"attn.proj.bias"]))
        h=F.gelu(lnlin(x,p+"norm2",p+"mlp.fc1"))
        if mode!='fp32': h=q(h,opd)
        x=st(x+F.linear(h,sd[p+"mlp.fc2.weight"] if mode=='fp32' else q(sd[p+"mlp.fc2.weight"],opd),sd[p+"mlp.fc2.bias"]))
    out=F.layer_norm(x[:,0],(d,),sd["norm.weight"],sd["norm.bias"],1e-6)
    return out,ratios
ref,_=run('fp32')
for m in ('plain','fold','fold_w11'):
    o,r=run(m)
    rel=float((o-ref).norm()/ref.norm()); cos=float((1-F.cosine_similarity(o,ref,dim=1)).max())
    print(model,opd,m,'rel-L2 %.3e  max(1-cos) %.3e'%(rel,cos),'max |mu|/sigma by layer:',[round(v,2) for v in r[:6]],'...',[round(v,2) for v in r[-2:]] if r else '')
