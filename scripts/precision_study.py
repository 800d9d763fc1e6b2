"""Precision study (CPU, not product code): TF32 operand-rounding emulation on tensor-regime inputs (see precision_raw.py)."""
import importlib, sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
synth = importlib.import_module('iros20-6d-pose-tracking_b200.synth')
import se3_oracle as O
torch.set_num_threads(8)

def rna_tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)
def rn_bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def ident(x): return x

def fold(sd, conv, bn):
    w = sd[conv+'.weight'].double(); b = sd[conv+'.bias'].double()
    g = sd[bn+'.weight'].double(); beta = sd[bn+'.bias'].double(); mu = sd[bn+'.running_mean'].double(); var = sd[bn+'.running_var'].double()
    s = g/torch.sqrt(var+1e-5)
    return (w*s[:,None,None,None]).float(), ((b-mu)*s+beta).float()

def selu(x): return F.selu(x)

def run(sd, A, B, rnd, dt=torch.float64):
    def conv(x, w, b, stride, pad):
        y = F.conv2d(rnd(x).to(dt), rnd(w).to(dt), None, stride=stride, padding=pad).float()
        return y + b[None,:,None,None]
    def cbr(x, p, stride, pad):
        w,b = fold(sd, p+'.0', p+'.1'); return rnd(selu(conv(x,w,b,stride,pad)))
    def block(x, p):
        w1,b1 = fold(sd,p+'.conv1',p+'.bn1'); w2,b2 = fold(sd,p+'.conv2',p+'.bn2')
        t = rnd(F.relu(conv(x,w1,b1,1,1)))
        return rnd(F.relu(conv(t,w2,b2,1,1)+x))
    a = cbr(rnd(A),'convA1',2,3); a = F.max_pool2d(a,3,2,1); a = block(a,'convA2')
    b = cbr(rnd(B),'convB1',2,3); b = F.max_pool2d(b,3,2,1); b = block(b,'convB2'); b = block(b,'convB3')
    ab = torch.cat((a,b),1); ab = cbr(ab,'convAB1',2,1); ab = block(ab,'convAB2')
    outs=[]
    for h in ('trans','rot'):
        x = cbr(ab,h+'_conv1',2,1); x = block(x,h+'_conv2')
        x = x.mean((2,3))
        outs.append(torch.tanh(F.linear(x, sd[h+'_out.0.weight'], sd[h+'_out.0.bias'])))
    return torch.cat(outs,1)

n = int(sys.argv[1]) if len(sys.argv)>1 else 8
for seed in (0,1):
    sd = synth.make_state_dict(seed)
    A,B = synth.tensor_pairs(n, seed=seed)
    with torch.no_grad():
        ref = O.forward(sd,A,B); ref = torch.cat((ref['trans'],ref['rot']),1)
        for name,r in (('fp32-fold',ident),('tf32',rna_tf32),('bf16',rn_bf16)):
            out = run(sd,A,B,r)
            err = (out-ref).abs(); tol = 1e-4+1e-3*ref.abs()
            print(seed, name, 'max abs err %.3e'%err.max().item(), 'max err/tol %.3f'%(err/tol).max().item(), 'ref absmax %.3f'%ref.abs().max().item())
