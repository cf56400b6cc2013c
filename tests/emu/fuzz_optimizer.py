"""Randomised runs of the host-emulated optimizer kernels: 1-3 tensors of awkward sizes (1 ... 40000 elements, multi-chunk),
element-offset (unaligned -> scalar path) views, bf16 / fp32 moments, fp32 master parameters, fused gradient scale, shuffled
chunk order; AdamW vs oracle/adamw_oracle.py and the global norm / clip coefficient.  Needs the library built by
fuzz_kernels.py (EMU_LIB or /tmp/libemu_fuzz.so).  Not collected by pytest.  End-of-round-1 run: 880 cases, 0 failures."""
import ctypes, math, random, sys, time
import torch
sys.path.insert(0, "/root/repo")
from oracle import adamw_oracle as O
lib=ctypes.CDLL("/tmp/libemu_fuzz.so")
p,i32,i64,f32=ctypes.c_void_p,ctypes.c_int,ctypes.c_int64,ctypes.c_float
lib.emu_adamw.argtypes=[p,p,i32,i32,f32,f32,f32,f32,f32,f32,f32,p]
lib.emu_grad_norm.argtypes=[p,p,i32,p,f32,p]
lib.emu_grad_scale.argtypes=[p,p,i32,p]
BF=torch.bfloat16
random.seed(3); t_end=time.time()+150; n=0
def sl(n_, dtype, scale=1.0, positive=False):
    off=random.choice([0,0,1,3,8]); base=(torch.rand(n_+off) if positive else torch.randn(n_+off))*scale
    return base.to(dtype)[off:]
while time.time()<t_end:
    torch.manual_seed(random.randrange(1<<30))
    nt=random.randint(1,3); sizes=[random.choice([1,5,8,63,64,1000,32768,32769,40000]) for _ in range(nt)]
    fp32=random.random()<0.5; master=random.random()<0.4; sdt=torch.float32 if fp32 else BF
    ps=[sl(s,BF,0.5) for s in sizes]; gs=[sl(s,BF,2.0) for s in sizes]; ms=[sl(s,sdt,0.1) for s in sizes]; vs=[sl(s,sdt,0.1,True) for s in sizes]
    mw=[ (p_.float()+torch.randn(p_.numel())*1e-4) for p_ in ps] if master else [None]*nt
    step=random.randint(1,50); lr=10**random.uniform(-5,-1); wd=random.choice([0.0,0.1]); gsc=random.choice([None,0.3])
    want=[O.adamw_step(mw[i] if master else ps[i], gs[i], ms[i], vs[i], step, lr, 0.9, 0.95, 1e-8, wd, grad_scale=gsc or 1.0) for i in range(nt)]
    rows=[]; cm=[]
    for i in range(nt):
        rows.append((ps[i].data_ptr(),gs[i].data_ptr(),ms[i].data_ptr(),vs[i].data_ptr(),sizes[i],mw[i].data_ptr() if master else 0))
        cm+= [(i,c) for c in range((sizes[i]+32767)//32768)]
    random.shuffle(cm)  # any block order
    table=torch.tensor(rows,dtype=torch.int64); cmap=torch.tensor(cm,dtype=torch.int32)
    # norm first (does not modify)
    part=torch.empty(len(cm)); out2=torch.empty(2)
    lib.emu_grad_norm(table.data_ptr(),cmap.data_ptr(),len(cm),part.data_ptr(),1.0,out2.data_ptr())
    tot,coef=O.grad_norm_and_coef(gs,1.0)
    assert abs(out2[0].item()-tot)<=1e-4*tot+1e-6 and abs(out2[1].item()-coef)<1e-5, (out2, tot, coef)
    gsc_t=torch.tensor([gsc],dtype=torch.float32) if gsc else None
    lib.emu_adamw(table.data_ptr(),cmap.data_ptr(),len(cm),int(fp32)|(2 if master else 0),lr,0.9,0.95,1e-8,wd,1-0.9**step,math.sqrt(1-0.95**step),gsc_t.data_ptr() if gsc else None)
    for i,(pw,mw_,vw) in enumerate(want):
        if master:
            torch.testing.assert_close(mw[i], pw, atol=1e-6, rtol=1e-5); assert torch.equal(ps[i], mw[i].to(BF))
        else:
            torch.testing.assert_close(ps[i].float(), pw.float(), atol=1e-6, rtol=8e-3)
        tol=dict(atol=1e-7,rtol=1e-5) if fp32 else dict(atol=1e-6,rtol=8e-3)
        torch.testing.assert_close(ms[i].float(), mw_.float(), **tol); torch.testing.assert_close(vs[i].float(), vw.float(), **tol)
    n+=1
print("optimizer fuzz cases passed:", n)
