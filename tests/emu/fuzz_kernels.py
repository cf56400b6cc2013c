"""Randomised runs of the host-emulated kernels (see cuda_emu.h): GEMV with padded strides, decode attention over random
head_dim / GQA group / context / splits / window / softcap / padding ranges, pull-reduce for world 1-8.
Not collected by pytest (minutes of runtime): python tests/emu/fuzz_kernels.py [seed] [seconds].
End-of-round-1 run: 412 cases, 0 failures."""
import ctypes, math, os, random, subprocess, sys, time
import torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
ROOT="/root/repo"; EMU=ROOT+"/tests/emu"
out=os.environ.get("EMU_LIB", "/tmp/libemu_fuzz.so")
r=subprocess.run(["g++","-std=c++20","-O1","-fPIC","-shared","-Wno-attributes","-I",EMU,"-I",ROOT+"/transformers_b200/csrc","-I","/usr/local/cuda/include",EMU+"/emu_kernels.cpp","-o",out,"-lpthread"],capture_output=True,text=True)
assert r.returncode==0, r.stderr[-2000:]
lib=ctypes.CDLL(out)
p,i32,i64,f32=ctypes.c_void_p,ctypes.c_int,ctypes.c_int64,ctypes.c_float
lib.emu_gemv.argtypes=[p,p,p,i32,i32,i32,i32,i32,i32]
lib.emu_attn_decode.argtypes=[p,p,p,p,p,i32,p,i32,i32,i32,i32,i32]+[i64]*10+[f32,f32,i32,p,p,i32]
lib.emu_pull_reduce.argtypes=[p,i32,i64,i64,p,p,i32]
BF=torch.bfloat16
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+120
n=0
while time.time()<t_end:
    kind=random.choice(["gemv","decode","pull"])
    torch.manual_seed(random.randrange(1<<30))
    if kind=="gemv":
        M=random.randint(1,4); N=random.randint(1,40); K=8*random.randint(1,140)
        pad=8*random.randint(0,2)
        xs=torch.randn(M,K+pad).to(BF); ws=torch.randn(N,K+pad).to(BF)*0.2
        x,w=xs[:,:K],ws[:,:K]
        y=torch.full((M,N),float("nan"),dtype=BF)
        assert lib.emu_gemv(x.data_ptr(),w.data_ptr(),y.data_ptr(),M,N,K,K+pad,K+pad,N)==0
        torch.testing.assert_close(y.float(), x.float()@w.float().t(), atol=3e-2, rtol=1e-2)
    elif kind=="decode":
        D=random.choice([64,128,256]); G=random.choice([1,2,4,8]); Hkv=random.randint(1,2); Hq=Hkv*G; B=random.randint(1,2)
        ctx=random.randint(1,60); cap=ctx+random.randint(0,5); nsplit=random.randint(1,5)
        window=random.choice([0,0,random.randint(1,70)]); softcap=random.choice([0.0,0.0,20.0])
        kc=torch.randn(B,Hkv,cap,D).to(BF); vc=torch.randn(B,Hkv,cap,D).to(BF); q=torch.randn(B,1,Hq,D).to(BF)
        outt=torch.full((B,1,Hq,D),float("nan"),dtype=BF); lse=torch.full((B,Hq,128),float("nan")); ws=torch.full((B*Hq*nsplit*(D+2),),float("nan"))
        use_rng=random.random()<0.5
        ks=torch.tensor([random.randint(0,ctx-1) for _ in range(B)],dtype=torch.int32) if use_rng else None
        ke=torch.tensor([random.randint(int(ks[b])+1,ctx) for b in range(B)],dtype=torch.int32) if use_rng else None
        scale=D**-0.5
        rc=lib.emu_attn_decode(q.data_ptr(),kc.data_ptr(),vc.data_ptr(),outt.data_ptr(),lse.data_ptr(),128,ws.data_ptr(),B,ctx,Hq,Hkv,D,q.stride(0),q.stride(2),kc.stride(0),kc.stride(2),kc.stride(1),vc.stride(0),vc.stride(2),vc.stride(1),outt.stride(0),outt.stride(2),scale,softcap,window,ks.data_ptr() if use_rng else None,ke.data_ptr() if use_rng else None,nsplit)
        assert rc==0
        k=kc[:,:,:ctx].transpose(1,2).float().repeat_interleave(G,dim=2); v=vc[:,:,:ctx].transpose(1,2).float().repeat_interleave(G,dim=2)
        s=torch.einsum("bhd,bkhd->bhk",q[:,0].float(),k)*scale
        if softcap: s=softcap*torch.tanh(s/softcap)
        idx=torch.arange(ctx); valid=torch.ones(B,ctx,dtype=torch.bool)
        if window: valid&=idx[None]>=ctx-window
        if use_rng: valid&=(idx[None]>=ks[:,None])&(idx[None]<ke[:,None])
        s=s.masked_fill(~valid[:,None],float("-inf"))
        pr=torch.nan_to_num(torch.softmax(s,-1),nan=0.0)
        want=torch.einsum("bhk,bkhd->bhd",pr,v)
        torch.testing.assert_close(outt[:,0].float(),want,atol=2e-2,rtol=2e-2)
        l=torch.logsumexp(s,-1)
        fin=torch.isfinite(l)
        torch.testing.assert_close(lse[...,0][fin],l[fin],atol=2e-3,rtol=2e-3)
        assert torch.isinf(lse[...,0][~fin]).all()
    else:
        world=random.choice([1,2,4,8]); rows=random.randint(1,9); cols=8*random.randint(1,40)
        bufs=[torch.randn(world*rows,cols).to(BF) for _ in range(world)]
        res=torch.randn(rows,cols).to(BF) if random.random()<0.5 else None
        ptrs=(ctypes.c_void_p*world)(*[b.data_ptr() for b in bufs]); rank=random.randrange(world)
        o=torch.full((rows,cols),float("nan"),dtype=BF)
        lib.emu_pull_reduce(ctypes.cast(ptrs,p),world,rank*rows*cols,rows*cols,res.data_ptr() if res is not None else None,o.data_ptr(),random.randint(1,4))
        want=res.float() if res is not None else torch.zeros(rows,cols)
        for b in bufs: want=want+b[rank*rows:(rank+1)*rows].float()
        assert torch.equal(o,want.to(BF))
    n+=1
print("fuzz cases passed:",n)
