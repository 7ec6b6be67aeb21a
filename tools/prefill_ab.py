import sys, torch
sys.path.insert(0, ".")
from eccq_b200.kv_cache import SimpleBlockManager, paged_prefill_attention_ecc
from eccq_b200 import _lib
lib=_lib.load(); P=_lib.ptr; dev="cuda"
for (Hh,Dh,Sq,Hkv,dt) in ((12,64,1024,12,torch.float32),(32,128,2048,8,torch.float32),(32,128,1024,8,torch.float16)):
    nblk=Sq//16
    m=SimpleBlockManager(nblk,16,1,Hkv,Dh,device=dev,codec="hamming84")
    m.block_table=torch.randperm(nblk).to(torch.int32).view(1,Sq//16).to(dev)
    kk=torch.randn(1,Sq,Hkv*Dh,device=dev,dtype=torch.float16)
    sp=_lib.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.eccq_cache_write(m.desc(),P(kk),P(kk.flip(1)),_lib.ECCQ_F16,P(m.block_table),Sq//16,None,1,Sq,0,_lib.INJECT_PHILOX_FAST,None,None,1e-3,1,0,None,sp)
    q=torch.randn(1,Hh,Sq,Dh,device=dev,dtype=dt)
    cl=torch.full((1,),Sq,dtype=torch.int32,device=dev)
    fn=lambda: paged_prefill_attention_ecc(q,m.k_cache,m.v_cache,m.block_table,cl,m.k_scales,m.v_scales,0,16,use_interpolation=True)
    for _ in range(3): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize(); print(Hh,Dh,Sq,dt,round(a.elapsed_time(b)/20,4),"ms")
