import sys, torch
sys.path.insert(0, '.')
from maed_b200 import ops
torch.manual_seed(0)
def run(n,H,Cin,C,k,res,label):
    x=torch.randn(n,H,H,Cin,device='cuda'); w=torch.randn(C,k*k*Cin,device='cuda')*0.1
    g=torch.ones(C,device='cuda'); b=torch.zeros(C,device='cuda')
    a=ops.split(x); wp=ops.split(w); rp=ops.split(torch.randn(n,H,H,C,device='cuda')) if res else None
    dbg=torch.zeros(4096*8,dtype=torch.int64,device='cuda')
    for _ in range(2): ops.conv_gn(a,wp,k,k,g,b,True,rp,dbg=dbg)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv_gn(a,wp,k,k,g,b,True,rp,dbg=dbg); e1.record(); torch.cuda.synchronize()
    d=dbg.cpu().reshape(-1,8)
    d=d[d[:,0]>0]
    t0=d[0,0].item()
    print(label,'kernel %.1f us, items for cta0/grp0: %d'%(e0.elapsed_time(e1)*1e3,len(d)))
    for i in range(min(len(d),6)):
        r=d[i]
        print('  item %d: start %7d | tile_full +%6d | pass1 end +%6d | bar +%5d | exch+mr +%6d | coef +%5d | pass2 +%6d'%(
            i,r[0]-t0,r[6]-r[0],r[1]-r[6],r[2]-r[1],r[3]-r[2],r[4]-r[3],r[5]-r[4]))
    if len(d)>1: print('  avg period per item (cycles):', ((d[-1,0]-d[0,0]).item())/(len(d)-1))
run(128,56,64,256,1,True,'s0 conv3 (64->256, res)')
run(128,56,64,64,3,False,'s0 conv2 3x3')
run(128,28,128,512,1,True,'s1 conv3')
run(128,14,256,1024,1,True,'s2 conv3')
run(128,14,256,256,3,False,'s2 conv2 3x3')
