import sys, importlib, numpy as np
sys.path.insert(0,'/root/repo')
m=importlib.import_module("micro-quad-slam_b200"); syn=importlib.import_module("micro-quad-slam_b200.synth")
from oracle import orc
o=orc.Oracle()
w=syn.scaled(syn.CONFIGS["c3"],n_flights=1,n_samples=3000)
d=syn.generate(w,flight_id0=17); p=w.params()
x,y=o.pose_integrate(d["t_ms"][0],d["of_rate_x"][0],d["of_rate_y"][0],d["h_m"][0],d["yaw_deg"][0],d["of_q"][0])
cells,origin=o.beam_cells(p,x,y,d["yaw_deg"][0],d["ranges"][0])
# touched box
ok=cells[...,0]>=0
xs=np.concatenate([cells[...,0][ok],origin[:,0]]); ys=np.concatenate([cells[...,1][ok],origin[:,1]])
bx0=(xs.min()//4)*4; by0=ys.min(); bw=xs.max()+1-bx0; bh=ys.max()+1-by0
P=(bw+3)//4*4
if (P//4)%2==0: P+=4
print("box",bw,bh,"pitch",P)
frames=range(200,3000,37)
def wavefronts(addrs):
    # addrs: byte addresses of active lanes in one instruction
    if len(addrs)==0: return 0
    words=np.unique(addrs>>2); banks=words&31
    return np.bincount(banks,minlength=32).max()
def ray_params(f):
    out=[]
    for b in range(32):
        if cells[f,b,0]<0: out.append(None); continue
        dx=int(cells[f,b,0]-origin[f,0]); dy=int(cells[f,b,1]-origin[f,1])
        adx,ady=abs(dx),abs(dy); xm=adx>=ady; mm=adx if xm else ady; n=ady if xm else adx
        out.append((dx,dy,xm,mm,n))
    return out
def cell_addr(f,rp,k):
    dx,dy,xm,mm,n=rp
    q=(k*n+mm//2)//mm if mm else 0
    sx=1 if dx>=0 else -1; sy=1 if dy>=0 else -1
    cx=origin[f,0]+(k if xm else q)*sx; cy=origin[f,1]+(q if xm else k)*sy
    return int((cy-by0)*P+(cx-bx0))
K0=7
res={}
tot_updates=0
for name in ("beams32","fan8x4","beams16x2","along_ray","fan8x4_perfan"):
    res[name]=[0,0]  # instrs, wavefronts
for f in frames:
    rps=ray_params(f)
    ms=[rp[3] if rp else -1 for rp in rps]
    mmax=max(ms)
    tot_updates+=sum(max(mm-K0,0) for mm in ms)
    # A: lanes = 32 beams, one step k per instruction, k=K0..mmax-1
    for k in range(K0,mmax):
        a=np.array([cell_addr(f,rps[b],k) for b in range(32) if rps[b] and k<ms[b]],dtype=np.int64)
        res["beams32"][0]+=1; res["beams32"][1]+=wavefronts(a)
    # B: per fan of 8 beams x 4 consecutive steps; instruction = (fan, k4)
    for fan in range(4):
        fm=max(ms[fan*8:(fan+1)*8])
        for k in range(K0,fm,4):
            a=[cell_addr(f,rps[b],k+j) for b in range(fan*8,fan*8+8) for j in range(4) if rps[b] and k+j<ms[b]]
            res["fan8x4_perfan"][0]+=1; res["fan8x4_perfan"][1]+=wavefronts(np.array(a,dtype=np.int64))
    # B2: like B but the 4 fans processed as separate instructions only when... same as B (fan8x4 = B)
    res["fan8x4"]=res["fan8x4_perfan"]
    # C: 16 beams x 2 steps: halves (fans 0,1) and (fans 2,3)
    for half in range(2):
        hm=max(ms[half*16:(half+1)*16])
        for k in range(K0,hm,2):
            a=[cell_addr(f,rps[b],k+j) for b in range(half*16,half*16+16) for j in range(2) if rps[b] and k+j<ms[b]]
            res["beams16x2"][0]+=1; res["beams16x2"][1]+=wavefronts(np.array(a,dtype=np.int64))
    # D: lanes along ray: per beam chunks of 32 steps
    for b in range(32):
        if not rps[b]: continue
        for k in range(K0,ms[b],32):
            a=[cell_addr(f,rps[b],k+j) for j in range(32) if k+j<ms[b]]
            res["along_ray"][0]+=1; res["along_ray"][1]+=wavefronts(np.array(a,dtype=np.int64))
nf=len(list(frames))
print("frames",nf,"free updates/frame",tot_updates/nf)
for k,(i,wv) in res.items():
    print(f"{k:16s} instr/frame {i/nf:7.1f}  wavefronts/frame {wv/nf:7.1f}  wf/instr {wv/max(i,1):.2f}  lanes/instr {tot_updates/max(i,1):.1f}")
