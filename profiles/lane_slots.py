# Where the lane slots of the neighbour sweeps go: the headline scene evaluated on the CPU.
# python profiles/lane_slots.py [small]  (needs only the scene generator of libpbf_b200, no GPU)
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import pbf_b200
n3=(256,128,256); grid=(512,256,512)
if len(sys.argv)>1 and sys.argv[1]=='small': n3=(128,64,128); grid=(256,128,256)
pos,vel=pbf_b200.dam_break(*n3)
p=pos[:,:3]
c=np.floor(p).astype(np.int64)
gx,gy,gz=grid
key=c[:,0]+c[:,2]*gx+c[:,1]*gx*gz
order=np.argsort(key,kind='stable')
ks=key[order]
N=len(ks)
cnt=np.bincount(ks,minlength=gx*gy*gz).astype(np.int32)
cnt3=cnt.reshape(gy,gz,gx)
# merged run counts for cell (x): cells x-1..x+1 in row (y+dy,z+dz)
pad=np.pad(cnt3,((1,1),(1,1),(1,1)))
run3=pad[:,:,:-2]+pad[:,:,1:-1]+pad[:,:,2:]   # shape (gy+2,gz+2,gx): sum over x-1..x+1 at padded rows
cs=c[order]
runs=np.zeros((N,9),np.int32)
o=0
for dy in (-1,0,1):
    for dz in (-1,0,1):
        runs[:,o]=run3[cs[:,1]+1+dy, cs[:,2]+1+dz, cs[:,0]]
        o+=1
tot=runs.sum(1)
print("particles",N,"candidates/particle (incl self)",tot.mean())
W=32
M=(N//W)*W
r=runs[:M].reshape(-1,W,9)
# current: sorted by iterations desc per lane, per slot warp-max of ceil(c/2)
it=(r+1)//2
its=-np.sort(-it,axis=2)
cur=its.max(1).sum(1)
print("current pair-iterations per warp (sorted runs):",cur.mean(), " row order:", it.max(1).sum(1).mean())
# flattened: warp max of ceil(total/2)
t=r.sum(2)
flat=((t+1)//2).max(1)
print("flattened (one run) per warp:",flat.mean(), " ideal mean ceil(total/2):",((t+1)//2).mean())
# three runs (one per dy: 3 z-rows merged... not contiguous) ; three column-runs variant: per x-column contiguous? same total
# 3 runs of ~11: per-slot max after sorting
r3=r.reshape(-1,W,3,3).sum(3)
it3=-np.sort(-((r3+1)//2),axis=2)
print("three runs per particle:",it3.max(1).sum(1).mean())
