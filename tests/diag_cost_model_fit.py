#!/usr/bin/env python
"""Least-squares fit of  time(layer) = F + a * tiles_per_CTA + b * active_chunks_per_CTA  to the per-layer CUDA-event times of
profiles/r1_final_step_breakdown_diag.json (batch 16, scenes 0..15), with tile / active-chunk counts recomputed on the CPU
from the same synthetic scenes.  F = per-launch fixed cost (launch gap, prologue, first index tile, tail imbalance),
a = per-tile fixed cost, b = per-chunk cost (MMA floor: 192 cycles at N = 32, 384 at N = 64).  Analysis helper for the
profile write-up: it runs the CPU rulebook restatement under oracle/ as a *counter* of tiles and chunks only."""
import numpy as np, sys, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from btcdet_b200 import synthetic as S
from oracle import oracle as O
B=16
scenes=[S.lidar_like(20000, seed=i) for i in range(B)]
vox=O.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
coords=vox[1]; shape=[41,1600,1408]
def chunks(nbr, cin):
    n,K=nbr.shape; T=(K*cin+31)//32; nt=(n+127)//128; tot=0
    for t in range(nt):
        m=(nbr[t*128:(t+1)*128]>=0).any(0)
        for c in range(T):
            klo=(32*c)//cin; khi=min((32*c+31)//cin,K-1)
            tot+=bool(m[klo:khi+1].any())
    return nt, tot
layers=[]
o,p,pn,_=O.get_indice_pairs(coords,B,shape,3,subm=True); t1=O.pairs_to_tables(p,pn,coords.shape[0],coords.shape[0])[0]
layers+= [("4->16",32)+chunks(t1,4), ("16->16",32)+chunks(t1,16)]
lv,sh=coords,shape
for (ks,st,pd,cin,cout) in [((3,3,3),(2,2,2),(1,1,1),16,32),((3,3,3),(2,2,2),(1,1,1),32,64),((3,3,3),(2,2,2),(0,1,1),64,64)]:
    oi,p,pn,osh=O.get_indice_pairs(lv,B,sh,list(ks),list(st),list(pd))
    t=O.pairs_to_tables(p,pn,lv.shape[0],oi.shape[0])[0]
    N=32 if cout<=32 else 64
    layers.append(("s %d->%d"%(cin,cout),N)+chunks(t,cin))
    lv,sh=oi,list(osh)
    o2,p2,pn2,_=O.get_indice_pairs(lv,B,sh,3,subm=True); t2=O.pairs_to_tables(p2,pn2,lv.shape[0],lv.shape[0])[0]
    c=chunks(t2,cout)
    layers.append(("subm %d->%d a"%(cout,cout),N)+c); layers.append(("subm %d->%d b"%(cout,cout),N)+c)
oi,p,pn,osh=O.get_indice_pairs(lv,B,sh,[3,1,1],[2,1,1],[0,0,0]); t=O.pairs_to_tables(p,pn,lv.shape[0],oi.shape[0])[0]
layers.append(("64->128",128)+chunks(t,64))
d=json.loads(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r1_final_step_breakdown_diag.json')).read())
us=[r["us"] for r in d["steps"] if r["step"].startswith("conv ")]
print(len(layers), len(us))
rows=[]
for (name,N,nt,tot),u in zip(layers,us):
    print("%-16s N=%3d tiles %5d (%.2f/CTA) chunks/tile %.1f  %.1f us"%(name,N,nt,nt/148,tot/nt,u)); rows.append((N,nt/148,tot/148,u))
for Nc in (32,64):
    A=np.array([[1,r[1],r[2]] for r in rows if r[0]==Nc]); y=np.array([r[3] for r in rows if r[0]==Nc])
    x,res,_,_=np.linalg.lstsq(A,y,rcond=None)
    print("N=%d: F=%.1f us/launch, a=%.2f us/tile, b=%.3f us/chunk; fit residuals"%(Nc,*x), np.round(A@x-y,1))
