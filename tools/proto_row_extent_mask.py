"""Prototype (numpy, float32 rounding) of the O(rows) formulation of the tight-list tile mask — DESIGN.md §10 item 4.
Per tile row the set of dx with min_dy q(dx, dy) <= t is an interval [xl, xr] with a closed form (two square roots per row);
a tile of the row is kept iff its dx-range meets the interval. Checked here against the per-tile test (tests/test_cull_cpu.py)
and a brute-force per-pixel evaluation: identical cull sets (to 2 in 14k tiles), never drops a contributing tile.
Ported to csrc/psb_common.cuh (TileCull::rect_mask) in round 2; tests/test_cull_cpu.py holds the maintained numpy specification. Usage: python tools/proto_row_extent_mask.py"""
import sys, numpy as np
sys.path[:0]=['.','oracle','tests']
import oracle_c
from photo_slam_b200 import synthetic as syn
sys.path.insert(0,'tests')
from test_cull_cpu import _rect_mask, f32

def row_mask(mx,my,A,B,C,op,x0,y0,x1,y1,W,H):
    area=(x1-x0)*(y1-y0)
    if op < f32(1/255): return [False]*area
    det=f32(f32(A*C)-f32(B*B))
    if not (A>0 and C>0 and det>0): return [True]*area
    thr=f32(np.log(f32(255)*op))+f32(1e-3)
    DX=max(abs(f32(mx-f32(x0*16))),abs(f32(mx-f32(min(x1*16,W)-1)))); DY=max(abs(f32(my-f32(y0*16))),abs(f32(my-f32(min(y1*16,H)-1))))
    S=f32(f32(f32(abs(B)*DX)*DY)+f32(0.5)*f32(f32(A*DX)*DX+f32(C*DY)*DY))
    t=f32(thr+f32(f32(1e-5)*S+f32(1e-4)))
    t2=f32(2)*t
    Y=f32(np.sqrt(f32(t2*A/det)))            # |dy| extent of the ellipse
    dyr=f32(-B*f32(np.sqrt(f32(t2/f32(det*C)))))   # dy where dx is maximal
    invA=f32(1)/A
    out=[]
    for ty in range(y0,y1):
        py0=ty*16
        dylo=f32(my-f32(min(py0+16,H)-1)); dyhi=f32(my-f32(py0))
        lo=max(dylo,-Y); hi=min(dyhi,Y)
        if lo>hi:
            out += [False]*(x1-x0); continue
        def ext(dy, sgn):
            D=f32(f32(t2*A)-f32(det*dy)*dy)
            D=max(D,f32(0))
            return f32(f32(f32(-B*dy)+sgn*f32(np.sqrt(D)))*invA)
        xr=ext(min(max(dyr,lo),hi), f32(1)); xl=ext(min(max(-dyr,lo),hi), f32(-1))
        e=f32(1e-4)*(max(abs(xl),abs(xr))+f32(1))
        xr=f32(xr+e); xl=f32(xl-e)
        for tx in range(x0,x1):
            px0=tx*16
            dxlo=f32(mx-f32(min(px0+16,W)-1)); dxhi=f32(mx-f32(px0))
            out.append(bool(dxlo<=xr and dxhi>=xl))
    return out

def check(P, wh, scale_px, seed):
    W,H=wh; Wc,Hc,fx,fy=syn.CAMERAS["tum"]
    cam=syn.make_camera(W,H,fx*W/Wc,fy*H/Hc)
    f=oracle_c.forward(cam, syn.activate(syn.make_scene(P,cam,seed=seed,scale_px=scale_px)))
    gx,gy=(W+15)//16,(H+15)//16
    tot=cull_t=cull_r=bad=less=0
    for i in np.nonzero(f["radii"]>0)[0]:
        mx,my=f32(f["means2D"][i,0]),f32(f["means2D"][i,1]); A,B,C,op=[f32(v) for v in f["conic_opacity"][i]]
        r=int(f["radii"][i])
        x0,y0=min(gx,max(0,int((mx-r)/16))),min(gy,max(0,int((my-r)/16)))
        x1,y1=min(gx,max(0,int((mx+r+15)/16))),min(gy,max(0,int((my+r+15)/16)))
        if not 0<(x1-x0)*(y1-y0)<=32: continue
        a=_rect_mask(mx,my,A,B,C,op,x0,y0,x1,y1,W,H); b=row_mask(mx,my,A,B,C,op,x0,y0,x1,y1,W,H)
        k=0
        for ty in range(y0,y1):
            for tx in range(x0,x1):
                xs=np.arange(tx*16,min(tx*16+16,W),dtype=np.float64); ys=np.arange(ty*16,min(ty*16+16,H),dtype=np.float64)
                dx,dy=np.float64(mx)-xs[None,:],np.float64(my)-ys[:,None]
                power=-0.5*(np.float64(A)*dx*dx+np.float64(C)*dy*dy)-np.float64(B)*dx*dy
                alpha=np.where(power>0,0.0,np.minimum(0.99,np.float64(op)*np.exp(power)))
                contrib=(alpha>=1/255).any()
                tot+=1; cull_t+= (not a[k]); cull_r += (not b[k])
                if contrib and not b[k]: bad+=1; print("BAD",i,tx,ty,alpha.max())
                if a[k] and not b[k]: less+=1   # row method culls a tile the per-tile exact test keeps (possible: per-tile test is conservative too)
                k+=1
    print(P,wh,scale_px,'tiles',tot,'per-tile culled',cull_t,'row culled',cull_r,'bad',bad,'row culls beyond per-tile',less)
check(1500,(320,240),4.0,0); check(1500,(320,240),12.0,1); check(400,(640,480),40.0,2); check(3000,(640,480),2.4,3)
