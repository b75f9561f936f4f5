"""Tile-visit statistics of an optimal path: python tile_stats.py <m> <n> (256 x 256 tiles)."""
import numpy as np, sys
rng=np.random.default_rng(1)
def nw_path(x,y,pxy=3,pgap=2):
    m,n=len(x),len(y)
    H=np.zeros((m+1,n+1),np.int32)
    H[:,0]=np.arange(m+1)*pgap; H[0,:]=np.arange(n+1)*pgap
    # row-wise DP using the min-plus scan trick: H[i,j]=min(diag, up, left+pgap)
    for i in range(1,m+1):
        cost=np.where(y==x[i-1],0,pxy).astype(np.int32)
        t=np.minimum(H[i-1,:-1]+cost, H[i-1,1:]+pgap)   # without left
        # left dependency: H[i,j]=min(t[j], H[i,j-1]+pgap) -> prefix min of (t[j]-j*pgap)
        idx=np.arange(1,n+1)*pgap
        a=np.concatenate(([H[i,0]], t-idx))
        H[i,1:]=np.minimum.accumulate(a)[1:]+idx
    i,j=m,n; path=[]
    while i>0 and j>0:
        path.append((i,j))
        c=0 if x[i-1]==y[j-1] else pxy
        if H[i,j]==H[i-1,j-1]+c: i-=1;j-=1
        elif H[i,j]==H[i-1,j]+pgap: i-=1
        else: j-=1
    return path
m=int(sys.argv[1]); n=int(sys.argv[2])
x=rng.integers(0,4,m); y=rng.integers(0,4,n)
path=nw_path(x,y)
T=256
tiles=[]; cur=None
for (i,j) in path:
    key=((i-1)//T,(j-1)//T)
    if key!=cur:
        cur=key; tiles.append([key,(i-1)%T,(j-1)%T,0])
    tiles[-1][3]+=1
steps=0; cells=0; walk=0; full=0
for key,li,lj,w in tiles:
    tq=lj//4+1; tmax=li//8
    steps+=tq+tmax; cells+=(li+1)*(lj+1); walk+=w
print("tiles",len(tiles),"per 256 diag",len(tiles)/((m+n)/2/T),"steps",steps,"avg steps/tile",steps/len(tiles),
      "useful lane-steps frac",cells/32/ (steps*32*32/32)/1.0 ,"walk",walk)
import collections
print(sorted(collections.Counter((min(li//64,3),min(lj//64,3)) for _,li,lj,_ in tiles).items()))
print("sum steps small tiles (cells<8192):",sum((lj//4+1+li//8) for _,li,lj,_ in tiles if (li+1)*(lj+1)<8192), "count", sum(1 for _,li,lj,_ in tiles if (li+1)*(lj+1)<8192))
