"""Instruction model of the tile traceback for different tile shapes: optimal paths of random
ACGT pairs (numpy DP), tiles visited, wavefront steps and warp-instructions per pair."""
import numpy as np, sys
sys.argv=[sys.argv[0],'1','1']
exec(open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'tile_stats.py')).read().split("m=int(sys.argv[1])")[0])
rng=np.random.default_rng(5)
def model(path,TR,TC,perstep_over=50,percell=6.4,walk=32):
    RT=TR//32
    tiles=[];cur=None
    for (i,j) in path:
        key=((i-1)//TR,(j-1)//TC)
        if key!=cur:
            cur=key;tiles.append([(i-1)%TR,(j-1)%TC,0])
        tiles[-1][2]+=1
    ins=0;steps=0
    for li,lj,w in tiles:
        tq=lj//4+1;tmax=li//RT
        ins+=(tq+tmax)*(RT*4*percell+perstep_over)+w*walk
        steps+=tq+tmax
    return len(tiles),steps,ins
for (m,n) in [(6000,6000),(8000,3000)]:
    x=rng.integers(0,4,m);y=rng.integers(0,4,n)
    path=nw_path(x,y)
    for TR,TC in [(256,256),(128,256),(128,128),(64,256),(64,128),(128,512),(64,512)]:
        nt,st,ins=model(path,TR,TC)
        print(m,n,(TR,TC),"tiles",nt,"steps",st,"warp-instr",int(ins), "ckpt B/cell", 4/TR+4/TC)
