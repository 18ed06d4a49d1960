import numpy as np, sys
sys.path.insert(0,'.')
from luminair_b200.backend import CudaBackend, ColumnBatch
from oracle import cfft as ocfft
from oracle.circle import CanonicCoset
from oracle.fields import P
be=CudaBackend(0)
for log in [1,2,3,4,5,8,13]:
    rng=np.random.Generator(np.random.PCG64(log))
    vals=rng.integers(0,P,size=(2,1<<log),dtype=np.uint64).astype(np.uint32)
    dom=CanonicCoset(log).circle_domain()
    cb=ColumnBatch(be.upload(vals.reshape(-1)),2,log)
    be.interpolate(cb)
    got=be.download(cb.buf).reshape(2,-1)
    want=ocfft.interpolate(vals,dom).astype(np.uint32)
    print(log,"interp ok",np.array_equal(got,want))
    cb2=ColumnBatch(be.upload(want.reshape(-1)),2,log)
    out=ColumnBatch(be.alloc(2<<log),2,log)
    be.evaluate(cb2,out)
    g2=be.download(out.buf).reshape(2,-1)
    print(log,"eval ok",np.array_equal(g2,vals), (g2!=vals).sum())
    if log<=3: print(g2[0],vals[0])
    out2=ColumnBatch(be.alloc(2<<(log+1)),2,log+1)
    be.evaluate(cb2,out2)
    w2=ocfft.evaluate(want,CanonicCoset(log+1).circle_domain()).astype(np.uint32)
    g3=be.download(out2.buf).reshape(2,-1)
    print(log,"lde ok",np.array_equal(g3,w2),(g3!=w2).sum())
