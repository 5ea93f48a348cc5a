import numpy as np, sys, time
sys.path.insert(0, "/root/repo")
import pymf_b200
from oracle import nmf_oracle as O
def rel(a,b): return float(np.linalg.norm(a-b)/np.linalg.norm(b))
for (d,n,k) in [(256,512,32),(300,1000,40),(512,640,128),(4096,2048,32)]:
    rng=np.random.RandomState(1)
    X=rng.random_sample((d,n)).astype(np.float32); W0=rng.random_sample((d,k)); H0=rng.random_sample((k,n))
    Wr,Hr=W0.copy(),H0.copy(); O.update_h(X.astype(np.float64),Wr,Hr)
    out={}
    for path in ("simt","tc"):
        e=pymf_b200.Engine(d,n,k,path=path); e.set_err_mode("trace"); e.upload_x(X); e.set_w(W0); e.set_h(H0)
        e.run(1,compute_w=False,compute_h=True,compute_err=False,early_stop=False)
        H=e.get_h(); fn=e.frobenius(); out[path]=(H,fn)
        print(d,n,k,path,"relH vs f64 after update_h: %.3e"%rel(H,Hr),"ferr",fn, "oracle", O.frobenius_norm(X.astype(np.float64),Wr,Hr), flush=True)
        e.close()
