"""Accuracy probe (checker, hence under tests/): float32 CUDA library against the NumPy float64 oracle after 1-3 outer
iterations at a few medium shapes; prints relative Frobenius errors.  Run on the GPU box: python tests/probes/errprobe.py"""
import os, sys, numpy as np, scipy.sparse as sps
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/exp-trmf-nips16_b200')
import cases
from oracle import abi, trmf_numpy as tn
CORELIB='/root/repo/exp-trmf-nips16_b200/trmf/corelib'
def run(Y,lags,W,H,L,dtype,**kw):
    return abi.run_train(os.path.join(CORELIB,'trmf_float32.so' if dtype==np.float32 else 'trmf_float64.so'),Y,lags,W,H,L,dtype=dtype,**kw)
for (T,n,k,lags,dens,iters) in [(700,450,40,[1,7,24],0.6,1),(2000,1500,40,[1,7,24],0.9,1),(2000,1500,40,[1,7,24],0.9,3),(1500,1200,64,[1,2,3,24],0.5,2),(3000,400,20,list(range(1,25)),0.9,2)]:
    p = cases.make_problem(T,n,k,lags,dens,seed=T+k)
    f32=lambda a: np.asarray(a,dtype=np.float32)
    Y = sps.csr_matrix((f32(p['Ysp'].data), p['Ysp'].indices, p['Ysp'].indptr), shape=p['Ysp'].shape)
    W0,H0,L0 = f32(p['W0']),f32(p['H0']),f32(p['L0'])
    kw = dict(lambdaI=0.5,lambdaAR=50.0,lambdaLag=0.5,max_iter=iters,period_Lag=1,missing=True)
    Wo,Ho,Lo = tn.train(Y.astype(np.float64),p['lags'],W0.astype(np.float64),H0.astype(np.float64),L0.astype(np.float64),**kw)
    Wr,Hr,Lr = abi.run_reference(Y,p['lags'],W0,H0,L0,dtype=np.float32,threads=8,**kw) if abi.ref_available(np.float32) else (Wo,Ho,Lo)
    for env in ({}, {'TRMF_B200_NO_GRAM_HV':'1'}, {'TRMF_B200_GENERIC_F':'1','TRMF_B200_NO_GRAM_HV':'1'}):
        for k_,v in env.items(): os.environ[k_]=v
        W,H,L = run(Y,p['lags'],W0,H0,L0,np.float32,**kw)
        for k_ in env: del os.environ[k_]
        print(f"T{T} n{n} k{k} it{iters} {str(sorted(env)):60s} W {cases.rel(W,Wo):.2e} H {cases.rel(H,Ho):.2e} L {cases.rel(L,Lo):.2e}")
    print(f"   reference float32 build vs float64 oracle:                         W {cases.rel(Wr,Wo):.2e} H {cases.rel(Hr,Ho):.2e} L {cases.rel(Lr,Lo):.2e}")
