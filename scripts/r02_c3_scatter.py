"""Run-to-run scatter of log Z on BASELINE config 3 (Rastrigin 10-D, nlive 2000, R 50, clustering on) in the CPU oracle:
the reference schedule (one death per iteration, per-cluster evidences: add_cluster / delete_cluster) against the batched
schedule (global evidence), 8 seeds each.  python scripts/r02_c3_scatter.py  (about 5 minutes on 8 cores)"""
import sys, numpy as np, multiprocessing as mp
sys.path.insert(0,'tests')
def one(job):
    mode, seed = job
    import oracle_lib as O
    box=dict(prior_lo=[-5.12]*10, prior_hi=[5.12]*10)
    s=O.make_settings(10,0,nlive=2000,num_repeats=50,seed=seed,do_clustering=True,batch_K=(0 if mode=='ref' else 1000))
    r,_=O.run(s,like='rastrigin',**box)
    return mode,seed,r.logZ,r.logZerr,r.ncluster
if __name__=='__main__':
    jobs=[('ref',s) for s in range(8)]+[('bat',s) for s in range(8)]
    with mp.Pool(8) as p: res=p.map(one,jobs)
    for m in ('ref','bat'):
        z=np.array([r[2] for r in res if r[0]==m]); e=np.array([r[3] for r in res if r[0]==m])
        print(m,'mean',z.mean(),'std',z.std(ddof=1),'reported err',e.mean(),'ncl',[r[4] for r in res if r[0]==m])
