import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import delta_graph_slam_b200 as d
from oracle import oracle_py as O
P0,P1=O.synth_traj(0),O.synth_traj(1)
s0=O.synth_scan(P0,noise_seed=1000); s1=O.synth_scan(P1,noise_seed=1001)
v0=O.voxelgrid(s0,0.1)['out']; v1=O.voxelgrid(s1,0.1)['out']
ref=O.Registration(O.NDT,resolution=1.0); ref.setInputTarget(v0); L=ref.ndt_leaves()
ndt=d.NormalDistributionsTransform(); ndt.setResolution(1.0); ndt.setInputTarget(v0); G=ndt.ndt_leaves()
print(len(L['idx']),len(G['idx']), np.array_equal(L['idx'],G['idx']))
bad=np.nonzero(L['n']!=G['n'])[0]
print("n mismatch", len(bad), L['n'][bad][:10], G['n'][bad][:10])
for b in bad[:3]:
    print("cov L",L['cov'][b]); print("cov G",G['cov'][b]); print(np.linalg.eigvalsh(L['cov'][b]))
valid=(L['n']>=6)&(G['n']>=6)
print("mean maxdiff",np.abs(G['mean']-L['mean']).max())
sc=np.abs(L['cov'][valid]).max(axis=(1,2),keepdims=True)
print("cov rel",(np.abs(G['cov'][valid]-L['cov'][valid])/sc).max())
sc=np.abs(L['icov'][valid]).max(axis=(1,2),keepdims=True)
r=(np.abs(G['icov'][valid]-L['icov'][valid])/sc).max(axis=(1,2)); print("icov rel",r.max(), np.argmax(r))
ref.setInputSource(v1); ndt.setInputSource(v1)
for p in ([0,0,0,0,0,0],[0.4,-0.1,0.03,0.01,-0.02,0.05]):
    p=np.array(p,float)
    s0_,g0,H0=ref.ndt_derivatives(p); s1_,g1,H1=ndt.ndt_derivatives(p)
    print("score",s0_,s1_,s1_-s0_); print("g",g0,g1); print("H rel",np.abs(H1-H0).max()/np.abs(H0).max())
    print(np.abs(H1-H0)/np.abs(H0).max())
