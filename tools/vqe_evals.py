import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from tyxonq_b200 import ucc
from tyxonq_b200.vqe import TFIMVqe
dev=torch.device('cuda',0)
i1,i2=ucc.random_integral(7,2077); ex,pids=ucc.uccsd_ex_ops(5,2)
sv=ucc.UCCStatevector(14,(5,5),ex,pids,ucc.hamiltonian_from_integral(i1,i2),device=dev)
np.random.seed(2077); p=np.random.rand(75)-0.5
for graph in (False, True):
    for _ in range(3): sv.energy_and_grad(p, graph=graph)
    t=time.perf_counter(); R=50
    for _ in range(R): e,g=sv.energy_and_grad(p, graph=graph)
    print('ucc graph',graph, R/(time.perf_counter()-t),'evals/s', e)
v=TFIMVqe(10,1,device=dev); q=np.random.default_rng(0).normal(size=(2,10))
for graph in (False, True):
    for _ in range(3): v.energy_and_grad(q, graph=graph)
    t=time.perf_counter(); R=100
    for _ in range(R): e,g=v.energy_and_grad(q, graph=graph)
    print('tfim graph',graph, R/(time.perf_counter()-t),'evals/s', e)
