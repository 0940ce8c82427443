import sys, torch
sys.path.insert(0,'/root/repo')
from l2hmc_b200 import ops, _lib
L, nb, nlf = 16, 64, 10
torch.manual_seed(0)
shape=[L]*4
x = ops.su3_project(torch.complex(torch.randn(nb,4,*shape,3,3,dtype=torch.float64,device='cuda'), torch.randn(nb,4,*shape,3,3,dtype=torch.float64,device='cuda')))
v = ops.su3_rand_momentum(nb, shape, 1, 0, 'cuda')
for carve in (-1, 0, 25, 50, 100):
    _lib.set_option('su3_force_carveout', carve)
    for _ in range(2): ops.su3_hmc_trajectory(x, v, 6.0, 0.1, nlf)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): ops.su3_hmc_trajectory(x, v, 6.0, 0.1, nlf)
    e1.record(); torch.cuda.synchronize()
    print('carveout', carve, 'traj ms', e0.elapsed_time(e1)/4)
