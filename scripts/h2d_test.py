import torch, time, numpy as np, sys
sys.path.insert(0, '.')
import parsmurf_b200 as psm
x = np.random.default_rng(0).standard_normal((1_000_000, 27))
xp = torch.from_numpy(x).pin_memory()
torch.cuda.synchronize()
for i in range(3):
    t0 = time.time(); xd = xp.to('cuda'); torch.cuda.synchronize(); t1 = time.time()
    print('torch pinned H2D 216MB: %.1f ms' % ((t1 - t0) * 1e3))
for i in range(2):
    t0 = time.time(); xd = torch.from_numpy(x).to('cuda'); torch.cuda.synchronize(); t1 = time.time()
    print('torch pageable H2D: %.1f ms' % ((t1 - t0) * 1e3))
c = psm.Context(0)
for i in range(3):
    t0 = time.time(); c.dataset_upload(xp.numpy()); t1 = time.time()
    print('psm upload (pinned): %.1f ms' % ((t1 - t0) * 1e3))
for i in range(2):
    t0 = time.time(); c.dataset_upload(x); t1 = time.time()
    print('psm upload (pageable): %.1f ms' % ((t1 - t0) * 1e3))
t0 = time.time(); c2 = psm.Context(0); t1 = time.time(); print('ctx create %.1f ms' % ((t1-t0)*1e3))
