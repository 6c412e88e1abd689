import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aide_b200.trainer import AideTrainer
dev = torch.device('cuda:0')
B, S = 8, 256
batches = [bench.make_batch(B, S, 1234 + i, device=dev) for i in range(2)]
for ts in (True, False):
    tr = AideTrainer('fuseunet', mode='parity', device=dev, seed=2, two_streams=ts)
    def step(i):
        b = batches[i % 2]; tr.step(b['x'], b['t1'], b['t2'], b['augs'], 0.25)
    for i in range(4): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): step(i)
    e1.record(); torch.cuda.synchronize()
    print('two_streams', ts, 'ms/step %.2f' % (e0.elapsed_time(e1) / 10), flush=True)
    del tr; torch.cuda.empty_cache()
