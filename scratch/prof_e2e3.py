import time, numpy as np, torch, sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hail_b200 as hb
from hail_b200 import _lib, statgen
from hail_b200.statgen import GroupBasis, _HostStream
N = 400000; Me = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rng = np.random.default_rng(0)
cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(9)]); y = rng.standard_normal((N, 1))
h = torch.empty((Me, (N + 3) // 4), dtype=torch.uint8, pin_memory=True)
h.random_(0, 255)
g = hb.HostBedGenotypes(h, N, 0)
for r in range(8):
    torch.cuda.synchronize()
    t0 = time.time(); st = _HostStream(g)
    t1 = time.time(); b = GroupBasis(y, cov, np.arange(N))
    t2 = time.time(); outs = st.run([b])
    t3 = time.time(); st.close()
    t4 = time.time()
    print("begin %.3f  prologue %.3f  run %.3f  close %.3f  total %.3f  -> %.3e g/s" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0, Me * N / (t4 - t0)))
# the same without the prologue in between (pure pipeline)
b = GroupBasis(y, cov, np.arange(N))
for r in range(3):
    torch.cuda.synchronize()
    t0 = time.time(); st = _HostStream(g); outs = st.run([b]); st.close(); t4 = time.time()
    print("no-prologue total %.3f -> %.1f GB/s" % (t4 - t0, h.numel() / 1e9 / (t4 - t0)))
