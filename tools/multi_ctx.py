"""Experiment: C contexts x S sequences on one GPU, driven by C host threads on C streams."""
import sys, os, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from __graft_entry__ import load_package
pkg = load_package()
C = int(sys.argv[1]); S = int(sys.argv[2]); K = 12; W = 3
n_frames = bench.PREFILL + W + K
worlds = bench.make_worlds(n_frames, 0, 8)
ctxs = []
for c in range(C):
    odo = bench.slot_odometry(worlds, S, c)
    steps = [bench.pack_step(worlds, S, f) for f in range(n_frames)]
    dev = [(torch.from_numpy(a).cuda(), co, torch.from_numpy(b).cuda(), so) for a, co, b, so in steps]
    q_all = np.array([[odo[s][f, :4] for s in range(S)] for f in range(n_frames)])
    t_all = np.array([[odo[s][f, 4:] for s in range(S)] for f in range(n_frames)])
    st = torch.cuda.Stream()
    mc = max(int(np.diff(x[1]).max()) for x in steps); ms = max(int(np.diff(x[3]).max()) for x in steps)
    R = pkg.Registrar(0.4, 0.8, batch=S, cap_corner_in=mc + 64, cap_surf_in=ms + 64, cap_map_corner=1 << 17, cap_map_surf=1 << 17)
    R.set_stream(st.cuda_stream)
    ctxs.append((R, dev, q_all, t_all, st))
def run(c, f0, f1, out):
    R, dev, q_all, t_all, st = ctxs[c]
    with torch.cuda.stream(st):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for f in range(f0, f1):
            a, co, b, so = dev[f]
            R.register_batch(a.data_ptr(), co, b.data_ptr(), so, q_all[f], t_all[f], device_ptrs=True)
        e1.record(st)
    out[c] = (e0, e1)
out = {}
th = [threading.Thread(target=run, args=(c, 0, bench.PREFILL + W, out)) for c in range(C)]
[t.start() for t in th]; [t.join() for t in th]
torch.cuda.synchronize()
out = {}
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(c, bench.PREFILL + W, n_frames, out)) for c in range(C)]
[t.start() for t in th]; [t.join() for t in th]
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = max(a.elapsed_time(b) for a, b in out.values())
print(f"contexts {C} x seqs {S}: {C*S*K/(ms*1e-3):.0f} reg/s (events max {ms/K:.3f} ms/step, wall {1e3*wall/K:.3f} ms/step)")
