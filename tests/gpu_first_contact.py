"""Diagnostic (not a pytest): first contact of the CUDA path with the oracle, verbose."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness, oracle
from conftest import load_package, rot_angle
s2m = load_package()
truth, odom, frames = harness.sequence(20261018, "HDL64", 10)
R = s2m.Registrar(0.4, 0.8, trace=True)
O = oracle.Oracle(0.4, 0.8, trace=True)
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
for f in range(10):
    c, s = frames[f]
    t0 = time.time(); rg, qg, tg = R.register(c, s, odom[f, :4], odom[f, 4:]); tgms = (time.time() - t0) * 1e3
    t0 = time.time(); ro, qo, to = O.register(c, s, odom[f, :4], odom[f, 4:]); toms = (time.time() - t0) * 1e3
    sg, so = R.stats, O.stats
    print(f"frame {f}: rc {rg}/{ro} ds {sg.n_corner_ds},{sg.n_surf_ds}/{so.n_corner_ds},{so.n_surf_ds} "
          f"map {sg.n_map_corner},{sg.n_map_surf}/{so.n_map_corner},{so.n_map_surf} "
          f"edge {list(sg.n_edge)}/{list(so.n_edge)} plane {list(sg.n_plane)}/{list(so.n_plane)} "
          f"it {list(sg.lm_iters)}/{list(so.lm_iters)} term {list(sg.lm_term)}/{list(so.lm_term)} "
          f"cost {sg.cost_initial[0]:.6f}->{sg.cost_final[1]:.6f} / {so.cost_initial[0]:.6f}->{so.cost_final[1]:.6f} "
          f"dt {np.linalg.norm(tg - to):.3e} dr {rot_angle(qg, qo):.3e}  gpu {tgms:.2f} ms cpu {toms:.1f} ms", flush=True)
    for cls in (0, 1):
        a, b = R.trace_cloud(cls), O.trace_cloud(cls)
        if a.shape != b.shape or not np.array_equal(bits(a), bits(b)):
            print("   ds cloud mismatch cls", cls, a.shape, b.shape)
    if rg == 0:
        for outer in (0, 1):
            for cls in (0, 1):
                ig, dg, ug = R.trace_knn(outer, cls); io, do, uo = O.trace_knn(outer, cls)
                gate = do[:, 4] < 1.0
                if len(ig) != len(io):
                    print("   knn len mismatch", len(ig), len(io)); continue
                bad = int((ig[gate] != io[gate]).any(1).sum()); badd = int((bits(dg[gate]) != bits(do[gate])).any(1).sum())
                print(f"   outer {outer} cls {cls}: gate {int(gate.sum())}/{len(gate)} idx-mismatch {bad} d2-mismatch {badd} used-mismatch {int((ug != uo).sum())}")
            pg, sg_, itg, ng, termg = R.trace_lm(outer); po, so_, ito, no, termo = O.trace_lm(outer)
            print(f"   outer {outer}: sums rel err {np.abs(sg_ - so_).max() / np.abs(so_).max():.2e} pose diff {np.abs(pg - po).max():.2e} iters {ng}/{no}")
    for cls in (0, 1):
        a, b = R.map_download(cls), O.get_map(cls)
        same = a.shape == b.shape and np.array_equal(bits(a), bits(b))
        nd = -1 if a.shape != b.shape else int((bits(a) != bits(b)).any(1).sum())
        print(f"   map cls {cls}: {a.shape[0]}/{b.shape[0]} identical {same} differing points {nd}")
print("launches", R.launch_count())
