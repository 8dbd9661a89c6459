import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("simulation-server_b200"); capi = importlib.import_module("simulation-server_b200.capi"); wl = importlib.import_module("simulation-server_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc, st, info = wl.long_vein(n)
sim = capi.Sim(sc, collect_stats=True, use_graph=False)
sim.upload_state(st)
prev = sim.stats(); done = 0
for target in (1, 5, 20, 50, 100, 150, 200):
    sim.step(target - done); done = target
    s = sim.stats()
    pos = np.stack(sim.download(capi.PARTICLE_POS), 1); r = np.hypot(pos[:, 0], pos[:, 2])
    prof = sim.profile_steps(1); done += 1
    s2 = sim.stats()
    print(f"step {done}: tri_tests/step {s2['triangle_tests']-s['triangle_tests']:.3e} pair_tests/step {s2['pair_tests']-s['pair_tests']:.3e} pair_hits {s2['pair_hits']-s['pair_hits']} vein_hits {s2['vein_hits']-s['vein_hits']} teleported {s2['teleported_cells']} r>38: {(r>38).mean():.3f} r>44: {(r>44).mean():.3f} r>50 {(r>50).mean():.3f}  vein_coll {prof['vein_collisions'][0]*1e3:.0f}us " + " ".join(f"{k} {prof[k][0]*1e3:.0f}us" for k in ('vein_cull_cells', 'vein_filter', 'vein_masking', 'wall_rebuild', 'particle_collisions', 'springs') if k in prof) + f" rebuilds {s2.get('wall_rebuilds', 0)}")
