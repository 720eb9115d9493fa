import sys, time
sys.path.insert(0, "/root/repo")
from parthenon_b200 import host, capi
ov = {"parthenon/mesh/nghost": 4, "parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 2,
      "burgers/num_scalars": 8, "burgers/recon": "weno5", "pb2/math": "fast"}
for d in (1, 2, 3):
    ov[f"parthenon/mesh/nx{d}"] = 128
    ov[f"parthenon/meshblock/nx{d}"] = 16
sim = host.Simulation(overrides=ov); sim.pre_execute()
for _ in range(20): sim.cycle()
sim.sync()
capi.profile(reset=True); capi.profile(enable=True)
n0 = capi.launch_count(); t0 = time.time()
N = 100
for _ in range(N): sim.cycle()
sim.sync(); wall = time.time() - t0
capi.profile(enable=False)
p = capi.profile()
print("wall ms/cycle", 1e3 * wall / N, "launches/cycle", (capi.launch_count() - n0) / N, "blocks", sim.info()["nbtotal"])
tot = 0
for k, (ms, n) in sorted(p.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {ms / N:8.3f} ms/cycle  {n / N:6.1f} launches/cycle")
    tot += ms
print("sum kernels ms/cycle", tot / N)
# host-side split of a cycle: Step (both stages + tagging) vs LoadBalancing/AMR + new dt
ts, tr, nre = 0.0, 0.0, 0
n_before = sim.info()["nbtotal"]
for _ in range(N):
    a = time.time(); sim.step(); sim.sync(); b = time.time(); sim.regrid(); sim.sync(); c = time.time()
    ts += b - a; tr += c - b
    n_now = sim.info()["nbtotal"]; nre += n_now != n_before; n_before = n_now
print(f"step {1e3 * ts / N:.3f} ms/cycle, regrid+dt {1e3 * tr / N:.3f} ms/cycle, {nre} remeshes in {N} cycles, blocks {n_before}")
