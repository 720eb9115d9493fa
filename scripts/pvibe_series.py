import sys, time
sys.path.insert(0, "/root/repo")
from parthenon_b200 import host
ov = {"parthenon/mesh/nghost": 4, "parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 2,
      "parthenon/time/tlim": 0.4, "burgers/num_scalars": 8, "burgers/recon": "weno5", "pb2/math": "fast"}
for d in (1, 2, 3):
    ov[f"parthenon/mesh/nx{d}"] = 128
    ov[f"parthenon/meshblock/nx{d}"] = 16
for rep in range(2):
    sim = host.Simulation(overrides=ov); sim.pre_execute(); sim.sync()
    c = 0; t0 = time.time(); nb = sim.info()["nbtotal"]; nre = 0
    while sim.time < 0.4:
        sim.cycle(); c += 1
        n = sim.info()["nbtotal"]; nre += n != nb; nb = n
        if c % 200 == 0:
            sim.sync(); t1 = time.time()
            print(rep, c, nb, f"{1e3 * (t1 - t0) / 200:.2f} ms/cycle", nre, "remeshes", f"dt {sim.dt:.3e}", flush=True)
            t0 = t1; nre = 0
    sim.close()
