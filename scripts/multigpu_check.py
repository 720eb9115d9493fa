#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multigpu_check.py
Every rank advances its Morton-contiguous share of a burgers mesh through the C++ host
framework (NCCL halo slabs, allreduce-min dt) in STRICT arithmetic and compares its blocks
bit-for-bit with the CPU oracle run on the whole mesh; history columns are compared after the
cross-rank reduction."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from parthenon_b200 import capi, host  # noqa: E402
from tests.test_burgers_sim_gpu import burgers_overrides  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    capi.check(capi.lib().pb2_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def new_id():  # an NCCL unique id builds exactly one communicator
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.check(capi.lib().pb2_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    ok = True
    for nx, nrb, nscal, ncyc in ((8, 4, 2, 3), (16, 2, 8, 2)):
        nccl_id = new_id()
        m = oracle.Mesh(3, (nx,) * 3, 4, (nrb,) * 3)
        B = oracle.Burgers(m, num_scalars=nscal)
        B.init()
        sim = host.Simulation(overrides=burgers_overrides(nx, nrb, 4, nscal, "weno5", "strict", True),
                              rank=rank, nranks=world, nccl_id=nccl_id)
        info = sim.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        _, nl = sim.exchange_elements("base")
        sim.pre_execute()
        good = sim.dt == B.dt and np.array_equal(sim.get_field("base", "U"), B.U[lo:hi])
        for _ in range(ncyc):
            B.step()
            sim.cycle()
            good = good and sim.dt == B.dt and np.array_equal(sim.get_field("base", "U"), B.U[lo:hi])
        h = sim.history()
        good = good and np.allclose(h, B.history(), rtol=1e-13, atol=0)
        print(f"rank {rank}/{world}: mesh {nx * nrb}^3, blocks {lo}..{hi - 1}, "
              f"{nl} Reals per exchange through NCCL slabs: {'bit-exact' if good else 'MISMATCH'}",
              flush=True)
        ok = ok and good
        sim.close()
    # multilevel meshes across devices: restriction / prolongation regions whose neighbour lives
    # on another GPU and flux corrections through their own NCCL slabs, against the committed
    # dumps of the reference itself (tests/golden/)
    from tests import helpers as H
    from tests.test_host_topology import deck_overrides
    gold = os.path.join(ROOT, "tests", "golden")
    cases = [("burgers", "burgers_s16_b8_l2_weno5", 3, (16, 16, 16), (8, 8, 8), 4, "U", 2,
              {"burgers/num_scalars": 1, "burgers/recon": "weno5", "pb2/math": "strict"}),
             ("advection", "advection_s16_b8_l3_gaussian", 3, (16, 16, 16), (8, 8, 8), 2, "advected", 3,
              {"Advection/profile": "smooth_gaussian", "Advection/amp": 1.0, "Advection/vy": -0.7,
               "Advection/vz": 0.4})]
    for app, name, ndim, nxm, nxb, ng, field, ncyc, extra in cases:
        nccl_id = new_id()
        g = np.load(os.path.join(gold, name + ".npz"))
        leaves, nrb = H.leaves_from_bounds(g["bounds"], nxm, nxb)
        ov = deck_overrides(ndim, nxb, ng, nrb, refinement="static")
        ov.update(extra)
        sim = host.Simulation(app=app, overrides=ov, leaves=leaves, rank=rank, nranks=world,
                              nccl_id=nccl_id)
        info = sim.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        sim.pre_execute()
        good = np.array_equal(sim.get_field("base", field), g["U_0"][lo:hi])
        for c in range(1, ncyc + 1):
            sim.cycle()
            good = good and np.array_equal(sim.get_field("base", field), g[f"U_{c}"][lo:hi])
        print(f"rank {rank}/{world}: {name}, blocks {lo}..{hi - 1} of {g['U_0'].shape[0]} "
              f"(multilevel, flux correction over NCCL): {'bit-exact' if good else 'MISMATCH'}",
              flush=True)
        ok = ok and good
        sim.close()
    # adaptive meshes across devices: refinement flags gathered over ranks, blocks migrating
    # between GPUs through NCCL slabs on every remesh, against the reference's adaptive dumps
    for name, ndim, nxm, nxb, numlevel, dc, ncyc in (
            ("advection_a32_b8_l3_2d", 2, (32, 32, 1), (8, 8, 1), 3, 3, 40),
            ("advection_a32_b8_l2_3d", 3, (32, 32, 32), (8, 8, 8), 2, 2, 8)):
        nccl_id = new_id()
        g = np.load(os.path.join(gold, name + ".npz"))
        nrb = [nxm[d] // nxb[d] if d < ndim else 1 for d in range(3)]
        ov = deck_overrides(ndim, nxb, 2, nrb, refinement="adaptive")
        ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": dc,
                   "Advection/profile": "hard_sphere"})
        sim = host.Simulation(app="advection", overrides=ov, rank=rank, nranks=world,
                              nccl_id=nccl_id)
        sim.pre_execute()
        good, moved = True, set()
        for c in range(ncyc + 1):
            if c:
                sim.step()
            info = sim.info()
            lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
            leaves, _ = H.leaves_from_bounds(g[f"bounds_{c}"], nxm, nxb)
            locs = np.array([sim.block(b)["loc"] for b in range(info["nblocks"])])
            good = good and info["nbtotal"] == len(leaves) and np.array_equal(locs, leaves[lo:hi]) \
                and np.array_equal(sim.get_field("base", "advected"), g[f"U_{c}"][lo:hi])
            moved.add((lo, hi))
            if c:
                sim.regrid()
        print(f"rank {rank}/{world}: {name}, adaptive, {ncyc} cycles, {len(moved)} different "
              f"gid ranges on this rank: {'bit-exact' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
        sim.close()
    # face / edge / node fields across devices: channel pieces from ownership masks that sender
    # and receiver derive independently, element forms of restriction / prolongation on regions
    # whose neighbour lives on another GPU, against the reference's dumps
    for name, ndim, nx, nb, ng in H.TECOMM + H.TECOMM_MULTILEVEL:
        nccl_id = new_id()
        g = np.load(os.path.join(gold, name + ".npz"))
        full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
        leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
        static = len(set(g["meta"][:, 1])) > 1
        ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static" if static else "none")
        sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves if static else None,
                              rank=rank, nranks=world, nccl_id=nccl_id)
        info = sim.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        good = all(np.array_equal(sim.get_field("base", f), g[k][lo:hi])
                   for f, k in (("face", "U_0"), ("edge", "U_1"), ("node", "U_2")))
        print(f"rank {rank}/{world}: {name}, face / edge / node fields, blocks {lo}..{hi - 1} of "
              f"{g['U_0'].shape[0]}: {'bit-exact' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
        sim.close()
    # adaptive remesh of face / edge / node fields across devices: NOT supported (the 2-GPU run of
    # round 2 did not reproduce the reference's dumps); the build must refuse it loudly
    for name, ndim, nx, nb, numlevel in H.TEAMR[:1]:
        nccl_id = new_id()
        ov = deck_overrides(ndim, (nb,) * 3, 2, (nx // nb,) * 3, refinement="adaptive")
        ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": 2})
        refused = False
        try:
            sim = host.Simulation(app="tecomm", overrides=ov, rank=rank, nranks=world,
                                  nccl_id=nccl_id)
            for c in (1, 2, 3):
                sim.tag_and_remesh(c)
            sim.close()
        except RuntimeError as e:
            refused = "single device" in str(e)
        print(f"rank {rank}/{world}: {name}, adaptive face / edge / node fields on {world} devices: "
              f"{'refused as documented' if refused else 'NOT REFUSED'}", flush=True)
        ok = ok and refused
    # sparse fields across devices: null-message flags travel with the slabs, a block allocates a
    # field when a non-null message arrives from another GPU
    from tests.test_oracle_golden import SPARSE
    for name, kw, ncyc in SPARSE:
        nccl_id = new_id()
        g = np.load(os.path.join(gold, name + ".npz"))
        ov = {f"parthenon/sparse/{k}": v for k, v in kw.items()}
        sim = host.Simulation(app="sparse_advection", overrides=ov, rank=rank, nranks=world,
                              nccl_id=nccl_id)
        info = sim.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        sim.pre_execute()
        dumped = {int(c): i for i, c in enumerate(g["cycles"])}

        def state():
            return np.stack([np.where(sim.allocation("base", f"sparse_{f}")[:, None, None, None],
                                      sim.get_field("base", f"sparse_{f}")[:, 0], np.nan)
                             for f in range(4)], axis=1)

        good = np.array_equal(state(), g["U_0"][lo:hi], equal_nan=True)
        for c in range(1, ncyc + 1):
            sim.cycle()
            if c in dumped:
                good = good and np.array_equal(state(), g[f"U_{c}"][lo:hi], equal_nan=True)
        print(f"rank {rank}/{world}: {name}, sparse fields, blocks {lo}..{hi - 1}, {ncyc} cycles: "
              f"{'bit-exact, allocation included' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
        sim.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        raise SystemExit("multi-GPU parity FAILED")
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
