"""bench.py --config advection2d | advection_amr | sparse3d: the BASELINE.json configurations
other than the burgers headline (configs[0], [2], [3]), one JSON line each in bench.py's format:
zone-cycles/s of the whole cycle (remesh included on adaptive meshes) plus the roofline of every
kernel class from the launches themselves (work units of the launch x algorithmic bytes per
unit / CUDA-event time, peak = MEASURED_PEAKS.json).

  advection2d    example/advection 2-D 256^2, 32^2 blocks, uniform, nghost 2 (configs[0])
  advection_amr  example/advection 3-D 128^3 base, 16^3 blocks, adaptive, 3 levels (configs[2]):
                 restriction / prolongation ghost fill, flux correction, remesh
  sparse3d       example/sparse_advection 3-D 256^3, 32^3 blocks, 4 sparse fields allocated /
                 deallocated as the blobs move (configs[3]; the reference aborts in 3-D)
Parity of these paths is pinned by tests/ (reference dumps, oracle); this file only measures."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def bytes_per_unit(kernel, ndim):
    """algorithmic bytes per unit of work (DESIGN.md "Kernels")"""
    return {
        # per value moved: read once, written once
        "copy_kernel": 16, "halo_uniform_kernel": 16, "pack_kernel": 16, "unpack_kernel": 16,
        # per coarse cell and component: 2^ndim fine reads + 1 coarse write
        "restrict_kernel": 8 * (2 ** ndim + 1),
        # per coarse cell and component: 1 coarse read (stencil neighbours are reuse) + 2^ndim writes
        "prolongate_kernel": 8 * (2 ** ndim + 1),
        # per coarse face value: 2^(ndim-1) fine faces read + 1 coarse face written
        "flux_correct_kernel": 8 * (2 ** (ndim - 1) + 1),
        # per ghost value: one read, one write
        "apply_bc_kernel": 16,
        # per cell and component: read u, write ndim face fluxes
        "advection_flux_kernel": 8 * (1 + ndim),
        # per cell and component: read ndim fluxes (the far face is the neighbour's near face) + write
        "flux_div_kernel": 8 * (ndim + 1),
        # per value: two reads, one write
        "weighted_sum_kernel": 24,
    }.get(kernel)


def run(args, emit):
    import torch

    from parthenon_b200 import capi, host
    from bench import ClockSampler, UNIT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    if args.gpus != 1:
        raise SystemExit("--config runs are single-GPU lines")
    L = capi.lib()
    capi.check(L.pb2_set_device(0))
    extra = dict(kv.split("=", 1) for kv in args.set)
    if args.config == "advection2d":
        ndim, app = 2, "advection"
        ov = {"parthenon/mesh/nx1": 256, "parthenon/mesh/nx2": 256, "parthenon/meshblock/nx1": 32,
              "parthenon/meshblock/nx2": 32, "Advection/profile": "hard_sphere"}
        workload = ("example/advection 2D 256^2 mesh, 32^2 meshblocks (64 blocks), uniform, nghost 2, "
                    "hard sphere, rk2, periodic")
        field = "advected"
    elif args.config == "advection_amr":
        ndim, app = 3, "advection"
        ov = {"parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 3,
              "parthenon/mesh/derefine_count": 10, "Advection/profile": "hard_sphere"}
        for d in (1, 2, 3):
            ov[f"parthenon/mesh/nx{d}"] = 128
            ov[f"parthenon/meshblock/nx{d}"] = 16
        workload = ("example/advection 3D 128^3 base mesh, 16^3 meshblocks, adaptive refinement with "
                    "3 levels (restriction / prolongation ghost fill, flux correction, remesh every "
                    "cycle), hard sphere, rk2, periodic")
        field = "advected"
    else:
        ndim, app = 3, "sparse_advection"
        ov = {}
        for d in (1, 2, 3):
            ov[f"parthenon/mesh/nx{d}"] = 256
            ov[f"parthenon/meshblock/nx{d}"] = 32
        workload = ("example/sparse_advection 3D 256^3 mesh, 32^3 meshblocks (512 blocks), 4 sparse "
                    "fields allocated / deallocated as the blobs move, rk2, periodic (the reference "
                    "supports 2-D only: no reference parity in 3-D, see DESIGN.md)")
        field = None
    ov.update(extra)
    sim = host.Simulation(app=app, overrides=ov)
    sim.pre_execute()
    info = sim.info()
    nb_cells = 1
    for d in (1, 2, 3)[:ndim]:
        nb_cells *= int(ov.get(f"parthenon/meshblock/nx{d}", 16))

    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        sim.cycle()
    sim.sync()
    capi.profile(reset=True)
    capi.profile(enable=True)
    n0 = capi.launch_count()
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    capi.check(L.pb2_event_create(C.byref(ev0)))
    capi.check(L.pb2_event_create(C.byref(ev1)))
    torch.cuda.synchronize()
    blocks = []
    alloc = []
    w0 = time.time()
    capi.check(L.pb2_event_record(ev0, sim.stream))
    for _ in range(args.steps):
        blocks.append(sim.info()["nbtotal"])
        sim.cycle()
    capi.check(L.pb2_event_record(ev1, sim.stream))
    sim.sync()
    torch.cuda.synchronize()
    w1 = time.time()
    ms = C.c_float()
    capi.check(L.pb2_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    sec = ms.value * 1e-3
    launches = capi.launch_count() - n0
    capi.profile(enable=False)
    prof, work = capi.profile(), capi.profile_work()
    windows = [(w0, w1)]
    if sampler.proc is not None and sampler.count(w0, w1) < 8:
        x0 = time.time()
        while time.time() - x0 < 1.5:
            sim.cycle()
        sim.sync()
        windows.append((x0, time.time()))
    clocks = sampler.stop(windows)
    clocks["samples_in_timed_region"] = sampler.count(w0, w1)
    if app == "sparse_advection":
        alloc = [int(sim.allocation("base", f"sparse_{f}").sum()) for f in range(4)]
    zone_cycles = float(sum(blocks)) * nb_cells
    value = zone_cycles / sec

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    kernels = {}
    for name, (kms, n) in prof.items():
        ent = {"ms_total": kms, "launches": n, "share": kms / (1e3 * sec),
               "ms_per_launch": kms / n}
        bpu, w = bytes_per_unit(name, ndim), work.get(name, 0.0)
        if bpu and w > 0:
            ent["work_per_launch"] = w / n
            ent["gbs"] = bpu * w / (kms * 1e-3) / 1e9
            ent["frac_of_hbm_peak"] = ent["gbs"] / peak
        kernels[name] = ent
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    roofline = None
    if dom and "gbs" in kernels[dom]:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["gbs"] / peak, "traffic": None,
                    "peak_source": peak_src, "share_of_step": kernels[dom]["share"]}
    device_ms = sum(v[0] for v in prof.values())
    line = {
        "metric": f"zone-cycles/s {args.config}", "value": value, "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (the example's own initial condition)",
        "config": {"workload": workload, "overrides": extra,
                   "blocks": {"first": blocks[0], "last": blocks[-1], "min": min(blocks),
                              "max": max(blocks)},
                   "l2": "small working sets fit the 126 MB L2 (advection2d: 0.6 MB per field); "
                         "per-kernel GB/s are then L2 figures, see DESIGN.md"},
        "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
        "launches_per_cycle": launches / args.steps,
        "device_busy_fraction": device_ms / (1e3 * sec),
        "roofline": roofline, "kernels": kernels, "cpu_baseline": None,
    }
    if alloc:
        line["config"]["allocated_block_field_pairs"] = {"now": sum(alloc), "of": 4 * info["nbtotal"]}
    emit(line)
    sim.close()
