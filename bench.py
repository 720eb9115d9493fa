#!/usr/bin/env python
"""bench.py — zone-cycles/s of the burgers 3-D benchmark (Parthenon-VIBE) on B200.

    python bench.py --gpus N --steps K --warmup W          (N = 1; torchrun for N > 1)
    python bench.py --impl reference ...                   (CPU arm: the oracle port)

A step is ONE cycle of the hot path (RK2: two stages of reconstruct + flux + update + ghost
exchange) over the whole mesh.  Workload: BASELINE.json configs[1] at N = 1 (256^3 mesh of 32^3
blocks, nghost 4, weno5, 8 scalars, periodic); per-GPU work is held at 256^3 for N > 1
(512x256x256, 512x512x256, 512^3 = configs[4]), i.e. weak scaling.  `value` times the loop with
the state resident in HBM; `e2e` adds, every step, the upload of the state from pinned host
memory and the read-back of the result.  See DESIGN.md "Measurement".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# rank 0 must print exactly one JSON line on stdout: keep NCCL's banner off it
os.environ["NCCL_DEBUG"] = os.environ.get("PB2_NCCL_DEBUG", "WARN")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "zone-cycles/s burgers 3D (256^3 per GPU, 32^3 meshblocks, weno5, 8 scalars, RK2)"
UNIT = "zone-cycles/s"
NCOMP = 11

# mesh (in cells) per GPU count: 256^3 per GPU, Morton-contiguous octants
MESH = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}


def overrides(mesh, block, math, fused=True):
    ov = {"parthenon/mesh/nghost": 4, "parthenon/mesh/refinement": "none",
          "burgers/num_scalars": NCOMP - 3, "burgers/recon": "weno5", "pb2/math": math,
          "pb2/fused_stage": "true" if fused else "false"}
    for d in range(3):
        ov[f"parthenon/mesh/nx{d + 1}"] = mesh[d]
        ov[f"parthenon/meshblock/nx{d + 1}"] = block
    return ov


# ---- algorithmic bytes per zone and launch of every kernel class (DESIGN.md "Kernels") ----
def algorithmic_bytes_per_zone(kernel, ghost_per_zone):
    c = NCOMP
    table = {
        # read U (C) + write one flux direction (C)
        "flux_x_kernel": 8 * (2 * c), "flux_march_kernel<y>": 8 * (2 * c),
        "flux_march_kernel<z>": 8 * (2 * c),
        # read 3 fluxes (3C) + u (C) + base (C, second stage only: 0.5 on average) + write
        # out (C) + derived (1)
        "update_kernel": 8 * (3 * c + c + 0.5 * c + c + 1),
        # x sweep: read u (C) + base (0.5 C on average: second stage only) + write out (C)
        "sweep_x_kernel": 8 * (c + 0.5 * c + c),
        # y / z sweeps: read u (C) + read-modify-write out (2C) (+ derived in the last one)
        "sweep_march_kernel<y>": 8 * (3 * c), "sweep_march_kernel<z>": 8 * (3 * c + 1),
        # every ghost value read once, written once
        "copy_kernel": 8 * 2 * c * ghost_per_zone,
        "halo_uniform_kernel": 8 * 2 * c * ghost_per_zone,
        "pack_kernel": 8 * 2 * c * ghost_per_zone, "unpack_kernel": 8 * 2 * c * ghost_per_zone,
    }
    return table.get(kernel)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def count(self, t0, t1):
        return sum(1 for t, _ in self.rows if t0 <= t <= t1)

    def stop(self, windows):
        """windows: [(t0, t1)] wall-clock intervals during which the timed loop (or the same
        loop continued, see main) was running on the GPU"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if any(a <= t <= b for a, b in windows)]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_oracle_rate(nx, block, steps, warmup):
    """the reference's algorithm restated on the CPU (oracle/, test infrastructure), all host
    threads: zone-cycles/s on a bounded sample of the workload"""
    import oracle
    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which
    # would otherwise turn the reference arm into a single-thread run
    oracle.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    nrb = nx // block
    m = oracle.Mesh(3, (block,) * 3, 4, (nrb,) * 3)
    B = oracle.Burgers(m, num_scalars=NCOMP - 3, recon="weno5", cfl=0.8)
    B.init()
    for _ in range(warmup):
        B.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        B.step()
    dt = time.perf_counter() - t0
    zones = nx ** 3
    return steps * zones / dt, dt / steps, oracle.lib().orc_get_max_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    nx = args.cpu_sample_nx
    rate, sec, cores = cpu_oracle_rate(nx, args.block, args.steps, args.warmup)
    sample = (f"{nx}^3 mesh of {args.block}^3 blocks ({(nx // args.block) ** 3} blocks), same deck, "
              f"{args.steps} cycles after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic burgers initial condition of the reference deck)",
        "config": {"workload": "benchmarks/burgers 3D, 32^3 meshblocks, uniform, weno5, 8 scalars, "
                               "rk2, periodic — CPU sample " + sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, default=32)
    ap.add_argument("--nx", type=int, default=0, help="cubic mesh override (N = 1 only)")
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--unfused", action="store_true", help="reference-shaped task list")
    ap.add_argument("--cpu-sample-nx", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (use torchrun)")

    import torch
    import torch.distributed as dist

    from parthenon_b200 import capi, host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    L = capi.lib()
    capi.check(L.pb2_set_device(local_rank))
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.check(L.pb2_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().numpy().tobytes())

    mesh = MESH[args.gpus] if not args.nx else (args.nx,) * 3
    zones = mesh[0] * mesh[1] * mesh[2]
    sim = host.Simulation(overrides=overrides(mesh, args.block, args.math, not args.unfused),
                          rank=rank, nranks=world, nccl_id=nccl_id)
    info = sim.info()
    sim.pre_execute()

    def barrier():
        sim.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        """CUDA events on the application's stream, barrier + sync on both sides, max over ranks"""
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        capi.check(L.pb2_event_create(C.byref(ev0)))
        capi.check(L.pb2_event_create(C.byref(ev1)))
        barrier()
        capi.check(L.pb2_event_record(ev0, sim.stream))
        for _ in range(steps):
            fn()
        if after is not None:
            # work still draining on the copy streams belongs to the timed region: wait for it
            # on the host, then close the interval on the (now idle) application stream
            after()
        capi.check(L.pb2_event_record(ev1, sim.stream))
        barrier()
        ms = C.c_float()
        capi.check(L.pb2_event_elapsed_ms(ev0, ev1, C.byref(ms)))
        L.pb2_event_destroy(ev0)
        L.pb2_event_destroy(ev1)
        t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3

    # ---- device-resident throughput -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        sim.cycle()
    capi.profile(reset=True)
    capi.profile(enable=True)
    n0 = capi.launch_count()
    w0 = time.time()
    sec = timed(sim.cycle, args.steps)
    w1 = time.time()
    launches = capi.launch_count() - n0
    capi.profile(enable=False)
    prof = capi.profile()
    windows = [(w0, w1)]
    clocks = None
    if rank == 0 and sampler.proc is not None and sampler.count(w0, w1) < 8 and world == 1:
        # nvidia-smi samples every 100 ms; a short timed region holds few samples, so the
        # SAME loop is continued (untimed) until the clocks under this load are on record
        x0 = time.time()
        while time.time() - x0 < 1.5:
            sim.cycle()
            sim.sync()
        windows.append((x0, time.time()))
    if rank == 0:
        clocks = sampler.stop(windows)
        clocks["samples_in_timed_region"] = sampler.count(w0, w1)
    value = args.steps * zones / sec

    # ---- end to end: state uploaded from pinned host memory and read back every step --------
    e2e = None
    if not args.no_e2e:
        # Every step takes one batch of state (interior cells of U) from PINNED HOST memory,
        # uploads it, fills ghosts, advances one RK2 cycle and returns the new state to host
        # memory.  Batches are independent (a stream of ensemble members), so the host API's two
        # lanes overlap the H2D of batch n+1 and the D2H of batch n-1 with the cycle of batch n
        # (full-duplex PCIe); the serial figure (one batch at a time, no overlap) is kept beside
        # it.  Every step moves nreal*8 bytes each way inside the timed region either way.
        nreal = sim.interior_size("base", "U")
        hin = [torch.empty(nreal, dtype=torch.float64).pin_memory() for _ in range(2)]
        hout = [torch.empty(nreal, dtype=torch.float64).pin_memory() for _ in range(2)]
        sim.download_interior("base", "U", hin[0].data_ptr(), nreal)
        sim.sync()
        hin[1].copy_(hin[0])

        def e2e_serial():
            sim.upload_interior("base", "U", hin[0].data_ptr(), nreal)
            sim.cycle()
            sim.download_interior("base", "U", hout[0].data_ptr(), nreal)

        state = {"i": 0}

        def e2e_piped():
            i = state["i"]
            # batch i+1 starts crossing PCIe now; batch i (prefetched one call earlier) joins
            # the application stream, cycles, and drains on the D2H stream
            sim.prefetch_interior("base", "U", hin[(i + 1) % 2].data_ptr(), nreal, (i + 1) % 2)
            sim.commit_interior("base", "U", i % 2)
            sim.cycle()
            sim.writeback_interior("base", "U", hout[i % 2].data_ptr(), nreal, i % 2)
            state["i"] = i + 1

        def drain():
            sim.lane_sync(0)
            sim.lane_sync(1)

        e2e_serial()
        ssec = timed(e2e_serial, min(args.steps, 5)) / min(args.steps, 5)
        sim.prefetch_interior("base", "U", hin[0].data_ptr(), nreal, 0)
        for _ in range(2):
            e2e_piped()
        esec = timed(e2e_piped, args.steps, after=drain)
        tot = torch.tensor([float(nreal)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": args.steps * zones / esec, "unit": UNIT,
               "h2d_bytes_per_step": int(8 * tot.item()), "d2h_bytes_per_step": int(8 * tot.item()),
               "ms_per_step": 1e3 * esec / args.steps,
               "serial_value": zones / ssec, "serial_ms_per_step": 1e3 * ssec,
               "what": "per step: pinned host U (interior cells) -> H2D -> scatter + ghost exchange "
                       "-> one RK2 cycle -> gather -> D2H into pinned host memory, through "
                       "pb2h_sim_prefetch_interior / _commit_interior / pb2h_sim_cycle / "
                       "_writeback_interior; independent batches double-buffered over two lanes so "
                       "the copies of neighbouring batches overlap the cycle (timed region ends "
                       "when the last D2H has landed); serial_* = same path one batch at a time "
                       "(pb2h_sim_upload_interior / _cycle / _download_interior)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    zones_rank = zones / world
    ghost_per_zone = ((args.block + 8) ** 3 - args.block ** 3) / args.block ** 3
    kernels = {}
    for name, (ms, n) in prof.items():
        bpz = algorithmic_bytes_per_zone(name, ghost_per_zone)
        ent = {"ms_total": ms, "launches": n, "share": ms / (1e3 * sec)}
        if bpz:
            ent["gbs"] = bpz * zones_rank / (ms / n * 1e-3) / 1e9
        kernels[name] = ent
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    roofline = None
    if dom and "gbs" in kernels[dom]:
        t = traffic.get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["gbs"] / peak, "traffic": t,
                    "peak_source": peak_src, "share_of_step": kernels[dom]["share"],
                    "note": "FP64-pipe bound kernel (33 WENO5-Z reconstructions per zone-stage); "
                            "HBM fraction reported as the contract asks, see DESIGN.md"}
    # whole-cycle HBM roofline under the compulsory-traffic model A_zc
    a_zc = 8 * ((5 * NCOMP + 2) + 4 * NCOMP * ghost_per_zone)
    cycle_roof = {"algorithmic_bytes_per_zone_cycle": a_zc, "achieved_gbs": value / world * a_zc / 1e9,
                  "frac_of_hbm_peak": value / world * a_zc / 1e9 / peak}

    # BASELINE.json's second metric: ghost-exchange GB/s vs the HBM peak.  A_exch = 2 G C 8 B per
    # block: every ghost value read once from its owner and written once.
    ghost = None
    hk = [k for k in ("halo_uniform_kernel", "copy_kernel") if k in kernels and "gbs" in kernels[k]]
    if hk:
        k = hk[0]
        ghost = {"kernel": k, "gbs": kernels[k]["gbs"], "frac_of_hbm_peak": kernels[k]["gbs"] / peak,
                 "bytes_per_exchange_per_gpu": 8 * 2 * NCOMP * ghost_per_zone * zones_rank,
                 "ms_per_exchange": prof[k][0] / prof[k][1],
                 "note": "same-device channels, sender interior -> receiver ghosts in one launch; "
                         "inter-GPU channels are pack_kernel / NCCL / unpack_kernel (see kernels)"}
    fp64_peak = capi.fp64_peak_tflops()
    cycle_roof["fp64_peak_tflops_measured"] = fp64_peak
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nx = args.cpu_sample_nx
        rate, csec, cores = cpu_oracle_rate(nx, args.block, 3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nx}^3 mesh of {args.block}^3 blocks, same deck, 3 cycles after 1 warm-up "
                         f"({csec:.2f} s per cycle); oracle/pb2_oracle.c with OpenMP"}

    line = {
        "metric": METRIC if not args.nx else METRIC.replace("256^3", f"{args.nx}^3"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic burgers initial condition of the reference deck)",
        "config": {
            "workload": f"benchmarks/burgers 3D {mesh[0]}x{mesh[1]}x{mesh[2]} mesh, {args.block}^3 "
                        f"meshblocks ({info['nbtotal']} blocks, {info['nblocks']} per GPU), uniform, "
                        f"nghost 4, weno5, 8 scalars, rk2, periodic",
            "math": args.math, "task_list": "reference-shaped" if args.unfused else "fused stage",
            "l2": "working set (>= 5.8 GB per GPU) exceeds the 126 MB L2; no flush needed",
            "partition": f"Morton-contiguous gid ranges, {world} rank(s)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "ghost_exchange": ghost, "cycle_roofline": cycle_roof,
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
