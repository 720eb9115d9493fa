#!/usr/bin/env python
"""bench.py — zone-cycles/s of the burgers 3-D benchmark (Parthenon-VIBE) on B200.

    python bench.py --gpus N --steps K --warmup W          (N = 1; torchrun for N > 1)
    python bench.py --impl reference ...                   (CPU arm: the reference's own build)
    python bench.py --scaling strong --gpus N              (512^3 in total at every N)
    python bench.py --config advection2d|advection_amr|sparse3d   (the other BASELINE configs)

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
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Rank 0 must print exactly ONE JSON line on stdout, and NCCL (NCCL_DEBUG=INFO, which the
# driver may set to see the communicator's ranks) writes its log to stdout too.  So file
# descriptor 1 is pointed at stderr for the whole process — NCCL_DEBUG is left alone — and the
# JSON line goes to the original stdout saved here.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)

import numpy as np  # noqa: E402

UNIT = "zone-cycles/s"
NCOMP = 11

# mesh (in cells) per GPU count: 256^3 per GPU, Morton-contiguous octants
MESH = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def metric_name(mesh, per_gpu):
    m = "x".join(str(x) for x in mesh) if len(set(mesh)) > 1 else f"{mesh[0]}^3"
    return (f"zone-cycles/s burgers 3D ({m}{' per GPU' if per_gpu else ''}, 32^3 meshblocks, "
            f"weno5, 8 scalars, RK2)")


def overrides(mesh, block, math, fused=True, extra=None):
    ov = {"parthenon/mesh/nghost": 4, "parthenon/mesh/refinement": "none",
          "burgers/num_scalars": NCOMP - 3, "burgers/recon": "weno5", "pb2/math": math,
          "pb2/fused_stage": "true" if fused else "false"}
    for d in range(3):
        ov[f"parthenon/mesh/nx{d + 1}"] = mesh[d]
        ov[f"parthenon/meshblock/nx{d + 1}"] = block
    if extra:
        ov.update(extra)
    return ov


# ---- algorithmic bytes per unit of work of every kernel class (DESIGN.md "Kernels") ---------
# unit = zone for the stencil kernels, value (one double moved) for the ghost kernels; the
# number of units of a launch comes from the launch itself (pb2_profile_get_work)
def algorithmic_bytes_per_unit(kernel, ghost_per_zone):
    c = NCOMP
    table = {
        # read U (C) + write one flux direction (C)
        "flux_x_kernel": 8 * (2 * c), "flux_march_kernel<y>": 8 * (2 * c),
        "flux_march_kernel<z>": 8 * (2 * c),
        # read 3 fluxes (3C) + u (C) + base (C, second stage only: 0.5 on average) + write
        # out (C) + derived (1)
        "update_kernel": 8 * (3 * c + c + 0.5 * c + c + 1),
        # x sweep: read u (C) + base (0.5 C on average: second stage only) + write out (C)
        "sweep_x_kernel": 8 * (c + 0.5 * c + c), "sweep_xpair_kernel": 8 * (c + 0.5 * c + c),
        # y sweep: read u (C) + read-modify-write out (2C)
        "sweep_march_kernel<y>": 8 * (3 * c), "sweep_chunk_kernel<y>": 8 * (3 * c),
        # z sweep: the same + derived (1) (+ the pushed ghosts, C G/N values written, when the
        # ghost exchange is folded into it)
        "sweep_march_kernel<z>": 8 * (3 * c + 1),
        "sweep_chunk_kernel<z>": 8 * (3 * c + 1 + c * ghost_per_zone),
        # per value moved: read once, written once
        "copy_kernel": 16, "halo_uniform_kernel": 16, "pack_kernel": 16, "unpack_kernel": 16,
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
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if any(a <= t <= b for a, b in windows)]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "power_w_median": float(np.median(pw)) if pw else None,
                "reasons": sorted(reasons)}


# ---- CPU arms -----------------------------------------------------------------------------------
def host_threads():
    return len(os.sched_getaffinity(0))


def cpu_oracle_rate(nx, block, steps, warmup):
    """the reference's algorithm restated on the CPU (oracle/, test infrastructure), all host
    threads: zone-cycles/s on a bounded sample of the workload"""
    import oracle
    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which
    # would otherwise turn the reference arm into a single-thread run
    oracle.lib().orc_set_num_threads(host_threads())
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


def reference_binaries():
    out = {}
    for layout in ("simdfor", "mdrange"):
        p = os.path.join(REF_DIR, f"burgers-benchmark.{layout}")
        if os.path.isfile(p) and os.access(p, os.X_OK):
            out[layout] = p
    return out


def run_reference_binary(exe, nx, block, steps, warmup, threads, timeout):
    """the UNMODIFIED reference (baseline/_ref, built by scripts/build_reference.sh: Kokkos
    OpenMP) on the benchmark deck: returns zone-cycles/wallsecond as the reference itself
    reports it (driver.cpp:57-63; its timer restarts after perf_cycle_offset cycles)"""
    import tempfile
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(threads), "OMP_PROC_BIND": "spread",
                "OMP_PLACES": "threads"})
    args = [exe, "-i", os.path.join(REF_DIR, "burgers.pin")]
    for d in (1, 2, 3):
        args += [f"parthenon/mesh/nx{d}={nx}", f"parthenon/meshblock/nx{d}={block}"]
    args += ["parthenon/mesh/refinement=none", "parthenon/mesh/numlevel=1",
             "parthenon/mesh/nghost=4", "burgers/num_scalars=8", "burgers/recon=weno5",
             f"parthenon/time/nlim={steps + warmup}", f"parthenon/time/perf_cycle_offset={warmup}",
             "parthenon/time/tlim=1e30", "parthenon/time/ncycle_out=1",
             "parthenon/output0/dt=-1", "parthenon/output1/dt=-1"]
    with tempfile.TemporaryDirectory() as tmp:  # the .hst and friends land in the cwd
        t0 = time.perf_counter()
        r = subprocess.run(args, cwd=tmp, env=env, capture_output=True, text=True,
                           timeout=timeout)
        wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"reference exited {r.returncode}: {r.stdout[-400:]} {r.stderr[-400:]}")
    m = re.search(r"zone-cycles/wallsecond\s*=\s*([0-9.eE+-]+)", r.stdout)
    if not m:
        raise RuntimeError("no zone-cycles/wallsecond in the reference's output: " + r.stdout[-400:])
    return float(m.group(1)), wall


def best_reference_layout(block, threads):
    """both loop layouts on a small mesh; the faster one is used for the timed run"""
    rates = {}
    for layout, exe in reference_binaries().items():
        try:
            rates[layout] = run_reference_binary(exe, 64, block, 3, 1, threads, 300)[0]
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"reference layout {layout} failed: {e}\n")
    if not rates:
        return None, rates
    return max(rates, key=rates.get), rates


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on this box's host
    cores, same deck as the GPU arm's N = 1 workload (256^3, 32^3 blocks)"""
    if rank != 0:
        return
    threads = host_threads()
    nx = args.nx or 256
    layout, rates = best_reference_layout(args.block, threads)
    kind = "reference"
    note = None
    if layout is not None:
        # bounded run: one cycle at 256^3 costs ~nx^3 / rate seconds; keep the whole run under
        # ~4 minutes by shrinking the mesh only if it must (stated in `sample`)
        per_cycle = nx ** 3 / rates[layout]
        if per_cycle * (args.steps + args.warmup) > 420:
            note = f"{nx}^3 would need {per_cycle:.1f} s per cycle on {threads} threads"
            nx //= 2
        try:
            rate, wall = run_reference_binary(reference_binaries()[layout], nx, args.block,
                                              args.steps, args.warmup, threads, 1500)
            sec = nx ** 3 / rate
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"reference run failed ({e}); falling back to the oracle port\n")
            layout = None
    if layout is None:
        kind = "port"
        nx = args.cpu_sample_nx
        rate, sec, threads = cpu_oracle_rate(nx, args.block, args.steps, args.warmup)
    sample = (f"{nx}^3 mesh of {args.block}^3 blocks ({(nx // args.block) ** 3} blocks), same deck, "
              f"{args.steps} cycles after {args.warmup} warm-up; "
              + (f"unmodified reference (Kokkos OpenMP, {layout} loops; 64^3 calibration "
                 f"{ {k: round(v) for k, v in rates.items()} } zc/s), {threads} threads, "
                 f"rate as printed by the reference's driver"
                 if kind == "reference" else "oracle/pb2_oracle.c with OpenMP")
              + (f"; {note}" if note else ""))
    line = {
        "impl": "reference", "metric": metric_name((256,) * 3, True), "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic burgers initial condition of the reference deck)",
        "config": {"workload": f"benchmarks/burgers 3D {nx}x{nx}x{nx} mesh, {args.block}^3 "
                               f"meshblocks ({(nx // args.block) ** 3} blocks), uniform, nghost 4, "
                               f"weno5, 8 scalars, rk2, periodic — CPU, " + sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---- parity of the benchmarked configuration, inside the bench run ------------------------------
def block_crcs(u):
    import zlib
    return [zlib.crc32(np.ascontiguousarray(u[b]).tobytes()) for b in range(u.shape[0])]


def parity_check(host, dist, rank, world, new_nccl_id, block, extra):
    """The kernels and the partition this run times, checked in the run itself on a smaller mesh
    of the same block shape: (1) the N-rank result against the SAME deck on one rank — must be
    bit-identical (every ghost that crossed NVLink equals the one a same-device copy delivers);
    (2) the fast arithmetic against pb2/math = strict (bit-exact to the reference, see tests/):
    per-element relative difference, floor 1e-3 of the component's largest magnitude."""
    cycles = 10 if world > 1 else 25
    mesh = tuple(x // 2 for x in MESH[world])  # 128^3 per GPU
    out = {"deck": f"{mesh[0]}x{mesh[1]}x{mesh[2]} mesh of {block}^3 blocks, {cycles} cycles, "
                   f"same deck and kernels as the timed run"}
    crcs = {}
    if world > 1:
        sim = host.Simulation(overrides=overrides(mesh, block, "fast", True, extra), rank=rank,
                              nranks=world, nccl_id=new_nccl_id())
        sim.pre_execute()
        sim.cycle(cycles)
        mine = block_crcs(sim.get_field("base", "U"))
        first = sim.info()["first_gid"]
        sim.close()
        parts = [None] * world
        dist.all_gather_object(parts, (first, mine))
        for f, lst in parts:
            for i, c in enumerate(lst):
                crcs[f + i] = c
    if rank != 0:
        return None
    ref = {}
    for math in ("fast", "strict"):
        s1 = host.Simulation(overrides=overrides(mesh, block, math, True, extra))
        s1.pre_execute()
        s1.cycle(cycles)
        ref[math] = s1.get_field("base", "U")
        s1.close()
    one = block_crcs(ref["fast"])
    if world > 1:
        out["n_rank_bit_identical_to_1_rank"] = bool(
            len(crcs) == len(one) and all(crcs[g] == one[g] for g in range(len(one))))
        out["blocks_compared"] = len(one)
    cmax = np.abs(ref["strict"]).max(axis=(0, 2, 3, 4), keepdims=True)
    rel = np.abs(ref["fast"] - ref["strict"]) / np.maximum(np.abs(ref["strict"]), 1e-3 * cmax)
    out["max_rel_fast_vs_strict"] = float(rel.max())
    out["tolerance"] = 1e-12
    out["ok"] = bool(out["max_rel_fast_vs_strict"] <= 1e-12 and
                     out.get("n_rank_bit_identical_to_1_rank", True))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: BASELINE configs[4] as named, 512^3 in total at every N")
    ap.add_argument("--config", default="burgers",
                    choices=["burgers", "advection2d", "advection_amr", "sparse3d"])
    ap.add_argument("--block", type=int, default=32)
    ap.add_argument("--nx", type=int, default=0, help="cubic mesh override")
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--unfused", action="store_true", help="reference-shaped task list")
    ap.add_argument("--set", action="append", default=[], metavar="block/key=value",
                    help="extra input-deck override (repeatable)")
    ap.add_argument("--cpu-sample-nx", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
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
    if args.config != "burgers":
        from scripts import bench_configs
        bench_configs.run(args, emit)
        return

    import torch
    import torch.distributed as dist

    from parthenon_b200 import capi, host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    L = capi.lib()
    capi.check(L.pb2_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_nccl_id():
        """one ncclUniqueId per communicator (per Simulation), made on rank 0"""
        if world == 1:
            return None
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.check(L.pb2_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    nccl_id = new_nccl_id()

    extra = dict(kv.split("=", 1) for kv in args.set)
    if args.nx:
        mesh = (args.nx,) * 3
    elif args.scaling == "strong":
        mesh = (512, 512, 512)
    else:
        mesh = MESH[args.gpus]
    zones = mesh[0] * mesh[1] * mesh[2]
    sim = host.Simulation(overrides=overrides(mesh, args.block, args.math, not args.unfused, extra),
                          rank=rank, nranks=world, nccl_id=nccl_id)
    info = sim.info()
    sim.pre_execute()
    halo_mode = sim.exchange_mode("base")  # (every rank asks: building the tables is collective)

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
    work = capi.profile_work()
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
        del hin, hout
    sim.close()

    # ---- parity of what was just timed ------------------------------------------------------
    parity = None
    if not args.no_parity and args.math == "fast" and not args.unfused:
        parity = parity_check(host, dist, rank, world, new_nccl_id, args.block, extra)

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

    def side_file(name):
        p = os.path.join(ROOT, "profiles", name)
        return json.load(open(p)) if os.path.exists(p) else {}

    traffic = side_file("dram_traffic.json")
    fp64_inst = side_file("fp64_inst.json")
    fp64_peak = capi.fp64_peak_tflops()
    ghost_per_zone = ((args.block + 8) ** 3 - args.block ** 3) / args.block ** 3
    kernels = {}
    for name, (ms, n) in prof.items():
        bpu = algorithmic_bytes_per_unit(name, ghost_per_zone)
        ent = {"ms_total": ms, "launches": n, "share": ms / (1e3 * sec)}
        w = work.get(name, 0.0)
        if bpu and w > 0:
            # bytes of THESE launches (their own block lists / region tables) over their time
            ent["work_per_launch"] = w / n
            ent["gbs"] = bpu * w / (ms * 1e-3) / 1e9
        kernels[name] = ent
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    roofline = None
    if dom and "gbs" in kernels[dom]:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["gbs"] / peak, "traffic": traffic.get(dom),
                    "peak_source": peak_src, "share_of_step": kernels[dom]["share"],
                    "note": "FP64-pipe bound kernel (11 WENO5-Z reconstructions per zone and "
                            "sweep); the HBM fraction is reported as the contract asks, the "
                            "binding roof is `fp64`, see DESIGN.md"}
        fi = fp64_inst.get(dom)
        if fi:
            # FP64 thread-instructions per zone (ncu smsp__inst_executed_pipe_fp64 of one launch
            # / zones, profiles/): 2 flops each, the convention of the FMA peak measured beside it
            tf = 2.0 * fi * work[dom] / (prof[dom][0] * 1e-3) / 1e12
            roofline["fp64"] = {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s (2 per FP64 "
                                "instruction)", "frac": tf / fp64_peak,
                                "fp64_inst_per_zone": fi,
                                "peak_source": "pb2_measure_fp64_peak, this run"}
    # whole-cycle HBM roofline under the compulsory-traffic model A_zc
    a_zc = 8 * ((5 * NCOMP + 2) + 4 * NCOMP * ghost_per_zone)
    cycle_roof = {"algorithmic_bytes_per_zone_cycle": a_zc, "achieved_gbs": value / world * a_zc / 1e9,
                  "frac_of_hbm_peak": value / world * a_zc / 1e9 / peak,
                  "fp64_peak_tflops_measured": fp64_peak}
    if fp64_inst:
        per_zc = 2 * sum(fp64_inst.get(k, 0) for k in
                         ("sweep_xpair_kernel", "sweep_chunk_kernel<y>", "sweep_chunk_kernel<z>"))
        if per_zc:
            cycle_roof["fp64_inst_per_zone_cycle"] = per_zc
            cycle_roof["frac_of_fp64_peak"] = 2.0 * per_zc * value / world / 1e12 / fp64_peak

    # BASELINE.json's second metric: ghost-exchange GB/s vs the HBM peak.  A_exch = 2 G C 8 B per
    # block: every ghost value read once from its owner and written once.
    ghost = None
    hk = [k for k in ("halo_uniform_kernel", "copy_kernel") if k in kernels and "gbs" in kernels[k]]
    if hk:
        k = hk[0]
        ghost = {"kernel": k, "gbs": kernels[k]["gbs"], "frac_of_hbm_peak": kernels[k]["gbs"] / peak,
                 "bytes_per_exchange_per_gpu": 16 * kernels[k]["work_per_launch"],
                 "ms_per_exchange": prof[k][0] / prof[k][1]}
    elif "sweep_chunk_kernel<z>" in kernels:
        ghost = {"kernel": "sweep_chunk_kernel<z>", "gbs": None,
                 "note": "same-device ghosts are stored by the last sweep of the stage (no exchange "
                         "pass, no re-read of the field): the exchange costs "
                         f"{8 * NCOMP * ghost_per_zone:.1f} extra bytes written per zone inside a "
                         "kernel bound by the FP64 pipe; stand-alone ghost fill (init / remesh): "
                         "halo_uniform_kernel, profiles/"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        layout, rates = best_reference_layout(args.block, threads)
        nx = args.cpu_sample_nx
        if layout is not None:
            try:
                rate, wall = run_reference_binary(reference_binaries()[layout], nx, args.block, 6, 2,
                                                  threads, 600)
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                       "sample": f"{nx}^3 mesh of {args.block}^3 blocks, same deck, 6 cycles after 2 "
                                 f"warm-up ({wall:.1f} s wall incl. start-up); unmodified reference, "
                                 f"Kokkos OpenMP, {layout} loops (64^3 calibration: "
                                 f"{ {k: round(v) for k, v in rates.items()} } zc/s)"}
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"reference baseline failed: {e}\n")
        if cpu is None:
            rate, csec, cores = cpu_oracle_rate(nx, args.block, 3, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{nx}^3 mesh of {args.block}^3 blocks, same deck, 3 cycles after 1 "
                             f"warm-up ({csec:.2f} s per cycle); oracle/pb2_oracle.c with OpenMP"}

    per_gpu = args.scaling == "weak" and not args.nx
    line = {
        "metric": metric_name((256,) * 3 if per_gpu else mesh, per_gpu),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic burgers initial condition of the reference deck)",
        "config": {
            "workload": f"benchmarks/burgers 3D {mesh[0]}x{mesh[1]}x{mesh[2]} mesh, {args.block}^3 "
                        f"meshblocks ({info['nbtotal']} blocks, {info['nblocks']} per GPU), uniform, "
                        f"nghost 4, weno5, 8 scalars, rk2, periodic",
            "math": args.math, "task_list": "reference-shaped" if args.unfused else "fused stage",
            "overrides": extra,
            "l2": "working set (>= 5.8 GB per GPU) exceeds the 126 MB L2; no flush needed",
            "partition": f"Morton-contiguous gid ranges, {world} rank(s)",
            "inter_gpu_halo": halo_mode},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "ghost_exchange": ghost, "cycle_roofline": cycle_roof,
        "kernels": kernels, "parity": parity,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
