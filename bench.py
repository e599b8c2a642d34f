#!/usr/bin/env python
"""Benchmark of the CT-HYB hot path (BASELINE.json metric: QMC moves/s for the whole box).

    python bench.py --gpus N --steps K --warmup W            # B200 engine (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

A *step* is one accumulation block of the hot path: every chain of the population advances by NCorr = 100 moves
(one sweep kernel) and is measured once (one measurement kernel) -- the body of ctqmc_measure's iNMeas loop
(reference CTQMC.F90:8422-8581).  A *move* is one iteration of the iNCorr loop: one proposal of any type,
accepted or not (CTQMC.F90:8428-8434).  value = chains * NCorr * K * N / time.

Timed region of `value`: chain state resident in HBM, K steps, (N > 1: plus the single all-reduce of the raw
accumulators that ends an accumulation), CUDA events on the launching stream, max over ranks.
`e2e`: the same K accumulation blocks as ONE call of the reference plugin's solve (w2dynamics_b200/solver_sequence.py, the call a
user of the reference makes) with HOST arrays: init_paras, init_solver(u, F(tau), mu from host memory), set_taudiffmax,
ctqmc_warmup(0), ctqmc_measure(1,1) with Nmeas = K, every get_resultset array read back to the host, dest_ctqmc -- all inside the
timed region.  The chain population is kept thermalised across solves (option reuse_population = ReuseMCConfig for all chains): a
revived population is rebuilt on the device from its operator lists with the new F(tau).  `e2e.per_solve_of_one_step` is the same
with Nmeas = 1: the fixed cost of a solve charged to a single accumulation block.
`roofline`: dominant kernel k_sweep, algorithmic bytes counted in-kernel per move (DESIGN.md "Roofline accounting")
divided by the kernel's summed event time.  `cpu_baseline`: the oracle (CPU restatement of the reference algorithm;
the Fortran reference cannot be built in this image) on all host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": "C1: 1-band Hubbard, Bethe bath, half filling, U=2, beta=10 (4-state Fock space)",
    "C2": "C2: 2-orbital density-density Kanamori, beta=50, 16-state Fock space (determinant-dominated)",
    "C3": "C3: 3-orbital t2g full Kanamori, beta=40, 64-state Fock space, blocks d<=3",
    "C4": "C4: 5-orbital d-shell full Slater-Coulomb, beta=20, 1024-state Fock space, 132 blocks d<=28",
    "C5": "C5: 1-band Hubbard, beta=200 (expansion order in the hundreds)",
}
DEFAULT_CHAINS = {"C1": 65536, "C2": 65536, "C3": 32768, "C4": 32768, "C5": 32768}     # per GPU
# (C5: the CTA scheduler serves global moves a warp-load at a time -- with 28 chains per CTA (4096 chains) 91 % of the warp
#  samples waited at the CTA barrier for single global moves; 222 chains per CTA keep all warps busy: 2.2 -> 7.9 M moves/s)
# moves per chain from the empty configuration before anything is timed, BOTH arms.  Round 2, late: 6000 / 8000 moves left C4 at
# <k_f> = 2.65 (equilibrium 2.9, oracle chains: flat from 50 000 moves on) and C5 at 61 (equilibrium 74, flat from 10 000 on the CPU but
# the 65-window sweep of a 32 768-chain population needs longer), so the timed steps still grew and the later legs (time-to-error, e2e)
# ran on more expensive configurations than `value`: C5 104 -> 165 -> 199 ms per step within one bench run.
THERMALISE = {"C1": 4000, "C2": 12000, "C3": 12000, "C4": 40000, "C5": 24000}
METRIC = "QMC moves/sec (box)"
UNIT = "moves/s"


def fp64_peak():
    """FP64 tensor-pipe yardstick measured on this pool's B200 (cuBLAS DGEMM 8192^3, tools/fp64_peak.py): MEASURED_PEAKS.json
    holds no FP64 number."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_fp64_yardstick.json")))
        return float(d["dgemm_8192_tflops"]), "measured cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r1_fp64_yardstick.json)"
    except Exception:
        return 37.0, "nominal B200 FP64 (no measurement file)"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


class OraclePopulation:
    """The oracle port run the way the reference runs: independent Markov chains, every chain advancing the SAME number of
    moves (one chain per MPI rank, Nmeas x NCorr moves each; moves/s = ranks x moves / mean(time-sampling)).
    Chains differ ~30x in cost per move on C4 (outer block of 1 vs 28 states), so each host thread owns several chains, times
    each one on its own for a slice, and the figure is  threads x chains / sum_i(seconds per move of chain i)  -- the
    equal-moves throughput of the population, not a time-weighted average that favours the cheap chains.
    The chains run the oracle's per-operator ket / bra caches (oracle/chain.cpp `trace_cached`: the reference's cost structure,
    Trace.F90:47-49, 5384-5470, 6109-6209); the library is compiled on this host with -O3 -march=native."""

    def __init__(self, cfg, nthreads, chains_per_thread=8):
        from w2dynamics_b200 import problems
        from oracle.oracle import OracleChain
        from oracle import oracle as orc
        orc.build()
        self.p = problems.make_problem(cfg)
        self.ncorr = int(self.p.qmc.get("NCorr", 100))
        self.nthreads, self.cpt, self.therm = nthreads, chains_per_thread, THERMALISE[cfg]
        self.chains = [OracleChain(self.p, rng="philox", seed=43890, chain=1000 + i, fast_trace=2) for i in range(nthreads * chains_per_thread)]
        self._threads(lambda t: [self.chains[t * self.cpt + j].warmup(self.therm) for j in range(self.cpt)])
        self.reset()

    def _threads(self, fn):
        ths = [threading.Thread(target=fn, args=(t,)) for t in range(self.nthreads)]
        [t.start() for t in ths]; [t.join() for t in ths]

    def reset(self):
        self.moves = [0] * len(self.chains)
        self.busy = [0.0] * len(self.chains)
        self.wall = 0.0
        for c in self.chains:
            c.measure(0, reset=True)

    def run(self, seconds):
        """Every thread gives each of its chains seconds / chains_per_thread of sweeps + measurements."""
        def work(t):
            for j in range(self.cpt):
                i = t * self.cpt + j
                t0 = time.perf_counter(); stop = t0 + seconds / self.cpt
                while True:
                    self.chains[i].measure(1, reset=False)
                    self.moves[i] += self.ncorr
                    now = time.perf_counter()
                    if now >= stop:
                        break
                self.busy[i] += now - t0
        t0 = time.perf_counter()
        self._threads(work)
        self.wall += time.perf_counter() - t0

    def rate(self):
        """(equal-moves moves/s of the population, its standard error from the chain-to-chain scatter of the cost per move)"""
        import numpy as np
        c = np.array([b / m for b, m in zip(self.busy, self.moves) if m > 0])
        r = self.nthreads / c.mean()
        return r, r * c.std(ddof=1) / np.sqrt(len(c)) / c.mean()

    def time_to_error(self):
        """CPU counterpart of the engine's time_to_error: SEM of G(tau) over the chains of this run -> wall seconds to 1e-3."""
        import numpy as np
        g = np.array([c.results()["gtau"] for c, m in zip(self.chains, self.moves) if m > 0])
        mid = g.shape[-1] // 2
        gm, sem = np.abs(g.mean(0)[..., mid]).mean(), (g.std(0, ddof=1)[..., mid] / np.sqrt(g.shape[0])).mean()
        gma, sema = np.abs(g.mean(0)).mean(), (g.std(0, ddof=1) / np.sqrt(g.shape[0])).mean()
        if not (gm > 0 and sem > 0):
            return None
        return {"seconds_to_1e-3_relative": self.wall * (sem / (1e-3 * gm)) ** 2, "seconds_to_1e-3_relative_tau_average": self.wall * (sema / (1e-3 * gma)) ** 2,
                "sample_seconds": self.wall, "chains": int(g.shape[0]), "moves": int(sum(self.moves))}

    def sample(self):
        tw = sum(self.moves) / (sum(self.busy) / self.nthreads)
        return ("%d chains on %d host threads (%d per thread), %.1f s each of NCorr=%d sweeps + measurements (%d moves), thermalised %d moves; "
                "per-operator trace caches, g++ -O3 -march=native on this host; equal-moves throughput (time-weighted average would be %.3g moves/s)"
                % (len(self.chains), self.nthreads, self.cpt, sum(self.busy) / len(self.chains), self.ncorr, sum(self.moves), self.therm, tw))


def host_cores():
    """(physical cores, hardware threads) of this host"""
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count() or 1, os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1, os.cpu_count() or 1


def cpu_baseline(cfg, seconds):
    cores, nthreads = host_cores()
    pop = OraclePopulation(cfg, nthreads)
    pop.run(seconds)
    v, se = pop.rate()
    return {"value": v, "stderr": se, "unit": UNIT, "cores": cores, "threads": nthreads, "kind": "port", "sample": pop.sample(), "time_to_error": pop.time_to_error()}


def run_reference(args, cfg):
    """Reference arm: the reference's own CPU algorithm for the path.  The Fortran module cannot be built in this
    image (no Fortran compiler), so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, nthreads = host_cores()
    pop = OraclePopulation(cfg, nthreads)
    slice_s = args.slice_seconds if args.slice_seconds > 0 else min(60.0 / max(1, args.steps + args.warmup), 10.0)     # one step = one slice; whole run ~ a minute
    for _ in range(args.warmup):
        pop.run(slice_s)
    pop.reset()
    t0 = time.time()
    for _ in range(args.steps):
        pop.run(slice_s)
    dt = time.time() - t0
    val, se = pop.rate()
    sample = pop.sample()
    ncorr = pop.ncorr
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "ncorr": ncorr, "note": "oracle port of the reference algorithm (Fortran reference not buildable here), compiled -O3 -march=native on this host; per-operator ket / bra caches with two frontiers = the reference's cost structure (Trace.F90:47-49, 5384-5470, 6109-6209)"},
        "cpu_baseline": {"value": val, "stderr": se, "unit": UNIT, "cores": cores, "threads": nthreads, "kind": "port", "sample": sample, "time_to_error": pop.time_to_error()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


_DIST = {}


def run_engine(args, cfg):
    import ctypes as C
    import numpy as np
    import torch
    import __graft_entry__ as g
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from w2dynamics_b200 import problems
    from w2dynamics_b200.CTQMC import MCTQMC
    from w2dynamics_b200.solver_sequence import PluginSequence, qmc_defaults

    p = problems.make_problem(cfg)
    ncorr = int(p.qmc.get("NCorr", 100))
    nchains = args.chains or DEFAULT_CHAINS[cfg]
    if args.strong:
        nchains = max(1024, nchains // world)      # strong scaling: the population of one GPU is split over the ranks
    therm = args.thermalise if args.thermalise >= 0 else THERMALISE[cfg]
    eng = MCTQMC()
    eng.set_option("nchains", nchains); eng.set_option("nbins", 32); eng.set_option("device", local)
    eng.set_option("reuse_population", 1)      # dest_ctqmc parks the thermalised population; the next init_solver revives it (ReuseMCConfig for all chains)
    for key in ("cta_warps", "ops_cap", "scheduler", "batch_chains", "lu_groups"):
        if os.environ.get("W2B_" + key.upper()):
            eng.set_option(key, int(os.environ["W2B_" + key.upper()]))
    if world > 1 and not _DIST.get("on"):
        eng.enable_distributed()               # NCCL communicator inside the library (w2b_comm_init); ctqmc_measure ends with its all-reduce
        _DIST["on"] = True
    elif world > 1:
        eng._dist = (rank, world)
    L = eng._load()

    # ---- the reference plugin's call sequence (w2dynamics_b200/solver_sequence.py = CtHybSolver.set_problem / solve, proved
    # call-identical to the unmodified plugin by tests/test_reference_plugin.py): host arrays in, host arrays out
    qmc = qmc_defaults(); qmc.update(p.qmc)
    ftau5 = np.ascontiguousarray(p.ftau.transpose(4, 0, 1, 2, 3))
    screening = np.zeros_like(ftau5)

    solver = PluginSequence(eng, dict(qmc), 43890)             # ONE CtHybSolver for the whole run: init_rnd once, the random streams run on

    class plugin:                                              # one DMFT-style iteration = set_problem + solve (cthyb: 305, 314)
        def __init__(self, nwarm, nmeas):
            solver.qmc.update(Nwarmups=nwarm, Nmeas=nmeas)
            solver.param_string = "".join("%s = %s #\n" % (k, v) for k, v in sorted(solver.qmc.items()))
            self.seq = solver

        def solve(self, it):
            self.seq.set_problem(p.hamiltonian, p.beta, p.nbands, p.quantum_numbers, p.u_matrix, ftau5, -p.muimp, screening)
            return self.seq.solve(it)
    plugin(therm, 1).solve(0)                                  # thermalisation, untimed: one solve; dest_ctqmc parks the population

    # ---- `value`: chain state resident in HBM, K accumulation blocks timed on the device
    # (the parameter string and arrays are the plugin's own: the parked population is revived only for a bit-identical local problem.
    #  Until late in round 2 this leg passed problems.pstring(), whose key set differs, silently got a FRESH, empty population and
    #  timed it 300 moves after its creation -- C5 at <k> = 61 and growing instead of 74.  Now checked: see `revived` below.)
    def revive():
        seq = plugin(0, args.steps)
        seq.seq.set_problem(p.hamiltonian, p.beta, p.nbands, p.quantum_numbers, p.u_matrix, ftau5, -p.muimp, screening)
        eng.init_solver(seq.seq.umatrix, seq.seq.ftau, seq.seq.muimp, seq.seq.screening, 1)
        k = eng.scalar("mean_order")
        if therm > 0 and not k > 0.0:
            raise SystemExit("bench.py: the thermalised population was not revived (empty chains after init_solver)")
        return k
    mean_k_start = revive()
    eng.set_taudiffmax(p.qmc["TaudiffMax"])
    eng.init_counters()
    ms = (C.c_double * 3)()

    def blocks(n):
        eng._ck(L.w2b_bench_blocks(C.c_longlong(n), C.c_longlong(ncorr), C.c_int(1), ms))
        return ms[0], ms[1], ms[2]

    def allreduce_accum():
        if world > 1:
            eng.allreduce_accumulators()

    blocks(args.warmup)                                        # W untimed warm-up steps
    allreduce_accum()
    torch.cuda.synchronize()
    blocked = eng.scalar("blocked") > 0 or os.environ.get("W2B_SCHEDULER") == "2"
    bytes0, launches0 = eng.scalar("cnt_alg_bytes"), eng.scalar("gpu_launches")
    flops0 = eng.scalar("trace_flops") if blocked else 0.0
    rounds0 = eng.scalar("rounds")
    if blocked:
        eng.set_option("time_kernels", 1)                      # CUDA events around every round kernel, on the launching stream
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_all, t_sweep, t_meas = blocks(args.steps)                # K timed steps (CUDA events inside, on the launching stream)
    allreduce_accum()                                          # N > 1: the single all-reduce that ends an accumulation
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = max(e0.elapsed_time(e1), t_all)
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    bytes1, launches1 = eng.scalar("cnt_alg_bytes"), eng.scalar("gpu_launches")
    flops1 = eng.scalar("trace_flops") if blocked else 0.0
    rounds1 = eng.scalar("rounds")
    kms = {k: eng.scalar("ms_" + k) for k in ("propose", "trace", "decide")} if blocked else {}
    if blocked:
        eng.set_option("time_kernels", 0)
    mean_k = eng.scalar("mean_order")
    kcap = eng.scalar("kcap")
    moves = float(nchains) * ncorr * args.steps * world
    value = moves / (elapsed_ms * 1e-3)
    err = int(eng.scalar("error_flags"))

    # ---- wall time to a fixed G(tau) error (second half of BASELINE.json's metric): sample K fresh accumulation
    # blocks, SEM of G(beta/2) over the chain-group bins, T * (eps / eps*)^2 with eps* = 1e-3 |G(beta/2)|
    tte = None
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng._ck(L.w2b_sample(C.c_longlong(args.steps), C.c_int(1)))
        allreduce_accum()
        torch.cuda.synchronize()
        t_sample = time.perf_counter() - t0
        eng._ck(L.w2b_finalize())
        gb = eng.bins("gtau")                                   # [bins, band, spin, tau]
        mid = gb.shape[-1] // 2
        gm = np.abs(gb.mean(0)[..., mid]).mean()
        sem = (gb.std(0, ddof=1)[..., mid] / np.sqrt(gb.shape[0])).mean()
        gm_all = np.abs(gb.mean(0)).mean()                      # the same over all tau bins (G(beta/2) is exponentially small for an insulator)
        sem_all = (gb.std(0, ddof=1) / np.sqrt(gb.shape[0])).mean()
        if gm > 0 and sem > 0:
            tte = {"seconds_to_1e-3_relative": t_sample * (sem / (1e-3 * gm)) ** 2, "sample_seconds": t_sample, "sem_g_beta_half": float(sem),
                   "g_beta_half": float(gm), "seconds_to_1e-3_relative_tau_average": t_sample * (sem_all / (1e-3 * gm_all)) ** 2,
                   "sem_g_tau_average": float(sem_all), "g_tau_average": float(gm_all), "bins": int(gb.shape[0]), "moves": moves}
    except Exception as exc:                                    # diagnostic only: never lose the throughput line
        tte = {"error": str(exc)[:200]}
    eng.dest_ctqmc()                                            # parks the population again

    # ---- e2e: the K steps as ONE call of the reference plugin's solve (what a user of the reference runs): init_paras +
    # init_solver(host u, F(tau), mu) + set_taudiffmax + ctqmc_warmup(0) + ctqmc_measure(1,1) with Nmeas = K accumulation blocks
    # (N > 1: + the library's all-reduce) + every result array of get_resultset read back to host memory + dest_ctqmc /
    # dest_paras -- everything inside the timed region.  The population is the one thermalised above (reuse_population: revived
    # and rebuilt from its operator lists with the uploaded F(tau)).  `per_solve_of_one_step`: the same with Nmeas = 1, i.e. the
    # fixed cost of a solve (upload, rebuild, normalisation, result copies) charged to a single accumulation block.
    def timed_solves(nmeas, nsolves, it0):
        seq = plugin(0, nmeas)
        for it in range(2):
            seq.solve(it0 + it)                                # warm-up solves (untimed)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        hb0, db0 = eng.scalar("h2d_bytes"), eng.scalar("d2h_bytes")
        t0 = time.perf_counter()
        e0.record()
        res = None
        for it in range(nsolves):
            res = seq.solve(it0 + 2 + it)
        e1.record()
        torch.cuda.synchronize()
        ms_ = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        hb1, db1 = eng.scalar("h2d_bytes"), eng.scalar("d2h_bytes")
        if world > 1:
            t = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, hb1 - hb0, db1 - db0, res
    e2e_ms, hb, db, res = timed_solves(args.steps, 1, 100)
    hb0 = db0 = 0.0; hb1, db1 = hb, db
    e2e_val = moves / (e2e_ms * 1e-3)
    res_bytes = int(sum(np.asarray(v).nbytes for v in res.values()))
    one_ms, _, _, _ = timed_solves(1, 5, 200)
    one_val = float(nchains) * ncorr * 5 * world / (one_ms * 1e-3)
    # stationarity check: mean expansion order after every leg (the population was revived once more) next to the one before `value`
    mean_k_end = revive()
    eng.set_option("reuse_population", 0)
    eng.dest_ctqmc(); eng.dest_paras()                         # drop the population (frees the device memory for the next config)

    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    alg_bytes = bytes1 - bytes0
    achieved = alg_bytes / (t_sweep * 1e-3) / 1e9 if t_sweep > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(cfg)
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(cfg, args.cpu_seconds)
    state_gb = nchains * (2 * p.nbands + 1) * kcap ** 2 * 8 / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "chains_per_gpu": nchains, "ncorr": ncorr, "moves_per_step": nchains * ncorr * world,
                   "mean_order_per_flavour": mean_k, "mean_order_before_timed_steps": mean_k_start, "mean_order_after_all_legs": mean_k_end, "thermalisation_moves": therm, "parallelism": "chains sharded over %d GPU(s), no data-path collective; one ncclAllReduce of the accumulators per accumulation (inside the library)" % world,
                   "l2": "inverse-matrix state %.1f GB per GPU > 126 MB L2 (inputs larger than L2, no flush needed)" % state_gb,
                   "error_flags": err},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int((hb1 - hb0) / args.steps), "d2h_bytes_per_step": int((db1 - db0) / args.steps),
                "ms_per_step": e2e_ms / args.steps, "result_bytes_per_solve": res_bytes,
                "path": "one call of the reference plugin's solve through the C ABI with Nmeas = steps: init_paras, init_solver(host arrays), set_taudiffmax, ctqmc_warmup(0), ctqmc_measure(1,1), all get_resultset arrays, dest_ctqmc; bytes are per step (totals / steps)",
                "per_solve_of_one_step": {"value": one_val, "ms_per_solve": one_ms / 5}},
        "gpu_launches": int(launches1 - launches0),
        "kernel_ms": {"k_sweep": t_sweep / args.steps, "k_measure": t_meas / args.steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_sweep", "algorithmic_bytes_per_launch": alg_bytes / args.steps, "launch_ms": t_sweep / args.steps,
                     "peak_source": peak_src},
        "clocks": clocks,
    }
    if blocked:
        # blocked traces: the sweep is rounds of k_propose / k_trace / k_decide (+ k_rare on a second stream); the dominant
        # kernel is k_trace, bound by the FP64 tensor pipe (DMMA): algorithmic flops = 2 d_out d_in columns per operator
        # product, counted in the kernel, over the summed CUDA-event durations of its launches in the timed region
        nround = max(1.0, rounds1 - rounds0)
        fpk, fsrc = fp64_peak()
        tfl = (flops1 - flops0) / (kms["trace"] * 1e-3) / 1e12 if kms["trace"] > 0 else 0.0
        line["kernel_ms"] = {"sweep": t_sweep / args.steps, "k_propose": kms["propose"] / args.steps, "k_trace": kms["trace"] / args.steps,
                             "k_decide": kms["decide"] / args.steps, "k_measure": t_meas / args.steps, "rounds_per_step": nround / args.steps}
        line["roofline"] = {"bound": "tensor", "achieved": tfl, "peak": fpk, "unit": "TFLOP/s", "frac": tfl / fpk, "traffic": traffic,
                            "kernel": "k_trace", "dtype": "f64 (DMMA.8x8x4)", "algorithmic_flops_per_launch": (flops1 - flops0) / nround,
                            "launch_ms": kms["trace"] / nround, "peak_source": fsrc,
                            "hbm_view": {"achieved_gbs": achieved, "peak_gbs": peak, "note": "determinant-path algorithmic bytes over the whole sweep"}}
    if tte:
        line["time_to_error"] = tte
    if cpu:
        line["cpu_baseline"] = cpu
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("W2B_BENCH_CONFIG", "C4"), choices=sorted(WORKLOADS) + ["all"],
                    help="workload (default C4 = the configuration BASELINE.json's metric is quoted on); `all`: one JSON line per config")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default per config)")
    ap.add_argument("--thermalise", type=int, default=-1, help="untimed thermalisation moves per chain")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: split the one-GPU chain population over the ranks (default: weak, fixed chains per GPU)")
    ap.add_argument("--slice-seconds", type=float, default=0.0, help="reference arm: wall time of one step (default: sized for a one-minute run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    for cfg in (sorted(WORKLOADS) if args.config == "all" else [args.config]):
        if args.impl == "reference":
            run_reference(args, cfg)
        else:
            line = run_engine(args, cfg)
            if line is not None:
                print(json.dumps(line), flush=True)
    if args.impl != "reference" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
