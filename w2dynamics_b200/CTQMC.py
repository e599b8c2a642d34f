"""Drop-in stand-in for the reference's f2py module ``w2dyn.auxiliaries.CTQMC``.

The reference's solver plugin does (``w2dyn/dmft/cthyb_solver.py:21``)::

    from w2dyn.auxiliaries.CTQMC import mctqmc as ctqmc

Redirecting that single import to ::

    from w2dynamics_b200.CTQMC import mctqmc as ctqmc

routes the Z-space CT-HYB hot path (everything between ``ctqmc.init_solver`` and
``ctqmc.dest_ctqmc``, ``cthyb_solver.py:316-400``) to the B200 engine through the C ABI in
``include/w2b_cthyb.h``.  Same subroutine names, argument meaning, array names and shapes
(Fortran order) as the f2py module; where the reference ``stop``s, this raises RuntimeError.

One process drives one GPU and a population of ``nchains`` Markov chains (the reference:
one chain per MPI rank).  With one process per GPU call ``mctqmc.enable_distributed()``
once (under ``torch.distributed``), or ``mctqmc.comm_init(unique_id, rank, nranks)`` with
an id obtained from ``mctqmc.comm_unique_id()`` on rank 0 and broadcast by any means
(mpi4py ``bcast``): chains are sharded over ranks and the raw accumulators are combined
inside the library with a single ``ncclAllReduce`` per ``ctqmc_measure``.

There is no CPU fallback: the CUDA library must be present and a GPU visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(_HERE, "libw2b_cthyb.so")


class MCTQMC:
    """Object with the attribute/method surface of ``CTQMC.mctqmc``."""

    _ARRAYS = {
        # name: shape builder (d = dims dict) -- extents of the reference's module arrays
        "gtau": lambda d: (d["nb"], 2, d["ntau"]),
        "gtau_full": lambda d: (d["nb"], 2, d["nb"], 2, d["ntau"]),
        "gtau_blocks": lambda d: (d["nb"], 2, d["ntau"], -1),                     # AccumGtau = 2 only; last axis = binning levels
        "gleg": lambda d: (d["nb"], 2, d["nleg"]),
        "gleg_full": lambda d: (d["nb"], 2, d["nb"], 2, d["nleg"]),
        "occ": lambda d: (d["nb"], 2, d["nb"], 2),
        "histo": lambda d: (d["nb"], 2, d["maxhisto"] + 1),
        "rhisto": lambda d: (100,), "lhisto": lambda d: (100,),
        "statessuperstates": lambda d: (d["nstates"],),
        "eigenstatesenergies": lambda d: (d["nstates"],),
        "occbasismapping": lambda d: (d["nstates"], d["nb"], 2),
        "outersuperstatehisto": lambda d: (d["nsst"],),
        "outerstatehisto": lambda d: (d["nstates"],),
        "signhistosuperstates": lambda d: (d["nsst"],),
        "signhistostates": lambda d: (d["nstates"],),
        "tracecontribsuperstates": lambda d: (d["nsst"],),
        "tracecontribstates": lambda d: (d["nstates"],),
        "accadd": lambda d: (16,), "accrem": lambda d: (16,),
        "accpair": lambda d: (d["nb"], 2, d["ntau"]),
        "densitymatrix": lambda d: (d["nstates"], d["nstates"]),
        "rho1": lambda d: (2 * d["nb"], 2 * d["nb"]),
        "rho2": lambda d: (2 * d["nb"],) * 4,
    }
    _SCALARS = ("mean_sign", "accglob", "accshift", "accadd4", "accrem4", "accflavc",
                "time_warmup", "time_sim", "time_sim_steps", "aborted")
    _INT_ARRAYS = ("statessuperstates", "occbasismapping", "accpair")

    def __init__(self, libpath=CUDA_LIB):
        self._libpath = libpath
        self._lib = None
        self._dims = None
        self._dist = None
        self._cache = {}
        # plain module variables the reference lets Python write (cthyb_solver.py:114, 323, 241)
        self.fancy_prog = False
        self.simid = 0
        self.fourpnt = 0
        self._nchains = int(os.environ.get("W2B_NCHAINS", "4096"))
        self._nbins = int(os.environ.get("W2B_NBINS", "32"))
        self._options = {}

    # ------------------------------------------------------------------ plumbing
    def _load(self):
        if self._lib is None:
            if not os.path.exists(self._libpath):
                raise RuntimeError("B200 CT-HYB engine library missing (%s): build it with "
                                   "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback"
                                   % self._libpath)
            L = C.CDLL(self._libpath)
            L.w2b_last_error.restype = C.c_char_p
            self._lib = L
        return self._lib

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("w2b_cthyb: " + self._load().w2b_last_error().decode())

    def set_option(self, key, value):
        """Engine knobs without a reference counterpart (nchains, nbins, kcap, ops_cap, device, chain_base)."""
        self._options[key] = value
        if self._lib is not None and key in ("time_kernels", "reuse_population"):       # run-time switches take effect immediately
            self._ck(self._lib.w2b_set_option(key.encode(), C.c_double(float(value))))
        if key == "nchains":
            self._nchains = int(value)
        if key == "nbins":
            self._nbins = int(value)

    def comm_unique_id(self):
        """Rank 0: the NCCL unique id (bytes) every rank must pass to comm_init."""
        n = C.c_size_t()
        buf = C.create_string_buffer(256)
        self._ck(self._load().w2b_comm_unique_id(buf, C.c_size_t(256), C.byref(n)))
        return buf.raw[:n.value]

    def comm_init(self, unique_id, rank, nranks):
        """Collective: join the library's communicator (before init_solver).  Chains shard over ranks by global chain id and every
        ctqmc_measure ends with one all-reduce of the raw accumulators inside the library."""
        if "device" in self._options:
            self._ck(self._load().w2b_set_option(b"device", C.c_double(float(self._options["device"]))))
        self._ck(self._load().w2b_comm_init(unique_id, C.c_size_t(len(unique_id)), C.c_int(int(rank)), C.c_int(int(nranks))))
        self._dist = (int(rank), int(nranks))
        self.simid = int(rank)

    def enable_distributed(self, group=None):
        """torch.distributed convenience: broadcast rank 0's unique id over the process group, then comm_init."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(box[0], rank, size)

    # ------------------------------------------------------------------ the mctqmc subroutines
    def init_rnd(self, nseed):
        self._ck(self._load().w2b_init_rnd(C.c_int(int(nseed))))

    def init_paras(self, parastr):
        b = parastr.encode() if isinstance(parastr, str) else bytes(parastr)
        self._pstring = b.decode()
        self._ck(self._load().w2b_init_paras(b, C.c_size_t(len(b))))

    def init_solver(self, u_matrix_in, ftau_full, muimp_full, screening_function=None, iter_no=1):
        L = self._load()
        u = np.asfortranarray(np.asarray(u_matrix_in, dtype=np.float64))
        ft = np.asfortranarray(np.asarray(ftau_full, dtype=np.float64))
        mu = np.asfortranarray(np.asarray(muimp_full, dtype=np.float64))
        nb, nftau = ft.shape[0], ft.shape[4]
        if u.shape != (2 * nb,) * 4 or mu.shape != (nb, 2, nb, 2) or ft.shape[:4] != (nb, 2, nb, 2):
            raise ValueError("init_solver: inconsistent array extents")
        # the library is a process-wide singleton like the reference's module: every engine states ALL its options, so that
        # nothing set through another MCTQMC object lingers
        opts = {"kcap": 0, "ops_cap": 0, "cta_warps": 0, "batch_chains": 0, "scheduler": 0, "chain_base": 0, "propose_run": 4,
                "prefetch": 1, "debug_same_key": 0, "time_kernels": 0, "reuse_population": 0, "lu_groups": 0}
        opts.update(self._options)
        opts.setdefault("nchains", self._nchains)
        opts.setdefault("nbins", self._nbins)
        for k, v in opts.items():
            self._ck(L.w2b_set_option(k.encode(), C.c_double(float(v))))
        dp = C.POINTER(C.c_double)
        self._ck(L.w2b_init_solver(u.ctypes.data_as(dp), ft.ctypes.data_as(dp), mu.ctypes.data_as(dp), None,
                                   C.c_int(int(iter_no)), C.c_int(nftau), C.c_int(nb)))
        self._dims = None
        self._cache = {}
        self._nb = nb

    def set_taudiffmax(self, taudiffmax):
        self._ck(self._load().w2b_set_taudiffmax(C.c_double(float(taudiffmax))))

    def init_counters(self):
        self._ck(self._load().w2b_init_counters())

    def ctqmc_warmup(self, nwarmups):
        self._ck(self._load().w2b_ctqmc_warmup(C.c_longlong(int(nwarmups))))

    def ctqmc_calibrate(self, list_, phase1pairnum, phase2taugrid, phase2pairnum, nprogress):
        self._ck(self._load().w2b_ctqmc_calibrate(C.c_int(int(bool(list_))), C.c_int(phase1pairnum), C.c_int(phase2taugrid),
                                                  C.c_int(phase2pairnum), C.c_int(nprogress)))

    def ctqmc_measure(self, isector=1, icomponent=1):
        L = self._load()
        self._cache = {}
        if isector != 1:
            raise RuntimeError("w2b_cthyb: only the partition-function sector (isector = 1) is implemented")
        self._ck(L.w2b_ctqmc_measure(C.c_int(isector), C.c_int(icomponent)))      # ends with the library's own all-reduce when a communicator exists

    def allreduce_accumulators(self):
        """The all-reduce by itself (split form w2b_sample / w2b_allreduce_accum / w2b_finalize; bench.py)."""
        self._ck(self._load().w2b_allreduce_accum())

    def dest_ctqmc(self):
        self._ck(self._load().w2b_dest_ctqmc())

    def dest_paras(self):
        self._ck(self._load().w2b_dest_paras())

    # ------------------------------------------------------------------ Monte-Carlo configurations
    def get_mc_config_scalars(self, chain=0):
        n, o, s = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._load().w2b_get_mc_config_scalars(C.c_int(chain), C.byref(n), C.byref(o), C.byref(s)))
        return n.value, o.value, s.value

    def get_mc_config(self, n, chain=0):
        taus = np.zeros(n)
        orbs, spins, cas, hyb = (np.zeros(n, dtype=np.int32) for _ in range(4))
        o, s = C.c_int(), C.c_int()
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        self._ck(self._load().w2b_get_mc_config(C.c_int(chain), C.c_int(n), taus.ctypes.data_as(dp), orbs.ctypes.data_as(ip),
                                                spins.ctypes.data_as(ip), cas.ctypes.data_as(ip), hyb.ctypes.data_as(ip),
                                                C.byref(o), C.byref(s)))
        return taus, orbs, spins, cas, hyb, o.value, s.value

    def set_mc_config(self, n, taus, orbs, spins, cas, hashybs, outer_sst, outer_state, chain=0):
        taus = np.ascontiguousarray(taus, dtype=np.float64)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (orbs, spins, cas, hashybs)]
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        self._ck(self._load().w2b_set_mc_config(C.c_int(chain), C.c_int(int(n)), taus.ctypes.data_as(dp),
                                                *[a.ctypes.data_as(ip) for a in arrs], C.c_int(int(outer_sst)), C.c_int(int(outer_state))))

    def set_empty_mc_config(self, outer_sst, outer_state, chain=0):
        self._ck(self._load().w2b_set_empty_mc_config(C.c_int(chain), C.c_int(int(outer_sst)), C.c_int(int(outer_state))))

    # ------------------------------------------------------------------ results (module variables of the reference)
    def _param(self, key):
        for line in self._pstring.split("\n"):
            if "=" in line:
                k, v = line.split("=", 1)
                if k.strip().startswith(key):
                    return v.split("#")[0].strip()
        raise KeyError(key)

    def scalar(self, name):
        out = C.c_double()
        self._ck(self._load().w2b_get_scalar(name.encode(), C.byref(out)))
        return out.value

    def _get_dims(self):
        if self._dims is None:
            self._dims = dict(nb=self._nb, ntau=int(float(self._param("Ntau"))), nleg=int(float(self._param("NLegMax"))),
                              maxhisto=int(float(self._param("MaxHisto"))), nstates=int(self.scalar("nstates")),
                              nsst=int(self.scalar("nsst")), nbins=int(self.scalar("nbins")))
        return self._dims

    def array(self, name, shape=None):
        n = C.c_size_t()
        L = self._load()
        self._ck(L.w2b_get_array(name.encode(), None, C.c_size_t(0), C.byref(n)))
        buf = np.zeros(n.value)
        self._ck(L.w2b_get_array(name.encode(), buf.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(n.value), C.byref(n)))
        if shape is not None:
            buf = buf.reshape(shape, order="F")
        return buf

    def bins(self, name):
        """Per-bin (chain-group) samples of gtau / gleg / occ / sign: array with a leading bin axis, so that
        ``DistributedSample(list(bins), comm)`` keeps its meaning (statistics.py:36-63)."""
        d = self._get_dims()
        if name == "sign":
            return self.array("sign_bins")
        if name == "giw":
            flat = self.array("giw_bins")
            nbins = d["nbins"]
            flat = (flat[0::2] + 1j * flat[1::2]).reshape(nbins, -1)
            return np.stack([flat[i].reshape((self._nb, 2, -1), order="F") for i in range(nbins)])
        shp = self._ARRAYS[name](d)
        flat = self.array(name + "_bins")
        nb = flat.size // int(np.prod(shp))
        return np.stack([flat[i * int(np.prod(shp)):(i + 1) * int(np.prod(shp))].reshape(shp, order="F") for i in range(nb)])

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name == "accpairtau":
            return self.array("accpairtau")
        if name == "giw":
            # MeasGiw != 0: Giw(NBands, 2, -Niw:Niw-1), complex (CTQMC.F90:573, 8815-8817; read at cthyb_solver.py:990); the C ABI hands
            # complex numbers over as (re, im) pairs
            if "giw" not in self._cache:
                raw = self.array("giw")
                self._cache["giw"] = np.asfortranarray((raw[0::2] + 1j * raw[1::2]).reshape((self._nb, 2, -1), order="F"))
            return self._cache["giw"]
        if name == "accpair" and int(self.scalar("naccpair")) != self._get_dims()["ntau"]:
            # calibration phase 2 (cthyb_solver.py:914-921): AccPair(Nb, 2, phase2taugrid)
            return np.asfortranarray(np.rint(self.array("accpair", (self._nb, 2, int(self.scalar("naccpair"))))).astype(np.int64))
        if name in MCTQMC._ARRAYS:
            if name not in self._cache:
                a = self.array(name, MCTQMC._ARRAYS[name](self._get_dims()))
                if name in MCTQMC._INT_ARRAYS:
                    a = np.asfortranarray(np.rint(a).astype(np.int64))
                self._cache[name] = a
            return self._cache[name]
        if name in MCTQMC._SCALARS:
            v = self.scalar(name)       # f2py hands module scalars back as 0-d arrays (the caller does `.copy()`, solver.py:1028)
            return np.int32(v) if name == "aborted" else np.float64(v)
        raise AttributeError(name)

    # ------------------------------------------------------------------ parity / bench hooks
    def debug_steps(self, chain, nsteps):
        log = np.zeros((nsteps, 20))
        self._ck(self._load().w2b_debug_steps(C.c_int(chain), C.c_longlong(nsteps), log.ctypes.data_as(C.POINTER(C.c_double))))
        return log

    def debug_chain_scalars(self, chain=0):
        out = np.zeros(8)
        self._ck(self._load().w2b_debug_chain_scalars(C.c_int(chain), out.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(trace_log=out[0], trace_sign=out[1], det_log=out[2], det_sign=out[3], cfgsign=out[4],
                    n_ops=int(out[5]), outer_sst=int(out[6]), error=int(out[7]))

    def debug_minv(self, chain, orbital, spin):
        k = C.c_int()
        L = self._load()
        self._ck(L.w2b_debug_minv(C.c_int(chain), C.c_int(orbital), C.c_int(spin), None, C.c_size_t(0), C.byref(k)))
        out = np.zeros(k.value * k.value)
        self._ck(L.w2b_debug_minv(C.c_int(chain), C.c_int(orbital), C.c_int(spin), out.ctypes.data_as(C.POINTER(C.c_double)),
                                  C.c_size_t(out.size), C.byref(k)))
        return out.reshape((k.value, k.value), order="F")

    def block(self, nsweep, do_measure=True):
        self._ck(self._load().w2b_block(C.c_longlong(int(nsweep)), C.c_int(int(do_measure))))

    def sync(self):
        self._ck(self._load().w2b_sync())


mctqmc = MCTQMC()
