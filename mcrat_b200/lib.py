"""ctypes binding of libmcrat_b200.so (the C ABI declared in include/mcrat_b200.h).

This is plumbing for the Python test / benchmark harness; the product boundary is
the C ABI itself (the host program is C).  Nothing here
computes: a missing or unloadable library is an error, there is no fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .synth import PHOTON_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libmcrat_b200.so")
DROPIN_PATH = os.path.join(CSRC, "libmcrat_b200_dropin.so")

HYDRO_FIELDS = ["r0", "r1", "r2", "r0_size", "r1_size", "r2_size", "r", "theta", "v0", "v1", "v2",
                "dens", "dens_lab", "pres", "temp", "gamma", "B0", "B1", "B2"]

ABI_VERSION = 2
RNG_PHILOX, RNG_REPLAY = 0, 1
LOOP_MODES = {"auto": 0, "streamed": 1, "persistent": 2, "streamed_global": 3, "persistent_stream": 4}

ERRORS = {-1: "ERR_CUDA", -2: "ERR_ARG", -3: "ERR_STATE", -4: "ERR_REPLAY", -5: "ERR_TABLE"}


class McratB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "ERR"), code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("dimensions", C.c_int), ("geometry", C.c_int), ("stokes_switch", C.c_int),
                ("tau_calculation", C.c_int), ("cyclosynch_switch", C.c_int), ("b_field_calc", C.c_int),
                ("epsilon_b", C.c_double), ("device", C.c_int), ("rng_mode", C.c_int), ("seed", C.c_uint64),
                ("shard", C.c_uint32), ("profile", C.c_int), ("stream", C.c_void_p), ("scan_index", C.c_int)]


class FrameStats(C.Structure):
    _fields_ = [("iterations", C.c_longlong), ("scatterings", C.c_longlong), ("relocations", C.c_longlong),
                ("photon_slots", C.c_longlong), ("cell_evals", C.c_longlong), ("box_evals", C.c_longlong),
                ("time_now", C.c_double),
                ("last_time_step", C.c_double), ("last_scattered_index", C.c_int), ("not_found", C.c_int),
                ("cs_host_pending", C.c_int), ("error", C.c_int), ("cs_emitted", C.c_int),
                ("scatt_cyclosynch_num_ph", C.c_int), ("cs_comptonized_weight", C.c_double),
                ("ref_equiv_evals", C.c_longlong)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class KernelTimes(C.Structure):
    _fields_ = [("scan_ms", C.c_double), ("scan_launches", C.c_longlong), ("pass_ms", C.c_double),
                ("pass_launches", C.c_longlong), ("event_ms", C.c_double), ("event_launches", C.c_longlong),
                ("other_ms", C.c_double), ("other_launches", C.c_longlong)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = [
    "mcrat_b200_abi_version", "mcrat_b200_device_count", "mcrat_b200_create", "mcrat_b200_destroy",
    "mcrat_b200_last_error", "mcrat_b200_synchronize", "mcrat_b200_set_hydro", "mcrat_b200_set_thermal_table",
    "mcrat_b200_build_thermal_table", "mcrat_b200_set_photons", "mcrat_b200_get_photons", "mcrat_b200_get_photon", "mcrat_b200_list_capacity",
    "mcrat_b200_set_num_shards", "mcrat_b200_num_shards", "mcrat_b200_get_shard_stats",
    "mcrat_b200_set_replay_uniforms", "mcrat_b200_replay_consumed", "mcrat_b200_find_containing_hydro_cell",
    "mcrat_b200_calc_mean_free_path", "mcrat_b200_photon_event", "mcrat_b200_update_photon_position",
    "mcrat_b200_ph_abs_cyclosynch", "mcrat_b200_calc_cyclosynch_r_limits", "mcrat_b200_set_cs_limits", "mcrat_b200_ph_min_max", "mcrat_b200_ph_scatt_stats",
    "mcrat_b200_average_photon_energy", "mcrat_b200_run_frame", "mcrat_b200_set_loop_mode",
    "mcrat_b200_rebin_cyclosynch_comp_photons", "mcrat_b200_set_cs_rebin_params", "mcrat_b200_get_kernel_times",
    "mcrat_b200_launch_count", "mcrat_b200_rescan_all", "mcrat_b200_measure_fp64_peak", "mcrat_b200_measure_hbm_peak",
    "mcrat_b200_selftest_div_by_c", "mcrat_b200_set_recheck_skip", "mcrat_b200_set_profile",
    "mcrat_b200_photon_emit_cyclosynch", "mcrat_b200_photon_emit_cyclosynch_single",
    "mcrat_b200_get_not_found", "mcrat_b200_get_sorted_indexes",
    "mcrat_b200_comm_unique_id", "mcrat_b200_comm_nccl_version", "mcrat_b200_comm_create", "mcrat_b200_comm_destroy",
    "mcrat_b200_comm_rank", "mcrat_b200_comm_size", "mcrat_b200_comm_collectives", "mcrat_b200_comm_bcast_thermal_table",
    "mcrat_b200_comm_build_thermal_table", "mcrat_b200_comm_reduce_frame_stats", "mcrat_b200_comm_photon_counts",
    "mcrat_b200_comm_gather_photons",
]


def build(force=False):
    """Compile the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".c")) or f == "Makefile"]
    srcs += [os.path.join(HERE, "..", "include", f) for f in ("mcrat_b200.h", "mcrat_b200_dropin.h")]
    srcs = [s for s in srcs if os.path.exists(s)]
    stale = force or not os.path.exists(LIB_PATH) or not os.path.exists(DROPIN_PATH) or \
        any(os.path.getmtime(s) > min(os.path.getmtime(LIB_PATH), os.path.getmtime(DROPIN_PATH)) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", CSRC, "-s", "all"])
    return LIB_PATH


_lib = None


def _pin_nccl():
    """The library binds NCCL at run time by soname (libnccl.so.2).  In a Python process that also imports torch the copy
    torch was built against (the nvidia-nccl wheel next to it) must be the one that gets loaded, whichever of the two
    comes first: the dynamic linker keeps one object per soname, and an older system NCCL loaded first would make a
    later `import torch` fail on missing symbols.  A C host (mcrat.c under MPI) has no such concern and uses the system's."""
    if os.environ.get("MCRAT_B200_NCCL"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MCRAT_B200_NCCL"] = cand
                return
    except Exception:
        pass


def load():
    """dlopen the library; raises if it is not built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        path = os.environ.get("MCRAT_B200_LIB", LIB_PATH)  # A/B experiments with alternative builds
        if not os.path.exists(path):
            raise FileNotFoundError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no CPU fallback)" % LIB_PATH)
        _pin_nccl()
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.mcrat_b200_last_error.restype = C.c_char_p
        L.mcrat_b200_last_error.argtypes = [C.c_void_p]
        L.mcrat_b200_replay_consumed.restype = C.c_longlong
        L.mcrat_b200_replay_consumed.argtypes = [C.c_void_p]
        L.mcrat_b200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.mcrat_b200_destroy.argtypes = [C.c_void_p]
        L.mcrat_b200_launch_count.restype = C.c_longlong
        L.mcrat_b200_launch_count.argtypes = [C.c_void_p]
        L.mcrat_b200_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.mcrat_b200_comm_destroy.argtypes = [C.c_void_p]
        L.mcrat_b200_comm_collectives.restype = C.c_longlong
        L.mcrat_b200_comm_collectives.argtypes = [C.c_void_p]
        for f in ("rank", "size"):
            getattr(L, "mcrat_b200_comm_" + f).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class HotPath:
    """Host-side mirror of the reference's function surface for one shard (one GPU).

    Method names follow the reference (findContainingHydroCell, calcMeanFreePath,
    photonEvent, updatePhotonPosition, phAbsCyclosynch, ...; Src/mclib.h:8-23,
    Src/mc_cyclosynch.h:84-92); each forwards to the C ABI.
    """

    def __init__(self, cfg, device=0, rng_mode=RNG_PHILOX, seed=0, shard=0, profile=False, stream=None,
                 num_shards=1, scan_index=False, loop_mode=None):
        self.L = load()
        c = Config(ABI_VERSION, cfg["dimensions"], cfg["geometry"], cfg["stokes"], cfg["tau_calculation"],
                   cfg["cyclosynch"], cfg["b_field_calc"], cfg["epsilon_b"], device, rng_mode, seed, shard,
                   1 if profile else 0, stream, 1 if scan_index else 0)
        self.ctx = C.c_void_p()
        rc = self.L.mcrat_b200_create(C.byref(c), C.byref(self.ctx))
        if rc != 0:
            raise McratB200Error(rc, self.L.mcrat_b200_last_error(None).decode())
        self.cfg = dict(cfg)
        self._keep = []
        if num_shards != 1:
            self.set_num_shards(num_shards)
        if loop_mode is not None:
            self.set_loop_mode(loop_mode)

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.L.mcrat_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise McratB200Error(rc, self.L.mcrat_b200_last_error(self.ctx).decode())

    # ---- inputs -----------------------------------------------------------------------
    def set_hydro(self, hydro):
        n = int(hydro["num_elements"])
        ptrs = (C.POINTER(C.c_double) * 19)()
        keep = []
        for i, f in enumerate(HYDRO_FIELDS):
            if f in hydro:
                a = np.ascontiguousarray(hydro[f], dtype=np.float64)
                if a.size != n:
                    raise ValueError("hydro field %s has %d elements, expected %d" % (f, a.size, n))
                keep.append(a)
                ptrs[i] = _dp(a)
            else:
                ptrs[i] = None
        dom = np.array(list(hydro.get("r0_domain", (0, 0))) + list(hydro.get("r1_domain", (0, 0))) +
                       list(hydro.get("r2_domain", (0, 0))), dtype=np.float64)
        self._ck(self.L.mcrat_b200_set_hydro(self.ctx, C.c_int(n), ptrs, _dp(dom), C.c_double(hydro.get("fps", 5.0)),
                                             C.c_int(hydro.get("scatt_frame_number", 0)),
                                             C.c_int(hydro.get("inj_frame_number", 0))))

    def set_thermal_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (221, 81)
        self._ck(self.L.mcrat_b200_set_thermal_table(self.ctx, _dp(t)))

    def build_thermal_table(self, calls=500000, seed=1):
        """Hot cross-section table built on the device (returns the 221 x 81 table and the kernel time in ms)."""
        t = np.zeros((221, 81), dtype=np.float64)
        ms = C.c_float(0)
        self._ck(self.L.mcrat_b200_build_thermal_table(self.ctx, C.c_longlong(calls), C.c_uint64(seed), _dp(t), C.byref(ms)))
        return t, ms.value

    def set_photons(self, photons):
        ph = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
        self._ck(self.L.mcrat_b200_set_photons(self.ctx, ph.ctypes.data_as(C.c_void_p), C.c_int(ph.size)))

    def set_photons_ptr(self, ptr, n):
        """Upload from a raw host pointer (e.g. pinned memory owned by the caller)."""
        self._ck(self.L.mcrat_b200_set_photons(self.ctx, C.c_void_p(ptr), C.c_int(n)))

    def get_photons(self, out=None):
        n = self.L.mcrat_b200_list_capacity(self.ctx)
        if out is None:
            out = np.zeros(n, dtype=PHOTON_DTYPE)
        self._ck(self.L.mcrat_b200_get_photons(self.ctx, out.ctypes.data_as(C.c_void_p), C.c_int(n)))
        return out

    def get_photons_ptr(self, ptr, n):
        self._ck(self.L.mcrat_b200_get_photons(self.ctx, C.c_void_p(ptr), C.c_int(n)))

    def get_photon(self, index):
        out = np.zeros(1, dtype=PHOTON_DTYPE)
        self._ck(self.L.mcrat_b200_get_photon(self.ctx, C.c_int(index), out.ctypes.data_as(C.c_void_p)))
        return out[0]

    def set_num_shards(self, n):
        """Sub-shards (independent reference 'ranks') the next set_photons() splits the list into."""
        self._ck(self.L.mcrat_b200_set_num_shards(self.ctx, C.c_int(n)))

    def num_shards(self):
        return int(self.L.mcrat_b200_num_shards(self.ctx))

    def set_loop_mode(self, mode):
        """'auto' | 'streamed' (four launches per iteration) | 'persistent' (one cooperative launch per frame)."""
        self._ck(self.L.mcrat_b200_set_loop_mode(self.ctx, C.c_int(LOOP_MODES[mode] if isinstance(mode, str) else mode)))

    def shard_stats(self, shard):
        st, first, count = FrameStats(), C.c_int(0), C.c_int(0)
        self._ck(self.L.mcrat_b200_get_shard_stats(self.ctx, C.c_int(shard), C.byref(st), C.byref(first), C.byref(count)))
        d = st.as_dict()
        d.update(first_slot=first.value, num_slots=count.value)
        return d

    def set_replay_uniforms(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._ck(self.L.mcrat_b200_set_replay_uniforms(self.ctx, _dp(u), C.c_size_t(u.size)))

    def replay_consumed(self):
        return int(self.L.mcrat_b200_replay_consumed(self.ctx))

    # ---- the reference's function surface -----------------------------------------------
    def findContainingHydroCell(self, find_nearest_block_switch):
        n = C.c_int(0)
        self._ck(self.L.mcrat_b200_find_containing_hydro_cell(self.ctx, C.c_int(find_nearest_block_switch), C.byref(n)))
        return n.value

    def calcMeanFreePath(self):
        i, t = C.c_int(0), C.c_double(0)
        self._ck(self.L.mcrat_b200_calc_mean_free_path(self.ctx, C.byref(i), C.byref(t)))
        return i.value, t.value

    def sortedIndexes(self, n=None):
        """photonList.sorted_indexes in full (slot indices by ascending time_to_scatter, ties in slot order)."""
        n = int(self.L.mcrat_b200_list_capacity(self.ctx)) if n is None else int(n)
        out = np.zeros(n, dtype=np.int32)
        self._ck(self.L.mcrat_b200_get_sorted_indexes(self.ctx, out.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(n)))
        return out

    def photonEvent(self, dt_max):
        ts, idx, sc, ab = C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.mcrat_b200_photon_event(self.ctx, C.c_double(dt_max), C.byref(ts), C.byref(idx), C.byref(sc),
                                                C.byref(ab)))
        return ts.value, idx.value, sc.value

    def updatePhotonPosition(self, t):
        self._ck(self.L.mcrat_b200_update_photon_position(self.ctx, C.c_double(t)))

    def phAbsCyclosynch(self):
        na, ns, w = C.c_int(0), C.c_int(0), C.c_double(0)
        self._ck(self.L.mcrat_b200_ph_abs_cyclosynch(self.ctx, C.byref(na), C.byref(ns), C.byref(w)))
        return w.value, na.value, ns.value

    def photonEmitCyclosynch(self, r_inj, ph_weight, maximum_photons, theta_min, theta_max):
        """All-cells mode (inject_single_switch = 0) -> (photons emitted, adjusted weight, cells in the shell)."""
        n, w, nc = C.c_int(0), C.c_double(0), C.c_int(0)
        self._ck(self.L.mcrat_b200_photon_emit_cyclosynch(self.ctx, C.c_double(r_inj), C.c_double(ph_weight), C.c_int(maximum_photons),
                                                          C.c_double(theta_min), C.c_double(theta_max), C.byref(n), C.byref(w),
                                                          C.byref(nc)))
        return n.value, w.value, nc.value

    def photonEmitCyclosynchSingle(self, scatt_ph_index):
        """inject_single_switch = 1 -> slot of the new pool photon."""
        slot = C.c_int(-1)
        self._ck(self.L.mcrat_b200_photon_emit_cyclosynch_single(self.ctx, C.c_int(scatt_ph_index), C.byref(slot)))
        return slot.value

    def rebinCyclosynchCompPhotons(self, max_photons):
        """-> (number of empty bins, num_cyclosynch_ph_emit, scatt_cyclosynch_num_ph), Src/mc_cyclosynch.c:600-710"""
        emit, scatt, nnull = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.mcrat_b200_rebin_cyclosynch_comp_photons(self.ctx, C.c_int(max_photons), C.byref(emit), C.byref(scatt),
                                                                 C.byref(nnull)))
        return nnull.value, emit.value, scatt.value

    def set_cs_rebin_params(self, e_perc=0.1, ang=0.5, ang_phi=10.0):
        self._ck(self.L.mcrat_b200_set_cs_rebin_params(self.ctx, C.c_double(e_perc), C.c_double(ang), C.c_double(ang_phi)))

    def set_cs_limits(self, max_photons, scatt_cyclosynch_num_ph=0):
        self._ck(self.L.mcrat_b200_set_cs_limits(self.ctx, C.c_int(max_photons), C.c_int(scatt_cyclosynch_num_ph)))

    def calcCyclosynchRLimits(self, frame_scatt, frame_inj, fps, r_inj, which):
        self.L.mcrat_b200_calc_cyclosynch_r_limits.restype = C.c_double
        return self.L.mcrat_b200_calc_cyclosynch_r_limits(C.c_int(frame_scatt), C.c_int(frame_inj), C.c_double(fps),
                                                          C.c_double(r_inj), which.encode())

    def phMinMax(self):
        v = [C.c_double(0) for _ in range(4)]
        self._ck(self.L.mcrat_b200_ph_min_max(self.ctx, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def phScattStats(self):
        mx, mn, avg, ravg = C.c_int(0), C.c_int(0), C.c_double(0), C.c_double(0)
        self._ck(self.L.mcrat_b200_ph_scatt_stats(self.ctx, C.byref(mx), C.byref(mn), C.byref(avg), C.byref(ravg)))
        return mx.value, mn.value, avg.value, ravg.value

    def averagePhotonEnergy(self):
        e = C.c_double(0)
        self._ck(self.L.mcrat_b200_average_photon_energy(self.ctx, C.byref(e)))
        return e.value

    # ---- device-resident frame loop ---------------------------------------------------------
    def run_frame(self, time_now, remaining_time, max_iters=-1, switch=1):
        st = FrameStats()
        self._ck(self.L.mcrat_b200_run_frame(self.ctx, C.c_double(time_now), C.c_double(remaining_time),
                                             C.c_longlong(max_iters), C.c_int(switch), C.byref(st)))
        return st.as_dict()

    def not_found(self, max_entries=32):
        """-> (slots, hydro coordinates [n, 3], total count) of the photons without a containing cell since the last call."""
        slots = np.zeros(max_entries, dtype=np.int32)
        h = np.zeros((max_entries, 3), dtype=np.float64)
        tot = C.c_int(0)
        n = self.L.mcrat_b200_get_not_found(self.ctx, C.c_int(max_entries), slots.ctypes.data_as(C.POINTER(C.c_int)), _dp(h), C.byref(tot))
        if n < 0:
            self._ck(n)
        return slots[:n], h[:n], tot.value

    # ---- measurement -----------------------------------------------------------------------------
    def synchronize(self):
        self._ck(self.L.mcrat_b200_synchronize(self.ctx))

    def launch_count(self):
        return int(self.L.mcrat_b200_launch_count(self.ctx))

    def kernel_times(self, reset=False):
        t = KernelTimes()
        self._ck(self.L.mcrat_b200_get_kernel_times(self.ctx, C.byref(t), C.c_int(1 if reset else 0)))
        return t.as_dict()

    def set_profile(self, on):
        self._ck(self.L.mcrat_b200_set_profile(self.ctx, C.c_int(1 if on else 0)))

    def rescan_all(self):
        ev, ms = C.c_longlong(0), C.c_float(0)
        self._ck(self.L.mcrat_b200_rescan_all(self.ctx, C.byref(ev), C.byref(ms)))
        return ev.value, ms.value

    def measure_fp64_peak(self):
        v = C.c_double(0)
        self._ck(self.L.mcrat_b200_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def set_recheck_skip(self, mode):
        self._ck(self.L.mcrat_b200_set_recheck_skip(self.ctx, int(mode)))

    def selftest_div_by_c(self, n, seed=1):
        v = C.c_longlong(-1)
        self._ck(self.L.mcrat_b200_selftest_div_by_c(self.ctx, C.c_longlong(n), C.c_uint(seed), C.byref(v)))
        return v.value

    def measure_hbm_peak(self):
        v = C.c_double(0)
        self._ck(self.L.mcrat_b200_measure_hbm_peak(self.ctx, C.byref(v)))
        return v.value


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it, the host distributes the bytes)."""
    L = load()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.mcrat_b200_comm_unique_id(buf, C.c_size_t(COMM_ID_BYTES))
    if rc != 0:
        raise McratB200Error(rc, L.mcrat_b200_last_error(None).decode())
    return buf.raw


class Comm:
    """The reference's MPI exchanges either side of the frame loop, over NCCL (include/mcrat_b200.h, mcrat_b200_comm_*):
    one communicator per HotPath context, one process per GPU.  `dist` (torch.distributed, any backend) is used only to
    hand the unique id from rank 0 to the others, the way mcrat.c would MPI_Bcast it."""

    def __init__(self, hot_path, nranks, rank, unique_id):
        self.hp = hot_path
        self.L = hot_path.L
        self.c = C.c_void_p()
        rc = self.L.mcrat_b200_comm_create(hot_path.ctx, int(nranks), int(rank), unique_id, C.c_size_t(len(unique_id)),
                                           C.byref(self.c))
        hot_path._ck(rc)

    @classmethod
    def from_torch_dist(cls, hot_path, dist):
        """unique id created on rank 0 and broadcast as an object over the existing process group."""
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
        box = [comm_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return cls(hot_path, world, rank, box[0])

    def close(self):
        if getattr(self, "c", None) is not None and self.c and self.hp.ctx:
            self.L.mcrat_b200_comm_destroy(self.c)
        self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rank(self):
        return int(self.L.mcrat_b200_comm_rank(self.c))

    @property
    def size(self):
        return int(self.L.mcrat_b200_comm_size(self.c))

    def collectives(self):
        return int(self.L.mcrat_b200_comm_collectives(self.c))

    def bcast_thermal_table(self, root=0):
        self.hp._ck(self.L.mcrat_b200_comm_bcast_thermal_table(self.c, C.c_int(root)))

    def build_thermal_table(self, calls=500000, seed=1):
        t = np.zeros((221, 81), dtype=np.float64)
        ms = C.c_float(0)
        self.hp._ck(self.L.mcrat_b200_comm_build_thermal_table(self.c, C.c_longlong(calls), C.c_uint64(seed), _dp(t), C.byref(ms)))
        return t, ms.value

    def reduce_frame_stats(self, stats):
        mine, tot = FrameStats(), FrameStats()
        for f, _ in FrameStats._fields_:
            setattr(mine, f, stats[f])
        self.hp._ck(self.L.mcrat_b200_comm_reduce_frame_stats(self.c, C.byref(mine), C.byref(tot)))
        return tot.as_dict()

    def photon_counts(self):
        n = self.size
        a, b, c = (np.zeros(n, dtype=np.int64) for _ in range(3))
        ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_longlong))
        self.hp._ck(self.L.mcrat_b200_comm_photon_counts(self.c, ip(a), ip(b), ip(c)))
        return {"list_capacity": a, "output_photons": b, "null_slots": c}

    def gather_photons(self, root=0, out=None, capacity=None):
        """-> (photons of all ranks in rank order, or None on non-receiving ranks; per-rank counts).

        Collective: every rank of the communicator calls it, with the same `root`.  A receiving rank passes a buffer
        (`out`) or a `capacity`, or neither: the C call itself finds out -- collectively -- whether every receiver has room;
        if not, nothing is transferred, every rank learns the total and comes back once more, the receivers with a buffer of
        that size (a too small `out` is replaced, the returned array is then not `out`).  The wrapper issues no collective
        of its own, so ranks may differ in what they pass."""
        n = self.size
        counts = np.zeros(n, dtype=np.int64)
        tot = C.c_longlong(0)
        recv = root == -1 or root == self.rank
        if recv and out is None and capacity is not None:
            out = np.zeros(int(capacity), dtype=PHOTON_DTYPE)
        for attempt in range(2):
            have = recv and out is not None
            ptr = out.ctypes.data_as(C.c_void_p) if have else None
            rc = self.L.mcrat_b200_comm_gather_photons(self.c, C.c_int(root), ptr, C.c_longlong(out.size if have else 0),
                                                       counts.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(tot))
            if rc == 0:
                return (out[:tot.value] if recv else None), counts
            if rc == -2 and attempt == 0 and tot.value > 0:
                # agreed by all ranks inside the call: some receiver lacks room.  Everybody calls again.
                if recv and (out is None or out.size < tot.value):
                    out = np.zeros(tot.value, dtype=PHOTON_DTYPE)
                continue
            self.hp._ck(rc)
        raise McratB200Error(-3, "comm_gather_photons: no agreement on the buffer sizes after two rounds")
