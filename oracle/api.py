"""ctypes bindings for the oracle (TEST INFRASTRUCTURE).

Two checkers live behind this module:

* :class:`Oracle`  -- ``oracle/liboracle.so``: the plain-C restatement of the
  reference's hot path with run-time configuration (oracle/mcrat_oracle.c).
* :class:`RefLib`  -- ``oracle/_ref/libmcrat_ref_<cfg>.so``: the reference's own
  unmodified sources for one compile-time configuration (oracle/build_ref.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module; the product (``mcrat_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# struct photon, Src/mcrat.h:142-171 (176 bytes with natural alignment)
PHOTON_DTYPE = np.dtype([
    ("type", "S1"),
    ("p0", "f8"), ("p1", "f8"), ("p2", "f8"), ("p3", "f8"),
    ("comv_p0", "f8"), ("comv_p1", "f8"), ("comv_p2", "f8"), ("comv_p3", "f8"),
    ("r0", "f8"), ("r1", "f8"), ("r2", "f8"),
    ("s0", "f8"), ("s1", "f8"), ("s2", "f8"), ("s3", "f8"),
    ("num_scatt", "f8"),
    ("recalc_properties", "i4"),
    ("weight", "f8"),
    ("nearest_block_index", "i4"),
    ("time_to_scatter", "f8"),
    ("total_optical_depth", "f8"),
], align=True)
assert PHOTON_DTYPE.itemsize == 176

HYDRO_FIELDS = ["r0", "r1", "r2", "r0_size", "r1_size", "r2_size", "r", "theta", "v0", "v1", "v2",
                "dens", "dens_lab", "pres", "temp", "gamma", "B0", "B1", "B2"]

N_PH_E, N_T = 220, 80  # Src/hot_x_section.h:2-10


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class FrameStats(C.Structure):
    _fields_ = [("iterations", C.c_longlong), ("scatterings", C.c_longlong), ("relocations", C.c_longlong),
                ("photon_slots", C.c_longlong), ("time_now", C.c_double), ("last_time_step", C.c_double),
                ("cs_emitted", C.c_int), ("scatt_cyclosynch_num_ph", C.c_int)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


# --------------------------------------------------------------------------------------
# restatement
# --------------------------------------------------------------------------------------
class McConfig(C.Structure):
    _fields_ = [("dimensions", C.c_int), ("geometry", C.c_int), ("stokes_switch", C.c_int),
                ("tau_calculation", C.c_int), ("cyclosynch_switch", C.c_int), ("b_field_calc", C.c_int),
                ("epsilon_b", C.c_double), ("cs_rebin_e_perc", C.c_double)]


class McHydro(C.Structure):
    _fields_ = ([("num_elements", C.c_int)] + [(f, C.POINTER(C.c_double)) for f in HYDRO_FIELDS] +
                [("r0_domain", C.c_double * 2), ("r1_domain", C.c_double * 2), ("r2_domain", C.c_double * 2),
                 ("fps", C.c_double), ("scatt_frame_number", C.c_int), ("inj_frame_number", C.c_int)])


class McPhotonList(C.Structure):
    _fields_ = [("photons", C.c_void_p), ("sorted_indexes", C.POINTER(C.c_int)), ("num_photons", C.c_int),
                ("num_null_photons", C.c_int), ("list_capacity", C.c_int)]


class McRanlxs(C.Structure):
    _fields_ = [("xdbl", C.c_double * 12), ("ydbl", C.c_double * 12), ("carry", C.c_double),
                ("xflt", C.c_float * 24), ("ir", C.c_uint), ("jr", C.c_uint), ("is_", C.c_uint),
                ("is_old", C.c_uint), ("pr", C.c_uint)]


class McRng(C.Structure):
    _fields_ = [("uniform", C.c_void_p), ("uniform_pos", C.c_void_p), ("get", C.c_void_p), ("set", C.c_void_p),
                ("kind", C.c_int), ("lxs", McRanlxs), ("replay", C.POINTER(C.c_double)), ("replay_n", C.c_size_t),
                ("replay_pos", C.c_size_t), ("tee", C.POINTER(C.c_double)), ("tee_cap", C.c_size_t),
                ("tee_n", C.c_size_t), ("key", C.c_uint32 * 2), ("hint_iter", C.c_uint64),
                ("hint_slot", C.c_uint32), ("hint_stream", C.c_uint32), ("hint_draw", C.c_uint64),
                ("ndraws", C.c_ulonglong)]


def build_oracle(force=False):
    """Compile oracle/liboracle.so (gcc; the checker, never shipped)."""
    lib = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("mcrat_oracle.c", "mc_mathlib.c", "mcrat_oracle.h", "mc_mathlib.h")]
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "liboracle.so"])
    return lib


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        L = C.CDLL(build_oracle())
        L.mc_oracle_new.restype = C.c_void_p
        L.mc_oracle_new.argtypes = [C.POINTER(McConfig)]
        L.mc_oracle_free.argtypes = [C.c_void_p]
        L.mc_oracle_set_thermal_table.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.mc_oracle_set_log.argtypes = [C.c_void_p, C.c_char_p]
        L.mc_oracle_set_iter.argtypes = [C.c_void_p, C.c_ulonglong]
        L.mc_oracle_get_iter.restype = C.c_ulonglong
        L.mc_oracle_get_iter.argtypes = [C.c_void_p]
        L.mc_oracle_set_iter.argtypes = [C.c_void_p, C.c_ulonglong]
        L.mc_oracle_get_iter.restype = C.c_ulonglong
        L.mc_oracle_get_iter.argtypes = [C.c_void_p]
        for name in ("mc_bessel_Kn",):
            getattr(L, name).restype = C.c_double
        L.mc_bessel_Kn.argtypes = [C.c_int, C.c_double]
        L.mc_klein_nishina_cross_section.restype = C.c_double
        L.mc_klein_nishina_cross_section.argtypes = [C.c_double]
        L.mc_photon_event.restype = C.c_double
        L.mc_ph_abs_cyclosynch.restype = C.c_double
        L.mc_average_photon_energy.restype = C.c_double
        L.mc_thermal_cross_section.restype = C.c_double
        L.mc_interpolate_thermal_hot_cross_section.restype = C.c_double
        L.mc_total_thermal_cross_section.restype = C.c_double
        L.mc_single_maxwell_juttner.restype = C.c_double
        L.mc_boosted_cross_section.restype = C.c_double
        L.mc_sample_thermal_electron.restype = C.c_double
        L.mc_hydro_element_volume.restype = C.c_double
        L.mc_magnetic_field_magnitude.restype = C.c_double
        L.mc_cyclosynch_r_limits.restype = C.c_double
        L.mc_ranlxs_get.restype = C.c_ulong
        L.mc_ranlxs_get_double.restype = C.c_double
        _oracle_lib = L
    return _oracle_lib


class OracleRng:
    """mc_rng handle: 'ranlxs0' (seeded), 'replay' (buffer) or 'philox' (keyed)."""

    def __init__(self, kind="ranlxs0", seed=0, shard=0, buf=None):
        self.L = oracle_lib()
        self.r = McRng()
        self._keep = None
        if kind == "ranlxs0":
            self.L.mc_rng_init_ranlxs0(C.byref(self.r), C.c_ulong(seed))
        elif kind == "replay":
            self._keep = np.ascontiguousarray(buf, dtype=np.float64)
            self.L.mc_rng_init_replay(C.byref(self.r), _dp(self._keep), C.c_size_t(self._keep.size))
        elif kind == "philox":
            self.L.mc_rng_init_philox(C.byref(self.r), C.c_uint64(seed), C.c_uint32(shard))
        else:
            raise ValueError(kind)
        self._tee = None

    def tee(self, capacity):
        self._tee = np.zeros(capacity, dtype=np.float64)
        self.L.mc_rng_set_tee(C.byref(self.r), _dp(self._tee), C.c_size_t(capacity))
        return self._tee

    def tee_values(self):
        n = int(self.r.tee_n)
        if n > self._tee.size:
            raise RuntimeError("tee buffer overflow: %d draws > %d" % (n, self._tee.size))
        return self._tee[:n].copy()

    @property
    def ndraws(self):
        return int(self.r.ndraws)

    @property
    def replay_pos(self):
        return int(self.r.replay_pos)

    def ref(self):
        return C.byref(self.r)


class _HydroHolder:
    """Keeps numpy arrays alive behind an mc_hydro struct."""

    def __init__(self, hydro):
        self.arrays = {}
        self.h = McHydro()
        n = int(hydro["num_elements"])
        self.h.num_elements = n
        for f in HYDRO_FIELDS:
            a = np.ascontiguousarray(hydro.get(f, np.zeros(n)), dtype=np.float64)
            if a.size != n:
                raise ValueError("hydro field %s has %d elements, expected %d" % (f, a.size, n))
            self.arrays[f] = a
            setattr(self.h, f, _dp(a))
        for k in ("r0_domain", "r1_domain", "r2_domain"):
            d = hydro.get(k, (0.0, 0.0))
            getattr(self.h, k)[0] = d[0]
            getattr(self.h, k)[1] = d[1]
        self.h.fps = float(hydro.get("fps", 5.0))
        self.h.scatt_frame_number = int(hydro.get("scatt_frame_number", 0))
        self.h.inj_frame_number = int(hydro.get("inj_frame_number", 0))


class Oracle:
    """The C restatement, configured at run time (cfg: dict from oracle.configs)."""

    def __init__(self, cfg, table=None, log=None):
        self.L = oracle_lib()
        self.cfg = dict(cfg)
        c = McConfig(cfg["dimensions"], cfg["geometry"], cfg["stokes"], cfg["tau_calculation"], cfg["cyclosynch"],
                     cfg["b_field_calc"], cfg["epsilon_b"], cfg.get("cs_rebin_e_perc", 0.1))
        self.o = C.c_void_p(self.L.mc_oracle_new(C.byref(c)))
        if table is not None:
            self.set_table(table)
        if log:
            self.L.mc_oracle_set_log(self.o, log.encode())
        self.hydro = None
        self.list = None

    def __del__(self):
        try:
            if self.list is not None:
                self.L.mc_list_free(C.byref(self.list))
            self.L.mc_oracle_free(self.o)
        except Exception:
            pass

    def set_iter(self, it):
        """Position of the keyed streams (while-loop iteration number of the next iteration)."""
        self.L.mc_oracle_set_iter(self.o, C.c_ulonglong(it))

    def get_iter(self):
        return int(self.L.mc_oracle_get_iter(self.o))

    def set_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (N_PH_E + 1, N_T + 1)
        self.L.mc_oracle_set_thermal_table(self.o, _dp(t))

    def set_hydro(self, hydro):
        self.hydro = _HydroHolder(hydro)

    def set_photons(self, photons):
        ph = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
        if self.list is None:
            self.list = McPhotonList()
            self.L.mc_list_init(C.byref(self.list))
        self.L.mc_list_set(C.byref(self.list), ph.ctypes.data_as(C.c_void_p), C.c_int(ph.size))
        # setPhotonList (Src/photons.c:83-108) leaves num_photons at n even if the array holds null photons; the
        # driver never gives it such an array, the tests do: restore the invariant verifyPhotonNum checks
        self.list.num_photons = ph.size - self.list.num_null_photons

    def rebin_cyclosynch_comp_photons(self, max_photons):
        """-> (return value, num_cyclosynch_ph_emit, scatt_cyclosynch_num_ph), Src/mc_cyclosynch.c:600-710"""
        emit, scatt = C.c_int(0), C.c_int(0)
        rc = self.L.mc_rebin_cyclosynch_comp_photons(self.o, C.byref(self.list), C.byref(emit), C.byref(scatt), C.c_int(max_photons))
        return rc, emit.value, scatt.value

    def photons(self):
        n = self.list.list_capacity
        out = np.zeros(n, dtype=PHOTON_DTYPE)
        if n:
            C.memmove(out.ctypes.data, self.list.photons, n * PHOTON_DTYPE.itemsize)
        return out

    def sorted_indexes(self):
        n = self.list.list_capacity
        return np.ctypeslib.as_array(self.list.sorted_indexes, shape=(n,)).copy()

    @property
    def checkinblock_evals(self):
        # instrumentation counter sits after the table arrays; expose through a tiny accessor instead
        raise NotImplementedError

    # --- boundary functions -----------------------------------------------------------
    def find_containing_hydro_cell(self, switch, rng):
        return self.L.mc_find_containing_hydro_cell(self.o, C.byref(self.list), C.byref(self.hydro.h),
                                                    C.c_int(switch), rng.ref())

    def calc_mean_free_path(self, rng):
        self.L.mc_calc_mean_free_path(self.o, C.byref(self.list), C.byref(self.hydro.h), rng.ref())

    def photon_event(self, dt_max, rng):
        idx, sc, ab = C.c_int(0), C.c_int(0), C.c_int(0)
        dt = self.L.mc_photon_event(self.o, C.byref(self.list), C.c_double(dt_max), C.byref(self.hydro.h),
                                    C.byref(idx), C.byref(sc), C.byref(ab), rng.ref())
        return dt, idx.value, sc.value

    def update_photon_position(self, t):
        self.L.mc_update_photon_position(C.byref(self.list), C.c_double(t))

    def run_frame(self, rng, time_now, remaining_time, max_iters=-1, switch=1, cs=None, scatt_cs_num=0):
        st = FrameStats()
        st.scatt_cyclosynch_num_ph = scatt_cs_num
        cs = cs or dict(r_inj=0.0, ph_weight=0.0, max_photons=0, theta_min=0.0, theta_max=0.0)
        self.L.mc_run_frame(self.o, C.byref(self.list), C.byref(self.hydro.h), rng.ref(), C.c_double(time_now),
                            C.c_double(remaining_time), C.c_longlong(max_iters), C.c_int(switch),
                            C.c_double(cs["r_inj"]), C.c_double(cs["ph_weight"]), C.c_int(cs["max_photons"]),
                            C.c_double(cs["theta_min"]), C.c_double(cs["theta_max"]), C.byref(st))
        return st.as_dict()

    def ph_abs_cyclosynch(self):
        na, ns = C.c_int(0), C.c_int(0)
        w = self.L.mc_ph_abs_cyclosynch(self.o, C.byref(self.list), C.byref(na), C.byref(ns), C.byref(self.hydro.h))
        return w, na.value, ns.value

    def photon_emit_cyclosynch(self, rng, r_inj, ph_weight, max_photons, theta_min, theta_max, single=0, scatt_idx=0):
        return self.L.mc_photon_emit_cyclosynch(self.o, C.byref(self.list), C.c_double(r_inj), C.c_double(ph_weight),
                                                C.c_int(max_photons), C.c_double(theta_min), C.c_double(theta_max),
                                                C.byref(self.hydro.h), rng.ref(), C.c_int(single), C.c_int(scatt_idx))

    # --- unit-level ---------------------------------------------------------------------
    def coord_to_hydro(self, x, y, z):
        out = (C.c_double * 3)()
        self.L.mc_coord_to_hydro(self.o, out, C.c_double(x), C.c_double(y), C.c_double(z))
        return np.array(out[:])

    def hydro_vector_to_cartesian(self, v0, v1, v2, x0, x1, x2):
        out = (C.c_double * 3)()
        self.L.mc_hydro_vector_to_cartesian(self.o, out, *[C.c_double(a) for a in (v0, v1, v2, x0, x1, x2)])
        return np.array(out[:])

    def lorentz_boost(self, boost, p, obj="p"):
        b = (C.c_double * 3)(*boost)
        pp = (C.c_double * 4)(*p)
        out = (C.c_double * 4)()
        self.L.mc_lorentz_boost(b, pp, out, C.c_char(obj.encode()))
        return np.array(out[:])

    def single_scatter(self, el, ph, s, rng):
        e = (C.c_double * 4)(*el)
        p = (C.c_double * 4)(*ph)
        ss = (C.c_double * 4)(*s)
        ok = self.L.mc_single_scatter(self.o, e, p, ss, rng.ref())
        return ok, np.array(p[:]), np.array(ss[:])

    def single_thermal_electron(self, temp, ph_p, rng):
        e = (C.c_double * 4)()
        p = (C.c_double * 4)(*ph_p)
        self.L.mc_single_thermal_electron(e, C.c_double(temp), p, rng.ref())
        return np.array(e[:])

    def find_containing_block(self, r0, r1, r2):
        return self.L.mc_find_containing_block(self.o, C.c_double(r0), C.c_double(r1), C.c_double(r2),
                                               C.byref(self.hydro.h))

    def thermal_cross_section(self, comv_e, temp, rng):
        return self.L.mc_thermal_cross_section(self.o, C.c_double(comv_e), C.c_double(temp), rng.ref())


# --------------------------------------------------------------------------------------
# reference sources (oracle/_ref)
# --------------------------------------------------------------------------------------
def ref_lib_path(name, timing=False):
    """timing=True: the -O3 -march=x86-64-v3 build of the same sources (bench.py's CPU arm only; see build_ref.py)."""
    return os.path.join(HERE, "_ref", "libmcrat_ref_%s%s.so" % (name, "_o3" if timing else ""))


def ref_available(name, timing=False):
    return os.path.exists(ref_lib_path(name, timing))


def host_runs_timing_build():
    """The timing build needs AVX2 and FMA (x86-64-v3)."""
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except Exception:
        return False
    return all(f in flags for f in ("avx2", "fma", "bmi2", "movbe"))


class RefLib:
    """One compile-time configuration of the reference's own sources."""

    def __init__(self, name, timing=False):
        path = ref_lib_path(name, timing)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.name = name
        L = self.L = C.CDLL(path)
        L.ref_hydro_new.restype = C.c_void_p
        L.ref_list_new.restype = C.c_void_p
        L.ref_rng_new.restype = C.c_void_p
        L.ref_rng_tee_count.restype = C.c_size_t
        L.ref_rng_draws.restype = C.c_ulonglong
        L.ref_rng_uniform.restype = C.c_double
        L.ref_rng_get.restype = C.c_ulong
        for f in ("ref_photonEvent", "ref_averagePhotonEnergy", "ref_hydroElementVolume",
                  "ref_kleinNishinaCrossSection", "ref_sampleThermalElectron", "ref_singleMaxwellJuttner",
                  "ref_boostedCrossSection", "ref_calculateTotalThermalCrossSection",
                  "ref_interpolateThermalHotCrossSection", "ref_getThermalCrossSection", "ref_calcCyclosynchRLimits",
                  "ref_getMagneticFieldMagnitude", "ref_phAbsCyclosynch"):
            getattr(L, f).restype = C.c_double
        cfg = (C.c_int * 8)()
        L.ref_get_config(cfg)
        self.config = list(cfg)
        self.hydro = None
        self.list = None
        self.rng = None
        self._keep = []

    def set_log(self, path):
        self.L.ref_set_log(path.encode() if path else None)

    def constants(self):
        out = (C.c_double * 10)()
        self.L.ref_get_constants(out)
        return list(out)

    # hydro / list / rng ---------------------------------------------------------------
    def set_hydro(self, hydro):
        if self.hydro:
            self.L.ref_hydro_free(self.hydro)
        n = int(hydro["num_elements"])
        arrs = []
        ptrs = (C.POINTER(C.c_double) * 19)()
        for i, f in enumerate(HYDRO_FIELDS):
            if f in hydro:
                a = np.ascontiguousarray(hydro[f], dtype=np.float64)
                arrs.append(a)
                ptrs[i] = _dp(a)
            else:
                ptrs[i] = None
        dom = np.array(list(hydro.get("r0_domain", (0, 0))) + list(hydro.get("r1_domain", (0, 0))) +
                       list(hydro.get("r2_domain", (0, 0))), dtype=np.float64)
        self.hydro = C.c_void_p(self.L.ref_hydro_new(C.c_int(n), ptrs, _dp(dom), C.c_double(hydro.get("fps", 5.0)),
                                                     C.c_int(hydro.get("scatt_frame_number", 0)),
                                                     C.c_int(hydro.get("inj_frame_number", 0))))
        self.num_elements = n

    def hydro_analytic(self, kind):
        self.L.ref_hydro_analytic(self.hydro, C.c_int(kind), None)

    def hydro_field(self, name):
        out = np.zeros(self.num_elements)
        self.L.ref_hydro_get(self.hydro, C.c_int(HYDRO_FIELDS.index(name)), _dp(out))
        return out

    def set_photons(self, photons):
        if self.list:
            self.L.ref_list_free(self.list)
        ph = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
        self.list = C.c_void_p(self.L.ref_list_new(ph.ctypes.data_as(C.c_void_p), C.c_int(ph.size)))

    def photons(self):
        n = self.L.ref_list_capacity(self.list)
        out = np.zeros(n, dtype=PHOTON_DTYPE)
        if n:
            self.L.ref_list_get(self.list, out.ctypes.data_as(C.c_void_p))
        return out

    def sorted_indexes(self):
        n = self.L.ref_list_capacity(self.list)
        out = np.zeros(n, dtype=np.int32)
        self.L.ref_list_get_sorted(self.list, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def new_rng(self, seed=0, replay=None, tee=None):
        r = C.c_void_p(self.L.ref_rng_new(C.c_ulong(seed)))
        if replay is not None:
            buf = np.ascontiguousarray(replay, dtype=np.float64)
            self._keep.append(buf)
            self.L.ref_rng_use_replay(r, _dp(buf), C.c_size_t(buf.size))
        teebuf = None
        if tee:
            teebuf = np.zeros(tee, dtype=np.float64)
            self._keep.append(teebuf)
            self.L.ref_rng_set_tee(r, _dp(teebuf), C.c_size_t(tee))
        return r, teebuf

    def tee_values(self, r, teebuf):
        n = self.L.ref_rng_tee_count(r)
        if n > teebuf.size:
            raise RuntimeError("tee overflow %d > %d" % (n, teebuf.size))
        return teebuf[:n].copy()

    def set_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (N_PH_E + 1, N_T + 1)
        self.L.ref_set_thermal_table(_dp(t))

    # boundary functions -----------------------------------------------------------------
    def find_containing_hydro_cell(self, switch, rng):
        return self.L.ref_findContainingHydroCell(self.list, self.hydro, C.c_int(switch), rng)

    def calc_mean_free_path(self, rng):
        self.L.ref_calcMeanFreePath(self.list, self.hydro, rng)

    def photon_event(self, dt_max, rng):
        idx, sc, ab = C.c_int(0), C.c_int(0), C.c_int(0)
        dt = self.L.ref_photonEvent(self.list, C.c_double(dt_max), self.hydro, C.byref(idx), C.byref(sc),
                                    C.byref(ab), rng)
        return dt, idx.value, sc.value

    def update_photon_position(self, t):
        self.L.ref_updatePhotonPosition(self.list, C.c_double(t))

    def run_frame(self, rng, time_now, remaining_time, max_iters=-1, switch=1, cs=None, scatt_cs_num=0):
        st = FrameStats()
        st.scatt_cyclosynch_num_ph = scatt_cs_num
        cs = cs or dict(r_inj=0.0, ph_weight=0.0, max_photons=0, theta_min=0.0, theta_max=0.0)
        self.L.ref_run_frame(self.list, self.hydro, rng, C.c_double(time_now), C.c_double(remaining_time),
                             C.c_longlong(max_iters), C.c_int(switch), C.c_double(cs["r_inj"]),
                             C.c_double(cs["ph_weight"]), C.c_int(cs["max_photons"]), C.c_double(cs["theta_min"]),
                             C.c_double(cs["theta_max"]), C.byref(st))
        return st.as_dict()

    def ph_abs_cyclosynch(self):
        na, ns = C.c_int(0), C.c_int(0)
        w = self.L.ref_phAbsCyclosynch(self.list, C.byref(na), C.byref(ns), self.hydro)
        return w, na.value, ns.value

    def photon_emit_cyclosynch(self, rng, r_inj, ph_weight, max_photons, theta_min, theta_max, single=0, scatt_idx=0):
        return self.L.ref_photonEmitCyclosynch(self.list, C.c_double(r_inj), C.c_double(ph_weight),
                                               C.c_int(max_photons), C.c_double(theta_min), C.c_double(theta_max),
                                               self.hydro, rng, C.c_int(single), C.c_int(scatt_idx))

    def rebin_cyclosynch_comp_photons(self, max_photons, theta_min=0.0, theta_max=0.0, rng=None):
        """rebinCyclosynchCompPhotons (Src/mc_cyclosynch.c:600-710) -> (return value, num_cyclosynch_ph_emit, scatt_cyclosynch_num_ph)"""
        emit, scatt = C.c_int(0), C.c_int(0)
        rc = self.L.ref_rebinCyclosynchCompPhotons(self.list, C.byref(emit), C.byref(scatt), C.c_int(max_photons),
                                                   C.c_double(theta_min), C.c_double(theta_max), rng)
        return rc, emit.value, scatt.value

    def photon_injection(self, rng, r_inj, ph_weight, min_photons, max_photons, spect, theta_min, theta_max):
        if not self.list:
            self.list = C.c_void_p(self.L.ref_list_new(None, C.c_int(0)))
        return self.L.ref_photonInjection(self.list, C.c_double(r_inj), C.c_double(ph_weight), C.c_int(min_photons),
                                          C.c_int(max_photons), C.c_char(spect.encode()), C.c_double(theta_min),
                                          C.c_double(theta_max), self.hydro, rng)

    # unit-level ---------------------------------------------------------------------------
    def coord_to_hydro(self, x, y, z):
        out = (C.c_double * 3)()
        self.L.ref_mcratCoordinateToHydroCoordinate(out, C.c_double(x), C.c_double(y), C.c_double(z))
        return np.array(out[:])

    def hydro_vector_to_cartesian(self, v0, v1, v2, x0, x1, x2):
        out = (C.c_double * 3)()
        self.L.ref_hydroVectorToCartesian(out, *[C.c_double(a) for a in (v0, v1, v2, x0, x1, x2)])
        return np.array(out[:])

    def lorentz_boost(self, boost, p, obj="p"):
        b = (C.c_double * 3)(*boost)
        pp = (C.c_double * 4)(*p)
        out = (C.c_double * 4)()
        self.L.ref_lorentzBoost(b, pp, out, C.c_char(obj.encode()))
        return np.array(out[:])

    def single_scatter(self, el, ph, s, rng):
        e = (C.c_double * 4)(*el)
        p = (C.c_double * 4)(*ph)
        ss = (C.c_double * 4)(*s)
        ok = self.L.ref_singleScatter(e, p, ss, rng)
        return ok, np.array(p[:]), np.array(ss[:])

    def single_thermal_electron(self, temp, ph_p, rng):
        e = (C.c_double * 4)()
        p = (C.c_double * 4)(*ph_p)
        self.L.ref_singleThermalElectron(e, C.c_double(temp), p, rng)
        return np.array(e[:])

    def find_containing_block(self, r0, r1, r2):
        return self.L.ref_findContainingBlock(C.c_double(r0), C.c_double(r1), C.c_double(r2), self.hydro)

    def thermal_cross_section(self, comv_e, temp, rng):
        return self.L.ref_getThermalCrossSection(C.c_double(comv_e), C.c_double(temp), rng)

    def kn_cross_section(self, x):
        return self.L.ref_kleinNishinaCrossSection(C.c_double(x))
