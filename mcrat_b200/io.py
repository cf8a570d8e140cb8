"""ctypes binding of libmcrat_b200_io.so (include/mcrat_b200_io.h): mc.par, mcrat_input.h and the
mc_proc / mcdata HDF5 layout.  Harness plumbing only -- the product is the C library."""
import ctypes as C
import os

import numpy as np

from .lib import CSRC, Config, build
from .synth import PHOTON_DTYPE

IO_LIB_PATH = os.path.join(CSRC, "libmcrat_b200_io.so")
MAX_BINS = 64

IO_EXPORTS = ["mcrat_b200_read_mc_par", "mcrat_b200_config_from_input_header", "mcrat_b200_print_photons",
              "mcrat_b200_merge_frame", "mcrat_b200_h5_dataset_length", "mcrat_b200_h5_read_dataset",
              "mcrat_b200_h5_list", "mcrat_b200_io_last_error"]


class McPar(C.Structure):
    _fields_ = [("fps", C.c_double), ("last_frame", C.c_int), ("r0_domain", C.c_double * 2), ("r1_domain", C.c_double * 2),
                ("r2_domain", C.c_double * 2), ("theta_jmin", C.c_double), ("theta_j", C.c_double), ("n_theta_j", C.c_int),
                ("frm0", C.c_int * MAX_BINS), ("frm2", C.c_int * MAX_BINS), ("inj_radius", C.c_double * MAX_BINS),
                ("spect", C.c_char), ("min_photons", C.c_int), ("max_photons", C.c_int), ("restart", C.c_char)]


class IoSwitches(C.Structure):
    _fields_ = [("comv_switch", C.c_int), ("save_type", C.c_int), ("stokes_switch", C.c_int), ("sim_switch", C.c_int),
                ("simulation_type", C.c_int), ("mc_path", C.c_char * 256), ("filepath", C.c_char * 256),
                ("fileroot", C.c_char * 256), ("mcpar", C.c_char * 64)]


class McratIoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("io error %d: %s" % (code, msg))
        self.code = code


_io = None


def load_io():
    global _io
    if _io is None:
        if not os.path.exists(IO_LIB_PATH):
            build()
        L = C.CDLL(IO_LIB_PATH)
        L.mcrat_b200_io_last_error.restype = C.c_char_p
        L.mcrat_b200_h5_dataset_length.restype = C.c_longlong
        _io = L
    return _io


def _ck(rc):
    if rc < 0:
        raise McratIoError(rc, load_io().mcrat_b200_io_last_error().decode())
    return rc


def read_mc_par(path):
    p = McPar()
    _ck(load_io().mcrat_b200_read_mc_par(path.encode(), C.byref(p)))
    n = p.n_theta_j
    return dict(fps=p.fps, last_frame=p.last_frame, r0_domain=tuple(p.r0_domain), r1_domain=tuple(p.r1_domain),
                r2_domain=tuple(p.r2_domain), theta_jmin=p.theta_jmin, theta_j=p.theta_j, n_theta_j=n,
                frm0=list(p.frm0[:n]), frm2=list(p.frm2[:n]), inj_radius=list(p.inj_radius[:n]),
                spect=p.spect.decode(), min_photons=p.min_photons, max_photons=p.max_photons, restart=p.restart.decode())


def config_from_input_header(path):
    cfg, sw = Config(), IoSwitches()
    _ck(load_io().mcrat_b200_config_from_input_header(path.encode(), C.byref(cfg), C.byref(sw)))
    return cfg, sw


def switches(comv=1, save_type=1, stokes=1):
    sw = IoSwitches()
    sw.comv_switch, sw.save_type, sw.stokes_switch = comv, save_type, stokes
    return sw


def print_photons(directory, angle_rank, frame, photons, sw):
    ph = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
    _ck(load_io().mcrat_b200_print_photons(directory.encode(), C.c_int(angle_rank), C.c_int(frame),
                                           ph.ctypes.data_as(C.c_void_p), C.c_int(ph.size), C.byref(sw)))


def merge_frame(directory, frame, ranks, sw):
    r = (C.c_int * len(ranks))(*ranks)
    _ck(load_io().mcrat_b200_merge_frame(directory.encode(), C.c_int(frame), r, C.c_int(len(ranks)), C.byref(sw)))


def h5_list(path, group=""):
    buf = C.create_string_buffer(1 << 16)
    n = _ck(load_io().mcrat_b200_h5_list(path.encode(), group.encode(), buf, C.c_size_t(len(buf))))
    names = [x for x in buf.value.decode().split("\n") if x]
    assert len(names) == n
    return names


def h5_read(path, name):
    L = load_io()
    n = _ck(L.mcrat_b200_h5_dataset_length(path.encode(), name.encode()))
    if name.split("/")[-1] == "PT":
        out = np.zeros(n, dtype=np.int8)
        _ck(L.mcrat_b200_h5_read_dataset(path.encode(), name.encode(), None, out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
    else:
        out = np.zeros(n, dtype=np.float64)
        _ck(L.mcrat_b200_h5_read_dataset(path.encode(), name.encode(), out.ctypes.data_as(C.c_void_p), None, C.c_size_t(n)))
    return out
