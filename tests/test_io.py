"""Host-side configuration and output surface: mc.par, mcrat_input.h,
mc_proc_<rank>.h5 / mcdata_<frame>.h5.  CPU only."""
import glob
import os

import numpy as np
import pytest

from mcrat_b200 import io as mio
from mcrat_b200 import lib, synth

from h5spec import H5File

SAMPLE_MC_PAR = """[Hydro/MHD Simulation Block]

5.               # Number of frames per second of hydro simulation (likely always the same)
3000\t\t# Last available hydro simulation frame  (get it from the last file in the data folder)
0 5e12\t\t# Max r0 coordinate limits of hydro simulation
0 2.5e12\t\t# Max r1 coordinate limit of hydro simulation
0 2e13\t\t# Max r2 coordinate limit of hydro simulation (if simulation is 3D)

[MCRaT Injection Angles Block]

0.               \t# The minimum off-axis angle to inject photons (in degrees)
6.               \t# The maximum off-axis angle to inject photons (in degrees)
3.\t\t\t# Number of angle bins to consider
200 200 200      \t# Frame at which photon injection starts for each angle bin
2 2 2            \t# Number of frames for which photons are injected for each angle bin
1e11 1.5e12 2e12\t# The radius at which the photons are injected for each angle bin

[MCRaT Photon Block]

b\t\t# Type of spectrum we inject with, w=wien b=blackbody
1000\t\t# Min number of photons
5000\t\t# Max number of photons

[Initialization/Continuation Block]

i\t\t# Initialize or continue simulation (i=initialize (delete all files) c=continue)         
"""

REAL_HDF5 = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "scipy", "io", "matlab", "tests", "data",
                                   "testhdf5_7.4_GLNX86.mat"))


def test_io_library_exports_every_declared_symbol():
    L = mio.load_io()
    header = open(os.path.join(os.path.dirname(lib.HERE), "include", "mcrat_b200_io.h")).read()
    for name in mio.IO_EXPORTS:
        assert name + "(" in header, name
        assert hasattr(L, name), name


def test_mc_par_matches_reference_sample(tmp_path):
    """sample_mc.par:1-25 parsed like readMcPar (Src/mcrat_io.c:1136-1237): frm2 = frm0 + count (:1201)."""
    p = tmp_path / "mc.par"
    p.write_text(SAMPLE_MC_PAR)
    ref = "/root/reference/sample_mc.par"
    if os.path.exists(ref):  # the embedded copy is the reference's own file
        assert open(ref).read().split() == SAMPLE_MC_PAR.split()
    par = mio.read_mc_par(str(p))
    assert par["fps"] == 5.0 and par["last_frame"] == 3000
    assert par["r0_domain"] == (0.0, 5e12) and par["r1_domain"] == (0.0, 2.5e12) and par["r2_domain"] == (0.0, 2e13)
    assert (par["theta_jmin"], par["theta_j"], par["n_theta_j"]) == (0.0, 6.0, 3)
    assert par["frm0"] == [200, 200, 200] and par["frm2"] == [202, 202, 202]
    assert par["inj_radius"] == [float(np.float32(1e11)), float(np.float32(1.5e12)), float(np.float32(2e12))]  # strtof, :1212
    assert (par["spect"], par["min_photons"], par["max_photons"], par["restart"]) == ("b", 1000, 5000, "i")


def test_mc_par_errors(tmp_path):
    p = tmp_path / "bad.par"
    p.write_text("[Hydro/MHD Simulation Block]\n\n5.\n3000\n0 5e12\n")
    with pytest.raises(mio.McratIoError) as e:
        mio.read_mc_par(str(p))
    assert "r1 domain" in str(e.value)
    with pytest.raises(mio.McratIoError):
        mio.read_mc_par(str(tmp_path / "missing.par"))


def test_config_from_the_shipped_input_header(tmp_path):
    """The reference's shipped Src/mcrat_input.h: 2.5-D cylindrical PLUTO, Stokes ON, COMV ON, CS OFF, SAVE_TYPE ON;
    defaults of Src/mcrat.h:262-427 for what it does not define (TAU_CALCULATION DIRECT, B_FIELD_CALC TOTAL_E, EPSILON_B 0.5)."""
    text = """
//#define SIMULATION_TYPE SPHERICAL_OUTFLOW
/* #define GEOMETRY SPHERICAL
   #define DIMENSIONS THREE */
#define SIMULATION_TYPE CYLINDRICAL_OUTFLOW
#define FILEPATH "/data/LEO_2.5D_MHD_PLUTO/BPT5/"
#define FILEROOT "data."
#define MC_PATH "MCRAT_TEST/"
#define     SIM_SWITCH                  PLUTO
#define     GEOMETRY                    CYLINDRICAL
#define     DIMENSIONS                  TWO_POINT_FIVE
////#define     B_FIELD_CALC                SIMULATION
#define     HYDRO_L_SCALE               1e12
#define     STOKES_SWITCH               ON
#define     COMV_SWITCH                 ON
#define     HYDRO_D_SCALE               1
#define     CYCLOSYNCHROTRON_SWITCH     OFF
#define     SAVE_TYPE                   ON
#define     MCPAR                   "mc.par"
"""
    p = tmp_path / "mcrat_input.h"
    p.write_text(text)
    cfg, sw = mio.config_from_input_header(str(p))
    assert (cfg.abi_version, cfg.dimensions, cfg.geometry, cfg.stokes_switch) == (lib.ABI_VERSION, 1, 2, 1)
    assert (cfg.tau_calculation, cfg.cyclosynch_switch, cfg.b_field_calc, cfg.epsilon_b) == (1, 0, 1, 0.5)
    assert (sw.comv_switch, sw.save_type, sw.stokes_switch, sw.sim_switch, sw.simulation_type) == (1, 1, 1, 2, 1)
    assert sw.mc_path == b"MCRAT_TEST/" and sw.mcpar == b"mc.par" and sw.fileroot == b"data."
    ref = "/root/reference/Src/mcrat_input.h"
    if os.path.exists(ref):  # the real file gives the same configuration
        cfg2, sw2 = mio.config_from_input_header(ref)
        assert (cfg2.dimensions, cfg2.geometry, cfg2.stokes_switch, cfg2.cyclosynch_switch) == (1, 2, 1, 0)
        assert (sw2.comv_switch, sw2.save_type, sw2.sim_switch) == (1, 1, 2)
    # the oracle's generated headers round-trip through the parser
    from oracle import configs
    for name, c in configs.CONFIGS.items():
        q = tmp_path / (name + ".h")
        q.write_text(configs.input_header(c) + "#define HYDRO_L_SCALE 1.0\n#define HYDRO_D_SCALE 1.0\n")
        cfg3, sw3 = mio.config_from_input_header(str(q))
        assert (cfg3.dimensions, cfg3.geometry, cfg3.stokes_switch, cfg3.tau_calculation, cfg3.cyclosynch_switch) == \
            (c["dimensions"], c["geometry"], c["stokes"], c["tau_calculation"], c["cyclosynch"]), name
        if c["cyclosynch"]:
            assert cfg3.b_field_calc == c["b_field_calc"], name


def test_config_header_errors_mirror_the_reference(tmp_path):
    p = tmp_path / "h.h"
    p.write_text('#define SIM_SWITCH FLASH\n#define GEOMETRY CARTESIAN\n#define HYDRO_L_SCALE 1\n#define HYDRO_D_SCALE 1\n#define MCPAR "mc.par"\n')
    with pytest.raises(mio.McratIoError) as e:  # Src/mcrat.h:408-410
        mio.config_from_input_header(str(p))
    assert "DIMENSIONS" in str(e.value)
    p.write_text('#define SIM_SWITCH FLASH\n#define DIMENSIONS TWO\n#define GEOMETRY POLAR\n#define HYDRO_L_SCALE 1\n'
                 '#define HYDRO_D_SCALE 1\n#define MCPAR "mc.par"\n')
    with pytest.raises(mio.McratIoError):
        mio.config_from_input_header(str(p))


@pytest.mark.skipif(not REAL_HDF5, reason="no libhdf5-written file in this environment")
def test_spec_parser_and_c_reader_agree_on_a_file_written_by_libhdf5():
    """Pins both readers on a file the real HDF5 library wrote (MATLAB 7.3 = HDF5 behind a 512-byte user block)."""
    path = REAL_HDF5[0]
    f = H5File(path)
    assert f.base == 512 and f.sb_version == 0
    tree = f.tree()
    assert list(tree) == ["testdouble"]
    want = tree["testdouble"]
    assert want.shape == (9, 1) and want.dtype == np.float64
    got = mio.h5_read(path, "testdouble")
    assert np.array_equal(got, want.ravel())
    assert np.allclose(got, np.arange(9) * np.pi / 4)  # scipy's test data: 0 .. 2 pi in pi/4 steps
    assert mio.h5_list(path) == ["testdouble"]


def _photons(n, seed):
    cfg, hydro, ph, frame = synth.workload("C2", scale=1.0 / 32, n_photons=n, seed=seed)
    rng = np.random.default_rng(seed)
    ph["num_scatt"] = rng.integers(0, 50, n)
    ph["type"] = rng.choice(np.frombuffer(b"ikc", dtype="S1"), n)
    ph["s1"], ph["s2"] = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    dead = rng.choice(n, n // 7, replace=False)  # null photons are not saved (Src/mcrat_io.c:154)
    ph["weight"][dead] = 0
    return ph


NAMES = ["P0", "P1", "P2", "P3", "COMV_P0", "COMV_P1", "COMV_P2", "COMV_P3", "R0", "R1", "R2", "S0", "S1", "S2", "S3",
         "NS", "PW", "PT"]
FIELD = dict(P0="p0", P1="p1", P2="p2", P3="p3", COMV_P0="comv_p0", COMV_P1="comv_p1", COMV_P2="comv_p2", COMV_P3="comv_p3",
             R0="r0", R1="r1", R2="r2", S0="s0", S1="s1", S2="s2", S3="s3", NS="num_scatt", PW="weight")


def _expect(ph, name):
    live = ph[ph["weight"] != 0]
    if name == "PT":
        return np.frombuffer(live["type"].tobytes(), dtype=np.int8)
    return live[FIELD[name]].astype(np.float64)


def test_mc_proc_and_mcdata_layout(tmp_path):
    """printPhotons + dirFileMerge: group per frame, the reference's dataset names and types, append on
    a second call for the same frame, ranks concatenated in id order; checked with the independent parser."""
    d = str(tmp_path)
    sw = mio.switches(comv=1, save_type=1, stokes=1)
    lists = {(r, fr): _photons(300 + 17 * r + fr, seed=10 * r + fr) for r in (0, 1, 2) for fr in (200, 201, 1000)}
    for (r, fr), ph in lists.items():
        mio.print_photons(d, r, fr, ph, sw)
    extra = _photons(123, seed=99)  # a later injection block writes into an existing frame group (:451-530)
    mio.print_photons(d, 1, 201, extra, sw)
    for r in (0, 1, 2):
        tree = H5File(os.path.join(d, "mc_proc_%d.h5" % r)).tree()
        assert sorted(tree) == ["1000", "200", "201"]  # strcmp order, as libhdf5 keeps links
        for fr in (200, 201, 1000):
            g = tree[str(fr)]
            assert sorted(g) == sorted(NAMES)
            for name in NAMES:
                want = _expect(lists[(r, fr)], name)
                if (r, fr) == (1, 201):
                    want = np.concatenate([want, _expect(extra, name)])
                assert g[name].dtype == (np.int8 if name == "PT" else np.float64)
                assert np.array_equal(g[name], want), (r, fr, name)
                assert np.array_equal(mio.h5_read(os.path.join(d, "mc_proc_%d.h5" % r), "%d/%s" % (fr, name)), want)
    mio.merge_frame(d, 201, [0, 1, 2, 7], sw)  # rank 7 never wrote: skipped
    tree = H5File(os.path.join(d, "mcdata_201.h5")).tree()
    assert sorted(tree) == sorted(NAMES)
    for name in NAMES:
        want = np.concatenate([_expect(lists[(0, 201)], name), _expect(lists[(1, 201)], name), _expect(extra, name),
                               _expect(lists[(2, 201)], name)])
        assert np.array_equal(tree[name], want), name
    assert mio.h5_list(os.path.join(d, "mcdata_201.h5")) == sorted(NAMES)


def test_switches_select_datasets_and_empty_lists_are_legal(tmp_path):
    d = str(tmp_path)
    sw = mio.switches(comv=0, save_type=0, stokes=0)
    ph = _photons(50, seed=3)
    mio.print_photons(d, 0, 5, ph, sw)
    ph["weight"] = 0
    mio.print_photons(d, 0, 6, ph, sw)  # nothing alive: zero-length datasets
    tree = H5File(os.path.join(d, "mc_proc_0.h5")).tree()
    assert sorted(tree["5"]) == sorted(["P0", "P1", "P2", "P3", "R0", "R1", "R2", "NS", "PW"])
    assert all(v.size == 0 for v in tree["6"].values())
    mio.merge_frame(d, 6, [0], sw)
    assert all(v.size == 0 for v in H5File(os.path.join(d, "mcdata_6.h5")).tree().values())
    with pytest.raises(mio.McratIoError):
        mio.h5_read(os.path.join(d, "mc_proc_0.h5"), "5/S0")


def test_many_frames_in_one_file(tmp_path):
    """Hundreds of frame groups (a full run) still form one valid symbol-table group."""
    d = str(tmp_path)
    sw = mio.switches(comv=0, save_type=1, stokes=0)
    ph = _photons(20, seed=1)
    for fr in range(200, 200 + 150):
        mio.print_photons(d, 3, fr, ph, sw)
    f = H5File(os.path.join(d, "mc_proc_3.h5"))
    tree = f.tree()
    assert len(tree) == 150 and f.leaf_k >= 75
    assert np.array_equal(tree["349"]["PW"], _expect(ph, "PW"))


def _hand_built_chunked_file(path, values, chunk):
    """A minimal HDF5 file with one 1-D *chunked*, extendible F64 dataset "PW" at the root, laid out by hand from the
    format specification (layout message v3 class 2, B-tree v1 node type 1) -- the storage the reference's per-rank
    files use (H5Pset_chunk + H5S_UNLIMITED, Src/mcrat_io.c:253-262)."""
    import struct
    UNDEF = 0xFFFFFFFFFFFFFFFF
    n = len(values)
    nchunks = (n + chunk - 1) // chunk
    K = 4
    # addresses
    a_root_ohdr = 96
    a_heap = a_root_ohdr + 16 + 24 + 8
    a_heap_data = a_heap + 32
    heap_size = 8 + 8 + 32
    a_btree = a_heap_data + heap_size
    a_snod = a_btree + 24 + 33 * 8 + 32 * 8
    a_dset = a_snod + 8 + 2 * K * 40
    dset_msgs = (8 + 8) + (8 + 24) + (8 + 24) + (8 + 24)
    a_cbtree = a_dset + 16 + dset_msgs
    a_chunks = a_cbtree + 24 + (2 * 32 + 1) * 24 + 2 * 32 * 8
    eof = a_chunks + nchunks * chunk * 8
    b = bytearray(eof)
    b[0:8] = b"\x89HDF\r\n\x1a\n"
    b[13], b[14] = 8, 8
    struct.pack_into("<HH", b, 16, K, 16)
    struct.pack_into("<4Q", b, 24, 0, UNDEF, eof, UNDEF)
    struct.pack_into("<QQIIQQ", b, 56, 0, a_root_ohdr, 1, 0, a_btree, a_heap)
    # root group
    struct.pack_into("<BBHII", b, a_root_ohdr, 1, 0, 2, 1, 32)
    struct.pack_into("<HHB3xQQ", b, a_root_ohdr + 16, 0x11, 16, 0, a_btree, a_heap)
    b[a_heap:a_heap + 4] = b"HEAP"
    struct.pack_into("<QQQ", b, a_heap + 8, heap_size, 16, a_heap_data)
    b[a_heap_data + 8:a_heap_data + 10] = b"PW"
    struct.pack_into("<QQ", b, a_heap_data + 16, 1, 32)
    b[a_btree:a_btree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", b, a_btree + 4, 0, 0, 1, UNDEF, UNDEF)
    struct.pack_into("<QQQ", b, a_btree + 24, 0, a_snod, 8)
    b[a_snod:a_snod + 4] = b"SNOD"
    struct.pack_into("<BBH", b, a_snod + 4, 1, 0, 1)
    struct.pack_into("<QQII", b, a_snod + 8, 8, a_dset, 0, 0)
    # dataset header: fill value, datatype, dataspace (with maximum = unlimited), chunked layout
    struct.pack_into("<BBHII", b, a_dset, 1, 0, 4, 1, dset_msgs)
    m = a_dset + 16
    struct.pack_into("<HHB3x", b, m, 0x05, 8, 1)
    b[m + 8:m + 12] = bytes([1, 1, 2, 1])
    m += 16
    struct.pack_into("<HHB3x", b, m, 0x03, 24, 1)
    b[m + 8:m + 12] = bytes([0x11, 0x20, 0x3F, 0])
    struct.pack_into("<IHHBBBBI", b, m + 12, 8, 0, 64, 52, 11, 0, 52, 1023)
    m += 32
    struct.pack_into("<HHB3x", b, m, 0x01, 24, 0)
    b[m + 8:m + 11] = bytes([1, 1, 1])                      # version 1, rank 1, flags: maximum dimensions present
    struct.pack_into("<QQ", b, m + 16, n, UNDEF)
    m += 32
    struct.pack_into("<HHB3x", b, m, 0x08, 24, 0)
    b[m + 8:m + 11] = bytes([3, 2, 2])                      # version 3, chunked, dimensionality = rank + 1
    struct.pack_into("<QII", b, m + 11, a_cbtree, chunk, 8)  # B-tree address, chunk dims (elements), element size
    # chunk B-tree: node type 1, leaf
    b[a_cbtree:a_cbtree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", b, a_cbtree + 4, 1, 0, nchunks, UNDEF, UNDEF)
    p = a_cbtree + 24
    data = np.zeros(nchunks * chunk)
    data[:n] = values
    for c in range(nchunks):
        struct.pack_into("<IIQQ", b, p, chunk * 8, 0, c * chunk, 0)  # key: chunk bytes, filter mask, offsets (dim 0, element)
        struct.pack_into("<Q", b, p + 24, a_chunks + c * chunk * 8)
        p += 32
    struct.pack_into("<IIQQ", b, p, 0, 0, nchunks * chunk, 0)        # final key
    b[a_chunks:eof] = data.astype("<f8").tobytes()
    open(path, "wb").write(bytes(b))


def test_reader_takes_chunked_extendible_datasets(tmp_path):
    """The reference's own mc_proc files store every dataset chunked (Src/mcrat_io.c:253-262); the C reader must take
    that layout (un-filtered), including a last chunk that extends past the dataset."""
    vals = np.linspace(1e48, 3e50, 1234)
    path = str(tmp_path / "chunked.h5")
    _hand_built_chunked_file(path, vals, chunk=500)
    assert mio.h5_list(path) == ["PW"]
    assert np.array_equal(mio.h5_read(path, "PW"), vals)


def test_mc_proc_datasets_are_chunked_and_extendible_like_the_reference(tmp_path):
    """printPhotons creates every dataset with H5Pset_chunk(dims = photons of the first write) and an unlimited maximum
    dimension, then H5Dset_extent + a hyperslab write per append (Src/mcrat_io.c:140, 254-263, 423-705); dirFileMerge
    creates plain contiguous datasets (:1649).  Checked with the independent parser: chunk size, maximum dimension,
    chunk B-tree invariants, data."""
    from h5spec import UNDEF
    d = str(tmp_path)
    sw = mio.switches(comv=1, save_type=1, stokes=1)
    first, second, third = _photons(211, seed=5), _photons(97, seed=6), _photons(500, seed=7)
    n1 = int((first["weight"] != 0).sum())
    for ph in (first, second, third):
        mio.print_photons(d, 0, 300, ph, sw)
    f = H5File(os.path.join(d, "mc_proc_0.h5"))
    g = f.tree()["300"]
    links = f.object(f.object(f.root_ohdr)[1]["300"])[1]
    for name in NAMES:
        kind, arr = f.object(links[name])
        assert kind == "dataset"
        assert f.last_dataset_info["chunk"] == n1, name                    # the first write's count, kept by the appends
        assert f.last_dataset_info["maxdims"] == (UNDEF,), name            # H5S_UNLIMITED
        want = np.concatenate([_expect(first, name), _expect(second, name), _expect(third, name)])
        assert np.array_equal(arr, want) and np.array_equal(g[name], want), name
        assert np.array_equal(mio.h5_read(os.path.join(d, "mc_proc_0.h5"), "300/" + name), want)
    mio.merge_frame(d, 300, [0], sw)
    m = H5File(os.path.join(d, "mcdata_300.h5"))
    for name, oh in m.object(m.root_ohdr)[1].items():
        m.object(oh)
        assert m.last_dataset_info["chunk"] is None and m.last_dataset_info["maxdims"] is None, name


def test_many_appends_build_a_chunk_tree_with_internal_nodes(tmp_path):
    """More chunks than one B-tree node holds (2K = 64): leaves linked left to right under an internal level."""
    d = str(tmp_path)
    sw = mio.switches(comv=0, save_type=1, stokes=0)
    blocks = [_photons(8, seed=100 + k) for k in range(150)]
    for ph in blocks:
        mio.print_photons(d, 2, 77, ph, sw)
    f = H5File(os.path.join(d, "mc_proc_2.h5"))
    g = f.tree()["77"]
    n0 = int((blocks[0]["weight"] != 0).sum())
    for name in ("P0", "R2", "NS", "PW", "PT"):
        want = np.concatenate([_expect(ph, name) for ph in blocks])
        assert want.size > 64 * n0, "the case must need more than one leaf"
        assert np.array_equal(g[name], want), name
        assert np.array_equal(mio.h5_read(os.path.join(d, "mc_proc_2.h5"), "77/" + name), want), name


def test_append_after_an_empty_first_write(tmp_path):
    """A frame group created by a write with no live photon (chunk size 1, since a chunk cannot be empty), then filled by an
    append: hundreds of one-element chunks, three B-tree levels for the larger block -- still the right data in both readers."""
    d = str(tmp_path)
    sw = mio.switches(comv=0, save_type=1, stokes=0)
    dead = _photons(40, seed=2)
    dead["weight"] = 0
    mio.print_photons(d, 0, 9, dead, sw)
    first, second = _photons(90, seed=3), _photons(5000, seed=4)
    mio.print_photons(d, 0, 9, first, sw)
    mio.print_photons(d, 0, 9, second, sw)
    f = H5File(os.path.join(d, "mc_proc_0.h5"))
    g = f.tree()["9"]
    links = f.object(f.object(f.root_ohdr)[1]["9"])[1]
    for name in ("P0", "R1", "NS", "PW", "PT"):
        want = np.concatenate([_expect(first, name), _expect(second, name)])
        assert want.size > 64 * 64, "needs a third level of the chunk tree"
        assert np.array_equal(g[name], want), name
        assert np.array_equal(mio.h5_read(os.path.join(d, "mc_proc_0.h5"), "9/" + name), want), name
        f.object(links[name])
        assert f.last_dataset_info["chunk"] == 1
