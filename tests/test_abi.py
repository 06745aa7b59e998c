"""The C-ABI library loads and exports every symbol include/s2m.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "s2m.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2m_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(s2m):
    assert declared_symbols() == sorted(s2m.EXPORTS)


def test_library_exports_every_symbol(s2m, built):
    lib = ctypes.CDLL(s2m.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes(s2m, built):
    # s2m_params: 2 floats + 11 ints; s2m_stats: 15 ints (+pad) + 4 doubles
    assert ctypes.sizeof(s2m.Params) == 52
    assert ctypes.sizeof(s2m.Stats) == 96
    p = s2m.default_params()
    assert abs(p.line_res - 0.4) < 1e-7 and abs(p.plane_res - 0.8) < 1e-7 and p.batch == 1


def test_strerror(s2m, built):
    L = s2m.load_library()
    assert L.s2m_strerror(0) == b"ok"
    assert b"not enough" in L.s2m_strerror(1)


def test_no_cpu_fallback(s2m, built):
    """Without a CUDA device the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(s2m.S2MError):
        s2m.Registrar()


def test_product_never_touches_oracle():
    pk = os.path.join(ROOT, "sc-a-loam_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "orc_" not in txt and "libs2m_oracle" not in txt, f


PCD_HEADER_37101 = (b"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                    b"TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 37101\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 37101\nDATA binary\n")


def test_pcd_layout_is_the_reference_sample_layout(s2m, built, tmp_path):
    """s2m_pcd_write emits, byte for byte, the 188-byte header of the reference's shipped
    utils/sample_data/KAIST03/Scans/000000.pcd (37101 points) followed by 16 B per point;
    s2m_pcd_read gives the floats back. Host-only: no GPU involved."""
    import numpy as np
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(37101, 4)).astype(np.float32)
    path = str(tmp_path / "a.pcd")
    s2m.pcd_write(path, pts)
    raw = open(path, "rb").read()
    assert len(PCD_HEADER_37101) == 188 and raw[:188] == PCD_HEADER_37101 and len(raw) == 188 + 16 * 37101
    assert np.array_equal(s2m.pcd_read(path).view(np.uint32), pts.view(np.uint32))
    s2m.pcd_write(path, np.zeros((0, 4), np.float32))
    assert len(s2m.pcd_read(path)) == 0
    import pytest
    with pytest.raises(s2m.S2MError):
        s2m.pcd_read(str(tmp_path / "missing.pcd"))
    open(path, "wb").write(PCD_HEADER_37101.replace(b"DATA binary", b"DATA ascii"))
    with pytest.raises(s2m.S2MError):
        s2m.pcd_read(path)
    # the reference's own files, when the tree is mounted (authoring container only)
    sample = "/root/reference/utils/sample_data/KAIST03/Scans/000000.pcd"
    if os.path.exists(sample):
        got = s2m.pcd_read(sample)
        raw = open(sample, "rb").read()
        assert raw[:188] == PCD_HEADER_37101
        assert np.array_equal(got.view(np.uint32), np.frombuffer(raw[188:188 + 16 * 37101], np.uint32).reshape(-1, 4))


def test_plain_c_consumer_of_the_header(s2m, built, tmp_path):
    """tests/c/abi_smoke.c includes include/s2m.h as C99, links libs2m.so and calls the host-only entry points."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(s2m.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-ls2m",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, str(tmp_path / "t.pcd")], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "s2m_params 52 bytes, s2m_stats 96 bytes" in out.stdout


def test_ros_node_source_compiles(s2m, built, tmp_path):
    """integration/alaserMapping_s2m.cpp (the drop-in for laserMapping.cpp:908-953) is
    type-checked against stand-ins for the ROS1 / PCL headers it includes (tests/c/ros_stubs: only the
    signatures it uses) with -Wall -Wextra -Werror, and linked against libs2m.so: every s2m_* call in the
    node matches include/s2m.h and resolves in the library."""
    import subprocess
    src = os.path.join(ROOT, "integration", "alaserMapping_s2m.cpp")
    inc = ["-I", os.path.join(ROOT, "tests", "c", "ros_stubs"), "-I", os.path.join(ROOT, "include")]
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only"] + inc + [src])
    libdir = os.path.dirname(s2m.LIB_PATH)
    exe = str(tmp_path / "alaserMapping_s2m")
    subprocess.check_call(["g++", "-std=c++17"] + inc + [src, "-o", exe, "-L", libdir, "-ls2m", "-lpthread",
                                                        "-Wl,-rpath," + libdir])
    nm = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    used = sorted(set(re.findall(r"\b(s2m_[a-z0-9_]+)", nm)))
    assert {"s2m_create", "s2m_register", "s2m_transform_cloud", "s2m_get_correction", "s2m_get_surround",
            "s2m_map_download", "s2m_destroy"} <= set(used)
