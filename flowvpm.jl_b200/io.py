"""`save(pfield, file_name; path, add_num, num, createpath, overwrite_time)` of the reference
(src/FLOWVPM_utils.jl:148-371) for fields whose truth lives on the GPU.

The reference writes an HDF5 file plus an XDMF descriptor for ParaView.  No HDF5 library exists in
this image, and none is needed: XDMF 3 describes raw binary files just as well.  This writer
keeps the reference's dataset names, attribute names, shapes ([np, 3] / [np]) and the XDMF layout,
but the DataItems are `Format="Binary"` windows (`Seek` = byte offset) into ONE little-endian
`<name>.<num>.bin` written next to the `.xmf`; the three header scalars of the HDF5 file (np, nt, t)
go into the first 24 bytes.  A `ResidentField` is downloaded once (one contiguous D2H of the mirror)
right before writing, so a time loop that never leaves the GPU can still dump every
`nsteps_save` steps (src/FLOWVPM_utils.jl:52,121-128).
"""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .particlefield import (C_INDEX, GAMMA_INDEX, J_INDEX, SIGMA_INDEX, STATIC_INDEX, U_INDEX,
                            VORTICITY_INDEX, X_INDEX)

VOL_ROW, CIRCULATION_ROW = 7, 8


def _rows(idx):
    if isinstance(idx, slice):
        return list(range(idx.start, idx.stop))
    return [int(idx)]


def _datasets(P, n, les):
    """(name, array[np, dim], xdmf attribute name or None, type) in the reference's order (:190-222)"""
    J = _rows(J_INDEX)
    out = [
        ("X", P[_rows(X_INDEX), :n], None, "Vector"),
        ("Gamma", P[_rows(GAMMA_INDEX), :n], "Gamma", "Vector"),
        ("sigma", P[_rows(SIGMA_INDEX), :n], "sigma", "Scalar"),
        ("circulation", P[[CIRCULATION_ROW], :n], "circulation", "Scalar"),
        ("vol", P[[VOL_ROW], :n], "vol", "Scalar"),
        ("static", P[_rows(STATIC_INDEX), :n], "static", "Scalar"),
        ("velocity", P[_rows(U_INDEX), :n], "velocity", "Vector"),
        ("velocity_gradient_x", P[J[0:3], :n], "velocity gradient x", "Vector"),
        ("velocity_gradient_y", P[J[3:6], :n], "velocity gradient y", "Vector"),
        ("velocity_gradient_z", P[J[6:9], :n], "velocity gradient z", "Vector"),
        ("vorticity", P[_rows(VORTICITY_INDEX), :n], "vorticity", "Vector"),
    ]
    if les:
        out.append(("C", P[_rows(C_INDEX), :n], "C", "Vector"))
    return out


def save(field, file_name, path="", add_num=True, num=-1, createpath=False, overwrite_time=None, les=True):
    """Returns `fname.xmf;` like the reference.  `field`: a ParticleField or a ResidentField (downloaded first).
    An empty field is saved as one dummy particle at the origin (:157-165)."""
    pf = field
    if hasattr(field, "download") and hasattr(field, "pfield"):
        field.download()
        pf = field.pfield
    n = pf.np
    P = np.asarray(pf.particles, dtype=np.float64)
    if n == 0:
        P = np.zeros((P.shape[0], 1))
        n = 1
    if createpath and path:
        os.makedirs(path, exist_ok=True)
    fname = file_name + ((f".{pf.nt}" if num == -1 else f".{num}") if add_num else "")
    binname = fname + ".bin"
    time = float(pf.t if overwrite_time is None else overwrite_time)
    items = []
    with open(os.path.join(path, binname), "wb") as fb:
        fb.write(struct.pack("<qqd", n, int(pf.nt), time))
        for name, a, attr, kind in _datasets(P, n, les):
            data = np.ascontiguousarray(a.T, dtype="<f8")     # [np, dim], row-major: XYZ per particle
            items.append((name, attr, kind, fb.tell(), data.shape[1]))
            fb.write(data.tobytes())

    def item(dim, seek):
        dims = f"{n} {dim}" if dim > 1 else f"{n}"
        return (f'\t\t\t\t\t<DataItem DataType="Float" Dimensions="{dims}" Format="Binary" Precision="8" '
                f'Endian="Little" Seek="{seek}">{binname}</DataItem>\n')

    with open(os.path.join(path, fname + ".xmf"), "w") as fx:
        fx.write('<?xml version="1.0" encoding="utf-8"?>\n')
        fx.write('<Xdmf xmlns:xi="http://www.w3.org/2001/XInclude" Version="3.0">\n')
        fx.write('\t<Domain>\n\t\t<Grid Name="particles" GridType="Uniform">\n')
        fx.write(f'\t\t\t\t<Time Value="{time}" />\n')
        for name, attr, kind, seek, dim in items:
            if name == "X":
                fx.write('\t\t\t\t<Geometry Type="XYZ">\n' + item(dim, seek) + '\t\t\t\t</Geometry>\n')
                fx.write(f'\t\t\t\t<Topology Dimensions="{n}" Type="Polyvertex"/>\n')
            else:
                fx.write(f'\t\t\t\t<Attribute Center="Node" Name="{attr}" Type="{kind}">\n' + item(dim, seek)
                         + '\t\t\t\t</Attribute>\n')
        fx.write('\t\t</Grid>\n\t</Domain>\n</Xdmf>\n')
    return fname + ".xmf;"


def read(xmf_path):
    """Read a file written by `save` back: dict(np, nt, t, X, Gamma, ..., each [np, dim]) -- what a restart
    (`read!` of the reference, src/FLOWVPM_utils.jl:423-470) or a test needs."""
    root = ET.parse(xmf_path).getroot()
    grid = root.find("Domain").find("Grid")
    here = os.path.dirname(os.path.abspath(xmf_path))
    out = {}

    def load(di):
        dims = [int(v) for v in di.get("Dimensions").split()]
        raw = np.fromfile(os.path.join(here, di.text.strip()), dtype="<f8", count=int(np.prod(dims)),
                          offset=int(di.get("Seek")))
        return raw.reshape(dims)

    out["X"] = load(grid.find("Geometry").find("DataItem"))
    for a in grid.findall("Attribute"):
        out[a.get("Name").replace(" ", "_")] = load(a.find("DataItem"))
    binfile = os.path.join(here, grid.find("Geometry").find("DataItem").text.strip())
    with open(binfile, "rb") as fb:
        out["np"], out["nt"], out["t"] = struct.unpack("<qqd", fb.read(24))
    assert float(grid.find("Time").get("Value")) == out["t"]
    return out
