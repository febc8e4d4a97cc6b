"""Dependency-free reader for the two VTK ImageData (.vti) flavours on the LBMCL path.

* what the reference (and this build's host) WRITES: inline ASCII ``DataArray`` bodies,
  ``header_type="UInt64"`` (reference lbmcl.hpp:261-334, ``storeData``);
* what the reference's golden fixtures ARE: Sailfish output, ``format="appended"``,
  base64 + zlib, ``header_type="UInt32"`` (reference target_results/8, target_results/32).

The reference's verify.py:44-52 reads both through ``pyvista.read(...).point_arrays``; pyvista and
vtk are not available here, so this module restates the small part of the VTK XML
format that is needed.  Returned arrays are in file order (x fastest, then y, then z) with the
file's own dtype, shaped ``(n_points,)`` or ``(n_points, n_components)``.
"""
from __future__ import annotations

import base64
import re
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

_VTK_DTYPES = {
    "Float32": np.float32,
    "Float64": np.float64,
    "Int32": np.int32,
    "UInt32": np.uint32,
    "Int64": np.int64,
    "UInt64": np.uint64,
    "UInt8": np.uint8,
    "Int8": np.int8,
}


def _b64_len(nbytes: int) -> int:
    return (nbytes + 2) // 3 * 4


def _decode_appended_array(blob: str, offset: int, header_dtype, compressed: bool, dtype) -> np.ndarray:
    """Decode one base64 ``AppendedData`` array that starts ``offset`` characters after '_'."""
    hsize = np.dtype(header_dtype).itemsize
    if not compressed:
        n = int(np.frombuffer(base64.b64decode(blob[offset:offset + _b64_len(hsize)]), header_dtype)[0])
        raw = base64.b64decode(blob[offset:offset + _b64_len(hsize + n)])[hsize:hsize + n]
        return np.frombuffer(raw, dtype=dtype).copy()
    # compressed: header = [n_blocks, block_size, last_block_size, csize_0 .. csize_{n-1}],
    # base64-encoded on its own, followed by the base64 of the concatenated zlib blocks.
    first = base64.b64decode(blob[offset:offset + _b64_len(3 * hsize)])
    n_blocks = int(np.frombuffer(first[:hsize], header_dtype)[0])
    hbytes = (3 + n_blocks) * hsize
    hchars = _b64_len(hbytes)
    header = np.frombuffer(base64.b64decode(blob[offset:offset + hchars])[:hbytes], header_dtype)
    csizes = [int(c) for c in header[3:3 + n_blocks]]
    total = sum(csizes)
    data = base64.b64decode(blob[offset + hchars:offset + hchars + _b64_len(total)])
    out = bytearray()
    pos = 0
    for c in csizes:
        out += zlib.decompress(data[pos:pos + c])
        pos += c
    return np.frombuffer(bytes(out), dtype=dtype).copy()


def read_vti(path: str) -> dict:
    """Return ``{"extent": (x0,x1,y0,y1,z0,z1), "dims": (nx,ny,nz), "arrays": {name: ndarray}}``."""
    with open(path, "r") as fh:
        text = fh.read()

    # The appended blob may contain characters that upset an XML parser only in theory (base64 is
    # XML safe), but cutting it out keeps ElementTree fast on the 32^3 fixtures.
    blob = None
    m = re.search(r"<AppendedData[^>]*>\s*_", text)
    if m is not None:
        end = text.index("</AppendedData>", m.end())
        blob = "".join(text[m.end():end].split())
        enc = re.search(r'encoding="([^"]+)"', m.group(0))
        if enc is None or enc.group(1) != "base64":
            raise ValueError(f"{path}: only base64 AppendedData is supported")
        text = text[:m.start()] + "<AppendedData/>" + text[end + len("</AppendedData>"):]

    root = ET.fromstring(text)
    if root.tag != "VTKFile" or root.attrib.get("type") != "ImageData":
        raise ValueError(f"{path}: not a VTK ImageData file")
    if root.attrib.get("byte_order", "LittleEndian") != "LittleEndian":
        raise ValueError(f"{path}: only little endian files are supported")
    header_dtype = _VTK_DTYPES[root.attrib.get("header_type", "UInt32")]
    compressed = "compressor" in root.attrib

    image = root.find("ImageData")
    piece = image.find("Piece")
    extent = tuple(int(v) for v in piece.attrib["Extent"].split())
    dims = (extent[1] - extent[0] + 1, extent[3] - extent[2] + 1, extent[5] - extent[4] + 1)
    n_points = dims[0] * dims[1] * dims[2]

    arrays = {}
    for da in piece.find("PointData").findall("DataArray"):
        name = da.attrib["Name"]
        dtype = _VTK_DTYPES[da.attrib["type"]]
        ncomp = int(da.attrib.get("NumberOfComponents", "1"))
        fmt = da.attrib.get("format", "ascii")
        if fmt == "ascii":
            # "nan"/"-nan"/"inf" as printed by iostreams are understood by float()
            vals = np.array([float(t) for t in (da.text or "").split()], dtype=np.float64).astype(dtype)
        elif fmt == "appended":
            if blob is None:
                raise ValueError(f"{path}: appended array without AppendedData")
            vals = _decode_appended_array(blob, int(da.attrib["offset"]), header_dtype, compressed, dtype)
        else:
            raise ValueError(f"{path}: unsupported DataArray format {fmt!r}")
        if vals.size != n_points * ncomp:
            raise ValueError(f"{path}: array {name!r} has {vals.size} values, expected {n_points * ncomp}")
        arrays[name] = vals.reshape(n_points, ncomp) if ncomp > 1 else vals
    return {"extent": extent, "dims": dims, "arrays": arrays}


def write_vti_ascii(path: str, rho: np.ndarray, v: np.ndarray, dim: int) -> None:
    """Write ``rho[N]`` / ``v[3][N]`` (N = dim^3, x fastest) the way reference lbmcl.hpp:261-334 does.

    Only the wet cube x,y,z in [1, dim-2] is written; values are printed in scientific notation with
    16 digits.  Used by the tests to produce files from CPU-side arrays; the product's writer is the C++
    host (lbmcl_b200/host/lbmb200.hpp).
    """
    n = dim ** 3
    rho = np.asarray(rho).reshape(dim, dim, dim)
    v = np.asarray(v).reshape(3, dim, dim, dim)
    assert rho.size == n
    type_str = "Float32" if rho.dtype == np.float32 else "Float64"
    ext = dim - 3
    sl = slice(1, dim - 1)

    def fmt(x) -> str:
        s = "%.16e" % float(x)
        return s  # Python prints nan / -nan as 'nan' / 'nan'; the reference prints 'nan' or '-nan'

    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n')
        fh.write('<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" header_type="UInt64">\n')
        fh.write(f'  <ImageData WholeExtent="0 {ext} 0 {ext} 0 {ext}" Origin="0 0 0" Spacing="1 1 1">\n')
        fh.write(f'    <Piece Extent="0 {ext} 0 {ext} 0 {ext}">\n')
        fh.write('      <PointData Scalars="rho">\n')
        fh.write(f'        <DataArray type="{type_str}" Name="rho" NumberOfComponents="1" format="ascii">\n')
        for z in range(1, dim - 1):
            for y in range(1, dim - 1):
                fh.write("".join(fmt(x) + " " for x in rho[z, y, sl]) + "\n")
        fh.write("        </DataArray>\n")
        fh.write(f'        <DataArray type="{type_str}" Name="v" NumberOfComponents="3" format="ascii">\n')
        for z in range(1, dim - 1):
            for y in range(1, dim - 1):
                row = []
                for x in range(1, dim - 1):
                    row.append(f"{fmt(v[0, z, y, x])} {fmt(v[1, z, y, x])} {fmt(v[2, z, y, x])} ")
                fh.write("".join(row) + "\n")
        fh.write("        </DataArray>\n      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n")
