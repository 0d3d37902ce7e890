"""The .pcd wire format of the SEE-VCN frame driver (host side, numpy only).

ref: SEE_VCN.save_pcd (see/surface_completion/SEE_VCN.py:267-280) writes the completed frame with
``o3d.io.write_point_cloud(fname, pcd, write_ascii=False)``; sc_multiproc.py:21-23 and the dataset adapters read it back
with ``o3d.io.read_point_cloud``.  open3d is not a dependency here: a binary PCD v0.7 with ``FIELDS x y z`` float32 is
a fixed ASCII header (the one of demo/demo_data/pcd/000001.pcd, kept verbatim in tests/golden/pcd_header.txt) followed
by N * 12 bytes.  The reader also accepts extra float32 fields (e.g. intensity) and the ascii flavour.
"""
import numpy as np

_HEADER = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
           "WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")


def pcd_header(num_points):
    return _HEADER.format(n=int(num_points))


def write_pcd(path, points):
    """points (N, >=3) -> binary PCD with the xyz columns as float32 (what save_pcd stores: '(N,3) shape' only)."""
    xyz = np.ascontiguousarray(np.asarray(points)[:, :3], dtype="<f4")
    with open(path, "wb") as f:
        f.write(pcd_header(len(xyz)).encode("ascii"))
        f.write(xyz.tobytes())


def read_pcd(path, fields=("x", "y", "z")):
    """-> (N, len(fields)) float32.  Binary or ascii PCD whose requested fields are 4-byte floats."""
    with open(path, "rb") as f:
        meta = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no DATA line")
            text = line.decode("ascii", "replace").strip()
            if not text or text.startswith("#"):
                continue
            key, _, rest = text.partition(" ")
            meta[key] = rest.split()
            if key == "DATA":
                break
        names = meta["FIELDS"]
        sizes = [int(v) for v in meta["SIZE"]]
        types = meta["TYPE"]
        counts = [int(v) for v in meta.get("COUNT", ["1"] * len(names))]
        n = int(meta["POINTS"][0]) if "POINTS" in meta else int(meta["WIDTH"][0]) * int(meta.get("HEIGHT", ["1"])[0])
        for name in fields:
            i = names.index(name)
            if not (sizes[i] == 4 and types[i] == "F" and counts[i] == 1):
                raise ValueError(f"{path}: field {name} is not a float32 scalar")
        kind = meta["DATA"][0]
        if kind == "binary":
            dt = np.dtype({"names": [f"f{i}" for i in range(len(names))],
                           "formats": [np.dtype((_np_type(types[i], sizes[i]), (counts[i],))) if counts[i] > 1
                                       else _np_type(types[i], sizes[i]) for i in range(len(names))]})
            raw = f.read(n * dt.itemsize)
            if len(raw) < n * dt.itemsize:
                raise ValueError(f"{path}: truncated ({len(raw)} of {n * dt.itemsize} bytes)")
            rec = np.frombuffer(raw, dtype=dt, count=n)
            return np.stack([rec[f"f{names.index(name)}"] for name in fields], axis=1).astype(np.float32)
        if kind == "ascii":
            table = np.loadtxt(f, dtype=np.float64, ndmin=2)
            cols = np.cumsum([0] + counts[:-1])
            return table[:n, [cols[names.index(name)] for name in fields]].astype(np.float32)
        raise ValueError(f"{path}: DATA {kind} is not supported (binary_compressed needs LZF)")


def _np_type(t, size):
    return {"F": {4: "<f4", 8: "<f8"}, "I": {1: "i1", 2: "<i2", 4: "<i4", 8: "<i8"},
            "U": {1: "u1", 2: "<u2", 4: "<u4", 8: "<u8"}}[t][size]
