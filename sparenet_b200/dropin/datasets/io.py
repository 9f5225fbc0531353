"""Drop-in for the reference's datasets/io.py (:15-80): IO.get / IO.put by file extension for the cloud formats on the path --
.pcd (ShapeNetCompletion, KITTI), .h5 (Completion3D), .npy, .txt; .png/.jpg through OpenCV like the reference.

The reference reads PCD through open3d (C++, not in this image): here the PCD container (v0.7: ASCII, binary and LZF
binary_compressed DATA sections; any FIELDS order, x/y/z extracted by name, F/I/U of size 1-8) is parsed directly with numpy --
no third-party dependency on the path.  `_read_pcd` returns float64 [n,3] like `np.array(pc.points)` (open3d stores doubles) and,
like open3d, drops points with a non-finite coordinate.  `_read_h5` multiplies by 0.9 like the reference (:63-65) and needs h5py
(absent here: the call raises ImportError; HDF5 is not re-implemented).  Parity note: open3d is un-vendored and absent, so the PCD
path is pinned to the format specification and to round trips, not to open3d's own output ("parity unpinned" for that dependency).
"""
import os
import struct

import numpy as np


def _lzf_decompress(src: bytes, out_len: int) -> bytes:
    """LZF (liblzf) as used by PCD's binary_compressed DATA: literal runs (ctrl < 32) and back references."""
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            ctrl += 1
            out[op:op + ctrl] = src[ip:ip + ctrl]
            ip += ctrl
            op += ctrl
        else:
            length = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if length == 7:
                length += src[ip]
                ip += 1
            ref -= src[ip]
            ip += 1
            for _ in range(length + 2):        # may overlap: byte by byte
                out[op] = out[ref]
                op += 1
                ref += 1
    return bytes(out[:op])


_NP = {("F", 4): np.float32, ("F", 8): np.float64, ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32, ("I", 8): np.int64,
       ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32, ("U", 8): np.uint64}


def read_pcd(file_path):
    with open(file_path, "rb") as f:
        raw = f.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        header[key.upper()] = val.split()
        if key.upper() == "DATA":
            break
    fields = header["FIELDS"]
    sizes = [int(v) for v in header["SIZE"]]
    types = header["TYPE"]
    counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
    n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header.get("HEIGHT", ["1"])[0])
    kind = header["DATA"][0].lower()
    names, dts = [], []
    for f_, s, t, c in zip(fields, sizes, types, counts):
        for j in range(c):
            names.append(f_ if c == 1 else f"{f_}_{j}")
            dts.append(np.dtype(_NP[(t.upper(), s)]))
    if kind == "ascii":
        body = raw[pos:].decode("ascii", "replace").split()
        arr = np.array(body[:n * len(names)], dtype=np.float64).reshape(n, len(names))
        cols = {nm: arr[:, i] for i, nm in enumerate(names)}
    elif kind == "binary":
        rec = np.dtype([(nm, dt) for nm, dt in zip(names, dts)])
        arr = np.frombuffer(raw, dtype=rec, count=n, offset=pos)
        cols = {nm: arr[nm] for nm in names}
    elif kind == "binary_compressed":
        csize, usize = struct.unpack_from("<II", raw, pos)
        data = _lzf_decompress(raw[pos + 8:pos + 8 + csize], usize)
        cols, off = {}, 0
        for nm, dt in zip(names, dts):         # field-major (structure of arrays) inside the compressed block
            cols[nm] = np.frombuffer(data, dtype=dt, count=n, offset=off)
            off += n * dt.itemsize
    else:
        raise ValueError(f"unsupported PCD DATA kind: {kind}")
    pts = np.stack([np.asarray(cols[c], dtype=np.float64) for c in ("x", "y", "z")], axis=1)
    return pts[np.isfinite(pts).all(axis=1)]


def write_pcd(file_path, points, binary=True):
    pts = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    n = pts.shape[0]
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
            f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'binary' if binary else 'ascii'}\n")
    with open(file_path, "wb") as f:
        f.write(head.encode("ascii"))
        if binary:
            f.write(pts.tobytes())
        else:
            f.write("".join(f"{p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n" for p in pts).encode("ascii"))


class IO:
    @classmethod
    def get(cls, file_path):
        ext = os.path.splitext(file_path)[1]
        if ext in (".png", ".jpg"):
            return cls._read_img(file_path)
        if ext == ".npy":
            return cls._read_npy(file_path)
        if ext == ".pcd":
            return cls._read_pcd(file_path)
        if ext == ".h5":
            return cls._read_h5(file_path)
        if ext == ".txt":
            return cls._read_txt(file_path)
        raise Exception("Unsupported file extension: %s" % ext)

    @classmethod
    def put(cls, file_path, file_content):
        ext = os.path.splitext(file_path)[1]
        if ext == ".pcd":
            return cls._write_pcd(file_path, file_content)
        if ext == ".h5":
            return cls._write_h5(file_path, file_content)
        raise Exception("Unsupported file extension: %s" % ext)

    @classmethod
    def _read_img(cls, file_path):
        import cv2
        return cv2.imread(file_path, cv2.IMREAD_UNCHANGED) / 255.0

    @classmethod
    def _read_npy(cls, file_path):
        return np.load(file_path)

    @classmethod
    def _read_pcd(cls, file_path):
        return read_pcd(file_path)

    @classmethod
    def _read_h5(cls, file_path):
        import h5py                          # not in this image: ImportError, loudly (HDF5 is not re-implemented)
        with h5py.File(file_path, "r") as f:
            return f["data"][()] * 0.9       # "avoid overflow while gridding" (:63-65)

    @classmethod
    def _read_txt(cls, file_path):
        return np.loadtxt(file_path)

    @classmethod
    def _write_pcd(cls, file_path, file_content):
        write_pcd(file_path, file_content)

    @classmethod
    def _write_h5(cls, file_path, file_content):
        import h5py
        with h5py.File(file_path, "w") as f:
            f.create_dataset("data", data=file_content)
