"""CMLW container: a flat list of named n-d arrays (little endian, C order).

The on-disk window-snapshot / golden-vector format of this repo.
  "CMLW0001" | u32 n | n x { u32 name_len | name | u32 dtype | u32 ndim | u64 dims[ndim] | data }
  dtype: 0=f32 1=f64 2=i32 3=u8 4=i64
The C++ twin used by the reference driver is oracle/cmlw_io.h.
"""
import struct
import numpy as np

_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f8"), 2: np.dtype("<i4"), 3: np.dtype("u1"), 4: np.dtype("<i8")}
_CODES = {v: k for k, v in _DTYPES.items()}
MAGIC = b"CMLW0001"


def save(path, arrays):
    """arrays: dict name -> array-like (f32/f64/i32/u8/i64)."""
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if dt not in _CODES:
                raise TypeError(f"cmlw: unsupported dtype {a.dtype} for {name}")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", _CODES[dt], a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(a.astype(dt, copy=False).tobytes())


def load(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a CMLW file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            code, nd = struct.unpack("<II", f.read(8))
            dims = struct.unpack(f"<{nd}Q", f.read(8 * nd)) if nd else ()
            dt = _DTYPES[code]
            cnt = int(np.prod(dims)) if nd else 1
            data = f.read(cnt * dt.itemsize)
            out[name] = np.frombuffer(data, dtype=dt).reshape(dims).copy()
    return out
