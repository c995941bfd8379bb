"""TDMW weight container: the sidecar file that replaces the frozen TorchScript `model.pt`.

The reference ships its network as a frozen TorchScript archive with zero named parameters
(cva_mvsnet/export_model.py:194-197), so a from-scratch runtime needs the raw tensors.  A `.tdmw`
file holds the *unfolded* fp32 parameters of one CVA-MVSNet checkpoint (conv weights + BatchNorm
gamma/beta/mean/var) under their `state_dict` names (prefix `cva_mvsnet.` stripped), plus the two
hyper-parameters that change the graph: `depth_num` and `view_aggregation`.

Layout (little endian):
    char[8]  magic  "TDMW0001"
    u32      n_tensors
    u32      depth_num[3]
    u32      view_aggregation
    repeat n_tensors:  u32 name_len, char[name_len] name, u32 ndim, u32 dims[ndim], u64 offset (in floats)
    u64      n_floats
    f32[n_floats] data
The C++ loader (`tandem_b200/csrc/weights.cpp`) reads the same layout and folds BatchNorm at load time.
"""
import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"TDMW0001"


def save_tdmw(path, tensors, depth_num, view_aggregation):
    """tensors: ordered mapping name -> np.ndarray(float32)."""
    hdr = bytearray()
    hdr += MAGIC
    hdr += struct.pack("<I", len(tensors))
    hdr += struct.pack("<III", *[int(d) for d in depth_num])
    hdr += struct.pack("<I", 1 if view_aggregation else 0)
    off = 0
    blobs = []
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype=np.float32)
        nb = name.encode()
        hdr += struct.pack("<I", len(nb)) + nb
        hdr += struct.pack("<I", a.ndim)
        for d in a.shape:
            hdr += struct.pack("<I", d)
        hdr += struct.pack("<Q", off)
        off += a.size
        blobs.append(a.reshape(-1))
    hdr += struct.pack("<Q", off)
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.concatenate(blobs).astype("<f4").tobytes())


def load_tdmw(path):
    """Returns (OrderedDict name -> np.ndarray float32, depth_num tuple, view_aggregation bool)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a TDMW file")
    p = 8
    (n,) = struct.unpack_from("<I", buf, p); p += 4
    depth_num = struct.unpack_from("<III", buf, p); p += 12
    (va,) = struct.unpack_from("<I", buf, p); p += 4
    metas = []
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, p); p += 4
        name = buf[p:p + ln].decode(); p += ln
        (nd,) = struct.unpack_from("<I", buf, p); p += 4
        dims = struct.unpack_from("<" + "I" * nd, buf, p); p += 4 * nd
        (off,) = struct.unpack_from("<Q", buf, p); p += 8
        metas.append((name, dims, off))
    (nf,) = struct.unpack_from("<Q", buf, p); p += 8
    data = np.frombuffer(buf, dtype="<f4", count=nf, offset=p)
    out = OrderedDict()
    for name, dims, off in metas:
        cnt = int(np.prod(dims)) if len(dims) else 1
        out[name] = data[off:off + cnt].reshape(dims).copy()
    return out, tuple(int(d) for d in depth_num), bool(va)
