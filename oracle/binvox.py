"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's binvox reader / writer (utils/binvox_rw.py),
the on-disk format of the ground-truth volumes that feed the IoU step (utils/data_loaders.py:84-87).

Pinned: oracle/make_golden_binvox.py imports the real /root/reference/utils/binvox_rw.py (behind a matplotlib stub,
the module only imports it for a plotting helper) and checks both functions byte-/bit-exact on the committed
fixtures tests/golden/binvox_cases.npz.
"""
import numpy as np


def read_header(buf):
    """utils/binvox_rw.py:106-116 -> (dims, translate, scale, offset of the RLE payload)"""
    pos, lines = 0, []
    for _ in range(5):
        end = buf.index(b"\n", pos)
        lines.append(buf[pos:end].strip())
        pos = end + 1
    if not lines[0].startswith(b"#binvox"):
        raise IOError("[ERROR] Not a binvox file")
    dims = [int(v) for v in lines[1].split(b" ")[1:]]
    translate = [float(v) for v in lines[2].split(b" ")[1:]]
    scale = [float(v) for v in lines[3].split(b" ")[1:]][0]
    return dims, translate, scale, pos


def decode(payload, dims, fix_coords=True):
    """utils/binvox_rw.py:119-153: (value, count) byte pairs -> dense bool volume; the file order is x, z, y
    (y fastest) and fix_coords transposes to x, y, z."""
    raw = np.frombuffer(payload, dtype=np.uint8)
    values, counts = raw[::2], raw[1::2]
    data = np.repeat(values, counts).astype(bool).reshape(dims)
    return np.transpose(data, (0, 2, 1)) if fix_coords else data


def encode(dense, axis_order="xyz"):
    """utils/binvox_rw.py:268-300 payload bytes of a dense volume, including the writer's quirk: a run whose length is
    a multiple of 255 is followed by a zero-length pair when the value then changes (ctr was reset to 0 by the dump),
    and a trailing remainder of 0 is not written."""
    d = np.asarray(dense).astype(int)
    if axis_order not in ("xzy", "xyz"):
        raise ValueError("[ERROR] Unsupported voxel model axis order")
    flat = d.flatten() if axis_order == "xzy" else np.transpose(d, (0, 2, 1)).flatten()
    out = bytearray()
    state, ctr = int(flat[0]), 0
    for c in flat.tolist():
        if c == state:
            ctr += 1
            if ctr == 255:
                out += bytes((state, ctr))
                ctr = 0
        else:
            out += bytes((state, ctr))
            state, ctr = c, 1
    if ctr > 0:
        out += bytes((state, ctr))
    return bytes(out)


def header_bytes(dims, translate=(0.0, 0.0, 0.0), scale=1.0):
    """utils/binvox_rw.py:253-262"""
    return ("#binvox 1\n" + "dim %s\n" % " ".join(map(str, dims)) + "translate %s\n" % " ".join(map(str, translate))
            + "scale %s\n" % str(scale) + "data\n").encode("latin-1")
