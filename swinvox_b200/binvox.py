"""binvox run-length streams <-> device volumes: the ground-truth side of the IoU step.

Drop-in for the two functions of the reference's utils/binvox_rw.py that sit on the evaluation path
(read_as_3d_array, utils/binvox_rw.py:119-153, as called by utils/data_loaders.py:84-87; write,
utils/binvox_rw.py:239-300) -- batched, decoded straight into the fp32 {0,1} device tensor the metric kernel
reads.  The text header is parsed on the host (five short lines); the run-length payload goes through the CUDA
kernels of libswinvox_b200 (svx_binvox_decode / svx_binvox_encode).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Voxels:
    """the reference's container (utils/binvox_rw.py:67-103): data, dims, translate, scale, axis_order"""

    def __init__(self, data, dims, translate, scale, axis_order):
        assert axis_order in ("xzy", "xyz")
        self.data, self.dims, self.translate, self.scale, self.axis_order = data, dims, translate, scale, axis_order


def read_header(buf):
    """utils/binvox_rw.py:106-116 on a bytes object -> (dims, translate, scale, payload offset)"""
    pos, lines = 0, []
    for _ in range(5):
        end = buf.find(b"\n", pos)
        if end < 0:
            raise IOError("[ERROR] Not a binvox file")
        lines.append(buf[pos:end].strip())
        pos = end + 1
    if not lines[0].startswith(b"#binvox"):
        raise IOError("[ERROR] Not a binvox file")
    dims = [int(v) for v in lines[1].split(b" ")[1:]]
    translate = [float(v) for v in lines[2].split(b" ")[1:]]
    scale = [float(v) for v in lines[3].split(b" ")[1:]][0]
    return dims, translate, scale, pos


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def decode_batch(files, device="cuda", fix_coords=True, out=None):
    """files: list of bytes objects (whole .binvox files, all with the same dims).  Returns (volumes, headers):
    volumes = fp32 {0,1} tensor [B, d0, d2, d1] on `device` (x, y, z order when fix_coords, as the data loader feeds
    the metric), headers = [(dims, translate, scale)].  One H2D copy of the concatenated payloads, one kernel."""
    lib = _lib.get()
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.SvxError("swinvox_b200.binvox decodes on the GPU only (no CPU fallback)")
    heads, chunks, offsets, pos = [], [], [0], 0
    for f in files:
        dims, tr, sc, off = read_header(f)
        heads.append((dims, tr, sc))
        pay = np.frombuffer(f, dtype=np.uint8, offset=off)
        if pay.size % 2:
            pay = pay[:-1]          # a trailing unpaired byte has no count (raw_data[::2] / [1::2] would not pair it either)
        total = int(pay[1::2].astype(np.int64).sum())
        if dims != heads[0][0]:
            raise ValueError("decode_batch needs volumes of one size per call")
        if total != dims[0] * dims[1] * dims[2]:
            raise ValueError(f"cannot reshape array of size {total} into shape {tuple(dims)}")   # numpy's reshape error
        chunks.append(pay)
        pos += pay.size
        offsets.append(pos)
    d0, d1, d2 = heads[0][0]
    B = len(files)
    payload = torch.from_numpy(np.concatenate(chunks)).to(device, non_blocking=True)
    offs = torch.tensor(offsets, dtype=torch.int64).to(device, non_blocking=True)
    shape = (B, d0, d2, d1) if fix_coords else (B, d0, d1, d2)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("decode_batch: `out` must be a contiguous fp32 tensor of shape %s" % (shape,))
    status = torch.empty(B, dtype=torch.int32, device=device)
    d = _lib.BinvoxDecodeDesc()
    d.payload, d.offsets, d.out, d.status = payload.data_ptr(), offs.data_ptr(), out.data_ptr(), status.data_ptr()
    d.B, d.d0, d.d1, d.d2, d.fix_coords = B, d0, d1, d2, 1 if fix_coords else 0
    _lib.check(lib.svx_binvox_decode(C.byref(d), _stream(device)), lib)
    payload.record_stream(torch.cuda.current_stream(device))
    return out, heads


def read_as_3d_array(fp, fix_coords=True, device="cuda"):
    """the reference's call (utils/binvox_rw.py:119): one file object -> Voxels whose .data is a bool device tensor"""
    vol, heads = decode_batch([fp.read()], device, fix_coords)
    dims, tr, sc = heads[0]
    return Voxels(vol[0] > 0, dims, tr, sc, "xyz" if fix_coords else "xzy")


def encode_batch(volumes, threshold=0.5, axis_order="xyz", translate=(0.0, 0.0, 0.0), scale=1.0):
    """volumes: fp32 device tensor [B, d0, d1, d2] (probabilities, or {0,1}); a voxel is set iff value >= threshold.
    Returns a list of B bytes objects: complete .binvox files, byte-identical to the reference writer's output."""
    lib = _lib.get()
    if volumes.device.type != "cuda":
        raise _lib.SvxError("swinvox_b200.binvox encodes on the GPU only (no CPU fallback)")
    if axis_order not in ("xzy", "xyz"):
        raise ValueError("[ERROR] Unsupported voxel model axis order")
    v = volumes.to(torch.float32).contiguous()
    B, d0, d1, d2 = v.shape
    P = d0 * d1 * d2
    payload = torch.empty(B, 2 * P, dtype=torch.uint8, device=v.device)
    nbytes = torch.empty(B, dtype=torch.int32, device=v.device)
    d = _lib.BinvoxEncodeDesc()
    d.volume, d.threshold, d.payload, d.nbytes = v.data_ptr(), float(threshold), payload.data_ptr(), nbytes.data_ptr()
    d.B, d.d0, d.d1, d.d2, d.axis_xyz = B, d0, d1, d2, 1 if axis_order == "xyz" else 0
    _lib.check(lib.svx_binvox_encode(C.byref(d), _stream(v.device)), lib)
    n = nbytes.cpu().tolist()
    pay = payload.cpu().numpy()
    head = ("#binvox 1\n" + "dim %s\n" % " ".join(map(str, (d0, d1, d2))) + "translate %s\n" % " ".join(map(str, translate))
            + "scale %s\n" % str(scale) + "data\n").encode("latin-1")
    return [head + pay[b, :n[b]].tobytes() for b in range(B)]
