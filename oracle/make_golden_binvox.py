"""Pins oracle/binvox.py against the real reference reader / writer and writes tests/golden/binvox_cases.npz.
Runs ONLY in the build container (needs /root/reference):  python -m oracle.make_golden_binvox"""
import io
import os
import sys
import types

import numpy as np

from . import binvox as OB

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "binvox_cases.npz")


def import_reference():
    # binvox_rw imports matplotlib at module level for a plotting helper only
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.path.insert(0, "/root/reference/utils")
    import binvox_rw
    return binvox_rw


def cases():
    rng = np.random.default_rng(0)
    out = {}
    out["empty"] = np.zeros((32, 32, 32), bool)
    out["full"] = np.ones((32, 32, 32), bool)
    out["sparse"] = rng.random((32, 32, 32)) < 0.03
    out["dense_noise"] = rng.random((32, 32, 32)) < 0.5
    blob = np.zeros((32, 32, 32), bool)
    blob[6:25, 9:20, 4:29] = True
    blob[10:14, 12:15, 0:32] = False
    out["blob"] = blob
    run255 = np.zeros(32 * 32 * 32, bool)      # runs that are exact multiples of 255 (the writer's zero-length pairs)
    run255[255:510] = True
    run255[510 + 765:510 + 765 + 255] = True
    run255[-255:] = True
    out["run255"] = run255.reshape(32, 32, 32)
    out["small_rect"] = rng.random((5, 7, 3)) < 0.4
    out["alternating"] = (np.arange(16 ** 3) % 2 == 0).reshape(16, 16, 16)
    return out


def main():
    ref = import_reference()
    store = {}
    for name, vol in cases().items():
        dims = list(vol.shape)
        for order in ("xyz", "xzy"):
            fp = io.BytesIO()
            ref.write(ref.Voxels(vol, dims, [0.0, 0.0, 0.0], 1.0, order), fp)
            blob = fp.getvalue()
            d2, tr, sc, off = OB.read_header(blob)
            assert d2 == dims
            assert OB.header_bytes(dims, [0.0, 0.0, 0.0], 1.0) == blob[:off]
            assert OB.encode(vol, order) == blob[off:], (name, order)
            for fix in (True, False):
                got_ref = ref.read_as_3d_array(io.BytesIO(blob), fix_coords=fix).data
                assert np.array_equal(OB.decode(blob[off:], d2, fix), got_ref), (name, order, fix)
            if order == "xyz" and len(set(dims)) == 1:
                # written from an xyz volume, read back with fix_coords -> the same volume (cubic grids only: the
                # reference reshapes the transposed stream with the untransposed dims, binvox_rw.py:144-147)
                assert np.array_equal(ref.read_as_3d_array(io.BytesIO(blob)).data, vol)
            store[f"{name}.{order}.file"] = np.frombuffer(blob, dtype=np.uint8)
        store[f"{name}.volume"] = vol
    np.savez_compressed(GOLDEN, **store)
    print("oracle.binvox == reference utils/binvox_rw.py on", len(cases()), "volumes x 2 axis orders; wrote", GOLDEN)


if __name__ == "__main__":
    main()
