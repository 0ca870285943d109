"""binvox run-length decode / encode.  CPU tier: the oracle against the golden files written by the
REAL reference writer (oracle/make_golden_binvox.py).  GPU tier: the CUDA kernels, through the C-ABI, bit-exact
against the oracle and the goldens (byte/index work: no tolerance)."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import binvox as OB

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "binvox_cases.npz")
CASES = ["empty", "full", "sparse", "dense_noise", "blob", "run255", "small_rect", "alternating"]


def load():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_files(name):
    g = load()
    vol = g[f"{name}.volume"]
    for order in ("xyz", "xzy"):
        blob = g[f"{name}.{order}.file"].tobytes()
        dims, tr, sc, off = OB.read_header(blob)
        assert dims == list(vol.shape) and tr == [0.0, 0.0, 0.0] and sc == 1.0
        assert OB.header_bytes(dims, tr, sc) + OB.encode(vol, order) == blob
        if len(set(dims)) == 1:
            back = OB.decode(blob[off:], dims, fix_coords=(order == "xyz"))
            assert np.array_equal(back, vol)


def test_oracle_zero_length_pair_quirk():
    """a run of exactly 255 followed by a change is written as (v,255)(v,0): utils/binvox_rw.py:283-294"""
    flat = np.zeros(4 ** 3, bool)
    v = np.ones((8, 8, 8), bool)
    v.reshape(-1)[255:] = False
    enc = OB.encode(v, "xzy")
    assert enc[:6] == bytes((1, 255, 1, 0, 0, 255)) and enc[-2:] == bytes((0, 2))
    assert np.array_equal(OB.decode(enc, [8, 8, 8], fix_coords=False), v) and flat.sum() == 0


@pytest.mark.gpu
def test_decode_kernel_matches_goldens_bit_exact():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200 import binvox as BV
    g = load()
    cubes = [n for n in CASES if g[f"{n}.volume"].shape == (32, 32, 32)]
    files = [g[f"{n}.xyz.file"].tobytes() for n in cubes]
    vols, heads = BV.decode_batch(files)               # one launch for the whole batch
    torch.cuda.synchronize()
    assert vols.dtype == torch.float32 and tuple(vols.shape) == (len(cubes), 32, 32, 32)
    for i, n in enumerate(cubes):
        assert np.array_equal(vols[i].cpu().numpy(), g[f"{n}.volume"].astype(np.float32)), n
        assert heads[i] == ([32, 32, 32], [0.0, 0.0, 0.0], 1.0)
    # raw file order (fix_coords=False) and non-cubic dims, one object at a time, against the oracle
    for n in CASES:
        for order in ("xyz", "xzy"):
            blob = g[f"{n}.{order}.file"].tobytes()
            dims, _, _, off = OB.read_header(blob)
            for fix in (True, False):
                got = BV.decode_batch([blob], fix_coords=fix)[0][0].cpu().numpy()
                assert np.array_equal(got, OB.decode(blob[off:], dims, fix).astype(np.float32)), (n, order, fix)
    v = BV.read_as_3d_array(io.BytesIO(files[2]))      # the reference's call signature
    assert v.axis_order == "xyz" and v.dims == [32, 32, 32] and np.array_equal(v.data.cpu().numpy(), g["sparse.volume"])


@pytest.mark.gpu
def test_encode_kernel_is_byte_identical_to_reference_writer():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200 import binvox as BV
    g = load()
    for n in CASES:
        vol = torch.from_numpy(g[f"{n}.volume"].astype(np.float32)).cuda()[None]
        for order in ("xyz", "xzy"):
            out = BV.encode_batch(vol, threshold=0.5, axis_order=order)
            assert out[0] == g[f"{n}.{order}.file"].tobytes(), (n, order)
    # batched, thresholding probabilities, random volumes against the oracle; then the round trip through decode
    rng = np.random.default_rng(5)
    probs = rng.random((16, 32, 32, 32)).astype(np.float32)
    probs[3] = 0.0
    probs[4, :, :, :16] = 0.9
    files = BV.encode_batch(torch.from_numpy(probs).cuda(), threshold=0.3)
    for b in range(16):
        want = OB.header_bytes([32, 32, 32], (0.0, 0.0, 0.0), 1.0) + OB.encode(probs[b] >= 0.3, "xyz")
        assert files[b] == want, b
    back, _ = BV.decode_batch(files)
    assert np.array_equal(back.cpu().numpy(), (probs >= 0.3).astype(np.float32))


@pytest.mark.gpu
def test_decode_rejects_malformed_streams():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200 import binvox as BV
    blob = load()["blob.xyz.file"].tobytes()
    with pytest.raises(ValueError, match="cannot reshape"):
        BV.decode_batch([blob[:-2]])
    with pytest.raises(IOError):
        BV.decode_batch([b"#notbinvox\n" + blob])
