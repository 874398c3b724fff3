"""``agx_np_sincosf`` (csrc/agx_common.cuh) reproduces numpy's float32 ``sin`` / ``cos`` bit for bit - the reference
evaluates latlon -> xyz in float32 numpy (generate/transforms.py:106-110) and EdgeDirection is ill-conditioned in
those bits.  Observable through the C ABI as the source records of ``agx_node_tables``: z = sin(lat), record[3] =
cos(lat), x = cos(lat) * cos(lon), y = cos(lat) * sin(lon) (float32 products, as numpy forms them).

Sweep: every 131st float32 bit pattern of [-2 pi, 2 pi] (16.6 M values, every binade down to the denormals), as
latitude and as longitude."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sweep() -> np.ndarray:
    top = np.float32(2 * np.pi).view(np.int32)
    bits = np.arange(0, int(top) + 1, 131, dtype=np.int64).astype(np.int32)
    pos = bits.view(np.float32)
    return np.concatenate([pos, -pos, np.array([np.pi / 2, np.pi, 2 * np.pi, 1e-30, 0.0], dtype=np.float32)])


def test_float32_sincos_bits_match_numpy():
    from anemoi_graphs_b200 import ops

    v = _sweep()
    rng = np.random.default_rng(0)
    other = rng.uniform(-2 * np.pi, 2 * np.pi, v.size).astype(np.float32)
    for lat, lon in ((v, other), (other, v)):
        x = np.stack([lat, lon], axis=1)
        rec = ops.NodeTables(torch.from_numpy(x).cuda()).xyzc.cpu().numpy()
        cl = np.cos(lat)
        np.testing.assert_array_equal(rec[:, 2].view(np.int32), np.sin(lat).view(np.int32))
        np.testing.assert_array_equal(rec[:, 3].view(np.int32), cl.view(np.int32))
        np.testing.assert_array_equal(rec[:, 0].view(np.int32), (cl * np.cos(lon)).view(np.int32))
        np.testing.assert_array_equal(rec[:, 1].view(np.int32), (cl * np.sin(lon)).view(np.int32))
