"""CPU tests of the host-side logic that needs no GPU: spec geometry, the tf-shaped adapter's graph bookkeeping,
flat-buffer bucketing and sharding."""
import numpy as np
import pytest
import torch

from imitation_from_observation_b200 import _lib, compat as tf, parallel, spec as S


def test_decoder_mirrors_encoder():
    s = S.context_skip_new(64, 64)
    d = s.dec_layers()
    assert [L["name"] for L in d] == ["d_h1", "d_h2", "d_h3", "d_h4"]
    assert d[0]["in_hw"] == (4, 4) and d[0]["out_hw"] == (8, 8) and d[0]["c_state"] + d[0]["c_skip"] == 1024
    assert d[3]["out_hw"] == (64, 64) and d[3]["c_out"] == 3 and d[3]["c_state"] + d[3]["c_skip"] == 128
    shapes = s.param_shapes()
    assert shapes["deconv/d_h1/w"] == (5, 5, 256, 1024) and shapes["conv/h0_conv/w"] == (5, 5, 3, 64)
    assert shapes["conv/h4_lin/Matrix"] == (8192, 1024) and shapes["deconv/d_h0_lin/Matrix"] == (1024, 8192)
    assert "conv_context/h0_conv/w" in shapes
    assert "conv_context/h0_conv/w" not in S.context_ae_real().param_shapes()
    r = S.context_ae_real()
    dr = r.dec_layers()
    assert [(L["stride"], L["in_hw"], L["out_hw"]) for L in dr] == [
        (2, (9, 16), (18, 32)), (1, (18, 32), (18, 32)), (2, (18, 32), (36, 64)), (1, (36, 64), (36, 64))]


def test_placeholder_chain_validation():
    ph = tf.placeholder(tf.uint8, (3, 2, 16, 16, 3))
    good = tf.multiply(tf.subtract(tf.image.convert_image_dtype(ph, dtype=tf.float32), 0.5), 2.0)
    assert good.root is ph and len(good.chain) == 3

    class Dummy:
        pass
    good._register_consumer(Dummy())
    assert len(ph.consumers) == 1
    with pytest.raises(_lib.CtxError):
        tf.subtract(ph, 0.5)._register_consumer(Dummy())
    fl = tf.placeholder(tf.float32, (3, 2, 16, 16, 3))
    with pytest.raises(_lib.CtxError):
        tf.multiply(fl, 2.0)._register_consumer(Dummy())
    with tf.variable_scope("contextmodel"):
        assert tf.current_scope() == "contextmodel/"
    assert tf.current_scope() == ""


def _offsets(spec):
    off, o = {}, 0
    for n, shp in spec.param_shapes().items():
        off[n] = o
        o += -(-int(np.prod(shp)) // 64) * 64
    return off, o


@pytest.mark.parametrize("spec", [S.context_skip_new(16, 16, 3, 4, 16), S.context_ae_real(12, 20, 3, 10, (8, 4, 4, 4))])
def test_scope_ranges_cover_flat_buffer(spec):
    off, total = _offsets(spec)
    ranges = parallel.scope_ranges(off, total)
    assert ranges[0][1] == 0 and ranges[-1][2] == total
    for (_, a, b), (_, c, d) in zip(ranges, ranges[1:]):
        assert b == c and a < b
    assert [r[0] for r in ranges] == [s for s in parallel.SCOPE_ORDER if any(k.startswith(s + "/") for k in off)]
    flat = torch.zeros(total)
    ar = parallel.BucketedAllReduce(flat, ranges)
    assert sum(b - a for _, a, b in ar.buckets) == total and not ar.enabled


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 250):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_transform_and_bank_layout():
    """arm_shaping.py:12-16 / train_script.py:16-22,67-97: resize + x/127.5-1, its inverse, and the time-major bank."""
    import numpy as np
    from imitation_from_observation_b200 import data
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (96, 128, 3)).astype(np.uint8)
    t = data.transform(img, 48, 64, True)
    assert t.shape == (48, 64, 3) and t.dtype == np.float64 and -1.0 <= t.min() and t.max() <= 1.0
    assert np.array_equal(data.transform(img, 96, 128, False), img)              # same size, no rescale: identity
    assert np.allclose(data.inverse_transform(t) * 255.0, (t + 1) * 127.5)
    vids = [[rng.randint(1, 256, (20, 20, 3)).astype(np.uint8) for _ in range(51)] for _ in range(3)] + [[img] * 40]
    bank = data.frames_to_bank(vids, nlen=25, nskip=2, idims=(16, 16))
    assert bank.shape == (25, 3, 16, 16, 3)                                        # the 40-frame video is skipped


def test_pixel_table_is_the_reference_transform():
    """DeviceBank's 256-entry table == float32(transform(v)) for every pixel value (arm_shaping.py:12-14)."""
    import numpy as np
    from imitation_from_observation_b200 import data, synthetic
    v = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    assert np.array_equal(data.pixel_table()[v], data.transform(v, 16, 16, True).astype(np.float32))
    assert np.array_equal(data.pixel_table()[v], synthetic.to_float(v).astype(np.float32))
