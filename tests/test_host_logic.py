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


def test_tf_checkpoint_round_trip_and_format_details(tmp_path):
    """tf_checkpoint.py (Saver V2 tensor bundles, written from the published formats — no TF-written file is available):
    its writer and reader agree, and the reader copes with what a TF-written index may contain that our writer never
    produces: prefix-compressed keys, several data blocks, Snappy-compressed blocks, extra variables (Adam slots)."""
    import struct
    import numpy as np
    import pytest
    from imitation_from_observation_b200 import tf_checkpoint as T

    # published check values: CRC-32C("123456789") = 0xE3069283; LevelDB's mask of it
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283
    assert T.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF

    rs = np.random.RandomState(0)
    sd = {"contextmodel/conv/h%d_conv/w" % j: rs.randn(5, 5, 3 + j, 4).astype(np.float32) for j in range(40)}
    sd["contextmodel/conv/h0_conv/biases"] = rs.randn(4).astype(np.float32)
    sd["global_step"] = np.array(7, dtype=np.int64)                          # a scalar
    sd["contextmodel/conv/h0_conv/w/Adam"] = rs.randn(5, 5, 3, 4).astype(np.float32)
    prefix = str(tmp_path / "model_10")
    T.write_checkpoint(prefix, sd)
    assert T.checkpoint_exists(prefix) and not T.checkpoint_exists(prefix + "x")
    got = T.read_checkpoint(prefix)
    assert sorted(got) == sorted(sd)
    for k in sd:
        assert got[k].dtype == sd[k].dtype and got[k].shape == sd[k].shape and np.array_equal(got[k], sd[k]), k
    assert T.list_variables(prefix)["global_step"] == (np.dtype(np.int64), ())
    only = T.read_checkpoint(prefix, ["contextmodel/conv/h3_conv/w"])
    assert list(only) == ["contextmodel/conv/h3_conv/w"]
    with pytest.raises(KeyError):
        T.read_checkpoint(prefix, ["nope"])
    # several data blocks were written (4 KB blocks, 43 entries of ~60 bytes each fit one; force more) and keys share prefixes
    T.write_table(str(tmp_path / "t"), [(b"key%04d" % i, b"v" * (i % 7)) for i in range(2000)], block_size=256)
    back = T.read_table(str(tmp_path / "t"))
    assert back == [(b"key%04d" % i, b"v" * (i % 7)) for i in range(2000)]
    # corruption is detected: flip one data byte -> tensor checksum; flip one index byte -> block checksum
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(T.CheckpointError):
        T.read_checkpoint(prefix)
    assert len(T.read_checkpoint(prefix, verify=False)) == len(sd)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 1
    open(str(tmp_path / "bad.index"), "wb").write(bytes(idx))
    with pytest.raises(T.CheckpointError):
        T.read_table(str(tmp_path / "bad.index"))

    # Snappy: literals of every length class, copies with 1- and 2-byte offsets, an overlapping (run-length) copy
    def lit(b):
        n = len(b) - 1
        if n < 60:
            return bytes([n << 2]) + b
        return bytes([60 << 2, n]) + b if n < 256 else bytes([61 << 2]) + struct.pack("<H", n) + b
    plain = b"abcdefgh" + b"abcdefgh"[:6] + b"x" * 70 + b"Z" * 300 + bytes(range(200)) + b"abcd"
    stream = (T._put_varint(len(plain)) + lit(b"abcdefgh") + bytes([((6 - 4) << 2) | 1, 8])          # copy len 6, offset 8
              + lit(b"x") + bytes([((64 - 1) << 2) | 2]) + struct.pack("<H", 1) + bytes([((5 - 1) << 2) | 2]) + struct.pack("<H", 1)
              + lit(b"Z" * 300) + lit(bytes(range(200)))
              + bytes([((4 - 1) << 2) | 2]) + struct.pack("<H", 300 + 200 + 70 + 6 + 8))                 # "abcd" from the very start
    assert T.snappy_decompress(stream) == plain
    # a table whose data block is stored Snappy-compressed (type byte 1), as TF's TableBuilder may write it
    payload = np.arange(6, dtype=np.float32).tobytes()
    block = T._build_block([(b"", b"\x08\x01"), (b"a/w", T._entry_bytes(1, (2, 3), 0, 24, T.mask_crc(T.crc32c(payload))))])
    comp = T._put_varint(len(block)) + lit(block)
    out = bytearray()

    def emit(contents, ctype):
        off = len(out)
        out.extend(contents + bytes([ctype]) + struct.pack("<I", T.mask_crc(T.crc32c(contents + bytes([ctype])))))
        return T._put_varint(off) + T._put_varint(len(contents))
    h = emit(comp, 1)
    meta = emit(T._build_block([]), 0)
    index = emit(T._build_block([(b"a/w", h)], 1), 0)
    footer = meta + index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC))
    open(str(tmp_path / "snap.index"), "wb").write(bytes(out))
    np.arange(6, dtype=np.float32).tofile(str(tmp_path / "snap.data-00000-of-00001"))
    got = T.read_checkpoint(str(tmp_path / "snap"), verify=True)
    assert list(got) == ["a/w"] and np.array_equal(got["a/w"], np.arange(6, dtype=np.float32).reshape(2, 3))


def test_saver_reads_and_writes_tf_checkpoints(tmp_path):
    """compat.Saver: save(format="tf") writes a tensor bundle, restore() finds it (before the .npz), accepts the
    checkpoint's own scope prefix (the trainer builds under "contextmodel/", base.py:134-138 does not) and ignores the
    Adam slot variables a TF-trained checkpoint also holds."""
    import numpy as np
    import pytest
    from imitation_from_observation_b200 import compat as tf, tf_checkpoint as T

    class Net:
        def __init__(self):
            self.shapes = {"conv/h0_conv/w": (5, 5, 3, 4), "conv/h0_conv/biases": (4,), "deconv/d_h0_lin/Matrix": (6, 8)}
            rs = np.random.RandomState(1)
            self.P = {k: rs.randn(*s).astype(np.float32) for k, s in self.shapes.items()}

        def state_dict(self):
            return {k: v.copy() for k, v in self.P.items()}

        def load_state_dict(self, sd, prefix=""):
            self.P = {k: np.asarray(sd[prefix + k], dtype=np.float32).copy() for k in self.shapes}

    class Model:
        def __init__(self, scope):
            self.scope, self.net = scope, Net()

    tf.reset_default_graph()
    m = Model("contextmodel/")
    tf._models.append(m)
    want = m.net.state_dict()
    path = str(tmp_path / "100" / "model_100_1.00")
    tf.train.Saver(format="tf").save(None, path)
    assert T.checkpoint_exists(path) and sorted(T.list_variables(path)) == sorted("contextmodel/" + k for k in want)
    # a TF-trained checkpoint also carries optimiser slots: add some and rewrite
    full = T.read_checkpoint(path)
    full["contextmodel/conv/h0_conv/w/Adam"] = np.zeros((5, 5, 3, 4), np.float32)
    full["beta1_power"] = np.array(0.9, np.float32)
    T.write_checkpoint(path, full)
    for scope in ("contextmodel/", ""):                    # same scope as saved, and the sampler's un-scoped build
        tf.reset_default_graph()
        m2 = Model(scope)
        m2.net.P = {k: np.zeros(s, np.float32) for k, s in m2.net.shapes.items()}
        tf._models.append(m2)
        tf.train.Saver().restore(None, path)
        for k in want:
            assert np.array_equal(m2.net.P[k], want[k]), (scope, k)
    # the default format stays the .npz, and restore still reads it
    tf.reset_default_graph()
    tf._models.append(m)
    p2 = str(tmp_path / "npz" / "model")
    tf.train.Saver().save(None, p2)
    assert not T.checkpoint_exists(p2)
    m.net.P = {k: np.zeros(s, np.float32) for k, s in m.net.shapes.items()}
    tf.train.Saver().restore(None, p2)
    assert all(np.array_equal(m.net.P[k], want[k]) for k in want)
    # a checkpoint of another model is rejected with the missing names
    tf.reset_default_graph()
    m3 = Model("contextmodel/")
    m3.net.shapes = dict(m3.net.shapes, **{"translate/trans_z/Matrix": (4, 4)})
    tf._models.append(m3)
    with pytest.raises(KeyError):
        tf.train.Saver().restore(None, path)
    tf.reset_default_graph()


def test_tf_checkpoint_v1_single_file(tmp_path):
    """The checkpoint the reference restores for InceptionV3 (slim's `inception_v3.ckpt`, base.py:130-132) is a single-file
    V1 checkpoint: SavedTensorSlices protos in one table.  Reader vs our writer (float_val and tensor_content forms), a
    hand-assembled variable stored in two slices, and the one restore path every caller uses (load_variables)."""
    import numpy as np
    import pytest
    from imitation_from_observation_b200 import tf_checkpoint as T

    rs = np.random.RandomState(3)
    sd = {"InceptionV3/Conv2d_1a_3x3/weights": rs.randn(3, 3, 3, 32).astype(np.float32),
          "InceptionV3/Conv2d_1a_3x3/BatchNorm/beta": rs.randn(32).astype(np.float32),
          "InceptionV3/Logits/Conv2d_1c_1x1/biases": rs.randn(1001).astype(np.float32),
          "global_step": np.asarray(7, dtype=np.int64), "f64": rs.randn(4, 5)}
    for content in (False, True):
        path = str(tmp_path / ("inception_v3_%d.ckpt" % content))
        T.write_checkpoint_v1(path, sd, use_tensor_content=content)
        assert T.checkpoint_exists(path) and not T._is_v2(path)
        lv = T.list_variables(path)
        assert lv["InceptionV3/Conv2d_1a_3x3/weights"] == (np.dtype(np.float32), (3, 3, 3, 32)) and lv["global_step"][1] == ()
        got = T.read_checkpoint(path)
        assert sorted(got) == sorted(sd)
        for k in sd:
            assert got[k].dtype == sd[k].dtype and np.array_equal(got[k], sd[k]), k
        assert list(T.read_checkpoint(path, ["f64"])) == ["f64"]
        with pytest.raises(KeyError):
            T.read_checkpoint(path, ["nope"])

    # a variable stored as two row slices ([0:2, :] and [2:5, :]) — what a partitioned save produces
    def ld(num, payload):
        return T._put_varint(num << 3 | 2) + T._put_varint(len(payload)) + payload

    def vi(num, v):
        return T._put_varint(num << 3) + T._put_varint(v)

    full = rs.randn(5, 4).astype(np.float32)
    shape = ld(2, vi(1, 5)) + ld(2, vi(1, 4))
    meta = ld(1, ld(1, b"w") + ld(2, shape) + vi(3, 1))
    entries = [(b"", ld(1, meta))]
    for i, (r0, n) in enumerate(((0, 2), (2, 3))):
        sl = ld(1, vi(1, r0) + vi(2, n)) + ld(1, b"")                 # extent rows [r0, r0+n), full columns
        tensor = vi(1, 1) + ld(5, full[r0:r0 + n].tobytes())
        entries.append((b"\x00w\x00\x01" + bytes([i]), ld(2, ld(1, b"w") + ld(2, sl) + ld(3, tensor))))
    T.write_table(str(tmp_path / "sliced.ckpt"), entries)
    assert np.array_equal(T.read_checkpoint(str(tmp_path / "sliced.ckpt"))["w"], full)
    T.write_table(str(tmp_path / "half.ckpt"), entries[:2])            # a missing slice is an error, not zeros
    with pytest.raises(T.CheckpointError):
        T.read_checkpoint(str(tmp_path / "half.ckpt"))

    # load_variables: V1 file, V2 bundle or .npz, under the first prefix that holds every name
    names = ["Conv2d_1a_3x3/weights", "Conv2d_1a_3x3/BatchNorm/beta"]
    v1 = str(tmp_path / "inception_v3_0.ckpt")
    vals, pre = T.load_variables(v1, names, prefixes=("", "InceptionV3/"))
    assert pre == "InceptionV3/" and np.array_equal(vals["InceptionV3/" + names[0]], sd["InceptionV3/" + names[0]])
    np.savez(str(tmp_path / "m.npz"), **{"contextmodel/a": np.ones(3, np.float32)})
    vals, pre = T.load_variables(str(tmp_path / "m"), ["a"], prefixes=("x/", "contextmodel/", ""))
    assert pre == "contextmodel/" and np.array_equal(vals["contextmodel/a"], np.ones(3, np.float32))
    with pytest.raises(KeyError):
        T.load_variables(str(tmp_path / "m"), ["b"], prefixes=("", "contextmodel/"))
    with pytest.raises(FileNotFoundError):
        T.load_variables(str(tmp_path / "absent"), ["a"])


def test_videos_to_bank_reads_mp4_like_the_reference(tmp_path):
    """train_script.py:58-96 on real mp4 files: 51-frame videos only, frames 1, 1 + nskip, ..., transform, black-frame drop."""
    cv2 = pytest.importorskip("cv2")
    from imitation_from_observation_b200 import data as D
    rs = np.random.RandomState(3)

    def write(name, nframes, black_at=None):
        path = str(tmp_path / name)
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (64, 48))
        if not vw.isOpened():
            pytest.skip("no mp4 encoder in this OpenCV build")
        for j in range(nframes):
            base = np.full((48, 64, 3), 40 + 3 * j, np.uint8)
            base[:, :, 1] = rs.randint(60, 200)
            vw.write(np.zeros((48, 64, 3), np.uint8) if j == black_at else base)
        vw.release()
        return path

    good = [write("a.mp4", 51), write("b.mp4", 51)]
    short = write("c.mp4", 30)
    black = write("d.mp4", 51, black_at=11)                     # frame 11 = 1 + 2 * 5 is sampled at nskip = 5
    frames = D.read_video(good[0])
    assert len(frames) == 51 and frames[0].shape == (48, 64, 3) and frames[0].dtype == np.uint8
    assert abs(int(frames[7][0, 0, 0]) - (40 + 21)) <= 3        # RGB order: channel 0 holds the ramp (lossy codec)
    bank = D.videos_to_bank(good + [short, black, str(tmp_path / "missing.mp4")], nlen=10, nskip=5, idims=(24, 32),
                            rng=np.random.RandomState(0))
    assert bank.shape == (10, 2, 24, 32, 3) and bank.dtype == np.float64
    assert bank.min() >= -1.0 and bank.max() <= 1.0
    # time order: the ramp brightens by 3 * nskip grey levels per kept frame
    ramp = (bank[:, 0, 0, 0, 0] + 1.0) * 127.5
    assert np.all(np.abs(np.diff(ramp) - 15.0) <= 4.0)
    # uint8 path of the inception trainer: no rescale, black frames are kept (train_script.py:73 `not self.inception`)
    bank8 = D.videos_to_bank(good + [black], nlen=10, nskip=5, idims=(24, 32), rescale=False, inception=True,
                             rng=np.random.RandomState(0))
    assert bank8.shape == (10, 3, 24, 32, 3) and bank8.dtype == np.uint8
