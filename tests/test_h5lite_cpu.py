"""h5lite (the dependency-free HDF5 reader / writer behind the Keras weight files and the
prediction / attribution outputs): round trips, the structures the HDF5 specification fixes
byte for byte, and a file written by the real libhdf5 (MATLAB 7.4 ``-v7.3`` output shipped with
scipy's test data) as an external pin of the reader."""
import os
import struct

import numpy as np
import pytest

from bpreveal_b200 import h5lite


def test_round_trip_of_everything_the_writer_can_write(tmp_path):
    rng = np.random.default_rng(0)
    w = h5lite.Writer()
    k = rng.standard_normal((25, 4, 64)).astype(np.float32)
    w.create_dataset("layers/conv1d/vars/0", data=k)
    w.create_dataset("layers/conv1d/vars/1", data=np.arange(64, dtype=np.float32))
    shap = rng.standard_normal((300, 50, 4)).astype(np.float16)
    w.create_dataset("shap", data=shap, chunks=(128, 50, 4), compression="gzip")
    ints = np.arange(1000, dtype=np.int64).reshape(10, 100)
    w.create_dataset("ints", data=ints, chunks=(3, 40), shuffle=True, compression=6)
    w.create_dataset("plainchunks", data=ints.astype(np.uint8), chunks=(4, 100))
    w.create_dataset("names", data=np.array([b"chr1", b"chrX_long"]))
    w.create_dataset("descriptions", data=["first window", "second \u00e9", ""])
    w.attrs("")["big"] = "x" * 200000   # beyond one object-header message: global heap
    w.attrs("")["tags"] = ["a", "bc"]
    w.create_dataset("scalar", data=np.float64(3.5))
    w.create_dataset("empty", data=np.zeros((0, 4), dtype=np.float32))
    w.attrs("")["keras_version"] = "3.5.0"
    w.attrs("layers")["layer_names"] = [b"a", b"bcd"]
    w.attrs("shap")["scale"] = np.float32(2.5)
    w.attrs("shap")["dims"] = np.array([1, 2, 3], dtype=np.int32)
    for i in range(700):  # more members than one symbol node of the default size holds
        w.create_dataset(f"many/d{i:03d}", data=np.full((3,), i, dtype=np.int32))
    w.create_dataset("late", shape=(5, 7), dtype=np.float32)
    path = str(tmp_path / "t.h5")
    late = w.save(path)
    late["/late"][...] = np.arange(35, dtype=np.float32).reshape(5, 7)
    late["/late"].flush()
    with h5lite.File(path) as f:
        assert sorted(f.keys()) == ["descriptions", "empty", "ints", "late", "layers", "many",
                                    "names", "plainchunks", "scalar", "shap"]
        assert list(f["descriptions"][...]) == ["first window", "second \u00e9", ""]
        assert f.attrs["big"] == "x" * 200000 and list(f.attrs["tags"]) == ["a", "bc"]
        assert np.array_equal(f["layers/conv1d/vars/0"][...], k)
        assert f["layers/conv1d/vars/0"].dtype == np.float32
        assert np.array_equal(f["shap"][...], shap) and f["shap"].dtype == np.float16
        assert np.array_equal(f["ints"][...], ints)
        assert np.array_equal(f["plainchunks"][...], ints.astype(np.uint8))
        assert list(f["names"][...]) == [b"chr1", b"chrX_long"]
        assert f["scalar"][...] == 3.5 and f["scalar"].shape == ()
        assert f["empty"].shape == (0, 4)
        assert f.attrs["keras_version"] == "3.5.0"  # str -> variable-length UTF-8, like h5py
        assert list(f["layers"].attrs["layer_names"]) == [b"a", b"bcd"]
        assert f["shap"].attrs["scale"] == np.float32(2.5)
        assert list(f["shap"].attrs["dims"]) == [1, 2, 3]
        assert len(f["many"]) == 700 and all(f[f"many/d{i:03d}"][0] == i for i in (0, 9, 350, 699))
        assert np.array_equal(f["late"][...], np.arange(35).reshape(5, 7))
        assert "layers/conv1d" in f and "layers/nothing" not in f
        seen = []
        f["layers"].visititems(lambda n, o: seen.append(n))
        assert seen == ["conv1d", "conv1d/vars", "conv1d/vars/0", "conv1d/vars/1"]
    # the same through bytes
    assert np.array_equal(h5lite.File(w.tobytes())["ints"][...], ints)


def test_many_chunks_need_a_two_level_chunk_tree(tmp_path):
    a = np.arange(200 * 6, dtype=np.float32).reshape(200, 6)
    w = h5lite.Writer()
    w.create_dataset("x", data=a, chunks=(2, 6), compression="gzip")  # 100 chunks > 64 per node
    f = h5lite.File(w.tobytes())
    assert np.array_equal(f["x"][...], a)
    assert np.array_equal(f["x"][10:12], a[10:12])


def test_written_bytes_follow_the_specification():
    """Known-answer checks of the fixed parts of the format (HDF5 File Format Specification 3.0,
    sections II.A superblock v0, III.A.1 B-tree v1 / symbol nodes, III.D local heap,
    IV.A.1 object header v1, IV.A.2.d datatype message)."""
    w = h5lite.Writer()
    w.create_dataset("a", data=np.array([1.0, 2.0], dtype=np.float32))
    raw = w.tobytes()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    # superblock v0: versions, offset / length sizes 8 / 8, group leaf K 4, internal K 16
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    assert struct.unpack_from("<HH", raw, 16) == (4, 16)
    base, freesp, eof, driver = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and freesp == h5lite.UNDEF and driver == h5lite.UNDEF and eof == len(raw)
    name_off, root_oh, cache, _ = struct.unpack_from("<QQII", raw, 56)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert cache == 1 and name_off == 0
    # root object header: version 1, one symbol-table message (type 0x11) naming the same nodes
    ver, _, nmsg, refc, hsize = struct.unpack_from("<BBHII", raw, root_oh)
    assert (ver, nmsg, refc) == (1, 1, 1)
    mtype, msize = struct.unpack_from("<HH", raw, root_oh + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", raw, root_oh + 24) == (btree, heap)
    assert raw[btree:btree + 4] == b"TREE" and raw[btree + 4] == 0 and raw[btree + 5] == 0
    assert raw[heap:heap + 4] == b"HEAP"
    dsize, free_head, dseg = struct.unpack_from("<QQQ", raw, heap + 8)
    assert free_head == 1  # H5HL_FREE_NULL: libhdf5 rejects any other "no free block" marker
    assert raw[dseg:dseg + 8] == bytes(8)           # offset 0 is the empty string
    assert raw[dseg + 8:dseg + 10] == b"a\0"
    snod = struct.unpack_from("<Q", raw, btree + 24 + 8)[0]
    assert raw[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", raw, snod + 6)[0] == 1
    # IEEE binary32 little-endian datatype message as the specification's example spells it
    dt = h5lite._encode_dtype(np.dtype("<f4"))
    assert dt == bytes([0x11, 0x20, 0x1F, 0x00, 4, 0, 0, 0, 0, 0, 32, 0, 23, 8, 0, 23, 127, 0, 0, 0])
    dt16 = h5lite._encode_dtype(np.dtype("<f2"))
    assert dt16 == bytes([0x11, 0x20, 0x0F, 0x00, 2, 0, 0, 0, 0, 0, 16, 0, 10, 5, 0, 10, 15, 0, 0, 0])
    di = h5lite._encode_dtype(np.dtype("<i4"))
    assert di == bytes([0x10, 0x08, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0])
    for d in ("<f2", "<f4", "<f8", "<i1", "<u1", "<i8", ">i4", ">f8", "S7"):
        assert h5lite._decode_dtype(h5lite._encode_dtype(np.dtype(d))).np_dtype == np.dtype(d)


def _v2_file():
    """A hand-assembled file in the NEW style (superblock v2, version-2 object headers with link
    messages, version-3 attribute with a variable-length string in a global heap): what h5py
    writes with libver='latest' for small groups.  Built here independently of h5lite.Writer."""
    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(b), 0) + b for t, b in msgs)
        head = b"OHDR" + bytes([2, 0x00]) + struct.pack("<B", len(body))
        return head + body + b"\0\0\0\0"  # checksum is not verified by readers of this kind

    data = np.arange(6, dtype="<i2").tobytes()
    blob = bytearray(48)
    gheap_addr = 48
    s = b"hello"
    gcol = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", 64) + \
        struct.pack("<HHIQ", 1, 0, 0, len(s)) + s.ljust(8, b"\0")
    blob += gcol.ljust(64, b"\0")
    data_addr = len(blob)
    blob += data
    blob += b"\0" * (-len(blob) % 8)
    dt = bytes([0x10, 0x08, 0, 0, 2, 0, 0, 0, 0, 0, 16, 0])
    sp = bytes([2, 2, 0, 1]) + struct.pack("<QQ", 2, 3)
    vl = bytes([0x19, 0x01, 0x01, 0x00]) + struct.pack("<I", 16) + \
        bytes([0x13, 0x10, 0, 0, 1, 0, 0, 0])
    asp = bytes([2, 0, 0, 0])
    aname = b"note\0"
    attr = bytes([3, 0]) + struct.pack("<HHH", len(aname), len(vl), len(asp)) + bytes([1]) + aname + \
        vl + asp + struct.pack("<IQI", len(s), gheap_addr, 1)
    ds_addr = len(blob)
    blob += ohdr([(0x01, sp), (0x03, dt), (0x08, bytes([3, 1]) + struct.pack("<QQ", data_addr, len(data))),
                  (0x0C, attr)])
    blob += b"\0" * (-len(blob) % 8)
    link = bytes([1, 0x00, 1]) + b"d" + struct.pack("<Q", ds_addr)
    grp_addr = len(blob)
    blob += ohdr([(0x06, link)])
    blob += b"\0" * (-len(blob) % 8)
    link2 = bytes([1, 0x00, 1]) + b"g" + struct.pack("<Q", grp_addr)
    root_addr = len(blob)
    blob += ohdr([(0x06, link2)])
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, h5lite.UNDEF, len(blob),
                                                                  root_addr) + b"\0\0\0\0"
    blob[:len(sb)] = sb
    return bytes(blob)


def test_reader_handles_new_style_files():
    f = h5lite.File(_v2_file())
    assert f.keys() == ["g"] and f["g"].keys() == ["d"]
    d = f["g/d"]
    assert d.shape == (2, 3) and d.dtype == np.dtype("<i2")
    assert np.array_equal(d[...], np.arange(6).reshape(2, 3))
    assert d.attrs["note"] == "hello"


def test_unsupported_features_are_named(tmp_path):
    raw = bytearray(_v2_file())
    raw[8] = 7  # superblock version
    with pytest.raises(h5lite.H5Unsupported, match="superblock version 7"):
        h5lite.File(bytes(raw))
    with pytest.raises(ValueError, match="signature"):
        h5lite.File(b"not an hdf5 file at all")
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.Writer().create_dataset("x", data=np.array([1 + 2j]))
    w = h5lite.Writer()
    w.create_dataset("a/b", data=np.zeros(3))
    with pytest.raises(ValueError, match="already exists"):
        w.create_dataset("a/b", data=np.zeros(3))


MATLAB = os.path.join(os.path.dirname(np.__file__), "..", "scipy", "io", "matlab", "tests", "data",
                      "testhdf5_7.4_GLNX86.mat")


@pytest.mark.skipif(not os.path.exists(MATLAB), reason="scipy's test data is not installed")
def test_reader_against_a_file_written_by_libhdf5():
    """MATLAB 7.4 wrote this with the HDF5 library itself: 512-byte user block, superblock v0,
    symbol-table root group, version-1 object header, layout message version 2, a fixed-length
    string attribute.  scipy's own test expects testdouble = 0, pi/4, ..., 2 pi."""
    f = h5lite.File(MATLAB)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.float64
    assert np.allclose(d[...].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    assert d.attrs["MATLAB_class"] == b"double"
