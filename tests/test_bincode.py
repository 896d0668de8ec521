"""bincode 2 (standard configuration) persistence of the mixture: the encoding rules against the known answers of the
bincode specification, the exact round trip through the serde structure of doc/Gpx_Tutorial.ipynb:421, and the layout of
a few hand-checked fragments.  No GPU (the device round trip is in tests/test_gpu_fit_api.py)."""
import json
import os
import struct

import pytest

from egobox_b200 import bincode as B

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _enc(t, v):
    out = bytearray()
    B._encode(out, t, v)
    return bytes(out)


def test_varint_known_answers():
    """bincode spec, "VarintEncoding": u < 251 one byte; 251 + u16 LE; 252 + u32 LE; 253 + u64 LE."""
    cases = {0: b"\x00", 250: b"\xfa", 251: b"\xfb\xfb\x00", 65535: b"\xfb\xff\xff", 65536: b"\xfc\x00\x00\x01\x00",
             2 ** 32 - 1: b"\xfc\xff\xff\xff\xff", 2 ** 32: b"\xfd\x00\x00\x00\x00\x01\x00\x00\x00",
             10847399533071556633: b"\xfd" + struct.pack("<Q", 10847399533071556633)}
    for v, want in cases.items():
        assert _enc(B.USIZE, v) == want
        assert B._decode(B.Reader(want), B.USIZE) == v


def test_primitive_layouts():
    assert _enc(B.F64, 1.0) == struct.pack("<d", 1.0)
    assert _enc(B.STR, "Matern52") == b"\x08Matern52"
    assert _enc(B.option(B.USIZE), None) == b"\x00" and _enc(B.option(B.USIZE), 3) == b"\x01\x03"
    assert _enc(B.RECOMBINATION, "Hard") == b"\x00"
    assert _enc(B.RECOMBINATION, {"Smooth": None}) == b"\x01\x00"
    assert _enc(B.RECOMBINATION, {"Smooth": 0.5}) == b"\x01\x01" + struct.pack("<d", 0.5)
    assert _enc(B.NB_CLUSTERS, {"Fixed": {"nb": 3}}) == b"\x00\x03"
    assert _enc(B.NB_CLUSTERS, {"Auto": {"max": None}}) == b"\x01\x00"
    # ndarray: v, dim (no length: a fixed-size array), data (sequence)
    assert _enc(B.array(2), {"v": 1, "dim": [2, 1], "data": [1.0, 2.0]}) == b"\x01\x02\x01\x02" + struct.pack("<2d", 1.0, 2.0)
    # Array1<(f64, f64)>: tuples carry no length
    assert _enc(B.array(1, B.PAIR), {"v": 1, "dim": [1], "data": [[0.01, 10.0]]}) == b"\x01\x01\x01" + struct.pack("<2d", 0.01, 10.0)
    # bitflags: the bits
    assert _enc(B.MIXTURE_PARAMS[1][3][1], "CONSTANT | LINEAR | QUADRATIC") == b"\x07"
    assert _enc(B.MIXTURE_PARAMS[1][4][1], "SQUAREDEXPONENTIAL | MATERN52") == b"\x09"
    assert B._decode(B.Reader(b"\x09"), B.MIXTURE_PARAMS[1][4][1]) == "SQUAREDEXPONENTIAL | MATERN52"
    # ThetaTuning::Full: variant 1, init, bounds
    tt = {"Full": {"init": {"v": 1, "dim": [1], "data": [0.1]}, "bounds": {"v": 1, "dim": [1], "data": [[0.01, 10.0]]}}}
    assert _enc(B.THETA_TUNING, tt) == (b"\x01" + b"\x01\x01\x01" + struct.pack("<d", 0.1) + b"\x01\x01\x01"
                                        + struct.pack("<2d", 0.01, 10.0))


def _notebook_model():
    """The complete serde JSON of a trained GpMixture stored in doc/Gpx_Tutorial.ipynb:421, re-assembled from the two
    transcribed fixtures (the expert and the mixture-level blocks), in the reference's key order."""
    expert = json.load(open(os.path.join(GOLDEN, "gpx_tutorial_linear_matern52.json")))
    mix = json.load(open(os.path.join(GOLDEN, "gpx_tutorial_mixture.json")))
    expert = {k: v for k, v in expert.items() if k != "source"}
    return {"recombination": mix["recombination"], "experts": [expert], "gmx": mix["gmx"], "gp_type": mix["gp_type"],
            "training_data": mix["training_data"], "params": mix["params"]}


def test_reference_model_round_trips_exactly():
    obj = _notebook_model()
    data = B.encode_mixture(obj)
    back = B.decode_mixture(data)
    assert back == obj                                  # every float bit for bit, every key, every order
    assert B.encode_mixture(back) == data
    # layout spot checks: recombination Hard, one expert, its typetag name first
    name = obj["experts"][0]["type_fullgp"].encode()
    assert data[:2] == b"\x00\x01" and data[2] == len(name) and data[3:3 + len(name)] == name
    # the rng state (4 x u64, here all > 2^32: 9 bytes each) closes the file
    s = obj["params"]["rng"]["s"]
    tail = b"".join(b"\xfd" + struct.pack("<Q", v) for v in s)
    assert data.endswith(tail)
    with pytest.raises(ValueError):
        B.decode_mixture(data + b"\x00")
    with pytest.raises(ValueError):
        B.decode_mixture(data[:-3])
