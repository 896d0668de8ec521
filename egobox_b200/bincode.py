"""bincode 2 ("standard" configuration) encoding of the reference's serde data model of a trained `GpMixture`.

The reference writes `.bin` files with `bincode::serde::encode_to_vec(self, bincode::config::standard())`
(crates/moe/src/algorithm.rs:511-524) and reads them back with `decode_from_slice` (:1096-1106); `Gpx.save` /
`Gpx.load` choose it for every file name that does not end in `.json` (python/src/gp_mix.rs:310-337).  bincode is
not self-describing: the bytes are the serde calls of the derived `Serialize` impls in field order, so this module
walks the SAME structure that `Gpx.to_dict()` builds for the JSON format, driven by a schema restating the Rust types:
  GpMixture (moe/src/algorithm.rs:429-443), Recombination / RegressionSpec / CorrelationSpec (moe/src/types.rs:17-93),
  NbClusters / GpType / GpMixtureValidParams (moe/src/parameters.rs:22-135), GaussianMixture (moe/src/gaussian_mixture.rs:28-45),
  GaussianProcess / GpInnerParams (gp/src/algorithm.rs:41-60, 165-192), NormalizedData (gp/src/utils.rs:8-17),
  ThetaTuning / GpValidParams (gp/src/parameters.rs:12-105), mean / correlation models as strings (`serde(into = "String")`).
bincode 2 standard configuration: little endian; u8 / bool / f64 raw; every other integer as a varint (< 251: one byte;
251 + u16, 252 + u32, 253 + u64); usize as u64; enum variant = u32 varint index; Option = one byte tag; sequences, strings
and maps carry a usize length; tuples, fixed-size arrays and struct fields carry nothing.  ndarray's serde form is the struct
{v: u8 = 1, dim: [usize; N], data: sequence}.  bitflags 2 serialises its bits as the integer for non-human-readable formats.
A `Box<dyn FullGpSurrogate>` goes through typetag's internally tagged form: the concrete struct with ONE extra leading
field, the variant name as a string (typetag::ser::TaggedSerializer::serialize_struct).

NOT VERIFIED against a file written by stock egobox: there is no Rust toolchain in this image and the reference holds no
`.bin` fixture, so what is tested is the encoding rules above (known answers of the bincode specification), the exact
round trip dict -> bytes -> dict, and save / load of models through this format (tests/test_bincode.py).
"""
from __future__ import annotations

import struct

# ---- schema ------------------------------------------------------------------------------------------------
F64, USIZE, U8, U64, STR = "f64", "usize", "u8", "u64", "str"


def option(t):
    return ("option", t)


def seq(t):
    return ("seq", t)


def array(ndim, elem=F64):
    return ("ndarray", ndim, elem)


def record(*fields):
    return ("struct", list(fields))


def enum(*variants):
    """variants: (name, None) unit | (name, type) newtype | (name, ("struct", ...)) struct variant."""
    return ("enum", list(variants))


def flags(*names):
    return ("flags", list(names))


PAIR = ("tuple", [F64, F64])
THETA_TUNING = enum(("Fixed", array(1)),
                    ("Full", record(("init", array(1)), ("bounds", array(1, PAIR)))),
                    ("Partial", record(("init", array(1)), ("bounds", array(1, PAIR)), ("active", seq(USIZE)))))
RECOMBINATION = enum(("Hard", None), ("Smooth", option(F64)))
NORMALIZED = record(("data", array(2)), ("mean", array(1)), ("std", array(1)))
GP_VALID_PARAMS = record(("theta_tuning", THETA_TUNING), ("mean", STR), ("corr", STR), ("kpls_dim", option(USIZE)),
                         ("n_start", USIZE), ("max_eval", USIZE), ("nugget", F64))
INNER_PARAMS = record(("sigma2", F64), ("beta", array(2)), ("gamma", array(2)), ("r_chol", array(2)), ("ft", array(2)),
                      ("ft_qr_r", array(2)))
GAUSSIAN_PROCESS = record(("theta", array(1)), ("likelihood", F64), ("inner_params", INNER_PARAMS), ("w_star", array(2)),
                          ("xt_norm", NORMALIZED), ("yt_norm", NORMALIZED),
                          ("training_data", ("tuple", [array(2), array(1)])), ("params", GP_VALID_PARAMS))
EXPERT = ("typetag", "type_fullgp", GAUSSIAN_PROCESS)
GAUSSIAN_MIXTURE = record(("weights", array(1)), ("means", array(2)), ("covariances", array(3)), ("precisions", array(3)),
                          ("precisions_chol", array(3)), ("heaviside_factor", F64), ("log_det", array(1)))
GP_TYPE = enum(("FullGp", None), ("SparseGp", "unsupported"))
NB_CLUSTERS = enum(("Fixed", record(("nb", USIZE))), ("Auto", record(("max", option(USIZE)))))
MIXTURE_PARAMS = record(("gp_type", GP_TYPE), ("n_clusters", NB_CLUSTERS), ("recombination", RECOMBINATION),
                        ("regression_spec", flags("CONSTANT", "LINEAR", "QUADRATIC")),
                        ("correlation_spec", flags("SQUAREDEXPONENTIAL", "ABSOLUTEEXPONENTIAL", "MATERN32", "MATERN52")),
                        ("theta_tunings", seq(THETA_TUNING)), ("kpls_dim", option(USIZE)), ("n_start", USIZE),
                        ("max_eval", USIZE), ("gmm", option("unsupported")), ("gmx", option(GAUSSIAN_MIXTURE)),
                        ("rng", record(("s", ("fixed", 4, U64)))))
GP_MIXTURE = record(("recombination", RECOMBINATION), ("experts", seq(EXPERT)), ("gmx", GAUSSIAN_MIXTURE),
                    ("gp_type", GP_TYPE), ("training_data", ("tuple", [array(2), array(1)])), ("params", MIXTURE_PARAMS))


# ---- primitives --------------------------------------------------------------------------------------------
def put_varint(out, v):
    v = int(v)
    if v < 0:
        raise ValueError("unsigned varint expected")
    if v < 251:
        out.append(v)
    elif v < 1 << 16:
        out.append(251)
        out += struct.pack("<H", v)
    elif v < 1 << 32:
        out.append(252)
        out += struct.pack("<I", v)
    elif v < 1 << 64:
        out.append(253)
        out += struct.pack("<Q", v)
    else:
        raise ValueError("integer too wide")


class Reader:
    def __init__(self, data):
        self.b, self.i = memoryview(data), 0

    def take(self, n):
        if self.i + n > len(self.b):
            raise ValueError("bincode: unexpected end of data")
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def varint(self):
        t = self.take(1)[0]
        if t < 251:
            return t
        if t == 251:
            return struct.unpack("<H", self.take(2))[0]
        if t == 252:
            return struct.unpack("<I", self.take(4))[0]
        if t == 253:
            return struct.unpack("<Q", self.take(8))[0]
        raise ValueError("bincode: unsupported varint tag %d" % t)


# ---- encode / decode ------------------------------------------------------------------------------------------
def _encode(out, t, v):
    if t == F64:
        out += struct.pack("<d", float(v))
    elif t in (USIZE, U64):
        put_varint(out, v)
    elif t == U8:
        out.append(int(v) & 0xFF)
    elif t == STR:
        raw = str(v).encode("utf-8")
        put_varint(out, len(raw))
        out += raw
    elif t == "unsupported":
        raise NotImplementedError("bincode: this part of the model (sparse GP / gmm) is not covered")
    else:
        kind = t[0]
        if kind == "option":
            if v is None:
                out.append(0)
            else:
                out.append(1)
                _encode(out, t[1], v)
        elif kind == "seq":
            put_varint(out, len(v))
            for e in v:
                _encode(out, t[1], e)
        elif kind == "fixed":
            assert len(v) == t[1]
            for e in v:
                _encode(out, t[2], e)
        elif kind == "tuple":
            assert len(v) == len(t[1])
            for tt, e in zip(t[1], v):
                _encode(out, tt, e)
        elif kind == "struct":
            for name, tt in t[1]:
                _encode(out, tt, v[name])
        elif kind == "ndarray":
            ndim, elem = t[1], t[2]
            out.append(int(v.get("v", 1)))
            dim = list(v["dim"])
            assert len(dim) == ndim, (dim, ndim)
            for d in dim:
                put_varint(out, d)
            data = v["data"]
            put_varint(out, len(data))
            if elem == F64:
                out += struct.pack("<%dd" % len(data), *[float(x) for x in data])
            else:
                for e in data:
                    _encode(out, elem, e)
        elif kind == "enum":
            names = [n for n, _ in t[1]]
            if isinstance(v, str):
                idx, payload = names.index(v), None
            else:
                (name, payload), = v.items()
                idx = names.index(name)
            put_varint(out, idx)
            pt = t[1][idx][1]
            if pt is not None:
                _encode(out, pt, payload)
        elif kind == "flags":
            bits = 0
            for part in str(v).split("|"):
                part = part.strip()
                if part:
                    bits |= 1 << t[1].index(part)
            out.append(bits)
        elif kind == "typetag":
            _encode(out, STR, v[t[1]])
            _encode(out, t[2], v)
        else:
            raise ValueError("bad schema %r" % (t,))


def _decode(r, t):
    if t == F64:
        return struct.unpack("<d", r.take(8))[0]
    if t in (USIZE, U64):
        return r.varint()
    if t == U8:
        return r.take(1)[0]
    if t == STR:
        return bytes(r.take(r.varint())).decode("utf-8")
    if t == "unsupported":
        raise NotImplementedError("bincode: this part of the model (sparse GP / gmm) is not covered")
    kind = t[0]
    if kind == "option":
        tag = r.take(1)[0]
        if tag > 1:
            raise ValueError("bincode: bad Option tag %d" % tag)
        return _decode(r, t[1]) if tag else None
    if kind == "seq":
        return [_decode(r, t[1]) for _ in range(r.varint())]
    if kind == "fixed":
        return [_decode(r, t[2]) for _ in range(t[1])]
    if kind == "tuple":
        return [_decode(r, tt) for tt in t[1]]
    if kind == "struct":
        return {name: _decode(r, tt) for name, tt in t[1]}
    if kind == "ndarray":
        ndim, elem = t[1], t[2]
        v = r.take(1)[0]
        dim = [r.varint() for _ in range(ndim)]
        n = r.varint()
        if elem == F64:
            data = list(struct.unpack("<%dd" % n, r.take(8 * n)))
        else:
            data = [_decode(r, elem) for _ in range(n)]
        return {"v": v, "dim": dim, "data": data}
    if kind == "enum":
        idx = r.varint()
        if idx >= len(t[1]):
            raise ValueError("bincode: enum variant %d out of range" % idx)
        name, pt = t[1][idx]
        return name if pt is None else {name: _decode(r, pt)}
    if kind == "flags":
        bits = r.take(1)[0]
        return " | ".join(n for i, n in enumerate(t[1]) if bits >> i & 1)
    if kind == "typetag":
        name = _decode(r, STR)
        body = _decode(r, t[2])
        out = {t[1]: name}
        out.update(body)
        return out
    raise ValueError("bad schema %r" % (t,))


def encode_mixture(obj) -> bytes:
    """`Gpx.to_dict()` -> the bytes of `bincode::serde::encode_to_vec(&GpMixture, standard())`."""
    out = bytearray()
    _encode(out, GP_MIXTURE, obj)
    return bytes(out)


def decode_mixture(data):
    r = Reader(data)
    obj = _decode(r, GP_MIXTURE)
    if r.i != len(r.b):
        raise ValueError("bincode: %d trailing bytes" % (len(r.b) - r.i))
    return obj
