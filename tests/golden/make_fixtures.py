"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted read-only; the GPU box never needs /root/reference).

  python tests/golden/make_fixtures.py

1. half_cheetah_2k.npz : the first 2000 transitions of the reference's own expert-data fixture
   examples/il/expert_data/half_cheetah_mujoco.bson (a BSON.jl-serialised ExperienceBuffer):
   s, sp [2000,17] f32, a [2000,6] f32, r [2000] f32, done [2000] u8, t [2000] i64.  Used as real-data
   input for buffer / GAE / PPO parity tests (episode starts come from t == 1, experience_buffer.jl:198-200).
2. ref_kats.json : known-answer values transcribed from the reference's own tests (file:line cited per entry).
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# ---- minimal BSON reader (documents, arrays, strings, binary, ints, doubles, bools, null) --------------
def _cstr(b, o):
    e = b.index(b"\x00", o)
    return b[o:e].decode(), e + 1


def _doc(b, o, as_list=False):
    size = struct.unpack_from("<i", b, o)[0]
    end = o + size - 1
    o += 4
    out = [] if as_list else {}
    while o < end:
        t = b[o]; o += 1
        k, o = _cstr(b, o)
        if t == 0x01:
            v = struct.unpack_from("<d", b, o)[0]; o += 8
        elif t == 0x02:
            n = struct.unpack_from("<i", b, o)[0]; v = b[o + 4:o + 4 + n - 1].decode(); o += 4 + n
        elif t == 0x03:
            v, o = _doc(b, o)
        elif t == 0x04:
            v, o = _doc(b, o, True)
        elif t == 0x05:
            n = struct.unpack_from("<i", b, o)[0]; v = b[o + 5:o + 5 + n]; o += 5 + n
        elif t == 0x08:
            v = bool(b[o]); o += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v = struct.unpack_from("<i", b, o)[0]; o += 4
        elif t == 0x12:
            v = struct.unpack_from("<q", b, o)[0]; o += 8
        else:
            raise ValueError(f"BSON type {t:#x}")
        if as_list:
            out.append(v)
        else:
            out[k] = v
    return out, end + 1


_JL = {"Float32": np.float32, "Float64": np.float64, "Int64": np.int64, "Bool": np.uint8, "UInt8": np.uint8}


def _jl_array(node):
    """BSON.jl `array` node (tag=array, type.name[-1] in _JL, size, data) -> numpy [batch, features...]."""
    name = node["type"]["name"][-1]
    size = [int(x) for x in node["size"]]
    return np.frombuffer(node["data"], dtype=_JL[name]).reshape(size[::-1])  # column-major -> batch-major


def half_cheetah(n=2000):
    raw = open(os.path.join(REF, "examples/il/expert_data/half_cheetah_mujoco.bson"), "rb").read()
    doc, _ = _doc(raw, 0)
    cols = doc["data"]["data"][0]  # the ExperienceBuffer's `data` Dict (first struct field)
    out = {}
    for k in ("s", "a", "sp", "r", "done", "t"):
        a = _jl_array(cols[k])[:n]
        out[k] = np.ascontiguousarray(a.reshape(n) if a.shape[1:] == (1,) else a)
    return out


KATS = {
    "circ_inds": {"cite": "test/experience_buffer_tests.jl:23-28",
                  "cases": [[4, 60, 100], [1, 100, 100], [1, 101, 100], [1, 120, 100], [90, 20, 100]]},
    "split_batches": {"cite": "test/experience_buffer_tests.jl:177-180",
                      "cases": [[100, [0.5, 0.5], [50, 50]], [100, [1.0], [100]],
                                [100, [1 / 3, 1 / 3, 1 / 3], [34, 33, 33]]]},
    "last_n_partial": {"cite": "test/experience_buffer_tests.jl:32-41 (capacity 100, 50 pushed)",
                       "cases": [[10, [41, 50]], [1, [50, 50]], [50, [1, 50]], [51, [1, 50]], [1000, [1, 50]]]},
    "last_n_full": {"cite": "test/experience_buffer_tests.jl:44-51 (capacity 100, 150 pushed -> next_ind 51)"},
    "priorities": {"cite": "test/experience_buffer_tests.jl:193-205",
                   "I": [1, 2, 3], "v": [1.0, 2.0, 3.0], "alpha": 0.6, "max_priority": 3.0},
    "schedule": {"cite": "test/util_tests.jl:55-86"},
    "whiten_space": {"cite": "test/spaces_tests.jl:34-40", "mu": 1.0, "sigma": 2.0, "x": 0.0, "y": -0.5},
    "gae_kat": {"cite": "test/gym/sampler_tests.jl:75-81 inputs (r=6, done=1, V=0, lambda=0.9, gamma=0.7, range 1:5); "
                        "values derived in SURVEY 8c (the reference test asserts nothing: parity unpinned)",
                "advantage": [14.60685920715332, 13.66168212890625, 12.161399841308594, 9.779999732971191, 6.0],
                "return": [16.638599395751953, 15.197999000549316, 13.139999389648438, 10.199999809265137, 6.0]},
    "gradient_penalty": {"cite": "test/extras_tests.jl:7-9 (out of scope; listed for completeness)"},
}


if __name__ == "__main__":
    hc = half_cheetah()
    np.savez_compressed(os.path.join(HERE, "half_cheetah_2k.npz"), **hc)
    print({k: (v.shape, str(v.dtype)) for k, v in hc.items()}, "episode starts:", np.flatnonzero(hc["t"] == 1)[:5])
    with open(os.path.join(HERE, "ref_kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
