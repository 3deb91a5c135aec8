import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    p = os.path.join(GOLDEN, name)
    if os.path.exists(p + ".gz"):
        with gzip.open(p + ".gz", "rt") as f:
            return json.load(f)
    with open(p) as f:
        return json.load(f)


def f32bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def case_id(c):
    if c["mode"] == "puct":
        return "puct-%s-%d-s%d" % (c["game"], c["sims"], c["salt"])
    return "gumbel-%s-n%d-m%d-%s-%s-s%d" % (c["game"], c["n"], c["m"], c["activation"],
                                            "reuse" if c["reuse"] else "fresh", c["salt"])
